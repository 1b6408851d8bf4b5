"""A deterministic stand-in environment + recording controller shared by make_reference_golden.py (which drives the
REFERENCE's EvaluationService with it) and tests/test_reference_golden.py (which drives this repo's)."""
import numpy as np


class FakeEnv:
    """Mimics the part of REALRobotEnv that EvaluationService touches (evaluate.py:148-177, 203-394)."""
    log = None                       # shared list, set by the driver

    def __init__(self):
        self.action_space, self.observation_space = "ACTION_SPACE", "OBSERVATION_SPACE"
        self.intrinsic_timesteps = self.extrinsic_timesteps = self.extrinsic_trials = None
        self.timestep, self.goal_idx = 0, -1

    def set_goals_dataset_path(self, path):
        self.log.append("env.set_goals_dataset_path(%s)" % path)

    def render(self, mode):
        self.log.append("env.render(%s)" % mode)

    def reset(self):
        self.log.append("env.reset")
        self.timestep = 0
        return {"t": 0}

    def set_goal(self):
        self.goal_idx += 1
        self.log.append("env.set_goal -> %d" % self.goal_idx)
        return {"t": 0, "goal": self.goal_idx}

    def step(self, action):
        self.timestep += 1
        limit = self.intrinsic_timesteps if self.goal_idx < 0 else self.extrinsic_timesteps
        done = self.timestep >= limit
        self.log.append("env.step(%s) -> t=%d done=%s" % (action, self.timestep, done))
        return {"t": self.timestep}, 0, done, {}

    def evaluateGoal(self):
        ch = ['2D', '2.5D', '3D'][self.goal_idx % 3]
        score = 1.0 / (1 + self.goal_idx)
        self.log.append("env.evaluateGoal -> %s %.4f" % (ch, score))
        return ch, score


def make_controller(base_policy, log):
    class Recorder(base_policy):
        def __init__(self, action_space, observation_space):
            log.append("controller(%s, %s)" % (action_space, observation_space))
            self.n = 0

        def start_intrinsic_phase(self):
            log.append("controller.start_intrinsic_phase")

        def end_intrinsic_phase(self, observation, reward, done):
            log.append("controller.end_intrinsic_phase(%s, %s, %s)" % (observation, reward, done))

        def start_extrinsic_phase(self):
            log.append("controller.start_extrinsic_phase")

        def end_extrinsic_phase(self):
            log.append("controller.end_extrinsic_phase")

        def start_extrinsic_trial(self):
            log.append("controller.start_extrinsic_trial")

        def end_extrinsic_trial(self, observation, reward, done):
            log.append("controller.end_extrinsic_trial(%s, %s, %s)" % (observation, reward, done))

        def step(self, observation, reward, done):
            self.n += 1
            log.append("controller.step(%s, %s, %s)" % (observation, reward, done))
            return "a%d" % self.n
    return Recorder


def plain(x):
    """JSON-friendly copy of an evaluation_state / score object."""
    if isinstance(x, dict):
        return {str(k): plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [plain(v) for v in x]
    if isinstance(x, (np.floating, float)):
        return float(x)
    if isinstance(x, (np.integer, int)):
        return int(x)
    return x


# ---- scripted object motion for the settle detector (generate_goals.py:21-68 runEnv)
def scripted_pose(script, s):
    """Pose (xyz + quaternion) of the single object after s simulation steps."""
    p = np.array([0.0, 0.0, 0.4, 0.0, 0.0, 0.0, 1.0])
    if script == "A":                         # moving for 5 steps
        p[2] -= 0.01 * min(s, 5)
    elif script == "B":                       # moving for 40 steps
        p[2] -= 0.001 * min(s, 40)
    elif script == "C":                       # at rest, one 2e-4 jolt after step 25
        p[0] += 2e-4 if s >= 25 else 0.0
    elif script == "D":                       # q and -q alternate: the same orientation
        p[3:] *= -1.0 if s % 2 else 1.0
    elif script == "E":                       # creeping just below the threshold
        p[1] += 0.9e-4 * s
    elif script == "F":                       # creeping just above it
        p[1] += 1.1e-4 * s
    elif script == "G":                       # orientation drifting by 2e-3 per step for 30 steps
        a = 1e-3 * min(s, 30)
        p[3:] = [np.sin(a), 0.0, 0.0, np.cos(a)]
    return p


SCRIPTS = ["A", "B", "C", "D", "E", "F", "G"]
MAX_TS = [31, 32, 45, 46, 47, 61, 62, 200]
