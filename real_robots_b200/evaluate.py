"""real_robots.evaluate — evaluation service (reference real_robots/evaluate.py:16-446).

`evaluate(Controller, ...)` keeps the reference's signature, phase order, callbacks, checks and return value
`(score_object, scores)`.  The AIcrowd event sync (evaluate.py:123-146) is a no-op when `aicrowd_api` is absent.
`evaluate_vectorised` is the added trial-parallel extrinsic phase: every env of a VecREALRobotEnv runs one goal,
ranks shard the goal list and the per-goal scores are all-gathered with torch.distributed (NCCL on GPUs, gloo in
the CPU tests) — the only collective anywhere near this path.
"""
import numpy as np

from .policy import BasePolicy


class EvaluationService:
    def __init__(self, Controller, environment='R1', action_type='macro_action', n_objects=1,
                 intrinsic_timesteps=15e6, extrinsic_timesteps=10e3, extrinsic_trials=50, visualize=True,
                 goals_dataset_path="./goals.npy.npz", video=None):
        self.ControllerClass = Controller
        self.intrinsic_timesteps = intrinsic_timesteps
        self.extrinsic_timesteps = extrinsic_timesteps
        self.extrinsic_trials = extrinsic_trials
        self.visualize = visualize
        self.goals_dataset_path = goals_dataset_path
        self.setup_gym_env(environment, action_type, n_objects)
        self.setup_controller()
        self.setup_evaluation_state()
        self.setup_scores()
        self.video = video
        self.videomaker = None
        if video:                                                           # evaluate.py:69-71: video = (intrinsic, extrinsic, debug)
            from .videomaker import VideoMaker
            self.videomaker = VideoMaker(self.env, *video)

    def setup_gym_env(self, environment, action_type, n_objects):           # evaluate.py:148-177
        if environment in ["R1", "R2"]:
            rnd = environment
        else:
            raise Exception("Environment type has to be either R1 or R2")
        if action_type == 'macro_action' and environment == 'R2':
            raise Exception("Action type cannot be macro_action in Round 2")
        if action_type in ['joints', 'cartesian', 'macro_action']:
            act = action_type[0].upper()
        else:
            raise Exception("Action type has to be either 'joints', 'cartesian',or 'macro_action'")
        if isinstance(n_objects, int) and 1 <= n_objects <= 3:
            n_obj = n_objects
        else:
            raise Exception("Number of objects has to be 1, 2 or 3.")
        from . import make
        self.env = make('REALRobot2020-{}{}{}-v0'.format(rnd, act, n_obj))
        self.env.set_goals_dataset_path(self.goals_dataset_path)
        self.env.intrinsic_timesteps = self.intrinsic_timesteps
        self.env.extrinsic_timesteps = self.extrinsic_timesteps
        self.env.extrinsic_trials = self.extrinsic_trials
        if self.visualize:
            self.env.render('human')

    def setup_controller(self):                                             # evaluate.py:179-191
        if not issubclass(self.ControllerClass, BasePolicy):
            raise Exception("Supplied Controller is not a Sub-Class of real_robots.policy.BasePolicy . Please ensure that "
                            "the supplied controller class is derived from real_robots.policy.BasePolicy")
        self.controller = self.ControllerClass(self.env.action_space, self.env.observation_space)

    def setup_evaluation_state(self):                                       # evaluate.py:100-121
        self.evaluation_state = {"state": "PENDING", "intrinsic_phase_state": "PENDING", "extrinsic_phase_state": "PENDING",
                                 "max_intrinsic_timesteps": self.intrinsic_timesteps, "max_extrinsic_timesteps": self.extrinsic_timesteps,
                                 "current_intrinsic_timestep": 0, "max_extrinsic_trials": self.extrinsic_trials,
                                 "num_extrinsic_trials_complete": 0, "progress_in_current_extrinsic_trial": 0,
                                 "evaluation_score": {"score": 0, "score_2D": 0, "score_2.5D": 0, "score_3D": 0, "score_total": 0},
                                 "score": {"score": 0, "score_secondary": 0}}

    def sync_evaluation_state(self):
        pass                                                                # aicrowd_api is not a dependency here

    def setup_scores(self):
        self.scores = {}

    def add_scores(self, challenge, score):
        self.scores.setdefault(challenge, []).append(score)

    def run_intrinsic_phase(self):                                          # evaluate.py:203-279
        try:
            self._run_intrinsic_phase()
        except Exception as e:
            self.evaluation_state["state"] = "ERROR"
            self.evaluation_state["intrinsic_phase_state"] = "INTRINSIC_PHASE_ERROR"
            raise e

    def _run_intrinsic_phase(self):
        if not self.intrinsic_timesteps:
            self.intrinsic_timesteps = 0
        if self.intrinsic_timesteps > 0:
            observation = self.env.reset()
            reward, done, steps = 0, False, 0
            self.evaluation_state["state"] = self.evaluation_state["intrinsic_phase_state"] = "INTRINSIC_PHASE_IN_PROGRESS"
            self.controller.start_intrinsic_phase()
            if self.videomaker:
                self.videomaker.start_intrinsic()
            while not done:
                action = self.controller.step(observation, reward, done)
                observation, reward, done, _ = self.env.step(action)
                steps += 1
                self.evaluation_state["current_intrinsic_timestep"] = steps
                if self.videomaker:
                    self.videomaker.update_intrinsic(steps)
            if self.videomaker:
                self.videomaker.end_intrinsic()
            self.evaluation_state["state"] = self.evaluation_state["intrinsic_phase_state"] = "INTRINSIC_PHASE_COMPLETE"
            self.controller.end_intrinsic_phase(observation, reward, done)
        else:
            print("[WARNING] Skipping Intrinsic Phase as intrinsic_timesteps = 0 or False")
            self.evaluation_state["state"] = "INTRINSIC_PHASE_SKIPPED"

    def run_extrinsic_trial(self, trial_number):                            # evaluate.py:281-332
        self.env.reset()
        reward, done = 0, False
        observation = self.env.set_goal()
        self.controller.start_extrinsic_trial()
        if self.videomaker:
            self.videomaker.start_trial(observation, trial_number)
        steps = 0
        while not done:
            action = self.controller.step(observation, reward, done)
            observation, reward, done, _ = self.env.step(action)
            steps += 1
            self.evaluation_state["progress_in_current_extrinsic_trial"] = float(steps) / self.extrinsic_timesteps
            if self.videomaker:
                self.videomaker.extrinsic_trial(observation, action, steps, self.build_score_object() if self.scores else None)
        if self.videomaker:
            self.videomaker.end_trial()
        self.add_scores(*self.env.evaluateGoal())
        self.evaluation_state["num_extrinsic_trials_complete"] = trial_number + 1
        self.controller.end_extrinsic_trial(observation, reward, done)

    def run_extrinsic_phase(self):                                          # evaluate.py:334-394
        try:
            self.evaluation_state["state"] = self.evaluation_state["extrinsic_phase_state"] = "EXTRINSIC_PHASE_IN_PROGRESS"
            self.controller.start_extrinsic_phase()
            for trial in range(int(self.extrinsic_trials)):
                self.run_extrinsic_trial(trial)
            self.evaluation_state["state"] = self.evaluation_state["extrinsic_phase_state"] = "EXTRINSIC_PHASE_COMPLETE"
            final = self.build_score_object()
            self.evaluation_state["score"] = {"score": final["score_total"], "score_secondary": final["score_2D"]}
            self.evaluation_state["meta"] = self.evaluation_state["evaluation_score"]
            self.evaluation_state["state"] = "EVALUATION_COMPLETE"                 # evaluate.py:376-388
            self.controller.end_extrinsic_phase()
        except Exception as e:
            self.evaluation_state["state"] = "ERROR"
            self.evaluation_state["extrinsic_phase_state"] = "EXTRINSIC_PHASE_ERROR"
            raise e
        return self.build_score_object()

    def build_score_object(self):                                           # evaluate.py:396-417
        return build_score_object(self.scores, self.evaluation_state)


def build_score_object(scores, evaluation_state=None):
    total_results = []
    score_object = {}
    for key in ['2D', '2.5D', '3D']:
        if key in scores.keys():
            results = scores[key]
            challenge_score = np.mean(results)
        else:
            results = []
            challenge_score = 0
        total_results += list(results)
        score_object["score_{}".format(key)] = challenge_score
    score_object["score_total"] = np.mean(total_results) if len(total_results) else float('nan')
    if evaluation_state is not None:
        evaluation_state["evaluation_score"] = score_object
    return score_object


def evaluate(Controller, environment='R1', action_type='macro_action', n_objects=1, intrinsic_timesteps=15e6,
             extrinsic_timesteps=10e3, extrinsic_trials=50, visualize=True, goals_dataset_path="./goals.npy.npz", video=None):
    """Drop-in for real_robots.evaluate (evaluate.py:420-446): returns (score_object, scores)."""
    service = EvaluationService(Controller, environment, action_type, n_objects, intrinsic_timesteps, extrinsic_timesteps,
                                extrinsic_trials, visualize, goals_dataset_path, video)
    service.run_intrinsic_phase()
    service.run_extrinsic_phase()
    return service.build_score_object(), service.scores


# ---------------------------------------------------------------------------------------------- vectorised / sharded
def shard_range(n_items, rank, world):
    """Contiguous shard [lo, hi) of n_items for `rank` of `world` (goals are partitioned, not replicated)."""
    per = (n_items + world - 1) // world
    lo = min(rank * per, n_items)
    return lo, min(lo + per, n_items)


def gather_scores(local_ids, local_scores, n_items, group=None):
    """all_gather of (goal id, score) pairs -> dense score vector [n_items] on every rank.
    Works on NCCL (CUDA tensors) and gloo (CPU tensors); without an initialised process group it is the identity."""
    import torch
    import torch.distributed as dist
    out = torch.full((n_items,), float('nan'), dtype=torch.float64, device=local_scores.device)
    if not (dist.is_available() and dist.is_initialized()):
        out[local_ids] = local_scores
        return out
    world = dist.get_world_size(group)
    per = (n_items + world - 1) // world
    ids = torch.full((per,), -1, dtype=torch.int64, device=local_scores.device)
    sc = torch.zeros(per, dtype=torch.float64, device=local_scores.device)
    ids[:local_ids.numel()] = local_ids
    sc[:local_scores.numel()] = local_scores
    all_ids = [torch.empty_like(ids) for _ in range(world)]
    all_sc = [torch.empty_like(sc) for _ in range(world)]
    dist.all_gather(all_ids, ids, group=group)
    dist.all_gather(all_sc, sc, group=group)
    for i, s in zip(all_ids, all_sc):
        ok = i >= 0
        out[i[ok]] = s[ok]
    return out


def evaluate_vectorised(policy_factory, goals, n_objects=3, action_type='macro_action', extrinsic_timesteps=1000,
                        device=0, challenge_of=None, render=False):
    """Trial-parallel extrinsic phase (BASELINE config 4): goal g of this rank's shard runs in env g - lo.

    goals: dict(initial [G, n_obj, 7] pos+quat, final [G, n_obj, 3]) numpy arrays (synthetic or from a Goal list).
    policy_factory(env) -> BatchedPolicy.  Returns (score_object, scores dict challenge -> list, score vector [G])."""
    import torch
    import torch.distributed as dist
    from .envs.vec_env import VecREALRobotEnv
    G = goals["initial"].shape[0]
    rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    lo, hi = shard_range(G, rank, world)
    n = max(hi - lo, 1)
    env = VecREALRobotEnv(n, device=device, objects=n_objects, action_type=action_type, additional_obs=True)
    env.extrinsic_timesteps = int(extrinsic_timesteps)
    env.goal_idx = 0
    dev = env.device
    init = torch.zeros(n, n_objects, 7, dtype=torch.float64, device=dev)
    final = torch.zeros(n, n_objects, 3, dtype=torch.float64, device=dev)
    if hi > lo:
        init[:hi - lo] = torch.from_numpy(np.ascontiguousarray(goals["initial"][lo:hi])).to(dev)
        final[:hi - lo] = torch.from_numpy(np.ascontiguousarray(goals["final"][lo:hi])).to(dev)
    else:
        init[:, :, 6] = 1.0
    env.reset(render=False)
    env.set_object_poses(init)
    env.goal_positions = final
    # evaluate.py:282-285: the controller's first observation is the one set_goal() returns, i.e. taken AFTER the objects
    # were moved to the goal's initial state
    obs = env._observe(env._all if render else None)
    policy = policy_factory(env)
    reward = torch.zeros(n, dtype=torch.float64, device=dev)
    done = torch.zeros(n, dtype=torch.bool, device=dev)
    for _ in range(int(extrinsic_timesteps)):
        action = policy.step(obs, reward, done)
        action.setdefault("render", render)
        obs, reward, done, _ = env.step(action)
    local = env.evaluate_goal(final)[:hi - lo]
    ids = torch.arange(lo, hi, dtype=torch.int64, device=dev)
    scores_vec = gather_scores(ids, local, G)
    sv = scores_vec.cpu().numpy()
    scores = {}
    for g in range(G):
        ch = challenge_of[g] if challenge_of is not None else '2D'
        scores.setdefault(ch, []).append(float(sv[g]))
    env.close()
    return build_score_object(scores), scores, sv
