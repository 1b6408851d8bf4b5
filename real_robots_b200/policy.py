"""real_robots.policy.BasePolicy — the controller callback interface (reference real_robots/policy.py:1-103)."""


class BasePolicy:
    def __init__(self, action_space, observation_space):
        self.action_space = action_space
        self.observation_space = observation_space

    def step(self, observation, reward, done):
        """Called every timestep; must return an action from `action_space` (policy.py:7-54)."""
        raise NotImplementedError("Please derive your custom controller from this class and implement `step`")

    def start_intrinsic_phase(self):
        pass

    def end_intrinsic_phase(self, observation, reward, done):
        pass

    def start_extrinsic_phase(self):
        pass

    def end_extrinsic_phase(self):
        pass

    def start_extrinsic_trial(self):
        pass

    def end_extrinsic_trial(self, observation, reward, done):
        pass


class BatchedPolicy:
    """Vectorised controller for VecREALRobotEnv / evaluate_vectorised: step() maps tensor observations [N, ...]
    to an action dict of tensors [N, ...]."""

    def __init__(self, env):
        self.env = env

    def step(self, observation, reward, done):
        raise NotImplementedError


class RandomBatchedPolicy(BatchedPolicy):
    """Vectorised real_robots/cli.py:12-20 RandomPolicy: draw from the action Box, redraw with p = 0.05 per step."""

    def __init__(self, env, seed=1234):
        import torch
        super().__init__(env)
        self.g = torch.Generator(device=env.device)
        self.g.manual_seed(seed)
        self.action = env.sample_random_actions(self.g)

    def step(self, observation, reward, done):
        import torch
        redraw = torch.rand(self.env.n_envs, device=self.env.device, generator=self.g) < 0.05
        new = self.env.sample_random_actions(self.g)
        self.action = torch.where(redraw[:, None], new, self.action)
        return {"joint_command": self.action, "render": True}
