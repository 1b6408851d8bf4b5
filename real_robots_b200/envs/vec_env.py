"""VecREALRobotEnv — the batched interface added on top of the reference's gym surface: N independent
REALRobot2020 envs stepped in lockstep on one GPU, PyTorch CUDA tensors in and out, zero-copy.

Semantics per env are those of REALRobotEnv.step_joints / step_cartesian / step_macro (reference env.py:326-467);
the request caches of the cartesian / macro action types (env.py:366-370, 396-403) are kept per env on the device.
"""
import math

import torch

from ..sim import BatchedSim
from ..util import euler_to_quat
from .robot import Kuka


class VecREALRobotEnv:
    intrinsic_timesteps = int(15e6)
    extrinsic_timesteps = int(10e3)

    def __init__(self, n_envs, device=0, objects=3, action_type='joints', additional_obs=True, **cfg):
        if action_type not in ('joints', 'cartesian', 'macro_action'):
            raise ValueError("action_type must be one 'joints', 'cartesian' or 'macro_action'")
        self.n_envs, self.action_type, self.additional_obs = int(n_envs), action_type, additional_obs
        self.robot = Kuka(additional_obs, objects)
        self.sim = BatchedSim(n_envs, objects, device=device, **cfg)
        self.robot._bind(self.sim)
        self.device = self.sim.device
        N, dev = self.n_envs, self.device
        self.timestep = torch.zeros(N, dtype=torch.int64, device=dev)
        self.goal_idx = -1
        self._zero9 = torch.zeros(N, 9, dtype=torch.float64, device=dev)
        self._all = torch.ones(N, dtype=torch.uint8, device=dev)
        self.goal_positions = None           # [N, n_obj, 3] when goals are set (evaluate)
        if action_type == 'cartesian':
            self.requested = torch.full((N, 7), float('nan'), dtype=torch.float64, device=dev)
            self.last_ik = torch.zeros(N, 11, dtype=torch.float64, device=dev)
        if action_type == 'macro_action':
            self.requested = torch.full((N, 2, 2), float('nan'), dtype=torch.float64, device=dev)
            self.plan = torch.zeros(N, 1000, 9, dtype=torch.float64, device=dev)
            self.plan_step = torch.full((N,), 1000, dtype=torch.int64, device=dev)
            q = euler_to_quat([0, 3.14, -1.57])
            self._plan_quat = torch.tensor(q, dtype=torch.float64, device=dev)
        low = torch.tensor(self.robot.min_joints, dtype=torch.float64, device=dev)
        high = torch.tensor(self.robot.max_joints, dtype=torch.float64, device=dev)
        self.joints_low, self.joints_high = low, high

    # ------------------------------------------------------------------ lifecycle
    def reset(self, mask=None, render=True):
        self.sim.reset(mask)
        if mask is None:
            self.timestep.zero_()
        else:
            self.timestep.masked_fill_(mask.to(self.device).bool(), 0)
        return self._observe(self._all if render else None)

    def close(self):
        self.sim.close()

    # ------------------------------------------------------------------ observations
    def _observe(self, render_mask):
        K = Kuka.ObsSpaces
        obs = {}
        if render_mask is None:
            # render off for every env: the reference returns its cached zero arrays (env.py:114-122, 273-275) and so do we —
            # one zero image per kind, broadcast over the env axis (read-only views), no kernel launch, no 3.46 GB of zero stores
            rgb, depth, seg = self._zero_images()
        else:
            # rendered before the state read-out: right after a step the rasteriser overlaps the tail of the solver (rr_render)
            rgb, depth, seg = self.sim.render(mask=render_mask, want_mask=self.additional_obs)
        joints, touch, objs = self.sim.read_state()
        obs[K.JOINT_POSITIONS], obs[K.TOUCH_SENSORS] = joints, touch
        obs[K.RETINA], obs[K.DEPTH] = rgb, depth
        if self.additional_obs:
            obs[K.MASK] = seg
            obs[K.OBJ_POS] = objs[:, :, :3]
            obs[K.GOAL_POS] = self.goal_positions
        return obs

    def _zero_images(self):
        if getattr(self, "_zimg", None) is None:
            N, dev = self.n_envs, self.device
            H, W = Kuka.eye_height, Kuka.eye_width
            self._zimg = (torch.zeros(1, H, W, 3, dtype=torch.uint8, device=dev).expand(N, H, W, 3),
                          torch.zeros(1, H, W, dtype=torch.float32, device=dev).expand(N, H, W),
                          torch.zeros(1, H, W, dtype=torch.int32, device=dev).expand(N, H, W))
        return self._zimg

    def _render_mask(self, r):
        if r is None:
            return None
        if isinstance(r, (bool, int)):
            return self._all if r else None
        return r.to(device=self.device, dtype=torch.uint8).reshape(self.n_envs).contiguous()

    # ------------------------------------------------------------------ stepping
    def step(self, actions):
        if self.action_type == 'joints':
            cmd = actions.get('joint_command')
            cmd = self._zero9 if cmd is None else cmd
        elif self.action_type == 'cartesian':
            cmd = self._cartesian(actions['cartesian_command'], actions['gripper_command'])
        else:
            cmd = self._macro(actions['macro_action'])
        return self.step_joints(cmd, actions.get('render', True))

    def step_joints(self, joint_command, render=True):
        self.sim.step(joint_command)
        obs = self._observe(self._render_mask(render))
        self.timestep += 1
        limit = self.intrinsic_timesteps if self.goal_idx < 0 else self.extrinsic_timesteps
        done = self.timestep >= limit
        reward = torch.zeros(self.n_envs, dtype=torch.float64, device=self.device)
        return obs, reward, done, {}

    def _cartesian(self, cart, grip):
        # env.py:358-386: IK only when the (coords, orientation) request differs from the cached one (exact equality)
        changed = (cart != self.requested).any(dim=1) | torch.isnan(self.requested).any(dim=1)
        if bool(changed.any()):
            ik = self.sim.ik(cart.contiguous(), 1000, 0.001, mask=changed)
            self.last_ik = torch.where(changed[:, None], ik, self.last_ik)
            self.requested = torch.where(changed[:, None], cart, self.requested)
        return torch.cat([self.last_ik[:, :7], grip], dim=1).contiguous()

    def _ik_xyz(self, xyz, mask):
        pq = torch.cat([xyz, self._plan_quat.expand(self.n_envs, 4)], dim=1).contiguous()
        return self.sim.ik(pq, 1000, 0.001, mask=mask)[:, :9]

    def _macro(self, macro):
        # env.py:388-467: same action -> next plan row; new action or exhausted plan -> generate_plan
        same = (macro == self.requested).flatten(1).all(dim=1)
        self.plan_step = torch.where(same, self.plan_step + 1, self.plan_step)
        need = (~same) | (self.plan_step >= 1000)
        if bool(need.any()):
            self._generate_plans(macro, need)
            self.requested = torch.where(need[:, None, None], macro, self.requested)
            self.plan_step = torch.where(need, torch.zeros_like(self.plan_step), self.plan_step)
        return self.plan[torch.arange(self.n_envs, device=self.device), self.plan_step].contiguous()

    def _generate_plans(self, macro, need):
        N, dev = self.n_envs, self.device
        p1, p2 = macro[:, 0], macro[:, 1]

        def at(p, z):
            return torch.cat([p, torch.full((N, 1), z, dtype=torch.float64, device=dev)], dim=1)
        home = torch.zeros(9, dtype=torch.float64, device=dev)
        home2 = home.clone()
        home2[5] = math.pi / 2
        home2[6] = math.pi / 2
        p1h, p1l, p2h = self._ik_xyz(at(p1, 0.6), need), self._ik_xyz(at(p1, 0.46), need), self._ik_xyz(at(p2, 0.6), need)
        dist = (p2 - p1).norm(dim=1)
        pieces = torch.clamp((dist / 0.05).to(torch.int64) + 1, max=500)
        max_pieces = int(pieces[need].max().item())
        chunk = (500 // pieces)
        iks = []
        piecesf = pieces.to(torch.float64)
        for i in range(max_pieces):
            # np.linspace(p1, p2, pieces + 1)[i + 1] bit for bit: (i + 1) / pieces as a correctly rounded tensor / tensor
            # division (python scalar / tensor is scalar * reciprocal: one ulp off for some i), times delta, plus start;
            # the end point is p2 itself
            frac = torch.full_like(piecesf, float(i + 1)) / piecesf
            way = torch.where((pieces <= i + 1)[:, None], p2, (p2 - p1) * frac[:, None] + p1)
            coord = at(way, 0.46)
            iks.append(self._ik_xyz(coord, need & (pieces > i)))
        iks = torch.stack(iks, dim=1)                                        # [N, max_pieces, 9]
        rows = torch.arange(500, device=dev)[None, :]
        idx = torch.minimum(rows // chunk[:, None], (pieces - 1)[:, None])     # joints[i*chunk:] = ik_i  (last wins)
        idx = idx.clamp(max=max_pieces - 1)      # envs that keep their plan may have more pieces than were solved
        sweep = torch.gather(iks, 1, idx[:, :, None].expand(N, 500, 9))
        plan = torch.cat([home2.expand(N, 100, 9), p1h[:, None].expand(N, 100, 9), p1l[:, None].expand(N, 50, 9), sweep,
                          p2h[:, None].expand(N, 50, 9), home2.expand(N, 100, 9), home.expand(N, 100, 9)], dim=1)
        self.plan = torch.where(need[:, None, None], plan, self.plan)

    # ------------------------------------------------------------------ goals / scoring (env.py:151-200)
    def set_object_poses(self, pos_quat, mask=None):
        """pos_quat [N, n_obj, 7]: BodyPart.reset_pose for every object (set_goal's initial_state)."""
        for o in range(self.sim.n_objects):
            self.sim.set_object_pose(o, pos_quat[:, o].contiguous(), mask)

    def evaluate_goal(self, goal_xyz):
        return self.sim.score(goal_xyz.contiguous())

    def fault_flags(self):
        """Per-env failure flags of the last step (BatchedSim.faults): a batched env cannot raise for one env the way
        Kuka.apply_action's assert does (robot.py:189), so non-finite commands / states are flagged instead."""
        return self.sim.faults()

    def sample_random_actions(self, generator=None):
        """U(low, high) of the joints action Box (cli.RandomPolicy draw)."""
        u = torch.rand(self.n_envs, 9, dtype=torch.float64, device=self.device, generator=generator)
        return self.joints_low + (self.joints_high - self.joints_low) * u


def gather_retinas(retina, factor=4, dst=None, group=None):
    """Policy hand-off after the step path (not in the reference): box-filter the retinas on the device by
    `factor` (240x320 -> 60x80 for 4) and collect every rank's envs on the learner.

    retina: u8 [n, 240, 320, 3] (CUDA for NCCL over NVLink, CPU for gloo).  Returns u8 [world * n, 240/f, 320/f, 3] on every
    rank (dst=None, all_gather) or on rank `dst` only (gather; None elsewhere).  Without an initialised process group it is
    just the down-sampling.  Equal n on all ranks (envs are sharded evenly)."""
    import torch.distributed as dist
    n, h, w, c = retina.shape
    if h % factor or w % factor:
        raise ValueError("factor must divide 240 and 320")
    small = retina.reshape(n, h // factor, factor, w // factor, factor, c).to(torch.float32).mean(dim=(2, 4))
    small = small.round_().clamp_(0, 255).to(torch.uint8).contiguous()
    if not (dist.is_available() and dist.is_initialized()):
        return small
    world = dist.get_world_size(group)
    if dst is None:
        parts = [torch.empty_like(small) for _ in range(world)]
        dist.all_gather(parts, small, group=group)
        return torch.cat(parts, dim=0)
    me = dist.get_rank(group)
    parts = [torch.empty_like(small) for _ in range(world)] if me == dst else None
    dist.gather(small, parts, dst=dst, group=group)
    return torch.cat(parts, dim=0) if me == dst else None
