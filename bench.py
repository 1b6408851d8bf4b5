#!/usr/bin/env python
"""bench.py — env-steps/sec incl. retina render at 4096 envs/GPU (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # CPU arm: the oracle port on all host cores

Our arm: one step = rr_step (action prologue + physics) + rr_read_state + rr_render (RGB + depth + mask) for 4096
envs with 3 objects per GPU, actions from the vectorised cli.RandomPolicy (p = 0.05 resample, seed 1234 + rank)
pre-generated on the device.  `value` is timed with CUDA events between barriers, max over ranks; `e2e` times
the same step through rr_step_host (pinned host buffers, H2D of the commands and D2H of every observation inside
the timed region).  The roofline entry is for the rasteriser kernel (k_raster): algorithmic bytes = the
844 800 B/env of image output it must write over its CUDA-event duration.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_ENVS = 4096
N_OBJECTS = 3
IMG_BYTES = 240 * 320 * (3 + 4 + 4)          # rgb u8x3 + depth f32 + mask i32 = 844 800
STATE_BYTES = 3824                            # BASELINE.md §4 physics bytes per env-step
METRIC = "env_steps_per_sec_with_retina_4096_envs_per_gpu"


def workload(envs):
    """config.workload of BOTH arms (the reference arm runs a bounded sample of the same workload on the host cores)."""
    return ("BASELINE configs[2]: REALRobot2020 R1J3 joints action, %d envs/GPU, 3 objects, random policy "
            "(p=0.05 resample), physics + touch + RGB/depth/mask retina 240x320" % envs)


_JSON_OUT = None


def quiet_stdout():
    """stdout carries exactly ONE JSON line: anything native libraries print to fd 1 on the way (the NCCL version
    banner at communicator creation) is sent to stderr, and the line goes to the original stdout."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(float(s[0])) for s in self.samples if s and s[0].replace(".", "").isdigit())
        mx = [int(float(s[1])) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            for i, n in enumerate(names):
                if len(s) > 3 + i and s[3 + i].lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.samples)}


def random_policy_actions(torch, steps, n, device, seed):
    """Vectorised real_robots/cli.py:12-20 RandomPolicy: sample from the action Box, resample with p = 0.05."""
    import math
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    lo = torch.full((9,), -math.pi * 0.944, dtype=torch.float64, device=device)
    hi = -lo.clone()
    for i in (0, 1, 3, 5):
        lo[i], hi[i] = -math.pi * 0.666, math.pi * 0.666
    lo[6], hi[6] = -math.pi * 0.972, math.pi * 0.972
    lo[7:] = 0.0
    hi[7:] = math.pi / 2
    acts = torch.empty(steps, n, 9, dtype=torch.float64, device=device)
    cur = lo + (hi - lo) * torch.rand(n, 9, dtype=torch.float64, device=device, generator=g)
    for t in range(steps):
        resample = torch.rand(n, device=device, generator=g) < 0.05
        new = lo + (hi - lo) * torch.rand(n, 9, dtype=torch.float64, device=device, generator=g)
        cur = torch.where(resample[:, None], new, cur)
        acts[t] = cur
    return acts


def oracle_worker(args):
    """One CPU process: one oracle env, `steps` env-steps with render; returns wall seconds for the timed part."""
    seed, steps, warm = args
    import numpy as np
    from oracle.orc import Oracle
    rng = np.random.default_rng(seed)
    o = Oracle(N_OBJECTS)
    lo = np.array([-2.09, -2.09, -2.96, -2.09, -2.96, -2.09, -3.05, 0, 0])
    hi = -lo
    hi[7:] = np.pi / 2
    a = rng.uniform(lo, hi)
    for _ in range(60):                 # same untimed pre-roll as the GPU arm: objects come to rest on the table
        o.step(a)
    for _ in range(warm):
        o.step(a); o.read(); o.render()
    t0 = time.perf_counter()
    for _ in range(steps):
        if rng.random() < 0.05:
            a = rng.uniform(lo, hi)
        o.step(a); o.read(); o.render()
    return time.perf_counter() - t0


def cpu_run(steps_per_worker, workers, pool=None):
    import multiprocessing as mp
    jobs = [(1234 + w, steps_per_worker, 3) for w in range(workers)]
    if pool is None:
        with mp.get_context("spawn").Pool(workers) as p:
            times = p.map(oracle_worker, jobs, chunksize=1)
    else:
        times = pool.map(oracle_worker, jobs, chunksize=1)
    wall = max(times)
    return workers * steps_per_worker / wall, wall


def run_reference(args):
    """CPU arm: pybullet is not installable here (absent from /root/reference and the wheelhouse), so this times
    the oracle port — the CPU restatement of the reference's step path — one process per host core."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import orc
    orc.build()
    cores = os.cpu_count() or 1
    per_step = 25                       # env-steps per worker per bench "step" (bounded sample)
    import multiprocessing as mp
    vals = []
    t_all = 0.0
    with mp.get_context("spawn").Pool(cores) as pool:      # one worker process per host core, reused by every bench step
        for _ in range(args.warmup):
            cpu_run(5, cores, pool)
        for _ in range(args.steps):
            v, wall = cpu_run(per_step, cores, pool)
            vals.append(v); t_all += wall
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_all / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 physics / f32 raster", "data": "synthetic",
        "config": {"workload": workload(args.envs), "n_envs_per_gpu": args.envs, "n_objects": N_OBJECTS,
                   "cpu_sample": "one oracle env per host core, %d env-steps per core per bench step" % per_step},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": "%d processes x %d env-steps with render per bench step (oracle port; pybullet not installable)" % (cores, per_step)},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def bind_to_gpu_numa(local):
    """Pin this rank to the CPUs of its GPU's NUMA node before any pinned host memory is allocated, so that the 3.46 GB of
    images a step copies to the host land in memory next to the GPU's PCIe root (matters for the e2e leg at N > 1).
    Returns a description for the JSON line; never fails the bench."""
    try:
        bdf = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bdf.count(":") == 2 and len(bdf.split(":")[0]) == 8:
            bdf = bdf[4:]                                    # nvidia-smi prints an 8-digit domain, sysfs uses 4
        base = "/sys/bus/pci/devices/" + bdf
        node = int(open(base + "/numa_node").read())
        cpus = set()
        for part in open(base + "/local_cpulist").read().strip().split(","):
            if part:
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if node < 0 or not allowed:
            return "unbound (no NUMA node for %s)" % bdf
        os.sched_setaffinity(0, allowed)
        return "node %d (%d cpus)" % (node, len(allowed))
    except Exception as ex:                                   # noqa: BLE001
        return "unbound (%s)" % type(ex).__name__


def run_ours(args):
    import torch
    import torch.distributed as dist
    from real_robots_b200.sim import BatchedSim

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    numa = bind_to_gpu_numa(local) if world > 1 else "unbound (single rank)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    K, W = args.steps, max(args.warmup, 3)
    sim = BatchedSim(args.envs, N_OBJECTS, device=local)
    acts = random_policy_actions(torch, K + W, args.envs, dev, 1234 + rank)

    def step(t):
        sim.step(acts[t])
        sim.render()              # before read_state: the rasteriser starts while the solver's heavy-env kernel still runs (it waits
        sim.read_state()          # per env inside the kernel); read_state joins that kernel first

    # untimed pre-roll: after reset the objects are in the air (robot.py:19-24) and land ~27 steps later; the timed
    # region should see the steady state of a random-policy rollout (resting and arm contacts), not the free fall
    for t in range(args.settle):
        sim.step(acts[t % (K + W)])
    for t in range(W):
        step(t)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    sim.profile(True)
    l0 = sim.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for t in range(W, W + K):
        step(t)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = sim.launch_count() - l0
    prof = sim.profile_read()
    sim.profile(False)
    sampler.stop_flag = True
    # the rasteriser by itself (after the timed region, not part of it): inside a step the look-ahead narrowphase and the solver's
    # heavy-env kernel run beside it, so its in-step duration is longer than its own
    raster_alone_ms = None
    if rank == 0:
        torch.cuda.synchronize()
        sim.profile(True)
        for t in range(10):                # the rollout goes on (the rasteriser's cost depends on the scene), with the device drained
            sim.step(acts[(W + K + t) % (W + K)])         # between the step's kernels and the render
            torch.cuda.synchronize()
            sim.render()
            sim.read_state()
        alone = sim.profile_read()
        sim.profile(False)
        raster_alone_ms = alone["raster"][0] / max(alone["raster"][1], 1)
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    # per rank: timed region and the sum of the kernels' own durations (the ranks roll out different seeds, and the solver's
    # duration follows the most contact-laden envs of a rank: the spread between ranks is workload, not communication)
    mine = torch.tensor([ms / K, sum(v[0] / max(v[1], 1) for v in prof.values())], dtype=torch.float64, device=dev)
    per_rank = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        dist.all_gather(per_rank, mine)
    else:
        per_rank = [mine]
    per_rank = [[round(float(x), 3) for x in t.tolist()] for t in per_rank]
    ms_max = float(tms.item())
    value = world * args.envs * K / (ms_max * 1e-3)

    # ---- end-to-end leg through the host-buffer C-ABI call (pinned host memory, copies inside the timed region)
    # The random policy is open loop, so the driver keeps two steps in flight (rr_step_host_begin / _end): the physics and first
    # render chunks of step t+1 run while the copy engine drains step t's images into the other set of pinned buffers.
    Ke = max(3, min(K, args.e2e_steps))
    host_acts_t = acts[:Ke + 2].cpu().pin_memory()
    host_acts = host_acts_t.numpy()
    mask_t = torch.ones(args.envs, dtype=torch.uint8).pin_memory()
    mask = mask_t.numpy()
    outs = [sim.alloc_host(images=True, want_mask=True, pinned=True) for _ in range(2)]
    for t in range(2):
        sim.step_host(host_acts[t], mask, out=outs[t])
    barrier()
    t0 = time.perf_counter()
    sim.step_host_begin(host_acts[2], mask, outs[0])
    for t in range(3, 2 + Ke):
        sim.step_host_begin(host_acts[t], mask, outs[(t - 2) & 1])
        sim.step_host_end()
    sim.step_host_end()
    te = time.perf_counter() - t0
    tt = torch.tensor([te], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_value = world * args.envs * Ke / float(tt.item())
    h2d = args.envs * 9 * 8 + args.envs
    d2h = args.envs * (9 * 8 + 4 * 8 + N_OBJECTS * 7 * 8 + IMG_BYTES)
    # ---- the ceiling of that leg: nothing but the device-to-host copies of one step's images into pinned memory, all ranks at once
    rgb_d, depth_d, seg_d = sim.images()
    pinned = outs[0]["_torch"][3:6]
    cstream = torch.cuda.Stream(device=dev)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ceil_ms = []
    for rep in range(4):
        barrier()
        with torch.cuda.stream(cstream):
            c0.record()
            for h, d in zip(pinned, (rgb_d, depth_d, seg_d)):
                h.copy_(d, non_blocking=True)
            c1.record()
        cstream.synchronize()
        if rep:
            ceil_ms.append(c0.elapsed_time(c1))
    ct = torch.tensor([max(ceil_ms) if world > 1 else min(ceil_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ct, op=dist.ReduceOp.MAX)
    d2h_ceiling_gbs = args.envs * IMG_BYTES / (float(ct.item()) * 1e-3) / 1e9           # per rank, with every rank copying
    e2e_ceiling = world * args.envs / (float(ct.item()) * 1e-3)
    # ---- N > 1: the score gather of evaluate() over NCCL on real rr_score output (the only collective of the port)
    collective = None
    if world > 1:
        from real_robots_b200.evaluate import gather_scores, shard_range
        G = world * 64
        lo, hi = shard_range(G, rank, world)
        goal = sim.read_state()[2][:, :, :3].contiguous() + 0.05
        local = sim.score(goal)[:hi - lo]
        vec = gather_scores(torch.arange(lo, hi, dtype=torch.int64, device=dev), local, G)
        ok = bool(torch.isfinite(vec).all()) and bool(torch.equal(vec[lo:hi], local)) and bool(((vec > 0) & (vec <= N_OBJECTS)).all())
        okt = torch.tensor([1.0 if ok else 0.0], dtype=torch.float64, device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)
        collective = {"what": "evaluate.gather_scores (all_gather over NCCL) of rr_score output, %d goals" % G, "ranks": world, "ok": bool(okt.item() == 1.0)}

    if rank == 0:
        peak, peak_src = measured_peaks()
        r_ms, r_n = prof["raster"]
        s_ms, s_n = prof["step"]
        raster_ms = r_ms / max(r_n, 1)
        achieved = args.envs * IMG_BYTES / (raster_ms * 1e-3) / 1e9 if raster_ms > 0 else 0.0
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 physics / f32 raster", "data": "synthetic",
            "config": {"workload": workload(args.envs),
                       "n_envs_per_gpu": args.envs, "n_objects": N_OBJECTS, "parallelism": "envs sharded, %d rank(s), no collective" % world, "numa": numa,
                       "static_layer": ("off (RR_STATIC_LAYER=0): every shape rasterised for every env" if os.environ.get("RR_STATIC_LAYER", "1")[:1] == "0" else
                                        "table + robot base rasterised once per context, every env's z-buffer starts from it (images bit-identical "
                                        "to rasterising everything; RR_STATIC_LAYER=0 disables)"),
                       "pre_roll": "%d untimed physics steps so that objects rest on the table (contacts active)" % args.settle,
                       "l2": "outputs larger than L2 (%.2f GB written per step per GPU)" % (args.envs * IMG_BYTES / 1e9)},
            "roofline": {"kernel": "k_raster", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": (args.traffic * args.envs / 4096 if args.traffic else None),
                         "algorithmic_bytes": args.envs * IMG_BYTES, "peak_source": peak_src,
                         "ms_per_launch": raster_ms, "launches": r_n,
                         "alone": {"ms_per_launch": raster_alone_ms, "achieved": (args.envs * IMG_BYTES / (raster_alone_ms * 1e-3) / 1e9) if raster_alone_ms else None,
                                   "frac": (args.envs * IMG_BYTES / (raster_alone_ms * 1e-3) / 1e9 / peak) if raster_alone_ms and peak else None,
                                   "how": "10 more steps after the timed region with a device synchronisation between rr_step and rr_render, so nothing runs beside the rasteriser; achieved / frac above are the in-step "
                                          "launches, which share the GPU with the look-ahead narrowphase and the solver's heavy-env kernel"},
                         "kernels_ms_per_step": {k: (v[0] / max(s_n, 1)) for k, v in prof.items()},
                         "kernels_note": "CUDA-event time per step on the stream each kernel runs on.  step = k_step_dyn + k_step_solve; the solver's "
                                         "heavy-env kernel and the look-ahead narrowphase of the next step (collide) run on side streams beside "
                                         "k_raster, so the entries overlap and do not add up to ms_per_step",
                         # SURVEY §8(d) secondary rates of the rasteriser: 17 578 scene triangles and 76 800 pixels per env
                         "secondary": {"triangles_per_s": args.envs * 17578 / (raster_ms * 1e-3) if raster_ms > 0 else None,
                                       "pixels_per_s": args.envs * 76800 / (raster_ms * 1e-3) if raster_ms > 0 else None,
                                       "image_write_GBps": achieved},
                         "physics_kernel": {"kernel": "k_step_dyn + k_step_solve", "ms_per_launch": s_ms / max(s_n, 1),
                                            "achieved_gbs": args.envs * STATE_BYTES / (s_ms / max(s_n, 1) * 1e-3) / 1e9 if s_ms > 0 else None}},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": Ke,
                    "api": "rr_step_host_begin / rr_step_host_end, two steps in flight, pinned host buffers",
                    "d2h_ceiling_gbs": d2h_ceiling_gbs, "ceiling": e2e_ceiling, "frac_of_ceiling": e2e_value / e2e_ceiling if e2e_ceiling else None,
                    "ceiling_how": "plain pinned D2H copy of one step's images (%.2f GB per rank), all ranks at once, slowest rank" % (args.envs * IMG_BYTES / 1e9)},
            "collective_check": collective,
            "gpu_launches": launches,
            "per_rank_ms": {"ms_per_step": [p[0] for p in per_rank], "kernel_sum_ms": [p[1] for p in per_rank]},
            "clocks": sampler.summary(),
        }
        if args.cpu_baseline:
            from oracle import orc
            orc.build()
            n_cpu = 3000                                # ~11 s of single-core CPU work
            wall = oracle_worker((1234, n_cpu, 3))
            line["cpu_baseline"] = {"value": n_cpu / wall, "unit": "env-steps/s", "cores": 1, "kind": "port",
                                    "sample": "1 oracle env x %d env-steps with render (%.1f s)" % (n_cpu, wall)}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=N_ENVS)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--settle", type=int, default=60, help="untimed physics-only pre-roll steps before the warm-up")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--traffic", type=float, default=4.80e9,
                    help="dram bytes (read + write) per 4096-env k_raster launch from the ncu capture in profiles/r02c_ncu_full_summary.json (0.45 GB read + 4.35 GB written)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        if int(os.environ.get("RANK", "0")) != 0:
            args.cpu_baseline = False
        run_ours(args)


if __name__ == "__main__":
    main()
