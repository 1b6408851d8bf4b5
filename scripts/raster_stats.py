"""Event counts of the rasteriser (RR_STATS tuning build): RR_LIB=real_robots_b200/variants/lib_stats.so python scripts/raster_stats.py"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
from real_robots_b200.sim import BatchedSim
from real_robots_b200 import _lib
N = 256
sim = BatchedSim(N, 3)
g = torch.Generator(device="cuda"); g.manual_seed(0)
act = (torch.rand(N, 9, dtype=torch.float64, device="cuda", generator=g) * 2 - 1) * 2.0
for i in range(60): sim.step(act)
lib = _lib.load()
out = (ctypes.c_ulonglong * 16)()
lib.rr_debug_raster_stats(out, 1)
sim.render()
lib.rr_debug_raster_stats(out, 1)
names = ["triangles tested", "after back-face cull", "records", "list entries small", "list entries medium", "list entries huge", "pixels scanned", "covered",
         "early-z rejected", "depth computed", "atomicMax issued", "pixels shaded", "-", "MISMATCH pre-test back / exact front", "MISMATCH pre-test front / exact back", "uncertain (exact test)"]
for n, v in zip(names, out):
    print("%-22s %12.1f per env" % (n, v / N))
