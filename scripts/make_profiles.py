"""Turn the ncu captures brought back in gpurun_out/ into the tracked summaries under profiles/.
usage: make_profiles.py <tag> <launches.csv> <full.ncu-rep>"""
import csv, json, os, subprocess, sys
tag, launches, rep = sys.argv[1:4]
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
prof = os.path.join(root, "profiles")
csrc = os.path.join(root, "real_robots_b200", "csrc")

# ---- launch list
rows = [r for r in csv.reader(l for l in open(launches) if not l.startswith("==")) if len(r) > 5]
hdr = rows[0]
ik, iv, im = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
agg = {}
for r in rows[1:]:
    if r[im] != "gpu__time_duration.sum":
        continue
    name = r[ik].split("(")[0]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += float(r[iv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(prof, tag + "_launches_summary.txt"), "w") as f:
    f.write("# ncu launch list (%s): ncu --metrics gpu__time_duration.sum --clock-control none -s 190 -c 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3\n" % tag)
    f.write("# cold-cache, serialised per-launch times: compare SHARES, not absolutes. unit: ns.  (bench + e2e legs: the e2e leg renders in 16 chunks of 256 envs)\n")
    f.write("%-40s %8s %14s %8s %12s\n" % ("kernel", "launches", "total", "share", "avg"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%-40s %8d %14.1f %7.2f%% %12.1f\n" % (k, v[0], v[1], 100 * v[1] / tot, v[1] / v[0]))
subprocess.run(["cp", launches, os.path.join(prof, tag + "_launches.csv")])

# ---- full capture: selected metrics per kernel
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
h, units = r[0], r[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "launch__grid_size", "launch__block_size",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
out = {"command": "ncu --set full --clock-control none --import-source on -k regex:\"k_raster|k_step_dyn|k_step_solve|k_collide\" -s 192 -c 4 (r02) | -s 194 -c 10 (r02c) python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 3",
       "units": {k: units[h.index(k)] for k in keys if k in h}, "kernels": []}
for v in r[2:]:
    d = {k: v[h.index(k)] for k in keys if k in h}
    st = [(h[i].split("stalled_")[1], float(v[i])) for i in range(len(h)) if "pcsamp_warps_issue_stalled" in h[i] and "not_issued" not in h[i]]
    tot_s = sum(x[1] for x in st) or 1.0
    d["stall_reasons_pct"] = {n: round(100 * x / tot_s, 1) for n, x in sorted(st, key=lambda t: -t[1])[:8]}
    out["kernels"].append(d)
json.dump(out, open(os.path.join(prof, tag + "_ncu_full_summary.json"), "w"), indent=1)

# ---- per-phase / per-line instruction and sample tables
def run(script, obj, kern, src, dst, extra=()):
    res = subprocess.run([sys.executable, os.path.join(root, "scripts", script), rep, os.path.join(csrc, obj), kern] + ([os.path.join(csrc, src)] if src else []) + list(extra),
                         capture_output=True, text=True).stdout
    open(os.path.join(prof, dst), "w").write(res)
run("ncu_phases.py", "rr_render.o", "k_raster", "rr_render.cu", tag + "_k_raster_phases.txt")
run("ncu_lines.py", "rr_render.o", "k_raster", None, tag + "_k_raster_hot_lines.txt", ["30"])
run("ncu_phases.py", "rr_physics.o", "k_step_solve$", "rr_physics.cu", tag + "_k_step_solve_phases.txt")
run("ncu_lines.py", "rr_physics.o", "k_step_solve$", None, tag + "_k_step_solve_hot_lines.txt", ["30"])
run("ncu_phases.py", "rr_physics.o", "k_step_solve_heavy", "rr_physics.cu", tag + "_k_step_solve_heavy_phases.txt")
run("ncu_phases.py", "rr_physics.o", "k_step_dyn", "rr_physics.cu", tag + "_k_step_dyn_phases.txt")
run("ncu_phases.py", "rr_collide.o", "k_collide", "rr_collide.cu", tag + "_k_collide_phases.txt")
print("profiles written for", tag)
