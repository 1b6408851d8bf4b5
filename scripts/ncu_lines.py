"""Aggregate an ncu SASS source page by CUDA source line using nvdisasm -g line info.
usage: ncu_lines.py <report.ncu-rep> <object.o> <kernel substring> [top]"""
import csv, re, subprocess, sys, tempfile, os, glob
rep, obj, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# map offset -> (file,line) inside the wanted function
line_of = {}
cur = None
infn = False
for l in sass.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        infn = re.search(kern.replace("$", r"(?![A-Za-z_])"), m.group(1)) is not None     # 'k_step_solve$' does not match k_step_solve_heavy
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m:
        line_of[int(m.group(1), 16)] = cur
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
if len(rows) < 3:
    sys.exit("no launch of a kernel matching %r in %s" % (kern, rep))
hdr = rows[1]
ia, ii, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base = None
agg = {}
tot_i = tot_s = 0
for r in rows[2:]:
    if len(r) <= isamp or not r[ia].startswith("0x"):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    key = line_of.get(a - base)
    n, t, s = int(r[ii]), int(r[it]), int(r[isamp])
    d = agg.setdefault(key, [0, 0, 0])
    d[0] += n
    d[1] += t
    d[2] += s
    tot_i += n
    tot_s += s
src = {}
print("total warp-instr %d, samples %d" % (tot_i, tot_s))
print("%-28s %12s %6s %8s %6s  %s" % ("file:line", "warp-instr", "%", "samples", "%", "avg thr"))
for key, d in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    f = "%s:%d" % key if key else "?"
    text = ""
    if key:
        p = [p for p in glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "real_robots_b200", "csrc", "*")) if os.path.basename(p) == key[0]]
        if p:
            if p[0] not in src:
                src[p[0]] = open(p[0]).read().splitlines()
            text = src[p[0]][key[1] - 1].strip()[:90]
    print("%-28s %12d %6.2f %8d %6.2f  %5.1f  %s" % (f, d[0], 100 * d[0] / tot_i, d[2], 100 * d[2] / max(tot_s, 1), d[1] / max(d[0], 1), text))
