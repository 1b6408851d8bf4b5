"""Aggregate an ncu SASS source page by *function / phase* of a kernel.

usage: ncu_phases.py <report.ncu-rep> <object.o> <kernel substring> <source.cu> [marker ...]

Instructions are attributed to the INNERMOST inlined source line nvdisasm reports; this script maps each such line to the
enclosing top-level function of <source.cu> (device helpers are inlined into the kernel, so their lines keep their own
function name), and lines inside the kernel itself to the last `// ----` phase marker comment above them.
"""
import csv, glob, os, re, subprocess, sys, tempfile

rep, obj, kern, srcfile = sys.argv[1:5]
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
sass = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
line_of, cur, infn = {}, None, False
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

# label every line of the source file
src = open(srcfile).read().splitlines()
label, fn = {}, "?"
for i, text in enumerate(src, 1):
    m = re.match(r"^(?:template.*>\s*)?(?:static\s+)?(?:__device__|__global__)[^(]*?(\w+)\s*\(", text)
    mk = re.search(r"\b(k_\w+)\s*\(", text) if text.startswith("__global__") else None
    if mk:
        fn = mk.group(1)
    elif m:
        fn = m.group(1)
    elif re.match(r"^k_\w+\(", text):
        fn = text.split("(")[0]
    m = re.match(r"^\s+// ---- (.*)", text)
    if m and fn.startswith("k_"):
        fn = fn.split(":")[0] + ":" + m.group(1)[:28]
    label[i] = fn
base_name = os.path.basename(srcfile)

csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "-k", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
if len(rows) < 3:
    sys.exit("no launch of a kernel matching %r in %s" % (kern, rep))
hdr = rows[1]
ia, ii, it, isamp = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
base, agg, tot_i, tot_s = None, {}, 0, 0
for r in rows[2:]:
    if len(r) <= isamp or not r[ia].startswith("0x"):
        continue
    a = int(r[ia], 16)
    if base is None:
        base = a
    key = line_of.get(a - base)
    if key and key[0] == base_name:
        lab = label.get(key[1], "?")
    elif key:
        lab = key[0]
    else:
        lab = "?"
    n, t, s = int(r[ii]), int(r[it]), int(r[isamp])
    d = agg.setdefault(lab, [0, 0, 0])
    d[0] += n; d[1] += t; d[2] += s
    tot_i += n; tot_s += s
print("total warp-instr %d, samples %d" % (tot_i, tot_s))
print("%-44s %12s %6s %8s %6s  %s" % ("function / phase", "warp-instr", "%", "samples", "%", "avg thr"))
for lab, d in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print("%-44s %12d %6.2f %8d %6.2f  %5.1f" % (lab, d[0], 100.0 * d[0] / tot_i, d[2], 100.0 * d[2] / max(tot_s, 1), d[1] / max(d[0], 1)))
