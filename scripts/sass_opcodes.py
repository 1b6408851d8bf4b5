"""Per-kernel SASS opcode summary of the shipped library ("commit a per-kernel SASS opcode summary under profiles/").
usage: python scripts/sass_opcodes.py [lib.so] > profiles/rNN_sass_opcodes.txt"""
import collections, os, re, subprocess, sys
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "real_robots_b200", "libreal_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, counts = None, collections.OrderedDict()
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", l)
    if m and kern:
        counts[kern][m.group(1)] += 1
print("# SASS opcode summary of %s (cuobjdump -sass, sm_100a).  Per kernel: instruction count, then the opcodes that matter for the" % os.path.basename(lib))
print("# questions asked of this code: fp64 / fp32 arithmetic, shuffles / votes / redux, shared / global / local memory, 128-bit streaming")
print("# stores, atomics, barriers, and the TMA / bulk-copy / tensor-core families (UBLKCP, UTMALDG, SYNCS, UTCHMMA/HMMA: absent =")
print("# not used; nothing on this path is a dense contraction, and the one TMA candidate measured slower — profiles/r02_raster_variants.txt).")
fam = [("fp64", r"^D(FMA|ADD|MUL|SETP|MNMX)"), ("fp32", r"^F(FMA|ADD|MUL|SETP|MNMX|SEL|CHK)"), ("int/logic", r"^(IMAD|IADD3|LOP3|SHF|LEA|ISETP|VIADD|SEL|PRMT|MOV|IABS|POPC|FLO|BREV|I2F|F2I|F2F|I2I|VIMNMX)"),
       ("mufu", r"^MUFU"), ("shuffle/vote/redux", r"^(SHFL|VOTE|REDUX|MATCH|WARPSYNC)"), ("lds/sts", r"^(LDS|STS)"), ("atoms", r"^ATOMS"), ("ldg/ld", r"^(LDG|LD)\b"),
       ("stg/st", r"^(STG|ST)\b"), ("stg.128 streaming (EF)", r"^STG\.E\.EF\.128"), ("local ldl/stl", r"^(LDL|STL)"), ("atomg/red", r"^(ATOMG|RED)"), ("bar", r"^BAR"),
       ("branch", r"^(BRA|BSSY|BSYNC|EXIT|CALL|RET)"), ("prefetch (CCTL)", r"^CCTL"), ("bulk/TMA (UBLKCP|UTMALDG|UTMASTG|SYNCS)", r"^(UBLKCP|UTMALDG|UTMASTG|SYNCS)"), ("tensor (HMMA|UTCHMMA|UTCMMA|IMMA)", r"^(HMMA|UTC|IMMA|QMMA)")]
for k, c in counts.items():
    tot = sum(c.values())
    print("\n%s: %d instructions" % (k, tot))
    row = []
    for name, pat in fam:
        n = sum(v for op, v in c.items() if re.match(pat, op))
        row.append("%s %d" % (name, n))
    print("   " + ";  ".join(row))
    print("   top: " + ", ".join("%s %d" % (op, v) for op, v in c.most_common(12)))
