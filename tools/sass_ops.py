#!/usr/bin/env python3
"""Static opcode histogram of kernels in a built library (no GPU): tools/sass_ops.py [lib.so] kernel_substr [...]"""
import collections, os, re, subprocess, sys, tempfile
args = sys.argv[1:]
lib = os.path.abspath(args.pop(0)) if args and args[0].endswith(".so") else os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "orb_slam_fusion_b200", "liborbx_b200.so")
with tempfile.TemporaryDirectory() as td:
    subprocess.check_call(["cuobjdump", "-xelf", "all", lib], cwd=td, stdout=subprocess.DEVNULL)
    cubin = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", cubin], capture_output=True, text=True).stdout
fn, per = None, collections.OrderedDict()
for l in sass.splitlines():
    m = re.match(r"\.text\.(\S+):", l)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m and fn:
        t = re.sub(r"^@!?U?P\d+\s+", "", m.group(2))
        per.setdefault(fn, collections.Counter())[t.split()[0].split(".")[0]] += 1
for want in args:
    for name, c in per.items():
        if want in name:
            uni = sum(v for k, v in c.items() if k.startswith("U") or k in ("S2UR", "R2UR", "LDCU"))
            print("%s: %d instructions, %d uniform-datapath (%s)" % (name[:60], sum(c.values()), uni,
                  ", ".join("%s %d" % (k, c[k]) for k in ("UMOV", "UIADD3", "ULEA", "S2UR", "LDCU", "UISETP", "ULOP3", "USHF") if c[k])))
            print("   " + ", ".join("%s %d" % kv for kv in c.most_common(16)))
