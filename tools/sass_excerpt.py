#!/usr/bin/env python3
"""Trimmed SASS evidence of the built library (cuobjdump + nvdisasm, no GPU needed): per kernel the static opcode histogram
and the instructions that prove the hardware paths the design claims -- UTMALDG (TMA tensor copies), SYNCS (mbarrier),
IDP (DP4A / DP2A), VIMNMX3 (16-bit SIMD min/max), VABSDIFF4, POPC, REDUX, LDGSTS.
usage: tools/sass_excerpt.py [liborbx_b200.so] > profiles/<round>_sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys
import tempfile

lib = os.path.abspath(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "orb_slam_fusion_b200", "liborbx_b200.so"))
KERNELS = ["k_import", "k_zero_counters", "k_resize_tma", "k_fast_blur", "k_octree", "k_plan", "k_describe", "k_knn2", "k_knn2_merge_keys", "k_top2_keys_ratio",
           "k_stereo_rowband", "k_stereo_refine", "k_window_topk", "k_projection_claim", "k_search_by_bow", "k_bow_descend"]
PROOF = re.compile(r"^(UTMALDG|SYNCS|IDP|VIMNMX3|VABSDIFF4|POPC|REDUX|LDGSTS|UBLKCP|BAR)")
with tempfile.TemporaryDirectory() as td:
    subprocess.check_call(["cuobjdump", "-xelf", "all", lib], cwd=td, stdout=subprocess.DEVNULL)
    cubin = [os.path.join(td, f) for f in os.listdir(td) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", cubin], capture_output=True, text=True).stdout
fn, per = None, collections.OrderedDict()
for l in sass.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", l) or re.match(r"\.text\.(\S+):", l)
    if m:
        fn = m.group(1)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
    if m and fn:
        per.setdefault(fn, []).append((m.group(1), m.group(2).strip()))
print("# SASS excerpt of %s (sm_100a), static counts" % os.path.basename(lib))
for want in KERNELS:
    for name, ins in per.items():
        dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
        short = dem.split("(")[0].split("::")[-1]
        if short.split("<")[0] != want:
            continue
        ops = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in ins)
        print("\n## %s   [%d instructions]" % (dem.split("(")[0], len(ins)))
        print("   " + ", ".join("%s %d" % kv for kv in ops.most_common(14)))
        seen = collections.Counter()
        for off, t in ins:
            t2 = re.sub(r"^@!?U?P\d+\s+", "", t)
            m = PROOF.match(t2)
            if m and seen[t2.split()[0]] < 2:
                seen[t2.split()[0]] += 1
                print("   /*%s*/ %s" % (off, t))
