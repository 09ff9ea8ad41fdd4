#!/bin/bash
# A/B of build variants (liborbx_b200_<name>.so next to the library) and environment knobs: per-stage ms of a 512-frame batch
cd /root/repo
echo base; python tools/stage_times.py
for n in "$@"; do echo $n; ORBX_LIB=/root/repo/orb_slam_fusion_b200/liborbx_b200_$n.so python tools/stage_times.py 2>&1 | tail -1; done
