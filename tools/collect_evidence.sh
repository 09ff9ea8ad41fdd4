#!/bin/bash
# Round evidence on one B200 (run through gpurun): bench line, reference arm, ncu launch list, ncu --set full of the
# extraction kernels (with the source page for per-opcode counts) and of k_knn2.  usage: tools/collect_evidence.sh <tag>
tag=${1:-r2_x}
out=gpurun_out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
python bench.py --impl reference > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 > $out/${tag}_ncu_launches.log 2>&1
python tools/profile_batch.py 128 > $out/${tag}_profile_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_import|k_zero_counters|k_resize_tma|k_fast_blur|k_octree|k_plan|k_describe" \
    --launch-skip 12 -c 12 -f -o $out/${tag}_extraction python tools/profile_batch.py 128 > $out/${tag}_ncu_full.log 2>&1
python tools/knn2_slice_time.py 10000000 > $out/${tag}_knn2_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"k_knn2$|k_knn2\(" --launch-skip 2 -c 1 -f -o $out/${tag}_knn2 python tools/knn2_slice_time.py 10000000 > $out/${tag}_ncu_knn2.log 2>&1
ls -la $out | tail -12
