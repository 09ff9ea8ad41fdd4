#!/bin/bash
# Round evidence on one B200 (run through gpurun): bench line, reference arm, ncu launch list, ncu --set full of the
# extraction kernels and of the bag-of-words search.  usage: tools/collect_evidence.sh <tag>
tag=${1:-r1_x}
out=gpurun_out
python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err || exit 1
python bench.py --impl reference > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/${tag}_launches_bench_steps2.csv \
    python bench.py --steps 2 --warmup 1 > $out/${tag}_ncu_launches.log 2>&1
python tools/profile_batch.py 128 > $out/${tag}_profile_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_import|k_resize_tma|k_fast_blur|k_octree|k_plan|k_describe" \
    --launch-skip 12 -c 12 -f -o $out/${tag}_extraction python tools/profile_batch.py 128 > $out/${tag}_ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_search_by_bow|k_bow_descend|k_bow_frame" --launch-skip 3 -c 3 -f \
    -o $out/${tag}_bow python bench.py --steps 1 --warmup 3 > $out/${tag}_ncu_bow.log 2>&1
ls -la $out
