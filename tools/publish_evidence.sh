#!/bin/bash
# Turn the artifacts tools/collect_evidence.sh <tag> left in gpurun_out/ into the tracked summaries under profiles/
# (run in the dev container; needs ncu for reading the reports, no GPU).  usage: tools/publish_evidence.sh <tag>
set -e
tag=${1:?tag}
cd "$(dirname "$0")/.."
o=gpurun_out p=profiles
cp $o/${tag}_bench.json $p/${tag}_bench.json
cp $o/${tag}_bench_reference_arm.json $p/${tag}_bench_reference_arm.json
cp $o/${tag}_launches_bench_steps2.csv $p/${tag}_launches_bench_steps2.csv
{
  echo "# ncu --set full --clock-control none --import-source on, one 128-frame batch of config 1 (tools/profile_batch.py 128), B200, round-2 code (tools/collect_evidence.sh ${tag}); counters by tools/ncu_summary.py"
  python tools/ncu_summary.py $o/${tag}_extraction.ncu-rep
  echo
  echo "# executed warp-instructions by opcode per kernel (ncu source page, tools/ncu_opmix.py)"
  ncu -i $o/${tag}_extraction.ncu-rep --page source --csv --print-source sass > /tmp/${tag}_sass.csv 2>/dev/null
  python tools/ncu_opmix.py /tmp/${tag}_sass.csv
} > $p/${tag}_extraction_kernels_ncu_full.txt
{
  echo "# ncu --set full --clock-control none of k_knn2, 1000 queries x 10 M rows (tools/knn2_slice_time.py 10000000), round-2 code"
  python tools/ncu_summary.py $o/${tag}_knn2.ncu-rep
  echo
  echo "# the same run without ncu:"
  cat $o/${tag}_knn2_plain.log
} > $p/${tag}_knn2_ncu_full.txt
python tools/sass_excerpt.py > $p/${tag}_sass_excerpt.txt
ls -la $p | grep ${tag}
