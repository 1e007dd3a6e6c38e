#!/bin/bash
# Round profile: ncu launch lists of the bench command and of the 128-D sweep cell, one
# `ncu --set full` capture per top kernel, condensed on the box into text (the .ncu-rep files
# are dropped when they would not fit the 64 MiB return limit).  Usage: bash tools/profile_round.sh TAG
tag=${1:-vXX}
out=gpurun_out
mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 600 --csv --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/ncu_l.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $out/launches128_$tag.csv python tools/dim_sweep.py --cells 128x8192 --cycles 2 > $out/ncu_l128.log 2>&1
cap() {  # name regex skip command...
  name=$1; rx=$2; skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o $out/prof_${name}_$tag -f "$@" > $out/ncu_$name.log 2>&1
  python tools/ncu_summary.py $out/prof_${name}_$tag.ncu-rep "$name ($tag)" "$*" > $out/${name}_${tag}_ncu.txt 2>&1
  python tools/sass_hot.py $out/prof_${name}_$tag.ncu-rep 40 > $out/${name}_${tag}_sass.txt 2>&1
}
cap step_pc k_step_pc 40 python bench.py --steps 1 --warmup 3 --no-cpu-baseline
cap basis_wy k_basis_wy 40 python bench.py --steps 1 --warmup 3 --no-cpu-baseline
cap normals k_normals 40 python bench.py --steps 1 --warmup 3 --no-cpu-baseline
cap stream_products k_stream_products 2 python tools/dim_sweep.py --cells 128x8192 --cycles 2
cap stream_accept k_stream_accept 2 python tools/dim_sweep.py --cells 128x8192 --cycles 2
cap basis_wy128 k_basis_wy 2 python tools/dim_sweep.py --cells 128x8192 --cycles 2
# keep the return under the limit: drop the largest reports first
while [ $(du -sm $out | cut -f1) -gt 55 ]; do
  f=$(ls -S $out/*.ncu-rep 2>/dev/null | head -1); [ -z "$f" ] && break; rm -f "$f"
done
ls -la $out
