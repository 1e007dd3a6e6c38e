#!/bin/bash
# Round-2 profile: ncu launch list of the bench command, one `ncu --set full` capture per top
# kernel (condensed on the box into text; the .ncu-rep files are dropped when they would not
# fit the 64 MiB return limit), and the DRAM traffic of one window.
# Usage: bash tools/profile_round2.sh TAG
tag=${1:-r2}
out=gpurun_out
mkdir -p $out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-cobaya-run"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 150 -c 400 --csv --log-file $out/launches_$tag.csv $B > $out/ncu_l.log 2>&1
cap() {  # name regex skip command...
  name=$1; rx=$2; skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 -o $out/prof_${name}_$tag -f "$@" > $out/ncu_$name.log 2>&1
  python tools/ncu_summary.py $out/prof_${name}_$tag.ncu-rep "$name ($tag)" "$*" > $out/${name}_${tag}_ncu.txt 2>&1
  python tools/sass_hot.py $out/prof_${name}_$tag.ncu-rep 40 > $out/${name}_${tag}_sass.txt 2>&1
}
cap step_pc2 k_step_pc2 12 $B
cap basis_wy k_basis_wy 12 $B
cap normals k_normals 12 $B
cap moments k_task_moments 1 $B
cap rowgemm k_rowgemm 2 python tools/dim_sweep.py --cells 512x8192 --cycles 1
cap basis_wy_big k_basis_wy_big 1 python tools/dim_sweep.py --cells 512x8192 --cycles 1
cap rows_gather k_rows_gather 2 $B
# SASS evidence: which tensor / TMA / async-copy instructions each kernel holds
cuobjdump -sass cobaya_b200/lib/libcobaya_b200.so | awk '/Function :/ {f=$3} /DMMA|UBLKCP|LDGSTS|SYNCS|UTMALDG|UTCHMMA|HMMA/ {c[f" "$2]++} END {for (k in c) print c[k], k}' | sed 's/([^)]*)//' | sort -k2 | awk '{n[$2" "$3]+=$1} END {for (k in n) print n[k], k}' | sort -k2 > $out/sass_evidence_$tag.txt 2>&1
while [ $(du -sm $out | cut -f1) -gt 55 ]; do
  f=$(ls -S $out/*.ncu-rep 2>/dev/null | head -1); [ -z "$f" ] && break; rm -f "$f"
done
ls -la $out | head -50
