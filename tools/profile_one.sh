#!/bin/bash
# One `ncu --set full` capture of one kernel of the bench command, condensed on the box.
# Usage: bash tools/profile_one.sh NAME REGEX SKIP TAG [bench args...]
name=$1; rx=$2; skip=$3; tag=$4; shift 4
out=gpurun_out
mkdir -p $out
ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c 1 \
    -o $out/prof_${name}_$tag -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-cobaya-run "$@" \
    > $out/ncu_$name.log 2>&1
python tools/ncu_summary.py $out/prof_${name}_$tag.ncu-rep "$name ($tag)" "bench.py --steps 1 --warmup 3 $*" > $out/${name}_${tag}_ncu.txt 2>&1
python tools/sass_hot.py $out/prof_${name}_$tag.ncu-rep 70 > $out/${name}_${tag}_sass.txt 2>&1
# DROP_REP=1: keep only the condensed summaries (gpurun_out/ is limited to 64 MiB per call)
[ -n "$DROP_REP" ] && rm -f $out/prof_${name}_$tag.ncu-rep
