#!/bin/bash
# run bench.py once per library variant; print value and per-kernel launch times
for v in "$@"; do
  COBAYA_B200_LIB=$PWD/cobaya_b200/lib/variants/$v.so timeout 200 python bench.py --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); a=d['roofline']['avg_launch_ms']; print('$v', '%.4g'%d['value'], 'basis %.4f step %.4f tape %.4f'%(a['basis'],a['step'],a['tape']), 'e2e %.4g'%d['e2e']['value'])
"
done
