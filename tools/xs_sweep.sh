#!/bin/bash
# x-stage run-length sweep (1 GPU): step / element-kernel / gs-pass times at 64^3, lx = 8
out=gpurun_out/${1:-r02c}_xs_sweep.jsonl
: > $out
run() { env "$@" python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.readline()); r = l['roofline']
print(json.dumps({'env': '$*', 'gdofs': round(l['value'], 3), 'ms_step': round(l['ms_per_step'], 4), 'kernel_ms': round(r['kernel_ms'], 4), 'gs_ms': round(r['gs_ms'], 4), 'classes_in_pass': r.get('gs_classes_in_pass'), 'parity': l.get('parity', {}).get('rel_l2_f')}))" >> $out; }
run B200_XSTAGE=0
run B200_XSTAGE=1
run B200_XSTAGE=2
cat $out
