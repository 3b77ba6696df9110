#!/bin/bash
# step time against the length of the timed region (the board's 1 kW power cap lowers the SM clock in long runs)
for s in ${@:-10 20 80}; do python bench.py --steps $s --warmup 5 --no-e2e --no-cpu 2>/dev/null | python -c "
import sys, json
l = json.loads(sys.stdin.readline()); r = l['roofline']
print('steps', l['steps'], 'ms_step', round(l['ms_per_step'], 4), 'GDOF/s', round(l['value'], 2), 'kernel', round(r['kernel_ms'], 4), 'gs', round(r['gs_ms'], 4), l['clocks'], l['parity']['rel_l2_f'])"; done
