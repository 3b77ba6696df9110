#!/bin/bash
# BASELINE configs[4] on N GPUs: polynomial-order sweep lx = 5..10 at ~1e8 DOF in total (weak bricks per GPU)
# usage: tools/lxsweep_mgpu.sh <ngpus> <out.jsonl>
N=${1:-8}; out=${2:-gpurun_out/lxsweep_n$N.jsonl}
: > $out
port=29700
for lx in 5 6 7 8 9 10; do
  ne=$(python -c "print(max(2, round((1.0e8 / $N / $lx**3) ** (1.0/3.0))))")
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port \
    bench.py --gpus $N --lx $lx --ne $ne --steps 20 --warmup 5 --no-e2e 2>/dev/null | grep '^{' >> $out
done
python - "$out" <<'PY'
import json, sys
for l in open(sys.argv[1]):
    d = json.loads(l); r = d["roofline"]
    print(d["config"]["lx"], d["config"]["elements_per_gpu"], "elements/GPU:", round(d["value"], 2), "GDOF/s,",
          round(d["ms_per_step"], 3), "ms/step, element kernel", round(r["achieved"]), "GB/s =", round(r["frac"], 3),
          "of peak, parity", d["parity"]["rel_l2_f"])
PY
