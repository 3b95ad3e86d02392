#!/bin/bash
# sparse push against copy-engine pushes (N ranks, weak scaling, final kernels)
cd "$(dirname "$0")/.."
O=gpurun_out
N=${N:-2}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
X="--no-cpu-baseline --no-configs --no-numpy-ref --gather window-copy"
for mode in ${MODES:-sparse copy}; do
  timeout 400 $TR bench.py --gpus $N --steps 30 --warmup 3 $X --push $mode > $O/r2c34_${mode}$N.json 2> $O/r2c34_${mode}$N.err
  python - $mode $N <<'PY'
import json,sys
mode,n=sys.argv[1],sys.argv[2]
try:
    d=json.loads([l for l in open(f"gpurun_out/r2c34_{mode}{n}.json").read().splitlines() if l.startswith("{")][-1])
    print(mode, n, "value", round(d["value"]), "ms/step", round(d["ms_per_step"],3), d.get("byte_check"), d["config"]["parallelism"][:110])
except Exception as e:
    print(mode, n, "failed", e); print(open(f"gpurun_out/r2c34_{mode}{n}.err").read()[-2500:])
PY
done
