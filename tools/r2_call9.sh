#!/bin/bash
# round 2, GPU call 9 (8 GPUs): weak / strong / band scaling at N = 8 with the byte check of every rank's frames
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 8 --steps 30 --warmup 3 > $O/r2c9_weak8.json 2> $O/r2c9_weak8.err
timeout 900 $TR bench.py --gpus 8 --steps 4 --warmup 3 --workload torus1m --scaling strong --frames 256 > $O/r2c9_strong8_torus.json 2> $O/r2c9_strong8_torus.err
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 3 --split bands --views 16 > $O/r2c9_bands8.json 2> $O/r2c9_bands8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > $O/r2c9_weak2.json 2> $O/r2c9_weak2.err
for f in weak8 strong8_torus bands8 weak2; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/r2c9_{f}.json").read().splitlines() if l.startswith("{")][-1])
    print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],3), d["scaling"], d.get("byte_check"), "pcie", round(d["e2e"]["pcie_gbs"],1), d["config"]["parallelism"][:90])
except Exception as e:
    print(f, "failed", e); print(open(f"gpurun_out/r2c9_{f}.err").read()[-1500:])
PY
done
