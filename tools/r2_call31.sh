#!/bin/bash
# round 2, 8 GPUs, final kernels: weak / strong (config 5) / row-band scaling with the byte check of every rank's frames
cd "$(dirname "$0")/.."
O=gpurun_out
N=${N:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
X="--no-cpu-baseline --no-configs --no-numpy-ref"
timeout 400 $TR bench.py --gpus $N --steps 30 --warmup 3 $X > $O/r2c31_weak$N.json 2> $O/r2c31_weak$N.err
timeout 500 $TR bench.py --gpus $N --steps 4 --warmup 3 --workload torus1m --scaling strong --frames 256 $X > $O/r2c31_strong${N}_torus.json 2> $O/r2c31_strong${N}_torus.err
timeout 400 $TR bench.py --gpus $N --steps 20 --warmup 3 --split bands --views 16 $X > $O/r2c31_bands$N.json 2> $O/r2c31_bands$N.err
timeout 300 python bench.py --gpus 1 --steps 4 --warmup 3 --workload torus1m --scaling strong --frames 256 $X > $O/r2c31_strong1_torus.json 2> $O/r2c31_strong1_torus.err
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --split bands --views 16 $X > $O/r2c31_bands1.json 2> $O/r2c31_bands1.err
for f in weak$N strong${N}_torus bands$N strong1_torus bands1; do python - $f <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads([l for l in open(f"gpurun_out/r2c31_{f}.json").read().splitlines() if l.startswith("{")][-1])
    print(f, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],3), d["scaling"], d.get("byte_check"), "pcie", round(d["e2e"]["pcie_gbs"],1), d["e2e"].get("pcie_frac"), d["config"]["parallelism"][:80])
except Exception as e:
    print(f, "failed", e); print(open(f"gpurun_out/r2c31_{f}.err").read()[-1500:])
PY
done
