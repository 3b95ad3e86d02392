#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "production_stencil or fused" 2>&1 | tail -30 > $O/r2c19_pytest.txt
cat $O/r2c19_pytest.txt
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k 'regex:k_' -s 15 -c 15 --csv --log-file $O/r2c19_launches.csv python tools/profile_step.py 64 3 diablo > $O/r2c19_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c19_launches.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value')
for r in rows[1:]:
    print(r[ik][:60], r[im], r[iv])
PY
