#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -30 > $O/r2c20_pytest.txt
cat $O/r2c20_pytest.txt
echo "== default"; timeout 300 python tools/profile_step.py 64 4 diablo 2>&1 | tail -3
echo "== B2R_CLIP_ELIDE=0"; B2R_CLIP_ELIDE=0 timeout 300 python tools/profile_step.py 64 4 diablo 2>&1 | tail -3
echo "== torus default"; timeout 300 python tools/profile_step.py 16 3 torus1m 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k 'regex:k_' -s 16 -c 16 --csv --log-file $O/r2c20_launches.csv python tools/profile_step.py 64 3 diablo > $O/r2c20_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2c20_launches.csv')) if len(r)>10]
h=rows[0]; ik=h.index('Kernel Name'); im=h.index('Metric Name'); iv=h.index('Metric Value')
for r in rows[1:]:
    print(r[ik][:40], r[im], r[iv])
PY
