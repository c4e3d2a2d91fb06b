#!/bin/bash
# compares the 2- and 4-epilogue-group cluster kernels: parity tests, then bench.py's headline
timeout 400 python -m pytest tests/test_gpu_trunk.py -x -q -m gpu 2>&1 | tail -30
for g in 2 4; do
  RUMPY_B200_CLUSTER_GROUPS=$g timeout 200 python bench.py --no-extra --no-train --steps 30 2>/dev/null > gpurun_out/groups_$g.json
  python - <<PY
import json
d = json.load(open('gpurun_out/groups_$g.json'))
print('groups $g:', d['value'], 'Mpix/s', d['ms_per_step'], 'ms', d['roofline']['us_per_launch'], 'us trunk kernel')
PY
done
