#!/bin/bash
python -m pytest tests/test_gpu_roi.py -x -q -m gpu 2>&1 | tail -6
python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench1.json 2> gpurun_out/r02_bench1.err; echo "rc=$?"; tail -3 gpurun_out/r02_bench1.err
python - <<'PY'
import json
l=json.loads(open('gpurun_out/r02_bench1.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','stages_ms','parity','sgg_frame','train_bwd','clip'):
    print(k, json.dumps(l.get(k))[:600])
print('e2e', l['e2e']); print('roofline', l['roofline'])
PY
