#!/bin/bash
# the driver's round-end sequence on one GPU: GPU tests, smoke(), bench (both arms)
python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
python - <<'PY'
import json
l=json.loads(open("gpurun_out/r02_bench_final.json").read().strip().splitlines()[-1])
print({k:l[k] for k in ("value","ms_per_step","gpu_launches")}, l["e2e"]["value"], l["roofline"]["frac"], l["roofline"]["other"]["roi_align_fwd"]["frac"], l["config"])
PY
