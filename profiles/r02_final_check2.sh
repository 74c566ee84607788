#!/bin/bash
# last check of the round: the whole GPU suite, smoke(), both bench arms with the driver's defaults
python -m pytest tests -q -m gpu 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/final_bench.err
python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
