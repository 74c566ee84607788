#!/bin/bash
# GPU run of the band-owner backward: parity tests, then timing of band (4/8/12/16 warps) against phase
set -x
python -m pytest tests/test_gpu_roi.py -x -q -m gpu -k "backward or config2 or random_shapes" 2>&1 | tail -15
for nb in 4 8 12 16; do I2V_BAND_WARPS=$nb python profiles/bench_bwd.py band; done
python profiles/bench_bwd.py phase
