#!/bin/bash
set -x
python -m pytest tests/test_gpu_roi.py -x -q -m gpu -k "backward or config2 or random_shapes" 2>&1 | tail -5
python profiles/bench_bwd.py phase band
