#!/bin/bash
# whole GPU suite + the config-3 frame timing
python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python profiles/run_vrd.py 2>&1 | tail -5
