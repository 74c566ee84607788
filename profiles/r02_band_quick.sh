#!/bin/bash
for nb in ${NBS:-8 12}; do I2V_BAND_WARPS=$nb python profiles/bench_bwd.py ${IMPLS:-band}; done
