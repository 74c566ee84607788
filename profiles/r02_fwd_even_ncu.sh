#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:lattice_fwd_even -s 2 -c 1 -o gpurun_out/r02_fwd_even -f python profiles/bench_fwd.py even > gpurun_out/r02_fwd_even_ncu.log 2>&1
ncu -i gpurun_out/r02_fwd_even.ncu-rep --page details 2>&1 | grep -E "Duration|Throughput|Issue|Eligible|Ipc|Executed Inst|L1/TEX|Registers|Warp Cycles|Active Warps" | head -30
