#!/bin/bash
ncu --set full --clock-control none --import-source on -k regex:lattice_fwd_chan -s 2 -c 1 -o gpurun_out/r02_fwd_chan -f python profiles/bench_fwd.py chan > gpurun_out/r02_fwd_chan_ncu.log 2>&1
ncu -i gpurun_out/r02_fwd_chan.ncu-rep --page details 2>&1 | grep -E "Duration|Throughput|Issue|Eligible|Ipc|Executed Inst|L1/TEX|Registers|Warp Cycles|Active Warps|Bank|Wavefronts|Stall" | head -40
ncu -i gpurun_out/r02_fwd_chan.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r02_fwd_chan_raw.csv
