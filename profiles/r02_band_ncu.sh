#!/bin/bash
# one ncu --set full capture of the band-owner backward at config 2 (I2V_BAND_WARPS from the environment)
ncu --set full --clock-control none --import-source on -k regex:lattice_bwd_band -s 2 -c 1 -o gpurun_out/r02_band_${I2V_BAND_WARPS:-8} -f python profiles/bench_bwd.py band > gpurun_out/r02_band_ncu.log 2>&1
ncu -i gpurun_out/r02_band_${I2V_BAND_WARPS:-8}.ncu-rep --page details 2>&1 | grep -E "Duration|Throughput|Issue|Eligible|IPC|Ipc|Executed Inst|L1/TEX|Registers|Warp Cycles|Active Warps" | head -40
