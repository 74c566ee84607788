#!/bin/bash
# ncu --set full of the proposal stage's per-frame kernels (order, NMS) at 32 frames
python profiles/bench_prop.py 32
ncu --set full --clock-control none --import-source on -k regex:"segment_sort|nms_scan|nms_relay|segment_topk" -s 8 -c 2 -o gpurun_out/r02_prop -f python profiles/bench_prop.py 32 > gpurun_out/r02_prop_ncu.log 2>&1
ncu -i gpurun_out/r02_prop.ncu-rep --page details 2>&1 | grep -E "^  [a-z_]+.*\(|Duration|Issue Slots|Executed Ipc Active|Executed Inst|Registers|Warp Cycles Per Issued|Active Warps" | head -60
ncu -i gpurun_out/r02_prop.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r02_prop_raw.csv
