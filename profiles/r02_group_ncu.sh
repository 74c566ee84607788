#!/bin/bash
# ncu --set full of the two largest non-fc6 kernels of a relation-stage frame group (4 frames x 64 detections)
ncu --set full --clock-control none --import-source on -k regex:"roi_pool_plane_bf16|linear_tcgen05_kernel|pair_conv1|rel_score" -s 8 -c 8 -o gpurun_out/r02_group -f python profiles/run_group.py > gpurun_out/r02_group_ncu.log 2>&1
ncu -i gpurun_out/r02_group.ncu-rep --page details 2>&1 | grep -E "^  [a-zA-Z_:<>]+.*\(|Duration|Issue Slots Busy|Executed Ipc Active|Executed Instructions|Registers Per|L1/TEX Cache Throughput|DRAM Throughput|Compute \(SM\) Throughput|Memory Throughput|Achieved Active Warps|Block Size|Grid Size" | head -150
ncu -i gpurun_out/r02_group.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/r02_group_raw.csv
