#!/bin/bash
# Round-2 evidence, second capture (after the proposal-stage, relation-stage and pipelining changes): the full GPU test
# suite, the bench line of both arms, the ncu launch lists of the bench command and of one relation-stage frame group, and
# `ncu --set full` summaries of the kernels that changed.  Run under gpurun (one GPU).
python -m pytest tests -q -m gpu 2>&1 | tail -2 > gpurun_out/r02_pytest_gpu.txt; cat gpurun_out/r02_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err || { tail -5 gpurun_out/r02_bench_1gpu.err; exit 1; }
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs --no-projection > gpurun_out/r02_ncu_launches.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_group_launches.csv \
    python profiles/run_group.py > gpurun_out/r02_group_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"segment_topk|nms_scan" -s 8 -c 2 -o gpurun_out/r02_prop -f \
    python profiles/bench_prop.py 32 > gpurun_out/r02_prop_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"roi_pool_plane_bf16|rel_score_tile|pair_conv1|linear_tcgen05_pair_kernel<0, 4" -s 5 -c 4 -o gpurun_out/r02_group -f \
    python profiles/run_group.py > gpurun_out/r02_group_ncu.log 2>&1
for r in r02_prop r02_group; do
  ncu -i gpurun_out/$r.ncu-rep --page details 2>&1 | grep -E "^  [a-zA-Z_:<>]+.*\(|Duration|Throughput|Issue Slots|Executed Ipc Active|Executed Instructions|Registers Per|Active Warps Per SM|Block Size|Grid Size|Dynamic Shared" > gpurun_out/$r.details.txt
done
TIME=1 python profiles/run_group.py; python profiles/bench_pool_rows.py; python profiles/bench_conv2.py; python profiles/bench_prop.py 32; python profiles/bench_fwd.py slab
