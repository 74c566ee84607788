#!/bin/bash
# Round-2 evidence: the bench line, the ncu launch list of the same command, and one `ncu --set full` capture each of the
# dominant backward kernel (phase), its band-owner alternative and the forward slab kernel.  Run under gpurun.
set -x
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs --no-projection > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lattice_bwd_phase -s 2 -c 1 -o gpurun_out/r02_bwd_phase -f \
    python profiles/bench_bwd.py phase > gpurun_out/r02_ncu_phase.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:lattice_fwd_slab -s 2 -c 1 -o gpurun_out/r02_fwd_slab -f \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-configs --no-projection > gpurun_out/r02_ncu_fwd.log 2>&1
for r in r02_bwd_phase r02_fwd_slab; do
  ncu -i gpurun_out/$r.ncu-rep --page details 2>&1 | grep -E "Duration|Throughput|Issue|Eligible|Ipc|Executed Inst|L1/TEX|Registers|Warp Cycles|Active Warps|DRAM|Bank|Shared" > gpurun_out/$r.details.txt
done
