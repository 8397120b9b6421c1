#!/bin/bash
# round-end profile artefacts (run on the GPU box, after the plain runs have exited 0): launch lists, DRAM traffic of the headline
# kernels, one ncu --set full capture of the fused reverse sweep at T = 3.  Numbers printed under ncu are never bench values.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_r02_cfg5_default.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_launch_cfg5.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_r02_cfg5_t3.csv python bench.py --workload cfg5_t3 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_r02_cfg1.csv python bench.py --workload cfg1_rbf_d6_m100_t16_rk4 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/launches_r02_cfg2.csv python bench.py --workload cfg2_df_d6_m100_t16_rk4 --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
$NCU --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:k_rollout -c 4 --csv --log-file gpurun_out/traffic_r02_cfg5_default.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_rollout_bwd -c 1 -o gpurun_out/bwd_tc_r02_final python bench.py --workload cfg5_t3 --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_rollout_fwd -c 1 -o gpurun_out/fwd_tc_r02_final python bench.py --workload cfg5_t3 --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1
ls -la gpurun_out
