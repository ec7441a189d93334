#!/bin/bash
# One single-GPU profiling pass of the round-2 kernels (run under gpurun): plain bench first, then the ncu passes.
set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err || exit 1
python tools/bench_gan_step.py --disc cuda > gpurun_out/r02_gan_step_n1.json 2> gpurun_out/r02_gan_n1.err
python tools/bench_gan_step.py --disc torch > gpurun_out/r02_gan_step_n1_torchD.json 2>> gpurun_out/r02_gan_n1.err
B="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches.csv $B > gpurun_out/r02_ncu_ll.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -c 1200 --csv --log-file gpurun_out/r02_dram.csv $B > gpurun_out/r02_ncu_dram.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:dense_block_kernel -s 30 -c 2 -o gpurun_out/r02_dense_full $B > gpurun_out/r02_ncu_full_dense.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 60 -c 6 -o gpurun_out/r02_conv_full $B > gpurun_out/r02_ncu_full_conv.log 2>&1
B2="python bench.py --mode train --workload cfg3 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_train_cfg3.csv $B2 > gpurun_out/r02_ncu_ll_train.log 2>&1
ls -la gpurun_out | tail -20
