#!/usr/bin/env bash
# 8-GPU measurement session (one gpurun call): partitioned tests at every world size, then the
# BASELINE.json configurations that name 8 GPUs.  Every output lands in gpurun_out/.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_partitioned.py -m gpu -q 2>&1 | tail -60 > gpurun_out/pytest_part_r02_n8.log
tail -3 gpurun_out/pytest_part_r02_n8.log
timeout 300 $TR --master-port 29601 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r02_cfg2_n8.json 2> gpurun_out/bench_r02_cfg2_n8.err
tail -c 200 gpurun_out/bench_r02_cfg2_n8.err
timeout 300 $TR --master-port 29602 bench.py --gpus 8 --steps 5 --warmup 3 --workload cfg4 --bd-steps 0 > gpurun_out/bench_r02_cfg4_n8.json 2> gpurun_out/bench_r02_cfg4_n8.err
tail -c 200 gpurun_out/bench_r02_cfg4_n8.err
timeout 500 $TR --master-port 29603 bench.py --gpus 8 --steps 1 --warmup 3 --e2e-warmup 1 --workload cfg5 --bd-steps 0 > gpurun_out/bench_r02_cfg5_n8.json 2> gpurun_out/bench_r02_cfg5_n8.err
tail -c 200 gpurun_out/bench_r02_cfg5_n8.err
if [ "${RBL_CFG5_BD:-0}" = "1" ]; then
  timeout 700 $TR --master-port 29604 bench.py --gpus 8 --steps 3 --warmup 3 --workload small --dtype single --no-cpu-baseline \
      --bd-workload cfg5 --bd-steps 1 --bd-warmup 0 --bd-profile-step 0 > gpurun_out/bench_r02_cfg5_bd_n8.json 2> gpurun_out/bench_r02_cfg5_bd_n8.err
  tail -c 200 gpurun_out/bench_r02_cfg5_bd_n8.err
fi
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | head -8
