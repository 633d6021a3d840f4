#!/usr/bin/env bash
# One-GPU evidence session: quick parity of the changed placement kernel, O(N) bandwidths, ncu full captures
# of the four default product kernels at 162 000 blobs, the launch list of a short bench run, cfg4 at N=1,
# cfg1 (tools/config_runs.py).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -m pytest tests/test_gpu_rigid.py tests/test_gpu_interface.py -m gpu -q -x 2>&1 | tail -3 > gpurun_out/pytest_r02n.log; tail -1 gpurun_out/pytest_r02n.log
python tools/on_kernels_bw.py > gpurun_out/on_kernels_r02n.jsonl 2>/dev/null
P="python tools/gpu_probe.py --reps 1"
$P --precisions single --walls 1 --variants 0 > gpurun_out/plain_n1.log 2>&1 && ncu --set full --clock-control none -k regex:rpy_matvec_sym_kernel -s 1 -c 1 -o gpurun_out/prof_r02_final_f32_wall $P --precisions single --walls 1 --variants 0 > gpurun_out/ncu_n1.log 2>&1
$P --precisions double --walls 1 --variants 0 > gpurun_out/plain_n2.log 2>&1 && ncu --set full --clock-control none -k regex:rpy_matvec_sym_kernel -s 1 -c 1 -o gpurun_out/prof_r02_final_f64_wall $P --precisions double --walls 1 --variants 0 > gpurun_out/ncu_n2.log 2>&1
$P --precisions single --walls 0 --variants 0 > gpurun_out/plain_n3.log 2>&1 && ncu --set full --clock-control none -k regex:rpy_matvec_sym_kernel -s 1 -c 1 -o gpurun_out/prof_r02_final_f32_free $P --precisions single --walls 0 --variants 0 > gpurun_out/ncu_n3.log 2>&1
$P --precisions double --walls 0 --variants 1 > gpurun_out/plain_n4.log 2>&1 && ncu --set full --clock-control none -k regex:rpy_matvec_sym_kernel -s 1 -c 1 -o gpurun_out/prof_r02_final_f64_free $P --precisions double --walls 0 --variants 1 > gpurun_out/ncu_n4.log 2>&1
# keep the call's output under gpurun's 64 MiB: raw metric pages as CSV, reports deleted
for f in f32_wall f64_wall f32_free f64_free; do
  ncu -i gpurun_out/prof_r02_final_$f.ncu-rep --page raw --csv > gpurun_out/ncu_r02_final_$f.raw.csv 2>/dev/null
  rm -f gpurun_out/prof_r02_final_$f.ncu-rep
done
B="python bench.py --steps 2 --warmup 3 --bd-steps 1 --bd-mixed 0 --bd-profile-step 0 --no-cpu-baseline --dtype single"
$B > gpurun_out/plain_n5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_r02_bench.csv $B > gpurun_out/ncu_n5.log 2>&1
python bench.py --steps 3 --warmup 3 --e2e-warmup 1 --workload cfg4 --bd-steps 0 > gpurun_out/bench_r02_cfg4_n1.json 2> gpurun_out/bench_r02_cfg4_n1.err
python tools/config_runs.py --which cfg1 --cpu > gpurun_out/config_runs_r02.jsonl 2> gpurun_out/config_runs_r02.err
tail -c 300 gpurun_out/config_runs_r02.err
