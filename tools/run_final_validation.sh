#!/usr/bin/env bash
# Final one-GPU validation of the round: smoke(), the whole GPU test suite, the default bench line, the O(N)
# bandwidth lines and one ncu --set full capture of the placement kernel.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 420 python -m pytest tests -m gpu -q -x 2>&1 | tail -40 > gpurun_out/pytest_gpu_r02_final.log; tail -2 gpurun_out/pytest_gpu_r02_final.log
timeout 400 python bench.py > gpurun_out/bench_r02_final_n1.json 2> gpurun_out/bench_r02_final_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
try:
    d = json.loads(open("gpurun_out/bench_r02_final_n1.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"], d["roofline"]["frac"], d["f64"]["value"], d["f64"]["roofline"]["frac"])
    print("parity", d["parity"]["rel_err"], d["f64"]["parity"]["rel_err"], "clocks", d["clocks"])
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["kind"])
    print("bd", {k: v.get("seconds_per_step") for k, v in d["bd_step"].items() if isinstance(v, dict)})
except Exception as e:
    print("no bench line:", e)
PY
timeout 120 python tools/on_kernels_bw.py > gpurun_out/on_kernels_r02_final.jsonl 2>/dev/null
timeout 200 ncu --set full --clock-control none -k regex:place_blobs_rows_kernel -c 2 -o gpurun_out/prof_r02_place_rows python tools/on_kernels_bw.py > gpurun_out/ncu_place_rows.log 2>&1
ncu -i gpurun_out/prof_r02_place_rows.ncu-rep --page raw --csv > gpurun_out/ncu_r02_place_rows.raw.csv 2>/dev/null
rm -f gpurun_out/prof_r02_place_rows.ncu-rep
grep place_blobs gpurun_out/on_kernels_r02_final.jsonl | cut -c1-200
