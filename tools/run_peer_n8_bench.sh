#!/usr/bin/env bash
# 8-GPU bench line with the peer-memory exchange (matvec leg self-checked against NCCL, parity against the
# oracle, BD leg).  Output: gpurun_out/bench_peer_cfg2_n8.json
mkdir -p gpurun_out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29621 \
  bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_peer_cfg2_n8.json 2> gpurun_out/bench_peer_cfg2_n8.err
echo rc=$?; tail -c 300 gpurun_out/bench_peer_cfg2_n8.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_peer_cfg2_n8.json").read().strip().splitlines()[-1])
print({k: d.get(k) for k in ("value", "ms_per_step", "comm_ms_per_step", "exchange")})
print("f64", {k: d["f64"].get(k) for k in ("value", "ms_per_step", "comm_ms_per_step", "exchange")})
print("parity", d["parity"]["rel_err"], d["f64"]["parity"]["rel_err"], "kernel_ms", d["roofline"]["kernel_ms"], d["f64"]["roofline"]["kernel_ms"])
print("bd", d["bd_step"].get("exchange"), {k: (v.get("seconds_per_step"), v.get("gmres_iterations")) for k, v in d["bd_step"].items() if isinstance(v, dict)})
PY
