#!/usr/bin/env bash
# N-GPU session for the peer-memory exchange (usage: run_peer_n2.sh N ["pytest -k expression"]): partitioned
# tests (peer exchange required), then the bench line at N (matvec leg through the library's collective
# operator + BD leg).
mkdir -p gpurun_out
export RBL_REQUIRE_PEER=${RBL_REQUIRE_PEER:-1}
N=${1:-2}
K=${2:-"world1 or ${N}-"}
timeout 500 python -m pytest tests/test_gpu_partitioned.py -m gpu -q -x -k "$K" 2>&1 | tail -80 > gpurun_out/pytest_peer_n${N}.log
tail -5 gpurun_out/pytest_peer_n${N}.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N} --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29611 bench.py --gpus ${N} --steps 20 --warmup 3 > gpurun_out/bench_peer_cfg2_n${N}.json 2> gpurun_out/bench_peer_cfg2_n${N}.err
echo "bench rc=$?"; tail -c 600 gpurun_out/bench_peer_cfg2_n${N}.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_peer_cfg2_n${N}.json").read().strip().splitlines()[-1])
    print({k: d.get(k) for k in ("value", "ms_per_step", "comm_ms_per_step", "exchange")})
    print("f64", {k: d["f64"].get(k) for k in ("value", "ms_per_step", "comm_ms_per_step", "exchange")})
    print("parity", d["parity"]["rel_err"], d["f64"]["parity"]["rel_err"])
    print("bd", {k: (v.get("seconds_per_step"), v.get("gmres_iterations")) for k, v in d["bd_step"].items() if isinstance(v, dict)})
except Exception as e:
    print("no bench line:", e)
PY
