#!/usr/bin/env bash
# One-GPU session: parity of every kernel variant (incl. the 512-thread, one-CTA-per-SM ones), their timings at
# cfg2 / cfg3 size, and the A/B of the two placement kernels at 6.42 M blobs.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_matvec.py tests/test_gpu_two_rhs.py tests/test_gpu_rigid.py -m gpu -q -x \
  -k "variant or both_kernels or ragged or golden or partial_products or placement or blob_positions or reference_members" 2>&1 | tail -15 > gpurun_out/pytest_variants_r02p.log
tail -3 gpurun_out/pytest_variants_r02p.log
python tools/gpu_probe.py --reps 3 --precisions single --walls 1,0 --variants all > gpurun_out/probe_r02p.jsonl 2> gpurun_out/probe_r02p.err
grep matvec gpurun_out/probe_r02p.jsonl | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['wall'], d['T'], d['threads'], d['rc'], d['ms_kernel'], d['gpairs_s'])"
python tools/sym2_sweep.py cfg3 > gpurun_out/sym2_r02p.jsonl 2> gpurun_out/sym2_r02p.err
python -c "
import json
for l in open('gpurun_out/sym2_r02p.jsonl'):
    d = json.loads(l)
    if 'variant' in d: print(d['precision'], d['variant'], round(d['two_rhs_ms'], 2), round(d['two_single_ms'], 2), d['rel_diff'])
    else: print(l.strip()[:300])"
for v in 0 1; do RBL_PLACE_VARIANT=$v python tools/on_kernels_bw.py 2>/dev/null | grep -E "place_blobs|fill" > gpurun_out/place_variant_$v.jsonl; done
python -c "
import json
for v in (0, 1):
    for l in open(f'gpurun_out/place_variant_{v}.jsonl'):
        d = json.loads(l); print(v, d['kernel'], d['precision'], round(d['ms'] * 1e3, 1), 'us', round(d['achieved_gbs']), 'GB/s')"
