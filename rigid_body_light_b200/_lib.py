"""ctypes view of the C ABI (include/rbl.h) for Python hosts that hold DEVICE buffers:
bench.py, the multi-GPU sharding layer and the GPU tests.  The pybind11 host class
(c_rigid) is the reference-facing boundary; this is the same library seen through its
`extern "C"` entry points, exactly what a Go/Java/Rust host would bind.

No CPU fallback: ``load()`` raises if librbl.so is missing.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

RBL_F32, RBL_F64 = 4, 8
STATUS = {0: "OK", 1: "INVALID", 2: "CUDA", 3: "BELOW_WALL", 4: "SINGULAR", 5: "STATE", 6: "NOMEM"}

# name -> (restype, argtypes); every symbol include/rbl.h declares
_c = ctypes
_vp, _i, _d, _sz, _i64 = _c.c_void_p, _c.c_int, _c.c_double, _c.c_size_t, _c.c_int64
_pi, _pd, _pvp, _pi64 = _c.POINTER(_c.c_int), _c.POINTER(_c.c_double), _c.POINTER(_c.c_void_p), _c.POINTER(_c.c_int64)
SYMBOLS = {
    "rbl_create": (_i, [_i, _i, _pvp]),
    "rbl_destroy": (None, [_vp]),
    "rbl_last_error": (_c.c_char_p, [_vp]),
    "rbl_precision": (_i, [_vp]),
    "rbl_sm_count": (_i, [_vp]),
    "rbl_version": (_c.c_char_p, []),
    "rbl_set_parameters": (_i, [_vp, _d, _d, _d, _d, _vp, _i]),
    "rbl_set_flags": (_i, [_vp, _i, _i]),
    "rbl_set_config": (_i, [_vp, _vp, _vp, _i]),
    "rbl_get_config": (_i, [_vp, _vp, _vp]),
    "rbl_set_K_mats": (_i, [_vp]),
    "rbl_n_bodies": (_i, [_vp]),
    "rbl_blobs_per_body": (_i, [_vp]),
    "rbl_blob_positions": (_i, [_vp, _vp]),
    "rbl_K_dot": (_i, [_vp, _vp, _vp]),
    "rbl_KT_dot": (_i, [_vp, _vp, _vp]),
    "rbl_Kinv_dot": (_i, [_vp, _vp, _vp]),
    "rbl_KTinv_dot": (_i, [_vp, _vp, _vp]),
    "rbl_apply_M": (_i, [_vp, _vp, _vp, _i, _vp]),
    "rbl_apply_PC": (_i, [_vp, _vp, _vp]),
    "rbl_apply_saddle": (_i, [_vp, _vp, _vp]),
    "rbl_evolve": (_i, [_vp, _vp]),
    "rbl_export_K_csc": (_i, [_vp, _vp, _vp, _vp]),
    "rbl_export_Kinv_csc": (_i, [_vp, _vp, _vp, _vp]),
    "rbl_gmres": (_i, [_vp, _vp, _vp, _d, _i, _i, _pi, _pd]),
    "rbl_lanczos_sqrt": (_i, [_vp, _vp, _vp, _d, _i, _pi]),
    "rbl_bd_step_seeded": (_i, [_vp, _vp, _vp, _c.c_uint64, _c.c_uint64, _d, _d, _i, _i, _d, _i, _vp, _pi, _pd]),
    "rbl_normals": (_i, [_vp, _c.c_uint64, _c.c_uint64, _c.c_uint64, _sz, _vp, _vp, _vp]),
    "rbl_dev_normals": (_i, [_vp, _c.c_uint64, _c.c_uint64, _c.c_uint64, _sz, _vp, _vp, _vp]),
    "rbl_apply_M2": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "rbl_dev_apply_M2": (_i, [_vp, _vp, _vp, _vp, _i, _vp, _vp]),
    "rbl_lanczos_sqrt2": (_i, [_vp, _vp, _vp, _vp, _vp, _d, _i, _pi]),
    "rbl_set_lanczos_pairing": (_i, [_vp, _i]),
    "rbl_set_noise_preconditioner": (_i, [_vp, _i]),
    "rbl_noise_selfcheck": (_i, [_vp, _pd, _pd, _pi]),
    "rbl_num_sym2_variants": (_i, [_vp]),
    "rbl_sym2_variant_info": (_i, [_vp, _i, _pi, _pi]),
    "rbl_set_sym2_variant": (_i, [_vp, _i]),
    "rbl_bd_step": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _i, _i, _d, _i, _vp, _pi, _pd]),
    "rbl_dev_apply_M": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "rbl_dev_blob_positions": (_i, [_vp, _vp]),
    "rbl_dev_K_dot": (_i, [_vp, _vp, _vp]),
    "rbl_dev_KT_dot": (_i, [_vp, _vp, _vp]),
    "rbl_dev_apply_PC": (_i, [_vp, _vp, _vp]),
    "rbl_dev_apply_saddle": (_i, [_vp, _vp, _vp]),
    "rbl_dev_apply_saddle_shard": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "rbl_dev_apply_M_part": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "rbl_dev_saddle_finish": (_i, [_vp, _vp, _vp, _vp, _vp]),
    "rbl_sync": (_i, [_vp]),
    "rbl_comm_unique_id": (_i, [_vp]),
    "rbl_comm_init": (_i, [_vp, _vp, _i, _i, _pi]),
    "rbl_comm_exchange": (_i, [_vp]),
    "rbl_comm_exchange_why": (ctypes.c_char_p, [_vp]),
    "rbl_comm_set_exchange": (_i, [_vp, _i]),
    "rbl_comm_profile": (_i, [_vp, _pd, _pi, _i]),
    "rbl_comm_world": (_i, [_vp]),
    "rbl_comm_rank": (_i, [_vp]),
    "rbl_stream": (_vp, [_vp]),
    "rbl_set_stream": (_i, [_vp, _vp]),
    "rbl_dev_alloc": (_i, [_vp, _sz, _pvp]),
    "rbl_dev_free": (_i, [_vp, _vp]),
    "rbl_pinned_alloc": (_i, [_vp, _sz, _pvp]),
    "rbl_pinned_free": (_i, [_vp, _vp]),
    "rbl_memcpy_h2d": (_i, [_vp, _vp, _vp, _sz]),
    "rbl_memcpy_d2h": (_i, [_vp, _vp, _vp, _sz]),
    "rbl_timer_start": (_i, [_vp]),
    "rbl_timer_stop": (_i, [_vp, _pd]),
    "rbl_flush_l2": (_i, [_vp]),
    "rbl_num_matvec_variants": (_i, [_vp]),
    "rbl_matvec_variant_info": (_i, [_vp, _i, _pi, _pi]),
    "rbl_set_matvec_variant": (_i, [_vp, _i]),
    "rbl_set_matvec_mode": (_i, [_vp, _i]),
    "rbl_num_sym_variants": (_i, [_vp]),
    "rbl_sym_variant_info": (_i, [_vp, _i, _pi, _pi]),
    "rbl_set_sym_variant": (_i, [_vp, _i]),
    "rbl_sym_variant_chunk": (_i, [_vp, _i]),
    "rbl_bd_phase_ms": (_i, [_vp, _pd, _i]),
    "rbl_M_RFD": (_i, [_vp, _vp, _d, _vp]),
    "rbl_M_RFD_from_U": (_i, [_vp, _vp, _vp, _d, _vp]),
    "rbl_KT_RFD_from_U": (_i, [_vp, _vp, _vp, _d, _vp]),
    "rbl_KTinv_RFD": (_i, [_vp, _vp, _d, _vp]),
    "rbl_M_RFD_cfgs": (_i, [_vp, _vp, _d, _vp, _vp]),
    "rbl_update_X_Q_out": (_i, [_vp, _vp, _vp, _vp]),
    "rbl_evolve_RFD": (_i, [_vp, _vp]),
    "rbl_set_rfd_delta": (_i, [_vp, _d]),
    "rbl_set_split_rand": (_i, [_vp, _i]),
    "rbl_set_split_weight": (_i, [_vp, _d]),
    "rbl_plan_cost_bounds": (_i, [_i, _i, _i, _i, _i, _d, _pi64, _pi64]),
    "rbl_set_mixed_precision": (_i, [_vp, _i]),
    "rbl_launch_count": (_i64, [_vp]),
    "rbl_product_count": (_i64, [_vp]),
    "rbl_bd_stats": (_i, [_vp, _pi, _pi]),
    "rbl_profile_matvec": (_i, [_vp, _i]),
    "rbl_matvec_profile": (_i, [_vp, _pd, _pi64, _i]),
    "rbl_fma_peak": (_i, [_vp, _i, _pd]),
}


def lib_path():
    return os.path.join(_HERE, "librbl.so")


def load():
    """dlopen librbl.so and declare every prototype.  Raises if it is not built."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise ImportError(f"{p} is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = ctypes.CDLL(p)
        for name, (res, args) in SYMBOLS.items():
            f = getattr(L, name)  # AttributeError here = header/library mismatch
            f.restype = res
            f.argtypes = args
        _LIB = L
    return _LIB


class RblError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__(f"[{STATUS.get(status, status)}] {msg}")
        self.status = status


class Context:
    """Thin owner of an ``rbl_ctx*`` for device-pointer callers."""

    def __init__(self, precision="single", device=-1):
        self.L = load()
        self.real = np.float32 if precision in ("single", "float32", RBL_F32) else np.float64
        h = ctypes.c_void_p()
        st = self.L.rbl_create(RBL_F32 if self.real == np.float32 else RBL_F64, device, ctypes.byref(h))
        if st != 0:
            raise RblError(st, self.L.rbl_last_error(None).decode())
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.rbl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, st):
        if st != 0:
            raise RblError(st, self.L.rbl_last_error(self.h).decode())

    def call(self, name, *args):
        self.check(getattr(self.L, name)(self.h, *args))

    # host-array helpers -----------------------------------------------------------
    def _arr(self, x):
        return np.ascontiguousarray(np.asarray(x, dtype=self.real).reshape(-1))

    def set_parameters(self, a, dt, kBT, eta, cfg):
        cfg = self._arr(cfg)
        self.call("rbl_set_parameters", a, dt, kBT, eta, cfg.ctypes.data, cfg.size // 3)

    def set_flags(self, block_pc, wall):
        self.call("rbl_set_flags", int(block_pc), int(wall))

    def set_config(self, X, Q):
        X, Q = self._arr(X), self._arr(Q)
        self.call("rbl_set_config", X.ctypes.data, Q.ctypes.data, X.size // 3)
        self.call("rbl_set_K_mats")

    def apply_M(self, F, r):
        F, r = self._arr(F), self._arr(r)
        out = np.empty_like(F)
        self.call("rbl_apply_M", F.ctypes.data, r.ctypes.data, F.size // 3, out.ctypes.data)
        return out

    def fma_peak(self, iters=4096):
        t = ctypes.c_double()
        self.call("rbl_fma_peak", iters, ctypes.byref(t))
        return t.value

    def launch_count(self):
        return int(self.L.rbl_launch_count(self.h))

    def timer_start(self):
        self.call("rbl_timer_start")

    def timer_stop(self):
        ms = ctypes.c_double()
        self.call("rbl_timer_stop", ctypes.byref(ms))
        return ms.value

    def matvec_profile(self, reset=True):
        ms, n = ctypes.c_double(), ctypes.c_int64()
        self.call("rbl_matvec_profile", ctypes.byref(ms), ctypes.byref(n), int(reset))
        return ms.value, n.value
