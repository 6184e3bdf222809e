"""ctypes binding of libgkg_b200.so (the C ABI declared in include/gkg_abi.h).

The library is built in-tree by ``build()`` (nvcc, sm_100a only) so that it travels to the
GPU box with the repo snapshot.  There is no CPU fallback: if the library is missing or a
call fails, a ``RuntimeError`` is raised.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GKG_LIB", os.path.join(HERE, "libgkg_b200.so"))   # override: A/B timing of two builds
CSRC = os.path.join(HERE, "csrc")

GKG_F32, GKG_BF16 = 0, 1
KNN_AUTO, KNN_EXACT_FP32, KNN_TCGEN05 = 0, 1, 2

_c = ctypes
_vp, _i32, _i64, _sz = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_size_t

# name -> (restype, argtypes); mirrors include/gkg_abi.h one to one
SIGNATURES = {
    "gkg_abi_version": (_i32, []),
    "gkg_last_error": (_c.c_char_p, []),
    "gkg_launch_count": (_c.c_uint64, []),
    "gkg_knn_workspace_bytes": (_sz, [_i32] * 10),
    "gkg_knn_graph": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _i32, _i32, _vp]
                      + [_i32] * 9 + [_vp, _sz, _vp]),
    "gkg_knn_prepare": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64] + [_i32] * 9 + [_vp, _sz, _vp]),
    "gkg_knn_select": (_i32, [_vp, _i64, _i64, _vp, _vp, _vp, _i32, _i32, _vp] + [_i32] * 10 + [_vp, _sz, _vp]),
    "gkg_knn_graph_debug": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _vp, _i32, _i32, _vp]
                            + [_i32] * 9 + [_vp, _sz, _vp] + [_i32, _vp, _vp, _i32, _i32]),
    "gkg_mr_aggregate_fwd": (_i32, [_vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp, _vp] + [_i32] * 7 + [_vp]),
    "gkg_mr_aggregate_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp] + [_i32] * 7 + [_vp]),
    "gkg_mr_aggregate_bwd_det": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp] + [_i32] * 7 + [_vp]),
    "gkg_fixed_to_float": (_i32, [_vp, _vp, _vp, _c.c_longlong, _vp]),
    "gkg_pool_keys_fwd": (_i32, [_vp, _i64, _i64, _vp] + [_i32] * 6 + [_vp]),
    "gkg_pool_keys_bwd": (_i32, [_vp, _vp] + [_i32] * 6 + [_vp]),
    "gkg_grouped_fc_supported": (_i32, [_i32]),
    "gkg_grouped_fc_pass_width": (_i32, [_i32]),
    "gkg_grouped_fc_fwd": (_i32, [_vp, _vp, _vp, _vp, _c.c_longlong, _i32, _i32, _vp]),
    "gkg_neighbor_gather_fwd": (_i32, [_vp, _i64, _i64, _vp, _vp] + [_i32] * 7 + [_vp]),
    "gkg_neighbor_gather_bwd": (_i32, [_vp, _vp, _vp] + [_i32] * 7 + [_vp]),
    "gkg_neighbor_sum_fwd": (_i32, [_vp, _i64, _i64, _vp, _vp] + [_i32] * 7 + [_vp]),
    "gkg_neighbor_sum_bwd": (_i32, [_vp, _vp, _vp] + [_i32] * 7 + [_vp]),
    "gkg_label_score_fwd": (_i32, [_vp] * 7 + [_i32] * 3 + [_vp]),
    "gkg_label_score_bwd": (_i32, [_vp] * 11 + [_i32] * 3 + [_vp]),
    "gkg_multilabel_loss": (_i32, [_vp] * 5 + [_c.c_longlong] + [_c.c_float] * 5 + [_vp]),
    "gkg_grouped_fc_pack_weights": (_i32, [_vp, _vp, _vp, _i32, _i32, _vp]),
    "gkg_grouped_fc_wgrad": (_i32, [_vp, _vp, _vp, _c.c_longlong, _i32, _vp]),
    "gkg_bn_workspace_bytes": (_sz, [_c.c_longlong, _i32]),
    "gkg_bn_stats": (_i32, [_vp, _c.c_longlong, _i32, _i32, _c.c_float, _c.c_float, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gkg_bn_act_forward": (_i32, [_vp] * 5 + [_c.c_longlong, _i32, _i32, _i32, _vp, _vp, _vp]),
    "gkg_bn_act_backward": (_i32, [_vp] * 7 + [_c.c_longlong, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gkg_bn_act_backward_reduce": (_i32, [_vp] * 7 + [_c.c_longlong, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _sz, _vp]),
    "gkg_bn_act_backward_elemt": (_i32, [_vp] * 8 + [_c.c_longlong, _c.c_longlong, _i32, _i32, _i32, _vp, _vp]),
    "gkg_column_sum": (_i32, [_vp, _c.c_longlong, _i32, _i32, _vp, _vp, _sz, _vp]),
    "gkg_bn_backward_reduce": (_i32, [_vp, _vp, _vp, _vp, _c.c_longlong, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
}

_lib = None
_lock = threading.Lock()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile csrc/*.cu into libgkg_b200.so (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", CSRC, "-j8"]
    if force:
        subprocess.run(["make", "-C", CSRC, "clean"], check=True, capture_output=not verbose)
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libgkg_b200.so failed:\n" + res.stdout[-4000:] + res.stderr[-4000:])
    if verbose:
        print(res.stdout)
    return LIB_PATH


def load():
    """Load the shared library (once) and attach the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(gkgnet_b200 has no CPU fallback for its CUDA kernels)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        if lib.gkg_abi_version() != 2:
            raise RuntimeError("libgkg_b200.so ABI version mismatch")
        _lib = lib
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().gkg_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def launch_count() -> int:
    return int(load().gkg_launch_count())
