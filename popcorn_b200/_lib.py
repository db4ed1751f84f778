"""ctypes binding of libpopcorn_b200.so (the C-ABI declared in include/popcorn_b200.h).

There is no CPU fallback and no alternative backend: if the shared library is missing or a call
fails, a RuntimeError is raised.  ``build()`` compiles it in-tree with nvcc for sm_100a.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POPCORN_B200_LIB") or os.path.join(_HERE, "libpopcorn_b200.so")   # override: development builds
CSRC = os.path.join(_HERE, "csrc")

_lib = None

_vp, _i, _ll, _sz, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_size_t, C.c_float

# name -> (restype, argtypes); must list every symbol include/popcorn_b200.h declares
SIGNATURES = {
    "pc_version": (_i, []),
    "pc_last_error": (C.c_char_p, []),
    "pc_launch_count": (_ll, [_i]),
    "pc_profile_enable": (_i, [_i]),
    "pc_profile_num": (_i, []),
    "pc_profile_name": (C.c_char_p, [_i]),
    "pc_profile_get": (_i, [_i, C.POINTER(C.c_double), C.POINTER(C.c_longlong), C.POINTER(C.c_double)]),
    "pc_memcpy2d_async": (_i, [_vp, _sz, _vp, _sz, _sz, _sz, _i, _vp]),
    "pc_dda_pack_floats": (_i, []),
    "pc_dda_pack_offset": (_i, [_i, _i]),
    "pc_head_pack_floats": (_i, [_i]),
    "pc_dda_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "pc_dda_tc_pack_base": (_i, []),
    "pc_dda_tc_pack_floats": (_i, []),
    "pc_dda_tc_pack": (_i, [_vp, _vp]),
    "pc_conv_tc_layer_floats": (_i, [_i, _i]),
    "pc_conv_tc_pack_layer": (_i, [_vp, _i, _i, _vp]),
    "pc_dda_forward": (_i, [_vp, _ll, _vp, _i, _i, _i, _i, _ll, _ll, _i, _i, _i, _i, _i, _i, _vp, _ll, _ll, _i, _vp, _sz, _vp]),
    "pc_head_dense_forward": (_i, [_vp, _i, _vp, _ll, _ll, _i, _vp, _ll, _i, _i, _i, _i, _vp, _vp, _ll, _i, _vp, _ll, _i,
                                   _vp, _vp, _i, _vp]),
    "pc_head_tc_pack_bytes": (_i, []),
    "pc_tc_operand_format": (_i, []),
    "pc_head_dense_forward_tc": (_i, [_vp, _i, _vp, _ll, _ll, _i, _vp, _ll, _i, _i, _i, _i, _vp, _vp, _ll, _i, _vp, _ll, _i,
                                      _vp, _vp, _i, _vp]),
    "pc_head_sparse_forward_tc": (_i, [_vp, _i, _vp, _ll, _ll, _vp, _vp, _vp, _ll, _ll, _vp, _vp, _vp, _vp]),
    "pc_infer_tile_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "pc_infer_tile_fused": (_i, [_vp, _vp, _ll, _vp, _vp, _i, _i, _i, _i, _ll, _ll, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "pc_compact_workspace_bytes": (_sz, [_ll]),
    "pc_sparse_mask_compact": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _sz, _vp]),
    "pc_head_sparse_forward": (_i, [_vp, _i, _vp, _ll, _ll, _vp, _vp, _vp, _ll, _ll, _vp, _vp, _vp, _vp]),
    "pc_head_bwd_workspace_bytes": (_sz, [_i]),
    "pc_head_sparse_backward": (_i, [_vp, _i, _vp, _ll, _ll, _vp, _vp, _vp, _ll, _ll, _vp, _f, _vp, _vp, _vp, _sz, _vp, _vp]),
    "pc_region_sum": (_i, [_vp, _vp, _ll, _i, _vp, _vp]),
    "pc_region_sum_backward": (_i, [_vp, _vp, _ll, _i, _vp, _vp]),
    "pc_region_scale": (_i, [_vp, _vp, _ll, _i, _vp, _vp]),
    "pc_accumulate_tile": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "pc_finalize_map": (_i, [_vp, _vp, _vp, _vp, _vp, _ll, _vp]),
    "pc_ingest_normalize": (_i, [_vp, _i, _i, _ll, _i, C.c_uint, _vp, _i, _ll, _i, _i, _i, _vp, _vp, _vp, _ll, _i, _vp]),
    "pc_conv3x3_layer": (_i, [_vp, _i, _ll, _i, _i, _i, _i, _i, _i, C.c_uint, _vp, _i, _ll, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _i, _i,
                               _vp, _ll, _i, _vp, _ll, _i, _vp]),
    "pc_convt2x2_layer": (_i, [_vp, _i, _ll, _i, _i, _i, _vp, _vp, _ll, _i, _vp]),
    "pc_conv_wgrad_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "pc_conv3x3_wgrad": (_i, [_vp, _i, _ll, _i, _i, _i, _i, _i, _i, C.c_uint, _vp, _i, _ll, _i, _i, _i, _i, _i, _vp, _ll, _i, _i, _i, _i,
                               _vp, _i, _vp, _sz, _vp]),
    "pc_relu_backward": (_i, [_vp, _ll, _i, _vp, _ll, _i, _vp, _ll, _i, _vp, _ll, _i, _i, _i, _i, _vp]),
    "pc_maxpool2x2_relu_backward": (_i, [_vp, _ll, _i, _vp, _ll, _i, _vp, _ll, _i, _vp, _ll, _i, _i, _i, _i, _vp]),
    "pc_convt2x2_dgrad": (_i, [_vp, _ll, _i, _vp, _i, _i, _i, _vp, _ll, _i, _vp]),
    "pc_convt_wgrad_workspace_bytes": (_sz, [_i, _i, _i]),
    "pc_convt2x2_wgrad": (_i, [_vp, _ll, _i, _vp, _ll, _i, _i, _i, _i, _vp, _i, _vp, _sz, _vp]),
    "pc_test_conv3x3": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "pc_test_convt2x2": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "pc_test_fma_peak": (_i, [_i, _i, _i, _vp, _vp]),
}


def build(verbose: bool = False) -> str:
    """Compile csrc/*.cu -> libpopcorn_b200.so (nvcc, -gencode arch=compute_100a,code=sm_100a -lineinfo)."""
    r = subprocess.run(["make", "-C", CSRC, "-j", "8"], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("popcorn_b200: building libpopcorn_b200.so failed")
    return LIB_PATH


def lib() -> C.CDLL:
    """The loaded library; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                f"popcorn_b200: {LIB_PATH} not found — build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C popcorn_b200/csrc`.  There is no CPU / PyTorch fallback for the hot path.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)   # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = lib().pc_last_error()
        raise RuntimeError(f"popcorn_b200 {what} failed (code {rc}): {msg.decode() if msg else ''}")
