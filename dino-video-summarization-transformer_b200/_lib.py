"""ctypes binding of libdvst.so (the C ABI declared in include/dvst.h).

The product path has no CPU / PyTorch fallback: if the shared library is missing, `lib()` raises.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DVST_LIB: load another build of the library (same-box A/B runs of two kernel versions); default = the in-tree build
LIB_PATH = os.environ.get("DVST_LIB") or os.path.join(_HERE, "csrc", "libdvst.so")

_lib = None

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_float = ctypes.c_float


class DvstError(RuntimeError):
    pass


# name -> (restype, argtypes); kept in one table so tests can check every symbol of include/dvst.h is exported.
SIGNATURES = {
    "dvst_last_error": (ctypes.c_char_p, []),
    "dvst_launch_count": (c_ll, []),
    "dvst_abi_version": (c_int, []),
    "dvst_gemm_bf16": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_ll,
                               c_void_p, c_ll, c_int, c_ll, c_int, c_void_p]),
    "dvst_set_gemm_cta_group": (None, [c_int]),
    "dvst_set_producer_ring": (None, [c_int]),
    "dvst_set_pdl": (None, [c_int]),
    "dvst_set_attention_impl": (None, [c_int]),
    "dvst_set_chunk_rows": (None, [c_ll]),
    "dvst_set_ln_fusion": (None, [c_int]),
    "dvst_set_temporal_fusion": (None, [c_int]),
    "dvst_gemm_ln_temporal_attention": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_ll,
                                                c_int, c_ll, c_void_p, c_ll, c_void_p]),
    "dvst_temporal_tile_geometry": (c_int, [c_int, c_void_p, c_void_p]),
    "dvst_profile_enable": (None, [c_int]),
    "dvst_profile_collect": (c_int, [c_void_p, c_void_p, c_void_p, c_int]),
    "dvst_create": (c_int, [c_void_p, c_void_p]),
    "dvst_destroy": (None, [c_void_p]),
    "dvst_set_weight": (c_int, [c_void_p, ctypes.c_char_p, c_void_p]),
    "dvst_finalize_weights": (c_int, [c_void_p]),
    "dvst_workspace_bytes": (c_ll, [c_void_p, c_int, c_int, c_int, c_int]),
    "dvst_forward_features": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p, c_ll,
                                      c_void_p]),
    "dvst_forward_attention_map": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_ll,
                                           c_void_p]),
    "dvst_pair_loss": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_float, c_float, c_int, c_int, c_void_p,
                               c_void_p]),
    "dvst_forward_windows": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                     c_ll, c_void_p]),
    "dvst_forward_windows_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                        c_int, c_int, c_void_p, c_void_p, c_ll, c_void_p]),
    "dvst_pair_loss_indexed": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_float, c_float,
                                       c_int, c_int, c_void_p, c_void_p]),
    "dvst_head_loss_scratch_bytes": (c_ll, [c_int]),
    "dvst_head_loss": (c_int, [c_void_p, c_void_p, c_ll, c_void_p, c_void_p, c_float, c_float, c_int, c_int, c_void_p,
                               c_void_p, c_ll, c_void_p]),
    "dvst_colsum": (c_int, [c_void_p, c_ll, c_void_p, c_int, c_int, c_void_p]),
    "dvst_center_ema": (c_int, [c_void_p, c_void_p, c_float, c_float, c_int, c_void_p]),
    "dvst_cast_bf16": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_ll, c_ll, c_int, c_void_p]),
    "dvst_l2norm_rows": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_ll, c_void_p, c_ll, c_int, c_void_p]),
    "dvst_layernorm768": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_ll, c_int, c_ll, c_ll, c_float,
                                  c_void_p]),
    "dvst_select_frames": (c_int, [c_void_p, c_ll, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "dvst_plan_windows": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p]),
    "dvst_preprocess_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_void_p]),
    "dvst_rowstats768": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_ll, c_void_p]),
    "dvst_row_offsets768": (c_int, [c_void_p, c_void_p, c_ll, c_void_p]),
    "dvst_gemm_bf16_ln": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                  c_float, c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_void_p, c_void_p, c_ll, c_int,
                                  c_void_p]),
    "dvst_knn_similarity": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_ll, c_void_p]),
    "dvst_knn_topk": (c_int, [c_void_p, c_ll, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "dvst_knn_vote": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_float, c_int, c_void_p,
                              c_void_p, c_void_p]),
    "dvst_attention_scratch_bytes": (c_ll, [c_int, c_int, c_int]),
    "dvst_attention_ws": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_int, c_ll, c_void_p,
                                  c_ll, c_void_p]),
    "dvst_attention_hp": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_int, c_ll, c_void_p,
                                  c_ll, c_void_p]),
    "dvst_attention": (c_int, [c_void_p, c_ll, c_void_p, c_ll, c_void_p, c_int, c_int, c_int, c_int, c_ll, c_void_p]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DvstError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback for the CUDA path)")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            if not hasattr(l, name) and os.environ.get("DVST_LIB_ALLOW_MISSING") == "1":
                continue   # same-box A/B runs against an older build of the library (DVST_LIB=...)
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        # experiment switches (A/B runs of bench.py on one box): DVST_LN_FUSION=0 forces the stand-alone LayerNorm
        if os.environ.get("DVST_LN_FUSION") == "0":
            l.dvst_set_ln_fusion(0)
        # DVST_PDL=1: programmatic dependent launch for the hot-path kernels (off by default)
        if os.environ.get("DVST_PDL") == "1":
            l.dvst_set_pdl(1)
        if os.environ.get("DVST_PRODUCER_RING"):
            l.dvst_set_producer_ring(int(os.environ["DVST_PRODUCER_RING"]))
        # DVST_TATTN_FUSION=0: qkv GEMM + temporal attention as two kernels (the round-1 form)
        if os.environ.get("DVST_TATTN_FUSION") == "0":
            l.dvst_set_temporal_fusion(0)
        # DVST_ATTENTION_IMPL: 1 default (split tcgen05 kernel), 3 first-generation one-pass, 2 two-pass, 0 mma.sync
        if os.environ.get("DVST_ATTENTION_IMPL"):
            l.dvst_set_attention_impl(int(os.environ["DVST_ATTENTION_IMPL"]))
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise DvstError(lib().dvst_last_error().decode())


def launch_count():
    return int(lib().dvst_launch_count())


PROF_KINDS = ("gemm", "attention", "layernorm", "elementwise", "head_loss", "gemm_hbm", "gemm_tattn")


def profile_enable(on):
    lib().dvst_profile_enable(1 if on else 0)


def profile_collect():
    """{kind: (ms, work, launches)} since the last collect (waits for the recorded events)."""
    n = len(PROF_KINDS)
    ms, work, cnt = (ctypes.c_double * n)(), (ctypes.c_double * n)(), (ctypes.c_longlong * n)()
    check(lib().dvst_profile_collect(ms, work, cnt, n))
    return {k: (ms[i], work[i], int(cnt[i])) for i, k in enumerate(PROF_KINDS)}


def ptr(t):
    return None if t is None else c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)
