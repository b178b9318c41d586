"""ctypes binding of libhsimae_b200.so (the C ABI in include/hsimae_b200.h).

There is no fallback: if the library is missing or a call fails, a
``RuntimeError`` carrying ``hsimae_last_error()`` is raised.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libhsimae_b200.so"

c_void_p, c_int, c_i32, c_i64, c_f32 = C.c_void_p, C.c_int, C.c_int32, C.c_int64, C.c_float


class Dims(C.Structure):
    _fields_ = [(n, c_i32) for n in (
        "img_size", "patch_size", "bands", "b_patch_size", "embed_dim", "depth", "s_depth", "num_heads",
        "dec_dim", "dec_depth", "dec_heads", "num_class", "qkv_bias", "norm_pix_loss")] + [("mlp_ratio", c_f32)]


class GemmDesc(C.Structure):
    _fields_ = [
        ("M", c_i32), ("N", c_i32), ("K", c_i32), ("epilogue", c_i32), ("impl", c_i32),
        ("A", c_void_p), ("lda", c_i32), ("B", c_void_p), ("ldb", c_i32),
        ("out0", c_void_p), ("ld0", c_i32), ("out1", c_void_p), ("ld1", c_i32),
        ("bias", c_void_p), ("resid", c_void_p), ("ldr", c_i32), ("resid2", c_void_p),
        ("gamma", c_void_p), ("beta", c_void_p), ("stats", c_void_p),
        ("ab", c_void_p), ("ldab", c_i32),
        ("rowscale", c_void_p), ("rs_mode", c_i32), ("rs_K", c_i32), ("rs_len_l", c_i32), ("rs_G", c_i32),
        ("scratch", c_void_p),
        ("A2", c_void_p), ("lda2", c_i32), ("B2", c_void_p), ("ldb2", c_i32),
    ]


class LnBwdDesc(C.Structure):
    _fields_ = [
        ("M", c_i32), ("N", c_i32), ("K", c_i32),
        ("A", c_void_p), ("lda", c_i32), ("B", c_void_p), ("ldb", c_i32),
        ("x", c_void_p), ("ldx", c_i32), ("stats", c_void_p), ("gamma", c_void_p),
        ("dx_in", c_void_p), ("ldi", c_i32), ("dx_out", c_void_p), ("ldo", c_i32), ("dxb", c_void_p), ("ldxb", c_i32),
        ("rowscale", c_void_p), ("rs_mode", c_i32), ("rs_K", c_i32), ("rs_len_l", c_i32), ("rs_G", c_i32),
        ("dgamma", c_void_p), ("dbeta", c_void_p),
    ]


class MlpDesc(C.Structure):
    _fields_ = [
        ("M", c_i32), ("D", c_i32), ("Hp", c_i32),
        ("X", c_void_p), ("ldx", c_i32), ("W13", c_void_p), ("ldw13", c_i32), ("b13", c_void_p),
        ("W2", c_void_p), ("ldw2", c_i32), ("b2", c_void_p),
        ("resid", c_void_p), ("ldr", c_i32), ("resid2", c_void_p),
        ("rowscale", c_void_p), ("rs_mode", c_i32), ("rs_K", c_i32), ("rs_len_l", c_i32), ("rs_G", c_i32),
        ("out", c_void_p), ("ldo", c_i32),
        ("gamma", c_void_p), ("beta", c_void_p), ("ln", c_void_p), ("ldln", c_i32), ("stats", c_void_p),
        ("g", c_void_p), ("ldg", c_i32),
    ]


class WgradDesc(C.Structure):
    _fields_ = [
        ("Mred", c_i32), ("Nout", c_i32), ("Kin", c_i32), ("impl", c_i32),
        ("Y", c_void_p), ("ldy", c_i32), ("X", c_void_p), ("ldx", c_i32),
        ("dst0", c_void_p), ("dst1", c_void_p), ("ld", c_i32), ("row_map", c_i32), ("rows_valid", c_i32),
        ("cols_valid", c_i32), ("bias0", c_void_p), ("bias1", c_void_p),
    ]


# name -> (restype, argtypes); mirrors include/hsimae_b200.h one to one
PROTOTYPES = {
    "hsimae_last_error": (C.c_char_p, []),
    "hsimae_abi_version": (c_int, []),
    "hsimae_launch_count": (c_i64, []),
    "hsimae_plan_create": (c_int, [C.POINTER(Dims), C.POINTER(c_void_p)]),
    "hsimae_plan_destroy": (None, [c_void_p]),
    "hsimae_plan_num_params": (c_int, [c_void_p]),
    "hsimae_plan_param_name": (C.c_char_p, [c_void_p, c_int]),
    "hsimae_plan_param_numel": (c_i64, [c_void_p, c_int]),
    "hsimae_plan_param_grad_offset": (c_i64, [c_void_p, c_int]),
    "hsimae_plan_param_has_grad": (c_int, [c_void_p, c_int]),
    "hsimae_plan_grad_arena_elems": (c_i64, [c_void_p]),
    "hsimae_plan_bf16_arena_elems": (c_i64, [c_void_p]),
    "hsimae_plan_f32_arena_elems": (c_i64, [c_void_p]),
    "hsimae_plan_pack_table_bytes": (c_i64, [c_void_p]),
    "hsimae_plan_hidden": (c_int, [c_void_p, c_int]),
    "hsimae_plan_grad_bucket": (c_int, [c_void_p, c_int, C.POINTER(c_i64), C.POINTER(c_i64)]),
    "hsimae_pack_params": (c_int, [c_void_p, C.POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p]),
    "hsimae_mask": (c_int, [c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_void_p, c_void_p, c_void_p,
                            c_void_p, c_void_p, c_void_p]),
    "hsimae_encoder_workspace_bytes": (c_i64, [c_void_p, c_i32, c_i32, c_i32, c_i32]),
    "hsimae_encoder_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_void_p,
                                       C.POINTER(c_void_p), c_i32, c_void_p, c_i64, c_void_p]),
    "hsimae_encoder_forward_scene": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i64, c_i32, c_void_p, c_i64,
                                             c_void_p]),
    "hsimae_encoder_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_void_p,
                                        C.POINTER(c_void_p), c_void_p, c_i64, c_void_p, c_i32, c_void_p]),
    "hsimae_encoder_latent": (c_int, [c_void_p, c_i32, c_i32, c_i32, c_i32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hsimae_decoder_workspace_bytes": (c_i64, [c_void_p, c_i32, c_i32, c_i32, c_i32]),
    "hsimae_decoder_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_void_p, c_void_p,
                                       c_void_p, c_i32, c_i32, c_void_p, c_i64, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p]),
    "hsimae_decoder_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_void_p, c_void_p, c_void_p,
                                        c_i64, c_void_p, c_void_p, c_void_p]),
    "hsimae_head_forward": (c_int, [c_void_p, c_void_p, c_i32, c_void_p, c_i32, c_void_p, c_void_p, c_void_p]),
    "hsimae_head_backward": (c_int, [c_void_p, c_void_p, c_i32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "hsimae_gemm": (c_int, [C.POINTER(GemmDesc), c_void_p]),
    "hsimae_gemm_lnbwd": (c_int, [C.POINTER(LnBwdDesc), c_void_p]),
    "hsimae_wgrad": (c_int, [C.POINTER(WgradDesc), c_void_p]),
    "hsimae_wgrad_group": (c_int, [C.POINTER(WgradDesc), c_i32, c_void_p]),
    "hsimae_mlp_fused": (c_int, [C.POINTER(MlpDesc), c_void_p]),
    "hsimae_set_pdl": (c_int, [c_int]),
    "hsimae_helper_join": (c_int, [c_void_p, c_void_p]),
    "hsimae_attention_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32, c_i32,
                                         c_i32, c_void_p]),
    "hsimae_attention_backward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_i32, c_i32,
                                          c_i32, c_i32, c_i32, c_i32, c_void_p]),
    "hsimae_adamw_tile_elems": (c_i32, []),
    "hsimae_adamw_step": (c_int, [c_void_p, c_i32, c_i32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                  c_void_p]),
    "hsimae_gwpca_workspace_bytes": (c_i64, [c_i32, c_i32, c_i32]),
    "hsimae_gwpca_moments": (c_int, [c_void_p, c_i32, c_i64, c_i32, c_i32, C.POINTER(c_i32), c_void_p, c_i64, c_void_p, c_void_p,
                                     c_void_p, c_void_p]),
    "hsimae_gwpca_project": (c_int, [c_void_p, c_i32, c_i64, c_i32, c_i32, C.POINTER(c_i32), c_i32, c_void_p, c_void_p, c_void_p,
                                     c_i32, c_void_p, c_i64, c_void_p]),
    "hsimae_gather_patches": (c_int, [c_void_p, c_void_p, c_void_p, c_i32, c_i32, c_void_p, c_void_p, c_void_p, c_i32, c_void_p, c_void_p]),
}

ABI_VERSION = 2   # include/hsimae_b200.h HSIMAE_ABI_VERSION
_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = Path(os.environ.get("HSIMAE_B200_LIB", LIB_PATH))
    if not path.exists():
        raise RuntimeError(
            f"hsimae_b200: {path} not found. Build it with `python -m hsimae_b200.build` "
            "(or __graft_entry__.build()); there is no CPU / PyTorch fallback for this path.")
    lib = C.CDLL(str(path))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)  # AttributeError => header/library mismatch, surface it
        fn.restype = res
        fn.argtypes = args
    if lib.hsimae_abi_version() != ABI_VERSION:
        raise RuntimeError("hsimae_b200: ABI version mismatch between _lib.py and the shared library")
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().hsimae_last_error()
        raise RuntimeError(f"hsimae_b200 {what} failed (status {status}): {msg.decode() if msg else '?'}")
