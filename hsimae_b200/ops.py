"""Single-operator entry points of the library on torch CUDA tensors (thin ctypes
wrappers; used by the parity tests and handy for profiling one kernel)."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

EPI_BIAS_BF16, EPI_BIAS_F32, EPI_RESID_LN, EPI_SWIGLU, EPI_DSWIGLU, EPI_DGATE = range(6)
IMPL_TC, IMPL_SIMT = 0, 1


def _p(t: Optional[torch.Tensor]):
    return t.data_ptr() if t is not None else None


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("hsimae_b200 operators run on CUDA tensors only")


def gemm(A, B, epilogue, *, impl=IMPL_TC, bias=None, resid=None, resid2=None, gamma=None, beta=None, ab=None,
         rowscale=None, rs_mode=0, rs_K=1, rs_len_l=1, rs_G=1, want_stats=True, A2=None, B2=None, keep_ab=True):
    """C = A[M,K] @ B[N,K]^T with the fused epilogue `epilogue`; returns a dict of outputs."""
    _need_cuda(A, B)
    L = _lib.load()
    M, K = A.shape
    N = B.shape[0]
    dev = A.device
    d = _lib.GemmDesc()
    d.M, d.N, d.K, d.epilogue, d.impl = M, N, K, epilogue, impl
    d.A, d.lda, d.B, d.ldb = _p(A), A.stride(0), _p(B), B.stride(0)
    out = {}
    if epilogue == EPI_BIAS_BF16:
        out["out"] = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        d.out0, d.ld0 = _p(out["out"]), N
    elif epilogue == EPI_BIAS_F32:
        out["out"] = torch.empty(M, N, dtype=torch.float32, device=dev)
        d.out0, d.ld0 = _p(out["out"]), N
    elif epilogue == EPI_RESID_LN:
        out["x"] = torch.empty(M, N, dtype=torch.float32, device=dev)
        d.out0, d.ld0 = _p(out["x"]), N
        if gamma is not None:
            out["ln"] = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
            d.out1, d.ld1 = _p(out["ln"]), N
            if want_stats:
                out["stats"] = torch.empty(M, 2, dtype=torch.float32, device=dev)
                d.stats = _p(out["stats"])
    elif epilogue == EPI_SWIGLU:
        out["ab"] = torch.empty(M, N, dtype=torch.bfloat16, device=dev) if keep_ab else None
        out["g"] = torch.empty(M, N // 2, dtype=torch.bfloat16, device=dev)
        d.out0, d.ld0, d.out1, d.ld1 = _p(out["ab"]), N, _p(out["g"]), N // 2
    elif epilogue == EPI_DSWIGLU:
        out["dab"] = torch.empty(M, 2 * N, dtype=torch.bfloat16, device=dev)
        d.out0, d.ld0 = _p(out["dab"]), 2 * N
        d.ab, d.ldab = _p(ab), ab.stride(0)
    elif epilogue == EPI_DGATE:
        # dab = dswiglu(A @ B^T, A2 @ B2^T + bias): the pre-activations are recomputed inside the kernel
        _need_cuda(A2, B2)
        out["dab"] = torch.empty(M, 2 * N, dtype=torch.bfloat16, device=dev)
        d.out0, d.ld0 = _p(out["dab"]), 2 * N
        d.A2, d.lda2, d.B2, d.ldb2 = _p(A2), A2.stride(0), _p(B2), B2.stride(0)
    d.bias = _p(bias)
    d.resid, d.ldr, d.resid2 = _p(resid), (resid.stride(0) if resid is not None else 0), _p(resid2)
    d.gamma, d.beta = _p(gamma), _p(beta)
    d.rowscale, d.rs_mode, d.rs_K, d.rs_len_l, d.rs_G = _p(rowscale), rs_mode, rs_K, rs_len_l, rs_G
    scratch = None
    if impl == IMPL_SIMT:
        scratch = torch.empty(M, N, dtype=torch.float32, device=dev)
        d.scratch = _p(scratch)
    _lib.check(L.hsimae_gemm(C.byref(d), _stream()), "gemm")
    return out


def gemm_lnbwd(A, B, x, stats, gamma, dx_in, *, want_dxb=True, rowscale=None, rs_mode=0, rs_K=1, rs_len_l=1, rs_G=1,
               dgamma=None, dbeta=None, inplace=False):
    """dy = A[M,K] @ B[N,K]^T followed by the LayerNorm backward of `x` (row statistics `stats` = (mean, rstd), weight `gamma`)
    in one launch: returns dx = dx_in + LN_bwd(dy), bf16(rowscale * dx) and the accumulated dgamma / dbeta."""
    _need_cuda(A, B)
    L = _lib.load()
    M, K = A.shape
    N = B.shape[0]
    dev = A.device
    d = _lib.LnBwdDesc()
    d.M, d.N, d.K = M, N, K
    d.A, d.lda, d.B, d.ldb = _p(A), A.stride(0), _p(B), B.stride(0)
    d.x, d.ldx, d.stats, d.gamma = _p(x), x.stride(0), _p(stats), _p(gamma)
    out = {"dx": dx_in if inplace else torch.empty(M, N, dtype=torch.float32, device=dev)}
    d.dx_in, d.ldi, d.dx_out, d.ldo = _p(dx_in), dx_in.stride(0), _p(out["dx"]), N
    if want_dxb:
        out["dxb"] = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        d.dxb, d.ldxb = _p(out["dxb"]), N
    out["dgamma"] = dgamma if dgamma is not None else torch.zeros(N, dtype=torch.float32, device=dev)
    out["dbeta"] = dbeta if dbeta is not None else torch.zeros(N, dtype=torch.float32, device=dev)
    d.dgamma, d.dbeta = _p(out["dgamma"]), _p(out["dbeta"])
    d.rowscale, d.rs_mode, d.rs_K, d.rs_len_l, d.rs_G = _p(rowscale), rs_mode, rs_K, rs_len_l, rs_G
    _lib.check(L.hsimae_gemm_lnbwd(C.byref(d), _stream()), "gemm_lnbwd")
    return out


def mlp_fused(X, W13, b13, W2, b2, resid, *, gamma=None, beta=None, resid2=None, keep_g=True,
              rowscale=None, rs_mode=0, rs_K=1, rs_len_l=1, rs_G=1):
    """out = resid + rs * (W2 (silu(W1 x) * W3 x) + b2) [+ resid2], LayerNorm of the result -- one kernel (csrc/block_fused.cu).
    W13 / b13 in the library's interleave (pack_interleaved); returns dict(x, ln, stats, g)."""
    _need_cuda(X, W13, W2, resid)
    L = _lib.load()
    M, D = X.shape
    Hp = W2.shape[1]
    dev = X.device
    d = _lib.MlpDesc()
    d.M, d.D, d.Hp = M, D, Hp
    d.X, d.ldx, d.W13, d.ldw13, d.b13 = _p(X), X.stride(0), _p(W13), W13.stride(0), _p(b13)
    d.W2, d.ldw2, d.b2 = _p(W2), W2.stride(0), _p(b2)
    d.resid, d.ldr, d.resid2 = _p(resid), resid.stride(0), _p(resid2)
    d.rowscale, d.rs_mode, d.rs_K, d.rs_len_l, d.rs_G = _p(rowscale), rs_mode, rs_K, rs_len_l, rs_G
    out = {"x": torch.empty(M, D, dtype=torch.float32, device=dev), "ln": None, "stats": None, "g": None}
    d.out, d.ldo = _p(out["x"]), D
    if gamma is not None:
        out["ln"] = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
        out["stats"] = torch.empty(M, 2, dtype=torch.float32, device=dev)
        d.gamma, d.beta, d.ln, d.ldln, d.stats = _p(gamma), _p(beta), _p(out["ln"]), D, _p(out["stats"])
    if keep_g:
        out["g"] = torch.empty(M, Hp, dtype=torch.bfloat16, device=dev)
        d.g, d.ldg = _p(out["g"]), Hp
    _lib.check(L.hsimae_mlp_fused(C.byref(d), _stream()), "mlp_fused")
    return out


def wgrad(Y, X, dst0, *, impl=IMPL_TC, dst1=None, row_map=0, rows_valid=None, cols_valid=None, bias0=None, bias1=None):
    """dst[map(n), k] += sum_m Y[m, n] X[m, k]  (accumulates in place)."""
    _need_cuda(Y, X, dst0)
    L = _lib.load()
    d = _lib.WgradDesc()
    d.Mred, d.Nout, d.Kin, d.impl = Y.shape[0], Y.shape[1], X.shape[1], impl
    d.Y, d.ldy, d.X, d.ldx = _p(Y), Y.stride(0), _p(X), X.stride(0)
    d.dst0, d.dst1, d.ld, d.row_map = _p(dst0), _p(dst1), dst0.stride(0), row_map
    d.rows_valid = rows_valid if rows_valid is not None else dst0.shape[0]
    d.cols_valid = cols_valid if cols_valid is not None else dst0.shape[1]
    d.bias0, d.bias1 = _p(bias0), _p(bias1)
    _lib.check(L.hsimae_wgrad(C.byref(d), _stream()), "wgrad")


def wgrad_group(jobs):
    """several `wgrad` problems in one launch; `jobs` = list of dicts with the keyword arguments of `wgrad` (Y, X, dst0, ...)"""
    L = _lib.load()
    arr = (_lib.WgradDesc * len(jobs))()
    for d, j in zip(arr, jobs):
        Y, X, dst0 = j["Y"], j["X"], j["dst0"]
        _need_cuda(Y, X, dst0)
        d.Mred, d.Nout, d.Kin, d.impl = Y.shape[0], Y.shape[1], X.shape[1], 0
        d.Y, d.ldy, d.X, d.ldx = _p(Y), Y.stride(0), _p(X), X.stride(0)
        d.dst0, d.dst1, d.ld, d.row_map = _p(dst0), _p(j.get("dst1")), dst0.stride(0), j.get("row_map", 0)
        d.rows_valid = j.get("rows_valid", dst0.shape[0])
        d.cols_valid = j.get("cols_valid", dst0.shape[1])
        d.bias0, d.bias1 = _p(j.get("bias0")), _p(j.get("bias1"))
    _lib.check(L.hsimae_wgrad_group(arr, len(jobs), _stream()), "wgrad_group")


def attention_forward(qkv, n, D, heads, K, nseq, length, seq_step, tok_step):
    _need_cuda(qkv)
    L = _lib.load()
    out = torch.empty(n * K, D, dtype=torch.bfloat16, device=qkv.device)
    lse = torch.empty(n * K, heads, dtype=torch.float32, device=qkv.device)
    _lib.check(L.hsimae_attention_forward(_p(qkv), _p(out), _p(lse), n, D, heads, K, nseq, length, seq_step, tok_step, _stream()),
               "attention_forward")
    return out, lse


def attention_backward(qkv, out, lse, dout, n, D, heads, K, nseq, length, seq_step, tok_step):
    _need_cuda(qkv)
    L = _lib.load()
    dqkv = torch.empty_like(qkv)
    _lib.check(L.hsimae_attention_backward(_p(qkv), _p(out), _p(lse), _p(dout), _p(dqkv), n, D, heads, K, nseq, length, seq_step,
                                           tok_step, _stream()), "attention_backward")
    return dqkv


def mask(noise_t, noise_l, len_t, len_l):
    """-> ids_keep i64 [N,K], ids_restore i64 [N,T*L], mask f32 [N,T*L]"""
    _need_cuda(noise_t, noise_l)
    L = _lib.load()
    n, T = noise_t.shape
    Lp = noise_l.shape[1]
    dev = noise_t.device
    ids_keep = torch.empty(n, len_t * len_l, dtype=torch.int64, device=dev)
    ids_restore = torch.empty(n, T * Lp, dtype=torch.int64, device=dev)
    m = torch.empty(n, T * Lp, dtype=torch.float32, device=dev)
    _lib.check(L.hsimae_mask(_p(noise_t.contiguous()), _p(noise_l.contiguous()), n, T, Lp, len_t, len_l, _p(ids_keep),
                             _p(ids_restore), _p(m), None, None, _stream()), "mask")
    return ids_keep, ids_restore, m


def pack_interleaved(w1: torch.Tensor, w3: torch.Tensor, hp: int) -> torch.Tensor:
    """[H, d] x 2 -> [2*hp, d] in the library's w1|w3 interleave (16 rows of w1, 16 of w3, ...); test helper."""
    H, d = w1.shape
    out = torch.zeros(2 * hp, d, dtype=w1.dtype, device=w1.device)
    idx = torch.arange(H, device=w1.device)
    out[(idx // 16) * 32 + idx % 16] = w1
    out[(idx // 16) * 32 + 16 + idx % 16] = w3
    return out
