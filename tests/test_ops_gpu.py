"""GPU: single-kernel parity.  tcgen05 GEMMs vs the CUDA-core checker vs a torch fp32 reference of the
same op (inputs are the same bf16 values, so differences are accumulation order + output rounding)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from oracle import hsimae_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"

# bf16 output rounding is 2^-9 relative; fp32 accumulation-order differences are ~1e-6
TOL_BF16 = 6e-3
TOL_F32 = 2e-5


def _rand_bf16(*shape, scale=1.0, seed=0):
    g = torch.Generator(device=DEV).manual_seed(seed)
    return (scale * torch.randn(*shape, device=DEV, generator=g)).to(torch.bfloat16)


@pytest.fixture(scope="module")
def ops():
    from hsimae_b200 import ops
    return ops


# ---------------------------------------------------------------- mask
@pytest.mark.parametrize("n,T,L,lt,ll", [(64, 4, 9, 3, 6), (64, 4, 9, 2, 9), (4096, 4, 9, 3, 6), (33, 4, 9, 2, 4), (33, 4, 9, 4, 2),
                                         (7, 4, 9, 3, 3), (5, 4, 9, 2, 2), (9, 4, 9, 4, 9), (16, 8, 16, 3, 5), (1, 2, 4, 2, 2)])
def test_mask_bit_exact(ops, n, T, L, lt, ll):
    torch.manual_seed(n + lt)
    nt, nl = torch.rand(n, T), torch.rand(n, L)
    if n > 2:  # exact ties (undefined in the reference; lowest index first here)
        nt[1, 1] = nt[1, 0]
        nl[2, L - 1] = nl[2, 0]
        nl[0, :] = 0.5
    k, r, m = ops.mask(nt.to(DEV), nl.to(DEV), lt, ll)
    ko, ro, mo = O.structured_mask(nt, nl, lt, ll)
    assert k.dtype == torch.int64 and r.dtype == torch.int64 and m.dtype == torch.float32
    assert torch.equal(k.cpu(), ko) and torch.equal(r.cpu(), ro) and torch.equal(m.cpu(), mo)


def test_mask_random_shapes_with_quantised_noise(ops):
    """randomised sweep (shapes, visible shapes, noise quantised to a few levels => many exact ties): the stable
    lowest-index-first tie rule must hold everywhere, bit for bit"""
    import random as pyrandom
    rng = pyrandom.Random(1234)
    g = torch.Generator().manual_seed(99)
    for _ in range(40):
        T, L, n = rng.randint(2, 8), rng.randint(2, 16), rng.randint(1, 300)
        lt, ll = rng.randint(1, T), rng.randint(1, L)
        levels = rng.choice([2, 3, 5, 1000000])
        nt = torch.floor(torch.rand(n, T, generator=g) * levels) / levels
        nl = torch.floor(torch.rand(n, L, generator=g) * levels) / levels
        k, r, m = ops.mask(nt.to(DEV), nl.to(DEV), lt, ll)
        ko, ro, mo = O.structured_mask(nt, nl, lt, ll)
        assert torch.equal(k.cpu(), ko) and torch.equal(r.cpu(), ro) and torch.equal(m.cpu(), mo), (T, L, n, lt, ll, levels)
        assert torch.equal(torch.sort(r.cpu(), 1).values, torch.arange(T * L).expand(n, -1))
        assert int(m.sum()) == n * (T * L - lt * ll)


def test_mask_kat(ops, golden):
    import hashlib
    z = golden("kat_masks.npz")
    k, r, m = ops.mask(torch.from_numpy(z["noise_t"]).to(DEV), torch.from_numpy(z["noise_l"]).to(DEV), 3, 6)
    assert hashlib.sha1(k.cpu().numpy().tobytes()).hexdigest()[:16] == "5ee5fac47baaef6e"
    assert hashlib.sha1(r.cpu().numpy().tobytes()).hexdigest()[:16] == "466597abc0b24641"
    assert torch.equal(k.cpu(), torch.from_numpy(z["large_ids_keep"]))
    assert torch.equal(m.cpu(), torch.from_numpy(z["large_mask"]))


def test_mask_empty_batch(ops):
    k, r, m = ops.mask(torch.empty(0, 4, device=DEV), torch.empty(0, 9, device=DEV), 3, 6)
    assert k.shape == (0, 18) and r.shape == (0, 36)


# ---------------------------------------------------------------- GEMM epilogues
GEMM_SHAPES = [(1000, 768, 256), (256, 256, 256), (130, 64, 64), (777, 192, 64), (512, 80, 64), (300, 256, 1376),
               (300, 64, 352), (129, 256, 768), (1, 128, 128), (2048, 144, 144), (640, 256, 80)]


@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
@pytest.mark.parametrize("epi", [0, 1])
def test_gemm_bias(ops, M, N, K, epi):
    A, B = _rand_bf16(M, K, seed=1), _rand_bf16(N, K, scale=0.05, seed=2)
    bias = torch.randn(N, device=DEV)
    ref = A.float() @ B.float().t() + bias
    tc = ops.gemm(A, B, epi, bias=bias)["out"]
    ck = ops.gemm(A, B, epi, bias=bias, impl=ops.IMPL_SIMT)["out"]
    tol = TOL_BF16 if epi == 0 else TOL_F32
    assert rel_err(ck.float(), ref) < tol
    assert rel_err(tc.float(), ref) < tol
    nobias = ops.gemm(A, B, epi)["out"]
    assert rel_err(nobias.float(), ref - bias) < tol


def test_gemm_strided_operands(ops):
    """operands that are column slices of wider buffers (lda != K), K not a multiple of 64"""
    M, N, K = 500, 176, 72
    Abig, Bbig = _rand_bf16(M, 80, seed=3), _rand_bf16(N, 80, scale=0.1, seed=4)
    A, B = Abig[:, :K], Bbig[:, :K]
    ref = A.float() @ B.float().t()
    assert rel_err(ops.gemm(A, B, 1)["out"], ref) < TOL_F32
    assert rel_err(ops.gemm(A, B, 1, impl=ops.IMPL_SIMT)["out"], ref) < TOL_F32


@pytest.mark.parametrize("M,N,K", [(1000, 256, 256), (300, 256, 688), (4096, 64, 176), (129, 128, 352), (77, 144, 144), (260, 64, 64)])
@pytest.mark.parametrize("variant", ["plain", "resid2+scale", "no_ln"])
def test_gemm_resid_layernorm(ops, M, N, K, variant):
    A, B = _rand_bf16(M, K, seed=5), _rand_bf16(N, K, scale=0.05, seed=6)
    bias = 0.1 * torch.randn(N, device=DEV)
    resid = 6.0 * torch.randn(M, N, device=DEV)
    gamma, beta = 1 + 0.1 * torch.randn(N, device=DEV), 0.1 * torch.randn(N, device=DEV)
    kw = dict(bias=bias, resid=resid, gamma=gamma, beta=beta)
    y = A.float() @ B.float().t() + bias
    if variant == "resid2+scale":
        Ktok, ll, G = 6, 3, 2          # spatial-mode row groups: rows -> (b, t)
        Mpad = (M + Ktok - 1) // Ktok
        scale = (torch.rand(Mpad * G, device=DEV) > 0.3).float() / 0.7
        r2 = torch.randn(M, N, device=DEV)
        kw.update(resid2=r2, rowscale=scale, rs_mode=1, rs_K=Ktok, rs_len_l=ll, rs_G=G)
        rows = torch.arange(M, device=DEV)
        s = scale[(rows // Ktok) * G + (rows % Ktok) // ll]
        x = resid + s[:, None] * y + r2
    else:
        x = resid + y
    if variant == "no_ln":
        kw.update(gamma=None, beta=None)
    for impl in (ops.IMPL_SIMT, ops.IMPL_TC):
        out = ops.gemm(A, B, ops.EPI_RESID_LN, impl=impl, **kw)
        assert rel_err(out["x"], x) < TOL_F32, impl
        if variant != "no_ln":
            ln = F.layer_norm(x, (N,), gamma, beta, 1e-5)
            assert rel_err(out["ln"].float(), ln) < TOL_BF16, impl
            mean, var = x.mean(1), x.var(1, unbiased=False)
            assert torch.allclose(out["stats"][:, 0], mean, atol=1e-4)
            assert torch.allclose(out["stats"][:, 1], (var + 1e-5).rsqrt(), rtol=1e-4)


@pytest.mark.parametrize("M,N,K", [(1000, 256, 768), (300, 256, 1376), (129, 192, 576), (40000, 256, 768), (77, 160, 64),
                                   (300, 64, 192), (5000, 64, 352), (1000, 32, 64), (1000, 128, 384)])   # decoder-width rows too
@pytest.mark.parametrize("variant", ["plain", "inplace+scale", "no_dxb"])
def test_gemm_lnbwd(ops, M, N, K, variant):
    _lnbwd_case(ops, M, N, K, variant)


def _lnbwd_case(ops, M, N, K, variant):
    """dgrad GEMM + LayerNorm backward in one launch (kEpiLnBwd) vs torch autograd of F.layer_norm in fp32 on the same
    bf16 operands.  Inside the kernel (t, xhat) make one round trip through a bf16 pair: tolerance = bf16 rounding."""
    A, B = _rand_bf16(M, K, seed=15), _rand_bf16(N, K, scale=0.05, seed=16)
    x = (3.0 * torch.randn(M, N, device=DEV) + torch.randn(M, 1, device=DEV)).requires_grad_(True)
    gamma = (1 + 0.2 * torch.randn(N, device=DEV)).requires_grad_(True)
    beta = torch.zeros(N, device=DEV, requires_grad=True)
    dx_in = torch.randn(M, N, device=DEV)
    dy = A.float() @ B.float().t()
    F.layer_norm(x, (N,), gamma, beta, 1e-5).backward(dy)
    xd = x.detach()
    mean, var = xd.mean(1), xd.var(1, unbiased=False)
    stats = torch.stack([mean, (var + 1e-5).rsqrt()], 1).contiguous()
    ref_dx = dx_in + x.grad
    kw = {}
    s = None
    if variant == "inplace+scale":
        Ktok, ll, G = 6, 3, 2
        Mpad = (M + Ktok - 1) // Ktok
        scale = (torch.rand(Mpad * G, device=DEV) > 0.3).float() / 0.7
        rows = torch.arange(M, device=DEV)
        s = scale[(rows // Ktok) * G + (rows % Ktok) // ll]
        kw.update(rowscale=scale, rs_mode=1, rs_K=Ktok, rs_len_l=ll, rs_G=G, inplace=True)
    if variant == "no_dxb":
        kw.update(want_dxb=False)
    dg0, db0 = torch.randn(N, device=DEV), torch.randn(N, device=DEV)      # the gradients accumulate
    out = ops.gemm_lnbwd(A, B, xd, stats, gamma.detach(), dx_in.clone(), dgamma=dg0.clone(), dbeta=db0.clone(), **kw)
    assert rel_err(out["dx"], ref_dx) < TOL_BF16
    if variant != "no_dxb":
        ref_b = ref_dx * s[:, None] if s is not None else ref_dx
        assert rel_err(out["dxb"].float(), ref_b) < 2 * TOL_BF16
    assert rel_err(out["dgamma"] - dg0, gamma.grad) < 1e-3
    assert rel_err(out["dbeta"] - db0, beta.grad) < 1e-3


@pytest.mark.parametrize("M,d,H", [(1000, 256, 684), (300, 64, 172), (129, 128, 344), (50, 144, 384)])
def test_gemm_swiglu_fwd_bwd(ops, M, d, H):
    hp = (H + 15) // 16 * 16
    x = _rand_bf16(M, d, seed=7)
    w1, w3 = _rand_bf16(H, d, scale=0.08, seed=8), _rand_bf16(H, d, scale=0.08, seed=9)
    b1, b3 = 0.1 * torch.randn(H, device=DEV), 0.1 * torch.randn(H, device=DEV)
    W = ops.pack_interleaved(w1, w3, hp)
    bias = ops.pack_interleaved(b1[:, None], b3[:, None], hp)[:, 0].contiguous()
    a_ref = x.float() @ w1.float().t() + b1
    b_ref = x.float() @ w3.float().t() + b3
    idx = torch.arange(H, device=DEV)
    ca, cb = (idx // 16) * 32 + idx % 16, (idx // 16) * 32 + 16 + idx % 16
    outs = {}
    for impl in (ops.IMPL_SIMT, ops.IMPL_TC):
        o = ops.gemm(x, W, ops.EPI_SWIGLU, impl=impl, bias=bias)
        outs[impl] = o
        assert rel_err(o["ab"][:, ca].float(), a_ref) < TOL_BF16
        assert rel_err(o["ab"][:, cb].float(), b_ref) < TOL_BF16
        a16, b16 = o["ab"][:, ca].float(), o["ab"][:, cb].float()
        assert rel_err(o["g"][:, :H].float(), F.silu(a16) * b16) < TOL_BF16
        if hp > H:
            assert float(o["g"][:, H:].abs().max()) == 0.0
    # backward of the gate: dg = dy @ W2 (N = hp columns), fused with d(silu(a)*b)
    dy = _rand_bf16(M, d, scale=0.1, seed=10)
    w2 = _rand_bf16(d, H, scale=0.08, seed=11)
    w2t = torch.zeros(hp, d, dtype=torch.bfloat16, device=DEV)
    w2t[:H] = w2.t()
    ab = outs[ops.IMPL_TC]["ab"]
    a16, b16 = ab[:, ca].float().requires_grad_(True), ab[:, cb].float().requires_grad_(True)
    dg = dy.float() @ w2.float()
    (F.silu(a16) * b16).backward(dg)
    for impl in (ops.IMPL_SIMT, ops.IMPL_TC):
        o = ops.gemm(dy, w2t, ops.EPI_DSWIGLU, impl=impl, ab=ab)
        assert rel_err(o["dab"][:, ca].float(), a16.grad) < TOL_BF16
        assert rel_err(o["dab"][:, cb].float(), b16.grad) < TOL_BF16


@pytest.mark.parametrize("M,d,H", [(1000, 256, 684), (300, 64, 172), (129, 128, 344), (4099, 256, 684), (77, 64, 172)])
def test_gemm_gate_recompute(ops, M, d, H):
    """Training path: forward keeps only g (gated on the fp32 pre-activations), backward recomputes a|b inside the
    d(gate) kernel.  Reference: torch fp32 autograd of silu(x W1^T + b1) * (x W3^T + b3) on the same bf16 inputs."""
    hp = (H + 15) // 16 * 16
    x = _rand_bf16(M, d, seed=7)
    w1, w3 = _rand_bf16(H, d, scale=0.08, seed=8), _rand_bf16(H, d, scale=0.08, seed=9)
    b1, b3 = 0.1 * torch.randn(H, device=DEV), 0.1 * torch.randn(H, device=DEV)
    W = ops.pack_interleaved(w1, w3, hp)
    bias = ops.pack_interleaved(b1[:, None], b3[:, None], hp)[:, 0].contiguous()
    a_ref = (x.float() @ w1.float().t() + b1).requires_grad_(True)
    b_ref = (x.float() @ w3.float().t() + b3).requires_grad_(True)
    g_ref = F.silu(a_ref) * b_ref
    o = ops.gemm(x, W, ops.EPI_SWIGLU, bias=bias, keep_ab=False)
    assert o["ab"] is None
    assert rel_err(o["g"][:, :H].float(), g_ref.detach()) < TOL_BF16
    if hp > H:
        assert float(o["g"][:, H:].abs().max()) == 0.0
    dy = _rand_bf16(M, d, scale=0.1, seed=10)
    w2 = _rand_bf16(d, H, scale=0.08, seed=11)
    w2t = torch.zeros(hp, d, dtype=torch.bfloat16, device=DEV)
    w2t[:H] = w2.t()
    g_ref.backward(dy.float() @ w2.float())
    idx = torch.arange(H, device=DEV)
    ca, cb = (idx // 16) * 32 + idx % 16, (idx // 16) * 32 + 16 + idx % 16
    dab = ops.gemm(dy, w2t, ops.EPI_DGATE, A2=x, B2=W, bias=bias)["dab"]
    assert rel_err(dab[:, ca].float(), a_ref.grad) < TOL_BF16
    assert rel_err(dab[:, cb].float(), b_ref.grad) < TOL_BF16
    if hp > H:   # padded hidden units: zero weights and bias => zero gradient
        pad = torch.ones(2 * hp, dtype=torch.bool, device=DEV)
        pad[ca] = False
        pad[cb] = False
        assert float(dab[:, pad].float().abs().max()) == 0.0


# ---------------------------------------------------------------- wgrad
@pytest.mark.parametrize("Mred,Nout,Kin", [(5000, 256, 256), (1000, 768, 256), (4097, 64, 64), (3000, 256, 688), (999, 192, 64),
                                           (2000, 64, 176), (63, 128, 128), (1, 64, 64), (8192, 144, 144)])
def test_wgrad_plain(ops, Mred, Nout, Kin):
    Y, X = _rand_bf16(Mred, Nout, scale=0.1, seed=12), _rand_bf16(Mred, Kin, seed=13)
    cols = Kin - 4 if Kin == 688 else Kin
    ref = Y.float().t() @ X.float()
    bref = Y.float().sum(0)
    for impl in (ops.IMPL_SIMT, ops.IMPL_TC):
        dst = torch.ones(Nout, cols, device=DEV)
        b = torch.zeros(Nout, device=DEV)
        ops.wgrad(Y, X, dst, impl=impl, bias0=b, cols_valid=cols)
        assert rel_err(dst - 1.0, ref[:, :cols]) < 2e-4, impl          # accumulates on top of the existing value
        assert rel_err(b, bref) < 2e-4, impl


def test_wgrad_interleaved_and_row_clip(ops):
    Mred, H, d = 3000, 172, 64
    hp = 176
    dab = _rand_bf16(Mred, 2 * hp, scale=0.1, seed=14)
    X = _rand_bf16(Mred, d, seed=15)
    idx = torch.arange(H, device=DEV)
    ca, cb = (idx // 16) * 32 + idx % 16, (idx // 16) * 32 + 16 + idx % 16
    for impl in (ops.IMPL_SIMT, ops.IMPL_TC):
        g1, g3 = torch.zeros(H, d, device=DEV), torch.zeros(H, d, device=DEV)
        b1, b3 = torch.zeros(H, device=DEV), torch.zeros(H, device=DEV)
        ops.wgrad(dab, X, g1, impl=impl, dst1=g3, row_map=1, rows_valid=H, bias0=b1, bias1=b3)
        assert rel_err(g1, dab[:, ca].float().t() @ X.float()) < 2e-4
        assert rel_err(g3, dab[:, cb].float().t() @ X.float()) < 2e-4
        assert rel_err(b1, dab[:, ca].float().sum(0)) < 2e-4 and rel_err(b3, dab[:, cb].float().sum(0)) < 2e-4
    # decoder_pred: 80 packed rows of which 72 exist
    Y, X2 = _rand_bf16(2000, 80, scale=0.1, seed=16), _rand_bf16(2000, 64, seed=17)
    for impl in (ops.IMPL_SIMT, ops.IMPL_TC):
        dst = torch.zeros(72, 64, device=DEV)
        ops.wgrad(Y, X2, dst, impl=impl, rows_valid=72)
        assert rel_err(dst, (Y.float().t() @ X2.float())[:72]) < 2e-4


# ---------------------------------------------------------------- attention
def _attn_ref(qkv, n, D, heads, K, groups):
    """groups: LongTensor [nseq, len] of row offsets inside a sample"""
    hd = D // heads
    q, k, v = qkv.float().reshape(n, K, 3, heads, hd).unbind(2)
    out = torch.zeros(n, K, heads, hd, device=qkv.device)
    for rows in groups:
        qs, ks, vs = q[:, rows].transpose(1, 2), k[:, rows].transpose(1, 2), v[:, rows].transpose(1, 2)
        att = torch.softmax(qs @ ks.transpose(-1, -2) * hd ** -0.5, -1)
        out[:, rows] = (att @ vs).transpose(1, 2)
    return out.reshape(n * K, D)


@pytest.mark.parametrize("n,D,heads,lt,ll,kind", [
    (37, 256, 16, 3, 6, "spatial"), (37, 256, 16, 3, 6, "spectral"), (37, 256, 16, 3, 6, "full"),
    (20, 256, 16, 2, 9, "spatial"), (20, 256, 16, 2, 9, "spectral"), (300, 64, 8, 4, 9, "full"),
    (9, 128, 8, 4, 9, "spatial"), (9, 128, 8, 4, 9, "spectral"), (5, 64, 4, 2, 4, "spectral"), (1, 64, 4, 4, 2, "spatial"),
    # head dimension 32 (two k-steps per score tile), sequences of 18 / 36 / 6 tokens; 3 query tiles with head dimension 16
    (9, 128, 4, 3, 6, "full"), (7, 256, 8, 4, 9, "full"), (9, 128, 4, 3, 6, "spatial"), (11, 256, 16, 4, 9, "full")])
def test_attention_fwd_bwd(ops, n, D, heads, lt, ll, kind):
    K = lt * ll
    if kind == "spatial":
        spec = (lt, ll, ll, 1); groups = [torch.arange(ll) + t * ll for t in range(lt)]
    elif kind == "spectral":
        spec = (ll, lt, 1, ll); groups = [torch.arange(lt) * ll + l for l in range(ll)]
    else:
        spec = (1, K, K, 1); groups = [torch.arange(K)]
    groups = [g.to(DEV) for g in groups]
    qkv = _rand_bf16(n * K, 3 * D, seed=18)
    out, lse = ops.attention_forward(qkv, n, D, heads, K, *spec)
    leaf = qkv.float().requires_grad_(True)
    ref = _attn_ref(leaf, n, D, heads, K, groups)
    assert rel_err(out.float(), ref) < TOL_BF16
    dout = _rand_bf16(n * K, D, scale=0.1, seed=19)
    ref.backward(dout.float())
    dqkv = ops.attention_backward(qkv, out, lse, dout, n, D, heads, K, *spec)
    assert rel_err(dqkv.float(), leaf.grad) < 2e-2   # bf16 O/dO inputs + bf16 output


@pytest.mark.parametrize("M,d,H", [(73728 // 4, 256, 684), (40000, 64, 172), (700, 128, 344)])
def test_wgrad_group_matches_single_launches(ops, M, d, H):
    """The four weight gradients of a block (dW2, dW1|dW3, dWproj, dWq|k|v + bias gradients) in one grouped launch ==
    the same four problems launched one by one == torch fp32 on the same bf16 operands."""
    hp = (H + 15) // 16 * 16
    dy, g = _rand_bf16(M, d, scale=0.1, seed=1), _rand_bf16(M, hp, scale=0.5, seed=2)
    dab, ln2 = _rand_bf16(M, 2 * hp, scale=0.1, seed=3), _rand_bf16(M, d, seed=4)
    dxm, ao = _rand_bf16(M, d, scale=0.1, seed=5), _rand_bf16(M, d, seed=6)
    dqkv, ln1 = _rand_bf16(M, 3 * d, scale=0.1, seed=7), _rand_bf16(M, d, seed=8)

    def fresh():
        z = lambda *s: torch.zeros(*s, device=DEV)
        return dict(w2=z(d, H), b2=z(d), w1=z(H, d), w3=z(H, d), b1=z(H), b3=z(H), wp=z(d, d), bp=z(d), wq=z(3 * d, d), bq=z(3 * d))

    def jobs(o):
        return [dict(Y=dy, X=g, dst0=o["w2"], cols_valid=H, bias0=o["b2"]),
                dict(Y=dab, X=ln2, dst0=o["w1"], dst1=o["w3"], row_map=1, rows_valid=H, bias0=o["b1"], bias1=o["b3"]),
                dict(Y=dxm, X=ao, dst0=o["wp"], bias0=o["bp"]),
                dict(Y=dqkv, X=ln1, dst0=o["wq"], bias0=o["bq"])]
    a, b = fresh(), fresh()
    ops.wgrad_group(jobs(a))
    for j in jobs(b):
        ops.wgrad(j.pop("Y"), j.pop("X"), j.pop("dst0"), **j)
    for k in a:
        assert rel_err(a[k], b[k]) < 2e-5, k          # same MMAs, different reduction-split boundaries / atomic order
    assert rel_err(a["w2"], dy.float().t() @ g.float()[:, :H]) < 2e-4
    assert rel_err(a["wp"], dxm.float().t() @ ao.float()) < 2e-4
    assert rel_err(a["wq"], dqkv.float().t() @ ln1.float()) < 2e-4
    assert rel_err(a["bq"], dqkv.float().sum(0)) < 2e-4
    idx = torch.arange(H, device=DEV)
    ca, cb = (idx // 16) * 32 + idx % 16, (idx // 16) * 32 + 16 + idx % 16
    assert rel_err(a["w1"], dab.float()[:, ca].t() @ ln2.float()) < 2e-4
    assert rel_err(a["w3"], dab.float()[:, cb].t() @ ln2.float()) < 2e-4


# ---------------------------------------------------------------- fused gated MLP (csrc/block_fused.cu)
@pytest.mark.parametrize("M,d,H,pair", [(1000, 256, 684, 1), (300, 64, 172, 1), (129, 128, 344, 1), (4099, 256, 684, 1), (77, 64, 172, 1),
                                        (128, 64, 172, 1), (37000, 256, 684, 1), (40000, 64, 172, 1), (300, 64, 172, 0), (700, 128, 344, 0)])
@pytest.mark.parametrize("variant", ["ln", "plain", "resid2_rs"])
def test_mlp_fused(ops, M, d, H, pair, variant, monkeypatch):
    """One kernel for  x + rs * (w2(silu(w1 h) * w3 h) + b2) [+ resid2]  and the next LayerNorm (Models.py:231-232, 305).
    Reference: torch fp32 on the same bf16 inputs with the gate output rounded to bf16 where the kernel rounds it
    (it is the bf16 A operand of the down-projection)."""
    if pair == 0:
        pytest.skip("single-CTA form is selected once per process (HSIMAE_FUSED_MLP_PAIR); covered by test_switches_gpu")
    hp = (H + 15) // 16 * 16
    h = _rand_bf16(M, d, seed=7)
    w1, w3 = _rand_bf16(H, d, scale=0.08, seed=8), _rand_bf16(H, d, scale=0.08, seed=9)
    g = torch.Generator(device=DEV).manual_seed(12)
    b1, b3 = 0.1 * torch.randn(H, device=DEV, generator=g), 0.1 * torch.randn(H, device=DEV, generator=g)
    w2 = torch.zeros(d, hp, dtype=torch.bfloat16, device=DEV)
    w2[:, :H] = _rand_bf16(d, H, scale=0.08, seed=11)
    b2 = 0.1 * torch.randn(d, device=DEV, generator=g)
    resid = 3.0 * torch.randn(M, d, device=DEV, generator=g)
    W13 = ops.pack_interleaved(w1, w3, hp)
    b13 = ops.pack_interleaved(b1[:, None], b3[:, None], hp)[:, 0].contiguous()
    gamma = beta = resid2 = rowscale = None
    kw = {}
    if variant != "plain":
        gamma, beta = 1 + 0.1 * torch.randn(d, device=DEV, generator=g), 0.1 * torch.randn(d, device=DEV, generator=g)
    if variant == "resid2_rs":
        resid2 = torch.randn(M, d, device=DEV, generator=g)
        K = 18
        nb = (M + K - 1) // K
        rowscale = (torch.rand(nb, device=DEV, generator=g) > 0.3).float() / 0.7
        kw = dict(rowscale=rowscale, rs_mode=3, rs_K=K, rs_len_l=9, rs_G=1)
    g_ref = F.silu(h.float() @ w1.float().t() + b1) * (h.float() @ w3.float().t() + b3)
    y = g_ref.to(torch.bfloat16).float() @ w2[:, :H].float().t() + b2
    if rowscale is not None:
        y = y * rowscale[torch.arange(M, device=DEV) // 18][:, None]
    x_ref = resid + y + (resid2 if resid2 is not None else 0)
    o = ops.mlp_fused(h, W13, b13, w2, b2, resid, gamma=gamma, beta=beta, resid2=resid2, **kw)
    assert rel_err(o["g"][:, :H].float(), g_ref) < TOL_BF16
    if hp > H:
        assert float(o["g"][:, H:].abs().max()) == 0.0
    assert rel_err(o["x"], x_ref) < 2e-4          # fp32 stream; the bf16 gate operand bounds the error of the update
    if gamma is not None:
        ln_ref = F.layer_norm(x_ref, (d,), gamma, beta, 1e-5)
        assert rel_err(o["ln"].float(), ln_ref) < TOL_BF16
        mean, var = x_ref.mean(-1), x_ref.var(-1, unbiased=False)
        assert rel_err(o["stats"][:, 0], mean) < 1e-3 and rel_err(o["stats"][:, 1], torch.rsqrt(var + 1e-5)) < 1e-3
    # inference form: no gate output, same stream; and the two-launch path it replaces
    o2 = ops.mlp_fused(h, W13, b13, w2, b2, resid, gamma=gamma, beta=beta, resid2=resid2, keep_g=False, **kw)
    assert o2["g"] is None and torch.equal(o2["x"], o["x"])
    u = ops.gemm(h, W13, ops.EPI_SWIGLU, bias=b13, keep_ab=False)
    v = ops.gemm(u["g"], w2, ops.EPI_RESID_LN, bias=b2, resid=resid, resid2=resid2, gamma=gamma, beta=beta, **kw)
    assert rel_err(o["x"], v["x"]) < 2e-5
