"""B200-native HSIMAE / DualViT / HSIViT behind the reference's `Models.py` surface.

Same constructor keywords, forward signatures, return tuples, parameter names /
shapes / registration order (so ``state_dict`` round-trips with reference
checkpoints and seeded construction yields identical weights) as
/root/reference/Models.py:309-634 (HSIMAE), :637-993 (DualViT), :996-1160
(HSIViT).  The modules are parameter containers; ``forward`` runs the sm_100a
CUDA library (hsimae_b200/csrc) through the C ABI in include/hsimae_b200.h.
There is no CPU implementation: calling ``forward`` without a CUDA device or
without the built library raises.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib
from .host import choose_visible_shape, draw_drop_factors, sincos_table, swiglu_hidden

__all__ = ["HSIMAE", "DualViT", "HSIViT", "PatchEmbed", "Attention", "SwiGLU", "Block", "DropPath"]


# ---------------------------------------------------------------------------
# parameter containers (names and construction order follow the reference)
# ---------------------------------------------------------------------------
class PatchEmbed(nn.Module):
    """Geometry + the Conv3d parameters of the patch embedding (Models.py:104-149)."""

    def __init__(self, img_size=224, patch_size=16, bands=32, b_patch_size=8, in_chans=1, embed_dim=768):
        super().__init__()
        if img_size % patch_size or bands % b_patch_size:
            raise AssertionError("image / band extents must be divisible by the patch extents")
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.bands, self.b_patch_size = bands, b_patch_size
        self.grid_size = img_size // patch_size
        self.b_grid_size = bands // b_patch_size
        self.input_size = (self.b_grid_size, self.grid_size, self.grid_size)
        self.num_patches = self.b_grid_size * self.grid_size ** 2
        self.in_chans = in_chans
        k = [b_patch_size, patch_size, patch_size]
        self.proj = nn.Conv3d(in_chans, embed_dim, kernel_size=k, stride=k)
        self.output_size = None


class Attention(nn.Module):
    """q / k / v / proj Linear parameters (Models.py:163-187)."""

    def __init__(self, dim, num_heads=8, qkv_bias=False):
        super().__init__()
        if dim % num_heads:
            raise AssertionError("dim should be divisible by num_heads")
        self.num_heads = num_heads
        self.scale = (dim // num_heads) ** -0.5
        self.q = nn.Linear(dim, dim, bias=qkv_bias)
        self.k = nn.Linear(dim, dim, bias=qkv_bias)
        self.v = nn.Linear(dim, dim, bias=qkv_bias)
        self.proj = nn.Linear(dim, dim)


class SwiGLU(nn.Module):
    """w1 / w2 / w3 Linear parameters of the gated MLP (Models.py:222-229)."""

    def __init__(self, dim, mlp_ratio):
        super().__init__()
        hidden = swiglu_hidden(dim, mlp_ratio)
        self.w1 = nn.Linear(dim, hidden, bias=True)
        self.w2 = nn.Linear(hidden, dim, bias=True)
        self.w3 = nn.Linear(dim, hidden, bias=True)


class DropPath(nn.Module):
    """Stochastic-depth rate holder (Models.py:254-266); the draw happens in host.draw_drop_factors."""

    def __init__(self, drop_prob: float = 0.0):
        super().__init__()
        self.drop_prob = drop_prob

    def extra_repr(self):
        return f"drop_prob={round(self.drop_prob, 3):0.3f}"


class Block(nn.Module):
    """Pre-LN block parameters (Models.py:269-301)."""

    def __init__(self, dim, num_heads, mlp_ratio=4.0, qkv_bias=False, norm_layer=nn.LayerNorm, drop_path=0.0):
        super().__init__()
        self.norm1 = norm_layer(dim)
        self.attn = Attention(dim, num_heads=num_heads, qkv_bias=qkv_bias)
        self.drop_path = DropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.norm2 = norm_layer(dim)
        self.mlp = SwiGLU(dim, mlp_ratio)

    @property
    def drop_rate(self) -> float:
        return float(getattr(self.drop_path, "drop_prob", 0.0))


# ---------------------------------------------------------------------------
# runtime: plan, arenas, workspaces
# ---------------------------------------------------------------------------
def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


_PROCESS_DEVICE = None   # the library configures its kernels (shared-memory limits, TMA descriptors) once per process


class _Runtime:
    """Per-model handle on the native plan plus the device arenas it packs into."""

    def __init__(self, dims: _lib.Dims):
        self.lib = _lib.load()
        self.dims = dims
        handle = C.c_void_p()
        _lib.check(self.lib.hsimae_plan_create(C.byref(dims), C.byref(handle)), "plan_create")
        self.plan = handle
        L = self.lib
        self.n_params = L.hsimae_plan_num_params(self.plan)
        self.names = [L.hsimae_plan_param_name(self.plan, i).decode() for i in range(self.n_params)]
        self.numels = [L.hsimae_plan_param_numel(self.plan, i) for i in range(self.n_params)]
        self.grad_off = [L.hsimae_plan_param_grad_offset(self.plan, i) for i in range(self.n_params)]
        self.grad_elems = L.hsimae_plan_grad_arena_elems(self.plan)
        self.buckets = []
        for i in range(6):   # arena order: pe + spatial thirds (0, 1, 2), spectral (3), fusion + norm + head (4), decoder (5)
            off, n = C.c_int64(), C.c_int64()
            _lib.check(L.hsimae_plan_grad_bucket(self.plan, i, C.byref(off), C.byref(n)), "grad_bucket")
            self.buckets.append((off.value, n.value))
        self.device = None
        self.wb = self.wf = self.table = None
        self.sig = None
        self.checked_ptrs = None
        self.ptr_array = (C.c_void_p * self.n_params)()

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                self.lib.hsimae_plan_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    def ensure_device(self, device: torch.device):
        if self.device == device:
            return
        global _PROCESS_DEVICE
        if _PROCESS_DEVICE is None:
            _PROCESS_DEVICE = device
        elif _PROCESS_DEVICE != device:
            raise RuntimeError(f"hsimae_b200 runs one process per GPU: this process already uses {_PROCESS_DEVICE}, "
                               f"got a model on {device} (launch one rank per device, e.g. with torchrun)")
        L = self.lib
        self.wb = torch.zeros(L.hsimae_plan_bf16_arena_elems(self.plan), dtype=torch.bfloat16, device=device)
        self.wf = torch.zeros(L.hsimae_plan_f32_arena_elems(self.plan), dtype=torch.float32, device=device)
        self.table = torch.zeros(max(L.hsimae_plan_pack_table_bytes(self.plan), 16), dtype=torch.uint8, device=device)
        self.device = device
        self.sig = None

    def pack(self, params: List[torch.Tensor], ptrs=None):
        """Refresh the packed bf16 / fp32 operand arenas if any parameter changed."""
        if ptrs is None:
            ptrs = [p.data_ptr() for p in params]
        sig = (ptrs, [p._version for p in params])
        if sig == self.sig:
            return
        if self.sig is None or ptrs != self.sig[0]:
            self.ptr_array[:] = ptrs
        _lib.check(self.lib.hsimae_pack_params(self.plan, self.ptr_array, _ptr(self.wb), _ptr(self.wf), _ptr(self.table),
                                               _stream()), "pack_params")
        self.sig = sig

    # -- thin wrappers -------------------------------------------------------
    def enc_ws(self, n, lt, ll, save, device):
        nbytes = self.lib.hsimae_encoder_workspace_bytes(self.plan, n, lt, ll, int(save))
        return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)

    def dec_ws(self, n, lt, ll, save, device):
        nbytes = self.lib.hsimae_decoder_workspace_bytes(self.plan, n, lt, ll, int(save))
        return torch.empty(max(nbytes, 256), dtype=torch.uint8, device=device)

    def drop_table(self, drops):
        if drops is None or all(d is None for d in drops):
            return None
        arr = (C.c_void_p * len(drops))()
        for i, d in enumerate(drops):
            arr[i] = d.data_ptr() if d is not None else None
        return arr

    def encoder_forward(self, imgs, n, lt, ll, ids32, drops, save, ws):
        tab = self.drop_table(drops)
        _lib.check(self.lib.hsimae_encoder_forward(self.plan, _ptr(self.wb), _ptr(self.wf), _ptr(imgs), n, lt, ll, _ptr(ids32),
                                                   tab, int(save), _ptr(ws), ws.numel(), _stream()), "encoder_forward")

    def encoder_backward(self, imgs, n, lt, ll, ids32, drops, ws, grads, stages=31):
        tab = self.drop_table(drops)
        _lib.check(self.lib.hsimae_encoder_backward(self.plan, _ptr(self.wb), _ptr(self.wf), _ptr(imgs), n, lt, ll, _ptr(ids32),
                                                    tab, _ptr(ws), ws.numel(), _ptr(grads), stages, _stream()),
                   "encoder_backward")

    def encoder_latent(self, n, lt, ll, save, ws):
        out = torch.empty(n, lt * ll, self.dims.embed_dim, dtype=torch.float32, device=ws.device)
        _lib.check(self.lib.hsimae_encoder_latent(self.plan, n, lt, ll, int(save), _ptr(self.wf), _ptr(ws), _ptr(out), _stream()),
                   "encoder_latent")
        return out

    def mask(self, noise_t, noise_l, lt, ll):
        n, T = noise_t.shape
        Lp = noise_l.shape[1]
        dev = noise_t.device
        K, P = lt * ll, T * Lp
        ids_keep = torch.empty(n, K, dtype=torch.int64, device=dev)
        ids_restore = torch.empty(n, P, dtype=torch.int64, device=dev)
        mask = torch.empty(n, P, dtype=torch.float32, device=dev)
        keep32 = torch.empty(n, K, dtype=torch.int32, device=dev)
        restore32 = torch.empty(n, P, dtype=torch.int32, device=dev)
        _lib.check(self.lib.hsimae_mask(_ptr(noise_t), _ptr(noise_l), n, T, Lp, lt, ll, _ptr(ids_keep), _ptr(ids_restore),
                                        _ptr(mask), _ptr(keep32), _ptr(restore32), _stream()), "mask")
        return ids_keep, ids_restore, mask, keep32, restore32


class _Saved:
    """Python-side record of one forward call (kept alive by the autograd node)."""
    __slots__ = ("rt", "imgs_full", "n_full", "ws_full", "drops_full", "pooled", "imgs_m", "n_m", "lt", "ll", "keep32",
                 "restore32", "ws_enc", "ws_dec", "drops_m", "dp", "consumed")


class _HsiFunction(torch.autograd.Function):
    """One autograd node for a whole model call.  Outputs: (loss, logits) where
    either may be a dummy when the corresponding branch is inactive; pixel
    outputs are returned out of band (they carry no gradient)."""

    @staticmethod
    def forward(ctx, saved: _Saved, loss: Optional[torch.Tensor], logits: Optional[torch.Tensor], *params):
        ctx.saved = saved
        ctx.n_params = len(params)
        ctx.set_materialize_grads(False)
        outs = []
        ctx.has_loss = loss is not None
        ctx.has_logits = logits is not None
        if loss is not None:
            outs.append(loss)
        if logits is not None:
            outs.append(logits)
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        s: _Saved = ctx.saved
        if s.consumed:
            # the stashed activations were released by the first backward (torch raises here as well)
            raise RuntimeError("hsimae_b200: trying to backward through the graph a second time: the saved activations of this "
                               "forward call have been freed (retain_graph / double backward are not supported); "
                               "sum the losses and call backward once, as Model_Finetuning.py:153-164 does")
        rt: _Runtime = s.rt
        gouts = list(gouts)
        g_loss = gouts.pop(0) if ctx.has_loss else None
        g_logits = gouts.pop(0) if ctx.has_logits else None
        dev = rt.device
        grads = torch.zeros(rt.grad_elems, dtype=torch.float32, device=dev)
        sync = s.dp.begin(grads) if s.dp is not None else None
        if g_loss is not None and s.ws_dec is not None:
            g = g_loss.detach().reshape(1).to(torch.float32).contiguous()
            _lib.check(rt.lib.hsimae_decoder_backward(rt.plan, _ptr(rt.wb), _ptr(rt.wf), s.n_m, s.lt, s.ll, _ptr(s.restore32),
                                                      _ptr(s.ws_enc), _ptr(s.ws_dec), s.ws_dec.numel(), _ptr(g), _ptr(grads),
                                                      _stream()), "decoder_backward")
            if sync is not None and g_logits is None:
                sync.ready(5)
            # backward-completion order of the arena regions; the spatial encoder goes out in thirds so that only the
            # last one (+ the patch embedding) is exposed after the final kernel
            # The spectral chain (stage 2 | 32) runs on the library's helper stream beside the spatial thirds and is joined
            # before the last third: its region (bucket 3) is final on this stream after the join.
            for stage, bucket in ((1, 4), (2 | 32, None), (4, 2), (8, 1), (-1, 3), (16, 0)):
                if stage < 0:
                    _lib.check(rt.lib.hsimae_helper_join(rt.plan, _stream()), "helper_join")
                else:
                    rt.encoder_backward(s.imgs_m, s.n_m, s.lt, s.ll, s.keep32, s.drops_m, s.ws_enc, grads, stage)
                if sync is not None and g_logits is None and bucket is not None:
                    sync.ready(bucket)
        if g_logits is not None and s.ws_full is not None:
            gl = g_logits.detach().to(torch.float32).contiguous()
            T, Lp = rt.dims.bands // rt.dims.b_patch_size, (rt.dims.img_size // rt.dims.patch_size) ** 2
            _lib.check(rt.lib.hsimae_head_backward(rt.plan, _ptr(rt.wf), s.n_full, _ptr(s.ws_full), _ptr(s.pooled), _ptr(gl),
                                                   _ptr(grads), _stream()), "head_backward")
            rt.encoder_backward(s.imgs_full, s.n_full, T, Lp, None, s.drops_full, s.ws_full, grads, 31 | 32)
            if sync is not None:
                for b in (5, 4, 3, 2, 1, 0):
                    sync.ready(b)
        if sync is not None:
            sync.finish()
        # the stashed activations are consumed: release them now (no retain_graph / double backward)
        s.ws_enc = s.ws_dec = s.ws_full = None
        s.consumed = True
        out = [None, None, None]
        needs = ctx.needs_input_grad[3:]
        for i in range(ctx.n_params):
            off = rt.grad_off[i]
            if off < 0 or not needs[i]:
                out.append(None)
            else:
                out.append(grads[off:off + rt.numels[i]].view(s_shapes(ctx, i)))
        return tuple(out)


def s_shapes(ctx, i):
    return ctx.saved.rt.shapes[i]


# ---------------------------------------------------------------------------
# model base
# ---------------------------------------------------------------------------
class _HsiBase(nn.Module):
    def _build_encoder(self, img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_ratio, norm_layer, bands,
                       b_patch_size, no_qkv_bias, s_depth, dpr):
        self.patch_embed = PatchEmbed(img_size, patch_size, bands, b_patch_size, in_chans, embed_dim)
        self.input_size = self.patch_embed.input_size
        self.pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches, embed_dim))

        def blk(rate):
            return Block(embed_dim, num_heads, mlp_ratio, qkv_bias=not no_qkv_bias, norm_layer=norm_layer, drop_path=rate)

        if s_depth > 0:
            self.blocks_1 = nn.ModuleList([blk(dpr[i]) for i in range(s_depth)])
            self.blocks_2 = nn.ModuleList([blk(dpr[i]) for i in range(s_depth)])
        if s_depth < 12:  # literal 12, as in the reference (Models.py:385)
            self.blocks = nn.ModuleList([blk(dpr[i]) for i in range(s_depth, depth)])
        self.norm = norm_layer(embed_dim)

    def _build_decoder(self, embed_dim, decoder_embed_dim, decoder_depth, decoder_num_heads, mlp_ratio, norm_layer, no_qkv_bias,
                       patch_size, in_chans):
        self.decoder_blocks = nn.ModuleList([
            Block(decoder_embed_dim, decoder_num_heads, mlp_ratio, qkv_bias=not no_qkv_bias, norm_layer=norm_layer)
            for _ in range(decoder_depth)])
        self.decoder_norm = norm_layer(decoder_embed_dim)
        self.decoder_pred = nn.Linear(decoder_embed_dim, self.b_pred_patch_size * patch_size ** 2 * in_chans, bias=True)

    def _init_all(self, decoder: bool):
        T, G = self.input_size[0], self.input_size[1]
        self.pos_embed.data.copy_(sincos_table(self.dim, T, G))
        self.pos_embed.requires_grad = False
        if decoder:
            self.decoder_pos_embed.data.copy_(sincos_table(self.dec_dim, T, G))
            self.decoder_pos_embed.requires_grad = False
        w = self.patch_embed.proj.weight.data
        if self.trunc_init:
            torch.nn.init.trunc_normal_(w)
            if decoder:
                torch.nn.init.trunc_normal_(self.mask_token, std=0.02)
        else:
            torch.nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
            if decoder:
                torch.nn.init.normal_(self.mask_token, std=0.02)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            if self.trunc_init:
                nn.init.trunc_normal_(m.weight, std=0.02)
            else:
                torch.nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def __getstate__(self):
        st = self.__dict__.copy()
        for k in ("_rt", "_rt_accessors", "_dp", "_last", "_last_imgs", "_last_stats"):
            st.pop(k, None)
        return st

    # -- native runtime --------------------------------------------------------
    def _runtime(self) -> _Runtime:
        rt = self.__dict__.get("_rt")
        if rt is None:
            pe = self.patch_embed
            if pe.in_chans != 1:
                raise RuntimeError("hsimae_b200 supports in_chans=1 (as every reference call site uses)")
            for m in self.modules():
                if isinstance(m, nn.LayerNorm) and (m.eps != 1e-5 or not m.elementwise_affine):
                    raise RuntimeError("hsimae_b200 supports nn.LayerNorm(eps=1e-5, affine) as norm_layer")
            d = _lib.Dims()
            d.img_size, d.patch_size = pe.img_size[0], pe.patch_size[0]
            d.bands, d.b_patch_size = pe.bands, pe.b_patch_size
            d.embed_dim, d.depth, d.s_depth = self.dim, self._depth, self.s_depth
            d.num_heads = self._num_heads
            d.dec_dim = getattr(self, "dec_dim", 0) or 0
            d.dec_depth = len(self.decoder_blocks) if hasattr(self, "decoder_blocks") else 0
            d.dec_heads = self._dec_heads
            d.num_class = self.cls_head.out_features if hasattr(self, "cls_head") else 0
            d.qkv_bias = int(self._qkv_bias)
            d.norm_pix_loss = int(getattr(self, "norm_pix_loss", False))
            d.mlp_ratio = float(self._mlp_ratio)
            rt = _Runtime(d)
            named = dict(self.named_parameters())
            missing = [n for n in rt.names if n not in named]
            if missing:
                raise RuntimeError(f"hsimae_b200: parameters missing from the module: {missing[:4]}...")
            rt.shapes = [tuple(named[n].shape) for n in rt.names]
            self.__dict__["_rt"] = rt
        return rt

    def _plan_params(self, rt: _Runtime) -> List[torch.Tensor]:
        """The parameters in the plan's order.  Walking `named_parameters()` costs ~0.6 ms per call for the 535 tensors
        of HSIMAE-Large (exposed GPU idle time when the training loop synchronises every step, as the reference's
        does with `loss.item()`), so the owning (sub-module, attribute) pairs are resolved once and each call only
        re-reads the current Parameter objects from their owners -- replaced parameters, `.to()`, `load_state_dict`
        are all seen; swapping a whole sub-module for another requires `model._rt_accessors = None`."""
        acc = self.__dict__.get("_rt_accessors")
        if acc is None:
            owners = {}
            for prefix, mod in self.named_modules():
                for name in mod._parameters:
                    owners[(prefix + "." if prefix else "") + name] = (mod._parameters, name)
            acc = [owners[n] for n in rt.names]
            self.__dict__["_rt_accessors"] = acc
        return [d[k] for d, k in acc]

    def invalidate_weight_cache(self) -> None:
        """Force the next forward to re-validate and re-pack the parameters.  Needed only after writes torch cannot see:
        `p.data.<op>_()` (the `.data` alias has its own version counter) or raw-pointer writes from foreign code.
        `load_state_dict`, optimiser steps, `.to()` and Parameter replacement are detected automatically."""
        rt = self.__dict__.get("_rt")
        if rt is not None:
            rt.sig = rt.checked_ptrs = None
        self.__dict__["_rt_accessors"] = None

    def _prepare(self, imgs: torch.Tensor):
        if not imgs.is_cuda:
            raise RuntimeError("hsimae_b200 runs on a CUDA (sm_100a) device only; there is no CPU path. "
                               "Move the model and inputs to 'cuda'.")
        rt = self._runtime()
        params = self._plan_params(rt)
        dev = imgs.device
        if dev.index != torch.cuda.current_device():
            raise RuntimeError(f"hsimae_b200: inputs are on {dev} but the current CUDA device is cuda:{torch.cuda.current_device()}; "
                               "call torch.cuda.set_device(...) first (kernels are launched on the current device's stream)")
        # device / dtype / layout can only change together with the storage pointer: validate when a pointer moved
        ptrs = [p.data_ptr() for p in params]
        if ptrs != rt.checked_ptrs or dev != rt.device:
            for p in params:
                if p.device != dev or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError("hsimae_b200: parameters must be contiguous fp32 tensors on the input's CUDA device")
            rt.checked_ptrs = ptrs
        rt.ensure_device(dev)
        rt.pack(params, ptrs)
        return rt, params

    def _check_imgs(self, imgs):
        pe = self.patch_embed
        if imgs.dim() != 5 or imgs.shape[1] != 1 or imgs.shape[2] != pe.bands or imgs.shape[3] != pe.img_size[0] \
                or imgs.shape[4] != pe.img_size[1]:
            raise AssertionError(f"Input of shape {tuple(imgs.shape)} doesn't match model "
                                 f"([N,1,{pe.bands},{pe.img_size[0]},{pe.img_size[1]}]).")
        return imgs.detach().to(torch.float32).contiguous()

    def _split_rates(self):
        r1 = [b.drop_rate for b in self.blocks_1] if hasattr(self, "blocks_1") else []
        rf = [b.drop_rate for b in self.blocks] if hasattr(self, "blocks") else []
        return r1, rf

    def _wants_grad(self, params):
        return torch.is_grad_enabled() and any(p.requires_grad for p in params)

    # -- passes -------------------------------------------------------------------
    def _full_pass(self, rt, imgs, save):
        """Unmasked encoder + classification head (Models.py:869-894, 964-973)."""
        n = imgs.shape[0]
        T, G = self.input_size[0], self.input_size[1]
        Lp = G * G
        r1, rf = self._split_rates()
        drops = draw_drop_factors(r1, rf, n, T, Lp, imgs.device, self.training)
        ws = rt.enc_ws(n, T, Lp, save, imgs.device)
        rt.encoder_forward(imgs, n, T, Lp, None, drops, save, ws)
        self.patch_embed.output_size = torch.Size((n, T, Lp, self.dim))
        pooled = torch.empty(n, T * self.dim, dtype=torch.float32, device=imgs.device)
        logits = torch.empty(n, self.cls_head.out_features, dtype=torch.float32, device=imgs.device)
        _lib.check(rt.lib.hsimae_head_forward(rt.plan, _ptr(rt.wf), n, _ptr(ws), int(save), _ptr(pooled), _ptr(logits),
                                              _stream()), "head_forward")
        return logits, pooled, ws, drops

    def _masked_pass(self, rt, imgs, mask_ratio, save, want_tokens=False):
        """Masked encoder + decoder + loss + pixel outputs (Models.py:537-634)."""
        n = imgs.shape[0]
        T, G = self.input_size[0], self.input_size[1]
        Lp = G * G
        dev = imgs.device
        lt, ll = choose_visible_shape(T, Lp, mask_ratio)
        noise_t = torch.rand(n, T, device=dev)
        noise_l = torch.rand(n, Lp, device=dev)
        ids_keep, ids_restore, mask, keep32, restore32 = rt.mask(noise_t, noise_l, lt, ll)
        self.len_t, self.len_l = torch.tensor(lt), torch.tensor(ll)
        r1, rf = self._split_rates()
        drops = draw_drop_factors(r1, rf, n, lt, ll, dev, self.training)
        ws_enc = rt.enc_ws(n, lt, ll, save, dev)
        rt.encoder_forward(imgs, n, lt, ll, keep32, drops, save, ws_enc)
        self.patch_embed.output_size = torch.Size((n, T, Lp, self.dim))
        ws_dec = rt.dec_ws(n, lt, ll, save, dev)
        loss = torch.empty(1, dtype=torch.float32, device=dev)
        pred_img = torch.empty_like(imgs)
        mask_img = torch.empty_like(imgs)
        P = T * Lp
        tokens = torch.empty(n, P, self.patch_embed.b_patch_size * self.patch_embed.patch_size[0] ** 2,
                             dtype=torch.float32, device=dev) if want_tokens else None
        _lib.check(rt.lib.hsimae_decoder_forward(rt.plan, _ptr(rt.wb), _ptr(rt.wf), _ptr(imgs), n, lt, ll, _ptr(restore32),
                                                 _ptr(mask), _ptr(ws_enc), int(save), int(save), _ptr(ws_dec), ws_dec.numel(),
                                                 _ptr(loss), _ptr(pred_img), _ptr(mask_img), _ptr(tokens), _stream()),
                   "decoder_forward")
        pe = self.patch_embed
        self.patch_info = (n, pe.bands, pe.img_size[0], pe.img_size[1], pe.patch_size[0], pe.b_patch_size, T, G, G)
        self.__dict__["_last_imgs"], self.__dict__["_last_stats"] = imgs, None
        aux = dict(ids_keep=ids_keep, ids_restore=ids_restore, mask=mask, keep32=keep32, restore32=restore32, lt=lt, ll=ll,
                   drops=drops, tokens=tokens, noise_t=noise_t, noise_l=noise_l)
        return loss.view(()), pred_img, mask_img, aux, ws_enc, ws_dec

    def _attach(self, rt, params, saved: _Saved, loss, logits):
        saved.rt = rt
        saved.dp = self.__dict__.get("_dp")
        outs = _HsiFunction.apply(saved, loss, logits, *params)
        outs = list(outs)
        loss_o = outs.pop(0) if loss is not None else None
        logits_o = outs.pop(0) if logits is not None else None
        return loss_o, logits_o


class _PatchLayout:
    """The reference's public layout helpers of the reconstruction models (pure index permutations, any device)."""

    def patchify(self, imgs):
        """[N,1,bands,H,W] -> [N, T*L, u*p*p]: token order (t, h, w), element order (u, p, q)  (Models.py:461-473)"""
        N, _, T, H, W = imgs.shape
        p, u = self.patch_embed.patch_size[0], self.b_pred_patch_size
        assert H == W and H % p == 0 and T % u == 0
        h = w = H // p
        t = T // u
        x = imgs.reshape(N, t, u, h, p, w, p).permute(0, 1, 3, 5, 2, 4, 6).reshape(N, t * h * w, u * p ** 2)
        self.patch_info = (N, T, H, W, p, u, t, h, w)
        return x

    def unpatchify(self, x):
        """inverse of :meth:`patchify` for the shape recorded by the last patchify / forward call  (Models.py:475-482)"""
        N, T, H, W, p, u, t, h, w = self.patch_info
        return x.reshape(N, t, h, w, u, p, p).permute(0, 1, 4, 2, 5, 3, 6).reshape(N, 1, T, H, W)

    def get_dim_patches(self, T, L, mask_ratio):
        """(len_t, len_l) as 0-dim LongTensors; consumes one `random.sample` draw  (Models.py:484-493)"""
        lt, ll = choose_visible_shape(T, L, mask_ratio)
        return torch.tensor(lt), torch.tensor(ll)

    # `self.mean` / `self.var` of the reference's forward_loss (Models.py:605-610, 951-955): the per-patch mean and the
    # per-patch STANDARD DEVIATION sqrt(var_unbiased + 1e-6) of the last reconstructed batch, [N, T*L, 1] each.  The fused
    # loss kernel keeps them in registers; these attributes are evaluated on first access from the batch of the last
    # forward call (a reference to it is kept, no copy), so the training loop pays nothing for them.
    def _patch_stats(self):
        imgs = self.__dict__.get("_last_imgs")
        if imgs is None or not self.norm_pix_loss:
            raise AttributeError("mean / var are set by a forward call with norm_pix_loss=True (Models.py:605-610)")
        cache = self.__dict__.get("_last_stats")
        if cache is None:
            info = self.__dict__.get("patch_info")
            target = self.patchify(imgs)
            if info is not None:
                self.patch_info = info
            mean = target.mean(dim=-1, keepdim=True)
            std = (target.var(dim=-1, keepdim=True) + 1.0e-6) ** 0.5
            cache = self.__dict__["_last_stats"] = (mean, std)
        return cache

    @property
    def mean(self):
        return self._patch_stats()[0]

    @property
    def var(self):
        return self._patch_stats()[1]


def _new_saved() -> _Saved:
    s = _Saved()
    for k in _Saved.__slots__:
        setattr(s, k, None)
    return s


# ---------------------------------------------------------------------------
# the three public model classes
# ---------------------------------------------------------------------------
class HSIMAE(_PatchLayout, _HsiBase):
    """Masked autoencoder with separate spatial / spectral encoders (Models.py:309-634)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=1024, depth=24, num_heads=16,
                 decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16, mlp_ratio=4.0, norm_layer=nn.LayerNorm,
                 norm_pix_loss=False, bands=16, b_patch_size=4, no_qkv_bias=False, trunc_init=False, s_depth=6, **kwargs):
        super().__init__()
        self.dim, self.dec_dim, self.s_depth = embed_dim, decoder_embed_dim, s_depth
        self.b_pred_patch_size = b_patch_size
        self.trunc_init, self.norm_pix_loss = trunc_init, norm_pix_loss
        self._depth, self._num_heads, self._dec_heads = depth, num_heads, decoder_num_heads
        self._mlp_ratio, self._qkv_bias = mlp_ratio, not no_qkv_bias
        self._build_encoder(img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_ratio, norm_layer, bands,
                            b_patch_size, no_qkv_bias, s_depth, [0.0] * max(depth, s_depth))
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.decoder_embed = nn.Linear(embed_dim, decoder_embed_dim, bias=True)
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches, decoder_embed_dim))
        self._build_decoder(embed_dim, decoder_embed_dim, decoder_depth, decoder_num_heads, mlp_ratio, norm_layer,
                            no_qkv_bias, patch_size, in_chans)
        self._init_all(decoder=True)

    def forward(self, imgs, mask_ratio=0.75):
        """-> (loss, pred [N,1,bands,H,W] de-normalised pixels, mask [N,1,bands,H,W])  (Models.py:627-634)"""
        x = self._check_imgs(imgs)
        rt, params = self._prepare(x)
        save = self._wants_grad(params)
        loss, pred_img, mask_img, aux, ws_enc, ws_dec = self._masked_pass(rt, x, mask_ratio, save)
        self.__dict__["_last"] = aux
        if save:
            s = _new_saved()
            s.imgs_m, s.n_m, s.lt, s.ll = x, x.shape[0], aux["lt"], aux["ll"]
            s.keep32, s.restore32, s.ws_enc, s.ws_dec, s.drops_m = aux["keep32"], aux["restore32"], ws_enc, ws_dec, aux["drops"]
            loss, _ = self._attach(rt, params, s, loss, None)
        return loss, pred_img, mask_img

    def forward_encoder(self, x, mask_ratio):
        """-> (latent [N,K,D], mask [N,T*L], ids_restore, ids_keep)  (Models.py:537-571); inference only."""
        x = self._check_imgs(x)
        rt, _ = self._prepare(x)
        n = x.shape[0]
        T, G = self.input_size[0], self.input_size[1]
        lt, ll = choose_visible_shape(T, G * G, mask_ratio)
        noise_t = torch.rand(n, T, device=x.device)
        noise_l = torch.rand(n, G * G, device=x.device)
        ids_keep, ids_restore, mask, keep32, _ = rt.mask(noise_t, noise_l, lt, ll)
        self.len_t, self.len_l = torch.tensor(lt), torch.tensor(ll)
        ws = rt.enc_ws(n, lt, ll, False, x.device)
        rt.encoder_forward(x, n, lt, ll, keep32, None, False, ws)
        return rt.encoder_latent(n, lt, ll, False, ws), mask, ids_restore, ids_keep


class DualViT(_PatchLayout, _HsiBase):
    """Dual-branch fine-tuning model: classification + reconstruction (Models.py:637-993)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=1024, depth=24, s_depth=6, num_heads=16,
                 mlp_ratio=4.0, norm_layer=nn.LayerNorm, bands=32, b_patch_size=8, num_class=100, no_qkv_bias=False,
                 trunc_init=False, drop_path=0., decoder_embed_dim=512, decoder_depth=8, decoder_num_heads=16,
                 norm_pix_loss=False, **kwargs):
        super().__init__()
        self.trunc_init, self.dim, self.dec_dim, self.s_depth = trunc_init, embed_dim, decoder_embed_dim, s_depth
        self.b_pred_patch_size, self.norm_pix_loss = b_patch_size, norm_pix_loss
        self._depth, self._num_heads, self._dec_heads = depth, num_heads, decoder_num_heads
        self._mlp_ratio, self._qkv_bias = mlp_ratio, not no_qkv_bias
        dpr = [x.item() for x in torch.linspace(0, drop_path, depth)]
        self._build_encoder(img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_ratio, norm_layer, bands,
                            b_patch_size, no_qkv_bias, s_depth, dpr)
        self.cls_head = nn.Linear(embed_dim * self.patch_embed.b_grid_size, num_class)
        self.decoder_embed = nn.Linear(embed_dim, decoder_embed_dim)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, decoder_embed_dim))
        self.decoder_pos_embed = nn.Parameter(torch.zeros(1, self.patch_embed.num_patches, decoder_embed_dim))
        self._build_decoder(embed_dim, decoder_embed_dim, decoder_depth, decoder_num_heads, mlp_ratio, norm_layer,
                            no_qkv_bias, patch_size, in_chans)
        self._init_all(decoder=True)

    def forward(self, imgs, imgs_u=None, mask_ratio=0.75):
        """train: (loss_rec, pred_rec, mask, class_pred); imgs_u=None: class_pred  (Models.py:975-993)"""
        x = self._check_imgs(imgs)
        rt, params = self._prepare(x)
        save = self._wants_grad(params)
        logits, pooled, ws_full, drops_full = self._full_pass(rt, x, save)
        s = _new_saved() if save else None
        if save:
            s.imgs_full, s.n_full, s.ws_full, s.drops_full, s.pooled = x, x.shape[0], ws_full, drops_full, pooled
        if imgs_u is None:
            if save:
                _, logits = self._attach(rt, params, s, None, logits)
            return logits
        both = torch.cat([x, self._check_imgs(imgs_u)], dim=0)
        loss, pred_img, mask_img, aux, ws_enc, ws_dec = self._masked_pass(rt, both, mask_ratio, save)
        aux["drops_full"] = drops_full
        self.__dict__["_last"] = aux
        if save:
            s.imgs_m, s.n_m, s.lt, s.ll = both, both.shape[0], aux["lt"], aux["ll"]
            s.keep32, s.restore32, s.ws_enc, s.ws_dec, s.drops_m = aux["keep32"], aux["restore32"], ws_enc, ws_dec, aux["drops"]
            loss, logits = self._attach(rt, params, s, loss, logits)
        return loss, pred_img, mask_img, logits


class HSIViT(_HsiBase):
    """Encoder-only classifier (Models.py:996-1160)."""

    def __init__(self, img_size=224, patch_size=16, in_chans=3, embed_dim=1024, depth=24, num_heads=16, mlp_ratio=4.0,
                 norm_layer=nn.LayerNorm, bands=16, b_patch_size=4, num_class=100, no_qkv_bias=False, trunc_init=False,
                 drop_rate=0., drop_path=0., s_depth=6, **kwargs):
        super().__init__()
        self.trunc_init, self.b_pred_patch_size = trunc_init, b_patch_size
        self._depth, self._num_heads, self._dec_heads = depth, num_heads, 0
        self._mlp_ratio, self._qkv_bias = mlp_ratio, not no_qkv_bias
        self.dim = embed_dim
        dpr = [x.item() for x in torch.linspace(0, drop_path, depth)]
        self._build_encoder(img_size, patch_size, in_chans, embed_dim, depth, num_heads, mlp_ratio, norm_layer, bands,
                            b_patch_size, no_qkv_bias, s_depth, dpr)
        self.cls_head = nn.Linear(embed_dim * self.patch_embed.b_grid_size, num_class)
        self.s_depth = s_depth
        self._init_all(decoder=False)

    def forward(self, imgs):
        """-> logits [N, num_class]  (Models.py:1158-1160)"""
        x = self._check_imgs(imgs)
        rt, params = self._prepare(x)
        save = self._wants_grad(params)
        logits, pooled, ws_full, drops_full = self._full_pass(rt, x, save)
        if save:
            s = _new_saved()
            s.imgs_full, s.n_full, s.ws_full, s.drops_full, s.pooled = x, x.shape[0], ws_full, drops_full, pooled
            _, logits = self._attach(rt, params, s, None, logits)
        return logits
