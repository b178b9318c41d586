// Plan (static layouts), parameter packing and the forward/backward orchestration
// of the HSIMAE compute path, exported through the C ABI of include/hsimae_b200.h.
//
// Reference call stack being replaced (all in /root/reference/Models.py):
//   HSIMAE.forward :627-634  -> forward_encoder :537-571 -> Block.forward :303-306
//                            -> forward_decoder :573-601 -> forward_loss :603-616 -> recons :618-625
//   DualViT.forward :975-993 -> forward_encoder :869-894, head :964-973, forward_mask_encoder :896-923
//   HSIViT.forward  :1158-1160
// and the autograd backward PyTorch derives from them (Model_Pretraining.py:101).
#include <string>
#include <vector>
#include <cstdlib>
#include <mutex>
#include "../../include/hsimae_b200.h"
#include "gemm.cuh"
#include "block_fused.cuh"
#include "kernels.cuh"

using namespace hsimae;
typedef __nv_bfloat16 bf16;

namespace {

struct ParamSlot {
  std::string name;
  int64_t numel;
  int64_t grad_off;  // -1: no gradient
};

struct BlockW {
  // bf16 arena (element offsets)
  int64_t wqkv, wqkv_t, wproj, wproj_t, w13, w13_t, w2, w2_t;
  // fp32 arena
  int64_t bqkv, bproj, b13, b2, g1, be1, g2, be2;
  // parameter slots (for gradients)
  int n1w, n1b, qw, qb, pw, pb, n2w, n2b, w1w, w1b, w3w, w3b, w2w, w2b;
};

struct JobT {  // packing job template; src filled from the parameter table
  int slot;
  int64_t dst_off;
  int rows, cols, pitch, kind, row_map, row_off;
};

}  // namespace

struct hsimae_plan {
  hsimae_dims dims;
  PatchGeom g;
  int D, H, Hp, heads, Dd, Hd, Hdp, dheads, n_fusion, PKp;
  std::vector<ParamSlot> slots;
  std::vector<JobT> jobs;
  std::vector<BlockW> b1, b2, bf, bd;
  int64_t f_pe_w, f_pe_b, f_pos, f_norm_g, f_norm_b, f_cls_w, f_cls_b, f_de_b, f_dpos, f_dnorm_g, f_dnorm_b, f_pred_b;
  int64_t w_de, w_de_t, w_pred, w_pred_t;
  int p_pe_w, p_pe_b, p_norm_w, p_norm_b, p_cls_w, p_cls_b, p_de_w, p_de_b, p_dnorm_w, p_dnorm_b, p_pred_w, p_pred_b;
  int64_t grad_elems, bf16_elems, f32_elems;
  int64_t bucket_end[6];  // gradient-arena boundaries (see hsimae_plan_grad_bucket): the spatial encoder is cut into three
  int sp_cut[2];          // first spatial block of the middle / last third (backward sub-stages 8 / 4)
  int max_job_elems;
  std::vector<const void*> last_ptrs;
  const void* last_table;
  bool debug_simt;
  bool helper_pending = false;   // the spectral backward chain is running on the helper stream and has not been joined yet
  bool recompute_gate;   // backward recomputes the gated up-projection instead of re-reading saved pre-activations
};

namespace {

int swiglu_hidden(int dim, float ratio) {
  // Models.py:225 with Block's arguments (Models.py:300-301): hidden=int(dim*ratio), multiple_of=ratio
  const int hidden = (int)(dim * ratio);
  const double m = ratio;
  const long two_thirds = (2L * hidden) / 3;
  const double q = floor((two_thirds + m - 1) / m);
  return (int)(m * q);
}

struct Builder {
  hsimae_plan* p;
  int64_t wb = 0, wf = 0, gr = 0;
  int add_slot(const std::string& name, int64_t numel, bool grad) {
    ParamSlot s{name, numel, -1};
    if (grad) { gr = align_up(gr, 4); s.grad_off = gr; gr += numel; }
    p->slots.push_back(s);
    return (int)p->slots.size() - 1;
  }
  int64_t take_b(int64_t n) { wb = align_up(wb, 64); int64_t o = wb; wb += n; return o; }
  int64_t take_f(int64_t n) { wf = align_up(wf, 16); int64_t o = wf; wf += n; return o; }
  void job(int slot, int64_t dst, int rows, int cols, int pitch, int kind, int row_map = 0, int row_off = 0) {
    p->jobs.push_back(JobT{slot, dst, rows, cols, pitch, kind, row_map, row_off});
    if (rows * cols > p->max_job_elems) p->max_job_elems = rows * cols;
  }
  void vec(int slot, int64_t dst, int n, int row_map = 0, int row_off = 0) {
    // plain vectors are packed as one row (64 elements per tile column block instead of one element per tile row)
    if (row_map == 0) job(slot, dst + row_off, 1, n, n, 0, 0, 0);
    else job(slot, dst, n, 1, 1, 0, row_map, row_off);
  }

  BlockW block(const std::string& pre, int d, int H, int Hp, bool qkv_bias) {
    BlockW w{};
    w.n1w = add_slot(pre + "norm1.weight", d, true);
    w.n1b = add_slot(pre + "norm1.bias", d, true);
    w.qw = add_slot(pre + "attn.q.weight", (int64_t)d * d, true);
    int kw = add_slot(pre + "attn.k.weight", (int64_t)d * d, true);
    int vw = add_slot(pre + "attn.v.weight", (int64_t)d * d, true);
    int kb = -1, vb = -1;
    w.qb = -1;
    if (qkv_bias) {
      w.qb = add_slot(pre + "attn.q.bias", d, true);
      kb = add_slot(pre + "attn.k.bias", d, true);
      vb = add_slot(pre + "attn.v.bias", d, true);
    }
    w.pw = add_slot(pre + "attn.proj.weight", (int64_t)d * d, true);
    w.pb = add_slot(pre + "attn.proj.bias", d, true);
    w.n2w = add_slot(pre + "norm2.weight", d, true);
    w.n2b = add_slot(pre + "norm2.bias", d, true);
    w.w1w = add_slot(pre + "mlp.w1.weight", (int64_t)H * d, true);
    w.w1b = add_slot(pre + "mlp.w1.bias", H, true);
    w.w3w = add_slot(pre + "mlp.w3.weight", (int64_t)H * d, true);
    w.w3b = add_slot(pre + "mlp.w3.bias", H, true);
    w.w2w = add_slot(pre + "mlp.w2.weight", (int64_t)d * H, true);
    w.w2b = add_slot(pre + "mlp.w2.bias", d, true);

    w.wqkv = take_b((int64_t)3 * d * d);   w.wqkv_t = take_b((int64_t)3 * d * d);
    w.wproj = take_b((int64_t)d * d);      w.wproj_t = take_b((int64_t)d * d);
    w.w13 = take_b((int64_t)2 * Hp * d);   w.w13_t = take_b((int64_t)2 * Hp * d);
    w.w2 = take_b((int64_t)d * Hp);        w.w2_t = take_b((int64_t)Hp * d);
    w.bqkv = take_f(3 * d); w.bproj = take_f(d); w.b13 = take_f(2 * Hp); w.b2 = take_f(d);
    w.g1 = take_f(d); w.be1 = take_f(d); w.g2 = take_f(d); w.be2 = take_f(d);

    vec(w.n1w, w.g1, d); vec(w.n1b, w.be1, d); vec(w.n2w, w.g2, d); vec(w.n2b, w.be2, d);
    const int qs[3] = {w.qw, kw, vw};
    const int bs[3] = {w.qb, kb, vb};
    for (int i = 0; i < 3; ++i) {
      job(qs[i], w.wqkv, d, d, d, 1, 0, i * d);
      job(qs[i], w.wqkv_t, d, d, 3 * d, 2, 0, i * d);
      if (qkv_bias) vec(bs[i], w.bqkv, d, 0, i * d);
    }
    job(w.pw, w.wproj, d, d, d, 1); job(w.pw, w.wproj_t, d, d, d, 2);
    vec(w.pb, w.bproj, d);
    job(w.w1w, w.w13, H, d, d, 1, 1); job(w.w3w, w.w13, H, d, d, 1, 2);
    job(w.w1w, w.w13_t, H, d, 2 * Hp, 2, 1); job(w.w3w, w.w13_t, H, d, 2 * Hp, 2, 2);
    vec(w.w1b, w.b13, H, 1); vec(w.w3b, w.b13, H, 2);
    job(w.w2w, w.w2, d, H, Hp, 1); job(w.w2w, w.w2_t, d, H, d, 2);
    vec(w.w2b, w.b2, d);
    return w;
  }
};

int build_plan(const hsimae_dims& dm, hsimae_plan* p) {
  HS_REQUIRE(dm.img_size > 0 && dm.patch_size > 0 && dm.img_size % dm.patch_size == 0, "img_size %d not divisible by patch_size %d", dm.img_size, dm.patch_size);
  HS_REQUIRE(dm.bands > 0 && dm.b_patch_size > 0 && dm.bands % dm.b_patch_size == 0, "bands %d not divisible by b_patch_size %d", dm.bands, dm.b_patch_size);
  HS_REQUIRE(dm.embed_dim % 16 == 0 && dm.embed_dim >= 16 && dm.embed_dim <= 256, "embed_dim %d unsupported (multiple of 16, <= 256)", dm.embed_dim);
  HS_REQUIRE(dm.num_heads > 0 && dm.embed_dim % dm.num_heads == 0, "embed_dim %d not divisible by num_heads %d", dm.embed_dim, dm.num_heads);
  HS_REQUIRE(dm.depth >= 0 && dm.s_depth >= 0, "negative depth");
  if (dm.dec_dim > 0) {
    HS_REQUIRE(dm.dec_dim % 16 == 0 && dm.dec_dim <= 256, "decoder_embed_dim %d unsupported (multiple of 16, <= 256)", dm.dec_dim);
    HS_REQUIRE(dm.dec_heads > 0 && dm.dec_dim % dm.dec_heads == 0, "decoder dim %d not divisible by heads %d", dm.dec_dim, dm.dec_heads);
  }
  p->dims = dm;
  PatchGeom& g = p->g;
  g.bands = dm.bands; g.img = dm.img_size; g.u = dm.b_patch_size; g.p = dm.patch_size;
  g.T = dm.bands / dm.b_patch_size; g.G = dm.img_size / dm.patch_size; g.L = g.G * g.G; g.P = g.T * g.L;
  g.PK = g.u * g.p * g.p; g.cube = g.bands * g.img * g.img;
  HS_REQUIRE(g.T <= 32 && g.L <= 64 && g.P <= 256, "token grid %dx%d unsupported", g.T, g.L);
  HS_REQUIRE(g.PK % 4 == 0 && g.PK <= 128, "patch of %d elements unsupported (multiple of 4, <= 128)", g.PK);
  HS_REQUIRE(g.cube % 4 == 0, "cube of %d elements unsupported (multiple of 4)", g.cube);
  p->D = dm.embed_dim; p->heads = dm.num_heads;
  p->H = swiglu_hidden(p->D, dm.mlp_ratio); p->Hp = (int)align_up(p->H, 16);
  p->Dd = dm.dec_dim; p->dheads = dm.dec_heads;
  p->Hd = p->Dd ? swiglu_hidden(p->Dd, dm.mlp_ratio) : 0; p->Hdp = (int)align_up(p->Hd, 16);
  HS_REQUIRE(p->H % 4 == 0 && (p->Dd == 0 || p->Hd % 4 == 0), "SwiGLU hidden width must be a multiple of 4 (mlp_ratio %.3f)", dm.mlp_ratio);
  // Models.py:356,385 -- split encoders iff s_depth > 0, fusion blocks iff s_depth < 12 (literal)
  p->n_fusion = (dm.s_depth < 12 && dm.depth > dm.s_depth) ? dm.depth - dm.s_depth : 0;
  p->PKp = (int)align_up(g.PK, 16);
  p->max_job_elems = 1;
  p->last_table = nullptr;
  p->debug_simt = getenv("HSIMAE_DEBUG_SIMT") && atoi(getenv("HSIMAE_DEBUG_SIMT")) != 0;
  // HSIMAE_SAVE_GATE=1 keeps the pre-activations (A/B measurements); the CUDA-core checker has no recompute kernel
  p->recompute_gate = !p->debug_simt && !(getenv("HSIMAE_SAVE_GATE") && atoi(getenv("HSIMAE_SAVE_GATE")) != 0) && p->D <= 256 && p->Dd <= 256;

  Builder b{p};
  const int D = p->D, Dd = p->Dd;
  p->p_pe_w = b.add_slot("patch_embed.proj.weight", (int64_t)D * g.PK, true);
  p->p_pe_b = b.add_slot("patch_embed.proj.bias", D, true);
  int p_pos = b.add_slot("pos_embed", (int64_t)g.P * D, false);
  p->f_pe_w = b.take_f((int64_t)D * g.PK); p->f_pe_b = b.take_f(D); p->f_pos = b.take_f((int64_t)g.P * D);
  b.vec(p->p_pe_w, p->f_pe_w, D * g.PK); b.vec(p->p_pe_b, p->f_pe_b, D); b.vec(p_pos, p->f_pos, g.P * D);
  const bool qb = dm.qkv_bias != 0;
  // the spatial encoder is the LAST thing backward finishes: its gradients go out in three pieces so that only the
  // all-reduce of the final third of the blocks (+ patch embedding) cannot hide behind compute
  p->sp_cut[0] = dm.s_depth / 3; p->sp_cut[1] = 2 * dm.s_depth / 3;
  for (int i = 0; i < dm.s_depth; ++i) {
    if (i == p->sp_cut[0]) p->bucket_end[0] = align_up(b.gr, 4);   // patch embedding + spatial blocks [0, cut0)
    if (i == p->sp_cut[1]) p->bucket_end[1] = align_up(b.gr, 4);   // spatial blocks [cut0, cut1)
    p->b1.push_back(b.block("blocks_1." + std::to_string(i) + ".", D, p->H, p->Hp, qb));
  }
  if (dm.s_depth <= p->sp_cut[0]) p->bucket_end[0] = align_up(b.gr, 4);
  if (dm.s_depth <= p->sp_cut[1]) p->bucket_end[1] = align_up(b.gr, 4);
  p->bucket_end[2] = align_up(b.gr, 4);   // spatial blocks [cut1, s_depth)
  for (int i = 0; i < dm.s_depth; ++i) p->b2.push_back(b.block("blocks_2." + std::to_string(i) + ".", D, p->H, p->Hp, qb));
  p->bucket_end[3] = align_up(b.gr, 4);   // spectral encoder
  for (int i = 0; i < p->n_fusion; ++i) p->bf.push_back(b.block("blocks." + std::to_string(i) + ".", D, p->H, p->Hp, qb));
  p->p_norm_w = b.add_slot("norm.weight", D, true);
  p->p_norm_b = b.add_slot("norm.bias", D, true);
  p->f_norm_g = b.take_f(D); p->f_norm_b = b.take_f(D);
  b.vec(p->p_norm_w, p->f_norm_g, D); b.vec(p->p_norm_b, p->f_norm_b, D);
  p->p_cls_w = p->p_cls_b = -1;
  if (dm.num_class > 0) {
    p->p_cls_w = b.add_slot("cls_head.weight", (int64_t)dm.num_class * g.T * D, true);
    p->p_cls_b = b.add_slot("cls_head.bias", dm.num_class, true);
    p->f_cls_w = b.take_f((int64_t)dm.num_class * g.T * D); p->f_cls_b = b.take_f(dm.num_class);
    b.vec(p->p_cls_w, p->f_cls_w, dm.num_class * g.T * D); b.vec(p->p_cls_b, p->f_cls_b, dm.num_class);
  }
  p->bucket_end[4] = align_up(b.gr, 4);   // fusion blocks + final norm + classification head
  p->p_de_w = -1;
  if (Dd > 0) {
    p->p_de_w = b.add_slot("decoder_embed.weight", (int64_t)Dd * D, true);
    p->p_de_b = b.add_slot("decoder_embed.bias", Dd, true);
    int p_dpos = b.add_slot("decoder_pos_embed", (int64_t)g.P * Dd, false);
    p->w_de = b.take_b((int64_t)Dd * D); p->w_de_t = b.take_b((int64_t)D * Dd);
    p->f_de_b = b.take_f(Dd); p->f_dpos = b.take_f((int64_t)g.P * Dd);
    b.job(p->p_de_w, p->w_de, Dd, D, D, 1); b.job(p->p_de_w, p->w_de_t, Dd, D, Dd, 2);
    b.vec(p->p_de_b, p->f_de_b, Dd); b.vec(p_dpos, p->f_dpos, g.P * Dd);
    for (int i = 0; i < dm.dec_depth; ++i)
      p->bd.push_back(b.block("decoder_blocks." + std::to_string(i) + ".", Dd, p->Hd, p->Hdp, qb));
    p->p_dnorm_w = b.add_slot("decoder_norm.weight", Dd, true);
    p->p_dnorm_b = b.add_slot("decoder_norm.bias", Dd, true);
    p->p_pred_w = b.add_slot("decoder_pred.weight", (int64_t)g.PK * Dd, true);
    p->p_pred_b = b.add_slot("decoder_pred.bias", g.PK, true);
    p->f_dnorm_g = b.take_f(Dd); p->f_dnorm_b = b.take_f(Dd); p->f_pred_b = b.take_f(p->PKp);
    p->w_pred = b.take_b((int64_t)p->PKp * Dd); p->w_pred_t = b.take_b((int64_t)Dd * p->PKp);
    b.vec(p->p_dnorm_w, p->f_dnorm_g, Dd); b.vec(p->p_dnorm_b, p->f_dnorm_b, Dd); b.vec(p->p_pred_b, p->f_pred_b, g.PK);
    b.job(p->p_pred_w, p->w_pred, g.PK, Dd, Dd, 1); b.job(p->p_pred_w, p->w_pred_t, g.PK, Dd, p->PKp, 2);
  }
  p->grad_elems = align_up(b.gr, 4);
  p->bucket_end[5] = p->grad_elems;       // decoder
  p->bf16_elems = align_up(b.wb, 64);
  p->f32_elems = align_up(b.wf, 16);
  return kOk;
}

// ---------------------------------------------------------------------------
// workspace layouts
// ---------------------------------------------------------------------------
struct Bump {
  uintptr_t base;
  int64_t off = 0;
  template <class T> T* take(int64_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + (uintptr_t)off);
    off += count * (int64_t)sizeof(T);
    return r;
  }
};

struct BlockStash {
  bf16* ln1; float* stats1; bf16* qkv; bf16* ao; float* lse;
  float* x_mid; float* stats2; bf16* ln2; bf16* ab; bf16* g; float* x_out;
};

BlockStash take_stash(Bump& b, int64_t M, int d, int Hp, int heads) {
  BlockStash s;
  s.ln1 = b.take<bf16>(M * d); s.stats1 = b.take<float>(M * 2);
  s.qkv = b.take<bf16>(M * 3 * d); s.ao = b.take<bf16>(M * d); s.lse = b.take<float>(M * heads);
  s.x_mid = b.take<float>(M * d); s.stats2 = b.take<float>(M * 2); s.ln2 = b.take<bf16>(M * d);
  s.ab = b.take<bf16>(M * 2 * Hp); s.g = b.take<bf16>(M * Hp); s.x_out = b.take<float>(M * d);
  return s;
}

void take_chain(Bump& b, std::vector<BlockStash>& v, int nblk, bool save, int64_t M, int d, int Hp, int heads) {
  v.clear();
  if (nblk == 0) return;
  if (save) { for (int i = 0; i < nblk; ++i) v.push_back(take_stash(b, M, d, Hp, heads)); }
  else { BlockStash s = take_stash(b, M, d, Hp, heads); for (int i = 0; i < nblk; ++i) v.push_back(s); }
}

struct BwdScratch {
  float* dxA; float* dxB; bf16* dxb; bf16* dab; bf16* dln; bf16* dqkv; bf16* dao;
  bf16* dxb2;   // bf16(rs1 * dx_mid): kept apart from dxb so that the block's four weight gradients can run as one launch
};

BwdScratch take_bwd(Bump& b, int64_t M, int d, int Hp) {
  BwdScratch s;
  s.dxA = b.take<float>(M * d); s.dxB = b.take<float>(M * d); s.dxb = b.take<bf16>(M * d);
  s.dab = b.take<bf16>(M * 2 * Hp); s.dln = b.take<bf16>(M * d); s.dqkv = b.take<bf16>(M * 3 * d); s.dao = b.take<bf16>(M * d);
  s.dxb2 = b.take<bf16>(M * d);
  return s;
}

struct EncLayout {
  int N, K, lt, ll; int64_t M;
  float* x0;
  std::vector<BlockStash> sp, sc, fu;
  bf16* latent; float* stats_f; float* x_final;
  bf16* dlatent; BwdScratch bw;
  BwdScratch bw2; float* dxC;   // split encoders: scratch of the spectral chain when it runs beside the spatial one; spatial gradient stream
  float* simt;
  int64_t bytes;
};

EncLayout enc_layout(const hsimae_plan* p, int N, int lt, int ll, bool save, const void* ws) {
  EncLayout L;
  L.N = N; L.lt = lt; L.ll = ll; L.K = lt * ll; L.M = (int64_t)N * L.K;
  Bump b{reinterpret_cast<uintptr_t>(ws)};
  const int D = p->D;
  L.x0 = b.take<float>(L.M * D);
  take_chain(b, L.sp, p->dims.s_depth, save, L.M, D, p->Hp, p->heads);
  take_chain(b, L.sc, p->dims.s_depth, save, L.M, D, p->Hp, p->heads);
  take_chain(b, L.fu, p->n_fusion, save, L.M, D, p->Hp, p->heads);
  L.latent = b.take<bf16>(L.M * D);
  L.stats_f = b.take<float>(L.M * 2);
  L.x_final = !L.fu.empty() ? L.fu.back().x_out : (!L.sc.empty() ? L.sc.back().x_out : L.x0);
  L.dlatent = nullptr; L.bw = BwdScratch{};
  L.bw2 = BwdScratch{}; L.dxC = nullptr;
  if (save) {
    L.dlatent = b.take<bf16>(L.M * D); L.bw = take_bwd(b, L.M, D, p->Hp);
    if (p->dims.s_depth > 0) { L.bw2 = take_bwd(b, L.M, D, p->Hp); L.dxC = b.take<float>(L.M * D); }
  }
  L.simt = p->debug_simt ? b.take<float>(L.M * (int64_t)(3 * D > 2 * p->Hp ? 3 * D : 2 * p->Hp)) : nullptr;
  L.bytes = align_up(b.off, 256);
  return L;
}

struct DecLayout {
  int N, K, P; int64_t M, Md;
  float* y; float* x0; bf16* ln0; float* stats0;
  std::vector<BlockStash> blk;
  bf16* dnorm; float* stats_n; float* x_final;
  float* pred; bf16* dpred_unit; float* loss_partial;
  bf16* dpred; bf16* dy; BwdScratch bw;
  float* simt;
  int64_t bytes;
};

DecLayout dec_layout(const hsimae_plan* p, int N, int lt, int ll, bool save, const void* ws) {
  DecLayout L;
  L.N = N; L.K = lt * ll; L.P = p->g.P; L.M = (int64_t)N * L.K; L.Md = (int64_t)N * L.P;
  Bump b{reinterpret_cast<uintptr_t>(ws)};
  const int Dd = p->Dd;
  L.y = b.take<float>(L.M * Dd);
  L.x0 = b.take<float>(L.Md * Dd);
  take_chain(b, L.blk, p->dims.dec_depth, save, L.Md, Dd, p->Hdp, p->dheads);
  L.dnorm = b.take<bf16>(L.Md * Dd);
  L.stats_n = b.take<float>(L.Md * 2);
  L.x_final = !L.blk.empty() ? L.blk.back().x_out : L.x0;
  L.pred = b.take<float>(L.Md * p->PKp);
  L.dpred_unit = b.take<bf16>(L.Md * p->PKp);
  L.loss_partial = b.take<float>(N);
  L.dpred = nullptr; L.dy = nullptr; L.bw = BwdScratch{};
  if (save) { L.dpred = b.take<bf16>(L.Md * p->PKp); L.dy = b.take<bf16>(L.M * Dd); L.bw = take_bwd(b, L.Md, Dd, p->Hdp); }
  L.simt = p->debug_simt ? b.take<float>(L.Md * (int64_t)(3 * Dd > 2 * p->Hdp ? 3 * Dd : 2 * p->Hdp) + L.Md * p->PKp) : nullptr;
  L.bytes = align_up(b.off, 256);
  return L;
}

// ---------------------------------------------------------------------------
// block forward / backward
// ---------------------------------------------------------------------------
struct Ctx {
  const hsimae_plan* p;
  const bf16* wb;
  const float* wf;
  float* grads;
  float* simt;
  cudaStream_t st;
  bool save = true;   // the pass keeps what backward needs
};

// HSIMAE_FUSED_MLP=0 restores the two-launch gated MLP (A/B measurements)
bool fused_mlp_enabled() {
  static const bool on = !(getenv("HSIMAE_FUSED_MLP") && atoi(getenv("HSIMAE_FUSED_MLP")) == 0);
  return on;
}

// The recompute-in-backward d(gate) kernel and the fused MLP kernel walk ALL hidden chunks of a row tile inside one CTA:
// serial latency when there are only a few row tiles (fine-tuning batches: 1152 rows = 9 tiles, d(gate) 29 us).  Keeping
// the pre-activations and using the per-output-tile kernels for such problems was measured and is NOT faster end to end
// (fine-tuning step 9.9 -> 10.3 ms: one more launch and the [M, 2Hp] write in forward): HSIMAE_RECOMPUTE_MIN_ROWS (default 0)
// keeps the switch for A/B measurements.  The choice depends on M only, so forward and backward agree.
bool recompute_gate_for(const hsimae_plan* p, int64_t M) {
  static const int min_rows = getenv("HSIMAE_RECOMPUTE_MIN_ROWS") ? atoi(getenv("HSIMAE_RECOMPUTE_MIN_ROWS")) : 0;
  return p->recompute_gate && M >= min_rows;
}

// ---------------------------------------------------------------------------
// The spatial and the spectral encoder are independent chains of full-device kernels between the patch embedding and
// the fusion blocks (Models.py:553-564).  Run back to back, every kernel boundary (~2.5 us of drain / launch / fill) and
// every partial last wave leaves SMs idle; run on TWO streams, the CTAs of one chain's next kernel fill the SMs the other
// chain's kernel is draining.  The helper stream is forked from and joined back into the caller's stream with events, so
// the ABI contract (all work ordered on the given stream, capturable) holds.  HSIMAE_OVERLAP=0 disables.
// ---------------------------------------------------------------------------
struct Helper { cudaStream_t st = nullptr; cudaEvent_t fork = nullptr, join = nullptr, aux = nullptr; bool ok = false; };
Helper& helper() {
  static Helper h;
  static std::once_flag once;
  std::call_once(once, [] {
    const bool on = !(getenv("HSIMAE_OVERLAP") && atoi(getenv("HSIMAE_OVERLAP")) == 0);
    if (!on) return;
    if (cudaStreamCreateWithFlags(&h.st, cudaStreamNonBlocking) != cudaSuccess) return;
    if (cudaEventCreateWithFlags(&h.fork, cudaEventDisableTiming) != cudaSuccess) return;
    if (cudaEventCreateWithFlags(&h.join, cudaEventDisableTiming) != cudaSuccess) return;
    if (cudaEventCreateWithFlags(&h.aux, cudaEventDisableTiming) != cudaSuccess) return;
    h.ok = true;
  });
  return h;
}

int run_gemm(const Ctx& c, GemmArgs& a, int epi) {
  a.ln_eps = 1e-5f;
  if (c.p->debug_simt) return gemm_simt(a, epi, c.simt, c.st);
  return gemm_tc(a, epi, c.st);
}
int run_wgrad(const Ctx& c, const WgradArgs& a) {
  if (c.p->debug_simt) return wgrad_simt(a, c.st);
  return wgrad_tc(a, c.st);
}
// the weight gradients of one block: one grouped launch (every operand stays live until the end of the block's backward)
int run_wgrad_group(const Ctx& c, const WgradArgs* jobs, int n) {
  if (c.p->debug_simt) { for (int i = 0; i < n; ++i) HS_TRY(wgrad_simt(jobs[i], c.st)); return kOk; }
  return wgrad_tc_group(jobs, n, c.st);
}

struct TailLN {
  const float* gamma; const float* beta;  // nullptr: no LayerNorm after the block
  bf16* ln; float* stats;
  const float* resid2;                    // extra residual summed into x_out
};

float* gptr(const Ctx& c, int slot) { return slot >= 0 ? c.grads + c.p->slots[slot].grad_off : nullptr; }

int block_forward(const Ctx& c, const BlockW& w, int64_t M, int N, int d, int Hp, int heads, const SeqSpec& seq,
                  BlockStash& s, const float* x_in, RowScale rs1, RowScale rs2, const TailLN& tail) {
  GemmArgs a{};
  // q|k|v projection (Models.py:194-208)
  a.M = (int)M; a.N = 3 * d; a.K = d; a.A = s.ln1; a.lda = d; a.B = c.wb + w.wqkv; a.ldb = d;
  a.out0 = s.qkv; a.ld0 = 3 * d; a.bias = c.wf + w.bqkv;
  HS_TRY(run_gemm(c, a, kEpiBiasBf16));
  // attention (Models.py:210-215)
  AttnArgs at{}; at.N = N; at.D = d; at.heads = heads; at.s = seq; at.qkv = s.qkv; at.out = s.ao; at.lse = s.lse;
  HS_TRY(launch_attn_fwd(at, c.st));
  // output projection + residual + norm2 (Models.py:216, 304-305)
  a = GemmArgs{};
  a.M = (int)M; a.N = d; a.K = d; a.A = s.ao; a.lda = d; a.B = c.wb + w.wproj; a.ldb = d;
  a.out0 = s.x_mid; a.ld0 = d; a.bias = c.wf + w.bproj; a.resid = x_in; a.ldr = d; a.rs = rs1;
  a.gamma = c.wf + w.g2; a.beta = c.wf + w.be2; a.out1 = s.ln2; a.ld1 = d; a.stats = s.stats2;
  HS_TRY(run_gemm(c, a, kEpiResidLN));
  // gated MLP (Models.py:232): up-projection, gate, down-projection + residual (+ other branch) + next LayerNorm
  GemmArgs dn{};
  dn.M = (int)M; dn.N = d; dn.K = Hp; dn.A = s.g; dn.lda = Hp; dn.B = c.wb + w.w2; dn.ldb = Hp;
  dn.out0 = s.x_out; dn.ld0 = d; dn.bias = c.wf + w.b2; dn.resid = s.x_mid; dn.ldr = d; dn.resid2 = tail.resid2; dn.rs = rs2;
  dn.gamma = tail.gamma; dn.beta = tail.beta; dn.out1 = tail.ln; dn.ld1 = d; dn.stats = tail.stats; dn.ln_eps = 1e-5f;
  const bool recompute = recompute_gate_for(c.p, M);
  if (!c.p->debug_simt && recompute && fused_mlp_enabled() && mlp_fused_supported(d, Hp)) {
    // one kernel; the gate output is only written when backward will read it (dW2)
    MlpFusedArgs m{};
    m.tail = dn; m.X = s.ln2; m.ldx = d; m.W13 = c.wb + w.w13; m.ldw = d; m.b13 = c.wf + w.b13;
    m.g = c.save ? s.g : nullptr; m.ldg = Hp;
    HS_TRY(mlp_fused(m, c.st));
    return kOk;
  }
  a = GemmArgs{};
  a.M = (int)M; a.N = 2 * Hp; a.K = d; a.A = s.ln2; a.lda = d; a.B = c.wb + w.w13; a.ldb = d;
  // the pre-activations are only kept for the checker path; the product path recomputes them in backward
  a.out0 = recompute ? nullptr : s.ab; a.ld0 = 2 * Hp; a.out1 = s.g; a.ld1 = Hp; a.bias = c.wf + w.b13;
  HS_TRY(run_gemm(c, a, kEpiSwiGLU));
  HS_TRY(run_gemm(c, dn, kEpiResidLN));
  return kOk;
}

// dx_src: gradient w.r.t. x_out (fp32); dxb: bf16(rs2 * dx_src).  On return dx_dst holds the gradient
// w.r.t. x_in and dxb holds bf16(rs_prev * dx_dst) (when want_dxb).
int block_backward(const Ctx& c, const BlockW& w, int64_t M, int N, int d, int H, int Hp, int heads, const SeqSpec& seq,
                   const BlockStash& s, const float* x_in, RowScale rs1, RowScale rs_prev, bool want_dxb,
                   const float* dx_src, float* dx_dst, BwdScratch& b) {
  GemmArgs a{};
  // d(gate): dab = dswiglu(dxb W2, ab)
  a.M = (int)M; a.N = Hp; a.K = d; a.A = b.dxb; a.lda = d; a.B = c.wb + w.w2_t; a.ldb = d;
  a.out0 = b.dab; a.ld0 = 2 * Hp;
  if (recompute_gate_for(c.p, M)) {
    a.A2 = s.ln2; a.lda2 = d; a.B2 = c.wb + w.w13; a.ldb2 = d; a.bias = c.wf + w.b13;
    HS_TRY(run_gemm(c, a, kEpiDGate));
  } else {
    a.ab = s.ab; a.ldab = 2 * Hp;
    HS_TRY(run_gemm(c, a, kEpiDSwiGLU));
  }
  WgradArgs wj[4];
  WgradArgs& g2 = wj[0];
  g2 = WgradArgs{};
  g2.Mred = (int)M; g2.Nout = d; g2.Kin = Hp; g2.Y = b.dxb; g2.ldy = d; g2.X = s.g; g2.ldx = Hp;
  g2.dst0 = gptr(c, w.w2w); g2.ld = H; g2.rows_valid = d; g2.cols_valid = H; g2.bias0 = gptr(c, w.w2b);
  // d(ln2) = dab W13
  a = GemmArgs{};
  a.M = (int)M; a.N = d; a.K = 2 * Hp; a.A = b.dab; a.lda = 2 * Hp; a.B = c.wb + w.w13_t; a.ldb = 2 * Hp;
  // ... followed by norm2 backward + residual, emitting bf16(rs1 * dx_mid) into its own buffer (dxb is still the dW2 operand):
  // one kernel where the row fits a tile (kEpiLnBwd), else the dgrad GEMM and ln_bwd_kernel
  GemmArgs f = a;
  f.out0 = dx_dst; f.ld0 = d; f.out1 = b.dxb2; f.ld1 = d; f.resid = dx_src; f.ldr = d; f.lnx = s.x_mid; f.ldx = d;
  f.stats = s.stats2; f.gamma = c.wf + w.g2; f.rs = rs1; f.dgamma = gptr(c, w.n2w); f.dbeta = gptr(c, w.n2b);
  const bool fuse_ln = !c.p->debug_simt && dx_src != nullptr && gemm_lnbwd_preferred(f);
  if (fuse_ln) {
    HS_TRY(gemm_tc_lnbwd(f, c.st));
  } else {
    a.out0 = b.dln; a.ld0 = d;
    HS_TRY(run_gemm(c, a, kEpiBiasBf16));
  }
  WgradArgs& g13 = wj[1];
  g13 = WgradArgs{};
  g13.Mred = (int)M; g13.Nout = 2 * Hp; g13.Kin = d; g13.Y = b.dab; g13.ldy = 2 * Hp; g13.X = s.ln2; g13.ldx = d;
  g13.dst0 = gptr(c, w.w1w); g13.dst1 = gptr(c, w.w3w); g13.ld = d; g13.row_map = 1; g13.rows_valid = H; g13.cols_valid = d;
  g13.bias0 = gptr(c, w.w1b); g13.bias1 = gptr(c, w.w3b);
  LnBwdArgs ln{};
  if (!fuse_ln) {
    ln.M = (int)M; ln.D = d; ln.dy = b.dln; ln.x = s.x_mid; ln.stats = s.stats2; ln.gamma = c.wf + w.g2;
    ln.dx_in = dx_src; ln.dx_out = dx_dst; ln.dxb = b.dxb2; ln.rs = rs1; ln.dgamma = gptr(c, w.n2w); ln.dbeta = gptr(c, w.n2b);
    HS_TRY(launch_ln_bwd(ln, c.st));
  }
  // attention output projection
  a = GemmArgs{};
  a.M = (int)M; a.N = d; a.K = d; a.A = b.dxb2; a.lda = d; a.B = c.wb + w.wproj_t; a.ldb = d; a.out0 = b.dao; a.ld0 = d;
  HS_TRY(run_gemm(c, a, kEpiBiasBf16));
  WgradArgs& gp = wj[2];
  gp = WgradArgs{};
  gp.Mred = (int)M; gp.Nout = d; gp.Kin = d; gp.Y = b.dxb2; gp.ldy = d; gp.X = s.ao; gp.ldx = d;
  gp.dst0 = gptr(c, w.pw); gp.ld = d; gp.rows_valid = d; gp.cols_valid = d; gp.bias0 = gptr(c, w.pb);
  // attention backward
  AttnArgs at{}; at.N = N; at.D = d; at.heads = heads; at.s = seq; at.qkv = s.qkv; at.out = s.ao; at.lse = s.lse;
  at.dout = b.dao; at.dqkv = b.dqkv;
  HS_TRY(launch_attn_bwd(at, c.st));
  WgradArgs& gq = wj[3];
  gq = WgradArgs{};
  gq.Mred = (int)M; gq.Nout = 3 * d; gq.Kin = d; gq.Y = b.dqkv; gq.ldy = 3 * d; gq.X = s.ln1; gq.ldx = d;
  gq.dst0 = gptr(c, w.qw); gq.ld = d; gq.rows_valid = 3 * d; gq.cols_valid = d; gq.bias0 = gptr(c, w.qb);
  // dW2 | dW1,dW3 | dWproj | dWq,k,v (+ their bias gradients): one launch, one wave
  HS_TRY(run_wgrad_group(c, wj, 4));
  // q|k|v projection
  a = GemmArgs{};
  a.M = (int)M; a.N = d; a.K = 3 * d; a.A = b.dqkv; a.lda = 3 * d; a.B = c.wb + w.wqkv_t; a.ldb = 3 * d;
  // ... followed by norm1 backward + residual, emitting bf16(rs_prev * dx_in) for the previous block
  if (fuse_ln) {
    a.out0 = dx_dst; a.ld0 = d; a.out1 = want_dxb ? b.dxb : nullptr; a.ld1 = d; a.resid = dx_dst; a.ldr = d; a.lnx = x_in; a.ldx = d;
    a.stats = s.stats1; a.gamma = c.wf + w.g1; a.rs = rs_prev; a.dgamma = gptr(c, w.n1w); a.dbeta = gptr(c, w.n1b);
    return gemm_tc_lnbwd(a, c.st);
  }
  a.out0 = b.dln; a.ld0 = d;
  HS_TRY(run_gemm(c, a, kEpiBiasBf16));
  ln = LnBwdArgs{};
  ln.M = (int)M; ln.D = d; ln.dy = b.dln; ln.x = x_in; ln.stats = s.stats1; ln.gamma = c.wf + w.g1;
  ln.dx_in = dx_dst; ln.dx_out = dx_dst; ln.dxb = want_dxb ? b.dxb : nullptr; ln.rs = rs_prev;
  ln.dgamma = gptr(c, w.n1w); ln.dbeta = gptr(c, w.n1b);
  HS_TRY(launch_ln_bwd(ln, c.st));
  return kOk;
}

RowScale make_rs(const float* const* drop, int idx, int mode, int K, int ll, int G) {
  RowScale r{};
  r.scale = drop ? drop[idx] : nullptr;
  r.mode = mode; r.K = K; r.len_l = ll; r.G = G;
  return r;
}

struct ChainSpec {
  const std::vector<BlockW>* w;
  std::vector<BlockStash>* st;
  SeqSpec seq;
  int mode, G, drop_base;
};

void enc_chains(const hsimae_plan* p, EncLayout& L, ChainSpec out[3]) {
  const int K = L.K, lt = L.lt, ll = L.ll, sd = p->dims.s_depth;
  out[0] = ChainSpec{&p->b1, &L.sp, SeqSpec{K, lt, ll, ll, 1}, 1, lt, 0};
  out[1] = ChainSpec{&p->b2, &L.sc, SeqSpec{K, ll, lt, 1, ll}, 2, ll, 2 * sd};
  out[2] = ChainSpec{&p->bf, &L.fu, SeqSpec{K, 1, K, K, 1}, 3, 1, 4 * sd};
}

int check_enc_args(const hsimae_plan* p, int n, int lt, int ll, const int32_t* ids) {
  HS_REQUIRE(p != nullptr, "null plan");
  HS_REQUIRE(n >= 0, "negative batch");
  HS_REQUIRE(lt >= 1 && lt <= p->g.T && ll >= 1 && ll <= p->g.L, "visible shape (%d,%d) outside (%d,%d)", lt, ll, p->g.T, p->g.L);
  if (!ids) HS_REQUIRE(lt == p->g.T && ll == p->g.L, "unmasked pass requires the full (%d,%d) token grid", p->g.T, p->g.L);
  HS_REQUIRE((int64_t)n * lt * ll < (1ll << 31) / 8, "batch too large for 32-bit row indexing");
  return kOk;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char* hsimae_last_error(void) { return hsimae::last_error(); }
int hsimae_abi_version(void) { return HSIMAE_ABI_VERSION; }
int64_t hsimae_launch_count(void) { return (int64_t)hsimae::launch_count(); }

int hsimae_plan_create(const hsimae_dims* dims, hsimae_plan** out) {
  HS_REQUIRE(dims && out, "null argument");
  hsimae_plan* p = new hsimae_plan();
  int s = build_plan(*dims, p);
  if (s != kOk) { delete p; *out = nullptr; return s; }
  *out = p;
  return kOk;
}
void hsimae_plan_destroy(hsimae_plan* plan) { delete plan; }
int hsimae_plan_num_params(const hsimae_plan* p) { return p ? (int)p->slots.size() : 0; }
const char* hsimae_plan_param_name(const hsimae_plan* p, int i) { return (p && i >= 0 && i < (int)p->slots.size()) ? p->slots[i].name.c_str() : nullptr; }
int64_t hsimae_plan_param_numel(const hsimae_plan* p, int i) { return (p && i >= 0 && i < (int)p->slots.size()) ? p->slots[i].numel : -1; }
int64_t hsimae_plan_param_grad_offset(const hsimae_plan* p, int i) { return (p && i >= 0 && i < (int)p->slots.size()) ? p->slots[i].grad_off : -1; }
int hsimae_plan_param_has_grad(const hsimae_plan* p, int i) { return (p && i >= 0 && i < (int)p->slots.size()) ? (p->slots[i].grad_off >= 0) : 0; }
int64_t hsimae_plan_grad_arena_elems(const hsimae_plan* p) { return p ? p->grad_elems : 0; }
int64_t hsimae_plan_bf16_arena_elems(const hsimae_plan* p) { return p ? p->bf16_elems : 0; }
int64_t hsimae_plan_f32_arena_elems(const hsimae_plan* p) { return p ? p->f32_elems : 0; }
namespace {
int job_tiles_c(const JobT& t) { return ceil_div(t.cols, kPackTileC); }
int job_tiles(const JobT& t) { return ceil_div(t.rows, kPackTileR) * job_tiles_c(t); }
int pack_tiles(const hsimae_plan* p) { int n = 0; for (const JobT& t : p->jobs) n += job_tiles(t); return n; }
}  // namespace
// device table: the jobs followed by the tile -> job map
int64_t hsimae_plan_pack_table_bytes(const hsimae_plan* p) {
  return p ? (int64_t)(p->jobs.size() * sizeof(PackJob) + (size_t)pack_tiles(p) * sizeof(int)) : 0;
}
int hsimae_plan_grad_bucket(const hsimae_plan* p, int i, int64_t* offset, int64_t* elems) {
  HS_REQUIRE(p && offset && elems && i >= 0 && i < 6, "grad_bucket: bad argument");
  const int64_t begin = i == 0 ? 0 : p->bucket_end[i - 1];
  *offset = begin; *elems = p->bucket_end[i] - begin;
  return kOk;
}
int hsimae_plan_hidden(const hsimae_plan* p, int decoder) { return p ? (decoder ? p->Hd : p->H) : 0; }

int hsimae_pack_params(hsimae_plan* p, const void* const* params, void* bf16_arena, void* f32_arena, void* table, void* stream) {
  HS_REQUIRE(p && params && bf16_arena && f32_arena && table, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const size_t np = p->slots.size();
  bool same = p->last_table == table && p->last_ptrs.size() == np;
  if (same) for (size_t i = 0; i < np; ++i) if (p->last_ptrs[i] != params[i]) { same = false; break; }
  if (!same) {
    std::vector<PackJob> jobs(p->jobs.size());
    std::vector<int> tile_job;
    for (size_t i = 0; i < jobs.size(); ++i) {
      const JobT& t = p->jobs[i];
      HS_REQUIRE(params[t.slot] != nullptr, "parameter '%s' is missing", p->slots[t.slot].name.c_str());
      jobs[i] = PackJob{(const float*)params[t.slot], t.dst_off, t.rows, t.cols, t.pitch, t.kind, t.row_map, t.row_off,
                        (int)tile_job.size(), job_tiles_c(t)};
      tile_job.insert(tile_job.end(), (size_t)job_tiles(t), (int)i);
    }
    HS_CHECK_CUDA(cudaMemcpyAsync(table, jobs.data(), jobs.size() * sizeof(PackJob), cudaMemcpyHostToDevice, st));
    HS_CHECK_CUDA(cudaMemcpyAsync(static_cast<char*>(table) + jobs.size() * sizeof(PackJob), tile_job.data(), tile_job.size() * sizeof(int),
                                  cudaMemcpyHostToDevice, st));
    HS_CHECK_CUDA(cudaStreamSynchronize(st));  // `jobs` is a stack-owned staging buffer
    p->last_ptrs.assign(params, params + np);
    p->last_table = table;
  }
  return launch_pack((const PackJob*)table, (int)p->jobs.size(), pack_tiles(p), (bf16*)bf16_arena, (float*)f32_arena, st);
}

int hsimae_mask(const float* noise_t, const float* noise_l, int32_t n, int32_t T, int32_t L, int32_t len_t, int32_t len_l,
                int64_t* ids_keep, int64_t* ids_restore, float* mask, int32_t* ids_keep32, int32_t* ids_restore32, void* stream) {
  if (n == 0) return kOk;
  HS_REQUIRE(noise_t && noise_l && ids_keep && ids_restore && mask, "mask: null argument");
  return launch_mask(noise_t, noise_l, n, T, L, len_t, len_l, ids_keep, ids_restore, mask, ids_keep32, ids_restore32, (cudaStream_t)stream);
}

int64_t hsimae_encoder_workspace_bytes(const hsimae_plan* p, int32_t n, int32_t lt, int32_t ll, int32_t save) {
  if (!p) return -1;
  return enc_layout(p, n, lt, ll, save != 0, nullptr).bytes;
}

static int encoder_forward_impl(hsimae_plan* p, const void* wb, const void* wf, const float* imgs, const float* scene, int scene_w,
                                long long pixel0, int32_t n, int32_t lt, int32_t ll, const int32_t* ids, const float* const* drop,
                                int32_t save, void* ws, int64_t ws_bytes, void* stream) {
  HS_TRY(check_enc_args(p, n, lt, ll, ids));
  HS_REQUIRE(wb && wf && (imgs || scene) && ws, "encoder_forward: null argument");
  EncLayout L = enc_layout(p, n, lt, ll, save != 0, ws);
  HS_REQUIRE(ws_bytes >= L.bytes, "encoder workspace too small: %lld < %lld", (long long)ws_bytes, (long long)L.bytes);
  if (n == 0) return kOk;
  Ctx c{p, (const bf16*)wb, (const float*)wf, nullptr, L.simt, (cudaStream_t)stream, save != 0};
  const float* wff = c.wf;
  const int D = p->D;
  ChainSpec ch[3]; enc_chains(p, L, ch);
  const bool split = !L.sp.empty();
  const bool fusion = !L.fu.empty();

  EmbedArgs e{};
  e.g = p->g; e.N = n; e.K = L.K; e.D = D; e.imgs = imgs; e.scene = scene; e.scene_w = scene_w; e.pixel0 = pixel0; e.W = wff + p->f_pe_w; e.bias = wff + p->f_pe_b; e.pos = wff + p->f_pos;
  e.ids_keep = ids; e.x = L.x0; e.eps = 1e-5f;
  if (split) {
    e.gamma_a = wff + p->b1[0].g1; e.beta_a = wff + p->b1[0].be1; e.ln_a = L.sp[0].ln1; e.stats_a = L.sp[0].stats1;
    e.gamma_b = wff + p->b2[0].g1; e.beta_b = wff + p->b2[0].be1; e.ln_b = L.sc[0].ln1; e.stats_b = L.sc[0].stats1;
  } else if (fusion) {
    e.gamma_a = wff + p->bf[0].g1; e.beta_a = wff + p->bf[0].be1; e.ln_a = L.fu[0].ln1; e.stats_a = L.fu[0].stats1;
  } else {
    e.gamma_a = wff + p->f_norm_g; e.beta_a = wff + p->f_norm_b; e.ln_a = L.latent; e.stats_a = L.stats_f;
  }
  HS_TRY(launch_embed_fwd(e, c.st));

  // spatial chain on the caller's stream, spectral chain on the helper stream (see helper())
  Helper& hp = helper();
  const bool overlap = split && hp.ok && !p->debug_simt;
  Ctx cctx[3] = {c, c, c};
  if (overlap) {
    cctx[1].st = hp.st;
    HS_CHECK_CUDA(cudaEventRecord(hp.fork, c.st));
    HS_CHECK_CUDA(cudaStreamWaitEvent(hp.st, hp.fork, 0));
  }
  TailLN final_ln{wff + p->f_norm_g, wff + p->f_norm_b, L.latent, L.stats_f, nullptr};
  const float* x_in = L.x0;
  for (int ci = 0; ci < 3; ++ci) {
    ChainSpec& cs = ch[ci];
    const int nb = (int)cs.st->size();
    if (nb == 0) continue;
    const Ctx& c = cctx[ci];
    if (ci == 1) x_in = L.x0;             // the spectral encoder restarts from the embedded tokens
    for (int i = 0; i < nb; ++i) {
      // the last spectral block adds the spatial result (x = x1 + x2): it waits for the spatial chain
      if (overlap && ci == 1 && i == nb - 1) HS_CHECK_CUDA(cudaStreamWaitEvent(hp.st, hp.aux, 0));
      TailLN tail{};
      if (i + 1 < nb) {
        const BlockW& nx = (*cs.w)[i + 1];
        tail = TailLN{wff + nx.g1, wff + nx.be1, (*cs.st)[i + 1].ln1, (*cs.st)[i + 1].stats1, nullptr};
      } else if (ci == 0) {
        tail = TailLN{nullptr, nullptr, nullptr, nullptr, nullptr};          // spatial result is only summed later
      } else if (ci == 1) {
        if (fusion) tail = TailLN{wff + p->bf[0].g1, wff + p->bf[0].be1, L.fu[0].ln1, L.fu[0].stats1, nullptr};
        else tail = final_ln;
        tail.resid2 = L.sp.back().x_out;                                     // x = x1 + x2 (Models.py:564)
      } else {
        tail = final_ln;
      }
      RowScale rs1 = make_rs(drop, cs.drop_base + 2 * i, cs.mode, L.K, L.ll, cs.G);
      RowScale rs2 = make_rs(drop, cs.drop_base + 2 * i + 1, cs.mode, L.K, L.ll, cs.G);
      HS_TRY(block_forward(c, (*cs.w)[i], L.M, n, D, p->Hp, p->heads, cs.seq, (*cs.st)[i], x_in, rs1, rs2, tail));
      x_in = (*cs.st)[i].x_out;
    }
    if (overlap && ci == 0) HS_CHECK_CUDA(cudaEventRecord(hp.aux, cctx[0].st));           // spatial chain enqueued
    if (overlap && ci == 1) {                                                             // join: the fusion blocks follow on the caller's stream
      HS_CHECK_CUDA(cudaEventRecord(hp.join, hp.st));
      HS_CHECK_CUDA(cudaStreamWaitEvent(cctx[0].st, hp.join, 0));
    }
  }
  return kOk;
}

int hsimae_encoder_forward(hsimae_plan* p, const void* wb, const void* wf, const float* imgs, int32_t n, int32_t lt, int32_t ll,
                           const int32_t* ids, const float* const* drop, int32_t save, void* ws, int64_t ws_bytes, void* stream) {
  HS_REQUIRE(imgs != nullptr, "encoder_forward: null argument");
  return encoder_forward_impl(p, wb, wf, imgs, nullptr, 0, 0, n, lt, ll, ids, drop, save, ws, ws_bytes, stream);
}

int hsimae_encoder_forward_scene(hsimae_plan* p, const void* wb, const void* wf, const float* scene, int32_t scene_h, int32_t scene_w,
                                 int64_t pixel0, int32_t n, void* ws, int64_t ws_bytes, void* stream) {
  HS_REQUIRE(p && scene, "encoder_forward_scene: null argument");
  const int img = p->g.img;
  HS_REQUIRE(scene_h >= img && scene_w >= img, "scene %dx%d smaller than the %dx%d window", scene_h, scene_w, img, img);
  const int64_t windows = (int64_t)(scene_h - img + 1) * (scene_w - img + 1);
  HS_REQUIRE(pixel0 >= 0 && n >= 0 && pixel0 + n <= windows, "window range [%lld, %lld) outside the %lld windows of the scene",
             (long long)pixel0, (long long)(pixel0 + n), (long long)windows);
  return encoder_forward_impl(p, wb, wf, nullptr, scene, scene_w, pixel0, n, p->g.T, p->g.L, nullptr, nullptr, 0, ws, ws_bytes, stream);
}

int hsimae_encoder_backward(hsimae_plan* p, const void* wb, const void* wf, const float* imgs, int32_t n, int32_t lt, int32_t ll,
                            const int32_t* ids, const float* const* drop, void* ws, int64_t ws_bytes, float* grads, int32_t stages,
                            void* stream) {
  HS_TRY(check_enc_args(p, n, lt, ll, ids));
  HS_REQUIRE(wb && wf && imgs && ws && grads, "encoder_backward: null argument");
  EncLayout L = enc_layout(p, n, lt, ll, true, ws);
  HS_REQUIRE(ws_bytes >= L.bytes, "encoder workspace too small: %lld < %lld", (long long)ws_bytes, (long long)L.bytes);
  if (n == 0) return kOk;
  Ctx c{p, (const bf16*)wb, (const float*)wf, grads, L.simt, (cudaStream_t)stream};
  const int D = p->D;
  ChainSpec ch[3]; enc_chains(p, L, ch);
  const bool split = !L.sp.empty();
  const bool fusion = !L.fu.empty();
  BwdScratch& b = L.bw;

  // rs2 of the block that produced x_final (needed for the bf16 operand of its down-projection backward)
  auto rs2_of = [&](int ci, int i) { return make_rs(drop, ch[ci].drop_base + 2 * i + 1, ch[ci].mode, L.K, L.ll, ch[ci].G); };
  auto rs1_of = [&](int ci, int i) { return make_rs(drop, ch[ci].drop_base + 2 * i, ch[ci].mode, L.K, L.ll, ch[ci].G); };
  RowScale none{};

  if (stages & 1) {
  // final norm backward (Models.py:570)
  LnBwdArgs ln{};
  ln.M = (int)L.M; ln.D = D; ln.dy = L.dlatent; ln.x = L.x_final; ln.stats = L.stats_f; ln.gamma = c.wf + p->f_norm_g;
  ln.dx_in = nullptr; ln.dx_out = b.dxA; ln.dgamma = gptr(c, p->p_norm_w); ln.dbeta = gptr(c, p->p_norm_b);
  if (fusion) { ln.dxb = b.dxb; ln.rs = rs2_of(2, (int)L.fu.size() - 1); }
  HS_TRY(launch_ln_bwd(ln, c.st));

  // fusion blocks, last to first
  for (int i = (int)L.fu.size() - 1; i >= 0; --i) {
    const float* x_in = i > 0 ? L.fu[i - 1].x_out : (split ? L.sc.back().x_out : L.x0);
    const bool want = i > 0;
    RowScale prev = i > 0 ? rs2_of(2, i - 1) : none;
    HS_TRY(block_backward(c, p->bf[i], L.M, n, D, p->H, p->Hp, p->heads, ch[2].seq, L.fu[i], x_in, rs1_of(2, i), prev, want,
                          b.dxA, b.dxA, b));
  }
  }  // stage 1
  const float* dx_embed_a = b.dxA;
  const float* dx_embed_b = nullptr;
  Helper& hp = helper();
  if (split) {
    const int sd = (int)L.sp.size();
    // spectral encoder: reads the summed gradient dxA, writes its own stream dxB.  With stage bit 32 it runs on the helper
    // stream (own scratch) beside the spatial chain and is joined before the patch-embedding backward / by hsimae_helper_join.
    if (stages & 2) {
      const bool async = (stages & 32) != 0 && hp.ok && !p->debug_simt;
      Ctx cs = c;
      BwdScratch& bs = async ? L.bw2 : b;
      if (async) {
        HS_CHECK_CUDA(cudaEventRecord(hp.fork, c.st));
        HS_CHECK_CUDA(cudaStreamWaitEvent(hp.st, hp.fork, 0));
        cs.st = hp.st;
      }
      HS_TRY(launch_scale_cast(b.dxA, bs.dxb, (int)L.M, D, rs2_of(1, sd - 1), cs.st));
      for (int i = sd - 1; i >= 0; --i) {
        const float* x_in = i > 0 ? L.sc[i - 1].x_out : L.x0;
        RowScale prev = i > 0 ? rs2_of(1, i - 1) : none;
        HS_TRY(block_backward(cs, p->b2[i], L.M, n, D, p->H, p->Hp, p->heads, ch[1].seq, L.sc[i], x_in, rs1_of(1, i), prev, i > 0,
                              i == sd - 1 ? b.dxA : b.dxB, b.dxB, bs));
      }
      if (async) { HS_CHECK_CUDA(cudaEventRecord(hp.join, hp.st)); p->helper_pending = true; }
    }  // stage 2
    // spatial encoder in three sub-stages (4: last third of the blocks, 8: middle, 16: first third).  Its gradient stream
    // is dxC: the summed gradient dxA stays read-only, the spectral chain may still be reading it.
    if (stages & 4) HS_TRY(launch_scale_cast(b.dxA, b.dxb, (int)L.M, D, rs2_of(0, sd - 1), c.st));
    for (int i = sd - 1; i >= 0; --i) {
      const int bit = i >= p->sp_cut[1] ? 4 : (i >= p->sp_cut[0] ? 8 : 16);
      if (!(stages & bit)) continue;
      const float* x_in = i > 0 ? L.sp[i - 1].x_out : L.x0;
      RowScale prev = i > 0 ? rs2_of(0, i - 1) : none;
      HS_TRY(block_backward(c, p->b1[i], L.M, n, D, p->H, p->Hp, p->heads, ch[0].seq, L.sp[i], x_in, rs1_of(0, i), prev, i > 0,
                            i == sd - 1 ? b.dxA : L.dxC, L.dxC, b));
    }
    dx_embed_a = L.dxC;
    dx_embed_b = b.dxB;
  }
  if (!(stages & 16)) return kOk;
  if (p->helper_pending) {   // the patch-embedding backward sums both chains' gradients
    HS_CHECK_CUDA(cudaStreamWaitEvent(c.st, hp.join, 0));
    p->helper_pending = false;
  }
  EmbedBwdArgs e{};
  e.g = p->g; e.N = n; e.K = L.K; e.D = D; e.imgs = imgs; e.ids_keep = ids; e.dx_a = dx_embed_a; e.dx_b = dx_embed_b;
  e.dW = gptr(c, p->p_pe_w); e.dbias = gptr(c, p->p_pe_b);
  return launch_embed_bwd(e, c.st);
}

int hsimae_helper_join(hsimae_plan* p, void* stream) {
  HS_REQUIRE(p != nullptr, "null plan");
  if (p->helper_pending) {
    HS_CHECK_CUDA(cudaStreamWaitEvent((cudaStream_t)stream, helper().join, 0));
    p->helper_pending = false;
  }
  return kOk;
}

__global__ void latent_kernel(const float* __restrict__ x, const float* __restrict__ stats, const float* __restrict__ g,
                              const float* __restrict__ be, float* __restrict__ out, int64_t M, int D) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < M * D; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t m = i / D; const int d = (int)(i - m * D);
    out[i] = fmaf((x[i] - stats[2 * m]) * stats[2 * m + 1], g[d], be[d]);
  }
}

int hsimae_encoder_latent(const hsimae_plan* p, int32_t n, int32_t lt, int32_t ll, int32_t save, const void* wf, const void* ws,
                          float* out, void* stream) {
  HS_REQUIRE(p && wf && ws && out, "encoder_latent: null argument");
  EncLayout L = enc_layout(p, n, lt, ll, save != 0, ws);
  if (n == 0) return kOk;
  const float* f = (const float*)wf;
  latent_kernel<<<2 * kNumSMs, 256, 0, (cudaStream_t)stream>>>(L.x_final, L.stats_f, f + p->f_norm_g, f + p->f_norm_b, out, L.M, p->D);
  HS_CHECK_LAUNCH("latent_kernel");
  return kOk;
}

int64_t hsimae_decoder_workspace_bytes(const hsimae_plan* p, int32_t n, int32_t lt, int32_t ll, int32_t save) {
  if (!p || p->Dd == 0) return -1;
  return dec_layout(p, n, lt, ll, save != 0, nullptr).bytes;
}

int hsimae_decoder_forward(hsimae_plan* p, const void* wb, const void* wf, const float* imgs, int32_t n, int32_t lt, int32_t ll,
                           const int32_t* ids_restore, const float* mask, const void* enc_ws, int32_t enc_save, int32_t save,
                           void* ws, int64_t ws_bytes, float* loss, float* pred_img, float* mask_img, float* pred_tokens,
                           void* stream) {
  HS_REQUIRE(p && p->Dd > 0, "this model has no decoder");
  HS_REQUIRE(wb && wf && imgs && ids_restore && mask && enc_ws && ws && loss, "decoder_forward: null argument");
  EncLayout E = enc_layout(p, n, lt, ll, enc_save != 0, enc_ws);
  DecLayout L = dec_layout(p, n, lt, ll, save != 0, ws);
  HS_REQUIRE(ws_bytes >= L.bytes, "decoder workspace too small: %lld < %lld", (long long)ws_bytes, (long long)L.bytes);
  HS_REQUIRE(L.P > L.K, "nothing is masked: the reconstruction loss is undefined (Models.py:615 divides by mask.sum())");
  if (n == 0) return kOk;
  Ctx c{p, (const bf16*)wb, (const float*)wf, nullptr, L.simt, (cudaStream_t)stream, save != 0};
  const float* wff = c.wf;
  const int D = p->D, Dd = p->Dd;
  // decoder_embed (Models.py:579)
  GemmArgs a{};
  a.M = (int)L.M; a.N = Dd; a.K = D; a.A = E.latent; a.lda = D; a.B = c.wb + p->w_de; a.ldb = D;
  a.out0 = L.y; a.ld0 = Dd; a.bias = wff + p->f_de_b;
  HS_TRY(run_gemm(c, a, kEpiBiasF32));
  // mean-fill + unshuffle + pos (+ first LayerNorm)
  FillArgs f{};
  f.N = n; f.K = L.K; f.P = L.P; f.D = Dd; f.y = L.y; f.ids_restore = ids_restore; f.pos = wff + p->f_dpos; f.x = L.x0; f.eps = 1e-5f;
  if (!L.blk.empty()) { f.gamma = wff + p->bd[0].g1; f.beta = wff + p->bd[0].be1; f.ln = L.blk[0].ln1; f.stats = L.blk[0].stats1; }
  else { f.gamma = wff + p->f_dnorm_g; f.beta = wff + p->f_dnorm_b; f.ln = L.dnorm; f.stats = L.stats_n; }
  HS_TRY(launch_fill_fwd(f, c.st));
  const SeqSpec seq{L.P, 1, L.P, L.P, 1};
  const float* x_in = L.x0;
  RowScale none{};
  for (size_t i = 0; i < L.blk.size(); ++i) {
    TailLN tail;
    if (i + 1 < L.blk.size()) tail = TailLN{wff + p->bd[i + 1].g1, wff + p->bd[i + 1].be1, L.blk[i + 1].ln1, L.blk[i + 1].stats1, nullptr};
    else tail = TailLN{wff + p->f_dnorm_g, wff + p->f_dnorm_b, L.dnorm, L.stats_n, nullptr};
    HS_TRY(block_forward(c, p->bd[i], L.Md, n, Dd, p->Hdp, p->dheads, seq, L.blk[i], x_in, none, none, tail));
    x_in = L.blk[i].x_out;
  }
  // decoder_pred (Models.py:600)
  a = GemmArgs{};
  a.M = (int)L.Md; a.N = p->PKp; a.K = Dd; a.A = L.dnorm; a.lda = Dd; a.B = c.wb + p->w_pred; a.ldb = Dd;
  a.out0 = L.pred; a.ld0 = p->PKp; a.bias = wff + p->f_pred_b;
  if (p->debug_simt) { Ctx c2 = c; c2.simt = L.simt + L.Md * (int64_t)(3 * Dd > 2 * p->Hdp ? 3 * Dd : 2 * p->Hdp); HS_TRY(run_gemm(c2, a, kEpiBiasF32)); }
  else HS_TRY(run_gemm(c, a, kEpiBiasF32));
  // loss + pixel outputs
  LossArgs l{};
  l.g = p->g; l.N = n; l.norm_pix = p->dims.norm_pix_loss; l.imgs = imgs; l.pred = L.pred; l.ldp = p->PKp; l.mask = mask;
  l.mask_sum = (float)((int64_t)n * (L.P - L.K));
  l.loss_partial = L.loss_partial; l.loss = loss; l.dpred = L.dpred_unit; l.ldd = p->PKp; l.pred_img = pred_img; l.mask_img = mask_img;
  HS_TRY(launch_loss(l, c.st));
  if (pred_tokens)
    HS_CHECK_CUDA(cudaMemcpy2DAsync(pred_tokens, (size_t)p->g.PK * 4, L.pred, (size_t)p->PKp * 4, (size_t)p->g.PK * 4, (size_t)L.Md,
                                    cudaMemcpyDeviceToDevice, c.st));
  return kOk;
}

int hsimae_decoder_backward(hsimae_plan* p, const void* wb, const void* wf, int32_t n, int32_t lt, int32_t ll,
                            const int32_t* ids_restore, void* enc_ws, void* ws, int64_t ws_bytes, const float* grad_loss,
                            float* grads, void* stream) {
  HS_REQUIRE(p && p->Dd > 0, "this model has no decoder");
  HS_REQUIRE(wb && wf && ids_restore && enc_ws && ws && grad_loss && grads, "decoder_backward: null argument");
  EncLayout E = enc_layout(p, n, lt, ll, true, enc_ws);
  DecLayout L = dec_layout(p, n, lt, ll, true, ws);
  HS_REQUIRE(ws_bytes >= L.bytes, "decoder workspace too small");
  if (n == 0) return kOk;
  Ctx c{p, (const bf16*)wb, (const float*)wf, grads, L.simt, (cudaStream_t)stream};
  const int D = p->D, Dd = p->Dd;
  BwdScratch& b = L.bw;
  RowScale none{};
  // dL/dpred = grad_loss * unit gradient
  HS_TRY(launch_scale_bf16(L.dpred_unit, L.dpred, L.Md * p->PKp, grad_loss, c.st));
  WgradArgs g{};
  g.Mred = (int)L.Md; g.Nout = p->PKp; g.Kin = Dd; g.Y = L.dpred; g.ldy = p->PKp; g.X = L.dnorm; g.ldx = Dd;
  g.dst0 = gptr(c, p->p_pred_w); g.ld = Dd; g.rows_valid = p->g.PK; g.cols_valid = Dd; g.bias0 = gptr(c, p->p_pred_b);
  HS_TRY(run_wgrad(c, g));
  GemmArgs a{};
  a.M = (int)L.Md; a.N = Dd; a.K = p->PKp; a.A = L.dpred; a.lda = p->PKp; a.B = c.wb + p->w_pred_t; a.ldb = p->PKp;
  a.out0 = b.dln; a.ld0 = Dd;
  HS_TRY(run_gemm(c, a, kEpiBiasBf16));
  LnBwdArgs ln{};
  ln.M = (int)L.Md; ln.D = Dd; ln.dy = b.dln; ln.x = L.x_final; ln.stats = L.stats_n; ln.gamma = c.wf + p->f_dnorm_g;
  ln.dx_out = b.dxA; ln.dxb = L.blk.empty() ? nullptr : b.dxb; ln.rs = none;
  ln.dgamma = gptr(c, p->p_dnorm_w); ln.dbeta = gptr(c, p->p_dnorm_b);
  HS_TRY(launch_ln_bwd(ln, c.st));
  const SeqSpec seq{L.P, 1, L.P, L.P, 1};
  for (int i = (int)L.blk.size() - 1; i >= 0; --i) {
    const float* x_in = i > 0 ? L.blk[i - 1].x_out : L.x0;
    HS_TRY(block_backward(c, p->bd[i], L.Md, n, Dd, p->Hd, p->Hdp, p->dheads, seq, L.blk[i], x_in, none, none, i > 0, b.dxA, b.dxA, b));
  }
  FillArgs f{};
  f.N = n; f.K = L.K; f.P = L.P; f.D = Dd; f.ids_restore = ids_restore; f.dx = b.dxA; f.dy = L.dy;
  HS_TRY(launch_fill_bwd(f, c.st));
  g = WgradArgs{};
  g.Mred = (int)L.M; g.Nout = Dd; g.Kin = D; g.Y = L.dy; g.ldy = Dd; g.X = E.latent; g.ldx = D;
  g.dst0 = gptr(c, p->p_de_w); g.ld = D; g.rows_valid = Dd; g.cols_valid = D; g.bias0 = gptr(c, p->p_de_b);
  HS_TRY(run_wgrad(c, g));
  a = GemmArgs{};
  a.M = (int)L.M; a.N = D; a.K = Dd; a.A = L.dy; a.lda = Dd; a.B = c.wb + p->w_de_t; a.ldb = Dd; a.out0 = E.dlatent; a.ld0 = D;
  if (p->debug_simt) { Ctx c2 = c; c2.simt = E.simt; HS_TRY(run_gemm(c2, a, kEpiBiasBf16)); }
  else HS_TRY(run_gemm(c, a, kEpiBiasBf16));
  return kOk;
}

int hsimae_head_forward(hsimae_plan* p, const void* wf, int32_t n, const void* enc_ws, int32_t enc_save, float* pooled, float* logits,
                        void* stream) {
  HS_REQUIRE(p && p->dims.num_class > 0, "this model has no classification head");
  HS_REQUIRE(wf && enc_ws && pooled && logits, "head_forward: null argument");
  EncLayout E = enc_layout(p, n, p->g.T, p->g.L, enc_save != 0, enc_ws);
  const float* f = (const float*)wf;
  HeadArgs h{};
  h.N = n; h.T = p->g.T; h.L = p->g.L; h.D = p->D; h.C = p->dims.num_class; h.x = E.x_final; h.stats = E.stats_f;
  h.gamma = f + p->f_norm_g; h.beta = f + p->f_norm_b; h.W = f + p->f_cls_w; h.bias = f + p->f_cls_b; h.z = pooled; h.logits = logits;
  return launch_head_fwd(h, (cudaStream_t)stream);
}

int hsimae_head_backward(hsimae_plan* p, const void* wf, int32_t n, void* enc_ws, const float* pooled, const float* dlogits,
                         float* grads, void* stream) {
  HS_REQUIRE(p && p->dims.num_class > 0, "this model has no classification head");
  HS_REQUIRE(wf && enc_ws && pooled && dlogits && grads, "head_backward: null argument");
  EncLayout E = enc_layout(p, n, p->g.T, p->g.L, true, enc_ws);
  const float* f = (const float*)wf;
  HeadArgs h{};
  h.N = n; h.T = p->g.T; h.L = p->g.L; h.D = p->D; h.C = p->dims.num_class; h.W = f + p->f_cls_w; h.z = const_cast<float*>(pooled);
  h.dlogits = dlogits; h.dW = grads + p->slots[p->p_cls_w].grad_off; h.dbias = grads + p->slots[p->p_cls_b].grad_off;
  h.dlatent = E.dlatent;
  return launch_head_bwd(h, (cudaStream_t)stream);
}

// ---- single operators -----------------------------------------------------------
int hsimae_gemm(const hsimae_gemm_desc* d, void* stream) {
  HS_REQUIRE(d != nullptr, "null descriptor");
  GemmArgs a{};
  a.M = d->M; a.N = d->N; a.K = d->K; a.A = (const bf16*)d->A; a.lda = d->lda; a.B = (const bf16*)d->B; a.ldb = d->ldb;
  a.out0 = d->out0; a.ld0 = d->ld0; a.out1 = d->out1; a.ld1 = d->ld1; a.bias = d->bias; a.resid = d->resid; a.ldr = d->ldr;
  a.resid2 = d->resid2; a.gamma = d->gamma; a.beta = d->beta; a.stats = d->stats; a.ab = (const bf16*)d->ab; a.ldab = d->ldab;
  a.A2 = (const bf16*)d->A2; a.lda2 = d->lda2; a.B2 = (const bf16*)d->B2; a.ldb2 = d->ldb2;
  a.rs.scale = d->rowscale; a.rs.mode = d->rs_mode; a.rs.K = d->rs_K > 0 ? d->rs_K : 1; a.rs.len_l = d->rs_len_l > 0 ? d->rs_len_l : 1; a.rs.G = d->rs_G;
  a.ln_eps = 1e-5f;
  if (d->impl == 1) return gemm_simt(a, d->epilogue, d->scratch, (cudaStream_t)stream);
  return gemm_tc(a, d->epilogue, (cudaStream_t)stream);
}

int hsimae_gemm_lnbwd(const hsimae_lnbwd_desc* d, void* stream) {
  HS_REQUIRE(d != nullptr, "null descriptor");
  GemmArgs a{};
  a.M = d->M; a.N = d->N; a.K = d->K; a.A = (const bf16*)d->A; a.lda = d->lda; a.B = (const bf16*)d->B; a.ldb = d->ldb;
  a.out0 = d->dx_out; a.ld0 = d->ldo; a.out1 = d->dxb; a.ld1 = d->ldxb; a.resid = d->dx_in; a.ldr = d->ldi;
  a.lnx = d->x; a.ldx = d->ldx; a.stats = const_cast<float*>(d->stats); a.gamma = d->gamma; a.dgamma = d->dgamma; a.dbeta = d->dbeta;
  a.rs.scale = d->rowscale; a.rs.mode = d->rs_mode; a.rs.K = d->rs_K > 0 ? d->rs_K : 1; a.rs.len_l = d->rs_len_l > 0 ? d->rs_len_l : 1; a.rs.G = d->rs_G;
  a.ln_eps = 1e-5f;
  return gemm_tc_lnbwd(a, (cudaStream_t)stream);
}

int hsimae_mlp_fused(const hsimae_mlp_desc* d, void* stream) {
  HS_REQUIRE(d != nullptr, "null descriptor");
  MlpFusedArgs m{};
  GemmArgs& t = m.tail;
  t.M = d->M; t.N = d->D; t.K = d->Hp; t.B = (const bf16*)d->W2; t.ldb = d->ldw2; t.bias = d->b2;
  t.out0 = d->out; t.ld0 = d->ldo; t.resid = d->resid; t.ldr = d->ldr; t.resid2 = d->resid2;
  t.gamma = d->gamma; t.beta = d->beta; t.out1 = d->ln; t.ld1 = d->ldln; t.stats = d->stats; t.ln_eps = 1e-5f;
  t.rs.scale = d->rowscale; t.rs.mode = d->rs_mode; t.rs.K = d->rs_K > 0 ? d->rs_K : 1; t.rs.len_l = d->rs_len_l > 0 ? d->rs_len_l : 1; t.rs.G = d->rs_G;
  m.X = (const bf16*)d->X; m.ldx = d->ldx; m.W13 = (const bf16*)d->W13; m.ldw = d->ldw13; m.b13 = d->b13;
  m.g = (bf16*)d->g; m.ldg = d->ldg;
  return mlp_fused(m, (cudaStream_t)stream);
}

int hsimae_wgrad(const hsimae_wgrad_desc* d, void* stream) {
  HS_REQUIRE(d != nullptr, "null descriptor");
  WgradArgs a{};
  a.Mred = d->Mred; a.Nout = d->Nout; a.Kin = d->Kin; a.Y = (const bf16*)d->Y; a.ldy = d->ldy; a.X = (const bf16*)d->X; a.ldx = d->ldx;
  a.dst0 = d->dst0; a.dst1 = d->dst1; a.ld = d->ld; a.row_map = d->row_map; a.rows_valid = d->rows_valid; a.cols_valid = d->cols_valid;
  a.bias0 = d->bias0; a.bias1 = d->bias1;
  if (d->impl == 1) return wgrad_simt(a, (cudaStream_t)stream);
  return wgrad_tc(a, (cudaStream_t)stream);
}

int hsimae_wgrad_group(const hsimae_wgrad_desc* d, int32_t n, void* stream) {
  HS_REQUIRE(d != nullptr && n >= 1 && n <= 4, "wgrad_group: 1..4 descriptors");
  WgradArgs jobs[4];
  for (int i = 0; i < n; ++i) {
    WgradArgs& a = jobs[i];
    a = WgradArgs{};
    a.Mred = d[i].Mred; a.Nout = d[i].Nout; a.Kin = d[i].Kin; a.Y = (const bf16*)d[i].Y; a.ldy = d[i].ldy; a.X = (const bf16*)d[i].X; a.ldx = d[i].ldx;
    a.dst0 = d[i].dst0; a.dst1 = d[i].dst1; a.ld = d[i].ld; a.row_map = d[i].row_map; a.rows_valid = d[i].rows_valid; a.cols_valid = d[i].cols_valid;
    a.bias0 = d[i].bias0; a.bias1 = d[i].bias1;
  }
  return wgrad_tc_group(jobs, n, (cudaStream_t)stream);
}

int hsimae_attention_forward(const void* qkv, void* out, float* lse, int32_t n, int32_t D, int32_t heads, int32_t K, int32_t nseq,
                             int32_t len, int32_t seq_step, int32_t tok_step, void* stream) {
  AttnArgs a{}; a.N = n; a.D = D; a.heads = heads; a.s = SeqSpec{K, nseq, len, seq_step, tok_step};
  a.qkv = (const bf16*)qkv; a.out = (bf16*)out; a.lse = lse;
  return launch_attn_fwd(a, (cudaStream_t)stream);
}

int hsimae_attention_backward(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, int32_t n, int32_t D,
                              int32_t heads, int32_t K, int32_t nseq, int32_t len, int32_t seq_step, int32_t tok_step, void* stream) {
  AttnArgs a{}; a.N = n; a.D = D; a.heads = heads; a.s = SeqSpec{K, nseq, len, seq_step, tok_step};
  a.qkv = (const bf16*)qkv; a.out = (bf16*)const_cast<void*>(out); a.lse = const_cast<float*>(lse); a.dout = (const bf16*)dout; a.dqkv = (bf16*)dqkv;
  return launch_attn_bwd(a, (cudaStream_t)stream);
}

}  // extern "C"
