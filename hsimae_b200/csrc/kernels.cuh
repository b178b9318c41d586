// Launchers of the non-GEMM kernels (HBM-bound / CUDA-core work).
#pragma once
#include "common.cuh"

namespace hsimae {

// Geometry of the patch grid (Models.py:107-149)
struct PatchGeom {
  int bands, img;        // cube is [bands, img, img]
  int u, p;              // spectral / spatial patch edge
  int T, G, L, P;        // spectral groups, grid side, G*G, T*L
  int PK;                // u*p*p elements per patch
  int cube;              // bands*img*img elements per sample
};

// ---- masking (Models.py:495-535) -------------------------------------------
int launch_mask(const float* noise_t, const float* noise_l, int N, int T, int L, int len_t, int len_l,
                int64_t* ids_keep, int64_t* ids_restore, float* mask, int32_t* ids_keep32, int32_t* ids_restore32,
                cudaStream_t stream);

// ---- patch embedding (Models.py:147-158, 547-550) ---------------------------
struct EmbedArgs {
  PatchGeom g;
  int N, K, D;                 // samples, kept tokens per sample, embed dim
  const float* imgs;           // [N, cube]   (nullptr when `scene` is given)
  // optional on-device sliding window: sample n is the img x img window whose top-left corner is pixel
  // (pixel0 + n) of a [scene_h, scene_w, bands] HWC scene, row-major over (scene_h - img + 1) x (scene_w - img + 1)
  // window positions (replaces the host-side cube materialisation of Utils/Preprocessing.py:205-213)
  const float* scene; int scene_w; long long pixel0;
  const float* W;              // [D, PK]  (Conv3d weight viewed 2-D)
  const float* bias;           // [D]
  const float* pos;            // [P, D]
  const int32_t* ids_keep;     // [N, K] or nullptr (identity, K == P)
  float* x;                    // [N*K, D] fp32 out
  const float* gamma_a; const float* beta_a; __nv_bfloat16* ln_a; float* stats_a;  // first LayerNorm consumer
  const float* gamma_b; const float* beta_b; __nv_bfloat16* ln_b; float* stats_b;  // optional second one
  float eps;
};
int launch_embed_fwd(const EmbedArgs& a, cudaStream_t stream);

struct EmbedBwdArgs {
  PatchGeom g;
  int N, K, D;
  const float* imgs;
  const int32_t* ids_keep;
  const float* dx_a;           // [N*K, D]
  const float* dx_b;           // optional, summed with dx_a
  float* dW;                   // [D, PK] accumulated
  float* dbias;                // [D] accumulated
};
int launch_embed_bwd(const EmbedBwdArgs& a, cudaStream_t stream);
// tensor-core (TF32 mma.sync) forms for the reference's 8x3x3 patch geometry (embed_mma.cu); the launchers above pick them
bool embed_fwd_mma_supported(const EmbedArgs& a);
int launch_embed_fwd_mma(const EmbedArgs& a, cudaStream_t stream);
bool embed_bwd_mma_supported(const EmbedBwdArgs& a);
int launch_embed_bwd_mma(const EmbedBwdArgs& a, cudaStream_t stream);

// ---- LayerNorm backward (+ residual-gradient add) ---------------------------
struct LnBwdArgs {
  int M, D;
  const __nv_bfloat16* dy;     // [M, D] gradient w.r.t. the LayerNorm output
  const float* x;              // [M, D] LayerNorm input
  const float* stats;          // [M, 2]
  const float* gamma;
  const float* dx_in;          // optional fp32 [M, D] added to the result
  float* dx_out;               // fp32 [M, D]
  __nv_bfloat16* dxb;          // optional bf16 copy of rs * dx_out (next GEMM operand)
  RowScale rs;
  float* dgamma; float* dbeta; // accumulated
};
int launch_ln_bwd(const LnBwdArgs& a, cudaStream_t stream);

// bf16(rs * x)
int launch_scale_cast(const float* x, __nv_bfloat16* out, int M, int D, RowScale rs, cudaStream_t stream);
// out = bf16(scale[0] * in)   (scale on device)
int launch_scale_bf16(const __nv_bfloat16* in, __nv_bfloat16* out, int64_t n, const float* scale, cudaStream_t stream);

// ---- attention over short token groups (Models.py:192-215) -------------------
struct SeqSpec {
  int K;        // token rows per sample
  int nseq;     // sequences per sample
  int len;      // tokens per sequence
  int seq_step; // row offset between sequence starts
  int tok_step; // row offset between tokens of a sequence
};
struct AttnArgs {
  int N, D, heads;
  SeqSpec s;
  const __nv_bfloat16* qkv;    // [N*K, 3D]
  __nv_bfloat16* out;          // [N*K, D]
  float* lse;                  // [N*K, heads]
  // backward
  const __nv_bfloat16* dout;   // [N*K, D]
  __nv_bfloat16* dqkv;         // [N*K, 3D]
};
int launch_attn_fwd(const AttnArgs& a, cudaStream_t stream);
int launch_attn_bwd(const AttnArgs& a, cudaStream_t stream);

// ---- decoder token fill / unshuffle (Models.py:583-592) -----------------------
struct FillArgs {
  int N, K, P, D;
  const float* y;              // [N*K, D] decoder-embedded visible tokens
  const int32_t* ids_restore;  // [N, P]
  const float* pos;            // [P, D]
  float* x;                    // [N*P, D]
  const float* gamma; const float* beta; __nv_bfloat16* ln; float* stats; float eps;
  // backward
  const float* dx;             // [N*P, D]
  __nv_bfloat16* dy;           // [N*K, D]
};
int launch_fill_fwd(const FillArgs& a, cudaStream_t stream);
int launch_fill_bwd(const FillArgs& a, cudaStream_t stream);

// ---- reconstruction loss + pixel outputs (Models.py:603-625) -----------------
struct LossArgs {
  PatchGeom g;
  int N;
  int norm_pix;
  const float* imgs;           // [N, cube]
  const float* pred; int ldp;  // [N*P, ldp] fp32 (ldp >= PK)
  const float* mask;           // [N, P]
  float mask_sum;              // number of masked patches over the batch
  float* loss_partial;         // [N]
  float* loss;                 // [1]
  __nv_bfloat16* dpred; int ldd;  // [N*P, ldd] unit-scale gradient (optional)
  float* pred_img;             // [N, cube] (optional)
  float* mask_img;             // [N, cube] (optional)
};
int launch_loss(const LossArgs& a, cudaStream_t stream);

// ---- classification head (Models.py:964-973) ----------------------------------
struct HeadArgs {
  int N, T, L, D, C;
  const float* x;              // [N*T*L, D] pre-norm encoder output
  const float* stats;          // [N*T*L, 2]
  const float* gamma; const float* beta;
  const float* W;              // [C, T*D]
  const float* bias;           // [C]
  float* z;                    // [N, T*D] pooled features (saved)
  float* logits;               // [N, C]
  // backward
  const float* dlogits;        // [N, C]
  float* dW; float* dbias;     // accumulated
  __nv_bfloat16* dlatent;      // [N*T*L, D] gradient w.r.t. the final-norm output
};
int launch_head_fwd(const HeadArgs& a, cudaStream_t stream);
int launch_head_bwd(const HeadArgs& a, cudaStream_t stream);

// ---- parameter packing ---------------------------------------------------------
struct PackJob {
  const float* src;
  int64_t dst_off;   // element offset in the bf16 arena (or fp32 arena for kind 0)
  int rows, cols;    // source shape
  int pitch;         // destination row pitch (elements)
  int kind;          // 0: fp32 copy, 1: bf16, 2: bf16 transposed
  int row_map;       // 0: plain, 1: w1-interleave (which=0), 2: w3-interleave (which=1)
  int row_off;       // added to the (mapped) row
  int tile0;         // index of this job's first [32 x 64] tile in the launch
  int tiles_c;       // tiles per tile row
};
constexpr int kPackTileR = 32, kPackTileC = 64;
// the device table is [njobs] PackJob followed by [ntiles] int32 tile -> job
int launch_pack(const PackJob* jobs_dev, int njobs, int ntiles, __nv_bfloat16* bf16_arena, float* f32_arena, cudaStream_t stream);

}  // namespace hsimae
