/* hsimae_b200 -- C ABI of the B200 (sm_100a) HSIMAE compute library.
 *
 * Drop-in boundary: the reference (Ryan21wy/HSIMAE) has no FFI/plugin layer; its
 * only stable interface is the Python module `Models` (HSIMAE, DualViT, HSIViT;
 * /root/reference/Models.py:309-634, 637-993, 996-1160).  The repo-root
 * `Models.py` mirrors that surface and drives the functions below through
 * ctypes with raw device pointers, explicit sizes and an explicit cudaStream_t.
 * No torch types cross this boundary.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero otherwise; the message is
 *     available (thread-local) from hsimae_last_error().
 *   - the caller owns every buffer (parameters, arenas, workspaces, outputs);
 *     the library allocates no device memory and keeps only cached TMA
 *     descriptors keyed by pointer/shape.
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no
 *     function synchronises, so the call sequence is CUDA-graph capturable.
 *   - all matrices are row-major; "bf16" is __nv_bfloat16; token rows of a
 *     sample are ordered (spectral group t, spatial position l), l fastest.
 */
#ifndef HSIMAE_B200_H_
#define HSIMAE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HSIMAE_ABI_VERSION 2

/* Constructor arguments of the reference model classes that shape the compute
 * path (Models.py:312-332, 640-663, 997-1016). */
typedef struct hsimae_dims {
  int32_t img_size, patch_size, bands, b_patch_size;
  int32_t embed_dim, depth, s_depth, num_heads;
  int32_t dec_dim, dec_depth, dec_heads; /* dec_dim == 0: no decoder (HSIViT) */
  int32_t num_class;                     /* 0: no classification head (HSIMAE) */
  int32_t qkv_bias;                      /* !no_qkv_bias */
  int32_t norm_pix_loss;
  float mlp_ratio;
} hsimae_dims;

typedef struct hsimae_plan hsimae_plan;

const char* hsimae_last_error(void);
int hsimae_abi_version(void);
/* number of kernels launched by this library so far in this process */
int64_t hsimae_launch_count(void);

/* ---- plan: static layout of parameters, packed operands and gradients ---- */
int hsimae_plan_create(const hsimae_dims* dims, hsimae_plan** out);
void hsimae_plan_destroy(hsimae_plan* plan);

/* Parameter table, in the library's canonical order.  Names are the reference's
 * state_dict keys (Models.py:342-424, 736).  `has_grad` is 0 for the frozen
 * position tables (Models.py:434-435). */
int hsimae_plan_num_params(const hsimae_plan* plan);
const char* hsimae_plan_param_name(const hsimae_plan* plan, int i);
int64_t hsimae_plan_param_numel(const hsimae_plan* plan, int i);
int64_t hsimae_plan_param_grad_offset(const hsimae_plan* plan, int i); /* fp32 elements into the gradient arena; -1 if none */
int hsimae_plan_param_has_grad(const hsimae_plan* plan, int i);
int64_t hsimae_plan_grad_arena_elems(const hsimae_plan* plan);  /* fp32 elements */
int64_t hsimae_plan_bf16_arena_elems(const hsimae_plan* plan);  /* packed GEMM operands */
int64_t hsimae_plan_f32_arena_elems(const hsimae_plan* plan);   /* packed biases / affine / tables */
int64_t hsimae_plan_pack_table_bytes(const hsimae_plan* plan);  /* device scratch for the packing job table */
int hsimae_plan_hidden(const hsimae_plan* plan, int decoder);   /* SwiGLU hidden width (Models.py:225) */
/* Gradient arena regions in the order their gradients are final during backward:
 * i = 3 decoder (after hsimae_decoder_backward), 2 fusion+norm+head (encoder stage 1),
 * 1 spectral encoder (stage 2), 0 patch embedding + spatial encoder (stage 4).
 * Lets the caller all-reduce one region while the next stage is still computing. */
int hsimae_plan_grad_bucket(const hsimae_plan* plan, int i, int64_t* offset, int64_t* elems);

/* fp32 master parameters -> packed bf16/fp32 arenas (both zero-initialised by
 * the caller once).  `params` is a HOST array of num_params device pointers. */
int hsimae_pack_params(hsimae_plan* plan, const void* const* params, void* bf16_arena, void* f32_arena,
                       void* pack_table_dev, void* stream);

/* ---- masking: replaces HSIMAE.spatial_spectral_masking (Models.py:495-535) ---- */
int hsimae_mask(const float* noise_t, const float* noise_l, int32_t n, int32_t T, int32_t L, int32_t len_t, int32_t len_l,
                int64_t* ids_keep, int64_t* ids_restore, float* mask, int32_t* ids_keep32, int32_t* ids_restore32,
                void* stream);

/* ---- encoder: replaces forward_encoder (Models.py:537-571, 869-894, 896-923, 1119-1145)
 * ids_keep32 == NULL selects the unmasked pass (len_t = T, len_l = L).
 * `drop` is NULL or a host array of 2*(2*s_depth+n_fusion) device pointers to the
 * already-scaled stochastic-depth factors, ordered blocks_1[i].{attn,mlp}...,
 * blocks_2[i]..., blocks[i]...; NULL entries mean identity (Models.py:235-251). */
int64_t hsimae_encoder_workspace_bytes(const hsimae_plan* plan, int32_t n, int32_t len_t, int32_t len_l, int32_t save);
int hsimae_encoder_forward(hsimae_plan* plan, const void* bf16_arena, const void* f32_arena, const float* imgs, int32_t n,
                           int32_t len_t, int32_t len_l, const int32_t* ids_keep32, const float* const* drop, int32_t save,
                           void* ws, int64_t ws_bytes, void* stream);
/* Unmasked inference pass over sliding windows of an HWC scene that is resident in HBM: sample i is the
 * img x img window whose top-left corner is window (pixel0 + i), windows enumerated row-major over
 * (scene_h - img + 1) x (scene_w - img + 1).  Replaces the host-side per-pixel cube materialisation of
 * Utils/Preprocessing.py:205-213 + the batch loop of Model_Finetuning.py:264-278 (workspace: save = 0 layout). */
int hsimae_encoder_forward_scene(hsimae_plan* plan, const void* bf16_arena, const void* f32_arena, const float* scene,
                                 int32_t scene_h, int32_t scene_w, int64_t pixel0, int32_t n, void* ws, int64_t ws_bytes,
                                 void* stream);
/* consumes the latent gradient left in the workspace by decoder/head backward.
 * `stages` is a bit mask (1: final norm + fusion blocks, 2: spectral encoder, 4 / 8 / 16: last / middle / first third of
 * the spatial encoder, 16 also the patch embedding); stages must be run in that order, HSIMAE_STAGES_ALL = all. */
#define HSIMAE_STAGES_ALL 31
/* Stage bit 32 (with bit 2): run the spectral chain on the library's helper stream beside the spatial chain (forked from `stream`
 * with an event, own scratch); it is joined back into `stream` before the patch-embedding backward of stage 16, or earlier by
 * hsimae_helper_join (data-parallel callers: the spectral gradient region is final on `stream` after the join). */
#define HSIMAE_STAGE_SPECTRAL_ASYNC 32
int hsimae_helper_join(hsimae_plan* plan, void* stream);
int hsimae_encoder_backward(hsimae_plan* plan, const void* bf16_arena, const void* f32_arena, const float* imgs, int32_t n,
                            int32_t len_t, int32_t len_l, const int32_t* ids_keep32, const float* const* drop, void* ws,
                            int64_t ws_bytes, float* grad_arena, int32_t stages, void* stream);
/* copy of the final-norm output (fp32 [n*K, D]) for inspection / parity tests */
int hsimae_encoder_latent(const hsimae_plan* plan, int32_t n, int32_t len_t, int32_t len_l, int32_t save, const void* f32_arena,
                          const void* ws, float* out, void* stream);

/* ---- decoder + loss + pixel outputs: forward_decoder, forward_loss, recons (Models.py:573-625) ---- */
int64_t hsimae_decoder_workspace_bytes(const hsimae_plan* plan, int32_t n, int32_t len_t, int32_t len_l, int32_t save);
int hsimae_decoder_forward(hsimae_plan* plan, const void* bf16_arena, const void* f32_arena, const float* imgs, int32_t n,
                           int32_t len_t, int32_t len_l, const int32_t* ids_restore32, const float* mask, const void* enc_ws,
                           int32_t enc_save, int32_t save, void* ws, int64_t ws_bytes, float* loss, float* pred_img,
                           float* mask_img, float* pred_tokens, void* stream);
/* grad_loss: device pointer to the scalar dL/dloss.  Leaves dL/dlatent in enc_ws. */
int hsimae_decoder_backward(hsimae_plan* plan, const void* bf16_arena, const void* f32_arena, int32_t n, int32_t len_t,
                            int32_t len_l, const int32_t* ids_restore32, void* enc_ws, void* ws, int64_t ws_bytes,
                            const float* grad_loss, float* grad_arena, void* stream);

/* ---- classification head: DualViT.head / HSIViT.head 'AGG' (Models.py:964-973, 1147-1156) ---- */
int hsimae_head_forward(hsimae_plan* plan, const void* f32_arena, int32_t n, const void* enc_ws, int32_t enc_save, float* pooled,
                        float* logits, void* stream);
int hsimae_head_backward(hsimae_plan* plan, const void* f32_arena, int32_t n, void* enc_ws, const float* pooled,
                         const float* dlogits, float* grad_arena, void* stream);

/* ---- single operators (unit parity tests; also the building blocks above) ---- */
/* C[M,N] = A[M,K] * B[N,K]^T with a fused epilogue (see csrc/gemm.cuh).  impl 0 = tcgen05, 1 = CUDA-core checker. */
typedef struct hsimae_gemm_desc {
  int32_t M, N, K, epilogue, impl;
  const void* A; int32_t lda;
  const void* B; int32_t ldb;
  void* out0; int32_t ld0;
  void* out1; int32_t ld1;
  const float* bias;
  const float* resid; int32_t ldr;
  const float* resid2;
  const float* gamma; const float* beta; float* stats;
  const void* ab; int32_t ldab;
  const float* rowscale; int32_t rs_mode, rs_K, rs_len_l, rs_G;
  float* scratch; /* impl 1: fp32 [M,N] */
  const void* A2; int32_t lda2; /* epilogue 5: bf16 [M,K] input of the gated projection */
  const void* B2; int32_t ldb2; /* epilogue 5: bf16 [2N,K] interleaved w1|w3; `bias` = its packed bias */
} hsimae_gemm_desc;
int hsimae_gemm(const hsimae_gemm_desc* d, void* stream);

/* dgrad GEMM + LayerNorm backward in ONE launch (csrc/gemm.cuh kEpiLnBwd; replaces the autograd of nn.LayerNorm behind
 * Models.py:304-305 together with the nn.Linear input gradient that feeds it):
 *   dy = A[M,K] * B[N,K]^T;  dx_out = dx_in + LN_bwd(dy; x, stats = (mean, rstd) per row, gamma);
 *   dxb (optional) = bf16(rowscale * dx_out);  dgamma += sum_rows dy * xhat;  dbeta += sum_rows dy.
 * 32 <= N <= 256, N % 32 == 0 (the row has to fit one accumulator tile); dx_in == dx_out is allowed.  The engine uses it from
 * two waves of 128-row tiles up (HSIMAE_LNBWD_MIN_ROWS), below that the dgrad GEMM and the LayerNorm backward are two launches. */
typedef struct hsimae_lnbwd_desc {
  int32_t M, N, K;
  const void* A; int32_t lda;
  const void* B; int32_t ldb;
  const float* x; int32_t ldx;
  const float* stats;
  const float* gamma;
  const float* dx_in; int32_t ldi;
  float* dx_out; int32_t ldo;
  void* dxb; int32_t ldxb;
  const float* rowscale; int32_t rs_mode, rs_K, rs_len_l, rs_G;
  float* dgamma; float* dbeta;
} hsimae_lnbwd_desc;
int hsimae_gemm_lnbwd(const hsimae_lnbwd_desc* d, void* stream);

/* Programmatic dependent launch of the library's kernels on (default, HSIMAE_PDL=0 disables) / off; returns the previous
 * setting.  Off is for per-kernel timing: overlapped prologues make consecutive kernels' profiler durations overlap. */
int hsimae_set_pdl(int on);

/* W[Nout,Kin] += Y[Mred,Nout]^T X[Mred,Kin] (+ column sums of Y into bias).  row_map 1 = interleaved w1|w3 rows. */
typedef struct hsimae_wgrad_desc {
  int32_t Mred, Nout, Kin, impl;
  const void* Y; int32_t ldy;
  const void* X; int32_t ldx;
  float* dst0; float* dst1; int32_t ld, row_map, rows_valid, cols_valid;
  float* bias0; float* bias1;
} hsimae_wgrad_desc;
int hsimae_wgrad(const hsimae_wgrad_desc* d, void* stream);
/* up to 4 independent problems (the weight gradients of one block) in ONE launch, one wave over the SMs */
int hsimae_wgrad_group(const hsimae_wgrad_desc* d, int32_t n, void* stream);

/* Gated MLP half of a block as ONE kernel (csrc/block_fused.cu; replaces the kEpiSwiGLU + kEpiResidLN pair of launches;
 * /root/reference/Models.py:231-232 SwiGLU.forward + the residual add of Block.forward :305):
 *   out = resid + rs * (W2 (silu(W1 x) * (W3 x)) + b2) [+ resid2];  ln = LayerNorm(out) * gamma + beta;  stats = (mean, rstd)
 * X bf16 [M, D]; W13 bf16 [2 Hp, D] (w1|w3 interleaved by 16 rows, zero rows for the padding), b13 fp32 [2 Hp];
 * W2 bf16 [D, Hp]; g (optional) receives the bf16 gate output [M, Hp] backward needs.  D in {64,128,192,256}, Hp % 16 == 0. */
typedef struct hsimae_mlp_desc {
  int32_t M, D, Hp;
  const void* X; int32_t ldx;
  const void* W13; int32_t ldw13;
  const float* b13;
  const void* W2; int32_t ldw2;
  const float* b2;
  const float* resid; int32_t ldr;
  const float* resid2;
  const float* rowscale; int32_t rs_mode, rs_K, rs_len_l, rs_G;
  float* out; int32_t ldo;
  const float* gamma; const float* beta; void* ln; int32_t ldln; float* stats;
  void* g; int32_t ldg;
} hsimae_mlp_desc;
int hsimae_mlp_fused(const hsimae_mlp_desc* d, void* stream);

int hsimae_attention_forward(const void* qkv, void* out, float* lse, int32_t n, int32_t D, int32_t heads, int32_t K, int32_t nseq,
                             int32_t len, int32_t seq_step, int32_t tok_step, void* stream);
int hsimae_attention_backward(const void* qkv, const void* out, const float* lse, const void* dout, void* dqkv, int32_t n,
                              int32_t D, int32_t heads, int32_t K, int32_t nseq, int32_t len, int32_t seq_step,
                              int32_t tok_step, void* stream);

/* ---- on-device pretraining data feed -------------------------------------------------------------
 * Replaces HSIdataset4PT.__getitem__ + DataLoader collation (/root/reference/Model_Pretraining.py:40-51, :76):
 *   out[i, 0, c, y, x] = (scene[num][h + y', w + x', c] - min) / (max - min),  (c_, h, w, num, max, min) = cut_info[index[i]]
 * with y' = img-1-y when flips[2i+1] (np.flip(data, 0)) and x' = img-1-x when flips[2i] (np.flip(data, 1)).
 * scenes: all scenes back to back, each [H, W, bands] fp32; scene_off[s] = element offset of scene s, scene_hw[2s..] = (H, W);
 * cut_info: int16 rows of 6 (Utils/Preprocessing.py:78,114); index: int64 [n]; flips: uint8 [n, 2] or NULL (no flips);
 * out: fp32 [n, 1, bands, img, img].  Windows must lie inside their scene (the host wrapper checks once). */
int hsimae_gather_patches(const float* scenes, const int64_t* scene_off, const int32_t* scene_hw, int32_t bands, int32_t img,
                          const int16_t* cut_info, const int64_t* index, const uint8_t* flips, int32_t n, float* out, void* stream);

/* ---- fused multi-tensor AdamW ---------------------------------------------------------------------
 * Replaces optimizer.step() of torch.optim.AdamW as the drivers configure it (/root/reference/Model_Pretraining.py:80-86,
 * 102; Model_Finetuning.py:98-104,165): one launch over a device job table
 *   struct { float* p; const float* g; float* m; float* v; int32 n; float reserved; int32 tile0; int32 pad; } jobs[njobs];
 *   int32 tile_job[ntiles];           tiles of hsimae_adamw_tile_elems() elements, tile0 = first tile of the job
 * (the table depends on pointers only: it survives learning-rate changes; decay = 1 - lr*wd is a launch scalar)
 * p *= decay; m += (1-beta1)(g-m); v = v*beta2 + (1-beta2) g*g; p += neg_step_size * m / (sqrt(v)/bc2_sqrt + eps)
 * with neg_step_size = -lr / (1 - beta1^t), bc2_sqrt = sqrt(1 - beta2^t): the operation order of torch's foreach
 * implementation, scalars evaluated in double on the host as torch does. */
int32_t hsimae_adamw_tile_elems(void);
int hsimae_adamw_step(const void* jobs, int32_t njobs, int32_t ntiles, float decay, float one_minus_beta1, float beta2,
                      float one_minus_beta2, float eps, float bias_correction2_sqrt, float neg_step_size, void* stream);

/* ---- group-wise PCA preprocessing -------------------------------------------------------------------
 * Replaces applyGWPCA (/root/reference/Utils/GroupWisePCA.py:20-34; callers Utils/Preprocessing.py:90-91,192-193):
 * global min-max scaling, contiguous band groups (split_data, :5-17), one sklearn PCA(n_components, whiten) per group.
 * The device makes the passes over the pixels, the caller solves the b x b symmetric eigenproblems in between:
 *   hsimae_gwpca_moments: X [n, c] row-major pixels (dtype 0 = float32, 1 = float64), raw (un-normalised) values ->
 *     mean [c], minmax [2] = (global min, global max), cov [ngroups, 64, 64] = centred covariance of every group's bands
 *     (divisor n - 1, rows / columns beyond the group's width are zero); all fp64, all on the device.
 *   hsimae_gwpca_project: out [n, ngroups * k_per_group] fp64, out[r, g * k + i] = sum_j (X[r, off_g + j] - mean[off_g + j]) * W[g * k + i, j]
 *     with W [ngroups * k, 64] fp64 on the device (component i of group g times the caller's scale: 1 / sqrt(eigenvalue)
 *     when whitening, 1 / (max - min) otherwise).  flip_u != 0 additionally negates every output column whose entry of
 *     largest magnitude (first on ties) is negative -- sklearn <= 1.4's svd_flip(u_based_decision=True), the reference's pin.
 * group_off: HOST array of ngroups + 1 band offsets (groups of 1..64 bands, at most 16 groups, at most 64 output channels).
 * ws: device scratch of hsimae_gwpca_workspace_bytes(c, ngroups, nout) bytes (need not be zeroed).  Every reduction has a
 * fixed order: results are deterministic.  Nothing synchronises. */
int64_t hsimae_gwpca_workspace_bytes(int32_t c, int32_t ngroups, int32_t nout);
int hsimae_gwpca_moments(const void* X, int32_t dtype, int64_t n, int32_t c, int32_t ngroups, const int32_t* group_off, void* ws,
                         int64_t ws_bytes, double* mean, double* minmax, double* cov, void* stream);
int hsimae_gwpca_project(const void* X, int32_t dtype, int64_t n, int32_t c, int32_t ngroups, const int32_t* group_off,
                         int32_t k_per_group, const double* mean, const double* W, double* out, int32_t flip_u, void* ws,
                         int64_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HSIMAE_B200_H_ */
