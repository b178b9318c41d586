// GEMM interface shared by the tcgen05 kernels (gemm_tc.cu) and the SIMT
// checker kernels (gemm_simt.cu).  All matrices are row-major.
//
//   forward / dgrad :  C[M,N] = A[M,K] * B[N,K]^T        (A, B bf16, K contiguous)
//   wgrad           :  W[Nout,Kin] += Y[Mred,Nout]^T * X[Mred,Kin]   (fp32 red.add)
//
// The epilogues are written once (epilogue.cuh) against an accumulator
// "provider" so that the tensor-core kernel (accumulator rows in TMEM) and the
// checker (accumulator rows in a global fp32 scratch) share them.
#pragma once
#include "common.cuh"

namespace hsimae {

enum Epilogue : int {
  kEpiBiasBf16 = 0,  // out0(bf16) = acc + bias
  kEpiBiasF32 = 1,   // out0(f32)  = acc + bias
  kEpiResidLN = 2,   // out0(f32)  = resid + rs*(acc+bias) [+ resid2];  out1(bf16) = LN(out0); stats
  kEpiSwiGLU = 3,    // out0(bf16, interleaved a|b) = acc + bias;  out1(bf16) = silu(a)*b
  kEpiDSwiGLU = 4,   // acc = dg;  out0(bf16, interleaved) = (dg*b*silu'(a) | dg*silu(a)), a,b from `ab`
  kEpiDGate = 5,     // like kEpiDSwiGLU, but a|b = A2 * B2^T + bias is RECOMPUTED by a second GEMM in the same kernel
  kNumEpilogues = 6,
  // dgrad + LayerNorm backward in one kernel (internal to the engine, tensor-core path only; the tile spans the row):
  //   dy = acc;  out0(f32) = resid + LN_bwd(dy; lnx, stats, gamma);  out1(bf16, optional) = rs * out0;
  //   dgamma += sum_rows dy * xhat;  dbeta += sum_rows dy
  kEpiLnBwd = 6,
};

struct GemmArgs {
  int M, N, K;
  const __nv_bfloat16* A;  int lda;   // [M,K]
  const __nv_bfloat16* B;  int ldb;   // [N,K]
  void* out0;              int ld0;
  void* out1;              int ld1;
  const float* bias;                   // [N] or nullptr
  const float* resid;      int ldr;    // fp32 [M,N] (kEpiResidLN)
  const float* resid2;                 // optional second residual, same ld
  RowScale rs;                         // stochastic-depth factor on (acc+bias)
  const float* gamma;                  // LN affine (nullptr => no LN / no out1)
  const float* beta;
  float* stats;                        // [M,2] (mean, rstd) or nullptr
  const __nv_bfloat16* ab; int ldab;   // kEpiDSwiGLU: saved pre-activations (interleaved)
  const __nv_bfloat16* A2; int lda2;   // kEpiDGate: [M,K] input of the gated projection
  const __nv_bfloat16* B2; int ldb2;   // kEpiDGate: [2N,K] interleaved w1|w3 (bias = its packed bias)
  float ln_eps;
  const float* lnx;        int ldx;    // kEpiLnBwd: fp32 [M,N] input of the LayerNorm being differentiated (stats = its mean / rstd)
  float* dgamma;                       // kEpiLnBwd: [N] accumulated with atomics (nullptr => skipped)
  float* dbeta;
};

struct WgradArgs {
  int Mred, Nout, Kin;                 // reduction length, output rows, output cols
  const __nv_bfloat16* Y;  int ldy;    // [Mred, Nout]
  const __nv_bfloat16* X;  int ldx;    // [Mred, Kin]
  float* dst0;                         // fp32 [*, ld] accumulated with red.add
  float* dst1;                         // second destination for the interleaved map
  int ld;
  int row_map;                         // 0: row r -> dst0[r];  1: interleave16 (a-rows -> dst0, b-rows -> dst1)
  int rows_valid;                      // rows (after mapping) that exist in dst
  int cols_valid;                      // columns that exist in dst
  float* bias0;                        // optional column-sum of Y (bias grad), same row map
  float* bias1;
};

// interleave granularity of the fused w1|w3 projection: packed column p holds
// hidden unit (p/32)*16 + p%16 of w1 when (p%32) < 16, else of w3.
constexpr int kGate = 16;

__host__ __device__ inline int packed_col(int which, int h) { return (h / kGate) * (2 * kGate) + which * kGate + (h % kGate); }

int gemm_tc(const GemmArgs& a, int epi, cudaStream_t stream);
int gemm_simt(const GemmArgs& a, int epi, float* scratch, cudaStream_t stream);
int gemm_tc_lnbwd(const GemmArgs& a, cudaStream_t stream);   // kEpiLnBwd
bool gemm_lnbwd_supported(const GemmArgs& a);   // shapes the fused dgrad + LayerNorm-backward kernel takes
bool gemm_lnbwd_preferred(const GemmArgs& a);   // ... and where the engine uses it (else: two launches)
int wgrad_tc(const WgradArgs& a, cudaStream_t stream);
int wgrad_tc_group(const WgradArgs* jobs, int njobs, cudaStream_t stream);   // up to 4 independent problems in one launch
int wgrad_simt(const WgradArgs& a, cudaStream_t stream);

}  // namespace hsimae
