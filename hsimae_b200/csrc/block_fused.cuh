// Fused transformer-block kernels (block_fused.cu): a CTA (pair) carries a 128-row token tile through several
// contractions of a block without the intermediate activations touching HBM.
#pragma once
#include "gemm.cuh"

namespace hsimae {

// Gated MLP half of a block (Models.py:232, 305): x_out = resid + rs * (w2(silu(w1 x) * w3 x) + b2) [+ resid2],
// followed by the LayerNorm of the next consumer -- the two GEMM launches of block_forward (kEpiSwiGLU, kEpiResidLN)
// as ONE kernel; the gate output only goes to HBM when backward needs it (g != nullptr).
struct MlpFusedArgs {
  GemmArgs tail;                       // the down-projection call: M, N = d, K = Hp, B = W2 [d, Hp], bias = b2, resid, rs,
                                       // resid2, out0 = x_out (fp32), gamma / beta / out1 / stats = next LayerNorm; A unused
  const __nv_bfloat16* X;   int ldx;   // [M, d]    input of the gated projection (LayerNorm-2 output)
  const __nv_bfloat16* W13; int ldw;   // [2 Hp, d] w1|w3 interleaved by 16 rows
  const float* b13;                    // [2 Hp]    packed bias
  __nv_bfloat16* g;         int ldg;   // [M, Hp]   gate output kept for backward, or nullptr
};

bool mlp_fused_supported(int d, int Hp);
int mlp_fused(const MlpFusedArgs& a, cudaStream_t stream);

}  // namespace hsimae
