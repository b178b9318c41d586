// Decoder token fill/unshuffle, reconstruction loss + pixel outputs,
// classification head, parameter packing.
#include "kernels.cuh"

namespace hsimae {

// ---------------------------------------------------------------------------
// Decoder fill.  Reference: Models.py:583-592 -- masked slots receive the mean
// of the sample's visible decoder-embedded tokens (the learned mask_token is
// never read), tokens are unshuffled with ids_restore, decoder_pos_embed is
// added; fused with the first decoder LayerNorm (Models.py:304).
// One CTA per sample, one warp per output token.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
fill_fwd_kernel(FillArgs a) {
  extern __shared__ float sm[];
  const int D = a.D, K = a.K, P = a.P;
  float* sy = sm;                 // [K][D]
  float* smean = sy + (size_t)K * D;  // [D]
  int* ssrc = reinterpret_cast<int*>(smean + D);   // [P] ids_restore of the sample: staged with the tokens, so that the
                                                   // per-token loop below has no dependent global load in it
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int vecs = D / 4;   // D % 4 == 0 (decoder widths are multiples of 16)
  for (int n = blockIdx.x; n < a.N; n += gridDim.x) {
    __syncthreads();
    const float4* src4 = reinterpret_cast<const float4*>(a.y + (size_t)n * K * D);
    for (int i = threadIdx.x; i < K * vecs; i += blockDim.x) reinterpret_cast<float4*>(sy)[i] = ld_stream_f4(src4 + i);
    for (int p = threadIdx.x; p < P; p += blockDim.x) ssrc[p] = a.ids_restore[(size_t)n * P + p];
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s += sy[k * D + d];
      smean[d] = s / K;
    }
    __syncthreads();
    for (int p = warp; p < P; p += nw) {
      const int src = ssrc[p];
      const float* row = src < K ? sy + (size_t)src * D : smean;
      const size_t m = (size_t)n * P + p;
      float v[8];
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = lane + 32 * j;
        v[j] = i < D ? row[i] + __ldg(a.pos + (size_t)p * D + i) : 0.f;
        s += v[j];
      }
      const float mean = warp_sum(s) / D;
      float sq = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) { const int i = lane + 32 * j; if (i < D) { const float dv = v[j] - mean; sq = fmaf(dv, dv, sq); } }
      const float rstd = rsqrtf(warp_sum(sq) / D + a.eps);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = lane + 32 * j;
        if (i < D) {
          a.x[m * D + i] = v[j];
          if (a.ln) a.ln[m * D + i] = __float2bfloat16_rn(fmaf((v[j] - mean) * rstd, a.gamma[i], a.beta[i]));
        }
      }
      if (lane == 0 && a.stats) { a.stats[2 * m] = mean; a.stats[2 * m + 1] = rstd; }
    }
  }
}

// dy[k] = dx[pos(k)] + (1/K) * sum_{masked p} dx[p]
__global__ void __launch_bounds__(256)
fill_bwd_kernel(FillArgs a) {
  extern __shared__ float sm[];
  const int D = a.D, K = a.K, P = a.P;
  float* sdx = sm;                    // [P][D]
  float* sacc = sdx + (size_t)P * D;  // [D]
  int* spos = reinterpret_cast<int*>(sacc + D);  // [K]
  int* ssrc = spos + K;                          // [P] ids_restore of the sample (no global loads inside the reductions)
  for (int n = blockIdx.x; n < a.N; n += gridDim.x) {
    __syncthreads();
    {
      const float4* src4 = reinterpret_cast<const float4*>(a.dx + (size_t)n * P * D);
      for (int i = threadIdx.x; i < P * D / 4; i += blockDim.x) reinterpret_cast<float4*>(sdx)[i] = ld_stream_f4(src4 + i);
    }
    for (int p = threadIdx.x; p < P; p += blockDim.x) {
      const int src = a.ids_restore[(size_t)n * P + p];
      ssrc[p] = src;
      if (src < K) spos[src] = p;
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      float s = 0.f;
      for (int p = 0; p < P; ++p)
        if (ssrc[p] >= K) s += sdx[p * D + d];
      sacc[d] = s / K;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * D; i += blockDim.x) {
      const int k = i / D, d = i - k * D;
      a.dy[(size_t)n * K * D + i] = __float2bfloat16_rn(sdx[spos[k] * D + d] + sacc[d]);
    }
  }
}

int launch_fill_fwd(const FillArgs& a, cudaStream_t stream) {
  HS_REQUIRE(a.D <= 256, "decoder fill: D=%d > 256 unsupported", a.D);
  if (a.N == 0) return kOk;
  HS_REQUIRE(a.D % 4 == 0, "decoder fill: D=%d must be a multiple of 4", a.D);
  const size_t smem = ((size_t)a.K * a.D + a.D) * sizeof(float) + (size_t)a.P * sizeof(int);
  HS_CHECK_CUDA(cudaFuncSetAttribute(fill_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = a.N < 8 * kNumSMs ? a.N : 8 * kNumSMs;
  fill_fwd_kernel<<<grid, 256, smem, stream>>>(a);
  HS_CHECK_LAUNCH("fill_fwd_kernel");
  return kOk;
}

int launch_fill_bwd(const FillArgs& a, cudaStream_t stream) {
  if (a.N == 0) return kOk;
  const size_t smem = ((size_t)a.P * a.D + a.D) * sizeof(float) + (size_t)(a.K + a.P) * sizeof(int);
  HS_CHECK_CUDA(cudaFuncSetAttribute(fill_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = a.N < 8 * kNumSMs ? a.N : 8 * kNumSMs;
  fill_bwd_kernel<<<grid, 256, smem, stream>>>(a);
  HS_CHECK_LAUNCH("fill_bwd_kernel");
  return kOk;
}

// ---------------------------------------------------------------------------
// Reconstruction loss + pixel-space outputs.  Reference: Models.py:603-625.
//   target = patchify(imgs); (norm_pix) target = (target-mean)/sqrt(var_unbiased+1e-6)
//   loss   = sum_p mask_p * mean_j (pred-target)^2 / sum(mask)
//   pred_img = unpatchify(pred*std+mean), mask_img = unpatchify(mask broadcast)
// Also emits the unit-scale gradient dL/dpred (bf16) for the backward pass.
// One CTA per sample (grid-stride), one warp per patch token.
// ---------------------------------------------------------------------------
__device__ __forceinline__ int cube_index2(const PatchGeom& g, int p, int j) {
  const int t = p / g.L, hw = p - t * g.L;
  const int h = hw / g.G, w = hw - h * g.G;
  const int pp2 = g.p * g.p;
  const int u = j / pp2, r = j - u * pp2;
  const int pp = r / g.p, q = r - pp * g.p;
  return ((t * g.u + u) * g.img + (h * g.p + pp)) * g.img + (w * g.p + q);
}

__global__ void __launch_bounds__(256)
loss_kernel(LossArgs a) {
  extern __shared__ float sm[];
  const PatchGeom g = a.g;
  float* sCube = sm;                  // input cube
  float* sPred = sCube + g.cube;      // de-normalised prediction, cube layout
  float* sMask = sPred + g.cube;      // mask, cube layout
  // cube offset of element j of patch p, built once per CTA: the (t, h, w, u, p, q) index arithmetic is five runtime
  // integer divisions per element, which was most of this kernel's instruction stream (139 -> see profiles/r02*)
  unsigned short* sIdx = reinterpret_cast<unsigned short*>(sMask + g.cube);   // [P][PK]
  __shared__ float swl[8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  const int PK = g.PK;
  const float inv_pk = 1.0f / PK;
  for (int i = threadIdx.x; i < g.P * PK; i += blockDim.x) {
    const int p = i / PK, j = i - p * PK;
    sIdx[i] = (unsigned short)cube_index2(g, p, j);
  }
  for (int n = blockIdx.x; n < a.N; n += gridDim.x) {
    __syncthreads();
    const float4* src = reinterpret_cast<const float4*>(a.imgs + (size_t)n * g.cube);
    for (int i = threadIdx.x; i < g.cube / 4; i += blockDim.x) reinterpret_cast<float4*>(sCube)[i] = ld_stream_f4(src + i);
    __syncthreads();
    float wl = 0.f;
    for (int p = warp; p < g.P; p += nw) {
      const size_t m = (size_t)n * g.P + p;
      const float mk = a.mask[m];
      float t[4], pr[4]; int idx[4];
      float s = 0.f;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int j = lane + 32 * r;
        if (j < PK) { idx[r] = sIdx[p * PK + j]; t[r] = sCube[idx[r]]; s += t[r]; pr[r] = a.pred[m * a.ldp + j]; } else { idx[r] = 0; t[r] = 0.f; pr[r] = 0.f; }
      }
      float mean = 0.f, std = 1.f;
      if (a.norm_pix) {
        mean = warp_sum(s) * inv_pk;
        float sq = 0.f;
#pragma unroll
        for (int r = 0; r < 4; ++r) { const int j = lane + 32 * r; if (j < PK) { const float dv = t[r] - mean; sq = fmaf(dv, dv, sq); } }
        const float var = warp_sum(sq) / (PK - 1);   // torch.var default: unbiased
        std = sqrtf(var + 1.0e-6f);
      }
      const float istd = 1.0f / std;
      float l = 0.f;
      const float gsc = mk * 2.0f * inv_pk / a.mask_sum;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int j = lane + 32 * r;
        if (j < PK) {
          const float tn = (t[r] - mean) * istd;
          const float df = pr[r] - tn;
          l = fmaf(df, df, l);
          if (a.dpred) a.dpred[m * a.ldd + j] = __float2bfloat16_rn(gsc * df);
          sPred[idx[r]] = fmaf(pr[r], std, mean);
          sMask[idx[r]] = mk;
        } else if (j < a.ldd && a.dpred) {
          a.dpred[m * a.ldd + j] = __float2bfloat16_rn(0.f);
        }
      }
      l = warp_sum(l) * inv_pk;
      wl += mk * l;
    }
    if (lane == 0) swl[warp] = wl;
    __syncthreads();
    if (threadIdx.x == 0) {
      float tot = 0.f;
      for (int w = 0; w < nw; ++w) tot += swl[w];
      a.loss_partial[n] = tot;
    }
    if (a.pred_img) {
      float4* dst = reinterpret_cast<float4*>(a.pred_img + (size_t)n * g.cube);
      for (int i = threadIdx.x; i < g.cube / 4; i += blockDim.x) dst[i] = reinterpret_cast<float4*>(sPred)[i];
    }
    if (a.mask_img) {
      float4* dst = reinterpret_cast<float4*>(a.mask_img + (size_t)n * g.cube);
      for (int i = threadIdx.x; i < g.cube / 4; i += blockDim.x) dst[i] = reinterpret_cast<float4*>(sMask)[i];
    }
  }
}

// deterministic final reduction of the per-sample partials
__global__ void __launch_bounds__(1024)
loss_reduce_kernel(const float* __restrict__ partial, int N, float denom, float* __restrict__ loss) {
  __shared__ double sh[1024];
  double s = 0.0;
  for (int i = threadIdx.x; i < N; i += 1024) s += (double)partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss = (float)(sh[0] / (double)denom);
}

int launch_loss(const LossArgs& a, cudaStream_t stream) {
  HS_REQUIRE(a.g.PK <= 128, "loss: patch of %d elements unsupported (max 128)", a.g.PK);
  HS_REQUIRE(a.g.cube % 4 == 0, "loss: cube size must be a multiple of 4");
  HS_REQUIRE(a.mask_sum > 0.f, "loss: no masked patches");
  if (a.N == 0) return kOk;
  HS_REQUIRE(a.g.cube < 65536, "loss: cube of %d elements unsupported", a.g.cube);
  const size_t smem = (size_t)3 * a.g.cube * sizeof(float) + (size_t)a.g.P * a.g.PK * sizeof(unsigned short);
  HS_CHECK_CUDA(cudaFuncSetAttribute(loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = a.N < 8 * kNumSMs ? a.N : 8 * kNumSMs;
  loss_kernel<<<grid, 256, smem, stream>>>(a);
  HS_CHECK_LAUNCH("loss_kernel");
  loss_reduce_kernel<<<1, 1024, 0, stream>>>(a.loss_partial, a.N, a.mask_sum, a.loss);
  HS_CHECK_LAUNCH("loss_reduce_kernel");
  return kOk;
}

// ---------------------------------------------------------------------------
// Classification head.  Reference: Models.py:964-973 ('AGG'): latent
// [N,T,L,C] -> [N,L,T*C] -> mean over L -> Linear(T*C -> classes).  The final
// encoder LayerNorm (Models.py:893) is applied here from the saved row stats.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_fwd_kernel(HeadArgs a) {
  extern __shared__ float sz[];  // [T*D]
  const int TD = a.T * a.D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int n = blockIdx.x; n < a.N; n += gridDim.x) {
    __syncthreads();
    for (int i = threadIdx.x; i < TD; i += blockDim.x) {
      const int t = i / a.D, d = i - t * a.D;
      float s = 0.f;
      for (int l = 0; l < a.L; ++l) {
        const size_t m = ((size_t)n * a.T + t) * a.L + l;
        s += fmaf((a.x[m * a.D + d] - a.stats[2 * m]) * a.stats[2 * m + 1], a.gamma[d], a.beta[d]);
      }
      s /= a.L;
      sz[i] = s;
      a.z[(size_t)n * TD + i] = s;
    }
    __syncthreads();
    for (int c = warp; c < a.C; c += nw) {
      float s = 0.f;
      for (int i = lane; i < TD; i += 32) s = fmaf(sz[i], __ldg(a.W + (size_t)c * TD + i), s);
      s = warp_sum(s);
      if (lane == 0) a.logits[(size_t)n * a.C + c] = s + a.bias[c];
    }
  }
}

// dlatent[n,t,l,d] = (1/L) sum_c dlogits[n,c] W[c, t*D+d]
__global__ void __launch_bounds__(256)
head_bwd_dx_kernel(HeadArgs a) {
  const int TD = a.T * a.D;
  for (int n = blockIdx.x; n < a.N; n += gridDim.x) {
    for (int i = threadIdx.x; i < TD; i += blockDim.x) {
      float s = 0.f;
      for (int c = 0; c < a.C; ++c) s = fmaf(a.dlogits[(size_t)n * a.C + c], __ldg(a.W + (size_t)c * TD + i), s);
      s /= a.L;
      const int t = i / a.D, d = i - t * a.D;
      const __nv_bfloat16 v = __float2bfloat16_rn(s);
      for (int l = 0; l < a.L; ++l) a.dlatent[(((size_t)n * a.T + t) * a.L + l) * a.D + d] = v;
    }
  }
}

// dW[c,i] += sum_n dlogits[n,c] z[n,i];  dbias[c] += sum_n dlogits[n,c]   (one CTA per class: deterministic)
__global__ void __launch_bounds__(256)
head_bwd_dw_kernel(HeadArgs a) {
  const int TD = a.T * a.D;
  const int c = blockIdx.x;
  for (int i = threadIdx.x; i < TD; i += blockDim.x) {
    float s = 0.f;
    for (int n = 0; n < a.N; ++n) s = fmaf(a.dlogits[(size_t)n * a.C + c], a.z[(size_t)n * TD + i], s);
    a.dW[(size_t)c * TD + i] += s;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int n = 0; n < a.N; ++n) s += a.dlogits[(size_t)n * a.C + c];
    a.dbias[c] += s;
  }
}

int launch_head_fwd(const HeadArgs& a, cudaStream_t stream) {
  if (a.N == 0) return kOk;
  const size_t smem = (size_t)a.T * a.D * sizeof(float);
  HS_REQUIRE(smem <= 200 * 1024, "head: T*D too large");
  HS_CHECK_CUDA(cudaFuncSetAttribute(head_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = a.N < 8 * kNumSMs ? a.N : 8 * kNumSMs;
  head_fwd_kernel<<<grid, 256, smem, stream>>>(a);
  HS_CHECK_LAUNCH("head_fwd_kernel");
  return kOk;
}

int launch_head_bwd(const HeadArgs& a, cudaStream_t stream) {
  if (a.N == 0) return kOk;
  int grid = a.N < 8 * kNumSMs ? a.N : 8 * kNumSMs;
  head_bwd_dx_kernel<<<grid, 256, 0, stream>>>(a);
  HS_CHECK_LAUNCH("head_bwd_dx_kernel");
  head_bwd_dw_kernel<<<a.C, 256, 0, stream>>>(a);
  HS_CHECK_LAUNCH("head_bwd_dw_kernel");
  return kOk;
}

// ---------------------------------------------------------------------------
// Parameter packing: fp32 master parameters -> bf16 GEMM operands (both
// orientations, fused q|k|v and interleaved w1|w3) + fp32 bias/affine arena.
// One launch over a device-resident job table.
// ---------------------------------------------------------------------------
// One CTA per [32 rows x 64 cols] tile of one job (1-D launch over all tiles; the tile -> job map follows the job
// table): coalesced fp32 reads, coalesced writes in either orientation (the transposed copy goes through shared
// memory so that the 32 destination elements of a source column leave as one 64-byte run).
__global__ void __launch_bounds__(256)
pack_kernel(const PackJob* __restrict__ jobs, const int* __restrict__ tile_job, __nv_bfloat16* __restrict__ wb, float* __restrict__ wf) {
  __shared__ float tile[kPackTileR][kPackTileC + 1];
  const PackJob j = jobs[tile_job[blockIdx.x]];
  const int t = blockIdx.x - j.tile0;
  const int r0 = (t / j.tiles_c) * kPackTileR, c0 = (t % j.tiles_c) * kPackTileC;
  auto dst_row = [&](int r) {
    int rm = r;
    if (j.row_map != 0) rm = (r / 16) * 32 + (j.row_map - 1) * 16 + (r % 16);
    return rm + j.row_off;
  };
  if (j.kind != 2) {
    for (int i = threadIdx.x; i < kPackTileR * kPackTileC; i += blockDim.x) {
      const int r = r0 + i / kPackTileC, c = c0 + i % kPackTileC;
      if (r < j.rows && c < j.cols) {
        const float v = j.src[(size_t)r * j.cols + c];
        if (j.kind == 0) wf[j.dst_off + (int64_t)dst_row(r) * j.pitch + c] = v;
        else wb[j.dst_off + (int64_t)dst_row(r) * j.pitch + c] = __float2bfloat16_rn(v);
      }
    }
    return;
  }
  for (int i = threadIdx.x; i < kPackTileR * kPackTileC; i += blockDim.x) {
    const int lr = i / kPackTileC, lc = i % kPackTileC;
    const int r = r0 + lr, c = c0 + lc;
    tile[lr][lc] = (r < j.rows && c < j.cols) ? j.src[(size_t)r * j.cols + c] : 0.f;
  }
  __syncthreads();
  // destination rows (mapped source rows) are the fast index of the transposed copy
  for (int i = threadIdx.x; i < kPackTileR * kPackTileC; i += blockDim.x) {
    const int lc = i / kPackTileR, lr = i % kPackTileR;
    const int r = r0 + lr, c = c0 + lc;
    if (r < j.rows && c < j.cols) wb[j.dst_off + (int64_t)c * j.pitch + dst_row(r)] = __float2bfloat16_rn(tile[lr][lc]);
  }
}

int launch_pack(const PackJob* jobs_dev, int njobs, int ntiles, __nv_bfloat16* bf16_arena, float* f32_arena, cudaStream_t stream) {
  if (njobs == 0 || ntiles == 0) return kOk;
  pack_kernel<<<ntiles, 256, 0, stream>>>(jobs_dev, reinterpret_cast<const int*>(jobs_dev + njobs), bf16_arena, f32_arena);
  HS_CHECK_LAUNCH("pack_kernel");
  return kOk;
}

}  // namespace hsimae
