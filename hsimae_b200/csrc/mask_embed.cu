// Structured spatial-spectral masking and the fused patch-embedding kernels.
#include "kernels.cuh"

namespace hsimae {

// ---------------------------------------------------------------------------
// Masking.  Reference: Models.py:495-535 (three argsort/argsort/gather rounds).
// Restated per sample as: rank the T spectral noises and the L spatial noises
// (stable, lowest index first); a token (t,l) is visible iff rank_t < len_t and
// rank_l < len_l; ids_shuffle lists tokens by (#dropped axes, raster index).
// Integer outputs are bit-exact with the reference (tests/test_mask_*).
// ---------------------------------------------------------------------------
constexpr int kMaxT = 32, kMaxL = 64;

__global__ void __launch_bounds__(128)
mask_kernel(const float* __restrict__ noise_t, const float* __restrict__ noise_l, int N, int T, int L, int len_t,
            int len_l, int64_t* __restrict__ ids_keep, int64_t* __restrict__ ids_restore, float* __restrict__ mask,
            int32_t* __restrict__ ids_keep32, int32_t* __restrict__ ids_restore32) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* nt = noise_t + (size_t)n * T;
  const float* nl = noise_l + (size_t)n * L;
  uint32_t keep_t = 0;
  uint64_t keep_l = 0;
  for (int i = 0; i < T; ++i) {
    const float v = nt[i];
    int rank = 0;
    for (int j = 0; j < T; ++j) { const float w = nt[j]; rank += (w < v) || (w == v && j < i); }
    if (rank < len_t) keep_t |= 1u << i;
  }
  for (int i = 0; i < L; ++i) {
    const float v = nl[i];
    int rank = 0;
    for (int j = 0; j < L; ++j) { const float w = nl[j]; rank += (w < v) || (w == v && j < i); }
    if (rank < len_l) keep_l |= 1ull << i;
  }
  const int P = T * L;
  const int K = len_t * len_l;
  const int n1 = len_t * (L - len_l) + (T - len_t) * len_l;  // tokens dropped on exactly one axis
  int c0 = 0, c1 = K, c2 = K + n1;
  for (int t = 0; t < T; ++t) {
    const int dt = (keep_t >> t) & 1u ? 0 : 1;
    for (int l = 0; l < L; ++l) {
      const int dropped = dt + ((keep_l >> l) & 1ull ? 0 : 1);
      const int p = t * L + l;
      int pos;
      if (dropped == 0) {
        pos = c0++;
        ids_keep[(size_t)n * K + pos] = p;
        if (ids_keep32) ids_keep32[(size_t)n * K + pos] = p;
      } else if (dropped == 1) pos = c1++;
      else pos = c2++;
      ids_restore[(size_t)n * P + p] = pos;
      if (ids_restore32) ids_restore32[(size_t)n * P + p] = pos;
      mask[(size_t)n * P + p] = dropped ? 1.0f : 0.0f;
    }
  }
}

int launch_mask(const float* noise_t, const float* noise_l, int N, int T, int L, int len_t, int len_l,
                int64_t* ids_keep, int64_t* ids_restore, float* mask, int32_t* ids_keep32, int32_t* ids_restore32,
                cudaStream_t stream) {
  HS_REQUIRE(N >= 0 && T >= 1 && T <= kMaxT && L >= 1 && L <= kMaxL, "mask: unsupported shape N=%d T=%d L=%d", N, T, L);
  HS_REQUIRE(len_t >= 1 && len_t <= T && len_l >= 1 && len_l <= L, "mask: bad visible shape (%d,%d)", len_t, len_l);
  if (N == 0) return kOk;
  mask_kernel<<<ceil_div(N, 128), 128, 0, stream>>>(noise_t, noise_l, N, T, L, len_t, len_l, ids_keep, ids_restore, mask,
                                                    ids_keep32, ids_restore32);
  HS_CHECK_LAUNCH("mask_kernel");
  return kOk;
}

// ---------------------------------------------------------------------------
// Patch embedding forward.  Reference: Conv3d(stride==kernel) + 'ncts->ntsc'
// (Models.py:147-158), gather of the visible tokens (:528), + pos_embed gather
// (:547-550), followed by the first LayerNorm of the consuming block(s)
// (Block.forward :304).  One CTA walks samples; the cube is read once with
// coalesced 128-bit loads, only visible tokens are embedded.
// ---------------------------------------------------------------------------
constexpr int kEmbedThreads = 256;
constexpr int kTokChunk = 12;

__device__ __forceinline__ int cube_index(const PatchGeom& g, int p, int j) {
  // token p = (t, h, w); element j = (u, pp, q)  (Models.py:469-471)
  const int t = p / g.L, hw = p - t * g.L;
  const int h = hw / g.G, w = hw - h * g.G;
  const int pp2 = g.p * g.p;
  const int u = j / pp2, r = j - u * pp2;
  const int pp = r / g.p, q = r - pp * g.p;
  return ((t * g.u + u) * g.img + (h * g.p + pp)) * g.img + (w * g.p + q);
}


// cooperative load of sample n's cube into smem ([band][y][x], the layout of the reference's [N,1,bands,H,W] input):
// either a contiguous cube, or gathered on the fly from an HWC scene (each pixel = `bands` contiguous floats)
__device__ __forceinline__ void stage_cube(const EmbedArgs& a, const PatchGeom& g, int n, float* sCube, int tid, int nthreads) {
  if (a.scene == nullptr) {
    const float4* src = reinterpret_cast<const float4*>(a.imgs + (size_t)n * g.cube);
    for (int i = tid; i < g.cube / 4; i += nthreads) reinterpret_cast<float4*>(sCube)[i] = ld_stream_f4(src + i);
  } else {
    const int wout = a.scene_w - g.img + 1;
    const long long pix = a.pixel0 + n;
    const int r = (int)(pix / wout), c = (int)(pix - (long long)r * wout);
    const int quads = g.bands / 4, plane = g.img * g.img;
    for (int i = tid; i < plane * quads; i += nthreads) {
      const int px = i / quads, qd = i - px * quads;
      const int y = px / g.img, x = px - y * g.img;
      const float4 v = __ldg(reinterpret_cast<const float4*>(a.scene + ((size_t)(r + y) * a.scene_w + (c + x)) * g.bands) + qd);
      float* d = sCube + (size_t)(qd * 4) * plane + px;
      d[0] = v.x; d[plane] = v.y; d[2 * plane] = v.z; d[3 * plane] = v.w;
    }
  }
}

__global__ void __launch_bounds__(kEmbedThreads)
embed_fwd_kernel(EmbedArgs a) {
  extern __shared__ float sm[];
  const PatchGeom g = a.g;
  const int D = a.D, K = a.K, PK = g.PK;
  float* sW = sm;                         // [PK][D]  (transposed weight)
  float* sCube = sW + (size_t)PK * D;     // [cube]
  float* sX = sCube + g.cube;             // [K][D]
  int* sIdx = reinterpret_cast<int*>(sX + (size_t)K * D);  // [K][PK] cube offsets of the kept tokens
  int* sTok = sIdx + (size_t)K * PK;      // [K]

  for (int i = threadIdx.x; i < PK * D; i += blockDim.x) {
    const int d = i / PK, j = i - d * PK;
    sW[(size_t)j * D + d] = a.W[i];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;

  for (int n = blockIdx.x; n < a.N; n += gridDim.x) {
    __syncthreads();
    stage_cube(a, g, n, sCube, threadIdx.x, blockDim.x);
    for (int i = threadIdx.x; i < K; i += blockDim.x) sTok[i] = a.ids_keep ? a.ids_keep[(size_t)n * K + i] : i;
    __syncthreads();
    for (int i = threadIdx.x; i < K * PK; i += blockDim.x) {
      const int k = i / PK, j = i - k * PK;
      sIdx[i] = cube_index(g, sTok[k], j);
    }
    __syncthreads();
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
      const float b = a.bias ? a.bias[d] : 0.f;
      for (int k0 = 0; k0 < K; k0 += kTokChunk) {
        float acc[kTokChunk];
#pragma unroll
        for (int k = 0; k < kTokChunk; ++k) acc[k] = 0.f;
        for (int j = 0; j < PK; ++j) {
          const float w = sW[(size_t)j * D + d];
#pragma unroll
          for (int k = 0; k < kTokChunk; ++k)
            if (k0 + k < K) acc[k] = fmaf(sCube[sIdx[(k0 + k) * PK + j]], w, acc[k]);
        }
#pragma unroll
        for (int k = 0; k < kTokChunk; ++k)
          if (k0 + k < K) sX[(size_t)(k0 + k) * D + d] = acc[k] + b + __ldg(a.pos + (size_t)sTok[k0 + k] * D + d);
      }
    }
    __syncthreads();
    // LayerNorm(s) + stores, one warp per token
    for (int k = warp; k < K; k += nwarps) {
      const float* xr = sX + (size_t)k * D;
      const size_t m = (size_t)n * K + k;
      float s = 0.f;
      for (int i = lane; i < D; i += 32) s += xr[i];
      const float mean = warp_sum(s) / D;
      float sq = 0.f;
      for (int i = lane; i < D; i += 32) { const float dv = xr[i] - mean; sq = fmaf(dv, dv, sq); }
      const float rstd = rsqrtf(warp_sum(sq) / D + a.eps);
      for (int i = lane; i < D; i += 32) {
        const float v = xr[i];
        a.x[m * D + i] = v;
        const float xh = (v - mean) * rstd;
        if (a.ln_a) a.ln_a[m * D + i] = __float2bfloat16_rn(fmaf(xh, a.gamma_a[i], a.beta_a[i]));
        if (a.ln_b) a.ln_b[m * D + i] = __float2bfloat16_rn(fmaf(xh, a.gamma_b[i], a.beta_b[i]));
      }
      if (lane == 0) {
        if (a.stats_a) { a.stats_a[2 * m] = mean; a.stats_a[2 * m + 1] = rstd; }
        if (a.stats_b) { a.stats_b[2 * m] = mean; a.stats_b[2 * m + 1] = rstd; }
      }
    }
  }
}

// Specialisation for a compile-time patch geometry (the reference's 8x3x3 patches of a 9x9 cube): the (u,p,q)
// loops unroll into immediate shared-memory offsets.  The kernel is bound by shared-memory loads, not FMAs, so every
// thread owns kEmbedCh output channels and TK tokens: one broadcast LDS of a cube value feeds kEmbedCh FMAs and one
// LDS of a weight feeds TK.  kEmbedSlots samples are in flight per CTA (64 threads each) sharing the transposed
// weight tile in shared memory.
constexpr int kEmbedSlots = 4;
constexpr int kSlotThreads = 64;
constexpr int kEmbedCh = 4;     // output channels per thread: each staged cube value feeds kEmbedCh FMAs

template <int U, int P, int IMG, int TK>
__global__ void __launch_bounds__(kSlotThreads * kEmbedSlots)
embed_fwd_fixed_kernel(EmbedArgs a) {
  extern __shared__ float sm[];
  const PatchGeom g = a.g;
  constexpr int PK = U * P * P;
  const int D = a.D, K = a.K;
  const int slot = threadIdx.x / kSlotThreads, tid = threadIdx.x % kSlotThreads;
  float* sW = sm;                                              // [PK][D] shared by all slots
  const size_t slot_floats = ((size_t)g.cube + (size_t)K * D + 2 * (size_t)K + 3) / 4 * 4;   // keep every slot 16-byte aligned
  float* sCube = sW + (size_t)PK * D + slot * slot_floats;     // [cube]
  float* sX = sCube + g.cube;                                  // [K][D]
  int* sBase = reinterpret_cast<int*>(sX + (size_t)K * D);     // [K] cube offset of each kept token
  int* sTok = sBase + K;                                       // [K]
  // transposed weight tile: independent 128-bit loads, several in flight per thread (PK is a multiple of 4)
  {
    const float4* W4 = reinterpret_cast<const float4*>(a.W);
#pragma unroll 4
    for (int i = threadIdx.x; i < PK * D / 4; i += blockDim.x) {
      const float4 w = __ldg(W4 + i);
      const int d = (4 * i) / PK, j = 4 * i - d * PK;
      sW[(size_t)j * D + d] = w.x; sW[(size_t)(j + 1) * D + d] = w.y; sW[(size_t)(j + 2) * D + d] = w.z; sW[(size_t)(j + 3) * D + d] = w.w;
    }
  }
  const int warp = tid >> 5, lane = tid & 31, nwarps = kSlotThreads >> 5;
  const int dq = D / kEmbedCh;                                 // thread t owns channels t, t + dq, t + 2 dq, t + 3 dq
  // LayerNorm affine of the channels this lane normalises (first 256-channel slab), loaded once
  float ga[8], ba[8], gb[8], bb[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = lane * 8 + j;
    ga[j] = (a.ln_a && c < D) ? a.gamma_a[c] : 0.f; ba[j] = (a.ln_a && c < D) ? a.beta_a[c] : 0.f;
    gb[j] = (a.ln_b && c < D) ? a.gamma_b[c] : 0.f; bb[j] = (a.ln_b && c < D) ? a.beta_b[c] : 0.f;
  }
  const int rounds = (a.N + gridDim.x * kEmbedSlots - 1) / (gridDim.x * kEmbedSlots);
  for (int rd = 0; rd < rounds; ++rd) {
    const int n = (rd * gridDim.x + blockIdx.x) * kEmbedSlots + slot;
    const bool live = n < a.N;
    __syncthreads();
    if (live) {
      stage_cube(a, g, n, sCube, tid, kSlotThreads);
      for (int i = tid; i < K; i += kSlotThreads) {
        const int tok = a.ids_keep ? a.ids_keep[(size_t)n * K + i] : i;
        sTok[i] = tok;
        sBase[i] = cube_index(g, tok, 0);
      }
    }
    __syncthreads();
    if (live) {
      for (int d = tid; d < dq; d += kSlotThreads) {
        float b[kEmbedCh];
#pragma unroll
        for (int c = 0; c < kEmbedCh; ++c) b[c] = a.bias ? a.bias[d + c * dq] : 0.f;
        for (int k0 = 0; k0 < K; k0 += TK) {
          float acc[TK][kEmbedCh];
          const float* base[TK];
#pragma unroll
          for (int k = 0; k < TK; ++k) {
            base[k] = sCube + sBase[k0 + k < K ? k0 + k : K - 1];
#pragma unroll
            for (int c = 0; c < kEmbedCh; ++c) acc[k][c] = 0.f;
          }
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int pp = 0; pp < P; ++pp)
#pragma unroll
              for (int q = 0; q < P; ++q) {
                float w[kEmbedCh];
#pragma unroll
                for (int c = 0; c < kEmbedCh; ++c) w[c] = sW[(size_t)((u * P + pp) * P + q) * D + d + c * dq];
#pragma unroll
                for (int k = 0; k < TK; ++k) {
                  const float v = base[k][(u * IMG + pp) * IMG + q];
#pragma unroll
                  for (int c = 0; c < kEmbedCh; ++c) acc[k][c] = fmaf(v, w[c], acc[k][c]);
                }
              }
#pragma unroll
          for (int k = 0; k < TK; ++k)
            if (k0 + k < K) {
#pragma unroll
              for (int c = 0; c < kEmbedCh; ++c)
                sX[(size_t)(k0 + k) * D + d + c * dq] = acc[k][c] + b[c] + __ldg(a.pos + (size_t)sTok[k0 + k] * D + d + c * dq);
            }
        }
      }
    }
    __syncthreads();
    if (live) {
      for (int k = warp; k < K; k += nwarps) {
        const float* xr = sX + (size_t)k * D;
        const size_t m = (size_t)n * K + k;
        // lane l owns channels [8 l, 8 l + 8) of every 256-channel slab: 128-bit shared loads and global stores
        float s = 0.f;
        for (int i = lane * 8; i < D; i += 256) {
          const float4 v0 = *reinterpret_cast<const float4*>(xr + i), v1 = *reinterpret_cast<const float4*>(xr + i + 4);
          s += (v0.x + v0.y) + (v0.z + v0.w) + (v1.x + v1.y) + (v1.z + v1.w);
        }
        const float mean = warp_sum(s) / D;
        float sq = 0.f;
        for (int i = lane * 8; i < D; i += 256) {
#pragma unroll
          for (int j = 0; j < 8; ++j) { const float dv = xr[i + j] - mean; sq = fmaf(dv, dv, sq); }
        }
        const float rstd = rsqrtf(warp_sum(sq) / D + a.eps);
        for (int i = lane * 8; i < D; i += 256) {
          float v[8], xh[8], o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) { v[j] = xr[i + j]; xh[j] = (v[j] - mean) * rstd; }
          *reinterpret_cast<float4*>(a.x + m * D + i) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(a.x + m * D + i + 4) = make_float4(v[4], v[5], v[6], v[7]);
          if (a.ln_a) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = i < 256 ? fmaf(xh[j], ga[j], ba[j]) : fmaf(xh[j], __ldg(a.gamma_a + i + j), __ldg(a.beta_a + i + j));
            *reinterpret_cast<uint4*>(a.ln_a + m * D + i) =
                make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
          }
          if (a.ln_b) {
#pragma unroll
            for (int j = 0; j < 8; ++j) o[j] = i < 256 ? fmaf(xh[j], gb[j], bb[j]) : fmaf(xh[j], __ldg(a.gamma_b + i + j), __ldg(a.beta_b + i + j));
            *reinterpret_cast<uint4*>(a.ln_b + m * D + i) =
                make_uint4(pack_bf16x2(o[0], o[1]), pack_bf16x2(o[2], o[3]), pack_bf16x2(o[4], o[5]), pack_bf16x2(o[6], o[7]));
          }
        }
        if (lane == 0) {
          if (a.stats_a) { a.stats_a[2 * m] = mean; a.stats_a[2 * m + 1] = rstd; }
          if (a.stats_b) { a.stats_b[2 * m] = mean; a.stats_b[2 * m + 1] = rstd; }
        }
      }
    }
  }
}

static size_t embed_fwd_smem(const EmbedArgs& a) {
  return ((size_t)a.g.PK * a.D + a.g.cube + (size_t)a.K * a.D) * sizeof(float) + ((size_t)a.K * a.g.PK + a.K) * sizeof(int);
}

int launch_embed_fwd(const EmbedArgs& a, cudaStream_t stream) {
  HS_REQUIRE(a.g.cube % 4 == 0, "embed: cube size %d must be a multiple of 4", a.g.cube);
  HS_REQUIRE(a.K >= 1 && a.K <= a.g.P, "embed: bad K=%d", a.K);
  HS_REQUIRE((a.imgs != nullptr) != (a.scene != nullptr), "embed: exactly one of imgs / scene must be given");
  if (a.scene) HS_REQUIRE(a.g.bands % 4 == 0 && a.scene_w >= a.g.img, "embed: scene needs bands %% 4 == 0 and width >= window");
  if (a.N == 0) return kOk;
  if (embed_fwd_mma_supported(a)) return launch_embed_fwd_mma(a, stream);
  const size_t smem = embed_fwd_smem(a);
  HS_REQUIRE(smem <= 227 * 1024, "embed: configuration needs %zu bytes of shared memory (> 227 KB)", smem);
  const int grid = a.N < kNumSMs ? a.N : kNumSMs;
  const size_t smem2 = ((size_t)a.g.PK * a.D + kEmbedSlots * (((size_t)a.g.cube + (size_t)a.K * a.D + 2 * (size_t)a.K + 3) / 4 * 4)) * sizeof(float);
  if (a.g.u == 8 && a.g.p == 3 && a.g.img == 9 && smem2 <= 227 * 1024 && a.D % 8 == 0) {
    HS_CHECK_CUDA(cudaFuncSetAttribute(embed_fwd_fixed_kernel<8, 3, 9, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    const int grid2 = ceil_div(a.N, kEmbedSlots) < kNumSMs ? ceil_div(a.N, kEmbedSlots) : kNumSMs;
    embed_fwd_fixed_kernel<8, 3, 9, 9><<<grid2, kSlotThreads * kEmbedSlots, smem2, stream>>>(a);
  } else {
    HS_CHECK_CUDA(cudaFuncSetAttribute(embed_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    embed_fwd_kernel<<<grid, kEmbedThreads, smem, stream>>>(a);
  }
  HS_CHECK_LAUNCH("embed_fwd_kernel");
  return kOk;
}

// ---------------------------------------------------------------------------
// Patch embedding backward: dW[d][j] += sum_m dx[m][d] * patch[m][j], db[d] += sum_m dx[m][d].
// (autograd of Models.py:157; the input needs no gradient.)
// Each thread owns one output channel d and PK accumulators; CTAs walk samples
// and flush with one atomicAdd per (d, j).
// ---------------------------------------------------------------------------
constexpr int kMaxPK = 128;

template <int PKC>
__global__ void __launch_bounds__(256)
embed_bwd_kernel(EmbedBwdArgs a) {
  extern __shared__ float sm[];
  const PatchGeom g = a.g;
  const int D = a.D, K = a.K, PK = g.PK;
  float* sCube = sm;
  int* sIdx = reinterpret_cast<int*>(sCube + g.cube);
  int* sTok = sIdx + (size_t)K * PK;
  // channel handled by this thread (D may exceed blockDim: loop over channel groups)
  for (int d0 = 0; d0 < D; d0 += blockDim.x) {
    const int d = d0 + threadIdx.x;
    float acc[PKC];
#pragma unroll
    for (int j = 0; j < PKC; ++j) acc[j] = 0.f;
    float bsum = 0.f;
    for (int n = blockIdx.x; n < a.N; n += gridDim.x) {
      __syncthreads();
      const float4* src = reinterpret_cast<const float4*>(a.imgs + (size_t)n * g.cube);
      for (int i = threadIdx.x; i < g.cube / 4; i += blockDim.x) reinterpret_cast<float4*>(sCube)[i] = ld_stream_f4(src + i);
      for (int i = threadIdx.x; i < K; i += blockDim.x) sTok[i] = a.ids_keep ? a.ids_keep[(size_t)n * K + i] : i;
      __syncthreads();
      for (int i = threadIdx.x; i < K * PK; i += blockDim.x) {
        const int k = i / PK, j = i - k * PK;
        sIdx[i] = cube_index(g, sTok[k], j);
      }
      __syncthreads();
      if (d < D) {
        for (int k = 0; k < K; ++k) {
          const size_t m = (size_t)n * K + k;
          float gv = a.dx_a[m * D + d];
          if (a.dx_b) gv += a.dx_b[m * D + d];
          bsum += gv;
#pragma unroll
          for (int j = 0; j < PKC; ++j)
            if (j < PK) acc[j] = fmaf(gv, sCube[sIdx[k * PK + j]], acc[j]);
        }
      }
    }
    if (d < D) {
#pragma unroll
      for (int j = 0; j < PKC; ++j)
        if (j < PK) atomicAdd(a.dW + (size_t)d * PK + j, acc[j]);
      if (a.dbias) atomicAdd(a.dbias + d, bsum);
    }
  }
}

// Each thread owns kBwdCh output channels (d, d + D/kBwdCh, ...) and their PK accumulators: one broadcast LDS of a
// cube value feeds kBwdCh FMAs (the loop is bound by shared-memory loads).
constexpr int kBwdCh = 2;

template <int U, int P, int IMG>
__global__ void __launch_bounds__(128)
embed_bwd_fixed_kernel(EmbedBwdArgs a) {
  extern __shared__ float sm[];
  const PatchGeom g = a.g;
  constexpr int PK = U * P * P;
  const int D = a.D, K = a.K;
  const int dq = D / kBwdCh;
  float* sCube = sm;
  int* sBase = reinterpret_cast<int*>(sCube + g.cube);
  for (int d0 = 0; d0 < dq; d0 += blockDim.x) {
    const int d = d0 + threadIdx.x;
    float acc[kBwdCh][PK];
    float bsum[kBwdCh];
#pragma unroll
    for (int c = 0; c < kBwdCh; ++c) {
      bsum[c] = 0.f;
#pragma unroll
      for (int j = 0; j < PK; ++j) acc[c][j] = 0.f;
    }
    for (int n = blockIdx.x; n < a.N; n += gridDim.x) {
      __syncthreads();
      const float4* src = reinterpret_cast<const float4*>(a.imgs + (size_t)n * g.cube);
      for (int i = threadIdx.x; i < g.cube / 4; i += blockDim.x) reinterpret_cast<float4*>(sCube)[i] = ld_stream_f4(src + i);
      for (int i = threadIdx.x; i < K; i += blockDim.x) sBase[i] = cube_index(g, a.ids_keep ? a.ids_keep[(size_t)n * K + i] : i, 0);
      __syncthreads();
      if (d < dq) {
        for (int k = 0; k < K; ++k) {
          const size_t m = (size_t)n * K + k;
          float gv[kBwdCh];
#pragma unroll
          for (int c = 0; c < kBwdCh; ++c) {
            gv[c] = a.dx_a[m * D + d + c * dq];
            if (a.dx_b) gv[c] += a.dx_b[m * D + d + c * dq];
            bsum[c] += gv[c];
          }
          const float* base = sCube + sBase[k];
#pragma unroll
          for (int u = 0; u < U; ++u)
#pragma unroll
            for (int pp = 0; pp < P; ++pp)
#pragma unroll
              for (int q = 0; q < P; ++q) {
                const float v = base[(u * IMG + pp) * IMG + q];
#pragma unroll
                for (int c = 0; c < kBwdCh; ++c) acc[c][(u * P + pp) * P + q] = fmaf(gv[c], v, acc[c][(u * P + pp) * P + q]);
              }
        }
      }
    }
    if (d < dq) {
#pragma unroll
      for (int c = 0; c < kBwdCh; ++c) {
        // PK is a multiple of 4 and every row of dW starts 16-byte aligned: one vector reduction per 4 accumulators
#pragma unroll
        for (int j = 0; j < PK; j += 4)
          red_add_f32x4(a.dW + (size_t)(d + c * dq) * PK + j, acc[c][j], acc[c][j + 1], acc[c][j + 2], acc[c][j + 3]);
        if (a.dbias) atomicAdd(a.dbias + d + c * dq, bsum[c]);
      }
    }
  }
}

int launch_embed_bwd(const EmbedBwdArgs& a, cudaStream_t stream) {
  HS_REQUIRE(a.g.PK <= kMaxPK, "embed_bwd: patch of %d elements unsupported (max %d)", a.g.PK, kMaxPK);
  if (a.N == 0) return kOk;
  if (embed_bwd_mma_supported(a)) return launch_embed_bwd_mma(a, stream);
  const size_t smem = (size_t)a.g.cube * sizeof(float) + ((size_t)a.K * a.g.PK + a.K) * sizeof(int);
  HS_REQUIRE(smem <= 227 * 1024, "embed_bwd: needs %zu bytes of shared memory", smem);
  const int grid = a.N < 2 * kNumSMs ? a.N : 2 * kNumSMs;   // two CTAs per SM
  if (a.g.u == 8 && a.g.p == 3 && a.g.img == 9 && a.D % kBwdCh == 0) {
    HS_CHECK_CUDA(cudaFuncSetAttribute(embed_bwd_fixed_kernel<8, 3, 9>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    // every CTA flushes a full [D, PK] partial with reductions: as few CTAs as keep the SMs busy
    embed_bwd_fixed_kernel<8, 3, 9><<<grid, 128, smem, stream>>>(a);
  } else if (a.g.PK <= 72) {
    HS_CHECK_CUDA(cudaFuncSetAttribute(embed_bwd_kernel<72>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    embed_bwd_kernel<72><<<grid, 256, smem, stream>>>(a);
  } else {
    HS_CHECK_CUDA(cudaFuncSetAttribute(embed_bwd_kernel<kMaxPK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    embed_bwd_kernel<kMaxPK><<<grid, 256, smem, stream>>>(a);
  }
  HS_CHECK_LAUNCH("embed_bwd_kernel");
  return kOk;
}

}  // namespace hsimae
