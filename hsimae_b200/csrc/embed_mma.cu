// Patch embedding on tensor cores (TF32 mma.sync), forward and weight gradient, for the reference's 8x3x3 patches of a
// 9x9 cube.
//
// Reference: PatchEmbed.forward, /root/reference/Models.py:147-158 -- Conv3d(1 -> D, kernel = stride = (8,3,3)) is the
// contraction [tokens, 72] x [72, D]; followed here by the gather of the visible tokens (:528), the position-table add
// (:547-550) and the first LayerNorm(s) of the consuming block(s) (:304).  PyTorch runs this convolution through
// cuDNN's TF32 path on Ampere and later (torch.backends.cudnn.allow_tf32 defaults to True), so TF32 operands with fp32
// accumulation are the reference's own arithmetic on a GPU.
//
// The CUDA-core kernels in mask_embed.cu are bound by shared-memory loads (one LDS per 4 FMAs): 310 us forward / 256 us
// backward at batch 4096 for ~190 MB of traffic (9 % of the HBM peak).  Here the [72, D] weight lives in REGISTERS as
// B fragments (each warp owns D/8 output columns), the cube values are gathered from shared memory straight into A
// fragments, and the LayerNorm statistics meet across the 8 warps in shared memory: no [tokens, D] staging tile.
#include "kernels.cuh"

namespace hsimae {

namespace {

constexpr int kU = 8, kP = 3, kImg = 9;
constexpr int kPK = kU * kP * kP;        // 72 = 9 k-steps of 8
constexpr int kEmbThreads = 256;         // 8 warps
constexpr int kMaxGroupTok = 80;         // tokens per CTA iteration (5 m-tiles of 16)
constexpr int kMaxGroupSamples = 4;

__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait0() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// cube offset of patch element j = (u, pp, q) relative to the patch's first element
__device__ __forceinline__ int patch_off(int j) {
  const int u = j / (kP * kP), r = j - u * (kP * kP);
  const int pp = r / kP, q = r - pp * kP;
  return (u * kImg + pp) * kImg + q;
}
// cube offset of the first element of token (t, h, w)
__device__ __forceinline__ int token_base(const PatchGeom& g, int tok) {
  const int t = tok / g.L, hw = tok - t * g.L;
  const int h = hw / g.G, w = hw - h * g.G;
  return ((t * kU) * kImg + h * kP) * kImg + w * kP;
}

// cooperative load of sample n's cube ([band][y][x]): contiguous cube (asynchronous copies) or gathered from an HWC scene
__device__ __forceinline__ void load_cube(const EmbedArgs& a, const PatchGeom& g, int n, float* sCube, int tid) {
  if (a.scene == nullptr) {
    const float4* src = reinterpret_cast<const float4*>(a.imgs + (size_t)n * g.cube);
    const uint32_t dst = smem_addr(sCube);
    for (int i = tid; i < g.cube / 4; i += kEmbThreads) cp_async16(dst + (uint32_t)i * 16u, src + i);
  } else {
    const int wout = a.scene_w - g.img + 1;
    const long long pix = a.pixel0 + n;
    const int r = (int)(pix / wout), c = (int)(pix - (long long)r * wout);
    const int quads = g.bands / 4, plane = g.img * g.img;
    for (int i = tid; i < plane * quads; i += kEmbThreads) {
      const int px = i / quads, qd = i - px * quads;
      const int y = px / g.img, x = px - y * g.img;
      const float4 v = __ldg(reinterpret_cast<const float4*>(a.scene + ((size_t)(r + y) * a.scene_w + (c + x)) * g.bands) + qd);
      float* d = sCube + (size_t)(qd * 4) * plane + px;
      d[0] = v.x; d[plane] = v.y; d[2 * plane] = v.z; d[3 * plane] = v.w;
    }
  }
}

// ---------------------------------------------------------------------------
// forward: NT = n-tiles (of 8 output columns) per warp, D = 64 NT
// ---------------------------------------------------------------------------
template <int NT>
__global__ void __launch_bounds__(kEmbThreads, 2)
embed_fwd_mma_kernel(EmbedArgs a, int S) {
  extern __shared__ float sm[];
  const PatchGeom g = a.g;
  const int D = a.D, K = a.K;
  float* sCubes = sm;                                                   // [2][S][cube]
  int* sBase = reinterpret_cast<int*>(sCubes + (size_t)2 * S * g.cube);  // [kMaxGroupTok] cube offset (incl. sample slot) of each token row
  int* sTok = sBase + kMaxGroupTok;                                     // [kMaxGroupTok] position of the token in the [T, L] grid
  float* sRed = reinterpret_cast<float*>(sTok + kMaxGroupTok);          // [2][8 warps][16 rows][2]: (sum, sum of squares) partials
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int col0 = warp * NT * 8;                                       // first output column of this warp

  // B fragments of the whole [72, D/8] weight slice: b0 = W[n][8 ks + tq], b1 = W[n][8 ks + tq + 4] with n = col0 + 8 nt + gq
  uint32_t breg[9][NT][2];
#pragma unroll
  for (int ks = 0; ks < 9; ++ks)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float* wr = a.W + (size_t)(col0 + nt * 8 + gq) * kPK + ks * 8 + tq;
      breg[ks][nt][0] = f2tf32(__ldg(wr)); breg[ks][nt][1] = f2tf32(__ldg(wr + 4));
    }
  int offA[9][2];
#pragma unroll
  for (int ks = 0; ks < 9; ++ks) { offA[ks][0] = patch_off(ks * 8 + tq); offA[ks][1] = patch_off(ks * 8 + tq + 4); }
  float2 bias[NT];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) bias[nt] = a.bias ? *reinterpret_cast<const float2*>(a.bias + col0 + nt * 8 + 2 * tq) : make_float2(0.f, 0.f);

  const int groups = (a.N + S - 1) / S;
  const bool async = a.scene == nullptr;
  int buf = 0;
  // prologue: the first group's cubes
  if ((int)blockIdx.x < groups) {
    const int n0 = blockIdx.x * S;
    for (int s = 0; s < S && n0 + s < a.N; ++s) load_cube(a, g, n0 + s, sCubes + ((size_t)buf * S + s) * g.cube, threadIdx.x);
    cp_async_commit();
  }
  int par = 0;
  for (int gi = blockIdx.x; gi < groups; gi += gridDim.x) {
    const int n0 = gi * S;
    const int ns = a.N - n0 < S ? a.N - n0 : S;
    const int ntok = ns * K;
    for (int i = threadIdx.x; i < ntok; i += kEmbThreads) {
      const int s = i / K, k = i - s * K;
      const int tok = a.ids_keep ? a.ids_keep[(size_t)(n0 + s) * K + k] : k;
      sTok[i] = tok;
      sBase[i] = (buf * S + s) * g.cube + token_base(g, tok);
    }
    cp_async_wait0();
    __syncthreads();
    // the next group's cubes travel while this one is computed (contiguous cubes only: a scene window is gathered synchronously)
    const int gnext = gi + gridDim.x;
    if (async && gnext < groups) {
      const int m0 = gnext * S;
      for (int s = 0; s < S && m0 + s < a.N; ++s) load_cube(a, g, m0 + s, sCubes + ((size_t)(buf ^ 1) * S + s) * g.cube, threadIdx.x);
      cp_async_commit();
    }
    const int mtiles = (ntok + 15) >> 4;
    for (int mt = 0; mt < mtiles; ++mt) {
      const int r0 = mt * 16 + gq, r1 = r0 + 8;
      const bool v0 = r0 < ntok, v1 = r1 < ntok;
      const float* p0 = sCubes + sBase[v0 ? r0 : 0];
      const float* p1 = sCubes + sBase[v1 ? r1 : 0];
      float acc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
#pragma unroll
      for (int ks = 0; ks < 9; ++ks) {
        uint32_t af[4];
        af[0] = f2tf32(p0[offA[ks][0]]); af[1] = f2tf32(p1[offA[ks][0]]);
        af[2] = f2tf32(p0[offA[ks][1]]); af[3] = f2tf32(p1[offA[ks][1]]);
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) mma_tf32(acc[nt], af, breg[ks][nt][0], breg[ks][nt][1]);
      }
      // + bias + position table; row statistics over the D columns (8 warps x NT x 8 columns)
      const float* pos0 = a.pos + (size_t)sTok[v0 ? r0 : 0] * D + col0 + 2 * tq;
      const float* pos1 = a.pos + (size_t)sTok[v1 ? r1 : 0] * D + col0 + 2 * tq;
      float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const float2 e0 = __ldg(reinterpret_cast<const float2*>(pos0 + nt * 8)), e1 = __ldg(reinterpret_cast<const float2*>(pos1 + nt * 8));
        acc[nt][0] += bias[nt].x + e0.x; acc[nt][1] += bias[nt].y + e0.y;
        acc[nt][2] += bias[nt].x + e1.x; acc[nt][3] += bias[nt].y + e1.y;
        s0 += acc[nt][0] + acc[nt][1]; q0 = fmaf(acc[nt][0], acc[nt][0], fmaf(acc[nt][1], acc[nt][1], q0));
        s1 += acc[nt][2] + acc[nt][3]; q1 = fmaf(acc[nt][2], acc[nt][2], fmaf(acc[nt][3], acc[nt][3], q1));
      }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      q0 += __shfl_xor_sync(0xffffffffu, q0, 1); q0 += __shfl_xor_sync(0xffffffffu, q0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      q1 += __shfl_xor_sync(0xffffffffu, q1, 1); q1 += __shfl_xor_sync(0xffffffffu, q1, 2);
      float* red = sRed + (size_t)par * 8 * 16 * 2;
      if (tq == 0) {
        *reinterpret_cast<float2*>(red + (warp * 16 + gq) * 2) = make_float2(s0, q0);
        *reinterpret_cast<float2*>(red + (warp * 16 + gq + 8) * 2) = make_float2(s1, q1);
      }
      __syncthreads();
      float ts0 = 0.f, tq0 = 0.f, ts1 = 0.f, tq1 = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) {
        const float2 x0 = *reinterpret_cast<const float2*>(red + (w * 16 + gq) * 2), x1 = *reinterpret_cast<const float2*>(red + (w * 16 + gq + 8) * 2);
        ts0 += x0.x; tq0 += x0.y; ts1 += x1.x; tq1 += x1.y;
      }
      par ^= 1;   // the next m-tile writes the other half of sRed: no second barrier
      const float invD = 1.0f / (float)D;
      const float mean0 = ts0 * invD, mean1 = ts1 * invD;
      const float rstd0 = rsqrtf(fmaxf(tq0 * invD - mean0 * mean0, 0.f) + a.eps), rstd1 = rsqrtf(fmaxf(tq1 * invD - mean1 * mean1, 0.f) + a.eps);
      const size_t m0g = (size_t)n0 * K + r0, m1g = (size_t)n0 * K + r1;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int c = col0 + nt * 8 + 2 * tq;
        if (v0) *reinterpret_cast<float2*>(a.x + m0g * D + c) = make_float2(acc[nt][0], acc[nt][1]);
        if (v1) *reinterpret_cast<float2*>(a.x + m1g * D + c) = make_float2(acc[nt][2], acc[nt][3]);
        const float h00 = (acc[nt][0] - mean0) * rstd0, h01 = (acc[nt][1] - mean0) * rstd0;
        const float h10 = (acc[nt][2] - mean1) * rstd1, h11 = (acc[nt][3] - mean1) * rstd1;
        if (a.ln_a) {
          const float2 ga = __ldg(reinterpret_cast<const float2*>(a.gamma_a + c)), be = __ldg(reinterpret_cast<const float2*>(a.beta_a + c));
          if (v0) *reinterpret_cast<uint32_t*>(a.ln_a + m0g * D + c) = pack_bf16x2(fmaf(h00, ga.x, be.x), fmaf(h01, ga.y, be.y));
          if (v1) *reinterpret_cast<uint32_t*>(a.ln_a + m1g * D + c) = pack_bf16x2(fmaf(h10, ga.x, be.x), fmaf(h11, ga.y, be.y));
        }
        if (a.ln_b) {
          const float2 ga = __ldg(reinterpret_cast<const float2*>(a.gamma_b + c)), be = __ldg(reinterpret_cast<const float2*>(a.beta_b + c));
          if (v0) *reinterpret_cast<uint32_t*>(a.ln_b + m0g * D + c) = pack_bf16x2(fmaf(h00, ga.x, be.x), fmaf(h01, ga.y, be.y));
          if (v1) *reinterpret_cast<uint32_t*>(a.ln_b + m1g * D + c) = pack_bf16x2(fmaf(h10, ga.x, be.x), fmaf(h11, ga.y, be.y));
        }
      }
      if (warp == 0 && tq == 0) {
        if (v0) { if (a.stats_a) *reinterpret_cast<float2*>(a.stats_a + 2 * m0g) = make_float2(mean0, rstd0); if (a.stats_b) *reinterpret_cast<float2*>(a.stats_b + 2 * m0g) = make_float2(mean0, rstd0); }
        if (v1) { if (a.stats_a) *reinterpret_cast<float2*>(a.stats_a + 2 * m1g) = make_float2(mean1, rstd1); if (a.stats_b) *reinterpret_cast<float2*>(a.stats_b + 2 * m1g) = make_float2(mean1, rstd1); }
      }
    }
    __syncthreads();   // every warp is done with this group's cubes and token tables
    if (async) buf ^= 1;
    else if (gnext < groups) {
      const int m0 = gnext * S;
      for (int s = 0; s < S && m0 + s < a.N; ++s) load_cube(a, g, m0 + s, sCubes + ((size_t)buf * S + s) * g.cube, threadIdx.x);
    }
  }
  cp_async_wait0();
}

// ---------------------------------------------------------------------------
// backward: dW[d][j] += sum_m dx[m][d] * patch[m][j],  db[d] += sum_m dx[m][d]     (autograd of Models.py:157)
//   MMA roles: M = d (each warp owns MT m-tiles of 16 channels, D = 128 MT), N = j (9 n-tiles), K = tokens (8 per step).
//   A = dx^T comes straight from global memory (every element is used by exactly one warp: nothing to share, 8 lanes
//   read 32 contiguous bytes), B = patch values gathered from the staged cubes.
// ---------------------------------------------------------------------------
template <int MT>
__global__ void __launch_bounds__(kEmbThreads, 2)
embed_bwd_mma_kernel(EmbedBwdArgs a, int S) {
  extern __shared__ float sm[];
  const PatchGeom g = a.g;
  const int D = a.D, K = a.K;
  float* sCubes = sm;                                                   // [2][S][cube]
  int* sBase = reinterpret_cast<int*>(sCubes + (size_t)2 * S * g.cube);  // [kMaxGroupTok]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, tq = lane & 3;
  const int d0 = warp * MT * 16;
  int offB[9];
#pragma unroll
  for (int nt = 0; nt < 9; ++nt) offB[nt] = patch_off(nt * 8 + gq);
  float acc[MT][9][4];
  float bsum[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    bsum[mt][0] = bsum[mt][1] = 0.f;
#pragma unroll
    for (int nt = 0; nt < 9; ++nt) { acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f; }
  }
  EmbedArgs ld{};   // load_cube reads only these fields
  ld.imgs = a.imgs; ld.scene = nullptr;
  const int groups = (a.N + S - 1) / S;
  int buf = 0;
  if ((int)blockIdx.x < groups) {
    const int n0 = blockIdx.x * S;
    for (int s = 0; s < S && n0 + s < a.N; ++s) load_cube(ld, g, n0 + s, sCubes + ((size_t)buf * S + s) * g.cube, threadIdx.x);
    cp_async_commit();
  }
  for (int gi = blockIdx.x; gi < groups; gi += gridDim.x) {
    const int n0 = gi * S;
    const int ns = a.N - n0 < S ? a.N - n0 : S;
    const int ntok = ns * K;
    for (int i = threadIdx.x; i < kMaxGroupTok; i += kEmbThreads) {
      if (i < ntok) {
        const int s = i / K, k = i - s * K;
        const int tok = a.ids_keep ? a.ids_keep[(size_t)(n0 + s) * K + k] : k;
        sBase[i] = (buf * S + s) * g.cube + token_base(g, tok);
      } else {
        sBase[i] = buf * S * g.cube;   // padding rows: any finite cube value (their dx operand is zero)
      }
    }
    cp_async_wait0();
    __syncthreads();
    const int gnext = gi + gridDim.x;
    if (gnext < groups) {
      const int m0 = gnext * S;
      for (int s = 0; s < S && m0 + s < a.N; ++s) load_cube(ld, g, m0 + s, sCubes + ((size_t)(buf ^ 1) * S + s) * g.cube, threadIdx.x);
      cp_async_commit();
    }
    const size_t mrow0 = (size_t)n0 * K;
    const int ksteps = (ntok + 7) >> 3;
    for (int ks = 0; ks < ksteps; ++ks) {
      const int t0 = ks * 8 + tq, t1 = t0 + 4;          // token columns of this lane's A elements
      const bool u0 = t0 < ntok, u1 = t1 < ntok;
      uint32_t af[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const int d = d0 + mt * 16 + gq;
        float x00 = 0.f, x10 = 0.f, x01 = 0.f, x11 = 0.f;   // (row d | d + 8, token t0 | t1)
        if (u0) { const float* p = a.dx_a + (mrow0 + t0) * D + d; x00 = __ldg(p); x10 = __ldg(p + 8); if (a.dx_b) { const float* r = a.dx_b + (mrow0 + t0) * D + d; x00 += __ldg(r); x10 += __ldg(r + 8); } }
        if (u1) { const float* p = a.dx_a + (mrow0 + t1) * D + d; x01 = __ldg(p); x11 = __ldg(p + 8); if (a.dx_b) { const float* r = a.dx_b + (mrow0 + t1) * D + d; x01 += __ldg(r); x11 += __ldg(r + 8); } }
        bsum[mt][0] += x00 + x01; bsum[mt][1] += x10 + x11;
        af[mt][0] = f2tf32(x00); af[mt][1] = f2tf32(x10); af[mt][2] = f2tf32(x01); af[mt][3] = f2tf32(x11);
      }
      const float* p0 = sCubes + sBase[t0 < kMaxGroupTok ? t0 : 0];
      const float* p1 = sCubes + sBase[t1 < kMaxGroupTok ? t1 : 0];
#pragma unroll
      for (int nt = 0; nt < 9; ++nt) {
        const uint32_t b0 = f2tf32(p0[offB[nt]]), b1 = f2tf32(p1[offB[nt]]);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) mma_tf32(acc[mt][nt], af[mt], b0, b1);
      }
    }
    __syncthreads();
    buf ^= 1;
  }
  cp_async_wait0();
  // flush: one vector reduction per accumulator pair (columns 2 tq, 2 tq + 1 of an n-tile)
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const int d = d0 + mt * 16 + gq;
#pragma unroll
    for (int nt = 0; nt < 9; ++nt) {
      const int j = nt * 8 + 2 * tq;
      asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(a.dW + (size_t)d * kPK + j), "f"(acc[mt][nt][0]), "f"(acc[mt][nt][1]) : "memory");
      asm volatile("red.global.add.v2.f32 [%0], {%1,%2};" ::"l"(a.dW + (size_t)(d + 8) * kPK + j), "f"(acc[mt][nt][2]), "f"(acc[mt][nt][3]) : "memory");
    }
    if (a.dbias) {
      float b0 = bsum[mt][0], b1 = bsum[mt][1];
      b0 += __shfl_xor_sync(0xffffffffu, b0, 1); b0 += __shfl_xor_sync(0xffffffffu, b0, 2);
      b1 += __shfl_xor_sync(0xffffffffu, b1, 1); b1 += __shfl_xor_sync(0xffffffffu, b1, 2);
      if (tq == 0) { atomicAdd(a.dbias + d, b0); atomicAdd(a.dbias + d + 8, b1); }
    }
  }
}

int group_samples(int K) {
  int S = kMaxGroupTok / K;
  if (S > kMaxGroupSamples) S = kMaxGroupSamples;
  return S < 1 ? 1 : S;
}
size_t embed_mma_smem(const PatchGeom& g, int S) {
  return (size_t)2 * S * g.cube * sizeof(float) + 2 * kMaxGroupTok * sizeof(int) + 2 * 8 * 16 * 2 * sizeof(float);
}
bool geometry_ok(const PatchGeom& g, int D, int K) {
  return g.u == kU && g.p == kP && g.img == kImg && D % 64 == 0 && D >= 64 && D <= 256 && K >= 1 && K <= kMaxGroupTok && g.cube % 4 == 0;
}

}  // namespace

// HSIMAE_EMBED_MMA=0 keeps the CUDA-core kernels (A/B measurements)
static bool embed_mma_enabled() {
  static const bool on = !(getenv("HSIMAE_EMBED_MMA") && atoi(getenv("HSIMAE_EMBED_MMA")) == 0);
  return on;
}

bool embed_fwd_mma_supported(const EmbedArgs& a) {
  if (!embed_mma_enabled() || !geometry_ok(a.g, a.D, a.K)) return false;
  if (a.scene && a.g.bands % 4 != 0) return false;
  return embed_mma_smem(a.g, group_samples(a.K)) <= 110 * 1024;
}

int launch_embed_fwd_mma(const EmbedArgs& a, cudaStream_t stream) {
  const int S = group_samples(a.K);
  const size_t smem = embed_mma_smem(a.g, S);
  const int groups = ceil_div(a.N, S);
  const int grid = groups < 2 * kNumSMs ? groups : 2 * kNumSMs;
  switch (a.D / 64) {
    case 1:
      HS_CHECK_CUDA(cudaFuncSetAttribute(embed_fwd_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      embed_fwd_mma_kernel<1><<<grid, kEmbThreads, smem, stream>>>(a, S); break;
    case 2:
      HS_CHECK_CUDA(cudaFuncSetAttribute(embed_fwd_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      embed_fwd_mma_kernel<2><<<grid, kEmbThreads, smem, stream>>>(a, S); break;
    case 3:
      HS_CHECK_CUDA(cudaFuncSetAttribute(embed_fwd_mma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      embed_fwd_mma_kernel<3><<<grid, kEmbThreads, smem, stream>>>(a, S); break;
    default:
      HS_CHECK_CUDA(cudaFuncSetAttribute(embed_fwd_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      embed_fwd_mma_kernel<4><<<grid, kEmbThreads, smem, stream>>>(a, S); break;
  }
  HS_CHECK_LAUNCH("embed_fwd_mma_kernel");
  return kOk;
}

bool embed_bwd_mma_supported(const EmbedBwdArgs& a) {
  // each warp owns whole 16-channel m-tiles: D = 128 or 256
  if (!embed_mma_enabled() || !geometry_ok(a.g, a.D, a.K) || a.D % 128 != 0) return false;
  return embed_mma_smem(a.g, group_samples(a.K)) <= 110 * 1024;
}

int launch_embed_bwd_mma(const EmbedBwdArgs& a, cudaStream_t stream) {
  const int S = group_samples(a.K);
  const size_t smem = embed_mma_smem(a.g, S);
  const int groups = ceil_div(a.N, S);
  // every CTA flushes a full [D, 72] partial with reductions: as few CTAs as keep the SMs busy
  const int grid = groups < 2 * kNumSMs ? groups : 2 * kNumSMs;
  if (a.D == 128) {
    HS_CHECK_CUDA(cudaFuncSetAttribute(embed_bwd_mma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    embed_bwd_mma_kernel<1><<<grid, kEmbThreads, smem, stream>>>(a, S);
  } else {
    HS_CHECK_CUDA(cudaFuncSetAttribute(embed_bwd_mma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    embed_bwd_mma_kernel<2><<<grid, kEmbThreads, smem, stream>>>(a, S);
  }
  HS_CHECK_LAUNCH("embed_bwd_mma_kernel");
  return kOk;
}

}  // namespace hsimae
