// LayerNorm backward and short-sequence attention (forward + backward).
#include <cstdlib>
#include "kernels.cuh"

namespace hsimae {

// ---------------------------------------------------------------------------
// LayerNorm backward fused with the residual-gradient add and the bf16 copy
// that feeds the next backward GEMM.  (autograd of Block.forward, Models.py:304-305)
//   dxhat = dy * gamma
//   dx    = dx_in + rstd * (dxhat - mean(dxhat) - xhat * mean(dxhat * xhat))
// One warp per row; dgamma/dbeta partials live in registers across the rows a
// warp visits, then one smem reduction + atomicAdd per CTA.
// ---------------------------------------------------------------------------
constexpr int kLnMaxPerLane = 8;  // D <= 256

__global__ void __launch_bounds__(256)
ln_bwd_kernel(LnBwdArgs a) {
  __shared__ float sg[8][256];
  __shared__ float sb[8][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = a.D;
  float gam[kLnMaxPerLane], dgam[kLnMaxPerLane], dbet[kLnMaxPerLane];
#pragma unroll
  for (int j = 0; j < kLnMaxPerLane; ++j) {
    const int i = lane + 32 * j;
    gam[j] = i < D ? a.gamma[i] : 0.f;
    dgam[j] = 0.f; dbet[j] = 0.f;
  }
  const float invD = 1.0f / D;
  for (int m = blockIdx.x * 8 + warp; m < a.M; m += gridDim.x * 8) {
    const float mean = a.stats[2 * (size_t)m], rstd = a.stats[2 * (size_t)m + 1];
    float xh[kLnMaxPerLane], dyv[kLnMaxPerLane];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < kLnMaxPerLane; ++j) {
      const int i = lane + 32 * j;
      if (i < D) {
        xh[j] = (a.x[(size_t)m * D + i] - mean) * rstd;
        dyv[j] = __bfloat162float(a.dy[(size_t)m * D + i]);
        const float t = dyv[j] * gam[j];
        s1 += t; s2 = fmaf(t, xh[j], s2);
        dgam[j] = fmaf(dyv[j], xh[j], dgam[j]);
        dbet[j] += dyv[j];
      } else { xh[j] = 0.f; dyv[j] = 0.f; }
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
    const float sc = a.dxb ? row_scale(a.rs, m) : 1.0f;
#pragma unroll
    for (int j = 0; j < kLnMaxPerLane; ++j) {
      const int i = lane + 32 * j;
      if (i < D) {
        float v = rstd * (dyv[j] * gam[j] - s1 - xh[j] * s2);
        if (a.dx_in) v += a.dx_in[(size_t)m * D + i];
        a.dx_out[(size_t)m * D + i] = v;
        if (a.dxb) a.dxb[(size_t)m * D + i] = __float2bfloat16_rn(sc * v);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < kLnMaxPerLane; ++j) { sg[warp][lane + 32 * j] = dgam[j]; sb[warp][lane + 32 * j] = dbet[j]; }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    float g = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { g += sg[w][i]; b += sb[w][i]; }
    if (a.dgamma) atomicAdd(a.dgamma + i, g);
    if (a.dbeta) atomicAdd(a.dbeta + i, b);
  }
}

// Vector variant for D = 32*VEC (VEC in {2,4,8}): each lane owns VEC contiguous elements (128-bit fp32 / 64..128-bit
// bf16 accesses), two rows per warp iteration so that both rows' loads are in flight before the first reduction.
template <int VEC> struct VecIO;
template <> struct VecIO<8> {
  static __device__ __forceinline__ void ldf(const float* p, float (&v)[8]) {
    // coherent streaming loads: dx_in may alias dx_out (in-place update by the owning thread)
    const float4 a = __ldcs(reinterpret_cast<const float4*>(p)), b = __ldcs(reinterpret_cast<const float4*>(p + 4));
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void ldh(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 t = ld_stream_u4(p);
    const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
  }
  static __device__ __forceinline__ void stf(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  static __device__ __forceinline__ void sth(__nv_bfloat16* p, const float (&v)[8]) {
    uint4 t; t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]); t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = t;
  }
};
template <> struct VecIO<4> {
  static __device__ __forceinline__ void ldf(const float* p, float (&v)[4]) { const float4 a = __ldcs(reinterpret_cast<const float4*>(p)); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
  static __device__ __forceinline__ void ldh(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 t = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
  }
  static __device__ __forceinline__ void stf(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
  static __device__ __forceinline__ void sth(__nv_bfloat16* p, const float (&v)[4]) {
    uint2 t; t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = t;
  }
};
template <> struct VecIO<2> {
  static __device__ __forceinline__ void ldf(const float* p, float (&v)[2]) { const float2 a = *reinterpret_cast<const float2*>(p); v[0] = a.x; v[1] = a.y; }
  static __device__ __forceinline__ void ldh(const __nv_bfloat16* p, float (&v)[2]) { const float2 a = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p)); v[0] = a.x; v[1] = a.y; }
  static __device__ __forceinline__ void stf(float* p, const float (&v)[2]) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
  static __device__ __forceinline__ void sth(__nv_bfloat16* p, const float (&v)[2]) { *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(v[0], v[1]); }
};

template <int VEC>
__global__ void __launch_bounds__(256)
ln_bwd_vec_kernel(LnBwdArgs a) {
  constexpr int D = 32 * VEC;
  __shared__ float sg[8][D];
  __shared__ float sb[8][D];
  pdl_wait();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e0 = lane * VEC;
  float gam[VEC], dgam[VEC], dbet[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) { gam[j] = a.gamma[e0 + j]; dgam[j] = 0.f; dbet[j] = 0.f; }
  const float invD = 1.0f / D;
  const int stride = gridDim.x * 8 * 2;
  for (int m0 = (blockIdx.x * 8 + warp) * 2; m0 < a.M; m0 += stride) {
    float xv[2][VEC], dyv[2][VEC], din[2][VEC], mean[2], rstd[2];
    bool ok[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int m = m0 + r;
      ok[r] = m < a.M;
      if (ok[r]) {
        const size_t off = (size_t)m * D + e0;
        VecIO<VEC>::ldf(a.x + off, xv[r]);
        VecIO<VEC>::ldh(a.dy + off, dyv[r]);
        if (a.dx_in) VecIO<VEC>::ldf(a.dx_in + off, din[r]);
        const float2 st = *reinterpret_cast<const float2*>(a.stats + 2 * (size_t)m);
        mean[r] = st.x; rstd[r] = st.y;
      }
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (!ok[r]) continue;   // warp-uniform
      const int m = m0 + r;
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        xv[r][j] = (xv[r][j] - mean[r]) * rstd[r];
        const float t = dyv[r][j] * gam[j];
        s1 += t; s2 = fmaf(t, xv[r][j], s2);
        dgam[j] = fmaf(dyv[r][j], xv[r][j], dgam[j]);
        dbet[j] += dyv[r][j];
      }
      s1 = warp_sum(s1) * invD;
      s2 = warp_sum(s2) * invD;
      float o[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        o[j] = rstd[r] * (dyv[r][j] * gam[j] - s1 - xv[r][j] * s2);
        if (a.dx_in) o[j] += din[r][j];
      }
      const size_t off = (size_t)m * D + e0;
      VecIO<VEC>::stf(a.dx_out + off, o);
      if (a.dxb) {
        const float sc = row_scale(a.rs, m);
#pragma unroll
        for (int j = 0; j < VEC; ++j) o[j] *= sc;
        VecIO<VEC>::sth(a.dxb + off, o);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < VEC; ++j) { sg[warp][e0 + j] = dgam[j]; sb[warp][e0 + j] = dbet[j]; }
  pdl_trigger();   // the row loop is done: the next kernel may start its prologue
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    float g = 0.f, b = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) { g += sg[w][i]; b += sb[w][i]; }
    if (a.dgamma) atomicAdd(a.dgamma + i, g);
    if (a.dbeta) atomicAdd(a.dbeta + i, b);
  }
}

int launch_ln_bwd(const LnBwdArgs& a, cudaStream_t stream) {
  HS_REQUIRE(a.D <= 32 * kLnMaxPerLane, "ln_bwd: D=%d > %d unsupported", a.D, 32 * kLnMaxPerLane);
  if (a.M == 0) return kOk;
  int grid = ceil_div(a.M, 16);
  if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
  // dx_in == dx_out (in-place) is fine: every element is read and written by the same thread
  if (a.D == 256) HS_CHECK_CUDA(launch_pdl(ln_bwd_vec_kernel<8>, dim3(grid), dim3(256), 0, stream, a));
  else if (a.D == 128) HS_CHECK_CUDA(launch_pdl(ln_bwd_vec_kernel<4>, dim3(grid), dim3(256), 0, stream, a));
  else if (a.D == 64) HS_CHECK_CUDA(launch_pdl(ln_bwd_vec_kernel<2>, dim3(grid), dim3(256), 0, stream, a));
  else ln_bwd_kernel<<<grid, 256, 0, stream>>>(a);
  HS_CHECK_LAUNCH("ln_bwd_kernel");
  return kOk;
}

__global__ void __launch_bounds__(256)
scale_cast_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out, int M, int D, RowScale rs) {
  const int64_t total = (int64_t)M * D / 4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)((i * 4) / D);
    const float s = row_scale(rs, m);
    const float4 v = reinterpret_cast<const float4*>(x)[i];
    uint2 o;
    o.x = pack_bf16x2(s * v.x, s * v.y);
    o.y = pack_bf16x2(s * v.z, s * v.w);
    reinterpret_cast<uint2*>(out)[i] = o;
  }
}

int launch_scale_cast(const float* x, __nv_bfloat16* out, int M, int D, RowScale rs, cudaStream_t stream) {
  HS_REQUIRE(D % 4 == 0, "scale_cast: D=%d must be a multiple of 4", D);
  if (M == 0) return kOk;
  int64_t total = (int64_t)M * D / 4;
  int grid = (int)((total + 255) / 256);
  if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
  scale_cast_kernel<<<grid, 256, 0, stream>>>(x, out, M, D, rs);
  HS_CHECK_LAUNCH("scale_cast_kernel");
  return kOk;
}

__global__ void __launch_bounds__(256)
scale_bf16_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out, int64_t n2, const float* __restrict__ scale) {
  const float s = *scale;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) {
    const float2 v = unpack_bf16x2(reinterpret_cast<const uint32_t*>(in)[i]);
    reinterpret_cast<uint32_t*>(out)[i] = pack_bf16x2(s * v.x, s * v.y);
  }
}

int launch_scale_bf16(const __nv_bfloat16* in, __nv_bfloat16* out, int64_t n, const float* scale, cudaStream_t stream) {
  HS_REQUIRE(n % 2 == 0, "scale_bf16: element count must be even");
  if (n == 0) return kOk;
  int64_t n2 = n / 2;
  int grid = (int)((n2 + 255) / 256);
  if (grid > 8 * kNumSMs) grid = 8 * kNumSMs;
  scale_bf16_kernel<<<grid, 256, 0, stream>>>(in, out, n2, scale);
  HS_CHECK_LAUNCH("scale_bf16_kernel");
  return kOk;
}

// ---------------------------------------------------------------------------
// Attention over short token groups.  Reference: Attention.forward,
// Models.py:192-215 -- softmax(q k^T * hd^-0.5) v per (group, head).
// The spatial / spectral / fusion encoders only differ in which of a sample's
// token rows form a group (SeqSpec), so no data is ever regrouped in HBM
// (the reference's einops rearranges at Models.py:553-563 disappear).
// A CTA stages whole samples (q|k|v rows, bf16) in shared memory; one thread
// owns one (token, head) pair.
// ---------------------------------------------------------------------------
template <int HD>
__device__ __forceinline__ void load_head(const __nv_bfloat16* p, float (&v)[HD]) {
#pragma unroll
  for (int i = 0; i < HD; i += 8) {
    const uint4 t = *reinterpret_cast<const uint4*>(p + i);
    const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[i] = a.x; v[i + 1] = a.y; v[i + 2] = b.x; v[i + 3] = b.y; v[i + 4] = c.x; v[i + 5] = c.y; v[i + 6] = d.x; v[i + 7] = d.y;
  }
}
template <int HD>
__device__ __forceinline__ void store_head(__nv_bfloat16* p, const float (&v)[HD]) {
#pragma unroll
  for (int i = 0; i < HD; i += 8) {
    uint4 t;
    t.x = pack_bf16x2(v[i], v[i + 1]); t.y = pack_bf16x2(v[i + 2], v[i + 3]);
    t.z = pack_bf16x2(v[i + 4], v[i + 5]); t.w = pack_bf16x2(v[i + 6], v[i + 7]);
    *reinterpret_cast<uint4*>(p + i) = t;
  }
}

// row (within the sample) of token `j` of the sequence that token-row `w` belongs to
__device__ __forceinline__ void seq_of(const SeqSpec& s, int w, int& base) {
  // tok_step == 1: rows [q*len, (q+1)*len)   (spatial / fusion / decoder)
  // tok_step  > 1: rows {q + j*tok_step}     (spectral)
  if (s.tok_step == 1) base = (w / s.len) * s.seq_step;
  else base = (w % s.tok_step) * s.seq_step;
}

template <int HD>
__global__ void __launch_bounds__(256)
attn_fwd_kernel(AttnArgs a, int spc) {
  extern __shared__ __align__(16) uint8_t smraw[];
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(smraw);
  const int D = a.D, K = a.s.K, H = a.heads;
  const int row3 = 3 * D;
  const float scale_log2 = rsqrtf((float)HD) * 1.4426950408889634f;
  for (int n0 = blockIdx.x * spc; n0 < a.N; n0 += gridDim.x * spc) {
    const int ns = (a.N - n0) < spc ? (a.N - n0) : spc;
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(a.qkv + (size_t)n0 * K * row3);
    const int nvec = ns * K * row3 / 8;
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) reinterpret_cast<uint4*>(sq)[i] = ld_stream_u4(src + i);
    __syncthreads();
    const int items = ns * K * H;
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const int h = it % H;
      const int tw = it / H;          // token row within the staged samples
      const int smp = tw / K, w = tw - smp * K;
      int base; seq_of(a.s, w, base);
      const __nv_bfloat16* srow = sq + (size_t)smp * K * row3;
      float q[HD];
      load_head<HD>(srow + (size_t)w * row3 + h * HD, q);
      float mx = -INFINITY, den = 0.f, o[HD];
#pragma unroll
      for (int i = 0; i < HD; ++i) o[i] = 0.f;
      for (int j = 0; j < a.s.len; ++j) {
        const __nv_bfloat16* kr = srow + (size_t)(base + j * a.s.tok_step) * row3 + D + h * HD;
        float kv[HD];
        load_head<HD>(kr, kv);
        float sc = 0.f;
#pragma unroll
        for (int i = 0; i < HD; ++i) sc = fmaf(q[i], kv[i], sc);
        sc *= scale_log2;
        const float nm = fmaxf(mx, sc);
        const float corr = exp2f(mx - nm);
        const float pj = exp2f(sc - nm);
        load_head<HD>(kr + D, kv);
        den = den * corr + pj;
#pragma unroll
        for (int i = 0; i < HD; ++i) o[i] = fmaf(pj, kv[i], o[i] * corr);
        mx = nm;
      }
      const float inv = 1.0f / den;
#pragma unroll
      for (int i = 0; i < HD; ++i) o[i] *= inv;
      const size_t m = (size_t)(n0 + smp) * K + w;
      store_head<HD>(a.out + m * D + h * HD, o);
      // log-sum-exp of the scaled scores, in log2 units
      if (a.lse) a.lse[m * H + h] = mx + log2f(den);
    }
  }
}

// Backward: dq_i = scale * sum_j dS_ij k_j ; dk_j = scale * sum_i dS_ij q_i ; dv_j = sum_i P_ij dO_i
// with P_ij = exp2(s_ij - lse_i), dS_ij = P_ij (dO_i . v_j - dO_i . O_i).
template <int HD>
__global__ void __launch_bounds__(256)
attn_bwd_kernel(AttnArgs a, int spc) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int D = a.D, K = a.s.K, H = a.heads;
  const int row3 = 3 * D;
  __nv_bfloat16* sq = reinterpret_cast<__nv_bfloat16*>(smraw);         // [spc*K][3D]
  __nv_bfloat16* sdo = sq + (size_t)spc * K * row3;                    // [spc*K][D]
  float* sdelta = reinterpret_cast<float*>(sdo + (size_t)spc * K * D); // [spc*K][H]   dO.O
  float* slse = sdelta + (size_t)spc * K * H;                          // [spc*K][H]
  const float scale = rsqrtf((float)HD);
  const float scale_log2 = scale * 1.4426950408889634f;
  for (int n0 = blockIdx.x * spc; n0 < a.N; n0 += gridDim.x * spc) {
    const int ns = (a.N - n0) < spc ? (a.N - n0) : spc;
    __syncthreads();
    {
      const uint4* src = reinterpret_cast<const uint4*>(a.qkv + (size_t)n0 * K * row3);
      const int nvec = ns * K * row3 / 8;
      for (int i = threadIdx.x; i < nvec; i += blockDim.x) reinterpret_cast<uint4*>(sq)[i] = ld_stream_u4(src + i);
      const uint4* src2 = reinterpret_cast<const uint4*>(a.dout + (size_t)n0 * K * D);
      const int nvec2 = ns * K * D / 8;
      for (int i = threadIdx.x; i < nvec2; i += blockDim.x) reinterpret_cast<uint4*>(sdo)[i] = ld_stream_u4(src2 + i);
    }
    const int items = ns * K * H;
    // delta_i = dO_i . O_i   (O read from global, bf16)
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const int h = it % H, tw = it / H;
      const size_t m = (size_t)n0 * K + tw;
      float ov[HD], dv[HD];
      load_head<HD>(a.out + m * D + h * HD, ov);
      load_head<HD>(a.dout + m * D + h * HD, dv);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < HD; ++i) s = fmaf(ov[i], dv[i], s);
      sdelta[it] = s;
      slse[it] = a.lse[m * H + h];
    }
    __syncthreads();
    for (int it = threadIdx.x; it < items; it += blockDim.x) {
      const int h = it % H, tw = it / H;
      const int smp = tw / K, w = tw - smp * K;
      int base; seq_of(a.s, w, base);
      const __nv_bfloat16* srow = sq + (size_t)smp * K * row3;
      const __nv_bfloat16* drow = sdo + (size_t)smp * K * D;
      const float* dl = sdelta + (size_t)smp * K * H;
      const float* ls = slse + (size_t)smp * K * H;
      // ---- as query i = w: dq
      float q[HD], dO[HD], acc[HD];
      load_head<HD>(srow + (size_t)w * row3 + h * HD, q);
      load_head<HD>(drow + (size_t)w * D + h * HD, dO);
#pragma unroll
      for (int i = 0; i < HD; ++i) acc[i] = 0.f;
      const float lse_i = ls[w * H + h], delta_i = dl[w * H + h];
      for (int j = 0; j < a.s.len; ++j) {
        const int rj = base + j * a.s.tok_step;
        float kv[HD], vv[HD];
        load_head<HD>(srow + (size_t)rj * row3 + D + h * HD, kv);
        load_head<HD>(srow + (size_t)rj * row3 + 2 * D + h * HD, vv);
        float sc = 0.f, dp = 0.f;
#pragma unroll
        for (int i = 0; i < HD; ++i) { sc = fmaf(q[i], kv[i], sc); dp = fmaf(dO[i], vv[i], dp); }
        const float pij = exp2f(sc * scale_log2 - lse_i);
        const float ds = pij * (dp - delta_i) * scale;
#pragma unroll
        for (int i = 0; i < HD; ++i) acc[i] = fmaf(ds, kv[i], acc[i]);
      }
      const size_t m = (size_t)(n0 + smp) * K + w;
      store_head<HD>(a.dqkv + m * row3 + h * HD, acc);
      // ---- as key/value j = w: dk, dv
      float kj[HD], vj[HD], dk[HD], dvv[HD];
      load_head<HD>(srow + (size_t)w * row3 + D + h * HD, kj);
      load_head<HD>(srow + (size_t)w * row3 + 2 * D + h * HD, vj);
#pragma unroll
      for (int i = 0; i < HD; ++i) { dk[i] = 0.f; dvv[i] = 0.f; }
      for (int i2 = 0; i2 < a.s.len; ++i2) {
        const int ri = base + i2 * a.s.tok_step;
        float qi[HD], doi[HD];
        load_head<HD>(srow + (size_t)ri * row3 + h * HD, qi);
        load_head<HD>(drow + (size_t)ri * D + h * HD, doi);
        float sc = 0.f, dp = 0.f;
#pragma unroll
        for (int i = 0; i < HD; ++i) { sc = fmaf(qi[i], kj[i], sc); dp = fmaf(doi[i], vj[i], dp); }
        const float pij = exp2f(sc * scale_log2 - ls[ri * H + h]);
        const float ds = pij * (dp - dl[ri * H + h]) * scale;
#pragma unroll
        for (int i = 0; i < HD; ++i) { dk[i] = fmaf(ds, qi[i], dk[i]); dvv[i] = fmaf(pij, doi[i], dvv[i]); }
      }
      store_head<HD>(a.dqkv + m * row3 + D + h * HD, dk);
      store_head<HD>(a.dqkv + m * row3 + 2 * D + h * HD, dvv);
    }
  }
}

static int attn_check(const AttnArgs& a) {
  HS_REQUIRE(a.heads > 0 && a.D % a.heads == 0, "attention: D=%d not divisible by heads=%d", a.D, a.heads);
  const int hd = a.D / a.heads;
  HS_REQUIRE(hd == 8 || hd == 16 || hd == 32, "attention: head dim %d unsupported (8, 16, 32)", hd);
  HS_REQUIRE(a.s.nseq * a.s.len == a.s.K, "attention: sequences (%d x %d) do not tile the %d token rows", a.s.nseq, a.s.len, a.s.K);
  HS_REQUIRE((a.s.K * 3 * a.D) % 8 == 0, "attention: sample row block must be 16-byte granular");
  return kOk;
}

template <int HD>
static int attn_fwd_launch(const AttnArgs& a, cudaStream_t stream) {
  const size_t per_sample = (size_t)a.s.K * 3 * a.D * 2;
  int spc = (int)((96 * 1024) / per_sample);
  if (spc < 1) spc = 1;
  const int want = ceil_div(a.N, 2 * kNumSMs);   // keep >= 2 CTAs per SM busy
  if (spc > want) spc = want < 1 ? 1 : want;
  const size_t smem = per_sample * spc;
  HS_REQUIRE(smem <= 227 * 1024, "attention: %zu bytes of shared memory needed", smem);
  HS_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = ceil_div(a.N, spc);
  if (grid > 4 * kNumSMs) grid = 4 * kNumSMs;
  attn_fwd_kernel<HD><<<grid, 256, smem, stream>>>(a, spc);
  HS_CHECK_LAUNCH("attn_fwd_kernel");
  return kOk;
}

template <int HD>
static int attn_bwd_launch(const AttnArgs& a, cudaStream_t stream) {
  const size_t per_sample = (size_t)a.s.K * (4 * a.D * 2 + 2 * a.heads * 4);
  int spc = (int)((96 * 1024) / per_sample);
  if (spc < 1) spc = 1;
  const int want = ceil_div(a.N, 2 * kNumSMs);
  if (spc > want) spc = want < 1 ? 1 : want;
  const size_t smem = per_sample * spc;
  HS_REQUIRE(smem <= 227 * 1024, "attention bwd: %zu bytes of shared memory needed", smem);
  HS_CHECK_CUDA(cudaFuncSetAttribute(attn_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = ceil_div(a.N, spc);
  if (grid > 4 * kNumSMs) grid = 4 * kNumSMs;
  attn_bwd_kernel<HD><<<grid, 256, smem, stream>>>(a, spc);
  HS_CHECK_LAUNCH("attn_bwd_kernel");
  return kOk;
}

bool attn_mma_supported(const AttnArgs& a);
int launch_attn_mma_fwd(const AttnArgs& a, cudaStream_t stream);
int launch_attn_mma_bwd(const AttnArgs& a, cudaStream_t stream);

bool attn_small_supported(const AttnArgs& a);
int launch_attn_small(const AttnArgs& a, bool bwd, cudaStream_t stream);

// groups of at most 4 tokens (spectral encoder): one thread per (sequence, head), attn_small.cu
// (HSIMAE_ATTN_SMALL=0 sends them through the block-diagonal mma path instead, for A/B measurements)
static bool attn_use_small(const AttnArgs& a) {
  static const bool off = getenv("HSIMAE_ATTN_SMALL") && atoi(getenv("HSIMAE_ATTN_SMALL")) == 0;
  static const bool force_simt = getenv("HSIMAE_ATTN_SIMT") && atoi(getenv("HSIMAE_ATTN_SIMT")) != 0;
  return !off && !force_simt && attn_small_supported(a);
}

// HSIMAE_ATTN_SIMT=1 forces the CUDA-core kernels (kept for groups longer than 40 tokens and as an A/B checker)
static bool attn_use_mma(const AttnArgs& a) {
  static const bool force_simt = getenv("HSIMAE_ATTN_SIMT") && atoi(getenv("HSIMAE_ATTN_SIMT")) != 0;
  return !force_simt && attn_mma_supported(a);
}

int launch_attn_fwd(const AttnArgs& a, cudaStream_t stream) {
  HS_TRY(attn_check(a));
  if (a.N == 0) return kOk;
  if (attn_use_small(a)) return launch_attn_small(a, false, stream);
  if (attn_use_mma(a)) return launch_attn_mma_fwd(a, stream);
  switch (a.D / a.heads) {
    case 8: return attn_fwd_launch<8>(a, stream);
    case 16: return attn_fwd_launch<16>(a, stream);
    default: return attn_fwd_launch<32>(a, stream);
  }
}

int launch_attn_bwd(const AttnArgs& a, cudaStream_t stream) {
  HS_TRY(attn_check(a));
  HS_REQUIRE(a.lse && a.dout && a.dqkv, "attention bwd: missing buffers");
  if (a.N == 0) return kOk;
  if (attn_use_small(a)) return launch_attn_small(a, true, stream);
  if (attn_use_mma(a)) return launch_attn_mma_bwd(a, stream);
  switch (a.D / a.heads) {
    case 8: return attn_bwd_launch<8>(a, stream);
    case 16: return attn_bwd_launch<16>(a, stream);
    default: return attn_bwd_launch<32>(a, stream);
  }
}

}  // namespace hsimae
