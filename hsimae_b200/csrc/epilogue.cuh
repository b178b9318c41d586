// GEMM epilogues, one accumulator ROW per thread.
//
// `Acc` provides   load<W>(col, float (&v)[W])   and   store<W>(col, const float (&v)[W])
// on the thread's accumulator row; both are executed by every thread of the
// warp (tcgen05.ld/st are warp-collective), global traffic is predicated on
// `valid`.  Column offsets are relative to the tile (n0 = first global column).
#pragma once
#include "gemm.cuh"

namespace hsimae {

template <int W>
__device__ __forceinline__ void load_f32_row(const float* p, float (&v)[W]) {
#pragma unroll
  for (int i = 0; i < W; i += 4) {
    float4 t = *reinterpret_cast<const float4*>(p + i);
    v[i] = t.x; v[i + 1] = t.y; v[i + 2] = t.z; v[i + 3] = t.w;
  }
}

// v[i] += vec[i], vec read with 128-bit (warp-uniform, L1-resident) loads
template <int W>
__device__ __forceinline__ void add_vec(const float* __restrict__ vec, float (&v)[W]) {
#pragma unroll
  for (int i = 0; i < W; i += 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(vec + i));
    v[i] += t.x; v[i + 1] += t.y; v[i + 2] += t.z; v[i + 3] += t.w;
  }
}

template <int W>
__device__ __forceinline__ void store_f32_row(float* p, const float (&v)[W]) {
#pragma unroll
  for (int i = 0; i < W; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
}

template <int W>
__device__ __forceinline__ void store_bf16_row(__nv_bfloat16* p, const float (&v)[W]) {
#pragma unroll
  for (int i = 0; i < W; i += 8) {
    uint4 t;
    t.x = pack_bf16x2(v[i], v[i + 1]);
    t.y = pack_bf16x2(v[i + 2], v[i + 3]);
    t.z = pack_bf16x2(v[i + 4], v[i + 5]);
    t.w = pack_bf16x2(v[i + 6], v[i + 7]);
    *reinterpret_cast<uint4*>(p + i) = t;
  }
}

template <int W>
__device__ __forceinline__ void load_bf16_row(const __nv_bfloat16* p, float (&v)[W]) {
#pragma unroll
  for (int i = 0; i < W; i += 8) {
    uint4 t = *reinterpret_cast<const uint4*>(p + i);
    float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[i] = a.x; v[i + 1] = a.y; v[i + 2] = b.x; v[i + 3] = b.y;
    v[i + 4] = c.x; v[i + 5] = c.y; v[i + 6] = d.x; v[i + 7] = d.y;
  }
}

// ---- chunk bodies ---------------------------------------------------------

template <int W, class Acc>
__device__ __forceinline__ void epi_bias_chunk(const GemmArgs& p, Acc& acc, int m, bool valid, int n0, int c, bool f32out) {
  float v[W];
  acc.template load<W>(c, v);
  const int n = n0 + c;
  if (!valid || n >= p.N) return;
  if (p.bias) add_vec<W>(p.bias + n, v);
  if (f32out) store_f32_row<W>(reinterpret_cast<float*>(p.out0) + (size_t)m * p.ld0 + n, v);
  else        store_bf16_row<W>(reinterpret_cast<__nv_bfloat16*>(p.out0) + (size_t)m * p.ld0 + n, v);
}

// pass 1 of the residual epilogue: x = resid + s*(acc+bias) [+resid2]; keep x in the accumulator.
template <int W, class Acc>
__device__ __forceinline__ float epi_resid_chunk(const GemmArgs& p, Acc& acc, int m, bool valid, int c, float s, bool keep) {
  float v[W];
  acc.template load<W>(c, v);
  float sum = 0.f;
  if (valid) {
    float r[W];
    load_f32_row<W>(p.resid + (size_t)m * p.ldr + c, r);
    if (p.bias) add_vec<W>(p.bias + c, v);
#pragma unroll
    for (int i = 0; i < W; ++i) v[i] = fmaf(s, v[i], r[i]);
    if (p.resid2) {
      load_f32_row<W>(p.resid2 + (size_t)m * p.ldr + c, r);
#pragma unroll
      for (int i = 0; i < W; ++i) v[i] += r[i];
    }
    store_f32_row<W>(reinterpret_cast<float*>(p.out0) + (size_t)m * p.ld0 + c, v);
#pragma unroll
    for (int i = 0; i < W; ++i) sum += v[i];
  }
  if (keep) acc.template store<W>(c, v);
  return sum;
}

template <int W, class Acc>
__device__ __forceinline__ float epi_sqdev_chunk(Acc& acc, int c, float mean) {
  float v[W];
  acc.template load<W>(c, v);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < W; ++i) { float d = v[i] - mean; s = fmaf(d, d, s); }
  return s;
}

template <int W, class Acc>
__device__ __forceinline__ void epi_norm_chunk(const GemmArgs& p, Acc& acc, int m, bool valid, int c, float mean, float rstd) {
  float v[W];
  acc.template load<W>(c, v);
  if (!valid) return;
#pragma unroll
  for (int i = 0; i < W; i += 4) {
    const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + c + i));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.beta + c + i));
    v[i] = fmaf((v[i] - mean) * rstd, g.x, b.x);
    v[i + 1] = fmaf((v[i + 1] - mean) * rstd, g.y, b.y);
    v[i + 2] = fmaf((v[i + 2] - mean) * rstd, g.z, b.z);
    v[i + 3] = fmaf((v[i + 3] - mean) * rstd, g.w, b.w);
  }
  store_bf16_row<W>(reinterpret_cast<__nv_bfloat16*>(p.out1) + (size_t)m * p.ld1 + c, v);
}

// SwiGLU forward: chunk of 32 packed columns = 16 a-values then 16 b-values.
template <class Acc>
__device__ __forceinline__ void epi_swiglu_chunk(const GemmArgs& p, Acc& acc, int m, bool valid, int n0, int c) {
  float v[32];
  acc.template load<32>(c, v);
  const int n = n0 + c;
  if (!valid || n >= p.N) return;
  if (p.bias) add_vec<32>(p.bias + n, v);
  store_bf16_row<32>(reinterpret_cast<__nv_bfloat16*>(p.out0) + (size_t)m * p.ld0 + n, v);
  float g[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    // gate on the values the backward pass will see (bf16-rounded pre-activations)
    float a = bf16_round(v[i]), b = bf16_round(v[16 + i]);
    g[i] = silu_f(a) * b;
  }
  store_bf16_row<16>(reinterpret_cast<__nv_bfloat16*>(p.out1) + (size_t)m * p.ld1 + (n >> 1), g);
}

// SwiGLU backward: chunk of 16 hidden columns -> 32 packed output columns.
template <class Acc>
__device__ __forceinline__ void epi_dswiglu_chunk(const GemmArgs& p, Acc& acc, int m, bool valid, int n0, int c) {
  float dg[16];
  acc.template load<16>(c, dg);
  const int h = n0 + c;
  if (!valid || h >= p.N) return;
  float ab[32];
  load_bf16_row<32>(p.ab + (size_t)m * p.ldab + 2 * h, ab);
  float o[32];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float a = ab[i], b = ab[16 + i];
    float sg = __fdividef(1.0f, 1.0f + __expf(-a));
    float si = a * sg;
    o[i] = dg[i] * b * (sg * (1.0f + a * (1.0f - sg)));
    o[16 + i] = dg[i] * si;
  }
  store_bf16_row<32>(reinterpret_cast<__nv_bfloat16*>(p.out0) + (size_t)m * p.ld0 + 2 * h, o);
}

// ---- per-tile driver --------------------------------------------------------
// m: global row of this thread; n0: first global column of the tile; width:
// tile width (multiple of 16).  For kEpiResidLN the tile must span the row.
template <int EPI, class Acc>
__device__ __forceinline__ void run_epilogue(const GemmArgs& p, Acc& acc, int m, int n0, int width) {
  const bool valid = m < p.M;
  if constexpr (EPI == kEpiBiasBf16 || EPI == kEpiBiasF32) {
    int c = 0;
    for (; c + 32 <= width; c += 32) epi_bias_chunk<32>(p, acc, m, valid, n0, c, EPI == kEpiBiasF32);
    for (; c + 16 <= width; c += 16) epi_bias_chunk<16>(p, acc, m, valid, n0, c, EPI == kEpiBiasF32);
  } else if constexpr (EPI == kEpiResidLN) {
    const bool ln = p.gamma != nullptr;
    const float s = valid ? row_scale(p.rs, m) : 1.0f;
    float sum = 0.f;
    int c = 0;
    for (; c + 32 <= width; c += 32) sum += epi_resid_chunk<32>(p, acc, m, valid, c, s, ln);
    for (; c + 16 <= width; c += 16) sum += epi_resid_chunk<16>(p, acc, m, valid, c, s, ln);
    if (ln) {
      acc.fence_store();
      const float inv = 1.0f / (float)width;
      const float mean = sum * inv;
      float sq = 0.f;
      for (c = 0; c + 32 <= width; c += 32) sq += epi_sqdev_chunk<32>(acc, c, mean);
      for (; c + 16 <= width; c += 16) sq += epi_sqdev_chunk<16>(acc, c, mean);
      const float rstd = rsqrtf(sq * inv + p.ln_eps);
      for (c = 0; c + 32 <= width; c += 32) epi_norm_chunk<32>(p, acc, m, valid, c, mean, rstd);
      for (; c + 16 <= width; c += 16) epi_norm_chunk<16>(p, acc, m, valid, c, mean, rstd);
      if (valid && p.stats) *reinterpret_cast<float2*>(p.stats + 2 * (size_t)m) = make_float2(mean, rstd);
    }
  } else if constexpr (EPI == kEpiSwiGLU) {
    for (int c = 0; c + 32 <= width; c += 32) epi_swiglu_chunk(p, acc, m, valid, n0, c);
  } else if constexpr (EPI == kEpiDSwiGLU) {
    for (int c = 0; c + 16 <= width; c += 16) epi_dswiglu_chunk(p, acc, m, valid, n0, c);
  }
}

// accumulator rows in a global fp32 scratch (SIMT checker path)
struct GmemAcc {
  float* row;
  template <int W> __device__ __forceinline__ void load(int c, float (&v)[W]) { load_f32_row<W>(row + c, v); }
  template <int W> __device__ __forceinline__ void store(int c, const float (&v)[W]) { store_f32_row<W>(row + c, v); }
  __device__ __forceinline__ void fence_store() {}
};

}  // namespace hsimae
