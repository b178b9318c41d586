// Attention over very short groups (len <= 4): the spectral encoder attends over the 2-4 kept spectral groups of one
// spatial position (/root/reference/Models.py:192-215 on the '(b l) t c' regrouping of :554,563).  A 16x16 mma tile is
// 87 % masked padding for such groups, so these run on CUDA cores: one thread owns one (sequence, head) -- all LEN x LEN
// scores, the softmax and P V (backward: dQ, dK, dV) stay in its registers -- while whole samples are staged through
// shared memory with 16-byte cp.async copies and leave with coalesced 16-byte stores.  The kernels are HBM-bound
// (q|k|v in, out + log-sum-exp out; backward: q|k|v, O, dO, lse in, dq|dk|dv out).
#include "kernels.cuh"

namespace hsimae {

namespace {

constexpr int kSmallThreads = 128;
// CTAs per SM (shared-memory and register budget).  Measured, B200, batch 4096 (fwd / bwd us):
//   2 tokens: 6 / 4 CTAs 34 / 79,  4 / 3 CTAs 37.5 / 90.5
//   3 tokens: 4 / 3 CTAs 37.5 / 96 (a 164-byte spill in backward), unconstrained registers 45 / 105, 6 / 4 CTAs 44.5 / 122
constexpr int small_fwd_per_sm(int len) { return len <= 2 ? 6 : 4; }
constexpr int small_bwd_per_sm(int len) { return len <= 2 ? 4 : 3; }

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }
__device__ __forceinline__ float fast_exp2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_log2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

template <int HD>
__device__ __forceinline__ void load_vec(const uint8_t* p, float (&v)[HD]) {
#pragma unroll
  for (int i = 0; i < HD; i += 8) {
    const uint4 t = *reinterpret_cast<const uint4*>(p + i * 2);
    const float2 a = unpack_bf16x2(t.x), b = unpack_bf16x2(t.y), c = unpack_bf16x2(t.z), d = unpack_bf16x2(t.w);
    v[i] = a.x; v[i + 1] = a.y; v[i + 2] = b.x; v[i + 3] = b.y; v[i + 4] = c.x; v[i + 5] = c.y; v[i + 6] = d.x; v[i + 7] = d.y;
  }
}
template <int HD>
__device__ __forceinline__ void store_vec(uint8_t* p, const float (&v)[HD]) {
#pragma unroll
  for (int i = 0; i < HD; i += 8)
    *reinterpret_cast<uint4*>(p + i * 2) = make_uint4(pack_bf16x2(v[i], v[i + 1]), pack_bf16x2(v[i + 2], v[i + 3]),
                                                      pack_bf16x2(v[i + 4], v[i + 5]), pack_bf16x2(v[i + 6], v[i + 7]));
}
template <int HD>
__device__ __forceinline__ float dot(const float (&a)[HD], const float (&b)[HD]) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < HD; ++i) s = fmaf(a[i], b[i], s);
  return s;
}

// ---------------------------------------------------------------------------
// forward: out = softmax(q k^T / sqrt(hd)) v, lse in log2 units of the scaled scores (same convention as attn_mma.cu)
// ---------------------------------------------------------------------------
template <int HD, int LEN>
__global__ void __launch_bounds__(kSmallThreads, small_fwd_per_sm(LEN))
attn_small_fwd_kernel(AttnArgs a, int spc) {
  extern __shared__ __align__(16) uint8_t smraw[];
  pdl_wait();
  const int D = a.D, K = a.s.K, H = a.heads;
  const int pitch = 3 * D * 2 + 16;
  uint8_t* sq = smraw;                                                   // [spc*K][pitch]
  float* slse = reinterpret_cast<float*>(sq + (size_t)spc * K * pitch);  // [spc*K][H]
  const float scale_log2 = rsqrtf((float)HD) * 1.4426950408889634f;
  const int row_vecs = 3 * D / 8, out_vecs = D / 8;
  const int items = a.s.nseq * H;                                        // (sequence, head) pairs of one sample
  const FastDiv fd_row(row_vecs), fd_out(out_vecs), fd_h(H), fd_items(items);

  for (int n0 = blockIdx.x * spc; n0 < a.N; n0 += gridDim.x * spc) {
    const int ns = (a.N - n0) < spc ? (a.N - n0) : spc;
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(a.qkv + (size_t)n0 * K * 3 * D);
    const uint32_t sq_addr = smem_addr(sq);
    for (int i = threadIdx.x; i < ns * K * row_vecs; i += blockDim.x) {
      int r, c; fd_row.divmod(i, r, c);
      cp_async16(sq_addr + (uint32_t)(r * pitch + c * 16), src + i);
    }
    cp_async_wait_all();
    __syncthreads();
    for (int it = threadIdx.x; it < ns * items; it += blockDim.x) {
      int smp, w; fd_items.divmod(it, smp, w);
      int seq, h; fd_h.divmod(w, seq, h);                             // head fastest: a warp reads contiguous 32-byte pieces of a row
      uint8_t* base = sq + (size_t)smp * K * pitch;
      int row[LEN];
#pragma unroll
      for (int i = 0; i < LEN; ++i) row[i] = seq * a.s.seq_step + i * a.s.tok_step;
      float k[LEN][HD], v[LEN][HD];
#pragma unroll
      for (int j = 0; j < LEN; ++j) {
        load_vec<HD>(base + (size_t)row[j] * pitch + (D + h * HD) * 2, k[j]);
        load_vec<HD>(base + (size_t)row[j] * pitch + (2 * D + h * HD) * 2, v[j]);
      }
#pragma unroll
      for (int i = 0; i < LEN; ++i) {
        float q[HD];
        load_vec<HD>(base + (size_t)row[i] * pitch + h * HD * 2, q);
        float s[LEN], mx = -INFINITY;
#pragma unroll
        for (int j = 0; j < LEN; ++j) { s[j] = dot<HD>(q, k[j]) * scale_log2; mx = fmaxf(mx, s[j]); }
        float sum = 0.f;
#pragma unroll
        for (int j = 0; j < LEN; ++j) { s[j] = fast_exp2(s[j] - mx); sum += s[j]; }
        const float inv = __fdividef(1.0f, sum);
        float o[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) {
          float acc = 0.f;
#pragma unroll
          for (int j = 0; j < LEN; ++j) acc = fmaf(s[j], v[j][d], acc);
          o[d] = acc * inv;
        }
        store_vec<HD>(base + (size_t)row[i] * pitch + h * HD * 2, o);      // the q slot of (row, head) is this thread's alone
        slse[(smp * K + row[i]) * H + h] = mx + fast_log2(sum);
      }
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(a.out + (size_t)n0 * K * D);
    for (int i = threadIdx.x; i < ns * K * out_vecs; i += blockDim.x) {
      int r, c; fd_out.divmod(i, r, c);
      dst[i] = *reinterpret_cast<const uint4*>(sq + (size_t)r * pitch + c * 16);
    }
    if (a.lse) {
      float* ldst = a.lse + (size_t)n0 * K * H;
      for (int i = threadIdx.x; i < ns * K * H; i += blockDim.x) ldst[i] = slse[i];
    }
  }
  pdl_trigger();
}

// ---------------------------------------------------------------------------
// backward:  P_ij = exp2(s_ij c - lse_i);  dS_ij = P_ij (dO_i.V_j - dO_i.O_i) / sqrt(hd)
//            dQ_i = sum_j dS_ij K_j;  dK_j = sum_i dS_ij Q_i;  dV_j = sum_i P_ij dO_i
// ---------------------------------------------------------------------------
template <int HD, int LEN>
__global__ void __launch_bounds__(kSmallThreads, small_bwd_per_sm(LEN))
attn_small_bwd_kernel(AttnArgs a, int spc) {
  extern __shared__ __align__(16) uint8_t smraw[];
  pdl_wait();
  const int D = a.D, K = a.s.K, H = a.heads;
  const int pitch = 3 * D * 2 + 16, pitch_o = 2 * D * 2 + 16;
  uint8_t* sq = smraw;                                               // [spc*K][pitch]   q|k|v  -> overwritten with dq|dk|dv
  uint8_t* so = sq + (size_t)spc * K * pitch;                        // [spc*K][pitch_o] O | dO
  float* slse = reinterpret_cast<float*>(so + (size_t)spc * K * pitch_o);   // [spc*K][H]
  const float scale = rsqrtf((float)HD);
  const float scale_log2 = scale * 1.4426950408889634f;
  const int row_vecs = 3 * D / 8, o_vecs = D / 8;
  const int items = a.s.nseq * H;
  const FastDiv fd_row(row_vecs), fd_o(o_vecs), fd_h(H), fd_items(items);

  for (int n0 = blockIdx.x * spc; n0 < a.N; n0 += gridDim.x * spc) {
    const int ns = (a.N - n0) < spc ? (a.N - n0) : spc;
    __syncthreads();
    {
      const uint32_t sq_addr = smem_addr(sq), so_addr = smem_addr(so);
      const uint4* src = reinterpret_cast<const uint4*>(a.qkv + (size_t)n0 * K * 3 * D);
      for (int i = threadIdx.x; i < ns * K * row_vecs; i += blockDim.x) {
        int r, c; fd_row.divmod(i, r, c);
        cp_async16(sq_addr + (uint32_t)(r * pitch + c * 16), src + i);
      }
      const uint4* src_o = reinterpret_cast<const uint4*>(a.out + (size_t)n0 * K * D);
      const uint4* src_d = reinterpret_cast<const uint4*>(a.dout + (size_t)n0 * K * D);
      for (int i = threadIdx.x; i < ns * K * o_vecs; i += blockDim.x) {
        int r, c; fd_o.divmod(i, r, c);
        cp_async16(so_addr + (uint32_t)(r * pitch_o + c * 16), src_o + i);
        cp_async16(so_addr + (uint32_t)(r * pitch_o + D * 2 + c * 16), src_d + i);
      }
      const float* lsrc = a.lse + (size_t)n0 * K * H;
      for (int i = threadIdx.x; i < ns * K * H; i += blockDim.x) slse[i] = lsrc[i];
      cp_async_wait_all();
    }
    __syncthreads();
    for (int it = threadIdx.x; it < ns * items; it += blockDim.x) {
      int smp, w; fd_items.divmod(it, smp, w);
      int seq, h; fd_h.divmod(w, seq, h);
      uint8_t* base = sq + (size_t)smp * K * pitch;
      const uint8_t* obase = so + (size_t)smp * K * pitch_o;
      const float* ls = slse + (size_t)smp * K * H;
      int row[LEN];
#pragma unroll
      for (int i = 0; i < LEN; ++i) row[i] = seq * a.s.seq_step + i * a.s.tok_step;
      float k[LEN][HD], v[LEN][HD], dk[LEN][HD], dv[LEN][HD];
#pragma unroll
      for (int j = 0; j < LEN; ++j) {
        load_vec<HD>(base + (size_t)row[j] * pitch + (D + h * HD) * 2, k[j]);
        load_vec<HD>(base + (size_t)row[j] * pitch + (2 * D + h * HD) * 2, v[j]);
#pragma unroll
        for (int d = 0; d < HD; ++d) { dk[j][d] = 0.f; dv[j][d] = 0.f; }
      }
#pragma unroll
      for (int i = 0; i < LEN; ++i) {
        float q[HD], o[HD], dout[HD];
        load_vec<HD>(base + (size_t)row[i] * pitch + h * HD * 2, q);
        load_vec<HD>(obase + (size_t)row[i] * pitch_o + h * HD * 2, o);
        load_vec<HD>(obase + (size_t)row[i] * pitch_o + (D + h * HD) * 2, dout);
        const float lse = ls[row[i] * H + h], delta = dot<HD>(dout, o);
        float dq[HD];
#pragma unroll
        for (int d = 0; d < HD; ++d) dq[d] = 0.f;
#pragma unroll
        for (int j = 0; j < LEN; ++j) {
          const float p = fast_exp2(fmaf(dot<HD>(q, k[j]), scale_log2, -lse));
          const float ds = p * (dot<HD>(dout, v[j]) - delta) * scale;
#pragma unroll
          for (int d = 0; d < HD; ++d) {
            dq[d] = fmaf(ds, k[j][d], dq[d]);
            dk[j][d] = fmaf(ds, q[d], dk[j][d]);
            dv[j][d] = fmaf(p, dout[d], dv[j][d]);
          }
        }
        store_vec<HD>(base + (size_t)row[i] * pitch + h * HD * 2, dq);     // q_i is not needed again
      }
#pragma unroll
      for (int j = 0; j < LEN; ++j) {
        store_vec<HD>(base + (size_t)row[j] * pitch + (D + h * HD) * 2, dk[j]);
        store_vec<HD>(base + (size_t)row[j] * pitch + (2 * D + h * HD) * 2, dv[j]);
      }
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(a.dqkv + (size_t)n0 * K * 3 * D);
    for (int i = threadIdx.x; i < ns * K * row_vecs; i += blockDim.x) {
      int r, c; fd_row.divmod(i, r, c);
      dst[i] = *reinterpret_cast<const uint4*>(sq + (size_t)r * pitch + c * 16);
    }
  }
  pdl_trigger();
}

template <int HD, int LEN>
int small_fwd(const AttnArgs& a, cudaStream_t stream) {
  const size_t per_sample = (size_t)a.s.K * (3 * a.D * 2 + 16) + (size_t)a.s.K * a.heads * 4;
  int spc = (int)(((216 / small_fwd_per_sm(LEN)) * 1024) / per_sample);
  if (spc < 1) spc = 1;
  const int want = ceil_div(a.N, 2 * small_fwd_per_sm(LEN) * kNumSMs);
  if (spc > want) spc = want < 1 ? 1 : want;
  const size_t smem = per_sample * spc;
  HS_REQUIRE(smem <= 227 * 1024, "attention(small): %zu bytes of shared memory needed", smem);
  HS_CHECK_CUDA(cudaFuncSetAttribute(attn_small_fwd_kernel<HD, LEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = ceil_div(a.N, spc);
  if (grid > 2 * small_fwd_per_sm(LEN) * kNumSMs) grid = 2 * small_fwd_per_sm(LEN) * kNumSMs;
  HS_CHECK_CUDA(launch_pdl(attn_small_fwd_kernel<HD, LEN>, dim3(grid), dim3(kSmallThreads), smem, stream, a, spc));
  HS_CHECK_LAUNCH("attn_small_fwd_kernel");
  return kOk;
}

template <int HD, int LEN>
int small_bwd(const AttnArgs& a, cudaStream_t stream) {
  const size_t per_sample = (size_t)a.s.K * ((3 * a.D * 2 + 16) + (2 * a.D * 2 + 16)) + (size_t)a.s.K * a.heads * 4;
  int spc = (int)(((216 / small_bwd_per_sm(LEN)) * 1024) / per_sample);
  if (spc < 1) spc = 1;
  const int want = ceil_div(a.N, 2 * small_bwd_per_sm(LEN) * kNumSMs);
  if (spc > want) spc = want < 1 ? 1 : want;
  const size_t smem = per_sample * spc;
  HS_REQUIRE(smem <= 227 * 1024, "attention(small) bwd: %zu bytes of shared memory needed", smem);
  HS_CHECK_CUDA(cudaFuncSetAttribute(attn_small_bwd_kernel<HD, LEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = ceil_div(a.N, spc);
  if (grid > 2 * small_bwd_per_sm(LEN) * kNumSMs) grid = 2 * small_bwd_per_sm(LEN) * kNumSMs;
  HS_CHECK_CUDA(launch_pdl(attn_small_bwd_kernel<HD, LEN>, dim3(grid), dim3(kSmallThreads), smem, stream, a, spc));
  HS_CHECK_LAUNCH("attn_small_bwd_kernel");
  return kOk;
}

template <int HD>
int small_dispatch(const AttnArgs& a, bool bwd, cudaStream_t stream) {
  switch (a.s.len) {
    case 1: return bwd ? small_bwd<HD, 1>(a, stream) : small_fwd<HD, 1>(a, stream);
    case 2: return bwd ? small_bwd<HD, 2>(a, stream) : small_fwd<HD, 2>(a, stream);
    case 3: return bwd ? small_bwd<HD, 3>(a, stream) : small_fwd<HD, 3>(a, stream);
    default: return bwd ? small_bwd<HD, 4>(a, stream) : small_fwd<HD, 4>(a, stream);
  }
}

}  // namespace

bool attn_small_supported(const AttnArgs& a) {
  const int hd = a.heads > 0 ? a.D / a.heads : 0;
  return (hd == 8 || hd == 16) && a.s.len >= 1 && a.s.len <= 4 && a.D % 8 == 0;
}

int launch_attn_small(const AttnArgs& a, bool bwd, cudaStream_t stream) {
  return a.D / a.heads == 8 ? small_dispatch<8>(a, bwd, stream) : small_dispatch<16>(a, bwd, stream);
}

}  // namespace hsimae
