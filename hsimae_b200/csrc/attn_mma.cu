// Tensor-core attention for the short token groups of HSIMAE (len <= 40).
//
// Reference: Attention.forward, /root/reference/Models.py:192-215 and its autograd
// backward.  Groups are tiny (2..36 tokens, head dim 8/16), so one WARP owns one
// 16-row tile of one (sample, head): QK^T, softmax and PV run on mma.sync
// m16n8k16 bf16 fragments fed by ldmatrix from whole samples staged in shared
// memory.  Groups shorter than 9 tokens are packed several per tile with a
// block-diagonal mask; the spatial / spectral / fusion encoders differ only in
// the row pattern of a group (SeqSpec), never in data layout.
#include "kernels.cuh"

namespace hsimae {

namespace {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kMaxNT = 5;        // key tiles of 8  -> groups of at most 40 tokens
constexpr int kAttnThreads = 256;

// Geometry of one warp-unit: 16 tile rows and NT*8 columns, both mapped onto token rows of the staged sample.
struct Unit {
  int s, nseq, seq_step, tok_step;
  bool packed;
  int G;          // packed: sequences per tile
  int seq0;       // packed: first sequence of the tile; unpacked: the sequence
  int pos0;       // unpacked: first position of the tile rows
  int NT;         // column tiles

  __device__ __forceinline__ bool row_valid(int r) const {
    return packed ? (r < G * s && seq0 + r / s < nseq) : (pos0 + r < s);
  }
  __device__ __forceinline__ int row_token(int r) const {   // token row inside the sample (clamped when invalid)
    if (!row_valid(r)) return packed ? seq0 * seq_step : seq0 * seq_step;
    return packed ? (seq0 + r / s) * seq_step + (r % s) * tok_step : seq0 * seq_step + (pos0 + r) * tok_step;
  }
  __device__ __forceinline__ bool col_valid(int c) const {
    return packed ? (c < G * s && seq0 + c / s < nseq) : (c < s);
  }
  __device__ __forceinline__ int col_token(int c) const {
    if (!col_valid(c)) return seq0 * seq_step;
    return packed ? (seq0 + c / s) * seq_step + (c % s) * tok_step : seq0 * seq_step + c * tok_step;
  }
  __device__ __forceinline__ bool pair_valid(int r, int c) const {
    if (!row_valid(r) || !col_valid(c)) return false;
    return packed ? (r / s == c / s) : true;
  }
};

__device__ __forceinline__ int units_per_head(const SeqSpec& q) {
  if (q.len <= 8) { const int G = 16 / q.len; return (q.nseq + G - 1) / G; }
  return q.nseq * ((q.len + 15) / 16);
}

__device__ __forceinline__ Unit make_unit(const SeqSpec& q, int u) {
  Unit t;
  t.s = q.len; t.nseq = q.nseq;
  // row of token `pos` of sequence `seq`:  seq*seq_step + pos*tok_step  (see ln_attn.cu::seq_of)
  t.seq_step = q.seq_step; t.tok_step = q.tok_step;
  t.packed = q.len <= 8;
  if (t.packed) { t.G = 16 / q.len; t.seq0 = u * t.G; t.pos0 = 0; t.NT = 2; }
  else { const int mt = (q.len + 15) / 16; t.G = 1; t.seq0 = u / mt; t.pos0 = (u % mt) * 16; t.NT = (q.len + 7) / 8; }
  return t;
}

// A fragment (16 rows x 16 k) from row-major smem: lane supplies the address of one 8x8 matrix row.
template <int HD>
__device__ __forceinline__ void load_a(const Unit& t, uint32_t sbase, int pitch, int col0, int kstep, int lane, uint32_t (&a)[4]) {
  const int mi = lane >> 3;
  const int r = (mi & 1) * 8 + (lane & 7);
  int co = kstep * 16 + (mi >> 1) * 8;
  if (HD == 8) co = 0;   // only the lower k-half exists; the upper registers are zeroed below
  const uint32_t addr = sbase + (uint32_t)(t.row_token(r) * pitch + (col0 + co) * 2);
  ldsm_x4(addr, a[0], a[1], a[2], a[3]);
  if (HD == 8) { a[2] = 0u; a[3] = 0u; }
}

// B fragments for C = A * X^T  (B[k = feature][n = token]): two column tiles (nt, nt+1) at k-step `kstep`.
template <int HD>
__device__ __forceinline__ void load_b_rows(const Unit& t, uint32_t sbase, int pitch, int col0, int kstep, int nt, int lane,
                                            uint32_t (&b)[4]) {
  const int mi = lane >> 3;
  const int c = (nt + (mi >> 1)) * 8 + (lane & 7);
  int co = kstep * 16 + (mi & 1) * 8;
  if (HD == 8) co = 0;
  const uint32_t addr = sbase + (uint32_t)(t.col_token(c) * pitch + (col0 + co) * 2);
  ldsm_x4(addr, b[0], b[1], b[2], b[3]);   // (nt: k lo, k hi), (nt+1: k lo, k hi)
  if (HD == 8) { b[1] = 0u; b[3] = 0u; }
}

// B fragments for C = A * X  (B[k = token][n = feature]): k-step = 16 column-tokens, two feature tiles (dn, dn+1).
__device__ __forceinline__ void load_b_cols(const Unit& t, uint32_t sbase, int pitch, int col0, int kstep, int dn, int lane,
                                            uint32_t (&b)[4]) {
  const int mi = lane >> 3;
  const int c = kstep * 16 + (mi & 1) * 8 + (lane & 7);
  const uint32_t addr = sbase + (uint32_t)(t.col_token(c) * pitch + (col0 + (dn + (mi >> 1)) * 8) * 2);
  ldsm_x4_t(addr, b[0], b[1], b[2], b[3]);  // (dn: k lo, k hi), (dn+1: k lo, k hi)
}

// scores for the unit: acc[nt] = rows(A source at colA) x cols(B source at colB)^T over the head dim
template <int HD>
__device__ __forceinline__ void tile_scores(const Unit& t, uint32_t sA, int pitchA, int colA, uint32_t sB, int pitchB, int colB,
                                            int lane, float (&acc)[kMaxNT][4]) {
#pragma unroll
  for (int nt = 0; nt < kMaxNT; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
  constexpr int KS = HD <= 16 ? 1 : HD / 16;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    uint32_t a[4];
    load_a<HD>(t, sA, pitchA, colA, ks, lane, a);
#pragma unroll
    for (int nt = 0; nt < kMaxNT; nt += 2) {
      if (nt < t.NT) {
        uint32_t b[4];
        load_b_rows<HD>(t, sB, pitchB, colB, ks, nt, lane, b);
        mma16816(acc[nt], a, b[0], b[1]);
        if (nt + 1 < kMaxNT && nt + 1 < t.NT) mma16816(acc[nt + 1], a, b[2], b[3]);
      }
    }
  }
}

// out[dn] (16 rows x HD) = P(16 x cols) * X(cols x HD) with P given as fp32 C-fragments
template <int HD>
__device__ __forceinline__ void tile_apply(const Unit& t, const float (&p)[kMaxNT][4], uint32_t sX, int pitchX, int colX, int lane,
                                           float (&out)[HD / 8][4]) {
#pragma unroll
  for (int dn = 0; dn < HD / 8; ++dn) { out[dn][0] = out[dn][1] = out[dn][2] = out[dn][3] = 0.f; }
#pragma unroll
  for (int ks = 0; ks < (kMaxNT + 1) / 2; ++ks) {
    if (2 * ks < t.NT) {
      uint32_t a[4];
      a[0] = pack_bf16x2(p[2 * ks][0], p[2 * ks][1]);
      a[1] = pack_bf16x2(p[2 * ks][2], p[2 * ks][3]);
      if (2 * ks + 1 < kMaxNT && 2 * ks + 1 < t.NT) {
        a[2] = pack_bf16x2(p[2 * ks + 1][0], p[2 * ks + 1][1]);
        a[3] = pack_bf16x2(p[2 * ks + 1][2], p[2 * ks + 1][3]);
      } else { a[2] = 0u; a[3] = 0u; }
#pragma unroll
      for (int dn = 0; dn < HD / 8; dn += 2) {
        uint32_t b[4];
        if (HD == 8) {
          // a single feature tile: lanes 16..31 would address feature tile 1 -> point them at tile 0 and ignore
          const int mi = lane >> 3;
          const int c = ks * 16 + (mi & 1) * 8 + (lane & 7);
          const uint32_t addr = sX + (uint32_t)(t.col_token(c) * pitchX + colX * 2);
          ldsm_x4_t(addr, b[0], b[1], b[2], b[3]);
          mma16816(out[0], a, b[0], b[1]);
        } else {
          load_b_cols(t, sX, pitchX, colX, ks, dn, lane, b);
          mma16816(out[dn], a, b[0], b[1]);
          mma16816(out[dn + 1], a, b[2], b[3]);
        }
      }
    }
  }
}

// write a (16 x HD) C-fragment tile as bf16 into row-major smem
template <int HD>
__device__ __forceinline__ void store_tile(const Unit& t, const float (&v)[HD / 8][4], float s0, float s1, uint8_t* sbase, int pitch,
                                           int col0, int lane) {
  const int g = lane >> 2, tq = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    const int r = g + half * 8;
    if (!t.row_valid(r)) continue;
    uint8_t* row = sbase + (size_t)t.row_token(r) * pitch + (size_t)col0 * 2;
    const float sc = half ? s1 : s0;
#pragma unroll
    for (int dn = 0; dn < HD / 8; ++dn)
      *reinterpret_cast<uint32_t*>(row + (dn * 8 + 2 * tq) * 2) = pack_bf16x2(v[dn][half * 2] * sc, v[dn][half * 2 + 1] * sc);
  }
}

__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// ---------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kAttnThreads)
attn_mma_fwd_kernel(AttnArgs a, int spc) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int D = a.D, K = a.s.K, H = a.heads;
  const int pitch = 3 * D * 2 + 16;                       // +16 B: ldmatrix rows land in distinct bank groups
  uint8_t* sq = smraw;                                    // [spc*K][pitch]
  float* slse = reinterpret_cast<float*>(sq + (size_t)spc * K * pitch);   // [spc*K][H]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const float scale_log2 = rsqrtf((float)HD) * 1.4426950408889634f;
  const int uph = units_per_head(a.s);
  const int row_vecs = 3 * D / 8;                          // uint4 per token row

  for (int n0 = blockIdx.x * spc; n0 < a.N; n0 += gridDim.x * spc) {
    const int ns = (a.N - n0) < spc ? (a.N - n0) : spc;
    __syncthreads();
    const uint4* src = reinterpret_cast<const uint4*>(a.qkv + (size_t)n0 * K * 3 * D);
    for (int i = threadIdx.x; i < ns * K * row_vecs; i += blockDim.x) {
      const int r = i / row_vecs, c = i - r * row_vecs;
      *reinterpret_cast<uint4*>(sq + (size_t)r * pitch + c * 16) = ld_stream_u4(src + i);
    }
    __syncthreads();
    const int total = ns * H * uph;
    for (int w = warp; w < total; w += nwarps) {
      const int u = w % uph, h = (w / uph) % H, smp = w / (uph * H);
      const Unit t = make_unit(a.s, u);
      const uint32_t sb = smem_addr(sq + (size_t)smp * K * pitch);
      float sc[kMaxNT][4];
      tile_scores<HD>(t, sb, pitch, h * HD, sb, pitch, D + h * HD, lane, sc);
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < kMaxNT; ++nt) {
        if (nt < t.NT) {
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int r = g + (e >> 1) * 8, c = nt * 8 + 2 * tq + (e & 1);
            sc[nt][e] = t.pair_valid(r, c) ? sc[nt][e] * scale_log2 : -INFINITY;
          }
          mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
          mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
        }
      }
      mx0 = quad_max(mx0); mx1 = quad_max(mx1);
      if (mx0 == -INFINITY) mx0 = 0.f;
      if (mx1 == -INFINITY) mx1 = 0.f;
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < kMaxNT; ++nt) {
        if (nt < t.NT) {
          sc[nt][0] = exp2f(sc[nt][0] - mx0); sc[nt][1] = exp2f(sc[nt][1] - mx0);
          sc[nt][2] = exp2f(sc[nt][2] - mx1); sc[nt][3] = exp2f(sc[nt][3] - mx1);
          sum0 += sc[nt][0] + sc[nt][1]; sum1 += sc[nt][2] + sc[nt][3];
        }
      }
      sum0 = quad_sum(sum0); sum1 = quad_sum(sum1);
      float o[HD / 8][4];
      tile_apply<HD>(t, sc, sb, pitch, 2 * D + h * HD, lane, o);
      // the q slot of these (rows, head) is read by this unit only: reuse it for the output
      __syncwarp();
      store_tile<HD>(t, o, sum0 > 0.f ? 1.0f / sum0 : 0.f, sum1 > 0.f ? 1.0f / sum1 : 0.f, sq + (size_t)smp * K * pitch, pitch, h * HD, lane);
      if (tq == 0) {
        if (t.row_valid(g)) slse[(smp * K + t.row_token(g)) * H + h] = mx0 + log2f(sum0);
        if (t.row_valid(g + 8)) slse[(smp * K + t.row_token(g + 8)) * H + h] = mx1 + log2f(sum1);
      }
    }
    __syncthreads();
    const int out_vecs = D / 8;
    uint4* dst = reinterpret_cast<uint4*>(a.out + (size_t)n0 * K * D);
    for (int i = threadIdx.x; i < ns * K * out_vecs; i += blockDim.x) {
      const int r = i / out_vecs, c = i - r * out_vecs;
      dst[i] = *reinterpret_cast<const uint4*>(sq + (size_t)r * pitch + c * 16);
    }
    if (a.lse) {
      float* ldst = a.lse + (size_t)n0 * K * H;
      for (int i = threadIdx.x; i < ns * K * H; i += blockDim.x) ldst[i] = slse[i];
    }
  }
}

// ---------------------------------------------------------------------------
// backward:  phase 1 (tile rows = queries): dQ;  phase 2 (tile rows = keys): dK, dV
//   P_ij = exp2(s_ij*c - lse_i),  dS_ij = P_ij (dO_i.V_j - dO_i.O_i) / sqrt(hd)
// ---------------------------------------------------------------------------
template <int HD>
__global__ void __launch_bounds__(kAttnThreads)
attn_mma_bwd_kernel(AttnArgs a, int spc) {
  extern __shared__ __align__(16) uint8_t smraw[];
  const int D = a.D, K = a.s.K, H = a.heads;
  const int pitch = 3 * D * 2 + 16;
  const int pitch_o = D * 2 + 16;
  uint8_t* sq = smraw;                                            // [spc*K][pitch]   q|k|v
  uint8_t* sdo = sq + (size_t)spc * K * pitch;                    // [spc*K][pitch_o] dO
  uint8_t* sdq = sdo + (size_t)spc * K * pitch_o;                 // [spc*K][pitch]   dq|dk|dv
  float* sdelta = reinterpret_cast<float*>(sdq + (size_t)spc * K * pitch);  // [spc*K][H]
  float* slse = sdelta + (size_t)spc * K * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, tq = lane & 3;
  const float scale = rsqrtf((float)HD);
  const float scale_log2 = scale * 1.4426950408889634f;
  const int uph = units_per_head(a.s);
  const int row_vecs = 3 * D / 8, o_vecs = D / 8;

  for (int n0 = blockIdx.x * spc; n0 < a.N; n0 += gridDim.x * spc) {
    const int ns = (a.N - n0) < spc ? (a.N - n0) : spc;
    __syncthreads();
    {
      const uint4* src = reinterpret_cast<const uint4*>(a.qkv + (size_t)n0 * K * 3 * D);
      for (int i = threadIdx.x; i < ns * K * row_vecs; i += blockDim.x) {
        const int r = i / row_vecs, c = i - r * row_vecs;
        *reinterpret_cast<uint4*>(sq + (size_t)r * pitch + c * 16) = ld_stream_u4(src + i);
      }
      const uint4* src2 = reinterpret_cast<const uint4*>(a.dout + (size_t)n0 * K * D);
      for (int i = threadIdx.x; i < ns * K * o_vecs; i += blockDim.x) {
        const int r = i / o_vecs, c = i - r * o_vecs;
        *reinterpret_cast<uint4*>(sdo + (size_t)r * pitch_o + c * 16) = ld_stream_u4(src2 + i);
      }
      // delta_i = dO_i . O_i per (row, head); lse
      for (int it = threadIdx.x; it < ns * K * H; it += blockDim.x) {
        const int h = it % H, r = it / H;
        const size_t m = (size_t)n0 * K + r;
        const __nv_bfloat16* po = a.out + m * D + h * HD;
        const __nv_bfloat16* pd = a.dout + m * D + h * HD;
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < HD; i += 8) {
          const uint4 x = *reinterpret_cast<const uint4*>(po + i), y = *reinterpret_cast<const uint4*>(pd + i);
          const float2 x0 = unpack_bf16x2(x.x), x1 = unpack_bf16x2(x.y), x2 = unpack_bf16x2(x.z), x3 = unpack_bf16x2(x.w);
          const float2 y0 = unpack_bf16x2(y.x), y1 = unpack_bf16x2(y.y), y2 = unpack_bf16x2(y.z), y3 = unpack_bf16x2(y.w);
          s += x0.x * y0.x + x0.y * y0.y + x1.x * y1.x + x1.y * y1.y + x2.x * y2.x + x2.y * y2.y + x3.x * y3.x + x3.y * y3.y;
        }
        sdelta[it] = s;
        slse[it] = a.lse[m * H + h];
      }
    }
    __syncthreads();
    const int total = ns * H * uph;
    for (int w = warp; w < total; w += nwarps) {
      const int u = w % uph, h = (w / uph) % H, smp = w / (uph * H);
      const Unit t = make_unit(a.s, u);
      const uint32_t sb = smem_addr(sq + (size_t)smp * K * pitch);
      const uint32_t sdb = smem_addr(sdo + (size_t)smp * K * pitch_o);
      uint8_t* sdq_s = sdq + (size_t)smp * K * pitch;
      const float* dl = sdelta + (size_t)smp * K * H;
      const float* ls = slse + (size_t)smp * K * H;
      // ---------------- phase 1: rows = queries
      {
        float sc[kMaxNT][4], dp[kMaxNT][4];
        tile_scores<HD>(t, sb, pitch, h * HD, sb, pitch, D + h * HD, lane, sc);           // Q K^T
        tile_scores<HD>(t, sdb, pitch_o, h * HD, sb, pitch, 2 * D + h * HD, lane, dp);    // dO V^T
        const int t0 = t.row_token(g), t1 = t.row_token(g + 8);
        const float lse0 = ls[t0 * H + h], lse1 = ls[t1 * H + h];
        const float de0 = dl[t0 * H + h], de1 = dl[t1 * H + h];
#pragma unroll
        for (int nt = 0; nt < kMaxNT; ++nt) {
          if (nt < t.NT) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int r = g + (e >> 1) * 8, c = nt * 8 + 2 * tq + (e & 1);
              const float lse = (e >> 1) ? lse1 : lse0, de = (e >> 1) ? de1 : de0;
              const float p = t.pair_valid(r, c) ? exp2f(sc[nt][e] * scale_log2 - lse) : 0.f;
              sc[nt][e] = p * (dp[nt][e] - de) * scale;     // dS
            }
          }
        }
        float dq[HD / 8][4];
        tile_apply<HD>(t, sc, sb, pitch, D + h * HD, lane, dq);                              // dS K
        store_tile<HD>(t, dq, 1.f, 1.f, sdq_s, pitch, h * HD, lane);
      }
      // ---------------- phase 2: rows = keys, columns = queries
      {
        float sc[kMaxNT][4], dp[kMaxNT][4];
        tile_scores<HD>(t, sb, pitch, D + h * HD, sb, pitch, h * HD, lane, sc);               // K Q^T
        tile_scores<HD>(t, sb, pitch, 2 * D + h * HD, sdb, pitch_o, h * HD, lane, dp);        // V dO^T
#pragma unroll
        for (int nt = 0; nt < kMaxNT; ++nt) {
          if (nt < t.NT) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int r = g + (e >> 1) * 8, c = nt * 8 + 2 * tq + (e & 1);
              const bool ok = t.pair_valid(r, c);
              const int tc = t.col_token(c);
              const float p = ok ? exp2f(sc[nt][e] * scale_log2 - ls[tc * H + h]) : 0.f;
              dp[nt][e] = p * (dp[nt][e] - dl[tc * H + h]) * scale;   // dS^T
              sc[nt][e] = p;                                           // P^T
            }
          }
        }
        float dk[HD / 8][4], dv[HD / 8][4];
        tile_apply<HD>(t, dp, sb, pitch, h * HD, lane, dk);                                   // dS^T Q
        tile_apply<HD>(t, sc, sdb, pitch_o, h * HD, lane, dv);                                // P^T dO
        store_tile<HD>(t, dk, 1.f, 1.f, sdq_s, pitch, D + h * HD, lane);
        store_tile<HD>(t, dv, 1.f, 1.f, sdq_s, pitch, 2 * D + h * HD, lane);
      }
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(a.dqkv + (size_t)n0 * K * 3 * D);
    for (int i = threadIdx.x; i < ns * K * row_vecs; i += blockDim.x) {
      const int r = i / row_vecs, c = i - r * row_vecs;
      dst[i] = *reinterpret_cast<const uint4*>(sdq + (size_t)r * pitch + c * 16);
    }
  }
}

template <int HD>
int fwd_launch(const AttnArgs& a, cudaStream_t stream) {
  const size_t per_sample = (size_t)a.s.K * (3 * a.D * 2 + 16) + (size_t)a.s.K * a.heads * 4;
  int spc = (int)((100 * 1024) / per_sample);
  if (spc < 1) spc = 1;
  const int want = ceil_div(a.N, 2 * kNumSMs);
  if (spc > want) spc = want < 1 ? 1 : want;
  const size_t smem = per_sample * spc;
  HS_REQUIRE(smem <= 227 * 1024, "attention: %zu bytes of shared memory needed", smem);
  HS_CHECK_CUDA(cudaFuncSetAttribute(attn_mma_fwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = ceil_div(a.N, spc);
  if (grid > 4 * kNumSMs) grid = 4 * kNumSMs;
  attn_mma_fwd_kernel<HD><<<grid, kAttnThreads, smem, stream>>>(a, spc);
  HS_CHECK_LAUNCH("attn_mma_fwd_kernel");
  return kOk;
}

template <int HD>
int bwd_launch(const AttnArgs& a, cudaStream_t stream) {
  const size_t per_sample = (size_t)a.s.K * (2 * (3 * a.D * 2 + 16) + (a.D * 2 + 16)) + (size_t)a.s.K * a.heads * 8;
  int spc = (int)((100 * 1024) / per_sample);
  if (spc < 1) spc = 1;
  const int want = ceil_div(a.N, 2 * kNumSMs);
  if (spc > want) spc = want < 1 ? 1 : want;
  const size_t smem = per_sample * spc;
  HS_REQUIRE(smem <= 227 * 1024, "attention bwd: %zu bytes of shared memory needed", smem);
  HS_CHECK_CUDA(cudaFuncSetAttribute(attn_mma_bwd_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = ceil_div(a.N, spc);
  if (grid > 4 * kNumSMs) grid = 4 * kNumSMs;
  attn_mma_bwd_kernel<HD><<<grid, kAttnThreads, smem, stream>>>(a, spc);
  HS_CHECK_LAUNCH("attn_mma_bwd_kernel");
  return kOk;
}

}  // namespace

bool attn_mma_supported(const AttnArgs& a) {
  const int hd = a.heads > 0 ? a.D / a.heads : 0;
  return (hd == 8 || hd == 16 || hd == 32) && a.s.len <= 8 * kMaxNT && a.D % 8 == 0;
}

int launch_attn_mma_fwd(const AttnArgs& a, cudaStream_t stream) {
  switch (a.D / a.heads) {
    case 8: return fwd_launch<8>(a, stream);
    case 16: return fwd_launch<16>(a, stream);
    default: return fwd_launch<32>(a, stream);
  }
}

int launch_attn_mma_bwd(const AttnArgs& a, cudaStream_t stream) {
  switch (a.D / a.heads) {
    case 8: return bwd_launch<8>(a, stream);
    case 16: return bwd_launch<16>(a, stream);
    default: return bwd_launch<32>(a, stream);
  }
}

}  // namespace hsimae
