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

// 16-byte asynchronous global->shared copies: every thread keeps all of its copies in flight (no register staging)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

constexpr int kMaxNT = 5;        // key tiles of 8  -> groups of at most 40 tokens
// forward CTAs per SM: five samples in flight per SM hide each other's load / store phases better than four
// (spatial 56 -> 51 us, decoder 54 -> 44 us); the backward's 68 KB of staging per sample allows three
#ifndef HSIMAE_ATTN_FWD_PER_SM
#define HSIMAE_ATTN_FWD_PER_SM 5
#endif
#ifndef HSIMAE_ATTN_UNROLL_FWD
#define HSIMAE_ATTN_UNROLL_FWD 1
#endif
#ifndef HSIMAE_ATTN_UNROLL_BWD
#define HSIMAE_ATTN_UNROLL_BWD 1
#endif
#define HS_PRAGMA_(x) _Pragma(#x)
#define HS_UNROLL(n) HS_PRAGMA_(unroll n)
#ifndef HSIMAE_ATTN_BWD_PER_SM
#define HSIMAE_ATTN_BWD_PER_SM 3
#endif
constexpr int kFwdPerSM = HSIMAE_ATTN_FWD_PER_SM;
constexpr int kAttnThreads = 128;   // small CTAs, several per SM: one CTA's load/store phases overlap the others' math

// Geometry of one warp-unit: 16 tile rows and NT*8 columns, both mapped onto token rows of the staged sample.
struct Unit {
  int s, nseq, seq_step, tok_step;
  uint32_t inv;   // ceil(2^16 / s): x / s == (x * inv) >> 16 for the small x (< 64) used here
  bool packed;
  __device__ __forceinline__ int div_s(int x) const { return (int)(((uint32_t)x * inv) >> 16); }
  __device__ __forceinline__ int mod_s(int x) const { return x - div_s(x) * s; }
  int G;          // packed: sequences per tile
  int seq0;       // packed: first sequence of the tile; unpacked: the sequence
  int pos0;       // unpacked: first position of the tile rows
  int NT;         // column tiles

  __device__ __forceinline__ bool row_valid(int r) const {
    return packed ? (r < G * s && seq0 + div_s(r) < nseq) : (pos0 + r < s);
  }
  // Invalid tile rows are clamped onto the first token of the sequence so that their ldmatrix addresses stay inside the
  // sample.  In the forward kernel another warp (the unit that owns that token) may be overwriting that token's Q slot with
  // its O tile at the same time: compute-sanitizer racecheck reports it; the values feed masked scores (madd = -inf) and
  // rows that are never stored, so either version of the bytes gives the same results.
  __device__ __forceinline__ int row_token(int r) const {   // token row inside the sample (clamped when invalid)
    if (!row_valid(r)) return packed ? seq0 * seq_step : seq0 * seq_step;
    return packed ? (seq0 + div_s(r)) * seq_step + mod_s(r) * tok_step : seq0 * seq_step + (pos0 + r) * tok_step;
  }
  __device__ __forceinline__ bool col_valid(int c) const {
    return packed ? (c < G * s && seq0 + div_s(c) < nseq) : (c < s);
  }
  __device__ __forceinline__ int col_token(int c) const {
    if (!col_valid(c)) return seq0 * seq_step;
    return packed ? (seq0 + div_s(c)) * seq_step + mod_s(c) * tok_step : seq0 * seq_step + c * tok_step;
  }
  __device__ __forceinline__ bool pair_valid(int r, int c) const {
    if (!row_valid(r) || !col_valid(c)) return false;
    return packed ? (div_s(r) == div_s(c)) : true;
  }
};

__device__ __forceinline__ int units_per_head(const SeqSpec& q) {
  if (q.len <= 8) { const int G = 16 / q.len; return (q.nseq + G - 1) / G; }
  return q.nseq * ((q.len + 15) / 16);
}

__device__ __forceinline__ Unit make_unit(const SeqSpec& q, int u) {
  Unit t;
  t.s = q.len; t.nseq = q.nseq;
  t.inv = (65536u + (uint32_t)q.len - 1u) / (uint32_t)q.len;
  // row of token `pos` of sequence `seq`:  seq*seq_step + pos*tok_step  (see ln_attn.cu::seq_of)
  t.seq_step = q.seq_step; t.tok_step = q.tok_step;
  t.packed = q.len <= 8;
  if (t.packed) { t.G = 16 / q.len; t.seq0 = u * t.G; t.pos0 = 0; t.NT = 2; }
  else { const int mt = (q.len + 15) / 16; t.G = 1; t.seq0 = u / mt; t.pos0 = (u % mt) * 16; t.NT = (q.len + 7) / 8; }
  return t;
}

// Backward works on GROUPS of units that share their key columns: one packed unit, or the ceil(len / 16) row tiles of one
// (longer) sequence.  dK / dV of a group's keys are accumulated in registers over the group's query units.
__device__ __forceinline__ int groups_per_head(const SeqSpec& q) {
  if (q.len <= 8) { const int G = 16 / q.len; return (q.nseq + G - 1) / G; }
  return q.nseq;
}
__device__ __forceinline__ int units_per_group(const SeqSpec& q) { return q.len <= 8 ? 1 : (q.len + 15) / 16; }

__device__ __forceinline__ float fast_exp2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float fast_log2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Per-lane view of a unit, computed ONCE per unit and reused for every sample / head: byte offsets of the rows this
// lane addresses in each ldmatrix (for the q|k|v pitch and for the dO pitch), which of its accumulator elements are
// inside the (block-diagonal) mask, and the token rows of the accumulator rows / columns it owns.
template <int NT>
struct Geo {
  static constexpr int NP = (NT + 1) / 2;
  uint32_t a_off, a_off_o;        // A operand (16 tile rows)
  uint32_t b_off[NP], b_off_o[NP];    // B = X^T operand: column-tile pairs
  uint32_t bc_off[NP], bc_off_o[NP];  // B = X operand (transposed load): k-steps of 16 column tokens
  float madd[NT][4];              // additive mask: 0 where accumulator element (nt, e) is a valid (row, col) pair, -inf elsewhere
  int r_tok[2]; bool r_ok[2];     // accumulator rows g, g+8
  int c_tok[NT][2];               // accumulator columns nt*8 + 2t + {0,1}
};

template <int HD, int NT>
__device__ __forceinline__ Geo<NT> make_geo(const Unit& t, int lane, int pitch, int pitch_o) {
  Geo<NT> q;
  const int mi = lane >> 3, g = lane >> 2, tq = lane & 3;
  const int a_tok = t.row_token((mi & 1) * 8 + (lane & 7));
  const int a_co = HD == 8 ? 0 : (mi >> 1) * 8;
  const int b_co = HD == 8 ? 0 : (mi & 1) * 8;
  const int bc_co = HD == 8 ? 0 : (mi >> 1) * 8;
  q.a_off = (uint32_t)(a_tok * pitch + a_co * 2);
  q.a_off_o = (uint32_t)(a_tok * pitch_o + a_co * 2);
#pragma unroll
  for (int i = 0; i < Geo<NT>::NP; ++i) {
    const int bt = t.col_token((2 * i + (mi >> 1)) * 8 + (lane & 7));
    const int ct = t.col_token(i * 16 + (mi & 1) * 8 + (lane & 7));
    q.b_off[i] = (uint32_t)(bt * pitch + b_co * 2);   q.b_off_o[i] = (uint32_t)(bt * pitch_o + b_co * 2);
    q.bc_off[i] = (uint32_t)(ct * pitch + bc_co * 2); q.bc_off_o[i] = (uint32_t)(ct * pitch_o + bc_co * 2);
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e)
      q.madd[nt][e] = t.pair_valid(g + (e >> 1) * 8, nt * 8 + 2 * tq + (e & 1)) ? 0.f : -INFINITY;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) { q.r_ok[hf] = t.row_valid(g + hf * 8); q.r_tok[hf] = t.row_token(g + hf * 8); }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { q.c_tok[nt][0] = t.col_token(nt * 8 + 2 * tq); q.c_tok[nt][1] = t.col_token(nt * 8 + 2 * tq + 1); }
  return q;
}

// Move a Geo to another row tile of the same sequence (same columns): only the row-dependent members change.
template <int HD, int NT>
__device__ __forceinline__ void geo_set_rows(Geo<NT>& q, const Unit& t, int lane, int pitch, int pitch_o) {
  const int mi = lane >> 3, g = lane >> 2, tq = lane & 3;
  const int a_tok = t.row_token((mi & 1) * 8 + (lane & 7));
  const int a_co = HD == 8 ? 0 : (mi >> 1) * 8;
  q.a_off = (uint32_t)(a_tok * pitch + a_co * 2);
  q.a_off_o = (uint32_t)(a_tok * pitch_o + a_co * 2);
#pragma unroll
  for (int nt = 0; nt < NT; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e)
      q.madd[nt][e] = t.pair_valid(g + (e >> 1) * 8, nt * 8 + 2 * tq + (e & 1)) ? 0.f : -INFINITY;
#pragma unroll
  for (int hf = 0; hf < 2; ++hf) { q.r_ok[hf] = t.row_valid(g + hf * 8); q.r_tok[hf] = t.row_token(g + hf * 8); }
}

// token rows of the KEY columns as accumulator rows: key tile kt (16 columns), accumulator rows g, g + 8
template <int NT>
struct KeyRows {
  static constexpr int KT = (NT + 1) / 2;
  int tok[KT][2]; bool ok[KT][2];
};
template <int NT>
__device__ __forceinline__ KeyRows<NT> make_key_rows(const Unit& t, int lane) {
  KeyRows<NT> k;
  const int g = lane >> 2;
#pragma unroll
  for (int kt = 0; kt < KeyRows<NT>::KT; ++kt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int c = kt * 16 + hf * 8 + g;
      k.ok[kt][hf] = c < NT * 8 && t.col_valid(c);
      k.tok[kt][hf] = t.col_token(c < NT * 8 ? c : 0);
    }
  return k;
}

// transpose of an 8x8 bf16 matrix held as C-fragment pairs (lane: row lane / 4, columns 2 (lane % 4), +1)
__device__ __forceinline__ uint32_t movm_t(uint32_t x) {
  uint32_t y;
  asm("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}

// scores: acc[nt] = rows(A) x cols(B)^T over the head dim.  sA / sB already include the operand's feature column.
//   A fragment (16 rows x 16 k): lane supplies the address of one 8x8 matrix row (ldmatrix.x4)
//   B fragments (B[k = feature][n = token]): one ldmatrix.x4 covers two column tiles
template <int HD, int NT>
__device__ __forceinline__ void tile_scores(uint32_t sA, uint32_t a_off, uint32_t sB, const uint32_t (&b_off)[(NT + 1) / 2],
                                            float (&acc)[NT][4]) {
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
  constexpr int KS = HD <= 16 ? 1 : HD / 16;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
    uint32_t a[4];
    ldsm_x4(sA + a_off + ks * 32, a[0], a[1], a[2], a[3]);
    if (HD == 8) { a[2] = 0u; a[3] = 0u; }   // only the lower k-half exists
#pragma unroll
    for (int i = 0; i < (NT + 1) / 2; ++i) {
      uint32_t b[4];
      ldsm_x4(sB + b_off[i] + ks * 32, b[0], b[1], b[2], b[3]);
      if (HD == 8) { b[1] = 0u; b[3] = 0u; }
      mma16816(acc[2 * i], a, b[0], b[1]);
      if (2 * i + 1 < NT) mma16816(acc[2 * i + 1], a, b[2], b[3]);
    }
  }
}

// out (16 rows x HD) = P(16 x cols) * X(cols x HD) with P given as fp32 C-fragments; sX includes the feature column.
//   B fragments (B[k = token][n = feature]) come from transposed ldmatrix: k-step = 16 column tokens, two feature tiles
template <int HD, int NT>
__device__ __forceinline__ void tile_apply(const float (&p)[NT][4], uint32_t sX, const uint32_t (&bc_off)[(NT + 1) / 2],
                                           float (&out)[HD / 8][4]) {
#pragma unroll
  for (int dn = 0; dn < HD / 8; ++dn) { out[dn][0] = out[dn][1] = out[dn][2] = out[dn][3] = 0.f; }
#pragma unroll
  for (int ks = 0; ks < (NT + 1) / 2; ++ks) {
    uint32_t a[4];
    a[0] = pack_bf16x2(p[2 * ks][0], p[2 * ks][1]);
    a[1] = pack_bf16x2(p[2 * ks][2], p[2 * ks][3]);
    if (2 * ks + 1 < NT) {
      a[2] = pack_bf16x2(p[2 * ks + 1 < NT ? 2 * ks + 1 : 0][0], p[2 * ks + 1 < NT ? 2 * ks + 1 : 0][1]);
      a[3] = pack_bf16x2(p[2 * ks + 1 < NT ? 2 * ks + 1 : 0][2], p[2 * ks + 1 < NT ? 2 * ks + 1 : 0][3]);
    } else { a[2] = 0u; a[3] = 0u; }
#pragma unroll
    for (int dn = 0; dn < HD / 8; dn += 2) {
      uint32_t b[4];
      // HD == 8: a single feature tile; lanes 16..31 re-address tile 0 and their matrices are ignored
      ldsm_x4_t(sX + bc_off[ks] + dn * 16, b[0], b[1], b[2], b[3]);
      mma16816(out[dn], a, b[0], b[1]);
      if (HD > 8) mma16816(out[dn + 1 < HD / 8 ? dn + 1 : dn], a, b[2], b[3]);
    }
  }
}

// write a (16 x HD) C-fragment tile as bf16 into row-major smem (srow includes the feature column)
template <int HD, int NT>
__device__ __forceinline__ void store_tile(const Geo<NT>& q, const float (&v)[HD / 8][4], float s0, float s1, uint8_t* sbase, int pitch,
                                           int col0, int lane) {
  const int tq = lane & 3;
#pragma unroll
  for (int half = 0; half < 2; ++half) {
    if (!q.r_ok[half]) continue;
    uint8_t* row = sbase + (size_t)q.r_tok[half] * pitch + (size_t)col0 * 2;
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
template <int HD, int NT>
__global__ void __launch_bounds__(kAttnThreads, kFwdPerSM)
attn_mma_fwd_kernel(AttnArgs a, int spc, int hgroups) {
  extern __shared__ __align__(16) uint8_t smraw[];
  pdl_wait();
  const int D = a.D, K = a.s.K, H = a.heads;
  const int pitch = 3 * D * 2 + 16;                       // +16 B: ldmatrix rows land in distinct bank groups
  uint8_t* sq = smraw;                                    // [spc*K][pitch]
  float* slse = reinterpret_cast<float*>(sq + (size_t)spc * K * pitch);   // [spc*K][H]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int tq = lane & 3;
  const float scale_log2 = rsqrtf((float)HD) * 1.4426950408889634f;
  const int uph = units_per_head(a.s);
  const int row_vecs = 3 * D / 8;                          // uint4 per token row
  const FastDiv fd_row(row_vecs), fd_out(D / 8), fd_hg(hgroups);
  Geo<NT> q;
  int cur_u = -1;

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
    // work item = (unit, sample, group of heads), unit slowest; every warp takes a contiguous range of items so
    // the per-lane unit geometry is rebuilt only when the unit changes
    const int hpg = H / hgroups;
    const int total = ns * uph * hgroups;
    const int per_warp = (total + nwarps - 1) / nwarps;
    const int w_end = (warp + 1) * per_warp < total ? (warp + 1) * per_warp : total;
    for (int w = warp * per_warp; w < w_end; ++w) {
      int wq, hg, u, smp; fd_hg.divmod(w, wq, hg); FastDiv(ns).divmod(wq, u, smp);
      if (u != cur_u) { q = make_geo<HD, NT>(make_unit(a.s, u), lane, pitch, 0); cur_u = u; }
      uint8_t* ssmp = sq + (size_t)smp * K * pitch;
      const uint32_t sb = smem_addr(ssmp);
HS_UNROLL(HSIMAE_ATTN_UNROLL_FWD)
      for (int h = hg * hpg; h < (hg + 1) * hpg; ++h) {
        float sc[NT][4];
        tile_scores<HD, NT>(sb + h * HD * 2, q.a_off, sb + (D + h * HD) * 2, q.b_off, sc);
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
          for (int e = 0; e < 4; ++e) sc[nt][e] = fmaf(sc[nt][e], scale_log2, q.madd[nt][e]);
          mx0 = fmaxf(mx0, fmaxf(sc[nt][0], sc[nt][1]));
          mx1 = fmaxf(mx1, fmaxf(sc[nt][2], sc[nt][3]));
        }
        mx0 = quad_max(mx0); mx1 = quad_max(mx1);
        if (mx0 == -INFINITY) mx0 = 0.f;
        if (mx1 == -INFINITY) mx1 = 0.f;
        float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
          sc[nt][0] = fast_exp2(sc[nt][0] - mx0); sc[nt][1] = fast_exp2(sc[nt][1] - mx0);
          sc[nt][2] = fast_exp2(sc[nt][2] - mx1); sc[nt][3] = fast_exp2(sc[nt][3] - mx1);
          sum0 += sc[nt][0] + sc[nt][1]; sum1 += sc[nt][2] + sc[nt][3];
        }
        sum0 = quad_sum(sum0); sum1 = quad_sum(sum1);
        float o[HD / 8][4];
        tile_apply<HD, NT>(sc, sb + (2 * D + h * HD) * 2, q.bc_off, o);
        // the q slot of these (rows, head) is read by this unit only: reuse it for the output
        __syncwarp();
        store_tile<HD, NT>(q, o, sum0 > 0.f ? __fdividef(1.0f, sum0) : 0.f, sum1 > 0.f ? __fdividef(1.0f, sum1) : 0.f, ssmp, pitch, h * HD, lane);
        if (tq == 0) {
          if (q.r_ok[0]) slse[(smp * K + q.r_tok[0]) * H + h] = mx0 + fast_log2(sum0);
          if (q.r_ok[1]) slse[(smp * K + q.r_tok[1]) * H + h] = mx1 + fast_log2(sum1);
        }
      }
    }
    __syncthreads();
    const int out_vecs = D / 8;
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
// backward, per (group of units sharing their keys, sample, head):
//   for every 16-row query tile of the group:   P_ij = exp2(s_ij*c - lse_i),  dS_ij = P_ij (dO_i.V_j - dO_i.O_i) / sqrt(hd)
//     dQ (tile rows) = dS K;   P^T and dS^T come from the SAME accumulator fragments through movmatrix (8x8 transposes),
//     dK += dS^T Q,  dV += P^T dO  accumulate in registers over the group's query tiles
//   (the first version recomputed K Q^T, V dO^T and the softmax a second time with keys as tile rows: 6 more MMAs, NT*4 exp2 and
//    2 NT*4 shared-memory reads of lse / delta per unit and head)
// ---------------------------------------------------------------------------
template <int HD, int NT>
__global__ void __launch_bounds__(kAttnThreads, HSIMAE_ATTN_BWD_PER_SM)
attn_mma_bwd_kernel(AttnArgs a, int spc, int hgroups) {
  extern __shared__ __align__(16) uint8_t smraw[];
  pdl_wait();
  constexpr int KT = (NT + 1) / 2;
  const int D = a.D, K = a.s.K, H = a.heads;
  const int pitch = 3 * D * 2 + 16;
  const int pitch_o = D * 2 + 16;
  uint8_t* sq = smraw;                                            // [spc*K][pitch]   q|k|v
  uint8_t* sdo = sq + (size_t)spc * K * pitch;                    // [spc*K][pitch_o] dO
  uint8_t* sdq = sdo + (size_t)spc * K * pitch_o;                 // [spc*K][pitch]   dq|dk|dv
  float* sdelta = reinterpret_cast<float*>(sdq + (size_t)spc * K * pitch);  // [spc*K][H]
  float* slse = sdelta + (size_t)spc * K * H;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int tq = lane & 3;
  const float scale = rsqrtf((float)HD);
  const float scale_log2 = scale * 1.4426950408889634f;
  const int gph = groups_per_head(a.s), upg = units_per_group(a.s);
  const int row_vecs = 3 * D / 8, o_vecs = D / 8;
  const FastDiv fd_row(row_vecs), fd_o(o_vecs), fd_h(H), fd_hg(hgroups);
  Geo<NT> q;
  KeyRows<NT> kr;
  int cur_g = -1, cur_i = 0;

  for (int n0 = blockIdx.x * spc; n0 < a.N; n0 += gridDim.x * spc) {
    const int ns = (a.N - n0) < spc ? (a.N - n0) : spc;
    __syncthreads();
    {
      const uint32_t sq_addr = smem_addr(sq), sdo_addr = smem_addr(sdo), so_addr = smem_addr(sdq), sl_addr = smem_addr(slse);
      const uint4* src = reinterpret_cast<const uint4*>(a.qkv + (size_t)n0 * K * 3 * D);
      for (int i = threadIdx.x; i < ns * K * row_vecs; i += blockDim.x) {
        int r, c; fd_row.divmod(i, r, c);
        cp_async16(sq_addr + (uint32_t)(r * pitch + c * 16), src + i);
      }
      const uint4* src2 = reinterpret_cast<const uint4*>(a.dout + (size_t)n0 * K * D);
      const uint4* src3 = reinterpret_cast<const uint4*>(a.out + (size_t)n0 * K * D);   // O parks in the (still unused) dq slot
      for (int i = threadIdx.x; i < ns * K * o_vecs; i += blockDim.x) {
        int r, c; fd_o.divmod(i, r, c);
        cp_async16(sdo_addr + (uint32_t)(r * pitch_o + c * 16), src2 + i);
        cp_async16(so_addr + (uint32_t)(r * pitch + c * 16), src3 + i);
      }
      const float* lsrc = a.lse + (size_t)n0 * K * H;
      for (int i = threadIdx.x; i < ns * K * H; i += blockDim.x) cp_async4(sl_addr + (uint32_t)i * 4u, lsrc + i);
      cp_async_wait_all();
      __syncthreads();
      // delta_i = dO_i . O_i per (row, head)
      for (int it = threadIdx.x; it < ns * K * H; it += blockDim.x) {
        int r, h; fd_h.divmod(it, r, h);
        const uint8_t* po = sdq + (size_t)r * pitch + (size_t)h * HD * 2;
        const uint8_t* pd = sdo + (size_t)r * pitch_o + (size_t)h * HD * 2;
        float sacc = 0.f;
#pragma unroll
        for (int i = 0; i < HD; i += 8) {
          const uint4 x = *reinterpret_cast<const uint4*>(po + i * 2), y = *reinterpret_cast<const uint4*>(pd + i * 2);
          const float2 x0 = unpack_bf16x2(x.x), x1 = unpack_bf16x2(x.y), x2 = unpack_bf16x2(x.z), x3 = unpack_bf16x2(x.w);
          const float2 y0 = unpack_bf16x2(y.x), y1 = unpack_bf16x2(y.y), y2 = unpack_bf16x2(y.z), y3 = unpack_bf16x2(y.w);
          sacc += x0.x * y0.x + x0.y * y0.y + x1.x * y1.x + x1.y * y1.y + x2.x * y2.x + x2.y * y2.y + x3.x * y3.x + x3.y * y3.y;
        }
        sdelta[it] = sacc;
      }
    }
    __syncthreads();
    const int hpg = H / hgroups;
    const int total = ns * gph * hgroups;
    const int per_warp = (total + nwarps - 1) / nwarps;
    const int w_end = (warp + 1) * per_warp < total ? (warp + 1) * per_warp : total;
    for (int w = warp * per_warp; w < w_end; ++w) {
      int wq, hg, grp, smp; fd_hg.divmod(w, wq, hg); FastDiv(ns).divmod(wq, grp, smp);
      if (grp != cur_g) {
        const Unit t0 = make_unit(a.s, grp * upg);
        q = make_geo<HD, NT>(t0, lane, pitch, pitch_o);
        kr = make_key_rows<NT>(t0, lane);
        cur_g = grp; cur_i = 0;
      }
      const uint32_t sb = smem_addr(sq + (size_t)smp * K * pitch);
      const uint32_t sdb = smem_addr(sdo + (size_t)smp * K * pitch_o);
      uint8_t* sdq_s = sdq + (size_t)smp * K * pitch;
      const float* dl = sdelta + (size_t)smp * K * H;
      const float* ls = slse + (size_t)smp * K * H;
HS_UNROLL(HSIMAE_ATTN_UNROLL_BWD)
      for (int h = hg * hpg; h < (hg + 1) * hpg; ++h) {
        const uint32_t cq = sb + h * HD * 2, ck = sb + (D + h * HD) * 2, cv = sb + (2 * D + h * HD) * 2, cdo = sdb + h * HD * 2;
        float dk[KT][HD / 8][4], dv[KT][HD / 8][4];
#pragma unroll
        for (int kt = 0; kt < KT; ++kt)
#pragma unroll
          for (int dn = 0; dn < HD / 8; ++dn)
#pragma unroll
            for (int e = 0; e < 4; ++e) { dk[kt][dn][e] = 0.f; dv[kt][dn][e] = 0.f; }
        for (int i = 0; i < upg; ++i) {
          if (i != cur_i) {   // only sequences longer than 16 tokens have more than one row tile
            geo_set_rows<HD, NT>(q, make_unit(a.s, grp * upg + i), lane, pitch, pitch_o);
            cur_i = i;
          }
          float sc[NT][4], dp[NT][4];
          tile_scores<HD, NT>(cq, q.a_off, ck, q.b_off, sc);            // Q K^T
          tile_scores<HD, NT>(cdo, q.a_off_o, cv, q.b_off, dp);         // dO V^T
          const float lse0 = ls[q.r_tok[0] * H + h], lse1 = ls[q.r_tok[1] * H + h];
          const float de0 = dl[q.r_tok[0] * H + h], de1 = dl[q.r_tok[1] * H + h];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float lse = (e >> 1) ? lse1 : lse0, de = (e >> 1) ? de1 : de0;
              const float p = fast_exp2(fmaf(sc[nt][e], scale_log2, q.madd[nt][e]) - lse);
              sc[nt][e] = p * (dp[nt][e] - de) * scale;     // dS
              dp[nt][e] = p;                                 // P
            }
          }
          {
            float dq[HD / 8][4];
            tile_apply<HD, NT>(sc, ck, q.bc_off, dq);                     // dS K
            store_tile<HD, NT>(q, dq, 1.f, 1.f, sdq_s, pitch, h * HD, lane);
          }
          // dS^T and P^T as A operands: [keys of tile kt] x [the 16 query rows of this tile]
          uint32_t tS[NT][2], tP[NT][2];
#pragma unroll
          for (int nt = 0; nt < NT; ++nt)
#pragma unroll
            for (int hf = 0; hf < 2; ++hf) {
              tS[nt][hf] = movm_t(pack_bf16x2(sc[nt][2 * hf], sc[nt][2 * hf + 1]));
              tP[nt][hf] = movm_t(pack_bf16x2(dp[nt][2 * hf], dp[nt][2 * hf + 1]));
            }
#pragma unroll
          for (int dn = 0; dn < HD / 8; dn += 2) {
            // B operands (k = the tile's 16 query tokens, n = features): Q and dO rows by transposed ldmatrix
            uint32_t bq[4], bo[4];
            ldsm_x4_t(cq + q.a_off + dn * 16, bq[0], bq[1], bq[2], bq[3]);
            ldsm_x4_t(cdo + q.a_off_o + dn * 16, bo[0], bo[1], bo[2], bo[3]);
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) {
              const bool two = 2 * kt + 1 < NT;
              const uint32_t aS[4] = {tS[2 * kt][0], two ? tS[two ? 2 * kt + 1 : 0][0] : 0u, tS[2 * kt][1], two ? tS[two ? 2 * kt + 1 : 0][1] : 0u};
              const uint32_t aP[4] = {tP[2 * kt][0], two ? tP[two ? 2 * kt + 1 : 0][0] : 0u, tP[2 * kt][1], two ? tP[two ? 2 * kt + 1 : 0][1] : 0u};
              mma16816(dk[kt][dn], aS, bq[0], bq[1]);                     // dS^T Q
              mma16816(dv[kt][dn], aP, bo[0], bo[1]);                     // P^T dO
              if (HD > 8) {
                mma16816(dk[kt][dn + 1 < HD / 8 ? dn + 1 : dn], aS, bq[2], bq[3]);
                mma16816(dv[kt][dn + 1 < HD / 8 ? dn + 1 : dn], aP, bo[2], bo[3]);
              }
            }
          }
        }
        // dK, dV of the group's keys: accumulator rows are key columns
#pragma unroll
        for (int kt = 0; kt < KT; ++kt)
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            if (!kr.ok[kt][hf]) continue;
            uint8_t* row = sdq_s + (size_t)kr.tok[kt][hf] * pitch;
#pragma unroll
            for (int dn = 0; dn < HD / 8; ++dn) {
              *reinterpret_cast<uint32_t*>(row + (D + h * HD + dn * 8 + 2 * tq) * 2) = pack_bf16x2(dk[kt][dn][hf * 2], dk[kt][dn][hf * 2 + 1]);
              *reinterpret_cast<uint32_t*>(row + (2 * D + h * HD + dn * 8 + 2 * tq) * 2) = pack_bf16x2(dv[kt][dn][hf * 2], dv[kt][dn][hf * 2 + 1]);
            }
          }
      }
    }
    __syncthreads();
    uint4* dst = reinterpret_cast<uint4*>(a.dqkv + (size_t)n0 * K * 3 * D);
    for (int i = threadIdx.x; i < ns * K * row_vecs; i += blockDim.x) {
      int r, c; fd_row.divmod(i, r, c);
      dst[i] = *reinterpret_cast<const uint4*>(sdq + (size_t)r * pitch + c * 16);
    }
  }
  pdl_trigger();
}

int host_units_per_head(const SeqSpec& q) {
  if (q.len <= 8) { const int G = 16 / q.len; return (q.nseq + G - 1) / G; }
  return q.nseq * ((q.len + 15) / 16);
}
int host_groups_per_head(const SeqSpec& q) {
  if (q.len <= 8) { const int G = 16 / q.len; return (q.nseq + G - 1) / G; }
  return q.nseq;
}
int host_col_tiles(const SeqSpec& q) { return q.len <= 8 ? 2 : (q.len + 7) / 8; }
// smallest divisor of H that gives every warp of the CTA at least two work items
int pick_hgroups(int H, int spc, int uph) {
  for (int d = 1; d <= H; ++d)
    if (H % d == 0 && spc * uph * d >= 2 * (kAttnThreads / 32)) return d;
  return H;
}

template <int HD, int NT>
int fwd_launch_nt(const AttnArgs& a, cudaStream_t stream) {
  const size_t per_sample = (size_t)a.s.K * (3 * a.D * 2 + 16) + (size_t)a.s.K * a.heads * 4;
  int spc = (int)(((216 / kFwdPerSM) * 1024) / per_sample);   // kFwdPerSM CTAs per SM
  if (spc < 1) spc = 1;
  const int want = ceil_div(a.N, 2 * kFwdPerSM * kNumSMs);
  if (spc > want) spc = want < 1 ? 1 : want;
  const size_t smem = per_sample * spc;
  HS_REQUIRE(smem <= 227 * 1024, "attention: %zu bytes of shared memory needed", smem);
  HS_CHECK_CUDA(cudaFuncSetAttribute(attn_mma_fwd_kernel<HD, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = ceil_div(a.N, spc);
  if (grid > 2 * kFwdPerSM * kNumSMs) grid = 2 * kFwdPerSM * kNumSMs;
  HS_CHECK_CUDA(launch_pdl(attn_mma_fwd_kernel<HD, NT>, dim3(grid), dim3(kAttnThreads), smem, stream, a, spc, pick_hgroups(a.heads, spc, host_units_per_head(a.s))));
  HS_CHECK_LAUNCH("attn_mma_fwd_kernel");
  return kOk;
}

template <int HD, int NT>
int bwd_launch_nt(const AttnArgs& a, cudaStream_t stream) {
  const size_t per_sample = (size_t)a.s.K * (2 * (3 * a.D * 2 + 16) + (a.D * 2 + 16)) + (size_t)a.s.K * a.heads * 8;
  int spc = (int)(((216 / HSIMAE_ATTN_BWD_PER_SM) * 1024) / per_sample);   // <= 72 KB per CTA: three CTAs per SM
  if (spc < 1) spc = 1;
  const int want = ceil_div(a.N, 2 * HSIMAE_ATTN_BWD_PER_SM * kNumSMs);
  if (spc > want) spc = want < 1 ? 1 : want;
  const size_t smem = per_sample * spc;
  HS_REQUIRE(smem <= 227 * 1024, "attention bwd: %zu bytes of shared memory needed", smem);
  HS_CHECK_CUDA(cudaFuncSetAttribute(attn_mma_bwd_kernel<HD, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = ceil_div(a.N, spc);
  if (grid > 2 * HSIMAE_ATTN_BWD_PER_SM * kNumSMs) grid = 2 * HSIMAE_ATTN_BWD_PER_SM * kNumSMs;
  HS_CHECK_CUDA(launch_pdl(attn_mma_bwd_kernel<HD, NT>, dim3(grid), dim3(kAttnThreads), smem, stream, a, spc, pick_hgroups(a.heads, spc, host_groups_per_head(a.s))));
  HS_CHECK_LAUNCH("attn_mma_bwd_kernel");
  return kOk;
}

template <int HD>
int fwd_launch(const AttnArgs& a, cudaStream_t stream) {
  switch (host_col_tiles(a.s)) {
    case 2: return fwd_launch_nt<HD, 2>(a, stream);
    case 3: return fwd_launch_nt<HD, 3>(a, stream);
    case 4: return fwd_launch_nt<HD, 4>(a, stream);
    default: return fwd_launch_nt<HD, 5>(a, stream);
  }
}
template <int HD>
int bwd_launch(const AttnArgs& a, cudaStream_t stream) {
  switch (host_col_tiles(a.s)) {
    case 2: return bwd_launch_nt<HD, 2>(a, stream);
    case 3: return bwd_launch_nt<HD, 3>(a, stream);
    case 4: return bwd_launch_nt<HD, 4>(a, stream);
    default: return bwd_launch_nt<HD, 5>(a, stream);
  }
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
