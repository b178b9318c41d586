// Fused transformer-block kernels for sm_100a: several contractions of a block on ONE resident 128-row token tile.
//
//   mlp_fused_kernel : LN2 output -> w1|w3 (tcgen05, TMEM) -> silu(a)*b in registers -> gate tile in shared memory
//                      -> w2 accumulated in TMEM over the hidden chunks -> + bias + residual (+ other branch)
//                      -> LayerNorm of the next consumer.  The [M, 2 Hp] pre-activations never exist and the gate
//                      output [M, Hp] is written once (training) or not at all (inference) and never read back.
//
// Reference being replaced: SwiGLU.forward + the residual add of Block.forward, /root/reference/Models.py:231-232, 305
// (the cuBLAS sgemm + elementwise kernels PyTorch dispatches for them).
//
// Pipeline per CTA (P = 2: CTA pair, cta_group::2 MMAs, each CTA streams HALF of every weight tile; the leader issues):
//   warp 0  TMA producer: the w1|w3 ring (one [128/P x 64] box per slot)
//   warp 1  MMA issuer:   G1(c): ab[c%2] = A * W13[c]^T   (N = 128: 64 hidden units, a|b interleaved by 16; A from smem)
//                         G2(c): out += g[c%2] * W2[:, c]^T (N = d, K = 64; A = the gate tile IN TENSOR MEMORY)
//                         issue order G1(0) G1(1) | G2(0) G1(2) | G2(1) G1(3) | ...  (the gate epilogue of chunk c runs
//                         under G2(c-1) + G1(c+1))
//   warp 2  TMA producer: the w2 ring (one [d/P x 64] box per slot)
//   warp 3  TMA producer: the tile's A block, k-block by k-block as the previous tile's last G1 releases it
//   warps 4-11  gate epilogue: 8 warps per chunk, each group of 4 takes 32 of the 64 hidden units; g = silu(a)*b is
//               written back as packed bf16 into the a|b stage it came from (tcgen05.st) and, in training, to HBM
//   warps 12-19 final epilogue on the [128 x d] accumulator: + bias + residual (+ other branch), next LayerNorm; two
//               warps per TMEM lane quarter, each half of the columns (the accumulator is locked while it is drained)
// Tensor memory: [0, d) output accumulator, [256, 384) and [384, 512) the two a|b stages.
#include "tc_device.cuh"
#include "block_fused.cuh"

namespace hsimae {

namespace ptx {
// A operand in tensor memory (lane = row, one 32-bit column = two consecutive K elements), B from shared memory
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 16-byte read of a kernel-lifetime constant in shared memory (the per-column parameter block): an explicit ld.shared the
// compiler may schedule freely -- through a generic pointer these became LD.E with the address arithmetic redone per load
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
}  // namespace ptx

// -DHSIMAE_TRACE: per CTA and warp role, cycles spent in each pipeline wait (hsimae_debug_trace_fused; tuning instrument)
#ifdef HSIMAE_TRACE
__device__ long long g_trace_fused[256 * 64];
#define FT_DECL long long ft_t0 = clock64(), ft_acc[7] = {0, 0, 0, 0, 0, 0, 0}
#define FT_WAIT(i, ...) { long long ft_t = clock64(); __VA_ARGS__; ft_acc[i] += clock64() - ft_t; }
#define FT_DUMP(slot) { long long* o = g_trace_fused + blockIdx.x * 64 + (slot) * 8; o[0] = clock64() - ft_t0; for (int i_ = 0; i_ < 7; ++i_) o[1 + i_] = ft_acc[i_]; }
#else
#define FT_DECL
#define FT_WAIT(i, ...) { __VA_ARGS__; }
#define FT_DUMP(slot)
#endif

constexpr int kMlpCtrlGateThreads = 384;   // 4 control warps + 8 gate-epilogue warps; then FW final-epilogue warps
constexpr int kMlpChunk = 64;                       // hidden units per chunk (128 interleaved a|b columns)
constexpr uint32_t kMlpAbCol = 256;                 // TMEM: [0, d) output accumulator, 256 + 128 s: a|b stage s

// Final epilogue of the fused MLP on one warp's 32 accumulator rows: x = resid + rs * (acc + b2) [+ resid2] goes out as the
// fp32 residual stream, LayerNorm(x) of the next consumer as bf16 (two-pass statistics from the accumulator, like
// tc_epilogue<kEpiResidLN>).  The accumulator is held for the whole function, so nothing in it should wait on a round trip:
//   * the residual tile comes through the copy engine (three rotating boxes per warp, updated in place and handed back
//     to TMA as the output tile), its rows prefetched into L2 a main loop earlier;
//   * every per-column parameter (bias, gamma, beta) is read from SHARED memory: this kernel leaves no L1, so a
//     warp-uniform __ldg per chunk would be an L2 round trip.
// (Measured alternative: results straight from registers, 16 bytes per lane into 32 different lines per instruction,
//  is slower -- 26k vs 19k cycles per tile.)
struct MlpFinal {
  uint32_t boxes;            // this warp's kFinalBoxes x 4 KB staging boxes
  uint64_t* rbar;            // their "filled" barriers
  uint32_t rphase = 0;
  bool pending = false;
  bool leader = ptx::elect_one();
};

// The 128 accumulator rows are drained by EIGHT warps: two per TMEM lane quarter, each taking half of the columns
// [c0, c0 + cw); the row statistics of the two halves meet in shared memory (named barrier 1 over the 256 threads).
template <int kFinalBoxes, class Wait>
__device__ __forceinline__ void mlp_final_epilogue(const GemmArgs& p, TmemAcc& acc, MlpFinal& st, const CUtensorMap* tmO0,
                                                   const CUtensorMap* tmO1, const CUtensorMap* tmR, uint32_t s_bias,
                                                   uint32_t s_gamma, uint32_t s_beta, float* s_red, bool split, int half,
                                                   int rowi, int m0, int lane, int width, int c0, int cw, int dbg, Wait wait_acc) {
  const int m = m0 + lane;
  const bool valid = m < p.M;
  const bool ln = p.gamma != nullptr;
  const float s = valid ? row_scale(p.rs, m) : 1.0f;
  const int nch = cw >> 5;
  const uint32_t sw = (uint32_t)(lane & 7);
  auto issue = [&](int ch) {
    if (st.leader && !(dbg & 2)) {
      const int b = ch % kFinalBoxes;
      ptx::mbar_expect_tx(st.rbar + b, kStageBufBytes);
      ptx::tma_load_2d_addr(st.boxes + (uint32_t)b * kStageBufBytes, tmR, st.rbar + b, c0 + ch * 32, m0);
    }
  };
  if (st.pending) { if (st.leader) ptx::bulk_wait_read0(); __syncwarp(); st.pending = false; }   // every box is free again
  for (int ch = 0; ch < kFinalBoxes && ch < nch; ++ch) issue(ch);
  wait_acc();
  float sum = 0.f, sumsq = 0.f;
  for (int ch = 0; ch < nch; ++ch) {
    const int b = ch % kFinalBoxes, c = c0 + ch * 32;
    const uint32_t row = st.boxes + (uint32_t)b * kStageBufBytes + (uint32_t)lane * 128u;
    float v[32];
    acc.template load<32>(c, v);
    if (!(dbg & 2)) { ptx::mbar_wait(st.rbar + b, (st.rphase >> b) & 1u); st.rphase ^= 1u << b; }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const uint4 r = ptx::ld_shared_v4(row + ((((uint32_t)j) ^ sw) << 4));
      const float4 bb = ptx::lds_f4(s_bias + (uint32_t)(c + 4 * j) * 4u);
      v[4 * j] = fmaf(s, v[4 * j] + bb.x, __uint_as_float(r.x));
      v[4 * j + 1] = fmaf(s, v[4 * j + 1] + bb.y, __uint_as_float(r.y));
      v[4 * j + 2] = fmaf(s, v[4 * j + 2] + bb.z, __uint_as_float(r.z));
      v[4 * j + 3] = fmaf(s, v[4 * j + 3] + bb.w, __uint_as_float(r.w));
    }
    if (p.resid2 && valid) {
      float r2[32];
      load_f32_row<32>(p.resid2 + (size_t)m * p.ldr + c, r2);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += r2[i];
    }
    if (valid) {
#pragma unroll
      for (int i = 0; i < 32; ++i) { sum += v[i]; sumsq = fmaf(v[i], v[i], sumsq); }
    }
    if (ln) acc.template store<32>(c, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) ptx::st_shared_v4(row + ((((uint32_t)j) ^ sw) << 4), pack4_f32(v + 4 * j));
    ptx::fence_proxy_async();
    __syncwarp();
    if (st.leader && !(dbg & 4)) {
      ptx::tma_store_2d(tmO0, st.boxes + (uint32_t)b * kStageBufBytes, c, m0);
      ptx::bulk_commit();
      // the box used one chunk ago is free once its store has been read: refill it (kFinalBoxes - 1 chunks ahead)
      if (ch >= 1 && ch + kFinalBoxes - 1 < nch) ptx::bulk_wait_read1();
    }
    st.pending = true;
    if (ch >= 1 && ch + kFinalBoxes - 1 < nch) issue(ch + kFinalBoxes - 1);
  }
  if (!ln) return;
  acc.fence_store();
  // Row statistics in ONE pass (sum and sum of squares gathered while the row was drained): a second trip through the
  // accumulator for the centred sum of squares is 8 more TMEM loads inside the section that holds the accumulator.  The rows
  // are residual-stream values (|mean| of the order of the standard deviation): E[x^2] - mean^2 loses nothing in fp32 here.
  // (Over BOTH column halves when the row is split between two warps.)
  const float inv = 1.0f / (float)width;
  if (split) {
    s_red[half * 128 + rowi] = sum;
    s_red[256 + half * 128 + rowi] = sumsq;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    sum += s_red[(half ^ 1) * 128 + rowi];
    sumsq += s_red[256 + (half ^ 1) * 128 + rowi];
    asm volatile("bar.sync 1, 256;" ::: "memory");   // the partials may be overwritten by the next tile
  }
  const float mean = sum * inv;
  const float rstd = rsqrtf(fmaxf(sumsq * inv - mean * mean, 0.f) + p.ln_eps);
  if (half == 0 && valid && p.stats) *reinterpret_cast<float2*>(p.stats + 2 * (size_t)m) = make_float2(mean, rstd);
  // bf16 output boxes of 64 columns alternate between the two staging boxes
  if (st.leader) ptx::bulk_wait_read0();
  __syncwarp();
  for (int k = 0; k * 64 < cw; ++k) {
    const uint32_t row = st.boxes + (uint32_t)(k % kFinalBoxes) * kStageBufBytes + (uint32_t)lane * 128u;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (k * 64 + h * 32 >= cw) break;   // a trailing 32-column box half is clipped at the tensor edge by TMA
      const int c = c0 + k * 64 + h * 32;
      float v[32];
      acc.template load<32>(c, v);
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const float4 g = ptx::lds_f4(s_gamma + (uint32_t)(c + i) * 4u);
        const float4 be = ptx::lds_f4(s_beta + (uint32_t)(c + i) * 4u);
        v[i] = fmaf((v[i] - mean) * rstd, g.x, be.x);
        v[i + 1] = fmaf((v[i + 1] - mean) * rstd, g.y, be.y);
        v[i + 2] = fmaf((v[i + 2] - mean) * rstd, g.z, be.z);
        v[i + 3] = fmaf((v[i + 3] - mean) * rstd, g.w, be.w);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) ptx::st_shared_v4(row + ((((uint32_t)(h * 4 + i)) ^ sw) << 4), pack8_bf16(v + 8 * i));
    }
    ptx::fence_proxy_async();
    __syncwarp();
    if (st.leader) {
      ptx::tma_store_2d(tmO1, st.boxes + (uint32_t)(k % kFinalBoxes) * kStageBufBytes, c0 + k * 64, m0);
      ptx::bulk_commit();
      ptx::bulk_wait_read1();
    }
    __syncwarp();
  }
  st.pending = true;
}

// FW = 4: one final-epilogue warp per TMEM lane quarter, three staging boxes each; FW = 8: two per quarter (each half of
// the columns), two boxes each.
template <int P, int FW>
__global__ void __launch_bounds__(kMlpCtrlGateThreads + 32 * FW, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW13,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmO0,
                 const __grid_constant__ CUtensorMap tmO1, const __grid_constant__ CUtensorMap tmR, MlpFusedArgs p,
                 int n13, int n2, int nch, int m_units, int dbg) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  constexpr int kFinalBoxes = FW == 4 ? 3 : 2;
  constexpr int kMlpThreads = kMlpCtrlGateThreads + 32 * FW;

  const int d = p.tail.N, Hp = p.tail.K;
  const int num_kb = d / kBlockK;
  const uint32_t a_bytes = (uint32_t)num_kb * kATileBytes;
  const uint32_t slot13 = (uint32_t)(2 * kMlpChunk / P) * 128u;   // [128/P rows of w1|w3] x [64 K]
  const uint32_t slot2 = (uint32_t)(d / P) * 128u;                // [d/P rows of w2] x [64 hidden]
  uint8_t* a_res = smem;
  uint8_t* ring13 = a_res + a_bytes;
  uint8_t* ring2 = ring13 + (size_t)n13 * slot13;
  uint8_t* staging = ring2 + (size_t)n2 * slot2;
  uint64_t* full13 = reinterpret_cast<uint64_t*>(staging + FW * kFinalBoxes * kStageBufBytes);
  uint64_t* empty13 = full13 + n13;
  uint64_t* full2 = empty13 + n13;
  uint64_t* empty2 = full2 + n2;
  uint64_t* abfull = empty2 + n2;      // [2] G1 of the chunk retired (both CTAs, multicast commit)
  uint64_t* chunk_done = abfull + 2;   // [2] leader: the gate values of both CTAs are in tensor memory
  uint64_t* a_full = chunk_done + 2;   // [4] k-block kb of the A tile has landed (leader)
  uint64_t* a_empty = a_full + 4;      // [4] the tile's last G1 has read k-block kb (both CTAs)
  uint64_t* out_full = a_empty + 4;
  uint64_t* out_empty = out_full + 1;
  uint64_t* rbar = out_empty + 1;      // [8][kFinalBoxes] residual box filled (final epilogue)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + FW * kFinalBoxes);
  // per-column parameters, copied once: [b13: 2 Hp][b2: d][gamma: d][beta: d]
  float* s_b13 = reinterpret_cast<float*>(reinterpret_cast<uintptr_t>(tmem_slot + 4 + 3) & ~(uintptr_t)15);
  float* s_b2 = s_b13 + 2 * Hp;
  float* s_gamma = s_b2 + d;
  float* s_beta = s_gamma + d;
  float* s_red = s_beta + d;           // [2][2][128] row-statistic partials of the two final-epilogue column halves
  // the final epilogue splits the columns between two warp groups when each half is whole 64-column output boxes
  const bool fsplit = FW == 8 && (d & 127) == 0;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = P == 2 ? (int)ptx::cluster_ctarank() : 0;
  const int unit0 = (int)blockIdx.x / P, unit_step = (int)gridDim.x / P;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmX); ptx::prefetch_tmap(&tmW13); ptx::prefetch_tmap(&tmW2);
    ptx::prefetch_tmap(&tmO0); ptx::prefetch_tmap(&tmO1); ptx::prefetch_tmap(&tmR);
    for (int i = 0; i < FW * kFinalBoxes; ++i) ptx::mbar_init(rbar + i, 1);
    for (int i = 0; i < n13; ++i) { ptx::mbar_init(full13 + i, 1); ptx::mbar_init(empty13 + i, 1); }
    for (int i = 0; i < n2; ++i) { ptx::mbar_init(full2 + i, 1); ptx::mbar_init(empty2 + i, 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(abfull + i, 1); ptx::mbar_init(chunk_done + i, 8 * P); }
    for (int i = 0; i < 4; ++i) { ptx::mbar_init(a_full + i, 1); ptx::mbar_init(a_empty + i, 1); }
    ptx::mbar_init(out_full, 1); ptx::mbar_init(out_empty, (fsplit ? 256 : 128) * P);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (P == 2) { ptx::tmem_alloc_pair(tmem_slot, 512); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  }
  ptx::pdl_trigger();
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (P == 2) ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();   // everything above touched only this CTA's shared / tensor memory
  for (int i = threadIdx.x; i < 2 * Hp; i += kMlpThreads) s_b13[i] = p.b13[i];
  for (int i = threadIdx.x; i < d; i += kMlpThreads) {
    s_b2[i] = p.tail.bias ? p.tail.bias[i] : 0.f;
    s_gamma[i] = p.tail.gamma ? p.tail.gamma[i] : 1.f;
    s_beta[i] = p.tail.gamma ? p.tail.beta[i] : 0.f;
  }
  __syncthreads();

  if (warp == 0) {
    // ---- w1|w3 ring ---------------------------------------------------------------------------------------------
    if (ptx::elect_one()) {
      int slot = 0; uint32_t phase = 0;
      FT_DECL;
      for (int mu = unit0; mu < m_units; mu += unit_step)
        for (int c = 0; c < nch; ++c)
          for (int kb = 0; kb < num_kb; ++kb) {
            FT_WAIT(1, ptx::mbar_wait(empty13 + slot, phase ^ 1u));
            uint8_t* dst = ring13 + (size_t)slot * slot13;
            const int r0 = c * 2 * kMlpChunk + crank * (2 * kMlpChunk / P);
            if constexpr (P == 2) {
              if (crank == 0) ptx::mbar_expect_tx(full13 + slot, 2u * slot13);
              ptx::tma_load_2d_pair(dst, &tmW13, ptx::mapa_rank(ptx::smem_u32(full13 + slot), 0), kb * kBlockK, r0);
            } else {
              ptx::mbar_expect_tx(full13 + slot, slot13);
              ptx::tma_load_2d(dst, &tmW13, full13 + slot, kb * kBlockK, r0);
            }
            if (++slot == n13) { slot = 0; phase ^= 1u; }
          }
      FT_DUMP(0);
    }
  } else if (warp == 2) {
    // ---- w2 ring -----------------------------------------------------------------------------------------------
    if (ptx::elect_one()) {
      int slot = 0; uint32_t phase = 0;
      FT_DECL;
      for (int mu = unit0; mu < m_units; mu += unit_step)
        for (int c = 0; c < nch; ++c) {
          FT_WAIT(0, ptx::mbar_wait(empty2 + slot, phase ^ 1u));
          uint8_t* dst = ring2 + (size_t)slot * slot2;
          if constexpr (P == 2) {
            if (crank == 0) ptx::mbar_expect_tx(full2 + slot, 2u * slot2);
            ptx::tma_load_2d_pair(dst, &tmW2, ptx::mapa_rank(ptx::smem_u32(full2 + slot), 0), c * kMlpChunk, crank * (d / P));
          } else {
            ptx::mbar_expect_tx(full2 + slot, slot2);
            ptx::tma_load_2d(dst, &tmW2, full2 + slot, c * kMlpChunk, 0);
          }
          if (++slot == n2) { slot = 0; phase ^= 1u; }
        }
      FT_DUMP(1);
    }
  } else if (warp == 3) {
    // ---- A tile, one 64-wide k-block at a time: k-block kb of the NEXT tile is requested as soon as the last G1 of
    // the current tile has read it, so the reload hides under the tile's remaining G1 / G2 work ------------------------
    if (ptx::elect_one()) {
      int t = 0;
      FT_DECL;
      for (int mu = unit0; mu < m_units; mu += unit_step, ++t) {
        const int m0 = (mu * P + crank) * kBlockM;   // may lie past the last row: TMA zero-fills the load, clips the stores
        for (int kb = 0; kb < num_kb; ++kb) {
          if (t > 0) FT_WAIT(0, ptx::mbar_wait(a_empty + kb, (uint32_t)(t - 1) & 1u));
          if constexpr (P == 2) {
            if (crank == 0) ptx::mbar_expect_tx(a_full + kb, 2u * kATileBytes);
            ptx::tma_load_2d_pair(a_res + (size_t)kb * kATileBytes, &tmX, ptx::mapa_rank(ptx::smem_u32(a_full + kb), 0), kb * kBlockK, m0);
          } else {
            ptx::mbar_expect_tx(a_full + kb, kATileBytes);
            ptx::tma_load_2d(a_res + (size_t)kb * kATileBytes, &tmX, a_full + kb, kb * kBlockK, m0);
          }
        }
      }
      FT_DUMP(5);
    }
  } else if (warp == 1) {
    // ---- MMA issuer (leader CTA) ----------------------------------------------------------------------------------
    if (crank == 0 && ptx::elect_one()) {
      const uint32_t idesc_ab = make_idesc(2 * kMlpChunk, false, false, kBlockM * P);
      const uint32_t idesc_out = make_idesc(d, false, false, kBlockM * P);
      const uint64_t adesc0 = make_smem_desc(ptx::smem_u32(a_res), 16, 1024);
      const uint64_t b13desc0 = make_smem_desc(ptx::smem_u32(ring13), 16, 1024);
      const uint64_t b2desc0 = make_smem_desc(ptx::smem_u32(ring2), 16, 1024);
      int s13 = 0, s2 = 0; uint32_t ph13 = 0, ph2 = 0;
      int gi0 = 0, t = 0;   // global chunk index of the tile's first chunk, tile counter
      FT_DECL;
      auto commit = [](uint64_t* bar) { if constexpr (P == 2) ptx::umma_commit_pair(bar); else ptx::umma_commit(bar); };
      auto g1 = [&](int c) {
        const int s = (gi0 + c) & 1;
        const uint32_t d_ab = tmem_base + kMlpAbCol + (uint32_t)s * 128u;
        for (int kb = 0; kb < num_kb; ++kb) {
          if (c == 0) FT_WAIT(4, ptx::mbar_wait(a_full + kb, (uint32_t)t & 1u));
          FT_WAIT(0, ptx::mbar_wait(full13 + s13, ph13));
          ptx::tc_fence_after();
          const uint64_t adesc = adesc0 + (uint64_t)(kb * (kATileBytes >> 4));
          const uint64_t bdesc = b13desc0 + (uint64_t)((uint32_t)s13 * (slot13 >> 4));
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            if constexpr (P == 2) ptx::umma_bf16_pair(d_ab, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_ab, (kb | k) != 0 ? 1u : 0u);
            else ptx::umma_bf16(d_ab, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_ab, (kb | k) != 0 ? 1u : 0u);
          }
          commit(empty13 + s13);
          if (c == nch - 1) commit(a_empty + kb);   // nothing of this tile reads k-block kb of A any more
          if (++s13 == n13) { s13 = 0; ph13 ^= 1u; }
        }
        commit(abfull + s);
      };
      // G2: the A operand is the gate tile the epilogue left IN TENSOR MEMORY, inside the a|b stage it was computed from:
      // hidden units [32 j, 32 j + 32) of the chunk as 16 packed bf16x2 columns at stage + 64 j
      auto g2 = [&](int c) {
        const int gi = gi0 + c, s = gi & 1;
        FT_WAIT(1, ptx::mbar_wait(chunk_done + s, (uint32_t)(gi >> 1) & 1u));
        if (c == 0 && t > 0) FT_WAIT(2, ptx::mbar_wait(out_empty, (uint32_t)(t - 1) & 1u));   // the previous tile's accumulator has been drained
        FT_WAIT(3, ptx::mbar_wait(full2 + s2, ph2));
        ptx::tc_fence_after();
        const uint32_t a_t = tmem_base + kMlpAbCol + (uint32_t)s * 128u;
        const uint64_t bdesc = b2desc0 + (uint64_t)((uint32_t)s2 * (slot2 >> 4));
#pragma unroll
        for (int k = 0; k < kMlpChunk / 16; ++k) {
          const uint32_t a_k = a_t + (uint32_t)(k >> 1) * 64u + (uint32_t)(k & 1) * 8u;
          if constexpr (P == 2) ptx::umma_bf16_ts_pair(tmem_base, a_k, bdesc + (uint64_t)(k * 2), idesc_out, (c | k) != 0 ? 1u : 0u);
          else ptx::umma_bf16_ts(tmem_base, a_k, bdesc + (uint64_t)(k * 2), idesc_out, (c | k) != 0 ? 1u : 0u);
        }
        commit(empty2 + s2);
        if (++s2 == n2) { s2 = 0; ph2 ^= 1u; }
        if (c == nch - 1) commit(out_full);
      };
      for (int mu = unit0; mu < m_units; mu += unit_step, ++t) {
        g1(0);
        if (nch > 1) g1(1);
        for (int c = 0; c < nch; ++c) {
          g2(c);
          if (c + 2 < nch) g1(c + 2);
        }
        gi0 += nch;
      }
      FT_DUMP(2);
    }
  } else if (warp < 12) {
    // ---- gate epilogue: g = silu(a) * b for 32 of the chunk's 64 hidden units; the bf16 values go back into the first
    // 16 columns of the 64 a|b columns they came from (the A operand of G2) and, in training, straight to HBM -------------
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const bool leader = ptx::elect_one();
    const uint32_t done_addr0 = P == 2 ? ptx::mapa_rank(ptx::smem_u32(chunk_done), 0) : 0u;
    const uint32_t sb13 = ptx::smem_u32(s_b13);
    int gi = 0;
    FT_DECL;
    for (int mu = unit0; mu < m_units; mu += unit_step) {
      const int m = (mu * P + crank) * kBlockM + q * 32 + lane;
      __nv_bfloat16* grow = (p.g != nullptr && m < p.tail.M) ? p.g + (size_t)m * p.ldg : nullptr;
      for (int c = 0; c < nch; ++c, ++gi) {
        const int s = gi & 1;
        const uint32_t ph = (uint32_t)(gi >> 1) & 1u;
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + kMlpAbCol + (uint32_t)s * 128u + (uint32_t)grp * 64u;
        FT_WAIT(0, ptx::mbar_wait(abfull + s, ph));
        ptx::tc_fence_after();
        float v0[32], v1[32], gp[16];
#ifdef HSIMAE_TRACE
        long long ft_a = clock64();
#endif
        ptx::tmem_ld32(tacc, v0);
        ptx::tmem_ld32(tacc + 32u, v1);
        ptx::tmem_ld_wait();
#ifdef HSIMAE_TRACE
        long long ft_b = clock64(); ft_acc[1] += ft_b - ft_a;
#endif
        const int hbase = c * kMlpChunk + grp * 32;
        const uint32_t bsh = sb13 + (uint32_t)(2 * hbase) * 4u;
        if (hbase + 32 <= Hp) {
          // whole 32-unit block inside the padded hidden width (every chunk but possibly the last)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            float g[16];
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              const float4 ba = ptx::lds_f4(bsh + (uint32_t)(32 * j + i) * 4u), bb = ptx::lds_f4(bsh + (uint32_t)(32 * j + 16 + i) * 4u);
              const float a0 = (j ? v1[i] : v0[i]) + ba.x, a1 = (j ? v1[i + 1] : v0[i + 1]) + ba.y;
              const float a2 = (j ? v1[i + 2] : v0[i + 2]) + ba.z, a3 = (j ? v1[i + 3] : v0[i + 3]) + ba.w;
              g[i] = a0 * ptx::sigmoid_fast(a0) * ((j ? v1[16 + i] : v0[16 + i]) + bb.x);
              g[i + 1] = a1 * ptx::sigmoid_fast(a1) * ((j ? v1[17 + i] : v0[17 + i]) + bb.y);
              g[i + 2] = a2 * ptx::sigmoid_fast(a2) * ((j ? v1[18 + i] : v0[18 + i]) + bb.z);
              g[i + 3] = a3 * ptx::sigmoid_fast(a3) * ((j ? v1[19 + i] : v0[19 + i]) + bb.w);
            }
            const uint4 lo = pack8_bf16(g), hi = pack8_bf16(g + 8);
            gp[8 * j] = __uint_as_float(lo.x); gp[8 * j + 1] = __uint_as_float(lo.y); gp[8 * j + 2] = __uint_as_float(lo.z); gp[8 * j + 3] = __uint_as_float(lo.w);
            gp[8 * j + 4] = __uint_as_float(hi.x); gp[8 * j + 5] = __uint_as_float(hi.y); gp[8 * j + 6] = __uint_as_float(hi.z); gp[8 * j + 7] = __uint_as_float(hi.w);
            if (grow != nullptr) {   // 32 contiguous bytes per thread = one full sector
              uint4* dst = reinterpret_cast<uint4*>(grow + hbase + 16 * j);
              dst[0] = lo; dst[1] = hi;
            }
          }
        } else {
          // tail of the hidden dimension: units beyond the padded width give a finite (zero) operand and are not stored
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int h0 = hbase + j * 16;
            const bool live = h0 < Hp;
            float g[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float a = (j ? v1[i] : v0[i]) + (live ? s_b13[2 * h0 + i] : 0.f), b = (j ? v1[16 + i] : v0[16 + i]) + (live ? s_b13[2 * h0 + 16 + i] : 0.f);
              g[i] = live ? a * ptx::sigmoid_fast(a) * b : 0.f;
            }
            const uint4 lo = pack8_bf16(g), hi = pack8_bf16(g + 8);
            gp[8 * j] = __uint_as_float(lo.x); gp[8 * j + 1] = __uint_as_float(lo.y); gp[8 * j + 2] = __uint_as_float(lo.z); gp[8 * j + 3] = __uint_as_float(lo.w);
            gp[8 * j + 4] = __uint_as_float(hi.x); gp[8 * j + 5] = __uint_as_float(hi.y); gp[8 * j + 6] = __uint_as_float(hi.z); gp[8 * j + 7] = __uint_as_float(hi.w);
            if (grow != nullptr && live) {
              uint4* dst = reinterpret_cast<uint4*>(grow + h0);
              dst[0] = lo; dst[1] = hi;
            }
          }
        }
#ifdef HSIMAE_TRACE
        long long ft_c = clock64(); ft_acc[2] += ft_c - ft_b;
#endif
        ptx::tmem_st16(tacc, gp);
        ptx::tmem_st_wait();
#ifdef HSIMAE_TRACE
        long long ft_d = clock64(); ft_acc[3] += ft_d - ft_c;
#endif
        ptx::tc_fence_before();
        __syncwarp();
        if (leader) {
          if constexpr (P == 2) ptx::mbar_arrive_cluster(done_addr0 + (uint32_t)s * 8u);
          else ptx::mbar_arrive(chunk_done + s);
        }
#ifdef HSIMAE_TRACE
        ft_acc[4] += clock64() - ft_d;
#endif
      }
    }
    if (warp == 4 && lane == 0) FT_DUMP(3);
  } else if (warp < 16 || fsplit) {
    // ---- final epilogue: + bias + residual (+ other branch), LayerNorm of the next consumer --------------------------
    const int q = warp & 3;
    const int half = (warp - 12) >> 2;
    const int cw = fsplit ? d / 2 : d, c0 = half * cw;
    MlpFinal st;
    st.boxes = ptx::smem_u32(staging) + (uint32_t)(warp - 12) * (kFinalBoxes * kStageBufBytes);
    st.rbar = rbar + (warp - 12) * kFinalBoxes;
    const uint32_t out_empty_addr = P == 2 ? ptx::mapa_rank(ptx::smem_u32(out_empty), 0) : 0u;
    int t = 0;
    FT_DECL;
    for (int mu = unit0; mu < m_units; mu += unit_step, ++t) {
      TmemAcc acc{tmem_base + ((uint32_t)(q * 32) << 16)};
      {
        // the tile's residual rows start moving HBM -> L2 a whole main loop before the boxes are requested: the
        // accumulator is held for the duration of pass 1, which then runs at L2 rather than HBM latency
        const int m = (mu * P + crank) * kBlockM + q * 32 + lane;
        if (m < p.tail.M && !(dbg & 1)) ptx::prefetch_l2_bulk(p.tail.resid + (size_t)m * p.tail.ldr + c0, (uint32_t)cw * 4u);
      }
      mlp_final_epilogue<kFinalBoxes>(p.tail, acc, st, &tmO0, &tmO1, &tmR, ptx::smem_u32(s_b2), ptx::smem_u32(s_gamma), ptx::smem_u32(s_beta), s_red, fsplit, half, q * 32 + lane,
                         (mu * P + crank) * kBlockM + q * 32, lane, d, c0, cw, dbg,
                         [&]() { FT_WAIT(0, ptx::mbar_wait(out_full, (uint32_t)t & 1u)); ptx::tc_fence_after(); });
      ptx::tc_fence_before();
      if constexpr (P == 2) ptx::mbar_arrive_cluster(out_empty_addr); else ptx::mbar_arrive(out_empty);
    }
    if (st.pending) { if (st.leader) ptx::bulk_wait_read0(); __syncwarp(); }
    if (warp == 12 && lane == 0) FT_DUMP(4);
  }

  ptx::pdl_trigger();
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (P == 2) ptx::cluster_sync_all();   // nobody leaves while the peer may still read its shared memory / barriers
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (P == 2) ptx::tmem_dealloc_pair(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
namespace {

struct MlpSmem { int n13, n2; size_t bytes; };

// shared memory: [A block][w1|w3 ring][w2 ring][8 x 2 residual / output boxes][barriers][parameters]
bool mlp_smem_plan(int d, int Hp, int P, int FW, MlpSmem* out) {
  const int num_kb = d / kBlockK;
  const int a_bytes = num_kb * kATileBytes;
  const int slot13 = 2 * kMlpChunk / P * 128, slot2 = d / P * 128;
  const int fixed = a_bytes + FW * (FW == 4 ? 3 : 2) * kStageBufBytes + 2048 + (2 * Hp + 3 * d + 512) * 4;
  int n2 = 2;
  int n13 = (kSmemMax - fixed - n2 * slot2) / slot13;
  if (n13 < num_kb) { n2 = 1; n13 = (kSmemMax - fixed - n2 * slot2) / slot13; }
  if (n13 > 3 * num_kb) n13 = 3 * num_kb;   // up to three chunks of w1|w3 in flight
  if (n13 > 12) n13 = 12;
  if (n13 < 1) return false;
  // spend what is left on the w2 ring
  while (n2 < 4 && fixed + n13 * slot13 + (n2 + 1) * slot2 <= kSmemMax) ++n2;
  // HSIMAE_FUSED_MLP_N13 / _N2: ring depths for A/B measurements (must still fit)
  if (getenv("HSIMAE_FUSED_MLP_N13")) n13 = atoi(getenv("HSIMAE_FUSED_MLP_N13"));
  if (getenv("HSIMAE_FUSED_MLP_N2")) n2 = atoi(getenv("HSIMAE_FUSED_MLP_N2"));
  out->n13 = n13; out->n2 = n2;
  out->bytes = (size_t)fixed + (size_t)n13 * slot13 + (size_t)n2 * slot2;
  return out->bytes <= (size_t)kSmemMax;
}

template <int P, int FW>
int launch_mlp_fused(const MlpFusedArgs& a, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    HS_CHECK_CUDA(cudaFuncSetAttribute(mlp_fused_kernel<P, FW>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    configured = true;
  }
  const GemmArgs& t = a.tail;
  const int d = t.N, Hp = t.K;
  MlpSmem sm;
  HS_REQUIRE(mlp_smem_plan(d, Hp, P, FW, &sm), "fused MLP: width %d does not fit in shared memory", d);
  const uint64_t M = (uint64_t)t.M;
  CUtensorMap tmX, tmW13, tmW2, tmO0, tmO1, tmR;
  HS_TRY(get_tmap(a.X, (uint64_t)d, M, (uint64_t)a.ldx, 64, kBlockM, &tmX));
  HS_TRY(get_tmap(a.W13, (uint64_t)d, (uint64_t)(2 * Hp), (uint64_t)a.ldw, 64, (uint32_t)(2 * kMlpChunk / P), &tmW13));
  HS_TRY(get_tmap(t.B, (uint64_t)Hp, (uint64_t)d, (uint64_t)t.ldb, 64, (uint32_t)(d / P), &tmW2));
  HS_TRY(get_tmap(t.out0, (uint64_t)d, M, (uint64_t)t.ld0, 32, 32, &tmO0, 4));
  if (t.gamma) HS_TRY(get_tmap(t.out1, (uint64_t)d, M, (uint64_t)t.ld1, 64, 32, &tmO1)); else tmO1 = tmO0;
  HS_TRY(get_tmap(t.resid, (uint64_t)d, M, (uint64_t)t.ldr, 32, 32, &tmR, 4));
  const int m_units = ceil_div(ceil_div(t.M, kBlockM), P);
  static const int dbg = getenv("HSIMAE_FUSED_MLP_DBG") ? atoi(getenv("HSIMAE_FUSED_MLP_DBG")) : 0;   // timing experiments only (wrong results)
  HS_TRY(launch_clustered(mlp_fused_kernel<P, FW>, pair_grid(m_units, P), kMlpCtrlGateThreads + 32 * FW, sm.bytes, P, stream, tmX, tmW13, tmW2, tmO0, tmO1,
                          tmR, a, sm.n13, sm.n2, ceil_div(Hp, kMlpChunk), m_units, dbg));
  HS_CHECK_LAUNCH("mlp_fused_kernel");
  return kOk;
}

}  // namespace

bool mlp_fused_supported(int d, int Hp) {
  MlpSmem sm;
  return d % 64 == 0 && d >= 64 && d <= 256 && Hp % 16 == 0 && Hp >= 16 && mlp_smem_plan(d, Hp, 2, 4, &sm);
}

int mlp_fused(const MlpFusedArgs& a, cudaStream_t stream) {
  const GemmArgs& t = a.tail;
  HS_REQUIRE(t.M > 0 && mlp_fused_supported(t.N, t.K), "fused MLP: unsupported shape M=%d d=%d Hp=%d", t.M, t.N, t.K);
  HS_REQUIRE(a.X && a.W13 && a.b13 && t.B && t.out0 && t.resid, "fused MLP: null argument");
  HS_REQUIRE(t.ld0 % 4 == 0 && (reinterpret_cast<uintptr_t>(t.out0) & 15) == 0 && (t.gamma == nullptr || (t.out1 != nullptr && t.ld1 % 8 == 0 &&
             (reinterpret_cast<uintptr_t>(t.out1) & 15) == 0)), "fused MLP: outputs must be 16-byte aligned rows");
  HS_REQUIRE(a.ldx % 8 == 0 && a.ldw % 8 == 0 && t.ldb % 8 == 0 && (a.g == nullptr || (a.ldg % 8 == 0 && (reinterpret_cast<uintptr_t>(a.g) & 15) == 0)), "fused MLP: rows must be 16-byte aligned");
  // pairs halve the weight bytes entering each SM (the binding rate of this kernel); a single CTA only when there is
  // one row block or when asked for (HSIMAE_FUSED_MLP_PAIR=0, A/B measurements)
  static const int pair = getenv("HSIMAE_FUSED_MLP_PAIR") ? atoi(getenv("HSIMAE_FUSED_MLP_PAIR")) : 1;
  // HSIMAE_FUSED_MLP_FW = 4 | 8 final-epilogue warps (A/B measurements)
  static const int fw = getenv("HSIMAE_FUSED_MLP_FW") ? atoi(getenv("HSIMAE_FUSED_MLP_FW")) : 4;
  MlpSmem sm;
  const bool single_ok = mlp_smem_plan(t.N, t.K, 1, 4, &sm);
  if (single_ok && (pair == 0 || t.M <= kBlockM)) return launch_mlp_fused<1, 4>(a, stream);
  if (fw == 8) return launch_mlp_fused<2, 8>(a, stream);
  return launch_mlp_fused<2, 4>(a, stream);
}

}  // namespace hsimae

#ifdef HSIMAE_TRACE
extern "C" int hsimae_debug_trace_fused(long long* host_out, int n_ll) {
  return (int)cudaMemcpyFromSymbol(host_out, hsimae::g_trace_fused, (size_t)n_ll * sizeof(long long));
}
#endif
