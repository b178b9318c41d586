// tcgen05 / TMEM / TMA bf16 GEMMs for sm_100a.
//
//   gemm_tc_kernel       : C[M,N] = A[M,K] * B[N,K]^T with fused epilogues.  Persistent, warp-specialised: warp 0 = TMA
//                          producer, warp 1 = MMA issuer (one thread chosen by elect.sync, accumulator stages in TMEM),
//                          warps 2.. = epilogue groups (one accumulator row per thread, output boxes through TMA store).
//                          Operands are K-major, 128B-swizzled smem tiles of 64 K-elements.
//   gemm_tc_ares_kernel  : the same for K <= 256 with the A block resident in shared memory.
//   gemm_tc_dgate_kernel : two contractions per accumulator stage (recomputed gate pre-activations + gate gradient).
//   wgrad_tc_kernel      : W[Nout,Kin] += Y^T X, reduction over the token dimension, both operands MN-major (the
//                          reduction index is the slow one in memory), split over the reduction across CTAs, fp32
//                          red.add epilogue straight into the gradient arena.
//   Every kernel has a CTA-pair form (P = 2: cluster of two, cta_group::2 MMAs over both shared memories) and is launched
//   with programmatic stream serialisation (prologue overlaps the previous kernel's drain).  What bounds them on B200 is
//   written up in DESIGN.md section 3.
//
// Reference being replaced: the cuBLAS sgemm calls behind nn.Linear forward /
// backward at /root/reference/Models.py:195-216, 232, 579, 600.
#include <cuda.h>
#include <mutex>
#include <unordered_map>
#include <string>
#include <cstdlib>
#include <type_traits>
#include "tc_device.cuh"

namespace hsimae {

// Build with -DHSIMAE_TRACE to accumulate, per CTA and warp role, the cycles spent in each pipeline wait
// (read back with hsimae_debug_trace; tuning instrument, not part of the product build).
#ifdef HSIMAE_TRACE
__device__ long long g_trace[256 * 16];
#define TR_DECL long long tr_t0 = clock64(), tr_acc[3] = {0, 0, 0}
#define TR_WAIT(i, ...) { long long tr_t = clock64(); __VA_ARGS__; tr_acc[i] += clock64() - tr_t; }
#define TR_DUMP(slot) { long long* tr_o = g_trace + blockIdx.x * 16 + (slot) * 4; tr_o[0] = clock64() - tr_t0; tr_o[1] = tr_acc[0]; tr_o[2] = tr_acc[1]; tr_o[3] = tr_acc[2]; }
#else
#define TR_DECL
#define TR_WAIT(i, ...) { __VA_ARGS__; }
#define TR_DUMP(slot)
#endif


// ---------------------------------------------------------------------------
// forward / dgrad kernel
//
// These GEMMs are short in K (64..1376) and wide in output bytes: the epilogue (element-wise math, TMEM round trips,
// output boxes) takes as long as the MMAs of a tile or longer.  The accumulator is therefore split into S TMEM stages
// (512/S columns each) and every stage has its OWN group of four epilogue warps: up to S tiles are being drained
// concurrently while the MMA thread fills the next free stage.
// ---------------------------------------------------------------------------
// P = 2: CTA-pair mode.  The two CTAs of a cluster own consecutive 128-row blocks; each loads its own A rows and HALF
// of the weight tile, the leader (cluster rank 0) issues 256-row cta_group::2 MMAs that read both shared memories and
// write 128 accumulator rows into each CTA's TMEM.  Weight bytes entering each SM per output row are halved: a streaming
// kernel is limited by SM ingress (~45 B/clk/SM), 48 KB per 64-deep k-block of a 128x256 tile vs 32 KB in a pair.
// Barrier ownership: `full` lives in the leader (both producers' TMA bytes land on it), `empty` / `tfull` are
// signalled in both CTAs by multicast tcgen05.commit, `tempty` lives in the leader (both CTAs' epilogues arrive).
template <int EPI, int S, int P>
__global__ void __launch_bounds__(64 + 128 * S, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1,
               const __grid_constant__ CUtensorMap tmR, const __grid_constant__ CUtensorMap tmX, GemmArgs p, int block_n, int stages,
               int n_blks, int num_tiles) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  constexpr uint32_t kAccStride = 512 / S;
  constexpr uint32_t kStagingBytes = 4 * S * StagingBufs<EPI, S>::value * kStageBufBytes;
  constexpr bool kLnBwd = EPI == kEpiLnBwd;                    // dgrad + LayerNorm backward (S == 2 only)
  constexpr bool kTmaResid = (EPI == kEpiResidLN || kLnBwd) && S == 2;

  const uint32_t b_bytes = (uint32_t)(block_n / P) * 128u;   // this CTA's share of the weight tile
  const uint32_t stage_bytes = kATileBytes + b_bytes;
  uint8_t* staging = smem + (size_t)stages * stage_bytes;   // 1024-byte aligned (stage_bytes is a multiple of 1024)
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
  uint64_t* empty = full + stages;
  uint64_t* tfull = empty + stages;
  uint64_t* tempty = tfull + S;
  uint64_t* rbar = tempty + S;                                   // [4 * S][kResidBoxes], only used when kTmaResid
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + (kTmaResid ? 4 * S * kResidBoxes : 0));
  float* gamma_s = reinterpret_cast<float*>(tmem_slot + 4);     // [256] LayerNorm weight (kLnBwd)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  const int crank = P == 2 ? (int)ptx::cluster_ctarank() : 0;
  const int unit0 = (int)blockIdx.x / P, unit_step = (int)gridDim.x / P;   // a "unit" walks tiles of 128*P rows

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    ptx::prefetch_tmap(&tmO0);
    ptx::prefetch_tmap(&tmO1);
    if constexpr (kTmaResid) ptx::prefetch_tmap(&tmR);
    if constexpr (kLnBwd) ptx::prefetch_tmap(&tmX);
    for (int i = 0; i < stages; ++i) { ptx::mbar_init(full + i, 1); ptx::mbar_init(empty + i, 1); }
    for (int i = 0; i < S; ++i) { ptx::mbar_init(tfull + i, 1); ptx::mbar_init(tempty + i, 128 * P); }
    if constexpr (kTmaResid) { for (int i = 0; i < 4 * S * kResidBoxes; ++i) ptx::mbar_init(rbar + i, 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (P == 2) { ptx::tmem_alloc_pair(tmem_slot, 512); ptx::tmem_relinquish_pair(); }
    else { ptx::tmem_alloc(tmem_slot, 512); ptx::tmem_relinquish(); }
  }
  ptx::pdl_trigger();
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (P == 2) ptx::cluster_sync_all();   // the peer's barriers exist before anything is signalled on them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  ptx::pdl_wait();   // everything above touched only this CTA's shared / tensor memory

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      TR_DECL;
      for (int tile = unit0; tile < num_tiles; tile += unit_step) {
        const int m_blk = tile / n_blks, n_blk = tile - m_blk * n_blks;
        const int m0 = (m_blk * P + crank) * kBlockM;
        const int n0 = n_blk * block_n + crank * (block_n / P);
        for (int kb = 0; kb < num_kb; ++kb) {
          TR_WAIT(0, ptx::mbar_wait(empty + stage, phase ^ 1u));
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          if constexpr (P == 2) {
            if (crank == 0) ptx::mbar_expect_tx(full + stage, 2u * stage_bytes);
            const uint32_t bar = ptx::mapa_rank(ptx::smem_u32(full + stage), 0);
            ptx::tma_load_2d_pair(sa, &tmA, bar, kb * kBlockK, m0);
            ptx::tma_load_2d_pair(sa + kATileBytes, &tmB, bar, kb * kBlockK, n0);
          } else {
            ptx::mbar_expect_tx(full + stage, stage_bytes);
            ptx::tma_load_2d(sa, &tmA, full + stage, kb * kBlockK, m0);
            ptx::tma_load_2d(sa + kATileBytes, &tmB, full + stage, kb * kBlockK, n0);
          }
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
      TR_DUMP(0);
    }
  } else if (warp == 1) {
    if (crank == 0 && ptx::elect_one()) {
      const uint32_t idesc = make_idesc(block_n, false, false, kBlockM * P);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      TR_DECL;
      for (int tile = unit0; tile < num_tiles; tile += unit_step, ++it) {
        const int as = it % S;
        const uint32_t aphase = (uint32_t)(it / S) & 1u;
        TR_WAIT(1, ptx::mbar_wait(tempty + as, aphase ^ 1u));
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)as * kAccStride;
        for (int kb = 0; kb < num_kb; ++kb) {
          TR_WAIT(2, ptx::mbar_wait(full + stage, phase));
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = make_smem_desc(sa, 16, 1024);
          const uint64_t bdesc = make_smem_desc(sa + kATileBytes, 16, 1024);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            // advance 16 K-elements = 32 bytes inside the 128B swizzle row
            if constexpr (P == 2) ptx::umma_bf16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
            else ptx::umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (P == 2) ptx::umma_commit_pair(empty + stage); else ptx::umma_commit(empty + stage);
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        if constexpr (P == 2) ptx::umma_commit_pair(tfull + as); else ptx::umma_commit(tfull + as);
      }
      TR_DUMP(1);
    }
  } else {
    const int q = warp & 3;            // TMEM lane quarter this warp may access
    const int grp = (warp - 2) >> 2;   // accumulator stage served by this warp's group
    Stager st{ptx::smem_u32(staging) + (uint32_t)(warp - 2) * (StagingBufs<EPI, S>::value * kStageBufBytes), lane, false};
    if constexpr (kTmaResid) { st.tm_resid = &tmR; st.rbar = rbar + (warp - 2) * kResidBoxes; }
    [[maybe_unused]] LnBwdColAcc col;
    if constexpr (kLnBwd) {
      st.tm_x = &tmX;
#pragma unroll
      for (int j = 0; j < 8; ++j) { col.dg[j] = 0.f; col.db[j] = 0.f; }
      const int e = (warp - 2) * 32 + lane;                        // 0 .. 128 S - 1
      for (int i = e; i < 256; i += 128 * S) gamma_s[i] = i < p.N ? p.gamma[i] : 0.f;
      ptx::named_bar_sync(1, 128 * S);
    }
    const uint32_t tempty_addr = P == 2 ? ptx::mapa_rank(ptx::smem_u32(tempty + grp), 0) : 0u;
    int it = grp;
    TR_DECL;
    for (int tile = unit0 + grp * unit_step; tile < num_tiles; tile += S * unit_step, it += S) {
      const int m_blk = tile / n_blks, n_blk = tile - m_blk * n_blks;
      const uint32_t aphase = (uint32_t)(it / S) & 1u;
      TmemAcc acc{tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)grp * kAccStride};
      const int n0 = n_blk * block_n;
      int width = p.N - n0; if (width > block_n) width = block_n;
      auto wait_acc = [&]() { TR_WAIT(0, ptx::mbar_wait(tfull + grp, aphase)); ptx::tc_fence_after(); };
      if constexpr (kLnBwd) {
        TR_WAIT(1, tc_lnbwd_epilogue(p, acc, st, &tmO0, &tmO1, (m_blk * P + crank) * kBlockM + q * 32, lane, width, ptx::smem_u32(gamma_s), col, wait_acc));
      } else {
        TR_WAIT(1, tc_epilogue<EPI, (S == 2 ? 1 : 0) | (kTmaResid ? 4 : 0)>(p, acc, st, &tmO0, &tmO1, (m_blk * P + crank) * kBlockM + q * 32, lane, n0, width, wait_acc));
      }
      ptx::tc_fence_before();
      if constexpr (P == 2) ptx::mbar_arrive_cluster(tempty_addr); else ptx::mbar_arrive(tempty + grp);
    }
    st.acquire();   // the staging buffers must outlive every store that reads them
    if constexpr (kLnBwd) {
      // gradients of the LayerNorm weight / bias: the warps' column partials meet in the (now idle) staging boxes, one
      // atomic per column and CTA
      constexpr uint32_t kWarpStaging = StagingBufs<EPI, S>::value * kStageBufBytes;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        ptx::st_shared_f32(st.base + (uint32_t)((j * 32 + kLnBwdCol(lane)) * 4), col.dg[j]);
        ptx::st_shared_f32(st.base + 1024u + (uint32_t)((j * 32 + kLnBwdCol(lane)) * 4), col.db[j]);
      }
      ptx::named_bar_sync(1, 128 * S);
      const uint32_t s0 = ptx::smem_u32(staging);
      for (int e = (warp - 2) * 32 + lane; e < p.N; e += 128 * S) {
        float g = 0.f, b = 0.f;
#pragma unroll
        for (int w = 0; w < 4 * S; ++w) {
          g += ptx::ld_shared_f32(s0 + (uint32_t)w * kWarpStaging + (uint32_t)(e * 4));
          b += ptx::ld_shared_f32(s0 + (uint32_t)w * kWarpStaging + 1024u + (uint32_t)(e * 4));
        }
        if (p.dgamma) atomicAdd(p.dgamma + e, g);
        if (p.dbeta) atomicAdd(p.dbeta + e, b);
      }
    }
    if (warp == 2 && lane == 0) TR_DUMP(2);
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
// A-resident variant for wide outputs with a short reduction (K <= 256: q|k|v, gated up-projection, d(gate)):
// a CTA owns whole 128-row blocks, loads the A block ONCE into shared memory and streams only the weight tiles
// through the ring while it walks the output tiles of that block -- the generic kernel re-fetches A for every
// output tile.  P = 2 is the CTA-pair mode described above (each CTA streams half of every weight tile).
// ---------------------------------------------------------------------------
// The MMA tile is up to 256 columns wide (one instruction stream per FLOP is what limits these short-K GEMMs: the
// single issuing thread needs ~400-600 cycles per 64-deep k-block for waits, descriptors and commits, so each k-block
// must carry >= 512 tensor-cycles of work).  Accumulators wider than 128 columns are drained by TWO epilogue groups,
// one per 128-column half, so all four groups stay busy: H = halves per tile, S = 4 / H accumulator stages.
template <int EPI, int P>
__global__ void __launch_bounds__(64 + 128 * 4, 1)
gemm_tc_ares_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1, GemmArgs p,
                    int block_n, int stages, int n_blks, int m_units) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  constexpr uint32_t kStagingBytes = 16 * kStageBufBytes;   // one box per epilogue warp
  const int H = block_n > 128 ? 2 : 1;
  const int S = 4 / H;
  const uint32_t acc_stride = 128u * (uint32_t)H;

  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  const uint32_t a_bytes = (uint32_t)num_kb * kATileBytes;
  const uint32_t b_bytes = (uint32_t)(block_n / P) * 128u;
  uint8_t* a_res = smem;
  uint8_t* ring = a_res + a_bytes;
  uint8_t* staging = ring + (size_t)stages * b_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
  uint64_t* empty = full + stages;
  uint64_t* tfull = empty + stages;
  uint64_t* tempty = tfull + 4;
  uint64_t* a_full = tempty + 4;
  uint64_t* a_empty = a_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = P == 2 ? (int)ptx::cluster_ctarank() : 0;
  const int unit0 = (int)blockIdx.x / P, unit_step = (int)gridDim.x / P;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA); ptx::prefetch_tmap(&tmB); ptx::prefetch_tmap(&tmO0); ptx::prefetch_tmap(&tmO1);
    for (int i = 0; i < stages; ++i) { ptx::mbar_init(full + i, 1); ptx::mbar_init(empty + i, 1); }
    for (int i = 0; i < S; ++i) { ptx::mbar_init(tfull + i, 1); ptx::mbar_init(tempty + i, 128 * H * P); }
    ptx::mbar_init(a_full, 1); ptx::mbar_init(a_empty, 1);
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

  if (warp == 0) {
    if (ptx::elect_one()) {
      // weight tiles stream continuously (they do not depend on the resident A block)
      int stage = 0; uint32_t phase = 0;
      TR_DECL;
      for (int mu = unit0; mu < m_units; mu += unit_step)
        for (int nb = 0; nb < n_blks; ++nb)
          for (int kb = 0; kb < num_kb; ++kb) {
            TR_WAIT(0, ptx::mbar_wait(empty + stage, phase ^ 1u));
            if constexpr (P == 2) {
              if (crank == 0) ptx::mbar_expect_tx(full + stage, 2u * b_bytes);
              ptx::tma_load_2d_pair(ring + (size_t)stage * b_bytes, &tmB, ptx::mapa_rank(ptx::smem_u32(full + stage), 0), kb * kBlockK,
                                    nb * block_n + crank * (block_n / P));
            } else {
              ptx::mbar_expect_tx(full + stage, b_bytes);
              ptx::tma_load_2d(ring + (size_t)stage * b_bytes, &tmB, full + stage, kb * kBlockK, nb * block_n);
            }
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
      TR_DUMP(0);
    }
  } else if (warp == 1) {
    // One elected thread owns the resident A block of its CTA -- reloaded once every MMA that read the previous block has
    // retired (a_empty, signalled in both CTAs of a pair) -- and, in the leader CTA, issues the MMAs.
    if (ptx::elect_one()) {
      const uint32_t idesc = make_idesc(block_n, false, false, kBlockM * P);
      const uint32_t a_full_addr = P == 2 ? ptx::mapa_rank(ptx::smem_u32(a_full), 0) : 0u;
      const uint64_t adesc0 = make_smem_desc(ptx::smem_u32(a_res), 16, 1024);
      const uint64_t bdesc0 = make_smem_desc(ptx::smem_u32(ring), 16, 1024);
      const uint32_t b_units = b_bytes >> 4;
      int stage = 0; uint32_t phase = 0, aph = 0;
      int it = 0;
      TR_DECL;
      for (int mu = unit0; mu < m_units; mu += unit_step) {
        const int m0 = (mu * P + crank) * kBlockM;   // may lie past the last row: TMA zero-fills the load and clips the stores
        if (mu != unit0) ptx::mbar_wait(a_empty, aph ^ 1u);
        if constexpr (P == 2) {
          if (crank == 0) ptx::mbar_expect_tx(a_full, 2u * a_bytes);
          for (int kb = 0; kb < num_kb; ++kb) ptx::tma_load_2d_pair(a_res + (size_t)kb * kATileBytes, &tmA, a_full_addr, kb * kBlockK, m0);
        } else {
          ptx::mbar_expect_tx(a_full, a_bytes);
          for (int kb = 0; kb < num_kb; ++kb) ptx::tma_load_2d(a_res + (size_t)kb * kATileBytes, &tmA, a_full, kb * kBlockK, m0);
        }
        if (crank == 0) {
          TR_WAIT(0, ptx::mbar_wait(a_full, aph));
          for (int nb = 0; nb < n_blks; ++nb, ++it) {
            const int as = it % S;
            TR_WAIT(1, ptx::mbar_wait(tempty + as, ((uint32_t)(it / S) & 1u) ^ 1u));
            ptx::tc_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)as * acc_stride;
            for (int kb = 0; kb < num_kb; ++kb) {
              TR_WAIT(2, ptx::mbar_wait(full + stage, phase));
              ptx::tc_fence_after();
              const uint64_t adesc = adesc0 + (uint64_t)(kb * (kATileBytes >> 4));
              const uint64_t bdesc = bdesc0 + (uint64_t)((uint32_t)stage * b_units);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                if constexpr (P == 2) ptx::umma_bf16_pair(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
                else ptx::umma_bf16(d_tmem, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0 ? 1u : 0u);
              }
              if constexpr (P == 2) ptx::umma_commit_pair(empty + stage); else ptx::umma_commit(empty + stage);
              if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
            if constexpr (P == 2) ptx::umma_commit_pair(tfull + as); else ptx::umma_commit(tfull + as);
          }
          if constexpr (P == 2) ptx::umma_commit_pair(a_empty); else ptx::umma_commit(a_empty);
        }
        aph ^= 1u;
      }
      TR_DUMP(1);
    }
  } else {
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int as = grp / H, half = grp % H;     // accumulator stage and 128-column half served by this group
    Stager st{ptx::smem_u32(staging) + (uint32_t)(warp - 2) * kStageBufBytes, lane, false};
    const uint32_t tempty_addr = P == 2 ? ptx::mapa_rank(ptx::smem_u32(tempty + as), 0) : 0u;
    const int my_units = unit0 < m_units ? (m_units - 1 - unit0) / unit_step + 1 : 0;
    const int my_tiles = my_units * n_blks;
    TR_DECL;
    for (int it = as; it < my_tiles; it += S) {
      const int mu = unit0 + (it / n_blks) * unit_step, nb = it % n_blks;
      const uint32_t aphase = (uint32_t)(it / S) & 1u;
      TmemAcc acc{tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * acc_stride + (uint32_t)half * 128u};
      const int n0 = nb * block_n + half * 128;
      int width = p.N - n0; if (width > 128) width = 128;
      if (width > block_n) width = block_n;
      auto wait_acc = [&]() { TR_WAIT(0, ptx::mbar_wait(tfull + as, aphase)); ptx::tc_fence_after(); };
      if (width > 0) {
        TR_WAIT(1, tc_epilogue<EPI, 2>(p, acc, st, &tmO0, &tmO1, (mu * P + crank) * kBlockM + q * 32, lane, n0, width, wait_acc));
      } else {
        wait_acc();   // nothing to drain in this half, but the stage is only free once the tile's MMAs have retired
      }
      ptx::tc_fence_before();
      if constexpr (P == 2) ptx::mbar_arrive_cluster(tempty_addr); else ptx::mbar_arrive(tempty + as);
    }
    st.acquire();
    if (warp == 2 && lane == 0) TR_DUMP(2);
  }

  ptx::pdl_trigger();
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (P == 2) ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (P == 2) ptx::tmem_dealloc_pair(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------
// d(gate) with recomputed pre-activations (kEpiDGate): per 64 hidden units one accumulator stage holds
//   [0,128)   a|b = A2 * B2^T   (the forward's gated up-projection, recomputed: 2 K N2 FLOP per row instead of
//                                writing 2 N2 bytes per row in forward and reading them back here)
//   [128,192) dg  = A  * B^T    (gradient w.r.t. the gate output)
// and the epilogue turns them into the interleaved d(a|b) tile.  Both A blocks stay resident (2 x K/64 tiles of
// 16 KB); the two weight tiles of a k-block travel in one ring stage.  P = 2: CTA pair, halves of both weight tiles.
// ---------------------------------------------------------------------------
constexpr int kGateChunk = 64;

template <int P>
__global__ void __launch_bounds__(64 + 128 * 4, 1)
gemm_tc_dgate_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
                     const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmB2,
                     const __grid_constant__ CUtensorMap tmO0, GemmArgs p, int stages, int n_chunks, int m_units) {
  constexpr int S = 2;
  constexpr uint32_t kAccStride = 256;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  constexpr uint32_t kHalfBoxBytes = 32 * 64;              // [32 rows x 64 bytes], 64B swizzle: 16 hidden units of d(a|b)
  constexpr uint32_t kStagingBytes = 16 * kHalfBoxBytes;   // one half box per epilogue warp

  const int num_kb = (p.K + kBlockK - 1) / kBlockK;
  const uint32_t a_bytes = (uint32_t)num_kb * kATileBytes;
  constexpr uint32_t b2_bytes = (uint32_t)(2 * kGateChunk / P) * 128u;   // this CTA's share of the w1|w3 tile
  constexpr uint32_t b_bytes = (uint32_t)(kGateChunk / P) * 128u;        // ... and of the w2^T tile
  constexpr uint32_t stage_bytes = b2_bytes + b_bytes;
  uint8_t* a_res = smem;                 // dy block
  uint8_t* a2_res = a_res + a_bytes;     // gate-input block
  uint8_t* ring = a2_res + a_bytes;
  uint8_t* staging = ring + (size_t)stages * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(staging + kStagingBytes);
  uint64_t* empty = full + stages;
  uint64_t* tfull = empty + stages;
  uint64_t* tempty = tfull + S;
  uint64_t* a_full = tempty + S;
  uint64_t* a_empty = a_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = P == 2 ? (int)ptx::cluster_ctarank() : 0;
  const int unit0 = (int)blockIdx.x / P, unit_step = (int)gridDim.x / P;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmA); ptx::prefetch_tmap(&tmA2); ptx::prefetch_tmap(&tmB); ptx::prefetch_tmap(&tmB2); ptx::prefetch_tmap(&tmO0);
    for (int i = 0; i < stages; ++i) { ptx::mbar_init(full + i, 1); ptx::mbar_init(empty + i, 1); }
    for (int i = 0; i < S; ++i) { ptx::mbar_init(tfull + i, 1); ptx::mbar_init(tempty + i, 256 * P); }
    ptx::mbar_init(a_full, 1); ptx::mbar_init(a_empty, 1);
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

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0; uint32_t phase = 0;
      TR_DECL;
      for (int mu = unit0; mu < m_units; mu += unit_step)
        for (int ch = 0; ch < n_chunks; ++ch)
          for (int kb = 0; kb < num_kb; ++kb) {
            TR_WAIT(0, ptx::mbar_wait(empty + stage, phase ^ 1u));
            uint8_t* sb = ring + (size_t)stage * stage_bytes;
            const int r2 = ch * 2 * kGateChunk + crank * (2 * kGateChunk / P), r1 = ch * kGateChunk + crank * (kGateChunk / P);
            if constexpr (P == 2) {
              if (crank == 0) ptx::mbar_expect_tx(full + stage, 2u * stage_bytes);
              const uint32_t bar = ptx::mapa_rank(ptx::smem_u32(full + stage), 0);
              ptx::tma_load_2d_pair(sb, &tmB2, bar, kb * kBlockK, r2);
              ptx::tma_load_2d_pair(sb + b2_bytes, &tmB, bar, kb * kBlockK, r1);
            } else {
              ptx::mbar_expect_tx(full + stage, stage_bytes);
              ptx::tma_load_2d(sb, &tmB2, full + stage, kb * kBlockK, r2);
              ptx::tma_load_2d(sb + b2_bytes, &tmB, full + stage, kb * kBlockK, r1);
            }
            if (++stage == stages) { stage = 0; phase ^= 1u; }
          }
      TR_DUMP(0);
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      const uint32_t idesc_ab = make_idesc(2 * kGateChunk, false, false, kBlockM * P);
      const uint32_t idesc_dg = make_idesc(kGateChunk, false, false, kBlockM * P);
      const uint32_t a_full_addr = P == 2 ? ptx::mapa_rank(ptx::smem_u32(a_full), 0) : 0u;
      const uint64_t adesc0 = make_smem_desc(ptx::smem_u32(a_res), 16, 1024);
      const uint64_t a2desc0 = make_smem_desc(ptx::smem_u32(a2_res), 16, 1024);
      const uint64_t bdesc0 = make_smem_desc(ptx::smem_u32(ring), 16, 1024);
      int stage = 0; uint32_t phase = 0, aph = 0;
      int it = 0;
      TR_DECL;
      for (int mu = unit0; mu < m_units; mu += unit_step) {
        const int m0 = (mu * P + crank) * kBlockM;
        if (mu != unit0) ptx::mbar_wait(a_empty, aph ^ 1u);
        if constexpr (P == 2) {
          if (crank == 0) ptx::mbar_expect_tx(a_full, 4u * a_bytes);
          for (int kb = 0; kb < num_kb; ++kb) {
            ptx::tma_load_2d_pair(a2_res + (size_t)kb * kATileBytes, &tmA2, a_full_addr, kb * kBlockK, m0);
            ptx::tma_load_2d_pair(a_res + (size_t)kb * kATileBytes, &tmA, a_full_addr, kb * kBlockK, m0);
          }
        } else {
          ptx::mbar_expect_tx(a_full, 2u * a_bytes);
          for (int kb = 0; kb < num_kb; ++kb) {
            ptx::tma_load_2d(a2_res + (size_t)kb * kATileBytes, &tmA2, a_full, kb * kBlockK, m0);
            ptx::tma_load_2d(a_res + (size_t)kb * kATileBytes, &tmA, a_full, kb * kBlockK, m0);
          }
        }
        if (crank == 0) {
          TR_WAIT(0, ptx::mbar_wait(a_full, aph));
          for (int ch = 0; ch < n_chunks; ++ch, ++it) {
            const int as = it % S;
            TR_WAIT(1, ptx::mbar_wait(tempty + as, ((uint32_t)(it / S) & 1u) ^ 1u));
            ptx::tc_fence_after();
            const uint32_t d_ab = tmem_base + (uint32_t)as * kAccStride, d_dg = d_ab + 2 * kGateChunk;
            for (int kb = 0; kb < num_kb; ++kb) {
              TR_WAIT(2, ptx::mbar_wait(full + stage, phase));
              ptx::tc_fence_after();
              const uint64_t koff = (uint64_t)(kb * (kATileBytes >> 4));
              const uint64_t b2desc = bdesc0 + (uint64_t)((uint32_t)stage * (stage_bytes >> 4));
              const uint64_t bdesc = b2desc + (uint64_t)(b2_bytes >> 4);
#pragma unroll
              for (int k = 0; k < kBlockK / 16; ++k) {
                const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
                if constexpr (P == 2) {
                  ptx::umma_bf16_pair(d_ab, a2desc0 + koff + (uint64_t)(k * 2), b2desc + (uint64_t)(k * 2), idesc_ab, accum);
                  ptx::umma_bf16_pair(d_dg, adesc0 + koff + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_dg, accum);
                } else {
                  ptx::umma_bf16(d_ab, a2desc0 + koff + (uint64_t)(k * 2), b2desc + (uint64_t)(k * 2), idesc_ab, accum);
                  ptx::umma_bf16(d_dg, adesc0 + koff + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_dg, accum);
                }
              }
              if constexpr (P == 2) ptx::umma_commit_pair(empty + stage); else ptx::umma_commit(empty + stage);
              if (++stage == stages) { stage = 0; phase ^= 1u; }
            }
            if constexpr (P == 2) ptx::umma_commit_pair(tfull + as); else ptx::umma_commit(tfull + as);
          }
          if constexpr (P == 2) ptx::umma_commit_pair(a_empty); else ptx::umma_commit(a_empty);
        }
        aph ^= 1u;
      }
      TR_DUMP(1);
    }
  } else {
    // Four epilogue groups: stage = group / 2, and each group drains one 32-unit half of the 64 hidden units, so
    // every accumulator stage is emptied by 8 warps (the epilogue, not the MMA, is the longer side of this kernel).
    const int q = warp & 3;
    const int grp = (warp - 2) >> 2;
    const int as = grp >> 1, half = grp & 1;
    const uint32_t box = ptx::smem_u32(staging) + (uint32_t)(warp - 2) * kHalfBoxBytes;
    const bool leader = ptx::elect_one();
    bool pending = false;
    const uint32_t tempty_addr = P == 2 ? ptx::mapa_rank(ptx::smem_u32(tempty + as), 0) : 0u;
    const int my_units = unit0 < m_units ? (m_units - 1 - unit0) / unit_step + 1 : 0;
    const int my_tiles = my_units * n_chunks;
    const uint32_t acc0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)as * kAccStride;
    TR_DECL;
    for (int it = as; it < my_tiles; it += S) {
      const int mu = unit0 + (it / n_chunks) * unit_step, ch = it % n_chunks;
      const uint32_t aphase = (uint32_t)(it / S) & 1u;
      const int h0 = ch * kGateChunk + half * (kGateChunk / 2);
      int width = p.N - h0; if (width > kGateChunk / 2) width = kGateChunk / 2;   // hidden units of this half (multiple of 16, may be <= 0)
      const int m0 = (mu * P + crank) * kBlockM + q * 32;
      TR_WAIT(0, ptx::mbar_wait(tfull + as, aphase));
      ptx::tc_fence_after();
      for (int c = 0; c < width; c += 16) {
        const int cc = half * (kGateChunk / 2) + c;      // hidden offset inside the chunk
        float ab[32], dg[16], o[32];
        ptx::tmem_ld32(acc0 + 2 * cc, ab);
        ptx::tmem_ld16(acc0 + 2 * kGateChunk + cc, dg);
        ptx::tmem_ld_wait();
        if (p.bias) add_vec<32>(p.bias + 2 * (h0 + c), ab);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float a = ab[i], b = ab[16 + i];
          const float sg = ptx::sigmoid_fast(a);
          o[i] = dg[i] * b * (sg * (1.0f + a * (1.0f - sg)));
          o[16 + i] = dg[i] * (a * sg);
        }
        if (pending) { if (leader) ptx::bulk_wait_read0(); __syncwarp(); pending = false; }
        // 64-byte row of the half box; 64B swizzle: 16-byte piece index XOR ((row >> 1) & 3)
#pragma unroll
        for (int i = 0; i < 4; ++i)
          ptx::st_shared_v4(box + (uint32_t)lane * 64u + (uint32_t)((i ^ ((lane >> 1) & 3)) << 4), pack8_bf16(o + 8 * i));
        ptx::fence_proxy_async();
        __syncwarp();
        if (leader) { ptx::tma_store_2d(&tmO0, box, 2 * (h0 + c), m0); ptx::bulk_commit(); }
        pending = true;
      }
      ptx::tc_fence_before();
      if constexpr (P == 2) ptx::mbar_arrive_cluster(tempty_addr); else ptx::mbar_arrive(tempty + as);
    }
    if (pending) { if (leader) ptx::bulk_wait_read0(); __syncwarp(); }
    if (warp == 2 && lane == 0) TR_DUMP(2);
  }

  ptx::pdl_trigger();
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (P == 2) ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (P == 2) ptx::tmem_dealloc_pair(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------
// wgrad kernel: one (output tile, reduction split) per CTA
// ---------------------------------------------------------------------------
constexpr int kBoxBytes = 64 * 128;  // one {64 x 64} bf16 TMA box
constexpr int kOnesBytes = 16 * 1024;  // all-ones B operand for the fused column sums

// Bias gradients ride along: dbias[n] = sum_m Y[m,n] * 1 is one more MMA per K-step against an all-ones
// B tile (N = 16) that lives in shared memory for the whole kernel -- a tile of ones is the same in every
// swizzle/major layout -- accumulated in 16 extra TMEM columns.  Only the c_blk == 0 tiles do it.
// P = 2: a CTA pair owns 256 output rows x bn columns of one reduction split; each CTA loads the Y box of its own 128
// output rows and HALF of the X columns (the operand that every row tile re-reads), the leader issues 256-row MMAs.
template <int P>
__global__ void __launch_bounds__(kGemmThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmX, WgradArgs p,
                int bn, int stages, int tiles_c, int num_tiles, int kb_total, int kb_per_split,
                uint32_t lbo, uint32_t sbo, uint32_t kstep_bytes) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);

  const int bn_cta = bn / P;                                    // X columns staged by this CTA
  const uint32_t a_bytes = 2 * kBoxBytes;
  const uint32_t b_bytes = (uint32_t)(bn_cta / 64) * kBoxBytes;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* ones = smem + (size_t)stages * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ones + kOnesBytes);
  uint64_t* empty = full + stages;
  uint64_t* tfull = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = P == 2 ? (int)ptx::cluster_ctarank() : 0;
  const int unit = (int)blockIdx.x / P;
  const int tile = unit % num_tiles;
  const int split = unit / num_tiles;
  const int r_blk = tile / tiles_c, c_blk = tile - r_blk * tiles_c;
  const int row0 = (r_blk * P + crank) * kBlockM;               // first output row of this CTA
  const int kb0 = split * kb_per_split;
  int kb1 = kb0 + kb_per_split; if (kb1 > kb_total) kb1 = kb_total;
  const int nkb = kb1 - kb0;
  const bool do_bias = p.bias0 != nullptr && c_blk == 0;

  if (do_bias) {
    for (int i = threadIdx.x; i < kOnesBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;  // bf16 1.0 x2
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core (async proxy)
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmY);
    ptx::prefetch_tmap(&tmX);
    for (int i = 0; i < stages; ++i) { ptx::mbar_init(full + i, 1); ptx::mbar_init(empty + i, 1); }
    ptx::mbar_init(tfull, 1);
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

  if (nkb > 0) {
    if (warp == 0) {
      if (ptx::elect_one()) {
        int stage = 0; uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(empty + stage, phase ^ 1u);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          if constexpr (P == 2) {
            if (crank == 0) ptx::mbar_expect_tx(full + stage, 2u * stage_bytes);
            const uint32_t bar = ptx::mapa_rank(ptx::smem_u32(full + stage), 0);
            for (int j = 0; j < 2; ++j) ptx::tma_load_2d_pair(sa + j * kBoxBytes, &tmY, bar, row0 + j * 64, kb * 64);
            for (int j = 0; j < bn_cta / 64; ++j)
              ptx::tma_load_2d_pair(sa + a_bytes + j * kBoxBytes, &tmX, bar, c_blk * bn + crank * bn_cta + j * 64, kb * 64);
          } else {
            ptx::mbar_expect_tx(full + stage, stage_bytes);
            for (int j = 0; j < 2; ++j) ptx::tma_load_2d(sa + j * kBoxBytes, &tmY, full + stage, row0 + j * 64, kb * 64);
            for (int j = 0; j < bn / 64; ++j)
              ptx::tma_load_2d(sa + a_bytes + j * kBoxBytes, &tmX, full + stage, c_blk * bn + j * 64, kb * 64);
          }
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      if (crank == 0 && ptx::elect_one()) {
        const uint32_t idesc = make_idesc(bn, true, true, kBlockM * P);
        const uint32_t idesc_ones = make_idesc(16, true, true, kBlockM * P);
        const uint64_t ones_desc = make_smem_desc(ptx::smem_u32(ones), lbo, sbo);
        int stage = 0; uint32_t phase = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(full + stage, phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * stage_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adesc = make_smem_desc(sa + k * kstep_bytes, lbo, sbo);
            const uint64_t bdesc = make_smem_desc(sa + a_bytes + k * kstep_bytes, lbo, sbo);
            if constexpr (P == 2) {
              ptx::umma_bf16_pair(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
              if (do_bias) ptx::umma_bf16_pair(tmem_base + 256, adesc, ones_desc, idesc_ones, (kb | k) != 0 ? 1u : 0u);
            } else {
              ptx::umma_bf16(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
              if (do_bias) ptx::umma_bf16(tmem_base + 256, adesc, ones_desc, idesc_ones, (kb | k) != 0 ? 1u : 0u);
            }
          }
          if constexpr (P == 2) ptx::umma_commit_pair(empty + stage); else ptx::umma_commit(empty + stage);
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        if constexpr (P == 2) ptx::umma_commit_pair(tfull); else ptx::umma_commit(tfull);
      }
    } else {
      const int q = warp & 3;
      ptx::mbar_wait(tfull, 0);
      ptx::tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16);
      const int r = row0 + q * 32 + lane;  // packed output row
      float* drow = nullptr;
      float* brow = nullptr;
      if (r < p.Nout) {
        if (p.row_map == 0) {
          if (r < p.rows_valid) { drow = p.dst0 + (size_t)r * p.ld; brow = p.bias0 ? p.bias0 + r : nullptr; }
        } else {
          const int which = (r % (2 * kGate)) / kGate;
          const int h = (r / (2 * kGate)) * kGate + (r % kGate);
          if (h < p.rows_valid) {
            drow = (which ? p.dst1 : p.dst0) + (size_t)h * p.ld;
            brow = p.bias0 ? (which ? p.bias1 : p.bias0) + h : nullptr;
          }
        }
      }
      for (int c = 0; c < bn; c += 16) {
        float v[16];
        ptx::tmem_ld16(tbase + c, v);
        ptx::tmem_ld_wait();
        const int col = c_blk * bn + c;
        if (drow != nullptr) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            if (col + i < p.cols_valid) red_add_f32x4(drow + col + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
      if (do_bias) {
        float v[16];
        ptx::tmem_ld16(tbase + 256, v);
        ptx::tmem_ld_wait();
        if (brow != nullptr) atomicAdd(brow, v[0]);
      }
    }
  }

  ptx::pdl_trigger();
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (P == 2) ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (P == 2) ptx::tmem_dealloc_pair(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------
// grouped wgrad: up to kMaxWgradJobs independent W += Y^T X problems in ONE launch (the four weight gradients of a
// transformer block: each alone is one partial wave with its own launch, drain and tail -- dWproj reaches 360 TFLOP/s
// on its own).  Same CTA program as wgrad_tc_kernel; a CTA (pair) looks its (job, tile, split) up in the job table.
// ---------------------------------------------------------------------------
constexpr int kMaxWgradJobs = 4;
struct WgradJob {
  WgradArgs p;
  int bn, tiles_c, num_tiles, kb_total, kb_per_split;
  int unit0;   // first unit (CTA or CTA pair) of this job in the launch
};
struct alignas(64) WgradGroupParams {
  CUtensorMap tmY[kMaxWgradJobs], tmX[kMaxWgradJobs];
  WgradJob job[kMaxWgradJobs];
  int njobs;
};

template <int P>
__global__ void __launch_bounds__(kGemmThreads, 1)
wgrad_group_kernel(const __grid_constant__ WgradGroupParams gp, int stages, uint32_t stage_bytes, uint32_t lbo, uint32_t sbo,
                   uint32_t kstep_bytes) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = P == 2 ? (int)ptx::cluster_ctarank() : 0;
  const int unit = (int)blockIdx.x / P;
  int ji = 0;
#pragma unroll
  for (int j = 1; j < kMaxWgradJobs; ++j)
    if (j < gp.njobs && unit >= gp.job[j].unit0) ji = j;
  const WgradJob& J = gp.job[ji];
  const WgradArgs& p = J.p;
  const CUtensorMap* tmY = &gp.tmY[ji];
  const CUtensorMap* tmX = &gp.tmX[ji];
  const int bn = J.bn;
  const int bn_cta = bn / P;                                    // X columns staged by this CTA
  const uint32_t a_bytes = 2 * kBoxBytes;
  const uint32_t b_bytes = (uint32_t)(bn_cta / 64) * kBoxBytes;
  const uint32_t job_stage_bytes = a_bytes + b_bytes;           // <= stage_bytes (the widest job of the launch)
  uint8_t* ones = smem + (size_t)stages * stage_bytes;
  uint64_t* full = reinterpret_cast<uint64_t*>(ones + kOnesBytes);
  uint64_t* empty = full + stages;
  uint64_t* tfull = empty + stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull + 1);

  const int local = unit - J.unit0;
  const int tile = local % J.num_tiles;
  const int split = local / J.num_tiles;
  const int r_blk = tile / J.tiles_c, c_blk = tile - r_blk * J.tiles_c;
  const int row0 = (r_blk * P + crank) * kBlockM;               // first output row of this CTA
  const int kb0 = split * J.kb_per_split;
  int kb1 = kb0 + J.kb_per_split; if (kb1 > J.kb_total) kb1 = J.kb_total;
  const int nkb = kb1 - kb0;
  const bool do_bias = p.bias0 != nullptr && c_blk == 0;

  if (do_bias) {
    for (int i = threadIdx.x; i < kOnesBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;  // bf16 1.0 x2
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(tmY);
    ptx::prefetch_tmap(tmX);
    for (int i = 0; i < stages; ++i) { ptx::mbar_init(full + i, 1); ptx::mbar_init(empty + i, 1); }
    ptx::mbar_init(tfull, 1);
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
  ptx::pdl_wait();

  if (nkb > 0) {
    if (warp == 0) {
      if (ptx::elect_one()) {
        int stage = 0; uint32_t phase = 0;
        for (int kb = kb0; kb < kb1; ++kb) {
          ptx::mbar_wait(empty + stage, phase ^ 1u);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          if constexpr (P == 2) {
            if (crank == 0) ptx::mbar_expect_tx(full + stage, 2u * job_stage_bytes);
            const uint32_t bar = ptx::mapa_rank(ptx::smem_u32(full + stage), 0);
            for (int j = 0; j < 2; ++j) ptx::tma_load_2d_pair(sa + j * kBoxBytes, tmY, bar, row0 + j * 64, kb * 64);
            for (int j = 0; j < bn_cta / 64; ++j)
              ptx::tma_load_2d_pair(sa + a_bytes + j * kBoxBytes, tmX, bar, c_blk * bn + crank * bn_cta + j * 64, kb * 64);
          } else {
            ptx::mbar_expect_tx(full + stage, job_stage_bytes);
            for (int j = 0; j < 2; ++j) ptx::tma_load_2d(sa + j * kBoxBytes, tmY, full + stage, row0 + j * 64, kb * 64);
            for (int j = 0; j < bn / 64; ++j)
              ptx::tma_load_2d(sa + a_bytes + j * kBoxBytes, tmX, full + stage, c_blk * bn + j * 64, kb * 64);
          }
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      if (crank == 0 && ptx::elect_one()) {
        const uint32_t idesc = make_idesc(bn, true, true, kBlockM * P);
        const uint32_t idesc_ones = make_idesc(16, true, true, kBlockM * P);
        const uint64_t ones_desc = make_smem_desc(ptx::smem_u32(ones), lbo, sbo);
        int stage = 0; uint32_t phase = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          ptx::mbar_wait(full + stage, phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + (size_t)stage * stage_bytes);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t adesc = make_smem_desc(sa + k * kstep_bytes, lbo, sbo);
            const uint64_t bdesc = make_smem_desc(sa + a_bytes + k * kstep_bytes, lbo, sbo);
            if constexpr (P == 2) {
              ptx::umma_bf16_pair(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
              if (do_bias) ptx::umma_bf16_pair(tmem_base + 256, adesc, ones_desc, idesc_ones, (kb | k) != 0 ? 1u : 0u);
            } else {
              ptx::umma_bf16(tmem_base, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
              if (do_bias) ptx::umma_bf16(tmem_base + 256, adesc, ones_desc, idesc_ones, (kb | k) != 0 ? 1u : 0u);
            }
          }
          if constexpr (P == 2) ptx::umma_commit_pair(empty + stage); else ptx::umma_commit(empty + stage);
          if (++stage == stages) { stage = 0; phase ^= 1u; }
        }
        if constexpr (P == 2) ptx::umma_commit_pair(tfull); else ptx::umma_commit(tfull);
      }
    } else {
      const int q = warp & 3;
      ptx::mbar_wait(tfull, 0);
      ptx::tc_fence_after();
      const uint32_t tbase = tmem_base + ((uint32_t)(q * 32) << 16);
      const int r = row0 + q * 32 + lane;  // packed output row
      float* drow = nullptr;
      float* brow = nullptr;
      if (r < p.Nout) {
        if (p.row_map == 0) {
          if (r < p.rows_valid) { drow = p.dst0 + (size_t)r * p.ld; brow = p.bias0 ? p.bias0 + r : nullptr; }
        } else {
          const int which = (r % (2 * kGate)) / kGate;
          const int h = (r / (2 * kGate)) * kGate + (r % kGate);
          if (h < p.rows_valid) {
            drow = (which ? p.dst1 : p.dst0) + (size_t)h * p.ld;
            brow = p.bias0 ? (which ? p.bias1 : p.bias0) + h : nullptr;
          }
        }
      }
      for (int c = 0; c < bn; c += 16) {
        float v[16];
        ptx::tmem_ld16(tbase + c, v);
        ptx::tmem_ld_wait();
        const int col = c_blk * bn + c;
        if (drow != nullptr) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            if (col + i < p.cols_valid) red_add_f32x4(drow + col + i, v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
      if (do_bias) {
        float v[16];
        ptx::tmem_ld16(tbase + 256, v);
        ptx::tmem_ld_wait();
        if (brow != nullptr) atomicAdd(brow, v[0]);
      }
    }
  }

  ptx::pdl_trigger();
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (P == 2) ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (P == 2) ptx::tmem_dealloc_pair(tmem_base, 512); else ptx::tmem_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------
// column sums (bias gradients): dst[map(n)] += sum_m Y[m][n]
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
colsum_kernel(const __nv_bfloat16* __restrict__ Y, int ldy, int Mred, int Nout, int rows_per_cta,
              float* dst0, float* dst1, int row_map, int rows_valid) {
  __shared__ float part[8][64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 64 + lane * 2;
  const int m0 = blockIdx.y * rows_per_cta;
  int m1 = m0 + rows_per_cta; if (m1 > Mred) m1 = Mred;
  float s0 = 0.f, s1 = 0.f;
  if (n < Nout) {
    for (int m = m0 + warp; m < m1; m += 8) {
      uint32_t u = *reinterpret_cast<const uint32_t*>(Y + (size_t)m * ldy + n);
      float2 f = unpack_bf16x2(u);
      s0 += f.x; s1 += f.y;
    }
  }
  part[warp][lane * 2] = s0; part[warp][lane * 2 + 1] = s1;
  __syncthreads();
  if (threadIdx.x < 64) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += part[w][threadIdx.x];
    const int r = blockIdx.x * 64 + threadIdx.x;
    if (r < Nout) {
      if (row_map == 0) { if (r < rows_valid) atomicAdd(dst0 + r, s); }
      else {
        const int which = (r % (2 * kGate)) / kGate;
        const int h = (r / (2 * kGate)) * kGate + (r % kGate);
        if (h < rows_valid) atomicAdd((which ? dst1 : dst0) + h, s);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// host side: tensor-map cache + launchers
// ---------------------------------------------------------------------------
namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

struct MapKey {
  const void* ptr; uint64_t d0, d1, pitch; uint32_t b0, b1, esz;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && pitch == o.pitch && b0 == o.b0 && b1 == o.b1 && esz == o.esz;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = std::hash<const void*>()(k.ptr);
    auto mix = [&h](uint64_t v) { h ^= std::hash<uint64_t>()(v) + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.d0); mix(k.d1); mix(k.pitch); mix(k.b0); mix(k.b1); mix(k.esz);
    return h;
  }
};

std::mutex g_map_mu;
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;

// 2-D tensor of bf16 (esz 2) or fp32 (esz 4), dim0 contiguous (d0 elements), d1 rows of `pitch` elements,
// box {b0, b1} with b0 * esz == 128 bytes (128B swizzle) or 64 bytes (64B swizzle).
}  // namespace

int get_tmap(const void* ptr, uint64_t d0, uint64_t d1, uint64_t pitch, uint32_t b0, uint32_t b1, CUtensorMap* out, uint32_t esz) {
  MapKey key{ptr, d0, d1, pitch, b0, b1, esz};
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) { *out = it->second; return kOk; }
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point not available"); return kCudaError; }
  HS_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand %p is not 16-byte aligned", ptr);
  HS_REQUIRE((pitch * esz) % 16 == 0, "TMA operand pitch %llu elements is not a multiple of 16 bytes", (unsigned long long)pitch);
  HS_REQUIRE((b0 * esz == 128 || b0 * esz == 64) && b1 <= 256, "bad TMA box {%u,%u}", b0, b1);
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {pitch * esz};
  cuuint32_t box[2] = {b0, b1};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMap m;
  CUresult r = fn(&m, esz == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, b0 * esz == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) ptr=%p dims={%llu,%llu} pitch=%llu box={%u,%u}", (int)r, ptr,
              (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)pitch, b0, b1);
    return kCudaError;
  }
  {
    std::lock_guard<std::mutex> lk(g_map_mu);
    if (g_maps.size() > 65536) g_maps.clear();
    g_maps[key] = m;
  }
  *out = m;
  return kOk;
}

namespace {

template <int EPI, int S, int P>
int launch_gemm_s(const GemmArgs& a, const CUtensorMap* tm, int block_n, int n_blks, int m_blks, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    HS_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<EPI, S, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    configured = true;
  }
  // shared memory: [operand pipeline stages][epilogue staging boxes][barriers]; as many stages as fit
  const int staging = 4 * S * StagingBufs<EPI, S>::value * kStageBufBytes;
  const int stage_bytes = kATileBytes + block_n / P * 128;
  const int num_kb = ceil_div(a.K, kBlockK);
  // epilogues that read per-row global inputs (saved pre-activations, residual) rely on L1 to merge each thread's
  // 16-byte loads of one line: leave ~36 KB of the 228 KB shared/L1 array to the cache for them (the two-stage
  // residual epilogue stages its rows through TMA instead)
  const int smem_cap = (EPI == kEpiDSwiGLU || (EPI == kEpiResidLN && S != 2)) ? kSmemMax - 36 * 1024 : kSmemMax;
  int stages = (smem_cap - staging - 2048 - (EPI == kEpiLnBwd ? 1024 : 0)) / stage_bytes;
  if (stages > 8) stages = 8;
  if (stages > 2 * num_kb) stages = 2 * num_kb;
  if (stages < 2) stages = 2;
  const size_t smem = (size_t)stages * stage_bytes + staging + 1024 + 512 + (EPI == kEpiLnBwd ? 1024 : 0);
  HS_REQUIRE(smem <= (size_t)kSmemMax, "gemm: tile N=%d needs %zu bytes of shared memory", block_n, smem);
  const int num_tiles = ceil_div(m_blks, P) * n_blks;
  HS_TRY(launch_clustered(gemm_tc_kernel<EPI, S, P>, pair_grid(num_tiles, P), 64 + 128 * S, smem, P, stream, tm[0], tm[1], tm[2],
                          tm[3], tm[4], tm[5], a, block_n, stages, n_blks, num_tiles));
  HS_CHECK_LAUNCH("gemm_tc_kernel");
  return kOk;
}

template <int EPI, int P>
int launch_gemm_ares(const GemmArgs& a, const CUtensorMap* tm, int block_n, int n_blks, int m_blks, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    HS_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_ares_kernel<EPI, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    configured = true;
  }
  const int num_kb = ceil_div(a.K, kBlockK);
  const int fixed = num_kb * kATileBytes + 16 * kStageBufBytes + 2048;
  // d(gate) reads its saved pre-activations row-wise through L1: keep ~36 KB of the array as cache
  const int cap = (EPI == kEpiDSwiGLU) ? kSmemMax - 36 * 1024 : kSmemMax;
  const int b_bytes = block_n / P * 128;
  int stages = (cap - fixed) / b_bytes;
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  const size_t smem = (size_t)fixed + (size_t)stages * b_bytes;
  HS_REQUIRE(smem <= (size_t)kSmemMax, "gemm(A-resident): %zu bytes of shared memory", smem);
  const int m_units = ceil_div(m_blks, P);
  HS_TRY(launch_clustered(gemm_tc_ares_kernel<EPI, P>, pair_grid(m_units, P), 64 + 128 * 4, smem, P, stream, tm[0], tm[1], tm[2],
                          tm[3], a, block_n, stages, n_blks, m_units));
  HS_CHECK_LAUNCH("gemm_tc_ares_kernel");
  return kOk;
}

// wide output + short reduction: keep the A block resident (HSIMAE_GEMM_ARES=0 disables, for A/B measurements)
// Plain bias epilogues with 256-column tiles run faster on the streaming kernel with CTA pairs (q|k|v: 36 vs 44 us --
// re-fetching the 16 KB A tiles costs less than the exposed reload of a resident A block per row block).
bool use_ares(const GemmArgs& a, int epi, int block_n, int n_blks) {
  static const bool enabled = !(getenv("HSIMAE_GEMM_ARES") && atoi(getenv("HSIMAE_GEMM_ARES")) == 0);
  if (epi == kEpiBiasBf16 && block_n > 128) return false;
  return enabled && n_blks >= 2 && a.K <= 256 && (epi == kEpiBiasBf16 || epi == kEpiSwiGLU || epi == kEpiDSwiGLU);
}

// CTA pairs (cta_group::2): each half of the weight tile must be whole 8-row swizzle atoms and a legal share of the
// MMA's N (HSIMAE_GEMM_PAIR=0 disables, for A/B measurements)
// Measured (B200, M = 73 728): pairs pay where the operand stream is the limit (long reductions into 256 columns:
// 53 -> 49 us at K = 1376), not where the epilogue is (HSIMAE_GEMM_PAIR = 0 | 1 | 2: never | default policy | wherever legal).
bool use_pair(int block_n, int m_blks, int K, int N, bool ares, bool light_epilogue) {
  static const int mode = getenv("HSIMAE_GEMM_PAIR") ? atoi(getenv("HSIMAE_GEMM_PAIR")) : 1;
  if (mode == 0 || block_n % 32 != 0 || m_blks < 2) return false;
  if (mode >= 2) return true;
  // less than one row block per SM: nothing streams long enough for the halved weight tiles to matter, and the
  // cluster launch / cross-CTA handshakes only add latency (small fine-tuning and inference batches)
  if (m_blks < kNumSMs) return false;
  // A-resident kernels stream their weight tiles at the per-SM L2 read rate (~42 B/clk); halving them pays once the
  // epilogue is light (gated projection that does not keep a|b: 62 -> 58 us), not when it is the limit anyway
  if (ares) return light_epilogue;
  return K >= 512 || N >= 512;
}

template <int EPI>
int launch_gemm(const GemmArgs& a, const CUtensorMap* tm, int block_n, int n_blks, int m_blks, bool ares, bool pair, cudaStream_t stream) {
  if (ares)
    return pair ? launch_gemm_ares<EPI, 2>(a, tm, block_n, n_blks, m_blks, stream) : launch_gemm_ares<EPI, 1>(a, tm, block_n, n_blks, m_blks, stream);
  // as many accumulator stages (= epilogue warp groups) as fit in the 512 TMEM columns
  if (block_n <= 128)
    return pair ? launch_gemm_s<EPI, 4, 2>(a, tm, block_n, n_blks, m_blks, stream) : launch_gemm_s<EPI, 4, 1>(a, tm, block_n, n_blks, m_blks, stream);
  return pair ? launch_gemm_s<EPI, 2, 2>(a, tm, block_n, n_blks, m_blks, stream) : launch_gemm_s<EPI, 2, 1>(a, tm, block_n, n_blks, m_blks, stream);
}

template <int P>
int launch_gemm_dgate(const GemmArgs& a, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    HS_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_dgate_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    configured = true;
  }
  const int num_kb = ceil_div(a.K, kBlockK);
  const int m_blks = ceil_div(a.M, kBlockM);
  const int fixed = 2 * num_kb * kATileBytes + 16 * 32 * 64 + 2048;
  const int stage_bytes = 3 * kGateChunk / P * 128;
  int stages = (kSmemMax - fixed) / stage_bytes;
  if (stages > 8) stages = 8;
  HS_REQUIRE(stages >= 2, "gemm(d-gate): K=%d leaves no room for the weight ring", a.K);
  const size_t smem = (size_t)fixed + (size_t)stages * stage_bytes;
  CUtensorMap tmA, tmA2, tmB, tmB2, tmO;
  HS_TRY(get_tmap(a.A, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.lda, 64, kBlockM, &tmA));
  HS_TRY(get_tmap(a.A2, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.lda2, 64, kBlockM, &tmA2));
  HS_TRY(get_tmap(a.B, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.ldb, 64, (uint32_t)(kGateChunk / P), &tmB));
  HS_TRY(get_tmap(a.B2, (uint64_t)a.K, (uint64_t)(2 * a.N), (uint64_t)a.ldb2, 64, (uint32_t)(2 * kGateChunk / P), &tmB2));
  HS_TRY(get_tmap(a.out0, (uint64_t)(2 * a.N), (uint64_t)a.M, (uint64_t)a.ld0, 32, 32, &tmO));   // half boxes: 32 packed columns
  const int m_units = ceil_div(m_blks, P);
  HS_TRY(launch_clustered(gemm_tc_dgate_kernel<P>, pair_grid(m_units, P), 64 + 128 * 4, smem, P, stream, tmA, tmA2, tmB, tmB2, tmO, a,
                          stages, ceil_div(a.N, kGateChunk), m_units));
  HS_CHECK_LAUNCH("gemm_tc_dgate_kernel");
  return kOk;
}

int env_int(const char* name, int dflt) {
  const char* s = getenv(name);
  return s ? atoi(s) : dflt;
}

}  // namespace

int gemm_check_args(const GemmArgs& a, int epi) {
  HS_REQUIRE(a.M > 0 && a.N > 0 && a.K > 0, "gemm: empty problem M=%d N=%d K=%d", a.M, a.N, a.K);
  HS_REQUIRE(a.N % 16 == 0, "gemm: N=%d must be a multiple of 16", a.N);
  HS_REQUIRE(a.lda % 8 == 0 && a.ldb % 8 == 0, "gemm: lda=%d ldb=%d must be multiples of 8", a.lda, a.ldb);
  HS_REQUIRE(epi >= 0 && epi < kNumEpilogues, "gemm: bad epilogue %d", epi);
  if (epi == kEpiResidLN) {
    HS_REQUIRE(a.N <= 256, "gemm: residual/LayerNorm epilogue needs the row in one tile (N=%d > 256)", a.N);
    HS_REQUIRE(a.resid != nullptr, "gemm: residual epilogue without residual");
  }
  if (epi == kEpiSwiGLU) HS_REQUIRE(a.N % 32 == 0 && a.out1 != nullptr, "gemm: SwiGLU epilogue needs N%%32==0 and out1");
  if (epi != kEpiSwiGLU) HS_REQUIRE(a.out0 != nullptr, "gemm: missing output");
  if (epi == kEpiDSwiGLU) HS_REQUIRE(a.ab != nullptr, "gemm: dSwiGLU epilogue needs saved pre-activations");
  if (epi == kEpiDGate) {
    HS_REQUIRE(a.A2 != nullptr && a.B2 != nullptr && a.out0 != nullptr, "gemm: d-gate epilogue needs the gate input and w1|w3");
    HS_REQUIRE(a.K <= 256 && a.lda2 % 8 == 0 && a.ldb2 % 8 == 0, "gemm: d-gate epilogue needs K <= 256 (K=%d) and 16-byte aligned rows", a.K);
  }
  return kOk;
}

int pick_block_n(int N, int K, int epi) {
  if (epi == kEpiResidLN) return N;   // LayerNorm needs the whole row in one tile
  if (N <= 128) return N;
  // long reductions are MMA/operand-bound: full-width 256-column MMAs, A streamed once
  if (K >= 512 && N % 256 == 0 && (epi == kEpiBiasBf16 || epi == kEpiBiasF32)) return 256;
  // short reductions with a wide output run A-resident (use_ares): 256-column MMAs drained as two 128-column halves
  static const int wide = getenv("HSIMAE_GEMM_ARES_N") ? atoi(getenv("HSIMAE_GEMM_ARES_N")) : 256;
  static const int wide_gate = getenv("HSIMAE_GEMM_ARES_N_GATE") ? atoi(getenv("HSIMAE_GEMM_ARES_N_GATE")) : 128;
  if (K <= 256 && N >= 512 && epi == kEpiBiasBf16 && wide == 256) return 256;
  // square short-K projections (attention-output dgrad, [M,256] x [256,256]): HSIMAE_GEMM_N256=1 runs them as one 256-column
  // tile on the streaming pair kernel instead of two 128-column tiles on the A-resident one -- measured slower (18.9 vs
  // 17.9 us, step 22.4 vs 22.1 ms), kept as a switch
  static const int n256 = getenv("HSIMAE_GEMM_N256") ? atoi(getenv("HSIMAE_GEMM_N256")) : 0;
  if (n256 && K <= 256 && N == 256 && epi == kEpiBiasBf16) return 256;
  if (K <= 256 && N >= 512 && (epi == kEpiSwiGLU || epi == kEpiDSwiGLU) && wide_gate == 256) return 256;
  // otherwise four 128-column accumulator stages (four epilogue warp groups).  Tiles start at multiples of 128 so
  // the 128-byte output boxes never straddle two tiles; the N tail is clipped by TMA.
  return 128;
}

int gemm_tc(const GemmArgs& a, int epi, cudaStream_t stream) {
  HS_TRY(gemm_check_args(a, epi));
  if (epi == kEpiDGate) {
    static const int pair_mode = getenv("HSIMAE_GEMM_PAIR") ? atoi(getenv("HSIMAE_GEMM_PAIR")) : 1;
    // pairs halve the weight tiles in shared memory, which is what leaves room for a useful ring next to two A blocks
    if (pair_mode != 0 && (a.M > kBlockM * kNumSMs || (pair_mode >= 2 && a.M > kBlockM))) return launch_gemm_dgate<2>(a, stream);
    return launch_gemm_dgate<1>(a, stream);
  }
  const int block_n = pick_block_n(a.N, a.K, epi);
  const int n_blks = ceil_div(a.N, block_n);
  const int m_blks = ceil_div(a.M, kBlockM);
  const bool ares = use_ares(a, epi, block_n, n_blks);
  // the 256-column residual epilogue keeps 96 KB of residual boxes: only a pair's half-size weight tiles leave a useful ring
  const bool pair = use_pair(block_n, m_blks, a.K, a.N, ares, epi == kEpiSwiGLU && a.out0 == nullptr) ||
                    ((epi == kEpiResidLN || epi == kEpiBiasBf16) && block_n > 128 && use_pair(block_n, m_blks, 512, a.N, false, false));
  CUtensorMap tm[6];
  HS_TRY(get_tmap(a.A, (uint64_t)a.K, (uint64_t)a.M, (uint64_t)a.lda, 64, kBlockM, &tm[0]));
  HS_TRY(get_tmap(a.B, (uint64_t)a.K, (uint64_t)a.N, (uint64_t)a.ldb, 64, (uint32_t)(pair ? block_n / 2 : block_n), &tm[1]));
  tm[4] = tm[0];   // residual map: only the residual / LayerNorm epilogue has one
  tm[5] = tm[0];   // LayerNorm-input map: only gemm_tc_lnbwd has one
  // output boxes: [32 rows x 128 bytes]
  const uint64_t M = (uint64_t)a.M, N = (uint64_t)a.N;
  switch (epi) {
    case kEpiBiasBf16:
      HS_TRY(get_tmap(a.out0, N, M, (uint64_t)a.ld0, 64, 32, &tm[2]));
      tm[3] = tm[2];
      return launch_gemm<kEpiBiasBf16>(a, tm, block_n, n_blks, m_blks, ares, pair, stream);
    case kEpiBiasF32:
      HS_TRY(get_tmap(a.out0, N, M, (uint64_t)a.ld0, 32, 32, &tm[2], 4));
      tm[3] = tm[2];
      return launch_gemm<kEpiBiasF32>(a, tm, block_n, n_blks, m_blks, false, pair, stream);
    case kEpiResidLN:
      HS_TRY(get_tmap(a.out0, N, M, (uint64_t)a.ld0, 32, 32, &tm[2], 4));
      if (a.gamma) HS_TRY(get_tmap(a.out1, N, M, (uint64_t)a.ld1, 64, 32, &tm[3]));
      else tm[3] = tm[2];
      HS_TRY(get_tmap(a.resid, N, M, (uint64_t)a.ldr, 32, 32, &tm[4], 4));
      return launch_gemm<kEpiResidLN>(a, tm, block_n, n_blks, m_blks, false, pair, stream);
    case kEpiSwiGLU:
      HS_TRY(get_tmap(a.out1, N / 2, M, (uint64_t)a.ld1, 64, 32, &tm[3]));
      if (a.out0) HS_TRY(get_tmap(a.out0, N, M, (uint64_t)a.ld0, 64, 32, &tm[2]));
      else tm[2] = tm[3];
      return launch_gemm<kEpiSwiGLU>(a, tm, block_n, n_blks, m_blks, ares, pair, stream);
    case kEpiDSwiGLU:
      HS_TRY(get_tmap(a.out0, 2 * N, M, (uint64_t)a.ld0, 64, 32, &tm[2]));
      tm[3] = tm[2];
      return launch_gemm<kEpiDSwiGLU>(a, tm, block_n, n_blks, m_blks, ares, pair, stream);
  }
  set_error("gemm: bad epilogue %d", epi);
  return kInvalidArgument;
}

// dgrad + LayerNorm backward in one launch (kEpiLnBwd): 256-column tiles, two accumulator stages, whole rows per tile.
// HSIMAE_LNBWD_FUSE=0 makes the engine fall back to the dgrad GEMM followed by ln_bwd_vec_kernel (A/B measurements, tests).
bool gemm_lnbwd_supported(const GemmArgs& a);
// Policy (engine): on by default from two waves of row tiles up.  Below that the epilogue's sixteen serial box steps of a tile are
// not hidden behind another tile's (fine-tuning step, 32 + 71 samples: 6.6 ms with two launches, 7.5 ms fused); at the pretraining
// batch the fused form wins (21.5 -> 20.5 ms per step).
bool gemm_lnbwd_preferred(const GemmArgs& a) {
  static const bool enabled = !(getenv("HSIMAE_LNBWD_FUSE") && atoi(getenv("HSIMAE_LNBWD_FUSE")) == 0);
  static const int min_n = getenv("HSIMAE_LNBWD_MIN_N") ? atoi(getenv("HSIMAE_LNBWD_MIN_N")) : 32;
  static const int min_rows = getenv("HSIMAE_LNBWD_MIN_ROWS") ? atoi(getenv("HSIMAE_LNBWD_MIN_ROWS")) : 2 * kNumSMs * kBlockM;
  return enabled && a.N >= min_n && a.M >= min_rows && gemm_lnbwd_supported(a);
}

bool gemm_lnbwd_supported(const GemmArgs& a) {
  return a.N >= 32 && a.N <= 256 && a.N % 32 == 0 && a.M > 0 && a.K > 0 && a.lda % 8 == 0 && a.ldb % 8 == 0 &&
         a.ld0 % 4 == 0 && a.ldr % 4 == 0 && a.ldx % 4 == 0 && (a.out1 == nullptr || a.ld1 % 8 == 0);
}

int gemm_tc_lnbwd(const GemmArgs& a, cudaStream_t stream) {
  HS_REQUIRE(gemm_lnbwd_supported(a), "gemm(ln-bwd): unsupported shape M=%d N=%d K=%d", a.M, a.N, a.K);
  HS_REQUIRE(a.A && a.B && a.out0 && a.resid && a.lnx && a.stats && a.gamma, "gemm(ln-bwd): null argument");
  const int block_n = a.N, n_blks = 1;
  const int m_blks = ceil_div(a.M, kBlockM);
  const bool pair = use_pair(block_n, m_blks, 512, a.N, false, false);
  const uint64_t M = (uint64_t)a.M, N = (uint64_t)a.N;
  CUtensorMap tm[6];
  HS_TRY(get_tmap(a.A, (uint64_t)a.K, M, (uint64_t)a.lda, 64, kBlockM, &tm[0]));
  HS_TRY(get_tmap(a.B, (uint64_t)a.K, N, (uint64_t)a.ldb, 64, (uint32_t)(pair ? block_n / 2 : block_n), &tm[1]));
  HS_TRY(get_tmap(a.out0, N, M, (uint64_t)a.ld0, 32, 32, &tm[2], 4));
  if (a.out1) HS_TRY(get_tmap(a.out1, N, M, (uint64_t)a.ld1, 64, 32, &tm[3]));
  else tm[3] = tm[2];
  HS_TRY(get_tmap(a.resid, N, M, (uint64_t)a.ldr, 32, 32, &tm[4], 4));
  HS_TRY(get_tmap(a.lnx, N, M, (uint64_t)a.ldx, 32, 32, &tm[5], 4));
  return pair ? launch_gemm_s<kEpiLnBwd, 2, 2>(a, tm, block_n, n_blks, m_blks, stream)
              : launch_gemm_s<kEpiLnBwd, 2, 1>(a, tm, block_n, n_blks, m_blks, stream);
}

int wgrad_check_args(const WgradArgs& a) {
  HS_REQUIRE(a.Mred > 0 && a.Nout > 0 && a.Kin > 0, "wgrad: empty problem");
  HS_REQUIRE(a.ldy % 8 == 0 && a.ldx % 8 == 0, "wgrad: ldy=%d ldx=%d must be multiples of 8", a.ldy, a.ldx);
  HS_REQUIRE(a.ld % 4 == 0 && a.cols_valid % 4 == 0, "wgrad: destination ld=%d cols=%d must be multiples of 4", a.ld, a.cols_valid);
  HS_REQUIRE(a.row_map == 0 || (a.row_map == 1 && a.dst1 != nullptr), "wgrad: bad row map");
  return kOk;
}

int launch_colsum(const WgradArgs& a, cudaStream_t stream) {
  if (!a.bias0) return kOk;
  int gx = ceil_div(a.Nout, 64);
  int gy = ceil_div(2 * kNumSMs, gx);
  int rows = ceil_div(a.Mred, gy);
  if (rows < 64) rows = 64;
  gy = ceil_div(a.Mred, rows);
  colsum_kernel<<<dim3(gx, gy), 256, 0, stream>>>(a.Y, a.ldy, a.Mred, a.Nout, rows, a.bias0, a.bias1, a.row_map, a.rows_valid);
  HS_CHECK_LAUNCH("colsum_kernel");
  return kOk;
}

template <int P>
int launch_wgrad(const WgradArgs& a, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    HS_CHECK_CUDA(cudaFuncSetAttribute(wgrad_tc_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  // output-column tile: multiple of 64, at most 256
  int kin64 = ceil_div(a.Kin, 64) * 64;
  int tiles_c = ceil_div(kin64, 256);
  int bn = ceil_div(kin64 / 64, tiles_c) * 64;
  if (P == 2 && bn % 128 != 0) bn += 64;       // each CTA of a pair stages whole 64-column boxes
  const int tiles_r = ceil_div(a.Nout, kBlockM * P);
  const int num_tiles = tiles_r * tiles_c;
  const int kb_total = ceil_div(a.Mred, 64);
  int splits = kNumSMs / (num_tiles * P);   // one wave: never more CTAs than SMs
  if (splits < 1) splits = 1;
  if (splits > kb_total) splits = kb_total;
  const int kb_per = ceil_div(kb_total, splits);
  splits = ceil_div(kb_total, kb_per);
  const int stage_bytes = (2 + bn / P / 64) * kBoxBytes;
  int stages = (kSmemBudget - kOnesBytes) / stage_bytes;
  if (stages > 8) stages = 8;
  const size_t smem = (size_t)stages * stage_bytes + kOnesBytes + 1024 + 256;
  CUtensorMap tmY, tmX;
  HS_TRY(get_tmap(a.Y, (uint64_t)a.Nout, (uint64_t)a.Mred, (uint64_t)a.ldy, 64, 64, &tmY));
  HS_TRY(get_tmap(a.X, (uint64_t)a.Kin, (uint64_t)a.Mred, (uint64_t)a.ldx, 64, 64, &tmX));
  // MN-major SWIZZLE_128B canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units:
  // LBO = distance between 64-element atoms along M/N (one TMA box = 8 KB),
  // SBO = distance between groups of 8 reduction rows (1 KB); a K=16 step spans two groups.
  const uint32_t lbo = kBoxBytes, sbo = 1024, kstep = 2048;
  // bias gradients are fused (ones-operand MMA); HSIMAE_WGRAD_FUSED_BIAS=0 selects the separate column-sum kernel (A/B debugging)
  static const bool fused = env_int("HSIMAE_WGRAD_FUSED_BIAS", 1) != 0;
  WgradArgs k = a;
  if (!fused) { k.bias0 = nullptr; k.bias1 = nullptr; }
  HS_TRY(launch_clustered(wgrad_tc_kernel<P>, num_tiles * splits * P, kGemmThreads, smem, P, stream, tmY, tmX, k, bn, stages, tiles_c,
                          num_tiles, kb_total, kb_per, lbo, sbo, kstep));
  HS_CHECK_LAUNCH("wgrad_tc_kernel");
  if (!fused) return launch_colsum(a, stream);
  return kOk;
}


// output-column tile of a weight-gradient problem: multiple of 64 (of 128 for CTA pairs), at most 256
void wgrad_tiling(const WgradArgs& a, int P, int* bn, int* tiles_c, int* tiles_r) {
  const int kin64 = ceil_div(a.Kin, 64) * 64;
  *tiles_c = ceil_div(kin64, 256);
  *bn = ceil_div(kin64 / 64, *tiles_c) * 64;
  if (P == 2 && *bn % 128 != 0) *bn += 64;
  *tiles_r = ceil_div(a.Nout, kBlockM * P);
}

template <int P>
int launch_wgrad_group(const WgradArgs* jobs, int njobs, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    HS_CHECK_CUDA(cudaFuncSetAttribute(wgrad_group_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured = true;
  }
  WgradGroupParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.njobs = njobs;
  int tiles[kMaxWgradJobs], total_tiles = 0, bn_max = 0;
  for (int j = 0; j < njobs; ++j) {
    WgradJob& J = gp.job[j];
    J.p = jobs[j];
    int tiles_r;
    wgrad_tiling(jobs[j], P, &J.bn, &J.tiles_c, &tiles_r);
    J.num_tiles = tiles_r * J.tiles_c;
    J.kb_total = ceil_div(jobs[j].Mred, 64);
    tiles[j] = J.num_tiles; total_tiles += J.num_tiles;
    if (J.bn > bn_max) bn_max = J.bn;
    HS_TRY(get_tmap(jobs[j].Y, (uint64_t)jobs[j].Nout, (uint64_t)jobs[j].Mred, (uint64_t)jobs[j].ldy, 64, 64, &gp.tmY[j]));
    HS_TRY(get_tmap(jobs[j].X, (uint64_t)jobs[j].Kin, (uint64_t)jobs[j].Mred, (uint64_t)jobs[j].ldx, 64, 64, &gp.tmX[j]));
  }
  // one wave: split every tile's reduction so that the launch has at most one CTA (pair) per SM (pair); every tile is the
  // same amount of work per reduction row, so the splits are handed out evenly and the remainder goes to the first jobs
  const int units_max = kNumSMs / P;
  int base = units_max / total_tiles; if (base < 1) base = 1;
  int extra = base * total_tiles < units_max ? (units_max - base * total_tiles) : 0;
  int unit = 0;
  for (int j = 0; j < njobs; ++j) {
    WgradJob& J = gp.job[j];
    int splits = base;
    if (extra >= tiles[j]) { splits += 1; extra -= tiles[j]; }
    if (splits > J.kb_total) splits = J.kb_total;
    J.kb_per_split = ceil_div(J.kb_total, splits);
    splits = ceil_div(J.kb_total, J.kb_per_split);
    J.unit0 = unit;
    unit += J.num_tiles * splits;
  }
  const int stage_bytes = (2 + bn_max / P / 64) * kBoxBytes;
  int stages = (kSmemBudget - kOnesBytes) / stage_bytes;
  if (stages > 8) stages = 8;
  const size_t smem = (size_t)stages * stage_bytes + kOnesBytes + 1024 + 256;
  const uint32_t lbo = kBoxBytes, sbo = 1024, kstep = 2048;   // MN-major SWIZZLE_128B canonical layout, see launch_wgrad
  HS_TRY(launch_clustered(wgrad_group_kernel<P>, unit * P, kGemmThreads, smem, P, stream, gp, stages, (uint32_t)stage_bytes, lbo, sbo, kstep));
  HS_CHECK_LAUNCH("wgrad_group_kernel");
  return kOk;
}

int wgrad_tc(const WgradArgs& a, cudaStream_t stream) {
  HS_TRY(wgrad_check_args(a));
  // CTA pairs halve the L2 reads of X (re-read by every row tile); they need at least two row tiles' worth of rows
  static const int pair = env_int("HSIMAE_WGRAD_PAIR", 1);
  if (pair != 0 && a.Nout > 2 * kBlockM && a.Mred >= 64 * kNumSMs) return launch_wgrad<2>(a, stream);
  return launch_wgrad<1>(a, stream);
}

// The weight gradients of one block in one launch.  HSIMAE_WGRAD_GROUP=0: one launch per problem (A/B measurements).
int wgrad_tc_group(const WgradArgs* jobs, int njobs, cudaStream_t stream) {
  HS_REQUIRE(njobs >= 1 && njobs <= kMaxWgradJobs, "wgrad group: %d jobs (1..%d supported)", njobs, kMaxWgradJobs);
  for (int j = 0; j < njobs; ++j) HS_TRY(wgrad_check_args(jobs[j]));
  static const int group = env_int("HSIMAE_WGRAD_GROUP", 1);
  static const bool fused_bias = env_int("HSIMAE_WGRAD_FUSED_BIAS", 1) != 0;
  if (group == 0 || !fused_bias || njobs == 1) {
    for (int j = 0; j < njobs; ++j) HS_TRY(wgrad_tc(jobs[j], stream));
    return kOk;
  }
  static const int pair = env_int("HSIMAE_WGRAD_PAIR", 1);
  bool any_tall = false;
  int mred_min = jobs[0].Mred;
  for (int j = 0; j < njobs; ++j) { any_tall |= jobs[j].Nout > 2 * kBlockM; if (jobs[j].Mred < mred_min) mred_min = jobs[j].Mred; }
  if (pair != 0 && any_tall && mred_min >= 64 * kNumSMs) return launch_wgrad_group<2>(jobs, njobs, stream);
  return launch_wgrad_group<1>(jobs, njobs, stream);
}

}  // namespace hsimae

#ifdef HSIMAE_TRACE
extern "C" int hsimae_debug_trace(long long* host_out, int n_ll) {
  return (int)cudaMemcpyFromSymbol(host_out, hsimae::g_trace, (size_t)n_ll * sizeof(long long));
}
#endif
