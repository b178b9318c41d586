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
// Pipeline per CTA (P = 2: CTA pair, cta_group::2 MMAs over both shared memories, each CTA streams HALF of every weight
// tile; the leader issues):
//   warp 0  TMA producer: the tile's A block (once per tile) and the w1|w3 ring (one [128/P x 64] box per slot)
//   warp 1  MMA issuer:   G1(c): ab[c%2] = A * W13[c]^T   (N = 128: 64 hidden units, a|b interleaved by 16)
//                         G2(c): out += g[c%2] * W2[:, c]^T (N = d, K = 64)
//                         issue order G1(0) G1(1) | G2(0) G1(2) | G2(1) G1(3) | ...  (the gate epilogue of chunk c runs
//                         under G2(c-1) + G1(c+1))
//   warp 2  TMA producer: the w2 ring (one [d/P x 64] box per slot)
//   warp 3  gate-output store (training): one TMA store per chunk straight from the MMA operand buffer
//   warps 4-11  gate epilogue: 8 warps per chunk, each group of 4 takes 32 of the 64 hidden units
//   warps 12-15 final epilogue: the residual / LayerNorm epilogue of gemm_tc.cu on the [128 x d] accumulator
#include "tc_device.cuh"
#include "block_fused.cuh"

namespace hsimae {

namespace ptx {
// cluster-scope release / acquire: the gate tile a peer CTA wrote into ITS shared memory is read by the pair MMA the
// leader issues, so the "chunk done" handshake must order memory across the two CTAs (one elected lane per warp pays
// for the fence, not every thread)
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_acq_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_acq_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_acq_cluster(bar, parity)) {
    if (++spins > (1u << 26)) { printf("hsimae: mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}
}  // namespace ptx

constexpr int kMlpThreads = 512;
constexpr int kMlpChunk = 64;                       // hidden units per chunk (128 interleaved a|b columns)
constexpr uint32_t kGBufBytes = kBlockM * 128;      // gate tile [128 rows x 64 bf16], 128B-swizzled K-major
constexpr uint32_t kMlpAbCol = 256;                 // TMEM: [0, d) output accumulator, 256 + 128 s: a|b stage s

template <int P>
__global__ void __launch_bounds__(kMlpThreads, 1)
mlp_fused_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW13,
                 const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmG,
                 const __grid_constant__ CUtensorMap tmO0, const __grid_constant__ CUtensorMap tmO1, MlpFusedArgs p,
                 int n13, int n2, int nch, int m_units) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);

  const int d = p.tail.N, Hp = p.tail.K;
  const int num_kb = d / kBlockK;
  const uint32_t a_bytes = (uint32_t)num_kb * kATileBytes;
  const uint32_t slot13 = (uint32_t)(2 * kMlpChunk / P) * 128u;   // [128/P rows of w1|w3] x [64 K]
  const uint32_t slot2 = (uint32_t)(d / P) * 128u;                // [d/P rows of w2] x [64 hidden]
  uint8_t* a_res = smem;
  uint8_t* ring13 = a_res + a_bytes;
  uint8_t* ring2 = ring13 + (size_t)n13 * slot13;
  uint8_t* gbuf = ring2 + (size_t)n2 * slot2;
  uint8_t* staging = gbuf + 2 * kGBufBytes;
  uint64_t* full13 = reinterpret_cast<uint64_t*>(staging + 4 * kStageBufBytes);
  uint64_t* empty13 = full13 + n13;
  uint64_t* full2 = empty13 + n13;
  uint64_t* empty2 = full2 + n2;
  uint64_t* abfull = empty2 + n2;      // [2] G1 of the chunk retired (both CTAs, multicast commit)
  uint64_t* chunk_done = abfull + 2;   // [2] leader: the gate tiles of both CTAs are in shared memory, a|b stage drained
  uint64_t* gempty = chunk_done + 2;   // [2] G2 of the chunk retired: the gate buffer may be rewritten
  uint64_t* gready = gempty + 2;       // [2] local: this CTA's gate tile is complete (-> store warp)
  uint64_t* gstored = gready + 2;      // [2] local: the TMA store has read the gate tile
  uint64_t* a_full = gstored + 2;
  uint64_t* a_empty = a_full + 1;
  uint64_t* out_full = a_empty + 1;
  uint64_t* out_empty = out_full + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_empty + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int crank = P == 2 ? (int)ptx::cluster_ctarank() : 0;
  const int unit0 = (int)blockIdx.x / P, unit_step = (int)gridDim.x / P;
  const bool save_g = p.g != nullptr;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&tmX); ptx::prefetch_tmap(&tmW13); ptx::prefetch_tmap(&tmW2); ptx::prefetch_tmap(&tmG);
    ptx::prefetch_tmap(&tmO0); ptx::prefetch_tmap(&tmO1);
    for (int i = 0; i < n13; ++i) { ptx::mbar_init(full13 + i, 1); ptx::mbar_init(empty13 + i, 1); }
    for (int i = 0; i < n2; ++i) { ptx::mbar_init(full2 + i, 1); ptx::mbar_init(empty2 + i, 1); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(abfull + i, 1); ptx::mbar_init(chunk_done + i, 8 * P); ptx::mbar_init(gempty + i, 1);
      ptx::mbar_init(gready + i, 8); ptx::mbar_init(gstored + i, 1);
    }
    ptx::mbar_init(a_full, 1); ptx::mbar_init(a_empty, 1); ptx::mbar_init(out_full, 1); ptx::mbar_init(out_empty, 128 * P);
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
    // ---- A block + w1|w3 ring --------------------------------------------------------------------------------
    if (ptx::elect_one()) {
      int slot = 0; uint32_t phase = 0, aph = 0;
      const uint32_t a_full_addr = P == 2 ? ptx::mapa_rank(ptx::smem_u32(a_full), 0) : 0u;
      for (int mu = unit0; mu < m_units; mu += unit_step) {
        const int m0 = (mu * P + crank) * kBlockM;   // may lie past the last row: TMA zero-fills the load, clips the stores
        if (mu != unit0) ptx::mbar_wait(a_empty, aph ^ 1u);   // every G1 of the previous tile has retired
        if constexpr (P == 2) {
          if (crank == 0) ptx::mbar_expect_tx(a_full, 2u * a_bytes);
          for (int kb = 0; kb < num_kb; ++kb) ptx::tma_load_2d_pair(a_res + (size_t)kb * kATileBytes, &tmX, a_full_addr, kb * kBlockK, m0);
        } else {
          ptx::mbar_expect_tx(a_full, a_bytes);
          for (int kb = 0; kb < num_kb; ++kb) ptx::tma_load_2d(a_res + (size_t)kb * kATileBytes, &tmX, a_full, kb * kBlockK, m0);
        }
        aph ^= 1u;
        for (int c = 0; c < nch; ++c)
          for (int kb = 0; kb < num_kb; ++kb) {
            ptx::mbar_wait(empty13 + slot, phase ^ 1u);
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
      }
    }
  } else if (warp == 2) {
    // ---- w2 ring -----------------------------------------------------------------------------------------------
    if (ptx::elect_one()) {
      int slot = 0; uint32_t phase = 0;
      for (int mu = unit0; mu < m_units; mu += unit_step)
        for (int c = 0; c < nch; ++c) {
          ptx::mbar_wait(empty2 + slot, phase ^ 1u);
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
    }
  } else if (warp == 1) {
    // ---- MMA issuer (leader CTA) ----------------------------------------------------------------------------------
    if (crank == 0 && ptx::elect_one()) {
      const uint32_t idesc_ab = make_idesc(2 * kMlpChunk, false, false, kBlockM * P);
      const uint32_t idesc_out = make_idesc(d, false, false, kBlockM * P);
      const uint64_t adesc0 = make_smem_desc(ptx::smem_u32(a_res), 16, 1024);
      const uint64_t b13desc0 = make_smem_desc(ptx::smem_u32(ring13), 16, 1024);
      const uint64_t b2desc0 = make_smem_desc(ptx::smem_u32(ring2), 16, 1024);
      const uint64_t gdesc0 = make_smem_desc(ptx::smem_u32(gbuf), 16, 1024);
      int s13 = 0, s2 = 0; uint32_t ph13 = 0, ph2 = 0;
      int gi0 = 0, t = 0;   // global chunk index of the tile's first chunk, tile counter
      auto commit = [](uint64_t* bar) { if constexpr (P == 2) ptx::umma_commit_pair(bar); else ptx::umma_commit(bar); };
      auto g1 = [&](int c) {
        const int s = (gi0 + c) & 1;
        const uint32_t d_ab = tmem_base + kMlpAbCol + (uint32_t)s * 128u;
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(full13 + s13, ph13);
          ptx::tc_fence_after();
          const uint64_t adesc = adesc0 + (uint64_t)(kb * (kATileBytes >> 4));
          const uint64_t bdesc = b13desc0 + (uint64_t)((uint32_t)s13 * (slot13 >> 4));
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k) {
            if constexpr (P == 2) ptx::umma_bf16_pair(d_ab, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_ab, (kb | k) != 0 ? 1u : 0u);
            else ptx::umma_bf16(d_ab, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_ab, (kb | k) != 0 ? 1u : 0u);
          }
          commit(empty13 + s13);
          if (++s13 == n13) { s13 = 0; ph13 ^= 1u; }
        }
        commit(abfull + s);
        if (c == nch - 1) commit(a_empty);
      };
      auto g2 = [&](int c) {
        const int gi = gi0 + c, s = gi & 1;
        ptx::mbar_wait_acq_cluster(chunk_done + s, (uint32_t)(gi >> 1) & 1u);
        if (c == 0 && t > 0) ptx::mbar_wait(out_empty, (uint32_t)(t - 1) & 1u);   // the previous tile's accumulator has been drained
        ptx::mbar_wait(full2 + s2, ph2);
        ptx::tc_fence_after();
        const uint64_t adesc = gdesc0 + (uint64_t)((uint32_t)s * (kGBufBytes >> 4));
        const uint64_t bdesc = b2desc0 + (uint64_t)((uint32_t)s2 * (slot2 >> 4));
#pragma unroll
        for (int k = 0; k < kMlpChunk / 16; ++k) {
          if constexpr (P == 2) ptx::umma_bf16_pair(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_out, (c | k) != 0 ? 1u : 0u);
          else ptx::umma_bf16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc_out, (c | k) != 0 ? 1u : 0u);
        }
        commit(empty2 + s2);
        if (++s2 == n2) { s2 = 0; ph2 ^= 1u; }
        commit(gempty + s);
        if (c == nch - 1) commit(out_full);
      };
      for (int mu = unit0; mu < m_units; mu += unit_step, ++t) {
        ptx::mbar_wait(a_full, (uint32_t)t & 1u);
        ptx::tc_fence_after();
        g1(0);
        if (nch > 1) g1(1);
        for (int c = 0; c < nch; ++c) {
          g2(c);
          if (c + 2 < nch) g1(c + 2);
        }
        gi0 += nch;
      }
    }
  } else if (warp == 3) {
    // ---- gate output -> HBM (training): the MMA operand tile is a legal TMA box ------------------------------------
    if (save_g && ptx::elect_one()) {
      int gi = 0;
      for (int mu = unit0; mu < m_units; mu += unit_step) {
        const int m0 = (mu * P + crank) * kBlockM;
        for (int c = 0; c < nch; ++c, ++gi) {
          const int s = gi & 1;
          ptx::mbar_wait(gready + s, (uint32_t)(gi >> 1) & 1u);
          ptx::tma_store_2d(&tmG, ptx::smem_u32(gbuf) + (uint32_t)s * kGBufBytes, c * kMlpChunk, m0);
          ptx::bulk_commit();
          if (gi > 0) { ptx::bulk_wait_read1(); ptx::mbar_arrive(gstored + ((gi - 1) & 1)); }
        }
      }
      ptx::bulk_wait_read0();
    }
  } else if (warp < 12) {
    // ---- gate epilogue: g = silu(a) * b for 32 of the chunk's 64 hidden units, written as the A operand of G2 ------
    const int q = warp & 3;
    const int grp = (warp - 4) >> 2;
    const bool leader = ptx::elect_one();
    const uint32_t done_addr0 = P == 2 ? ptx::mapa_rank(ptx::smem_u32(chunk_done), 0) : ptx::smem_u32(chunk_done);
    const uint32_t row_off = (uint32_t)(q * 32 + lane) * 128u;
    const uint32_t sw = (uint32_t)(lane & 7);
    int gi = 0;
    for (int mu = unit0; mu < m_units; mu += unit_step) {
      for (int c = 0; c < nch; ++c, ++gi) {
        const int s = gi & 1;
        const uint32_t ph = (uint32_t)(gi >> 1) & 1u;
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + kMlpAbCol + (uint32_t)s * 128u + (uint32_t)grp * 64u;
        const uint32_t grow = ptx::smem_u32(gbuf) + (uint32_t)s * kGBufBytes + row_off;
        ptx::mbar_wait(abfull + s, ph);
        ptx::tc_fence_after();
        uint4 o[4];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int h0 = c * kMlpChunk + grp * 32 + j * 16;
          if (h0 < Hp) {
            float v[32], g[16];
            ptx::tmem_ld32(tacc + 32u * j, v);
            ptx::tmem_ld_wait();
            add_vec<32>(p.b13 + 2 * h0, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) g[i] = v[i] * ptx::sigmoid_fast(v[i]) * v[16 + i];
            o[2 * j] = pack8_bf16(g); o[2 * j + 1] = pack8_bf16(g + 8);
          } else {
            o[2 * j] = make_uint4(0u, 0u, 0u, 0u); o[2 * j + 1] = make_uint4(0u, 0u, 0u, 0u);   // padding: finite operand for G2
          }
        }
        // the buffer's previous tile (chunk gi - 2) has been consumed by G2 and, when it is kept, read by the store
        ptx::mbar_wait(gempty + s, ph ^ 1u);
        if (save_g) ptx::mbar_wait(gstored + s, ph ^ 1u);
#pragma unroll
        for (int i = 0; i < 4; ++i) ptx::st_shared_v4(grow + ((((uint32_t)(grp * 4 + i)) ^ sw) << 4), o[i]);
        ptx::fence_proxy_async();   // generic-proxy writes -> visible to the tensor core / TMA (async proxy)
        ptx::tc_fence_before();
        __syncwarp();
        if (leader) {
          if (save_g) ptx::mbar_arrive(gready + s);
          if constexpr (P == 2) ptx::mbar_arrive_release_cluster(done_addr0 + (uint32_t)s * 8u);
          else ptx::mbar_arrive(chunk_done + s);
        }
      }
    }
  } else {
    // ---- final epilogue: + bias + residual (+ other branch), LayerNorm of the next consumer --------------------------
    const int q = warp & 3;
    Stager st{ptx::smem_u32(staging) + (uint32_t)(warp - 12) * kStageBufBytes, lane, false};
    const uint32_t out_empty_addr = P == 2 ? ptx::mapa_rank(ptx::smem_u32(out_empty), 0) : 0u;
    int t = 0;
    for (int mu = unit0; mu < m_units; mu += unit_step, ++t) {
      TmemAcc acc{tmem_base + ((uint32_t)(q * 32) << 16)};
      tc_epilogue<kEpiResidLN, 1>(p.tail, acc, st, &tmO0, &tmO1, (mu * P + crank) * kBlockM + q * 32, lane, 0, d,
                                  [&]() { ptx::mbar_wait(out_full, (uint32_t)t & 1u); ptx::tc_fence_after(); });
      ptx::tc_fence_before();
      if constexpr (P == 2) ptx::mbar_arrive_cluster(out_empty_addr); else ptx::mbar_arrive(out_empty);
    }
    st.acquire();
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

// shared memory: [A block][w1|w3 ring][w2 ring][2 gate tiles][4 staging boxes][barriers]
bool mlp_smem_plan(int d, int P, MlpSmem* out) {
  const int num_kb = d / kBlockK;
  const int a_bytes = num_kb * kATileBytes;
  const int slot13 = 2 * kMlpChunk / P * 128, slot2 = d / P * 128;
  const int fixed = a_bytes + 2 * (int)kGBufBytes + 4 * kStageBufBytes + 2048;
  int n2 = 2;
  int n13 = (kSmemMax - fixed - n2 * slot2) / slot13;
  if (n13 < num_kb) { n2 = 1; n13 = (kSmemMax - fixed - n2 * slot2) / slot13; }
  if (n13 > 2 * num_kb) n13 = 2 * num_kb;   // two chunks of w1|w3 in flight
  if (n13 > 8) n13 = 8;
  if (n13 < 1) return false;
  // spend what is left on the w2 ring
  while (n2 < 4 && fixed + n13 * slot13 + (n2 + 1) * slot2 <= kSmemMax) ++n2;
  out->n13 = n13; out->n2 = n2;
  out->bytes = (size_t)fixed + (size_t)n13 * slot13 + (size_t)n2 * slot2;
  return out->bytes <= (size_t)kSmemMax;
}

template <int P>
int launch_mlp_fused(const MlpFusedArgs& a, cudaStream_t stream) {
  static bool configured = false;
  if (!configured) {
    HS_CHECK_CUDA(cudaFuncSetAttribute(mlp_fused_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemMax));
    configured = true;
  }
  const GemmArgs& t = a.tail;
  const int d = t.N, Hp = t.K;
  MlpSmem sm;
  HS_REQUIRE(mlp_smem_plan(d, P, &sm), "fused MLP: width %d does not fit in shared memory", d);
  const uint64_t M = (uint64_t)t.M;
  CUtensorMap tmX, tmW13, tmW2, tmG, tmO0, tmO1;
  HS_TRY(get_tmap(a.X, (uint64_t)d, M, (uint64_t)a.ldx, 64, kBlockM, &tmX));
  HS_TRY(get_tmap(a.W13, (uint64_t)d, (uint64_t)(2 * Hp), (uint64_t)a.ldw, 64, (uint32_t)(2 * kMlpChunk / P), &tmW13));
  HS_TRY(get_tmap(t.B, (uint64_t)Hp, (uint64_t)d, (uint64_t)t.ldb, 64, (uint32_t)(d / P), &tmW2));
  HS_TRY(get_tmap(t.out0, (uint64_t)d, M, (uint64_t)t.ld0, 32, 32, &tmO0, 4));
  if (t.gamma) HS_TRY(get_tmap(t.out1, (uint64_t)d, M, (uint64_t)t.ld1, 64, 32, &tmO1)); else tmO1 = tmO0;
  if (a.g) HS_TRY(get_tmap(a.g, (uint64_t)Hp, M, (uint64_t)a.ldg, 64, kBlockM, &tmG)); else tmG = tmX;
  const int m_units = ceil_div(ceil_div(t.M, kBlockM), P);
  HS_TRY(launch_clustered(mlp_fused_kernel<P>, pair_grid(m_units, P), kMlpThreads, sm.bytes, P, stream, tmX, tmW13, tmW2, tmG, tmO0,
                          tmO1, a, sm.n13, sm.n2, ceil_div(Hp, kMlpChunk), m_units));
  HS_CHECK_LAUNCH("mlp_fused_kernel");
  return kOk;
}

}  // namespace

bool mlp_fused_supported(int d, int Hp) {
  MlpSmem sm;
  return d % 64 == 0 && d >= 64 && d <= 256 && Hp % 16 == 0 && Hp >= 16 && mlp_smem_plan(d, 2, &sm);
}

int mlp_fused(const MlpFusedArgs& a, cudaStream_t stream) {
  const GemmArgs& t = a.tail;
  HS_REQUIRE(t.M > 0 && mlp_fused_supported(t.N, t.K), "fused MLP: unsupported shape M=%d d=%d Hp=%d", t.M, t.N, t.K);
  HS_REQUIRE(a.X && a.W13 && a.b13 && t.B && t.out0 && t.resid, "fused MLP: null argument");
  HS_REQUIRE(a.ldx % 8 == 0 && a.ldw % 8 == 0 && t.ldb % 8 == 0 && (a.g == nullptr || a.ldg % 8 == 0), "fused MLP: rows must be 16-byte aligned");
  // pairs halve the weight bytes entering each SM (the binding rate of this kernel); a single CTA only when there is
  // one row block or when asked for (HSIMAE_FUSED_MLP_PAIR=0, A/B measurements)
  static const int pair = getenv("HSIMAE_FUSED_MLP_PAIR") ? atoi(getenv("HSIMAE_FUSED_MLP_PAIR")) : 1;
  MlpSmem sm;
  const bool single_ok = mlp_smem_plan(t.N, 1, &sm);
  if (single_ok && (pair == 0 || t.M <= kBlockM)) return launch_mlp_fused<1>(a, stream);
  return launch_mlp_fused<2>(a, stream);
}

}  // namespace hsimae
