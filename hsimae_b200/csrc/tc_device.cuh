// tcgen05 / TMEM / TMA device helpers shared by the tensor-core kernels (gemm_tc.cu, block_fused.cu):
// PTX wrappers, the TMEM accumulator view, TMA-store staging, the fused epilogues and the UMMA descriptors.
#pragma once
#include <cuda.h>
#include <type_traits>
#include "epilogue.cuh"

namespace hsimae {

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// One lane of the (converged) warp.  Unlike `lane == 0`, elect.sync tells the compiler that exactly one thread runs the
// region, so tcgen05 / TMA instructions take their operands straight from uniform registers (no per-instruction
// "waterfall" loops): the MMA issue loop shrinks from ~80 to ~25 instructions per k-block.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// Programmatic dependent launch: kernels launched with the programmatic-stream-serialization attribute may start their
// prologue (barrier init, TMEM allocation, descriptor prefetch) while the previous kernel on the stream drains;
// pdl_wait() blocks until that kernel has completed and its writes are visible, pdl_trigger() lets the next one go.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 26)) { printf("hsimae: mbarrier wait timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
  }
}

__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_addr(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
// barrier among a subset of the CTA's warps (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
// ---- CTA-pair (cta_group::2) helpers: two CTAs of a cluster on one TPC drive ONE 256-row MMA; each loads its own
// 128 A rows and HALF of the B tile, the tensor core reads both shared memories.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
  uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  // default semantics (.release at CTA scope): the accumulator reads this orders are TMEM reads, fenced by
  // tcgen05.fence::before_thread_sync; a cluster-scope release would cost a GPU-wide MEMBAR per thread and tile
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// load into THIS CTA's shared memory, completion bytes signalled on a barrier that may live in the peer CTA
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrives (once the MMAs issued so far have retired) on the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}

// pull `bytes` (multiple of 16) starting at a 16-byte aligned global address into L2, asynchronously
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// silu(a) = a * sigmoid(a) with sigmoid(a) = 0.5 + 0.5 tanh(a / 2): one MUFU op instead of exp + reciprocal
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_fast(float a) { return fmaf(tanh_approx(0.5f * a), 0.5f, 0.5f); }

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
}
__device__ __forceinline__ void tmem_relinquish() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),
        "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]),
        "=f"(v[16]), "=f"(v[17]), "=f"(v[18]), "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]),
        "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]), "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]),
        "=f"(v[8]), "=f"(v[9]), "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31};"
      ::"f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
        "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]),
        "f"(v[16]), "f"(v[17]), "f"(v[18]), "f"(v[19]), "f"(v[20]), "f"(v[21]), "f"(v[22]), "f"(v[23]),
        "f"(v[24]), "f"(v[25]), "f"(v[26]), "f"(v[27]), "f"(v[28]), "f"(v[29]), "f"(v[30]), "f"(v[31]),
        "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15};"
      ::"f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]),
        "f"(v[8]), "f"(v[9]), "f"(v[10]), "f"(v[11]), "f"(v[12]), "f"(v[13]), "f"(v[14]), "f"(v[15]),
        "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace ptx

// accumulator row of this thread in TMEM (lane = row, column = n)
struct TmemAcc {
  uint32_t base;  // (lane_base << 16) | column_base
  template <int W> __device__ __forceinline__ void load(int c, float (&v)[W]) {
    if constexpr (W == 32) ptx::tmem_ld32(base + c, v); else ptx::tmem_ld16(base + c, v);
    ptx::tmem_ld_wait();
  }
  template <int W> __device__ __forceinline__ void store(int c, const float (&v)[W]) {
    if constexpr (W == 32) ptx::tmem_st32(base + c, v); else ptx::tmem_st16(base + c, v);
  }
  __device__ __forceinline__ void fence_store() { ptx::tmem_st_wait(); }
};

// ---------------------------------------------------------------------------
// Tensor-core epilogues with coalesced output: every epilogue warp owns 32
// accumulator rows; it assembles [32 rows x 128 B] output boxes in its own
// swizzled shared-memory staging buffers and hands them to the TMA store unit
// (one elected lane), which also clips the M / N tails.  Same arithmetic as
// the row-wise reference epilogues in epilogue.cuh (the checker path).
// ---------------------------------------------------------------------------
constexpr int kStageBufBytes = 32 * 128;   // one box
// staging boxes per epilogue warp: the gate epilogue fills two boxes at once (a|b and gate); the others reuse a
// single box (the TMA unit has read it long before the next one is assembled), which leaves the shared memory
// to the operand pipeline.
// The residual / LayerNorm epilogue of the two-stage (256-column) kernel rotates three boxes per warp: each is filled
// with a residual tile by TMA, updated in place and handed back to TMA as the output tile.
constexpr int kResidBoxes = 3;
template <int EPI, int S = 4> struct StagingBufs {
  // the LayerNorm-backward epilogue adds a fourth box for its bf16 output
  static constexpr int value = EPI == kEpiSwiGLU ? 2 : (EPI == kEpiResidLN && S == 2 ? kResidBoxes : (EPI == kEpiLnBwd ? kResidBoxes + 1 : 1));
};

struct Stager {
  uint32_t base;     // smem address of this warp's two 4 KB buffers (1024-byte aligned)
  int lane;
  bool pending;      // a committed store may still be reading the buffers
  bool leader = ptx::elect_one();   // the one lane that issues (and later waits for) this warp's bulk stores
  const CUtensorMap* tm_resid = nullptr;   // TMA-staged residual (kEpiResidLN, S == 2): fp32 [M, N], boxes of 32 x 32
  const CUtensorMap* tm_x = nullptr;       // kEpiLnBwd: the LayerNorm input, same box shape
  uint64_t* rbar = nullptr;                // this warp's kResidBoxes "residual box filled" barriers
  uint32_t rphase = 0;                     // their phase bits
  // wait until the TMA unit has finished reading every box this warp handed over
  __device__ __forceinline__ void acquire() {
    if (pending) {
      if (leader) ptx::bulk_wait_read0();
      __syncwarp();
      pending = false;
    }
  }
  // single-buffer use: wait until the previous store has been read, then refill buffer 0
  __device__ __forceinline__ int begin_box() { acquire(); return 0; }
  __device__ __forceinline__ void end_box(const CUtensorMap* tm, int col, int row) { flush(0, tm, col, row); }
  // 16-byte piece j (0..7) of this lane's 128-byte box row; 128B-swizzle: piece index XOR (row & 7)
  __device__ __forceinline__ void put(int b, int j, uint4 v) {
    ptx::st_shared_v4(base + (uint32_t)b * kStageBufBytes + (uint32_t)lane * 128u + (uint32_t)((j ^ (lane & 7)) << 4), v);
  }
  __device__ __forceinline__ void flush(int b, const CUtensorMap* tm, int col, int row) {
    ptx::fence_proxy_async();
    __syncwarp();
    if (leader) { ptx::tma_store_2d(tm, base + (uint32_t)b * kStageBufBytes, col, row); ptx::bulk_commit(); }
    pending = true;
  }
};

__device__ __forceinline__ uint4 pack8_bf16(const float* v) {
  uint4 t;
  t.x = pack_bf16x2(v[0], v[1]); t.y = pack_bf16x2(v[2], v[3]); t.z = pack_bf16x2(v[4], v[5]); t.w = pack_bf16x2(v[6], v[7]);
  return t;
}
__device__ __forceinline__ uint4 pack4_f32(const float* v) {
  return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
}

// A staged box is only safe to refill after the previous store from it was read; with two buffers used
// alternately the wait (issued after the chunk's math) is almost always already satisfied.
template <int W, bool F32>
__device__ __forceinline__ void tc_bias_chunk(const GemmArgs& p, TmemAcc& acc, Stager& st, const CUtensorMap* tmO, int m0, int n0,
                                              int c, int width) {
  float v[W];
  acc.template load<W>(c, v);
  if (p.bias) add_vec<W>(p.bias + n0 + c, v);
  if constexpr (F32) {
    const int b = st.begin_box();
#pragma unroll
    for (int i = 0; i < W / 4; ++i) st.put(b, i, pack4_f32(v + 4 * i));
    st.end_box(tmO, n0 + c, m0);
  } else {
    const int b = (c & 63) == 0 ? st.begin_box() : 0;
    const int j0 = (c & 63) >> 3;
#pragma unroll
    for (int i = 0; i < W / 8; ++i) st.put(b, j0 + i, pack8_bf16(v + 8 * i));
    if (((c + W) & 63) == 0 || c + W >= width) st.end_box(tmO, n0 + (c & ~63), m0);
  }
}

// `wait_acc()` blocks until the tile's accumulator is complete; epilogues that read per-row global inputs issue
// the first loads BEFORE calling it, so that latency overlaps the tail of the MMA.
template <int EPI, int MODE, class Wait>
__device__ __forceinline__ void tc_epilogue(const GemmArgs& p, TmemAcc& acc, Stager& st, const CUtensorMap* tmO0,
                                            const CUtensorMap* tmO1, int m0, int lane, int n0, int width, Wait wait_acc) {
  const int m = m0 + lane;
  const bool valid = m < p.M;
  constexpr bool PREFETCH = (MODE & 1) != 0;   // fetch the residual one chunk ahead (needs the registers of the S=2 variant)
  constexpr bool G_DIRECT = (MODE & 2) != 0;   // gate values go to global memory straight from registers (one staging box per warp)
  if constexpr (EPI == kEpiBiasBf16 || EPI == kEpiBiasF32) {
    constexpr bool F32 = EPI == kEpiBiasF32;
    wait_acc();
    int c = 0;
    for (; c + 32 <= width; c += 32) tc_bias_chunk<32, F32>(p, acc, st, tmO0, m0, n0, c, width);
    for (; c + 16 <= width; c += 16) tc_bias_chunk<16, F32>(p, acc, st, tmO0, m0, n0, c, width);
  } else if constexpr (EPI == kEpiSwiGLU) {
    // tile of packed (a|b interleaved by 16) columns; buffer 0: a|b boxes (64 packed columns), buffer 1: the gate box.
    // The pre-activations are only written when the caller keeps them (out0): the training path recomputes them in
    // backward (gemm_tc_dgate_kernel) because HBM writes (3.9 TB/s) are the scarce resource of this kernel.
    const bool keep_ab = p.out0 != nullptr;
    wait_acc();
    for (int c = 0; c + 32 <= width; c += 32) {
      float v[32];
      acc.template load<32>(c, v);
      if (p.bias) add_vec<32>(p.bias + n0 + c, v);
      float g[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        // gate on what backward will see: the bf16 copy it re-reads, or the fp32 values it recomputes
        const float a = keep_ab ? bf16_round(v[i]) : v[i], b = keep_ab ? bf16_round(v[16 + i]) : v[16 + i];
        g[i] = a * ptx::sigmoid_fast(a) * b;
      }
      if (keep_ab) {
        if ((c & 63) == 0) st.acquire();
        const int j0 = (c & 63) >> 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) st.put(0, j0 + i, pack8_bf16(v + 8 * i));
      }
      if (G_DIRECT && keep_ab) {
        // the single staging box is taken by a|b: 32 contiguous bytes per thread = one full sector
        if (valid) store_bf16_row<16>(reinterpret_cast<__nv_bfloat16*>(p.out1) + (size_t)m * p.ld1 + ((n0 + c) >> 1), g);
      } else {
        if (!keep_ab && c == 0) st.acquire();
        st.put(G_DIRECT ? 0 : 1, (c >> 5) * 2, pack8_bf16(g));
        st.put(G_DIRECT ? 0 : 1, (c >> 5) * 2 + 1, pack8_bf16(g + 8));
      }
      if (keep_ab && (((c + 32) & 63) == 0 || c + 32 >= width)) st.flush(0, tmO0, n0 + (c & ~63), m0);
    }
    if (!(G_DIRECT && keep_ab)) st.flush(G_DIRECT ? 0 : 1, tmO1, n0 >> 1, m0);
  } else if constexpr (EPI == kEpiDSwiGLU) {
    // tile of hidden columns; every 16 of them become 32 packed output columns.  The saved pre-activations of the
    // NEXT chunk are fetched while the current one is being processed (the first fetch overlaps the MMA tail).
    const __nv_bfloat16* abrow = p.ab + (size_t)(valid ? m : 0) * p.ldab + 2 * n0;
    // the tile's saved pre-activations start moving HBM -> L2 while the MMAs of the tile are still running
    if (valid) ptx::prefetch_l2_bulk(abrow, (uint32_t)width * 4u);
    uint4 nxt[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) nxt[i] = valid ? *reinterpret_cast<const uint4*>(abrow + 8 * i) : make_uint4(0u, 0u, 0u, 0u);  // L1-allocating: the 4 loads share one line
    wait_acc();
    for (int c = 0; c + 16 <= width; c += 16) {
      uint4 cur[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) cur[i] = nxt[i];
      if (c + 32 <= width && valid) {
#pragma unroll
        for (int i = 0; i < 4; ++i) nxt[i] = *reinterpret_cast<const uint4*>(abrow + 2 * (c + 16) + 8 * i);
      }
      float dg[16];
      acc.template load<16>(c, dg);
      float ab[32], o[32];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float2 x0 = unpack_bf16x2(cur[i].x), x1 = unpack_bf16x2(cur[i].y), x2 = unpack_bf16x2(cur[i].z), x3 = unpack_bf16x2(cur[i].w);
        ab[8 * i] = x0.x; ab[8 * i + 1] = x0.y; ab[8 * i + 2] = x1.x; ab[8 * i + 3] = x1.y;
        ab[8 * i + 4] = x2.x; ab[8 * i + 5] = x2.y; ab[8 * i + 6] = x3.x; ab[8 * i + 7] = x3.y;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float a = ab[i], b = ab[16 + i];
        const float sg = ptx::sigmoid_fast(a);
        o[i] = dg[i] * b * (sg * (1.0f + a * (1.0f - sg)));
        o[16 + i] = dg[i] * (a * sg);
      }
      const int b = (c & 31) == 0 ? st.begin_box() : 0;
      const int j0 = ((c & 31) >> 4) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) st.put(b, j0 + i, pack8_bf16(o + 8 * i));
      if (((c + 16) & 31) == 0 || c + 16 >= width) st.end_box(tmO0, 2 * (n0 + (c & ~31)), m0);
    }
  } else if constexpr (EPI == kEpiResidLN) {
    // the tile spans the whole row (n0 == 0, width == N)
    const bool ln = p.gamma != nullptr;
    const float s = valid ? row_scale(p.rs, m) : 1.0f;
    float sum = 0.f;
    const float* rrow = p.resid + (size_t)(valid ? m : 0) * p.ldr;
    constexpr bool TMA_RESID = (MODE & 4) != 0;
    const bool staged = TMA_RESID && (width & 31) == 0 && st.tm_resid != nullptr;
    if (TMA_RESID && staged) {
      // Residual tiles travel through the copy engine: box (c mod 3) is filled with columns [32c, 32c+32) of this warp's
      // 32 rows two chunks ahead of its use (the first two while the tile's MMAs are still running), every thread
      // updates its own 128-byte row in place, and the same box goes back out as the new residual stream.  No
      // per-thread global loads remain on the critical path (they were 30 % of this kernel's stall samples).
      const int nch = width >> 5;
      auto issue = [&](int c) {
        if (st.leader) {
          const int b = c % kResidBoxes;
          ptx::mbar_expect_tx(st.rbar + b, kStageBufBytes);
          ptx::tma_load_2d_addr(st.base + (uint32_t)b * kStageBufBytes, st.tm_resid, st.rbar + b, c * 32, m0);
        }
      };
      st.acquire();                    // every box is free again (stores of the previous tile have been read)
      issue(0);
      if (nch > 1) issue(1);
      wait_acc();
      for (int ch = 0; ch < nch; ++ch) {
        const int b = ch % kResidBoxes, c = ch * 32;
        const uint32_t row = st.base + (uint32_t)b * kStageBufBytes + (uint32_t)lane * 128u;
        float v[32];
        acc.template load<32>(c, v);
        ptx::mbar_wait(st.rbar + b, (st.rphase >> b) & 1u);
        st.rphase ^= 1u << b;
        uint4 rr[8];   // all eight loads first: the volatile accessors keep program order, one exposed latency instead of eight
#pragma unroll
        for (int j = 0; j < 8; ++j) rr[j] = ptx::ld_shared_v4(row + (uint32_t)((j ^ (lane & 7)) << 4));
        if (p.bias) add_vec<32>(p.bias + c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[4 * j] = fmaf(s, v[4 * j], __uint_as_float(rr[j].x));
          v[4 * j + 1] = fmaf(s, v[4 * j + 1], __uint_as_float(rr[j].y));
          v[4 * j + 2] = fmaf(s, v[4 * j + 2], __uint_as_float(rr[j].z));
          v[4 * j + 3] = fmaf(s, v[4 * j + 3], __uint_as_float(rr[j].w));
        }
        if (p.resid2 && valid) {
          float r2[32];
          load_f32_row<32>(p.resid2 + (size_t)m * p.ldr + c, r2);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += r2[i];
        }
        if (valid) {
#pragma unroll
          for (int i = 0; i < 32; ++i) sum += v[i];
        }
        if (ln) acc.template store<32>(c, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) ptx::st_shared_v4(row + (uint32_t)((j ^ (lane & 7)) << 4), pack4_f32(v + 4 * j));
        ptx::fence_proxy_async();
        __syncwarp();
        if (st.leader) {
          ptx::tma_store_2d(tmO0, st.base + (uint32_t)b * kStageBufBytes, c, m0);
          ptx::bulk_commit();
          // the box used one chunk ago is free once its store has been read: refill it for chunk ch + 2
          if (ch + 2 < nch) ptx::bulk_wait_read1();
        }
        st.pending = true;
        if (ch + 2 < nch) issue(ch + 2);
      }
    } else {
    // residual chunk of 32 columns, fetched one chunk ahead (the first fetch overlaps the MMA tail)
    float rnext[32];   // only live when PREFETCH
    if constexpr (PREFETCH) { if (width >= 32) load_f32_row<32>(rrow, rnext); }
    wait_acc();
    auto pass1 = [&](auto wtag, int c) {
      constexpr int W = decltype(wtag)::value;
      float v[W];
      acc.template load<W>(c, v);
      if (valid) {
        float r[W];
        if constexpr (W == 32 && PREFETCH) {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = rnext[i];
          if (c + 64 <= width) load_f32_row<32>(rrow + c + 32, rnext);
        } else {
          load_f32_row<W>(rrow + c, r);
        }
        if (p.bias) add_vec<W>(p.bias + c, v);
#pragma unroll
        for (int i = 0; i < W; ++i) v[i] = fmaf(s, v[i], r[i]);
        if (p.resid2) {
          load_f32_row<W>(p.resid2 + (size_t)m * p.ldr + c, r);
#pragma unroll
          for (int i = 0; i < W; ++i) v[i] += r[i];
        }
#pragma unroll
        for (int i = 0; i < W; ++i) sum += v[i];
      }
      if (ln) acc.template store<W>(c, v);
      const int b = st.begin_box();
#pragma unroll
      for (int i = 0; i < W / 4; ++i) st.put(b, i, pack4_f32(v + 4 * i));
      st.end_box(tmO0, c, m0);
    };
    int c = 0;
    for (; c + 32 <= width; c += 32) pass1(std::integral_constant<int, 32>{}, c);
    for (; c + 16 <= width; c += 16) pass1(std::integral_constant<int, 16>{}, c);
    }
    int c = 0;
    if (ln) {
      acc.fence_store();
      const float inv = 1.0f / (float)width;
      const float mean = sum * inv;
      float sq = 0.f;
      for (c = 0; c + 32 <= width; c += 32) sq += epi_sqdev_chunk<32>(acc, c, mean);
      for (; c + 16 <= width; c += 16) sq += epi_sqdev_chunk<16>(acc, c, mean);
      const float rstd = rsqrtf(sq * inv + p.ln_eps);
      auto pass3 = [&](auto wtag, int c) {
        constexpr int W = decltype(wtag)::value;
        float v[W];
        acc.template load<W>(c, v);
#pragma unroll
        for (int i = 0; i < W; i += 4) {
          const float4 g = __ldg(reinterpret_cast<const float4*>(p.gamma + c + i));
          const float4 be = __ldg(reinterpret_cast<const float4*>(p.beta + c + i));
          v[i] = fmaf((v[i] - mean) * rstd, g.x, be.x);
          v[i + 1] = fmaf((v[i + 1] - mean) * rstd, g.y, be.y);
          v[i + 2] = fmaf((v[i + 2] - mean) * rstd, g.z, be.z);
          v[i + 3] = fmaf((v[i + 3] - mean) * rstd, g.w, be.w);
        }
        const int b = (c & 63) == 0 ? st.begin_box() : 0;
        const int j0 = (c & 63) >> 3;
#pragma unroll
        for (int i = 0; i < W / 8; ++i) st.put(b, j0 + i, pack8_bf16(v + 8 * i));
        if (((c + W) & 63) == 0 || c + W >= width) st.end_box(tmO1, c & ~63, m0);
      };
      for (c = 0; c + 32 <= width; c += 32) pass3(std::integral_constant<int, 32>{}, c);
      for (; c + 16 <= width; c += 16) pass3(std::integral_constant<int, 16>{}, c);
      if (valid && p.stats) *reinterpret_cast<float2*>(p.stats + 2 * (size_t)m) = make_float2(mean, rstd);
    }
  }
}

// ---------------------------------------------------------------------------
// dgrad + LayerNorm backward (kEpiLnBwd).  The accumulator row of a thread is dy, the gradient w.r.t. the LayerNorm
// output; the tile spans the whole row.  Two passes over the accumulator, sixteen 32-column steps, three fp32 boxes per
// warp rotating through the copy engine two steps ahead of their use:
//   pass A (steps 0..7):  box = LayerNorm input x.  xhat = (x - mean) rstd, t = dy gamma; row sums s1 = sum t,
//                         s2 = sum t xhat; column sums of dy xhat and dy (gradients of gamma / beta) over the warp's 32
//                         rows, transposed through the (now dead) x box and the idle bf16 box (box_colsum), kept in
//                         registers across tiles; (t, xhat) go back into the accumulator packed as a bf16 pair (the
//                         separate kernel read dy as bf16 too).
//   pass B (steps 8..15): box = incoming residual gradient; dx = rstd (t - s1/N - xhat s2/N) + dx_in written in place
//                         and handed back to TMA as the outgoing residual gradient; bf16(rs dx) -- the next dgrad's
//                         operand -- through the fourth box, 64 columns per store.
// Reference: the autograd of nn.LayerNorm at /root/reference/Models.py:304-305 (norm1 / norm2 of Block.forward).
// ---------------------------------------------------------------------------
// Column sums over the warp's 32 rows of a [32 rows x 32 fp32] box (128B-swizzled rows: 16-byte piece j of row r sits at
// piece j ^ (r & 7)).  Lane l reads piece (l & 7) of the eight rows r = (l >> 3) + 4 i -- every quarter warp reads one whole
// 128-byte row, no bank conflicts -- two butterfly steps join the four row groups, and the lane keeps column
// 4 (l & 7) + (l >> 3) of the chunk (kLnBwdCol).  A shuffle-only transposed reduction of the register rows (31 shuffles
// and 62 selects per quantity) measured 875 cycles per call here; this one is eight loads deep.
__device__ __forceinline__ float box_colsum(uint32_t box, int lane) {
  const int p = lane & 7, g = lane >> 3;
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  uint4 v[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int r = g + 4 * i;
    v[i] = ptx::ld_shared_v4(box + (uint32_t)(r * 128 + ((p ^ (r & 7)) << 4)));
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a0 += __uint_as_float(v[i].x); a1 += __uint_as_float(v[i].y); a2 += __uint_as_float(v[i].z); a3 += __uint_as_float(v[i].w);
  }
#pragma unroll
  for (int h = 8; h <= 16; h <<= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, h); a1 += __shfl_xor_sync(0xffffffffu, a1, h);
    a2 += __shfl_xor_sync(0xffffffffu, a2, h); a3 += __shfl_xor_sync(0xffffffffu, a3, h);
  }
  return g == 0 ? a0 : (g == 1 ? a1 : (g == 2 ? a2 : a3));
}
__device__ __forceinline__ int kLnBwdCol(int lane) { return 4 * (lane & 7) + (lane >> 3); }   // column (within its chunk) a lane accumulates

struct LnBwdColAcc { float dg[8], db[8]; };   // column 32 * j + kLnBwdCol(lane)

template <class Wait>
__device__ __forceinline__ void tc_lnbwd_epilogue(const GemmArgs& p, TmemAcc& acc, Stager& st, const CUtensorMap* tmO0,
                                                  const CUtensorMap* tmO1, int m0, int lane, int width, uint32_t gamma_smem,
                                                  LnBwdColAcc& col, Wait wait_acc) {
  const int m = m0 + lane;
  const bool valid = m < p.M;
  const int nch = width >> 5;            // <= 8
  const int nsteps = 2 * nch;
  float mean = 0.f, rstd = 0.f, rsc = 1.f;
  if (valid) {
    const float2 ms = *reinterpret_cast<const float2*>(p.stats + 2 * (size_t)m);
    mean = ms.x; rstd = ms.y;
    rsc = row_scale(p.rs, m);
  }
  auto issue = [&](int k) {
    if (st.leader) {
      const int b = k % kResidBoxes;
      ptx::mbar_expect_tx(st.rbar + b, kStageBufBytes);
      if (k < nch) ptx::tma_load_2d_addr(st.base + (uint32_t)b * kStageBufBytes, st.tm_x, st.rbar + b, k * 32, m0);
      else ptx::tma_load_2d_addr(st.base + (uint32_t)b * kStageBufBytes, st.tm_resid, st.rbar + b, (k - nch) * 32, m0);
    }
  };
  st.acquire();                    // every box is free again (stores of the previous tile have been read)
  issue(0);
  issue(1);                        // nsteps >= 2
  wait_acc();
  float s1 = 0.f, s2 = 0.f;
  const uint32_t box_b = st.base + (uint32_t)kResidBoxes * kStageBufBytes + (uint32_t)lane * 128u;   // this lane's row of the fourth box
#pragma unroll
  for (int ch = 0; ch < 8; ++ch) {   // unrolled: the column accumulators are registers
    if (ch >= nch) break;
    const int b = ch % kResidBoxes, c = ch * 32;
    const uint32_t row = st.base + (uint32_t)b * kStageBufBytes + (uint32_t)lane * 128u;
    float dy[32], pg[32];
    acc.template load<32>(c, dy);
    ptx::mbar_wait(st.rbar + b, (st.rphase >> b) & 1u);
    st.rphase ^= 1u << b;
    uint4 xr[8], gr[8];            // all loads first: the volatile accessors keep program order
#pragma unroll
    for (int j = 0; j < 8; ++j) xr[j] = ptx::ld_shared_v4(row + (uint32_t)((j ^ (lane & 7)) << 4));
#pragma unroll
    for (int j = 0; j < 8; ++j) gr[j] = ptx::ld_shared_v4(gamma_smem + (uint32_t)((c + 4 * j) * 4));
    float wf[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float xs[4] = {__uint_as_float(xr[j].x), __uint_as_float(xr[j].y), __uint_as_float(xr[j].z), __uint_as_float(xr[j].w)};
      const float gs[4] = {__uint_as_float(gr[j].x), __uint_as_float(gr[j].y), __uint_as_float(gr[j].z), __uint_as_float(gr[j].w)};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = 4 * j + e;
        const float xh = (xs[e] - mean) * rstd;
        const float t = dy[i] * gs[e];
        s1 += t;
        s2 = fmaf(t, xh, s2);
        pg[i] = dy[i] * xh;
        wf[i] = __uint_as_float(pack_bf16x2(t, xh));
      }
    }
    acc.template store<32>(c, wf);
    // gradients of gamma / beta: dy * xhat goes into the x box (dead now), dy into the bf16 box (idle in this pass)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      ptx::st_shared_v4(row + (uint32_t)((j ^ (lane & 7)) << 4), pack4_f32(pg + 4 * j));
      ptx::st_shared_v4(box_b + (uint32_t)((j ^ (lane & 7)) << 4), pack4_f32(dy + 4 * j));
    }
    __syncwarp();
    col.dg[ch] += box_colsum(st.base + (uint32_t)b * kStageBufBytes, lane);
    col.db[ch] += box_colsum(st.base + (uint32_t)kResidBoxes * kStageBufBytes, lane);
    ptx::fence_proxy_async();      // the x box was written through the generic proxy; the copy engine refills it next
    __syncwarp();                  // every lane has read both boxes: they may be rewritten / refilled
    if (ch + 2 < nsteps) issue(ch + 2);   // box (ch + 2) % 3 was last read one chunk ago
  }
  acc.fence_store();
  const float inv = 1.0f / (float)width;
  const float a1 = s1 * inv, a2 = s2 * inv;
  const bool want_b = p.out1 != nullptr;
#pragma unroll 1
  for (int ch = 0; ch < nch; ++ch) {
    const int k = nch + ch;
    const int b = k % kResidBoxes, c = ch * 32;
    const uint32_t row = st.base + (uint32_t)b * kStageBufBytes + (uint32_t)lane * 128u;
    float wf[32], v[32];
    acc.template load<32>(c, wf);
    ptx::mbar_wait(st.rbar + b, (st.rphase >> b) & 1u);
    st.rphase ^= 1u << b;
    uint4 dr[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) dr[j] = ptx::ld_shared_v4(row + (uint32_t)((j ^ (lane & 7)) << 4));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float ds[4] = {__uint_as_float(dr[j].x), __uint_as_float(dr[j].y), __uint_as_float(dr[j].z), __uint_as_float(dr[j].w)};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int i = 4 * j + e;
        const float2 tx = unpack_bf16x2(__float_as_uint(wf[i]));   // (t, xhat)
        v[i] = fmaf(rstd, tx.x - a1 - tx.y * a2, ds[e]);
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) ptx::st_shared_v4(row + (uint32_t)((j ^ (lane & 7)) << 4), pack4_f32(v + 4 * j));
    ptx::fence_proxy_async();
    __syncwarp();
    if (st.leader) {
      ptx::tma_store_2d(tmO0, st.base + (uint32_t)b * kStageBufBytes, c, m0);
      ptx::bulk_commit();
      ptx::bulk_wait_read1();   // everything but the store just issued has been read: box (k + 2) % 3 and the bf16 box are free
    }
    __syncwarp();
    st.pending = true;
    if (k + 2 < nsteps) issue(k + 2);
    if (want_b) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= rsc;
      const int j0 = (ch & 1) * 4;
#pragma unroll
      for (int i = 0; i < 4; ++i) ptx::st_shared_v4(box_b + (uint32_t)(((j0 + i) ^ (lane & 7)) << 4), pack8_bf16(v + 8 * i));
      if ((ch & 1) || ch + 1 == nch) {
        ptx::fence_proxy_async();
        __syncwarp();
        if (st.leader) {
          ptx::tma_store_2d(tmO1, st.base + (uint32_t)kResidBoxes * kStageBufBytes, c & ~63, m0);
          ptx::bulk_commit();
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// descriptors
// ---------------------------------------------------------------------------
// Shared-memory matrix descriptor (sm_100 "version 1"), SWIZZLE_128B.
//   bits [0,14)  start address >> 4      bits [16,30) leading byte offset >> 4
//   bits [32,46) stride byte offset >> 4 bits [46,48) version = 1
//   bits [61,64) layout type (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16, bf16 x bf16 -> f32, M=128.
__host__ __device__ inline uint32_t make_idesc(int n, bool a_mn_major, bool b_mn_major, int m = 128) {
  uint32_t d = 0;
  d |= 1u << 4;                       // accumulator format f32
  d |= 1u << 7;                       // A = bf16
  d |= 1u << 10;                      // B = bf16
  d |= (a_mn_major ? 1u : 0u) << 15;  // A major
  d |= (b_mn_major ? 1u : 0u) << 16;  // B major
  d |= (uint32_t)(n >> 3) << 17;      // N / 8
  d |= (uint32_t)(m >> 4) << 24;      // M / 16 (256: both CTAs of a pair)
  return d;
}

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;             // 64 bf16 = one 128B swizzle row
constexpr int kATileBytes = kBlockM * kBlockK * 2;  // 16 KB
constexpr int kGemmThreads = 192;       // wgrad kernel: producer + MMA + 4 epilogue warps
constexpr int kSmemBudget = 200 * 1024;

// ---------------------------------------------------------------------------
// host side (defined in gemm_tc.cu): cached tensor maps and cluster launches
// ---------------------------------------------------------------------------
// 2-D tensor of bf16 (esz 2) or fp32 (esz 4), dim0 contiguous (d0 elements), d1 rows of `pitch` elements,
// box {b0, b1} with b0 * esz == 128 bytes (128B swizzle) or 64 bytes (64B swizzle).
int get_tmap(const void* ptr, uint64_t d0, uint64_t d1, uint64_t pitch, uint32_t b0, uint32_t b1, CUtensorMap* out, uint32_t esz = 2);

constexpr int kSmemMax = 227 * 1024;

// launches on `grid` CTAs, as clusters of P
template <class Kernel, class... Args>
int launch_clustered(Kernel kernel, int grid, int threads, size_t smem, int P, cudaStream_t stream, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = P; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  const bool pdl = pdl_enabled();
  cfg.attrs = attr; cfg.numAttrs = pdl ? 2 : 1;
  HS_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
  return kOk;
}

// grid for `units` work items of P CTAs each: at most one CTA per SM, whole clusters
inline int pair_grid(int units, int P) {
  int g = units * P;
  if (g > kNumSMs) g = kNumSMs / P * P;
  return g;
}


}  // namespace hsimae
