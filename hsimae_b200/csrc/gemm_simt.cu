// CUDA-core checker GEMMs.  NOT a product path: they exist so that the tcgen05
// kernels in gemm_tc.cu can be validated on the GPU against an independent,
// obviously-correct implementation that runs the very same epilogues
// (epilogue.cuh) -- see tests/test_gemm_gpu.py.  They are only reachable
// through hsimae_gemm_check / hsimae_wgrad_check and the HSIMAE_DEBUG_SIMT=1
// debugging switch of the engine.
#include "epilogue.cuh"

namespace hsimae {

// C_scratch[M,N] (fp32) = A[M,K] * B[N,K]^T, 64x64 tile per CTA, 16x16 threads, 4x4 micro-tile.
__global__ void __launch_bounds__(256)
simt_gemm_kernel(const __nv_bfloat16* __restrict__ A, int lda, const __nv_bfloat16* __restrict__ B, int ldb,
                 float* __restrict__ C, int ldc, int M, int N, int K) {
  __shared__ float sa[16][65];
  __shared__ float sb[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int i = threadIdx.x; i < 64 * 16; i += 256) {
      int r = i >> 4, c = i & 15;
      int m = m0 + r, n = n0 + r, k = k0 + c;
      sa[c][r] = (m < M && k < K) ? __bfloat162float(A[(size_t)m * lda + k]) : 0.f;
      sb[c][r] = (n < N && k < K) ? __bfloat162float(B[(size_t)n * ldb + k]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sa[k][ty * 4 + i]; b[i] = sb[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int m = m0 + ty * 4 + i, n = n0 + tx * 4 + j;
      if (m < M && n < N) C[(size_t)m * ldc + n] = acc[i][j];
    }
}

template <int EPI>
__global__ void __launch_bounds__(128)
simt_epilogue_kernel(GemmArgs p, float* scratch, int ldc) {
  const int m = blockIdx.x * 128 + threadIdx.x;
  if (m >= p.M) return;  // the global-scratch accumulator needs no warp-collective access
  GmemAcc acc{scratch + (size_t)m * ldc};
  run_epilogue<EPI>(p, acc, m, 0, p.N);
}

int gemm_check_args(const GemmArgs& a, int epi);

int gemm_simt(const GemmArgs& a, int epi, float* scratch, cudaStream_t stream) {
  HS_TRY(gemm_check_args(a, epi));
  HS_REQUIRE(scratch != nullptr, "gemm_simt needs an fp32 scratch of M*N elements");
  dim3 grid(ceil_div(a.N, 64), ceil_div(a.M, 64));
  simt_gemm_kernel<<<grid, 256, 0, stream>>>(a.A, a.lda, a.B, a.ldb, scratch, a.N, a.M, a.N, a.K);
  HS_CHECK_LAUNCH("simt_gemm_kernel");
  const int eg = ceil_div(a.M, 128);
  switch (epi) {
    case kEpiBiasBf16: simt_epilogue_kernel<kEpiBiasBf16><<<eg, 128, 0, stream>>>(a, scratch, a.N); break;
    case kEpiBiasF32:  simt_epilogue_kernel<kEpiBiasF32><<<eg, 128, 0, stream>>>(a, scratch, a.N); break;
    case kEpiResidLN:  simt_epilogue_kernel<kEpiResidLN><<<eg, 128, 0, stream>>>(a, scratch, a.N); break;
    case kEpiSwiGLU:   simt_epilogue_kernel<kEpiSwiGLU><<<eg, 128, 0, stream>>>(a, scratch, a.N); break;
    case kEpiDSwiGLU:  simt_epilogue_kernel<kEpiDSwiGLU><<<eg, 128, 0, stream>>>(a, scratch, a.N); break;
  }
  HS_CHECK_LAUNCH("simt_epilogue_kernel");
  return kOk;
}

// W[map(n), k] += sum_m Y[m,n] X[m,k]; one 32x32 output tile per CTA column, reduction split over blockIdx.z.
__global__ void __launch_bounds__(256)
simt_wgrad_kernel(WgradArgs p, int rows_per_split) {
  __shared__ float sy[32][33];
  __shared__ float sx[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  const int n0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int m0 = blockIdx.z * rows_per_split;
  int m1 = m0 + rows_per_split; if (m1 > p.Mred) m1 = p.Mred;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  float bsum = 0.f;
  for (int mb = m0; mb < m1; mb += 32) {
    for (int i = threadIdx.x; i < 32 * 32; i += 256) {
      int r = i >> 5, c = i & 31;
      int m = mb + r;
      sy[r][c] = (m < m1 && n0 + c < p.Nout) ? __bfloat162float(p.Y[(size_t)m * p.ldy + n0 + c]) : 0.f;
      sx[r][c] = (m < m1 && k0 + c < p.Kin) ? __bfloat162float(p.X[(size_t)m * p.ldx + k0 + c]) : 0.f;
    }
    __syncthreads();
    for (int r = 0; r < 32; ++r) {
      float x = sx[r][tx];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] = fmaf(sy[r][ty * 4 + i], x, acc[i]);
    }
    if (blockIdx.x == 0 && ty == 0)
      for (int r = 0; r < 32; ++r) bsum += sy[r][tx];
    __syncthreads();
  }
  auto map_row = [&](int r, float* d0, float* d1, int stride) -> float* {
    if (r >= p.Nout) return nullptr;
    if (p.row_map == 0) return r < p.rows_valid ? d0 + (size_t)r * stride : nullptr;
    const int which = (r % (2 * kGate)) / kGate;
    const int h = (r / (2 * kGate)) * kGate + (r % kGate);
    return h < p.rows_valid ? (which ? d1 : d0) + (size_t)h * stride : nullptr;
  };
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float* row = map_row(n0 + ty * 4 + i, p.dst0, p.dst1, p.ld);
    int k = k0 + tx;
    if (row && k < p.cols_valid) atomicAdd(row + k, acc[i]);
  }
  if (p.bias0 && blockIdx.x == 0 && ty == 0) {
    float* b = map_row(n0 + tx, p.bias0, p.bias1, 1);
    if (b) atomicAdd(b, bsum);
  }
}

int wgrad_check_args(const WgradArgs& a);

int wgrad_simt(const WgradArgs& a, cudaStream_t stream) {
  HS_TRY(wgrad_check_args(a));
  int gx = ceil_div(a.Kin, 32), gy = ceil_div(a.Nout, 32);
  int gz = ceil_div(4 * kNumSMs, gx * gy);
  int rows = ceil_div(a.Mred, gz);
  rows = ceil_div(rows, 32) * 32;
  gz = ceil_div(a.Mred, rows);
  simt_wgrad_kernel<<<dim3(gx, gy, gz), 256, 0, stream>>>(a, rows);
  HS_CHECK_LAUNCH("simt_wgrad_kernel");
  return kOk;
}

}  // namespace hsimae
