// Group-wise PCA preprocessing on the device.
//
// Reference: applyGWPCA, /root/reference/Utils/GroupWisePCA.py:20-34 -- flatten the [H, W, C] scene to [n, C] pixels, min-max
// normalise with the GLOBAL extrema (:23), split the bands into contiguous groups by repeated halving (split_data, :5-17),
// fit_transform an sklearn PCA(n_components = nc // group, whiten) on every group (:27-30) and concatenate (:32).
// PCA of a group = eigen-decomposition of its b x b covariance (b <= 64 bands): the device does the two passes over the
// pixels (moments, projection), the host only the tiny symmetric eigenproblems (hsimae_b200/gwpca.py).  The min-max
// scaling is affine, so everything is computed on the RAW pixels and the scale is folded into the projection weights.
//
// All arithmetic is fp64 (the reference works on float64 arrays); every reduction has a fixed order (per-CTA partials
// summed by a second kernel in CTA order), so results are run-to-run deterministic.  HBM-bound: X is read twice
// (moments) + once (projection); algorithmic bytes per pixel = 3 * C * sizeof(in) + nout * 8.
#include "kernels.cuh"
#include "../../include/hsimae_b200.h"

namespace hsimae {

namespace {

constexpr int kMaxGroups = 16;
constexpr int kGB = 64;                       // bands per group, padded
constexpr int kStatBlocks = 2 * kNumSMs;      // CTAs of the column pass
constexpr int kGramBlocks = kNumSMs / 2;      // CTAs per group of the covariance pass (x ngroups)
constexpr int kGramRows = 32;
constexpr int kProjRows = 16;
constexpr int kAbsBlocks = kNumSMs;

struct Groups {
  int n;
  int off[kMaxGroups + 1];
};

template <class T> __device__ __forceinline__ double ldx(const T* p) { return (double)__ldg(p); }

// ---- pass 1: column sums, global min / max ----------------------------------------------------------------------
// part_sum [kStatBlocks, c], part_mm [kStatBlocks, 2]
template <class T>
__global__ void __launch_bounds__(256)
colstats_kernel(const T* __restrict__ X, int64_t n, int c, double* __restrict__ part_sum, double* __restrict__ part_mm) {
  __shared__ double s_mn[256], s_mx[256];
  const int64_t rows_per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = rows_per * blockIdx.x, r1 = min(n, r0 + rows_per);
  double mn = INFINITY, mx = -INFINITY;
  for (int j = threadIdx.x; j < c; j += blockDim.x) {      // consecutive threads read consecutive bands of one pixel
    double s = 0.0;
    for (int64_t r = r0; r < r1; ++r) {
      const double v = ldx(X + r * c + j);
      s += v; mn = fmin(mn, v); mx = fmax(mx, v);
    }
    part_sum[(size_t)blockIdx.x * c + j] = s;
  }
  s_mn[threadIdx.x] = mn; s_mx[threadIdx.x] = mx;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_mn[threadIdx.x] = fmin(s_mn[threadIdx.x], s_mn[threadIdx.x + o]);
      s_mx[threadIdx.x] = fmax(s_mx[threadIdx.x], s_mx[threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part_mm[2 * blockIdx.x] = s_mn[0]; part_mm[2 * blockIdx.x + 1] = s_mx[0]; }
}

__global__ void colstats_finish_kernel(const double* __restrict__ part_sum, const double* __restrict__ part_mm, int nblk, int64_t n, int c,
                                       double* __restrict__ mean, double* __restrict__ minmax) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < c) {
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += part_sum[(size_t)b * c + j];
    mean[j] = s / (double)n;
  }
  if (j == 0) {
    double mn = INFINITY, mx = -INFINITY;
    for (int b = 0; b < nblk; ++b) { mn = fmin(mn, part_mm[2 * b]); mx = fmax(mx, part_mm[2 * b + 1]); }
    minmax[0] = mn; minmax[1] = mx;
  }
}

// ---- pass 2: centred Gram matrix of every band group ------------------------------------------------------------
// grid (kGramBlocks, ngroups); thread (tj, tk) owns the 4x4 block (4tj.., 4tk..) of the 64x64 matrix, upper triangle only.
// part [kGramBlocks, ngroups, 64*64]
template <class T>
__global__ void __launch_bounds__(256)
gram_kernel(const T* __restrict__ X, int64_t n, int c, Groups G, const double* __restrict__ mean, double* __restrict__ part) {
  __shared__ __align__(16) double tile[kGramRows][kGB];
  const int g = blockIdx.y, off = G.off[g], b = G.off[g + 1] - off;
  const int tj = threadIdx.x >> 4, tk = threadIdx.x & 15;
  const bool active = tk >= tj && 4 * tj < b && 4 * tk < b;
  double acc[4][4] = {};
  const int64_t ntiles = (n + kGramRows - 1) / kGramRows;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t r0 = t * kGramRows;
    __syncthreads();
    for (int i = threadIdx.x; i < kGramRows * kGB; i += blockDim.x) {
      const int r = i >> 6, j = i & 63;
      double v = 0.0;
      if (j < b && r0 + r < n) v = ldx(X + (r0 + r) * c + off + j) - mean[off + j];
      tile[r][j] = v;
    }
    __syncthreads();
    if (active) {
#pragma unroll 4
      for (int r = 0; r < kGramRows; ++r) {
        const double2 a01 = *reinterpret_cast<const double2*>(&tile[r][4 * tj]), a23 = *reinterpret_cast<const double2*>(&tile[r][4 * tj + 2]);
        const double2 b01 = *reinterpret_cast<const double2*>(&tile[r][4 * tk]), b23 = *reinterpret_cast<const double2*>(&tile[r][4 * tk + 2]);
        const double a[4] = {a01.x, a01.y, a23.x, a23.y}, bb[4] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
        for (int p = 0; p < 4; ++p)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[p][q] = fma(a[p], bb[q], acc[p][q]);
      }
    }
  }
  double* dst = part + ((size_t)blockIdx.x * gridDim.y + g) * (kGB * kGB);
  if (tk >= tj) {
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int q = 0; q < 4; ++q) dst[(4 * tj + p) * kGB + 4 * tk + q] = acc[p][q];
  }
}

// cov [ngroups, 64, 64] = sum over CTAs / (n - 1), mirrored to the full symmetric matrix
__global__ void gram_finish_kernel(const double* __restrict__ part, int nblk, int ngroups, int64_t n, double* __restrict__ cov) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= ngroups * kGB * kGB) return;
  const int g = e / (kGB * kGB), jk = e - g * kGB * kGB;
  int j = jk >> 6, k = jk & 63;
  if ((k >> 2) < (j >> 2)) { const int t = j; j = k; k = t; }
  double s = 0.0;
  for (int b = 0; b < nblk; ++b) s += part[((size_t)b * ngroups + g) * (kGB * kGB) + j * kGB + k];
  cov[e] = s / (double)(n > 1 ? n - 1 : 1);
}

// ---- projection: out[r, o] = sum_j (X[r, off_g + j] - mean[off_g + j]) * W[o, j],  g = o / k_per_group ----------------
template <class T>
__global__ void __launch_bounds__(256)
project_kernel(const T* __restrict__ X, int64_t n, int c, Groups G, int kper, const double* __restrict__ mean,
               const double* __restrict__ W, double* __restrict__ out) {
  extern __shared__ __align__(16) double smem[];
  const int nout = G.n * kper;
  double* sW = smem;                       // [nout][kGB + 1]
  double* sX = smem + nout * (kGB + 1);    // [kProjRows][c]
  for (int i = threadIdx.x; i < nout * kGB; i += blockDim.x) sW[(i >> 6) * (kGB + 1) + (i & 63)] = W[i];
  const int64_t ntiles = (n + kProjRows - 1) / kProjRows;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const int64_t r0 = t * kProjRows;
    const int rows = (int)min((int64_t)kProjRows, n - r0);
    __syncthreads();
    for (int i = threadIdx.x; i < rows * c; i += blockDim.x) sX[i] = ldx(X + r0 * c + i) - mean[i % c];
    __syncthreads();
    for (int i = threadIdx.x; i < rows * nout; i += blockDim.x) {
      const int r = i / nout, o = i - r * nout;
      const int g = o / kper, off = G.off[g], b = G.off[g + 1] - off;
      const double* x = sX + r * c + off;
      const double* w = sW + o * (kGB + 1);
      double s = 0.0;
      for (int j = 0; j < b; ++j) s = fma(x[j], w[j], s);
      out[(r0 + r) * nout + o] = s;
    }
  }
}

// ---- u-based sign convention: flip every column so that its entry of largest magnitude is positive -------------
// (sklearn.utils.extmath.svd_flip(u_based_decision=True): argmax |u[:, k]| over the samples, FIRST index on ties)
__global__ void __launch_bounds__(256)
absmax_kernel(const double* __restrict__ out, int64_t n, int nout, double* __restrict__ part_val, long long* __restrict__ part_idx) {
  // thread = (row lane, column): 256 / nout_pad rows in flight, columns fastest => coalesced
  __shared__ double s_v[256];
  __shared__ long long s_i[256];
  const int cols = nout;                  // <= 64
  const int lanes = 256 / cols;
  const int col = threadIdx.x % cols, lane = threadIdx.x / cols;
  const int64_t rows_per = (n + gridDim.x - 1) / gridDim.x;
  const int64_t r0 = rows_per * blockIdx.x, r1 = min(n, r0 + rows_per);
  double best = -1.0; long long bi = -1;
  if (lane < lanes) {
    for (int64_t r = r0 + lane; r < r1; r += lanes) {
      const double a = fabs(out[r * cols + col]);
      if (a > best) { best = a; bi = r; }       // rows ascend per thread: strict > keeps the first
    }
  }
  s_v[threadIdx.x] = best; s_i[threadIdx.x] = bi;
  __syncthreads();
  if (threadIdx.x < cols) {
    for (int l = 1; l < lanes; ++l) {
      const double v = s_v[l * cols + col]; const long long i = s_i[l * cols + col];
      if (i >= 0 && (v > best || (v == best && i < bi) || bi < 0)) { best = v; bi = i; }
    }
    part_val[blockIdx.x * cols + col] = best;
    part_idx[blockIdx.x * cols + col] = bi;
  }
}

__global__ void absmax_finish_kernel(const double* __restrict__ out, const double* __restrict__ part_val, const long long* __restrict__ part_idx,
                                     int nblk, int nout, double* __restrict__ signs) {
  const int col = threadIdx.x;
  if (col >= nout) return;
  double best = -1.0; long long bi = -1;
  for (int b = 0; b < nblk; ++b) {          // CTAs own ascending row ranges: strict > keeps the first index
    const double v = part_val[b * nout + col]; const long long i = part_idx[b * nout + col];
    if (i >= 0 && v > best) { best = v; bi = i; }
  }
  double s = 1.0;
  if (bi >= 0) { const double v = out[bi * nout + col]; s = v < 0.0 ? -1.0 : 1.0; }   // np.sign(0) = 0 cannot occur for a max |.| > 0
  signs[col] = s;
}

__global__ void flip_kernel(double* __restrict__ out, int64_t total, int nout, const double* __restrict__ signs) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] *= signs[i % nout];
}

int fill_groups(Groups& G, int32_t c, int32_t ngroups, const int32_t* group_off) {
  HS_REQUIRE(group_off != nullptr, "gwpca: group_off is null");
  HS_REQUIRE(ngroups >= 1 && ngroups <= kMaxGroups, "gwpca: %d band groups (supported: 1..%d)", ngroups, kMaxGroups);
  G.n = ngroups;
  for (int i = 0; i <= ngroups; ++i) G.off[i] = group_off[i];
  HS_REQUIRE(G.off[0] >= 0 && G.off[ngroups] <= c, "gwpca: band groups exceed the %d bands", c);
  for (int i = 0; i < ngroups; ++i)
    HS_REQUIRE(G.off[i + 1] - G.off[i] >= 1 && G.off[i + 1] - G.off[i] <= kGB, "gwpca: group %d has %d bands (supported: 1..%d)", i,
               G.off[i + 1] - G.off[i], kGB);
  return kOk;
}

size_t moments_ws_bytes(int c, int ngroups) {
  return ((size_t)kStatBlocks * c + 2 * kStatBlocks + (size_t)kGramBlocks * ngroups * kGB * kGB) * sizeof(double);
}

}  // namespace

}  // namespace hsimae

extern "C" int64_t hsimae_gwpca_workspace_bytes(int32_t c, int32_t ngroups, int32_t nout) {
  using namespace hsimae;
  if (c < 1 || ngroups < 1 || ngroups > kMaxGroups || nout < 0 || nout > 64) return -1;
  const size_t flip = (size_t)kAbsBlocks * 64 * (sizeof(double) + sizeof(long long)) + 64 * sizeof(double);
  const size_t m = moments_ws_bytes(c, ngroups);
  return (int64_t)(m > flip ? m : flip);
}

extern "C" int hsimae_gwpca_moments(const void* X, int32_t dtype, int64_t n, int32_t c, int32_t ngroups, const int32_t* group_off,
                                    void* ws, int64_t ws_bytes, double* mean, double* minmax, double* cov, void* stream) {
  using namespace hsimae;
  Groups G;
  HS_TRY(fill_groups(G, c, ngroups, group_off));
  HS_REQUIRE(X && ws && mean && minmax && cov, "gwpca_moments: null argument");
  HS_REQUIRE(dtype == 0 || dtype == 1, "gwpca_moments: dtype %d (0 = float32, 1 = float64)", dtype);
  HS_REQUIRE(n >= 2 && c >= 1, "gwpca_moments: need at least 2 pixels and 1 band (n=%lld, c=%d)", (long long)n, c);
  HS_REQUIRE(ws_bytes >= (int64_t)moments_ws_bytes(c, ngroups), "gwpca_moments: workspace of %lld bytes is too small", (long long)ws_bytes);
  cudaStream_t st = (cudaStream_t)stream;
  double* part_sum = (double*)ws;
  double* part_mm = part_sum + (size_t)kStatBlocks * c;
  double* part_gram = part_mm + 2 * kStatBlocks;
  if (dtype == 0) colstats_kernel<float><<<kStatBlocks, 256, 0, st>>>((const float*)X, n, c, part_sum, part_mm);
  else colstats_kernel<double><<<kStatBlocks, 256, 0, st>>>((const double*)X, n, c, part_sum, part_mm);
  HS_CHECK_LAUNCH("gwpca colstats_kernel");
  colstats_finish_kernel<<<ceil_div(c, 128), 128, 0, st>>>(part_sum, part_mm, kStatBlocks, n, c, mean, minmax);
  HS_CHECK_LAUNCH("gwpca colstats_finish_kernel");
  dim3 grid(kGramBlocks, ngroups);
  if (dtype == 0) gram_kernel<float><<<grid, 256, 0, st>>>((const float*)X, n, c, G, mean, part_gram);
  else gram_kernel<double><<<grid, 256, 0, st>>>((const double*)X, n, c, G, mean, part_gram);
  HS_CHECK_LAUNCH("gwpca gram_kernel");
  gram_finish_kernel<<<ceil_div(ngroups * kGB * kGB, 256), 256, 0, st>>>(part_gram, kGramBlocks, ngroups, n, cov);
  HS_CHECK_LAUNCH("gwpca gram_finish_kernel");
  return kOk;
}

extern "C" int hsimae_gwpca_project(const void* X, int32_t dtype, int64_t n, int32_t c, int32_t ngroups, const int32_t* group_off,
                                    int32_t k_per_group, const double* mean, const double* W, double* out, int32_t flip_u, void* ws,
                                    int64_t ws_bytes, void* stream) {
  using namespace hsimae;
  Groups G;
  HS_TRY(fill_groups(G, c, ngroups, group_off));
  if (n == 0) return kOk;
  HS_REQUIRE(X && mean && W && out, "gwpca_project: null argument");
  HS_REQUIRE(dtype == 0 || dtype == 1, "gwpca_project: dtype %d (0 = float32, 1 = float64)", dtype);
  const int nout = ngroups * k_per_group;
  HS_REQUIRE(k_per_group >= 1 && nout <= 64, "gwpca_project: %d groups x %d components (supported: up to 64 output channels)", ngroups, k_per_group);
  const size_t smem = ((size_t)nout * (kGB + 1) + (size_t)kProjRows * c) * sizeof(double);
  HS_REQUIRE(smem <= 200 * 1024, "gwpca_project: %d bands do not fit the shared-memory row tile", c);
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ntiles = (n + kProjRows - 1) / kProjRows;
  const int grid = (int)(ntiles < 4 * kNumSMs ? ntiles : 4 * kNumSMs);
  if (dtype == 0) {
    if (smem > 48 * 1024) HS_CHECK_CUDA(cudaFuncSetAttribute(project_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    project_kernel<float><<<grid, 256, smem, st>>>((const float*)X, n, c, G, k_per_group, mean, W, out);
  } else {
    if (smem > 48 * 1024) HS_CHECK_CUDA(cudaFuncSetAttribute(project_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    project_kernel<double><<<grid, 256, smem, st>>>((const double*)X, n, c, G, k_per_group, mean, W, out);
  }
  HS_CHECK_LAUNCH("gwpca project_kernel");
  if (flip_u) {
    const size_t need = (size_t)kAbsBlocks * nout * (sizeof(double) + sizeof(long long)) + nout * sizeof(double);
    HS_REQUIRE(ws && ws_bytes >= (int64_t)need, "gwpca_project: workspace of %lld bytes is too small for the sign pass", (long long)ws_bytes);
    double* part_val = (double*)ws;
    long long* part_idx = (long long*)(part_val + (size_t)kAbsBlocks * nout);
    double* signs = (double*)(part_idx + (size_t)kAbsBlocks * nout);
    absmax_kernel<<<kAbsBlocks, 256, 0, st>>>(out, n, nout, part_val, part_idx);
    HS_CHECK_LAUNCH("gwpca absmax_kernel");
    absmax_finish_kernel<<<1, 64, 0, st>>>(out, part_val, part_idx, kAbsBlocks, nout, signs);
    HS_CHECK_LAUNCH("gwpca absmax_finish_kernel");
    flip_kernel<<<4 * kNumSMs, 256, 0, st>>>(out, n * nout, nout, signs);
    HS_CHECK_LAUNCH("gwpca flip_kernel");
  }
  return kOk;
}
