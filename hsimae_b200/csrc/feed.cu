// On-device pretraining data feed.
//
// Reference: HSIdataset4PT.__getitem__, /root/reference/Model_Pretraining.py:40-51 -- per sample a Python slice
// `cube[h:h+9, w:w+9, :]` of scene `num`, `(x - min) / (max - min)`, optional horizontal (np.flip(data, 1): the W axis)
// and vertical (np.flip(data, 0): the H axis) flips, HWC -> [1, C, H, W]; batches are assembled by a DataLoader with
// num_workers = 0.  Here the scenes stay resident in HBM and ONE launch assembles the whole batch: one CTA per
// sample reads the 9 contiguous pixel rows of the window (channels fastest, 128-byte lines), transposes through
// shared memory and writes the [C, H, W] cube with contiguous stores.  The flips are index reversals on the read
// side; which samples flip is decided on the host with the reference's RNG calls (hsimae_b200/feed.py).
// Pure data movement + one IEEE subtract/divide per element: results are bit-identical to the reference for fp32 scenes.
#include "kernels.cuh"
#include "../../include/hsimae_b200.h"

namespace hsimae {

namespace {

constexpr int kFeedThreads = 256;

__global__ void __launch_bounds__(kFeedThreads)
gather_patches_kernel(const float* __restrict__ scenes, const int64_t* __restrict__ scene_off, const int32_t* __restrict__ scene_hw,
                      int bands, int img, const int16_t* __restrict__ cut_info, const int64_t* __restrict__ index,
                      const uint8_t* __restrict__ flips, int B, float* __restrict__ out) {
  extern __shared__ float tile[];   // [img*img][bands + 1]
  const int px = img * img, pitch = bands + 1;
  for (int n = blockIdx.x; n < B; n += gridDim.x) {
    const int16_t* ci = cut_info + (size_t)index[n] * 6;      // (c, h, w, scene, max, min), Utils/Preprocessing.py:78,114
    const int h0 = ci[1], w0 = ci[2], sc = ci[3];
    const float mx = (float)ci[4], mn = (float)ci[5];
    const float range = mx - mn;
    const bool hflip = flips && flips[2 * n], vflip = flips && flips[2 * n + 1];
    const int W = scene_hw[2 * sc + 1];
    const float* src = scenes + scene_off[sc];
    __syncthreads();
    // element (y, x, c) of the window: consecutive threads walk the channels of one pixel, then the next pixel of the row
    for (int i = threadIdx.x; i < px * bands; i += blockDim.x) {
      const int p = i / bands, c = i - p * bands;
      const int y = p / img, x = p - y * img;
      const float v = __ldg(src + ((size_t)(h0 + y) * W + (w0 + x)) * bands + c);
      tile[p * pitch + c] = (v - mn) / range;
    }
    __syncthreads();
    float* dst = out + (size_t)n * bands * px;
    for (int i = threadIdx.x; i < px * bands; i += blockDim.x) {
      const int c = i / px, p = i - c * px;
      const int y = p / img, x = p - y * img;
      const int ys = vflip ? img - 1 - y : y, xs = hflip ? img - 1 - x : x;
      dst[i] = tile[(ys * img + xs) * pitch + c];
    }
  }
}

}  // namespace

}  // namespace hsimae

extern "C" int hsimae_gather_patches(const float* scenes, const int64_t* scene_off, const int32_t* scene_hw, int32_t bands, int32_t img,
                                     const int16_t* cut_info, const int64_t* index, const uint8_t* flips, int32_t n, float* out,
                                     void* stream) {
  using namespace hsimae;
  if (n == 0) return kOk;
  HS_REQUIRE(scenes && scene_off && scene_hw && cut_info && index && out, "gather_patches: null argument");
  HS_REQUIRE(bands > 0 && img > 0 && n > 0, "gather_patches: bad shape bands=%d img=%d n=%d", bands, img, n);
  const size_t smem = (size_t)img * img * (bands + 1) * sizeof(float);
  HS_REQUIRE(smem <= 200 * 1024, "gather_patches: window of %d x %d x %d floats does not fit in shared memory", img, img, bands);
  if (smem > 48 * 1024) HS_CHECK_CUDA(cudaFuncSetAttribute(gather_patches_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = n < 16 * kNumSMs ? n : 16 * kNumSMs;
  gather_patches_kernel<<<grid, kFeedThreads, smem, (cudaStream_t)stream>>>(scenes, scene_off, scene_hw, bands, img, cut_info, index, flips, n, out);
  HS_CHECK_LAUNCH("gather_patches_kernel");
  return kOk;
}
