// The post-render epilogue of xvr's training iteration in ONE launch.
//
// /root/reference/src/xvr/model/trainer.py:292-302 (render_samples) follows every DRR batch with
//     mask = img > 0;  img = img.sum(dim=1, keepdim=True);
//     keep = mean(mask) > img_threshold                                   (no label channels)
//     keep = mean(any(mask[:, 1:], dim=1)) > mask_threshold               (label channels; channel 0 = background)
// and XrayTransforms (utils/preprocess.py:28-29) then needs the batch-global min and max of the kept images: ~20
// element-wise / reduction launches over the (B,C,H,W) batch.  Here one CTA per sample reads its C x N pixels once,
// writes the channel sum and reduces {foreground fraction, keep, min, max} with a fixed tree (deterministic).
#include "common.cuh"

namespace xvr {

__global__ void __launch_bounds__(1024)
render_epilogue_kernel(const float* __restrict__ img, int C, int N, float img_threshold, float mask_threshold,
                       float* __restrict__ sum_img, float* __restrict__ stats) {
  __shared__ float s_cnt[32], s_min[32], s_max[32];
  const int b = blockIdx.x;
  const float* x = img + (int64_t)b * C * N;
  float cnt = 0.f, mn = INFINITY, mx = -INFINITY;
  for (int n = threadIdx.x; n < N; n += 1024) {
    float s = 0.f;
    bool fg = false;
    for (int c = 0; c < C; ++c) {
      const float v = __ldg(x + (int64_t)c * N + n);
      s += v;
      if (C == 1 || c > 0) fg = fg || v > 0.f;
    }
    if (sum_img) sum_img[(int64_t)b * N + n] = s;
    cnt += fg ? 1.f : 0.f;
    mn = fminf(mn, s);
    mx = fmaxf(mx, s);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  }
  if ((threadIdx.x & 31) == 0) {
    s_cnt[threadIdx.x >> 5] = cnt;
    s_min[threadIdx.x >> 5] = mn;
    s_max[threadIdx.x >> 5] = mx;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    cnt = s_cnt[threadIdx.x];
    mn = s_min[threadIdx.x];
    mx = s_max[threadIdx.x];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if (threadIdx.x == 0) {
      const float frac = cnt / (float)N;  // counts are exact integers in fp32 (N < 2^24)
      stats[b * 4 + 0] = frac;
      stats[b * 4 + 1] = frac > (C == 1 ? img_threshold : mask_threshold) ? 1.f : 0.f;
      stats[b * 4 + 2] = mn;
      stats[b * 4 + 3] = mx;
    }
  }
}

}  // namespace xvr

using namespace xvr;

// img (B,C,N) -> sum_img (B,N) = channel sum (NULL allowed when C == 1: the image is its own sum),
// stats (B,4) = {foreground fraction, keep (0/1), min, max of the channel sum}.
extern "C" int xvr_render_epilogue(const float* img, int B, int C, int N, float img_threshold, float mask_threshold,
                                   float* sum_img, float* stats, void* stream) {
  if (!img || !stats || B <= 0 || C < 1 || N < 1 || N >= (1 << 24) || (C > 1 && !sum_img)) {
    set_last_error("xvr_render_epilogue: invalid argument");
    return XVR_ERR_INVALID;
  }
  render_epilogue_kernel<<<B, 1024, 0, (cudaStream_t)stream>>>(img, C, N, img_threshold, mask_threshold, sum_img, stats);
  return check_launch("xvr_render_epilogue");
}
