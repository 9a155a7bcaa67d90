// Image-similarity kernels: global and sliding-patch normalised cross correlation (forward + backward) and the
// 3x3 Sobel operator (forward + transpose).
//
// Replace DiffDRR 0.6.0 metrics.NormalizedCrossCorrelation2d / MultiscaleNormalizedCrossCorrelation2d /
// GradientNormalizedCrossCorrelation2d / Sobel as called at /root/reference/src/xvr/model/loss.py:16,27 and
// /root/reference/src/xvr/registrar/base.py:119-122.  The reference unfolds every p x p window into a
// (B, (H-p+1)(W-p+1), p, p) tensor (an 81x / 121x blow-up) and z-scores it; here a thread owns a window and
// reads it from a shared-memory tile.  The backward pass is a gather: d ncc / d x_i over the windows containing
// pixel i is  sum_w A_w (x1_i - mu1_w) - B_w (x2_i - mu2_w)  with four per-window coefficients saved by the
// forward pass, i.e. a p x p gather from a shared-memory tile -- no atomics, deterministic.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace xvr {

constexpr int NCC_T = 16;  // windows per CTA edge

// per-window statistics -> ncc and the gradient coefficients w.r.t. the second / first image
struct WindowStats {
  float ncc, A, B2, B1, mu1, mu2;
};

__device__ __forceinline__ WindowStats window_stats(float s1, float s2, float s11, float s22, float s12, float c1,
                                                    float c2, float inv_n, float eps) {
  // sums are of pivot-shifted values (x - c): variances are shift invariant, means get the pivot back
  const float m1 = s1 * inv_n, m2 = s2 * inv_n;
  const float v1 = fmaxf(s11 * inv_n - m1 * m1, 0.f) + eps;
  const float v2 = fmaxf(s22 * inv_n - m2 * m2, 0.f) + eps;
  const float cov = s12 * inv_n - m1 * m2;
  const float r1 = rsqrtf(v1), r2 = rsqrtf(v2);
  WindowStats w;
  w.A = r1 * r2;
  w.ncc = cov * w.A;
  w.mu1 = m1 + c1;
  w.mu2 = m2 + c2;
  w.B2 = w.ncc / v2;
  w.B1 = w.ncc / v1;
  return w;
}

// ---------------------------------------------------------------------------------------------- patch NCC
// grid (tiles_x, tiles_y, B*C), block NCC_T x NCC_T.  partial[(bc, tile)] = sum of window nccs of the tile.
// coef2 / coef1 (nullable): (B*C, 4, nH, nW) coefficient maps for the gradient w.r.t. x2 / x1.
__global__ void __launch_bounds__(NCC_T* NCC_T)
patch_ncc_fwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int H, int W, int p, float eps,
                     float* __restrict__ partial, float* __restrict__ coef2, float* __restrict__ coef1) {
  extern __shared__ float tile[];
  const int nH = H - p + 1, nW = W - p + 1;
  const int TS = NCC_T + p - 1;
  float* t1 = tile;
  float* t2 = tile + TS * TS;
  const int bc = blockIdx.z;
  const int oy = blockIdx.y * NCC_T, ox = blockIdx.x * NCC_T;
  const float* a = x1 + (int64_t)bc * H * W;
  const float* b = x2 + (int64_t)bc * H * W;
  const int tid = threadIdx.y * NCC_T + threadIdx.x;
  for (int k = tid; k < TS * TS; k += NCC_T * NCC_T) {
    const int r = k / TS, c = k - r * TS;
    const int y = oy + r, x = ox + c;
    const bool in = y < H && x < W;
    t1[k] = in ? __ldg(a + y * W + x) : 0.f;
    t2[k] = in ? __ldg(b + y * W + x) : 0.f;
  }
  __syncthreads();
  const int wy = oy + threadIdx.y, wx = ox + threadIdx.x;
  float ncc = 0.f;
  if (wy < nH && wx < nW) {
    const int base = threadIdx.y * TS + threadIdx.x;
    const int ctr = base + (p / 2) * TS + p / 2;
    const float c1 = t1[ctr], c2 = t2[ctr];
    float s1 = 0.f, s2 = 0.f, s11 = 0.f, s22 = 0.f, s12 = 0.f;
    for (int r = 0; r < p; ++r) {
      for (int c = 0; c < p; ++c) {
        const float u = t1[base + r * TS + c] - c1, v = t2[base + r * TS + c] - c2;
        s1 += u;
        s2 += v;
        s11 = fmaf(u, u, s11);
        s22 = fmaf(v, v, s22);
        s12 = fmaf(u, v, s12);
      }
    }
    const WindowStats w = window_stats(s1, s2, s11, s22, s12, c1, c2, 1.0f / (float)(p * p), eps);
    ncc = w.ncc;
    const int64_t plane = (int64_t)nH * nW;
    const int64_t o = (int64_t)bc * 4 * plane + (int64_t)wy * nW + wx;
    if (coef2) {  // gradient w.r.t. x2: A (x1 - mu1) - B2 (x2 - mu2)
      coef2[o] = w.A;
      coef2[o + plane] = w.B2;
      coef2[o + 2 * plane] = w.mu1;
      coef2[o + 3 * plane] = w.mu2;
    }
    if (coef1) {  // gradient w.r.t. x1: A (x2 - mu2) - B1 (x1 - mu1)
      coef1[o] = w.A;
      coef1[o + plane] = w.B1;
      coef1[o + 2 * plane] = w.mu2;
      coef1[o + 3 * plane] = w.mu1;
    }
  }
  // deterministic block sum
  __shared__ float red[NCC_T * NCC_T / 32];
  ncc = warp_sum(ncc);
  if ((tid & 31) == 0) red[tid >> 5] = ncc;
  __syncthreads();
  if (tid < 32) {
    float v = tid < NCC_T * NCC_T / 32 ? red[tid] : 0.f;
    v = warp_sum(v);
    if (tid == 0) partial[(int64_t)bc * gridDim.x * gridDim.y + blockIdx.y * gridDim.x + blockIdx.x] = v;
  }
}

// grad[b,c,y,x] = gscale[b] * sum_w ( A_w (xa - mua_w) - B_w (xb - mub_w) ), windows w containing (y,x).
// For the gradient w.r.t. x2: xa = x1, xb = x2; w.r.t. x1: xa = x2, xb = x1.  (The differences are formed per
// window: factoring xa * sum(A_w) - sum(A_w mua_w) cancels catastrophically on flat windows where A_w ~ 1/eps.)
__global__ void __launch_bounds__(NCC_T* NCC_T)
patch_ncc_bwd_kernel(const float* __restrict__ xa, const float* __restrict__ xb, const float* __restrict__ coef,
                     const float* __restrict__ gscore, float scale, int C, int H, int W, int p,
                     float* __restrict__ grad, int accumulate) {
  extern __shared__ float tile[];
  const int nH = H - p + 1, nW = W - p + 1;
  const int TS = NCC_T + p - 1;
  const int bc = blockIdx.z;
  const int oy = blockIdx.y * NCC_T, ox = blockIdx.x * NCC_T;
  const int64_t plane = (int64_t)nH * nW;
  const float* cf = coef + (int64_t)bc * 4 * plane;
  const int tid = threadIdx.y * NCC_T + threadIdx.x;
  // window (wy, wx) covers pixels wy..wy+p-1: pixel y needs windows y-p+1..y -> tile origin oy-p+1
  for (int k = tid; k < TS * TS; k += NCC_T * NCC_T) {
    const int r = k / TS, c = k - r * TS;
    const int wy = oy - (p - 1) + r, wx = ox - (p - 1) + c;
    const bool in = wy >= 0 && wy < nH && wx >= 0 && wx < nW;
    const int64_t o = (int64_t)wy * nW + wx;
    tile[k] = in ? __ldg(cf + o) : 0.f;
    tile[TS * TS + k] = in ? __ldg(cf + plane + o) : 0.f;
    tile[2 * TS * TS + k] = in ? __ldg(cf + 2 * plane + o) : 0.f;
    tile[3 * TS * TS + k] = in ? __ldg(cf + 3 * plane + o) : 0.f;
  }
  __syncthreads();
  const int y = oy + threadIdx.y, x = ox + threadIdx.x;
  if (y >= H || x >= W) return;
  const int64_t i = (int64_t)bc * H * W + (int64_t)y * W + x;
  const float va = __ldg(xa + i), vb = __ldg(xb + i);
  float acc = 0.f;
  const int base = threadIdx.y * TS + threadIdx.x;
  for (int r = 0; r < p; ++r) {
    for (int c = 0; c < p; ++c) {
      const int k = base + r * TS + c;
      acc = fmaf(tile[k], va - tile[2 * TS * TS + k], acc);
      acc = fmaf(-tile[TS * TS + k], vb - tile[3 * TS * TS + k], acc);
    }
  }
  const float g = __ldg(gscore + bc / C) * scale;
  const float v = g * acc;
  grad[i] = accumulate ? grad[i] + v : v;
}

// ---------------------------------------------------------------------------------------------- global NCC
// One CTA per (b,c): two passes over the image (mean, then centred moments).  stats[(bc)] = {ncc, A, B2, B1, mu1, mu2}.
__global__ void __launch_bounds__(1024)
global_ncc_fwd_kernel(const float* __restrict__ x1, const float* __restrict__ x2, int n, float eps,
                      float* __restrict__ stats) {
  __shared__ float red[32][5];
  __shared__ float mean[2];
  const float* a = x1 + (int64_t)blockIdx.x * n;
  const float* b = x2 + (int64_t)blockIdx.x * n;
  const int tid = threadIdx.x;
  float s1 = 0.f, s2 = 0.f;
  for (int i = tid; i < n; i += 1024) {
    s1 += __ldg(a + i);
    s2 += __ldg(b + i);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((tid & 31) == 0) {
    red[tid >> 5][0] = s1;
    red[tid >> 5][1] = s2;
  }
  __syncthreads();
  if (tid < 32) {
    const float u = warp_sum(red[tid][0]), v = warp_sum(red[tid][1]);
    if (tid == 0) {
      mean[0] = u / (float)n;
      mean[1] = v / (float)n;
    }
  }
  __syncthreads();
  const float c1 = mean[0], c2 = mean[1];
  float q[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
  for (int i = tid; i < n; i += 1024) {
    const float u = __ldg(a + i) - c1, v = __ldg(b + i) - c2;
    q[0] += u;
    q[1] += v;
    q[2] = fmaf(u, u, q[2]);
    q[3] = fmaf(v, v, q[3]);
    q[4] = fmaf(u, v, q[4]);
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const float v = warp_sum(q[k]);
    if ((tid & 31) == 0) red[tid >> 5][k] = v;
  }
  __syncthreads();
  if (tid < 32) {
    float r[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) r[k] = warp_sum(red[tid][k]);
    if (tid == 0) {
      const WindowStats w = window_stats(r[0], r[1], r[2], r[3], r[4], c1, c2, 1.0f / (float)n, eps);
      float* o = stats + (int64_t)blockIdx.x * 6;
      o[0] = w.ncc; o[1] = w.A; o[2] = w.B2; o[3] = w.B1; o[4] = w.mu1; o[5] = w.mu2;
    }
  }
}

// grad[bc, i] (+)= gscore[b] * scale * (A (xa_i - mua) - B (xb_i - mub))
__global__ void __launch_bounds__(256)
global_ncc_bwd_kernel(const float* __restrict__ xa, const float* __restrict__ xb, const float* __restrict__ stats,
                      int which, const float* __restrict__ gscore, float scale, int C, int n,
                      float* __restrict__ grad, int accumulate) {
  const int bc = blockIdx.y;
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float* s = stats + (int64_t)bc * 6;
  const float A = s[1], B = which == 2 ? s[2] : s[3];
  const float mua = which == 2 ? s[4] : s[5], mub = which == 2 ? s[5] : s[4];
  const int64_t o = (int64_t)bc * n + i;
  const float v = __ldg(gscore + bc / C) * scale * (A * (__ldg(xa + o) - mua) - B * (__ldg(xb + o) - mub));
  grad[o] = accumulate ? grad[o] + v : v;
}

// score[b] = scale * sum over the C*T partials of image b (fixed tree); accumulate adds into score.
__global__ void __launch_bounds__(256)
ncc_finish_kernel(const float* __restrict__ partial, int per_image, int stride, float scale,
                  float* __restrict__ score, int accumulate) {
  __shared__ float red[8];
  const float* p = partial + (int64_t)blockIdx.x * per_image * stride;
  float acc = 0.f;
  for (int i = threadIdx.x; i < per_image; i += 256) acc += p[(int64_t)i * stride];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) score[blockIdx.x] = accumulate ? score[blockIdx.x] + v * scale : v * scale;
  }
}

// ---------------------------------------------------------------------------------------------- Sobel
// out[b,0] = Gx * x, out[b,1] = Gy * x (cross-correlation, zero padding 1), Gx = [[1,0,-1],[2,0,-2],[1,0,-1]],
// Gy = [[1,2,1],[0,0,0],[-1,-2,-1]].
__global__ void __launch_bounds__(256)
sobel_fwd_kernel(const float* __restrict__ x, int H, int W, float* __restrict__ out) {
  const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (px >= W || py >= H) return;
  const float* im = x + (int64_t)blockIdx.z * H * W;
  auto at = [&](int y, int xx) -> float { return (y >= 0 && y < H && xx >= 0 && xx < W) ? __ldg(im + y * W + xx) : 0.f; };
  const float a = at(py - 1, px - 1), b = at(py - 1, px), c = at(py - 1, px + 1);
  const float d = at(py, px - 1), f = at(py, px + 1);
  const float g = at(py + 1, px - 1), h = at(py + 1, px), i = at(py + 1, px + 1);
  float* o = out + (int64_t)blockIdx.z * 2 * H * W + py * W + px;
  o[0] = (a - c) + 2.f * (d - f) + (g - i);
  o[(int64_t)H * W] = (a - g) + 2.f * (b - h) + (c - i);
}

// transpose: gx[b,y,x] = sum over output pixels that read (y,x)
__global__ void __launch_bounds__(256)
sobel_bwd_kernel(const float* __restrict__ gout, int H, int W, float* __restrict__ gx) {
  const int px = blockIdx.x * 32 + (threadIdx.x & 31), py = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (px >= W || py >= H) return;
  const float* g0 = gout + (int64_t)blockIdx.z * 2 * H * W;
  const float* g1 = g0 + (int64_t)H * W;
  auto at = [&](const float* g, int y, int xx) -> float {
    return (y >= 0 && y < H && xx >= 0 && xx < W) ? __ldg(g + y * W + xx) : 0.f;
  };
  // out(y', x') reads in(y'+dy, x'+dx) with weight K[dy+1][dx+1]  =>  in(y,x) feeds out(y-dy, x-dx)
  float acc = 0.f;
  const float kx[3][3] = {{1.f, 0.f, -1.f}, {2.f, 0.f, -2.f}, {1.f, 0.f, -1.f}};
  const float ky[3][3] = {{1.f, 2.f, 1.f}, {0.f, 0.f, 0.f}, {-1.f, -2.f, -1.f}};
#pragma unroll
  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
    for (int dx = -1; dx <= 1; ++dx) {
      acc += kx[dy + 1][dx + 1] * at(g0, py - dy, px - dx) + ky[dy + 1][dx + 1] * at(g1, py - dy, px - dx);
    }
  gx[(int64_t)blockIdx.z * H * W + py * W + px] = acc;
}


// ---------------------------------------------------------------------------------------------- registration similarity
// One registration iteration scores  S = sum_b  w_g NCC(f, y) + w_p NCC_p(f, y) + w_s NCC_q(Sobel f, Sobel y)  with
// y = XrayTransforms(x) = ((x - min x) / (max x - min x + e) - mean) / std  of the raw DRR batch x and a fixed,
// already transformed X-ray f (/root/reference/src/xvr/registrar/base.py:245-252, utils/preprocess.py:5-29), and
// immediately differentiates it.  Composed from tensor ops that is ~60 launches around six real kernels; the three
// kernels below supply the ends (standardise + Sobel in, score out, standardise-backward out) so that the whole
// value-and-gradient evaluation is nine launches.  The two ends that touch every pixel run as ONE THREAD-BLOCK
// CLUSTER of RS_CLUSTER CTAs each: a registration batch is a single 256 x 256 image, the batch-global min / max / tie
// counts / gradient sums are exchanged through distributed shared memory behind cluster barriers (fixed order:
// deterministic), so the reductions need no second launch and the pixels are spread over RS_CLUSTER SMs.
// (Round 1 ran them as one CTA: three passes over 65 536 pixels on a single SM cost ~150 us each and made the fused
// path slower than the ~50 small launches it replaces.)

// Fixed-tree reductions over a 1024-thread CTA; every thread receives the result.
template <typename T, typename Op>
__device__ __forceinline__ T cta1024_reduce(T v, Op op, T* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();  // red may still be read from the previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  v = red[threadIdx.x & 31];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

constexpr int RS_CLUSTER = 8;  // CTAs per cluster (portable maximum)

// value of `mine` from every CTA of the cluster, combined in rank order (deterministic); `slot` is a __shared__ word
template <typename T, typename Op>
__device__ __forceinline__ T cluster_combine(cg::cluster_group& cluster, T* slot, T mine, Op op) {
  if (threadIdx.x == 0) *slot = mine;
  cluster.sync();
  T v = *cluster.map_shared_rank(slot, 0);
  for (unsigned r = 1; r < cluster.num_blocks(); ++r) v = op(v, *cluster.map_shared_rank(slot, r));
  cluster.sync();  // nobody overwrites its slot (or exits) while a peer still reads it
  return v;
}

// stats = {min, max, #elements equal to min, #elements equal to max, max - min + e}; ones (B) <- 1;
// y (B,1,H,W) <- XrayTransforms(x); sob (B,2,H,W) <- Sobel(y) with zero padding in y-space (conv2d padding=1).
__global__ void __cluster_dims__(RS_CLUSTER, 1, 1) __launch_bounds__(1024)
regsim_prep_kernel(const float* __restrict__ x, int B, int H, int W, float std_eps, float mean, float inv_std,
                   float* __restrict__ stats, float* __restrict__ ones, float* __restrict__ y,
                   float* __restrict__ sob) {
  __shared__ float redf[32];
  __shared__ int redi[32];
  __shared__ float slot_f;
  __shared__ int slot_i;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = threadIdx.x;
  const int gtid = (int)cluster.block_rank() * 1024 + tid, stride = (int)cluster.num_blocks() * 1024;
  const int HW = H * W, n = B * HW;
  float lo = INFINITY, hi = -INFINITY;
  for (int i = gtid; i < n; i += stride) {
    const float v = __ldg(x + i);
    lo = fminf(lo, v);
    hi = fmaxf(hi, v);
  }
  lo = cta1024_reduce(lo, [](float a, float b) { return fminf(a, b); }, redf);
  lo = cluster_combine(cluster, &slot_f, lo, [](float a, float b) { return fminf(a, b); });
  hi = cta1024_reduce(hi, [](float a, float b) { return fmaxf(a, b); }, redf);
  hi = cluster_combine(cluster, &slot_f, hi, [](float a, float b) { return fmaxf(a, b); });
  int nlo = 0, nhi = 0;
  for (int i = gtid; i < n; i += stride) {
    const float v = __ldg(x + i);
    nlo += v == lo;
    nhi += v == hi;
  }
  nlo = cta1024_reduce(nlo, [](int a, int b) { return a + b; }, redi);
  nlo = cluster_combine(cluster, &slot_i, nlo, [](int a, int b) { return a + b; });
  nhi = cta1024_reduce(nhi, [](int a, int b) { return a + b; }, redi);
  nhi = cluster_combine(cluster, &slot_i, nhi, [](int a, int b) { return a + b; });
  const float r = (hi - lo) + std_eps;
  if (gtid == 0) {
    stats[0] = lo;
    stats[1] = hi;
    stats[2] = (float)nlo;
    stats[3] = (float)nhi;
    stats[4] = r;
  }
  for (int b = gtid; b < B; b += stride) ones[b] = 1.f;
  for (int i = gtid; i < n; i += stride) {
    const int b = i / HW, rem = i - b * HW;
    const int py = rem / W, px = rem - py * W;
    const float* im = x + (int64_t)b * HW;
    auto at = [&](int yy, int xx) -> float {
      if (yy < 0 || yy >= H || xx < 0 || xx >= W) return 0.f;
      return (__fdiv_rn(__ldg(im + yy * W + xx) - lo, r) - mean) * inv_std;
    };
    const float a = at(py - 1, px - 1), bb = at(py - 1, px), c = at(py - 1, px + 1);
    const float d = at(py, px - 1), e = at(py, px), f = at(py, px + 1);
    const float g = at(py + 1, px - 1), h = at(py + 1, px), k = at(py + 1, px + 1);
    y[i] = e;
    float* o = sob + (int64_t)b * 2 * HW + rem;
    o[0] = (a - c) + 2.f * (d - f) + (g - k);
    o[HW] = (a - g) + 2.f * (bb - h) + (c - k);
  }
}

// score[0] = w_g sum_b NCC_b + w_p sum(partial_p) + w_s sum(partial_s)  (weights already hold the 1/count factors)
__global__ void __launch_bounds__(1024)
regsim_score_kernel(const float* __restrict__ gstat, int B, float w_g, const float* __restrict__ partial_p, int n_p,
                    float w_p, const float* __restrict__ partial_s, int n_s, float w_s, float* __restrict__ score) {
  __shared__ float redf[32];
  const int tid = threadIdx.x;
  float acc_g = 0.f, acc_p = 0.f, acc_s = 0.f;
  for (int b = tid; b < B; b += 1024) acc_g += gstat[(int64_t)b * 6];
  for (int i = tid; i < n_p; i += 1024) acc_p += partial_p[i];
  for (int i = tid; i < n_s; i += 1024) acc_s += partial_s[i];
  const float v = cta1024_reduce(w_g * acc_g + w_p * acc_p + w_s * acc_s, [](float a, float b) { return a + b; }, redf);
  if (tid == 0) score[0] = v;
}

// g_y (B,1,H,W): dS/dy from the NCC terms on entry; Sobel^T of g_sob (B,2,H,W) is added, then the chain through
// y = ((x - lo) / r - mean) * inv_std with lo = min x, r = max x - min x + e (the extrema receive their share evenly
// over ties, as torch's min()/max() backward distributes it):
//   dS/dx_j = a g_j + [x_j = lo] (c S1 - a S0) / n_lo - [x_j = hi] c S1 / n_hi,
//   a = inv_std / r, c = inv_std / r^2, S0 = sum g, S1 = sum g (x - lo).
__global__ void __cluster_dims__(RS_CLUSTER, 1, 1) __launch_bounds__(1024)
regsim_bwd_finish_kernel(const float* __restrict__ x, const float* __restrict__ g_sob, const float* __restrict__ stats,
                         int B, int H, int W, float inv_std, float* __restrict__ g_y, float* __restrict__ grad) {
  __shared__ float redf[32];
  __shared__ float slot_f;
  cg::cluster_group cluster = cg::this_cluster();
  const int tid = (int)cluster.block_rank() * 1024 + threadIdx.x, stride = (int)cluster.num_blocks() * 1024;
  const int HW = H * W, n = B * HW;
  const float lo = stats[0], hi = stats[1], nlo = stats[2], nhi = stats[3], r = stats[4];
  const float kx[3][3] = {{1.f, 0.f, -1.f}, {2.f, 0.f, -2.f}, {1.f, 0.f, -1.f}};
  const float ky[3][3] = {{1.f, 2.f, 1.f}, {0.f, 0.f, 0.f}, {-1.f, -2.f, -1.f}};
  float s0 = 0.f, s1 = 0.f;
  for (int i = tid; i < n; i += stride) {
    const int b = i / HW, rem = i - b * HW;
    const int py = rem / W, px = rem - py * W;
    const float* g0 = g_sob + (int64_t)b * 2 * HW;
    const float* g1 = g0 + HW;
    float acc = 0.f;
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int yy = py - dy, xx = px - dx;
        if (yy >= 0 && yy < H && xx >= 0 && xx < W)
          acc += kx[dy + 1][dx + 1] * __ldg(g0 + yy * W + xx) + ky[dy + 1][dx + 1] * __ldg(g1 + yy * W + xx);
      }
    const float g = g_y[i] + acc;
    g_y[i] = g;
    s0 += g;
    s1 = fmaf(g, __ldg(x + i) - lo, s1);
  }
  float S0 = cta1024_reduce(s0, [](float a, float b) { return a + b; }, redf);
  S0 = cluster_combine(cluster, &slot_f, S0, [](float a, float b) { return a + b; });
  float S1 = cta1024_reduce(s1, [](float a, float b) { return a + b; }, redf);
  S1 = cluster_combine(cluster, &slot_f, S1, [](float a, float b) { return a + b; });
  const float a = inv_std / r;
  const float cS1 = inv_std * S1 / (r * r);
  const float t_lo = (cS1 - a * S0) / nlo, t_hi = -cS1 / nhi;
  for (int i = tid; i < n; i += stride) {  // every thread re-reads only what it wrote itself
    const float v = __ldg(x + i);
    float out = a * g_y[i];
    if (v == lo) out += t_lo;
    if (v == hi) out += t_hi;
    grad[i] = out;
  }
}

// carve-up of the caller's workspace (in floats)
struct RegSimLayout {
  int64_t stats, ones, y, sob, gstat, part_p, part_s, coef_p, coef_s, g_y, g_sob, total;
  int tiles_p, tiles_s;
};

static RegSimLayout regsim_layout(int B, int H, int W, int p, int q) {
  RegSimLayout L;
  const int64_t HW = (int64_t)H * W;
  auto tiles = [&](int k) { return ((H - k + 1 + NCC_T - 1) / NCC_T) * ((W - k + 1 + NCC_T - 1) / NCC_T); };
  L.tiles_p = tiles(p);
  L.tiles_s = tiles(q);
  int64_t o = 0;
  auto take = [&](int64_t n) { const int64_t at = o; o += (n + 3) & ~(int64_t)3; return at; };
  L.stats = take(8);
  L.ones = take(B);
  L.y = take(B * HW);
  L.sob = take(2 * B * HW);
  L.gstat = take((int64_t)B * 6);
  L.part_p = take((int64_t)B * L.tiles_p);
  L.part_s = take((int64_t)B * 2 * L.tiles_s);
  L.coef_p = take((int64_t)B * 4 * (H - p + 1) * (W - p + 1));
  L.coef_s = take((int64_t)B * 2 * 4 * (H - q + 1) * (W - q + 1));
  L.g_y = take(B * HW);
  L.g_sob = take(2 * B * HW);
  L.total = o;
  return L;
}

}  // namespace xvr

using namespace xvr;

static size_t ncc_smem(int p, int planes) { return (size_t)planes * (NCC_T + p - 1) * (NCC_T + p - 1) * sizeof(float); }

// score (B,) (+)= weight * NCC_p(x1, x2).  patch <= 0 selects the global (whole-image) NCC.
// workspace: >= B*C*max(6, tiles) floats.  coef2 / coef1 (nullable) receive what xvr_ncc_bwd needs for the
// gradient w.r.t. x2 / x1: (B*C,4,nH,nW) floats for a patch NCC, (B*C,6) for the global one.
extern "C" int xvr_ncc_fwd(const float* x1, const float* x2, int B, int C, int H, int W, int patch, float eps,
                           float weight, int accumulate, float* score, float* workspace, float* coef2, float* coef1,
                           void* stream) {
  if (!x1 || !x2 || !score || !workspace || B <= 0 || C <= 0 || H <= 0 || W <= 0 || patch > H || patch > W ||
      patch > 64) {
    set_last_error("xvr_ncc_fwd: invalid argument (patch must be <= min(H, W, 64))");
    return XVR_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (patch <= 0) {
    float* stats = coef2 ? coef2 : (coef1 ? coef1 : workspace);
    global_ncc_fwd_kernel<<<B * C, 1024, 0, st>>>(x1, x2, H * W, eps, stats);
    int rc = check_launch("xvr_ncc_fwd/global");
    if (rc) return rc;
    if (coef2 && coef1) cudaMemcpyAsync(coef1, coef2, (size_t)B * C * 6 * sizeof(float), cudaMemcpyDeviceToDevice, st);
    ncc_finish_kernel<<<B, 256, 0, st>>>(stats, C, 6, weight / (float)C, score, accumulate);
    return check_launch("xvr_ncc_fwd/finish");
  }
  const int nH = H - patch + 1, nW = W - patch + 1;
  dim3 grid((nW + NCC_T - 1) / NCC_T, (nH + NCC_T - 1) / NCC_T, B * C);
  const size_t smem = ncc_smem(patch, 2);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(patch_ncc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  patch_ncc_fwd_kernel<<<grid, dim3(NCC_T, NCC_T), smem, st>>>(x1, x2, H, W, patch, eps, workspace, coef2, coef1);
  int rc = check_launch("xvr_ncc_fwd/patch");
  if (rc) return rc;
  const int tiles = grid.x * grid.y;
  ncc_finish_kernel<<<B, 256, 0, st>>>(workspace, C * tiles, 1, weight / ((float)C * nH * nW), score, accumulate);
  return check_launch("xvr_ncc_fwd/finish");
}

// grad (B,C,H,W) (+)= gscore[b] * weight * d NCC_p / d x_which, from the coefficients saved by xvr_ncc_fwd.
extern "C" int xvr_ncc_bwd(const float* x1, const float* x2, const float* coef, int which, const float* gscore,
                           int B, int C, int H, int W, int patch, float weight, int accumulate, float* grad,
                           void* stream) {
  if (!x1 || !x2 || !coef || !gscore || !grad || (which != 1 && which != 2) || B <= 0 || C <= 0) {
    set_last_error("xvr_ncc_bwd: invalid argument");
    return XVR_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const float* xa = which == 2 ? x1 : x2;
  const float* xb = which == 2 ? x2 : x1;
  if (patch <= 0) {
    const int n = H * W;
    global_ncc_bwd_kernel<<<dim3((n + 255) / 256, B * C), 256, 0, st>>>(xa, xb, coef, which, gscore,
                                                                        weight / ((float)C * n), C, n, grad, accumulate);
    return check_launch("xvr_ncc_bwd/global");
  }
  const int nH = H - patch + 1, nW = W - patch + 1;
  dim3 grid((W + NCC_T - 1) / NCC_T, (H + NCC_T - 1) / NCC_T, B * C);
  const size_t smem = ncc_smem(patch, 4);
  if (smem > 48 * 1024)
    cudaFuncSetAttribute(patch_ncc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  patch_ncc_bwd_kernel<<<grid, dim3(NCC_T, NCC_T), smem, st>>>(
      xa, xb, coef, gscore, weight / ((float)C * nH * nW * patch * patch), C, H, W, patch, grad, accumulate);
  return check_launch("xvr_ncc_bwd/patch");
}

extern "C" int xvr_sobel_fwd(const float* x, int B, int H, int W, float* out, void* stream) {
  if (!x || !out || B <= 0 || H <= 0 || W <= 0) {
    set_last_error("xvr_sobel_fwd: invalid argument");
    return XVR_ERR_INVALID;
  }
  sobel_fwd_kernel<<<dim3((W + 31) / 32, (H + 7) / 8, B), 256, 0, (cudaStream_t)stream>>>(x, H, W, out);
  return check_launch("xvr_sobel_fwd");
}

extern "C" int xvr_sobel_bwd(const float* gout, int B, int H, int W, float* gx, void* stream) {
  if (!gout || !gx || B <= 0 || H <= 0 || W <= 0) {
    set_last_error("xvr_sobel_bwd: invalid argument");
    return XVR_ERR_INVALID;
  }
  sobel_bwd_kernel<<<dim3((W + 31) / 32, (H + 7) / 8, B), 256, 0, (cudaStream_t)stream>>>(gout, H, W, gx);
  return check_launch("xvr_sobel_bwd");
}

// ---- value and gradient of the registration similarity in nine launches (see regsim_prep_kernel).
static bool regsim_args_ok(int B, int H, int W, int p, int q) {
  return B > 0 && H > 0 && W > 0 && p >= 1 && q >= 1 && p <= H && p <= W && q <= H && q <= W && p <= 64 && q <= 64 &&
         (int64_t)B * H * W <= ((int64_t)1 << 24);
}

extern "C" long long xvr_regsim_workspace_floats(int B, int H, int W, int patch_mncc, int patch_gncc) {
  if (!regsim_args_ok(B, H, W, patch_mncc, patch_gncc)) return -1;
  return regsim_layout(B, H, W, patch_mncc, patch_gncc).total;
}

// fixed (B,1,H,W): the transformed target X-ray; fixed_sobel (B,2,H,W) = Sobel(fixed); moving (B,1,H,W): raw DRRs.
// score[0] = sum_b w_global NCC + w_patch NCC_p + w_grad NCC_q(Sobel) of (fixed, XrayTransforms(moving));
// grad (B,1,H,W) = d score / d moving.  The transform is Standardize(std_eps) -> (. - mean) * inv_std.
extern "C" int xvr_regsim(const float* fixed, const float* fixed_sobel, const float* moving, int B, int H, int W,
                          float std_eps, float mean, float inv_std, int patch_mncc, int patch_gncc, float w_global,
                          float w_patch, float w_grad, float ncc_eps, float* workspace, long long workspace_floats,
                          float* score, float* grad, void* stream) {
  if (!fixed || !fixed_sobel || !moving || !workspace || !score || !grad ||
      !regsim_args_ok(B, H, W, patch_mncc, patch_gncc)) {
    set_last_error("xvr_regsim: invalid argument (patches must be <= min(H, W, 64), B*H*W <= 2^24)");
    return XVR_ERR_INVALID;
  }
  const int p = patch_mncc, q = patch_gncc;
  const RegSimLayout L = regsim_layout(B, H, W, p, q);
  if (workspace_floats < L.total) {
    set_last_error("xvr_regsim: workspace smaller than xvr_regsim_workspace_floats()");
    return XVR_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* ws = workspace;
  const int HW = H * W;
  int rc;
  regsim_prep_kernel<<<RS_CLUSTER, 1024, 0, st>>>(moving, B, H, W, std_eps, mean, inv_std, ws + L.stats, ws + L.ones, ws + L.y,
                                         ws + L.sob);
  if ((rc = check_launch("xvr_regsim/prep"))) return rc;

  global_ncc_fwd_kernel<<<B, 1024, 0, st>>>(fixed, ws + L.y, HW, ncc_eps, ws + L.gstat);
  if ((rc = check_launch("xvr_regsim/global"))) return rc;
  const int nHp = H - p + 1, nWp = W - p + 1, nHq = H - q + 1, nWq = W - q + 1;
  const size_t smem_fwd = ncc_smem(p > q ? p : q, 2), smem_bwd = ncc_smem(p > q ? p : q, 4);
  if (smem_fwd > 48 * 1024)
    cudaFuncSetAttribute(patch_ncc_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fwd);
  if (smem_bwd > 48 * 1024)
    cudaFuncSetAttribute(patch_ncc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bwd);
  patch_ncc_fwd_kernel<<<dim3((nWp + NCC_T - 1) / NCC_T, (nHp + NCC_T - 1) / NCC_T, B), dim3(NCC_T, NCC_T),
                         ncc_smem(p, 2), st>>>(fixed, ws + L.y, H, W, p, ncc_eps, ws + L.part_p, ws + L.coef_p, nullptr);
  if ((rc = check_launch("xvr_regsim/patch"))) return rc;
  patch_ncc_fwd_kernel<<<dim3((nWq + NCC_T - 1) / NCC_T, (nHq + NCC_T - 1) / NCC_T, B * 2), dim3(NCC_T, NCC_T),
                         ncc_smem(q, 2), st>>>(fixed_sobel, ws + L.sob, H, W, q, ncc_eps, ws + L.part_s, ws + L.coef_s,
                                               nullptr);
  if ((rc = check_launch("xvr_regsim/gradient patch"))) return rc;
  regsim_score_kernel<<<1, 1024, 0, st>>>(ws + L.gstat, B, w_global, ws + L.part_p, B * L.tiles_p,
                                          w_patch / ((float)nHp * nWp), ws + L.part_s, B * 2 * L.tiles_s,
                                          w_grad / (2.f * nHq * nWq), score);
  if ((rc = check_launch("xvr_regsim/score"))) return rc;

  // d/dy of the three terms (gather form, see patch_ncc_bwd_kernel), then back through Sobel and the transform
  global_ncc_bwd_kernel<<<dim3((HW + 255) / 256, B), 256, 0, st>>>(fixed, ws + L.y, ws + L.gstat, 2, ws + L.ones,
                                                                   w_global / (float)HW, 1, HW, ws + L.g_y, 0);
  if ((rc = check_launch("xvr_regsim/global bwd"))) return rc;
  patch_ncc_bwd_kernel<<<dim3((W + NCC_T - 1) / NCC_T, (H + NCC_T - 1) / NCC_T, B), dim3(NCC_T, NCC_T), ncc_smem(p, 4),
                         st>>>(fixed, ws + L.y, ws + L.coef_p, ws + L.ones, w_patch / ((float)nHp * nWp * p * p), 1, H,
                               W, p, ws + L.g_y, 1);
  if ((rc = check_launch("xvr_regsim/patch bwd"))) return rc;
  patch_ncc_bwd_kernel<<<dim3((W + NCC_T - 1) / NCC_T, (H + NCC_T - 1) / NCC_T, B * 2), dim3(NCC_T, NCC_T),
                         ncc_smem(q, 4), st>>>(fixed_sobel, ws + L.sob, ws + L.coef_s, ws + L.ones,
                                               w_grad / (2.f * nHq * nWq * q * q), 2, H, W, q, ws + L.g_sob, 0);
  if ((rc = check_launch("xvr_regsim/gradient patch bwd"))) return rc;
  regsim_bwd_finish_kernel<<<RS_CLUSTER, 1024, 0, st>>>(moving, ws + L.g_sob, ws + L.stats, B, H, W, inv_std, ws + L.g_y, grad);
  return check_launch("xvr_regsim/finish");
}
