// dL/dvolume of the fused trilinear DRR, atomics-free and deterministic.
//
// The reference obtains this gradient from grid_sampler_3d_backward, which scatters every sample into its 8
// corner voxels with fastAtomicAdd (ATen/native/cuda/GridSampler.cuh:263-280, safe_add_3d): non-deterministic,
// and 8 L2 atomics per sample.  Here the adjoint is evaluated in GATHER form: a thread owns one voxel, and for
// every pose it inverts the projection (voxel -> detector pixel), visits the few rays whose samples can fall
// within one voxel of it, re-creates those samples with the forward kernel's own arithmetic and sums
//        dL/dV[p] = sum_b sum_r c_{b,r} sum_k prod_a max(0, 1 - |x_{b,r,k,a} - p_a|),   c = g * L * w(span).
// Each voxel is written exactly once by its owner; the summation order is fixed.
#include "common.cuh"

namespace xvr {

struct VolGradParams {
  int D0, D1, D2;
  DetectorGeom geom;                  // cam2vox (B,3,4) + detector basis
  const float* __restrict__ vox2cam;  // (B,3,4) inverse of cam2vox
  const float4* __restrict__ info;    // (B,N,3): {amin, span, c, 0}, {d0, d1, d2, 0}, {1/d0, 1/d1, 1/d2, 0}
  const float* __restrict__ gout;     // (B,N)
  float4* __restrict__ info_out;
  int B, H, W, n_points, step_mode;
  float eps;
  float* __restrict__ gvol;
  int accumulate;
};

__device__ __forceinline__ float vg_step_weight(int mode, float span, int n) {
  if (mode == 0) return __fdiv_rn(span, (float)(n - 1));
  if (mode == 1) return __fdiv_rn(span, (float)n);
  return __fdiv_rn(1.0f, (float)n);
}

// Pre-pass: per ray {amin, span, g * L * w(span)} exactly as trilinear_fwd_kernel derives them.
__global__ void __launch_bounds__(256) ray_info_kernel(const VolGradParams p) {
  const int64_t ray = (int64_t)blockIdx.x * 256 + threadIdx.x;
  const int N = p.H * p.W;
  if (ray >= (int64_t)p.B * N) return;
  const int b = (int)(ray / N), n = (int)(ray - (int64_t)b * N);
  float s[3], d[3], L;
  generate_ray(p.geom, b, n, p.eps, s, d, L);
  const float lo[3] = {0.f, 0.f, 0.f};
  const float hi[3] = {(float)(p.D0 - 1), (float)(p.D1 - 1), (float)(p.D2 - 1)};
  const AlphaRange ar = alpha_range(s, d, lo, hi);
  const float span = ar.amax - ar.amin;
  // rays that stay more than one voxel away from the volume contribute nothing (the forward skips them too)
  const float plo[3] = {-1.f, -1.f, -1.f};
  const float phi[3] = {(float)p.D0, (float)p.D1, (float)p.D2};
  const AlphaRange pr = alpha_range(s, d, plo, phi);
  const float c = (pr.amin < pr.amax) ? __ldg(p.gout + ray) * L * vg_step_weight(p.step_mode, span, p.n_points) : 0.f;
  p.info_out[ray * 3 + 0] = make_float4(ar.amin, span, c, 0.f);
  p.info_out[ray * 3 + 1] = make_float4(d[0], d[1], d[2], 0.f);
  p.info_out[ray * 3 + 2] = make_float4(1.0f / d[0], 1.0f / d[1], 1.0f / d[2], 0.f);
}

__global__ void __launch_bounds__(256) volume_grad_kernel(const VolGradParams p) {
  // one thread per voxel, axis 2 fastest (coalesced store); 64 x 4 voxel tiles of (axis 2, axis 1)
  const int z = blockIdx.x * 64 + (threadIdx.x & 63);
  const int y = blockIdx.y * 4 + (threadIdx.x >> 6);
  const int x = blockIdx.z;
  if (z >= p.D2 || y >= p.D1) return;
  const float pv[3] = {(float)x, (float)y, (float)z};
  const int N = p.H * p.W;
  const int np = p.n_points;
  const float lstep = 1.0f / (float)(np - 1);
  const float inv_vx = 1.0f / p.geom.v[0], inv_uy = 1.0f / p.geom.u[1];
  const float sdd = p.geom.o[2];
  float acc = 0.f;

  for (int b = 0; b < p.B; ++b) {
    const float* Gi = p.vox2cam + b * 12;
    float q[3], row[3][3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
#pragma unroll
      for (int c = 0; c < 3; ++c) row[a][c] = __ldg(Gi + a * 4 + c);
      q[a] = fmaf(row[a][2], pv[2], fmaf(row[a][1], pv[1], fmaf(row[a][0], pv[0], __ldg(Gi + a * 4 + 3))));
    }
    if (!(q[2] > 1e-3f)) continue;  // behind the source
    const float m = sdd / q[2];
    // continuous pixel coordinates of the voxel centre and the half-extent of its +-1 voxel support
    const float cj = (q[0] * m - p.geom.o[0]) * inv_vx;
    const float ci = (q[1] * m - p.geom.o[1]) * inv_uy;
    float rj = 0.f, ri = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      rj += fabsf(row[0][a] - (q[0] / q[2]) * row[2][a]);
      ri += fabsf(row[1][a] - (q[1] / q[2]) * row[2][a]);
    }
    rj = rj * m * fabsf(inv_vx) * 1.01f + 1e-3f;
    ri = ri * m * fabsf(inv_uy) * 1.01f + 1e-3f;
    const int j0 = max(0, (int)ceilf(cj - rj)), j1 = min(p.W - 1, (int)floorf(cj + rj));
    const int i0 = max(0, (int)ceilf(ci - ri)), i1 = min(p.H - 1, (int)floorf(ci + ri));
    if (i0 > i1 || j0 > j1) continue;
    // the source is the translation column of cam2vox (what generate_ray uses)
    const float* G = p.geom.cam2vox + b * 12;
    const float s0 = __ldg(G + 3), s1 = __ldg(G + 7), s2 = __ldg(G + 11);
    const float4* inf = p.info + (int64_t)b * N * 3;
    for (int i = i0; i <= i1; ++i) {
      for (int j = j0; j <= j1; ++j) {
        const float4* q4 = inf + (int64_t)(i * p.W + j) * 3;
        const float4 a = __ldg(q4);
        if (a.z == 0.f) continue;
        const float4 d = __ldg(q4 + 1), r = __ldg(q4 + 2);
        // alpha window in which the ray is within one voxel of p on every axis
        const float e0 = pv[0] - s0, e1 = pv[1] - s1, e2 = pv[2] - s2;
        const float x0 = (e0 - 1.f) * r.x, x1 = (e0 + 1.f) * r.x;
        const float y0 = (e1 - 1.f) * r.y, y1 = (e1 + 1.f) * r.y;
        const float z0 = (e2 - 1.f) * r.z, z1 = (e2 + 1.f) * r.z;
        const float alo = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
        const float ahi = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
        if (!(alo < ahi)) continue;
        // sample indices: alpha_k = amin + u_k span, u_k ~ k/(n-1); pad the range slightly, the weights decide
        const float sc = (float)(np - 1) / a.y;
        const float k0f = (alo - a.x) * sc, k1f = (ahi - a.x) * sc;
        const float klo = fmaxf(fminf(k0f, k1f) - 0.02f, 0.f), khi = fminf(fmaxf(k0f, k1f) + 0.02f, (float)(np - 1));
        if (!(klo <= khi)) continue;
        float wsum = 0.f;
        for (int k = (int)ceilf(klo); k <= (int)floorf(khi); ++k) {
          const float u = linspace01(k, np, lstep);
          const float alpha = fmaf(u, a.y, a.x);
          const float dx = fabsf(fmaf(alpha, d.x, s0) - pv[0]);
          const float dy = fabsf(fmaf(alpha, d.y, s1) - pv[1]);
          const float dz = fabsf(fmaf(alpha, d.z, s2) - pv[2]);
          if (dx < 1.f && dy < 1.f && dz < 1.f) wsum = fmaf((1.f - dx) * (1.f - dy), 1.f - dz, wsum);
        }
        acc = fmaf(a.z, wsum, acc);
      }
    }
  }
  const int64_t o = ((int64_t)x * p.D1 + y) * p.D2 + z;
  p.gvol[o] = p.accumulate ? p.gvol[o] + acc : acc;
}

}  // namespace xvr

using namespace xvr;

// gvol (D0,D1,D2) (+)= dL/dvolume of xvr_trilinear_drr_fwd for upstream gradient gout (B,1,H*W).
// vox2cam (B,3,4) is the inverse of cam2vox; workspace holds 3*B*H*W float4.  The detector basis must be axis
// aligned in the camera frame (row step along y, column step along x), as DiffDRR's detector is.
extern "C" int xvr_trilinear_drr_bwd_volume(const float* cam2vox, const float* vox2cam, const float* cam2world,
                                            const float* det9, int B, int det_h, int det_w, int n_points,
                                            int step_mode, float eps, const float* gout, int D0, int D1, int D2,
                                            float* workspace, float* gvol, int accumulate, void* stream) {
  if (!cam2vox || !vox2cam || !cam2world || !det9 || !gout || !workspace || !gvol || B <= 0 || det_h <= 0 ||
      det_w <= 0 || n_points < 2 || D0 < 2 || D1 < 2 || D2 < 2) {
    set_last_error("xvr_trilinear_drr_bwd_volume: invalid argument");
    return XVR_ERR_INVALID;
  }
  if (det9[3] != 0.f || det9[5] != 0.f || det9[7] != 0.f || det9[8] != 0.f || det9[4] == 0.f || det9[6] == 0.f) {
    set_last_error("xvr_trilinear_drr_bwd_volume: detector basis must be axis aligned (row step = (0,dy,0), "
                   "column step = (dx,0,0))");
    return XVR_ERR_INVALID;
  }
  VolGradParams p = {};
  p.D0 = D0; p.D1 = D1; p.D2 = D2;
  p.geom.cam2vox = cam2vox;
  p.geom.cam2world = cam2world;
  for (int a = 0; a < 3; ++a) { p.geom.o[a] = det9[a]; p.geom.u[a] = det9[3 + a]; p.geom.v[a] = det9[6 + a]; }
  p.geom.W = det_w;
  p.vox2cam = vox2cam;
  p.info = (const float4*)workspace;
  p.info_out = (float4*)workspace;
  p.gout = gout;
  p.B = B; p.H = det_h; p.W = det_w; p.n_points = n_points; p.step_mode = step_mode; p.eps = eps;
  p.gvol = gvol;
  p.accumulate = accumulate;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rays = (int64_t)B * det_h * det_w;
  ray_info_kernel<<<(unsigned)((rays + 255) / 256), 256, 0, st>>>(p);
  int rc = check_launch("xvr_trilinear_drr_bwd_volume/info");
  if (rc) return rc;
  dim3 grid((D2 + 63) / 64, (D1 + 3) / 4, D0);
  volume_grad_kernel<<<grid, 256, 0, st>>>(p);
  return check_launch("xvr_trilinear_drr_bwd_volume");
}
