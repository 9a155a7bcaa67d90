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
    // Candidate rays = pixel box of the 8 projected corners of the voxel's +-1 support (a convex set in front of the
    // source projects inside the box of its projected vertices).  A support that reaches the source plane has no
    // finite window -- every ray is a candidate (voxels around a source placed inside the volume: every ray's first
    // sample, alpha = 0, is the source itself) -- and one entirely behind the source is never reached.
    // (Until the end of round 1 the window came from the projection linearised at the voxel centre with 1 % of
    // padding: only a bound when the support's depth is < 1 % of the voxel's, and voxels behind the source plane were
    // skipped -- 6 % off with the source inside the volume, 3e-4 off after a partial fix; scripts/check_gather_window.py.)
    float jmin = INFINITY, jmax = -INFINITY, imin = INFINITY, imax = -INFINITY;
    int in_front = 0;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float sx = (c & 1) ? 1.f : -1.f, sy = (c & 2) ? 1.f : -1.f, sz = (c & 4) ? 1.f : -1.f;
      const float q0 = q[0] + sx * row[0][0] + sy * row[0][1] + sz * row[0][2];
      const float q1 = q[1] + sx * row[1][0] + sy * row[1][1] + sz * row[1][2];
      const float q2 = q[2] + sx * row[2][0] + sy * row[2][1] + sz * row[2][2];
      if (q2 > 1e-3f * sdd) {
        const float m = sdd / q2;
        const float cj = (q0 * m - p.geom.o[0]) * inv_vx, ci = (q1 * m - p.geom.o[1]) * inv_uy;
        jmin = fminf(jmin, cj); jmax = fmaxf(jmax, cj);
        imin = fminf(imin, ci); imax = fmaxf(imax, ci);
        ++in_front;
      }
    }
    if (in_front == 0) continue;  // behind the source
    int j0 = 0, j1 = p.W - 1, i0 = 0, i1 = p.H - 1;
    if (in_front == 8) {
      j0 = max(0, (int)ceilf(jmin - 1e-3f)); j1 = min(p.W - 1, (int)floorf(jmax + 1e-3f));
      i0 = max(0, (int)ceilf(imin - 1e-3f)); i1 = min(p.H - 1, (int)floorf(imax + 1e-3f));
    }
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


// ------------------------------------------------------------------------------------------------------------------
// Version 2 (the default; xvr_set_volgrad_version(1) selects the gather above as an independent cross-check):
// brick-local scatter, no atomics, deterministic.  Passes the oracle-parity / determinism tests including
// anisotropic voxels, a shifted / reversed / non-square detector and a source inside the volume, agrees with
// version 1 to 8e-8 relative L2 at config-2 scale and is 1.7x faster there (12.2 vs 21.0 ms per 8 poses).
//
// A warp owns a 16^3 brick of the gradient volume in shared memory.  For every pose it projects the brick (grown by
// the one-voxel support of the trilinear hat) onto the detector, walks the rays of that pixel window and
// re-creates, with the forward kernel's own arithmetic, the samples that fall inside the grown brick; each sample
// adds c * w to those of its 8 corners the brick owns.  Every voxel is accumulated by exactly one brick and written
// once.  What makes plain read-modify-write safe inside the warp: the 32 lanes work on rays that are S pixels apart
// in both detector directions ("colour classes", S^2 passes per pose), and S is chosen per brick and pose such that
// two such rays stay more than two voxels apart (L-infinity) wherever they cross the brick -- their 2x2x2 corner
// sets can never overlap, so no two lanes ever touch the same word; passes are separated by __syncwarp().
// Work per sample ~90 instructions instead of the ~2000 the voxel-centric gather spends per sample it finds.
constexpr int VG_B = 16;  // brick edge

__global__ void __launch_bounds__(32) volume_grad_brick_kernel(const VolGradParams p) {
  __shared__ float acc[VG_B * VG_B * VG_B];
  const int lane = threadIdx.x;
  const int nb1 = (p.D1 + VG_B - 1) / VG_B, nb2 = (p.D2 + VG_B - 1) / VG_B;
  const int bx = blockIdx.x / (nb1 * nb2), by = (blockIdx.x / nb2) % nb1, bz = blockIdx.x % nb2;
  const int lo[3] = {bx * VG_B, by * VG_B, bz * VG_B};
  for (int i = lane; i < VG_B * VG_B * VG_B; i += 32) acc[i] = 0.f;
  __syncwarp();

  const int N = p.H * p.W;
  const int np = p.n_points;
  const float lstep = 1.0f / (float)(np - 1);
  const float inv_vx = 1.0f / p.geom.v[0], inv_uy = 1.0f / p.geom.u[1];
  const float sdd = p.geom.o[2];
  // open box of sample positions that can reach an owned voxel: (lo - 1, lo + B) per axis
  const float elo[3] = {(float)lo[0] - 1.f, (float)lo[1] - 1.f, (float)lo[2] - 1.f};
  const float ehi[3] = {(float)(lo[0] + VG_B), (float)(lo[1] + VG_B), (float)(lo[2] + VG_B)};

  for (int b = 0; b < p.B; ++b) {
    // ---- pixel window of the grown brick and its nearest depth: lanes 0..7 take one corner each
    const float* Gi = p.vox2cam + b * 12;
    const float* G = p.geom.cam2vox + b * 12;
    float cj = 0.f, ci = 0.f, depth = 0.f;
    {
      const int c = lane & 7;
      const float px = (c & 1) ? ehi[0] : elo[0], py = (c & 2) ? ehi[1] : elo[1], pz = (c & 4) ? ehi[2] : elo[2];
      float q[3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
        q[a] = fmaf(__ldg(Gi + a * 4 + 2), pz, fmaf(__ldg(Gi + a * 4 + 1), py, fmaf(__ldg(Gi + a * 4), px, __ldg(Gi + a * 4 + 3))));
      depth = q[2];
      const float m = sdd / q[2];
      cj = (q[0] * m - p.geom.o[0]) * inv_vx;
      ci = (q[1] * m - p.geom.o[1]) * inv_uy;
    }
    float jmin = cj, jmax = cj, imin = ci, imax = ci, dmin = depth;
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      jmin = fminf(jmin, __shfl_xor_sync(0xffffffffu, jmin, o));
      jmax = fmaxf(jmax, __shfl_xor_sync(0xffffffffu, jmax, o));
      imin = fminf(imin, __shfl_xor_sync(0xffffffffu, imin, o));
      imax = fmaxf(imax, __shfl_xor_sync(0xffffffffu, imax, o));
      dmin = fminf(dmin, __shfl_xor_sync(0xffffffffu, dmin, o));
    }
    int j0, j1, i0, i1, S;
    if (dmin > 1e-3f * sdd) {
      j0 = max(0, (int)floorf(jmin));
      j1 = min(p.W - 1, (int)ceilf(jmax));
      i0 = max(0, (int)floorf(imin));
      i1 = min(p.H - 1, (int)ceilf(imax));
      // Two rays S pixels apart: at depth alpha * sdd their world-space distance is >= S * alpha * pixel * cos^2(tilt)
      // (distance between two lines through the source; tilt = largest angle between a ray and the detector normal).
      // In voxel units that is >= ... / (largest voxel spacing), and the L-infinity norm is >= Euclid / sqrt(3).
      float sp2 = 0.f;  // largest squared column norm of vox2cam's 3x3 block = (largest voxel spacing in mm)^2
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float c0 = __ldg(Gi + a), c1 = __ldg(Gi + 4 + a), c2 = __ldg(Gi + 8 + a);
        sp2 = fmaxf(sp2, fmaf(c0, c0, fmaf(c1, c1, c2 * c2)));
      }
      const float xmax = fmaxf(fabsf(p.geom.o[0]), fabsf(p.geom.o[0] + p.geom.v[0] * (float)(p.W - 1)));
      const float ymax = fmaxf(fabsf(p.geom.o[1]), fabsf(p.geom.o[1] + p.geom.u[1] * (float)(p.H - 1)));
      const float cos2 = sdd * sdd / (sdd * sdd + xmax * xmax + ymax * ymax);
      const float pixel = fminf(fabsf(p.geom.u[1]), fabsf(p.geom.v[0]));
      const float sep = cos2 * (dmin / sdd) * pixel * rsqrtf(sp2) * 0.57735027f;  // per pixel of separation
      S = (int)ceilf(2.05f / fmaxf(sep, 1e-6f));
      S = max(1, min(S, 1 << 14));
    } else {  // the grown brick reaches behind the source: every ray may hit it, one ray per pass
      j0 = 0; j1 = p.W - 1; i0 = 0; i1 = p.H - 1; S = 1 << 14;
    }
    if (j0 > j1 || i0 > i1) continue;
    const float s0 = __ldg(G + 3), s1 = __ldg(G + 7), s2 = __ldg(G + 11);
    const float4* inf = p.info + (int64_t)b * N * 3;
    const int Sj = min(S, j1 - j0 + 1), Si = min(S, i1 - i0 + 1);  // colour classes that exist in this window

    for (int cls = 0; cls < Si * Sj; ++cls) {
      const int ci0 = i0 + cls / Sj, cj0 = j0 + cls % Sj;
      const int na = (i1 - ci0) / S + 1, nbj = (j1 - cj0) / S + 1;  // rays of this class: na x nbj, S pixels apart
      for (int t0 = 0; t0 < na * nbj; t0 += 32) {
        const int t = t0 + lane;
        int klo = 0, khi = -1;
        float amin = 0.f, span = 0.f, coef = 0.f, d0 = 0.f, d1 = 0.f, d2 = 0.f;
        if (t < na * nbj) {
          const int i = ci0 + (t / nbj) * S, j = cj0 + (t % nbj) * S;
          const float4* q4 = inf + (int64_t)(i * p.W + j) * 3;
          const float4 a = __ldg(q4);
          if (a.z != 0.f) {
            const float4 d = __ldg(q4 + 1);
            amin = a.x; span = a.y; coef = a.z; d0 = d.x; d1 = d.y; d2 = d.z;
            // alpha window in which the ray is inside the grown brick.  span < 0 is legitimate: a ray that misses
            // the box [0, D-1] but passes within the zero padding is marched "backwards" by the reference (and by the
            // forward kernel), with a negative step weight
            const float aend = a.x + a.y;
            float alo = fminf(a.x, aend), ahi = fmaxf(a.x, aend);
            const float sv[3] = {s0, s1, s2}, dv3[3] = {d.x, d.y, d.z};
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) {
              if (fabsf(dv3[ax]) > 1e-12f) {
                const float r = 1.0f / dv3[ax];
                const float x0 = (elo[ax] - sv[ax]) * r, x1 = (ehi[ax] - sv[ax]) * r;
                alo = fmaxf(alo, fminf(x0, x1));
                ahi = fminf(ahi, fmaxf(x0, x1));
              } else if (!(sv[ax] > elo[ax] && sv[ax] < ehi[ax])) {
                ahi = -INFINITY;
              }
            }
            if (alo <= ahi && a.y != 0.f) {
              // alpha_k = amin + u_k span, u_k ~ k / (n - 1); pad by one sample, the ownership test decides
              const float sc = (float)(np - 1) / a.y;
              const float k0f = (alo - a.x) * sc, k1f = (ahi - a.x) * sc;
              klo = max(0, (int)floorf(fminf(k0f, k1f)) - 1);
              khi = min(np - 1, (int)ceilf(fmaxf(k0f, k1f)) + 1);
            }
          }
        }
        for (int k = klo; k <= khi; ++k) {
          const float u = linspace01(k, np, lstep);
          const float alpha = fmaf(u, span, amin);
          const float x = fmaf(alpha, d0, s0), y = fmaf(alpha, d1, s1), z = fmaf(alpha, d2, s2);
          const float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
          const int vx = (int)fx0 - lo[0], vy = (int)fy0 - lo[1], vz = (int)fz0 - lo[2];  // brick-local corner 000
          if (vx < -1 || vx >= VG_B || vy < -1 || vy >= VG_B || vz < -1 || vz >= VG_B) continue;
          const float fx = x - fx0, fy = y - fy0, fz = z - fz0;
          const float wx[2] = {1.f - fx, fx}, wy[2] = {1.f - fy, fy}, wz[2] = {1.f - fz, fz};
#pragma unroll
          for (int ox = 0; ox < 2; ++ox) {
            const int ax_ = vx + ox;
            if ((unsigned)ax_ >= (unsigned)VG_B || lo[0] + ax_ >= p.D0) continue;
#pragma unroll
            for (int oy = 0; oy < 2; ++oy) {
              const int ay_ = vy + oy;
              if ((unsigned)ay_ >= (unsigned)VG_B || lo[1] + ay_ >= p.D1) continue;
              const float wxy = coef * wx[ox] * wy[oy];
#pragma unroll
              for (int oz = 0; oz < 2; ++oz) {
                const int az_ = vz + oz;
                if ((unsigned)az_ >= (unsigned)VG_B || lo[2] + az_ >= p.D2) continue;
                float* cell = acc + (ax_ * VG_B + ay_) * VG_B + az_;
                *cell = fmaf(wxy, wz[oz], *cell);
              }
            }
          }
        }
        __syncwarp();
      }
    }
  }
  __syncwarp();
  for (int i = lane; i < VG_B * VG_B * VG_B; i += 32) {
    const int vx = i / (VG_B * VG_B), vy = (i / VG_B) % VG_B, vz = i % VG_B;
    const int gx = lo[0] + vx, gy = lo[1] + vy, gz = lo[2] + vz;
    if (gx < p.D0 && gy < p.D1 && gz < p.D2) {
      const int64_t o = ((int64_t)gx * p.D1 + gy) * p.D2 + gz;
      p.gvol[o] = p.accumulate ? p.gvol[o] + acc[i] : acc[i];
    }
  }
}

}  // namespace xvr

using namespace xvr;

// gvol (D0,D1,D2) (+)= dL/dvolume of xvr_trilinear_drr_fwd for upstream gradient gout (B,1,H*W).
// vox2cam (B,3,4) is the inverse of cam2vox; workspace holds 3*B*H*W float4.  The detector basis must be axis
// aligned in the camera frame (row step along y, column step along x), as DiffDRR's detector is.
extern "C" int xvr_trilinear_drr_bwd_volume(const float* cam2vox, const float* vox2cam, const float* cam2world,
                                            const float* det9, int B, int det_h, int det_w, int n_points,
                                            int step_mode, float eps, const float* gout, int D0, int D1, int D2,
                                            float* workspace, float* gvol, int accumulate, int opts, void* stream) {
  if (!cam2vox || !vox2cam || !cam2world || !det9 || !gout || !workspace || !gvol || B <= 0 || det_h <= 0 ||
      det_w <= 0 || n_points < 2 || D0 < 2 || D1 < 2 || D2 < 2 || (opts & ~XVR_OPT_KNOWN)) {
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
  if (!(opts & XVR_OPT_VOLGRAD_GATHER)) {
    const int64_t bricks = (int64_t)((D0 + VG_B - 1) / VG_B) * ((D1 + VG_B - 1) / VG_B) * ((D2 + VG_B - 1) / VG_B);
    volume_grad_brick_kernel<<<(unsigned)bricks, 32, 0, st>>>(p);
    return check_launch("xvr_trilinear_drr_bwd_volume/brick");
  }
  dim3 grid((D2 + 63) / 64, (D1 + 3) / 4, D0);
  volume_grad_kernel<<<grid, 256, 0, st>>>(p);
  return check_launch("xvr_trilinear_drr_bwd_volume");
}
