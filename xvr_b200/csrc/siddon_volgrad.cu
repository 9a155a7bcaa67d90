// dL/dvolume of the Siddon renderer.
//
// The reference gets it from autograd through grid_sample(mode="nearest"): every segment of every ray scatters
// g * L * (alpha_{m+1} - alpha_m) into the voxel its midpoint resolves to, with atomicAdd (non-deterministic).  Two
// formulations here:
//   * fused DRR path (xvr_siddon_drr_bwd_volume): brick-local accumulation, no atomics, deterministic.  A warp owns a
//     16^3 brick of the gradient in shared memory; per pose it projects the brick (grown by one voxel) onto the
//     detector, and the rays of that pixel window walk ONLY the part of their traversal that lies in the grown brick
//     -- the same plane crossings, midpoints and certified voxel indices the forward kernel computes (setup_ray with
//     a clip range keeps a contiguous run of the ray's sorted crossing list).  Lanes work on rays S pixels apart
//     ("colour classes"), S chosen so that two such rays stay more than a voxel diagonal apart inside the brick and
//     therefore never meet in a voxel: plain read-modify-write, every voxel written once by its owner.
//   * ray entry point (source / target tensors, label channels: xvr_siddon_rays_bwd with gvol): no detector geometry
//     to derive ownership from, so it is the reference's own formulation -- one RED.ADD per segment (siddon.cu).
//
// COMPILED WITH -fmad=false like siddon.cu (see siddon_common.cuh).
#include "siddon_common.cuh"

namespace xvr {

constexpr int SVG_B = 16;  // brick edge

struct SiddonVolGradParams {
  SiddonParams sp;                    // fused geometry, volume shape, voxel shift, eps, index tolerance
  const float* __restrict__ vox2cam;  // (B,3,4) inverse of cam2vox
  const float* __restrict__ gout;     // (B,N)
  float* __restrict__ gvol;           // (D0,D1,D2)
  int H, W;
  int accumulate;
};

template <bool HALF>
__global__ void __launch_bounds__(32) siddon_volume_grad_brick_kernel(const SiddonVolGradParams q) {
  __shared__ float acc[SVG_B * SVG_B * SVG_B];
  const SiddonParams& p = q.sp;
  const int lane = threadIdx.x;
  const int D0 = p.vol.D0, D1 = p.vol.D1, D2 = p.vol.D2;
  const int nb1 = (D1 + SVG_B - 1) / SVG_B, nb2 = (D2 + SVG_B - 1) / SVG_B;
  const int bx = blockIdx.x / (nb1 * nb2), by = (blockIdx.x / nb2) % nb1, bz = blockIdx.x % nb2;
  const int lo[3] = {bx * SVG_B, by * SVG_B, bz * SVG_B};
  for (int i = lane; i < SVG_B * SVG_B * SVG_B; i += 32) acc[i] = 0.f;
  __syncwarp();

  const DetectorGeom& geom = p.geom;
  const float inv_vx = 1.0f / geom.v[0], inv_uy = 1.0f / geom.u[1];
  const float sdd = geom.o[2];
  // voxel i covers [i - shift, i + 1 - shift] along its axis; the brick grown by one voxel on every side
  const float elo[3] = {(float)lo[0] - p.voxel_shift - 1.f, (float)lo[1] - p.voxel_shift - 1.f,
                        (float)lo[2] - p.voxel_shift - 1.f};
  const float ehi[3] = {(float)(lo[0] + SVG_B) - p.voxel_shift + 1.f, (float)(lo[1] + SVG_B) - p.voxel_shift + 1.f,
                        (float)(lo[2] + SVG_B) - p.voxel_shift + 1.f};
  const IndexConsts kc = index_consts(p);
  const int N = p.N;

  for (int b = 0; b < p.B; ++b) {
    // ---- pixel window of the grown brick and its nearest depth: lanes 0..7 take one corner each
    const float* Gi = q.vox2cam + b * 12;
    float cj = 0.f, ci = 0.f, depth = 0.f;
    {
      const int c = lane & 7;
      const float px = (c & 1) ? ehi[0] : elo[0], py = (c & 2) ? ehi[1] : elo[1], pz = (c & 4) ? ehi[2] : elo[2];
      float v[3];
#pragma unroll
      for (int a = 0; a < 3; ++a)
        v[a] = fmaf(__ldg(Gi + a * 4 + 2), pz, fmaf(__ldg(Gi + a * 4 + 1), py, fmaf(__ldg(Gi + a * 4), px, __ldg(Gi + a * 4 + 3))));
      depth = v[2];
      const float m = sdd / v[2];
      cj = (v[0] * m - geom.o[0]) * inv_vx;
      ci = (v[1] * m - geom.o[1]) * inv_uy;
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
      j1 = min(q.W - 1, (int)ceilf(jmax));
      i0 = max(0, (int)floorf(imin));
      i1 = min(q.H - 1, (int)ceilf(imax));
      // Two rays S pixels apart are, at depth alpha * sdd, >= S * alpha * pixel * cos^2(tilt) apart in world space
      // (volgrad.cu derives the bound); in voxel units / (largest voxel spacing).  Two rays can only meet in a voxel
      // where they are closer than its diagonal sqrt(3): require S * separation-per-pixel > sqrt(3), with 5 % margin.
      float sp2 = 0.f;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float c0 = __ldg(Gi + a), c1 = __ldg(Gi + 4 + a), c2 = __ldg(Gi + 8 + a);
        sp2 = fmaxf(sp2, fmaf(c0, c0, fmaf(c1, c1, c2 * c2)));
      }
      const float xmax = fmaxf(fabsf(geom.o[0]), fabsf(geom.o[0] + geom.v[0] * (float)(q.W - 1)));
      const float ymax = fmaxf(fabsf(geom.o[1]), fabsf(geom.o[1] + geom.u[1] * (float)(q.H - 1)));
      const float cos2 = sdd * sdd / (sdd * sdd + xmax * xmax + ymax * ymax);
      const float pixel = fminf(fabsf(geom.u[1]), fabsf(geom.v[0]));
      const float sep = cos2 * (dmin / sdd) * pixel * rsqrtf(sp2);  // voxels of separation per pixel
      S = (int)ceilf(1.05f * 1.7320508f / fmaxf(sep, 1e-6f));
      S = max(1, min(S, 1 << 14));
    } else {  // the grown brick reaches behind the source: every ray may hit it, one ray per pass
      j0 = 0; j1 = q.W - 1; i0 = 0; i1 = q.H - 1; S = 1 << 14;
    }
    if (j0 > j1 || i0 > i1) continue;
    const int Sj = min(S, j1 - j0 + 1), Si = min(S, i1 - i0 + 1);

    for (int cls = 0; cls < Si * Sj; ++cls) {
      const int ci0 = i0 + cls / Sj, cj0 = j0 + cls % Sj;
      const int na = (i1 - ci0) / S + 1, nbj = (j1 - cj0) / S + 1;
      for (int t0 = 0; t0 < na * nbj; t0 += 32) {
        const int t = t0 + lane;
        if (t < na * nbj) {
          const int i = ci0 + (t / nbj) * S, j = cj0 + (t % nbj) * S;
          const int64_t ray = (int64_t)b * N + (i * q.W + j);
          const float g = __ldg(q.gout + ray);
          if (g != 0.f) {
            // alpha range in which the ray is inside the grown brick
            float s[3], d[3], L;
            generate_ray(geom, b, i * q.W + j, p.eps, s, d, L);
            float clo = -INFINITY, chi = INFINITY;
#pragma unroll
            for (int a = 0; a < 3; ++a) {
              if (fabsf(d[a]) > 1e-12f) {
                const float x0 = (elo[a] - s[a]) / d[a], x1 = (ehi[a] - s[a]) / d[a];
                clo = fmaxf(clo, fminf(x0, x1));
                chi = fminf(chi, fmaxf(x0, x1));
              } else if (!(s[a] > elo[a] && s[a] < ehi[a])) {
                chi = -INFINITY;
              }
            }
            if (clo <= chi) {
              RaySetup r;
              setup_ray<HALF>(p, b, ray, r, clo, chi);  // re-generates the same s, d, L
              const float coef = g * r.L;
              float prev;
              if (!r.empty && pop_next<HALF>(p, r, prev) >= 0) {
                for (;;) {
                  float next;
                  if (pop_next<HALF>(p, r, next) < 0) break;
                  const float mid = __fmul_rn(__fadd_rn(prev, next), 0.5f);
                  int idx[3];
                  midpoint_voxel_axes(p, kc, mid, r.s, r.d, r.tol, idx);
                  const int vx = idx[0] - lo[0], vy = idx[1] - lo[1], vz = idx[2] - lo[2];
                  if ((idx[0] | idx[1] | idx[2]) >= 0 && (unsigned)vx < (unsigned)SVG_B &&
                      (unsigned)vy < (unsigned)SVG_B && (unsigned)vz < (unsigned)SVG_B) {
                    float* cell = acc + (vx * SVG_B + vy) * SVG_B + vz;
                    *cell = *cell + coef * __fsub_rn(next, prev);
                  }
                  prev = next;
                }
              }
            }
          }
        }
        __syncwarp();
      }
    }
  }
  __syncwarp();
  for (int i = lane; i < SVG_B * SVG_B * SVG_B; i += 32) {
    const int vx = i / (SVG_B * SVG_B), vy = (i / SVG_B) % SVG_B, vz = i % SVG_B;
    const int gx = lo[0] + vx, gy = lo[1] + vy, gz = lo[2] + vz;
    if (gx < D0 && gy < D1 && gz < D2) {
      const int64_t o = ((int64_t)gx * D1 + gy) * D2 + gz;
      q.gvol[o] = q.accumulate ? q.gvol[o] + acc[i] : acc[i];
    }
  }
}

}  // namespace xvr

using namespace xvr;

// gvol (D0,D1,D2) (+)= dL/dvolume of xvr_siddon_drr_fwd for upstream gradient gout (B,1,H*W); atomics-free and
// deterministic.  vox2cam (B,3,4) is the inverse of cam2vox.  The detector basis must be axis aligned in the camera
// frame (row step along y, column step along x), as DiffDRR's detector is.
extern "C" int xvr_siddon_drr_bwd_volume(const float* cam2vox, const float* vox2cam, const float* cam2world,
                                         const float* det9, int B, int det_h, int det_w, float voxel_shift, float eps,
                                         const float* gout, int D0, int D1, int D2, float* gvol, int accumulate,
                                         int opts, void* stream) {
  if (!cam2vox || !vox2cam || !cam2world || !det9 || !gout || !gvol || B <= 0 || det_h <= 0 || det_w <= 0 || D0 < 1 ||
      D1 < 1 || D2 < 1 || (opts & ~XVR_OPT_KNOWN) || ((opts >> XVR_OPT_SIDDON_TOL_SHIFT) & 0xF) > 4) {
    set_last_error("xvr_siddon_drr_bwd_volume: invalid argument");
    return XVR_ERR_INVALID;
  }
  if (det9[3] != 0.f || det9[5] != 0.f || det9[7] != 0.f || det9[8] != 0.f || det9[4] == 0.f || det9[6] == 0.f) {
    set_last_error("xvr_siddon_drr_bwd_volume: detector basis must be axis aligned (row step = (0,dy,0), "
                   "column step = (dx,0,0))");
    return XVR_ERR_INVALID;
  }
  if ((int64_t)D0 * D1 * D2 >= (int64_t)1 << 31) {
    set_last_error("xvr_siddon_drr_bwd_volume: volume too large for 32-bit voxel offsets");
    return XVR_ERR_INVALID;
  }
  static const float kScale[5] = {1.0f, 1e30f, 0.5f, 0.25f, 0.125f};
  SiddonVolGradParams q = {};
  SiddonParams& p = q.sp;
  p.vol.data = gvol;  // never read: the adjoint does not depend on the voxel values
  p.vol.D0 = D0;
  p.vol.D1 = D1;
  p.vol.D2 = D2;
  p.vol.s0 = D1 * D2;
  p.vol.s1 = D2;
  p.C = 1;
  p.fused = true;
  p.geom.cam2vox = cam2vox;
  p.geom.cam2world = cam2world;
  for (int a = 0; a < 3; ++a) {
    p.geom.o[a] = det9[a];
    p.geom.u[a] = det9[3 + a];
    p.geom.v[a] = det9[6 + a];
  }
  p.geom.W = det_w;
  p.B = B;
  p.N = det_h * det_w;
  p.voxel_shift = voxel_shift;
  p.eps = eps;
  p.index_tol_scale = kScale[(opts >> XVR_OPT_SIDDON_TOL_SHIFT) & 0xF];
  p.idx_bias = (int)(0u - 0x4B400000u * (unsigned)(p.vol.s0 + p.vol.s1 + 1));
  {
    const int size[3] = {D0, D1, D2};
    for (int a = 0; a < 3; ++a) p.rsize[a] = 1.0f / (float)size[a];
  }
  q.vox2cam = vox2cam;
  q.gout = gout;
  q.gvol = gvol;
  q.H = det_h;
  q.W = det_w;
  q.accumulate = accumulate;
  const int64_t bricks = (int64_t)((D0 + SVG_B - 1) / SVG_B) * ((D1 + SVG_B - 1) / SVG_B) * ((D2 + SVG_B - 1) / SVG_B);
  const bool half = fabsf(voxel_shift) <= 4.f && 2.f * voxel_shift == floorf(2.f * voxel_shift);
  cudaStream_t st = (cudaStream_t)stream;
  if (half) siddon_volume_grad_brick_kernel<true><<<(unsigned)bricks, 32, 0, st>>>(q);
  else siddon_volume_grad_brick_kernel<false><<<(unsigned)bricks, 32, 0, st>>>(q);
  return check_launch("xvr_siddon_drr_bwd_volume");
}
