// Shared device helpers for the xvr_b200 DRR hot path (sm_100a).
//
// Coordinates: the CT volume is a dense fp32 array vol[D0][D1][D2] (D2 contiguous). Ray end points are
// given in voxel-index coordinates (what xvr obtains from drr.affine_inverse at
// /root/reference/src/xvr/model/trainer.py:285), so component a of a point indexes volume axis a.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define XVR_OK 0
#define XVR_ERR_INVALID -1
#define XVR_ERR_CUDA -2

// Per-call options word `opts` of the entry points that have variants (include/xvr_b200.h XVR_OPT_*; 0 = default).
#define XVR_OPT_KSPLIT_MASK 0x7        // 0: automatic, 1..4: 1/2/4/8 lanes share one ray (trilinear forward)
#define XVR_OPT_SIDDON_WALK 0x10       // Siddon: voxel indices from the integer walk where its certificate holds
#define XVR_OPT_VOLGRAD_GATHER 0x20    // dL/dvolume: voxel-centric gather instead of the brick-local scatter
#define XVR_OPT_NO_TRIM 0x40           // trilinear: march all n_points samples even outside the box of non-zero voxels
#define XVR_OPT_SIDDON_TOL_SHIFT 8     // bits 8..11, test hook: certificate tolerance 0: x1, 1: always exact, 2..4: x1/2, 1/4, 1/8
#define XVR_OPT_LABEL_BRICKS 0x80      // trilinear forward with labels: a brick table follows the label volume (see header)
#define XVR_OPT_KNOWN 0xFF7

namespace xvr {

void set_last_error(const char* msg);
int check_launch(const char* what);

struct Vol {
  const float* __restrict__ data;
  int D0, D1, D2;
  int s0, s1;  // element strides of axis 0 / 1 (D1*D2, D2)
  cudaTextureObject_t tex;  // optional layered-2D copy: layer = axis 0, height = axis 1, width = axis 2
  const int* __restrict__ bbox;  // optional DEVICE int[6]: first / last index of a non-zero voxel per axis (lo0 lo1 lo2 hi0 hi1 hi2)
  const uint8_t* __restrict__ occ;  // optional (nb0,nb1,nb2) distance field over OCC_BRICK^3 bricks: Chebyshev distance in
  int nb0, nb1, nb2;                // bricks to the nearest OCCUPIED one (grown by two voxels, it holds a non-zero voxel)
};

#ifndef XVR_OCC_BRICK
#define XVR_OCC_BRICK 8
#endif
constexpr int OCC_BRICK = XVR_OCC_BRICK;  // even (the occupancy is built from flags of 2^3-voxel cells)
constexpr int OCC_DIST_CAP = 24;          // largest distance the field stores (bricks)

// Opaque handle behind xvr_volume_* (include/xvr_b200.h): a block-linear layered array + point-sampled texture.
// The array has D0 + 2 layers: volume layer x lives in array layer x + 1, array layers 0 and D0 + 1 are zero, so
// that grid_sample's zero padding along axis 0 is a clamped layer index instead of a predicated fetch.
struct VolumeTexture {
  cudaArray_t array;
  cudaTextureObject_t tex;
  int D0, D1, D2;
  int* bbox;  // DEVICE int[6], refreshed by every upload: the box of the volume's non-zero voxels (empty: lo = D, hi = -1)
  uint8_t* occ;  // DEVICE (nb0,nb1,nb2) brick distance field (0 = occupied), refreshed by every upload
  int nb0, nb1, nb2;
  uint8_t* occ_tmp;  // scratch of the distance transform, same size
  uint8_t* cells;    // scratch: any-non-zero flags of 2^3-voxel cells
};

// The 2x2 (axis1, axis2) footprint around texel-corner (u, v) of one layer in a single TEX instruction.
// With u = iz + 1, v = iy + 1 the request sits exactly between four texel centres, so the footprint selection is
// not subject to the sampler's fixed-point coordinate rounding; the returned values are the stored fp32 texels
// (.x = (iy+1, iz), .y = (iy+1, iz+1), .z = (iy, iz+1), .w = (iy, iz)); out-of-range texels read 0 (border mode).
__device__ __forceinline__ float4 gather_yz(cudaTextureObject_t tex, int layer, float u, float v) {
  float4 r;
  asm volatile("tld4.r.a2d.v4.f32.f32 {%0,%1,%2,%3}, [%4, {%5,%6,%7,%7}];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(tex), "r"(layer), "f"(u), "f"(v));
  return r;
}

// In-kernel ray generation (the fused DRR path): the detector point of pixel (i, j) in the camera frame is
// c = o + i*u + j*v; its voxel-space target is G c, the source is G's translation column and the world-mm ray
// length is |M_rot c| (DiffDRR detector.py + drr.py as sequenced at /root/reference/src/xvr/model/trainer.py:283-285).
struct DetectorGeom {
  const float* __restrict__ cam2vox;    // (B,3,4) row-major, camera -> voxel-index coordinates
  const float* __restrict__ cam2world;  // (B,3,4) row-major, camera -> world mm (rotation part gives the length)
  float o[3], u[3], v[3];
  int W;
};

__device__ __forceinline__ void camera_point(const DetectorGeom& g, int n, float c[3]) {
  const int i = n / g.W, j = n - i * g.W;
#pragma unroll
  for (int a = 0; a < 3; ++a) c[a] = fmaf((float)j, g.v[a], fmaf((float)i, g.u[a], g.o[a]));
}

// source s, direction d = t - s + eps (voxel coords) and world ray length L of ray n of pose b
__device__ __forceinline__ void generate_ray(const DetectorGeom& g, int b, int n, float eps, float s[3], float d[3],
                                             float& L) {
  float c[3];
  camera_point(g, n, c);
  const float* G = g.cam2vox + b * 12;
  const float* M = g.cam2world + b * 12;
  float w[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    s[a] = __ldg(G + a * 4 + 3);
    const float t = fmaf(__ldg(G + a * 4 + 2), c[2], fmaf(__ldg(G + a * 4 + 1), c[1], fmaf(__ldg(G + a * 4), c[0], s[a])));
    d[a] = (t - s[a]) + eps;
    w[a] = fmaf(__ldg(M + a * 4 + 2), c[2], fmaf(__ldg(M + a * 4 + 1), c[1], __ldg(M + a * 4) * c[0]));
  }
  L = sqrtf(fmaf(w[0], w[0], fmaf(w[1], w[1], w[2] * w[2])));
}

// Slab test of the segment s + alpha*d, alpha in [0,1], against the box [lo, hi]^3 (per axis).
// Restates DiffDRR renderers._get_alpha_minmax: per-axis (plane - s)/d with IEEE division, min/max over the
// two planes, max/min over axes, then the clamp to [0,1].  Also reports which plane is active for the
// analytic pose gradient (axis index, plane coordinate; axis = -1 when the clamp is active).
struct AlphaRange {
  float amin, amax;
  int axis_min, axis_max;      // active axis, or -1 if clamped
  float plane_min, plane_max;  // plane coordinate of the active crossing
};

__device__ __forceinline__ AlphaRange alpha_range(const float s[3], const float d[3], const float lo[3],
                                                  const float hi[3]) {
  AlphaRange r;
  r.amin = -INFINITY;
  r.amax = INFINITY;
  r.axis_min = -1;
  r.axis_max = -1;
  r.plane_min = 0.f;
  r.plane_max = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float a0 = __fdiv_rn(lo[a] - s[a], d[a]);
    float a1 = __fdiv_rn(hi[a] - s[a], d[a]);
    float mn, mx, pmn, pmx;
    if (a0 <= a1) { mn = a0; mx = a1; pmn = lo[a]; pmx = hi[a]; }
    else          { mn = a1; mx = a0; pmn = hi[a]; pmx = lo[a]; }
    if (mn > r.amin) { r.amin = mn; r.axis_min = a; r.plane_min = pmn; }
    if (mx < r.amax) { r.amax = mx; r.axis_max = a; r.plane_max = pmx; }
  }
  if (r.amin < 0.f) { r.amin = 0.f; r.axis_min = -1; }
  if (r.amax > 1.f) { r.amax = 1.f; r.axis_max = -1; }
  return r;
}

// Walk through the brick distance field along s + t d from t0 towards t1: (a lower bound within one probe offset of) the
// t at which the ray enters the first occupied brick, t0 if it starts in one, +inf if there is none.
// In a brick at Chebyshev distance D >= 1 from the nearest occupied brick the ray may advance (D - 1) bricks along its
// fastest axis -- every brick it can reach is free -- and in any case to the exit of the brick it is in; the next brick
// is then probed a hundredth of a brick further on (a brick cut by less than that is covered by the two voxels its
// neighbours' occupancy is grown by).  A handful of steps through a wide margin of air, one per brick next to the body.
// If the step budget runs out the walk reports the point it reached: conservative (less is trimmed), never wrong.
__device__ __forceinline__ float first_occupied_brick(const Vol& v, const float s[3], const float d[3], float t0,
                                                      float t1) {
  constexpr float BK = (float)OCC_BRICK, INV = 1.0f / (float)OCC_BRICK;
  const int nb[3] = {v.nb0, v.nb1, v.nb2};
  float invd[3], off[3], up[3];
  float dmax = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const bool moving = d[a] != 0.f;
    invd[a] = moving ? 1.0f / d[a] : 0.f;
    off[a] = moving ? 0.f : INFINITY;  // an axis the ray does not move along never ends a brick
    up[a] = d[a] > 0.f ? 1.f : 0.f;
    dmax = fmaxf(dmax, fabsf(d[a]));
  }
  if (!(dmax > 0.f)) return t0;
  const float jump = BK / dmax, probe = 0.01f * jump;
  float t = t0, entered = t0;
  for (int it = 0; it < 4 * OCC_DIST_CAP + 64; ++it) {
    float cf[3];
    int c[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      cf[a] = floorf(fmaf(t, d[a], s[a]) * INV);
      c[a] = min(max((int)cf[a], 0), nb[a] - 1);
    }
    const int D = __ldg(v.occ + ((int64_t)c[0] * nb[1] + c[1]) * nb[2] + c[2]);
    if (D == 0) return entered;
    float t_exit = INFINITY;
#pragma unroll
    for (int a = 0; a < 3; ++a) t_exit = fminf(t_exit, fmaf((cf[a] + up[a]) * BK - s[a], invd[a], off[a]));
    entered = fmaxf(t_exit, t);  // (a position outside the grid is looked up in the nearest brick: keep moving on)
    t = fmaxf(entered + probe, fmaf((float)(D - 1), jump, t));
    if (!(entered < t1)) return INFINITY;
    if (t > t1) t = t1;  // the last stretch ends in the brick of t1: probe that one too
  }
  return entered;
}

// Where along the ray s + alpha d can a NON-ZERO voxel be touched?  Narrows [first, last] (in: the alpha range of
// interest, out: from the entry into the first occupied region to the exit from the last one); false if nowhere.
//   stage 1 -- the box of the volume's non-zero voxels, widened to (lo - 2, hi + 2): a trilinear sample at x reads the
//     voxels floor(x), floor(x) + 1, a Siddon midpoint resolves to a voxel within one of floor(x); the second voxel is
//     for rays that run along a face of the box within rounding (slab test on one side, positions on the other);
//   stage 2 -- the distance field over OCC_BRICK^3 bricks (a brick is occupied if, grown by two voxels, it holds a
//     non-zero voxel: same two reasons), sphere-traced in from BOTH ends up to the first occupied brick each way
//     (first_occupied_brick): a ray through a body surrounded by air takes a handful of steps through the air (one
//     that meets nothing occupied walks its whole length once and is dropped).  Empty bricks BETWEEN occupied ones are
//     not cut out.
// Everything outside [first, last] contributes exact zeros to a line integral and to its derivatives; callers add their
// own margin along the ray for the rounding of the alphas.
__device__ __forceinline__ bool occupied_alpha_range(const Vol& v, const float s[3], const float d[3], float& first,
                                                     float& last) {
  float tin = first, tout = last;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float lo = (float)(__ldg(v.bbox + a) - 2), hi = (float)(__ldg(v.bbox + 3 + a) + 2);
    if (d[a] != 0.f) {
      const float a0 = (lo - s[a]) / d[a], a1 = (hi - s[a]) / d[a];
      tin = fmaxf(tin, fminf(a0, a1));
      tout = fminf(tout, fmaxf(a0, a1));
    } else if (!(s[a] > lo && s[a] < hi)) {
      return false;
    }
  }
  if (!(tin <= tout)) return false;
  if (!v.occ) {
    first = tin;
    last = tout;
    return true;
  }
  const float f = first_occupied_brick(v, s, d, tin, tout);
  if (f == INFINITY) return false;
  const float nd[3] = {-d[0], -d[1], -d[2]};
  const float lb = first_occupied_brick(v, s, nd, -tout, -tin);
  // the brick found above is on the way back too, unless the reversed walk rounds a grazed face the other way
  const float l = lb == INFINITY ? tout : -lb;
  first = fminf(f, l);
  last = fmaxf(f, l);
  return true;
}

// torch.linspace(0, 1, n)[k] in fp32 (ATen RangeFactories: symmetric evaluation around the midpoint).
// The upper half is ONE fused operation, spelled out so that every kernel (texture, staged, backward, volume
// adjoint) forms bit-identical sample parameters whatever the compiler's contraction choices are.
__device__ __forceinline__ float linspace_tail(float step, float r) { return __fmaf_rn(-step, r, 1.0f); }
__device__ __forceinline__ float linspace01(int k, int n, float step) {
  return (k < n / 2) ? step * (float)k : linspace_tail(step, (float)(n - 1 - k));
}

// Trilinear blend of the 8 corner values c[x][y][z] at fractional offsets (fx, fy, fz): along z (contiguous), then
// y, then x.  GRAD also returns dV/dx.  Shared by every gather path (texture, global loads, staged bricks) so that
// they produce bit-identical samples.
template <bool GRAD>
__device__ __forceinline__ float trilinear_interp(float c000, float c001, float c010, float c011, float c100,
                                                  float c101, float c110, float c111, float fx, float fy, float fz,
                                                  float g[3]) {
  float dz00 = c001 - c000, dz01 = c011 - c010, dz10 = c101 - c100, dz11 = c111 - c110;
  float c00 = fmaf(fz, dz00, c000), c01 = fmaf(fz, dz01, c010);
  float c10 = fmaf(fz, dz10, c100), c11 = fmaf(fz, dz11, c110);
  float dy0 = c01 - c00, dy1 = c11 - c10;
  float c0 = fmaf(fy, dy0, c00), c1 = fmaf(fy, dy1, c10);
  float dx = c1 - c0;
  if (GRAD) {
    g[0] = dx;
    g[1] = fmaf(fx, dy1 - dy0, dy0);
    float gz0 = fmaf(fy, dz01 - dz00, dz00), gz1 = fmaf(fy, dz11 - dz10, dz10);
    g[2] = fmaf(fx, gz1 - gz0, gz0);
  }
  return fmaf(fx, dx, c0);
}

// One trilinear sample with zero padding (grid_sample mode="bilinear", padding_mode="zeros",
// align_corners=True on a grid normalised with dims = shape-1, i.e. the sampler coordinate IS the voxel index;
// ATen/native/cuda/GridSampler.cuh:23-31 and the out-of-bounds handling at :225-227).
// GRAD additionally returns the spatial gradient dV/dx (zero-padded corners differentiate as zeros,
// which is what grid_sampler_3d_backward computes).
template <bool GRAD, bool TEX>
__device__ __forceinline__ float sample_trilinear(const Vol& v, float x, float y, float z, float g[3]) {
  float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
  int ix = (int)fx0, iy = (int)fy0, iz = (int)fz0;
  float fx = x - fx0, fy = y - fy0, fz = z - fz0;
  float c000, c001, c010, c011, c100, c101, c110, c111;
  if (TEX) {
    // two unconditional gathers fetch the 8 corners; axis-1/2 padding comes from the border mode, axis-0 padding
    // from the zero layers at both ends of the array (any out-of-range x clamps onto one of them)
    const float tu = fz0 + 1.0f, tv = fy0 + 1.0f;
    const unsigned last = (unsigned)(v.D0 + 1);
    const float4 a = gather_yz(v.tex, (int)min((unsigned)(ix + 1), last), tu, tv);
    const float4 b = gather_yz(v.tex, (int)min((unsigned)(ix + 2), last), tu, tv);
    c010 = a.x; c011 = a.y; c001 = a.z; c000 = a.w;
    c110 = b.x; c111 = b.y; c101 = b.z; c100 = b.w;
  } else if ((unsigned)ix < (unsigned)(v.D0 - 1) && (unsigned)iy < (unsigned)(v.D1 - 1) &&
      (unsigned)iz < (unsigned)(v.D2 - 1)) {
    const float* p = v.data + ((int64_t)ix * v.s0 + iy * v.s1 + iz);
    c000 = __ldg(p);
    c001 = __ldg(p + 1);
    c010 = __ldg(p + v.s1);
    c011 = __ldg(p + v.s1 + 1);
    c100 = __ldg(p + v.s0);
    c101 = __ldg(p + v.s0 + 1);
    c110 = __ldg(p + v.s0 + v.s1);
    c111 = __ldg(p + v.s0 + v.s1 + 1);
  } else {
    // Border / outside: fetch each corner only if it is inside the volume, else 0.
    // Coordinates far outside (|x| huge, NaN) fall through with every corner rejected.
    if (!(x > -1.f && x < (float)v.D0 && y > -1.f && y < (float)v.D1 && z > -1.f && z < (float)v.D2)) {
      if (GRAD) { g[0] = g[1] = g[2] = 0.f; }
      return 0.f;
    }
    bool x0 = ix >= 0, x1 = ix + 1 < v.D0;
    bool y0 = iy >= 0, y1 = iy + 1 < v.D1;
    bool z0 = iz >= 0, z1 = iz + 1 < v.D2;
    const float* p = v.data + ((int64_t)ix * v.s0 + iy * v.s1 + iz);
    c000 = (x0 && y0 && z0) ? __ldg(p) : 0.f;
    c001 = (x0 && y0 && z1) ? __ldg(p + 1) : 0.f;
    c010 = (x0 && y1 && z0) ? __ldg(p + v.s1) : 0.f;
    c011 = (x0 && y1 && z1) ? __ldg(p + v.s1 + 1) : 0.f;
    c100 = (x1 && y0 && z0) ? __ldg(p + v.s0) : 0.f;
    c101 = (x1 && y0 && z1) ? __ldg(p + v.s0 + 1) : 0.f;
    c110 = (x1 && y1 && z0) ? __ldg(p + v.s0 + v.s1) : 0.f;
    c111 = (x1 && y1 && z1) ? __ldg(p + v.s0 + v.s1 + 1) : 0.f;
  }
  return trilinear_interp<GRAD>(c000, c001, c010, c011, c100, c101, c110, c111, fx, fy, fz, g);
}

// Adjoint of sample_trilinear w.r.t. the volume: coef * (trilinear weight) into each of the 8 corners that lie inside
// the volume (zero-padded corners receive nothing) -- grid_sampler_3d_backward's safe_add_3d
// (ATen/native/cuda/GridSampler.cuh:263-280), one RED.ADD.F32 per corner.  Used by the ray entry points, which have
// no detector geometry to derive an atomics-free ownership from (the fused DRR path has: csrc/volgrad.cu).
__device__ __forceinline__ void scatter_trilinear(float* __restrict__ gvol, const Vol& v, float x, float y, float z,
                                                  float coef) {
  if (!(x > -1.f && x < (float)v.D0 && y > -1.f && y < (float)v.D1 && z > -1.f && z < (float)v.D2)) return;
  const float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
  const int ix = (int)fx0, iy = (int)fy0, iz = (int)fz0;
  const float fx = x - fx0, fy = y - fy0, fz = z - fz0;
  const float wx[2] = {1.f - fx, fx}, wy[2] = {1.f - fy, fy}, wz[2] = {1.f - fz, fz};
#pragma unroll
  for (int ox = 0; ox < 2; ++ox) {
    if ((unsigned)(ix + ox) >= (unsigned)v.D0) continue;
#pragma unroll
    for (int oy = 0; oy < 2; ++oy) {
      if ((unsigned)(iy + oy) >= (unsigned)v.D1) continue;
#pragma unroll
      for (int oz = 0; oz < 2; ++oz) {
        if ((unsigned)(iz + oz) >= (unsigned)v.D2) continue;
        atomicAdd(gvol + ((int64_t)(ix + ox) * v.s0 + (iy + oy) * v.s1 + (iz + oz)), coef * wx[ox] * wy[oy] * wz[oz]);
      }
    }
  }
}

// Nearest label lookup at the same sampler coordinate (grid_sample mode="nearest", align_corners=True,
// zero padding -> channel 0 when outside; GridSampler.cuh nearest branch uses nearbyint = ties-to-even).
__device__ __forceinline__ int sample_label(const uint8_t* __restrict__ lab, const Vol& v, float x, float y,
                                            float z) {
  float rx = nearbyintf(x), ry = nearbyintf(y), rz = nearbyintf(z);
  if (!(rx >= 0.f && rx < (float)v.D0 && ry >= 0.f && ry < (float)v.D1 && rz >= 0.f && rz < (float)v.D2))
    return 0;
  return (int)__ldg(lab + ((int64_t)(int)rx * v.s0 + (int)ry * v.s1 + (int)rz));
}

// The same lookup through a table of OCC_BRICK^3-voxel bricks: bricks[b] = the label every voxel of brick b GROWN BY ONE
// VOXEL carries (outside the volume counts as 0), or 255 if they differ.  A position in brick b rounds to a voxel of the
// grown brick, so a uniform brick answers for it -- one byte from a table that stays in L1 (neighbouring lanes read the
// same entry) instead of a scattered byte from a volume-sized array; mixed bricks (label boundaries) take the exact path.
__device__ __forceinline__ int sample_label_bricked(const uint8_t* __restrict__ lab, const uint8_t* __restrict__ bricks,
                                                    const Vol& v, float x, float y, float z) {
  constexpr float INV = 1.0f / (float)OCC_BRICK;
  const int nb0 = (v.D0 + OCC_BRICK - 1) / OCC_BRICK, nb1 = (v.D1 + OCC_BRICK - 1) / OCC_BRICK,
            nb2 = (v.D2 + OCC_BRICK - 1) / OCC_BRICK;
  // (NaN / far-away positions convert to a clamped brick; an edge brick is uniform only if it is all 0, the exact
  // answer for anything outside -- and the exact path rejects them itself)
  const int b0 = min(max((int)floorf(x * INV), 0), nb0 - 1), b1 = min(max((int)floorf(y * INV), 0), nb1 - 1),
            b2 = min(max((int)floorf(z * INV), 0), nb2 - 1);
  const int c = __ldg(bricks + ((int64_t)b0 * nb1 + b1) * nb2 + b2);
  return c != 255 ? c : sample_label(lab, v, x, y, z);
}

// Thread -> detector pixel mapping.  A CTA of 256 threads covers a (256>>cta_w_log2) x (1<<cta_w_log2) pixel
// tile; inside it each warp covers a (32>>lane_w_log2) x (1<<lane_w_log2) sub-tile, so that the 32 lanes of a
// gather touch as few 128-byte lines of the volume as the view geometry allows.
struct TileMap {
  int W, H;         // detector width/height used for the mapping (N = H*W); W == 0 -> linear ray index
  int lane_w_log2;  // log2 of the warp sub-tile width
  int cta_w_log2;   // log2 of the CTA tile width
  int tiles_x, tiles_y;
  int ks_log2;      // log2 of the number of lanes that share one ray (each takes a slice of its samples)
};

// With ks_log2 > 0 a warp holds 32 >> ks_log2 rays; lane l serves ray (l mod rays-per-warp), sample slice
// l / rays-per-warp (small batches: more threads in flight than rays).
__device__ __forceinline__ int tile_ray_index(const TileMap& m, int tile, int tid, int N) {
  const int lane = tid & 31, warp = tid >> 5;
  const int rpw_log2 = 5 - m.ks_log2;
  const int r = lane & ((1 << rpw_log2) - 1);
  if (m.W == 0) {
    const int n = (tile * 8 + warp) * (1 << rpw_log2) + r;
    return n < N ? n : -1;
  }
  const int lw = m.lane_w_log2, cw = m.cta_w_log2;
  const int lane_j = r & ((1 << lw) - 1), lane_i = r >> lw;
  const int wpr_log2 = cw - lw;  // warps per tile row
  const int warp_j = warp & ((1 << wpr_log2) - 1), warp_i = warp >> wpr_log2;
  const int ty = tile / m.tiles_x, tx = tile - ty * m.tiles_x;
  const int j = (tx << cw) + (warp_j << lw) + lane_j;
  const int i = ty * ((256 >> cw) >> m.ks_log2) + warp_i * ((1 << rpw_log2) >> lw) + lane_i;
  return (i < m.H && j < m.W) ? i * m.W + j : -1;
}

// sum over the lanes that share a ray (lanes congruent modulo rays-per-warp)
__device__ __forceinline__ float ksplit_sum(float v, int ks_log2) {
  for (int o = 16; o >= (32 >> ks_log2); o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace xvr
