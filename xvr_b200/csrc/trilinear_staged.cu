// Trilinear DRR forward (+ Jacobian) with the volume staged brick by brick in shared memory -- the variant
// north_star describes ("TMA-staged volume bricks").  OPT-IN and NOT YET RUN ON A GPU (written after round 1's GPU
// budget was spent; DESIGN.md 5.1, profiles/r1_brick_staging_model.md for the model that motivates it).
//
// Same contract as xvr_trilinear_drr_fwd (csrc/trilinear.cu): same rays, same alpha_k, same sample positions, same
// 8-corner blend (trilinear_interp) accumulated in the same order -- only the SOURCE of the 8 corners differs, so
// the output must be bit-identical to the texture kernel's, which is what its tests assert.
//
// A CTA owns a 16 x 16 detector tile of one pose; warp w takes columns 2w, 2w+1 (16 rows x 2 columns per warp: the
// detector rows run along the contiguous volume axis in the AP/PA set-up, so a warp's 8-corner loads spread over the
// banks).  The tile's rays form a thin frustum; the CTA walks it slab by slab (ST_K voxel layers) along the
// frustum's dominant volume axis A.  Per slab:
//   1. warp 0 bounds the slab's corner indices on the other two axes from the four corner rays of the tile (the
//      pixel -> plane-point map is projective, so the tile's image on a plane x_A = const is the convex hull of its
//      corner images) and publishes the box;
//   2. every thread issues the bulk copies (cp.async.bulk, the TMA unit's linear mode) of its rows of the box --
//      a row is the box's extent along the contiguous axis, 16-byte aligned -- and zero-fills what lies outside the
//      volume (= grid_sample's zero padding); completion is tracked by one mbarrier (expect-tx per thread);
//   3. every thread interpolates those of its samples whose cell lies in the slab from 8 shared-memory loads.
// Anything the box cannot serve is served from global memory with the reference arithmetic (sample_trilinear's
// load path): boxes larger than the buffer, samples outside the box (rounding at its faces), rays that run
// against the frustum's direction or only graze the zero padding, volumes whose rows are not 16-byte multiples,
// and -- bounded spin -- a barrier that does not complete.  Nothing here can wait forever.
#include "common.cuh"

namespace xvr {

// Tuning knobs (override with XVR_B200_NVCC_FLAGS="-DXVR_ST_K=8 -DXVR_ST_CAP=24576 -DXVR_ST_MIN_CTAS=2")
#ifndef XVR_ST_K
#define XVR_ST_K 4
#endif
#ifndef XVR_ST_CAP
#define XVR_ST_CAP 12288
#endif
#ifndef XVR_ST_MIN_CTAS
#define XVR_ST_MIN_CTAS 3
#endif
constexpr int ST_T = 16;             // detector tile edge per CTA
constexpr int ST_K = XVR_ST_K;       // voxel layers per slab
constexpr int ST_CAP = XVR_ST_CAP;   // floats in the staging buffer (48 KB -> three to four CTAs per SM)

struct StagedParams {
  Vol vol;
  DetectorGeom geom;
  int B, H, W, n_points, step_mode;
  float eps;
  int tiles_x, tiles_y;
  float* __restrict__ out;          // (B,1,H*W)
  float* __restrict__ jac;          // (B,7,H*W), nullable
  unsigned long long* __restrict__ stats;  // nullable: {samples from shared memory, samples from global memory, barrier time-outs}
};

struct BoxDesc {
  int lo[3];   // first staged voxel index per axis (may be -1 / -4: zero padding)
  int E[3];    // extent per axis; E[2] is a multiple of 4 and lo[2] a multiple of 4 (16-byte rows)
  int staged;  // 0: this slab is served from global memory
};

__device__ __forceinline__ float st_step_weight(int mode, float span, int n) {
  if (mode == 0) return __fdiv_rn(span, (float)(n - 1));
  if (mode == 1) return __fdiv_rn(span, (float)n);
  return __fdiv_rn(1.0f, (float)n);
}
__device__ __forceinline__ float st_step_weight_dspan(int mode, int n) {
  if (mode == 0) return 1.0f / (float)(n - 1);
  if (mode == 1) return 1.0f / (float)n;
  return 0.f;
}

// ---- mbarrier / bulk-copy wrappers (PTX ISA 8.x, sm_90+)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// component `a` of a 3-vector without dynamic indexing (keeps the arrays in registers)
__device__ __forceinline__ float pick3(const float v[3], int a) { return a == 0 ? v[0] : (a == 1 ? v[1] : v[2]); }

// slab number of a cell index (cells -ST_K .. -1 -> 0, 0 .. ST_K-1 -> 1, ...); first cell of slab n is (n-1)*ST_K
__device__ __forceinline__ int slab_of(int cell) { return (max(cell, -ST_K) + ST_K) / ST_K; }

template <bool JAC, int STAGES>
__global__ void __launch_bounds__(256, STAGES == 1 ? XVR_ST_MIN_CTAS : 2) trilinear_fwd_staged_kernel(const StagedParams p) {
  extern __shared__ __align__(16) float box[];  // STAGES * ST_CAP floats
  __shared__ __align__(8) unsigned long long mbar_storage[2];
  __shared__ BoxDesc descs[3];
  __shared__ int t_first, t_last;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles = p.tiles_x * p.tiles_y;
  const int b = blockIdx.x / tiles;
  const int tile = blockIdx.x - b * tiles;
  const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
  const int i0 = ty * ST_T, j0 = tx * ST_T;
  const int pi = i0 + (lane & 15), pj = j0 + warp * 2 + (lane >> 4);
  const bool inside = pi < p.H && pj < p.W;
  const int N = p.H * p.W;
  const int n = inside ? pi * p.W + pj : 0;
  const int np = p.n_points;
  const int size[3] = {p.vol.D0, p.vol.D1, p.vol.D2};
  const uint32_t bar = smem_u32(&mbar_storage[0]);  // the second barrier sits 8 bytes further

  if (tid == 0) {
    mbar_init(bar, 256);
    mbar_init(bar + 8u, 256);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    t_first = INT_MAX;
    t_last = INT_MIN;
  }

  // ---- this thread's ray, exactly as trilinear_fwd_kernel sets it up
  float s[3], d[3], L;
  generate_ray(p.geom, b, n, p.eps, s, d, L);
  const float lo0[3] = {0.f, 0.f, 0.f};
  const float hi0[3] = {(float)(size[0] - 1), (float)(size[1] - 1), (float)(size[2] - 1)};
  const AlphaRange ar = alpha_range(s, d, lo0, hi0);
  const float span = ar.amax - ar.amin;
  const float w = st_step_weight(p.step_mode, span, np);
  bool live = inside;
  if (live) {  // rays that never come within one voxel of the volume are an exact 0
    const float plo[3] = {-1.f, -1.f, -1.f};
    const float phi[3] = {(float)size[0], (float)size[1], (float)size[2]};
    const AlphaRange pr = alpha_range(s, d, plo, phi);
    live = pr.amin < pr.amax;
  }

  // ---- the frustum's marching axis and direction, from the tile's central ray (uniform over the CTA)
  int A;
  bool forward;
  {
    float cs[3], cd[3], cl;
    generate_ray(p.geom, b, min(i0 + ST_T / 2, p.H - 1) * p.W + min(j0 + ST_T / 2, p.W - 1), p.eps, cs, cd, cl);
    A = fabsf(cd[1]) > fabsf(cd[0]) ? 1 : 0;
    if (fabsf(cd[2]) > fabsf(pick3(cd, A))) A = 2;
    forward = pick3(cd, A) > 0.f;
  }
  const int O1 = A == 0 ? 1 : 0, O2 = A == 2 ? 1 : 2;  // the other two axes, ascending

  // ---- corner rays of the tile, held by lanes 0..3 (and 4..7) of warp 0 for the per-slab box
  float qs[3] = {0.f, 0.f, 0.f}, qd[3] = {1.f, 1.f, 1.f};
  if (warp == 0 && lane < 8) {
    const int ci = (lane & 1) ? min(i0 + ST_T - 1, p.H - 1) : i0;
    const int cj = (lane & 2) ? min(j0 + ST_T - 1, p.W - 1) : j0;
    float ql;
    generate_ray(p.geom, b, ci * p.W + cj, p.eps, qs, qd, ql);
  }
  const float qsA = pick3(qs, A), qdA = pick3(qd, A), qs1 = pick3(qs, O1), qd1 = pick3(qd, O1), qs2 = pick3(qs, O2),
              qd2 = pick3(qd, O2);
  const float sA = pick3(s, A), dA = pick3(d, A);

  const float lstep = 1.0f / (float)(np - 1);
  float sumV = 0.f;
  float Aj[3] = {0.f, 0.f, 0.f}, Uj[3] = {0.f, 0.f, 0.f};
  unsigned n_shared = 0, n_global = 0, n_timeout = 0;

  auto accumulate = [&](float u, float v, const float g[3]) {
    sumV += v;
    if (JAC) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        Aj[a] += g[a];
        Uj[a] = fmaf(u, g[a], Uj[a]);
      }
    }
  };

  // A ray is "regular" when it advances along A in the frustum's direction; anything else (a ray that only grazes
  // the zero padding and is marched with a negative span, a ray running the other way) is marched here and now
  // from global memory, samples in the same order.
  const bool regular = live && span > 0.f && dA != 0.f && ((dA > 0.f) == forward);
  if (live && !regular) {
    for (int k = 0; k < np; ++k) {
      const float u = linspace01(k, np, lstep);
      const float alpha = fmaf(u, span, ar.amin);
      float g[3];
      const float v = sample_trilinear<JAC, false>(p.vol, fmaf(alpha, d[0], s[0]), fmaf(alpha, d[1], s[1]),
                                                   fmaf(alpha, d[2], s[2]), g);
      accumulate(u, v, g);
    }
    n_global += np;
  }

  __syncthreads();  // barrier initialised, t_first / t_last reset
  if (regular) {
    const float a_end = fmaf(1.0f, span, ar.amin);
    const int c0 = (int)floorf(fmaf(ar.amin, dA, sA)), c1 = (int)floorf(fmaf(a_end, dA, sA));
    const int sa = slab_of(c0), sb = slab_of(c1);
    atomicMin(&t_first, forward ? sa : -sa);
    atomicMax(&t_last, forward ? sb : -sb);
  }
  __syncthreads();
  const int tf = t_first, tl = t_last;  // travel index t = +-slab number, increasing along the rays

  // next sample of this ray: index k, and (once computed) its position
  int k = 0;
  bool have = false;
  float cu = 0.f, cx = 0.f, cy = 0.f, cz = 0.f;
  int ct = 0;
  bool broken = false;  // a barrier wait timed out: this thread stops trusting the staging buffers

  // ---- 1. box of slab t (warp 0; lanes 0-3: corner rays at plane loA, lanes 4-7: at plane loA + ST_K)
  auto publish_box = [&](int t, BoxDesc& desc) {
    const int slab = forward ? t : -t;
    const int loA = (slab - 1) * ST_K;
    float v1 = 0.f, v2 = 0.f;
    bool ok = true;
    if (lane < 8) {
      const float plane = (float)(loA + ((lane & 4) ? ST_K : 0));
      ok = fabsf(qdA) > 1e-12f;
      const float al = ok ? (plane - qsA) / qdA : 0.f;
      v1 = fmaf(al, qd1, qs1);
      v2 = fmaf(al, qd2, qs2);
      ok = ok && fabsf(v1) < 1e8f && fabsf(v2) < 1e8f;
    }
    float mn1 = lane < 8 ? v1 : INFINITY, mx1 = lane < 8 ? v1 : -INFINITY;
    float mn2 = lane < 8 ? v2 : INFINITY, mx2 = lane < 8 ? v2 : -INFINITY;
    const unsigned bad = __ballot_sync(0xffffffffu, !ok);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      mn1 = fminf(mn1, __shfl_xor_sync(0xffffffffu, mn1, o));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
      mn2 = fminf(mn2, __shfl_xor_sync(0xffffffffu, mn2, o));
      mx2 = fmaxf(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
    }
    if (lane == 0) {
      // cells loA .. loA+K-1 need layers loA .. loA+K; on the other axes one cell of margin each side (the
      // per-sample positions are rounded differently from this bound) plus the upper corner
      const int l1 = (int)floorf(mn1) - 1, h1 = (int)floorf(mx1) + 2;
      const int l2 = (int)floorf(mn2) - 1, h2 = (int)floorf(mx2) + 2;
      int lo[3], hi[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {  // nothing beyond one layer of zero padding is ever read
        lo[a] = max(a == A ? loA : (a == O1 ? l1 : l2), -1);
        hi[a] = min(a == A ? loA + ST_K : (a == O1 ? h1 : h2), size[a]);
      }
      lo[2] &= ~3;  // 16-byte rows (two's complement: -1 -> -4)
      hi[2] = ((hi[2] + 4) & ~3) - 1;
      long long elems = 1;
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        desc.lo[a] = lo[a];
        desc.E[a] = hi[a] - lo[a] + 1;
        elems *= (long long)max(desc.E[a], 0);
      }
      desc.staged = bad == 0 && elems > 0 && elems <= ST_CAP && (size[2] & 3) == 0 && ((size_t)p.vol.data & 15) == 0;
    }
  };

  // ---- 2. stage a box: one bulk copy per row that intersects the volume, zeros elsewhere.  Called by every thread
  // (CTA-uniform on desc.staged): every thread arrives at `bar_addr` exactly once per staged box.
  auto stage_box = [&](const BoxDesc& desc, float* buf, uint32_t bar_addr) {
    if (desc.staged == 0) return;
    if (broken) {
      mbar_arrive(bar_addr);
      return;
    }
    // the buffer was read (and its padding written) through the generic proxy; the bulk copies write it through the
    // async proxy
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    const int bl0 = desc.lo[0], bl1 = desc.lo[1], bl2 = desc.lo[2];
    const int E0 = desc.E[0], E1 = desc.E[1], E2 = desc.E[2];
    const int rows = E0 * E1;
    const int c_lo = max(bl2, 0), c_hi = min(bl2 + E2, size[2]);  // multiples of 4
    uint32_t bytes = 0;
    for (int r = tid; r < rows; r += 256) {
      const int g0 = bl0 + r / E1, g1 = bl1 + r % E1;
      if ((unsigned)g0 < (unsigned)size[0] && (unsigned)g1 < (unsigned)size[1] && c_hi > c_lo)
        bytes += (uint32_t)(c_hi - c_lo) * 4u;
    }
    if (bytes) mbar_arrive_expect_tx(bar_addr, bytes); else mbar_arrive(bar_addr);
    for (int r = tid; r < rows; r += 256) {
      const int g0 = bl0 + r / E1, g1 = bl1 + r % E1;
      float* row = buf + (size_t)r * E2;
      const bool in = (unsigned)g0 < (unsigned)size[0] && (unsigned)g1 < (unsigned)size[1] && c_hi > c_lo;
      const int z_lo = in ? c_lo - bl2 : E2, z_hi = in ? c_hi - bl2 : E2;  // [z_lo, z_hi) comes from the volume
      for (int c = 0; c < z_lo; ++c) row[c] = 0.f;
      for (int c = z_hi; c < E2; ++c) row[c] = 0.f;
      if (in)
        bulk_g2s(smem_u32(row + z_lo), p.vol.data + ((int64_t)g0 * p.vol.s0 + (int64_t)g1 * p.vol.s1 + c_lo),
                 (uint32_t)(c_hi - c_lo) * 4u, bar_addr);
    }
  };

  // bounded wait for the bulk copies of a staged box; false (and `broken`) when the barrier does not complete
  auto wait_box = [&](uint32_t bar_addr, uint32_t parity) -> bool {
    if (broken) return false;
    bool done = false;
    for (int spin = 0; spin < (1 << 16) && !done; ++spin) done = mbar_try_wait(bar_addr, parity);
    if (!done) {
      broken = true;
      ++n_timeout;
    }
    return done;
  };

  // ---- 3. this ray's samples whose cell lies in slab t, from `buf` when `staged`
  auto march_slab = [&](int t, const BoxDesc& desc, const float* buf, bool staged) {
    if (!regular) return;
    const int bl0 = desc.lo[0], bl1 = desc.lo[1], bl2 = desc.lo[2];
    const int E0 = desc.E[0], E1 = desc.E[1], E2 = desc.E[2];
    while (k < np) {
      if (!have) {
        cu = linspace01(k, np, lstep);
        const float alpha = fmaf(cu, span, ar.amin);
        cx = fmaf(alpha, d[0], s[0]);
        cy = fmaf(alpha, d[1], s[1]);
        cz = fmaf(alpha, d[2], s[2]);
        const float pa = A == 0 ? cx : (A == 1 ? cy : cz);
        const int cs = slab_of((int)floorf(pa));
        ct = forward ? cs : -cs;
        have = true;
      }
      if (ct > t) break;  // belongs to a later slab
      float g[3], v;
      const float fx0 = floorf(cx), fy0 = floorf(cy), fz0 = floorf(cz);
      const int lx = (int)fx0 - bl0, ly = (int)fy0 - bl1, lz = (int)fz0 - bl2;
      if (staged && ct == t && (unsigned)lx < (unsigned)(E0 - 1) && (unsigned)ly < (unsigned)(E1 - 1) &&
          (unsigned)lz < (unsigned)(E2 - 1)) {
        const float* q = buf + ((size_t)lx * E1 + ly) * E2 + lz;
        const int sy = E2, sx = E1 * E2;
        v = trilinear_interp<JAC>(q[0], q[1], q[sy], q[sy + 1], q[sx], q[sx + 1], q[sx + sy], q[sx + sy + 1],
                                  cx - fx0, cy - fy0, cz - fz0, g);
        ++n_shared;
      } else {
        v = sample_trilinear<JAC, false>(p.vol, cx, cy, cz, g);
        ++n_global;
      }
      accumulate(cu, v, g);
      have = false;
      ++k;
    }
  };

  if (STAGES == 1) {
    // One buffer: publish -> barrier -> stage -> barrier + wait -> march.  Box descriptors are double-buffered so
    // that warp 0 can publish slab t+1 while slower warps still copy slab t's descriptor into registers.
    uint32_t phase = 0;
    for (int t = tf; t <= tl; ++t) {
      BoxDesc& desc = descs[(t - tf) & 1];
      if (warp == 0) publish_box(t, desc);
      __syncthreads();  // everyone is done with the previous slab's buffer; the new box is published
      bool staged = desc.staged != 0;
      if (staged) {  // CTA-uniform
        stage_box(desc, box, bar);
        __syncthreads();  // the zero fills are visible
        staged = wait_box(bar, phase & 1u);
        ++phase;
      }
      march_slab(t, desc, box, staged);
    }
  } else {
    // Two buffers: while slab t is marched from buffer t&1, the bulk copies of slab t+1 fill buffer (t+1)&1.  One
    // barrier per slab.  Descriptors live in three slots: slot (t+1)%3 is rewritten while slab t-1's readers are
    // still possible only for slots (t-1)%3 and t%3.
    uint32_t uses0 = 0u, uses1 = 0u;  // completed phases of each barrier
    if (tf <= tl) {
      if (warp == 0) publish_box(tf, descs[0]);
      __syncthreads();
      stage_box(descs[0], box, bar);
    }
    for (int t = tf; t <= tl; ++t) {
      const int i = t - tf;
      BoxDesc& desc = descs[i % 3];
      BoxDesc& next = descs[(i + 1) % 3];
      if (t < tl && warp == 0) publish_box(t + 1, next);
      __syncthreads();  // slab t-1 fully marched (its buffer is free), slab t's zero fills and the next box visible
      if (t < tl) stage_box(next, box + (size_t)((i + 1) & 1) * ST_CAP, bar + 8u * (uint32_t)((i + 1) & 1));
      bool staged = desc.staged != 0;
      if (staged) {
        staged = wait_box(bar + 8u * (uint32_t)(i & 1), ((i & 1) ? uses1 : uses0) & 1u);
        if (i & 1) ++uses1; else ++uses0;
      }
      march_slab(t, desc, box + (size_t)(i & 1) * ST_CAP, staged);
    }
  }

  if (p.stats && (n_shared | n_global | n_timeout)) {
    atomicAdd(p.stats + 0, (unsigned long long)n_shared);
    atomicAdd(p.stats + 1, (unsigned long long)n_global);
    atomicAdd(p.stats + 2, (unsigned long long)n_timeout);
  }
  if (!inside) return;

  // ---- epilogue: identical to trilinear_fwd_kernel's
  const int64_t ray = (int64_t)b * N + n;
  p.out[ray] = sumV * L * w;
  if (JAC) {
    const float T = fmaf(Aj[0], d[0], fmaf(Aj[1], d[1], Aj[2] * d[2]));
    const float Q = fmaf(Uj[0], d[0], fmaf(Uj[1], d[1], Uj[2] * d[2]));
    const float P = T - Q;
    const float wp = st_step_weight_dspan(p.step_mode, np);
    float js[3], jt[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float Bv = fmaf(ar.amin, Aj[a], span * Uj[a]);
      js[a] = w * (Aj[a] - Bv);
      jt[a] = w * Bv;
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (ar.axis_min == a) {
        const float dads = (ar.amin - 1.f) / d[a], dadt = -ar.amin / d[a];
        js[a] += (w * P - sumV * wp) * dads;
        jt[a] += (w * P - sumV * wp) * dadt;
      }
      if (ar.axis_max == a) {
        const float dads = (ar.amax - 1.f) / d[a], dadt = -ar.amax / d[a];
        js[a] += (w * Q + sumV * wp) * dads;
        jt[a] += (w * Q + sumV * wp) * dadt;
      }
    }
    float* j = p.jac + (int64_t)b * 7 * N + n;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      j[(int64_t)a * N] = L * js[a];
      j[(int64_t)(3 + a) * N] = L * jt[a];
    }
    j[(int64_t)6 * N] = sumV * w;
  }
}

}  // namespace xvr

using namespace xvr;

// Same arguments as xvr_trilinear_drr_fwd without the texture handle and tile shape, plus `stages` (1: one staging
// buffer, three CTAs per SM; 2: double-buffered, the copies of slab t+1 overlap the march of slab t, two CTAs per
// SM) and an optional device counter triple `stats` = {samples served from shared memory, from global memory,
// barrier time-outs} the caller zeroes.
extern "C" int xvr_trilinear_drr_fwd_staged(const float* volume, int D0, int D1, int D2, const float* cam2vox,
                                            const float* cam2world, const float* det9, int B, int det_h, int det_w,
                                            int n_points, int step_mode, float eps, int stages, float* out, float* jac,
                                            unsigned long long* stats, void* stream) {
  if (!volume || !cam2vox || !cam2world || !det9 || !out || B <= 0 || det_h <= 0 || det_w <= 0 || D0 < 2 || D1 < 2 ||
      D2 < 2 || n_points < 2 || step_mode < 0 || step_mode > 2 || (stages != 1 && stages != 2)) {
    set_last_error("xvr_trilinear_drr_fwd_staged: invalid argument");
    return XVR_ERR_INVALID;
  }
  if ((int64_t)D0 * D1 * D2 >= (int64_t)1 << 31) {
    set_last_error("xvr_trilinear_drr_fwd_staged: volume too large for 32-bit voxel offsets");
    return XVR_ERR_INVALID;
  }
  StagedParams p = {};
  p.vol.data = volume;
  p.vol.D0 = D0;
  p.vol.D1 = D1;
  p.vol.D2 = D2;
  p.vol.s0 = D1 * D2;
  p.vol.s1 = D2;
  p.vol.tex = 0;
  p.geom.cam2vox = cam2vox;
  p.geom.cam2world = cam2world;
  for (int a = 0; a < 3; ++a) {
    p.geom.o[a] = det9[a];
    p.geom.u[a] = det9[3 + a];
    p.geom.v[a] = det9[6 + a];
  }
  p.geom.W = det_w;
  p.B = B;
  p.H = det_h;
  p.W = det_w;
  p.n_points = n_points;
  p.step_mode = step_mode;
  p.eps = eps;
  p.tiles_x = (det_w + ST_T - 1) / ST_T;
  p.tiles_y = (det_h + ST_T - 1) / ST_T;
  p.out = out;
  p.jac = jac;
  p.stats = stats;
  const int64_t grid = (int64_t)B * p.tiles_x * p.tiles_y;
  if (grid >= (int64_t)1 << 31) {
    set_last_error("xvr_trilinear_drr_fwd_staged: grid too large");
    return XVR_ERR_INVALID;
  }
  const int smem = stages * ST_CAP * (int)sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    kernel<<<(unsigned)grid, 256, smem, st>>>(p);
  };
  if (stages == 1) {
    if (jac) launch(trilinear_fwd_staged_kernel<true, 1>); else launch(trilinear_fwd_staged_kernel<false, 1>);
  } else {
    if (jac) launch(trilinear_fwd_staged_kernel<true, 2>); else launch(trilinear_fwd_staged_kernel<false, 2>);
  }
  return check_launch("xvr_trilinear_drr_fwd_staged");
}
