// Trilinear DRR forward (+ Jacobian) with the volume staged brick by brick in shared memory by the TMA unit -- the
// formulation north_star names ("TMA-staged volume bricks into shared memory").  Version 3 (round 2); the round-1
// version (per-thread cp.async.bulk rows behind CTA-wide barriers) measured 40 ms per launch at config 2 against
// 15.2 ms for the texture kernel and was dropped.
//
// STATUS (measured on the B200, profiles/r2_staged.md): bit-identical to the texture kernel on every parity case, 90-95 %
// of the config-2 samples served from shared memory, no time-outs -- and 3.3x SLOWER than the texture kernel (49.4 ms
// against 15.1 ms per 116-pose launch for the best tile / ring shape).  The reason is geometric, not a tuning gap: at
// config 2 neighbouring rays are 1.4-2.8 voxels apart, so the axis-aligned box around a 16 x 16-ray frustum slab holds
// 3-4 voxels per voxel that a sample actually touches (35-70 B of L2->SM fill per sample, 130-250 GB per launch,
// against 79 GB for the texture path, which fetches only the footprints it needs).  OPT-IN (XVR_B200_STAGED=1); the
// texture kernel stays the default.  It would pay on denser ray bundles (detector pixels <= one voxel apart).
//
// Same contract as xvr_trilinear_drr_fwd (csrc/trilinear.cu): same rays, same alpha_k, same sample positions, same
// 8-corner blend (trilinear_interp) accumulated in the same order -- only the SOURCE of the 8 corners differs, so the
// output is bit-identical to the texture kernel's, which is what its tests assert.
//
// Structure.  A CTA owns a ST_TI x ST_TJ detector tile of one pose: NW consumer warps (one ray per thread, a warp =
// ST_TI rows x 32/ST_TI columns) and ONE producer warp.  The tile's rays form a thin frustum that advances along its
// dominant volume axis A; the frustum is cut into stages of ST_K cell layers (ST_K + 1 voxel layers).  The producer
// bounds each stage's footprint on the other two axes from the tile's four corner rays (the pixel -> plane map is
// projective, so the tile's image on a plane x_A = const is the convex hull of its corner images), picks the smallest
// box of a fixed menu that covers it and issues ONE cp.async.bulk.tensor.3d (SASS UTMALDG) into a ring of ST_R
// shared-memory stages; out-of-volume voxels are zero-filled by the TMA unit (= grid_sample's zero padding).
// Completion is an mbarrier per ring slot (expect-tx); consumers never meet at a CTA-wide barrier: each warp walks
// its samples k = 0..n-1 in lock step, waits for the stage its lanes are in, and publishes the lowest stage it
// still needs in a shared progress word that the producer polls before it reuses a slot.
// Anything a box cannot serve is served from global memory with the reference arithmetic (sample_trilinear's load
// path): samples outside their box, rays that run against the frustum's direction, lanes that run more than a ring
// ahead of their warp, boxes beyond the menu, and -- every wait is bounded -- a barrier that does not complete.
#include <cuda.h>
#include <stdio.h>

#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace xvr {

#ifndef XVR_ST_TI
#define XVR_ST_TI 16  // tile rows = rays along a warp's lanes (16 or 32)
#endif
#ifndef XVR_ST_NW
#define XVR_ST_NW 8  // consumer warps per CTA
#endif
#ifndef XVR_ST_K
#define XVR_ST_K 4  // cell layers per stage (power of two)
#endif
#ifndef XVR_ST_R
#define XVR_ST_R 2  // ring slots
#endif
#ifndef XVR_ST_MIN_CTAS
#define XVR_ST_MIN_CTAS 2
#endif
#ifndef XVR_ST_AREA
#define XVR_ST_AREA 2560  // largest box footprint EP * EQ (floats per voxel layer)
#endif
constexpr int ST_TI = XVR_ST_TI;
constexpr int ST_CPW = 32 / ST_TI;       // detector columns per warp
constexpr int ST_NW = XVR_ST_NW;
constexpr int ST_TJ = ST_NW * ST_CPW;    // tile columns
constexpr int ST_NC = ST_NW * 32;        // consumer threads
constexpr int ST_K = XVR_ST_K;
constexpr int ST_KLOG = ST_K == 2 ? 1 : (ST_K == 4 ? 2 : 3);
constexpr int ST_L = ST_K + 1;           // voxel layers per stage
constexpr int ST_LP = (ST_L + 3) & ~3;   // ... padded to the TMA unit's 16-byte inner extent when the layers are innermost
constexpr int ST_R = XVR_ST_R;
constexpr int ST_AREA = XVR_ST_AREA;
constexpr int ST_STAGE_FLOATS = ST_L * ST_AREA;
constexpr int ST_MENU = 9;               // box extents 16, 24, ..., 80 on each lateral axis
constexpr int ST_EMIN = 16, ESTEP = 8;
static_assert((1 << ST_KLOG) == ST_K && ST_K % 4 == 0, "ST_K must be 4 or 8 (stage origins along axis 2 are 16-byte aligned)");
static_assert(ST_TI == 16 || ST_TI == 32, "a warp covers 16 or 32 detector rows");
static_assert((ST_STAGE_FLOATS * 4) % 128 == 0, "TMA destinations are 128-byte aligned");

struct StagedParams {
  Vol vol;
  DetectorGeom geom;
  int B, H, W, n_points, step_mode;
  float eps;
  int tiles_x, tiles_y;
  const CUtensorMap* __restrict__ maps;  // [3 axes][ST_MENU (EP)][ST_MENU (EQ)], device memory
  float* __restrict__ out;               // (B,1,H*W)
  float* __restrict__ jac;               // (B,7,H*W), nullable
  unsigned long long* __restrict__ stats;  // nullable: {samples from shared memory, from global memory, time-outs}
};

// One ring slot's box: lateral origin and strides of the voxel layout the TMA unit wrote.
struct __align__(16) StageDesc {
  int lo;      // loP | loQ << 16  (both biased by +4096 so that they are non-negative)
  int sP;      // element stride of the first lateral axis
  int sA;      // element stride of the marching axis
  int lim;     // (EP - 1) | (EQ - 1) << 16, 0 when the stage is not staged (every sample fails the box test)
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

// ---- mbarrier / TMA wrappers (PTX ISA 8.x, sm_90+)
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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(dst), "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
      : "memory");
}

template <int A>
struct Axes {  // the two lateral axes (ascending) and the element stride of the second one inside a staged box
  static constexpr int P = A == 0 ? 1 : 0;
  static constexpr int Q = A == 2 ? 1 : 2;
  static constexpr int SQ = A == 2 ? ST_LP : 1;
  static constexpr int LAYERS = A == 2 ? ST_LP : ST_L;  // voxel layers a staged box holds (A = 2: padded)
};

struct CtaShared {
  unsigned long long full[ST_R];  // mbarriers: slot filled
  StageDesc desc[ST_R];
  int progress[ST_NW];            // lowest stage each consumer warp still needs
  int t_first, t_last;
  int issued;                     // highest stage the producer has issued (its slot's previous load had completed)
  int broken;                     // a wait timed out somewhere: everybody stops trusting the ring
};

// ---------------------------------------------------------------------------------------------------------------
template <int A, bool JAC>
__device__ __forceinline__ void consumer_march(const StagedParams& p, CtaShared& sh, const float* __restrict__ ring,
                                               int tf, bool forward, bool regular, const float s[3], const float d[3],
                                               float amin, float span, int warp, float& sumV, float Aj[3], float Uj[3],
                                               unsigned& n_shared, unsigned& n_global, unsigned& n_timeout) {
  using AX = Axes<A>;
  const int np = p.n_points;
  const float lstep = 1.0f / (float)(np - 1);
  const int size_a = A == 0 ? p.vol.D0 : (A == 1 ? p.vol.D1 : p.vol.D2);
  // travel index of cell c: forward c + ST_K (cells >= -1 stay non-negative), backward (size_a + ST_K - 1) - c;
  // stage = travel index >> ST_KLOG; the stage's first REAL cell layer is the smaller end either way
  const int off = forward ? ST_K : size_a + ST_K - 1;
  const int sgn = forward ? 1 : -1;
  const uint32_t full0 = smem_u32(&sh.full[0]);

  int cur_t = INT_MIN;           // stage whose descriptor this lane holds
  int d_lo = 0, d_sP = 0, d_sA = 0, d_lim = 0, d_first = 0;  // ... and the descriptor: origins, strides, limits, first layer
  const float* d_base = ring;
  int w_ready = tf - 1;          // highest stage this warp has seen filled (warp-uniform)
  int w_rel = tf;                // progress published so far (warp-uniform)
  bool trust = true;             // false once a wait timed out (CTA-wide flag mirrored per warp)
#ifdef XVR_ST_DEBUG
  int why = 0;
#endif

  if (!__any_sync(0xffffffffu, regular)) {  // nothing to march here: do not hold the producer back
    if ((threadIdx.x & 31) == 0) *(volatile int*)&sh.progress[warp] = INT_MAX;
    return;
  }

  // Every lane walks its own samples k = 0 .. np-1 in order.  Rays of a warp that enter and leave through the faces
  // the frustum marches between stay at the same depth for the same k; rays clipped by a side face do not, so a lane
  // whose next sample lies in a stage beyond the ring window STALLS (keeps its k) until the lanes behind it have
  // caught up -- the lanes of the warp's lowest stage always proceed.
  int k = 0;
  while (__any_sync(0xffffffffu, regular && k < np)) {
    const bool act = regular && k < np;
    bool stall = false;
    const float u = linspace01(min(k, np - 1), np, lstep);
    const float alpha = fmaf(u, span, amin);
    const float x = fmaf(alpha, d[0], s[0]);
    const float y = fmaf(alpha, d[1], s[1]);
    const float z = fmaf(alpha, d[2], s[2]);
    const float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
    const int ix = (int)fx0, iy = (int)fy0, iz = (int)fz0;
    const int ca = A == 0 ? ix : (A == 1 ? iy : iz);
    const int ip = AX::P == 0 ? ix : iy;
    const int iq = AX::Q == 1 ? iy : iz;
    const int tU = act ? (off + sgn * ca) >> ST_KLOG : cur_t;

    if (__any_sync(0xffffffffu, tU != cur_t)) {
      // ---- slow path, entered by the whole warp whenever a lane changes stage (every ~ST_K samples)
      const int tmin = __reduce_min_sync(0xffffffffu, act ? tU : INT_MAX);
      const int tmax = __reduce_max_sync(0xffffffffu, act ? tU : INT_MIN);
      if (trust && tmin > w_rel) {  // stages below tmin are done with: let the producer reuse their slots
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
          __threadfence_block();
          *(volatile int*)&sh.progress[warp] = tmin;
        }
        w_rel = tmin;
      }
      // the ring holds stages w_rel .. w_rel + ST_R - 1: lanes beyond it are served from global memory
      const int thi = min(tmax, tmin + ST_R - 1);
      // stages below tmin are of no interest (and their slots may be several phases on): wait for tmin .. thi only,
      // each of which cannot have been overwritten because this warp has not released it
      w_ready = max(w_ready, tmin - 1);
      while (trust && w_ready < thi) {
        const int q = w_ready + 1 - tf;
        const uint32_t bar = full0 + 8u * (uint32_t)(q % ST_R);
        const uint32_t parity = (uint32_t)(q / ST_R) & 1u;
        // A parity wait only means "use q / ST_R of this slot has completed" once the previous use is known to have
        // completed -- else a warp that skipped that use (rays that start deep in the volume) would sail through a
        // barrier that is still one phase behind.  The producer issues a stage only after its slot's previous load
        // has landed, and says so in `issued`.
        bool done = false;
        for (int spin = 0; spin < (1 << 22) && !done; ++spin) {
          done = *(volatile int*)&sh.issued >= w_ready + 1;
          if (!done) __nanosleep(32);
        }
        if (done) {
          done = false;
          for (int spin = 0; spin < (1 << 20) && !done; ++spin) done = mbar_try_wait(bar, parity);
        }
        if (!done || *(volatile int*)&sh.broken) {
          if (!done) {
            *(volatile int*)&sh.broken = 1;
            if ((threadIdx.x & 31) == 0) ++n_timeout;
          }
          trust = false;
          break;
        }
        ++w_ready;
      }
      if (tU != cur_t && act) {
#ifdef XVR_ST_DEBUG
        why = (trust && tU <= w_ready) ? 0 : 2;
#endif
        if (trust && tU <= w_ready) {
          const int slot = (tU - tf) % ST_R;
          const int4 dd = *reinterpret_cast<const int4*>(&sh.desc[slot]);
          d_lo = dd.x;
          d_sP = dd.y;
          d_sA = dd.z;
          d_lim = dd.w;
          d_base = ring + (size_t)slot * ST_STAGE_FLOATS;
          // first real voxel layer of the stage: forward (tU << KLOG) - ST_K, backward size_a - 1 - (tU << KLOG) ... - (K - 1)
          d_first = forward ? (tU << ST_KLOG) - ST_K : (size_a + ST_K - 1) - (tU << ST_KLOG) - (ST_K - 1);
          cur_t = tU;
#ifdef XVR_ST_DEBUG
          if (d_lim == 0) why = 1;
#endif
        } else if (trust) {
          stall = true;  // beyond the ring window: wait for the lanes behind (asks again next time round)
        } else {
          d_lim = 0;  // the ring is broken (a wait timed out): global memory from here on
        }
      }
    }

    float g[3], v;
    const int lp = ip - ((d_lo & 0xffff) - 4096), lq = iq - ((int)((unsigned)d_lo >> 16) - 4096);
    if (!act || stall) continue;
    ++k;
    if ((unsigned)lp < (unsigned)(d_lim & 0xffff) && (unsigned)lq < ((unsigned)d_lim >> 16)) {
      const float* q0 = d_base + (ca - d_first) * d_sA + lp * d_sP + lq * AX::SQ;
      const float* q1 = q0 + d_sA;  // next voxel layer along the marching axis
      float c000, c001, c010, c011, c100, c101, c110, c111;
      if (A == 0) {  // x = A, y = P, z = Q
        c000 = q0[0]; c001 = q0[AX::SQ]; c010 = q0[d_sP]; c011 = q0[d_sP + AX::SQ];
        c100 = q1[0]; c101 = q1[AX::SQ]; c110 = q1[d_sP]; c111 = q1[d_sP + AX::SQ];
      } else if (A == 1) {  // x = P, y = A, z = Q
        c000 = q0[0]; c001 = q0[AX::SQ]; c100 = q0[d_sP]; c101 = q0[d_sP + AX::SQ];
        c010 = q1[0]; c011 = q1[AX::SQ]; c110 = q1[d_sP]; c111 = q1[d_sP + AX::SQ];
      } else {  // x = P, y = Q, z = A
        c000 = q0[0]; c010 = q0[AX::SQ]; c100 = q0[d_sP]; c110 = q0[d_sP + AX::SQ];
        c001 = q1[0]; c011 = q1[AX::SQ]; c101 = q1[d_sP]; c111 = q1[d_sP + AX::SQ];
      }
      v = trilinear_interp<JAC>(c000, c001, c010, c011, c100, c101, c110, c111, x - fx0, y - fy0, z - fz0, g);
      ++n_shared;
    } else {
      v = sample_trilinear<JAC, false>(p.vol, x, y, z, g);
      ++n_global;
#ifdef XVR_ST_DEBUG
      if (p.stats) atomicAdd(p.stats + 4 + (why == 0 ? (d_lim == 0 ? 1 : 3) : why), 1ull);  // 5 unstaged, 6 ahead of ring, 7 box miss
#endif
    }
    sumV += v;
    if (JAC) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        Aj[a] += g[a];
        Uj[a] = fmaf(u, g[a], Uj[a]);
      }
    }
  }
  // done: nothing of the ring is needed any more
  __syncwarp();
  if ((threadIdx.x & 31) == 0) {
    __threadfence_block();
    *(volatile int*)&sh.progress[warp] = INT_MAX;
  }
}

// The producer warp: for every stage tf .. tl, wait until its ring slot is free, bound the box, issue the TMA load.
template <int A>
__device__ __forceinline__ void producer_loop(const StagedParams& p, CtaShared& sh, float* ring, int tf, int tl,
                                              bool forward, int b, int i0, int j0) {
  using AX = Axes<A>;
  const int lane = threadIdx.x & 31;
  const int size[3] = {p.vol.D0, p.vol.D1, p.vol.D2};
  const int size_a = size[A];
  // corner rays of the tile: lanes 0..3 (near plane of a stage) and 4..7 (far plane)
  float qs[3] = {0.f, 0.f, 0.f}, qd[3] = {1.f, 1.f, 1.f};
  if (lane < 8) {
    const int ci = (lane & 1) ? min(i0 + ST_TI - 1, p.H - 1) : i0;
    const int cj = (lane & 2) ? min(j0 + ST_TJ - 1, p.W - 1) : j0;
    float ql;
    generate_ray(p.geom, b, ci * p.W + cj, p.eps, qs, qd, ql);
  }
  const float qsA = qs[A], qdA = qd[A], qsP = qs[AX::P], qdP = qd[AX::P], qsQ = qs[AX::Q], qdQ = qd[AX::Q];
  const uint32_t full0 = smem_u32(&sh.full[0]);
  const bool aligned = ((size_t)p.vol.data & 15) == 0 && (size[2] & 3) == 0;

  int issued = 0;
  for (int t = tf; t <= tl; ++t, ++issued) {
    const int q = t - tf;
    const int slot = q % ST_R;
    // ---- the slot is free once its previous load has landed (nobody may have waited for it: stages that no ray
    // samples) and every consumer warp's progress has passed stage t - ST_R
    if (q >= ST_R) {
      {
        const uint32_t bar = full0 + 8u * (uint32_t)slot;
        const uint32_t parity = (uint32_t)(q / ST_R - 1) & 1u;
        bool landed = false;
        for (int spin = 0; spin < (1 << 22) && !landed; ++spin) landed = mbar_try_wait(bar, parity);
        if (!landed) {
          *(volatile int*)&sh.broken = 1;
          break;
        }
      }
      bool free_ = false;
      for (int spin = 0; spin < (1 << 22) && !free_; ++spin) {
        int pr = lane < ST_NW ? *(volatile int*)&sh.progress[lane] : INT_MAX;
        pr = __reduce_min_sync(0xffffffffu, pr);
        free_ = pr > t - ST_R;
        if (!free_) {
          if (*(volatile int*)&sh.broken) break;
          __nanosleep(64);
        }
      }
      if (!free_) {
        *(volatile int*)&sh.broken = 1;
        break;
      }
    }
    // ---- box of stage t: real voxel layers first .. first + ST_K; lateral bounds from the corner rays on the planes
    // that bound every sample whose cell lies in the stage
    const int first = forward ? (t << ST_KLOG) - ST_K : (size_a + ST_K - 1) - (t << ST_KLOG) - (ST_K - 1);
    float vP = 0.f, vQ = 0.f;
    bool ok = true;
    if (lane < 8) {
      const float plane = (float)(first + ((lane & 4) ? ST_K : 0));
      ok = fabsf(qdA) > 1e-12f;
      const float al = ok ? (plane - qsA) / qdA : 0.f;
      vP = fmaf(al, qdP, qsP);
      vQ = fmaf(al, qdQ, qsQ);
      ok = ok && al > 0.f && fabsf(vP) < 1e6f && fabsf(vQ) < 1e6f;
    }
    float mnP = lane < 8 ? vP : INFINITY, mxP = lane < 8 ? vP : -INFINITY;
    float mnQ = lane < 8 ? vQ : INFINITY, mxQ = lane < 8 ? vQ : -INFINITY;
    const unsigned bad = __ballot_sync(0xffffffffu, !ok);
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      mnP = fminf(mnP, __shfl_xor_sync(0xffffffffu, mnP, o));
      mxP = fmaxf(mxP, __shfl_xor_sync(0xffffffffu, mxP, o));
      mnQ = fminf(mnQ, __shfl_xor_sync(0xffffffffu, mnQ, o));
      mxQ = fmaxf(mxQ, __shfl_xor_sync(0xffffffffu, mxQ, o));
    }
    if (lane == 0) {
      // a sample at lateral position v reads voxels floor(v), floor(v) + 1; nothing beyond one layer of zero
      // padding is ever read (samples lie inside the volume)
      const int loP = max((int)floorf(mnP), -1), hiP = min((int)floorf(mxP) + 1, size[AX::P]);
      // The TMA unit faults (illegal instruction) unless the innermost tensor coordinate is 16-byte aligned
      // (measured: scripts/tma_probe2.cu): the contiguous axis 2 is the second lateral axis when A = 0 or 1 -- align
      // its origin down to a multiple of 4 voxels (-1 -> -4, two's complement) -- and the marching axis itself when
      // A = 2, whose stage origins are multiples of ST_K given D2 % 4 == 0.
      const int loQ = A == 2 ? max((int)floorf(mnQ), -1) : (max((int)floorf(mnQ), -1) & ~3);
      const int hiQ = min((int)floorf(mxQ) + 1, size[AX::Q]);
      const int nP = hiP - loP + 1, nQ = hiQ - loQ + 1;
      const int mP = max(0, (nP - ST_EMIN + ESTEP - 1) / ESTEP), mQ = max(0, (nQ - ST_EMIN + ESTEP - 1) / ESTEP);
      const int EP = ST_EMIN + ESTEP * mP, EQ = ST_EMIN + ESTEP * mQ;
      const bool staged = aligned && bad == 0 && nP > 0 && nQ > 0 && mP < ST_MENU && mQ < ST_MENU &&
                          AX::LAYERS * EP * EQ <= ST_STAGE_FLOATS;
      StageDesc dsc;
      dsc.lo = (loP + 4096) | ((loQ + 4096) << 16);
      // layouts the TMA unit writes (innermost first): A=0 {Q, P, layer}, A=1 {Q, layer, P}, A=2 {layer, Q, P}
      dsc.sP = A == 0 ? EQ : (A == 1 ? ST_L * EQ : ST_LP * EQ);
      dsc.sA = A == 0 ? EP * EQ : (A == 1 ? EQ : 1);
      dsc.lim = staged ? ((EP - 1) | ((EQ - 1) << 16)) : 0;
      sh.desc[slot] = dsc;
      const uint32_t bar = full0 + 8u * (uint32_t)slot;
      if (staged) {
        // the slot was read through the generic proxy; the TMA unit writes it through the async proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_arrive_expect_tx(bar, (uint32_t)(AX::LAYERS * EP * EQ * 4));
        const CUtensorMap* map = p.maps + ((A * ST_MENU + mP) * ST_MENU + mQ);
        const uint32_t dst = smem_u32(ring + (size_t)slot * ST_STAGE_FLOATS);
        // tensor coordinates are (axis 2, axis 1, axis 0)
        int c[3];
        c[2 - A] = first;
        c[2 - AX::P] = loP;
        c[2 - AX::Q] = loQ;
        tma_load_3d(dst, map, c[0], c[1], c[2], bar);
      } else {
        mbar_arrive(bar);
      }
      __threadfence_block();
      *(volatile int*)&sh.issued = t;
    }
    __syncwarp();
  }
  // every load this warp issued must have landed before the CTA may retire (its shared memory is re-used)
  for (int q = max(0, issued - ST_R); q < issued; ++q) {
    const uint32_t bar = full0 + 8u * (uint32_t)(q % ST_R);
    const uint32_t parity = (uint32_t)(q / ST_R) & 1u;
    bool done = false;
    for (int spin = 0; spin < (1 << 22) && !done; ++spin) done = mbar_try_wait(bar, parity);
  }
}

template <bool JAC>
__global__ void __launch_bounds__(ST_NC + 32, XVR_ST_MIN_CTAS) trilinear_fwd_staged_kernel(const __grid_constant__ StagedParams p) {
  extern __shared__ __align__(128) float ring_raw[];  // ST_R * ST_STAGE_FLOATS floats (+ 128 bytes of slack)
  __shared__ CtaShared sh;
  float* ring = reinterpret_cast<float*>(((uintptr_t)ring_raw + 127) & ~(uintptr_t)127);  // TMA destinations: 128 B

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool producer = warp == ST_NW;
  const int tiles = p.tiles_x * p.tiles_y;
  const int b = blockIdx.x / tiles;
  const int tile = blockIdx.x - b * tiles;
  const int ty = tile / p.tiles_x, tx = tile - ty * p.tiles_x;
  const int i0 = ty * ST_TI, j0 = tx * ST_TJ;
  const int pi = i0 + (lane & (ST_TI - 1)), pj = j0 + warp * ST_CPW + (lane / ST_TI);
  const bool inside = !producer && pi < p.H && pj < p.W;
  const int N = p.H * p.W;
  const int n = inside ? pi * p.W + pj : 0;
  const int np = p.n_points;

  if (tid == 0) {
    for (int r = 0; r < ST_R; ++r) mbar_init(smem_u32(&sh.full[r]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    sh.t_first = INT_MAX;
    sh.t_last = INT_MIN;
    sh.broken = 0;
  }

  // ---- this thread's ray, exactly as trilinear_fwd_kernel sets it up
  float s[3], d[3], L;
  generate_ray(p.geom, b, n, p.eps, s, d, L);
  const float lo0[3] = {0.f, 0.f, 0.f};
  const float hi0[3] = {(float)(p.vol.D0 - 1), (float)(p.vol.D1 - 1), (float)(p.vol.D2 - 1)};
  const AlphaRange ar = alpha_range(s, d, lo0, hi0);
  const float span = ar.amax - ar.amin;
  const float w = st_step_weight(p.step_mode, span, np);
  bool live = inside;
  if (live) {  // rays that never come within one voxel of the volume are an exact 0
    const float plo[3] = {-1.f, -1.f, -1.f};
    const float phi[3] = {(float)p.vol.D0, (float)p.vol.D1, (float)p.vol.D2};
    const AlphaRange pr = alpha_range(s, d, plo, phi);
    live = pr.amin < pr.amax;
  }

  // ---- the frustum's marching axis and direction, from the tile's central ray (uniform over the CTA)
  int A;
  bool forward;
  {
    float cs[3], cd[3], cl;
    generate_ray(p.geom, b, min(i0 + ST_TI / 2, p.H - 1) * p.W + min(j0 + ST_TJ / 2, p.W - 1), p.eps, cs, cd, cl);
    A = fabsf(cd[1]) > fabsf(cd[0]) ? 1 : 0;
    if (fabsf(cd[2]) > fabsf(A == 1 ? cd[1] : cd[0])) A = 2;
    forward = (A == 0 ? cd[0] : (A == 1 ? cd[1] : cd[2])) > 0.f;
  }
  const float sA = A == 0 ? s[0] : (A == 1 ? s[1] : s[2]), dA = A == 0 ? d[0] : (A == 1 ? d[1] : d[2]);
  const int size_a = A == 0 ? p.vol.D0 : (A == 1 ? p.vol.D1 : p.vol.D2);

  float sumV = 0.f;
  float Aj[3] = {0.f, 0.f, 0.f}, Uj[3] = {0.f, 0.f, 0.f};
  unsigned n_shared = 0, n_global = 0, n_timeout = 0;

  // A ray is "regular" when it advances along A in the frustum's direction; anything else (a ray that only grazes
  // the zero padding and is marched with a negative span, a ray running the other way) is marched here and now
  // from global memory, samples in the same order.
  const bool regular = live && span > 0.f && dA != 0.f && ((dA > 0.f) == forward);
  if (live && !regular) {
    const float lstep = 1.0f / (float)(np - 1);
    for (int k = 0; k < np; ++k) {
      const float u = linspace01(k, np, lstep);
      const float alpha = fmaf(u, span, ar.amin);
      float g[3];
      const float v = sample_trilinear<JAC, false>(p.vol, fmaf(alpha, d[0], s[0]), fmaf(alpha, d[1], s[1]),
                                                   fmaf(alpha, d[2], s[2]), g);
      sumV += v;
      if (JAC) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          Aj[a] += g[a];
          Uj[a] = fmaf(u, g[a], Uj[a]);
        }
      }
    }
    n_global += np;
#ifdef XVR_ST_DEBUG
    if (p.stats) atomicAdd(p.stats + 4, (unsigned long long)np);  // irregular rays
#endif
  }

  __syncthreads();  // barriers initialised, t_first / t_last reset
  if (regular) {
    const float a_end = fmaf(1.0f, span, ar.amin);
    const int c0 = (int)floorf(fmaf(ar.amin, dA, sA)), c1 = (int)floorf(fmaf(a_end, dA, sA));
    const int off = forward ? ST_K : size_a + ST_K - 1;
    const int sgn = forward ? 1 : -1;
    atomicMin(&sh.t_first, (off + sgn * c0) >> ST_KLOG);
    atomicMax(&sh.t_last, (off + sgn * c1) >> ST_KLOG);
  }
  if (!producer && lane == 0) sh.progress[warp] = INT_MIN;  // set below once t_first is known
  __syncthreads();
  const int tf = sh.t_first, tl = sh.t_last;
  if (!producer && lane == 0) sh.progress[warp] = tf;
  if (tid == 0) sh.issued = tf - 1;
  __syncthreads();  // the last CTA-wide barrier: from here on producer and consumers only meet through the ring

  if (tf <= tl) {
    if (producer) {
      if (A == 0) producer_loop<0>(p, sh, ring, tf, tl, forward, b, i0, j0);
      else if (A == 1) producer_loop<1>(p, sh, ring, tf, tl, forward, b, i0, j0);
      else producer_loop<2>(p, sh, ring, tf, tl, forward, b, i0, j0);
    } else {
      if (A == 0) consumer_march<0, JAC>(p, sh, ring, tf, forward, regular, s, d, ar.amin, span, warp, sumV, Aj, Uj, n_shared, n_global, n_timeout);
      else if (A == 1) consumer_march<1, JAC>(p, sh, ring, tf, forward, regular, s, d, ar.amin, span, warp, sumV, Aj, Uj, n_shared, n_global, n_timeout);
      else consumer_march<2, JAC>(p, sh, ring, tf, forward, regular, s, d, ar.amin, span, warp, sumV, Aj, Uj, n_shared, n_global, n_timeout);
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

// ---------------------------------------------------------------------------------------------------------------
// Tensor maps: one per (marching axis, box extent on the first lateral axis, on the second), in device memory,
// built once per (volume pointer, shape) and cached.  cuTensorMapEncodeTiled comes from the driver through the
// runtime's entry-point query, so the library does not link against libcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct MapKey {
  const void* ptr;
  int D0, D1, D2, device;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && D0 == o.D0 && D1 == o.D1 && D2 == o.D2 && device == o.device;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    return std::hash<const void*>()(k.ptr) ^ ((size_t)k.D0 * 1000003u) ^ ((size_t)k.D1 * 10007u) ^ (size_t)k.D2 ^
           ((size_t)k.device << 20);
  }
};
static std::mutex g_map_mutex;
static std::unordered_map<MapKey, CUtensorMap*, MapKeyHash> g_maps;  // device arrays of 3 * ST_MENU^2 maps

static const CUtensorMap* tensor_maps(const float* volume, int D0, int D1, int D2, cudaStream_t st) {
  int device = 0;
  cudaGetDevice(&device);
  const MapKey key = {volume, D0, D1, D2, device};
  std::lock_guard<std::mutex> lock(g_map_mutex);
  auto it = g_maps.find(key);
  if (it != g_maps.end()) return it->second;
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
      set_last_error("xvr_trilinear_drr_fwd_staged: cuTensorMapEncodeTiled is not available from this driver");
      return nullptr;
    }
    encode = (EncodeTiledFn)fn;
  }
  const int count = 3 * ST_MENU * ST_MENU;
  CUtensorMap* host = new CUtensorMap[count];
  const cuuint64_t dims[3] = {(cuuint64_t)D2, (cuuint64_t)D1, (cuuint64_t)D0};
  const cuuint64_t strides[2] = {(cuuint64_t)D2 * 4, (cuuint64_t)D1 * D2 * 4};
  const cuuint32_t ones[3] = {1, 1, 1};
  for (int A = 0; A < 3; ++A)
    for (int mP = 0; mP < ST_MENU; ++mP)
      for (int mQ = 0; mQ < ST_MENU; ++mQ) {
        const cuuint32_t EP = ST_EMIN + ESTEP * mP, EQ = ST_EMIN + ESTEP * mQ;
        cuuint32_t box[3];  // (axis 2, axis 1, axis 0) extents; lateral axes P < Q
        if (A == 0) { box[0] = EQ; box[1] = EP; box[2] = ST_L; }       // P = axis 1, Q = axis 2
        else if (A == 1) { box[0] = EQ; box[1] = ST_L; box[2] = EP; }  // P = axis 0, Q = axis 2
        else { box[0] = ST_L; box[1] = EQ; box[2] = EP; }              // P = axis 0, Q = axis 1
        // a box that does not fit the stage is never requested; encode a minimal legal one in its place
        if (A == 2) box[0] = ST_LP;  // the inner extent is 16-byte granular: 5 layers -> 8
        // a box that does not fit the stage is never requested; encode a minimal legal one in its place
        if ((A == 2 ? ST_LP : ST_L) * EP * EQ > (cuuint32_t)ST_STAGE_FLOATS) { box[0] = A == 2 ? 4 : 16; box[1] = 1; box[2] = 1; }
        CUresult r = encode(&host[(A * ST_MENU + mP) * ST_MENU + mQ], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)volume,
                            dims, strides, box, ones, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                            CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
          char msg[160];
          snprintf(msg, sizeof(msg), "xvr_trilinear_drr_fwd_staged: cuTensorMapEncodeTiled failed (%d) for axis %d box %u x %u",
                   (int)r, A, EP, EQ);
          set_last_error(msg);
          delete[] host;
          return nullptr;
        }
      }
  CUtensorMap* dev = nullptr;
  if (cudaMalloc(&dev, sizeof(CUtensorMap) * count) != cudaSuccess ||
      cudaMemcpy(dev, host, sizeof(CUtensorMap) * count, cudaMemcpyHostToDevice) != cudaSuccess) {
    set_last_error("xvr_trilinear_drr_fwd_staged: could not upload the tensor maps");
    delete[] host;
    cudaGetLastError();
    return nullptr;
  }
  delete[] host;
  if (g_maps.size() >= 64) {  // volumes come and go (training subjects): keep the table bounded
    for (auto& kv : g_maps) cudaFree(kv.second);
    g_maps.clear();
  }
  g_maps[key] = dev;
  (void)st;
  return dev;
}

}  // namespace xvr

using namespace xvr;

// Same arguments as xvr_trilinear_drr_fwd without the texture handle and tile shape, plus an optional device counter
// triple `stats` = {samples served from shared memory, from global memory, barrier time-outs} the caller zeroes.
// Needs a 16-byte aligned volume whose contiguous extent is a multiple of 4 (the TMA unit's stride granularity);
// other volumes are an invalid argument here and go through xvr_trilinear_drr_fwd.
extern "C" int xvr_trilinear_drr_fwd_staged(const float* volume, int D0, int D1, int D2, const float* cam2vox,
                                            const float* cam2world, const float* det9, int B, int det_h, int det_w,
                                            int n_points, int step_mode, float eps, float* out, float* jac,
                                            unsigned long long* stats, void* stream) {
  if (!volume || !cam2vox || !cam2world || !det9 || !out || B <= 0 || det_h <= 0 || det_w <= 0 || D0 < 2 || D1 < 2 ||
      D2 < 2 || n_points < 2 || step_mode < 0 || step_mode > 2) {
    set_last_error("xvr_trilinear_drr_fwd_staged: invalid argument");
    return XVR_ERR_INVALID;
  }
  if (((size_t)volume & 15) != 0 || (D2 & 3) != 0) {
    set_last_error("xvr_trilinear_drr_fwd_staged: the TMA unit needs a 16-byte aligned volume with D2 % 4 == 0");
    return XVR_ERR_INVALID;
  }
  if ((int64_t)D0 * D1 * D2 >= (int64_t)1 << 31) {
    set_last_error("xvr_trilinear_drr_fwd_staged: volume too large for 32-bit voxel offsets");
    return XVR_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  StagedParams p = {};
  p.maps = tensor_maps(volume, D0, D1, D2, st);
  if (!p.maps) return XVR_ERR_CUDA;
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
  p.tiles_x = (det_w + ST_TJ - 1) / ST_TJ;
  p.tiles_y = (det_h + ST_TI - 1) / ST_TI;
  p.out = out;
  p.jac = jac;
  p.stats = stats;
  const int64_t grid = (int64_t)B * p.tiles_x * p.tiles_y;
  if (grid >= (int64_t)1 << 31) {
    set_last_error("xvr_trilinear_drr_fwd_staged: grid too large");
    return XVR_ERR_INVALID;
  }
  const int smem = ST_R * ST_STAGE_FLOATS * (int)sizeof(float) + 128;
  auto launch = [&](auto kernel) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    kernel<<<(unsigned)grid, ST_NC + 32, smem, st>>>(p);
  };
  if (jac) launch(trilinear_fwd_staged_kernel<true>); else launch(trilinear_fwd_staged_kernel<false>);
  return check_launch("xvr_trilinear_drr_fwd_staged");
}
