// Trilinear DRR renderer: forward (+ per-ray analytic Jacobian) and recompute backward.
//
// Replaces DiffDRR 0.6.0 renderers.Trilinear.forward and the autograd backward of the grid_sample it is
// built on, as called from /root/reference/src/xvr/model/trainer.py:288 (drr.renderer(vol, source, target,
// raylen, mask=seg)) and, through DRR.forward / Registration.forward, from
// /root/reference/src/xvr/registrar/base.py:249.  One thread owns one ray and marches its n_points samples
// in registers; nothing of size (B, N, n_points) is ever materialised.
#include "common.cuh"

#ifndef XVR_TRI_MIN_CTAS
#define XVR_TRI_MIN_CTAS 3
#endif
#ifndef XVR_TRI_PIPE
#define XVR_TRI_PIPE 4
#endif
#ifndef XVR_TRI_UNROLL
#define XVR_TRI_UNROLL 4
#endif
#define XVR_PRAGMA(x) _Pragma(#x)
#define XVR_UNROLL(n) XVR_PRAGMA(unroll n)

namespace xvr {

struct TrilinearParams {
  Vol vol;
  const uint8_t* __restrict__ labels;  // nullable, same shape as vol
  const uint8_t* __restrict__ label_bricks;  // nullable: per-brick uniform label / 255 (sample_label_bricked)
  int C;                               // output channels (1 without labels)
  const float* __restrict__ source;    // (B,1,3) voxel coords
  const float* __restrict__ target;    // (B,N,3)
  const float* __restrict__ raylen;    // (B,N) world-mm ray length (trainer.py:284)
  bool fused;                          // rays generated in-kernel from `geom` instead of source/target/raylen
  DetectorGeom geom;
  int B, N;
  int n_points;
  int step_mode;  // 0: span/(n-1)   1: span/n   2: 1/n
  float eps;
  TileMap map;
  int tiles_per_pose;
  float* __restrict__ out;  // (B,C,N)
  float* __restrict__ jac;  // (B,7,N): dI/ds(3), dI/dt(3), dI/draylen; nullable
  // backward only
  const float* __restrict__ gout;  // (B,C,N)
  float* __restrict__ gtarget;     // (B,N,3)
  float* __restrict__ gsrc_ray;    // (B,3,N) per-ray source gradient (reduced by reduce_rows)
  float* __restrict__ graylen;     // (B,N)
  float* __restrict__ gvol;        // (D0,D1,D2), accumulated into (ray entry point: RED.ADD scatter), nullable
};

__device__ __forceinline__ float step_weight(int mode, float span, int n) {
  if (mode == 0) return __fdiv_rn(span, (float)(n - 1));
  if (mode == 1) return __fdiv_rn(span, (float)n);
  return __fdiv_rn(1.0f, (float)n);
}

// d(step weight)/d(span)
__device__ __forceinline__ float step_weight_dspan(int mode, int n) {
  if (mode == 0) return 1.0f / (float)(n - 1);
  if (mode == 1) return 1.0f / (float)n;
  return 0.f;
}

// Rays whose segment never comes within one voxel of the volume sample only zero padding: the result is an
// exact 0 and the march can be skipped.
__device__ __forceinline__ bool misses_padded_box(const float s[3], const float d[3], const Vol& v) {
  float lo[3] = {-1.f, -1.f, -1.f};
  float hi[3] = {(float)v.D0, (float)v.D1, (float)v.D2};
  AlphaRange r = alpha_range(s, d, lo, hi);
  return !(r.amin < r.amax);
}

// Empty-space trimming of a ray's sample range [kb, ke): the samples before the ray enters the occupied part of the
// volume and after it leaves it have 8 zero corners -- value 0, gradient 0, an exact no-op for every running sum
// (occupied_alpha_range, common.cuh).  Two samples of margin along the ray for the rounding of alpha_k; margin samples
// are marched as usual and add their zeros: the rendered image and the Jacobian are bit-identical to the full march.
__device__ __forceinline__ void trim_sample_range(const Vol& v, const float s[3], const float d[3], float amin,
                                                  float span, int np, int& kb, int& ke) {
  if (!(span > 0.f) || ke <= kb) return;  // rays marched "backwards" through the padding: rare, left alone
  const float lstep = 1.0f / (float)(np - 1);
  float first = fmaf(lstep * (float)kb, span, amin), last = fmaf(lstep * (float)(ke - 1), span, amin);
  if (!occupied_alpha_range(v, s, d, first, last)) {
    ke = kb;
    return;
  }
  const float sc = (float)(np - 1) / span;
  const float kin = fminf(fmaxf((first - amin) * sc, -4.f), (float)np + 4.f);
  const float kout = fminf(fmaxf((last - amin) * sc, -4.f), (float)np + 4.f);
  kb = max(kb, (int)floorf(kin) - 2);
  ke = min(ke, (int)ceilf(kout) + 3);
  if (ke < kb) ke = kb;
}

template <bool JAC, bool LABELS, bool TEX>
__global__ void __launch_bounds__(256, XVR_TRI_MIN_CTAS) trilinear_fwd_kernel(const TrilinearParams p) {
  extern __shared__ float chan_acc[];  // LABELS: [C][256]
  const int b = blockIdx.x / p.tiles_per_pose;
  const int tile = blockIdx.x - b * p.tiles_per_pose;
  const int tid = threadIdx.x;
  const int n = tile_ray_index(p.map, tile, tid, p.N);
  if (LABELS) {
    for (int c = 0; c < p.C; ++c) chan_acc[c * 256 + tid] = 0.f;
  }
  const int ks = LABELS ? 0 : p.map.ks_log2;
  if (n < 0 && ks == 0) return;
  const bool live = n >= 0;
  const int64_t ray = (int64_t)b * p.N + (live ? n : 0);

  float s[3], d[3], L;
  if (p.fused) {
    generate_ray(p.geom, b, live ? n : 0, p.eps, s, d, L);
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      s[a] = __ldg(p.source + b * 3 + a);
      d[a] = (__ldg(p.target + ray * 3 + a) - s[a]) + p.eps;
    }
    L = __ldg(p.raylen + ray);
  }
  const int np = p.n_points;

  float lo[3] = {0.f, 0.f, 0.f};
  float hi[3] = {(float)(p.vol.D0 - 1), (float)(p.vol.D1 - 1), (float)(p.vol.D2 - 1)};
  const AlphaRange ar = alpha_range(s, d, lo, hi);
  const float span = ar.amax - ar.amin;
  const float w = step_weight(p.step_mode, span, np);

  // Running sums.  With JAC: A = sum grad V, U = sum u grad V; everything else the Jacobian needs is linear in
  // them (alpha_k = amin + u_k span): sum alpha grad V = amin A + span U, sum grad V . d = d . A, sum u grad V . d = d . U.
  float sumV = 0.f;
  float A[3] = {0.f, 0.f, 0.f}, U[3] = {0.f, 0.f, 0.f};

  // The samples this ray marches (empty-space trimming), then this lane's share of them (all of them unless several
  // lanes share the ray): the samples k = part (mod lanes per ray).  Interleaved, not contiguous slices: the lanes of a
  // ray then sample neighbouring positions in the same iteration (their TLD4 footprints share cache lines), every lane
  // gets the same share of the TRIMMED range, and -- the residue classes being absolute -- each lane adds up the same
  // samples in the same order whatever the trimming cut away (bit-identical with and without it).
  const int part = (tid & 31) >> (5 - ks);
  const int stride = 1 << ks;
  int kbeg = 0, kend = np;
  if (p.vol.bbox) trim_sample_range(p.vol, s, d, ar.amin, span, np, kbeg, kend);
  kbeg += (part - kbeg) & (stride - 1);

  if (live && !misses_padded_box(s, d, p.vol)) {
    const float lstep = 1.0f / (float)(np - 1);
    auto sample = [&](float u) {
      const float alpha = fmaf(u, span, ar.amin);
      const float x = fmaf(alpha, d[0], s[0]);
      const float y = fmaf(alpha, d[1], s[1]);
      const float z = fmaf(alpha, d[2], s[2]);
      float g[3];
      const float v = sample_trilinear<JAC, TEX>(p.vol, x, y, z, g);
      if (LABELS) {
        const int c = p.label_bricks ? sample_label_bricked(p.labels, p.label_bricks, p.vol, x, y, z)
                                     : sample_label(p.labels, p.vol, x, y, z);
        chan_acc[c * 256 + tid] += v;
      } else {
        sumV += v;
      }
      if (JAC) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          A[a] += g[a];
          U[a] = fmaf(u, g[a], U[a]);
        }
      }
    };
    // torch.linspace evaluates the two halves of [0,1] from their own ends (linspace01): one loop per half, each
    // with a float counter, instead of a select + int->float conversion per sample
    const int half = np / 2;
    int k = kbeg;
    const int k1 = min(kend, half);
    const float fstride = (float)stride;
    float kf = (float)k;
#if XVR_TRI_PIPE > 0
    if (TEX) {
      // Software pipeline, XVR_TRI_PIPE samples deep.  The march waits on texture LATENCY (long-scoreboard 73 % of the
      // stall samples, the tex queue never full, the pipe 73 % busy), and under the 80-register cap of 3 CTAs per SM the
      // compiler's schedule of the plain unrolled loop keeps only 4 TLD4 per lane in flight.  Here a ring of PIPE slots
      // holds the 8 corners of PIPE samples; a turn of the loop blends all of them and refills them (ptxas groups the
      // refills at the end of the turn: 2 * PIPE gathers in flight per lane).  u and the position are formed again when
      // a slot is consumed (same operations, same bits), and samples are still added up in order of k: bit-identical
      // to the plain loop.  Measured at config 2 (ms per step): plain 7.11, PIPE 3 / 4 / 5 / 6 = 6.89 / 6.74 / 6.76 /
      // 6.83 (from 5 on the slots spill inside the loop).
      constexpr int PIPE = XVR_TRI_PIPE;
      struct Slot { float4 a, b; int ch; };  // ch: label channel of the sample (LABELS only; fetched with the corners)
      auto issue = [&](float u, Slot& q) {
        const float alpha = fmaf(u, span, ar.amin);
        const float x = fmaf(alpha, d[0], s[0]), y = fmaf(alpha, d[1], s[1]), z = fmaf(alpha, d[2], s[2]);
        const float fx0 = floorf(x), fy0 = floorf(y), fz0 = floorf(z);
        const float tu = fz0 + 1.0f, tv = fy0 + 1.0f;
        const unsigned last = (unsigned)(p.vol.D0 + 1);
        const int ix = (int)fx0;
        q.a = gather_yz(p.vol.tex, (int)min((unsigned)(ix + 1), last), tu, tv);
        q.b = gather_yz(p.vol.tex, (int)min((unsigned)(ix + 2), last), tu, tv);
        if (LABELS)
          q.ch = p.label_bricks ? sample_label_bricked(p.labels, p.label_bricks, p.vol, x, y, z)
                                : sample_label(p.labels, p.vol, x, y, z);
      };
      auto consume = [&](float u, const Slot& q) {
        const float alpha = fmaf(u, span, ar.amin);
        const float x = fmaf(alpha, d[0], s[0]), y = fmaf(alpha, d[1], s[1]), z = fmaf(alpha, d[2], s[2]);
        const float fx = x - floorf(x), fy = y - floorf(y), fz = z - floorf(z);
        float g[3];
        const float v = trilinear_interp<JAC>(q.a.w, q.a.z, q.a.x, q.a.y, q.b.w, q.b.z, q.b.x, q.b.y, fx, fy, fz, g);
        if (LABELS) chan_acc[q.ch * 256 + tid] += v;
        else sumV += v;
        if (JAC) {
#pragma unroll
          for (int a = 0; a < 3; ++a) {
            A[a] += g[a];
            U[a] = fmaf(u, g[a], U[a]);
          }
        }
      };
      // one half of the linspace: `cnt` is its float counter, moving by `cstep` per sample, u = ufun(cnt)
      auto run = [&](int kstop, float& cnt, float cstep, auto ufun) {
        if (k + (2 * PIPE - 1) * stride < kstop) {  // enough samples to fill the pipe and turn it over once
          Slot q[PIPE];
#pragma unroll
          for (int i = 0; i < PIPE; ++i) issue(ufun(cnt + (float)i * cstep), q[i]);
          cnt += (float)PIPE * cstep;
          // invariant: q[i] holds sample k + i*stride; cnt is the counter of sample k + PIPE*stride
          for (; k + (2 * PIPE - 1) * stride < kstop; k += PIPE * stride, cnt += (float)PIPE * cstep) {
#pragma unroll
            for (int i = 0; i < PIPE; ++i) {
              consume(ufun(cnt + (float)(i - PIPE) * cstep), q[i]);  // (the counters are small integers: exact)
              issue(ufun(cnt + (float)i * cstep), q[i]);
            }
          }
#pragma unroll
          for (int i = 0; i < PIPE; ++i) consume(ufun(cnt + (float)(i - PIPE) * cstep), q[i]);
          k += PIPE * stride;
        }
        for (; k < kstop; k += stride, cnt += cstep) sample(ufun(cnt));  // fewer than 2 * PIPE left
      };
      run(k1, kf, fstride, [&](float c) { return lstep * c; });
      float rf = (float)(np - 1 - k);
      run(kend, rf, -fstride, [&](float c) { return linspace_tail(lstep, c); });
    } else
#endif
    {
    XVR_UNROLL(XVR_TRI_UNROLL)
    for (; k < k1; k += stride, kf += fstride) sample(lstep * kf);
    float rf = (float)(np - 1 - k);
    XVR_UNROLL(XVR_TRI_UNROLL)
    for (; k < kend; k += stride, rf -= fstride) sample(linspace_tail(lstep, rf));
    }
  }

  if (ks > 0) {  // combine the slices (fixed order: deterministic); slice 0 writes
    sumV = ksplit_sum(sumV, ks);
    if (JAC) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        A[a] = ksplit_sum(A[a], ks);
        U[a] = ksplit_sum(U[a], ks);
      }
    }
    if (!live || part != 0) return;
  }

  if (LABELS) {
    for (int c = 0; c < p.C; ++c) {
      const float sv = chan_acc[c * 256 + tid];
      p.out[((int64_t)b * p.C + c) * p.N + n] = sv * L * w;
      sumV += sv;
    }
  } else {
    p.out[ray] = sumV * L * w;
  }

  if (JAC) {
    // I = L * w(span) * sumV,  x_k = s + alpha_k d,  alpha_k = amin + u_k span,  d = t - s + eps.
    // dI/ds = L [ sumV w' (dspan/ds) + w (A - Bv + P damin/ds + Q damax/ds) ],  P = T - Q
    // dI/dt = L [ sumV w' (dspan/dt) + w (     Bv + P damin/dt + Q damax/dt) ]
    // a crossing alpha = (plane - s_a)/d_a has dalpha/ds_a = (alpha-1)/d_a, dalpha/dt_a = -alpha/d_a.
    const float T = fmaf(A[0], d[0], fmaf(A[1], d[1], A[2] * d[2]));
    const float Q = fmaf(U[0], d[0], fmaf(U[1], d[1], U[2] * d[2]));
    const float P = T - Q;
    const float wp = step_weight_dspan(p.step_mode, np);
    float js[3], jt[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float Bv = fmaf(ar.amin, A[a], span * U[a]);
      js[a] = w * (A[a] - Bv);
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
    float* j = p.jac + (int64_t)b * 7 * p.N + n;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      j[(int64_t)a * p.N] = L * js[a];
      j[(int64_t)(3 + a) * p.N] = L * jt[a];
    }
    j[(int64_t)6 * p.N] = sumV * w;
  }
}

// Recompute backward: given dL/dout (B,C,N) re-march every ray and emit dL/dtarget (B,N,3), the per-ray
// dL/dsource (B,3,N) and dL/draylen (B,N).  Handles label channels (the upstream gradient of a sample is the
// one of the channel its label selects).
template <bool LABELS, bool TEX, bool VOLGRAD>
__global__ void __launch_bounds__(256) trilinear_bwd_kernel(const TrilinearParams p) {
  extern __shared__ float chan_g[];  // LABELS: [C][256] upstream gradient per channel
  const int b = blockIdx.x / p.tiles_per_pose;
  const int tile = blockIdx.x - b * p.tiles_per_pose;
  const int tid = threadIdx.x;
  const int n = tile_ray_index(p.map, tile, tid, p.N);
  if (n < 0) return;
  const int64_t ray = (int64_t)b * p.N + n;

  float s[3], d[3], L;
  if (p.fused) {
    generate_ray(p.geom, b, n, p.eps, s, d, L);
  } else {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      s[a] = __ldg(p.source + b * 3 + a);
      d[a] = (__ldg(p.target + ray * 3 + a) - s[a]) + p.eps;
    }
    L = __ldg(p.raylen + ray);
  }
  const int np = p.n_points;
  float g1 = 0.f;
  if (LABELS) {
    for (int c = 0; c < p.C; ++c) chan_g[c * 256 + tid] = __ldg(p.gout + ((int64_t)b * p.C + c) * p.N + n);
  } else {
    g1 = __ldg(p.gout + ray);
  }

  float lo[3] = {0.f, 0.f, 0.f};
  float hi[3] = {(float)(p.vol.D0 - 1), (float)(p.vol.D1 - 1), (float)(p.vol.D2 - 1)};
  const AlphaRange ar = alpha_range(s, d, lo, hi);
  const float span = ar.amax - ar.amin;
  const float w = step_weight(p.step_mode, span, np);

  // all sums carry the upstream gradient of the sample's channel
  float sumV = 0.f;
  float A[3] = {0.f, 0.f, 0.f}, Bv[3] = {0.f, 0.f, 0.f}, T = 0.f, Q = 0.f;
  if (!misses_padded_box(s, d, p.vol)) {
    const float lstep = 1.0f / (float)(np - 1);
#pragma unroll 4
    for (int k = 0; k < np; ++k) {
      const float u = linspace01(k, np, lstep);
      const float alpha = fmaf(u, span, ar.amin);
      const float x = fmaf(alpha, d[0], s[0]);
      const float y = fmaf(alpha, d[1], s[1]);
      const float z = fmaf(alpha, d[2], s[2]);
      float g[3];
      const float v = sample_trilinear<true, TEX>(p.vol, x, y, z, g);
      float go = g1;
      if (LABELS) go = chan_g[sample_label(p.labels, p.vol, x, y, z) * 256 + tid];
      if (VOLGRAD) scatter_trilinear(p.gvol, p.vol, x, y, z, go * L * w);
      sumV = fmaf(go, v, sumV);
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float ga = go * g[a];
        g[a] = ga;
        A[a] += ga;
        Bv[a] = fmaf(alpha, ga, Bv[a]);
      }
      const float gd = fmaf(g[0], d[0], fmaf(g[1], d[1], g[2] * d[2]));
      T += gd;
      Q = fmaf(u, gd, Q);
    }
  }
  const float P = T - Q;
  const float wp = step_weight_dspan(p.step_mode, np);
  float js[3], jt[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    js[a] = w * (A[a] - Bv[a]);
    jt[a] = w * Bv[a];
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
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    p.gtarget[ray * 3 + a] = L * jt[a];
    p.gsrc_ray[((int64_t)b * 3 + a) * p.N + n] = L * js[a];
  }
  p.graylen[ray] = sumV * w;
}

// Backward through a saved Jacobian: gtarget = g * J_t, gsrc_ray = g * J_s, graylen = g * J_L.
__global__ void __launch_bounds__(256)
jac_bwd_kernel(const float* __restrict__ jac, const float* __restrict__ gout, int B, int N,
               float* __restrict__ gtarget, float* __restrict__ gsrc_ray, float* __restrict__ graylen) {
  const int64_t ray = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (ray >= (int64_t)B * N) return;
  const int b = (int)(ray / N);
  const int n = (int)(ray - (int64_t)b * N);
  const float g = __ldg(gout + ray);
  const float* j = jac + (int64_t)b * 7 * N + n;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    gsrc_ray[((int64_t)b * 3 + a) * N + n] = g * __ldg(j + (int64_t)a * N);
    gtarget[ray * 3 + a] = g * __ldg(j + (int64_t)(3 + a) * N);
  }
  graylen[ray] = g * __ldg(j + (int64_t)6 * N);
}

// Deterministic row sums: out[r] = sum_n in[r, n]; one CTA per row, fixed summation tree.
__global__ void __launch_bounds__(1024) reduce_rows_kernel(const float* __restrict__ in, int N,
                                                           float* __restrict__ out) {
  __shared__ float part[32];
  const float* row = in + (int64_t)blockIdx.x * N;
  float acc = 0.f;
  for (int n = threadIdx.x; n < N; n += 1024) acc += __ldg(row + n);
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = part[threadIdx.x];
    v = warp_sum(v);
    if (threadIdx.x == 0) out[blockIdx.x] = v;
  }
}

// Fused-path backward: dL/dG (B,3,4) from the saved per-ray Jacobian.  t_r = G [c_r; 1], s = G[:,3]  =>
// dL/dG[:, :3] = sum_r g_r J_t,r (x) c_r,  dL/dG[:, 3] = sum_r g_r (J_t,r + J_s,r).  gridDim.y CTAs share a pose
// (each reduces a contiguous slice of its rays with a fixed tree and writes 12 partial sums); a second pass adds the
// slices in order, so the result is deterministic for a given launch shape.
__global__ void __launch_bounds__(1024)
drr_jac_bwd_kernel(const float* __restrict__ jac, const float* __restrict__ gout, int N, DetectorGeom geom,
                   float* __restrict__ partial) {
  __shared__ float part[32][12];
  const int b = blockIdx.x, slice = blockIdx.y, S = gridDim.y;
  const int per = (N + S - 1) / S;
  const int n0 = slice * per, n1 = min(N, n0 + per);
  float acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.f;
  const float* j = jac + (int64_t)b * 7 * N;
  for (int n = n0 + threadIdx.x; n < n1; n += 1024) {
    const float g = __ldg(gout + (int64_t)b * N + n);
    float c[3];
    camera_point(geom, n, c);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float gt = g * __ldg(j + (int64_t)(3 + a) * N + n);
      const float gs = g * __ldg(j + (int64_t)a * N + n);
      acc[a * 4 + 0] = fmaf(gt, c[0], acc[a * 4 + 0]);
      acc[a * 4 + 1] = fmaf(gt, c[1], acc[a * 4 + 1]);
      acc[a * 4 + 2] = fmaf(gt, c[2], acc[a * 4 + 2]);
      acc[a * 4 + 3] += gt + gs;
    }
  }
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    const float v = warp_sum(acc[k]);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      const float v = warp_sum(part[threadIdx.x][k]);
      if (threadIdx.x == 0) partial[((int64_t)b * S + slice) * 12 + k] = v;
    }
  }
}

__global__ void drr_jac_bwd_finish_kernel(const float* __restrict__ partial, int B, int S, float* __restrict__ gG) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // one thread per (pose, matrix entry)
  if (i >= B * 12) return;
  const int b = i / 12, k = i - b * 12;
  float v = 0.f;
  for (int s = 0; s < S; ++s) v += partial[((int64_t)b * S + s) * 12 + k];
  gG[i] = v;
}

static int fill_geom(DetectorGeom& g, const float* cam2vox, const float* cam2world, const float* det9, int W) {
  if (!cam2vox || !cam2world || !det9 || W <= 0) {
    set_last_error("xvr_drr: null geometry argument");
    return XVR_ERR_INVALID;
  }
  g.cam2vox = cam2vox;
  g.cam2world = cam2world;
  for (int a = 0; a < 3; ++a) {
    g.o[a] = det9[a];
    g.u[a] = det9[3 + a];
    g.v[a] = det9[6 + a];
  }
  g.W = W;
  return XVR_OK;
}

static int fill_map(TileMap& m, int N, int det_h, int det_w, int lane_w_log2, int cta_w_log2, int ks_log2,
                    int* tiles) {
  if (det_w > 0 && det_h > 0) {
    if ((int64_t)det_h * det_w != N) {
      set_last_error("detector hint H*W != N");
      return XVR_ERR_INVALID;
    }
    if (lane_w_log2 < 0 || lane_w_log2 > 5 || cta_w_log2 < lane_w_log2 || cta_w_log2 > 8 ||
        (8 - cta_w_log2) < (5 - lane_w_log2)) {
      set_last_error("invalid tile shape");
      return XVR_ERR_INVALID;
    }
    if (ks_log2 > 5 - lane_w_log2) ks_log2 = 5 - lane_w_log2;
    m.W = det_w;
    m.H = det_h;
    m.lane_w_log2 = lane_w_log2;
    m.cta_w_log2 = cta_w_log2;
    m.ks_log2 = ks_log2;
    const int tw = 1 << cta_w_log2, th = (256 >> cta_w_log2) >> ks_log2;
    m.tiles_x = (det_w + tw - 1) / tw;
    m.tiles_y = (det_h + th - 1) / th;
    *tiles = m.tiles_x * m.tiles_y;
  } else {
    m.W = 0;
    m.H = 0;
    m.lane_w_log2 = 5;
    m.cta_w_log2 = 8;
    m.ks_log2 = ks_log2;
    m.tiles_x = (N + (256 >> ks_log2) - 1) / (256 >> ks_log2);
    m.tiles_y = 1;
    *tiles = m.tiles_x;
  }
  return XVR_OK;
}

static int fill_common(TrilinearParams& p, const float* volume, const void* voltex, int D0, int D1, int D2,
                       const uint8_t* labels,
                       int C, const float* source, const float* target, const float* raylen, int B, int N,
                       int n_points, int step_mode, float eps, int det_h, int det_w, int lane_w_log2,
                       int cta_w_log2, int opts, bool allow_ksplit = true) {
  if (!volume || ((!source || !target || !raylen) && !p.fused) || B <= 0 || N <= 0 || D0 < 2 || D1 < 2 || D2 < 2 ||
      n_points < 2 || n_points >= (1 << 26) || step_mode < 0 || step_mode > 2 || C < 1 || (labels && C > 255) || (!labels && C != 1) ||
      (opts & ~XVR_OPT_KNOWN) || (opts & XVR_OPT_KSPLIT_MASK) > 4) {
    set_last_error("xvr_trilinear: invalid argument");
    return XVR_ERR_INVALID;
  }
  if ((int64_t)D0 * D1 * D2 >= (int64_t)1 << 31) {
    set_last_error("xvr_trilinear: volume too large for 32-bit voxel offsets");
    return XVR_ERR_INVALID;
  }
  p.vol.data = volume;
  p.vol.D0 = D0;
  p.vol.D1 = D1;
  p.vol.D2 = D2;
  p.vol.s0 = D1 * D2;
  p.vol.s1 = D2;
  p.vol.tex = 0;
  p.vol.bbox = nullptr;
  if (voltex) {
    const VolumeTexture* vt = (const VolumeTexture*)voltex;
    if (vt->D0 != D0 || vt->D1 != D1 || vt->D2 != D2) {
      set_last_error("xvr_trilinear: volume texture shape differs from the volume");
      return XVR_ERR_INVALID;
    }
    p.vol.tex = vt->tex;
    if (!(opts & XVR_OPT_NO_TRIM)) {  // empty-space trimming: the handle knows where its non-zero voxels are
      p.vol.bbox = vt->bbox;
      p.vol.occ = vt->occ;
      p.vol.nb0 = vt->nb0;
      p.vol.nb1 = vt->nb1;
      p.vol.nb2 = vt->nb2;
    }
  }
  p.labels = labels;
  // XVR_OPT_LABEL_BRICKS: the caller's label buffer continues, at the next multiple of 256 bytes after the D0*D1*D2 label
  // bytes, with the (nb0,nb1,nb2) brick table of sample_label_bricked (nb = ceil(D / 8))
  p.label_bricks = (labels && (opts & XVR_OPT_LABEL_BRICKS))
                       ? labels + (((size_t)D0 * D1 * D2 + 255) / 256) * 256 : nullptr;
  p.C = C;
  p.source = source;
  p.target = target;
  p.raylen = raylen;
  p.B = B;
  p.N = N;
  p.n_points = n_points;
  p.step_mode = step_mode;
  p.eps = eps;
  // Small batches (B = 1 registration): let 2, 4 or 8 lanes share a ray (interleaved samples of the trimmed range).  The
  // kernel keeps 3 CTAs per SM resident, 444 in all: take the FEWEST lanes per ray that still give every slot a CTA.
  // (Round 1 picked the split whose last wave was fullest; with trimmed rays a CTA's work varies from nothing to a full
  // march, waves mean little, and every extra lane per ray costs texture locality: config 3 at 256^2, 1 / 2 / 4 / 8
  // lanes per ray = under-filled / 0.250 / 0.252 / 0.266 ms per iteration.)
  int ks = 0;
  if (!labels && allow_ksplit) {
    if (opts & XVR_OPT_KSPLIT_MASK) {
      ks = (opts & XVR_OPT_KSPLIT_MASK) - 1;
    } else {
      const int64_t slots = 148 * XVR_TRI_MIN_CTAS;
      while (ks < 3 && ((((int64_t)B * N) << ks) + 255) / 256 < slots) ++ks;
    }
  }
  return fill_map(p.map, N, det_h, det_w, lane_w_log2, cta_w_log2, ks, &p.tiles_per_pose);
}

}  // namespace xvr

using namespace xvr;

extern "C" int xvr_trilinear_rays_fwd(const float* volume, const void* voltex, int D0, int D1, int D2,
                                      const uint8_t* labels, int C,
                                      const float* source, const float* target, const float* raylen, int B,
                                      int N, int n_points, int step_mode, float eps, int det_h, int det_w,
                                      int lane_w_log2, int cta_w_log2, float* out, float* jac, int opts,
                                      void* stream) {
  TrilinearParams p = {};
  int rc = fill_common(p, volume, voltex, D0, D1, D2, labels, C, source, target, raylen, B, N, n_points, step_mode, eps,
                       det_h, det_w, lane_w_log2, cta_w_log2, opts);
  if (rc) return rc;
  if (!out) {
    set_last_error("xvr_trilinear_rays_fwd: out is null");
    return XVR_ERR_INVALID;
  }
  p.out = out;
  p.jac = jac;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t grid = (int64_t)B * p.tiles_per_pose;
  if (grid >= (int64_t)1 << 31) {
    set_last_error("xvr_trilinear_rays_fwd: grid too large");
    return XVR_ERR_INVALID;
  }
  const size_t smem = labels ? (size_t)C * 256 * sizeof(float) : 0;
  const bool tex = p.vol.tex != 0;
  if (labels) {
    // with jac: the Jacobian of the channel SUM (what a caller that collapses the channels differentiates)
    auto k = jac ? (tex ? trilinear_fwd_kernel<true, true, true> : trilinear_fwd_kernel<true, true, false>)
                 : (tex ? trilinear_fwd_kernel<false, true, true> : trilinear_fwd_kernel<false, true, false>);
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<(unsigned)grid, 256, smem, st>>>(p);
  } else if (jac) {
    auto k = tex ? trilinear_fwd_kernel<true, false, true> : trilinear_fwd_kernel<true, false, false>;
    k<<<(unsigned)grid, 256, 0, st>>>(p);
  } else {
    auto k = tex ? trilinear_fwd_kernel<false, false, true> : trilinear_fwd_kernel<false, false, false>;
    k<<<(unsigned)grid, 256, 0, st>>>(p);
  }
  return check_launch("xvr_trilinear_rays_fwd");
}

extern "C" int xvr_trilinear_rays_bwd(const float* volume, const void* voltex, int D0, int D1, int D2,
                                      const uint8_t* labels, int C,
                                      const float* source, const float* target, const float* raylen, int B,
                                      int N, int n_points, int step_mode, float eps, int det_h, int det_w,
                                      int lane_w_log2, int cta_w_log2, const float* gout, float* gsource,
                                      float* gtarget, float* graylen, float* workspace, float* gvol, void* stream) {
  TrilinearParams p = {};
  int rc = fill_common(p, volume, voltex, D0, D1, D2, labels, C, source, target, raylen, B, N, n_points, step_mode, eps,
                       det_h, det_w, lane_w_log2, cta_w_log2, 0, false);
  if (rc) return rc;
  if (!gout || !gsource || !gtarget || !graylen || !workspace) {
    set_last_error("xvr_trilinear_rays_bwd: null gradient buffer");
    return XVR_ERR_INVALID;
  }
  p.gout = gout;
  p.gtarget = gtarget;
  p.gsrc_ray = workspace;  // (B,3,N)
  p.graylen = graylen;
  p.gvol = gvol;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t grid = (int64_t)B * p.tiles_per_pose;
  const size_t smem = labels ? (size_t)C * 256 * sizeof(float) : 0;
  const bool tex = p.vol.tex != 0;
  auto launch = [&](auto k) {
    if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    k<<<(unsigned)grid, 256, smem, st>>>(p);
  };
  if (labels) {
    if (gvol) { if (tex) launch(trilinear_bwd_kernel<true, true, true>); else launch(trilinear_bwd_kernel<true, false, true>); }
    else { if (tex) launch(trilinear_bwd_kernel<true, true, false>); else launch(trilinear_bwd_kernel<true, false, false>); }
  } else {
    if (gvol) { if (tex) launch(trilinear_bwd_kernel<false, true, true>); else launch(trilinear_bwd_kernel<false, false, true>); }
    else { if (tex) launch(trilinear_bwd_kernel<false, true, false>); else launch(trilinear_bwd_kernel<false, false, false>); }
  }
  rc = check_launch("xvr_trilinear_rays_bwd");
  if (rc) return rc;
  reduce_rows_kernel<<<B * 3, 1024, 0, st>>>(workspace, N, gsource);
  return check_launch("xvr_trilinear_rays_bwd/reduce");
}

// out[r] = sum_n in[r, n] with a fixed summation tree (deterministic); shared by the Siddon entry points.
extern "C" int xvr_reduce_rows(const float* in, int rows, int N, float* out, void* stream) {
  if (!in || !out || rows <= 0 || N <= 0) {
    set_last_error("xvr_reduce_rows: invalid argument");
    return XVR_ERR_INVALID;
  }
  reduce_rows_kernel<<<rows, 1024, 0, (cudaStream_t)stream>>>(in, N, out);
  return check_launch("xvr_reduce_rows");
}

extern "C" int xvr_rays_jac_bwd(const float* jac, const float* gout, int B, int N, float* gsource,
                                float* gtarget, float* graylen, float* workspace, void* stream) {
  if (!jac || !gout || !gsource || !gtarget || !graylen || !workspace || B <= 0 || N <= 0) {
    set_last_error("xvr_rays_jac_bwd: invalid argument");
    return XVR_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t rays = (int64_t)B * N;
  jac_bwd_kernel<<<(unsigned)((rays + 255) / 256), 256, 0, st>>>(jac, gout, B, N, gtarget, workspace, graylen);
  int rc = check_launch("xvr_rays_jac_bwd");
  if (rc) return rc;
  reduce_rows_kernel<<<B * 3, 1024, 0, st>>>(workspace, N, gsource);
  return check_launch("xvr_rays_jac_bwd/reduce");
}

// Fused DRR forward: rays are generated in-kernel from the per-pose camera->voxel matrix and the detector basis
// (det9 = origin, row step, column step of the pixel grid in the camera frame; a HOST array of 9 floats).
extern "C" int xvr_trilinear_drr_fwd(const float* volume, const void* voltex, int D0, int D1, int D2,
                                     const float* cam2vox, const float* cam2world, const float* det9, int B,
                                     int det_h, int det_w, int n_points, int step_mode, float eps, int lane_w_log2,
                                     int cta_w_log2, float* out, float* jac, int opts, void* stream) {
  TrilinearParams p = {};
  p.fused = true;
  int rc = fill_geom(p.geom, cam2vox, cam2world, det9, det_w);
  if (rc) return rc;
  rc = fill_common(p, volume, voltex, D0, D1, D2, nullptr, 1, nullptr, nullptr, nullptr, B, det_h * det_w, n_points,
                   step_mode, eps, det_h, det_w, lane_w_log2, cta_w_log2, opts);
  if (rc) return rc;
  if (!out) {
    set_last_error("xvr_trilinear_drr_fwd: out is null");
    return XVR_ERR_INVALID;
  }
  p.out = out;
  p.jac = jac;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t grid = (int64_t)B * p.tiles_per_pose;
  if (grid >= (int64_t)1 << 31) {
    set_last_error("xvr_trilinear_drr_fwd: grid too large");
    return XVR_ERR_INVALID;
  }
  const bool tex = p.vol.tex != 0;
  if (jac) {
    auto k = tex ? trilinear_fwd_kernel<true, false, true> : trilinear_fwd_kernel<true, false, false>;
    k<<<(unsigned)grid, 256, 0, st>>>(p);
  } else {
    auto k = tex ? trilinear_fwd_kernel<false, false, true> : trilinear_fwd_kernel<false, false, false>;
    k<<<(unsigned)grid, 256, 0, st>>>(p);
  }
  return check_launch("xvr_trilinear_drr_fwd");
}

// The fused DRR with label channels = DRR.forward(..., mask_to_channels=True) / trainer.py:283-289 with mask=seg: rays
// generated in the kernel, out (B,C,H*W), jac (B,7,H*W) = the Jacobian of the channel SUM or NULL (what a caller that
// collapses the channels differentiates, trainer.py:294; backward = xvr_drr_jac_bwd on the summed upstream gradient).
// labels / C / XVR_OPT_LABEL_BRICKS as in xvr_trilinear_rays_fwd.
extern "C" int xvr_trilinear_drr_fwd_labels(const float* volume, const void* voltex, int D0, int D1, int D2,
                                            const uint8_t* labels, int C, const float* cam2vox,
                                            const float* cam2world, const float* det9, int B, int det_h, int det_w,
                                            int n_points, int step_mode, float eps, int lane_w_log2, int cta_w_log2,
                                            float* out, float* jac, int opts, void* stream) {
  TrilinearParams p = {};
  p.fused = true;
  int rc = fill_geom(p.geom, cam2vox, cam2world, det9, det_w);
  if (rc) return rc;
  if (!labels) {
    set_last_error("xvr_trilinear_drr_fwd_labels: labels is null (use xvr_trilinear_drr_fwd)");
    return XVR_ERR_INVALID;
  }
  rc = fill_common(p, volume, voltex, D0, D1, D2, labels, C, nullptr, nullptr, nullptr, B, det_h * det_w, n_points,
                   step_mode, eps, det_h, det_w, lane_w_log2, cta_w_log2, opts);
  if (rc) return rc;
  if (!out) {
    set_last_error("xvr_trilinear_drr_fwd_labels: out is null");
    return XVR_ERR_INVALID;
  }
  p.out = out;
  p.jac = jac;
  const int64_t grid = (int64_t)B * p.tiles_per_pose;
  if (grid >= (int64_t)1 << 31) {
    set_last_error("xvr_trilinear_drr_fwd_labels: grid too large");
    return XVR_ERR_INVALID;
  }
  const size_t smem = (size_t)C * 256 * sizeof(float);
  const bool tex = p.vol.tex != 0;
  auto k = jac ? (tex ? trilinear_fwd_kernel<true, true, true> : trilinear_fwd_kernel<true, true, false>)
               : (tex ? trilinear_fwd_kernel<false, true, true> : trilinear_fwd_kernel<false, true, false>);
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k<<<(unsigned)grid, 256, smem, (cudaStream_t)stream>>>(p);
  return check_launch("xvr_trilinear_drr_fwd_labels");
}

namespace xvr {
__global__ void __launch_bounds__(256) trilinear_count_kernel(const TrilinearParams p, unsigned long long* counter) {
  const int64_t ray = (int64_t)blockIdx.x * 256 + threadIdx.x;
  unsigned long long mine = 0;
  if (ray < (int64_t)p.B * p.N) {
    const int b = (int)(ray / p.N), n = (int)(ray - (int64_t)b * p.N);
    float s[3], d[3], L;
    generate_ray(p.geom, b, n, p.eps, s, d, L);
    const float lo[3] = {0.f, 0.f, 0.f};
    const float hi[3] = {(float)(p.vol.D0 - 1), (float)(p.vol.D1 - 1), (float)(p.vol.D2 - 1)};
    const AlphaRange ar = alpha_range(s, d, lo, hi);
    if (!misses_padded_box(s, d, p.vol)) {
      int kb = 0, ke = p.n_points;
      if (p.vol.bbox) trim_sample_range(p.vol, s, d, ar.amin, ar.amax - ar.amin, p.n_points, kb, ke);
      mine = (unsigned long long)(ke - kb);
    }
  }
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(counter, mine);
}
}  // namespace xvr

// Number of samples xvr_trilinear_drr_fwd marches for this batch (bench.py: the executed share under empty-space
// trimming; rays that miss the padded volume march none).  counter: DEVICE unsigned long long the caller zeroes.
extern "C" int xvr_trilinear_drr_count(const void* voltex, int D0, int D1, int D2, const float* cam2vox,
                                       const float* cam2world, const float* det9, int B, int det_h, int det_w,
                                       int n_points, float eps, unsigned long long* counter, int opts, void* stream) {
  if (!counter || B <= 0 || det_h <= 0 || det_w <= 0 || D0 < 2 || D1 < 2 || D2 < 2 || n_points < 2 ||
      (opts & ~XVR_OPT_KNOWN)) {
    set_last_error("xvr_trilinear_drr_count: invalid argument");
    return XVR_ERR_INVALID;
  }
  TrilinearParams p = {};
  p.fused = true;
  int rc = fill_geom(p.geom, cam2vox, cam2world, det9, det_w);
  if (rc) return rc;
  p.vol.D0 = D0;
  p.vol.D1 = D1;
  p.vol.D2 = D2;
  if (voltex && !(opts & XVR_OPT_NO_TRIM)) {
    const VolumeTexture* vt = (const VolumeTexture*)voltex;
    p.vol.bbox = vt->bbox;
    p.vol.occ = vt->occ;
    p.vol.nb0 = vt->nb0;
    p.vol.nb1 = vt->nb1;
    p.vol.nb2 = vt->nb2;
  }
  p.B = B;
  p.N = det_h * det_w;
  p.n_points = n_points;
  p.eps = eps;
  const int64_t rays = (int64_t)B * p.N;
  trilinear_count_kernel<<<(unsigned)((rays + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p, counter);
  return check_launch("xvr_trilinear_drr_count");
}

// Backward of any fused DRR forward that saved its per-ray Jacobian: gG (B,3,4) = dL/d(cam2vox).
// workspace: NULL (one CTA per pose) or >= 12 * B * xvr_drr_jac_bwd_slices(B, det_h * det_w) floats: small batches
// (registration, B = 1) then spread each pose over several CTAs.
extern "C" int xvr_drr_jac_bwd_slices(int B, int N) {
  if (B <= 0 || N <= 0) return 1;
  int S = (2 * 148 + B - 1) / B;             // about two CTAs per SM in total
  const int most = (N + 2047) / 2048;         // at least two rays per thread and slice
  if (S > most) S = most;
  if (S > 64) S = 64;
  return S < 1 ? 1 : S;
}

extern "C" int xvr_drr_jac_bwd(const float* jac, const float* gout, const float* det9, int B, int det_h, int det_w,
                               float* gG, float* workspace, void* stream) {
  if (!jac || !gout || !gG || B <= 0 || det_h <= 0) {
    set_last_error("xvr_drr_jac_bwd: invalid argument");
    return XVR_ERR_INVALID;
  }
  DetectorGeom g = {};
  int rc = fill_geom(g, jac, jac, det9, det_w);  // the matrices are not read by the reduction
  if (rc) return rc;
  const int N = det_h * det_w;
  const int S = workspace ? xvr_drr_jac_bwd_slices(B, N) : 1;
  cudaStream_t st = (cudaStream_t)stream;
  if (S == 1) {
    drr_jac_bwd_kernel<<<dim3(B, 1), 1024, 0, st>>>(jac, gout, N, g, gG);
    return check_launch("xvr_drr_jac_bwd");
  }
  drr_jac_bwd_kernel<<<dim3(B, S), 1024, 0, st>>>(jac, gout, N, g, workspace);
  rc = check_launch("xvr_drr_jac_bwd");
  if (rc) return rc;
  drr_jac_bwd_finish_kernel<<<(B * 12 + 127) / 128, 128, 0, st>>>(workspace, B, S, gG);
  return check_launch("xvr_drr_jac_bwd/finish");
}
