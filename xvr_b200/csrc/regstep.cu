// The scalar tail of one registration iteration in one launch each: parameters -> camera matrices (and back), and
// optimiser + learning-rate scheduler + trajectory log.
//
// xvr's loop (/root/reference/src/xvr/registrar/base.py:245-278) runs, per iteration, a few hundred launches that
// each touch a handful of floats: convert(rot, xyz) and the compose / affine-inverse chain in front of the
// renderer, their autograd mirror image behind it, torch.optim.Adam(maximize=True) on two 3-vectors,
// ReduceLROnPlateau and the bookkeeping of the stopping rule.  At B = 1 the renderer needs ~0.25 ms; those
// launches cost more than that even inside a CUDA graph.  Here each group is a single one-thread-per-pose kernel.
#include "common.cuh"

namespace xvr {

// ------------------------------------------------------------------------------------------ Euler pose -> camera
struct EulerCameraParams {
  const float* __restrict__ rot;  // (B,3) Euler angles, radians
  const float* __restrict__ xyz;  // (B,3)
  int B;
  int axis[3];           // 0/1/2 = X/Y/Z of the convention string, R = R_axis0(a0) R_axis1(a1) R_axis2(a2)
  int rotated_frame;     // 1: matrix = [R | R xyz] (DiffDRR's convert), 0: [R | xyz]
  float reorient[16];    // R0: camera frame -> pose frame (row-major 4x4)
  float affinv[16];      // world mm -> voxel index (row-major 4x4)
  float* __restrict__ cam2world;  // (B,3,4) = (P R0)[:3]
  float* __restrict__ cam2vox;    // (B,3,4) = (Ainv P R0)[:3]
  // backward
  const float* __restrict__ gcam2vox;  // (B,3,4)
  float* __restrict__ grot;            // (B,3)
  float* __restrict__ gxyz;            // (B,3)
};

// elementary rotation about `axis` and its derivative w.r.t. the angle (row-major 3x3)
__device__ __forceinline__ void elementary(int axis, float angle, float R[9], float dR[9]) {
  float s, c;
  sincosf(angle, &s, &c);
  for (int k = 0; k < 9; ++k) { R[k] = 0.f; dR[k] = 0.f; }
  if (axis == 0) {
    R[0] = 1.f; R[4] = c; R[5] = -s; R[7] = s; R[8] = c;
    dR[4] = -s; dR[5] = -c; dR[7] = c; dR[8] = -s;
  } else if (axis == 1) {
    R[0] = c; R[2] = s; R[4] = 1.f; R[6] = -s; R[8] = c;
    dR[0] = -s; dR[2] = c; dR[6] = -c; dR[8] = -s;
  } else {
    R[0] = c; R[1] = -s; R[3] = s; R[4] = c; R[8] = 1.f;
    dR[0] = -s; dR[1] = -c; dR[3] = c; dR[4] = -s;
  }
}

__device__ __forceinline__ void mul33(const float A[9], const float B[9], float C[9]) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = fmaf(A[i * 3], B[j], fmaf(A[i * 3 + 1], B[3 + j], A[i * 3 + 2] * B[6 + j]));
}

template <bool BACKWARD>
__global__ void euler_camera_kernel(const EulerCameraParams p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  float E[3][9], dE[3][9];
  for (int k = 0; k < 3; ++k) elementary(p.axis[k], p.rot[b * 3 + k], E[k], dE[k]);
  float R01[9], R[9], R12[9];
  mul33(E[0], E[1], R01);
  mul33(R01, E[2], R);
  const float t_in[3] = {p.xyz[b * 3], p.xyz[b * 3 + 1], p.xyz[b * 3 + 2]};
  if (!BACKWARD) {
    float P[12];  // top three rows of the pose matrix
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) P[i * 4 + j] = R[i * 3 + j];
      P[i * 4 + 3] = p.rotated_frame ? fmaf(R[i * 3], t_in[0], fmaf(R[i * 3 + 1], t_in[1], R[i * 3 + 2] * t_in[2])) : t_in[i];
    }
    // M = P R0 (rows 0..2; row 3 of both factors is 0 0 0 1)
    float M[12];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j)
        M[i * 4 + j] = fmaf(P[i * 4], p.reorient[j], fmaf(P[i * 4 + 1], p.reorient[4 + j],
                            fmaf(P[i * 4 + 2], p.reorient[8 + j], P[i * 4 + 3] * p.reorient[12 + j])));
    for (int k = 0; k < 12; ++k) p.cam2world[b * 12 + k] = M[k];
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j)
        p.cam2vox[b * 12 + i * 4 + j] = fmaf(p.affinv[i * 4], M[j], fmaf(p.affinv[i * 4 + 1], M[4 + j],
                                             fmaf(p.affinv[i * 4 + 2], M[8 + j], j == 3 ? p.affinv[i * 4 + 3] : 0.f)));
    return;
  }
  // G = A3 P4 R0 with A3 = affinv[:3, :], P4 = [P; 0 0 0 1]  =>  dL/dP[:3, :] = (A3[:, :3])^T gG R0^T
  const float* g = p.gcam2vox + b * 12;
  float T[12];  // (A3[:, :3])^T gG   (3x4)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j)
      T[i * 4 + j] = fmaf(p.affinv[i], g[j], fmaf(p.affinv[4 + i], g[4 + j], p.affinv[8 + i] * g[8 + j]));
  float gP[12];  // T R0^T
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j)
      gP[i * 4 + j] = fmaf(T[i * 4], p.reorient[j * 4], fmaf(T[i * 4 + 1], p.reorient[j * 4 + 1],
                           fmaf(T[i * 4 + 2], p.reorient[j * 4 + 2], T[i * 4 + 3] * p.reorient[j * 4 + 3])));
  float gR[9], gt[3] = {gP[3], gP[7], gP[11]};
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) gR[i * 3 + j] = gP[i * 4 + j] + (p.rotated_frame ? gt[i] * t_in[j] : 0.f);
  for (int j = 0; j < 3; ++j)
    p.gxyz[b * 3 + j] = p.rotated_frame ? fmaf(R[j], gt[0], fmaf(R[3 + j], gt[1], R[6 + j] * gt[2])) : gt[j];
  // dR/da0 = dE0 E1 E2, dR/da1 = E0 dE1 E2, dR/da2 = E0 E1 dE2
  mul33(E[1], E[2], R12);
  float D[9], tmp[9];
  float ga[3];
  mul33(dE[0], R12, D);
  ga[0] = 0.f;
  for (int k = 0; k < 9; ++k) ga[0] = fmaf(gR[k], D[k], ga[0]);
  mul33(E[0], dE[1], tmp);
  mul33(tmp, E[2], D);
  ga[1] = 0.f;
  for (int k = 0; k < 9; ++k) ga[1] = fmaf(gR[k], D[k], ga[1]);
  mul33(R01, dE[2], D);
  ga[2] = 0.f;
  for (int k = 0; k < 9; ++k) ga[2] = fmaf(gR[k], D[k], ga[2]);
  for (int k = 0; k < 3; ++k) p.grot[b * 3 + k] = ga[k];
}

// ------------------------------------------------------------------------------------- optimiser + scheduler + log
struct RegUpdateParams {
  float* rot;  // (n_rot) parameters, updated in place
  float* xyz;  // (3)
  const float* grot;
  const float* gxyz;
  int n_rot;
  float* m_rot; float* v_rot;  // Adam moments
  float* m_xyz; float* v_xyz;
  double* state;  // [step, best, num_bad, current_lr, n_plateaus, active, lr_rot, lr_xyz]
  const float* loss;
  float* rows;    // (max_rows, 1 + n_rot + 3 + 2) trajectory log
  float* count;   // rows written so far (float, as the host code keeps it)
  int max_rows;
  double beta1, beta2, adam_eps;
  double factor, patience, threshold, min_lr, sched_eps, max_n_plateaus;
};

// torch.optim.Adam(maximize=True) on one parameter vector, in the operation order of registrar.adam_maximize_
__device__ __forceinline__ void adam_vec(float* p, const float* grad, float* m, float* v, int n, double lr, double bias1,
                                         double bias2, const RegUpdateParams& q, bool on) {
  const float w = (float)(1.0 - q.beta1), b2 = (float)q.beta2, omb2 = (float)(1.0 - q.beta2);
  const float sb2 = (float)sqrt(bias2), step = (float)(lr / bias1), eps = (float)q.adam_eps;
  for (int i = 0; i < n; ++i) {
    const float g = -grad[i];
    const float mn = m[i] + w * (g - m[i]);  // torch.lerp(m, g, 1 - beta1)
    const float vn = v[i] * b2 + omb2 * g * g;
    const float denom = sqrtf(vn) / sb2 + eps;
    const float pn = p[i] - step * (mn / denom);
    if (on) { m[i] = mn; v[i] = vn; p[i] = pn; }
  }
}

__global__ void reg_update_kernel(const RegUpdateParams q) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double* s = q.state;
  const double active = s[5];
  const bool on = active > 0.0;
  // ---- Adam(maximize=True), gated by `active`
  const double t = s[0] + active;
  const double bias1 = 1.0 - pow(q.beta1, t), bias2 = 1.0 - pow(q.beta2, t);
  adam_vec(q.rot, q.grot, q.m_rot, q.v_rot, q.n_rot, s[6], bias1, bias2, q, on);
  adam_vec(q.xyz, q.gxyz, q.m_xyz, q.v_xyz, 3, s[7], bias1, bias2, q, on);
  s[0] = t;
  // ---- ReduceLROnPlateau(mode="max", threshold_mode="rel") + the reference loop's stopping rule
  const double metric = (double)q.loss[0];
  const bool better = metric > s[1] * (1.0 + q.threshold);
  const double best = better ? metric : s[1];
  double num_bad = better ? 0.0 : s[2] + 1.0;
  const bool reduce = num_bad > q.patience;
  double lr[2];
  for (int k = 0; k < 2; ++k) {
    const double cur = s[6 + k];
    const double cand = fmax(cur * q.factor, q.min_lr);
    lr[k] = (reduce && (cur - cand > q.sched_eps)) ? cand : cur;
  }
  if (reduce) num_bad = 0.0;
  const bool dropped = lr[0] < s[3];
  const double current = dropped ? lr[0] : s[3];
  const double n_plateaus = s[4] + (dropped ? 1.0 : 0.0);
  const double new_active = n_plateaus >= q.max_n_plateaus ? 0.0 : active;
  if (on) {
    s[1] = best; s[2] = num_bad; s[3] = current; s[4] = n_plateaus; s[5] = new_active; s[6] = lr[0]; s[7] = lr[1];
    // ---- trajectory row: similarity before the update, parameters after it, the next learning rates
    const int idx = (int)q.count[0];
    if (idx >= 0 && idx < q.max_rows) {
      float* row = q.rows + (int64_t)idx * (1 + q.n_rot + 3 + 2);
      row[0] = q.loss[0];
      for (int i = 0; i < q.n_rot; ++i) row[1 + i] = q.rot[i];
      for (int i = 0; i < 3; ++i) row[1 + q.n_rot + i] = q.xyz[i];
      row[1 + q.n_rot + 3] = (float)lr[0];
      row[1 + q.n_rot + 4] = (float)lr[1];
    }
    q.count[0] += 1.0f;
  }
}

}  // namespace xvr

using namespace xvr;

static int fill_euler(EulerCameraParams& p, const float* rot, const float* xyz, int B, const int* axes,
                      int rotated_frame, const float* reorient16, const float* affinv16) {
  if (!rot || !xyz || B <= 0 || !axes || !reorient16 || !affinv16) {
    set_last_error("xvr_euler_camera: invalid argument");
    return XVR_ERR_INVALID;
  }
  for (int k = 0; k < 3; ++k) {
    if (axes[k] < 0 || axes[k] > 2) {
      set_last_error("xvr_euler_camera: axes must be 0 (X), 1 (Y) or 2 (Z)");
      return XVR_ERR_INVALID;
    }
    p.axis[k] = axes[k];
  }
  p.rot = rot;
  p.xyz = xyz;
  p.B = B;
  p.rotated_frame = rotated_frame;
  for (int k = 0; k < 16; ++k) {
    p.reorient[k] = reorient16[k];
    p.affinv[k] = affinv16[k];
  }
  return XVR_OK;
}

// (rot, xyz) (B,3)+(B,3) Euler pose -> cam2world, cam2vox (B,3,4): convert(rot, xyz, "euler_angles", convention),
// reorient.compose(pose) and the affine inverse in one launch.  axes/reorient16/affinv16 are HOST arrays.
extern "C" int xvr_euler_camera_fwd(const float* rot, const float* xyz, int B, const int* axes, int rotated_frame,
                                    const float* reorient16, const float* affinv16, float* cam2world, float* cam2vox,
                                    void* stream) {
  EulerCameraParams p = {};
  int rc = fill_euler(p, rot, xyz, B, axes, rotated_frame, reorient16, affinv16);
  if (rc) return rc;
  if (!cam2world || !cam2vox) {
    set_last_error("xvr_euler_camera_fwd: null output");
    return XVR_ERR_INVALID;
  }
  p.cam2world = cam2world;
  p.cam2vox = cam2vox;
  euler_camera_kernel<false><<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(p);
  return check_launch("xvr_euler_camera_fwd");
}

// grot, gxyz (B,3) from dL/dcam2vox (B,3,4) (cam2world carries no gradient: ray lengths are rotation invariant)
extern "C" int xvr_euler_camera_bwd(const float* rot, const float* xyz, int B, const int* axes, int rotated_frame,
                                    const float* reorient16, const float* affinv16, const float* gcam2vox,
                                    float* grot, float* gxyz, void* stream) {
  EulerCameraParams p = {};
  int rc = fill_euler(p, rot, xyz, B, axes, rotated_frame, reorient16, affinv16);
  if (rc) return rc;
  if (!gcam2vox || !grot || !gxyz) {
    set_last_error("xvr_euler_camera_bwd: null buffer");
    return XVR_ERR_INVALID;
  }
  p.gcam2vox = gcam2vox;
  p.grot = grot;
  p.gxyz = gxyz;
  euler_camera_kernel<true><<<(B + 63) / 64, 64, 0, (cudaStream_t)stream>>>(p);
  return check_launch("xvr_euler_camera_bwd");
}

// One registration update: Adam(maximize) on (rot, xyz), ReduceLROnPlateau(mode="max"), the stopping rule of
// registrar/base.py:262-278 and one row of the trajectory log.  state8 (DEVICE, 8 doubles) = {Adam step, best,
// num_bad, smallest lr seen, n_plateaus, active, lr_rot, lr_xyz}; hyper9 (HOST, 9 doubles) = {beta1, beta2, adam
// eps, factor, patience, threshold, min_lr, scheduler eps, max_n_plateaus}.
extern "C" int xvr_reg_update(float* rot, float* xyz, const float* grot, const float* gxyz, int n_rot, float* m_rot,
                              float* v_rot, float* m_xyz, float* v_xyz, double* state8, const float* loss,
                              float* log_rows, float* log_count, int max_rows, const double* hyper9, void* stream) {
  if (!rot || !xyz || !grot || !gxyz || n_rot <= 0 || !m_rot || !v_rot || !m_xyz || !v_xyz || !state8 || !loss ||
      !log_rows || !log_count || max_rows <= 0 || !hyper9) {
    set_last_error("xvr_reg_update: invalid argument");
    return XVR_ERR_INVALID;
  }
  RegUpdateParams q = {};
  q.rot = rot; q.xyz = xyz; q.grot = grot; q.gxyz = gxyz; q.n_rot = n_rot;
  q.m_rot = m_rot; q.v_rot = v_rot; q.m_xyz = m_xyz; q.v_xyz = v_xyz;
  q.state = state8; q.loss = loss; q.rows = log_rows; q.count = log_count; q.max_rows = max_rows;
  q.beta1 = hyper9[0]; q.beta2 = hyper9[1]; q.adam_eps = hyper9[2];
  q.factor = hyper9[3]; q.patience = hyper9[4]; q.threshold = hyper9[5]; q.min_lr = hyper9[6];
  q.sched_eps = hyper9[7]; q.max_n_plateaus = hyper9[8];
  reg_update_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(q);
  return check_launch("xvr_reg_update");
}
