// Pose parameterisation -> SE(3) matrix / camera matrices in ONE launch each way, for every parameterisation of
// diffdrr.pose.convert that has a closed form: euler_angles (any convention), axis_angle, so3_log_map, se3_log_map,
// quaternion, rotation_6d and quaternion_adjugate (xvr's training default, /root/reference/src/xvr/config/trainer.py:17;
// the network head calls convert at /root/reference/src/xvr/model/network.py:49-54, the sampler at
// model/sampler.py:29-31, the registration module through DRR.forward at registrar/base.py:249).  rotation_10d needs a
// symmetric 4x4 eigen-decomposition and stays on torch.linalg.eigh.
//
// In PyTorch the chain convert -> make_matrix -> reorient.compose -> affine_inverse @ ... is ~25 launches of a few
// floats each, and ~50 more in its autograd mirror image; here the forward is one thread per pose and the backward one
// thread per (pose, parameter): the map is evaluated on dual numbers (value + derivative w.r.t. that one parameter),
// i.e. forward-mode differentiation of exactly the code the forward runs, contracted with the incoming gradient.
#include "common.cuh"

namespace xvr {

enum PoseKind {
  POSE_EULER = 0,
  POSE_AXIS_ANGLE = 1,
  POSE_SO3_LOG = 2,
  POSE_SE3_LOG = 3,
  POSE_QUATERNION = 4,
  POSE_ROTATION_6D = 5,
  POSE_QUATERNION_ADJUGATE = 6,
};

struct Dual {
  float v, d;
};
__device__ __forceinline__ Dual operator+(Dual a, Dual b) { return {a.v + b.v, a.d + b.d}; }
__device__ __forceinline__ Dual operator-(Dual a, Dual b) { return {a.v - b.v, a.d - b.d}; }
__device__ __forceinline__ Dual operator-(Dual a) { return {-a.v, -a.d}; }
__device__ __forceinline__ Dual operator*(Dual a, Dual b) { return {a.v * b.v, fmaf(a.d, b.v, a.v * b.d)}; }
__device__ __forceinline__ Dual operator/(Dual a, Dual b) {
  const float q = a.v / b.v;
  return {q, (a.d - q * b.d) / b.v};
}

// scalar helpers shared by the float and the dual instantiation
__device__ __forceinline__ float lift(float, float c) { return c; }
__device__ __forceinline__ Dual lift(Dual, float c) { return {c, 0.f}; }
__device__ __forceinline__ float value(float a) { return a; }
__device__ __forceinline__ float value(Dual a) { return a.v; }
__device__ __forceinline__ void xsincos(float a, float& s, float& c) { sincosf(a, &s, &c); }
__device__ __forceinline__ void xsincos(Dual a, Dual& s, Dual& c) {
  float sv, cv;
  sincosf(a.v, &sv, &cv);
  s = {sv, cv * a.d};
  c = {cv, -sv * a.d};
}
__device__ __forceinline__ float xsqrt(float a) { return sqrtf(a); }
__device__ __forceinline__ Dual xsqrt(Dual a) {
  const float r = sqrtf(a.v);
  return {r, r > 0.f ? 0.5f * a.d / r : 0.f};  // torch's norm: subgradient 0 at the origin
}
// clamp_min by a constant: the derivative vanishes where the clamp is active (torch.clamp's autograd)
__device__ __forceinline__ float xclamp_min(float a, float c) { return a < c ? c : a; }
__device__ __forceinline__ Dual xclamp_min(Dual a, float c) { return a.v < c ? Dual{c, 0.f} : a; }

template <class T>
__device__ __forceinline__ void mat33_mul(const T A[9], const T B[9], T C[9]) {
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

template <class T>
__device__ __forceinline__ void elementary_rotation(int axis, T angle, T R[9]) {
  T s, c;
  xsincos(angle, s, c);
  const T o = lift(angle, 1.f), z = lift(angle, 0.f);
  if (axis == 0) {
    R[0] = o; R[1] = z; R[2] = z; R[3] = z; R[4] = c; R[5] = -s; R[6] = z; R[7] = s; R[8] = c;
  } else if (axis == 1) {
    R[0] = c; R[1] = z; R[2] = s; R[3] = z; R[4] = o; R[5] = z; R[6] = -s; R[7] = z; R[8] = c;
  } else {
    R[0] = c; R[1] = -s; R[2] = z; R[3] = s; R[4] = c; R[5] = z; R[6] = z; R[7] = z; R[8] = o;
  }
}

// I + a K + b K^2 with K = hat(w)
template <class T>
__device__ __forceinline__ void rodrigues(const T w[3], T a, T b, T R[9]) {
  const T z = lift(a, 0.f), o = lift(a, 1.f);
  const T K[9] = {z, -w[2], w[1], w[2], z, -w[0], -w[1], w[0], z};
  T K2[9];
  mat33_mul(K, K, K2);
#pragma unroll
  for (int k = 0; k < 9; ++k) R[k] = a * K[k] + b * K2[k];
  R[0] = R[0] + o;
  R[4] = R[4] + o;
  R[8] = R[8] + o;
}

template <class T>
__device__ __forceinline__ void quaternion_matrix(const T q[4], T R[9]) {
  const T w = q[0], x = q[1], y = q[2], z = q[3];
  const T o = lift(w, 1.f);
  const T k = lift(w, 2.f) / (w * w + x * x + y * y + z * z);
  R[0] = o - k * (y * y + z * z); R[1] = k * (x * y - z * w);     R[2] = k * (x * z + y * w);
  R[3] = k * (x * y + z * w);     R[4] = o - k * (x * x + z * z); R[5] = k * (y * z - x * w);
  R[6] = k * (x * z - y * w);     R[7] = k * (y * z + x * w);     R[8] = o - k * (x * x + y * y);
}

// Top three rows P (3x4, row-major) of convert(rot, xyz, parameterization, convention) -- xvr_b200/pose.py::_rotation
// and convert, i.e. the oracle's pose_from_params -- for one pose.
template <class T>
__device__ void pose_rows(int kind, const int axis[3], int rotated_frame, const T* rot, const T xyz[3], T P[12]) {
  T R[9];
  T t[3] = {xyz[0], xyz[1], xyz[2]};
  bool rotate_t = rotated_frame != 0;
  if (kind == POSE_EULER) {
    T E0[9], E1[9], E2[9], R01[9];
    elementary_rotation(axis[0], rot[0], E0);
    elementary_rotation(axis[1], rot[1], E1);
    elementary_rotation(axis[2], rot[2], E2);
    mat33_mul(E0, E1, R01);
    mat33_mul(R01, E2, R);
  } else if (kind == POSE_AXIS_ANGLE || kind == POSE_SO3_LOG || kind == POSE_SE3_LOG) {
    const float eps = kind == POSE_AXIS_ANGLE ? 1e-12f : 1e-4f;
    const T theta = xsqrt(xclamp_min(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2], eps));
    T s, c;
    xsincos(theta, s, c);
    const T o = lift(theta, 1.f);
    const T b = (o - c) / (theta * theta);
    rodrigues(rot, s / theta, b, R);
    if (kind == POSE_SE3_LOG) {  // t = V(w) xyz, never rotated afterwards
      T V[9];
      rodrigues(rot, b, (theta - s) / (theta * theta * theta), V);
#pragma unroll
      for (int i = 0; i < 3; ++i) t[i] = V[i * 3] * xyz[0] + V[i * 3 + 1] * xyz[1] + V[i * 3 + 2] * xyz[2];
      rotate_t = false;
    }
  } else if (kind == POSE_QUATERNION) {
    quaternion_matrix(rot, R);
  } else if (kind == POSE_ROTATION_6D) {
    // rows b1 = normalize(a1), b2 = normalize(a2 - (b1 . a2) b1), b3 = b1 x b2  (F.normalize: x / max(|x|, 1e-12))
    T b1[3], b2[3];
    const T n1 = xclamp_min(xsqrt(rot[0] * rot[0] + rot[1] * rot[1] + rot[2] * rot[2]), 1e-12f);
#pragma unroll
    for (int i = 0; i < 3; ++i) b1[i] = rot[i] / n1;
    const T dot = b1[0] * rot[3] + b1[1] * rot[4] + b1[2] * rot[5];
#pragma unroll
    for (int i = 0; i < 3; ++i) b2[i] = rot[3 + i] - dot * b1[i];
    const T n2 = xclamp_min(xsqrt(b2[0] * b2[0] + b2[1] * b2[1] + b2[2] * b2[2]), 1e-12f);
#pragma unroll
    for (int i = 0; i < 3; ++i) b2[i] = b2[i] / n2;
    R[0] = b1[0]; R[1] = b1[1]; R[2] = b1[2];
    R[3] = b2[0]; R[4] = b2[1]; R[5] = b2[2];
    R[6] = b1[1] * b2[2] - b1[2] * b2[1];
    R[7] = b1[2] * b2[0] - b1[0] * b2[2];
    R[8] = b1[0] * b2[1] - b1[1] * b2[0];
  } else {  // POSE_QUATERNION_ADJUGATE: the column of the symmetric 4x4 with the largest norm, normalised
    const int sym[16] = {0, 1, 2, 3, 1, 4, 5, 6, 2, 5, 7, 8, 3, 6, 8, 9};
    int pick = 0;
    float best = -1.f;
    T nbest = lift(rot[0], 0.f);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      T n2 = lift(rot[0], 0.f);
#pragma unroll
      for (int i = 0; i < 4; ++i) n2 = n2 + rot[sym[i * 4 + j]] * rot[sym[i * 4 + j]];
      const T n = xsqrt(n2);
      if (value(n) > best) {  // first maximum wins, as torch.argmax does
        best = value(n);
        pick = j;
        nbest = n;
      }
    }
    T q[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // A[i][pick] without dynamic register indexing
      T a = rot[sym[i * 4]];
      if (pick == 1) a = rot[sym[i * 4 + 1]];
      if (pick == 2) a = rot[sym[i * 4 + 2]];
      if (pick == 3) a = rot[sym[i * 4 + 3]];
      q[i] = a / nbest;
    }
    quaternion_matrix(q, R);
  }
#pragma unroll
  for (int i = 0; i < 3; ++i) {
#pragma unroll
    for (int j = 0; j < 3; ++j) P[i * 4 + j] = R[i * 3 + j];
    P[i * 4 + 3] = rotate_t ? R[i * 3] * t[0] + R[i * 3 + 1] * t[1] + R[i * 3 + 2] * t[2] : t[i];
  }
}

struct PoseParams {
  const float* __restrict__ rot;  // (B,n_rot)
  const float* __restrict__ xyz;  // (B,3)
  int B, kind, n_rot;
  int axis[3];
  int rotated_frame;
  float angle_scale;    // pi/180 for convert(..., degrees=True) (Euler angles only), else 1
  int has_camera;       // reorient / affinv valid
  float reorient[16];   // R0: camera frame -> pose frame (row-major 4x4)
  float affinv[16];     // world mm -> voxel index
  float* __restrict__ pose;       // (B,4,4) or null
  float* __restrict__ cam2world;  // (B,3,4) = (P R0)[:3] or null
  float* __restrict__ cam2vox;    // (B,3,4) = (Ainv P R0)[:3] or null
  const float* __restrict__ gpose;     // (B,4,4) or null
  const float* __restrict__ gcam2vox;  // (B,3,4) or null
  float* __restrict__ grot;            // (B,n_rot)
  float* __restrict__ gxyz;            // (B,3)
};

__global__ void pose_fwd_kernel(const PoseParams p) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  float rot[10], xyz[3];
  for (int k = 0; k < p.n_rot; ++k) rot[k] = p.rot[b * p.n_rot + k] * (p.kind == POSE_EULER ? p.angle_scale : 1.f);
  for (int k = 0; k < 3; ++k) xyz[k] = p.xyz[b * 3 + k];
  float P[12];
  pose_rows<float>(p.kind, p.axis, p.rotated_frame, rot, xyz, P);
  if (p.pose) {
    float* o = p.pose + b * 16;
    for (int k = 0; k < 12; ++k) o[k] = P[k];
    o[12] = 0.f; o[13] = 0.f; o[14] = 0.f; o[15] = 1.f;
  }
  if (!p.has_camera) return;
  float M[12];  // P R0 (row 3 of both factors is 0 0 0 1)
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 4; ++j)
      M[i * 4 + j] = fmaf(P[i * 4], p.reorient[j], fmaf(P[i * 4 + 1], p.reorient[4 + j],
                          fmaf(P[i * 4 + 2], p.reorient[8 + j], P[i * 4 + 3] * p.reorient[12 + j])));
  if (p.cam2world)
    for (int k = 0; k < 12; ++k) p.cam2world[b * 12 + k] = M[k];
  if (p.cam2vox)
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 4; ++j)
        p.cam2vox[b * 12 + i * 4 + j] = fmaf(p.affinv[i * 4], M[j], fmaf(p.affinv[i * 4 + 1], M[4 + j],
                                             fmaf(p.affinv[i * 4 + 2], M[8 + j], j == 3 ? p.affinv[i * 4 + 3] : 0.f)));
}

// one thread per (pose, parameter): dP/dparam by forward-mode, contracted with dL/dP
__global__ void pose_bwd_kernel(const PoseParams p) {
  const int np = p.n_rot + 3;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= p.B * np) return;
  const int b = idx / np, i = idx - b * np;
  const float scale = p.kind == POSE_EULER ? p.angle_scale : 1.f;
  Dual rot[10], xyz[3];
  for (int k = 0; k < p.n_rot; ++k) rot[k] = {p.rot[b * p.n_rot + k] * scale, k == i ? scale : 0.f};
  for (int k = 0; k < 3; ++k) xyz[k] = {p.xyz[b * 3 + k], (p.n_rot + k) == i ? 1.f : 0.f};
  Dual P[12];
  pose_rows<Dual>(p.kind, p.axis, p.rotated_frame, rot, xyz, P);
  // dL/dP (3x4): straight from gpose, and through G = A3 P4 R0  =>  (A3[:, :3])^T gG R0^T
  float acc = 0.f;
  if (p.gpose) {
    const float* g = p.gpose + b * 16;
    for (int k = 0; k < 12; ++k) acc = fmaf(g[k], P[k].d, acc);
  }
  if (p.gcam2vox && p.has_camera) {
    const float* g = p.gcam2vox + b * 12;
    float T[12];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c)
        T[r * 4 + c] = fmaf(p.affinv[r], g[c], fmaf(p.affinv[4 + r], g[4 + c], p.affinv[8 + r] * g[8 + c]));
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c) {
        const float gP = fmaf(T[r * 4], p.reorient[c * 4], fmaf(T[r * 4 + 1], p.reorient[c * 4 + 1],
                              fmaf(T[r * 4 + 2], p.reorient[c * 4 + 2], T[r * 4 + 3] * p.reorient[c * 4 + 3])));
        acc = fmaf(gP, P[r * 4 + c].d, acc);
      }
  }
  if (i < p.n_rot) p.grot[b * p.n_rot + i] = acc;
  else p.gxyz[b * 3 + (i - p.n_rot)] = acc;
}

static int fill_pose(PoseParams& p, const float* rot, const float* xyz, int B, int kind, int n_rot, const int* axes,
                     int rotated_frame, float angle_scale, const float* reorient16, const float* affinv16) {
  static const int kRot[7] = {3, 3, 3, 3, 4, 6, 10};
  if (!rot || !xyz || B <= 0 || kind < 0 || kind > 6 || n_rot != kRot[kind] || (kind == POSE_EULER && !axes) ||
      ((reorient16 == nullptr) != (affinv16 == nullptr))) {
    set_last_error("xvr_pose: invalid argument (kind 0..6 = euler_angles, axis_angle, so3_log_map, se3_log_map, "
                   "quaternion, rotation_6d, quaternion_adjugate with 3,3,3,3,4,6,10 rotation components)");
    return XVR_ERR_INVALID;
  }
  p.rot = rot;
  p.xyz = xyz;
  p.B = B;
  p.kind = kind;
  p.n_rot = n_rot;
  for (int k = 0; k < 3; ++k) {
    p.axis[k] = axes ? axes[k] : 0;
    if (p.axis[k] < 0 || p.axis[k] > 2) {
      set_last_error("xvr_pose: Euler axes must be 0, 1 or 2");
      return XVR_ERR_INVALID;
    }
  }
  p.rotated_frame = rotated_frame;
  p.angle_scale = angle_scale;
  p.has_camera = reorient16 != nullptr;
  for (int k = 0; k < 16; ++k) {
    p.reorient[k] = reorient16 ? reorient16[k] : 0.f;
    p.affinv[k] = affinv16 ? affinv16[k] : 0.f;
  }
  return XVR_OK;
}

}  // namespace xvr

using namespace xvr;

extern "C" int xvr_pose_fwd(const float* rot, const float* xyz, int B, int kind, int n_rot, const int* axes,
                            int rotated_frame, float angle_scale, const float* reorient16, const float* affinv16,
                            float* pose, float* cam2world, float* cam2vox, void* stream) {
  PoseParams p = {};
  int rc = fill_pose(p, rot, xyz, B, kind, n_rot, axes, rotated_frame, angle_scale, reorient16, affinv16);
  if (rc) return rc;
  if (!pose && !cam2world && !cam2vox) {
    set_last_error("xvr_pose_fwd: no output buffer");
    return XVR_ERR_INVALID;
  }
  if ((cam2world || cam2vox) && !p.has_camera) {
    set_last_error("xvr_pose_fwd: camera matrices need reorient16 and affinv16");
    return XVR_ERR_INVALID;
  }
  p.pose = pose;
  p.cam2world = cam2world;
  p.cam2vox = cam2vox;
  pose_fwd_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p);
  return check_launch("xvr_pose_fwd");
}

extern "C" int xvr_pose_bwd(const float* rot, const float* xyz, int B, int kind, int n_rot, const int* axes,
                            int rotated_frame, float angle_scale, const float* reorient16, const float* affinv16,
                            const float* gpose, const float* gcam2vox, float* grot, float* gxyz, void* stream) {
  PoseParams p = {};
  int rc = fill_pose(p, rot, xyz, B, kind, n_rot, axes, rotated_frame, angle_scale, reorient16, affinv16);
  if (rc) return rc;
  if ((!gpose && !gcam2vox) || !grot || !gxyz || (gcam2vox && !p.has_camera)) {
    set_last_error("xvr_pose_bwd: null gradient buffer");
    return XVR_ERR_INVALID;
  }
  p.gpose = gpose;
  p.gcam2vox = gcam2vox;
  p.grot = grot;
  p.gxyz = gxyz;
  const int n = B * (n_rot + 3);
  pose_bwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p);
  return check_launch("xvr_pose_bwd");
}
