/* xvr_b200 -- C-ABI of libxvr_b200.so: hand-written sm_100a kernels for xvr's DRR hot path.
 *
 * The reference (eigenvivek/xvr) has no FFI: its hot path is the Python API of the un-vendored dependency
 * diffdrr==0.6.0 (/root/reference/pyproject.toml:14).  Each entry point below names the DiffDRR symbol it
 * replaces and the xvr call site that pins its interface.  Conventions:
 *   - every pointer is a DEVICE pointer to contiguous fp32 (uint8 for label volumes) unless marked HOST;
 *   - the caller owns all buffers; nothing is allocated except by xvr_volume_create;
 *   - `stream` is a cudaStream_t passed as void* (NULL = default stream); all work is asynchronous on it and
 *     graph-capturable;
 *   - return value 0 = OK, -1 = invalid argument, -2 = CUDA error; xvr_last_error() gives the message
 *     (thread-local).  Nothing aborts, nothing falls back to the CPU.
 *   - volume[D0][D1][D2] with D2 contiguous; ray end points are in voxel-index coordinates of that array
 *     (what drr.affine_inverse(...) returns at /root/reference/src/xvr/model/trainer.py:285).
 */
#ifndef XVR_B200_H
#define XVR_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Per-call options word: the last argument before `stream` of every entry point that has variants.  0 = the
 * library default; unknown bits are an invalid argument.  (ABI 1 had process-global setters for these.) */
#define XVR_OPT_KSPLIT(log2) ((log2) + 1) /* trilinear forward: 2^log2 (0..3) lanes share one ray; 0 in the field = automatic
                                           * (small batches, B = 1 registration, split rays so that the SMs stay full) */
#define XVR_OPT_SIDDON_WALK 0x10          /* Siddon: voxel indices from the integer walk (opt-in: same indices, measured slower) */
#define XVR_OPT_VOLGRAD_GATHER 0x20       /* dL/dvolume: voxel-centric gather (cross-check) instead of the brick-local scatter */
#define XVR_OPT_NO_TRIM 0x40              /* trilinear forward: march all n_points samples, also those outside the box of
                                           * the volume's non-zero voxels (the default skips them: exact zeros) */
#define XVR_OPT_LABEL_BRICKS 0x80         /* trilinear forward with label channels: the `labels` buffer continues, at the next
                                           * multiple of 256 bytes after its D0*D1*D2 label bytes, with a (ceil(D0/8),
                                           * ceil(D1/8),ceil(D2/8)) uint8 table: the label every voxel of that 8^3 brick
                                           * grown by one voxel carries (outside the volume = 0), or 255 if they differ.
                                           * Uniform bricks answer the nearest-label lookup from the table (L1-resident)
                                           * instead of the volume-sized array; same channels to the bit. */
#define XVR_OPT_SIDDON_TOL(code) ((code) << 8) /* test hook: tolerance of the fast voxel-index certificate; 0 production,
                                           * 1 always the reference's exact arithmetic, 2/3/4 = x 1/2, 1/4, 1/8 (margin probes) */

int xvr_abi_version(void);
const char* xvr_last_error(void);
/* kernels launched by this library since it was loaded */
long long xvr_launch_count(void);

/* ---- Volume texture: a block-linear (layered cudaArray) copy of the CT volume for the TLD4 corner gathers.
 * Replaces nothing in the reference (which samples with F.grid_sample on the linear tensor); it is the HBM layout
 * the trilinear kernels prefer.  Optional: pass NULL as `voltex` below to gather from the linear volume. */
int xvr_volume_create(int D0, int D1, int D2, void** handle_out);
/* the same handle without the texture copy (no D0 <= 2046 limit, no second copy of the volume): only the non-zero box
 * and the brick distance field that xvr_volume_upload records -- what the Siddon entries take as `occupancy` */
int xvr_occupancy_create(int D0, int D1, int D2, void** handle_out);
int xvr_volume_upload(void* handle, const float* volume, void* stream);
int xvr_volume_destroy(void* handle);
/* Empty-space trimming.  transform_hu_to_density (trainer.py:196-197) maps air to exactly 0 and CT volumes carry wide
 * margins of it.  Every upload therefore also records, on `stream` and graph-capturably, (i) the box of the volume's
 * NON-ZERO voxels and (ii) a distance field over 8^3-voxel bricks: per brick the Chebyshev distance, in bricks, to the
 * nearest brick that -- grown by two voxels -- holds a non-zero voxel.  The forward kernels that are given the handle
 * (xvr_trilinear_*_fwd via `voltex`, xvr_siddon_drr_fwd / xvr_siddon_trace via `occupancy`) sphere-trace that field in
 * from both ends of every ray and march only between the entry into the first occupied brick and the exit from the last
 * one.  What is left out are samples with 8 zero corners / segments in zero voxels: exact zeros for every running sum,
 * so images and Jacobians are bit-identical to the full march.  XVR_OPT_NO_TRIM switches it off per call.
 * xvr_volume_bbox: bbox6 = HOST int[6] = lo0 lo1 lo2 hi0 hi1 hi2 (lo = D, hi = -1 for an all-zero volume);
 * synchronises `stream`. */
int xvr_volume_bbox(void* handle, int* bbox6, void* stream);
/* samples xvr_trilinear_drr_fwd marches for a batch (counter: zeroed DEVICE unsigned long long) -- bench.py's executed
 * share under the trimming */
int xvr_trilinear_drr_count(const void* voltex, int D0, int D1, int D2, const float* cam2vox, const float* cam2world,
                            const float* det9, int B, int det_h, int det_w, int n_points, float eps,
                            unsigned long long* counter, int opts, void* stream);

/* ---- Trilinear renderer = diffdrr.renderers.Trilinear.forward
 * call site /root/reference/src/xvr/model/trainer.py:288  drr.renderer(vol, source, target, raylen, mask=seg)
 *   source (B,1,3), target (B,N,3), raylen (B,N); labels (D0,D1,D2) uint8 or NULL with C = max label + 1 (else 1)
 *   step_mode 0: span/(n-1)  1: span/n  2: 1/n ;  det_h*det_w == N selects compact detector tiles (0,0 = linear)
 *   out (B,C,N);  jac (B,7,N) or NULL: per-ray d(sum over channels of out)/d(source xyz, target xyz, raylen) --
 *   with labels this serves callers that collapse the channels (trainer.py:294), others use xvr_*_rays_bwd */
int xvr_trilinear_rays_fwd(const float* volume, const void* voltex, int D0, int D1, int D2, const uint8_t* labels,
                           int C, const float* source, const float* target, const float* raylen, int B, int N,
                           int n_points, int step_mode, float eps, int det_h, int det_w, int lane_w_log2,
                           int cta_w_log2, float* out, float* jac, int opts, void* stream);
/* autograd backward of the above (= grid_sample backward + glue): re-marches the rays.
 *   gout (B,C,N) -> gsource (B,1,3), gtarget (B,N,3), graylen (B,N); workspace (B,3,N);
 *   gvol NULL, or (D0,D1,D2) that dL/dvolume is ADDED to: the reference's own scatter (one RED.ADD per corner,
 *   grid_sampler_3d_backward's safe_add_3d) -- rays given as tensors carry no detector geometry to derive an
 *   atomics-free ownership from (the fused path has one: xvr_trilinear_drr_bwd_volume) */
int xvr_trilinear_rays_bwd(const float* volume, const void* voltex, int D0, int D1, int D2, const uint8_t* labels,
                           int C, const float* source, const float* target, const float* raylen, int B, int N,
                           int n_points, int step_mode, float eps, int det_h, int det_w, int lane_w_log2,
                           int cta_w_log2, const float* gout, float* gsource, float* gtarget, float* graylen,
                           float* workspace, float* gvol, void* stream);
/* backward through a Jacobian saved by a *_rays_fwd call: 28 bytes per ray instead of a second march */
int xvr_rays_jac_bwd(const float* jac, const float* gout, int B, int N, float* gsource, float* gtarget,
                     float* graylen, float* workspace, void* stream);

/* ---- Fused DRR = diffdrr.drr.DRR.forward (detector -> ray length -> affine_inverse -> Trilinear -> reshape),
 * the sequence xvr restates at /root/reference/src/xvr/model/trainer.py:283-289 and reaches through
 * Registration.forward at /root/reference/src/xvr/registrar/base.py:249.
 *   cam2vox, cam2world (B,3,4) row-major; det9 HOST float[9] = detector origin, row step, column step (camera mm)
 *   out (B,1,H*W); jac (B,7,H*W) or NULL */
int xvr_trilinear_drr_fwd(const float* volume, const void* voltex, int D0, int D1, int D2, const float* cam2vox,
                          const float* cam2world, const float* det9, int B, int det_h, int det_w, int n_points,
                          int step_mode, float eps, int lane_w_log2, int cta_w_log2, float* out, float* jac,
                          int opts, void* stream);
/* the same with label channels = DRR.forward(..., mask_to_channels=True) / trainer.py:283-289 with mask=seg, rays
 * generated in the kernel: out (B,C,H*W); jac (B,7,H*W) or NULL = the Jacobian of the channel SUM (what a caller that
 * collapses the channels differentiates, trainer.py:294; backward = xvr_drr_jac_bwd on the summed upstream gradient).
 * labels / C / XVR_OPT_LABEL_BRICKS as in xvr_trilinear_rays_fwd. */
int xvr_trilinear_drr_fwd_labels(const float* volume, const void* voltex, int D0, int D1, int D2, const uint8_t* labels,
                                 int C, const float* cam2vox, const float* cam2world, const float* det9, int B,
                                 int det_h, int det_w, int n_points, int step_mode, float eps, int lane_w_log2,
                                 int cta_w_log2, float* out, float* jac, int opts, void* stream);
/* gG (B,3,4) = dL/d cam2vox from the saved Jacobian and gout (B,1,H*W).  workspace: NULL, or
 * 12 * B * xvr_drr_jac_bwd_slices(B, H*W) floats so that small batches spread each pose over several CTAs */
int xvr_drr_jac_bwd_slices(int B, int N);
int xvr_drr_jac_bwd(const float* jac, const float* gout, const float* det9, int B, int det_h, int det_w, float* gG,
                    float* workspace, void* stream);

/* dL/dvolume of xvr_trilinear_drr_fwd = the `grad_input` half of grid_sampler_3d_backward (fastAtomicAdd scatter,
 * ATen/native/cuda/GridSampler.cuh:263-280), here in gather form: one owner thread per voxel, no atomics,
 * deterministic.  vox2cam (B,3,4) = inverse of cam2vox; workspace 12*B*H*W floats; gvol (D0,D1,D2) (+)= gradient. */
int xvr_trilinear_drr_bwd_volume(const float* cam2vox, const float* vox2cam, const float* cam2world,
                                 const float* det9, int B, int det_h, int det_w, int n_points, int step_mode,
                                 float eps, const float* gout, int D0, int D1, int D2, float* workspace, float* gvol,
                                 int accumulate, int opts, void* stream);
/* Variant of xvr_trilinear_drr_fwd with the volume staged brick by brick in shared memory by the TMA unit
 * (csrc/trilinear_staged.cu): a CTA owns a detector tile of one pose, a producer warp issues one
 * cp.async.bulk.tensor.3d per stage of the tile's frustum into a ring of shared-memory stages (hardware zero fill =
 * grid_sample's zero padding), consumer warps interpolate from 8 shared-memory loads per sample.  Same arithmetic and
 * summation order, so out / jac are bit-identical to xvr_trilinear_drr_fwd.  Needs a 16-byte aligned volume with
 * D2 % 4 == 0.  stats: NULL or a zeroed DEVICE unsigned long long[3] = {samples served from shared memory, from global
 * memory, barrier time-outs}. */
int xvr_trilinear_drr_fwd_staged(const float* volume, int D0, int D1, int D2, const float* cam2vox,
                                 const float* cam2world, const float* det9, int B, int det_h, int det_w, int n_points,
                                 int step_mode, float eps, float* out, float* jac, unsigned long long* stats,
                                 void* stream);
/* (two formulations, both atomics-free and deterministic, same result: brick-local scatter in shared memory by
 * default, voxel-centric gather with XVR_OPT_VOLGRAD_GATHER; csrc/volgrad.cu) */

/* ---- Siddon renderer = diffdrr.renderers.Siddon.forward (same call sites, --renderer siddon)
 * Traversed voxel indices are bit-identical to the reference's sort + grid_sample(nearest) formulation. */
int xvr_siddon_rays_fwd(const float* volume, int D0, int D1, int D2, const uint8_t* labels, int C,
                        const float* source, const float* target, const float* raylen, int B, int N,
                        float voxel_shift, float eps, int det_h, int det_w, int lane_w_log2, int cta_w_log2,
                        float* out, float* jac, int opts, void* stream);
/* fused Siddon DRR = diffdrr.drr.DRR.forward with renderer="siddon" (registrar/base.py:249, trainer.py:283-289 with
 * --renderer siddon): rays generated in-kernel as in xvr_trilinear_drr_fwd, out (B,1,H*W), jac (B,7,H*W) or NULL;
 * backward = xvr_drr_jac_bwd.
 * `occupancy` (NULL = none): a handle of xvr_occupancy_create (or xvr_volume_create) the volume was uploaded to.  With
 * it every ray's plane crossings are restricted to the stretch from its entry into the first occupied 8^3 brick to its
 * exit from the last one: the segments dropped are air (value exactly 0), i.e. exact zeros for the line integral and for
 * the Jacobian sums, so image and Jacobian are bit-identical to the full traversal (XVR_OPT_NO_TRIM switches it off). */
int xvr_siddon_drr_fwd(const float* volume, const void* occupancy, int D0, int D1, int D2, const float* cam2vox,
                       const float* cam2world, const float* det9, int B, int det_h, int det_w, float voxel_shift,
                       float eps, int lane_w_log2, int cta_w_log2, float* out, float* jac, int opts, void* stream);
int xvr_siddon_rays_bwd(const float* volume, int D0, int D1, int D2, const uint8_t* labels, int C,
                        const float* source, const float* target, const float* raylen, int B, int N,
                        float voxel_shift, float eps, int det_h, int det_w, int lane_w_log2, int cta_w_log2,
                        const float* gout, float* gsource, float* gtarget, float* graylen, float* workspace,
                        float* gvol /* NULL, or (D0,D1,D2) accumulated into: one RED.ADD per segment */, int opts,
                        void* stream);
/* dL/dvolume of xvr_siddon_drr_fwd, atomics-free and deterministic (csrc/siddon_volgrad.cu): a warp owns a 16^3 brick
 * of the gradient in shared memory, projects it onto the detector and lets the rays of that pixel window walk the
 * part of their traversal inside the brick -- same crossings and certified voxel indices as the forward; lanes work
 * on rays far enough apart never to meet in a voxel.  vox2cam (B,3,4) = inverse of cam2vox; gvol (+)= gradient. */
int xvr_siddon_drr_bwd_volume(const float* cam2vox, const float* vox2cam, const float* cam2world, const float* det9,
                              int B, int det_h, int det_w, float voxel_shift, float eps, const float* gout, int D0,
                              int D1, int D2, float* gvol, int accumulate, int opts, void* stream);
/* the traversal itself (test hook, and bench.py's segment count): idx/seg (B,N,trace_max), count (B,N);
 * trace_max = 0 with idx = seg = NULL writes the per-ray segment counts only */
int xvr_siddon_trace(const float* volume, const void* occupancy /* nullable, as in xvr_siddon_drr_fwd */, int D0, int D1,
                     int D2, const float* source, const float* target, int B, int N, float voxel_shift, float eps,
                     int trace_max, int32_t* idx, float* seg, int32_t* count, int opts, void* stream);
/* Voxel index of a segment: the certified evaluation of the midpoint (default).  With XVR_OPT_SIDDON_WALK the forward
 * (without label channels) and trace kernels take it from an integer walk (+-stride at every plane crossing) whenever
 * min_a |d_a| * segment length / 2 exceeds the rounding budget of the certified evaluation -- provably the same index
 * (scripts/siddon_cheap_certificate.py; bit-exact on the B200, tests/test_siddon_gpu.py) but measured SLOWER at config 5
 * (near-ties between axes fail the certificate in a third of the warp-wide iterations): kept as a cross-check. */

/* test hook: the hoisted-reciprocal division of the traversal vs IEEE division on random operands;
 * mismatches is a DEVICE counter the caller zeroes */
int xvr_selftest_division(int blocks, int per_thread, unsigned seed, unsigned long long* mismatches, void* stream);

/* ---- Similarity = diffdrr.metrics.{NormalizedCrossCorrelation2d, MultiscaleNormalizedCrossCorrelation2d,
 * GradientNormalizedCrossCorrelation2d, Sobel}; call sites /root/reference/src/xvr/model/loss.py:16,27 and
 * /root/reference/src/xvr/registrar/base.py:119-122.
 *   x1, x2 (B,C,H,W); patch <= 0 = whole-image NCC; score (B,) (+)= weight * NCC
 *   workspace >= B*C*max(6, ceil(H/16)*ceil(W/16)) floats
 *   coef2 / coef1: NULL or (B*C,4,H-p+1,W-p+1) [(B*C,6) when patch <= 0], consumed by xvr_ncc_bwd for d/dx2, d/dx1 */
int xvr_ncc_fwd(const float* x1, const float* x2, int B, int C, int H, int W, int patch, float eps, float weight,
                int accumulate, float* score, float* workspace, float* coef2, float* coef1, void* stream);
int xvr_ncc_bwd(const float* x1, const float* x2, const float* coef, int which, const float* gscore, int B, int C,
                int H, int W, int patch, float weight, int accumulate, float* grad, void* stream);
int xvr_sobel_fwd(const float* x, int B, int H, int W, float* out /* (B,2,H,W) */, void* stream);
int xvr_sobel_bwd(const float* gout, int B, int H, int W, float* gx /* (B,1,H,W) */, void* stream);

/* ---- Value and gradient of the registration similarity in nine launches (parity-green on the B200): the chain
 *   XrayTransforms(moving) -> beta * mNCC([None, p]) + (1 - beta) * GradNCC(q, sigma = 0) -> .sum() -> backward
 * of /root/reference/src/xvr/registrar/base.py:245-252 (transform: utils/preprocess.py:5-29, no Equalize, no resize).
 *   fixed (B,1,H,W) the transformed target, fixed_sobel (B,2,H,W) = xvr_sobel_fwd(fixed), moving (B,1,H,W) raw DRRs;
 *   transform = (x - min x) / (max x - min x + std_eps) over the whole batch, then (. - mean) * inv_std;
 *   score[0] = sum_b w_global NCC + w_patch NCC_p + w_grad NCC_q(Sobel);  grad (B,1,H,W) = d score / d moving;
 *   workspace: xvr_regsim_workspace_floats(B, H, W, p, q) floats (-1 for invalid sizes). */
long long xvr_regsim_workspace_floats(int B, int H, int W, int patch_mncc, int patch_gncc);
int xvr_regsim(const float* fixed, const float* fixed_sobel, const float* moving, int B, int H, int W, float std_eps,
               float mean, float inv_std, int patch_mncc, int patch_gncc, float w_global, float w_patch,
               float w_grad, float ncc_eps, float* workspace, long long workspace_floats, float* score, float* grad,
               void* stream);

/* ---- The scalar ends of one registration iteration, one launch each (/root/reference/src/xvr/registrar/base.py:245-278).
 * xvr_euler_camera_*: convert(rot, xyz, "euler_angles", convention) -> reorient.compose(pose) -> affine inverse,
 *   i.e. the camera matrices the fused renderer consumes, and the matching backward.  rot, xyz (B,3) DEVICE;
 *   axes HOST int[3] (0/1/2 = X/Y/Z of the convention string); reorient16, affinv16 HOST row-major 4x4;
 *   cam2world, cam2vox, gcam2vox (B,3,4); grot, gxyz (B,3).
 * xvr_reg_update: torch.optim.Adam(maximize=True) on (rot, xyz) + ReduceLROnPlateau(mode="max") + the stopping rule
 *   and trajectory row of run_test_time_optimization.  state8 DEVICE double[8] = {Adam step, best, num_bad, smallest
 *   lr seen, n_plateaus, active, lr_rot, lr_xyz}; hyper9 HOST double[9] = {beta1, beta2, Adam eps, factor, patience,
 *   threshold, min_lr, scheduler eps, max_n_plateaus}; log_rows (max_rows, 1 + n_rot + 3 + 2), log_count DEVICE float. */
int xvr_euler_camera_fwd(const float* rot, const float* xyz, int B, const int* axes, int rotated_frame,
                         const float* reorient16, const float* affinv16, float* cam2world, float* cam2vox,
                         void* stream);
int xvr_euler_camera_bwd(const float* rot, const float* xyz, int B, const int* axes, int rotated_frame,
                         const float* reorient16, const float* affinv16, const float* gcam2vox, float* grot,
                         float* gxyz, void* stream);
/* xvr_pose_*: the same for every closed-form parameterisation of diffdrr.pose.convert (call sites
 *   /root/reference/src/xvr/model/network.py:49-54 -- the regressor head, quaternion_adjugate by default,
 *   config/trainer.py:17 --, model/sampler.py:29-31, registrar/base.py:168-170), one launch each way.
 *   kind 0..6 = euler_angles, axis_angle, so3_log_map, se3_log_map, quaternion, rotation_6d, quaternion_adjugate with
 *   n_rot = 3,3,3,3,4,6,10; rot (B,n_rot), xyz (B,3); axes HOST int[3] (Euler only, else NULL); angle_scale = pi/180 for
 *   degrees=True (Euler only); reorient16 / affinv16 HOST 4x4 or both NULL.  Outputs (any may be NULL): pose (B,4,4),
 *   cam2world, cam2vox (B,3,4; need the camera constants).  Backward: gpose (B,4,4) and/or gcam2vox (B,3,4) -> grot, gxyz
 *   (one thread per pose and parameter, forward-mode differentiation of the forward's own code). */
int xvr_pose_fwd(const float* rot, const float* xyz, int B, int kind, int n_rot, const int* axes, int rotated_frame,
                 float angle_scale, const float* reorient16, const float* affinv16, float* pose, float* cam2world,
                 float* cam2vox, void* stream);
int xvr_pose_bwd(const float* rot, const float* xyz, int B, int kind, int n_rot, const int* axes, int rotated_frame,
                 float angle_scale, const float* reorient16, const float* affinv16, const float* gpose,
                 const float* gcam2vox, float* grot, float* gxyz, void* stream);
int xvr_reg_update(float* rot, float* xyz, const float* grot, const float* gxyz, int n_rot, float* m_rot, float* v_rot,
                   float* m_xyz, float* v_xyz, double* state8, const float* loss, float* log_rows, float* log_count,
                   int max_rows, const double* hyper9, void* stream);

/* ---- HU -> density = diffdrr.data.transform_hu_to_density, called every step at
 * /root/reference/src/xvr/model/trainer.py:196-197.  xvr_hu_stats reduces {min soft, max soft, min bone, max bone}
 * once per HU volume (independent of the multiplier; workspace = 4 ints, stats = 4 floats, both DEVICE);
 * xvr_hu_to_density is the one-pass piecewise map, shifted and scaled to [0,1]. */
int xvr_hu_stats(const float* hu, long long n, float air, float bone, int* workspace, float* stats, void* stream);
int xvr_hu_to_density(const float* hu, long long n, float air, float bone, float multiplier,
                      const float* multiplier_dev /* NULL, or DEVICE float overriding `multiplier` */,
                      const float* stats, float* out, void* stream);

/* ---- The post-render epilogue of the training iteration (/root/reference/src/xvr/model/trainer.py:292-302 +
 * utils/preprocess.py:28-29) in one launch: img (B,C,N) -> sum_img (B,N) = img.sum(dim=1) (NULL allowed when C == 1),
 * stats (B,4) = {foreground fraction, keep (0/1: fraction > img_threshold without label channels, fraction of pixels
 * with any foreground channel > mask_threshold with them), min, max of the channel sum}; deterministic. */
int xvr_render_epilogue(const float* img, int B, int C, int N, float img_threshold, float mask_threshold,
                        float* sum_img, float* stats, void* stream);

/* out[r] = sum_n in[r,n], fixed summation tree */
int xvr_reduce_rows(const float* in, int rows, int N, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
