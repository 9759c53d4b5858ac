/* bts_b200 -- C ABI of the B200-native BTS hot path (libbts_b200.so).
 *
 * This is the native boundary underneath the Python module surface `bts` (BtsModel / encoder / bts /
 * local_planar_guidance / reduction_1x1 / silog_loss, reference pytorch/bts.py).  The reference's only
 * native interface is the TensorFlow custom op (tensorflow/custom_layer/local_planar_guidance.h:22-49:
 * LocalPlanarGuidanceKernel / LocalPlanarGuidanceGradKernel functors taking raw float pointers + sizes);
 * bts_lpg_fwd / bts_lpg_bwd replace exactly those two functors (layout=BTS_LAYOUT_NHWC, tf_compat=1
 * reproduces the op bit-for-formula); every other entry point replaces an ATen/cuDNN call sequence of
 * pytorch/bts.py cited per function.
 *
 * Conventions (all entry points):
 *   - plain pointers are DEVICE pointers (fp32 unless noted); no torch / TF types.
 *   - asynchronous on the caller-supplied CUDA stream (void* == cudaStream_t); never allocates,
 *     never synchronises (the reference op calls d.synchronize() after each launch, .cu:91,170 -- we do not).
 *   - returns 0 on success, a negative BTS_E* on bad arguments, a positive cudaError_t on launch failure.
 *   - `*_h` variants take HOST pointers, do the H2D/D2H copies themselves and synchronise; they are the
 *     reference-facing plugin form used for end-to-end (e2e) measurements and by non-torch callers.
 */
#ifndef BTS_B200_H_
#define BTS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BTS_LAYOUT_NCHW 0 /* plane (B,4,h,w)  -- pytorch/bts.py:135-138 */
#define BTS_LAYOUT_NHWC 1 /* plane (B,h,w,4)  -- custom_layer/local_planar_guidance.cc:102-107 */

#define BTS_EINVAL (-1)   /* bad argument (null pointer, non-positive size, unsupported upratio) */
#define BTS_EALIGN (-2)   /* pointer not 16-byte aligned where the kernel needs it */

/* library / build identification: returns e.g. 100 for sm_100a; writes a short version string */
int bts_version(char *buf, int buflen);

/* ---- Local planar guidance --------------------------------------------------------------------
 * Forward  (pytorch/bts.py:132-146; custom_layer/local_planar_guidance.cu:33-72, .cc:74-115):
 *     depth[b,y,x] = n4 / ((n1*u(x) + n2*v(y)) + n3),  plane of patch (y/r, x/r),
 *     u(x) = ((x mod r) - (r-1)/2)/r  (bit-exact dyadic grid), v(y) likewise.
 * `focal` is unused by the reference in both implementations and is not a parameter.
 * Optional fused by-products of bts.forward (pytorch/bts.py:228-229,242-243,256); pass NULL to skip:
 *     scaled = depth / max_depth            (B,H,W)
 *     ds     = scaled[:, ::ds_stride, ::ds_stride]   (B,H/ds_stride,W/ds_stride)  nearest down-sample
 * `depth` itself may be NULL when only the by-products are wanted.  upratio r must be 1 or even. */
int bts_lpg_fwd(const float *plane, float *depth, int B, int h, int w, int r, int layout, void *stream);
int bts_lpg_fwd_fused(const float *plane, float *depth, float *scaled, float *ds, float max_depth,
                      int ds_stride, int B, int h, int w, int r, int layout, void *stream);

/* Backward (autograd of pytorch/bts.py:146; custom_layer/local_planar_guidance.cu:95-150, .cc:241-298):
 *     dplane[b,:,i,j] = sum over the r x r tile of (-dY*n4*u/den^2, -dY*n4*v/den^2, -dY*n4/den^2, dY/den)
 * tf_compat=1 drops n4 from the first three (the reference TF kernel's formula).
 * Fused form: dY_total = d_depth + (d_scaled + scatter(d_ds)) / max_depth, any of the three may be NULL. */
int bts_lpg_bwd(const float *dy, const float *plane, float *dplane, int B, int h, int w, int r,
                int layout, int tf_compat, void *stream);
int bts_lpg_bwd_fused(const float *d_depth, const float *d_scaled, const float *d_ds, float max_depth,
                      int ds_stride, const float *plane, float *dplane, int B, int h, int w, int r,
                      int layout, int tf_compat, void *stream);

/* Host-pointer plugin form of the TF op pair (input (B,h,w,4) NHWC, output (B,h*r,w*r)); copies in/out
 * and synchronises.  Mirrors LocalPlanarGuidanceOp::Compute / LocalPlanarGuidanceGradOp::Compute
 * (custom_layer/local_planar_guidance.cc:190-229, 364-416). */
int bts_lpg_fwd_h(const float *plane_host, float *depth_host, int B, int h, int w, int r, int layout);
int bts_lpg_bwd_h(const float *dy_host, const float *plane_host, float *dplane_host, int B, int h, int w,
                  int r, int layout, int tf_compat);

/* ---- silog loss (pytorch/bts.py:41-48) ---------------------------------------------------------
 * mask: uint8/bool, n elements.  ws: >= 4 doubles of device workspace, zeroed by the call.
 * fwd:  d_i = ln est_i - ln gt_i over mask; loss = 10*sqrt(mean(d^2) - lambda*mean(d)^2)  -> loss[0]
 *       ws keeps (sum d, sum d^2, N) for the backward.
 * bwd:  dest_i = gout[0] * mask_i * (10/sqrt(S)) * (d_i - lambda*m1) / (N*est_i)    */
int bts_silog_fwd(const float *est, const float *gt, const uint8_t *mask, long long n, float lambda,
                  double *ws, float *loss, void *stream);
int bts_silog_bwd(const float *est, const float *gt, const uint8_t *mask, long long n, float lambda,
                  const double *ws, const float *gout, float *dest, void *stream);

/* ---- plane-coefficient head tail (pytorch/bts.py:112-120 + 223-229 per scale) -------------------
 * c3: (B,3,h,w) output of reduc.plane_params.  Computes theta=sig(c0)*pi/3, phi=sig(c1)*2pi,
 * dist=sig(c2)*max_depth, n=(sin th cos ph, sin th sin ph, cos th), n^=n/max(|n|,1e-12), plane=(n^,dist),
 * then LPG(r) -> scaled=(depth/max_depth) (B,H,W) and ds (optional, nearest down-sample by ds_stride).
 * plane_out (B,4,h,w) optional (saved for backward / inspection). */
int bts_plane_head_fwd(const float *c3, float *plane_out, float *scaled, float *ds, float max_depth,
                       int ds_stride, int B, int h, int w, int r, void *stream);
/* backward of the above: inputs d_scaled (B,H,W), d_ds (optional), c3; output dc3 (B,3,h,w) */
int bts_plane_head_bwd(const float *d_scaled, const float *d_ds, const float *c3, float *dc3,
                       float max_depth, int ds_stride, int B, int h, int w, int r, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* BTS_B200_H_ */
