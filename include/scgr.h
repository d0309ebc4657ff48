/* scgr.h -- C ABI of the B200-native differentiable Gaussian rasterizer (libscgr.so).
 *
 * Drop-in boundary for the operator SCGaussian calls from
 *   reference gaussian_renderer/__init__.py:15      (import of the external rasterizer package)
 *   reference gaussian_renderer/__init__.py:38-51   (GaussianRasterizationSettings -> ScgrView)
 *   reference gaussian_renderer/__init__.py:100-108 (GaussianRasterizer.forward -> scgr_forward_*)
 * The reference's own native layer is an external torch extension (`_C.rasterize_gaussians`,
 * `_C.rasterize_gaussians_backward`, `_C.mark_visible`) with torch::Tensor
 * signatures; this header is the torch-free equivalent: plain device pointers, sizes and a
 * stream.  The Python host side (scgaussian_b200/rasterizer.py, re-exported as the package
 * `diff_gaussian_rasterization`) binds it with ctypes; INTEGRATION.md shows the stub.
 *
 * Ownership: the library never allocates device memory.  The caller owns inputs, outputs,
 * gradients and the three scratch buffers (geometry / binning / image -- the counterparts of the
 * reference extension's geomBuffer / binningBuffer / imgBuffer), whose sizes it queries with
 * scgr_*_bytes().  The library keeps no global mutable state besides a thread-local error string,
 * the thread-local auxiliary streams, a launch counter and the (off by default) profiling log.
 *
 * Streams: every entry point enqueues its work on `stream` and returns; only scgr_forward() waits
 * on the host, once, for the instance count (the reference extension blocks at the same point of
 * every forward on a D2H copy of num_rendered).  Stage 1 additionally uses a library-owned,
 * highest-priority auxiliary stream (one per host thread and device) for the depth sort; it is
 * forked from and joined back into `stream` with events, so callers see plain stream semantics.
 * The number of (Gaussian, tile) instances R is produced on the device; scgr_forward_geometry()
 * also copies it asynchronously to status_host (pinned host memory) so that the caller can size
 * the binning buffer.  scgr_forward_render() takes the *capacity* of the binning buffer; when R >
 * capacity it renders nothing and raises the overflow flag readable at status_host[1] after the
 * stream has been synchronised -- the caller then grows the buffer and calls it again.
 *
 * Errors: every function returns 0 on success, non-zero otherwise; scgr_last_error() gives the
 * message (thread-local).  Nothing is thrown across the boundary.
 *
 * All device pointers are fp32 / int32 unless stated, row-major, contiguous; 16-byte alignment
 * of `shs` enables the vectorised load path (torch allocations always satisfy it).
 */
#ifndef SCGR_H_
#define SCGR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCGR_VERSION 200   /* major*10000 + minor*100 + patch */
#define SCGR_TILE 16        /* BLOCK_X = BLOCK_Y of the external rasterizer's config.h */

typedef void* scgr_stream_t;   /* cudaStream_t */

/* Counterpart of GaussianRasterizationSettings (12 fields; reference
 * gaussian_renderer/__init__.py:38-51).  Matrices are the torch tensors' storage as-is:
 * viewmatrix = W2C^T, projmatrix = (P*W2C)^T row-major (reference scene/cameras.py:60-62). */
typedef struct ScgrView {
    int32_t image_height;
    int32_t image_width;
    float tanfovx;
    float tanfovy;
    const float* bg;          /* device [3] */
    float scale_modifier;
    const float* viewmatrix;  /* device [16] */
    const float* projmatrix;  /* device [16] */
    int32_t sh_degree;        /* active degree, 0..3 */
    const float* campos;      /* device [3] */
    int32_t prefiltered;
    int32_t debug;            /* non-zero: synchronise + check after every kernel */
} ScgrView;

/* The per-Gaussian inputs of GaussianRasterizer.forward (reference
 * gaussian_renderer/__init__.py:100-108).  Exactly one of shs / colors_precomp / (sh_dc, sh_rest) and exactly one
 * of (scales, rotations) / cov3D_precomp is non-NULL.  means2D is never read (it only exists on
 * the Python side as the gradient hook) and is therefore absent here. */
typedef struct ScgrGaussians {
    int32_t P;
    int32_t sh_coeffs;             /* M: coefficients stored per Gaussian in shs ((max_deg+1)^2) */
    const float* means3D;          /* [P,3] */
    const float* opacities;        /* [P] (the [P,1] tensor) */
    const float* shs;              /* [P,M,3] or NULL */
    const float* colors_precomp;   /* [P,3] or NULL */
    const float* scales;           /* [P,3] or NULL */
    const float* rotations;        /* [P,4] (r,x,y,z), used as given, or NULL */
    const float* cov3D_precomp;    /* [P,6] xx,xy,xz,yy,yz,zz or NULL */
    /* SURVEY.md section 8(f) row f2, second half -- the SH coefficients read STRAIGHT from the hybrid model's four
     * arrays (reference scene/gaussian_model.py:131-140 `get_features` cats them on every render): Gaussian i < sh_n0
     * takes row i of set 0, Gaussian i >= sh_n0 row i - sh_n0 of set 1; coefficient 0 from sh_dc, 1.. from sh_rest.
     * Selected by shs == NULL && colors_precomp == NULL with sh_dc[0] or sh_dc[1] non-NULL; needs sh_coeffs == 16.
     * The 192 bytes per Gaussian that the assembled copy costs each way are never written or re-read. */
    const float* sh_dc[2];         /* [n_k, 1, 3] or NULL for an empty set */
    const float* sh_rest[2];       /* [n_k, 15, 3] */
    int32_t sh_n0;                 /* size of set 0 */
} ScgrGaussians;

/* Gradients returned by _RasterizeGaussians.backward.  Every
 * non-NULL array is written in full (zeros for culled Gaussians): no caller-side memset needed
 * (unless `accumulate` is set, see below). */
typedef struct ScgrGrads {
    float* dL_dmeans3D;        /* [P,3] */
    float* dL_dmeans2D;        /* [P,3]  NDC-scaled screen-space gradient, z = 0 */
    float* dL_dshs;            /* [P,M,3] or NULL */
    float* dL_dcolors_precomp; /* [P,3] or NULL */
    float* dL_dopacities;      /* [P] */
    float* dL_dscales;         /* [P,3] or NULL */
    float* dL_drotations;      /* [P,4] or NULL */
    float* dL_dcov3D_precomp;  /* [P,6] or NULL */
    /* Optional, for batches of views (the reference accumulates the same quantities over
     * sequential iterations, scene/gaussian_model.py:932-934):
     * densification_stats [P,2] (8-byte aligned) receives {|dL_dmeans2D[i, 0:2]| * visible_i, visible_i} with visible_i = radii[i] > 0
     * -- the two per-view terms of add_densification_stats -- so that one SUM all-reduce over a flat buffer carries
     * them; it requires `radii` (the forward's int32 [P] output).  With accumulate != 0 every parameter gradient and
     * the statistics are ADDED to what the arrays hold (gradient accumulation over the views a rank renders before
     * the single all-reduce); dL_dmeans2D is per view and is always overwritten. */
    float* densification_stats;
    const int32_t* radii;
    int32_t accumulate;
    /* Optional [P]: 1 for every Gaussian that received any gradient in this view, 0 otherwise (added when `accumulate`
     * is set): summed over the ranks it tells which rows of the gradient arrays are non-zero anywhere -- the rest need
     * not cross the NVSwitch (scgr_nvls_allreduce_rows). */
    float* live_count;
    /* split SH layout (ScgrGaussians.sh_dc / sh_rest): the gradient rows go straight to the model's arrays, dL_dshs NULL */
    float* dL_dsh_dc[2];       /* [n_k, 1, 3] */
    float* dL_dsh_rest[2];     /* [n_k, 15, 3] */
} ScgrGrads;

int scgr_version(void);
const char* scgr_last_error(void);

/* Scratch sizes in bytes (each buffer must be 256-byte aligned). */
size_t scgr_geometry_bytes(int32_t P);
size_t scgr_binning_bytes(int32_t P, int32_t image_width, int32_t image_height, int64_t capacity);
size_t scgr_image_bytes(int32_t image_width, int32_t image_height);

/* Stage 1 of GaussianRasterizer.forward: preprocess (cull, project, EWA cov2D, SH->RGB), depth
 * ordering of the Gaussians, prefix sum of tiles touched.  Writes radii[P] (int32) and the
 * geometry scratch; enqueues an async copy of {R, 0} to status_host[0..1] (int64, pinned host
 * memory; may be NULL). */
int scgr_forward_geometry(const ScgrView* view, const ScgrGaussians* g, void* geometry_scratch,
                          int32_t* radii, int64_t* status_host, scgr_stream_t stream);

/* Stage 2: instance emission, stable tile partition, per-tile ranges, alpha compositing.
 * out_color[3,H,W], out_depth[1,H,W] (un-normalised sum d*alpha*T), out_alpha[1,H,W] are written in
 * full.  `capacity` = number of instances the binning scratch was sized for.  Enqueues an async
 * copy of {R, overflow} to status_host[0..1] (may be NULL). */
int scgr_forward_render(const ScgrView* view, const ScgrGaussians* g, void* geometry_scratch,
                        void* binning_scratch, int64_t capacity, void* image_scratch,
                        float* out_color, float* out_depth, float* out_alpha,
                        int64_t* status_host, scgr_stream_t stream);

/* GaussianRasterizer.forward in ONE call: stage 1, the forward's single host wait (for R; the
 * reference's forward blocks at the same point on a D2H copy), stage 2 -- without returning to the
 * caller in between.  `status_host` (required) is int64[2] in pinned host memory; when it is
 * device-mapped (cudaHostAlloc memory under unified addressing is) the device stores R into it
 * directly and the host spins on the word instead of synchronising the stream.  `binning_scratch`
 * must have been sized for `capacity` instances by the caller *before* the call (e.g. from the
 * previous view's R plus headroom); it may be NULL with capacity 0.
 * With a device-mapped status word stage 2 is enqueued before R is known (its kernels read R on the
 * device and refuse to run past `capacity`), so the GPU does not idle between the stages; the host
 * wait only decides the return value.
 * Returns 0 when the images were rendered, SCGR_NEED_CAPACITY when stage 1 completed but
 * R = status_host[0] exceeds `capacity`: the geometry scratch is valid, stage 2 did nothing; the
 * caller allocates >= R and finishes with scgr_forward_render().  Any other value is an error. */
#define SCGR_NEED_CAPACITY 3
int scgr_forward(const ScgrView* view, const ScgrGaussians* g, void* geometry_scratch, int32_t* radii,
                 void* binning_scratch, int64_t capacity, void* image_scratch,
                 float* out_color, float* out_depth, float* out_alpha,
                 int64_t* status_host, scgr_stream_t stream);

/* _RasterizeGaussians.backward: needs the inputs and the three scratch buffers of the matching
 * forward, untouched. */
int scgr_backward(const ScgrView* view, const ScgrGaussians* g, const void* geometry_scratch,
                  const void* binning_scratch, int64_t capacity, const void* image_scratch,
                  const float* dL_dcolor, const float* dL_ddepth, const float* dL_dalpha,
                  const ScgrGrads* grads, scgr_stream_t stream);

/* GaussianRasterizer.markVisible: present[i] = view-space z of means3D[i] > 0.2. */
int scgr_mark_visible(const float* means3D, int32_t P, const float* viewmatrix, uint8_t* present,
                      scgr_stream_t stream);

/* ---- the photometric loss of reference train.py:160-161, fused ----
 *   Ll1  = l1_loss(image, gt)                                   reference utils/loss_utils.py:40-41
 *   loss = (1 - lambda_dssim) * Ll1 + lambda_dssim * (1 - ssim(image, gt))
 *   ssim = reference utils/loss_utils.py:56-94 (window 11, sigma 1.5, zero padding, mean over all elements)
 * image / gt: [C,H,W] fp32 device (a batch folds into C: the filter is depthwise).
 * scgr_photometric_forward writes out3 = {Ll1, ssim, loss} (device) and, when want_grad != 0, keeps the
 * per-pixel SSIM partial derivatives in `scratch` (scgr_photometric_scratch_bytes, 256-byte aligned).
 * scgr_photometric_backward writes dL_dimage[C,H,W] = upstream * d loss / d image, `upstream` being a
 * device scalar (NULL = 1): the autograd scalar is consumed without a host synchronisation.
 * gt receives no gradient (the reference's gt image is data). */
size_t scgr_photometric_scratch_bytes(int32_t C, int32_t H, int32_t W);
int scgr_photometric_forward(const float* image, const float* gt, int32_t C, int32_t H, int32_t W,
                             float lambda_dssim, void* scratch, int32_t want_grad, float* out3,
                             scgr_stream_t stream);
int scgr_photometric_backward(const float* image, const float* gt, int32_t C, int32_t H, int32_t W,
                              float lambda_dssim, const void* scratch, const float* upstream,
                              float* dL_dimage, scgr_stream_t stream);

/* ---- BASELINE config 5's loss path: the match-prior loss on the rendered depth ----
 * reference scene/gaussian_model.py:241-282 `GaussianModel.get_matchloss_from_renderdepth(cam0, depth0, loss_state)`,
 * called by reference train.py:164-165 (`loss += render_match_loss * 0.3`).  For the rendered view and each other view
 * (one ScgrMatchPair): match_depth = grid_sample(depth, uv0 / (width, height) * 2 - 1)  (bilinear, zeros padding,
 * align_corners False);  world = rays_o + rays_d * (match_depth / cam_rays_d.z);  xyz = intr1 (w2c1 [world; 1]);
 * xy = xyz[:2] / (xyz[2] + 1e-8);  inside = 0 < xy < (width, height);
 * pair loss = sum(mean_xy(|xy - uv1| / (width, height)) * inside * valid) / (sum(inside * valid) + 1e-8);
 * out[0] = sum over the pairs.  All arrays device fp32; matrices by value (row-major).  `scratch`: 8 floats (device)
 * written by the forward and read by the backward.  The backward zero-fills dL_ddepth [H,W] and scatters
 * upstream * d loss / d depth into it (`upstream` a device scalar, NULL = 1). */
#define SCGR_MATCH_MAX_PAIRS 8
typedef struct ScgrMatchPair {
    int32_t n;                 /* matches of this pair */
    const float* uv0;          /* [n,2] matched pixels in the rendered view (match_data["uv"]) */
    const float* rays_o;       /* [n,3] */
    const float* rays_d;       /* [n,3] */
    const float* cam_rays_d;   /* [n,3] (z is used) */
    const float* uv1;          /* [n,2] the same matches in the other view */
    const float* valid;        /* [n] blender_mask0 * blender_mask1 (> 0 counts), or NULL: all valid */
    float w2c1[12];            /* rows 0..2 of the other view's world-to-camera matrix */
    float intr1[9];            /* the other view's 3x3 intrinsics */
} ScgrMatchPair;
int scgr_match_loss_forward(const float* depth, int32_t H, int32_t W, float width, float height, const ScgrMatchPair* pairs,
                            int32_t n_pairs, float* scratch, float* out, scgr_stream_t stream);
int scgr_match_loss_backward(const float* depth, int32_t H, int32_t W, float width, float height, const ScgrMatchPair* pairs,
                             int32_t n_pairs, const float* scratch, const float* upstream, float* dL_ddepth,
                             scgr_stream_t stream);

/* ---- SURVEY.md section 8(f) row f1, second half: the DTU background term of reference train.py:150-158, :167-168 ----
 * scgr_bg_mask: mask[y,x] = 1 where max_c gt[c,y,x] < threshold for the pixel AND the (window - 1) pixels above it in
 * its column that exist (the reference's `for i in range(1, 50): bg_mask[:, i:] *= bg_mask_clone[:, :-i]`, window 50);
 * gt [C,H,W] is zeroed IN PLACE where masked (`gt_image[bg_mask.repeat(3,1,1)] = 0.`); count[0] (device float) receives
 * the number of masked pixels.  mask: uint8 [H,W] (a torch.bool tensor's storage).
 * scgr_masked_mean_forward: out2 = {values[mask].mean(), count}  (`rendered_alpha[bg_mask].mean()`; NaN for an empty
 * mask, as torch); scratch from scgr_masked_mean_scratch_bytes(n), 256-byte aligned.
 * scgr_masked_mean_backward: dL_dvalues[i] = mask[i] ? upstream / count : 0  (upstream a device scalar, NULL = 1). */
int scgr_bg_mask(float* gt, int32_t C, int32_t H, int32_t W, float threshold, int32_t window, uint8_t* mask, float* count,
                 scgr_stream_t stream);
size_t scgr_masked_mean_scratch_bytes(int64_t n);
int scgr_masked_mean_forward(const float* values, const uint8_t* mask, int64_t n, void* scratch, float* out2,
                             scgr_stream_t stream);
int scgr_masked_mean_backward(const uint8_t* mask, int64_t n, const float* out2, const float* upstream, float* dL_dvalues,
                              scgr_stream_t stream);

/* ---- the path's single collective (sum of the per-view gradients over the ranks) as
 * a two-shot all-reduce through NVSwitch multicast: rank r reduces its 1/world of the buffer with
 * multimem.ld_reduce (in-switch fp32 sum of the replicas) and broadcasts it with multimem.st.
 * `multicast_ptr` is the multicast address of a symmetric-memory buffer of n_floats fp32 (n_floats a
 * multiple of 4 * world) replicated on every rank; the caller brackets the call with cross-rank barriers
 * on the same stream (replicas complete before, all shards broadcast after).  Without multicast
 * support the host side falls back to ncclAllReduce on the same buffer. */
int scgr_nvls_allreduce(void* multicast_ptr, size_t n_floats, int32_t rank, int32_t world,
                        scgr_stream_t stream);
/* Row-sparse companion: the same two-shot reduction for an array of n_rows rows of row_floats fp32 (a multiple of 4;
 * rows 16-byte aligned), restricted to the rows i with live_count[i] != 0.  `live_count` is this rank's LOCAL copy of
 * an array that is already identical on every rank (ScgrGrads.live_count summed by scgr_nvls_allreduce): rows nobody
 * wrote are exact zeros everywhere and need no traffic.  Rank r reduces rows [r ceil(n / world), (r + 1) ceil(n / world)).
 * Same bracketing by cross-rank barriers as above. */
int scgr_nvls_allreduce_rows(void* multicast_rows, const float* live_count, int64_t n_rows, int32_t row_floats,
                             int32_t rank, int32_t world, scgr_stream_t stream);

/* The whole collective of a step in ONE launch: cross-rank barrier -> dense shot over the first
 * dense_floats of the buffer -> barrier -> row-sparse shot over the rows block -> barrier, the barriers taken INSIDE the
 * kernel on flag words in symmetric memory (instead of three host-issued barrier kernels between two launches).
 * flags[q] is the address, in rank q's replica, of a flag array of `world` uint32 words, zero before the first call (rank
 * r release-stores into word r of every array and waits on its own); sync_local is 4 zeroed uint32 words of ordinary
 * device memory owned by the caller; epoch is 1 for the first call and grows by 3 per call (same on every rank).  The
 * grid is sized to be co-resident.  A wait that exceeds ~10 s raises sync_local[2] instead of hanging the GPU. */
#define SCGR_NVLS_MAX_WORLD 8
typedef struct ScgrNvlsFused {
    void* multicast_ptr;          /* multicast address of the flat buffer */
    size_t dense_floats;          /* multiple of 4 * world */
    void* multicast_rows;         /* multicast address of the rows block, or NULL (dense only) */
    const float* live_count;      /* this rank's own mapping of the live counts (inside the dense part) */
    int64_t n_rows;
    int32_t row_floats;           /* multiple of 4 */
    int32_t rank, world;          /* world <= SCGR_NVLS_MAX_WORLD */
    uint32_t* flags[SCGR_NVLS_MAX_WORLD];
    uint32_t* sync_local;
    uint32_t epoch;
    /* optional: peer_ptrs[q] = rank q's peer-to-peer mapping of the flat buffer (all `world` entries, world 2 / 4 / 8).  Non-NULL
     * selects plain peer loads / stores instead of the multicast instructions for both shots: faster with few ranks, where
     * the switch's reduction engine runs far below the link rate.  Same sums (one reduction per element, in rank order). */
    void* peer_ptrs[SCGR_NVLS_MAX_WORLD];
} ScgrNvlsFused;
int scgr_nvls_allreduce_fused(const ScgrNvlsFused* args, scgr_stream_t stream);

/* ---- simple_knn._C.distCUDA2 (reference scene/gaussian_model.py:20, :444) ----
 * out[i] = mean over the 3 nearest other points of |p_i - p_j|^2; points [n,3] fp32 device, out [n].
 * One-shot initialisation helper (exact tiled all-pairs scan, n <= 2^20), not on the per-step path. */
int scgr_knn3_mean_dist2(const float* points, int32_t n, float* out, scgr_stream_t stream);

/* ---- activations + hybrid assembly, fused ----
 * What GaussianModel.get_xyz / get_scaling / get_rotation / get_opacity / get_features compute on every
 * render() call (reference scene/gaussian_model.py:105-152), one launch forward and one backward:
 *   means3D = cat(rayo + rayd * zval, bg_xyz)            :124-129
 *   scales  = cat(exp(_scaling), exp(bg_scaling))        :105-113
 *   rotations = cat(normalize(_rotation), normalize(bg_rotation))   :115-122 (F.normalize, eps 1e-12)
 *   opacities = cat(sigmoid(_opacity), sigmoid(bg_opacity))         :142-150
 *   shs     = cat(cat(_features_dc, bg_features_dc), cat(_features_rest, bg_features_rest), dim=1)  :131-140
 * The model is two sets of raw (pre-activation) parameters; set[0] comes first in the assembled arrays.
 * A set with rayo != NULL is ray-based (position = rayo + rayd * zval, only zval trained: reference :493);
 * otherwise its position is `xyz`.  A set with n == 0 may leave its pointers NULL (the reference's model before
 * bg Gaussians exist / a plain 3DGS model with the ray set empty). */
typedef struct ScgrModelSet {
    int32_t n;
    const float* xyz;            /* [n,3]   free positions (reference bg_xyz), or NULL */
    const float* rayo;           /* [n,3]   ray origins (reference _rayo), or NULL */
    const float* rayd;           /* [n,3]   ray directions (reference _rayd) */
    const float* zval;           /* [n,1]   depth along the ray (reference _zval) */
    const float* scaling;        /* [n,3]   log scale */
    const float* rotation;       /* [n,4]   unnormalised quaternion (r,x,y,z) */
    const float* opacity;        /* [n,1]   logit */
    const float* features_dc;    /* [n,1,3] */
    const float* features_rest;  /* [n,sh_rest,3], NULL when sh_rest == 0 */
} ScgrModelSet;
typedef struct ScgrModel {
    int32_t sh_rest;             /* (max_sh_degree + 1)^2 - 1 rows of features_rest: 15 at degree 3 */
    ScgrModelSet set[2];
} ScgrModel;
typedef struct ScgrActivated {   /* the operator's inputs, P = set[0].n + set[1].n (16-byte aligned) */
    float* means3D;              /* [P,3] */
    float* scales;               /* [P,3] */
    float* rotations;            /* [P,4] */
    float* opacities;            /* [P,1] */
    float* shs;                  /* [P,sh_rest+1,3], or NULL: the operator takes the split SH layout (ScgrGaussians.sh_dc / sh_rest) */
} ScgrActivated;
typedef struct ScgrActivatedGrads {   /* what scgr_backward produced (ScgrGrads), all required */
    const float* dL_dmeans3D;
    const float* dL_dscales;
    const float* dL_drotations;
    const float* dL_dopacities;
    const float* dL_dshs;        /* NULL with the split SH layout: scgr_backward wrote dL/dfeatures_* itself */
} ScgrActivatedGrads;
typedef struct ScgrModelSetGrads {    /* written in full for a set with n > 0 */
    float* dL_dxyz;              /* [n,3] free set only */
    float* dL_dzval;             /* [n,1] ray-based set only */
    float* dL_dscaling;          /* [n,3] */
    float* dL_drotation;         /* [n,4] */
    float* dL_dopacity;          /* [n,1] */
    float* dL_dfeatures_dc;      /* [n,1,3] */
    float* dL_dfeatures_rest;    /* [n,sh_rest,3] */
} ScgrModelSetGrads;
typedef struct ScgrModelGrads {
    ScgrModelSetGrads set[2];
} ScgrModelGrads;
int scgr_assemble_forward(const ScgrModel* model, const ScgrActivated* out, scgr_stream_t stream);
int scgr_assemble_backward(const ScgrModel* model, const ScgrActivatedGrads* grads, const ScgrModelGrads* out,
                           scgr_stream_t stream);

/* ---- the per-iteration consumers of radii / screenspace_points.grad ----
 * reference train.py:192-193 and scene/gaussian_model.py:932-934, for every Gaussian i with update_filter[i]
 * (NULL: radii[i] > 0, which is what the reference passes as visibility_filter):
 *   max_radii2D[i] = max(max_radii2D[i], radii[i]);  xyz_gradient_accum[i] += |dL_dmeans2D[i, 0:2]|;  denom[i] += 1
 * in place, one launch, no host synchronisation (each boolean-mask statement of the reference is a nonzero + D2H).
 * dL_dmeans2D [P,3] (screenspace_points.grad), update_filter [P] bytes (torch.bool) or NULL, radii [P] int32 or
 * NULL (then max_radii2D is left alone and update_filter is required), xyz_gradient_accum / denom [P,1] or both NULL,
 * max_radii2D [P] fp32 or NULL. */
int scgr_densification_stats(const float* dL_dmeans2D, const uint8_t* update_filter, const int32_t* radii, int32_t P,
                             float* xyz_gradient_accum, float* denom, float* max_radii2D, scgr_stream_t stream);

/* ---- the optimizer step, fused ----
 * torch.optim.Adam(groups, lr=0.0, eps=1e-15).step() as the reference runs it on its two optimizers every
 * iteration (reference scene/gaussian_model.py:491-510, train.py:204-208): no weight decay, no amsgrad.  All groups
 * in ONE launch.  `step` is the group's step count AFTER this update (torch increments first); bias corrections
 * are formed in double on the host exactly as torch/optim/adam.py does.  param / exp_avg / exp_avg_sq are updated
 * in place; n < 2^31 per group; groups with n == 0 are skipped. */
#define SCGR_ADAM_MAX_GROUPS 16
typedef struct ScgrAdamGroup {
    float* param;
    const float* grad;
    float* exp_avg;
    float* exp_avg_sq;
    int64_t n;
    double lr;                   /* divided by the bias correction in double, as torch does, before rounding to fp32 */
    int32_t step;
} ScgrAdamGroup;
int scgr_adam_step(const ScgrAdamGroup* groups, int32_t n_groups, double beta1, double beta2, double eps,
                   scgr_stream_t stream);

/* ---- SURVEY.md section 8(f) row f4, second half: prune compaction ----
 * reference scene/gaussian_model.py:777-820 (`_prune_optimizer`, `prune_points`) evaluates `t[valid_points_mask]` on
 * every per-Gaussian array of the model -- 12 parameters, their exp_avg / exp_avg_sq, _rayo, _rayd,
 * xyz_gradient_accum, denom, max_radii2D -- each a nonzero + host synchronisation + gather.  Here: ONE launch for
 * all arrays from one index list:  dst[a][j, 0:row_floats] = src[a][index[j], 0:row_floats],  j < n_out.
 * `index` is a device int64 array (what torch.nonzero returns); rows are fp32 (int / bool state is compacted by the
 * host side).  src and dst must not overlap.  n_out * row_floats < 2^40 per array; arrays with row_floats == 0 are
 * skipped. */
#define SCGR_GATHER_MAX_ARRAYS 48
typedef struct ScgrRowGather {
    const float* src;
    float* dst;
    int32_t row_floats;
} ScgrRowGather;
int scgr_gather_rows(const ScgrRowGather* arrays, int32_t n_arrays, const int64_t* index, int64_t n_out,
                     scgr_stream_t stream);

/* ---- SURVEY.md section 8(f) row f4, third part: densification append ----
 * reference scene/gaussian_model.py:822-862 (`cat_tensors_to_optimizer`, `densification_postfix`) appends the new
 * Gaussians to the 6 parameter groups of the free set, extends exp_avg / exp_avg_sq with zeros and resets the
 * statistics: 18 torch.cat + 15 zero fills.  Here: ONE launch over flat segments, dst[0:n_floats] = src[0:n_floats]
 * (src == NULL: zeros).  Segments must not overlap; n_floats < 2^40; empty segments are skipped. */
#define SCGR_COPY_MAX_SEGMENTS 48
typedef struct ScgrSegmentCopy {
    float* dst;
    const float* src;
    int64_t n_floats;
} ScgrSegmentCopy;
int scgr_copy_segments(const ScgrSegmentCopy* segments, int32_t n_segments, scgr_stream_t stream);

/* Launch accounting and per-kernel timing (the reference has no tracing at all, SURVEY.md section 5;
 * bench.py uses this for the live roofline numbers).  scgr_kernel_launch_count(): kernels this
 * library has launched since load.  scgr_profile_enable(1): from now on every kernel is bracketed
 * by CUDA events on its stream; scgr_profile_fetch() synchronises them, returns up to max_entries
 * (name, milliseconds) pairs in launch order (names are static strings) and clears the log;
 * returns the number of entries, or -1 on error. */
long long scgr_kernel_launch_count(void);
int scgr_profile_enable(int on);
int scgr_profile_fetch(const char** names, float* ms, int max_entries);

/* Introspection for stage-level parity tests (device pointers into the scratch buffers).
 * record  : [P] x 12 floats {x, y, cA, cB | cC, opacity, depth, pmin2 | r, g, b, bits(u32: radius | flags << 28)}
 *           with cA = -0.5 log2(e) conicA, cB = -log2(e) conicB, cC = -0.5 log2(e) conicC
 * point_list : [R] uint32 Gaussian ids, tile-major, depth-ordered;  ranges : [tiles] x {start,end} uint32,
 *           an empty tile holds (0xffffffff, 0)
 * n_contrib : [H*W] uint32;  final_T : [H*W] float;  tiles_touched : [P] uint32 */
typedef struct ScgrDebugViews {
    const float* record;
    const uint32_t* tiles_touched;
    const uint32_t* depth_order;      /* [P] Gaussian ids in ascending (depth, id) */
    const uint32_t* point_list;
    const uint32_t* ranges;
    const uint32_t* n_contrib;
    const float* final_T;
    const int64_t* num_rendered;      /* device int64[2] = {R, overflow} */
} ScgrDebugViews;
int scgr_debug_views(int32_t P, int32_t image_width, int32_t image_height, int64_t capacity,
                     const void* geometry_scratch, const void* binning_scratch,
                     const void* image_scratch, ScgrDebugViews* out);

#ifdef __cplusplus
}
#endif
#endif /* SCGR_H_ */
