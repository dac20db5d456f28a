/*
 * splat_b200.h — C ABI of the B200-native (sm_100a) differentiable Gaussian-splat rasterizer.
 *
 * Drop-in boundary for the native module the reference imports at
 * gaussian_renderer/__init__.py:14 (`from diff_gaussian_rasterization import ...`): that package's
 * `_C.rasterize_gaussians`, `_C.rasterize_gaussians_backward` and `_C.mark_visible` (third-party
 * ingra14m/depth-diff-gaussian-rasterization @ f2d8fa9, pinned at reference README.md:28; call sites
 * gaussian_renderer/__init__.py:94-102 and :106-114; surface restated in SURVEY.md §8b).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to contiguous fp32 (unless typed otherwise) on the device that is
 *    current when the call is made; `stream` is a cudaStream_t passed as void* (NULL = legacy default);
 *  - matrices are the reference's row-vector tensors (scene/cameras.py:68-73): m[4*c + r] = element (r,c);
 *  - scratch memory is owned by the CALLER: the library asks for it through `sfb_alloc_fn` callbacks, the
 *    same contract as the external rasterizer's std::function<char*(size_t)> resize lambdas, so that the
 *    host framework (torch) owns every byte and the three buffers can live in an autograd ctx;
 *  - every entry point returns 0 on success, a negative code on error (sfb_last_error() has the text).
 *    There is no CPU fallback: a missing/failed GPU is an error.
 */
#ifndef SPLAT_B200_H
#define SPLAT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Must return a device pointer to at least `bytes` bytes, 256-byte aligned, valid until the caller
 * releases the buffer it belongs to (after backward).  Called at most once per buffer per forward. */
typedef void* (*sfb_alloc_fn)(void* user, size_t bytes);

#define SFB_OK 0
#define SFB_ERR_CUDA -1
#define SFB_ERR_ARG -2
#define SFB_ERR_ALLOC -3

/* ABI version of this header; bumped on any signature change. */
int sfb_abi_version(void);
/* Text of the last error raised on the calling thread ("" if none). */
const char* sfb_last_error(void);

/* Forward.  Replaces _C.rasterize_gaussians (reference call site
 * gaussian_renderer/__init__.py:94-102).
 *   P                 number of Gaussians;  sh_degree active degree (0..3);  M = coefficients per Gaussian
 *                     in `shs` ([P][M][3]);  exactly one of shs / colors_precomp ([P][3]) is non-NULL;
 *                     exactly one of (scales [P][3], rotations [P][4]) / cov3D_precomp ([P][6]) is non-NULL
 *   opacities [P]     bg [3]   viewmatrix, projmatrix [16]   campos [3]
 *   out_color [3][H][W]   out_depth [1][H][W]   radii [P] (int32)      — caller-allocated outputs
 *   out_alpha [1][H][W] or NULL — optional fused coverage image sum(alpha*T): what the reference computes with
 *                     a second full pass (colours = 1, bg = 0; gaussian_renderer/__init__.py:104-115)
 *   geom/binning/img  scratch allocators; the returned base pointers + *num_rendered must be handed to
 *                     sfb_rasterize_backward unchanged.
 * One blocking 4-byte device->host read (num_rendered) per call, like the reference. */
int sfb_rasterize_forward(
    int P, int sh_degree, int M, int W, int H,
    const float* bg, const float* means3D, const float* shs, const float* colors_precomp,
    const float* opacities, const float* scales, float scale_modifier, const float* rotations,
    const float* cov3D_precomp, const float* viewmatrix, const float* projmatrix, const float* campos,
    float tan_fovx, float tan_fovy, int prefiltered,
    float* out_color, float* out_depth, float* out_alpha, int* radii,
    sfb_alloc_fn geom_alloc, void* geom_user,
    sfb_alloc_fn binning_alloc, void* binning_user,
    sfb_alloc_fn img_alloc, void* img_user,
    int* num_rendered, int debug, void* stream);

/* View-parallel training: one process per GPU, splats replicated, rank r renders camera r — the serial
 * per-view loop of train.py:169 — and the per-splat gradients of all views are summed (train.py:242-252).  sfb_xchg
 * describes the symmetric (peer-mapped) buffers that sum travels through; the caller allocates one buffer of
 * sfb_xchg_bytes() bytes per rank (same size everywhere), zero-filled once, and maps every rank's buffer into every
 * process (e.g. torch.distributed._symmetric_memory: buffer_ptrs / multicast_ptr).  Pass it to sfb_rasterize_backward
 * (the geometry kernel then writes its packed gradient records into `local` and its colour gradients into EVERY
 * rank's buffer as it computes them) and then call sfb_xchg_finish once per step on every rank. */
#define SFB_XCHG_MAX_RANKS 16
typedef struct sfb_xchg {
  int rank, world;                  /* this rank, number of ranks (<= SFB_XCHG_MAX_RANKS) */
  int P, ngeo;                      /* Gaussians; floats per packed record: 12 with shs (11 used), 16 with colors_precomp (14 used) */
  void* local;                      /* this rank's buffer */
  void* peers[SFB_XCHG_MAX_RANKS];  /* rank r's buffer as mapped into this process (peers[rank] may equal local) */
  void* mc;                         /* NVSwitch multicast mapping of the buffers (multimem.* instructions), or NULL */
  int max_ctas;                     /* 0: default grid of the exchange kernels (2 CTAs per SM).  The kernel's CTAs wait for the
                                       other ranks on the device, so ALL ranks' CTAs must be able to run at the same time:
                                       a test that emulates several ranks on ONE device passes (2 * SMs) / world here */
  const void* campos_views;         /* [world][3] floats: every rank's camera centre (fused exchange with shs), or NULL */
} sfb_xchg;

/* Backward.  Replaces _C.rasterize_gaussians_backward.  dL_dout_color [3][H][W] is the
 * cotangent of out_color; dL_dout_alpha [1][H][W] (or NULL) the cotangent of the fused out_alpha; dL_dout_depth
 * [1][H][W] (or NULL) the cotangent of out_depth — the depth image is one more composited channel (value = the splat's
 * view-space z, no background), so its gradient reaches the opacities, the 2-D geometry and, through z, the means
 * (the reference's depth losses, train.py:195-229; whether the pinned extension propagates it is not checkable from the
 * reference tree, SURVEY A.9-1: pass NULL for a forward-only depth).  Outputs (all caller-allocated, fully written by the call — no pre-zeroing
 * needed): dL_dmeans2D [P][3] (xy = gradient w.r.t. the NDC-scaled screen mean, z = 0; this is what
 * lands in viewspace_points.grad, scene/gaussian_model.py:429), dL_dcolors [P][3], dL_dopacity [P][1],
 * dL_dmeans3D [P][3], dL_dcov3D [P][6], dL_dsh [P][M][3] (may be NULL when shs is NULL),
 * dL_dscales [P][3], dL_drotations [P][4] (may be NULL when cov3D_precomp is given).  dL_dcolors may be NULL when
 * shs is given and dL_dcov3D may be NULL when scales / rotations are given (the reference returns those two
 * gradients only for the corresponding precomputed inputs); a NULL output is simply not written.
 * flags: SFB_BWD_ACC_FRESH = this is the FIRST backward on these scratch buffers since the forward that filled them
 * (the forward leaves the per-splat gradient accumulators inside geom_buffer cleared, so the backward can skip its
 * 48 B/Gaussian memset); pass 0 when in doubt or when running backward again on the same buffers (retain_graph).
 * xchg (or NULL) + xchg_epoch: view-parallel exchange mode (sfb_xchg above; needs scales / rotations).
 *   - With dL_dmeans3D, dL_dopacity, dL_dscales, dL_drotations and dL_dsh (shs; needs xchg->campos_views) or dL_dcolors
 *     (colors_precomp) given, they receive the SUM OVER THE RANKS: the geometry backward and the exchange run as ONE
 *     persistent kernel (chunks of 1024 Gaussians flow through "differentiate -> push / reduce over NVLink -> rebuild SH
 *     rows -> unpack" while later chunks are still being differentiated); every rank must make this call with the same
 *     xchg_epoch (1, 2, 3, ... +1 per step).  No sfb_xchg_finish.
 *   - With those outputs NULL the parameter gradients stay in the symmetric buffers as packed records (+ pushed colour
 *     gradients) and sfb_xchg_finish sums them afterwards.
 *   dL_dmeans2D (a per-view quantity) is written either way. */
#define SFB_BWD_ACC_FRESH 1
/* SFB_BWD_SH_FACTORED (with shs): dL_dcolors (required) receives the gradient w.r.t. the SH-evaluated colour BEFORE its
 * max(0, .) clamp — i.e. the render backward's colour gradient with the clamped channels zeroed — and dL_dsh may be
 * NULL (not written).  dL_dsh[i][k][c] = basis_k(normalize(means3D[i] - campos)) * dL_dcolors[i][c]: the view-parallel
 * trainer exchanges these 3 floats per Gaussian and view instead of the 3*M-float rows (sfb_sh_grad_combine). */
#define SFB_BWD_SH_FACTORED 2
int sfb_rasterize_backward(
    int P, int sh_degree, int M, int num_rendered, int W, int H,
    const float* bg, const float* means3D, const float* shs, const float* colors_precomp,
    const float* scales, float scale_modifier, const float* rotations, const float* cov3D_precomp,
    const float* viewmatrix, const float* projmatrix, const float* campos,
    float tan_fovx, float tan_fovy, const int* radii,
    void* geom_buffer, void* binning_buffer, void* img_buffer,
    const float* dL_dout_color, const float* dL_dout_alpha, const float* dL_dout_depth,
    float* dL_dmeans2D, float* dL_dcolors, float* dL_dopacity, float* dL_dmeans3D, float* dL_dcov3D,
    float* dL_dsh, float* dL_dscales, float* dL_drotations,
    int debug, int flags,
    const sfb_xchg* xchg /* or NULL */, unsigned xchg_epoch,   /* exchange mode: parameter-gradient outputs may be NULL */
    void* stream);

size_t sfb_xchg_bytes(int P, int world, int ngeo, int with_colour_tables /* 1 with shs */);
/* Second half of the exchange, same `epoch` as the backward of this step (1, 2, 3, ... identical on every rank, +1 per
 * step): waits (on the device) until every rank's backward has finished, sums the packed records over the ranks
 * (multimem.ld_reduce / multimem.st through the switch when x->mc, peer loads / stores otherwise), rebuilds
 * dL_dsh [P][M][3] = sum_v basis(dir_v) (x) colour gradient of view v (campos_views [world][3]; with shs), and writes
 * the summed dL_dmeans3D [P][3], dL_dopacity [P], dL_dscales [P][3], dL_drotations [P][4] (+ dL_dcolors [P][3] without
 * shs).  Asynchronous on `stream`; no NCCL involved. */
int sfb_xchg_finish(const sfb_xchg* x, unsigned epoch, int sh_degree, int M, const float* means3D,
                    const float* campos_views, float* dL_dmeans3D, float* dL_dopacity, float* dL_dscales,
                    float* dL_drotations, float* dL_dcolors, float* dL_dsh, void* stream);

/* Synchronises `stream` and reads this rank's exchange error word: 0 = every sfb_xchg_finish so far completed;
 * otherwise (1: a rank never announced its backward | 2: a rank never broadcast its sums | 3: the fused kernel ran out
 * of ready work for too long) | epoch << 8 — the device-side
 * waits are bounded (about 2 s), a kernel that gives up leaves its outputs untouched and records the failure here. */
int sfb_xchg_status(const sfb_xchg* x, unsigned* status, void* stream);
/* Measurement hooks.  sfb_xchg_timeline: synchronises `stream` and returns twelve values about the last exchange kernel
 * on this rank.  [0..5]: device timestamps (globaltimer, ns).  sfb_xchg_finish: first CTA started, last CTA past the
 * first cross-rank barrier, done with the slice reduction, with the SH rows, past the second barrier, done unpacking.
 * Fused kernel: first CTA started, last geometry chunk / NVLink unit / SH chunk / unpack chunk finished, last CTA done.
 * [6..11] (fused kernel only): SM cycles, summed over the CTAs, spent choosing work (incl. waiting), in geometry chunks,
 * in their flag releases, in NVLink units, SH chunks, unpack chunks.  sfb_xchg_tune: share of the CTAs that start on the
 * slice reduction (eighths of the grid; 0, the default: 2 with two ranks, 1 with more) and reduction round trips in flight per thread (4, default, or 16);
 * process-wide. */
int sfb_xchg_timeline(const sfb_xchg* x, unsigned long long* ns12, void* stream);
void sfb_xchg_tune(int nred_eighths, int depth);

/* Multi-view sum of SH gradients from their factored form (view-parallel training; the serial loop
 * it replaces is train.py:169-242, whose loss.backward() accumulates the V per-view dL_dsh into features.grad):
 *     dL_dsh[i][k][c] = sum_{v < V} basis_k(normalize(means3D[i] - campos[v])) * dL_dcolor_views[v][i][c]
 * basis = the real SH basis of utils/sh_utils.py:57-112 up to sh_degree; coefficients k >= (sh_degree+1)^2 of the
 * [P][M][3] output are written as 0.  campos [V][3] (device), dL_dcolor_views [V][P][3] = the dL_dcolors outputs of
 * V backward calls run with SFB_BWD_SH_FACTORED.  1 <= V <= 64.  Views are summed in index order (reproducible). */
int sfb_sh_grad_combine(int P, int V, int sh_degree, int M, const float* means3D, const float* campos,
                        const float* dL_dcolor_views, float* dL_dsh, void* stream);

/* Replaces _C.mark_visible: present[i] = 1 iff the view-space z of means3D[i] is > 0.2. */
int sfb_mark_visible(int P, const float* means3D, const float* viewmatrix, const float* projmatrix,
                     uint8_t* present, void* stream);

/* Inspection of the opaque scratch buffers (parity tests compare these with the oracle bit-for-bit).
 * Each output may be NULL.  means2D [P][2], depths [P], cov3D [P][6], conic_opacity [P][4], rgb [P][3],
 * clamped [P][3] (uint8), tiles_touched [P] (uint32).  The 3D covariance is not kept in the scratch (the
 * backward recomputes it): it is re-derived here from the forward's scales / rotations (or cov3D_precomp). */
int sfb_export_geom(int P, const void* geom_buffer, const float* scales, float scale_modifier,
                    const float* rotations, const float* cov3D_precomp, float* means2D, float* depths,
                    float* cov3D, float* conic_opacity, float* rgb, uint8_t* clamped, uint32_t* tiles_touched,
                    void* stream);
/* point_list_keys [R] (uint64: tile<<32 | depth bits), point_list [R] (uint32), ranges [T][2] (uint32). */
int sfb_export_binning(int P, int num_rendered, int W, int H, const void* geom_buffer,
                       const void* binning_buffer, uint64_t* point_list_keys, uint32_t* point_list,
                       uint32_t* ranges, void* stream);
/* The staging primitive of the render kernels on its own (tests): out [n][16] = for every list entry i the 64-byte
 * shared-memory row the TMA unit delivers (cp.async.bulk.tensor ... tile::gather4 on a tensor map over the 48-byte
 * record table): table[idx[i]][0..11] followed by four zeros.  table [P][12] fp32, idx [n] (each < P). */
int sfb_debug_gather_rows(int P, const float* table, int n, const uint32_t* idx, float* out, void* stream);
/* final_T [H][W], n_contrib [H][W] (uint32; position+1 of the last contributor in the tile list). */
int sfb_export_img(int W, int H, const void* img_buffer, float* final_T, uint32_t* n_contrib, void* stream);

/* Per-kernel device timing with CUDA events recorded on the launching stream (bench.py's roofline leg).
 * While enabled, every kernel of a forward (which = 0) / backward (which = 1) call is bracketed by an event
 * pair; sfb_profile_read waits for the records of the last such call and writes their durations (ms),
 * returning the record count; sfb_profile_name(which, i) names record i ("tile_sort.scatter", ...).  The other entry
 * points (loss, densify, activate, knn) append their records to the list of the last forward / backward call;
 * sfb_profile_enable(1) starts empty lists (which = 0 current).  At most 48 records per list. */
void sfb_profile_enable(int on);
int sfb_profile_count(int which);
int sfb_profile_read(int which, float* ms, int max_records);
const char* sfb_profile_name(int which, int i);

/* Number of kernel launches issued by the last forward / backward on the calling thread. */
int sfb_last_launch_count(void);

/* ---------------------------------------------------------------------------------------------------------------
 * SURVEY.md §8f rows adjacent to the rasterizer (the steps right after it in the reference's training iteration).
 * --------------------------------------------------------------------------------------------------------------- */

/* Fused photometric loss and its gradient.  Replaces utils/loss_utils.py:18-19 (l1_loss) and :33-76 (gaussian,
 * create_window, ssim, _ssim: 11x11 Gaussian window, sigma 1.5, zero padding 5, C1 = 0.01^2, C2 = 0.03^2,
 * size_average=True) as combined at train.py:183-184, plus the optional mask term of train.py:189-193:
 *     loss = (1 - lambda_dssim) * mean|img - gt| + lambda_dssim * (1 - ssim(img, gt))
 *            [+ lambda_mask * mean|clamp(opacity, 0, 1) - gt_mask|]
 *   img, gt [C][H][W];  opacity, gt_mask [H][W] or both NULL
 *   out_scalars [4] (device): {l1, ssim, mask_l1, loss}   (ssim = 0 when lambda_dssim == 0: it is not evaluated)
 *   dL_dimg [C][H][W] or NULL (forward only), dL_dopacity [H][W] or NULL: gradients of `loss`, times grad_scale
 *                     (the upstream cotangent, e.g. 1 / #views for train.py:242's mean), fully written
 *   scratch           device buffer of sfb_loss_scratch_bytes(C, H, W) bytes, 256-byte aligned, caller-owned
 * The scalars are reduced in a fixed order (bit-reproducible run to run). */
size_t sfb_loss_scratch_bytes(int C, int H, int W);
/* host-side: the 11 fp32 weights of the 1-D window (= loss_utils.gaussian(11, 1.5), bit for bit) */
void sfb_loss_window(float* out11);
int sfb_l1_ssim_loss(int C, int H, int W, const float* img, const float* gt, float lambda_dssim,
                     const float* opacity, const float* gt_mask, float lambda_mask, float grad_scale,
                     float* out_scalars, float* dL_dimg, float* dL_dopacity, void* scratch, void* stream);

/* Densification statistics of one view.  Replaces GaussianModel.add_densification_stats
 * (scene/gaussian_model.py:427-430) and the max_radii2D update of train.py:280-282.  For every i with
 * update_filter[i] != 0 (update_filter == NULL: radii[i] > 0, the reference's visibility_filter):
 *     xyz_gradient_accum[i] += hypot(dL_dmeans2D[i][0], dL_dmeans2D[i][1]);  denom[i] += 1;
 *     max_radii2D[i] = max(max_radii2D[i], radii[i])          (when max_radii2D and radii are non-NULL)
 * dL_dmeans2D [P][3] is viewspace_points.grad (the rasterizer backward's dL_dmeans2D); radii [P] int32. */
int sfb_densify_stats(int P, const float* dL_dmeans2D, const int* radii, const uint8_t* update_filter,
                      float* xyz_gradient_accum, float* denom, float* max_radii2D, void* stream);

/* Selection predicates of GaussianModel.densify_and_prune / densify_and_clone / densify_and_split
 * (scene/gaussian_model.py:355-425) for the P existing Gaussians:
 *     g = xyz_gradient_accum / denom (NaN -> 0);  smax = max(scale)
 *     clone = |g| >= grad_threshold && smax <= dense_extent          dense_extent = percent_dense * scene_extent
 *     split =  g  >= grad_threshold && smax >  dense_extent
 *     prune = opacity < min_opacity || (max_screen_size > 0 && (max_radii2D > max_screen_size || smax > big_extent))
 *                                                                    big_extent = 0.1 * scene_extent
 * scales [P][3], opacity [P]: activated values, or the raw parameters when raw_params != 0 (exp / sigmoid applied
 * here, scene/gaussian_model.py:53-58).  Masks are uint8 [P]; counts [3] (device, may be NULL) = #clone, #split,
 * #prune.  The optimizer-state surgery that consumes the masks stays with the caller. */
int sfb_densify_masks(int P, const float* xyz_gradient_accum, const float* denom, const float* scales,
                      const float* opacity, const float* max_radii2D, int raw_params, float grad_threshold,
                      float dense_extent, float big_extent, float min_opacity, float max_screen_size,
                      uint8_t* clone_mask, uint8_t* split_mask, uint8_t* prune_mask, uint32_t* counts, void* stream);

/* Fused parameter activations — the step right before the rasterizer: the static branch of
 * get_gaussian_dict (train.py:42-50), i.e. the GaussianModel getters of scene/gaussian_model.py:64-86:
 *     scales    = exp(raw_scaling) [+ scale_offset]      raw_scaling [P][3], or [P][1] when isotropic != 0 (:64-68);
 *                                                        scale_offset [P][3] or NULL = ret['scales'] of train.py:73
 *     rotations = raw_rotation / max(||raw_rotation||_2, 1e-12)                    [P][4]   (:70-72, F.normalize)
 *     opacity   = sigmoid(raw_opacity)                                             [P]      (:83-85)
 *     features  = cat(f_dc [P][1][3], f_rest [P][M-1][3], dim=1)                   [P][M][3] (:78-82); NULL: skipped
 * One kernel.  rotations / features (in and out) must be 16-byte aligned. */
int sfb_activate_forward(int P, int M, int isotropic, const float* raw_scaling, const float* raw_rotation,
                         const float* raw_opacity, const float* f_dc, const float* f_rest, const float* scale_offset,
                         float* scales, float* rotations, float* opacity, float* features, void* stream);
/* Its backward (what autograd runs through ExpBackward / DivBackward+NormBackward / SigmoidBackward / CatBackward):
 * each dL_draw_* / dL_df_* output may be NULL (skipped); every non-NULL output is fully written.  The gradient
 * w.r.t. scale_offset is dL_dscales itself. */
int sfb_activate_backward(int P, int M, int isotropic, const float* raw_scaling, const float* raw_rotation,
                          const float* raw_opacity, const float* dL_dscales, const float* dL_drotations,
                          const float* dL_dopacity, const float* dL_dfeatures, float* dL_draw_scaling,
                          float* dL_draw_rotation, float* dL_draw_opacity, float* dL_df_dc, float* dL_df_rest,
                          void* stream);

/* Initial scales from the point cloud.  Replaces simple_knn's distCUDA2 (un-vendored dependency,
 * README.md:29; only call site scene/gaussian_model.py:105): mean_dist2[i] = mean of the squared Euclidean distances
 * from points[i] to its 3 nearest OTHER points (self excluded by index, so duplicates count with distance 0;
 * missing neighbours when P < 4 enter as FLT_MAX, as in simple_knn).  points [P][3]; scratch: device buffer of
 * sfb_knn_scratch_bytes(P) bytes, 256-byte aligned, caller-owned.  Exact (not approximate) and asynchronous: no host
 * round trip.  Distances are fp32: fma(dz, dz, fma(dy, dy, dx * dx)). */
size_t sfb_knn_scratch_bytes(int P);
int sfb_knn3_mean_dist2(int P, const float* points, float* mean_dist2, void* scratch, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SPLAT_B200_H */
