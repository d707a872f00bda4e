/*
 * ga_b200.h -- C ABI of libga_b200.so: B200 (sm_100a) kernels for the two hot
 * paths of GaussianAnything.  Plain pointers and sizes only; no torch types.
 *
 * All pointers are DEVICE pointers unless stated otherwise.  No entry point
 * allocates, frees or synchronises; everything is enqueued on `stream`
 * (a cudaStream_t passed as void*).  Return value: 0 on success, a negative
 * GA_ERR_* code on a bad argument, or a positive cudaError_t from the launch.
 *
 * ---------------------------------------------------------------------------
 * Part 1: surfel (2D Gaussian) rasteriser.
 * Replaces the native module `diff_surfel_rasterization._C` that the reference
 * binds at /root/reference/nsr/gs_surfel.py:15 and calls at
 * /root/reference/nsr/gs_surfel.py:100-114 (`_C.rasterize_gaussians`,
 * `_C.rasterize_gaussians_backward`), batched over every (batch item, view)
 * of the Python loop at /root/reference/nsr/gs_surfel.py:65,74.
 * ---------------------------------------------------------------------------
 */
#ifndef GA_B200_H
#define GA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GA_ERR_BADARG   (-1)
#define GA_ERR_WORKSPACE (-2)   /* workspace too small for the stated sizes */
#define GA_ERR_SIZE     (-3)    /* image larger than 4080 px or P*views overflow */

#define GA_RASTER_REC_FLOATS 24   /* packed per-(view,surfel) record, 96 bytes */
#define GA_RASTER_GRAD_FLOATS 18  /* per-(view,surfel) gradient accumulator */

/* Byte offsets of the sections inside the forward workspace.  The forward
 * pass fills them; the backward pass reads them (the workspace is the
 * equivalent of upstream's geomBuffer/binningBuffer/imgBuffer). */
typedef struct GaRasterLayout {
    size_t total_bytes;
    size_t status;      /* int32[16]: [0]=instances D, [1]=overflow flag, [2]=tiles sorted out of smem */
    size_t rec;         /* float[NV*P][24]  Tu3 Tv3 Tw3 | xy2 opacity | normal3 | r | bbox x0 x1 y0 y1 | g b - - */
    size_t depth;       /* float[NV*P]      view-space z (0 when culled) */
    size_t rect;        /* uint32[NV*P]     x0 | y0<<8 | x1<<16 | y1<<24 (tile units) */
    size_t tile_count;  /* uint32[NV*T]     scratch: fill cursor */
    size_t tile_start;  /* uint32[NV*T+1]   exclusive scan == tile ranges [start,end) */
    size_t keys;        /* uint64[max_instances]  (depth bits<<32 | surfel), sorted per tile */
    size_t ids;         /* uint32[max_instances]  sorted surfel index per instance */
    size_t final_T;     /* float[NV][3][H*W]  T, M1, M2 */
    size_t n_contrib;   /* int32[NV][2][H*W]  last contributor, median contributor */
} GaRasterLayout;

/* Fills *layout for NV = batch*views images of H x W, P surfels per batch item
 * and room for max_instances (surfel,tile) pairs over all images.
 * Host-only, no CUDA call. */
int ga_raster_layout(int batch, int P, int views, int H, int W,
                     int64_t max_instances, GaRasterLayout *layout);

/*
 * Forward.  gauss13: [batch][P][13] = xyz3 opacity1 scale2 quat4(wxyz) rgb3, the
 * layout of /root/reference/nsr/gs_surfel.py:68-72.  viewmats / projmats:
 * [batch*views][16] exactly as the reference passes `viewmatrix` /
 * `projmatrix` (row-vector convention).  bg: [3].
 * Outputs: out_color [NV][3][H][W], out_allmap [NV][7][H][W] (channel order of
 * /root/reference/nsr/gs_surfel.py:121-142), out_radii int32 [NV][P].
 * If the instance count exceeds max_instances, status[1] is set and the
 * images are undefined (no out-of-bounds access happens); the caller re-runs
 * with a larger workspace.
 */
int ga_raster_forward(const float *gauss13, int batch, int P, int views,
                      const float *viewmats, const float *projmats, const float *bg,
                      int H, int W, float scale_modifier,
                      float *out_color, float *out_allmap, int32_t *out_radii,
                      void *workspace, size_t workspace_bytes, int64_t max_instances,
                      void *stream);

/* Bytes of scratch the backward needs (gradient accumulators). */
size_t ga_raster_backward_scratch_bytes(int batch, int P, int views);

/*
 * Backward.  dL_dcolor [NV][3][H][W], dL_dallmap [NV][7][H][W]; grad_gauss13
 * [batch][P][13] is OVERWRITTEN with the gradient summed over the views of
 * each batch item (same column order as gauss13).  workspace must be the one
 * the matching forward filled.
 */
int ga_raster_backward(const float *gauss13, int batch, int P, int views,
                       const float *viewmats, const float *projmats, const float *bg,
                       int H, int W, float scale_modifier,
                       const int32_t *radii,
                       const float *dL_dcolor, const float *dL_dallmap,
                       const void *workspace, size_t workspace_bytes, int64_t max_instances,
                       void *scratch, size_t scratch_bytes,
                       float *grad_gauss13, void *stream);

/* Measurement aid: when enabled, cudaEvents are recorded around every kernel
 * stage of the next forward/backward; ga_profile_read synchronises on them and
 * returns per-stage milliseconds: [0] preprocess, [1] binning, [2] render fwd,
 * [3] render bwd (+accumulator memset), [4] per-surfel bwd.  Returns the number
 * of stages written (0 if profiling never ran). */
int ga_profile_enable(int on);
int ga_profile_read(float *ms, int n);

/* Library self-description (host only). */
const char *ga_b200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GA_B200_H */
