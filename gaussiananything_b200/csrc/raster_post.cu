// Per-view post-processing of the rasteriser outputs, fused into one kernel each way.
// Restates /root/reference/nsr/gs_surfel.py:121-163 (alpha slice, normals camera->world by
// (n^T @ view[:3,:3].T), median depth with nan_to_num, distortion slice, image clamp) for all B x V views at once;
// the reference runs ~10 small torch kernels per view for this.
#include "../../include/ga_b200.h"
#include <cuda_runtime.h>

namespace {

// image [NV,3,HW] | alpha [NV,1,HW] | depth [NV,1,HW] | normal [NV,3,HW] | dist [NV,1,HW]
__global__ void __launch_bounds__(256)
post_fwd_kernel(const float *__restrict__ color, const float *__restrict__ allmap, const float *__restrict__ viewmats,
                float *__restrict__ image, float *__restrict__ alpha, float *__restrict__ depth,
                float *__restrict__ normal, float *__restrict__ dist, int HW)
{
    const int v = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const float *c = color + (size_t)v * 3 * HW, *a = allmap + (size_t)v * 7 * HW, *R = viewmats + v * 16;
    float *im = image + (size_t)v * 3 * HW, *nm = normal + (size_t)v * 3 * HW;
#pragma unroll
    for (int k = 0; k < 3; k++) im[(size_t)k * HW + i] = fminf(fmaxf(c[(size_t)k * HW + i], 0.f), 1.f);
    alpha[(size_t)v * HW + i] = a[(size_t)1 * HW + i];
    const float n0 = a[(size_t)2 * HW + i], n1 = a[(size_t)3 * HW + i], n2 = a[(size_t)4 * HW + i];
#pragma unroll
    for (int d = 0; d < 3; d++) nm[(size_t)d * HW + i] = n0 * R[4 * d] + n1 * R[4 * d + 1] + n2 * R[4 * d + 2];
    const float md = a[(size_t)5 * HW + i];
    depth[(size_t)v * HW + i] = (isnan(md) || isinf(md)) ? 0.f : md;       // torch.nan_to_num(x, 0, 0) also zeroes +inf
    dist[(size_t)v * HW + i] = a[(size_t)6 * HW + i];
}

__global__ void __launch_bounds__(256)
post_bwd_kernel(const float *__restrict__ color, const float *__restrict__ allmap, const float *__restrict__ viewmats,
                const float *__restrict__ g_image, const float *__restrict__ g_alpha, const float *__restrict__ g_depth,
                const float *__restrict__ g_normal, const float *__restrict__ g_dist,
                float *__restrict__ g_color, float *__restrict__ g_allmap, int HW)
{
    const int v = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= HW) return;
    const float *c = color + (size_t)v * 3 * HW, *a = allmap + (size_t)v * 7 * HW, *R = viewmats + v * 16;
    float *gc = g_color + (size_t)v * 3 * HW, *ga = g_allmap + (size_t)v * 7 * HW;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        const float x = c[(size_t)k * HW + i];
        const float g = g_image ? g_image[((size_t)v * 3 + k) * HW + i] : 0.f;
        gc[(size_t)k * HW + i] = (x >= 0.f && x <= 1.f) ? g : 0.f;          // clamp passes the gradient inside [0,1]
    }
    ga[i] = 0.f;                                                            // expected depth is not an output
    ga[(size_t)1 * HW + i] = g_alpha ? g_alpha[(size_t)v * HW + i] : 0.f;
    float gn[3] = {0.f, 0.f, 0.f};
    if (g_normal) {
#pragma unroll
        for (int d = 0; d < 3; d++) gn[d] = g_normal[((size_t)v * 3 + d) * HW + i];
    }
#pragma unroll
    for (int k = 0; k < 3; k++) ga[(size_t)(2 + k) * HW + i] = gn[0] * R[k] + gn[1] * R[4 + k] + gn[2] * R[8 + k];
    const float md = a[(size_t)5 * HW + i];
    ga[(size_t)5 * HW + i] = (g_depth && !(isnan(md) || isinf(md))) ? g_depth[(size_t)v * HW + i] : 0.f;
    ga[(size_t)6 * HW + i] = g_dist ? g_dist[(size_t)v * HW + i] : 0.f;
}

}  // namespace

extern "C" int ga_render_post_forward(const float *color, const float *allmap, const float *viewmats, int num_views,
                                      int H, int W, float *image, float *alpha, float *depth, float *normal,
                                      float *dist, void *stream)
{
    if (!color || !allmap || !viewmats || !image || !alpha || !depth || !normal || !dist || num_views <= 0 || H <= 0 || W <= 0)
        return GA_ERR_BADARG;
    const int HW = H * W;
    dim3 grid((HW + 255) / 256, num_views);
    post_fwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(color, allmap, viewmats, image, alpha, depth, normal, dist, HW);
    return (int)cudaGetLastError();
}

extern "C" int ga_render_post_backward(const float *color, const float *allmap, const float *viewmats, int num_views,
                                       int H, int W, const float *g_image, const float *g_alpha, const float *g_depth,
                                       const float *g_normal, const float *g_dist, float *g_color, float *g_allmap,
                                       void *stream)
{
    if (!color || !allmap || !viewmats || !g_color || !g_allmap || num_views <= 0 || H <= 0 || W <= 0) return GA_ERR_BADARG;
    const int HW = H * W;
    dim3 grid((HW + 255) / 256, num_views);
    post_bwd_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(color, allmap, viewmats, g_image, g_alpha, g_depth, g_normal,
                                                            g_dist, g_color, g_allmap, HW);
    return (int)cudaGetLastError();
}
