// K3 / K4: forward and backward alpha compositing of surfels, one CTA per
// 16x16 tile of one image, all images of the batch in one launch.
//
// Restates upstream forward.cu / backward.cu renderCUDA
// (github.com/hbb1/diff-surfel-rasterization; called from
// /root/reference/nsr/gs_surfel.py:100-114).  Differences in HOW, not WHAT:
//  * each warp owns an 8x4 pixel block; for every group of 32 staged surfels
//    the lanes test the surfels' conservative cull boxes (computed in K1)
//    against the warp's block and the warp only evaluates the hits.  The cull
//    box contains every pixel that can reach alpha >= 1/255, so results are
//    identical to evaluating every (pixel, surfel) pair.
//  * the backward recomputes the forward per tile, reduces each surfel's
//    gradient over the warp with shuffles, over the CTA in shared memory, and
//    issues one global atomic per (tile, surfel, component).
#include "raster_common.cuh"

#define CHUNK 256

struct PixelGeom {
    float s0, s1, p2, rho3d, rho2d, dx, dy, depth, G, alpha;
    bool use3d;
};

// Evaluates one (pixel, surfel) pair up to alpha; returns false when the pair
// is skipped (upstream's `continue` conditions).
__device__ __forceinline__ bool eval_pair(const float4 a, const float4 b, const float4 c,
                                          float pfx, float pfy, PixelGeom &o,
                                          float &k0, float &k1, float &k2,
                                          float &l0, float &l1, float &l2)
{
    // Tu = a.xyz, Tv = (a.w, b.x, b.y), Tw = (b.z, b.w, c.x), xy = (c.y, c.z), opacity = c.w
    k0 = pfx * b.z - a.x; k1 = pfx * b.w - a.y; k2 = pfx * c.x - a.z;
    l0 = pfy * b.z - a.w; l1 = pfy * b.w - b.x; l2 = pfy * c.x - b.y;
    const float p0 = k1 * l2 - k2 * l1, p1 = k2 * l0 - k0 * l2, p2 = k0 * l1 - k1 * l0;
    if (p2 == 0.0f) return false;
    const float ip = __fdividef(1.0f, p2);
    o.p2 = p2;
    o.s0 = p0 * ip; o.s1 = p1 * ip;
    o.rho3d = o.s0 * o.s0 + o.s1 * o.s1;
    o.dx = c.y - pfx; o.dy = c.z - pfy;
    o.rho2d = GA_FILTER_INV_SQUARE * (o.dx * o.dx + o.dy * o.dy);
    o.use3d = o.rho3d <= o.rho2d;
    const float rho = fminf(o.rho3d, o.rho2d);
    o.depth = o.use3d ? (o.s0 * b.z + o.s1 * b.w) + c.x : c.x;
    if (o.depth < GA_NEAR_N) return false;
    const float power = -0.5f * rho;
    if (power > 0.0f) return false;
    o.G = __expf(power);
    o.alpha = fminf(0.99f, c.w * o.G);
    return o.alpha >= 1.0f / 255.0f;
}

__global__ void __launch_bounds__(256)
render_fwd_kernel(RasterDims d, RasterWs ws, const float *__restrict__ bg,
                  float *__restrict__ out_color, float *__restrict__ out_allmap)
{
    __shared__ float4 s_rec[6][CHUNK];
    if (ws.status[1]) return;
    const int view = blockIdx.z;
    const int tile = blockIdx.y * d.gx + blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wx0 = blockIdx.x * GA_BLOCK_X + (warp & 1) * 8;
    const int wy0 = blockIdx.y * GA_BLOCK_Y + (warp >> 1) * 4;
    const int pxi = wx0 + (lane & 7), pyi = wy0 + (lane >> 3);
    const bool inside = pxi < d.W && pyi < d.H;
    const float pfx = (float)pxi, pfy = (float)pyi;
    const float bx_lo = (float)wx0, bx_hi = (float)(wx0 + 7);
    const float by_lo = (float)wy0, by_hi = (float)(wy0 + 3);

    const uint32_t start = ws.tile_start[(size_t)view * d.T + tile];
    const uint32_t end = ws.tile_start[(size_t)view * d.T + tile + 1];
    const int total = (int)(end - start);
    const float *rec_base = ws.rec + (size_t)view * d.P * GA_REC_F;

    bool done = !inside;
    float T = 1.0f, C0 = 0, C1 = 0, C2 = 0, N0 = 0, N1 = 0, N2 = 0;
    float Dacc = 0, M1 = 0, M2 = 0, dist = 0, median_depth = 0;
    int last_contributor = 0, median_contributor = -1;

    for (int c0 = 0; c0 < total; c0 += CHUNK) {
        if (__syncthreads_count(done) == 256) break;
        const int cnt = min(CHUNK, total - c0);
        if ((int)threadIdx.x < cnt) {
            const uint32_t id = ws.ids[start + c0 + threadIdx.x];
            const float4 *src = reinterpret_cast<const float4 *>(rec_base + (size_t)id * GA_REC_F);
#pragma unroll
            for (int q = 0; q < 6; q++) s_rec[q][threadIdx.x] = __ldg(src + q);
        }
        __syncthreads();
        for (int g0 = 0; g0 < cnt; g0 += 32) {
            if (__all_sync(0xffffffffu, done)) break;
            const int j = g0 + lane;
            bool hit = false;
            if (j < cnt) {
                const float4 bb = s_rec[4][j];
                hit = !(bb.y < bx_lo || bb.x > bx_hi || bb.w < by_lo || bb.z > by_hi);
            }
            unsigned mask = __ballot_sync(0xffffffffu, hit);
            while (mask) {
                const int jj = g0 + __ffs(mask) - 1;
                mask &= mask - 1;
                const float4 a = s_rec[0][jj], b = s_rec[1][jj], c = s_rec[2][jj];
                PixelGeom pg;
                float k0, k1, k2, l0, l1, l2;
                bool ok = !done && eval_pair(a, b, c, pfx, pfy, pg, k0, k1, k2, l0, l1, l2);
                float test_T = 0.f;
                if (ok) {
                    test_T = T * (1 - pg.alpha);
                    if (test_T < 0.0001f) { done = true; ok = false; }
                }
                if (__any_sync(0xffffffffu, ok)) {
                    const float4 nr = s_rec[3][jj], gb = s_rec[5][jj];
                    if (ok) {
                        const int contributor = c0 + jj + 1;
                        const float w = pg.alpha * T;
                        const float A = 1 - T;
                        const float m = GA_FAR_N / (GA_FAR_N - GA_NEAR_N) * (1 - GA_NEAR_N / pg.depth);
                        dist += (m * m * A + M2 - 2 * m * M1) * w;
                        Dacc += pg.depth * w;
                        M1 += m * w;
                        M2 += m * m * w;
                        if (T > 0.5f) { median_depth = pg.depth; median_contributor = contributor; }
                        N0 += nr.x * w; N1 += nr.y * w; N2 += nr.z * w;
                        C0 += nr.w * w; C1 += gb.x * w; C2 += gb.y * w;
                        T = test_T;
                        last_contributor = contributor;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (inside) {
        const size_t HW = (size_t)d.H * d.W;
        const size_t pix = (size_t)pyi * d.W + pxi;
        float *fT = ws.final_T + (size_t)view * 3 * HW;
        int32_t *nc = ws.n_contrib + (size_t)view * 2 * HW;
        fT[pix] = T; fT[pix + HW] = M1; fT[pix + 2 * HW] = M2;
        nc[pix] = last_contributor; nc[pix + HW] = median_contributor;
        float *oc = out_color + (size_t)view * 3 * HW;
        oc[pix] = C0 + T * bg[0]; oc[pix + HW] = C1 + T * bg[1]; oc[pix + 2 * HW] = C2 + T * bg[2];
        float *oa = out_allmap + (size_t)view * 7 * HW;
        oa[pix] = Dacc;
        oa[pix + HW] = 1 - T;
        oa[pix + 2 * HW] = N0; oa[pix + 3 * HW] = N1; oa[pix + 4 * HW] = N2;
        oa[pix + 5 * HW] = median_depth;
        oa[pix + 6 * HW] = dist;
    }
}

cudaError_t ga_launch_render_fwd(const RasterDims &d, const RasterWs &w, const float *bg,
                                 float *out_color, float *out_allmap, cudaStream_t s)
{
    dim3 grid(d.gx, d.gy, d.NV);
    render_fwd_kernel<<<grid, 256, 0, s>>>(d, w, bg, out_color, out_allmap);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// K4 backward
// ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__global__ void __launch_bounds__(256)
render_bwd_kernel(RasterDims d, RasterWs ws, const float *__restrict__ bg,
                  const float *__restrict__ dL_dcolor, const float *__restrict__ dL_dallmap,
                  float *__restrict__ grad_acc)
{
    __shared__ float4 s_rec[6][CHUNK];
    __shared__ float s_acc[CHUNK][GA_GRAD_F + 1];
    __shared__ uint32_t s_id[CHUNK];
    __shared__ int s_touched[CHUNK];
    __shared__ int s_maxc;
    if (ws.status[1]) return;
    const int view = blockIdx.z;
    const int tile = blockIdx.y * d.gx + blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int wx0 = blockIdx.x * GA_BLOCK_X + (warp & 1) * 8;
    const int wy0 = blockIdx.y * GA_BLOCK_Y + (warp >> 1) * 4;
    const int pxi = wx0 + (lane & 7), pyi = wy0 + (lane >> 3);
    const bool inside = pxi < d.W && pyi < d.H;
    const float pfx = (float)pxi, pfy = (float)pyi;
    const float bx_lo = (float)wx0, bx_hi = (float)(wx0 + 7);
    const float by_lo = (float)wy0, by_hi = (float)(wy0 + 3);

    const uint32_t start = ws.tile_start[(size_t)view * d.T + tile];
    const size_t HW = (size_t)d.H * d.W;
    const size_t pix = inside ? (size_t)pyi * d.W + pxi : 0;
    const float *fT = ws.final_T + (size_t)view * 3 * HW;
    const int32_t *nc = ws.n_contrib + (size_t)view * 2 * HW;
    const float *rec_base = ws.rec + (size_t)view * d.P * GA_REC_F;
    float *acc_base = grad_acc + (size_t)view * d.P * GA_GRAD_F;

    const float T_final = inside ? fT[pix] : 0.f;
    float T = T_final;
    const int last_contributor = inside ? nc[pix] : 0;
    const int median_contributor = inside ? nc[pix + HW] : 0;
    float dpx0 = 0, dpx1 = 0, dpx2 = 0, dL_ddepth = 0, dL_daccum = 0, dL_dreg = 0;
    float dn0 = 0, dn1 = 0, dn2 = 0, dL_dmedian = 0;
    if (inside) {
        const float *gc = dL_dcolor + (size_t)view * 3 * HW;
        const float *ga = dL_dallmap + (size_t)view * 7 * HW;
        dpx0 = gc[pix]; dpx1 = gc[pix + HW]; dpx2 = gc[pix + 2 * HW];
        dL_ddepth = ga[pix]; dL_daccum = ga[pix + HW];
        dn0 = ga[pix + 2 * HW]; dn1 = ga[pix + 3 * HW]; dn2 = ga[pix + 4 * HW];
        dL_dmedian = ga[pix + 5 * HW]; dL_dreg = ga[pix + 6 * HW];
    }
    const float final_D = inside ? fT[pix + HW] : 0.f, final_D2 = inside ? fT[pix + 2 * HW] : 0.f;
    const float final_A = 1 - T_final;
    const float bg_dot_dpixel = bg[0] * dpx0 + bg[1] * dpx1 + bg[2] * dpx2;
    float ar0 = 0, ar1 = 0, ar2 = 0, lc0 = 0, lc1 = 0, lc2 = 0;
    float last_alpha = 0, last_depth = 0, ln0 = 0, ln1 = 0, ln2 = 0;
    float accum_depth_rec = 0, accum_alpha_rec = 0, an0 = 0, an1 = 0, an2 = 0, last_dL_dT = 0;

    // nothing behind the deepest contributor of the tile can receive gradient
    if (threadIdx.x == 0) s_maxc = 0;
    __syncthreads();
    {
        int m = last_contributor;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (lane == 0) atomicMax(&s_maxc, m);
    }
    __syncthreads();
    const int total = s_maxc;          // list positions [0,total) matter

    for (int hi = total; hi > 0; hi -= CHUNK) {
        const int lo = max(0, hi - CHUNK);
        const int cnt = hi - lo;
        // stage positions lo..hi-1; slot t holds position hi-1-t (back to front)
        if ((int)threadIdx.x < cnt) {
            const uint32_t id = ws.ids[start + (hi - 1 - threadIdx.x)];
            s_id[threadIdx.x] = id;
            const float4 *src = reinterpret_cast<const float4 *>(rec_base + (size_t)id * GA_REC_F);
#pragma unroll
            for (int q = 0; q < 6; q++) s_rec[q][threadIdx.x] = __ldg(src + q);
#pragma unroll
            for (int f = 0; f < GA_GRAD_F; f++) s_acc[threadIdx.x][f] = 0.f;
            s_touched[threadIdx.x] = 0;
        }
        __syncthreads();
        for (int g0 = 0; g0 < cnt; g0 += 32) {
            const int j = g0 + lane;
            bool hit = false;
            if (j < cnt) {
                const float4 bb = s_rec[4][j];
                hit = !(bb.y < bx_lo || bb.x > bx_hi || bb.w < by_lo || bb.z > by_hi);
            }
            unsigned mask = __ballot_sync(0xffffffffu, hit);
            while (mask) {
                const int jj = g0 + __ffs(mask) - 1;
                mask &= mask - 1;
                const int contributor = hi - 1 - jj;       // 0-based list position
                const float4 a = s_rec[0][jj], b = s_rec[1][jj], c = s_rec[2][jj];
                PixelGeom pg;
                float k0, k1, k2, l0, l1, l2;
                const bool ok = inside && contributor < last_contributor &&
                                eval_pair(a, b, c, pfx, pfy, pg, k0, k1, k2, l0, l1, l2);
                if (!__any_sync(0xffffffffu, ok)) continue;
                float g[GA_GRAD_F];
#pragma unroll
                for (int f = 0; f < GA_GRAD_F; f++) g[f] = 0.f;
                if (ok) {
                    const float4 nr = s_rec[3][jj], gb = s_rec[5][jj];
                    const float alpha = pg.alpha, G = pg.G, c_d = pg.depth, opa = c.w;
                    T = T / (1.f - alpha);
                    const float w = alpha * T;
                    float dL_dalpha = 0.0f;
                    ar0 = last_alpha * lc0 + (1.f - last_alpha) * ar0; lc0 = nr.w;
                    ar1 = last_alpha * lc1 + (1.f - last_alpha) * ar1; lc1 = gb.x;
                    ar2 = last_alpha * lc2 + (1.f - last_alpha) * ar2; lc2 = gb.y;
                    dL_dalpha += (nr.w - ar0) * dpx0 + (gb.x - ar1) * dpx1 + (gb.y - ar2) * dpx2;
                    g[15] = w * dpx0; g[16] = w * dpx1; g[17] = w * dpx2;
                    float dL_dz = 0.0f;
                    const float m_d = GA_FAR_N / (GA_FAR_N - GA_NEAR_N) * (1 - GA_NEAR_N / c_d);
                    const float dmd_dd = (GA_FAR_N * GA_NEAR_N) / ((GA_FAR_N - GA_NEAR_N) * c_d * c_d);
                    if (contributor == median_contributor - 1) dL_dz += dL_dmedian;
                    const float dL_dweight = (final_D2 + m_d * m_d * final_A - 2 * m_d * final_D) * dL_dreg;
                    dL_dalpha += dL_dweight - last_dL_dT;
                    last_dL_dT = dL_dweight * alpha + (1 - alpha) * last_dL_dT;
                    const float dL_dmd = 2.0f * (T * alpha) * (m_d * final_A - final_D) * dL_dreg;
                    dL_dz += dL_dmd * dmd_dd;
                    accum_depth_rec = last_alpha * last_depth + (1.f - last_alpha) * accum_depth_rec;
                    last_depth = c_d;
                    dL_dalpha += (c_d - accum_depth_rec) * dL_ddepth;
                    accum_alpha_rec = last_alpha + (1.f - last_alpha) * accum_alpha_rec;
                    dL_dalpha += (1 - accum_alpha_rec) * dL_daccum;
                    an0 = last_alpha * ln0 + (1.f - last_alpha) * an0; ln0 = nr.x;
                    an1 = last_alpha * ln1 + (1.f - last_alpha) * an1; ln1 = nr.y;
                    an2 = last_alpha * ln2 + (1.f - last_alpha) * an2; ln2 = nr.z;
                    dL_dalpha += (nr.x - an0) * dn0 + (nr.y - an1) * dn1 + (nr.z - an2) * dn2;
                    g[11] = w * dn0; g[12] = w * dn1; g[13] = w * dn2;
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;
                    const float dL_dG = opa * dL_dalpha;       // clamp passed through (upstream)
                    dL_dz += w * dL_ddepth;
                    if (pg.use3d) {
                        const float Tw0 = b.z, Tw1 = b.w;
                        const float dL_ds0 = dL_dG * -G * pg.s0 + dL_dz * Tw0;
                        const float dL_ds1 = dL_dG * -G * pg.s1 + dL_dz * Tw1;
                        const float ip = __fdividef(1.0f, pg.p2);
                        const float q0 = dL_ds0 * ip, q1 = dL_ds1 * ip;
                        const float q2 = -(q0 * pg.s0 + q1 * pg.s1);
                        const float dk0 = l1 * q2 - l2 * q1, dk1 = l2 * q0 - l0 * q2, dk2 = l0 * q1 - l1 * q0;
                        const float dl0 = q1 * k2 - q2 * k1, dl1 = q2 * k0 - q0 * k2, dl2 = q0 * k1 - q1 * k0;
                        g[0] = -dk0; g[1] = -dk1; g[2] = -dk2;
                        g[3] = -dl0; g[4] = -dl1; g[5] = -dl2;
                        g[6] = pfx * dk0 + pfy * dl0 + dL_dz * pg.s0;
                        g[7] = pfx * dk1 + pfy * dl1 + dL_dz * pg.s1;
                        g[8] = pfx * dk2 + pfy * dl2 + dL_dz;
                    } else {
                        g[9] = dL_dG * (-G * GA_FILTER_INV_SQUARE * pg.dx);
                        g[10] = dL_dG * (-G * GA_FILTER_INV_SQUARE * pg.dy);
                        g[8] = dL_dz;
                    }
                    g[14] = G * dL_dalpha;
                }
#pragma unroll
                for (int f = 0; f < GA_GRAD_F; f++) g[f] = warp_sum(g[f]);
                if (lane < GA_GRAD_F) {
                    float v = g[0];
#pragma unroll
                    for (int f = 1; f < GA_GRAD_F; f++) if (lane == f) v = g[f];
                    atomicAdd(&s_acc[jj][lane], v);
                }
                if (lane == 0) s_touched[jj] = 1;
            }
        }
        __syncthreads();
        if ((int)threadIdx.x < cnt && s_touched[threadIdx.x]) {
            float *dst = acc_base + (size_t)s_id[threadIdx.x] * GA_GRAD_F;
#pragma unroll
            for (int f = 0; f < GA_GRAD_F; f++) {
                const float v = s_acc[threadIdx.x][f];
                if (v != 0.f) atomicAdd(dst + f, v);
            }
        }
        __syncthreads();
    }
}

cudaError_t ga_launch_render_bwd(const RasterDims &d, const RasterWs &w, const float *bg,
                                 const float *dL_dcolor, const float *dL_dallmap,
                                 float *grad_acc, cudaStream_t s)
{
    dim3 grid(d.gx, d.gy, d.NV);
    render_bwd_kernel<<<grid, 256, 0, s>>>(d, w, bg, dL_dcolor, dL_dallmap, grad_acc);
    return cudaGetLastError();
}
