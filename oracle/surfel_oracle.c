/*
 * surfel_oracle.c -- CPU restatement of the 2D-Gaussian (surfel) rasteriser.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under gaussiananything_b200/ may import,
 * link or execute this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs do, and only as the checker
 * (or as the timed CPU baseline), never as the product path.
 *
 * PARITY UNPINNED.  The arithmetic of this path lives in the third-party
 * package `diff_surfel_rasterization` (github.com/hbb1/diff-surfel-rasterization,
 * installed from git HEAD by the reference, /root/reference/README.md:156-159;
 * NOT vendored: /root/reference/.gitmodules:1-6 is commented out).  The
 * reference ships no tests or golden vectors for it.
 * This file restates that package's published algorithm
 * (cuda_rasterizer/{auxiliary.h,forward.cu,backward.cu,rasterizer_impl.cu})
 * and anchors on the reference's call site and conventions:
 *   /root/reference/nsr/gs_surfel.py:85-114     (call, argument meaning)
 *   /root/reference/nsr/gs_surfel.py:121-142    (allmap channel order)
 *   /root/reference/nsr/lsgm/flow_matching_trainer.py:2174-2228 (camera layout)
 *   /root/reference/utils/gs_utils/graphics_utils.py:38-85      (projection)
 * Gradients are cross-checked against fp64 autograd of an independent torch
 * restatement (oracle/surfel_torch.py) in tests/test_oracle_surfel.py; forward
 * and backward are also checked there against closed-form ray / plane geometry
 * and its float64 finite differences (test_known_answer_*), which pins that
 * this file computes what the remembered conventions say -- not that the
 * conventions are upstream's: the status stays "parity unpinned".
 *
 * Canonical arithmetic: every float op in the per-surfel stage (so_preprocess)
 * is a single IEEE-754 binary32 operation evaluated left to right with NO
 * fused multiply-add (build with -ffp-contract=off); the CUDA kernel K1 is
 * compiled the same way (--fmad=false), which is what makes radii,
 * tiles_touched, sort keys and tile ranges bit-exact between the two.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* Launchers such as torchrun export OMP_NUM_THREADS=1; the CPU baseline wants every host core. */
int so_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

#define BLOCK_X 16
#define BLOCK_Y 16
#define NEAR_N 0.2f
#define FAR_N 100.0f
#define FILTER_SIZE 0.707106f
#define FILTER_INV_SQUARE 2.0f
#define CUTOFF 3.0f

/*
 * The two judgement calls of this (unpinned) restatement are switchable, here and in the CUDA path
 * (ga_raster_set_variant, include/ga_b200.h), so pinning against upstream is a one-line flip of a default:
 *   radius formula       0: ceil(max(max(ex, ey), 3 * FilterSize))       (default; our reading of upstream forward.cu)
 *                        1: ceil(3 * max(max(ex, ey), FilterSize))       (SURVEY.md App. A item 5's alternative)
 *   quaternion gradient  0: vjp at the normalised quaternion, NOT chained through the normalisation (default)
 *                        1: chained through q / |q|
 * Both are no-ops for the reference's call except (radius) the tile rectangle of sub-pixel surfels.
 */
#ifndef SO_RADIUS_FORMULA
#define SO_RADIUS_FORMULA 0
#endif
#ifndef SO_QUAT_NORM_GRAD
#define SO_QUAT_NORM_GRAD 0
#endif
static int g_radius_formula = SO_RADIUS_FORMULA, g_quat_norm_grad = SO_QUAT_NORM_GRAD;
void so_set_variant(int radius_formula, int quat_norm_grad)
{
    g_radius_formula = radius_formula ? 1 : 0;
    g_quat_norm_grad = quat_norm_grad ? 1 : 0;
}
void so_get_variant(int *radius_formula, int *quat_norm_grad)
{
    *radius_formula = g_radius_formula; *quat_norm_grad = g_quat_norm_grad;
}

static inline int imin(int a, int b) { return a < b ? a : b; }
static inline int imax(int a, int b) { return a > b ? a : b; }

/* upstream auxiliary.h getRect: tile rectangle touched by (p, max_radius). */
static void get_rect(float px, float py, int max_radius, int gx, int gy,
                     int *x0, int *y0, int *x1, int *y1)
{
    *x0 = imin(gx, imax(0, (int)((px - (float)max_radius) / (float)BLOCK_X)));
    *y0 = imin(gy, imax(0, (int)((py - (float)max_radius) / (float)BLOCK_Y)));
    *x1 = imin(gx, imax(0, (int)((px + (float)max_radius + (float)(BLOCK_X - 1)) / (float)BLOCK_X)));
    *y1 = imin(gy, imax(0, (int)((py + (float)max_radius + (float)(BLOCK_Y - 1)) / (float)BLOCK_Y)));
}

/* quaternion (w,x,y,z) -> rotation, normalised inside (upstream quat_to_rotmat).
 * R is returned as R[row][col]. */
static void quat_to_rotmat(const float *q, float R[3][3])
{
    float n2 = ((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3];
    float s = 1.0f / sqrtf(n2);
    float w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
    R[0][0] = 1.f - 2.f * (y * y + z * z);
    R[1][0] = 2.f * (x * y + w * z);
    R[2][0] = 2.f * (x * z - w * y);
    R[0][1] = 2.f * (x * y - w * z);
    R[1][1] = 1.f - 2.f * (x * x + z * z);
    R[2][1] = 2.f * (y * z + w * x);
    R[0][2] = 2.f * (x * z + w * y);
    R[1][2] = 2.f * (y * z - w * x);
    R[2][2] = 1.f - 2.f * (x * x + y * y);
}

/*
 * Per-surfel stage (upstream forward.cu preprocessCUDA + compute_transmat +
 * compute_aabb).  viewmatrix / projmatrix are the 16 floats exactly as the
 * reference passes them (row-vector convention: point_row * M; i.e. the
 * kernel reads them column-major, /root/reference/nsr/gs_surfel.py:76-77).
 *
 * Outputs (caller allocated):
 *   transmat[9P]  rows Tu,Tv,Tw      normal_opacity[4P]   xy[2P]   depth[P]
 *   radii[P]      tiles_touched[P]   rect[4P] = x0,y0,x1,y1
 */
void so_preprocess(int P, const float *means3D, const float *opacities,
                   const float *scales, const float *rotations,
                   const float *vm, const float *pm, int H, int W,
                   float scale_modifier,
                   float *transmat, float *normal_opacity, float *xy,
                   float *depth, int *radii, int *tiles_touched, int *rect)
{
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const float hw = 0.5f * (float)W, hh = 0.5f * (float)H;
    const float cw = 0.5f * (float)(W - 1), ch = 0.5f * (float)(H - 1);
    for (int i = 0; i < P; i++) {
        radii[i] = 0;
        tiles_touched[i] = 0;
        rect[4 * i] = rect[4 * i + 1] = rect[4 * i + 2] = rect[4 * i + 3] = 0;
        depth[i] = 0.f;
        xy[2 * i] = xy[2 * i + 1] = 0.f;
        for (int k = 0; k < 9; k++) transmat[9 * i + k] = 0.f;
        for (int k = 0; k < 4; k++) normal_opacity[4 * i + k] = 0.f;

        const float px = means3D[3 * i], py = means3D[3 * i + 1], pz = means3D[3 * i + 2];
        /* in_frustum: view-space z only (upstream auxiliary.h in_frustum) */
        float vx = ((vm[0] * px + vm[4] * py) + vm[8] * pz) + vm[12];
        float vy = ((vm[1] * px + vm[5] * py) + vm[9] * pz) + vm[13];
        float vz = ((vm[2] * px + vm[6] * py) + vm[10] * pz) + vm[14];
        if (vz <= NEAR_N) continue;

        float R[3][3];
        quat_to_rotmat(rotations + 4 * i, R);
        const float sx = scale_modifier * scales[2 * i], sy = scale_modifier * scales[2 * i + 1];
        float L0[3] = {R[0][0] * sx, R[1][0] * sx, R[2][0] * sx};
        float L1[3] = {R[0][1] * sy, R[1][1] * sy, R[2][1] * sy};
        float L2[3] = {R[0][2], R[1][2], R[2][2]};

        /* B = M^T * A, A[k][j] = pm[4k+j]; only columns j = 0,1,3 are needed */
        float B[3][4];
        const int cols[3] = {0, 1, 3};
        for (int c = 0; c < 3; c++) {
            int j = cols[c];
            B[0][j] = (L0[0] * pm[j] + L0[1] * pm[4 + j]) + L0[2] * pm[8 + j];
            B[1][j] = (L1[0] * pm[j] + L1[1] * pm[4 + j]) + L1[2] * pm[8 + j];
            B[2][j] = ((px * pm[j] + py * pm[4 + j]) + pz * pm[8 + j]) + pm[12 + j];
        }
        float Tu[3], Tv[3], Tw[3];
        for (int r = 0; r < 3; r++) {
            Tu[r] = B[r][0] * hw + B[r][3] * cw;
            Tv[r] = B[r][1] * hh + B[r][3] * ch;
            Tw[r] = B[r][3];
        }
        float nx = (vm[0] * L2[0] + vm[4] * L2[1]) + vm[8] * L2[2];
        float ny = (vm[1] * L2[0] + vm[5] * L2[1]) + vm[9] * L2[2];
        float nz = (vm[2] * L2[0] + vm[6] * L2[1]) + vm[10] * L2[2];
        /* upstream stores transMats before any further rejection */
        for (int r = 0; r < 3; r++) {
            transmat[9 * i + r] = Tu[r];
            transmat[9 * i + 3 + r] = Tv[r];
            transmat[9 * i + 6 + r] = Tw[r];
        }
        /* DUAL_VISIABLE */
        float cosv = -((vx * nx + vy * ny) + vz * nz);
        if (cosv == 0.f) continue;
        float mult = cosv > 0.f ? 1.f : -1.f;
        nx = mult * nx; ny = mult * ny; nz = mult * nz;

        /* compute_aabb, cutoff = 3 */
        const float t0 = CUTOFF * CUTOFF, t1 = CUTOFF * CUTOFF, t2 = -1.0f;
        float d = (t0 * (Tw[0] * Tw[0]) + t1 * (Tw[1] * Tw[1])) + t2 * (Tw[2] * Tw[2]);
        if (d == 0.0f) continue;
        float inv = 1.0f / d;
        float f0 = inv * t0, f1 = inv * t1, f2 = inv * t2;
        float cx = (f0 * (Tu[0] * Tw[0]) + f1 * (Tu[1] * Tw[1])) + f2 * (Tu[2] * Tw[2]);
        float cy = (f0 * (Tv[0] * Tw[0]) + f1 * (Tv[1] * Tw[1])) + f2 * (Tv[2] * Tw[2]);
        float hx0 = cx * cx - ((f0 * (Tu[0] * Tu[0]) + f1 * (Tu[1] * Tu[1])) + f2 * (Tu[2] * Tu[2]));
        float hy0 = cy * cy - ((f0 * (Tv[0] * Tv[0]) + f1 * (Tv[1] * Tv[1])) + f2 * (Tv[2] * Tv[2]));
        float ex = sqrtf(fmaxf(1e-4f, hx0)), ey = sqrtf(fmaxf(1e-4f, hy0));
        float radius = g_radius_formula == 0 ? ceilf(fmaxf(fmaxf(ex, ey), CUTOFF * FILTER_SIZE))
                                             : ceilf(CUTOFF * fmaxf(fmaxf(ex, ey), FILTER_SIZE));

        int x0, y0, x1, y1;
        get_rect(cx, cy, (int)radius, gx, gy, &x0, &y0, &x1, &y1);
        if ((x1 - x0) * (y1 - y0) == 0) continue;

        depth[i] = vz;
        radii[i] = (int)radius;
        xy[2 * i] = cx; xy[2 * i + 1] = cy;
        normal_opacity[4 * i] = nx; normal_opacity[4 * i + 1] = ny;
        normal_opacity[4 * i + 2] = nz; normal_opacity[4 * i + 3] = opacities[i];
        tiles_touched[i] = (y1 - y0) * (x1 - x0);
        rect[4 * i] = x0; rect[4 * i + 1] = y0; rect[4 * i + 2] = x1; rect[4 * i + 3] = y1;
    }
}

typedef struct { uint64_t key; uint32_t id; } kv_t;
static int kv_cmp(const void *a, const void *b)
{
    const kv_t *x = (const kv_t *)a, *y = (const kv_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->id < y->id ? -1 : (x->id > y->id);   /* == stable LSD radix order */
}

/*
 * Binning (upstream rasterizer_impl.cu duplicateWithKeys + SortPairs +
 * identifyTileRanges).  D = sum(tiles_touched).  keys[D] = (tile<<32)|depth bits,
 * ids[D] = surfel index, ranges[2*tiles] = [start,end).
 * A stable sort of the duplication order (ascending surfel index) by key is
 * the same as sorting by (key, id).  Returns D.
 */
int64_t so_bin(int P, const float *depth, const int *radii, const int *rect,
               int H, int W, int64_t D, uint64_t *keys, uint32_t *ids, int32_t *ranges)
{
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    kv_t *kv = (kv_t *)malloc(sizeof(kv_t) * (size_t)(D > 0 ? D : 1));
    int64_t off = 0;
    for (int i = 0; i < P; i++) {
        if (radii[i] <= 0) continue;
        uint32_t db; memcpy(&db, depth + i, 4);
        for (int y = rect[4 * i + 1]; y < rect[4 * i + 3]; y++)
            for (int x = rect[4 * i]; x < rect[4 * i + 2]; x++) {
                uint64_t key = (uint64_t)(y * gx + x);
                key = (key << 32) | (uint64_t)db;
                if (off < D) { kv[off].key = key; kv[off].id = (uint32_t)i; }
                off++;
            }
    }
    if (off != D) { free(kv); return -off; }
    qsort(kv, (size_t)D, sizeof(kv_t), kv_cmp);
    for (int t = 0; t < gx * gy; t++) ranges[2 * t] = ranges[2 * t + 1] = 0;
    for (int64_t j = 0; j < D; j++) {
        keys[j] = kv[j].key; ids[j] = kv[j].id;
        int tile = (int)(kv[j].key >> 32);
        if (j == 0) ranges[2 * tile] = 0;
        else {
            int prev = (int)(kv[j - 1].key >> 32);
            if (prev != tile) { ranges[2 * prev + 1] = (int32_t)j; ranges[2 * tile] = (int32_t)j; }
        }
        if (j == D - 1) ranges[2 * tile + 1] = (int32_t)D;
    }
    free(kv);
    return D;
}

/*
 * Forward composite (upstream forward.cu renderCUDA).  One "thread" per pixel,
 * iterating the pixel's tile list front to back.
 *   out_color[3HW], out_allmap[7HW] (0 depth*w, 1 alpha, 2-4 normal, 5 median
 *   depth, 6 distortion), state: final_T[3HW] = {T, M1, M2},
 *   n_contrib[2HW] = {last contributor, median contributor}.
 */
void so_render_forward(int H, int W, const int32_t *ranges, const uint32_t *ids,
                       const float *xy, const float *transmat,
                       const float *normal_opacity, const float *colors,
                       const float *bg, float *out_color, float *out_allmap,
                       float *final_T, int32_t *n_contrib)
{
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const int HW = H * W;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; tile++) {
        const int tx = tile % gx, ty = tile / gx;
        const int r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        for (int ly = 0; ly < BLOCK_Y; ly++)
            for (int lx = 0; lx < BLOCK_X; lx++) {
                const int pxi = tx * BLOCK_X + lx, pyi = ty * BLOCK_Y + ly;
                if (pxi >= W || pyi >= H) continue;
                const int pix = W * pyi + pxi;
                const float pfx = (float)pxi, pfy = (float)pyi;
                float T = 1.0f, C[3] = {0, 0, 0}, N[3] = {0, 0, 0};
                float Dacc = 0, M1 = 0, M2 = 0, dist = 0, median_depth = 0;
                int contributor = 0, last_contributor = 0, median_contributor = -1;
                for (int j = r0; j < r1; j++) {
                    contributor++;
                    const uint32_t g = ids[j];
                    const float *Tu = transmat + 9 * g, *Tv = Tu + 3, *Tw = Tu + 6;
                    float k0 = pfx * Tw[0] - Tu[0], k1 = pfx * Tw[1] - Tu[1], k2 = pfx * Tw[2] - Tu[2];
                    float l0 = pfy * Tw[0] - Tv[0], l1 = pfy * Tw[1] - Tv[1], l2 = pfy * Tw[2] - Tv[2];
                    float p0 = k1 * l2 - k2 * l1, p1 = k2 * l0 - k0 * l2, p2 = k0 * l1 - k1 * l0;
                    if (p2 == 0.0f) continue;
                    float s0 = p0 / p2, s1 = p1 / p2;
                    float rho3d = s0 * s0 + s1 * s1;
                    float dx = xy[2 * g] - pfx, dy = xy[2 * g + 1] - pfy;
                    float rho2d = FILTER_INV_SQUARE * (dx * dx + dy * dy);
                    float rho = fminf(rho3d, rho2d);
                    float depth = (rho3d <= rho2d) ? (s0 * Tw[0] + s1 * Tw[1]) + Tw[2] : Tw[2];
                    if (depth < NEAR_N) continue;
                    const float *no = normal_opacity + 4 * g;
                    float power = -0.5f * rho;
                    if (power > 0.0f) continue;
                    float alpha = fminf(0.99f, no[3] * expf(power));
                    if (alpha < 1.0f / 255.0f) continue;
                    float test_T = T * (1 - alpha);
                    if (test_T < 0.0001f) break;            /* done = true */
                    float w = alpha * T;
                    float A = 1 - T;
                    float m = FAR_N / (FAR_N - NEAR_N) * (1 - NEAR_N / depth);
                    dist += (m * m * A + M2 - 2 * m * M1) * w;
                    Dacc += depth * w;
                    M1 += m * w;
                    M2 += m * m * w;
                    if (T > 0.5f) { median_depth = depth; median_contributor = contributor; }
                    for (int c = 0; c < 3; c++) N[c] += no[c] * w;
                    for (int c = 0; c < 3; c++) C[c] += colors[3 * g + c] * w;
                    T = test_T;
                    last_contributor = contributor;
                }
                final_T[pix] = T; final_T[pix + HW] = M1; final_T[pix + 2 * HW] = M2;
                n_contrib[pix] = last_contributor; n_contrib[pix + HW] = median_contributor;
                for (int c = 0; c < 3; c++) out_color[c * HW + pix] = C[c] + T * bg[c];
                out_allmap[0 * HW + pix] = Dacc;
                out_allmap[1 * HW + pix] = 1 - T;
                for (int c = 0; c < 3; c++) out_allmap[(2 + c) * HW + pix] = N[c];
                out_allmap[5 * HW + pix] = median_depth;
                out_allmap[6 * HW + pix] = dist;
            }
    }
}

static inline void atomic_add_d(double *p, double v)
{
#pragma omp atomic
    *p += v;
}

/*
 * Backward composite (upstream backward.cu renderCUDA): back-to-front over the
 * same tile list; per-pixel math in binary32 like upstream, accumulation over
 * pixels in binary64 (upstream uses float atomicAdd, order undefined).
 * Gradient buffers are double and must be zeroed by the caller:
 *   dL_dtransmat[9P], dL_dmean2D[2P], dL_dnormal[3P], dL_dopacity[P], dL_dcolor[3P]
 */
void so_render_backward(int H, int W, const int32_t *ranges, const uint32_t *ids,
                        const float *xy, const float *transmat,
                        const float *normal_opacity, const float *colors,
                        const float *bg, const float *final_T, const int32_t *n_contrib,
                        const float *dL_dcolor_px, const float *dL_dallmap_px,
                        double *dL_dtransmat, double *dL_dmean2D, double *dL_dnormal,
                        double *dL_dopacity, double *dL_dcolor)
{
    const int gx = (W + BLOCK_X - 1) / BLOCK_X, gy = (H + BLOCK_Y - 1) / BLOCK_Y;
    const int HW = H * W;
#pragma omp parallel for schedule(dynamic, 1)
    for (int tile = 0; tile < gx * gy; tile++) {
        const int tx = tile % gx, ty = tile / gx;
        const int r0 = ranges[2 * tile], r1 = ranges[2 * tile + 1];
        for (int ly = 0; ly < BLOCK_Y; ly++)
            for (int lx = 0; lx < BLOCK_X; lx++) {
                const int pxi = tx * BLOCK_X + lx, pyi = ty * BLOCK_Y + ly;
                if (pxi >= W || pyi >= H) continue;
                const int pix = W * pyi + pxi;
                const float pfx = (float)pxi, pfy = (float)pyi;
                const float T_final = final_T[pix];
                float T = T_final;
                const int last_contributor = n_contrib[pix];
                const int median_contributor = n_contrib[pix + HW];
                float accum_rec[3] = {0, 0, 0}, dL_dpixel[3];
                for (int c = 0; c < 3; c++) dL_dpixel[c] = dL_dcolor_px[c * HW + pix];
                const float dL_ddepth = dL_dallmap_px[0 * HW + pix];
                const float dL_daccum = dL_dallmap_px[1 * HW + pix];
                float dL_dnormal2D[3];
                for (int c = 0; c < 3; c++) dL_dnormal2D[c] = dL_dallmap_px[(2 + c) * HW + pix];
                const float dL_dmedian_depth = dL_dallmap_px[5 * HW + pix];
                const float dL_dreg = dL_dallmap_px[6 * HW + pix];
                float last_depth = 0, last_normal[3] = {0, 0, 0};
                float accum_depth_rec = 0, accum_alpha_rec = 0, accum_normal_rec[3] = {0, 0, 0};
                const float final_D = final_T[pix + HW], final_D2 = final_T[pix + 2 * HW];
                const float final_A = 1 - T_final;
                float last_dL_dT = 0, last_alpha = 0, last_color[3] = {0, 0, 0};
                float bg_dot_dpixel = 0;
                for (int c = 0; c < 3; c++) bg_dot_dpixel += bg[c] * dL_dpixel[c];

                int contributor = r1 - r0;
                for (int j = r1 - 1; j >= r0; j--) {
                    contributor--;
                    if (contributor >= last_contributor) continue;
                    const uint32_t g = ids[j];
                    const float *Tu = transmat + 9 * g, *Tv = Tu + 3, *Tw = Tu + 6;
                    float k0 = pfx * Tw[0] - Tu[0], k1 = pfx * Tw[1] - Tu[1], k2 = pfx * Tw[2] - Tu[2];
                    float l0 = pfy * Tw[0] - Tv[0], l1 = pfy * Tw[1] - Tv[1], l2 = pfy * Tw[2] - Tv[2];
                    float p0 = k1 * l2 - k2 * l1, p1 = k2 * l0 - k0 * l2, p2 = k0 * l1 - k1 * l0;
                    if (p2 == 0.0f) continue;
                    float s0 = p0 / p2, s1 = p1 / p2;
                    float rho3d = s0 * s0 + s1 * s1;
                    float dx = xy[2 * g] - pfx, dy = xy[2 * g + 1] - pfy;
                    float rho2d = FILTER_INV_SQUARE * (dx * dx + dy * dy);
                    float rho = fminf(rho3d, rho2d);
                    float c_d = (rho3d <= rho2d) ? (s0 * Tw[0] + s1 * Tw[1]) + Tw[2] : Tw[2];
                    if (c_d < NEAR_N) continue;
                    const float *no = normal_opacity + 4 * g;
                    const float opa = no[3];
                    float power = -0.5f * rho;
                    if (power > 0.0f) continue;
                    const float G = expf(power);
                    const float alpha = fminf(0.99f, opa * G);
                    if (alpha < 1.0f / 255.0f) continue;

                    T = T / (1.f - alpha);
                    const float w = alpha * T;
                    float dL_dalpha = 0.0f;
                    for (int c = 0; c < 3; c++) {
                        const float col = colors[3 * g + c];
                        accum_rec[c] = last_alpha * last_color[c] + (1.f - last_alpha) * accum_rec[c];
                        last_color[c] = col;
                        dL_dalpha += (col - accum_rec[c]) * dL_dpixel[c];
                        atomic_add_d(&dL_dcolor[3 * g + c], (double)(w * dL_dpixel[c]));
                    }
                    float dL_dz = 0.0f, dL_dweight = 0.0f;
                    const float m_d = FAR_N / (FAR_N - NEAR_N) * (1 - NEAR_N / c_d);
                    const float dmd_dd = (FAR_N * NEAR_N) / ((FAR_N - NEAR_N) * c_d * c_d);
                    if (contributor == median_contributor - 1) dL_dz += dL_dmedian_depth;
                    dL_dweight += (final_D2 + m_d * m_d * final_A - 2 * m_d * final_D) * dL_dreg;
                    dL_dalpha += dL_dweight - last_dL_dT;
                    last_dL_dT = dL_dweight * alpha + (1 - alpha) * last_dL_dT;
                    const float dL_dmd = 2.0f * (T * alpha) * (m_d * final_A - final_D) * dL_dreg;
                    dL_dz += dL_dmd * dmd_dd;

                    accum_depth_rec = last_alpha * last_depth + (1.f - last_alpha) * accum_depth_rec;
                    last_depth = c_d;
                    dL_dalpha += (c_d - accum_depth_rec) * dL_ddepth;
                    accum_alpha_rec = last_alpha * 1.0f + (1.f - last_alpha) * accum_alpha_rec;
                    dL_dalpha += (1 - accum_alpha_rec) * dL_daccum;
                    for (int c = 0; c < 3; c++) {
                        accum_normal_rec[c] = last_alpha * last_normal[c] + (1.f - last_alpha) * accum_normal_rec[c];
                        last_normal[c] = no[c];
                        dL_dalpha += (no[c] - accum_normal_rec[c]) * dL_dnormal2D[c];
                        atomic_add_d(&dL_dnormal[3 * g + c], (double)(w * dL_dnormal2D[c]));
                    }
                    dL_dalpha *= T;
                    last_alpha = alpha;
                    dL_dalpha += (-T_final / (1.f - alpha)) * bg_dot_dpixel;

                    /* the 0.99 clamp is passed straight through, as upstream does */
                    const float dL_dG = opa * dL_dalpha;
                    dL_dz += w * dL_ddepth;

                    if (rho3d <= rho2d) {
                        const float dL_ds0 = dL_dG * -G * s0 + dL_dz * Tw[0];
                        const float dL_ds1 = dL_dG * -G * s1 + dL_dz * Tw[1];
                        const float dsx_pz = dL_ds0 / p2, dsy_pz = dL_ds1 / p2;
                        const float q0 = dsx_pz, q1 = dsy_pz, q2 = -(dsx_pz * s0 + dsy_pz * s1);
                        /* dL_dk = cross(l, dL_dp), dL_dl = cross(dL_dp, k) */
                        const float dk0 = l1 * q2 - l2 * q1, dk1 = l2 * q0 - l0 * q2, dk2 = l0 * q1 - l1 * q0;
                        const float dl0 = q1 * k2 - q2 * k1, dl1 = q2 * k0 - q0 * k2, dl2 = q0 * k1 - q1 * k0;
                        atomic_add_d(&dL_dtransmat[9 * g + 0], (double)(-dk0));
                        atomic_add_d(&dL_dtransmat[9 * g + 1], (double)(-dk1));
                        atomic_add_d(&dL_dtransmat[9 * g + 2], (double)(-dk2));
                        atomic_add_d(&dL_dtransmat[9 * g + 3], (double)(-dl0));
                        atomic_add_d(&dL_dtransmat[9 * g + 4], (double)(-dl1));
                        atomic_add_d(&dL_dtransmat[9 * g + 5], (double)(-dl2));
                        atomic_add_d(&dL_dtransmat[9 * g + 6], (double)(pfx * dk0 + pfy * dl0 + dL_dz * s0));
                        atomic_add_d(&dL_dtransmat[9 * g + 7], (double)(pfx * dk1 + pfy * dl1 + dL_dz * s1));
                        atomic_add_d(&dL_dtransmat[9 * g + 8], (double)(pfx * dk2 + pfy * dl2 + dL_dz));
                    } else {
                        const float dG_ddelx = -G * FILTER_INV_SQUARE * dx;
                        const float dG_ddely = -G * FILTER_INV_SQUARE * dy;
                        atomic_add_d(&dL_dmean2D[2 * g + 0], (double)(dL_dG * dG_ddelx));
                        atomic_add_d(&dL_dmean2D[2 * g + 1], (double)(dL_dG * dG_ddely));
                        atomic_add_d(&dL_dtransmat[9 * g + 8], (double)dL_dz);
                    }
                    atomic_add_d(&dL_dopacity[g], (double)(G * dL_dalpha));
                }
            }
    }
}

/*
 * Per-surfel backward (upstream backward.cu compute_transmat_aabb +
 * preprocessCUDA): chains dL/dT (+ the mean2D / AABB-centre path) and
 * dL/dnormal to means3D, scales, rotations.  Differences to upstream, both
 * no-ops for the reference's call (scale_modifier == 1, unit quaternions):
 * the scale modifier is differentiated through, and like upstream the
 * quaternion gradient does not chain through the in-kernel normalisation.
 * Outputs are ACCUMULATED (+=) so a multi-view sum is one pass per view.
 */
void so_preprocess_backward(int P, const float *means3D, const float *scales,
                            const float *rotations, const float *vm, const float *pm,
                            int H, int W, float scale_modifier, const int *radii,
                            const float *transmat,
                            const double *dL_dtransmat, const double *dL_dmean2D,
                            const double *dL_dnormal,
                            double *dL_dmeans3D, double *dL_dscales, double *dL_drots)
{
    const double hw = 0.5 * W, hh = 0.5 * H, cw = 0.5 * (W - 1), ch = 0.5 * (H - 1);
    for (int i = 0; i < P; i++) {
        if (!(radii[i] > 0)) continue;
        double G[3][3]; /* G[c][r] = dL/dT[c][r], c: 0=Tu 1=Tv 2=Tw */
        for (int c = 0; c < 3; c++)
            for (int r = 0; r < 3; r++) G[c][r] = dL_dtransmat[9 * i + 3 * c + r];
        const float *Tm = transmat + 9 * i;
        const double gmx = dL_dmean2D[2 * i], gmy = dL_dmean2D[2 * i + 1];
        if (gmx != 0 || gmy != 0) {
            const double t[3] = {9.0, 9.0, -1.0};
            double Tu[3], Tv[3], Tw[3];
            for (int r = 0; r < 3; r++) { Tu[r] = Tm[r]; Tv[r] = Tm[3 + r]; Tw[r] = Tm[6 + r]; }
            double d = 0;
            for (int r = 0; r < 3; r++) d += t[r] * Tw[r] * Tw[r];
            double f[3], dL_df[3], dL_dd = 0;
            for (int r = 0; r < 3; r++) f[r] = t[r] / d;
            for (int r = 0; r < 3; r++) {
                G[0][r] += gmx * f[r] * Tw[r];
                G[1][r] += gmy * f[r] * Tw[r];
                G[2][r] += gmx * f[r] * Tu[r] + gmy * f[r] * Tv[r];
                dL_df[r] = gmx * Tu[r] * Tw[r] + gmy * Tv[r] * Tw[r];
                dL_dd += dL_df[r] * f[r];
            }
            dL_dd *= (-1.0 / d);
            for (int r = 0; r < 3; r++) G[2][r] += dL_dd * t[r] * Tw[r] * 2.0;
        }
        /* dL_dM[r][k] = sum_c AN[k][c] * G[c][r] ; AN = A * ndc2pix */
        double dM[3][3];
        for (int r = 0; r < 3; r++)
            for (int k = 0; k < 3; k++) {
                double an0 = pm[4 * k + 0] * hw + pm[4 * k + 3] * cw;
                double an1 = pm[4 * k + 1] * hh + pm[4 * k + 3] * ch;
                double an2 = pm[4 * k + 3];
                dM[r][k] = an0 * G[0][r] + an1 * G[1][r] + an2 * G[2][r];
            }
        float Rf[3][3];
        quat_to_rotmat(rotations + 4 * i, Rf);
        const double sx = (double)scale_modifier * scales[2 * i], sy = (double)scale_modifier * scales[2 * i + 1];
        /* normal path: recompute the dual-visible sign */
        const float px = means3D[3 * i], py = means3D[3 * i + 1], pz = means3D[3 * i + 2];
        float vx = ((vm[0] * px + vm[4] * py) + vm[8] * pz) + vm[12];
        float vy = ((vm[1] * px + vm[5] * py) + vm[9] * pz) + vm[13];
        float vz = ((vm[2] * px + vm[6] * py) + vm[10] * pz) + vm[14];
        float nx = (vm[0] * Rf[0][2] + vm[4] * Rf[1][2]) + vm[8] * Rf[2][2];
        float ny = (vm[1] * Rf[0][2] + vm[5] * Rf[1][2]) + vm[9] * Rf[2][2];
        float nz = (vm[2] * Rf[0][2] + vm[6] * Rf[1][2]) + vm[10] * Rf[2][2];
        float cosv = -((vx * nx + vy * ny) + vz * nz);
        const double mult = cosv > 0.f ? 1.0 : -1.0;
        const double gn[3] = {dL_dnormal[3 * i], dL_dnormal[3 * i + 1], dL_dnormal[3 * i + 2]};
        double dtn[3];
        dtn[0] = mult * (vm[0] * gn[0] + vm[1] * gn[1] + vm[2] * gn[2]);
        dtn[1] = mult * (vm[4] * gn[0] + vm[5] * gn[1] + vm[6] * gn[2]);
        dtn[2] = mult * (vm[8] * gn[0] + vm[9] * gn[1] + vm[10] * gn[2]);
        /* dL_dR[row k][col c] */
        double dR[3][3];
        for (int k = 0; k < 3; k++) { dR[k][0] = dM[0][k] * sx; dR[k][1] = dM[1][k] * sy; dR[k][2] = dtn[k]; }
        double ds0 = 0, ds1 = 0;
        for (int k = 0; k < 3; k++) { ds0 += dM[0][k] * Rf[k][0]; ds1 += dM[1][k] * Rf[k][1]; }
        dL_dscales[2 * i] += scale_modifier * ds0;
        dL_dscales[2 * i + 1] += scale_modifier * ds1;
        for (int k = 0; k < 3; k++) dL_dmeans3D[3 * i + k] += dM[2][k];
        /* quaternion vjp at the normalised quaternion (w,x,y,z) */
        const float *q = rotations + 4 * i;
        double n2 = (double)q[0] * q[0] + (double)q[1] * q[1] + (double)q[2] * q[2] + (double)q[3] * q[3];
        double s = 1.0 / sqrt(n2);
        double w = q[0] * s, x = q[1] * s, y = q[2] * s, z = q[3] * s;
        /* R[r][c] as in quat_to_rotmat; d/dw etc. */
        double gw = 2 * (z * (dR[1][0] - dR[0][1]) + y * (dR[0][2] - dR[2][0]) + x * (dR[2][1] - dR[1][2]));
        double gxq = 2 * (-2 * x * (dR[1][1] + dR[2][2]) + y * (dR[1][0] + dR[0][1]) + z * (dR[2][0] + dR[0][2]) + w * (dR[2][1] - dR[1][2]));
        double gyq = 2 * (x * (dR[1][0] + dR[0][1]) - 2 * y * (dR[0][0] + dR[2][2]) + z * (dR[2][1] + dR[1][2]) + w * (dR[0][2] - dR[2][0]));
        double gzq = 2 * (x * (dR[2][0] + dR[0][2]) + y * (dR[2][1] + dR[1][2]) - 2 * z * (dR[0][0] + dR[1][1]) + w * (dR[1][0] - dR[0][1]));
        if (g_quat_norm_grad) {          /* chain through q_hat = q / |q|: g_raw = (g - q_hat (q_hat . g)) / |q| */
            const double dot = w * gw + x * gxq + y * gyq + z * gzq;
            gw = (gw - w * dot) * s; gxq = (gxq - x * dot) * s; gyq = (gyq - y * dot) * s; gzq = (gzq - z * dot) * s;
        }
        dL_drots[4 * i] += gw; dL_drots[4 * i + 1] += gxq; dL_drots[4 * i + 2] += gyq; dL_drots[4 * i + 3] += gzq;
    }
}
