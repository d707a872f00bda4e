// tcgen05 GEMM for the DiT denoiser: C[M,N] = A[M,K] * W[N,K]^T, bf16 operands,
// fp32 accumulation in TMEM, fused epilogues.  Replaces the cuBLAS nn.Linear +
// separate bias / GELU / gate / residual / qk-RMSNorm / permute kernels of
// /root/reference/dit/dit_models_xformers.py:765-787,
// /root/reference/vit/vision_transformer.py:215-303 and
// /root/reference/ldm/modules/attention.py:484-561.
//
// Persistent kernel, one CTA per SM looping over 128 x BN output tiles:
//   warp 0   : TMA producer  (cp.async.bulk.tensor, 128B-swizzled K-major tiles)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2-9: epilogue, one accumulator row per thread (tcgen05.ld 32x32b), half the columns per warp
// smem ring of kStages {A 128x64, W BNx64} bf16 tiles with full/empty mbarriers; the TMEM accumulator is
// double buffered (acc_full/acc_empty) so the epilogue of one tile overlaps the main loop of the next.
#include "../../include/ga_b200.h"
#include "device_once.cuh"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int BM = 128, BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kThreads = 64 + 32 * kEpiWarps;          // TMA warp + MMA warp + 8 epilogue warps

template <int BN> struct GemmCfg {
    static constexpr int kStages = (BN >= 256) ? 3 : (BN >= 192 ? 4 : (BN >= 128 ? 5 : 7));    // + 36 KB of epilogue staging
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kRing = kStages * (kABytes + kBBytes);
    static constexpr int kSmem = kRing + 1024 + 36 * 1024;
    static constexpr int kTmemCols = (2 * BN <= 128) ? 128 : (2 * BN <= 256 ? 256 : 512);     // double-buffered accumulator (power of two)
};

// exact-erf GELU to ~3e-7 (Abramowitz-Stegun 7.1.28: erf x = 1 - (1 + a1 x + .. + a6 x^6)^-16), 1 MUFU + ~16 FP32
// ops per element; the library erff() costs about twice that and would make the GELU epilogue slower than the MMAs.
__device__ __forceinline__ float gelu_erf(float v)
{
    const float x = fabsf(v) * 0.70710678118654752f;
    float p = 0.0000430638f;
    p = fmaf(p, x, 0.0002765672f);
    p = fmaf(p, x, 0.0001520143f);
    p = fmaf(p, x, 0.0092705272f);
    p = fmaf(p, x, 0.0422820123f);
    p = fmaf(p, x, 0.0705230784f);
    p = fmaf(p, x, 1.0f);
    p = p * p; p = p * p; p = p * p; p = p * p;
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(p));
    const float erf_abs = 1.0f - r;
    return 0.5f * v * (1.0f + copysignf(erf_abs, v));
}

__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// ---- epilogue ---------------------------------------------------------------
// tcgen05.ld hands every thread ONE accumulator row, so storing straight from that layout makes each warp store
// touch 32 different rows (32 sectors per instruction) -- measured, the store queue then bounds the whole GEMM
// (epilogue warps busy ~90 % of the time, tensor pipe ~30 %).  Each epilogue warp therefore owns a 4 KB staging
// buffer (32 rows x 128 B, 16-byte chunks XOR-swizzled by row&7 so both phases are bank-conflict free):
//   phase A (thread = row)          : TMEM -> registers -> staging
//   phase B (8 lanes = one 128 B row): staging -> bias / GELU / gate / residual -> coalesced global stores
// The epilogue mode is a template parameter: with a run-time switch inside the unrolled loops the kernel grew to
// 5k instructions and a fifth of the epilogue's stall samples were instruction-cache misses.
constexpr int kStageBytes = 4096;          // per epilogue warp
constexpr int kHeadParams = 128;           // floats per epilogue warp: 64 bias + 64 norm weights (HEADS mode)

__device__ __forceinline__ void stage_rows(uint32_t stg, int lane, const uint32_t (&r)[32])
{
    const uint32_t row = stg + lane * 128;
#pragma unroll
    for (int ch = 0; ch < 8; ch++)
        sts128(row + ((ch ^ (lane & 7)) << 4), r[4 * ch], r[4 * ch + 1], r[4 * ch + 2], r[4 * ch + 3]);
}

// per-lane column parameters of phase B (this lane's 4 columns nn .. nn+3), fetched BEFORE the TMEM load so
// their latency hides under it
struct ColParams { float bias[4]; float gate[4]; };

template <int MODE>
__device__ __forceinline__ void load_col_params(const GaGemmEpilogue &ep, int lane, int m0w, int n, int N, ColParams &cp)
{
    const int nn = n + (lane & 7) * 4;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        cp.bias[j] = (ep.bias && nn + j < N) ? __ldg(ep.bias + nn + j) : 0.f;
        cp.gate[j] = 1.f;
    }
    if (MODE == GA_EPI_RESID_GATE_F32 && ep.gate) {
        const int b = m0w / ep.rows_per_batch;           // used when the warp's 32 rows sit in one batch element
#pragma unroll
        for (int j = 0; j < 4; j++) if (nn + j < N) cp.gate[j] = __ldg(ep.gate + (size_t)b * ep.gate_ld + nn + j);
    }
}

// 32 fp32 accumulator columns [n, n+32) of the warp's 32 rows [m0w, m0w+32), already staged by stage_rows()
template <int MODE>
__device__ __forceinline__ void epilogue32(const GaGemmEpilogue &ep, uint32_t stg, int lane, int m0w, int n, int M, int N,
                                           const ColParams &cp)
{
    const int ch = lane & 7, rsub = lane >> 3;
    const int nn = n + ch * 4;                                 // this lane's 4 columns
    const bool vec = (nn + 4 <= N) && (ep.ld_out % 4) == 0;
    bool one_batch = true;
    if (MODE == GA_EPI_RESID_GATE_F32 && ep.gate)
        one_batch = (m0w + 31) / ep.rows_per_batch == m0w / ep.rows_per_batch;
    uint4 acc[8];
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int rr = it * 4 + rsub;
        acc[it] = lds128(stg + rr * 128 + ((ch ^ (rr & 7)) << 4));
    }
    // residual rows are read up front, all eight in flight at once: interleaved with the stores below the compiler
    // must assume they alias and the loop degenerates into eight serial L2 round trips per chunk
    float4 res[8];
    if (MODE == GA_EPI_RESID_GATE_F32 && vec) {
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int m = m0w + it * 4 + rsub;
            res[it] = (m < M) ? *reinterpret_cast<const float4 *>(reinterpret_cast<const float *>(ep.out) +
                                                                   (size_t)m * ep.ld_out + nn)
                              : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    if (nn >= N) return;
#pragma unroll
    for (int it = 0; it < 8; it++) {
        const int m = m0w + it * 4 + rsub;
        if (m >= M) continue;
        float v[4] = {__uint_as_float(acc[it].x) + cp.bias[0], __uint_as_float(acc[it].y) + cp.bias[1],
                      __uint_as_float(acc[it].z) + cp.bias[2], __uint_as_float(acc[it].w) + cp.bias[3]};
        if (MODE == GA_EPI_BF16 || MODE == GA_EPI_GELU_BF16) {
            if (MODE == GA_EPI_GELU_BF16) {
#pragma unroll
                for (int j = 0; j < 4; j++) v[j] = gelu_erf(v[j]);
            }
            __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(ep.out) + (size_t)m * ep.ld_out + nn;
            if (vec) {
                *reinterpret_cast<uint2 *>(dst) = make_uint2(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]));
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) if (nn + j < N) dst[j] = __float2bfloat16(v[j]);
            }
        } else if (MODE == GA_EPI_F32) {
            float *dst = reinterpret_cast<float *>(ep.out) + (size_t)m * ep.ld_out + nn;
            if (vec) {
                *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) if (nn + j < N) dst[j] = v[j];
            }
        } else if (MODE == GA_EPI_RESID_GATE_F32) {
            // x[m, n] += gate[b, n] * (acc + bias)
            float *dst = reinterpret_cast<float *>(ep.out) + (size_t)m * ep.ld_out + nn;
            float g[4] = {cp.gate[0], cp.gate[1], cp.gate[2], cp.gate[3]};
            if (!one_batch) {
                const float *gp = ep.gate + (size_t)(m / ep.rows_per_batch) * ep.gate_ld + nn;
#pragma unroll
                for (int j = 0; j < 4; j++) if (nn + j < N) g[j] = __ldg(gp + j);
            }
            if (vec) {
                float4 x = res[it];
                x.x += g[0] * v[0]; x.y += g[1] * v[1]; x.z += g[2] * v[2]; x.w += g[3] * v[3];
                *reinterpret_cast<float4 *>(dst) = x;
            } else {
#pragma unroll
                for (int j = 0; j < 4; j++) if (nn + j < N) dst[j] += g[j] * v[j];
            }
        }
    }
}

// HEADS epilogue for one head (64 columns [n, n+64)) of the warp's 32 rows: two TMEM passes keep the live set at
// 32 values.  which = n / inner (+ first_part): 0 q, 1 k, 2 v ; head = (n % inner) / 64   ("(K H D)" column
// layout, vit/vision_transformer.py:191,255).  q/k: per-head RMSNorm (dit/norm.py:27-40), fp32, then * weight;
// a token's 64 bf16 are one 128 B line of the [B, H, tok_pitch, 64] layout.  v is stored transposed per head,
// [B, H, 64, tok_pitch], so that P*V is a K-major x K-major contraction: staged as [d][32 tokens].
__device__ __forceinline__ void epilogue_head(const GaGemmEpilogue &ep, uint32_t stg, float *hp, uint32_t taddr, int lane,
                                              int m0w, int n, int M)
{
    const int inner = ep.heads * 64;
    const int which = n / inner + ep.first_part;
    const int head = (n % inner) / 64;
    const float *w = which == 0 ? ep.qn_w : (which == 1 ? ep.kn_w : nullptr);
    hp[lane] = ep.bias ? __ldg(ep.bias + n + lane) : 0.f;
    hp[32 + lane] = ep.bias ? __ldg(ep.bias + n + 32 + lane) : 0.f;
    hp[64 + lane] = w ? __ldg(w + lane) : 1.f;
    hp[96 + lane] = w ? __ldg(w + 32 + lane) : 1.f;
    __syncwarp();
    const int rpb = ep.rows_per_batch;
    const int b_first = m0w / rpb;
    const bool one_batch = (m0w + 31) / rpb == b_first;
    uint32_t r[32];
    float rs = 1.0f;
    if (w) {
        float ss = 0.f;
#pragma unroll
        for (int h2 = 0; h2 < 2; h2++) {
            tmem_ld_32x32b_x32(taddr + h2 * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i++) {
                const float x = __uint_as_float(r[i]) + hp[h2 * 32 + i];
                ss += x * x;
            }
        }
        rs = rsqrtf(ss * (1.0f / 64.0f) + ep.eps);
    }
    if (which <= 1) {
#pragma unroll
        for (int h2 = 0; h2 < 2; h2++) {
            tmem_ld_32x32b_x32(taddr + h2 * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int c4 = 0; c4 < 4; c4++) {
                uint32_t pk[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const int i = c4 * 8 + j * 2;
                    const float x0 = (__uint_as_float(r[i]) + hp[h2 * 32 + i]) * rs * hp[64 + h2 * 32 + i];
                    const float x1 = (__uint_as_float(r[i + 1]) + hp[h2 * 32 + i + 1]) * rs * hp[64 + h2 * 32 + i + 1];
                    pk[j] = pack_bf16(x0, x1);
                }
                sts128(stg + lane * 128 + (((h2 * 4 + c4) ^ (lane & 7)) << 4), pk[0], pk[1], pk[2], pk[3]);
            }
        }
        warp_sync_smem();
        __nv_bfloat16 *base = reinterpret_cast<__nv_bfloat16 *>(which == 0 ? ep.q : ep.k);
        const int ch = lane & 7, rsub = lane >> 3;
        uint4 x[8];
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int rr = it * 4 + rsub;
            x[it] = lds128(stg + rr * 128 + ((ch ^ (rr & 7)) << 4));
        }
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int m = m0w + it * 4 + rsub;
            if (m >= M) continue;
            const int b = one_batch ? b_first : m / rpb;
            const int t = m - b * rpb;
            *reinterpret_cast<uint4 *>(base + (((size_t)b * ep.heads + head) * ep.tok_pitch + t) * 64 + ch * 8) = x[it];
        }
    } else {
        const int t0 = m0w - b_first * rpb;
        const bool fast = one_batch && (m0w + 32 <= M) && (ep.tok_pitch % 8) == 0 && (t0 % 8) == 0;
        __nv_bfloat16 *vt = reinterpret_cast<__nv_bfloat16 *>(ep.vt);
#pragma unroll
        for (int h2 = 0; h2 < 2; h2++) {
            tmem_ld_32x32b_x32(taddr + h2 * 32, r);
            tmem_ld_wait();
            if (fast) {
#pragma unroll
                for (int i = 0; i < 32; i++) {
                    const __nv_bfloat16 hv = __float2bfloat16(__uint_as_float(r[i]) + hp[h2 * 32 + i]);
                    asm volatile("st.shared.b16 [%0], %1;\n" ::"r"(stg + (uint32_t)((h2 * 32 + i) * 64 + lane * 2)),
                                 "h"(*reinterpret_cast<const unsigned short *>(&hv)));
                }
            } else {
                const int m = m0w + lane;
                if (m < M) {
                    const int b = m / rpb, t = m - b * rpb;
                    __nv_bfloat16 *dst = vt + (((size_t)b * ep.heads + head) * 64 + h2 * 32) * ep.tok_pitch + t;
#pragma unroll
                    for (int i = 0; i < 32; i++)
                        dst[(size_t)i * ep.tok_pitch] = __float2bfloat16(__uint_as_float(r[i]) + hp[h2 * 32 + i]);
                }
            }
        }
        if (fast) {
            warp_sync_smem();
            const int ch = lane & 3, dsub = lane >> 2;
            uint4 x[8];
#pragma unroll
            for (int it = 0; it < 8; it++) x[it] = lds128(stg + (it * 8 + dsub) * 64 + ch * 16);
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int dd = it * 8 + dsub;
                *reinterpret_cast<uint4 *>(vt + (((size_t)b_first * ep.heads + head) * 64 + dd) * ep.tok_pitch + t0 + ch * 8) =
                    x[it];
            }
        }
    }
    warp_sync_smem();       // staging buffer and hp[] are rewritten by the next call
}

// One epilogue warp drains its half of the BN accumulator columns for its 32 rows.
template <int BN, int MODE>
__device__ __forceinline__ void epilogue_tile(const GaGemmEpilogue &ep, uint32_t stg, float *hp, uint32_t trow, int lane,
                                              int chalf, int m0w, int n0, int M, int N)
{
    if (m0w >= M) return;                                          // warp-uniform: these 32 rows are padding
    if (MODE == GA_EPI_HEADS) {
#pragma unroll 1
        for (int c = chalf * (BN / 2); c < (chalf + 1) * (BN / 2); c += 64)
            if (n0 + c < N) epilogue_head(ep, stg, hp, trow + c, lane, m0w, n0 + c, M);
    } else {
        const int c0 = chalf * (BN / 2), c1 = (chalf + 1) * (BN / 2);
        if (n0 + c0 >= N) return;                                  // warp-uniform
        // software pipeline over the 32-column chunks: the TMEM load of chunk c+1 is in flight while chunk c is
        // staged and stored
        uint32_t r[32];
        ColParams cp;
        load_col_params<MODE>(ep, lane, m0w, n0 + c0, N, cp);
        tmem_ld_32x32b_x32(trow + c0, r);
#pragma unroll 1
        for (int c = c0; c < c1; c += 32) {
            if (n0 + c >= N) break;                                // warp-uniform
            tmem_ld_wait();
            stage_rows(stg, lane, r);
            const bool more = (c + 32 < c1) && (n0 + c + 32 < N);
            ColParams cpn;
            if (more) {
                tmem_ld_32x32b_x32(trow + c + 32, r);
                load_col_params<MODE>(ep, lane, m0w, n0 + c + 32, N, cpn);
            }
            warp_sync_smem();
            epilogue32<MODE>(ep, stg, lane, m0w, n0 + c, M, N, cp);
            warp_sync_smem();
            if (more) cp = cpn;
        }
    }
}

// Persistent: grid = min(#tiles, #SMs); every role loops over the CTA's tiles.  The TMEM accumulator is double
// buffered (2*BN columns), so the epilogue of tile i overlaps the TMA/MMA main loop of tile i+1.
// CS > 1: a cluster of CS CTAs works on CS vertically adjacent tiles (same n-block).  Each CTA loads its own A
// tile and 1/CS of the shared W tile, multicast to every CTA of the cluster, so the L2 -> SM operand traffic per
// CTA drops from (128 + BN) to (128 + BN/CS) rows per k-block -- the main loop is L2-bandwidth bound otherwise.
template <int BN, int CS, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const GaGemmEpilogue ep, const int M, const int N, const int K)
{
    using Cfg = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[Cfg::kStages], empty_bar[Cfg::kStages], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = smem, *smem_b = smem + Cfg::kStages * Cfg::kABytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = (K + BK - 1) / BK;
    const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
    const int num_mg = (num_m + CS - 1) / CS;          // groups of CS m-blocks
    const int tiles = num_mg * num_n;                  // work items per CLUSTER
    const int crank = CS > 1 ? (int)cluster_ctarank() : 0;
    const int cid = blockIdx.x / CS, ncl = gridDim.x / CS;
    constexpr uint16_t kMask = (uint16_t)((1u << CS) - 1);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tma_a);
        prefetch_tmap(&tma_b);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < Cfg::kStages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], CS); }
            for (int a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], kEpiWarps); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<Cfg::kTmemCols>(&tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    if (CS > 1) cluster_sync_all();                    // peers' barriers exist before anyone multicasts into them
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    pdl_wait();                     // inputs (A, the residual stream, the gate table) come from earlier kernels
    pdl_launch_dependents();        // let the next kernel run its own prologue under our main loop

    if (warp == 0) {
        if (elect_one()) {
            int it = 0;
            for (int tile = cid; tile < tiles; tile += ncl) {
                const int m0 = ((tile % num_mg) * CS + crank) * BM, n0 = (tile / num_mg) * BN;
                for (int kb = 0; kb < nk; kb++, it++) {
                    const int s = it % Cfg::kStages;
                    const uint32_t ph = (it / Cfg::kStages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);       // all CS consumers released this stage
                    mbar_expect_tx(&full_bar[s], Cfg::kABytes + Cfg::kBBytes);
                    tma_load_2d(smem_a + s * Cfg::kABytes, &tma_a, &full_bar[s], kb * BK, m0);
                    if (CS == 1) {
                        tma_load_2d(smem_b + s * Cfg::kBBytes, &tma_b, &full_bar[s], kb * BK, n0);
                    } else {
                        constexpr int kPart = Cfg::kBBytes / CS;        // this CTA's slice of the W tile
                        tma_load_2d_mcast(smem_b + s * Cfg::kBBytes + crank * kPart, &tma_b, &full_bar[s], kb * BK,
                                          n0 + crank * (BN / CS), kMask);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
            int it = 0, lt = 0;
            for (int tile = cid; tile < tiles; tile += ncl, lt++) {
                const int a = lt & 1;
                mbar_wait(&acc_empty[a], ((lt >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t tacc = tmem + (uint32_t)(a * BN);
                for (int kb = 0; kb < nk; kb++, it++) {
                    const int s = it % Cfg::kStages;
                    mbar_wait(&full_bar[s], (it / Cfg::kStages) & 1);
                    tc_fence_after();
                    const uint64_t ad = umma_desc_k_sw128(smem_u32(smem_a + s * Cfg::kABytes));
                    const uint64_t bd = umma_desc_k_sw128(smem_u32(smem_b + s * Cfg::kBBytes));
#pragma unroll
                    for (int k = 0; k < BK / 16; k++)
                        umma_bf16_ss(tacc, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    if (CS == 1) umma_commit(&empty_bar[s]);            // frees the smem stage when these MMAs retire
                    else umma_commit_mcast(&empty_bar[s], kMask);       // ... in every CTA that multicasts into it
                }
                umma_commit(&acc_full[a]);          // accumulator of this tile complete
            }
        }
    } else {
        const int ew = warp - 2;
        const int q = warp & 3;                     // TMEM lane quarter this warp may access
        const int chalf = ew >> 2;                  // which half of the BN columns this warp drains
        const uint32_t stg = smem_u32(smem + Cfg::kRing + ew * kStageBytes);
        float *hp = reinterpret_cast<float *>(smem + Cfg::kRing + kEpiWarps * kStageBytes) + ew * kHeadParams;
        int lt = 0;
        for (int tile = cid; tile < tiles; tile += ncl, lt++) {
            const int m0 = ((tile % num_mg) * CS + crank) * BM, n0 = (tile / num_mg) * BN;
            const int a = lt & 1;
            mbar_wait(&acc_full[a], (lt >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN);
            epilogue_tile<BN, MODE>(ep, stg, hp, trow, lane, chalf, m0 + q * 32, n0, M, N);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[a]);
        }
    }
    __syncthreads();
    if (CS > 1) cluster_sync_all();                    // nobody exits while a peer can still write into its smem
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem);
    }
}

// ---------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): a cluster of two CTAs computes a 256 x BN tile.  Each CTA stages its own
// 128 rows of A and HALF of the W tile (BN/2 rows), so the operand bytes an SM has to ingest per MMA cycle halve
// for W -- the 1-CTA kernel's main loop is bound by exactly that ingest.  The leader issues one M=256 MMA per
// k-step; commits are multicast to both CTAs' barriers; both CTAs run their own TMA producer and epilogue.
// ---------------------------------------------------------------------------
template <int BN> struct PairCfg {
    static constexpr int kABytes = BM * BK * 2;                 // 16 KB
    static constexpr int kBBytes = (BN / 2) * BK * 2;           // this CTA's half of the W tile
    static constexpr int kStages = (BN >= 256) ? 5 : 7;
    static constexpr int kRing = kStages * (kABytes + kBBytes);
    static constexpr int kSmem = kRing + 1024 + 36 * 1024;
    static constexpr int kTmemCols = 2 * BN;
};

template <int BN, int MODE>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tn_pair_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                         const GaGemmEpilogue ep, const int M, const int N, const int K)
{
    using Cfg = PairCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[Cfg::kStages], empty_bar[Cfg::kStages], acc_full[2], acc_empty[2];
    __shared__ uint32_t tmem_slot;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = smem, *smem_b = smem + Cfg::kStages * Cfg::kABytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nk = (K + BK - 1) / BK;
    const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
    const int num_mg = (num_m + 1) / 2;
    const int tiles = num_mg * num_n;                   // work items per pair
    const int crank = (int)cluster_ctarank();
    const bool leader = crank == 0;
    const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tma_a);
        prefetch_tmap(&tma_b);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < Cfg::kStages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            for (int a = 0; a < 2; a++) { mbar_init(&acc_full[a], 1); mbar_init(&acc_empty[a], 2 * kEpiWarps); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc_2cta<Cfg::kTmemCols>(&tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    pdl_wait();                     // inputs come from earlier kernels; everything above overlapped their tail
    pdl_launch_dependents();

    if (warp == 0) {
        if (elect_one()) {
            int it = 0;
            for (int tile = cid; tile < tiles; tile += ncl) {
                const int m0 = ((tile % num_mg) * 2 + crank) * BM, n0 = (tile / num_mg) * BN + crank * (BN / 2);
                for (int kb = 0; kb < nk; kb++, it++) {
                    const int s = it % Cfg::kStages;
                    const uint32_t ph = (it / Cfg::kStages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    if (leader) mbar_expect_tx(&full_bar[s], 2 * (Cfg::kABytes + Cfg::kBBytes));   // both CTAs' bytes
                    tma_load_2d_2cta(smem_a + s * Cfg::kABytes, &tma_a, &full_bar[s], kb * BK, m0);
                    tma_load_2d_2cta(smem_b + s * Cfg::kBBytes, &tma_b, &full_bar[s], kb * BK, n0);
                }
            }
        }
    } else if (warp == 1) {
        if (leader && elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(2 * BM, BN);
            int it = 0, lt = 0;
            for (int tile = cid; tile < tiles; tile += ncl, lt++) {
                const int a = lt & 1;
                mbar_wait(&acc_empty[a], ((lt >> 1) & 1) ^ 1);      // both CTAs' epilogues drained this accumulator
                tc_fence_after();
                const uint32_t tacc = tmem + (uint32_t)(a * BN);
                for (int kb = 0; kb < nk; kb++, it++) {
                    const int s = it % Cfg::kStages;
                    mbar_wait(&full_bar[s], (it / Cfg::kStages) & 1);
                    tc_fence_after();
                    const uint64_t ad = umma_desc_k_sw128(smem_u32(smem_a + s * Cfg::kABytes));
                    const uint64_t bd = umma_desc_k_sw128(smem_u32(smem_b + s * Cfg::kBBytes));
#pragma unroll
                    for (int k = 0; k < BK / 16; k++)
                        umma_bf16_ss_2cta(tacc, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                    umma_commit_2cta(&empty_bar[s]);
                }
                umma_commit_2cta(&acc_full[a]);
            }
        }
    } else {
        const int ew = warp - 2;
        const int q = warp & 3;
        const int chalf = ew >> 2;
        const uint32_t stg = smem_u32(smem + Cfg::kRing + ew * kStageBytes);
        float *hp = reinterpret_cast<float *>(smem + Cfg::kRing + kEpiWarps * kStageBytes) + ew * kHeadParams;
        int lt = 0;
        for (int tile = cid; tile < tiles; tile += ncl, lt++) {
            const int m0 = ((tile % num_mg) * 2 + crank) * BM, n0 = (tile / num_mg) * BN;
            const int a = lt & 1;
            mbar_wait(&acc_full[a], (lt >> 1) & 1);
            tc_fence_after();
            const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN);
            epilogue_tile<BN, MODE>(ep, stg, hp, trow, lane, chalf, m0 + q * 32, n0, M, N);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (leader) mbar_arrive(&acc_empty[a]);
                else mbar_arrive_remote(&acc_empty[a], 0);
            }
        }
    }
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc_2cta<Cfg::kTmemCols>(tmem);
    }
}

// ---- host side -------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

}  // namespace

// 2-D bf16 row-major [rows, cols] tensor map with a {64 cols x box_rows} box, 128B swizzle.
int ga_make_tmap_bf16(CUtensorMap *map, const void *ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows)
{
    EncodeTiledFn enc = get_encode();
    if (!enc) return GA_ERR_BADARG;
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) || (ld_elems * 2) % 16) return GA_ERR_BADARG;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {ld_elems * 2};
    cuuint32_t box[2] = {64, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1000 + (int)r;
}

static int sm_count() { return ga_sm_count(); }

template <int BN, int CS, int MODE>
static int launch_gemm_mode(const CUtensorMap &ta, const CUtensorMap &tb, const GaGemmEpilogue &ep, int M, int N, int K,
                            cudaStream_t s)
{
    using Cfg = GemmCfg<BN>;
    static GaPerDevice attr_set;
    if (ga_first_use_on_device(attr_set)) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN, CS, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::kSmem);
        if (e != cudaSuccess) return (int)e;
    }
    const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
    const int items = ((num_m + CS - 1) / CS) * num_n;              // work items per cluster
    int clusters = sm_count() / CS;
    if (items < clusters) clusters = items;
    dim3 grid(clusters * CS);
    if (CS == 1)
        return (int)ga_launch_pdl(gemm_bf16_tn_kernel<BN, CS, MODE>, grid, dim3(kThreads), (size_t)Cfg::kSmem, s, ta, tb, ep,
                                  M, N, K);
    return (int)ga_launch_cluster(gemm_bf16_tn_kernel<BN, CS, MODE>, grid, dim3(kThreads), (size_t)Cfg::kSmem, s,
                                  (unsigned)CS, ta, tb, ep, M, N, K);
}

template <int BN, int CS>
static int launch_gemm(const void *A, int lda, const void *W, int ldw, const GaGemmEpilogue &ep, int M, int N, int K,
                       cudaStream_t s)
{
    CUtensorMap ta, tb;
    int rc = ga_make_tmap_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM);
    if (rc) return rc;
    rc = ga_make_tmap_bf16(&tb, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)(BN / CS));
    if (rc) return rc;
    switch (ep.mode) {
    case GA_EPI_BF16: return launch_gemm_mode<BN, CS, GA_EPI_BF16>(ta, tb, ep, M, N, K, s);
    case GA_EPI_GELU_BF16: return launch_gemm_mode<BN, CS, GA_EPI_GELU_BF16>(ta, tb, ep, M, N, K, s);
    case GA_EPI_F32: return launch_gemm_mode<BN, CS, GA_EPI_F32>(ta, tb, ep, M, N, K, s);
    case GA_EPI_RESID_GATE_F32: return launch_gemm_mode<BN, CS, GA_EPI_RESID_GATE_F32>(ta, tb, ep, M, N, K, s);
    case GA_EPI_HEADS:
        if (BN == 128 || BN == 256)
            return launch_gemm_mode<((BN == 128 || BN == 256) ? BN : 128), CS, GA_EPI_HEADS>(ta, tb, ep, M, N, K, s);
        return GA_ERR_BADARG;
    default: return GA_ERR_BADARG;
    }
}

template <int BN, int MODE>
static int launch_gemm_pair_mode(const CUtensorMap &ta, const CUtensorMap &tb, const GaGemmEpilogue &ep, int M, int N, int K,
                                 cudaStream_t s)
{
    using Cfg = PairCfg<BN>;
    static GaPerDevice attr_set;
    if (ga_first_use_on_device(attr_set)) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_pair_kernel<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::kSmem);
        if (e != cudaSuccess) return (int)e;
    }
    const int num_m = (M + BM - 1) / BM, num_n = (N + BN - 1) / BN;
    const int items = ((num_m + 1) / 2) * num_n;
    int pairs = sm_count() / 2;
    if (items < pairs) pairs = items;
    return (int)ga_launch_cluster(gemm_bf16_tn_pair_kernel<BN, MODE>, dim3(pairs * 2), dim3(kThreads), (size_t)Cfg::kSmem, s,
                                  2u, ta, tb, ep, M, N, K);
}

template <int BN>
static int launch_gemm_pair(const void *A, int lda, const void *W, int ldw, const GaGemmEpilogue &ep, int M, int N, int K,
                            cudaStream_t s)
{
    CUtensorMap ta, tb;
    int rc = ga_make_tmap_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM);
    if (rc) return rc;
    rc = ga_make_tmap_bf16(&tb, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)(BN / 2));
    if (rc) return rc;
    switch (ep.mode) {
    case GA_EPI_BF16: return launch_gemm_pair_mode<BN, GA_EPI_BF16>(ta, tb, ep, M, N, K, s);
    case GA_EPI_GELU_BF16: return launch_gemm_pair_mode<BN, GA_EPI_GELU_BF16>(ta, tb, ep, M, N, K, s);
    case GA_EPI_F32: return launch_gemm_pair_mode<BN, GA_EPI_F32>(ta, tb, ep, M, N, K, s);
    case GA_EPI_RESID_GATE_F32: return launch_gemm_pair_mode<BN, GA_EPI_RESID_GATE_F32>(ta, tb, ep, M, N, K, s);
    case GA_EPI_HEADS: return launch_gemm_pair_mode<BN, GA_EPI_HEADS>(ta, tb, ep, M, N, K, s);
    default: return GA_ERR_BADARG;
    }
}

extern "C" int ga_gemm_bf16_tn(const void *A, int lda, const void *W, int ldw, int M, int N, int K,
                               const GaGemmEpilogue *epi, int block_n, void *stream)
{
    if (!A || !W || !epi || M <= 0 || N <= 0 || K <= 0) return GA_ERR_BADARG;
    // block_n = tile width {64,128,256} + 1000 * cluster size {1 (default), 2}; 9000 + width = CTA pair
    const int cs = block_n >= 1000 ? block_n / 1000 : 1;
    const int bn = block_n % 1000;
    if (epi->mode == GA_EPI_HEADS && (N % 64 != 0 || epi->heads <= 0 || bn < 128)) return GA_ERR_BADARG;
    if (bn != 64 && bn != 128 && bn != 192 && bn != 256) return GA_ERR_BADARG;
    if (bn == 192 && (cs != 1 || epi->mode == GA_EPI_HEADS)) return GA_ERR_BADARG;        // 96 columns per epilogue warp: no whole heads
    cudaStream_t s = (cudaStream_t)stream;
    if (cs == 1) {
        if (bn == 64) return launch_gemm<64, 1>(A, lda, W, ldw, *epi, M, N, K, s);
        if (bn == 128) return launch_gemm<128, 1>(A, lda, W, ldw, *epi, M, N, K, s);
        if (bn == 192) return launch_gemm<192, 1>(A, lda, W, ldw, *epi, M, N, K, s);
        return launch_gemm<256, 1>(A, lda, W, ldw, *epi, M, N, K, s);
    }
    if (cs == 2) {
        if (bn == 128) return launch_gemm<128, 2>(A, lda, W, ldw, *epi, M, N, K, s);
        if (bn == 256) return launch_gemm<256, 2>(A, lda, W, ldw, *epi, M, N, K, s);
    }
    if (cs == 9) {          // 9xxx: CTA pair (cta_group::2), 256 x bn tile per pair
        if (bn == 128) return launch_gemm_pair<128>(A, lda, W, ldw, *epi, M, N, K, s);
        if (bn == 256) return launch_gemm_pair<256>(A, lda, W, ldw, *epi, M, N, K, s);
    }
    return GA_ERR_BADARG;
}
