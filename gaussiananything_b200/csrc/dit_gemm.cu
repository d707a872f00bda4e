// tcgen05 GEMM for the DiT denoiser: C[M,N] = A[M,K] * W[N,K]^T, bf16 operands,
// fp32 accumulation in TMEM, fused epilogues.  Replaces the cuBLAS nn.Linear +
// separate bias / GELU / gate / residual / qk-RMSNorm / permute kernels of
// /root/reference/dit/dit_models_xformers.py:765-787,
// /root/reference/vit/vision_transformer.py:215-303 and
// /root/reference/ldm/modules/attention.py:484-561.
//
// One CTA computes one 128 x BN output tile:
//   warp 0   : TMA producer  (cp.async.bulk.tensor, 128B-swizzled K-major tiles)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2-5: epilogue, one accumulator row per thread (tcgen05.ld 32x32b)
// smem ring of kStages {A 128x64, W BNx64} bf16 tiles, full/empty mbarriers.
#include "../../include/ga_b200.h"
#include "sm100_ptx.cuh"

using namespace sm100;

namespace {

constexpr int BM = 128, BK = 64;
constexpr int kThreads = 192;

template <int BN> struct GemmCfg {
    static constexpr int kStages = (BN >= 256) ? 4 : 4;
    static constexpr int kABytes = BM * BK * 2;
    static constexpr int kBBytes = BN * BK * 2;
    static constexpr int kSmem = kStages * (kABytes + kBBytes) + 1024;
};

__device__ __forceinline__ float gelu_erf(float v) { return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f)); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// epilogue for 64 consecutive columns [n, n+64) of one row (v[] = accumulator)
__device__ __forceinline__ void epilogue64(const GaGemmEpilogue &ep, int m, int n, int M, int N, float (&v)[64])
{
    if (m >= M || n >= N) return;
    if (ep.bias) {
#pragma unroll
        for (int i = 0; i < 64; i++) v[i] += (n + i < N) ? __ldg(ep.bias + n + i) : 0.f;
    }
    const bool full = (n + 64 <= N);
    switch (ep.mode) {
    case GA_EPI_BF16:
    case GA_EPI_GELU_BF16: {
        if (ep.mode == GA_EPI_GELU_BF16) {
#pragma unroll
            for (int i = 0; i < 64; i++) v[i] = gelu_erf(v[i]);
        }
        __nv_bfloat16 *dst = reinterpret_cast<__nv_bfloat16 *>(ep.out) + (size_t)m * ep.ld_out + n;
        if (full && (ep.ld_out % 8) == 0) {
            uint4 *d4 = reinterpret_cast<uint4 *>(dst);
#pragma unroll
            for (int i = 0; i < 8; i++)
                d4[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                                   pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
        } else {
            for (int i = 0; i < 64 && n + i < N; i++) dst[i] = __float2bfloat16(v[i]);
        }
        break;
    }
    case GA_EPI_F32: {
        float *dst = reinterpret_cast<float *>(ep.out) + (size_t)m * ep.ld_out + n;
        for (int i = 0; i < 64 && n + i < N; i++) dst[i] = v[i];
        break;
    }
    case GA_EPI_RESID_GATE_F32: {
        // x[m, n] += gate[b, n] * (acc + bias)
        float *dst = reinterpret_cast<float *>(ep.out) + (size_t)m * ep.ld_out + n;
        const float *g = ep.gate ? ep.gate + (size_t)(m / ep.rows_per_batch) * ep.gate_ld + n : nullptr;
        if (full && (ep.ld_out % 4) == 0) {
            float4 *d4 = reinterpret_cast<float4 *>(dst);
#pragma unroll
            for (int i = 0; i < 16; i++) {
                float4 x = d4[i];
                const float g0 = g ? __ldg(g + 4 * i) : 1.f, g1 = g ? __ldg(g + 4 * i + 1) : 1.f;
                const float g2 = g ? __ldg(g + 4 * i + 2) : 1.f, g3 = g ? __ldg(g + 4 * i + 3) : 1.f;
                x.x += g0 * v[4 * i]; x.y += g1 * v[4 * i + 1]; x.z += g2 * v[4 * i + 2]; x.w += g3 * v[4 * i + 3];
                d4[i] = x;
            }
        } else {
            for (int i = 0; i < 64 && n + i < N; i++) dst[i] += (g ? __ldg(g + i) : 1.f) * v[i];
        }
        break;
    }
    case GA_EPI_HEADS: {
        // 64 columns == one attention head of q, k or v.  Column layout "(K H D)":
        // which = n / inner, head = (n % inner) / 64   (vit/vision_transformer.py:191, 255)
        const int inner = ep.heads * 64;
        const int which = n / inner + ep.first_part;        // 0 = q, 1 = k, 2 = v
        const int head = (n % inner) / 64;
        const int b = m / ep.rows_per_batch, t = m % ep.rows_per_batch;
        const size_t bh = (size_t)b * ep.heads + head;
        if (which <= 1) {
            // per-head RMSNorm (dit/norm.py:27-40, eps 1e-5), fp32, then * weight
            const float *w = which == 0 ? ep.qn_w : ep.kn_w;
            if (w) {
                float ss = 0.f;
#pragma unroll
                for (int i = 0; i < 64; i++) ss += v[i] * v[i];
                const float r = rsqrtf(ss * (1.0f / 64.0f) + ep.eps);
#pragma unroll
                for (int i = 0; i < 64; i++) v[i] = v[i] * r * __ldg(w + i);
            }
            __nv_bfloat16 *base = reinterpret_cast<__nv_bfloat16 *>(which == 0 ? ep.q : ep.k);
            uint4 *d4 = reinterpret_cast<uint4 *>(base + (bh * ep.tok_pitch + t) * 64);
#pragma unroll
            for (int i = 0; i < 8; i++)
                d4[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                                   pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
        } else {
            // V is stored transposed per head, [B, H, 64, tok_pitch], so that P*V is a
            // K-major x K-major tcgen05 contraction (keys contiguous)
            __nv_bfloat16 *base = reinterpret_cast<__nv_bfloat16 *>(ep.vt) + bh * 64 * ep.tok_pitch + t;
#pragma unroll
            for (int i = 0; i < 64; i++) base[(size_t)i * ep.tok_pitch] = __float2bfloat16(v[i]);
        }
        break;
    }
    default: break;
    }
}

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                    const GaGemmEpilogue ep, const int M, const int N, const int K)
{
    using Cfg = GemmCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t full_bar[Cfg::kStages], empty_bar[Cfg::kStages], acc_bar;
    __shared__ uint32_t tmem_slot;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *smem_a = smem, *smem_b = smem + Cfg::kStages * Cfg::kABytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int nk = (K + BK - 1) / BK;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tma_a);
        prefetch_tmap(&tma_b);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < Cfg::kStages; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
            mbar_init(&acc_bar, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<BN>(&tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            for (int kb = 0; kb < nk; kb++) {
                const int s = kb % Cfg::kStages;
                const uint32_t ph = (kb / Cfg::kStages) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                mbar_expect_tx(&full_bar[s], Cfg::kABytes + Cfg::kBBytes);
                tma_load_2d(smem_a + s * Cfg::kABytes, &tma_a, &full_bar[s], kb * BK, m0);
                tma_load_2d(smem_b + s * Cfg::kBBytes, &tma_b, &full_bar[s], kb * BK, n0);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_bf16(BM, BN);
            for (int kb = 0; kb < nk; kb++) {
                const int s = kb % Cfg::kStages;
                const uint32_t ph = (kb / Cfg::kStages) & 1;
                mbar_wait(&full_bar[s], ph);
                tc_fence_after();
                const uint64_t ad = umma_desc_k_sw128(smem_u32(smem_a + s * Cfg::kABytes));
                const uint64_t bd = umma_desc_k_sw128(smem_u32(smem_b + s * Cfg::kBBytes));
#pragma unroll
                for (int k = 0; k < BK / 16; k++)
                    umma_bf16_ss(tmem, ad + (uint64_t)(k * 2), bd + (uint64_t)(k * 2), idesc, (kb | k) != 0);
                umma_commit(&empty_bar[s]);     // frees the smem stage when these MMAs retire
            }
            umma_commit(&acc_bar);              // accumulator complete
        }
    } else {
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const int row = q * 32 + lane;
        mbar_wait(&acc_bar, 0);
        tc_fence_after();
        const uint32_t trow = tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < BN; c += 64) {
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(trow + c, r0);
            tmem_ld_32x32b_x32(trow + c + 32, r1);
            tmem_ld_wait();
            float v[64];
#pragma unroll
            for (int i = 0; i < 32; i++) { v[i] = __uint_as_float(r0[i]); v[32 + i] = __uint_as_float(r1[i]); }
            epilogue64(ep, m0 + row, n0 + c, M, N, v);
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<BN>(tmem);
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

template <int BN>
static int launch_gemm(const CUtensorMap &ta, const CUtensorMap &tb, const GaGemmEpilogue &ep, int M, int N, int K,
                       cudaStream_t s)
{
    using Cfg = GemmCfg<BN>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tn_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             Cfg::kSmem);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN);
    gemm_bf16_tn_kernel<BN><<<grid, kThreads, Cfg::kSmem, s>>>(ta, tb, ep, M, N, K);
    return (int)cudaGetLastError();
}

extern "C" int ga_gemm_bf16_tn(const void *A, int lda, const void *W, int ldw, int M, int N, int K,
                               const GaGemmEpilogue *epi, int block_n, void *stream)
{
    if (!A || !W || !epi || M <= 0 || N <= 0 || K <= 0) return GA_ERR_BADARG;
    if (epi->mode == GA_EPI_HEADS && (N % 64 != 0 || epi->heads <= 0)) return GA_ERR_BADARG;
    if (block_n != 64 && block_n != 128 && block_n != 256) return GA_ERR_BADARG;
    CUtensorMap ta, tb;
    int rc = ga_make_tmap_bf16(&ta, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, BM);
    if (rc) return rc;
    rc = ga_make_tmap_bf16(&tb, W, (uint64_t)N, (uint64_t)K, (uint64_t)ldw, (uint32_t)block_n);
    if (rc) return rc;
    cudaStream_t s = (cudaStream_t)stream;
    if (block_n == 64) return launch_gemm<64>(ta, tb, *epi, M, N, K, s);
    if (block_n == 128) return launch_gemm<128>(ta, tb, *epi, M, N, K, s);
    return launch_gemm<256>(ta, tb, *epi, M, N, K, s);
}
