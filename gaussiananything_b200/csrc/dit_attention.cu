// tcgen05 flash attention (head_dim 64, no mask, no dropout) for the DiT
// self- and cross-attention: softmax(q k^T / 8) v.  Replaces
// xformers.ops.memory_efficient_attention at
// /root/reference/vit/vision_transformer.py:297 and
// /root/reference/ldm/modules/attention.py:538-546.
//
// Inputs are the head-split tensors the QKV GEMM epilogue writes:
//   Q  [B*H, tok_pitch_q, 64]   K [B*H, tok_pitch_k, 64]   Vt [B*H, 64, tok_pitch_k]   (bf16)
// Output O [B, Nq, H*64] bf16 (token-major: the A operand of the out-projection GEMM).
//
// One CTA = 128 queries of one (batch, head); keys in blocks of 128:
//   warp 0   : TMA producer (Q once, K/V double buffered)
//   warp 1   : single-thread tcgen05.mma issuer: S = Q K^T (128x128x64) into TMEM,
//              then O_blk = P V (128x64x128) with P staged in 128B-swizzled smem
//   warps 2-5: softmax, one query row per thread: S from TMEM -> registers,
//              online max/sum in the exp2 domain, P (bf16) -> smem, O_blk from
//              TMEM accumulated in registers with the running rescale.
// S and O_blk are double buffered in TMEM (2x128 + 2x64 columns) so the next
// block's QK^T overlaps this block's softmax.
#include "../../include/ga_b200.h"
#include "sm100_ptx.cuh"

using namespace sm100;

int ga_make_tmap_bf16(CUtensorMap *map, const void *ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows);

namespace {

constexpr int AQ = 128, AK = 128, HD = 64;
constexpr int kQBytes = AQ * HD * 2;            // 16 KB
constexpr int kKBytes = AK * HD * 2;            // 16 KB
constexpr int kVBytes = HD * AK * 2;            // 16 KB (two 64-key sub-tiles of 8 KB)
constexpr int kPBytes = AQ * AK * 2;            // 32 KB (two 64-key sub-tiles of 16 KB)
constexpr int kSmemAttn = kQBytes + 2 * kKBytes + 2 * kVBytes + 2 * kPBytes + 1024;
constexpr int kAttnThreads = 192;

__device__ __forceinline__ uint32_t pack2(float a, float b)
{
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                const __grid_constant__ CUtensorMap tma_vt, __nv_bfloat16 *__restrict__ out,
                const int Nq, const int Nk, const int pitch_q, const int pitch_k, const int heads,
                const float scale_log2)
{
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint64_t q_full, k_full[2], k_empty[2], v_full[2], v_empty[2], s_full[2], p_full[2], o_full[2];
    __shared__ uint32_t tmem_slot;
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t *sQ = smem;
    uint8_t *sK = sQ + kQBytes;
    uint8_t *sV = sK + 2 * kKBytes;
    uint8_t *sP = sV + 2 * kVBytes;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bh = blockIdx.y;
    const int q0 = blockIdx.x * AQ;
    const int nb = (Nk + AK - 1) / AK;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tma_q); prefetch_tmap(&tma_k); prefetch_tmap(&tma_vt);
    }
    if (warp == 1) {
        if (lane == 0) {
            mbar_init(&q_full, 1);
            for (int s = 0; s < 2; s++) {
                mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
                mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
                mbar_init(&s_full[s], 1); mbar_init(&o_full[s], 1);
                mbar_init(&p_full[s], 128);
            }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<512>(&tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_slot;
    const uint32_t tS[2] = {tmem, tmem + 128}, tO[2] = {tmem + 256, tmem + 320};

    if (warp == 0) {
        if (elect_one()) {
            mbar_expect_tx(&q_full, kQBytes);
            tma_load_2d(sQ, &tma_q, &q_full, 0, bh * pitch_q + q0);
            for (int j = 0; j < nb; j++) {
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                mbar_wait(&k_empty[s], ph ^ 1);
                mbar_expect_tx(&k_full[s], kKBytes);
                tma_load_2d(sK + s * kKBytes, &tma_k, &k_full[s], 0, bh * pitch_k + j * AK);
                mbar_wait(&v_empty[s], ph ^ 1);
                mbar_expect_tx(&v_full[s], kVBytes);
                tma_load_2d(sV + s * kVBytes, &tma_vt, &v_full[s], j * AK, bh * HD);
                tma_load_2d(sV + s * kVBytes + kVBytes / 2, &tma_vt, &v_full[s], j * AK + 64, bh * HD);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc_s = umma_idesc_bf16(AQ, AK);
            constexpr uint32_t idesc_o = umma_idesc_bf16(AQ, HD);
            const uint64_t qd = umma_desc_k_sw128(smem_u32(sQ));
            mbar_wait(&q_full, 0);
            auto issue_s = [&](int j) {
                const int s = j & 1;
                mbar_wait(&k_full[s], (j >> 1) & 1);
                tc_fence_after();
                const uint64_t kd = umma_desc_k_sw128(smem_u32(sK + s * kKBytes));
#pragma unroll
                for (int k = 0; k < HD / 16; k++)
                    umma_bf16_ss(tS[s], qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), idesc_s, k != 0);
                umma_commit(&k_empty[s]);
                umma_commit(&s_full[s]);
            };
            issue_s(0);
            for (int j = 0; j < nb; j++) {
                const int s = j & 1;
                const uint32_t ph = (j >> 1) & 1;
                if (j + 1 < nb) issue_s(j + 1);
                mbar_wait(&p_full[s], ph);
                mbar_wait(&v_full[s], ph);
                tc_fence_after();
                const uint64_t pd = umma_desc_k_sw128(smem_u32(sP + s * kPBytes));
                const uint64_t vd = umma_desc_k_sw128(smem_u32(sV + s * kVBytes));
#pragma unroll
                for (int kk = 0; kk < AK / 16; kk++) {
                    const uint64_t sub_p = (uint64_t)((kk >> 2) * ((kPBytes / 2) >> 4));
                    const uint64_t sub_v = (uint64_t)((kk >> 2) * ((kVBytes / 2) >> 4));
                    umma_bf16_ss(tO[s], pd + sub_p + (uint64_t)(2 * (kk & 3)), vd + sub_v + (uint64_t)(2 * (kk & 3)),
                                 idesc_o, kk != 0);
                }
                umma_commit(&v_empty[s]);
                umma_commit(&o_full[s]);
            }
        }
    } else {
        const int qd4 = warp & 3;
        const int row = qd4 * 32 + lane;
        const uint32_t lane_off = (uint32_t)(qd4 * 32) << 16;
        float m_run = -INFINITY, l_run = 0.f, corr = 0.f;
        float o_acc[HD];
#pragma unroll
        for (int i = 0; i < HD; i++) o_acc[i] = 0.f;

        auto accumulate_o = [&](int j, float c) {
            const int s = j & 1;
            mbar_wait(&o_full[s], (j >> 1) & 1);
            tc_fence_after();
            uint32_t r0[32], r1[32];
            tmem_ld_32x32b_x32(tO[s] + lane_off, r0);
            tmem_ld_32x32b_x32(tO[s] + lane_off + 32, r1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; i++) {
                o_acc[i] = o_acc[i] * c + __uint_as_float(r0[i]);
                o_acc[32 + i] = o_acc[32 + i] * c + __uint_as_float(r1[i]);
            }
        };

        for (int j = 0; j < nb; j++) {
            const int s = j & 1;
            const uint32_t ph = (j >> 1) & 1;
            mbar_wait(&s_full[s], ph);
            tc_fence_after();
            float sv[AK];
#pragma unroll
            for (int c = 0; c < AK; c += 32) {
                uint32_t r[32];
                tmem_ld_32x32b_x32(tS[s] + lane_off + c, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; i++) sv[c + i] = __uint_as_float(r[i]);
            }
            const int kbase = j * AK;
            float mx = -INFINITY;
            if (kbase + AK <= Nk) {
#pragma unroll
                for (int i = 0; i < AK; i++) mx = fmaxf(mx, sv[i]);
            } else {
#pragma unroll
                for (int i = 0; i < AK; i++) {
                    if (kbase + i >= Nk) sv[i] = -INFINITY;
                    mx = fmaxf(mx, sv[i]);
                }
            }
            const float corr_prev = corr;
            const float m_new = fmaxf(m_run, mx * scale_log2);
            corr = exp2f(m_run - m_new);
            m_run = m_new;
            float lsum = 0.f;
            uint8_t *prow = sP + s * kPBytes + (row >> 3) * 1024 + (row & 7) * 128;
#pragma unroll
            for (int c8 = 0; c8 < AK / 8; c8++) {
                float p[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    p[i] = exp2f(sv[c8 * 8 + i] * scale_log2 - m_new);
                    lsum += p[i];
                }
                const int sub = c8 >> 3, q16 = c8 & 7;
                uint4 *dst = reinterpret_cast<uint4 *>(prow + sub * (kPBytes / 2) + ((q16 ^ (row & 7)) << 4));
                *dst = make_uint4(pack2(p[0], p[1]), pack2(p[2], p[3]), pack2(p[4], p[5]), pack2(p[6], p[7]));
            }
            l_run = l_run * corr + lsum;
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&p_full[s]);
            if (j > 0) accumulate_o(j - 1, corr_prev);
        }
        accumulate_o(nb - 1, corr);
        const int q = q0 + row;
        if (q < Nq) {
            const float inv = 1.0f / l_run;
            const int b = bh / heads, h = bh % heads;
            uint4 *dst = reinterpret_cast<uint4 *>(out + ((size_t)b * Nq + q) * (size_t)(heads * HD) + h * HD);
#pragma unroll
            for (int i = 0; i < 8; i++)
                dst[i] = make_uint4(pack2(o_acc[8 * i] * inv, o_acc[8 * i + 1] * inv),
                                    pack2(o_acc[8 * i + 2] * inv, o_acc[8 * i + 3] * inv),
                                    pack2(o_acc[8 * i + 4] * inv, o_acc[8 * i + 5] * inv),
                                    pack2(o_acc[8 * i + 6] * inv, o_acc[8 * i + 7] * inv));
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<512>(tmem);
    }
}

}  // namespace

extern "C" int ga_attention_bf16(const void *Q, const void *K, const void *Vt, void *out, int batch, int heads,
                                 int Nq, int Nk, int pitch_q, int pitch_k, float softmax_scale, void *stream)
{
    if (!Q || !K || !Vt || !out || batch <= 0 || heads <= 0 || Nq <= 0 || Nk <= 0) return GA_ERR_BADARG;
    if (pitch_q < Nq || pitch_k < Nk || pitch_k % 128 != 0) return GA_ERR_BADARG;
    const uint64_t BH = (uint64_t)batch * heads;
    CUtensorMap tq, tk, tv;
    int rc = ga_make_tmap_bf16(&tq, Q, BH * pitch_q, HD, HD, AQ);
    if (rc) return rc;
    rc = ga_make_tmap_bf16(&tk, K, BH * pitch_k, HD, HD, AK);
    if (rc) return rc;
    rc = ga_make_tmap_bf16(&tv, Vt, BH * HD, (uint64_t)pitch_k, (uint64_t)pitch_k, HD);
    if (rc) return rc;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemAttn);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid((Nq + AQ - 1) / AQ, (unsigned)BH);
    const float scale_log2 = softmax_scale * 1.4426950408889634f;
    attn_fwd_kernel<<<grid, kAttnThreads, kSmemAttn, (cudaStream_t)stream>>>(
        tq, tk, tv, reinterpret_cast<__nv_bfloat16 *>(out), Nq, Nk, pitch_q, pitch_k, heads, scale_log2);
    return (int)cudaGetLastError();
}
