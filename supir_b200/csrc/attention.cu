// supir_b200 — K6 of SURVEY.md §2a: softmax(Q K^T / sqrt(d)) V for head_dim 64, no mask, on tcgen05.
//
// Replaces F.scaled_dot_product_attention / xformers.memory_efficient_attention in CrossAttention
// (sgm/modules/attention.py:273-277, 357-359) for self-attention, 77-token cross-attention and ZeroCrossAttn
// (SUPIR/modules/SUPIR_v0.py:146).
//
// One CTA = one (batch, head, 128-query block); 192 threads:
//   warp 0 lane 0 : TMA producer (Q once; K/V 128-row blocks through a 3-stage ring)
//   warp 1 lane 0 : tcgen05.mma issuer:  S_j = Q K_j^T  (128x128x64)  -> TMEM S[j&1]
//                                        O_j = P_j V_j   (128x64x128)  -> TMEM O[j&1]   (not accumulated in TMEM)
//   warps 2..5    : one query row per thread. Pass 1 over S: row max; pass 2: p = exp2((s - m) * scale), row sum, P written
//                   to shared memory as the bf16 K-major 128B-swizzled A operand of the PV MMA; the running output is
//                   kept in registers: O = (O + O_{j-1}) * alpha_j, so the rescale never touches TMEM.
// QK^T of block j+1 is issued before PV of block j, so the tensor pipe works while the softmax warps are busy.
// V is consumed in its natural [kv, d] layout as an MN-major B operand (no transpose pass).
#include "common.cuh"
#include "supir_b200.h"

namespace supir {

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box);

static constexpr int ATT_BM = 128;   // queries per CTA
static constexpr int ATT_BN = 128;   // keys per block
static constexpr int ATT_D = 64;
static constexpr int ATT_STAGES = 3;
static constexpr int ATT_THREADS = 192;

struct AttnParams {
    int Lq, Lk, H;
    long long ldo;
    __nv_bfloat16* out;
    float scale_log2;
    uint32_t desc_hi;      // smem descriptor template (upper word)
    uint32_t idesc_qk;     // 128x128, A K-major, B K-major
    uint32_t idesc_pv;     // 128x64,  A K-major, B MN-major
};

struct AttnSmem {
    static constexpr int Q_BYTES = ATT_BM * ATT_D * 2;          // 16 KB
    static constexpr int K_BYTES = ATT_BN * ATT_D * 2;          // 16 KB
    static constexpr int V_BYTES = ATT_BN * ATT_D * 2;          // 16 KB
    static constexpr int P_BYTES = ATT_BM * ATT_BN * 2;         // 32 KB (two 64-column halves)
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = OFF_Q + Q_BYTES;
    static constexpr int OFF_V = OFF_K + ATT_STAGES * K_BYTES;
    static constexpr int OFF_P = OFF_V + ATT_STAGES * V_BYTES;
    static constexpr int OFF_BAR = OFF_P + 2 * P_BYTES;
    static constexpr int TOTAL = OFF_BAR + 256 + 1024;
};

__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_d64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem + AttnSmem::OFF_Q;
    uint8_t* sK = smem + AttnSmem::OFF_K;
    uint8_t* sV = smem + AttnSmem::OFF_V;
    uint8_t* sP = smem + AttnSmem::OFF_P;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnSmem::OFF_BAR);
    uint64_t* q_full = bars;                // 1
    uint64_t* kv_full = bars + 1;           // [3]
    uint64_t* kv_empty = bars + 4;          // [3]
    uint64_t* s_full = bars + 7;            // [2]
    uint64_t* s_empty = bars + 9;           // [2]
    uint64_t* p_full = bars + 11;           // [2]
    uint64_t* o_full = bars + 13;           // [2]
    uint64_t* o_empty = bars + 15;          // [2]
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qblk = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
    const int nblk = (p.Lk + ATT_BN - 1) / ATT_BN;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        mbar_init(q_full, 1);
        for (int s = 0; s < ATT_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&s_empty[i], 4);
            mbar_init(&p_full[i], 4);
            mbar_init(&o_full[i], 1);
            mbar_init(&o_empty[i], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    const uint32_t tS = tmem_base;          // S[i] at column i*128
    const uint32_t tO = tmem_base + 256;    // O[i] at column 256 + i*64

    if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer ----------------
            mbar_expect_tx(q_full, AttnSmem::Q_BYTES);
            tma_load_2d(sQ, &tmQ, q_full, head * ATT_D, batch * p.Lq + qblk * ATT_BM);
            int stage = 0;
            uint32_t phase = 0;
            for (int j = 0; j < nblk; ++j) {
                mbar_wait(&kv_empty[stage], phase ^ 1);
                mbar_expect_tx(&kv_full[stage], AttnSmem::K_BYTES + AttnSmem::V_BYTES);
                tma_load_2d(sK + stage * AttnSmem::K_BYTES, &tmK, &kv_full[stage], head * ATT_D, batch * p.Lk + j * ATT_BN);
                tma_load_2d(sV + stage * AttnSmem::V_BYTES, &tmV, &kv_full[stage], head * ATT_D, batch * p.Lk + j * ATT_BN);
                if (++stage == ATT_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ---------------- MMA issuer ----------------
            const uint64_t dtemplate = (uint64_t)p.desc_hi << 32;
            const uint64_t qdesc = dtemplate | ((smem_u32(sQ) >> 4) & 0x3FFF);
            mbar_wait(q_full, 0);
            int stage = 0;
            uint32_t kv_phase = 0;
            int pv_stage = 0;
            for (int j = 0; j <= nblk; ++j) {
                if (j < nblk) {
                    const int sb = j & 1;
                    mbar_wait(&kv_full[stage], kv_phase);
                    mbar_wait(&s_empty[sb], ((j >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint64_t kdesc = dtemplate | ((smem_u32(sK + stage * AttnSmem::K_BYTES) >> 4) & 0x3FFF);
#pragma unroll
                    for (int k = 0; k < ATT_D / 16; ++k)
                        umma_bf16(tS + sb * ATT_BN, qdesc + 2 * k, kdesc + 2 * k, p.idesc_qk, k != 0);
                    umma_commit(&s_full[sb]);
                    if (++stage == ATT_STAGES) { stage = 0; kv_phase ^= 1; }
                }
                if (j > 0) {
                    const int jj = j - 1, ob = jj & 1;
                    mbar_wait(&p_full[ob], (jj >> 1) & 1);
                    mbar_wait(&o_empty[ob], ((jj >> 1) & 1) ^ 1);
                    tc_fence_after();
                    const uint32_t pbase = smem_u32(sP + ob * AttnSmem::P_BYTES);
                    const uint32_t vbase = smem_u32(sV + pv_stage * AttnSmem::V_BYTES);
#pragma unroll
                    for (int k = 0; k < ATT_BN / 16; ++k) {
                        // A = P: K-major, two 64-column halves of 16 KB, 32 B per 16-k step inside a half
                        const uint32_t pa = pbase + (k >> 2) * (AttnSmem::P_BYTES / 2) + (k & 3) * 32;
                        // B = V: MN-major (d contiguous), 16 kv rows of 128 B per step
                        const uint32_t va = vbase + k * 16 * 128;
                        umma_bf16(tO + ob * ATT_D, dtemplate | ((pa >> 4) & 0x3FFF), dtemplate | ((va >> 4) & 0x3FFF),
                                  p.idesc_pv, k != 0);
                    }
                    umma_commit(&o_full[ob]);
                    umma_commit(&kv_empty[pv_stage]);
                    if (++pv_stage == ATT_STAGES) pv_stage = 0;
                }
            }
        }
    } else {
        // ---------------- softmax / output warps ----------------
        const int quad = warp & 3;
        const int row = quad * 32 + lane;                       // query row inside the tile
        const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
        float o_acc[ATT_D];
#pragma unroll
        for (int i = 0; i < ATT_D; ++i) o_acc[i] = 0.f;
        float m_run = -INFINITY, l_run = 0.f;
        uint8_t* p_row_base = nullptr;
        for (int j = 0; j < nblk; ++j) {
            const int sb = j & 1;
            mbar_wait(&s_full[sb], (j >> 1) & 1);
            tc_fence_after();
            const int kv_valid = min(ATT_BN, p.Lk - j * ATT_BN);
            // pass 1: row max
            float mx = -INFINITY;
#pragma unroll 1
            for (int c0 = 0; c0 < ATT_BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tS + sb * ATT_BN + lane_off + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (c0 + i < kv_valid) mx = fmaxf(mx, __uint_as_float(r[i]));
            }
            const float m_new = fmaxf(m_run, mx * p.scale_log2);
            const float alpha = exp2f(m_run - m_new);   // 0 on the first block (m_run = -inf)
            // pass 2: probabilities -> P (bf16, swizzled K-major), row sum
            float lsum = 0.f;
            p_row_base = sP + sb * AttnSmem::P_BYTES + row * 128;
#pragma unroll 1
            for (int c0 = 0; c0 < ATT_BN; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tS + sb * ATT_BN + lane_off + c0, r);
                tmem_ld_wait();
                float pv[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float e = (c0 + i < kv_valid) ? exp2f(__uint_as_float(r[i]) * p.scale_log2 - m_new) : 0.f;
                    pv[i] = e;
                    lsum += e;
                }
                uint8_t* half_base = p_row_base + (c0 >> 6) * (AttnSmem::P_BYTES / 2);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = ((c0 & 63) >> 3) + q;            // 16-byte chunk index inside the 128-byte row
                    uint4 u;
                    u.x = pack_bf16x2(pv[q * 8 + 0], pv[q * 8 + 1]);
                    u.y = pack_bf16x2(pv[q * 8 + 2], pv[q * 8 + 3]);
                    u.z = pack_bf16x2(pv[q * 8 + 4], pv[q * 8 + 5]);
                    u.w = pack_bf16x2(pv[q * 8 + 6], pv[q * 8 + 7]);
                    *reinterpret_cast<uint4*>(half_base + ((chunk ^ (row & 7)) << 4)) = u;
                }
            }
            tc_fence_before();
            fence_proxy_async_smem();   // make the generic-proxy writes of P visible to the tensor-core (async) proxy
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&s_empty[sb]);
                mbar_arrive(&p_full[sb]);
            }
            // fold in the previous block's PV product, then rescale to the new running max
            if (j > 0) {
                const int ob = (j - 1) & 1;
                mbar_wait(&o_full[ob], ((j - 1) >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c0 = 0; c0 < ATT_D; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld_32x32(tO + ob * ATT_D + lane_off + c0, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o_acc[c0 + i] += __uint_as_float(r[i]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&o_empty[ob]);
            }
#pragma unroll
            for (int i = 0; i < ATT_D; ++i) o_acc[i] *= alpha;
            l_run = l_run * alpha + lsum;
            m_run = m_new;
        }
        {
            const int ob = (nblk - 1) & 1;
            mbar_wait(&o_full[ob], ((nblk - 1) >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int c0 = 0; c0 < ATT_D; c0 += 32) {
                uint32_t r[32];
                tmem_ld_32x32(tO + ob * ATT_D + lane_off + c0, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o_acc[c0 + i] += __uint_as_float(r[i]);
            }
        }
        const int qrow = qblk * ATT_BM + row;
        if (qrow < p.Lq) {
            const float inv = 1.f / l_run;
            __nv_bfloat16* dst = p.out + ((long long)batch * p.Lq + qrow) * p.ldo + head * ATT_D;
#pragma unroll
            for (int q = 0; q < ATT_D / 8; ++q) {
                uint4 u;
                u.x = pack_bf16x2(o_acc[q * 8 + 0] * inv, o_acc[q * 8 + 1] * inv);
                u.y = pack_bf16x2(o_acc[q * 8 + 2] * inv, o_acc[q * 8 + 3] * inv);
                u.z = pack_bf16x2(o_acc[q * 8 + 4] * inv, o_acc[q * 8 + 5] * inv);
                u.w = pack_bf16x2(o_acc[q * 8 + 6] * inv, o_acc[q * 8 + 7] * inv);
                reinterpret_cast<uint4*>(dst)[q] = u;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

static long long g_att_desc_override = -1;
static long long g_att_idesc_pv_override = -1;

}  // namespace supir

using namespace supir;

extern "C" int supir_debug_set_attention_descriptors(long long smem_desc_template, long long idesc_pv) {
    g_att_desc_override = smem_desc_template;
    g_att_idesc_pv_override = idesc_pv;
    return SUPIR_OK;
}

extern "C" int supir_attention_bf16(const void* q, long long ldq, const void* k, long long ldk, const void* v,
                                    long long ldv, void* out, long long ldo, int B, int H, int Lq, int Lk, int head_dim,
                                    float scale, void* stream) {
    SUPIR_REQUIRE(q && k && v && out, "supir_attention_bf16: null pointer");
    SUPIR_REQUIRE(head_dim == 64, "supir_attention_bf16: head_dim %d unsupported (64 only)", head_dim);
    SUPIR_REQUIRE(B > 0 && H > 0 && Lq > 0 && Lk > 0, "supir_attention_bf16: bad shape");
    SUPIR_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "supir_attention_bf16: leading dims must be multiples of 8");
    static bool attr_set = false;
    if (!attr_set) {
        SUPIR_CHECK_CUDA(cudaFuncSetAttribute(attention_d64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::TOTAL));
        attr_set = true;
    }
    CUtensorMap tmQ, tmK, tmV;
    const uint32_t box[2] = {ATT_D, ATT_BM};
    int rc;
    {
        const uint64_t dims[2] = {(uint64_t)H * ATT_D, (uint64_t)B * Lq};
        const uint64_t str[1] = {(uint64_t)ldq};
        if ((rc = make_tmap_bf16(&tmQ, q, 2, dims, str, box))) return rc;
    }
    {
        const uint64_t dims[2] = {(uint64_t)H * ATT_D, (uint64_t)B * Lk};
        const uint64_t str[1] = {(uint64_t)ldk};
        if ((rc = make_tmap_bf16(&tmK, k, 2, dims, str, box))) return rc;
        const uint64_t strv[1] = {(uint64_t)ldv};
        if ((rc = make_tmap_bf16(&tmV, v, 2, dims, strv, box))) return rc;
    }
    AttnParams p{};
    p.Lq = Lq; p.Lk = Lk; p.H = H;
    p.ldo = ldo;
    p.out = reinterpret_cast<__nv_bfloat16*>(out);
    p.scale_log2 = scale * 1.4426950408889634f;
    const uint64_t dt = g_att_desc_override >= 0 ? (uint64_t)g_att_desc_override
                                                 : (((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61));
    p.desc_hi = (uint32_t)(dt >> 32);
    p.idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN, 0, 0);
    p.idesc_pv = g_att_idesc_pv_override >= 0 ? (uint32_t)g_att_idesc_pv_override : umma_idesc_bf16(ATT_BM, ATT_D, 0, 1);
    dim3 grid((Lq + ATT_BM - 1) / ATT_BM, H, B);
    attention_d64_kernel<<<grid, ATT_THREADS, AttnSmem::TOTAL, reinterpret_cast<cudaStream_t>(stream)>>>(tmQ, tmK, tmV, p);
    count_launch();
    SUPIR_CHECK_CUDA(cudaGetLastError());
    return SUPIR_OK;
}
