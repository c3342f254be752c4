// supir_b200 — K6 of SURVEY.md §2a: softmax(Q K^T / sqrt(d)) V for head_dim 64, no mask, on tcgen05.
//
// Replaces F.scaled_dot_product_attention / xformers.memory_efficient_attention in CrossAttention
// (sgm/modules/attention.py:273-277, 357-359) for self-attention, 77-token cross-attention and ZeroCrossAttn
// (SUPIR/modules/SUPIR_v0.py:146).
//
// One CTA = one (batch, head) and 256 queries = two 128-row tiles A and B; 384 threads (3 warpgroups; the softmax
// warpgroups take the registers the control warpgroup gives up via setmaxnreg, so a whole S row lives in registers):
//   warp 0 lane 0 : TMA producer (Q_A, Q_B once; K/V 128-row blocks through a 3-stage ring, shared by both tiles)
//   warp 1 lane 0 : tcgen05.mma issuer. Per tile X and key block j:
//                     S_X = Q_X K_j^T   (128x128x64, TMEM, overwritten every block)
//                     O_X += P_X V_j    (128x64x128, ACCUMULATED in TMEM across blocks)
//                   with P_X read straight from TENSOR MEMORY (A-in-TMEM form of tcgen05.mma: P never touches shared
//                   memory, whose bandwidth the QK/PV operand reads already saturate) and V consumed in its natural
//                   [kv, d] layout as an MN-major B operand.
//   warps 4..7    : softmax group of tile A, warps 8..11: tile B — one query row per thread. The whole S row (128 fp32) is
//                   read from TMEM once into registers; p = ex2((s - m) * scale) is packed to bf16 pairs and stored back to
//                   TMEM (tcgen05.st, 64 columns per tile) as the A operand of the PV MMA.
// Online softmax without touching O every block: the exponent reference m is only moved when the running row maximum
// exceeds it by more than 8 (p <= 2^8 stays exact enough in bf16/fp32); only then the warp rescales its 32 rows of O in
// TMEM (tcgen05.ld / tcgen05.st) and its row sum. Two independent tiles keep the tensor pipe and the MUFU pipe busy
// while the other tile waits on a dependency; the ex2 throughput (16/clk/SM) is this kernel's roofline, not the MMA.
#include "common.cuh"
#include "supir_b200.h"

namespace supir {

int make_tmap_bf16(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_elems,
                   const uint32_t* box);

static constexpr int ATT_BM = 128;   // queries per tile (two tiles per CTA)
static constexpr int ATT_BN = 128;   // keys per block
static constexpr int ATT_D = 64;
static constexpr float ATT_RESCALE_THRESHOLD = 8.0f;   // log2 units

struct AttnParams {
    int Lq, Lk, H;
    long long ldo;
    __nv_bfloat16* out;
    float scale_log2;
    uint32_t desc_hi;      // smem descriptor template (upper word)
    uint32_t idesc_qk;     // 128x128, A K-major, B K-major
    uint32_t idesc_pv;     // 128x64,  A K-major, B MN-major
};

// TILES = 128-query tiles per CTA, STAGES = K/V ring depth. <2, 4>: one CTA per SM, two tiles ping-pong (self-attention).
// <1, 1>: single-block keys (77-token cross-attention): a light CTA (48 KB of shared memory, 256 TMEM columns, 256
// threads) so that two or three share an SM and one CTA's load / store latency hides behind another's math.
template <int TILES, int STAGES>
struct AttnSmem {
    static constexpr int Q_BYTES = ATT_BM * ATT_D * 2;          // 16 KB per tile
    static constexpr int K_BYTES = ATT_BN * ATT_D * 2;          // 16 KB
    static constexpr int V_BYTES = ATT_BN * ATT_D * 2;          // 16 KB
    static constexpr int OFF_Q = 0;
    static constexpr int OFF_K = OFF_Q + TILES * Q_BYTES;
    static constexpr int OFF_V = OFF_K + STAGES * K_BYTES;
    static constexpr int OFF_BAR = OFF_V + STAGES * V_BYTES;
    static constexpr int TOTAL = OFF_BAR + 256 + 1024;
};

template <int TILES, int STAGES>
__global__ void __launch_bounds__(128 + 128 * TILES, TILES == 1 ? 2 : 1)
attention_d64_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                     const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
    using AttnSmem = supir::AttnSmem<TILES, STAGES>;
    constexpr int ATT_STAGES = STAGES;
    constexpr uint32_t TMEM_COLS = TILES == 2 ? 512 : 256;
    constexpr uint32_t COL_O = TILES * ATT_BN, COL_P = TILES * ATT_BN + TILES * ATT_D;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem + AttnSmem::OFF_Q;
    uint8_t* sK = smem + AttnSmem::OFF_K;
    uint8_t* sV = smem + AttnSmem::OFF_V;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnSmem::OFF_BAR);
    uint64_t* q_full = bars;                // 1
    uint64_t* kv_full = bars + 1;                       // [ATT_STAGES]
    uint64_t* kv_empty = kv_full + ATT_STAGES;          // [ATT_STAGES]
    uint64_t* s_full = kv_empty + ATT_STAGES;           // [2] per tile
    uint64_t* p_full = s_full + 2;                      // [2] per tile
    uint64_t* pv_done = p_full + 2;                     // [2] per tile
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(pv_done + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qblk = blockIdx.x, head = blockIdx.y, batch = blockIdx.z;
    const int nblk = (p.Lk + ATT_BN - 1) / ATT_BN;
    const int q0 = qblk * TILES * ATT_BM;                           // first query row of this CTA (within the batch element)
    const bool tileB_valid = TILES == 2 && (q0 + ATT_BM) < p.Lq;  // CTA-uniform

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK);
        tma_prefetch_desc(&tmV);
        mbar_init(q_full, 1);
        for (int s = 0; s < ATT_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 4);
            mbar_init(&pv_done[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        tmem_alloc(tmem_ptr, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // TMEM columns: S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384) P_A [384,448) P_B [448,512) (bf16 pairs)

    if (warp < 4) {
      // register budgets: <2,*> 384 threads x 168 -> 56 / 224;  <1,*> 256 threads x 128 (two CTAs per SM) -> 56 / 200
      asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
      if (warp == 0) {
        if (lane == 0) {
            // ---------------- TMA producer ----------------
            mbar_expect_tx(q_full, (tileB_valid ? 2 : 1) * AttnSmem::Q_BYTES);
            tma_load_2d(sQ, &tmQ, q_full, head * ATT_D, batch * p.Lq + q0);
            if (tileB_valid) tma_load_2d(sQ + AttnSmem::Q_BYTES, &tmQ, q_full, head * ATT_D, batch * p.Lq + q0 + ATT_BM);
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
            const int ntile = tileB_valid ? 2 : 1;
            auto issue_qk = [&](int x, int stage) {
                const uint64_t qdesc = dtemplate | ((smem_u32(sQ + x * AttnSmem::Q_BYTES) >> 4) & 0x3FFF);
                const uint64_t kdesc = dtemplate | ((smem_u32(sK + stage * AttnSmem::K_BYTES) >> 4) & 0x3FFF);
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k)
                    umma_bf16(tmem_base + x * ATT_BN, qdesc + 2 * k, kdesc + 2 * k, p.idesc_qk, k != 0);
                umma_commit(&s_full[x]);
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            for (int x = 0; x < ntile; ++x) issue_qk(x, 0);
            int stage = 0;                 // stage holding block j
            uint32_t kv_phase = 0;
            for (int j = 0; j < nblk; ++j) {
                int nstage = stage + 1;
                uint32_t nphase = kv_phase;
                if (nstage == ATT_STAGES) { nstage = 0; nphase ^= 1; }
                if (j + 1 < nblk) {
                    mbar_wait(&kv_full[nstage], nphase);
                }
                for (int x = 0; x < ntile; ++x) {
                    mbar_wait(&p_full[x], j & 1);       // P_X(j) written, S_X(j) consumed
                    tc_fence_after();
                    if (j + 1 < nblk) issue_qk(x, nstage);
                    const uint32_t vbase = smem_u32(sV + stage * AttnSmem::V_BYTES);
#pragma unroll
                    for (int k = 0; k < ATT_BN / 16; ++k) {
                        // A = P in TMEM: 16 k-elements = 8 columns per step; B = V: MN-major (d contiguous), 16 kv rows of 128 B
                        const uint32_t va = vbase + k * 16 * 128;
                        umma_bf16_ts(tmem_base + COL_O + x * ATT_D, tmem_base + COL_P + x * 64 + k * 8,
                                     dtemplate | ((va >> 4) & 0x3FFF), p.idesc_pv, (j | k) != 0);
                    }
                    umma_commit(&pv_done[x]);
                }
                umma_commit(&kv_empty[stage]);
                stage = nstage;
                kv_phase = nphase;
            }
        }
      }
    } else {
        if (TILES == 2) asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        // ---------------- softmax / output warps ----------------
        const int x = (warp - 4) >> 2;                          // tile 0 (A) or 1 (B)
        if (x == 0 || tileB_valid) {
            const int quad = warp & 3;
            const int row = quad * 32 + lane;                   // query row inside the tile
            const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
            const uint32_t tS = tmem_base + x * ATT_BN + lane_off;
            const uint32_t tO = tmem_base + COL_O + x * ATT_D + lane_off;
            const uint32_t tP = tmem_base + COL_P + x * 64 + lane_off;
            float m_ref = 0.f, l_run = 0.f;
            for (int j = 0; j < nblk; ++j) {
                mbar_wait(&s_full[x], j & 1);
                tc_fence_after();
                uint32_t s[128];
                tmem_ld_32x32_at<0>(tS + 0, s);
                tmem_ld_32x32_at<32>(tS + 32, s);
                tmem_ld_32x32_at<64>(tS + 64, s);
                tmem_ld_32x32_at<96>(tS + 96, s);
                tmem_ld_wait();
                const int kv_valid = min(ATT_BN, p.Lk - j * ATT_BN);
                if (kv_valid != ATT_BN) {
#pragma unroll
                    for (int i = 0; i < 128; ++i)
                        if (i >= kv_valid) s[i] = 0xff800000u;   // -inf
                }
                // row maximum with 8 independent chains (a single chain would serialise 128 dependent FMNMX)
                float mxa[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) mxa[t] = __uint_as_float(s[t]);
#pragma unroll
                for (int i = 8; i < 128; i += 8) {
#pragma unroll
                    for (int t = 0; t < 8; ++t) mxa[t] = fmaxf(mxa[t], __uint_as_float(s[i + t]));
                }
                float mx = fmaxf(fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3])),
                                 fmaxf(fmaxf(mxa[4], mxa[5]), fmaxf(mxa[6], mxa[7])));
                mx *= p.scale_log2;
                // previous PV of this tile must have finished: it reads P (about to be overwritten) and updates O
                if (j > 0) {
                    mbar_wait(&pv_done[x], (j - 1) & 1);
                    tc_fence_after();
                }
                if (j == 0) {
                    m_ref = mx;
                } else {
                    const bool need = mx > m_ref + ATT_RESCALE_THRESHOLD;
                    if (__any_sync(0xffffffffu, need)) {
                        const float alpha = need ? ex2_approx(m_ref - mx) : 1.0f;
                        if (need) m_ref = mx;
                        l_run *= alpha;
#pragma unroll
                        for (int c0 = 0; c0 < ATT_D; c0 += 32) {
                            uint32_t o[32];
                            tmem_ld_32x32(tO + c0, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                            tmem_st_32x32(tO + c0, o);
                        }
                        tmem_st_wait();
                    }
                }
                float ls[8];
#pragma unroll
                for (int t = 0; t < 8; ++t) ls[t] = 0.f;
#pragma unroll
                for (int c0 = 0; c0 < ATT_BN; c0 += 64) {
                    uint32_t pk[32];                                     // 64 probabilities as bf16 pairs
#pragma unroll
                    for (int i = 0; i < 64; i += 8) {
                        float e[8];
#pragma unroll
                        for (int t = 0; t < 8; ++t) {
                            e[t] = ex2_approx(fmaf(__uint_as_float(s[c0 + i + t]), p.scale_log2, -m_ref));
                            ls[t] += e[t];
                        }
#pragma unroll
                        for (int t = 0; t < 4; ++t) pk[(i >> 1) + t] = pack_bf16x2(e[2 * t], e[2 * t + 1]);
                    }
                    tmem_st_32x32(tP + (c0 >> 1), pk);
                }
                tmem_st_wait();
                l_run += ((ls[0] + ls[1]) + (ls[2] + ls[3])) + ((ls[4] + ls[5]) + (ls[6] + ls[7]));
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&p_full[x]);
            }
            // epilogue: O / l
            mbar_wait(&pv_done[x], (nblk - 1) & 1);
            tc_fence_after();
            const int qrow = q0 + x * ATT_BM + row;
            const float inv = 1.f / l_run;
            __nv_bfloat16* dst = p.out + ((long long)batch * p.Lq + qrow) * p.ldo + head * ATT_D;
#pragma unroll
            for (int c0 = 0; c0 < ATT_D; c0 += 32) {
                uint32_t o[32];
                tmem_ld_32x32(tO + c0, o);
                tmem_ld_wait();
                if (qrow < p.Lq) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        uint4 u;
                        u.x = pack_bf16x2(__uint_as_float(o[q * 8 + 0]) * inv, __uint_as_float(o[q * 8 + 1]) * inv);
                        u.y = pack_bf16x2(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv);
                        u.z = pack_bf16x2(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv);
                        u.w = pack_bf16x2(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv);
                        reinterpret_cast<uint4*>(dst + c0)[q] = u;
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

static long long g_att_desc_override = -1;
static long long g_att_idesc_pv_override = -1;

template <int TILES, int STAGES>
static int launch_attention(const CUtensorMap& tmQ, const CUtensorMap& tmK, const CUtensorMap& tmV, const AttnParams& p, int B,
                            cudaStream_t st) {
    using S = AttnSmem<TILES, STAGES>;
    static bool attr_set = false;
    if (!attr_set) {
        SUPIR_CHECK_CUDA(cudaFuncSetAttribute(attention_d64_kernel<TILES, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL));
        attr_set = true;
    }
    dim3 grid((p.Lq + TILES * ATT_BM - 1) / (TILES * ATT_BM), p.H, B);
    attention_d64_kernel<TILES, STAGES><<<grid, 128 + 128 * TILES, S::TOTAL, st>>>(tmQ, tmK, tmV, p);
    count_launch();
    SUPIR_CHECK_CUDA(cudaGetLastError());
    return SUPIR_OK;
}

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
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // one key block (77-token cross-attention): light one-tile CTAs, two per SM. One-tile CTAs measured no better than the
    // two-tile kernel for multi-block self-attention (745 vs 762 TFLOP/s at 4096 tokens), so that keeps the ping-pong.
    if (Lk <= ATT_BN) return launch_attention<1, 1>(tmQ, tmK, tmV, p, B, st);
    return launch_attention<2, 4>(tmQ, tmK, tmV, p, B, st);
}
