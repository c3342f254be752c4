/* supir_b200 — C ABI of the B200 (sm_100a) backend for SUPIR's EDM sampling hot path.
 *
 * The reference (Fanghua-Yu/SUPIR) has no FFI: its hot path is a tree of torch.nn modules that dispatch to
 * cuDNN/cuBLAS/ATen. This header is the boundary a maintainer binds instead (ctypes stub in INTEGRATION.md): each entry
 * point names the reference call site(s) it replaces. Conventions:
 *   - plain device pointers + sizes, no torch types; the caller (PyTorch host code) owns every buffer;
 *   - activations are channels-last: images [B, H, W, C] ("NHWC"), tokens [B, L, C] — the same memory;
 *   - bf16 storage unless a name says f32; statistics / softmax / accumulation are fp32 (fp64 for GroupNorm sums);
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream and allocates nothing;
 *   - return 0 on success, negative on error; supir_last_error() returns the message (thread-local).
 */
#ifndef SUPIR_B200_H
#define SUPIR_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------------------------ */
/* library                                                                                                            */
/* ------------------------------------------------------------------------------------------------------------------ */
const char* supir_last_error(void);
int supir_version(void);
/* number of kernels launched by this library since load / since the last reset (bench.py's gpu_launches) */
long long supir_launch_count(void);
void supir_reset_launch_count(void);

/* ------------------------------------------------------------------------------------------------------------------ */
/* K2/K3/K4: Linear, conv1x1, conv3x3 on tcgen05 tensor cores (gemm.cu)                                               */
/* ------------------------------------------------------------------------------------------------------------------ */
typedef struct supir_epilogue {
    const float* bias;      /* [N] fp32 or NULL                      (nn.Linear / nn.Conv2d bias)                      */
    const float* rowvec;    /* [nbatch, rowvec_ld] fp32 or NULL: added per batch element                              */
                            /*   (ResBlock "h + emb_out[..., None, None]", openaimodel.py:343-353)                     */
    int rows_per_batch;     /* GEMM mode: rows per batch element for rowvec (conv mode uses the image index)           */
    int rowvec_ld;          /* leading dim of rowvec (0 -> N)                                                          */
    const void* residual;   /* bf16 [M, ldr] or NULL: added after the bf16 rounding of the GEMM result                 */
    long long ldr;
    int act;                /* 0 none; 1 SiLU; 2 GEGLU (accumulator columns in groups of 32 = 16 value | 16 gate,      */
                            /*   output has N/2 columns; attention.py:84-92, exact-erf GELU)                            */
    int out_f32;            /* 1: out is fp32 instead of bf16                                                          */
} supir_epilogue;

/* out[M, N] = epilogue(A[M, K] @ W[N, K]^T); A, W bf16 row-major with leading dims lda/ldw (elements, multiples of 8).
 * Replaces nn.Linear (attention.py:213-218, 87, 106, 587, 611; openaimodel.py:287-293) and 1x1 nn.Conv2d on NHWC data
 * (openaimodel.py:317; SUPIR_v0.py:48,87; model.py:124-126,164-175; autoencoder.py:297-298). */
int supir_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, void* out, long long ldc, int M, int N,
                    int K, const supir_epilogue* ep, void* stream);

/* 3x3, stride 1, zero pad 1 convolution on NHWC bf16 as an implicit GEMM (TMA performs the im2col and the padding).
 * x: [B, H, W, ldx>=Cin]; Wp: packed weights [Cout, 3, 3, Cin] (= torch weight.permute(0,2,3,1)); out: [B,H,W,ldc].
 * Replaces nn.Conv2d(k=3,p=1) in ResBlock/ZeroSFT/VAE (openaimodel.py:263,300-307; SUPIR_v0.py:79,82-83; model.py:108-117). */
int supir_conv3x3_bf16(const void* x, long long ldx, const void* Wp, void* out, long long ldc, int B, int H, int W,
                       int Cin, int Cout, const supir_epilogue* ep, void* stream);

/* debugging / tuning knob: force the N tile (64/128/256), 0 = heuristic */
int supir_set_gemm_tile_n(int bn);
/* debugging: override the UMMA shared-memory descriptor template / instruction descriptor (-1 = built-in default) */
int supir_debug_set_umma_descriptors(long long smem_desc_template, long long idesc);

#ifdef __cplusplus
}
#endif
#endif /* SUPIR_B200_H */
