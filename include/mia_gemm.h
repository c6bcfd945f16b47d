/* C ABI of the tensor-core GEMM of libmia_scan.so (sm_100a: tcgen05.mma + TMEM accumulators + TMA tensor-map loads).
 *
 *   C[M, N] = act( A[M, K] . W[N, K]^T + bias[N] )
 *
 * It replaces the library GEMMs (cuBLAS through nn.Linear / torch.einsum, cuDNN through kernel==stride nn.Conv2d) on the
 * projections of the hot path:
 *   SS2D      in_proj / out_proj / x_proj   R2GenCSR/VMamba/classification/models/vmamba.py:751, 775, 386
 *   Mamba     in_proj / x_proj / out_proj   CXPMRG_Bench_MambaXray_VL/arm/Finetuning/mamba_simple.py:408-414, 686, 708
 *   SmallPatchEmbed conv 16/16, 4/4, 1x1    HD_Xray_Pretrain_MAE/pretrain/patch_embed.py:25-41 (one GEMM each over patches)
 *   timm Block qkv / proj / fc1 / fc2       HD_Xray_Pretrain_MAE/pretrain/models/mae.py:64-66, 82-84
 *
 * A (M x K, row pitch lda) and W (N x K, row pitch ldw: nn.Linear's weight layout) hold bf16 or fp16 with K contiguous;
 * accumulation is fp32 in tensor memory; bias is fp32 (may be NULL); C (row pitch ldc) is written in the input dtype or in
 * fp32.  Device pointers only, 16-byte aligned A and W, lda / ldw multiples of 8 elements; any M, N; K % 8 == 0 is implied
 * by the pitch rule (a K tail inside the last 64-wide block is zero-filled by TMA).  Asynchronous on `cuda_stream`.
 * Returns 0 or a negative MIA_GEMM_E* code; mia_gemm_last_error() holds the message (thread-local).
 */
#ifndef MIA_GEMM_H_
#define MIA_GEMM_H_

#ifdef __cplusplus
extern "C" {
#endif

enum { MIA_GEMM_F32 = 0, MIA_GEMM_F16 = 1, MIA_GEMM_BF16 = 2 };          /* same numbering as MIA_F32 / MIA_F16 / MIA_BF16 */
enum { MIA_ACT_NONE = 0, MIA_ACT_RELU = 1, MIA_ACT_GELU = 2, MIA_ACT_SILU = 3 };   /* GELU: exact (erf) form, nn.GELU() default */
enum { MIA_GEMM_OK = 0, MIA_GEMM_EINVAL = -1, MIA_GEMM_ECUDA = -2 };

int mia_gemm_tn(const void *A, const void *W, const float *bias, void *C, int M, int N, int K, long long lda, long long ldw,
                long long ldc, int in_dtype, int out_dtype, int act, void *cuda_stream);
/* General form: each operand may also be stored with its M / N index contiguous ("MN-major"), which is what the two
 * backward contractions of a linear layer need with NO transposed copies:
 *   a_mn_major = 0: A is [M][K] (pitch lda)      a_mn_major = 1: A is stored [K][M] (pitch lda), i.e. A^T read in place
 *   b_mn_major = 0: B is [N][K] (pitch ldb)      b_mn_major = 1: B is stored [K][N] (pitch ldb)
 *   dX[tok, in]  = dY[tok, out] . W[out, in]      -> A = dY (0),            B = W stored [K = out][N = in]   (1)
 *   dW[out, in]  = dY[tok, out]^T . X[tok, in]    -> A = dY stored [K][M] (1), B = X stored [K = tok][N = in] (1)
 * The tcgen05 instruction descriptor's a_major / b_major bits select the layout; TMA boxes are [64 k][64 mn] for it. */
int mia_gemm(const void *A, const void *B, const float *bias, void *C, int M, int N, int K, long long lda, long long ldb, long long ldc,
             int a_mn_major, int b_mn_major, int in_dtype, int out_dtype, int act, void *cuda_stream);
const char *mia_gemm_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* MIA_GEMM_H_ */
