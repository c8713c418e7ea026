/* libdvst.so - C ABI of the B200-native scoring path (TimeSformer forward + DINO head/loss).
 *
 * Conventions: every function returns 0 on success, DVST_ERR otherwise (message via dvst_last_error(), thread
 * local); nothing throws.  All data pointers are caller-owned DEVICE pointers unless stated otherwise; `stream` is a
 * cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream).  No call synchronises the device.
 */
#ifndef DVST_H_
#define DVST_H_

#ifdef __cplusplus
extern "C" {
#endif

#define DVST_ABI_VERSION 1
#define DVST_ERR 1

const char* dvst_last_error(void);
/* Number of kernel launches issued by the library so far (bench.py's `gpu_launches`). */
long long dvst_launch_count(void);
int dvst_abi_version(void);

/* out[M,N] = epilogue(A[M,K] x W[N,K]^T).  bf16 operands, fp32 accumulate (tcgen05 / TMEM).
 * Replaces torch.nn.Linear as used by models/timesformer.py:41-52,63-64,71,83,108,133 and
 * vision_transformer.py:254-266.
 *   split    : 0 plain; 1 = operands hold [hi | lo] bf16 halves (lo at column K), 3-term product.
 *   bias     : fp32 [N] or NULL.   residual : fp32 [M,ldr] added before the store, or NULL (may alias out).
 *   out_kind : 0 bf16, 1 fp32, 2 bf16 hi/lo pair (lo written at column + lo_off).   act : 0 none, 1 exact GELU. */
int dvst_gemm_bf16(const void* A, long long lda, const void* W, long long ldw, const float* bias, int M, int N, int K,
                   int split, const float* residual, long long ldr, void* out, long long ldo, int out_kind,
                   long long lo_off, int act, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DVST_H_ */
