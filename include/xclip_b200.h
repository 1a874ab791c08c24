/*
 * xclip_b200.h - C-ABI of the B200-native CLIP training hot path.
 *
 * This is the drop-in boundary: plain `extern "C"` entry points,
 * raw device pointers + explicit shapes, a cudaStream_t, int return codes.  The
 * CALLER (PyTorch, through x_clip_b200/_lib.py) owns every buffer; the library
 * never allocates persistent device memory, never synchronises the host, never
 * throws and never exits.  All pointers are device pointers unless stated.
 *
 * Each entry point names the reference call site it replaces
 * (lucidrains/x-clip v0.14.4, paths relative to /root/reference).
 *
 * Conventions
 *   - "bf16" buffers hold __nv_bfloat16, "f32" buffers hold float.
 *   - Row-major everywhere; `ld*` are leading dimensions in ELEMENTS.
 *   - return 0 on success; otherwise an XCLIP_ERR_* code and xclip_last_error()
 *     returns a thread-local message.
 *   - Work is enqueued on `stream` and is asynchronous w.r.t. the host.
 */
#ifndef XCLIP_B200_H_
#define XCLIP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XCLIP_OK 0
#define XCLIP_ERR_INVALID 1      /* bad argument / unsupported shape */
#define XCLIP_ERR_CUDA 2         /* CUDA runtime / driver error */
#define XCLIP_ERR_UNSUPPORTED 3  /* device is not sm_100 */

#define XCLIP_MAJOR_K 0  /* operand stored with the contraction index contiguous */
#define XCLIP_MAJOR_MN 1 /* operand stored transposed (its M / N index contiguous) */

typedef void* xclip_stream_t; /* cudaStream_t */

/* ---- library ---------------------------------------------------------- */
int xclip_abi_version(void);
const char* xclip_last_error(void);
/* Binds to the current CUDA device, checks it is sm_100, resolves the driver
 * entry point used for TMA descriptors.  Idempotent. */
int xclip_init(void);
/* number of kernels this library has launched since the last reset (host counter) */
long long xclip_launch_count(void);
void xclip_launch_count_reset(void);

/* ---- dense contraction (tcgen05) --------------------------------------
 * C[M,N] (+)= alpha * A * B^T (+ bias[N]) (+ residual[row % res_row_mod or row, N])
 * Replaces nn.Linear fwd + its autograd dgrad/wgrad:
 *   x_clip/x_clip.py:191,195 (FeedForward), :209,:210 (Attention to_qkv/to_out),
 *   :358 (patch embedding, with bias), :368 (to_cls_tokens), :556,:570 (latent proj).
 * a_major = K : A is [M,K] (lda >= K).   a_major = MN : A is stored as [K,M] (lda >= M).
 * b_major = K : B is [N,K] (ldb >= K) - the nn.Linear [out,in] layout.
 * b_major = MN: B is stored as [K,N] (ldb >= N).
 * c_dtype 0 = bf16, 1 = f32.  accumulate != 0 (f32 only): C += result (atomic; enables
 * split-K over `K` when the output has too few tiles to fill the GPU - the wgrad case).
 * Requirements: N % 8 == 0, lda/ldb/ldr % 8 == 0, ldc % 8 == 0 (bf16) or % 4 (f32),
 * 16-byte aligned base pointers.  bias is f32, residual is bf16.
 */
int xclip_gemm_bf16(const void* a, int64_t lda, int a_major, const void* b, int64_t ldb,
                    int b_major, void* c, int64_t ldc, int c_dtype, int M, int N, int K,
                    float alpha, const float* bias, const void* residual, int64_t ldr,
                    int res_row_mod, int accumulate, xclip_stream_t stream);

#ifdef __cplusplus
}
#endif

#endif /* XCLIP_B200_H_ */
