/* fsb200 — C ABI of libfsb200.so, the B200 (sm_100a) backend for the Fengshen data-parallel pretraining step.
 *
 * This is the drop-in boundary. The reference reaches its native code through pybind11
 * torch-extension modules that take torch::Tensor (fengshen/models/megatron/fused_kernels/
 * scaled_masked_softmax.cpp:70-83, scaled_upper_triang_masked_softmax.cpp:62-70) and through third-party
 * extensions (flash_attn_cuda, deepspeed.ops.adam.FusedAdam, cuBLAS via F.linear). Every entry below names the
 * reference call site it replaces. Conventions:
 *   - plain pointers + sizes only; all pointers are DEVICE pointers unless stated; caller owns every buffer;
 *   - the library never allocates device memory and never synchronises: work is enqueued on `stream`;
 *   - row-major tensors; "ld*" are row strides in ELEMENTS;
 *   - bf16 activations/weights, fp32 statistics / optimizer state;
 *   - return 0 on success, negative fsb_status on error; fsb_last_error() gives a thread-local message;
 *   - no silent no-ops: an unsupported shape/dtype is an error (cf. the silent `default: break` at
 *     scaled_masked_softmax.h:448 in the reference).
 */
#ifndef FSB200_H_
#define FSB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* fsb_stream_t; /* cudaStream_t */

typedef enum {
  FSB_OK = 0,
  FSB_ERR_INVALID = -1,     /* bad shape / alignment / argument */
  FSB_ERR_CUDA = -2,        /* CUDA runtime / driver error      */
  FSB_ERR_UNSUPPORTED = -3  /* valid request this build does not implement */
} fsb_status;

typedef enum { FSB_BF16 = 0, FSB_F32 = 1 } fsb_dtype;

/* ---- library ---------------------------------------------------------------------------------------------- */
int fsb_version(void);               /* 1000*major + minor */
const char* fsb_last_error(void);    /* thread-local, never NULL */
int fsb_num_sms(void);               /* SM count of the current device (148 on B200) */

/* ---- GEMM (tcgen05 + TMA + TMEM) --------------------------------------------------------------------------
 * Replaces F.linear / torch.baddbmm / torch.bmm -> cuBLAS on the hot path:
 *   ColumnParallelLinear.forward  fengshen/models/megatron/mpu/layers.py:347-360  (Y = X W^T)
 *   RowParallelLinear.forward     fengshen/models/megatron/mpu/layers.py:451-470
 *   ParallelLinear (LM head)      fengshen/models/megatron/layers/transformer.py:136-172
 * and their autograd transposes (dgrad / wgrad).
 *
 *   FSB_GEMM_NT : D[M,N] = A[M,K] * B[N,K]^T      forward       (X W^T)
 *   FSB_GEMM_NN : D[M,N] = A[M,K] * B[K,N]        data grad     (dY W)
 *   FSB_GEMM_TN : D[M,N] = A[K,M]^T * B[K,N]      weight grad   (dY^T X)
 *
 * A, B bf16. D bf16 or fp32 (d_dtype). Optional fused epilogue, applied in this order:
 *   acc (+ bias[N]) -> activation -> (+ D_old if accumulate) -> store.
 * bias: fp32 or bf16 vector of length N (bias_dtype), may be NULL.
 * Requirements: pointers 16-byte aligned; lda/ldb/ldd multiples of 8 elements.
 * aux (bf16, may be NULL): receives acc + bias BEFORE the activation (saved for the activation's backward).
 * `batch` > 1 runs independent GEMMs with element strides stride_a/b/d between them (use 1 and 0 otherwise).
 */
typedef enum { FSB_GEMM_NT = 0, FSB_GEMM_NN = 1, FSB_GEMM_TN = 2 } fsb_gemm_layout;
typedef enum {
  FSB_EPI_NONE = 0,
  FSB_EPI_GELU_TANH = 1, /* gelu_new / bias_gelu: layers/activations.py:60-77 */
  FSB_EPI_GELU_ERF = 2   /* erf_gelu: layers/activations.py:98-117; HF BERT hidden_act="gelu" */
} fsb_gemm_epilogue;

int fsb_gemm_bf16(int layout, int64_t M, int64_t N, int64_t K,
                  const void* A, int64_t lda, const void* B, int64_t ldb,
                  void* D, int64_t ldd, int d_dtype,
                  const void* bias, int bias_dtype, int epilogue, int accumulate,
                  void* aux, int64_t ldaux,
                  int64_t batch, int64_t stride_a, int64_t stride_b, int64_t stride_d, int64_t stride_aux,
                  fsb_stream_t stream);

/* ---- fused scaled-dot-product attention (tcgen05, flash-style online softmax) ------------------------------
 * Replaces ParallelSelfAttention.flash_attention (fengshen/models/megatron/layers/transformer.py:410-456; 3P
 * flash_attn_cuda.fwd/bwd, layers/flash_attention.py:31-47,81-101) and the legacy baddbmm -> FusedScaleMaskSoftmax ->
 * bmm path (transformer.py:307-408). q/k/v are read in place from the packed QKV projection output:
 *   element (b, s, head, d) of X lives at X + ((b*seq + s)*x_row_stride + head*x_head_stride + d)  (elements).
 * o: same addressing with o_*_stride. lse: fp32 [batch, nheads, seq_q], log2 domain (internal, consumed by bwd).
 * kv_mask: optional uint8 [batch, seq_kv], 1 = attend (HF additive padding mask), NULL = none.
 * causal=1 masks key > query (requires seq_q == seq_kv). head_dim in {64, 128}.
 */
int fsb_sdpa_fwd(const void* q, const void* k, const void* v, void* o, float* lse,
                 int64_t batch, int64_t seq_q, int64_t seq_kv, int nheads, int head_dim,
                 int64_t q_row_stride, int64_t k_row_stride, int64_t v_row_stride, int64_t o_row_stride,
                 int64_t q_head_stride, int64_t k_head_stride, int64_t v_head_stride, int64_t o_head_stride,
                 float scale, int causal, const uint8_t* kv_mask, fsb_stream_t stream);

/* Backward of fsb_sdpa_fwd. o/lse are the forward outputs; delta: fp32 scratch [batch, nheads, seq_q] (written here);
 * dq/dk/dv are written with their own strides (e.g. the three slices of a packed dQKV buffer). Deterministic. */
int fsb_sdpa_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout,
                 const float* lse, float* delta, void* dq, void* dk, void* dv,
                 int64_t batch, int64_t seq_q, int64_t seq_kv, int nheads, int head_dim,
                 int64_t q_row_stride, int64_t k_row_stride, int64_t v_row_stride, int64_t o_row_stride,
                 int64_t do_row_stride, int64_t dq_row_stride, int64_t dk_row_stride, int64_t dv_row_stride,
                 int64_t q_head_stride, int64_t k_head_stride, int64_t v_head_stride, int64_t o_head_stride,
                 int64_t do_head_stride, int64_t dq_head_stride, int64_t dk_head_stride, int64_t dv_head_stride,
                 float scale, int causal, const uint8_t* kv_mask, fsb_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* FSB200_H_ */
