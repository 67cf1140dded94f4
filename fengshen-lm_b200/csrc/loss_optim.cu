// fsb200 — fused softmax-cross-entropy (forward + backward in one launch) and the flat-shard AdamW update.
//   cross-entropy : torch.nn.CrossEntropyLoss()(shift_logits, shift_labels) at fengshen/models/llama/modeling_llama.py:334-339
//                   (mean over labels != -100). The shift-by-one is done by index arithmetic here instead of the
//                   `.contiguous()` copy of the shifted logits at :336.
//   AdamW         : deepspeed.ops.adam.FusedAdam(adam_w_mode=True) selected at fengshen/models/model_utils.py:69-72,
//                   restated in torch.optim.AdamW's operation order (the CPU oracle's optimiser), on the rank's flat fp32
//                   shard {master, m, v} with 16-bit gradient in and 16-bit parameter out — the data flow of the ZeRO-1/2
//                   optimizer described in SURVEY.md Appendix D. Optional device-side gradient scale (clip coefficient).
//   grad-norm     : local sum of squares -> (all-reduced by the host) -> clip coefficient, all on device.
#include "host_common.h"
#include "ptx.cuh"

namespace fsb {

constexpr int XENT_THREADS = 512;

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// labels are indexed with the causal shift: row t = (b, s) uses labels[b*S + s + shift] and is ignored when s + shift >= S.
__global__ void count_valid_kernel(const int64_t* __restrict__ labels, int64_t rows, int seq_len, int shift,
                                   int ignore_index, int* __restrict__ count) {
  __shared__ int sm[32];
  int c = 0;
  for (int64_t t = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; t < rows; t += int64_t(gridDim.x) * blockDim.x) {
    const int s = int(t % seq_len);
    if (s + shift < seq_len && labels[t + shift] != ignore_index) ++c;
  }
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x < 32) {
    c = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0;
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (threadIdx.x == 0) atomicAdd(count, c);
  }
}

// One CTA per row. Pass 1: online max / sum-exp over the row (16-byte loads). Pass 2: re-read (L2-resident) and write
// dlogits = (softmax - onehot) * grad_scale / n_valid (may alias logits). row_loss[t] = lse - logit[label] (0 if ignored).
__global__ void __launch_bounds__(XENT_THREADS) xent_fwd_bwd_kernel(
    const __nv_bfloat16* logits, const int64_t* __restrict__ labels, __nv_bfloat16* dlogits /* may alias logits */,
    float* __restrict__ row_loss, const int* __restrict__ n_valid, int64_t rows, int V, int64_t ld, int seq_len, int shift,
    int ignore_index, float grad_scale) {
  __shared__ float red_m[XENT_THREADS / 32], red_s[XENT_THREADS / 32];
  __shared__ float sh_max, sh_sum;
  const int nv = V >> 3;  // V % 8 == 0 enforced by the host
  for (int64_t t = blockIdx.x; t < rows; t += gridDim.x) {
    const int s = int(t % seq_len);
    const bool in_range = (s + shift < seq_len);
    const int64_t lab = in_range ? labels[t + shift] : ignore_index;
    const bool valid = in_range && lab != ignore_index;
    const uint4* row = reinterpret_cast<const uint4*>(logits + t * ld);
    uint4* drow = dlogits ? reinterpret_cast<uint4*>(dlogits + t * ld) : nullptr;
    if (!valid) {  // ignored row: zero gradient, zero loss (CTA-uniform branch)
      if (drow) for (int i = threadIdx.x; i < nv; i += XENT_THREADS) drow[i] = make_uint4(0, 0, 0, 0);
      if (threadIdx.x == 0) row_loss[t] = 0.f;
      continue;
    }
    float lab_logit = 0.f;  // read before pass 2 may overwrite it (dlogits may alias logits)
    if (threadIdx.x == 0) lab_logit = __bfloat162float(logits[t * ld + lab]);
    float m = -INFINITY, sum = 0.f;
    // four independent 16-byte loads in flight per thread and iteration (the single-load form left the row fetch latency-
    // bound: 4.1 TB/s of DRAM-level traffic on [32768, 50264]); one running-max update per 32 values
    for (int i0 = threadIdx.x; i0 < nv; i0 += 4 * XENT_THREADS) {
      uint4 q[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * XENT_THREADS;
        q[u] = i < nv ? row[i] : make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u);   // bf16 -inf pairs
      }
      float f[4][8];
      float lm = -INFINITY;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        unpack8(q[u], f[u]);
#pragma unroll
        for (int j = 0; j < 8; ++j) lm = fmaxf(lm, f[u][j]);
      }
      const float nm = fmaxf(m, lm);
      float acc = 0.f;
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc += __expf(f[u][j] - nm);    // exp(-inf) = 0 for the padding lanes
      sum = sum * __expf(m - nm) + acc;
      m = nm;
    }
    // block combine
    const float wm = warp_max(m);
    sum = warp_sum_f(m == -INFINITY ? 0.f : sum * __expf(m - wm));
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) { red_m[w] = wm; red_s[w] = sum; }
    __syncthreads();
    if (w == 0) {
      float mm = l < XENT_THREADS / 32 ? red_m[l] : -INFINITY;
      float ss = l < XENT_THREADS / 32 ? red_s[l] : 0.f;
      const float gm = warp_max(mm);
      ss = warp_sum_f(mm == -INFINITY ? 0.f : ss * __expf(mm - gm));
      if (l == 0) { sh_max = gm; sh_sum = ss; }
    }
    __syncthreads();
    const float gmax = sh_max, gsum = sh_sum;
    const float lse = gmax + logf(gsum);
    if (threadIdx.x == 0) row_loss[t] = lse - lab_logit;
    if (drow) {
      const float scale = grad_scale / float(*n_valid);
      const float inv = 1.f / gsum;
      for (int i0 = threadIdx.x; i0 < nv; i0 += 4 * XENT_THREADS) {
        uint4 q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * XENT_THREADS;
          if (i < nv) q[u] = row[i];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * XENT_THREADS;
          if (i >= nv) continue;
          float f[8];
          unpack8(q[u], f);
          const int c0 = i * 8;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float p = __expf(f[j] - gmax) * inv;
            if (c0 + j == lab) p -= 1.f;
            f[j] = p * scale;
          }
          drow[i] = pack8(f);
        }
      }
    }
  }
}

// loss = sum(row_loss) / n_valid   (single CTA, deterministic)
__global__ void __launch_bounds__(1024) loss_reduce_kernel(const float* __restrict__ row_loss, int64_t rows,
                                                           const int* __restrict__ n_valid, float* __restrict__ loss) {
  __shared__ float sm[32];
  float s = 0.f;
  for (int64_t i = threadIdx.x; i < rows; i += 1024) s += row_loss[i];
  s = warp_sum_f(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = warp_sum_f(sm[threadIdx.x]);
    if (threadIdx.x == 0) *loss = s / fmaxf(float(*n_valid), 1.f);
  }
}

// ------------------------------------------------------------------------------------------------ AdamW (flat shard)
struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt;  // bc1 = 1-b1^t, bc2_sqrt = sqrt(1-b2^t)
};

template <bool kGradF32>
__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                                                    const void* __restrict__ g_, __nv_bfloat16* __restrict__ p16,
                                                    int64_t n, AdamArgs a, const float* __restrict__ grad_scale,
                                                    const float* __restrict__ hyper) {
  const float gs = grad_scale ? *grad_scale : 1.f;
  const int64_t nvec = n >> 2;
  if (hyper != nullptr) {   // per-step scalars from DEVICE memory: the launch is then identical every step (CUDA-graph replay)
    a.lr = hyper[0]; a.bc1 = hyper[1]; a.bc2_sqrt = hyper[2];
  }
  const float step = a.lr / a.bc1;
  const float decay = 1.f - a.lr * a.weight_decay;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < nvec; i += int64_t(gridDim.x) * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float g[4];
    if (kGradF32) {
      float4 gg = reinterpret_cast<const float4*>(g_)[i];
      g[0] = gg.x; g[1] = gg.y; g[2] = gg.z; g[3] = gg.w;
    } else {
      uint2 gg = reinterpret_cast<const uint2*>(g_)[i];
      g[0] = bf16lo(gg.x); g[1] = bf16hi(gg.x); g[2] = bf16lo(gg.y); g[3] = bf16hi(gg.y);
    }
    float pa[4] = {pp.x, pp.y, pp.z, pp.w}, ma[4] = {mm.x, mm.y, mm.z, mm.w}, va[4] = {vv.x, vv.y, vv.z, vv.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float gj = g[j] * gs;
      pa[j] *= decay;                                        // p.mul_(1 - lr*wd)
      ma[j] = a.beta1 * ma[j] + (1.f - a.beta1) * gj;        // exp_avg.lerp_(grad, 1-b1)
      va[j] = a.beta2 * va[j] + (1.f - a.beta2) * gj * gj;   // exp_avg_sq
      const float denom = sqrtf(va[j]) / a.bc2_sqrt + a.eps;
      pa[j] -= step * (ma[j] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = make_float4(pa[0], pa[1], pa[2], pa[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(ma[0], ma[1], ma[2], ma[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(va[0], va[1], va[2], va[3]);
    if (p16) {
      uint2 o;
      o.x = pack_bf16x2(pa[0], pa[1]);
      o.y = pack_bf16x2(pa[2], pa[3]);
      reinterpret_cast<uint2*>(p16)[i] = o;
    }
  }
}

// partial[blockIdx.x] = sum of squares of this CTA's slice
template <bool kF32>
__global__ void __launch_bounds__(256) sumsq_kernel(const void* __restrict__ x_, int64_t n, float* __restrict__ partial) {
  __shared__ float sm[8];
  float s = 0.f;
  const int64_t nvec = n >> 2;
  for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < nvec; i += int64_t(gridDim.x) * blockDim.x) {
    if (kF32) {
      float4 q = reinterpret_cast<const float4*>(x_)[i];
      s += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
    } else {
      uint2 q = reinterpret_cast<const uint2*>(x_)[i];
      float a = bf16lo(q.x), b = bf16hi(q.x), c = bf16lo(q.y), d = bf16hi(q.y);
      s += a * a + b * b + c * c + d * d;
    }
  }
  s = warp_sum_f(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = threadIdx.x < 8 ? sm[threadIdx.x] : 0.f;
    s = warp_sum_f(s);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
  }
}
// out[0] (+)= sum(partial)
__global__ void __launch_bounds__(1024) sum_partials_kernel(const float* __restrict__ partial, int n, float* __restrict__ out,
                                                            int accumulate) {
  __shared__ float sm[32];
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += 1024) s += partial[i];
  s = warp_sum_f(s);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    s = warp_sum_f(sm[threadIdx.x]);
    if (threadIdx.x == 0) out[0] = accumulate ? out[0] + s : s;
  }
}
// coef = min(1, max_norm / (sqrt(sumsq) + 1e-6)) ; norm_out = sqrt(sumsq)   (torch.nn.utils.clip_grad_norm_ form)
__global__ void clip_coef_kernel(const float* __restrict__ sumsq, float max_norm, float* __restrict__ coef,
                                 float* __restrict__ norm_out) {
  const float nrm = sqrtf(*sumsq);
  if (norm_out) *norm_out = nrm;
  float c = max_norm / (nrm + 1e-6f);
  *coef = (max_norm > 0.f && c < 1.f) ? c : 1.f;
}

}  // namespace fsb

using namespace fsb;

extern "C" int fsb_softmax_xent_fwd_bwd(const void* logits, const int64_t* labels, void* dlogits, float* row_loss,
                                        float* loss, int* n_valid, int64_t rows, int64_t vocab, int64_t ld,
                                        int64_t seq_len, int shift, int ignore_index, float grad_scale,
                                        fsb_stream_t st_) {
  cudaStream_t st = (cudaStream_t)st_;
  FSB_REQUIRE(logits && labels && row_loss && loss && n_valid, "xent: null pointer");
  FSB_REQUIRE(rows > 0 && vocab > 0 && vocab % 8 == 0 && ld % 8 == 0 && ld >= vocab, "xent: vocab/ld must be multiples of 8");
  FSB_REQUIRE(seq_len > 0 && rows % seq_len == 0 && shift >= 0, "xent: rows must be a multiple of seq_len");
  FSB_REQUIRE(aligned16(logits) && aligned16(dlogits), "xent: alignment");
  cudaError_t e = cudaMemsetAsync(n_valid, 0, sizeof(int), st);
  if (e != cudaSuccess) { set_error("xent: memset failed: %s", cudaGetErrorString(e)); return FSB_ERR_CUDA; }
  int cg = int((rows + 255) / 256);
  if (cg > 4 * num_sms()) cg = 4 * num_sms();
  count_valid_kernel<<<cg, 256, 0, st>>>(labels, rows, int(seq_len), shift, ignore_index, n_valid);
  FSB_CUDA_LAUNCH_CHECK();
  int grid = int(rows < int64_t(8) * num_sms() ? rows : int64_t(8) * num_sms());
  xent_fwd_bwd_kernel<<<grid, XENT_THREADS, 0, st>>>((const __nv_bfloat16*)logits, labels, (__nv_bfloat16*)dlogits,
                                                     row_loss, n_valid, rows, int(vocab), ld, int(seq_len), shift,
                                                     ignore_index, grad_scale);
  FSB_CUDA_LAUNCH_CHECK();
  loss_reduce_kernel<<<1, 1024, 0, st>>>(row_loss, rows, n_valid, loss);
  FSB_CUDA_LAUNCH_CHECK();
  return FSB_OK;
}

extern "C" int fsb_adamw_flat(float* master, float* exp_avg, float* exp_avg_sq, const void* grad, int grad_dtype,
                              void* param16, int64_t n, float lr, float beta1, float beta2, float eps,
                              float weight_decay, int64_t step, const float* grad_scale, const float* hyper,
                              fsb_stream_t st) {
  FSB_REQUIRE(master && exp_avg && exp_avg_sq && grad && n > 0 && (step >= 1 || hyper != nullptr), "adamw: bad args");
  if (step < 1) step = 1;   // placeholder when the device scalars carry the bias corrections
  FSB_REQUIRE(n % 4 == 0, "adamw: n=%ld must be a multiple of 4 (pad the flat shard)", (long)n);
  FSB_REQUIRE(aligned16(master) && aligned16(exp_avg) && aligned16(exp_avg_sq) &&
                  (reinterpret_cast<uintptr_t>(grad) & 7) == 0 && (reinterpret_cast<uintptr_t>(param16) & 7) == 0,
              "adamw: alignment");
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bc1 = float(1.0 - pow(double(beta1), double(step)));
  a.bc2_sqrt = float(sqrt(1.0 - pow(double(beta2), double(step))));
  int64_t blocks = (n / 4 + 255) / 256;
  const int64_t cap = int64_t(num_sms()) * 16;
  const int g = int(blocks < cap ? blocks : cap);
  if (grad_dtype == FSB_F32)
    adamw_kernel<true><<<g, 256, 0, (cudaStream_t)st>>>(master, exp_avg, exp_avg_sq, grad, (__nv_bfloat16*)param16, n, a,
                                                        grad_scale, hyper);
  else
    adamw_kernel<false><<<g, 256, 0, (cudaStream_t)st>>>(master, exp_avg, exp_avg_sq, grad, (__nv_bfloat16*)param16, n, a,
                                                         grad_scale, hyper);
  FSB_CUDA_LAUNCH_CHECK();
  return FSB_OK;
}

extern "C" size_t fsb_sumsq_workspace_bytes(void) { return size_t(num_sms()) * 8 * sizeof(float); }

extern "C" int fsb_sumsq(const void* x, int dtype, int64_t n, float* out, int accumulate, void* workspace,
                         size_t workspace_bytes, fsb_stream_t st) {
  FSB_REQUIRE(x && out && workspace && n > 0 && n % 4 == 0, "sumsq: bad args (n %% 4 == 0 required)");
  const int g = num_sms() * 8;
  FSB_REQUIRE(workspace_bytes >= size_t(g) * sizeof(float), "sumsq: workspace too small");
  if (dtype == FSB_F32)
    sumsq_kernel<true><<<g, 256, 0, (cudaStream_t)st>>>(x, n, (float*)workspace);
  else
    sumsq_kernel<false><<<g, 256, 0, (cudaStream_t)st>>>(x, n, (float*)workspace);
  FSB_CUDA_LAUNCH_CHECK();
  sum_partials_kernel<<<1, 1024, 0, (cudaStream_t)st>>>((const float*)workspace, g, out, accumulate);
  FSB_CUDA_LAUNCH_CHECK();
  return FSB_OK;
}

extern "C" int fsb_clip_coef(const float* sumsq, float max_norm, float* coef, float* norm_out, fsb_stream_t st) {
  FSB_REQUIRE(sumsq && coef, "clip_coef: null pointer");
  clip_coef_kernel<<<1, 1, 0, (cudaStream_t)st>>>(sumsq, max_norm, coef, norm_out);
  FSB_CUDA_LAUNCH_CHECK();
  return FSB_OK;
}
