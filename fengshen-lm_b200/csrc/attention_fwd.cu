// fsb200 — fused scaled-dot-product attention, forward (flash-style online softmax on tcgen05 / TMEM).
// Replaces ParallelSelfAttention.flash_attention (fengshen/models/megatron/layers/transformer.py:410-456, 3P
// flash_attn_cuda.fwd) and the baddbmm -> FusedScaleMaskSoftmax -> bmm path (transformer.py:307-408): same math,
//   O = softmax(scale * Q K^T + mask) V,  mask = causal and/or key-padding,
// without the three repacking copies of q/k/v (transformer.py:419-429): Q, K, V are read straight out of the packed
// QKV projection output through TMA tensor maps (any row/head stride), O is written in [token, head*dim] layout.
//
// CTA = one (batch, head, pair of 128-row Q tiles). 12 warps:
//   warps 0-3 / 4-7 : softmax warpgroup for Q tile 0 / 1 — ONE THREAD PER ROW (TMEM lane == row): no shuffles
//   warp 8          : TMA producer (Q once, K/V tiles through a multi-stage ring)
//   warp 9          : tcgen05.mma issuer  S_i = Q_i K_j^T  (128 x 64 x D)  and  O_i(j) = P_i V_j  (128 x D x 64)
// The two Q tiles ping-pong: while warpgroup 0 does exp2 on S_0 the tensor core computes S_1 / PV_1 and vice versa.
// P is written as bf16 into shared memory in the canonical K-major SWIZZLE_128B layout and fed back as the A operand;
// V is consumed as an MN-major B operand straight from its row-major tile. O accumulates in registers (fp32).
#include "host_common.h"
#include "ptx.cuh"

namespace fsb {

constexpr int ATT_BQ = 128;   // rows per Q tile (= TMEM lanes)
constexpr int ATT_NQ = 2;     // Q tiles per CTA
constexpr int ATT_BKV = 64;   // keys per inner step
constexpr int ATT_THREADS = 384;

template <int D>
struct AttFwdSmem {
  static constexpr int STAGES = (D == 128) ? 4 : 8;  // K/V ring depth: must cover the TMA round trip
  static constexpr int Q_BYTES = ATT_BQ * D * 2;       // per slot
  static constexpr int KV_BYTES = ATT_BKV * D * 2;     // per tensor per stage
  static constexpr int P_BYTES = ATT_BQ * ATT_BKV * 2; // per slot
  static constexpr int OFF_Q = 0;
  static constexpr int OFF_K = OFF_Q + ATT_NQ * Q_BYTES;
  static constexpr int OFF_V = OFF_K + STAGES * KV_BYTES;
  static constexpr int OFF_P = OFF_V + STAGES * KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + ATT_NQ * P_BYTES;
  // q_full, k_full[S], k_empty[S], v_full[S], v_empty[S], s_full[2], p_ready[2], o_full[2]
  static constexpr int NBAR = 1 + 4 * STAGES + 6;
  static constexpr int TOTAL = OFF_BAR + NBAR * 8 + 16 + 1024;
};

struct AttFwdParams {
  __nv_bfloat16* o;
  float* lse;               // [batch, nheads, seq_q], log2 domain
  const uint8_t* kv_mask;   // [batch, seq_kv] (1 = attend) or nullptr
  int64_t o_row_stride, o_head_stride;
  int q_col0, k_col0, v_col0;           // column (element) offset of head 0 inside each tensor map
  int q_head_stride, k_head_stride, v_head_stride;
  int seq_q, seq_kv, nheads, batch;
  int causal;
  float scale_log2;  // softmax scale * log2(e)
};

__device__ __forceinline__ float ex2_approx_f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int kRegs>
__device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kRegs)); }
template <int kRegs>
__device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kRegs)); }

template <int D>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const AttFwdParams p) {
  using S = AttFwdSmem<D>;
  constexpr int STAGES = S::STAGES;
  constexpr int TMEM_COLS = (D == 128) ? 512 : 256;
  constexpr int TM_S = 0;                 // S_i at TM_S + i*64
  constexpr int TM_O = ATT_NQ * ATT_BKV;  // O_i at TM_O + i*D
  constexpr uint32_t IDESC_S = make_idesc_bf16(ATT_BQ, ATT_BKV, 0, 0);
  constexpr uint32_t IDESC_PV = make_idesc_bf16(ATT_BQ, D, 0, 1);

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::OFF_BAR);
  uint64_t* q_full = bars;
  uint64_t* k_full = q_full + 1;
  uint64_t* k_empty = k_full + STAGES;
  uint64_t* v_full = k_empty + STAGES;
  uint64_t* v_empty = v_full + STAGES;
  uint64_t* s_full = v_empty + STAGES;
  uint64_t* p_ready = s_full + 2;
  uint64_t* o_full = p_ready + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_pairs = gridDim.x;
  const int pair = num_pairs - 1 - blockIdx.x;  // heavy (late) causal tiles first
  const int head = blockIdx.y, b = blockIdx.z;
  const int q0 = pair * (ATT_NQ * ATT_BQ);

  // number of KV steps per slot
  const int n_all = (p.seq_kv + ATT_BKV - 1) / ATT_BKV;
  int n_kv[ATT_NQ];
#pragma unroll
  for (int i = 0; i < ATT_NQ; ++i) {
    const int qs = q0 + i * ATT_BQ;
    if (qs >= p.seq_q) n_kv[i] = 0;
    else if (p.causal) n_kv[i] = min(n_all, (min(qs + ATT_BQ, p.seq_q) + ATT_BKV - 1) / ATT_BKV);
    else n_kv[i] = n_all;
  }
  const int n_total = max(n_kv[0], n_kv[1]);

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1); mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1); mbar_init(&p_ready[i], ATT_BQ); mbar_init(&o_full[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<TMEM_COLS>(tmem_ptr_smem);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp >= 8) {
    reg_dec<40>();
    if (warp == 8 && lane == 0) {
      // ===================== TMA producer =====================
      const int qc = p.q_col0 + head * p.q_head_stride;
      const int kc = p.k_col0 + head * p.k_head_stride;
      const int vc = p.v_col0 + head * p.v_head_stride;
      int active = (n_kv[0] > 0) + (n_kv[1] > 0);
      mbar_expect_tx(q_full, active * S::Q_BYTES);
#pragma unroll
      for (int i = 0; i < ATT_NQ; ++i) {
        if (n_kv[i] == 0) continue;
#pragma unroll
        for (int h = 0; h < D / 64; ++h)
          tma_load_3d(smem + S::OFF_Q + i * S::Q_BYTES + h * (ATT_BQ * 128), &tmQ, q_full, qc + h * 64, q0 + i * ATT_BQ, b);
      }
      int st = 0; uint32_t ph = 0;
      for (int j = 0; j < n_total; ++j) {
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], S::KV_BYTES);
#pragma unroll
        for (int h = 0; h < D / 64; ++h)
          tma_load_3d(smem + S::OFF_K + st * S::KV_BYTES + h * (ATT_BKV * 128), &tmK, &k_full[st], kc + h * 64,
                      j * ATT_BKV, b);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], S::KV_BYTES);
#pragma unroll
        for (int h = 0; h < D / 64; ++h)
          tma_load_3d(smem + S::OFF_V + st * S::KV_BYTES + h * (ATT_BKV * 128), &tmV, &v_full[st], vc + h * 64,
                      j * ATT_BKV, b);
        if (++st == STAGES) { st = 0; ph ^= 1; }
      }
    } else if (warp == 9) {
      // ===================== MMA issuer: warp-uniform loop, one elected lane issues =====================
      const uint64_t dsc_q = make_smem_desc_sw128(smem_u32(smem + S::OFF_Q), 0, 1024);
      const uint64_t dsc_k = make_smem_desc_sw128(smem_u32(smem + S::OFF_K), 0, 1024);
      const uint64_t dsc_p = make_smem_desc_sw128(smem_u32(smem + S::OFF_P), 0, 1024);
      const uint64_t dsc_v = make_smem_desc_sw128(smem_u32(smem + S::OFF_V), ATT_BKV * 128, 1024);  // MN-major
      auto issue_S = [&](int slot, int st) {   // caller: inside elect_one()
        const uint64_t da = dsc_q + uint64_t(slot) * (S::Q_BYTES >> 4), db = dsc_k + uint64_t(st) * (S::KV_BYTES >> 4);
#pragma unroll
        for (int kk = 0; kk < D / 16; ++kk)
          umma_bf16(tmem_base + TM_S + slot * ATT_BKV, da + (((kk / 4) * (ATT_BQ * 128) + (kk % 4) * 32) >> 4),
                    db + (((kk / 4) * (ATT_BKV * 128) + (kk % 4) * 32) >> 4), IDESC_S, kk != 0);
      };
      auto issue_PV = [&](int slot, int st) {
        const uint64_t da = dsc_p + uint64_t(slot) * (S::P_BYTES >> 4), db = dsc_v + uint64_t(st) * (S::KV_BYTES >> 4);
#pragma unroll
        for (int kk = 0; kk < ATT_BKV / 16; ++kk)
          umma_bf16(tmem_base + TM_O + slot * D, da + ((kk * 32) >> 4), db + ((kk * 2048) >> 4), IDESC_PV, kk != 0);
      };
      mbar_wait(q_full, 0);
      int st = 0; uint32_t ph = 0;
      if (n_total > 0) {
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
        if (elect_one()) {
          for (int i = 0; i < ATT_NQ; ++i)
            if (n_kv[i] > 0) { issue_S(i, 0); umma_commit(&s_full[i]); }
          umma_commit(&k_empty[0]);
        }
        __syncwarp();
      }
      for (int j = 0; j < n_total; ++j) {
        int st1 = st + 1; uint32_t ph1 = ph;
        if (st1 == STAGES) { st1 = 0; ph1 ^= 1; }
        bool v_waited = false;
        for (int i = 0; i < ATT_NQ; ++i) {
          if (j >= n_kv[i]) continue;
          mbar_wait(&p_ready[i], j & 1);
          if (!v_waited) { mbar_wait(&v_full[st], ph); v_waited = true; }
          const bool more = j + 1 < n_kv[i];
          if (more) mbar_wait(&k_full[st1], ph1);
          tc_fence_after();
          if (elect_one()) {
            issue_PV(i, st);
            umma_commit(&o_full[i]);
            if (more) { issue_S(i, st1); umma_commit(&s_full[i]); }
          }
          __syncwarp();
        }
        if (elect_one()) {
          umma_commit(&v_empty[st]);
          if (j + 1 < n_total) umma_commit(&k_empty[st1]);
        }
        __syncwarp();
        st = st1; ph = ph1;
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    reg_inc<216>();
    const int slot = warp >> 2;
    const int quad = warp & 3;
    const int r_in = quad * 32 + lane;            // row inside the tile == TMEM lane
    const int q_row = q0 + slot * ATT_BQ + r_in;  // position in the sequence
    const int n_mine = n_kv[slot];
    const uint32_t t_lane = tmem_base + (uint32_t(quad * 32) << 16);
    uint8_t* sP = smem + S::OFF_P + slot * S::P_BYTES + r_in * 128;
    const int sw = r_in & 7;
    const uint8_t* mrow = p.kv_mask ? p.kv_mask + int64_t(b) * p.seq_kv : nullptr;

    float m_run = -INFINITY, l_run = 0.f;
    float o[D];
#pragma unroll
    for (int d = 0; d < D; ++d) o[d] = 0.f;

    for (int j = 0; j < n_mine; ++j) {
      mbar_wait(&s_full[slot], j & 1);
      tc_fence_after();
      uint32_t r0[32], r1[32];
      tmem_ld32(t_lane + TM_S + slot * ATT_BKV, r0);
      tmem_ld32(t_lane + TM_S + slot * ATT_BKV + 32, r1);
      tmem_ld_wait();
      const int kv0 = j * ATT_BKV;
      const bool need_mask = (p.causal && kv0 + ATT_BKV - 1 > q0 + slot * ATT_BQ) || (kv0 + ATT_BKV > p.seq_kv) || mrow;
      float mx = -INFINITY;
      if (!need_mask) {
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const float a = __uint_as_float(r0[c]) * p.scale_log2, bq = __uint_as_float(r1[c]) * p.scale_log2;
          r0[c] = __float_as_uint(a); r1[c] = __float_as_uint(bq);
          mx = fmaxf(mx, fmaxf(a, bq));
        }
      } else {
#pragma unroll
        for (int c = 0; c < 64; ++c) {
          float s = __uint_as_float(c < 32 ? r0[c & 31] : r1[c & 31]) * p.scale_log2;
          const int col = kv0 + c;
          bool keep = col < p.seq_kv && !(p.causal && col > q_row);
          if (keep && mrow) keep = mrow[col] != 0;
          s = keep ? s : -INFINITY;
          if (c < 32) r0[c & 31] = __float_as_uint(s); else r1[c & 31] = __float_as_uint(s);
          mx = fmaxf(mx, s);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = ex2_approx_f(m_run - m_use);  // m_run == -inf -> 0
      float sum = 0.f;
      uint32_t pk[32];
#pragma unroll
      for (int c = 0; c < 64; c += 2) {
        const float a = ex2_approx_f(__uint_as_float(c < 32 ? r0[c & 31] : r1[c & 31]) - m_use);
        const float bb = ex2_approx_f(__uint_as_float(c + 1 < 32 ? r0[(c + 1) & 31] : r1[(c + 1) & 31]) - m_use);
        pk[c >> 1] = pack_bf16x2(a, bb);
        // sum what the tensor core will actually see (bf16-rounded), keeps rows normalised
        sum += bf16lo(pk[c >> 1]) + bf16hi(pk[c >> 1]);
      }
#pragma unroll
      for (int ch = 0; ch < 8; ++ch) {
        uint4 q4 = make_uint4(pk[ch * 4], pk[ch * 4 + 1], pk[ch * 4 + 2], pk[ch * 4 + 3]);
        *reinterpret_cast<uint4*>(sP + ((ch ^ sw) << 4)) = q4;
      }
      fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tc_fence_before();     // order our tcgen05.ld of S / O before the MMAs that overwrite them
      mbar_arrive(&p_ready[slot]);
      l_run = l_run * alpha + sum;
      m_run = m_new;

      mbar_wait(&o_full[slot], j & 1);
      tc_fence_after();
#pragma unroll
      for (int ch = 0; ch < D / 32; ++ch) {
        uint32_t t[32];
        tmem_ld32(t_lane + TM_O + slot * D + ch * 32, t);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) o[ch * 32 + c] = o[ch * 32 + c] * alpha + __uint_as_float(t[c]);
      }
    }
    // ---- epilogue: normalise, store O (bf16) and LSE (log2 domain)
    if (n_mine > 0 && q_row < p.seq_q) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      __nv_bfloat16* op = p.o + (int64_t(b) * p.seq_q + q_row) * p.o_row_stride + int64_t(head) * p.o_head_stride;
#pragma unroll
      for (int d = 0; d < D; d += 8) {
        float f[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) f[t] = o[d + t] * inv;
        *reinterpret_cast<uint4*>(op + d) = pack8(f);
      }
      p.lse[(int64_t(b) * p.nheads + head) * p.seq_q + q_row] = l_run > 0.f ? m_run + log2f(l_run) : INFINITY;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) {
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}

// Tensor map over a packed activation buffer: dims {row_width, seq, batch}; box {64, box_rows, 1}.
int make_attn_tmap(CUtensorMap* tm, const void* base, int64_t row_stride, int64_t width, int64_t seq, int64_t batch,
                   int box_rows) {
  uint64_t dims[3] = {uint64_t(width), uint64_t(seq), uint64_t(batch)};
  uint64_t strides[2] = {uint64_t(row_stride) * 2, uint64_t(seq) * uint64_t(row_stride) * 2};
  uint32_t box[3] = {64, uint32_t(box_rows), 1};
  return make_tmap_bf16(tm, base, 3, dims, strides, box);
}

template <int D>
static int launch_attn_fwd(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const AttFwdParams& p,
                           cudaStream_t st) {
  using S = AttFwdSmem<D>;
  static bool configured = false;
  auto kern = attn_fwd_kernel<D>;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
    if (e != cudaSuccess) {
      set_error("sdpa_fwd: cudaFuncSetAttribute(%d) failed: %s", S::TOTAL, cudaGetErrorString(e));
      return FSB_ERR_CUDA;
    }
    configured = true;
  }
  dim3 grid((p.seq_q + ATT_NQ * ATT_BQ - 1) / (ATT_NQ * ATT_BQ), p.nheads, p.batch);
  kern<<<grid, ATT_THREADS, S::TOTAL, st>>>(tq, tk, tv, p);
  FSB_CUDA_LAUNCH_CHECK();
  return FSB_OK;
}

}  // namespace fsb

using namespace fsb;

extern "C" int fsb_sdpa_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int64_t batch,
                            int64_t seq_q, int64_t seq_kv, int nheads, int head_dim, int64_t q_row_stride,
                            int64_t k_row_stride, int64_t v_row_stride, int64_t o_row_stride, int64_t q_head_stride,
                            int64_t k_head_stride, int64_t v_head_stride, int64_t o_head_stride, float scale, int causal,
                            const uint8_t* kv_mask, fsb_stream_t st) {
  FSB_REQUIRE(q && k && v && o && lse, "sdpa_fwd: null pointer");
  FSB_REQUIRE(head_dim == 64 || head_dim == 128, "sdpa_fwd: head_dim %d unsupported (64 or 128)", head_dim);
  FSB_REQUIRE(batch > 0 && seq_q > 0 && seq_kv > 0 && nheads > 0 && batch < 65536 && nheads < 65536, "sdpa_fwd: bad dims");
  FSB_REQUIRE(!causal || seq_q == seq_kv, "sdpa_fwd: causal needs seq_q == seq_kv");
  FSB_REQUIRE(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), "sdpa_fwd: 16-byte alignment required");
  FSB_REQUIRE((q_row_stride | k_row_stride | v_row_stride | o_row_stride | q_head_stride | k_head_stride |
               v_head_stride | o_head_stride) % 8 == 0,
              "sdpa_fwd: strides must be multiples of 8 elements");
  CUtensorMap tq, tk, tv;
  int rc;
  if ((rc = make_attn_tmap(&tq, q, q_row_stride, (nheads - 1) * q_head_stride + head_dim, seq_q, batch, ATT_BQ))) return rc;
  if ((rc = make_attn_tmap(&tk, k, k_row_stride, (nheads - 1) * k_head_stride + head_dim, seq_kv, batch, ATT_BKV))) return rc;
  if ((rc = make_attn_tmap(&tv, v, v_row_stride, (nheads - 1) * v_head_stride + head_dim, seq_kv, batch, ATT_BKV))) return rc;
  AttFwdParams p;
  p.o = (__nv_bfloat16*)o; p.lse = lse; p.kv_mask = kv_mask;
  p.o_row_stride = o_row_stride; p.o_head_stride = o_head_stride;
  p.q_col0 = 0; p.k_col0 = 0; p.v_col0 = 0;
  p.q_head_stride = int(q_head_stride); p.k_head_stride = int(k_head_stride); p.v_head_stride = int(v_head_stride);
  p.seq_q = int(seq_q); p.seq_kv = int(seq_kv); p.nheads = nheads; p.batch = int(batch);
  p.causal = causal;
  p.scale_log2 = scale * 1.4426950408889634f;
  return head_dim == 128 ? launch_attn_fwd<128>(tq, tk, tv, p, (cudaStream_t)st)
                         : launch_attn_fwd<64>(tq, tk, tv, p, (cudaStream_t)st);
}
