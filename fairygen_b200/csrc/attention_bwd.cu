// fairygen_b200 — flash-attention BACKWARD on tcgen05 tensor cores (sm_100a), head_dim 128, non-causal.
//
// Config 5 of BASELINE.json (stage-2 motion-LoRA fine-tune step) back-propagates through the same
// flash_attention()/AttentionModule the forward replaces (reference animation/diffsynth/models/
// wan_video_dit.py:27-60, 113-120; the reference gets this backward from torch autograd over FA2/SDPA).
//
//   P  = 2^(q·kᵀ·scale·log2e − L)            L = log2-domain log-sum-exp saved by fgb_attn_fwd_ex
//   dV = Pᵀ·dO          dP = dO·Vᵀ          dS = P ∘ (dP − Δ),   Δ = rowsum(dO ∘ O)
//   dQ = scale·dS·K     dK = scale·dSᵀ·Q
//
// Two launches of ONE templated kernel, so nothing is accumulated across CTAs (no atomics, deterministic):
//   kDKV = true   CTA = (128 keys, head): K and V tiles resident in smem, Q/dO/L/Δ streamed 64 query rows at a
//                 time. The tensor core computes the TRANSPOSED score blocks Sᵀ = K·Qᵀ and dPᵀ = V·dOᵀ (keys on the
//                 128 TMEM lanes, queries on columns), so that Pᵀ and dSᵀ — written back into TMEM as packed bf16
//                 by the 128 softmax threads, one key row each — are directly the TMEM A operands of
//                 dV += Pᵀ·dO and dK += dSᵀ·Q (B = the same dO / Q smem tiles read MN-major). No smem round trip.
//   kDKV = false  CTA = (128 queries, head): Q and dO resident, K/V streamed 64 keys at a time; S = Q·Kᵀ,
//                 dP = dO·Vᵀ, dS (TMEM, bf16) is the A operand of dQ += dS·K. L and Δ are per-thread scalars.
// Warp roles (320 threads): warps 0-3 / 4-7 softmax + epilogue for score columns [0,32) / [32,64) of every streamed
// sub-tile (TMEM lane quarter = warp % 4; two warps per SM sub-partition hide each other's MUFU / TMEM latencies — no
// row statistics are exchanged in the backward, L and Δ come from the forward), warp 8 MMA issuer, warp 9 TMA.
// The two score blocks are double-buffered in TMEM, so the MMAs of sub-tile i+1 run under the softmax of i.
// TMEM: [S|dP] x 2 buffers (256 columns), accumulators dV (128) and dK or dQ (128).
// Out-of-range rows need no masks: TMA zero-fills them, so they contribute exact zeros to every product.
#include "common.cuh"
#include "host.h"

namespace fgb {

constexpr int kBwdThreads = 320;   // warps 0-7 softmax / epilogue, warp 8 MMA issuer, warp 9 TMA producer
constexpr int kBwdMmaWarp = 8, kBwdTmaWarp = 9;
constexpr int kRes = 128;                       // resident rows per CTA
constexpr int kSub = 64;                        // streamed rows per step
constexpr int kResBox = kRes * 64 * 2;          // 16 KB: [128 rows][64 cols] box
constexpr int kResBytes = 2 * kResBox;          // 32 KB tile
constexpr int kSubBox = kSub * 64 * 2;          // 8 KB: [64 rows][64 cols] box
constexpr int kSubBytes = 2 * kSubBox;          // 16 KB tile
constexpr int kBwdStages = 3;
constexpr int kStatBytes = kSub * 4;            // 256 B of L or Δ per sub-tile
constexpr int kBwdSmem = 2 * kResBytes + kBwdStages * (2 * kSubBytes + 2 * kStatBytes) + 1024 + 256;

struct AttnBwdParams {
  const float* lse;     // [heads][ld_stat]
  const float* delta;   // [heads][ld_stat]
  __nv_bfloat16* out0;  // dV (kDKV) / unused
  __nv_bfloat16* out1;  // dK (kDKV) / dQ
  int64_t ld0, ld1;
  int64_t ld_stat;
  int32_t rows_res;     // rows of the resident operand (s_kv for kDKV, s_q otherwise)
  int32_t rows_str;     // rows of the streamed operand
  float scale, scale_log2;
};

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 1-D bulk copy global -> shared, completion on an mbarrier (bytes multiple of 16, both sides 16-byte aligned)
__device__ __forceinline__ void bulk_load_1d(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <bool kDKV>
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_r1, const __grid_constant__ CUtensorMap tmap_r2,
                const __grid_constant__ CUtensorMap tmap_t1, const __grid_constant__ CUtensorMap tmap_t2,
                const AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_r1 = smem;                                   // resident A operand of the score MMA (K or Q)
  uint8_t* smem_r2 = smem + kResBytes;                       // resident A operand of the dP MMA (V or dO)
  uint8_t* smem_t1 = smem + 2 * kResBytes;                   // [stages] streamed Q (kDKV) / K
  uint8_t* smem_t2 = smem_t1 + kBwdStages * kSubBytes;       // [stages] streamed dO (kDKV) / V
  float* smem_stat = reinterpret_cast<float*>(smem_t2 + kBwdStages * kSubBytes);  // [stages][2][64] (L, Δ), kDKV only
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(smem_stat) + kBwdStages * 2 * kStatBytes);
  uint64_t* res_full = bars;                 // [1]
  uint64_t* full = bars + 1;                 // [stages] TMA -> MMA / softmax
  uint64_t* empty = full + kBwdStages;       // [stages] MMA -> TMA
  uint64_t* s_full = empty + kBwdStages;     // [2] score buffers written
  uint64_t* p_ready = s_full + 2;            // [2] P / dS stored by 128 threads
  uint64_t* acc_done = p_ready + 2;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int r0 = blockIdx.x * kRes;
  const int n_sub = (p.rows_str + kSub - 1) / kSub;

  if (warp == kBwdTmaWarp && lane == 0) {
    tma_prefetch_desc(&tmap_r1);
    tma_prefetch_desc(&tmap_r2);
    tma_prefetch_desc(&tmap_t1);
    tma_prefetch_desc(&tmap_t2);
    mbar_init(res_full, 1);
    for (int i = 0; i < kBwdStages; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 256);
    }
    mbar_init(acc_done, 1);
    fence_mbar_init();
  }
  if (warp == kBwdMmaWarp) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // TMEM columns: buffer b: scores at b*128 + [0,64), dP at b*128 + [64,128); acc0 (dV) at 256; acc1 (dK / dQ) at 384.

  if (warp == kBwdTmaWarp) {
    if (elect_one()) {
      // ------------------------------- TMA producer -------------------------------
      mbar_expect_tx(res_full, 2 * kResBytes);
      for (int b = 0; b < 2; ++b) {
        tma_load_2d(smem_r1 + b * kResBox, &tmap_r1, res_full, head * 128 + b * 64, r0, kEvictFirst);
        tma_load_2d(smem_r2 + b * kResBox, &tmap_r2, res_full, head * 128 + b * 64, r0, kEvictFirst);
      }
      int st = 0;
      uint32_t ph = 0;
      for (int i = 0; i < n_sub; ++i) {
        mbar_wait(&empty[st], ph ^ 1);
        mbar_expect_tx(&full[st], 2 * kSubBytes + (kDKV ? 2 * kStatBytes : 0));
        for (int b = 0; b < 2; ++b) {
          tma_load_2d(smem_t1 + st * kSubBytes + b * kSubBox, &tmap_t1, &full[st], head * 128 + b * 64, i * kSub, kEvictLast);
          tma_load_2d(smem_t2 + st * kSubBytes + b * kSubBox, &tmap_t2, &full[st], head * 128 + b * 64, i * kSub, kEvictLast);
        }
        if (kDKV) {
          const int64_t off = static_cast<int64_t>(head) * p.ld_stat + static_cast<int64_t>(i) * kSub;
          bulk_load_1d(smem_stat + st * 2 * kSub, p.lse + off, kStatBytes, &full[st]);
          bulk_load_1d(smem_stat + st * 2 * kSub + kSub, p.delta + off, kStatBytes, &full[st]);
        }
        if (++st == kBwdStages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == kBwdMmaWarp) {
    if (elect_one()) {
      // ------------------------------- MMA issuer ---------------------------------
      constexpr uint32_t idesc_s = make_idesc_bf16(128, kSub, 0, 0);    // resident (K-major) x streamed (K-major), N = 64
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, 128, 0, 1);   // TMEM bf16 A x streamed tile read MN-major, N = 128
      const uint64_t r1_desc = make_sdesc_sw128(smem_u32(smem_r1), 16, 1024);
      const uint64_t r2_desc = make_sdesc_sw128(smem_u32(smem_r2), 16, 1024);
      const uint64_t t1_desc = make_sdesc_sw128(smem_u32(smem_t1), 16, 1024);
      const uint64_t t2_desc = make_sdesc_sw128(smem_u32(smem_t2), 16, 1024);
      const uint64_t t1_mn = make_sdesc_sw128(smem_u32(smem_t1), kSubBox, 1024);
      const uint64_t t2_mn = make_sdesc_sw128(smem_u32(smem_t2), kSubBox, 1024);
      auto issue_scores = [&](int buf, int st) {
        const uint64_t st_off = static_cast<uint64_t>((st * kSubBytes) >> 4);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // 16 head-dim elements per MMA
          const uint64_t ra = static_cast<uint64_t>(((kk >> 2) * kResBox + (kk & 3) * 32) >> 4);
          const uint64_t tb = static_cast<uint64_t>(((kk >> 2) * kSubBox + (kk & 3) * 32) >> 4);
          umma_ss(tmem_base + buf * 128, r1_desc + ra, t1_desc + st_off + tb, idesc_s, kk != 0);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t ra = static_cast<uint64_t>(((kk >> 2) * kResBox + (kk & 3) * 32) >> 4);
          const uint64_t tb = static_cast<uint64_t>(((kk >> 2) * kSubBox + (kk & 3) * 32) >> 4);
          umma_ss(tmem_base + buf * 128 + 64, r2_desc + ra, t2_desc + st_off + tb, idesc_s, kk != 0);
        }
        tc_commit(&s_full[buf]);
      };
      mbar_wait(res_full, 0);
      mbar_wait(&full[0], 0);
      tc_fence_after();
      issue_scores(0, 0);
      int st = 0, st_next = 1 % kBwdStages;
      uint32_t ph_next = (kBwdStages == 1) ? 1u : 0u;
      for (int i = 0; i < n_sub; ++i) {
        const int buf = i & 1;
        if (i + 1 < n_sub) {
          // The other score buffer is free: the accumulate MMAs that read its P / dS (sub-tile i-1) were issued
          // in the previous iteration and the tensor pipe executes in order.
          mbar_wait(&full[st_next], ph_next);
          tc_fence_after();
          issue_scores(buf ^ 1, st_next);
        }
        mbar_wait(&p_ready[buf], (i >> 1) & 1);
        tc_fence_after();
        const uint64_t st_off = static_cast<uint64_t>((st * kSubBytes) >> 4);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {  // 16 streamed rows per MMA (8 packed bf16x2 TMEM columns of A)
          const uint64_t kb = static_cast<uint64_t>((kk * 2048) >> 4);
          if (kDKV) umma_ts(tmem_base + 256, tmem_base + buf * 128 + kk * 8, t2_mn + st_off + kb, idesc_acc, (i | kk) != 0);
          umma_ts(tmem_base + 384, tmem_base + buf * 128 + 64 + kk * 8, t1_mn + st_off + kb, idesc_acc, (i | kk) != 0);
        }
        tc_commit(&empty[st]);
        st = st_next;
        if (++st_next == kBwdStages) { st_next = 0; ph_next ^= 1; }
      }
      tc_commit(acc_done);
    }
  } else {
    // ------------------------------ softmax / epilogue warps ------------------------------
    const int quarter = warp & 3;
    const int wg = warp >> 2;        // 32-column half of every 64-column sub-tile handled by this warpgroup
    const int r_local = quarter * 32 + lane;
    const int row = r0 + r_local;
    const uint32_t lane_bits = static_cast<uint32_t>(quarter * 32) << 16;
    float l_row = 0.f, d_row = 0.f;
    if (!kDKV) {
      // rows >= s_q were written by the forward / delta kernels up to ld_stat (finite values)
      const int64_t off = static_cast<int64_t>(head) * p.ld_stat + (row < p.ld_stat ? row : 0);
      l_row = p.lse[off];
      d_row = p.delta[off];
    }
    int st = 0;
    uint32_t ph = 0;
    for (int i = 0; i < n_sub; ++i) {
      const int buf = i & 1;
      const uint32_t t_s = tmem_base + lane_bits + buf * 128;
      if (kDKV) mbar_wait(&full[st], ph);  // the L / Δ rows of this sub-tile have landed in smem
      mbar_wait(&s_full[buf], (i >> 1) & 1);
      tc_fence_after();
      const float* stat = smem_stat + st * 2 * kSub;
      {
        const int c = wg;              // this warpgroup's 32 score columns
        uint32_t s[32], dp[32], pk[16], dk[16];
        tmem_ld32(t_s + c * 32, s);
        tmem_ld32(t_s + 64 + c * 32, dp);
        tmem_ld_wait();
#pragma unroll
        for (int e = 0; e < 32; e += 4) {
          float lv[4], dv[4];
          if (kDKV) {
            const float4 l4 = *reinterpret_cast<const float4*>(stat + c * 32 + e);
            const float4 d4 = *reinterpret_cast<const float4*>(stat + kSub + c * 32 + e);
            lv[0] = l4.x; lv[1] = l4.y; lv[2] = l4.z; lv[3] = l4.w;
            dv[0] = d4.x; dv[1] = d4.y; dv[2] = d4.z; dv[3] = d4.w;
          } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) { lv[t] = l_row; dv[t] = d_row; }
          }
          float pv[4], ds[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            pv[t] = ex2f(fmaf(__uint_as_float(s[e + t]), p.scale_log2, -lv[t]));
            ds[t] = pv[t] * (__uint_as_float(dp[e + t]) - dv[t]);
          }
          pk[(e >> 1)] = pack_bf16(pv[0], pv[1]);
          pk[(e >> 1) + 1] = pack_bf16(pv[2], pv[3]);
          dk[(e >> 1)] = pack_bf16(ds[0], ds[1]);
          dk[(e >> 1) + 1] = pack_bf16(ds[2], ds[3]);
        }
        // P over the consumed score columns, dS over the consumed dP columns (16 packed columns per chunk)
        if (kDKV) {
          uint32_t (&lo)[8] = *reinterpret_cast<uint32_t (*)[8]>(&pk[0]);
          uint32_t (&hi)[8] = *reinterpret_cast<uint32_t (*)[8]>(&pk[8]);
          tmem_st8(t_s + c * 16, lo);
          tmem_st8(t_s + c * 16 + 8, hi);
        }
        {
          uint32_t (&lo)[8] = *reinterpret_cast<uint32_t (*)[8]>(&dk[0]);
          uint32_t (&hi)[8] = *reinterpret_cast<uint32_t (*)[8]>(&dk[8]);
          tmem_st8(t_s + 64 + c * 16, lo);
          tmem_st8(t_s + 64 + c * 16 + 8, hi);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[buf]);
      if (++st == kBwdStages) { st = 0; ph ^= 1; }
    }
    // ---- epilogue: accumulators -> bf16 -> global (one row per thread, 16-byte stores)
    mbar_wait(acc_done, 0);
    tc_fence_after();
    auto store_acc = [&](uint32_t tcol, __nv_bfloat16* out, int64_t ld, float mul) {
      __nv_bfloat16* orow = out + static_cast<int64_t>(row) * ld + head * 128;
#pragma unroll 1
      for (int c = 2 * wg; c < 2 * wg + 2; ++c) {   // each warpgroup stores 64 of the 128 output columns
        uint32_t a[32];
        tmem_ld32(tmem_base + lane_bits + tcol + c * 32, a);
        tmem_ld_wait();
        if (row < p.rows_res) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            uint4 o;
            o.x = pack_bf16(__uint_as_float(a[g * 8 + 0]) * mul, __uint_as_float(a[g * 8 + 1]) * mul);
            o.y = pack_bf16(__uint_as_float(a[g * 8 + 2]) * mul, __uint_as_float(a[g * 8 + 3]) * mul);
            o.z = pack_bf16(__uint_as_float(a[g * 8 + 4]) * mul, __uint_as_float(a[g * 8 + 5]) * mul);
            o.w = pack_bf16(__uint_as_float(a[g * 8 + 6]) * mul, __uint_as_float(a[g * 8 + 7]) * mul);
            *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = o;
          }
        }
      }
    };
    if (kDKV) store_acc(256, p.out0, p.ld0, 1.0f);
    store_acc(384, p.out1, p.ld1, p.scale);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kBwdMmaWarp) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// Δ[h][row] = sum_d dO[row, h, d] * O[row, h, d]; one warp per (row, head). Rows in [s_q, ld_stat) get 0.
__global__ void __launch_bounds__(256)
attn_delta_kernel(const __nv_bfloat16* __restrict__ o, int64_t ldo, const __nv_bfloat16* __restrict__ dout, int64_t ld_do,
                  float* __restrict__ delta, int64_t ld_stat, int s_q, int heads) {
  const int64_t gw = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t row = gw / heads;
  const int h = static_cast<int>(gw % heads);
  if (row >= ld_stat) return;
  float acc = 0.f;
  if (row < s_q) {
    const uint2 a = *reinterpret_cast<const uint2*>(o + row * ldo + h * 128 + lane * 4);
    const uint2 b = *reinterpret_cast<const uint2*>(dout + row * ld_do + h * 128 + lane * 4);
    acc = bf16_lo(a.x) * bf16_lo(b.x) + bf16_hi(a.x) * bf16_hi(b.x) + bf16_lo(a.y) * bf16_lo(b.y) + bf16_hi(a.y) * bf16_hi(b.y);
    acc = warp_sum(acc);
  }
  if (lane == 0) delta[static_cast<int64_t>(h) * ld_stat + row] = acc;
}

template <bool kDKV>
static int launch_bwd(dim3 grid, cudaStream_t stream, const CUtensorMap& r1, const CUtensorMap& r2, const CUtensorMap& t1,
                      const CUtensorMap& t2, const AttnBwdParams& p) {
  auto kfn = attn_bwd_kernel<kDKV>;
  static unsigned long long configured = 0;  // per template instance and device
  if (first_use_on_device(configured)) {
    FGB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmem));
  }
  kfn<<<grid, kBwdThreads, kBwdSmem, stream>>>(r1, r2, t1, t2, p);
  FGB_LAUNCH_CHECK("attn_bwd_kernel");
  return FGB_OK;
}

}  // namespace fgb

extern "C" int fgb_attn_bwd(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                            const void* o, int64_t ldo, const void* dout, int64_t ld_do, const void* lse, void* delta,
                            int64_t ld_stat, void* dq, int64_t ld_dq, void* dk, int64_t ld_dk, void* dv, int64_t ld_dv,
                            int32_t s_q, int32_t s_kv, int32_t heads, float scale, void* stream) {
  using namespace fgb;
  FGB_CHECK_ARG(ctx, "fgb_attn_bwd: ctx is NULL");
  FGB_CHECK_ARG(q && k && v && o && dout && lse && delta && dq && dk && dv, "fgb_attn_bwd: NULL tensor pointer");
  FGB_CHECK_ARG(s_q > 0 && s_kv > 0 && heads > 0 && heads <= 65535, "fgb_attn_bwd: bad problem s_q=%d s_kv=%d heads=%d", s_q, s_kv, heads);
  const int64_t width = static_cast<int64_t>(heads) * FGB_HEAD_DIM;
  FGB_CHECK_ARG(ldq >= width && ldk >= width && ldv >= width && ldo >= width && ld_do >= width && ld_dq >= width &&
                    ld_dk >= width && ld_dv >= width, "fgb_attn_bwd: leading dimension < heads*128");
  FGB_CHECK_ARG(ld_stat % kSub == 0 && ld_stat >= s_q, "fgb_attn_bwd: ld_stat=%lld must be a multiple of 64 and >= s_q", (long long)ld_stat);
  FGB_CHECK_ARG(aligned16(lse) && aligned16(delta) && aligned16(dq) && aligned16(dk) && aligned16(dv) && aligned16(o) && aligned16(dout),
                "fgb_attn_bwd: pointers must be 16-byte aligned");
  FGB_CHECK_ARG(ld_dq % 8 == 0 && ld_dk % 8 == 0 && ld_dv % 8 == 0 && ldo % 4 == 0 && ld_do % 4 == 0, "fgb_attn_bwd: leading dimensions must be multiples of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);

  {
    const int64_t warps = ld_stat * heads;
    attn_delta_kernel<<<static_cast<unsigned>((warps * 32 + 255) / 256), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(o), ldo, static_cast<const __nv_bfloat16*>(dout), ld_do, static_cast<float*>(delta),
        ld_stat, s_q, heads);
    FGB_LAUNCH_CHECK("attn_delta_kernel");
  }

  CUtensorMap q128, do128, k128, v128, q64, do64, k64, v64;
  int rc;
  if ((rc = make_tmap_bf16_2d(ctx, &q128, q, s_q, width, ldq, kRes))) return rc;
  if ((rc = make_tmap_bf16_2d(ctx, &do128, dout, s_q, width, ld_do, kRes))) return rc;
  if ((rc = make_tmap_bf16_2d(ctx, &k128, k, s_kv, width, ldk, kRes))) return rc;
  if ((rc = make_tmap_bf16_2d(ctx, &v128, v, s_kv, width, ldv, kRes))) return rc;
  if ((rc = make_tmap_bf16_2d(ctx, &q64, q, s_q, width, ldq, kSub))) return rc;
  if ((rc = make_tmap_bf16_2d(ctx, &do64, dout, s_q, width, ld_do, kSub))) return rc;
  if ((rc = make_tmap_bf16_2d(ctx, &k64, k, s_kv, width, ldk, kSub))) return rc;
  if ((rc = make_tmap_bf16_2d(ctx, &v64, v, s_kv, width, ldv, kSub))) return rc;

  AttnBwdParams p;
  p.lse = static_cast<const float*>(lse);
  p.delta = static_cast<const float*>(delta);
  p.ld_stat = ld_stat;
  p.scale = scale;
  p.scale_log2 = scale * 1.4426950408889634f;

  // dK, dV: resident = keys, streamed = queries
  p.out0 = static_cast<__nv_bfloat16*>(dv);
  p.ld0 = ld_dv;
  p.out1 = static_cast<__nv_bfloat16*>(dk);
  p.ld1 = ld_dk;
  p.rows_res = s_kv;
  p.rows_str = s_q;
  rc = launch_bwd<true>(dim3((s_kv + kRes - 1) / kRes, heads), st, k128, v128, q64, do64, p);
  if (rc) return rc;
  // dQ: resident = queries, streamed = keys
  p.out0 = nullptr;
  p.ld0 = 0;
  p.out1 = static_cast<__nv_bfloat16*>(dq);
  p.ld1 = ld_dq;
  p.rows_res = s_q;
  p.rows_str = s_kv;
  return launch_bwd<false>(dim3((s_q + kRes - 1) / kRes, heads), st, q128, do128, k64, v64, p);
}
