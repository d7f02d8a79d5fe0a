// fairygen_b200 — non-causal flash-attention forward on tcgen05 tensor cores (sm_100a), head_dim 128.
//
//   o[s, h, :] = softmax(q[s, h, :] · k[:, h, :]ᵀ · scale) · v[:, h, :]
//
// Replaces flash_attention()/AttentionModule of the reference (animation/diffsynth/models/
// wan_video_dit.py:27-60, 113-120) for self-attention (:145; S = 27 280 tokens at 704x1280x121) and
// cross-attention against the 512-token umT5 context (:179).
//
// Design: one CTA per (256 query rows, head); 320 threads.
//   warps 0-3 / 4-7  softmax warpgroup for query tile 0 / 1 (one query row per thread): tcgen05.ld
//                    the 128x128 fp32 score tile S from TMEM, online softmax with exp2 and a LAZY
//                    running-max (O is only rescaled when the max grows by more than 2^8), write
//                    P as packed bf16 back into TMEM over S (tcgen05.st), normalise O at the end.
//   warp 8           MMA issuer (one thread): S_i = Q_i·K_jᵀ (SS, both K-major) and
//                    O_i += P_i·V_j (A = P from TMEM, B = V from smem, MN-major), the two query
//                    tiles ping-pong so tensor cores work on one tile while the other is in softmax.
//   warp 9           TMA producer: Q once, then a 2-stage ring of K and V tiles (128 keys each).
// TMEM (512 columns): S0 | S1 | O0 | O1, 128 fp32 columns each; P_i aliases the first 64 columns of S_i.
#include "common.cuh"
#include "host.h"

namespace fgb {

constexpr int kAttnThreads = 320;
constexpr int kTile = 128;                  // query rows per tile == keys per KV tile == head_dim
constexpr int kBoxBytes = kTile * 64 * 2;   // one TMA box: 128 rows x 64 bf16 = 16 KB
constexpr int kTileBytes = 2 * kBoxBytes;   // 128 x 128 bf16 = 32 KB (two 64-column boxes)
constexpr int kKVStages = 2;
constexpr int kAttnSmem = 2 * kTileBytes /*Q*/ + 2 * kKVStages * kTileBytes /*K,V*/ + 1024 + 256;

struct AttnParams {
  __nv_bfloat16* o;
  int64_t ldo;
  int32_t s_q, s_kv;
  float scale_log2;
};

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                                   // [2 tiles][2 boxes][128][64]
  uint8_t* smem_k = smem + 2 * kTileBytes;                  // [stages][2 boxes][128][64]
  uint8_t* smem_v = smem_k + kKVStages * kTileBytes;        // [stages][2 boxes][128][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + kKVStages * kTileBytes);
  uint64_t* q_full = bars;          // [1]
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* v_full = bars + 3;      // [2]
  uint64_t* k_empty = bars + 5;     // [2]
  uint64_t* v_empty = bars + 7;     // [2]
  uint64_t* s_full = bars + 9;      // [2] per query tile: S_i written by the tensor core
  uint64_t* p_ready = bars + 11;    // [2] per query tile: P_i stored (and O_i rescaled) by 128 threads
  uint64_t* pv_done = bars + 13;    // [2] per query tile: O_i += P_i V_j finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int head = blockIdx.y;
  const int q0 = blockIdx.x * (2 * kTile);
  const int n_kv = (p.s_kv + kTile - 1) / kTile;

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&pv_done[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    if (lane == 0) {
      // ------------------------------- TMA producer -------------------------------
      mbar_expect_tx(q_full, 2 * kTileBytes);
      for (int i = 0; i < 2; ++i)
        for (int b = 0; b < 2; ++b)
          tma_load_2d(smem_q + i * kTileBytes + b * kBoxBytes, &tmap_q, q_full, head * 128 + b * 64,
                      q0 + i * kTile, kEvictFirst);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], kTileBytes);
        for (int b = 0; b < 2; ++b)
          tma_load_2d(smem_k + st * kTileBytes + b * kBoxBytes, &tmap_k, &k_full[st], head * 128 + b * 64,
                      j * kTile, kEvictLast);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], kTileBytes);
        for (int b = 0; b < 2; ++b)
          tma_load_2d(smem_v + st * kTileBytes + b * kBoxBytes, &tmap_v, &v_full[st], head * 128 + b * 64,
                      j * kTile, kEvictLast);
      }
    }
  } else if (warp == 8) {
    if (lane == 0) {
      // ------------------------------- MMA issuer ---------------------------------
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);  // P (TMEM)   x V (MN-major)
      auto issue_s = [&](int i, int st) {
        const uint32_t qa = smem_u32(smem_q + i * kTileBytes);
        const uint32_t ka = smem_u32(smem_k + st * kTileBytes);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // 16 head-dim elements per MMA
          const uint32_t off = (kk >> 2) * kBoxBytes + (kk & 3) * 32;
          umma_ss(tmem_base + i * 128, make_sdesc_sw128(qa + off, 16, 1024), make_sdesc_sw128(ka + off, 16, 1024),
                  idesc_s, kk != 0);
        }
        tc_commit(&s_full[i]);
      };
      auto issue_pv = [&](int i, int st, int j) {
        const uint32_t va = smem_u32(smem_v + st * kTileBytes);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // 16 keys per MMA: 16 rows of 128 B in each d-half box
          umma_ts(tmem_base + 256 + i * 128, tmem_base + i * 128 + kk * 8,
                  make_sdesc_sw128(va + kk * 2048, kBoxBytes, 1024), idesc_pv, (j | kk) != 0);
        }
        tc_commit(&pv_done[i]);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      issue_s(1, 0);
      tc_commit(&k_empty[0]);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1;
        const int st1 = (j + 1) & 1;
        mbar_wait(&v_full[st], (j >> 1) & 1);
        for (int i = 0; i < 2; ++i) {
          mbar_wait(&p_ready[i], j & 1);
          tc_fence_after();
          issue_pv(i, st, j);
          if (i == 1) tc_commit(&v_empty[st]);
          if (j + 1 < n_kv) {
            if (i == 0) {
              mbar_wait(&k_full[st1], ((j + 1) >> 1) & 1);
              tc_fence_after();
            }
            issue_s(i, st1);
            if (i == 1) tc_commit(&k_empty[st1]);
          }
        }
      }
    }
  } else {
    // ------------------------------ softmax warpgroups ------------------------------
    const int i = warp >> 2;        // query tile
    const int quarter = warp & 3;   // TMEM lane quarter
    const int row = q0 + i * kTile + quarter * 32 + lane;
    const uint32_t lane_bits = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_s = tmem_base + lane_bits + i * 128;
    const uint32_t t_o = tmem_base + lane_bits + 256 + i * 128;
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      mbar_wait(&s_full[i], j & 1);
      tc_fence_after();
      uint32_t sr[4][32];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(t_s + c * 32, sr[c]);
      tmem_ld_wait();
      const int valid = p.s_kv - j * kTile;  // keys of this tile that exist
      if (valid < kTile) {
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c * 32 + e >= valid) sr[c][e] = 0xff800000u;  // -inf
      }
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int e = 0; e < 32; ++e) mx = fmaxf(mx, __uint_as_float(sr[c][e]));
      const float m_tile = mx * p.scale_log2;
      if (j == 0) {
        m = m_tile;
      } else {
        const float m_new = fmaxf(m, m_tile);
        if (__any_sync(0xffffffffu, m_new - m > 8.0f)) {
          // rescale the running sum and the O accumulator of this row by 2^(m - m_new)
          const float alpha = fast_exp2(m - m_new);
          l *= alpha;
          m = m_new;
          mbar_wait(&pv_done[i], (j - 1) & 1);  // O_i must be quiescent
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t orr[32];
            tmem_ld32(t_o + c * 32, orr);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) orr[e] = __float_as_uint(__uint_as_float(orr[e]) * alpha);
            tmem_st32(t_o + c * 32, orr);
          }
        }
      }
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {  // 64 keys -> 32 packed columns per store
        uint32_t pk[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int src = c * 64 + 2 * e;
          const float p0 = fast_exp2(fmaf(__uint_as_float(sr[src >> 5][src & 31]), p.scale_log2, -m));
          const float p1 = fast_exp2(fmaf(__uint_as_float(sr[(src + 1) >> 5][(src + 1) & 31]), p.scale_log2, -m));
          sum += p0 + p1;
          pk[e] = pack_bf16(p0, p1);
        }
        tmem_st32(t_s + c * 32, pk);
      }
      l += sum;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[i]);
    }
    // final normalisation: O / l -> bf16 -> global
    mbar_wait(&pv_done[i], (n_kv - 1) & 1);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    __nv_bfloat16* orow = p.o + static_cast<int64_t>(row) * p.ldo + head * 128;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t orr[32];
      tmem_ld32(t_o + c * 32, orr);
      tmem_ld_wait();
      if (row < p.s_q) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 o;
          o.x = pack_bf16(__uint_as_float(orr[g * 8 + 0]) * inv_l, __uint_as_float(orr[g * 8 + 1]) * inv_l);
          o.y = pack_bf16(__uint_as_float(orr[g * 8 + 2]) * inv_l, __uint_as_float(orr[g * 8 + 3]) * inv_l);
          o.z = pack_bf16(__uint_as_float(orr[g * 8 + 4]) * inv_l, __uint_as_float(orr[g * 8 + 5]) * inv_l);
          o.w = pack_bf16(__uint_as_float(orr[g * 8 + 6]) * inv_l, __uint_as_float(orr[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = o;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

}  // namespace fgb

extern "C" int fgb_attn_fwd(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                            int64_t ldv, void* o, int64_t ldo, int32_t s_q, int32_t s_kv, int32_t heads, float scale,
                            void* stream) {
  using namespace fgb;
  FGB_CHECK_ARG(ctx, "fgb_attn_fwd: ctx is NULL");
  FGB_CHECK_ARG(q && k && v && o, "fgb_attn_fwd: NULL tensor pointer");
  FGB_CHECK_ARG(s_q > 0 && s_kv > 0 && heads > 0, "fgb_attn_fwd: empty problem s_q=%d s_kv=%d heads=%d", s_q, s_kv, heads);
  FGB_CHECK_ARG(heads <= 65535, "fgb_attn_fwd: too many heads");
  const int64_t width = static_cast<int64_t>(heads) * FGB_HEAD_DIM;
  FGB_CHECK_ARG(ldq >= width && ldk >= width && ldv >= width && ldo >= width, "fgb_attn_fwd: leading dimension < heads*128");
  FGB_CHECK_ARG(aligned16(o) && ldo % 8 == 0, "fgb_attn_fwd: o must be 16-byte aligned with ldo %% 8 == 0");

  CUtensorMap tq, tk, tv;
  int rc = make_tmap_bf16_2d(ctx, &tq, q, s_q, width, ldq, kTile);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(ctx, &tk, k, s_kv, width, ldk, kTile);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(ctx, &tv, v, s_kv, width, ldv, kTile);
  if (rc) return rc;

  AttnParams p;
  p.o = static_cast<__nv_bfloat16*>(o);
  p.ldo = ldo;
  p.s_q = s_q;
  p.s_kv = s_kv;
  p.scale_log2 = scale * 1.4426950408889634f;

  static bool configured = false;
  if (!configured) {
    FGB_CUDA(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
    configured = true;
  }
  dim3 grid((s_q + 2 * kTile - 1) / (2 * kTile), heads);
  attn_fwd_kernel<<<grid, kAttnThreads, kAttnSmem, static_cast<cudaStream_t>(stream)>>>(tq, tk, tv, p);
  FGB_LAUNCH_CHECK("attn_fwd_kernel");
  return FGB_OK;
}
