// fairygen_b200 — flash-attention forward, cluster variant: ONE 128-row query tile per CTA, the score tile S
// DOUBLE-BUFFERED in TMEM, K/V tiles multicast across a 2-CTA cluster.
//
// Why: in attention.cu (two query tiles per CTA, TMEM = 2 x (S + O)) the tensor pipe of a tile has to wait for that
// tile's softmax (S_i(j+1) overwrites P_i(j)), so every barrier / TMEM / commit latency of the loop
// softmax -> PV -> S -> softmax is exposed with only two tiles in flight. Here TMEM holds S[0] | S[1] | O (384 columns):
// S(j+1) = Q·K(j+1)ᵀ is issued BEFORE the softmax of tile j has finished, so the softmax warps always find their next
// score tile ready and the tensor pipe only ever waits for P(j). With one tile per CTA every CTA would fetch all of K
// and V on its own (64 B/clk/SM from L2 at full tensor rate, above what the L2 can deliver to 148 SMs), so two CTAs
// (adjacent query tiles of one head) form a cluster: each loads HALF of every K/V tile and multicasts it into both
// CTAs' shared memory; a ring slot is recycled when BOTH CTAs' MMAs have consumed it (multicast tcgen05.commit).
//
// Warp roles (320 threads) as in attention.cu: warps 0-3 / 4-7 softmax for key columns [0,64) / [64,128) of the tile
// (lazy running max, row max swapped through smem, partial exp2 emulation), warp 8 MMA issuer, warp 9 TMA producer.
#include <cstdlib>

#include "attention_common.cuh"

namespace fgb {

constexpr int kCStages = 2;
constexpr int kHalfBox = 64 * 64 * 2;   // TMA box of the multicast loads: 64 rows x 64 bf16 = 8 KB
constexpr int kClusterSmem = kTileBytes /*Q*/ + 2 * kCStages * kTileBytes /*K,V*/ + 1024 /*align*/ + 256 /*barriers*/ +
                             2 * 2 * kTile * 4 /*row-max exchange*/;

__device__ __forceinline__ void tma_load_2d_mc(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int32_t c0, int32_t c1,
                                               uint16_t cta_mask, uint64_t hint) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5, %6;"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask),
      "l"(hint)
      : "memory");
}
// arrive (once all prior tcgen05 ops of this thread are done) on the same-offset barrier of every CTA in cta_mask
__device__ __forceinline__ void tc_commit_mc(uint64_t* bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"(cta_mask)
               : "memory");
}

template <int EMU>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kAttnThreads, 1)
attn_fwd_cluster_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                        const __grid_constant__ CUtensorMap tmap_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_q = smem;                                   // [2 boxes][128][64]
  uint8_t* smem_k = smem + kTileBytes;                      // [stages][2 boxes][128][64]
  uint8_t* smem_v = smem_k + kCStages * kTileBytes;         // [stages][2 boxes][128][64]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + kCStages * kTileBytes);
  uint64_t* q_full = bars;                     // [1]
  uint64_t* k_full = bars + 1;                 // [stages]  both CTAs' halves have landed
  uint64_t* v_full = k_full + kCStages;        // [stages]
  uint64_t* k_empty = v_full + kCStages;       // [stages]  count 2: this CTA's and the peer's MMAs are done with the slot
  uint64_t* v_empty = k_empty + kCStages;      // [stages]
  uint64_t* s_full = v_empty + kCStages;       // [2] score buffer written by the tensor core
  uint64_t* p_ready = s_full + 2;              // [1] P stored (and O rescaled) by 256 threads
  uint64_t* pv_done = p_ready + 1;             // [1] O += P V finished
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);
  float* xchg = reinterpret_cast<float*>(bars + 32);   // [2 slots][2 column halves][128 rows]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t cta = cluster_ctarank();
  const int n_kv_all = (p.s_kv + kTile - 1) / kTile;
  int unit = blockIdx.x >> 1, kv_lo = 0, kv_hi = n_kv_all, part = -1;
  if (unit >= p.n_full) {
    part = unit - p.n_full;
    const int chunk = part % p.split;
    unit = p.n_full + part / p.split;
    kv_lo = static_cast<int>(static_cast<int64_t>(chunk) * n_kv_all / p.split);
    kv_hi = static_cast<int>(static_cast<int64_t>(chunk + 1) * n_kv_all / p.split);
  }
  const int head = unit / p.n_pairs;
  const int q0 = (unit % p.n_pairs) * (2 * kTile) + static_cast<int>(cta) * kTile;
  const int n_kv = kv_hi - kv_lo;

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    mbar_init(q_full, 1);
    for (int i = 0; i < kCStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 2);
      mbar_init(&v_empty[i], 2);
    }
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_ready, 256);
    mbar_init(pv_done, 1);
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the peer's barriers exist before anything of ours can signal them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;   // columns: S[0] 0-127 | S[1] 128-255 | O 256-383

  if (warp == 9) {
    if (elect_one()) {
      // ------------------------------- TMA producer -------------------------------
      mbar_expect_tx(q_full, kTileBytes);
      for (int b = 0; b < 2; ++b) tma_load_2d(smem_q + b * kBoxBytes, &tmap_q, q_full, head * 128 + b * 64, q0, kEvictFirst);
      int st = 0;
      uint32_t ph = 0;
      for (int j = 0; j < n_kv; ++j) {
        // this CTA fetches rows [64*cta, 64*cta+64) of the tile and multicasts them into both CTAs
        mbar_wait_cluster(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], kTileBytes);
        for (int b = 0; b < 2; ++b)
          tma_load_2d_mc(smem_k + st * kTileBytes + b * kBoxBytes + cta * kHalfBox, &tmap_k, &k_full[st], head * 128 + b * 64,
                         (kv_lo + j) * kTile + static_cast<int>(cta) * 64, 0x3, kEvictLast);
        mbar_wait_cluster(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], kTileBytes);
        for (int b = 0; b < 2; ++b)
          tma_load_2d_mc(smem_v + st * kTileBytes + b * kBoxBytes + cta * kHalfBox, &tmap_v, &v_full[st], head * 128 + b * 64,
                         (kv_lo + j) * kTile + static_cast<int>(cta) * 64, 0x3, kEvictLast);
        if (++st == kCStages) { st = 0; ph ^= 1; }
      }
    }
  } else if (warp == 8) {
    if (elect_one()) {
      // ------------------------------- MMA issuer ---------------------------------
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_pv = make_idesc_bf16(128, 128, 0, 1);  // P (TMEM)   x V (MN-major)
      const uint64_t q_desc = make_sdesc_sw128(smem_u32(smem_q), 16, 1024);
      const uint64_t k_desc = make_sdesc_sw128(smem_u32(smem_k), 16, 1024);
      const uint64_t v_desc = make_sdesc_sw128(smem_u32(smem_v), kBoxBytes, 1024);
      auto issue_s = [&](int buf, int st) {
        const uint64_t kd = k_desc + static_cast<uint64_t>((st * kTileBytes) >> 4);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // 16 head-dim elements per MMA
          const uint64_t off = static_cast<uint64_t>(((kk >> 2) * kBoxBytes + (kk & 3) * 32) >> 4);
          umma_ss(tmem_base + buf * 128, q_desc + off, kd + off, idesc_s, kk != 0);
        }
        tc_commit(&s_full[buf]);
        tc_commit_mc(&k_empty[st], 0x3);
      };
      mbar_wait(q_full, 0);
      mbar_wait_cluster(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      int st = 0, st_next = 1 % kCStages;
      uint32_t ph = 0, ph_next = (kCStages == 1) ? 1u : 0u;
      for (int j = 0; j < n_kv; ++j) {
        const int buf = j & 1;
        if (j + 1 < n_kv) {
          // S(j+1) goes into the other score buffer: its previous contents (P of tile j-1) were consumed by PV(j-1),
          // issued in the previous iteration — the tensor pipe executes in order.
          mbar_wait_cluster(&k_full[st_next], ph_next);
          tc_fence_after();
          issue_s(buf ^ 1, st_next);
        }
        mbar_wait_cluster(&v_full[st], ph);
        mbar_wait(p_ready, j & 1);
        tc_fence_after();
        const uint64_t vd = v_desc + static_cast<uint64_t>((st * kTileBytes) >> 4);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // 16 keys per MMA; P lives in columns [0,32) and [64,96) of the score buffer
          umma_ts(tmem_base + 256, tmem_base + buf * 128 + (kk & 3) * 8 + (kk >> 2) * 64, vd + static_cast<uint64_t>((kk * 2048) >> 4),
                  idesc_pv, (j | kk) != 0);
        tc_commit(pv_done);
        tc_commit_mc(&v_empty[st], 0x3);
        st = st_next;
        ph = ph_next;
        if (++st_next == kCStages) { st_next = 0; ph_next ^= 1; }
      }
    }
  } else {
    // ------------------------------ softmax warpgroups ------------------------------
    const int wg = warp >> 2;       // key-column half handled by this warpgroup
    const int quarter = warp & 3;   // TMEM lane quarter
    const int r_local = quarter * 32 + lane;
    const int row = q0 + r_local;
    const uint32_t lane_bits = static_cast<uint32_t>(quarter * 32) << 16;
    const uint32_t t_o = tmem_base + lane_bits + 256;
    float m = -INFINITY, l = 0.f;
    int xcount = 0;
    const uint64_t scale2 = pack2(p.scale_log2, p.scale_log2);

    auto swap_rows = [&](float mine) -> float {
      float* slot = xchg + (xcount & 1) * (2 * kTile);
      ++xcount;
      slot[wg * kTile + r_local] = mine;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
      return slot[(wg ^ 1) * kTile + r_local];
    };

    auto sweep = [&](auto exps_tag, uint32_t t_s, float m_used, float& row_max, uint32_t (&pk)[32]) -> float {
      constexpr bool EXPS = decltype(exps_tag)::value;
      uint32_t buf[2][32];
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
      const uint64_t negm2 = pack2(-m_used, -m_used);
      tmem_ld32(t_s, buf[0]);
      tmem_ld32(t_s + 32, buf[1]);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t (&cur)[32] = buf[c];
#pragma unroll
        for (int e = 0; e < 32; e += 2)
          mx[(e >> 1) & 3] = fmaxf(mx[(e >> 1) & 3], fmaxf(__uint_as_float(cur[e]), __uint_as_float(cur[e + 1])));
        if (EXPS) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const uint64_t x2 = fma2(pack2(__uint_as_float(cur[2 * e]), __uint_as_float(cur[2 * e + 1])), scale2, negm2);
            float p0, p1;
            if ((e & 7) < EMU) {
              exp2_emulated(x2, p0, p1);
            } else {
              float x0, x1;
              unpack2(x2, x0, x1);
              p0 = fast_exp2(x0);
              p1 = fast_exp2(x1);
            }
            acc[e & 3] = add2(acc[e & 3], pack2(p0, p1));
            pk[c * 16 + e] = pack_bf16(p0, p1);
          }
        }
      }
      row_max = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      float a0, a1, b0, b1;
      unpack2(add2(acc[0], acc[1]), a0, a1);
      unpack2(add2(acc[2], acc[3]), b0, b1);
      return (a0 + a1) + (b0 + b1);
    };

    for (int j = 0; j < n_kv; ++j) {
      const int sb = j & 1;
      const uint32_t t_s = tmem_base + lane_bits + sb * 128 + wg * 64;  // this thread's 64 score columns
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      const int valid = p.s_kv - (kv_lo + j) * kTile - wg * 64;  // keys of this half-tile that exist
      if (valid < 64) {
#pragma unroll 1
        for (int c = (valid < 0 ? 0 : valid) >> 5; c < 2; ++c) {
          uint32_t fix[32];
          tmem_ld32(t_s + c * 32, fix);
          tmem_ld_wait();
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c * 32 + e >= valid) fix[e] = 0xff800000u;
          tmem_st32(t_s + c * 32, fix);
        }
        tmem_st_wait();
      }
      uint32_t pk[32];
      float row_max, sum;
      bool redo = false;
      if (j == 0) {
        sweep(TagFalse{}, t_s, 0.f, row_max, pk);  // exact maximum of the first tile
        row_max = fmaxf(row_max, swap_rows(row_max));
        m = row_max * p.scale_log2;
        redo = true;
      } else {
        sum = sweep(TagTrue{}, t_s, m, row_max, pk);     // speculative: exponentials against the stale maximum
        row_max = fmaxf(row_max, swap_rows(row_max));   // maximum of the whole 128-key row
        const float m_new = fmaxf(m, row_max * p.scale_log2);
        if (__any_sync(0xffffffffu, m_new - m > 8.0f)) {
          // rare: rescale the running sum and this warpgroup's half of O by 2^(m - m_new), then redo the tile
          const float alpha = fast_exp2(m - m_new);
          l *= alpha;
          m = m_new;
          mbar_wait(pv_done, (j - 1) & 1);  // O must be quiescent
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t orr[32];
            tmem_ld32(t_o + wg * 64 + c * 32, orr);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 32; ++e) orr[e] = __float_as_uint(__uint_as_float(orr[e]) * alpha);
            tmem_st32(t_o + wg * 64 + c * 32, orr);
          }
          redo = true;
        }
      }
      if (redo) sum = sweep(TagTrue{}, t_s, m, row_max, pk);
      l += sum;
      tmem_st32(t_s, pk);   // P over this thread's own first 32 (consumed) score columns
      // Flow control: the score tile j+1 is ready long before P(j) is consumed, so without this wait the 256 arrivals
      // of tile j+1 could complete a second phase of p_ready before the MMA thread has observed the first.
      if (j > 0) mbar_wait(pv_done, (j - 1) & 1);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);
    }

    // ---- epilogue: each warpgroup normalises and stores its 64 of the 128 output columns
    const float l_row = l + swap_rows(l);
    mbar_wait(pv_done, (n_kv - 1) & 1);
    tc_fence_after();
    if (part >= 0) {
      const int pr = static_cast<int>(cta) * kTile + r_local;
      float* prow = p.part_o + (static_cast<int64_t>(part) * (2 * kTile) + pr) * 128 + wg * 64;
      if (wg == 0) p.part_ml[static_cast<int64_t>(part) * (2 * kTile) + pr] = make_float2(m, l_row);
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t orr[32];
        tmem_ld32(t_o + wg * 64 + c * 32, orr);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 8; ++g)
          *reinterpret_cast<uint4*>(prow + c * 32 + g * 4) = make_uint4(orr[g * 4], orr[g * 4 + 1], orr[g * 4 + 2], orr[g * 4 + 3]);
      }
    } else {
      const float inv_l = 1.0f / l_row;
      if (wg == 0 && p.lse != nullptr && row < p.ld_lse) p.lse[static_cast<int64_t>(head) * p.ld_lse + row] = m + __log2f(l_row);
      __nv_bfloat16* orow = row < p.s_q ? out_row(p, row, head) + wg * 64 : nullptr;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t orr[32];
        tmem_ld32(t_o + wg * 64 + c * 32, orr);
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
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
  cluster_sync_all();   // do not exit while the peer can still multicast into this CTA's smem / barriers
}

template <int EMU>
static int launch_one(int grid, cudaStream_t stream, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                      const AttnParams& p) {
  auto kfn = attn_fwd_cluster_kernel<EMU>;
  static bool configured = false;
  if (!configured) {
    FGB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kClusterSmem));
    configured = true;
  }
  kfn<<<grid, kAttnThreads, kClusterSmem, stream>>>(tq, tk, tv, p);
  FGB_LAUNCH_CHECK("attn_fwd_cluster_kernel");
  return FGB_OK;
}

int launch_attn_cluster(int emu, int units, cudaStream_t stream, const CUtensorMap& tq, const CUtensorMap& tk64,
                        const CUtensorMap& tv64, const AttnParams& p) {
  const int grid = 2 * units;   // one 2-CTA cluster per work unit
  switch (emu) {
    case 0: return launch_one<0>(grid, stream, tq, tk64, tv64, p);
    case 1: return launch_one<1>(grid, stream, tq, tk64, tv64, p);
    case 2: return launch_one<2>(grid, stream, tq, tk64, tv64, p);
    case 4: return launch_one<4>(grid, stream, tq, tk64, tv64, p);
    default: return launch_one<3>(grid, stream, tq, tk64, tv64, p);
  }
}

}  // namespace fgb
