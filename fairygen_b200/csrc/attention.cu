// fairygen_b200 — non-causal flash-attention forward on tcgen05 tensor cores (sm_100a), head_dim 128.
//
//   o[s, h, :] = softmax(q[s, h, :] · k[:, h, :]ᵀ · scale) · v[:, h, :]
//
// Replaces flash_attention()/AttentionModule of the reference (animation/diffsynth/models/
// wan_video_dit.py:27-60, 113-120) for self-attention (:145; S = 27 280 tokens at 704x1280x121) and
// cross-attention against the 512-token umT5 context (:179).
//
// Design: one CTA per (256 query rows = two 128-row tiles, head); 320 threads.
//   warp 8           MMA issuer (one thread): S_i = Q_i·K_jᵀ (SS, both K-major) and O_i += P_i·V_j (A = P from TMEM,
//                    B = V from smem, MN-major); the two query tiles alternate in the tensor pipe.
//   warp 9           TMA producer: Q once, then a 2-stage ring of K and V tiles (128 keys each).
//   warps 0-3 / 4-7  softmax warpgroup for key columns [0,64) / [64,128) of EVERY score tile, one query row per
//                    thread: both warpgroups work on the same tile at the same time and then move to the other
//                    tile together. The MUFU (16 ex2/clk/SM) needs as long for a 128x128 tile as the tensor pipe
//                    needs for its two MMAs, so the softmax must never stall: with two warps per SM sub-partition
//                    active on one tile their dependency / MUFU-queue stalls hide each other, and while they
//                    process tile 1 the tensor pipe finishes PV and the next S of tile 0 (a single warp per
//                    sub-partition, as in a one-warpgroup-per-tile split, issued only every ~4th cycle).
//                    Online softmax with exp2 against a LAZY running max: exponentials are computed speculatively
//                    against the stale max while the half-row maximum is found on the side; the two threads that
//                    share a row swap their maxima through smem (one 64-thread named barrier per tile), and only if
//                    the max grew by more than 2^8 are O and l rescaled and the tile redone. Each thread keeps the
//                    row sum of its own columns; P goes back into TMEM over the thread's own consumed score
//                    columns as packed bf16 (tcgen05.st); O is normalised at the end.
// TMEM (512 columns): S0 | S1 | O0 | O1, 128 fp32 columns each; P_i aliases columns [0,32) and [64,96) of S_i.
#include <cstdlib>
#include <vector>

#include "attention_common.cuh"

namespace fgb {

#ifndef FGB_ATTN_TRACE
#define FGB_ATTN_TRACE 0
#endif
#if FGB_ATTN_TRACE
// Debug build only (make EXTRA=-DFGB_ATTN_TRACE=1): softmax thread 0 of CTA 0 stamps the phases of its first items with the
// global nanosecond timer; tools/kcheck (KCHECK_TRACE=1) prints them through fgb_debug_attn_trace.
__device__ unsigned long long g_attn_trace[16 * 16];
__device__ __forceinline__ void trace_stamp(uint32_t item, int slot) {
  if (blockIdx.x == 0 && threadIdx.x == 0 && item < 16) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_attn_trace[item * 16 + slot] = t;
  }
}
#define FGB_TRACE(item, slot) trace_stamp(item, slot)
#else
#define FGB_TRACE(item, slot)
#endif

// EMU: of every 8 score pairs, EMU are exponentiated by exp2_emulated, the rest by MUFU.EX2.
// PAIR: two CTAs of a cluster (one TPC) work on 512 query rows of one head with tcgen05.mma.cta_group::2: every score / PV
// MMA has M = 256 (tile i of both CTAs), each CTA stages only HALF of every K tile (64 of its 128 keys) and half of every V
// tile (64 of its 128 head-dim columns) — the shared-memory fill and the operand reads of the tensor core per SM drop by a
// third, the L2 -> SM traffic for K/V by half. The leader CTA issues all MMAs; K/V `full` barriers live in the leader (2-CTA
// TMA signals them), `empty` / s_full / pv_done are multicast commits, P-ready collects one arrival per softmax warp of both
// CTAs. Everything a softmax thread does is CTA-local and identical in both variants.
template <int EMU, bool PAIR>
__global__ void __launch_bounds__(kAttnThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const __grid_constant__ CUtensorMap tmap_o, const AttnParams p) {
  constexpr int kEmu = EMU;
  constexpr int kFrac = (EMU == 9) ? 3 : EMU;   // pairs of every 8 that take the FMA-pipe exp2
  constexpr int kStages = PAIR ? 4 : 2;                       // K / V ring depth
  constexpr int kKVBytes = PAIR ? kTileBytes / 2 : kTileBytes;   // bytes of one K (or V) stage in THIS CTA
  constexpr int kKBox = PAIR ? kBoxBytes / 2 : kBoxBytes;        // one K box: 64 (pair) or 128 key rows of 128 B
  constexpr int kItemTiles = PAIR ? 4 : 2;                       // 128-row query tiles per work item
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  if (smem - smem_raw > kAttnAlignSlack) __trap();          // the slack left below the 227 KB limit must cover the round-up
  uint8_t* smem_q = smem;                                   // [2 tiles][2 boxes][128][64]
  uint8_t* smem_k = smem + 2 * kTileBytes;                  // [stages][2 boxes][128 or 64][64]
  uint8_t* smem_v = smem_k + kStages * kKVBytes;            // [stages][2 boxes][128][64] or [stages][128][64]
  uint8_t* smem_o = smem_v + kStages * kKVBytes;            // [2 boxes][128][64]: one finished output tile on its way to a TMA store
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_o + kTileBytes);
  uint64_t* q_full = bars;          // [1] this CTA's Q tiles landed
  uint64_t* q_empty = bars + 1;     // [1] the item's last score MMAs have read Q: the next item's Q may land
  uint64_t* q_pair = bars + 2;      // [1] PAIR, leader: the peer's Q tiles landed (remote arrive)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);
  //                         bars + 4, 5: four vote words
  uint64_t* s_full = bars + 6;      // [2] per query tile: S_i written by the tensor core
  uint64_t* p_ready = bars + 8;     // [2] per query tile: first half of every thread's P_i (32 of its 64 keys) stored
  uint64_t* p_ready2 = bars + 10;   // [2] per query tile: second half stored (and O_i rescaled) — PV starts on the first half
  uint64_t* pv_done = bars + 12;    // [2] per query tile: O_i += P_i V_j finished
  uint64_t* k_full = bars + 14;     // [4]
  uint64_t* v_full = bars + 18;     // [4]
  uint64_t* k_empty = bars + 22;    // [4]
  uint64_t* v_empty = bars + 26;    // [4]
  float* xchg = reinterpret_cast<float*>(bars + 32);   // [2 slots][2 column halves][128 rows]

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kv_all = (p.s_kv + kTile - 1) / kTile;
  // Persistent CTA: work items blockIdx.x, blockIdx.x + gridDim.x, ... of the list [n_full whole units | split units x split
  // key chunks]. Every role walks the same list; mbarrier parities follow running counters (items, KV tiles), so the TMA
  // thread prefetches the next item's Q / K / V and the tensor core starts its first score tiles while the softmax warps
  // are still writing the previous item's output: prologue and epilogue of consecutive items overlap (they dominated the
  // 4-tile cross-attention launches, where a CTA used to live for 8 tile steps only).
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;   // 0 = leader
  int head = 0, q0 = 0, kv_lo = 0, n_kv = 0, part = -1;
  auto decode_item = [&](int w) {
    int unit = w, kv_hi = n_kv_all;
    kv_lo = 0;
    part = -1;
    if (unit >= p.n_full) {
      part = unit - p.n_full;
      const int chunk = part % p.split;
      unit = p.n_full + part / p.split;
      kv_lo = static_cast<int>(static_cast<int64_t>(chunk) * n_kv_all / p.split);
      kv_hi = static_cast<int>(static_cast<int64_t>(chunk + 1) * n_kv_all / p.split);
    }
    head = unit / p.n_pairs;
    q0 = (unit % p.n_pairs) * (kItemTiles * kTile) + static_cast<int>(rank) * (2 * kTile);   // first query row of THIS CTA
    n_kv = kv_hi - kv_lo;   // KV tiles of this item; local tile jj is global tile kv_lo + jj
  };
  const int w0 = PAIR ? (blockIdx.x >> 1) : blockIdx.x;        // first work item of this CTA (pair) and the stride
  const int w_stride = PAIR ? (gridDim.x >> 1) : gridDim.x;

  if (warp == 9 && lane == 0) {
    tma_prefetch_desc(&tmap_q);
    tma_prefetch_desc(&tmap_k);
    tma_prefetch_desc(&tmap_v);
    tma_prefetch_desc(&tmap_o);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(q_pair, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], PAIR ? 16 : 256);    // PAIR: one arrival per softmax warp of both CTAs
      mbar_init(&p_ready2[i], PAIR ? 16 : 256);
      mbar_init(&pv_done[i], 1);
    }
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 8) {
    if (PAIR) tmem_alloc_2sm<512>(tmem_slot);
    else tmem_alloc<512>(tmem_slot);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();    // both CTAs' barriers exist before any remote signal
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 9) {
    if (elect_one()) {
      // ------------------------------- TMA producer -------------------------------
      uint32_t kf[kStages], vf[kStages];   // PAIR: the leader's full barriers as cluster addresses
      if (PAIR) {
#pragma unroll
        for (int sI = 0; sI < kStages; ++sI) {
          asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(kf[sI]) : "r"(smem_u32(&k_full[sI])));
          asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(vf[sI]) : "r"(smem_u32(&v_full[sI])));
        }
      }
      uint32_t t = 0;    // KV tiles loaded so far (all items): ring stage t % kStages, phase (t / kStages) & 1
      uint32_t it = 0;   // items so far
      for (int w = w0; w < p.n_items; w += w_stride, ++it) {
        decode_item(w);
        if (PAIR) mbar_wait_parked_cluster(q_empty, (it & 1) ^ 1);
        else mbar_wait_parked(q_empty, (it & 1) ^ 1);
        mbar_expect_tx(q_full, 2 * kTileBytes);
        for (int i = 0; i < 2; ++i)
          for (int b = 0; b < 2; ++b)
            tma_load_2d(smem_q + i * kTileBytes + b * kBoxBytes, &tmap_q, q_full, head * 128 + b * 64,
                        q0 + i * kTile, kEvictFirst);
        for (int j = 0; j < n_kv; ++j, ++t) {
          const int st = t % kStages;
          const uint32_t ph = (t / kStages) & 1;
          if (PAIR) {
            // this CTA's half of the tile: keys [rank*64, +64) of K (both head-dim boxes), head-dim columns [rank*64, +64) of V
            mbar_wait_parked_cluster(&k_empty[st], ph ^ 1);
            if (rank == 0) mbar_expect_tx(&k_full[st], 2 * kKVBytes);
            uint32_t kfl = kf[0], vfl = vf[0];
#pragma unroll
            for (int sI = 1; sI < kStages; ++sI) {
              kfl = (st == sI) ? kf[sI] : kfl;
              vfl = (st == sI) ? vf[sI] : vfl;
            }
            for (int b = 0; b < 2; ++b)
              tma_load_2d_2sm(smem_k + st * kKVBytes + b * kKBox, &tmap_k, kfl, head * 128 + b * 64,
                              (kv_lo + j) * kTile + static_cast<int>(rank) * 64, kEvictLast);
            mbar_wait_parked_cluster(&v_empty[st], ph ^ 1);
            if (rank == 0) mbar_expect_tx(&v_full[st], 2 * kKVBytes);
            tma_load_2d_2sm(smem_v + st * kKVBytes, &tmap_v, vfl, head * 128 + static_cast<int>(rank) * 64, (kv_lo + j) * kTile,
                            kEvictLast);
          } else {
            mbar_wait_parked(&k_empty[st], ph ^ 1);
            mbar_expect_tx(&k_full[st], kTileBytes);
            for (int b = 0; b < 2; ++b)
              tma_load_2d(smem_k + st * kTileBytes + b * kBoxBytes, &tmap_k, &k_full[st], head * 128 + b * 64,
                          (kv_lo + j) * kTile, kEvictLast);
            mbar_wait_parked(&v_empty[st], ph ^ 1);
            mbar_expect_tx(&v_full[st], kTileBytes);
            for (int b = 0; b < 2; ++b)
              tma_load_2d(smem_v + st * kTileBytes + b * kBoxBytes, &tmap_v, &v_full[st], head * 128 + b * 64,
                          (kv_lo + j) * kTile, kEvictLast);
          }
        }
      }
    }
  } else if (warp == 8) {
    if (rank == 0 && elect_one()) {
      // ------------------------------- MMA issuer (PAIR: the leader, for both CTAs) ---------------------------------
      constexpr uint32_t idesc_s = make_idesc_bf16(PAIR ? 256 : 128, 128, 0, 0);   // Q (K-major) x K (K-major)
      constexpr uint32_t idesc_pv = make_idesc_bf16(PAIR ? 256 : 128, 128, 0, 1);  // P (TMEM)   x V (MN-major)
      // Descriptors are built once; per MMA only a compile-time offset is added to the 14-bit address field.
      const uint64_t q_desc = make_sdesc_sw128(smem_u32(smem_q), 16, 1024);
      const uint64_t k_desc = make_sdesc_sw128(smem_u32(smem_k), 16, 1024);
      const uint64_t v_desc = make_sdesc_sw128(smem_u32(smem_v), kBoxBytes, 1024);
      auto commit = [&](uint64_t* bar) {
        if (PAIR) tc_commit_2sm(bar, 0x3);   // same-offset barrier in both CTAs
        else tc_commit(bar);
      };
      auto wait = [&](uint64_t* bar, uint32_t parity) {
        if (PAIR) mbar_wait_parked_cluster(bar, parity);
        else mbar_wait_parked(bar, parity);
      };
      auto issue_s = [&](int i, int st) {
        const uint64_t qd = q_desc + static_cast<uint64_t>((i * kTileBytes) >> 4);
        const uint64_t kd = k_desc + static_cast<uint64_t>((st * kKVBytes) >> 4);
        const uint32_t d_tmem = tmem_base + i * 128;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {  // 16 head-dim elements per MMA
          const uint64_t off_q = static_cast<uint64_t>(((kk >> 2) * kBoxBytes + (kk & 3) * 32) >> 4);
          const uint64_t off_k = static_cast<uint64_t>(((kk >> 2) * kKBox + (kk & 3) * 32) >> 4);
          if (PAIR) umma_ss_2sm(d_tmem, qd + off_q, kd + off_k, idesc_s, kk != 0);
          else umma_ss(d_tmem, qd + off_q, kd + off_k, idesc_s, kk != 0);
        }
        commit(&s_full[i]);
      };
      auto issue_pv = [&](int i, int st, uint32_t g, bool first) {   // g: running KV-tile count (barrier parity)
        const uint64_t vd = v_desc + static_cast<uint64_t>((st * kKVBytes) >> 4);
        const uint32_t d_tmem = tmem_base + 256 + i * 128;
        const uint32_t p_tmem = tmem_base + i * 128;
        // 16 keys per MMA: 16 rows of 128 B in each d-half box. P arrives in two halves: MMAs 0,1 / 4,5 consume the first 32
        // keys of each warpgroup's 64, MMAs 2,3 / 6,7 the second — the tensor core starts PV while the softmax warps are
        // still exponentiating the second half of the tile.
        wait(&p_ready[i], g & 1);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 1) {
            wait(&p_ready2[i], g & 1);
            tc_fence_after();
          }
#pragma unroll
          for (int wgi = 0; wgi < 2; ++wgi)
#pragma unroll
            for (int tt = 0; tt < 2; ++tt) {
              const int kk = wgi * 4 + h * 2 + tt;
              const uint32_t acc = !(first && (h | wgi | tt) == 0);
              if (PAIR) umma_ts_2sm(d_tmem, p_tmem + (kk & 3) * 8 + (kk >> 2) * 64, vd + static_cast<uint64_t>((kk * 2048) >> 4), idesc_pv, acc);
              else umma_ts(d_tmem, p_tmem + (kk & 3) * 8 + (kk >> 2) * 64, vd + static_cast<uint64_t>((kk * 2048) >> 4), idesc_pv, acc);
            }
        }
        commit(&pv_done[i]);
      };
      uint32_t g = 0;    // KV tiles issued so far (all items)
      uint32_t it = 0;
      for (int w = w0; w < p.n_items; w += w_stride, ++it) {
        decode_item(w);
        mbar_wait_parked(q_full, it & 1);
        if (PAIR) mbar_wait_parked_cluster(q_pair, it & 1);
        wait(&k_full[g % kStages], (g / kStages) & 1);
        tc_fence_after();
        issue_s(0, g % kStages);
        issue_s(1, g % kStages);
        commit(&k_empty[g % kStages]);
        if (n_kv == 1) commit(q_empty);
        for (int j = 0; j < n_kv; ++j, ++g) {
          const int st = g % kStages;
          const int st1 = (g + 1) % kStages;
          wait(&v_full[st], (g / kStages) & 1);
          for (int i = 0; i < 2; ++i) {
            issue_pv(i, st, g, j == 0);
            if (i == 1) commit(&v_empty[st]);
            if (j + 1 < n_kv) {
              if (i == 0) {
                wait(&k_full[st1], ((g + 1) / kStages) & 1);
                tc_fence_after();
              }
              issue_s(i, st1);
              if (i == 1) {
                commit(&k_empty[st1]);
                if (j + 2 == n_kv) commit(q_empty);   // the item's last score MMAs: Q is free for the next item
              }
            }
          }
        }
      }
    }
  } else {
    // ------------------------------ softmax warpgroups ------------------------------
    const int wg = warp >> 2;       // key-column half handled by this warpgroup
    const int quarter = warp & 3;   // TMEM lane quarter
    const int r_local = quarter * 32 + lane;
    const uint32_t lane_bits = static_cast<uint32_t>(quarter * 32) << 16;
    float m[2] = {-INFINITY, -INFINITY}, l[2] = {0.f, 0.f};
    int xcount = 0;
    const uint64_t scale2 = pack2(p.scale_log2, p.scale_log2);

    // Swap a per-row value with the thread of the other warpgroup that shares this row (same lane, warp ^ 4).
    // Slots alternate, so a slot is rewritten only after the partner has passed the next barrier (i.e. read it).
    auto swap_rows = [&](float mine) -> float {
      float* slot = xchg + (xcount & 1) * (2 * kTile);
      ++xcount;
      slot[wg * kTile + r_local] = mine;
      asm volatile("bar.sync %0, 64;" ::"r"(1 + quarter) : "memory");
      return slot[(wg ^ 1) * kTile + r_local];
    };

    // "my part of P is in TMEM": per thread on the CTA's own barrier, or (PAIR) one arrival per warp on the leader's barrier
    auto signal_p = [&](uint64_t* bar) {
      tc_fence_before();
      if (PAIR) {
        __syncwarp();
        if (lane == 0) {
          if (rank == 0) mbar_arrive(bar);
          else mbar_arrive_cluster(bar, 0);
        }
      } else {
        mbar_arrive(bar);
      }
    };
    auto wait_mma = [&](uint64_t* bar, uint32_t parity) {   // barriers the (possibly remote) tensor-core commits signal
      if (PAIR) mbar_wait_cluster(bar, parity);
      else mbar_wait(bar, parity);
    };
    // One pass over this thread's 64 scores, 32 at a time (the second TMEM load is in flight while the first chunk is
    // processed): tracks the maximum and, if EXPS, writes P = 2^(s*scale - m_used) as packed bf16 into pk and
    // returns the sum. KIND selects how the exponentials are made (ExpMixed / ExpMixedClamp / ExpMufu).
    // early_bar != nullptr: the first 32 keys' P (pk[0..15]) are stored to TMEM and announced on early_bar as soon as they
    // exist, before the second 32 scores are touched.
    auto sweep = [&](auto exps_tag, auto max_tag, auto kind_tag, uint32_t t_s, float m_used, float& row_max, uint32_t (&pk)[32],
                     uint64_t* early_bar = nullptr) -> float {
      constexpr bool EXPS = decltype(exps_tag)::value;
      constexpr bool MAXV = decltype(max_tag)::value;
      constexpr int KIND = decltype(kind_tag)::value;
      uint32_t buf[2][32];
      float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      uint64_t acc[4] = {0ull, 0ull, 0ull, 0ull};
      const uint64_t negm2 = pack2(-m_used, -m_used);
      tmem_ld32(t_s, buf[0]);
      tmem_ld32(t_s + 32, buf[1]);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t (&cur)[32] = buf[c];
        if (c == 0) tmem_ld_wait();   // waits for both loads (tcgen05.wait::ld has no partial form)
        if (MAXV && !(kEmu == 9 && EXPS)) {   // DEBUG (FGB_ATTN_EMU=9): no running-max work after the first tile (timing only)
#pragma unroll
          for (int e = 0; e < 32; e += 2)
            mx[(e >> 1) & 3] = fmaxf(mx[(e >> 1) & 3], fmaxf(__uint_as_float(cur[e]), __uint_as_float(cur[e + 1])));
        }
        if (EXPS) {
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const uint64_t x2 = fma2(pack2(__uint_as_float(cur[2 * e]), __uint_as_float(cur[2 * e + 1])), scale2, negm2);
            float p0, p1;
            if (EMU == 7) {          // DEBUG (FGB_ATTN_EMU=7): no exponential at all — isolates the MUFU cost
              unpack2(x2, p0, p1);
            } else if (KIND != ExpMufu::value && (e & 7) < kFrac) {
              exp2_emulated<KIND == ExpMixedClamp::value>(x2, p0, p1);
            } else {
              float x0, x1;
              unpack2(x2, x0, x1);
              p0 = fast_exp2(x0);
              p1 = fast_exp2(x1);
            }
            acc[e & 3] = add2(acc[e & 3], pack2(p0, p1));
            pk[c * 16 + e] = pack_bf16(p0, p1);
          }
          if (c == 0 && early_bar != nullptr) {
            tmem_st16(t_s, pk);
            tmem_st_wait();
            signal_p(early_bar);
          }
        }
      }
      row_max = fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3]));
      float a0, a1, b0, b1;
      unpack2(add2(acc[0], acc[1]), a0, a1);
      unpack2(add2(acc[2], acc[3]), b0, b1);
      return (a0 + a1) + (b0 + b1);
    };

    // Last, partial KV tile (rare path, kept out of line): overwrite the scores of keys that do not exist with -inf.
    auto mask_partial = [&](uint32_t t_s, int valid) {
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
    };
    // One decision per CTA (the two threads of a row must agree; per CTA keeps it simple): AND of `ok` over the 256 softmax threads.
    volatile int* vote = reinterpret_cast<volatile int*>(bars + 4);
    auto vote_all = [&](int slot, bool ok) -> bool {
      if (threadIdx.x == 0) vote[slot] = 1;
      asm volatile("bar.sync 9, 256;" ::: "memory");
      if (!ok) vote[slot] = 0;
      asm volatile("bar.sync 9, 256;" ::: "memory");
      return vote[slot] != 0;
    };

    uint32_t g0 = 0;   // KV tiles of the items before this one: s_full / p_ready / pv_done complete once per tile
    uint32_t it = 0;
    for (int w = w0; w < p.n_items; w += w_stride, ++it) {
    decode_item(w);
    FGB_TRACE(it, 0);
    m[0] = m[1] = -INFINITY;
    l[0] = l[1] = 0.f;
    if (PAIR) {
      // the leader's MMA thread must know that the PEER's Q tiles have landed too: one remote arrival per item
      mbar_wait(q_full, it & 1);
      if (rank == 1 && threadIdx.x == 0) mbar_arrive_cluster(q_pair, 0);
    }
    // ---- bounded-score softmax: a fixed per-row reference instead of a running maximum (AttnParams::kmax)
    int mode = 2;   // 0: reference from the Cauchy-Schwarz bound, 1: anchored on the first tile's maximum, 2: running maximum
    bool head_bound = false;
    if (p.kmax != nullptr && p.qmax != nullptr) {
      // head-level bound: one reference for all rows, nothing to read from the Q tile, nothing to agree on
      const float b = sqrtf(__ldg(p.qmax + head) * __ldg(p.kmax + head)) * p.scale_log2 * 1.0001f;
      if (b <= kBoundFixedMax) {     // NaN fails
        head_bound = true;
        mode = 0;
        m[0] = m[1] = b <= kBoundDirect ? b : kWindowLo - b;
        if (p.stats != nullptr && threadIdx.x == 0) atomicAdd(p.stats + 0, 1);
      }
    }
    if (p.kmax != nullptr && !head_bound) {
      mbar_wait(q_full, it & 1);
      FGB_TRACE(it, 1);
      const float kmax2 = __ldg(p.kmax + head);   // max_j ||k_j||^2 of this head (fgb_head_norm_max)
      float bnd[2];
      bool ok = true;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        // this thread's query row: 2 boxes x 128 B (the 128-byte swizzle only permutes 16-byte chunks inside the row)
        float ss = 0.f;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const uint4* qr = reinterpret_cast<const uint4*>(smem_q + i * kTileBytes + b * kBoxBytes + r_local * 128);
#pragma unroll
          for (int ch = 0; ch < 8; ++ch) {
            const uint4 v = qr[(ch + r_local) & 7];   // staggered chunk order: neighbouring rows hit different banks
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) ss += bf16_lo(w[e]) * bf16_lo(w[e]) + bf16_hi(w[e]) * bf16_hi(w[e]);
          }
        }
        bnd[i] = sqrtf(ss * kmax2) * p.scale_log2 * 1.0001f;
        ok = ok && (bnd[i] <= kBoundFixedMax);   // NaN bounds fail the test and end up on the running-max path
        m[i] = bnd[i] <= kBoundDirect ? bnd[i] : kWindowLo - bnd[i];
      }
      if (vote_all((it & 1) * 2, ok)) {
        mode = 0;
      } else {
        // the Cauchy-Schwarz bound alone leaves no safe window: anchor it on the exact maximum of the first KV tile
        ok = true;
#pragma unroll 1
        for (int i = 0; i < 2; ++i) {
          const uint32_t t_s = tmem_base + lane_bits + i * 128 + wg * 64;
          wait_mma(&s_full[i], g0 & 1);
          tc_fence_after();
          const int valid = p.s_kv - kv_lo * kTile - wg * 64;
          if (valid < 64) mask_partial(t_s, valid);
          uint32_t pk[32];
          float row_max;
          sweep(TagFalse{}, TagTrue{}, ExpMufu{}, t_s, 0.f, row_max, pk);
          const float m0 = fmaxf(row_max, swap_rows(row_max)) * p.scale_log2;
          ok = ok && (bnd[i] - m0 <= kWindowLo + kWindowHi);
          m[i] = fmaxf(bnd[i] - kWindowHi, fminf(bnd[i], m0 + 20.0f));
        }
        mode = vote_all((it & 1) * 2 + 1, ok) ? 1 : 2;
        if (mode == 2) m[0] = m[1] = -INFINITY;
      }
      if (p.stats != nullptr && threadIdx.x == 0) atomicAdd(p.stats + mode, 1);
    }
    FGB_TRACE(it, 2);

    // One (KV tile j, query tile i) step of this thread. MODE and LAST are compile-time so that the steady-state loop of
    // each mode is one compact instruction stream (a runtime dispatch inside the loop cost instruction-cache misses):
    // only the last KV tile of a CTA can be partial, so the -inf masking and the MUFU-only sweep live in the peeled copy.
    auto tile_step = [&](auto mode_tag, auto last_tag, int j, int i) {
      constexpr int MODE = decltype(mode_tag)::value;
      constexpr bool LAST = decltype(last_tag)::value;
      const uint32_t t_s = tmem_base + lane_bits + i * 128 + wg * 64;  // this thread's 64 score columns
      const uint32_t t_o = tmem_base + lane_bits + 256 + i * 128;
      wait_mma(&s_full[i], (g0 + j) & 1);
      tc_fence_after();
      bool partial = false;
      if (LAST) {
        const int valid = p.s_kv - (kv_lo + j) * kTile - wg * 64;  // keys of this half-tile that exist
        partial = valid < 64;
        if (partial) mask_partial(t_s, valid);
      }
      uint32_t pk[32];
      float row_max, sum;
      if (MODE < 2) {            // no maximum, no exchange, no rescale: one pass of exponentials
        if (LAST && partial) sum = sweep(TagTrue{}, TagFalse{}, ExpMufu{}, t_s, m[i], row_max, pk, &p_ready[i]);
        else if (MODE == 0) sum = sweep(TagTrue{}, TagFalse{}, ExpMixed{}, t_s, m[i], row_max, pk, &p_ready[i]);
        else sum = sweep(TagTrue{}, TagFalse{}, ExpMixedClamp{}, t_s, m[i], row_max, pk, &p_ready[i]);
        l[i] += sum;
        tmem_st16(t_s + 16, pk + 16);
        tmem_st_wait();
        signal_p(&p_ready2[i]);
        return;
      }
      if (EMU == 8) {            // DEBUG (FGB_ATTN_EMU=8): no softmax work at all — the tensor / barrier skeleton alone
        signal_p(&p_ready[i]);
        signal_p(&p_ready2[i]);
        return;
      }
      bool redo = false;
      auto exp_sweep = [&]() -> float {   // exponentials against m[i] with the side maximum; masked tiles take the MUFU
        return (LAST && partial) ? sweep(TagTrue{}, TagTrue{}, ExpMufu{}, t_s, m[i], row_max, pk)
                                 : sweep(TagTrue{}, TagTrue{}, ExpMixedClamp{}, t_s, m[i], row_max, pk);
      };
      if (j == 0) {
        sweep(TagFalse{}, TagTrue{}, ExpMufu{}, t_s, 0.f, row_max, pk);  // exact maximum of the first tile
        row_max = fmaxf(row_max, swap_rows(row_max));
        m[i] = row_max * p.scale_log2;
        redo = true;
      } else {
        // Speculate that the running maximum has not grown by more than 2^8: exponentiate against the stale
        // maximum while the true maximum is computed on the side (MUFU and ALU pipes in parallel).
        sum = exp_sweep();
        if (kEmu != 9) row_max = fmaxf(row_max, swap_rows(row_max));   // maximum of the whole 128-key row
        const float m_new = kEmu == 9 ? m[i] : fmaxf(m[i], row_max * p.scale_log2);
        // both threads of a row see the same m_new, so both warps of the pair take the same branch
        if (__any_sync(0xffffffffu, m_new - m[i] > 8.0f)) {
          // rare: rescale the running sum and this warpgroup's half of the O accumulator by 2^(m - m_new), redo the tile
          const float alpha = fast_exp2(m[i] - m_new);
          l[i] *= alpha;
          m[i] = m_new;
          wait_mma(&pv_done[i], (g0 + j - 1) & 1);  // O_i must be quiescent
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
      if (redo) sum = exp_sweep();  // S is still intact in TMEM: P has not been stored yet
      l[i] += sum;
      tmem_st32(t_s, pk);   // P over this thread's own first 32 (consumed) score columns
      tmem_st_wait();
      signal_p(&p_ready[i]);    // the running-max path needs the whole row before any P exists: both halves at once
      signal_p(&p_ready2[i]);
    };
    auto run_tiles = [&](auto mode_tag) {
      // the peeled copy (masking, MUFU-only sweep) runs only when the item's last KV tile really is partial: a full last tile
      // goes through the loop body, so that an item touches ONE copy of the tile step — with 4 KV tiles per item
      // (cross-attention) the second copy pushed the per-item code footprint past the instruction cache (8.8 % no-instruction
      // stalls against 1.4 % in self-attention, profiles/r02_ncu_attn_cross_summary.csv)
      const bool tail_partial = p.s_kv - (kv_lo + n_kv - 1) * kTile < kTile;
      const int n_plain = tail_partial ? n_kv - 1 : n_kv;
      for (int j = 0; j < n_plain; ++j) {
        tile_step(mode_tag, TagFalse{}, j, 0);   // both warpgroups work on query tile 0 together, then on tile 1
        if (j < 4) FGB_TRACE(it, 3 + 2 * j);
        tile_step(mode_tag, TagFalse{}, j, 1);
        if (j < 4) FGB_TRACE(it, 4 + 2 * j);
      }
      if (tail_partial) {
        tile_step(mode_tag, TagTrue{}, n_kv - 1, 0);
        tile_step(mode_tag, TagTrue{}, n_kv - 1, 1);
      }
    };
    if (mode == 0) run_tiles(ModeTag<0>{});
    else if (mode == 1) run_tiles(ModeTag<1>{});
    else run_tiles(ModeTag<2>{});

    // ---- epilogue: each warpgroup normalises and stores its 64 of the 128 output columns of both tiles
#pragma unroll 1
    for (int i = 0; i < 2; ++i) {
      const int row = q0 + i * kTile + r_local;
      const uint32_t t_o = tmem_base + lane_bits + 256 + i * 128 + wg * 64;
      const float l_row = l[i] + swap_rows(l[i]);   // both halves used the same running max
      wait_mma(&pv_done[i], (g0 + n_kv - 1) & 1);
      tc_fence_after();
      if (part >= 0) {
        // split-KV CTA: un-normalised O and (m, l) go to the workspace; attn_combine_kernel merges the chunks
        const int pr = static_cast<int>(rank) * (2 * kTile) + i * kTile + r_local;   // row inside the work item
        float* prow = p.part_o + (static_cast<int64_t>(part) * (kItemTiles * kTile) + pr) * 128 + wg * 64;
        if (wg == 0) p.part_ml[static_cast<int64_t>(part) * (kItemTiles * kTile) + pr] = make_float2(m[i], l_row);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t orr[32];
          tmem_ld32(t_o + c * 32, orr);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<uint4*>(prow + c * 32 + g * 4) = make_uint4(orr[g * 4], orr[g * 4 + 1], orr[g * 4 + 2], orr[g * 4 + 3]);
        }
      } else {
        // final normalisation: O / l -> bf16 -> global (or the owning peer's buffer)
        const float inv_l = 1.0f / l_row;
        if (wg == 0 && p.lse != nullptr && row < p.ld_lse) p.lse[static_cast<int64_t>(head) * p.ld_lse + row] = m[i] + __log2f(l_row);
        if (p.tma_out) {
          // stage the tile in shared memory (128-byte swizzle: chunk ^ (row & 7), conflict-free for one row per lane) and let
          // the TMA write full 128-byte lines: a per-thread store of 16 bytes per row touched 32 lines per warp instruction
          // and held the warps for ~2.8 us per work item (profiles/r02_ncu_attn_cross_before_tma_store.txt)
          // both 32-column chunks are fetched and normalised into packed registers BEFORE the buffer hand-over, so that for the
          // second tile this work overlaps the TMA store of the first one reading the staging buffer (an item's epilogue was
          // 2.35 us of a 9.5 us cross-attention item, profiles/r02_attn_cross_trace.log)
          uint32_t pko[32];
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint32_t orr[32];
            tmem_ld32(t_o + c * 32, orr);
            tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 16; ++e)
              pko[c * 16 + e] = pack_bf16(__uint_as_float(orr[2 * e]) * inv_l, __uint_as_float(orr[2 * e + 1]) * inv_l);
          }
          if (threadIdx.x == 0) tma_store_wait_read<0>();          // the previous tile's store has read the buffer
          asm volatile("bar.sync 9, 256;" ::: "memory");
          uint8_t* srow = smem_o + wg * kBoxBytes + r_local * 128;
#pragma unroll
          for (int g = 0; g < 8; ++g)
            *reinterpret_cast<uint4*>(srow + ((g ^ (r_local & 7)) << 4)) = make_uint4(pko[4 * g], pko[4 * g + 1], pko[4 * g + 2], pko[4 * g + 3]);
          fence_proxy_async_smem();
          asm volatile("bar.sync 9, 256;" ::: "memory");
          if (threadIdx.x == 0) {      // (one store per warpgroup behind 128-thread barriers measured slower: profiles/r02_attn_head_bound_ab.log)
            tma_store_2d(&tmap_o, smem_o, head * 128, q0 + i * kTile);                    // rows >= s_q are clipped by the tensor map
            tma_store_2d(&tmap_o, smem_o + kBoxBytes, head * 128 + 64, q0 + i * kTile);
            tma_store_commit();
          }
        } else {
          __nv_bfloat16* orow = row < p.s_q ? out_row(p, row, head) + wg * 64 : nullptr;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
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
      }
      FGB_TRACE(it, 11 + i);
    }
    FGB_TRACE(it, 13);
    g0 += n_kv;
    }   // items
    if (threadIdx.x == 0) tma_store_wait<0>();
  }

  __syncwarp();   // the single-thread roles rejoin their warps before the (aligned) barrier
  tc_fence_before();
  if (PAIR) cluster_sync_all();   // the leader's MMAs read the peer's shared memory and TMEM: nobody leaves early
  else __syncthreads();
  if (warp == 8) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2sm<512>(tmem_base);
    else tmem_dealloc<512>(tmem_base);
  }
}

constexpr int kDefaultEmu = 3;   // 3 of every 8 score pairs use the FMA-pipe exp2: best in-step (power-capped) time, see DESIGN.md §3.2

// Merge the `split` key-chunk partials of the split units: one warp per query row.
//   M = max_c m_c;  w_c = 2^(m_c - M);  o = sum_c w_c O_c / sum_c w_c l_c
__global__ void __launch_bounds__(256)
attn_combine_kernel(const AttnParams p, int n_split_units) {
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int u = gw / p.item_rows;
  const int r_local = gw % p.item_rows;
  if (u >= n_split_units) return;
  const int unit = p.n_full + u;
  const int head = unit / p.n_pairs;
  const int row = (unit % p.n_pairs) * p.item_rows + r_local;
  if (row >= p.s_q && (p.lse == nullptr || row >= p.ld_lse)) return;
  float mmax = -INFINITY;
  for (int c = 0; c < p.split; ++c)
    mmax = fmaxf(mmax, p.part_ml[(static_cast<int64_t>(u) * p.split + c) * p.item_rows + r_local].x);
  float lsum = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = 0; c < p.split; ++c) {
    const int64_t pr = (static_cast<int64_t>(u) * p.split + c) * p.item_rows + r_local;
    const float2 ml = p.part_ml[pr];
    const float w = (ml.x == -INFINITY) ? 0.f : exp2f(ml.x - mmax);
    lsum += w * ml.y;
    const float4 o = *reinterpret_cast<const float4*>(p.part_o + pr * 128 + lane * 4);
    acc.x += w * o.x; acc.y += w * o.y; acc.z += w * o.z; acc.w += w * o.w;
  }
  const float inv = 1.0f / lsum;
  uint2 out;
  out.x = pack_bf16(acc.x * inv, acc.y * inv);
  out.y = pack_bf16(acc.z * inv, acc.w * inv);
  if (row < p.s_q) *reinterpret_cast<uint2*>(out_row(p, row, head) + lane * 4) = out;
  if (p.lse != nullptr && lane == 0 && row < p.ld_lse) p.lse[static_cast<int64_t>(head) * p.ld_lse + row] = mmax + log2f(lsum);
}

template <int EMU, bool PAIR>
static int launch_attn(int grid, cudaStream_t stream, const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv,
                       const CUtensorMap& to, const AttnParams& p) {
  auto kfn = attn_fwd_kernel<EMU, PAIR>;
  static unsigned long long configured = 0;  // per template instance and device
  if (first_use_on_device(configured)) {
    FGB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
  }
  if (PAIR) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(kAttnThreads);
    cfg.dynamicSmemBytes = kAttnSmem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    FGB_CUDA(cudaLaunchKernelEx(&cfg, kfn, tq, tk, tv, to, p));
  } else {
    kfn<<<grid, kAttnThreads, kAttnSmem, stream>>>(tq, tk, tv, to, p);
  }
  FGB_LAUNCH_CHECK("attn_fwd_kernel");
  return FGB_OK;
}

// One CTA per SM (256 query rows per work item; the default), or CTA pairs (512 rows per item, tcgen05 cta_group::2;
// FGB_ATTN_PAIR=1). Measured on the headline self-attention (profiles/r02_attn_pair_ab.log): the pair variant is bit-correct
// but 45 % SLOWER (9.83 vs 6.75 ms) — every P-ready / S-ready handshake of the two-tile ping-pong then crosses the cluster
// (remote mbarrier arrive + multicast commit, a few hundred cycles each way) and the loop is latency-bound, not operand-bound.
// The environment is read on every call so that a test can exercise both variants in one process.
static bool attn_use_pair(const fgb_ctx* ctx) {
  const char* e = getenv("FGB_ATTN_PAIR");
  return e != nullptr && atoi(e) != 0 && ctx->sm_count >= 2;
}

// How the last, partly filled wave of CTAs is cut along the keys: returns the split factor g (1 = no split) and the
// number of units that are split. With U units on n_sm SMs (one CTA per SM), the last U mod n_sm units would occupy a
// whole wave on their own; cutting each into g key chunks makes the tail ceil(rem*g/n_sm)/g of a wave instead.
static void plan_split(int units, int n_kv, int n_sm, int* split, int* n_split_units) {
  *split = 1;
  *n_split_units = 0;
  const int rem = units % n_sm;
  if (rem == 0 || n_kv < 16) return;
  double best = 1.0;   // tail length in waves without a split
  int best_g = 1;
  const int g_max = n_kv / 8 < 16 ? n_kv / 8 : 16;
  for (int g = 2; g <= g_max; ++g) {
    const int waves = (rem * g + n_sm - 1) / n_sm;
    const double tail = static_cast<double>(waves) / g + 0.01 * g;   // + prologue/epilogue/combine cost per chunk
    if (tail < best - 0.05) {
      best = tail;
      best_g = g;
    }
  }
  if (best_g > 1) {
    *split = best_g;
    *n_split_units = rem;
  }
}

}  // namespace fgb

extern "C" int64_t fgb_attn_workspace_bytes(fgb_ctx* ctx, int32_t s_q, int32_t s_kv, int32_t heads) {
  using namespace fgb;
  if (!ctx || s_q <= 0 || s_kv <= 0 || heads <= 0) return 0;
  const bool pair = attn_use_pair(ctx);
  const int item_rows = pair ? 4 * kTile : 2 * kTile;
  const int n_pairs = (s_q + item_rows - 1) / item_rows;
  int split, n_split;
  plan_split(n_pairs * heads, (s_kv + kTile - 1) / kTile, pair ? ctx->sm_count / 2 : ctx->sm_count, &split, &n_split);
  return static_cast<int64_t>(n_split) * split * item_rows * (128 * 4 + 8);
}

static int attn_fwd_impl(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                         int64_t ldv, void* o, int64_t ldo, int32_t s_q, int32_t s_kv, int32_t heads, float scale,
                         void* lse, int64_t ld_lse, void* workspace, int64_t workspace_bytes, void* const* o_peers,
                         int32_t n_peers, int32_t rows_per_peer, int32_t col_offset, const void* kmax, const void* qmax, void* stream) {
  using namespace fgb;
  FGB_CHECK_ARG(ctx, "fgb_attn_fwd: ctx is NULL");
  FGB_CHECK_ARG(q && k && v && (o || o_peers), "fgb_attn_fwd: NULL tensor pointer");
  FGB_CHECK_ARG(s_q > 0 && s_kv > 0 && heads > 0, "fgb_attn_fwd: empty problem s_q=%d s_kv=%d heads=%d", s_q, s_kv, heads);
  FGB_CHECK_ARG(heads <= 65535, "fgb_attn_fwd: too many heads");
  const int64_t width = static_cast<int64_t>(heads) * FGB_HEAD_DIM;
  FGB_CHECK_ARG(ldq >= width && ldk >= width && ldv >= width && ldo >= width, "fgb_attn_fwd: leading dimension < heads*128");
  FGB_CHECK_ARG(ldo % 8 == 0 && (o_peers || aligned16(o)), "fgb_attn_fwd: o must be 16-byte aligned with ldo %% 8 == 0");
  FGB_CHECK_ARG(workspace == nullptr || aligned16(workspace), "fgb_attn_fwd: workspace must be 16-byte aligned");
  if (o_peers) {
    FGB_CHECK_ARG(n_peers > 0 && n_peers <= FGB_MAX_PEERS && rows_per_peer > 0 && static_cast<int64_t>(rows_per_peer) * n_peers >= s_q &&
                      col_offset >= 0 && col_offset % 8 == 0 && ldo >= col_offset + width,
                  "fgb_attn_fwd_scatter: peers=%d rows_per_peer=%d col_offset=%d", n_peers, rows_per_peer, col_offset);
    for (int i = 0; i < n_peers; ++i) FGB_CHECK_ARG(o_peers[i] && aligned16(o_peers[i]), "fgb_attn_fwd_scatter: peer output %d", i);
  }

  CUtensorMap tq, tk, tv;
  int rc = make_tmap_bf16_2d(ctx, &tq, q, s_q, width, ldq, kTile);
  if (rc) return rc;
  const bool pair = attn_use_pair(ctx);
  const int item_rows = pair ? 4 * kTile : 2 * kTile;
  const int workers = pair ? ctx->sm_count / 2 : ctx->sm_count;
  rc = make_tmap_bf16_2d(ctx, &tk, k, s_kv, width, ldk, pair ? kTile / 2 : kTile);   // PAIR: each CTA loads 64 of a tile's 128 keys
  if (rc) return rc;
  rc = make_tmap_bf16_2d(ctx, &tv, v, s_kv, width, ldv, kTile);
  if (rc) return rc;
  CUtensorMap to = tq;
  if (!o_peers) {
    if ((rc = make_tmap_bf16_2d(ctx, &to, o, s_q, width, ldo, kTile))) return rc;
  }

  AttnParams p;
  p.o = static_cast<__nv_bfloat16*>(o);
  p.ldo = ldo;
  p.s_q = s_q;
  p.s_kv = s_kv;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.lse = static_cast<float*>(lse);
  p.ld_lse = ld_lse;
  p.tma_out = o_peers ? 0 : 1;
  p.kmax = static_cast<const float*>(kmax);
  p.qmax = kmax ? static_cast<const float*>(qmax) : nullptr;
  p.stats = kmax ? ctx->attn_stats : nullptr;
  p.rows_per_peer = o_peers ? rows_per_peer : 0;
  p.col_offset = col_offset;
  for (int i = 0; i < FGB_MAX_PEERS; ++i) p.o_peers[i] = (o_peers && i < n_peers) ? static_cast<__nv_bfloat16*>(o_peers[i]) : nullptr;
  FGB_CHECK_ARG(lse == nullptr || ld_lse >= s_q, "fgb_attn_fwd: ld_lse=%lld < s_q", (long long)ld_lse);
  p.item_rows = item_rows;
  p.n_pairs = (s_q + item_rows - 1) / item_rows;
  const int64_t units64 = static_cast<int64_t>(p.n_pairs) * heads;
  FGB_CHECK_ARG(units64 < (1ll << 30), "fgb_attn_fwd: problem too large");
  const int units = static_cast<int>(units64);
  int split = 1, n_split = 0;
  // the key split needs the caller's scratch (the library never allocates); without it every unit runs whole
  if (workspace != nullptr) plan_split(units, (s_kv + kTile - 1) / kTile, workers, &split, &n_split);
  const int64_t need = static_cast<int64_t>(n_split) * split * item_rows * (128 * 4 + 8);
  if (need > workspace_bytes) {
    split = 1;
    n_split = 0;
  }
  p.split = split;
  p.n_full = units - n_split;
  p.part_o = static_cast<float*>(workspace);
  p.part_ml = reinterpret_cast<float2*>(static_cast<char*>(workspace) + static_cast<int64_t>(n_split) * split * item_rows * 128 * 4);

  // fraction of exponentials moved off the MUFU (EMU of every 8 pairs); FGB_ATTN_EMU overrides for tuning
  static int emu = -1;
  if (emu < 0) {
    const char* env = getenv("FGB_ATTN_EMU");
    emu = env ? atoi(env) : kDefaultEmu;
    if (emu < 0 || emu > 9) emu = kDefaultEmu;
  }
  p.n_items = p.n_full + n_split * split;
  const int grid = (p.n_items < workers ? p.n_items : workers) * (pair ? 2 : 1);   // persistent: one CTA (pair) per SM (pair)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
#define FGB_ATTN_CASE(E)                                                                                  \
  case E:                                                                                                 \
    rc = pair ? launch_attn<E, true>(grid, st, tq, tk, tv, to, p) : launch_attn<E, false>(grid, st, tq, tk, tv, to, p); \
    break;
  switch (emu) {
    FGB_ATTN_CASE(0) FGB_ATTN_CASE(1) FGB_ATTN_CASE(2) FGB_ATTN_CASE(3) FGB_ATTN_CASE(4) FGB_ATTN_CASE(5) FGB_ATTN_CASE(6)
    FGB_ATTN_CASE(7) FGB_ATTN_CASE(8)
    default:
      rc = pair ? launch_attn<9, true>(grid, st, tq, tk, tv, to, p) : launch_attn<9, false>(grid, st, tq, tk, tv, to, p);
      break;
  }
#undef FGB_ATTN_CASE
  if (rc) return rc;
  if (n_split > 0) {
    const int warps = n_split * item_rows;
    attn_combine_kernel<<<(warps * 32 + 255) / 256, 256, 0, st>>>(p, n_split);
    FGB_LAUNCH_CHECK("attn_combine_kernel");
  }
  return FGB_OK;
}

// Host-side walk of the work list of one attention launch (no device needed): the persistent CTAs take items w0, w0 + stride, ...
// of [n_full whole units | split units x split key chunks]; every (unit, KV tile) must be visited exactly once, chunks must be
// non-empty, and the partial buffers must fit the workspace fgb_attn_workspace_bytes() asks for. The item -> (unit, KV range)
// arithmetic below is the kernel's decode_item, restated (the kernel keeps its lambda so that its code does not change).
extern "C" int fgb_attn_schedule_check(int32_t s_q, int32_t s_kv, int32_t heads, int32_t sm_count, int32_t with_workspace,
                                       int32_t* split_out, int32_t* n_split_units_out) {
  using namespace fgb;
  if (s_q <= 0 || s_kv <= 0 || heads <= 0 || sm_count <= 0) return set_error(FGB_ERR_INVALID, "fgb_attn_schedule_check: bad shape");
  const int item_rows = 2 * kTile;
  const int n_pairs = (s_q + item_rows - 1) / item_rows;
  const int units = n_pairs * heads;
  const int n_kv_all = (s_kv + kTile - 1) / kTile;
  int split = 1, n_split = 0;
  if (with_workspace) plan_split(units, n_kv_all, sm_count, &split, &n_split);
  fgb_ctx ctx;
  ctx.sm_count = sm_count;
  const int64_t need = static_cast<int64_t>(n_split) * split * item_rows * (128 * 4 + 8);
  if (with_workspace && need != fgb_attn_workspace_bytes(&ctx, s_q, s_kv, heads))
    return set_error(FGB_ERR_INVALID, "attention schedule: workspace %lld != fgb_attn_workspace_bytes", (long long)need);
  const int n_full = units - n_split, n_items = n_full + n_split * split;
  std::vector<int> seen(static_cast<size_t>(units) * n_kv_all, 0);
  const int workers = n_items < sm_count ? n_items : sm_count;
  int visited = 0;
  for (int cta = 0; cta < workers; ++cta)
    for (int w = cta; w < n_items; w += workers, ++visited) {
      int unit = w, kv_lo = 0, kv_hi = n_kv_all;
      if (unit >= n_full) {
        const int part = unit - n_full, chunk = part % split;
        unit = n_full + part / split;
        kv_lo = static_cast<int>(static_cast<int64_t>(chunk) * n_kv_all / split);
        kv_hi = static_cast<int>(static_cast<int64_t>(chunk + 1) * n_kv_all / split);
        if (part >= n_split * split) return set_error(FGB_ERR_INVALID, "attention schedule: partial slot %d outside the workspace", part);
      }
      if (unit < 0 || unit >= units || kv_hi <= kv_lo)
        return set_error(FGB_ERR_INVALID, "attention schedule: item %d -> unit %d, KV tiles [%d,%d)", w, unit, kv_lo, kv_hi);
      for (int j = kv_lo; j < kv_hi; ++j) ++seen[static_cast<size_t>(unit) * n_kv_all + j];
    }
  if (visited != n_items) return set_error(FGB_ERR_INVALID, "attention schedule: %d of %d items visited", visited, n_items);
  for (size_t i = 0; i < seen.size(); ++i)
    if (seen[i] != 1) return set_error(FGB_ERR_INVALID, "attention schedule: (unit, KV tile) %zu visited %d times", i, seen[i]);
  if (split_out) *split_out = split;
  if (n_split_units_out) *n_split_units_out = n_split;
  return FGB_OK;
}

#if FGB_ATTN_TRACE
extern "C" int fgb_debug_attn_trace(unsigned long long* host, int n) {
  return cudaMemcpyFromSymbol(host, fgb::g_attn_trace, sizeof(unsigned long long) * (n < 256 ? n : 256)) == cudaSuccess ? 0 : 1;
}
#endif

extern "C" int fgb_attn_set_stats(fgb_ctx* ctx, void* counts_dev) {
  if (!ctx) return fgb::set_error(FGB_ERR_INVALID, "fgb_attn_set_stats: ctx is NULL");
  if (counts_dev && (reinterpret_cast<uintptr_t>(counts_dev) & 3u)) return fgb::set_error(FGB_ERR_INVALID, "fgb_attn_set_stats: unaligned");
  ctx->attn_stats = static_cast<int32_t*>(counts_dev);
  return FGB_OK;
}

extern "C" int fgb_attn_fwd_ex(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                               int64_t ldv, void* o, int64_t ldo, int32_t s_q, int32_t s_kv, int32_t heads, float scale,
                               void* lse, int64_t ld_lse, void* workspace, int64_t workspace_bytes, void* stream) {
  return attn_fwd_impl(ctx, q, ldq, k, ldk, v, ldv, o, ldo, s_q, s_kv, heads, scale, lse, ld_lse, workspace, workspace_bytes, nullptr, 0,
                       0, 0, nullptr, nullptr, stream);
}

extern "C" int fgb_attn_fwd_bounded(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                    int64_t ldv, void* o, int64_t ldo, int32_t s_q, int32_t s_kv, int32_t heads, float scale,
                                    const void* kmax, void* lse, int64_t ld_lse, void* workspace, int64_t workspace_bytes,
                                    void* const* o_peers, int32_t n_peers, int32_t rows_per_peer, int32_t col_offset, void* stream) {
  if (!kmax) return fgb::set_error(FGB_ERR_INVALID, "fgb_attn_fwd_bounded: kmax is NULL");
  return attn_fwd_impl(ctx, q, ldq, k, ldk, v, ldv, o, ldo, s_q, s_kv, heads, scale, lse, ld_lse, workspace, workspace_bytes, o_peers,
                       n_peers, rows_per_peer, col_offset, kmax, nullptr, stream);
}

extern "C" int fgb_attn_fwd_bounded_qk(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                       int64_t ldv, void* o, int64_t ldo, int32_t s_q, int32_t s_kv, int32_t heads, float scale,
                                       const void* kmax, const void* qmax, void* lse, int64_t ld_lse, void* workspace,
                                       int64_t workspace_bytes, void* const* o_peers, int32_t n_peers, int32_t rows_per_peer,
                                       int32_t col_offset, void* stream) {
  if (!kmax) return fgb::set_error(FGB_ERR_INVALID, "fgb_attn_fwd_bounded_qk: kmax is NULL");
  return attn_fwd_impl(ctx, q, ldq, k, ldk, v, ldv, o, ldo, s_q, s_kv, heads, scale, lse, ld_lse, workspace, workspace_bytes, o_peers,
                       n_peers, rows_per_peer, col_offset, kmax, qmax, stream);
}

extern "C" int fgb_attn_fwd_scatter(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                    int64_t ldv, void* const* o_peers, int32_t n_peers, int64_t ldo, int32_t rows_per_peer,
                                    int32_t col_offset, int32_t s_q, int32_t s_kv, int32_t heads, float scale, void* workspace,
                                    int64_t workspace_bytes, void* stream) {
  if (!o_peers) return fgb::set_error(FGB_ERR_INVALID, "fgb_attn_fwd_scatter: o_peers is NULL");
  return attn_fwd_impl(ctx, q, ldq, k, ldk, v, ldv, nullptr, ldo, s_q, s_kv, heads, scale, nullptr, 0, workspace, workspace_bytes, o_peers,
                       n_peers, rows_per_peer, col_offset, nullptr, nullptr, stream);
}

extern "C" int fgb_attn_fwd(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                            int64_t ldv, void* o, int64_t ldo, int32_t s_q, int32_t s_kv, int32_t heads, float scale,
                            void* stream) {
  return fgb_attn_fwd_ex(ctx, q, ldq, k, ldk, v, ldv, o, ldo, s_q, s_kv, heads, scale, nullptr, 0, nullptr, 0, stream);
}
