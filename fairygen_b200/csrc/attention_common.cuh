// fairygen_b200 — parameters and arithmetic helpers of the attention-forward kernel (attention.cu; also used by the
// experimental variants kept under experiments/).
#pragma once

#include "common.cuh"
#include "host.h"

namespace fgb {

constexpr int kAttnThreads = 320;
constexpr int kTile = 128;                  // query rows per tile == keys per KV tile == head_dim
constexpr int kBoxBytes = kTile * 64 * 2;   // one TMA box: 128 rows x 64 bf16 = 16 KB
constexpr int kTileBytes = 2 * kBoxBytes;   // 128 x 128 bf16 = 32 KB (two 64-column boxes)
constexpr int kKVStages = 2;
// Q + K/V rings + one 128 x 128 bf16 staging tile for the TMA-store epilogue + barriers + row exchange = 231 680 bytes; the
// remaining 768 bytes up to the 227 KB limit are the alignment slack (the dynamic segment is declared 1024-byte aligned and
// the kernel traps if the slack would not do).
constexpr int kAttnAlignSlack = 768;
constexpr int kAttnSmem = 2 * kTileBytes /*Q*/ + 2 * kKVStages * kTileBytes /*K,V*/ + kTileBytes /*O staging*/ + kAttnAlignSlack + 256 /*barriers*/ +
                          2 * 2 * kTile * 4 /*row-max exchange*/;
static_assert(kAttnSmem <= 232448, "attention kernel shared memory exceeds 227 KB");

struct AttnParams {
  __nv_bfloat16* o;
  int64_t ldo;
  int32_t s_q, s_kv;
  float scale_log2;
  // Work list: CTA b < n_full handles unit b over all KV tiles; the remaining CTAs handle the LAST units of
  // the list, each split `split` ways along the keys (wave-quantisation fix, see fgb_attn_fwd_ex). A unit is
  // (head, pair of query tiles): unit = head * n_pairs + pair.
  int32_t n_pairs, n_full, split;
  int32_t item_rows;  // query rows of one work item: 256 (one CTA) or 512 (CTA pair)
  int32_t n_items;  // n_full + (split units) * split: the work list the persistent CTAs walk with stride gridDim.x
  float* part_o;    // [split CTA][256 rows][128] un-normalised fp32 partial outputs
  float2* part_ml;  // [split CTA][256 rows] (running max in log2 units, row sum)
  float* lse;       // optional [heads][ld_lse]: log2-domain log-sum-exp of the scaled scores (for the backward pass);
  int64_t ld_lse;   // rows in [s_q, ld_lse) get the value of an all-zero query row, so the backward needs no masks
  // Fused Ulysses return exchange: when rows_per_peer > 0 the output row of global token t goes straight into the
  // token-major buffer of the rank that owns t (peer memory over NVLink): o_peers[t / rows_per_peer] + (t %
  // rows_per_peer) * ldo + (col_offset + head*128); `o` is unused.
  __nv_bfloat16* o_peers[FGB_MAX_PEERS];
  int32_t rows_per_peer, col_offset;
  // Bounded-score softmax: kmax[head] = max_j ||k_j||^2 over all keys of the head (fp32, NULL = running-max softmax).
  // |s_ij| <= B_i = ||q_i||·sqrt(kmax)·scale·log2e (Cauchy-Schwarz). P = 2^(s - R_i) with a FIXED per-row reference R_i needs no
  // running max, no rescale and no exchange; it is exact as long as the row's largest score lands inside the exponent window
  // that bf16 (P) and fp32 (l, O) share: s_max - R_i in [-120, +100] (the row sum of 27 280 terms <= 2^100 stays below 2^115).
  //   mode 0 (B_i <= 110 for every row of the CTA):  R_i = B_i (B_i <= 60) or 120 - B_i  ->  s - R_i in [-120, 100] for EVERY score
  //   mode 1 (otherwise): the exact maximum m0 of the row's first KV tile anchors the window (m0 <= s_max <= B_i):
  //           R_i = max(B_i - 100, min(B_i, m0 + 20)), valid when B_i - m0 <= 220
  //   mode 2: a CTA with a row that fails both runs the running-max path. stats (optional, int32[3]) counts CTAs per mode.
  const float* kmax;
  // Optional qmax[head] = max_i ||q_i||^2 over ALL query rows of the head (a by-product of the q/k norm kernels). When
  // B_head = sqrt(qmax·kmax)·scale·log2e <= 110 the whole head runs mode 0 with ONE reference for every row (R = B_head, or
  // 120 - B_head above 60): no per-row norm pass over the Q tile and no CTA vote at the start of each work item — 1.3 us of a
  // 9.5 us cross-attention item (profiles/r02_attn_cross_trace.log). P, l and O all carry 8 exponent bits, so a looser
  // reference costs no precision inside the window. Heads that fail the test take the per-row path above.
  const float* qmax;
  int32_t* stats;
  int32_t tma_out;   // 1: `o` is written through the output tensor map (staged tiles, full lines); 0: per-row stores (peers / partials)
};

constexpr float kBoundDirect = 60.0f;     // B <= this: R = B
constexpr float kBoundFixedMax = 110.0f;  // B <= this: R = 120 - B (mode 0)
constexpr float kWindowLo = 120.0f;       // s_max - R >= -kWindowLo
constexpr float kWindowHi = 100.0f;       // s_max - R <= +kWindowHi

__device__ __forceinline__ __nv_bfloat16* out_row(const AttnParams& p, int row, int head) {
  if (p.rows_per_peer > 0) {
    const int peer = row / p.rows_per_peer;
    return p.o_peers[peer] + static_cast<int64_t>(row - peer * p.rows_per_peer) * p.ldo + p.col_offset + head * 128;
  }
  return p.o + static_cast<int64_t>(row) * p.ldo + head * 128;
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for a pair on the FMA/ALU pipes instead of the MUFU (which is as busy as the tensor pipe in this
// kernel): round-to-nearest split x = n + f via the 1.5*2^23 trick, degree-3 minimax polynomial for 2^f
// on [-0.5, 0.5] (max relative error 7.5e-5, far below the bf16 rounding of P), exponent add in integer.
// CLAMP = false: the caller guarantees x in [-126, 127] (bounded-score mode 0). CLAMP = true: x is clamped at -126 from below
// with a NaN-propagating max (a NaN score stays NaN, as on the MUFU lanes), so the result of a very negative score is 2^-126
// instead of 0 — which is why tiles with masked (-inf) scores take the MUFU-only sweep.
template <bool CLAMP>
__device__ __forceinline__ void exp2_emulated(uint64_t x2, float& r0, float& r1) {
  if (CLAMP) {
    float x0, x1;
    unpack2(x2, x0, x1);
    asm("max.NaN.f32 %0, %0, %1;" : "+f"(x0) : "f"(-126.0f));
    asm("max.NaN.f32 %0, %0, %1;" : "+f"(x1) : "f"(-126.0f));
    x2 = pack2(x0, x1);
  }
  const uint64_t kMagic = pack2(12582912.0f, 12582912.0f);
  const uint64_t kNegMagic = pack2(-12582912.0f, -12582912.0f);
  const uint64_t kMinusOne = pack2(-1.0f, -1.0f);
  const uint64_t c0 = pack2(0.9999280572f, 0.9999280572f), c1 = pack2(0.6932609677f, 0.6932609677f);
  const uint64_t c2 = pack2(0.2426111251f, 0.2426111251f), c3 = pack2(0.0551716685f, 0.0551716685f);
  const uint64_t t = add2(x2, kMagic);            // integer part lands in the low mantissa bits
  const uint64_t n = add2(t, kNegMagic);          // round(x) as float
  const uint64_t f = fma2(n, kMinusOne, x2);      // x - round(x)
  uint64_t p = fma2(f, c3, c2);
  p = fma2(p, f, c1);
  p = fma2(p, f, c0);
  float p0, p1, t0, t1;
  unpack2(p, p0, p1);
  unpack2(t, t0, t1);
  r0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  r1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}

struct TagTrue { static constexpr bool value = true; };
struct TagFalse { static constexpr bool value = false; };
template <int M>
struct ModeTag { static constexpr int value = M; };
// how a sweep turns scores into exponentials
struct ExpMixed { static constexpr int value = 0; };       // MUFU + FMA-pipe emulation, no clamp (mode 0, full tiles)
struct ExpMixedClamp { static constexpr int value = 1; };  // MUFU + clamped emulation (anchored / running-max references)
struct ExpMufu { static constexpr int value = 2; };        // MUFU only: 2^(-inf) = 0 exactly (tiles with masked keys)


}  // namespace fgb
