// fairygen_b200 — bf16 GEMM with fused epilogues on tcgen05 tensor cores (sm_100a).
//
//   C[m,n] = epilogue(A[m,k] · W[n,k]ᵀ + bias[n])
//
// Replaces every nn.Linear of the DiT block (reference: animation/diffsynth/models/wan_video_dit.py
// :140-146 q/k/v/o, :176-185 cross q/k/v/o, :208-209 ffn, :258 head, :305 patch embedding as a
// GEMM, :307-318 time/text embeddings) together with the elementwise ops that follow them
// (GELU-tanh :208, GateModule :192-193, residual add :226).
//
// Design (one persistent CTA per SM, warp-specialised, 192 threads):
//   warp 0      TMA producer: A tile [128 x 64] and W tile [256 x 64] (128-byte swizzle) into a
//               4-stage shared-memory ring, completion on mbarriers.
//   warp 1      MMA issuer: one thread issues tcgen05.mma (M=128, N=256, K=16, bf16 -> fp32) with
//               the accumulator in TMEM; two 256-column accumulators so the epilogue of tile i
//               overlaps the main loop of tile i+1. tcgen05.commit frees ring slots.
//   warps 2-5   epilogue: tcgen05.ld (one accumulator row per thread), bias / GELU / gate /
//               residual in fp32 with the reference's bf16 rounding points, 16-byte global stores.
// Tiles are rasterised in groups of 8 M-blocks so the CTAs resident at one time share A and W
// tiles through the 126 MB L2.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "common.cuh"
#include "host.h"

namespace fgb {

constexpr int kBM = 128;
constexpr int kBN = 256;
constexpr int kBK = 64;
constexpr int kStages = 4;
constexpr int kGroupM = 8;
constexpr int kABytes = kBM * kBK * 2;  // 16 KB
constexpr int kBBytes = kBN * kBK * 2;  // 32 KB
constexpr int kGemmThreads = 192;
constexpr int kGemmSmem = kStages * (kABytes + kBBytes) + 1024 /*align slack*/ + 256 /*barriers*/;

struct GemmParams {
  const __nv_bfloat16* bias;
  __nv_bfloat16* c;
  const __nv_bfloat16* gate0;
  const __nv_bfloat16* gate1;
  int64_t ldc;
  int32_t m, n, k;
  int32_t rows_gate0;
  int32_t m_blocks, n_blocks, tiles, k_blocks;
  int32_t k2_blocks;   // extra K-blocks taken from the second operand pair (tmap_a2, tmap_b2): acc += A2 · W2ᵀ (or A2 · W2)
  // Convolution as a GEMM over shifted rows (fgb_conv_taps_bf16): K-block kb belongs to tap kb / tap_kblocks and reads the A
  // rows [m_blk*128 + a_row0 + tap_off[tap], +128) — TMA zero-fills rows outside the tensor. tap_kblocks = 0: plain GEMM.
  int32_t tap_kblocks;
  int32_t a_row0;
  int32_t grid_h, grid_w;   // > 0: C rows are positions of a [.., grid_h, grid_w] grid whose 1-wide border is written as zero
  int32_t tap_off[27];
};

__device__ __forceinline__ void tile_coords(const GemmParams& p, int tile, int& m_blk, int& n_blk) {
  const int group_size = kGroupM * p.n_blocks;
  const int group = tile / group_size;
  const int first_m = group * kGroupM;
  const int gsz = min(p.m_blocks - first_m, kGroupM);
  const int in_group = tile - group * group_size;
  m_blk = first_m + in_group % gsz;
  n_blk = in_group / gsz;
}

// two elements per instruction on the FMA pipe (fp32x2); tanh stays on the MUFU
__device__ __forceinline__ void gelu_tanh_2(float& x0, float& x1) {
  const float kAlpha = 0.7978845608028654f, kAB = 0.7978845608028654f * 0.044715f;
  const uint64_t x2 = pack2(x0, x1);
  const uint64_t t = fma2(mul2(x2, x2), pack2(kAB, kAB), pack2(kAlpha, kAlpha));   // alpha * (1 + beta x^2)
  float i0, i1;
  unpack2(mul2(t, x2), i0, i1);
  float t0, t1;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(i0));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(i1));
  const uint64_t h = mul2(x2, pack2(0.5f, 0.5f));
  unpack2(fma2(h, pack2(t0, t1), h), x0, x1);
}

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3)))  — nn.GELU(approximate='tanh')
  const float kAlpha = 0.7978845608028654f, kBeta = 0.044715f;
  float inner = kAlpha * (x + kBeta * x * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(inner));
  return 0.5f * x * (1.0f + t);
}

// B_MN = false: W is [n, k] (nn.Linear weight, K-major B operand)            C = A · Wᵀ   (forward)
// B_MN = true : W is [k, n] (the same nn.Linear weight read as [out=k, in=n])  C = A · W    (dgrad: dX = dY · W)
//               B tile = 4 TMA boxes of [64 k-rows][64 n-cols], fed to the tensor core as an MN-major operand.
// TAPS = true: convolution mode (fgb_conv_taps_bf16): per-tap row offsets in the producer, grid-border masking in the epilogue.
// A compile-time switch so that the DiT instantiations carry none of it.
template <int EPI, bool B_MN, bool TAPS>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_a2, const __grid_constant__ CUtensorMap tmap_b2, const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (kABytes + kBBytes));
  uint64_t* full = bars;                   // [kStages]  TMA -> MMA
  uint64_t* empty = bars + kStages;        // [kStages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * kStages;      // [2]   MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * kStages + 2; // [2]   epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    if (p.k2_blocks > 0) {
      tma_prefetch_desc(&tmap_a2);
      tma_prefetch_desc(&tmap_b2);
    }
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 128);
    }
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // ------------------------------- TMA producer -------------------------------
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x) {
        int m_blk, n_blk;
        tile_coords(p, tile, m_blk, n_blk);
        const int k_total = p.k_blocks + p.k2_blocks;
        for (int kb = 0; kb < k_total; ++kb) {
          mbar_wait(&empty[stage], phase ^ 1);
          mbar_expect_tx(&full[stage], kABytes + kBBytes);
          // the last k2_blocks K-blocks come from the second operand pair (a rank-r LoRA term folded into the GEMM)
          const bool second = kb >= p.k_blocks;
          const CUtensorMap* ma = second ? &tmap_a2 : &tmap_a;
          const CUtensorMap* mb = second ? &tmap_b2 : &tmap_b;
          const int kc = (second ? kb - p.k_blocks : kb) * kBK;
          if constexpr (TAPS) {
            const int tap = kb / p.tap_kblocks;
            tma_load_2d(smem_a + stage * kABytes, ma, &full[stage], (kb - tap * p.tap_kblocks) * kBK,
                        m_blk * kBM + p.a_row0 + p.tap_off[tap], kEvictNormal);
          } else {
            tma_load_2d(smem_a + stage * kABytes, ma, &full[stage], kc, m_blk * kBM, kEvictNormal);
          }
          if (B_MN) {
#pragma unroll
            for (int b = 0; b < kBN / 64; ++b)
              tma_load_2d(smem_b + stage * kBBytes + b * (kBK * 128), mb, &full[stage], n_blk * kBN + b * 64, kc, kEvictLast);
          } else {
            tma_load_2d(smem_b + stage * kBBytes, mb, &full[stage], kc, n_blk * kBN, kEvictLast);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (elect_one()) {
      // ------------------------------- MMA issuer ---------------------------------
      constexpr uint32_t idesc = make_idesc_bf16(kBM, kBN, 0, B_MN ? 1 : 0);
      // descriptors are built once; per MMA only an offset is added to the 14-bit address field
      const uint64_t a_desc0 = make_sdesc_sw128(smem_u32(smem_a), 16, 1024);
      // MN-major B: rows are k (128 B = 64 n each), the next 64 n-columns live kBK*128 bytes further (LBO)
      const uint64_t b_desc0 = B_MN ? make_sdesc_sw128(smem_u32(smem_b), kBK * 128, 1024) : make_sdesc_sw128(smem_u32(smem_b), 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      int local = 0;
      for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++local) {
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kBN;
        const int k_total = p.k_blocks + p.k2_blocks;
        for (int kb = 0; kb < k_total; ++kb) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = a_desc0 + static_cast<uint64_t>((stage * kABytes) >> 4);
          const uint64_t bdesc = b_desc0 + static_cast<uint64_t>((stage * kBBytes) >> 4);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_ss(tmem_d, adesc + static_cast<uint64_t>((k * 32) >> 4),
                    bdesc + static_cast<uint64_t>((B_MN ? k * 2048 : k * 32) >> 4), idesc, (kb | k) != 0);
          tc_commit(&empty[stage]);  // ring slot reusable once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        tc_commit(&tmem_full[acc]);  // accumulator complete
      }
    }
  } else {
    // --------------------------------- epilogue -----------------------------------
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are visible to this warp
    int local = 0;
    for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++local) {
      int m_blk, n_blk;
      tile_coords(p, tile, m_blk, n_blk);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * kBM + quarter * 32 + lane;
      const bool row_ok = row < p.m;
      bool border = false;
      if (TAPS && p.grid_w > 0) {
        const int pos = row % (p.grid_h * p.grid_w), gy = pos / p.grid_w, gx = pos - gy * p.grid_w;
        border = gy == 0 || gy == p.grid_h - 1 || gx == 0 || gx == p.grid_w - 1;
      }
      __nv_bfloat16* crow = p.c + static_cast<int64_t>(row) * p.ldc;
      const __nv_bfloat16* gate = nullptr;
      if (EPI == FGB_EPI_GATED_RESIDUAL) gate = (row < p.rows_gate0) ? p.gate0 : p.gate1;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kBN;
      // residual epilogues read C: the 64 bytes of the NEXT 32-column chunk are requested before the current chunk is
      // processed, so their latency hides under the TMEM load and the math (with K = 3072 the epilogue of a tile is as
      // long as its main loop, and a serial load -> use chain per chunk cost 15 %)
      constexpr bool kReadsC = (EPI == FGB_EPI_GATED_RESIDUAL || EPI == FGB_EPI_RESIDUAL);
      uint4 xcur[4], xnext[4];
      auto load_c = [&](int c, uint4 (&dst)[4]) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = n_blk * kBN + c * 32 + g * 8;
          dst[g] = (row_ok && col < p.n) ? *reinterpret_cast<const uint4*>(crow + col) : make_uint4(0, 0, 0, 0);
        }
      };
      if (kReadsC) load_c(0, xcur);
#pragma unroll 1
      for (int c = 0; c < kBN / 32; ++c) {
        const int col0 = n_blk * kBN + c * 32;
        if (col0 >= p.n) break;  // warp-uniform
        uint32_t r[32];
        tmem_ld32(taddr + c * 32, r);
        if (kReadsC && c + 1 < kBN / 32) load_c(c + 1, xnext);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = col0 + g * 8;
          if (col >= p.n) break;
          float y[8];
          uint4 bv = make_uint4(0, 0, 0, 0);
          if (p.bias) bv = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
          const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            y[2 * i] = round_bf16(__uint_as_float(r[g * 8 + 2 * i]) + bf16_lo(bw[i]));
            y[2 * i + 1] = round_bf16(__uint_as_float(r[g * 8 + 2 * i + 1]) + bf16_hi(bw[i]));
          }
          if (EPI == FGB_EPI_BIAS_GELU_TANH) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) gelu_tanh_2(y[i], y[i + 1]);
          }
          if (kReadsC) {
            if (row_ok) {
              const uint4 xv = xcur[g];
              const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
              if (EPI == FGB_EPI_GATED_RESIDUAL) {
                const uint4 gv = __ldg(reinterpret_cast<const uint4*>(gate + col));
                const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  y[2 * i] = bf16_lo(xw[i]) + round_bf16(bf16_lo(gw[i]) * y[2 * i]);
                  y[2 * i + 1] = bf16_hi(xw[i]) + round_bf16(bf16_hi(gw[i]) * y[2 * i + 1]);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  y[2 * i] = bf16_lo(xw[i]) + y[2 * i];
                  y[2 * i + 1] = bf16_hi(xw[i]) + y[2 * i + 1];
                }
              }
            }
          }
          if (row_ok) {
            uint4 o;
            if (TAPS && border) {
#pragma unroll
              for (int i = 0; i < 8; ++i) y[i] = 0.f;
            }
            o.x = pack_bf16(y[0], y[1]);
            o.y = pack_bf16(y[2], y[3]);
            o.z = pack_bf16(y[4], y[5]);
            o.w = pack_bf16(y[6], y[7]);
            *reinterpret_cast<uint4*>(crow + col) = o;
          }
        }
        if (kReadsC) {
#pragma unroll
          for (int g = 0; g < 4; ++g) xcur[g] = xnext[g];
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

template <int EPI, bool B_MN = false, bool TAPS = false>
static int launch_gemm(fgb_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p,
                       cudaStream_t stream, const CUtensorMap* ta2 = nullptr, const CUtensorMap* tb2 = nullptr) {
  auto kfn = gemm_bf16_kernel<EPI, B_MN, TAPS>;
  static unsigned long long configured = 0;  // per template instance and device
  if (first_use_on_device(configured)) {
    FGB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kGemmSmem));
  }
  const int grid = p.tiles < ctx->sm_count ? p.tiles : ctx->sm_count;
  kfn<<<grid, kGemmThreads, kGemmSmem, stream>>>(ta, tb, ta2 ? *ta2 : ta, tb2 ? *tb2 : tb, p);
  FGB_LAUNCH_CHECK("gemm_bf16_kernel");
  return FGB_OK;
}

// =====================================================================================================================
// 2-CTA variant (the DiT forward GEMMs): a cluster of two CTAs on one TPC computes a 256 x 256 output tile with
// tcgen05.mma.cta_group::2 (M = 256, N = 256, K = 16).  Each CTA owns 128 of the 256 rows (accumulator: its own TMEM, 2 x 256
// columns, double-buffered) and loads its 128 x 64 slice of A plus HALF of the 256 x 64 W tile: per SM the operand traffic of a
// k-block drops from 48 KB to 32 KB (smem fill, L2 reads) and the tensor core reads 64 instead of 96 bytes per clock from
// shared memory.  The leader CTA issues the MMAs for the pair; `full` barriers live in the leader (both CTAs' TMA loads signal
// them), `empty` / `tmem_full` are signalled in both CTAs by multicast commits, `tmem_empty` collects one arrival per epilogue
// warp of both CTAs.  The epilogue stages 32 x 64 bf16 blocks per warp in shared memory (128-byte swizzle, conflict-free) and
// writes them with TMA stores: full 128-byte lines, no per-thread row-strided stores.  Tiles are rasterised in supertiles of
// group_m x band_n tiles so that a W band and an A group stay L2-resident while they are being reused.
constexpr int kPairBM = 128;       // rows per CTA
constexpr int kPairTM = 256;       // rows per cluster tile
#ifndef FGB_PAIR_STAGES
#define FGB_PAIR_STAGES 5
#endif
constexpr int kPairStages = FGB_PAIR_STAGES;
constexpr int kPairABytes = kPairBM * kBK * 2;          // 16 KB
constexpr int kPairBBytesMax = (256 / 2) * kBK * 2;     // 16 KB: this CTA's half of a 256-column W tile
constexpr int kPairStageBytes = kPairABytes + kPairBBytesMax;
constexpr int kPairStoreBytes = 32 * 128;               // one staged block: 32 rows x 64 bf16
constexpr int kPairStagingBytes = 4 * 2 * kPairStoreBytes;   // 4 epilogue warps x 2 buffers
constexpr int kPairThreads = 192;
constexpr int kPairSmem = kPairStages * kPairStageBytes + kPairStagingBytes + 1024 /*align*/ + 256 /*barriers*/;

struct PairParams {
  const __nv_bfloat16* bias;
  __nv_bfloat16* c;
  const __nv_bfloat16* gate0;
  const __nv_bfloat16* gate1;
  int64_t ldc;
  int32_t m, n;
  int32_t rows_gate0;
  int32_t m_tiles, n_tiles, tiles, k_blocks;
  int32_t group_m, band_n;   // supertile: band_n tile columns x group_m tile rows, m fastest inside
  // SCATTER (fused Ulysses exchange of the q|k|v projection, fgb_gemm_qkv_scatter): column c of the [m, 3*dim] output belongs
  // to group c / dim (q, k, v) and head (c % dim) / 128; the 64-column block goes to the receive matrix of the peer that owns
  // the head — peer = head / hpr, columns (group*hpr + head % hpr)*128 + c % 128 — through that peer's tensor map (whose base is
  // already this rank's row window). rowsq[g*m + row] += sum of squares of the (bias-added, bf16-rounded) q / k row: the
  // receiver's RMSNorm needs the statistics of the FULL row, which no single rank holds after the head split.
  float* rowsq;
  int32_t dim, hpr;
  // Stream-K tail (fgb_gemm_bf16_sk; sk_tiles > 0): tiles [0, dp_tiles) are handed out whole, one per cluster and wave
  // (dp_tiles is a multiple of the cluster count); the K-blocks of the last sk_tiles tiles — what would be a partly filled
  // last wave — are cut into units of kSkUnit K-blocks and spread evenly over the first sk_clusters clusters, so that wave
  // takes sk_tiles / sk_clusters of a tile time instead of a whole one. The cluster whose range starts a tile owns it; the
  // others dump their fp32 partial accumulators into sk_partial (one slot per cluster) and raise a flag per epilogue warp.
  int32_t dp_tiles, sk_tiles, sk_clusters, sk_upt;   // sk_upt = units per tile
  // Half-width tail (BN = 256 only, exclusive with stream-K): when the last wave holds at most half as many tiles as there are
  // clusters, each of its hw_tiles tiles is handed out as TWO 256 x 128 items (N = 128 MMAs, 64-row W boxes), so that wave takes
  // ~0.6 of a tile time instead of a whole one — no partial sums, no workspace, no fix-up.
  int32_t hw_tiles;
  float* sk_partial;
  int32_t* sk_flags;
};

constexpr int kSkUnit = 4;           // K-blocks per stream-K unit (256 elements of K)
constexpr int kSkMaxSplit = 4;       // at most this many clusters per tail tile (bounds the owner's reduction)

struct PairItem {
  int32_t tile, kb0, kb1;
  int32_t role;        // 0 = whole tile or owner of a split tile (normal epilogue), 1 = contributor (partial dump)
  int32_t unit_end;    // owner: first unit after its own range; tile_end: first unit after the tile — the contributors are the
  int32_t tile_end;    // clusters after this one whose unit ranges start before tile_end
  int32_t n_off, bn;   // half-width tail items: column offset inside the tile and width (128); bn = 0: the kernel's full width
};

struct PeerMaps {
  CUtensorMap m[FGB_MAX_PEERS];
};

__host__ __device__ __forceinline__ void pair_tile_coords(const PairParams& p, int tile, int& mt, int& nt) {
  const int band_tiles = p.m_tiles * p.band_n;          // tiles of a full band
  const int band = tile / band_tiles;
  const int n0 = band * p.band_n;
  const int bn = min(p.n_tiles - n0, p.band_n);         // width of this band
  const int in_band = tile - band * band_tiles;
  const int group_tiles = p.group_m * bn;
  const int group = in_band / group_tiles;
  const int m0 = group * p.group_m;
  const int gm = min(p.m_tiles - m0, p.group_m);
  const int in_group = in_band - group * group_tiles;
  mt = m0 + in_group % gm;
  nt = n0 + in_group / gm;
}

__host__ __device__ __forceinline__ int sk_range_begin(const PairParams& p, int c) {
  return static_cast<int>(static_cast<int64_t>(c) * (p.sk_tiles * p.sk_upt) / p.sk_clusters);
}

// The it-th work item of `cluster` (all three roles of a CTA walk the same list): whole tiles first, then at most two
// stream-K segments — the end of one tail tile (contributor, or owner if the range starts on the tile boundary) and the
// beginning of the next (owner).
__host__ __device__ __forceinline__ bool pair_next_item(const PairParams& p, int cluster, int n_clusters, int it, PairItem& w) {
  const int tile = cluster + it * n_clusters;
  if (tile < p.dp_tiles) {
    w.tile = tile;
    w.kb0 = 0;
    w.kb1 = p.k_blocks;
    w.role = 0;
    w.unit_end = w.tile_end = 0;
    w.n_off = w.bn = 0;
    return true;
  }
  w.n_off = w.bn = 0;
  if (p.hw_tiles > 0) {
    const int h = tile - p.dp_tiles;
    if (h >= 2 * p.hw_tiles) return false;
    w.tile = p.dp_tiles + (h >> 1);
    w.kb0 = 0;
    w.kb1 = p.k_blocks;
    w.role = 0;
    w.unit_end = w.tile_end = 0;
    w.n_off = (h & 1) * 128;
    w.bn = 128;
    return true;
  }
  if (p.sk_tiles == 0 || cluster >= p.sk_clusters) return false;
  const int seg = it - p.dp_tiles / n_clusters;
  if (seg > 1) return false;
  const int u0 = sk_range_begin(p, cluster), u1 = sk_range_begin(p, cluster + 1);
  if (u0 == u1) return false;
  int j = u0 / p.sk_upt;
  int first = u0, last = min(u1, (j + 1) * p.sk_upt);
  if (seg == 1) {
    if (u1 <= (j + 1) * p.sk_upt) return false;
    ++j;
    first = j * p.sk_upt;
    last = u1;
  }
  w.tile = p.dp_tiles + j;
  w.kb0 = (first - j * p.sk_upt) * kSkUnit;
  w.kb1 = min((last - j * p.sk_upt) * kSkUnit, p.k_blocks);
  w.role = first == j * p.sk_upt ? 0 : 1;
  w.unit_end = last;
  w.tile_end = (j + 1) * p.sk_upt;
  return true;
}

__device__ __forceinline__ void st_release_gpu(int32_t* p, int32_t v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int32_t ld_acquire_gpu(const int32_t* p) {
  int32_t v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// BN = 256: the default. BN = 128: half-width tiles for problems whose 256-wide tile count leaves most of the last wave idle
// (e.g. N = 3072 on the 6820 rows of a Ulysses SP4 rank: 324 tiles = 4.38 waves of 74 clusters; 648 half tiles = 8.76).
template <int EPI, int BN, bool SCATTER>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPairThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_bh, const __grid_constant__ CUtensorMap tmap_c,
                 const __grid_constant__ PeerMaps peer_maps, const PairParams p) {
  constexpr int kPairBN = BN;
  constexpr int kPairBBytes = (BN / 2) * kBK * 2;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kPairStages * kPairABytes;
  uint8_t* smem_st = smem + kPairStages * kPairStageBytes;                    // 1024-byte aligned (stage sizes are multiples of 1 KB)
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_st + kPairStagingBytes);
  uint64_t* full = bars;                               // [stages]  TMA (both CTAs) -> MMA; used in the leader
  uint64_t* empty = bars + kPairStages;                // [stages]  MMA -> TMA, multicast to both CTAs
  uint64_t* tmem_full = bars + 2 * kPairStages;        // [2]       MMA -> epilogue, multicast to both CTAs
  uint64_t* tmem_empty = bars + 2 * kPairStages + 2;   // [2]       epilogue warps of both CTAs -> MMA; used in the leader
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPairStages + 4);
  uint64_t* sk_bar = bars + 2 * kPairStages + 6;       // [4 warps][2]  stream-K landing slots of the epilogue warps

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();             // 0 = leader
  const int n_clusters = gridDim.x >> 1;
  const int cluster = blockIdx.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    tma_prefetch_desc(&tmap_c);
    if (p.hw_tiles > 0) tma_prefetch_desc(&tmap_bh);
    for (int s = 0; s < kPairStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], 8);
    }
    for (int a = 0; a < 8; ++a) mbar_init(&sk_bar[a], 1);
    fence_mbar_init();
  }
  if (warp == 1) tmem_alloc_2sm<512>(tmem_slot);
  tc_fence_before();
  cluster_sync_all();    // barriers of BOTH CTAs are initialised before any remote signal can arrive
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (elect_one()) {
      // ------------------------------- TMA producer (both CTAs) -------------------------------
      uint32_t full_leader[kPairStages];
#pragma unroll
      for (int s = 0; s < kPairStages; ++s)
        asm volatile("mapa.shared::cluster.u32 %0, %1, 0;" : "=r"(full_leader[s]) : "r"(smem_u32(&full[s])));
      int stage = 0;
      uint32_t phase = 0;
      PairItem w;
      for (int it = 0; pair_next_item(p, cluster, n_clusters, it, w); ++it) {
        int mt, nt;
        pair_tile_coords(p, w.tile, mt, nt);
        const int row_a = mt * kPairTM + static_cast<int>(rank) * kPairBM;
        const int bn = w.bn ? w.bn : kPairBN;
        const int row_b = nt * kPairBN + w.n_off + static_cast<int>(rank) * (bn / 2);
        const CUtensorMap* mb = w.bn ? &tmap_bh : &tmap_b;          // 64-row boxes for a half-width item
        const uint32_t stage_tx = 2 * (kPairABytes + (bn / 2) * kBK * 2);
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait_cluster(&empty[stage], phase ^ 1);
          if (rank == 0) mbar_expect_tx(&full[stage], stage_tx);   // both CTAs' bytes land on the leader's barrier
          uint32_t fl = full_leader[0];
#pragma unroll
          for (int s = 1; s < kPairStages; ++s) fl = (stage == s) ? full_leader[s] : fl;
          tma_load_2d_2sm(smem_a + stage * kPairABytes, &tmap_a, fl, kb * kBK, row_a, kEvictNormal);
          tma_load_2d_2sm(smem_b + stage * kPairBBytes, mb, fl, kb * kBK, row_b, kEvictLast);
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0 && elect_one()) {
      // ------------------------------- MMA issuer (leader only) -------------------------------
      constexpr uint32_t idesc_full = make_idesc_bf16(kPairTM, kPairBN, 0, 0);
      constexpr uint32_t idesc_half = make_idesc_bf16(kPairTM, 128, 0, 0);
      const uint64_t a_desc0 = make_sdesc_sw128(smem_u32(smem_a), 16, 1024);
      const uint64_t b_desc0 = make_sdesc_sw128(smem_u32(smem_b), 16, 1024);
      int stage = 0;
      uint32_t phase = 0;
      PairItem w;
      for (int local = 0; pair_next_item(p, cluster, n_clusters, local, w); ++local) {
        const uint32_t idesc = w.bn ? idesc_half : idesc_full;
        const int acc = local & 1;
        const uint32_t acc_phase = (local >> 1) & 1;
        mbar_wait_cluster(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * kPairBN;
        for (int kb = w.kb0; kb < w.kb1; ++kb) {
          mbar_wait_cluster(&full[stage], phase);
          tc_fence_after();
          const uint64_t adesc = a_desc0 + static_cast<uint64_t>((stage * kPairABytes) >> 4);
          const uint64_t bdesc = b_desc0 + static_cast<uint64_t>((stage * kPairBBytes) >> 4);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_ss_2sm(tmem_d, adesc + static_cast<uint64_t>((k * 32) >> 4), bdesc + static_cast<uint64_t>((k * 32) >> 4), idesc,
                        ((kb - w.kb0) | k) != 0);
          tc_commit_2sm(&empty[stage], 0x3);     // the slot is reusable in BOTH CTAs once these MMAs have read it
          if (++stage == kPairStages) { stage = 0; phase ^= 1; }
        }
        tc_commit_2sm(&tmem_full[acc], 0x3);     // accumulator halves complete in both CTAs
      }
    }
  } else {
    // --------------------------------- epilogue (both CTAs) -----------------------------------
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32) are visible to this warp
    uint8_t* stage_buf = smem_st + quarter * (2 * kPairStoreBytes);
    int buf = 0;
    constexpr bool kReadsC = (EPI == FGB_EPI_GATED_RESIDUAL || EPI == FGB_EPI_RESIDUAL);
    // stream-K: this warp's [32 rows x BN] fp32 block inside a cluster's partial slot, as float4 [BN/32][8][32 lanes] (a warp
    // instruction moves 512 contiguous bytes; the owner's same-position thread reads back exactly what was written)
    constexpr int kSkWarpVec = (kPairBN / 32) * 8 * 32;               // float4 per warp block
    constexpr int kSkSlotVec = 8 * kSkWarpVec;                        // 2 CTAs x 4 warps
    const int sk_warp = static_cast<int>(rank) * 4 + quarter;
    PairItem w;
    for (int local = 0; pair_next_item(p, cluster, n_clusters, local, w); ++local) {
      int mt, nt;
      pair_tile_coords(p, w.tile, mt, nt);
      const int acc = local & 1;
      const uint32_t acc_phase = (local >> 1) & 1;
      mbar_wait_cluster(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * kPairBN;
      if (w.role == 1) {
        // contributor: fp32 partial accumulators -> this cluster's slot, then the flag of this warp (release)
        float4* dst = reinterpret_cast<float4*>(p.sk_partial) + static_cast<int64_t>(cluster) * kSkSlotVec + sk_warp * kSkWarpVec + lane;
#pragma unroll 1
        for (int c32 = 0; c32 < kPairBN / 32; ++c32) {
          uint32_t r[32];
          tmem_ld32(taddr + c32 * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 8; ++j)
            __stcg(dst + (c32 * 8 + j) * 32, make_float4(__uint_as_float(r[4 * j]), __uint_as_float(r[4 * j + 1]),
                                                         __uint_as_float(r[4 * j + 2]), __uint_as_float(r[4 * j + 3])));
        }
        tc_fence_before();
        __threadfence();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive_cluster(&tmem_empty[acc], 0);
          st_release_gpu(p.sk_flags + cluster * 8 + sk_warp, 1);
        }
        continue;
      }
      // owner of a split tile: add the contributors' partials (the clusters after this one whose unit ranges begin inside the
      // tile) into the accumulator before the normal epilogue. A split tile is always the LAST item of its owner, so the operand
      // ring is idle by now (every stage consumed: tmem_full fired) and serves as the landing zone: each warp pulls its
      // 32 x BN fp32 block of one contributor in two bulk copies (the copy engine keeps tens of KB in flight per warp, where
      // per-thread loads would pay one L2 round trip per 32 columns), double-buffered across (contributor, half).
      const bool split = w.tile_end > w.unit_end;
      if (split) {
        constexpr int kHalfBytes = kSkWarpVec * 16 / 2;
        uint8_t* land = smem + quarter * (2 * kHalfBytes);        // 4 warps x 2 x 16 KB (BN = 256) inside the 160 KB ring
        uint64_t* lbar = sk_bar + quarter * 2;
        auto next_contributor = [&](int c) {
          for (++c; c < p.sk_clusters; ++c) {
            const int b = sk_range_begin(p, c);
            if (b >= w.tile_end) return -1;
            if (sk_range_begin(p, c + 1) != b) return c;
          }
          return -1;
        };
        auto issue = [&](int c, int half, int slot) {           // lane 0 only
          const uint8_t* src = reinterpret_cast<const uint8_t*>(p.sk_partial) +
                               (static_cast<int64_t>(c) * kSkSlotVec + sk_warp * kSkWarpVec) * 16 + half * kHalfBytes;
          mbar_expect_tx(&lbar[slot], kHalfBytes);
          bulk_load(land + slot * kHalfBytes, src, kHalfBytes, &lbar[slot]);
        };
        auto wait_flag = [&](int c) {
          uint32_t spins = 0;
          while (ld_acquire_gpu(p.sk_flags + c * 8 + sk_warp) == 0) {
            if (++spins > FGB_SPIN_LIMIT) __trap();
          }
          fence_proxy_async_all();    // the partials were written through the generic proxy; the bulk copy reads through the async proxy
        };
        int c_issue = next_contributor(cluster), h_issue = 0, n_issued = 0, n_done = 0;
        int c_cur = c_issue;
        if (lane == 0) {
          wait_flag(c_issue);
          issue(c_issue, 0, 0);
          issue(c_issue, 1, 1);
        }
        n_issued = 2;
        c_issue = next_contributor(c_issue);
        while (c_cur >= 0) {
          for (int half = 0; half < 2; ++half, ++n_done) {
            const int slot = n_done & 1;
            mbar_wait(&lbar[slot], (n_done >> 1) & 1);
            const float4* src = reinterpret_cast<const float4*>(land + slot * kHalfBytes) + lane;
#pragma unroll 1
            for (int cc = 0; cc < kPairBN / 64; ++cc) {
              const int c32 = half * (kPairBN / 64) + cc;
              uint32_t r[32];
              tmem_ld32(taddr + c32 * 32, r);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 v = src[(cc * 8 + j) * 32];
                r[4 * j] = __float_as_uint(__uint_as_float(r[4 * j]) + v.x);
                r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + v.y);
                r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + v.z);
                r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + v.w);
              }
              tmem_st32(taddr + c32 * 32, r);
            }
            tmem_st_wait();
            __syncwarp();                 // every lane is done with this landing slot
            if (c_issue >= 0 && lane == 0) {
              if (h_issue == 0) wait_flag(c_issue);
              issue(c_issue, h_issue, slot);
            }
            if (c_issue >= 0) {
              ++n_issued;
              if (++h_issue == 2) {
                h_issue = 0;
                c_issue = next_contributor(c_issue);
              }
            }
          }
          c_cur = next_contributor(c_cur);
        }
      }
      const int row0 = mt * kPairTM + static_cast<int>(rank) * kPairBM + quarter * 32;   // first row of this warp's block
      const int row = row0 + lane;
      const bool row_ok = row < p.m;
      const __nv_bfloat16* crow = p.c + static_cast<int64_t>(row) * p.ldc;
      const __nv_bfloat16* gate = nullptr;
      if (EPI == FGB_EPI_GATED_RESIDUAL) gate = (row < p.rows_gate0) ? p.gate0 : p.gate1;
      uint4 xcur[4], xnext[4];
      const int bn = w.bn ? w.bn : kPairBN;          // columns of this item
      const int col_base = nt * kPairBN + w.n_off;   // its first column
      auto load_c = [&](int c32, uint4 (&dst)[4]) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int col = col_base + c32 * 32 + g * 8;
          dst[g] = (row_ok && col < p.n) ? *reinterpret_cast<const uint4*>(crow + col) : make_uint4(0, 0, 0, 0);
        }
      };
      if (kReadsC) load_c(0, xcur);
      float ss = 0.f;   // SCATTER: sum of squares of this row's outputs in this tile (a tile lies inside one of q / k / v)
#pragma unroll 1
      for (int c64 = 0; c64 < bn / 64; ++c64) {
        if (col_base + c64 * 64 >= p.n) break;   // warp-uniform
        uint8_t* sbuf = stage_buf + buf * kPairStoreBytes;
        // the TMA store that last read this buffer must be done with it (at most one newer store may still be in flight)
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int c32 = c64 * 2 + half;
          const int col0 = col_base + c32 * 32;
          uint32_t r[32];
          tmem_ld32(taddr + c32 * 32, r);
          if (kReadsC && c32 + 1 < bn / 32) load_c(c32 + 1, xnext);
          tmem_ld_wait();
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int col = col0 + g * 8;
            float y[8];
            uint4 bv = make_uint4(0, 0, 0, 0);
            if (p.bias && col < p.n) bv = __ldg(reinterpret_cast<const uint4*>(p.bias + col));
            const uint32_t bw[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              y[2 * i] = round_bf16(__uint_as_float(r[g * 8 + 2 * i]) + bf16_lo(bw[i]));
              y[2 * i + 1] = round_bf16(__uint_as_float(r[g * 8 + 2 * i + 1]) + bf16_hi(bw[i]));
            }
            if (EPI == FGB_EPI_BIAS_GELU_TANH) {
#pragma unroll
              for (int i = 0; i < 8; i += 2) gelu_tanh_2(y[i], y[i + 1]);
            }
            if (SCATTER) {
#pragma unroll
              for (int i = 0; i < 8; ++i) ss = fmaf(y[i], y[i], ss);
            }
            if (kReadsC) {
              const uint4 xv = xcur[g];
              const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w};
              if (EPI == FGB_EPI_GATED_RESIDUAL) {
                uint4 gv = make_uint4(0, 0, 0, 0);
                if (col < p.n) gv = __ldg(reinterpret_cast<const uint4*>(gate + col));
                const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  y[2 * i] = bf16_lo(xw[i]) + round_bf16(bf16_lo(gw[i]) * y[2 * i]);
                  y[2 * i + 1] = bf16_hi(xw[i]) + round_bf16(bf16_hi(gw[i]) * y[2 * i + 1]);
                }
              } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  y[2 * i] = bf16_lo(xw[i]) + y[2 * i];
                  y[2 * i + 1] = bf16_hi(xw[i]) + y[2 * i + 1];
                }
              }
            }
            uint4 o;
            o.x = pack_bf16(y[0], y[1]);
            o.y = pack_bf16(y[2], y[3]);
            o.z = pack_bf16(y[4], y[5]);
            o.w = pack_bf16(y[6], y[7]);
            // 128-byte swizzle of the staged block: 16-byte chunk index XOR (row & 7) — what the TMA store expects, and
            // it spreads the 32 lanes of one store instruction over all banks
            const int chunk = (half * 4 + g) ^ (lane & 7);
            *reinterpret_cast<uint4*>(sbuf + lane * 128 + chunk * 16) = o;
          }
          if (kReadsC) {
#pragma unroll
            for (int g = 0; g < 4; ++g) xcur[g] = xnext[g];
          }
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const int gcol = col_base + c64 * 64;
          if (SCATTER) {
            const int grp = gcol / p.dim, head = (gcol - grp * p.dim) >> 7;
            const int peer = head / p.hpr;
            tma_store_2d(&peer_maps.m[peer], sbuf, (grp * p.hpr + head - peer * p.hpr) * 128 + (gcol & 127), row0);
          } else {
            tma_store_2d(&tmap_c, sbuf, gcol, row0);   // rows >= m and columns >= n are clipped by the tensor map
          }
          tma_store_commit();
        }
        buf ^= 1;
      }
      if (SCATTER) {
        const int grp = col_base / p.dim;
        if (grp < 2 && row_ok) atomicAdd(p.rowsq + static_cast<int64_t>(grp) * p.m + row, ss);
      }
      // this warp has read its part of the accumulator: one arrival per warp on the leader's barrier
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_cluster(&tmem_empty[acc], 0);
        if (split) {   // the partials are consumed: lower the flags for the next launch that uses this workspace
          for (int c = cluster + 1; c < p.sk_clusters; ++c) {
            const int b = sk_range_begin(p, c);
            if (b >= w.tile_end) break;
            if (sk_range_begin(p, c + 1) == b) continue;
            p.sk_flags[c * 8 + sk_warp] = 0;
          }
        }
      }
    }
    if (lane == 0) tma_store_wait<0>();
  }

  __syncwarp();   // the single-thread roles rejoin their warps before the (aligned) barrier
  tc_fence_before();
  cluster_sync_all();   // the leader's MMAs read the peer's shared memory; nobody leaves before everything is consumed
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm<512>(tmem_base);
  }
}

template <int EPI, int BN, bool SCATTER = false>
static int launch_gemm_pair(fgb_ctx* ctx, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tc, const PairParams& p,
                            cudaStream_t stream, const PeerMaps* pm = nullptr, const CUtensorMap* tbh = nullptr) {
  auto kfn = gemm_pair_kernel<EPI, BN, SCATTER>;
  static PeerMaps no_peers{};
  static unsigned long long configured = 0;  // per template instance and device
  if (first_use_on_device(configured)) {
    FGB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, kPairSmem));
  }
  int clusters = ctx->sm_count / 2;
  if (p.sk_tiles == 0 && p.hw_tiles == 0 && p.tiles < clusters) clusters = p.tiles;   // split tails are laid out for the full cluster count
  kfn<<<2 * clusters, kPairThreads, kPairSmem, stream>>>(ta, tb, tbh ? *tbh : tb, tc, pm ? *pm : no_peers, p);
  FGB_LAUNCH_CHECK("gemm_pair_kernel");
  return FGB_OK;
}

// Supertile shape, from a sweep on the four GEMM shapes of the denoise step (tools/gemm_sweep.sh, profiles/r02_gemm_sweep*.log):
// bands of at most 14 tile columns (a 3.6 K-column W band: 22 MB at K = 3072) and SMALL row groups — 4 tile rows, 1 for the
// long-K FFN2 — so that the ~74 tiles in flight reuse each W tile a few times in quick succession while it is hot in L2 and the
// A rows they share stay small. FGB_GEMM_BAND_N / FGB_GEMM_GROUP_M override, for tuning.
static void pair_supertile(int m_tiles, int n_tiles, int k, int* group_m, int* band_n) {
  static int env_gm = -1, env_bn = -1;
  if (env_gm < 0) {
    const char* e = getenv("FGB_GEMM_GROUP_M");
    env_gm = e ? atoi(e) : 0;
    e = getenv("FGB_GEMM_BAND_N");
    env_bn = e ? atoi(e) : 0;
  }
  int bn = n_tiles < 14 ? n_tiles : 14;
  const int bands = (n_tiles + bn - 1) / bn;
  bn = (n_tiles + bands - 1) / bands;      // even out the bands (36 tile columns -> 3 bands of 12)
  int gm = k >= 8192 ? 1 : 4;     // long K: single tile rows (profiles/r02_gemm_sweep4.log: FFN2 1225 vs 1203 TFLOP/s for 1 vs 2 rows)
  if (gm > m_tiles) gm = m_tiles;
  *group_m = env_gm > 0 ? env_gm : gm;
  *band_n = env_bn > 0 ? env_bn : bn;
}

// Stream-K workspace: [flags: one int32 per (cluster, epilogue warp), padded to 4 KB][one fp32 256 x 256 slot per cluster].
constexpr int64_t kSkFlagBytes = 4096;
constexpr int64_t kSkSlotBytes = 256 * 256 * 4;
static int64_t sk_workspace_bytes(const fgb_ctx* ctx) { return kSkFlagBytes + static_cast<int64_t>(ctx->sm_count / 2) * kSkSlotBytes; }

// Decides whether the last, partly filled wave of a 2-CTA GEMM is cut along K (see PairParams). FGB_GEMM_SK=0 switches it off;
// FGB_GEMM_SK_MAXFRAC (default 0.9): a tail that fills more than this fraction of the clusters stays whole (nothing to gain);
// FGB_GEMM_SK_MIN_KB (default 96 K-blocks = K >= 6144): the fix-up (partial dump, flag, bulk pull, accumulate) costs ~15 us, and
// under the power cap a partly filled wave runs at a higher clock than a full one, so a split only pays when a tile is long —
// measured on the step's shapes at 6820 rows (profiles/r02_streamk_ab.log): FFN2 (K = 14 336) 13.85 -> 12.92 ms per 30
// launches, the K = 3072 projections 3.59 -> 3.96 ms.
static void pair_plan_streamk(fgb_ctx* ctx, PairParams& pp, void* ws, int64_t ws_bytes) {
  pp.dp_tiles = pp.tiles;
  pp.sk_tiles = pp.sk_clusters = 0;
  pp.sk_upt = 1;
  pp.sk_partial = nullptr;
  pp.sk_flags = nullptr;
  if (ctx->sk_min_kblocks < 0) {      // first use: defaults, overridable from the environment (A/B runs) or fgb_gemm_streamk_tune
    const char* e = getenv("FGB_GEMM_SK_MAXFRAC");
    ctx->sk_max_frac = e ? atof(e) : 0.9;
    e = getenv("FGB_GEMM_SK_MIN_KB");
    ctx->sk_min_kblocks = e ? atoi(e) : 96;
    e = getenv("FGB_GEMM_SK");
    ctx->sk_enabled = e ? atoi(e) : 1;
  }
  const int enabled = ctx->sk_enabled, min_kb = ctx->sk_min_kblocks;
  const double max_frac = ctx->sk_max_frac;
  const int clusters = ctx->sm_count / 2;
  if (!enabled || !ws || ws_bytes < sk_workspace_bytes(ctx) || clusters * 8 * 4 > kSkFlagBytes) return;
  const int tail = pp.tiles % clusters;
  if (tail == 0 || tail > max_frac * clusters || pp.k_blocks < 2 * kSkUnit || pp.k_blocks < min_kb) return;
  pp.dp_tiles = pp.tiles - tail;
  pp.sk_tiles = tail;
  pp.sk_upt = (pp.k_blocks + kSkUnit - 1) / kSkUnit;
  pp.sk_clusters = clusters < kSkMaxSplit * tail ? clusters : kSkMaxSplit * tail;
  pp.sk_flags = static_cast<int32_t*>(ws);
  pp.sk_partial = reinterpret_cast<float*>(static_cast<char*>(ws) + kSkFlagBytes);
}

// Half-width tail (see PairParams::hw_tiles): for 256-wide launches whose last wave is at most half full and is not already
// split along K. FGB_GEMM_HW=0 switches it off.
static void pair_plan_half_tail(const fgb_ctx* ctx, PairParams& pp, int bn) {
  pp.hw_tiles = 0;
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("FGB_GEMM_HW");
    enabled = e ? atoi(e) : 1;
  }
  const int clusters = ctx->sm_count / 2;
  if (!enabled || bn != 256 || pp.sk_tiles > 0 || pp.n % 256 != 0) return;
  const int tail = pp.tiles % clusters;
  if (tail == 0 || 2 * tail > clusters || pp.tiles < clusters) return;
  pp.dp_tiles = pp.tiles - tail;
  pp.hw_tiles = tail;
}

static int gemm_impl(fgb_ctx* ctx, const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* c, int64_t ldc,
                     int32_t m, int32_t n, int32_t k, int32_t epilogue, const void* gate0, const void* gate1, int32_t rows_gate0,
                     const void* a2, int64_t lda2, const void* w2, int64_t ldw2, int32_t k2, void* sk_ws, int64_t sk_ws_bytes,
                     void* stream);

// Everything about a 2-CTA launch that depends on the SHAPE only (tile width, tile counts, raster, stream-K / half-width tail):
// one place for gemm_impl and for the host-side schedule check (fgb_gemm_schedule_check). Returns the tile width.
static int pair_plan_shape(fgb_ctx* ctx, PairParams& pp, int m, int n, int k, void* sk_ws, int64_t sk_ws_bytes) {
  // tile width: 256 columns unless half-width tiles shorten the last, partly filled wave by more than they cost — a half-width
  // tile runs at ~0.85 of the full-width rate per FLOP (profiles/r02_gemm_sweep3.log: 1105 vs 1281 TFLOP/s at 6820 rows), so
  // this only pays for small problems (a handful of tiles); the Ulysses-rank shapes (324 tiles = 4.38 waves) stay at 256
  const int clusters = ctx->sm_count / 2;
  const int m_tiles = (m + kPairTM - 1) / kPairTM;
  const int waves256 = (m_tiles * ((n + 255) / 256) + clusters - 1) / clusters;
  const int waves128 = (m_tiles * ((n + 127) / 128) + clusters - 1) / clusters;
  static int env_bn = -1;
  if (env_bn < 0) {
    const char* e = getenv("FGB_GEMM_BN");
    env_bn = e ? atoi(e) : 0;
  }
  const int bn = env_bn ? env_bn : ((0.5 * 1.3 * waves128 < waves256) ? 128 : 256);
  pp.m = m;
  pp.n = n;
  pp.m_tiles = m_tiles;
  pp.n_tiles = (n + bn - 1) / bn;
  pp.tiles = pp.m_tiles * pp.n_tiles;
  pp.k_blocks = (k + kBK - 1) / kBK;
  pair_supertile(pp.m_tiles, (pp.n_tiles * bn + 255) / 256, k, &pp.group_m, &pp.band_n);
  pp.band_n = pp.band_n * 256 / bn;      // the band is sized in columns
  pair_plan_streamk(ctx, pp, sk_ws, sk_ws_bytes);
  pair_plan_half_tail(ctx, pp, bn);
  return bn;
}

}  // namespace fgb

extern "C" int64_t fgb_gemm_workspace_bytes(fgb_ctx* ctx) { return ctx ? fgb::sk_workspace_bytes(ctx) : 0; }

extern "C" int fgb_gemm_streamk_tune(fgb_ctx* ctx, int32_t min_k, double max_tail_frac) {
  FGB_CHECK_ARG(ctx && min_k >= 0 && max_tail_frac >= 0.0 && max_tail_frac <= 1.0, "fgb_gemm_streamk_tune: bad argument");
  const char* e = getenv("FGB_GEMM_SK");
  ctx->sk_enabled = e ? atoi(e) : 1;
  ctx->sk_min_kblocks = (min_k + fgb::kBK - 1) / fgb::kBK;
  ctx->sk_max_frac = max_tail_frac;
  return FGB_OK;
}

extern "C" int fgb_gemm_bf16_sk(fgb_ctx* ctx, const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* c,
                                int64_t ldc, int32_t m, int32_t n, int32_t k, int32_t epilogue, const void* gate0, const void* gate1,
                                int32_t rows_gate0, void* workspace, int64_t workspace_bytes, void* stream) {
  FGB_CHECK_ARG(!workspace || fgb::aligned16(workspace), "fgb_gemm_bf16_sk: workspace must be 16-byte aligned");
  return fgb::gemm_impl(ctx, a, lda, w, ldw, bias, c, ldc, m, n, k, epilogue, gate0, gate1, rows_gate0, nullptr, 0, nullptr, 0, 0,
                        workspace, workspace_bytes, stream);
}

// Host-side walk of the work list of one 2-CTA GEMM launch (no device needed): every (tile, K-block, column half) must be
// computed exactly once; a split tile has exactly one owner, the owner's item is its cluster's last, and the contributors the
// owner will wait for are exactly the clusters that dump a partial for that tile; a cluster contributes at most once.
extern "C" int fgb_gemm_schedule_check(int32_t m, int32_t n, int32_t k, int32_t sm_count, int32_t with_workspace, int32_t min_k,
                                       int32_t* n_split_tiles, int32_t* n_half_items) {
  using namespace fgb;
  if (m < kPairTM || n <= 0 || k <= 0 || sm_count < 2) return set_error(FGB_ERR_INVALID, "fgb_gemm_schedule_check: bad shape");
  fgb_ctx ctx;
  ctx.sm_count = sm_count;
  ctx.sk_enabled = 1;
  ctx.sk_min_kblocks = (min_k + kBK - 1) / kBK;
  ctx.sk_max_frac = 0.9;
  PairParams pp{};
  static char fake_ws[16];
  const int bn = pair_plan_shape(&ctx, pp, m, n, k, with_workspace ? fake_ws : nullptr, with_workspace ? sk_workspace_bytes(&ctx) : 0);
  int clusters = sm_count / 2;
  if (pp.sk_tiles == 0 && pp.hw_tiles == 0 && pp.tiles < clusters) clusters = pp.tiles;
  const int halves = bn / 128;     // coverage is counted per 128-column half
  std::vector<int> cover(static_cast<size_t>(pp.tiles) * pp.k_blocks * halves, 0);
  std::vector<int> owner(pp.tiles, -1), contrib_of_cluster(clusters, -1);
  std::vector<std::vector<int>> contributors(pp.tiles);
  int splits = 0, half_items = 0;
  for (int c = 0; c < clusters; ++c) {
    PairItem w;
    int n_items = 0;
    for (int it = 0; pair_next_item(pp, c, clusters, it, w); ++it) ++n_items;
    for (int it = 0; pair_next_item(pp, c, clusters, it, w); ++it) {
      if (w.tile < 0 || w.tile >= pp.tiles || w.kb0 < 0 || w.kb1 > pp.k_blocks || w.kb0 >= w.kb1)
        return set_error(FGB_ERR_INVALID, "schedule: cluster %d item %d out of range (tile %d kb [%d,%d))", c, it, w.tile, w.kb0, w.kb1);
      int mt, nt;
      pair_tile_coords(pp, w.tile, mt, nt);
      if (mt < 0 || mt >= pp.m_tiles || nt < 0 || nt >= pp.n_tiles) return set_error(FGB_ERR_INVALID, "schedule: tile %d -> (%d,%d)", w.tile, mt, nt);
      const int h0 = w.bn ? w.n_off / 128 : 0, h1 = w.bn ? h0 + w.bn / 128 : halves;
      if (w.bn) ++half_items;
      for (int kb = w.kb0; kb < w.kb1; ++kb)
        for (int h = h0; h < h1; ++h) ++cover[(static_cast<size_t>(w.tile) * pp.k_blocks + kb) * halves + h];
      const bool partial = w.kb0 > 0 || w.kb1 < pp.k_blocks;
      if (w.role == 1) {
        if (w.kb0 == 0) return set_error(FGB_ERR_INVALID, "schedule: contributor of tile %d starts at K-block 0", w.tile);
        if (contrib_of_cluster[c] >= 0) return set_error(FGB_ERR_INVALID, "schedule: cluster %d contributes twice", c);
        contrib_of_cluster[c] = w.tile;
        contributors[w.tile].push_back(c);
      } else {
        if (w.kb0 != 0 && partial) return set_error(FGB_ERR_INVALID, "schedule: owner of tile %d starts at K-block %d", w.tile, w.kb0);
        if (!w.bn) {
          if (owner[w.tile] >= 0) return set_error(FGB_ERR_INVALID, "schedule: tile %d has two owners", w.tile);
          owner[w.tile] = c;
        }
        if (w.tile_end > w.unit_end) {     // split: must be the cluster's last item, and the wait list must match the dumps
          ++splits;
          if (it != n_items - 1) return set_error(FGB_ERR_INVALID, "schedule: split tile %d is not the last item of cluster %d", w.tile, c);
        }
      }
    }
  }
  for (size_t i = 0; i < cover.size(); ++i)
    if (cover[i] != 1) return set_error(FGB_ERR_INVALID, "schedule: (tile, K-block, half) %zu computed %d times (m=%d n=%d k=%d)", i, cover[i], m, n, k);
  // the owner's wait list, computed the way the kernel does, against the clusters that really dump a partial for the tile
  for (int t = 0; t < pp.tiles; ++t) {
    if (contributors[t].empty()) continue;
    const int o = owner[t];
    if (o < 0) return set_error(FGB_ERR_INVALID, "schedule: split tile %d has no owner", t);
    const int tile_end = (t - pp.dp_tiles + 1) * pp.sk_upt;
    std::vector<int> waits;
    for (int c = o + 1; c < pp.sk_clusters; ++c) {
      const int b = sk_range_begin(pp, c);
      if (b >= tile_end) break;
      if (sk_range_begin(pp, c + 1) == b) continue;
      waits.push_back(c);
    }
    std::vector<int> dumps = contributors[t];
    std::sort(dumps.begin(), dumps.end());
    if (waits != dumps) return set_error(FGB_ERR_INVALID, "schedule: tile %d: owner %d waits for %zu clusters, %zu dump", t, o, waits.size(), dumps.size());
    if (static_cast<int>(dumps.size()) > kSkMaxSplit) return set_error(FGB_ERR_INVALID, "schedule: tile %d split %zu ways", t, dumps.size() + 1);
  }
  if (n_split_tiles) *n_split_tiles = splits;
  if (n_half_items) *n_half_items = half_items;
  return FGB_OK;
}

extern "C" int fgb_gemm_bf16(fgb_ctx* ctx, const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias,
                             void* c, int64_t ldc, int32_t m, int32_t n, int32_t k, int32_t epilogue,
                             const void* gate0, const void* gate1, int32_t rows_gate0, void* stream) {
  return fgb_gemm_bf16_ex(ctx, a, lda, w, ldw, bias, c, ldc, m, n, k, epilogue, gate0, gate1, rows_gate0, nullptr, 0, nullptr, 0, 0,
                          stream);
}

extern "C" int fgb_gemm_bf16_ex(fgb_ctx* ctx, const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias,
                                void* c, int64_t ldc, int32_t m, int32_t n, int32_t k, int32_t epilogue, const void* gate0,
                                const void* gate1, int32_t rows_gate0, const void* a2, int64_t lda2, const void* w2,
                                int64_t ldw2, int32_t k2, void* stream) {
  return fgb::gemm_impl(ctx, a, lda, w, ldw, bias, c, ldc, m, n, k, epilogue, gate0, gate1, rows_gate0, a2, lda2, w2, ldw2, k2, nullptr, 0,
                        stream);
}

static int fgb::gemm_impl(fgb_ctx* ctx, const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, void* c, int64_t ldc,
                          int32_t m, int32_t n, int32_t k, int32_t epilogue, const void* gate0, const void* gate1, int32_t rows_gate0,
                          const void* a2, int64_t lda2, const void* w2, int64_t ldw2, int32_t k2, void* sk_ws, int64_t sk_ws_bytes,
                          void* stream) {
  using namespace fgb;
  FGB_CHECK_ARG(ctx, "fgb_gemm_bf16: ctx is NULL");
  FGB_CHECK_ARG(a && w && c, "fgb_gemm_bf16: NULL matrix pointer");
  FGB_CHECK_ARG(m > 0 && n > 0 && k > 0, "fgb_gemm_bf16: empty problem m=%d n=%d k=%d", m, n, k);
  FGB_CHECK_ARG(n % 8 == 0 && k % 8 == 0, "fgb_gemm_bf16: n=%d and k=%d must be multiples of 8", n, k);
  FGB_CHECK_ARG(lda >= k && ldw >= k && ldc >= n, "fgb_gemm_bf16: leading dimension too small");
  FGB_CHECK_ARG(ldc % 8 == 0 && aligned16(c), "fgb_gemm_bf16: C must be 16-byte aligned with ldc %% 8 == 0");
  FGB_CHECK_ARG(!bias || aligned16(bias), "fgb_gemm_bf16: bias must be 16-byte aligned");
  FGB_CHECK_ARG(epilogue >= 0 && epilogue <= 3, "fgb_gemm_bf16: unknown epilogue %d", epilogue);
  if (epilogue == FGB_EPI_GATED_RESIDUAL)
    FGB_CHECK_ARG(gate0 && gate1 && aligned16(gate0) && aligned16(gate1),
                  "fgb_gemm_bf16: gated residual needs 16-byte aligned gate0/gate1");

  CUtensorMap ta, tb, ta2, tb2;
  int rc;
  // the 2-CTA kernel takes the plain forward GEMMs with at least one full cluster tile of rows; everything else (tiny M: the
  // time / text embeddings; the LoRA K-extension of the trainer) stays on the 1-CTA kernel
  static int pair_enabled = -1;
  if (pair_enabled < 0) {
    const char* e = getenv("FGB_GEMM_PAIR");
    pair_enabled = e ? atoi(e) : 1;
  }
  if (pair_enabled && k2 == 0 && m >= kPairTM && ctx->sm_count >= 2) {
    CUtensorMap tc;
    if ((rc = make_tmap_bf16_2d(ctx, &ta, a, m, k, lda, kPairBM))) return rc;
    PairParams pp;
    const int bn = pair_plan_shape(ctx, pp, m, n, k, sk_ws, sk_ws_bytes);
    if ((rc = make_tmap_bf16_2d(ctx, &tb, w, n, k, ldw, bn / 2))) return rc;
    if ((rc = make_tmap_bf16_2d(ctx, &tc, c, m, n, ldc, 32))) return rc;
    pp.bias = static_cast<const __nv_bfloat16*>(bias);
    pp.c = static_cast<__nv_bfloat16*>(c);
    pp.gate0 = static_cast<const __nv_bfloat16*>(gate0);
    pp.gate1 = static_cast<const __nv_bfloat16*>(gate1);
    pp.ldc = ldc;
    pp.rows_gate0 = rows_gate0;
    pp.rowsq = nullptr;
    pp.dim = pp.hpr = 0;
    CUtensorMap tbh = tb;
    if (pp.hw_tiles > 0 && (rc = make_tmap_bf16_2d(ctx, &tbh, w, n, k, ldw, 64))) return rc;
    cudaStream_t ps = static_cast<cudaStream_t>(stream);
#define FGB_PAIR_CASE(E)                                                       \
  case E:                                                                      \
    return bn == 128 ? launch_gemm_pair<E, 128>(ctx, ta, tb, tc, pp, ps) : launch_gemm_pair<E, 256>(ctx, ta, tb, tc, pp, ps, nullptr, &tbh);
    switch (epilogue) {
      FGB_PAIR_CASE(FGB_EPI_BIAS)
      FGB_PAIR_CASE(FGB_EPI_BIAS_GELU_TANH)
      FGB_PAIR_CASE(FGB_EPI_GATED_RESIDUAL)
      default:
        return bn == 128 ? launch_gemm_pair<FGB_EPI_RESIDUAL, 128>(ctx, ta, tb, tc, pp, ps)
                         : launch_gemm_pair<FGB_EPI_RESIDUAL, 256>(ctx, ta, tb, tc, pp, ps, nullptr, &tbh);
    }
#undef FGB_PAIR_CASE
  }
  rc = make_tmap_bf16_2d(ctx, &ta, a, m, k, lda, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(ctx, &tb, w, n, k, ldw, kBN);
  if (rc) return rc;
  if (k2 > 0) {
    FGB_CHECK_ARG(a2 && w2 && k2 % 8 == 0 && lda2 >= k2 && ldw2 >= k2, "fgb_gemm_bf16_ex: bad second operand pair (k2=%d)", k2);
    if ((rc = make_tmap_bf16_2d(ctx, &ta2, a2, m, k2, lda2, kBM))) return rc;
    if ((rc = make_tmap_bf16_2d(ctx, &tb2, w2, n, k2, ldw2, kBN))) return rc;
  }

  GemmParams p;
  p.tap_kblocks = 0;
  p.a_row0 = 0;
  p.grid_h = p.grid_w = 0;
  p.k2_blocks = k2 > 0 ? (k2 + kBK - 1) / kBK : 0;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.c = static_cast<__nv_bfloat16*>(c);
  p.gate0 = static_cast<const __nv_bfloat16*>(gate0);
  p.gate1 = static_cast<const __nv_bfloat16*>(gate1);
  p.ldc = ldc;
  p.m = m;
  p.n = n;
  p.k = k;
  p.rows_gate0 = rows_gate0;
  p.m_blocks = (m + kBM - 1) / kBM;
  p.n_blocks = (n + kBN - 1) / kBN;
  p.tiles = p.m_blocks * p.n_blocks;
  p.k_blocks = (k + kBK - 1) / kBK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const CUtensorMap* pa2 = k2 > 0 ? &ta2 : nullptr;
  const CUtensorMap* pb2 = k2 > 0 ? &tb2 : nullptr;
  switch (epilogue) {
    case FGB_EPI_BIAS: return launch_gemm<FGB_EPI_BIAS>(ctx, ta, tb, p, s, pa2, pb2);
    case FGB_EPI_BIAS_GELU_TANH: return launch_gemm<FGB_EPI_BIAS_GELU_TANH>(ctx, ta, tb, p, s, pa2, pb2);
    case FGB_EPI_GATED_RESIDUAL: return launch_gemm<FGB_EPI_GATED_RESIDUAL>(ctx, ta, tb, p, s, pa2, pb2);
    default: return launch_gemm<FGB_EPI_RESIDUAL>(ctx, ta, tb, p, s, pa2, pb2);
  }
}

extern "C" int fgb_gemm_qkv_scatter(fgb_ctx* ctx, const void* a, int64_t lda, const void* w, int64_t ldw, const void* bias, int32_t m,
                                    int32_t dim, int32_t k, void* const* peer_recv, int32_t world, int32_t rank, float* rowsq,
                                    void* workspace, int64_t workspace_bytes, void* stream) {
  using namespace fgb;
  FGB_CHECK_ARG(ctx && a && w && peer_recv && rowsq, "fgb_gemm_qkv_scatter: NULL argument");
  FGB_CHECK_ARG(m > 0 && k > 0 && k % 8 == 0 && dim > 0 && dim % 256 == 0 && lda >= k && ldw >= k,
                "fgb_gemm_qkv_scatter: m=%d dim=%d (multiple of 256) k=%d", m, dim, k);
  const int heads = dim / 128;
  FGB_CHECK_ARG(world > 0 && world <= FGB_MAX_PEERS && heads % world == 0 && rank >= 0 && rank < world,
                "fgb_gemm_qkv_scatter: heads=%d world=%d rank=%d", heads, world, rank);
  FGB_CHECK_ARG(!bias || aligned16(bias), "fgb_gemm_qkv_scatter: bias must be 16-byte aligned");
  FGB_CHECK_ARG(ctx->sm_count >= 2, "fgb_gemm_qkv_scatter: needs CTA pairs");
  const int hpr = heads / world;
  const int64_t ld_recv = static_cast<int64_t>(3) * hpr * 128;
  const int n = 3 * dim;
  CUtensorMap ta, tb;
  PeerMaps pm{};
  int rc;
  if ((rc = make_tmap_bf16_2d(ctx, &ta, a, m, k, lda, kPairBM))) return rc;
  if ((rc = make_tmap_bf16_2d(ctx, &tb, w, n, k, ldw, 128))) return rc;
  for (int q = 0; q < world; ++q) {
    FGB_CHECK_ARG(peer_recv[q] && aligned16(peer_recv[q]), "fgb_gemm_qkv_scatter: peer receive matrix %d", q);
    // this rank's row window [rank*m, rank*m + m) of peer q's receive matrix: the map clips the padded tile rows at m
    const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(peer_recv[q]) + static_cast<int64_t>(rank) * m * ld_recv;
    if ((rc = make_tmap_bf16_2d(ctx, &pm.m[q], base, m, ld_recv, ld_recv, 32))) return rc;
  }
  for (int q = world; q < FGB_MAX_PEERS; ++q) pm.m[q] = pm.m[0];
  PairParams pp;
  pp.bias = static_cast<const __nv_bfloat16*>(bias);
  pp.c = nullptr;
  pp.gate0 = pp.gate1 = nullptr;
  pp.ldc = 0;
  pp.m = m;
  pp.n = n;
  pp.rows_gate0 = 0;
  pp.rowsq = rowsq;
  pp.dim = dim;
  pp.hpr = hpr;
  pp.m_tiles = (m + kPairTM - 1) / kPairTM;
  pp.n_tiles = (n + 255) / 256;
  pp.tiles = pp.m_tiles * pp.n_tiles;
  pp.k_blocks = (k + kBK - 1) / kBK;
  pair_supertile(pp.m_tiles, pp.n_tiles, k, &pp.group_m, &pp.band_n);
  FGB_CHECK_ARG(!workspace || aligned16(workspace), "fgb_gemm_qkv_scatter: workspace must be 16-byte aligned");
  pair_plan_streamk(ctx, pp, workspace, workspace_bytes);
  pair_plan_half_tail(ctx, pp, 256);
  CUtensorMap tbh = tb;
  if (pp.hw_tiles > 0 && (rc = make_tmap_bf16_2d(ctx, &tbh, w, n, k, ldw, 64))) return rc;
  return launch_gemm_pair<FGB_EPI_BIAS, 256, true>(ctx, ta, tb, pm.m[0], pp, static_cast<cudaStream_t>(stream), &pm, &tbh);
}

extern "C" int fgb_gemm_dgrad(fgb_ctx* ctx, const void* dy, int64_t ld_dy, const void* w, int64_t ldw, void* dx, int64_t ld_dx,
                              int32_t m, int32_t n_in, int32_t k_out, void* stream) {
  return fgb_gemm_dgrad_ex(ctx, dy, ld_dy, w, ldw, dx, ld_dx, m, n_in, k_out, nullptr, 0, nullptr, 0, 0, stream);
}

extern "C" int fgb_gemm_dgrad_ex(fgb_ctx* ctx, const void* dy, int64_t ld_dy, const void* w, int64_t ldw, void* dx, int64_t ld_dx,
                                 int32_t m, int32_t n_in, int32_t k_out, const void* u, int64_t ld_u, const void* a1, int64_t ld_a1,
                                 int32_t k2, void* stream) {
  using namespace fgb;
  FGB_CHECK_ARG(ctx, "fgb_gemm_dgrad: ctx is NULL");
  FGB_CHECK_ARG(dy && w && dx, "fgb_gemm_dgrad: NULL matrix pointer");
  FGB_CHECK_ARG(m > 0 && n_in > 0 && k_out > 0, "fgb_gemm_dgrad: empty problem m=%d n_in=%d k_out=%d", m, n_in, k_out);
  FGB_CHECK_ARG(n_in % 8 == 0 && k_out % 8 == 0, "fgb_gemm_dgrad: n_in=%d and k_out=%d must be multiples of 8", n_in, k_out);
  FGB_CHECK_ARG(ld_dy >= k_out && ldw >= n_in && ld_dx >= n_in, "fgb_gemm_dgrad: leading dimension too small");
  FGB_CHECK_ARG(ld_dx % 8 == 0 && aligned16(dx), "fgb_gemm_dgrad: dx must be 16-byte aligned with ld_dx %% 8 == 0");
  CUtensorMap ta, tb, ta2, tb2;
  int rc = make_tmap_bf16_2d(ctx, &ta, dy, m, k_out, ld_dy, kBM);
  if (rc) return rc;
  rc = make_tmap_bf16_2d(ctx, &tb, w, k_out, n_in, ldw, kBK);   // box = [64 k-rows][64 n-cols]
  if (rc) return rc;
  if (k2 > 0) {
    FGB_CHECK_ARG(u && a1 && k2 % 8 == 0 && ld_u >= k2 && ld_a1 >= n_in, "fgb_gemm_dgrad_ex: bad second operand pair (k2=%d)", k2);
    if ((rc = make_tmap_bf16_2d(ctx, &ta2, u, m, k2, ld_u, kBM))) return rc;
    if ((rc = make_tmap_bf16_2d(ctx, &tb2, a1, k2, n_in, ld_a1, kBK))) return rc;
  }
  GemmParams p;
  p.tap_kblocks = 0;
  p.a_row0 = 0;
  p.grid_h = p.grid_w = 0;
  p.k2_blocks = k2 > 0 ? (k2 + kBK - 1) / kBK : 0;
  p.bias = nullptr;
  p.c = static_cast<__nv_bfloat16*>(dx);
  p.gate0 = p.gate1 = nullptr;
  p.ldc = ld_dx;
  p.m = m;
  p.n = n_in;
  p.k = k_out;
  p.rows_gate0 = 0;
  p.m_blocks = (m + kBM - 1) / kBM;
  p.n_blocks = (n_in + kBN - 1) / kBN;
  p.tiles = p.m_blocks * p.n_blocks;
  p.k_blocks = (k_out + kBK - 1) / kBK;
  return launch_gemm<FGB_EPI_BIAS, true>(ctx, ta, tb, p, static_cast<cudaStream_t>(stream), k2 > 0 ? &ta2 : nullptr,
                                         k2 > 0 ? &tb2 : nullptr);
}

extern "C" int fgb_conv_taps_bf16(fgb_ctx* ctx, const void* x, int64_t ldx, int64_t x_rows, int64_t a_row0, const void* w, int64_t ldw,
                                  const void* bias, void* out, int64_t ldo, int32_t m, int32_t n, int32_t cin, int32_t taps,
                                  const int32_t* tap_offsets, int32_t grid_h, int32_t grid_w, int32_t epilogue, void* stream) {
  using namespace fgb;
  FGB_CHECK_ARG(ctx && x && w && out && tap_offsets, "fgb_conv_taps_bf16: NULL argument");
  FGB_CHECK_ARG(m > 0 && n > 0 && n % 8 == 0 && cin > 0 && cin % kBK == 0 && taps >= 1 && taps <= 27,
                "fgb_conv_taps_bf16: m=%d n=%d cin=%d taps=%d (n %% 8, cin %% 64, 1 <= taps <= 27)", m, n, cin, taps);
  FGB_CHECK_ARG(x_rows > 0 && ldx >= cin && ldw >= static_cast<int64_t>(taps) * cin && ldo >= n && ldo % 8 == 0 && aligned16(out) &&
                    (!bias || aligned16(bias)), "fgb_conv_taps_bf16: leading dimensions / alignment");
  FGB_CHECK_ARG(epilogue == FGB_EPI_BIAS || epilogue == FGB_EPI_RESIDUAL, "fgb_conv_taps_bf16: epilogue must be BIAS or RESIDUAL");
  FGB_CHECK_ARG((grid_h == 0 && grid_w == 0) || (grid_h >= 3 && grid_w >= 3), "fgb_conv_taps_bf16: grid %d x %d", grid_h, grid_w);
  FGB_CHECK_ARG(a_row0 > -(1ll << 30) && a_row0 < (1ll << 30) && x_rows < (1ll << 31), "fgb_conv_taps_bf16: row range");
  CUtensorMap ta, tb;
  int rc = make_tmap_bf16_2d(ctx, &ta, x, x_rows, cin, ldx, kBM);
  if (rc) return rc;
  if ((rc = make_tmap_bf16_2d(ctx, &tb, w, n, static_cast<int64_t>(taps) * cin, ldw, kBN))) return rc;
  GemmParams p;
  p.tap_kblocks = cin / kBK;
  p.a_row0 = static_cast<int32_t>(a_row0);
  p.grid_h = grid_h;
  p.grid_w = grid_w;
  for (int i = 0; i < 27; ++i) p.tap_off[i] = i < taps ? tap_offsets[i] : 0;
  p.k2_blocks = 0;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.c = static_cast<__nv_bfloat16*>(out);
  p.gate0 = p.gate1 = nullptr;
  p.ldc = ldo;
  p.m = m;
  p.n = n;
  p.k = taps * cin;
  p.rows_gate0 = 0;
  p.m_blocks = (m + kBM - 1) / kBM;
  p.n_blocks = (n + kBN - 1) / kBN;
  p.tiles = p.m_blocks * p.n_blocks;
  p.k_blocks = taps * p.tap_kblocks;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (epilogue == FGB_EPI_RESIDUAL) return launch_gemm<FGB_EPI_RESIDUAL, false, true>(ctx, ta, tb, p, s);
  return launch_gemm<FGB_EPI_BIAS, false, true>(ctx, ta, tb, p, s);
}
