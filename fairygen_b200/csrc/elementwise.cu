// fairygen_b200 — fused memory-bound kernels of the DiT block (sm_100a).
//
// These replace chains of ATen elementwise/reduction kernels in the reference
// (animation/diffsynth/models/wan_video_dit.py, "DIT"; pipelines/wan_video.py, "PIPE";
// diffusion/flow_match.py, "FM") with one pass over HBM each: every row is read once with 16-byte
// loads (12 in flight per lane at D = 3072), reduced with warp shuffles, and written once.
// Arithmetic is fp32 with the reference's bf16 rounding points reproduced in registers.
#include <cstdlib>

#include "common.cuh"
#include "host.h"

namespace fgb {

constexpr int kRowWarps = 4;  // rows per CTA (one warp per row)

struct PeerPtrs {
  void* p[FGB_MAX_PEERS];
};
// Where the Ulysses exchange wants a row: peers.p == all NULL -> in place. Otherwise the 128-wide head `h` of local
// row r goes to peer h / (heads/world), row (rank*s_local + r), group `grp` of that peer's receive matrix
// [world*s_local][groups][heads/world][128] (see sp_scatter_heads_kernel).
struct ScatterSpec {
  PeerPtrs peers;
  int s_local, heads, world, rank, grp, groups;
};

// Squared norm of one 128-wide head = the sum of `ss` over the 16 lanes that hold it. (A one-instruction variant —
// redux.sync.add on a rounded-up fixed-point image — was measured SLOWER than these four shuffle + add steps on sm_100a:
// qk_norm_rope 130 -> 161 us alone, profiles/r02_attn_head_bound_ab.log.)
__device__ __forceinline__ float head_ss(float ss) {
#pragma unroll
  for (int off = 8; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  return ss;
}

// LayerNorm arithmetic shared by the one-row-per-warp kernel and the row-streaming kernel (same bits from both). The kernels
// are issue-bound at the capped in-step clock, so statistics and normalisation run as packed fp32x2 operations (two elements per
// FADD2 / FFMA2): each lane's result is the IEEE result of the scalar operation; only the order of the partial sums differs
// from a single sequential accumulator (even and odd elements are summed separately).
template <int NV>
__device__ __forceinline__ void ln_row_stats(const uint4 (&v)[NV], float eps, float& mean, float& rstd) {
  constexpr int D = NV * 256;
  // One pass for both moments, shifted by the row's first element K so that E[(x-K)^2] - E[x-K]^2 does not cancel:
  // mean = K + s/D, var = q/D - (s/D)^2.
  const float K = __shfl_sync(0xffffffffu, bf16_lo(v[0].x), 0);
  const uint64_t nk2 = pack2(-K, -K);
  uint64_t s2 = pack2(0.f, 0.f), q2 = pack2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t d2 = add2(pack2(bf16_lo(w[j]), bf16_hi(w[j])), nk2);
      s2 = add2(s2, d2);
      q2 = fma2(d2, d2, q2);
    }
  }
  float sa, sb, qa, qb;
  unpack2(s2, sa, sb);
  unpack2(q2, qa, qb);
  const float s1 = warp_sum(sa + sb) * (1.0f / D);
  const float q1 = warp_sum(qa + qb) * (1.0f / D);
  mean = K + s1;
  rstd = rsqrtf(fmaxf(q1 - s1 * s1, 0.f) + eps);
}

// One 16-byte vector of the row: normalise and apply (weight, bias) [AFFINE: F.layer_norm, one rounding] or the adaLN
// modulation row (scale, shift) with the reference's bf16 rounding points (DIT:63-64).
template <bool AFFINE>
__device__ __forceinline__ uint4 ln_apply(const uint4& xv, const uint4& av, const uint4& bv, float rstd, float nmr) {
  const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, aw[4] = {av.x, av.y, av.z, av.w}, bw[4] = {bv.x, bv.y, bv.z, bv.w};
  const uint64_t r2 = pack2(rstd, rstd), n2 = pack2(nmr, nmr);
  uint32_t o[4];
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    uint64_t t2 = fma2(pack2(bf16_lo(xw[w]), bf16_hi(xw[w])), r2, n2);
    if (AFFINE) t2 = fma2(t2, pack2(bf16_lo(aw[w]), bf16_hi(aw[w])), pack2(bf16_lo(bw[w]), bf16_hi(bw[w])));
    float lo, hi;
    unpack2(t2, lo, hi);
    const uint32_t n = pack_bf16(lo, hi);
    // norm(x) -> bf16 ; (1 + scale) -> bf16 ; product -> bf16 ; + shift -> bf16, two elements per op
    o[w] = AFFINE ? n : add_bf16x2(mul_bf16x2(n, add_bf16x2(0x3F803F80u, aw[w])), bw[w]);
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm (no affine) + adaLN modulate, or LayerNorm with affine.      DIT:205-207, 63-64, 224-227
// NV = dim / 256 16-byte vectors per lane.
// ---------------------------------------------------------------------------------------------
template <int NV, bool AFFINE>
__global__ void __launch_bounds__(kRowWarps * 32, 3)   // <= 168 registers (the unrolled scale/shift loads would otherwise take 255): 12 warps/SM
ln_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ y, int64_t ldy, int rows,
          float eps, const __nv_bfloat16* __restrict__ shift0, const __nv_bfloat16* __restrict__ scale0,
          const __nv_bfloat16* __restrict__ shift1, const __nv_bfloat16* __restrict__ scale1, int rows_mod0) {
  const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  constexpr int D = NV * 256;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx);
  uint4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = ldg_nc_v4(xr + i * 32 + lane);
  float mean, rstd;
  ln_row_stats<NV>(v, eps, mean, rstd);
  // AFFINE: (shift0, scale0) carry (bias, weight). Otherwise pick the modulation row of this token.
  const bool first = AFFINE || row < rows_mod0;
  const uint4* sh = reinterpret_cast<const uint4*>(first ? shift0 : shift1);
  const uint4* sc = reinterpret_cast<const uint4*>(first ? scale0 : scale1);
  uint4* yr = reinterpret_cast<uint4*>(y + static_cast<int64_t>(row) * ldy);
  const float nmr = -mean * rstd;
#pragma unroll
  for (int i = 0; i < NV; ++i) yr[i * 32 + lane] = ln_apply<AFFINE>(v[i], __ldg(sc + i * 32 + lane), __ldg(sh + i * 32 + lane), rstd, nmr);
}

// RMSNorm arithmetic shared by every kernel that normalises a row (same bits from all of them): sum of squares and the
// scale step as packed fp32x2 operations (see ln_row_stats).
template <int NV>
__device__ __forceinline__ float rms_row_sumsq(const uint4 (&v)[NV]) {
  uint64_t q2 = pack2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t w[4] = {v[i].x, v[i].y, v[i].z, v[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint64_t x2 = pack2(bf16_lo(w[j]), bf16_hi(w[j]));
      q2 = fma2(x2, x2, q2);
    }
  }
  float a, b;
  unpack2(q2, a, b);
  return a + b;     // this lane's part; the caller sums over the warp
}
// norm -> bf16, * weight -> bf16 (DIT:99-110), two elements per operation
__device__ __forceinline__ void rms_scale_weight(const uint4& xv, const uint4& wv, float rs, uint32_t (&o)[4]) {
  const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, ww[4] = {wv.x, wv.y, wv.z, wv.w};
  const uint64_t r2 = pack2(rs, rs);
#pragma unroll
  for (int w = 0; w < 4; ++w) {
    float lo, hi;
    unpack2(mul2(pack2(bf16_lo(xw[w]), bf16_hi(xw[w])), r2), lo, hi);
    o[w] = mul_bf16x2(pack_bf16(lo, hi), ww[w]);
  }
}

// ---------------------------------------------------------------------------------------------
// RMSNorm over the full row (all heads) + optional 3-D RoPE, in place.      DIT:91-110, 140-144
// ---------------------------------------------------------------------------------------------
template <int NV, bool SCATTER, bool HMAX = false>
__global__ void __launch_bounds__(kRowWarps * 32)
rmsnorm_rope_kernel(__nv_bfloat16* __restrict__ x, int64_t ldx, int rows, float eps,
                    const __nv_bfloat16* __restrict__ weight, const float2* __restrict__ rope_tab, int gf, int gh,
                    int gw, int token_offset, const ScatterSpec sc, float* __restrict__ hmax2 = nullptr) {
  const int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  // hmax2 (optional): hmax2[h] = max over rows of ||out[row, h*128 : +128]||^2 — the query bound of fgb_attn_fwd_bounded_qk
  __shared__ float hred[HMAX ? kRowWarps : 1][HMAX ? 2 * NV : 1];
  if (HMAX) {
    if ((lane & 15) == 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) hred[threadIdx.x >> 5][2 * i + (lane >> 4)] = 0.f;
    }
    if (row >= rows) {     // the block reduction below needs every warp
      __syncthreads();
      return;
    }
  } else if (row >= rows) {
    return;
  }
  constexpr int D = NV * 256;
  uint4* xr = reinterpret_cast<uint4*>(x + static_cast<int64_t>(row) * ldx);
  uint4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = xr[i * 32 + lane];
  const float sq = rms_row_sumsq<NV>(v);
  const float rs = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);

  // Every vector of this lane starts at column (i*32+lane)*8, i.e. head-local complex lanes
  // c0..c0+3 with c0 = (lane % 16) * 4 — the same four rotation angles for all NV vectors.
  float cs[4], sn[4];
  bool rotate = false;
  if (rope_tab != nullptr) {
    const int t = token_offset + row;
    if (t < gf * gh * gw) {
      rotate = true;
      const int fi = t / (gh * gw);
      const int hi = (t / gw) % gh;
      const int wi = t % gw;
      const int c0 = (lane & 15) * 4;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j;
        const int pos = c < 22 ? fi : (c < 43 ? hi : wi);
        const float2 e = __ldg(rope_tab + pos * 64 + c);
        cs[j] = e.x;
        sn[j] = e.y;
      }
    }
  }
  const uint4* wr = reinterpret_cast<const uint4*>(weight);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    uint32_t o[4];
    rms_scale_weight(v[i], __ldg(wr + i * 32 + lane), rs, o);
    if (rotate) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float a = bf16_lo(o[j]), b = bf16_hi(o[j]);
        o[j] = pack_bf16(a * cs[j] - b * sn[j], a * sn[j] + b * cs[j]);
      }
    }
    const uint4 out = make_uint4(o[0], o[1], o[2], o[3]);
    if (HMAX) {
      float ss = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) ss += bf16_lo(o[j]) * bf16_lo(o[j]) + bf16_hi(o[j]) * bf16_hi(o[j]);
      const float hs = head_ss(ss);
      if ((lane & 15) == 0) hred[threadIdx.x >> 5][2 * i + (lane >> 4)] = hs;
    }
    if (SCATTER) {
      // fused Ulysses exchange: the normalised, rotated head slice goes straight to the peer that owns the head
      const int col = (i * 32 + lane) * 8;
      const int head = col >> 7;
      const int hpr = sc.heads / sc.world;
      const int peer = head / hpr;
      __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(sc.peers.p[peer]) +
                           ((static_cast<int64_t>(sc.rank) * sc.s_local + row) * sc.groups + sc.grp) * hpr * 128 +
                           (head - peer * hpr) * 128 + (col & 127);
      *reinterpret_cast<uint4*>(dst) = out;
    } else {
      xr[i * 32 + lane] = out;
    }
  }
  if (HMAX) {
    __syncthreads();
    if (threadIdx.x < 2 * NV) {
      float mx = 0.f;
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) mx = fmaxf(mx, hred[w][threadIdx.x]);
      if (mx > __ldcg(hmax2 + threadIdx.x))
        atomicMax(reinterpret_cast<unsigned int*>(hmax2) + threadIdx.x, __float_as_uint(mx));   // non-negative floats order like their bits
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Single-GPU self-attention prologue in ONE pass over the fused q|k|v rows: RMSNorm + weight + 3-D RoPE on q and on k
// (DIT:140-144) and the key bound kmax2[h] = max_rows ||k[row, h]||^2 of the bounded-score softmax — what used to be two
// rmsnorm_rope launches plus a head_norm_max pass that read k a second time. One warp per row, grid-stride; the per-head
// running maxima stay in registers (lane 0 and 16 of each vector index own a head).
// ---------------------------------------------------------------------------------------------
template <int NV, bool QMAX>
__global__ void __launch_bounds__(kRowWarps * 32)
qk_norm_rope_kernel(__nv_bfloat16* __restrict__ qkv, int64_t ld, int rows, float eps, const __nv_bfloat16* __restrict__ wq,
                    const __nv_bfloat16* __restrict__ wk, const float2* __restrict__ rope_tab, int gf, int gh, int gw, int token_offset,
                    float* __restrict__ kmax2, float* __restrict__ qmax2) {
  const int lane = threadIdx.x & 31;
  constexpr int D = NV * 256;
  float best[NV], bestq[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) best[i] = bestq[i] = 0.f;
  for (int row = blockIdx.x * kRowWarps + (threadIdx.x >> 5); row < rows; row += gridDim.x * kRowWarps) {
    float cs[4], sn[4];
    bool rotate = false;
    if (rope_tab != nullptr) {
      const int t = token_offset + row;
      if (t < gf * gh * gw) {
        rotate = true;
        const int fi = t / (gh * gw), hi = (t / gw) % gh, wi = t % gw;
        const int c0 = (lane & 15) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c0 + j;
          const int pos = c < 22 ? fi : (c < 43 ? hi : wi);
          const float2 e = __ldg(rope_tab + pos * 64 + c);
          cs[j] = e.x;
          sn[j] = e.y;
        }
      }
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      uint4* xr = reinterpret_cast<uint4*>(qkv + static_cast<int64_t>(row) * ld + g * D);
      uint4 v[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = xr[i * 32 + lane];
      const float sq = rms_row_sumsq<NV>(v);
      const float rs = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
      const uint4* wr = reinterpret_cast<const uint4*>(g == 0 ? wq : wk);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        uint32_t o[4];
        rms_scale_weight(v[i], __ldg(wr + i * 32 + lane), rs, o);
        if (rotate) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float a = bf16_lo(o[j]), b = bf16_hi(o[j]);
            o[j] = pack_bf16(a * cs[j] - b * sn[j], a * sn[j] + b * cs[j]);
          }
        }
        xr[i * 32 + lane] = make_uint4(o[0], o[1], o[2], o[3]);
        if (g == 1 || QMAX) {
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < 4; ++j) ss += bf16_lo(o[j]) * bf16_lo(o[j]) + bf16_hi(o[j]) * bf16_hi(o[j]);
          const float hs = head_ss(ss);
          if (g == 1) best[i] = fmaxf(best[i], hs);
          else bestq[i] = fmaxf(bestq[i], hs);
        }
      }
    }
  }
  __shared__ float red[QMAX ? 2 : 1][kRowWarps][2 * NV];
  if ((lane & 15) == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      red[0][threadIdx.x >> 5][2 * i + (lane >> 4)] = best[i];
      if (QMAX) red[1][threadIdx.x >> 5][2 * i + (lane >> 4)] = bestq[i];
    }
  }
  __syncthreads();
  if (threadIdx.x < (QMAX ? 4 : 2) * NV) {
    const int which = threadIdx.x / (2 * NV), h = threadIdx.x % (2 * NV);
    float* dst = which == 0 ? kmax2 : qmax2;
    if (dst != nullptr) {
      float mx = 0.f;
#pragma unroll
      for (int w = 0; w < kRowWarps; ++w) mx = fmaxf(mx, red[which][w][h]);
      // the maximum only grows: after the first CTAs almost nobody exceeds it any more, so look before the atomic (a stale,
      // lower value just costs one redundant atomic)
      if (mx > __ldcg(dst + h)) atomicMax(reinterpret_cast<unsigned int*>(dst) + h, __float_as_uint(mx));   // non-negative floats order like their bits
    }
  }
}

// ---------------------------------------------------------------------------------------------
// out[h] = max over rows of ||x[row, h*128 : (h+1)*128]||^2  — the key-norm bound of the bounded-score softmax
// (fgb_attn_fwd_bounded). One warp per row (grid-stride); a head is 16 consecutive lanes of one 16-byte vector index.
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(256)
head_norm_max_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, int rows, int heads, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  float best[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) best[i] = 0.f;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float f[8];
      float ss = 0.f;
      if (i * 32 + lane < heads * 16) {   // an odd head count (3 heads per rank at SP8) leaves the last half-warp idle
        unpack8(ldg_nc_v4(xr + i * 32 + lane), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) ss = fmaf(f[e], f[e], ss);
      }
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);   // stays inside the 16-lane half
      best[i] = fmaxf(best[i], ss);
    }
  }
  // block-level maximum first (8 warps), then one atomic per head and CTA: a few hundred contended atomics per head
  // instead of one per warp
  __shared__ float red[8][2 * NV];
  if ((lane & 15) == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) red[threadIdx.x >> 5][2 * i + (lane >> 4)] = best[i];
  }
  __syncthreads();
  if (threadIdx.x < 2 * NV && static_cast<int>(threadIdx.x) < heads) {
    float m = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) m = fmaxf(m, red[w][threadIdx.x]);
    atomicMax(reinterpret_cast<unsigned int*>(out) + threadIdx.x, __float_as_uint(m));   // non-negative floats order like their bits
  }
}

// ---------------------------------------------------------------------------------------------
// CFG combine + flow-match Euler update + first-frame restore.          PIPE:302, 307-309; FM:144-154
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float cfg_fm_one(float x, float np, float nn, bool has_neg, float cfg, float dsig) {
  float n = np;
  if (has_neg) n = round_bf16(nn + round_bf16(cfg * round_bf16(np - nn)));
  return x + round_bf16(n * dsig);  // caller rounds the sum
}

__global__ void cfg_fm_step_kernel(__nv_bfloat16* __restrict__ lat, const __nv_bfloat16* __restrict__ npos,
                                   const __nv_bfloat16* __restrict__ nneg, const __nv_bfloat16* __restrict__ first,
                                   float cfg, float dsig, int channels, int frames, int hw, int vec) {
  // one thread per `vec` (8 or 1) consecutive elements of the innermost (hw) axis
  const int64_t per_plane = hw / vec;
  const int64_t total = static_cast<int64_t>(channels) * frames * per_plane;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t plane = idx / per_plane;  // c * frames + f
    const int f = static_cast<int>(plane % frames);
    const int c = static_cast<int>(plane / frames);
    const int64_t off = idx * vec;
    if (vec == 8) {
      uint4 out;
      if (f == 0 && first != nullptr) {
        out = *reinterpret_cast<const uint4*>(first + static_cast<int64_t>(c) * hw + (idx % per_plane) * 8);
      } else {
        float x[8], a[8], b[8];
        unpack8(*reinterpret_cast<const uint4*>(lat + off), x);
        unpack8(ldg_nc_v4(npos + off), a);
        if (nneg) unpack8(ldg_nc_v4(nneg + off), b);
#pragma unroll
        for (int e = 0; e < 8; ++e) x[e] = cfg_fm_one(x[e], a[e], nneg ? b[e] : 0.f, nneg != nullptr, cfg, dsig);
        out = pack8(x);
      }
      *reinterpret_cast<uint4*>(lat + off) = out;
    } else {
      if (f == 0 && first != nullptr) {
        lat[off] = first[static_cast<int64_t>(c) * hw + (idx % per_plane)];
      } else {
        const float r = cfg_fm_one(__bfloat162float(lat[off]), __bfloat162float(npos[off]),
                                   nneg ? __bfloat162float(nneg[off]) : 0.f, nneg != nullptr, cfg, dsig);
        lat[off] = __float2bfloat16_rn(r);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// patchify (im2row for the k=s=(1,2,2) Conv3d) and unpatchify.                DIT:305, 338-351
// ---------------------------------------------------------------------------------------------
__global__ void patchify_rows_kernel(const __nv_bfloat16* __restrict__ lat, __nv_bfloat16* __restrict__ rows_out,
                                     int64_t ld_rows, int channels, int gf, int gh, int gw, int token_offset,
                                     int rows) {
  const int64_t total = static_cast<int64_t>(rows) * channels;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int c = static_cast<int>(idx % channels);
  const int r = static_cast<int>(idx / channels);
  const int t = token_offset + r;
  uint2 out = make_uint2(0, 0);
  if (t < gf * gh * gw) {
    const int f = t / (gh * gw), h = (t / gw) % gh, w = t % gw;
    const int64_t W2 = 2 * gw, H2 = 2 * gh;
    const __nv_bfloat16* src = lat + ((static_cast<int64_t>(c) * gf + f) * H2 + 2 * h) * W2 + 2 * w;
    out.x = *reinterpret_cast<const uint32_t*>(src);       // y = 0, z = 0..1
    out.y = *reinterpret_cast<const uint32_t*>(src + W2);  // y = 1, z = 0..1
  }
  *reinterpret_cast<uint2*>(rows_out + static_cast<int64_t>(r) * ld_rows + c * 4) = out;
}

__global__ void unpatchify_kernel(const __nv_bfloat16* __restrict__ rows_in, int64_t ld_rows,
                                  __nv_bfloat16* __restrict__ out, int channels, int gf, int gh, int gw) {
  // one thread per (c, f, row of the latent, patch column) -> two adjacent output elements
  const int64_t H2 = 2 * gh;
  const int64_t total = static_cast<int64_t>(channels) * gf * H2 * gw;
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= total) return;
  const int w = static_cast<int>(idx % gw);
  const int hy = static_cast<int>((idx / gw) % H2);
  const int f = static_cast<int>((idx / (gw * H2)) % gf);
  const int c = static_cast<int>(idx / (gw * H2 * gf));
  const int h = hy >> 1, y = hy & 1;
  const int64_t t = (static_cast<int64_t>(f) * gh + h) * gw + w;
  const __nv_bfloat16* src = rows_in + t * ld_rows + y * 2 * channels + c;
  __nv_bfloat162 v;
  v.x = src[0];         // z = 0
  v.y = src[channels];  // z = 1
  *reinterpret_cast<__nv_bfloat162*>(out + ((static_cast<int64_t>(c) * gf + f) * H2 + hy) * (2 * gw) + 2 * w) = v;
}

// ---------------------------------------------------------------------------------------------
// small embedding helpers                                                   DIT:67-71, 312-318
// ---------------------------------------------------------------------------------------------
__global__ void sinusoid_kernel(const float* __restrict__ t, __nv_bfloat16* __restrict__ out, int rows, int dim) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= rows * dim) return;
  const int r = idx / dim, j = idx % dim, half = dim / 2;
  const int i = j < half ? j : j - half;
  const double w = pow(10000.0, -static_cast<double>(i) / static_cast<double>(half));
  const double a = static_cast<double>(t[r]) * w;
  out[idx] = __float2bfloat16_rn(static_cast<float>(j < half ? cos(a) : sin(a)));
}

__global__ void silu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= n) return;
  const float v = __bfloat162float(x[idx]);
  y[idx] = __float2bfloat16_rn(v / (1.0f + expf(-v)));
}

__global__ void add_bcast_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                 __nv_bfloat16* __restrict__ out, int64_t n, int64_t cols, int64_t period) {
  const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (idx >= n) return;
  out[idx] = __float2bfloat16_rn(__bfloat162float(a[idx]) + __bfloat162float(b[(idx % cols) % period]));
}

// ---------------------------------------------------------------------------------------------
// Ulysses head re-partition pack / unpack                       xdit_context_parallel.py:125-146
// ---------------------------------------------------------------------------------------------
template <bool PACK>
__global__ void sp_heads_kernel(__nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ buf,
                                int s_local, int heads, int groups, int world) {
  // x is [s_local][groups][heads][128] (e.g. groups = 3 for a fused q|k|v row);
  // buf is [world][s_local][groups][heads/world][128]; 16 vectors of 16 B per head row.
  const int hpr = heads / world;
  const int gh = groups * heads;
  const int64_t total = static_cast<int64_t>(s_local) * gh * 16;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int vec = static_cast<int>(idx & 15);
    const int g_head = static_cast<int>((idx >> 4) % gh);
    const int64_t s = (idx >> 4) / gh;
    const int grp = g_head / heads, head = g_head % heads;
    const int peer = head / hpr, hh = head % hpr;
    uint4* xp = reinterpret_cast<uint4*>(x + s * ldx + static_cast<int64_t>(g_head) * 128) + vec;
    uint4* bp = reinterpret_cast<uint4*>(
                    buf + (((static_cast<int64_t>(peer) * s_local + s) * groups + grp) * hpr + hh) * 128) + vec;
    if (PACK) *bp = *xp; else *xp = *bp;
  }
}

// ---------------------------------------------------------------------------------------------
// Ulysses exchange over NVLink peer memory (no NCCL on the data path)
// ---------------------------------------------------------------------------------------------
// x [s_local][groups][heads][128] of THIS rank -> recv buffer of every peer, laid out [world*s_local tokens][groups]
// [heads/world][128]: the head slice owned by peer `q` goes to rows [rank*s_local, (rank+1)*s_local) of q's buffer.
// 16-byte stores; consecutive threads write consecutive 16 B of one 256-byte head row (full NVLink packets).
__global__ void sp_scatter_heads_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, PeerPtrs peers, int s_local,
                                        int heads, int groups, int world, int rank, int grp0, int groups_total) {
  const int hpr = heads / world;
  const int gh = groups * heads;
  const int64_t total = static_cast<int64_t>(s_local) * gh * 16;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int vec = static_cast<int>(idx & 15);
    const int g_head = static_cast<int>((idx >> 4) % gh);
    const int64_t s = (idx >> 4) / gh;
    const int grp = g_head / heads, head = g_head % heads;
    const int peer = head / hpr, hh = head % hpr;
    const uint4 v = ldg_nc_v4(reinterpret_cast<const uint4*>(x + s * ldx + static_cast<int64_t>(g_head) * 128) + vec);
    __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(peers.p[peer]) +
                         ((static_cast<int64_t>(rank) * s_local + s) * groups_total + grp0 + grp) * hpr * 128 + hh * 128;
    reinterpret_cast<uint4*>(dst)[vec] = v;
  }
}

// Reverse direction (backward of the Ulysses exchange): x [s_pad][groups][heads/world][128] — this rank's heads, ALL tokens
// (e.g. dq|dk|dv from fgb_attn_bwd) -> the token-major matrix [rows][groups][heads][128] of the rank that owns each token.
__global__ void sp_return_heads_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, PeerPtrs peers, int64_t ld_dst, int rows,
                                       int s_pad, int heads, int groups, int world, int rank) {
  const int hpr = heads / world;
  const int gh = groups * hpr;
  const int64_t total = static_cast<int64_t>(s_pad) * gh * 16;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int vec = static_cast<int>(idx & 15);
    const int g_head = static_cast<int>((idx >> 4) % gh);
    const int64_t s = (idx >> 4) / gh;
    const int grp = g_head / hpr, hh = g_head % hpr;
    const int peer = static_cast<int>(s / rows);
    const uint4 v = ldg_nc_v4(reinterpret_cast<const uint4*>(x + s * ldx + static_cast<int64_t>(g_head) * 128) + vec);
    __nv_bfloat16* dst = static_cast<__nv_bfloat16*>(peers.p[peer]) + (s - static_cast<int64_t>(peer) * rows) * ld_dst +
                         static_cast<int64_t>(grp) * heads * 128 + (rank * hpr + hh) * 128;
    reinterpret_cast<uint4*>(dst)[vec] = v;
  }
}

// Cross-GPU barrier for the exchange: thread q publishes `epoch` in peer q's flag slot [rank] (release, system scope)
// and then waits until peer q has published `epoch` in ours (acquire). All earlier peer stores of this stream are
// complete (kernel boundary) and made visible by the fence. Bounded: after `limit` clocks (default ~30 s; ranks can be
// seconds apart at start-up) the waiting thread writes the epoch to *status and gives up, so the host can name the failure
// (fgb_sp_barrier_status); without a status word the kernel traps instead of hanging the box.
constexpr long long kBarrierClocks = 50000000000ll;

__device__ __forceinline__ void sp_epoch_exchange(const PeerPtrs& flags, int q, int rank, int epoch, int* status, long long limit) {
  __threadfence_system();
  int* remote = static_cast<int*>(flags.p[q]) + rank;
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(remote), "r"(epoch) : "memory");
  const int* mine = static_cast<const int*>(flags.p[rank]) + q;
  const long long t0 = clock64();
  for (unsigned spins = 1;; ++spins) {
    int v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(mine) : "memory");
    if (v - epoch >= 0) break;
    // slow path only: an exchange that has already lost a peer does not wait another `limit` at each of its later barriers
    if ((spins & 1023u) == 0 && status != nullptr && *reinterpret_cast<volatile int*>(status) != 0) break;
    if (clock64() - t0 > limit) {   // a peer died or never launched: report instead of hanging
      if (status == nullptr) __trap();
      atomicCAS(status, 0, epoch);  // the FIRST epoch that timed out stays in the status word
      break;
    }
  }
  __threadfence_system();
}

__global__ void sp_barrier_kernel(PeerPtrs flags, int world, int rank, int epoch, int* __restrict__ status, long long limit) {
  const int q = threadIdx.x;
  if (q >= world) return;
  sp_epoch_exchange(flags, q, rank, epoch, status, limit);
}

// ---------------------------------------------------------------------------------------------
// Fused Ulysses exchange, receiver side (the sender is the q|k|v GEMM epilogue, fgb_gemm_qkv_scatter)
// ---------------------------------------------------------------------------------------------
// Barrier 0 of a block with the row statistics riding along: every rank pushes the sums of squares of its q and k rows
// (rowsq [2][rows], filled by the GEMM epilogue's atomics) into every peer's stats matrix [2][s_pad] at its own row window,
// zeroes rowsq and kmax2 for the next use, then runs the epoch exchange of sp_barrier_kernel. One block.
__global__ void __launch_bounds__(1024)
sp_stats_barrier_kernel(PeerPtrs flags, PeerPtrs stats, float* __restrict__ rowsq, int rows, int s_pad, float* __restrict__ kmax2, int hpr,
                        int world, int rank, int epoch, int* __restrict__ status) {
  for (int idx = threadIdx.x; idx < 2 * rows; idx += blockDim.x) {
    const int g = idx / rows, r = idx - g * rows;
    const float v = rowsq[idx];
    rowsq[idx] = 0.f;
    for (int q = 0; q < world; ++q) static_cast<float*>(stats.p[q])[static_cast<int64_t>(g) * s_pad + rank * rows + r] = v;
  }
  if (static_cast<int>(threadIdx.x) < hpr) kmax2[threadIdx.x] = 0.f;
  __threadfence_system();
  __syncthreads();
  const int q = threadIdx.x;
  if (q >= world) return;
  sp_epoch_exchange(flags, q, rank, epoch, status, kBarrierClocks);
}

// RMSNorm (statistics of the FULL row, received with the data) + weight + 3-D RoPE on the received q and k groups, in place:
// recv [s_pad][3][hpr][128]; the full-row mean square of token t is stats[g*s_pad + t] / dim. Also leaves
// kmax2[h] = max over the real tokens of ||k[t, h]||^2 (the key bound of the bounded-score softmax). One warp per token row,
// grid-stride (the running maxima stay in registers).
__global__ void __launch_bounds__(256)
recv_norm_rope_kernel(__nv_bfloat16* __restrict__ recv, int s_pad, int tokens, int hpr, const float* __restrict__ stats, int dim,
                      float eps, const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk,
                      const float2* __restrict__ rope_tab, int gf, int gh, int gw, float* __restrict__ kmax2, float* __restrict__ qmax2) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  const int vecs = hpr * 16;                  // 16-byte vectors per group and row
  const int64_t ld = static_cast<int64_t>(3) * hpr * 128;
  const float inv_d = 1.0f / dim;
  float best[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // per pass (head pair) running max of ||k||^2; hpr <= 12
  float bestq[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};  // the same for q (qmax2 != NULL)
  const int c0 = (lane & 15) * 4;
  for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < s_pad; t += warps) {
    float cs[4], sn[4];
    const bool rotate = rope_tab != nullptr && t < gf * gh * gw;
    if (rotate) {
      const int fi = t / (gh * gw), hi = (t / gw) % gh, wi = t % gw;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j;
        const int pos = c < 22 ? fi : (c < 43 ? hi : wi);
        const float2 e = __ldg(rope_tab + pos * 64 + c);
        cs[j] = e.x;
        sn[j] = e.y;
      }
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const float rs = rsqrtf(__ldg(stats + static_cast<int64_t>(g) * s_pad + t) * inv_d + eps);
      uint4* row = reinterpret_cast<uint4*>(recv + static_cast<int64_t>(t) * ld + g * hpr * 128);
      const uint4* wr = reinterpret_cast<const uint4*>(g == 0 ? wq : wk);
#pragma unroll
      for (int pass = 0; pass < 6; ++pass) {
        const int c = pass * 32 + lane;
        float ss = 0.f;
        if (c < vecs) {
          uint32_t o[4];
          rms_scale_weight(row[c], __ldg(wr + c), rs, o);
          if (rotate) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float a = bf16_lo(o[j]), b = bf16_hi(o[j]);
              o[j] = pack_bf16(a * cs[j] - b * sn[j], a * sn[j] + b * cs[j]);
            }
          }
          row[c] = make_uint4(o[0], o[1], o[2], o[3]);
          if (g == 1 || qmax2 != nullptr) {
#pragma unroll
            for (int j = 0; j < 4; ++j) ss += bf16_lo(o[j]) * bf16_lo(o[j]) + bf16_hi(o[j]) * bf16_hi(o[j]);
          }
        }
        if ((g == 1 || qmax2 != nullptr) && pass * 32 < vecs) {   // warp-uniform
          const float hs = head_ss(ss);
          if (g == 1) {
            if (t < tokens) best[pass] = fmaxf(best[pass], hs);
          } else {
            bestq[pass] = fmaxf(bestq[pass], hs);     // padded query rows are computed (and discarded) too: keep them inside the bound
          }
        }
      }
    }
  }
  __shared__ float red[2][8][12];
  if ((lane & 15) == 0) {
#pragma unroll
    for (int pass = 0; pass < 6; ++pass) {
      red[0][threadIdx.x >> 5][2 * pass + (lane >> 4)] = best[pass];
      red[1][threadIdx.x >> 5][2 * pass + (lane >> 4)] = bestq[pass];
    }
  }
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < 2 * 12) {
    const int which = threadIdx.x / 12, h = threadIdx.x % 12;
    float* dst = which == 0 ? kmax2 : qmax2;
    if (h < hpr && dst != nullptr) {
      float mx = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) mx = fmaxf(mx, red[which][w][h]);
      atomicMax(reinterpret_cast<unsigned int*>(dst) + h, __float_as_uint(mx));   // non-negative floats order like their bits
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Row-streaming variants of the three row kernels above (LayerNorm+modulate / affine, RMSNorm(+RoPE), fused q|k RMSNorm+RoPE
// with the key bound) for the large launches of the denoise step.
//
// The one-row-per-warp kernels keep one row per warp in flight and only while that warp is in its load phase (64-79 % of
// the HBM copy rate, BENCH_r01). Here one persistent CTA per SM decouples the two: a single producer thread keeps a ring of
// kSlots rows in flight with 1-D bulk copies (cp.async.bulk, completion on mbarriers) — 132 KB per SM at D = 3072,
// independent of what the consumers are doing — and 11 consumer warps take rows out of shared memory (16 bytes per lane,
// conflict-free), release the slot as soon as the row is in registers, and do the same arithmetic and the same 16-byte
// stores as before. Items are dealt round-robin over the CTAs: item n of CTA b is global item b + n*gridDim.x.
// ---------------------------------------------------------------------------------------------
constexpr int kStreamWarps = 11;      // + 1 producer warp = 12 warps: registers are allocated per 4 warps, a 13th warp would cap a thread at 128
constexpr int kStreamThreads = (kStreamWarps + 1) * 32;

template <int NV>
struct StreamCfg {
  static constexpr int kRowBytes = NV * 512;
  // A slot is always consumed by the SAME warp (kSlots is a multiple of kStreamWarps and a warp takes every kStreamWarps-th item),
  // so a warp looks at a slot's `full` barrier for fill n only after it consumed fill n-1 itself. With a free slot -> warp
  // mapping a fast warp could test the parity of fill n while fill n-1 had not even landed (bulk copies complete out of
  // order when one of them misses the TLB), see the opposite parity as "done", read a stale row and arrive on `empty` once
  // too often: wrong rows and, a few launches in a hundred at > 128 MB, a dead ring (caught by the spin-limit trap).
  static constexpr int kSlots = (2 * kStreamWarps * kRowBytes <= 160 * 1024) ? 2 * kStreamWarps : kStreamWarps;
  static constexpr int kSmem = kSlots * kRowBytes + 2 * kSlots * 8 + 128;
};

// OP: struct with `void row(int item, int lane, uint4 (&v)[NV])` (consumes one row / row segment held in registers) and
// `void finish(int warp, int lane)` (after the CTA's last row; may use the barrier below), plus
// `const __nv_bfloat16* src(int item) const` (where the item's NV*512 bytes start).
template <int NV, class OP>
__global__ void __launch_bounds__(kStreamThreads, 1) row_stream_kernel(int items, OP op) {
  using C = StreamCfg<NV>;
  extern __shared__ uint8_t stream_smem_raw[];
  uint8_t* ring = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(stream_smem_raw) + 127) & ~uintptr_t(127));
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + C::kSlots * C::kRowBytes);
  uint64_t* empty = full + C::kSlots;
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < C::kSlots; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  if (warp == kStreamWarps) {
    if (elect_one()) {
      int k = 0;
      for (int item = blockIdx.x; item < items; item += gridDim.x, ++k) {
        const int slot = k % C::kSlots;
        mbar_wait(&empty[slot], ((k / C::kSlots) & 1) ^ 1);
        mbar_expect_tx(&full[slot], C::kRowBytes);
        bulk_load(ring + slot * C::kRowBytes, op.src(item), C::kRowBytes, &full[slot]);
      }
    }
  } else {
    for (int k = warp;; k += kStreamWarps) {
      const int item = blockIdx.x + k * gridDim.x;
      if (item >= items) break;
      const int slot = k % C::kSlots;
      mbar_wait(&full[slot], (k / C::kSlots) & 1);
      const uint4* src = reinterpret_cast<const uint4*>(ring + slot * C::kRowBytes);
      uint4 v[NV];
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = src[i * 32 + lane];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[slot]);     // the row is in registers: the slot can be refilled
      op.row(item, lane, v);
    }
  }
  __syncwarp();
  op.finish(warp, lane);
}

template <int NV, bool AFFINE>
struct LnStreamOp {
  const __nv_bfloat16* x;
  int64_t ldx;
  __nv_bfloat16* y;
  int64_t ldy;
  float eps;
  const __nv_bfloat16 *shift0, *scale0, *shift1, *scale1;
  int rows_mod0;
  __device__ __forceinline__ const __nv_bfloat16* src(int item) const { return x + static_cast<int64_t>(item) * ldx; }
  __device__ __forceinline__ void row(int row, int lane, uint4 (&v)[NV]) {
    float mean, rstd;
    ln_row_stats<NV>(v, eps, mean, rstd);
    const bool first = AFFINE || row < rows_mod0;
    const uint4* sh = reinterpret_cast<const uint4*>(first ? shift0 : shift1);
    const uint4* sc = reinterpret_cast<const uint4*>(first ? scale0 : scale1);
    uint4* yr = reinterpret_cast<uint4*>(y + static_cast<int64_t>(row) * ldy);
    const float nmr = -mean * rstd;
#pragma unroll
    for (int i = 0; i < NV; ++i) yr[i * 32 + lane] = ln_apply<AFFINE>(v[i], __ldg(sc + i * 32 + lane), __ldg(sh + i * 32 + lane), rstd, nmr);
  }
  __device__ __forceinline__ void finish(int, int) {}
};

// SEGS = 1: RMSNorm(+RoPE) of [rows, D] in place (weight w0). SEGS = 2: the q and k groups of the fused q|k|v rows (item =
// row*2 + group; weights w0 / w1) plus the key bound kmax2[h] = max ||k[row, h]||^2.
template <int NV, int SEGS>
struct RmsStreamOp {
  __nv_bfloat16* x;
  int64_t ldx;
  float eps;
  const __nv_bfloat16 *w0, *w1;
  const float2* rope_tab;
  int gf, gh, gw, token_offset;
  float *kmax2, *qmax2;
  float best[SEGS == 2 ? NV : 1], bestq[SEGS == 2 ? NV : 1];
  __device__ __forceinline__ __nv_bfloat16* dst(int item) const {
    return SEGS == 2 ? x + static_cast<int64_t>(item >> 1) * ldx + (item & 1) * (NV * 256) : x + static_cast<int64_t>(item) * ldx;
  }
  __device__ __forceinline__ const __nv_bfloat16* src(int item) const { return dst(item); }
  __device__ __forceinline__ void row(int item, int lane, uint4 (&v)[NV]) {
    constexpr int D = NV * 256;
    const int r = SEGS == 2 ? item >> 1 : item;
    const int g = SEGS == 2 ? item & 1 : 0;
    const float sq = rms_row_sumsq<NV>(v);
    const float rs = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
    float cs[4], sn[4];
    bool rotate = false;
    if (rope_tab != nullptr) {
      const int t = token_offset + r;
      if (t < gf * gh * gw) {
        rotate = true;
        const int fi = t / (gh * gw), hi = (t / gw) % gh, wi = t % gw;
        const int c0 = (lane & 15) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int c = c0 + j;
          const int pos = c < 22 ? fi : (c < 43 ? hi : wi);
          const float2 e = __ldg(rope_tab + pos * 64 + c);
          cs[j] = e.x;
          sn[j] = e.y;
        }
      }
    }
    const uint4* wr = reinterpret_cast<const uint4*>(g == 0 ? w0 : w1);
    uint4* xr = reinterpret_cast<uint4*>(dst(item));
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      uint32_t o[4];
      rms_scale_weight(v[i], __ldg(wr + i * 32 + lane), rs, o);
      if (rotate) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float a = bf16_lo(o[j]), b = bf16_hi(o[j]);
          o[j] = pack_bf16(a * cs[j] - b * sn[j], a * sn[j] + b * cs[j]);
        }
      }
      xr[i * 32 + lane] = make_uint4(o[0], o[1], o[2], o[3]);
      if (SEGS == 2) {
        float ss = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) ss += bf16_lo(o[j]) * bf16_lo(o[j]) + bf16_hi(o[j]) * bf16_hi(o[j]);
        const float hs = head_ss(ss);
        if (g == 1) best[i] = fmaxf(best[i], hs);
        else bestq[i] = fmaxf(bestq[i], hs);
      }
    }
  }
  __device__ __forceinline__ void finish(int warp, int lane) {
    if (SEGS != 2) return;
    __shared__ float red[2][kStreamWarps][2 * NV];
    if (warp < kStreamWarps && (lane & 15) == 0) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        red[0][warp][2 * i + (lane >> 4)] = best[i];
        red[1][warp][2 * i + (lane >> 4)] = bestq[i];
      }
    }
    __syncthreads();
    if (threadIdx.x < 4 * NV) {
      const int which = threadIdx.x / (2 * NV), h = threadIdx.x % (2 * NV);
      float* dst = which == 0 ? kmax2 : qmax2;
      if (dst != nullptr) {
        float mx = 0.f;
#pragma unroll
        for (int w = 0; w < kStreamWarps; ++w) mx = fmaxf(mx, red[which][w][h]);
        if (mx > __ldcg(dst + h)) atomicMax(reinterpret_cast<unsigned int*>(dst) + h, __float_as_uint(mx));   // non-negative floats order like their bits
      }
    }
  }
};

// rows below this stay on the one-row-per-warp kernels: a persistent grid needs a few rows per warp.
// Which kernels stream (FGB_EW_STREAM: bit 0 = LayerNorm kernels, bit 1 = RMSNorm kernels; default 1): measured inside the
// power-capped denoise step (profiles/r02_ew_stream_ab.log), the two LayerNorm kernels gain (ln_modulate 8.35 -> 7.83 ms,
// ln_affine 4.92 -> 4.04 ms per step) while the in-place RMSNorm kernels lose (10.2 -> 12.3 ms, 3.9 -> 4.7 ms); alone at full
// clocks the one-row-per-warp kernels already run at 0.91-0.93 of the copy rate and the ring does not help either.
constexpr int kStreamMinRows = 2048;
static bool stream_enabled(int bit) {
  static int mask = -1;
  if (mask < 0) {
    const char* e = getenv("FGB_EW_STREAM");
    mask = e ? atoi(e) : 1;
  }
  return (mask >> bit) & 1;
}

template <int NV, class OP>
static int launch_row_stream(const fgb_ctx* ctx, cudaStream_t s, int items, const OP& op) {
  auto kfn = row_stream_kernel<NV, OP>;
  static unsigned long long configured = 0;
  if (first_use_on_device(configured))
    FGB_CUDA(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, StreamCfg<NV>::kSmem));
  const int grid = items < ctx->sm_count ? items : ctx->sm_count;
  kfn<<<grid, kStreamThreads, StreamCfg<NV>::kSmem, s>>>(items, op);
  FGB_LAUNCH_CHECK("row_stream_kernel");
  return FGB_OK;
}

static inline int grid_1d(int64_t n, int block) { return static_cast<int>((n + block - 1) / block); }

template <bool AFFINE>
static int launch_ln(const fgb_ctx* ctx, int nv, dim3 grid, cudaStream_t s, const __nv_bfloat16* x, int64_t ldx, __nv_bfloat16* y,
                     int64_t ldy, int rows, float eps, const __nv_bfloat16* sh0, const __nv_bfloat16* sc0,
                     const __nv_bfloat16* sh1, const __nv_bfloat16* sc1, int rows_mod0) {
  if (rows >= kStreamMinRows && stream_enabled(0)) {
#define FGB_LNS_CASE(NV)                                                                                     \
  case NV:                                                                                                   \
    return launch_row_stream<NV>(ctx, s, rows, LnStreamOp<NV, AFFINE>{x, ldx, y, ldy, eps, sh0, sc0, sh1, sc1, rows_mod0});
    switch (nv) {
      FGB_LNS_CASE(1) FGB_LNS_CASE(2) FGB_LNS_CASE(3) FGB_LNS_CASE(4) FGB_LNS_CASE(6) FGB_LNS_CASE(8) FGB_LNS_CASE(12) FGB_LNS_CASE(16)
      FGB_LNS_CASE(20)
      default: break;
    }
#undef FGB_LNS_CASE
  }
#define FGB_LN_CASE(NV)                                                                                    \
  case NV:                                                                                                 \
    ln_kernel<NV, AFFINE><<<grid, kRowWarps * 32, 0, s>>>(x, ldx, y, ldy, rows, eps, sh0, sc0, sh1, sc1, rows_mod0); \
    break;
  switch (nv) {
    FGB_LN_CASE(1) FGB_LN_CASE(2) FGB_LN_CASE(3) FGB_LN_CASE(4) FGB_LN_CASE(6) FGB_LN_CASE(8) FGB_LN_CASE(12) FGB_LN_CASE(16)
    FGB_LN_CASE(20)
    default:
      return set_error(FGB_ERR_UNSUPPORTED, "layer norm: dim %d is not one of 256*{1,2,3,4,6,8,12,16,20}", nv * 256);
  }
#undef FGB_LN_CASE
  FGB_LAUNCH_CHECK("ln_kernel");
  return FGB_OK;
}

}  // namespace fgb

using namespace fgb;
typedef __nv_bfloat16 bf16;

extern "C" int fgb_ln_modulate(fgb_ctx* ctx, const void* x, int64_t ldx, void* y, int64_t ldy, int32_t rows,
                               int32_t dim, float eps, const void* shift0, const void* scale0, const void* shift1,
                               const void* scale1, int32_t rows_mod0, void* stream) {
  FGB_CHECK_ARG(ctx && x && y && shift0 && scale0 && shift1 && scale1, "fgb_ln_modulate: NULL argument");
  FGB_CHECK_ARG(rows > 0 && dim > 0 && dim % 256 == 0, "fgb_ln_modulate: rows=%d dim=%d (dim must be a multiple of 256)", rows, dim);
  FGB_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && aligned16(x) && aligned16(y) && aligned16(shift0) && aligned16(scale0) &&
                    aligned16(shift1) && aligned16(scale1),
                "fgb_ln_modulate: operands must be 16-byte aligned");
  dim3 grid((rows + kRowWarps - 1) / kRowWarps);
  return launch_ln<false>(ctx, dim / 256, grid, static_cast<cudaStream_t>(stream), static_cast<const bf16*>(x), ldx,
                          static_cast<bf16*>(y), ldy, rows, eps, static_cast<const bf16*>(shift0),
                          static_cast<const bf16*>(scale0), static_cast<const bf16*>(shift1),
                          static_cast<const bf16*>(scale1), rows_mod0);
}

extern "C" int fgb_ln_affine(fgb_ctx* ctx, const void* x, int64_t ldx, void* y, int64_t ldy, int32_t rows, int32_t dim,
                             float eps, const void* weight, const void* bias, void* stream) {
  FGB_CHECK_ARG(ctx && x && y && weight && bias, "fgb_ln_affine: NULL argument");
  FGB_CHECK_ARG(rows > 0 && dim > 0 && dim % 256 == 0, "fgb_ln_affine: rows=%d dim=%d (dim must be a multiple of 256)", rows, dim);
  FGB_CHECK_ARG(ldx % 8 == 0 && ldy % 8 == 0 && aligned16(x) && aligned16(y) && aligned16(weight) && aligned16(bias),
                "fgb_ln_affine: operands must be 16-byte aligned");
  dim3 grid((rows + kRowWarps - 1) / kRowWarps);
  return launch_ln<true>(ctx, dim / 256, grid, static_cast<cudaStream_t>(stream), static_cast<const bf16*>(x), ldx,
                         static_cast<bf16*>(y), ldy, rows, eps, static_cast<const bf16*>(bias),
                         static_cast<const bf16*>(weight), static_cast<const bf16*>(bias),
                         static_cast<const bf16*>(weight), rows);
}

extern "C" int fgb_rmsnorm_rope(fgb_ctx* ctx, void* x, int64_t ldx, int32_t rows, int32_t dim, float eps,
                                const void* weight, const void* rope_tab, int32_t gf, int32_t gh, int32_t gw,
                                int32_t token_offset, void* stream) {
  FGB_CHECK_ARG(ctx && x && weight, "fgb_rmsnorm_rope: NULL argument");
  FGB_CHECK_ARG(rows > 0 && dim > 0 && dim % 256 == 0, "fgb_rmsnorm_rope: rows=%d dim=%d (dim must be a multiple of 256)", rows, dim);
  FGB_CHECK_ARG(ldx % 8 == 0 && aligned16(x) && aligned16(weight), "fgb_rmsnorm_rope: operands must be 16-byte aligned");
  if (rope_tab) {
    FGB_CHECK_ARG(gf > 0 && gh > 0 && gw > 0 && gf <= 1024 && gh <= 1024 && gw <= 1024,
                  "fgb_rmsnorm_rope: grid (%d,%d,%d) outside the 1024-position RoPE table", gf, gh, gw);
    FGB_CHECK_ARG(token_offset >= 0, "fgb_rmsnorm_rope: negative token offset");
  }
  dim3 grid((rows + kRowWarps - 1) / kRowWarps);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bf16* xp = static_cast<bf16*>(x);
  const bf16* wp = static_cast<const bf16*>(weight);
  const float2* tab = static_cast<const float2*>(rope_tab);
  if (rows >= kStreamMinRows && stream_enabled(1)) {
#define FGB_RMSS_CASE(NV)                                                                                       \
  case NV: {                                                                                                    \
    RmsStreamOp<NV, 1> op{xp, ldx, eps, wp, wp, tab, gf, gh, gw, token_offset, nullptr, nullptr, {0.f}, {0.f}};                  \
    return launch_row_stream<NV>(ctx, s, rows, op);                                                             \
  }
    switch (dim / 256) {
      FGB_RMSS_CASE(1) FGB_RMSS_CASE(2) FGB_RMSS_CASE(3) FGB_RMSS_CASE(4) FGB_RMSS_CASE(6) FGB_RMSS_CASE(8) FGB_RMSS_CASE(12)
      FGB_RMSS_CASE(16) FGB_RMSS_CASE(20)
      default: break;
    }
#undef FGB_RMSS_CASE
  }
  ScatterSpec none{};
#define FGB_RMS_CASE(NV)                                                                                         \
  case NV:                                                                                                       \
    rmsnorm_rope_kernel<NV, false><<<grid, kRowWarps * 32, 0, s>>>(xp, ldx, rows, eps, wp, tab, gf, gh, gw, token_offset, none); \
    break;
  switch (dim / 256) {
    FGB_RMS_CASE(1) FGB_RMS_CASE(2) FGB_RMS_CASE(3) FGB_RMS_CASE(4) FGB_RMS_CASE(6) FGB_RMS_CASE(8) FGB_RMS_CASE(12) FGB_RMS_CASE(16)
    FGB_RMS_CASE(20)
    default:
      return set_error(FGB_ERR_UNSUPPORTED, "rmsnorm: dim %d is not one of 256*{1,2,3,4,6,8,12,16,20}", dim);
  }
#undef FGB_RMS_CASE
  FGB_LAUNCH_CHECK("rmsnorm_rope_kernel");
  return FGB_OK;
}

extern "C" int fgb_rmsnorm_hmax(fgb_ctx* ctx, void* x, int64_t ldx, int32_t rows, int32_t dim, float eps, const void* weight, void* hmax2,
                               void* stream) {
  FGB_CHECK_ARG(ctx && x && weight && hmax2, "fgb_rmsnorm_hmax: NULL argument");
  FGB_CHECK_ARG(rows > 0 && dim > 0 && dim % 256 == 0, "fgb_rmsnorm_hmax: rows=%d dim=%d (dim must be a multiple of 256)", rows, dim);
  FGB_CHECK_ARG(ldx % 8 == 0 && aligned16(x) && aligned16(weight), "fgb_rmsnorm_hmax: operands must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FGB_CUDA(cudaMemsetAsync(hmax2, 0, sizeof(float) * (dim / 128), s));
  dim3 grid((rows + kRowWarps - 1) / kRowWarps);
  ScatterSpec none{};
#define FGB_RMSH_CASE(NV)                                                                                        \
  case NV:                                                                                                       \
    rmsnorm_rope_kernel<NV, false, true><<<grid, kRowWarps * 32, 0, s>>>(static_cast<bf16*>(x), ldx, rows, eps, static_cast<const bf16*>(weight), \
                                                                   nullptr, 1, 1, 1, 0, none, static_cast<float*>(hmax2)); \
    break;
  switch (dim / 256) {
    FGB_RMSH_CASE(1) FGB_RMSH_CASE(2) FGB_RMSH_CASE(3) FGB_RMSH_CASE(4) FGB_RMSH_CASE(6) FGB_RMSH_CASE(8) FGB_RMSH_CASE(12) FGB_RMSH_CASE(16)
    FGB_RMSH_CASE(20)
    default:
      return set_error(FGB_ERR_UNSUPPORTED, "rmsnorm: dim %d is not one of 256*{1,2,3,4,6,8,12,16,20}", dim);
  }
#undef FGB_RMSH_CASE
  FGB_LAUNCH_CHECK("rmsnorm_rope_kernel");
  return FGB_OK;
}

extern "C" int fgb_cfg_fm_step(fgb_ctx* ctx, void* latents, const void* noise_pos, const void* noise_neg,
                               const void* first_frame, float cfg_scale, float sigma_delta, int32_t channels,
                               int32_t frames, int32_t hw, void* stream) {
  FGB_CHECK_ARG(ctx && latents && noise_pos, "fgb_cfg_fm_step: NULL argument");
  FGB_CHECK_ARG(channels > 0 && frames > 0 && hw > 0, "fgb_cfg_fm_step: empty latent");
  const bool vec8 = hw % 8 == 0 && aligned16(latents) && aligned16(noise_pos) && (!noise_neg || aligned16(noise_neg)) &&
                    (!first_frame || aligned16(first_frame));
  const int vec = vec8 ? 8 : 1;
  const int64_t total = static_cast<int64_t>(channels) * frames * (hw / vec);
  int grid = grid_1d(total, 256);
  if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
  cfg_fm_step_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<bf16*>(latents), static_cast<const bf16*>(noise_pos), static_cast<const bf16*>(noise_neg),
      static_cast<const bf16*>(first_frame), cfg_scale, sigma_delta, channels, frames, hw, vec);
  FGB_LAUNCH_CHECK("cfg_fm_step_kernel");
  return FGB_OK;
}

extern "C" int fgb_patchify_rows(fgb_ctx* ctx, const void* latents, void* rows_out, int64_t ld_rows, int32_t channels,
                                 int32_t gf, int32_t gh, int32_t gw, int32_t token_offset, int32_t rows, void* stream) {
  FGB_CHECK_ARG(ctx && latents && rows_out, "fgb_patchify_rows: NULL argument");
  FGB_CHECK_ARG(channels > 0 && gf > 0 && gh > 0 && gw > 0 && rows > 0 && token_offset >= 0, "fgb_patchify_rows: bad shape");
  FGB_CHECK_ARG(ld_rows >= channels * 4 && ld_rows % 4 == 0 && (reinterpret_cast<uintptr_t>(rows_out) & 7) == 0 &&
                    (reinterpret_cast<uintptr_t>(latents) & 3) == 0,
                "fgb_patchify_rows: alignment");
  const int64_t total = static_cast<int64_t>(rows) * channels;
  patchify_rows_kernel<<<grid_1d(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(latents), static_cast<bf16*>(rows_out), ld_rows, channels, gf, gh, gw, token_offset, rows);
  FGB_LAUNCH_CHECK("patchify_rows_kernel");
  return FGB_OK;
}

extern "C" int fgb_unpatchify(fgb_ctx* ctx, const void* head_rows, int64_t ld_rows, void* out, int32_t channels,
                              int32_t gf, int32_t gh, int32_t gw, void* stream) {
  FGB_CHECK_ARG(ctx && head_rows && out, "fgb_unpatchify: NULL argument");
  FGB_CHECK_ARG(channels > 0 && gf > 0 && gh > 0 && gw > 0 && ld_rows >= 4 * channels, "fgb_unpatchify: bad shape");
  FGB_CHECK_ARG((reinterpret_cast<uintptr_t>(out) & 3) == 0, "fgb_unpatchify: out must be 4-byte aligned");
  const int64_t total = static_cast<int64_t>(channels) * gf * 2 * gh * gw;
  unpatchify_kernel<<<grid_1d(total, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(head_rows), ld_rows, static_cast<bf16*>(out), channels, gf, gh, gw);
  FGB_LAUNCH_CHECK("unpatchify_kernel");
  return FGB_OK;
}

extern "C" int fgb_sinusoidal_embedding(fgb_ctx* ctx, const void* timesteps_f32, void* out, int32_t rows, int32_t dim,
                                        void* stream) {
  FGB_CHECK_ARG(ctx && timesteps_f32 && out, "fgb_sinusoidal_embedding: NULL argument");
  FGB_CHECK_ARG(rows > 0 && dim > 0 && dim % 2 == 0, "fgb_sinusoidal_embedding: rows=%d dim=%d", rows, dim);
  sinusoid_kernel<<<grid_1d(static_cast<int64_t>(rows) * dim, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const float*>(timesteps_f32), static_cast<bf16*>(out), rows, dim);
  FGB_LAUNCH_CHECK("sinusoid_kernel");
  return FGB_OK;
}

extern "C" int fgb_silu(fgb_ctx* ctx, const void* x, void* y, int64_t n, void* stream) {
  FGB_CHECK_ARG(ctx && x && y && n > 0, "fgb_silu: bad argument");
  silu_kernel<<<grid_1d(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(x),
                                                                                static_cast<bf16*>(y), n);
  FGB_LAUNCH_CHECK("silu_kernel");
  return FGB_OK;
}

extern "C" int fgb_add_bcast(fgb_ctx* ctx, const void* a, const void* b, void* out, int64_t rows, int64_t cols,
                             int64_t period, void* stream) {
  FGB_CHECK_ARG(ctx && a && b && out && rows > 0 && cols > 0 && period > 0, "fgb_add_bcast: bad argument");
  const int64_t n = rows * cols;
  add_bcast_kernel<<<grid_1d(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(a), static_cast<const bf16*>(b), static_cast<bf16*>(out), n, cols, period);
  FGB_LAUNCH_CHECK("add_bcast_kernel");
  return FGB_OK;
}

static int sp_heads(fgb_ctx* ctx, bool pack, void* x, int64_t ldx, void* buf, int32_t s_local, int32_t heads,
                    int32_t groups, int32_t world, void* stream) {
  FGB_CHECK_ARG(ctx && x && buf, "fgb_sp_(un)pack_heads: NULL argument");
  FGB_CHECK_ARG(s_local > 0 && heads > 0 && groups > 0 && world > 0 && heads % world == 0,
                "fgb_sp_(un)pack_heads: heads=%d must divide by world=%d", heads, world);
  FGB_CHECK_ARG(ldx % 8 == 0 && ldx >= static_cast<int64_t>(groups) * heads * 128 && aligned16(x) && aligned16(buf),
                "fgb_sp_(un)pack_heads: alignment / leading dimension");
  const int64_t total = static_cast<int64_t>(s_local) * groups * heads * 16;
  int grid = grid_1d(total, 256);
  if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (pack)
    sp_heads_kernel<true><<<grid, 256, 0, s>>>(static_cast<bf16*>(x), ldx, static_cast<bf16*>(buf), s_local, heads, groups, world);
  else
    sp_heads_kernel<false><<<grid, 256, 0, s>>>(static_cast<bf16*>(x), ldx, static_cast<bf16*>(buf), s_local, heads, groups, world);
  FGB_LAUNCH_CHECK("sp_heads_kernel");
  return FGB_OK;
}

extern "C" int fgb_sp_pack_heads(fgb_ctx* ctx, const void* x, int64_t ldx, void* send, int32_t s_local, int32_t heads,
                                 int32_t groups, int32_t world, void* stream) {
  return sp_heads(ctx, true, const_cast<void*>(x), ldx, send, s_local, heads, groups, world, stream);
}

extern "C" int fgb_sp_unpack_heads(fgb_ctx* ctx, const void* recv, void* x, int64_t ldx, int32_t s_local, int32_t heads,
                                   int32_t groups, int32_t world, void* stream) {
  return sp_heads(ctx, false, x, ldx, const_cast<void*>(recv), s_local, heads, groups, world, stream);
}

extern "C" int fgb_sp_scatter_heads(fgb_ctx* ctx, const void* x, int64_t ldx, void* const* peer_bufs, int32_t s_local,
                                    int32_t heads, int32_t groups, int32_t group_first, int32_t groups_total, int32_t world,
                                    int32_t rank, void* stream) {
  FGB_CHECK_ARG(ctx && x && peer_bufs, "fgb_sp_scatter_heads: NULL argument");
  FGB_CHECK_ARG(group_first >= 0 && group_first + groups <= groups_total, "fgb_sp_scatter_heads: groups [%d,%d) of %d", group_first,
                group_first + groups, groups_total);
  FGB_CHECK_ARG(s_local > 0 && heads > 0 && groups > 0 && world > 0 && world <= FGB_MAX_PEERS && heads % world == 0 && rank >= 0 &&
                    rank < world, "fgb_sp_scatter_heads: heads=%d world=%d rank=%d", heads, world, rank);
  FGB_CHECK_ARG(ldx % 8 == 0 && ldx >= static_cast<int64_t>(groups) * heads * 128 && aligned16(x), "fgb_sp_scatter_heads: alignment");
  PeerPtrs pp;
  for (int i = 0; i < FGB_MAX_PEERS; ++i) pp.p[i] = i < world ? peer_bufs[i] : nullptr;
  for (int i = 0; i < world; ++i) FGB_CHECK_ARG(pp.p[i] && aligned16(pp.p[i]), "fgb_sp_scatter_heads: peer buffer %d", i);
  const int64_t total = static_cast<int64_t>(s_local) * groups * heads * 16;
  int grid = grid_1d(total, 256);
  if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
  sp_scatter_heads_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(x), ldx, pp, s_local, heads,
                                                                             groups, world, rank, group_first, groups_total);
  FGB_LAUNCH_CHECK("sp_scatter_heads_kernel");
  return FGB_OK;
}

extern "C" int fgb_sp_barrier_status(fgb_ctx* ctx, void* const* peer_flags, int32_t world, int32_t rank, int32_t epoch, void* status,
                                     int64_t timeout_clocks, void* stream) {
  FGB_CHECK_ARG(ctx && peer_flags && world > 0 && world <= FGB_MAX_PEERS && rank >= 0 && rank < world, "fgb_sp_barrier: bad argument");
  FGB_CHECK_ARG(timeout_clocks >= 0, "fgb_sp_barrier: timeout_clocks=%lld", static_cast<long long>(timeout_clocks));
  PeerPtrs pp;
  for (int i = 0; i < FGB_MAX_PEERS; ++i) pp.p[i] = i < world ? peer_flags[i] : nullptr;
  for (int i = 0; i < world; ++i) FGB_CHECK_ARG(pp.p[i], "fgb_sp_barrier: peer flag array %d is NULL", i);
  sp_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(pp, world, rank, epoch, static_cast<int*>(status),
                                                                     timeout_clocks > 0 ? timeout_clocks : kBarrierClocks);
  FGB_LAUNCH_CHECK("sp_barrier_kernel");
  return FGB_OK;
}

extern "C" int fgb_sp_barrier(fgb_ctx* ctx, void* const* peer_flags, int32_t world, int32_t rank, int32_t epoch, void* stream) {
  return fgb_sp_barrier_status(ctx, peer_flags, world, rank, epoch, nullptr, 0, stream);
}

extern "C" int fgb_rmsnorm_rope_scatter(fgb_ctx* ctx, const void* x, int64_t ldx, int32_t rows, int32_t dim, float eps,
                                        const void* weight, const void* rope_tab, int32_t gf, int32_t gh, int32_t gw,
                                        int32_t token_offset, void* const* peer_bufs, int32_t world, int32_t rank, int32_t group,
                                        int32_t groups_total, void* stream) {
  FGB_CHECK_ARG(ctx && x && weight && peer_bufs, "fgb_rmsnorm_rope_scatter: NULL argument");
  FGB_CHECK_ARG(rows > 0 && dim > 0 && dim % 256 == 0, "fgb_rmsnorm_rope_scatter: rows=%d dim=%d", rows, dim);
  FGB_CHECK_ARG(ldx % 8 == 0 && aligned16(x) && aligned16(weight), "fgb_rmsnorm_rope_scatter: operands must be 16-byte aligned");
  const int heads = dim / 128;
  FGB_CHECK_ARG(world > 0 && world <= FGB_MAX_PEERS && heads % world == 0 && rank >= 0 && rank < world && group >= 0 && group < groups_total,
                "fgb_rmsnorm_rope_scatter: heads=%d world=%d rank=%d group=%d/%d", heads, world, rank, group, groups_total);
  if (rope_tab)
    FGB_CHECK_ARG(gf > 0 && gh > 0 && gw > 0 && gf <= 1024 && gh <= 1024 && gw <= 1024 && token_offset >= 0,
                  "fgb_rmsnorm_rope_scatter: grid (%d,%d,%d) outside the RoPE table", gf, gh, gw);
  ScatterSpec sc{};
  for (int i = 0; i < world; ++i) {
    FGB_CHECK_ARG(peer_bufs[i] && aligned16(peer_bufs[i]), "fgb_rmsnorm_rope_scatter: peer buffer %d", i);
    sc.peers.p[i] = peer_bufs[i];
  }
  sc.s_local = rows;
  sc.heads = heads;
  sc.world = world;
  sc.rank = rank;
  sc.grp = group;
  sc.groups = groups_total;
  dim3 grid((rows + kRowWarps - 1) / kRowWarps);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  bf16* xp = const_cast<bf16*>(static_cast<const bf16*>(x));
  const bf16* wp = static_cast<const bf16*>(weight);
  const float2* tab = static_cast<const float2*>(rope_tab);
#define FGB_RMS_CASE(NV)                                                                                                    \
  case NV:                                                                                                                  \
    rmsnorm_rope_kernel<NV, true><<<grid, kRowWarps * 32, 0, s>>>(xp, ldx, rows, eps, wp, tab, gf, gh, gw, token_offset, sc); \
    break;
  switch (dim / 256) {
    FGB_RMS_CASE(1) FGB_RMS_CASE(2) FGB_RMS_CASE(3) FGB_RMS_CASE(4) FGB_RMS_CASE(6) FGB_RMS_CASE(8) FGB_RMS_CASE(12) FGB_RMS_CASE(16)
    FGB_RMS_CASE(20)
    default:
      return set_error(FGB_ERR_UNSUPPORTED, "rmsnorm: dim %d is not one of 256*{1,2,3,4,6,8,12,16,20}", dim);
  }
#undef FGB_RMS_CASE
  FGB_LAUNCH_CHECK("rmsnorm_rope_kernel<scatter>");
  return FGB_OK;
}

extern "C" int fgb_head_norm_max(fgb_ctx* ctx, const void* x, int64_t ldx, int32_t rows, int32_t heads, void* out_f32, void* stream) {
  FGB_CHECK_ARG(ctx && x && out_f32 && rows > 0 && heads > 0, "fgb_head_norm_max: bad argument");
  FGB_CHECK_ARG(ldx % 8 == 0 && ldx >= static_cast<int64_t>(heads) * 128 && aligned16(x), "fgb_head_norm_max: x must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FGB_CUDA(cudaMemsetAsync(out_f32, 0, sizeof(float) * heads, s));
  int grid = (rows + 7) / 8;
  if (grid > ctx->sm_count * 4) grid = ctx->sm_count * 4;
  const bf16* xp = static_cast<const bf16*>(x);
  float* op = static_cast<float*>(out_f32);
#define FGB_HNM_CASE(NV)                                                     \
  case NV:                                                                   \
    head_norm_max_kernel<NV><<<grid, 256, 0, s>>>(xp, ldx, rows, heads, op); \
    break;
  switch ((heads + 1) / 2) {
    FGB_HNM_CASE(1) FGB_HNM_CASE(2) FGB_HNM_CASE(3) FGB_HNM_CASE(4) FGB_HNM_CASE(6) FGB_HNM_CASE(8) FGB_HNM_CASE(12) FGB_HNM_CASE(16)
    FGB_HNM_CASE(20)
    default:
      return set_error(FGB_ERR_UNSUPPORTED, "fgb_head_norm_max: unsupported head count %d", heads);
  }
#undef FGB_HNM_CASE
  FGB_LAUNCH_CHECK("head_norm_max_kernel");
  return FGB_OK;
}

extern "C" int fgb_qk_norm_rope(fgb_ctx* ctx, void* qkv, int64_t ld, int32_t rows, int32_t dim, float eps, const void* wq, const void* wk,
                               const void* rope_tab, int32_t gf, int32_t gh, int32_t gw, int32_t token_offset, void* kmax2,
                               void* qmax2, void* stream) {
  FGB_CHECK_ARG(ctx && qkv && wq && wk && kmax2, "fgb_qk_norm_rope: NULL argument");
  FGB_CHECK_ARG(rows > 0 && dim > 0 && dim % 256 == 0 && ld >= 2 * static_cast<int64_t>(dim) && ld % 8 == 0, "fgb_qk_norm_rope: rows=%d dim=%d",
                rows, dim);
  FGB_CHECK_ARG(aligned16(qkv) && aligned16(wq) && aligned16(wk), "fgb_qk_norm_rope: operands must be 16-byte aligned");
  if (rope_tab)
    FGB_CHECK_ARG(gf > 0 && gh > 0 && gw > 0 && gf <= 1024 && gh <= 1024 && gw <= 1024 && token_offset >= 0,
                  "fgb_qk_norm_rope: grid (%d,%d,%d) outside the RoPE table", gf, gh, gw);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FGB_CUDA(cudaMemsetAsync(kmax2, 0, sizeof(float) * (dim / 128), s));
  if (qmax2) FGB_CUDA(cudaMemsetAsync(qmax2, 0, sizeof(float) * (dim / 128), s));
  float* qp = static_cast<float*>(qmax2);
  const int grid = (rows + kRowWarps - 1) / kRowWarps;   // one row per warp: many short CTAs keep more loads in flight than a grid-stride loop
  bf16* xp = static_cast<bf16*>(qkv);
  const bf16* wqp = static_cast<const bf16*>(wq);
  const bf16* wkp = static_cast<const bf16*>(wk);
  const float2* tab = static_cast<const float2*>(rope_tab);
  float* kp = static_cast<float*>(kmax2);
  if (rows >= kStreamMinRows && stream_enabled(1)) {
#define FGB_QKS_CASE(NV)                                                                                        \
  case NV: {                                                                                                    \
    RmsStreamOp<NV, 2> op{xp, ld, eps, wqp, wkp, tab, gf, gh, gw, token_offset, kp, qp, {0.f}, {0.f}};                      \
    return launch_row_stream<NV>(ctx, s, 2 * rows, op);                                                         \
  }
    switch (dim / 256) {
      FGB_QKS_CASE(1) FGB_QKS_CASE(2) FGB_QKS_CASE(3) FGB_QKS_CASE(4) FGB_QKS_CASE(6) FGB_QKS_CASE(8) FGB_QKS_CASE(12) FGB_QKS_CASE(16)
      FGB_QKS_CASE(20)
      default: break;
    }
#undef FGB_QKS_CASE
  }
#define FGB_QK_CASE(NV)                                                                                                            \
  case NV:                                                                                                                         \
    if (qp) qk_norm_rope_kernel<NV, true><<<grid, kRowWarps * 32, 0, s>>>(xp, ld, rows, eps, wqp, wkp, tab, gf, gh, gw, token_offset, kp, qp); \
    else qk_norm_rope_kernel<NV, false><<<grid, kRowWarps * 32, 0, s>>>(xp, ld, rows, eps, wqp, wkp, tab, gf, gh, gw, token_offset, kp, qp); \
    break;
  switch (dim / 256) {
    FGB_QK_CASE(1) FGB_QK_CASE(2) FGB_QK_CASE(3) FGB_QK_CASE(4) FGB_QK_CASE(6) FGB_QK_CASE(8) FGB_QK_CASE(12) FGB_QK_CASE(16) FGB_QK_CASE(20)
    default:
      return set_error(FGB_ERR_UNSUPPORTED, "fgb_qk_norm_rope: dim %d is not one of 256*{1,2,3,4,6,8,12,16,20}", dim);
  }
#undef FGB_QK_CASE
  FGB_LAUNCH_CHECK("qk_norm_rope_kernel");
  return FGB_OK;
}

extern "C" int fgb_sp_stats_barrier(fgb_ctx* ctx, void* const* peer_flags, void* const* peer_stats, void* rowsq, int32_t rows,
                                    int32_t s_pad, void* kmax2, int32_t hpr, int32_t world, int32_t rank, int32_t epoch, void* status,
                                    void* stream) {
  FGB_CHECK_ARG(ctx && peer_flags && peer_stats && rowsq && kmax2, "fgb_sp_stats_barrier: NULL argument");
  FGB_CHECK_ARG(world > 0 && world <= FGB_MAX_PEERS && rank >= 0 && rank < world && rows > 0 && s_pad == rows * world && hpr > 0 && hpr <= 12,
                "fgb_sp_stats_barrier: rows=%d s_pad=%d hpr=%d world=%d rank=%d", rows, s_pad, hpr, world, rank);
  PeerPtrs fl, st;
  for (int i = 0; i < FGB_MAX_PEERS; ++i) {
    fl.p[i] = i < world ? peer_flags[i] : nullptr;
    st.p[i] = i < world ? peer_stats[i] : nullptr;
  }
  for (int i = 0; i < world; ++i) FGB_CHECK_ARG(fl.p[i] && st.p[i], "fgb_sp_stats_barrier: peer %d", i);
  sp_stats_barrier_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(fl, st, static_cast<float*>(rowsq), rows, s_pad,
                                                                            static_cast<float*>(kmax2), hpr, world, rank, epoch,
                                                                            static_cast<int*>(status));
  FGB_LAUNCH_CHECK("sp_stats_barrier_kernel");
  return FGB_OK;
}

extern "C" int fgb_recv_norm_rope(fgb_ctx* ctx, void* recv, int32_t s_pad, int32_t tokens, int32_t hpr, const void* stats, int32_t dim,
                                  float eps, const void* wq, const void* wk, const void* rope_tab, int32_t gf, int32_t gh, int32_t gw,
                                  void* kmax2, void* qmax2, void* stream) {
  FGB_CHECK_ARG(ctx && recv && stats && wq && wk && kmax2, "fgb_recv_norm_rope: NULL argument");
  if (qmax2) FGB_CUDA(cudaMemsetAsync(qmax2, 0, sizeof(float) * hpr, static_cast<cudaStream_t>(stream)));
  FGB_CHECK_ARG(s_pad > 0 && tokens > 0 && tokens <= s_pad && hpr > 0 && hpr <= 12 && dim > 0, "fgb_recv_norm_rope: s_pad=%d tokens=%d hpr=%d",
                s_pad, tokens, hpr);
  FGB_CHECK_ARG(aligned16(recv) && aligned16(wq) && aligned16(wk), "fgb_recv_norm_rope: operands must be 16-byte aligned");
  if (rope_tab)
    FGB_CHECK_ARG(gf > 0 && gh > 0 && gw > 0 && gf <= 1024 && gh <= 1024 && gw <= 1024, "fgb_recv_norm_rope: grid (%d,%d,%d) outside the RoPE table",
                  gf, gh, gw);
  int grid = (s_pad + 7) / 8;
  if (grid > ctx->sm_count * 6) grid = ctx->sm_count * 6;
  recv_norm_rope_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<bf16*>(recv), s_pad, tokens, hpr, static_cast<const float*>(stats), dim, eps, static_cast<const bf16*>(wq),
      static_cast<const bf16*>(wk), static_cast<const float2*>(rope_tab), gf, gh, gw, static_cast<float*>(kmax2), static_cast<float*>(qmax2));
  FGB_LAUNCH_CHECK("recv_norm_rope_kernel");
  return FGB_OK;
}

extern "C" int fgb_sp_return_heads(fgb_ctx* ctx, const void* x, int64_t ldx, void* const* peer_bufs, int64_t ld_dst, int32_t rows,
                                   int32_t s_pad, int32_t heads, int32_t groups, int32_t world, int32_t rank, void* stream) {
  FGB_CHECK_ARG(ctx && x && peer_bufs, "fgb_sp_return_heads: NULL argument");
  FGB_CHECK_ARG(rows > 0 && s_pad == rows * world && heads > 0 && groups > 0 && world > 0 && world <= FGB_MAX_PEERS && heads % world == 0 &&
                    rank >= 0 && rank < world, "fgb_sp_return_heads: rows=%d s_pad=%d heads=%d world=%d rank=%d", rows, s_pad, heads, world, rank);
  FGB_CHECK_ARG(ldx % 8 == 0 && ldx >= static_cast<int64_t>(groups) * (heads / world) * 128 && aligned16(x) && ld_dst % 8 == 0 &&
                    ld_dst >= static_cast<int64_t>(groups) * heads * 128, "fgb_sp_return_heads: alignment / leading dimensions");
  PeerPtrs pp;
  for (int i = 0; i < FGB_MAX_PEERS; ++i) pp.p[i] = i < world ? peer_bufs[i] : nullptr;
  for (int i = 0; i < world; ++i) FGB_CHECK_ARG(pp.p[i] && aligned16(pp.p[i]), "fgb_sp_return_heads: peer buffer %d", i);
  const int64_t total = static_cast<int64_t>(s_pad) * groups * (heads / world) * 16;
  int grid = grid_1d(total, 256);
  if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
  sp_return_heads_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(x), ldx, pp, ld_dst, rows, s_pad,
                                                                            heads, groups, world, rank);
  FGB_LAUNCH_CHECK("sp_return_heads_kernel");
  return FGB_OK;
}
