// fairygen_b200 — memory-bound kernels of the stage-2 motion-LoRA fine-tune step (BASELINE config 5).
//
// The reference trains only the `lora_B2` matrices of the 300 adapted Linears (animation/diffsynth/diffusion/
// training_module.py:266-352, "TMOD") with the flow-matching SFT loss (diffusion/loss.py:5-21, "LOSS") and gets
// every gradient from torch autograd. Here the backward of each DiTBlock op (models/wan_video_dit.py, "DIT") is one
// fused pass over HBM, mirroring the forward kernels in elementwise.cu: one warp per row, the row held in registers,
// 16-byte accesses, warp-shuffle reductions. Gradients are stored in bf16 like autograd's (the activations are bf16),
// reductions and the B2 gradient accumulate in fp32.
#include <mma.h>

#include "common.cuh"
#include "host.h"

namespace fgb {

constexpr int kRowWarpsT = 4;

// ---------------------------------------------------------------------------------------------
// LayerNorm backward (+ modulate scale or affine weight) + residual-gradient add.
//   forward (DIT:63-64, 205-207, 224-227):  y = n * g + shift,  n = (x - mean) * rstd,  g = 1 + scale  or  weight
//   dn = dy * g;   dx = rstd * (dn - mean(dn) - n * mean(dn * n));   out = dres + dx
// ---------------------------------------------------------------------------------------------
template <int NV, bool AFFINE>
__global__ void __launch_bounds__(kRowWarpsT * 32)
ln_bwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ dy, int64_t ld_dy,
              const __nv_bfloat16* dres, int64_t ld_dres, __nv_bfloat16* out, int64_t ld_out, int rows, float eps,
              const __nv_bfloat16* __restrict__ g0, const __nv_bfloat16* __restrict__ g1, int rows_mod0) {
  const int row = blockIdx.x * kRowWarpsT + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  constexpr int D = NV * 256;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx);
  const uint4* dyr = reinterpret_cast<const uint4*>(dy + static_cast<int64_t>(row) * ld_dy);
  uint4 v[NV], dv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = ldg_nc_v4(xr + i * 32 + lane);
    dv[i] = ldg_nc_v4(dyr + i * 32 + lane);
  }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8];
    unpack8(v[i], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) sum += f[e];
  }
  const float mean = warp_sum(sum) * (1.0f / D);
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8];
    unpack8(v[i], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) sq += (f[e] - mean) * (f[e] - mean);
  }
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
  const uint4* gp = reinterpret_cast<const uint4*>((AFFINE || row < rows_mod0) ? g0 : g1);
  float s1 = 0.f, s2 = 0.f;  // sum(dn), sum(dn * n)
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8], d[8], g[8];
    unpack8(v[i], f);
    unpack8(dv[i], d);
    unpack8(__ldg(gp + i * 32 + lane), g);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float gg = AFFINE ? g[e] : round_bf16(1.0f + g[e]);
      const float dn = d[e] * gg;
      s1 += dn;
      s2 += dn * (f[e] - mean) * rstd;
    }
  }
  s1 = warp_sum(s1) * (1.0f / D);
  s2 = warp_sum(s2) * (1.0f / D);
  const uint4* rr = dres ? reinterpret_cast<const uint4*>(dres + static_cast<int64_t>(row) * ld_dres) : nullptr;
  uint4* orow = reinterpret_cast<uint4*>(out + static_cast<int64_t>(row) * ld_out);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8], d[8], g[8], r[8];
    unpack8(v[i], f);
    unpack8(dv[i], d);
    unpack8(__ldg(gp + i * 32 + lane), g);
    if (rr) unpack8(rr[i * 32 + lane], r);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float gg = AFFINE ? g[e] : round_bf16(1.0f + g[e]);
      const float n = (f[e] - mean) * rstd;
      const float dx = rstd * (d[e] * gg - s1 - n * s2);
      f[e] = rr ? r[e] + round_bf16(dx) : dx;
    }
    orow[i * 32 + lane] = pack8(f);
  }
}

// ---------------------------------------------------------------------------------------------
// RMSNorm (+ 3-D RoPE) backward, in place on dy.      forward: DIT:91-110 (rmsnorm_rope_kernel)
//   y = rot(n * w),  n = x * rs,  rs = rsqrt(mean(x^2) + eps)
//   d(nw) = rot^-1(dy);  dn = d(nw) * w;  dx = rs * (dn - n * mean(dn * n))
// ---------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(kRowWarpsT * 32)
rmsnorm_rope_bwd_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ dy, int64_t ld_dy,
                        int rows, float eps, const __nv_bfloat16* __restrict__ weight, const float2* __restrict__ rope_tab,
                        int gf, int gh, int gw, int token_offset) {
  const int row = blockIdx.x * kRowWarpsT + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  constexpr int D = NV * 256;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx);
  uint4* dr = reinterpret_cast<uint4*>(dy + static_cast<int64_t>(row) * ld_dy);
  uint4 v[NV], dv[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = ldg_nc_v4(xr + i * 32 + lane);
    dv[i] = dr[i * 32 + lane];
  }
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8];
    unpack8(v[i], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) sq += f[e] * f[e];
  }
  const float rs = rsqrtf(warp_sum(sq) * (1.0f / D) + eps);
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
  const uint4* wr = reinterpret_cast<const uint4*>(weight);
  // dn = rot^-1(dy) * w is formed twice (once for the reduction, once for the result) instead of being kept:
  // the row and its gradient already occupy 96 registers at D = 3072
  auto grad_n = [&](int i, float (&d)[8]) {
    float w[8];
    unpack8(dv[i], d);
    unpack8(__ldg(wr + i * 32 + lane), w);
    if (rotate) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {  // transpose of the rotation
        const float a = d[2 * j], b = d[2 * j + 1];
        d[2 * j] = a * cs[j] + b * sn[j];
        d[2 * j + 1] = -a * sn[j] + b * cs[j];
      }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e] *= w[e];
  };
  float s2 = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8], d[8];
    unpack8(v[i], f);
    grad_n(i, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) s2 += d[e] * f[e] * rs;
  }
  s2 = warp_sum(s2) * (1.0f / D);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float f[8], d[8];
    unpack8(v[i], f);
    grad_n(i, d);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] = rs * (d[e] - f[e] * rs * s2);
    dr[i * 32 + lane] = pack8(f);
  }
}

// ---------------------------------------------------------------------------------------------
// GELU(tanh) forward / backward on the FFN hidden (DIT:208), gate scaling of a gradient (DIT:192-193)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float tanh_fast(float x) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x));
  return t;
}

template <bool BWD>
__global__ void gelu_kernel(const __nv_bfloat16* __restrict__ z, const __nv_bfloat16* __restrict__ dh,
                            __nv_bfloat16* __restrict__ out, int64_t n8) {
  const float kAlpha = 0.7978845608028654f, kBeta = 0.044715f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float x[8], d[8];
    unpack8(ldg_nc_v4(reinterpret_cast<const uint4*>(z) + i), x);
    if (BWD) unpack8(ldg_nc_v4(reinterpret_cast<const uint4*>(dh) + i), d);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float t = tanh_fast(kAlpha * (x[e] + kBeta * x[e] * x[e] * x[e]));
      if (BWD) {
        const float dt = (1.0f - t * t) * kAlpha * (1.0f + 3.0f * kBeta * x[e] * x[e]);
        x[e] = d[e] * (0.5f * (1.0f + t) + 0.5f * x[e] * dt);
      } else {
        x[e] = 0.5f * x[e] * (1.0f + t);
      }
    }
    reinterpret_cast<uint4*>(out)[i] = pack8(x);
  }
}

__global__ void mul_gate_kernel(const __nv_bfloat16* __restrict__ dx, int64_t ld_dx, __nv_bfloat16* __restrict__ out,
                                int64_t ld_out, int rows, int dim8, const __nv_bfloat16* __restrict__ gate0,
                                const __nv_bfloat16* __restrict__ gate1, int rows_gate0) {
  const int64_t total = static_cast<int64_t>(rows) * dim8;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / dim8), c = static_cast<int>(i % dim8);
    float d[8], g[8];
    unpack8(ldg_nc_v4(reinterpret_cast<const uint4*>(dx + static_cast<int64_t>(r) * ld_dx) + c), d);
    unpack8(__ldg(reinterpret_cast<const uint4*>(r < rows_gate0 ? gate0 : gate1) + c), g);
#pragma unroll
    for (int e = 0; e < 8; ++e) d[e] *= g[e];
    reinterpret_cast<uint4*>(out + static_cast<int64_t>(r) * ld_out)[c] = pack8(d);
  }
}

// ---------------------------------------------------------------------------------------------
// Stage-2 LoRA (TMOD:317-352):  y = W x + b + s * B1 (A1 x) + s * (B2 * mask * 2) (A1 x),  only B2 trainable.
//   merge:  W_eff = W + s * (B1 + B2 * mask * 2) A1        (one bf16 weight per step; forward AND dgrad use it)
//   wgrad:  dB2 += (dyᵀ t) * mask * 2 * s,  t = A1 x       (rank-R GEMM over the tokens, fp32 accumulate)
// Both are memory-bound rank-32 updates: warp-level mma (wmma bf16 m16n16k16) keeps them at HBM speed without
// occupying the tcgen05 pipeline that the big GEMMs use.
// ---------------------------------------------------------------------------------------------
using namespace nvcuda;

template <int NR>  // R = 16 * NR
__global__ void __launch_bounds__(256)
lora_merge_kernel(const __nv_bfloat16* __restrict__ w, int64_t ldw, const __nv_bfloat16* __restrict__ a1, int64_t lda,
                  const __nv_bfloat16* __restrict__ b1, const __nv_bfloat16* __restrict__ b2,
                  const uint8_t* __restrict__ mask, float mask_mul, float scaling, __nv_bfloat16* __restrict__ w_eff,
                  int64_t ld_eff, int n, int k) {
  constexpr int R = 16 * NR;
  // operands and the fp32 result tile share one buffer (the result overwrites the operands after the last MMA)
  constexpr int kOperandBytes = (64 * R + R * 128) * 2, kAccBytes = 64 * 136 * 4;
  __shared__ __align__(32) unsigned char s_raw[kOperandBytes > kAccBytes ? kOperandBytes : kAccBytes];
  __nv_bfloat16* s_b = reinterpret_cast<__nv_bfloat16*>(s_raw);              // (B1 + B2*mask*2) * scaling, rows n0..n0+63
  __nv_bfloat16* s_a = s_b + 64 * R;                                         // A1[:, k0..k0+127]
  float* s_acc = reinterpret_cast<float*>(s_raw);
  const int n0 = blockIdx.y * 64, k0 = blockIdx.x * 128;
  // every global read of this tile is issued up front (one round trip): the W tile into registers, A1 and B into smem
  uint4 wv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = threadIdx.x + j * 256;
    wv[j] = ldg_nc_v4(w + static_cast<int64_t>(n0 + (i >> 4)) * ldw + k0 + (i & 15) * 8);
  }
  for (int i = threadIdx.x; i < R * 16; i += 256)
    *reinterpret_cast<uint4*>(s_a + (i >> 4) * 128 + (i & 15) * 8) = ldg_nc_v4(a1 + static_cast<int64_t>(i >> 4) * lda + k0 + (i & 15) * 8);
  for (int i = threadIdx.x; i < 64 * R; i += 256) {
    const int64_t g = static_cast<int64_t>(n0 + i / R) * R + (i % R);
    float v = b1 ? __bfloat162float(b1[g]) : 0.f;
    if (b2) {
      // B2 * mask -> bf16, * scale_factor -> bf16 (TMOD:343-346), summed with B1 in fp32
      const float m = mask ? static_cast<float>(mask[g]) : 1.0f;
      v += round_bf16(round_bf16(__bfloat162float(b2[g]) * m) * mask_mul);
    }
    s_b[i] = __float2bfloat16_rn(v * scaling);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5;
  const int wn = (warp & 3) * 16, wk = (warp >> 2) * 64;
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) wmma::fill_fragment(acc[j], 0.f);
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::row_major> fa;
    wmma::load_matrix_sync(fa, s_b + wn * R + r * 16, R);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> fb;
      wmma::load_matrix_sync(fb, s_a + r * 16 * 128 + wk + j * 16, 128);
      wmma::mma_sync(acc[j], fa, fb, acc[j]);
    }
  }
  __syncthreads();   // every warp is done reading the operands
#pragma unroll
  for (int j = 0; j < 4; ++j) wmma::store_matrix_sync(s_acc + wn * 136 + wk + j * 16, acc[j], 136, wmma::mem_row_major);
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int i = threadIdx.x + j * 256;
    const int r = i >> 4, c = (i & 15) * 8;
    float f[8];
    unpack8(wv[j], f);
#pragma unroll
    for (int e = 0; e < 8; ++e) f[e] += s_acc[r * 136 + c + e];
    *reinterpret_cast<uint4*>(w_eff + static_cast<int64_t>(n0 + r) * ld_eff + k0 + c) = pack8(f);
  }
}

template <int NR>
__global__ void __launch_bounds__(128)
lora_wgrad_kernel(const __nv_bfloat16* __restrict__ dy, int64_t ld_dy, const __nv_bfloat16* __restrict__ t, int64_t ld_t,
                  float* __restrict__ db, const uint8_t* __restrict__ mask, float mul, int rows, int n, int rows_per_cta,
                  int transpose_out) {
  constexpr int R = 16 * NR;
  constexpr int kChunk = 64;                       // tokens per step
  constexpr int kTV = kChunk * R / 8 / 128;        // 16-byte vectors of t per thread per chunk (R = 16/32/64 -> 1/2/4)
  constexpr int kLdDy = 64 + 8, kLdT = R + 8;     // 16 bytes of padding per row: the transposed (col-major) A fragments
  __shared__ __align__(32) __nv_bfloat16 s_dy[kChunk * kLdDy];   // of wmma would otherwise hit the same banks 16 ways
  __shared__ __align__(32) __nv_bfloat16 s_t[kChunk * kLdT];
  __shared__ __align__(32) float s_out[64 * R];
  const int n0 = blockIdx.x * 64;
  const int s_begin = blockIdx.y * rows_per_cta;
  const int s_end = min(rows, s_begin + rows_per_cta);
  const int warp = threadIdx.x >> 5;
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[NR];
#pragma unroll
  for (int j = 0; j < NR; ++j) wmma::fill_fragment(acc[j], 0.f);
  // register double buffer: the loads of chunk c+1 are in flight while chunk c is multiplied out of shared memory;
  // rows past the end and columns past n are zeros
  uint4 rdy[4], rt[kTV];
  auto fetch = [&](int s0) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = threadIdx.x + j * 128;
      const int r = i >> 3, c = (i & 7) * 8;
      rdy[j] = (s0 + r < s_end && n0 + c < n) ? ldg_nc_v4(dy + static_cast<int64_t>(s0 + r) * ld_dy + n0 + c) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int j = 0; j < kTV; ++j) {
      const int i = threadIdx.x + j * 128;
      const int r = i / (R / 8), c = (i % (R / 8)) * 8;
      rt[j] = (s0 + r < s_end) ? ldg_nc_v4(t + static_cast<int64_t>(s0 + r) * ld_t + c) : make_uint4(0, 0, 0, 0);
    }
  };
  fetch(s_begin);
  for (int s0 = s_begin; s0 < s_end; s0 += kChunk) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int i = threadIdx.x + j * 128;
      *reinterpret_cast<uint4*>(s_dy + (i >> 3) * kLdDy + (i & 7) * 8) = rdy[j];
    }
#pragma unroll
    for (int j = 0; j < kTV; ++j) {
      const int i = threadIdx.x + j * 128;
      *reinterpret_cast<uint4*>(s_t + (i / (R / 8)) * kLdT + (i % (R / 8)) * 8) = rt[j];
    }
    __syncthreads();
    if (s0 + kChunk < s_end) fetch(s0 + kChunk);
#pragma unroll
    for (int ks = 0; ks < kChunk / 16; ++ks) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __nv_bfloat16, wmma::col_major> fa;  // A[n_i, s_k] = dy[s_k, n_i]
      wmma::load_matrix_sync(fa, s_dy + ks * 16 * kLdDy + warp * 16, kLdDy);
#pragma unroll
      for (int j = 0; j < NR; ++j) {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, __nv_bfloat16, wmma::row_major> fb;
        wmma::load_matrix_sync(fb, s_t + ks * 16 * kLdT + j * 16, kLdT);
        wmma::mma_sync(acc[j], fa, fb, acc[j]);
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int j = 0; j < NR; ++j) wmma::store_matrix_sync(s_out + warp * 16 * R + j * 16, acc[j], R, wmma::mem_row_major);
  __syncthreads();
  for (int i = threadIdx.x; i < 64 * R; i += 128) {
    const int r = i / R, c = i % R;
    if (n0 + r >= n) continue;
    const int64_t g = static_cast<int64_t>(n0 + r) * R + c;
    const float m = mask ? static_cast<float>(mask[g]) : 1.0f;
    const float v = s_out[i] * m * mul;
    // transpose_out: db is [R, n] (the layout of lora_A: dA = uᵀ·X computed as (Xᵀ·u)ᵀ)
    if (v != 0.f) atomicAdd(db + (transpose_out ? static_cast<int64_t>(c) * n + (n0 + r) : g), v);
  }
}

// B2eff[n, r] = bf16(bf16(bf16(B2 * mask) * mask_mul) * scaling) into a (possibly strided, e.g. block-diagonal) view
__global__ void lora_b2_eff_kernel(const __nv_bfloat16* __restrict__ b2, const uint8_t* __restrict__ mask, float mask_mul,
                                   float scaling, __nv_bfloat16* __restrict__ out, int64_t ld_out, int64_t n_rows, int rank) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n_rows * rank) return;
  const float m = mask ? static_cast<float>(mask[i]) : 1.0f;
  const float v = round_bf16(round_bf16(__bfloat162float(b2[i]) * m) * mask_mul);
  out[(i / rank) * ld_out + (i % rank)] = __float2bfloat16_rn(v * scaling);
}

// All adapted Linears of the model in one launch: entry e = {offset into the flat B2 / mask buffers, rows, destination pointer,
// destination row stride}; blockIdx.y = entry, grid-stride over its elements.
__global__ void lora_b2_eff_batched_kernel(const __nv_bfloat16* __restrict__ b2, const uint8_t* __restrict__ mask,
                                           const int64_t* __restrict__ table, float mask_mul, float scaling, int rank) {
  const int64_t* e = table + 4 * static_cast<int64_t>(blockIdx.y);
  const int64_t off = e[0], total = e[1] * rank, ld_out = e[3];
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(static_cast<uintptr_t>(e[2]));
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float m = mask ? static_cast<float>(mask[off + i]) : 1.0f;
    const float v = round_bf16(round_bf16(__bfloat162float(b2[off + i]) * m) * mask_mul);
    out[(i / rank) * ld_out + (i % rank)] = __float2bfloat16_rn(v * scaling);
  }
}

// counter-based Bernoulli mask: keep[i] = hash(seed, i) / 2^32 > drop_prob   (torch.rand_like(...) > p, TMOD:342)
__global__ void bernoulli_mask_kernel(uint8_t* __restrict__ out, int64_t n, float drop_prob, uint64_t seed) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * static_cast<uint64_t>(i + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const float u = static_cast<float>(z >> 40) * (1.0f / 16777216.0f);
  out[i] = u > drop_prob ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// Flow-matching SFT loss (LOSS:5-21; FlowMatchScheduler.add_noise / training_target, flow_match.py:164-179)
// ---------------------------------------------------------------------------------------------
__global__ void fm_noise_target_kernel(const __nv_bfloat16* __restrict__ x0, const __nv_bfloat16* __restrict__ noise,
                                       float sigma, __nv_bfloat16* __restrict__ latents, __nv_bfloat16* __restrict__ target,
                                       int64_t n) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const float a = __bfloat162float(x0[i]), b = __bfloat162float(noise[i]);
  latents[i] = __float2bfloat16_rn(round_bf16((1.0f - sigma) * a) + round_bf16(sigma * b));
  target[i] = __float2bfloat16_rn(b - a);
}

__global__ void __launch_bounds__(256)
mse_loss_grad_kernel(const __nv_bfloat16* __restrict__ pred, const __nv_bfloat16* __restrict__ target, float weight,
                     float* __restrict__ loss, __nv_bfloat16* __restrict__ dpred, int64_t n) {
  float part = 0.f;
  const float inv_n = 1.0f / static_cast<float>(n);
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float d = __bfloat162float(pred[i]) - __bfloat162float(target[i]);
    part += d * d;
    if (dpred) dpred[i] = __float2bfloat16_rn(2.0f * weight * inv_n * d);
  }
  part = warp_sum(part);
  __shared__ float s[8];
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += s[i];
    atomicAdd(loss, tot * inv_n * weight);
  }
}

// d_rows[t, y*2C + z*C + c] = dpred[c, f, 2h+y, 2w+z]: adjoint of unpatchify_kernel (DIT:346-351)
__global__ void unpatchify_bwd_kernel(const __nv_bfloat16* __restrict__ dpred, __nv_bfloat16* __restrict__ d_rows,
                                      int64_t ld_rows, int channels, int gf, int gh, int gw) {
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
  const __nv_bfloat162 v =
      *reinterpret_cast<const __nv_bfloat162*>(dpred + ((static_cast<int64_t>(c) * gf + f) * H2 + hy) * (2 * gw) + 2 * w);
  __nv_bfloat16* dst = d_rows + t * ld_rows + y * 2 * channels + c;
  dst[0] = v.x;
  dst[channels] = v.y;
}

// AdamW on the bf16 B2 parameters with fp32 moments (the reference trains with torch.optim.AdamW, train.py runner)
__global__ void adamw_kernel(__nv_bfloat16* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, int64_t n, float lr, float beta1, float beta2, float eps, float wd,
                             float bc1, float bc2) {
  const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  float w = __bfloat162float(p[i]);
  const float gi = g[i];
  const float mi = beta1 * m[i] + (1.0f - beta1) * gi;
  const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  w *= 1.0f - lr * wd;
  w -= lr * (mi / bc1) / (sqrtf(vi / bc2) + eps);
  p[i] = __float2bfloat16_rn(w);
}

static inline int grid_for(int64_t n, int block, int cap) {
  int64_t g = (n + block - 1) / block;
  return static_cast<int>(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace fgb

using namespace fgb;
typedef __nv_bfloat16 bf16;

#define FGB_NV_SWITCH(nv, CASE)                                                                                \
  switch (nv) {                                                                                                \
    CASE(1) CASE(2) CASE(4) CASE(6) CASE(8) CASE(12) CASE(16) CASE(20)                                         \
    default:                                                                                                   \
      return set_error(FGB_ERR_UNSUPPORTED, "dim %d is not one of 256*{1,2,4,6,8,12,16,20}", (nv) * 256);      \
  }

extern "C" int fgb_ln_bwd(fgb_ctx* ctx, const void* x, int64_t ldx, const void* dy, int64_t ld_dy, const void* dres,
                          int64_t ld_dres, void* out, int64_t ld_out, int32_t rows, int32_t dim, float eps,
                          const void* g0, const void* g1, int32_t rows_mod0, int32_t affine, void* stream) {
  FGB_CHECK_ARG(ctx && x && dy && out && g0, "fgb_ln_bwd: NULL argument");
  FGB_CHECK_ARG(rows > 0 && dim > 0 && dim % 256 == 0, "fgb_ln_bwd: rows=%d dim=%d (dim must be a multiple of 256)", rows, dim);
  FGB_CHECK_ARG(ldx % 8 == 0 && ld_dy % 8 == 0 && ld_out % 8 == 0 && (!dres || ld_dres % 8 == 0) && aligned16(x) && aligned16(dy) &&
                    aligned16(out) && aligned16(g0) && (!g1 || aligned16(g1)) && (!dres || aligned16(dres)),
                "fgb_ln_bwd: operands must be 16-byte aligned");
  if (!g1) g1 = g0;
  dim3 grid((rows + kRowWarpsT - 1) / kRowWarpsT);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define CASE(NV)                                                                                                       \
  case NV:                                                                                                             \
    if (affine)                                                                                                        \
      ln_bwd_kernel<NV, true><<<grid, kRowWarpsT * 32, 0, s>>>((const bf16*)x, ldx, (const bf16*)dy, ld_dy, (const bf16*)dres, \
                                                               ld_dres, (bf16*)out, ld_out, rows, eps, (const bf16*)g0,  \
                                                               (const bf16*)g1, rows_mod0);                            \
    else                                                                                                               \
      ln_bwd_kernel<NV, false><<<grid, kRowWarpsT * 32, 0, s>>>((const bf16*)x, ldx, (const bf16*)dy, ld_dy, (const bf16*)dres, \
                                                                ld_dres, (bf16*)out, ld_out, rows, eps, (const bf16*)g0, \
                                                                (const bf16*)g1, rows_mod0);                           \
    break;
  FGB_NV_SWITCH(dim / 256, CASE)
#undef CASE
  FGB_LAUNCH_CHECK("ln_bwd_kernel");
  return FGB_OK;
}

extern "C" int fgb_rmsnorm_rope_bwd(fgb_ctx* ctx, const void* x, int64_t ldx, void* dy, int64_t ld_dy, int32_t rows,
                                    int32_t dim, float eps, const void* weight, const void* rope_tab, int32_t gf, int32_t gh,
                                    int32_t gw, int32_t token_offset, void* stream) {
  FGB_CHECK_ARG(ctx && x && dy && weight, "fgb_rmsnorm_rope_bwd: NULL argument");
  FGB_CHECK_ARG(rows > 0 && dim > 0 && dim % 256 == 0, "fgb_rmsnorm_rope_bwd: rows=%d dim=%d", rows, dim);
  FGB_CHECK_ARG(ldx % 8 == 0 && ld_dy % 8 == 0 && aligned16(x) && aligned16(dy) && aligned16(weight),
                "fgb_rmsnorm_rope_bwd: operands must be 16-byte aligned");
  if (rope_tab)
    FGB_CHECK_ARG(gf > 0 && gh > 0 && gw > 0 && gf <= 1024 && gh <= 1024 && gw <= 1024 && token_offset >= 0,
                  "fgb_rmsnorm_rope_bwd: grid (%d,%d,%d) outside the RoPE table", gf, gh, gw);
  dim3 grid((rows + kRowWarpsT - 1) / kRowWarpsT);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define CASE(NV)                                                                                                    \
  case NV:                                                                                                          \
    rmsnorm_rope_bwd_kernel<NV><<<grid, kRowWarpsT * 32, 0, s>>>((const bf16*)x, ldx, (bf16*)dy, ld_dy, rows, eps,  \
                                                                 (const bf16*)weight, (const float2*)rope_tab, gf, gh, gw, \
                                                                 token_offset);                                     \
    break;
  FGB_NV_SWITCH(dim / 256, CASE)
#undef CASE
  FGB_LAUNCH_CHECK("rmsnorm_rope_bwd_kernel");
  return FGB_OK;
}

extern "C" int fgb_gelu_tanh(fgb_ctx* ctx, const void* z, void* h, int64_t n, void* stream) {
  FGB_CHECK_ARG(ctx && z && h && n > 0 && n % 8 == 0 && aligned16(z) && aligned16(h), "fgb_gelu_tanh: bad argument (n %% 8, 16-byte alignment)");
  gelu_kernel<false><<<grid_for(n / 8, 256, ctx->sm_count * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(z), nullptr, static_cast<bf16*>(h), n / 8);
  FGB_LAUNCH_CHECK("gelu_kernel");
  return FGB_OK;
}

extern "C" int fgb_gelu_tanh_bwd(fgb_ctx* ctx, const void* z, const void* dh, void* dz, int64_t n, void* stream) {
  FGB_CHECK_ARG(ctx && z && dh && dz && n > 0 && n % 8 == 0 && aligned16(z) && aligned16(dh) && aligned16(dz),
                "fgb_gelu_tanh_bwd: bad argument (n %% 8, 16-byte alignment)");
  gelu_kernel<true><<<grid_for(n / 8, 256, ctx->sm_count * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(z), static_cast<const bf16*>(dh), static_cast<bf16*>(dz), n / 8);
  FGB_LAUNCH_CHECK("gelu_kernel");
  return FGB_OK;
}

extern "C" int fgb_mul_gate(fgb_ctx* ctx, const void* dx, int64_t ld_dx, void* out, int64_t ld_out, int32_t rows, int32_t dim,
                            const void* gate0, const void* gate1, int32_t rows_gate0, void* stream) {
  FGB_CHECK_ARG(ctx && dx && out && gate0 && gate1, "fgb_mul_gate: NULL argument");
  FGB_CHECK_ARG(rows > 0 && dim > 0 && dim % 8 == 0 && ld_dx % 8 == 0 && ld_out % 8 == 0 && aligned16(dx) && aligned16(out) &&
                    aligned16(gate0) && aligned16(gate1), "fgb_mul_gate: dim %% 8 and 16-byte alignment required");
  mul_gate_kernel<<<grid_for(static_cast<int64_t>(rows) * (dim / 8), 256, ctx->sm_count * 16), 256, 0,
                    static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(dx), ld_dx, static_cast<bf16*>(out), ld_out,
                                                         rows, dim / 8, static_cast<const bf16*>(gate0),
                                                         static_cast<const bf16*>(gate1), rows_gate0);
  FGB_LAUNCH_CHECK("mul_gate_kernel");
  return FGB_OK;
}

extern "C" int fgb_lora_merge(fgb_ctx* ctx, const void* w, int64_t ldw, const void* a1, int64_t lda, const void* b1,
                              const void* b2, const void* mask, float mask_mul, float scaling, void* w_eff, int64_t ld_eff,
                              int32_t n, int32_t k, int32_t rank, void* stream) {
  FGB_CHECK_ARG(ctx && w && a1 && w_eff && (b1 || b2), "fgb_lora_merge: NULL argument");
  FGB_CHECK_ARG(n > 0 && k > 0 && n % 64 == 0 && k % 128 == 0, "fgb_lora_merge: n=%d must divide by 64 and k=%d by 128", n, k);
  FGB_CHECK_ARG(rank == 16 || rank == 32 || rank == 64, "fgb_lora_merge: rank %d not in {16, 32, 64}", rank);
  FGB_CHECK_ARG(ldw % 8 == 0 && lda % 8 == 0 && ld_eff % 8 == 0 && aligned16(w) && aligned16(a1) && aligned16(w_eff),
                "fgb_lora_merge: operands must be 16-byte aligned");
  dim3 grid(k / 128, n / 64);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define LAUNCH(NR)                                                                                                     \
  lora_merge_kernel<NR><<<grid, 256, 0, s>>>((const bf16*)w, ldw, (const bf16*)a1, lda, (const bf16*)b1, (const bf16*)b2, \
                                             (const uint8_t*)mask, mask_mul, scaling, (bf16*)w_eff, ld_eff, n, k)
  if (rank == 16) LAUNCH(1); else if (rank == 32) LAUNCH(2); else LAUNCH(4);
#undef LAUNCH
  FGB_LAUNCH_CHECK("lora_merge_kernel");
  return FGB_OK;
}

extern "C" int fgb_lora_wgrad(fgb_ctx* ctx, const void* dy, int64_t ld_dy, const void* t, int64_t ld_t, void* db_f32,
                              const void* mask, float mul, int32_t rows, int32_t n, int32_t rank, int32_t transpose_out, void* stream) {
  FGB_CHECK_ARG(ctx && dy && t && db_f32, "fgb_lora_wgrad: NULL argument");
  FGB_CHECK_ARG(rows > 0 && n > 0 && n % 8 == 0, "fgb_lora_wgrad: rows=%d n=%d (n %% 8)", rows, n);
  FGB_CHECK_ARG(rank == 16 || rank == 32 || rank == 64, "fgb_lora_wgrad: rank %d not in {16, 32, 64}", rank);
  FGB_CHECK_ARG(!(transpose_out && mask), "fgb_lora_wgrad: a mask applies to the [n, rank] layout only");
  FGB_CHECK_ARG(ld_dy % 8 == 0 && ld_t % 8 == 0 && aligned16(dy) && aligned16(t), "fgb_lora_wgrad: operands must be 16-byte aligned");
  const int n_tiles = (n + 63) / 64;
  // enough token chunks to fill the machine ~10x over (each CTA keeps only one 12 KB chunk in flight); multiples of 64 rows
  int splits = (10 * ctx->sm_count + n_tiles - 1) / n_tiles;
  int rows_per_cta = ((rows + splits - 1) / splits + 31) / 32 * 32;
  rows_per_cta = (rows_per_cta + 63) / 64 * 64;
  if (rows_per_cta < 128) rows_per_cta = 128;
  splits = (rows + rows_per_cta - 1) / rows_per_cta;
  dim3 grid(n_tiles, splits);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
#define LAUNCH(NR)                                                                                                  \
  lora_wgrad_kernel<NR><<<grid, 128, 0, s>>>((const bf16*)dy, ld_dy, (const bf16*)t, ld_t, (float*)db_f32,          \
                                             (const uint8_t*)mask, mul, rows, n, rows_per_cta, transpose_out)
  if (rank == 16) LAUNCH(1); else if (rank == 32) LAUNCH(2); else LAUNCH(4);
#undef LAUNCH
  FGB_LAUNCH_CHECK("lora_wgrad_kernel");
  return FGB_OK;
}

extern "C" int fgb_bernoulli_mask(fgb_ctx* ctx, void* out_u8, int64_t n, float drop_prob, uint64_t seed, void* stream) {
  FGB_CHECK_ARG(ctx && out_u8 && n > 0 && drop_prob >= 0.f && drop_prob < 1.f, "fgb_bernoulli_mask: bad argument");
  bernoulli_mask_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint8_t*>(out_u8), n, drop_prob, seed);
  FGB_LAUNCH_CHECK("bernoulli_mask_kernel");
  return FGB_OK;
}

extern "C" int fgb_fm_noise_target(fgb_ctx* ctx, const void* x0, const void* noise, float sigma, void* latents, void* target,
                                   int64_t n, void* stream) {
  FGB_CHECK_ARG(ctx && x0 && noise && latents && target && n > 0, "fgb_fm_noise_target: bad argument");
  fm_noise_target_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x0), static_cast<const bf16*>(noise), sigma, static_cast<bf16*>(latents), static_cast<bf16*>(target), n);
  FGB_LAUNCH_CHECK("fm_noise_target_kernel");
  return FGB_OK;
}

extern "C" int fgb_mse_loss_grad(fgb_ctx* ctx, const void* pred, const void* target, float weight, void* loss_f32, void* dpred,
                                 int64_t n, void* stream) {
  FGB_CHECK_ARG(ctx && pred && target && loss_f32 && n > 0, "fgb_mse_loss_grad: bad argument");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  FGB_CUDA(cudaMemsetAsync(loss_f32, 0, sizeof(float), s));
  mse_loss_grad_kernel<<<grid_for(n, 256, ctx->sm_count * 8), 256, 0, s>>>(static_cast<const bf16*>(pred), static_cast<const bf16*>(target),
                                                                         weight, static_cast<float*>(loss_f32), static_cast<bf16*>(dpred), n);
  FGB_LAUNCH_CHECK("mse_loss_grad_kernel");
  return FGB_OK;
}

extern "C" int fgb_unpatchify_bwd(fgb_ctx* ctx, const void* dpred, void* d_rows, int64_t ld_rows, int32_t channels, int32_t gf,
                                  int32_t gh, int32_t gw, void* stream) {
  FGB_CHECK_ARG(ctx && dpred && d_rows, "fgb_unpatchify_bwd: NULL argument");
  FGB_CHECK_ARG(channels > 0 && gf > 0 && gh > 0 && gw > 0 && ld_rows >= 4 * channels, "fgb_unpatchify_bwd: bad shape");
  FGB_CHECK_ARG((reinterpret_cast<uintptr_t>(dpred) & 3) == 0, "fgb_unpatchify_bwd: dpred must be 4-byte aligned");
  const int64_t total = static_cast<int64_t>(channels) * gf * 2 * gh * gw;
  unpatchify_bwd_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(dpred), static_cast<bf16*>(d_rows), ld_rows, channels, gf, gh, gw);
  FGB_LAUNCH_CHECK("unpatchify_bwd_kernel");
  return FGB_OK;
}

extern "C" int fgb_adamw_step(fgb_ctx* ctx, void* param_bf16, const void* grad_f32, void* m_f32, void* v_f32, int64_t n, float lr,
                              float beta1, float beta2, float eps, float weight_decay, int32_t step, void* stream) {
  FGB_CHECK_ARG(ctx && param_bf16 && grad_f32 && m_f32 && v_f32 && n > 0 && step >= 1, "fgb_adamw_step: bad argument");
  const float bc1 = 1.0f - powf(beta1, static_cast<float>(step)), bc2 = 1.0f - powf(beta2, static_cast<float>(step));
  adamw_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<bf16*>(param_bf16), static_cast<const float*>(grad_f32), static_cast<float*>(m_f32), static_cast<float*>(v_f32), n, lr,
      beta1, beta2, eps, weight_decay, bc1, bc2);
  FGB_LAUNCH_CHECK("adamw_kernel");
  return FGB_OK;
}

extern "C" int fgb_lora_b2_eff_batched(fgb_ctx* ctx, const void* b2_flat, const void* mask_flat, const void* table, int32_t n_entries,
                                       int32_t rank, float mask_mul, float scaling, void* stream) {
  FGB_CHECK_ARG(ctx && b2_flat && table && n_entries > 0 && n_entries <= 65535 && rank > 0, "fgb_lora_b2_eff_batched: bad argument");
  lora_b2_eff_batched_kernel<<<dim3(48, n_entries), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(b2_flat), static_cast<const uint8_t*>(mask_flat), static_cast<const int64_t*>(table), mask_mul, scaling, rank);
  FGB_LAUNCH_CHECK("lora_b2_eff_batched_kernel");
  return FGB_OK;
}

extern "C" int fgb_lora_b2_eff(fgb_ctx* ctx, const void* b2, const void* mask, float mask_mul, float scaling, void* out,
                               int64_t ld_out, int64_t n_rows, int32_t rank, void* stream) {
  FGB_CHECK_ARG(ctx && b2 && out && n_rows > 0 && rank > 0 && ld_out >= rank, "fgb_lora_b2_eff: bad argument");
  lora_b2_eff_kernel<<<static_cast<unsigned>((n_rows * rank + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(b2), static_cast<const uint8_t*>(mask), mask_mul, scaling, static_cast<bf16*>(out), ld_out, n_rows, rank);
  FGB_LAUNCH_CHECK("lora_b2_eff_kernel");
  return FGB_OK;
}
