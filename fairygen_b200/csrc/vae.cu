// fairygen_b200 — kernels of the Wan2.2 VAE38 decoder around its convolutions (sm_100a).            SURVEY §8(f) row 1
//
// Reference: animation/diffsynth/models/wan_video_vae.py ("VAE"). Feature maps live channels-last on a zero-bordered grid,
//   G[t][y][x][c],  y in [0, H+2), x in [0, W+2), c in [0, Cp)   (Cp = channels rounded up to 64, padding channels zero),
// so that a causal 3x3x3 convolution is a GEMM over shifted rows of the flattened grid (fgb_conv_taps_bf16 in gemm.cu: one
// K-block group per tap, the tap's row offset added to the TMA coordinate, border rows written as zero by the epilogue) and
// the two cached frames of CausalConv3d (VAE:44-52, 288-301) are simply the two frames stored in front of the new ones.
// This file holds the memory-bound rest: latent de-normalisation into the grid, RMS_norm + SiLU, nearest-exact 2x
// up-sampling (with the frame interleave of the temporal up-sampling), the DupUp3D shortcut, the softmax of the single-head
// attention block, and un-patchify + tile blending + clamp. One pass each, 16-byte accesses along the channels, fp32 math.
#include "common.cuh"
#include "host.h"

namespace fgb {

static inline int vae_grid(int64_t n, int block, int cap) {
  const int64_t g = (n + block - 1) / block;
  return static_cast<int>(g < cap ? (g < 1 ? 1 : g) : cap);
}

// ---------------------------------------------------------------------------------------------
// z / (1/std) + mean (VAE:1328-1331), [C][T][H][W] -> grid interior (border and padding channels are left as they are: zero)
// ---------------------------------------------------------------------------------------------
__global__ void vae_latent_rows_kernel(const __nv_bfloat16* __restrict__ z, const float* __restrict__ mean, const float* __restrict__ inv_std,
                                       __nv_bfloat16* __restrict__ out, int C, int T, int H, int W, int Cp) {
  const int64_t total = static_cast<int64_t>(T) * H * W * C;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % C);
    int64_t r = idx / C;
    const int x = static_cast<int>(r % W);
    r /= W;
    const int y = static_cast<int>(r % H), t = static_cast<int>(r / H);
    const float v = round_bf16(__bfloat162float(z[((static_cast<int64_t>(c) * T + t) * H + y) * W + x]) / inv_std[c]) + mean[c];
    out[((static_cast<int64_t>(t) * (H + 2) + y + 1) * (W + 2) + x + 1) * Cp + c] = __float2bfloat16_rn(v);
  }
}

// ---------------------------------------------------------------------------------------------
// RMS_norm (VAE:67-70: x / max(|x|_2, 1e-12) * sqrt(C) * gamma) followed by SiLU (VAE:274-277). One warp per grid row.
// Zero rows (the border) stay zero, so the output is again a zero-bordered grid.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) vae_norm_silu_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                            int64_t rows, int Cp, float sqrt_c, const __nv_bfloat16* __restrict__ gamma,
                                                            int silu) {
  const int64_t row = blockIdx.x * 4ll + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const uint4* xr = reinterpret_cast<const uint4*>(x + row * Cp);
  const int nvec = Cp / 8;
  // rows of up to 1024 channels (every decoder width) are held in registers between the two passes: 4 x 16 bytes per lane
  uint4 held[4];
  const bool fits = nvec <= 128;
  float ss = 0.f;
  if (fits) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = lane + 32 * k;
      held[k] = i < nvec ? ldg_nc_v4(xr + i) : make_uint4(0, 0, 0, 0);
      float f[8];
      unpack8(held[k], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
    }
  } else {
    for (int i = lane; i < nvec; i += 32) {
      float f[8];
      unpack8(xr[i], f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
    }
  }
  ss = warp_sum(ss);
  const float inv = sqrt_c / fmaxf(sqrtf(ss), 1e-12f);
  uint4* orow = reinterpret_cast<uint4*>(out + row * Cp);
  const uint4* gr = reinterpret_cast<const uint4*>(gamma);
  auto finish = [&](int i, const uint4& xv) {
    float f[8], g[8];
    unpack8(xv, f);
    unpack8(ldg_nc_v4(gr + i), g);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y = round_bf16(f[j] * inv * g[j]);
      if (silu) y = y / (1.f + __expf(-y));
      f[j] = y;
    }
    orow[i] = pack8(f);
  };
  if (fits) {
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (lane + 32 * k < nvec) finish(lane + 32 * k, held[k]);
  } else {
    for (int i = lane; i < nvec; i += 32) finish(i, xr[i]);
  }
}

// ---------------------------------------------------------------------------------------------
// Upsample(scale 2, nearest-exact) per frame (VAE:73-79, 98-100) into the interior of a grid of twice the size. With halves = 2
// the source is the output of the temporal up-sampling conv (VAE:147-156): frame t' of the result is the channel half t' % 2 of
// source frame t' / 2.
// ---------------------------------------------------------------------------------------------
__global__ void vae_upsample2x_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int Cp, int T_dst, int H,
                                      int W, int halves) {
  const int vec = Cp / 8;
  const int64_t total = static_cast<int64_t>(T_dst) * (2 * H) * (2 * W) * vec;
  const int64_t ld_src = static_cast<int64_t>(halves) * Cp;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % vec);
    int64_t r = idx / vec;
    const int x = static_cast<int>(r % (2 * W));
    r /= 2 * W;
    const int y = static_cast<int>(r % (2 * H)), t = static_cast<int>(r / (2 * H));
    const int64_t srow = (static_cast<int64_t>(t / halves) * (H + 2) + (y >> 1) + 1) * (W + 2) + (x >> 1) + 1;
    const uint4 v = ldg_nc_v4(reinterpret_cast<const uint4*>(src + srow * ld_src + static_cast<int64_t>(t % halves) * Cp) + c);
    const int64_t drow = (static_cast<int64_t>(t) * (2 * H + 2) + y + 1) * (2 * W + 2) + x + 1;
    reinterpret_cast<uint4*>(dst + drow * Cp)[c] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// main += DupUp3D(x) (VAE:417-439, 511-512): channel j = ((c_out*ft + a)*2 + b)*2 + d of repeat_interleave(x, repeats) lands at
// (t*ft + a, 2y + b, 2x + d); on the first chunk the first ft-1 frames are dropped.
// ---------------------------------------------------------------------------------------------
__global__ void vae_dup_up_add_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ main, int Cin_p, int Cout, int Cout_p,
                                      int repeats, int ft, int skip, int T_out, int H, int W) {
  // one thread = 8 consecutive output channels of one position: one 16-byte read-modify-write of `main`, 8 gathered x values
  const int groups = (Cout + 7) / 8;
  const int64_t total = static_cast<int64_t>(T_out) * (2 * H) * (2 * W) * groups;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c0 = static_cast<int>(idx % groups) * 8;
    int64_t r = idx / groups;
    const int xo = static_cast<int>(r % (2 * W));
    r /= 2 * W;
    const int yo = static_cast<int>(r % (2 * H)), to = static_cast<int>(r / (2 * H));
    const int tt = to + skip, a = tt % ft, t = tt / ft;
    const int sub = (a * 2 + (yo & 1)) * 2 + (xo & 1);          // j = c * (4 ft) + sub
    const int64_t srow = (static_cast<int64_t>(t) * (H + 2) + (yo >> 1) + 1) * (W + 2) + (xo >> 1) + 1;
    const int64_t drow = (static_cast<int64_t>(to) * (2 * H + 2) + yo + 1) * (2 * W + 2) + xo + 1;
    const __nv_bfloat16* xs = x + srow * Cin_p;
    uint4* d = reinterpret_cast<uint4*>(main + drow * Cout_p + c0);
    float f[8];
    unpack8(*d, f);
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (c0 + i < Cout) f[i] = round_bf16(f[i] + __bfloat162float(xs[((c0 + i) * 4 * ft + sub) / repeats]));
    *d = pack8(f);
  }
}

// ---------------------------------------------------------------------------------------------
// Row softmax of the attention block's scores (VAE:331-337, scale 1/sqrt(C)), in place on bf16 [rows][ld]; keys are grid
// positions of one frame: the border positions and the columns past the frame are excluded.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) vae_attn_softmax_kernel(__nv_bfloat16* __restrict__ s, int64_t ld, int rows, int n_cols, int gh, int gw,
                                                               float scale) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  __nv_bfloat16* sr = s + static_cast<int64_t>(row) * ld;
  const int tokens = gh * gw;
  auto live = [&](int col) {
    if (col >= tokens) return false;
    const int y = col / gw, x = col - y * gw;
    return y > 0 && y < gh - 1 && x > 0 && x < gw - 1;
  };
  float mx = -INFINITY;
  for (int c = lane; c < n_cols; c += 32)
    if (live(c)) mx = fmaxf(mx, __bfloat162float(sr[c]) * scale);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
  float sum = 0.f;
  for (int c = lane; c < n_cols; c += 32)
    if (live(c)) sum += __expf(__bfloat162float(sr[c]) * scale - mx);
  sum = warp_sum(sum);
  const float inv = sum > 0.f ? 1.f / sum : 0.f;
  for (int c = lane; c < n_cols; c += 32)
    sr[c] = __float2bfloat16_rn(live(c) ? __expf(__bfloat162float(sr[c]) * scale - mx) * inv : 0.f);
}

// ---------------------------------------------------------------------------------------------
// Un-patchify 'b (c r q) f h w -> b c f (h q) (w r)' (VAE:214-224) of the head output (12 of Cp channels) into the video, either
// clamped to [-1, 1] directly (single_decode, VAE:1212-1215) or accumulated with the blending ramps of tiled_decode
// (VAE:1081-1100, 1128-1150): values += v * mask, weight += mask, mask = min(ramp_y, ramp_x).
// ---------------------------------------------------------------------------------------------
struct VaeOutParams {
  int T, H, W, Cp;            // head grid: T frames of H x W positions (interior), Cp channels per row
  int t0, y0, x0;             // where the tile starts in the video (frames, pixels)
  int VT, VH, VW;             // video size
  int top_bound, bottom_bound, left_bound, right_bound, border_y, border_x;
};

__device__ __forceinline__ float vae_ramp(int i, int n, int first_is_bound, int last_is_bound, int border) {
  float m = 1.f;
  if (!first_is_bound && i < border) m = static_cast<float>(i + 1) / border;
  if (!last_is_bound && i >= n - border) m = static_cast<float>(n - i) / border;   // flip((arange+1)/border)
  return m;
}

template <bool BLEND>
__global__ void vae_unpatchify_kernel(const __nv_bfloat16* __restrict__ head, float* __restrict__ values, float* __restrict__ weight,
                                      VaeOutParams p) {
  const int PH = 2 * p.H, PW = 2 * p.W;
  const int64_t total = static_cast<int64_t>(p.T) * PH * PW;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(idx % PW);
    int64_t r = idx / PW;
    const int Y = static_cast<int>(r % PH), t = static_cast<int>(r / PH);
    const int q = Y & 1, rr = X & 1;
    const __nv_bfloat16* hrow = head + ((static_cast<int64_t>(t) * (p.H + 2) + (Y >> 1) + 1) * (p.W + 2) + (X >> 1) + 1) * p.Cp;
    const int vy = p.y0 + Y, vx = p.x0 + X, vt = p.t0 + t;
    if (vy >= p.VH || vx >= p.VW || vt >= p.VT) continue;
    float m = 1.f;
    if (BLEND) m = fminf(vae_ramp(Y, PH, p.top_bound, p.bottom_bound, p.border_y), vae_ramp(X, PW, p.left_bound, p.right_bound, p.border_x));
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = __bfloat162float(hrow[(c * 2 + rr) * 2 + q]);
      const int64_t o = ((static_cast<int64_t>(c) * p.VT + vt) * p.VH + vy) * p.VW + vx;
      if (BLEND) values[o] += v * m;
      else values[o] = fminf(fmaxf(v, -1.f), 1.f);
    }
    if (BLEND) weight[(static_cast<int64_t>(vt) * p.VH + vy) * p.VW + vx] += m;
  }
}

__global__ void vae_blend_finish_kernel(float* __restrict__ values, const float* __restrict__ weight, int64_t plane, int channels) {
  const int64_t total = plane * channels;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x)
    values[idx] = fminf(fmaxf(values[idx] / weight[idx % plane], -1.f), 1.f);
}

// =============================================================================================
// Encoder side (VideoVAE38_.encode, VAE:1298-1323; Encoder3d_38, VAE:620-733).
// =============================================================================================

// patchify 'b c f (h q) (w r) -> b (c r q) f h w' (VAE:199-211): video bf16 [3][T][H][W] -> grid [T][H/2+2][W/2+2][Cp], channel
// (c*2 + r)*2 + q holds pixel (2y + q, 2x + r) of colour c.
__global__ void vae_patchify_rows_kernel(const __nv_bfloat16* __restrict__ video, __nv_bfloat16* __restrict__ out, int T, int H, int W,
                                         int Cp) {
  const int64_t total = static_cast<int64_t>(T) * H * W * 3;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(idx % W);
    int64_t r = idx / W;
    const int Y = static_cast<int>(r % H);
    r /= H;
    const int t = static_cast<int>(r % T), c = static_cast<int>(r / T);
    const int ch = (c * 2 + (X & 1)) * 2 + (Y & 1);
    const int64_t row = (static_cast<int64_t>(t) * (H / 2 + 2) + (Y >> 1) + 1) * (W / 2 + 2) + (X >> 1) + 1;
    out[row * Cp + ch] = video[idx];
  }
}

// space-to-depth by 2 in h, w: dst[t][y+1][x+1][(py*2 + px)*Cp + c] = src[t][2y+py+1][2x+px+1][c]. A stride-2 3x3 convolution
// with ZeroPad2d((0,1,0,1)) (Resample 'downsample2d/3d', VAE:106-117, 240-249) is then a stride-1 convolution with 2x2 taps
// over 4*Cp channels of this grid (kernel rows / columns 2t + p, the fourth one zero), i.e. one fgb_conv_taps_bf16 launch.
__global__ void vae_space_to_depth_kernel(const __nv_bfloat16* __restrict__ src, __nv_bfloat16* __restrict__ dst, int Cp, int T, int H, int W) {
  const int vec = Cp / 8;
  const int h2 = H / 2, w2 = W / 2;
  const int64_t total = static_cast<int64_t>(T) * H * W * vec;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(idx % vec);
    int64_t r = idx / vec;
    const int X = static_cast<int>(r % W);
    r /= W;
    const int Y = static_cast<int>(r % H), t = static_cast<int>(r / H);
    const uint4 v = ldg_nc_v4(reinterpret_cast<const uint4*>(src + ((static_cast<int64_t>(t) * (H + 2) + Y + 1) * (W + 2) + X + 1) * Cp) + c);
    const int64_t drow = (static_cast<int64_t>(t) * (h2 + 2) + (Y >> 1) + 1) * (w2 + 2) + (X >> 1) + 1;
    reinterpret_cast<uint4*>(dst + drow * (4ll * Cp) + ((Y & 1) * 2 + (X & 1)) * Cp)[c] = v;
  }
}

// main += AvgDown3D(x) (VAE:363-395, 469-474): the (ft, fs, fs) sub-positions of x fold into channels k = ((c*ft + a)*fs + b)*fs + d,
// consecutive groups of `group` such channels are averaged into one output channel. pad_front zero frames precede x when its
// frame count is not a multiple of ft (the single-frame first chunk).
__global__ void vae_avg_down_add_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ main, int Cin_p, int Cout, int Cout_p,
                                        int group, int ft, int fs, int pad_front, int T_out, int H_out, int W_out) {
  const int64_t total = static_cast<int64_t>(T_out) * H_out * W_out * Cout;
  const int Hin = H_out * fs, Win = W_out * fs;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int co = static_cast<int>(idx % Cout);
    int64_t r = idx / Cout;
    const int xo = static_cast<int>(r % W_out);
    r /= W_out;
    const int yo = static_cast<int>(r % H_out), to = static_cast<int>(r / H_out);
    float sum = 0.f;
    for (int g = 0; g < group; ++g) {
      const int k = co * group + g;
      const int c = k / (ft * fs * fs), rem = k % (ft * fs * fs);
      const int a = rem / (fs * fs), b = (rem / fs) % fs, d = rem % fs;
      const int ti = to * ft + a - pad_front;
      if (ti < 0) continue;
      sum += __bfloat162float(x[((static_cast<int64_t>(ti) * (Hin + 2) + yo * fs + b + 1) * (Win + 2) + xo * fs + d + 1) * Cin_p + c]);
    }
    __nv_bfloat16* dptr = main + ((static_cast<int64_t>(to) * (H_out + 2) + yo + 1) * (W_out + 2) + xo + 1) * Cout_p + co;
    *dptr = __float2bfloat16_rn(__bfloat162float(*dptr) + round_bf16(sum / group));
  }
}

// Latent mean out of the 1x1x1 `conv1` grid (first z of its 2z channels), normalised (mu - mean) * inv_std (VAE:1313-1320), into
// fp32 [z][VT][Vh][Vw] at (t0, y0, x0): directly, or blended like tiled_encode (VAE:1181-1203).
template <bool BLEND>
__global__ void vae_latent_out_kernel(const __nv_bfloat16* __restrict__ grid, const float* __restrict__ mean, const float* __restrict__ inv_std,
                                      float* __restrict__ values, float* __restrict__ weight, int Z, VaeOutParams p) {
  const int64_t total = static_cast<int64_t>(p.T) * p.H * p.W;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int X = static_cast<int>(idx % p.W);
    int64_t r = idx / p.W;
    const int Y = static_cast<int>(r % p.H), t = static_cast<int>(r / p.H);
    const int vy = p.y0 + Y, vx = p.x0 + X, vt = p.t0 + t;
    if (vy >= p.VH || vx >= p.VW || vt >= p.VT) continue;
    const __nv_bfloat16* row = grid + ((static_cast<int64_t>(t) * (p.H + 2) + Y + 1) * (p.W + 2) + X + 1) * p.Cp;
    float m = 1.f;
    if (BLEND) m = fminf(vae_ramp(Y, p.H, p.top_bound, p.bottom_bound, p.border_y), vae_ramp(X, p.W, p.left_bound, p.right_bound, p.border_x));
    for (int c = 0; c < Z; ++c) {
      const float v = (__bfloat162float(row[c]) - mean[c]) * inv_std[c];
      const int64_t o = ((static_cast<int64_t>(c) * p.VT + vt) * p.VH + vy) * p.VW + vx;
      if (BLEND) values[o] += round_bf16(v) * m;
      else values[o] = v;
    }
    if (BLEND) weight[(static_cast<int64_t>(vt) * p.VH + vy) * p.VW + vx] += m;
  }
}

__global__ void vae_blend_divide_kernel(float* __restrict__ values, const float* __restrict__ weight, int64_t plane, int channels) {
  const int64_t total = plane * channels;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x)
    values[idx] = values[idx] / weight[idx % plane];
}

}  // namespace fgb

using namespace fgb;
using bf16 = __nv_bfloat16;

extern "C" int fgb_vae_latent_rows(fgb_ctx* ctx, const void* z, const void* mean_f32, const void* inv_std_f32, void* grid, int32_t channels,
                                   int32_t frames, int32_t h, int32_t w, int32_t cp, void* stream) {
  FGB_CHECK_ARG(ctx && z && mean_f32 && inv_std_f32 && grid, "fgb_vae_latent_rows: NULL argument");
  FGB_CHECK_ARG(channels > 0 && frames > 0 && h > 0 && w > 0 && cp >= channels && cp % 8 == 0, "fgb_vae_latent_rows: C=%d T=%d H=%d W=%d Cp=%d",
                channels, frames, h, w, cp);
  const int64_t total = static_cast<int64_t>(frames) * h * w * channels;
  vae_latent_rows_kernel<<<vae_grid(total, 256, ctx->sm_count * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(z), static_cast<const float*>(mean_f32), static_cast<const float*>(inv_std_f32), static_cast<bf16*>(grid),
      channels, frames, h, w, cp);
  FGB_LAUNCH_CHECK("vae_latent_rows_kernel");
  return FGB_OK;
}

extern "C" int fgb_vae_norm_silu(fgb_ctx* ctx, const void* x, void* out, int64_t rows, int32_t channels, int32_t cp, const void* gamma,
                                 int32_t silu, void* stream) {
  FGB_CHECK_ARG(ctx && x && out && gamma, "fgb_vae_norm_silu: NULL argument");
  FGB_CHECK_ARG(rows > 0 && channels > 0 && cp >= channels && cp % 8 == 0 && aligned16(x) && aligned16(out) && aligned16(gamma),
                "fgb_vae_norm_silu: rows=%lld C=%d Cp=%d (Cp %% 8, 16-byte alignment)", (long long)rows, channels, cp);
  vae_norm_silu_kernel<<<static_cast<unsigned>((rows + 3) / 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(out), rows, cp, sqrtf(static_cast<float>(channels)), static_cast<const bf16*>(gamma), silu);
  FGB_LAUNCH_CHECK("vae_norm_silu_kernel");
  return FGB_OK;
}

extern "C" int fgb_vae_upsample2x(fgb_ctx* ctx, const void* src, void* dst, int32_t cp, int32_t frames_dst, int32_t h, int32_t w,
                                  int32_t halves, void* stream) {
  FGB_CHECK_ARG(ctx && src && dst, "fgb_vae_upsample2x: NULL argument");
  FGB_CHECK_ARG(cp > 0 && cp % 8 == 0 && frames_dst > 0 && h > 0 && w > 0 && (halves == 1 || halves == 2) && frames_dst % halves == 0 &&
                    aligned16(src) && aligned16(dst), "fgb_vae_upsample2x: Cp=%d T=%d H=%d W=%d halves=%d", cp, frames_dst, h, w, halves);
  const int64_t total = static_cast<int64_t>(frames_dst) * 4 * h * w * (cp / 8);
  vae_upsample2x_kernel<<<vae_grid(total, 256, ctx->sm_count * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(src), static_cast<bf16*>(dst), cp, frames_dst, h, w, halves);
  FGB_LAUNCH_CHECK("vae_upsample2x_kernel");
  return FGB_OK;
}

extern "C" int fgb_vae_dup_up_add(fgb_ctx* ctx, const void* x, void* main, int32_t cin, int32_t cin_p, int32_t cout, int32_t cout_p,
                                  int32_t factor_t, int32_t first_chunk, int32_t frames_out, int32_t h, int32_t w, void* stream) {
  FGB_CHECK_ARG(ctx && x && main, "fgb_vae_dup_up_add: NULL argument");
  FGB_CHECK_ARG(cin > 0 && cout > 0 && cin_p >= cin && cout_p >= cout && (factor_t == 1 || factor_t == 2) && frames_out > 0 && h > 0 && w > 0 &&
                    (cout * factor_t * 4) % cin == 0, "fgb_vae_dup_up_add: Cin=%d Cout=%d factor_t=%d", cin, cout, factor_t);
  FGB_CHECK_ARG(cout_p % 8 == 0 && aligned16(main), "fgb_vae_dup_up_add: main rows must be 16-byte aligned");
  const int64_t total = static_cast<int64_t>(frames_out) * 4 * h * w * ((cout + 7) / 8);
  vae_dup_up_add_kernel<<<vae_grid(total, 256, ctx->sm_count * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(main), cin_p, cout, cout_p, cout * factor_t * 4 / cin, factor_t,
      first_chunk ? factor_t - 1 : 0, frames_out, h, w);
  FGB_LAUNCH_CHECK("vae_dup_up_add_kernel");
  return FGB_OK;
}

extern "C" int fgb_vae_attn_softmax(fgb_ctx* ctx, void* scores, int64_t ld, int32_t rows, int32_t n_cols, int32_t grid_h, int32_t grid_w,
                                    float scale, void* stream) {
  FGB_CHECK_ARG(ctx && scores && rows > 0 && n_cols > 0 && ld >= n_cols && grid_h >= 3 && grid_w >= 3, "fgb_vae_attn_softmax: bad argument");
  vae_attn_softmax_kernel<<<(rows + 3) / 4, 128, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<bf16*>(scores), ld, rows, n_cols, grid_h,
                                                                                        grid_w, scale);
  FGB_LAUNCH_CHECK("vae_attn_softmax_kernel");
  return FGB_OK;
}

extern "C" int fgb_vae_unpatchify(fgb_ctx* ctx, const void* head, int32_t frames, int32_t h, int32_t w, int32_t cp, void* values_f32,
                                  void* weight_f32, int32_t t0, int32_t y0, int32_t x0, int32_t video_frames, int32_t video_h,
                                  int32_t video_w, int32_t bounds_tblr, int32_t border_y, int32_t border_x, void* stream) {
  FGB_CHECK_ARG(ctx && head && values_f32, "fgb_vae_unpatchify: NULL argument");
  FGB_CHECK_ARG(frames > 0 && h > 0 && w > 0 && cp >= 12 && t0 >= 0 && y0 >= 0 && x0 >= 0 && video_frames > 0 && video_h > 0 && video_w > 0,
                "fgb_vae_unpatchify: bad geometry");
  FGB_CHECK_ARG(!weight_f32 || (border_y > 0 && border_x > 0), "fgb_vae_unpatchify: blending needs positive border widths");
  VaeOutParams p{frames, h, w, cp, t0, y0, x0, video_frames, video_h, video_w, (bounds_tblr >> 3) & 1, (bounds_tblr >> 2) & 1,
                 (bounds_tblr >> 1) & 1, bounds_tblr & 1, border_y, border_x};
  const int64_t total = static_cast<int64_t>(frames) * 4 * h * w;
  const int grid = vae_grid(total, 256, ctx->sm_count * 16);
  if (weight_f32)
    vae_unpatchify_kernel<true><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(head), static_cast<float*>(values_f32),
                                                                                     static_cast<float*>(weight_f32), p);
  else
    vae_unpatchify_kernel<false><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(head),
                                                                                      static_cast<float*>(values_f32), nullptr, p);
  FGB_LAUNCH_CHECK("vae_unpatchify_kernel");
  return FGB_OK;
}

extern "C" int fgb_vae_blend_finish(fgb_ctx* ctx, void* values_f32, const void* weight_f32, int64_t plane, int32_t channels, void* stream) {
  FGB_CHECK_ARG(ctx && values_f32 && weight_f32 && plane > 0 && channels > 0, "fgb_vae_blend_finish: bad argument");
  vae_blend_finish_kernel<<<vae_grid(plane * channels, 256, ctx->sm_count * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<float*>(values_f32), static_cast<const float*>(weight_f32), plane, channels);
  FGB_LAUNCH_CHECK("vae_blend_finish_kernel");
  return FGB_OK;
}

extern "C" int fgb_vae_patchify_rows(fgb_ctx* ctx, const void* video, void* grid, int32_t frames, int32_t h, int32_t w, int32_t cp,
                                     void* stream) {
  FGB_CHECK_ARG(ctx && video && grid, "fgb_vae_patchify_rows: NULL argument");
  FGB_CHECK_ARG(frames > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0 && cp >= 12, "fgb_vae_patchify_rows: T=%d H=%d W=%d Cp=%d (even H, W)",
                frames, h, w, cp);
  const int64_t total = static_cast<int64_t>(frames) * h * w * 3;
  vae_patchify_rows_kernel<<<vae_grid(total, 256, ctx->sm_count * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(video), static_cast<bf16*>(grid), frames, h, w, cp);
  FGB_LAUNCH_CHECK("vae_patchify_rows_kernel");
  return FGB_OK;
}

extern "C" int fgb_vae_space_to_depth(fgb_ctx* ctx, const void* src, void* dst, int32_t cp, int32_t frames, int32_t h, int32_t w, void* stream) {
  FGB_CHECK_ARG(ctx && src && dst, "fgb_vae_space_to_depth: NULL argument");
  FGB_CHECK_ARG(cp > 0 && cp % 8 == 0 && frames > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0 && aligned16(src) && aligned16(dst),
                "fgb_vae_space_to_depth: Cp=%d T=%d H=%d W=%d (even H, W)", cp, frames, h, w);
  const int64_t total = static_cast<int64_t>(frames) * h * w * (cp / 8);
  vae_space_to_depth_kernel<<<vae_grid(total, 256, ctx->sm_count * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(src), static_cast<bf16*>(dst), cp, frames, h, w);
  FGB_LAUNCH_CHECK("vae_space_to_depth_kernel");
  return FGB_OK;
}

extern "C" int fgb_vae_avg_down_add(fgb_ctx* ctx, const void* x, void* main, int32_t cin, int32_t cin_p, int32_t cout, int32_t cout_p,
                                    int32_t factor_t, int32_t factor_s, int32_t pad_front, int32_t frames_out, int32_t h_out, int32_t w_out,
                                    void* stream) {
  FGB_CHECK_ARG(ctx && x && main, "fgb_vae_avg_down_add: NULL argument");
  FGB_CHECK_ARG(cin > 0 && cout > 0 && cin_p >= cin && cout_p >= cout && (factor_t == 1 || factor_t == 2) && (factor_s == 1 || factor_s == 2) &&
                    pad_front >= 0 && pad_front < factor_t && frames_out > 0 && h_out > 0 && w_out > 0 &&
                    (cin * factor_t * factor_s * factor_s) % cout == 0,
                "fgb_vae_avg_down_add: Cin=%d Cout=%d factor_t=%d factor_s=%d pad_front=%d", cin, cout, factor_t, factor_s, pad_front);
  const int64_t total = static_cast<int64_t>(frames_out) * h_out * w_out * cout;
  vae_avg_down_add_kernel<<<vae_grid(total, 256, ctx->sm_count * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), static_cast<bf16*>(main), cin_p, cout, cout_p, cin * factor_t * factor_s * factor_s / cout, factor_t, factor_s,
      pad_front, frames_out, h_out, w_out);
  FGB_LAUNCH_CHECK("vae_avg_down_add_kernel");
  return FGB_OK;
}

extern "C" int fgb_vae_latent_out(fgb_ctx* ctx, const void* grid, int32_t frames, int32_t h, int32_t w, int32_t cp, const void* mean_f32,
                                  const void* inv_std_f32, int32_t z_dim, void* values_f32, void* weight_f32, int32_t t0, int32_t y0, int32_t x0,
                                  int32_t out_frames, int32_t out_h, int32_t out_w, int32_t bounds_tblr, int32_t border_y, int32_t border_x,
                                  void* stream) {
  FGB_CHECK_ARG(ctx && grid && mean_f32 && inv_std_f32 && values_f32, "fgb_vae_latent_out: NULL argument");
  FGB_CHECK_ARG(frames > 0 && h > 0 && w > 0 && z_dim > 0 && cp >= z_dim && t0 >= 0 && y0 >= 0 && x0 >= 0 && out_frames > 0 && out_h > 0 && out_w > 0,
                "fgb_vae_latent_out: bad geometry");
  FGB_CHECK_ARG(!weight_f32 || (border_y > 0 && border_x > 0), "fgb_vae_latent_out: blending needs positive border widths");
  VaeOutParams p{frames, h, w, cp, t0, y0, x0, out_frames, out_h, out_w, (bounds_tblr >> 3) & 1, (bounds_tblr >> 2) & 1,
                 (bounds_tblr >> 1) & 1, bounds_tblr & 1, border_y, border_x};
  const int64_t total = static_cast<int64_t>(frames) * h * w;
  const int blocks = vae_grid(total, 128, ctx->sm_count * 16);
  if (weight_f32)
    vae_latent_out_kernel<true><<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(grid), static_cast<const float*>(mean_f32), static_cast<const float*>(inv_std_f32),
        static_cast<float*>(values_f32), static_cast<float*>(weight_f32), z_dim, p);
  else
    vae_latent_out_kernel<false><<<blocks, 128, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const bf16*>(grid), static_cast<const float*>(mean_f32), static_cast<const float*>(inv_std_f32),
        static_cast<float*>(values_f32), nullptr, z_dim, p);
  FGB_LAUNCH_CHECK("vae_latent_out_kernel");
  return FGB_OK;
}

extern "C" int fgb_vae_blend_divide(fgb_ctx* ctx, void* values_f32, const void* weight_f32, int64_t plane, int32_t channels, void* stream) {
  FGB_CHECK_ARG(ctx && values_f32 && weight_f32 && plane > 0 && channels > 0, "fgb_vae_blend_divide: bad argument");
  vae_blend_divide_kernel<<<vae_grid(plane * channels, 256, ctx->sm_count * 16), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<float*>(values_f32), static_cast<const float*>(weight_f32), plane, channels);
  FGB_LAUNCH_CHECK("vae_blend_divide_kernel");
  return FGB_OK;
}
