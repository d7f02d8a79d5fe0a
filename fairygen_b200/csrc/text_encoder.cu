// fairygen_b200 — kernels of the umT5 text encoder around the tcgen05 GEMMs (sm_100a).          SURVEY §8(f) row 2
//
// Reference: animation/diffsynth/models/wan_video_text_encoder.py ("TENC"): 24 pre-norm blocks of
//   x += o(softmax(q·kᵀ + pos_bias + key_mask)·v)      T5Attention, no 1/sqrt(d) scaling, head_dim 64      TENC:59-95
//   x += fc2(fc1(n) * gelu_tanh(gate(n)))               T5FeedForward                                        TENC:108-113
// with T5LayerNorm (RMS, no bias; TENC:33-38) and a per-layer relative-position bias (TENC:159-193).
// The Linears run on fgb_gemm_bf16 (q|k|v and gate|fc1 fused along N, residual adds in the epilogue); this file holds the
// rest. The workload is 512 tokens x 1-2 prompts: every tensor here is a few MB, so the kernels are latency- rather than
// bandwidth-sized — one pass each, 16-byte accesses where rows are contiguous, fp32 math, bf16 storage.
#include "common.cuh"
#include "host.h"

namespace fgb {

static inline int te_grid(int64_t n, int block) { return static_cast<int>((n + block - 1) / block); }

// ---------------------------------------------------------------------------------------------
// token_embedding(ids)                                                                  TENC:246
// ---------------------------------------------------------------------------------------------
__global__ void embedding_rows_kernel(const __nv_bfloat16* __restrict__ table, int64_t ld_table, int vocab,
                                      const int64_t* __restrict__ ids, int n, int dim, __nv_bfloat16* out, int64_t ldo) {
  const int vec_per_row = dim / 8;
  const int64_t total = static_cast<int64_t>(n) * vec_per_row;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(idx / vec_per_row), c = static_cast<int>(idx % vec_per_row);
    const int64_t id = ids[r];
    uint4 v = make_uint4(0, 0, 0, 0);  // ids outside the table never read memory (the host mirror rejects them first)
    if (id >= 0 && id < vocab) v = ldg_nc_v4(reinterpret_cast<const uint4*>(table + id * ld_table) + c);
    reinterpret_cast<uint4*>(out + static_cast<int64_t>(r) * ldo)[c] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// T5LayerNorm: y = w * bf16(x * rsqrt(mean(x^2) + eps))                                 TENC:33-38
// One warp per row, the row re-read from L1/L2 for the second pass (rows are <= 8 KB).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) t5_layer_norm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* out,
                                                            int64_t ldo, int rows, int dim, float eps,
                                                            const __nv_bfloat16* __restrict__ weight) {
  const int row = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<int64_t>(row) * ldx);
  const int nvec = dim / 8;
  float ss = 0.f;
  for (int i = lane; i < nvec; i += 32) {
    float f[8];
    unpack8(xr[i], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
  }
  ss = warp_sum(ss);
  const float r = rsqrtf(ss / static_cast<float>(dim) + eps);
  const uint4* wr = reinterpret_cast<const uint4*>(weight);
  uint4* orow = reinterpret_cast<uint4*>(out + static_cast<int64_t>(row) * ldo);
  for (int i = lane; i < nvec; i += 32) {
    float f[8], w[8];
    unpack8(xr[i], f);
    unpack8(ldg_nc_v4(wr + i), w);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = w[j] * round_bf16(f[j] * r);
    orow[i] = pack8(f);
  }
}

// ---------------------------------------------------------------------------------------------
// h = fc1(x) * gelu_tanh(gate(x)) from the fused GEMM output [rows, gate(F) | fc1(F)]           TENC:18-22, 109
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float te_gelu_tanh(float x) {
  const float u = 0.7978845608028654f * (x + 0.044715f * x * x * x);
  return 0.5f * x * (1.f + tanhf(u));
}

__global__ void geglu_kernel(const __nv_bfloat16* __restrict__ gf, int64_t ld, __nv_bfloat16* out, int64_t ldo, int rows, int F) {
  const int vec_per_row = F / 8;
  const int64_t total = static_cast<int64_t>(rows) * vec_per_row;
  for (int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; idx < total;
       idx += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(idx / vec_per_row), c = static_cast<int>(idx % vec_per_row);
    const uint4* row = reinterpret_cast<const uint4*>(gf + static_cast<int64_t>(r) * ld);
    float g[8], f[8];
    unpack8(ldg_nc_v4(row + c), g);
    unpack8(ldg_nc_v4(row + vec_per_row + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = f[j] * round_bf16(te_gelu_tanh(g[j]));   // GELU output is a bf16 tensor in the reference
    reinterpret_cast<uint4*>(out + static_cast<int64_t>(r) * ldo)[c] = pack8(f);
  }
}

// ---------------------------------------------------------------------------------------------
// Relative-position bias as a function of (key - query): table[h][rel + (s_q-1)] = emb[bucket(rel)][h]   TENC:159-169
// The bucket of every rel comes from the host mirror (integer table, built once per sequence length).
// ---------------------------------------------------------------------------------------------
__global__ void t5_bias_table_kernel(const __nv_bfloat16* __restrict__ emb, const int32_t* __restrict__ bucket_of_rel, int n_rel,
                                     int heads, int buckets, float* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= heads * n_rel) return;
  const int h = idx / n_rel, r = idx % n_rel;
  const int b = bucket_of_rel[r];
  out[idx] = (b >= 0 && b < buckets) ? __bfloat162float(emb[b * heads + h]) : 0.f;
}

// ---------------------------------------------------------------------------------------------
// T5 self-attention, head_dim 64: o = softmax(q·kᵀ + bias[key - query] (+ -inf on masked keys))·v       TENC:59-95
//
// CTA = (32 query rows, one head, one sample), 8 warps x 4 rows; keys in chunks of 128 staged in shared memory as bf16
// with a 34-word row stride (lane j reads row j with 8-byte loads: conflict-free per half-warp), online softmax over chunks
// in fp32. A warp works on its 4 rows together so that every K / V word fetched from shared memory feeds 4 rows: QK — each
// lane owns 4 keys of the chunk, the query rows arrive as 16-byte broadcasts; PV — probabilities go through a per-warp
// shared strip and come back 4 at a time as broadcasts, each lane owns 2 of the 64 output dims. Chunks without a live key
// (the padded tail of a prompt) are skipped before they are staged. Plain FMA pipe on purpose: 512 x 512 x 64 heads is
// ~4 GFLOP per layer against ~200 GFLOP of Linears.
// ---------------------------------------------------------------------------------------------
constexpr int kTeRows = 32;    // query rows per CTA
constexpr int kTeKeys = 128;   // keys per staged chunk
constexpr int kTeStride = 34;  // 32-bit words per staged K / V row (64 bf16 + 2 pad words; even: rows stay 8-byte aligned)
constexpr int kTeSmemBytes = (2 * kTeKeys * kTeStride + kTeRows * 64 + 8 * 4 * kTeKeys) * 4;

__global__ void __launch_bounds__(256) t5_attention_kernel(const __nv_bfloat16* __restrict__ q, int64_t ldq,
                                                           const __nv_bfloat16* __restrict__ k, int64_t ldk,
                                                           const __nv_bfloat16* __restrict__ v, int64_t ldv, __nv_bfloat16* o,
                                                           int64_t ldo, int s_q, int s_kv, const float* __restrict__ bias,
                                                           const uint8_t* __restrict__ key_mask) {
  extern __shared__ __align__(16) uint32_t te_smem[];
  uint32_t* sk = te_smem;                                                   // [128][34]
  uint32_t* sv = sk + kTeKeys * kTeStride;                                  // [128][34]
  float* sq = reinterpret_cast<float*>(sv + kTeKeys * kTeStride);           // [32][64]
  float* sp = sq + kTeRows * 64;                                            // [8 warps][4 rows][128]
  const int head = blockIdx.y, sample = blockIdx.z;
  const int row0 = blockIdx.x * kTeRows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_rel = s_q + s_kv - 1;
  const float* bias_h = bias ? bias + static_cast<int64_t>(head) * n_rel + (s_q - 1) : nullptr;
  const uint8_t* mask = key_mask ? key_mask + static_cast<int64_t>(sample) * s_kv : nullptr;
  const __nv_bfloat16* qb = q + static_cast<int64_t>(sample) * s_q * ldq + head * 64;
  const __nv_bfloat16* kb = k + static_cast<int64_t>(sample) * s_kv * ldk + head * 64;
  const __nv_bfloat16* vb = v + static_cast<int64_t>(sample) * s_kv * ldv + head * 64;

  // query tile -> fp32 shared (8 threads per row, 8 values each); rows past s_q are zeros and never stored
  {
    const int r = threadIdx.x >> 3, c = threadIdx.x & 7;
    float f[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (row0 + r < s_q) unpack8(ldg_nc_v4(reinterpret_cast<const uint4*>(qb + static_cast<int64_t>(row0 + r) * ldq) + c), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) sq[r * 64 + c * 8 + j] = f[j];
  }

  float m[4], l[4], acc0[4], acc1[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m[i] = -INFINITY;
    l[i] = 0.f;
    acc0[i] = acc1[i] = 0.f;
  }
  const int rl0 = warp * 4;                 // first of this warp's rows inside the tile
  float* spw = sp + warp * 4 * kTeKeys;

  for (int key0 = 0; key0 < s_kv; key0 += kTeKeys) {
    // barrier (previous chunk consumed; first time round: query tile written) + "does this chunk hold a live key?"
    const int probe = key0 + static_cast<int>(threadIdx.x);
    const int any_live = __syncthreads_or(threadIdx.x < kTeKeys && probe < s_kv && (mask == nullptr || mask[probe] != 0));
    if (!any_live) continue;                // block-uniform
    // stage K and V rows [key0, key0+128): 8 threads per row, one 16-byte load each, stored as 4 words at stride 34
    for (int t = threadIdx.x; t < kTeKeys * 8; t += 256) {
      const int r = t >> 3, c = t & 7;
      uint4 kv4 = make_uint4(0, 0, 0, 0), vv4 = make_uint4(0, 0, 0, 0);
      if (key0 + r < s_kv) {
        kv4 = ldg_nc_v4(reinterpret_cast<const uint4*>(kb + static_cast<int64_t>(key0 + r) * ldk) + c);
        vv4 = ldg_nc_v4(reinterpret_cast<const uint4*>(vb + static_cast<int64_t>(key0 + r) * ldv) + c);
      }
      uint2* dk = reinterpret_cast<uint2*>(sk + r * kTeStride + c * 4);
      uint2* dv = reinterpret_cast<uint2*>(sv + r * kTeStride + c * 4);
      dk[0] = make_uint2(kv4.x, kv4.y); dk[1] = make_uint2(kv4.z, kv4.w);
      dv[0] = make_uint2(vv4.x, vv4.y); dv[1] = make_uint2(vv4.z, vv4.w);
    }
    __syncthreads();

    bool live[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int key = key0 + t * 32 + lane;
      live[t] = key < s_kv && (mask == nullptr || mask[key] != 0);
    }

    // ---- scores of 4 rows x (4 keys per lane)
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int t = 0; t < 4; ++t) s[i][t] = 0.f;
#pragma unroll 4
    for (int c = 0; c < 16; ++c) {          // 4 of the 64 dims per step
      float4 qq[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) qq[i] = reinterpret_cast<const float4*>(sq + (rl0 + i) * 64)[c];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const uint2 kw = *reinterpret_cast<const uint2*>(sk + (t * 32 + lane) * kTeStride + c * 2);
        const float k0 = bf16_lo(kw.x), k1 = bf16_hi(kw.x), k2 = bf16_lo(kw.y), k3 = bf16_hi(kw.y);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          s[i][t] = fmaf(qq[i].x, k0, s[i][t]);
          s[i][t] = fmaf(qq[i].y, k1, s[i][t]);
          s[i][t] = fmaf(qq[i].z, k2, s[i][t]);
          s[i][t] = fmaf(qq[i].w, k3, s[i][t]);
        }
      }
    }

    // ---- bias, mask, online softmax; probabilities to the warp's strip
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int row = row0 + rl0 + i;
      float cmax = -INFINITY;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int key = key0 + t * 32 + lane;
        if (live[t]) {
          if (bias_h && row < s_q) s[i][t] += __ldg(bias_h + (key - row));
          cmax = fmaxf(cmax, s[i][t]);
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) cmax = fmaxf(cmax, __shfl_xor_sync(0xffffffffu, cmax, off));
      const float m_new = fmaxf(m[i], cmax);        // finite: the chunk holds a live key
      const float rescale = __expf(m[i] - m_new);   // m = -inf before the first live chunk: exp(-inf) = 0
      float psum = 0.f;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const float p = live[t] ? __expf(s[i][t] - m_new) : 0.f;
        psum += p;
        spw[i * kTeKeys + t * 32 + lane] = p;
      }
      psum = warp_sum(psum);
      l[i] = l[i] * rescale + psum;
      m[i] = m_new;
      acc0[i] *= rescale;
      acc1[i] *= rescale;
    }
    __syncwarp();

    // ---- o += P·V for the 4 rows (keys past s_kv: V staged as zeros, p = 0)
    const int kmax = min(kTeKeys, s_kv - key0);
    for (int j = 0; j < kmax; j += 4) {
      const uint32_t v0 = sv[(j + 0) * kTeStride + lane], v1 = sv[(j + 1) * kTeStride + lane];
      const uint32_t v2 = sv[(j + 2) * kTeStride + lane], v3 = sv[(j + 3) * kTeStride + lane];
      const float a0 = bf16_lo(v0), b0 = bf16_hi(v0), a1 = bf16_lo(v1), b1 = bf16_hi(v1);
      const float a2 = bf16_lo(v2), b2 = bf16_hi(v2), a3 = bf16_lo(v3), b3 = bf16_hi(v3);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float4 pp = reinterpret_cast<const float4*>(spw + i * kTeKeys)[j >> 2];
        acc0[i] = fmaf(pp.x, a0, acc0[i]); acc1[i] = fmaf(pp.x, b0, acc1[i]);
        acc0[i] = fmaf(pp.y, a1, acc0[i]); acc1[i] = fmaf(pp.y, b1, acc1[i]);
        acc0[i] = fmaf(pp.z, a2, acc0[i]); acc1[i] = fmaf(pp.z, b2, acc1[i]);
        acc0[i] = fmaf(pp.w, a3, acc0[i]); acc1[i] = fmaf(pp.w, b3, acc1[i]);
      }
    }
    __syncwarp();                            // the strip is rewritten in the next chunk
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = row0 + rl0 + i;
    if (row >= s_q) break;
    const float inv = l[i] > 0.f ? 1.f / l[i] : 0.f;
    uint32_t* orow = reinterpret_cast<uint32_t*>(o + (static_cast<int64_t>(sample) * s_q + row) * ldo + head * 64);
    orow[lane] = pack_bf16(acc0[i] * inv, acc1[i] * inv);
  }
}

}  // namespace fgb

using namespace fgb;
using bf16 = __nv_bfloat16;

extern "C" int fgb_embedding_rows(fgb_ctx* ctx, const void* table, int64_t ld_table, int32_t vocab, const void* ids, int32_t n,
                                  int32_t dim, void* out, int64_t ldo, void* stream) {
  FGB_CHECK_ARG(ctx && table && ids && out, "fgb_embedding_rows: NULL argument");
  FGB_CHECK_ARG(n > 0 && vocab > 0 && dim > 0 && dim % 8 == 0 && ld_table >= dim && ld_table % 8 == 0 && ldo >= dim && ldo % 8 == 0 &&
                    aligned16(table) && aligned16(out), "fgb_embedding_rows: n=%d vocab=%d dim=%d (dim %% 8, 16-byte alignment)", n, vocab, dim);
  int grid = te_grid(static_cast<int64_t>(n) * (dim / 8), 256);
  if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
  embedding_rows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(table), ld_table, vocab,
                                                                             static_cast<const int64_t*>(ids), n, dim,
                                                                             static_cast<bf16*>(out), ldo);
  FGB_LAUNCH_CHECK("embedding_rows_kernel");
  return FGB_OK;
}

extern "C" int fgb_t5_layer_norm(fgb_ctx* ctx, const void* x, int64_t ldx, void* out, int64_t ldo, int32_t rows, int32_t dim, float eps,
                                 const void* weight, void* stream) {
  FGB_CHECK_ARG(ctx && x && out && weight, "fgb_t5_layer_norm: NULL argument");
  FGB_CHECK_ARG(rows > 0 && dim > 0 && dim % 8 == 0 && ldx >= dim && ldo >= dim && ldx % 8 == 0 && ldo % 8 == 0 && aligned16(x) &&
                    aligned16(out) && aligned16(weight), "fgb_t5_layer_norm: rows=%d dim=%d (dim %% 8, 16-byte alignment)", rows, dim);
  t5_layer_norm_kernel<<<te_grid(rows, 4), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(x), ldx, static_cast<bf16*>(out), ldo, rows, dim, eps, static_cast<const bf16*>(weight));
  FGB_LAUNCH_CHECK("t5_layer_norm_kernel");
  return FGB_OK;
}

extern "C" int fgb_geglu(fgb_ctx* ctx, const void* gate_fc1, int64_t ld, void* out, int64_t ldo, int32_t rows, int32_t ffn_dim,
                         void* stream) {
  FGB_CHECK_ARG(ctx && gate_fc1 && out, "fgb_geglu: NULL argument");
  FGB_CHECK_ARG(rows > 0 && ffn_dim > 0 && ffn_dim % 8 == 0 && ld >= 2 * static_cast<int64_t>(ffn_dim) && ld % 8 == 0 && ldo >= ffn_dim &&
                    ldo % 8 == 0 && aligned16(gate_fc1) && aligned16(out), "fgb_geglu: rows=%d ffn_dim=%d (ffn_dim %% 8, 16-byte alignment)",
                rows, ffn_dim);
  int grid = te_grid(static_cast<int64_t>(rows) * (ffn_dim / 8), 256);
  if (grid > ctx->sm_count * 16) grid = ctx->sm_count * 16;
  geglu_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const bf16*>(gate_fc1), ld, static_cast<bf16*>(out), ldo,
                                                                    rows, ffn_dim);
  FGB_LAUNCH_CHECK("geglu_kernel");
  return FGB_OK;
}

extern "C" int fgb_t5_bias_table(fgb_ctx* ctx, const void* emb, const void* bucket_of_rel, int32_t n_rel, int32_t heads, int32_t buckets,
                                 void* out, void* stream) {
  FGB_CHECK_ARG(ctx && emb && bucket_of_rel && out, "fgb_t5_bias_table: NULL argument");
  FGB_CHECK_ARG(n_rel > 0 && heads > 0 && buckets > 0, "fgb_t5_bias_table: n_rel=%d heads=%d buckets=%d", n_rel, heads, buckets);
  t5_bias_table_kernel<<<te_grid(static_cast<int64_t>(heads) * n_rel, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(emb), static_cast<const int32_t*>(bucket_of_rel), n_rel, heads, buckets, static_cast<float*>(out));
  FGB_LAUNCH_CHECK("t5_bias_table_kernel");
  return FGB_OK;
}

extern "C" int fgb_t5_attention(fgb_ctx* ctx, const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                                int64_t ldo, int32_t batch, int32_t s_q, int32_t s_kv, int32_t heads, const void* bias,
                                const void* key_mask, void* stream) {
  FGB_CHECK_ARG(ctx && q && k && v && o, "fgb_t5_attention: NULL tensor pointer");
  FGB_CHECK_ARG(batch > 0 && s_q > 0 && s_kv > 0 && heads > 0 && heads <= 65535 && batch <= 65535,
                "fgb_t5_attention: batch=%d s_q=%d s_kv=%d heads=%d", batch, s_q, s_kv, heads);
  const int64_t width = static_cast<int64_t>(heads) * 64;
  FGB_CHECK_ARG(ldq >= width && ldk >= width && ldv >= width && ldo >= width && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 2 == 0 &&
                    aligned16(q) && aligned16(k) && aligned16(v) && (reinterpret_cast<uintptr_t>(o) & 3u) == 0,
                "fgb_t5_attention: head_dim is 64; rows must be 16-byte aligned (ld %% 8) and at least heads*64 wide");
  static unsigned long long configured = 0;   // > 48 KB of dynamic shared memory needs the opt-in, once per device
  if (first_use_on_device(configured))
    FGB_CUDA(cudaFuncSetAttribute(t5_attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTeSmemBytes));
  dim3 grid((s_q + kTeRows - 1) / kTeRows, heads, batch);
  t5_attention_kernel<<<grid, 256, kTeSmemBytes, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const bf16*>(q), ldq, static_cast<const bf16*>(k), ldk, static_cast<const bf16*>(v), ldv, static_cast<bf16*>(o), ldo, s_q,
      s_kv, static_cast<const float*>(bias), static_cast<const uint8_t*>(key_mask));
  FGB_LAUNCH_CHECK("t5_attention_kernel");
  return FGB_OK;
}
