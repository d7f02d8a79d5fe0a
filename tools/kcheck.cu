// kcheck — standalone bring-up harness for the tensor-core kernels behind the C ABI (no torch).
// Compares fgb_gemm_bf16 / fgb_attn_fwd with naive fp32 CUDA-core kernels on the same bf16 inputs
// and times them with CUDA events. Test infrastructure only; not part of the product path.
//
//   kcheck gemm M N K EPI [iters]
//   kcheck attn S_Q S_KV HEADS [iters] [check=1]
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../include/fairygen_b200.h"

#define CK(x)                                                                            \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess) {                                                             \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);    \
      exit(2);                                                                           \
    }                                                                                    \
  } while (0)
#define FK(x)                                                         \
  do {                                                                \
    int r_ = (x);                                                     \
    if (r_) {                                                         \
      printf("fgb error %d: %s\n", r_, fgb_last_error());             \
      exit(3);                                                        \
    }                                                                 \
  } while (0)

typedef __nv_bfloat16 bf16;

__global__ void fill_kernel(bf16* p, int64_t n, uint32_t seed, float scale) {
  int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t h = (uint32_t)i * 2654435761u + seed * 40503u;
  h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
  float u = (h & 0xFFFFFF) / 16777216.0f;  // [0,1)
  p[i] = __float2bfloat16((u * 2.f - 1.f) * scale);
}

__device__ float gelu_ref(float x) { return 0.5f * x * (1.f + tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x))); }
__device__ float rb(float x) { return __bfloat162float(__float2bfloat16(x)); }

__global__ void gemm_ref_kernel(const bf16* A, const bf16* W, const bf16* bias, const bf16* Cin, float* Cout, int M,
                                int N, int K, int epi, const bf16* g0, const bf16* g1, int rows_g0) {
  int n = blockIdx.x * blockDim.x + threadIdx.x;
  int m = blockIdx.y;
  if (n >= N || m >= M) return;
  float acc = 0.f;
  for (int k = 0; k < K; ++k) acc += __bfloat162float(A[(int64_t)m * K + k]) * __bfloat162float(W[(int64_t)n * K + k]);
  float y = rb(acc + (bias ? __bfloat162float(bias[n]) : 0.f));
  float x = Cin ? __bfloat162float(Cin[(int64_t)m * N + n]) : 0.f;
  float out;
  if (epi == 0) out = y;
  else if (epi == 1) out = gelu_ref(y);
  else if (epi == 2) out = x + rb(__bfloat162float((m < rows_g0 ? g0 : g1)[n]) * y);
  else out = x + y;
  Cout[(int64_t)m * N + n] = rb(out);
}

// one warp per (query row, head): exact two-pass softmax in fp32
__global__ void attn_ref_kernel(const bf16* q, const bf16* k, const bf16* v, float* o, int s_q, int s_kv, int heads,
                                float scale, float* lse_ref) {
  int row = blockIdx.x, h = blockIdx.y, lane = threadIdx.x;
  int64_t W = (int64_t)heads * 128;
  float qv[4];
  for (int i = 0; i < 4; ++i) qv[i] = __bfloat162float(q[row * W + h * 128 + lane * 4 + i]);
  extern __shared__ float sc[];
  float mx = -INFINITY;
  for (int j = 0; j < s_kv; ++j) {
    float d = 0.f;
    for (int i = 0; i < 4; ++i) d += qv[i] * __bfloat162float(k[j * W + h * 128 + lane * 4 + i]);
    for (int o2 = 16; o2; o2 >>= 1) d += __shfl_xor_sync(~0u, d, o2);
    d *= scale;
    if (lane == 0) sc[j] = d;
    mx = fmaxf(mx, d);
  }
  __syncwarp();
  float l = 0.f, acc[4] = {0, 0, 0, 0};
  for (int j = 0; j < s_kv; ++j) {
    float p = expf(sc[j] - mx);
    l += p;
    float pb = rb(p);
    for (int i = 0; i < 4; ++i) acc[i] += pb * __bfloat162float(v[j * W + h * 128 + lane * 4 + i]);
  }
  for (int i = 0; i < 4; ++i) o[row * W + h * 128 + lane * 4 + i] = acc[i] / l;
  if (lane == 0) lse_ref[(int64_t)h * ((s_q + 63) / 64 * 64) + row] = (mx + logf(l)) * 1.4426950408889634f;
}

static void compare(const std::vector<bf16>& got, const std::vector<float>& ref, const char* what) {
  double num = 0, den = 0, maxabs = 0;
  int64_t bad = -1;
  for (size_t i = 0; i < ref.size(); ++i) {
    double g = __bfloat162float(got[i]), r = ref[i];
    double d = fabs(g - r);
    if (!(d == d)) { d = 1e30; }
    if (d > maxabs) { maxabs = d; bad = i; }
    num += (g - r) * (g - r);
    den += r * r;
  }
  double rel = sqrt(num / (den + 1e-30));
  printf("%s: rel_l2=%.3e max_abs=%.3e (at %lld: got %f ref %f) %s\n", what, rel, maxabs, (long long)bad,
         bad >= 0 ? __bfloat162float(got[bad]) : 0.f, bad >= 0 ? ref[bad] : 0.f, rel < 5e-3 ? "PASS" : "FAIL");
}

#ifdef KCHECK_TRACE
extern "C" int fgb_debug_attn_trace(unsigned long long* host, int n);
#endif

int main(int argc, char** argv) {
  if (argc < 2) { printf("usage: kcheck gemm|attn ...\n"); return 1; }
  fgb_ctx* ctx = nullptr;
  FK(fgb_create(0, &ctx));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  if (!strcmp(argv[1], "gemm") && argc >= 6) {
    int M = atoi(argv[2]), N = atoi(argv[3]), K = atoi(argv[4]), epi = atoi(argv[5]);
    int iters = argc > 6 ? atoi(argv[6]) : 0;
    int check = argc > 7 ? atoi(argv[7]) : 1;
    bf16 *A, *W, *bias, *C, *C0, *g0, *g1;
    float* Cref;
    CK(cudaMalloc(&A, (size_t)M * K * 2));
    CK(cudaMalloc(&W, (size_t)N * K * 2));
    CK(cudaMalloc(&bias, (size_t)N * 2));
    CK(cudaMalloc(&g0, (size_t)N * 2));
    CK(cudaMalloc(&g1, (size_t)N * 2));
    CK(cudaMalloc(&C, (size_t)M * N * 2));
    CK(cudaMalloc(&C0, (size_t)M * N * 2));
    CK(cudaMalloc(&Cref, (size_t)M * N * 4));
    auto fill = [&](bf16* p, int64_t n, uint32_t seed, float s) { fill_kernel<<<(n + 255) / 256, 256>>>(p, n, seed, s); };
    fill(A, (int64_t)M * K, 1, 1.0f);
    fill(W, (int64_t)N * K, 2, 1.0f / sqrtf((float)K));
    fill(bias, N, 3, 0.5f);
    fill(g0, N, 4, 1.0f);
    fill(g1, N, 5, 1.0f);
    fill(C0, (int64_t)M * N, 6, 1.0f);
    CK(cudaMemcpy(C, C0, (size_t)M * N * 2, cudaMemcpyDeviceToDevice));
    int rows_g0 = M / 3;
    FK(fgb_gemm_bf16(ctx, A, K, W, K, bias, C, N, M, N, K, epi, g0, g1, rows_g0, nullptr));
    FK(fgb_sync_check(ctx, nullptr));
    if (check) {
      dim3 g((N + 127) / 128, M);
      gemm_ref_kernel<<<g, 128>>>(A, W, bias, C0, Cref, M, N, K, epi, g0, g1, rows_g0);
      CK(cudaDeviceSynchronize());
      std::vector<bf16> got((size_t)M * N);
      std::vector<float> ref((size_t)M * N);
      CK(cudaMemcpy(got.data(), C, got.size() * 2, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(ref.data(), Cref, ref.size() * 4, cudaMemcpyDeviceToHost));
      char name[128];
      snprintf(name, sizeof name, "gemm M=%d N=%d K=%d epi=%d", M, N, K, epi);
      compare(got, ref, name);
    }
    if (iters > 0) {
      for (int i = 0; i < 3; ++i) FK(fgb_gemm_bf16(ctx, A, K, W, K, bias, C, N, M, N, K, epi, g0, g1, rows_g0, nullptr));
      CK(cudaEventRecord(e0));
      for (int i = 0; i < iters; ++i) FK(fgb_gemm_bf16(ctx, A, K, W, K, bias, C, N, M, N, K, epi, g0, g1, rows_g0, nullptr));
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      ms /= iters;
      printf("gemm M=%d N=%d K=%d epi=%d: %.3f ms  %.1f TFLOP/s\n", M, N, K, epi, ms, 2.0 * M * N * K / ms * 1e-9);
    }
  } else if (!strcmp(argv[1], "attn") && argc >= 5) {
    int SQ = atoi(argv[2]), SKV = atoi(argv[3]), H = atoi(argv[4]);
    int iters = argc > 5 ? atoi(argv[5]) : 0;
    int check = argc > 6 ? atoi(argv[6]) : 1;
    int64_t W = (int64_t)H * 128;
    bf16 *q, *k, *v, *o;
    float* oref;
    CK(cudaMalloc(&q, (size_t)SQ * W * 2));
    CK(cudaMalloc(&k, (size_t)SKV * W * 2));
    CK(cudaMalloc(&v, (size_t)SKV * W * 2));
    CK(cudaMalloc(&o, (size_t)SQ * W * 2));
    CK(cudaMemset(o, 0xff, (size_t)SQ * W * 2));
    auto fill = [&](bf16* p, int64_t n, uint32_t seed, float s) { fill_kernel<<<(n + 255) / 256, 256>>>(p, n, seed, s); };
    fill(q, SQ * W, 11, 2.0f);
    fill(k, SKV * W, 12, 2.0f);
    fill(v, SKV * W, 13, 1.0f);
    float scale = 1.0f / sqrtf(128.f);
    // key-split workspace (KCHECK_NOSPLIT=1 runs every unit whole) and the saved log-sum-exp rows
    int64_t ws_bytes = getenv("KCHECK_NOSPLIT") ? 0 : fgb_attn_workspace_bytes(ctx, SQ, SKV, H);
    void* ws = nullptr;
    float* lse = nullptr;
    if (ws_bytes > 0) CK(cudaMalloc(&ws, ws_bytes));
    const int64_t LDS = (SQ + 63) / 64 * 64;
    CK(cudaMalloc(&lse, (size_t)LDS * H * 4));
    printf("attn workspace %lld bytes\n", (long long)ws_bytes);
    // KCHECK_BOUNDED=1: bounded-score softmax (the engine's default): key-norm bound first, then fgb_attn_fwd_bounded
    float* kmax2 = nullptr;
    float* qmax2 = nullptr;   // KCHECK_BOUNDED=2: also the head-level query bound (fgb_attn_fwd_bounded_qk)
    if (getenv("KCHECK_BOUNDED")) {
      CK(cudaMalloc(&kmax2, H * 4));
      FK(fgb_head_norm_max(ctx, k, W, SKV, H, kmax2, nullptr));
      if (atoi(getenv("KCHECK_BOUNDED")) >= 2) {
        CK(cudaMalloc(&qmax2, H * 4));
        FK(fgb_head_norm_max(ctx, q, W, SQ, H, qmax2, nullptr));
      }
    }
    auto run = [&]() {
      if (kmax2) return fgb_attn_fwd_bounded_qk(ctx, q, W, k, W, v, W, o, W, SQ, SKV, H, scale, kmax2, qmax2, lse, LDS, ws, ws_bytes, nullptr, 0, 0, 0, nullptr);
      return fgb_attn_fwd_ex(ctx, q, W, k, W, v, W, o, W, SQ, SKV, H, scale, lse, LDS, ws, ws_bytes, nullptr);
    };
    FK(run());
    FK(fgb_sync_check(ctx, nullptr));
    if (check) {
      CK(cudaMalloc(&oref, (size_t)SQ * W * 4));
      dim3 g(SQ, H);
      float* lse_ref;
      CK(cudaMalloc(&lse_ref, (size_t)LDS * H * 4));
      CK(cudaFuncSetAttribute(attn_ref_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SKV * 4));
      attn_ref_kernel<<<g, 32, SKV * 4>>>(q, k, v, oref, SQ, SKV, H, scale, lse_ref);
      CK(cudaGetLastError());
      CK(cudaDeviceSynchronize());
      {
        std::vector<float> a((size_t)LDS * H), b((size_t)LDS * H);
        CK(cudaMemcpy(a.data(), lse, a.size() * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(b.data(), lse_ref, b.size() * 4, cudaMemcpyDeviceToHost));
        double mx = 0;
        for (int h = 0; h < H; ++h)
          for (int r = 0; r < SQ; ++r) mx = fmax(mx, fabs((double)a[h * LDS + r] - b[h * LDS + r]));
        printf("lse max_abs_err=%.3e %s\n", mx, mx < 2e-3 ? "PASS" : "FAIL");
      }
      std::vector<bf16> got((size_t)SQ * W);
      std::vector<float> ref((size_t)SQ * W);
      CK(cudaMemcpy(got.data(), o, got.size() * 2, cudaMemcpyDeviceToHost));
      CK(cudaMemcpy(ref.data(), oref, ref.size() * 4, cudaMemcpyDeviceToHost));
      char name[128];
      snprintf(name, sizeof name, "attn s_q=%d s_kv=%d heads=%d", SQ, SKV, H);
      compare(got, ref, name);
    }
    if (iters > 0) {
      for (int i = 0; i < 2; ++i) FK(run());
      CK(cudaEventRecord(e0));
      for (int i = 0; i < iters; ++i) FK(run());
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      ms /= iters;
      printf("attn s_q=%d s_kv=%d heads=%d: %.3f ms  %.1f TFLOP/s\n", SQ, SKV, H, ms,
             4.0 * SQ * SKV * (double)W / ms * 1e-9);
    }
#ifdef KCHECK_TRACE
    {   // debug library build (-DFGB_ATTN_TRACE=1): phase stamps of softmax thread 0 of CTA 0, first 16 items of the last launch
      unsigned long long tr[256];
      if (fgb_debug_attn_trace(tr, 256) == 0) {
        for (int it = 1; it < 16; ++it) {
          printf("item %2d: start +%5lld | q %5lld bound %5lld |", it, (long long)(tr[it * 16] - tr[(it - 1) * 16]),
                 (long long)(tr[it * 16 + 1] - tr[it * 16]), (long long)(tr[it * 16 + 2] - tr[it * 16 + 1]));
          for (int s = 3; s <= 10; ++s) printf(" %5lld", (long long)(tr[it * 16 + s] - tr[it * 16 + s - 1]));
          printf(" | epi %5lld %5lld %5lld\n", (long long)(tr[it * 16 + 11] - tr[it * 16 + 10]),
                 (long long)(tr[it * 16 + 12] - tr[it * 16 + 11]), (long long)(tr[it * 16 + 13] - tr[it * 16 + 12]));
        }
      }
    }
#endif
  } else if (!strcmp(argv[1], "attnbwd") && argc >= 5) {
    // timing only (parity lives in tests/test_kernels_gpu.py against torch autograd)
    int SQ = atoi(argv[2]), SKV = atoi(argv[3]), H = atoi(argv[4]);
    int iters = argc > 5 ? atoi(argv[5]) : 3;
    int64_t W = (int64_t)H * 128;
    const int64_t LDS = (SQ + 63) / 64 * 64;
    bf16 *q, *k, *v, *o, *dout, *dq, *dk, *dv;
    float *lse, *delta;
    for (bf16** p : {&q, &o, &dout, &dq}) CK(cudaMalloc(p, (size_t)SQ * W * 2));
    for (bf16** p : {&k, &v, &dk, &dv}) CK(cudaMalloc(p, (size_t)SKV * W * 2));
    CK(cudaMalloc(&lse, (size_t)LDS * H * 4));
    CK(cudaMalloc(&delta, (size_t)LDS * H * 4));
    auto fill = [&](bf16* p, int64_t n, uint32_t seed, float s) { fill_kernel<<<(n + 255) / 256, 256>>>(p, n, seed, s); };
    fill(q, SQ * W, 11, 2.0f);
    fill(k, SKV * W, 12, 2.0f);
    fill(v, SKV * W, 13, 1.0f);
    fill(dout, SQ * W, 14, 1.0f);
    float scale = 1.0f / sqrtf(128.f);
    FK(fgb_attn_fwd_ex(ctx, q, W, k, W, v, W, o, W, SQ, SKV, H, scale, lse, LDS, nullptr, 0, nullptr));
    auto run = [&]() {
      return fgb_attn_bwd(ctx, q, W, k, W, v, W, o, W, dout, W, lse, delta, LDS, dq, W, dk, W, dv, W, SQ, SKV, H, scale, nullptr);
    };
    FK(run());
    FK(fgb_sync_check(ctx, nullptr));
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) FK(run());
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ms /= iters;
    // counted FLOPs: 2.5 x forward (5 S x S x 128 products: S, dP, dV, dK, dQ); the two-kernel scheme executes 7
    printf("attnbwd s_q=%d s_kv=%d heads=%d: %.3f ms  %.1f TFLOP/s counted (2.5x fwd), %.1f executed (3.5x fwd)\n", SQ, SKV, H, ms,
           10.0 * SQ * SKV * (double)W / ms * 1e-9, 14.0 * SQ * SKV * (double)W / ms * 1e-9);
  } else {
    printf("bad arguments\n");
    return 1;
  }
  FK(fgb_destroy(ctx));
  printf("done\n");
  return 0;
}
