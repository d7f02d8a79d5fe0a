// ubench — per-SMSP instruction throughput probes on sm_100a (exp2 variants, conversions, min/max) used to
// budget the attention softmax. One CTA of W warps per SM, each warp runs N independent chains.
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdint>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)
constexpr int ITERS = 4096;

template <int OP>
__global__ void k(float* out, long long* cyc) {
  float v[8];
  for (int i = 0; i < 8; ++i) v[i] = -0.001f * (threadIdx.x + i + 1);
  uint32_t h[8];
  for (int i = 0; i < 8; ++i) h[i] = 0xB800B400u + threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 2) asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(v[i]), "f"(v[(i + 1) & 7]));
      if (OP == 3) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(v[(i + 1) & 7]), "f"(v[(i + 2) & 7]));
      if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(v[i]) : "f"(v[(i + 1) & 7]), "f"(v[(i + 2) & 7]));
      if (OP == 5) { uint64_t a; asm volatile("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(v[i]), "f"(v[(i+1)&7])); asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(a)); asm volatile("mov.b64 {%0,%1}, %2;" : "=f"(v[i]), "=f"(v[(i+1)&7]) : "l"(a)); }
      if (OP == 6) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h[i]) : "f"(v[i]), "f"(v[(i + 1) & 7]));
      if (OP == 7) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(h[i]));
      if (OP == 8) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(v[i]));
      if (OP == 9) asm volatile("shl.b32 %0, %0, 23; add.s32 %0, %0, %1;" : "+r"(h[i]) : "r"(h[(i+1)&7]));
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += v[i] + __uint_as_float(h[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
  float* out; long long* cyc; CK(cudaMalloc(&out, 148 * 1024 * 4)); CK(cudaMalloc(&cyc, 8));
  const char* names[] = {"ex2.f32", "ex2.f16x2", "cvt.bf16x2.f32", "max3.f32", "fma.f32", "fma.f32x2", "cvt.f16x2.f32", "ex2.bf16x2", "tanh.f32", "shl+add"};
  for (int warps : {4, 8, 16}) {
    for (int op = 0; op < 10; ++op) {
      long long c = 0;
      switch (op) {
#define C(N) case N: k<N><<<148, warps * 32>>>(out, cyc); break;
        C(0) C(1) C(2) C(3) C(4) C(5) C(6) C(7) C(8) C(9)
      }
      CK(cudaDeviceSynchronize());
      CK(cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost));
      double per_smsp = (double)c / (ITERS * 8.0 * (warps / 4.0));  // cycles per warp-instruction per SMSP
      printf("warps/SM=%2d %-16s %.2f cycles per warp-instr per SMSP\n", warps, names[op], per_smsp);
    }
  }
  return 0;
}
