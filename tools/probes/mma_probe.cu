// Micro-benchmark (sm_100a): latency of a dependent chain and throughput of independent chains of
// mma.sync.m16n8k8 tf32 and m16n8k16 bf16/f16 (the legacy tensor path the 3xTF32 chain kernels use).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float* d, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float* d, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int CH, bool BF>
__global__ void probe(float* out, long long* cyc, int iters) {
  uint32_t a[4] = {0x3f800000u + threadIdx.x, 0x3f000000u, 0x3e800000u, 0x3f800000u}, b[2] = {0x3f800000u, 0x3f000000u};
  float acc[CH][4];
#pragma unroll
  for (int c = 0; c < CH; c++) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int c = 0; c < CH; c++) { if (BF) mma_bf16(acc[c], a, b); else mma_tf32(acc[c], a, b); }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < CH; c++) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH, bool BF>
void run(const char* name, int warps) {
  float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  probe<CH, BF><<<1, 32 * warps>>>(out, cyc, iters);
  probe<CH, BF><<<1, 32 * warps>>>(out, cyc, iters);
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%s warps=%d chains/warp=%d: %.1f cycles per loop iteration (%.2f cycles per MMA per warp; SM issues one MMA per %.2f cycles)\n",
         name, warps, CH, (double)h / iters, (double)h / iters / CH, (double)h / iters / CH / warps);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  run<1, false>("tf32 m16n8k8 ", 1);  run<2, false>("tf32 m16n8k8 ", 1);  run<4, false>("tf32 m16n8k8 ", 1);  run<8, false>("tf32 m16n8k8 ", 1);
  run<4, false>("tf32 m16n8k8 ", 4);  run<4, false>("tf32 m16n8k8 ", 8);  run<8, false>("tf32 m16n8k8 ", 8);  run<8, false>("tf32 m16n8k8 ", 16);
  run<1, true>("bf16 m16n8k16", 1);   run<4, true>("bf16 m16n8k16", 1);   run<8, true>("bf16 m16n8k16", 8);   run<8, true>("bf16 m16n8k16", 16);
  return 0;
}
