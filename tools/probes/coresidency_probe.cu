// Which resource keeps a "guest" CTA (a small tcgen05 GEMM CTA of another stream) from sharing an SM with a resident
// "hog" CTA (the persistent LSTM recurrence CTA)?  Synthetic kernels with the same footprints, one resource varied at a
// time; every CTA records %smid and %globaltimer at its start and end, the host counts the guest CTAs that STARTED on an
// SM while a hog CTA was running there.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o coresidency_probe coresidency_probe.cu
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <vector>

struct Rec {
  unsigned long long t0, t1;
  unsigned smid, pad;
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned smid() {
  unsigned s;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(s));
  return s;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t n, bool relinquish) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(n) : "memory");
  if (relinquish) asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t a, uint32_t n) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(a), "r"(n) : "memory");
}

// NR floats kept live per thread (register pressure), tm0 + tm1 TMEM columns (0 = none), spins for `ns` nanoseconds
template <int NR, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) occupy(Rec* rec, int tm0, int tm1, int relinq, unsigned long long ns, float seed,
                                                    float* sink) {
  extern __shared__ uint8_t dsm[];
  __shared__ uint32_t slot[2];
  const unsigned long long t0 = gtime();
  if (threadIdx.x < 32) {
    if (tm0) tmem_alloc(&slot[0], tm0, tm1 == 0 && relinq);
    if (tm1) tmem_alloc(&slot[1], tm1, relinq);
  }
  __syncthreads();
  const unsigned long long t_alloc = gtime();
  float r[NR];
#pragma unroll
  for (int i = 0; i < NR; i++) r[i] = seed * (float)(i + threadIdx.x);
  while (gtime() - t_alloc < ns) {
#pragma unroll
    for (int i = 0; i < NR; i++) r[i] = r[i] * 1.0001f + r[(i + 1) % NR];
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NR; i++) s += r[i];
  if (s == 123.456f) sink[threadIdx.x] = s + dsm[threadIdx.x];
  __syncthreads();
  if (threadIdx.x < 32) {
    if (tm0) tmem_dealloc(slot[0], tm0);
    if (tm1) tmem_dealloc(slot[1], tm1);
  }
  if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) {
    unsigned w;
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(w));
    sink[64 + (threadIdx.x >> 5)] = (float)w;   // hardware warp slots of CTA 0 (host prints them)
  }
  if (threadIdx.x == 0) {
    rec[blockIdx.x].t0 = t0;
    rec[blockIdx.x].t1 = gtime();
    rec[blockIdx.x].smid = smid();
    rec[blockIdx.x].pad = (unsigned)(t_alloc - t0);
  }
}


typedef void (*Kern)(Rec*, int, int, int, unsigned long long, float, float*);
struct Variant {
  const char* name;
  Kern k;
  int threads;
};
static const Variant HOGS[] = {
    {"hog 9 warps NR150", occupy<150, 288>, 288}, {"hog 9 warps NR40 ", occupy<40, 288>, 288},
    {"hog 8 warps NR150", occupy<150, 256>, 256}, {"hog 9 warps NR100", occupy<100, 288>, 288},
    {"hog 9 warps NR85 ", occupy<85, 288>, 288},  {"hog 9 warps NR120", occupy<120, 288>, 288},
};
static const Variant GUESTS[] = {
    {"guest 6 warps NR60", occupy<60, 192>, 192},  {"guest 3 warps NR60", occupy<60, 96>, 96},
    {"guest 4 warps NR60", occupy<60, 128>, 128},  {"guest 4 warps NR140", occupy<140, 128>, 128},
    {"guest 1 warp NR20 ", occupy<20, 32>, 32},    {"guest 10 warps NR60", occupy<60, 320>, 320},
    {"guest 6 warps NR80", occupy<80, 192>, 192},
};

static void launch(const Variant& v, int grid, int smem_kb, int cluster, cudaStream_t st, Rec* rec, int tm0, int tm1,
                   unsigned long long ns, float* sink) {
  cudaFuncSetAttribute(v.k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
  cudaFuncSetAttribute(v.k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(v.threads);
  cfg.dynamicSmemBytes = smem_kb * 1024;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = cluster;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, v.k, rec, tm0, tm1, 1, ns, 0.5f, sink);
  if (e != cudaSuccess) printf("launch error: %s\n", cudaGetErrorString(e));
}

int main() {
  const int NHOG = 128, NGUEST = 592;
  Rec *dh, *dg;
  float *sink_h, *sink_g;
  cudaMalloc(&dh, NHOG * sizeof(Rec));
  cudaMalloc(&dg, NGUEST * sizeof(Rec));
  cudaMalloc(&sink_h, 4096);
  cudaMalloc(&sink_g, 4096);
  int lo, hi;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  cudaStream_t s_hi, s_lo;
  cudaStreamCreateWithPriority(&s_hi, cudaStreamNonBlocking, hi);
  cudaStreamCreateWithPriority(&s_lo, cudaStreamNonBlocking, lo);
  for (auto& v : HOGS) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, v.k);
    printf("%s: %d regs\n", v.name, fa.numRegs);
  }
  for (auto& v : GUESTS) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, v.k);
    printf("%s: %d regs\n", v.name, fa.numRegs);
  }
  struct Cfg {
    int hog, guest, gsm, gtm;
  };
  const Cfg cfgs[] = {
      {0, 0, 98, 128}, {1, 0, 98, 128}, {2, 0, 98, 128}, {3, 0, 98, 128}, {4, 0, 98, 128}, {5, 0, 98, 128},
      {0, 1, 98, 128}, {0, 2, 98, 128}, {0, 3, 98, 128}, {0, 4, 8, 0},    {0, 4, 8, 32},
      {2, 1, 98, 128}, {2, 2, 98, 128}, {2, 3, 98, 128}, {2, 5, 98, 128}, {2, 6, 98, 128},
      {3, 1, 98, 128}, {3, 2, 98, 128}, {3, 3, 98, 128}, {3, 6, 98, 128},
      {4, 2, 98, 128}, {4, 3, 98, 128}, {4, 5, 98, 128}, {4, 6, 98, 128},
      {5, 1, 98, 128}, {5, 2, 98, 128}, {5, 6, 98, 128}, {1, 4, 8, 0},    {1, 0, 98, 0},
  };
  for (const Cfg& c : cfgs) {
    for (int rep = 0; rep < 2; rep++) {
      cudaMemset(dh, 0, NHOG * sizeof(Rec));
      cudaMemset(dg, 0, NGUEST * sizeof(Rec));
      cudaMemset(sink_h, 0, 4096);
      cudaMemset(sink_g, 0, 4096);
      cudaDeviceSynchronize();
      launch(HOGS[c.hog], NHOG, 120, 4, s_hi, dh, 64, 256, 400000ull, sink_h);
      launch(GUESTS[c.guest], NGUEST, c.gsm, 1, s_lo, dg, c.gtm, 0, 20000ull, sink_g);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("error %s\n", cudaGetErrorString(e));
        return 1;
      }
      if (rep == 0) continue;  // first repetition warms the attribute / module state up
      std::vector<Rec> h(NHOG), g(NGUEST);
      float wh[16], wg[16];
      cudaMemcpy(h.data(), dh, NHOG * sizeof(Rec), cudaMemcpyDeviceToHost);
      cudaMemcpy(g.data(), dg, NGUEST * sizeof(Rec), cudaMemcpyDeviceToHost);
      cudaMemcpy(wh, sink_h + 64, sizeof(wh), cudaMemcpyDeviceToHost);
      cudaMemcpy(wg, sink_g + 64, sizeof(wg), cudaMemcpyDeviceToHost);
      unsigned long long tmin = ~0ull, hog_end = 0, guest_end = 0, g_first = ~0ull;
      for (auto& r : h) tmin = std::min(tmin, r.t0), hog_end = std::max(hog_end, r.t1);
      for (auto& r : g) g_first = std::min(g_first, r.t0), guest_end = std::max(guest_end, r.t1);
      int co_started = 0, before_end = 0, stuck = 0;
      for (auto& r : g) {
        if (r.t0 < hog_end) before_end++;
        for (auto& q : h)
          if (q.smid == r.smid && r.t0 >= q.t0 && r.t0 < q.t1) {
            co_started++;
            if (r.pad > 100000u) stuck++;  // waited > 100 us in tcgen05.alloc
          }
      }
      printf("%s + %s (%3d KB, %3d cols): beside a hog %3d (stuck in alloc %3d), started before hog end %3d | first guest %6.1f, "
             "hog end %6.1f, guest end %6.1f us | hog warp slots",
             HOGS[c.hog].name, GUESTS[c.guest].name, c.gsm, c.gtm, co_started, stuck, before_end,
             ((long long)g_first - (long long)tmin) * 1e-3, (hog_end - tmin) * 1e-3, ((long long)guest_end - (long long)tmin) * 1e-3);
      for (int i = 0; i < HOGS[c.hog].threads / 32; i++) printf(" %d", (int)wh[i]);
      printf(" | guest");
      for (int i = 0; i < GUESTS[c.guest].threads / 32; i++) printf(" %d", (int)wg[i]);
      printf("\n");
    }
  }
  return 0;
}
