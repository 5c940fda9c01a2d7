// Persistent "chain" kernel: the whole post-pooling part of a step -- hierarchical fusion (fusion.py:188-343: seq-1
// cross attention, 2-token modality attention, three Linear-ReLU-Dropout-LayerNorm stages) and the multi-dimensional
// NIG head (deer.py:198-266) -- as ONE launch per direction instead of ~50 (forward) / ~70 (backward) dependent launches
// of 4-7 us each.  M = batch rows only (256 in the benchmark), so every layer is a handful of 32x32 tiles: the chain is
// pure launch + memory latency, 0.5 ms of a 4.2 ms step when run as separate kernels.
//
// The host hands the kernel a PROGRAM (kernel parameter, <= 32 KB): a list of ops grouped into dependency levels.
//   * GEMM ops run on the fused 3xTF32 tile routine of gemm_x3.cuh (bias / activation / dropout epilogue, backward gate
//     prologue, bias-gradient side output, batch strides), fp32-grade at every size;
//   * LayerNorm forward / backward and the 2-token attention core are row-wise ops;
// One CTA per SM (grid <= 148) walks the levels: the tiles of all ops of a level are dealt round-robin to the CTAs, then
// a grid-wide barrier (one atomic + spin on a counter the host zeroes) separates the levels.  All CTAs are co-resident
// by construction (grid <= SM count, one CTA per SM), so the barrier cannot deadlock against itself; kernels of other
// streams never wait on this one.
#include <string.h>

#include "gemm_x3.cuh"

namespace deer {
namespace chain {

constexpr int MAX_OPS = 72;
constexpr int MAX_LEVELS = 48;
constexpr int NT = 256;

enum Kind { K_GEMM = 0, K_LN_FWD = 1, K_LN_BWD = 2, K_MHA2_FWD = 3, K_MHA2_BWD = 4, K_AXPY = 5 };

struct Op {
  int kind, tiles, tiles_n, pad_;
  x3::Args g;   // K_GEMM: the GEMM; other kinds reuse its fields (see the op routines below)
};

struct Program {
  int nops, nlevels;
  unsigned int* barrier;            // zero-initialised counter (one per launch)
  int* error;                       // set to 1 if a barrier wait times out (device-side watchdog)
  int level_begin[MAX_LEVELS + 1];  // ops [level_begin[l], level_begin[l+1]) form level l
  Op ops[MAX_OPS];
};

using GCfg = x3::Cfg<32, 32, 2, 1, 4>;
constexpr int SMEM_FLOATS = GCfg::SMEM1;

template <bool AK, bool BKC, bool GATE>
__device__ __noinline__ void gemm_tile(const x3::Args& a, int bx, int by, int bz, float* smem) {
  x3::x3_tile<32, 32, 2, 1, 4, AK, BKC, GATE, true>(a, bx, by, bz, smem);
}

__device__ __forceinline__ void run_gemm(const Op& op, int tile, float* smem) {
  const x3::Args a = op.g;   // private copy: the tile routine reads its fields in inner loops
  const int per_batch = op.tiles_n * (int)((a.M + 31) / 32);
  const int bz = tile / per_batch, r = tile % per_batch;
  const int by = r / op.tiles_n, bx = r % op.tiles_n;
  const bool ak = !a.transA, bk = a.transB != 0, gate = a.gate != nullptr;
  if (ak) {
    if (bk) { if (gate) gemm_tile<true, true, true>(a, bx, by, bz, smem); else gemm_tile<true, true, false>(a, bx, by, bz, smem); }
    else    { if (gate) gemm_tile<true, false, true>(a, bx, by, bz, smem); else gemm_tile<true, false, false>(a, bx, by, bz, smem); }
  } else {
    if (bk) { if (gate) gemm_tile<false, true, true>(a, bx, by, bz, smem); else gemm_tile<false, true, false>(a, bx, by, bz, smem); }
    else    { if (gate) gemm_tile<false, false, true>(a, bx, by, bz, smem); else gemm_tile<false, false, false>(a, bx, by, bz, smem); }
  }
}

// ---- LayerNorm forward: g.A = x [M,N], g.B = gamma, g.bias = beta, g.C = y, g.gate = mean out, g.colsum = rstd out,
//      g.beta = eps; 8 rows per tile (one warp per row), N <= 512
__device__ __forceinline__ void run_ln_fwd(const Op& op, int tile) {
  const x3::Args& a = op.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = tile * 8 + warp, N = a.N;
  if (row >= a.M) return;
  const float* xr = a.A + (long long)row * a.lda;
  float v[16];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int j = lane + 32 * i;
    v[i] = j < N ? xr[j] : 0.f;
    s += v[i];
  }
  const float mu = warp_sum(s) / N;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int j = lane + 32 * i;
    const float d = j < N ? v[i] - mu : 0.f;
    q = fmaf(d, d, q);
  }
  const float rs = rsqrtf(warp_sum(q) / N + a.beta);
  float* yr = a.C + (long long)row * a.ldc;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int j = lane + 32 * i;
    if (j < N) yr[j] = (v[i] - mu) * rs * a.B[j] + a.bias[j];
  }
  if (lane == 0) {
    const_cast<float*>(a.gate)[row] = mu;
    a.colsum[row] = rs;
  }
}

// ---- LayerNorm backward: g.A = dy, g.B = x, g.bias = gamma, g.gate = mean, g.colsum = rstd (in), g.C = dx,
//      dgamma / dbeta accumulated (atomics) at (float*)g.drop_step and ((float*)g.drop_step) + N_pad -- carried in the
//      two spare 8-byte fields drop_seed / drop_offset; g.beta = 1: dx += ; 32 rows per tile (4 per warp)
__device__ __forceinline__ void run_ln_bwd(const Op& op, int tile, float* smem) {
  const x3::Args& a = op.g;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, N = a.N;
  float* dgamma = reinterpret_cast<float*>(a.drop_seed);
  float* dbeta = reinterpret_cast<float*>(a.drop_offset);
  float* sg = smem;            // [512] partial dgamma
  float* sb = smem + 512;      // [512] partial dbeta
  for (int j = threadIdx.x; j < 1024; j += NT) smem[j] = 0.f;
  __syncthreads();
  float ag[16], ab[16];
#pragma unroll
  for (int i = 0; i < 16; i++) ag[i] = ab[i] = 0.f;
  for (int rr = 0; rr < 4; rr++) {
    const int row = tile * 32 + warp * 4 + rr;
    if (row >= a.M) break;
    const float* dyr = a.A + (long long)row * a.lda;
    const float* xr = a.B + (long long)row * a.ldb;
    const float mu = a.gate[row], rs = a.colsum[row];
    float xh[16], gdy[16];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int j = lane + 32 * i;
      const float dy = j < N ? dyr[j] : 0.f;
      xh[i] = j < N ? (xr[j] - mu) * rs : 0.f;
      gdy[i] = j < N ? dy * a.bias[j] : 0.f;
      s1 += gdy[i];
      s2 = fmaf(gdy[i], xh[i], s2);
      ag[i] = fmaf(dy, xh[i], ag[i]);
      ab[i] += dy;
    }
    s1 = warp_sum(s1) / N;
    s2 = warp_sum(s2) / N;
    float* dxr = a.C + (long long)row * a.ldc;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int j = lane + 32 * i;
      if (j < N) {
        const float d = rs * (gdy[i] - s1 - xh[i] * s2);
        dxr[j] = a.beta != 0.f ? dxr[j] + d : d;
      }
    }
  }
  if (dgamma) {
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int j = lane + 32 * i;
      if (j < N) {
        atomicAdd(sg + j, ag[i]);
        atomicAdd(sb + j, ab[i]);
      }
    }
    __syncthreads();
    for (int j = threadIdx.x; j < N; j += NT) {
      atomicAdd(dgamma + j, sg[j]);
      atomicAdd(dbeta + j, sb[j]);
    }
  }
}

// ---- 2-token attention core (csrc/mha2.cu): g.A = qkv [B,2,3E], g.C = ctx_mean [B,E], g.colsum = attw [B,2,2] (may be
//      null), g.bias -> probs [B,heads,2,2] (written); g.N = E, g.K = heads; one sample per tile, one warp per head
__device__ __forceinline__ void run_mha2_fwd(const Op& op, int tile, float* smem) {
  const x3::Args& a = op.g;
  const int b = tile, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int E = a.N, heads = a.K, d = E / heads;
  float* pw = smem;   // [heads][4]
  if (h < heads) {
    const float scale = rsqrtf((float)d);
    const float* t0 = a.A + (long long)b * 6 * E;
    const float* t1 = t0 + 3 * E;
    const int o = h * d;
    float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
    for (int j = lane; j < d; j += 32) {
      const float q0 = t0[o + j] * scale, q1 = t1[o + j] * scale;
      const float k0 = t0[E + o + j], k1 = t1[E + o + j];
      s00 = fmaf(q0, k0, s00); s01 = fmaf(q0, k1, s01); s10 = fmaf(q1, k0, s10); s11 = fmaf(q1, k1, s11);
    }
    s00 = warp_sum(s00); s01 = warp_sum(s01); s10 = warp_sum(s10); s11 = warp_sum(s11);
    const float m0 = fmaxf(s00, s01), m1 = fmaxf(s10, s11);
    float e00 = expf(s00 - m0), e01 = expf(s01 - m0), e10 = expf(s10 - m1), e11 = expf(s11 - m1);
    const float i0 = 1.f / (e00 + e01), i1 = 1.f / (e10 + e11);
    e00 *= i0; e01 *= i0; e10 *= i1; e11 *= i1;
    for (int j = lane; j < d; j += 32) {
      const float v0 = t0[2 * E + o + j], v1 = t1[2 * E + o + j];
      a.C[(long long)b * a.ldc + o + j] = 0.5f * ((e00 + e10) * v0 + (e01 + e11) * v1);   // mean over the two tokens
    }
    if (lane == 0) {
      pw[h * 4 + 0] = e00; pw[h * 4 + 1] = e01; pw[h * 4 + 2] = e10; pw[h * 4 + 3] = e11;
      float* pr = const_cast<float*>(a.bias) + ((long long)b * heads + h) * 4;
      pr[0] = e00; pr[1] = e01; pr[2] = e10; pr[3] = e11;
    }
  }
  __syncthreads();
  if (threadIdx.x < 4 && a.colsum) {
    float s = 0.f;
    for (int k = 0; k < heads; k++) s += pw[k * 4 + threadIdx.x];
    a.colsum[(long long)b * 4 + threadIdx.x] = s / heads;
  }
}

// backward: g.A = qkv, g.B = d(ctx_mean) [B,E] (pitch ldb), g.gate = d(attw) [B,2,2] or null, g.bias = probs,
//           g.C = dqkv [B,2,3E]
__device__ __forceinline__ void run_mha2_bwd(const Op& op, int tile) {
  const x3::Args& a = op.g;
  const int b = tile, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int E = a.N, heads = a.K, d = E / heads;
  if (h >= heads) return;
  const float scale = rsqrtf((float)d);
  const float* t0 = a.A + (long long)b * 6 * E;
  const float* t1 = t0 + 3 * E;
  float* g0 = a.C + (long long)b * 6 * E;
  float* g1 = g0 + 3 * E;
  const int o = h * d;
  const float* pr = a.bias + ((long long)b * heads + h) * 4;
  const float p00 = pr[0], p01 = pr[1], p10 = pr[2], p11 = pr[3];
  const float* dcm = a.B + (long long)b * a.ldb + o;
  float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float v0 = t0[2 * E + o + j], v1 = t1[2 * E + o + j];
    const float c = 0.5f * dcm[j];   // d(ctx_0) = d(ctx_1) = d(ctx_mean) / 2
    a00 = fmaf(c, v0, a00); a01 = fmaf(c, v1, a01); a10 = fmaf(c, v0, a10); a11 = fmaf(c, v1, a11);
    g0[2 * E + o + j] = (p00 + p10) * c;
    g1[2 * E + o + j] = (p01 + p11) * c;
  }
  a00 = warp_sum(a00); a01 = warp_sum(a01); a10 = warp_sum(a10); a11 = warp_sum(a11);
  if (a.gate) {
    const float ih = 1.f / heads;
    a00 += a.gate[(long long)b * 4 + 0] * ih; a01 += a.gate[(long long)b * 4 + 1] * ih;
    a10 += a.gate[(long long)b * 4 + 2] * ih; a11 += a.gate[(long long)b * 4 + 3] * ih;
  }
  const float r0 = p00 * a00 + p01 * a01, r1 = p10 * a10 + p11 * a11;
  const float ds00 = p00 * (a00 - r0) * scale, ds01 = p01 * (a01 - r0) * scale;
  const float ds10 = p10 * (a10 - r1) * scale, ds11 = p11 * (a11 - r1) * scale;
  for (int j = lane; j < d; j += 32) {
    const float q0 = t0[o + j], q1 = t1[o + j];
    const float k0 = t0[E + o + j], k1 = t1[E + o + j];
    g0[o + j] = ds00 * k0 + ds01 * k1;
    g1[o + j] = ds10 * k0 + ds11 * k1;
    g0[E + o + j] = ds00 * q0 + ds10 * q1;
    g1[E + o + j] = ds01 * q0 + ds11 * q1;
  }
}

// ---- y (+)= x over M rows of N floats: g.A = x (pitch lda), g.C = y (pitch ldc), g.beta = 1: accumulate; 32 rows per tile
__device__ __forceinline__ void run_axpy(const Op& op, int tile) {
  const x3::Args& a = op.g;
  const int r0 = tile * 32, r1 = min(a.M, r0 + 32);
  for (int idx = threadIdx.x; idx < (r1 - r0) * a.N; idx += NT) {
    const int r = r0 + idx / a.N, c = idx % a.N;
    const float v = a.A[(long long)r * a.lda + c];
    float* y = a.C + (long long)r * a.ldc + c;
    *y = a.beta != 0.f ? *y + v : v;
  }
}

__device__ __forceinline__ bool grid_barrier(unsigned int* ctr, unsigned int target, int* error) {
  __shared__ int s_ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(ctr, 1u);
    const long long t0 = clock64();
    int ok = 1;
    while (true) {
      unsigned int v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
      if (v >= target) break;
      if (clock64() - t0 > (1ll << 31)) {   // ~1 s: a CTA never arrived (device-side watchdog, never hang the GPU)
        ok = 0;
        if (error) *error = 1;
        break;
      }
    }
    s_ok = ok;
  }
  __syncthreads();
  return s_ok != 0;
}

__global__ void __launch_bounds__(NT, 1) chain_kernel(const __grid_constant__ Program prog) {
  DEER_PDL_ENTRY();
  extern __shared__ __align__(16) float smem[];
  unsigned int arrived = 0;
  for (int L = 0; L < prog.nlevels; L++) {
    const int o0 = prog.level_begin[L], o1 = prog.level_begin[L + 1];
    int total = 0;
    for (int o = o0; o < o1; o++) total += prog.ops[o].tiles;
    for (int w = blockIdx.x; w < total; w += gridDim.x) {
      int o = o0, t = w;
      while (t >= prog.ops[o].tiles) {
        t -= prog.ops[o].tiles;
        o++;
      }
      const Op& op = prog.ops[o];
      switch (op.kind) {
        case K_GEMM: run_gemm(op, t, smem); break;
        case K_LN_FWD: run_ln_fwd(op, t); break;
        case K_LN_BWD: run_ln_bwd(op, t, smem); break;
        case K_MHA2_FWD: run_mha2_fwd(op, t, smem); break;
        case K_MHA2_BWD: run_mha2_bwd(op, t); break;
        case K_AXPY: run_axpy(op, t); break;
        default: break;
      }
      __syncthreads();   // the next tile reuses the shared-memory ring
    }
    if (L + 1 < prog.nlevels) {
      arrived += gridDim.x;
      if (!grid_barrier(prog.barrier, arrived, prog.error)) return;
    }
  }
}

}  // namespace chain
}  // namespace deer

using namespace deer;

extern "C" {

int deer_chain_max_ops(void) { return chain::MAX_OPS; }
int deer_chain_max_levels(void) { return chain::MAX_LEVELS; }

int deer_chain_run(const deer_chain_program* u, void* stream) {
  DEER_CHECK_ARG(u && u->nops > 0 && u->nops <= chain::MAX_OPS && u->nlevels > 0 && u->nlevels <= chain::MAX_LEVELS &&
                     u->barrier != nullptr,
                 "chain_run: bad program");
  static_assert(sizeof(deer_chain_op) == sizeof(chain::Op), "chain op ABI mismatch");
  static_assert(sizeof(deer_chain_program) == sizeof(chain::Program), "chain program ABI mismatch");
  chain::Program prog;
  memcpy(&prog, u, sizeof(prog));
  int max_tiles = 1;
  for (int l = 0; l < prog.nlevels; l++) {
    DEER_CHECK_ARG(prog.level_begin[l] < prog.level_begin[l + 1] && prog.level_begin[l + 1] <= prog.nops,
                   "chain_run: bad level table");
    int t = 0;
    for (int o = prog.level_begin[l]; o < prog.level_begin[l + 1]; o++) {
      chain::Op& op = prog.ops[o];
      DEER_CHECK_ARG(op.tiles > 0, "chain_run: op without tiles");
      if (op.kind == chain::K_GEMM) {
        x3::Args& a = op.g;
        auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
        DEER_CHECK_ARG(a.A && a.B && a.C && al16(a.A) && al16(a.B) && a.lda % 4 == 0 && a.ldb % 4 == 0 && a.sA % 4 == 0 &&
                           a.sB % 4 == 0 && (!a.gate || (al16(a.gate) && a.ldgate % 4 == 0 && a.sGate % 4 == 0)),
                       "chain_run: GEMM operands must be 16-byte aligned with pitches that are multiples of 4");
        a.splitk = 1;
        if (a.drop_ld <= 0) a.drop_ld = a.ldc;
        if (!a.gate) { a.gate_mode = 0; a.gate_scale = 1.f; }
        op.tiles_n = (a.N + 31) / 32;
        DEER_CHECK_ARG(op.tiles == op.tiles_n * ((a.M + 31) / 32) * a.batch, "chain_run: GEMM tile count mismatch");
      } else if (op.kind == chain::K_LN_FWD || op.kind == chain::K_LN_BWD) {
        DEER_CHECK_ARG(op.g.N <= 512, "chain_run: LayerNorm width > 512");
      } else if (op.kind == chain::K_MHA2_FWD || op.kind == chain::K_MHA2_BWD) {
        DEER_CHECK_ARG(op.g.K > 0 && op.g.K <= 8 && op.g.N % op.g.K == 0, "chain_run: attention heads");
      }
      t += op.tiles;
    }
    if (t > max_tiles) max_tiles = t;
  }
  const int grid = max_tiles < kNumSMs ? max_tiles : kNumSMs;
  constexpr int smem = chain::SMEM_FLOATS * (int)sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(chain::chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cuda_status(e, "chain_kernel smem attribute");
    attr_done = true;
  }
  DEER_LAUNCH(chain::chain_kernel, grid, chain::NT, smem, stream, prog);
  g_engine_calls[DEER_ENGINE_TF32X3]++;
  return DEER_OK;
}

}  // extern "C"
