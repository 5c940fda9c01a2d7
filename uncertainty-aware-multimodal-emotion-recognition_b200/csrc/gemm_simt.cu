// fp32 CUDA-core GEMM: exact-fp32 engine for small / unaligned shapes (K=84, K=10, N=4, N=1 ...) and the
// reference engine the tcgen05 path is validated against on the device.
//   C[b] = act(opA(A[b]) opB(B[b]) + bias[b] + beta*C[b])
// 64x64 block tile, BK=16, 256 threads, 4x4 register micro-tile.  Split-K (atomic accumulation) is used when
// the output has too few tiles to fill 148 SMs and K is long (weight gradients: K = B*T).
#include "common.cuh"

namespace deer {

constexpr int BM = 64, BN = 64, BK = 16;
int g_small_engine = 2;  // 1: register-staged 32x32 kernel, 2: cp.async-pipelined one (deer_set_option DEER_OPT_SMALL_GEMM)

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, long long lda,
                                                        const float* __restrict__ B, long long ldb,
                                                        float* __restrict__ C, long long ldc, int M, int N, int K,
                                                        const float* __restrict__ bias, int act, float beta,
                                                        int splitk, long long sA, long long sB, long long sC,
                                                        long long sBias) {
  DEER_PDL_ENTRY();
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int z = blockIdx.z;
  const int batch = z / splitk, ks = z % splitk;
  A += batch * sA;
  B += batch * sB;
  C += batch * sC;
  if (bias) bias += batch * sBias;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16x16 threads, each 4x4 outputs
  // K range of this split (multiples of BK)
  const int ktiles = (K + BK - 1) / BK;
  const int per = (ktiles + splitk - 1) / splitk;
  const int kbeg = ks * per * BK;
  const int kend = min(K, (ks + 1) * per * BK);

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
    // ---- load A tile (BM x BK) into As[k][m]
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int m, k;
      if (!TA) {  // A[m,k], k contiguous
        k = tid & 15;
        m = (tid >> 4) + 16 * j;
      } else {  // A stored [K,M], m contiguous
        m = tid & 63;
        k = (tid >> 6) + 4 * j;
      }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < M && gk < kend) v = TA ? A[(long long)gk * lda + gm] : A[(long long)gm * lda + gk];
      As[k][m] = v;
    }
    // ---- load B tile (BK x BN) into Bs[k][n]
#pragma unroll
    for (int j = 0; j < 4; j++) {
      int n, k;
      if (TB) {  // B stored [N,K], k contiguous
        k = tid & 15;
        n = (tid >> 4) + 16 * j;
      } else {  // B stored [K,N], n contiguous
        n = tid & 63;
        k = (tid >> 6) + 4 * j;
      }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < N && gk < kend) v = TB ? B[(long long)gn * ldb + gk] : B[(long long)gk * ldb + gn];
      Bs[k][n] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; k++) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; i++) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; j++) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  // ---- epilogue
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int gm = m0 + ty * 4 + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const int gn = n0 + tx * 4 + j;
      if (gn >= N) continue;
      float* c = C + (long long)gm * ldc + gn;
      if (splitk > 1) {
        // accumulate semantics only (beta==1, no act); bias added by split 0
        float v = acc[i][j];
        if (bias && ks == 0) v += bias[gn];
        atomicAdd(c, v);
      } else {
        float v = acc[i][j];
        if (bias) v += bias[gn];
        if (beta != 0.f) v += beta * (*c);
        *c = act_apply(v, act);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Small-problem variant: 32x32 output tile per CTA, FOUR 64-thread groups that each reduce a quarter of K (intra-CTA
// split-K, partials summed through shared memory), register prefetch of the next k-tile.  The ~100 post-pooling
// GEMMs of a step (M = batch rows, N,K <= 768) have too few 64x64 tiles to occupy 148 SMs and a serial K loop that
// is pure load->sync->FMA latency; this shape gives 4x more CTAs and a 4x shorter dependent chain, still exact fp32
// and with the full bias / beta / activation epilogue (no atomics).
constexpr int SM_T = 32, SM_K = 16, SM_G = 4, SM_LD = SM_T + 4;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_small_kernel(const float* __restrict__ A, long long lda,
                                                         const float* __restrict__ B, long long ldb,
                                                         float* __restrict__ C, long long ldc, int M, int N, int K,
                                                         const float* __restrict__ bias, int act, float beta,
                                                         long long sA, long long sB, long long sC, long long sBias) {
  DEER_PDL_ENTRY();
  __shared__ __align__(16) float sm[2 * SM_G * SM_K * SM_LD];  // operand tiles, later the 3 x 1024 partial sums
  float(*As)[SM_K][SM_LD] = reinterpret_cast<float(*)[SM_K][SM_LD]>(sm);
  float(*Bs)[SM_K][SM_LD] = reinterpret_cast<float(*)[SM_K][SM_LD]>(sm + SM_G * SM_K * SM_LD);
  const int batch = blockIdx.z;
  A += batch * sA;
  B += batch * sB;
  C += batch * sC;
  if (bias) bias += batch * sBias;
  const int m0 = blockIdx.y * SM_T, n0 = blockIdx.x * SM_T;
  const int g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int tx = t & 7, ty = t >> 3;  // 8x8 threads, 4x4 outputs each
  // K range of this group, in whole k-tiles
  const int ktiles = (K + SM_K - 1) / SM_K;
  const int per = (ktiles + SM_G - 1) / SM_G;
  const int kbeg = g * per * SM_K;
  const int kend = min(K, (g + 1) * per * SM_K);

  // element (row r of the 32-wide tile dim, k) handled by this thread in load slot j (8 slots per operand)
  float ra[8], rb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int m, k;
      if (!TA) { k = t & 15; m = (t >> 4) + 4 * j; } else { m = t & 31; k = (t >> 5) + 2 * j; }
      const int gm = m0 + m, gk = k0 + k;
      ra[j] = (gm < M && gk < kend) ? (TA ? __ldg(A + (long long)gk * lda + gm) : __ldg(A + (long long)gm * lda + gk)) : 0.f;
      int n, kk;
      if (TB) { kk = t & 15; n = (t >> 4) + 4 * j; } else { n = t & 31; kk = (t >> 5) + 2 * j; }
      const int gn = n0 + n, gk2 = k0 + kk;
      rb[j] = (gn < N && gk2 < kend) ? (TB ? __ldg(B + (long long)gn * ldb + gk2) : __ldg(B + (long long)gk2 * ldb + gn)) : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int j = 0; j < 8; j++) {
      int m, k;
      if (!TA) { k = t & 15; m = (t >> 4) + 4 * j; } else { m = t & 31; k = (t >> 5) + 2 * j; }
      As[g][k][m] = ra[j];
      int n, kk;
      if (TB) { kk = t & 15; n = (t >> 4) + 4 * j; } else { n = t & 31; kk = (t >> 5) + 2 * j; }
      Bs[g][kk][n] = rb[j];
    }
  };
  auto group_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(g + 1) : "memory"); };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  if (kbeg < kend) {
    fetch(kbeg);
    for (int k0 = kbeg; k0 < kend; k0 += SM_K) {
      stash();
      group_sync();
      if (k0 + SM_K < kend) fetch(k0 + SM_K);  // in flight while this tile is multiplied
#pragma unroll
      for (int k = 0; k < SM_K; k++) {
        const float4 a = *reinterpret_cast<const float4*>(&As[g][k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bs[g][k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      }
      group_sync();
    }
  }
  // ---- cross-group reduction through shared memory (reuses the operand tiles: 4608 floats >= 3 x 1024 partials)
  __syncthreads();
  float* red = sm;
  if (g > 0) {
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < 4; j++) red[(g - 1) * 1024 + (ty * 4 + i) * 32 + tx * 4 + j] = acc[i][j];
  }
  __syncthreads();
  if (g == 0) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int gm = m0 + ty * 4 + i;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const int gn = n0 + tx * 4 + j;
        const int o = (ty * 4 + i) * 32 + tx * 4 + j;
        float v = acc[i][j] + red[o] + red[1024 + o] + red[2048 + o];
        if (gm < M && gn < N) {
          float* c = C + (long long)gm * ldc + gn;
          if (bias) v += bias[gn];
          if (beta != 0.f) v += beta * (*c);
          *c = act_apply(v, act);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Small-problem engine, second generation: same decomposition (32x32 output tile, four 64-thread groups that each own
// a quarter of K, partial sums combined through shared memory) but the operands travel global -> shared with
// cp.async (16 B, zero-filled tails) through a 3-deep ring of 32-wide k-stages per group, so a group has up to
// 24 KB of loads in flight instead of one register-staged k-tile: the ~100 post-pooling GEMMs of a step were pure
// load -> barrier -> FMA latency.  Operand tiles keep their global orientation (k-contiguous rows, or k-major panels)
// and the inner loop reads 4x4 register blocks with 128-bit shared loads either way.  Exact fp32 FMA accumulation.
constexpr int S2_T = 32;            // output tile edge
constexpr int S2_K = 32;            // k per stage
constexpr int S2_G = 4;             // k-groups per CTA
constexpr int S2_S = 3;             // stages per group
constexpr int S2_LD = S2_K + 4;     // padded row (16-byte aligned, conflict-free 128-bit reads)
constexpr int S2_TILE = S2_T * S2_LD;                                   // floats per operand tile
constexpr int S2_SMEM = S2_G * S2_S * 2 * S2_TILE * (int)sizeof(float);  // 110,592 B: two CTAs per SM

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, int src_bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// AK / BK: operand rows are k-contiguous in global memory (A stored [M,K] / B stored [N,K]); otherwise the operand is
// stored k-major ([K,M] / [K,N]).
template <bool AK, bool BK_>
__global__ void __launch_bounds__(256, 2) gemm_small2_kernel(const float* __restrict__ A, long long lda,
                                                              const float* __restrict__ B, long long ldb,
                                                              float* __restrict__ C, long long ldc, int M, int N, int K,
                                                              const float* __restrict__ bias, int act, float beta,
                                                              long long sA, long long sB, long long sC,
                                                              long long sBias) {
  DEER_PDL_ENTRY();
  extern __shared__ __align__(16) float s2[];
  const int batch = blockIdx.z;
  A += batch * sA;
  B += batch * sB;
  C += batch * sC;
  if (bias) bias += batch * sBias;
  const int m0 = blockIdx.y * S2_T, n0 = blockIdx.x * S2_T;
  const int g = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int tx = t & 7, ty = t >> 3;
  float* gA = s2 + (size_t)g * S2_S * 2 * S2_TILE;  // [stage][A tile | B tile]
  // K range of this group in whole stages
  const int kst = (K + S2_K - 1) / S2_K;
  const int per = (kst + S2_G - 1) / S2_G;
  const int kbeg = min(K, g * per * S2_K);
  const int kend = min(K, (g + 1) * per * S2_K);
  const int nst = (kend - kbeg + S2_K - 1) / S2_K;

  auto issue = [&](int st) {  // stage st of this group -> ring slot st % S2_S
    float* dA = gA + (size_t)(st % S2_S) * 2 * S2_TILE;
    float* dB = dA + S2_TILE;
    const int k0 = kbeg + st * S2_K;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const int f = t + 64 * q, r = f >> 3, c4 = (f & 7) * 4;
      {
        // AK: r = tile row (m), c4 = k offset; else: r = k offset, c4 = tile row offset
        const int gm = m0 + (AK ? r : c4), gk = k0 + (AK ? c4 : r);
        const int lim = AK ? (gm < M ? kend - gk : 0) : (gk < kend ? M - gm : 0);
        const int bytes = max(0, min(4, lim)) * 4;
        const float* src = bytes ? (AK ? A + (long long)gm * lda + gk : A + (long long)gk * lda + gm) : A;
        cp_async16(dA + r * S2_LD + c4, src, bytes);
      }
      {
        const int gn = n0 + (BK_ ? r : c4), gk = k0 + (BK_ ? c4 : r);
        const int lim = BK_ ? (gn < N ? kend - gk : 0) : (gk < kend ? N - gn : 0);
        const int bytes = max(0, min(4, lim)) * 4;
        const float* src = bytes ? (BK_ ? B + (long long)gn * ldb + gk : B + (long long)gk * ldb + gn) : B;
        cp_async16(dB + r * S2_LD + c4, src, bytes);
      }
    }
  };
  auto group_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(g + 1) : "memory"); };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;

  // output element (i, j) of this thread: row ty*4 + i; column tx + 8*j when B rows are k-contiguous (the eight lanes
  // of a quarter warp then read eight different 16-byte bank groups), tx*4 + j when B is k-major (one 128-bit read)
#pragma unroll
  for (int st = 0; st < S2_S - 1; st++) {
    if (st < nst) issue(st);
    cp_async_commit();
  }
  for (int st = 0; st < nst; st++) {
    cp_async_wait<S2_S - 2>();  // this thread's copies of stage st have landed
    group_sync();               // ... and everyone else's; slot (st-1) % S is free again
    if (st + S2_S - 1 < nst) issue(st + S2_S - 1);
    cp_async_commit();
    const float* tA = gA + (size_t)(st % S2_S) * 2 * S2_TILE;
    const float* tB = tA + S2_TILE;
#pragma unroll
    for (int kc = 0; kc < S2_K; kc += 4) {
      float a[4][4], b[4][4];  // [i or j][kk]
#pragma unroll
      for (int x = 0; x < 4; x++) {
        if (AK) {
          const float4 v = *reinterpret_cast<const float4*>(tA + (ty * 4 + x) * S2_LD + kc);
          a[x][0] = v.x; a[x][1] = v.y; a[x][2] = v.z; a[x][3] = v.w;
        } else {
          const float4 v = *reinterpret_cast<const float4*>(tA + (kc + x) * S2_LD + ty * 4);
          a[0][x] = v.x; a[1][x] = v.y; a[2][x] = v.z; a[3][x] = v.w;
        }
        if (BK_) {
          const float4 v = *reinterpret_cast<const float4*>(tB + (tx + 8 * x) * S2_LD + kc);
          b[x][0] = v.x; b[x][1] = v.y; b[x][2] = v.z; b[x][3] = v.w;
        } else {
          const float4 v = *reinterpret_cast<const float4*>(tB + (kc + x) * S2_LD + tx * 4);
          b[0][x] = v.x; b[1][x] = v.y; b[2][x] = v.z; b[3][x] = v.w;
        }
      }
#pragma unroll
      for (int kk = 0; kk < 4; kk++)
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
          for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a[i][kk], b[j][kk], acc[i][j]);
    }
  }
  cp_async_wait<0>();
  // ---- combine the four k-groups through shared memory; all 256 threads finish four outputs each (coalesced rows)
  __syncthreads();
  float* red = s2;  // [4][32][33]
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) red[(g * 32 + ty * 4 + i) * 33 + (BK_ ? tx + 8 * j : tx * 4 + j)] = acc[i][j];
  __syncthreads();
#pragma unroll
  for (int q = 0; q < 4; q++) {
    const int o = threadIdx.x + 256 * q, r = o >> 5, c = o & 31;
    const int gm = m0 + r, gn = n0 + c;
    if (gm < M && gn < N) {
      float v = red[r * 33 + c] + red[(32 + r) * 33 + c] + red[(64 + r) * 33 + c] + red[(96 + r) * 33 + c];
      float* cp = C + (long long)gm * ldc + gn;
      if (bias) v += bias[gn];
      if (beta != 0.f) v += beta * (*cp);
      *cp = act ? act_apply1(v, act) : v;
    }
  }
}

static bool small2_ok(const float* A, long long lda, const float* B, long long ldb, long long sA, long long sB) {
  return ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0 && lda % 4 == 0 && ldb % 4 == 0 &&
         sA % 4 == 0 && sB % 4 == 0;
}

int gemm_simt(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
              long long ldc, int M, int N, int K, const float* bias, int act, float beta, int batch, long long sA,
              long long sB, long long sC, long long sBias, cudaStream_t stream) {
  // small problems (fewer 64x64 tiles than SMs): 32x32 tiles with intra-CTA split-K
  if ((long long)((N + BN - 1) / BN) * ((M + BM - 1) / BM) * batch < kNumSMs && K <= 4096) {
    dim3 sgrid((N + SM_T - 1) / SM_T, (M + SM_T - 1) / SM_T, batch);
    if (g_small_engine == 2 && K >= 64 && small2_ok(A, lda, B, ldb, sA, sB)) {
      static bool attr_done = false;
      if (!attr_done) {
        cudaFuncSetAttribute(gemm_small2_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM);
        cudaFuncSetAttribute(gemm_small2_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM);
        cudaFuncSetAttribute(gemm_small2_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM);
        cudaFuncSetAttribute(gemm_small2_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, S2_SMEM);
        attr_done = true;
      }
      // template flags: operand rows k-contiguous?  A stored [M,K] <=> !transA; B stored [N,K] <=> transB
#define GO2(AKF, BKF)                                                                                              \
  DEER_LAUNCH((gemm_small2_kernel<AKF, BKF>), sgrid, 256, S2_SMEM, stream, A, lda, B, ldb, C, ldc, M, N, K, bias, act, \
              beta, sA, sB, sC, sBias)
      if (!transA && transB) GO2(true, true);
      else if (!transA && !transB) GO2(true, false);
      else if (transA && transB) GO2(false, true);
      else GO2(false, false);
#undef GO2
      return DEER_OK;
    }
#define GOS(TA, TB)                                                                                              \
  DEER_LAUNCH((gemm_small_kernel<TA, TB>), sgrid, 256, 0, stream, A, lda, B, ldb, C, ldc, M, N, K, bias, act, beta, \
              sA, sB, sC, sBias)
    if (!transA && !transB) GOS(false, false);
    else if (!transA && transB) GOS(false, true);
    else if (transA && !transB) GOS(true, false);
    else GOS(true, true);
#undef GOS
    return DEER_OK;
  }
  const int gx = (N + BN - 1) / BN, gy = (M + BM - 1) / BM;
  int splitk = 1;
  if (beta == 1.f && act == DEER_ACT_NONE) {
    const long long tiles = (long long)gx * gy * batch;
    if (tiles < 2 * kNumSMs && K >= 2048) {
      splitk = (int)((4LL * kNumSMs + tiles - 1) / tiles);
      const int maxsplit = K / 256;
      if (splitk > maxsplit) splitk = maxsplit;
      if (splitk < 1) splitk = 1;
    }
  }
  dim3 grid(gx, gy, batch * splitk);
#define GO(TA, TB)                                                                                                 \
  DEER_LAUNCH((gemm_simt_kernel<TA, TB>), grid, 256, 0, stream, A, lda, B, ldb, C, ldc, M, N, K, bias, act, beta, \
              splitk, sA, sB, sC, sBias)
  if (!transA && !transB) GO(false, false);
  else if (!transA && transB) GO(false, true);
  else if (transA && !transB) GO(true, false);
  else GO(true, true);
#undef GO
  return DEER_OK;
}

}  // namespace deer
