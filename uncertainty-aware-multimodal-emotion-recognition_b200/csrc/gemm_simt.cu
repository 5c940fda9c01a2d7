// fp32 CUDA-core GEMM: exact-fp32 engine for small / unaligned shapes (K=84, K=10, N=4, N=1 ...) and the
// reference engine the tcgen05 path is validated against on the device.
//   C[b] = act(opA(A[b]) opB(B[b]) + bias[b] + beta*C[b])
// 64x64 block tile, BK=16, 256 threads, 4x4 register micro-tile.  Split-K (atomic accumulation) is used when
// the output has too few tiles to fill 148 SMs and K is long (weight gradients: K = B*T).
#include "common.cuh"

namespace deer {

constexpr int BM = 64, BN = 64, BK = 16;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) gemm_simt_kernel(const float* __restrict__ A, long long lda,
                                                        const float* __restrict__ B, long long ldb,
                                                        float* __restrict__ C, long long ldc, int M, int N, int K,
                                                        const float* __restrict__ bias, int act, float beta,
                                                        int splitk, long long sA, long long sB, long long sC,
                                                        long long sBias) {
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

int gemm_simt(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
              long long ldc, int M, int N, int K, const float* bias, int act, float beta, int batch, long long sA,
              long long sB, long long sC, long long sBias, cudaStream_t stream) {
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
