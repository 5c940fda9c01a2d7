// Error-compensated 3xTF32 tensor-core GEMM with fused prologue / epilogue: the engine of the post-pooling chain
// (fusion Q/K/V + MLP layers, fusion.py:188-343; NIG head layers, deer.py:30-108,198-266; pooled model,
// complete_project.py:61-588) -- M = batch rows, K and N <= 1536: ~25 dependent GEMMs forward and ~50 backward per
// step.  Precision: every fp32 operand x is split in registers into hi = tf32(x), lo = tf32(x - hi) and each product is
// accumulated as a_lo*b_hi + a_hi*b_lo + a_hi*b_hi in the fp32 accumulators of mma.sync.m16n8k8.tf32 (the dropped
// a_lo*b_lo term is 2^-22 relative): fp32-grade results at every M, so the precision policy of the chain no longer
// depends on the batch size.
//
//   C[b] = dropout( act( opA(A_eff[b]) opB(B[b]) + bias[b] + beta*C[b] ) )        A_eff = A (.) gatefn(gate)
//
// Fusions (each one removes an elementwise kernel + an HBM/L2 round trip from the serial chain):
//   * epilogue: bias, activation, inverted dropout (same Philox stream as deer_dropout over the flat output index);
//   * A-operand prologue ("gate"): the backward of Linear -> ReLU -> Dropout needs dz = dy (.) [d > 0]/(1-p), where d is
//     the layer's saved (post-dropout) output -- applied while the fragments are read, dz is never materialised
//     (tanh / sigmoid derivative from the saved output likewise);
//   * bias gradient: with transA (dW = dz^T x) the column sums of A_eff over the contraction are accumulated by the
//     first column of CTAs into `colsum` (db), so a layer's backward is exactly two launches (dx; dW + db).
//
// Tiling: small problems (the chain at B = 256) use 32x32 output tiles, 8 warps = 2 (M) x 4 k-groups that each own a
// 32-wide slice of every 128-wide k-stage (the serial k-loop is 4x shorter; partial sums combined through shared
// memory); large M uses 64x64 tiles, 4 warps, no intra-CTA split.  Operands travel global -> shared by cp.async through
// a 3-stage ring in their global orientation (k-contiguous rows or k-major panels), padded so fragment reads are
// bank-conflict-free.  Cross-CTA split-K (atomic accumulation) for weight gradients with a long contraction.
#include "gemm_x3.cuh"

namespace deer {
namespace x3 {

template <int BM, int BN, int WM, int WN, int WK, bool AK, bool BKC, bool GATE, bool VEC>
static cudaError_t launch_one(const Args& a, dim3 grid, cudaStream_t stream) {
  using C_ = Cfg<BM, BN, WM, WN, WK>;
  auto kern = gemm_x3_kernel<BM, BN, WM, WN, WK, AK, BKC, GATE, VEC>;
  constexpr int smem = (GATE ? C_::SMEM1 : C_::SMEM0) * (int)sizeof(float);
  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr_done = true;
  }
  return launch_kernel(kern, grid, dim3(C_::NT), (size_t)smem, stream, a);
}

template <int BM, int BN, int WM, int WN, int WK>
static cudaError_t launch_cfg(const Args& a, dim3 grid, bool vec, cudaStream_t stream) {
  const bool ak = !a.transA, bk = a.transB != 0, gate = a.gate != nullptr;
#define X3_GO(AKF, BKF, GF, VF) return launch_one<BM, BN, WM, WN, WK, AKF, BKF, GF, VF>(a, grid, stream)
#define X3_V(AKF, BKF, GF) \
  do {                     \
    if (vec) X3_GO(AKF, BKF, GF, true); \
    else X3_GO(AKF, BKF, GF, false);    \
  } while (0)
#define X3_G(AKF, BKF)        \
  do {                        \
    if (gate) X3_V(AKF, BKF, true); \
    else X3_V(AKF, BKF, false);     \
  } while (0)
  if (ak && bk) X3_G(true, true);
  else if (ak && !bk) X3_G(true, false);
  else if (!ak && bk) X3_G(false, true);
  else X3_G(false, false);
#undef X3_G
#undef X3_V
#undef X3_GO
  return cudaSuccess;
}

}  // namespace x3

int gemm_x3(const x3::Args& in, cudaStream_t stream) {
  x3::Args a = in;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = al16(a.A) && al16(a.B) && a.lda % 4 == 0 && a.ldb % 4 == 0 && a.sA % 4 == 0 && a.sB % 4 == 0 &&
                   (!a.gate || (al16(a.gate) && a.ldgate % 4 == 0 && a.sGate % 4 == 0));
  const long long t32 = cdiv(a.M, 32) * cdiv(a.N, 32) * a.batch;
  const long long t64 = cdiv(a.M, 64) * cdiv(a.N, 64) * a.batch;
  const bool small = t64 < 2 * kNumSMs;        // not enough 64x64 tiles for two waves: 32x32 tiles + intra-CTA split-K
  const long long tiles = small ? t32 : t64;
  const int ks_stage = small ? 128 : 32;
  a.splitk = 1;
  if (a.beta == 1.f && a.act == DEER_ACT_NONE && a.drop_p == 0.f && tiles < kNumSMs && a.K >= 8 * ks_stage) {
    int s = (int)((2LL * kNumSMs + tiles - 1) / tiles);
    const int maxs = a.K / (4 * ks_stage);
    if (s > maxs) s = maxs;
    if (s > 1) a.splitk = s;
  }
  cudaError_t e;
  if (small) {
    dim3 grid((unsigned)cdiv(a.N, 32), (unsigned)cdiv(a.M, 32), (unsigned)(a.batch * a.splitk));
    e = x3::launch_cfg<32, 32, 2, 1, 4>(a, grid, vec, stream);
  } else {
    dim3 grid((unsigned)cdiv(a.N, 64), (unsigned)cdiv(a.M, 64), (unsigned)(a.batch * a.splitk));
    e = x3::launch_cfg<64, 64, 2, 2, 1>(a, grid, vec, stream);
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  g_engine_calls[DEER_ENGINE_TF32X3]++;
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_status(e, "gemm_x3_kernel");
  return DEER_OK;
}

// plain-GEMM entry used by the deer_gemm dispatcher (engine DEER_GEMM_TF32X3)
int gemm_x3_plain(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                  long long ldc, int M, int N, int K, const float* bias, int act, float beta, int batch, long long sA,
                  long long sB, long long sC, long long sBias, cudaStream_t stream) {
  x3::Args a = {};
  a.A = A; a.B = B; a.C = C; a.bias = bias;
  a.lda = lda; a.ldb = ldb; a.ldc = ldc;
  a.sA = sA; a.sB = sB; a.sC = sC; a.sBias = sBias;
  a.M = M; a.N = N; a.K = K; a.batch = batch;
  a.transA = transA; a.transB = transB; a.act = act; a.beta = beta;
  a.gate_scale = 1.f;
  return gemm_x3(a, stream);
}

}  // namespace deer

using namespace deer;

extern "C" {

int deer_gemm_x3(const deer_gemm_x3_args* u, void* stream) {
  DEER_CHECK_ARG(u && u->A && u->B && u->C, "gemm_x3: null pointer");
  DEER_CHECK_ARG(u->M > 0 && u->N > 0 && u->K > 0 && u->batch > 0, "gemm_x3: empty shape");
  DEER_CHECK_ARG(u->lda >= (u->transA ? u->M : u->K) && u->ldb >= (u->transB ? u->K : u->N) && u->ldc >= u->N,
                 "gemm_x3: leading dimension too small");
  DEER_CHECK_ARG(u->act >= 0 && u->act <= 3, "gemm_x3: bad activation");
  DEER_CHECK_ARG(u->beta == 0.f || u->beta == 1.f, "gemm_x3: beta must be 0 or 1");
  DEER_CHECK_ARG(u->drop_p >= 0.f && u->drop_p < 1.f, "gemm_x3: dropout probability must be in [0,1)");
  DEER_CHECK_ARG(u->gate == nullptr || (u->gate_mode >= 1 && u->gate_mode <= 3 && u->ldgate >= (u->transA ? u->M : u->K)),
                 "gemm_x3: bad gate operand");
  DEER_CHECK_ARG(u->colsum == nullptr || u->transA, "gemm_x3: colsum (bias gradient) needs transA (dW = dz^T x)");
  x3::Args a = {};
  a.A = u->A; a.B = u->B; a.C = u->C; a.bias = u->bias; a.gate = u->gate; a.colsum = u->colsum;
  a.drop_step = u->drop_step;
  a.lda = u->lda; a.ldb = u->ldb; a.ldc = u->ldc; a.ldgate = u->ldgate;
  a.sA = u->sA; a.sB = u->sB; a.sC = u->sC; a.sBias = u->sBias; a.sGate = u->sGate; a.sColsum = u->sColsum;
  a.drop_seed = u->drop_seed; a.drop_offset = u->drop_offset;
  a.drop_ld = u->drop_ld > 0 ? u->drop_ld : u->ldc;
  a.drop_col0 = u->drop_col0;
  a.drop_batch_stride = u->drop_batch_stride;
  a.M = u->M; a.N = u->N; a.K = u->K; a.batch = u->batch;
  a.transA = u->transA; a.transB = u->transB; a.act = u->act;
  a.beta = u->beta; a.drop_p = u->drop_p;
  a.gate_mode = u->gate ? u->gate_mode : 0;
  a.gate_scale = u->gate ? u->gate_scale : 1.f;
  return gemm_x3(a, (cudaStream_t)stream);
}

}  // extern "C"
