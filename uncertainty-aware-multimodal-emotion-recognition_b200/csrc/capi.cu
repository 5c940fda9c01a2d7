// Library-level entry points: version, error string, launch counter, and the GEMM dispatcher.
#include <stdarg.h>

#include "common.cuh"

namespace deer {

std::atomic<long long> g_launches{0};
std::atomic<long long> g_engine_calls[8];   // per-engine dispatch counters (deer_gemm_engine_count)
int g_pdl = 0;   // measured: PDL made the B=1024 inference 10 % slower and the train step 2 % slower (early-resident
                 // dependents compete for SM slots); kept as an ablation switch
static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int gemm_simt(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
              long long ldc, int M, int N, int K, const float* bias, int act, float beta, int batch, long long sA,
              long long sB, long long sC, long long sBias, cudaStream_t stream);
// returns DEER_ERR_UNSUPPORTED (without setting an error) when the shape/alignment cannot use the TMA path
int gemm_tcgen05(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                 long long ldc, int M, int N, int K, const float* bias, int act, float beta, int batch, long long sA,
                 long long sB, long long sC, long long sBias, cudaStream_t stream);
int gemm_x3_plain(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                  long long ldc, int M, int N, int K, const float* bias, int act, float beta, int batch, long long sA,
                  long long sB, long long sC, long long sBias, cudaStream_t stream);
extern int g_small_engine;
extern int g_h16_pair;
extern int g_nig_pipe;
extern int g_lstm_dual;
extern int g_lstm_keep16;
extern int g_lstm_stasync;
extern int g_lstm_carveout;
extern int g_lstm_xin;
extern int g_lstm_xin_dual;
extern int g_lstm_halfsplit;
extern int g_lstm_colsplit;
extern int g_tf32_pair;
bool gemm_tf32_pair_supported(int transA, int transB, int M, int N, int K, long long ldc, const float* bias, int act,
                              float beta);
int gemm_tf32_pair_rowterm(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                           long long ldc, int M, int N, int K, const float* w, const float* v, int nb, int nt, int time_major,
                           cudaStream_t stream);
int gemm_tf32_pair(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                   long long ldc, int M, int N, int K, const float* bias, int act, float beta, cudaStream_t stream);
void gemm_tcgen05_set_round(int on);
void lstm_cluster_set_option(int ts, int tile);
void lstm_cluster_set_profile(long long* buf);
bool gemm_tcgen05_supported(const float* A, long long lda, int transA, const float* B, long long ldb, int transB,
                            const float* C, long long ldc, int M, int N, int K, int batch, long long sA, long long sB,
                            long long sC);

}  // namespace deer

using namespace deer;

namespace deer {
__global__ void timestamp_kernel(unsigned long long* slots, int index) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  slots[index] = t;
}
}  // namespace deer

extern "C" {

int deer_version(void) { return 200; }
long long deer_gemm_engine_count(int engine) {
  return (engine >= 0 && engine < 8) ? g_engine_calls[engine].load() : -1;
}
const char* deer_last_error(void) { return g_err; }
long long deer_launch_count(void) { return g_launches.load(); }

int deer_timestamp(unsigned long long* slots, int index, void* stream) {
  DEER_CHECK_ARG(slots && index >= 0, "timestamp: bad args");
  DEER_LAUNCH(timestamp_kernel, 1, 1, 0, stream, slots, index);
  return DEER_OK;
}

int deer_lstm_set_profile_buffer(long long* device_buf) {
  lstm_cluster_set_profile(device_buf);
  return DEER_OK;
}

int deer_set_option(int option, int value) {
  switch (option) {
    case DEER_OPT_TMA_TF32_ROUND:
      gemm_tcgen05_set_round(value);
      return DEER_OK;
    case DEER_OPT_LSTM_TS:
      lstm_cluster_set_option(value, -1);
      return DEER_OK;
    case DEER_OPT_LSTM_TILE:
      lstm_cluster_set_option(-1, value);
      return DEER_OK;
    case DEER_OPT_PDL:
      g_pdl = value ? 1 : 0;
      return DEER_OK;
    case DEER_OPT_TF32_PAIR:
      g_tf32_pair = value ? 1 : 0;
      return DEER_OK;
    case DEER_OPT_H16_PAIR:
      g_h16_pair = value ? 1 : 0;
      return DEER_OK;
    case DEER_OPT_LSTM_COLSPLIT:
      g_lstm_colsplit = value ? 1 : 0;
      return DEER_OK;
    case DEER_OPT_LSTM_DUAL:
      g_lstm_dual = value;
      return DEER_OK;
    case DEER_OPT_LSTM_HALFSPLIT:
      g_lstm_halfsplit = value ? 1 : 0;
      return DEER_OK;
    case DEER_OPT_LSTM_CARVEOUT:
      g_lstm_carveout = value ? 1 : 0;
      return DEER_OK;
    case DEER_OPT_LSTM_XIN:
      g_lstm_xin = value ? 1 : 0;
      g_lstm_xin_dual = (value == 3) ? 0 : 1;
      return DEER_OK;
    case DEER_OPT_LSTM_STASYNC:
      g_lstm_stasync = value ? 1 : 0;
      return DEER_OK;
    case DEER_OPT_LSTM_KEEP16:
      g_lstm_keep16 = value ? 1 : 0;
      return DEER_OK;
    case DEER_OPT_NIG_PIPELINE:
      g_nig_pipe = value ? 1 : 0;
      return DEER_OK;
    case DEER_OPT_SMALL_GEMM:
      g_small_engine = value == 1 ? 1 : 2;
      return DEER_OK;
    default:
      set_error("set_option: unknown option %d", option);
      return DEER_ERR_INVALID;
  }
}

int deer_gemm_rowterm(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                      long long ldc, int M, int N, int K, const float* w, const float* v, int nb, int nt, int time_major,
                      void* stream) {
  DEER_CHECK_ARG(A && B && C && w && v && M > 0 && N > 0 && K > 0 && nb > 0 && nt > 0, "gemm_rowterm: bad args");
  DEER_CHECK_ARG(lda >= (transA ? M : K) && ldb >= (transB ? K : N) && ldc >= N, "gemm_rowterm: leading dimension too small");
  g_engine_calls[DEER_ENGINE_TF32_PAIR]++;
  return gemm_tf32_pair_rowterm(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, w, v, nb, nt, time_major,
                                (cudaStream_t)stream);
}

int deer_gemm(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
              long long ldc, int M, int N, int K, const float* bias, int act, float beta, int batch, long long sA,
              long long sB, long long sC, long long sBias, int engine, void* stream) {
  DEER_CHECK_ARG(A && B && C, "gemm: null pointer");
  DEER_CHECK_ARG(M > 0 && N > 0 && K > 0 && batch > 0, "gemm: empty shape");
  // a leading dimension smaller than the row length = overlapping rows (sliding-window operand); only the CTA-pair TMA
  // kernels implement it (TMA reads any 16-byte-multiple pitch; overlapping C rows need the reduce-add epilogue)
  const bool overlapA = !transA && lda < K, overlapB = !transB && ldb < N, overlapC = ldc < N;
  DEER_CHECK_ARG((lda >= (transA ? M : K) || overlapA) && (ldb >= (transB ? K : N) || overlapB) && lda > 0 && ldb > 0 &&
                     ldc > 0,
                 "gemm: leading dimension too small");
  if (overlapA || overlapB || overlapC) {
    if (engine == DEER_GEMM_SIMT || batch != 1 || (overlapC && beta != 1.f) ||
        !gemm_tcgen05_supported(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, batch, sA, sB, sC) ||
        !gemm_tf32_pair_supported(transA, transB, M, N, K, ldc, bias, act, beta)) {
      set_error("gemm: overlapping-row operands need the CTA-pair TF32 engine (M=%d N=%d K=%d lda=%lld ldb=%lld ldc=%lld)",
                M, N, K, lda, ldb, ldc);
      return DEER_ERR_UNSUPPORTED;
    }
    g_engine_calls[DEER_ENGINE_TF32_PAIR]++;
    return gemm_tf32_pair(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, act, beta, (cudaStream_t)stream);
  }
  DEER_CHECK_ARG(act >= 0 && act <= 3, "gemm: bad activation");
  DEER_CHECK_ARG(beta == 0.f || beta == 1.f, "gemm: beta must be 0 or 1");
  cudaStream_t st = (cudaStream_t)stream;
  if (engine == DEER_GEMM_TF32X3)
    return gemm_x3_plain(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, act, beta, batch, sA, sB, sC, sBias, st);
  if (engine == DEER_GEMM_SIMT) {
    g_engine_calls[DEER_ENGINE_SIMT]++;
    return gemm_simt(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, act, beta, batch, sA, sB, sC, sBias, st);
  }
  const bool ok = gemm_tcgen05_supported(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, batch, sA, sB, sC);
  if (engine == DEER_GEMM_TF32 && !ok) {
    set_error("gemm: shape/alignment not supported by the tcgen05 engine (M=%d N=%d K=%d lda=%lld ldb=%lld)", M, N, K,
              lda, ldb);
    return DEER_ERR_UNSUPPORTED;
  }
  if (ok && gemm_tf32_pair_supported(transA, transB, M, N, K, ldc, bias, act, beta)) {
    g_engine_calls[DEER_ENGINE_TF32_PAIR]++;
    return gemm_tf32_pair(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, act, beta, st);
  }
  if (ok) {
    g_engine_calls[DEER_ENGINE_TF32]++;
    return gemm_tcgen05(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, act, beta, batch, sA, sB, sC, sBias, st);
  }
  g_engine_calls[DEER_ENGINE_SIMT]++;
  return gemm_simt(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, bias, act, beta, batch, sA, sB, sC, sBias, st);
}

}  // extern "C"
