// Bidirectional LSTM layer recurrence, time-major (encoders.py:82-89, call :380).
//
// Engine DEER_GEMM_SIMT ("stepwise"): per time step one recurrent GEMM launch covering both directions
// (gates_t += h_{t-1} W_hh^T, fp32) followed by one fused gate-nonlinearity/cell kernel.  It is exact fp32 and
// is the on-device reference the persistent cluster kernel (lstm_persistent.cu) is validated against.
//
// Layouts: gates [T,B,2,4H] (gate order i,f,g,o as nn.LSTM), h_out [T,B,2H], c_all [T,B,2,H].
// Direction 0 walks t = 0..T-1, direction 1 walks t = T-1..0; both advance in the same launch.
#include "common.cuh"

namespace deer {

int gemm_simt(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
              long long ldc, int M, int N, int K, const float* bias, int act, float beta, int batch, long long sA,
              long long sB, long long sC, long long sBias, cudaStream_t stream);
int gemm_tcgen05(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                 long long ldc, int M, int N, int K, const float* bias, int act, float beta, int batch, long long sA,
                 long long sB, long long sC, long long sBias, cudaStream_t stream);
bool gemm_tcgen05_supported(const float* A, long long lda, int transA, const float* B, long long ldb, int transB,
                            const float* C, long long ldc, int M, int N, int K, int batch, long long sA, long long sB,
                            long long sC);

bool lstm_persistent_supported(const float* gates, const float* h_out, const float* c_all, const float* w_fwd,
                               const float* w_rev, int T, int B, int H);
int lstm_fwd_persistent(float* gates, const float* w_fwd, const float* w_rev, float* h_out, float* c_all, int T, int B,
                        int keep, cudaStream_t stream);

// recurrent step GEMM for both directions: fp32 SIMT (one batched launch) or TF32 tcgen05 (one launch per direction)
static int step_gemm(bool tf32, const float* A, long long lda, const float* B, long long ldb, int transB, float* C,
                     long long ldc, int M, int N, int K, float beta, long long sA, long long sB, long long sC,
                     cudaStream_t stream) {
  if (tf32 && gemm_tcgen05_supported(A, lda, 0, B, ldb, transB, C, ldc, M, N, K, 1, 0, 0, 0) &&
      gemm_tcgen05_supported(A + sA, lda, 0, B + sB, ldb, transB, C + sC, ldc, M, N, K, 1, 0, 0, 0)) {
    for (int d = 0; d < 2; d++) {
      int rc = gemm_tcgen05(A + d * sA, lda, 0, B + d * sB, ldb, transB, C + d * sC, ldc, M, N, K, nullptr,
                            DEER_ACT_NONE, beta, 1, 0, 0, 0, 0, stream);
      if (rc) return rc;
    }
    return DEER_OK;
  }
  return gemm_simt(A, lda, 0, B, ldb, transB, C, ldc, M, N, K, nullptr, DEER_ACT_NONE, beta, 2, sA, sB, sC, 0, stream);
}

// one thread per (b, dir, j)
__global__ void __launch_bounds__(256) lstm_cell_fwd_kernel(float* __restrict__ gates, float* __restrict__ h_out,
                                                            float* __restrict__ c_all, float* __restrict__ c_work,
                                                            int step, int T, int B, int H) {
  DEER_PDL_ENTRY();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * 2 * H) return;
  const int j = (int)(idx % H);
  const int dir = (int)((idx / H) & 1);
  const int b = (int)(idx / (2 * H));
  const int t = dir ? T - 1 - step : step;
  float* g = gates + (((long long)t * B + b) * 2 + dir) * 4 * H;
  const float gi = sigmoid_f(g[j]);
  const float gf = sigmoid_f(g[H + j]);
  const float gg = tanh_f(g[2 * H + j]);
  const float go = sigmoid_f(g[3 * H + j]);
  float c_prev = 0.f;
  if (step > 0) {
    if (c_all) {
      const int tp = dir ? t + 1 : t - 1;
      c_prev = c_all[(((long long)tp * B + b) * 2 + dir) * H + j];
    } else {
      c_prev = c_work[((long long)b * 2 + dir) * H + j];
    }
  }
  const float c = fmaf(gf, c_prev, gi * gg);
  const float h = go * tanh_f(c);
  g[j] = gi;
  g[H + j] = gf;
  g[2 * H + j] = gg;
  g[3 * H + j] = go;
  if (c_all) c_all[(((long long)t * B + b) * 2 + dir) * H + j] = c;
  else c_work[((long long)b * 2 + dir) * H + j] = c;
  h_out[((long long)t * B + b) * 2 * H + dir * H + j] = h;
}

// gates (post-activation) -> pre-activation gradients in place
__global__ void __launch_bounds__(256) lstm_cell_bwd_kernel(float* __restrict__ gates, const float* __restrict__ c_all,
                                                            const float* __restrict__ dh_out,
                                                            const float* __restrict__ dh_work,
                                                            float* __restrict__ dc_work, int step, int first, int T,
                                                            int B, int H) {
  DEER_PDL_ENTRY();
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)B * 2 * H) return;
  const int j = (int)(idx % H);
  const int dir = (int)((idx / H) & 1);
  const int b = (int)(idx / (2 * H));
  const int t = dir ? T - 1 - step : step;
  float* g = gates + (((long long)t * B + b) * 2 + dir) * 4 * H;
  const float gi = g[j], gf = g[H + j], gg = g[2 * H + j], go = g[3 * H + j];
  const long long wi = ((long long)b * 2 + dir) * H + j;
  float dh = dh_out[((long long)t * B + b) * 2 * H + dir * H + j];
  float dc = 0.f;
  if (!first) {
    dh += dh_work[wi];
    dc = dc_work[wi];
  }
  const float c = c_all[(((long long)t * B + b) * 2 + dir) * H + j];
  float c_prev = 0.f;
  if (step > 0) {
    const int tp = dir ? t + 1 : t - 1;
    c_prev = c_all[(((long long)tp * B + b) * 2 + dir) * H + j];
  }
  const float tc = tanh_f(c);
  const float d_o = dh * tc;
  dc = fmaf(dh * go, 1.f - tc * tc, dc);
  const float d_i = dc * gg, d_g = dc * gi, d_f = dc * c_prev;
  dc_work[wi] = dc * gf;
  g[j] = d_i * gi * (1.f - gi);
  g[H + j] = d_f * gf * (1.f - gf);
  g[2 * H + j] = d_g * (1.f - gg * gg);
  g[3 * H + j] = d_o * go * (1.f - go);
}

int lstm_fwd_stepwise(float* gates, const float* w_hh, long long w_stride, float* h_out, float* c_all, float* c_work, int T, int B, int H,
                      bool tf32, cudaStream_t stream) {
  const long long n = (long long)B * 2 * H;
  const unsigned grid = (unsigned)cdiv(n, 256);
  const long long row_h = (long long)B * 2 * H, row_g = (long long)B * 8 * H;
  for (int step = 0; step < T; step++) {
    if (step > 0) {
      const int t0 = step, t1 = T - 1 - step;
      const int p0 = step - 1, p1 = T - step;
      const float* A = h_out + p0 * row_h;
      const long long sA = (long long)(p1 - p0) * row_h + H;
      float* C = gates + t0 * row_g;
      const long long sC = (long long)(t1 - t0) * row_g + 4 * H;
      int rc = step_gemm(tf32, A, 2 * H, w_hh, H, 1, C, 8 * H, B, 4 * H, H, 1.f, sA, w_stride, sC, stream);
      if (rc) return rc;
    }
    DEER_LAUNCH(lstm_cell_fwd_kernel, grid, 256, 0, stream, gates, h_out, c_all, c_work, step, T, B, H);
  }
  return DEER_OK;
}

int lstm_bwd_stepwise(float* gates, const float* w_hh, long long w_stride, const float* c_all, const float* dh_out, float* dh_work,
                      float* dc_work, int T, int B, int H, bool tf32, cudaStream_t stream) {
  const long long n = (long long)B * 2 * H;
  const unsigned grid = (unsigned)cdiv(n, 256);
  const long long row_g = (long long)B * 8 * H;
  for (int step = T - 1; step >= 0; step--) {
    DEER_LAUNCH(lstm_cell_bwd_kernel, grid, 256, 0, stream, gates, c_all, dh_out, dh_work, dc_work, step,
                step == T - 1 ? 1 : 0, T, B, H);
    if (step > 0) {
      // dh_work[b,dir,:] = dgates_t[b,dir,:] @ W_hh[dir]   ([4H,H], no transpose)
      const int t0 = step, t1 = T - 1 - step;
      const float* A = gates + t0 * row_g;
      const long long sA = (long long)(t1 - t0) * row_g + 4 * H;
      int rc = step_gemm(tf32, A, 8 * H, w_hh, H, 0, dh_work, 2 * H, B, H, 4 * H, 0.f, sA, w_stride, H, stream);
      if (rc) return rc;
    }
  }
  return DEER_OK;
}

}  // namespace deer

using namespace deer;

extern "C" {

int deer_lstm_fwd(float* gates, const float* w_hh_fwd, const float* w_hh_rev, float* h_out, float* c_out, float* c_work,
                  int T, int B, int H, int engine, void* stream) {
  DEER_CHECK_ARG(gates && w_hh_fwd && w_hh_rev && h_out && T > 0 && B > 0 && H > 0, "lstm_fwd: bad args");
  DEER_CHECK_ARG(c_out || c_work, "lstm_fwd: need c_out or c_work");
  // natural-layout engines: SIMT = exact-fp32 stepwise reference; AUTO/TF32/DEER_LSTM_PERSISTENT_V1 = the round-1
  // 8-CTA TF32 forward kernel (h exchange through L2/TMA multicast) when H == 256; DEER_LSTM_STEPWISE_TF32 = stepwise
  // with the tcgen05 GEMM per step.  The production path is deer_lstm_cluster_fwd/bwd (lstm_cluster.cu).
  if (engine != DEER_GEMM_SIMT && engine != DEER_LSTM_STEPWISE_TF32 &&
      lstm_persistent_supported(gates, h_out, c_out, w_hh_fwd, w_hh_rev, T, B, H))
    return lstm_fwd_persistent(gates, w_hh_fwd, w_hh_rev, h_out, c_out, T, B, c_out != nullptr, (cudaStream_t)stream);
  return lstm_fwd_stepwise(gates, w_hh_fwd, (long long)(w_hh_rev - w_hh_fwd), h_out, c_out, c_work, T, B, H,
                           engine == DEER_LSTM_STEPWISE_TF32, (cudaStream_t)stream);
}

int deer_lstm_bwd(float* gates, const float* w_hh_fwd, const float* w_hh_rev, const float* c_all, const float* dh_out,
                  float* dh_work, float* dc_work, int T, int B, int H, int engine, void* stream) {
  DEER_CHECK_ARG(gates && w_hh_fwd && w_hh_rev && c_all && dh_out && dh_work && dc_work && T > 0 && B > 0 && H > 0,
                 "lstm_bwd: bad args");
  return lstm_bwd_stepwise(gates, w_hh_fwd, (long long)(w_hh_rev - w_hh_fwd), c_all, dh_out, dh_work, dc_work, T, B, H,
                           engine == DEER_LSTM_STEPWISE_TF32, (cudaStream_t)stream);
}

}  // extern "C"
