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

// one thread per (b, dir, j)
__global__ void __launch_bounds__(256) lstm_cell_fwd_kernel(float* __restrict__ gates, float* __restrict__ h_out,
                                                            float* __restrict__ c_all, float* __restrict__ c_work,
                                                            int step, int T, int B, int H) {
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
                      cudaStream_t stream) {
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
      int rc = gemm_simt(A, 2 * H, 0, w_hh, H, 1, C, 8 * H, B, 4 * H, H, nullptr, DEER_ACT_NONE, 1.f, 2, sA,
                         w_stride, sC, 0, stream);
      if (rc) return rc;
    }
    DEER_LAUNCH(lstm_cell_fwd_kernel, grid, 256, 0, stream, gates, h_out, c_all, c_work, step, T, B, H);
  }
  return DEER_OK;
}

int lstm_bwd_stepwise(float* gates, const float* w_hh, long long w_stride, const float* c_all, const float* dh_out, float* dh_work,
                      float* dc_work, int T, int B, int H, cudaStream_t stream) {
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
      int rc = gemm_simt(A, 8 * H, 0, w_hh, H, 0, dh_work, 2 * H, B, H, 4 * H, nullptr, DEER_ACT_NONE, 0.f, 2, sA,
                         w_stride, H, 0, stream);
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
  (void)engine;
  return lstm_fwd_stepwise(gates, w_hh_fwd, (long long)(w_hh_rev - w_hh_fwd), h_out, c_out, c_work, T, B, H,
                           (cudaStream_t)stream);
}

int deer_lstm_bwd(float* gates, const float* w_hh_fwd, const float* w_hh_rev, const float* c_all, const float* dh_out,
                  float* dh_work, float* dc_work, int T, int B, int H, int engine, void* stream) {
  DEER_CHECK_ARG(gates && w_hh_fwd && w_hh_rev && c_all && dh_out && dh_work && dc_work && T > 0 && B > 0 && H > 0,
                 "lstm_bwd: bad args");
  (void)engine;
  return lstm_bwd_stepwise(gates, w_hh_fwd, (long long)(w_hh_rev - w_hh_fwd), c_all, dh_out, dh_work, dc_work, T, B, H,
                           (cudaStream_t)stream);
}

}  // extern "C"
