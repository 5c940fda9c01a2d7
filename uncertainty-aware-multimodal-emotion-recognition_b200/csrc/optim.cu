// Trainer step over flat fp32 buffers (training.py:121-150 AdamW groups, :219 clip_grad_norm_, :224 step):
// one sum-of-squares reduction for the global gradient norm and one fused clip + AdamW update per LR group.
// HBM-bound: AdamW reads p,g,m,v (16 B/param) and writes p,m,v (12 B/param).  No host synchronisation: the clip
// coefficient is computed on the device from the reduced norm.
#include "common.cuh"

namespace deer {

__global__ void __launch_bounds__(256) sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
  DEER_PDL_ENTRY();
  __shared__ float red[32];
  float s = 0.f;
  const long long n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x4[i];
    s += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    s += x[i] * x[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

__global__ void __launch_bounds__(256) adamw_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                    float* __restrict__ m, float* __restrict__ v, long long n, float lr,
                                                    float beta1, float beta2, float eps, float wd, float bc1,
                                                    float bc2_sqrt, const float* __restrict__ sumsq, float max_norm,
                                                    float grad_scale, const long long* __restrict__ step_dev,
                                                    const float* __restrict__ lr_dev) {
  DEER_PDL_ENTRY();
  // CUDA-graph replays: the step count (bias corrections) and the learning rate come from device memory, so one
  // captured training step stays valid across steps and LR-schedule changes
  if (step_dev) {
    const float t = (float)(*step_dev + 1);
    bc1 = 1.f - exp2f(t * log2f(beta1));
    bc2_sqrt = sqrtf(1.f - exp2f(t * log2f(beta2)));
  }
  if (lr_dev) lr *= *lr_dev;
  float clip = 1.f;
  if (sumsq && max_norm > 0.f) {
    const float total_norm = sqrtf(*sumsq) * grad_scale;
    clip = fminf(1.f, max_norm / (total_norm + 1e-6f));
  }
  const float gs = grad_scale * clip;
  const float step_size = lr / bc1;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i] * (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    pi -= step_size * mi / denom;
    p[i] = pi;
    m[i] = mi;
    v[i] = vi;
  }
}
// zero-fill of the trainer's flat gradient buffer (+ an optional second small buffer: the gradient-norm accumulator)
__global__ void __launch_bounds__(256) fill_zero_kernel(float* __restrict__ x, long long n, float* __restrict__ y,
                                                        long long ny) {
  DEER_PDL_ENTRY();
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  const bool al = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  const long long n4 = al ? (n >> 2) : 0;
  float4* x4 = reinterpret_cast<float4*>(x);
  for (long long i = tid; i < n4; i += nth) x4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (long long i = (n4 << 2) + tid; i < n; i += nth) x[i] = 0.f;
  if (y)
    for (long long i = tid; i < ny; i += nth) y[i] = 0.f;
}

__global__ void step_increment_kernel(long long* step) {
  DEER_PDL_ENTRY();
  *step += 1;
}

}  // namespace deer

using namespace deer;

extern "C" {

int deer_sumsq(const float* x, long long n, float* out, void* stream) {
  DEER_CHECK_ARG(x && out && n > 0, "sumsq: bad args");
  DEER_CHECK_ARG((reinterpret_cast<uintptr_t>(x) & 15) == 0, "sumsq: x must be 16-byte aligned");
  long long g = cdiv(n, 256 * 8);
  if (g > kNumSMs * 8) g = kNumSMs * 8;
  if (g < 1) g = 1;
  DEER_LAUNCH(sumsq_kernel, (unsigned)g, 256, 0, stream, x, n, out);
  return DEER_OK;
}

int deer_fill_zero(float* x, long long n, float* y, long long ny, void* stream) {
  DEER_CHECK_ARG(x && n > 0 && (y == nullptr || ny > 0), "fill_zero: bad args");
  long long g = cdiv(n, 256 * 8);
  if (g > kNumSMs * 8) g = kNumSMs * 8;
  if (g < 1) g = 1;
  DEER_LAUNCH(fill_zero_kernel, (unsigned)g, 256, 0, stream, x, n, y, ny);
  return DEER_OK;
}

int deer_step_increment(long long* step, void* stream) {
  DEER_CHECK_ARG(step, "step_increment: null pointer");
  DEER_LAUNCH(step_increment_kernel, 1, 1, 0, stream, step);
  return DEER_OK;
}

int deer_adamw(float* p, const float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
               float eps, float weight_decay, int step, const float* sumsq, float max_norm, float grad_scale,
               const long long* step_dev, const float* lr_dev, void* stream) {
  DEER_CHECK_ARG(p && g && m && v && n > 0 && (step >= 1 || step_dev), "adamw: bad args");
  const float bc1 = step_dev ? 1.f : 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = step_dev ? 1.f : sqrtf(1.f - powf(beta2, (float)step));
  long long gr = cdiv(n, 256 * 4);
  if (gr > kNumSMs * 8) gr = kNumSMs * 8;
  if (gr < 1) gr = 1;
  DEER_LAUNCH(adamw_kernel, (unsigned)gr, 256, 0, stream, p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1,
              bc2_sqrt, sumsq, max_norm, grad_scale, step_dev, lr_dev);
  return DEER_OK;
}

}  // extern "C"
