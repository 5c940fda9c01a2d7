// Core of nn.MultiheadAttention for the TWO-token modality sequence of TrimodalFusion (fusion.py:315-332):
// per (batch, head): 2x2 scaled-dot-product scores, softmax over keys, context = P V, head-averaged weights.
// One block per batch element, one warp per head (the in_proj / out_proj GEMMs are done by deer_gemm).
#include "common.cuh"

namespace deer {

// qkv [B,2,3E]; ctx [B,2,E]; attw [B,2,2]; probs [B,heads,2,2]
__global__ void mha2_fwd_kernel(const float* __restrict__ qkv, float* __restrict__ ctx, float* __restrict__ ctx_mean,
                                float* __restrict__ attw, float* __restrict__ probs, int B, int E, int heads) {
  DEER_PDL_ENTRY();
  __shared__ float pw[32][4];
  const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = E / heads;
  const float scale = rsqrtf((float)d);
  const float* t0 = qkv + (long long)b * 6 * E;
  const float* t1 = t0 + 3 * E;
  const int o = h * d;
  float s00 = 0.f, s01 = 0.f, s10 = 0.f, s11 = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float q0 = t0[o + j] * scale, q1 = t1[o + j] * scale;
    const float k0 = t0[E + o + j], k1 = t1[E + o + j];
    s00 = fmaf(q0, k0, s00);
    s01 = fmaf(q0, k1, s01);
    s10 = fmaf(q1, k0, s10);
    s11 = fmaf(q1, k1, s11);
  }
  s00 = warp_sum(s00);
  s01 = warp_sum(s01);
  s10 = warp_sum(s10);
  s11 = warp_sum(s11);
  const float m0 = fmaxf(s00, s01), m1 = fmaxf(s10, s11);
  float e00 = expf(s00 - m0), e01 = expf(s01 - m0), e10 = expf(s10 - m1), e11 = expf(s11 - m1);
  const float i0 = 1.f / (e00 + e01), i1 = 1.f / (e10 + e11);
  e00 *= i0;
  e01 *= i0;
  e10 *= i1;
  e11 *= i1;
  for (int j = lane; j < d; j += 32) {
    const float v0 = t0[2 * E + o + j], v1 = t1[2 * E + o + j];
    const float c0 = e00 * v0 + e01 * v1, c1 = e10 * v0 + e11 * v1;
    if (ctx) {
      ctx[((long long)b * 2 + 0) * E + o + j] = c0;
      ctx[((long long)b * 2 + 1) * E + o + j] = c1;
    }
    if (ctx_mean) ctx_mean[(long long)b * E + o + j] = 0.5f * (c0 + c1);  // attended.mean(dim=1), fusion.py:335
  }
  if (lane == 0) {
    pw[h][0] = e00;
    pw[h][1] = e01;
    pw[h][2] = e10;
    pw[h][3] = e11;
    float* pr = probs + ((long long)b * heads + h) * 4;
    pr[0] = e00;
    pr[1] = e01;
    pr[2] = e10;
    pr[3] = e11;
  }
  __syncthreads();
  if (threadIdx.x < 4 && attw) {
    float a = 0.f;
    for (int k = 0; k < heads; k++) a += pw[k][threadIdx.x];
    attw[(long long)b * 4 + threadIdx.x] = a / heads;
  }
}

__global__ void mha2_bwd_kernel(const float* __restrict__ dctx, const float* __restrict__ dctx_mean,
                                const float* __restrict__ dattw,
                                const float* __restrict__ qkv, const float* __restrict__ probs,
                                float* __restrict__ dqkv, int B, int E, int heads) {
  DEER_PDL_ENTRY();
  const int b = blockIdx.x, h = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int d = E / heads;
  const float scale = rsqrtf((float)d);
  const float* t0 = qkv + (long long)b * 6 * E;
  const float* t1 = t0 + 3 * E;
  float* g0 = dqkv + (long long)b * 6 * E;
  float* g1 = g0 + 3 * E;
  const int o = h * d;
  const float* pr = probs + ((long long)b * heads + h) * 4;
  const float p00 = pr[0], p01 = pr[1], p10 = pr[2], p11 = pr[3];
  const float* dc0 = dctx ? dctx + ((long long)b * 2 + 0) * E + o : nullptr;
  const float* dc1 = dctx ? dctx + ((long long)b * 2 + 1) * E + o : nullptr;
  const float* dcm = dctx_mean ? dctx_mean + (long long)b * E + o : nullptr;
  // dp_ij = dctx_i . v_j (+ dattw_ij / heads)
  float a00 = 0.f, a01 = 0.f, a10 = 0.f, a11 = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float v0 = t0[2 * E + o + j], v1 = t1[2 * E + o + j];
    float c0 = dc0 ? dc0[j] : 0.f, c1 = dc1 ? dc1[j] : 0.f;
    if (dcm) {
      const float hm = 0.5f * dcm[j];
      c0 += hm;
      c1 += hm;
    }
    a00 = fmaf(c0, v0, a00);
    a01 = fmaf(c0, v1, a01);
    a10 = fmaf(c1, v0, a10);
    a11 = fmaf(c1, v1, a11);
    g0[2 * E + o + j] = p00 * c0 + p10 * c1;  // dv_0
    g1[2 * E + o + j] = p01 * c0 + p11 * c1;  // dv_1
  }
  a00 = warp_sum(a00);
  a01 = warp_sum(a01);
  a10 = warp_sum(a10);
  a11 = warp_sum(a11);
  if (dattw) {
    const float ih = 1.f / heads;
    a00 += dattw[(long long)b * 4 + 0] * ih;
    a01 += dattw[(long long)b * 4 + 1] * ih;
    a10 += dattw[(long long)b * 4 + 2] * ih;
    a11 += dattw[(long long)b * 4 + 3] * ih;
  }
  const float r0 = p00 * a00 + p01 * a01, r1 = p10 * a10 + p11 * a11;
  const float ds00 = p00 * (a00 - r0) * scale, ds01 = p01 * (a01 - r0) * scale;
  const float ds10 = p10 * (a10 - r1) * scale, ds11 = p11 * (a11 - r1) * scale;
  for (int j = lane; j < d; j += 32) {
    const float q0 = t0[o + j], q1 = t1[o + j];
    const float k0 = t0[E + o + j], k1 = t1[E + o + j];
    g0[o + j] = ds00 * k0 + ds01 * k1;      // dq_0
    g1[o + j] = ds10 * k0 + ds11 * k1;      // dq_1
    g0[E + o + j] = ds00 * q0 + ds10 * q1;  // dk_0
    g1[E + o + j] = ds01 * q0 + ds11 * q1;  // dk_1
  }
}

}  // namespace deer

using namespace deer;

extern "C" {

int deer_mha2_fwd(const float* qkv, float* ctx, float* ctx_mean, float* attw, float* probs, int B, int E, int heads,
                  void* stream) {
  DEER_CHECK_ARG(qkv && (ctx || ctx_mean) && probs && B > 0 && E > 0 && heads > 0 && heads <= 32 && E % heads == 0,
                 "mha2_fwd: bad args");
  DEER_LAUNCH(mha2_fwd_kernel, B, heads * 32, 0, stream, qkv, ctx, ctx_mean, attw, probs, B, E, heads);
  return DEER_OK;
}

int deer_mha2_bwd(const float* dctx, const float* dctx_mean, const float* dattw, const float* qkv, const float* probs,
                  float* dqkv, int B, int E, int heads, void* stream) {
  DEER_CHECK_ARG((dctx || dctx_mean) && qkv && probs && dqkv && B > 0 && E > 0 && heads > 0 && heads <= 32 && E % heads == 0,
                 "mha2_bwd: bad args");
  DEER_LAUNCH(mha2_bwd_kernel, B, heads * 32, 0, stream, dctx, dctx_mean, dattw, qkv, probs, dqkv, B, E, heads);
  return DEER_OK;
}

}  // extern "C"
