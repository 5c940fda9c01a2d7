// Integer / reduction kernels either side of the hot path.
//
//   deer_linguistic_features   EnhancedTextEncoder.extract_linguistic_features (encoders.py:648-699): per-sample integer
//                              statistics of the valid token ids -> [B,10] fp32, bit-exact with the reference's CPU result.
//   deer_metrics_moments       one pass over (pred, target): the fp64 moments from which CCC (metrics.py:59-103), MAE
//                              (:105-114), RMSE (:116-125) and Cohen's d (:190-211) follow on the host.
//   deer_uce_*                 uncertainty_calibration_error (metrics.py:214-279): per-sample means, exact order statistics
//                              by 8-bit radix select (the quantile bin boundaries) and the binned sums.
// All of them are HBM-bound single-pass streams over at most 12 B per (sample, dimension).
#include "common.cuh"

namespace deer {

// ------------------------------------------------------------------ linguistic features
// One warp per sample; T tokens staged in shared memory (valid ones compacted), then an all-pairs multiplicity count:
// T <= 512 in every caller (BERT max_length 128), so O(T^2 / 32) compares per lane is a few hundred instructions.
constexpr int LING_MAX_T = 1024;
constexpr int LING_WARPS = 4;

__global__ void __launch_bounds__(LING_WARPS * 32) linguistic_features_kernel(const long long* __restrict__ ids,
                                                                              const long long* __restrict__ mask,
                                                                              float* __restrict__ out, int B, int T,
                                                                              int max_length) {
  DEER_PDL_ENTRY();
  extern __shared__ long long tok[];  // [LING_WARPS][T]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int b = blockIdx.x * LING_WARPS + w;
  if (b >= B) return;
  long long* my = tok + (size_t)w * T;
  // ordered compaction of the valid tokens (mask.bool(): any non-zero value)
  int n = 0;
  for (int t0 = 0; t0 < T; t0 += 32) {
    const int t = t0 + lane;
    const bool v = t < T && mask[(size_t)b * T + t] != 0;
    const unsigned bal = __ballot_sync(0xffffffffu, v);
    if (v) my[n + __popc(bal & ((1u << lane) - 1u))] = ids[(size_t)b * T + t];
    n += __popc(bal);
  }
  __syncwarp();
  int uniq = 0, maxmult = 0, punct = 0, special = 0;
  long long maxid = -1;
  for (int j = lane; j < n; j += 32) {
    const long long v = my[j];
    int mult = 0;
    bool first = true;
    for (int i = 0; i < n; i++) {
      const bool eq = my[i] == v;
      mult += eq;
      first = first && !(eq && i < j);
    }
    uniq += first;
    maxmult = max(maxmult, mult);
    maxid = v > maxid ? v : maxid;
    punct += (v >= 999 && v <= 1030);
    special += (v >= 100 && v <= 999);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uniq += __shfl_xor_sync(0xffffffffu, uniq, o);
    punct += __shfl_xor_sync(0xffffffffu, punct, o);
    special += __shfl_xor_sync(0xffffffffu, special, o);
    maxmult = max(maxmult, __shfl_xor_sync(0xffffffffu, maxmult, o));
    const long long om = __shfl_xor_sync(0xffffffffu, maxid, o);
    maxid = om > maxid ? om : maxid;
  }
  if (lane < 10) {
    float f = 0.f;
    const int den = n > 0 ? n : 1;
    if (n > 0) {
      // python-float quotients (double) stored into a float32 tensor; tensor quotients are fp32 divisions
      if (lane == 0) f = (float)((double)n / (double)max_length);
      if (lane == 1) f = (float)((double)uniq / (double)den);
      if (lane == 2) f = __fdiv_rn((float)n, (float)(maxid + 1));  // mean(bincount): sum = n, len = max id + 1
      if (lane == 3) f = (float)maxmult;
      if (lane == 4) f = __fdiv_rn((float)punct, (float)den);
      if (lane == 5) f = __fdiv_rn((float)special, (float)den);
    }
    out[(size_t)b * 10 + lane] = f;
  }
}

// ------------------------------------------------------------------ regression-metric moments
constexpr int MOM_THREADS = 192;  // multiple of every supported D
constexpr int NMOM = DEER_METRICS_NMOM;

__global__ void __launch_bounds__(MOM_THREADS) metrics_moments_kernel(const float* __restrict__ pred,
                                                                      const float* __restrict__ target, long long total,
                                                                      int D, double* __restrict__ out) {
  DEER_PDL_ENTRY();
  __shared__ double red[NMOM][MOM_THREADS];
  const int tid = threadIdx.x;
  double m[NMOM];
#pragma unroll
  for (int i = 0; i < NMOM; i++) m[i] = 0.0;
  const long long stride = (long long)gridDim.x * MOM_THREADS;  // multiple of D: the dimension is fixed per thread
  for (long long e = (long long)blockIdx.x * MOM_THREADS + tid; e < total; e += stride) {
    const float pf = __ldcs(pred + e), tf = __ldcs(target + e);
    if (pf != pf || tf != tf) continue;  // the reference drops NaN pairs (metrics.py:79-84)
    const double p = pf, t = tf, d = t - p;
    m[0] += 1.0;
    m[1] += t;
    m[2] += p;
    m[3] += t * t;
    m[4] += p * p;
    m[5] += t * p;
    m[6] += fabs(d);
    m[7] += d * d;
  }
#pragma unroll
  for (int i = 0; i < NMOM; i++) red[i][tid] = m[i];
  __syncthreads();
  const int first = (int)(((long long)blockIdx.x * MOM_THREADS) % D);
  for (int wk = tid; wk < D * NMOM; wk += MOM_THREADS) {
    const int d = wk / NMOM, s = wk % NMOM;
    int t0 = (d - first) % D;
    if (t0 < 0) t0 += D;
    double acc = 0.0;
    for (int t = t0; t < MOM_THREADS; t += D) acc += red[s][t];
    if (acc != 0.0) atomicAdd(out + d * NMOM + s, acc);
  }
}

// ------------------------------------------------------------------ uncertainty calibration error
__device__ __forceinline__ unsigned f2key(float f) {  // order-preserving float -> uint32
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
constexpr unsigned UCE_INVALID = 0xffffffffu;  // key of a dropped sample (sorts last; no finite float maps to it)

// per sample: err = mean_d |pred - target|, u = mean_d uncertainty, both in fp32 with numpy's left-to-right row sum
// (metrics.py:236-242); samples with NaN err / NaN or infinite u are dropped (:245)
__global__ void __launch_bounds__(256) uce_prepare_kernel(const float* __restrict__ pred,
                                                          const float* __restrict__ target,
                                                          const float* __restrict__ uncert, long long N, int D,
                                                          float* __restrict__ err_m, unsigned* __restrict__ keys,
                                                          unsigned long long* __restrict__ n_valid) {
  DEER_PDL_ENTRY();
  unsigned long long cnt = 0;
  for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < N; b += (long long)gridDim.x * blockDim.x) {
    float se = 0.f, su = 0.f;
    for (int d = 0; d < D; d++) {
      se = __fadd_rn(se, fabsf(__fsub_rn(pred[b * D + d], target[b * D + d])));
      su = __fadd_rn(su, uncert[b * D + d]);
    }
    if (D > 1) {
      se = __fdiv_rn(se, (float)D);
      su = __fdiv_rn(su, (float)D);
    }
    const bool ok = !(se != se) && !(su != su) && !isinf(su);
    err_m[b] = se;
    keys[b] = ok ? f2key(su) : UCE_INVALID;
    cnt += ok;
  }
  cnt = __reduce_add_sync(0xffffffffu, (unsigned)cnt);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(n_valid, cnt);
}

constexpr int UCE_MAX_RANKS = DEER_UCE_MAX_RANKS;
struct SelState {
  unsigned prefix[UCE_MAX_RANKS];
  unsigned long long k[UCE_MAX_RANKS];
};

// one 8-bit digit of the radix select: for every requested rank, the histogram of the current digit over the keys
// that share the rank's already-fixed high bits
__global__ void __launch_bounds__(256) uce_select_hist_kernel(const unsigned* __restrict__ keys, long long N, int pass,
                                                              int R, const SelState* __restrict__ st,
                                                              unsigned long long* __restrict__ hist) {
  DEER_PDL_ENTRY();
  __shared__ unsigned h[UCE_MAX_RANKS][256];
  __shared__ unsigned pre[UCE_MAX_RANKS];
  for (int i = threadIdx.x; i < R * 256; i += blockDim.x) (&h[0][0])[i] = 0u;
  if (threadIdx.x < R) pre[threadIdx.x] = st->prefix[threadIdx.x];
  __syncthreads();
  const int shift = 24 - 8 * pass;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const unsigned k = __ldcs(keys + i);
    const unsigned dig = (k >> shift) & 0xffu;
    if (pass == 0) {
      atomicAdd(&h[0][dig], 1u);  // every rank shares the first histogram
    } else {
      const unsigned hi = k >> (shift + 8);
      for (int r = 0; r < R; r++)
        if (hi == pre[r]) atomicAdd(&h[r][dig], 1u);
    }
  }
  __syncthreads();
  const int nr = pass == 0 ? 1 : R;
  for (int i = threadIdx.x; i < nr * 256; i += blockDim.x) {
    const unsigned v = (&h[0][0])[i];
    if (v) atomicAdd(hist + i, (unsigned long long)v);
  }
}

__global__ void uce_select_init_kernel(SelState init, SelState* __restrict__ st) {
  DEER_PDL_ENTRY();
  if (threadIdx.x == 0) *st = init;
}

// single block: walk each rank's histogram, fix the digit, carry the residual rank; clears the histogram for the next pass
__global__ void __launch_bounds__(32) uce_select_scan_kernel(int pass, int R, SelState* __restrict__ st,
                                                             unsigned long long* __restrict__ hist,
                                                             float* __restrict__ values) {
  DEER_PDL_ENTRY();
  const int r = threadIdx.x;
  if (r < R) {
    const unsigned long long* h = hist + (pass == 0 ? 0 : r * 256);
    unsigned long long k = st->k[r], cum = 0;
    int d = 0;
    for (; d < 255; d++) {
      const unsigned long long c = h[d];
      if (k < cum + c) break;
      cum += c;
    }
    st->k[r] = k - cum;
    const unsigned p = (st->prefix[r] << 8) | (unsigned)d;
    st->prefix[r] = p;
    if (pass == 3) values[r] = key2f(p);
  }
  __syncwarp();
  for (int i = threadIdx.x; i < R * 256; i += 32) hist[i] = 0ull;
}

// bins [edge_i, edge_{i+1}) on the sample-mean uncertainty (metrics.py:262-264): count, sum(1-u), sum(1-err)
__global__ void __launch_bounds__(256) uce_bins_kernel(const unsigned* __restrict__ keys,
                                                       const float* __restrict__ err_m, long long N, int n_bins,
                                                       const double* __restrict__ edges, double* __restrict__ out) {
  DEER_PDL_ENTRY();
  __shared__ double sedge[DEER_UCE_MAX_BINS + 1];
  __shared__ double acc[3][DEER_UCE_MAX_BINS];
  if (threadIdx.x <= n_bins) sedge[threadIdx.x] = edges[threadIdx.x];
  if (threadIdx.x < 3 * DEER_UCE_MAX_BINS) (&acc[0][0])[threadIdx.x] = 0.0;
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const unsigned k = __ldcs(keys + i);
    if (k == UCE_INVALID) continue;
    const float uf = key2f(k), ef = __ldcs(err_m + i);
    const double u = uf;
    for (int j = 0; j < n_bins; j++)
      if (u >= sedge[j] && u < sedge[j + 1]) {
        atomicAdd(&acc[0][j], 1.0);
        atomicAdd(&acc[1][j], (double)__fsub_rn(1.f, uf));
        atomicAdd(&acc[2][j], (double)__fsub_rn(1.f, ef));
      }
  }
  __syncthreads();
  if (threadIdx.x < 3 * DEER_UCE_MAX_BINS) {
    const int s = threadIdx.x / DEER_UCE_MAX_BINS, j = threadIdx.x % DEER_UCE_MAX_BINS;
    const double v = acc[s][j];
    if (j < n_bins && v != 0.0) atomicAdd(out + s * n_bins + j, v);
  }
}

static int stream_grid1(long long n, int threads) {
  long long g = cdiv(n, threads);
  const long long cap = (long long)kNumSMs * 8;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

}  // namespace deer

using namespace deer;

extern "C" {

int deer_linguistic_features(const long long* input_ids, const long long* attention_mask, float* features, int B, int T,
                             int max_length, void* stream) {
  DEER_CHECK_ARG(input_ids && attention_mask && features && B > 0 && T > 0 && max_length > 0,
                 "linguistic_features: bad args");
  if (T > LING_MAX_T) {
    set_error("linguistic_features: T=%d > %d unsupported", T, LING_MAX_T);
    return DEER_ERR_UNSUPPORTED;
  }
  const size_t smem = (size_t)LING_WARPS * T * sizeof(long long);
  DEER_LAUNCH(linguistic_features_kernel, (int)cdiv(B, LING_WARPS), LING_WARPS * 32, smem, stream, input_ids,
              attention_mask, features, B, T, max_length);
  return DEER_OK;
}

int deer_metrics_moments(const float* pred, const float* target, long long N, int D, double* moments, void* stream) {
  DEER_CHECK_ARG(pred && target && moments && N > 0, "metrics_moments: bad args");
  if (D < 1 || D > 8 || MOM_THREADS % D != 0) {
    set_error("metrics_moments: D=%d unsupported (need a divisor of %d, <=8)", D, MOM_THREADS);
    return DEER_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaMemsetAsync(moments, 0, sizeof(double) * NMOM * D, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_status(e, "metrics_moments memset");
  DEER_LAUNCH(metrics_moments_kernel, stream_grid1(N * D, MOM_THREADS), MOM_THREADS, 0, stream, pred, target, N * D, D,
              moments);
  return DEER_OK;
}

int deer_uce_prepare(const float* pred, const float* target, const float* uncert, long long N, int D, float* err_mean,
                     unsigned* keys, unsigned long long* n_valid, void* stream) {
  DEER_CHECK_ARG(pred && target && uncert && err_mean && keys && n_valid && N > 0 && D >= 1, "uce_prepare: bad args");
  cudaError_t e = cudaMemsetAsync(n_valid, 0, sizeof(unsigned long long), (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_status(e, "uce_prepare memset");
  DEER_LAUNCH(uce_prepare_kernel, stream_grid1(N, 256), 256, 0, stream, pred, target, uncert, N, D, err_mean, keys,
              n_valid);
  return DEER_OK;
}

int deer_uce_select(const unsigned* keys, long long N, const long long* ranks_host, int R, float* values,
                    void* workspace, long long workspace_bytes, void* stream) {
  DEER_CHECK_ARG(keys && ranks_host && values && workspace && N > 0 && R >= 1 && R <= UCE_MAX_RANKS,
                 "uce_select: bad args");
  const size_t need = sizeof(SelState) + sizeof(unsigned long long) * UCE_MAX_RANKS * 256;
  DEER_CHECK_ARG((size_t)workspace_bytes >= need, "uce_select: workspace too small (DEER_UCE_WORKSPACE_BYTES)");
  SelState init;
  for (int r = 0; r < UCE_MAX_RANKS; r++) {
    init.prefix[r] = 0u;
    init.k[r] = r < R ? (unsigned long long)ranks_host[r] : 0ull;
    if (r < R) DEER_CHECK_ARG(ranks_host[r] >= 0 && ranks_host[r] < N, "uce_select: rank out of range");
  }
  SelState* st = reinterpret_cast<SelState*>(workspace);
  unsigned long long* hist = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(workspace) + sizeof(SelState));
  cudaStream_t s = (cudaStream_t)stream;
  DEER_LAUNCH(uce_select_init_kernel, 1, 32, 0, stream, init, st);
  cudaError_t e = cudaMemsetAsync(hist, 0, sizeof(unsigned long long) * UCE_MAX_RANKS * 256, s);
  if (e != cudaSuccess) return cuda_status(e, "uce_select memset");
  for (int pass = 0; pass < 4; pass++) {
    DEER_LAUNCH(uce_select_hist_kernel, stream_grid1(N, 256), 256, 0, stream, keys, N, pass, R, st, hist);
    DEER_LAUNCH(uce_select_scan_kernel, 1, 32, 0, stream, pass, R, st, hist, values);
  }
  return DEER_OK;
}

int deer_uce_bins(const unsigned* keys, const float* err_mean, long long N, int n_bins, const double* edges,
                  double* bin_sums, void* stream) {
  DEER_CHECK_ARG(keys && err_mean && edges && bin_sums && N > 0 && n_bins >= 1 && n_bins <= DEER_UCE_MAX_BINS,
                 "uce_bins: bad args");
  cudaError_t e = cudaMemsetAsync(bin_sums, 0, sizeof(double) * 3 * n_bins, (cudaStream_t)stream);
  if (e != cudaSuccess) return cuda_status(e, "uce_bins memset");
  DEER_LAUNCH(uce_bins_kernel, stream_grid1(N, 256), 256, 0, stream, keys, err_mean, N, n_bins, edges, bin_sums);
  return DEER_OK;
}

}  // extern "C"
