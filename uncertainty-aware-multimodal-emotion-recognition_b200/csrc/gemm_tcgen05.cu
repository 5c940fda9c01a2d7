// placeholder until the tcgen05/TMA engine lands: reports "unsupported" so DEER_GEMM_AUTO uses the SIMT engine.
#include "common.cuh"
namespace deer {
bool gemm_tcgen05_supported(const float*, long long, int, const float*, long long, int, const float*, long long, int,
                            int, int, int, long long, long long, long long) {
  return false;
}
int gemm_tcgen05(const float*, long long, int, const float*, long long, int, float*, long long, int, int, int,
                 const float*, int, float, int, long long, long long, long long, long long, cudaStream_t) {
  return DEER_ERR_UNSUPPORTED;
}
}  // namespace deer
