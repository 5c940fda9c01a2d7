// tcgen05 / TMEM / TMA GEMM engine (sm_100a).
//
//   C = act(opA(A) opB(B) + bias + beta*C)       fp32 in HBM, TF32 tensor-core math, fp32 accumulation in TMEM
//
// Every dense contraction of the DEER path that is big enough to matter runs here: the time-batched LSTM input
// projections ([B*T,84|512] x [.,2048]), the attention-pooling scorers, the Conv1d taps, every nn.Linear, and their
// dgrad / wgrad transposes.  Operands stay fp32 in HBM (the tolerance of the path is 1e-3 relative, which BF16
// operands do not meet after ~15 chained layers; TF32 keeps 10 mantissa bits) and are fed to `tcgen05.mma
// kind::tf32` straight from 128B-swizzled shared memory filled by TMA:
//
//   warp 0      TMA producer   : cp.async.bulk.tensor.2d -> smem ring (STAGES x (A 16 KB + B 16 KB)), mbarrier tx
//   warp 1      MMA issuer     : one elected thread, 4 x tcgen05.mma (M128 N128 K8) per 32-wide K block,
//                                tcgen05.commit releases the smem slot / signals the epilogue; owns TMEM alloc
//   warps 2..5  epilogue       : tcgen05.ld 32x32b.x32 (one accumulator row per thread), bias / activation /
//                                beta*C, 128-bit stores (or red.add for split-K)
//
// Both operand majors are supported without any transposition pass: a K-major operand (row = M or N index, K
// contiguous) is one TMA box [128 rows x 32 k]; an MN-major operand (row = k, M or N contiguous; this is what the
// weight-gradient GEMM dY^T X and the input-gradient GEMM dY W need) is four boxes [32 k x 32 mn] and the UMMA
// shared-memory descriptor / instruction descriptor carry the major bits.  Tails in M, N, K are handled by TMA
// out-of-bounds zero fill plus masked stores.  Two CTAs co-reside per SM (96 KB smem, 128 TMEM columns each) so one
// CTA's prologue/epilogue overlaps the other's main loop.  Split-K (grid.z) covers weight gradients whose reduction
// runs over B*T = 76800 rows but whose output has fewer tiles than SMs.
#include "tc_ptx.cuh"

namespace deer {
namespace tc {

constexpr int BLOCK_M = 128, BLOCK_N = 128, BLOCK_K = 32;  // 32 fp32 = one 128-byte swizzle row
constexpr int UMMA_K = 8;                                  // tf32: 32 bytes of K per instruction
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BLOCK_M * BLOCK_K * 4;          // 16 KB, same for A and B
constexpr int STAGE_BYTES = 2 * TILE_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int NUM_THREADS = 192;
constexpr int TMEM_COLS = 128;

struct Params {
  float* C;
  long long ldc;
  const float* bias;
  int M, N, K;
  int act;
  float beta;
  int splits;
  int kblocks_per_split;
};

// A_MN / B_MN: operand is MN-major (stored [K, M] resp. [K, N]).
template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 2)
    gemm_tf32_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const Params p) {
  DEER_PDL_ENTRY();
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment required by the 128B swizzle atoms
  // 1024-byte alignment for the 128B swizzle atoms, applied as an OFFSET on the __shared__ array: going through
  // uintptr_t would make every later access a generic LD/ST instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.y * BLOCK_M, n0 = blockIdx.x * BLOCK_N;
  const int total_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int kb_begin = blockIdx.z * p.kblocks_per_split;
  const int kb_end = min(total_kb, kb_begin + p.kblocks_per_split);
  const int num_kb = kb_end - kb_begin;  // host guarantees >= 1

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(tmem_full_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================================================================== TMA producer
    const bool leader = elect_one();
    for (int i = 0; i < num_kb; i++) {
      const int s = i % STAGES;
      const uint32_t ph = (i / STAGES) & 1;
      mbar_wait(&empty_bar[s], ph ^ 1);
      uint8_t* sa = smem + s * STAGE_BYTES;
      uint8_t* sb = sa + TILE_BYTES;
      const int k0 = (kb_begin + i) * BLOCK_K;
      if (leader) {
        mbar_expect_tx(&full_bar[s], STAGE_BYTES);
        if (!A_MN) {
          tma_load_2d(sa, &tmap_a, &full_bar[s], k0, m0);  // box {32 k, 128 rows}
        } else {
#pragma unroll
          for (int j = 0; j < BLOCK_M / 32; j++)            // box {32 m, 32 k} x4
            tma_load_2d(sa + j * 4096, &tmap_a, &full_bar[s], m0 + 32 * j, k0);
        }
        if (!B_MN) {
          tma_load_2d(sb, &tmap_b, &full_bar[s], k0, n0);
        } else {
#pragma unroll
          for (int j = 0; j < BLOCK_N / 32; j++)
            tma_load_2d(sb + j * 4096, &tmap_b, &full_bar[s], n0 + 32 * j, k0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    // whole warp runs the loop (warp-uniform control flow keeps the tcgen05 operands in uniform registers: a lane-0
    // branch makes ptxas wrap every MMA in an ELECT/R2UR waterfall, ~60 cycles per instruction); one lane issues
    {
      constexpr uint32_t idesc = make_idesc(A_MN ? 1 : 0, B_MN ? 1 : 0, BLOCK_M, BLOCK_N);
      const uint32_t tb = warp_uniform(tmem_base);
      const bool leader = elect_one();
      for (int i = 0; i < num_kb; i++) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sb = sa + TILE_BYTES;
        if (leader) {
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; k++) {
            // K-major : 8-row x 128 B swizzle atoms stacked every 1024 B (SBO); +32 B per K step inside the atom.
            // MN-major: rows are k, 128 B = 32 mn elements; 32-element MN chunks every 4096 B (LBO); the 8 k-rows of
            //           one instruction are two 4-row (512 B) swizzle atoms (SBO); +1024 B per K step.
            const uint64_t ad = A_MN ? make_smem_desc(sa + k * 1024, 4096, 512, 1) : make_smem_desc(sa + k * 32, 16, 1024, 2);
            const uint64_t bd = B_MN ? make_smem_desc(sb + k * 1024, 4096, 512, 1) : make_smem_desc(sb + k * 32, 16, 1024, 2);
            umma_tf32(tb, ad, bd, idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // smem slot reusable once these MMAs have read it
        }
        __syncwarp();
      }
      if (leader) umma_commit(tmem_full_bar);  // accumulator complete
      __syncwarp();
    }
  } else {
    // ===================================================================== epilogue (warps 2..5)
    // TMEM holds one accumulator ROW per thread; storing it directly would make every warp store touch 32 rows
    // (32 half-used sectors per request: measured 32 sectors/request, the kernel was L1-wavefront bound).  Each warp
    // instead transposes its 32x32 chunk through a padded shared-memory tile (the operand ring is idle once
    // tmem_full fires) and writes 4 complete 128-byte row segments per store instruction.
    const int g = warp & 3;          // TMEM lane group this warp may access
    const int row = g * 32 + lane;
    const int gm = m0 + row;
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    const bool row_ok = gm < p.M;
    float* crow = p.C + (long long)gm * p.ldc;
    constexpr int TLD = 36;          // padded tile row (floats): conflict-free 128-bit writes by row and reads by column
    float* tile = reinterpret_cast<float*>(smem) + g * (32 * TLD);
    const int cq = (lane & 7) * 4;   // column quad of this lane in the read-back phase
    const int rsub = lane >> 3;      // row within each group of 4 rows
#pragma unroll 1
    for (int c = 0; c < BLOCK_N / 32; c++) {
      const int gn0 = n0 + c * 32;
      if (gn0 >= p.N) break;  // warp-uniform
      const bool full = gn0 + 32 <= p.N;
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(c * 32), v);
      tmem_ld_wait();
      if (p.splits > 1) {
        if (!row_ok) continue;
        // split-K partial sums: accumulate semantics (beta == 1, no activation); split 0 carries the bias
#pragma unroll
        for (int j = 0; j < 32; j++) {
          if (gn0 + j < p.N) {
            float x = v[j];
            if (p.bias && blockIdx.z == 0) x += __ldg(p.bias + gn0 + j);
            atomicAdd(crow + gn0 + j, x);
          }
        }
      } else if (full) {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(tile + lane * TLD + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncwarp();
        float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.bias) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gn0 + cq));
#pragma unroll
        for (int it = 0; it < 8; it++) {
          const int r = it * 4 + rsub;
          const int grow = m0 + g * 32 + r;
          if (grow < p.M) {
            float4 o = *reinterpret_cast<const float4*>(tile + r * TLD + cq);
            float4* dst = reinterpret_cast<float4*>(p.C + (long long)grow * p.ldc + gn0 + cq);
            o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
            if (p.beta != 0.f) {
              const float4 old = __ldcg(dst);
              o.x += p.beta * old.x; o.y += p.beta * old.y; o.z += p.beta * old.z; o.w += p.beta * old.w;
            }
            if (p.act != DEER_ACT_NONE) o = act_apply4(o, p.act);
            *dst = o;
          }
        }
        __syncwarp();
      } else {
        if (!row_ok) continue;
#pragma unroll
        for (int j = 0; j < 32; j++) {
          if (gn0 + j < p.N) {
            float x = v[j];
            if (p.bias) x += __ldg(p.bias + gn0 + j);
            if (p.beta != 0.f) x += p.beta * crow[gn0 + j];
            crow[gn0 + j] = p.act != DEER_ACT_NONE ? act_apply1(x, p.act) : x;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      ptr = nullptr;
    return reinterpret_cast<EncodeTiledFn>(ptr);
  }();
  return fn;
}

static int g_tma_tf32_round = 1;  // 1: CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 (TMA rounds fp32 -> tf32), 0: raw FLOAT32 bits

// 2-D map over a row-major fp32 matrix with `rows` rows of `cols` contiguous elements, row pitch `ld` elements.
bool make_map(CUtensorMap* map, const float* base, long long rows, long long cols, long long ld, int box_cols,
              int box_rows, bool mn_major) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, g_tma_tf32_round ? CU_TENSOR_MAP_DATA_TYPE_TFLOAT32 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                  const_cast<float*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace tc

void gemm_tcgen05_set_round(int on) { tc::g_tma_tf32_round = on ? 1 : 0; }

bool gemm_tcgen05_supported(const float* A, long long lda, int transA, const float* B, long long ldb, int transB,
                            const float* C, long long ldc, int M, int N, int K, int batch, long long, long long,
                            long long) {
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  if (batch != 1) return false;
  if (!al16(A) || !al16(B) || !al16(C)) return false;
  if ((lda & 3) || (ldb & 3) || (ldc & 3)) return false;
  if (M < 32 || N < 32 || K < 32) return false;  // tiny problems: the fp32 SIMT engine is exact and as fast
  (void)transA;
  (void)transB;
  return true;
}

int gemm_tcgen05(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                 long long ldc, int M, int N, int K, const float* bias, int act, float beta, int batch, long long,
                 long long, long long, long long, cudaStream_t stream) {
  using namespace tc;
  if (batch != 1) return DEER_ERR_UNSUPPORTED;
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return DEER_ERR_UNSUPPORTED;
  CUtensorMap ma, mb;
  bool ok;
  if (!transA) ok = make_map(&ma, A, M, K, lda, BLOCK_K, BLOCK_M, false);  // stored [M,K]: K-major
  else ok = make_map(&ma, A, K, M, lda, 32, BLOCK_K, true);                // stored [K,M]: MN-major
  if (transB) ok = ok && make_map(&mb, B, N, K, ldb, BLOCK_K, BLOCK_N, false);  // stored [N,K]: K-major
  else ok = ok && make_map(&mb, B, K, N, ldb, 32, BLOCK_K, true);               // stored [K,N]: MN-major
  if (!ok) {
    set_error("gemm_tcgen05: cuTensorMapEncodeTiled failed (M=%d N=%d K=%d lda=%lld ldb=%lld)", M, N, K, lda, ldb);
    return DEER_ERR_UNSUPPORTED;
  }
  const int gx = (N + BLOCK_N - 1) / BLOCK_N, gy = (M + BLOCK_M - 1) / BLOCK_M;
  const int total_kb = (K + BLOCK_K - 1) / BLOCK_K;
  int splits = 1;
  if (beta == 1.f && act == DEER_ACT_NONE) {
    const int tiles = gx * gy;
    if (tiles < kNumSMs && total_kb >= 64) {
      splits = (2 * kNumSMs + tiles - 1) / tiles;
      const int max_splits = total_kb / 16;
      if (splits > max_splits) splits = max_splits;
      if (splits < 1) splits = 1;
    }
  }
  int per = (total_kb + splits - 1) / splits;
  splits = (total_kb + per - 1) / per;  // no empty split
  Params p{C, ldc, bias, M, N, K, act, beta, splits, per};
  dim3 grid(gx, gy, splits);
#define DEER_TC_GO(AM, BM)                                                                                      \
  do {                                                                                                          \
    static bool attr = false;                                                                                   \
    if (!attr) {                                                                                                \
      cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<AM, BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           SMEM_BYTES);                                                         \
      if (e != cudaSuccess) return cuda_status(e, "gemm_tcgen05 smem attribute");                               \
      attr = true;                                                                                              \
    }                                                                                                           \
    DEER_LAUNCH((gemm_tf32_kernel<AM, BM>), grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, p);                  \
  } while (0)
  if (!transA && transB) DEER_TC_GO(false, false);
  else if (!transA && !transB) DEER_TC_GO(false, true);
  else if (transA && transB) DEER_TC_GO(true, false);
  else DEER_TC_GO(true, true);
#undef DEER_TC_GO
  return DEER_OK;
}

}  // namespace deer
