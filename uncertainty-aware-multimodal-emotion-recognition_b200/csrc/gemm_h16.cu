// Persistent tcgen05 GEMM engine for 16-bit operands (both FP16 or both BF16), FP32 accumulation.
//
//   C = act(opA(A) opB(B) + bias + beta*C)     C fp32 in HBM (+ optional 16-bit shadow copy C16 for the next GEMM)
//
// Used for the time-batched LSTM contractions (input projections, dx, dW_ih, dW_hh), whose operands exist as 16-bit
// shadows written by their producers (the persistent LSTM kernels emit h as FP16 and dpre as BF16; weights and the
// network input are cast once per step).  For this model's value ranges FP16 carries the same 11 significant bits as
// TF32; gradients use BF16 for the exponent range.  Versus the TF32 engine: half the operand bytes through L2 (the
// TF32 engine is L2-bandwidth bound at 128 B/cycle/SM) and twice the tensor rate.
//
//   tile 128 x 256, K block 64 (one 128-byte swizzle row), 4-stage TMA ring (4 x 48 KB), M128 N256 K16 MMAs
//   persistent CTAs (one per SM) over (tile, K-split) work items; TWO 256-column TMEM accumulators so the epilogue of
//   work item i overlaps the main loop of item i+1
//   warp 0  TMA producer     warp 1  MMA issuer (owns TMEM)     warps 2..9  epilogue (2 per TMEM lane group)
//   epilogue: tcgen05.ld -> shared-memory transpose -> bias / beta*C / activation -> 128-byte coalesced row segments
//   (fp32 and, optionally, 16-bit); split-K partials go out as fp32 atomics (beta == 1, no activation)
// Both operand majors: K-major (row = M/N index, K contiguous: one box per stage) and MN-major (row = k, M/N contiguous:
// 64-element boxes; what dY^T X and dY W need) -- no transposition pass anywhere.
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

#include "tc_ptx.cuh"

namespace deer {
namespace h16 {

using namespace deer::tc;

constexpr int BM = 128, BN = 256, BK = 64, UK = 16, STAGES = 4;
constexpr int A_BYTES = BM * BK * 2;          // 16 KB
constexpr int B_BYTES = BN * BK * 2;          // 32 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int STAGING_BYTES = 8 * 4096;     // one 32x32 fp32 tile (128B-swizzled, TMA-store source) per epilogue warp
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGING_BYTES + 256 + 1024;
constexpr int NUM_THREADS = 320;            // TMA warp, MMA warp, 8 epilogue warps

struct Params {
  float* C;
  long long ldc;
  void* C16;            // optional 16-bit copy of C (row pitch ldc16 elements)
  long long ldc16;
  int c16_bf;           // 1: bf16, 0: fp16
  const float* bias;
  int M, N, K;
  int act;
  float beta;
  int splits, kb_per_split, tiles_m, tiles_n;
  uint32_t idesc;
  long long* prof;      // optional per-role wait/busy cycle counters of block 0 (debug), else null
  int debug;            // bit 0: skip the fp32 stores (timing experiments only)
  int kb_phase;         // split-precision mode (pair kernel): k-blocks of ONE pass over K; the k loop makes three passes
                        // (A_hi,B_hi), (A_lo,B_hi), (A_hi,B_lo) over the hi / lo operand maps.  0: plain GEMM
  const uint32_t* drop_mask;   // optional (pair kernel, fp32 output, beta = 0, N % 32 == 0): keep bits of an inverted dropout
  float drop_scale;            // over C, one word per (row, 32 columns): C = keep ? C * drop_scale : 0 in the epilogue
  // optional rank-1-per-sample row term (pair kernel, fp32 output, beta = 0, N % 4 == 0): C[m, :] += rt_w[b, t] * rt_v[b, :]
  // with row m = (b, t) of a [rt_B, rt_T] (batch-major) or [rt_T, rt_B] (rt_tm: time-major) grid -- the attention pooling's
  // own input gradient (weights x upstream gradient) added while the scorer's input-gradient GEMM writes dx
  const float* rt_w;           // [rt_B, rt_T]
  const float* rt_v;           // [rt_B, N]
  int rt_B, rt_T, rt_tm;
};
#define H16_T0() const long long _t0 = p.prof ? clock64() : 0
#define H16_ACC(slot)                                                   \
  do {                                                                  \
    if (p.prof && blockIdx.x == 0 && lane == 0) p.prof[slot] += clock64() - _t0; \
  } while (0)

__device__ __forceinline__ void umma_16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
    gemm_h16_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const __grid_constant__ CUtensorMap tmap_c, const Params p) {
  DEER_PDL_ENTRY();
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms, applied as an OFFSET on the __shared__ array: going through
  // uintptr_t would make every later access a generic LD/ST instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* staging = reinterpret_cast<float*>(smem + STAGES * STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + STAGING_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total_kb = (p.K + BK - 1) / BK;
  const int total_work = p.tiles_m * p.tiles_n * p.splits;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_c);
    for (int s = 0; s < STAGES; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int w, int& m0, int& n0, int& kb0, int& nkb) {
    const int split = w % p.splits, tile = w / p.splits;
    m0 = (tile / p.tiles_n) * BM;
    n0 = (tile % p.tiles_n) * BN;
    kb0 = split * p.kb_per_split;
    nkb = min(total_kb, kb0 + p.kb_per_split) - kb0;
  };

  if (warp == 0) {
    // ===================================================================== TMA producer
    const bool leader = elect_one();
    uint32_t g = 0;  // running stage counter across work items
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      int m0, n0, kb0, nkb;
      decode(w, m0, n0, kb0, nkb);
      for (int i = 0; i < nkb; i++, g++) {
        const int s = g % STAGES;
        {
          H16_T0();
          mbar_wait(&empty_bar[s], ((g / STAGES) & 1) ^ 1);
          H16_ACC(0);
        }
        if (leader) {
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          const int k0 = (kb0 + i) * BK;
          mbar_expect_tx(&full_bar[s], STAGE_BYTES);
          if (!A_MN) {
            tma_load_2d(sa, &tmap_a, &full_bar[s], k0, m0);            // box {64 k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BM / 64; j++)                           // box {64 m, 64 k} x2
              tma_load_2d(sa + j * 8192, &tmap_a, &full_bar[s], m0 + 64 * j, k0);
          }
          if (!B_MN) {
            tma_load_2d(sb, &tmap_b, &full_bar[s], k0, n0);            // box {64 k, 256 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BN / 64; j++)                           // box {64 n, 64 k} x4
              tma_load_2d(sb + j * 8192, &tmap_b, &full_bar[s], n0 + 64 * j, k0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer
    const uint32_t tb = warp_uniform(tmem_base);
    const bool leader = elect_one();
    uint32_t g = 0;
    int it = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, it++) {
      int m0, n0, kb0, nkb;
      decode(w, m0, n0, kb0, nkb);
      const int acc = it & 1;
      {
        H16_T0();
        mbar_wait(&tmem_empty[acc], (uint32_t)(((it >> 1) & 1) ^ 1));   // epilogue drained this accumulator
        H16_ACC(1);
      }
      tc_fence_after();
      for (int i = 0; i < nkb; i++, g++) {
        const int s = g % STAGES;
        {
          H16_T0();
          mbar_wait(&full_bar[s], (g / STAGES) & 1);
          H16_ACC(2);
        }
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + s * STAGE_BYTES);
        const uint32_t sb = sa + A_BYTES;
        if (leader) {
#pragma unroll
          for (int k = 0; k < BK / UK; k++) {
            // K-major : 8-row x 128 B swizzle atoms every 1024 B (SBO); +32 B per K step inside the atom
            // MN-major: rows are k (128 B = 64 mn elements); 64-element mn chunks every 8192 B (LBO); 8-k-row atoms
            //           every 1024 B (SBO); +2048 B (16 k rows) per K step
            const uint64_t ad = A_MN ? make_smem_desc(sa + k * 2048, 8192, 1024, 2) : make_smem_desc(sa + k * 32, 16, 1024, 2);
            const uint64_t bd = B_MN ? make_smem_desc(sb + k * 2048, 8192, 1024, 2) : make_smem_desc(sb + k * 32, 16, 1024, 2);
            umma_16(tb + acc * BN, ad, bd, p.idesc, (i > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (i == nkb - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..9)
    // Two warps per TMEM lane group, each taking every other 32-column chunk: tcgen05.ld (one accumulator row per
    // lane) -> bias / activation in registers -> 128B-swizzled shared tile -> ONE TMA tensor store (or TMA reduce-add
    // for beta == 1 and for split-K partial sums) per 32x32 chunk.  No per-element global store instructions, no
    // atomics, and tile tails are clipped by the TMA unit.  (Measured: the previous LDS/STG read-back loop was
    // instruction-issue bound at ~0.35 instructions per output float and cost as much as the whole main loop.)
    const int gq = warp & 3;                 // TMEM lane group of this warp
    const int half = (warp - 2) >> 2;        // which of the two warps of the group
    uint8_t* tile = reinterpret_cast<uint8_t*>(staging) + (warp - 2) * 4096;   // [32 rows][128 B], SWIZZLE_128B
    const int act = p.act;
    const bool reduce = p.beta != 0.f || p.splits > 1;
    int it = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x, it++) {
      int m0, n0, kb0, nkb;
      decode(w, m0, n0, kb0, nkb);
      const int acc = it & 1;
      {
        H16_T0();
        mbar_wait(&tmem_full[acc], (uint32_t)((it >> 1) & 1));
        if (warp == 2) H16_ACC(3);
      }
      tc_fence_after();
      H16_T0();
      const int row_base = m0 + gq * 32;
      const bool add_bias = p.bias != nullptr && (w % p.splits) == 0;
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        const int gn0 = n0 + c * 32;
        if (gn0 >= p.N || row_base >= p.M) break;  // warp-uniform: nothing of this chunk exists
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(gq * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
        tmem_ld_wait();
        if (add_bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (gn0 + j < p.N) {   // bias is 16-byte aligned and N is a multiple of 4 (checked on the host)
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gn0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
        }
        if (act != DEER_ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 o = act_apply4(make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]), act);
            v[j] = o.x; v[j + 1] = o.y; v[j + 2] = o.z; v[j + 3] = o.w;
          }
        }
        // the previous TMA store of this warp has finished READING the tile (waited below before the loop continues)
#pragma unroll
        for (int j = 0; j < 8; j++)
          *reinterpret_cast<float4*>(tile + lane * 128 + ((j ^ (lane & 7)) << 4)) =
              make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (reduce) tma_reduce_add_2d(&tmap_c, tile, gn0, row_base);
          else tma_store_2d(&tmap_c, tile, gn0, row_base);
          tma_store_commit();
          tma_store_wait_read();
        }
        __syncwarp();
      }
      // every TMEM read of this accumulator has completed (tcgen05.wait::ld): hand it back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      if (warp == 2) H16_ACC(4);
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------- CTA-pair engine
// cta_group::2 variant: a cluster of two CTAs (one TPC) owns a 256 x 256 output tile.  Each CTA stages its own 128
// rows of A and HALF of the B tile (128 of the 256 N rows); the leader CTA's single thread issues M256 N256 K16 MMAs
// that read both CTAs' shared memory and write 128 accumulator rows into each CTA's TMEM.  Per CTA and k-block the
// shared-memory traffic drops from 48 KB written + 48 KB read to 32 + 32 KB: the single-CTA kernel is bound by exactly
// that traffic (TMA fill + tensor-core operand reads against 128 B/clk/SM), not by the tensor pipe.
//   full barriers live in the leader (both CTAs' TMA loads signal them, .cta_group::2); the MMA commits multicast to the
//   slot-empty and accumulator-full barriers of BOTH CTAs; both CTAs' epilogue warps arrive on the leader's
//   accumulator-empty barrier.  Everything else (roles, TMEM double buffering, TMA-store epilogue) is as above.
constexpr int P_STAGES = 5;
constexpr int P_STAGING_BYTES = 2 * STAGING_BYTES;   // TWO 32x32 fp32 store tiles per epilogue warp (double buffered)
constexpr int PB_BYTES = (BN / 2) * BK * 2;            // 16 KB: this CTA's half of the B tile
constexpr int P_STAGE_BYTES = A_BYTES + PB_BYTES;      // 32 KB
constexpr int P_SMEM_BYTES = P_STAGES * P_STAGE_BYTES + P_STAGING_BYTES + 256 + 1024;
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;        // shared::cluster address of the same offset in the even CTA

template <int ELEM>
__device__ __forceinline__ void umma_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  if (ELEM == 2)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"((uint16_t)3)
      : "memory");
}
// TMA tile load whose completion bytes are counted on the LEADER CTA's mbarrier (same offset, peer bit cleared)
__device__ __forceinline__ void tma_load_2d_pair(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar) & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive on the barrier at the same offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(bar)), "r"(rank));
  // default semantics (release at CTA scope): the .release.cluster form compiles to MEMBAR.ALL.GPU + ERRBAR, which
  // waited on the warp's outstanding global traffic (11 % of the stall samples); the hand-back only orders TMEM reads,
  // and those are complete (tcgen05.wait::ld + tcgen05.fence::before_thread_sync) before this arrive
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}

// 16 TMEM lanes x 32 columns in the mma-accumulator fragment layout: lane t holds row t/4 (registers 4n, 4n+1) and row
// t/4 + 8 (registers 4n+2, 4n+3) at columns 8n + 2*(t%4) + {0,1}, n = 0..3.  Four lanes therefore own 32 contiguous
// bytes of one output row: a warp-wide 8-byte store writes eight full 32-byte sectors, straight from registers.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void red_add_v2(float* p, float a, float b) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(a), "f"(b) : "memory");
}

// ELEM = operand element size: 2 (FP16 / BF16, kind::f16, 64-element k-blocks) or 4 (fp32 read as TF32, kind::tf32,
// 32-element k-blocks).  In BYTES both look the same to the pipeline (128-byte swizzle rows, 32 bytes of K per MMA);
// only the MN-major box geometry differs (TF32 needs the 32-byte-atom swizzle: 32x32 boxes instead of 64x64).
template <int ELEM, bool A_MN, bool B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
    gemm_h16_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_c, const __grid_constant__ CUtensorMap tmap_a2,
                         const __grid_constant__ CUtensorMap tmap_b2, const Params p) {
  DEER_PDL_ENTRY();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  float* staging = reinterpret_cast<float*>(smem + P_STAGES * P_STAGE_BYTES);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + P_STAGES * P_STAGE_BYTES + P_STAGING_BYTES);
  uint64_t* empty_bar = full_bar + P_STAGES;
  uint64_t* tmem_full = empty_bar + P_STAGES;  // [2]
  uint64_t* tmem_empty = tmem_full + 2;        // [2]  (the leader's copy is the live one)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
  constexpr int BKE = 128 / ELEM;                  // k elements per block (one 128-byte swizzle row)
  constexpr int MNB = BKE;                         // MN-major box: MNB mn-elements x BKE k-rows
  constexpr int MNB_BYTES = MNB * 128;             // 8 KB (16-bit) / 4 KB (TF32)
  constexpr int MN_KSTEP = (ELEM == 2 ? 16 : 8) * 128;   // bytes per MMA K step in an MN-major panel
  constexpr int MN_SBO = ELEM == 2 ? 1024 : 512;
  constexpr int MN_LAYOUT = ELEM == 2 ? 2 : 1;
  // split precision (p.kb_phase > 0): every operand is a pair of FP16 matrices x = hi + lo (hi = fp16(x), lo =
  // fp16(x - hi): 22 significant bits); the k loop runs three passes over K accumulating A_hi B_hi + A_lo B_hi +
  // A_hi B_lo in the fp32 TMEM accumulator -- fp32-grade products at 3 MMAs on the 16-bit tensor rate
  const int total_kb = p.kb_phase > 0 ? 3 * p.kb_phase : (p.K + BKE - 1) / BKE;
  const int total_work = p.tiles_m * p.tiles_n * p.splits;   // tiles_m counts 256-row pair tiles here

  if (threadIdx.x == 0) {
    prefetch_tmap(&tmap_a);
    prefetch_tmap(&tmap_b);
    prefetch_tmap(&tmap_c);
    if (p.kb_phase > 0) {
      prefetch_tmap(&tmap_a2);
      prefetch_tmap(&tmap_b2);
    }
    for (int s = 0; s < P_STAGES; s++) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int b = 0; b < 2; b++) {
      mbar_init(&tmem_full[b], 1);
      mbar_init(&tmem_empty[b], 16);   // 8 epilogue warps of each CTA
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // both CTAs' barriers exist before any remote arrival / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode = [&](int w, int& m0, int& n0, int& kb0, int& nkb) {
    const int split = w % p.splits, tile = w / p.splits;
    m0 = (tile / p.tiles_n) * (2 * BM);
    n0 = (tile % p.tiles_n) * BN;
    kb0 = split * p.kb_per_split;
    nkb = min(total_kb, kb0 + p.kb_per_split) - kb0;
  };

  if (warp == 0) {
    // ===================================================================== TMA producer (both CTAs)
    const bool leader = elect_one();
    uint32_t g = 0;
    for (int w = cluster_id; w < total_work; w += num_clusters) {
      int m0, n0, kb0, nkb;
      decode(w, m0, n0, kb0, nkb);
      const int mr = m0 + (int)rank * BM;          // this CTA's 128 rows of A
      const int nr = n0 + (int)rank * (BN / 2);    // this CTA's 128 rows of B
      for (int i = 0; i < nkb; i++, g++) {
        const int s = g % P_STAGES;
        mbar_wait(&empty_bar[s], ((g / P_STAGES) & 1) ^ 1);
        if (leader) {
          uint8_t* sa = smem + s * P_STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          int kbi = kb0 + i;
          const CUtensorMap* ta = &tmap_a;
          const CUtensorMap* tb = &tmap_b;
          if (p.kb_phase > 0) {                     // pass 0: (hi, hi), pass 1: (lo, hi), pass 2: (hi, lo)
            const int pass = kbi / p.kb_phase;
            kbi -= pass * p.kb_phase;
            if (pass == 1) ta = &tmap_a2;
            if (pass == 2) tb = &tmap_b2;
          }
          const int k0 = kbi * BKE;
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * P_STAGE_BYTES);   // both CTAs' bytes land on this barrier
          if (!A_MN) {
            tma_load_2d_pair(sa, ta, &full_bar[s], k0, mr);                 // box {BKE k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < BM / MNB; j++)                               // boxes {MNB m, BKE k}
              tma_load_2d_pair(sa + j * MNB_BYTES, ta, &full_bar[s], mr + MNB * j, k0);
          }
          if (!B_MN) {
            tma_load_2d_pair(sb, tb, &full_bar[s], k0, nr);                 // box {BKE k, 128 rows}
          } else {
#pragma unroll
            for (int j = 0; j < (BN / 2) / MNB; j++)                         // boxes {MNB n, BKE k}
              tma_load_2d_pair(sb + j * MNB_BYTES, tb, &full_bar[s], nr + MNB * j, k0);
          }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (leader CTA only)
    if (rank == 0) {
      const uint32_t tb = warp_uniform(tmem_base);
      const bool leader = elect_one();
      uint32_t g = 0;
      int it = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters, it++) {
        int m0, n0, kb0, nkb;
        decode(w, m0, n0, kb0, nkb);
        const int acc = it & 1;
        mbar_wait(&tmem_empty[acc], (uint32_t)(((it >> 1) & 1) ^ 1));   // both CTAs drained this accumulator
        tc_fence_after();
        for (int i = 0; i < nkb; i++, g++) {
          const int s = g % P_STAGES;
          mbar_wait(&full_bar[s], (g / P_STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * P_STAGE_BYTES);
          const uint32_t sb = sa + A_BYTES;
          if (leader) {
#pragma unroll
            for (int k = 0; k < BK / UK; k++) {
              const uint64_t ad = A_MN ? make_smem_desc(sa + k * MN_KSTEP, MNB_BYTES, MN_SBO, MN_LAYOUT)
                                       : make_smem_desc(sa + k * 32, 16, 1024, 2);
              const uint64_t bd = B_MN ? make_smem_desc(sb + k * MN_KSTEP, MNB_BYTES, MN_SBO, MN_LAYOUT)
                                       : make_smem_desc(sb + k * 32, 16, 1024, 2);
              umma_pair<ELEM>(tb + acc * BN, ad, bd, p.idesc, (i > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit_pair(&empty_bar[s]);
            if (i == nkb - 1) umma_commit_pair(&tmem_full[acc]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================================================================== epilogue (warps 2..9, both CTAs)
    const int gq = warp & 3;
    const int half = (warp - 2) >> 2;
    uint8_t* tile0 = reinterpret_cast<uint8_t*>(staging) + (warp - 2) * 8192;   // two 4 KB tiles per warp
    uint32_t nstore = 0;                                                         // stores issued by this warp
    const int act = p.act;
    const bool reduce = p.beta != 0.f || p.splits > 1;
    int it = 0;
    for (int w = cluster_id; w < total_work; w += num_clusters, it++) {
      int m0, n0, kb0, nkb;
      decode(w, m0, n0, kb0, nkb);
      const int acc = it & 1;
      mbar_wait(&tmem_full[acc], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const int row_base = m0 + (int)rank * BM + gq * 32;
      const bool add_bias = p.bias != nullptr && (w % p.splits) == 0;
      if (p.debug & 4) {
        // ---- direct epilogue (ablation, DEER_H16_DEBUG=4; measured SLOWER than the TMA-store epilogue: 110 vs 100 us): TMEM -> registers (accumulator-fragment layout) -> full-sector global stores.  No
        // shared-memory staging and no TMA store: the TMA unit and the shared-memory port stay with the operand ring
        // (measured: the TMA-store epilogue added 37 us to the 64 us the rest of the kernel needs at K = 512).
        const int lr = lane >> 2, lc = (lane & 3) * 2;
#pragma unroll 1
        for (int c = half; c < BN / 32; c += 2) {
          const int gn0 = n0 + c * 32;
          if (gn0 >= p.N || row_base >= p.M) break;
          float v[32];
          const uint32_t ta = tmem_base + ((uint32_t)(gq * 32) << 16) + (uint32_t)(acc * BN + c * 32);
          tmem_ld_16x256b_x4(ta, v);
          tmem_ld_16x256b_x4(ta + (16u << 16), v + 16);
          tmem_ld_wait();
          if (p.debug & 2) continue;
#pragma unroll
          for (int n = 0; n < 4; n++) {
            const int col = gn0 + 8 * n + lc;
            if (col >= p.N) continue;   // N is even (checked on the host): col + 1 < N as well
            float2 b2 = make_float2(0.f, 0.f);
            if (add_bias) b2 = __ldg(reinterpret_cast<const float2*>(p.bias + col));
#pragma unroll
            for (int h2 = 0; h2 < 2; h2++) {
#pragma unroll
              for (int q = 0; q < 2; q++) {
                const int row = row_base + 16 * h2 + 8 * q + lr;
                float x = v[16 * h2 + 4 * n + 2 * q] + b2.x, y = v[16 * h2 + 4 * n + 2 * q + 1] + b2.y;
                if (act != DEER_ACT_NONE) {
                  x = act_apply1(x, act);
                  y = act_apply1(y, act);
                }
                if (row < p.M && !(p.debug & 1)) {
                  float* dst = p.C + (long long)row * p.ldc + col;
                  if (reduce) red_add_v2(dst, x, y);
                  else __stcs(reinterpret_cast<float2*>(dst), make_float2(x, y));
                }
              }
            }
          }
        }
      } else if (p.C16 != nullptr) {
        // ---- 16-bit output (tmap_c describes C16): 64 columns per store -- the same 4 KB, 128-byte-row staging tile
        // carries twice the columns, so the kernel writes half the bytes (the fp32 store stream is what bounds the
        // small-K projections)
#pragma unroll 1
        for (int c = half; c < BN / 64; c += 2) {
          const int gn0 = n0 + c * 64;
          if (gn0 >= p.N || row_base >= p.M) break;
          float v[64];
          const uint32_t ta = tmem_base + ((uint32_t)(gq * 32) << 16) + (uint32_t)(acc * BN + c * 64);
          tmem_ld32(ta, v);
          tmem_ld32(ta + 32, v + 32);
          tmem_ld_wait();
          if (add_bias) {
#pragma unroll
            for (int j = 0; j < 64; j += 4) {
              if (gn0 + j < p.N) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gn0 + j));
                v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
              }
            }
          }
          if (act != DEER_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 64; j += 4) {
              const float4 o = act_apply4(make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]), act);
              v[j] = o.x; v[j + 1] = o.y; v[j + 2] = o.z; v[j + 3] = o.w;
            }
          }
          if (p.debug & 2) continue;
          uint8_t* tile = tile0 + (nstore & 1) * 4096;
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; j++) {
            uint4 o;
            if (p.c16_bf) {
              o.x = pack_bf2(v[8 * j], v[8 * j + 1]); o.y = pack_bf2(v[8 * j + 2], v[8 * j + 3]);
              o.z = pack_bf2(v[8 * j + 4], v[8 * j + 5]); o.w = pack_bf2(v[8 * j + 6], v[8 * j + 7]);
            } else {
              o.x = pack_h2(v[8 * j], v[8 * j + 1]); o.y = pack_h2(v[8 * j + 2], v[8 * j + 3]);
              o.z = pack_h2(v[8 * j + 4], v[8 * j + 5]); o.w = pack_h2(v[8 * j + 6], v[8 * j + 7]);
            }
            *reinterpret_cast<uint4*>(tile + lane * 128 + ((j ^ (lane & 7)) << 4)) = o;
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0 && !(p.debug & 1)) {
            tma_store_2d(&tmap_c, tile, gn0, row_base);
            tma_store_commit();
          }
          nstore++;
        }
      } else
#pragma unroll 1
      for (int c = half; c < BN / 32; c += 2) {
        const int gn0 = n0 + c * 32;
        if (gn0 >= p.N || row_base >= p.M) break;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(gq * 32) << 16) + (uint32_t)(acc * BN + c * 32), v);
        tmem_ld_wait();
        if (add_bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (gn0 + j < p.N) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + gn0 + j));
              v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
            }
          }
        }
        if (act != DEER_ACT_NONE) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 o = act_apply4(make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]), act);
            v[j] = o.x; v[j + 1] = o.y; v[j + 2] = o.z; v[j + 3] = o.w;
          }
        }
        if (p.rt_w) {
          const int row = row_base + lane;
          if (row < p.M) {
            const int b = p.rt_tm ? row % p.rt_B : row / p.rt_T;
            const int t = p.rt_tm ? row / p.rt_B : row % p.rt_T;
            const float w = __ldg(p.rt_w + (long long)b * p.rt_T + t);
            const float* vrow = p.rt_v + (long long)b * p.N + gn0;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              if (gn0 + j < p.N) {
                const float4 u = __ldg(reinterpret_cast<const float4*>(vrow + j));
                v[j] = fmaf(w, u.x, v[j]); v[j + 1] = fmaf(w, u.y, v[j + 1]);
                v[j + 2] = fmaf(w, u.z, v[j + 2]); v[j + 3] = fmaf(w, u.w, v[j + 3]);
              }
            }
          }
        }
        if (p.drop_mask) {
          // the backward of an inverted dropout on this GEMM's output (nn.LSTM's inter-layer dropout applied to dx): this
          // lane holds 32 consecutive columns of one row = one word of the keep mask the forward pass wrote
          const int row = row_base + lane;
          const uint32_t w = row < p.M ? __ldg(p.drop_mask + (long long)row * (p.N >> 5) + (gn0 >> 5)) : 0u;
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = ((w >> j) & 1u) ? v[j] * p.drop_scale : 0.f;
        }
        if (p.debug & 2) continue;   // timing experiments: TMEM read only
        // double-buffered store tiles: before overwriting a tile only the store issued TWO chunks ago must have been
        // read out by the TMA unit; the previous chunk's store keeps draining while this one is staged (the wait
        // right after every store was 29 % of all stall samples)
        uint8_t* tile = tile0 + (nstore & 1) * 4096;
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; j++)
          *reinterpret_cast<float4*>(tile + lane * 128 + ((j ^ (lane & 7)) << 4)) =
              make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0 && !(p.debug & 1)) {
          if (reduce) tma_reduce_add_2d(&tmap_c, tile, gn0, row_base);
          else tma_store_2d(&tmap_c, tile, gn0, row_base);
          tma_store_commit();
        }
        nstore++;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&tmem_empty[acc], 0);   // the leader's MMA warp owns the hand-back barrier
    }
    if (lane == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // the leader's MMAs read this CTA's shared memory; nobody leaves or frees TMEM early
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
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
// 2-D map over a row-major 16-bit matrix: `rows` rows of `cols` contiguous elements, pitch `ld` elements
static bool make_map16(CUtensorMap* map, const void* base, int bf, long long rows, long long cols, long long ld,
                       int box_cols, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, bf ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                  const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// 2-D map over the fp32 output: box = one 32x32 chunk, 128B swizzle (matches the epilogue's shared tile)
static bool make_map_c(CUtensorMap* map, float* base, long long rows, long long cols, long long ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

}  // namespace h16

static long long* g_h16_prof = nullptr;
int g_h16_pair = 1;   // 1 (default): cta_group::2 CTA-pair kernel when M > 128; 0: single-CTA kernel (DEER_OPT_H16_PAIR)
void gemm_h16_set_profile(long long* buf) { g_h16_prof = buf; }

int gemm_h16_ex(const void* A, const void* A_lo, long long lda, int transA, int a_bf, const void* B, const void* B_lo,
                long long ldb, int transB, int b_bf, float* C, long long ldc, void* C16, long long ldc16, int c16_bf, int M,
                int N, int K, const float* bias, int act, float beta, cudaStream_t stream);
int gemm_h16(const void* A, long long lda, int transA, int a_bf, const void* B, long long ldb, int transB, int b_bf,
             float* C, long long ldc, void* C16, long long ldc16, int c16_bf, int M, int N, int K, const float* bias,
             int act, float beta, cudaStream_t stream) {
  return gemm_h16_ex(A, nullptr, lda, transA, a_bf, B, nullptr, ldb, transB, b_bf, C, ldc, C16, ldc16, c16_bf, M, N, K,
                     bias, act, beta, stream);
}

// keep mask handed to the NEXT gemm_h16_ex call of this thread (set and cleared by gemm_h16_dropmask)
static thread_local const uint32_t* t_drop_mask = nullptr;
static thread_local float t_drop_scale = 1.f;

// A_lo / B_lo non-NULL: split-precision product (A + A_lo)(B + B_lo) ~ A B + A_lo B + A B_lo on FP16 hi / lo operand
// pairs (CTA-pair kernel only; beta = 0)
int gemm_h16_ex(const void* A, const void* A_lo, long long lda, int transA, int a_bf, const void* B, const void* B_lo,
                long long ldb, int transB, int b_bf, float* C, long long ldc, void* C16, long long ldc16, int c16_bf, int M,
                int N, int K, const float* bias, int act, float beta, cudaStream_t stream) {
  using namespace h16;
  const bool split = A_lo != nullptr;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  // output: fp32 C, or (C == NULL) a 16-bit C16 alone -- the CTA-pair kernel's 64-column store tiles
  const bool out16 = C == nullptr && C16 != nullptr;
  if (out16) ldc = ldc16;
  if (split && (B_lo == nullptr || a_bf || b_bf || beta != 0.f || !al16(A_lo) || !al16(B_lo) ||
                !(g_h16_pair && M > BM && (N & 1) == 0 && (ldc & 1) == 0))) {
    set_error("gemm_h16: the split-precision product needs FP16 hi/lo pairs for both operands, beta = 0 and the CTA-pair "
              "kernel (M > 128, even N)");
    return DEER_ERR_UNSUPPORTED;
  }
  if (!al16(A) || !al16(B) || !al16(C) || (lda & 7) || (ldb & 7) || (!out16 && (ldc & 3)) || (bias && !al16(bias)) ||
      (C16 && (!al16(C16) || (ldc16 & 7)))) {
    set_error("gemm_h16: operands must be 16-byte aligned with leading dimensions that are multiples of 16 bytes");
    return DEER_ERR_UNSUPPORTED;
  }
  if ((C16 != nullptr && !out16) || (C == nullptr && !out16) || (beta != 0.f && act != DEER_ACT_NONE) ||
      (bias && (N & 3)) || (out16 && (beta != 0.f || !(g_h16_pair && M > BM && (N & 1) == 0)))) {
    set_error("gemm_h16: unsupported epilogue (fp32 AND 16-bit output, activation on an accumulating GEMM, bias with "
              "N % 4, or a 16-bit output outside the CTA-pair kernel / with beta != 0)");
    return DEER_ERR_UNSUPPORTED;
  }
  CUtensorMap ma, mb, mc;
  bool ok = out16 ? make_map16(&mc, C16, c16_bf, M, N, ldc16, 64, 32) : make_map_c(&mc, C, M, N, ldc);
  if (!ok) {
    set_error("gemm_h16: cuTensorMapEncodeTiled failed for C (M=%d N=%d ldc=%lld)", M, N, ldc);
    return DEER_ERR_UNSUPPORTED;
  }
  if (!transA) ok = make_map16(&ma, A, a_bf, M, K, lda, BK, BM);   // stored [M,K]: K-major
  else ok = make_map16(&ma, A, a_bf, K, M, lda, 64, BK);           // stored [K,M]: MN-major
  if (transB) ok = ok && make_map16(&mb, B, b_bf, N, K, ldb, BK, BN);  // stored [N,K]: K-major
  else ok = ok && make_map16(&mb, B, b_bf, K, N, ldb, 64, BK);         // stored [K,N]: MN-major
  if (!ok) {
    set_error("gemm_h16: cuTensorMapEncodeTiled failed (M=%d N=%d K=%d lda=%lld ldb=%lld)", M, N, K, lda, ldb);
    return DEER_ERR_UNSUPPORTED;
  }
  const bool pair = g_h16_pair && M > BM && (N & 1) == 0 && (ldc & 1) == 0;
  if (pair && transB) {   // the pair kernel stages HALF of the K-major B tile per CTA: box of 128 rows
    if (!make_map16(&mb, B, b_bf, N, K, ldb, BK, BN / 2)) {
      set_error("gemm_h16: cuTensorMapEncodeTiled failed for B (pair box)");
      return DEER_ERR_UNSUPPORTED;
    }
  }
  CUtensorMap ma2 = ma, mb2 = mb;
  if (split) {   // the lo operands: same geometry as their hi partners
    bool ok2 = transA ? make_map16(&ma2, A_lo, 0, K, M, lda, 64, BK) : make_map16(&ma2, A_lo, 0, M, K, lda, BK, BM);
    ok2 = ok2 && (transB ? make_map16(&mb2, B_lo, 0, N, K, ldb, BK, BN / 2) : make_map16(&mb2, B_lo, 0, K, N, ldb, 64, BK));
    if (!ok2) {
      set_error("gemm_h16: cuTensorMapEncodeTiled failed for the lo operands");
      return DEER_ERR_UNSUPPORTED;
    }
  }
  const int tile_rows = pair ? 2 * BM : BM;
  const int tiles_m = (M + tile_rows - 1) / tile_rows, tiles_n = (N + BN - 1) / BN;
  const int kb_phase = (K + BK - 1) / BK;
  const int total_kb = split ? 3 * kb_phase : kb_phase;
  int splits = 1;
  if (beta == 1.f && act == DEER_ACT_NONE && C16 == nullptr) {
    const int tiles = tiles_m * tiles_n;
    const int units = pair ? kNumSMs / 2 : kNumSMs;   // schedulable CTAs / CTA pairs
    if (tiles < units && total_kb >= 32) {
      splits = (2 * units + tiles - 1) / tiles;
      const int max_splits = total_kb / 8;
      if (splits > max_splits) splits = max_splits;
      if (splits < 1) splits = 1;
    }
  }
  int per = (total_kb + splits - 1) / splits;
  splits = (total_kb + per - 1) / per;
  // instruction descriptor: D fp32, A/B formats, majors, N = 256, M = 128
  const uint32_t idesc = (1u << 4) | ((uint32_t)(a_bf ? 1 : 0) << 7) | ((uint32_t)(b_bf ? 1 : 0) << 10) |
                         ((uint32_t)(transA ? 1 : 0) << 15) | ((uint32_t)(transB ? 0 : 1) << 16) |
                         ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(tile_rows >> 4) << 24);
  Params p{C, ldc, C16, ldc16, c16_bf, bias, M, N, K, act, beta, splits, per, tiles_m, tiles_n, idesc, g_h16_prof,
           getenv("DEER_H16_DEBUG") ? atoi(getenv("DEER_H16_DEBUG")) : 0, split ? kb_phase : 0, t_drop_mask, t_drop_scale};
  if (p.drop_mask && (!pair || out16 || beta != 0.f || (N & 31) || ldc < N || p.debug)) {
    set_error("gemm_h16: the dropout-mask epilogue needs the CTA-pair kernel (M > 128), fp32 output, beta = 0 and N %% 32 == 0");
    return DEER_ERR_UNSUPPORTED;
  }
  const int work = tiles_m * tiles_n * splits;
  if (pair) {
    const int clusters = work < kNumSMs / 2 ? work : kNumSMs / 2;
#define DEER_H16_PAIR_GO(AM, BMN)                                                                                    \
  do {                                                                                                               \
    static bool attr = false;                                                                                        \
    if (!attr) {                                                                                                     \
      cudaError_t e = cudaFuncSetAttribute(gemm_h16_pair_kernel<2, AM, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                           P_SMEM_BYTES);                                                            \
      if (e != cudaSuccess) return cuda_status(e, "gemm_h16 pair smem attribute");                                   \
      attr = true;                                                                                                   \
    }                                                                                                                \
    DEER_LAUNCH((gemm_h16_pair_kernel<2, AM, BMN>), 2 * clusters, NUM_THREADS, P_SMEM_BYTES, stream, ma, mb, mc, ma2, mb2, p);     \
  } while (0)
    if (!transA && transB) DEER_H16_PAIR_GO(false, false);
    else if (!transA && !transB) DEER_H16_PAIR_GO(false, true);
    else if (transA && transB) DEER_H16_PAIR_GO(true, false);
    else DEER_H16_PAIR_GO(true, true);
#undef DEER_H16_PAIR_GO
    return DEER_OK;
  }
  const int grid = work < kNumSMs ? work : kNumSMs;
#define DEER_H16_GO(AM, BMN)                                                                                       \
  do {                                                                                                             \
    static bool attr = false;                                                                                      \
    if (!attr) {                                                                                                   \
      cudaError_t e = cudaFuncSetAttribute(gemm_h16_kernel<AM, BMN>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                           SMEM_BYTES);                                                            \
      if (e != cudaSuccess) return cuda_status(e, "gemm_h16 smem attribute");                                      \
      attr = true;                                                                                                 \
    }                                                                                                              \
    DEER_LAUNCH((gemm_h16_kernel<AM, BMN>), grid, NUM_THREADS, SMEM_BYTES, stream, ma, mb, mc, p);                     \
  } while (0)
  if (!transA && transB) DEER_H16_GO(false, false);
  else if (!transA && !transB) DEER_H16_GO(false, true);
  else if (transA && transB) DEER_H16_GO(true, false);
  else DEER_H16_GO(true, true);
#undef DEER_H16_GO
  return DEER_OK;
}

int gemm_tf32_pair(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                   long long ldc, int M, int N, int K, const float* bias, int act, float beta, cudaStream_t stream);
bool gemm_tf32_pair_supported(int transA, int transB, int M, int N, int K, long long ldc, const float* bias, int act,
                              float beta);
// ------------------------------------------------------------------------------------------------- TF32 on CTA pairs
// Same kernel with fp32 operands read as TF32 (ELEM = 4): the 128x128-tile TF32 engine (gemm_tcgen05.cu) is bound by
// shared-memory / L2 operand traffic; 256x256 pair tiles move a quarter of the operand bytes per FLOP and CTA.
int g_tf32_pair = 1;   // deer_set_option(DEER_OPT_TF32_PAIR)
// row term handed to the NEXT gemm_tf32_pair call of this thread (set and cleared by gemm_tf32_pair_rowterm)
static thread_local const float* t_rt_w = nullptr;
static thread_local const float* t_rt_v = nullptr;
static thread_local int t_rt_B = 0, t_rt_T = 0, t_rt_tm = 0;
bool gemm_tf32_pair_supported(int transA, int transB, int M, int N, int K, long long ldc, const float* bias, int act,
                              float beta) {
  (void)transA; (void)transB;
  if (!g_tf32_pair) return false;
  if (M <= 256 || N < 128 || K < 64) return false;                    // too small to feed 74 CTA pairs
  if ((N & 3) || (ldc & 3)) return false;
  if (bias && (reinterpret_cast<uintptr_t>(bias) & 15)) return false;
  if (beta != 0.f && act != DEER_ACT_NONE) return false;
  return true;
}
int gemm_tf32_pair(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                   long long ldc, int M, int N, int K, const float* bias, int act, float beta, cudaStream_t stream) {
  using namespace h16;
  constexpr int BKE = 32;
  CUtensorMap ma, mb, mc;
  bool ok = make_map_c(&mc, C, M, N, ldc);
  if (!transA) ok = ok && tc::make_map(&ma, A, M, K, lda, BKE, BM, false);          // stored [M,K]: K-major
  else ok = ok && tc::make_map(&ma, A, K, M, lda, 32, BKE, true);                    // stored [K,M]: MN-major
  if (transB) ok = ok && tc::make_map(&mb, B, N, K, ldb, BKE, BN / 2, false);       // stored [N,K]: K-major, half tile
  else ok = ok && tc::make_map(&mb, B, K, N, ldb, 32, BKE, true);                    // stored [K,N]: MN-major
  if (!ok) {
    set_error("gemm_tf32_pair: cuTensorMapEncodeTiled failed (M=%d N=%d K=%d lda=%lld ldb=%lld)", M, N, K, lda, ldb);
    return DEER_ERR_UNSUPPORTED;
  }
  const int tiles_m = (M + 2 * BM - 1) / (2 * BM), tiles_n = (N + BN - 1) / BN;
  const int total_kb = (K + BKE - 1) / BKE;
  int splits = 1;
  if (beta == 1.f && act == DEER_ACT_NONE) {
    const int tiles = tiles_m * tiles_n;
    const int units = kNumSMs / 2;
    if (tiles < units && total_kb >= 64) {
      splits = (2 * units + tiles - 1) / tiles;
      const int max_splits = total_kb / 16;
      if (splits > max_splits) splits = max_splits;
      if (splits < 1) splits = 1;
    }
  }
  int per = (total_kb + splits - 1) / splits;
  splits = (total_kb + per - 1) / per;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(transA ? 1 : 0) << 15) |
                         ((uint32_t)(transB ? 0 : 1) << 16) | ((uint32_t)(BN >> 3) << 17) |
                         ((uint32_t)((2 * BM) >> 4) << 24);
  Params p{C, ldc, nullptr, 0, 0, bias, M, N, K, act, beta, splits, per, tiles_m, tiles_n, idesc, nullptr,
           getenv("DEER_H16_DEBUG") ? atoi(getenv("DEER_H16_DEBUG")) : 0, 0, nullptr, 1.f, t_rt_w, t_rt_v, t_rt_B, t_rt_T,
           t_rt_tm};
  if (p.rt_w && (beta != 0.f || (N & 3) || ldc < N || p.debug || (long long)t_rt_B * t_rt_T != M ||
                 (reinterpret_cast<uintptr_t>(p.rt_v) & 15))) {
    set_error("gemm_tf32_pair: the row-term epilogue needs beta = 0, N %% 4 == 0, rows = B x T and a 16-byte aligned v");
    return DEER_ERR_UNSUPPORTED;
  }
  const int work = tiles_m * tiles_n * splits;
  const int clusters = work < kNumSMs / 2 ? work : kNumSMs / 2;
#define DEER_TF32_PAIR_GO(AM, BMN)                                                                                  \
  do {                                                                                                              \
    static bool attr = false;                                                                                       \
    if (!attr) {                                                                                                    \
      cudaError_t e = cudaFuncSetAttribute(gemm_h16_pair_kernel<4, AM, BMN>,                                        \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, P_SMEM_BYTES);              \
      if (e != cudaSuccess) return cuda_status(e, "gemm_tf32_pair smem attribute");                                 \
      attr = true;                                                                                                  \
    }                                                                                                               \
    DEER_LAUNCH((gemm_h16_pair_kernel<4, AM, BMN>), 2 * clusters, NUM_THREADS, P_SMEM_BYTES, stream, ma, mb, mc, ma, mb, p); \
  } while (0)
  if (!transA && transB) DEER_TF32_PAIR_GO(false, false);
  else if (!transA && !transB) DEER_TF32_PAIR_GO(false, true);
  else if (transA && transB) DEER_TF32_PAIR_GO(true, false);
  else DEER_TF32_PAIR_GO(true, true);
#undef DEER_TF32_PAIR_GO
  return DEER_OK;
}

// C = A B^T-or-B + w (x) v per sample (see Params::rt_w): the TF32 CTA-pair GEMM with the row-term epilogue
int gemm_tf32_pair_rowterm(const float* A, long long lda, int transA, const float* B, long long ldb, int transB, float* C,
                           long long ldc, int M, int N, int K, const float* w, const float* v, int nb, int nt, int time_major,
                           cudaStream_t stream) {
  if (!gemm_tf32_pair_supported(transA, transB, M, N, K, ldc, nullptr, DEER_ACT_NONE, 0.f)) {
    set_error("gemm_rowterm: shape outside the TF32 CTA-pair engine (M=%d N=%d K=%d)", M, N, K);
    return DEER_ERR_UNSUPPORTED;
  }
  t_rt_w = w; t_rt_v = v; t_rt_B = nb; t_rt_T = nt; t_rt_tm = time_major;
  const int rc = gemm_tf32_pair(A, lda, transA, B, ldb, transB, C, ldc, M, N, K, nullptr, DEER_ACT_NONE, 0.f, stream);
  t_rt_w = t_rt_v = nullptr; t_rt_B = t_rt_T = t_rt_tm = 0;
  return rc;
}

// fp32 -> FP16 hi / lo pair of a row-major matrix: hi = fp16(x), lo = fp16(x - hi) (22 significant bits together);
// same geometry as cast16 (columns cols..cols_pad-1 zero-filled).  row_scale (optional): row r is multiplied by
// row_scale[r] first (the text encoder's attention mask, encoders.py:733-735 -- the masked copy is never materialised)
__global__ void __launch_bounds__(256) cast_split16_kernel(const float* __restrict__ src, long long ld_src,
                                                           const float* __restrict__ row_scale,
                                                           uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                           long long ld_dst, long long rows, int cols, int cols_pad) {
  DEER_PDL_ENTRY();
  const int quads = cols_pad >> 2;                     // cols_pad % 4 == 0
  const long long total = rows * quads;
  const bool vec = (ld_src & 3) == 0 && (reinterpret_cast<uintptr_t>(src) & 15) == 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / quads;
    const int c = (int)(i % quads) * 4;
    float v[4];
    if (vec && c + 3 < cols) {
      const float4 x = __ldcs(reinterpret_cast<const float4*>(src + r * ld_src + c));
      v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] = (c + j < cols) ? src[r * ld_src + c + j] : 0.f;
    }
    if (row_scale) {
      const float sc = __ldg(row_scale + r);
#pragma unroll
      for (int j = 0; j < 4; j++) v[j] *= sc;
    }
    __half h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      h[j] = __float2half_rn(v[j]);
      l[j] = __float2half_rn(v[j] - __half2float(h[j]));
    }
    *reinterpret_cast<uint2*>(hi + r * ld_dst + c) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(lo + r * ld_dst + c) = *reinterpret_cast<const uint2*>(l);
  }
}

// fp32 -> 16-bit cast of a row-major matrix; dst row pitch ld_dst >= cols (columns cols..cols_pad-1 are zero-filled)
__global__ void __launch_bounds__(256) cast16_kernel(const float* __restrict__ src, long long ld_src,
                                                     uint16_t* __restrict__ dst, long long ld_dst, long long rows,
                                                     int cols, int cols_pad, int bf) {
  DEER_PDL_ENTRY();
  const long long per_row = cols_pad / 2;
  const long long total = rows * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / per_row;
    const int c = (int)(i % per_row) * 2;
    const float a = c < cols ? src[r * ld_src + c] : 0.f;
    const float b = c + 1 < cols ? src[r * ld_src + c + 1] : 0.f;
    uint32_t v;
    if (bf) {
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      v = *reinterpret_cast<uint32_t*>(&h);
    } else {
      __half2 h = __floats2half2_rn(a, b);
      v = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint32_t*>(dst + r * ld_dst + c) = v;
  }
}

}  // namespace deer

using namespace deer;

namespace deer {
// LSTM layer operands in one pass: W_ih of both directions (nn.LSTM row order g*H+u, fp32 [4H,In]) -> gate-interleaved
// (row 4u+g) 16-bit [2*4H, Kp] (columns >= In zero-filled), and b_ih + b_hh -> gate-interleaved fp32 [2*4H].
// Replaces 2 row permutations + 2 bias sums + 2 bias permutations + 1 cast (7 launches at the head of every layer).
__global__ void __launch_bounds__(256) lstm_prep_kernel(const float* __restrict__ w_f, const float* __restrict__ w_r,
                                                        const float* __restrict__ bi_f, const float* __restrict__ bh_f,
                                                        const float* __restrict__ bi_r, const float* __restrict__ bh_r,
                                                        uint16_t* __restrict__ w16, float* __restrict__ b_il, int H, int In,
                                                        int Kp, int bf) {
  DEER_PDL_ENTRY();
  const int G = 4 * H;
  const long long per_row = Kp / 2;
  const long long total = 2LL * G * per_row;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int ro = (int)(i / per_row);            // output row: d*G + 4u + g
    const int c = (int)(i % per_row) * 2;
    const int d = ro / G, rr = ro % G;
    const int srow = (rr & 3) * H + (rr >> 2);    // g*H + u
    const float* w = d ? w_r : w_f;
    const float a = c < In ? w[(long long)srow * In + c] : 0.f;
    const float b = c + 1 < In ? w[(long long)srow * In + c + 1] : 0.f;
    uint32_t v;
    if (bf) {
      __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
      v = *reinterpret_cast<uint32_t*>(&h);
    } else {
      __half2 h = __floats2half2_rn(a, b);
      v = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint32_t*>(w16 + (long long)ro * Kp + c) = v;
    if (b_il != nullptr && c == 0) b_il[ro] = (d ? bi_r : bi_f)[srow] + (d ? bh_r : bh_f)[srow];
  }
}
// The reverse for the gradients: gate-interleaved dW_ih [2,4H,In], dW_hh [2,4H,H], db [2,4H] ACCUMULATED into the
// natural-order targets of both directions (8 launches at the tail of every layer's backward -> 1).
struct LstmGradTargets {
  float* dwi[2];
  float* dwh[2];
  float* dbi[2];
  float* dbh[2];
};
__global__ void __launch_bounds__(256) lstm_unprep_kernel(const float* __restrict__ dwi_il, const float* __restrict__ dwh_il,
                                                          const float* __restrict__ db_il, const LstmGradTargets tg, int H,
                                                          int In) {
  DEER_PDL_ENTRY();
  const int G = 4 * H;
  const long long n_wi = 2LL * G * In, n_wh = 2LL * G * H, n_b = 2LL * G;
  const long long total = n_wi + n_wh + n_b;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    if (i < n_wi) {
      const int k = (int)(i % In);
      const int ro = (int)(i / In);
      const int d = ro / G, rr = ro % G;
      tg.dwi[d][(long long)((rr & 3) * H + (rr >> 2)) * In + k] += dwi_il[i];
    } else if (i < n_wi + n_wh) {
      const long long j = i - n_wi;
      const int k = (int)(j % H);
      const int ro = (int)(j / H);
      const int d = ro / G, rr = ro % G;
      tg.dwh[d][(long long)((rr & 3) * H + (rr >> 2)) * H + k] += dwh_il[j];
    } else {
      const int ro = (int)(i - n_wi - n_wh);
      const int d = ro / G, rr = ro % G;
      const int srow = (rr & 3) * H + (rr >> 2);
      const float v = db_il[ro];
      tg.dbi[d][srow] += v;     // b_ih and b_hh receive the same gradient
      tg.dbh[d][srow] += v;
    }
  }
}
}  // namespace deer

extern "C" {

int deer_gemm_h16(const void* A, long long lda, int transA, int a_bf16, const void* B, long long ldb, int transB,
                  int b_bf16, float* C, long long ldc, void* C16, long long ldc16, int c16_bf16, int M, int N, int K,
                  const float* bias, int act, float beta, void* stream) {
  DEER_CHECK_ARG(A && B && (C || C16), "gemm_h16: null pointer");
  DEER_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_h16: empty shape");
  // a leading dimension smaller than the row length = overlapping rows (the sliding-window Conv1d operands): K-major A
  // (lda < K), MN-major B (ldb < N) and an accumulating fp32 C (ldc < N with beta = 1: TMA reduce-add) are supported
  DEER_CHECK_ARG(lda > 0 && ldb > 0 && (transA ? lda >= M : true) && (transB ? ldb >= K : true) &&
                     (C ? (ldc >= N || (ldc > 0 && beta == 1.f)) : ldc16 >= N),
                 "gemm_h16: leading dimension too small");
  DEER_CHECK_ARG(act >= 0 && act <= 3, "gemm_h16: bad activation");
  DEER_CHECK_ARG(beta == 0.f || beta == 1.f, "gemm_h16: beta must be 0 or 1");
  // tcgen05.mma kind::f16 takes ONE 16-bit format for both operands (a mixed descriptor traps as an illegal instruction)
  DEER_CHECK_ARG((a_bf16 != 0) == (b_bf16 != 0), "gemm_h16: A and B must both be FP16 or both be BF16");
  g_engine_calls[DEER_ENGINE_H16]++;
  return gemm_h16(A, lda, transA, a_bf16, B, ldb, transB, b_bf16, C, ldc, C16, ldc16, c16_bf16, M, N, K, bias, act, beta,
                  (cudaStream_t)stream);
}

int deer_gemm_h16_dropmask(const void* A, long long lda, int transA, int a_bf16, const void* B, long long ldb, int transB,
                           int b_bf16, float* C, long long ldc, int M, int N, int K, const void* keep_mask, float scale,
                           void* stream) {
  DEER_CHECK_ARG(A && B && C && keep_mask && M > 0 && N > 0 && K > 0 && scale > 0.f, "gemm_h16_dropmask: bad args");
  DEER_CHECK_ARG(lda > 0 && ldb > 0 && ldc >= N && (transA ? lda >= M : true) && (transB ? ldb >= K : true),
                 "gemm_h16_dropmask: leading dimension too small");
  DEER_CHECK_ARG((a_bf16 != 0) == (b_bf16 != 0), "gemm_h16_dropmask: A and B must both be FP16 or both be BF16");
  DEER_CHECK_ARG((reinterpret_cast<uintptr_t>(keep_mask) & 3) == 0, "gemm_h16_dropmask: mask alignment");
  g_engine_calls[DEER_ENGINE_H16]++;
  t_drop_mask = reinterpret_cast<const uint32_t*>(keep_mask);
  t_drop_scale = scale;
  const int rc = gemm_h16(A, lda, transA, a_bf16, B, ldb, transB, b_bf16, C, ldc, nullptr, 0, 0, M, N, K, nullptr, DEER_ACT_NONE,
                          0.f, (cudaStream_t)stream);
  t_drop_mask = nullptr;
  t_drop_scale = 1.f;
  return rc;
}

int deer_gemm_h16_split(const void* A_hi, const void* A_lo, long long lda, int transA, const void* B_hi, const void* B_lo,
                        long long ldb, int transB, float* C, long long ldc, int M, int N, int K, const float* bias, int act,
                        void* stream) {
  DEER_CHECK_ARG(A_hi && A_lo && B_hi && B_lo && C, "gemm_h16_split: null pointer");
  DEER_CHECK_ARG(M > 0 && N > 0 && K > 0, "gemm_h16_split: empty shape");
  // lda < K with !transA = overlapping rows (sliding-window operand of the Conv1d taps): any 16-byte-multiple pitch
  DEER_CHECK_ARG(lda > 0 && (transA ? lda >= M : true) && ldb >= (transB ? K : N) && ldc >= N,
                 "gemm_h16_split: leading dimension too small");
  DEER_CHECK_ARG(act >= 0 && act <= 3, "gemm_h16_split: bad activation");
  g_engine_calls[DEER_ENGINE_H16_SPLIT]++;
  return gemm_h16_ex(A_hi, A_lo, lda, transA, 0, B_hi, B_lo, ldb, transB, 0, C, ldc, nullptr, 0, 0, M, N, K, bias, act,
                     0.f, (cudaStream_t)stream);
}

int deer_cast_split16(const float* src, long long ld_src, const float* row_scale, void* hi, void* lo, long long ld_dst,
                      long long rows, int cols, int cols_pad, void* stream) {
  DEER_CHECK_ARG(src && hi && lo && rows > 0 && cols > 0 && cols_pad >= cols && (cols_pad & 3) == 0 && ld_dst >= cols_pad &&
                     (ld_dst & 3) == 0,
                 "cast_split16: bad args");
  DEER_CHECK_ARG(((reinterpret_cast<uintptr_t>(hi) | reinterpret_cast<uintptr_t>(lo)) & 7) == 0,
                 "cast_split16: outputs must be 8-byte aligned");
  long long g = cdiv(rows * (cols_pad / 4), 256);
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  DEER_LAUNCH(cast_split16_kernel, (unsigned)g, 256, 0, stream, src, ld_src, row_scale, reinterpret_cast<uint16_t*>(hi),
              reinterpret_cast<uint16_t*>(lo), ld_dst, rows, cols, cols_pad);
  return DEER_OK;
}

int deer_gemm_h16_set_profile_buffer(long long* device_buf) {
  gemm_h16_set_profile(device_buf);
  return DEER_OK;
}

int deer_cast16(const float* src, long long ld_src, void* dst, long long ld_dst, long long rows, int cols, int cols_pad,
                int bf16, void* stream) {
  DEER_CHECK_ARG(src && dst && rows > 0 && cols > 0 && cols_pad >= cols && (cols_pad & 1) == 0 && ld_dst >= cols_pad &&
                     (ld_dst & 1) == 0,
                 "cast16: bad args");
  long long g = cdiv(rows * (cols_pad / 2), 256);
  if (g > kNumSMs * 16) g = kNumSMs * 16;
  DEER_LAUNCH(cast16_kernel, (unsigned)g, 256, 0, stream, src, ld_src, reinterpret_cast<uint16_t*>(dst), ld_dst, rows, cols,
              cols_pad, bf16);
  return DEER_OK;
}

int deer_lstm_prep(const float* w_ih_fwd, const float* w_ih_rev, const float* b_ih_fwd, const float* b_hh_fwd,
                   const float* b_ih_rev, const float* b_hh_rev, void* w16, float* b_il, int H, int In, int Kp, int bf16,
                   void* stream) {
  DEER_CHECK_ARG(w_ih_fwd && w_ih_rev && w16 && H > 0 && In > 0 && Kp >= In && (Kp & 1) == 0, "lstm_prep: bad args");
  DEER_CHECK_ARG(b_il == nullptr || (b_ih_fwd && b_hh_fwd && b_ih_rev && b_hh_rev), "lstm_prep: biases");
  long long g = cdiv(2LL * 4 * H * (Kp / 2), 256);
  if (g > kNumSMs * 8) g = kNumSMs * 8;
  DEER_LAUNCH(lstm_prep_kernel, (unsigned)g, 256, 0, stream, w_ih_fwd, w_ih_rev, b_ih_fwd, b_hh_fwd, b_ih_rev, b_hh_rev,
              reinterpret_cast<uint16_t*>(w16), b_il, H, In, Kp, bf16);
  return DEER_OK;
}

int deer_lstm_unprep(const float* dwi_il, const float* dwh_il, const float* db_il, float* dw_ih_fwd, float* dw_ih_rev,
                     float* dw_hh_fwd, float* dw_hh_rev, float* db_ih_fwd, float* db_hh_fwd, float* db_ih_rev,
                     float* db_hh_rev, int H, int In, void* stream) {
  DEER_CHECK_ARG(dwi_il && dwh_il && db_il && dw_ih_fwd && dw_ih_rev && dw_hh_fwd && dw_hh_rev && db_ih_fwd &&
                     db_hh_fwd && db_ih_rev && db_hh_rev && H > 0 && In > 0,
                 "lstm_unprep: bad args");
  DEER_CHECK_ARG(db_ih_fwd != db_hh_fwd && db_ih_rev != db_hh_rev, "lstm_unprep: b_ih and b_hh targets must differ");
  LstmGradTargets tg{{dw_ih_fwd, dw_ih_rev}, {dw_hh_fwd, dw_hh_rev}, {db_ih_fwd, db_ih_rev}, {db_hh_fwd, db_hh_rev}};
  long long g = cdiv(2LL * 4 * H * (In + H + 1), 256);
  if (g > kNumSMs * 8) g = kNumSMs * 8;
  DEER_LAUNCH(lstm_unprep_kernel, (unsigned)g, 256, 0, stream, dwi_il, dwh_il, db_il, tg, H, In);
  return DEER_OK;
}

}  // extern "C"
