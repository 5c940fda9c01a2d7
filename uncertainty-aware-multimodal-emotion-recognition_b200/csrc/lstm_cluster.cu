// Persistent bidirectional-LSTM recurrence, forward and backward-through-time, for H = 256
// (encoders.py:82-89: nn.LSTM(84|512 -> 256, num_layers=2, bidirectional); call :380; autograd of it).
//
// One thread-block CLUSTER of 4 CTAs owns one (direction, N-row batch tile) recurrence for all T steps; no per-step
// launch, no grid sync, nothing but the h / dh exchange crosses CTAs:
//
//   forward   gates_t^T[4H, N] = pre_t^T + W_hh[4H,256] h_{t-1}^T[256, N]
//     * CTA r owns hidden units 64r..64r+63: their 256 gate rows of W_hh stay resident for the whole sequence as the
//       A operand of tcgen05.mma (FP16, either 128 KB of 128B-swizzled shared memory or 256 TMEM columns); rows are
//       ordered 4*unit+gate so the four gates of a unit sit in four adjacent TMEM lanes of one warp.  FP16 operands
//       carry the same 11 significant bits as TF32 for h in (-1,1) and for the weights, at twice the tensor rate and
//       half the exchange bytes; accumulation is FP32 in TMEM.
//     * per step: 2 accumulators (M128 x N) x 16 K-steps; the gate epilogue reads its TMEM lane (one gate row, N batch
//       columns), adds the time-batched input projection (prefetched one step ahead), applies sigmoid/tanh on the
//       MUFU pipe, transposes through a per-warp shared tile so one thread holds i,f,g,o of (unit, N/4 columns),
//       updates the cell state held in REGISTERS across all T steps and writes h_t / gates / c_t.
//     * h_t is all-gathered through DISTRIBUTED SHARED MEMORY: every warp packs its 8 units x N columns into 16-byte
//       FP16 chunks and stores them straight into the swizzled B-operand buffer of all 4 CTAs (st.shared::cluster),
//       then releases one remote mbarrier arrival per destination; buffers are double-buffered so no "free" handshake
//       is needed (a peer can only be one step ahead).
//
//   backward  dh_{t-1}^T[256, N] = W_hh^T[256, 4H] dpre_t^T[4H, N]
//     * CTA r owns the same 64 units: it holds dc in registers, turns (dh_out_t + dh_rec, saved gates, c_t, c_{t-1})
//       into the pre-activation gradients dpre_t of ITS 256 gate rows (written to HBM for the weight-gradient GEMMs
//       and, as BF16, into the B-operand tile), multiplies by its K-slice of W_hh^T (A operand [256 hid x 256 k],
//       BF16, resident) and REDUCE-SCATTERS the partial dh over DSMEM: each accumulator row is stored to the CTA that
//       owns that hidden unit; the owner sums the 4 partials at the start of the next step.
//
// Warp roles (288 threads): warps 0..7 compute (TMEM lane group = warp % 4, accumulator = warp / 4), warp 8 lane 0
// issues tcgen05.mma / tcgen05.commit.  Grid = 4 x ceil(B/N) x 2 directions CTAs (B=256, N=16: 128 CTAs).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "tc_ptx.cuh"

namespace deer {
namespace tc {

constexpr int QH = 256;        // hidden size
constexpr int QC = 4;          // cluster size
constexpr int QU = QH / QC;    // hidden units per CTA (64)
constexpr int QTHREADS = 256;  // BPTT kernel: 8 compute warps, warp 0 also issues the MMAs (no control warp, see there)
constexpr int QW_BYTES = 2 * 128 * QH * 2;  // resident A operand: 2 accumulators x 128 rows x 256 k x 16 bit = 128 KB
constexpr int Q_WCOL = 64;     // TMEM column where the resident A operand starts (TS mode)

// ------------------------------------------------------------------------------------------------- extra PTX
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t raddr, uint4 v) {
  asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
// async-proxy bulk copy local shared -> (remote) shared of a cluster peer; completes `bytes` on the peer's mbarrier
__device__ __forceinline__ void bulk_copy_to_peer(uint32_t dst_cluster, uint32_t src_cta, uint32_t bytes,
                                                  uint32_t mbar_cluster) {
  asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   dst_cluster),
               "r"(src_cta), "r"(bytes), "r"(mbar_cluster)
               : "memory");
}
// asynchronous 16-byte store into the shared memory of a cluster peer; its completion (16 bytes) is counted on the
// peer's mbarrier -- no staging fence, no bulk-copy descriptor through the TMA unit
__device__ __forceinline__ void st_async_v4(uint32_t dst_cluster, uint4 v, uint32_t mbar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(
                   dst_cluster),
               "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(mbar_cluster)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t raddr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_cluster(bar, parity)) {
    if (++spins > (1u << 26)) mbar_timeout_trap();
  }
}
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
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
template <int N>
__device__ __forceinline__ void tmem_ld_row(uint32_t taddr, float* v) {
  if constexpr (N == 8) {
    tmem_ld8(taddr, v);
  } else if constexpr (N == 16) {
    tmem_ld16(taddr, v);
  } else {
    tmem_ld32(taddr, v);
  }
}
// kind::f16 instruction descriptor: D fp32, A/B format fmt (0 = F16, 1 = BF16), both K-major
__host__ __device__ constexpr uint32_t make_idesc_f16(uint32_t fmt, int M, int N) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
  __half2 h = __floats2half2_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
template <bool BF>
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  return BF ? pack_bf2(a, b) : pack_h2(a, b);
}

// The gate epilogue is MUFU-bound (16 MUFU lanes per SM and clock; one ex2 + one rcp per sigmoid).  Two sigmoids share
// ONE reciprocal: 1/a = b * rcp(a b), 1/b = a * rcp(a b); with the exponent clamped at 2^60 the product stays finite
// (sigmoid < 1e-18 there).  Flush-to-zero MUFU forms: no denormal pre/post-scaling instructions around them.
__device__ __forceinline__ float ex2_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// (1 / (1 + 2^t0), 1 / (1 + 2^t1))
__device__ __forceinline__ void sigmoid_pair_ex2(float t0, float t1, float& s0, float& s1) {
  const float a0 = 1.f + ex2_ftz(fminf(t0, 60.f));
  const float a1 = 1.f + ex2_ftz(fminf(t1, 60.f));
  const float R = rcp_ftz(a0 * a1);
  s0 = a1 * R;
  s1 = a0 * R;
}
// (tanh x0, tanh x1) = 1 - 2 / (1 + e^{2x})
__device__ __forceinline__ void tanh_pair(float x0, float x1, float& t0, float& t1) {
  float s0, s1;
  sigmoid_pair_ex2(x0 * 2.8853900817779268f, x1 * 2.8853900817779268f, s0, s1);
  t0 = fmaf(-2.f, s0, 1.f);
  t1 = fmaf(-2.f, s1, 1.f);
}

// byte offset of 16-byte chunk `c` (0..7) of row `row` inside a K-major SWIZZLE_128B tile whose 8-row groups are 1024 B apart
__device__ __forceinline__ uint32_t sw128(int row, int c) {
  return (uint32_t)((row >> 3) * 1024 + (row & 7) * 128 + ((c ^ (row & 7)) << 4));
}

template <int N>
struct QLayout {
  static constexpr int NQ = N / 4;                   // batch columns per cell thread
  static constexpr int ROWF = N + 4;                 // padded fp32 row (floats) of the transpose / partial tiles
  static constexpr int HB_BYTES = N * QH * 2;        // one B-operand tile: N rows x 256 k x 16 bit
  static constexpr int ACT_BYTES = 8 * 32 * ROWF * 4;  // forward: per-warp activation transpose tiles
  static constexpr int SH_BYTES = 2 * 8 * N * 16;      // forward: double-buffered per-warp fp16 h blocks [N cols][8 units]
  // backward: the partial dh rows travel as BF16 (DSMEM moves ~20 B/cycle/SM, so the reduce-scatter volume is the
  // longest stage of a BPTT step; BF16 halves it and matches the operand precision of the MMA that follows)
  static constexpr int PART_BYTES = QC * QU * N * 2;        // one buffer of partial-dh slots [src][unit][N] bf16
  static constexpr int PSTAGE_BYTES = 2 * 8 * 32 * N * 2;   // double-buffered per-warp partial rows [32][N] bf16
};

// Layouts.  `pre` / `dpre` are GATE-INTERLEAVED [T,B,2,H,4] (column 4*unit+gate: the input-projection GEMM runs on
// row-interleaved W_ih), so one warp's 8 units x 4 gates of a sample are one 128-byte line.  The activated gates and
// the cell states saved for BPTT are private to the forward/backward kernel pair and stored BLOCKED in the order the
// cell threads hold them: gact [T,2,ntiles,4 CTA,8 warp,4 gate,N/16,32 lane,4], c [T,2,ntiles,4,8,N/16,32,4] -- every
// access to them is one 128-bit load/store per thread, 512 contiguous bytes per warp.
struct LstmClusterParams {
  float* gates;        // fwd: pre-activations in; bwd: dpre out           [T,B,2,H,4]
  const float* w_fwd;  // [4H,H] (natural nn.LSTM row order g*H+u)
  const float* w_rev;  // [4H,H]
  float* h_out;        // fwd: [T,B,2H] out
  float* gact;         // blocked activated gates: fwd out (or null), bwd in
  float* c_all;        // blocked cell states: fwd out (or null), bwd in
  const float* dh_out; // bwd: [T,B,2H]
  float* db;           // bwd: [2,H,4] bias gradient (column sums of dpre), accumulated with atomics; may be null
  // optional 16-bit shadows for the tcgen05 16-bit GEMM engine (null to skip): forward h as FP16 and/or BF16
  // [T,B,2H]; backward dpre as BF16 [T,B,2,H,4]
  __half* h16;
  __nv_bfloat16* hb16;
  __nv_bfloat16* dpre16;
  const __half* pre16;  // fwd: FP16 pre-activations [T,B,2,H,4] (read instead of `gates` when non-null)
  int T, B, ntiles, keep;
  long long* prof;     // optional clock64 trace of block 0 (tools/lstm_probe.py --prof), else null
  int keep16;          // 1: the kept gates / cell states are FP16 (same blocked order, half the bytes), 0: fp32
  int stasync;         // forward h all-gather: 1 = per-lane st.async stores into the peers' B operand, 0 = bulk copies
  // XIN kernels (input projection inside the recurrence, first layer): x as time-major FP16 rows [T*B, xk] (xk % 8 == 0,
  // xk <= 128), both directions' gate-interleaved FP16 W_ih [2*4H, xk] and the gate-interleaved bias b_ih + b_hh [2*4H]
  const __half* x16;
  const __half* wih16;
  const float* bias_il;
  int xk;
};
constexpr int Q_PROF_S0 = 64, Q_PROF_STEPS = 4, Q_PROF_SLOTS = 8;
// (only in the FAST = false instantiation -- the probe's: the clock read and its guards were 3 % of all instructions the
//  production kernel issued)
#define Q_PROF(slot)                                                                               \
  do {                                                                                             \
    if constexpr (!FAST) {                                                                         \
      if (p.prof && blockIdx.x == 0 && s >= Q_PROF_S0 && s < Q_PROF_S0 + Q_PROF_STEPS)             \
        p.prof[(s - Q_PROF_S0) * Q_PROF_SLOTS + (slot)] = clock64();                               \
    }                                                                                              \
  } while (0)

// ================================================================================================ forward
// G = 1: one batch tile of N columns per CTA (8 compute warps + the MMA warp).
// G = 2 (inference, large batches): TWO independent batch sub-tiles of N columns per CTA, each with its own 8 compute
// warps, accumulators, h tiles and barriers, sharing the resident W_hh and the MMA warp.  A step of one tile is a serial
// chain (MMA -> gate epilogue -> cell update -> exchange); with two tiles in flight one tile's MUFU-bound epilogue runs
// while the other's MMAs and DSMEM exchange are in flight, instead of every unit idling in turn (a monolithic 32-column
// tile took 2.75 us per step; two waves of CTAs were needed at B = 1024).
// CS = 2 (training, N = 16): the SAME tile and MMAs, but two sets of 8 compute warps that each take one half of the
// tile's batch columns (a warp may read its TMEM lane quadrant at any column): the gate epilogue / cell update /
// exchange of a step -- a latency-bound chain with only two warps per scheduler -- runs at half the length per warp.
// The kept gate / cell layouts stay those of the 8-warp kernel, so the BPTT kernel is unchanged.
// G = 2 with CS = 2 ("half split", HS; N = 16): the two column halves of ONE 16-column tile run as two INDEPENDENT
// recurrences -- own 8 compute warps, accumulators, B-operand tiles (8 valid rows of a 16-row tile; the MMA stays N = 16,
// the tensor pipe is 90 % idle anyway) and barriers, sharing the resident W_hh and the MMA warp.  A step of one half is a
// serial chain MMA (420 cycles) -> gate epilogue -> cell update -> DSMEM all-gather (750-900 cycles from "h computed" to
// "the peers' MMA warps see it", whichever transport is used); with two halves in flight one half's exchange flies while
// the other half's gates are computed, on all 128 SMs of a B = 256 batch.  Same kept layouts as the 8-warp kernel.
// FAST: the batch is a whole number of tiles and no clock trace is requested -- the per-column bounds checks (a branch
// around every global load / store of the inner loop) and the trace guards are compiled out.
// XIN: the layer's INPUT PROJECTION runs inside the recurrence (first layer, In <= 128): the CTA keeps its 256 rows of the
// gate-interleaved FP16 W_ih in TMEM beside W_hh (128 more columns), the MMA warp fetches the step's [N x In] FP16 tile of x one step ahead and
// issues W_ih x_t into the (double-buffered) accumulators BEFORE h_{t-1} arrives -- off the step's critical path, on a
// tensor pipe that is 90 % idle -- and the gate epilogue adds the bias.  The FP16 pre-activation tensor [T,B,2,4H] (315 MB
// written by a GEMM and read back here at B = 256; 1.26 GB each way at B = 1024) never exists.
// With two sub-tiles (G = 2) the XIN kernel runs TWO MMA warps, one per sub-tile: with a single issuer the input part of one
// sub-tile sits in front of the other's recurrent MMAs (measured, layer-0 forward at B = 1024: 1637 us with one issuer,
// 1204 us with two, 1338 us on the GEMM path).  Without XIN a second issuer LOSES (the no-keep forward of one layer at
// B = 1024: 0.99 -> 1.49 ms), so that kernel keeps one.
template <int N, bool TS, int G, int CS = 1, bool FAST = false, bool XIN = false>
__global__ void __cluster_dims__(QC, 1, 1)
    __launch_bounds__((G == 2 && CS == 2 ? 512 : G * CS * 256) + ((XIN && G == 2) ? 64 : 32), 1)
    lstm_fwd_cluster_kernel(const LstmClusterParams p) {
  DEER_PDL_ENTRY();
  static_assert(CS == 1 || (CS == 2 && N == 16), "column split: one 16-column tile, two warp sets");
  static_assert(!XIN || (TS && CS == 1 && N == 16), "XIN: 16-column tiles, TMEM-resident W_hh");
  constexpr bool HS = (G == 2 && CS == 2);
  using L = QLayout<N>;
  constexpr int NW = N / CS;             // batch columns per compute warp
  constexpr int NQ = NW / 4;             // batch columns per cell thread
  constexpr int ROWF = NW + 4;           // padded fp32 row of the per-warp transpose tile
  constexpr int NCW = HS ? 16 : 8 * G * CS;   // compute warps
  constexpr uint32_t XBYTES = HS ? L::HB_BYTES / 2 : L::HB_BYTES;   // bytes of h landing in one B-operand tile per step
  constexpr int MMAW = NCW;              // index of the (first) MMA-issuing warp
  constexpr int NMW = (XIN && G == 2) ? 2 : 1;   // MMA warps
  constexpr int NTHREADS = NCW * 32 + 32 * NMW;
  constexpr int ACT_TOTAL = NCW * 32 * ROWF * 4;   // per-warp activation transpose tiles
  constexpr int SH_TOTAL = NCW * 2 * NW * 16;      // double-buffered per-warp fp16 h blocks [NW cols][8 units]
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms, applied as an OFFSET on the __shared__ array: going through
  // uintptr_t would make every later access a generic LD/ST instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* wsm = smem;                                   // SS: 128 KB resident W (unused in TS mode)
  uint8_t* hbuf_all = smem + (TS ? 0 : QW_BYTES);        // per group: 2 x HB_BYTES
  float* stage_act_all = reinterpret_cast<float*>(hbuf_all + G * 2 * L::HB_BYTES);
  __half* stage_h_all = reinterpret_cast<__half*>(reinterpret_cast<uint8_t*>(stage_act_all) + ACT_TOTAL);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stage_h_all) + SH_TOTAL);
  uint64_t* h_full_all = bars;            // [G][2]
  uint64_t* mma_done_all = bars + 2 * G;  // [G]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * G);
  // XIN: per group two [16 k-chunks][N rows][16 B] x tiles (the B-operand layout of the h tiles)
  constexpr int XIN_OFF = (((TS ? 0 : QW_BYTES) + G * 2 * L::HB_BYTES + ACT_TOTAL + SH_TOTAL + 3 * G * 8 + 16) + 1023) & ~1023;
  constexpr int XT_BYTES = 16 * N * 16;
  uint8_t* xbuf_all = smem + XIN_OFF;
  constexpr uint32_t ACC_BUF = XIN ? (uint32_t)(G * 2 * N) : 0u;   // XIN: accumulators double-buffered by step parity

  const int warp_id = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = (G == 2 && warp_id >= 8 && warp_id < 16) ? 1 : 0;   // batch sub-tile of this compute warp
  const int chalf = (CS == 2 && warp_id >= 8 && warp_id < 16) ? 1 : 0;  // column half of this compute warp
  const int warp = warp_id >= MMAW ? 8 : (warp_id & 7);              // role index: 0..7 compute, 8 = MMA issuer
  const int wslot = warp_id < NCW ? warp_id : 0;                     // private staging slot of a compute warp
  uint8_t* hbuf = hbuf_all + grp * 2 * L::HB_BYTES;
  uint64_t* h_full = h_full_all + 2 * grp;
  uint64_t* mma_done = mma_done_all + grp;
  const uint32_t r = cluster_ctarank();
  const int cid = blockIdx.x / QC;
  const int ctile = cid % p.ntiles, dir = cid / p.ntiles;   // p.ntiles counts CTA tiles of G*N (HS: N) batch columns
  const int tile = HS ? ctile : ctile * G + grp;
  const int ntiles_all = HS ? p.ntiles : p.ntiles * G;      // N-column tiles per direction (the kept layouts' unit)
  const int b0 = tile * N;
  const int T = p.T, B = p.B;
  const float* __restrict__ W = dir ? p.w_rev : p.w_fwd;

  // ---- one-time setup
  if (threadIdx.x == 0) {
    for (int g = 0; g < G; g++) {
      mbar_init(&h_full_all[2 * g + 0], 1);
      mbar_init(&h_full_all[2 * g + 1], 1);
      mbar_init(&mma_done_all[g], 1);
    }
    fence_barrier_init();
    // each phase of h_full[b] = one arming arrival + N x 256 fp16 of h landing from the 4 CTAs (async proxy)
    for (int g = 0; g < G; g++) {
      if (T > 1) mbar_expect_tx(&h_full_all[2 * g + 0], XBYTES);
      if (T > 2) mbar_expect_tx(&h_full_all[2 * g + 1], XBYTES);
    }
  }
  // two allocations (accumulators 64 columns, resident operand 256) instead of one 512-column block: the 192 columns
  // left over let a 128-column GEMM CTA of a concurrent stream share the SM instead of spinning in tcgen05.alloc
  if (warp_id == MMAW) {
    if constexpr (TS) {
      tmem_alloc_more_follow(tmem_slot, (XIN && G == 2) ? 128 : 64);
      if constexpr (XIN) {
        tmem_alloc_more_follow(tmem_slot + 1, 256);
        tmem_alloc(tmem_slot + 2, 128);     // W_ih: 2 accumulators x 64 columns (K padded to 128 fp16)
      } else {
        tmem_alloc(tmem_slot + 1, 256);
      }
    } else {
      tmem_alloc(tmem_slot, 64);
    }
  }
  if constexpr (XIN) {
    for (int idx = threadIdx.x; idx < G * 2 * XT_BYTES / 16; idx += NTHREADS)   // chunks beyond xk stay zero for good
      reinterpret_cast<uint4*>(xbuf_all)[idx] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
  }
  if constexpr (!TS) {
    // W_hh rows of this CTA -> fp16, K-major SWIZZLE_128B: [acc a][k-block 4][128 rows x 128 B]; row m = 4*unit + gate
    for (int idx = threadIdx.x; idx < 2 * 128 * 32; idx += NTHREADS) {
      const int kc = idx & 31, m = (idx >> 5) & 127, a = idx >> 12;
      const int g = m & 3, u = a * 32 + (m >> 2);
      const float4* src = reinterpret_cast<const float4*>(W + (size_t)(g * QH + (int)r * QU + u) * QH + kc * 8);
      const float4 x0 = __ldg(src), x1 = __ldg(src + 1);
      uint4 v;
      v.x = pack_h2(x0.x, x0.y); v.y = pack_h2(x0.z, x0.w); v.z = pack_h2(x1.x, x1.y); v.w = pack_h2(x1.z, x1.w);
      *reinterpret_cast<uint4*>(wsm + a * 65536 + (kc >> 3) * 16384 + sw128(m, kc & 7)) = v;
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_w = TS ? tmem_slot[1] : 0u;   // resident A operand (TS mode)
  const uint32_t tmem_wih = XIN ? tmem_slot[2] : 0u;
  if constexpr (XIN) {
    if (warp_id < 8) {
      // this CTA's 256 gate-interleaved rows of W_ih (rows dir*4H + 256 r + a*128 + m) -> TMEM, lane = row m of
      // accumulator a, 64 columns = 128 fp16 (zero beyond xk)
      const int a = warp >> 2, sub = warp & 3, m = sub * 32 + lane;
      const int nchw = p.xk >> 3;
      const uint4* src = reinterpret_cast<const uint4*>(p.wih16 + ((size_t)dir * (4 * QH) + (size_t)r * 256 + a * 128 + m) * p.xk);
#pragma unroll 1
      for (int ch = 0; ch < 2; ch++) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 8; i++) {
          const int kc = ch * 8 + i;
          const uint4 x = kc < nchw ? __ldg(src + kc) : make_uint4(0u, 0u, 0u, 0u);
          reinterpret_cast<uint4*>(v)[i] = x;
        }
        tmem_st32(tmem_wih + ((uint32_t)(sub * 32) << 16) + (uint32_t)(a * 64 + ch * 32), v);
      }
      tmem_st_wait();
    }
  }
  if constexpr (TS) {
    if (warp_id < 8) {
      // resident A operand in TMEM: lane = row m of accumulator a, 128 columns = 256 fp16 (2 per column, low half first)
      const int a = warp >> 2, sub = warp & 3, m = sub * 32 + lane;
      const int g = m & 3, u = a * 32 + (m >> 2);
      const float* src = W + (size_t)(g * QH + (int)r * QU + u) * QH;
#pragma unroll 1
      for (int ch = 0; ch < 4; ch++) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 16; i++) {
          const float4 x = __ldg(reinterpret_cast<const float4*>(src + ch * 64) + i);
          reinterpret_cast<uint32_t*>(v)[2 * i] = pack_h2(x.x, x.y);
          reinterpret_cast<uint32_t*>(v)[2 * i + 1] = pack_h2(x.z, x.w);
        }
        tmem_st32(tmem_w + ((uint32_t)(sub * 32) << 16) + (uint32_t)(a * 128 + ch * 32), v);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  cluster_sync_all();  // every CTA's barriers are initialised before any peer signals them

  if (warp == 8) {
    // =================================================================== MMA issuer (whole warp, one elected lane issues)
    {
      constexpr uint32_t idesc = make_idesc_f16(0, 128, N);
      const uint32_t tb = warp_uniform(tmem_base);
      const uint32_t tw = warp_uniform(tmem_w);
      const uint32_t twi = warp_uniform(tmem_wih);
      const bool leader = elect_one();
      // XIN: fetch the [N x xk] FP16 tile of x for step s of sub-tile g into its B-operand tile and issue W_ih x_s into
      // the accumulator buffer of that step's parity (this warp wrote the tile itself: generic -> async proxy fence only)
      const int nch = XIN ? (p.xk >> 3) : 0;
      const int nkx = XIN ? ((p.xk + 15) >> 4) : 0;
      auto x_load = [&](const int s, const int g, uint4 (&v)[N / 2]) {
        if constexpr (XIN) {
          // <= 16 chunks per row x N rows / 32 lanes = N/2 per lane, all in flight at once; issued BEFORE this warp waits
          // for h, so the round trip to L2 / DRAM costs the recurrence nothing
          const int t = dir ? T - 1 - s : s;
          const int bg0 = (ctile * G + g) * N;
#pragma unroll
          for (int i = 0; i < N / 2; i++) {
            const int idx = lane + 32 * i;
            const int n = idx % N, kc = idx / N;
            v[i] = make_uint4(0u, 0u, 0u, 0u);
            if (kc < nch && (FAST || bg0 + n < B))
              v[i] = __ldg(reinterpret_cast<const uint4*>(p.x16 + ((size_t)t * B + bg0 + n) * p.xk) + kc);
          }
        }
      };
      auto x_issue = [&](const int s, const int g, const uint4 (&v)[N / 2]) {
        if constexpr (XIN) {
          uint8_t* xb = xbuf_all + (2 * g + (s & 1)) * XT_BYTES;
#pragma unroll
          for (int i = 0; i < N / 2; i++) {
            const int idx = lane + 32 * i;
            const int n = idx % N, kc = idx / N;
            if (kc < nch) *reinterpret_cast<uint4*>(xb + kc * (N * 16) + n * 16) = v[i];
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (leader) {
            const uint32_t tgx = tb + (uint32_t)(s & 1) * ACC_BUF + (uint32_t)(g * 2 * N);
#pragma unroll
            for (int a = 0; a < 2; a++) {
#pragma unroll 1
              for (int k = 0; k < nkx; k++) {
                const uint64_t bd = make_smem_desc(smem_u32(xb) + (2 * k) * (N * 16), N * 16, 128, 0);
                umma_f16_ts(tgx + a * N, twi + a * 64 + k * 8, bd, idesc, k > 0 ? 1u : 0u);
              }
            }
          }
          __syncwarp();
        }
      };
      // (XIN: iteration s = -1 only issues the input part of step 0 -- ONE call site of the input part keeps the kernel's
      //  register count where it was)
      for (int s = XIN ? -1 : 0; s < T; s++) {
#pragma unroll
        for (int g = 0; g < G; g++) {   // the batch sub-tiles take turns on the tensor core (NMW = 2: one warp each)
          if (NMW == 2 && g != warp_id - MMAW) continue;
          uint64_t* hf = h_full_all + 2 * g;
          uint64_t* md = mma_done_all + g;
          uint4 xv[N / 2];
          if (XIN && s + 1 < T) x_load(s + 1, g, xv);
          if (s <= 0) {
            if constexpr (XIN) {
              if (s == 0) {
                if (leader) umma_commit(md);   // the gates of step 0 are the input projection alone: W_ih x_0, in flight
                __syncwarp();
              }
              if (s + 1 < T) x_issue(s + 1, g, xv);
            } else {
              if (leader) mbar_arrive(md);  // h_{-1} = 0: the gates of step 0 are the input projection alone
            }
            continue;
          }
          if (p.stasync) {
            // the tile was written by the peers' (generic-proxy) st.async stores: acquire at cluster scope, then order
            // them before this warp's async-proxy reads (tcgen05.mma operand fetch)
            mbar_wait_cluster(&hf[(s - 1) & 1], (uint32_t)(((s - 1) >> 1) & 1));
            fence_proxy_async_smem();
          } else {
            mbar_wait(&hf[(s - 1) & 1], (uint32_t)(((s - 1) >> 1) & 1));
          }
          if (leader && g == 0) Q_PROF(0);
          if (leader && s + 2 < T) mbar_expect_tx(&hf[(s - 1) & 1], XBYTES);  // re-arm for h_{s+1}
          tc_fence_after();
          const uint32_t hb = smem_u32(hbuf_all) + (2 * g + ((s - 1) & 1)) * L::HB_BYTES;
          const uint32_t tg = tb + (uint32_t)(s & 1) * ACC_BUF + (uint32_t)(g * 2 * N);
          if (leader) {
#pragma unroll
            for (int a = 0; a < 2; a++) {
#pragma unroll
              for (int k = 0; k < 16; k++) {
                // B operand, K-major no-swizzle: [32 k-chunks][N rows][16 B]; 8x16B core matrices, SBO 128 B, LBO N*16 B
                const uint64_t bd = make_smem_desc(hb + (2 * k) * (N * 16), N * 16, 128, 0);
                if constexpr (TS) {
                  // (XIN: W_ih x_s is already in the accumulator)
                  umma_f16_ts(tg + a * N, tw + a * 128 + k * 8, bd, idesc, (XIN || k > 0) ? 1u : 0u);
                } else {
                  const uint64_t ad =
                      make_smem_desc(smem_u32(wsm) + a * 65536 + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024, 2);
                  umma_f16_ss(tg + a * N, ad, bd, idesc, k > 0 ? 1u : 0u);
                }
              }
            }
            umma_commit(md);
            if (g == 0) Q_PROF(1);
          }
          __syncwarp();
          // the next step's input part: its accumulator buffer was read out before h_{s-1} left, its x tile was consumed
          // by MMAs that completed before this step's; all of it is off the critical path
          if (XIN && s + 1 < T) x_issue(s + 1, g, xv);
        }
      }
    }
  } else {
    // =================================================================== gate / cell epilogue
    const int a = warp >> 2, sub = warp & 3;
    const int j = lane >> 2, g = lane & 3;        // gate-row role: unit j of this warp, gate g
    const int ul = a * 32 + sub * 8 + j;          // unit inside the CTA
    const int ug = (int)r * QU + ul;              // unit inside the direction
    const uint32_t tacc = tmem_base + ((uint32_t)(sub * 32) << 16) +
                          (uint32_t)(grp * 2 * N + a * N + (HS ? 0 : chalf * NW));
    // XIN: bias of this thread's gate row (the projection GEMM's epilogue added it before)
    const float xbias = XIN ? __ldg(p.bias_il + dir * (4 * QH) + (int)r * 256 + a * 128 + sub * 32 + lane) : 0.f;
    float* sa = stage_act_all + wslot * (32 * ROWF);
    __half* sh_base = stage_h_all + wslot * (2 * NW * 8);
    const int bw0 = b0 + chalf * NW;              // first batch column of this warp
    const float sc = (g == 2) ? 2.f : 1.f;        // tanh(x) = 2*sigmoid(2x) - 1 keeps the warp convergent
    const int q = g;                              // cell role: columns [q*NQ, q*NQ+NQ) of unit j
    float cst[NQ];
#pragma unroll
    for (int i = 0; i < NQ; i++) cst[i] = 0.f;
    // pre-activations of the NEXT step, as loaded (fp32 bits, or an FP16 value in the low half): converting at load time
    // would make the convert instruction wait for the DRAM round trip inside the current step
    uint32_t pre[NW];
    const bool pre_is16 = p.pre16 != nullptr;
    const int il0 = 4 * ((int)r * QU + a * 32 + sub * 8);  // first interleaved gate column of this warp
    auto load_pre = [&](int s) {
      const int t = dir ? T - 1 - s : s;
      const long long off = (((long long)t * B + bw0) * 2 + dir) * (4 * QH) + il0 + lane;
      if (pre_is16) {   // FP16 pre-activations (the projection GEMM's 16-bit epilogue): half the bytes of the only
                        // stream this kernel reads; one 64-byte request per warp and sample
        const unsigned short* src = reinterpret_cast<const unsigned short*>(p.pre16) + off;
#pragma unroll
        for (int n = 0; n < NW; n++)
          pre[n] = (FAST || bw0 + n < B) ? (uint32_t)__ldcs(src + (long long)n * (8 * QH)) : 0u;
      } else {
        const float* src = p.gates + off;
#pragma unroll
        for (int n = 0; n < NW; n++)
          pre[n] = (FAST || bw0 + n < B) ? __float_as_uint(__ldcs(src + (long long)n * (8 * QH))) : 0u;
      }
    };
    // blocked save area of this warp: ((((t*2+dir)*ntiles+tile)*4+r)*8+warp) blocks of 4*NQ*32 (gates) / NQ*32 (c)
    const long long blk_w = ((long long)dir * ntiles_all + tile) * 32 + (int)r * 8 + warp;
    const long long blk_t = 2LL * ntiles_all * 32;
    if constexpr (!XIN) load_pre(0);
    for (int s = 0; s < T; s++) {
      const int t = dir ? T - 1 - s : s;
      __half* sh = sh_base + (s & 1) * (NW * 8);
      mbar_wait(mma_done, (uint32_t)(s & 1));
      if (warp_id == 0 && lane == 0) Q_PROF(2);
      tc_fence_after();
      float x[NW];
      if (XIN || s > 0) {
        tmem_ld_row<NW>(tacc + (uint32_t)(s & 1) * ACC_BUF, x);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int n = 0; n < NW; n++) x[n] = 0.f;
      }
      tc_fence_before();
      if constexpr (XIN) {
#pragma unroll
        for (int n = 0; n < NW; n++) x[n] += xbias;
      } else
      if (pre_is16) {
#pragma unroll
        for (int n = 0; n < NW; n++) x[n] += __half2float(__ushort_as_half((unsigned short)pre[n]));
      } else {
#pragma unroll
        for (int n = 0; n < NW; n++) x[n] += __uint_as_float(pre[n]);
      }
      {
        const float nsl = -1.4426950408889634f * sc;   // sigmoid(sc x) = 1 / (1 + 2^(nsl x))
#pragma unroll
        for (int n = 0; n < NW; n += 2) {
          float s0, s1;
          sigmoid_pair_ex2(x[n] * nsl, x[n + 1] * nsl, s0, s1);
          x[n] = fmaf(sc, s0, 1.f - sc);
          x[n + 1] = fmaf(sc, s1, 1.f - sc);
        }
      }
#pragma unroll
      for (int n = 0; n < NW; n += 4)
        *reinterpret_cast<float4*>(sa + lane * ROWF + n) = make_float4(x[n], x[n + 1], x[n + 2], x[n + 3]);
      if (!XIN && s + 1 < T) load_pre(s + 1);  // a full step ahead of its use: DRAM latency never lands on the critical path
      __syncwarp();
      if (warp_id == 0 && lane == 0) Q_PROF(3);
      float gi[NQ], gf[NQ], gg[NQ], go[NQ];
      if constexpr (NQ == 2) {
        const float2 vi = *reinterpret_cast<const float2*>(sa + (4 * j + 0) * ROWF + q * NQ);
        const float2 vf = *reinterpret_cast<const float2*>(sa + (4 * j + 1) * ROWF + q * NQ);
        const float2 vg = *reinterpret_cast<const float2*>(sa + (4 * j + 2) * ROWF + q * NQ);
        const float2 vo = *reinterpret_cast<const float2*>(sa + (4 * j + 3) * ROWF + q * NQ);
        gi[0] = vi.x; gi[1] = vi.y; gf[0] = vf.x; gf[1] = vf.y;
        gg[0] = vg.x; gg[1] = vg.y; go[0] = vo.x; go[1] = vo.y;
      } else
#pragma unroll
      for (int i = 0; i < NQ; i += 4) {
        const float4 vi = *reinterpret_cast<const float4*>(sa + (4 * j + 0) * ROWF + q * NQ + i);
        const float4 vf = *reinterpret_cast<const float4*>(sa + (4 * j + 1) * ROWF + q * NQ + i);
        const float4 vg = *reinterpret_cast<const float4*>(sa + (4 * j + 2) * ROWF + q * NQ + i);
        const float4 vo = *reinterpret_cast<const float4*>(sa + (4 * j + 3) * ROWF + q * NQ + i);
        gi[i] = vi.x; gi[i + 1] = vi.y; gi[i + 2] = vi.z; gi[i + 3] = vi.w;
        gf[i] = vf.x; gf[i + 1] = vf.y; gf[i + 2] = vf.z; gf[i + 3] = vf.w;
        gg[i] = vg.x; gg[i + 1] = vg.y; gg[i + 2] = vg.z; gg[i + 3] = vg.w;
        go[i] = vo.x; go[i + 1] = vo.y; go[i + 2] = vo.z; go[i + 3] = vo.w;
      }
      float hv[NQ];
#pragma unroll
      for (int i = 0; i < NQ; i += 2) {
        const float c0 = fmaf(gf[i], cst[i], gi[i] * gg[i]);
        const float c1 = fmaf(gf[i + 1], cst[i + 1], gi[i + 1] * gg[i + 1]);
        cst[i] = c0;
        cst[i + 1] = c1;
        float t0, t1;
        tanh_pair(c0, c1, t0, t1);
        hv[i] = go[i] * t0;
        hv[i + 1] = go[i + 1] * t1;
        sh[(q * NQ + i) * 8 + j] = __float2half_rn(hv[i]);
        sh[(q * NQ + i + 1) * 8 + j] = __float2half_rn(hv[i + 1]);
      }
      if (warp_id == 0 && lane == 0) Q_PROF(4);
      if (s + 1 < T) {
        // all-gather FIRST (it is the only thing the next step waits for): this warp's [N cols x 8 units] fp16 block is
        // k-chunk 8r+4a+sub of every CTA's B operand; one async-proxy bulk copy per destination, completion counted on
        // the destination's h_full barrier
        const uint32_t dst = smem_u32(hbuf) + (s & 1) * L::HB_BYTES + ((int)r * 8 + 4 * a + sub) * (N * 16) +
                             (HS ? 0 : chalf * (NW * 16));   // HS: the half's own tile, rows 0 .. 7
        if (p.stasync) {
          // lane n sends column n of the block (8 units x fp16 = one 16-byte row of the B operand) straight to the four
          // CTAs: posted stores, completion counted on each destination's h_full barrier
          __syncwarp();
          if (lane < NW) {
            const uint4 hv8 = *reinterpret_cast<const uint4*>(sh + lane * 8);
            const uint32_t bar = smem_u32(&h_full[s & 1]);
#pragma unroll
            for (uint32_t d = 0; d < QC; d++) st_async_v4(mapa_u32(dst + lane * 16, d), hv8, mapa_u32(bar, d));
          }
        } else {
        fence_proxy_async_smem();
        __syncwarp();
        if (lane < QC) {
          bulk_copy_to_peer(mapa_u32(dst, (uint32_t)lane), smem_u32(sh), NW * 16,
                            mapa_u32(smem_u32(&h_full[s & 1]), (uint32_t)lane));
        }
        }
        if (warp_id == 0 && lane == 0) Q_PROF(5);
      } else {
        __syncwarp();
      }
      // ---- everything below is off the recurrence's critical path: global stores of this step's results
      const long long row0 = (long long)t * B + bw0 + q * NQ;
      const long long blk = (long long)t * blk_t + blk_w;
#pragma unroll
      for (int i = 0; i < NQ; i++)
        if (FAST || bw0 + q * NQ + i < B) p.h_out[(row0 + i) * (2 * QH) + dir * QH + ug] = hv[i];
      if constexpr (CS == 2) {
        if (p.keep) {
          // the 8-warp kernel's blocked layout ([gate][32 lanes][4 columns] per warp block): this thread's two columns
          // are one half of the float4 that lane (j, chalf*2 + q/2) of the 8-warp kernel would store
          const int lane8 = j * 4 + chalf * 2 + (q >> 1);
          const int eo = (q & 1) * 2;
          if (p.keep16) {
            // FP16 kept state: this thread's two columns are one 32-bit word of each of the (i,f) / (g,o) 16-byte groups
            // and of the 8-byte c group that lane8 of the 8-warp kernel stores
            uint32_t* gw = reinterpret_cast<uint32_t*>(reinterpret_cast<__half*>(p.gact) + blk * 512);
            uint32_t* cw = reinterpret_cast<uint32_t*>(reinterpret_cast<__half*>(p.c_all) + blk * 128);
            const int h2 = eo >> 1;
            __stcs(gw + lane8 * 4 + h2, pack_h2(gi[0], gi[1]));
            __stcs(gw + lane8 * 4 + 2 + h2, pack_h2(gf[0], gf[1]));
            __stcs(gw + (32 + lane8) * 4 + h2, pack_h2(gg[0], gg[1]));
            __stcs(gw + (32 + lane8) * 4 + 2 + h2, pack_h2(go[0], go[1]));
            __stcs(cw + lane8 * 2 + h2, pack_h2(cst[0], cst[1]));
          } else {
          float* gs = p.gact + blk * 512 + lane8 * 4 + eo;
          float* cs = p.c_all + blk * 128 + lane8 * 4 + eo;
          __stcs(reinterpret_cast<float2*>(gs + 0 * 128), make_float2(gi[0], gi[1]));
          __stcs(reinterpret_cast<float2*>(gs + 1 * 128), make_float2(gf[0], gf[1]));
          __stcs(reinterpret_cast<float2*>(gs + 2 * 128), make_float2(gg[0], gg[1]));
          __stcs(reinterpret_cast<float2*>(gs + 3 * 128), make_float2(go[0], go[1]));
          __stcs(reinterpret_cast<float2*>(cs), make_float2(cst[0], cst[1]));
          }
        }
      } else
      if (p.keep) {
        if (p.keep16) {
          // FP16 kept state (gates in (0,1) / (-1,1), cell state: 11 significant bits, far inside what the BF16 operands
          // of BPTT carry): per 4-column chunk one 16-byte store of (i,f) and one of (g,o) per thread -- 512 contiguous
          // bytes per warp -- and one 8-byte store of c; 2.5 KB per (sample, step, direction) instead of 5 KB
          uint4* gs = reinterpret_cast<uint4*>(reinterpret_cast<__half*>(p.gact) + blk * (4 * NQ * 32)) + lane;
          uint2* cs = reinterpret_cast<uint2*>(reinterpret_cast<__half*>(p.c_all) + blk * (NQ * 32)) + lane;
#pragma unroll
          for (int i = 0; i < NQ; i += 4) {
            uint4 a, b;
            a.x = pack_h2(gi[i], gi[i + 1]); a.y = pack_h2(gi[i + 2], gi[i + 3]);
            a.z = pack_h2(gf[i], gf[i + 1]); a.w = pack_h2(gf[i + 2], gf[i + 3]);
            b.x = pack_h2(gg[i], gg[i + 1]); b.y = pack_h2(gg[i + 2], gg[i + 3]);
            b.z = pack_h2(go[i], go[i + 1]); b.w = pack_h2(go[i + 2], go[i + 3]);
            __stcs(gs + (i / 4 * 2 + 0) * 32, a);
            __stcs(gs + (i / 4 * 2 + 1) * 32, b);
            __stcs(cs + (i / 4) * 32, make_uint2(pack_h2(cst[i], cst[i + 1]), pack_h2(cst[i + 2], cst[i + 3])));
          }
        } else {
        float4* gs = reinterpret_cast<float4*>(p.gact + blk * (4 * NQ * 32)) + lane;
        float4* cs = reinterpret_cast<float4*>(p.c_all + blk * (NQ * 32)) + lane;
#pragma unroll
        for (int i = 0; i < NQ; i += 4) {
          __stcs(gs + (0 * NQ + i) * 8, make_float4(gi[i], gi[i + 1], gi[i + 2], gi[i + 3]));
          __stcs(gs + (1 * NQ + i) * 8, make_float4(gf[i], gf[i + 1], gf[i + 2], gf[i + 3]));
          __stcs(gs + (2 * NQ + i) * 8, make_float4(gg[i], gg[i + 1], gg[i + 2], gg[i + 3]));
          __stcs(gs + (3 * NQ + i) * 8, make_float4(go[i], go[i + 1], go[i + 2], go[i + 3]));
          __stcs(cs + i * 8, make_float4(cst[i], cst[i + 1], cst[i + 2], cst[i + 3]));
        }
        }
      }
      if (p.h16 || p.hb16) {
        // the warp's [N cols x 8 units] fp16 block doubles as the source of the 16-bit shadows of h: one 16-byte
        // store per sample row (the tile is double-buffered and only READ by the bulk copies in flight)
        if (lane < NW && (FAST || bw0 + lane < B)) {
          const uint4 hv8 = *reinterpret_cast<const uint4*>(sh + lane * 8);
          const long long o = ((long long)t * B + bw0 + lane) * (2 * QH) + dir * QH + (int)r * QU + a * 32 + sub * 8;
          if (p.h16) *reinterpret_cast<uint4*>(p.h16 + o) = hv8;
          if (p.hb16) {
            const __half2* hp = reinterpret_cast<const __half2*>(&hv8);
            uint4 bv;
            uint32_t* bp = reinterpret_cast<uint32_t*>(&bv);
#pragma unroll
            for (int e = 0; e < 4; e++) {
              const float2 f = __half22float2(hp[e]);
              bp[e] = pack_bf2(f.x, f.y);
            }
            *reinterpret_cast<uint4*>(p.hb16 + o) = bv;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // no CTA retires while a peer may still address its shared memory
  if (warp_id == MMAW) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (XIN && G == 2) ? 128 : 64);
    if constexpr (TS) tmem_dealloc(tmem_w, 256);
    if constexpr (XIN) tmem_dealloc(tmem_wih, 128);
  }
}

// ================================================================================================ backward
// K16: the kept gates / cell states are FP16 (DEER_OPT_LSTM_KEEP16): they are loaded a step ahead as RAW bits and
// converted only when the step that uses them starts (a convert at the load would wait for the DRAM round trip inside
// the current step -- measured 1.48 -> 1.93 us per step)
template <int N, bool TS, bool K16 = false, bool FAST = false>
// 160 registers (not the 255 that 8 warps would allow): 2 warps x 5120 registers per sub-partition leave 6144 for the
// warps of a GEMM CTA of the concurrent stream
__global__ void __cluster_dims__(QC, 1, 1) __maxnreg__(160)
    lstm_bwd_cluster_kernel(const LstmClusterParams p) {
  DEER_PDL_ENTRY();
  using L = QLayout<N>;
  constexpr int NQ = L::NQ, ROWF = L::ROWF;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms, applied as an OFFSET on the __shared__ array: going through
  // uintptr_t would make every later access a generic LD/ST instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* wsm = smem;                                   // SS: 128 KB resident W^T slice
  uint8_t* bsm = smem + (TS ? 0 : QW_BYTES);             // B operand: dpre tile [N rows x 256 k] bf16
  __nv_bfloat16* part = reinterpret_cast<__nv_bfloat16*>(bsm + L::HB_BYTES);  // [2][QC src][QU units][N]
  __nv_bfloat16* pstage = part + 2 * (L::PART_BYTES / 2);                      // [8 warps][2][32][N]
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(pstage) + L::PSTAGE_BYTES);
  uint64_t* part_full = bars;     // [2]
  uint64_t* b_ready = bars + 2;
  uint64_t* mma_done = bars + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t r = cluster_ctarank();
  const int cid = blockIdx.x / QC;
  const int tile = cid % p.ntiles, dir = cid / p.ntiles;
  const int b0 = tile * N;
  const int T = p.T, B = p.B;
  const float* __restrict__ W = dir ? p.w_rev : p.w_fwd;

  if (threadIdx.x == 0) {
    mbar_init(&part_full[0], 1);
    mbar_init(&part_full[1], 1);
    mbar_init(b_ready, 8);
    mbar_init(mma_done, 1);
    fence_barrier_init();
    // each phase of part_full[b] = one arming arrival + the 4 sources' [64 x ROWF] partial-dh rows (async proxy)
    if (T > 1) mbar_expect_tx(&part_full[0], L::PART_BYTES);
    if (T > 2) mbar_expect_tx(&part_full[1], L::PART_BYTES);
  }
  // two allocations (accumulators 64 columns, resident operand 256) instead of one 512-column block: the 192 columns
  // left over let a 128-column GEMM CTA of a concurrent stream share the SM instead of spinning in tcgen05.alloc
  if (warp == 0) {
    if constexpr (TS) {
      tmem_alloc_more_follow(tmem_slot, 64);
      tmem_alloc(tmem_slot + 1, 256);
    } else {
      tmem_alloc(tmem_slot, 64);
    }
  }
  if constexpr (!TS) {
    // A[m = hidden unit (2 x 128)][k = 4*unit_local + gate] = W_hh[gate*256 + 64r + unit_local][m], bf16, K-major SW128
    for (int idx = threadIdx.x; idx < 2 * 32 * 128; idx += QTHREADS) {
      const int m = idx & 127, kc = (idx >> 7) & 31, a = idx >> 12;
      float f[8];
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const int ul = kc * 2 + (e >> 2), g = e & 3;
        f[e] = __ldg(W + (size_t)(g * QH + (int)r * QU + ul) * QH + a * 128 + m);
      }
      uint4 v;
      v.x = pack_bf2(f[0], f[1]); v.y = pack_bf2(f[2], f[3]); v.z = pack_bf2(f[4], f[5]); v.w = pack_bf2(f[6], f[7]);
      *reinterpret_cast<uint4*>(wsm + a * 65536 + (kc >> 3) * 16384 + sw128(m, kc & 7)) = v;
    }
    fence_proxy_async_smem();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_w = TS ? tmem_slot[1] : 0u;   // resident A operand (TS mode)
  if constexpr (TS) {
    {
      const int a = warp >> 2, sub = warp & 3, m = a * 128 + sub * 32 + lane;
#pragma unroll 1
      for (int ch = 0; ch < 4; ch++) {  // 64 k = 16 units x 4 gates per chunk
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; i++) {
          const int k0 = ch * 64 + 2 * i;
          const float f0 = __ldg(W + (size_t)((k0 & 3) * QH + (int)r * QU + (k0 >> 2)) * QH + m);
          const float f1 = __ldg(W + (size_t)(((k0 + 1) & 3) * QH + (int)r * QU + ((k0 + 1) >> 2)) * QH + m);
          reinterpret_cast<uint32_t*>(v)[i] = pack_bf2(f0, f1);
        }
        tmem_st32(tmem_w + ((uint32_t)(sub * 32) << 16) + (uint32_t)(a * 128 + ch * 32), v);
      }
      tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  cluster_sync_all();

  // No dedicated control warp in this kernel: warp 0 issues the MMAs of a step between handing over its own part of the
  // B operand and its global stores.  A 9th warp costs a whole register slot (156 registers x 32 lanes) in ONE of the SM's
  // four sub-partitions (warps are dealt round-robin, registers are allocated per sub-partition): with 3 x 5120 of its
  // 16384 registers taken no warp of another kernel fits there, so no CTA of a concurrent stream could ever share the SM
  // with a recurrence CTA (tools/probes/coresidency_probe.cu).  With 8 warps every sub-partition keeps 6144 registers free.
  constexpr uint32_t idesc = make_idesc_f16(1, 128, N);
  const uint32_t tb = warp_uniform(tmem_base);
  const uint32_t tw = warp_uniform(tmem_w);
  const bool leader = elect_one();
  auto issue_step = [&](const int s) {
    mbar_wait(b_ready, (uint32_t)(s & 1));
    if (leader) Q_PROF(0);
    // every cell thread has consumed the partials of step s-1: re-arm that buffer for step s+1
    if (leader && s >= 1 && s + 2 < T) mbar_expect_tx(&part_full[(s - 1) & 1], L::PART_BYTES);
    tc_fence_after();
    const uint32_t bb = smem_u32(bsm);
    if (leader) {
#pragma unroll
      for (int a = 0; a < 2; a++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
          const uint64_t bd = make_smem_desc(bb + (k >> 2) * (N * 128) + (k & 3) * 32, 16, 1024, 2);
          if constexpr (TS) {
            umma_f16_ts(tb + a * N, tw + a * 128 + k * 8, bd, idesc, k > 0 ? 1u : 0u);
          } else {
            const uint64_t ad = make_smem_desc(smem_u32(wsm) + a * 65536 + (k >> 2) * 16384 + (k & 3) * 32, 16, 1024, 2);
            umma_f16_ss(tb + a * N, ad, bd, idesc, k > 0 ? 1u : 0u);
          }
        }
      }
      umma_commit(mma_done);
      Q_PROF(1);
    }
    __syncwarp();
  };
  {
    const int a = warp >> 2, sub = warp & 3;
    const int j = lane >> 2, q = lane & 3;        // cell role: unit j of this warp, columns [q*NQ, q*NQ+NQ)
    const int ul = a * 32 + sub * 8 + j;
    const int ug = (int)r * QU + ul;
    // row role (partial-dh scatter): accumulator a, TMEM lane sub*32+lane -> hidden unit a*128 + sub*32 + lane
    const uint32_t tacc = tmem_base + ((uint32_t)(sub * 32) << 16) + (uint32_t)(a * N);
    const uint32_t dst_cta = (uint32_t)(2 * a + (sub >> 1));
    float dc[NQ];
#pragma unroll
    for (int i = 0; i < NQ; i++) dc[i] = 0.f;

    const long long blk_w = ((long long)dir * p.ntiles + tile) * 32 + (int)r * 8 + warp;
    const long long blk_t = 2LL * p.ntiles * 32;
    float sdb[4] = {0.f, 0.f, 0.f, 0.f};          // bias gradient of (unit, 4 gates) over this thread's columns, all t
    // Operands of a step -- the kept gates / cell states and dh_out -- are fetched TWO steps ahead into one of two
    // register sets (the loop body is instantiated once per set, so no register is ever indexed dynamically): ncu put
    // 22 % of all stall samples of the one-step-ahead version on the first FP16 -> FP32 convert of a step, i.e. on the
    // DRAM round trip of loads issued only half a step (~0.7 us) earlier.
    constexpr bool DEEP = K16 && N == 16;         // (32-column tiles / fp32 kept state: two sets would spill -- one step ahead)
    struct Kept {
      uint4 ra[NQ / 4], rb[NQ / 4];               // K16: raw FP16 bits of the (i,f) / (g,o) gates
      uint2 rc[NQ / 4], rcp[NQ / 4];              //      ... and of c_t / c_{t-1}
      float4 f[K16 ? 1 : 6 * (NQ / 4)];           // fp32 kept state: i, f, g, o, c, c_prev per 4-column chunk
      float vdh[NQ];
    };
    auto load_step = [&](int s, Kept& k) {
      const int t = dir ? s : T - 1 - s;            // reverse of the forward order
      const int tp = dir ? t + 1 : t - 1;           // forward-previous time step (c_{prev})
      const bool first = dir ? (t == T - 1) : (t == 0);
      const long long blk = (long long)t * blk_t + blk_w;
      if constexpr (K16) {
        const uint4* gs = reinterpret_cast<const uint4*>(reinterpret_cast<const __half*>(p.gact) + blk * (4 * NQ * 32)) + lane;
        const uint2* cs = reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.c_all) + blk * (NQ * 32)) + lane;
        const uint2* cps = reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(p.c_all) +
                                                          ((long long)tp * blk_t + blk_w) * (NQ * 32)) + lane;
#pragma unroll
        for (int i = 0; i < NQ; i += 4) {
          k.ra[i / 4] = __ldcs(gs + (i / 4 * 2 + 0) * 32);
          k.rb[i / 4] = __ldcs(gs + (i / 4 * 2 + 1) * 32);
          k.rc[i / 4] = __ldcs(cs + (i / 4) * 32);
          k.rcp[i / 4] = first ? make_uint2(0u, 0u) : __ldcs(cps + (i / 4) * 32);
        }
      } else {
        const float4* gs = reinterpret_cast<const float4*>(p.gact + blk * (4 * NQ * 32)) + lane;
        const float4* cs = reinterpret_cast<const float4*>(p.c_all + blk * (NQ * 32)) + lane;
        const float4* cps = reinterpret_cast<const float4*>(p.c_all + ((long long)tp * blk_t + blk_w) * (NQ * 32)) + lane;
#pragma unroll
        for (int i = 0; i < NQ; i += 4) {
          float4* f = k.f + 6 * (i / 4);
          f[0] = __ldcs(gs + (0 * NQ + i) * 8);
          f[1] = __ldcs(gs + (1 * NQ + i) * 8);
          f[2] = __ldcs(gs + (2 * NQ + i) * 8);
          f[3] = __ldcs(gs + (3 * NQ + i) * 8);
          f[4] = __ldcs(cs + i * 8);
          f[5] = first ? make_float4(0.f, 0.f, 0.f, 0.f) : __ldcs(cps + i * 8);
        }
      }
#pragma unroll
      for (int i = 0; i < NQ; i++) {
        const int b = b0 + q * NQ + i;
        k.vdh[i] = (FAST || b < B) ? __ldcs(p.dh_out + ((long long)t * B + b) * (2 * QH) + dir * QH + ug) : 0.f;
      }
    };
    auto body = [&](const int s, Kept& kept) {
      const int t = dir ? s : T - 1 - s;
      float vi[NQ], vf[NQ], vg[NQ], vo[NQ], vc[NQ], vcp[NQ];
      float dh[NQ];
#pragma unroll
      for (int i = 0; i < NQ; i++) dh[i] = kept.vdh[i];
      if constexpr (K16) {
        auto h2f = [](uint32_t v) { return __half22float2(*reinterpret_cast<const __half2*>(&v)); };
#pragma unroll
        for (int i = 0; i < NQ; i += 4) {
          const uint4 a = kept.ra[i / 4], b = kept.rb[i / 4];
          const uint2 c = kept.rc[i / 4], cp = kept.rcp[i / 4];
          float2 f;
          f = h2f(a.x); vi[i] = f.x; vi[i + 1] = f.y; f = h2f(a.y); vi[i + 2] = f.x; vi[i + 3] = f.y;
          f = h2f(a.z); vf[i] = f.x; vf[i + 1] = f.y; f = h2f(a.w); vf[i + 2] = f.x; vf[i + 3] = f.y;
          f = h2f(b.x); vg[i] = f.x; vg[i + 1] = f.y; f = h2f(b.y); vg[i + 2] = f.x; vg[i + 3] = f.y;
          f = h2f(b.z); vo[i] = f.x; vo[i + 1] = f.y; f = h2f(b.w); vo[i + 2] = f.x; vo[i + 3] = f.y;
          f = h2f(c.x); vc[i] = f.x; vc[i + 1] = f.y; f = h2f(c.y); vc[i + 2] = f.x; vc[i + 3] = f.y;
          f = h2f(cp.x); vcp[i] = f.x; vcp[i + 1] = f.y; f = h2f(cp.y); vcp[i + 2] = f.x; vcp[i + 3] = f.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < NQ; i += 4) {
          const float4* f = kept.f + 6 * (i / 4);
          vi[i] = f[0].x; vi[i + 1] = f[0].y; vi[i + 2] = f[0].z; vi[i + 3] = f[0].w;
          vf[i] = f[1].x; vf[i + 1] = f[1].y; vf[i + 2] = f[1].z; vf[i + 3] = f[1].w;
          vg[i] = f[2].x; vg[i + 1] = f[2].y; vg[i + 2] = f[2].z; vg[i + 3] = f[2].w;
          vo[i] = f[3].x; vo[i + 1] = f[3].y; vo[i + 2] = f[3].z; vo[i + 3] = f[3].w;
          vc[i] = f[4].x; vc[i + 1] = f[4].y; vc[i + 2] = f[4].z; vc[i + 3] = f[4].w;
          vcp[i] = f[5].x; vcp[i + 1] = f[5].y; vcp[i + 2] = f[5].z; vcp[i + 3] = f[5].w;
        }
      }
      if constexpr (DEEP) {
        if (s + 2 < T) load_step(s + 2, kept);   // this set is free again: its next use is two steps away
      }
      if (s > 0) {
        mbar_wait(&part_full[(s - 1) & 1], (uint32_t)(((s - 1) >> 1) & 1));
        if (warp == 0 && lane == 0) Q_PROF(2);
        const __nv_bfloat16* ps = part + ((s - 1) & 1) * (QC * QU * N) + ul * N + q * NQ;
#pragma unroll
        for (int src = 0; src < QC; src++) {
#pragma unroll
          for (int i = 0; i < NQ; i += 4) {
            const uint2 v = *reinterpret_cast<const uint2*>(ps + src * (QU * N) + i);
            const float2 lo = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.x));
            const float2 hi = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&v.y));
            dh[i] += lo.x; dh[i + 1] += lo.y; dh[i + 2] += hi.x; dh[i + 3] += hi.y;
          }
        }
      }
      float* gout = p.gates + (((long long)t * B + b0 + q * NQ) * 2 + dir) * (4 * QH) + 4 * ug;
      const int kb = ul >> 4, ch = (ul & 15) >> 1;
      float4 dp[NQ];
      uint2 dp16[NQ];
      float tcv[NQ];
#pragma unroll
      for (int i = 0; i < NQ; i += 2) tanh_pair(vc[i], vc[i + 1], tcv[i], tcv[i + 1]);
#pragma unroll
      for (int i = 0; i < NQ; i++) {
        const float tc_ = tcv[i];
        const float d_o = dh[i] * tc_;
        const float dcc = fmaf(dh[i] * vo[i], 1.f - tc_ * tc_, dc[i]);
        dc[i] = dcc * vf[i];
        dp[i].x = dcc * vg[i] * vi[i] * (1.f - vi[i]);
        dp[i].y = dcc * vcp[i] * vf[i] * (1.f - vf[i]);
        dp[i].z = dcc * vi[i] * (1.f - vg[i] * vg[i]);
        dp[i].w = d_o * vo[i] * (1.f - vo[i]);
        dp16[i].x = pack_bf2(dp[i].x, dp[i].y);
        dp16[i].y = pack_bf2(dp[i].z, dp[i].w);
        if (s + 1 < T) {
          const int n = q * NQ + i;
          *reinterpret_cast<uint2*>(bsm + kb * (N * 128) + sw128(n, ch) + (ul & 1) * 8) = dp16[i];
        }
      }
      if (s + 1 < T) {
        // hand the BF16 dpre tile to the tensor core FIRST; the global stores of dpre follow off the critical path
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(b_ready);
        if (warp == 0 && lane == 0) Q_PROF(3);
        if (warp == 0) issue_step(s);   // waits for the other 7 warps' tiles, then launches the step's 32 MMAs
      }
      auto store_step = [&]() {
#pragma unroll
        for (int i = 0; i < NQ; i++) {
          if (FAST || b0 + q * NQ + i < B) {
            if (p.gates) __stcs(reinterpret_cast<float4*>(gout + (long long)i * (8 * QH)), dp[i]);
            if (p.dpre16)
              *reinterpret_cast<uint2*>(p.dpre16 + (((long long)t * B + b0 + q * NQ + i) * 2 + dir) * (4 * QH) + 4 * ug) = dp16[i];
            sdb[0] += dp[i].x; sdb[1] += dp[i].y; sdb[2] += dp[i].z; sdb[3] += dp[i].w;
          }
        }
      };
      // warps 1-7 store in the shadow of the MMAs; warp 0 has just spent that time issuing them, so its stores wait until
      // its partial-dh rows are on their way to the peers (they would otherwise delay the next step of the whole cluster)
      const bool store_late = (warp == 0) && (s + 1 < T);
      if (!store_late) store_step();
      if (s + 1 < T) {
        if constexpr (!DEEP) load_step(s + 1, kept);   // one step ahead: streams in while the tensor core and the exchange run
        mbar_wait(mma_done, (uint32_t)(s & 1));
        if (warp == 0 && lane == 0) Q_PROF(4);
        tc_fence_after();
        float x[N];
        tmem_ld_row<N>(tacc, x);
        tmem_ld_wait();
        // reduce-scatter: this warp's 32 accumulator rows (hidden units of CTA dst_cta) go to that CTA's slot
        // [buf][src r][rows] with one async-proxy bulk copy; completion is counted on its part_full barrier
        __nv_bfloat16* st = pstage + (warp * 2 + (s & 1)) * (32 * N);
#pragma unroll
        for (int n = 0; n < N; n += 8) {
          uint4 v;
          v.x = pack_bf2(x[n], x[n + 1]); v.y = pack_bf2(x[n + 2], x[n + 3]);
          v.z = pack_bf2(x[n + 4], x[n + 5]); v.w = pack_bf2(x[n + 6], x[n + 7]);
          *reinterpret_cast<uint4*>(st + lane * N + n) = v;
        }
        fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          const uint32_t slot =
              smem_u32(part) + (uint32_t)(((s & 1) * QC + (int)r) * QU + (sub & 1) * 32) * (N * 2);
          bulk_copy_to_peer(mapa_u32(slot, dst_cta), smem_u32(st), 32 * N * 2,
                            mapa_u32(smem_u32(&part_full[s & 1]), dst_cta));
        }
        if (warp == 0 && lane == 0) Q_PROF(5);
      }
      if (store_late) store_step();
    };
    if constexpr (DEEP) {
      Kept set0, set1;
      load_step(0, set0);
      if (T > 1) load_step(1, set1);
      for (int s = 0; s < T; s += 2) {
        body(s, set0);
        if (s + 1 < T) body(s + 1, set1);
      }
    } else {
      Kept set0;
      load_step(0, set0);
      for (int s = 0; s < T; s++) body(s, set0);
    }
    if (p.db) {
      // db[dir, unit, gate] += sum over this cluster's samples and all T steps: fold the 4 column groups of a unit
#pragma unroll
      for (int e = 0; e < 4; e++) {
        sdb[e] += __shfl_xor_sync(0xffffffffu, sdb[e], 1);
        sdb[e] += __shfl_xor_sync(0xffffffffu, sdb[e], 2);
      }
      if (q == 0) {
#pragma unroll
        for (int e = 0; e < 4; e++) atomicAdd(p.db + dir * (4 * QH) + 4 * ug + e, sdb[e]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 64);
    if constexpr (TS) tmem_dealloc(tmem_w, 256);
  }
}

template <int N, bool TS, int G = 1, int CS = 1, bool XIN = false>
constexpr int fwd_smem_bytes() {
  using L = QLayout<N>;
  constexpr int NW = N / CS, NCW = (G == 2 && CS == 2) ? 16 : 8 * G * CS;
  if constexpr (XIN) {   // (same expression as XIN_OFF in the kernel) + W_ih + the x tiles + alignment slack
    constexpr int off = (((TS ? 0 : QW_BYTES) + G * 2 * L::HB_BYTES + NCW * 32 * (NW + 4) * 4 + NCW * 2 * NW * 16 + 3 * G * 8 + 16) + 1023) & ~1023;
    constexpr int needx = off + G * 2 * (16 * N * 16) + 1024;
    return needx > 120 * 1024 ? needx : 120 * 1024;
  }
  // >= 120 KB even in TS mode: one LSTM CTA per SM (two would not fit their 2 x 320 TMEM columns and the second would
  // spin in tcgen05.alloc); a TF32 GEMM CTA (100 KB, 128 columns) of another stream still fits beside it
  constexpr int need = (TS ? 0 : QW_BYTES) + G * 2 * L::HB_BYTES + NCW * 32 * (NW + 4) * 4 + NCW * 2 * NW * 16 + 64 + 1024;
  return need > 120 * 1024 ? need : 120 * 1024;
}
template <int N, bool TS>
constexpr int bwd_smem_bytes() {
  using L = QLayout<N>;
  constexpr int need = (TS ? 0 : QW_BYTES) + L::HB_BYTES + 2 * L::PART_BYTES + L::PSTAGE_BYTES + 64 + 1024;
  return need > 120 * 1024 ? need : 120 * 1024;
}

}  // namespace tc

static int g_lstm_ts = 1;      // resident operand in TMEM (1) or in shared memory (0)
static int g_lstm_tile = 0;    // 0 auto, else forced N (16 or 32)
int g_lstm_colsplit = 0;       // DEER_OPT_LSTM_COLSPLIT: 16 compute warps (two column halves) on 16-column tiles; measured: no gain (1.444 vs 1.438 us/step: the step is issue-bound, not latency-bound), so off
int g_lstm_dual = 1;           // DEER_OPT_LSTM_DUAL: two interleaved 16-column sub-tiles per CTA for no-keep 32-column tiles
int g_lstm_keep16 = 1;         // DEER_OPT_LSTM_KEEP16: FP16 (1, default) or fp32 (0) kept gates / cell states
int g_lstm_halfsplit = 0;      // DEER_OPT_LSTM_HALFSPLIT: forward recurrence on 16-column tiles as two independent 8-column halves.
                               // The KERNEL is 9 % faster that way (1.46 -> 1.33 us/step at B = 256), but its 17 warps x 96
                               // registers leave the video / text stream's kernels less room on the same SMs: the whole
                               // training step measured 4.30 ms with it vs 4.26 ms without (same box, back to back), so it is
                               // off by default.  (The same split of the BPTT kernel was built and measured slower by itself:
                               // 0.498 vs 0.402 ms per layer -- its kept-state reads become 4-byte accesses.)
int g_lstm_stasync = 0;        // DEER_OPT_LSTM_STASYNC: forward h all-gather by st.async stores (1) or bulk copies (0, default).
                               // Same-box A/B at the end of round 2: training step 4.02 ms (bulk) vs 4.06-4.12 ms (st.async);
                               // kernel alone at B = 256: 1.41 vs 1.44 us/step; B = 1024 inference 0.977 vs 0.960 ms per layer;
                               // B = 1024 training forward 1.39 vs 1.55 ms.  From issue to the peers' MMA warps seeing the tile
                               // both transports take 750-900 cycles.
int g_lstm_carveout = 1;       // DEER_OPT_LSTM_CARVEOUT: ask for the maximum shared-memory carve-out while the recurrence runs
static long long* g_lstm_prof = nullptr;
void lstm_cluster_set_profile(long long* buf) { g_lstm_prof = buf; }
void lstm_cluster_set_option(int ts, int tile) {
  if (ts >= 0) g_lstm_ts = ts ? 1 : 0;
  if (tile >= 0) g_lstm_tile = tile;
}

bool lstm_cluster_supported(const float* gates, const float* w_fwd, const float* w_rev, int H) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  return H == tc::QH && al16(gates) && al16(w_fwd) && al16(w_rev);
}

static int pick_tile(int B) {
  if (g_lstm_tile == 16 || g_lstm_tile == 32) return g_lstm_tile;
  // 4-CTA clusters: at most ~33 are co-resident on 148 SMs; prefer the narrow tile while one wave still covers B
  return (2 * ((B + 15) / 16) <= 32) ? 16 : 32;
}

template <int N, bool TS, int G, int CS, bool FAST, bool XIN = false>
static int launch_fwd_f(const tc::LstmClusterParams& p, cudaStream_t stream) {
  constexpr int smem = tc::fwd_smem_bytes<N, TS, G, CS, XIN>();
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc::lstm_fwd_cluster_kernel<N, TS, G, CS, FAST, XIN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cuda_status(e, "lstm_fwd_cluster smem attribute");
    // The L1 / shared-memory split of an SM is fixed while CTAs are resident, and the driver picks the smallest carve-out
    // that fits the kernel: ask for the maximum, so that a <= 100 KB GEMM CTA of the other stream can join the recurrence
    // CTA (what decides whether it does is the register file, see lstm_bwd_cluster_kernel).
    if (g_lstm_carveout)
      cudaFuncSetAttribute(tc::lstm_fwd_cluster_kernel<N, TS, G, CS, FAST, XIN>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
    attr = true;
  }
  DEER_LAUNCH((tc::lstm_fwd_cluster_kernel<N, TS, G, CS, FAST, XIN>), tc::QC * p.ntiles * 2,
              ((G == 2 && CS == 2) ? 512 : G * CS * 256) + ((XIN && G == 2) ? 64 : 32), smem, stream, p);
  return DEER_OK;
}
template <int N, bool TS, int G = 1, int CS = 1, bool XIN = false>
static int launch_fwd(const tc::LstmClusterParams& p, cudaStream_t stream) {
  // whole tiles (of the columns ONE CTA covers) and no clock trace: the check-free instantiation
  constexpr int cols = (G == 2 && CS == 2) ? N : G * N;
  if (p.B % cols == 0 && p.prof == nullptr) return launch_fwd_f<N, TS, G, CS, true, XIN>(p, stream);
  return launch_fwd_f<N, TS, G, CS, false, XIN>(p, stream);
}
template <int N, bool TS, bool K16, bool FAST>
static int launch_bwd_f(const tc::LstmClusterParams& p, cudaStream_t stream) {
  constexpr int smem = tc::bwd_smem_bytes<N, TS>();
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(tc::lstm_bwd_cluster_kernel<N, TS, K16, FAST>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return cuda_status(e, "lstm_bwd_cluster smem attribute");
    if (g_lstm_carveout)   // see launch_fwd_f
      cudaFuncSetAttribute(tc::lstm_bwd_cluster_kernel<N, TS, K16, FAST>, cudaFuncAttributePreferredSharedMemoryCarveout,
                           cudaSharedmemCarveoutMaxShared);
    attr = true;
  }
  DEER_LAUNCH((tc::lstm_bwd_cluster_kernel<N, TS, K16, FAST>), tc::QC * p.ntiles * 2, tc::QTHREADS, smem, stream, p);
  return DEER_OK;
}
template <int N, bool TS, bool K16>
static int launch_bwd(const tc::LstmClusterParams& p, cudaStream_t stream) {
  if (p.B % N == 0 && p.prof == nullptr) return launch_bwd_f<N, TS, K16, true>(p, stream);
  return launch_bwd_f<N, TS, K16, false>(p, stream);
}

int lstm_cluster_tile(int B) { return pick_tile(B); }

int lstm_fwd_cluster(const float* pre_il, const void* pre_f16, const float* w_fwd, const float* w_rev, float* h_out,
                     float* gact, float* c_blk, void* h16, void* hb16, int T, int B, cudaStream_t stream) {
  int N = pick_tile(B);
  const int keep = (gact != nullptr && c_blk != nullptr) ? 1 : 0;
  // experiment (DEER_OPT_LSTM_DUAL = 2): training forward of a one-wave batch on HALF the SMs, as dual sub-tiles of 32
  // columns per CTA; the kept layouts are the 16-column kernel's when ceil(B/16) == 2 ceil(B/32)
  const bool dual_keep = g_lstm_dual == 2 && keep && N == 16 && g_lstm_ts && (B + 15) / 16 == 2 * ((B + 31) / 32);
  if (dual_keep) N = 32;
  tc::LstmClusterParams p{const_cast<float*>(pre_il), w_fwd, w_rev, h_out, gact, c_blk, nullptr, nullptr,
                          reinterpret_cast<__half*>(h16), reinterpret_cast<__nv_bfloat16*>(hb16), nullptr,
                          reinterpret_cast<const __half*>(pre_f16), T, B, (B + N - 1) / N, keep, g_lstm_prof,
                          g_lstm_keep16, g_lstm_stasync};
  if (N == 16) {
    if (g_lstm_ts && g_lstm_halfsplit) return launch_fwd<16, true, 2, 2>(p, stream);
    if (g_lstm_ts && g_lstm_colsplit && !g_lstm_keep16) return launch_fwd<16, true, 1, 2>(p, stream);
    return g_lstm_ts ? launch_fwd<16, true>(p, stream) : launch_fwd<16, false>(p, stream);
  }
  // inference (nothing kept for BPTT): the 32 batch columns of a CTA run as two interleaved 16-column sub-tiles.  The
  // kept gate / cell layouts are those of the 32-column backward kernel, so training keeps the monolithic tile.
  if ((!keep || dual_keep) && g_lstm_ts && g_lstm_dual) return launch_fwd<16, true, 2>(p, stream);
  return g_lstm_ts ? launch_fwd<32, true>(p, stream) : launch_fwd<32, false>(p, stream);
}

// which XIN kernel (input projection inside the recurrence) serves a batch: 1 = 16-column tiles (one wave, B <= 256),
// 0 = none (the projection stays a GEMM).  Measured (tools/xin_probe.py, layer-0 forward incl. operand preparation):
// B = 256 training 552 -> 458 us, B = 256 inference 552 -> 410 us; the dual-sub-tile inference kernel of larger batches
// LOSES (B = 1024: 1342 -> 1637 us: its single MMA warp already serves two recurrences and the input part lands on their
// critical path), so it keeps the GEMM.
int g_lstm_xin = 1;            // DEER_OPT_LSTM_XIN
int g_lstm_xin_dual = 1;       // (DEER_OPT_LSTM_XIN = 3: in-kernel projection for 16-column tiles only)
int lstm_cluster_xin_mode(int B, int keep, int xk) {
  if (!g_lstm_xin || !g_lstm_ts || g_lstm_halfsplit || (g_lstm_colsplit && !g_lstm_keep16) || xk <= 0 || xk > 128 || (xk & 7))
    return 0;
  if (pick_tile(B) != 16) return (!keep && g_lstm_dual && g_lstm_xin_dual) ? 2 : 0;
  return (g_lstm_dual == 2 && keep) ? 0 : 1;
}
int lstm_fwd_cluster_xin(const void* x16, int xk, const void* wih16, const float* bias_il, const float* w_fwd,
                         const float* w_rev, float* h_out, float* gact, float* c_blk, void* h16, void* hb16, int T, int B,
                         cudaStream_t stream) {
  const int keep = (gact != nullptr && c_blk != nullptr) ? 1 : 0;
  const int mode = lstm_cluster_xin_mode(B, keep, xk);
  if (mode == 0) {
    set_error("lstm_cluster_fwd_xin: no in-kernel projection variant for B=%d keep=%d xk=%d", B, keep, xk);
    return DEER_ERR_UNSUPPORTED;
  }
  const int N = pick_tile(B);
  tc::LstmClusterParams p{nullptr, w_fwd, w_rev, h_out, gact, c_blk, nullptr, nullptr,
                          reinterpret_cast<__half*>(h16), reinterpret_cast<__nv_bfloat16*>(hb16), nullptr,
                          nullptr, T, B, (B + N - 1) / N, keep, g_lstm_prof, g_lstm_keep16, g_lstm_stasync,
                          reinterpret_cast<const __half*>(x16), reinterpret_cast<const __half*>(wih16), bias_il, xk};
  if (mode == 1) return launch_fwd<16, true, 1, 1, true>(p, stream);
  return launch_fwd<16, true, 2, 1, true>(p, stream);
}

int lstm_bwd_cluster(const float* gact, const float* c_blk, const float* dh_out, const float* w_fwd, const float* w_rev,
                     float* dpre_il, float* db_il, void* dpre16, int T, int B, cudaStream_t stream) {
  const int N = pick_tile(B);
  tc::LstmClusterParams p{dpre_il, w_fwd, w_rev, nullptr, const_cast<float*>(gact), const_cast<float*>(c_blk), dh_out,
                          db_il, nullptr, nullptr, reinterpret_cast<__nv_bfloat16*>(dpre16), nullptr, T, B,
                          (B + N - 1) / N, 1, g_lstm_prof, g_lstm_keep16, 0};
  if (g_lstm_keep16) {
    if (N == 16) return g_lstm_ts ? launch_bwd<16, true, true>(p, stream) : launch_bwd<16, false, true>(p, stream);
    return launch_bwd<32, true, true>(p, stream);
  }
  if (N == 16) return g_lstm_ts ? launch_bwd<16, true, false>(p, stream) : launch_bwd<16, false, false>(p, stream);
  return launch_bwd<32, true, false>(p, stream);  // the N=32 tiles + 128 KB of smem-resident weights exceed 227 KB
}

}  // namespace deer

using namespace deer;

extern "C" {

int deer_lstm_cluster_tile(int B) { return B > 0 ? lstm_cluster_tile(B) : DEER_ERR_INVALID; }

int deer_lstm_cluster_fwd(const float* pre_il, const float* w_hh_fwd, const float* w_hh_rev, float* h_out, void* gact,
                          void* c_blk, void* h_f16, void* h_bf16, int T, int B, int H, void* stream) {
  DEER_CHECK_ARG(pre_il && w_hh_fwd && w_hh_rev && h_out && T > 0 && B > 0, "lstm_cluster_fwd: bad args");
  DEER_CHECK_ARG((gact == nullptr) == (c_blk == nullptr), "lstm_cluster_fwd: gact and c_blk go together");
  DEER_CHECK_ARG(((reinterpret_cast<uintptr_t>(gact) | reinterpret_cast<uintptr_t>(c_blk)) & 15) == 0,
                 "lstm_cluster_fwd: gact / c_blk must be 16-byte aligned");
  if (!lstm_cluster_supported(pre_il, w_hh_fwd, w_hh_rev, H)) {
    set_error("lstm_cluster_fwd: needs H == 256 and 16-byte aligned pointers");
    return DEER_ERR_UNSUPPORTED;
  }
  DEER_CHECK_ARG(((reinterpret_cast<uintptr_t>(h_f16) | reinterpret_cast<uintptr_t>(h_bf16)) & 15) == 0,
                 "lstm_cluster_fwd: 16-bit shadows must be 16-byte aligned");
  return lstm_fwd_cluster(pre_il, nullptr, w_hh_fwd, w_hh_rev, h_out, reinterpret_cast<float*>(gact),
                          reinterpret_cast<float*>(c_blk), h_f16, h_bf16, T, B, (cudaStream_t)stream);
}

int deer_lstm_cluster_fwd_pre16(const void* pre_il_f16, const float* w_hh_fwd, const float* w_hh_rev, float* h_out,
                                void* gact, void* c_blk, void* h_f16, void* h_bf16, int T, int B, int H,
                                void* stream) {
  DEER_CHECK_ARG(pre_il_f16 && w_hh_fwd && w_hh_rev && h_out && T > 0 && B > 0, "lstm_cluster_fwd_pre16: bad args");
  DEER_CHECK_ARG((gact == nullptr) == (c_blk == nullptr), "lstm_cluster_fwd_pre16: gact and c_blk go together");
  if (!lstm_cluster_supported(h_out, w_hh_fwd, w_hh_rev, H)) {
    set_error("lstm_cluster_fwd_pre16: needs H == 256 and 16-byte aligned pointers");
    return DEER_ERR_UNSUPPORTED;
  }
  DEER_CHECK_ARG(((reinterpret_cast<uintptr_t>(h_f16) | reinterpret_cast<uintptr_t>(h_bf16) |
                   reinterpret_cast<uintptr_t>(pre_il_f16)) & 15) == 0,
                 "lstm_cluster_fwd_pre16: 16-bit buffers must be 16-byte aligned");
  return lstm_fwd_cluster(nullptr, pre_il_f16, w_hh_fwd, w_hh_rev, h_out, reinterpret_cast<float*>(gact),
                          reinterpret_cast<float*>(c_blk), h_f16, h_bf16, T, B, (cudaStream_t)stream);
}

int deer_lstm_cluster_xin_mode(int B, int keep, int xk) { return B > 0 ? lstm_cluster_xin_mode(B, keep, xk) : DEER_ERR_INVALID; }

int deer_lstm_cluster_fwd_xin(const void* x_f16, int xk, const void* w_ih_il_f16, const float* bias_il, const float* w_hh_fwd,
                              const float* w_hh_rev, float* h_out, void* gact, void* c_blk, void* h_f16, void* h_bf16, int T,
                              int B, int H, void* stream) {
  DEER_CHECK_ARG(x_f16 && w_ih_il_f16 && bias_il && w_hh_fwd && w_hh_rev && h_out && T > 0 && B > 0,
                 "lstm_cluster_fwd_xin: bad args");
  DEER_CHECK_ARG((gact == nullptr) == (c_blk == nullptr), "lstm_cluster_fwd_xin: gact and c_blk go together");
  if (!lstm_cluster_supported(h_out, w_hh_fwd, w_hh_rev, H)) {
    set_error("lstm_cluster_fwd_xin: needs H == 256 and 16-byte aligned pointers");
    return DEER_ERR_UNSUPPORTED;
  }
  DEER_CHECK_ARG(((reinterpret_cast<uintptr_t>(h_f16) | reinterpret_cast<uintptr_t>(h_bf16) | reinterpret_cast<uintptr_t>(x_f16) |
                   reinterpret_cast<uintptr_t>(w_ih_il_f16) | reinterpret_cast<uintptr_t>(gact) |
                   reinterpret_cast<uintptr_t>(c_blk)) & 15) == 0,
                 "lstm_cluster_fwd_xin: buffers must be 16-byte aligned");
  return lstm_fwd_cluster_xin(x_f16, xk, w_ih_il_f16, bias_il, w_hh_fwd, w_hh_rev, h_out, reinterpret_cast<float*>(gact),
                              reinterpret_cast<float*>(c_blk), h_f16, h_bf16, T, B, (cudaStream_t)stream);
}

int deer_lstm_cluster_bwd(const void* gact, const void* c_blk, const float* dh_out, const float* w_hh_fwd,
                          const float* w_hh_rev, float* dpre_il, float* db_il, void* dpre_bf16, int T, int B, int H,
                          void* stream) {
  DEER_CHECK_ARG(gact && c_blk && dh_out && w_hh_fwd && w_hh_rev && (dpre_il || dpre_bf16) && T > 0 && B > 0,
                 "lstm_cluster_bwd: bad args");
  if (!lstm_cluster_supported(dpre_il, w_hh_fwd, w_hh_rev, H)) {
    set_error("lstm_cluster_bwd: needs H == 256 and 16-byte aligned pointers");
    return DEER_ERR_UNSUPPORTED;
  }
  DEER_CHECK_ARG((reinterpret_cast<uintptr_t>(dpre_bf16) & 15) == 0, "lstm_cluster_bwd: dpre_bf16 must be 16-byte aligned");
  return lstm_bwd_cluster(reinterpret_cast<const float*>(gact), reinterpret_cast<const float*>(c_blk), dh_out, w_hh_fwd,
                          w_hh_rev, dpre_il, db_il, dpre_bf16, T, B, (cudaStream_t)stream);
}

}  // extern "C"
