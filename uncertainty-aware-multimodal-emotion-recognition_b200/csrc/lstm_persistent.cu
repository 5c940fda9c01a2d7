// Persistent bidirectional-LSTM forward recurrence for H = 256 (encoders.py:82-89: nn.LSTM(84|512 -> 256, bidirectional)).
//
// One thread-block CLUSTER of 8 CTAs owns one (direction, 64-row batch tile) recurrence for all T steps:
//   * CTA r keeps the recurrent weights of ITS 32 hidden units (4 gates x 32 rows x 256 k, TF32, 128 KB) resident in
//     shared memory for the whole sequence (loaded once by TMA), as the B operand of tcgen05.mma;
//   * every step it computes gates[64, 4x32] = pre_t + h_{t-1}[64,256] W_r^T with 32 tcgen05.mma (M64 N128 K8,
//     kind::tf32) accumulating in TMEM.  The accumulator is PRE-LOADED with the time-batched input projection pre_t
//     (tcgen05.st) while the h exchange is in flight, so the critical path holds only the recurrent part;
//   * the gate-nonlinearity epilogue (one batch row per thread: tcgen05.ld -> sigmoid/tanh on the MUFU pipe -> cell
//     state kept in REGISTERS across all T steps) writes h_t (and, for training, the activated gates and c_t) straight
//     to the time-major outputs with 128-bit stores;
//   * h_t is all-gathered across the cluster through L2: each CTA TMA-loads its own freshly written [64 x 32] slice
//     with .multicast::cluster into the K-major, 128B-swizzled A-operand buffer of ALL 8 CTAs (one L2 read per slice,
//     TF32 rounding done by the TMA unit); mbarrier transaction counts gate the next step's MMA; tcgen05.commit
//     multicast tells every producer when the h buffer may be overwritten.  No per-step kernel launch, no grid sync.
// Grid = 8 x ceil(B/64) x 2 directions CTAs (B=256: 64 SMs, B=1024: 256 CTAs in two waves).
#include "tc_ptx.cuh"

namespace deer {
namespace tc {

constexpr int LH = 256;                 // hidden size
constexpr int LC = 8;                   // cluster size
constexpr int LHS = LH / LC;            // hidden units per CTA (32)
constexpr int LBT = 64;                 // batch rows per cluster (UMMA M)
constexpr int LN = 4 * LHS;             // gate columns per CTA (UMMA N = 128)
constexpr int LW_BYTES = LN * LH * 4;   // 128 KB
constexpr int LHB_BYTES = LBT * LH * 4; // 64 KB
constexpr int LSMEM = LW_BYTES + LHB_BYTES + 1024 + 256;
constexpr int LTHREADS = 160;           // warp 0: MMA/TMA, warps 1..4: epilogue

struct LstmParams {
  float* gates;   // [T,B,2,4H] in: pre-activations, out (keep): activated gates
  float* h_out;   // [T,B,2H]
  float* c_all;   // [T,B,2,H] or null
  int T, B, ntiles, keep;
};

__global__ void __cluster_dims__(LC, 1, 1) __launch_bounds__(LTHREADS, 1)
    lstm_fwd_persistent_kernel(const __grid_constant__ CUtensorMap tmap_w_fwd,
                               const __grid_constant__ CUtensorMap tmap_w_rev,
                               const __grid_constant__ CUtensorMap tmap_h, const LstmParams p) {
  DEER_PDL_ENTRY();
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms, applied as an OFFSET on the __shared__ array: going through
  // uintptr_t would make every later access a generic LD/ST instead of LDS/STS
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* wbuf = smem;                       // 8 K-blocks x [128 n-rows x 128 B]
  uint8_t* hbuf = smem + LW_BYTES;            // 8 K-blocks x [64 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + LW_BYTES + LHB_BYTES);
  uint64_t* w_full = bars + 0;
  uint64_t* h_full = bars + 1;
  uint64_t* acc_ready = bars + 2;
  uint64_t* mma_done = bars + 3;
  uint64_t* h_free = bars + 4;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t r = cluster_ctarank();
  const int cid = blockIdx.x / LC;
  const int tile = cid % p.ntiles, dir = cid / p.ntiles;
  const int b0 = tile * LBT;
  const int T = p.T, B = p.B;
  const CUtensorMap* tmap_w = dir ? &tmap_w_rev : &tmap_w_fwd;

  // ---- one-time setup
  for (int i = threadIdx.x; i < LHB_BYTES / 16; i += LTHREADS) reinterpret_cast<float4*>(hbuf)[i] = make_float4(0, 0, 0, 0);
  fence_proxy_async_smem();  // zeroed h_{-1} visible to the tensor core (async proxy)
  if (threadIdx.x == 0) {
    prefetch_tmap(tmap_w);
    prefetch_tmap(&tmap_h);
    mbar_init(w_full, 1);
    mbar_init(h_full, 1);
    mbar_init(acc_ready, 4);
    mbar_init(mma_done, 1);
    mbar_init(h_free, LC);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers exist before any peer multicasts into them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =================================================================== MMA issuer (+ one-time weight load)
    if (lane == 0) {
      mbar_expect_tx(w_full, LW_BYTES);
      for (int kb = 0; kb < 8; kb++)
        for (int g = 0; g < 4; g++)  // box {32 k, 32 rows}: rows g*256 + r*32 .. +32 of W_hh, k block kb
          tma_load_2d(wbuf + kb * 16384 + g * 4096, tmap_w, w_full, kb * 32, g * LH + (int)r * LHS);
      constexpr uint32_t idesc = make_idesc(0, 0, LBT, LN);
      mbar_wait(w_full, 0);
      for (int s = 0; s < T; s++) {
        if (s > 0) mbar_wait(h_full, (uint32_t)((s - 1) & 1));
        mbar_wait(acc_ready, (uint32_t)(s & 1));
        tc_fence_after();
        const uint32_t ha = smem_u32(hbuf), wa = smem_u32(wbuf);
#pragma unroll 1
        for (int kb = 0; kb < 8; kb++) {
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const uint64_t ad = make_smem_desc(ha + kb * 8192 + k * 32, 16, 1024, 2);
            const uint64_t bd = make_smem_desc(wa + kb * 16384 + k * 32, 16, 1024, 2);
            umma_tf32(tmem_base, ad, bd, idesc, 1u);  // accumulate onto the pre-loaded input projection
          }
        }
        umma_commit(mma_done);
        umma_commit_mc(h_free, (uint16_t)0xFF);  // every CTA learns that this CTA no longer reads h_{s-1}
      }
    }
  } else {
    // =================================================================== epilogue: one batch row per thread
    const int q = warp & 3;                  // TMEM sub-partition of this warp
    const int m = q * 16 + lane;             // batch row inside the tile (lanes 16..31 idle: M=64 uses 16 lanes/subpart.)
    const bool active = lane < 16;
    const int b = b0 + m;
    const bool valid = active && b < B;
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
    float c_state[LHS];
#pragma unroll
    for (int i = 0; i < LHS; i++) c_state[i] = 0.f;

    auto preload = [&](int s) {
      const int t = dir ? T - 1 - s : s;
      const float* src = p.gates + (((long long)t * B + b) * 2 + dir) * (4 * LH) + r * LHS;
#pragma unroll 1
      for (int g = 0; g < 4; g++) {
        float v[32];
        if (valid) {
#pragma unroll
          for (int j = 0; j < 8; j++) {
            const float4 x = __ldcs(reinterpret_cast<const float4*>(src + g * LH) + j);
            v[4 * j] = x.x; v[4 * j + 1] = x.y; v[4 * j + 2] = x.z; v[4 * j + 3] = x.w;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; j++) v[j] = 0.f;
        }
        tmem_st32(taddr + (uint32_t)(g * 32), v);
      }
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_ready);
    };

    preload(0);
    for (int s = 0; s < T; s++) {
      const int t = dir ? T - 1 - s : s;
      mbar_wait(mma_done, (uint32_t)(s & 1));
      tc_fence_after();
      const long long row = (long long)t * B + b;
      float* gdst = p.gates + (row * 2 + dir) * (4 * LH) + r * LHS;
      float* cdst = p.c_all ? p.c_all + (row * 2 + dir) * LH + r * LHS : nullptr;
      float* hdst = p.h_out + row * (2 * LH) + dir * LH + r * LHS;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        float gi[8], gf[8], gg[8], go[8];
        tmem_ld8(taddr + (uint32_t)(0 * 32 + c * 8), gi);
        tmem_ld8(taddr + (uint32_t)(1 * 32 + c * 8), gf);
        tmem_ld8(taddr + (uint32_t)(2 * 32 + c * 8), gg);
        tmem_ld8(taddr + (uint32_t)(3 * 32 + c * 8), go);
        tmem_ld_wait();
        float hv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
          gi[u] = sigmoid_f(gi[u]);
          gf[u] = sigmoid_f(gf[u]);
          gg[u] = tanh_f(gg[u]);
          go[u] = sigmoid_f(go[u]);
          const float cn = fmaf(gf[u], c_state[c * 8 + u], gi[u] * gg[u]);
          c_state[c * 8 + u] = cn;
          hv[u] = go[u] * tanh_f(cn);
        }
        if (valid) {
          float4* h4 = reinterpret_cast<float4*>(hdst + c * 8);
          h4[0] = make_float4(hv[0], hv[1], hv[2], hv[3]);
          h4[1] = make_float4(hv[4], hv[5], hv[6], hv[7]);
          if (p.keep) {
            float4* d;
            d = reinterpret_cast<float4*>(gdst + 0 * LH + c * 8);
            d[0] = make_float4(gi[0], gi[1], gi[2], gi[3]); d[1] = make_float4(gi[4], gi[5], gi[6], gi[7]);
            d = reinterpret_cast<float4*>(gdst + 1 * LH + c * 8);
            d[0] = make_float4(gf[0], gf[1], gf[2], gf[3]); d[1] = make_float4(gf[4], gf[5], gf[6], gf[7]);
            d = reinterpret_cast<float4*>(gdst + 2 * LH + c * 8);
            d[0] = make_float4(gg[0], gg[1], gg[2], gg[3]); d[1] = make_float4(gg[4], gg[5], gg[6], gg[7]);
            d = reinterpret_cast<float4*>(gdst + 3 * LH + c * 8);
            d[0] = make_float4(go[0], go[1], go[2], go[3]); d[1] = make_float4(go[4], go[5], go[6], go[7]);
            if (cdst) {
              float4* c4 = reinterpret_cast<float4*>(cdst + c * 8);
              c4[0] = make_float4(c_state[c * 8 + 0], c_state[c * 8 + 1], c_state[c * 8 + 2], c_state[c * 8 + 3]);
              c4[1] = make_float4(c_state[c * 8 + 4], c_state[c * 8 + 5], c_state[c * 8 + 6], c_state[c * 8 + 7]);
            }
          }
        }
      }
      if (s + 1 < T) {
        // publish h_t: device-scope fence by every writer, CTA-wide rendezvous of the epilogue warps, then ONE thread
        // all-gathers this CTA's slice into all 8 CTAs (TMA multicast through L2).
        __threadfence();
        named_bar_sync(1, 128);
        if (warp == 1 && lane == 0) {
          fence_proxy_async_all();
          mbar_wait(h_free, (uint32_t)(s & 1));            // all 8 CTAs finished reading h_{t-1}
          mbar_expect_tx(h_full, LHB_BYTES);               // this CTA expects 8 slices x 8 KB
          tma_load_2d_mc(hbuf + r * 8192, &tmap_h, h_full, dir * LH + (int)r * LHS, t * B + b0, (uint16_t)0xFF);
        }
        preload(s + 1);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // peers may still signal this CTA's barriers
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 128);
  }
}

}  // namespace tc

bool lstm_persistent_supported(const float* gates, const float* h_out, const float* c_all, const float* w_fwd,
                               const float* w_rev, int T, int B, int H) {
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  (void)T;
  (void)B;
  return H == tc::LH && al16(gates) && al16(h_out) && (c_all == nullptr || al16(c_all)) && al16(w_fwd) && al16(w_rev);
}

int lstm_fwd_persistent(float* gates, const float* w_fwd, const float* w_rev, float* h_out, float* c_all, int T, int B,
                        int keep, cudaStream_t stream) {
  using namespace tc;
  CUtensorMap mwf, mwr, mh;
  bool ok = make_map(&mwf, w_fwd, 4 * LH, LH, LH, 32, 32, false) && make_map(&mwr, w_rev, 4 * LH, LH, LH, 32, 32, false) &&
            make_map(&mh, h_out, (long long)T * B, 2 * LH, 2 * LH, 32, LBT, false);
  if (!ok) {
    set_error("lstm_fwd_persistent: tensor map creation failed");
    return DEER_ERR_UNSUPPORTED;
  }
  static bool attr = false;
  if (!attr) {
    cudaError_t e = cudaFuncSetAttribute(lstm_fwd_persistent_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, LSMEM);
    if (e != cudaSuccess) return cuda_status(e, "lstm_fwd_persistent smem attribute");
    attr = true;
  }
  const int ntiles = (B + LBT - 1) / LBT;
  LstmParams p{gates, h_out, c_all, T, B, ntiles, keep};
  DEER_LAUNCH(lstm_fwd_persistent_kernel, LC * ntiles * 2, LTHREADS, LSMEM, stream, mwf, mwr, mh, p);
  return DEER_OK;
}

}  // namespace deer
