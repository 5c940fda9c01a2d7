// Device side of the fused 3xTF32 GEMM engine (see gemm_x3.cu): argument block, tile configuration and the CTA-wide tile
// routine, shared by the stand-alone kernel (gemm_x3.cu) and the persistent chain kernel (chain.cu).
#pragma once
#include "common.cuh"

namespace deer {
namespace x3 {

using Args = deer_gemm_x3_args;   // include/deer_b200.h (splitk: filled in by the dispatcher)

__device__ __forceinline__ void cp16(float* dst, const float* src, int bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp4(float* dst, const float* src, int bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
  const float r = x - __uint_as_float(hi);
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float* d, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ float gate_fn(float g, int mode, float scale) {
  switch (mode) {
    case 1: return g > 0.f ? scale : 0.f;        // ReLU (and inverted-dropout) mask from the saved output
    case 2: return (1.f - g * g) * scale;        // tanh'
    case 3: return g * (1.f - g) * scale;        // sigmoid'
    default: return scale;
  }
}

constexpr int STAGES = 3;

template <int BM, int BN, int WM, int WN, int WK>
struct Cfg {
  static constexpr int NT = 32 * WM * WN * WK;
  static constexpr int KS = 32 * WK;                    // k per stage
  static constexpr int MT = BM / WM / 16;               // m16 tiles per warp
  static constexpr int NTL = BN / WN / 8;               // n8 tiles per warp
  static constexpr int A_TILE = (BM * (KS + 4) > KS * (BM + 8)) ? BM * (KS + 4) : KS * (BM + 8);
  static constexpr int B_TILE = (BN * (KS + 4) > KS * (BN + 8)) ? BN * (KS + 4) : KS * (BN + 8);
  static constexpr int RED = WK * BM * (BN + 4);
  static constexpr int RING0 = STAGES * (A_TILE + B_TILE), RING1 = STAGES * (2 * A_TILE + B_TILE);
  static constexpr int SMEM0 = (RING0 > RED ? RING0 : RED) + BM;   // floats, no gate operand (+ column-sum scratch)
  static constexpr int SMEM1 = (RING1 > RED ? RING1 : RED) + BM;   // with the gate operand
};

// rows x KS operand tile, global -> shared.  KC: rows are k-contiguous in global memory (tile stored [rows][KS+4]);
// otherwise the operand is k-major (tile stored [KS][rows+8]).  VEC: 16-byte copies (aligned operand), else 4-byte.
template <int ROWS, int KS, int NT, bool KC, bool VEC>
__device__ __forceinline__ void load_tile(float* dst, const float* __restrict__ src, long long ld, int r0, int rmax,
                                          int k0, int kend, int tid) {
  constexpr int LDK = KS + 4, LDR = ROWS + 8;
  if constexpr (VEC) {
    constexpr int CNT = ROWS * KS / 4;
#pragma unroll
    for (int idx0 = 0; idx0 < CNT; idx0 += NT) {
      const int idx = idx0 + tid;
      if (CNT % NT != 0 && idx >= CNT) break;
      if constexpr (KC) {
        const int r = idx / (KS / 4), c4 = (idx % (KS / 4)) * 4;
        const int gr = r0 + r, gk = k0 + c4;
        const int lim = gr < rmax ? kend - gk : 0;
        const int bytes = max(0, min(4, lim)) * 4;
        cp16(dst + r * LDK + c4, bytes ? src + (long long)gr * ld + gk : src, bytes);
      } else {
        const int r = idx / (ROWS / 4), c4 = (idx % (ROWS / 4)) * 4;
        const int gk = k0 + r, gr = r0 + c4;
        const int lim = gk < kend ? rmax - gr : 0;
        const int bytes = max(0, min(4, lim)) * 4;
        cp16(dst + r * LDR + c4, bytes ? src + (long long)gk * ld + gr : src, bytes);
      }
    }
  } else {
    constexpr int CNT = ROWS * KS;
#pragma unroll 4
    for (int idx = tid; idx < CNT; idx += NT) {
      if constexpr (KC) {
        const int r = idx / KS, c = idx % KS;
        const int gr = r0 + r, gk = k0 + c;
        const bool ok = gr < rmax && gk < kend;
        cp4(dst + r * LDK + c, ok ? src + (long long)gr * ld + gk : src, ok ? 4 : 0);
      } else {
        const int r = idx / ROWS, c = idx % ROWS;
        const int gk = k0 + r, gr = r0 + c;
        const bool ok = gr < rmax && gk < kend;
        cp4(dst + r * LDR + c, ok ? src + (long long)gk * ld + gr : src, ok ? 4 : 0);
      }
    }
  }
}

// One BM x BN output tile (tile coordinates bx = n-tile, by = m-tile, bz = batch * splitk + split) by the whole CTA.
// Callable from a persistent kernel: ends with every thread past its last shared-memory access EXCEPT the epilogue's
// reads of the reduction buffer -- callers that reuse `smem` must __syncthreads() first.
template <int BM, int BN, int WM, int WN, int WK, bool AK, bool BKC, bool GATE, bool VEC>
__device__ __forceinline__ void x3_tile(const Args& p, const int bx, const int by, const int bz, float* smem) {
  using C_ = Cfg<BM, BN, WM, WN, WK>;
  constexpr int NT = C_::NT, KS = C_::KS, MT = C_::MT, NTL = C_::NTL;
  constexpr int LDAK = KS + 4, LDAM = BM + 8, LDBK = KS + 4, LDBN = BN + 8;
  float* sAb = smem;                                     // [STAGES][A_TILE]
  float* sBb = sAb + STAGES * C_::A_TILE;                // [STAGES][B_TILE]
  float* sGb = sBb + STAGES * C_::B_TILE;                // [STAGES][A_TILE] (GATE)
  float* scol = smem + (GATE ? C_::SMEM1 : C_::SMEM0) - BM;       // [BM] column sums

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wk = warp % WK, wn = (warp / WK) % WN, wm = warp / (WK * WN);
  const int z = bz, batch = z / p.splitk, ks = z % p.splitk;
  const float* __restrict__ A = p.A + batch * p.sA;
  const float* __restrict__ B = p.B + batch * p.sB;
  const float* __restrict__ G = GATE ? p.gate + batch * p.sGate : nullptr;
  float* __restrict__ C = p.C + batch * p.sC;
  const float* __restrict__ bias = p.bias ? p.bias + batch * p.sBias : nullptr;
  const int M = p.M, N = p.N, K = p.K;
  const int m0 = by * BM, n0 = bx * BN;
  // K range of this split in whole stages
  const int nstage_all = (K + KS - 1) / KS;
  const int per = (nstage_all + p.splitk - 1) / p.splitk;
  const int kbeg = min(K, ks * per * KS), kend = min(K, (ks + 1) * per * KS);
  const int nst = (kend - kbeg + KS - 1) / KS;
  const bool want_col = p.colsum != nullptr && bx == 0;
  const int gate_mode = p.gate_mode;
  const float gate_scale = p.gate_scale;

  auto issue = [&](int st) {
    const int slot = st % STAGES, k0 = kbeg + st * KS;
    load_tile<BM, KS, NT, AK, VEC>(sAb + slot * C_::A_TILE, A, p.lda, m0, M, k0, kend, tid);
    load_tile<BN, KS, NT, BKC, VEC>(sBb + slot * C_::B_TILE, B, p.ldb, n0, N, k0, kend, tid);
    if constexpr (GATE) load_tile<BM, KS, NT, AK, VEC>(sGb + slot * C_::A_TILE, G, p.ldgate, m0, M, k0, kend, tid);
  };

  float acc[MT][NTL][4];
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NTL; j++)
#pragma unroll
      for (int q = 0; q < 4; q++) acc[i][j][q] = 0.f;
  float csum = 0.f;   // this thread's share of one column sum (want_col)

#pragma unroll
  for (int st = 0; st < STAGES - 1; st++) {
    if (st < nst) issue(st);
    cp_commit();
  }
  const int mw = wm * (BM / WM), nw = wn * (BN / WN), kw = wk * 32;
  for (int st = 0; st < nst; st++) {
    cp_wait<STAGES - 2>();
    __syncthreads();              // stage st has landed for everyone; slot (st-1) % STAGES is free again
    if (st + STAGES - 1 < nst) issue(st + STAGES - 1);
    cp_commit();
    const float* tA = sAb + (st % STAGES) * C_::A_TILE;
    const float* tB = sBb + (st % STAGES) * C_::B_TILE;
    const float* tG = sGb + (st % STAGES) * C_::A_TILE;
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      const int kc = kw + kk * 8 + t;
      uint32_t ah[MT][4], al[MT][4], bh[NTL][2], bl[NTL][2];
#pragma unroll
      for (int i = 0; i < MT; i++) {
        const int r = mw + i * 16 + g;
        float a0, a1, a2, a3;
        if constexpr (AK) {
          a0 = tA[r * LDAK + kc]; a1 = tA[(r + 8) * LDAK + kc]; a2 = tA[r * LDAK + kc + 4]; a3 = tA[(r + 8) * LDAK + kc + 4];
          if constexpr (GATE) {
            a0 *= gate_fn(tG[r * LDAK + kc], gate_mode, gate_scale);
            a1 *= gate_fn(tG[(r + 8) * LDAK + kc], gate_mode, gate_scale);
            a2 *= gate_fn(tG[r * LDAK + kc + 4], gate_mode, gate_scale);
            a3 *= gate_fn(tG[(r + 8) * LDAK + kc + 4], gate_mode, gate_scale);
          }
        } else {
          a0 = tA[kc * LDAM + r]; a1 = tA[kc * LDAM + r + 8]; a2 = tA[(kc + 4) * LDAM + r]; a3 = tA[(kc + 4) * LDAM + r + 8];
          if constexpr (GATE) {
            a0 *= gate_fn(tG[kc * LDAM + r], gate_mode, gate_scale);
            a1 *= gate_fn(tG[kc * LDAM + r + 8], gate_mode, gate_scale);
            a2 *= gate_fn(tG[(kc + 4) * LDAM + r], gate_mode, gate_scale);
            a3 *= gate_fn(tG[(kc + 4) * LDAM + r + 8], gate_mode, gate_scale);
          }
        }
        split_tf32(a0, ah[i][0], al[i][0]);
        split_tf32(a1, ah[i][1], al[i][1]);
        split_tf32(a2, ah[i][2], al[i][2]);
        split_tf32(a3, ah[i][3], al[i][3]);
      }
#pragma unroll
      for (int j = 0; j < NTL; j++) {
        const int n = nw + j * 8 + g;
        float b0, b1;
        if constexpr (BKC) {
          b0 = tB[n * LDBK + kc]; b1 = tB[n * LDBK + kc + 4];
        } else {
          b0 = tB[kc * LDBN + n]; b1 = tB[(kc + 4) * LDBN + n];
        }
        split_tf32(b0, bh[j][0], bl[j][0]);
        split_tf32(b1, bh[j][1], bl[j][1]);
      }
#pragma unroll
      for (int i = 0; i < MT; i++)
#pragma unroll
        for (int j = 0; j < NTL; j++) {
          mma_tf32(acc[i][j], al[i], bh[j]);   // small terms first
          mma_tf32(acc[i][j], ah[i], bl[j]);
          mma_tf32(acc[i][j], ah[i], bh[j]);
        }
    }
    if constexpr (!AK) {
      if (want_col) {   // bias gradient: column sums of A_eff over this stage's k rows (tile stored [k][m])
        constexpr int KQ = NT / BM, ROWS_PER = KS / KQ;
        const int m = tid % BM, kq = tid / BM;
#pragma unroll 4
        for (int r = kq * ROWS_PER; r < (kq + 1) * ROWS_PER; r++) {
          float v = tA[r * LDAM + m];
          if constexpr (GATE) v *= gate_fn(tG[r * LDAM + m], gate_mode, gate_scale);
          csum += v;
        }
      }
    }
  }
  cp_wait<0>();
  __syncthreads();   // every warp is done with the operand ring: reuse it for the k-group reduction
  float* red = smem;   // [WK][BM][BN+4]
  constexpr int LDR = BN + 4;
#pragma unroll
  for (int i = 0; i < MT; i++)
#pragma unroll
    for (int j = 0; j < NTL; j++) {
      const int r = mw + i * 16 + g, c = nw + j * 8 + 2 * t;
      *reinterpret_cast<float2*>(red + (wk * BM + r) * LDR + c) = make_float2(acc[i][j][0], acc[i][j][1]);
      *reinterpret_cast<float2*>(red + (wk * BM + r + 8) * LDR + c) = make_float2(acc[i][j][2], acc[i][j][3]);
    }
  if (want_col && tid < BM) scol[tid] = 0.f;
  __syncthreads();
  if constexpr (!AK) {
    if (want_col) {
      atomicAdd(scol + tid % BM, csum);
    }
  }
  // ---- epilogue: combine the k-groups, bias / beta / activation / dropout, coalesced 128-bit stores
  const bool split = p.splitk > 1;
  const bool vec_c = ((reinterpret_cast<uintptr_t>(C) & 15) == 0) && (p.ldc % 4 == 0);
  const bool drop = p.drop_p > 0.f;
  const unsigned long long dstep = (drop && p.drop_step) ? *p.drop_step : 0ull;
  const uint32_t thr = (uint32_t)fminf(p.drop_p * 4294967296.f, 4294967040.f);
  const float dscale = drop ? 1.f / (1.f - p.drop_p) : 1.f;
  const uint2 dkey = make_uint2((uint32_t)p.drop_seed, (uint32_t)(p.drop_seed >> 32));
  constexpr int OUT4 = BM * BN / 4;
#pragma unroll
  for (int idx0 = 0; idx0 < OUT4; idx0 += NT) {
    const int idx = idx0 + tid;
    const int r = idx / (BN / 4), c4 = (idx % (BN / 4)) * 4;
    const int gm = m0 + r, gn = n0 + c4;
    if (gm >= M || gn >= N) continue;
    float4 v = *reinterpret_cast<const float4*>(red + r * LDR + c4);
#pragma unroll
    for (int w = 1; w < WK; w++) {
      const float4 u = *reinterpret_cast<const float4*>(red + (w * BM + r) * LDR + c4);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    float vv[4] = {v.x, v.y, v.z, v.w};
    float* cp = C + (long long)gm * p.ldc + gn;
    const int nval = min(4, N - gn);
    if (split) {      // cross-CTA split-K: accumulate (C already holds beta*C; bias from split 0)
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (q < nval) atomicAdd(cp + q, vv[q] + ((bias && ks == 0) ? bias[gn + q] : 0.f));
      continue;
    }
    if (bias) {
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (q < nval) vv[q] += bias[gn + q];
    }
    if (p.beta != 0.f) {
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (q < nval) vv[q] += p.beta * cp[q];
    }
    if (p.act) {
#pragma unroll
      for (int q = 0; q < 4; q++) vv[q] = act_apply1(vv[q], p.act);
    }
    if (drop) {
      const long long flat = (long long)gm * p.drop_ld + batch * p.drop_batch_stride + p.drop_col0 + gn;
      if ((flat & 3) == 0) {
        const unsigned long long c = (unsigned long long)(flat >> 2) + p.drop_offset;
        const uint4 rr = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)dstep, (uint32_t)(dstep >> 32)), dkey);
        vv[0] = rr.x >= thr ? vv[0] * dscale : 0.f;
        vv[1] = rr.y >= thr ? vv[1] * dscale : 0.f;
        vv[2] = rr.z >= thr ? vv[2] * dscale : 0.f;
        vv[3] = rr.w >= thr ? vv[3] * dscale : 0.f;
      } else {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const long long f = flat + q;
          const unsigned long long c = (unsigned long long)(f >> 2) + p.drop_offset;
          const uint4 rr = philox4x32_10(make_uint4((uint32_t)c, (uint32_t)(c >> 32), (uint32_t)dstep, (uint32_t)(dstep >> 32)), dkey);
          const uint32_t w4[4] = {rr.x, rr.y, rr.z, rr.w};
          vv[q] = w4[f & 3] >= thr ? vv[q] * dscale : 0.f;
        }
      }
    }
    if (vec_c && nval == 4) {
      *reinterpret_cast<float4*>(cp) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    } else {
#pragma unroll
      for (int q = 0; q < 4; q++)
        if (q < nval) cp[q] = vv[q];
    }
  }
  if constexpr (!AK) {
    if (want_col) {
      __syncthreads();
      if (tid < BM && m0 + tid < M) atomicAdd(p.colsum + batch * p.sColsum + m0 + tid, scol[tid]);
    }
  }
}


template <int BM, int BN, int WM, int WN, int WK, bool AK, bool BKC, bool GATE, bool VEC>
__global__ void __launch_bounds__(32 * WM * WN * WK) gemm_x3_kernel(const Args p) {
  DEER_PDL_ENTRY();
  extern __shared__ __align__(16) float smem_x3[];
  x3_tile<BM, BN, WM, WN, WK, AK, BKC, GATE, VEC>(p, blockIdx.x, blockIdx.y, blockIdx.z, smem_x3);
}

}  // namespace x3
}  // namespace deer
