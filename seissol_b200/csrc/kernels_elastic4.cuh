// Kernel generation 4 of the elastic element update: HALF-COLUMN threads.
//
// Same algorithm, shared-memory plan and TMA staging as kernels_elastic.cuh (k_local / k_neigh), but a
// quantity column of a group of V elements is owned by TWO threads: thread half H keeps the rows k with
// (k >> 1) & 1 == H (seissol_b200/codegen/gen_ops4.py).  28 instead of 56 accumulators per thread
// (order 6) buy
//   * double precision without register spills at 128 registers (16 resident warps per SM instead of 8
//     with 255 registers and local-memory traffic),
//   * single precision on packed FFMA2 (V = 2 elements per thread) inside the same budget.
// The two halves of a CTA run different instruction streams (different immediates); a half occupies whole
// warps, so there is no divergence.  Cross-thread data (dQ(d), I, face projections) is exchanged through the
// same shared-memory regions as before; the 9x9 flux-matrix product J = F.A is evaluated by both halves in
// 4-row chunks that feed the generated rDivM code directly (JChunks).
//
// Replaces: Spacetime::computeBatchedAder (src/Kernels/LinearCK/Time.cpp:142-217),
// Local::computeBatchedIntegral (src/Kernels/LinearCK/Local.cpp:255-395),
// Neighbor::computeBatchedNeighborsIntegral (src/Kernels/LinearCK/Neighbor.cpp:111-188).
#pragma once
#include "kernels_elastic.cuh"

namespace ssb {

__host__ __device__ constexpr int halfThreads(int NP) { return (kNQ * NP + 31) / 32 * 32; }
#ifndef SSB_MINB
#define SSB_MINB 2  // resident CTAs per SM the register allocation is sized for
#endif

// lane 1 of two packed values into an interleaved 128-bit slot ([k][V], V = 2): only the odd words change
__device__ __forceinline__ void stLane1(float* p, float2 a, float2 b) { p[1] = a.y; p[3] = b.y; }
__device__ __forceinline__ void stLane1(float*, float, float) {}
__device__ __forceinline__ void stLane1(double*, double, double) {}
__device__ __forceinline__ void stLane1(double* p, double2 a, double2 b) { p[1] = a.y; p[3] = b.y; }

template <typename R, typename G, int NP>
struct AderCtx4 {
  using P = LocalPlan<R, G, NP>;
  using PT = typename Pk<R>::T;
  static constexpr int V = Pk<R>::V;
  static constexpr int BP = P::BP;
  static constexpr int O = G::kOrder;
  const LocalParams& p;
  PT (&s)[3][3];
  PT (&icol)[P::LD1 / 2];  // own rows < LD1 of the I column, accumulated in registers
  R* RQ; R* RD;
  int r[3]; int q; int h; bool active; bool wantDer; const long long (&el)[V];
  __device__ __forceinline__ R* buf(int d) const { return d == 0 ? RQ : ((d & 1) ? RD : RD + P::offD2); }
  __device__ __forceinline__ static constexpr int ld(int d) { return d == 0 ? BP : pad4i(G::derRows(d)); }
  __device__ __forceinline__ const R* src(int d, int c) const { return buf(d - 1) + r[c] * ld(d - 1) * V; }
  template <int N>
  __device__ __forceinline__ void finish(int d, PT (&out)[N / 2]) {
    const R cd = static_cast<R>(p.coeffs[d]);
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
      icol[i] = pk_fmac(cd, out[i], icol[i]);
    }
    if (wantDer && active) {
#pragma unroll
      for (int v = 0; v < V; ++v) {
        R* const dbase = el[v] >= 0 ? static_cast<R*>(const_cast<void*>(p.derivs.at(el[v]))) : nullptr;
        if (dbase != nullptr) {  // NULL table entry: no derivatives kept for this element
          R* dst = dbase + p.derOffset[d] + q * p.derLd[d];
#pragma unroll
          for (int pc = 0; pc < N / 4; ++pc) {
            st2lane(dst + 4 * pc + 2 * h, v, out[2 * pc], out[2 * pc + 1]);
          }
        }
      }
    }
    if (d < O - 1) {
      R* dcol = buf(d) + q * N * V;
      if (active) {
#pragma unroll
        for (int pc = 0; pc < N / 4; ++pc) {
          stv2(dcol + (4 * pc + 2 * h) * V, out[2 * pc], out[2 * pc + 1]);
        }
      }
      __syncthreads();
    }
  }
};

// ---------------------------------------------------------------------------------------
template <typename R, typename G, int NP, bool ADER, bool LOCAL, int H>
__device__ __forceinline__ void localBody4(const LocalParams& p, unsigned char* smemRaw) {
  using P = LocalPlan<R, G, NP>;
  using PT = typename Pk<R>::T;
  constexpr int V = P::V, E = P::E, BP = P::BP, B2P = P::B2P, TEN = P::TEN;
  constexpr int HT = halfThreads(NP);
  constexpr int NPC = BP / 4;        // 128-bit pieces of a half column
  constexpr int NPC1 = P::LD1 / 4;   // pieces that receive derivative contributions
  R* const smem = reinterpret_cast<R*>(smemRaw);
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + P::barOffset);

  const int tid = threadIdx.x;
  const int t = tid - H * HT;
  const bool active = t < kNQ * NP;
  const int g = active ? t % NP : 0;
  const int q = active ? t / NP : 0;
  R* const RQ = smem + P::offQ + g * P::SEQ;
  R* const RD = smem + P::offD + g * P::SED;

  const long long nBatches = (p.n + E - 1) / E;
  uint32_t phase = 0;
  for (long long batch = blockIdx.x; batch < nBatches; batch += gridDim.x) {
    const long long slot = batch * E;

    // ---- TMA loads by the first E threads of the CTA (all in half 0)
    long long myElem = -1;
    R* myRQ = nullptr;
    R* myRD = nullptr;
    if (H == 0 && tid < E) {
      const long long sl = slot + tid;
      if (sl < p.n) {
        myElem = p.perm != nullptr ? p.perm[sl] : sl;
      }
      myRQ = smem + P::offQ + (tid / V) * P::SEQ + (tid % V) * TEN;
      myRD = smem + P::offD + (tid / V) * P::SED + (tid % V) * TEN;
      if (myElem >= 0) {
        mbar_arrive_expect_tx(bar, TEN * sizeof(R) * (ADER ? 1 : 2));
        tmaLoadTensor<R, BP>(myRQ, static_cast<const R*>(p.dofs.at(myElem)), p.ldg, bar);
        if (!ADER) {
          tmaLoadTensor<R, BP>(myRD, static_cast<const R*>(p.idofs.at(myElem)), p.ldg, bar);
        }
      } else {
        mbar_arrive_expect_tx(bar, 0);
      }
    }

    long long el[V];
    const R* LIg[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const long long sl = slot + g * V + v;
      el[v] = (active && sl < p.n) ? (p.perm != nullptr ? p.perm[sl] : sl) : -1;
      LIg[v] = el[v] >= 0 ? static_cast<const R*>(p.localInt.at(el[v])) : nullptr;
    }
    const int r0 = kStarRow[q][0], r1 = kStarRow[q][1], r2 = kStarRow[q][2];
    PT s[3][3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        R tv[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          tv[v] = (LIg[v] != nullptr && c < kStarCnt[q]) ? LIg[v][24 * j + kStarOff[q] + c] : R(0);
        }
        s[j][c] = pk_make(tv);
      }
    }
    unsigned skipMask = 0;  // bit 4 v + f: element v gets no local flux through face f
#pragma unroll
    for (int v = 0; v < V; ++v) {
      uchar4 ft = make_uchar4(0, 0, 0, 0);
      if (LOCAL && p.faceTypes != nullptr && el[v] >= 0) {
        ft = *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * el[v]);
      }
      const bool none = el[v] < 0;  // FaceType::DynamicRupture (3) gets no local flux
      skipMask |= ((none || ft.x == 3) ? 1u : 0u) << (4 * v);
      skipMask |= ((none || ft.y == 3) ? 2u : 0u) << (4 * v);
      skipMask |= ((none || ft.z == 3) ? 4u : 0u) << (4 * v);
      skipMask |= ((none || ft.w == 3) ? 8u : 0u) << (4 * v);
    }

    mbar_wait(bar, phase);
    phase ^= 1;

    PT acc[BP / 2];  // own rows of the Q column (LOCAL only)
    const SrcSplit<R> stagedQ{RQ + q * BP, RQ + TEN + q * BP};
    const SrcSplit<R> stagedD{RD + q * BP, RD + TEN + q * BP};

    if constexpr (ADER) {
      const bool wantDer = p.derivs.present();
      if (H == 0 && wantDer && tid < E && myElem >= 0) {
        // dQ(0) = Q copy into the derivative family (Time.cpp:109-116), from the staged tensor
        R* const dbase = static_cast<R*>(const_cast<void*>(p.derivs.at(myElem)));
        if (dbase != nullptr) {
          R* dst = dbase + p.derOffset[0];
          for (int c = 0; c < kNQ; ++c) {
            tma_store(dst + c * p.derLd[0], myRQ + c * BP, BP * sizeof(R));
          }
          tma_store_commit();
          tma_store_wait_read();
        }
      }
      PT icol[P::LD1 / 2];
      {
        PT qcol[BP / 2];
#pragma unroll
        for (int pc = 0; pc < NPC; ++pc) {
          PT tt[2];
          stagedQ.load2(tt, 4 * pc + 2 * H);
          qcol[2 * pc] = tt[0]; qcol[2 * pc + 1] = tt[1];
        }
        const R c0 = static_cast<R>(p.coeffs[0]);
#pragma unroll
        for (int i = 0; i < P::LD1 / 2; ++i) {
          icol[i] = pk_mulc(c0, qcol[i]);
        }
        if constexpr (V > 1) {
          __syncthreads();  // everybody has read its staged rows: interleave in place
          if (active) {
#pragma unroll
            for (int pc = 0; pc < NPC; ++pc) {
              stv2(RQ + (q * BP + 4 * pc + 2 * H) * V, qcol[2 * pc], qcol[2 * pc + 1]);
            }
          }
          __syncthreads();
        }
      }

      AderCtx4<R, G, NP> ctx{p, s, icol, RQ, RD, {r0, r1, r2}, q, H, active, wantDer, el};
      G::template ader_run<R, H>(ctx);

      __syncthreads();  // nobody reads dQ(>=1) any more: RD becomes the interleaved I
      PT iown[BP / 2];  // own rows of I
      {
        const R c0 = static_cast<R>(p.coeffs[0]);
#pragma unroll
        for (int i = 0; i < P::LD1 / 2; ++i) {
          iown[i] = icol[i];
        }
#pragma unroll
        for (int pc = 0; pc < NPC; ++pc) {
          PT tt[2];
          ldv2(tt, RQ + (q * BP + 4 * pc + 2 * H) * V);
          if constexpr (LOCAL) {
            acc[2 * pc] = tt[0]; acc[2 * pc + 1] = tt[1];
          }
          if (pc >= NPC1) {  // rows >= LD1 never receive a derivative contribution: I = c0 * Q there
            iown[2 * pc] = pk_mulc(c0, tt[0]); iown[2 * pc + 1] = pk_mulc(c0, tt[1]);
          }
        }
        if (active) {
#pragma unroll
          for (int pc = 0; pc < NPC; ++pc) {
            stv2(RD + (q * BP + 4 * pc + 2 * H) * V, iown[2 * pc], iown[2 * pc + 1]);
          }
        }
      }
      __syncthreads();  // I visible in RD; nobody reads Q from RQ any more
      if (active) {     // stage I per element in RQ for the bulk store
#pragma unroll
        for (int pc = 0; pc < NPC; ++pc) {
#pragma unroll
          for (int v = 0; v < V; ++v) {
            st2lane(RQ + v * TEN + q * BP + 4 * pc + 2 * H, v, iown[2 * pc], iown[2 * pc + 1]);
          }
        }
      }
      fence_async_smem();
      __syncthreads();
      if (H == 0 && tid < E && myElem >= 0) {
        tmaStoreTensor<R, BP>(static_cast<R*>(const_cast<void*>(p.idofs.at(myElem))), myRQ, p.ldg);
      }
    } else {
      // I staged in RD -> interleaved in place; Q staged in RQ -> registers
      if constexpr (V > 1) {
        PT irows[BP / 2];
#pragma unroll
        for (int pc = 0; pc < NPC; ++pc) {
          PT tt[2];
          stagedD.load2(tt, 4 * pc + 2 * H);
          irows[2 * pc] = tt[0]; irows[2 * pc + 1] = tt[1];
        }
        __syncthreads();
        if (active) {
#pragma unroll
          for (int pc = 0; pc < NPC; ++pc) {
            stv2(RD + (q * BP + 4 * pc + 2 * H) * V, irows[2 * pc], irows[2 * pc + 1]);
          }
        }
      }
#pragma unroll
      for (int pc = 0; pc < NPC; ++pc) {
        PT tt[2];
        stagedQ.load2(tt, 4 * pc + 2 * H);
        acc[2 * pc] = tt[0]; acc[2 * pc + 1] = tt[1];
      }
      __syncthreads();  // interleaved I visible; RQ free for the face buffers
    }

    if constexpr (LOCAL) {
      const R* const Ie = RD;  // interleaved I
      const bool wantProj = p.faceProj.present();
      G::template volume<R, H>(acc, Ie + r0 * BP * V, Ie + r1 * BP * V, Ie + r2 * BP * V, s);

      if (ADER) {
        if (H == 0 && tid < E) {
          tma_store_wait_read();  // the bulk store of I has left RQ
        }
        __syncthreads();
      }
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        PT a[kNQ];
#pragma unroll
        for (int c = 0; c < kNQ; ++c) {
          R tv[V];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            tv[v] = ((skipMask >> (4 * v + f)) & 1u) ? R(0) : LIg[v][72 + 81 * f + 9 * q + c];
          }
          a[c] = pk_make(tv);
        }
        R* const Fe = RQ + (f & 1) * kNQ * B2P * V;
        {
          PT F[B2P / 2];
          G::template fMrT<R, H>(f, F, SrcPk<R>{Ie + q * BP * V});
          if (active) {
#pragma unroll
            for (int pc = 0; pc < B2P / 4; ++pc) {
              stv2(Fe + (q * B2P + 4 * pc + 2 * H) * V, F[2 * pc], F[2 * pc + 1]);
            }
          }
        if constexpr (V > 1) {
          // face projections for the neighbor kernel (kernels_faceproj.cuh): the packed rows leave the registers
          // element by element (the shared-memory copy is lane-interleaved, no bulk store possible)
          if (wantProj && active) {
#pragma unroll
            for (int v = 0; v < V; ++v) {
              if (el[v] >= 0) {
                R* const dst = static_cast<R*>(const_cast<void*>(p.faceProj.at(el[v]))) + (f * kNQ + q) * B2P;
#pragma unroll
                for (int pc = 0; pc < B2P / 4; ++pc) {
                  st2lane(dst + 4 * pc + 2 * H, v, F[2 * pc], F[2 * pc + 1]);
                }
              }
            }
          }
        }
        }
        if constexpr (V == 1) {
          // face projections for the neighbor kernel (kernels_faceproj.cuh), see k_local
          if (wantProj) {
            fence_async_smem();
            if (f >= 1 && H == 0 && tid < E) {
              tma_store_wait_read();
            }
          }
        }
        __syncthreads();
        if constexpr (V == 1) {
          if (wantProj && H == 0 && tid < E && myElem >= 0) {
            tma_store(static_cast<R*>(const_cast<void*>(p.faceProj.at(myElem))) + f * kNQ * B2P,
                      myRQ + (f & 1) * kNQ * B2P, kNQ * B2P * sizeof(R));
            tma_store_commit();
          }
        }
        G::template rDivM<R, H>(f, acc, JChunks<R, B2P>{Fe, a});
      }

      // Q_new staged per element in RD (the last reader of I passed the barrier of face 3)
      if (active) {
#pragma unroll
        for (int pc = 0; pc < NPC; ++pc) {
#pragma unroll
          for (int v = 0; v < V; ++v) {
            st2lane(RD + v * TEN + q * BP + 4 * pc + 2 * H, v, acc[2 * pc], acc[2 * pc + 1]);
          }
        }
      }
      fence_async_smem();
      __syncthreads();
      if (H == 0 && tid < E && myElem >= 0) {
        tmaStoreTensor<R, BP>(static_cast<R*>(const_cast<void*>(p.dofs.at(myElem))), myRD, p.ldg);
      }
    }
    if (H == 0 && tid < E) {
      tma_store_wait_read();
    }
    __syncthreads();
  }  // batch loop
}

template <typename R, typename G, int NP, bool ADER, bool LOCAL>
__global__ void __launch_bounds__(2 * halfThreads(NP), SSB_MINB) k_local4(const LocalParams p) {
  using P = LocalPlan<R, G, NP>;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  if (threadIdx.x == 0) {
    uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + P::barOffset);
    mbar_init(bar, P::E);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x < halfThreads(NP)) {
    localBody4<R, G, NP, ADER, LOCAL, 0>(p, smemRaw);
  } else {
    localBody4<R, G, NP, ADER, LOCAL, 1>(p, smemRaw);
  }
}

// ---------------------------------------------------------------------------------------
// k_neigh4: per group V staged neighbor tiles (TMA target, re-used for the elements' own Q at the end),
// one buffer for G = rT_j . I_nb and one for F = fP_h . G (both interleaved [c][k][V]).
template <typename R, typename G, int NP, int H>
__device__ __forceinline__ void neighBody4(const NeighParams& p, unsigned char* smemRaw) {
  using P = NeighPlan<R, G, NP>;
  using PT = typename Pk<R>::T;
  constexpr int V = P::V, E = P::E, BP = P::BP, B2P = P::B2P, TEN = P::TEN;
  constexpr int HT = halfThreads(NP);
  R* const smem = reinterpret_cast<R*>(smemRaw);
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + P::barOffset);

  const int tid = threadIdx.x;
  const int t = tid - H * HT;
  const bool active = t < kNQ * NP;
  const int g = active ? t % NP : 0;
  const int q = active ? t / NP : 0;

  const long long nBatches = (p.n + E - 1) / E;
  uint32_t phase = 0;
  for (long long batch = blockIdx.x; batch < nBatches; batch += gridDim.x) {
    const long long slot = batch * E;

    long long el[V];
    const R* NIg[V];
    unsigned ftReg[V];   // face f regular? bit f
    unsigned relJ[V];    // 2 bits per face: neighbor's face id j
    unsigned relH[V];    // 2 bits per face: orientation h
#pragma unroll
    for (int v = 0; v < V; ++v) {
      const long long sl = slot + g * V + v;
      el[v] = (active && sl < p.n) ? (p.perm != nullptr ? p.perm[sl] : sl) : -1;
      NIg[v] = el[v] >= 0 ? static_cast<const R*>(p.neighInt.at(el[v])) : nullptr;
      ftReg[v] = 0; relJ[v] = 0; relH[v] = 0;
      if (el[v] >= 0) {
        uchar4 ft = make_uchar4(0, 0, 0, 0);
        if (p.faceTypes != nullptr) {
          ft = *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * el[v]);
        }
        ftReg[v] = (ft.x == 0 ? 1u : 0u) | (ft.y == 0 ? 2u : 0u) | (ft.z == 0 ? 4u : 0u) | (ft.w == 0 ? 8u : 0u);
        if (p.nbPtrs != nullptr) {   // a NULL FaceNeighbors entry loads no tile: treat the face like a non-regular one
#pragma unroll
          for (int f = 0; f < 4; ++f) {
            if (p.nbPtrs[4 * el[v] + f] == nullptr) {
              ftReg[v] &= ~(1u << f);
            }
          }
        }
        const uint2 rr = *reinterpret_cast<const uint2*>(p.faceRel + 8 * el[v]);
        relJ[v] = (rr.x & 3u) | (((rr.x >> 16) & 3u) << 2) | ((rr.y & 3u) << 4) | (((rr.y >> 16) & 3u) << 6);
        relH[v] = ((rr.x >> 8) & 3u) | (((rr.x >> 24) & 3u) << 2) | (((rr.y >> 8) & 3u) << 4) | (((rr.y >> 24) & 3u) << 6);
      }
    }

    long long myElem = -1;
    R* myTile = nullptr;
    uchar4 myFt = make_uchar4(0, 0, 0, 0);
    if (H == 0 && tid < E) {
      const long long sl = slot + tid;
      if (sl < p.n) {
        myElem = p.perm != nullptr ? p.perm[sl] : sl;
      }
      myTile = smem + P::offN + (tid / V) * P::SEQ + (tid % V) * TEN;
      if (myElem >= 0 && p.faceTypes != nullptr) {
        myFt = *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * myElem);
      }
    }
    auto loadTile = [&](const R* src) {
      if (src == nullptr) {
        mbar_arrive_expect_tx(bar, 0);
        return;
      }
      mbar_arrive_expect_tx(bar, TEN * sizeof(R));
      tmaLoadTensor<R, BP>(myTile, src, p.ldg, bar);
    };
    auto neighborOf = [&](int f) -> const R* {
      if (myElem < 0) {
        return nullptr;
      }
      const uint8_t ty = f == 0 ? myFt.x : (f == 1 ? myFt.y : (f == 2 ? myFt.z : myFt.w));
      if (ty != 0) {
        return nullptr;
      }
      if (p.nbPtrs != nullptr) {
        return static_cast<const R*>(p.nbPtrs[4 * myElem + f]);
      }
      return static_cast<const R*>(p.idofs.at(p.nbIdx[4 * myElem + f]));
    };
    if (H == 0 && tid < E) {
      loadTile(neighborOf(0));
    }

    PT acc[BP / 2];
#pragma unroll
    for (int i = 0; i < BP / 2; ++i) {
      acc[i] = pk_zero<R>();
    }
    R* const tiles = smem + P::offN + g * P::SEQ;
    const SrcSplit<R> nbCol{tiles + q * BP, tiles + TEN + q * BP};
    R* const Gb = smem + P::offF + g * P::SEF;
    R* const Fb = Gb + kNQ * B2P * V;

#pragma unroll
    for (int f = 0; f < 4; ++f) {
      const int j0 = (relJ[0] >> (2 * f)) & 3, jL = (relJ[V - 1] >> (2 * f)) & 3;
      const int h0 = (relH[0] >> (2 * f)) & 3, hL = (relH[V - 1] >> (2 * f)) & 3;
      mbar_wait(bar, (phase + f) & 1);
      {
        // G = rT_j . I_nb for the own rows.  The lanes of a group share the instruction stream: when their j
        // differ (rare after sorting) the operator runs a second time and only lane 1 is stored.
        PT Gt[B2P / 2];
        G::template rT<R, H>(j0, Gt, nbCol);
        if (active) {
#pragma unroll
          for (int pc = 0; pc < B2P / 4; ++pc) {
            stv2(Gb + (q * B2P + 4 * pc + 2 * H) * V, Gt[2 * pc], Gt[2 * pc + 1]);
          }
        }
        if constexpr (V > 1) {
          if (jL != j0) {
            G::template rT<R, H>(jL, Gt, nbCol);
            if (active) {
#pragma unroll
              for (int pc = 0; pc < B2P / 4; ++pc) {
                stLane1(Gb + (q * B2P + 4 * pc + 2 * H) * V, Gt[2 * pc], Gt[2 * pc + 1]);
              }
            }
          }
        }
      }
      __syncthreads();  // G visible; the tiles are no longer read
      if (H == 0 && tid < E) {
        // next neighbor, or finally the element's own Q
        loadTile(f < 3 ? neighborOf(f + 1) : (myElem >= 0 ? static_cast<const R*>(p.dofs.at(myElem)) : nullptr));
      }
      // column q of AminusT_f straight from global memory; latency hidden by fP and the barrier
      PT a[kNQ];
#pragma unroll
      for (int c = 0; c < kNQ; ++c) {
        R tv[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          tv[v] = ((ftReg[v] >> f) & 1u) ? NIg[v][81 * f + 9 * q + c] : R(0);
        }
        a[c] = pk_make(tv);
      }
      {
        // faces without a regular neighbor: no tile was loaded, G came from stale shared memory -> F exactly zero
        bool keep[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
          keep[v] = ((ftReg[v] >> f) & 1u) != 0;
        }
        PT F[B2P / 2];
        const SrcPk<R> gcol{Gb + q * B2P * V};
        G::template fP<R, H>(h0, F, gcol);
        if (active) {
#pragma unroll
          for (int pc = 0; pc < B2P / 4; ++pc) {
            stv2(Fb + (q * B2P + 4 * pc + 2 * H) * V, pk_keep(F[2 * pc], keep), pk_keep(F[2 * pc + 1], keep));
          }
        }
        if constexpr (V > 1) {
          if (hL != h0) {
            G::template fP<R, H>(hL, F, gcol);
            if (active) {
#pragma unroll
              for (int pc = 0; pc < B2P / 4; ++pc) {
                stLane1(Fb + (q * B2P + 4 * pc + 2 * H) * V, pk_keep(F[2 * pc], keep), pk_keep(F[2 * pc + 1], keep));
              }
            }
          }
        }
      }
      __syncthreads();  // F visible
      G::template rDivM<R, H>(f, acc, JChunks<R, B2P>{Fb, a});
    }

    // Q += acc: the own Q tensors arrived in the tiles
    mbar_wait(bar, phase);
    phase ^= 1;
    if (active) {
#pragma unroll
      for (int pc = 0; pc < BP / 4; ++pc) {
        PT tt[2];
        nbCol.load2(tt, 4 * pc + 2 * H);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          st2lane(tiles + v * TEN + q * BP + 4 * pc + 2 * H, v, pk_add(tt[0], acc[2 * pc]), pk_add(tt[1], acc[2 * pc + 1]));
        }
      }
    }
    fence_async_smem();
    __syncthreads();
    if (H == 0 && tid < E && myElem >= 0) {
      tmaStoreTensor<R, BP>(static_cast<R*>(const_cast<void*>(p.dofs.at(myElem))), myTile, p.ldg);
      tma_store_wait_read();
    }
    __syncthreads();
  }  // batch loop
}

template <typename R, typename G, int NP>
__global__ void __launch_bounds__(2 * halfThreads(NP), SSB_MINB) k_neigh4(const NeighParams p) {
  using P = NeighPlan<R, G, NP>;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  if (threadIdx.x == 0) {
    uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + P::barOffset);
    mbar_init(bar, P::E);
    mbar_fence_init();
  }
  __syncthreads();
  if (threadIdx.x < halfThreads(NP)) {
    neighBody4<R, G, NP, 0>(p, smemRaw);
  } else {
    neighBody4<R, G, NP, 1>(p, smemRaw);
  }
}

}  // namespace ssb
