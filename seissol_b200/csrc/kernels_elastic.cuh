// Hand-written sm_100a kernels for the elastic ADER-DG element update.
//
//   k_local : Cauchy-Kovalevskaya predictor (ader) + volume + local flux, fused, the
//             time-integrated DOFs `I` never leave the SM between the stages.
//             Replaces Spacetime::computeBatchedAder (src/Kernels/LinearCK/Time.cpp:142-217)
//             and Local::computeBatchedIntegral (src/Kernels/LinearCK/Local.cpp:255-395).
//   k_neigh : neighbor flux for all 4 faces and all 48 (h,j,i) relations in ONE launch.
//             Replaces Neighbor::computeBatchedNeighborsIntegral
//             (src/Kernels/LinearCK/Neighbor.cpp:111-188: up to 4x48 launches).
//   k_taylor: I = sum_d c_d dQ(d)  (Time::evaluateBatched, Time.cpp:257-292).
//
// Mapping: a CTA owns E elements; thread (e, q) owns quantity column q of
// element e in registers.  Element tensors live in shared memory as [e][q][k] with a
// per-element stride padded so that 128-bit accesses of consecutive lanes (= consecutive
// elements) are bank-conflict free; they arrive/leave by 1-D TMA bulk copies
// (cp.async.bulk + mbarrier) straight from/to the per-element pointers of the batched
// interface.  All constant operators are immediates in the instruction stream
// (seissol_b200/codegen/gen_ops.py).
//
// The instruction stream of one thread is ~10^4 straight-line FFMAs (no reuse), so the
// kernels are bound by instruction fetch unless >= 4 warps per SM sub-partition are
// resident (tools/ubench/gen_ifetch.py): the shared-memory plan below is
// squeezed to two regions per element (4 KB for order 6 / float) so that 16 warps fit.
#pragma once
#include "common.cuh"
#include "params.cuh"

namespace ssb {

constexpr int kNQ = 9;
__host__ __device__ constexpr int blockThreads(int E) { return (kNQ * E + 31) / 32 * 32; }
__host__ __device__ constexpr int imax(int a, int b) { return a > b ? a : b; }

// Shared-memory plan of k_local, usable on host and device.  A thread owns quantity column q of
// a GROUP of V elements (V = 2 packed floats, 1 double); NP groups per CTA.  Two regions per
// group, both either "staged" (V per-element tensors [v][q][k], what TMA moves) or
// "interleaved" ([q][k][v], what the packed arithmetic reads):
//   RQ : Q staged (TMA in) -> Q interleaved (= dQ(0)) -> I staged (TMA out) -> face buffers F0|F1
//   RD : dQ(odd) at +0, dQ(even>=2) at +offD2 -> I interleaved -> Q_new staged (TMA out)
// Without the predictor: RD = I staged -> interleaved, RQ = Q staged -> F0|F1.
template <typename R, typename G, int NP>
struct LocalPlan {
  static constexpr int V = Pk<R>::V;
  static constexpr int E = NP * V;
  static constexpr int BP = G::kBP;
  static constexpr int B2P = G::kB2P;
  static constexpr int LD1 = pad4i(G::derRows(1));
  static constexpr int LD2 = pad4i(G::derRows(2));
  static constexpr int offD2 = kNQ * LD1 * V;
  static constexpr int TEN = kNQ * BP;  // one element tensor (reals)
  static constexpr int SEQ = padOdd16(V * imax(TEN, 2 * kNQ * B2P), sizeof(R));
  static constexpr int SED = padOdd16(V * imax(kNQ * (LD1 + LD2), TEN), sizeof(R));
  static constexpr int offQ = 0;
  static constexpr int offD = NP * SEQ;
  static constexpr int endReals = NP * (SEQ + SED);
  static constexpr size_t barOffset = (size_t(endReals) * sizeof(R) + 15) / 16 * 16;
  static constexpr size_t bytes = barOffset + 16;
};

// ---------------------------------------------------------------------------------------
// Per-thread context handed to the generated ADER driver (gen::ader_run): it tells the
// generated code where dQ(d-1) lives in shared memory and what to do with a finished dQ(d).
template <typename R, typename G, int NP>
struct AderCtx {
  using P = LocalPlan<R, G, NP>;
  using PT = typename Pk<R>::T;
  static constexpr int V = Pk<R>::V;
  static constexpr int BP = P::BP;
  static constexpr int O = G::kOrder;
  const LocalParams& p;
  PT (&s)[3][3];
  PT (&icol)[P::LD1];  // rows < LD1 of the own I column, accumulated in registers
  R* RQ; R* RD;
  int r[3]; int q; bool active; bool wantDer; const long long (&el)[V];
  __device__ __forceinline__ R* buf(int d) const { return d == 0 ? RQ : ((d & 1) ? RD : RD + P::offD2); }
  __device__ __forceinline__ static constexpr int ld(int d) { return d == 0 ? BP : pad4i(G::derRows(d)); }
  __device__ __forceinline__ const R* src(int d, int c) const { return buf(d - 1) + r[c] * ld(d - 1) * V; }
  template <int N>
  __device__ __forceinline__ void finish(int d, PT (&out)[N]) {
    const R cd = static_cast<R>(p.coeffs[d]);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      icol[k] = pk_fmac(cd, out[k], icol[k]);
    }
    if (wantDer && active) {
      // derivative family (LTS / dynamic rupture consumers): straight to global memory
#pragma unroll
      for (int v = 0; v < V; ++v) {
        // a NULL entry in the derivative table: this element keeps no derivatives (Time.cpp:150-160)
        R* const dbase = el[v] >= 0 ? static_cast<R*>(const_cast<void*>(p.derivs.at(el[v]))) : nullptr;
        if (dbase != nullptr) {
          R* dst = dbase + p.derOffset[d] + q * p.derLd[d];
#pragma unroll
          for (int k4 = 0; k4 < N / 4; ++k4) {
            st4lane(dst + 4 * k4, v, out[4 * k4], out[4 * k4 + 1], out[4 * k4 + 2], out[4 * k4 + 3]);
          }
        }
      }
    }
    if (d < O - 1) {
      // dQ(d) overwrites dQ(d-2), which everybody finished reading before the barrier of step d-1
      R* dcol = buf(d) + q * N * V;
      if (active) {
#pragma unroll
        for (int k4 = 0; k4 < N / 4; ++k4) {
          stv4(dcol + 4 * k4 * V, out[4 * k4], out[4 * k4 + 1], out[4 * k4 + 2], out[4 * k4 + 3]);
        }
      }
      __syncthreads();
    }
  }
};

// J = F . A in chunks of 4 rows for the generated rDivM code: chunk c = rows 4c..4c+3 of column q of
// (F . AplusT/AminusT), F interleaved in shared memory ([c][k][V]), a = column q of the 9x9 flux matrix.
template <typename R, int B2P>
struct JChunks {
  using PT = typename Pk<R>::T;
  static constexpr int V = Pk<R>::V;
  const R* Fe;
  const PT (&a)[kNQ];
  template <int C>
  __device__ __forceinline__ void chunk(PT (&x)[4]) const {
#pragma unroll
    for (int c = 0; c < kNQ; ++c) {
      PT t[4];
      ldv4(t, Fe + (c * B2P + 4 * C) * V);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        x[i] = c == 0 ? pk_mul(a[c], t[i]) : pk_fma(a[c], t[i], x[i]);
      }
    }
  }
};

// own column q of the V staged element tensors of a group -> packed registers
template <typename R, int BP, int TEN>
__device__ __forceinline__ void loadStagedColumn(typename Pk<R>::T (&col)[BP], const R* staged, int q) {
  const SrcSplit<R> src{staged + q * BP, staged + TEN + q * BP};
#pragma unroll
  for (int k4 = 0; k4 < BP / 4; ++k4) {
    typename Pk<R>::T t[4];
    src.load4(t, 4 * k4);
    col[4 * k4] = t[0]; col[4 * k4 + 1] = t[1]; col[4 * k4 + 2] = t[2]; col[4 * k4 + 3] = t[3];
  }
}

// bulk copy of one element tensor ([9][BP] reals, global leading dimension ldg) global <-> shared
template <typename R, int BP>
__device__ __forceinline__ void tmaLoadTensor(R* dst, const R* src, int ldg, uint64_t* bar) {
  if (ldg == BP) {
    tma_load(dst, src, kNQ * BP * sizeof(R), bar);
  } else {
    for (int c = 0; c < kNQ; ++c) {
      tma_load(dst + c * BP, src + c * ldg, BP * sizeof(R), bar);
    }
  }
}
template <typename R, int BP>
__device__ __forceinline__ void tmaStoreTensor(R* dst, const R* src, int ldg) {
  if (ldg == BP) {
    tma_store(dst, src, kNQ * BP * sizeof(R));
  } else {
    for (int c = 0; c < kNQ; ++c) {
      tma_store(dst + c * ldg, src + c * BP, BP * sizeof(R));
    }
  }
  tma_store_commit();
}

// ---------------------------------------------------------------------------------------
#ifndef SSB_MINB_LOCAL
#define SSB_MINB_LOCAL 2  // resident CTAs per SM the register allocation of k_local is sized for
#endif
template <typename R, typename G, int NP, bool ADER, bool LOCAL>
__global__ void __launch_bounds__(blockThreads(NP), SSB_MINB_LOCAL) k_local(const LocalParams p) {
  using P = LocalPlan<R, G, NP>;
  using PT = typename Pk<R>::T;
  constexpr int V = P::V, E = P::E, BP = P::BP, B2P = P::B2P, TEN = P::TEN;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  R* const smem = reinterpret_cast<R*>(smemRaw);
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + P::barOffset);

  const int tid = threadIdx.x;
  const bool active = tid < kNQ * NP;
  const int g = active ? tid % NP : 0;  // group (pair) index inside the CTA
  const int q = active ? tid / NP : 0;
  R* const RQ = smem + P::offQ + g * P::SEQ;
  R* const RD = smem + P::offD + g * P::SED;

  if (tid == 0) {
    mbar_init(bar, E);
    mbar_fence_init();
  }
  __syncthreads();

  // Batches of E elements; with gridDim.x < #batches the CTAs are persistent and walk the batches
  // in lock step, which keeps all SMs inside the same code region (instruction-cache sharing).
  const long long nBatches = (p.n + E - 1) / E;
  uint32_t phase = 0;
  for (long long batch = blockIdx.x; batch < nBatches; batch += gridDim.x) {
  const long long slot = batch * E;

  // ---- TMA loads: thread t < E moves element slot+t (lane t%V of group t/V).
  //      Predictor: Q -> RQ.  Without predictor: Q -> RQ, I -> RD.
  long long myElem = -1;
  R* myRQ = nullptr;
  R* myRD = nullptr;
  if (tid < E) {
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

  // ---- the V elements of this thread's group; their star coefficients for column q straight
  //      from global memory while the bulk copies are in flight
  long long el[V];
  const R* LIg[V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const long long sl = slot + g * V + v;
    el[v] = sl < p.n ? (p.perm != nullptr ? p.perm[sl] : sl) : -1;
    LIg[v] = el[v] >= 0 ? static_cast<const R*>(p.localInt.at(el[v])) : nullptr;
  }
  const int r0 = kStarRow[q][0], r1 = kStarRow[q][1], r2 = kStarRow[q][2];
  PT s[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      R t[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        t[v] = (LIg[v] != nullptr && c < kStarCnt[q]) ? LIg[v][24 * j + kStarOff[q] + c] : R(0);
      }
      s[j][c] = pk_make(t);
    }
  }
  bool skipFace[4][V];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    uchar4 t = make_uchar4(0, 0, 0, 0);
    if (LOCAL && p.faceTypes != nullptr && el[v] >= 0) {
      t = *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * el[v]);
    }
    const bool none = el[v] < 0;  // FaceType::DynamicRupture (3) gets no local flux
    skipFace[0][v] = none || t.x == 3; skipFace[1][v] = none || t.y == 3;
    skipFace[2][v] = none || t.z == 3; skipFace[3][v] = none || t.w == 3;
  }

  mbar_wait(bar, phase);
  phase ^= 1;

  PT acc[BP];  // own Q column (LOCAL only)

  if constexpr (ADER) {
    const bool wantDer = p.derivs.present();
    // dQ(0) = Q copy into the derivative family (Time.cpp:109-116), from the staged tensor
    if (wantDer && tid < E && myElem >= 0) {
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
    PT icol[P::LD1];
    {
      PT qcol[BP];
      loadStagedColumn<R, BP, TEN>(qcol, RQ, q);
      const R c0 = static_cast<R>(p.coeffs[0]);
#pragma unroll
      for (int k = 0; k < P::LD1; ++k) {
        icol[k] = pk_mulc(c0, qcol[k]);
      }
      if constexpr (V > 1) {
        __syncthreads();  // everybody has read its staged columns: interleave in place
        if (active) {
#pragma unroll
          for (int k4 = 0; k4 < BP / 4; ++k4) {
            stv4(RQ + (q * BP + 4 * k4) * V, qcol[4 * k4], qcol[4 * k4 + 1], qcol[4 * k4 + 2], qcol[4 * k4 + 3]);
          }
        }
        __syncthreads();
      }
    }

    AderCtx<R, G, NP> ctx{p, s, icol, RQ, RD, {r0, r1, r2}, q, active, wantDer, el};
    G::template ader_run<R>(ctx);

    __syncthreads();  // nobody reads dQ(>=1) any more: RD becomes the interleaved I
    {
      const R c0 = static_cast<R>(p.coeffs[0]);
      if (active) {
#pragma unroll
        for (int k4 = 0; k4 < P::LD1 / 4; ++k4) {
          stv4(RD + (q * BP + 4 * k4) * V, icol[4 * k4], icol[4 * k4 + 1], icol[4 * k4 + 2], icol[4 * k4 + 3]);
        }
      }
      // rows >= LD1 never receive a derivative contribution: I = c0 * Q there
#pragma unroll
      for (int k4 = P::LD1 / 4; k4 < BP / 4; ++k4) {
        PT t[4];
        ldv4(t, RQ + (q * BP + 4 * k4) * V);
        if constexpr (LOCAL) {
          acc[4 * k4] = t[0]; acc[4 * k4 + 1] = t[1]; acc[4 * k4 + 2] = t[2]; acc[4 * k4 + 3] = t[3];
        }
        if (active) {
          stv4(RD + (q * BP + 4 * k4) * V, pk_mulc(c0, t[0]), pk_mulc(c0, t[1]), pk_mulc(c0, t[2]), pk_mulc(c0, t[3]));
        }
      }
      if constexpr (LOCAL) {
#pragma unroll
        for (int k4 = 0; k4 < P::LD1 / 4; ++k4) {
          PT t[4];
          ldv4(t, RQ + (q * BP + 4 * k4) * V);
          acc[4 * k4] = t[0]; acc[4 * k4 + 1] = t[1]; acc[4 * k4 + 2] = t[2]; acc[4 * k4 + 3] = t[3];
        }
      }
    }
    __syncthreads();  // I visible in RD; nobody reads Q from RQ any more
    // stage I per element in RQ for the bulk store
    if (active) {
#pragma unroll
      for (int k4 = 0; k4 < BP / 4; ++k4) {
        PT t[4];
        ldv4(t, RD + (q * BP + 4 * k4) * V);
#pragma unroll
        for (int v = 0; v < V; ++v) {
          st4lane(RQ + v * TEN + q * BP + 4 * k4, v, t[0], t[1], t[2], t[3]);
        }
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid < E && myElem >= 0) {
      tmaStoreTensor<R, BP>(static_cast<R*>(const_cast<void*>(p.idofs.at(myElem))), myRQ, p.ldg);
    }
  } else {
    // I staged in RD -> interleaved in place; Q staged in RQ -> registers
    if constexpr (V > 1) {
      PT icolFull[BP];
      loadStagedColumn<R, BP, TEN>(icolFull, RD, q);
      __syncthreads();
      if (active) {
#pragma unroll
        for (int k4 = 0; k4 < BP / 4; ++k4) {
          stv4(RD + (q * BP + 4 * k4) * V, icolFull[4 * k4], icolFull[4 * k4 + 1], icolFull[4 * k4 + 2],
               icolFull[4 * k4 + 3]);
        }
      }
    }
    loadStagedColumn<R, BP, TEN>(acc, RQ, q);
    __syncthreads();  // interleaved I visible; RQ free for the face buffers
  }

  if constexpr (LOCAL) {
    const R* const Ie = RD;  // interleaved I
    const bool wantProj = p.faceProj.present();
    // ---- volume: Q += sum_i kDivM_i . (I . star_i)
    G::template volume<R>(acc, Ie + r0 * BP * V, Ie + r1 * BP * V, Ie + r2 * BP * V, s);

    if (ADER) {
      if (tid < E) {
        tma_store_wait_read();  // the bulk store of I has left RQ
      }
      __syncthreads();
    }
    // ---- local flux: Q += rDivM_f . ((fMrT_f . I) . AplusT_f), skipped on DR faces
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      // column q of AplusT_f straight from global memory; latency hidden by fMrT (single precision) --
      // in double precision the registers are needed for the accumulators, so it is loaded after fMrT
      PT a[kNQ];
      auto loadA = [&]() {
#pragma unroll
        for (int c = 0; c < kNQ; ++c) {
          R t[V];
#pragma unroll
          for (int v = 0; v < V; ++v) {
            t[v] = skipFace[f][v] ? R(0) : LIg[v][72 + 81 * f + 9 * q + c];
          }
          a[c] = pk_make(t);
        }
      };
      if constexpr (sizeof(R) == 4) {
        loadA();
      }
      R* const Fe = RQ + (f & 1) * kNQ * B2P * V;
      {
        PT F[B2P];
        G::template fMrT<R>(f, F, SrcPk<R>{Ie + q * BP * V});
        if (active) {
#pragma unroll
          for (int k4 = 0; k4 < B2P / 4; ++k4) {
            stv4(Fe + (q * B2P + 4 * k4) * V, F[4 * k4], F[4 * k4 + 1], F[4 * k4 + 2], F[4 * k4 + 3]);
          }
        }
      }
      if constexpr (sizeof(R) == 8) {
        loadA();
      }
      if constexpr (V == 1) {
        // face projections for the neighbor kernel (kernels_faceproj.cuh): F_f of an element is one contiguous
        // [9][B2P] block of its face buffer and leaves by a bulk store.  The store of face f-1 has to be done reading
        // its buffer before anybody passes the barrier below and overwrites it with F_{f+1}.
        if (wantProj) {
          fence_async_smem();
          if (f >= 1 && tid < E) {
            tma_store_wait_read();
          }
        }
      }
      __syncthreads();
      if constexpr (V == 1) {
        if (wantProj && tid < E && myElem >= 0) {
          tma_store(static_cast<R*>(const_cast<void*>(p.faceProj.at(myElem))) + f * kNQ * B2P,
                    myRQ + (f & 1) * kNQ * B2P, kNQ * B2P * sizeof(R));
          tma_store_commit();
        }
      }
      G::template rDivM<R>(f, acc, JChunks<R, B2P>{Fe, a});
    }

    // ---- Q_new staged per element in RD (the last reader of I passed the barrier of face 3)
    if (active) {
#pragma unroll
      for (int k4 = 0; k4 < BP / 4; ++k4) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          st4lane(RD + v * TEN + q * BP + 4 * k4, v, acc[4 * k4], acc[4 * k4 + 1], acc[4 * k4 + 2], acc[4 * k4 + 3]);
        }
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid < E && myElem >= 0) {
      tmaStoreTensor<R, BP>(static_cast<R*>(const_cast<void*>(p.dofs.at(myElem))), myRD, p.ldg);
    }
  }
  if (tid < E) {
    tma_store_wait_read();  // shared memory must stay valid until the bulk stores have read it
  }
  __syncthreads();
  }  // batch loop
}

// ---------------------------------------------------------------------------------------
// k_neigh shared-memory plan per group: V staged neighbor tiles (TMA target, re-used for the
// elements' own Q at the end) and two interleaved face buffers.
template <typename R, typename G, int NP>
struct NeighPlan {
  static constexpr int V = Pk<R>::V;
  static constexpr int E = NP * V;
  static constexpr int BP = G::kBP;
  static constexpr int B2P = G::kB2P;
  static constexpr int TEN = kNQ * BP;
  static constexpr int SEQ = padOdd16(V * TEN, sizeof(R));
  static constexpr int SEF = padOdd16(V * 2 * kNQ * B2P, sizeof(R));
  static constexpr int offN = 0;
  static constexpr int offF = NP * SEQ;
  static constexpr int endReals = NP * (SEQ + SEF);
  static constexpr size_t barOffset = (size_t(endReals) * sizeof(R) + 15) / 16 * 16;
  static constexpr size_t bytes = barOffset + 16;
};

#ifndef SSB_MINB_NEIGH
#define SSB_MINB_NEIGH 2  // resident CTAs per SM the register allocation of k_neigh is sized for
#endif
template <typename R, typename G, int NP>
__global__ void __launch_bounds__(blockThreads(NP), SSB_MINB_NEIGH) k_neigh(const NeighParams p) {
  using P = NeighPlan<R, G, NP>;
  using PT = typename Pk<R>::T;
  constexpr int V = P::V, E = P::E, BP = P::BP, B2P = P::B2P, TEN = P::TEN;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  R* const smem = reinterpret_cast<R*>(smemRaw);
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + P::barOffset);

  const int tid = threadIdx.x;
  const bool active = tid < kNQ * NP;
  const int g = active ? tid % NP : 0;
  const int q = active ? tid / NP : 0;
  if (tid == 0) {
    mbar_init(bar, E);
    mbar_fence_init();
  }
  __syncthreads();

  const long long nBatches = (p.n + E - 1) / E;
  uint32_t phase = 0;
  for (long long batch = blockIdx.x; batch < nBatches; batch += gridDim.x) {
  const long long slot = batch * E;

  // per-lane face metadata of this thread's group
  long long el[V];
  const R* NIg[V];
  uint8_t ft[V][4], fj[V][4], fh[V][4];
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const long long sl = slot + g * V + v;
    el[v] = sl < p.n ? (p.perm != nullptr ? p.perm[sl] : sl) : -1;
    NIg[v] = el[v] >= 0 ? static_cast<const R*>(p.neighInt.at(el[v])) : nullptr;
    if (el[v] >= 0) {
      uchar4 t = make_uchar4(0, 0, 0, 0);
      if (p.faceTypes != nullptr) {
        t = *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * el[v]);
      }
      ft[v][0] = t.x; ft[v][1] = t.y; ft[v][2] = t.z; ft[v][3] = t.w;
      const uint2 rr = *reinterpret_cast<const uint2*>(p.faceRel + 8 * el[v]);
      fj[v][0] = rr.x & 0xff; fh[v][0] = (rr.x >> 8) & 0xff; fj[v][1] = (rr.x >> 16) & 0xff; fh[v][1] = (rr.x >> 24) & 0xff;
      fj[v][2] = rr.y & 0xff; fh[v][2] = (rr.y >> 8) & 0xff; fj[v][3] = (rr.y >> 16) & 0xff; fh[v][3] = (rr.y >> 24) & 0xff;
    } else {
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        ft[v][f] = 255; fj[v][f] = 0; fh[v][f] = 0;
      }
    }
  }

  // loader threads (tid < E): element slot+tid = lane tid%V of group tid/V
  long long myElem = -1;
  R* myTile = nullptr;
  if (tid < E) {
    const long long sl = slot + tid;
    if (sl < p.n) {
      myElem = p.perm != nullptr ? p.perm[sl] : sl;
    }
    myTile = smem + P::offN + (tid / V) * P::SEQ + (tid % V) * TEN;
  }
  uchar4 myFt = make_uchar4(0, 0, 0, 0);  // face types of the element this loader thread moves
  if (myElem >= 0 && p.faceTypes != nullptr) {
    myFt = *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * myElem);
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
    const uint8_t t = f == 0 ? myFt.x : (f == 1 ? myFt.y : (f == 2 ? myFt.z : myFt.w));
    if (t != 0) {
      return nullptr;
    }
    if (p.nbPtrs != nullptr) {
      return static_cast<const R*>(p.nbPtrs[4 * myElem + f]);
    }
    return static_cast<const R*>(p.idofs.at(p.nbIdx[4 * myElem + f]));
  };
  if (tid < E) {
    loadTile(neighborOf(0));
  }

  PT acc[BP];
#pragma unroll
  for (int k = 0; k < BP; ++k) {
    acc[k] = pk_zero<R>();
  }
  const R* const tiles = smem + P::offN + g * P::SEQ;
  const SrcSplit<R> nbCol{tiles + q * BP, tiles + TEN + q * BP};

#pragma unroll
  for (int f = 0; f < 4; ++f) {
    // column q of AminusT_f straight from global memory; latency hidden by rT/fP
    PT a[kNQ];
#pragma unroll
    for (int c = 0; c < kNQ; ++c) {
      R t[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        t[v] = ft[v][f] == 0 ? NIg[v][81 * f + 9 * q + c] : R(0);
      }
      a[c] = pk_make(t);
    }
    mbar_wait(bar, (phase + f) & 1);
    PT F[B2P];
    {
      // F = fP_h . (rT_j . I_nb); the lanes of a group share the instruction stream, so when
      // their (j, h) differ (rare after sorting) the operators are applied once per lane.
      PT Gt[B2P];
#pragma unroll
      for (int i = 0; i < B2P; ++i) {
        Gt[i] = pk_zero<R>();
        F[i] = pk_zero<R>();
      }
      G::template rT<R>(fj[0][f], Gt, nbCol);
      if constexpr (V > 1) {
        if (fj[V - 1][f] != fj[0][f]) {
          PT G1[B2P];
#pragma unroll
          for (int i = 0; i < B2P; ++i) {
            G1[i] = pk_zero<R>();
          }
          G::template rT<R>(fj[V - 1][f], G1, nbCol);
#pragma unroll
          for (int i = 0; i < B2P; ++i) {
            Gt[i] = pk_sel01(Gt[i], G1[i]);
          }
        }
      }
      G::template fP<R>(fh[0][f], F, Gt);
      if constexpr (V > 1) {
        if (fh[V - 1][f] != fh[0][f]) {
          PT F1[B2P];
#pragma unroll
          for (int i = 0; i < B2P; ++i) {
            F1[i] = pk_zero<R>();
          }
          G::template fP<R>(fh[V - 1][f], F1, Gt);
#pragma unroll
          for (int i = 0; i < B2P; ++i) {
            F[i] = pk_sel01(F[i], F1[i]);
          }
        }
      }
    }
    {
      // faces without a regular neighbor: no tile was loaded, F was computed from stale shared memory -> exact zero
      bool keep[V];
#pragma unroll
      for (int v = 0; v < V; ++v) {
        keep[v] = ft[v][f] == 0;
        if (p.nbPtrs != nullptr && el[v] >= 0) {
          keep[v] = keep[v] && p.nbPtrs[4 * el[v] + f] != nullptr;   // a NULL FaceNeighbors entry loads no tile either
        }
      }
#pragma unroll
      for (int i = 0; i < B2P; ++i) {
        F[i] = pk_keep(F[i], keep);
      }
    }
    R* const Fe = smem + P::offF + g * P::SEF + (f & 1) * kNQ * B2P * V;
    if (active) {
#pragma unroll
      for (int k4 = 0; k4 < B2P / 4; ++k4) {
        stv4(Fe + (q * B2P + 4 * k4) * V, F[4 * k4], F[4 * k4 + 1], F[4 * k4 + 2], F[4 * k4 + 3]);
      }
    }
    __syncthreads();  // F visible; the tiles are no longer read
    if (tid < E) {
      // next neighbor, or finally the element's own Q
      loadTile(f < 3 ? neighborOf(f + 1) : (myElem >= 0 ? static_cast<const R*>(p.dofs.at(myElem)) : nullptr));
    }
    G::template rDivM<R>(f, acc, JChunks<R, B2P>{Fe, a});
  }

  // Q += acc: the own Q tensors arrived in the tiles
  mbar_wait(bar, phase);
  phase ^= 1;
  if (active) {
#pragma unroll
    for (int k4 = 0; k4 < BP / 4; ++k4) {
      PT t[4];
      nbCol.load4(t, 4 * k4);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        st4lane(const_cast<R*>(tiles) + v * TEN + q * BP + 4 * k4, v, pk_add(t[0], acc[4 * k4]),
                pk_add(t[1], acc[4 * k4 + 1]), pk_add(t[2], acc[4 * k4 + 2]), pk_add(t[3], acc[4 * k4 + 3]));
      }
    }
  }
  fence_async_smem();
  __syncthreads();
  if (tid < E && myElem >= 0) {
    tmaStoreTensor<R, BP>(static_cast<R*>(const_cast<void*>(p.dofs.at(myElem))), myTile, p.ldg);
    tma_store_wait_read();
  }
  __syncthreads();
  }  // batch loop
}

// ---------------------------------------------------------------------------------------
// Taylor sum I = sum_d c_d dQ(d): pure streaming (1 638 reals moved per 1 764 flop at order 6).  The derivative family
// of an element is one contiguous block and I another, so both cross the SM boundary as 1-D TMA bulk copies (fully
// coalesced whatever the row counts of the dQ(d) are); the sum itself reads shared memory.  A CTA owns E elements.
constexpr int kTaylorE = 4;
template <typename R>
__global__ void __launch_bounds__(256) k_taylor(const TaylorParams p) {
  constexpr int E = kTaylorE;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  R* const smem = reinterpret_cast<R*>(smemRaw);
  const int tenI = kNQ * p.ldg;
  const int per = p.derSize + tenI;                      // reals per element: family | I
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + (size_t(E) * per * sizeof(R) + 15) / 16 * 16);
  const int tid = threadIdx.x;
  if (tid == 0) {
    mbar_init(bar, E);
    mbar_fence_init();
  }
  __syncthreads();
  const long long first = static_cast<long long>(blockIdx.x) * E;
  if (tid < E) {
    const long long el = first + tid;
    if (el < p.n) {
      mbar_arrive_expect_tx(bar, p.derSize * sizeof(R));
      tma_load(smem + tid * per, p.derivs.at(el), p.derSize * sizeof(R), bar);
    } else {
      mbar_arrive_expect_tx(bar, 0);
    }
  }
  mbar_wait(bar, 0);
  const int vecPerCol = p.ldg / 4;
  const int items = E * kNQ * vecPerCol;
  for (int i = tid; i < items; i += 256) {
    const int e = i / (kNQ * vecPerCol);
    const int rem = i - e * (kNQ * vecPerCol);
    const int c = rem / vecPerCol;
    const int k0 = 4 * (rem - c * vecPerCol);
    const R* der = smem + e * per;
    R out[4] = {R(0), R(0), R(0), R(0)};
    if (first + e < p.n) {
      for (int d = 0; d < p.order; ++d) {
        if (k0 < p.rows[d]) {
          R t[4];
          ld4(t, der + p.derOffset[d] + c * p.derLd[d] + k0);
          const R cd = static_cast<R>(p.coeffs[d]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            out[j] = (k0 + j < p.rows[d]) ? fma(cd, t[j], out[j]) : out[j];
          }
        }
      }
    }
    st4(smem + e * per + p.derSize + c * p.ldg + k0, out[0], out[1], out[2], out[3]);
  }
  fence_async_smem();
  __syncthreads();
  if (tid < E && first + tid < p.n) {
    tma_store(const_cast<void*>(p.idofs.at(first + tid)), smem + tid * per + p.derSize, tenI * sizeof(R));
    tma_store_commit();
    tma_store_wait_read();
  }
}

// ---------------------------------------------------------------------------------------
// LTS buffer accumulation: buffers[i] = reset ? idofs[i] : buffers[i] + idofs[i] for cells whose neighbors in a
// larger time cluster read the sum of two sub-steps (LtsSetup::accumulateBuffers; TimeCluster.cpp:508-530,
// device::algorithms kernel_streamBatchedData / kernel_accumulateBatchedData).

template <typename R>
__global__ void __launch_bounds__(256) k_lts_accumulate(const AccumulateParams p) {
  const long long total = p.n * p.vec4PerElem;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const long long el = i / p.vec4PerElem;
    const int off = 4 * static_cast<int>(i - el * p.vec4PerElem);
    R* dst = static_cast<R*>(const_cast<void*>(p.buffers.at(el))) + off;
    R a[4];
    ld4(a, static_cast<const R*>(p.idofs.at(el)) + off);
    if (!p.reset) {
      R b[4];
      ld4(b, dst);
      a[0] += b[0]; a[1] += b[1]; a[2] += b[2]; a[3] += b[3];
    }
    st4(dst, a[0], a[1], a[2], a[3]);
  }
}

}  // namespace ssb
