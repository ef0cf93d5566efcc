// Generation 5 of the predictor / local kernels: PERSISTENT, SOFTWARE-PIPELINED CTAs.
//
//   k_ader5  : Cauchy-Kovalevskaya predictor  (Spacetime::computeBatchedAder, src/Kernels/LinearCK/Time.cpp:142-217)
//   k_local5 : volume + local flux            (Local::computeBatchedIntegral, src/Kernels/LinearCK/Local.cpp:255-395)
//
// Why (profiles/r2_ncu_fp_summary.txt, ncu source page): in generation 3 a CTA loads its batch, waits, computes,
// stores and exits -- 40-50 % of the warp samples sit in the first and last 5 % of the instructions (the mbarrier
// wait behind a cold DRAM round trip, the store drain, the CTA turnover), so on average only half of the 16 resident
// warps issue.  Here a CTA stays resident and walks its batches; the shared-memory regions of generation 3 are re-used
// so that the NEXT batch's input is in flight while the current batch is computed, with NO extra shared memory:
//   ader : Q (region RQ) is dead after derivative step 1 (61 % of the FMAs) except for its rows >= LD1, which are
//          kept in registers -> the next Q is prefetched into RQ during steps 2..O-1; I is staged in RD and leaves by
//          an asynchronous bulk store that has the whole step 1 of the next batch to drain.
//   local: the update is accumulated from zero and leaves as a bulk REDUCE-ADD into Q in global memory
//          (cp.reduce.async.bulk .add, SASS UBLKRED: the addition Q + update happens once, in L2, with the same
//          rounding as an FADD/DADD) -> Q is never loaded; I (region RD) is dead after the last fMrT product and the
//          next I is prefetched there while the last rDivM product and the epilogue run.
// The bulk copies of a batch are issued by lanes spread over ALL warps of the CTA (not by warp 0 alone): a warp that
// lags behind the others runs through the instruction stream on its own and costs a second fetch stream.
// One element per thread column (scalar registers); float and double.
#pragma once
#include "common.cuh"
#include "params.cuh"
#include "kernels_elastic.cuh"

namespace ssb {

// bulk reduce-add shared -> global (bulk async-group completion), SASS UBLKRED.G.S.ADD.{F32,F64}.RN
__device__ __forceinline__ void tma_reduce_add(float* gdst, const float* smemSrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smemSrc)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add(double* gdst, const double* smemSrc, uint32_t bytes) {
  asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f64 [%0], [%1], %2;" ::"l"(gdst),
               "r"(smem_u32(smemSrc)), "r"(bytes)
               : "memory");
}
template <typename R, int BP>
__device__ __forceinline__ void tmaReduceTensor(R* dst, const R* src, int ldg) {
  if (ldg == BP) {
    tma_reduce_add(dst, src, kNQ * BP * sizeof(R));
  } else {
    for (int c = 0; c < kNQ; ++c) {
      tma_reduce_add(dst + c * ldg, src + c * BP, BP * sizeof(R));
    }
  }
  tma_store_commit();
}

// The element a loader lane moves: loaders are lanes 0..ceil(E/NW)-1 of every warp.
template <int E, int THREADS>
struct LoaderMap {
  static constexpr int NW = THREADS / 32;
  static constexpr int LPW = (E + NW - 1) / NW;
  __device__ __forceinline__ static int elem(int tid) {
    const int lane = tid & 31, warp = tid >> 5;
    const int e = lane * NW + warp;
    return (lane < LPW && e < E) ? e : -1;
  }
};

template <typename R, typename G, int E>
struct AderCtx5 {
  using P = LocalPlan<R, G, E>;
  static constexpr int BP = P::BP;
  static constexpr int O = G::kOrder;
  const LocalParams& p;
  R (&s)[3][3];
  R (&sNext)[3][3];
  R (&icol)[P::LD1];
  R* RQ; R* RD;
  int r[3]; int q; bool active; bool wantDer;
  const long long (&el)[1];
  // pipeline state
  uint64_t* bar; int le; R* myRQ; long long nextElem; bool haveNext; const R* nextLI; R* derBase;
  __device__ __forceinline__ R* buf(int d) const { return d == 0 ? RQ : ((d & 1) ? RD : RD + P::offD2); }
  __device__ __forceinline__ static constexpr int ld(int d) { return d == 0 ? BP : pad4i(G::derRows(d)); }
  __device__ __forceinline__ const R* src(int d, int c) const { return buf(d - 1) + r[c] * ld(d - 1); }
  template <int N>
  __device__ __forceinline__ void finish(int d, R (&out)[N]) {
    if (d == 1) {
      // RD still feeds the bulk store of the previous batch's I, RQ is read by step 1 of every thread
      if (le >= 0) {
        tma_store_wait_read();
      }
      __syncthreads();
      if (haveNext) {
        if (le >= 0) {
          if (nextElem >= 0) {
            mbar_arrive_expect_tx(bar, P::TEN * sizeof(R));
            tmaLoadTensor<R, BP>(myRQ, static_cast<const R*>(p.dofs.at(nextElem)), p.ldg, bar);
          } else {
            mbar_arrive_expect_tx(bar, 0);
          }
        }
        // star coefficients of the next batch: in flight during steps 2..O-1
#pragma unroll
        for (int j = 0; j < 3; ++j) {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            sNext[j][c] = (nextLI != nullptr && c < kStarCnt[q]) ? nextLI[24 * j + kStarOff[q] + c] : R(0);
          }
        }
      }
    }
    const R cd = static_cast<R>(p.coeffs[d]);
#pragma unroll
    for (int k = 0; k < N; ++k) {
      icol[k] = fma(cd, out[k], icol[k]);
    }
    if (derBase != nullptr) {
      // derivative family (LTS / dynamic rupture consumers): straight to global memory; a NULL table entry means the
      // element keeps no derivatives (src/Kernels/LinearCK/Time.cpp:150-160)
      R* dst = derBase + p.derOffset[d] + q * p.derLd[d];
#pragma unroll
      for (int k4 = 0; k4 < N / 4; ++k4) {
        st4(dst + 4 * k4, out[4 * k4], out[4 * k4 + 1], out[4 * k4 + 2], out[4 * k4 + 3]);
      }
    }
    if (d < O - 1) {
      R* dcol = buf(d) + q * N;
      if (active) {
#pragma unroll
        for (int k4 = 0; k4 < N / 4; ++k4) {
          st4(dcol + 4 * k4, out[4 * k4], out[4 * k4 + 1], out[4 * k4 + 2], out[4 * k4 + 3]);
        }
      }
      __syncthreads();
    }
  }
};

#ifndef SSB_MINB_LOCAL5
#define SSB_MINB_LOCAL5 2
#endif

template <typename R, typename G, int E>
__global__ void __launch_bounds__(blockThreads(E), SSB_MINB_LOCAL5) k_ader5(const LocalParams p) {
  static_assert(Pk<R>::V == 1, "generation 5 is scalar");
  using P = LocalPlan<R, G, E>;
  using LM = LoaderMap<E, blockThreads(E)>;
  constexpr int BP = P::BP, TEN = P::TEN, LD1 = P::LD1;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  R* const smem = reinterpret_cast<R*>(smemRaw);
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + P::barOffset);

  const int tid = threadIdx.x;
  const bool active = tid < kNQ * E;
  const int g = active ? tid % E : 0;
  const int q = active ? tid / E : 0;
  R* const RQ = smem + P::offQ + g * P::SEQ;
  R* const RD = smem + P::offD + g * P::SED;
  const int le = LM::elem(tid);
  R* const myRQ = smem + P::offQ + (le >= 0 ? le : 0) * P::SEQ;
  R* const myRD = smem + P::offD + (le >= 0 ? le : 0) * P::SED;

  if (tid == 0) {
    mbar_init(bar, E);
    mbar_fence_init();
  }
  __syncthreads();

  const long long nBatches = (p.n + E - 1) / E;
  auto elemOf = [&](long long batch, int e) -> long long {
    const long long sl = batch * E + e;
    return (batch < nBatches && sl < p.n) ? (p.perm != nullptr ? p.perm[sl] : sl) : -1;
  };
  const bool wantDer = p.derivs.present();
  const int r0 = kStarRow[q][0], r1 = kStarRow[q][1], r2 = kStarRow[q][2];

  long long batch = blockIdx.x;
  if (batch >= nBatches) {
    return;
  }
  // ---- prologue: the first batch's Q and star coefficients
  long long myElem = le >= 0 ? elemOf(batch, le) : -1;
  if (le >= 0) {
    if (myElem >= 0) {
      mbar_arrive_expect_tx(bar, TEN * sizeof(R));
      tmaLoadTensor<R, BP>(myRQ, static_cast<const R*>(p.dofs.at(myElem)), p.ldg, bar);
    } else {
      mbar_arrive_expect_tx(bar, 0);
    }
  }
  long long el[1] = {active ? elemOf(batch, g) : -1};
  R s[3][3];
  {
    const R* LI = el[0] >= 0 ? static_cast<const R*>(p.localInt.at(el[0])) : nullptr;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        s[j][c] = (LI != nullptr && c < kStarCnt[q]) ? LI[24 * j + kStarOff[q] + c] : R(0);
      }
    }
  }

  uint32_t phase = 0;
  for (; batch < nBatches; batch += gridDim.x) {
    const long long nextBatch = batch + gridDim.x;
    const bool haveNext = nextBatch < nBatches;
    const long long nextMy = (haveNext && le >= 0) ? elemOf(nextBatch, le) : -1;
    const long long nextEl = (haveNext && active) ? elemOf(nextBatch, g) : -1;
    const R* nextLI = nextEl >= 0 ? static_cast<const R*>(p.localInt.at(nextEl)) : nullptr;

    mbar_wait(bar, phase);
    phase ^= 1;

    // dQ(0) = Q copy into the derivative family (Time.cpp:109-116), from the staged tensor
    if (wantDer && le >= 0 && myElem >= 0) {
      R* const base = static_cast<R*>(const_cast<void*>(p.derivs.at(myElem)));
      if (base != nullptr) {
        R* dst = base + p.derOffset[0];
        for (int c = 0; c < kNQ; ++c) {
          tma_store(dst + c * p.derLd[0], myRQ + c * BP, BP * sizeof(R));
        }
        tma_store_commit();
      }
    }
    R* const derBase = (wantDer && active && el[0] >= 0) ? static_cast<R*>(const_cast<void*>(p.derivs.at(el[0]))) : nullptr;
    R icol[LD1];
    R itail[BP - LD1];  // rows >= LD1 never receive a derivative contribution: I = c0 * Q there
    {
      const R c0 = static_cast<R>(p.coeffs[0]);
      const R* qc = RQ + q * BP;
#pragma unroll
      for (int k4 = 0; k4 < BP / 4; ++k4) {
        R t[4];
        ld4(t, qc + 4 * k4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (4 * k4 + i < LD1) {
            icol[4 * k4 + i] = c0 * t[i];
          } else {
            itail[4 * k4 + i - LD1] = c0 * t[i];
          }
        }
      }
    }
    R sNext[3][3];
    AderCtx5<R, G, E> ctx{p, s, sNext, icol, RQ, RD, {r0, r1, r2}, q, active, wantDer, el,
                          bar, le, myRQ, nextMy, haveNext, nextLI, derBase};
    G::template ader_run<R>(ctx);

    __syncthreads();  // nobody reads dQ(>=1) any more: RD stages I
    if (active) {
      R* col = RD + q * BP;
#pragma unroll
      for (int k4 = 0; k4 < BP / 4; ++k4) {
        R t[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          t[i] = 4 * k4 + i < LD1 ? icol[(4 * k4 + i) < LD1 ? 4 * k4 + i : 0] : itail[(4 * k4 + i) >= LD1 ? 4 * k4 + i - LD1 : 0];
        }
        st4(col + 4 * k4, t[0], t[1], t[2], t[3]);
      }
    }
    fence_async_smem();
    __syncthreads();
    if (le >= 0 && myElem >= 0) {
      tmaStoreTensor<R, BP>(static_cast<R*>(const_cast<void*>(p.idofs.at(myElem))), myRD, p.ldg);
    }
    // roll over to the next batch
    myElem = nextMy;
    el[0] = nextEl;
    if (haveNext) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          s[j][c] = sNext[j][c];
        }
      }
    }
  }
  if (le >= 0) {
    tma_store_wait_read();
  }
}

// ---------------------------------------------------------------------------------------
template <typename R, typename G, int E>
__global__ void __launch_bounds__(blockThreads(E), SSB_MINB_LOCAL5) k_local5(const LocalParams p) {
  static_assert(Pk<R>::V == 1, "generation 5 is scalar");
  using P = LocalPlan<R, G, E>;
  using LM = LoaderMap<E, blockThreads(E)>;
  constexpr int BP = P::BP, B2P = P::B2P, TEN = P::TEN;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  R* const smem = reinterpret_cast<R*>(smemRaw);
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + P::barOffset);

  const int tid = threadIdx.x;
  const bool active = tid < kNQ * E;
  const int g = active ? tid % E : 0;
  const int q = active ? tid / E : 0;
  R* const RQ = smem + P::offQ + g * P::SEQ;  // face buffers F0|F1, finally the staged update
  R* const RD = smem + P::offD + g * P::SED;  // I
  const int le = LM::elem(tid);
  R* const myRQ = smem + P::offQ + (le >= 0 ? le : 0) * P::SEQ;
  R* const myRD = smem + P::offD + (le >= 0 ? le : 0) * P::SED;

  if (tid == 0) {
    mbar_init(bar, E);
    mbar_fence_init();
  }
  __syncthreads();

  const long long nBatches = (p.n + E - 1) / E;
  // element ids as 32-bit values inside the kernel (the launcher rejects batches of 2^31 elements and more): four
  // 64-bit ids are live across the whole batch, and the kernel sits at its register limit
  auto elemOf = [&](long long batch, int e) -> int {
    const long long sl = batch * E + e;
    return (batch < nBatches && sl < p.n) ? (p.perm != nullptr ? p.perm[sl] : static_cast<int>(sl)) : -1;
  };
  const bool wantProj = p.faceProj.present();
  const int r0 = kStarRow[q][0], r1 = kStarRow[q][1], r2 = kStarRow[q][2];

  long long batch = blockIdx.x;
  if (batch >= nBatches) {
    return;
  }
  int myElem = le >= 0 ? elemOf(batch, le) : -1;
  if (le >= 0) {
    if (myElem >= 0) {
      mbar_arrive_expect_tx(bar, TEN * sizeof(R));
      tmaLoadTensor<R, BP>(myRD, static_cast<const R*>(p.idofs.at(myElem)), p.ldg, bar);
    } else {
      mbar_arrive_expect_tx(bar, 0);
    }
  }
  int el = active ? elemOf(batch, g) : -1;
  const R* LI = el >= 0 ? static_cast<const R*>(p.localInt.at(el)) : nullptr;
  const int sCnt = kStarCnt[q], sOff = kStarOff[q];   // loop invariant: looked up once (they are loads from a table)
  R s[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s[j][c] = (LI != nullptr && c < sCnt) ? LI[24 * j + sOff + c] : R(0);
    }
  }
  auto faceTypesOf = [&](int e) -> uchar4 {
    return (p.faceTypes != nullptr && e >= 0) ? *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * e) : make_uchar4(0, 0, 0, 0);
  };
  uchar4 ft = faceTypesOf(el);

  uint32_t phase = 0;
  for (; batch < nBatches; batch += gridDim.x) {
    const long long nextBatch = batch + gridDim.x;
    const bool haveNext = nextBatch < nBatches;
    // the next batch's element ids, integration-data pointer and face types are requested at the top of face 2 and
    // consumed at face 3 / the bottom of the loop: early enough to hide the DRAM round trips (a load that is used right
    // away cost 9 % of the warp samples), late enough not to hold five registers through the volume stage
    int nextMy = -1, nextEl = -1;
    const R* nextLI = nullptr;
    uchar4 ftNext = make_uchar4(0, 0, 0, 0);
    bool skipFace[4];
    {
      const bool none = el < 0;  // FaceType::DynamicRupture (3) gets no local flux
      skipFace[0] = none || ft.x == 3; skipFace[1] = none || ft.y == 3;
      skipFace[2] = none || ft.z == 3; skipFace[3] = none || ft.w == 3;
    }

    mbar_wait(bar, phase);
    phase ^= 1;

    R acc[BP];
#pragma unroll
    for (int k = 0; k < BP; ++k) {
      acc[k] = R(0);
    }
    const R* const Ie = RD;
    // ---- volume: update = sum_i kDivM_i . (I . star_i)
    G::template volume<R>(acc, Ie + r0 * BP, Ie + r1 * BP, Ie + r2 * BP, s);

    // the bulk reduce-add of the previous batch's update has had the whole volume stage to finish reading RQ
    if (le >= 0) {
      tma_store_wait_read();
    }
    __syncthreads();

    // ---- local flux: update += rDivM_f . ((fMrT_f . I) . AplusT_f), skipped on DR faces
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      if (f == 2) {
        nextMy = (haveNext && le >= 0) ? elemOf(nextBatch, le) : -1;
        nextEl = (haveNext && active) ? elemOf(nextBatch, g) : -1;
        nextLI = nextEl >= 0 ? static_cast<const R*>(p.localInt.at(nextEl)) : nullptr;
        ftNext = faceTypesOf(nextEl);
      }
      R a[kNQ];
      auto loadA = [&]() {
#pragma unroll
        for (int c = 0; c < kNQ; ++c) {
          a[c] = skipFace[f] ? R(0) : LI[72 + 81 * f + 9 * q + c];
        }
      };
      if constexpr (sizeof(R) == 4) {
        loadA();
      }
      R* const Fe = RQ + (f & 1) * kNQ * B2P;
      {
        R F[B2P];
        G::template fMrT<R>(f, F, SrcPk<R>{Ie + q * BP});
        if (active) {
#pragma unroll
          for (int k4 = 0; k4 < B2P / 4; ++k4) {
            st4(Fe + q * B2P + 4 * k4, F[4 * k4], F[4 * k4 + 1], F[4 * k4 + 2], F[4 * k4 + 3]);
          }
        }
      }
      if constexpr (sizeof(R) == 8) {
        loadA();
      }
      // The bulk store of the face projection of face f-1 reads the OTHER buffer, which face f+1 overwrites after
      // the barrier: it has to be done before anybody passes.
      if (wantProj) {
        fence_async_smem();
        if (le >= 0 && f >= 1) {
          tma_store_wait_read();
        }
      }
      __syncthreads();
      if (wantProj && le >= 0 && myElem >= 0) {
        tma_store(static_cast<R*>(const_cast<void*>(p.faceProj.at(myElem))) + f * kNQ * B2P,
                  myRQ + (f & 1) * kNQ * B2P, kNQ * B2P * sizeof(R));
        tma_store_commit();
      }
      if (f == 3 && haveNext) {
        // I is dead (the last fMrT passed the barrier): the next batch's I lands in RD while rDivM_3 and the
        // epilogue run; its star coefficients travel to registers meanwhile
        if (le >= 0) {
          if (nextMy >= 0) {
            mbar_arrive_expect_tx(bar, TEN * sizeof(R));
            tmaLoadTensor<R, BP>(myRD, static_cast<const R*>(p.idofs.at(nextMy)), p.ldg, bar);
          } else {
            mbar_arrive_expect_tx(bar, 0);
          }
        }
        // ... and its star coefficients are pulled into the L2 now and loaded at the bottom of the loop, when the
        // accumulators are dead (loading them here spilled them to local memory behind a full-latency stall)
        if (nextLI != nullptr) {
#pragma unroll
          for (int j = 0; j < 3; ++j) {
            asm volatile("prefetch.global.L2 [%0];" ::"l"(nextLI + 24 * j + sOff));
          }
        }
      }
      G::template rDivM<R>(f, acc, JChunks<R, B2P>{Fe, a});
    }

    // ---- the update leaves as a bulk reduce-add into Q
    if (wantProj && le >= 0) {
      tma_store_wait_read();  // face projections of faces 2 and 3 have left RQ
    }
    __syncthreads();          // everybody is done reading the face buffers
    if (active) {
      R* col = RQ + q * BP;
#pragma unroll
      for (int k4 = 0; k4 < BP / 4; ++k4) {
        st4(col + 4 * k4, acc[4 * k4], acc[4 * k4 + 1], acc[4 * k4 + 2], acc[4 * k4 + 3]);
      }
    }
    fence_async_smem();
    __syncthreads();
    if (le >= 0 && myElem >= 0) {
      tmaReduceTensor<R, BP>(static_cast<R*>(const_cast<void*>(p.dofs.at(myElem))), myRQ, p.ldg);
    }
    myElem = nextMy;
    el = nextEl;
    LI = nextLI;
    ft = ftNext;
    if (haveNext) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          s[j][c] = (LI != nullptr && c < sCnt) ? LI[24 * j + sOff + c] : R(0);
        }
      }
    }
  }
  if (le >= 0) {
    tma_store_wait_read();
  }
}

}  // namespace ssb
