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
#pragma once
#include "common.cuh"

namespace ssb {

// ---------------------------------------------------------------------------------------
struct LocalParams {
  ElemArray dofs;      // Q, in/out                         (LTS::Dofs)
  ElemArray idofs;     // I: out if ader, in if !ader       (LTS::Buffers / Idofs scratch)
  ElemArray derivs;    // optional derivative family, out   (LTS::Derivatives)
  ElemArray localInt;  // LocalIntegrationData: star[3][24], nApNm1[4][81]
  const uint8_t* faceTypes;  // [n][4] FaceType, may be null (= all regular)
  const int* perm;           // optional processing order
  long long n;
  int ldg;             // leading dimension of Q/I columns in global memory (reals)
  int derOffset[8];    // offset of dQ(d) inside the derivative family (reals)
  int derLd[8];        // leading dimension of dQ(d) in global memory (reals)
  double coeffs[8];    // Taylor integration coefficients c_d
};

struct NeighParams {
  ElemArray dofs;       // Q in/out
  ElemArray neighInt;   // NeighboringIntegrationData: nAmNm1[4][81]
  const void* const* nbPtrs;  // [n][4] pointers to the neighbors' I (LTS::FaceNeighbors) or null
  const long long* nbIdx;     // else [n][4] indices into idofs
  ElemArray idofs;
  const uint8_t* faceTypes;   // [n][4] or null
  const uint8_t* faceRel;     // [n][4][2] = {neighbor's face j, orientation h}
  const int* perm;
  long long n;
  int ldg;
};

struct TaylorParams {
  ElemArray derivs;
  ElemArray idofs;
  long long n;
  int ldg;
  int order;
  int rows[8];
  int derOffset[8];
  int derLd[8];
  double coeffs[8];
};

constexpr int kNQ = 9;
__host__ __device__ constexpr int blockThreads(int E) { return (kNQ * E + 31) / 32 * 32; }

// Shared-memory plan of k_local, usable on host and device.
template <typename R, typename G, int E>
struct LocalPlan {
  static constexpr int BP = G::kBP;
  static constexpr int B2P = G::kB2P;
  static constexpr int SEQ = padOdd16(kNQ * BP, sizeof(R));                 // Q / I per element
  static constexpr int D0LD = pad4i(G::derRows(1));
  static constexpr int SED0 = padOdd16(kNQ * D0LD, sizeof(R));              // dQ(odd)
  static constexpr int SEF = padOdd16(kNQ * B2P, sizeof(R));                // one face buffer
  static constexpr int SELI = padOdd16(72 + 324, sizeof(R));                // star + nApNm1
  static constexpr bool kFInQ = 2 * SEF <= SEQ;                             // F aliases the Q region
  static constexpr int offQ = 0;
  static constexpr int offI = offQ + E * SEQ;
  static constexpr int offD0 = offI + E * SEQ;
  static constexpr int offLI = offD0 + E * SED0;
  static constexpr int offF = kFInQ ? offQ : offLI + E * SELI;
  static constexpr int endReals = kFInQ ? offLI + E * SELI : offF + 2 * E * SEF;
  static constexpr size_t barOffset = (size_t(endReals) * sizeof(R) + 15) / 16 * 16;
  static constexpr size_t bytes = barOffset + 16;
};

// ---------------------------------------------------------------------------------------
// Per-thread context handed to the generated ADER driver (gen::ader_run): it tells the
// generated code where dQ(d-1) lives in shared memory and what to do with a finished dQ(d).
template <typename R, typename G, int E>
struct AderCtx {
  using P = LocalPlan<R, G, E>;
  static constexpr int BP = P::BP;
  static constexpr int O = G::kOrder;
  const LocalParams& p;
  R (&s)[3][3];
  R* Qe; R* Ie; R* D0e; R* D1e; R* smem;
  int r[3]; int q; int tid; bool active; bool wantDer; long long myElem;
  __device__ __forceinline__ R* buf(int d) const { return d == 0 ? Qe : ((d & 1) ? D0e : D1e); }
  __device__ __forceinline__ static constexpr int ld(int d) { return d == 0 ? BP : pad4i(G::derRows(d)); }
  __device__ __forceinline__ const R* src(int d, int c) const { return buf(d - 1) + r[c] * ld(d - 1); }
      template <int N>
  __device__ __forceinline__ void finish(int d, R (&out)[N]) {
    const R cd = static_cast<R>(p.coeffs[d]);
    R* icol = Ie + q * BP;
    if (active) {  // padding threads alias (e=0,q=0): they must not repeat the read-modify-write
#pragma unroll
      for (int v = 0; v < N / 4; ++v) {
        R t[4];
        ld4(t, icol + 4 * v);
        st4(icol + 4 * v, fma(cd, out[4 * v], t[0]), fma(cd, out[4 * v + 1], t[1]),
            fma(cd, out[4 * v + 2], t[2]), fma(cd, out[4 * v + 3], t[3]));
      }
    }
    if (d < O - 1 || wantDer) {
      // the bulk stores issued earlier must have finished reading the buffer we overwrite
      if (wantDer && tid < E) {
        tma_store_wait_read();
      }
      R* dcol = buf(d) + q * N;
      if (active) {
#pragma unroll
        for (int v = 0; v < N / 4; ++v) {
          st4(dcol + 4 * v, out[4 * v], out[4 * v + 1], out[4 * v + 2], out[4 * v + 3]);
        }
      }
      if (wantDer) {
        fence_async_smem();
      }
      __syncthreads();
      if (wantDer && tid < E && myElem >= 0) {
        R* dst = static_cast<R*>(const_cast<void*>(p.derivs.at(myElem))) + p.derOffset[d];
        const R* srcb = ((d & 1) ? smem + P::offD0 + tid * P::SED0 : smem + P::offQ + tid * P::SEQ);
        for (int c = 0; c < kNQ; ++c) {
          tma_store(dst + c * p.derLd[d], srcb + c * N, N * sizeof(R));
        }
        tma_store_commit();
      }
    }
  }
};

// ---------------------------------------------------------------------------------------
template <typename R, typename G, int E, bool ADER, bool LOCAL>
__global__ void __launch_bounds__(blockThreads(E)) k_local(const LocalParams p) {
  using P = LocalPlan<R, G, E>;
  constexpr int BP = P::BP, B2P = P::B2P, O = G::kOrder;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  R* const smem = reinterpret_cast<R*>(smemRaw);
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + P::barOffset);

  const int tid = threadIdx.x;
  const bool active = tid < kNQ * E;
  const int e = active ? tid % E : 0;
  const int q = active ? tid / E : 0;
  const long long slot = static_cast<long long>(blockIdx.x) * E;

  R* const Qe = smem + P::offQ + e * P::SEQ;
  R* const Ie = smem + P::offI + e * P::SEQ;
  R* const D0e = smem + P::offD0 + e * P::SED0;
  R* const D1e = Qe;  // dQ(even >= 2) reuses the Q region (Q column is held in registers)
  const R* const LIe = smem + P::offLI + e * P::SELI;

  if (tid == 0) {
    mbar_init(bar, E);
    mbar_fence_init();
  }
  __syncthreads();

  // ---- TMA loads: thread t < E fetches element slot+t
  long long myElem = -1;
  if (tid < E) {
    const long long s = slot + tid;
    if (s < p.n) {
      myElem = p.perm != nullptr ? p.perm[s] : s;
    }
    uint32_t bytes = 0;
    if (myElem >= 0) {
      const uint32_t colBytes = BP * sizeof(R);
      bytes = (72 + 324) * sizeof(R) + kNQ * colBytes * ((ADER ? 1 : 2));
      mbar_arrive_expect_tx(bar, bytes);
      R* const qDst = smem + P::offQ + tid * P::SEQ;
      R* const iDst = smem + P::offI + tid * P::SEQ;
      const R* qSrc = static_cast<const R*>(p.dofs.at(myElem));
      tma_load(smem + P::offLI + tid * P::SELI, p.localInt.at(myElem), (72 + 324) * sizeof(R), bar);
      if (p.ldg == BP) {
        tma_load(qDst, qSrc, kNQ * colBytes, bar);
        if (!ADER) {
          tma_load(iDst, p.idofs.at(myElem), kNQ * colBytes, bar);
        }
      } else {
        for (int c = 0; c < kNQ; ++c) {
          tma_load(qDst + c * BP, qSrc + c * p.ldg, colBytes, bar);
        }
        if (!ADER) {
          const R* iSrc = static_cast<const R*>(p.idofs.at(myElem));
          for (int c = 0; c < kNQ; ++c) {
            tma_load(iDst + c * BP, iSrc + c * p.ldg, colBytes, bar);
          }
        }
      }
    } else {
      mbar_arrive_expect_tx(bar, 0);
    }
  }
  mbar_wait(bar, 0);

  // ---- per-thread star coefficients of column q (3 source columns, zero padded)
  const int r0 = kStarRow[q][0], r1 = kStarRow[q][1], r2 = kStarRow[q][2];
  R s[3][3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    s[j][0] = LIe[24 * j + kStarOff[q] + 0];
    s[j][1] = LIe[24 * j + kStarOff[q] + 1];
    s[j][2] = kStarCnt[q] == 3 ? LIe[24 * j + kStarOff[q] + 2] : R(0);
  }

  R acc[BP];  // own Q column
#pragma unroll
  for (int v = 0; v < BP / 4; ++v) {
    R t[4];
    ld4(t, Qe + q * BP + 4 * v);
    acc[4 * v + 0] = t[0]; acc[4 * v + 1] = t[1]; acc[4 * v + 2] = t[2]; acc[4 * v + 3] = t[3];
  }

  if constexpr (ADER) {
    const bool wantDer = p.derivs.present();
    // I = c0 * Q (own column)
    {
      const R c0 = static_cast<R>(p.coeffs[0]);
      if (active) {
#pragma unroll
        for (int v = 0; v < BP / 4; ++v) {
          st4(Ie + q * BP + 4 * v, c0 * acc[4 * v], c0 * acc[4 * v + 1], c0 * acc[4 * v + 2], c0 * acc[4 * v + 3]);
        }
      }
    }
    // dQ(0) = Q copy into the derivative family (Time.cpp:109-116)
    if (wantDer && tid < E && myElem >= 0) {
      R* dst = static_cast<R*>(const_cast<void*>(p.derivs.at(myElem))) + p.derOffset[0];
      for (int c = 0; c < kNQ; ++c) {
        tma_store(dst + c * p.derLd[0], smem + P::offQ + tid * P::SEQ + c * BP, BP * sizeof(R));
      }
      tma_store_commit();
    }

    // the dQ(0) bulk store reads the Q region, which step 2 overwrites: handled by the
    // wait in finish(1) (executed before the barrier that precedes step 2's writes).
    AderCtx<R, G, E> ctx{p, s, Qe, Ie, D0e, D1e, smem, {r0, r1, r2}, q, tid, active, wantDer, myElem};
    G::template ader_run<R>(ctx);

    if (wantDer && tid < E) {
      tma_store_wait_read();
    }
    fence_async_smem();
    __syncthreads();  // I complete and visible; D buffers free
    if (tid < E && myElem >= 0) {
      R* dst = static_cast<R*>(const_cast<void*>(p.idofs.at(myElem)));
      const R* srcb = smem + P::offI + tid * P::SEQ;
      if (p.ldg == BP) {
        tma_store(dst, srcb, kNQ * BP * sizeof(R));
      } else {
        for (int c = 0; c < kNQ; ++c) {
          tma_store(dst + c * p.ldg, srcb + c * BP, BP * sizeof(R));
        }
      }
      tma_store_commit();
    }
  } else {
    if constexpr (P::kFInQ) {
      __syncthreads();  // everyone has its Q column before F overwrites the Q region
    }
  }

  if constexpr (LOCAL) {
    // ---- volume: Q += sum_i kDivM_i . (I . star_i)
    G::template volume<R>(acc, Ie + r0 * BP, Ie + r1 * BP, Ie + r2 * BP, s);

    // ---- local flux: Q += rDivM_f . ((fMrT_f . I) . AplusT_f), skipped on DR faces
    uint8_t ft[4] = {0, 0, 0, 0};
    if (p.faceTypes != nullptr) {
      const long long sl = slot + e;
      if (sl < p.n) {
        const long long el = p.perm != nullptr ? p.perm[sl] : sl;
        const uchar4 t = *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * el);
        ft[0] = t.x; ft[1] = t.y; ft[2] = t.z; ft[3] = t.w;
      }
    }
    R* const Fbase = smem + P::offF + e * P::SEF;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      R F[B2P];
      G::template fMrT<R>(f, F, Ie + q * BP);
      R* const Fe = Fbase + (f & 1) * E * P::SEF;
      if (active) {
#pragma unroll
        for (int v = 0; v < B2P / 4; ++v) {
          st4(Fe + q * B2P + 4 * v, F[4 * v], F[4 * v + 1], F[4 * v + 2], F[4 * v + 3]);
        }
      }
      __syncthreads();
      const bool skip = ft[f] == 3;  // FaceType::DynamicRupture
      R J[B2P];
#pragma unroll
      for (int c = 0; c < kNQ; ++c) {
        const R a = skip ? R(0) : LIe[72 + 81 * f + 9 * q + c];
#pragma unroll
        for (int v = 0; v < B2P / 4; ++v) {
          R t[4];
          ld4(t, Fe + c * B2P + 4 * v);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            J[4 * v + i] = c == 0 ? a * t[i] : fma(a, t[i], J[4 * v + i]);
          }
        }
      }
      G::template rDivM<R>(f, acc, J);
    }

    // ---- write back Q
    __syncthreads();  // all F reads done (F may alias the Q region)
    if (active) {
#pragma unroll
      for (int v = 0; v < BP / 4; ++v) {
        st4(Qe + q * BP + 4 * v, acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid < E && myElem >= 0) {
      R* dst = static_cast<R*>(const_cast<void*>(p.dofs.at(myElem)));
      const R* srcb = smem + P::offQ + tid * P::SEQ;
      if (p.ldg == BP) {
        tma_store(dst, srcb, kNQ * BP * sizeof(R));
      } else {
        for (int c = 0; c < kNQ; ++c) {
          tma_store(dst + c * p.ldg, srcb + c * BP, BP * sizeof(R));
        }
      }
      tma_store_commit();
    }
  }
  if (tid < E) {
    tma_store_wait_read();  // shared memory must stay valid until the bulk stores have read it
  }
}

// ---------------------------------------------------------------------------------------
template <typename R, typename G, int E>
struct NeighPlan {
  static constexpr int BP = G::kBP;
  static constexpr int B2P = G::kB2P;
  static constexpr int SEQ = padOdd16(kNQ * BP, sizeof(R));
  static constexpr int SEF = padOdd16(kNQ * B2P, sizeof(R));
  static constexpr int SENI = padOdd16(324, sizeof(R));
  static constexpr int offA = 0;
  static constexpr int offB = offA + E * SEQ;
  static constexpr int offF = offB + E * SEQ;
  static constexpr int offNI = offF + 2 * E * SEF;
  static constexpr int endReals = offNI + E * SENI;
  static constexpr size_t barOffset = (size_t(endReals) * sizeof(R) + 15) / 16 * 16;
  static constexpr size_t bytes = barOffset + 32;
};

template <typename R, typename G, int E>
__global__ void __launch_bounds__(blockThreads(E)) k_neigh(const NeighParams p) {
  using P = NeighPlan<R, G, E>;
  constexpr int BP = P::BP, B2P = P::B2P;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  R* const smem = reinterpret_cast<R*>(smemRaw);
  uint64_t* const bars = reinterpret_cast<uint64_t*>(smemRaw + P::barOffset);  // [0]=Q+NI, [1]=bufB, [2]=bufA

  const int tid = threadIdx.x;
  const bool active = tid < kNQ * E;
  const int e = active ? tid % E : 0;
  const int q = active ? tid / E : 0;
  const long long slot = static_cast<long long>(blockIdx.x) * E;

  if (tid == 0) {
    mbar_init(&bars[0], E);
    mbar_init(&bars[1], E);
    mbar_init(&bars[2], E);
    mbar_fence_init();
  }
  __syncthreads();

  // element handled by this thread's lane (for face metadata) and by the loader thread
  long long el = -1;
  {
    const long long sl = slot + e;
    if (sl < p.n) {
      el = p.perm != nullptr ? p.perm[sl] : sl;
    }
  }
  uint8_t ft[4] = {0, 0, 0, 0};
  uint8_t fj[4] = {0, 0, 0, 0};
  uint8_t fh[4] = {0, 0, 0, 0};
  if (el >= 0) {
    if (p.faceTypes != nullptr) {
      const uchar4 t = *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * el);
      ft[0] = t.x; ft[1] = t.y; ft[2] = t.z; ft[3] = t.w;
    }
    const uint2 rr = *reinterpret_cast<const uint2*>(p.faceRel + 8 * el);
    fj[0] = rr.x & 0xff; fh[0] = (rr.x >> 8) & 0xff; fj[1] = (rr.x >> 16) & 0xff; fh[1] = (rr.x >> 24) & 0xff;
    fj[2] = rr.y & 0xff; fh[2] = (rr.y >> 8) & 0xff; fj[3] = (rr.y >> 16) & 0xff; fh[3] = (rr.y >> 24) & 0xff;
  } else {
    ft[0] = ft[1] = ft[2] = ft[3] = 255;
  }

  const uint32_t colBytes = BP * sizeof(R);
  // loader: neighbor tensor of face f of element slot+tid -> buffer
  auto loadNeighbor = [&](int f, R* dstBase, uint64_t* bar) {
    // called by threads tid < E only; here e == tid
    const R* src = nullptr;
    if (el >= 0 && ft[f] == 0) {
      if (p.nbPtrs != nullptr) {
        src = static_cast<const R*>(p.nbPtrs[4 * el + f]);
      } else {
        src = static_cast<const R*>(p.idofs.at(p.nbIdx[4 * el + f]));
      }
    }
    if (src == nullptr) {
      mbar_arrive_expect_tx(bar, 0);
      return;
    }
    mbar_arrive_expect_tx(bar, kNQ * colBytes);
    R* dst = dstBase + tid * P::SEQ;
    if (p.ldg == BP) {
      tma_load(dst, src, kNQ * colBytes, bar);
    } else {
      for (int c = 0; c < kNQ; ++c) {
        tma_load(dst + c * BP, src + c * p.ldg, colBytes, bar);
      }
    }
  };

  R* const bufA = smem + P::offA;
  R* const bufB = smem + P::offB;
  if (tid < E) {
    if (el >= 0) {
      mbar_arrive_expect_tx(&bars[0], kNQ * colBytes + 324 * sizeof(R));
      const R* qSrc = static_cast<const R*>(p.dofs.at(el));
      R* qDst = bufA + tid * P::SEQ;
      if (p.ldg == BP) {
        tma_load(qDst, qSrc, kNQ * colBytes, &bars[0]);
      } else {
        for (int c = 0; c < kNQ; ++c) {
          tma_load(qDst + c * BP, qSrc + c * p.ldg, colBytes, &bars[0]);
        }
      }
      tma_load(smem + P::offNI + tid * P::SENI, p.neighInt.at(el), 324 * sizeof(R), &bars[0]);
    } else {
      mbar_arrive_expect_tx(&bars[0], 0);
    }
    loadNeighbor(0, bufB, &bars[1]);
  }

  mbar_wait(&bars[0], 0);
  R acc[BP];
#pragma unroll
  for (int v = 0; v < BP / 4; ++v) {
    R t[4];
    ld4(t, bufA + e * P::SEQ + q * BP + 4 * v);
    acc[4 * v + 0] = t[0]; acc[4 * v + 1] = t[1]; acc[4 * v + 2] = t[2]; acc[4 * v + 3] = t[3];
  }
  __syncthreads();  // bufA free
  if (tid < E) {
    loadNeighbor(1, bufA, &bars[2]);
  }

  const R* const NIe = smem + P::offNI + e * P::SENI;
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    R* const buf = (f & 1) ? bufA : bufB;
    uint64_t* const bar = (f & 1) ? &bars[2] : &bars[1];
    mbar_wait(bar, (f >> 1) & 1);
    const bool regular = ft[f] == 0;
    R F[B2P];
    {
      R Gt[B2P];
#pragma unroll
      for (int i = 0; i < B2P; ++i) {
        Gt[i] = R(0);
        F[i] = R(0);
      }
      if (regular) {
        G::template rT<R>(fj[f], Gt, buf + e * P::SEQ + q * BP);
        G::template fP<R>(fh[f], F, Gt);
      }
    }
    R* const Fe = smem + P::offF + (f & 1) * E * P::SEF + e * P::SEF;
    if (active) {
#pragma unroll
      for (int v = 0; v < B2P / 4; ++v) {
        st4(Fe + q * B2P + 4 * v, F[4 * v], F[4 * v + 1], F[4 * v + 2], F[4 * v + 3]);
      }
    }
    __syncthreads();  // F visible; `buf` no longer read
    if (f + 2 < 4 && tid < E) {
      loadNeighbor(f + 2, buf, bar);
    }
    R J[B2P];
#pragma unroll
    for (int c = 0; c < kNQ; ++c) {
      const R a = regular ? NIe[81 * f + 9 * q + c] : R(0);
#pragma unroll
      for (int v = 0; v < B2P / 4; ++v) {
        R t[4];
        ld4(t, Fe + c * B2P + 4 * v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          J[4 * v + i] = c == 0 ? a * t[i] : fma(a, t[i], J[4 * v + i]);
        }
      }
    }
    G::template rDivM<R>(f, acc, J);
  }

  // write back through bufB (its last reader finished before the barrier of face 2)
  if (active) {
#pragma unroll
    for (int v = 0; v < BP / 4; ++v) {
      st4(bufB + e * P::SEQ + q * BP + 4 * v, acc[4 * v], acc[4 * v + 1], acc[4 * v + 2], acc[4 * v + 3]);
    }
  }
  fence_async_smem();
  __syncthreads();
  if (tid < E && el >= 0) {
    R* dst = static_cast<R*>(const_cast<void*>(p.dofs.at(el)));
    const R* srcb = bufB + tid * P::SEQ;
    if (p.ldg == BP) {
      tma_store(dst, srcb, kNQ * colBytes);
    } else {
      for (int c = 0; c < kNQ; ++c) {
        tma_store(dst + c * p.ldg, srcb + c * BP, colBytes);
      }
    }
    tma_store_commit();
    tma_store_wait_read();
  }
}

// ---------------------------------------------------------------------------------------
// Taylor sum I = sum_d c_d dQ(d): pure streaming, one thread per (element, column, 4 rows).
template <typename R>
__global__ void __launch_bounds__(256) k_taylor(const TaylorParams p) {
  const int vecPerCol = (p.rows[0] + 3) / 4;
  const long long total = p.n * kNQ * vecPerCol;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += 256ll * gridDim.x) {
    const long long el = i / (kNQ * vecPerCol);
    const int rem = static_cast<int>(i - el * (kNQ * vecPerCol));
    const int c = rem / vecPerCol;
    const int k0 = 4 * (rem - c * vecPerCol);
    const R* der = static_cast<const R*>(p.derivs.at(el));
    R out[4] = {R(0), R(0), R(0), R(0)};
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
    R* dst = static_cast<R*>(const_cast<void*>(p.idofs.at(el))) + c * p.ldg + k0;
    st4(dst, out[0], out[1], out[2], out[3]);
  }
}

}  // namespace ssb
