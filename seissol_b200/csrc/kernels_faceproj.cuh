// Neighbor flux from FACE PROJECTIONS (generation 5 of the neighbor kernel).
//
// Reference: Neighbor::computeBatchedNeighborsIntegral (src/Kernels/LinearCK/Neighbor.cpp:111-188),
//   Q_i += rDivM_i . fP_h . rT_j . I_nb . AminusT_i      (kernel::neighboringFlux(h,j,i), codegen/kernels/aderdg.py:414-461)
//
// The face mass matrix is diagonal and identical on the four faces, fMrT_f = D . rT_f (checked by the generator), so the
// LOCAL kernel already holds F_f = fMrT_f . I = D . (rT_f . I) for every face f of its element while it evaluates the
// local flux.  It stores these four [9][B2P] blocks (`face_proj`, 4 x 864 B at order 6 / float) next to I, and this
// kernel gathers the block of face j of each neighbor instead of the neighbor's whole I:
//   Q_i += rDivM_i . (fP_h D^-1) . (F_j,nb . AminusT_i)
// * 4 x 864 B gathered per element instead of 4 x 2016 B,
// * the 16 rT_j bodies (7.5 k instructions = 119 KB of a 206 KB kernel) and 39 % of the FMAs are gone,
// * the 9x9 flux-solver product is applied FIRST (it commutes with the left operators), so a thread keeps its column
//   in registers from the tile to the accumulator: no shared-memory exchange and no barrier between the stages.
// Rounding differs from the reference's evaluation order by a few ulp (tests hold the kernel to 2e-6 / 5e-15 against the
// oracle's reference order).
#pragma once
#include "common.cuh"
#include "params.cuh"
#include "kernels_elastic.cuh"
#include "kernels_elastic5.cuh"

namespace ssb {

template <typename R, typename G, int E>
struct NeighFpPlan {
  static constexpr int BP = G::kBP;
  static constexpr int B2P = G::kB2P;
  static constexpr int TEN = kNQ * BP;   // one element tensor
  static constexpr int PRJ = kNQ * B2P;  // one face projection
  static constexpr int AM = 4 * 81;      // nAmNm1[4][81]
  static_assert(3 * PRJ >= TEN, "the own Q tile re-uses the space of the projections of faces 0..2");
  // per element: projections of faces 0..3 | nAmNm1; stride padded to an odd multiple of 16 bytes
  static constexpr int SE = padOdd16(4 * PRJ + AM, sizeof(R));
  static constexpr int offA = 4 * PRJ;
  static constexpr size_t barOffset = (size_t(E) * SE * sizeof(R) + 15) / 16 * 16;
  static constexpr size_t bytes = barOffset + 16;
  static constexpr int threads = (kNQ * E + 31) / 32 * 32;
};

#ifndef SSB_MINB_NEIGHFP
#define SSB_MINB_NEIGHFP 2
#endif
template <typename R, typename G, int E>
__global__ void __launch_bounds__((NeighFpPlan<R, G, E>::threads), SSB_MINB_NEIGHFP) k_neigh_fp(const NeighParams p) {
  using P = NeighFpPlan<R, G, E>;
  constexpr int BP = P::BP, B2P = P::B2P, TEN = P::TEN, PRJ = P::PRJ;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  R* const smem = reinterpret_cast<R*>(smemRaw);
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + P::barOffset);  // bar[0]: projections + nAmNm1, bar[1]: own Q

  const int tid = threadIdx.x;
  const bool active = tid < kNQ * E;
  const int g = active ? tid % E : 0;
  const int q = active ? tid / E : 0;
  if (tid == 0) {
    mbar_init(bar, E);
    mbar_init(bar + 1, E);
    mbar_fence_init();
  }
  __syncthreads();

  const long long nBatches = (p.n + E - 1) / E;
  uint32_t phase = 0;
  for (long long batch = blockIdx.x; batch < nBatches; batch += gridDim.x) {
    const long long slot = batch * E;

    // ---- loader threads: element slot+tid.  One bulk copy per regular face (block j of the neighbor's projections)
    //      and one for nAmNm1.
    long long myElem = -1;
    R* const myTile = smem + (tid < E ? tid : 0) * P::SE;
    if (tid < E) {
      const long long sl = slot + tid;
      if (sl < p.n) {
        myElem = p.perm != nullptr ? p.perm[sl] : sl;
      }
      if (myElem >= 0) {
        uchar4 t = make_uchar4(0, 0, 0, 0);
        if (p.faceTypes != nullptr) {
          t = *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * myElem);
        }
        const uint8_t ftl[4] = {t.x, t.y, t.z, t.w};
        const uint2 rr = *reinterpret_cast<const uint2*>(p.faceRel + 8 * myElem);
        const int fjl[4] = {int(rr.x & 0xff), int((rr.x >> 16) & 0xff), int(rr.y & 0xff), int((rr.y >> 16) & 0xff)};
        const R* src[4];
        uint32_t bytes = P::AM * sizeof(R);
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          src[f] = nullptr;
          if (ftl[f] == 0) {
            const R* base = p.nbPtrs != nullptr ? static_cast<const R*>(p.nbPtrs[4 * myElem + f])
                                                : static_cast<const R*>(p.faceProj.at(p.nbIdx[4 * myElem + f]));
            if (base != nullptr) {
              src[f] = base + fjl[f] * PRJ;
              bytes += PRJ * sizeof(R);
            }
          }
        }
        mbar_arrive_expect_tx(bar, bytes);
        tma_load(myTile + P::offA, p.neighInt.at(myElem), P::AM * sizeof(R), bar);
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          if (src[f] != nullptr) {
            tma_load(myTile + f * PRJ, src[f], PRJ * sizeof(R), bar);
          }
        }
      } else {
        mbar_arrive_expect_tx(bar, 0);
      }
    }

    // ---- per-thread face metadata of element g
    const long long sl = slot + g;
    const long long el = sl < p.n ? (p.perm != nullptr ? p.perm[sl] : sl) : -1;
    bool keep[4] = {false, false, false, false};
    int fh[4] = {0, 0, 0, 0};
    if (el >= 0) {
      uchar4 t = make_uchar4(0, 0, 0, 0);
      if (p.faceTypes != nullptr) {
        t = *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * el);
      }
      const uint2 rr = *reinterpret_cast<const uint2*>(p.faceRel + 8 * el);
      fh[0] = (rr.x >> 8) & 0xff; fh[1] = (rr.x >> 24) & 0xff; fh[2] = (rr.y >> 8) & 0xff; fh[3] = (rr.y >> 24) & 0xff;
      keep[0] = t.x == 0; keep[1] = t.y == 0; keep[2] = t.z == 0; keep[3] = t.w == 0;
      if (p.nbPtrs != nullptr) {
#pragma unroll
        for (int f = 0; f < 4; ++f) {
          keep[f] = keep[f] && p.nbPtrs[4 * el + f] != nullptr;
        }
      }
    }

    R acc[BP];
#pragma unroll
    for (int k = 0; k < BP; ++k) {
      acc[k] = R(0);
    }
    const R* const tile = smem + g * P::SE;

    mbar_wait(bar, phase);
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      // x = column q of (F_nb . AminusT_f)
      R x[B2P];
#pragma unroll
      for (int k = 0; k < B2P; ++k) {
        x[k] = R(0);
      }
      const R* const A = tile + P::offA + 81 * f + 9 * q;
      const R* const Fnb = tile + f * PRJ;
#pragma unroll
      for (int c = 0; c < kNQ; ++c) {
        const R a = A[c];
#pragma unroll
        for (int k4 = 0; k4 < B2P / 4; ++k4) {
          R t[4];
          ld4(t, Fnb + c * B2P + 4 * k4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            x[4 * k4 + i] = fma(a, t[i], x[4 * k4 + i]);
          }
        }
      }
      // faces without a regular neighbor loaded no tile: x came from stale shared memory -> exact zero (select)
#pragma unroll
      for (int k = 0; k < B2P; ++k) {
        x[k] = keep[f] ? x[k] : R(0);
      }
      R F[B2P];
      G::template fPD<R>(fh[f], F, x);
      if (f == 3) {
        // everybody is done with the tiles of faces 0..2: the own Q tensor lands there while face 3 is computed
        __syncthreads();
        if (tid < E) {
          if (myElem >= 0) {
            mbar_arrive_expect_tx(bar + 1, TEN * sizeof(R));
            tmaLoadTensor<R, BP>(myTile, static_cast<const R*>(p.dofs.at(myElem)), p.ldg, bar + 1);
          } else {
            mbar_arrive_expect_tx(bar + 1, 0);
          }
        }
      }
      G::template rDivMr<R>(f, acc, F);
    }

    // ---- Q += acc
    mbar_wait(bar + 1, phase);
    phase ^= 1;
    if (active && el >= 0) {
      R* const col = const_cast<R*>(tile) + q * BP;
#pragma unroll
      for (int k4 = 0; k4 < BP / 4; ++k4) {
        R t[4];
        ld4(t, col + 4 * k4);
        st4(col + 4 * k4, t[0] + acc[4 * k4], t[1] + acc[4 * k4 + 1], t[2] + acc[4 * k4 + 2], t[3] + acc[4 * k4 + 3]);
      }
    }
    fence_async_smem();
    __syncthreads();
    if (tid < E && myElem >= 0) {
      tmaStoreTensor<R, BP>(static_cast<R*>(const_cast<void*>(p.dofs.at(myElem))), myTile, p.ldg);
      tma_store_wait_read();
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Taylor sum + face projection for LTS neighbors that are read through their derivatives (src/Kernels/TimeCommon.cpp:89-125,
// Time::evaluateBatched Time.cpp:257-292): I = sum_d c_d dQ(d) of the NEIGHBOR's family over the reader's (sub-)interval,
// then F = fMrT_j I for the one face j the reader sees -- the block the face-projection neighbor kernel gathers.
// One item per (cell, face) that needs it; E items per CTA; the family arrives by one bulk copy, I lives in shared memory.
constexpr int kTaylorFpE = 8;
template <typename R, typename G>
__global__ void __launch_bounds__(256) k_taylor_fp(const TaylorParams p) {
  constexpr int E = kTaylorFpE;
  constexpr int BP = G::kBP, B2P = G::kB2P, PRJ = kNQ * B2P;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  R* const smem = reinterpret_cast<R*>(smemRaw);
  const int tenI = kNQ * BP;
  const int per = p.derSize + tenI + PRJ;                  // reals per item: family | I | projection
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
  // ---- I = sum_d c_d dQ(d), column-major [9][BP] behind the family
  const int vecPerCol = BP / 4;
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
    st4(smem + e * per + p.derSize + c * BP + k0, out[0], out[1], out[2], out[3]);
  }
  __syncthreads();
  // ---- F = fMrT_j . I: thread (e, q) owns column q
  if (tid < E * kNQ) {
    const int e = tid / kNQ, q = tid - e * kNQ;
    if (first + e < p.n) {
      const int j = p.faceJ[first + e];
      R F[B2P];
      G::template fMrT<R>(j, F, SrcPk<R>{smem + e * per + p.derSize + q * BP});
      R* dst = smem + e * per + p.derSize + tenI + q * B2P;
#pragma unroll
      for (int k4 = 0; k4 < B2P / 4; ++k4) {
        st4(dst + 4 * k4, F[4 * k4], F[4 * k4 + 1], F[4 * k4 + 2], F[4 * k4 + 3]);
      }
    }
  }
  fence_async_smem();
  __syncthreads();
  if (tid < E && first + tid < p.n) {
    tma_store(const_cast<void*>(p.idofs.at(first + tid)), smem + tid * per + p.derSize + tenI, PRJ * sizeof(R));
    tma_store_commit();
    tma_store_wait_read();
  }
}

// =======================================================================================================================
// k_neigh_fp2: the same product as k_neigh_fp on PERSISTENT, SOFTWARE-PIPELINED CTAs.
//
// ncu on k_neigh_fp (profiles/r2_ncu_fp_summary.txt): 40 % of the warp samples sit in the mbarrier wait behind the
// gather (perm -> face metadata -> neighbor index -> bulk copy: three dependent DRAM round trips), long_scoreboard
// 3.45 stall cycles per issue, DRAM at 41 % -- the kernel is latency-bound, not bandwidth- or issue-bound.  Here
//  * a CTA stays resident and walks its batches; the four face slots of an element form a ring: the tiles of faces
//    0/1 of batch b+1 are requested as soon as every thread is done with them in batch b (barrier S1), the tiles of
//    faces 2/3 of batch b at the top of batch b (they are needed only after faces 0 and 1 are computed);
//  * the face metadata of batch b+1 is fetched at the top of batch b and the element ids (perm) two batches ahead,
//    so that no dependent load chain is on the critical path;
//  * the 9x9 flux-solver columns go straight from global memory to registers one face ahead (they need no staging: a
//    thread reads 36 contiguous bytes of its own element) -> 3.4 KB of shared memory per element instead of 4.75 KB:
//    28 elements per CTA, 16 warps per SM;
//  * the update leaves as a bulk reduce-add into Q (kernels_elastic5.cuh): Q is never loaded; the update is staged in
//    the slots of faces 2/3 (+ a small tail), which are dead after barrier S3.
template <typename R, typename G, int E>
struct NeighFp2Plan {
  static constexpr int BP = G::kBP;
  static constexpr int B2P = G::kB2P;
  static constexpr int TEN = kNQ * BP;
  static constexpr int PRJ = kNQ * B2P;
  static constexpr int TAIL = TEN > 2 * PRJ ? pad4i(TEN - 2 * PRJ) : 0;
  static constexpr int SE = padOdd16(4 * PRJ + TAIL, sizeof(R));
  static constexpr int offStage = 2 * PRJ;
  static constexpr size_t barOffset = (size_t(E) * SE * sizeof(R) + 15) / 16 * 16;
  static constexpr size_t bytes = barOffset + 32;
  static constexpr int threads = (kNQ * E + 31) / 32 * 32;
};

#ifndef SSB_MINB_NEIGHFP2
#define SSB_MINB_NEIGHFP2 2
#endif
template <typename R, typename G, int E>
__global__ void __launch_bounds__((NeighFp2Plan<R, G, E>::threads), SSB_MINB_NEIGHFP2) k_neigh_fp2(const NeighParams p) {
  using P = NeighFp2Plan<R, G, E>;
  using LM = LoaderMap<E, P::threads>;
  constexpr int BP = P::BP, B2P = P::B2P, PRJ = P::PRJ;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  R* const smem = reinterpret_cast<R*>(smemRaw);
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + P::barOffset);  // [0]: slot 0, [1]: slot 1, [2]: slots 2+3

  const int tid = threadIdx.x;
  const bool active = tid < kNQ * E;
  const int g = active ? tid % E : 0;
  const int q = active ? tid / E : 0;
  const int le = LM::elem(tid);
  R* const myTile = smem + (le >= 0 ? le : 0) * P::SE;
  const R* const tile = smem + g * P::SE;
  if (tid == 0) {
    mbar_init(bar, E);
    mbar_init(bar + 1, E);
    mbar_init(bar + 2, E);
    mbar_fence_init();
  }
  __syncthreads();

  const long long nBatches = (p.n + E - 1) / E;
  const long long stride = gridDim.x;
  auto slotElem = [&](long long batch, int e) -> long long {
    const long long sl = batch * E + e;
    return (batch < nBatches && sl < p.n) ? (p.perm != nullptr ? p.perm[sl] : sl) : -1;
  };
  // compute-thread metadata: bits 2f..2f+1 = orientation h of face f, bit 8+f = face f has a regular neighbor
  auto faceMeta = [&](long long elem) -> unsigned {
    if (elem < 0) {
      return 0u;
    }
    uchar4 t = make_uchar4(0, 0, 0, 0);
    if (p.faceTypes != nullptr) {
      t = *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * elem);
    }
    const uint2 rr = *reinterpret_cast<const uint2*>(p.faceRel + 8 * elem);
    unsigned m = ((rr.x >> 8) & 3u) | (((rr.x >> 24) & 3u) << 2) | (((rr.y >> 8) & 3u) << 4) | (((rr.y >> 24) & 3u) << 6);
    bool k0 = t.x == 0, k1 = t.y == 0, k2 = t.z == 0, k3 = t.w == 0;
    if (p.nbPtrs != nullptr) {
      k0 = k0 && p.nbPtrs[4 * elem + 0] != nullptr; k1 = k1 && p.nbPtrs[4 * elem + 1] != nullptr;
      k2 = k2 && p.nbPtrs[4 * elem + 2] != nullptr; k3 = k3 && p.nbPtrs[4 * elem + 3] != nullptr;
    }
    return m | (k0 ? 0x100u : 0u) | (k1 ? 0x200u : 0u) | (k2 ? 0x400u : 0u) | (k3 ? 0x800u : 0u);
  };
  // loader metadata: the four projection blocks this element gathers (nullptr: nothing to load)
  auto faceSrc = [&](long long elem, const R* (&src)[4]) {
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      src[f] = nullptr;
    }
    if (elem < 0) {
      return;
    }
    uchar4 t = make_uchar4(0, 0, 0, 0);
    if (p.faceTypes != nullptr) {
      t = *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * elem);
    }
    const uint8_t ftl[4] = {t.x, t.y, t.z, t.w};
    const uint2 rr = *reinterpret_cast<const uint2*>(p.faceRel + 8 * elem);
    const int fjl[4] = {int(rr.x & 0xff), int((rr.x >> 16) & 0xff), int(rr.y & 0xff), int((rr.y >> 16) & 0xff)};
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      if (ftl[f] == 0) {
        const R* base = p.nbPtrs != nullptr ? static_cast<const R*>(p.nbPtrs[4 * elem + f])
                                            : static_cast<const R*>(p.faceProj.at(p.nbIdx[4 * elem + f]));
        src[f] = base != nullptr ? base + fjl[f] * PRJ : nullptr;
      }
    }
  };
  auto issue = [&](int slot, const R* s, uint64_t* b, uint32_t extra) {
    // one arrival per loader lane and barrier phase; `extra`: bytes of a second copy signalled on the same barrier
    if (s != nullptr) {
      mbar_arrive_expect_tx(b, PRJ * sizeof(R) + extra);
      tma_load(myTile + slot * PRJ, s, PRJ * sizeof(R), b);
    } else {
      mbar_arrive_expect_tx(b, extra);
    }
  };

  long long batch = blockIdx.x;
  if (batch >= nBatches) {
    return;
  }
  // ---- prologue (exposed once per CTA): metadata of the first batch, tiles of faces 0/1, element ids of the second
  long long el = active ? slotElem(batch, g) : -1;
  unsigned hk = faceMeta(el);
  long long myElem = le >= 0 ? slotElem(batch, le) : -1;
  const R* src[4];
  faceSrc(le >= 0 ? myElem : -1, src);
  if (le >= 0) {
    issue(0, src[0], bar, 0);
    issue(1, src[1], bar + 1, 0);
  }
  long long elN = active ? slotElem(batch + stride, g) : -1;
  long long myElemN = le >= 0 ? slotElem(batch + stride, le) : -1;
  R a[kNQ];
  auto loadA = [&](long long elem, unsigned meta, int f) {
    const R* A = elem >= 0 ? static_cast<const R*>(p.neighInt.at(elem)) + 81 * f + 9 * q : nullptr;
    const bool k = (meta >> (8 + f)) & 1u;
#pragma unroll
    for (int c = 0; c < kNQ; ++c) {
      a[c] = k ? A[c] : R(0);
    }
  };
  loadA(el, hk, 0);

  uint32_t phase = 0;
  for (; batch < nBatches; batch += stride) {
    const bool haveNext = batch + stride < nBatches;
    // tiles of faces 2/3 of THIS batch: their slots staged the previous update, whose bulk store has to be done reading
    if (le >= 0) {
      tma_store_wait_read();
      if (src[2] != nullptr) {
        mbar_arrive_expect_tx(bar + 2, PRJ * sizeof(R) * (src[3] != nullptr ? 2 : 1));
        tma_load(myTile + 2 * PRJ, src[2], PRJ * sizeof(R), bar + 2);
        if (src[3] != nullptr) {
          tma_load(myTile + 3 * PRJ, src[3], PRJ * sizeof(R), bar + 2);
        }
      } else {
        issue(3, src[3], bar + 2, 0);
      }
    }
    // metadata of the next batch (its element ids arrived a batch ago): consumed at barrier S1
    const unsigned hkN = faceMeta(elN);
    const R* srcN[4];
    faceSrc(le >= 0 ? myElemN : -1, srcN);

    R acc[BP];
#pragma unroll
    for (int k = 0; k < BP; ++k) {
      acc[k] = R(0);
    }
    long long elN2 = -1, myElemN2 = -1;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      if (f < 3) {
        mbar_wait(bar + f, phase);   // f = 2: slots 2 and 3
      }
      // x = column q of (F_nb . AminusT_f)
      R x[B2P];
#pragma unroll
      for (int k = 0; k < B2P; ++k) {
        x[k] = R(0);
      }
      const R* const Fnb = tile + f * PRJ;
#pragma unroll
      for (int c = 0; c < kNQ; ++c) {
#pragma unroll
        for (int k4 = 0; k4 < B2P / 4; ++k4) {
          R t[4];
          ld4(t, Fnb + c * B2P + 4 * k4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            x[4 * k4 + i] = fma(a[c], t[i], x[4 * k4 + i]);
          }
        }
      }
      // faces without a regular neighbor loaded no tile: x came from stale shared memory -> exact zero (select)
      const bool keep = (hk >> (8 + f)) & 1u;
#pragma unroll
      for (int k = 0; k < B2P; ++k) {
        x[k] = keep ? x[k] : R(0);
      }
      // the flux-solver column of the next face (of the next element after face 3) travels during fPD / rDivM
      if (f < 3) {
        loadA(el, hk, f + 1);
      } else {
        loadA(elN, hkN, 0);
      }
      if (f == 1) {
        __syncthreads();  // S1: slots 0 and 1 are dead -> the next batch's tiles of faces 0/1
        if (haveNext && le >= 0) {
          issue(0, srcN[0], bar, 0);
          issue(1, srcN[1], bar + 1, 0);
        }
      }
      if (f == 3) {
        __syncthreads();  // S3: slots 2 and 3 are dead -> they stage the update
        elN2 = active ? slotElem(batch + 2 * stride, g) : -1;
        myElemN2 = le >= 0 ? slotElem(batch + 2 * stride, le) : -1;
      }
      R F[B2P];
      G::template fPD<R>(int((hk >> (2 * f)) & 3u), F, x);
      G::template rDivMr<R>(f, acc, F);
    }

    // ---- the update leaves as a bulk reduce-add into Q
    if (active) {
      R* const col = const_cast<R*>(tile) + P::offStage + q * BP;
#pragma unroll
      for (int k4 = 0; k4 < BP / 4; ++k4) {
        st4(col + 4 * k4, acc[4 * k4], acc[4 * k4 + 1], acc[4 * k4 + 2], acc[4 * k4 + 3]);
      }
    }
    fence_async_smem();
    __syncthreads();
    if (le >= 0 && myElem >= 0) {
      tmaReduceTensor<R, BP>(static_cast<R*>(const_cast<void*>(p.dofs.at(myElem))), myTile + P::offStage, p.ldg);
    }
    // roll over
    phase ^= 1;
    el = elN; hk = hkN; myElem = myElemN;
    elN = elN2; myElemN = myElemN2;
#pragma unroll
    for (int f = 0; f < 4; ++f) {
      src[f] = srcN[f];
    }
  }
  if (le >= 0) {
    tma_store_wait_read();
  }
}

}  // namespace ssb
