// Hand-written sm_100a kernels for the viscoelastic2 (anelastic attenuation, M mechanisms) element update.
//
//   k_visco_ader  : Cauchy-Kovalevskaya predictor with memory variables.  Replaces linearckanelastic
//                   Spacetime::computeAder / computeBatchedAder (src/Kernels/LinearCKAnelastic/Time.cpp:50-190,
//                   generated kernel::derivative, codegen/kernels/equations/viscoelastic2.py:330-423).
//   k_visco_local : volumeExt + localFluxExt + local.  Replaces Local::computeIntegral / computeBatchedIntegral
//                   (src/Kernels/LinearCKAnelastic/Local.cpp:63-110; viscoelastic2.py:213-300).
//   k_visco_neigh : neighborFluxExt (all faces, all relations, one launch) + neighbor.  Replaces
//                   Neighbor::computeNeighborsIntegral (src/Kernels/LinearCKAnelastic/Neighbor.cpp:71-160).
//
// Mapping: a CTA owns E elements; thread (e, q), q = 0..14, owns column q of the EXTENDED tensor Qext (B x 15: nine
// elastic quantities + six anelastic source columns).  Threads q < 9 ("elastic role") additionally own column q of
// Q / I, threads q >= 9 ("anelastic role", p = q - 9) own the M memory-variable columns (p, m).  The constant
// operators are the same immediate-operand code as the elastic kernels (gen/ops_o5.cuh); the per-element matrices
// (star 9x15 sparse, AplusT/AminusT 9x15, E 6xMx9, w, W) are read from LocalIntegrationData /
// NeighboringIntegrationData in global memory.  Element tensors arrive by 1-D TMA bulk copies.
//
// Layouts (column-major, leading dimension ldg): Q, I [9][ldg]; Qane, Iane [M][6][ldg] (column p + 6 m);
// local_int = star[3][33] | nApNm1[4][135] | E[162] | w[M] | W[M]; neigh_int = nAmNm1[4][135] | w[M]
// (src/Initializer/Typedefs.h:187-210 with Equations/viscoelastic2/Model/IntegrationData.h).
#pragma once
#include "kernels_elastic.cuh"

namespace ssb {

constexpr int kVNE = 15;           // extended quantities
constexpr int kVNA = 6 * kVM;      // memory-variable columns
constexpr int kVStar = 33;
constexpr int kVA = 135;
constexpr int kVLocalInt = 3 * kVStar + 4 * kVA + 54 * kVM + 2 * kVM;  // 807
constexpr int kVNeighInt = 4 * kVA + kVM;                              // 543
constexpr int kVOffA = 3 * kVStar;
constexpr int kVOffE = kVOffA + 4 * kVA;
constexpr int kVOffw = kVOffE + 54 * kVM;
constexpr int kVOffW = kVOffw + kVM;

// CSC layout of the 9 x 15 star matrix (codegen/matrices/matrices_viscoelastic.xml): rows, count, offset per column
__device__ constexpr int kVStarRow[15][3] = {{6, 7, 8}, {6, 7, 8}, {6, 7, 8}, {6, 7, 6}, {7, 8, 7}, {6, 8, 6}, {0, 3, 5},
                                             {1, 3, 4}, {2, 4, 5}, {6, 6, 6}, {7, 7, 7}, {8, 8, 8}, {6, 7, 6}, {7, 8, 7},
                                             {6, 8, 6}};
__device__ constexpr int kVStarCnt[15] = {3, 3, 3, 2, 2, 2, 3, 3, 3, 1, 1, 1, 2, 2, 2};
__device__ constexpr int kVStarOff[15] = {0, 3, 6, 9, 11, 13, 15, 18, 21, 24, 25, 26, 27, 29, 31};

#ifndef SSB_VISCO_MINB
#define SSB_VISCO_MINB 2   // resident CTAs per SM the register allocation is sized for
#endif
template <typename R, typename G, int E>
struct ViscoPlan {
  static constexpr int BP = G::kBP;
  static constexpr int B2P = G::kB2P;
  static constexpr int TQ = kNQ * BP;    // Q / I tensor
  static constexpr int TA = kVNA * BP;   // Qane / Iane tensor
  static constexpr int TF = kNQ * B2P;   // face projection
  // predictor: dQ ping-pong | dQane | Iane
  static constexpr int aderReals = padOdd16(2 * TQ + 2 * TA, sizeof(R));
  // local: I | Iane | two face buffers | Q | Qane (own DOFs, staged for the final read-modify-write)
  static constexpr int localReals = padOdd16(2 * TQ + 2 * TA + 2 * TF, sizeof(R));
  // neigh: neighbor tile | two face buffers | Q | Qane
  static constexpr int neighReals = padOdd16(2 * TQ + TA + 2 * TF, sizeof(R));
  static constexpr int threads = (kVNE * E + 31) / 32 * 32;
  // two mbarriers behind the element data
  __host__ __device__ static constexpr size_t bytes(int perElem) { return (size_t(E) * perElem * sizeof(R) + 15) / 16 * 16 + 16; }
};

template <typename R>
__device__ __forceinline__ void tmaLoadCols(R* dst, const R* src, int cols, int rows, int ldg, uint64_t* bar) {
  if (ldg == rows) {
    tma_load(dst, src, cols * rows * sizeof(R), bar);
  } else {
    for (int c = 0; c < cols; ++c) {
      tma_load(dst + c * rows, src + c * ldg, rows * sizeof(R), bar);
    }
  }
}
template <typename R>
__device__ __forceinline__ void tmaStoreCols(R* dst, const R* src, int cols, int rows, int ldg) {
  if (ldg == rows) {
    tma_store(dst, src, cols * rows * sizeof(R));
  } else {
    for (int c = 0; c < cols; ++c) {
      tma_store(dst + c * ldg, src + c * rows, rows * sizeof(R));
    }
  }
  tma_store_commit();
}

// star coefficients of column q (3 matrices x up to 3 entries)
template <typename R>
__device__ __forceinline__ void viscoStar(R (&s)[3][3], const R* li, int q) {
#pragma unroll
  for (int j = 0; j < 3; ++j) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s[j][c] = (li != nullptr && c < kVStarCnt[q]) ? li[kVStar * j + kVStarOff[q] + c] : R(0);
    }
  }
}

// ---------------------------------------------------------------------------------------
template <typename R, typename G, int E>
__global__ void __launch_bounds__((ViscoPlan<R, G, E>::threads), SSB_VISCO_MINB) k_visco_ader(const ViscoLocalParams p) {
  using P = ViscoPlan<R, G, E>;
  constexpr int BP = P::BP, TQ = P::TQ, TA = P::TA, O = G::kOrder;
  constexpr int R1 = pad4i(G::derRows(1));  // rows of dQext(d): kDivMT lowers the polynomial degree
  extern __shared__ __align__(128) unsigned char smemRaw[];
  R* const smem = reinterpret_cast<R*>(smemRaw);
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + P::bytes(P::aderReals) - 16);
  const int tid = threadIdx.x;
  const bool active = tid < kVNE * E;
  const int e = active ? tid % E : 0;
  const int q = active ? tid / E : 0;
  R* const SQ = smem + e * P::aderReals;  // two dQ buffers
  R* const SA = SQ + 2 * TQ;              // dQane(d), updated in place
  R* const SI = SA + TA;                  // Iane
  if (tid == 0) {
    mbar_init(bar, E);
    mbar_fence_init();
  }
  __syncthreads();
  const long long el = static_cast<long long>(blockIdx.x) * E + e;
  const bool valid = active && el < p.n;
  if (tid < E) {
    if (el < p.n) {
      mbar_arrive_expect_tx(bar, (TQ + TA) * sizeof(R));
      tmaLoadCols<R>(SQ, static_cast<const R*>(p.dofs.at(el)), kNQ, BP, p.ldg, bar);
      tmaLoadCols<R>(SA, static_cast<const R*>(p.dofsAne.at(el)), kVNA, BP, p.ldg, bar);
    } else {
      mbar_arrive_expect_tx(bar, 0);
    }
  }
  const R* li = valid ? static_cast<const R*>(p.localInt.at(el)) : nullptr;
  R s[3][3];
  viscoStar<R>(s, li, q);
  const int r0 = kVStarRow[q][0], r1 = kVStarRow[q][1], r2 = kVStarRow[q][2];
  const bool roleA = q < kNQ;
  const int pa = roleA ? 0 : q - kNQ;
  R Eq[kVNA];  // elastic role: E[c][q], c = q' + 6 m  (E[q' + 6 (m + M p)], viscoelastic2.py:74-81)
  R wv[kVM], Wv[kVM];
#pragma unroll
  for (int c = 0; c < kVNA; ++c) {
    Eq[c] = (li != nullptr && roleA) ? li[kVOffE + kVNA * q + c] : R(0);
  }
#pragma unroll
  for (int m = 0; m < kVM; ++m) {
    wv[m] = li != nullptr ? li[kVOffw + m] : R(0);
    Wv[m] = li != nullptr ? li[kVOffW + m] : R(0);
  }
  mbar_wait(bar, 0);

  const R c0 = static_cast<R>(p.coeffs[0]);
  R icol[BP];
  if (roleA) {
#pragma unroll
    for (int k4 = 0; k4 < BP / 4; ++k4) {
      R t[4];
      ld4(t, SQ + q * BP + 4 * k4);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        icol[4 * k4 + i] = c0 * t[i];
      }
    }
  } else {
#pragma unroll
    for (int m = 0; m < kVM; ++m) {
      const int col = pa + 6 * m;
#pragma unroll
      for (int k4 = 0; k4 < BP / 4; ++k4) {
        R t[4];
        ld4(t, SA + col * BP + 4 * k4);
        if (active) {
          st4(SI + col * BP + 4 * k4, c0 * t[0], c0 * t[1], c0 * t[2], c0 * t[3]);
        }
      }
    }
  }

#pragma unroll 1
  for (int d = 1; d < O; ++d) {
    const R cd = static_cast<R>(p.coeffs[d]);
    const R* src = SQ + ((d - 1) & 1) * TQ;
    R* dst = SQ + (d & 1) * TQ;
    R ext[R1];
    G::template ader_step_1<R>(ext, src + r0 * BP, src + r1 * BP, src + r2 * BP, s);
    if (roleA) {
      // dQ(d)[:, q] = dQext(d)[:, q] + sum_c dQane(d-1)[:, c] E[c][q]
      R out[BP];
#pragma unroll
      for (int k = 0; k < BP; ++k) {
        out[k] = k < R1 ? ext[k] : R(0);
      }
#pragma unroll
      for (int c = 0; c < kVNA; ++c) {
#pragma unroll
        for (int k4 = 0; k4 < BP / 4; ++k4) {
          R t[4];
          ld4(t, SA + c * BP + 4 * k4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            out[4 * k4 + i] = fma(Eq[c], t[i], out[4 * k4 + i]);
          }
        }
      }
      if (active) {
#pragma unroll
        for (int k4 = 0; k4 < BP / 4; ++k4) {
          st4(dst + q * BP + 4 * k4, out[4 * k4], out[4 * k4 + 1], out[4 * k4 + 2], out[4 * k4 + 3]);
        }
      }
#pragma unroll
      for (int k = 0; k < BP; ++k) {
        icol[k] = fma(cd, out[k], icol[k]);
      }
    }
    __syncthreads();  // every elastic thread has read dQane(d-1)
    if (!roleA && active) {
      // dQane(d)[:, p, m] = w[m] dQext(d)[:, 9 + p] + W[m] dQane(d-1)[:, p, m];  Iane += c_d dQane(d)
#pragma unroll
      for (int m = 0; m < kVM; ++m) {
        const int col = pa + 6 * m;
#pragma unroll
        for (int k4 = 0; k4 < BP / 4; ++k4) {
          R t[4], u[4];
          ld4(t, SA + col * BP + 4 * k4);
          ld4(u, SI + col * BP + 4 * k4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int k = 4 * k4 + i;
            t[i] = Wv[m] * t[i];
            if (k < R1) {
              t[i] = fma(wv[m], ext[k < R1 ? k : 0], t[i]);
            }
            u[i] = fma(cd, t[i], u[i]);
          }
          st4(SA + col * BP + 4 * k4, t[0], t[1], t[2], t[3]);
          st4(SI + col * BP + 4 * k4, u[0], u[1], u[2], u[3]);
        }
      }
    }
    __syncthreads();  // dQ(d), dQane(d) complete
  }

  // I -> first dQ buffer (nobody reads the dQ buffers any more), then bulk stores
  if (roleA && active) {
#pragma unroll
    for (int k4 = 0; k4 < BP / 4; ++k4) {
      st4(SQ + q * BP + 4 * k4, icol[4 * k4], icol[4 * k4 + 1], icol[4 * k4 + 2], icol[4 * k4 + 3]);
    }
  }
  fence_async_smem();
  __syncthreads();
  if (tid < E && el < p.n) {
    tmaStoreCols<R>(static_cast<R*>(const_cast<void*>(p.idofs.at(el))), SQ, kNQ, BP, p.ldg);
    tmaStoreCols<R>(static_cast<R*>(const_cast<void*>(p.idofsAne.at(el))), SI, kVNA, BP, p.ldg);
    tma_store_wait_read();
  }
}

// ---------------------------------------------------------------------------------------
template <typename R, typename G, int E>
__global__ void __launch_bounds__((ViscoPlan<R, G, E>::threads), SSB_VISCO_MINB) k_visco_local(const ViscoLocalParams p) {
  using P = ViscoPlan<R, G, E>;
  constexpr int BP = P::BP, B2P = P::B2P, TQ = P::TQ, TA = P::TA, TF = P::TF;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  R* const smem = reinterpret_cast<R*>(smemRaw);
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + P::bytes(P::localReals) - 16);  // [0]: I, Iane  [1]: Q, Qane
  const int tid = threadIdx.x;
  const bool active = tid < kVNE * E;
  const int e = active ? tid % E : 0;
  const int q = active ? tid / E : 0;
  R* const SIq = smem + e * P::localReals;  // I
  R* const SI = SIq + TQ;                    // Iane
  R* const SF = SI + TA;                     // two face buffers
  R* const SQ = SF + 2 * TF;                 // own Q: arrives while volume and fluxes are computed
  R* const SA = SQ + TQ;                     // own Qane
  if (tid == 0) {
    mbar_init(bar, E);
    mbar_init(bar + 1, E);
    mbar_fence_init();
  }
  __syncthreads();
  const long long el = static_cast<long long>(blockIdx.x) * E + e;
  const bool valid = active && el < p.n;
  if (tid < E) {
    if (el < p.n) {
      mbar_arrive_expect_tx(bar, (TQ + TA) * sizeof(R));
      tmaLoadCols<R>(SIq, static_cast<const R*>(p.idofs.at(el)), kNQ, BP, p.ldg, bar);
      tmaLoadCols<R>(SI, static_cast<const R*>(p.idofsAne.at(el)), kVNA, BP, p.ldg, bar);
      mbar_arrive_expect_tx(bar + 1, (TQ + TA) * sizeof(R));
      tmaLoadCols<R>(SQ, static_cast<const R*>(p.dofs.at(el)), kNQ, BP, p.ldg, bar + 1);
      tmaLoadCols<R>(SA, static_cast<const R*>(p.dofsAne.at(el)), kVNA, BP, p.ldg, bar + 1);
    } else {
      mbar_arrive_expect_tx(bar, 0);
      mbar_arrive_expect_tx(bar + 1, 0);
    }
  }
  const R* li = valid ? static_cast<const R*>(p.localInt.at(el)) : nullptr;
  R s[3][3];
  viscoStar<R>(s, li, q);
  const int r0 = kVStarRow[q][0], r1 = kVStarRow[q][1], r2 = kVStarRow[q][2];
  const bool roleA = q < kNQ;
  const int pa = roleA ? 0 : q - kNQ;
  uchar4 ft = make_uchar4(0, 0, 0, 0);
  if (valid && p.faceTypes != nullptr) {
    ft = *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * el);
  }
  mbar_wait(bar, 0);

  R ext[BP];
#pragma unroll
  for (int k = 0; k < BP; ++k) {
    ext[k] = R(0);
  }
  G::template volume<R>(ext, SIq + r0 * BP, SIq + r1 * BP, SIq + r2 * BP, s);
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const uint8_t ty = f == 0 ? ft.x : (f == 1 ? ft.y : (f == 2 ? ft.z : ft.w));
    const bool skip = li == nullptr || ty == 3;  // FaceType::DynamicRupture gets no local flux
    R a[kNQ];
#pragma unroll
    for (int c = 0; c < kNQ; ++c) {
      a[c] = skip ? R(0) : li[kVOffA + kVA * f + kNQ * q + c];
    }
    R* const Fe = SF + (f & 1) * TF;
    if (roleA) {
      R F[B2P];
      G::template fMrT<R>(f, F, SrcPk<R>{SIq + q * BP});
      if (active) {
#pragma unroll
        for (int k4 = 0; k4 < B2P / 4; ++k4) {
          st4(Fe + q * B2P + 4 * k4, F[4 * k4], F[4 * k4 + 1], F[4 * k4 + 2], F[4 * k4 + 3]);
        }
      }
    }
    __syncthreads();
    G::template rDivM<R>(f, ext, JChunks<R, B2P>{Fe, a});
  }
  // kernel::local (viscoelastic2.py:259-277) on the staged own DOFs: every thread updates its own columns
  mbar_wait(bar + 1, 0);
  if (valid) {
    if (roleA) {
      R Eq[kVNA];
#pragma unroll
      for (int c = 0; c < kVNA; ++c) {
        Eq[c] = li[kVOffE + kVNA * q + c];
      }
#pragma unroll
      for (int c = 0; c < kVNA; ++c) {
#pragma unroll
        for (int k4 = 0; k4 < BP / 4; ++k4) {
          R t[4];
          ld4(t, SI + c * BP + 4 * k4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            ext[4 * k4 + i] = fma(Eq[c], t[i], ext[4 * k4 + i]);
          }
        }
      }
#pragma unroll
      for (int k4 = 0; k4 < BP / 4; ++k4) {
        R t[4];
        ld4(t, SQ + q * BP + 4 * k4);
        st4(SQ + q * BP + 4 * k4, t[0] + ext[4 * k4], t[1] + ext[4 * k4 + 1], t[2] + ext[4 * k4 + 2], t[3] + ext[4 * k4 + 3]);
      }
    } else {
#pragma unroll
      for (int m = 0; m < kVM; ++m) {
        const int col = pa + 6 * m;
        const R wm = li[kVOffw + m], Wm = li[kVOffW + m];
#pragma unroll
        for (int k4 = 0; k4 < BP / 4; ++k4) {
          R t[4], u[4];
          ld4(t, SA + col * BP + 4 * k4);
          ld4(u, SI + col * BP + 4 * k4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            t[i] = fma(Wm, u[i], fma(wm, ext[4 * k4 + i], t[i]));
          }
          st4(SA + col * BP + 4 * k4, t[0], t[1], t[2], t[3]);
        }
      }
    }
  }
  fence_async_smem();
  __syncthreads();
  if (tid < E && el < p.n) {
    tmaStoreCols<R>(static_cast<R*>(const_cast<void*>(p.dofs.at(el))), SQ, kNQ, BP, p.ldg);
    tmaStoreCols<R>(static_cast<R*>(const_cast<void*>(p.dofsAne.at(el))), SA, kVNA, BP, p.ldg);
    tma_store_wait_read();
  }
}

// ---------------------------------------------------------------------------------------
template <typename R, typename G, int E>
__global__ void __launch_bounds__((ViscoPlan<R, G, E>::threads), SSB_VISCO_MINB) k_visco_neigh(const ViscoNeighParams p) {
  using P = ViscoPlan<R, G, E>;
  constexpr int BP = P::BP, B2P = P::B2P, TQ = P::TQ, TF = P::TF;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  R* const smem = reinterpret_cast<R*>(smemRaw);
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + P::bytes(P::neighReals) - 16);
  const int tid = threadIdx.x;
  const bool active = tid < kVNE * E;
  const int e = active ? tid % E : 0;
  const int q = active ? tid / E : 0;
  R* const tile = smem + e * P::neighReals;
  R* const SF = tile + TQ;
  R* const SQ = SF + 2 * TF;   // own Q and Qane: arrive while the four faces are computed (bar[1])
  R* const SA = SQ + TQ;
  if (tid == 0) {
    mbar_init(bar, E);
    mbar_init(bar + 1, E);
    mbar_fence_init();
  }
  __syncthreads();
  const long long sl = static_cast<long long>(blockIdx.x) * E + e;
  const long long el = sl < p.n ? (p.perm != nullptr ? p.perm[sl] : sl) : -1;
  const bool valid = active && el >= 0;
  const R* ni = valid ? static_cast<const R*>(p.neighInt.at(el)) : nullptr;
  uchar4 ft = make_uchar4(255, 255, 255, 255);
  uint2 rr = make_uint2(0, 0);
  if (el >= 0) {
    ft = p.faceTypes != nullptr ? *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * el) : make_uchar4(0, 0, 0, 0);
    rr = *reinterpret_cast<const uint2*>(p.faceRel + 8 * el);
  }
  auto faceType = [&](int f) -> uint8_t { return f == 0 ? ft.x : (f == 1 ? ft.y : (f == 2 ? ft.z : ft.w)); };
  auto loadTile = [&](int f) {
    if (el < 0 || faceType(f) != 0) {
      mbar_arrive_expect_tx(bar, 0);
      return;
    }
    const R* src = p.nbPtrs != nullptr ? static_cast<const R*>(p.nbPtrs[4 * el + f])
                                       : static_cast<const R*>(p.idofs.at(p.nbIdx[4 * el + f]));
    mbar_arrive_expect_tx(bar, TQ * sizeof(R));
    tmaLoadCols<R>(tile, src, kNQ, BP, p.ldg, bar);
  };
  if (tid < E) {
    loadTile(0);
    if (el >= 0) {
      mbar_arrive_expect_tx(bar + 1, (TQ + P::TA) * sizeof(R));
      tmaLoadCols<R>(SQ, static_cast<const R*>(p.dofs.at(el)), kNQ, BP, p.ldg, bar + 1);
      tmaLoadCols<R>(SA, static_cast<const R*>(p.dofsAne.at(el)), kVNA, BP, p.ldg, bar + 1);
    } else {
      mbar_arrive_expect_tx(bar + 1, 0);
    }
  }
  const bool roleA = q < kNQ;
  const int pa = roleA ? 0 : q - kNQ;
  R ext[BP];
#pragma unroll
  for (int k = 0; k < BP; ++k) {
    ext[k] = R(0);
  }
#pragma unroll
  for (int f = 0; f < 4; ++f) {
    const bool regular = valid && faceType(f) == 0;
    R a[kNQ];
#pragma unroll
    for (int c = 0; c < kNQ; ++c) {
      a[c] = regular ? ni[kVA * f + kNQ * q + c] : R(0);
    }
    const unsigned w32 = f < 2 ? rr.x : rr.y;
    const int j = (w32 >> (16 * (f & 1))) & 0xff;
    const int h = (w32 >> (16 * (f & 1) + 8)) & 0xff;
    mbar_wait(bar, f & 1);
    R* const Fe = SF + (f & 1) * TF;
    if (roleA) {
      R Gt[B2P], F[B2P];
#pragma unroll
      for (int i = 0; i < B2P; ++i) {
        Gt[i] = R(0);
        F[i] = R(0);
      }
      if (regular) {
        G::template rT<R>(j, Gt, SrcPk<R>{tile + q * BP});
        G::template fP<R>(h, F, Gt);
      }
      if (active) {
#pragma unroll
        for (int k4 = 0; k4 < B2P / 4; ++k4) {
          st4(Fe + q * B2P + 4 * k4, F[4 * k4], F[4 * k4 + 1], F[4 * k4 + 2], F[4 * k4 + 3]);
        }
      }
    }
    __syncthreads();  // F visible; the tile is no longer read
    if (tid < E && f < 3) {
      loadTile(f + 1);
    }
    G::template rDivM<R>(f, ext, JChunks<R, B2P>{Fe, a});
  }
  // kernel::neighbor (viscoelastic2.py:317-334) on the staged own DOFs
  mbar_wait(bar + 1, 0);
  if (valid) {
    if (roleA) {
#pragma unroll
      for (int k4 = 0; k4 < BP / 4; ++k4) {
        R t[4];
        ld4(t, SQ + q * BP + 4 * k4);
        st4(SQ + q * BP + 4 * k4, t[0] + ext[4 * k4], t[1] + ext[4 * k4 + 1], t[2] + ext[4 * k4 + 2], t[3] + ext[4 * k4 + 3]);
      }
    } else {
#pragma unroll
      for (int m = 0; m < kVM; ++m) {
        const int col = pa + 6 * m;
        const R wm = ni[4 * kVA + m];
#pragma unroll
        for (int k4 = 0; k4 < BP / 4; ++k4) {
          R t[4];
          ld4(t, SA + col * BP + 4 * k4);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            t[i] = fma(wm, ext[4 * k4 + i], t[i]);
          }
          st4(SA + col * BP + 4 * k4, t[0], t[1], t[2], t[3]);
        }
      }
    }
  }
  fence_async_smem();
  __syncthreads();
  if (tid < E && el >= 0) {
    tmaStoreCols<R>(static_cast<R*>(const_cast<void*>(p.dofs.at(el))), SQ, kNQ, BP, p.ldg);
    tmaStoreCols<R>(static_cast<R*>(const_cast<void*>(p.dofsAne.at(el))), SA, kVNA, BP, p.ldg);
    tma_store_wait_read();
  }
}

// ---------------------------------------------------------------------------------------
// k_visco_neigh_dr: the DynamicRupture branch of the viscoelastic neighbor integral
// (src/Kernels/LinearCKAnelastic/Neighbor.cpp:107-115: dynamicRupture::kernel::nodalFlux into Qext, then kernel::neighbor):
//   Qext = sum over the element's fault faces of  V3mTo2nTWDivM(side, rel) . godunov . fluxSolver      [B x 15]
//   Q += Qext[:, :9];  Qane[:, p, m] += w[m] Qext[:, 9 + p]
// kernel::neighbor is linear in Qext, so regular faces (k_visco_neigh) and fault faces are two launches.
// One warp per element like k_neigh_dr (kernels_dr.cuh): lanes are the fault points for T = godunov . fluxSolver
// [P x 15] (shared memory), then rows k = lane, lane + 32 of W(variant) . T from the dense table of the operator.
template <typename R, typename G, typename GD, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k_visco_neigh_dr(const ViscoNeighDrParams p) {
  constexpr int B = G::kB, P = GD::kP;
  constexpr int TS = 16;  // reals per point of T in shared memory (15 padded to 16: four 128-bit words)
  __shared__ __align__(16) R sT[WARPS][2 * 32 * TS];
  __shared__ __align__(16) R sFs[WARPS][136];
  static_assert(P <= 64 && B <= 64, "two points / rows per lane");
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long el = static_cast<long long>(blockIdx.x) * WARPS + warp;
  if (el >= p.n) {
    return;
  }
  const R* const table = GD::template nodalFluxTable<R>();
  const uchar4 ft = *reinterpret_cast<const uchar4*>(p.faceTypes + 4 * el);
  R acc[2][kVNE];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
#pragma unroll
    for (int q = 0; q < kVNE; ++q) {
      acc[r][q] = R(0);
    }
  }
  bool any = false;
#pragma unroll 1
  for (int f = 0; f < 4; ++f) {
    const uint8_t ty = f == 0 ? ft.x : (f == 1 ? ft.y : (f == 2 ? ft.z : ft.w));
    const R* god = static_cast<const R*>(p.godunov[4 * el + f]);
    if (ty != 3 || god == nullptr) {
      continue;  // warp-uniform
    }
    any = true;
    const R* fs = static_cast<const R*>(p.fluxSolver[4 * el + f]);
    const int variant = p.sideRel[8 * el + 2 * f] + 4 * p.sideRel[8 * el + 2 * f + 1];
    __syncwarp();
    for (int i = lane; i < 9 * kVNE; i += 32) {
      sFs[warp][i] = fs[i];
    }
    __syncwarp();
    // ---- step 1: T = godunov . fluxSolver for the points lane, lane + 32
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int pt = lane + 32 * half;
      R t[TS];
#pragma unroll
      for (int q = 0; q < TS; ++q) {
        t[q] = R(0);
      }
      if (pt < P) {
#pragma unroll
        for (int c = 0; c < 9; ++c) {
          const R gv = god[c * p.ppad + pt];
#pragma unroll
          for (int q = 0; q < kVNE; ++q) {
            t[q] = fma(gv, sFs[warp][c + 9 * q], t[q]);
          }
        }
      }
      R* dst = &sT[warp][pt * TS];
#pragma unroll
      for (int q4 = 0; q4 < TS / 4; ++q4) {
        st4(dst + 4 * q4, t[4 * q4], t[4 * q4 + 1], t[4 * q4 + 2], t[4 * q4 + 3]);
      }
    }
    __syncwarp();
    // ---- step 2: acc += W(variant) . T
    const R* w = table + static_cast<size_t>(variant) * P * 64 + lane;
#pragma unroll 4
    for (int pt = 0; pt < P; ++pt) {
      const R w0 = w[pt * 64];
      const R w1 = w[pt * 64 + 32];
      R t[TS];
#pragma unroll
      for (int q4 = 0; q4 < TS / 4; ++q4) {
        R u[4];
        ld4(u, &sT[warp][pt * TS + 4 * q4]);
        t[4 * q4] = u[0]; t[4 * q4 + 1] = u[1]; t[4 * q4 + 2] = u[2]; t[4 * q4 + 3] = u[3];
      }
#pragma unroll
      for (int q = 0; q < kVNE; ++q) {
        acc[0][q] = fma(w0, t[q], acc[0][q]);
        acc[1][q] = fma(w1, t[q], acc[1][q]);
      }
    }
  }
  if (!any) {
    return;
  }
  R* const dofs = static_cast<R*>(const_cast<void*>(p.dofs.at(el)));
  R* const ane = static_cast<R*>(const_cast<void*>(p.dofsAne.at(el)));
  const R* const wv = static_cast<const R*>(p.neighInt.at(el)) + 4 * kVA;
  R wm[kVM];
#pragma unroll
  for (int m = 0; m < kVM; ++m) {
    wm[m] = wv[m];
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int k = lane + 32 * r;
    if (k < B) {
#pragma unroll
      for (int q = 0; q < kNQ; ++q) {
        dofs[q * p.ldg + k] += acc[r][q];
      }
#pragma unroll
      for (int m = 0; m < kVM; ++m) {
#pragma unroll
        for (int a = 0; a < 6; ++a) {
          ane[(a + 6 * m) * p.ldg + k] += wm[m] * acc[r][kNQ + a];
        }
      }
    }
  }
}

}  // namespace ssb
