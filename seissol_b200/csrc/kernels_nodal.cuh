// Nodal boundary conditions of the local integral: Dirichlet (easi boundary map), FreeSurfaceGravity and Analytical faces.
//
// Reference: Local::computeIntegral, src/Kernels/LinearCK/Local.cpp:157-252 (batched :350-455 through
// DirichletBoundary::evaluateOnDevice, src/Kernels/LinearCK/DirichletBoundary.h:67-118), generated kernels
// projectToNodalBoundaryRotated, createEasiBoundaryGhostCells (codegen/kernels/nodalbc.py:55-85) and localFluxNodal
// (codegen/kernels/aderdg.py:336-352).  Per boundary face (cell c, local face f):
//   N      = V3mTo2nFace_f . I_c                  time-integrated DOFs at the B2 face nodes        [B2 x 9]
//   INodal = N . Tinv^T                           rotated into the face-aligned frame
//   Dirichlet (FaceType 4):          INodal(l, a) = sum_b map(a, b, l) INodal(l, b) + const(a, l)
//   FreeSurfaceGravity (FaceType 2): INodal(l, 0..2) = 2 p - INodal(l, 0..2),  p = -rho g displacement(l)
//   Analytical (FaceType 7):         INodal given by the caller (host callback, Local.cpp:409-452)
//   Q_c   += project2nFaceTo3m_f . (INodal . AminusT_f)                                             [B x 9]
// One warp per boundary face: lanes are the face nodes for the first three steps and the basis rows (lane, lane + 32)
// for the last; the two dense operators come from lane-fastest tables (gen_nodal.py).  Boundary faces are a surface
// set (few per cent of the cells), so this kernel is far from the roofline of the volume kernels by design.  Two
// boundary faces of ONE cell may run concurrently: Q is updated with atomic adds.
#pragma once
#include "common.cuh"
#include "params.cuh"

namespace ssb {

struct NodalBcParams {
  ElemArray dofs, idofs, neighInt;
  const long long* cell;      // [m]
  const uint8_t* face;        // [m]
  const uint8_t* kind;        // [m] FaceType: 4 = Dirichlet, 2 = FreeSurfaceGravity (others: face skipped); NULL: every
                              // face carries GIVEN nodal values (FaceType::Analytical, `analytical`)
  const void* tinv;           // [m][9][9] row-major Tinv(p, q)
  const void* easiMap;        // [m][9][9][B2] (a, b, l), Dirichlet faces
  const void* easiConst;      // [m][9][B2]
  const void* displacement;   // [m][B2], FreeSurfaceGravity faces
  const void* rhoG;           // [m] rho * g
  const void* analytical;     // [m][9][B2] (a, l): time-integrated boundary values at the face nodes (Analytical faces)
  long long m;
  int ldg;
};

template <typename R, typename G, typename GN, int WARPS>
__global__ void __launch_bounds__(32 * WARPS) k_local_nodal_bc(const NodalBcParams p) {
  constexpr int B = G::kB, B2 = GN::kB2;
  __shared__ R sI[WARPS][9 * B];        // I of the cell, [q][k]
  __shared__ R sT[WARPS][81 + 3];       // Tinv, then AminusT
  __shared__ R sZ[WARPS][32 * 9];       // node values [l][9]
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const long long i = static_cast<long long>(blockIdx.x) * WARPS + warp;
  if (i >= p.m) {
    return;
  }
  constexpr int kAnalytical = 7;   // FaceType::Analytical
  const int kind = p.kind != nullptr ? p.kind[i] : kAnalytical;
  if (kind != 4 && kind != 2 && !(kind == kAnalytical && p.analytical != nullptr)) {
    return;   // warp-uniform
  }
  const long long c = p.cell[i];
  const int f = p.face[i];
  R in[9];
  if (kind == kAnalytical) {
    // the boundary values at the face nodes are GIVEN (Local.cpp:409-452: the host evaluates the analytical solution at
    // the nodes and integrates it over the time step, DirichletBoundary::evaluateTimeDependent; the device applies
    // localFluxNodal to that scratch array)
    const R* const given = static_cast<const R*>(p.analytical) + static_cast<size_t>(i) * 9 * B2 + lane;
#pragma unroll
    for (int a = 0; a < 9; ++a) {
      in[a] = lane < B2 ? given[a * B2] : R(0);
    }
  } else {
  const R* const Ic = static_cast<const R*>(p.idofs.at(c));
  for (int q = 0; q < 9; ++q) {
    for (int k = lane; k < B; k += 32) {
      sI[warp][q * B + k] = Ic[q * p.ldg + k];
    }
  }
  const R* const tinv = static_cast<const R*>(p.tinv) + 81 * i;
  for (int e = lane; e < 81; e += 32) {
    sT[warp][e] = tinv[e];
  }
  __syncwarp();
  // ---- N(l, q) = sum_k V(l, k) I(k, q), lane = node l
  R nq[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) {
    nq[q] = R(0);
  }
  const R* V = GN::template nodalV<R>() + static_cast<size_t>(f) * B * 32 + lane;
#pragma unroll 4
  for (int k = 0; k < B; ++k) {
    const R v = V[k * 32];
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      nq[q] = fma(v, sI[warp][q * B + k], nq[q]);
    }
  }
  // ---- INodal(l, p') = sum_q N(l, q) Tinv(p', q)
#pragma unroll
  for (int pp = 0; pp < 9; ++pp) {
    R a = R(0);
#pragma unroll
    for (int q = 0; q < 9; ++q) {
      a = fma(nq[q], sT[warp][9 * pp + q], a);
    }
    in[pp] = a;
  }
  }
  // ---- the boundary condition at the node
  if (lane < B2 && kind != kAnalytical) {
    if (kind == 4) {
      const R* map = static_cast<const R*>(p.easiMap) + static_cast<size_t>(i) * 81 * B2 + lane;
      const R* cst = static_cast<const R*>(p.easiConst) + static_cast<size_t>(i) * 9 * B2 + lane;
      R out[9];
#pragma unroll
      for (int a = 0; a < 9; ++a) {
        R s = cst[a * B2];
#pragma unroll
        for (int b = 0; b < 9; ++b) {
          s = fma(map[(a * 9 + b) * B2], in[b], s);
        }
        out[a] = s;
      }
#pragma unroll
      for (int a = 0; a < 9; ++a) {
        in[a] = out[a];
      }
    } else {
      const R pressure = -static_cast<const R*>(p.rhoG)[i] * static_cast<const R*>(p.displacement)[i * B2 + lane];
      in[0] = R(2) * pressure - in[0];
      in[1] = R(2) * pressure - in[1];
      in[2] = R(2) * pressure - in[2];
    }
  }
  __syncwarp();   // everybody is done with Tinv
  const R* const A = static_cast<const R*>(p.neighInt.at(c)) + 81 * f;
  for (int e = lane; e < 81; e += 32) {
    sT[warp][e] = A[e];   // AminusT, memory A[o + 9 p']
  }
  __syncwarp();
  // ---- Z(l, p') = sum_o INodal(l, o) AminusT(o, p')
#pragma unroll
  for (int pp = 0; pp < 9; ++pp) {
    R a = R(0);
#pragma unroll
    for (int o = 0; o < 9; ++o) {
      a = fma(in[o], sT[warp][o + 9 * pp], a);
    }
    sZ[warp][lane * 9 + pp] = lane < B2 ? a : R(0);
  }
  __syncwarp();
  // ---- Q(k, p') += sum_l P(k, l) Z(l, p'), lane = rows k = lane, lane + 32
  R acc[2][9];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
#pragma unroll
    for (int pp = 0; pp < 9; ++pp) {
      acc[r][pp] = R(0);
    }
  }
  const R* P = GN::template nodalP<R>() + static_cast<size_t>(f) * B2 * 64 + lane;
#pragma unroll 3
  for (int l = 0; l < B2; ++l) {
    const R p0 = P[l * 64], p1 = P[l * 64 + 32];
#pragma unroll
    for (int pp = 0; pp < 9; ++pp) {
      const R zv = sZ[warp][l * 9 + pp];
      acc[0][pp] = fma(p0, zv, acc[0][pp]);
      acc[1][pp] = fma(p1, zv, acc[1][pp]);
    }
  }
  R* const Q = static_cast<R*>(const_cast<void*>(p.dofs.at(c)));
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int k = lane + 32 * r;
    if (k < B) {
#pragma unroll
      for (int pp = 0; pp < 9; ++pp) {
        atomicAdd(Q + pp * p.ldg + k, acc[r][pp]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Volume energies of the energy output: EnergyOutput::computeVolumeEnergies (src/ResultWriter/EnergyOutput.cpp:347-486).
// The reference evaluates Q at (O+1)^3 quadrature points of every element (kernel::evalAtQP) and sums
//   kinetic 0.5 rho |v|^2,  elastic 0.5 sigma : eps(sigma),  acoustic p^2 / (2 K) (elements with mu = 0),  momentum rho v
// with weights 6 |V| w_qp.  All integrands are products of two expansions in the ORTHOGONAL modal basis, and the quadrature
// is exact for them (degree 2 (O - 1) <= 2 (O + 1) - 1), so the same integrals are
//   int f g = 6 |V| sum_k M_kk f_k g_k,        int f = 6 |V| M_00 f_0   (phi_0 = 1)
// evaluated here straight from the modal coefficients: one pass over Q (HBM-bound), 16 (float) / 32 (double) lanes per
// element with one 128-bit load per lane and quantity, sums in double.
// out[8] (accumulated with atomic adds, the caller zeroes it): elastic kinetic, elastic, acoustic kinetic, acoustic,
// momentum x / y / z, out[7] unused (the gravitational energy needs the face displacements of the boundary faces).
struct VolumeEnergyParams {
  ElemArray dofs;
  const double* volume;     // [n]
  const double* material;   // [n][3] lambda, mu, rho
  double* out;              // [8]
  long long n;
  int ldg;
};

template <typename R, typename G, typename GN>
__global__ void __launch_bounds__(256) k_volume_energies(const VolumeEnergyParams p) {
  constexpr int B = G::kB;
  constexpr int W = 16 / int(sizeof(R));            // reals per 128-bit load
  constexpr int L = sizeof(R) == 4 ? 16 : 32;       // lanes per element: L * W rows >= B (B <= 56)
  constexpr int EPW = 32 / L;                       // elements per warp and iteration
  static_assert(L * W >= B, "one 128-bit load per lane and quantity covers a column");
  __shared__ double sAcc[8][8];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sub = lane / L, sl = lane % L;
  const bool rows = sl * W < B;
  double mass[W];
#pragma unroll
  for (int i = 0; i < W; ++i) {
    mass[i] = sl * W + i < B ? GN::massDiag()[sl * W + i] : 0.0;
  }
  const double mass0 = GN::massDiag()[0];
  double tot[7] = {0, 0, 0, 0, 0, 0, 0};
  const long long stride = static_cast<long long>(gridDim.x) * 8 * EPW;
  for (long long base = (static_cast<long long>(blockIdx.x) * 8 + warp) * EPW; base < p.n; base += stride) {
    const long long el = base + sub;
    const bool valid = el < p.n;
    const R* const Q = valid ? static_cast<const R*>(p.dofs.at(el)) : nullptr;
    R q[9][W];
#pragma unroll
    for (int a = 0; a < 9; ++a) {
      if (valid && rows) {
        if constexpr (sizeof(R) == 4) {
          const float4 t = *reinterpret_cast<const float4*>(Q + a * p.ldg + sl * W);
          q[a][0] = t.x; q[a][1] = t.y; q[a][2] = t.z; q[a][3] = t.w;
        } else {
          const double2 t = *reinterpret_cast<const double2*>(Q + a * p.ldg + sl * W);
          q[a][0] = t.x; q[a][1] = t.y;
        }
      } else {
#pragma unroll
        for (int i = 0; i < W; ++i) {
          q[a][i] = R(0);
        }
      }
    }
    double kin = 0, sq = 0, tr = 0, pp = 0;
#pragma unroll
    for (int i = 0; i < W; ++i) {
      // the point values and their products are `real` in the reference too (EnergyOutput.cpp:435-438, :463-481); the
      // sums are double.  In single precision this keeps the kernel on the HBM side of the conversion / DFMA rates.
      const double m = mass[i];
      const R xx = q[0][i], yy = q[1][i], zz = q[2][i], xy = q[3][i], yz = q[4][i], xz = q[5][i];
      const R u = q[6][i], v = q[7][i], w = q[8][i];
      const R trace = xx + yy + zz;
      kin += m * static_cast<double>(u * u + v * v + w * w);
      sq += m * static_cast<double>(xx * xx + yy * yy + zz * zz + R(2) * (xy * xy + yz * yz + xz * xz));
      tr += m * static_cast<double>(trace * trace);
      pp += m * static_cast<double>(xx * xx);
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) {   // stays inside the L-lane group of the element
      kin += __shfl_xor_sync(0xffffffffu, kin, o);
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
      tr += __shfl_xor_sync(0xffffffffu, tr, o);
      pp += __shfl_xor_sync(0xffffffffu, pp, o);
    }
    if (sl == 0 && valid) {
      const double jac = 6.0 * p.volume[el];
      const double lambda = p.material[3 * el], mu = p.material[3 * el + 1], rho = p.material[3 * el + 2];
      if (fabs(mu) < 10e-14) {   // acoustic element (EnergyOutput.cpp:443-450)
        tot[2] += jac * 0.5 * rho * kin;
        tot[3] += jac * pp / (2.0 * lambda);
      } else {
        tot[0] += jac * 0.5 * rho * kin;
        const double factor = -lambda / (2.0 * mu * (3.0 * lambda + 2.0 * mu));
        tot[1] += jac * 0.5 * (sq / (2.0 * mu) + factor * tr);
        const double m0 = jac * mass0 * rho;   // row 0 of the columns 6..8 is element 0 of this lane's loads
        tot[4] += m0 * static_cast<double>(q[6][0]);
        tot[5] += m0 * static_cast<double>(q[7][0]);
        tot[6] += m0 * static_cast<double>(q[8][0]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 7; ++i) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tot[i] += __shfl_xor_sync(0xffffffffu, tot[i], o);
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < 7; ++i) {
      sAcc[warp][i] = tot[i];
    }
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    double s = 0;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) {
      s += sAcc[wv][threadIdx.x];
    }
    atomicAdd(p.out + threadIdx.x, s);
  }
}

}  // namespace ssb
