// Cell-local and fault-local matrix set-up ON THE DEVICE: what the element-update kernels consume
// per element -- the star matrices and the flux solvers AplusT / AminusT of LocalIntegrationData /
// NeighboringIntegrationData -- and per fault face -- ImpedancesAndEta, TinvT, fluxSolverPlus / Minus -- from vertex
// coordinates and material parameters.
//
// Reference (host code there): src/Initializer/Model/CellLocalMatrices.cpp:66-298 (star = sum_d grad[d] M_d^T; AplusT =
// fluxScale Tinv^T (QgodLocal A~^T) T^T, fluxScale = -2 |S| / (6 |V|)), src/Equations/elastic/Model/ElasticSetup.h:28-215
// (transposed coefficient matrices, transposed Godunov state from the eigenvector basis, face rotation),
// src/Model/Common.h:319-364 (free surface), src/Numerical/Transformation.cpp:66-282, src/Geometry/MeshTools.cpp:44-132,
// src/Initializer/Model/DynamicRuptureMatrices.cpp:111-477 (impedances :341-359, TinvT :427, fluxSolver :461-469).
// One thread per (element, face) / per fault face, all arithmetic in double precision (set-up runs once; its results
// are stored in the precision of the context).  Quantity order: sxx syy szz sxy syz sxz u v w.
#pragma once
#include "common.cuh"
#include "params.cuh"

namespace ssb {

struct SetupCellParams {
  const double* coords;        // [n][4][3] vertex coordinates
  const long long* neighbors;  // [n][4] index into `material` of the face neighbor, < 0: none (own material)
  const double* material;      // [*][3] lambda, mu, rho; rows 0..n-1 are the own cells
  const uint8_t* faceTypes;    // [n][4] or null
  ElemArray localInt;          // out: star[3][24] | nApNm1[4][81]
  ElemArray neighInt;          // out: nAmNm1[4][81]
  long long n;
};

struct SetupDrParams {
  const double* coordsPlus;    // [m][4][3] vertices of the plus-side element
  const double* coordsMinus;   // [m][4][3]
  const uint8_t* faceInfo;     // [m][3] {plusSide, minusSide, faceRelation}
  const double* materialPlus;  // [m][3] lambda, mu, rho
  const double* materialMinus;
  void* tinvT;                 // out [m][81]  DRGodunovData::dataTinvT, column-major TinvT[a + 9 b]
  void* impAndEta;             // out [m][11]  zp zs zpNeig zsNeig etaP etaS invEtaS invZp invZs invZpNeig invZsNeig
  void* fluxSolverPlus;        // out [m][81]  column-major [c + 9 q]
  void* fluxSolverMinus;
  void* doubledSurfaceArea;    // out [m] or null
  long long m;
};

// Fault part of the energy output: EnergyOutput::computeDynamicRuptureEnergies (src/ResultWriter/EnergyOutput.cpp:262-345)
struct FaultEnergyParams {
  const void* frictionalEnergy;     // [m][ppad] DREnergyOutput::frictionalEnergy
  const void* accumulatedSlip;      // [m][ppad] DREnergyOutput::accumulatedSlip
  const void* doubledSurfaceArea;   // [m]
  const double* materialPlus;       // [m][3] lambda, mu, rho
  const double* materialMinus;
  const uint8_t* plusSideOnThisRank;  // [m] or null (= all faces)
  double* out;                      // [4] accumulated: frictional work, potency, seismic moment, unused
  long long m;
  int points, ppad;
};

// one warp per fault face, lanes over the points; sums in double
template <typename R>
__global__ void __launch_bounds__(256) k_fault_energies(const FaultEnergyParams p) {
  __shared__ double sAcc[8][3];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  double work = 0, potency = 0, moment = 0;
  const long long stride = static_cast<long long>(gridDim.x) * 8;
  for (long long i = static_cast<long long>(blockIdx.x) * 8 + warp; i < p.m; i += stride) {
    if (p.plusSideOnThisRank != nullptr && p.plusSideOnThisRank[i] == 0) {
      continue;   // warp-uniform
    }
    const R* const fe = static_cast<const R*>(p.frictionalEnergy) + i * p.ppad;
    const R* const as = static_cast<const R*>(p.accumulatedSlip) + i * p.ppad;
    double w = 0, s = 0;
    for (int k = lane; k < p.points; k += 32) {
      w += static_cast<double>(fe[k]);
      s += static_cast<double>(as[k]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      w += __shfl_xor_sync(0xffffffffu, w, o);
      s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    if (lane == 0) {
      const double muPlus = p.materialPlus[3 * i + 1], muMinus = p.materialMinus[3 * i + 1];
      const double mu = 2.0 * muPlus * muMinus / (muPlus + muMinus);
      const double inc = s * 0.5 * static_cast<double>(static_cast<const R*>(p.doubledSurfaceArea)[i]) / p.points;
      work += w;
      potency += inc;
      moment += inc * mu;
    }
  }
  if (lane == 0) {
    sAcc[warp][0] = work; sAcc[warp][1] = potency; sAcc[warp][2] = moment;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) {
      t += sAcc[wv][threadIdx.x];
    }
    atomicAdd(p.out + threadIdx.x, t);
  }
}

namespace setup {

struct M9 {
  double a[9][9];
  __device__ void zero() {
    for (int i = 0; i < 9; ++i) {
      for (int j = 0; j < 9; ++j) {
        a[i][j] = 0.0;
      }
    }
  }
};
__device__ inline void mul(const M9& x, const M9& y, M9& out) {
  for (int i = 0; i < 9; ++i) {
    for (int j = 0; j < 9; ++j) {
      double s = 0.0;
      for (int k = 0; k < 9; ++k) {
        s += x.a[i][k] * y.a[k][j];
      }
      out.a[i][j] = s;
    }
  }
}
__device__ inline void transpose(const M9& x, M9& out) {
  for (int i = 0; i < 9; ++i) {
    for (int j = 0; j < 9; ++j) {
      out.a[i][j] = x.a[j][i];
    }
  }
}
// getTransposedCoefficientMatrix (ElasticSetup.h:28-78)
__device__ inline void coefficientMatrixT(double lam, double mu, double rho, int dim, M9& m) {
  m.zero();
  const double l2m = lam + 2.0 * mu, ri = 1.0 / rho;
  if (dim == 0) {
    m.a[6][0] = -l2m; m.a[6][1] = -lam; m.a[6][2] = -lam; m.a[7][3] = -mu; m.a[8][5] = -mu;
    m.a[0][6] = -ri; m.a[3][7] = -ri; m.a[5][8] = -ri;
  } else if (dim == 1) {
    m.a[7][0] = -lam; m.a[7][1] = -l2m; m.a[7][2] = -lam; m.a[6][3] = -mu; m.a[8][4] = -mu;
    m.a[1][7] = -ri; m.a[3][6] = -ri; m.a[4][8] = -ri;
  } else {
    m.a[8][0] = -lam; m.a[8][1] = -lam; m.a[8][2] = -l2m; m.a[7][4] = -mu; m.a[6][5] = -mu;
    m.a[2][8] = -ri; m.a[5][6] = -ri; m.a[4][7] = -ri;
  }
}
// solve X A = B for X (A, B 9 x 9): Gaussian elimination with partial pivoting on A^T X^T = B^T
__device__ inline void solveRight(const M9& A, const M9& B, M9& X) {
  double m[9][18];
  for (int i = 0; i < 9; ++i) {
    for (int j = 0; j < 9; ++j) {
      m[i][j] = A.a[j][i];
      m[i][9 + j] = B.a[j][i];
    }
  }
  for (int c = 0; c < 9; ++c) {
    int piv = c;
    for (int r = c + 1; r < 9; ++r) {
      if (fabs(m[r][c]) > fabs(m[piv][c])) {
        piv = r;
      }
    }
    if (piv != c) {
      for (int j = 0; j < 18; ++j) {
        const double t = m[c][j]; m[c][j] = m[piv][j]; m[piv][j] = t;
      }
    }
    const double inv = 1.0 / m[c][c];
    for (int r = 0; r < 9; ++r) {
      if (r != c) {
        const double f = m[r][c] * inv;
        for (int j = c; j < 18; ++j) {
          m[r][j] -= f * m[c][j];
        }
      }
    }
    for (int j = c; j < 18; ++j) {
      m[c][j] *= inv;
    }
  }
  for (int i = 0; i < 9; ++i) {
    for (int j = 0; j < 9; ++j) {
      X.a[j][i] = m[i][9 + j];
    }
  }
}
// getTransposedGodunovState (ElasticSetup.h:80-180), elastic-elastic: QgodLocal = 1 - god, QgodNeighbor = god,
// god = (R^T)^-1 chi R^T, i.e. god^T... evaluated as: solve R^T god = chi R^T
__device__ inline void godunovStateT(double lam, double mu, double rho, double lamN, double muN, double rhoN, M9& qLoc,
                                     M9& qNb) {
  M9 R;
  R.zero();
  R.a[0][0] = lam + 2 * mu; R.a[1][0] = lam; R.a[2][0] = lam; R.a[6][0] = sqrt((lam + 2 * mu) / rho);
  R.a[3][1] = mu; R.a[7][1] = sqrt(mu / rho);
  R.a[5][2] = mu; R.a[8][2] = sqrt(mu / rho);
  R.a[4][3] = R.a[1][4] = R.a[2][5] = lam + 2 * mu;
  R.a[5][6] = muN; R.a[8][6] = -sqrt(muN / rhoN);
  R.a[3][7] = muN; R.a[7][7] = -sqrt(muN / rhoN);
  R.a[0][8] = lamN + 2 * muN; R.a[1][8] = lamN; R.a[2][8] = lamN; R.a[6][8] = -sqrt((lamN + 2 * muN) / rhoN);
  // god = solve(R^T, chi R^T): (R^T) god = chi R^T  <=>  god^T R = R chi^T  (chi diagonal)  -> X A = B with X = god^T
  M9 B;
  for (int i = 0; i < 9; ++i) {
    for (int j = 0; j < 9; ++j) {
      B.a[i][j] = j < 3 ? R.a[i][j] : 0.0;   // R chi (chi = diag(1, 1, 1, 0, ...))
    }
  }
  M9 godT;
  solveRight(R, B, godT);
  for (int i = 0; i < 9; ++i) {
    for (int j = 0; j < 9; ++j) {
      qNb.a[i][j] = godT.a[j][i];
      qLoc.a[i][j] = (i == j ? 1.0 : 0.0) - godT.a[j][i];
    }
  }
}
// getTransposedFreeSurfaceGodunovState (src/Model/Common.h:319-364), elastic
__device__ inline void freeSurfaceGodunovStateT(double lam, double mu, double rho, M9& q) {
  q.zero();
  const double cp = sqrt((lam + 2 * mu) / rho), cs = sqrt(mu / rho);
  const double S[3] = {-cp / (lam + 2 * mu), -cs / mu, -cs / mu};
  const int tr[3] = {0, 3, 5};
  for (int i = 0; i < 3; ++i) {
    q.a[tr[i]][6 + i] = S[i];
    q.a[6 + i][6 + i] = 1.0;
  }
}
// getFaceRotationMatrix (ElasticSetup.h:182-198)
__device__ inline void rotationMatrices(const double (&v)[3][3], M9& T, M9& Ti) {
  T.zero();
  Ti.zero();
  const int pa[6] = {0, 1, 2, 0, 1, 0}, pb[6] = {0, 1, 2, 1, 2, 2};   // Voigt order xx yy zz xy yz xz
  for (int col = 0; col < 6; ++col) {
    const int a = pa[col], b = pb[col];
    for (int row = 0; row < 6; ++row) {
      const int i = pa[row], j = pb[row];
      T.a[row][col] = a == b ? v[a][i] * v[b][j] : v[a][i] * v[b][j] + v[b][i] * v[a][j];
    }
  }
  for (int a = 0; a < 3; ++a) {
    for (int i = 0; i < 3; ++i) {
      T.a[6 + i][6 + a] = v[a][i];
      Ti.a[6 + a][6 + i] = v[a][i];
    }
  }
  for (int row = 0; row < 6; ++row) {
    const int a = pa[row], b = pb[row];
    for (int col = 0; col < 6; ++col) {
      const int i = pa[col], j = pb[col];
      Ti.a[row][col] = i == j ? v[a][i] * v[b][j] : v[a][i] * v[b][j] + v[a][j] * v[b][i];
    }
  }
}
__device__ inline void cross(const double (&a)[3], const double (&b)[3], double (&c)[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ inline double norm3(const double (&a)[3]) { return sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]); }

// face f of a tetrahedron (reference face-to-vertex table, MeshTools / FACE2NODES): unit normal, tangents, area; volume
__device__ inline void faceGeometry(const double* c, int f, double (&frame)[3][3], double& surface, double& volume) {
  const int f2n[4][3] = {{0, 2, 1}, {0, 1, 3}, {0, 3, 2}, {1, 2, 3}};
  double e1[3], e2[3], e3[3];
  for (int d = 0; d < 3; ++d) {
    e1[d] = c[3 + d] - c[d];
    e2[d] = c[6 + d] - c[d];
    e3[d] = c[9 + d] - c[d];
  }
  double cr[3];
  cross(e2, e3, cr);
  volume = fabs(e1[0] * cr[0] + e1[1] * cr[1] + e1[2] * cr[2]) / 6.0;
  const double* a = c + 3 * f2n[f][0];
  const double* b = c + 3 * f2n[f][1];
  const double* cc = c + 3 * f2n[f][2];
  double t1[3], u[3], nrm[3], t2[3];
  for (int d = 0; d < 3; ++d) {
    t1[d] = b[d] - a[d];
    u[d] = cc[d] - a[d];
  }
  cross(t1, u, nrm);
  const double nl = norm3(nrm);
  surface = 0.5 * nl;
  cross(nrm, t1, t2);
  const double l1 = norm3(t1), l2 = norm3(t2);
  for (int d = 0; d < 3; ++d) {
    frame[0][d] = nrm[d] / nl;
    frame[1][d] = t1[d] / l1;
    frame[2][d] = t2[d] / l2;
  }
}

}  // namespace setup

// CSC pattern of the 24-entry star matrix (row, column), column by column (codegen/matrices/star.xml; kStarRow / kStarCnt)
template <typename R>
__global__ void __launch_bounds__(128) k_setup_cell(const SetupCellParams p) {
  using namespace setup;
  const long long t = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long e = t >> 2;
  const int f = static_cast<int>(t & 3);
  if (e >= p.n) {
    return;
  }
  const double* c = p.coords + 12 * e;
  const double lam = p.material[3 * e], mu = p.material[3 * e + 1], rho = p.material[3 * e + 2];
  R* const li = static_cast<R*>(const_cast<void*>(p.localInt.at(e)));
  R* const ni = static_cast<R*>(const_cast<void*>(p.neighInt.at(e)));
  if (f == 0) {
    // star_i = sum_d grad[i][d] M_d^T, grad = inverse of the edge matrix (tetrahedronGlobalToReferenceJacobian)
    double m[3][3];
    for (int r = 0; r < 3; ++r) {
      for (int d = 0; d < 3; ++d) {
        m[d][r] = c[3 * (r + 1) + d] - c[d];   // columns = edges
      }
    }
    const double det = m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
                       m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
    double g[3][3];
    g[0][0] = (m[1][1] * m[2][2] - m[1][2] * m[2][1]) / det; g[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) / det;
    g[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) / det; g[1][0] = (m[1][2] * m[2][0] - m[1][0] * m[2][2]) / det;
    g[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) / det; g[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) / det;
    g[2][0] = (m[1][0] * m[2][1] - m[1][1] * m[2][0]) / det; g[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) / det;
    g[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) / det;
    M9 A0, A1, A2;
    coefficientMatrixT(lam, mu, rho, 0, A0);
    coefficientMatrixT(lam, mu, rho, 1, A1);
    coefficientMatrixT(lam, mu, rho, 2, A2);
    for (int i = 0; i < 3; ++i) {
      int k = 0;
      for (int col = 0; col < 9; ++col) {
        for (int r = 0; r < kStarCnt[col]; ++r) {
          const int row = kStarRow[col][r];
          li[24 * i + k] = static_cast<R>(g[i][0] * A0.a[row][col] + g[i][1] * A1.a[row][col] + g[i][2] * A2.a[row][col]);
          ++k;
        }
      }
    }
  }
  double frame[3][3], surface, volume;
  faceGeometry(c, f, frame, surface, volume);
  const long long nb = p.neighbors[4 * e + f];
  const long long mrow = nb >= 0 ? nb : e;
  const double lamN = p.material[3 * mrow], muN = p.material[3 * mrow + 1], rhoN = p.material[3 * mrow + 2];
  const uint8_t ft = p.faceTypes != nullptr ? p.faceTypes[4 * e + f] : 0;
  M9 qLoc, qNb, T, Ti, At, tmp, tmp2, TiT, TT;
  godunovStateT(lam, mu, rho, lamN, muN, rhoN, qLoc, qNb);
  if (ft == 1) {   // FaceType::FreeSurface
    freeSurfaceGodunovStateT(lam, mu, rho, qLoc);
    qNb.zero();
  }
  rotationMatrices(frame, T, Ti);
  transpose(Ti, TiT);
  transpose(T, TT);
  coefficientMatrixT(lam, mu, rho, 0, At);
  const double scale = -2.0 * surface / (6.0 * volume);
  // AplusT = scale Tinv^T (QgodLocal A~^T) T^T ; memory A[i + 9 j] = M(i, j)
  mul(qLoc, At, tmp);
  mul(TiT, tmp, tmp2);
  mul(tmp2, TT, tmp);
  for (int i = 0; i < 9; ++i) {
    for (int j = 0; j < 9; ++j) {
      li[72 + 81 * f + i + 9 * j] = ft == 3 ? R(0) : static_cast<R>(scale * tmp.a[i][j]);   // DynamicRupture: no local flux
    }
  }
  mul(qNb, At, tmp);
  mul(TiT, tmp, tmp2);
  mul(tmp2, TT, tmp);
  for (int i = 0; i < 9; ++i) {
    for (int j = 0; j < 9; ++j) {
      ni[81 * f + i + 9 * j] = static_cast<R>(scale * tmp.a[i][j]);
    }
  }
}

template <typename R>
__global__ void __launch_bounds__(128) k_setup_dr(const SetupDrParams p) {
  using namespace setup;
  const long long face = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (face >= p.m) {
    return;
  }
  const int plusSide = p.faceInfo[3 * face], minusSide = p.faceInfo[3 * face + 1];
  const double lamP = p.materialPlus[3 * face], muP = p.materialPlus[3 * face + 1], rhoP = p.materialPlus[3 * face + 2];
  const double lamM = p.materialMinus[3 * face], muM = p.materialMinus[3 * face + 1], rhoM = p.materialMinus[3 * face + 2];
  // impedances and eta (DynamicRuptureMatrices.cpp:341-359)
  const double zp = rhoP * sqrt((lamP + 2 * muP) / rhoP), zs = rhoP * sqrt(muP / rhoP);
  const double zpN = rhoM * sqrt((lamM + 2 * muM) / rhoM), zsN = rhoM * sqrt(muM / rhoM);
  R* imp = static_cast<R*>(p.impAndEta) + 11 * face;
  imp[0] = R(zp); imp[1] = R(zs); imp[2] = R(zpN); imp[3] = R(zsN);
  imp[4] = R(1.0 / (1.0 / zp + 1.0 / zpN)); imp[5] = R(1.0 / (1.0 / zs + 1.0 / zsN)); imp[6] = R(1.0 / zs + 1.0 / zsN);
  imp[7] = R(1.0 / zp); imp[8] = R(1.0 / zs); imp[9] = R(1.0 / zpN); imp[10] = R(1.0 / zsN);
  // face frame of the plus side; areas and volumes of both sides
  double frame[3][3], surfP, volP, frameM[3][3], surfM, volM;
  faceGeometry(p.coordsPlus + 12 * face, plusSide, frame, surfP, volP);
  faceGeometry(p.coordsMinus + 12 * face, minusSide, frameM, surfM, volM);
  M9 T, Ti, A, tmp, TT;
  rotationMatrices(frame, T, Ti);
  R* tinvT = static_cast<R*>(p.tinvT) + 81 * face;
  for (int a = 0; a < 9; ++a) {
    for (int b = 0; b < 9; ++b) {
      tinvT[a + 9 * b] = static_cast<R>(Ti.a[b][a]);   // TinvT(a, b) = Tinv(b, a)
    }
  }
  if (p.doubledSurfaceArea != nullptr) {
    static_cast<R*>(p.doubledSurfaceArea)[face] = static_cast<R>(2.0 * surfM);
  }
  // fluxSolver(q, p) = fluxScale sum_k star0(q, k) T(p, k)  (dynamic_rupture.py:88, rotateFluxMatrix)
  transpose(T, TT);
  for (int side = 0; side < 2; ++side) {
    coefficientMatrixT(side == 0 ? lamP : lamM, side == 0 ? muP : muM, side == 0 ? rhoP : rhoM, 0, A);
    mul(A, TT, tmp);
    const double scale = side == 0 ? -2.0 * surfP / (6.0 * volP) : 2.0 * surfM / (6.0 * volM);
    R* fs = static_cast<R*>(side == 0 ? p.fluxSolverPlus : p.fluxSolverMinus) + 81 * face;
    for (int q = 0; q < 9; ++q) {
      for (int c = 0; c < 9; ++c) {
        fs[c + 9 * q] = static_cast<R>(scale * tmp.a[c][q]);   // column-major [c + 9 q]: row c, column q
      }
    }
  }
}

}  // namespace ssb
