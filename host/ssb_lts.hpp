// C++ host of the clustered local-time-stepping (rate 2) element update on ONE rank of an x-slab partition, above the
// C ABI (include/seissol_b200.h), with the halo exchange on NCCL.  The C++ twin of seissol_b200/lts.py.
//
// What it replaces in the reference: the LTS part of the time-cluster step (src/Solver/TimeStepping/TimeCluster.cpp:
// computeLocalIntegration :358-538 with buffer reset / accumulation, computeNeighboringIntegration :556-998 with
// TimeCommon::computeIntegrals, src/Kernels/TimeCommon.cpp:40-125), the per-cell LtsSetup bits that drive it
// (src/Initializer/LtsSetup.h:21-111, InitProcedure/Internal/LtsSetup.cpp:96-189), the actor order of rate-2 clusters
// (AbstractTimeCluster.cpp:35-199) and the ghost/copy exchange of buffers and/or derivatives
// (DirectGhostTimeCluster.cpp:22-105), which moves from MPI onto ncclSend/ncclRecv over NVLink on a communication
// stream: only the corrections that read the ghost rows of a cluster wait for its message.
//
//   cluster c advances with dt_c = 2^c dt_0.  For a cell e in cluster c and a regular face with neighbor n:
//     cl(n) == c     : the neighbor's time-integrated DOFs of the same interval; if the neighbor is an own cell that
//                      accumulates for a larger cluster, integrate its derivatives over the full interval instead;
//     cl(n) == c + 1 : integrate its derivatives over my sub-interval relative to its expansion point (for a ghost cell
//                      the derivative family arrives with the halo of cluster c + 1);
//     cl(n) == c - 1 : the sum of its two steps (own cell: its accumulation row; ghost cell: accumulated on arrival).
//
// The mesh partition (which cells a rank owns, its ghost cells, the copy / ghost regions) is INPUT: a `Layout`, read
// from the file seissol_b200/partition.py::export_layout writes (in SeisSol it comes from the mesh reader and the
// clustering, both outside the hot path).
#pragma once
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <map>
#include <numeric>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef SSB_WITH_NCCL
#include <nccl.h>
#endif

#include "ssb_host.hpp"

namespace seissol_b200::lts {

using i64 = long long;

struct Region {
  i64 peer, side, cluster, start, count;
};

// seissol_b200/partition.py::RankLayout
struct Layout {
  i64 rank = 0, world = 1, nOwn = 0, nGhost = 0, nClusters = 1;
  std::vector<i64> gid, cluster;      // [nOwn + nGhost]
  std::vector<i64> neighbors;         // [nOwn][4] index into [own | ghost] (own index on boundary faces)
  std::vector<uint8_t> faceRel;       // [nOwn][4][2]
  std::vector<uint8_t> faceTypes;     // [nOwn][4]
  std::vector<Region> send, recv;

  static Layout read(const std::string& path) {
    FILE* f = std::fopen(path.c_str(), "rb");
    if (f == nullptr) {
      throw std::runtime_error("cannot open layout file " + path);
    }
    auto rd = [&](void* p, size_t bytes) {
      if (bytes > 0 && std::fread(p, 1, bytes, f) != bytes) {
        std::fclose(f);
        throw std::runtime_error("short layout file " + path);
      }
    };
    i64 h[8];
    rd(h, sizeof(h));
    if (h[0] != 0x5353424C54533031ll) {   // "SSBLTS01"
      std::fclose(f);
      throw std::runtime_error("not a layout file: " + path);
    }
    Layout l;
    l.rank = h[1]; l.world = h[2]; l.nOwn = h[3]; l.nGhost = h[4]; l.nClusters = h[5];
    const i64 nSend = h[6], nRecv = h[7], n = l.nOwn, na = l.nOwn + l.nGhost;
    l.gid.resize(na); l.cluster.resize(na); l.neighbors.resize(4 * n); l.faceRel.resize(8 * n); l.faceTypes.resize(4 * n);
    rd(l.gid.data(), na * 8); rd(l.cluster.data(), na * 8); rd(l.neighbors.data(), 4 * n * 8);
    rd(l.faceRel.data(), 8 * n); rd(l.faceTypes.data(), 4 * n);
    l.send.resize(nSend); l.recv.resize(nRecv);
    rd(l.send.data(), nSend * sizeof(Region)); rd(l.recv.data(), nRecv * sizeof(Region));
    std::fclose(f);
    return l;
  }
};

struct ClusterRange {
  i64 start, accEnd, derEnd, end;   // [start, accEnd): accumulate; [start, derEnd): keep derivatives
};
struct CellFace {
  i64 cell;
  int face;
};
struct SendList {
  i64 peer, side;
  std::vector<i64> rows;
};
struct RecvSlice {
  i64 peer, side, start, count;
};

// seissol_b200/lts.py::LtsPlan / build_plan
struct Plan {
  i64 n = 0, nGhost = 0, nClusters = 1, world = 1, rank = 0;
  std::vector<i64> gid, cluster, clusterExt, neighbors;
  std::vector<uint8_t> faceRel, faceTypes;
  std::vector<ClusterRange> ranges;
  std::vector<i64> derRow, accRow;          // [n + nGhost]
  std::vector<int> kind;                    // [n][4]
  std::vector<i64> scratchRow;              // [n][4]
  std::vector<std::vector<CellFace>> taylorFull, taylorSub;
  i64 nDer = 0, nAcc = 0, nScratch = 0, nDerOwn = 0, nAccOwn = 0;
  std::vector<std::vector<SendList>> send, sendDer;     // per cluster
  std::vector<std::vector<RecvSlice>> recv, recvDer;
  std::vector<std::vector<i64>> ghostAccCells, ghostAccRows;

  i64 updatesPerMacroStep() const {
    i64 u = 0;
    for (i64 c = 0; c < nClusters; ++c) {
      u += (ranges[c].end - ranges[c].start) << (nClusters - 1 - c);
    }
    return u;
  }
};

inline Plan buildPlan(const Layout& lay) {
  const i64 n = lay.nOwn, ng = lay.nGhost, C = lay.nClusters;
  const auto& clAll = lay.cluster;
  auto regularL = [&](i64 e, int f) { return lay.faceTypes[4 * e + f] == 0; };
  std::vector<int> group(n);
  for (i64 e = 0; e < n; ++e) {
    bool larger = false, smaller = false;
    for (int f = 0; f < 4; ++f) {
      if (!regularL(e, f)) {
        continue;
      }
      const i64 d = clAll[lay.neighbors[4 * e + f]] - clAll[e];
      if (d > 1 || d < -1) {
        throw std::runtime_error("neighboring cells may differ by at most one time cluster");
      }
      larger = larger || d == 1;
      smaller = smaller || d == -1;
    }
    group[e] = larger ? 0 : (smaller ? 1 : 2);
  }
  std::vector<i64> order(n);
  std::iota(order.begin(), order.end(), 0);
  std::stable_sort(order.begin(), order.end(), [&](i64 a, i64 b) {
    if (clAll[a] != clAll[b]) return clAll[a] < clAll[b];
    if (group[a] != group[b]) return group[a] < group[b];
    return lay.gid[a] < lay.gid[b];
  });
  std::vector<i64> newIndex(n + ng);
  for (i64 i = 0; i < n; ++i) {
    newIndex[order[i]] = i;
  }
  for (i64 g = 0; g < ng; ++g) {
    newIndex[n + g] = n + g;
  }
  Plan p;
  p.n = n; p.nGhost = ng; p.nClusters = C; p.world = lay.world; p.rank = lay.rank;
  p.gid.resize(n); p.cluster.resize(n); p.clusterExt.resize(n + ng); p.neighbors.resize(4 * n);
  p.faceRel.resize(8 * n); p.faceTypes.resize(4 * n);
  std::vector<int> groupO(n);
  for (i64 i = 0; i < n; ++i) {
    const i64 o = order[i];
    p.gid[i] = lay.gid[o];
    p.cluster[i] = p.clusterExt[i] = clAll[o];
    groupO[i] = group[o];
    for (int f = 0; f < 4; ++f) {
      p.neighbors[4 * i + f] = newIndex[lay.neighbors[4 * o + f]];
      p.faceTypes[4 * i + f] = lay.faceTypes[4 * o + f];
      p.faceRel[8 * i + 2 * f] = lay.faceRel[8 * o + 2 * f];
      p.faceRel[8 * i + 2 * f + 1] = lay.faceRel[8 * o + 2 * f + 1];
    }
  }
  for (i64 g = 0; g < ng; ++g) {
    p.clusterExt[n + g] = clAll[n + g];
  }
  p.ranges.resize(C);
  {
    i64 pos = 0;
    for (i64 c = 0; c < C; ++c) {
      ClusterRange r{pos, pos, pos, pos};
      while (pos < n && p.cluster[pos] == c) {
        ++pos;
      }
      r.end = pos;
      for (i64 i = r.start; i < r.end; ++i) {
        r.accEnd += groupO[i] == 0 ? 1 : 0;
        r.derEnd += groupO[i] <= 1 ? 1 : 0;
      }
      p.ranges[c] = r;
    }
  }
  auto regular = [&](i64 e, int f) { return p.faceTypes[4 * e + f] == 0; };
  auto nbOf = [&](i64 e, int f) { return p.neighbors[4 * e + f]; };
  // ghost cells seen from a smaller / larger own cluster
  std::vector<char> ghostDer(ng, 0), ghostAcc(ng, 0), sendsDer(n, 0);
  for (i64 e = 0; e < n; ++e) {
    for (int f = 0; f < 4; ++f) {
      const i64 nb = nbOf(e, f);
      if (!regular(e, f) || nb < n) {
        continue;
      }
      const i64 d = p.clusterExt[nb] - p.cluster[e];
      if (d == 1) {
        ghostDer[nb - n] = 1;
      } else if (d == -1) {
        ghostAcc[nb - n] = 1;
        sendsDer[e] = 1;
      }
    }
  }
  p.derRow.assign(n + ng, -1);
  p.accRow.assign(n + ng, -1);
  for (i64 i = 0; i < n; ++i) {
    if (groupO[i] <= 1) p.derRow[i] = p.nDerOwn++;
    if (groupO[i] == 0) p.accRow[i] = p.nAccOwn++;
  }
  p.nDer = p.nDerOwn;
  p.nAcc = p.nAccOwn;
  for (i64 g = 0; g < ng; ++g) {
    if (ghostDer[g]) p.derRow[n + g] = p.nDer++;
    if (ghostAcc[g]) p.accRow[n + g] = p.nAcc++;
  }
  p.kind.assign(4 * n, 0);
  p.scratchRow.assign(4 * n, -1);
  p.taylorFull.resize(C);
  p.taylorSub.resize(C);
  for (i64 e = 0; e < n; ++e) {
    for (int f = 0; f < 4; ++f) {
      if (!regular(e, f)) {
        continue;
      }
      const i64 nb = nbOf(e, f);
      const i64 d = p.clusterExt[nb] - p.cluster[e];
      int k = 0;
      if (d == 0 && nb < n && groupO[nb] == 0) {
        k = 1;
      } else if (d == 1) {
        k = 2;
      }
      p.kind[4 * e + f] = k;
      if (k > 0) {
        if (p.derRow[nb] < 0) {
          throw std::runtime_error("a neighbor that is integrated in time must keep its derivatives");
        }
        p.scratchRow[4 * e + f] = p.nScratch++;
        (k == 1 ? p.taylorFull : p.taylorSub)[p.cluster[e]].push_back(CellFace{e, f});
      }
      if (d == -1 && p.accRow[nb] < 0) {
        throw std::runtime_error("a neighbor in a smaller cluster must accumulate");
      }
    }
  }
  p.send.resize(C); p.recv.resize(C); p.sendDer.resize(C); p.recvDer.resize(C);
  p.ghostAccCells.resize(C); p.ghostAccRows.resize(C);
  for (const Region& reg : lay.send) {
    SendList s{reg.peer, reg.side, {}}, sd{reg.peer, reg.side, {}};
    for (i64 i = reg.start; i < reg.start + reg.count; ++i) {
      const i64 e = newIndex[i];
      s.rows.push_back(e);
      if (sendsDer[e]) {
        sd.rows.push_back(p.derRow[e]);
      }
    }
    p.send[reg.cluster].push_back(s);
    if (!sd.rows.empty()) {
      p.sendDer[reg.cluster].push_back(sd);
    }
  }
  for (const Region& reg : lay.recv) {
    p.recv[reg.cluster].push_back(RecvSlice{reg.peer, reg.side, reg.start, reg.count});
    i64 first = -1, cnt = 0;
    for (i64 i = reg.start; i < reg.start + reg.count; ++i) {
      if (p.derRow[i] >= 0) {
        first = first < 0 ? p.derRow[i] : first;
        ++cnt;
      }
    }
    if (cnt > 0) {
      p.recvDer[reg.cluster].push_back(RecvSlice{reg.peer, reg.side, first, cnt});
    }
  }
  for (i64 g = 0; g < ng; ++g) {
    if (ghostAcc[g]) {
      p.ghostAccCells[clAll[n + g]].push_back(n + g);
      p.ghostAccRows[clAll[n + g]].push_back(p.accRow[n + g]);
    }
  }
  return p;
}

// MonomialBasis::integrate (src/Numerical/TimeBasis.h:82-96) for [t0, t1] relative to the expansion point
inline std::vector<double> integrateCoeffs(int order, double t0, double t1) {
  std::vector<double> c(order);
  double fact = 1.0;
  for (int d = 0; d < order; ++d) {
    fact *= d + 1;
    c[d] = (std::pow(t1, d + 1) - std::pow(t0, d + 1)) / fact;
  }
  return c;
}

template <typename T>
struct DeviceVector {
  T* p = nullptr;
  size_t n = 0;
  DeviceVector() = default;
  explicit DeviceVector(size_t count, bool zero = true) { alloc(count, zero); }
  DeviceVector(const DeviceVector&) = delete;
  DeviceVector& operator=(const DeviceVector&) = delete;
  DeviceVector(DeviceVector&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DeviceVector& operator=(DeviceVector&& o) noexcept {
    if (this != &o) {
      cudaFree(p);
      p = o.p; n = o.n;
      o.p = nullptr; o.n = 0;
    }
    return *this;
  }
  ~DeviceVector() { cudaFree(p); }
  void alloc(size_t count, bool zero = true) {
    cudaFree(p);
    n = count;
    checkCuda(cudaMalloc(reinterpret_cast<void**>(&p), (count > 0 ? count : 1) * sizeof(T)), "cudaMalloc");
    if (zero) {
      checkCuda(cudaMemset(p, 0, (count > 0 ? count : 1) * sizeof(T)), "cudaMemset");
    }
  }
  void upload(const std::vector<T>& h) {
    alloc(h.size(), false);
    if (!h.empty()) {
      checkCuda(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice), "cudaMemcpy H2D");
    }
  }
};

// NCCL communicator of the box (one process per GPU); without SSB_WITH_NCCL only single-rank plans run
struct Comm {
#ifdef SSB_WITH_NCCL
  ncclComm_t comm = nullptr;
#endif
  int rank = 0, world = 1;
};

// seissol_b200/lts.py::CudaLtsBackend: predict / exchange / correct of one cluster through the C ABI
template <typename Real>
class Backend {
  public:
  // dofs [n][9 ld], localInt [n][396], neighInt [n][324]: device arrays in PLAN order (plan.gid)
  Backend(Context& ctx, const Plan& plan, Real* dofs, const Real* localInt, const Real* neighInt, bool faceProj,
          const Comm* comm = nullptr)
      : ctx_(ctx), p_(plan), dofs_(dofs), li_(localInt), ni_(neighInt), fp_(faceProj), comm_(comm) {
    const i64 n = p_.n;
    w_ = ctx.dofsSize();
    ders_ = ctx.derivativesSize();
    const int b2 = ctx.order() * (ctx.order() + 1) / 2;
    prj_ = 9 * ((b2 + 3) / 4 * 4);
    fpw_ = 4 * prj_;
    halow_ = fp_ ? fpw_ : w_;
    derivs_.alloc(size_t(std::max<i64>(p_.nDer, 1)) * ders_);
    faceTypes_.upload(p_.faceTypes);
    faceRel_.upload(p_.faceRel);
    idofs_.alloc(size_t(fp_ ? n : n + p_.nGhost) * w_);
    if (fp_) {
      faceProj_.alloc(size_t(n + p_.nGhost) * fpw_);
    }
    acc_.alloc(size_t(std::max<i64>(p_.nAcc, 1)) * halow_);
    scratch_.alloc(size_t(std::max<i64>(p_.nScratch, 1)) * (fp_ ? prj_ : w_));
    halo_ = fp_ ? faceProj_.p : idofs_.p;
    // LTS::FaceNeighbors: where the neighbor flux of (cell, face) reads.  Face-projection path: the kernel adds
    // j * prj reals to the entry, so a Taylor scratch row (one block) is entered as (row - j * prj).
    std::vector<const void*> ptr(4 * n);
    for (i64 e = 0; e < n; ++e) {
      for (int f = 0; f < 4; ++f) {
        const i64 nb = p_.neighbors[4 * e + f];
        const int k = p_.kind[4 * e + f];
        const int j = p_.faceRel[8 * e + 2 * f];
        const bool smaller = p_.faceTypes[4 * e + f] == 0 && p_.clusterExt[nb] == p_.cluster[e] - 1;
        const Real* q;
        if (k > 0) {
          q = scratch_.p + p_.scratchRow[4 * e + f] * (fp_ ? prj_ : w_) - (fp_ ? j * prj_ : 0);
        } else if (smaller) {
          q = acc_.p + p_.accRow[nb] * halow_;
        } else {
          q = halo_ + nb * halow_;
        }
        ptr[4 * e + f] = q;
      }
    }
    nbPtrs_.upload(ptr);
    taylor_.resize(2 * p_.nClusters);
    for (i64 c = 0; c < p_.nClusters; ++c) {
      for (int which = 0; which < 2; ++which) {
        const auto& list = which == 0 ? p_.taylorFull[c] : p_.taylorSub[c];
        TaylorBatch& tb = taylor_[2 * c + which];
        tb.count = i64(list.size());
        std::vector<const void*> src(list.size()), dst(list.size());
        std::vector<uint8_t> fj(list.size());
        for (size_t i = 0; i < list.size(); ++i) {
          const i64 e = list[i].cell;
          const int f = list[i].face;
          src[i] = derivs_.p + p_.derRow[p_.neighbors[4 * e + f]] * ders_;
          dst[i] = scratch_.p + p_.scratchRow[4 * e + f] * (fp_ ? prj_ : w_);
          fj[i] = p_.faceRel[8 * e + 2 * f];
        }
        tb.src.upload(src); tb.dst.upload(dst); tb.faceJ.upload(fj);
      }
    }
    // processing order of the neighbor kernel per cluster: sorted by the (j, h) face relations (no divergence)
    perm_.resize(p_.nClusters);
    for (i64 c = 0; c < p_.nClusters; ++c) {
      const ClusterRange& r = p_.ranges[c];
      std::vector<int> idx(r.end - r.start);
      std::iota(idx.begin(), idx.end(), 0);
      auto key = [&](int i) {
        const uint8_t* rel = &p_.faceRel[8 * (r.start + i)];
        return (rel[0] + 4 * rel[2] + 16 * rel[4] + 64 * rel[6]) * 81 + rel[1] + 3 * rel[3] + 9 * rel[5] + 27 * rel[7];
      };
      std::stable_sort(idx.begin(), idx.end(), [&](int a, int b) { return key(a) < key(b); });
      perm_[c].upload(idx);
    }
    // halo: gather lists and persistent pack buffers
    sendIdx_.resize(p_.nClusters); sendDerIdx_.resize(p_.nClusters); ghostAccSrc_.resize(p_.nClusters);
    for (i64 c = 0; c < p_.nClusters; ++c) {
      for (const SendList& s : p_.send[c]) {
        sendIdx_[c].emplace_back();
        Pack& pk = sendIdx_[c].back();
        pk.peer = s.peer; pk.side = s.side; pk.count = i64(s.rows.size());
        std::vector<const void*> src(s.rows.size());
        for (size_t i = 0; i < s.rows.size(); ++i) src[i] = halo_ + s.rows[i] * halow_;
        pk.src.upload(src);
        pk.buf.alloc(s.rows.size() * halow_);
      }
      for (const SendList& s : p_.sendDer[c]) {
        sendDerIdx_[c].emplace_back();
        Pack& pk = sendDerIdx_[c].back();
        pk.peer = s.peer; pk.side = s.side; pk.count = i64(s.rows.size());
        std::vector<const void*> src(s.rows.size());
        for (size_t i = 0; i < s.rows.size(); ++i) src[i] = derivs_.p + s.rows[i] * ders_;
        pk.src.upload(src);
        pk.buf.alloc(s.rows.size() * ders_);
      }
      std::vector<const void*> src(p_.ghostAccCells[c].size());
      for (size_t i = 0; i < src.size(); ++i) src[i] = halo_ + p_.ghostAccCells[c][i] * halow_;
      ghostAccSrc_[c].upload(src);
    }
    if (p_.world > 1) {
      checkCuda(cudaStreamCreateWithFlags(&commStream_, cudaStreamNonBlocking), "cudaStreamCreate");
      evPred_.resize(p_.nClusters); evHalo_.resize(p_.nClusters);
      for (i64 c = 0; c < p_.nClusters; ++c) {
        checkCuda(cudaEventCreateWithFlags(&evPred_[c], cudaEventDisableTiming), "cudaEventCreate");
        checkCuda(cudaEventCreateWithFlags(&evHalo_[c], cudaEventDisableTiming), "cudaEventCreate");
      }
      pending_.assign(p_.nClusters, 0);
    }
  }

  i64 launches = 0;
  cudaStream_t stream = nullptr;   // the compute stream

  void predict(i64 c, const ClusterRange& r, const std::vector<double>& coeffs, bool reset) {
    ssb_ctx* h = ctx_.handle();
    const ssb_array_t none{nullptr, nullptr, 0};
    if (r.derEnd > r.start) {
      const i64 d0 = p_.derRow[r.start];
      check(ssb_ader(h, rows(dofs_, r.start, w_), rows(idofs_.p, r.start, w_), rows(derivs_.p, d0, ders_), rows(li_, r.start, 396),
                     coeffs.data(), r.derEnd - r.start, nullptr, stream), "ssb_ader");
      ++launches;
    }
    if (r.end > r.derEnd) {
      check(ssb_ader(h, rows(dofs_, r.derEnd, w_), rows(idofs_.p, r.derEnd, w_), none, rows(li_, r.derEnd, 396), coeffs.data(),
                     r.end - r.derEnd, nullptr, stream), "ssb_ader");
      ++launches;
    }
    if (fp_) {
      check(ssb_local_faceproj(h, rows(dofs_, r.start, w_), rows(idofs_.p, r.start, w_), rows(li_, r.start, 396),
                               rows(faceProj_.p, r.start, fpw_), faceTypes_.p + 4 * r.start, r.end - r.start, nullptr, stream),
            "ssb_local_faceproj");
    } else {
      check(ssb_local(h, rows(dofs_, r.start, w_), rows(idofs_.p, r.start, w_), rows(li_, r.start, 396),
                      faceTypes_.p + 4 * r.start, r.end - r.start, nullptr, stream), "ssb_local");
    }
    ++launches;
    if (r.accEnd > r.start) {
      const i64 a0 = p_.accRow[r.start];
      if (fp_) {
        check(ssb_accumulate_rows(h, rows(acc_.p, a0, fpw_), rows(faceProj_.p, r.start, fpw_), reset ? 1 : 0, r.accEnd - r.start,
                                  int(fpw_), stream), "ssb_accumulate_rows");
      } else {
        check(ssb_lts_accumulate(h, rows(acc_.p, a0, w_), rows(idofs_.p, r.start, w_), reset ? 1 : 0, r.accEnd - r.start, stream),
              "ssb_lts_accumulate");
      }
      ++launches;
    }
  }

  // Halo of cluster c on the communication stream behind the prediction of cluster c: per region the copy cells' rows
  // (and, where a peer integrates them, their derivative families) are packed into a persistent buffer and sent; the
  // rows are received straight into the ghost rows the neighbor kernel reads; ghost rows that a larger own cluster
  // reads are accumulated over the two steps on arrival.
  void exchange(i64 c, bool reset) {
#ifdef SSB_WITH_NCCL
    ssb_ctx* h = ctx_.handle();
    checkCuda(cudaEventRecord(evPred_[c], stream), "cudaEventRecord");
    checkCuda(cudaStreamWaitEvent(commStream_, evPred_[c], 0), "cudaStreamWaitEvent");
    const ncclDataType_t ty = sizeof(Real) == 4 ? ncclFloat : ncclDouble;
    for (Pack& pk : sendIdx_[c]) {
      check(ssb_accumulate_rows(h, rows(pk.buf.p, 0, halow_), ssb_array_t{pk.src.p, nullptr, 0}, 1, pk.count, int(halow_),
                                commStream_), "pack");
    }
    for (Pack& pk : sendDerIdx_[c]) {
      check(ssb_accumulate_rows(h, rows(pk.buf.p, 0, ders_), ssb_array_t{pk.src.p, nullptr, 0}, 1, pk.count, int(ders_),
                                commStream_), "pack derivatives");
    }
    ncclGroupStart();
    for (Pack& pk : sendIdx_[c]) {
      ncclSend(pk.buf.p, size_t(pk.count) * halow_, ty, int(pk.peer), comm_->comm, commStream_);
    }
    for (Pack& pk : sendDerIdx_[c]) {
      ncclSend(pk.buf.p, size_t(pk.count) * ders_, ty, int(pk.peer), comm_->comm, commStream_);
    }
    // NCCL matches the messages of a (sender, receiver) pair in order: regions are posted in the same order on both
    // ends (integrals by side, then derivatives by side; the receiver's side s mirrors the sender's side 1 - s)
    auto bySideDesc = [](const RecvSlice& a, const RecvSlice& b) { return a.side > b.side; };
    std::vector<RecvSlice> rv = p_.recv[c], rd = p_.recvDer[c];
    std::stable_sort(rv.begin(), rv.end(), bySideDesc);
    std::stable_sort(rd.begin(), rd.end(), bySideDesc);
    for (const RecvSlice& s : rv) {
      ncclRecv(halo_ + s.start * halow_, size_t(s.count) * halow_, ty, int(s.peer), comm_->comm, commStream_);
    }
    for (const RecvSlice& s : rd) {
      ncclRecv(derivs_.p + s.start * ders_, size_t(s.count) * ders_, ty, int(s.peer), comm_->comm, commStream_);
    }
    ncclGroupEnd();
    if (ghostAccSrc_[c].n > 0) {
      const i64 a0 = p_.ghostAccRows[c][0];
      check(ssb_accumulate_rows(h, rows(acc_.p, a0, halow_), ssb_array_t{ghostAccSrc_[c].p, nullptr, 0}, reset ? 1 : 0,
                                i64(ghostAccSrc_[c].n), int(halow_), commStream_), "ghost accumulation");
      ++launches;
    }
    checkCuda(cudaEventRecord(evHalo_[c], commStream_), "cudaEventRecord");
    pending_[c] = 1;
#else
    (void)c; (void)reset;
    throw std::runtime_error("built without NCCL: single-rank plans only");
#endif
  }

  void correct(i64 c, const ClusterRange& r, const std::vector<double>& full, const std::vector<double>& sub) {
    ssb_ctx* h = ctx_.handle();
    if (p_.world > 1) {
      for (i64 k = c - 1; k <= c + 1; ++k) {   // ghost rows of my cluster, accumulated smaller ones, derivatives of larger ones
        if (k >= 0 && k < p_.nClusters && pending_[k]) {
          checkCuda(cudaStreamWaitEvent(stream, evHalo_[k], 0), "cudaStreamWaitEvent");
          if (k == c) pending_[k] = 0;
        }
      }
    }
    for (int which = 0; which < 2; ++which) {
      TaylorBatch& tb = taylor_[2 * c + which];
      if (tb.count == 0) {
        continue;
      }
      const double* cf = which == 0 ? full.data() : sub.data();
      if (fp_) {
        check(ssb_taylor_faceproj(h, ssb_array_t{tb.src.p, nullptr, 0}, ssb_array_t{tb.dst.p, nullptr, 0}, tb.faceJ.p, cf,
                                  tb.count, stream), "ssb_taylor_faceproj");
      } else {
        check(ssb_taylor(h, ssb_array_t{tb.src.p, nullptr, 0}, ssb_array_t{tb.dst.p, nullptr, 0}, cf, tb.count, stream),
              "ssb_taylor");
      }
      ++launches;
    }
    const ssb_array_t none{nullptr, nullptr, 0};
    if (fp_) {
      check(ssb_neigh_faceproj(h, rows(dofs_, r.start, w_), nbPtrs_.p + 4 * r.start, nullptr, none, rows(ni_, r.start, 324),
                               faceTypes_.p + 4 * r.start, faceRel_.p + 8 * r.start, r.end - r.start, perm_[c].p, stream),
            "ssb_neigh_faceproj");
    } else {
      check(ssb_neigh(h, rows(dofs_, r.start, w_), nbPtrs_.p + 4 * r.start, nullptr, none, rows(ni_, r.start, 324),
                      faceTypes_.p + 4 * r.start, faceRel_.p + 8 * r.start, r.end - r.start, perm_[c].p, stream), "ssb_neigh");
    }
    ++launches;
  }

  private:
  struct TaylorBatch {
    DeviceVector<const void*> src, dst;
    DeviceVector<uint8_t> faceJ;
    i64 count = 0;
  };
  struct Pack {
    i64 peer = 0, side = 0, count = 0;
    DeviceVector<const void*> src;
    DeviceVector<Real> buf;
  };
  template <typename T>
  static ssb_array_t rows(T* base, i64 first, i64 width) {
    return ssb_array_t{nullptr, base + first * width, i64(width * sizeof(Real))};
  }
  Context& ctx_;
  const Plan& p_;
  Real* dofs_;
  const Real* li_;
  const Real* ni_;
  bool fp_;
  const Comm* comm_;
  i64 w_ = 0, ders_ = 0, prj_ = 0, fpw_ = 0, halow_ = 0;
  DeviceVector<Real> derivs_, idofs_, faceProj_, acc_, scratch_;
  Real* halo_ = nullptr;
  DeviceVector<uint8_t> faceTypes_, faceRel_;
  DeviceVector<const void*> nbPtrs_;
  std::vector<TaylorBatch> taylor_;
  std::vector<DeviceVector<int>> perm_;
  std::vector<std::vector<Pack>> sendIdx_, sendDerIdx_;
  std::vector<DeviceVector<const void*>> ghostAccSrc_;
  cudaStream_t commStream_ = nullptr;
  std::vector<cudaEvent_t> evPred_, evHalo_;
  std::vector<char> pending_;
};

// seissol_b200/lts.py::LtsSolver: one call advances every cluster to t + dt_{C-1}
template <typename BackendT>
class Solver {
  public:
  Solver(const Plan& plan, BackendT& be, int order, double dt0) : p_(plan), be_(be), order_(order), dt0_(dt0) {}
  void macroStep() {
    const i64 C = p_.nClusters, fine = i64(1) << (C - 1);
    for (i64 s = 0; s < fine; ++s) {
      for (i64 c = 0; c < C; ++c) {
        const ClusterRange& r = p_.ranges[c];
        if (s % (i64(1) << c) == 0) {
          const bool reset = s % (i64(1) << (c + 1)) == 0;
          if (r.end > r.start) {
            be_.predict(c, r, integrateCoeffs(order_, 0.0, dt0_ * double(i64(1) << c)), reset);
          }
          if (p_.world > 1 && (!p_.send[c].empty() || !p_.recv[c].empty())) {
            be_.exchange(c, reset);
          }
        }
      }
      for (i64 c = 0; c < C; ++c) {
        const ClusterRange& r = p_.ranges[c];
        if ((s + 1) % (i64(1) << c) == 0 && r.end > r.start) {
          const double dt = dt0_ * double(i64(1) << c);
          const i64 start = (s + 1) - (i64(1) << c);
          const double off = double(start % (i64(1) << (c + 1))) * dt0_;
          be_.correct(c, r, integrateCoeffs(order_, 0.0, dt), integrateCoeffs(order_, off, off + dt));
        }
      }
    }
  }

  private:
  const Plan& p_;
  BackendT& be_;
  int order_;
  double dt0_;
};

// seissol_b200/mesh.py::element_values: deterministic bounded values per (global element id, entry index)
inline double elementValue(i64 gid, i64 entry, i64 streamId, double scale) {
  uint64_t h = uint64_t(gid) * 0x9E3779B97F4A7C15ull + (uint64_t(entry) + (uint64_t(streamId) << 20)) * 0xC2B2AE3D27D4EB4Full;
  h ^= h >> 33; h *= 0xFF51AFD7ED558CCDull; h ^= h >> 33; h *= 0xC4CEB9FE1A85EC53ull; h ^= h >> 33;
  const double u = double(h >> 40) / double(1 << 24);
  return (2.0 * u - 1.0) * scale;
}

}  // namespace seissol_b200::lts
