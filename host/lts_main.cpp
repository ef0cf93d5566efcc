// ssb_lts_proxy: clustered LTS (rate 2) of the elastic element update on one rank of a partitioned mesh, C++ host
// (host/ssb_lts.hpp) above the C ABI, halo exchange on NCCL (one process per GPU; RANK / WORLD_SIZE / LOCAL_RANK as
// torchrun sets them).  The partition of the mesh is input: one layout file per rank (seissol_b200/partition.py::
// export_layout); the element data are the deterministic values of seissol_b200/mesh.py::element_values, so that any
// partition -- and the Python driver -- see identical inputs.
//
//   ssb_lts_proxy --layout PREFIX [--order 6] [--precision s|d] [--steps 2] [--warmup 0] [--dt0 0.5] [--faceproj 0|1]
//                 [--dump PREFIX] [--idfile PATH]
//   reads PREFIX_<rank>.bin, writes <dump>_<rank>_Q.bin / _gid.bin, prints one JSON line per rank.
//   --plan-only PREFIX: no device needed; writes the LTS plan of the rank (PREFIX_<rank>_plan.bin: int64 arrays gid,
//   cluster, neighbors, kind, scratch_row, der_row, acc_row, ranges) for the comparison with seissol_b200/lts.py.
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <thread>

#include "ssb_lts.hpp"

using namespace seissol_b200;
using lts::i64;

namespace {

struct Args {
  std::string layout, dump, planDump, idfile = "/tmp/ssb_lts_nccl_id";
  int order = 6, steps = 2, warmup = 0, faceproj = -1;
  char precision = 's';
  double dt0 = 0.5;
};

int envInt(const char* name, int def) {
  const char* e = std::getenv(name);
  return e != nullptr ? std::atoi(e) : def;
}

#ifdef SSB_WITH_NCCL
// rank 0 publishes the ncclUniqueId in a file (one box, shared /tmp), the other ranks poll for it
void initComm(lts::Comm& comm, const std::string& idfile) {
  ncclUniqueId id;
  if (comm.rank == 0) {
    ncclGetUniqueId(&id);
    const std::string tmp = idfile + ".tmp";
    std::ofstream(tmp, std::ios::binary).write(reinterpret_cast<const char*>(&id), sizeof(id));
    std::rename(tmp.c_str(), idfile.c_str());
  } else {
    for (int tries = 0;; ++tries) {
      std::ifstream in(idfile, std::ios::binary);
      if (in && in.read(reinterpret_cast<char*>(&id), sizeof(id))) {
        break;
      }
      if (tries > 600) {
        throw std::runtime_error("no NCCL id file " + idfile);
      }
      std::this_thread::sleep_for(std::chrono::milliseconds(100));
    }
  }
  if (ncclCommInitRank(&comm.comm, comm.world, id, comm.rank) != ncclSuccess) {
    throw std::runtime_error("ncclCommInitRank failed");
  }
}
#endif

template <typename Real>
int run(const Args& a) {
  lts::Comm comm;
  comm.rank = envInt("RANK", 0);
  comm.world = envInt("WORLD_SIZE", 1);
  const lts::Layout lay = lts::Layout::read(a.layout + "_" + std::to_string(comm.rank) + ".bin");
  if (lay.world != comm.world || lay.rank != comm.rank) {
    throw std::runtime_error("layout file belongs to another rank / world size");
  }
  if (!a.planDump.empty()) {
    const lts::Plan plan = lts::buildPlan(lay);
    std::ofstream out(a.planDump + "_" + std::to_string(comm.rank) + "_plan.bin", std::ios::binary);
    auto wr = [&](const std::vector<i64>& v) { out.write(reinterpret_cast<const char*>(v.data()), v.size() * sizeof(i64)); };
    std::vector<i64> kind(plan.kind.begin(), plan.kind.end()), ranges, counts;
    for (const auto& r : plan.ranges) {
      ranges.insert(ranges.end(), {r.start, r.accEnd, r.derEnd, r.end});
    }
    counts = {plan.n, plan.nGhost, plan.nClusters, plan.nDer, plan.nAcc, plan.nScratch};
    for (i64 c = 0; c < plan.nClusters; ++c) {
      i64 sd = 0, rd = 0;
      for (const auto& s : plan.sendDer[c]) sd += i64(s.rows.size());
      for (const auto& s : plan.recvDer[c]) rd += s.count;
      counts.insert(counts.end(), {sd, rd, i64(plan.ghostAccCells[c].size())});
    }
    wr(counts); wr(plan.gid); wr(plan.cluster); wr(plan.neighbors); wr(kind); wr(plan.scratchRow); wr(plan.derRow);
    wr(plan.accRow); wr(ranges);
    return 0;
  }
  checkCuda(cudaSetDevice(envInt("LOCAL_RANK", 0)), "cudaSetDevice");
  if (comm.world > 1) {
#ifdef SSB_WITH_NCCL
    initComm(comm, a.idfile);
#else
    throw std::runtime_error("built without NCCL: single rank only");
#endif
  }
  Context ctx(a.order, int(sizeof(Real)));
  const lts::Plan plan = lts::buildPlan(lay);
  const i64 n = plan.n, B = ctx.basisFunctions(), ld = ctx.leadingDimension(), w = i64(ctx.dofsSize());
  // element data in plan order (tests/test_lts.py::_cuda_lts_rank)
  std::vector<Real> q(size_t(n) * w, Real(0)), li(size_t(n) * 396), ni(size_t(n) * 324);
  for (i64 e = 0; e < n; ++e) {
    const i64 g = plan.gid[e];
    for (i64 c = 0; c < 9; ++c) {
      for (i64 k = 0; k < B; ++k) {
        q[e * w + c * ld + k] = Real(lts::elementValue(g, c * B + k, 0, 1.0));
      }
    }
    for (i64 i = 0; i < 396; ++i) li[e * 396 + i] = Real(lts::elementValue(g, i, 1, 1e-2));
    for (i64 i = 0; i < 324; ++i) ni[e * 324 + i] = Real(lts::elementValue(g, i, 2, 1e-2));
  }
  lts::DeviceVector<Real> dq, dli, dni;
  dq.upload(q); dli.upload(li); dni.upload(ni);
  const bool faceproj = a.faceproj >= 0 ? a.faceproj != 0 : a.order >= 6;
  lts::Backend<Real> be(ctx, plan, dq.p, dli.p, dni.p, faceproj, &comm);
  lts::Solver<lts::Backend<Real>> solver(plan, be, a.order, a.dt0);
  for (int i = 0; i < a.warmup; ++i) {
    solver.macroStep();
  }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  checkCuda(cudaDeviceSynchronize(), "sync");
  cudaEventRecord(e0, be.stream);
  for (int i = 0; i < a.steps; ++i) {
    solver.macroStep();
  }
  cudaEventRecord(e1, be.stream);
  checkCuda(cudaDeviceSynchronize(), "sync");
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  if (!a.dump.empty()) {
    checkCuda(cudaMemcpy(q.data(), dq.p, q.size() * sizeof(Real), cudaMemcpyDeviceToHost), "D2H");
    const std::string pre = a.dump + "_" + std::to_string(comm.rank);
    std::ofstream(pre + "_Q.bin", std::ios::binary).write(reinterpret_cast<const char*>(q.data()), q.size() * sizeof(Real));
    std::ofstream(pre + "_gid.bin", std::ios::binary)
        .write(reinterpret_cast<const char*>(plan.gid.data()), plan.gid.size() * sizeof(i64));
  }
  const double updates = double(plan.updatesPerMacroStep()) * a.steps;
  std::cout << "{\"name\":\"lts\",\"rank\":" << comm.rank << ",\"world\":" << comm.world << ",\"cells\":" << n
            << ",\"ghost-cells\":" << plan.nGhost << ",\"clusters\":" << plan.nClusters << ",\"macro-steps\":" << a.steps
            << ",\"face-projections\":" << (faceproj ? 1 : 0) << ",\"launches\":" << be.launches << ",\"time\":" << ms * 1e-3
            << ",\"elem-updates-per-s\":" << (ms > 0 ? updates / (ms * 1e-3) : 0.0) << "}" << std::endl;
#ifdef SSB_WITH_NCCL
  if (comm.comm != nullptr) {
    ncclCommDestroy(comm.comm);
  }
#endif
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  Args a;
  for (int i = 1; i < argc; ++i) {
    const std::string k = argv[i];
    auto val = [&]() -> std::string {
      if (i + 1 >= argc) {
        std::cerr << "missing value for " << k << std::endl;
        std::exit(2);
      }
      return argv[++i];
    };
    if (k == "--layout") a.layout = val();
    else if (k == "--order") a.order = std::atoi(val().c_str());
    else if (k == "--precision") a.precision = val()[0];
    else if (k == "--steps") a.steps = std::atoi(val().c_str());
    else if (k == "--warmup") a.warmup = std::atoi(val().c_str());
    else if (k == "--dt0") a.dt0 = std::atof(val().c_str());
    else if (k == "--faceproj") a.faceproj = std::atoi(val().c_str());
    else if (k == "--dump") a.dump = val();
    else if (k == "--idfile") a.idfile = val();
    else if (k == "--plan-only") a.planDump = val();
    else {
      std::cerr << "usage: ssb_lts_proxy --layout PREFIX [--order O] [--precision s|d] [--steps N] [--warmup N] [--dt0 x] "
                   "[--faceproj 0|1] [--dump PREFIX] [--idfile PATH]" << std::endl;
      return 2;
    }
  }
  if (a.layout.empty()) {
    std::cerr << "usage: ssb_lts_proxy --layout PREFIX ..." << std::endl;
    return 2;
  }
  try {
    return a.precision == 'd' ? run<double>(a) : run<float>(a);
  } catch (const std::exception& e) {
    std::cerr << "ssb_lts_proxy: " << e.what() << std::endl;
    return 1;
  }
}
