// ssb_proxy: C++ host of the SeisSol proxy on the B200-native kernels.
//
//   ssb_proxy <cells> <timesteps> <kernel[,kernel...]> [-f plain|json] [--order N] [--precision s|d] [--dump PREFIX]
//
// Same command line, kernel names, composition, timing protocol and output as the reference proxy
// (app/Proxy/Main.cpp, src/Proxy/{Common,Runner,Allocator,KernelDevice}.cpp): one untimed warm-up pass, then `timesteps`
// timed passes between two stream synchronisations; `all` = local then neigh, `local` = ader then localwoader.
// Fake data as fakeData (src/Proxy/Allocator.cpp:49-134): per cell and face (side U{0..3}, orientation U{0..2},
// neighbor U{0..cells-1}) drawn from std::mt19937(cells) in the reference's order; values are bounded
// instead of fillWithStuff.  --dump writes the inputs and the final DOFs as raw arrays for the parity test
// (tests/test_host_cpp.py).  Kernels: all, local, neigh, ader, localwoader, neigh_dr, godunov_dr, all_dr (elastic).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <random>
#include <sstream>
#include <string>
#include <vector>

#include "ssb_host.hpp"

using namespace seissol_b200;

namespace {

enum class Kernel { All, Local, Neighbor, Ader, LocalWOAder, NeighborDR, GodunovDR, AllDR };

bool str2kernel(const std::string& s, std::vector<Kernel>& out) {
  std::stringstream ss(s);
  std::string part;
  while (std::getline(ss, part, ',')) {
    if (part == "all") out.push_back(Kernel::All);
    else if (part == "local") out.push_back(Kernel::Local);
    else if (part == "neigh") out.push_back(Kernel::Neighbor);
    else if (part == "ader") out.push_back(Kernel::Ader);
    else if (part == "localwoader") out.push_back(Kernel::LocalWOAder);
    else if (part == "neigh_dr") out.push_back(Kernel::NeighborDR);
    else if (part == "godunov_dr") out.push_back(Kernel::GodunovDR);
    else if (part == "all_dr") out.push_back(Kernel::AllDR);
    else {
      std::cerr << "Unknown kernel string: " << part << std::endl;
      return false;
    }
  }
  return !out.empty();
}

// static flop counts of the reference's generated kernels (kernel::X::NonZeroFlops / HardwareFlops; neighboringFlux:
// mean over the 48 relations), elastic, per order
struct Flops { double derNz, derHw, volNz, volHw, lfNz, lfHw, nfNz, nfHw; };
Flops flopsOf(int order) {
  switch (order) {
    case 4: return {6642, 34146, 4806, 19584, 18972, 56448, 4 * 5073.0, 4 * 16992.0};
    case 5: return {20943, 98001, 13986, 56448, 40320, 102528, 4 * 10671.0, 4 * 29952.0};
    case 6: return {56286, 248292, 34839, 127872, 79200, 246528, 4 * 20805.0, 4 * 73728.0};
    default: return {0, 0, 0, 0, 0, 0, 0, 0};
  }
}

template <typename T>
void upload(void*& dev, const std::vector<T>& host) {
  checkCuda(cudaMalloc(&dev, host.size() * sizeof(T)), "cudaMalloc");
  checkCuda(cudaMemcpy(dev, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice), "cudaMemcpy");
}
template <typename T>
void dumpRaw(const std::string& path, const std::vector<T>& v) {
  std::ofstream f(path, std::ios::binary);
  f.write(reinterpret_cast<const char*>(v.data()), static_cast<std::streamsize>(v.size() * sizeof(T)));
}

template <typename real>
int run(std::size_t cells, std::size_t timesteps, const std::vector<Kernel>& kernels, const std::string& kernelStr, int order,
        bool json, const std::string& dump) {
  Context ctx(order, sizeof(real));
  const int B = ctx.basisFunctions(), ld = ctx.leadingDimension();
  const std::size_t qs = ctx.dofsSize();

  // ---- fake data (Allocator.cpp:49-134)
  std::mt19937 rng(static_cast<unsigned>(cells));
  std::uniform_int_distribution<unsigned> sideDist(0, 3), orientationDist(0, 2);
  std::uniform_int_distribution<std::size_t> cellDist(0, cells - 1);
  std::vector<std::uint8_t> faceTypes(4 * cells, 0), faceRel(8 * cells);
  std::vector<long long> neighbors(4 * cells);
  for (std::size_t c = 0; c < cells; ++c) {
    for (int f = 0; f < 4; ++f) {
      faceRel[8 * c + 2 * f] = static_cast<std::uint8_t>(sideDist(rng));
      faceRel[8 * c + 2 * f + 1] = static_cast<std::uint8_t>(orientationDist(rng));
      neighbors[4 * c + f] = static_cast<long long>(cellDist(rng));
    }
  }
  std::mt19937 vrng(42);
  std::uniform_real_distribution<double> u(-1.0, 1.0);
  const double scale = 1e-2;
  std::vector<real> dofs(cells * qs, real(0)), buffers(cells * qs, real(0)), localInt(cells * 396), neighInt(cells * 324);
  for (std::size_t c = 0; c < cells; ++c) {
    for (int q = 0; q < 9; ++q) {
      for (int k = 0; k < B; ++k) {
        dofs[c * qs + q * ld + k] = static_cast<real>(u(vrng));
      }
    }
  }
  for (auto& v : localInt) v = static_cast<real>(u(vrng) * scale);
  for (auto& v : neighInt) v = static_cast<real>(u(vrng) * scale);

  Layer layer;
  layer.size = cells;
  layer.realSize = sizeof(real);
  layer.dofsStride = qs;
  upload(layer.dofs, dofs);
  upload(layer.buffers, buffers);
  upload(layer.localIntegration, localInt);
  upload(layer.neighboringIntegration, neighInt);
  void* p = nullptr;
  upload(p, faceTypes); layer.faceTypes = static_cast<std::uint8_t*>(p);
  upload(p, faceRel); layer.faceRelations = static_cast<std::uint8_t*>(p);
  upload(p, neighbors); layer.faceNeighbors = static_cast<long long*>(p);
  // face projections (order 6: the neighbor kernel gathers 4 x 864 B per cell instead of 4 x 2016 B; switched off with
  // SSB_FACEPROJ=0).  Not for the *_dr kernels: their faces are all DynamicRupture.
  {
    const char* e = std::getenv("SSB_FACEPROJ");
    bool useFaceProj = e != nullptr ? e[0] != '0' : order >= 6;
    for (Kernel k : kernels) {
      useFaceProj = useFaceProj && (k == Kernel::All || k == Kernel::Local || k == Kernel::Neighbor || k == Kernel::Ader || k == Kernel::LocalWOAder);
    }
    if (useFaceProj) {
      const int b2p = (order * (order + 1) / 2 + 3) / 4 * 4;
      layer.faceProjectionsStride = static_cast<std::size_t>(4 * 9 * b2p);
      checkCuda(cudaMalloc(&layer.faceProjections, cells * layer.faceProjectionsStride * sizeof(real)), "cudaMalloc");
      checkCuda(cudaMemset(layer.faceProjections, 0, cells * layer.faceProjectionsStride * sizeof(real)), "cudaMemset");
    }
  }

  // processing order of the neighbor kernel: cells grouped by their four (neighbor face, orientation) pairs -- the
  // reference's recorder groups by the same key (NeighIntegrationRecorder.cpp:142-144)
  {
    std::vector<std::pair<int, int>> keyed(cells);
    for (std::size_t c = 0; c < cells; ++c) {
      int kj = 0, kh = 0;
      for (int f = 3; f >= 0; --f) {
        kj = kj * 4 + faceRel[8 * c + 2 * f];
        kh = kh * 3 + faceRel[8 * c + 2 * f + 1];
      }
      keyed[c] = {kj * 81 + kh, static_cast<int>(c)};
    }
    std::stable_sort(keyed.begin(), keyed.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    std::vector<int> perm(cells);
    for (std::size_t c = 0; c < cells; ++c) perm[c] = keyed[c].second;
    upload(p, perm);
    layer.permutation = static_cast<int*>(p);
  }

  // ---- fault-side fake data (src/Proxy/Allocator.cpp:193-294): 4 fault faces per cell, each picking random plus /
  //      minus cells' derivative families; every cell face maps to a random fault face.  Only for the *_dr kernels.
  bool needDr = false;
  for (Kernel k : kernels) {
    needDr = needDr || k == Kernel::NeighborDR || k == Kernel::GodunovDR || k == Kernel::AllDR;
  }
  kernels::FaultLayer fault;
  if (needDr) {
    const std::size_t faces = 4 * cells, ders = ctx.derivativesSize();
    const int pp = ctx.faultPointsPadded();
    fault.faces = faces;
    std::vector<real> derivs(cells * ders);
    std::uniform_real_distribution<double> u01(0.0, 1.0);
    for (auto& v : derivs) v = static_cast<real>(u01(vrng));
    void* dDer = nullptr;
    upload(dDer, derivs);
    layer.derivatives = dDer;
    layer.derivativesStride = ders;
    std::vector<const void*> plus(faces), minus(faces), god(4 * cells), fs(4 * cells);
    std::vector<std::uint8_t> info(3 * faces), sideRel(8 * cells), drTypes(4 * cells, 3);
    std::vector<real> tinv(81 * faces), fluxSolver(81 * faces), imposed(faces * 9 * pp);
    for (auto& v : tinv) v = static_cast<real>(u(vrng));
    for (auto& v : fluxSolver) v = static_cast<real>(u(vrng) * scale);
    for (auto& v : imposed) v = static_cast<real>(u(vrng) * scale);
    void *dTinv = nullptr, *dFs = nullptr, *dImp = nullptr;
    upload(dTinv, tinv);
    upload(dFs, fluxSolver);
    upload(dImp, imposed);
    std::uniform_int_distribution<std::size_t> faceDist(0, faces - 1);
    for (std::size_t fc = 0; fc < faces; ++fc) {
      plus[fc] = static_cast<const char*>(dDer) + cellDist(rng) * ders * sizeof(real);
      minus[fc] = static_cast<const char*>(dDer) + cellDist(rng) * ders * sizeof(real);
      info[3 * fc] = static_cast<std::uint8_t>(sideDist(rng));
      info[3 * fc + 1] = static_cast<std::uint8_t>(sideDist(rng));
      info[3 * fc + 2] = static_cast<std::uint8_t>(orientationDist(rng));
    }
    for (std::size_t c = 0; c < 4 * cells; ++c) {
      const std::size_t fc = faceDist(rng);
      god[c] = static_cast<const char*>(dImp) + fc * 9 * pp * sizeof(real);
      fs[c] = static_cast<const char*>(dFs) + fc * 81 * sizeof(real);
      sideRel[2 * c] = static_cast<std::uint8_t>(sideDist(rng));
      sideRel[2 * c + 1] = static_cast<std::uint8_t>(orientationDist(rng));
    }
    fault.tinvT = dTinv;
    upload(p, plus); fault.timeDerivativePlus = static_cast<const void* const*>(p);
    upload(p, minus); fault.timeDerivativeMinus = static_cast<const void* const*>(p);
    upload(p, god); fault.godunov = static_cast<const void* const*>(p);
    upload(p, fs); fault.fluxSolver = static_cast<const void* const*>(p);
    upload(p, info); fault.faceInformation = static_cast<std::uint8_t*>(p);
    upload(p, sideRel); fault.sideRelation = static_cast<std::uint8_t*>(p);
    upload(p, drTypes); fault.drFaceTypes = static_cast<std::uint8_t*>(p);
    const std::size_t qi = faces * order * 9 * pp;
    checkCuda(cudaMalloc(&fault.qInterpolatedPlus, qi * sizeof(real)), "cudaMalloc");
    checkCuda(cudaMalloc(&fault.qInterpolatedMinus, qi * sizeof(real)), "cudaMalloc");
    // fault faces grouped by their (plus side, minus side, relation) operator variant
    std::vector<std::pair<int, int>> keyed(faces);
    for (std::size_t fc = 0; fc < faces; ++fc) {
      keyed[fc] = {info[3 * fc] + 4 * (info[3 * fc + 1] + 4 * info[3 * fc + 2]), static_cast<int>(fc)};
    }
    std::stable_sort(keyed.begin(), keyed.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
    std::vector<int> fperm(faces);
    for (std::size_t fc = 0; fc < faces; ++fc) fperm[fc] = keyed[fc].second;
    upload(p, fperm);
    fault.permutation = static_cast<int*>(p);
  }

  kernels::Spacetime spacetime(ctx);
  kernels::Local local(ctx);
  kernels::Neighbor neighbor(ctx);
  kernels::DynamicRupture dynamicRupture(ctx);
  // Gauss-Legendre points on [0, dt] (src/Numerical/Quadrature.h:76-90) by Newton on the Legendre polynomial; row s of
  // the coefficient table = tau_s^d / d! (src/Numerical/TimeBasis.h:71-80)
  std::vector<double> drCoeffs(order * order);
  for (int s = 0; s < order; ++s) {
    double x = std::cos(M_PI * (s + 0.75) / (order + 0.5));
    for (int it = 0; it < 100; ++it) {
      double p0 = 1.0, p1 = x;
      for (int n = 2; n <= order; ++n) {
        const double pn = ((2.0 * n - 1.0) * x * p1 - (n - 1.0) * p0) / n;
        p0 = p1;
        p1 = pn;
      }
      const double dp = order * (x * p1 - p0) / (x * x - 1.0);
      const double dx = p1 / dp;
      x -= dx;
      if (std::fabs(dx) < 1e-15) break;
    }
    const double tau = 0.5 * 1.0 * (1.0 - x);   // points in ascending order on [0, 1]
    double f = 1.0;
    for (int d = 0; d < order; ++d) {
      if (d > 0) f *= d;
      drCoeffs[s * order + d] = std::pow(tau, d) / f;
    }
  }
  cudaStream_t stream;
  checkCuda(cudaStreamCreate(&stream), "cudaStreamCreate");
  const double dt = 1.0;  // src/Proxy/Constants.h
  std::vector<double> coeffs(order);
  double fact = 1.0;
  for (int d = 0; d < order; ++d) {
    fact *= (d + 1);
    coeffs[d] = std::pow(dt, d + 1) / fact;  // MonomialBasis::integrate (src/Numerical/TimeBasis.h:82-96)
  }
  // all_dr = local, godunov_dr, neigh_dr (src/Proxy/KernelHost.h:29-56); the fault faces replace the regular neighbors
  auto pass = [&]() {
    for (Kernel k : kernels) {
      if (k == Kernel::All || k == Kernel::Local || k == Kernel::Ader || k == Kernel::AllDR) spacetime.computeBatchedAder(coeffs.data(), dt, layer, stream);
      if (k == Kernel::All || k == Kernel::Local || k == Kernel::LocalWOAder || k == Kernel::AllDR) local.computeBatchedIntegral(layer, dt, stream);
      if (k == Kernel::All || k == Kernel::Neighbor) neighbor.computeBatchedNeighborsIntegral(layer, stream);
      if (k == Kernel::GodunovDR || k == Kernel::AllDR) dynamicRupture.batchedSpaceTimeInterpolation(fault, drCoeffs.data(), stream);
      if (k == Kernel::NeighborDR || k == Kernel::AllDR) dynamicRupture.computeBatchedNodalFlux(layer, fault, stream);
    }
  };
  pass();  // warm-up pass (Runner.cpp:93-95)
  checkCuda(cudaStreamSynchronize(stream), "sync");
  const auto t0 = std::chrono::steady_clock::now();
  for (std::size_t t = 0; t < timesteps; ++t) {
    pass();
  }
  checkCuda(cudaStreamSynchronize(stream), "sync");
  const double time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  // ---- output (Runner.cpp:113-151, Common.cpp:91-151)
  const Flops fl = flopsOf(order);
  double nz = 0, hw = 0, reals = 0;
  const double bq = double(B) * 9;
  for (Kernel k : kernels) {
    const bool a = k == Kernel::All || k == Kernel::Local || k == Kernel::Ader;
    const bool l = k == Kernel::All || k == Kernel::Local || k == Kernel::LocalWOAder;
    const bool n = k == Kernel::All || k == Kernel::Neighbor;
    if (a) { nz += fl.derNz; hw += fl.derHw; reals += 2 * bq + 72; }
    if (l) { nz += fl.volNz + fl.lfNz; hw += fl.volHw + fl.lfHw; reals += 3 * bq + 72 + 324; }
    if (n) { nz += fl.nfNz; hw += fl.nfHw; reals += 6 * bq + 324; }
    // dynamic rupture, per cell = 4 fault faces (order 6 counts of the generated kernels, SURVEY.md §8: interpolation
    // 2 x 346 113 per face, nodal flux 56 387 per face on average); all_dr additionally runs the local kernels
    if (k == Kernel::AllDR) { nz += fl.derNz + fl.volNz + fl.lfNz; hw += fl.derHw + fl.volHw + fl.lfHw; reals += 5 * bq + 144 + 324; }
    if ((k == Kernel::GodunovDR || k == Kernel::AllDR) && order == 6) { nz += 4 * 692226.0; }
    if ((k == Kernel::NeighborDR || k == Kernel::AllDR) && order == 6) { nz += 4 * 56387.0; }
  }
  const double total = double(cells) * double(timesteps);
  // `gib`: the reference's legacy estimate with this build's padded tensor sizes (Time.cpp:222-232, Local.cpp:508-520,
  // Neighbor.cpp:218-226); `gib-kernel`: the algorithmic bytes of SURVEY.md §8(d) (unpadded, what the kernels have to move)
  double legacy = 0;
  const double qsz = double(qs);
  for (Kernel k : kernels) {
    const bool a = k == Kernel::All || k == Kernel::Local || k == Kernel::Ader || k == Kernel::AllDR;
    const bool l = k == Kernel::All || k == Kernel::Local || k == Kernel::LocalWOAder || k == Kernel::AllDR;
    const bool n = k == Kernel::All || k == Kernel::Neighbor;
    if (a) legacy += 3 * qsz + 72;
    if (l) legacy += 72 + 4 * 81 + qsz;
    if (n) legacy += 6 * qsz + 4 * 81;
  }
  const double gnz = nz * total * 1e-9, ghw = hw * total * 1e-9;
  const double gib = legacy * sizeof(real) * total / (1024.0 * 1024.0 * 1024.0);
  const double gibKernel = reals * sizeof(real) * total / (1024.0 * 1024.0 * 1024.0);
  const double cycles = 0.0;  // the reference derives cycles from time and always reports 0 (src/Proxy/Tools.cpp:12-40)
  if (!json) {
    std::cout << "=================================================\n";
    std::cout << "===            PERFORMANCE SUMMARY            ===\n";
    std::cout << "=================================================\n";
    std::cout << "seissol proxy mode                  : " << kernelStr << '\n';
    std::cout << "time for seissol proxy              : " << time << '\n';
    std::cout << "cycles                              : " << cycles << '\n' << '\n';
    std::cout << "GFLOP (libxsmm)                     : " << 0.0 << '\n';
    std::cout << "GFLOP (pspamm)                      : " << 0.0 << '\n';
    std::cout << "GFLOP (libxsmm + pspamm)            : " << 0.0 << '\n';
    std::cout << "GFLOP (non-zero) for seissol proxy  : " << gnz << '\n';
    std::cout << "GFLOP (hardware) for seissol proxy  : " << ghw << '\n';
    std::cout << "GiB (estimate) for seissol proxy    : " << gib << '\n';
    std::cout << "GiB (kernels) for seissol proxy     : " << gibKernel << '\n' << '\n';
    std::cout << "FLOPS/cycle (non-zero)              : " << 0.0 << '\n';
    std::cout << "FLOPS/cycle (hardware)              : " << 0.0 << '\n';
    std::cout << "Bytes/cycle (estimate)              : " << 0.0 << '\n';
    std::cout << "Bytes/cycle (kernel)                : " << 0.0 << '\n' << '\n';
    std::cout << "GFLOPS (non-zero) for seissol proxy : " << gnz / time << '\n';
    std::cout << "GFLOPS (hardware) for seissol proxy : " << ghw / time << '\n';
    std::cout << "GiB/s (estimate) for seissol proxy  : " << gib / time << '\n';
    std::cout << "GiB/s (kernel) for seissol proxy    : " << gibKernel / time << '\n';
    std::cout << "element updates/s                   : " << total / time << '\n';
    std::cout << "=================================================\n\n";
  } else {
    // the 17 keys of the reference's JSON line in its order (src/Proxy/Common.cpp:135-151; with cycles = 0 the
    // reference's per-cycle fields are not finite, they are written as 0 here), then our addition
    std::cout << "{\"name\":\"" << kernelStr << "\",\"time\":" << time << ",\"cycles\":" << cycles
              << ",\"gflop-libxsmm\":0,\"gflop-pspamm\":0,\"gflop-nz\":" << gnz << ",\"gflop-hw\":" << ghw
              << ",\"gib\":" << gib << ",\"gib-kernel\":" << gibKernel
              << ",\"gflopcycle-nz\":0,\"gflopcycle-hw\":0,\"gibcycle\":0,\"gibcycle-kernel\":0"
              << ",\"gflops-nz\":" << gnz / time << ",\"gflops-hw\":" << ghw / time << ",\"gibs\":" << gib / time
              << ",\"gibs-kernel\":" << gibKernel / time << ",\"elem-updates-per-s\":" << total / time << "}\n";
  }
  if (!dump.empty()) {
    std::vector<real> out(cells * qs);
    checkCuda(cudaMemcpy(out.data(), layer.dofs, out.size() * sizeof(real), cudaMemcpyDeviceToHost), "cudaMemcpy");
    dumpRaw(dump + "_Q0.bin", dofs);
    dumpRaw(dump + "_Q.bin", out);
    dumpRaw(dump + "_localInt.bin", localInt);
    dumpRaw(dump + "_neighInt.bin", neighInt);
    dumpRaw(dump + "_faceRel.bin", faceRel);
    dumpRaw(dump + "_neighbors.bin", neighbors);
    std::ofstream meta(dump + "_meta.txt");
    meta << cells << ' ' << timesteps << ' ' << order << ' ' << sizeof(real) << ' ' << B << ' ' << ld << '\n';
  }
  return 0;
}

// viscoelastic2, order 5, 3 mechanisms (BASELINE.json configs[2]): same protocol on the anelastic kernel classes
template <typename real>
int runVisco(std::size_t cells, std::size_t timesteps, const std::vector<Kernel>& kernels, const std::string& kernelStr, bool json,
             const std::string& dump) {
  const int order = 5;
  Context ctx(order, sizeof(real), false, 1, 3);
  const int B = ctx.basisFunctions(), ld = ctx.leadingDimension();
  const std::size_t qs = ctx.dofsSize(), as = 18 * static_cast<std::size_t>(ld);
  std::mt19937 rng(static_cast<unsigned>(cells));
  std::uniform_int_distribution<unsigned> sideDist(0, 3), orientationDist(0, 2);
  std::uniform_int_distribution<std::size_t> cellDist(0, cells - 1);
  std::vector<std::uint8_t> faceTypes(4 * cells, 0), faceRel(8 * cells);
  std::vector<long long> neighbors(4 * cells);
  for (std::size_t c = 0; c < cells; ++c) {
    for (int f = 0; f < 4; ++f) {
      faceRel[8 * c + 2 * f] = static_cast<std::uint8_t>(sideDist(rng));
      faceRel[8 * c + 2 * f + 1] = static_cast<std::uint8_t>(orientationDist(rng));
      neighbors[4 * c + f] = static_cast<long long>(cellDist(rng));
    }
  }
  std::mt19937 vrng(43);
  std::uniform_real_distribution<double> u(-1.0, 1.0);
  const double scale = 1e-2;
  std::vector<real> dofs(cells * qs, real(0)), ane(cells * as, real(0)), localInt(cells * 807), neighInt(cells * 543);
  for (std::size_t c = 0; c < cells; ++c) {
    for (int q = 0; q < 9; ++q)
      for (int k = 0; k < B; ++k) dofs[c * qs + q * ld + k] = static_cast<real>(u(vrng));
    for (int q = 0; q < 18; ++q)
      for (int k = 0; k < B; ++k) ane[c * as + q * ld + k] = static_cast<real>(u(vrng));
  }
  for (auto& v : localInt) v = static_cast<real>(u(vrng) * scale);
  for (auto& v : neighInt) v = static_cast<real>(u(vrng) * scale);
  kernels::ViscoLayer layer;
  layer.size = cells;
  layer.realSize = sizeof(real);
  layer.dofsStride = qs;
  layer.aneStride = as;
  upload(layer.dofs, dofs);
  upload(layer.dofsAne, ane);
  std::vector<real> zq(cells * qs, real(0)), za(cells * as, real(0));
  upload(layer.buffers, zq);
  upload(layer.buffersAne, za);
  upload(layer.localIntegration, localInt);
  upload(layer.neighboringIntegration, neighInt);
  void* p = nullptr;
  upload(p, faceTypes); layer.faceTypes = static_cast<std::uint8_t*>(p);
  upload(p, faceRel); layer.faceRelations = static_cast<std::uint8_t*>(p);
  upload(p, neighbors); layer.faceNeighbors = static_cast<long long*>(p);
  kernels::anelastic::Spacetime spacetime(ctx);
  kernels::anelastic::Local local(ctx);
  kernels::anelastic::Neighbor neighbor(ctx);
  cudaStream_t stream;
  checkCuda(cudaStreamCreate(&stream), "cudaStreamCreate");
  std::vector<double> coeffs(order);
  double fact = 1.0;
  for (int d = 0; d < order; ++d) {
    fact *= (d + 1);
    coeffs[d] = 1.0 / fact;
  }
  auto pass = [&]() {
    for (Kernel k : kernels) {
      if (k == Kernel::All || k == Kernel::Local || k == Kernel::Ader) spacetime.computeBatchedAder(coeffs.data(), 1.0, layer, stream);
      if (k == Kernel::All || k == Kernel::Local || k == Kernel::LocalWOAder) local.computeBatchedIntegral(layer, 1.0, stream);
      if (k == Kernel::All || k == Kernel::Neighbor) neighbor.computeBatchedNeighborsIntegral(layer, stream);
    }
  };
  pass();
  checkCuda(cudaStreamSynchronize(stream), "sync");
  const auto t0 = std::chrono::steady_clock::now();
  for (std::size_t t = 0; t < timesteps; ++t) pass();
  checkCuda(cudaStreamSynchronize(stream), "sync");
  const double time = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  double nz = 0;   // kernel::X::NonZeroFlops of the generated viscoelastic2 kernels (order 5, 3 mechanisms)
  for (Kernel k : kernels) {
    if (k == Kernel::All || k == Kernel::Local || k == Kernel::Ader) nz += 120489;
    if (k == Kernel::All || k == Kernel::Local || k == Kernel::LocalWOAder) nz += 15924 + 67680 + 14175;
    if (k == Kernel::All || k == Kernel::Neighbor) nz += 4 * 17511.0 + 1575;
  }
  const double total = double(cells) * double(timesteps);
  if (json) {
    std::cout << "{\"name\":\"" << kernelStr << "\",\"equations\":\"viscoelastic2\",\"time\":" << time << ",\"cycles\":0,\"gflop-nz\":"
              << nz * total * 1e-9 << ",\"gflops-nz\":" << nz * total * 1e-9 / time << ",\"elem-updates-per-s\":" << total / time << "}\n";
  } else {
    std::cout << "seissol proxy mode                  : " << kernelStr << " (viscoelastic2)\n"
              << "time for seissol proxy              : " << time << "\n"
              << "GFLOPS (non-zero) for seissol proxy : " << nz * total * 1e-9 / time << "\n"
              << "element updates/s                   : " << total / time << "\n";
  }
  if (!dump.empty()) {
    std::vector<real> out(cells * qs), outA(cells * as);
    checkCuda(cudaMemcpy(out.data(), layer.dofs, out.size() * sizeof(real), cudaMemcpyDeviceToHost), "cudaMemcpy");
    checkCuda(cudaMemcpy(outA.data(), layer.dofsAne, outA.size() * sizeof(real), cudaMemcpyDeviceToHost), "cudaMemcpy");
    dumpRaw(dump + "_Q0.bin", dofs);
    dumpRaw(dump + "_Qane0.bin", ane);
    dumpRaw(dump + "_Q.bin", out);
    dumpRaw(dump + "_Qane.bin", outA);
    dumpRaw(dump + "_localInt.bin", localInt);
    dumpRaw(dump + "_neighInt.bin", neighInt);
    dumpRaw(dump + "_faceRel.bin", faceRel);
    dumpRaw(dump + "_neighbors.bin", neighbors);
    std::ofstream meta(dump + "_meta.txt");
    meta << cells << ' ' << timesteps << ' ' << order << ' ' << sizeof(real) << ' ' << B << ' ' << ld << '\n';
  }
  return 0;
}

}  // namespace

int main(int argc, char* argv[]) {
  std::vector<std::string> pos;
  bool json = false;
  int order = 6;
  std::string precision = "s", dump, equations = "elastic";
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    if ((a == "-f" || a == "--format") && i + 1 < argc) json = std::string(argv[++i]) == "json";
    else if (a == "--order" && i + 1 < argc) order = std::atoi(argv[++i]);
    else if (a == "--precision" && i + 1 < argc) precision = argv[++i];
    else if (a == "--dump" && i + 1 < argc) dump = argv[++i];
    else if (a == "--equations" && i + 1 < argc) equations = argv[++i];
    else pos.push_back(a);
  }
  if (pos.size() != 3) {
    std::cerr << "usage: ssb_proxy <cells> <timesteps> <kernel[,kernel]> [-f plain|json] [--order 4|5|6] [--precision s|d] "
                 "[--dump PREFIX] [--equations elastic|viscoelastic2]\nkernels: all, local, neigh, ader, localwoader, neigh_dr, godunov_dr, all_dr" << std::endl;
    return -1;
  }
  std::vector<Kernel> kernels;
  if (!str2kernel(pos[2], kernels)) {
    return -1;
  }
  try {
    const std::size_t cells = std::stoull(pos[0]), timesteps = std::stoull(pos[1]);
    if (equations == "viscoelastic2") {
      for (Kernel k : kernels) {
        if (k == Kernel::NeighborDR || k == Kernel::GodunovDR || k == Kernel::AllDR) {
          std::cerr << "ssb_proxy: the dynamic-rupture kernels are built for the elastic equations" << std::endl;
          return -1;
        }
      }
      return precision == "d" ? runVisco<double>(cells, timesteps, kernels, pos[2], json, dump)
                              : runVisco<float>(cells, timesteps, kernels, pos[2], json, dump);
    }
    if (precision == "d") {
      return run<double>(cells, timesteps, kernels, pos[2], order, json, dump);
    }
    return run<float>(cells, timesteps, kernels, pos[2], order, json, dump);
  } catch (const std::exception& e) {
    std::cerr << "ssb_proxy: " << e.what() << std::endl;
    return 1;
  }
}
