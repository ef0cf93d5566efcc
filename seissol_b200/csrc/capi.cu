// C ABI of libseissol_b200.so (declared in include/seissol_b200.h).  Thin argument
// validation + dispatch onto the per-(order, real) kernel instantiations.  No CPU path.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "../../include/seissol_b200.h"
#include "api_internal.h"
#include "kernels_setup.cuh"
#include "kernels_nodal.cuh"

using namespace ssb;


struct ssb_ctx {
  const KernelSet* ks;       // query + taylor + default
  const KernelSet* ksAder;   // predictor alone
  const KernelSet* ksLocal;  // volume + local flux (with or without the predictor)
  const KernelSet* ksNeigh;
  const DrKernelSet* dr;
  const ViscoKernelSet* visco;
  int ppad;
  int order;
  int realSize;
  int equation;
  int ldq;
  int derOffset[8];
  int derLd[8];
  int derSize;
};

namespace {

int basisFunctions(int order) { return order * (order + 1) * (order + 2) / 6; }

void compactLayout(ssb_ctx* c) {
  c->ldq = (basisFunctions(c->order) + 3) / 4 * 4;
  int off = 0;
  for (int d = 0; d < c->order; ++d) {
    const int ld = (basisFunctions(c->order - d) + 3) / 4 * 4;
    c->derOffset[d] = off;
    c->derLd[d] = ld;
    off += 9 * ld;
  }
  c->derSize = off;
}

void referenceLayout(ssb_ctx* c) {
  const int per = 64 / c->realSize;  // 64-byte aligned leading dimensions
  auto align = [&](int rows) { return (rows + per - 1) / per * per; };
  c->ldq = align(basisFunctions(c->order));
  int off = 0;
  for (int d = 0; d < c->order; ++d) {
    const int ld = align(basisFunctions(c->order - d));
    c->derOffset[d] = off;
    c->derLd[d] = ld;
    off += 9 * ld;
  }
  c->derSize = off;
}

// strided arrays: base and stride must be 16-byte aligned (bulk copies, 128-bit accesses); pointer tables cannot be
// checked on the host
bool aligned16(const ssb_array_t& a) {
  if (a.ptrs != nullptr || a.base == nullptr) {
    return true;
  }
  return reinterpret_cast<uintptr_t>(a.base) % 16 == 0 && a.stride_bytes % 16 == 0;
}

ElemArray conv(const ssb_array_t& a) {
  ElemArray r;
  r.ptrs = a.ptrs;
  r.base = static_cast<const char*>(a.base);
  r.strideBytes = a.stride_bytes;
  return r;
}

const DrKernelSet* lookupDr(int order, int realSize) {
  if (realSize == 4) {
    switch (order) {
      case 4: return drkernelset_o4_f32();
      case 5: return drkernelset_o5_f32();
      case 6: return drkernelset_o6_f32();
      default: return nullptr;
    }
  }
  if (realSize == 8) {
    switch (order) {
      case 4: return drkernelset_o4_f64();
      case 5: return drkernelset_o5_f64();
      case 6: return drkernelset_o6_f64();
      default: return nullptr;
    }
  }
  return nullptr;
}

// Kernel generation per entry point: 3 = full-column threads (kernels_elastic.cuh), 4 = half-column threads
// (kernels_elastic4.cuh).  Defaults from the measurements in DESIGN.md §5: single precision runs generation 3
// everywhere; double precision runs the predictor on generation 3 (255 registers, no spills) and
// volume/local flux/neighbor flux on generation 4 (generation 3 needs > 255 registers there).  The environment
// variable SSB_KGEN = 3 | 4 forces one generation for everything (A/B measurements only).
int forcedGeneration() {
  const char* e = std::getenv("SSB_KGEN");
  if (e != nullptr && (e[0] == '3' || e[0] == '4')) {
    return e[0] - '0';
  }
  return 0;
}

const KernelSet* lookup4(int order, int realSize) {
#ifdef SSB_HAVE_V4_F32
  if (realSize == 4) {
    switch (order) {
      case 4: return kernelset4_o4_f32();
      case 5: return kernelset4_o5_f32();
      case 6: return kernelset4_o6_f32();
      default: return nullptr;
    }
  }
#endif
  if (realSize == 8) {
    switch (order) {
      case 4: return kernelset4_o4_f64();
      case 5: return kernelset4_o5_f64();
      case 6: return kernelset4_o6_f64();
      default: return nullptr;
    }
  }
  return nullptr;
}

const KernelSet* lookup(int order, int realSize) {
  if (realSize == 4) {
    switch (order) {
      case 4: return kernelset_o4_f32();
      case 5: return kernelset_o5_f32();
      case 6: return kernelset_o6_f32();
      default: return nullptr;
    }
  }
  if (realSize == 8) {
    switch (order) {
      case 4: return kernelset_o4_f64();
      case 5: return kernelset_o5_f64();
      case 6: return kernelset_o6_f64();
      default: return nullptr;
    }
  }
  return nullptr;
}

int fillLocal(ssb_ctx* c, LocalParams& p, ssb_array_t dofs, ssb_array_t idofs, ssb_array_t derivs,
              ssb_array_t localInt, const uint8_t* faceTypes, const double* coeffs, long long n,
              const int* perm) {
  if (c == nullptr || n < 0) {
    return SSB_EINVAL;
  }
  std::memset(&p, 0, sizeof(p));
  if (n == 0) {
    return 0;  // empty batch: nothing to validate, the launcher is a no-op
  }
  p.dofs = conv(dofs);
  p.idofs = conv(idofs);
  p.derivs = conv(derivs);
  p.localInt = conv(localInt);
  if (!p.dofs.present() || !p.idofs.present() || !p.localInt.present()) {
    return SSB_EINVAL;
  }
  if (!aligned16(dofs) || !aligned16(idofs) || !aligned16(derivs) || !aligned16(localInt)) {
    return SSB_EINVAL;
  }
  p.faceTypes = faceTypes;
  p.perm = perm;
  p.n = n;
  p.ldg = c->ldq;
  for (int d = 0; d < c->order; ++d) {
    p.derOffset[d] = c->derOffset[d];
    p.derLd[d] = c->derLd[d];
    p.coeffs[d] = coeffs != nullptr ? coeffs[d] : 0.0;
  }
  return 0;
}

}  // namespace

extern "C" {

int ssb_create(ssb_ctx** out, int order, int real_size, int equation, int mechanisms) {
  if (out == nullptr) {
    return SSB_EINVAL;
  }
  *out = nullptr;
  const ViscoKernelSet* visco = nullptr;
  if (equation == 1) {
    // viscoelastic2: order 5 with 3 mechanisms is the compiled configuration (BASELINE.json configs[2])
    if (mechanisms != kVM || order != 5) {
      return SSB_EUNSUPPORTED;
    }
    visco = real_size == 4 ? viscokernelset_o5_f32() : (real_size == 8 ? viscokernelset_o5_f64() : nullptr);
    if (visco == nullptr) {
      return SSB_EUNSUPPORTED;
    }
  } else if (equation != 0 || mechanisms != 0) {
    return SSB_EUNSUPPORTED;
  }
  const KernelSet* ks = lookup(order, real_size);
  if (ks == nullptr) {
    return SSB_EUNSUPPORTED;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
    (void)cudaGetLastError();
    return SSB_ENODEVICE;
  }
  ssb_ctx* c = new (std::nothrow) ssb_ctx();
  if (c == nullptr) {
    return SSB_EINVAL;
  }
  const KernelSet* ks4 = lookup4(order, real_size);
  const int forced = forcedGeneration();
  c->ks = ks;
  c->ksAder = ks;
  c->ksLocal = (real_size == 8 && ks4 != nullptr) ? ks4 : ks;
  c->ksNeigh = c->ksLocal;
  if (const char* e = std::getenv("SSB_ADER_GEN")) {  // A/B measurements: predictor on generation 4
    if (e[0] == '4' && ks4 != nullptr) {
      c->ksAder = ks4;
    }
  }
  if (forced == 3 || (forced == 4 && ks4 != nullptr)) {
    c->ks = c->ksAder = c->ksLocal = c->ksNeigh = forced == 4 ? ks4 : ks;
  }
  c->visco = visco;
  c->dr = visco != nullptr ? nullptr : lookupDr(order, real_size);
  // fault points per face padded to 4 (viscoelastic2: only the nodal flux of the imposed state is on this path)
  c->ppad = c->dr != nullptr ? c->dr->PP : ((order + 1) * (order + 1) + 3) / 4 * 4;
  c->order = order;
  c->realSize = real_size;
  c->equation = equation;
  compactLayout(c);
  *out = c;
  return 0;
}

void ssb_destroy(ssb_ctx* c) { delete c; }

int ssb_set_layout(ssb_ctx* c, int ld_q, const int* der_offset, const int* der_ld) {
  if (c == nullptr || ld_q < c->ks->BP || ld_q % 4 != 0) {
    return SSB_EINVAL;
  }
  c->ldq = ld_q;
  if (der_offset != nullptr && der_ld != nullptr) {
    int end = 0;
    for (int d = 0; d < c->order; ++d) {
      if (der_ld[d] % 4 != 0 || der_offset[d] % 4 != 0 || der_ld[d] < (basisFunctions(c->order - d) + 3) / 4 * 4) {
        return SSB_EINVAL;
      }
      // the family is moved as ONE block [0, end): the dQ(d) must follow each other without overlap
      if (der_offset[d] < end) {
        return SSB_EINVAL;
      }
      c->derOffset[d] = der_offset[d];
      c->derLd[d] = der_ld[d];
      end = der_offset[d] + 9 * der_ld[d] > end ? der_offset[d] + 9 * der_ld[d] : end;
    }
    c->derSize = end;
  }
  return 0;
}

int ssb_layout_reference(ssb_ctx* c) {
  if (c == nullptr) {
    return SSB_EINVAL;
  }
  referenceLayout(c);
  const int per = 64 / c->realSize;
  const int points = c->dr != nullptr ? c->dr->P : (c->order + 1) * (c->order + 1);
  c->ppad = (points + per - 1) / per * per;
  return 0;
}

int ssb_query(const ssb_ctx* c, int* out) {
  if (c == nullptr || out == nullptr) {
    return SSB_EINVAL;
  }
  std::memset(out, 0, 24 * sizeof(int));
  out[0] = c->ks->B;
  out[1] = c->ldq;
  out[2] = 9 * c->ldq;
  out[3] = c->derSize;
  for (int d = 0; d < c->order; ++d) {
    out[4 + d] = c->derOffset[d];
    out[12 + d] = c->derLd[d];
  }
  out[20] = c->ksLocal->eLocal;
  out[21] = c->ksNeigh->eNeigh;
  out[22] = c->dr != nullptr ? c->dr->P : (c->order + 1) * (c->order + 1);
  out[23] = c->ppad;
  return 0;
}

int ssb_ader(ssb_ctx* c, ssb_array_t dofs, ssb_array_t idofs, ssb_array_t derivs, ssb_array_t local_int,
             const double* coeffs, long long n, const int* perm, void* stream) {
  LocalParams p;
  if (coeffs == nullptr) {
    return SSB_EINVAL;
  }
  const int rc = fillLocal(c, p, dofs, idofs, derivs, local_int, nullptr, coeffs, n, perm);
  if (rc != 0) {
    return rc;
  }
  return c->ksAder->local(p, true, false, static_cast<cudaStream_t>(stream));
}

int ssb_local(ssb_ctx* c, ssb_array_t dofs, ssb_array_t idofs, ssb_array_t local_int,
              const uint8_t* face_types, long long n, const int* perm, void* stream) {
  LocalParams p;
  ssb_array_t none = {nullptr, nullptr, 0};
  const int rc = fillLocal(c, p, dofs, idofs, none, local_int, face_types, nullptr, n, perm);
  if (rc != 0) {
    return rc;
  }
  return c->ksLocal->local(p, false, true, static_cast<cudaStream_t>(stream));
}

int ssb_ader_local(ssb_ctx* c, ssb_array_t dofs, ssb_array_t idofs, ssb_array_t derivs,
                   ssb_array_t local_int, const uint8_t* face_types, const double* coeffs, long long n,
                   const int* perm, void* stream) {
  LocalParams p;
  if (coeffs == nullptr) {
    return SSB_EINVAL;
  }
  const int rc = fillLocal(c, p, dofs, idofs, derivs, local_int, face_types, coeffs, n, perm);
  if (rc != 0) {
    return rc;
  }
  return c->ksLocal->local(p, true, true, static_cast<cudaStream_t>(stream));
}

int ssb_local_faceproj(ssb_ctx* c, ssb_array_t dofs, ssb_array_t idofs, ssb_array_t local_int, ssb_array_t face_proj,
                       const uint8_t* face_types, long long n, const int* perm, void* stream) {
  LocalParams p;
  ssb_array_t none = {nullptr, nullptr, 0};
  const int rc = fillLocal(c, p, dofs, idofs, none, local_int, face_types, nullptr, n, perm);
  if (rc != 0) {
    return rc;
  }
  p.faceProj = conv(face_proj);
  if (n > 0 && !p.faceProj.present()) {
    return SSB_EINVAL;
  }
  return c->ksLocal->local(p, false, true, static_cast<cudaStream_t>(stream));
}

int ssb_ader_local_faceproj(ssb_ctx* c, ssb_array_t dofs, ssb_array_t idofs, ssb_array_t derivs, ssb_array_t local_int,
                            ssb_array_t face_proj, const uint8_t* face_types, const double* coeffs, long long n,
                            const int* perm, void* stream) {
  LocalParams p;
  if (coeffs == nullptr) {
    return SSB_EINVAL;
  }
  const int rc = fillLocal(c, p, dofs, idofs, derivs, local_int, face_types, coeffs, n, perm);
  if (rc != 0) {
    return rc;
  }
  p.faceProj = conv(face_proj);
  if (n > 0 && !p.faceProj.present()) {
    return SSB_EINVAL;
  }
  return c->ksLocal->local(p, true, true, static_cast<cudaStream_t>(stream));
}

int ssb_neigh_faceproj(ssb_ctx* c, ssb_array_t dofs, const void* const* nb_ptrs, const long long* nb_idx,
                       ssb_array_t face_proj, ssb_array_t neigh_int, const uint8_t* face_types,
                       const uint8_t* face_rel, long long n, const int* perm, void* stream) {
  if (c == nullptr || n < 0 || face_rel == nullptr) {
    return SSB_EINVAL;
  }
  if (n == 0) {
    return 0;
  }
  NeighParams p;
  std::memset(&p, 0, sizeof(p));
  p.dofs = conv(dofs);
  p.neighInt = conv(neigh_int);
  p.faceProj = conv(face_proj);
  p.nbPtrs = nb_ptrs;
  p.nbIdx = nb_idx;
  if (!p.dofs.present() || !p.neighInt.present() ||
      (nb_ptrs == nullptr && (nb_idx == nullptr || !p.faceProj.present()))) {
    return SSB_EINVAL;
  }
  if (!aligned16(dofs) || !aligned16(face_proj) || !aligned16(neigh_int)) {
    return SSB_EINVAL;
  }
  p.useFaceProj = 1;
  p.faceTypes = face_types;
  p.faceRel = face_rel;
  p.perm = perm;
  p.n = n;
  p.ldg = c->ldq;
  // the face-projection kernel lives in the full-column translation units (both precisions)
  return c->ks->neigh(p, static_cast<cudaStream_t>(stream));
}

int ssb_taylor(ssb_ctx* c, ssb_array_t derivs, ssb_array_t idofs, const double* coeffs, long long n,
               void* stream) {
  if (c == nullptr || coeffs == nullptr || n < 0) {
    return SSB_EINVAL;
  }
  TaylorParams p;
  std::memset(&p, 0, sizeof(p));
  p.derivs = conv(derivs);
  p.idofs = conv(idofs);
  if (n == 0) {
    return 0;
  }
  if (!p.derivs.present() || !p.idofs.present() || !aligned16(derivs) || !aligned16(idofs)) {
    return SSB_EINVAL;
  }
  p.n = n;
  p.ldg = c->ldq;
  p.order = c->order;
  p.derSize = c->derSize;
  for (int d = 0; d < c->order; ++d) {
    p.rows[d] = basisFunctions(c->order - d);
    p.derOffset[d] = c->derOffset[d];
    p.derLd[d] = c->derLd[d];
    p.coeffs[d] = coeffs[d];
  }
  return c->ks->taylor(p, static_cast<cudaStream_t>(stream));
}

int ssb_taylor_faceproj(ssb_ctx* c, ssb_array_t derivs, ssb_array_t proj_out, const uint8_t* face_j,
                        const double* coeffs, long long n, void* stream) {
  if (c == nullptr || coeffs == nullptr || n < 0 || c->ks->taylorFp == nullptr) {
    return SSB_EINVAL;
  }
  if (n == 0) {
    return 0;
  }
  TaylorParams p;
  std::memset(&p, 0, sizeof(p));
  p.derivs = conv(derivs);
  p.idofs = conv(proj_out);
  p.faceJ = face_j;
  if (!p.derivs.present() || !p.idofs.present() || face_j == nullptr) {
    return SSB_EINVAL;
  }
  p.n = n;
  p.ldg = c->ldq;
  p.order = c->order;
  p.derSize = c->derSize;
  for (int d = 0; d < c->order; ++d) {
    p.rows[d] = basisFunctions(c->order - d);
    p.derOffset[d] = c->derOffset[d];
    p.derLd[d] = c->derLd[d];
    p.coeffs[d] = coeffs[d];
  }
  return c->ks->taylorFp(p, static_cast<cudaStream_t>(stream));
}

int ssb_accumulate_rows(ssb_ctx* c, ssb_array_t dst, ssb_array_t src, int reset, long long n, int reals_per_row,
                        void* stream) {
  if (c == nullptr || n < 0 || reals_per_row <= 0 || reals_per_row % 4 != 0) {
    return SSB_EINVAL;
  }
  if (n == 0) {
    return 0;
  }
  AccumulateParams p;
  std::memset(&p, 0, sizeof(p));
  p.buffers = conv(dst);
  p.idofs = conv(src);
  if (!p.buffers.present() || !p.idofs.present()) {
    return SSB_EINVAL;
  }
  p.n = n;
  p.vec4PerElem = reals_per_row / 4;
  p.reset = reset;
  return c->ks->accumulate(p, static_cast<cudaStream_t>(stream));
}

int ssb_lts_accumulate(ssb_ctx* c, ssb_array_t buffers, ssb_array_t idofs, int reset, long long n, void* stream) {
  if (c == nullptr || n < 0) {
    return SSB_EINVAL;
  }
  if (n == 0) {
    return 0;
  }
  AccumulateParams p;
  std::memset(&p, 0, sizeof(p));
  p.buffers = conv(buffers);
  p.idofs = conv(idofs);
  if (!p.buffers.present() || !p.idofs.present()) {
    return SSB_EINVAL;
  }
  p.n = n;
  p.vec4PerElem = 9 * c->ldq / 4;
  p.reset = reset;
  return c->ks->accumulate(p, static_cast<cudaStream_t>(stream));
}

int ssb_neigh(ssb_ctx* c, ssb_array_t dofs, const void* const* nb_ptrs, const long long* nb_idx,
              ssb_array_t idofs, ssb_array_t neigh_int, const uint8_t* face_types, const uint8_t* face_rel,
              long long n, const int* perm, void* stream) {
  if (c == nullptr || n < 0 || face_rel == nullptr) {
    return SSB_EINVAL;
  }
  NeighParams p;
  std::memset(&p, 0, sizeof(p));
  p.dofs = conv(dofs);
  p.neighInt = conv(neigh_int);
  p.idofs = conv(idofs);
  p.nbPtrs = nb_ptrs;
  p.nbIdx = nb_idx;
  if (n == 0) {
    return 0;
  }
  if (!p.dofs.present() || !p.neighInt.present() ||
      (nb_ptrs == nullptr && (nb_idx == nullptr || !p.idofs.present()))) {
    return SSB_EINVAL;
  }
  if (!aligned16(dofs) || !aligned16(idofs) || !aligned16(neigh_int)) {
    return SSB_EINVAL;
  }
  p.faceTypes = face_types;
  p.faceRel = face_rel;
  p.perm = perm;
  p.n = n;
  p.ldg = c->ldq;
  return c->ksNeigh->neigh(p, static_cast<cudaStream_t>(stream));
}

int ssb_set_dr_layout(ssb_ctx* c, int ppad) {
  if (c == nullptr || c->dr == nullptr || ppad < c->dr->PP || ppad % 4 != 0) {
    return SSB_EINVAL;
  }
  c->ppad = ppad;
  return 0;
}

int ssb_dr_interp(ssb_ctx* c, ssb_array_t der_plus, ssb_array_t der_minus, const void* tinv_t,
                  const uint8_t* face_info, void* q_plus, void* q_minus, const double* coeffs, long long n,
                  const int* perm, void* stream) {
  if (c == nullptr || c->dr == nullptr || n < 0 || coeffs == nullptr) {
    return SSB_EINVAL;
  }
  if (n == 0) {
    return 0;
  }
  DrInterpParams p;
  std::memset(&p, 0, sizeof(p));
  p.derPlus = conv(der_plus);
  p.derMinus = conv(der_minus);
  if (!p.derPlus.present() || !p.derMinus.present() || tinv_t == nullptr || face_info == nullptr ||
      q_plus == nullptr || q_minus == nullptr || !aligned16(der_plus) || !aligned16(der_minus)) {
    return SSB_EINVAL;
  }
  p.tinvT = tinv_t;
  p.faceInfo = face_info;
  p.qPlus = q_plus;
  p.qMinus = q_minus;
  p.perm = perm;
  p.n = n;
  p.ppad = c->ppad;
  p.derSize = c->derSize;
  for (int d = 0; d < c->order; ++d) {
    p.derOffset[d] = c->derOffset[d];
    p.derLd[d] = c->derLd[d];
    for (int s = 0; s < c->order; ++s) {
      p.coeffs[s][d] = coeffs[s * c->order + d];
    }
  }
  return c->dr->interp(p, static_cast<cudaStream_t>(stream));
}

namespace {
int friction(ssb_ctx* c, const ssb_friction_data_t* d, const ssb_friction_params_t* q, long long n, int law,
             void* stream) {
  if (c == nullptr || c->dr == nullptr || n < 0 || d == nullptr || q == nullptr) {
    return SSB_EINVAL;
  }
  if (n == 0) {
    return 0;
  }
  FrictionParams p;
  std::memset(&p, 0, sizeof(p));
  p.qPlus = d->q_plus; p.qMinus = d->q_minus; p.imposedPlus = d->imposed_plus; p.imposedMinus = d->imposed_minus;
  p.impAndEta = d->imp_and_eta; p.initialStress = d->initial_stress;
  p.mu = d->mu; p.slip1 = d->slip1; p.slip2 = d->slip2; p.slipRate1 = d->slip_rate1; p.slipRate2 = d->slip_rate2;
  p.slipRateMagnitude = d->slip_rate_magnitude; p.accumulatedSlipMagnitude = d->accumulated_slip_magnitude;
  p.traction1 = d->traction1; p.traction2 = d->traction2; p.ruptureTime = d->rupture_time;
  p.peakSlipRate = d->peak_slip_rate; p.dynStressTime = d->dyn_stress_time;
  p.ruptureTimePending = d->rupture_time_pending; p.dynStressTimePending = d->dyn_stress_time_pending;
  p.dC = d->d_c; p.muS = d->mu_s; p.muD = d->mu_d; p.cohesion = d->cohesion; p.forcedRuptureTime = d->forced_rupture_time;
  p.rsA = d->rs_a; p.rsSl0 = d->rs_sl0; p.rsF0 = d->rs_f0; p.rsMuW = d->rs_mu_w; p.rsB = d->rs_b; p.rsSrW = d->rs_sr_w;
  p.stateVariable = d->state_variable; p.convergenceInner = d->convergence_inner; p.convergenceOuter = d->convergence_outer;
  p.resampleT = d->resample_t;
  p.initialPressure = d->initial_pressure;
  p.nucleationStress = d->nucleation_stress;
  p.nucleationPressure = d->nucleation_pressure;
  p.energySlip = d->energy_slip; p.energyAccumulatedSlip = d->energy_accumulated_slip;
  p.frictionalEnergy = d->frictional_energy; p.spaceWeights = d->space_weights;
  p.doubledSurfaceArea = d->doubled_surface_area;
  p.energiesFromAcrossFaultVelocities = q->energies_from_across_fault_velocities;
  if (p.frictionalEnergy != nullptr && (p.energySlip == nullptr || p.energyAccumulatedSlip == nullptr ||
                                        p.spaceWeights == nullptr || p.doubledSurfaceArea == nullptr)) {
    return SSB_EINVAL;   // the energy output needs all of its arrays
  }
  p.nucleationCount = q->nucleation_count;
  if (p.nucleationCount < 0 || p.nucleationCount > 4 || (p.nucleationCount > 0 && p.nucleationStress == nullptr)) {
    return SSB_EINVAL;
  }
  for (int i = 0; i < 4; ++i) {
    p.nucleationT0[i] = q->nucleation_t0[i];
    p.nucleationS0[i] = q->nucleation_s0[i];
  }
  const void* common[] = {p.qPlus, p.qMinus, p.imposedPlus, p.imposedMinus, p.impAndEta, p.initialStress, p.mu,
                          p.slip1, p.slip2, p.slipRate1, p.slipRate2, p.slipRateMagnitude, p.accumulatedSlipMagnitude,
                          p.traction1, p.traction2, p.ruptureTime, p.peakSlipRate, p.dynStressTime,
                          p.ruptureTimePending, p.dynStressTimePending, p.resampleT};
  for (const void* x : common) {
    if (x == nullptr) {
      return SSB_EINVAL;
    }
  }
  if (law == 0) {
    if (p.dC == nullptr || p.muS == nullptr || p.muD == nullptr || p.cohesion == nullptr || p.forcedRuptureTime == nullptr) {
      return SSB_EINVAL;
    }
  } else {
    if (p.rsA == nullptr || p.rsSl0 == nullptr || p.rsF0 == nullptr || p.rsMuW == nullptr || p.rsB == nullptr ||
        p.rsSrW == nullptr || p.stateVariable == nullptr || p.convergenceInner == nullptr || p.convergenceOuter == nullptr) {
      return SSB_EINVAL;
    }
  }
  p.n = n;
  p.ppad = c->ppad;
  for (int i = 0; i < 8; ++i) {
    p.deltaT[i] = q->delta_t[i];
    p.timeWeights[i] = q->time_weights[i];
  }
  p.fullUpdateTime = q->full_update_time;
  p.etaPDamp = q->eta_p_damp;
  p.healingThreshold = q->healing_threshold;
  p.t0 = q->t0;
  p.rsSr0 = q->rs_sr0;
  p.rsSlipRateTolerance = q->rs_slip_rate_tolerance;
  p.rsStateTolerance = q->rs_state_tolerance;
  p.rsFloatResolutionExit = (c->realSize == 4 && q->rs_float_resolution_exit != 0) ? 1 : 0;
  p.rsMaxNumberSlipRateUpdates = q->rs_max_number_slip_rate_updates;
  p.rsNumberStateVariableUpdates = q->rs_number_state_variable_updates;
  return c->dr->friction(p, law, static_cast<cudaStream_t>(stream));
}
}  // namespace

int ssb_friction_lsw(ssb_ctx* c, const ssb_friction_data_t* d, const ssb_friction_params_t* q, long long n,
                     void* stream) {
  return friction(c, d, q, n, 0, stream);
}

int ssb_friction_rsfvw(ssb_ctx* c, const ssb_friction_data_t* d, const ssb_friction_params_t* q, long long n,
                       void* stream) {
  return friction(c, d, q, n, 1, stream);
}

int ssb_neigh_dr(ssb_ctx* c, ssb_array_t dofs, const void* const* godunov, const void* const* flux_solver,
                 const uint8_t* side_rel, const uint8_t* face_types, long long n, void* stream) {
  if (c == nullptr || c->dr == nullptr || n < 0) {
    return SSB_EINVAL;
  }
  if (n == 0) {
    return 0;
  }
  NeighDrParams p;
  std::memset(&p, 0, sizeof(p));
  p.dofs = conv(dofs);
  if (!p.dofs.present() || godunov == nullptr || flux_solver == nullptr || side_rel == nullptr || face_types == nullptr) {
    return SSB_EINVAL;
  }
  p.godunov = godunov;
  p.fluxSolver = flux_solver;
  p.sideRel = side_rel;
  p.faceTypes = face_types;
  p.n = n;
  p.ldg = c->ldq;
  p.ppad = c->ppad;
  return c->dr->neighDr(p, static_cast<cudaStream_t>(stream));
}

namespace {
int fillVisco(ssb_ctx* c, ViscoLocalParams& p, ssb_array_t dofs, ssb_array_t dofs_ane, ssb_array_t idofs,
              ssb_array_t idofs_ane, ssb_array_t local_int, const uint8_t* face_types, const double* coeffs, long long n) {
  if (c == nullptr || c->visco == nullptr || n < 0) {
    return SSB_EINVAL;
  }
  std::memset(&p, 0, sizeof(p));
  if (n == 0) {
    return 0;
  }
  p.dofs = conv(dofs);
  p.dofsAne = conv(dofs_ane);
  p.idofs = conv(idofs);
  p.idofsAne = conv(idofs_ane);
  p.localInt = conv(local_int);
  if (!p.dofs.present() || !p.dofsAne.present() || !p.idofs.present() || !p.idofsAne.present() || !p.localInt.present()) {
    return SSB_EINVAL;
  }
  p.faceTypes = face_types;
  p.n = n;
  p.ldg = c->ldq;
  for (int d = 0; d < c->order; ++d) {
    p.coeffs[d] = coeffs != nullptr ? coeffs[d] : 0.0;
  }
  return 0;
}
}  // namespace

int ssb_visco_ader(ssb_ctx* c, ssb_array_t dofs, ssb_array_t dofs_ane, ssb_array_t idofs, ssb_array_t idofs_ane,
                   ssb_array_t local_int, const double* coeffs, long long n, void* stream) {
  if (coeffs == nullptr) {
    return SSB_EINVAL;
  }
  ViscoLocalParams p;
  const int rc = fillVisco(c, p, dofs, dofs_ane, idofs, idofs_ane, local_int, nullptr, coeffs, n);
  if (rc != 0 || n == 0) {
    return rc;
  }
  return c->visco->ader(p, static_cast<cudaStream_t>(stream));
}

int ssb_visco_local(ssb_ctx* c, ssb_array_t dofs, ssb_array_t dofs_ane, ssb_array_t idofs, ssb_array_t idofs_ane,
                    ssb_array_t local_int, const uint8_t* face_types, long long n, void* stream) {
  ViscoLocalParams p;
  const int rc = fillVisco(c, p, dofs, dofs_ane, idofs, idofs_ane, local_int, face_types, nullptr, n);
  if (rc != 0 || n == 0) {
    return rc;
  }
  return c->visco->local(p, static_cast<cudaStream_t>(stream));
}

int ssb_visco_neigh(ssb_ctx* c, ssb_array_t dofs, ssb_array_t dofs_ane, const void* const* nb_ptrs,
                    const long long* nb_idx, ssb_array_t idofs, ssb_array_t neigh_int, const uint8_t* face_types,
                    const uint8_t* face_rel, long long n, const int* perm, void* stream) {
  if (c == nullptr || c->visco == nullptr || n < 0 || face_rel == nullptr) {
    return SSB_EINVAL;
  }
  if (n == 0) {
    return 0;
  }
  ViscoNeighParams p;
  std::memset(&p, 0, sizeof(p));
  p.dofs = conv(dofs);
  p.dofsAne = conv(dofs_ane);
  p.idofs = conv(idofs);
  p.neighInt = conv(neigh_int);
  p.nbPtrs = nb_ptrs;
  p.nbIdx = nb_idx;
  if (!p.dofs.present() || !p.dofsAne.present() || !p.neighInt.present() ||
      (nb_ptrs == nullptr && (nb_idx == nullptr || !p.idofs.present()))) {
    return SSB_EINVAL;
  }
  p.faceTypes = face_types;
  p.faceRel = face_rel;
  p.perm = perm;
  p.n = n;
  p.ldg = c->ldq;
  return c->visco->neigh(p, static_cast<cudaStream_t>(stream));
}

int ssb_visco_neigh_dr(ssb_ctx* c, ssb_array_t dofs, ssb_array_t dofs_ane, ssb_array_t neigh_int,
                       const void* const* godunov, const void* const* flux_solver, const uint8_t* side_rel,
                       const uint8_t* face_types, long long n, void* stream) {
  if (c == nullptr || c->visco == nullptr || n < 0) {
    return SSB_EINVAL;
  }
  if (n == 0) {
    return 0;
  }
  if (godunov == nullptr || flux_solver == nullptr || side_rel == nullptr || face_types == nullptr) {
    return SSB_EINVAL;
  }
  ViscoNeighDrParams p;
  std::memset(&p, 0, sizeof(p));
  p.dofs = conv(dofs);
  p.dofsAne = conv(dofs_ane);
  p.neighInt = conv(neigh_int);
  if (!p.dofs.present() || !p.dofsAne.present() || !p.neighInt.present()) {
    return SSB_EINVAL;
  }
  p.godunov = godunov;
  p.fluxSolver = flux_solver;
  p.sideRel = side_rel;
  p.faceTypes = face_types;
  p.n = n;
  p.ldg = c->ldq;
  p.ppad = c->ppad;
  return c->visco->neighDr(p, static_cast<cudaStream_t>(stream));
}

int ssb_local_nodal_bc(ssb_ctx* c, ssb_array_t dofs, ssb_array_t idofs, ssb_array_t neigh_int, const long long* cell,
                       const uint8_t* face, const uint8_t* kind, const void* tinv, const void* easi_map,
                       const void* easi_const, const void* displacement, const void* rho_g, long long m, void* stream) {
  if (c == nullptr || m < 0 || c->equation != 0 || c->ks->nodalBc == nullptr) {
    return SSB_EINVAL;
  }
  if (m == 0) {
    return 0;
  }
  NodalBcParams p;
  std::memset(&p, 0, sizeof(p));
  p.dofs = conv(dofs); p.idofs = conv(idofs); p.neighInt = conv(neigh_int);
  if (!p.dofs.present() || !p.idofs.present() || !p.neighInt.present() || cell == nullptr || face == nullptr ||
      kind == nullptr || tinv == nullptr) {
    return SSB_EINVAL;
  }
  p.cell = cell; p.face = face; p.kind = kind; p.tinv = tinv;
  p.easiMap = easi_map; p.easiConst = easi_const; p.displacement = displacement; p.rhoG = rho_g;
  p.m = m;
  p.ldg = c->ldq;
  return c->ks->nodalBc(p, static_cast<cudaStream_t>(stream));
}

int ssb_local_nodal_flux(ssb_ctx* c, ssb_array_t dofs, ssb_array_t neigh_int, const long long* cell, const uint8_t* face,
                         const void* inodal, long long m, void* stream) {
  if (c == nullptr || m < 0 || c->equation != 0 || c->ks->nodalBc == nullptr) {
    return SSB_EINVAL;
  }
  if (m == 0) {
    return 0;
  }
  NodalBcParams p;
  std::memset(&p, 0, sizeof(p));
  p.dofs = conv(dofs); p.neighInt = conv(neigh_int);
  if (!p.dofs.present() || !p.neighInt.present() || cell == nullptr || face == nullptr || inodal == nullptr) {
    return SSB_EINVAL;
  }
  p.cell = cell; p.face = face; p.kind = nullptr; p.analytical = inodal;
  p.m = m;
  p.ldg = c->ldq;
  return c->ks->nodalBc(p, static_cast<cudaStream_t>(stream));
}

int ssb_volume_energies(ssb_ctx* c, ssb_array_t dofs, const double* volume, const double* material, long long n,
                        double* out, void* stream) {
  if (c == nullptr || n < 0 || c->equation != 0 || c->ks->volumeEnergies == nullptr) {
    return SSB_EINVAL;
  }
  if (n == 0) {
    return 0;
  }
  VolumeEnergyParams p;
  std::memset(&p, 0, sizeof(p));
  p.dofs = conv(dofs);
  if (!p.dofs.present() || volume == nullptr || material == nullptr || out == nullptr) {
    return SSB_EINVAL;
  }
  p.volume = volume; p.material = material; p.out = out;
  p.n = n;
  p.ldg = c->ldq;
  return c->ks->volumeEnergies(p, static_cast<cudaStream_t>(stream));
}

int ssb_fault_energies(ssb_ctx* c, const void* frictional_energy, const void* accumulated_slip,
                       const void* doubled_surface_area, const double* material_plus, const double* material_minus,
                       const uint8_t* plus_side_on_this_rank, long long m, double* out, void* stream) {
  if (c == nullptr || m < 0 || c->equation != 0) {
    return SSB_EINVAL;
  }
  if (m == 0) {
    return 0;
  }
  if (frictional_energy == nullptr || accumulated_slip == nullptr || doubled_surface_area == nullptr ||
      material_plus == nullptr || material_minus == nullptr || out == nullptr) {
    return SSB_EINVAL;
  }
  FaultEnergyParams p;
  std::memset(&p, 0, sizeof(p));
  p.frictionalEnergy = frictional_energy; p.accumulatedSlip = accumulated_slip;
  p.doubledSurfaceArea = doubled_surface_area;
  p.materialPlus = material_plus; p.materialMinus = material_minus;
  p.plusSideOnThisRank = plus_side_on_this_rank;
  p.out = out;
  p.m = m;
  p.points = (c->order + 1) * (c->order + 1);
  p.ppad = c->ppad;
  return faultEnergiesEntry(p, c->realSize, static_cast<cudaStream_t>(stream));
}

int ssb_setup_cell_matrices(ssb_ctx* c, const double* coords, const long long* neighbors, const double* material,
                            const uint8_t* face_types, ssb_array_t local_int, ssb_array_t neigh_int, long long n, void* stream) {
  if (c == nullptr || n < 0 || c->equation != 0) {
    return SSB_EINVAL;
  }
  if (n == 0) {
    return 0;
  }
  SetupCellParams p;
  std::memset(&p, 0, sizeof(p));
  p.coords = coords; p.neighbors = neighbors; p.material = material; p.faceTypes = face_types;
  p.localInt = conv(local_int); p.neighInt = conv(neigh_int);
  p.n = n;
  if (coords == nullptr || neighbors == nullptr || material == nullptr || !p.localInt.present() || !p.neighInt.present()) {
    return SSB_EINVAL;
  }
  return setupCellEntry(p, c->realSize, static_cast<cudaStream_t>(stream));
}

int ssb_setup_dr_matrices(ssb_ctx* c, const double* coords_plus, const double* coords_minus, const uint8_t* face_info,
                          const double* material_plus, const double* material_minus, void* tinv_t, void* imp_and_eta,
                          void* flux_solver_plus, void* flux_solver_minus, void* doubled_surface_area, long long m,
                          void* stream) {
  if (c == nullptr || m < 0 || c->equation != 0) {
    return SSB_EINVAL;
  }
  if (m == 0) {
    return 0;
  }
  if (coords_plus == nullptr || coords_minus == nullptr || face_info == nullptr || material_plus == nullptr ||
      material_minus == nullptr || tinv_t == nullptr || imp_and_eta == nullptr || flux_solver_plus == nullptr ||
      flux_solver_minus == nullptr) {
    return SSB_EINVAL;
  }
  SetupDrParams p;
  std::memset(&p, 0, sizeof(p));
  p.coordsPlus = coords_plus; p.coordsMinus = coords_minus; p.faceInfo = face_info;
  p.materialPlus = material_plus; p.materialMinus = material_minus;
  p.tinvT = tinv_t; p.impAndEta = imp_and_eta; p.fluxSolverPlus = flux_solver_plus; p.fluxSolverMinus = flux_solver_minus;
  p.doubledSurfaceArea = doubled_surface_area;
  p.m = m;
  return setupDrEntry(p, c->realSize, static_cast<cudaStream_t>(stream));
}

const char* ssb_error_string(int code) {
  switch (code) {
    case 0: return "success";
    case SSB_EINVAL: return "invalid argument";
    case SSB_EUNSUPPORTED: return "configuration (order, precision, equation) not compiled in";
    case SSB_ENODEVICE: return "no CUDA device available (there is no CPU fallback)";
    default: return code > 0 ? cudaGetErrorString(static_cast<cudaError_t>(code)) : "unknown error";
  }
}

const char* ssb_version(void) { return "seissol_b200 0.1 (sm_100a)"; }
}
