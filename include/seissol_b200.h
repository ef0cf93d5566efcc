/* seissol_b200 — C ABI of the B200-native ADER-DG element-update path.
 *
 * This is the drop-in boundary: a C-ABI shared library (libseissol_b200.so) whose
 * entry points are what SeisSol's device back-end binds for the element update.
 * Every function replaces one batched entry point of the reference (file:line cited
 * per function; all paths relative to the reference tree).  All data pointers are
 * DEVICE pointers; the caller owns all memory; nothing is allocated on the hot path;
 * all launches are stream-ordered and CUDA-graph-capture safe (no syncs, no
 * allocation), matching how the reference enqueues work on StreamRuntime::stream()
 * (src/Parallel/Runtime/Stream.h:163-190).
 *
 * Devices and threads: an entry point launches on the CURRENT CUDA device of the calling thread (the device the
 * stream and all pointers must belong to); one-time kernel attributes are set per device on first use, so one process
 * may drive several GPUs.  A context holds no device state and may be shared by host threads; calls on one stream
 * must not be issued concurrently.
 * Alignment: every per-element block (base + i*stride, or pointer-table entry) must be 16-byte aligned -- the kernels
 * move them with cp.async.bulk and 128-bit accesses; misaligned strided arrays are rejected with SSB_EINVAL, pointer
 * table entries are the caller's responsibility.  face_types rows are read as 4-byte, face_rel rows as 8-byte words.
 *
 * Return value: 0 on success, a positive cudaError_t code from the launch, or a
 * negative SSB_E* code for invalid arguments.  There is NO CPU fallback: without a
 * usable CUDA device every compute entry point fails.
 */
#ifndef SEISSOL_B200_H
#define SEISSOL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SSB_EINVAL (-1)      /* bad argument */
#define SSB_EUNSUPPORTED (-2) /* (order, precision, equation) not compiled in */
#define SSB_ENODEVICE (-3)   /* no CUDA device / driver */

/* One logical per-element array.  Either `ptrs` (device table of device pointers:
 * the reference's `real**` batched tables, src/Initializer/BatchRecorders/DataTypes/
 * Table.h:25-61, ConditionalKey.h:16-33) or `base` + index*stride_bytes. */
typedef struct ssb_array {
  const void* const* ptrs;
  const void* base;
  long long stride_bytes;
} ssb_array_t;

typedef struct ssb_ctx ssb_ctx;

/* One context per (ConvergenceOrder, real, equation) — the reference compiles one
 * binary per such configuration (cmake/process_users_input.cmake:53-92).
 * real_size: 4 or 8.  equation: 0 = elastic, 1 = viscoelastic2 (mechanisms = 3). */
int ssb_create(ssb_ctx** ctx, int order, int real_size, int equation, int mechanisms);
void ssb_destroy(ssb_ctx* ctx);

/* Memory layout of the modal tensors in global memory (column-major).
 * ld_q: leading dimension of Q/I columns (reals); der_offset[d], der_ld[d]: placement of
 * dQ(d) inside the derivative family.  Default (never called): compact layout, every
 * column padded to a multiple of 4 reals.  ssb_layout_reference() selects the reference's
 * 64-byte aligned layout (yateto/arch.py:197-198). */
int ssb_set_layout(ssb_ctx* ctx, int ld_q, const int* der_offset, const int* der_ld);
int ssb_layout_reference(ssb_ctx* ctx);
/* out[0]=#basis functions B, [1]=ld_q, [2]=Q size (reals), [3]=derivative family size,
 * [4..4+O) = der_offset, [12..12+O) = der_ld, [20] = elements per CTA (local), [21] = (neigh),
 * [22] = fault points P, [23] = ppad */
int ssb_query(const ssb_ctx* ctx, int* out24);

/* ader: Cauchy-Kovalevskaya predictor.  Replaces Spacetime::computeBatchedAder
 * (src/Kernels/LinearCK/Time.cpp:142-217; per-cell twin computeAder :65-140).
 * dofs: Q [in]; idofs: time-integrated I [out]; derivs: optional derivative family [out]
 * (pass a zeroed ssb_array_t to skip); local_int: LocalIntegrationData (star[3][24],
 * nApNm1[4][81], src/Initializer/Typedefs.h:187-196); coeffs: HOST array of `order`
 * Taylor integration coefficients (src/Numerical/TimeBasis.h:82-96).
 * perm: optional device array giving the processing order (may be NULL). */
int ssb_ader(ssb_ctx* ctx, ssb_array_t dofs, ssb_array_t idofs, ssb_array_t derivs,
             ssb_array_t local_int, const double* coeffs, long long n, const int* perm,
             void* stream);

/* local (without ader): volume + local flux.  Replaces Local::computeBatchedIntegral
 * (src/Kernels/LinearCK/Local.cpp:255-395; per-cell computeIntegral :124-253).
 * face_types: device uint8 [n][4] FaceType values (src/Initializer/BasicTypedefs.h),
 * DynamicRupture (3) faces get no local flux; may be NULL (all regular). */
int ssb_local(ssb_ctx* ctx, ssb_array_t dofs, ssb_array_t idofs, ssb_array_t local_int,
              const uint8_t* face_types, long long n, const int* perm, void* stream);

/* ader + local fused (the proxy's `local` kernel, src/Proxy/KernelDevice.cpp:69-88). */
int ssb_ader_local(ssb_ctx* ctx, ssb_array_t dofs, ssb_array_t idofs, ssb_array_t derivs,
                   ssb_array_t local_int, const uint8_t* face_types, const double* coeffs,
                   long long n, const int* perm, void* stream);

/* Taylor sum I = sum_d c_d dQ(d).  Replaces Time::evaluateBatched
 * (src/Kernels/LinearCK/Time.cpp:257-292) and TimeCommon::computeBatchedIntegrals
 * (src/Kernels/TimeCommon.cpp:89-125). */
int ssb_taylor(ssb_ctx* ctx, ssb_array_t derivs, ssb_array_t idofs, const double* coeffs,
               long long n, void* stream);

/* LTS buffer accumulation: buffers = reset ? idofs : buffers + idofs, for cells whose neighbors in a larger time
 * cluster read the sum over two sub-steps (LtsSetup::accumulateBuffers, src/Initializer/LtsSetup.h:21-111).  Replaces
 * device.algorithms.streamBatchedData / accumulateBatchedData in TimeCluster::computeLocalIntegrationDevice
 * (src/Solver/TimeStepping/TimeCluster.cpp:508-530). */
int ssb_lts_accumulate(ssb_ctx* ctx, ssb_array_t buffers, ssb_array_t idofs, int reset, long long n, void* stream);

/* Neighbor flux, all four faces and all 48 face relations in one launch.  Replaces
 * Neighbor::computeBatchedNeighborsIntegral (src/Kernels/LinearCK/Neighbor.cpp:111-188).
 * Neighbor data either as nb_ptrs (device [n][4] pointers to the neighbors' I,
 * LTS::FaceNeighborsDevice) or nb_idx (device [n][4] indices into `idofs`).
 * neigh_int: NeighboringIntegrationData (nAmNm1[4][81]); face_rel: device uint8 [n][4][2] =
 * {neighbor's local face id, orientation} (CellLocalInformation.h:26-27). */
int ssb_neigh(ssb_ctx* ctx, ssb_array_t dofs, const void* const* nb_ptrs, const long long* nb_idx,
              ssb_array_t idofs, ssb_array_t neigh_int, const uint8_t* face_types,
              const uint8_t* face_rel, long long n, const int* perm, void* stream);

/* Face-projection path of the neighbor flux (GTS, or LTS neighbors of the same cluster).
 * The face mass matrix of the reference element is diagonal and the same on all four faces (fMrT_f = D rT_f,
 * codegen/matrices/matrices_*.xml), so the local kernel already holds F_f = fMrT_f I = D (rT_f I) for every face f of
 * its element while it evaluates the local flux (codegen/kernels/aderdg.py:373-385).  ssb_local_faceproj /
 * ssb_ader_local_faceproj additionally store these four [9][B2P] blocks per element (face_proj: 4*9*B2P reals per
 * element, B2P = O(O+1)/2 padded to 4; entries of faces with FaceType DynamicRupture are written as well), and
 * ssb_neigh_faceproj evaluates  Q_i += rDivM_i (fP_h D^-1) (F_j,nb AminusT_i)  from the block of face j of each
 * neighbor: 4 x 864 B gathered per element instead of 4 x 2016 B (order 6, float) and no rT_j products.
 * Same result as ssb_local + ssb_neigh (src/Kernels/LinearCK/Neighbor.cpp:111-188) up to rounding.
 * nb_idx: device [n][4] indices into face_proj; or nb_ptrs: device [n][4] pointers to the NEIGHBOR's face_proj block
 * (NULL entry = face contributes nothing).  The pointers must be 16-byte aligned. */
int ssb_local_faceproj(ssb_ctx* ctx, ssb_array_t dofs, ssb_array_t idofs, ssb_array_t local_int,
                       ssb_array_t face_proj, const uint8_t* face_types, long long n, const int* perm,
                       void* stream);
int ssb_ader_local_faceproj(ssb_ctx* ctx, ssb_array_t dofs, ssb_array_t idofs, ssb_array_t derivs,
                            ssb_array_t local_int, ssb_array_t face_proj, const uint8_t* face_types,
                            const double* coeffs, long long n, const int* perm, void* stream);
int ssb_neigh_faceproj(ssb_ctx* ctx, ssb_array_t dofs, const void* const* nb_ptrs, const long long* nb_idx,
                       ssb_array_t face_proj, ssb_array_t neigh_int, const uint8_t* face_types,
                       const uint8_t* face_rel, long long n, const int* perm, void* stream);

/* LTS with face projections.  ssb_taylor_faceproj: for a neighbor that is read through its derivative family
 * (LtsSetup::neighborHasDerivatives, src/Initializer/LtsSetup.h:21-111; TimeCommon::computeBatchedIntegrals,
 * src/Kernels/TimeCommon.cpp:89-125) integrate the family over the reader's interval (coeffs as ssb_taylor) and store
 * F = fMrT_j I of the one face j (face_j: device uint8 [n]) the reader sees: proj_out entries are [9][B2P] blocks.  The
 * neighbor table of ssb_neigh_faceproj then holds (block pointer - j * 9 * B2P reals) for such faces.
 * ssb_accumulate_rows: dst = reset ? src : dst + src on rows of reals_per_row reals (multiple of 4) -- the buffer
 * accumulation of ssb_lts_accumulate for face-projection rows. */
int ssb_taylor_faceproj(ssb_ctx* ctx, ssb_array_t derivs, ssb_array_t proj_out, const uint8_t* face_j,
                        const double* coeffs, long long n, void* stream);
int ssb_accumulate_rows(ssb_ctx* ctx, ssb_array_t dst, ssb_array_t src, int reset, long long n, int reals_per_row,
                        void* stream);

/* ------------------------------------------------------------------------------------------
 * Dynamic rupture.  Fault tensors use the reference's layout: point index fastest, padded to
 * `ppad` points (default: multiple of 4; ssb_layout_reference: 64-byte aligned like
 * dr::misc::NumPaddedPoints, src/DynamicRupture/Misc.h:63-66):
 * QInterpolated{Plus,Minus}[face][O][9][ppad], ImposedState{Plus,Minus}[face][9][ppad],
 * every per-point array [face][ppad], bool arrays as uint8 [face][ppad]. */
int ssb_set_dr_layout(ssb_ctx* ctx, int ppad);

/* godunov_dr: space-time interpolation at the fault.  Replaces
 * kernels::DynamicRupture::batchedSpaceTimeInterpolation (src/Kernels/DynamicRupture.cpp:110-168).
 * der_plus/der_minus: per FACE the derivative family of the plus / minus element
 * (DynamicRupture::TimeDerivative{Plus,Minus}Device); tinv_t: device [n][81] (DRGodunovData::dataTinvT);
 * face_info: device uint8 [n][3] = DRFaceInformation{plusSide, minusSide, faceRelation};
 * coeffs: HOST [O][O], row s = point-evaluation coefficients tau_s^d/d! of time point s
 * (src/Numerical/TimeBasis.h:71-80 at the Gauss-Legendre points, DynamicRupture.cpp:62-70). */
int ssb_dr_interp(ssb_ctx* ctx, ssb_array_t der_plus, ssb_array_t der_minus, const void* tinv_t,
                  const uint8_t* face_info, void* q_plus, void* q_minus, const double* coeffs,
                  long long n_faces, const int* perm, void* stream);

/* FrictionLawData as a structure of DEVICE arrays (src/Memory/Descriptor/DynamicRupture.h:47-265). */
typedef struct ssb_friction_data {
  const void* q_plus; const void* q_minus;            /* QInterpolatedPlus/Minus */
  void* imposed_plus; void* imposed_minus;            /* ImposedStatePlus/Minus [n][9][ppad] */
  const void* imp_and_eta;                            /* [n][11] ImpedancesAndEta (DynamicRupture/Typedefs.h:23-35) */
  const void* initial_stress;                         /* InitialStressInFaultCS [n][6][ppad]; written when nucleating */
  void* mu; void* slip1; void* slip2; void* slip_rate1; void* slip_rate2; void* slip_rate_magnitude;
  void* accumulated_slip_magnitude; void* traction1; void* traction2; void* rupture_time;
  void* peak_slip_rate; void* dyn_stress_time;
  uint8_t* rupture_time_pending; uint8_t* dyn_stress_time_pending;
  const void* d_c; const void* mu_s; const void* mu_d; const void* cohesion; const void* forced_rupture_time; /* LSW */
  const void* rs_a; const void* rs_sl0; const void* rs_f0; const void* rs_mu_w; const void* rs_b; const void* rs_sr_w;
  void* state_variable; uint8_t* convergence_inner; uint8_t* convergence_outer;                      /* RS */
  const void* resample_t;                             /* [P][P], resample_t[j*P+p] = init::resample[p][j] */
  /* nucleation + initial pressure (all optional, NULL = absent).  With nucleation_count > 0 in the parameters the kernel
   * applies common::adjustInitialStress (FrictionSolverCommon.h:503-531) before every time point like
   * BaseFrictionLaw::evaluate (CpuImpl/BaseFrictionLaw.h:106-118): initial_stress and initial_pressure are then UPDATED
   * in place.  initial_pressure enters the normal stress of both laws (CpuImpl/LinearSlipWeakening.h:147-156,
   * CpuImpl/RateAndState.h:418-433). */
  void* initial_pressure;                             /* InitialPressure [n][ppad] */
  const void* nucleation_stress;                      /* NucleationStressInFaultCS [n][nucleation_count][6][ppad] */
  const void* nucleation_pressure;                    /* NucleationPressure [n][nucleation_count][ppad] */
  /* friction energy output (drParameters.isFrictionEnergyRequired; all NULL = off): common::computeFrictionEnergy
   * (FrictionSolverCommon.h:632-722, called at the end of BaseFrictionLaw::evaluate :173-182) accumulates DREnergyOutput
   * (src/Initializer/Typedefs.h:237-241) from the interpolated fields of the step. */
  void* energy_slip;                                  /* DREnergyOutput::slip [n][3][ppad], in/out */
  void* energy_accumulated_slip;                      /* DREnergyOutput::accumulatedSlip [n][ppad], in/out */
  void* frictional_energy;                            /* DREnergyOutput::frictionalEnergy [n][ppad], in/out */
  const void* space_weights;                          /* [P] quadrature weights of the fault points */
  const void* doubled_surface_area;                   /* [n] DRGodunovData::doubledSurfaceArea */
} ssb_friction_data_t;

/* Scalars of FrictionSolver::evaluate and FrictionLawParameters (DynamicRupture/Misc.h:254-287). */
typedef struct ssb_friction_params {
  double delta_t[8];       /* FrictionSolver::computeDeltaT (FrictionSolver.cpp:20-37) */
  double time_weights[8];
  double full_update_time;
  double eta_p_damp;       /* drParameters.etaDamp while etaDampEnd > fullUpdateTime, else 1 */
  double healing_threshold;
  double t0;
  double rs_sr0;
  double rs_slip_rate_tolerance;
  double rs_state_tolerance;
  int rs_max_number_slip_rate_updates;
  int rs_number_state_variable_updates;
  int nucleation_count;    /* drParameters.nucleationCount (<= 4); 0 = no nucleation */
  int energies_from_across_fault_velocities;  /* drParameters.energiesFromAcrossFaultVelocities (friction energy output) */
  double nucleation_t0[4]; /* drParameters.t0[i]: effective rise time of nucleation i (t0[0] is also `t0` above) */
  double nucleation_s0[4]; /* drParameters.s0[i]: onset time of nucleation i */
  /* Rate-and-state in SINGLE precision: the reference's absolute Newton tolerances (1e-8) lie below the float
   * resolution of slip rates of O(1): its float build never converges and always runs all 10 x 60 iterations.
   * 0 (default): the same behaviour, iteration for iteration.  1: an iteration also counts as converged when the
   * residual g = (shear - |normal| mu) / eta - V is at the rounding level of its terms,
   * |g| <= 8 eps ((shear + |normal| mu) / eta + V), i.e. when further iterations only move the iterate among
   * neighbouring floats.  Ignored in double precision. */
  int rs_float_resolution_exit;
  int reserved;
} ssb_friction_params_t;

/* FrictionSolver::evaluate (src/DynamicRupture/FrictionLaws/FrictionSolver.h:41) for
 * LinearSlipWeakeningLaw<NoSpecialization> and RateAndStateBase<FastVelocityWeakeningLaw<NoTP>>;
 * one thread per fault quadrature point, convergence decided per face like the CPU implementation. */
int ssb_friction_lsw(ssb_ctx* ctx, const ssb_friction_data_t* data, const ssb_friction_params_t* params,
                     long long n_faces, void* stream);
int ssb_friction_rsfvw(ssb_ctx* ctx, const ssb_friction_data_t* data, const ssb_friction_params_t* params,
                       long long n_faces, void* stream);

/* neigh_dr: nodal flux of the imposed state into the adjacent elements.  Replaces the DynamicRupture
 * branch of Neighbor::computeBatchedNeighborsIntegral (src/Kernels/LinearCK/Neighbor.cpp:150-186).
 * godunov / flux_solver: device [n][4] pointer tables (CellDRMapping, src/Initializer/Typedefs.h:265-270),
 * NULL entries for non-DR faces; side_rel: device uint8 [n][4][2] = {side, faceRelation}. */
int ssb_neigh_dr(ssb_ctx* ctx, ssb_array_t dofs, const void* const* godunov, const void* const* flux_solver,
                 const uint8_t* side_rel, const uint8_t* face_types, long long n, void* stream);

/* ------------------------------------------------------------------------------------------
 * viscoelastic2 (anelastic attenuation; context created with equation = 1, mechanisms = 3, order 5).
 * Additional per-element arrays: dofs_ane = Qane [M][6][ld] (LTS::DofsAne), idofs_ane = time-integrated memory
 * variables Iane (LocalTmp::timeIntegratedAne of the reference, here a caller-provided array because ader and
 * local are separate launches); local_int = LocalIntegrationData of the viscoelastic2 build:
 * star[3][33] | nApNm1[4][135] | E[162] | w[M] | W[M]; neigh_int = nAmNm1[4][135] | w[M]
 * (src/Initializer/Typedefs.h:187-210, src/Equations/viscoelastic2/Model/IntegrationData.h).
 *
 * ssb_visco_ader  replaces linearckanelastic Spacetime::computeBatchedAder (src/Kernels/LinearCKAnelastic/Time.cpp:122-190,
 *                 per-cell computeAder :50-120);
 * ssb_visco_local replaces Local::computeBatchedIntegral (src/Kernels/LinearCKAnelastic/Local.cpp:112-250, per-cell :63-110);
 * ssb_visco_neigh replaces Neighbor::computeBatchedNeighborsIntegral (src/Kernels/LinearCKAnelastic/Neighbor.cpp:162-260,
 *                 per-cell :71-130), regular faces;
 * ssb_visco_neigh_dr: its DynamicRupture branch (:107-115, batched :171-234): for every face of FaceType DynamicRupture
 *                 Qext += V3mTo2nTWDivM(side, rel) . godunov . fluxSolver, then kernel::neighbor (Q += Qext[:, :9],
 *                 Qane += w[m] Qext[:, 9:]) -- linear in Qext, hence a launch of its own next to ssb_visco_neigh.
 *                 godunov: device [n][4] pointers to the imposed state of the face [9][ppad] (CellDRMapping::godunov,
 *                 NULL = no fault face); flux_solver: [n][4] pointers to the 9 x 15 flux solver (column-major
 *                 [c + 9 q]); side_rel: uint8 [n][4][2] = {side, faceRelation}; ppad = ssb_query()[23]. */
int ssb_visco_ader(ssb_ctx* ctx, ssb_array_t dofs, ssb_array_t dofs_ane, ssb_array_t idofs, ssb_array_t idofs_ane,
                   ssb_array_t local_int, const double* coeffs, long long n, void* stream);
int ssb_visco_local(ssb_ctx* ctx, ssb_array_t dofs, ssb_array_t dofs_ane, ssb_array_t idofs, ssb_array_t idofs_ane,
                    ssb_array_t local_int, const uint8_t* face_types, long long n, void* stream);
int ssb_visco_neigh(ssb_ctx* ctx, ssb_array_t dofs, ssb_array_t dofs_ane, const void* const* nb_ptrs,
                    const long long* nb_idx, ssb_array_t idofs, ssb_array_t neigh_int, const uint8_t* face_types,
                    const uint8_t* face_rel, long long n, const int* perm, void* stream);

int ssb_visco_neigh_dr(ssb_ctx* ctx, ssb_array_t dofs, ssb_array_t dofs_ane, ssb_array_t neigh_int,
                       const void* const* godunov, const void* const* flux_solver, const uint8_t* side_rel,
                       const uint8_t* face_types, long long n, void* stream);

/* Nodal boundary conditions of the local integral (Local::computeIntegral, src/Kernels/LinearCK/Local.cpp:157-252; batched
 * :350-455 through DirichletBoundary::evaluateOnDevice): for m boundary faces (cell[i], face[i]) of FaceType kind[i]
 *   INodal = V3mTo2nFace_f . I . Tinv^T;   Dirichlet (4): INodal(l, a) = sum_b map(a, b, l) INodal(l, b) + const(a, l)
 *   (createEasiBoundaryGhostCells);   FreeSurfaceGravity (2): INodal(l, 0..2) = 2 p - INodal(l, 0..2), p = -rho g
 *   displacement(l);   Q += project2nFaceTo3m_f . INodal . AminusT_f (localFluxNodal; neigh_int holds AminusT as for
 *   ssb_neigh, with the identity in place of Tinv for these faces as the reference sets it up).
 * Call it after ssb_local (the reference evaluates these faces inside the local integral).  All face arrays are DEVICE
 * arrays in the precision of the context: tinv [m][9][9] row-major Tinv(p, q); easi_map [m][9][9][B2] (a, b, l) and
 * easi_const [m][9][B2] (Dirichlet faces; B2 = O(O+1)/2 face nodes); displacement [m][B2] and rho_g [m]
 * (FreeSurfaceGravity faces).  Faces of any other kind are skipped.  Two boundary faces of one cell are added with
 * atomic adds (their order is not fixed).  FaceType::Analytical faces: ssb_local_nodal_flux below. */
int ssb_local_nodal_bc(ssb_ctx* ctx, ssb_array_t dofs, ssb_array_t idofs, ssb_array_t neigh_int, const long long* cell,
                       const uint8_t* face, const uint8_t* kind, const void* tinv, const void* easi_map,
                       const void* easi_const, const void* displacement, const void* rho_g, long long m, void* stream);

/* FaceType::Analytical faces of the local integral (Local::evaluateBatchedTimeDependentBc, Local.cpp:397-455): the
 * reference evaluates the analytical solution at the face nodes on the HOST (a user callback per node and time point,
 * integrated over the time step by DirichletBoundary::evaluateTimeDependent, DirichletBoundary.h:121-165), stores the
 * result in LTS::AnalyticScratch and launches the generated localFluxNodal on the device.  This entry point is that launch
 * for m faces (cell[i], face[i]):   Q += project2nFaceTo3m_f . INodal_i . AminusT_f.
 * inodal: DEVICE array [m][9][B2] (quantity a, node l; B2 = O(O+1)/2) in the precision of the context. */
int ssb_local_nodal_flux(ssb_ctx* ctx, ssb_array_t dofs, ssb_array_t neigh_int, const long long* cell,
                         const uint8_t* face, const void* inodal, long long m, void* stream);

/* Volume part of the energy output (EnergyOutput::computeVolumeEnergies, src/ResultWriter/EnergyOutput.cpp:347-486) for n
 * elements: kinetic energy 0.5 rho |v|^2, elastic strain energy 0.5 sigma : eps(sigma), acoustic energy p^2 / (2 K) of
 * elements with mu = 0, total momentum rho v, integrated over every element.  The reference evaluates Q at (O+1)^3
 * quadrature points per element; the integrands are quadratic in the orthogonal modal basis, so the kernel takes the same
 * integrals from the modal coefficients in one streaming pass over Q (sums in double).  volume [n] and material [n][3] =
 * lambda, mu, rho are DEVICE arrays of doubles.  out: DEVICE double[8], ACCUMULATED (zero it first): [0] elastic kinetic
 * energy, [1] elastic energy, [2] acoustic kinetic energy, [3] acoustic energy, [4..6] momentum x, y, z, [7] unused (the
 * gravitational energy of free-surface-gravity faces and the plastic moment are not covered; the fault's frictional
 * energy is an output of ssb_friction_*).  Ghost / duplicate cells are the caller's to leave out, as :405-408 does. */
int ssb_volume_energies(ssb_ctx* ctx, ssb_array_t dofs, const double* volume, const double* material, long long n,
                        double* out, void* stream);

/* Fault part of the energy output (EnergyOutput::computeDynamicRuptureEnergies, src/ResultWriter/EnergyOutput.cpp:262-345)
 * for m fault faces: total frictional work = sum of DREnergyOutput::frictionalEnergy over the points, potency = sum over
 * the faces of (sum of accumulatedSlip over the points) * 0.5 * doubledSurfaceArea / P, seismic moment = potency
 * increment * harmonic mean of the two shear moduli.  frictional_energy / accumulated_slip [m][ppad] are the arrays
 * ssb_friction_* updates (ssb_friction_data_t::frictional_energy, energy_accumulated_slip), doubled_surface_area [m] in the
 * precision of the context; material_plus / minus [m][3] = lambda, mu, rho (doubles, as for ssb_setup_dr_matrices);
 * plus_side_on_this_rank [m] or NULL (all faces).  out: DEVICE double[4], ACCUMULATED: [0] total frictional work, [1]
 * potency, [2] seismic moment, [3] unused (the static frictional work, computeStaticWork, is not covered). */
int ssb_fault_energies(ssb_ctx* ctx, const void* frictional_energy, const void* accumulated_slip,
                       const void* doubled_surface_area, const double* material_plus, const double* material_minus,
                       const uint8_t* plus_side_on_this_rank, long long m, double* out, void* stream);

/* ------------------------------------------------------------------------------------------
 * Set-up on the device (elastic): the per-element and per-fault-face matrices the entry points above consume, from
 * vertex coordinates and material parameters; all inputs are DEVICE arrays of doubles, results are stored in the
 * precision of the context.
 * ssb_setup_cell_matrices replaces initializeCellLocalMatrices (src/Initializer/Model/CellLocalMatrices.cpp:66-298) for
 * the elastic model: local_int = star[3][24] | nApNm1[4][81], neigh_int = nAmNm1[4][81].  coords [n][4][3];
 * neighbors [n][4]: row of `material` of the face neighbor (< 0: none, the own material is used); material [*][3] =
 * lambda, mu, rho (rows 0..n-1 = own cells, further rows = ghost cells); face_types [n][4] or NULL: FreeSurface (1) gets
 * the free-surface Godunov state (src/Model/Common.h:319-364), DynamicRupture (3) no local flux matrix.
 * ssb_setup_dr_matrices replaces the matrix part of initializeDynamicRuptureMatrices
 * (src/Initializer/Model/DynamicRuptureMatrices.cpp:111-477: impedances and eta :341-359, TinvT :427, flux solvers
 * :461-469): coords_plus / coords_minus [m][4][3] of the two elements of every fault face, face_info [m][3] = {plusSide,
 * minusSide, faceRelation}, material_plus / minus [m][3]; out: tinv_t [m][81], imp_and_eta [m][11], flux_solver_plus /
 * minus [m][81] (column-major [c + 9 q]), doubled_surface_area [m] (may be NULL). */
int ssb_setup_cell_matrices(ssb_ctx* ctx, const double* coords, const long long* neighbors, const double* material,
                            const uint8_t* face_types, ssb_array_t local_int, ssb_array_t neigh_int, long long n,
                            void* stream);
int ssb_setup_dr_matrices(ssb_ctx* ctx, const double* coords_plus, const double* coords_minus, const uint8_t* face_info,
                          const double* material_plus, const double* material_minus, void* tinv_t, void* imp_and_eta,
                          void* flux_solver_plus, void* flux_solver_minus, void* doubled_surface_area, long long m,
                          void* stream);

const char* ssb_error_string(int code);
const char* ssb_version(void);

#ifdef __cplusplus
}
#endif
#endif
