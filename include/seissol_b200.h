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
 * [4..4+O) = der_offset, [12..12+O) = der_ld, [20] = elements per CTA (local), [21] = (neigh) */
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

/* Neighbor flux, all four faces and all 48 face relations in one launch.  Replaces
 * Neighbor::computeBatchedNeighborsIntegral (src/Kernels/LinearCK/Neighbor.cpp:111-188).
 * Neighbor data either as nb_ptrs (device [n][4] pointers to the neighbors' I,
 * LTS::FaceNeighborsDevice) or nb_idx (device [n][4] indices into `idofs`).
 * neigh_int: NeighboringIntegrationData (nAmNm1[4][81]); face_rel: device uint8 [n][4][2] =
 * {neighbor's local face id, orientation} (CellLocalInformation.h:26-27). */
int ssb_neigh(ssb_ctx* ctx, ssb_array_t dofs, const void* const* nb_ptrs, const long long* nb_idx,
              ssb_array_t idofs, ssb_array_t neigh_int, const uint8_t* face_types,
              const uint8_t* face_rel, long long n, const int* perm, void* stream);

const char* ssb_error_string(int code);
const char* ssb_version(void);

#ifdef __cplusplus
}
#endif
#endif
