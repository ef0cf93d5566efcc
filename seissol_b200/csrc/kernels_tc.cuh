// The Cauchy-Kovalevskaya predictor (`ader`) on the 5th-generation tensor cores, single precision.
//
// Reference: Spacetime::computeBatchedAder (src/Kernels/LinearCK/Time.cpp:142-217), generated kernel::derivative
//   dQ(t) = sum_c kDivMT_c . dQ(t-1) . star_c,  I = sum_t c_t dQ(t)                  (codegen/kernels/aderdg.py:463-532)
//
// The CUDA-core kernels (kernels_elastic.cuh) carry every operator as immediates: ~10^4 straight-line FMAs per thread, an
// instruction stream of 60-200 KB that the instruction caches cannot hold -- they issue at 0.57 IPC, 42 % of the FMA peak
// (DESIGN.md §4).  Here the constant operator is a GEMM operand instead:
//
//   Y[(e, q)][(c, k')] = sum_k dQ(t-1)[k, (e, q)] . kDivMT_c[k', k]     tcgen05.mma kind::tf32, M = 128 rows = 14 elements x 9
//                                                                       quantities, N = 3 x rows(t), K = rows(t-1)
//   dQ(t)[k', (e, q')] = sum_c sum_q Y[(e, q)][(c, k')] . star_c[q, q'] CUDA cores: a thread owns one column (e, q') as in
//                                                                       the CUDA-core kernels; the 2-3 source columns q of
//                                                                       the same element come through shared memory
//
// with a 3-term tf32 split (a b ~ ah bh + ah bl + al bh, round-to-nearest hi parts: relative error ~ 2^-22 per product,
// the accuracy of an fp32 FMA chain of this length) and fp32 accumulation in TMEM.  The loop over steps is ~1.3 k
// instructions per thread, resident in the instruction cache.
//
// One persistent CTA per SM, kNW independent "workers" of 256 threads: one TMEM lane = one column (e, q), shared by two
// threads that own the even / the odd 4-row chunks of the column (16 warps per SM hide the shared-memory and TMEM
// latencies; one thread per column ran at 0.09 IPC per warp).  A worker walks through
// load -> split -> [MMA -> TMEM -> exchange -> star -> split] x (O-1) -> store for its tile while the other worker's MMAs
// run, so tensor pipe and CUDA cores overlap.  All operators (hi and lo, every step) stay resident in shared
// memory (80 KB at order 6); a worker owns one 56 KB region that is in turn the landing zone of the bulk copies, the A
// operand, the exchange buffer and the staging area of the bulk store.
#pragma once
#include "common.cuh"
#include "params.cuh"
#include "umma.cuh"
#include "kernels_elastic.cuh"

namespace ssb {
namespace tc {

constexpr int kTE = 14;     // elements per tile: 126 of the 128 MMA rows
constexpr int kNW = 2;      // workers per CTA
constexpr int kRows = 128;  // MMA rows (TMEM lanes) per worker
constexpr int kWT = 256;    // threads per worker: two per row
__host__ __device__ constexpr int pad8i(int n) { return (n + 7) / 8 * 8; }
__host__ __device__ constexpr int pad16i(int n) { return (n + 15) / 16 * 16; }
// row stride of the exchange buffer (floats): a multiple of 4 whose quarter is odd -> 128-bit accesses of 8 consecutive
// rows hit 8 distinct 16-byte bank groups
__host__ __device__ constexpr int xStride(int rows) {
  const int u = (rows + 3) / 4;
  return (u % 2 == 1 ? u : u + 1) * 4;
}
// chunks (of 4 rows) a thread owns out of n: chunk k4 belongs to half h = k4 & 1, the thread's j-th chunk is k4 = 2 j + h
__host__ __device__ constexpr int ownChunks(int n) { return (n + 1) / 2; }

template <typename G>
struct AderTc {
  static constexpr int O = G::kOrder, B = G::kB, BP = G::kBP, TEN = kNQ * BP;
  __host__ __device__ static constexpr int rowsIn(int t) { return G::derRows(t - 1); }
  __host__ __device__ static constexpr int rowsOut(int t) { return G::derRows(t); }
  __host__ __device__ static constexpr int K(int t) { return pad8i(rowsIn(t)); }
  __host__ __device__ static constexpr int RP(int t) { return pad4i(rowsOut(t)); }
  __host__ __device__ static constexpr int N(int t) { return pad16i(3 * RP(t)); }
  __host__ __device__ static constexpr int S(int t) { return xStride(rowsOut(t)); }
  // float offset of [hi_t | lo_t] inside the operator blob (gen_tc.py)
  __host__ __device__ static constexpr int opOffset(int t) {
    int o = 0;
    for (int u = 1; u < t; ++u) {
      o += 2 * N(u) * K(u);
    }
    return o;
  }
  static constexpr int opsFloats = opOffset(O);
  __host__ __device__ static constexpr int regionFloats() {
    int m = TEN * kTE;
    for (int t = 1; t < O; ++t) {
      m = imax(m, 2 * kRows * K(t));
      m = imax(m, 3 * kRows * S(t));
    }
    return (m + 31) / 32 * 32;
  }
  static constexpr int tmemColsPerWorker = N(1) <= 32 ? 32 : (N(1) <= 64 ? 64 : (N(1) <= 128 ? 128 : 256));
  static constexpr int tmemCols = kNW * tmemColsPerWorker;
  static constexpr int CH = BP / 4;            // chunks of a full column
  static constexpr int OC = ownChunks(CH);     // ... owned by one thread
  static constexpr size_t barOffset = (size_t(opsFloats) + size_t(kNW) * regionFloats()) * sizeof(float);
  static constexpr size_t bytes = barOffset + 128;
  static_assert(tmemCols <= 512, "TMEM");
  static_assert(K(1) >= B, "the first operand holds a whole column");
};

__device__ __forceinline__ void worker_sync(int w) { asm volatile("bar.sync %0, 256;" ::"r"(w + 1) : "memory"); }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "elect.sync _|P1, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P1;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&r)[4]) {
  uint32_t a, b, c, d;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(a), "=r"(b), "=r"(c), "=r"(d)
               : "r"(taddr)
               : "memory");
  r[0] = __uint_as_float(a); r[1] = __uint_as_float(b); r[2] = __uint_as_float(c); r[3] = __uint_as_float(d);
}

// tf32 split with round-to-nearest: x = hi + lo exactly in fp32, the tensor core reads the upper 19 bits of each
// (cvt.rna.tf32.f32 expands to five instructions for its Inf/NaN handling; the data here is finite)
__device__ __forceinline__ float tf32_rna(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}

// The thread's chunks of one column (NV chunks hold data, zero beyond) -> its row of the A operand (KP columns):
// canonical K-major, no swizzle, element (r, k) at float offset (k / 4) * 512 + r * 4 + k % 4  (LBO = 2048 B, SBO = 128 B)
template <int KP, int NV, int NOWN>
__device__ __forceinline__ void writeOperandA(float* region, int row, int h, const float (&v)[NOWN][4]) {
  float* const hiBase = region + row * 4;
  float* const loBase = hiBase + kRows * KP;
#pragma unroll
  for (int j = 0; j < ownChunks(KP / 4); ++j) {
    const int k4 = 2 * j + h;
    if (k4 < KP / 4) {
      float hi[4], lo[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float x = (j < NOWN && k4 < NV) ? v[j < NOWN ? j : 0][i] : 0.f;
        hi[i] = tf32_rna(x);
        lo[i] = x - hi[i];
      }
      st4(hiBase + k4 * 512, hi[0], hi[1], hi[2], hi[3]);
      st4(loBase + k4 * 512, lo[0], lo[1], lo[2], lo[3]);
    }
  }
}

template <typename G>
struct AderWorker {
  using C = AderTc<G>;
  static constexpr int O = C::O, BP = C::BP, TEN = C::TEN;
  const LocalParams& p;
  float* region;
  const float* ops;
  uint64_t* mmaBar;
  uint32_t tmemD;      // this worker's accumulator columns, lane field of this warp included
  uint32_t mmaPhase;
  int w, row, h, q;
  bool issuer;         // warp 0 of the worker (warp-uniform)
  bool wantDer;
  float* derBase;      // derivative family of this thread's element or null
  float s[3][3];
  int srcRow[3];       // exchange rows of the star sources of column q
  float icol[C::OC][4];

  // step T: dQ(T-1) is in the A operand; returns the thread's chunks of dQ(T) of its column in out[]
  template <int T>
  __device__ __forceinline__ void step(float (&out)[ownChunks(C::RP(T) / 4)][4]) {
    constexpr int K = C::K(T), RP = C::RP(T), N = C::N(T), S = C::S(T), NCH = RP / 4, OCH = ownChunks(NCH);
    if (issuer) {
      if (elect_one()) {
        umma::fence_after_sync();
        const uint32_t idesc = umma::instr_desc_tf32(128, N);
        const uint32_t aH = umma::smem_addr(region), aL = aH + kRows * K * 4;
        const uint32_t bH = umma::smem_addr(ops + C::opOffset(T)), bL = bH + N * K * 4;
        const uint64_t dah = umma::smem_desc(aH, 2048, 128), dal = umma::smem_desc(aL, 2048, 128);
        const uint64_t dbh = umma::smem_desc(bH, N * 16, 128), dbl = umma::smem_desc(bL, N * 16, 128);
        const uint32_t d = tmemD & 0x0000ffffu;
#pragma unroll
        for (int ks = 0; ks < K / 8; ++ks) {
          const uint64_t oa = static_cast<uint64_t>((ks * 2 * 2048) >> 4), ob = static_cast<uint64_t>((ks * 2 * N * 16) >> 4);
          umma::mma_tf32(d, dal + oa, dbh + ob, idesc, ks > 0);
          umma::mma_tf32(d, dah + oa, dbl + ob, idesc, true);
          umma::mma_tf32(d, dah + oa, dbh + ob, idesc, true);
        }
        umma::commit(mmaBar);
      }
      __syncwarp();
    }
    mbar_wait(mmaBar, mmaPhase);
    mmaPhase ^= 1;
    umma::fence_after_sync();
    // ---- own chunks of Y[(e, q)][(c, k')] -> exchange buffer X[c][row][S]   (the A operand is dead: the MMAs are done)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float y[OCH][4];
#pragma unroll
      for (int j = 0; j < OCH; ++j) {
        const int k4 = 2 * j + h;
        if (k4 < NCH) {   // warp-uniform
          tmem_ld4(tmemD + c * RP + 4 * k4, y[j]);
        }
      }
      umma::tmem_ld_wait();
      float* dst = region + (c * kRows + row) * S;
#pragma unroll
      for (int j = 0; j < OCH; ++j) {
        const int k4 = 2 * j + h;
        if (k4 < NCH) {
          st4(dst + 4 * k4, y[j][0], y[j][1], y[j][2], y[j][3]);
        }
      }
    }
    umma::fence_before_sync();
    worker_sync(w);
    // ---- star: out[k'] = sum_c sum_i s[c][i] . Y_c[k'][source i]
#pragma unroll
    for (int j = 0; j < OCH; ++j) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        out[j][i] = 0.f;
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float* src = region + (c * kRows + srcRow[i]) * S + 4 * h;
        const float sv = s[c][i];
#pragma unroll
        for (int j = 0; j < OCH; ++j) {
          if (2 * j + h < NCH) {
            float t4[4];
            ld4(t4, src + 8 * j);
#pragma unroll
            for (int l = 0; l < 4; ++l) {
              out[j][l] = fmaf(sv, t4[l], out[j][l]);
            }
          }
        }
      }
    }
    // ---- Taylor sum and the derivative family
    const float ct = static_cast<float>(p.coeffs[T]);
#pragma unroll
    for (int j = 0; j < OCH; ++j) {
#pragma unroll
      for (int l = 0; l < 4; ++l) {
        icol[j][l] = fmaf(ct, out[j][l], icol[j][l]);
      }
    }
    if (wantDer && derBase != nullptr) {
      float* dst = derBase + p.derOffset[T] + q * p.derLd[T] + 4 * h;
#pragma unroll
      for (int j = 0; j < OCH; ++j) {
        if (2 * j + h < NCH) {
          st4(dst + 8 * j, out[j][0], out[j][1], out[j][2], out[j][3]);
        }
      }
    }
    worker_sync(w);   // everybody has read the exchange buffer: the region may be rewritten
  }

  template <int T>
  __device__ __forceinline__ void chain(const float (&prev)[ownChunks(C::RP(T - 1) / 4)][4]) {
    // prev = dQ(T-1) (T >= 2): operand of step T
    writeOperandA<C::K(T), C::RP(T - 1) / 4>(region, row, h, prev);
    fence_async_smem();
    worker_sync(w);
    float out[ownChunks(C::RP(T) / 4)][4];
    step<T>(out);
    if constexpr (T + 1 < O) {
      chain<T + 1>(out);
    }
  }
};

template <typename G, typename GT>
__global__ void __launch_bounds__(kNW* kWT, 1) k_ader_tc(const LocalParams p) {
  using C = AderTc<G>;
  constexpr int O = C::O, BP = C::BP, TEN = C::TEN, CH = C::CH, OC = C::OC;
  extern __shared__ __align__(128) unsigned char smemRaw[];
  float* const smem = reinterpret_cast<float*>(smemRaw);
  float* const ops = smem;
  // bar[0]: operators, bar[1 + 2 w]: tile loaded, bar[2 + 2 w]: MMAs done
  uint64_t* const bar = reinterpret_cast<uint64_t*>(smemRaw + C::barOffset);
  uint32_t* const tmemSlot = reinterpret_cast<uint32_t*>(bar + 1 + 2 * kNW);

  const int tid = threadIdx.x;
  const int w = tid / kWT, wt = tid % kWT;
  const int row = wt & (kRows - 1), h = wt >> 7;
  float* const region = smem + C::opsFloats + w * C::regionFloats();
  if (tid == 0) {
    mbar_init(bar, 1);
    for (int i = 0; i < kNW; ++i) {
      mbar_init(bar + 1 + 2 * i, kTE);
      mbar_init(bar + 2 + 2 * i, 1);
    }
    mbar_fence_init();
  }
  if (tid < 32) {
    umma::tmem_alloc(tmemSlot, C::tmemCols);
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  if (tid == 0) {
    mbar_arrive_expect_tx(bar, C::opsFloats * sizeof(float));
    tma_load(ops, GT::aderOps(), C::opsFloats * sizeof(float), bar);
  }
  const uint32_t tmemBase = *tmemSlot;

  AderWorker<G> W{p, region, ops, bar + 2 + 2 * w};
  W.tmemD = tmemBase + w * C::tmemColsPerWorker + ((static_cast<uint32_t>(row >> 5) * 32u) << 16);
  W.mmaPhase = 0;
  W.w = w; W.row = row; W.h = h;
  W.issuer = __shfl_sync(0xffffffffu, wt >> 5, 0) == 0;
  const int e = row / kNQ, q = row - (row / kNQ) * kNQ;
  W.q = q;
  W.wantDer = p.derivs.present();
  const bool rowUsed = row < kTE * kNQ;
#pragma unroll
  for (int i = 0; i < 3; ++i) {
    W.srcRow[i] = rowUsed ? e * kNQ + kStarRow[q][i] : row;
  }
  uint64_t* const ldBar = bar + 1 + 2 * w;
  uint32_t ldPhase = 0;
  bool opsReady = false;

  const long long nTiles = (p.n + kTE - 1) / kTE;
  for (long long tile = static_cast<long long>(blockIdx.x) * kNW + w; tile < nTiles;
       tile += static_cast<long long>(gridDim.x) * kNW) {
    const long long slot = tile * kTE;
    // ---- bulk loads: thread wt < kTE moves element slot + wt
    long long myElem = -1;
    float* const mySlot = region + wt * TEN;
    if (wt < kTE) {
      const long long sl = slot + wt;
      if (sl < p.n) {
        myElem = p.perm != nullptr ? p.perm[sl] : sl;
      }
      if (myElem >= 0) {
        mbar_arrive_expect_tx(ldBar, TEN * sizeof(float));
        tmaLoadTensor<float, BP>(mySlot, static_cast<const float*>(p.dofs.at(myElem)), p.ldg, ldBar);
      } else {
        mbar_arrive_expect_tx(ldBar, 0);
      }
    }
    // ---- this thread's element and its star entries of column q (global loads under the bulk copies)
    long long el = -1;
    if (rowUsed && slot + e < p.n) {
      el = p.perm != nullptr ? p.perm[slot + e] : slot + e;
    }
    const float* const LI = el >= 0 ? static_cast<const float*>(p.localInt.at(el)) : nullptr;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        W.s[c][i] = (LI != nullptr && i < kStarCnt[q]) ? LI[24 * c + kStarOff[q] + i] : 0.f;
      }
    }
    W.derBase = (W.wantDer && el >= 0) ? static_cast<float*>(const_cast<void*>(p.derivs.at(el))) : nullptr;

    mbar_wait(ldBar, ldPhase);
    ldPhase ^= 1;
    // dQ(0) = Q into the derivative family (Time.cpp:109-116), straight from the landing zone
    if (W.wantDer && wt < kTE && myElem >= 0) {
      float* const dbase = static_cast<float*>(const_cast<void*>(p.derivs.at(myElem)));
      if (dbase != nullptr) {
        float* dst = dbase + p.derOffset[0];
        for (int c = 0; c < kNQ; ++c) {
          tma_store(dst + c * p.derLd[0], mySlot + c * BP, BP * sizeof(float));
        }
        tma_store_commit();
        tma_store_wait_read();
      }
    }
    // ---- own chunks of the column -> registers; I = c0 Q so far
    {
      float col[OC][4];
      const float* src = region + e * TEN + q * BP + 4 * h;
#pragma unroll
      for (int j = 0; j < OC; ++j) {
        if (el >= 0 && 2 * j + h < CH) {
          ld4(col[j], src + 8 * j);
        } else {
          col[j][0] = col[j][1] = col[j][2] = col[j][3] = 0.f;
        }
      }
      const float c0 = static_cast<float>(p.coeffs[0]);
#pragma unroll
      for (int j = 0; j < OC; ++j) {
#pragma unroll
        for (int l = 0; l < 4; ++l) {
          W.icol[j][l] = c0 * col[j][l];
        }
      }
      worker_sync(w);   // everybody holds its column (and dQ(0) has left): the landing zone becomes the A operand
      if (!opsReady) {
        mbar_wait(bar, 0);
        opsReady = true;
      }
      writeOperandA<C::K(1), CH>(region, row, h, col);
    }
    fence_async_smem();
    worker_sync(w);
    {
      float out[ownChunks(C::RP(1) / 4)][4];
      W.template step<1>(out);
      if constexpr (O > 2) {
        W.template chain<2>(out);
      }
    }
    // ---- I staged per element for the bulk store (the last exchange buffer is dead: step() ends with a barrier)
    if (el >= 0) {
      float* dst = region + e * TEN + q * BP + 4 * h;
#pragma unroll
      for (int j = 0; j < OC; ++j) {
        if (2 * j + h < CH) {
          st4(dst + 8 * j, W.icol[j][0], W.icol[j][1], W.icol[j][2], W.icol[j][3]);
        }
      }
    }
    fence_async_smem();
    worker_sync(w);
    if (wt < kTE && myElem >= 0) {
      tmaStoreTensor<float, BP>(static_cast<float*>(const_cast<void*>(p.idofs.at(myElem))), mySlot, p.ldg);
      tma_store_wait_read();
    }
    worker_sync(w);   // the staging area is free again
  }
  umma::fence_before_sync();
  __syncthreads();
  if (tid < 32) {
    umma::tmem_dealloc(tmemBase, C::tmemCols);
  }
}

}  // namespace tc
}  // namespace ssb
