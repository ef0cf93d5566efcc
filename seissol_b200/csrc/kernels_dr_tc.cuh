// godunov_dr on the 5th-generation tensor cores (tcgen05 + TMEM), single precision.
//
// Reference: kernels::DynamicRupture::batchedSpaceTimeInterpolation (src/Kernels/DynamicRupture.cpp:110-168),
//   QInterpolated±[s] = V3mTo2n(side, rel) . (sum_d coeff[s][d] dQ±(d)) . TinvT     (codegen/kernels/dynamic_rupture.py:108-172)
//
// V3mTo2n is the one DENSE operator of the path ((O+1)^2 x B, 49 x 56 at order 6, 16 (side, relation) variants): as
// immediate-operand CUDA-core code it is 43 k instructions per kernel (680 KB, far beyond the instruction caches) and
// runs at 19 % of the HBM ceiling.  Here it is a GEMM on the tensor cores:
//
//   D[(face, s, q)][point] = sum_k A[(face, s, q)][k] . B[point][k]
//     A = the rotated Taylor sums, rows = (face, time point s, quantity q): 9 O rows per face side, FPT faces per
//         128-row tile, K = basis functions.  The CUDA cores rotate the derivative families in place in shared memory
//         (TinvT commutes with the left operator), then the two threads that own a row (= one TMEM lane) form its Taylor
//         sums in registers, split them into tf32 hi / lo and write them straight into TENSOR MEMORY (tcgen05.st): the
//         A operand never touches shared memory (MMA with A from TMEM).
//     B = V3mTo2n(variant), precomputed by the generator (gen_dr.py) in the canonical K-major no-swizzle layout, fetched
//         by bulk copies when the variant changes (faces arrive sorted by variant),
//   as THREE kind::tf32 products per K step (a b ~ al bh + ah bl + ah bh: relative error ~ 2^-21; plain TF32 has 2^-11,
//   the single-precision bar is 1e-5), accumulated in TMEM in fp32 and read back with tcgen05.ld: a thread stores half
//   of one contiguous row of QInterpolated.
//
// One persistent CTA per SM with NW independent workers of 256 threads; a worker owns a tile of FPT faces from the bulk
// loads (two stages: the next tile is in flight while this one is computed) to the stores, so the MMAs of one worker run
// under the CUDA-core work of the other.
#pragma once
#include "common.cuh"
#include "params.cuh"
#include "umma.cuh"

namespace ssb {

template <typename G, typename GD>
struct DrTcCfg {
  static constexpr int O = G::kOrder;
  static constexpr int B = G::kB;
  static constexpr int P = GD::kP;
  static constexpr int K = GD::kTcK;
  static constexpr int N = GD::kTcN;
  static constexpr int KG = K / 4;                 // core matrices along K
  static constexpr int LBO = 144;                  // bytes between core matrices along K (128 + 16 of padding)
  static constexpr int SBO = KG * LBO;             // bytes between 8-row groups
  static constexpr int RPF = 9 * O;                // A rows per face side
  static constexpr int FPT = 128 / RPF;            // faces per 128-row tile
  static constexpr int B_FLOATS = GD::kTcOperand;  // N rows
  static constexpr int NW = 2;                     // workers per CTA
  static constexpr int WT = 256;                   // threads per worker: two per row
  static constexpr int THREADS = NW * WT;
  static constexpr int TMEM_A = 2 * K;             // hi | lo
  static constexpr int TMEM_PER_WORKER = TMEM_A + 2 * N;   // A (hi | lo) and two accumulators (plus side, minus side)
  static constexpr int TMEM_COLS = NW * TMEM_PER_WORKER <= 128 ? 128 : (NW * TMEM_PER_WORKER <= 256 ? 256 : 512);
  static constexpr int CH = K / 4;                 // 4-column chunks of an A row
  static constexpr int OC = (CH + 1) / 2;          // ... owned by one of the two threads of the row
  __host__ __device__ static constexpr int totalRowGroups() {   // groups of 4 rows over all dQ(d)
    int t = 0;
    for (int d = 0; d < O; ++d) {
      t += (G::derRows(d) + 3) / 4;
    }
    return t;
  }
  static_assert(FPT >= 1 && NW * TMEM_PER_WORKER <= 512, "tile and TMEM budget");
  static_assert(B_FLOATS == (N / 8) * KG * 36, "generator and kernel disagree on the operand layout");
  // shared memory per worker (floats): B hi (plus, minus) | B lo (plus, minus) | two stages {families, rotations, meta}
  static constexpr int META = 4 * FPT + 4;         // per stage: {face id (2 words), variant plus, variant minus} per face
  __host__ __device__ static int stageFloats(int rawStride) { return 2 * FPT * rawStride + FPT * 108 + META; }
  __host__ __device__ static int workerFloats(int rawStride) { return 4 * B_FLOATS + 2 * stageFloats(rawStride); }
  __host__ __device__ static size_t bytes(int rawStride) {
    return size_t(NW) * workerFloats(rawStride) * sizeof(float) + 256;
  }
};

namespace drtc {
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
__device__ __forceinline__ void tmem_st4(uint32_t taddr, float a, float b, float c, float d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(a)),
               "r"(__float_as_uint(b)), "r"(__float_as_uint(c)), "r"(__float_as_uint(d))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] . B[smem]
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmemD, uint32_t tmemA, uint64_t descB, uint32_t idesc, bool accumulate) {
  const uint32_t acc = accumulate ? 1u : 0u;
  const uint32_t zero = 0u;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t"
      "}\n" ::"r"(tmemD),
      "r"(tmemA), "l"(descB), "r"(idesc), "r"(acc), "r"(zero)
      : "memory");
}
// round-to-nearest tf32 (the data is finite; cvt.rna.tf32.f32 expands to five instructions for its Inf/NaN handling)
__device__ __forceinline__ float tf32_rna(float x) { return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u); }
}  // namespace drtc

template <typename G, typename GD>
__global__ void __launch_bounds__((DrTcCfg<G, GD>::THREADS), 1) k_dr_interp_tc(const DrInterpParams p) {
  using C = DrTcCfg<G, GD>;
  constexpr int O = C::O, K = C::K, N = C::N, RPF = C::RPF, FPT = C::FPT, CH = C::CH, OC = C::OC;
  constexpr int TG = C::totalRowGroups();
  extern __shared__ __align__(128) unsigned char smemRaw[];
  float* const smem = reinterpret_cast<float*>(smemRaw);
  const int stageFloats = C::stageFloats(p.rawStride);
  const int workerFloats = C::workerFloats(p.rawStride);
  // per worker: bar[0..1]: stage full, bar[2]: B operand, bar[3]: MMA done
  uint64_t* const barAll = reinterpret_cast<uint64_t*>(smem + C::NW * workerFloats);
  uint32_t* const tmemSlot = reinterpret_cast<uint32_t*>(barAll + 4 * C::NW);

  const int tid = threadIdx.x;
  const int w = tid / C::WT, wt = tid % C::WT;
  const int row = wt & 127, h = wt >> 7;
  float* const wbase = smem + w * workerFloats;
  float* const sBh = wbase;                       // [side][B_FLOATS]
  float* const sBl = wbase + 2 * C::B_FLOATS;
  float* const sStage = wbase + 4 * C::B_FLOATS;
  uint64_t* const bar = barAll + 4 * w;

  if (tid == 0) {
    for (int i = 0; i < C::NW; ++i) {
      mbar_init(barAll + 4 * i + 0, 1);
      mbar_init(barAll + 4 * i + 1, 1);
      mbar_init(barAll + 4 * i + 2, 1);
      mbar_init(barAll + 4 * i + 3, 1);
    }
    mbar_fence_init();
  }
  if (tid < 32) {
    umma::tmem_alloc(tmemSlot, C::TMEM_COLS);
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmemBase = *tmemSlot + w * C::TMEM_PER_WORKER;
  const uint32_t laneBits = (static_cast<uint32_t>(row >> 5) * 32u) << 16;
  const uint32_t tmemA = tmemBase;            // hi: columns [0, K), lo: [K, 2K)
  const uint32_t tmemD0 = tmemBase + C::TMEM_A;
  const uint32_t tmemD1 = tmemD0 + N;
  const bool issuer = __shfl_sync(0xffffffffu, wt >> 5, 0) == 0;

  // row role: (face slot f, time point s, quantity q)
  const int f = row / RPF;
  const int rem = row - f * RPF;
  const int q = rem / O, s = rem - (rem / O) * O;   // s fastest: the lanes of a quarter warp read 1-2 distinct columns
  const bool rowUsed = f < FPT;

  const long long nTiles = (p.n + FPT - 1) / FPT;
  const long long tile0 = static_cast<long long>(blockIdx.x) * C::NW + w;
  const long long tileStride = static_cast<long long>(gridDim.x) * C::NW;

  // loader role: lane wt < 2 FPT of warp 0 moves one (face, side) family of a tile, its even lanes write the metadata;
  // the other warps fetch the rotations.  Split in two so that the dependent global loads (perm -> pointer table /
  // face info / TinvT) are in flight long before their results are needed:
  //   prefetchTile: global loads into registers;  commitTile: bulk copies + metadata / rotations into the stage
  constexpr int TPT = (FPT * 81 + (C::WT - 32) - 1) / (C::WT - 32);
  struct Ids {            // the faces this thread deals with as a loader: read from `perm` TWO tiles ahead
    long long mover;      // lane wt < 2 FPT of warp 0: face of slot wt / 2
    long long rot[TPT];   // the other warps: face whose rotation entry they fetch
  };
  struct Pre {
    long long face;
    const void* src;
    int vp, vm;
    float tv[TPT];
  };
  auto faceOf = [&](long long tile, int ff) -> long long {
    const long long slot = tile * FPT + ff;
    return (tile < nTiles && slot < p.n) ? (p.perm != nullptr ? p.perm[slot] : slot) : -1;
  };
  auto loadIds = [&](long long tile, Ids& ids) {
    ids.mover = wt < 2 * FPT ? faceOf(tile, wt >> 1) : -1;
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
      const int i = wt - 32 + t * (C::WT - 32);
      ids.rot[t] = (wt >= 32 && i < FPT * 81) ? faceOf(tile, i / 81) : -1;
    }
  };
  auto prefetchTile = [&](const Ids& ids, Pre& pre) {
    pre.face = ids.mover;
    pre.src = nullptr;
    pre.vp = pre.vm = 0;
    if (pre.face >= 0) {
      pre.src = (wt & 1) == 0 ? p.derPlus.at(pre.face) : p.derMinus.at(pre.face);
      const int i0 = p.faceInfo[3 * pre.face], i1 = p.faceInfo[3 * pre.face + 1], i2 = p.faceInfo[3 * pre.face + 2];
      pre.vp = i0;
      pre.vm = i1 + 4 * i2;
    }
#pragma unroll
    for (int t = 0; t < TPT; ++t) {
      const int i = wt - 32 + t * (C::WT - 32);
      pre.tv[t] = ids.rot[t] >= 0 ? static_cast<const float*>(p.tinvT)[81 * ids.rot[t] + (i - (i / 81) * 81)] : 0.f;
    }
  };
  auto commitTile = [&](int st, const Pre& pre) {
    float* const stage = sStage + st * stageFloats;
    float* const sTinv = stage + 2 * FPT * p.rawStride;
    int* const sMeta = reinterpret_cast<int*>(sTinv + FPT * 108);
    if (wt < 32) {   // the loader warp; lanes >= 2 FPT only take part in the byte count
      uint32_t bytes = 0;
      if (pre.face >= 0) {
        bytes = p.derSize * sizeof(float);
        tma_load(stage + wt * p.rawStride, pre.src, bytes, bar + st);
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
      }
      if (wt < 2 * FPT && (wt & 1) == 0) {
        const int ff = wt >> 1;
        sMeta[4 * ff + 0] = static_cast<int>(pre.face & 0xffffffffll);
        sMeta[4 * ff + 1] = static_cast<int>(pre.face >> 32);
        sMeta[4 * ff + 2] = pre.vp;
        sMeta[4 * ff + 3] = pre.vm;
      }
      if (wt == 0) {
        mbar_arrive_expect_tx(bar + st, bytes);   // total bytes of the stage on ONE arrival
      }
    } else {
#pragma unroll
      for (int t = 0; t < TPT; ++t) {
        const int i = wt - 32 + t * (C::WT - 32);
        if (i < FPT * 81) {
          const int ff = i / 81, e = i - ff * 81;
          sTinv[ff * 108 + (e / 9) * 12 + e % 9] = pre.tv[t];   // column qq of TinvT in 12 floats: three 128-bit reads
        }
      }
    }
  };

  uint32_t phaseB = 0, phaseMma = 0;
  int loaded0 = -1, loaded1 = -1;   // operator variants resident in shared memory (plus, minus); issuer warp

  // ---- helpers (SIDE is a compile-time constant everywhere: no dynamically indexed local arrays)
  // Taylor sums of the rotated rows: own 4-column chunks (k4 = 2 j + h) of A row (f, s, q)
  auto taylor = [&](const float* stage, int side, bool rowLive, float (&a)[OC][4]) {
#pragma unroll
    for (int j = 0; j < OC; ++j) {
      a[j][0] = a[j][1] = a[j][2] = a[j][3] = 0.f;
    }
    if (rowLive) {
      const float* fam = stage + (2 * f + side) * p.rawStride;
#pragma unroll
      for (int d = 0; d < O; ++d) {
        const float cf = static_cast<float>(p.coeffs[s][d]);
        const float* col = fam + p.derOffset[d] + q * p.derLd[d];
#pragma unroll
        for (int j = 0; j < OC; ++j) {
          const int k4 = 2 * j + h;
          if (4 * (2 * j) < G::derRows(d) && 4 * k4 < G::derRows(d)) {
            float t4[4];
            ld4(t4, col + 4 * k4);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float v = 4 * k4 + i < G::derRows(d) ? t4[i] : 0.f;
              a[j][i] = fmaf(cf, v, a[j][i]);
            }
          }
        }
      }
    }
  };
  // split into tf32 hi / lo and write to tensor memory: hi at column k, lo at column K + k
  auto storeA = [&](const float (&a)[OC][4]) {
#pragma unroll
    for (int j = 0; j < OC; ++j) {
      const int k4 = 2 * j + h;
      if (k4 < CH) {   // warp-uniform
        float hi[4], lo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          hi[i] = drtc::tf32_rna(a[j][i]);
          lo[i] = a[j][i] - hi[i];
        }
        drtc::tmem_st4(tmemA + laneBits + 4 * k4, hi[0], hi[1], hi[2], hi[3]);
        drtc::tmem_st4(tmemA + laneBits + K + 4 * k4, lo[0], lo[1], lo[2], lo[3]);
      }
    }
    drtc::tmem_st_wait();
    umma::fence_before_sync();
  };
  // the issuer warp: operator variant v of `side` into shared memory if it is not there, then the 3 x K/8 MMAs into D
  auto issue = [&](int side, int v, uint32_t dcol) {
    if (issuer) {
      int& loaded = side == 0 ? loaded0 : loaded1;
      const bool reload = loaded != v;   // warp-uniform
      if (drtc::elect_one()) {
        float* const bh = sBh + side * C::B_FLOATS;
        float* const bl = sBl + side * C::B_FLOATS;
        if (reload) {
          // the MMAs that read the old operand were waited for (barrier [3]) before this point
          mbar_arrive_expect_tx(bar + 2, 2 * C::B_FLOATS * sizeof(float));
          tma_load(bh, GD::v3CanonHi(v), C::B_FLOATS * sizeof(float), bar + 2);
          tma_load(bl, GD::v3CanonLo(v), C::B_FLOATS * sizeof(float), bar + 2);
          mbar_wait(bar + 2, phaseB);
        }
        umma::fence_after_sync();
        const uint32_t idesc = umma::instr_desc_tf32(128, N);
        const uint64_t dbh = umma::smem_desc(umma::smem_addr(bh), C::LBO, C::SBO);
        const uint64_t dbl = umma::smem_desc(umma::smem_addr(bl), C::LBO, C::SBO);
#pragma unroll
        for (int ks = 0; ks < K / 8; ++ks) {
          const uint64_t ob = static_cast<uint64_t>((ks * 2 * C::LBO) >> 4);
          drtc::mma_tf32_ts(dcol, tmemA + K + 8 * ks, dbh + ob, idesc, ks > 0);
          drtc::mma_tf32_ts(dcol, tmemA + 8 * ks, dbl + ob, idesc, true);
          drtc::mma_tf32_ts(dcol, tmemA + 8 * ks, dbh + ob, idesc, true);
        }
        umma::commit(bar + 3);
      }
      __syncwarp();
      if (reload) {
        phaseB ^= 1;
        loaded = v;
      }
    }
  };
  auto waitMma = [&]() {
    mbar_wait(bar + 3, phaseMma);
    phaseMma ^= 1;
    umma::fence_after_sync();
  };
  // row (f, s, q) -> one contiguous row of QInterpolated, columns [0, H0) / [H0, PP) per thread of the row
  auto epilogue = [&](int side, bool mine, long long myFace, uint32_t dcol) {
    constexpr int PP = GD::kPP;                         // padded points of a row (52 at order 6)
    constexpr int H0 = ((PP / 4 + 1) / 2) * 4;
    constexpr int NCHK = H0 / 4;
    const int c0 = h == 0 ? 0 : H0;
    float r4[NCHK][4];
#pragma unroll
    for (int c = 0; c < NCHK; ++c) {
      if (c0 + 4 * c < N) {   // warp-uniform
        drtc::tmem_ld4(dcol + laneBits + c0 + 4 * c, r4[c]);
      }
    }
    umma::tmem_ld_wait();
    if (mine) {
      float* dst = static_cast<float*>(side == 0 ? p.qPlus : p.qMinus) + ((myFace * O + s) * 9 + q) * static_cast<long long>(p.ppad);
#pragma unroll
      for (int c = 0; c < NCHK; ++c) {
        if (c0 + 4 * c < PP) {
          *reinterpret_cast<float4*>(dst + c0 + 4 * c) = make_float4(r4[c][0], r4[c][1], r4[c][2], r4[c][3]);
        }
      }
    }
  };

  Pre pre;
  Ids ids;
  if (tile0 < nTiles) {
    loadIds(tile0, ids);
    prefetchTile(ids, pre);
    commitTile(0, pre);
    loadIds(tile0 + tileStride, ids);
  }
  uint32_t it = 0;
  for (long long tile = tile0; tile < nTiles; tile += tileStride, ++it) {
    const int st = it & 1;
    float* const stage = sStage + st * stageFloats;
    float* const sTinv = stage + 2 * FPT * p.rawStride;
    const int* const sMeta = reinterpret_cast<const int*>(sTinv + FPT * 108);
    const bool haveNext = tile + tileStride < nTiles;
    if (haveNext) {
      prefetchTile(ids, pre);   // global loads only: consumed by commitTile below, under the MMAs
      loadIds(tile + 2 * tileStride, ids);
    }
    mbar_wait(bar + st, (it >> 1) & 1);
    drtc::worker_sync(w);   // metadata and rotations written by the loader threads one iteration ago are visible to all
    // the tile's faces (-1: none) and operator variants; `uniform`: all faces of the tile share both variants (the
    // caller sorts the faces by variant, so this is the normal case)
    long long myFace = -1;
    int myV0 = 0, myV1 = 0;
    const int v0 = sMeta[2], v1 = sMeta[3];
    bool uniform = true;
    int nFaces = 0;
#pragma unroll
    for (int ff = 0; ff < FPT; ++ff) {
      const long long fc = (static_cast<long long>(sMeta[4 * ff + 1]) << 32) | static_cast<unsigned>(sMeta[4 * ff]);
      if (fc >= 0) {
        ++nFaces;
        uniform = uniform && sMeta[4 * ff + 2] == v0 && sMeta[4 * ff + 3] == v1;
      }
      if (ff == f) {
        myFace = fc;
        myV0 = sMeta[4 * ff + 2];
        myV1 = sMeta[4 * ff + 3];
      }
    }
    const bool rowLive = rowUsed && myFace >= 0;
    // ---- phase 1: rotate every row of every dQ(d) of both sides in place, row <- row . TinvT
    //      (one item = 4 consecutive rows of one dQ(d): every column of TinvT is read once per item, 128-bit accesses)
    for (int item = wt; item < 2 * nFaces * TG; item += C::WT) {
      const int fs = item / TG;
      int r = item - fs * TG;
      int d = 0;
      while (r >= (G::derRows(d) + 3) / 4) {
        r -= (G::derRows(d) + 3) / 4;
        ++d;
      }
      const float* T = sTinv + (fs >> 1) * 108;
      float* rowp = stage + fs * p.rawStride + p.derOffset[d] + 4 * r;
      const int ld = p.derLd[d];
      float x[9][4];
#pragma unroll
      for (int c = 0; c < 9; ++c) {
        ld4(x[c], rowp + c * ld);
      }
#pragma unroll
      for (int qq = 0; qq < 9; ++qq) {
        float t[12];
        ld4(*reinterpret_cast<float(*)[4]>(t), T + 12 * qq);
        ld4(*reinterpret_cast<float(*)[4]>(t + 4), T + 12 * qq + 4);
        ld4(*reinterpret_cast<float(*)[4]>(t + 8), T + 12 * qq + 8);
        float y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float a = x[0][i] * t[0];
#pragma unroll
          for (int c = 1; c < 9; ++c) {
            a = fmaf(x[c][i], t[c], a);
          }
          y[i] = a;
        }
        st4(rowp + qq * ld, y[0], y[1], y[2], y[3]);   // in place: every input column is in registers already
      }
    }
    drtc::worker_sync(w);

    if (uniform) {
      // ---- fast path: the CUDA-core work of side 1 and the stores of side 0 run under the MMAs
      float a[OC][4];
      taylor(stage, 0, rowLive, a);
      storeA(a);
      drtc::worker_sync(w);
      issue(0, v0, tmemD0);
      if (haveNext) {
        commitTile(st ^ 1, pre);   // the other stage is free (its tile was finished behind a barrier)
      }
      taylor(stage, 1, rowLive, a);
      waitMma();
      storeA(a);
      drtc::worker_sync(w);
      issue(1, v1, tmemD1);
      epilogue(0, rowLive, myFace, tmemD0);
      waitMma();
      epilogue(1, rowLive, myFace, tmemD1);
    } else {
      // ---- faces with different operator variants in one tile: one MMA pass per distinct variant, sequentially
      if (haveNext) {
        commitTile(st ^ 1, pre);
      }
#pragma unroll 1
      for (int side = 0; side < 2; ++side) {
        float a[OC][4];
        taylor(stage, side, rowLive, a);
        storeA(a);
        drtc::worker_sync(w);
#pragma unroll 1
        for (int pass = 0; pass < FPT; ++pass) {
          const long long fc = (static_cast<long long>(sMeta[4 * pass + 1]) << 32) | static_cast<unsigned>(sMeta[4 * pass]);
          if (fc < 0) {
            continue;
          }
          const int v = sMeta[4 * pass + 2 + side];
          bool first = true;
          for (int ff = 0; ff < pass; ++ff) {
            if (sMeta[4 * ff + 2 + side] == v) {
              first = false;   // handled by an earlier pass (earlier slots of a tile are always occupied)
            }
          }
          if (!first) {
            continue;
          }
          issue(side, v, tmemD0);
          waitMma();
          epilogue(side, rowLive && (side == 0 ? myV0 : myV1) == v, myFace, tmemD0);
          umma::fence_before_sync();
          drtc::worker_sync(w);   // every row of D has been read: the next pass may overwrite it
          umma::fence_after_sync();
        }
      }
    }
    umma::fence_before_sync();
    drtc::worker_sync(w);   // the stage and both accumulators are free again
    umma::fence_after_sync();
  }
  umma::fence_before_sync();
  __syncthreads();
  if (tid < 32) {
    umma::tmem_dealloc(*tmemSlot, C::TMEM_COLS);
  }
}

}  // namespace ssb
