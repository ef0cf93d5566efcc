// godunov_dr on the 5th-generation tensor cores (tcgen05 + TMEM), single precision.
//
// Reference: kernels::DynamicRupture::batchedSpaceTimeInterpolation (src/Kernels/DynamicRupture.cpp:110-168),
//   QInterpolated±[s] = V3mTo2n(side, rel) . (sum_d coeff[s][d] dQ±(d)) . TinvT     (codegen/kernels/dynamic_rupture.py:108-172)
//
// V3mTo2n is the one DENSE operator of the path ((O+1)^2 x B, 49 x 56 at order 6, 16 (side, relation) variants): as
// immediate-operand SIMT code it is 43 k instructions per kernel (680 KB, far beyond the instruction caches) and runs at
// 19 % of the HBM ceiling.  Here it is a GEMM on the tensor cores:
//
//   D[(face, s, q)][point] = sum_k A[(face, s, q)][k] . B[point][k]
//     A = the rotated Taylor sums, rows = (face, time point s, quantity q): 9 O rows per face side, FPT faces per 128-row
//         tile, K = basis functions; computed by the CUDA cores (rotation with TinvT FIRST -- it commutes with the left
//         operator -- then the Taylor sums) and written to shared memory in the canonical K-major layout (umma.cuh),
//     B = V3mTo2n(variant), precomputed in that layout by the generator (gen_dr.py), fetched by bulk copies when the
//         variant changes (faces arrive sorted by variant),
//   as THREE kind::tf32 products per K step (a b ~ ah bh + ah bl + al bh, hi/lo split: relative error ~ 2^-21, the 1e-5
//   bar of single precision needs more than plain TF32's 2^-11), accumulated in TMEM (128 lanes x N columns, fp32),
//   read back with tcgen05.ld: a thread owns row r = one (face, s, q) and stores its (O+1)^2 point values as one
//   contiguous row of QInterpolated.
//
// One persistent CTA per SM, warp-specialised:
//   warp 8 (producer): face ids / operator variants / rotation matrices of the next tile into shared memory, both
//     derivative families of every face by bulk copies (two stages: the next tile is in flight while this one is computed);
//   warps 0..7 (consumers): rotate the families in place (one (k, d) row per thread), Taylor sums -> A operand,
//     thread 0 issues the MMAs, all eight warps drain TMEM (lane quadrant = warp % 4, column half = warp / 4).
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
  static constexpr int A_FLOATS = 16 * KG * 36;    // 128 rows
  static constexpr int B_FLOATS = GD::kTcOperand;  // N rows
  static constexpr int TMEM_COLS = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
  static constexpr int CONSUMERS = 256;
  static constexpr int THREADS = CONSUMERS + 32;
  static constexpr int SH = (O + 1) / 2;           // time points of the first half
  __host__ __device__ static constexpr int totalRows() {
    int t = 0;
    for (int d = 0; d < O; ++d) {
      t += G::derRows(d);
    }
    return t;
  }
  static_assert(FPT >= 1 && 2 * FPT * B <= CONSUMERS && FPT * totalRows() <= CONSUMERS, "thread mappings");
  static_assert(B_FLOATS == (N / 8) * KG * 36, "generator and kernel disagree on the operand layout");
  static_assert((N / 2) % 8 == 0, "each column half is read in chunks of 8");
  // shared memory (floats): A hi | A lo | B hi (plus, minus) | B lo (plus, minus) | per stage: families, rotations, meta
  static constexpr int offAh = 0;
  static constexpr int offAl = offAh + A_FLOATS;
  static constexpr int offBh = offAl + A_FLOATS;
  static constexpr int offBl = offBh + 2 * B_FLOATS;
  static constexpr int offStage = offBl + 2 * B_FLOATS;
  static constexpr int META = 4 * FPT + 4;         // floats per stage for {face id (2 words), variant plus, variant minus} per face
  __host__ __device__ static int stageFloats(int rawStride) { return 2 * FPT * rawStride + FPT * 84 + META; }
  __host__ __device__ static size_t bytes(int rawStride) {
    return (size_t(offStage) + 2 * size_t(stageFloats(rawStride))) * sizeof(float) + 128;
  }
};

__device__ __forceinline__ void consumer_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

template <typename G, typename GD>
__global__ void __launch_bounds__((DrTcCfg<G, GD>::THREADS), 1) k_dr_interp_tc(const DrInterpParams p) {
  using C = DrTcCfg<G, GD>;
  constexpr int O = C::O, B = C::B, K = C::K, N = C::N, KG = C::KG, RPF = C::RPF, FPT = C::FPT, SH = C::SH;
  constexpr int TR = C::totalRows();
  extern __shared__ __align__(128) unsigned char smemRaw[];
  float* const smem = reinterpret_cast<float*>(smemRaw);
  float* const sAh = smem + C::offAh;
  float* const sAl = smem + C::offAl;
  float* const sBh = smem + C::offBh;
  float* const sBl = smem + C::offBl;
  const int stageFloats = C::stageFloats(p.rawStride);
  float* const sStage = smem + C::offStage;
  // bar[0..1]: stage full, bar[2..3]: stage empty, bar[4]: B operand, bar[5]: MMA done
  uint64_t* const bar = reinterpret_cast<uint64_t*>(sStage + 2 * stageFloats);
  uint32_t* const tmemSlot = reinterpret_cast<uint32_t*>(bar + 6);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  // zero the A operand once: rows of missing faces, rows >= FPT * RPF and the K padding are never written
  for (int i = tid; i < 2 * C::A_FLOATS; i += C::THREADS) {
    sAh[i] = 0.f;
  }
  if (tid == 0) {
    mbar_init(bar + 0, 1); mbar_init(bar + 1, 1);
    mbar_init(bar + 2, 1); mbar_init(bar + 3, 1);
    mbar_init(bar + 4, 1); mbar_init(bar + 5, 1);
    mbar_fence_init();
  }
  if (warp == 0) {
    umma::tmem_alloc(tmemSlot, C::TMEM_COLS);
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = *tmemSlot;
  const long long nTiles = (p.n + FPT - 1) / FPT;

  if (warp == 8) {
    // =================================================== producer ===================================================
    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++it) {
      const int st = it & 1;
      float* const stage = sStage + st * stageFloats;
      float* const sTinv = stage + 2 * FPT * p.rawStride;
      int* const sMeta = reinterpret_cast<int*>(sTinv + FPT * 84);
      if (it >= 2) {
        mbar_wait(bar + 2 + st, ((it >> 1) - 1) & 1);   // the consumers are done with this stage
      }
      // every lane knows the faces of the tile; lanes 0 .. 2 FPT - 1 move one (face, side) family each
      long long fcs[FPT];
#pragma unroll
      for (int f = 0; f < FPT; ++f) {
        const long long slot = tile * FPT + f;
        fcs[f] = slot < p.n ? (p.perm != nullptr ? p.perm[slot] : slot) : -1;
      }
      long long face = -1;
#pragma unroll
      for (int f = 0; f < FPT; ++f) {
        if (lane < 2 * FPT && (lane >> 1) == f) {
          face = fcs[f];
        }
      }
      uint32_t bytes = 0;
      if (face >= 0) {
        const int side = lane & 1;
        tma_load(stage + lane * p.rawStride, side == 0 ? p.derPlus.at(face) : p.derMinus.at(face), p.derSize * sizeof(float),
                 bar + st);
        bytes = p.derSize * sizeof(float);
      }
      if (lane < 2 * FPT && (lane & 1) == 0) {
        const int f = lane >> 1;
        int vp = 0, vm = 0;
        if (face >= 0) {
          const int i0 = p.faceInfo[3 * face], i1 = p.faceInfo[3 * face + 1], i2 = p.faceInfo[3 * face + 2];
          vp = i0;
          vm = i1 + 4 * i2;
        }
        sMeta[4 * f + 0] = static_cast<int>(face & 0xffffffffll);
        sMeta[4 * f + 1] = static_cast<int>(face >> 32);
        sMeta[4 * f + 2] = vp;
        sMeta[4 * f + 3] = vm;
      }
      // the faces' 9x9 rotations
      for (int i = lane; i < FPT * 81; i += 32) {
        const int f = i / 81, e = i - f * 81;
        long long fc = -1;
#pragma unroll
        for (int g = 0; g < FPT; ++g) {
          fc = g == f ? fcs[g] : fc;
        }
        sTinv[f * 84 + e] = fc >= 0 ? static_cast<const float*>(p.tinvT)[81 * fc + e] : 0.f;
      }
      // total bytes of the stage, then ONE arrival (release: the generic writes above become visible with it)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        bytes += __shfl_xor_sync(0xffffffffu, bytes, o);
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_expect_tx(bar + st, bytes);
      }
    }
  } else {
    // =================================================== consumers ==================================================
    const uint32_t idesc = umma::instr_desc_tf32(128, N);
    uint32_t phaseB = 0, phaseMma = 0;
    int loadedVariant[2] = {-1, -1};   // meaningful in thread 0 only
    // phase 1 role: one (face, d, k) row of the families
    int r1f = 0, r1d = 0, r1k = 0;
    const bool r1 = tid < FPT * TR;
    if (r1) {
      r1f = tid / TR;
      int r = tid - r1f * TR;
      while (r >= G::derRows(r1d)) {
        r -= G::derRows(r1d);
        ++r1d;
      }
      r1k = r;
    }
    // phase 2 role: (face, basis function k, half of the time points)
    const bool r2 = tid < 2 * FPT * B;
    const int r2h = tid / (FPT * B);
    const int r2f = (tid - r2h * FPT * B) / B;
    const int r2k = tid - r2h * FPT * B - r2f * B;
    // epilogue role
    const int erow = (warp & 3) * 32 + lane;
    const int ecol0 = (warp >> 2) * (N / 2);
    const int ef = erow / RPF;
    const int erem = erow - ef * RPF;
    const int es = erem / 9, eq = erem - (erem / 9) * 9;

    uint32_t it = 0;
    for (long long tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++it) {
      const int st = it & 1;
      float* const stage = sStage + st * stageFloats;
      float* const sTinv = stage + 2 * FPT * p.rawStride;
      const int* const sMeta = reinterpret_cast<const int*>(sTinv + FPT * 84);
      mbar_wait(bar + st, (it >> 1) & 1);
      long long face[FPT];
      int variant[2][FPT];
#pragma unroll
      for (int f = 0; f < FPT; ++f) {
        face[f] = (static_cast<long long>(sMeta[4 * f + 1]) << 32) | static_cast<unsigned>(sMeta[4 * f]);
        variant[0][f] = sMeta[4 * f + 2];
        variant[1][f] = sMeta[4 * f + 3];
      }
      // ---- phase 1: rotate every row of every dQ(d) of both sides in place, row <- row . TinvT
      if (r1 && face[r1f] >= 0) {
        const float* T = sTinv + r1f * 84;
#pragma unroll
        for (int side = 0; side < 2; ++side) {
          float* row = stage + (2 * r1f + side) * p.rawStride + p.derOffset[r1d] + r1k;
          const int ld = p.derLd[r1d];
          float x[9], y[9];
#pragma unroll
          for (int c = 0; c < 9; ++c) {
            x[c] = row[c * ld];
          }
#pragma unroll
          for (int q = 0; q < 9; ++q) {
            float a = x[0] * T[9 * q];
#pragma unroll
            for (int c = 1; c < 9; ++c) {
              a = fmaf(x[c], T[c + 9 * q], a);
            }
            y[q] = a;
          }
#pragma unroll
          for (int c = 0; c < 9; ++c) {
            row[c * ld] = y[c];
          }
        }
      }
      consumer_sync();

#pragma unroll 1
      for (int side = 0; side < 2; ++side) {
        // ---- phase 2: Taylor sums of the rotated rows -> A operand (hi / lo), rows (f, s, q), column k
        if (r2 && face[r2f] >= 0) {
          const float* fam = stage + (2 * r2f + side) * p.rawStride;
          const int s0 = r2h == 0 ? 0 : SH;
          const int ns = r2h == 0 ? SH : O - SH;
          float acc[SH][9];
#pragma unroll
          for (int s = 0; s < SH; ++s) {
#pragma unroll
            for (int q = 0; q < 9; ++q) {
              acc[s][q] = 0.f;
            }
          }
#pragma unroll
          for (int d = 0; d < O; ++d) {
            if (r2k < G::derRows(d)) {
              float y[9];
#pragma unroll
              for (int q = 0; q < 9; ++q) {
                y[q] = fam[p.derOffset[d] + q * p.derLd[d] + r2k];
              }
#pragma unroll
              for (int s = 0; s < SH; ++s) {
                const float cf = s < ns ? static_cast<float>(p.coeffs[s0 + s][d]) : 0.f;
#pragma unroll
                for (int q = 0; q < 9; ++q) {
                  acc[s][q] = fmaf(cf, y[q], acc[s][q]);
                }
              }
            }
          }
          const int colOff = (r2k >> 2) * 36 + (r2k & 3);
#pragma unroll
          for (int s = 0; s < SH; ++s) {
            if (s < ns) {
#pragma unroll
              for (int q = 0; q < 9; ++q) {
                const int row = r2f * RPF + (s0 + s) * 9 + q;
                const int off = (row >> 3) * (KG * 36) + (row & 7) * 4 + colOff;
                const float hi = umma::tf32_hi(acc[s][q]);
                sAh[off] = hi;
                sAl[off] = umma::tf32_lo(acc[s][q], hi);
              }
            }
          }
        }
        fence_async_smem();   // generic-proxy writes of A -> visible to the tensor core
        consumer_sync();
        if (side == 1 && tid == 0) {
          mbar_arrive(bar + 2 + st);   // both families of the stage are consumed: the producer may refill it
        }

        // ---- one MMA pass per distinct operator variant among the faces of the tile (faces arrive sorted: usually one)
#pragma unroll 1
        for (int pass = 0; pass < FPT; ++pass) {
          if (face[pass] < 0) {
            continue;
          }
          bool first = true;
#pragma unroll
          for (int f = 0; f < FPT; ++f) {
            if (f < pass && face[f] >= 0 && variant[side][f] == variant[side][pass]) {
              first = false;   // handled by an earlier pass
            }
          }
          if (!first) {
            continue;
          }
          const int v = variant[side][pass];
          if (tid == 0) {
            float* const bh = sBh + side * C::B_FLOATS;
            float* const bl = sBl + side * C::B_FLOATS;
            if (loadedVariant[side] != v) {
              // the MMAs that read the old operand were waited for (barrier [5]) before this point
              mbar_arrive_expect_tx(bar + 4, 2 * C::B_FLOATS * sizeof(float));
              tma_load(bh, GD::v3CanonHi(v), C::B_FLOATS * sizeof(float), bar + 4);
              tma_load(bl, GD::v3CanonLo(v), C::B_FLOATS * sizeof(float), bar + 4);
              mbar_wait(bar + 4, phaseB);
              phaseB ^= 1;
              loadedVariant[side] = v;
            }
            umma::fence_after_sync();
            const uint32_t aH = umma::smem_addr(sAh), aL = umma::smem_addr(sAl);
            const uint32_t bH = umma::smem_addr(bh), bL = umma::smem_addr(bl);
#pragma unroll
            for (int ks = 0; ks < K / 8; ++ks) {
              const uint32_t off = ks * 2 * C::LBO;
              const uint64_t dah = umma::smem_desc(aH + off, C::LBO, C::SBO), dal = umma::smem_desc(aL + off, C::LBO, C::SBO);
              const uint64_t dbh = umma::smem_desc(bH + off, C::LBO, C::SBO), dbl = umma::smem_desc(bL + off, C::LBO, C::SBO);
              umma::mma_tf32(tmem, dah, dbh, idesc, ks > 0);
              umma::mma_tf32(tmem, dah, dbl, idesc, true);
              umma::mma_tf32(tmem, dal, dbh, idesc, true);
            }
            umma::commit(bar + 5);
          }
          mbar_wait(bar + 5, phaseMma);
          phaseMma ^= 1;
          umma::fence_after_sync();
          // ---- epilogue: row erow = (f, s, q) -> half of one contiguous row of QInterpolated
          const bool mine = erow < FPT * RPF && face[ef < FPT ? ef : 0] >= 0 && variant[side][ef < FPT ? ef : 0] == v;
          float* dst = nullptr;
          if (mine) {
            dst = static_cast<float*>(side == 0 ? p.qPlus : p.qMinus) +
                  ((face[ef] * O + es) * 9 + eq) * static_cast<long long>(p.ppad);
          }
#pragma unroll
          for (int c = 0; c < N / 2; c += 8) {
            const int c0 = ecol0 + c;
            uint32_t r[8];
            umma::tmem_ld8(tmem + ((static_cast<uint32_t>(warp & 3) * 32u) << 16) + c0, r);
            umma::tmem_ld_wait();
            if (mine) {
              if (c0 < p.ppad) {
                *reinterpret_cast<float4*>(dst + c0) = make_float4(__uint_as_float(r[0]), __uint_as_float(r[1]),
                                                                    __uint_as_float(r[2]), __uint_as_float(r[3]));
              }
              if (c0 + 4 < p.ppad) {
                *reinterpret_cast<float4*>(dst + c0 + 4) = make_float4(__uint_as_float(r[4]), __uint_as_float(r[5]),
                                                                        __uint_as_float(r[6]), __uint_as_float(r[7]));
              }
            }
          }
          umma::fence_before_sync();
          consumer_sync();   // every row of D has been read: the next pass may overwrite it, the next side may rewrite A
          umma::fence_after_sync();
        }
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) {
    umma::tmem_dealloc(tmem, C::TMEM_COLS);
  }
}

}  // namespace ssb
