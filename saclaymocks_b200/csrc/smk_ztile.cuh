// Tile transform of the contiguous (z) passes: LINES lines of M complex points in shared memory, layout [line][point].
// Kept apart from the kernels (smk_boxes.cu) so that tests/fft_stage_host.cpp can run the very same stage code thread by
// thread on the host (tests/test_fft_stage_cpu.py).
#pragma once
#include <math.h>
#include <vector>
#include "smk_fft.cuh"

namespace smk {

// ------------------------------------------------------------------ contiguous (z) pass
// INVERSE selects the tile height of the NZ = 1536 passes, which differs between the two directions.
template <int M, bool INVERSE = false>
struct ZTraits {
  using P = typename PlanFor<M>::type;
  // Lines per tile of the NZ = 1536 z passes, measured on B200 (tools/ab_check.sh): smaller tiles mean more, smaller CTAs
  // per SM (8 lines: four of 128 threads and 50 KB; 4 lines: eight of 64 threads), whose load / transform / store phases
  // interleave better.  16 -> 8 lines: c2r z 0.787 -> 0.772 ms, r2c z + Philox 1.186 -> 1.147 ms; 8 -> 4 lines (after the
  // twiddle loads had gone): c2r z 0.690 -> 0.674 ms but r2c z 1.139 -> 1.172 ms, hence 4 for the inverse pass only
  // (with the register-resident tail, 4 against 8 lines: 0.633 against 0.664 ms).
#ifndef SMK_Z_LINES
#define SMK_Z_LINES 8
#endif
#ifndef SMK_Z_LINES_C2R
#define SMK_Z_LINES_C2R 4
#endif
  static constexpr int LINES = (M > 1024) ? 4 : ((M == 768) ? (INVERSE ? SMK_Z_LINES_C2R : SMK_Z_LINES) : 16);
  static constexpr int NT_ = (M % 3 == 0) ? LINES * M / 48 / 32 * 32 : LINES * M / 32;
  static constexpr int NT = NT_ < 64 ? 64 : (NT_ > 512 ? 512 : NT_);
  // PERM: two-stage plan R0.R1 with a radix-32 first stage, run without the re-sorting last stage (which would need
  // all of a thread's butterflies in registers at once).  Natural index k then sits at position (k % R0) R1 + k / R0;
  // with one float2 of padding after every R1 points, consecutive k are 25 float2 apart for R1 = 24: the permuted
  // reads of the store / post-processing loops (lane = k) stay free of bank conflicts.
  static constexpr bool PERM = (P::S == 2 && P::radix(0) >= 32);
  static constexpr int R0 = P::radix(0), R1 = PERM ? P::radix(1) : 1;
  static constexpr int LP_ = PERM ? M + M / R1 : M;
  // pitch: the lanes of a half-warp are LINES lines x 16 / LINES consecutive positions of one stage; with a pitch of
  // 16 / LINES (mod 16) float2 they fall on 16 different bank pairs (LINES == 16: any odd pitch)
  static constexpr int pitch(int least) {
    int lp = least;
    while (lp % 16 != (16 / LINES) % 16) ++lp;
    return lp;
  }
  static constexpr int LP = (LINES == 16) ? LP_ + 1 - LP_ % 2 : pitch(LP_);
  __device__ static __forceinline__ int idx(int p) { return PERM ? p + p / R1 : p; }         // padded position
  __device__ static __forceinline__ int nat(int k) { return PERM ? idx((k % R0) * R1 + k / R0) : k; }   // where output k sits
};

template <int M> using ZFwd = ZTraits<M, false>;   // tiles of the forward (r2c) pass
template <int M> using ZInv = ZTraits<M, true>;    // tiles of the inverse (c2r) pass

// Where a non-final stage of the z tile takes its twiddles from (smk_fft.cuh TW_*).  A warp of the z pass holds several
// lines, so a table load fans out over a few addresses and costs one L1 wavefront each -- 15 loads per radix-16
// butterfly, about a quarter of the kernel's wavefronts.  Stage 0 loads one twiddle per output and thread and derives
// those of the thread's other butterflies by a constant rotation; later stages use compile-time constants when a thread
// can take all the butterflies of one block (NZ = 1536: 16.16.3, the 3 butterflies of each 48-point block).
#ifndef SMK_Z_TW
#define SMK_Z_TW 2   // 0: table loads everywhere, 1: split / constant twiddles, 2: and stage 0 reads the compact table
#endif
// The z passes' twiddle table: W_NZ[k] = exp(-2 pi i k / NZ), k < NZ, followed by the compact first-stage table
// T[q][j] = W_M^(q j) (M = NZ / 2, q < Z_SPLIT_ROWS, j < Z_SPLIT_ROW) read by TW_SPLIT.
constexpr int Z_SPLIT_ROW = 64, Z_SPLIT_ROWS = 32;
inline void z_twiddle_table(int nz, std::vector<float2>& h) {
  const double PI = 3.14159265358979323846;
  h.resize((size_t)nz + Z_SPLIT_ROWS * Z_SPLIT_ROW);
  for (int k = 0; k < nz; ++k) {
    const double a = -2.0 * PI * (double)k / (double)nz;
    h[k] = make_float2((float)cos(a), (float)sin(a));
  }
  const int m = nz / 2;
  for (int q = 0; q < Z_SPLIT_ROWS; ++q)
    for (int j = 0; j < Z_SPLIT_ROW; ++j) {
      const double a = -2.0 * PI * (double)((long long)q * j % m) / (double)m;
      h[(size_t)nz + q * Z_SPLIT_ROW + j] = make_float2((float)cos(a), (float)sin(a));
    }
}
template <class P, int STAGE, int LINES, int NT>
__host__ __device__ constexpr int z_tw_mode() {
  constexpr int R = P::radix(STAGE), MQ = P::sub(STAGE) / R, NB = P::N / R;
  constexpr int TPT = (NB * LINES + NT - 1) / NT, JSTEP = NT / LINES;
  if (!SMK_Z_TW) return TW_TABLE;
  if (STAGE == 0) return TPT > 1 ? TW_SPLIT : TW_TABLE;
  return ((NB * LINES) % NT == 0 && TPT == MQ && NB == JSTEP * MQ) ? TW_CONST : TW_TABLE;
}

// Stage S of the M-point transform of every line of the tile ([line][idx(point)], pitch LP; natural-order input, output
// k of a line ends at ZTraits<M, INV>::nat(k)).  Reads src and writes dst: the same shared-memory tile on the GPU; the host
// emulation (tests/fft_stage_host.cpp) hands the re-sorting last stage a copy as src, standing in for its barrier.
template <int M, bool INV, int S>
__device__ __forceinline__ void z_tile_stage(const float2* src, float2* dst, const float2* __restrict__ tw) {
  using ZT = ZTraits<M, INV>;
  using P = typename ZT::P;
  constexpr int LINES = ZT::LINES, NT = ZT::NT, LP = ZT::LP;
  if constexpr (ZT::PERM) {
    auto ld = [&](int line, int pos, int, int) { return src[line * LP + ZT::idx(pos)]; };
    auto st = [&](int line, int pos, float2 val) { dst[line * LP + ZT::idx(pos)] = val; };
    dif_stage<P, S, INV, LINES, NT, OUT_INPLACE, decltype(ld), decltype(st), NoPre, 1>(ld, st, tw, 2);
  } else {
    auto ld = [&](int line, int pos, int, int) { return src[line * LP + pos]; };
    auto st = [&](int line, int pos, float2 val) { dst[line * LP + pos] = val; };
    if constexpr (S == P::S - 1) {
      dif_stage<P, S, INV, LINES, NT, OUT_RESORT>(ld, st, tw, 2);
    } else {
      // unpadded tile: only the first stage can do without per-task table loads (TW_CONST needs the padded layout)
      constexpr int TW = (S == 0) ? z_tw_mode<P, 0, LINES, NT>() : TW_TABLE;
      dif_stage<P, S, INV, LINES, NT, OUT_INPLACE, decltype(ld), decltype(st), NoPre, 0, 0, TW,
                (SMK_Z_TW >= 2) ? Z_SPLIT_ROW : 0>(ld, st, tw, 2);
    }
  }
}

// all stages, a barrier after each
template <int M, bool INV, int S = 0>
__device__ __forceinline__ void z_tile_fft(float2* sm, const float2* __restrict__ tw) {
  if constexpr (S < PlanFor<M>::type::S) {
    z_tile_stage<M, INV, S>(sm, sm, tw);
    __syncthreads();
    z_tile_fft<M, INV, S + 1>(sm, tw);
  }
}

// Inverse z pass with the last DIF stage fused into the store (FUSE, plans of >= 2 stages): the last stage has no
// twiddles and its butterfly b = pos(n0) / RL produces the natural outputs n0 + q M/RL, so a warp whose lanes take
// CONSECUTIVE n0 runs it straight from shared memory into coalesced global stores -- the re-sort (one shared-memory
// write + read of the tile) and the separate store loop's read disappear (8 -> 6 passes over the tile in shared
// memory).  Consecutive n0 read positions M/R0 apart: one float2 of padding after every M/R0 positions makes that
// stride odd in units of 8 bytes, i.e. conflict free for the 16 lanes of a half-warp.  Where a thread can hold a whole
// block of the stage before last (TAIL, e.g. NZ = 1536), the last two stages run in registers instead: 4 passes.
template <int M>
struct C2RTraits {
  using ZT = ZTraits<M, true>;
  using P = typename ZT::P;
  static constexpr int R0 = P::radix(0), RL = P::radix(P::S - 1), BLK = M / R0, NB = M / RL;
  static constexpr bool FUSE = (P::S >= 2) && !ZT::PERM && (BLK % 2 == 0) && (BLK % RL == 0) && (M >= 32);
  static constexpr int LP_ = M + R0;
  static constexpr int LP = FUSE ? (ZT::LINES == 16 ? (LP_ + 1 - LP_ % 2) : ZT::pitch(LP_)) : ZT::LP;
  __device__ static __forceinline__ int idx(int p) { return FUSE ? p + p / BLK : ZT::idx(p); }
  // TAIL: the last TWO stages run in registers (c2r_tail below).  In TW_CONST mode the thread of block b holds the
  // block's RL butterflies of stage S-2, i.e. all R x RL points of the block -- which are exactly the inputs of the
  // block's last-stage butterflies, so these follow without another trip through shared memory and their outputs go
  // from the registers to global memory (4 instead of 6 passes over the tile, one barrier less).
#ifndef SMK_Z_TAIL
#define SMK_Z_TAIL 1
#endif
  static constexpr int S1 = P::S >= 3 ? P::S - 2 : 0;
  static constexpr bool TAIL =
      SMK_Z_TAIL && FUSE && P::S >= 3 && z_tw_mode<P, S1, ZT::LINES, ZT::NT>() == TW_CONST && BLK % P::sub(S1) == 0;
  // lanes along the butterflies of one line (instead of along the lines) when a half-warp fits into a line: the tail's
  // stores are then 128-byte runs; the padded layout's odd strides keep shared memory conflict free either way
  static constexpr bool JFAST = TAIL && (ZT::NT / ZT::LINES) % 16 == 0;
};

// stages [S0, S1) in place through the padded index, a barrier after each
template <class CT, int S0, int S1, int LINES, int NT, int LP>
__device__ __forceinline__ void c2r_stages(float2* sm, const float2* __restrict__ tw) {
  if constexpr (S0 < S1) {
    auto ld = [&](int line, int ppos, int, int) { return sm[line * LP + ppos]; };          // padded positions
    auto st = [&](int line, int ppos, float2 val) { sm[line * LP + ppos] = val; };
    dif_stage<typename CT::P, S0, true, LINES, NT, OUT_INPLACE, decltype(ld), decltype(st), NoPre, 0, CT::BLK,
              z_tw_mode<typename CT::P, S0, LINES, NT>(), (SMK_Z_TW >= 2) ? Z_SPLIT_ROW : 0, CT::JFAST>(ld, st, tw, 2);
    __syncthreads();
    c2r_stages<CT, S0 + 1, S1, LINES, NT, LP>(sm, tw);
  }
}

// last (twiddle-free) stage of the fused inverse z pass, run per natural output index n0 < NB straight out of shared
// memory: v[q] = output n0 + q NB of the line
template <class CT>
__device__ __forceinline__ void c2r_last_butterfly(const float2* row, int n0, float2 (&v)[CT::RL]) {
  const int pb = CT::P::pos(n0);
#pragma unroll
  for (int t = 0; t < CT::RL; ++t) v[t] = row[CT::idx(pb + t)];
  Butterfly<CT::RL, true>::run(v);
}

// The last two stages of the fused inverse z pass in registers (C2RTraits::TAIL).  Thread (line, b) loads the R x RL
// points of block b (position RL t + i = input t of the stage-(S-2) butterfly i), runs the RL radix-R butterflies with
// their compile-time twiddles W_MB^(i q), then the R last-stage butterflies over (v[0][m], .., v[RL-1][m]), and hands
// every result to emit(line, n, z) with its natural output index n = nat(b MB) + m N/MB + q2 N/RL.
template <class CT, int LINES, int NT, int LP, class Emit>
__device__ __forceinline__ void c2r_tail(const float2* sm, Emit emit) {
  using P = typename CT::P;
  constexpr int S1 = CT::S1, R = P::radix(S1), MB = P::sub(S1), RL = CT::RL, JSTEP = NT / LINES;
  static_assert(MB == R * RL && P::N == JSTEP * MB, "one block of the stage before last per thread");
  const int line = CT::JFAST ? threadIdx.x / JSTEP : threadIdx.x % LINES;
  const int b = CT::JFAST ? threadIdx.x % JSTEP : threadIdx.x / LINES;
  const float2* blk = sm + line * LP + b * MB + b / (CT::BLK / MB);   // padded start of block b
  float2 v[RL][R];
#pragma unroll
  for (int i = 0; i < RL; ++i)
#pragma unroll
    for (int t = 0; t < R; ++t) v[i][t] = blk[i + t * RL];
  constexpr TwConst<MB, 1, RL, R> K{};
#pragma unroll
  for (int i = 0; i < RL; ++i) {
    Butterfly<R, true>::run(v[i]);
    if (i > 0) {
#pragma unroll
      for (int q = 1; q < R; ++q) v[i][q] = twc<true>(v[i][q], K.c[i][q], K.s[i][q]);
    }
  }
  const int nb = P::nat(b * MB);
#pragma unroll
  for (int m = 0; m < R; ++m) {
    float2 u[RL];
#pragma unroll
    for (int t = 0; t < RL; ++t) u[t] = v[t][m];
    Butterfly<RL, true>::run(u);
#pragma unroll
    for (int q2 = 0; q2 < RL; ++q2) emit(line, nb + m * (P::N / MB) + q2 * (P::N / RL), u[q2]);
  }
}

}  // namespace smk
