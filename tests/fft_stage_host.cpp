// Runs the inverse z pass's in-place DIF stages (smk_ztile.cuh: c2r_stages + c2r_last_butterfly, i.e. the code
// c2r_z_kernel executes between its load and its store) thread by thread on the host and compares every line with a
// float64 inverse DFT.  Checks the index logic of the padded layout and of the three twiddle sources (table, split,
// compile-time constants) without a GPU.  Built and run by tests/test_fft_stage_cpu.py with g++ -Itests/host_stub.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "smk_ztile.cuh"

using namespace smk;

template <class CT, int S, int SEND, int LINES, int NT, int LP>
static void run_stages(float2* sm, const float2* tw) {
  if constexpr (S < SEND) {
    for (unsigned tid = 0; tid < (unsigned)NT; ++tid) {   // a barrier separates the stages on the GPU
      threadIdx.x = tid;
      c2r_stages<CT, S, S + 1, LINES, NT, LP>(sm, tw);
    }
    run_stages<CT, S + 1, SEND, LINES, NT, LP>(sm, tw);
  }
}

// Shared-memory wavefronts of one stage's loads: the k-th 8-byte load of the 32 threads of a warp is one request; each
// half-warp needs as many passes as the most loaded of the 16 bank pairs (distinct addresses only).  Returns the
// average over the stage's requests (2.0 = conflict free).  Same dif_stage instantiation as c2r_stages.
template <class CT, int S, int LINES, int NT, int LP>
static double stage_wavefronts(float2* sm, const float2* tw, long long* tw_requests, long long* tw_wavefronts) {
  std::vector<std::vector<int>> log(NT);
  std::vector<std::vector<const void*>> glog(NT);
  for (unsigned tid = 0; tid < (unsigned)NT; ++tid) {
    threadIdx.x = tid;
    smk_host_ldg_log = &glog[tid];
    auto ld = [&](int line, int ppos, int, int) { log[tid].push_back(line * LP + ppos); return sm[line * LP + ppos]; };
    auto st = [&](int, int, float2) {};
    dif_stage<typename CT::P, S, true, LINES, NT, OUT_INPLACE, decltype(ld), decltype(st), NoPre, 0, CT::BLK,
              z_tw_mode<typename CT::P, S, LINES, NT>(), (SMK_Z_TW >= 2) ? Z_SPLIT_ROW : 0, CT::JFAST>(ld, st, tw, 2);
  }
  smk_host_ldg_log = nullptr;
  // twiddle loads: one L1 wavefront per distinct 128-byte line of a warp's request
  long long greq = 0, glines = 0;
  for (int w = 0; w < NT / 32; ++w)
    for (size_t k = 0; k < glog[w * 32].size(); ++k, ++greq) {
      std::vector<uintptr_t> lines;
      for (int l = 0; l < 32; ++l) {
        if (k >= glog[w * 32 + l].size()) continue;
        const uintptr_t a = ((uintptr_t)glog[w * 32 + l][k] - (uintptr_t)tw) / 128;
        bool dup = false;
        for (uintptr_t x : lines) dup |= (x == a);
        if (!dup) lines.push_back(a);
      }
      glines += (long long)lines.size();
    }
  *tw_requests = greq;
  *tw_wavefronts = glines;
  long long req = 0, wf = 0;
  for (int w = 0; w < NT / 32; ++w)
    for (size_t k = 0; k < log[w * 32].size(); ++k, ++req)
      for (int h = 0; h < 2; ++h) {
        int cnt[16] = {0}, seen[16][16];
        for (int l = 0; l < 16; ++l) {
          const std::vector<int>& v = log[w * 32 + h * 16 + l];
          if (k >= v.size()) continue;
          const int a = v[k], b = a & 15;
          bool dup = false;
          for (int i = 0; i < cnt[b]; ++i) dup |= (seen[b][i] == a);
          if (!dup) seen[b][cnt[b]++] = a;
        }
        int mx = 0;
        for (int b = 0; b < 16; ++b) mx = cnt[b] > mx ? cnt[b] : mx;
        wf += mx;
      }
  return req ? (double)wf / req : 0.;
}

template <int M>
static int check() {
  using CT = C2RTraits<M>;
  using ZT = typename CT::ZT;
  using P = typename CT::P;
  if constexpr (!CT::FUSE) {
    printf("M=%d not fused: skipped\n", M);
    return 0;
  } else {
    constexpr int LINES = ZT::LINES, NT = ZT::NT, LP = CT::LP, RL = CT::RL, NB = CT::NB;
    std::vector<float2> sm((size_t)LINES * LP, make_float2(0.f, 0.f)), tw;
    z_twiddle_table(2 * M, tw);   // what smk_capi.cu uploads for NZ = 2 M
    std::vector<double> zr(LINES * M), zi(LINES * M);
    srand(1234 + M);
    for (int l = 0; l < LINES; ++l)
      for (int n = 0; n < M; ++n) {
        zr[l * M + n] = rand() / (double)RAND_MAX - 0.5;
        zi[l * M + n] = rand() / (double)RAND_MAX - 0.5;
        sm[(size_t)l * LP + CT::idx(n)] = make_float2((float)zr[l * M + n], (float)zi[l * M + n]);
      }
    long long treq[2] = {0, 0}, twf[2] = {0, 0};
    const double wf0 = stage_wavefronts<CT, 0, LINES, NT, LP>(sm.data(), tw.data(), &treq[0], &twf[0]);
    double wf1 = 0.;
    if constexpr (P::S > 2) wf1 = stage_wavefronts<CT, 1, LINES, NT, LP>(sm.data(), tw.data(), &treq[1], &twf[1]);
    std::vector<float2> outs((size_t)LINES * M, make_float2(1e30f, 1e30f));
    if constexpr (CT::TAIL) {
      // the kernel's path: stages [0, S-2) through shared memory, the last two in registers (c2r_tail)
      run_stages<CT, 0, P::S - 2, LINES, NT, LP>(sm.data(), tw.data());
      auto emit = [&](int line, int n, float2 z) { outs[(size_t)line * M + n] = z; };
      for (unsigned tid = 0; tid < (unsigned)NT; ++tid) {
        threadIdx.x = tid;
        c2r_tail<CT, LINES, NT, LP>(sm.data(), emit);
      }
    } else {
      run_stages<CT, 0, P::S - 1, LINES, NT, LP>(sm.data(), tw.data());
      for (int l = 0; l < LINES; ++l)
        for (int n0 = 0; n0 < NB; ++n0) {
          float2 v[RL];
          c2r_last_butterfly<CT>(sm.data() + (size_t)l * LP, n0, v);
          for (int q = 0; q < RL; ++q) outs[(size_t)l * M + n0 + q * NB] = v[q];
        }
    }
    double worst = 0., scale = 0.;
    for (int l = 0; l < LINES; ++l) {
      const float2* out = outs.data() + (size_t)l * M;
      for (int n = 0; n < M; ++n) {
        double re = 0., im = 0.;
        for (int k = 0; k < M; ++k) {
          const double a = 2. * M_PI * (double)((long long)k * n % M) / M, c = cos(a), s = sin(a);
          re += zr[l * M + k] * c - zi[l * M + k] * s;
          im += zr[l * M + k] * s + zi[l * M + k] * c;
        }
        worst = fmax(worst, fmax(fabs(out[n].x - re), fabs(out[n].y - im)));
        scale = fmax(scale, fmax(fabs(re), fabs(im)));
      }
    }
    const double rel = worst / scale;
    printf("M=%d lines=%d threads=%d tail=%d jfast=%d tw_modes=%d,%d lds_wavefronts_per_request=%.2f,%.2f "
           "twiddle_loads_per_tile=%lld,%lld twiddle_wavefronts_per_tile=%lld,%lld (tile data: %d) max_err/max=%.3g %s\n",
           M, LINES, NT, (int)CT::TAIL, (int)CT::JFAST, z_tw_mode<P, 0, LINES, NT>(), P::S > 2 ? z_tw_mode<P, 1, LINES, NT>() : -1, wf0, wf1, treq[0],
           treq[1], twf[0], twf[1], LINES * M * 8 / 128, rel, rel < 2e-6 ? "ok" : "FAIL");
    return rel < 2e-6 ? 0 : 1;
  }
}

// z_tile_fft (forward r2c tiles, and the inverse tiles that are not fused): every stage for all threads, the stage's
// input taken from a copy of the tile (what the barrier inside the re-sorting last stage guarantees on the GPU)
template <int M, bool INV, int S>
static void run_tile_stages(std::vector<float2>& sm, const float2* tw) {
  using ZT = ZTraits<M, INV>;
  if constexpr (S < ZT::P::S) {
    const std::vector<float2> src = sm;
    for (unsigned tid = 0; tid < (unsigned)ZT::NT; ++tid) {
      threadIdx.x = tid;
      z_tile_stage<M, INV, S>(src.data(), sm.data(), tw);
    }
    run_tile_stages<M, INV, S + 1>(sm, tw);
  }
}

template <int M, bool INV>
static int check_tile() {
  using ZT = ZTraits<M, INV>;
  constexpr int LINES = ZT::LINES, LP = ZT::LP;
  std::vector<float2> sm((size_t)LINES * LP, make_float2(0.f, 0.f)), tw;
  z_twiddle_table(2 * M, tw);   // what smk_capi.cu uploads for NZ = 2 M
  std::vector<double> zr(LINES * M), zi(LINES * M);
  srand(99 + M);
  for (int l = 0; l < LINES; ++l)
    for (int n = 0; n < M; ++n) {
      zr[l * M + n] = rand() / (double)RAND_MAX - 0.5;
      zi[l * M + n] = rand() / (double)RAND_MAX - 0.5;
      sm[(size_t)l * LP + ZT::idx(n)] = make_float2((float)zr[l * M + n], (float)zi[l * M + n]);
    }
  run_tile_stages<M, INV, 0>(sm, tw.data());
  double worst = 0., scale = 0.;
  for (int l = 0; l < LINES; ++l)
    for (int k = 0; k < M; ++k) {
      double re = 0., im = 0.;
      for (int n = 0; n < M; ++n) {
        const double a = (INV ? 2. : -2.) * M_PI * (double)((long long)k * n % M) / M, c = cos(a), sn = sin(a);
        re += zr[l * M + n] * c - zi[l * M + n] * sn;
        im += zr[l * M + n] * sn + zi[l * M + n] * c;
      }
      const float2 got = sm[(size_t)l * LP + ZT::nat(k)];
      worst = fmax(worst, fmax(fabs(got.x - re), fabs(got.y - im)));
      scale = fmax(scale, fmax(fabs(re), fabs(im)));
    }
  const double rel = worst / scale;
  printf("tile M=%d %s lines=%d threads=%d perm=%d max_err/max=%.3g %s\n", M, INV ? "inverse" : "forward", LINES, ZT::NT,
         (int)ZT::PERM, rel, rel < 2e-6 ? "ok" : "FAIL");
  return rel < 2e-6 ? 0 : 1;
}

// The strided (x / y) passes: tile [point][line] in shared memory, first stage from "global" memory, last stage back to
// it in natural order -- the dif_stage instantiations of strided_tile (smk_boxes.cu), with that file's tile shapes
// (StridedTraits: LINES kz columns, NT threads) and first-stage batch B0 (0 = all tasks at once, 2 = the k-factor passes).
template <class P, int S, int SEND, bool INV, int LINES, int NT>
static void run_strided_middle(float2* sm, const float2* tw) {
  if constexpr (S < SEND) {
    auto ld_s = [&](int line, int pos, int, int) { return sm[pos * LINES + line]; };
    auto st_s = [&](int line, int pos, float2 val) { sm[pos * LINES + line] = val; };
    for (unsigned tid = 0; tid < (unsigned)NT; ++tid) {
      threadIdx.x = tid;
      dif_stage<P, S, INV, LINES, NT, OUT_INPLACE>(ld_s, st_s, tw, 1);
    }
    run_strided_middle<P, S + 1, SEND, INV, LINES, NT>(sm, tw);
  }
}

template <int N, int LINES, int NT, bool INV, int B0>
static int check_strided() {
  using P = typename PlanFor<N>::type;
  std::vector<float2> in((size_t)N * LINES), out((size_t)N * LINES, make_float2(1e30f, 1e30f)), sm((size_t)N * LINES), tw(N);
  for (int k = 0; k < N; ++k) tw[k] = make_float2((float)cos(2. * M_PI * k / N), (float)-sin(2. * M_PI * k / N));
  srand(7 + N + LINES);
  for (auto& v : in) v = make_float2((float)(rand() / (double)RAND_MAX - 0.5), (float)(rand() / (double)RAND_MAX - 0.5));
  auto ld_g = [&](int line, int n, int, int) { return in[(size_t)n * LINES + line]; };
  auto st_g = [&](int line, int k, float2 val) { out[(size_t)k * LINES + line] = val; };
  auto ld_s = [&](int line, int pos, int, int) { return sm[pos * LINES + line]; };
  auto st_s = [&](int line, int pos, float2 val) { sm[pos * LINES + line] = val; };
  if constexpr (P::S == 1) {
    for (unsigned tid = 0; tid < (unsigned)NT; ++tid) {
      threadIdx.x = tid;
      dif_stage<P, 0, INV, LINES, NT, OUT_NATURAL>(ld_g, st_g, tw.data(), 1);
    }
  } else {
    for (unsigned tid = 0; tid < (unsigned)NT; ++tid) {
      threadIdx.x = tid;
      dif_stage<P, 0, INV, LINES, NT, OUT_INPLACE, decltype(ld_g), decltype(st_s), NoPre, B0>(ld_g, st_s, tw.data(), 1);
    }
    run_strided_middle<P, 1, P::S - 1, INV, LINES, NT>(sm.data(), tw.data());
    for (unsigned tid = 0; tid < (unsigned)NT; ++tid) {
      threadIdx.x = tid;
      dif_stage<P, P::S - 1, INV, LINES, NT, OUT_NATURAL>(ld_s, st_g, tw.data(), 1);
    }
  }
  // float64 DFT of every column through the table of exp(i 2 pi j / N), j = k n mod N
  std::vector<double> cs(N), sn(N);
  for (int j = 0; j < N; ++j) { cs[j] = cos(2. * M_PI * j / N); sn[j] = (INV ? 1. : -1.) * sin(2. * M_PI * j / N); }
  double worst = 0., scale = 0.;
  for (int l = 0; l < LINES; ++l)
    for (int k = 0; k < N; ++k) {
      double re = 0., im = 0.;
      for (int n = 0; n < N; ++n) {
        const int j = (int)((long long)k * n % N);
        const float2 x = in[(size_t)n * LINES + l];
        re += x.x * cs[j] - x.y * sn[j];
        im += x.x * sn[j] + x.y * cs[j];
      }
      const float2 got = out[(size_t)k * LINES + l];
      worst = fmax(worst, fmax(fabs(got.x - re), fabs(got.y - im)));
      scale = fmax(scale, fmax(fabs(re), fabs(im)));
    }
  const double rel = worst / scale;
  printf("strided N=%d %s lines=%d threads=%d batch=%d stages=%d max_err/max=%.3g %s\n", N, INV ? "inverse" : "forward", LINES,
         NT, B0, P::S, rel, rel < 3e-6 ? "ok" : "FAIL");
  return rel < 3e-6 ? 0 : 1;
}

int main() {
  int bad = 0;
  // strided passes: (N, LINES, NT) of StridedTraits<N> for the lengths of BASELINE's configurations and the small ones
  bad += check_strided<16, 16, 64, false, 0>() + check_strided<64, 16, 64, true, 0>() + check_strided<256, 16, 128, true, 2>();
  bad += check_strided<512, 16, 256, false, 0>() + check_strided<512, 16, 256, true, 2>();
  bad += check_strided<1024, 16, 512, true, 0>() + check_strided<2048, 8, 512, true, 2>();
  bad += check_strided<2560, 8, 640, false, 0>() + check_strided<2560, 8, 640, true, 2>();
  // every tile shape the C ABI dispatches (NZ/2): forward tiles, the inverse tiles that are not fused, the fused inverse
  bad += check_tile<4, false>() + check_tile<8, false>() + check_tile<12, false>() + check_tile<16, false>();
  bad += check_tile<32, false>() + check_tile<48, false>() + check_tile<64, false>() + check_tile<128, false>();
  bad += check_tile<256, false>() + check_tile<512, false>() + check_tile<768, false>();
  bad += check_tile<4, true>() + check_tile<8, true>() + check_tile<12, true>() + check_tile<16, true>();
  bad += check_tile<64, true>() + check_tile<512, true>();
  bad += check<32>() + check<48>() + check<128>() + check<256>() + check<384>() + check<768>() + check<2048>();
  return bad;
}
