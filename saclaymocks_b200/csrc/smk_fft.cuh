// Shared-memory-staged mixed-radix FFT building blocks (sm_100a).
//
// Every 1-D transform of the hot path (make_boxes.py:53,87; merge_spectra.py:311,321 in the
// reference, done there by FFTW) is computed here as an in-place decimation-in-frequency
// FFT on a tile of LINES independent lines held in shared memory.  Threads map to
// (line, butterfly) with the line index fastest, so that all lanes of a half-warp execute the
// SAME butterfly on DIFFERENT lines: every shared-memory access of a stage is then a row of
// consecutive float2 (strided passes, layout [point][line]) or a column with odd pitch
// (contiguous pass, layout [line][point], pitch M+1) -- conflict free for any index pattern,
// which is what lets the digit-reversed in-place algorithm be used without transposes.
// Twiddles are float2 tables computed in fp64 on the host (W[k] = exp(-2 pi i k / N)).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SMK_HD __host__ __device__ __forceinline__
#ifndef SMK_PACKED
#define SMK_PACKED 1
#endif

namespace smk {

// Complex add / subtract are the bulk of a butterfly.  On the device they are ONE packed instruction each
// (add.rn.f32x2 / sub.rn.f32x2 -> FADD2, sm_100+: the two float32 additions with the same IEEE rounding as the scalar
// pair, so results do not change), which takes the instruction count of the issue-bound z passes down.
SMK_HD float2 cadd(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && SMK_PACKED
  float2 d;
  asm("{ .reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
#else
  return make_float2(a.x + b.x, a.y + b.y);
#endif
}
SMK_HD float2 csub(float2 a, float2 b) {
#if defined(__CUDA_ARCH__) && SMK_PACKED
  float2 d;
  asm("{ .reg .b64 ra, rb, rd;\n\t"
      "mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5};\n\t"
      "sub.rn.f32x2 rd, ra, rb;\n\t"
      "mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
#else
  return make_float2(a.x - b.x, a.y - b.y);
#endif
}
SMK_HD float2 cmul(float2 a, float2 b) {
  return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
SMK_HD float2 cconj(float2 a) { return make_float2(a.x, -a.y); }
SMK_HD float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
// multiply by -i (forward) or +i (inverse)
template <bool INV>
SMK_HD float2 mul_mi(float2 a) {
  return INV ? make_float2(-a.y, a.x) : make_float2(a.y, -a.x);
}

// a / b for float32 from a reciprocal estimate r ~ 1/b and one fused residual correction (Markstein): correctly
// rounded except for rare near-ties, at 3 instructions instead of the ~10 of the IEEE division subroutine.  The
// reference divides in float32 (box /= N, k_i k_j / kk); the boxes are compared at 1e-5 relative L2.
__device__ __forceinline__ float rcp_approx(float b) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
  return r;
}
__device__ __forceinline__ float fdiv_fast(float a, float b, float r) {
  float q = a * r;
  float e = fmaf(-q, b, a);
  return fmaf(e, r, q);
}
// both components of a float2 at once (FMUL2 + 2 FFMA2; b and r go in as lane-broadcast operands): the same three
// roundings per component as fdiv_fast
__device__ __forceinline__ float2 fdiv_fast2(float2 a, float b, float r) {
#if defined(__CUDA_ARCH__) && SMK_PACKED
  float2 d;
  asm("{ .reg .b64 ra, rr, rnb, rq, re, rd;\n\t"
      ".reg .f32 nb;\n\t"
      "neg.f32 nb, %4;\n\t"
      "mov.b64 ra, {%2, %3}; mov.b64 rr, {%5, %5}; mov.b64 rnb, {nb, nb};\n\t"
      "mul.rn.f32x2 rq, ra, rr;\n\t"
      "fma.rn.f32x2 re, rq, rnb, ra;\n\t"
      "fma.rn.f32x2 rd, re, rr, rq;\n\t"
      "mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b), "f"(r));
  return d;
#else
  return make_float2(fdiv_fast(a.x, b, r), fdiv_fast(a.y, b, r));
#endif
}
// a * b + c per component (FFMA2)
__device__ __forceinline__ float2 cfma(float2 a, float2 b, float2 c) {
#if defined(__CUDA_ARCH__) && SMK_PACKED
  float2 d;
  asm("{ .reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
#else
  return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
#endif
}

// ---------------------------------------------------------------- radix butterflies
// y_q = sum_t x_t w^(t q),  w = exp(-2 pi i / R) (forward) or its conjugate (INV)
template <int R, bool INV>
struct Butterfly;

template <bool INV>
struct Butterfly<2, INV> {
  static SMK_HD void run(float2 (&v)[2]) {
    float2 a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};

template <bool INV>
struct Butterfly<3, INV> {
  static SMK_HD void run(float2 (&v)[3]) {
    const float S3 = 0.86602540378443864676f;
    float2 s = cadd(v[1], v[2]), d = csub(v[1], v[2]);
    float2 m = make_float2(v[0].x - 0.5f * s.x, v[0].y - 0.5f * s.y);
    float2 n = mul_mi<INV>(cscale(d, S3));   // -i n (fwd), +i n (inv)
    v[0] = cadd(v[0], s);
    v[1] = cadd(m, n);
    v[2] = csub(m, n);
  }
};

template <bool INV>
struct Butterfly<4, INV> {
  static SMK_HD void run(float2 (&v)[4]) {
    float2 t0 = cadd(v[0], v[2]), t1 = csub(v[0], v[2]);
    float2 t2 = cadd(v[1], v[3]), t3 = mul_mi<INV>(csub(v[1], v[3]));
    v[0] = cadd(t0, t2);
    v[2] = csub(t0, t2);
    v[1] = cadd(t1, t3);
    v[3] = csub(t1, t3);
  }
};

template <bool INV>
struct Butterfly<5, INV> {
  static SMK_HD void run(float2 (&v)[5]) {
    const float C1 = 0.30901699437494742410f, C2 = -0.80901699437494742410f;
    const float S1 = 0.95105651629515357212f, S2 = 0.58778525229247312917f;
    float2 a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
    float2 b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
    float2 m1 = make_float2(v[0].x + C1 * a1.x + C2 * a2.x, v[0].y + C1 * a1.y + C2 * a2.y);
    float2 m2 = make_float2(v[0].x + C2 * a1.x + C1 * a2.x, v[0].y + C2 * a1.y + C1 * a2.y);
    float2 n1 = mul_mi<INV>(make_float2(S1 * b1.x + S2 * b2.x, S1 * b1.y + S2 * b2.y));
    float2 n2 = mul_mi<INV>(make_float2(S2 * b1.x - S1 * b2.x, S2 * b1.y - S1 * b2.y));
    v[0] = cadd(v[0], cadd(a1, a2));
    v[1] = cadd(m1, n1);
    v[4] = csub(m1, n1);
    v[2] = cadd(m2, n2);
    v[3] = csub(m2, n2);
  }
};

template <bool INV>
struct Butterfly<8, INV> {
  static SMK_HD void run(float2 (&v)[8]) {
    const float H = 0.70710678118654752440f;
    float2 e[4] = {v[0], v[2], v[4], v[6]};
    float2 o[4] = {v[1], v[3], v[5], v[7]};
    Butterfly<4, INV>::run(e);
    Butterfly<4, INV>::run(o);
    // o_q *= w8^q
    float2 o1 = INV ? make_float2((o[1].x - o[1].y) * H, (o[1].x + o[1].y) * H)
                    : make_float2((o[1].x + o[1].y) * H, (o[1].y - o[1].x) * H);
    float2 o2 = mul_mi<INV>(o[2]);
    float2 o3 = INV ? make_float2((-o[3].x - o[3].y) * H, (o[3].x - o[3].y) * H)
                    : make_float2((o[3].y - o[3].x) * H, (-o[3].x - o[3].y) * H);
    v[0] = cadd(e[0], o[0]);
    v[4] = csub(e[0], o[0]);
    v[1] = cadd(e[1], o1);
    v[5] = csub(e[1], o1);
    v[2] = cadd(e[2], o2);
    v[6] = csub(e[2], o2);
    v[3] = cadd(e[3], o3);
    v[7] = csub(e[3], o3);
  }
};

template <bool INV>
struct Butterfly<16, INV> {
  // 4 x 4 decomposition: y[q1 + 4 q2] = sum_b w4^(b q2) w16^(b q1) [ sum_a x[4a+b] w4^(a q1) ]
  static SMK_HD float2 tw16(float2 v, float c, float s) {   // v * (c - i s) fwd, (c + i s) inv
    return INV ? make_float2(v.x * c - v.y * s, v.y * c + v.x * s) : make_float2(v.x * c + v.y * s, v.y * c - v.x * s);
  }
  static SMK_HD void run(float2 (&v)[16]) {
    const float H = 0.70710678118654752440f, C1 = 0.92387953251128675613f, S1 = 0.38268343236508977173f;
    float2 u[4][4];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      float2 g[4] = {v[b], v[4 + b], v[8 + b], v[12 + b]};
      Butterfly<4, INV>::run(g);
#pragma unroll
      for (int q1 = 0; q1 < 4; ++q1) u[b][q1] = g[q1];
    }
    // twiddles w16^(b q1): (b,q1) = (1,1):1 (1,2):2 (1,3):3 (2,1):2 (2,2):4 (2,3):6 (3,1):3 (3,2):6 (3,3):9
    u[1][1] = tw16(u[1][1], C1, S1);
    u[1][2] = tw16(u[1][2], H, H);
    u[1][3] = tw16(u[1][3], S1, C1);
    u[2][1] = tw16(u[2][1], H, H);
    u[2][2] = mul_mi<INV>(u[2][2]);
    u[2][3] = tw16(u[2][3], -H, H);
    u[3][1] = tw16(u[3][1], S1, C1);
    u[3][2] = tw16(u[3][2], -H, H);
    u[3][3] = tw16(u[3][3], -C1, -S1);
#pragma unroll
    for (int q1 = 0; q1 < 4; ++q1) {
      float2 g[4] = {u[0][q1], u[1][q1], u[2][q1], u[3][q1]};
      Butterfly<4, INV>::run(g);
#pragma unroll
      for (int q2 = 0; q2 < 4; ++q2) v[q1 + 4 * q2] = g[q2];
    }
  }
};

// v * w^e with w = exp(-2 pi i / L) (forward) or its conjugate (INV); c = cos(2 pi e / L), s = sin(2 pi e / L)
template <bool INV>
SMK_HD float2 twc(float2 v, float c, float s) {
  return INV ? make_float2(v.x * c - v.y * s, v.y * c + v.x * s) : make_float2(v.x * c + v.y * s, v.y * c - v.x * s);
}

template <bool INV>
struct Butterfly<24, INV> {
  // 3 x 8 decomposition, x index t = 3a + b (a < 8, b < 3), output q = q1 + 8 q2 (q1 < 8, q2 < 3):
  //   y[q1 + 8 q2] = sum_b w3^(b q2) w24^(b q1) [ sum_a x[3a+b] w8^(a q1) ]
  static SMK_HD void run(float2 (&v)[24]) {
    float2 u[3][8];
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      float2 g[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) g[a] = v[3 * a + b];
      Butterfly<8, INV>::run(g);
#pragma unroll
      for (int q1 = 0; q1 < 8; ++q1) u[b][q1] = g[q1];
    }
    // cos/sin(2 pi e / 24), e = 0..14
    const float C[15] = {1.f, 0.96592582628906828675f, 0.86602540378443864676f, 0.70710678118654752440f, 0.5f,
                         0.25881904510252076235f, 0.f, -0.25881904510252076235f, -0.5f, -0.70710678118654752440f,
                         -0.86602540378443864676f, -0.96592582628906828675f, -1.f, -0.96592582628906828675f,
                         -0.86602540378443864676f};
    const float S[15] = {0.f, 0.25881904510252076235f, 0.5f, 0.70710678118654752440f, 0.86602540378443864676f,
                         0.96592582628906828675f, 1.f, 0.96592582628906828675f, 0.86602540378443864676f,
                         0.70710678118654752440f, 0.5f, 0.25881904510252076235f, 0.f, -0.25881904510252076235f, -0.5f};
#pragma unroll
    for (int q1 = 1; q1 < 8; ++q1) {
      u[1][q1] = twc<INV>(u[1][q1], C[q1], S[q1]);
      u[2][q1] = twc<INV>(u[2][q1], C[2 * q1], S[2 * q1]);
    }
#pragma unroll
    for (int q1 = 0; q1 < 8; ++q1) {
      float2 g[3] = {u[0][q1], u[1][q1], u[2][q1]};
      Butterfly<3, INV>::run(g);
#pragma unroll
      for (int q2 = 0; q2 < 3; ++q2) v[q1 + 8 * q2] = g[q2];
    }
  }
};

template <bool INV>
struct Butterfly<32, INV> {
  // 4 x 8 decomposition, x index t = 4a + b (a < 8, b < 4), output q = q1 + 8 q2 (q1 < 8, q2 < 4):
  //   y[q1 + 8 q2] = sum_b w4^(b q2) w32^(b q1) [ sum_a x[4a+b] w8^(a q1) ]
  static SMK_HD void run(float2 (&v)[32]) {
    float2 u[4][8];
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      float2 g[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) g[a] = v[4 * a + b];
      Butterfly<8, INV>::run(g);
#pragma unroll
      for (int q1 = 0; q1 < 8; ++q1) u[b][q1] = g[q1];
    }
    // cos/sin(2 pi e / 32), e = 0..21
    const float C[22] = {1.f, 0.98078528040323044913f, 0.92387953251128675613f, 0.83146961230254523708f,
                         0.70710678118654752440f, 0.55557023301960222474f, 0.38268343236508977173f,
                         0.19509032201612826785f, 0.f, -0.19509032201612826785f, -0.38268343236508977173f,
                         -0.55557023301960222474f, -0.70710678118654752440f, -0.83146961230254523708f,
                         -0.92387953251128675613f, -0.98078528040323044913f, -1.f, -0.98078528040323044913f,
                         -0.92387953251128675613f, -0.83146961230254523708f, -0.70710678118654752440f,
                         -0.55557023301960222474f};
    const float S[22] = {0.f, 0.19509032201612826785f, 0.38268343236508977173f, 0.55557023301960222474f,
                         0.70710678118654752440f, 0.83146961230254523708f, 0.92387953251128675613f,
                         0.98078528040323044913f, 1.f, 0.98078528040323044913f, 0.92387953251128675613f,
                         0.83146961230254523708f, 0.70710678118654752440f, 0.55557023301960222474f,
                         0.38268343236508977173f, 0.19509032201612826785f, 0.f, -0.19509032201612826785f,
                         -0.38268343236508977173f, -0.55557023301960222474f, -0.70710678118654752440f,
                         -0.83146961230254523708f};
#pragma unroll
    for (int b = 1; b < 4; ++b)
#pragma unroll
      for (int q1 = 1; q1 < 8; ++q1) u[b][q1] = twc<INV>(u[b][q1], C[b * q1], S[b * q1]);
#pragma unroll
    for (int q1 = 0; q1 < 8; ++q1) {
      float2 g[4] = {u[0][q1], u[1][q1], u[2][q1], u[3][q1]};
      Butterfly<4, INV>::run(g);
#pragma unroll
      for (int q2 = 0; q2 < 4; ++q2) v[q1 + 8 * q2] = g[q2];
    }
  }
};

// ---------------------------------------------------------------- plans
// A plan factorises N into up to five radices (1 = unused).  Stage s works in place on
// sub-transforms of size sub(s) = N / (R0..R(s-1)); after the last stage, position p holds
// the natural output index nat(p) (mixed-radix digit reversal).
template <int R0, int R1 = 1, int R2 = 1, int R3 = 1, int R4 = 1>
struct PlanT {
  static constexpr int S = (R0 > 1) + (R1 > 1) + (R2 > 1) + (R3 > 1) + (R4 > 1);
  static constexpr int N = R0 * R1 * R2 * R3 * R4;
  __host__ __device__ static constexpr int radix(int s) {
    return s == 0 ? R0 : s == 1 ? R1 : s == 2 ? R2 : s == 3 ? R3 : R4;
  }
  __host__ __device__ static constexpr int sub(int s) {
    int m = N;
    for (int i = 0; i < s; ++i) m /= radix(i);
    return m;
  }
  // natural output index of in-place position p
  __host__ __device__ static constexpr int nat(int p) {
    int k = 0, w = 1;
    for (int s = 0; s < S; ++s) {
      int m = sub(s) / radix(s);
      int q = p / m;
      p -= q * m;
      k += q * w;
      w *= radix(s);
    }
    return k;
  }
  // in-place position of natural index k
  __host__ __device__ static constexpr int pos(int k) {
    int p = 0;
    for (int s = 0; s < S; ++s) {
      int q = k % radix(s);
      k /= radix(s);
      p += q * (sub(s) / radix(s));
    }
    return p;
  }
};

template <int N>
struct PlanFor;
#define SMK_PLAN(N_, ...)          \
  template <>                      \
  struct PlanFor<N_> {             \
    using type = PlanT<__VA_ARGS__>; \
  };
SMK_PLAN(4, 4)
SMK_PLAN(8, 8)
SMK_PLAN(12, 4, 3)
SMK_PLAN(16, 4, 4)
SMK_PLAN(24, 8, 3)
SMK_PLAN(32, 8, 4)
SMK_PLAN(48, 4, 4, 3)
SMK_PLAN(64, 8, 8)
SMK_PLAN(96, 8, 4, 3)
SMK_PLAN(128, 8, 4, 4)
SMK_PLAN(256, 16, 16)
SMK_PLAN(384, 8, 4, 4, 3)
// 512 and 1024 run as two-stage plans with a radix-32 first stage: a strided pass then touches shared memory twice
// per element (first stage writes, last stage reads) instead of four times, loads half the twiddles and has one
// barrier less, at 64 payload registers per thread (2 CTAs of 256 threads per SM).  Measured on B200 against the
// three-stage plans 8.8.8 / 16.16.4 (-DSMK_PLAN_512=3 -DSMK_PLAN_1024=3, tools/plan_sweep.sh): 512-point passes
// y 0.616 -> 0.572 ms, x with table 0.854 -> 0.762 ms, forward y 0.646 -> 0.523 ms; 1024-point passes y 3.40 -> 2.85 ms,
// x with k-factors 3.51 -> 3.00 ms (x with table 4.83 -> 5.04 ms: the 32 weights spill at 128 registers).
#ifndef SMK_PLAN_512
#define SMK_PLAN_512 2
#endif
#ifndef SMK_PLAN_1024
#define SMK_PLAN_1024 2
#endif
#if SMK_PLAN_512 == 2
SMK_PLAN(512, 32, 16)
#else
SMK_PLAN(512, 8, 8, 8)
#endif
// 768 (the z pass, NZ = 1536): SMK_PLAN_768=2 selects 32.24, run by the z kernels without the re-sorting last stage
// (outputs stay in digit-reversed positions of a padded line and are read back permuted, smk_boxes.cu ZTraits).
// Measured on B200 (tools/z_sweep.sh): parity green, but c2r z 0.790 -> 0.827 ms -- two shared-memory round trips fewer
// do not pay for the half-idle second round of the radix-32 stage (384 butterflies on 256 threads) -- so 16.16.3 stays.
#ifndef SMK_PLAN_768
#define SMK_PLAN_768 3
#endif
#if SMK_PLAN_768 == 2
SMK_PLAN(768, 32, 24)
#else
SMK_PLAN(768, 16, 16, 3)
#endif
#if SMK_PLAN_1024 == 2
SMK_PLAN(1024, 32, 32)
#else
SMK_PLAN(1024, 16, 16, 4)
#endif
SMK_PLAN(2048, 16, 16, 8)
#ifndef SMK_PLAN_2560
#define SMK_PLAN_2560 4
#endif
#if SMK_PLAN_2560 == 3
SMK_PLAN(2560, 32, 16, 5)
#elif SMK_PLAN_2560 == 31
SMK_PLAN(2560, 16, 32, 5)
#else
SMK_PLAN(2560, 16, 16, 2, 5)
#endif
SMK_PLAN(4096, 16, 16, 16)
#undef SMK_PLAN

// ---------------------------------------------------------------- compile-time twiddles
// cos / sin of 2 pi num / den evaluated in constant expressions (double Taylor series after reduction to [-pi, pi]:
// absolute error below 1e-15, i.e. exact after rounding to float except for rare ties).
__host__ __device__ constexpr double cx_cos_sin(long long num, long long den, bool want_sin) {
  num %= den;
  if (2 * num > den) num -= den;
  const double x = 6.283185307179586476925286766559 * (double)num / (double)den, x2 = x * x;
  double term = want_sin ? x : 1.0, sum = term;
  for (int k = want_sin ? 1 : 0; k < 44; k += 2) {
    term *= -x2 / (double)((k + 1) * (k + 2));
    sum += term;
  }
  return sum;
}
// c[i][q], s[i][q] = cos, sin of 2 pi (q i STEP) / L for i < T, q < R
template <int L, int STEP, int T, int R>
struct TwConst {
  float c[T][R], s[T][R];
  __host__ __device__ constexpr TwConst() : c{}, s{} {
    for (int i = 0; i < T; ++i)
      for (int q = 0; q < R; ++q) {
        c[i][q] = (float)cx_cos_sin((long long)q * i * STEP, L, false);
        s[i][q] = (float)cx_cos_sin((long long)q * i * STEP, L, true);
      }
  }
};

// ---------------------------------------------------------------- one DIF stage
// Tile element (line, pos) is accessed through the functors:
//   ld(line, pos)        -> float2     (input position, pre-stage)
//   st(line, idx, value)
// OUT selects what the store functor receives as idx and when it runs:
//   OUT_INPLACE   idx = the in-place position (same set the task read);
//   OUT_NATURAL   last stage only: idx = natural output index (stage writes elsewhere,
//                 e.g. straight to global memory);
//   OUT_RESORT    last stage only: idx = natural output index in the SAME buffer, so a
//                 __syncthreads() separates all reads of the stage from all writes.
// tw is the table W_L, L = N * twmul.
enum { OUT_INPLACE = 0, OUT_NATURAL = 1, OUT_RESORT = 2 };

struct NoPre {
  __device__ __forceinline__ float2 operator()(int, int, int, int, float2 v) const { return v; }
};

// ld(line, pos, i, t): i = task slot of this thread within the current batch, t = input index of the butterfly (both compile-time after
// unrolling, so a caller can keep per-element side data in a register array indexed [i][t]);
// pre(line, pos, i, t, v) runs after ALL loads of the thread have been issued (fused multiply of the first stage).
// PADBLK > 0 (in-place stages only): the tile is stored with one element of padding after every PADBLK = N / R0
// positions (position p at p + p / PADBLK); the stage hands ld / st the PADDED position, computed from the block index
// once per butterfly instead of a division per element.
// TW selects where the twiddles W_sub^(o q) of a non-final stage come from.  With several lines per warp the lanes of
// a table load read a few scattered addresses (one per butterfly offset o), each its own L1 wavefront, which on the
// shared-memory-bound z passes costs as much as the data:
//   TW_TABLE  one table load per output and task (the general case);
//   TW_SPLIT  (stage 0, where o = j0 + i JSTEP for the thread's task i) one load W^(q j0) per output, shared by the
//             thread's tasks: task i multiplies it by the compile-time constant W^(q i JSTEP).  With SPLIT_ROW > 0 the
//             load comes from a compact table T[q][j0] = W_N^(q j0) (rows of SPLIT_ROW entries, stored behind W_L at
//             tw + L): the lanes of a request then read neighbouring entries, one wavefront instead of one per j0;
//   TW_CONST  the thread takes the MQ butterflies of ONE block b = j0 (o = i is a compile-time value): no loads at
//             all.  Needs MQ tasks per thread; tasks of neighbouring threads are then a whole sub-transform apart, so
//             the tile needs the padded layout to stay free of bank conflicts.
enum { TW_TABLE = 0, TW_SPLIT = 1, TW_CONST = 2 };

template <class P, int STAGE, bool INV, int LINES, int NT, int OUT, class Load, class Store, class Pre = NoPre,
          int BATCH = 0, int PADBLK = 0, int TW = TW_TABLE, int SPLIT_ROW = 0, bool JFAST = false>
__device__ __forceinline__ void dif_stage(Load ld, Store st, const float2* __restrict__ tw, int twmul,
                                          Pre pre = Pre()) {
  constexpr int R = P::radix(STAGE);
  constexpr int M = P::sub(STAGE);
  constexpr int MQ = M / R;
  constexpr int NB = P::N / R;
  constexpr int NTASK = NB * LINES;
  constexpr bool LAST = (STAGE == P::S - 1);
  static_assert(OUT == OUT_INPLACE || LAST, "natural-order store only on the last stage");
  static_assert(NT % LINES == 0, "threads per block must be a multiple of the lines per tile");
  constexpr int TPT = (NTASK + NT - 1) / NT;
  // tasks whose loads are issued together (default: all of the thread's tasks; a smaller batch trades loads in flight
  // for registers, which lets the fused-multiply passes keep 3 CTAs per SM)
  constexpr int TB = (BATCH > 0 && BATCH < TPT) ? BATCH : TPT;
  static_assert(OUT != OUT_RESORT || TB == TPT, "a re-sorting stage reads everything before it writes");
  static_assert(PADBLK == 0 || (OUT == OUT_INPLACE && (STAGE == 0 ? MQ == PADBLK : PADBLK % M == 0)),
                "padded layout: in-place stages of a plan whose first stage has stride PADBLK");
  // padded position of element t of the butterfly at (b, o): stage 0 spans the pads (one per t), later stages sit
  // inside one PADBLK block (b * M / PADBLK pads before it)
  constexpr int TPAD = (PADBLK > 0 && STAGE == 0) ? 1 : 0;
  constexpr bool EVEN = (NTASK % NT == 0);
  constexpr int JSTEP = NT / LINES;
  static_assert(TW != TW_SPLIT || (STAGE == 0 && !LAST), "split twiddles: first stage of a multi-stage plan");
  static_assert(SPLIT_ROW == 0 || NT / LINES <= SPLIT_ROW, "compact twiddle table: one row entry per butterfly slot j0");
  static_assert(TW != TW_CONST || (!LAST && EVEN && TB == TPT && TPT == MQ && NB == JSTEP * MQ),
                "constant twiddles: one block of MQ butterflies per thread");
  // a thread always works on the same line: task = threadIdx.x + i*NT  =>  line = threadIdx.x % LINES.  JFAST swaps the
  // roles (neighbouring lanes = neighbouring butterflies of ONE line): for [line][point] tiles in the padded layout,
  // whose strides are odd, when the stage's results leave for global memory straight from the registers.
  const int line = JFAST ? threadIdx.x / JSTEP : threadIdx.x % LINES;
  const int j0 = JFAST ? threadIdx.x % JSTEP : threadIdx.x / LINES;
  // butterfly index of the thread's task i
  auto task_j = [&](int i) { return TW == TW_CONST ? j0 * MQ + i : j0 + i * JSTEP; };
  auto task_base = [&](int b, int o) {
    return b * M + o + ((PADBLK > 0 && STAGE > 0) ? b / (PADBLK > 0 ? PADBLK / M : 1) : 0);
  };
#pragma unroll
  for (int i0 = 0; i0 < TPT; i0 += TB) {
    // phase 1: the loads of this batch (keeps TB*R independent loads in flight)
    float2 v[TB][R];
#pragma unroll
    for (int ii = 0; ii < TB; ++ii) {
      const int j = task_j(i0 + ii);
      if (i0 + ii < TPT && (EVEN || j < NB)) {
        const int b = j / MQ, o = j - b * MQ;
        const int base = task_base(b, o);
#pragma unroll
        for (int t = 0; t < R; ++t) v[ii][t] = ld(line, base + t * (MQ + TPAD), ii, t);
      }
    }
    if (OUT == OUT_RESORT) __syncthreads();
    if constexpr (TW == TW_SPLIT) {
      // all butterflies, then output q of every task from ONE table load, then the stores
#pragma unroll
      for (int ii = 0; ii < TB; ++ii) {
        const int j = task_j(i0 + ii);
        if (i0 + ii < TPT && (EVEN || j < NB)) {
#pragma unroll
          for (int t = 0; t < R; ++t) v[ii][t] = pre(line, j + t * MQ, ii, t, v[ii][t]);
          Butterfly<R, INV>::run(v[ii]);
        }
      }
      constexpr TwConst<P::N, JSTEP, TPT, R> K{};   // W_N^(q i JSTEP)
      const int oc = j0 * twmul;                    // W_N^(q j0) = W_L[q j0 twmul]
#pragma unroll
      for (int q = 1; q < R; ++q) {
        float2 w = SPLIT_ROW > 0 ? __ldg(tw + P::N * twmul + q * SPLIT_ROW + j0) : __ldg(tw + q * oc);
        if (INV) w.y = -w.y;
#pragma unroll
        for (int ii = 0; ii < TB; ++ii) {
          if (i0 + ii < TPT && (EVEN || task_j(i0 + ii) < NB)) {
            const float2 wi = (i0 + ii == 0) ? w : twc<INV>(w, K.c[i0 + ii][q], K.s[i0 + ii][q]);
            v[ii][q] = cmul(v[ii][q], wi);
          }
        }
      }
#pragma unroll
      for (int ii = 0; ii < TB; ++ii) {
        const int j = task_j(i0 + ii);
        if (i0 + ii < TPT && (EVEN || j < NB)) {
#pragma unroll
          for (int q = 0; q < R; ++q) st(line, j + q * (MQ + TPAD), v[ii][q]);
        }
      }
    } else {
#pragma unroll
      for (int ii = 0; ii < TB; ++ii) {
        const int j = task_j(i0 + ii);
        if (i0 + ii < TPT && (EVEN || j < NB)) {
          const int b = j / MQ, o = j - b * MQ;
#pragma unroll
          for (int t = 0; t < R; ++t) v[ii][t] = pre(line, b * M + o + t * MQ, ii, t, v[ii][t]);
          Butterfly<R, INV>::run(v[ii]);
          if (!LAST) {
            if constexpr (TW == TW_CONST) {
              constexpr TwConst<M, 1, MQ, R> K{};   // W_sub^(q o), o = i0 + ii
              if (i0 + ii > 0) {
#pragma unroll
                for (int q = 1; q < R; ++q) v[ii][q] = twc<INV>(v[ii][q], K.c[i0 + ii][q], K.s[i0 + ii][q]);
              }
            } else {
              const int oc = o * ((P::N / M) * twmul);   // W_sub^(o q) = W_L[q * oc]
#pragma unroll
              for (int q = 1; q < R; ++q) {
                float2 w = __ldg(tw + q * oc);
                if (INV) w.y = -w.y;
                v[ii][q] = cmul(v[ii][q], w);
              }
            }
          }
          if (OUT != OUT_INPLACE) {
            const int nb = P::nat(b * M);
#pragma unroll
            for (int q = 0; q < R; ++q) st(line, nb + q * (P::N / R), v[ii][q]);
          } else {
            const int base = task_base(b, o);
#pragma unroll
            for (int q = 0; q < R; ++q) st(line, base + q * (MQ + TPAD), v[ii][q]);
          }
        }
      }
    }
  }
}

// Stages [S0, S1) in place in shared memory; element (line,pos) at sm[line*LS + pos*PS].
// A __syncthreads() follows every stage.
template <class P, int S0, int S1, bool INV, int LINES, int LS, int PS, int NT>
__device__ __forceinline__ void dif_stages_smem(float2* sm, const float2* __restrict__ tw, int twmul) {
  if constexpr (S0 < S1) {
    auto ld = [&](int line, int pos, int, int) { return sm[line * LS + pos * PS]; };
    auto st = [&](int line, int pos, float2 val) { sm[line * LS + pos * PS] = val; };
    dif_stage<P, S0, INV, LINES, NT, OUT_INPLACE>(ld, st, tw, twmul);
    __syncthreads();
    dif_stages_smem<P, S0 + 1, S1, INV, LINES, LS, PS, NT>(sm, tw, twmul);
  }
}

// Last stage with the outputs re-sorted to natural order inside the same buffer.
template <class P, bool INV, int LINES, int LS, int PS, int NT>
__device__ __forceinline__ void dif_last_resort_smem(float2* sm, const float2* __restrict__ tw, int twmul) {
  auto ld = [&](int line, int pos, int, int) { return sm[line * LS + pos * PS]; };
  auto st = [&](int line, int k, float2 val) { sm[line * LS + k * PS] = val; };
  dif_stage<P, P::S - 1, INV, LINES, NT, OUT_RESORT>(ld, st, tw, twmul);
  __syncthreads();
}

}  // namespace smk
