// Philox4x32-10 counter-based generator (Salmon et al. 2011) + Box-Muller.
// Replaces the single-threaded MT19937 np.random.normal loop of DrawGRF_boxk
// (bin/make_boxes.py:46-48): the stream is a pure function of (seed, counter), so the
// noise of a cell does not depend on how the box is split across GPUs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace smk {

__host__ __device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
    uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += W0; k1 += W1;
  }
}

// four N(0,1) variates for 64-bit counter `ctr` under 64-bit key `seed`
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t ctr) {
  uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
  philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
  const float S = 5.9604644775390625e-8f;                 // 2^-24
  float u0 = ((c[0] >> 8) + 1u) * S, u1 = (c[1] >> 8) * S;   // u0 in (0,1], u1 in [0,1)
  float u2 = ((c[2] >> 8) + 1u) * S, u3 = (c[3] >> 8) * S;
  // fast-math Box-Muller: __logf / __sincosf have absolute errors of ~2^-21, far below anything the second
  // moments of a Gaussian field can resolve; the full-precision libm calls made this kernel ALU-bound
  float r0 = sqrtf(-2.0f * __logf(u0)), r1 = sqrtf(-2.0f * __logf(u2));
  float s0, c0, s1, c1;
  __sincosf(6.283185307179586f * u1, &s0, &c0);
  __sincosf(6.283185307179586f * u3, &s1, &c1);
  return make_float4(r0 * c0, r0 * s0, r1 * c1, r1 * s1);
}

}  // namespace smk
