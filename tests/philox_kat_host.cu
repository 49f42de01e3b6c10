// Known-answer test of saclaymocks_b200/csrc/smk_philox.cuh: Philox4x32-10 against the vectors published with
// Random123 (Salmon et al. 2011, examples/kat_vectors, lines "philox4x32 10 ...").  Compiled with nvcc, runs on the
// CPU (tests/test_philox_cpu.py); the device code path is the same __host__ __device__ function.
// Optional arguments: 6 hex words (counter[4], key[2]) -> prints the four output words (used by the test to compare
// with its own Python restatement on random inputs).
#include <stdio.h>
#include <stdlib.h>

#include "../saclaymocks_b200/csrc/smk_philox.cuh"

struct Kat { uint32_t c[4], k[2], out[4]; };
static const Kat KATS[] = {
    {{0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u}, {0x00000000u, 0x00000000u},
     {0x6627e8d5u, 0xe169c58du, 0xbc57ac4cu, 0x9b00dbd8u}},
    {{0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu}, {0xffffffffu, 0xffffffffu},
     {0x408f276du, 0x41c83b0eu, 0xa20bc7c6u, 0x6d5451fdu}},
    {{0x243f6a88u, 0x85a308d3u, 0x13198a2eu, 0x03707344u}, {0xa4093822u, 0x299f31d0u},
     {0xd16cfe09u, 0x94fdccebu, 0x5001e420u, 0x24126ea1u}},
};

int main(int argc, char** argv) {
  if (argc == 7) {
    uint32_t c[4], k[2];
    for (int i = 0; i < 4; ++i) c[i] = (uint32_t)strtoul(argv[1 + i], nullptr, 16);
    for (int i = 0; i < 2; ++i) k[i] = (uint32_t)strtoul(argv[5 + i], nullptr, 16);
    smk::philox4x32_10(c, k[0], k[1]);
    printf("%08x %08x %08x %08x\n", c[0], c[1], c[2], c[3]);
    return 0;
  }
  int bad = 0;
  for (const Kat& t : KATS) {
    uint32_t c[4] = {t.c[0], t.c[1], t.c[2], t.c[3]};
    smk::philox4x32_10(c, t.k[0], t.k[1]);
    const bool ok = c[0] == t.out[0] && c[1] == t.out[1] && c[2] == t.out[2] && c[3] == t.out[3];
    printf("%s  %08x %08x %08x %08x\n", ok ? "ok " : "BAD", c[0], c[1], c[2], c[3]);
    bad += !ok;
  }
  printf("philox4x32-10 known-answer vectors: %d failed\n", bad);
  return bad;
}
