// Host stand-in for <cuda_runtime.h>: lets g++ compile the kernels' shared-memory FFT stage code (smk_fft.cuh,
// smk_ztile.cuh) so that tests/fft_stage_host.cpp can run it thread by thread on the CPU.  Test infrastructure only.
#pragma once
#include <cmath>
#include <cstdint>
struct float2 { float x, y; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
#define __host__
#define __device__
#define __forceinline__ inline
struct smk_host_dim3 { unsigned x, y, z; };
static smk_host_dim3 threadIdx;                 // set by the emulation loop before each "thread" runs
#include <vector>
static std::vector<const void*>* smk_host_ldg_log = nullptr;   // when set, every __ldg address of the running "thread"
template <class T> static inline T __ldg(const T* p) {
  if (smk_host_ldg_log) smk_host_ldg_log->push_back(p);
  return *p;
}
static inline void __syncthreads() {}           // the harness runs one stage at a time for all threads
