// Micro-benchmarks behind the design of the staged skewer gather: shared-memory wavefronts of overlapping
// (broadcast) window loads at 32 / 64 / 128 bits, FFMA vs FFMA2 (fma.rn.f32x2) issue rates, and both together.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lds_ffma_bench tools/micro/lds_ffma_bench.cu && ./lds_ffma_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm volatile("{ .reg .b64 ra, rb, rc, rd;\n\t"
      "mov.b64 ra, {%2, %3}; mov.b64 rb, {%4, %5}; mov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
      "mov.b64 {%0, %1}, rd; }"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

constexpr int ITER = 2048;

// MODE 0: 8 x LDS.32 at base + c, base = floor(step * lane)         (the window read of the gather)
// MODE 1: 4 x LDS.64 at 2*(floor(step * lane / 2)) + 2c
// MODE 2: 2 x LDS.128 at 4*(floor(step * lane / 4)) + 4c
// MODE 3: 8 x LDS.32, all lanes distinct consecutive addresses (no overlap) -- the reference point
template <int MODE>
__global__ void lds_kernel(float* out, long long* cyc, float step) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i * 1e-3f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int base = (int)(step * lane) + warp * 64;
  if (MODE == 1) base &= ~1;
  if (MODE == 2) base &= ~3;
  if (MODE == 3) base = lane + warp * 64;
  float acc = 0.f;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    const float* p = sm + base + (it & 7) * 32;
    if (MODE == 0 || MODE == 3) {
      float v[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) v[c] = p[MODE == 3 ? c * 32 : c];
#pragma unroll
      for (int c = 0; c < 8; ++c) acc += v[c];
    } else if (MODE == 1) {
      float2 v[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) v[c] = *reinterpret_cast<const float2*>(p + 2 * c);
#pragma unroll
      for (int c = 0; c < 4; ++c) acc += v[c].x + v[c].y;
    } else {
      float4 v[2];
#pragma unroll
      for (int c = 0; c < 2; ++c) v[c] = *reinterpret_cast<const float4*>(p + 4 * c);
#pragma unroll
      for (int c = 0; c < 2; ++c) acc += v[c].x + v[c].y + v[c].z + v[c].w;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// FMA issue: NCH independent chains per thread; MODE 0 = FFMA (3 register operands), 1 = FFMA2
template <int MODE, int NCH>
__global__ void fma_kernel(float* out, long long* cyc, float a, float b) {
  float2 x[NCH];
  for (int i = 0; i < NCH; ++i) x[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
  float2 aa = make_float2(a, a * 1.0001f), bb = make_float2(b, b * 0.9999f);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int i = 0; i < NCH; ++i) {
        if (MODE == 0) { x[i].x = fmaf(x[i].x, aa.x, bb.x); x[i].y = fmaf(x[i].y, aa.y, bb.y); }
        else x[i] = ffma2(x[i], aa, bb);
      }
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < NCH; ++i) s += x[i].x + x[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// gather-like mix: per "row": 8 LDS.32 (overlapping) + 20 FFMA2 (4 pixels x (4 z pairs + 1 row accumulate))
__global__ void mix_kernel(float* out, long long* cyc, float step, int nlds) {
  extern __shared__ float sm[];
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) sm[i] = (float)i * 1e-3f;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int base = (int)(step * lane) + warp * 64;
  float2 w[4][4], acc[4];
  for (int q = 0; q < 4; ++q) { acc[q] = make_float2(0.f, 0.f); for (int j = 0; j < 4; ++j) w[q][j] = make_float2(0.1f * q + j, 0.2f * j + lane); }
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    const float* p = sm + base + (it & 7) * 32;
    float2 r[4];
    if (nlds == 8) {
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = make_float2(p[2 * j], p[2 * j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) r[j] = *reinterpret_cast<const float2*>(p + 2 * j - (base & 1));
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float2 s = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 4; ++j) s = ffma2(w[q][j], r[j], s);
      acc[q] = ffma2(w[q][0], s, acc[q]);
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0].x + acc[1].y + acc[2].x + acc[3].y;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <class F>
static double run(F launch, int blocks) {
  long long* cyc; float* out;
  cudaMalloc(&cyc, blocks * sizeof(long long));
  cudaMalloc(&out, blocks * 1024 * sizeof(float));
  launch(out, cyc);
  launch(out, cyc);
  cudaDeviceSynchronize();
  long long* h = new long long[blocks];
  cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
  double m = 0;
  for (int i = 0; i < blocks; ++i) m += h[i];
  cudaFree(cyc); cudaFree(out); delete[] h;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(e));
  return m / blocks;
}

int main() {
  const int blocks = 148;
  printf("== shared-memory window loads (one CTA per SM; cycles per warp-level load instruction, SM-wide)\n");
  for (int nw : {4, 8, 16}) {
    const int nt = nw * 32;
    double c0 = run([&](float* o, long long* c) { lds_kernel<0><<<blocks, nt, 32768>>>(o, c, 0.37f); }, blocks);
    double c1 = run([&](float* o, long long* c) { lds_kernel<1><<<blocks, nt, 32768>>>(o, c, 0.37f); }, blocks);
    double c2 = run([&](float* o, long long* c) { lds_kernel<2><<<blocks, nt, 32768>>>(o, c, 0.37f); }, blocks);
    double c3 = run([&](float* o, long long* c) { lds_kernel<3><<<blocks, nt, 32768>>>(o, c, 0.37f); }, blocks);
    printf("warps %2d: LDS.32 overlap %.2f | LDS.64 overlap %.2f | LDS.128 overlap %.2f | LDS.32 distinct %.2f  cyc/instr/SM\n", nw,
           c0 / (ITER * 8.0 * nw), c1 / (ITER * 4.0 * nw), c2 / (ITER * 2.0 * nw), c3 / (ITER * 8.0 * nw));
    printf("          per 8-float window per warp: LDS.32 %.2f | LDS.64 %.2f | LDS.128 %.2f cycles (SM-wide)\n",
           c0 / (ITER * 1.0 * nw), c1 / (ITER * 1.0 * nw), c2 / (ITER * 1.0 * nw));
  }
  printf("== FMA issue (cycles per warp instruction per SMSP; FFMA2 carries two FMAs per lane)\n");
  for (int nw : {4, 8, 16}) {
    const int nt = nw * 32;
    double f1 = run([&](float* o, long long* c) { fma_kernel<0, 8><<<blocks, nt>>>(o, c, 1.0001f, 0.5f); }, blocks);
    double f2 = run([&](float* o, long long* c) { fma_kernel<1, 8><<<blocks, nt>>>(o, c, 1.0001f, 0.5f); }, blocks);
    printf("warps %2d: FFMA %.2f | FFMA2 %.2f cyc/instr/SMSP\n", nw, f1 / (ITER * 4.0 * 8 * 2 * (nw / 4.0)),
           f2 / (ITER * 4.0 * 8 * (nw / 4.0)));
  }
  printf("== gather-like mix per row (8 LDS.32 or 4 LDS.64 + 20 FFMA2), cycles per row per warp and SMSP\n");
  for (int nw : {4, 8, 12, 16}) {
    const int nt = nw * 32;
    double m8 = run([&](float* o, long long* c) { mix_kernel<<<blocks, nt, 32768>>>(o, c, 0.37f, 8); }, blocks);
    double m4 = run([&](float* o, long long* c) { mix_kernel<<<blocks, nt, 32768>>>(o, c, 0.37f, 4); }, blocks);
    printf("warps %2d: 8xLDS.32 %.1f | 4xLDS.64 %.1f cycles per row per warp-on-SMSP (FMA-pipe floor 40)\n", nw,
           m8 / (ITER * (nw / 4.0)), m4 / (ITER * (nw / 4.0)));
  }
  return 0;
}
