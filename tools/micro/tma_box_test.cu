// Minimal 3-D TMA box load (cp.async.bulk.tensor.3d + mbarrier) from a [nx][ny][nz] float tensor: the building block of
// the staged skewer gather, checked against the host.   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o tma_box_test tools/micro/tma_box_test.cu
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

struct alignas(64) Params {
  CUtensorMap map[2];
  float* out;
  int xw, yw, zl, x0, y0, z0;
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void box_kernel(const __grid_constant__ Params t, int variant) {
  extern __shared__ __align__(128) unsigned char sm[];
  float* box = reinterpret_cast<float*>(sm);
  const int n = t.xw * t.yw * t.zl;
  const unsigned bar = smem_u32(sm + ((n * 4 + 127) / 128) * 128 * 2);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * n * 4) : "memory");
    for (int f = 0; f < 2; ++f)
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                   ::"r"(smem_u32(box + f * ((n + 31) / 32 * 32))), "l"(&t.map[f]), "r"(t.z0), "r"(t.y0), "r"(t.x0), "r"(bar) : "memory");
  }
  unsigned done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(0) : "memory");
  for (int f = 0; f < 2; ++f)
    for (int i = threadIdx.x; i < n; i += blockDim.x) t.out[f * n + i] = box[f * ((n + 31) / 32 * 32) + i];
}

// variant B: tensor maps in global memory (device array), variant C: one map as a top-level __grid_constant__ parameter
__global__ void box_kernel_ptr(const CUtensorMap* maps, float* out, int n, int x0, int y0, int z0) {
  extern __shared__ __align__(128) unsigned char sm[];
  float* box = reinterpret_cast<float*>(sm);
  const unsigned bar = smem_u32(sm + ((n * 4 + 127) / 128) * 128);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(box)), "l"(maps), "r"(z0), "r"(y0), "r"(x0), "r"(bar) : "memory");
  }
  unsigned done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(0) : "memory");
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = box[i];
}
__global__ void box_kernel_direct(const __grid_constant__ CUtensorMap map, float* out, int n, int x0, int y0, int z0) {
  extern __shared__ __align__(128) unsigned char sm[];
  float* box = reinterpret_cast<float*>(sm);
  const unsigned bar = smem_u32(sm + ((n * 4 + 127) / 128) * 128);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n * 4) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(box)), "l"(&map), "r"(z0), "r"(y0), "r"(x0), "r"(bar) : "memory");
  }
  unsigned done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(0) : "memory");
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = box[i];
}

// variant 3: plain 1-D bulk copy (no tensor map)
__global__ void bulk_kernel(const float* src, float* out, int n) {
  extern __shared__ __align__(128) unsigned char sm[];
  float* box = reinterpret_cast<float*>(sm);
  const unsigned bar = smem_u32(sm + ((n * 4 + 127) / 128) * 128);
  if (threadIdx.x == 0) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(n * 4) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(box)), "l"(src), "r"(n * 4), "r"(bar) : "memory");
  }
  unsigned done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(0) : "memory");
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = box[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 0;
  const int zlv = argc > 2 ? atoi(argv[2]) : 20;
  const int promo = argc > 3 ? atoi(argv[3]) : 1;
  const int nx = 40, ny = 48, nz = 96, zl = zlv;
  const int xw = argc > 4 ? atoi(argv[4]) : 10, yw = argc > 5 ? atoi(argv[5]) : 9;
  const int x0 = argc > 6 ? atoi(argv[6]) : 7, y0 = argc > 7 ? atoi(argv[7]) : -2, z0 = argc > 8 ? atoi(argv[8]) : 85;   // partly out of range
  const size_t ncell = (size_t)nx * ny * nz;
  float* h = (float*)malloc(2 * ncell * sizeof(float));
  for (size_t i = 0; i < 2 * ncell; ++i) h[i] = (float)(i % 100003) * 0.25f + 1.f;
  float *d, *out;
  cudaMalloc(&d, 2 * ncell * sizeof(float));
  cudaMemcpy(d, h, 2 * ncell * sizeof(float), cudaMemcpyHostToDevice);
  const int n = xw * yw * zl;
  cudaMalloc(&out, 2 * n * sizeof(float));
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
  printf("entry point: %s, query %d, ptr %p\n", cudaGetErrorString(e), (int)q, fp);
  Params t{};
  for (int f = 0; f < 2; ++f) {
    cuuint64_t dims[3] = {(cuuint64_t)nz, (cuuint64_t)ny, (cuuint64_t)nx};
    cuuint64_t strides[2] = {(cuuint64_t)nz * 4, (cuuint64_t)nz * ny * 4};
    cuuint32_t boxd[3] = {(cuuint32_t)zl, (cuuint32_t)yw, (cuuint32_t)xw};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult rc = ((EncodeFn)fp)(&t.map[f], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d + f * ncell, dims, strides, boxd, es,
                                 CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                                 CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode field %d: %d\n", f, (int)rc);
  }
  t.out = out; t.xw = xw; t.yw = yw; t.zl = zl; t.x0 = x0; t.y0 = y0; t.z0 = z0;
  const size_t smem = ((n * 4 + 127) / 128) * 128 * 2 + 64;
  cudaFuncSetAttribute(box_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  printf("variant %d zl %d promo %d\n", variant, zl, promo);
  if (variant == 0) {
    box_kernel<<<1, 128, smem>>>(t, 0);
  } else if (variant == 1) {
    CUtensorMap* dm;
    cudaMalloc(&dm, 2 * sizeof(CUtensorMap));
    cudaMemcpy(dm, t.map, 2 * sizeof(CUtensorMap), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(box_kernel_ptr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int f = 0; f < 2; ++f) box_kernel_ptr<<<1, 128, smem>>>(dm + f, out + f * n, n, x0, y0, z0);
  } else if (variant == 3) {
    cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    bulk_kernel<<<1, 128, smem>>>(d, out, n);
    e = cudaDeviceSynchronize();
    printf("bulk kernel: %s\n", cudaGetErrorString(e));
    float* hb = (float*)malloc(n * sizeof(float));
    cudaMemcpy(hb, out, n * sizeof(float), cudaMemcpyDeviceToHost);
    int badb = 0;
    for (int i = 0; i < n; ++i) badb += hb[i] != h[i];
    printf("bulk check: %d mismatches of %d\n", badb, n);
    return badb != 0;
  } else {
    cudaFuncSetAttribute(box_kernel_direct, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int f = 0; f < 2; ++f) box_kernel_direct<<<1, 128, smem>>>(t.map[f], out + f * n, n, x0, y0, z0);
  }
  e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  // ---- variants, each in its own try (a sticky error ends the process: run the binary with an argument per variant)
  float* ho = (float*)malloc(2 * n * sizeof(float));
  cudaMemcpy(ho, out, 2 * n * sizeof(float), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int f = 0; f < 2; ++f)
    for (int a = 0; a < xw; ++a)
      for (int b = 0; b < yw; ++b)
        for (int c = 0; c < zl; ++c) {
          const int x = x0 + a, y = y0 + b, z = z0 + c;
          const bool in = x >= 0 && x < nx && y >= 0 && y < ny && z >= 0 && z < nz;
          const float want = in ? h[f * ncell + ((size_t)x * ny + y) * nz + z] : 0.f;
          if (ho[f * n + (a * yw + b) * zl + c] != want) ++bad;
        }
  printf("box check: %d mismatches of %d\n", bad, 2 * n);
  return bad != 0;
}
