// Microbenchmark behind the index-table design (profiles/r02_notes.md): random 32-byte gathers from a table of
// a given size, 4 independent probes per thread per iteration (like lookup_kernel), optionally followed by a
// dependent 4-byte gather from a second array.  Prints G probes/s.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_bench gather_bench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ld32(const uint4* p, uint4& a, uint4& c) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w)
               : "l"(p));
}
__device__ __forceinline__ uint32_t mix(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}

template <int DEP, int WIDTH>  // WIDTH: bytes per probe (32 or 4)
__global__ void __launch_bounds__(256) gather(const uint4* __restrict__ tab, uint32_t nb, const uint32_t* __restrict__ second,
                                              uint32_t n2, uint32_t n, uint32_t* __restrict__ out) {
  for (uint32_t i0 = (blockIdx.x * 256 + threadIdx.x) * 4; i0 < n; i0 += gridDim.x * 256 * 4) {
    uint32_t acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const uint32_t b = __umulhi(mix(i0 + u), nb);
      if (WIDTH == 32) {
        uint4 a, c;
        ld32(tab + 2 * (size_t)b, a, c);
        acc[u] = a.x ^ a.y ^ a.z ^ a.w ^ c.x ^ c.y ^ c.z ^ c.w;
      } else {
        acc[u] = __ldg(reinterpret_cast<const uint32_t*>(tab) + (size_t)b * 8 + (i0 & 7));
      }
    }
    if (DEP) {
#pragma unroll
      for (int u = 0; u < 4; ++u) acc[u] = __ldg(second + __umulhi(mix(acc[u] + i0 + u), n2));
    }
    *reinterpret_cast<uint4*>(out + i0) = make_uint4(acc[0], acc[1], acc[2], acc[3]);
  }
}

int main() {
  const uint32_t n = 48u << 20;  // probes per launch
  uint32_t* out;
  cudaMalloc(&out, (size_t)n * 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int sizes_mb[] = {8, 16, 27, 32, 48, 64, 80, 96, 128, 192, 256, 512};
  printf("%8s %14s %14s %14s %14s\n", "MB", "32B G/s", "32B+dep4B G/s", "4B G/s", "4B+dep4B G/s");
  for (int mb : sizes_mb) {
    const uint32_t nb = (uint32_t)((size_t)mb << 20) / 32;
    uint4* tab;
    uint32_t* second;
    const uint32_t n2 = 6u << 20;  // 24 MB of 4-byte entries
    cudaMalloc(&tab, (size_t)nb * 32);
    cudaMalloc(&second, (size_t)n2 * 4);
    cudaMemset(tab, 1, (size_t)nb * 32);
    cudaMemset(second, 2, (size_t)n2 * 4);
    float ms[4];
    for (int v = 0; v < 4; ++v) {
      for (int rep = 0; rep < 3; ++rep) {
        if (rep == 1) cudaEventRecord(e0);
        switch (v) {
          case 0: gather<0, 32><<<sms * 8, 256>>>(tab, nb, second, n2, n, out); break;
          case 1: gather<1, 32><<<sms * 8, 256>>>(tab, nb, second, n2, n, out); break;
          case 2: gather<0, 4><<<sms * 8, 256>>>(tab, nb, second, n2, n, out); break;
          default: gather<1, 4><<<sms * 8, 256>>>(tab, nb, second, n2, n, out); break;
        }
      }
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      cudaEventElapsedTime(&ms[v], e0, e1);
      ms[v] /= 2;
    }
    printf("%8d %14.1f %14.1f %14.1f %14.1f\n", mb, n / ms[0] * 1e-6, n / ms[1] * 1e-6, n / ms[2] * 1e-6, n / ms[3] * 1e-6);
    cudaFree(tab);
    cudaFree(second);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
