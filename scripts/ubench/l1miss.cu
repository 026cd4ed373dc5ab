// What does an L1 miss cost the LSU data stage, and does prefetch.global.L1 take that cost away?
// Every warp walks its own 4 rows -- 4 KB rings, 512 KB per SM (> L1, cyclic: every new sector misses L1), 76 MB in
// all (< 126 MB L2: every miss hits L2) -- with the DRR kernel's shape:
// 4 quarter-warps on 4 different rows, 8 lanes x 16 B contiguous per quarter, the window sliding by STEP
// bytes per request (STEP = 128: every sector new; 32: one new sector per quarter and request).
// Variants: plain LDG.128; prefetch.global.L1 of the window D requests ahead (all lanes / one lane per quarter);
// prefetch.global.L2 only.  Prints clk per warp-wide request per SM (1024 threads per SM, UNROLL loads in flight).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1miss l1miss.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

constexpr int ITERS = 4096;
constexpr int UNROLL = 4;

template <int MODE>  // 0 plain, 1 prefetch L1 all lanes, 2 prefetch L1 one lane per quarter, 3 prefetch L2 all lanes
__global__ void __launch_bounds__(1024) walk(const float4* __restrict__ base, size_t row_elems, int step_elems, int dist,
                                             float* out, long long* clk)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = lane >> 3, l = lane & 7;
  // this warp's 4 rows
  const size_t row0 = ((size_t)blockIdx.x * 32 + warp) * 4;
  const float4* p = base + (row0 + q) * row_elems;
  const uint32_t ring = (uint32_t)row_elems - 1u;  // row_elems is a power of two
  float acc = 0.f;
  const long long t0 = clock64();
  for (int i = 0; i < ITERS; i += UNROLL)
  {
    float4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
    {
      const float4* a = p + (((uint32_t)(i + u) * step_elems + l) & ring);
      const float4* pf = p + (((uint32_t)(i + u + dist) * step_elems + l) & ring);
      if (MODE == 1 || (MODE == 2 && l == 0))
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
      if (MODE == 3)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
      v[u] = __ldg(a);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      acc += (v[u].x + v[u].y) + (v[u].z + v[u].w);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0)
    clk[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, const float4* base, size_t row_elems, int step_bytes, int dist, float* out, long long* clk)
{
  for (int rep = 0; rep < 2; ++rep)
    walk<MODE><<<148, 1024>>>(base, row_elems, step_bytes / 16, dist, out, clk);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
  printf("%-64s step %3d B: %6.2f clk per request per SM\n", name, step_bytes, avg / (32.0 * ITERS));
}

int main()
{
  // 148 CTAs x 32 warps x 4 rows = 18944 rows of 4 KB
  const size_t row_bytes = 4096;
  const size_t rows = 148 * 32 * 4;
  const size_t total = rows * row_bytes;
  float4* base;
  float* out;
  long long* clk;
  if (cudaMalloc(&base, total) != cudaSuccess)
  {
    printf("alloc failed\n");
    return 1;
  }
  cudaMemset(base, 0, total);
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&clk, 148 * 8);
  const size_t row_elems = row_bytes / 16;
  printf("buffer %.1f MB, %zu KB per SM\n", total / 1e6, total / 148 / 1024);
  for (int step : {128, 64, 32, 16})
  {
    run<0>("LDG.128 plain", base, row_elems, step, 0, out, clk);
    run<1>("LDG.128 + prefetch.global.L1 8 requests ahead, all lanes", base, row_elems, step, 8, out, clk);
    run<2>("LDG.128 + prefetch.global.L1 8 requests ahead, 1 lane per quarter", base, row_elems, step, 8, out, clk);
    run<3>("LDG.128 + prefetch.global.L2 8 requests ahead, all lanes", base, row_elems, step, 8, out, clk);
  }
  cudaFree(base);
  return 0;
}
