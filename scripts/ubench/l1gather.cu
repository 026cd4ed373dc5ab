// L1 gather-throughput microbenchmark for sm_100a: cycles per warp-wide LDG for the
// access patterns of the DRR inner loop (L1-resident footprint, 1 CTA of 1024 threads per SM).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l1gather l1gather.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

constexpr int ITERS = 2048;
constexpr int UNROLL = 8;

__device__ __forceinline__ float total(float v) { return v; }
__device__ __forceinline__ float total(float2 v) { return v.x + v.y; }
__device__ __forceinline__ float total(float4 v) { return (v.x + v.y) + (v.z + v.w); }

template <typename T>
__global__ void __launch_bounds__(1024) gather(const T* __restrict__ base, const uint32_t* __restrict__ offs, int nsets,
                                               float* out, long long* clk, uint32_t lane_mask)
{
  // offs: [nsets][32] element offsets for the lanes of a warp; every thread keeps UNROLL of them in
  // registers and re-issues the same patterns shifted by a multiple of 128 elements (alignment-preserving)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc = 0.f;
  uint32_t o[UNROLL];
#pragma unroll
  for (int u = 0; u < UNROLL; ++u)
    o[u] = offs[((u + warp * UNROLL) % nsets) * 32 + lane];
  const long long t0 = clock64();
  for (int i = 0; i < ITERS; i += UNROLL)
  {
    const uint32_t shift = (uint32_t)((i >> 3) & 7) * 128u;
    T v[UNROLL];
    const bool on = (lane_mask >> lane) & 1u;  // predicated-off lanes issue no access
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
    {
      v[u] = T();
      if (on)
        v[u] = __ldg(base + ((o[u] + shift) & 2047u));
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      acc += total(v[u]);
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0)
    clk[blockIdx.x] = t1 - t0;
}

template <typename T>
void run(const char* name, const std::vector<uint32_t>& h_offs, int nsets, size_t n_elems, uint32_t lane_mask = 0xffffffffu)
{
  T* base;
  uint32_t* offs;
  float* out;
  long long* clk;
  cudaMalloc(&base, n_elems * sizeof(T));
  cudaMemset(base, 0, n_elems * sizeof(T));
  cudaMalloc(&offs, h_offs.size() * 4);
  cudaMemcpy(offs, h_offs.data(), h_offs.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&out, 148 * 1024 * 4);
  cudaMalloc(&clk, 148 * 8);
  for (int rep = 0; rep < 2; ++rep)
    gather<T><<<148, 1024>>>(base, offs, nsets, out, clk, lane_mask);
  cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; ++i) avg += h[i];
  avg /= 148;
    printf("%-44s %9.0f clk -> %.2f clk per warp-wide LDG per SM\n", name, avg, avg / (32.0 * ITERS));
  cudaFree(base); cudaFree(offs); cudaFree(out); cudaFree(clk);
}

static uint32_t rng_state = 12345;
static uint32_t rnd() { rng_state = rng_state * 1664525u + 1013904223u; return rng_state >> 8; }

int main()
{
  const int nsets = 64;
  const size_t n_elems = 2048;  // 32 KB of float4: L1 resident
  std::vector<uint32_t> o(nsets * 32);
  // P0: 32 consecutive elements, aligned
  for (int s = 0; s < nsets; ++s) for (int l = 0; l < 32; ++l) o[s * 32 + l] = (s * 32 + l) % n_elems;
  run<float4>("LDG.128 coalesced (512B contiguous)", o, nsets, n_elems);
  run<float2>("LDG.64  coalesced (256B contiguous)", o, nsets, n_elems);
  run<float>("LDG.32  coalesced (128B contiguous)", o, nsets, n_elems);
  // P1: DRR-like: 4 rows (quarters) x 8 lanes, lanes 0.49 element apart, rows at independent random bases
  for (int s = 0; s < nsets; ++s)
    for (int q = 0; q < 4; ++q)
    {
      const uint32_t b = rnd() % (n_elems - 16);
      const float frac = (rnd() % 1000) / 1000.0f;
      for (int l = 0; l < 8; ++l) o[s * 32 + q * 8 + l] = b + (uint32_t)(frac + 0.49f * l);
    }
  run<float4>("LDG.128 DRR-like (8 lanes ~4.4 recs, 4 rows)", o, nsets, n_elems);
  // P2: same but each quarter aligned so that it never crosses a 128B line
  for (int s = 0; s < nsets; ++s)
    for (int q = 0; q < 4; ++q)
    {
      const uint32_t b = (rnd() % (n_elems / 8 - 1)) * 8;
      const float frac = (rnd() % 1000) / 1000.0f;
      for (int l = 0; l < 8; ++l) o[s * 32 + q * 8 + l] = b + (uint32_t)(frac + 0.49f * l);
    }
  run<float4>("LDG.128 DRR-like, no line crossing", o, nsets, n_elems);
  // predication: does a quarter-warp without active lanes cost a data-stage pass?
  run<float4>("  same, quarters 0-1 active (16 lanes)", o, nsets, n_elems, 0x0000ffffu);
  run<float4>("  same, quarter 0 active (8 lanes)", o, nsets, n_elems, 0x000000ffu);
  run<float4>("  same, even lanes active (16 lanes)", o, nsets, n_elems, 0x55555555u);
  run<float4>("  same, 1 lane per quarter active", o, nsets, n_elems, 0x01010101u);
  // P3: all 32 lanes the same record
  for (int s = 0; s < nsets; ++s) { const uint32_t b = rnd() % n_elems; for (int l = 0; l < 32; ++l) o[s * 32 + l] = b; }
  run<float4>("LDG.128 broadcast (1 record)", o, nsets, n_elems);
  // P4: each quarter broadcast of one record (4 records per request)
  for (int s = 0; s < nsets; ++s) for (int q = 0; q < 4; ++q) { const uint32_t b = rnd() % n_elems; for (int l = 0; l < 8; ++l) o[s * 32 + q * 8 + l] = b; }
  run<float4>("LDG.128 one record per quarter", o, nsets, n_elems);
  // P6: DRR-like, but a fraction of the quarters is split between two row/plane segments (lanes 0..k-1 on one run of
  // records, lanes k..7 on another, far away): what a quarter-warp costs when its rays straddle a b-row or c-plane boundary
  for (int pct : {25, 50, 100})
  {
    for (int s = 0; s < nsets; ++s)
      for (int q = 0; q < 4; ++q)
      {
        const uint32_t b0 = rnd() % (n_elems - 16), b1 = rnd() % (n_elems - 16);
        const float frac = (rnd() % 1000) / 1000.0f;
        const bool split = (int)(rnd() % 100) < pct;
        const int k = 1 + (int)(rnd() % 7);
        for (int l = 0; l < 8; ++l) o[s * 32 + q * 8 + l] = ((split && l >= k) ? b1 : b0) + (uint32_t)(frac + 0.49f * l);
      }
    char name[96];
    snprintf(name, sizeof(name), "LDG.128 DRR-like, %d %% of the quarters on 2 segments", pct);
    run<float4>(name, o, nsets, n_elems);
  }
  // P7: every quarter on 4 segments (2 rows x 2 planes)
  for (int s = 0; s < nsets; ++s)
    for (int q = 0; q < 4; ++q)
    {
      uint32_t b[4];
      for (int j = 0; j < 4; ++j) b[j] = rnd() % (n_elems - 16);
      const float frac = (rnd() % 1000) / 1000.0f;
      for (int l = 0; l < 8; ++l) o[s * 32 + q * 8 + l] = b[l >> 1] + (uint32_t)(frac + 0.49f * l);
    }
  run<float4>("LDG.128 DRR-like, every quarter on 4 segments", o, nsets, n_elems);
  // P5: fully scattered
  for (int s = 0; s < nsets; ++s) for (int l = 0; l < 32; ++l) o[s * 32 + l] = rnd() % n_elems;
  run<float4>("LDG.128 scattered (32 lines)", o, nsets, n_elems);
  return 0;
}
