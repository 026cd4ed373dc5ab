// Pipe-throughput microbenchmark for sm_100a: warp-instructions / clk / SM for the
// FP32 / integer instructions the DRR inner loop is made of (scalar vs packed f32x2).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu ; run on the GPU box.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

constexpr int ITERS = 4096;
constexpr int ILP = 8;

template <int OP>
__global__ void __launch_bounds__(1024) bench(float* out, float seed, long long* clk)
{
  float v[ILP];
  float2 p[ILP];
  int iv[ILP];
#pragma unroll
  for (int k = 0; k < ILP; ++k)
  {
    v[k] = seed + threadIdx.x + k;
    p[k] = make_float2(v[k], v[k] + 1.f);
    iv[k] = threadIdx.x + k;
  }
  const float c = seed * 0.5f, d = seed * 0.25f;
  const float2 c2 = make_float2(c, c), d2 = make_float2(d, d);
  const long long t0 = clock64();
  for (int i = 0; i < ITERS; ++i)
  {
#pragma unroll
    for (int k = 0; k < ILP; ++k)
    {
      if (OP == 0) v[k] = fmaf(v[k], c, d);                       // FFMA
      if (OP == 1) v[k] = __fadd_rn(v[k], c);                     // FADD
      if (OP == 2) p[k] = __ffma2_rn(p[k], c2, d2);               // FFMA2
      if (OP == 3) p[k] = __fadd2_rn(p[k], c2);                   // FADD2
      if (OP == 4) v[k] = fminf(v[k], c + k);                     // FMNMX
      if (OP == 5) iv[k] = iv[k] * 3 + (int)threadIdx.x;          // IMAD
      if (OP == 6) iv[k] = (iv[k] ^ (int)threadIdx.x) + 7;        // LOP3/IADD3
      if (OP == 7) v[k] = __fadd_rd(v[k], c);                     // FADD.RM
      if (OP == 8) p[k] = __fadd2_rd(p[k], c2);                   // FADD2.RM
      if (OP == 9) { v[k] = fmaf(v[k], c, d); iv[k] = (iv[k] ^ (int)threadIdx.x) + 7; }   // FFMA + ALU mix
      if (OP == 10) { p[k] = __ffma2_rn(p[k], c2, d2); iv[k] = (iv[k] ^ (int)threadIdx.x) + 7; }  // FFMA2 + ALU mix
      if (OP == 11) { v[k] = fmaf(v[k], c, d); p[k] = __ffma2_rn(p[k], c2, d2); }        // FFMA + FFMA2
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < ILP; ++k)
    s += v[k] + p[k].x + p[k].y + (float)iv[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0)
    clk[blockIdx.x] = t1 - t0;
}

template <int OP>
void run(const char* name, int ops_per_iter)
{
  float* out;
  long long* clk;
  int nsm = 148;
  cudaMalloc(&out, sizeof(float) * nsm * 2 * 1024);
  cudaMalloc(&clk, sizeof(long long) * nsm * 2);
  bench<OP><<<nsm * 2, 1024>>>(out, 1.0f, clk);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<OP><<<nsm * 2, 1024>>>(out, 1.0f, clk);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  long long h[296];
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 296; ++i) avg += h[i];
  avg /= 296;
  // wall-clock based: 296 CTAs x 32 warps over 148 SMs (register use decides whether 1 or 2 CTAs are co-resident,
  // so the per-CTA clock64 span is printed for information only)
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double warp_instr_per_sm = 2.0 * 32.0 * ITERS * ILP * ops_per_iter;
  const double clk_wall = ms * 1e-3 * khz * 1e3;
  printf("%-14s  %8.3f ms  (CTA span %9.0f clk)  -> %.2f warp-instr/clk/SM at %d MHz (x32 or x64 lane-ops for packed)\n", name, ms,
         avg, warp_instr_per_sm / clk_wall, khz / 1000);
  cudaFree(out);
  cudaFree(clk);
}

int main()
{
  run<0>("FFMA", 1);
  run<1>("FADD", 1);
  run<2>("FFMA2", 1);
  run<3>("FADD2", 1);
  run<4>("FMNMX", 1);
  run<5>("IMAD", 1);
  run<6>("LOP3+IADD3", 2);
  run<7>("FADD.RM", 1);
  run<8>("FADD2.RM", 1);
  run<9>("FFMA+ALU2", 3);
  run<10>("FFMA2+ALU2", 3);
  run<11>("FFMA+FFMA2", 2);
  return 0;
}
