// Microbenchmark for the experiment VERDICT r1 asked for (Next #3): is the DRR marching loop faster when a CTA stages
// the footprint of its beam in shared memory (cp.async.bulk + mbarrier ring, one producer warp) and the rays sample from
// there with LDS.128, than when every lane gathers its two 16-byte records per sample from global memory through L1
// (what drr_pax_kernel does)?  Same synthetic "beam" geometry, same per-sample arithmetic (position chain, floor by the
// magic-constant add, record index, 2 x 128-bit loads, difference-form trilinear value, sequential sum) in both modes:
//
//   mode 0  LDG   rec = ic * Sc + ib * Sb + ia;  q0 = ldg(rec), q1 = ldg(rec + Sc)
//   mode 1  SMEM  slabs of P (+1 shared) planes of an nA x nB record box per CTA in a ring of R slots; a producer warp
//                 issues one cp.async.bulk per box row, consumers wait on the slab's mbarrier and read LDS.128
//
// A CTA = 16 x 16 rays (warp = 8 x 4, quarter-warps along the stack's fast axis, as in drr.cu); rays are `pix` voxels
// apart (0.49 for C2), diverge with depth like a cone beam and are tilted by (slope_a, slope_b) voxels per plane; CTAs of
// the same tile for `npose` jittered "poses" are adjacent in launch order (the L2 sharing of the real kernel).
//
//   ./smem_stage mode [pix=0.40] [step=1.25] [P=4] [R=4] [npose=100] [slope=0.06]
//
// Prints ms per launch, Gsamples/s and clk per warp-sample per SM (8.0 is the 128 B/clk/SM register-delivery floor).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

constexpr int NA = 513, NB = 513, NC = 402;   // C2 stack (volume 512 x 512 x 400, AP view: c = y ... sizes only matter for L2 / HBM)
constexpr float kMagic = 12582912.0f;
constexpr int kThreads = 256;
constexpr int kBoxA = 12, kBoxB = 12;          // record box per plane (covers 16 rays * 0.49 + tilt + margins)
constexpr int kPitch = 13;
constexpr float kDiv = 0.0015f;                // footprint grows 60 % from the near to the far side (C2: 0.34 -> 0.65 voxel per pixel)                     // row pitch in records: odd, so that runs on neighbouring rows use other banks

struct Args
{
  const float4* stack;
  float* out;
  int tiles_x, tiles_y, npose, nsamp;
  float pix, step, slope;
  int P, logP, R;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar)
{
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  while (!ok)
  {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// beam of one CTA: ray (i, j) of the 16 x 16 tile at plane c sits at
//   a = a0 + (i * pix) * (1 + c * div) + slope * c,  b = b0 + (j * pix) * (1 + c * div) - slope * c
struct Beam
{
  float a0, b0, div, slope, pix;
};

__device__ __forceinline__ Beam cta_beam(const Args& g, int& pose, int& tile)
{
  pose = blockIdx.x % g.npose;
  tile = blockIdx.x / g.npose;
  const int tx = tile % g.tiles_x, ty = tile / g.tiles_x;
  // jitter of the "pose": a few voxels, deterministic
  const float ja = (float)((pose * 37) % 23) - 11.0f, jb = (float)((pose * 53) % 19) - 9.0f;
  Beam b;
  b.pix = g.pix;
  b.div = kDiv;
  b.slope = g.slope * ((pose & 1) ? 1.0f : -1.0f);
  b.a0 = 30.0f + tx * 16.0f * g.pix + ja;
  b.b0 = 30.0f + ty * 16.0f * g.pix + jb;
  return b;
}

__device__ __forceinline__ float pax_lerp(const float4& q0, const float4& q1, float wa, float wb, float wc)
{
  const float p0 = fmaf(wa, fmaf(wb, q0.w, q0.y), fmaf(wb, q0.z, q0.x));
  const float p1 = fmaf(wa, fmaf(wb, q1.w, q1.y), fmaf(wb, q1.z, q1.x));
  return fmaf(wc, p1 - p0, p0);
}

// ------------------------------------------------------------------------------------------------ mode 0: LDG
__global__ void __launch_bounds__(kThreads, 5) beam_ldg(const Args g)
{
  int pose, tile;
  const Beam bm = cta_beam(g, pose, tile);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = (warp & 1) * 8 + (lane & 7), j = (warp >> 1) * 4 + (lane >> 3);
  // position chain like the reference loop: x += step vector (f32 adds)
  float a = bm.a0 + i * bm.pix, b = bm.b0 + j * bm.pix, c = 1.3f;
  const float sa = (i * bm.pix * bm.div + bm.slope) * g.step, sb = (j * bm.pix * bm.div - bm.slope) * g.step, sc = g.step;
  const uint32_t Sb = NA, Sc = NA * NB;
  const uint32_t K = 0u - 0x4B400000u * (Sc + Sb + 1u);
  float sum = 0.f;
  float4 q0, q1;
  float wa, wb, wc;
  {
    const float ta = __fadd_rd(a, kMagic), tb = __fadd_rd(b, kMagic), tc = __fadd_rd(c, kMagic);
    const uint32_t rec = __float_as_uint(tc) * Sc + __float_as_uint(tb) * Sb + __float_as_uint(ta) + K;
    wa = a - (ta - kMagic), wb = b - (tb - kMagic), wc = c - (tc - kMagic);
    q0 = __ldg(g.stack + rec);
    q1 = __ldg(g.stack + rec + Sc);
    a += sa, b += sb, c += sc;
  }
  for (int s = 1; s < g.nsamp; ++s)
  {
    const float ta = __fadd_rd(a, kMagic), tb = __fadd_rd(b, kMagic), tc = __fadd_rd(c, kMagic);
    const uint32_t rec = __float_as_uint(tc) * Sc + __float_as_uint(tb) * Sb + __float_as_uint(ta) + K;
    const float nwa = a - (ta - kMagic), nwb = b - (tb - kMagic), nwc = c - (tc - kMagic);
    const float4 n0 = __ldg(g.stack + rec);
    const float4 n1 = __ldg(g.stack + rec + Sc);
    a += sa, b += sb, c += sc;
    sum = __fadd_rn(sum, pax_lerp(q0, q1, wa, wb, wc));
    q0 = n0, q1 = n1, wa = nwa, wb = nwb, wc = nwc;
  }
  sum = __fadd_rn(sum, pax_lerp(q0, q1, wa, wb, wc));
  g.out[(size_t)blockIdx.x * kThreads + threadIdx.x] = sum;
}

// ------------------------------------------------------------------------------------------------ mode 1: SMEM
// slab t holds planes [t * P, t * P + P] (P + 1 planes: a sample with floor(c) in the slab finds both its planes there)
__global__ void __launch_bounds__(kThreads + 32, 5) beam_smem(const Args g)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int P = g.P, R = g.R;
  const int slab_recs = (P + 1) * kBoxB * kPitch;
  float4* ring = reinterpret_cast<float4*>(smem_raw);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + (size_t)R * slab_recs * sizeof(float4));
  uint64_t* empty = full + R;
  int* org = reinterpret_cast<int*>(empty + R);   // per slot: a origin, b origin

  int pose, tile;
  const Beam bm = cta_beam(g, pose, tile);
  const int c_last = (int)(1.3f + g.step * (g.nsamp - 1)) + 1;
  const int n_slabs = c_last / P + 1;

  if (threadIdx.x == 0)
  {
    for (int r = 0; r < R; ++r)
    {
      mbar_init(full + r, 1);
      mbar_init(empty + r, kThreads);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == kThreads / 32)
  {
    // ---- producer warp
    for (int t = 0; t < n_slabs; ++t)
    {
      const int slot = t % R;
      if (t >= R)
        mbar_wait(empty + slot, ((t / R) - 1) & 1);
      // box origin: the beam's footprint over planes [t P, t P + P + 1) (rays 0 .. 15 in both directions)
      const float c0 = (float)(t * P), c1 = (float)(t * P + P + 1);
      const float alo = bm.a0 + fminf(bm.slope * c0, bm.slope * c1), blo = bm.b0 + fminf(-bm.slope * c0, -bm.slope * c1);
      const int ao = (int)floorf(alo - 0.5f), bo = (int)floorf(blo - 0.5f);
      // box of this slab: the beam widens with depth (rows of `na` records, `nb` rows per plane)
      const int nab = min(kBoxA, (int)(15.0f * bm.pix * (1.0f + c1 * bm.div) + fabsf(bm.slope) * (float)(P + 1) + 3.0f));
      if (lane == 0)
      {
        org[2 * slot] = ao;
        org[2 * slot + 1] = bo;
        mbar_expect_tx(full + slot, (uint32_t)((P + 1) * nab * nab * sizeof(float4)));
      }
      __syncwarp();
      float4* dst = ring + (size_t)slot * slab_recs;
      for (int row = lane; row < (P + 1) * nab; row += 32)
      {
        const int pl = row / nab, rb = row % nab;
        const size_t src = ((size_t)(t * P + pl) * NB + (size_t)(bo + rb)) * NA + (size_t)ao;
        bulk_g2s(dst + (size_t)(pl * kBoxB + rb) * kPitch, g.stack + src, nab * sizeof(float4), full + slot);
      }
    }
    return;
  }

  // ---- consumers
  const int i = (warp & 1) * 8 + (lane & 7), j = (warp >> 1) * 4 + (lane >> 3);
  float a = bm.a0 + i * bm.pix, b = bm.b0 + j * bm.pix, c = 1.3f;
  const float sa = (i * bm.pix * bm.div + bm.slope) * g.step, sb = (j * bm.pix * bm.div - bm.slope) * g.step, sc = g.step;
  float sum = 0.f;
  int cur = -1;
  uint32_t base = 0;   // shared-memory byte address of record (a origin, b origin, plane 0) of the current slab
  int ao = 0, bo = 0, c_org = 0;
  const uint32_t ring_addr = smem_u32(ring);
  for (int s = 0; s < g.nsamp; ++s)
  {
    const float ta = __fadd_rd(a, kMagic), tb = __fadd_rd(b, kMagic), tc = __fadd_rd(c, kMagic);
    const int ia = (int)(__float_as_uint(ta) - 0x4B400000u), ib = (int)(__float_as_uint(tb) - 0x4B400000u),
              ic = (int)(__float_as_uint(tc) - 0x4B400000u);
    const float wa = a - (ta - kMagic), wb = b - (tb - kMagic), wc = c - (tc - kMagic);
    const int slab = ic >> g.logP;
    if (slab != cur)
    {
      if (cur >= 0)
        mbar_arrive(empty + (cur % R));
      // a thread may skip a slab (step > P never happens here, but be safe): release the ones in between
      for (int k = cur + 1; k < slab && cur >= 0; ++k)
      {
        mbar_wait(full + (k % R), (k / R) & 1);
        mbar_arrive(empty + (k % R));
      }
      cur = slab;
      const int slot = slab % R;
      mbar_wait(full + slot, (slab / R) & 1);
      ao = org[2 * slot];
      bo = org[2 * slot + 1];
      c_org = slab << g.logP;
      base = ring_addr + (uint32_t)(slot * slab_recs) * 16u;
    }
    const uint32_t off = (uint32_t)(((ic - c_org) * kBoxB + (ib - bo)) * kPitch + (ia - ao)) * 16u;
    float4 q0, q1;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(q0.x), "=f"(q0.y), "=f"(q0.z), "=f"(q0.w) : "r"(base + off));
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(q1.x), "=f"(q1.y), "=f"(q1.z), "=f"(q1.w)
                 : "r"(base + off + (uint32_t)(kBoxB * kPitch) * 16u));
    sum = __fadd_rn(sum, pax_lerp(q0, q1, wa, wb, wc));
    a += sa, b += sb, c += sc;
  }
  // release what is left so that the producer can finish
  if (cur >= 0)
    mbar_arrive(empty + (cur % R));
  for (int k = cur + 1; k < n_slabs; ++k)
  {
    mbar_wait(full + (k % R), (k / R) & 1);
    mbar_arrive(empty + (k % R));
  }
  g.out[(size_t)blockIdx.x * kThreads + threadIdx.x] = sum;
}

__global__ void fill(float4* p, size_t n)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const float v = (float)((i * 2654435761u) >> 40) * 1e-6f;
    p[i] = make_float4(v, 0.001f, 0.002f, 0.0005f);
  }
}

int main(int argc, char** argv)
{
  const int mode = argc > 1 ? atoi(argv[1]) : 0;
  Args g;
  g.pix = argc > 2 ? (float)atof(argv[2]) : 0.40f;
  g.step = argc > 3 ? (float)atof(argv[3]) : 1.25f;
  g.P = argc > 4 ? atoi(argv[4]) : 4;
  g.R = argc > 5 ? atoi(argv[5]) : 4;
  g.npose = argc > 6 ? atoi(argv[6]) : 100;
  g.slope = argc > 7 ? (float)atof(argv[7]) : 0.06f;
  g.logP = 0;
  while ((1 << g.logP) < g.P)
    ++g.logP;
  if ((1 << g.logP) != g.P)
  {
    printf("P must be a power of two\n");
    return 1;
  }
  g.tiles_x = g.tiles_y = 30;
  g.nsamp = (int)((NC - 6) / g.step);
  const size_t nrec = (size_t)NA * NB * (NC + 24);   // slack: the last slab reads a few planes past the last sample
  float4* stack;
  CK(cudaMalloc(&stack, nrec * sizeof(float4)));
  fill<<<148 * 8, 256>>>(stack, nrec);
  const int nblocks = g.tiles_x * g.tiles_y * g.npose;
  float* out;
  CK(cudaMalloc(&out, (size_t)nblocks * kThreads * sizeof(float)));
  g.stack = stack;
  g.out = out;
  // the footprint must fit the box: 15 rays * pix * (1 + div * NC) + |slope| * (P + 1) + margins
  const float span = 15.0f * g.pix * (1.0f + kDiv * (NC + g.P)) + fabsf(g.slope) * (g.P + 1) + 2.0f;
  if (mode == 1 && span > (float)kBoxA)
  {
    printf("footprint %.1f records does not fit the %d-record box\n", span, kBoxA);
    return 1;
  }
  const size_t smem = (size_t)g.R * (g.P + 1) * kBoxB * kPitch * sizeof(float4) + 2 * g.R * sizeof(uint64_t) + 2 * g.R * sizeof(int);
  if (mode == 1)
    CK(cudaFuncSetAttribute(beam_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep)
  {
    CK(cudaEventRecord(e0));
    if (mode == 0)
      beam_ldg<<<nblocks, kThreads>>>(g);
    else
      beam_smem<<<nblocks, kThreads + 32, smem>>>(g);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    CK(cudaGetLastError());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (rep > 0 && ms < best)
      best = ms;
  }
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const double samples = (double)nblocks * kThreads * g.nsamp;
  const double warp_samples_per_sm = samples / 32.0 / 148.0;
  printf("{\"mode\": \"%s\", \"pix\": %.2f, \"step\": %.2f, \"P\": %d, \"R\": %d, \"npose\": %d, \"slope\": %.3f, \"smem_bytes\": %zu, "
         "\"ms\": %.3f, \"Gsamples_per_s\": %.1f, \"clk_per_warp_sample\": %.2f}\n",
         mode ? "smem" : "ldg", g.pix, g.step, g.P, g.R, g.npose, g.slope, mode ? smem : (size_t)0, best, samples / best / 1e6,
         best * 1e-3 * clk_khz * 1e3 / warp_samples_per_sm);
  return 0;
}
