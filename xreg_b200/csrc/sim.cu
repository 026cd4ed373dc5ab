// 2D similarity metrics on DRR batches for sm_100a: NCC, Grad-NCC, Patch-NCC,
// Patch-Grad-NCC.  Replaces the OpenCL kernels + ViennaCL GEMV reductions of
// lib/regi/sim_metrics_2d/xregImgSimMetric2D{NCC,GradImg,GradNCC,PatchNCC,PatchGradNCC}OCL.cpp
// with the arithmetic of the CPU classes (…CPU.cpp):
//   * Gaussian (OpenCV fixed kernels, separable, reflect-101) + Sobel 3x3 are
//     evaluated in f32 with the oracle's operation order -> gradient images are
//     bit-identical to the CPU restatement;
//   * all moment reductions are accumulated in f64 (raw moments, one pass),
//     reduced with warp shuffles + one block-level step, and combined in a
//     fixed order by a finalize kernel (deterministic, no float atomics);
//   * patch statistics use separable box sums (running sums down the columns,
//     windowed sums across the strip in shared memory): O(P * d) instead of
//     the reference's O(P * d^2) per image.
#include "sim.h"

#include <cooperative_groups.h>
#include <cstring>

namespace xrc
{

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ int reflect101(int i, int n)
{
  if ((unsigned)i < (unsigned)n)  // interior: the common case
    return i;
  if (n == 1)
    return 0;
  while (i < 0 || i >= n)
    i = (i < 0) ? -i : (2 * (n - 1) - i);
  return i;
}

// (row, col) of flat index i in a window `w` columns wide, advanced by the CTA size without div / mod
struct WinIdx
{
  int r, c, dr, dc, w;
  __device__ WinIdx(int i, int stride, int w_) : r(i / w_), c(i % w_), dr(stride / w_), dc(stride % w_), w(w_) {}
  __device__ __forceinline__ void next()
  {
    r += dr;
    c += dc;
    if (c >= w)
    {
      c -= w;
      ++r;
    }
  }
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum NV values over a 256-thread CTA; result valid in thread 0
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* sh /* [8 * NV] */)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; ++k)
  {
    v[k] = warp_sum(v[k]);
    if (lane == 0)
      sh[warp * NV + k] = v[k];
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    const int nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k)
    {
      double s = 0.0;
      for (int w = 0; w < nw; ++w)
        s += sh[w * NV + k];
      v[k] = s;
    }
  }
}

// ----------------------------------------------------------------------------
// Gaussian + Sobel (+ optional NCC moments)   xregImgSimMetric2DGradImgCPU.cpp:32-102
// ----------------------------------------------------------------------------
constexpr int T = kGradTile;
constexpr int kHMax = kMaxGaussWidth / 2;

template <bool MOMENTS>
__global__ void __launch_bounds__(256) grad_kernel(const GradArgs a)
{
  // H: row-filtered window, B: blurred window (image coordinates, clipped to the image)
  __shared__ float Hs[(T + 2 + 2 * kHMax) * (T + 2)];
  __shared__ float Bs[(T + 2) * (T + 2)];
  __shared__ double red[8 * 6];
  __shared__ float cf[kMaxGaussWidth];

  if (threadIdx.x < kMaxGaussWidth)
    cf[threadIdx.x] = a.coeffs[threadIdx.x];
  __syncthreads();

  const int rows = (int)a.rows, cols = (int)a.cols;
  const int img = blockIdx.y;
  const int tile = blockIdx.x;
  const int r0 = (tile / (int)a.tiles_x) * T, c0 = (tile % (int)a.tiles_x) * T;
  const float* __restrict__ src = a.src + (size_t)img * rows * cols;
  const int h = a.gauss_width / 2;

  const int bc0 = max(c0 - 1, 0), bc1 = min(c0 + T, cols - 1);  // window columns (H and B)
  const int br0 = max(r0 - 1, 0), br1 = min(r0 + T, rows - 1);  // B rows
  const int bw = bc1 - bc0 + 1, bh = br1 - br0 + 1;

  if (a.gauss_width > 1)
  {
    const int hr0 = max(r0 - 1 - h, 0), hr1 = min(r0 + T + h, rows - 1);
    const int hh = hr1 - hr0 + 1;
    WinIdx wa(threadIdx.x, blockDim.x, bw);
    for (int i = threadIdx.x; i < hh * bw; i += blockDim.x, wa.next())
    {
      const int r = hr0 + wa.r, c = bc0 + wa.c;
      const float* __restrict__ row = src + (size_t)r * cols;
      float s = fmul(cf[h], __ldg(row + c));
      for (int j = 1; j <= h; ++j)
        s = fadd(s, fmul(cf[h + j], fadd(__ldg(row + reflect101(c - j, cols)), __ldg(row + reflect101(c + j, cols)))));
      Hs[i] = s;
    }
    __syncthreads();
    WinIdx wb(threadIdx.x, blockDim.x, bw);
    for (int i = threadIdx.x; i < bh * bw; i += blockDim.x, wb.next())
    {
      const int r = br0 + wb.r, cc = wb.c;
      float s = fmul(cf[h], Hs[(r - hr0) * bw + cc]);
      for (int j = 1; j <= h; ++j)
        s = fadd(s, fmul(cf[h + j], fadd(Hs[(reflect101(r - j, rows) - hr0) * bw + cc],
                                               Hs[(reflect101(r + j, rows) - hr0) * bw + cc])));
      Bs[i] = s;
    }
  }
  else
  {
    WinIdx wc(threadIdx.x, blockDim.x, bw);
    for (int i = threadIdx.x; i < bh * bw; i += blockDim.x, wc.next())
      Bs[i] = __ldg(src + (size_t)(br0 + wc.r) * cols + bc0 + wc.c);
  }
  __syncthreads();

  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int i = threadIdx.x; i < T * T; i += blockDim.x)
  {
    const int r = r0 + i / T, c = c0 + i % T;
    if (r < rows && c < cols)
    {
      const int rm = (reflect101(r - 1, rows) - br0) * bw, rc = (r - br0) * bw, rp = (reflect101(r + 1, rows) - br0) * bw;
      const int cm = reflect101(c - 1, cols) - bc0, cc = c - bc0, cp = reflect101(c + 1, cols) - bc0;
      const float pmm = Bs[rm + cm], pm0 = Bs[rm + cc], pmp = Bs[rm + cp];
      const float p0m = Bs[rc + cm], p0p = Bs[rc + cp];
      const float ppm = Bs[rp + cm], pp0 = Bs[rp + cc], ppp = Bs[rp + cp];
      // cv::Sobel 3x3 order (verified against cv2: (dm + dp) + 2 d0)
      const float gx = fadd(fadd(fsub(pmp, pmm), fsub(ppp, ppm)), fmul(2.0f, fsub(p0p, p0m)));
      const float sm = fadd(fadd(pmm, pmp), fmul(2.0f, pm0));
      const float sp = fadd(fadd(ppm, ppp), fmul(2.0f, pp0));
      const float gy = fsub(sp, sm);
      const size_t o = (size_t)r * cols + c;
      if (a.gx)
      {
        a.gx[(size_t)img * rows * cols + o] = gx;
        a.gy[(size_t)img * rows * cols + o] = gy;
      }
      if (MOMENTS)
      {
        if (!a.mask || a.mask[o])
        {
          const double dx = gx, dy = gy;
          acc[0] += dx;
          acc[1] += dx * dx;
          acc[2] += dx * (double)__ldg(a.f0x + o);
          acc[3] += dy;
          acc[4] += dy * dy;
          acc[5] += dy * (double)__ldg(a.f0y + o);
        }
      }
    }
  }
  if (MOMENTS)
  {
    block_sum<6>(acc, red);
    if (threadIdx.x == 0)
    {
      double* p = a.partials + ((size_t)img * gridDim.x + tile) * 6;
#pragma unroll
      for (int k = 0; k < 6; ++k)
        p[k] = acc[k];
    }
  }
}

// Fast path for the Gaussian widths the reference uses (0 / 3 / 5 / 7): one WARP streams a strip of
// 32 - 2 (HW + 1) output columns down a band of rows.  Lane = column (halo lanes load the reflect-101
// neighbours, so every lane simply filters the extended signal); the horizontal taps come from warp
// shuffles, the vertical taps from a register ring of row-filtered values, the Sobel operands from a
// 3-row ring of (right - left) and (left + right + 2 centre).  No shared memory, no barriers, ~60
// instructions per pixel instead of ~400 for the tiled kernel above.  Operation order is the oracle's.
template <int HW, bool MOMENTS>
__global__ void __launch_bounds__(256) grad_fast_kernel(const GradArgs a)
{
  constexpr int HALO = HW + 1;
  constexpr int OW = 32 - 2 * HALO;
  const int rows = (int)a.rows, cols = (int)a.cols;
  const int lane = threadIdx.x & 31;
  const int n_cstrips = (cols + OW - 1) / OW;
  const int band_rows = (int)a.band_rows;
  const int n_bands = (rows + band_rows - 1) / band_rows;
  const int unit = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (unit >= n_cstrips * n_bands)
    return;
  const int cstrip = unit % n_cstrips, band = unit / n_cstrips;
  const int img = blockIdx.y;
  const size_t npix = (size_t)rows * cols;
  const float* __restrict__ src = a.src + (size_t)img * npix;
  const int c = cstrip * OW + lane - HALO;
  const int cl = reflect101(c, cols);
  const int r0 = band * band_rows, r1 = min(r0 + band_rows, rows);
  const bool lane_out = (lane >= HALO) && (lane < 32 - HALO) && (c < cols);

  float cf[HW + 1];
#pragma unroll
  for (int j = 0; j <= HW; ++j)
    cf[j] = a.coeffs[HW + j];

  float Hr[2 * HW + 1];
  float D[3], S[3];
#pragma unroll
  for (int k = 0; k < 2 * HW + 1; ++k)
    Hr[k] = 0.f;
  D[0] = D[1] = D[2] = S[0] = S[1] = S[2] = 0.f;
  double acc[6] = {0, 0, 0, 0, 0, 0};

  float x_next = __ldg(src + (size_t)reflect101(r0 - HALO, rows) * cols + cl);
  for (int rr = r0 - HALO; rr < r1 + HALO; ++rr)
  {
    const float x = x_next;
    if (rr + 1 < r1 + HALO)
      x_next = __ldg(src + (size_t)reflect101(rr + 1, rows) * cols + cl);
    // row filter (cv::GaussianBlur, separable, fixed kernel; xregImgSimMetric2DGradImgCPU.cpp:93-96)
    float hval = x;
    if (HW > 0)
    {
      hval = fmul(cf[0], x);
#pragma unroll
      for (int j = 1; j <= HW; ++j)
      {
        const float xl = __shfl_sync(0xffffffffu, x, lane - j), xr = __shfl_sync(0xffffffffu, x, lane + j);
        hval = fadd(hval, fmul(cf[j], fadd(xl, xr)));
      }
    }
#pragma unroll
    for (int k = 0; k < 2 * HW; ++k)
      Hr[k] = Hr[k + 1];
    Hr[2 * HW] = hval;
    if (rr < r0 - HALO + 2 * HW)
      continue;  // ring not full yet
    // column filter -> blurred row rb = rr - HW
    float b = Hr[HW];
    if (HW > 0)
    {
      b = fmul(cf[0], Hr[HW]);
#pragma unroll
      for (int j = 1; j <= HW; ++j)
        b = fadd(b, fmul(cf[j], fadd(Hr[HW - j], Hr[HW + j])));
    }
    const float bl = __shfl_sync(0xffffffffu, b, lane - 1), br = __shfl_sync(0xffffffffu, b, lane + 1);
    D[0] = D[1];
    D[1] = D[2];
    D[2] = fsub(br, bl);
    S[0] = S[1];
    S[1] = S[2];
    S[2] = fadd(fadd(bl, br), fmul(2.0f, b));
    const int ro = rr - HW - 1;  // output row once blurred rows ro - 1, ro, ro + 1 are in the rings
    if (ro < r0 || !lane_out)
      continue;
    // cv::Sobel 3x3 (verified against cv2): gx = (dm + dp) + 2 d0, gy = sp - sm
    const float gx = fadd(fadd(D[0], D[2]), fmul(2.0f, D[1]));
    const float gy = fsub(S[2], S[0]);
    const size_t o = (size_t)ro * cols + c;
    if (a.gx)
    {
      a.gx[(size_t)img * npix + o] = gx;
      a.gy[(size_t)img * npix + o] = gy;
    }
    if (MOMENTS)
    {
      if (!a.mask || a.mask[o])
      {
        const double dx = gx, dy = gy;
        acc[0] += dx;
        acc[1] += dx * dx;
        acc[2] += dx * (double)__ldg(a.f0x + o);
        acc[3] += dy;
        acc[4] += dy * dy;
        acc[5] += dy * (double)__ldg(a.f0y + o);
      }
    }
  }
  if (MOMENTS)
  {
#pragma unroll
    for (int k = 0; k < 6; ++k)
      acc[k] = warp_sum(acc[k]);
    if (lane == 0)
    {
      double* p = a.partials + ((size_t)img * (n_cstrips * n_bands) + unit) * 6;
#pragma unroll
      for (int k = 0; k < 6; ++k)
        p[k] = acc[k];
    }
  }
}

template <int HW>
static void launch_grad_fast(const GradArgs& a, cudaStream_t st)
{
  const uint32_t units = grad_num_parts(a.rows, a.cols, a.gauss_width, a.band_rows);
  const dim3 grid((units + 7) / 8, a.n_imgs);
  if (a.partials)
    grad_fast_kernel<HW, true><<<grid, 256, 0, st>>>(a);
  else
    grad_fast_kernel<HW, false><<<grid, 256, 0, st>>>(a);
}

int launch_grad(const GradArgs& a_in, cudaStream_t st)
{
  GradArgs a = a_in;
  if (!a.n_imgs)
    return XRC_OK;
  if (grad_fast_path(a.gauss_width))
  {
    a.band_rows = grad_band_rows(a.rows, a.cols, a.gauss_width, a.n_imgs);
    switch (a.gauss_width / 2)
    {
      case 0: launch_grad_fast<0>(a, st); break;
      case 1: launch_grad_fast<1>(a, st); break;
      case 2: launch_grad_fast<2>(a, st); break;
      default: launch_grad_fast<3>(a, st); break;
    }
    count_launch();
    XRC_CUDA(cudaGetLastError());
    return XRC_OK;
  }
  a.tiles_x = (a.cols + T - 1) / T;
  a.tiles_y = (a.rows + T - 1) / T;
  const dim3 grid(a.tiles_x * a.tiles_y, a.n_imgs);
  if (a.partials)
    grad_kernel<true><<<grid, 256, 0, st>>>(a);
  else
    grad_kernel<false><<<grid, 256, 0, st>>>(a);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

// ----------------------------------------------------------------------------
// plain NCC moments   xregImgSimMetric2DNCCCPU.cpp:134-209
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) moments_kernel(const MomentArgs a)
{
  __shared__ double red[8 * 3];
  const int img = blockIdx.y;
  const float* __restrict__ m = a.src + (size_t)img * a.npix;
  const uint64_t b = (uint64_t)blockIdx.x * kMomChunk;
  const uint64_t e = min(b + (uint64_t)kMomChunk, a.npix);
  double acc[3] = {0, 0, 0};
  for (uint64_t i = b + threadIdx.x; i < e; i += blockDim.x)
  {
    if (!a.mask || a.mask[i])
    {
      const double v = m[i];
      acc[0] += v;
      acc[1] += v * v;
      acc[2] += v * (double)__ldg(a.f0 + i);
    }
  }
  block_sum<3>(acc, red);
  if (threadIdx.x == 0)
  {
    double* p = a.partials + ((size_t)img * gridDim.x + blockIdx.x) * 3;
    p[0] = acc[0];
    p[1] = acc[1];
    p[2] = acc[2];
  }
}

int launch_moments(const MomentArgs& a, cudaStream_t st)
{
  if (!a.n_imgs)
    return XRC_OK;
  const dim3 grid(a.n_chunks, a.n_imgs);
  moments_kernel<<<grid, 256, 0, st>>>(a);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

// one CTA per image: fixed-order reduction of the partials (thread t sums partials t, t + 256, ...; warp
// shuffles; the 8 warp totals in order), then
// ncc = (Smf - mu_m * Sf0) / (N sigma_f sigma_m), sim = 0.5 (1 - ncc)  (:182,205)
constexpr int kFinalizeThreads = 256;

__global__ void __launch_bounds__(kFinalizeThreads) ncc_finalize_kernel(const NccFinalizeArgs a)
{
  __shared__ double wsum[kFinalizeThreads / 32][6];
  const int img = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nv = 3 * (int)a.n_dirs;
  double s[6] = {0, 0, 0, 0, 0, 0};
  const double* p = a.partials + (size_t)img * a.n_parts * nv;
  for (uint32_t k = threadIdx.x; k < a.n_parts; k += kFinalizeThreads)
  {
    double v[6];
#pragma unroll
    for (int q = 0; q < 6; ++q)
      v[q] = (q < nv) ? p[(size_t)k * nv + q] : 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q)
      s[q] += v[q];
  }
#pragma unroll
  for (int q = 0; q < 6; ++q)
    s[q] = warp_sum(s[q]);
  if (lane == 0)
  {
#pragma unroll
    for (int q = 0; q < 6; ++q)
      wsum[warp][q] = s[q];
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
#pragma unroll
    for (int q = 0; q < 6; ++q)
    {
      double t = wsum[0][q];
      for (int w = 1; w < kFinalizeThreads / 32; ++w)
        t += wsum[w][q];
      s[q] = t;
    }
    float sim_dir[2] = {0.f, 0.f};
    if (a.ssd)
    {
      // sum (f - m)^2 / N from the raw moments (f64: the cancellation costs ~1e-12 relative)
      sim_dir[0] = (float)(((s[1] - 2.0 * s[2]) + a.sf0[0]) / a.n_eff);
    }
    for (uint32_t d = 0; d < (a.ssd ? 0u : a.n_dirs); ++d)
    {
      const double Sm = s[3 * d], Smm = s[3 * d + 1], Smf = s[3 * d + 2];
      const double mu = Sm / a.n_eff;
      double var = (Smm - Sm * mu) / (a.n_eff - 1.0);
      if (!(var > 0.0))
        var = 0.0;
      const float sd = fmaxf(1.0e-6f, (float)sqrt(var));
      const float num = (float)(Smf - mu * a.sf0[d]);
      const float ncc = num / (((float)a.n_eff * a.f_sd[d]) * sd);
      sim_dir[d] = (1.0f - ncc) * 0.5f;
    }
    const float sim = (a.n_dirs == 1) ? sim_dir[0] : (float)(0.5 * ((double)sim_dir[0] + (double)sim_dir[1]));
    a.sims[img] = sim;
    if (a.sims_host)
      a.sims_host[img] = sim;
  }
}

int launch_ncc_finalize(const NccFinalizeArgs& a, cudaStream_t st)
{
  if (!a.n_imgs)
    return XRC_OK;
  ncc_finalize_kernel<<<a.n_imgs, kFinalizeThreads, 0, st>>>(a);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

// ----------------------------------------------------------------------------
// Patch NCC   xregImgSimMetric2DPatchNCCCPU.cpp:74-300,332-441,558-619
//
// One CTA = (column strip, image, direction).  Thread t owns input column
// strip*W + t and keeps running column sums over the last d = 2r+1 rows of
// the NQ quantities needed by the patch statistics; for every completed
// window row the CTA publishes them in shared memory (double buffered, one
// barrier per row) and threads t < W add the d neighbouring columns to obtain
// the box sums of the patch whose top-left corner is (y-d+1, strip*W + t).
// ----------------------------------------------------------------------------
template <int MODE, bool FIXED>
struct PatchQ;

// moving image quantities
template <>
struct PatchQ<0, false>
{
  static constexpr int N = 3;  // m, m^2, m f
};
template <>
struct PatchQ<1, false>
{
  static constexpr int N = 4;  // m, m^2, M m, M m f
};
template <>
struct PatchQ<2, false>
{
  static constexpr int N = 3;  // M m, M m^2, M m f
};
// fixed image quantities
template <>
struct PatchQ<0, true>
{
  static constexpr int N = 2;  // f, f^2
};
template <>
struct PatchQ<1, true>
{
  static constexpr int N = 4;  // f, f^2, M f, M
};
template <>
struct PatchQ<2, true>
{
  static constexpr int N = 3;  // M f, M f^2, M
};

template <int MODE, bool FIXED>
__device__ __forceinline__ void patch_quantities(double mv, double fv, double M, double* q)
{
  if (FIXED)
  {
    if (MODE == 0)
    {
      q[0] = fv;
      q[1] = fv * fv;
    }
    else if (MODE == 1)
    {
      q[0] = fv;
      q[1] = fv * fv;
      q[2] = M * fv;
      q[3] = M;
    }
    else
    {
      q[0] = M * fv;
      q[1] = M * fv * fv;
      q[2] = M;
    }
  }
  else
  {
    if (MODE == 0)
    {
      q[0] = mv;
      q[1] = mv * mv;
      q[2] = mv * fv;
    }
    else if (MODE == 1)
    {
      q[0] = mv;
      q[1] = mv * mv;
      q[2] = M * mv;
      q[3] = M * mv * fv;
    }
    else
    {
      q[0] = M * mv;
      q[1] = M * mv * mv;
      q[2] = M * mv * fv;
    }
  }
}

// mean / clamped std-dev from box sums (detail::ComputePatchMeanStdDev, :558-619)
__device__ __forceinline__ void stats_from_sums(double S, double SS, double cnt, float& mean, float& sd,
                                                double* mean_d = nullptr)
{
  double mu = S, var = 0.0;
  if (cnt > 1.0)
  {
    mu = S / cnt;
    var = (SS - S * mu) / (cnt - 1.0);
    if (!(var > 0.0))
      var = 0.0;
  }
  mean = (float)mu;
  if (mean_d)
    *mean_d = mu;
  sd = fmaxf(1.0e-6f, (float)sqrt(var));
}

// Box sums by running sums in BOTH directions (O(1) per pixel and quantity):
//   * vertical: thread t owns C adjacent input columns of its strip and keeps, per
//     quantity, the running sum V of the last d rows (add the row entering the window,
//     subtract the row leaving it), f64;
//   * horizontal: per output row, an exclusive prefix sum P of V across the strip
//     (C-element local prefix, warp shuffle scan, warp totals through shared memory)
//     is published in shared memory and the d-column window sum is P[j + d] - P[j]:
//     one shared-memory read per output and quantity instead of d.
// One CTA = (row band, column strip, image, direction); a band first accumulates the
// d-1 rows above its first window (loads only), so bands give the grid enough CTAs
// without repeating the horizontal work.  Loads of the next row are issued before the
// scan of the current one.
// ST1: patch stride 1 (the reference apps' setting) known at compile time -- no integer divisions in the output section
template <int MODE, bool FIXED, int C, bool ST1>
__global__ void __launch_bounds__(kPatchThreads, (MODE == 0) ? 4 : 1) patch_kernel(const PatchArgs a)
{
  constexpr int NQ = PatchQ<MODE, FIXED>::N;
  constexpr int NT = kPatchThreads;
  constexpr int NW = NT / 32;
  // exclusive prefix at local column j = C * t + i is stored at Pex[buf][q][i][t]; local columns run to C * NT (incl.)
  __shared__ double Pex[2][NQ][C][NT + 1];
  __shared__ double wtot[2][NQ][NW];
  __shared__ double red[8];

  const int rows = (int)a.rows, cols = (int)a.cols;
  const int r = (int)a.radius, d = 2 * r + 1;
  const int W_out = C * NT - 2 * r;
  const int strip = blockIdx.x % a.n_strips, band = blockIdx.x / a.n_strips;
  const int img = blockIdx.y, dir = blockIdx.z;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int c0 = strip * W_out;            // first input column == first stride-1 centre column of the strip
  const int ncc_all = cols - 2 * r;        // stride-1 centre columns
  const int nrr_all = rows - 2 * r;        // stride-1 centre rows
  const int st = ST1 ? 1 : (int)a.stride;
  const int ncc_s = ST1 ? ncc_all : ((ncc_all - 1) / st + 1);
  const size_t npix = (size_t)rows * cols;
  const int i_begin = band * (int)a.band_rows, i_end = min(i_begin + (int)a.band_rows, nrr_all);

  const float* __restrict__ m = FIXED ? nullptr : (a.mov[dir] + (size_t)img * npix);
  const float* __restrict__ f = a.fix[dir];
  const uint8_t* __restrict__ mask = a.mask;
  const double n_full = (double)d * (double)d;
  const double inv_n = 1.0 / n_full, inv_nm1 = 1.0 / (n_full - 1.0);

  bool col_ok[C];
#pragma unroll
  for (int i = 0; i < C; ++i)
    col_ok[i] = (c0 + C * t + i) < cols;

  struct RowVals
  {
    float mv[C], fv[C], Mv[C];
  };
  auto load_row = [&](int y, RowVals& v) {
#pragma unroll
    for (int i = 0; i < C; ++i)
    {
      v.mv[i] = 0.f;
      v.fv[i] = 0.f;
      v.Mv[i] = 0.f;
      if (col_ok[i] && y < rows)
      {
        const size_t o = (size_t)y * cols + c0 + C * t + i;
        v.fv[i] = __ldg(f + o);
        if (!FIXED)
          v.mv[i] = __ldg(m + o);
        v.Mv[i] = (MODE != 0) ? (mask[o] ? 1.f : 0.f) : 1.f;
      }
    }
  };

  double V[C][NQ];
#pragma unroll
  for (int i = 0; i < C; ++i)
#pragma unroll
    for (int k = 0; k < NQ; ++k)
      V[i][k] = 0.0;
  auto accumulate = [&](const RowVals& v, double sign) {
#pragma unroll
    for (int i = 0; i < C; ++i)
    {
      double q[NQ];
      patch_quantities<MODE, FIXED>((double)v.mv[i], (double)v.fv[i], (double)v.Mv[i], q);
#pragma unroll
      for (int k = 0; k < NQ; ++k)
        V[i][k] += sign * q[k];
    }
  };

  // rows above the first window of the band
  for (int y = i_begin; y < i_begin + d - 1; ++y)
  {
    RowVals v;
    load_row(y, v);
    accumulate(v, 1.0);
  }

  double total = 0.0;
  RowVals v_new, v_old;
  load_row(i_begin + d - 1, v_new);
  load_row(i_begin, v_old);
  for (int i = i_begin; i < i_end; ++i)
  {
    const int buf = i & 1;
    accumulate(v_new, 1.0);  // window rows i .. i + d - 1 complete
    // local inclusive prefix over my C columns, then exclusive scan of the block totals
    double L[C][NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k)
    {
      L[0][k] = V[0][k];
#pragma unroll
      for (int c = 1; c < C; ++c)
        L[c][k] = L[c - 1][k] + V[c][k];
    }
    double incl[NQ];
#pragma unroll
    for (int k = 0; k < NQ; ++k)
    {
      double x = L[C - 1][k];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const double y2 = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o)
          x += y2;
      }
      incl[k] = x;
      if (lane == 31)
        wtot[buf][k][warp] = x;
    }
    // subtract the row that leaves the window and prefetch the next rows while the scan settles
    accumulate(v_old, -1.0);
    RowVals n_new, n_old;
    if (i + 1 < i_end)
    {
      load_row(i + d, n_new);
      load_row(i + 1, n_old);
    }
    __syncthreads();
    double ex[NQ];  // exclusive prefix of my first column
#pragma unroll
    for (int k = 0; k < NQ; ++k)
    {
      // totals of the warps to my left, added in warp order (a fixed-length chain under warp-uniform predicates:
      // the counted loop cost ~60 instructions per quantity)
      double off = 0.0;
#pragma unroll
      for (int w = 0; w < NW - 1; ++w)
      {
        const double wt = wtot[buf][k][w];
        if (w < warp)
          off += wt;
      }
      ex[k] = off + (incl[k] - L[C - 1][k]);
      Pex[buf][k][0][t] = ex[k];
#pragma unroll
      for (int c = 1; c < C; ++c)
        Pex[buf][k][c][t] = ex[k] + L[c - 1][k];
      if (t == NT - 1)
        Pex[buf][k][0][NT] = ex[k] + L[C - 1][k];  // local column C * NT
    }
    __syncthreads();
    if (i % st == 0 || FIXED)
    {
      float vout[C];       // per-patch values of this thread's columns (reference-order combine)
      bool vhave[C];
#pragma unroll
      for (int cc = 0; cc < C; ++cc)
      {
        vhave[cc] = false;
        vout[cc] = 0.0f;
        const int j = C * t + cc;  // local column == patch's left column
        const int c = c0 + j;
        if (j < W_out && c < ncc_all && (FIXED || c % st == 0))
        {
          double S[NQ];
#pragma unroll
          for (int k = 0; k < NQ; ++k)
          {
            const int je = j + d;
            const double lo = (cc == 0) ? ex[k] : (ex[k] + L[cc - 1][k]);
            S[k] = Pex[buf][k][je % C][je / C] - lo;
          }
          const size_t pa = (size_t)i * ncc_all + c;  // index on the stride-1 grid
          if (FIXED)
          {
            float mean, sd;
            double mean_d;
            if (MODE == 0)
            {
              stats_from_sums(S[0], S[1], n_full, mean, sd, &mean_d);
              a.o_mean[dir][pa] = mean_d;
              a.o_den[dir][pa] = sd * (float)n_full;
            }
            else if (MODE == 1)
            {
              stats_from_sums(S[0], S[1], n_full, mean, sd, &mean_d);
              a.o_mean[dir][pa] = mean_d;
              a.o_den[dir][pa] = sd * (float)n_full;
              a.o_smask[dir][pa] = S[2];
              if (dir == 0)
                a.o_nmask[pa] = (float)S[3];
            }
            else
            {
              stats_from_sums(S[0], S[1], S[2], mean, sd, &mean_d);
              a.o_mean[dir][pa] = mean_d;
              a.o_den[dir][pa] = sd * (float)S[2];
              a.o_smask[dir][pa] = S[0];
              if (dir == 0)
                a.o_nmask[pa] = (float)S[2];
            }
          }
          else
          {
            const size_t pk = (size_t)(i / st) * ncc_s + (c / st);  // strided patch index
            const float w = a.weights ? __ldg(a.weights + pk) : 1.0f;
            float val = 0.0f;  // cur_mov_img_patch_ncc_vals_[k] stays 0 for a skipped patch (:247-250)
            if (!a.weight_patch_sims || (fabsf(w) > 1.0e-6f))
            {
              float sd_m;
              double num;
              const double mu_f = __ldg(a.f_mean[dir] + pa);
              const float den_f = __ldg(a.f_den[dir] + pa);
              if (MODE == 0)
              {
                // mean / unbiased variance from raw moments; reciprocals of the constant counts
                const double mu = S[0] * inv_n;
                double var = (S[1] - S[0] * mu) * inv_nm1;
                if (!(var > 0.0))
                  var = 0.0;
                sd_m = fmaxf(1.0e-6f, sqrtf((float)var));
                // sum (m - mu_m)(f - mu_f) = Smf - mu_f Sm   (sum (f - mu_f) = 0)
                num = S[2] - mu_f * S[0];
              }
              else
              {
                const double nM = (double)__ldg(a.n_mask + pa);
                const double SfM = __ldg(a.f_smask[dir] + pa);
                double SmM, SmfM, mu_md;
                float mu_m;
                if (MODE == 1)
                {
                  stats_from_sums(S[0], S[1], n_full, mu_m, sd_m, &mu_md);
                  SmM = S[2];
                  SmfM = S[3];
                }
                else
                {
                  stats_from_sums(S[0], S[1], nM, mu_m, sd_m, &mu_md);
                  SmM = S[0];
                  SmfM = S[2];
                }
                num = SmfM - mu_f * SmM - mu_md * SfM + mu_md * mu_f * nM;
              }
              const float accv = (den_f != 0.0f) ? ((float)num / (sd_m * den_f)) : 0.0f;
              const float sv = 1.0f - accv;
              val = (a.weight_patch_sims ? w : 1.0f) * sv;
              total += (double)val;
            }
            vout[cc] = val;
            vhave[cc] = true;
          }
        }
      }
      if (!FIXED && a.vals)
      {
        // [img][dir][strided patch, row-major]: the reference's patch order (xregImgSimMetric2DPatchCommon.cpp:275-290)
        const size_t row0 = ((size_t)img * a.n_dirs + dir) * a.n_patches + (size_t)(i / st) * ncc_s;
        const int cfirst = c0 + C * t;
        // stride 1: the C columns of a thread are adjacent patches -> one 8-byte store when aligned
        if (C == 2 && st == 1 && vhave[0] && vhave[C - 1] && (((row0 + cfirst) & 1) == 0))
        {
          *reinterpret_cast<float2*>(a.vals + row0 + cfirst) = make_float2(vout[0], vout[C - 1]);
        }
        else
        {
#pragma unroll
          for (int cc = 0; cc < C; ++cc)
            if (vhave[cc])
              a.vals[row0 + (size_t)((cfirst + cc) / st)] = vout[cc];
        }
      }
    }
    v_new = n_new;
    v_old = n_old;
  }
  if (!FIXED)
  {
    double v[1] = {total};
    __syncthreads();
    block_sum<1>(v, red);
    if (t == 0)
      a.partials[((size_t)img * a.n_dirs + dir) * a.n_parts + blockIdx.x] = v[0];
  }
}

template <bool FIXED, int C, bool ST1>
static int launch_patch_cs(const PatchArgs& a, cudaStream_t st)
{
  const uint32_t n_imgs = FIXED ? 1 : a.n_imgs;
  const dim3 grid(a.n_parts, n_imgs, a.n_dirs);
  switch (a.mask_mode)
  {
    case 0: patch_kernel<0, FIXED, C, ST1><<<grid, kPatchThreads, 0, st>>>(a); break;
    case 1: patch_kernel<1, FIXED, C, ST1><<<grid, kPatchThreads, 0, st>>>(a); break;
    case 2: patch_kernel<2, FIXED, C, ST1><<<grid, kPatchThreads, 0, st>>>(a); break;
    default: XRC_FAIL(XRC_ERR_INVALID, "bad patch mask mode");
  }
  return XRC_OK;
}

template <bool FIXED, int C>
static int launch_patch_c(const PatchArgs& a, cudaStream_t st)
{
  if (a.stride == 1)
    XRC_TRY((launch_patch_cs<FIXED, C, true>(a, st)));
  else
    XRC_TRY((launch_patch_cs<FIXED, C, false>(a, st)));
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

template <bool FIXED>
static int launch_patch_impl(const PatchArgs& a_in, cudaStream_t st)
{
  PatchArgs a = a_in;
  if (!(FIXED ? 1u : a.n_imgs))
    return XRC_OK;
  const PatchPlan pl = patch_plan(a.rows, a.cols, a.radius, (FIXED ? 1u : a.n_imgs) * a.n_dirs);
  a.n_strips = pl.n_strips;
  a.n_parts = pl.n_strips * pl.n_bands;
  a.band_rows = pl.band_rows;
  return (pl.cols_per_thread == 2) ? launch_patch_c<FIXED, 2>(a, st) : launch_patch_c<FIXED, 1>(a, st);
}

int launch_patch(const PatchArgs& a, cudaStream_t st) { return launch_patch_impl<false>(a, st); }
int launch_patch_fixed_stats(const PatchArgs& a, cudaStream_t st) { return launch_patch_impl<true>(a, st); }

// image score = sum_k w_k s_k / divisor per direction (:262-287), then
// 0.5 (x + y) for the gradient variant (xregImgSimMetric2DPatchGradNCCCPU.cpp:222)
// one warp per image: lane l sums partials l, l + 32, ... in order, then a shuffle tree (fixed order)
__global__ void __launch_bounds__(32) patch_finalize_kernel(const PatchFinalizeArgs a)
{
  const uint32_t img = blockIdx.x;
  float sd[2] = {0.f, 0.f};
  if (a.seq_sums)
  {
    // reference order: f32 sum / f32 divisor (:268-287), then 0.5 * (x + y)
    if (threadIdx.x == 0)
    {
      for (uint32_t d = 0; d < a.n_dirs; ++d)
      {
        const float s = a.seq_sums[(size_t)img * a.n_dirs + d];
        sd[d] = a.divide ? __fdiv_rn(s, a.divisor_f) : s;
      }
      const float sim = (a.n_dirs == 1) ? sd[0] : (float)(0.5 * (double)__fadd_rn(sd[0], sd[1]));
      a.sims[img] = sim;
      if (a.sims_host)
        a.sims_host[img] = sim;
    }
    return;
  }
  for (uint32_t d = 0; d < a.n_dirs; ++d)
  {
    double s = 0.0;
    const double* p = a.partials + ((size_t)img * a.n_dirs + d) * a.n_parts;
    for (uint32_t k = threadIdx.x; k < a.n_parts; k += 32)
      s += p[k];
    s = warp_sum(s);
    sd[d] = (float)(s / a.divisor);
  }
  if (threadIdx.x == 0)
  {
    const float sim = (a.n_dirs == 1) ? sd[0] : (float)(0.5 * ((double)sd[0] + (double)sd[1]));
    a.sims[img] = sim;
    if (a.sims_host)
      a.sims_host[img] = sim;
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The reference combines the per-patch values with `Scalar patch_sims_sum = 0; for (s : vals) patch_sims_sum += s;`
// (xregImgSimMetric2DPatchNCCCPU.cpp:262-266): a sequential f32 sum whose rounding errors do not average out
// (206 116 addends at C2; with mask-coverage weights they are near-constant and the error is systematic, ~3e-5 of the
// similarity value).  To agree with the reference to ~1e-7 instead of ~1e-5 the same sum is evaluated here, exactly,
// but not serially:
//   while the running sum S stays inside one binade [2^e, 2^(e+1)) its ulp u = 2^(e-23) is constant and
//   RN(S + x) = S + u * rint(x / u) unless x / u lies exactly half-way between two integers (a tie, whose
//   resolution depends on the parity of S / u).  So a chunk of addends can be rounded to the grid independently,
//   added as integers (associative: warp scan), and accepted if (a) no addend is a tie, (b) every prefix
//   S/u + sum q stays in [2^23 + 1, 2^24 - 1] (strictly inside the binade, so that the grid really is u).
//   A chunk that fails is summed serially (one warp, 128 dependent adds) and the chunks after it are judged again
//   against the new S.
// One CTA (8 warps) per sequence; a round is 4096 elements = 32 chunks, one per quarter-warp (16 consecutive
// addends per lane); lane w of warp 0 judges chunk w.  Bitwise equal to the
// literal loop (patch_seqsum_serial_kernel; tests/test_gpu_metrics.py::test_seqsum_emulation_is_bit_exact).
constexpr int kSeqChunk = 128;   // addends per chunk: the unit that is accepted or falls back to the literal loop
constexpr int kSeqChunks = 32;   // chunks in flight per round (lane w of warp 0 judges chunk w)
constexpr int kSeqPerLane = 16;  // consecutive addends per lane: a chunk is a quarter-warp (8 lanes)
constexpr int kSeqWarps = kSeqChunks * kSeqChunk / (32 * kSeqPerLane);  // 8

// lane l of warp w holds elements [i, i + 16), i = round + 512 w + 16 l; the tail is padded with +0 (neutral: the
// running sum is never -0, a sum started at +0 cannot reach it)
__device__ __forceinline__ void seq_load(const float* __restrict__ v, uint64_t n, uint64_t i, float (&x)[kSeqPerLane])
{
  if (i + kSeqPerLane <= n && ((reinterpret_cast<uintptr_t>(v + i) & 15u) == 0))
  {
#pragma unroll
    for (int k = 0; k < kSeqPerLane / 4; ++k)
    {
      const float4 q = __ldg(reinterpret_cast<const float4*>(v + i) + k);
      x[4 * k] = q.x; x[4 * k + 1] = q.y; x[4 * k + 2] = q.z; x[4 * k + 3] = q.w;
    }
  }
  else
  {
#pragma unroll
    for (int j = 0; j < kSeqPerLane; ++j)
      x[j] = (i + j < n) ? __ldg(v + i + j) : 0.0f;
  }
}

// the literal loop over one chunk (quarter-warp `qtr`): every lane of the warp runs the same chain, the addends
// come by shuffle
__device__ __forceinline__ float seq_chunk_serial(float S, const float (&x)[kSeqPerLane], int qtr)
{
#pragma unroll 1
  for (int l = 8 * qtr; l < 8 * qtr + 8; ++l)
  {
#pragma unroll
    for (int j = 0; j < kSeqPerLane; ++j)
      S = __fadd_rn(S, __shfl_sync(0xffffffffu, x[j], l));
  }
  return S;
}

__global__ void __launch_bounds__(kSeqWarps * 32) patch_seqsum_kernel(const SeqSumArgs a)
{
  static_assert(kSeqChunks == 32 && kSeqPerLane * 8 == kSeqChunk, "lane w of warp 0 judges chunk w; a chunk is 8 lanes");
  // per-chunk records, double buffered by iteration parity: one barrier per iteration is enough (a warp can only
  // overwrite set k % 2 in iteration k after passing barrier k - 1, which every warp reaches after its reads of
  // iteration k - 2)
  __shared__ int sh_sum[2][kSeqChunks], sh_min[2][kSeqChunks], sh_max[2][kSeqChunks], sh_bad[2][kSeqChunks];
  __shared__ float sh_S_serial[2];
  const float* __restrict__ v = a.vals + (size_t)blockIdx.x * a.n;
  const uint64_t n = a.n;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int qtr = lane >> 3, chunk = 4 * warp + qtr;  // this lane's chunk within the round
  constexpr uint64_t kRound = (uint64_t)kSeqChunks * kSeqChunk;
  const uint64_t mine = (uint64_t)threadIdx.x * kSeqPerLane;  // offset of this lane's elements inside a round
  float S = 0.0f;
  int it = 0;  // iteration parity
  // rounds are fixed 4096-element windows, the same lanes always own the same part of the window: the loads do not
  // depend on the sum and are issued two rounds ahead into three register buffers used in rotation (the loop is
  // unrolled by hand: a register move from a buffer whose load is in flight would wait for it)
  auto run_round = [&](const uint64_t base, const float (&x)[kSeqPerLane]) {
    int first = 0;  // chunks [0, first) of this round are already in S
    while (first < kSeqChunks && base + (uint64_t)first * kSeqChunk < n)
    {
      it ^= 1;
      const float aS = fabsf(S);
      // fast path needs a normal, finite S with room for the scaling below: 2^-100 <= |S| < 2^100
      const bool fast = (aS >= 7.888609e-31f) && (aS < 1.2676506e30f);
      const int e = (__float_as_int(aS) >> 23) - 127;                     // |S| in [2^e, 2^(e+1))
      int ok = first;
      if (fast)
      {
        if (4 * warp + 3 >= first)
        {
          const float scale = __int_as_float((127 + 23 - e) << 23);       // 1 / u = 2^(23 - e), exact
          const float sgn = (S < 0.0f) ? -1.0f : 1.0f;                    // RN is symmetric: work with |S|, sgn * x
          // 16 independent roundings first (instruction-level parallelism: only two warps share a scheduler), then
          // the integer prefix chain
          int q[kSeqPerLane];
          unsigned badbits = 0;
#pragma unroll
          for (int j = 0; j < kSeqPerLane; ++j)
          {
            const float t = __fmul_rn(sgn * x[j], scale);                 // exact (power of two) or flushed towards 0 when tiny
            // rint and float -> int by the 1.5 * 2^23 trick (exact for |t| < 2^22): FRND / F2I run on the
            // quarter-rate conversion pipe
            const float tm = __fadd_rn(t, 12582912.0f);
            const float r = __fadd_rn(tm, -12582912.0f);
            // |t| <= 2^18 keeps every integer below inside int32 with room to spare (a larger addend cannot stay in
            // the binade for long anyway); NaN fails the comparison too
            const bool b = !(fabsf(t) <= 262144.0f) || (fabsf(__fsub_rn(t, r)) == 0.5f);
            badbits |= b ? (1u << j) : 0u;
            q[j] = __float_as_int(tm) - 0x4B400000;
          }
          const bool bad = badbits != 0u;
          int p = 0, mn = 0x7fffffff, mx = -0x7fffffff - 1;
#pragma unroll
          for (int j = 0; j < kSeqPerLane; ++j)
          {
            p += q[j];
            mn = min(mn, p);
            mx = max(mx, p);
          }
          if (bad)
          {
            // keep the integers below finite; the chunk is rejected anyway
            p = 0;
            mn = 0;
            mx = 0;
          }
          // prefix over the 8 lanes of the chunk
          int incl = p;
#pragma unroll
          for (int o = 1; o < 8; o <<= 1)
          {
            const int y = __shfl_up_sync(0xffffffffu, incl, o, 8);
            if ((lane & 7) >= o)
              incl += y;
          }
          const int pre = incl - p;
          mn += pre;
          mx += pre;
#pragma unroll
          for (int o = 1; o < 8; o <<= 1)
          {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o, 8));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o, 8));
          }
          const int tot = __shfl_sync(0xffffffffu, incl, 7, 8);
          const unsigned any_bad = (__ballot_sync(0xffffffffu, bad) >> (8 * qtr)) & 0xffu;
          if ((lane & 7) == 0)
          {
            sh_sum[it][chunk] = tot;   // |tot| <= 128 * 2^18 = 2^25
            sh_min[it][chunk] = mn;
            sh_max[it][chunk] = mx;
            sh_bad[it][chunk] = any_bad != 0u;
          }
        }
        __syncthreads();
        // every warp judges all chunks (lane w: chunk w): accepted if every chunk of [first, w) is and its own
        // prefixes stay inside the binade.  32 sums of at most 2^25 fit int32.
        const int A = (__float_as_int(aS) & 0x7fffff) | 0x800000;  // |S| / u, in [2^23, 2^24)
        const bool pending = lane >= first;
        const bool exists = base + (uint64_t)lane * kSeqChunk < n;
        const int my_sum = pending ? sh_sum[it][lane] : 0;
        const int my_min = sh_min[it][lane], my_max = sh_max[it][lane], my_bad = sh_bad[it][lane];
        int incl = my_sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
          const int y = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o)
            incl += y;
        }
        // beyond the first rejected chunk the prefixes are meaningless (and may have wrapped); they are not used
        const int A_in = A + incl - my_sum;
        const bool valid = !pending || (exists && !my_bad && ((long long)A_in + my_min >= (1ll << 23) + 1) &&
                                        ((long long)A_in + my_max <= (1ll << 24) - 1));
        const unsigned fails = __ballot_sync(0xffffffffu, !valid);
        ok = fails ? (__ffs(fails) - 1) : 32;  // >= first: the lanes below first vote valid
        const int add = __shfl_sync(0xffffffffu, incl, ok > 0 ? ok - 1 : 0);
        const int A_out = A + (ok > 0 ? add : 0);
        const float u = __int_as_float((127 + e - 23) << 23);
        const float mag = __fmul_rn((float)A_out, u);  // A_out < 2^24: exact
        S = (S < 0.0f) ? -mag : mag;
      }
      if (ok < kSeqChunks && base + (uint64_t)ok * kSeqChunk < n)
      {
        // the first chunk that did not pass: the literal loop, by the warp that holds it; then the rest of the round
        // is judged again against the new sum
        if (warp == (ok >> 2))
        {
          const float Sn = seq_chunk_serial(S, x, ok & 3);
          if (lane == 0)
            sh_S_serial[it] = Sn;
        }
        __syncthreads();
        S = sh_S_serial[it];
        first = ok + 1;
      }
      else
      {
        first = kSeqChunks;
      }
    }
  };
  float xa[kSeqPerLane], xb[kSeqPerLane], xc[kSeqPerLane];
  seq_load(v, n, mine, xa);
  seq_load(v, n, kRound + mine, xb);
  for (uint64_t base = 0; base < n; base += 3 * kRound)
  {
    seq_load(v, n, base + 2 * kRound + mine, xc);
    run_round(base, xa);
    if (base + kRound >= n)
      break;
    seq_load(v, n, base + 3 * kRound + mine, xa);
    run_round(base + kRound, xb);
    if (base + 2 * kRound >= n)
      break;
    seq_load(v, n, base + 4 * kRound + mine, xb);
    run_round(base + 2 * kRound, xc);
  }
  if (threadIdx.x == 0)
    a.out[blockIdx.x] = S;
}

// ---------------------------------------------------------------------------------------------------------------
// Two-phase form of the same evaluation (the default).  patch_seqsum_kernel above carries S from round to round: 51
// dependent rounds of a barrier-separated scan at C2 (206 116 addends), ~85 us however few sequences there are -- the
// largest fixed cost of a small population's step (13 poses per GPU on 8 GPUs).  Here the binade of S at the start
// of every chunk is PREDICTED from a float64 prefix sum, which removes the dependency from everything but a final
// cheap walk, and a thread-block cluster of up to 8 CTAs shares the independent part of one sequence:
//   pass 1 (every CTA of the cluster, its rounds of 16 384 addends): float64 round totals, exchanged through
//     distributed shared memory -> the float64 sum in front of every round;
//   pass 2 (same rounds): a block scan gives every chunk of 128 addends its predicted start value P; with
//     key = sign | exponent of (float)P the chunk's addends are rounded to that binade's grid and the chunk records
//     (sum, min / max prefix, key, bad) exactly as above, into the shared memory of the cluster's CTA 0;
//   walk (warp 0 of CTA 0): goes through the records 32 at a time with the TRUE f32 sum S: a chunk is accepted if its
//     key equals S's sign | exponent and its prefixes stay strictly inside the binade -- the same exactness argument
//     as above; the prediction only decides which grid was prepared, never the result -- otherwise the chunk is
//     summed by the literal loop (128 dependent adds) and the walk goes on with the new S.  Chunks the prediction
//     already expects to fail (S = 0 at the start, binade crossings, ties) keep their addends in shared memory.
// Bitwise equal to the literal loop on the same adversarial sequences (tests/test_gpu_metrics.py), for every
// cluster size.
constexpr int kS2Threads = 1024;
constexpr int kS2Warps = kS2Threads / 32;
constexpr int kS2RoundChunks = kS2Threads / 8;                 // 128 chunks
constexpr int kS2Round = kS2RoundChunks * kSeqChunk;           // 16 384 addends per round
constexpr int kS2Stash = 48;                     // chunks whose addends stay in shared memory for the literal loop
constexpr int kS2Margin = 2048;                  // ulps: how close to a binade edge the prediction is not trusted
constexpr int kS2MaxRounds = 96;                 // per sequence (bounds the shared-memory tables; longer -> chained kernel)
constexpr unsigned kS2BadBit = 0x200u;           // meta: bits 0-8 key (sign | exponent), bit 9 bad, bits 16-23 stash slot + 1
struct __align__(16) S2Rec
{
  int tot, mn, mx;   // pass 2: chunk sum, min / max prefix (grid units); after the window scan: inclusive sum, lo, hi
  unsigned meta;
};
struct S2Fixed   // front of the dynamic shared memory; the records follow
{
  float stash[kS2Stash][kSeqChunk];
  float stage[kSeqChunk];
  double wt[2][kS2Warps];
  double rt[kS2MaxRounds];    // float64 totals of this CTA's rounds (local round index)
  double run[kS2MaxRounds];   // float64 sum in front of this CTA's rounds
  int nstash;
  int pad[3];
};
constexpr size_t kS2FixedSmem = sizeof(S2Fixed);
constexpr size_t kS2MaxSmem = 200 * 1024;
static_assert(kS2FixedSmem % 16 == 0, "records must stay 16-byte aligned");

__global__ void __launch_bounds__(kS2Threads) patch_seqsum2_kernel(const SeqSumArgs a)
{
  namespace cg = cooperative_groups;
  extern __shared__ __align__(16) unsigned char s2_smem[];
  S2Fixed& sh = *reinterpret_cast<S2Fixed*>(s2_smem);
  S2Rec* const recs = reinterpret_cast<S2Rec*>(s2_smem + kS2FixedSmem);   // [n_chunks], used in CTA 0 of the cluster
  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t C = cluster.num_blocks(), rank = cluster.block_rank();
  const uint32_t seq = blockIdx.x / C;
  const float* __restrict__ v = a.vals + (size_t)seq * a.n;
  const uint64_t n = a.n;
  const uint32_t n_chunks = (uint32_t)((n + kSeqChunk - 1) / kSeqChunk);
  const uint32_t n_rounds = (uint32_t)((n + kS2Round - 1) / kS2Round);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, qtr = lane >> 3;
  if (tid == 0)
    sh.nstash = 0;

  // ---- pass 1: float64 totals of my rounds ----
  int par = 0;
  for (uint32_t R = rank, lr = 0; R < n_rounds; R += C, ++lr, par ^= 1)
  {
    float x[kSeqPerLane];
    seq_load(v, n, (uint64_t)R * kS2Round + (uint64_t)tid * kSeqPerLane, x);
    float s[kSeqPerLane / 2];
#pragma unroll
    for (int j = 0; j < kSeqPerLane / 2; ++j)
      s[j] = x[2 * j] + x[2 * j + 1];
    const float ls = ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
    const double d = warp_sum((double)ls);
    if (lane == 0)
      sh.wt[par][warp] = d;
    __syncthreads();  // one barrier per round: wt is double buffered by round parity
    if (warp == 0)
    {
      const double w = warp_sum(sh.wt[par][lane]);
      if (lane == 0)
        sh.rt[lr] = w;
    }
  }
  cluster.sync();
  // the float64 sum in front of each of my rounds: prefix over all rounds' totals (read from their owners)
  if (warp == 0)
  {
    double carry = 0.0;
    for (uint32_t R0 = 0; R0 < n_rounds; R0 += 32)
    {
      const uint32_t R = R0 + lane;
      double val = 0.0;
      if (R < n_rounds)
        val = cluster.map_shared_rank(sh.rt, R % C)[R / C];
      double inc = val;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1)
      {
        const double y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
          inc += y;
      }
      if (R < n_rounds && (R % C) == rank)
        sh.run[R / C] = carry + (inc - val);
      carry += __shfl_sync(0xffffffffu, inc, 31);
    }
  }
  __syncthreads();

  // ---- pass 2: chunk records under the predicted binade, into CTA 0 ----
  S2Rec* const recs0 = cluster.map_shared_rank(recs, 0);
  float* const stash0 = cluster.map_shared_rank(&sh.stash[0][0], 0);
  int* const nstash0 = cluster.map_shared_rank(&sh.nstash, 0);
  for (uint32_t R = rank, lr = 0; R < n_rounds; R += C, ++lr, par ^= 1)
  {
    float x[kSeqPerLane];
    seq_load(v, n, (uint64_t)R * kS2Round + (uint64_t)tid * kSeqPerLane, x);
    float s[kSeqPerLane / 2];
#pragma unroll
    for (int j = 0; j < kSeqPerLane / 2; ++j)
      s[j] = x[2 * j] + x[2 * j + 1];
    double inc = (double)(((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7])));
#pragma unroll
    for (int o = 1; o < 8; o <<= 1)
    {
      const double y = __shfl_up_sync(0xffffffffu, inc, o, 8);
      if ((lane & 7) >= o)
        inc += y;
    }
    const double c0 = __shfl_sync(0xffffffffu, inc, 7), c1 = __shfl_sync(0xffffffffu, inc, 15);
    const double c2 = __shfl_sync(0xffffffffu, inc, 23), c3 = __shfl_sync(0xffffffffu, inc, 31);
    const double qpre = (qtr == 0) ? 0.0 : (qtr == 1) ? c0 : (qtr == 2) ? (c0 + c1) : ((c0 + c1) + c2);
    if (lane == 0)
      sh.wt[par][warp] = ((c0 + c1) + c2) + c3;
    __syncthreads();
    double wi = sh.wt[par][lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const double y = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o)
        wi += y;
    }
    const double wprev = __shfl_sync(0xffffffffu, wi, warp > 0 ? warp - 1 : 0);
    const double P = sh.run[lr] + ((warp > 0) ? wprev : 0.0) + qpre;

    const float S0 = (float)P;
    const unsigned key = __float_as_uint(S0) >> 23;
    const int eb = (int)(key & 0xffu);
    const bool fast = (eb >= 27) && (eb < 227);                             // 2^-100 <= |S0| < 2^100
    // sign(S0) / u = +-2^(23 - e), exact (RN is symmetric: work with |S|, sgn * x)
    const float sscale = __int_as_float(((fast ? (277 - eb) : 127) << 23) | (int)((key & 0x100u) << 23));
    int p = 0, mn = 0x7fffffff, mx = -0x7fffffff - 1;
    float tsum = 0.0f, tie = 1.0f;
#pragma unroll
    for (int j = 0; j < kSeqPerLane; ++j)
    {
      const float t = __fmul_rn(x[j], sscale);            // exact (power of two) or flushed towards 0 when tiny
      const float tm = __fadd_rn(t, 12582912.0f);         // rint by the 1.5 * 2^23 trick (exact for |t| < 2^22)
      const float r = __fadd_rn(tm, -12582912.0f);
      tsum = __fadd_rn(tsum, fabsf(t));                                       // NaN / Inf propagate
      tie = fminf(tie, fabsf(__fadd_rn(fabsf(__fsub_rn(t, r)), -0.5f)));      // 0 iff some t lies exactly half-way
      p += __float_as_int(tm) - 0x4B400000;
      mn = min(mn, p);
      mx = max(mx, p);
    }
    // sum |t| <= 2^21 keeps every t inside the trick's exact range and every integer below far inside int32
    // (|chunk sum| <= 2^24); NaN fails the comparison.  (fminf drops a NaN operand, tsum catches it.)
    const bool bad = !fast || !(tsum <= 2097152.0f) || (tie == 0.0f);
    const bool any_bad = ((__ballot_sync(0xffffffffu, bad) >> (8 * qtr)) & 0xffu) != 0u;
    if (any_bad)
    {
      p = 0;
      mn = 0;
      mx = 0;
    }
    int incl = p;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1)
    {
      const int y = __shfl_up_sync(0xffffffffu, incl, o, 8);
      if ((lane & 7) >= o)
        incl += y;
    }
    const int pre = incl - p;
    mn += pre;
    mx += pre;
#pragma unroll
    for (int o = 1; o < 8; o <<= 1)
    {
      mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o, 8));
      mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o, 8));
    }
    const int tot = __shfl_sync(0xffffffffu, incl, 7, 8);
    const uint32_t gc = R * kS2RoundChunks + (uint32_t)(tid >> 3);
    // does the prediction itself expect this chunk to go to the literal loop?  Then keep its addends on chip.
    int slot = -1;
    if ((lane & 7) == 0 && gc < n_chunks)
    {
      const long long A0 = (long long)((__float_as_uint(S0) & 0x7fffffu) | 0x800000u);
      const bool likely = any_bad || (A0 + mn < (1ll << 23) + 1 + kS2Margin) || (A0 + mx > (1ll << 24) - 1 - kS2Margin);
      if (likely)
      {
        const int sl = atomicAdd(nstash0, 1);
        slot = (sl < kS2Stash) ? sl : -1;
      }
      S2Rec rec;
      rec.tot = tot;
      rec.mn = mn;
      rec.mx = mx;
      rec.meta = key | (any_bad ? kS2BadBit : 0u) | ((unsigned)(slot + 1) << 16);
      recs0[gc] = rec;
    }
    slot = __shfl_sync(0xffffffffu, slot, 0, 8);
    if (slot >= 0)
    {
      float4* dst = reinterpret_cast<float4*>(stash0 + (size_t)slot * kSeqChunk + (lane & 7) * kSeqPerLane);
#pragma unroll
      for (int k = 0; k < kSeqPerLane / 4; ++k)
        dst[k] = make_float4(x[4 * k], x[4 * k + 1], x[4 * k + 2], x[4 * k + 3]);
    }
  }
  cluster.sync();  // all records and stashed chunks have landed in CTA 0; nobody reads a peer's memory after this
  if (rank != 0)
    return;

  // ---- window scan (all warps): per window of 32 chunks the inclusive sum of the chunk sums, and the range of
  // |S| / u at the window's start (minus what the walk has accepted before) for which the chunk's prefixes stay
  // strictly inside the binade.  Units differ between chunks of different keys; only the run of chunks whose key
  // matches S is ever used, and beyond the first rejected chunk nothing is.
  for (uint32_t w0 = 32u * warp; w0 < n_chunks; w0 += 32u * kS2Warps)
  {
    const uint32_t gc = w0 + (uint32_t)lane;
    const bool exists = gc < n_chunks;
    S2Rec rec;
    rec.tot = 0;
    rec.mn = 0;
    rec.mx = 0;
    rec.meta = kS2BadBit;
    if (exists)
      rec = recs[gc];
    int incl = rec.tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
      const int y = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o)
        incl += y;
    }
    if (exists)
    {
      const int excl = incl - rec.tot;
      rec.mn = (int)((1u << 23) + 1u - (unsigned)rec.mn - (unsigned)excl);   // lo: wraps only beyond a rejected chunk
      rec.mx = (int)((1u << 24) - 1u - (unsigned)rec.mx - (unsigned)excl);   // hi
      rec.tot = incl;
      recs[gc] = rec;
    }
  }
  __syncthreads();
  if (warp != 0)
    return;

  // ---- the walk with the true sum ----
  float S = 0.0f;
  for (uint32_t w0 = 0; w0 < n_chunks; w0 += 32)
  {
    const uint32_t gc = w0 + (uint32_t)lane;
    S2Rec rec;
    rec.tot = 0;
    rec.mn = 0x7fffffff;   // lo > hi: never valid
    rec.mx = -0x7fffffff - 1;
    rec.meta = kS2BadBit;
    if (gc < n_chunks)
      rec = recs[gc];
    const int n_here = (int)min(32u, n_chunks - w0);
    int first = 0;
    int base = 0;  // inclusive sum up to chunk first - 1: what lo / hi / tot of the later chunks still contain
    while (first < n_here)
    {
      const unsigned sb = __float_as_uint(S);
      const unsigned key = sb >> 23;
      const int A = (int)((sb & 0x7fffffu) | 0x800000u);  // |S| / u when S is normal
      const int Arel = A - base;
      // a record's key can only equal S's when the prediction was in the fast range and the chunk is not bad
      const bool valid = (lane < first) || (((rec.meta & 0x3ffu) == key) && (Arel >= rec.mn) && (Arel <= rec.mx));
      const unsigned fails = __ballot_sync(0xffffffffu, !valid);
      const int ok = fails ? (__ffs(fails) - 1) : 32;  // >= first
      if (ok > first)
      {
        const int inc = __shfl_sync(0xffffffffu, rec.tot, ok - 1);
        const int A_out = Arel + inc;                                    // in [2^23 + 1, 2^24 - 1]
        const float u = __int_as_float((int)((key & 0xffu) - 23u) << 23);
        const float mag = __fmul_rn((float)A_out, u);                   // exact
        S = (key & 0x100u) ? -mag : mag;
      }
      if (ok < n_here)
      {
        // the literal loop over chunk w0 + ok: every lane runs the same chain on broadcast shared-memory reads
        const unsigned meta = __shfl_sync(0xffffffffu, rec.meta, ok);
        base = __shfl_sync(0xffffffffu, rec.tot, ok);
        const int slot = (int)((meta >> 16) & 0xffu) - 1;
        const float4* src = reinterpret_cast<const float4*>(sh.stage);
        if (slot >= 0)
        {
          src = reinterpret_cast<const float4*>(sh.stash[slot]);
        }
        else
        {
          const uint64_t i0 = (uint64_t)(w0 + ok) * kSeqChunk + 4u * lane;
          float4 q;
          q.x = (i0 < n) ? __ldg(v + i0) : 0.0f;
          q.y = (i0 + 1 < n) ? __ldg(v + i0 + 1) : 0.0f;
          q.z = (i0 + 2 < n) ? __ldg(v + i0 + 2) : 0.0f;
          q.w = (i0 + 3 < n) ? __ldg(v + i0 + 3) : 0.0f;
          __syncwarp();
          reinterpret_cast<float4*>(sh.stage)[lane] = q;
          __syncwarp();
        }
#pragma unroll 8
        for (int k = 0; k < kSeqChunk / 4; ++k)
        {
          const float4 q = src[k];
          S = __fadd_rn(S, q.x);
          S = __fadd_rn(S, q.y);
          S = __fadd_rn(S, q.z);
          S = __fadd_rn(S, q.w);
        }
        first = ok + 1;
      }
      else
      {
        first = n_here;
      }
    }
  }
  if (lane == 0)
    a.out[seq] = S;
}

__global__ void patch_seqsum_serial_kernel(const SeqSumArgs a)
{
  const float* __restrict__ v = a.vals + (size_t)blockIdx.x * a.n;
  float S = 0.0f;
  for (uint64_t k = 0; k < a.n; ++k)
    S = __fadd_rn(S, v[k]);
  a.out[blockIdx.x] = S;
}

int launch_seqsum(const SeqSumArgs& a, cudaStream_t st)
{
  if (!a.n_seq)
    return XRC_OK;
  const size_t n_chunks = (size_t)((a.n + kSeqChunk - 1) / kSeqChunk);
  const size_t n_rounds = (size_t)((a.n + kS2Round - 1) / kS2Round);
  const size_t smem2 = kS2FixedSmem + n_chunks * sizeof(S2Rec);
  if (a.serial == 1)
  {
    patch_seqsum_serial_kernel<<<a.n_seq, 1, 0, st>>>(a);
  }
  else if (a.serial == 2 || smem2 > kS2MaxSmem || n_rounds > (size_t)kS2MaxRounds || (a.serial == 0 && a.n_seq > 74u))
  {
    // many sequences fill the GPU by themselves: the chain's 256-thread CTAs (several per SM) then beat the two-phase
    // form's 1024-thread CTAs (measured at C2's length: 100 sequences 85 vs 94 us, 200 sequences 104 vs 178 us)
    patch_seqsum_kernel<<<a.n_seq, kSeqWarps * 32, 0, st>>>(a);   // the round-to-round chain (any length)
  }
  else
  {
    static bool attr_set[64] = {};
    int dev = 0;
    XRC_CUDA(cudaGetDevice(&dev));
    if (dev >= 0 && dev < 64 && !attr_set[dev])
    {
      XRC_CUDA(cudaFuncSetAttribute(patch_seqsum2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kS2MaxSmem));
      attr_set[dev] = true;
    }
    // CTAs per sequence: few sequences (a small population) spread each one over a cluster; serial = 10 + c forces c
    unsigned c = (a.n_seq <= 18u) ? 8u : (a.n_seq <= 37u) ? 4u : (a.n_seq <= 74u) ? 2u : 1u;
    if (a.serial >= 11 && a.serial <= 18)
      c = (unsigned)(a.serial - 10);
    while (c > 1u && (c > n_rounds || (c & (c - 1u))))
      --c;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(a.n_seq * c);
    cfg.blockDim = dim3(kS2Threads);
    cfg.dynamicSmemBytes = smem2;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    XRC_CUDA(cudaLaunchKernelEx(&cfg, patch_seqsum2_kernel, a));
  }
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

// Patch subsets (set_patches_to_use / random patches): dst[seq][j] = vals[seq][subset[j]], the per-patch values of the
// LOCAL patch list in its own order (xregImgSimMetric2DPatchNCCCPU.cpp:204), ready for patch_seqsum_kernel
__global__ void patch_gather_kernel(const float* __restrict__ vals, const uint32_t* __restrict__ subset, float* __restrict__ dst,
                                    uint64_t n_patches, uint32_t n_subset)
{
  const float* __restrict__ v = vals + (size_t)blockIdx.y * n_patches;
  float* __restrict__ d = dst + (size_t)blockIdx.y * n_subset;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n_subset; j += gridDim.x * blockDim.x)
    d[j] = v[subset[j]];
}

int launch_patch_gather(const float* vals, const uint32_t* subset, float* dst, uint64_t n_patches, uint32_t n_subset,
                        uint32_t n_seq, cudaStream_t st)
{
  if (!n_seq || !n_subset)
    return XRC_OK;
  const uint32_t bx = std::min<uint32_t>((n_subset + 255u) / 256u, 64u);
  patch_gather_kernel<<<dim3(bx, n_seq), 256, 0, st>>>(vals, subset, dst, n_patches, n_subset);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Log remap of a projection (pre-processing, once per fixed image): itk::DiscreteGaussianImageFilter's passes for the
// default I0 (restated from ITK 5.1.1's published algorithm, DESIGN.md section 4.7: double accumulation over the taps in
// kernel order, zero-flux Neumann border, float image between the passes), the two reductions, and the map itself.
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) itk_gauss_kernel(const ItkGaussArgs a)
{
  const size_t n = (size_t)a.rows * a.cols;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const int64_t y = (int64_t)(i / a.cols), x = (int64_t)(i % a.cols);
    double s = 0.0;
    for (int t = 0; t <= 2 * a.radius; ++t)
    {
      int64_t yy = y, xx = x;
      if (a.along_x)
        xx = min(max(x + t - a.radius, (int64_t)0), (int64_t)a.cols - 1);
      else
        yy = min(max(y + t - a.radius, (int64_t)0), (int64_t)a.rows - 1);
      s = __dadd_rn(s, __dmul_rn(a.k[t], (double)__ldg(a.src + (size_t)yy * a.cols + xx)));
    }
    a.dst[i] = (float)s;
  }
}

int launch_itk_gauss(const ItkGaussArgs& a, cudaStream_t st)
{
  const size_t n = (size_t)a.rows * a.cols;
  if (!n)
    return XRC_OK;
  itk_gauss_kernel<<<(unsigned)std::min<size_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(a);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

__global__ void __launch_bounds__(1024) minmax_kernel(const float* __restrict__ src, uint64_t n, float scale, float eps, float* out2)
{
  __shared__ float s_mx[32], s_mn[32];
  float mx = -3.402823466e+38f, mn = 3.402823466e+38f;   // mn: smallest value > eps
  for (uint64_t i = threadIdx.x; i < n; i += 1024)
  {
    const float v = __fmul_rn(__ldg(src + i), scale);
    mx = fmaxf(mx, v);
    if (v > eps)
      mn = fminf(mn, v);
  }
  for (int o = 16; o > 0; o >>= 1)
  {
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0)
  {
    s_mx[threadIdx.x >> 5] = mx;
    s_mn[threadIdx.x >> 5] = mn;
  }
  __syncthreads();
  if (threadIdx.x < 32)
  {
    mx = s_mx[threadIdx.x];
    mn = s_mn[threadIdx.x];
    for (int o = 16; o > 0; o >>= 1)
    {
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    }
    if (threadIdx.x == 0)
    {
      out2[0] = mx;
      out2[1] = (mn == 3.402823466e+38f) ? 0.0f : mn;   // no positive pixel: min_pos stays 0 (:112)
    }
  }
}

int launch_minmax(const float* src, uint64_t n, float scale, float eps, float* out2, cudaStream_t st)
{
  minmax_kernel<<<1, 1024, 0, st>>>(src, n, scale, eps, out2);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

__global__ void __launch_bounds__(256) log_map_kernel(const float* __restrict__ src, float* __restrict__ dst, uint64_t n,
                                                      float scale, int log_mode, float eps, float I0, float out_max)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
  {
    const float v = __fmul_rn(__ldg(src + i), scale);
    // -std::log(x / I0) in float (:140): the quotient is an IEEE division, the logarithm is evaluated in double and
    // rounded once (glibc's logf is correctly rounded in all but rare cases; CUDA's logf is only within 1 ulp)
    dst[i] = log_mode ? ((v > eps) ? -(float)log((double)__fdiv_rn(v, I0)) : out_max) : v;
  }
}

int launch_log_map(const float* src, float* dst, uint64_t n, float scale, int log_mode, float eps, float I0, float out_max,
                   cudaStream_t st)
{
  if (!n)
    return XRC_OK;
  log_map_kernel<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(src, dst, n, scale, log_mode, eps, I0,
                                                                                         out_max);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// Down-sampling of a projection: itk::BSplineDecompositionImageFilter + cubic BSplineInterpolateImageFunction restated
// (DESIGN.md section 4.8).  One thread per image line for the recursive prefilter (sequential by nature, a few hundred
// lines of a few hundred pixels, once per registration level), one thread per output pixel for the 16-tap evaluation;
// double arithmetic, uncontracted, in the oracle's operation order.
// ---------------------------------------------------------------------------------------------------------------
__global__ void f32_to_f64_kernel(const float* __restrict__ src, double* __restrict__ dst, uint64_t n)
{
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    dst[i] = (double)src[i];
}

int launch_f32_to_f64(const float* src, double* dst, uint64_t n, cudaStream_t st)
{
  if (!n)
    return XRC_OK;
  f32_to_f64_kernel<<<(unsigned)std::min<uint64_t>((n + 255) / 256, 148 * 8), 256, 0, st>>>(src, dst, n);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

__global__ void __launch_bounds__(64) bspline_prefilter_kernel(double* __restrict__ c, uint32_t rows, uint32_t cols, int along_x,
                                                               double zpow, long long horizon)
{
  const uint32_t line = blockIdx.x * blockDim.x + threadIdx.x;
  const long long n = along_x ? cols : rows;
  const uint32_t n_lines = along_x ? rows : cols;
  if (line >= n_lines || n == 1)
    return;
  const size_t stride = along_x ? 1 : cols;
  double* p = c + (along_x ? (size_t)line * cols : (size_t)line);
#define C_(i) p[(size_t)(i) * stride]
  const double z = __dsub_rn(sqrt(3.0), 2.0);
  const double c0 = __dmul_rn(__dsub_rn(1.0, z), __dsub_rn(1.0, __ddiv_rn(1.0, z)));
  for (long long i = 0; i < n; ++i)
    C_(i) = __dmul_rn(C_(i), c0);
  {
    double zn = z, sum;
    if (horizon < n)
    {
      sum = C_(0);
      for (long long i = 1; i < horizon; ++i)
      {
        sum = __dadd_rn(sum, __dmul_rn(zn, C_(i)));
        zn = __dmul_rn(zn, z);
      }
      C_(0) = sum;
    }
    else
    {
      const double iz = __ddiv_rn(1.0, z);
      double z2n = zpow;
      sum = __dadd_rn(C_(0), __dmul_rn(z2n, C_(n - 1)));
      z2n = __dmul_rn(z2n, __dmul_rn(z2n, iz));
      for (long long i = 1; i <= n - 2; ++i)
      {
        sum = __dadd_rn(sum, __dmul_rn(__dadd_rn(zn, z2n), C_(i)));
        zn = __dmul_rn(zn, z);
        z2n = __dmul_rn(z2n, iz);
      }
      C_(0) = __ddiv_rn(sum, __dsub_rn(1.0, __dmul_rn(zn, zn)));
    }
  }
  for (long long i = 1; i < n; ++i)
    C_(i) = __dadd_rn(C_(i), __dmul_rn(z, C_(i - 1)));
  C_(n - 1) = __dmul_rn(__ddiv_rn(z, __dsub_rn(__dmul_rn(z, z), 1.0)), __dadd_rn(__dmul_rn(z, C_(n - 2)), C_(n - 1)));
  for (long long i = n - 2; i >= 0; --i)
    C_(i) = __dmul_rn(z, __dsub_rn(C_(i + 1), C_(i)));
#undef C_
}

int launch_bspline_prefilter(double* c, uint32_t rows, uint32_t cols, int along_x, double zpow, int64_t horizon, cudaStream_t st)
{
  const uint32_t n_lines = along_x ? rows : cols;
  if (!n_lines)
    return XRC_OK;
  bspline_prefilter_kernel<<<(n_lines + 63) / 64, 64, 0, st>>>(c, rows, cols, along_x, zpow, (long long)horizon);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

__device__ __forceinline__ long long mirror_index(long long i, long long n)
{
  if (n == 1)
    return 0;
  if (i < 0)
    i = -i;
  if (i >= n)
    i = (n - 1) - (i - (n - 1));
  return i;
}

__global__ void __launch_bounds__(256) bspline_resample_kernel(const double* __restrict__ c, uint32_t rows, uint32_t cols,
                                                               float* __restrict__ out, uint32_t orows, uint32_t ocols, double factor)
{
  const size_t n_out = (size_t)orows * ocols;
  for (size_t o = (size_t)blockIdx.x * blockDim.x + threadIdx.x; o < n_out; o += (size_t)gridDim.x * blockDim.x)
  {
    const long long oy = (long long)(o / ocols), ox = (long long)(o % ocols);
    const double xs[2] = {__ddiv_rn((double)ox, factor), __ddiv_rn((double)oy, factor)};
    const long long len[2] = {(long long)cols, (long long)rows};
    float v = 0.0f;
    if (xs[0] >= -0.5 && xs[0] < (double)cols - 0.5 && xs[1] >= -0.5 && xs[1] < (double)rows - 0.5)
    {
      double w[2][4];
      long long idx[2][4];
#pragma unroll
      for (int d = 0; d < 2; ++d)
      {
        const long long i0 = (long long)floorf((float)xs[d]) - 1;
        const double t = __dsub_rn(xs[d], (double)(i0 + 1));
        w[d][3] = __dmul_rn(__dmul_rn(__dmul_rn(1.0 / 6.0, t), t), t);
        w[d][0] = __dsub_rn(__dadd_rn(1.0 / 6.0, __dmul_rn(__dmul_rn(0.5, t), __dsub_rn(t, 1.0))), w[d][3]);
        w[d][2] = __dsub_rn(__dadd_rn(t, w[d][0]), __dmul_rn(2.0, w[d][3]));
        w[d][1] = __dsub_rn(__dsub_rn(__dsub_rn(1.0, w[d][0]), w[d][2]), w[d][3]);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          idx[d][k] = mirror_index(i0 + k, len[d]);
      }
      double acc = 0.0;
#pragma unroll
      for (int pp = 0; pp < 16; ++pp)
      {
        const int kx = pp & 3, ky = pp >> 2;
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(w[0][kx], w[1][ky]), c[(size_t)idx[1][ky] * cols + (size_t)idx[0][kx]]));
      }
      v = (float)acc;
    }
    out[o] = v;
  }
}

int launch_bspline_resample(const double* c, uint32_t rows, uint32_t cols, float* out, uint32_t orows, uint32_t ocols,
                            double factor, cudaStream_t st)
{
  const size_t n_out = (size_t)orows * ocols;
  if (!n_out)
    return XRC_OK;
  bspline_resample_kernel<<<(unsigned)std::min<size_t>((n_out + 255) / 256, 148 * 8), 256, 0, st>>>(c, rows, cols, out, orows,
                                                                                                      ocols, factor);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

int launch_patch_finalize(const PatchFinalizeArgs& a, cudaStream_t st)
{
  if (!a.n_imgs)
    return XRC_OK;
  patch_finalize_kernel<<<a.n_imgs, 32, 0, st>>>(a);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

}  // namespace xrc
