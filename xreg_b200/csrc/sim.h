// Launchers of the similarity-metric kernels (internal; see sim.cu).
#pragma once

#include <cstdlib>

#include "common.h"

namespace xrc
{

constexpr int kGradTile = 32;       // output tile edge of the gradient kernel
constexpr int kMaxGaussWidth = 31;  // widest supported smoothing kernel
constexpr int kGradBandRows = 64;   // output rows per warp of the fast gradient kernel (throughput regime)
constexpr int kGradBandRowsMin = 8; // ... when only a few images are in flight (population 1)
constexpr int kPatchThreads = 256;  // threads per CTA of the patch kernel (each owns 1 or 2 input columns)
constexpr int kPatchBandRowsMin = 8;  // ... with few images in flight (population 1)
constexpr int kMomChunk = 4096;     // pixels per CTA of the plain moments kernel

struct GradArgs
{
  const float* src;       // n_imgs x rows x cols
  uint32_t n_imgs, rows, cols;
  int gauss_width;        // 0 = off
  float coeffs[kMaxGaussWidth];
  float* gx;              // optional outputs, n_imgs x rows x cols
  float* gy;
  // optional NCC moments against zero-mean fixed gradients
  const float* f0x;
  const float* f0y;
  const uint8_t* mask;
  double* partials;       // n_imgs x n_tiles x 6
  uint32_t tiles_x, tiles_y;
  uint32_t band_rows;     // fast kernel: output rows per warp (filled by the launcher)
};

struct MomentArgs
{
  const float* src;  // n_imgs x npix
  uint32_t n_imgs;
  uint64_t npix;
  const float* f0;
  const uint8_t* mask;
  double* partials;  // n_imgs x n_chunks x 3
  uint32_t n_chunks;
};

// per image: sims[i] from (Sm, Smm, Smf) partial sums, for n_dirs directions
struct NccFinalizeArgs
{
  const double* partials;  // n_imgs x n_parts x (3 * n_dirs)
  uint32_t n_imgs, n_parts, n_dirs;
  double n_eff;            // N or mask_len
  double sf0[2];           // sum of zero-mean fixed (over mask)
  float f_sd[2];
  float* sims;
  float* sims_host;        // optional host-mapped copy (pinned): saves the D2H memcpy of the scalars
  int ssd;                 // 1: sim = (Smm - 2 Smf + sf0[0]) / n_eff with f0 = the (masked) fixed image, sf0[0] = sum f^2
};

struct PatchArgs
{
  const float* mov[2];     // per direction: n_imgs x rows x cols
  const float* fix[2];     // per direction: rows x cols
  const uint8_t* mask;     // rows x cols or null
  uint32_t n_imgs, n_dirs, rows, cols;
  uint32_t radius, stride;
  uint32_t n_strips, n_parts;  // filled by the launcher from patch_plan(); n_parts = n_strips * n_bands
  uint32_t band_rows;          // ditto
  int mask_mode;           // 0 none, 1 mask in correlation only, 2 mask in stats too
  // fixed per-patch statistics on the stride-1 grid (rows-2r) x (cols-2r), per direction
  const double* f_mean[2];  // f64: keeps sum(m - mu_m)(f - mu_f) = Smf - mu_f Sm free of f32 rounding of mu_f
  const float* f_den[2];   // sigma_f * n
  const double* f_smask[2]; // sum over mask of f (mask modes)
  const float* n_mask;     // mask count per patch (mask modes)
  const float* weights;    // per strided patch or null
  int weight_patch_sims;
  double* partials;        // n_imgs x n_dirs x n_parts
  // reference-order combine: per-patch values w_k * s_k (0 for skipped patches) in the reference's patch order,
  // [img][dir][strided patch], or null (f64 combine through `partials` only)
  float* vals;
  uint64_t n_patches;      // strided patch count (stride of `vals` per image and direction)
  // fixed-stats mode outputs (when mov[0] == nullptr)
  double* o_mean[2];
  float* o_den[2];
  double* o_smask[2];
  float* o_nmask;
};

struct PatchFinalizeArgs
{
  const double* partials;
  uint32_t n_imgs, n_dirs, n_parts;
  double divisor;  // num_patches (mean), total weight, or 1
  // reference-order combine: seq_sums[img * n_dirs + dir] = the f32 sequential sum of the per-patch values
  // (patch_seqsum), divided in f32 by divisor_f when divide != 0 (xregImgSimMetric2DPatchNCCCPU.cpp:262-287)
  const float* seq_sums;
  float divisor_f;
  int divide;
  float* sims;
  float* sims_host;  // optional host-mapped copy (pinned)
};

// out[s] = (((0 + v[s][0]) + v[s][1]) + ...) in f32, round to nearest even at every step: the reference's
// `patch_sims_sum += s` loop.  serial != 0: one thread runs the literal loop (verification of the parallel emulation).
struct SeqSumArgs
{
  const float* vals;   // n_seq x n
  uint64_t n;
  uint32_t n_seq;
  float* out;          // n_seq
  int serial;         // 0: two-phase emulation, 1: literal loop, 2: chained emulation
};
int launch_seqsum(const SeqSumArgs& a, cudaStream_t st);
// dst[seq][j] = vals[seq][subset[j]] for n_seq sequences of n_patches values (patch subsets)
int launch_patch_gather(const float* vals, const uint32_t* subset, float* dst, uint64_t n_patches, uint32_t n_subset,
                        uint32_t n_seq, cudaStream_t st);

int launch_grad(const GradArgs& a, cudaStream_t st);
int launch_moments(const MomentArgs& a, cudaStream_t st);
int launch_ncc_finalize(const NccFinalizeArgs& a, cudaStream_t st);
int launch_patch(const PatchArgs& a, cudaStream_t st);
int launch_patch_fixed_stats(const PatchArgs& a, cudaStream_t st);
int launch_patch_finalize(const PatchFinalizeArgs& a, cudaStream_t st);

// ---- log remap of a projection (ImageIntensLogTransFilter, lib/image/xregImageIntensLogTrans.cpp:55-144)
struct ItkGaussArgs
{
  const float* src;
  float* dst;
  uint32_t rows, cols;
  int radius, along_x;
  double k[65];   // symmetric kernel, 2 radius + 1 taps (itk::GaussianOperator, max width 32)
};
int launch_itk_gauss(const ItkGaussArgs& a, cudaStream_t st);
// out[0] = max over all pixels, out[1] = smallest pixel > eps (0 if none), of src * scale (scale == 1: src itself)
int launch_minmax(const float* src, uint64_t n, float scale, float eps, float* out2, cudaStream_t st);
// dst = src * scale (log_mode 0) or (v > eps ? -log(v / I0) : out_max) with v = src * scale (log_mode 1)
int launch_log_map(const float* src, float* dst, uint64_t n, float scale, int log_mode, float eps, float I0, float out_max,
                   cudaStream_t st);

// ---- down-sampling of a projection (DownsampleImage, lib/itk/xregITKResampleUtils.h:49-112)
// cubic B-spline prefilter (itk::BSplineDecompositionImageFilter) of a rows x cols double image, in place, along x or y;
// zpow = z^(n - 1) for the line length n (only short lines use it; computed on the host like the oracle's pow)
int launch_bspline_prefilter(double* c, uint32_t rows, uint32_t cols, int along_x, double zpow, int64_t horizon, cudaStream_t st);
int launch_f32_to_f64(const float* src, double* dst, uint64_t n, cudaStream_t st);
// out[oy][ox] = cubic B-spline value at the continuous input index (ox / factor, oy / factor), 0 outside the buffer
int launch_bspline_resample(const double* c, uint32_t rows, uint32_t cols, float* out, uint32_t orows, uint32_t ocols,
                            double factor, cudaStream_t st);

// Gaussian widths served by the warp-streaming gradient kernel (the reference's apps use 5; 0 = no smoothing)
inline bool grad_fast_path(int gauss_width) { return gauss_width <= 1 || gauss_width == 3 || gauss_width == 5 || gauss_width == 7; }
// rows per warp of the fast kernel: long bands amortise the halo rows, short bands give a small batch
// enough warps to hide the per-row load latency (population 1: 61 us with 64-row bands at 384 x 384)
inline uint32_t grad_band_rows(uint32_t rows, uint32_t cols, int gauss_width, uint32_t n_imgs)
{
  const uint32_t ow = 32 - 2 * ((gauss_width > 1 ? gauss_width / 2 : 0) + 1);
  const uint64_t strips = (uint64_t)((cols + ow - 1) / ow) * n_imgs;
  uint32_t band = kGradBandRows;
  while (band > (uint32_t)kGradBandRowsMin && strips * ((rows + band - 1) / band) < 148u * 32u)
    band /= 2;
  return band;
}
// per-image partial-sum slots the gradient kernel writes (Grad-NCC moments)
inline uint32_t grad_num_parts(uint32_t rows, uint32_t cols, int gauss_width, uint32_t band_rows = kGradBandRows)
{
  if (grad_fast_path(gauss_width))
  {
    const uint32_t ow = 32 - 2 * ((gauss_width > 1 ? gauss_width / 2 : 0) + 1);
    return ((cols + ow - 1) / ow) * ((rows + band_rows - 1) / band_rows);
  }
  return ((rows + kGradTile - 1) / kGradTile) * ((cols + kGradTile - 1) / kGradTile);
}

// decomposition of the patch grid into CTAs (column strips x row bands)
struct PatchPlan
{
  uint32_t cols_per_thread, n_strips, n_bands, band_rows;
};
// n_units = images x directions in flight: small batches get short bands (more CTAs; each band re-reads the
// d - 1 rows above it, loads only) so that the serial row loop is not the latency of a population-1 evaluation
// resident CTAs per SM the plan counts on (3: every variant reaches it; XRC_PATCH_SLOTS overrides, measurement only)
inline uint32_t patch_ctas_per_sm()
{
  static const uint32_t k = [] {
    const char* e = getenv("XRC_PATCH_SLOTS");
    const int v = e ? atoi(e) : 0;
    return (uint32_t)((v >= 1 && v <= 8) ? v : 3);
  }();
  return k;
}
inline PatchPlan patch_plan(uint32_t rows, uint32_t cols, uint32_t radius, uint32_t n_units)
{
  PatchPlan p;
  p.cols_per_thread = (cols > (uint32_t)kPatchThreads) ? 2u : 1u;
  const uint32_t w_out = p.cols_per_thread * kPatchThreads - 2 * radius;
  p.n_strips = (cols - 2 * radius + w_out - 1) / w_out;
  const uint32_t nrr = rows - 2 * radius;
  // Every band costs its rows plus the 2 r rows above them (loads only: counted half), and the grid runs in waves
  // of 148 SMs x 3 resident CTAs.  Pick the number of bands that minimises waves x rows per band; bands shorter
  // than kPatchBandRowsMin are not worth their pre-roll.
  const uint64_t slots = 148u * patch_ctas_per_sm();
  uint64_t best_cost = ~0ull;
  uint32_t best_bands = 1;
  const uint32_t max_bands = (nrr + kPatchBandRowsMin - 1) / kPatchBandRowsMin;
  for (uint32_t nb = 1; nb <= max_bands; ++nb)
  {
    const uint32_t br = (nrr + nb - 1) / nb;
    const uint32_t bands = (nrr + br - 1) / br;
    const uint64_t ctas = (uint64_t)p.n_strips * bands * n_units;
    const uint64_t waves = (ctas + slots - 1) / slots;
    const uint64_t cost = waves * (2ull * br + 2ull * radius);
    if (cost < best_cost)
    {
      best_cost = cost;
      best_bands = bands;
      p.band_rows = br;
    }
    if (ctas >= 8 * slots)
      break;  // many waves already: finer bands only add pre-roll
  }
  p.n_bands = best_bands;
  return p;
}
// upper bound of n_strips * n_bands over all batch sizes (sizes the partial-sum buffer)
inline uint32_t patch_max_parts(uint32_t rows, uint32_t cols, uint32_t radius)
{
  const PatchPlan p = patch_plan(rows, cols, radius, 1u << 30);
  return p.n_strips * ((rows - 2 * radius + kPatchBandRowsMin - 1) / kPatchBandRowsMin);
}

}  // namespace xrc
