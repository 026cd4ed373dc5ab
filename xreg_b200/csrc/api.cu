// C ABI of libxreg_cuda.so (include/xreg_cuda.h): handle management, argument
// validation, error strings, stream-ordered orchestration of the kernels in
// drr.cu / sim.cu.  Host-side set-up arithmetic (affine inverse, Gaussian
// coefficients, fixed-image statistics) follows the same reference lines as
// the oracle but is written independently of it; nothing here touches oracle/.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <limits>
#include <memory>

#include "common.h"
#include "sim.h"

namespace xrc
{

static thread_local std::string g_err;
std::atomic<uint64_t> g_launch_count{0};

void set_error(const std::string& msg) { g_err = msg; }

// Transform<float,3,Affine>::inverse() as used at xregRayCastLineIntCPU.cpp:309:
// cofactor inverse of the linear part, translation = -(inv * t).  f32, no FMA
// (host x86-64 baseline code generation).
static inline float cof3(const float m[9], int i, int j)
{
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  const volatile float a = m[3 * i1 + j1] * m[3 * i2 + j2];
  const volatile float b = m[3 * i1 + j2] * m[3 * i2 + j1];
  return a - b;
}

void affine_inverse_f32(const float a[12], float out[12])
{
  float m[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      m[3 * r + c] = a[4 * r + c];
  const float c0 = cof3(m, 0, 0), c1 = cof3(m, 1, 0), c2 = cof3(m, 2, 0);
  const volatile float p0 = c0 * m[0], p1 = c1 * m[3], p2 = c2 * m[6];
  const volatile float s01 = p0 + p1;
  const float det = s01 + p2;
  const float invdet = 1.0f / det;
  float mi[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      mi[3 * i + j] = cof3(m, j, i) * invdet;
  const float t[3] = {a[3], a[7], a[11]};
  for (int r = 0; r < 3; ++r)
  {
    for (int c = 0; c < 3; ++c)
      out[4 * r + c] = mi[3 * r + c];
    const volatile float q0 = mi[3 * r] * t[0], q1 = mi[3 * r + 1] * t[1], q2 = mi[3 * r + 2] * t[2];
    const volatile float q01 = q0 + q1;
    out[4 * r + 3] = -(q01 + q2);
  }
}

// cv::getGaussianKernel(n, sigma<=0, CV_32F) of OpenCV 3.4 (fixed tables for n <= 7)
static int gauss_coeffs(int width, float* cf)
{
  static const float t1[] = {1.f};
  static const float t3[] = {0.25f, 0.5f, 0.25f};
  static const float t5[] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
  static const float t7[] = {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f};
  const float* tab = (width == 1) ? t1 : (width == 3) ? t3 : (width == 5) ? t5 : (width == 7) ? t7 : nullptr;
  if (tab)
  {
    memcpy(cf, tab, sizeof(float) * width);
    return 0;
  }
  const double sigma = ((width - 1) * 0.5 - 1) * 0.3 + 0.8;
  const double scale2x = -0.5 / (sigma * sigma);
  double sum = 0;
  for (int i = 0; i < width; ++i)
  {
    const double x = i - (width - 1) * 0.5;
    cf[i] = (float)exp(scale2x * x * x);
    sum += cf[i];
  }
  sum = 1.0 / sum;
  for (int i = 0; i < width; ++i)
    cf[i] = (float)(cf[i] * sum);
  return 0;
}

}  // namespace xrc

using namespace xrc;

// ----------------------------------------------------------------------------
// handles
// ----------------------------------------------------------------------------
struct xrc_ctx
{
  int device = 0;
  cudaStream_t stream = nullptr;
  bool owns_stream = false;
};

struct xrc_rc
{
  xrc_ctx* ctx = nullptr;
  int layout = XRC_LAYOUT_DEFAULT;
  std::vector<DeviceVolume> vols;
  std::vector<xrc_cam> cams;
  xrc_cam* d_cams = nullptr;
  uint32_t rows = 0, cols = 0;
  uint32_t max_projs = 0, num_projs = 0;
  bool allocated = false;
  float* d_buf_own = nullptr;   // own projection buffer; a borrower (use_other_proj_buf) has none
  xrc_rc* other = nullptr;      // lender of the projection buffer, resolved at every use (rc_proj_buf)
  uint32_t n_borrowers = 0;     // ray casters whose `other` is this one: it cannot be destroyed before them
  // multi-GPU tile sharding (xrc_rc_peer_attach): the projection buffers of all ranks, peers' opened over CUDA IPC
  uint32_t peer_n = 0, peer_rank = 0;
  float* peer_bufs[kMaxPeers] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // contiguous tile ranges of the ranks, balanced by measured work (xrc_rc_plan_tiles): rank r owns tiles
  // [tile_begin[r], tile_begin[r + 1]); empty until planned
  std::vector<uint32_t> tile_begin;
  // per-tile cost multipliers learnt from the ranks' measured kernel times (xrc_rc_plan_tiles_timed); empty = all 1
  std::vector<double> tile_mult;
  std::vector<double> tile_work;   // the work estimate of the last plan (samples + set-up), before the multipliers
  // exchange block at the tail of the own projection buffer (xchg.cu): flags + the gathered similarity values
  size_t xchg_off = 0;             // byte offset of the block in d_buf_own (the same on every rank: same sizes)
  uint32_t xchg_epoch = 0;         // the last epoch this rank signalled; all ranks count the same calls
  float* h_gather = nullptr;       // host-mapped: max_projs gathered values, then one status word
  float* h_poses = nullptr;   // pinned staging: max_projs x 12 floats
  uint32_t* h_cam_idx = nullptr;
  float* d_poses = nullptr;
  uint32_t* d_cam_idx = nullptr;
  uint32_t* d_zero_idx = nullptr;      // all-zero camera indices
  const float* ext_poses = nullptr;    // caller-owned device poses (xrc_rc_set_poses_device)
  const uint32_t* ext_cam_idx = nullptr;
  bool ext_mirrored = false;           // h_poses / h_cam_idx hold a host copy of ext_poses (xrc_rc_set_poses_device_mirrored)
  cudaEvent_t staged = nullptr;
  bool staged_pending = false;
  float step_size = 1.0f;
  int interp = XRC_INTERP_LINEAR;
  int kernel_id = XRC_KERNEL_SUM;
  int store_method = XRC_STORE_REPLACE;
  float default_bg = 0.0f;
  bool use_bg = false;
  float* d_bg = nullptr;
  int order = 0;
  bool skip_empty = true;  // empty-space trimming (exact; xrc_rc_set_skip_empty)
  int gap_mode = 0;        // interior gaps: 0 no, 1 yes (xrc_rc_set_skip_empty(rc, 3))
  // latency regime (xrc_obj_fn with <= kInlinePoses projections): the poses in h_poses / h_cam_idx travel in the
  // kernel parameters; d_poses is refreshed lazily if another entry point needs it
  bool inline_poses = false;
};

struct xrc_sm
{
  xrc_ctx* ctx = nullptr;
  int kind = XRC_SM_NCC;
  uint32_t rows = 0, cols = 0;
  std::vector<float> h_fixed;
  std::vector<uint8_t> h_mask;
  bool has_mask = false;
  bool fixed_dirty = true;  // fixed image or mask changed: re-derive fixed statistics
  uint32_t gauss_width = 5;
  // patch parameters (ImgSimMetric2DPatchCommon.h defaults)
  uint32_t radius = 5, stride = 1;
  int compute_mean = 0, weight_sims = 1, mask_stats = 0;
  std::vector<float> h_weights;
  bool has_weights = false;
  // binding
  xrc_rc* rc = nullptr;
  const float* host_src = nullptr;
  const float* dev_src = nullptr;
  uint32_t proj_offset = 0;
  // resources
  bool allocated = false;
  uint32_t max_imgs = 0, n_imgs = 0;
  float* d_fixed = nullptr;
  uint8_t* d_mask = nullptr;
  float* d_f0[2] = {nullptr, nullptr};  // zero-mean fixed (NCC) / fixed gradients
  float* d_fg[2] = {nullptr, nullptr};  // raw fixed gradients (patch-grad) or alias of d_fixed
  double sf0[2] = {0, 0};
  float f_sd[2] = {1, 1};
  double n_eff = 0;
  float* d_mov = nullptr;               // staging for host-bound moving images
  float* d_g[2] = {nullptr, nullptr};   // moving gradient images (patch-grad)
  double* d_partials = nullptr;
  size_t partials_len = 0;
  float* d_sims = nullptr;
  float* h_sims = nullptr;              // pinned
  // patch fixed stats (stride-1 grid)
  double* d_pmean[2] = {nullptr, nullptr};
  float* d_pden[2] = {nullptr, nullptr};
  double* d_psmask[2] = {nullptr, nullptr};
  float* d_pnmask = nullptr;
  float* d_weights = nullptr;
  double divisor = 1.0;
  uint32_t n_strips = 0;
  // reference-order combine of the per-patch values (XRC_COMBINE_*)
  int combine_mode = XRC_COMBINE_REFERENCE;
  float* d_vals = nullptr;      // max_imgs x n_dirs x n_patches
  float* d_seq = nullptr;       // max_imgs x n_dirs
  uint64_t n_patches = 0;
  float divisor_f = 1.0f;
  int divide_f = 0;
  // patch subset (ImgSimMetric2DPatchCommon::set_patches_to_use / random patches): local list of global patch indices
  std::vector<uint32_t> h_subset;
  bool subset_dirty = false;
  uint32_t* d_subset = nullptr;
  size_t d_subset_cap = 0;
  float* d_sub_vals = nullptr;   // max_imgs x n_dirs x |subset|
  size_t d_sub_vals_cap = 0;
  float sub_divisor_f = 1.0f;
};

// projection buffer in use: own, or the lender's as it is NOW (the lender may have been re-allocated since)
static float* rc_proj_buf(const xrc_rc* rc)
{
  const xrc_rc* r = rc;
  for (int hops = 0; r->other && hops < 8; ++hops)
    r = r->other;
  return r->d_buf_own;
}

static uint32_t rc_proj_capacity(const xrc_rc* rc)
{
  const xrc_rc* r = rc;
  for (int hops = 0; r->other && hops < 8; ++hops)
    r = r->other;
  return r->allocated ? r->max_projs : 0u;
}

static int use_device(const xrc_ctx* ctx)
{
  XRC_CUDA(cudaSetDevice(ctx->device));
  return XRC_OK;
}

template <class T>
static void dfree(T*& p)
{
  if (p)
    cudaFree(p);
  p = nullptr;
}

extern "C" {

const char* xrc_last_error(void) { return g_err.c_str(); }
int xrc_version(void) { return XRC_VERSION; }
uint64_t xrc_launch_count(void) { return g_launch_count.load(); }

// ---------------------------------------------------------------- context
static int ctx_create_common(int device, void* stream, bool external, xrc_ctx** out)
{
  XRC_CHECK_ARG(out, "xrc_ctx_create: null output");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    XRC_FAIL(XRC_ERR_CUDA, std::string("no CUDA device available (there is no CPU fallback): ") +
                               cudaGetErrorString(e));
  XRC_CHECK_ARG(device >= 0 && device < n, "xrc_ctx_create: device ordinal out of range");
  cudaDeviceProp prop;
  XRC_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    XRC_FAIL(XRC_ERR_CUDA, std::string("device ") + prop.name +
                               " is not compute capability 10.x; this library ships sm_100a code only");
  XRC_CUDA(cudaSetDevice(device));
  std::unique_ptr<xrc_ctx> c(new xrc_ctx);
  c->device = device;
  if (external)
  {
    c->stream = (cudaStream_t)stream;
    c->owns_stream = false;
  }
  else
  {
    XRC_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->owns_stream = true;
  }
  *out = c.release();
  return XRC_OK;
}

int xrc_ctx_create(int device, xrc_ctx** out) { return ctx_create_common(device, nullptr, false, out); }
int xrc_ctx_create_on_stream(int device, void* cuda_stream, xrc_ctx** out)
{
  return ctx_create_common(device, cuda_stream, true, out);
}

int xrc_ctx_destroy(xrc_ctx* ctx)
{
  if (!ctx)
    return XRC_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->owns_stream)
    cudaStreamDestroy(ctx->stream);
  delete ctx;
  return XRC_OK;
}

int xrc_ctx_synchronize(xrc_ctx* ctx)
{
  XRC_CHECK_ARG(ctx, "null context");
  XRC_TRY(use_device(ctx));
  XRC_CUDA(cudaStreamSynchronize(ctx->stream));
  return XRC_OK;
}

int xrc_ctx_device(const xrc_ctx* ctx, int* device)
{
  XRC_CHECK_ARG(ctx && device, "null argument");
  *device = ctx->device;
  return XRC_OK;
}

int xrc_ctx_stream(const xrc_ctx* ctx, void** cuda_stream)
{
  XRC_CHECK_ARG(ctx && cuda_stream, "null argument");
  *cuda_stream = (void*)ctx->stream;
  return XRC_OK;
}

// ---------------------------------------------------------------- ray caster
int xrc_rc_create(xrc_ctx* ctx, xrc_rc** out)
{
  XRC_CHECK_ARG(ctx && out, "xrc_rc_create: null argument");
  xrc_rc* rc = new xrc_rc;
  rc->ctx = ctx;
  *out = rc;
  return XRC_OK;
}

static void rc_peer_detach(xrc_rc* rc)
{
  for (uint32_t r = 0; r < rc->peer_n; ++r)
    if (r != rc->peer_rank && rc->peer_bufs[r])
      cudaIpcCloseMemHandle(rc->peer_bufs[r]);
  for (uint32_t r = 0; r < kMaxPeers; ++r)
    rc->peer_bufs[r] = nullptr;
  rc->peer_n = 0;
  rc->tile_begin.clear();
  rc->tile_mult.clear();
  rc->tile_work.clear();
}

static void rc_free_vols(xrc_rc* rc)
{
  for (auto& v : rc->vols)
    free_volume(&v);
  rc->vols.clear();
}

int xrc_rc_destroy(xrc_rc* rc)
{
  if (!rc)
    return XRC_OK;
  if (rc->n_borrowers)
    XRC_FAIL(XRC_ERR_INVALID, "xrc_rc_destroy: other ray casters still use this one's projection buffer "
                              "(xrc_rc_use_other_proj_buf); destroy them first");
  if (rc->other)
    --rc->other->n_borrowers;
  cudaSetDevice(rc->ctx->device);
  cudaStreamSynchronize(rc->ctx->stream);
  rc_peer_detach(rc);
  rc_free_vols(rc);
  dfree(rc->d_cams);
  dfree(rc->d_buf_own);
  dfree(rc->d_poses);
  dfree(rc->d_cam_idx);
  dfree(rc->d_zero_idx);
  dfree(rc->d_bg);
  if (rc->h_gather)
    cudaFreeHost(rc->h_gather);
  if (rc->h_poses)
    cudaFreeHost(rc->h_poses);
  if (rc->h_cam_idx)
    cudaFreeHost(rc->h_cam_idx);
  if (rc->staged)
    cudaEventDestroy(rc->staged);
  delete rc;
  return XRC_OK;
}

int xrc_rc_set_layout(xrc_rc* rc, int layout)
{
  XRC_CHECK_ARG(rc, "null ray caster");
  XRC_CHECK_ARG(layout >= XRC_LAYOUT_DEFAULT && layout <= XRC_LAYOUT_PAX, "unknown layout");
  XRC_CHECK_ARG(rc->vols.empty(), "xrc_rc_set_layout must be called before xrc_rc_set_volumes");
  rc->layout = layout;
  return XRC_OK;
}

static int rc_set_volumes_impl(xrc_rc* rc, uint32_t n, const float* const* ptrs, const uint64_t (*dims)[3],
                               const float (*idx_to_phys)[12], bool on_device, const float* hu_lower = nullptr)
{
  XRC_CHECK_ARG(rc && ptrs && dims && idx_to_phys, "xrc_rc_set_volumes: null argument");
  XRC_CHECK_ARG(n > 0, "xrc_rc_set_volumes: need at least one volume");
  // validate everything before touching the volumes in place: a failure leaves the previous ones usable
  for (uint32_t i = 0; i < n; ++i)
  {
    XRC_CHECK_ARG(ptrs[i], "xrc_rc_set_volumes: null volume pointer");
    for (int k = 0; k < 3; ++k)
      XRC_CHECK_ARG(dims[i][k] >= 1 && dims[i][k] < (1u << 22), "xrc_rc_set_volumes: bad volume dimension");
    XRC_CHECK_ARG(dims[i][0] * dims[i][1] < (1ull << 31), "xrc_rc_set_volumes: slice too large");
  }
  XRC_TRY(use_device(rc->ctx));
  cudaStream_t st = rc->ctx->stream;
  XRC_CUDA(cudaStreamSynchronize(st));
  // the new volumes are built aside and swapped in only when all of them exist
  std::vector<DeviceVolume> fresh(n);
  int status = XRC_OK;
  for (uint32_t i = 0; i < n && status == XRC_OK; ++i)
  {
    DeviceVolume& v = fresh[i];
    for (int k = 0; k < 3; ++k)
      v.dims[k] = dims[i][k];
    // default: principal-axis stacks; volumes whose padded record count exceeds 32 bits use one XY-quad stack
    int layout = rc->layout;
    if (layout == XRC_LAYOUT_DEFAULT)
    {
      const uint64_t dmax = std::max(v.dims[0], std::max(v.dims[1], v.dims[2]));
      layout = ((v.dims[0] + 2) * (v.dims[1] + 2) * (v.dims[2] + 2) + dmax * dmax < (1ull << 32)) ? XRC_LAYOUT_PAX
                                                                                                   : XRC_LAYOUT_QUAD;
    }
    memcpy(v.idx_to_phys, idx_to_phys[i], sizeof(float) * 12);
    affine_inverse_f32(v.idx_to_phys, v.phys_to_idx);
    const size_t nvox = (size_t)v.dims[0] * v.dims[1] * v.dims[2];
    const float* d_src = ptrs[i];
    float* staging = nullptr;
    status = [&]() -> int {
      if (!on_device)
      {
        XRC_CUDA(cudaMalloc(&staging, nvox * sizeof(float)));
        XRC_CUDA(cudaMemcpyAsync(staging, ptrs[i], nvox * sizeof(float), cudaMemcpyHostToDevice, st));
        d_src = staging;
      }
      if (hu_lower)
      {
        // HU -> linear attenuation on the device before repacking (HUToLinAttFilter, lib/image/xregHUToLinAtt.cpp:45-69);
        // a device-resident source must not be modified: convert a copy
        if (!staging)
        {
          XRC_CUDA(cudaMalloc(&staging, nvox * sizeof(float)));
          XRC_CUDA(cudaMemcpyAsync(staging, ptrs[i], nvox * sizeof(float), cudaMemcpyDeviceToDevice, st));
          d_src = staging;
        }
        launch_hu_to_lin_att(staging, nvox, *hu_lower, st);
      }
      int s = repack_volume(d_src, &v, layout, st);
      if (s == XRC_ERR_NOMEM && rc->layout == XRC_LAYOUT_DEFAULT && layout == XRC_LAYOUT_PAX)
      {
        // not even the f32 copy the stacks are built from fits: one XY-quad stack instead (4x instead of up to 12x)
        cudaGetLastError();
        free_volume(&v);
        layout = XRC_LAYOUT_QUAD;
        s = repack_volume(d_src, &v, layout, st);
      }
      if (s == XRC_OK && layout == XRC_LAYOUT_PAX)
        s = build_occupancy(d_src, &v, st);
      return s;
    }();
    cudaStreamSynchronize(st);
    if (staging)
      cudaFree(staging);   // on every exit path
  }
  if (status != XRC_OK)
  {
    for (auto& v : fresh)
      free_volume(&v);
    return status;
  }
  rc_free_vols(rc);
  rc->vols.swap(fresh);
  return XRC_OK;
}

int xrc_rc_set_volumes(xrc_rc* rc, uint32_t n, const float* const* host_ptrs, const uint64_t (*dims)[3],
                       const float (*idx_to_phys)[12])
{
  return rc_set_volumes_impl(rc, n, host_ptrs, dims, idx_to_phys, false);
}

int xrc_rc_set_volumes_hu(xrc_rc* rc, uint32_t n, const float* const* host_ptrs, const uint64_t (*dims)[3],
                          const float (*idx_to_phys)[12], float hu_lower)
{
  return rc_set_volumes_impl(rc, n, host_ptrs, dims, idx_to_phys, false, &hu_lower);
}

int xrc_rc_set_volumes_device(xrc_rc* rc, uint32_t n, const float* const* dev_ptrs, const uint64_t (*dims)[3],
                              const float (*idx_to_phys)[12])
{
  return rc_set_volumes_impl(rc, n, dev_ptrs, dims, idx_to_phys, true);
}

static int rc_wait_staging(xrc_rc* rc);

int xrc_rc_set_cameras(xrc_rc* rc, uint32_t n, const xrc_cam* cams)
{
  XRC_CHECK_ARG(rc && cams && n > 0, "xrc_rc_set_cameras: bad argument");
  for (uint32_t i = 0; i < n; ++i)
  {
    XRC_CHECK_ARG(cams[i].rows > 0 && cams[i].cols > 0, "xrc_rc_set_cameras: empty detector");
    // xregRayCastBaseCPU.cpp:60-70
    XRC_CHECK_ARG(cams[i].rows == cams[0].rows && cams[i].cols == cams[0].cols,
                  "xrc_rc_set_cameras: all camera models must have the same detector size");
    XRC_CHECK_ARG(cams[i].frame_type >= 0 && cams[i].frame_type <= 2, "xrc_rc_set_cameras: bad frame type");
  }
  XRC_CHECK_ARG(!rc->allocated || (cams[0].rows == rc->rows && cams[0].cols == rc->cols),
                "xrc_rc_set_cameras: detector size changed after allocation");
  XRC_TRY(use_device(rc->ctx));
  cudaStream_t st = rc->ctx->stream;
  XRC_CUDA(cudaStreamSynchronize(st));
  if (n < rc->cams.size() && rc->allocated)
  {
    // stored per-projection camera indices may now be out of range: back to camera 0 until the caller sets them again
    // (a caller's device index array, xrc_rc_set_poses_device, is clamped inside the kernels)
    XRC_TRY(rc_wait_staging(rc));
    for (uint32_t p = 0; p < rc->max_projs; ++p)
      if (rc->h_cam_idx[p] >= n)
        rc->h_cam_idx[p] = 0;
    XRC_CUDA(cudaMemcpy(rc->d_cam_idx, rc->h_cam_idx, sizeof(uint32_t) * rc->max_projs, cudaMemcpyHostToDevice));
  }
  rc->cams.assign(cams, cams + n);
  rc->rows = cams[0].rows;
  rc->cols = cams[0].cols;
  dfree(rc->d_cams);
  XRC_CUDA(cudaMalloc(&rc->d_cams, sizeof(xrc_cam) * n));
  XRC_CUDA(cudaMemcpy(rc->d_cams, cams, sizeof(xrc_cam) * n, cudaMemcpyHostToDevice));
  return XRC_OK;
}

int xrc_rc_allocate(xrc_rc* rc, uint32_t max_projs)
{
  XRC_CHECK_ARG(rc, "null ray caster");
  XRC_CHECK_ARG(!rc->vols.empty(), "xrc_rc_allocate: set volumes first");
  XRC_CHECK_ARG(!rc->cams.empty(), "xrc_rc_allocate: set camera models first");
  XRC_CHECK_ARG(max_projs > 0, "xrc_rc_allocate: need at least one projection");
  XRC_TRY(use_device(rc->ctx));
  XRC_CUDA(cudaStreamSynchronize(rc->ctx->stream));
  rc_peer_detach(rc);   // the peers' mappings of the old buffer are the callers' to drop (xrc_rc_peer_detach on every rank)
  dfree(rc->d_buf_own);
  dfree(rc->d_poses);
  dfree(rc->d_cam_idx);
  dfree(rc->d_zero_idx);
  rc->ext_poses = nullptr;
  rc->ext_cam_idx = nullptr;
  if (rc->h_poses)
    cudaFreeHost(rc->h_poses);
  if (rc->h_cam_idx)
    cudaFreeHost(rc->h_cam_idx);
  rc->h_poses = nullptr;
  rc->h_cam_idx = nullptr;
  const size_t npix = (size_t)rc->rows * rc->cols;
  if (!rc->other)
  {
    rc->xchg_off = (npix * max_projs * sizeof(float) + 255u) & ~(size_t)255u;
    const size_t bytes = rc->xchg_off + kXchgBlockBytes + sizeof(float) * max_projs;
    XRC_CUDA(cudaMalloc(&rc->d_buf_own, bytes));
    XRC_CUDA(cudaMemsetAsync(rc->d_buf_own, 0, bytes, rc->ctx->stream));
    rc->xchg_epoch = 0;
    if (rc->h_gather)
      cudaFreeHost(rc->h_gather);
    rc->h_gather = nullptr;
    XRC_CUDA(cudaHostAlloc(&rc->h_gather, sizeof(float) * ((size_t)max_projs + 1), cudaHostAllocMapped));
    memset(rc->h_gather, 0, sizeof(float) * ((size_t)max_projs + 1));
  }
  else
  {
    XRC_CHECK_ARG(rc_proj_capacity(rc) >= max_projs && rc->other->rows == rc->rows && rc->other->cols == rc->cols,
                  "xrc_rc_allocate: shared projection buffer is too small");
  }
  XRC_CUDA(cudaMalloc(&rc->d_poses, sizeof(float) * 12 * max_projs));
  XRC_CUDA(cudaMalloc(&rc->d_cam_idx, sizeof(uint32_t) * max_projs));
  XRC_CUDA(cudaMemsetAsync(rc->d_cam_idx, 0, sizeof(uint32_t) * max_projs, rc->ctx->stream));
  XRC_CUDA(cudaMalloc(&rc->d_zero_idx, sizeof(uint32_t) * max_projs));
  XRC_CUDA(cudaMemsetAsync(rc->d_zero_idx, 0, sizeof(uint32_t) * max_projs, rc->ctx->stream));
  XRC_CUDA(cudaHostAlloc(&rc->h_poses, sizeof(float) * 12 * max_projs, cudaHostAllocDefault));
  XRC_CUDA(cudaHostAlloc(&rc->h_cam_idx, sizeof(uint32_t) * max_projs, cudaHostAllocDefault));
  if (!rc->staged)
    XRC_CUDA(cudaEventCreateWithFlags(&rc->staged, cudaEventDisableTiming));
  // identity poses until the caller sets them
  for (uint32_t p = 0; p < max_projs; ++p)
  {
    float* m = rc->h_poses + 12 * (size_t)p;
    memset(m, 0, sizeof(float) * 12);
    m[0] = m[5] = m[10] = 1.0f;
    rc->h_cam_idx[p] = 0;
  }
  XRC_CUDA(cudaMemcpyAsync(rc->d_poses, rc->h_poses, sizeof(float) * 12 * max_projs, cudaMemcpyHostToDevice,
                           rc->ctx->stream));
  XRC_CUDA(cudaEventRecord(rc->staged, rc->ctx->stream));
  rc->staged_pending = true;
  rc->max_projs = max_projs;
  rc->num_projs = max_projs;
  rc->allocated = true;
  return XRC_OK;
}

static int rc_upload_poses(xrc_rc* rc, uint32_t n);

int xrc_rc_set_num_projs(xrc_rc* rc, uint32_t n)
{
  XRC_CHECK_ARG(rc, "null ray caster");
  XRC_CHECK_ARG(rc->allocated, "xrc_rc_set_num_projs: allocate first (capacity is fixed by xrc_rc_allocate)");
  XRC_CHECK_ARG(n <= rc->max_projs, "xrc_rc_set_num_projs: exceeds allocated capacity");
  if (rc->inline_poses && n != rc->num_projs)
  {
    // keep the device copy coherent with what the caller last set before the count changes
    XRC_TRY(use_device(rc->ctx));
    XRC_TRY(rc_upload_poses(rc, rc->num_projs));
  }
  rc->num_projs = n;
  return XRC_OK;
}

int xrc_rc_num_projs(const xrc_rc* rc, uint32_t* n)
{
  XRC_CHECK_ARG(rc && n, "null argument");
  *n = rc->num_projs;
  return XRC_OK;
}

int xrc_rc_max_projs_possible(const xrc_rc* rc, uint64_t* n)
{
  XRC_CHECK_ARG(rc && n, "null argument");
  XRC_CHECK_ARG(!rc->cams.empty(), "xrc_rc_max_projs_possible: set camera models first");
  XRC_TRY(use_device(rc->ctx));
  size_t free_b = 0, total_b = 0;
  XRC_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const size_t per_proj = (size_t)rc->rows * rc->cols * sizeof(float) * 3 + 64;  // DRR + two gradient images
  *n = (uint64_t)((free_b * 9 / 10) / per_proj);
  return XRC_OK;
}

static int rc_upload_poses(xrc_rc* rc, uint32_t n)
{
  cudaStream_t st = rc->ctx->stream;
  rc->ext_poses = nullptr;
  rc->ext_cam_idx = nullptr;
  rc->inline_poses = false;
  XRC_CUDA(cudaMemcpyAsync(rc->d_poses, rc->h_poses, sizeof(float) * 12 * n, cudaMemcpyHostToDevice, st));
  XRC_CUDA(cudaMemcpyAsync(rc->d_cam_idx, rc->h_cam_idx, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, st));
  XRC_CUDA(cudaEventRecord(rc->staged, st));
  rc->staged_pending = true;
  return XRC_OK;
}

static int rc_wait_staging(xrc_rc* rc)
{
  if (rc->staged_pending)
  {
    XRC_CUDA(cudaEventSynchronize(rc->staged));
    rc->staged_pending = false;
  }
  return XRC_OK;
}

int xrc_rc_set_poses(xrc_rc* rc, uint32_t n, const float* cam_to_phys, const uint32_t* cam_idx)
{
  XRC_CHECK_ARG(rc && cam_to_phys, "xrc_rc_set_poses: null argument");
  XRC_CHECK_ARG(rc->allocated, "xrc_rc_set_poses: allocate first");
  XRC_CHECK_ARG(n == rc->num_projs, "xrc_rc_set_poses: pose count must equal num_projs");
  if (cam_idx)
    for (uint32_t i = 0; i < n; ++i)
      XRC_CHECK_ARG(cam_idx[i] < rc->cams.size(), "xrc_rc_set_poses: camera index out of range");
  XRC_TRY(use_device(rc->ctx));
  XRC_TRY(rc_wait_staging(rc));
  memcpy(rc->h_poses, cam_to_phys, sizeof(float) * 12 * n);
  if (cam_idx)
    memcpy(rc->h_cam_idx, cam_idx, sizeof(uint32_t) * n);
  else
    memset(rc->h_cam_idx, 0, sizeof(uint32_t) * n);
  return rc_upload_poses(rc, n);
}

int xrc_rc_distribute_poses(xrc_rc* rc, uint32_t n_poses, const float* cam_to_phys)
{
  XRC_CHECK_ARG(rc && cam_to_phys, "xrc_rc_distribute_poses: null argument");
  XRC_CHECK_ARG(rc->allocated, "xrc_rc_distribute_poses: allocate first");
  const uint32_t n_cams = (uint32_t)rc->cams.size();
  // xregRayCastInterface.cpp:103
  XRC_CHECK_ARG((uint64_t)n_poses * n_cams == rc->num_projs,
                "xrc_rc_distribute_poses: n_poses * n_cams must equal num_projs");
  XRC_TRY(use_device(rc->ctx));
  XRC_TRY(rc_wait_staging(rc));
  uint32_t g = 0;
  for (uint32_t c = 0; c < n_cams; ++c)
  {
    for (uint32_t p = 0; p < n_poses; ++p, ++g)
    {
      memcpy(rc->h_poses + 12 * (size_t)g, cam_to_phys + 12 * (size_t)p, sizeof(float) * 12);
      rc->h_cam_idx[g] = c;
    }
  }
  return rc_upload_poses(rc, g);
}

int xrc_rc_set_poses_device(xrc_rc* rc, uint32_t n, const float* dev_cam_to_phys, const uint32_t* dev_cam_idx)
{
  XRC_CHECK_ARG(rc && dev_cam_to_phys, "xrc_rc_set_poses_device: null argument");
  XRC_CHECK_ARG(rc->allocated, "xrc_rc_set_poses_device: allocate first");
  XRC_CHECK_ARG(n == rc->num_projs, "xrc_rc_set_poses_device: pose count must equal num_projs");
  rc->ext_poses = dev_cam_to_phys;
  rc->ext_cam_idx = dev_cam_idx ? dev_cam_idx : rc->d_zero_idx;
  rc->ext_mirrored = false;
  rc->inline_poses = false;
  return XRC_OK;
}

int xrc_rc_set_poses_device_mirrored(xrc_rc* rc, uint32_t n, const float* dev_cam_to_phys, const uint32_t* dev_cam_idx,
                                     const float* host_cam_to_phys, const uint32_t* host_cam_idx)
{
  XRC_CHECK_ARG(host_cam_to_phys, "xrc_rc_set_poses_device_mirrored: null host mirror");
  XRC_CHECK_ARG((dev_cam_idx == nullptr) == (host_cam_idx == nullptr),
                "xrc_rc_set_poses_device_mirrored: camera indices must be given on both sides or on neither");
  XRC_TRY(xrc_rc_set_poses_device(rc, n, dev_cam_to_phys, dev_cam_idx));
  if (host_cam_idx)
    for (uint32_t i = 0; i < n; ++i)
      XRC_CHECK_ARG(host_cam_idx[i] < rc->cams.size(), "xrc_rc_set_poses_device_mirrored: camera index out of range");
  XRC_TRY(use_device(rc->ctx));
  XRC_TRY(rc_wait_staging(rc));
  memcpy(rc->h_poses, host_cam_to_phys, sizeof(float) * 12 * n);
  if (host_cam_idx)
    memcpy(rc->h_cam_idx, host_cam_idx, sizeof(uint32_t) * n);
  else
    memset(rc->h_cam_idx, 0, sizeof(uint32_t) * n);
  rc->ext_mirrored = true;
  return XRC_OK;
}

int xrc_rc_set_params(xrc_rc* rc, float step_size, int interp, int kernel_id, int store_method, float default_bg)
{
  XRC_CHECK_ARG(rc, "null ray caster");
  XRC_CHECK_ARG(step_size > 0.0f, "xrc_rc_set_params: step size must be positive");
  XRC_CHECK_ARG(interp >= XRC_INTERP_LINEAR && interp <= XRC_INTERP_BSPLINE, "xrc_rc_set_params: bad interpolation id");
  if (interp != XRC_INTERP_LINEAR && interp != XRC_INTERP_NN)
    XRC_FAIL(XRC_ERR_UNSUPPORTED, "sinc / B-spline interpolation is not supported on the GPU (the reference's OpenCL ray caster "
                                  "supports linear only, xregRayCastBaseOCL.cpp:338-341; here: linear and nearest neighbour)");
  XRC_CHECK_ARG(kernel_id == XRC_KERNEL_SUM || kernel_id == XRC_KERNEL_MAX, "xrc_rc_set_params: unsupported line integral kernel");
  XRC_CHECK_ARG(store_method == XRC_STORE_REPLACE || store_method == XRC_STORE_ACCUM, "xrc_rc_set_params: bad store method");
  rc->step_size = step_size;
  rc->interp = interp;
  rc->kernel_id = kernel_id;
  rc->store_method = store_method;
  rc->default_bg = default_bg;
  return XRC_OK;
}

int xrc_rc_set_bg_projs(xrc_rc* rc, const float* const* host_imgs, int use_bg)
{
  XRC_CHECK_ARG(rc, "null ray caster");
  if (!use_bg)
  {
    rc->use_bg = false;
    return XRC_OK;
  }
  XRC_CHECK_ARG(!rc->cams.empty(), "xrc_rc_set_bg_projs: set camera models first");
  XRC_CHECK_ARG(host_imgs || rc->d_bg, "xrc_rc_set_bg_projs: no background images given");
  XRC_TRY(use_device(rc->ctx));
  const size_t npix = (size_t)rc->rows * rc->cols;
  if (host_imgs)
  {
    XRC_CUDA(cudaStreamSynchronize(rc->ctx->stream));
    dfree(rc->d_bg);
    XRC_CUDA(cudaMalloc(&rc->d_bg, npix * rc->cams.size() * sizeof(float)));
    for (size_t c = 0; c < rc->cams.size(); ++c)
    {
      XRC_CHECK_ARG(host_imgs[c], "xrc_rc_set_bg_projs: need one background image per camera");
      XRC_CUDA(cudaMemcpy(rc->d_bg + c * npix, host_imgs[c], npix * sizeof(float), cudaMemcpyHostToDevice));
    }
  }
  rc->use_bg = true;
  return XRC_OK;
}

// XRC_LAYOUT_PAX stacks are built on demand.  The kernel decides per CTA which stack it marches (the principal axis of
// the ray through its tile centre, pax_cta_prologue), and the low bits of a sample depend on that choice (the plane of
// the bilinear step differs between stacks), so for results to be a pure function of (volume, camera, pose) every
// stack a CTA can choose must exist before the launch.  Host-visible poses: the index-space ray direction is affine in
// (col, row), so axis i can never win anywhere on the detector if some axis k with constant sign s satisfies
// s d_k >= (1 + eps) |d_i| at the four detector corners; every axis not excluded that way is built (a superset of the
// kernel's choices; eps covers the different rounding of kernel and host).  Poses only the device knows
// (xrc_rc_set_poses_device without a host mirror) get all three stacks.  The kernel still has a fallback for a missing
// stack (any built one; it reports the wish through host-mapped memory and the next compute() honours it), which only
// an out-of-memory condition or a wrong mirror can trigger.
static int rc_prepare_stacks(xrc_rc* rc, uint32_t vol_idx)
{
  DeviceVolume& v = rc->vols[vol_idx];
  if (v.layout != XRC_LAYOUT_PAX || (v.pax[0] && v.pax[1] && v.pax[2]))
    return XRC_OK;
  bool need[3] = {false, false, false};
  if (v.h_want)
    for (int k = 0; k < 3; ++k)
      if (*(volatile uint32_t*)(v.h_want + k))
      {
        need[k] = true;
        v.h_want[k] = 0u;
      }
  if (rc->ext_poses && !rc->ext_mirrored)
    need[0] = need[1] = need[2] = true;
  else
  {
    const float* A = v.phys_to_idx;
    const float eps = 1.0e-3f;
    for (uint32_t p = 0; p < rc->num_projs && !(need[0] && need[1] && need[2]); ++p)
    {
      const float* P = rc->h_poses + 12 * (size_t)p;
      const xrc_cam& cam = rc->cams[std::min<size_t>(rc->h_cam_idx[p], rc->cams.size() - 1)];
      float X[12];
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 4; ++c)
          X[4 * r + c] = A[4 * r] * P[c] + A[4 * r + 1] * P[4 + c] + A[4 * r + 2] * P[8 + c] + ((c == 3) ? A[4 * r + 3] : 0.0f);
      float src[3];
      for (int r = 0; r < 3; ++r)
        src[r] = X[4 * r] * cam.pinhole[0] + X[4 * r + 1] * cam.pinhole[1] + X[4 * r + 2] * cam.pinhole[2] + X[4 * r + 3];
      const float det_z = ((cam.frame_type == 1) ? -1.0f : 1.0f) * cam.focal_len;
      float d[4][3];
      for (int g = 0; g < 4; ++g)
      {
        const float col = (g & 1) ? (float)(cam.cols - 1) : 0.0f, row = (g & 2) ? (float)(cam.rows - 1) : 0.0f;
        float cv[3], w[3];
        for (int r = 0; r < 3; ++r)
          cv[r] = det_z * (cam.intrins_inv[3 * r] * col + cam.intrins_inv[3 * r + 1] * row + cam.intrins_inv[3 * r + 2]);
        if (cam.frame_type == 2)
          cv[2] -= cam.focal_len;
        for (int r = 0; r < 3; ++r)
          w[r] = cam.extrins_inv[4 * r] * cv[0] + cam.extrins_inv[4 * r + 1] * cv[1] + cam.extrins_inv[4 * r + 2] * cv[2] +
                 cam.extrins_inv[4 * r + 3];
        for (int r = 0; r < 3; ++r)
          d[g][r] = X[4 * r] * w[0] + X[4 * r + 1] * w[1] + X[4 * r + 2] * w[2] + X[4 * r + 3] - src[r];
      }
      for (int i = 0; i < 3; ++i)
      {
        bool dominated = false;
        for (int k = 0; k < 3 && !dominated; ++k)
        {
          if (k == i)
            continue;
          for (int sgn = -1; sgn <= 1 && !dominated; sgn += 2)
          {
            bool all = true;
            for (int g = 0; g < 4; ++g)
              all = all && ((float)sgn * d[g][k] >= (1.0f + eps) * fabsf(d[g][i])) && ((float)sgn * d[g][k] > 0.0f);
            dominated = all;
          }
        }
        if (!dominated)
          need[i] = true;
      }
    }
  }
  for (int k = 0; k < 3; ++k)
  {
    if (!need[k] || v.pax[k])
      continue;
    const int s = build_pax_stack(&v, k, rc->ctx->stream);
    if (s == XRC_ERR_NOMEM && (v.pax[0] || v.pax[1] || v.pax[2]))
    {
      cudaGetLastError();   // no room for another stack: the kernel keeps falling back to one that exists
      continue;
    }
    if (s == XRC_ERR_NOMEM && v.src && rc->layout == XRC_LAYOUT_DEFAULT)
    {
      // not even one padded stack fits: the unpadded XY-quad stack of the measurement layouts instead (same
      // samples to 1 ulp of the lerp chain, slower for views along x; reported by xrc_rc_volume_layout)
      cudaGetLastError();
      float* lin = v.src;
      v.src = nullptr;
      if (v.occ)
        cudaFree(v.occ);
      v.occ = nullptr;
      const int s2 = repack_volume(lin, &v, XRC_LAYOUT_QUAD, rc->ctx->stream);
      cudaStreamSynchronize(rc->ctx->stream);
      cudaFree(lin);
      XRC_TRY(s2);
      return XRC_OK;
    }
    XRC_TRY(s);
  }
  if (!(v.pax[0] || v.pax[1] || v.pax[2]))
    XRC_TRY(build_pax_stack(&v, 2, rc->ctx->stream));   // nothing to project yet: any stack will do
  return XRC_OK;
}

static void rc_fill_args(xrc_rc* rc, uint32_t vol_idx, DrrArgs* a)
{
  const DeviceVolume& v = rc->vols[vol_idx];
  memset(a, 0, sizeof(*a));
  a->vol = v.data;
  a->tex = v.tex;
  a->nx = (int)v.dims[0];
  a->ny = (int)v.dims[1];
  a->nz = (int)v.dims[2];
  memcpy(a->phys_to_idx, v.phys_to_idx, sizeof(float) * 12);
  a->cams = rc->d_cams;
  a->n_cams = (uint32_t)rc->cams.size();
  a->poses = rc->ext_poses ? rc->ext_poses : rc->d_poses;
  a->cam_idx = rc->ext_poses ? rc->ext_cam_idx : rc->d_cam_idx;
  a->n_projs = rc->num_projs;
  a->rows = rc->rows;
  a->cols = rc->cols;
  a->step_size = rc->step_size;
  a->out = rc_proj_buf(rc);
  a->init_mode = rc->use_bg ? 1 : ((rc->store_method == XRC_STORE_REPLACE) ? 0 : 2);
  a->default_bg = rc->default_bg;
  a->bg = rc->d_bg;
  a->order = rc->order & 1;
  a->variant = rc->order >> 1;
  for (int k = 0; k < 3; ++k)
  {
    a->pax[k] = v.pax[k];
    a->pax_sb[k] = v.pax_sb[k];
    a->pax_sc[k] = v.pax_sc[k];
  }
  a->pax_want = v.h_want;
  if (rc->inline_poses && !rc->ext_poses && rc->num_projs <= kInlinePoses)
  {
    a->use_inline = 1;
    memcpy(a->inl_poses, rc->h_poses, sizeof(float) * 12 * rc->num_projs);
    memcpy(a->inl_cam, rc->h_cam_idx, sizeof(uint32_t) * rc->num_projs);
  }
  a->occ = rc->skip_empty ? v.occ : nullptr;
  // interior gaps: on request only (xrc_rc_set_skip_empty(rc, 3)).  v.occ_fill tells how sparse the volume is, but sparse is
  // not enough: the C2 phantom reduced to its 12 bones (fill 0.04) has few gaps ALONG its rays -- 7 % fewer samples
  // fetched, 5 % slower for the segment-wise marching -- while two organs far apart along the view direction gain
  a->gaps = (rc->skip_empty && v.occ && rc->gap_mode == 1) ? 1 : 0;
  a->occ_wx = v.occ_wx;
  a->occ_ny = v.occ_ny;
  for (int k = 0; k < 3; ++k)
  {
    a->occ_lo[k] = v.occ_lo[k];
    a->occ_hi[k] = v.occ_hi[k];
  }
}

// where voxel (ix, iy, iz) lives as a plain float in whatever payload the volume has (nearest-neighbour interpolation)
static int rc_fill_nn(const DeviceVolume& v, DrrArgs* a)
{
  const uint32_t nx = (uint32_t)v.dims[0], ny = (uint32_t)v.dims[1];
  a->nn_off = 0;
  switch (v.layout)
  {
    case XRC_LAYOUT_LINEAR:   // padded (nx + 1) x (ny + 1) x (nz + 1)
      a->nn_base = (const float*)v.data;
      a->nn_s[0] = 1; a->nn_s[1] = nx + 1; a->nn_s[2] = (nx + 1) * (ny + 1);
      return XRC_OK;
    case XRC_LAYOUT_QUAD:     // float4 {v(x, y), ...} per voxel
      a->nn_base = (const float*)v.data;
      a->nn_s[0] = 4; a->nn_s[1] = 4 * nx; a->nn_s[2] = 4 * nx * ny;
      return XRC_OK;
    case XRC_LAYOUT_OCT:      // two float4 per voxel, v(x, y, z) first
      a->nn_base = (const float*)v.data;
      a->nn_s[0] = 8; a->nn_s[1] = 8 * nx; a->nn_s[2] = 8 * nx * ny;
      return XRC_OK;
    case XRC_LAYOUT_PAX:
      for (int k = 0; k < 3; ++k)
        if (v.pax[k])
        {
          // stack k: rec = (ic + 1) Sc + (ib + 1) Sb + (ia + 1), c = axis k, a = (k + 1) % 3, b = (k + 2) % 3; the record's
          // first component is v(ia, ib, ic) itself (the difference form keeps it)
          const int ka = (k + 1) % 3, kb = (k + 2) % 3;
          XRC_CHECK_ARG((uint64_t)4 * v.pax_sc[k] * (v.dims[k] + 2) < (1ull << 32), "nearest-neighbour interpolation: volume too large");
          a->nn_base = (const float*)v.pax[k];
          a->nn_s[ka] = 4; a->nn_s[kb] = 4 * v.pax_sb[k]; a->nn_s[k] = 4 * v.pax_sc[k];
          a->nn_off = 4 * (v.pax_sc[k] + v.pax_sb[k] + 1);
          return XRC_OK;
        }
      XRC_FAIL(XRC_ERR_INVALID, "nearest-neighbour interpolation: no volume stack built");
    default:
      XRC_FAIL(XRC_ERR_UNSUPPORTED, "nearest-neighbour interpolation is not available for the texture layouts");
  }
}

int xrc_rc_compute(xrc_rc* rc, uint32_t vol_idx)
{
  XRC_CHECK_ARG(rc, "null ray caster");
  XRC_CHECK_ARG(rc->allocated, "xrc_rc_compute: resources not allocated (xregRayCastLineIntCPU.cpp:296)");
  XRC_CHECK_ARG(vol_idx < rc->vols.size(), "xrc_rc_compute: volume index out of range");
  XRC_CHECK_ARG(rc_proj_buf(rc) && rc->num_projs <= rc_proj_capacity(rc),
                "xrc_rc_compute: the shared projection buffer (xrc_rc_use_other_proj_buf) is gone or too small");
  XRC_TRY(use_device(rc->ctx));
  XRC_TRY(rc_prepare_stacks(rc, vol_idx));
  DrrArgs a;
  rc_fill_args(rc, vol_idx, &a);
  if (rc->interp == XRC_INTERP_NN)
  {
    XRC_TRY(rc_fill_nn(rc->vols[vol_idx], &a));
    return launch_drr(a, kLayoutNN, rc->kernel_id, rc->ctx->stream);
  }
  return launch_drr(a, rc->vols[vol_idx].layout, rc->kernel_id, rc->ctx->stream);
}

// RayCasterDepthCPU::compute (lib/ray_cast/xregRayCastDepthCPU.cpp:236-272) on this ray caster's volumes / cameras / poses
int xrc_rc_compute_depth(xrc_rc* rc, uint32_t vol_idx, float collision_thresh, uint32_t num_backtracking_steps)
{
  XRC_CHECK_ARG(rc, "null ray caster");
  XRC_CHECK_ARG(rc->allocated, "xrc_rc_compute_depth: resources not allocated (xregRayCastDepthCPU.cpp:238)");
  XRC_CHECK_ARG(vol_idx < rc->vols.size(), "xrc_rc_compute_depth: volume index out of range");
  XRC_CHECK_ARG(rc_proj_buf(rc) && rc->num_projs <= rc_proj_capacity(rc),
                "xrc_rc_compute_depth: the shared projection buffer (xrc_rc_use_other_proj_buf) is gone or too small");
  XRC_CHECK_ARG(num_backtracking_steps <= 64, "xrc_rc_compute_depth: more than 64 refinement steps halve the step to nothing");
  if (rc->interp != XRC_INTERP_LINEAR && rc->interp != XRC_INTERP_NN)
    XRC_FAIL(XRC_ERR_UNSUPPORTED, "xrc_rc_compute_depth: linear and nearest-neighbour interpolation only");
  XRC_TRY(use_device(rc->ctx));
  XRC_TRY(rc_prepare_stacks(rc, vol_idx));
  DrrArgs a;
  rc_fill_args(rc, vol_idx, &a);
  XRC_TRY(rc_fill_nn(rc->vols[vol_idx], &a));
  a.depth_thresh = collision_thresh;
  a.depth_backtrack = num_backtracking_steps;
  return launch_depth(a, rc->interp == XRC_INTERP_NN ? 1 : 0, rc->ctx->stream);
}

// ---- multi-GPU tile sharding, one process per GPU
int xrc_rc_peer_export(xrc_rc* rc, uint8_t handle[XRC_IPC_HANDLE_BYTES])
{
  XRC_CHECK_ARG(rc && handle, "null argument");
  XRC_CHECK_ARG(rc->allocated && rc->d_buf_own, "xrc_rc_peer_export: allocate first (a ray caster that borrows its buffer cannot export it)");
  static_assert(sizeof(cudaIpcMemHandle_t) == XRC_IPC_HANDLE_BYTES, "IPC handle size");
  XRC_TRY(use_device(rc->ctx));
  XRC_CUDA(cudaStreamSynchronize(rc->ctx->stream));   // the buffer (and its exchange block) is zeroed before anyone can map it
  cudaIpcMemHandle_t h;
  XRC_CUDA(cudaIpcGetMemHandle(&h, rc->d_buf_own));
  memcpy(handle, &h, sizeof(h));
  return XRC_OK;
}

int xrc_rc_peer_attach(xrc_rc* rc, uint32_t n_ranks, uint32_t rank, const uint8_t* handles)
{
  XRC_CHECK_ARG(rc && handles, "null argument");
  XRC_CHECK_ARG(n_ranks >= 1 && n_ranks <= kMaxPeers && rank < n_ranks, "xrc_rc_peer_attach: 1 <= ranks <= 8, rank < ranks");
  XRC_CHECK_ARG(rc->allocated && rc->d_buf_own, "xrc_rc_peer_attach: allocate first");
  XRC_TRY(use_device(rc->ctx));
  XRC_CUDA(cudaStreamSynchronize(rc->ctx->stream));
  rc_peer_detach(rc);
  rc->peer_rank = rank;
  for (uint32_t r = 0; r < n_ranks; ++r)
  {
    if (r == rank)
    {
      rc->peer_bufs[r] = rc->d_buf_own;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * XRC_IPC_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
    {
      rc->peer_n = r;   // close what was opened so far
      rc_peer_detach(rc);
      cudaGetLastError();
      XRC_FAIL(XRC_ERR_CUDA, std::string("xrc_rc_peer_attach: cudaIpcOpenMemHandle failed for rank ") + std::to_string(r) +
                                 ": " + cudaGetErrorString(e));
    }
    rc->peer_bufs[r] = (float*)p;
  }
  rc->peer_n = n_ranks;
  return XRC_OK;
}

int xrc_rc_peer_detach(xrc_rc* rc)
{
  XRC_CHECK_ARG(rc, "null ray caster");
  XRC_TRY(use_device(rc->ctx));
  XRC_CUDA(cudaStreamSynchronize(rc->ctx->stream));
  rc_peer_detach(rc);
  return XRC_OK;
}

// Which tiles does each rank ray cast?  Contiguous ranges of the row-major tile list, so that the tiles in flight on a
// GPU stay neighbours (their beams -- and, with the pose jitter of a population, the beams of the neighbouring tiles'
// other poses -- overlap in L2: dealing the tiles round robin measured no faster than sharding the poses), cut where the
// MEASURED work balances: one count-only pass of the current projections gives the samples each tile fetches (exact
// integers, identical on every rank, so all ranks derive the same plan without talking), plus a constant per ray for
// its set-up.  Planned once per attach (first xrc_rc_compute_tiles) or on request; an optimiser's later populations
// stay around the same pose, so the plan stays balanced.
static int rc_count_tile_work(xrc_rc* rc, uint32_t vol_idx)
{
  XRC_TRY(use_device(rc->ctx));
  XRC_TRY(rc_prepare_stacks(rc, vol_idx));
  XRC_CHECK_ARG(rc->vols[vol_idx].layout == XRC_LAYOUT_PAX, "xrc_rc_plan_tiles: needs the default volume layout");
  const uint32_t nt = ((rc->cols + 15) / 16) * ((rc->rows + 15) / 16);
  cudaStream_t st = rc->ctx->stream;
  unsigned long long* d_cnt = nullptr;
  XRC_CUDA(cudaMalloc(&d_cnt, (nt + 1) * sizeof(unsigned long long)));
  cudaMemsetAsync(d_cnt, 0, (nt + 1) * sizeof(unsigned long long), st);
  DrrArgs a;
  rc_fill_args(rc, vol_idx, &a);
  a.sample_counter = d_cnt + nt;
  a.tile_counter = d_cnt;
  a.count_only = 1;
  int status = launch_drr(a, XRC_LAYOUT_PAX, rc->kernel_id, st);
  std::vector<unsigned long long> cnt(nt + 1, 0ull);
  if (status == XRC_OK)
  {
    cudaMemcpyAsync(cnt.data(), d_cnt, (nt + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess)
    {
      set_error("xrc_rc_plan_tiles: device failure");
      status = XRC_ERR_CUDA;
    }
  }
  cudaFree(d_cnt);
  XRC_TRY(status);
  // work of a tile: its fetched samples + 40 sample-equivalents per ray for set-up, trimming and the store
  rc->tile_work.resize(nt);
  for (uint32_t t = 0; t < nt; ++t)
    rc->tile_work[t] = (double)cnt[t] + 40.0 * 256.0 * (double)rc->num_projs;
  if (rc->tile_mult.size() != nt)
    rc->tile_mult.assign(nt, 1.0);
  return XRC_OK;
}

// contiguous ranges of equal (work x multiplier)
static void rc_cut_tiles(xrc_rc* rc)
{
  const uint32_t nt = (uint32_t)rc->tile_work.size();
  double total = 0.0;
  for (uint32_t t = 0; t < nt; ++t)
    total += rc->tile_work[t] * rc->tile_mult[t];
  rc->tile_begin.assign(rc->peer_n + 1, nt);
  rc->tile_begin[0] = 0;
  double acc = 0.0;
  uint32_t r = 1;
  for (uint32_t t = 0; t < nt && r < rc->peer_n; ++t)
  {
    acc += rc->tile_work[t] * rc->tile_mult[t];
    while (r < rc->peer_n && acc >= total * (double)r / (double)rc->peer_n)
      rc->tile_begin[r++] = t + 1;
  }
}

int xrc_rc_plan_tiles(xrc_rc* rc, uint32_t vol_idx)
{
  XRC_CHECK_ARG(rc, "null ray caster");
  XRC_CHECK_ARG(rc->allocated && rc->peer_n >= 1, "xrc_rc_plan_tiles: allocate and attach first");
  XRC_CHECK_ARG(vol_idx < rc->vols.size(), "xrc_rc_plan_tiles: volume index out of range");
  XRC_TRY(rc_count_tile_work(rc, vol_idx));
  rc_cut_tiles(rc);
  return XRC_OK;
}

// Feedback from the clock: samples are only a proxy of a tile's cost (beams near the detector's edge coalesce worse,
// long central beams hit L2 more often), so a plan balanced by samples leaves the ranks' kernels a few percent apart.
// rank_ms[r] = the time rank r's ray-casting kernel took under the CURRENT plan (every rank passes the same n_ranks
// numbers, e.g. all-gathered CUDA-event times, so that every rank derives the same new plan).  Each rank's tiles get
// their cost multiplier scaled by (measured time / planned share) ^ 0.7, and the ranges are cut again on the sample
// counts of the current projections.  Two or three rounds settle within ~1 %.
int xrc_rc_plan_tiles_timed(xrc_rc* rc, uint32_t vol_idx, const float* rank_ms)
{
  XRC_CHECK_ARG(rc && rank_ms, "null argument");
  XRC_CHECK_ARG(rc->allocated && rc->peer_n >= 1, "xrc_rc_plan_tiles_timed: allocate and attach first");
  XRC_CHECK_ARG(vol_idx < rc->vols.size(), "xrc_rc_plan_tiles_timed: volume index out of range");
  XRC_CHECK_ARG(rc->tile_begin.size() == (size_t)rc->peer_n + 1 && !rc->tile_work.empty(),
                "xrc_rc_plan_tiles_timed: no plan yet (the times must belong to a plan)");
  double sum_ms = 0.0, sum_w = 0.0;
  std::vector<double> share(rc->peer_n, 0.0);
  for (uint32_t r = 0; r < rc->peer_n; ++r)
  {
    XRC_CHECK_ARG(rank_ms[r] > 0.0f && rank_ms[r] < 1.0e9f, "xrc_rc_plan_tiles_timed: times must be positive and finite");
    for (uint32_t t = rc->tile_begin[r]; t < rc->tile_begin[r + 1]; ++t)
      share[r] += rc->tile_work[t] * rc->tile_mult[t];
    sum_ms += rank_ms[r];
    sum_w += share[r];
  }
  for (uint32_t r = 0; r < rc->peer_n; ++r)
  {
    if (!(share[r] > 0.0))
      continue;
    const double f = pow(((double)rank_ms[r] / sum_ms) / (share[r] / sum_w), 0.7);
    for (uint32_t t = rc->tile_begin[r]; t < rc->tile_begin[r + 1]; ++t)
      rc->tile_mult[t] *= f;
  }
  XRC_TRY(rc_count_tile_work(rc, vol_idx));
  rc_cut_tiles(rc);
  return XRC_OK;
}

int xrc_rc_tile_plan(const xrc_rc* rc, uint32_t* tile_begin /* n_ranks + 1 */)
{
  XRC_CHECK_ARG(rc && tile_begin, "null argument");
  XRC_CHECK_ARG(rc->tile_begin.size() == (size_t)rc->peer_n + 1, "xrc_rc_tile_plan: no plan yet (xrc_rc_plan_tiles / first xrc_rc_compute_tiles)");
  for (size_t i = 0; i < rc->tile_begin.size(); ++i)
    tile_begin[i] = rc->tile_begin[i];
  return XRC_OK;
}

// this rank's tiles of ALL current projections, each stored in its owner's buffer (its global index there)
static void rc_fill_tile_args(xrc_rc* rc, DrrArgs* a)
{
  a->tile_first = rc->tile_begin[rc->peer_rank];
  a->tile_stride = 1;
  a->tile_count = rc->tile_begin[rc->peer_rank + 1] - rc->tile_begin[rc->peer_rank];
  a->peer_n = rc->peer_n;
  a->peer_base = rc->num_projs / rc->peer_n;
  a->peer_extra = rc->num_projs % rc->peer_n;
  for (uint32_t r = 0; r < kMaxPeers; ++r)
    a->peer_out[r] = rc->peer_bufs[r];
}

int xrc_rc_compute_tiles(xrc_rc* rc, uint32_t vol_idx)
{
  XRC_CHECK_ARG(rc, "null ray caster");
  XRC_CHECK_ARG(rc->allocated, "xrc_rc_compute_tiles: resources not allocated");
  XRC_CHECK_ARG(rc->peer_n >= 1, "xrc_rc_compute_tiles: attach the ranks' projection buffers first (xrc_rc_peer_attach)");
  XRC_CHECK_ARG(vol_idx < rc->vols.size(), "xrc_rc_compute_tiles: volume index out of range");
  if (rc->interp != XRC_INTERP_LINEAR)
    XRC_FAIL(XRC_ERR_UNSUPPORTED, "xrc_rc_compute_tiles: the tile-sharded path is linear interpolation only");
  XRC_TRY(use_device(rc->ctx));
  XRC_TRY(rc_prepare_stacks(rc, vol_idx));
  XRC_CHECK_ARG(rc->vols[vol_idx].layout == XRC_LAYOUT_PAX, "xrc_rc_compute_tiles: needs the default volume layout");
  if (rc->tile_begin.size() != (size_t)rc->peer_n + 1)
    XRC_TRY(xrc_rc_plan_tiles(rc, vol_idx));
  DrrArgs a;
  rc_fill_args(rc, vol_idx, &a);
  rc_fill_tile_args(rc, &a);
  return launch_drr(a, XRC_LAYOUT_PAX, rc->kernel_id, rc->ctx->stream);
}

int xrc_rc_tile_samples(xrc_rc* rc, uint32_t vol_idx, uint64_t* algorithmic, uint64_t* fetched)
{
  XRC_CHECK_ARG(rc && (algorithmic || fetched), "null argument");
  XRC_CHECK_ARG(rc->allocated && rc->peer_n >= 1, "xrc_rc_tile_samples: allocate and attach first");
  XRC_CHECK_ARG(vol_idx < rc->vols.size() && rc->vols[vol_idx].layout == XRC_LAYOUT_PAX, "xrc_rc_tile_samples: bad volume");
  XRC_TRY(use_device(rc->ctx));
  XRC_TRY(rc_prepare_stacks(rc, vol_idx));
  if (rc->tile_begin.size() != (size_t)rc->peer_n + 1)
    XRC_TRY(xrc_rc_plan_tiles(rc, vol_idx));
  cudaStream_t st = rc->ctx->stream;
  unsigned long long* d_cnt = nullptr;
  XRC_CUDA(cudaMalloc(&d_cnt, 2 * sizeof(unsigned long long)));
  cudaMemsetAsync(d_cnt, 0, 2 * sizeof(unsigned long long), st);
  DrrArgs a;
  rc_fill_args(rc, vol_idx, &a);
  rc_fill_tile_args(rc, &a);
  a.peer_n = 0;   // counting only: nothing is stored
  a.sample_counter = d_cnt;
  int status = launch_ray_info(a, st);
  if (status == XRC_OK)
  {
    a.sample_counter = d_cnt + 1;
    a.count_only = 1;
    status = launch_drr(a, XRC_LAYOUT_PAX, rc->kernel_id, st);
  }
  unsigned long long cnt[2] = {0, 0};
  if (status == XRC_OK)
  {
    cudaMemcpyAsync(cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess)
    {
      set_error("xrc_rc_tile_samples: device failure");
      status = XRC_ERR_CUDA;
    }
  }
  cudaFree(d_cnt);
  if (algorithmic)
    *algorithmic = cnt[0];
  if (fetched)
    *fetched = cnt[1];
  return status;
}

int xrc_rc_volume_layout(const xrc_rc* rc, uint32_t vol_idx, int* layout)
{
  XRC_CHECK_ARG(rc && layout, "null argument");
  XRC_CHECK_ARG(vol_idx < rc->vols.size(), "xrc_rc_volume_layout: volume index out of range");
  *layout = rc->vols[vol_idx].layout;
  return XRC_OK;
}

int xrc_rc_volume_bytes(const xrc_rc* rc, uint64_t* bytes)
{
  XRC_CHECK_ARG(rc && bytes, "null argument");
  uint64_t b = 0;
  for (const auto& v : rc->vols)
    b += v.bytes + (v.occ ? sizeof(uint32_t) * (uint64_t)v.occ_wx * v.occ_ny * v.occ_nz : 0);
  *bytes = b;
  return XRC_OK;
}

int xrc_rc_device_buf(xrc_rc* rc, float** dev_ptr)
{
  XRC_CHECK_ARG(rc && dev_ptr, "null argument");
  XRC_CHECK_ARG(rc->allocated, "xrc_rc_device_buf: allocate first");
  *dev_ptr = rc_proj_buf(rc);
  return XRC_OK;
}

int xrc_rc_read_projs(xrc_rc* rc, uint32_t first, uint32_t count, float* host_dst)
{
  XRC_CHECK_ARG(rc && host_dst, "null argument");
  XRC_CHECK_ARG(rc->allocated, "xrc_rc_read_projs: allocate first");
  XRC_CHECK_ARG((uint64_t)first + count <= rc->max_projs, "xrc_rc_read_projs: range exceeds capacity");
  XRC_TRY(use_device(rc->ctx));
  const size_t npix = (size_t)rc->rows * rc->cols;
  XRC_CUDA(cudaMemcpyAsync(host_dst, rc_proj_buf(rc) + first * npix, count * npix * sizeof(float), cudaMemcpyDeviceToHost,
                           rc->ctx->stream));
  XRC_CUDA(cudaStreamSynchronize(rc->ctx->stream));
  return XRC_OK;
}

int xrc_rc_use_other_proj_buf(xrc_rc* rc, xrc_rc* other)
{
  XRC_CHECK_ARG(rc && other && rc != other, "xrc_rc_use_other_proj_buf: bad argument");
  XRC_CHECK_ARG(rc->ctx == other->ctx, "xrc_rc_use_other_proj_buf: ray casters live on different contexts");
  XRC_CHECK_ARG(!rc->allocated, "xrc_rc_use_other_proj_buf: must be called before allocation");
  for (const xrc_rc* r = other; r; r = r->other)
    XRC_CHECK_ARG(r != rc, "xrc_rc_use_other_proj_buf: circular buffer sharing");
  if (rc->other)
    --rc->other->n_borrowers;
  rc->other = other;
  ++other->n_borrowers;
  return XRC_OK;
}

int xrc_rc_ray_info(xrc_rc* rc, uint32_t vol_idx, uint8_t* host_mask, uint32_t* host_steps, uint64_t* total_samples)
{
  XRC_CHECK_ARG(rc, "null ray caster");
  XRC_CHECK_ARG(rc->allocated, "xrc_rc_ray_info: allocate first");
  XRC_CHECK_ARG(vol_idx < rc->vols.size(), "xrc_rc_ray_info: volume index out of range");
  XRC_TRY(use_device(rc->ctx));
  cudaStream_t st = rc->ctx->stream;
  const size_t n = (size_t)rc->rows * rc->cols * rc->num_projs;
  uint8_t* d_mask = nullptr;
  uint32_t* d_steps = nullptr;
  unsigned long long* d_cnt = nullptr;
  int status = XRC_OK;
  do
  {
    if (host_mask && cudaMalloc(&d_mask, n) != cudaSuccess) { status = XRC_ERR_NOMEM; break; }
    if (host_steps && cudaMalloc(&d_steps, n * sizeof(uint32_t)) != cudaSuccess) { status = XRC_ERR_NOMEM; break; }
    if (cudaMalloc(&d_cnt, sizeof(unsigned long long)) != cudaSuccess) { status = XRC_ERR_NOMEM; break; }
    cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), st);
    DrrArgs a;
    rc_fill_args(rc, vol_idx, &a);
    a.ray_mask = d_mask;
    a.ray_steps = d_steps;
    a.sample_counter = d_cnt;
    status = launch_ray_info(a, st);
    if (status != XRC_OK)
      break;
    unsigned long long cnt = 0;
    if (host_mask)
      cudaMemcpyAsync(host_mask, d_mask, n, cudaMemcpyDeviceToHost, st);
    if (host_steps)
      cudaMemcpyAsync(host_steps, d_steps, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, st);
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
    {
      set_error(std::string("xrc_rc_ray_info: ") + cudaGetErrorString(e));
      status = XRC_ERR_CUDA;
      break;
    }
    if (total_samples)
      *total_samples = cnt;
  } while (0);
  if (status == XRC_ERR_NOMEM)
    set_error("xrc_rc_ray_info: out of device memory");
  cudaFree(d_mask);
  cudaFree(d_steps);
  cudaFree(d_cnt);
  return status;
}

int xrc_rc_set_skip_empty(xrc_rc* rc, int enable)
{
  XRC_CHECK_ARG(rc, "null ray caster");
  XRC_CHECK_ARG(enable >= 0 && enable <= 3, "xrc_rc_set_skip_empty: 0 off, 1 (or 2) on, 3 on with interior gaps");
  rc->skip_empty = enable != 0;
  rc->gap_mode = (enable == 3) ? 1 : 0;
  return XRC_OK;
}

int xrc_rc_fetched_samples(xrc_rc* rc, uint32_t vol_idx, uint64_t* fetched)
{
  XRC_CHECK_ARG(rc && fetched, "null argument");
  XRC_CHECK_ARG(rc->allocated, "xrc_rc_fetched_samples: allocate first");
  XRC_CHECK_ARG(vol_idx < rc->vols.size(), "xrc_rc_fetched_samples: volume index out of range");
  XRC_CHECK_ARG(rc->vols[vol_idx].layout == XRC_LAYOUT_PAX, "xrc_rc_fetched_samples: only for the PAX layout");
  XRC_TRY(use_device(rc->ctx));
  cudaStream_t st = rc->ctx->stream;
  unsigned long long* d_cnt = nullptr;
  XRC_CUDA(cudaMalloc(&d_cnt, sizeof(unsigned long long)));
  cudaMemsetAsync(d_cnt, 0, sizeof(unsigned long long), st);
  if (rc_prepare_stacks(rc, vol_idx) != XRC_OK)
  {
    cudaFree(d_cnt);
    return XRC_ERR_CUDA;
  }
  DrrArgs a;
  rc_fill_args(rc, vol_idx, &a);
  a.sample_counter = d_cnt;
  a.count_only = 1;
  int status = launch_drr(a, XRC_LAYOUT_PAX, rc->kernel_id, st);
  unsigned long long cnt = 0;
  if (status == XRC_OK)
  {
    cudaMemcpyAsync(&cnt, d_cnt, sizeof(cnt), cudaMemcpyDeviceToHost, st);
    const cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess)
    {
      set_error(std::string("xrc_rc_fetched_samples: ") + cudaGetErrorString(e));
      status = XRC_ERR_CUDA;
    }
  }
  cudaFree(d_cnt);
  *fetched = cnt;
  return status;
}

// internal tuning hook (not part of the documented ABI surface): CTA ordering
int xrc_rc_set_cta_order(xrc_rc* rc, int order)
{
  XRC_CHECK_ARG(rc && order >= 0 && order < (1 << 16), "bad argument");
  rc->order = order;  // bit 0: CTA order; bits 1..: kernel variant (measurement only)
  return XRC_OK;
}

// ---------------------------------------------------------------- projection pre-processing
// itk::GaussianOperator<double>::GenerateCoefficients (ITK 5.1.1, restated: discrete Gaussian e^-t I_n(t) from the modified
// Bessel functions -- Abramowitz & Stegun 9.8.1-9.8.4 polynomials, downward recurrence with accuracy 40 --, summed until
// 1 - max error, at most max_width wide, normalised).  Returns the radius; k[0 .. 2 radius].
static double itk_bessel_i0(double y)
{
  const double d = fabs(y);
  if (d < 3.75)
  {
    double m = y / 3.75;
    m *= m;
    return 1.0 + m * (3.5156229 + m * (3.0899424 + m * (1.2067492 + m * (0.2659732 + m * (0.360768e-1 + m * 0.45813e-2)))));
  }
  const double m = 3.75 / d;
  return (exp(d) / sqrt(d)) *
         (0.39894228 + m * (0.1328592e-1 + m * (0.225319e-2 + m * (-0.157565e-2 + m * (0.916281e-2 + m * (-0.2057706e-1 +
          m * (0.2635537e-1 + m * (-0.1647633e-1 + m * 0.392377e-2))))))));
}
static double itk_bessel_i1(double y)
{
  const double d = fabs(y);
  double acc;
  if (d < 3.75)
  {
    double m = y / 3.75;
    m *= m;
    acc = d * (0.5 + m * (0.87890594 + m * (0.51498869 + m * (0.15084934 + m * (0.2658733e-1 + m * (0.301532e-2 + m * 0.32411e-3))))));
  }
  else
  {
    const double m = 3.75 / d;
    acc = 0.2282967e-1 + m * (-0.2895312e-1 + m * (0.1787654e-1 - m * 0.420059e-2));
    acc = 0.39894228 + m * (-0.3988024e-1 + m * (-0.362018e-2 + m * (0.163801e-2 + m * (-0.1031555e-1 + m * acc))));
    acc *= (exp(d) / sqrt(d));
  }
  return (y < 0.0) ? -acc : acc;
}
static double itk_bessel_i(int n, double y)
{
  if (y == 0.0)
    return 0.0;
  const double toy = 2.0 / fabs(y);
  double qip = 0.0, acc = 0.0, qi = 1.0;
  for (int j = 2 * (n + (int)sqrt(40.0 * n)); j > 0; j--)
  {
    const double qim = qip + j * toy * qi;
    qip = qi;
    qi = qim;
    if (fabs(qi) > 1.0e10)
    {
      acc *= 1.0e-10;
      qi *= 1.0e-10;
      qip *= 1.0e-10;
    }
    if (j == n)
      acc = qip;
  }
  acc *= itk_bessel_i0(y) / qi;
  return (y < 0.0 && (n & 1)) ? -acc : acc;
}
static int itk_gaussian_coeffs(double variance, double max_error, int max_width, double* k)
{
  double half[40];
  const double et = exp(-variance), cap = 1.0 - max_error;
  int n = 0;
  half[n++] = et * itk_bessel_i0(variance);
  double sum = half[0];
  half[n++] = et * itk_bessel_i1(variance);
  sum += half[1] * 2.0;
  for (int i = 2; sum < cap; i++)
  {
    half[n++] = et * itk_bessel_i(i, variance);
    sum += half[i] * 2.0;
    if (half[i] <= 0.0 || n > max_width || n >= 33)
      break;
  }
  for (int i = 0; i < n; ++i)
    half[i] /= sum;
  const int r = n - 1;
  for (int i = 0; i <= r; ++i)
    k[r + i] = k[r - i] = half[i];
  return r;
}

// ImageIntensLogTransFilter::GenerateData (lib/image/xregImageIntensLogTrans.cpp:55-144) on the device
int xrc_log_remap(xrc_ctx* ctx, const float* host_img, uint32_t rows, uint32_t cols, int normalize_zero_one,
                  int use_max_intensity_as_I0, float I0, float* host_out, float* I0_used)
{
  XRC_CHECK_ARG(ctx && host_img && host_out && rows > 0 && cols > 0, "xrc_log_remap: bad argument");
  XRC_TRY(use_device(ctx));
  const size_t n = (size_t)rows * cols;
  cudaStream_t st = ctx->stream;
  float *d_img = nullptr, *d_a = nullptr, *d_b = nullptr, *d_mm = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_img);
    cudaFree(d_a);
    cudaFree(d_b);
    cudaFree(d_mm);
  };
  if (cudaMalloc(&d_img, n * sizeof(float)) != cudaSuccess || cudaMalloc(&d_a, n * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&d_b, n * sizeof(float)) != cudaSuccess || cudaMalloc(&d_mm, 2 * sizeof(float)) != cudaSuccess)
  {
    cleanup();
    cudaGetLastError();
    XRC_FAIL(XRC_ERR_NOMEM, "xrc_log_remap: out of device memory");
  }
  const float eps = 1.0e-6f;
  float mm[2] = {0.f, 0.f};
  float scale = 1.0f, I0_to_use = I0;
  int status = XRC_OK;
  auto sync_mm = [&]() -> int {
    if (cudaMemcpyAsync(mm, d_mm, sizeof(mm), cudaMemcpyDeviceToHost, st) != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess)
    {
      set_error("xrc_log_remap: device failure");
      return XRC_ERR_CUDA;
    }
    return XRC_OK;
  };
  if (cudaMemcpyAsync(d_img, host_img, n * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess)
    status = XRC_ERR_CUDA;
  if (status == XRC_OK && normalize_zero_one)
  {
    // scale to [0, 1] by 1 / max (:74-85); I0 = 1 when the maximum is to be I0 (:89-92)
    status = launch_minmax(d_img, n, 1.0f, eps, d_mm, st);
    if (status == XRC_OK)
      status = sync_mm();
    scale = 1.0f / mm[0];
    if (use_max_intensity_as_I0)
      I0_to_use = 1.0f;
  }
  else if (status == XRC_OK && use_max_intensity_as_I0)
  {
    // I0 = max of the image smoothed with a discrete Gaussian of variance 2 (:95-105)
    ItkGaussArgs g;
    memset(&g, 0, sizeof(g));
    g.rows = rows;
    g.cols = cols;
    g.radius = itk_gaussian_coeffs(2.0, 0.01, 32, g.k);
    g.src = d_img;
    g.dst = d_a;
    g.along_x = 0;
    status = launch_itk_gauss(g, st);
    g.src = d_a;
    g.dst = d_b;
    g.along_x = 1;
    if (status == XRC_OK)
      status = launch_itk_gauss(g, st);
    if (status == XRC_OK)
      status = launch_minmax(d_b, n, 1.0f, eps, d_mm, st);
    if (status == XRC_OK)
      status = sync_mm();
    I0_to_use = mm[0];
  }
  if (status == XRC_OK)
  {
    // smallest pixel above eps of the (scaled) image -> what zero pixels map to (:108-133)
    status = launch_minmax(d_img, n, scale, eps, d_mm, st);
    if (status == XRC_OK)
      status = sync_mm();
  }
  if (status == XRC_OK)
  {
    const volatile float q = mm[1] / I0_to_use;
    const float out_max = -logf(q);
    status = launch_log_map(d_img, d_a, n, scale, 1, eps, I0_to_use, out_max, st);
  }
  if (status == XRC_OK && (cudaMemcpyAsync(host_out, d_a, n * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                           cudaStreamSynchronize(st) != cudaSuccess))
  {
    set_error("xrc_log_remap: device failure");
    status = XRC_ERR_CUDA;
  }
  cleanup();
  if (status == XRC_OK && I0_used)
    *I0_used = I0_to_use;
  return status;
}

// DownsampleImage (lib/itk/xregITKResampleUtils.h:49-112, the cubic B-spline default of :181-188) on the device
int xrc_downsample_size(uint32_t rows, uint32_t cols, double factor, uint32_t* out_rows, uint32_t* out_cols)
{
  XRC_CHECK_ARG(out_rows && out_cols && factor > 0.0, "xrc_downsample_size: bad argument");
  *out_cols = (uint32_t)(unsigned long)((double)cols * factor + 0.5);   // :101-105
  *out_rows = (uint32_t)(unsigned long)((double)rows * factor + 0.5);
  return XRC_OK;
}

int xrc_downsample_image(xrc_ctx* ctx, const float* host_img, uint32_t rows, uint32_t cols, double factor, double sigma,
                         float* host_out)
{
  XRC_CHECK_ARG(ctx && host_img && host_out && rows > 0 && cols > 0 && factor > 0.0, "xrc_downsample_image: bad argument");
  uint32_t orows = 0, ocols = 0;
  XRC_TRY(xrc_downsample_size(rows, cols, factor, &orows, &ocols));
  XRC_CHECK_ARG(orows > 0 && ocols > 0, "xrc_downsample_image: the factor leaves no pixel");
  XRC_TRY(use_device(ctx));
  const size_t n = (size_t)rows * cols, n_out = (size_t)orows * ocols;
  cudaStream_t st = ctx->stream;
  float *d_img = nullptr, *d_a = nullptr, *d_b = nullptr, *d_out = nullptr;
  double* d_c = nullptr;
  auto cleanup = [&]() {
    cudaFree(d_img);
    cudaFree(d_a);
    cudaFree(d_b);
    cudaFree(d_out);
    cudaFree(d_c);
  };
  if (cudaMalloc(&d_img, n * sizeof(float)) != cudaSuccess || cudaMalloc(&d_a, n * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&d_b, n * sizeof(float)) != cudaSuccess || cudaMalloc(&d_out, n_out * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&d_c, n * sizeof(double)) != cudaSuccess)
  {
    cleanup();
    cudaGetLastError();
    XRC_FAIL(XRC_ERR_NOMEM, "xrc_downsample_image: out of device memory");
  }
  int status = XRC_OK;
  if (cudaMemcpyAsync(d_img, host_img, n * sizeof(float), cudaMemcpyHostToDevice, st) != cudaSuccess)
    status = XRC_ERR_CUDA;
  const float* d_src = d_img;
  if (status == XRC_OK && (factor < 1.0) && (fabs(sigma) > 1.0e-6))   // smooth before down-sampling (:68-86)
  {
    const double sg = (sigma < 0.0) ? (0.5 / factor) : sigma;
    ItkGaussArgs g;
    memset(&g, 0, sizeof(g));
    g.rows = rows;
    g.cols = cols;
    g.radius = itk_gaussian_coeffs(sg * sg, 0.01, 32, g.k);
    g.src = d_img;
    g.dst = d_a;
    g.along_x = 0;
    status = launch_itk_gauss(g, st);
    g.src = d_a;
    g.dst = d_b;
    g.along_x = 1;
    if (status == XRC_OK)
      status = launch_itk_gauss(g, st);
    d_src = d_b;
  }
  // itk::BSplineDecompositionImageFilter (order 3), dimension 0 first, then the 16-tap evaluation per output pixel
  const double z = sqrt(3.0) - 2.0;
  const int64_t horizon = (int64_t)ceil(log(1.0e-10) / log(fabs(z)));
  if (status == XRC_OK)
    status = launch_f32_to_f64(d_src, d_c, n, st);
  if (status == XRC_OK)
    status = launch_bspline_prefilter(d_c, rows, cols, 1, pow(z, (double)((int64_t)cols - 1)), horizon, st);
  if (status == XRC_OK)
    status = launch_bspline_prefilter(d_c, rows, cols, 0, pow(z, (double)((int64_t)rows - 1)), horizon, st);
  if (status == XRC_OK)
    status = launch_bspline_resample(d_c, rows, cols, d_out, orows, ocols, factor, st);
  if (status == XRC_OK && (cudaMemcpyAsync(host_out, d_out, n_out * sizeof(float), cudaMemcpyDeviceToHost, st) != cudaSuccess ||
                           cudaStreamSynchronize(st) != cudaSuccess))
  {
    set_error("xrc_downsample_image: device failure");
    status = XRC_ERR_CUDA;
  }
  cleanup();
  return status;
}

// ---------------------------------------------------------------- metrics
static bool sm_is_patch_kind(int kind) { return kind == XRC_SM_PATCH_NCC || kind == XRC_SM_PATCH_GRAD_NCC; }

int xrc_sm_create(xrc_ctx* ctx, int kind, xrc_sm** out)
{
  XRC_CHECK_ARG(ctx && out, "xrc_sm_create: null argument");
  XRC_CHECK_ARG(kind >= XRC_SM_NCC && kind <= XRC_SM_SSD, "xrc_sm_create: unknown metric kind");
  xrc_sm* sm = new xrc_sm;
  sm->ctx = ctx;
  sm->kind = kind;
  *out = sm;
  return XRC_OK;
}

static void sm_free_resources(xrc_sm* sm)
{
  dfree(sm->d_fixed);
  dfree(sm->d_mask);
  for (int d = 0; d < 2; ++d)
  {
    dfree(sm->d_f0[d]);
    if (sm->d_fg[d] != nullptr && sm->kind == XRC_SM_PATCH_GRAD_NCC)
      cudaFree(sm->d_fg[d]);
    sm->d_fg[d] = nullptr;
    dfree(sm->d_g[d]);
    dfree(sm->d_pmean[d]);
    dfree(sm->d_pden[d]);
    dfree(sm->d_psmask[d]);
  }
  dfree(sm->d_pnmask);
  dfree(sm->d_mov);
  dfree(sm->d_partials);
  dfree(sm->d_sims);
  dfree(sm->d_weights);
  dfree(sm->d_vals);
  dfree(sm->d_seq);
  dfree(sm->d_subset);
  dfree(sm->d_sub_vals);
  sm->d_subset_cap = sm->d_sub_vals_cap = 0;
  sm->subset_dirty = !sm->h_subset.empty();
  if (sm->h_sims)
    cudaFreeHost(sm->h_sims);
  sm->h_sims = nullptr;
  sm->allocated = false;
}

int xrc_sm_destroy(xrc_sm* sm)
{
  if (!sm)
    return XRC_OK;
  cudaSetDevice(sm->ctx->device);
  cudaStreamSynchronize(sm->ctx->stream);
  sm_free_resources(sm);
  delete sm;
  return XRC_OK;
}

int xrc_sm_set_fixed(xrc_sm* sm, const float* host_img, uint32_t rows, uint32_t cols)
{
  XRC_CHECK_ARG(sm && host_img && rows > 0 && cols > 0, "xrc_sm_set_fixed: bad argument");
  XRC_CHECK_ARG(!sm->allocated || (rows == sm->rows && cols == sm->cols),
                "xrc_sm_set_fixed: image size changed after allocation");
  sm->rows = rows;
  sm->cols = cols;
  sm->h_fixed.assign(host_img, host_img + (size_t)rows * cols);
  sm->fixed_dirty = true;
  return XRC_OK;
}

int xrc_sm_set_mask(xrc_sm* sm, const uint8_t* host_mask)
{
  XRC_CHECK_ARG(sm, "null metric");
  if (host_mask)
  {
    XRC_CHECK_ARG(sm->rows && sm->cols, "xrc_sm_set_mask: set the fixed image first");
    sm->h_mask.assign(host_mask, host_mask + (size_t)sm->rows * sm->cols);
    sm->has_mask = true;
  }
  else
  {
    sm->h_mask.clear();
    sm->has_mask = false;
  }
  sm->fixed_dirty = true;
  return XRC_OK;
}

int xrc_sm_set_grad_params(xrc_sm* sm, uint32_t gauss_width)
{
  XRC_CHECK_ARG(sm, "null metric");
  XRC_CHECK_ARG(gauss_width == 0 || (gauss_width & 1), "xrc_sm_set_grad_params: smoothing kernel width must be odd or 0");
  if (gauss_width > (uint32_t)kMaxGaussWidth)
    XRC_FAIL(XRC_ERR_UNSUPPORTED, "xrc_sm_set_grad_params: smoothing kernel wider than 31 is not supported");
  sm->gauss_width = gauss_width;
  sm->fixed_dirty = true;
  return XRC_OK;
}

int xrc_seqsum_f32(xrc_ctx* ctx, const float* host_vals, uint32_t n_seq, uint64_t n, int serial, float* host_out)
{
  XRC_CHECK_ARG(ctx && host_vals && host_out && n_seq > 0, "xrc_seqsum_f32: bad argument");
  XRC_TRY(use_device(ctx));
  float* d_v = nullptr;
  float* d_o = nullptr;
  const size_t len = (size_t)n_seq * n;
  XRC_CUDA(cudaMalloc(&d_v, std::max<size_t>(len, 1) * sizeof(float)));
  if (cudaMalloc(&d_o, n_seq * sizeof(float)) != cudaSuccess)
  {
    cudaFree(d_v);
    XRC_FAIL(XRC_ERR_NOMEM, "xrc_seqsum_f32: out of device memory");
  }
  int status = XRC_OK;
  if (cudaMemcpyAsync(d_v, host_vals, len * sizeof(float), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
    status = XRC_ERR_CUDA;
  if (status == XRC_OK)
  {
    SeqSumArgs q;
    q.vals = d_v;
    q.n = n;
    q.n_seq = n_seq;
    q.out = d_o;
    q.serial = serial;
    status = launch_seqsum(q, ctx->stream);
  }
  if (status == XRC_OK && (cudaMemcpyAsync(host_out, d_o, n_seq * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess ||
                           cudaStreamSynchronize(ctx->stream) != cudaSuccess))
  {
    set_error("xrc_seqsum_f32: device failure");
    status = XRC_ERR_CUDA;
  }
  cudaFree(d_v);
  cudaFree(d_o);
  return status;
}

int xrc_sm_set_combine_mode(xrc_sm* sm, int mode)
{
  XRC_CHECK_ARG(sm, "null metric");
  XRC_CHECK_ARG(mode == XRC_COMBINE_REFERENCE || mode == XRC_COMBINE_REFERENCE_SERIAL || mode == XRC_COMBINE_F64,
                "xrc_sm_set_combine_mode: unknown mode");
  sm->combine_mode = mode;
  return XRC_OK;
}

int xrc_sm_set_patch_params(xrc_sm* sm, uint32_t radius, uint32_t stride, int compute_mean_of_patch_sims,
                            int weight_patch_sims_in_combine, int use_mask_for_patch_stats, const float* weights,
                            uint64_t n_weights)
{
  XRC_CHECK_ARG(sm, "null metric");
  XRC_CHECK_ARG(stride >= 1, "xrc_sm_set_patch_params: stride must be >= 1");
  if (2 * radius + 64 > (uint32_t)kPatchThreads)
    XRC_FAIL(XRC_ERR_UNSUPPORTED, "xrc_sm_set_patch_params: patch radius above 96 is not supported");
  XRC_CHECK_ARG(!sm->allocated || (radius == sm->radius && stride == sm->stride),
                "xrc_sm_set_patch_params: patch grid cannot change after allocation (patches_setup_)");
  sm->radius = radius;
  sm->stride = stride;
  sm->compute_mean = compute_mean_of_patch_sims;
  sm->weight_sims = weight_patch_sims_in_combine;
  sm->mask_stats = use_mask_for_patch_stats;
  if (weights)
  {
    sm->h_weights.assign(weights, weights + n_weights);
    sm->has_weights = true;
  }
  else
  {
    sm->h_weights.clear();
    sm->has_weights = false;
  }
  sm->fixed_dirty = true;
  return XRC_OK;
}

int xrc_sm_set_patch_subset(xrc_sm* sm, const uint64_t* patch_inds, uint64_t n)
{
  XRC_CHECK_ARG(sm, "null metric");
  XRC_CHECK_ARG(sm_is_patch_kind(sm->kind), "xrc_sm_set_patch_subset: not a patch metric");
  XRC_CHECK_ARG(n == 0 || patch_inds, "xrc_sm_set_patch_subset: null index list");
  XRC_CHECK_ARG(n < (1ull << 31), "xrc_sm_set_patch_subset: too many indices");
  if (sm->allocated)
    for (uint64_t j = 0; j < n; ++j)
      XRC_CHECK_ARG(patch_inds[j] < sm->n_patches, "xrc_sm_set_patch_subset: patch index outside the patch grid");
  sm->h_subset.resize((size_t)n);
  for (uint64_t j = 0; j < n; ++j)
  {
    XRC_CHECK_ARG(patch_inds[j] < (1ull << 32), "xrc_sm_set_patch_subset: patch index outside the patch grid");
    sm->h_subset[(size_t)j] = (uint32_t)patch_inds[j];
  }
  sm->subset_dirty = true;
  return XRC_OK;
}

int xrc_sm_bind_ray_caster(xrc_sm* sm, xrc_rc* rc, uint32_t proj_offset)
{
  XRC_CHECK_ARG(sm && rc, "xrc_sm_bind_ray_caster: null argument");
  XRC_CHECK_ARG(sm->ctx == rc->ctx, "xrc_sm_bind_ray_caster: metric and ray caster live on different contexts");
  // xregImgSimMetric2DCPU.cpp:45-70: re-binding is only allowed to the same ray caster
  XRC_CHECK_ARG(!sm->allocated || sm->rc == rc, "xrc_sm_bind_ray_caster: cannot switch ray caster after allocation");
  sm->rc = rc;
  sm->host_src = nullptr;
  sm->dev_src = nullptr;
  sm->proj_offset = proj_offset;
  return XRC_OK;
}

int xrc_sm_bind_host(xrc_sm* sm, const float* host_buf, uint32_t proj_offset)
{
  XRC_CHECK_ARG(sm && host_buf, "xrc_sm_bind_host: null argument");
  XRC_CHECK_ARG(!sm->allocated || sm->d_mov, "xrc_sm_bind_host: metric was allocated for a device source");
  sm->rc = nullptr;
  sm->dev_src = nullptr;
  sm->host_src = host_buf;
  sm->proj_offset = proj_offset;
  return XRC_OK;
}

int xrc_sm_bind_device(xrc_sm* sm, const float* dev_buf, uint32_t proj_offset)
{
  XRC_CHECK_ARG(sm && dev_buf, "xrc_sm_bind_device: null argument");
  sm->rc = nullptr;
  sm->host_src = nullptr;
  sm->dev_src = dev_buf;
  sm->proj_offset = proj_offset;
  return XRC_OK;
}

static bool sm_is_patch(const xrc_sm* sm) { return sm->kind == XRC_SM_PATCH_NCC || sm->kind == XRC_SM_PATCH_GRAD_NCC; }
static bool sm_is_grad(const xrc_sm* sm) { return sm->kind == XRC_SM_GRAD_NCC || sm->kind == XRC_SM_PATCH_GRAD_NCC; }

static void sm_fill_grad_args(const xrc_sm* sm, GradArgs* g)
{
  memset(g, 0, sizeof(*g));
  g->rows = sm->rows;
  g->cols = sm->cols;
  g->gauss_width = (int)sm->gauss_width;
  if (sm->gauss_width > 1)
    gauss_coeffs((int)sm->gauss_width, g->coeffs);
}

// zero-mean fixed image + statistics (ImgSimMetric2DNCCCPU::process_mask,
// xregImgSimMetric2DNCCCPU.cpp:211-236) for one direction
static void ncc_fixed_stats(const std::vector<float>& f, const std::vector<uint8_t>* mask, std::vector<float>* f0,
                            float* sd_out, double* sf0_out, double* n_eff)
{
  const size_t n = f.size();
  double s = 0, cnt = 0;
  for (size_t i = 0; i < n; ++i)
  {
    if (!mask || (*mask)[i])
    {
      s += f[i];
      cnt += 1;
    }
  }
  const float mean = (float)(s / cnt);
  f0->resize(n);
  double ss = 0, s0 = 0;
  for (size_t i = 0; i < n; ++i)
  {
    const float z = f[i] - mean;
    (*f0)[i] = z;
    if (!mask || (*mask)[i])
    {
      ss += (double)z * z;
      s0 += z;
    }
  }
  const float sd = (float)sqrt(ss / (cnt - 1.0));
  *sd_out = std::max(1.0e-6f, sd);
  *sf0_out = s0;
  *n_eff = cnt;
}

// Re-derive everything that depends on the fixed image / mask / parameters
// (process_updated_mask semantics, xregImgSimMetric2D.h:146-148).
static int sm_prepare_fixed(xrc_sm* sm)
{
  cudaStream_t st = sm->ctx->stream;
  const size_t npix = (size_t)sm->rows * sm->cols;
  XRC_CUDA(cudaMemcpyAsync(sm->d_fixed, sm->h_fixed.data(), npix * sizeof(float), cudaMemcpyHostToDevice, st));
  if (sm->has_mask)
  {
    if (!sm->d_mask)
      XRC_CUDA(cudaMalloc(&sm->d_mask, npix));
    XRC_CUDA(cudaMemcpyAsync(sm->d_mask, sm->h_mask.data(), npix, cudaMemcpyHostToDevice, st));
  }
  const std::vector<uint8_t>* mask = sm->has_mask ? &sm->h_mask : nullptr;
  const int n_dirs = sm_is_grad(sm) ? 2 : 1;

  // fixed gradient images (ImgSimMetric2DGradImgCPU::allocate_resources, :32-66)
  std::vector<float> fg[2];
  if (sm_is_grad(sm))
  {
    GradArgs g;
    sm_fill_grad_args(sm, &g);
    g.src = sm->d_fixed;
    g.n_imgs = 1;
    float* tx = sm->d_fg[0];
    float* ty = sm->d_fg[1];
    if (sm->kind == XRC_SM_GRAD_NCC)
    {
      tx = sm->d_f0[0];  // overwritten by the zero-mean version below
      ty = sm->d_f0[1];
    }
    g.gx = tx;
    g.gy = ty;
    XRC_TRY(launch_grad(g, st));
    if (sm->kind == XRC_SM_GRAD_NCC)
    {
      fg[0].resize(npix);
      fg[1].resize(npix);
      XRC_CUDA(cudaMemcpyAsync(fg[0].data(), tx, npix * sizeof(float), cudaMemcpyDeviceToHost, st));
      XRC_CUDA(cudaMemcpyAsync(fg[1].data(), ty, npix * sizeof(float), cudaMemcpyDeviceToHost, st));
      XRC_CUDA(cudaStreamSynchronize(st));
    }
  }

  if (sm->kind == XRC_SM_SSD)
  {
    // ImgSimMetric2DSSDCPU::process_mask (xregImgSimMetric2DSSDCPU.cpp:91-110): fixed image zeroed outside the mask;
    // the divisor is the full pixel count
    std::vector<float> f0(sm->h_fixed);
    double sff = 0.0;
    for (size_t i = 0; i < npix; ++i)
    {
      if (mask && !(*mask)[i])
        f0[i] = 0.0f;
      sff += (double)f0[i] * f0[i];
    }
    sm->sf0[0] = sff;
    sm->n_eff = (double)npix;
    XRC_CUDA(cudaMemcpyAsync(sm->d_f0[0], f0.data(), npix * sizeof(float), cudaMemcpyHostToDevice, st));
    XRC_CUDA(cudaStreamSynchronize(st));  // f0 is a local
  }
  else if (sm->kind == XRC_SM_NCC || sm->kind == XRC_SM_GRAD_NCC)
  {
    for (int d = 0; d < n_dirs; ++d)
    {
      std::vector<float> f0;
      ncc_fixed_stats((sm->kind == XRC_SM_NCC) ? sm->h_fixed : fg[d], mask, &f0, &sm->f_sd[d], &sm->sf0[d], &sm->n_eff);
      XRC_CUDA(cudaMemcpyAsync(sm->d_f0[d], f0.data(), npix * sizeof(float), cudaMemcpyHostToDevice, st));
      XRC_CUDA(cudaStreamSynchronize(st));  // f0 is a local
    }
  }
  else
  {
    // per-patch statistics of the fixed image(s) (ImgSimMetric2DPatchNCCCPU::process_mask, :332-441)
    PatchArgs p;
    memset(&p, 0, sizeof(p));
    p.fix[0] = sm->d_fg[0];
    p.fix[1] = sm->d_fg[1];
    p.mask = sm->has_mask ? sm->d_mask : nullptr;
    p.n_dirs = n_dirs;
    p.rows = sm->rows;
    p.cols = sm->cols;
    p.radius = sm->radius;
    p.stride = 1;  // statistics on the full stride-1 grid
    p.n_strips = sm->n_strips;
    p.mask_mode = !sm->has_mask ? 0 : (sm->mask_stats ? 2 : 1);
    for (int d = 0; d < 2; ++d)
    {
      p.o_mean[d] = sm->d_pmean[d];
      p.o_den[d] = sm->d_pden[d];
      p.o_smask[d] = sm->d_psmask[d];
    }
    p.o_nmask = sm->d_pnmask;
    XRC_TRY(launch_patch_fixed_stats(p, st));

    // weights + divisor (xregImgSimMetric2DPatchNCCCPU.cpp:262-287)
    const uint32_t r = sm->radius, s = sm->stride;
    const uint64_t ncr = (sm->rows - 1 - 2 * r) / s + 1, ncc = (sm->cols - 1 - 2 * r) / s + 1;
    const uint64_t np = ncr * ncc;
    if (sm->has_weights)
    {
      XRC_CHECK_ARG(sm->h_weights.size() == np, "patch weights: expected one weight per patch of the grid");
      XRC_CUDA(cudaMemcpyAsync(sm->d_weights, sm->h_weights.data(), np * sizeof(float), cudaMemcpyHostToDevice, st));
    }
    // the reference's own divisor: f32, tot_wgt accumulated sequentially over ALL patches (:268-284)
    sm->divide_f = (sm->compute_mean || sm->weight_sims) ? 1 : 0;
    if (sm->compute_mean)
    {
      sm->divisor_f = (float)np;
    }
    else if (sm->weight_sims)
    {
      volatile float tw = 0.0f;  // volatile: one rounding per addition, in order
      for (uint64_t k = 0; k < np; ++k)
        tw = tw + (sm->has_weights ? sm->h_weights[k] : 1.0f);
      sm->divisor_f = tw;
    }
    if (sm->compute_mean)
    {
      sm->divisor = (double)np;
    }
    else if (sm->weight_sims)
    {
      // tot_wgt (:277-284), summed in f64 like the kernel's sum of w_k s_k so that the ratio carries no
      // f32 accumulation error (the reference's two sequential f32 sums have correlated rounding errors)
      double tw = 0.0;
      for (uint64_t k = 0; k < np; ++k)
        tw += sm->has_weights ? (double)sm->h_weights[k] : 1.0;
      sm->divisor = tw;
    }
    else
    {
      sm->divisor = 1.0;
    }
  }
  XRC_CUDA(cudaStreamSynchronize(st));
  sm->fixed_dirty = false;
  return XRC_OK;
}

int xrc_sm_allocate(xrc_sm* sm, uint32_t max_imgs)
{
  XRC_CHECK_ARG(sm, "null metric");
  XRC_CHECK_ARG(max_imgs > 0, "xrc_sm_allocate: need at least one moving image");
  XRC_CHECK_ARG(!sm->h_fixed.empty(), "xrc_sm_allocate: set the fixed image first");
  XRC_CHECK_ARG(sm->rc || sm->host_src || sm->dev_src, "xrc_sm_allocate: bind a moving-image source first");
  if (sm->rc)
  {
    XRC_CHECK_ARG(sm->rc->allocated, "xrc_sm_allocate: the ray caster must be allocated first (xregImgSimMetric2D.h:116-118)");
    XRC_CHECK_ARG(sm->rc->rows == sm->rows && sm->rc->cols == sm->cols,
                  "xrc_sm_allocate: fixed image and detector sizes differ");
  }
  if (sm_is_patch(sm))
  {
    XRC_CHECK_ARG(2 * sm->radius + 1 <= sm->rows && 2 * sm->radius + 1 <= sm->cols,
                  "xrc_sm_allocate: patch diameter exceeds the image (xregImgSimMetric2DPatchCommon.cpp:272-273)");
  }
  XRC_TRY(use_device(sm->ctx));
  XRC_CUDA(cudaStreamSynchronize(sm->ctx->stream));
  sm_free_resources(sm);

  const size_t npix = (size_t)sm->rows * sm->cols;
  const int n_dirs = sm_is_grad(sm) ? 2 : 1;
  XRC_CUDA(cudaMalloc(&sm->d_fixed, npix * sizeof(float)));
  XRC_CUDA(cudaMalloc(&sm->d_sims, max_imgs * sizeof(float)));
  XRC_CUDA(cudaMemsetAsync(sm->d_sims, 0, max_imgs * sizeof(float), sm->ctx->stream));
  XRC_CUDA(cudaHostAlloc(&sm->h_sims, max_imgs * sizeof(float), cudaHostAllocMapped | cudaHostAllocPortable));
  memset(sm->h_sims, 0, max_imgs * sizeof(float));
  if (sm->host_src)
    XRC_CUDA(cudaMalloc(&sm->d_mov, npix * max_imgs * sizeof(float)));

  size_t parts_per_img = 0;
  if (sm->kind == XRC_SM_NCC || sm->kind == XRC_SM_SSD)
  {
    XRC_CUDA(cudaMalloc(&sm->d_f0[0], npix * sizeof(float)));
    parts_per_img = ((npix + kMomChunk - 1) / kMomChunk) * 3;
  }
  else if (sm->kind == XRC_SM_GRAD_NCC)
  {
    XRC_CUDA(cudaMalloc(&sm->d_f0[0], npix * sizeof(float)));
    XRC_CUDA(cudaMalloc(&sm->d_f0[1], npix * sizeof(float)));
    // the smoothing width may still change after allocation: size for the largest decomposition
    const size_t parts = std::max(grad_num_parts(sm->rows, sm->cols, 7, kGradBandRowsMin), grad_num_parts(sm->rows, sm->cols, 9));
    parts_per_img = parts * 6;
  }
  else
  {
    if (sm->kind == XRC_SM_PATCH_GRAD_NCC)
    {
      for (int d = 0; d < 2; ++d)
      {
        XRC_CUDA(cudaMalloc(&sm->d_fg[d], npix * sizeof(float)));
        XRC_CUDA(cudaMalloc(&sm->d_g[d], npix * max_imgs * sizeof(float)));
      }
    }
    else
    {
      sm->d_fg[0] = sm->d_fixed;  // alias, not owned
    }
    const uint32_t r = sm->radius;
    const size_t grid1 = (size_t)(sm->rows - 2 * r) * (sm->cols - 2 * r);
    for (int d = 0; d < n_dirs; ++d)
    {
      XRC_CUDA(cudaMalloc(&sm->d_pmean[d], grid1 * sizeof(double)));
      XRC_CUDA(cudaMalloc(&sm->d_pden[d], grid1 * sizeof(float)));
      XRC_CUDA(cudaMalloc(&sm->d_psmask[d], grid1 * sizeof(double)));
    }
    XRC_CUDA(cudaMalloc(&sm->d_pnmask, grid1 * sizeof(float)));
    sm->n_strips = patch_max_parts(sm->rows, sm->cols, r);  // capacity: partial sums per image and direction
    parts_per_img = (size_t)n_dirs * sm->n_strips;
    const uint64_t np = (uint64_t)((sm->rows - 1 - 2 * r) / sm->stride + 1) * ((sm->cols - 1 - 2 * r) / sm->stride + 1);
    XRC_CUDA(cudaMalloc(&sm->d_weights, np * sizeof(float)));
    sm->n_patches = np;
    if (sm->combine_mode != XRC_COMBINE_F64)
    {
      XRC_CUDA(cudaMalloc(&sm->d_vals, (size_t)max_imgs * n_dirs * np * sizeof(float)));
      XRC_CUDA(cudaMalloc(&sm->d_seq, (size_t)max_imgs * n_dirs * sizeof(float)));
    }
  }
  sm->partials_len = parts_per_img * max_imgs;
  XRC_CUDA(cudaMalloc(&sm->d_partials, sm->partials_len * sizeof(double)));
  sm->max_imgs = max_imgs;
  sm->n_imgs = max_imgs;
  sm->allocated = true;
  sm->fixed_dirty = true;
  return sm_prepare_fixed(sm);
}

int xrc_sm_set_num_imgs(xrc_sm* sm, uint32_t n)
{
  XRC_CHECK_ARG(sm, "null metric");
  XRC_CHECK_ARG(!sm->allocated || n <= sm->max_imgs, "xrc_sm_set_num_imgs: exceeds allocated capacity");
  sm->n_imgs = n;
  return XRC_OK;
}

static int sm_source(xrc_sm* sm, const float** src)
{
  const size_t npix = (size_t)sm->rows * sm->cols;
  cudaStream_t st = sm->ctx->stream;
  if (sm->rc)
  {
    XRC_CHECK_ARG((uint64_t)sm->proj_offset + sm->n_imgs <= sm->rc->max_projs &&
                      (uint64_t)sm->proj_offset + sm->n_imgs <= rc_proj_capacity(sm->rc),
                  "metric reads beyond the ray caster's projection buffer");
    *src = rc_proj_buf(sm->rc) + (size_t)sm->proj_offset * npix;
  }
  else if (sm->host_src)
  {
    XRC_CHECK_ARG(sm->d_mov, "metric was not allocated for a host source");
    XRC_CUDA(cudaMemcpyAsync(sm->d_mov, sm->host_src + (size_t)sm->proj_offset * npix, npix * sm->n_imgs * sizeof(float),
                             cudaMemcpyHostToDevice, st));
    *src = sm->d_mov;
  }
  else if (sm->dev_src)
  {
    *src = sm->dev_src + (size_t)sm->proj_offset * npix;
  }
  else
  {
    XRC_FAIL(XRC_ERR_INVALID, "metric has no moving-image source bound");
  }
  return XRC_OK;
}

int xrc_sm_compute(xrc_sm* sm)
{
  XRC_CHECK_ARG(sm, "null metric");
  XRC_CHECK_ARG(sm->allocated, "xrc_sm_compute: resources not allocated");
  XRC_TRY(use_device(sm->ctx));
  if (sm->fixed_dirty)
    XRC_TRY(sm_prepare_fixed(sm));
  if (!sm->n_imgs)
    return XRC_OK;
  cudaStream_t st = sm->ctx->stream;
  const size_t npix = (size_t)sm->rows * sm->cols;
  const float* src = nullptr;
  XRC_TRY(sm_source(sm, &src));
  const uint8_t* mask = sm->has_mask ? sm->d_mask : nullptr;

  if (sm->kind == XRC_SM_NCC || sm->kind == XRC_SM_SSD)
  {
    MomentArgs m;
    memset(&m, 0, sizeof(m));
    m.src = src;
    m.n_imgs = sm->n_imgs;
    m.npix = npix;
    m.f0 = sm->d_f0[0];
    m.mask = mask;
    m.partials = sm->d_partials;
    m.n_chunks = (uint32_t)((npix + kMomChunk - 1) / kMomChunk);
    XRC_TRY(launch_moments(m, st));
    NccFinalizeArgs f;
    memset(&f, 0, sizeof(f));
    f.partials = sm->d_partials;
    f.n_imgs = sm->n_imgs;
    f.n_parts = m.n_chunks;
    f.n_dirs = 1;
    f.ssd = (sm->kind == XRC_SM_SSD) ? 1 : 0;
    f.n_eff = sm->n_eff;
    f.sf0[0] = sm->sf0[0];
    f.f_sd[0] = sm->f_sd[0];
    f.sims = sm->d_sims;
    f.sims_host = sm->h_sims;
    return launch_ncc_finalize(f, st);
  }
  if (sm->kind == XRC_SM_GRAD_NCC)
  {
    GradArgs g;
    sm_fill_grad_args(sm, &g);
    g.src = src;
    g.n_imgs = sm->n_imgs;
    g.f0x = sm->d_f0[0];
    g.f0y = sm->d_f0[1];
    g.mask = mask;
    g.partials = sm->d_partials;
    XRC_TRY(launch_grad(g, st));
    NccFinalizeArgs f;
    memset(&f, 0, sizeof(f));
    f.partials = sm->d_partials;
    f.n_imgs = sm->n_imgs;
    f.n_parts = grad_num_parts(sm->rows, sm->cols, (int)sm->gauss_width,
                               grad_band_rows(sm->rows, sm->cols, (int)sm->gauss_width, sm->n_imgs));
    f.n_dirs = 2;
    f.n_eff = sm->n_eff;
    for (int d = 0; d < 2; ++d)
    {
      f.sf0[d] = sm->sf0[d];
      f.f_sd[d] = sm->f_sd[d];
    }
    f.sims = sm->d_sims;
    f.sims_host = sm->h_sims;
    return launch_ncc_finalize(f, st);
  }

  // patch variants
  const int n_dirs = sm_is_grad(sm) ? 2 : 1;
  PatchArgs p;
  memset(&p, 0, sizeof(p));
  if (sm->kind == XRC_SM_PATCH_GRAD_NCC)
  {
    GradArgs g;
    sm_fill_grad_args(sm, &g);
    g.src = src;
    g.n_imgs = sm->n_imgs;
    g.gx = sm->d_g[0];
    g.gy = sm->d_g[1];
    XRC_TRY(launch_grad(g, st));
    p.mov[0] = sm->d_g[0];
    p.mov[1] = sm->d_g[1];
  }
  else
  {
    p.mov[0] = src;
  }
  p.fix[0] = sm->d_fg[0];
  p.fix[1] = sm->d_fg[1];
  p.mask = mask;
  p.n_imgs = sm->n_imgs;
  p.n_dirs = n_dirs;
  p.rows = sm->rows;
  p.cols = sm->cols;
  p.radius = sm->radius;
  p.stride = sm->stride;
  p.n_strips = sm->n_strips;
  p.mask_mode = !sm->has_mask ? 0 : (sm->mask_stats ? 2 : 1);
  for (int d = 0; d < 2; ++d)
  {
    p.f_mean[d] = sm->d_pmean[d];
    p.f_den[d] = sm->d_pden[d];
    p.f_smask[d] = sm->d_psmask[d];
  }
  p.n_mask = sm->d_pnmask;
  p.weights = sm->has_weights ? sm->d_weights : nullptr;
  p.weight_patch_sims = sm->weight_sims;
  p.partials = sm->d_partials;
  // a patch subset is always combined in the reference's order (its f64 form would need the gathered values anyway)
  const bool use_subset = !sm->h_subset.empty();
  const bool ref_order = sm->combine_mode != XRC_COMBINE_F64 || use_subset;
  if (use_subset)
  {
    const size_t ns = sm->h_subset.size();
    if (sm->subset_dirty)
    {
      for (size_t j = 0; j < ns; ++j)
        XRC_CHECK_ARG(sm->h_subset[j] < sm->n_patches, "patch subset: index outside the patch grid");
      if (sm->d_subset_cap < ns)
      {
        XRC_CUDA(cudaStreamSynchronize(st));
        dfree(sm->d_subset);
        XRC_CUDA(cudaMalloc(&sm->d_subset, ns * sizeof(uint32_t)));
        sm->d_subset_cap = ns;
      }
      const size_t need = (size_t)sm->max_imgs * n_dirs * ns;
      if (sm->d_sub_vals_cap < need)
      {
        XRC_CUDA(cudaStreamSynchronize(st));
        dfree(sm->d_sub_vals);
        XRC_CUDA(cudaMalloc(&sm->d_sub_vals, need * sizeof(float)));
        sm->d_sub_vals_cap = need;
      }
      // pageable source: the copy is staged before the call returns, h_subset may change afterwards
      XRC_CUDA(cudaMemcpyAsync(sm->d_subset, sm->h_subset.data(), ns * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
      // divisor of the subset (xregImgSimMetric2DPatchNCCCPU.cpp:268-285): subset size, or the sequential f32 sum of
      // the subset's weights in subset order
      if (sm->compute_mean)
        sm->sub_divisor_f = (float)ns;
      else if (sm->weight_sims)
      {
        volatile float tw = 0.0f;
        for (size_t j = 0; j < ns; ++j)
          tw = tw + (sm->has_weights ? sm->h_weights[sm->h_subset[j]] : 1.0f);
        sm->sub_divisor_f = tw;
      }
      sm->subset_dirty = false;
    }
  }
  if (ref_order && !sm->d_vals)
  {
    // the mode was switched on after allocation
    XRC_CUDA(cudaMalloc(&sm->d_vals, (size_t)sm->max_imgs * n_dirs * sm->n_patches * sizeof(float)));
    XRC_CUDA(cudaMalloc(&sm->d_seq, (size_t)sm->max_imgs * n_dirs * sizeof(float)));
  }
  p.vals = ref_order ? sm->d_vals : nullptr;
  p.n_patches = sm->n_patches;
  XRC_TRY(launch_patch(p, st));
  PatchFinalizeArgs f;
  memset(&f, 0, sizeof(f));
  if (ref_order)
  {
    SeqSumArgs q;
    q.vals = sm->d_vals;
    q.n = sm->n_patches;
    q.n_seq = sm->n_imgs * n_dirs;
    q.out = sm->d_seq;
    q.serial = (sm->combine_mode == XRC_COMBINE_REFERENCE_SERIAL) ? 1 : 0;
    if (use_subset)
    {
      XRC_TRY(launch_patch_gather(sm->d_vals, sm->d_subset, sm->d_sub_vals, sm->n_patches, (uint32_t)sm->h_subset.size(),
                                  q.n_seq, st));
      q.vals = sm->d_sub_vals;
      q.n = sm->h_subset.size();
    }
    XRC_TRY(launch_seqsum(q, st));
    f.seq_sums = sm->d_seq;
    f.divisor_f = use_subset ? sm->sub_divisor_f : sm->divisor_f;
    f.divide = sm->divide_f;
  }
  f.partials = sm->d_partials;
  f.n_imgs = sm->n_imgs;
  f.n_dirs = n_dirs;
  {
    const PatchPlan pl = patch_plan(sm->rows, sm->cols, sm->radius, sm->n_imgs * n_dirs);
    f.n_parts = pl.n_strips * pl.n_bands;
  }
  f.divisor = sm->divisor;
  f.sims = sm->d_sims;
    f.sims_host = sm->h_sims;
  return launch_patch_finalize(f, st);
}

int xrc_sm_read_sims(xrc_sm* sm, float* host_dst, uint32_t n)
{
  XRC_CHECK_ARG(sm && host_dst, "null argument");
  XRC_CHECK_ARG(sm->allocated, "xrc_sm_read_sims: allocate first");
  XRC_CHECK_ARG(n <= sm->max_imgs, "xrc_sm_read_sims: more values than allocated images");
  XRC_TRY(use_device(sm->ctx));
  XRC_CUDA(cudaMemcpyAsync(sm->h_sims, sm->d_sims, n * sizeof(float), cudaMemcpyDeviceToHost, sm->ctx->stream));
  XRC_CUDA(cudaStreamSynchronize(sm->ctx->stream));
  memcpy(host_dst, sm->h_sims, n * sizeof(float));
  return XRC_OK;
}

int xrc_sm_device_sims(xrc_sm* sm, float** dev_ptr)
{
  XRC_CHECK_ARG(sm && dev_ptr, "null argument");
  XRC_CHECK_ARG(sm->allocated, "xrc_sm_device_sims: allocate first");
  *dev_ptr = sm->d_sims;
  return XRC_OK;
}

int xrc_sm_read_grads(xrc_sm* sm, uint32_t img, float* host_gx, float* host_gy)
{
  XRC_CHECK_ARG(sm && host_gx && host_gy, "null argument");
  XRC_CHECK_ARG(sm->allocated, "xrc_sm_read_grads: allocate first");
  XRC_CHECK_ARG(sm_is_grad(sm), "xrc_sm_read_grads: not a gradient metric");
  XRC_CHECK_ARG(img < sm->n_imgs, "xrc_sm_read_grads: image index out of range");
  XRC_TRY(use_device(sm->ctx));
  cudaStream_t st = sm->ctx->stream;
  const size_t npix = (size_t)sm->rows * sm->cols;
  const float* src = nullptr;
  XRC_TRY(sm_source(sm, &src));
  float* tmp = nullptr;
  XRC_CUDA(cudaMalloc(&tmp, 2 * npix * sizeof(float)));
  GradArgs g;
  sm_fill_grad_args(sm, &g);
  g.src = src + (size_t)img * npix;
  g.n_imgs = 1;
  g.gx = tmp;
  g.gy = tmp + npix;
  int s = launch_grad(g, st);
  if (s == XRC_OK)
  {
    cudaMemcpyAsync(host_gx, tmp, npix * sizeof(float), cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(host_gy, tmp + npix, npix * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess)
    {
      set_error("xrc_sm_read_grads: device failure");
      s = XRC_ERR_CUDA;
    }
  }
  cudaFree(tmp);
  return s;
}

// ---------------------------------------------------------------- fused batch
int xrc_eval_batch_async(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views)
{
  XRC_CHECK_ARG(rc && sms && n_views > 0, "xrc_eval_batch: bad argument");
  for (uint32_t v = 0; v < n_views; ++v)
    XRC_CHECK_ARG(sms[v] && sms[v]->rc == rc, "xrc_eval_batch: every metric must be bound to the ray caster");
  XRC_TRY(xrc_rc_compute(rc, vol_idx));
  for (uint32_t v = 0; v < n_views; ++v)
    XRC_TRY(xrc_sm_compute(sms[v]));
  return XRC_OK;
}

int xrc_eval_batch(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views, uint32_t n_per_view,
                   float* sims_out)
{
  XRC_CHECK_ARG(sims_out, "xrc_eval_batch: null output");
  XRC_TRY(xrc_eval_batch_async(rc, vol_idx, sms, n_views));
  cudaStream_t st = rc->ctx->stream;
  for (uint32_t v = 0; v < n_views; ++v)
    XRC_CHECK_ARG(n_per_view <= sms[v]->max_imgs, "xrc_eval_batch: n_per_view exceeds metric capacity");
  // the finalize kernels also write the scalars to h_sims (host-mapped pinned memory): no D2H copy to wait for
  XRC_CUDA(cudaStreamSynchronize(st));
  for (uint32_t v = 0; v < n_views; ++v)
    memcpy(sims_out + (size_t)v * n_per_view, sms[v]->h_sims, n_per_view * sizeof(float));
  return XRC_OK;
}

// ---------------------------------------------------------------- whole objective
// ExpSO3 / ExpSE3 (lib/transforms/xregRigidUtils.cpp:40-85, xregRotUtils): Rodrigues' formula in f32
void xrc_exp_se3(const float x[6], float out[12])
{
  const float wx = x[0], wy = x[1], wz = x[2];
  const float W[9] = {0.f, -wz, wy, wz, 0.f, -wx, -wy, wx, 0.f};
  float W2[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      W2[3 * r + c] = (W[3 * r] * W[c] + W[3 * r + 1] * W[3 + c]) + W[3 * r + 2] * W[6 + c];
  const float theta = sqrtf((wx * wx + wy * wy) + wz * wz);
  float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, A[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  if (theta > 1.0e-14f)
  {
    const float th2 = theta * theta;
    const float s = sinf(theta), c1 = 1.0f - cosf(theta);
    const float ra = s / theta, rb = c1 / th2;                     // R = I + sin/theta W + (1-cos)/theta^2 W^2
    const float aa = c1 / th2, ab = (theta - s) / (theta * th2);  // A = I + (1-cos)/theta^2 W + (theta-sin)/theta^3 W^2
    for (int i = 0; i < 9; ++i)
    {
      R[i] = (R[i] + ra * W[i]) + rb * W2[i];
      A[i] = (A[i] + aa * W[i]) + ab * W2[i];
    }
  }
  for (int r = 0; r < 3; ++r)
  {
    out[4 * r] = R[3 * r];
    out[4 * r + 1] = R[3 * r + 1];
    out[4 * r + 2] = R[3 * r + 2];
    out[4 * r + 3] = (A[3 * r] * x[3] + A[3 * r + 1] * x[4]) + A[3 * r + 2] * x[5];
  }
}

// c = a * b for row-major 3x4 affine transforms
static void affine_mul(const float a[12], const float b[12], float c[12])
{
  float t[12];
  for (int r = 0; r < 3; ++r)
  {
    for (int k = 0; k < 4; ++k)
    {
      float v = (a[4 * r] * b[k] + a[4 * r + 1] * b[4 + k]) + a[4 * r + 2] * b[8 + k];
      if (k == 3)
        v += a[4 * r + 3];
      t[4 * r + k] = v;
    }
  }
  memcpy(c, t, sizeof(t));
}

// first half of the objective: size the objects for the population, hand over the poses, enqueue the ray cast and every
// view's metric on the ray caster's stream.  No synchronisation.
static int obj_fn_set_poses(xrc_rc* rc, uint32_t n_views, uint32_t n_poses, const float* cam_to_phys);

static int obj_fn_enqueue(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses,
                          const float* cam_to_phys, uint32_t n_objs = 1, const uint32_t* obj_vols = nullptr, int use_bg = -1,
                          bool drr_only = false)
{
  XRC_CHECK_ARG(rc && sms && cam_to_phys, "xrc_obj_fn: null argument");
  XRC_CHECK_ARG(rc->allocated, "xrc_obj_fn: ray caster resources not allocated");
  XRC_CHECK_ARG(n_views == rc->cams.size(), "xrc_obj_fn: need one metric per camera model / view");
  XRC_CHECK_ARG((uint64_t)n_poses * n_views <= rc->max_projs, "xrc_obj_fn: population exceeds the allocated projections");
  // Intensity2D3DRegi::setup(): view v's metric reads projections [v * pop, (v + 1) * pop)
  if (rc->num_projs != n_poses * n_views)
    XRC_TRY(xrc_rc_set_num_projs(rc, n_poses * n_views));
  for (uint32_t v = 0; v < n_views; ++v)
  {
    XRC_CHECK_ARG(sms[v] && sms[v]->rc == rc, "xrc_obj_fn: every metric must be bound to the ray caster");
    XRC_CHECK_ARG(n_poses <= sms[v]->max_imgs, "xrc_obj_fn: population exceeds the metric's capacity");
    if (sms[v]->n_imgs != n_poses)
      XRC_TRY(xrc_sm_set_num_imgs(sms[v], n_poses));
    if (sms[v]->proj_offset != v * n_poses)
      XRC_TRY(xrc_sm_bind_ray_caster(sms[v], rc, v * n_poses));
  }
  rc->ext_poses = nullptr;  // host poses take over from a caller's device buffer
  if (n_objs <= 1 && !obj_vols && use_bg < 0)
  {
    XRC_TRY(obj_fn_set_poses(rc, n_views, n_poses, cam_to_phys));
    if (drr_only)
      return xrc_rc_compute(rc, vol_idx);
    return xrc_eval_batch_async(rc, vol_idx, sms, n_views);
  }
  // several moving objects (xregIntensity2D3DRegi.cpp:594-629): the first is stored with REPLACE (on the background
  // projections when there is a static volume), the others are accumulated; the caller's settings are restored
  const int saved_store = rc->store_method;
  const bool saved_bg = rc->use_bg;
  int status = XRC_OK;
  for (uint32_t j = 0; j < n_objs && status == XRC_OK; ++j)
  {
    rc->store_method = (j == 0) ? XRC_STORE_REPLACE : XRC_STORE_ACCUM;
    rc->use_bg = (j == 0) ? ((use_bg < 0) ? saved_bg : (use_bg != 0)) : false;
    status = obj_fn_set_poses(rc, n_views, n_poses, cam_to_phys + 12 * (size_t)j * n_poses);
    if (status == XRC_OK)
      status = xrc_rc_compute(rc, obj_vols ? obj_vols[j] : vol_idx);
  }
  rc->store_method = saved_store;
  rc->use_bg = saved_bg;
  XRC_TRY(status);
  for (uint32_t v = 0; v < n_views; ++v)
    XRC_TRY(xrc_sm_compute(sms[v]));
  return XRC_OK;
}

// hand one object's population to the ray caster, replicated over the views camera-major (xregRayCastInterface.cpp:97-114)
static int obj_fn_set_poses(xrc_rc* rc, uint32_t n_views, uint32_t n_poses, const float* cam_to_phys)
{
  if (n_poses * n_views <= kInlinePoses)
  {
    // latency regime: no H2D copy, the poses ride in the kernel parameters
    XRC_TRY(use_device(rc->ctx));
    XRC_TRY(rc_wait_staging(rc));
    uint32_t g = 0;
    for (uint32_t c = 0; c < n_views; ++c)
      for (uint32_t p = 0; p < n_poses; ++p, ++g)
      {
        memcpy(rc->h_poses + 12 * (size_t)g, cam_to_phys + 12 * (size_t)p, sizeof(float) * 12);
        rc->h_cam_idx[g] = c;
      }
    rc->inline_poses = true;
    return XRC_OK;
  }
  return xrc_rc_distribute_poses(rc, n_poses, cam_to_phys);
}

// second half: wait for the stream, collect the per-view scalars: per_view[v * stride + p], p < n_poses
static int obj_fn_finish(xrc_rc* rc, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses, float* per_view, size_t stride)
{
  XRC_TRY(use_device(rc->ctx));
  // the finalize kernels also write the scalars to h_sims (host-mapped pinned memory): no D2H copy to wait for
  XRC_CUDA(cudaStreamSynchronize(rc->ctx->stream));
  for (uint32_t v = 0; v < n_views; ++v)
    memcpy(per_view + (size_t)v * stride, sms[v]->h_sims, n_poses * sizeof(float));
  return XRC_OK;
}

// ImgSimMetric2DCombineMean::compute: f32 running sum over views, then / n_views
static void combine_mean(const float* pv, uint32_t n_views, uint32_t n_poses, float* sims_out)
{
  for (uint32_t p = 0; p < n_poses; ++p)
  {
    float acc = 0.0f;
    for (uint32_t v = 0; v < n_views; ++v)
    {
      const volatile float t = acc + pv[(size_t)v * n_poses + p];
      acc = t;
    }
    sims_out[p] = acc / (float)n_views;
  }
}

int xrc_obj_fn(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses,
               const float* cam_to_phys, float* sims_out, float* per_view_out)
{
  XRC_CHECK_ARG(sims_out, "xrc_obj_fn: null output");
  if (!n_poses)
    return XRC_OK;
  XRC_TRY(obj_fn_enqueue(rc, vol_idx, sms, n_views, n_poses, cam_to_phys));
  std::vector<float> tmp;
  float* pv = per_view_out;
  if (!pv)
  {
    tmp.resize((size_t)n_views * n_poses);
    pv = tmp.data();
  }
  XRC_TRY(obj_fn_finish(rc, sms, n_views, n_poses, pv, n_poses));
  combine_mean(pv, n_views, n_poses, sims_out);
  return XRC_OK;
}

int xrc_obj_fn_objects(xrc_rc* rc, uint32_t n_objs, const uint32_t* vol_idx, xrc_sm* const* sms, uint32_t n_views,
                       uint32_t n_poses, const float* cam_to_phys, int use_bg_projs, float* sims_out, float* per_view_out)
{
  XRC_CHECK_ARG(rc && vol_idx && sims_out && n_objs > 0, "xrc_obj_fn_objects: bad argument");
  for (uint32_t j = 0; j < n_objs; ++j)
    XRC_CHECK_ARG(vol_idx[j] < rc->vols.size(), "xrc_obj_fn_objects: volume index out of range");
  XRC_CHECK_ARG(!use_bg_projs || rc->d_bg, "xrc_obj_fn_objects: no background projections set (xrc_rc_set_bg_projs)");
  if (!n_poses)
    return XRC_OK;
  XRC_TRY(obj_fn_enqueue(rc, 0, sms, n_views, n_poses, cam_to_phys, n_objs, vol_idx, use_bg_projs ? 1 : 0));
  std::vector<float> tmp;
  float* pv = per_view_out;
  if (!pv)
  {
    tmp.resize((size_t)n_views * n_poses);
    pv = tmp.data();
  }
  XRC_TRY(obj_fn_finish(rc, sms, n_views, n_poses, pv, n_poses));
  combine_mean(pv, n_views, n_poses, sims_out);
  return XRC_OK;
}

// Device d's share of the camera-major (view, pose) unit list, u = v * n_poses + p (the reference's global projection
// index, xregRayCastInterface.cpp:97-114): of units [u0, u1), view v contributes the poses [p0, p1).
static void unit_range(uint32_t u0, uint32_t u1, uint32_t v, uint32_t n_poses, uint32_t& p0, uint32_t& p1)
{
  const uint64_t lo = (uint64_t)v * n_poses, hi = lo + n_poses;
  const uint64_t a = std::max<uint64_t>(u0, lo), b = std::min<uint64_t>(u1, hi);
  p0 = p1 = 0;
  if (b > a)
  {
    p0 = (uint32_t)(a - lo);
    p1 = (uint32_t)(b - lo);
  }
}

// units [u0, u1) on one device: size the ray caster for them, bind each view's metric to its run of projections, hand
// over the poses (camera-major within the chunk) and enqueue the ray cast.  No synchronisation.
static int obj_fn_enqueue_units(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses,
                                const float* cam_to_phys, uint32_t u0, uint32_t u1)
{
  XRC_CHECK_ARG(rc && sms && cam_to_phys, "xrc_obj_fn_multi: null argument");
  XRC_CHECK_ARG(rc->allocated, "xrc_obj_fn_multi: ray caster resources not allocated");
  XRC_CHECK_ARG(n_views == rc->cams.size(), "xrc_obj_fn_multi: need one metric per camera model / view");
  const uint32_t n = u1 - u0;
  XRC_CHECK_ARG(n <= rc->max_projs, "xrc_obj_fn_multi: a device's share exceeds its allocated projections");
  if (rc->num_projs != n)
    XRC_TRY(xrc_rc_set_num_projs(rc, n));
  uint32_t off = 0;
  for (uint32_t v = 0; v < n_views; ++v)
  {
    uint32_t p0, p1;
    unit_range(u0, u1, v, n_poses, p0, p1);
    if (p1 == p0)
      continue;
    XRC_CHECK_ARG(sms[v] && sms[v]->rc == rc, "xrc_obj_fn_multi: every metric must be bound to its device's ray caster");
    XRC_CHECK_ARG(p1 - p0 <= sms[v]->max_imgs, "xrc_obj_fn_multi: a device's share exceeds a metric's capacity");
    if (sms[v]->n_imgs != p1 - p0)
      XRC_TRY(xrc_sm_set_num_imgs(sms[v], p1 - p0));
    if (sms[v]->proj_offset != off)
      XRC_TRY(xrc_sm_bind_ray_caster(sms[v], rc, off));
    off += p1 - p0;
  }
  rc->ext_poses = nullptr;
  XRC_TRY(use_device(rc->ctx));
  XRC_TRY(rc_wait_staging(rc));
  uint32_t g = 0;
  for (uint32_t v = 0; v < n_views; ++v)
  {
    uint32_t p0, p1;
    unit_range(u0, u1, v, n_poses, p0, p1);
    for (uint32_t p = p0; p < p1; ++p, ++g)
    {
      memcpy(rc->h_poses + 12 * (size_t)g, cam_to_phys + 12 * (size_t)p, sizeof(float) * 12);
      rc->h_cam_idx[g] = v;
    }
  }
  if (n <= kInlinePoses)
    rc->inline_poses = true;  // latency regime: the poses ride in the kernel parameters
  else
    XRC_TRY(rc_upload_poses(rc, n));
  return xrc_rc_compute(rc, vol_idx);
}

// chunk d of n_units units cut into n_dev contiguous balanced chunks: the first n_units % n_dev chunks take one more
static void unit_chunk(uint32_t n_units, uint32_t n_dev, uint32_t d, uint32_t& u0, uint32_t& u1)
{
  const uint32_t base = n_units / n_dev, extra = n_units % n_dev;
  u0 = d * base + std::min(d, extra);
  u1 = u0 + base + (d < extra ? 1u : 0u);
}

int xrc_obj_fn_multi_share(uint32_t n_dev, uint32_t n_views, uint32_t n_poses, uint32_t dev, uint32_t view,
                           uint32_t* first_pose, uint32_t* count)
{
  XRC_CHECK_ARG(n_dev > 0 && dev < n_dev && view < n_views && first_pose && count, "xrc_obj_fn_multi_share: bad argument");
  XRC_CHECK_ARG((uint64_t)n_views * n_poses < (1ull << 32), "xrc_obj_fn_multi_share: too many projections");
  uint32_t u0, u1, p0, p1;
  unit_chunk(n_views * n_poses, n_dev, dev, u0, u1);
  unit_range(u0, u1, view, n_poses, p0, p1);
  *first_pose = p0;
  *count = p1 - p0;
  return XRC_OK;
}

// wait for a device's chunk and scatter its scalars: unit u = v * n_poses + p goes to dst[u - dst_first_unit]
static int obj_fn_finish_units(xrc_rc* rc, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses, uint32_t u0, uint32_t u1,
                               float* dst, uint32_t dst_first_unit)
{
  XRC_TRY(use_device(rc->ctx));
  // the finalize kernels also write the scalars to h_sims (host-mapped pinned memory): no D2H copy to wait for
  XRC_CUDA(cudaStreamSynchronize(rc->ctx->stream));
  for (uint32_t v = 0; v < n_views; ++v)
  {
    uint32_t p0, p1;
    unit_range(u0, u1, v, n_poses, p0, p1);
    if (p1 > p0)
      memcpy(dst + ((size_t)v * n_poses + p0 - dst_first_unit), sms[v]->h_sims, (p1 - p0) * sizeof(float));
  }
  return XRC_OK;
}

// Tile-sharded objective, one process per GPU: every rank ray casts ITS TILES of all n_views x n_poses projections and
// writes each projection into its owner's buffer (NVLink peer stores, xrc_rc_compute_tiles) ...
int xrc_obj_fn_tiles_enqueue_drr(xrc_rc* rc, uint32_t vol_idx, uint32_t n_views, uint32_t n_poses, const float* cam_to_phys)
{
  XRC_CHECK_ARG(rc && cam_to_phys && n_views > 0, "xrc_obj_fn_tiles_enqueue_drr: bad argument");
  XRC_CHECK_ARG(rc->allocated, "xrc_obj_fn_tiles_enqueue_drr: ray caster resources not allocated");
  XRC_CHECK_ARG(n_views == rc->cams.size(), "xrc_obj_fn_tiles_enqueue_drr: need one view per camera model");
  XRC_CHECK_ARG((uint64_t)n_poses * n_views <= rc->max_projs, "xrc_obj_fn_tiles_enqueue_drr: population exceeds the allocated projections");
  if (!n_poses)
    return XRC_OK;
  if (rc->num_projs != n_poses * n_views)
    XRC_TRY(xrc_rc_set_num_projs(rc, n_poses * n_views));
  rc->ext_poses = nullptr;
  XRC_TRY(obj_fn_set_poses(rc, n_views, n_poses, cam_to_phys));
  return xrc_rc_compute_tiles(rc, vol_idx);
}

// ... and, after a barrier across the ranks on their streams, scores the units it owns: view v's metric reads the
// projections [v * n_poses + p0, v * n_poses + p1) of the unit range at their GLOBAL indices in this rank's buffer.
int xrc_obj_fn_units_enqueue_metrics(xrc_rc* rc, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses, uint32_t first_unit,
                                     uint32_t n_units)
{
  XRC_CHECK_ARG(rc && sms && n_views > 0, "xrc_obj_fn_units_enqueue_metrics: bad argument");
  XRC_CHECK_ARG((uint64_t)first_unit + n_units <= (uint64_t)n_views * n_poses, "xrc_obj_fn_units_enqueue_metrics: unit range outside the list");
  const uint32_t u0 = first_unit, u1 = first_unit + n_units;
  for (uint32_t v = 0; v < n_views; ++v)
  {
    uint32_t p0, p1;
    unit_range(u0, u1, v, n_poses, p0, p1);
    if (p1 == p0)
      continue;
    XRC_CHECK_ARG(sms[v] && sms[v]->rc == rc, "xrc_obj_fn_units_enqueue_metrics: every metric must be bound to the ray caster");
    XRC_CHECK_ARG(p1 - p0 <= sms[v]->max_imgs, "xrc_obj_fn_units_enqueue_metrics: share exceeds a metric's capacity");
    if (sms[v]->n_imgs != p1 - p0)
      XRC_TRY(xrc_sm_set_num_imgs(sms[v], p1 - p0));
    if (sms[v]->proj_offset != v * n_poses + p0)
      XRC_TRY(xrc_sm_bind_ray_caster(sms[v], rc, v * n_poses + p0));
    XRC_TRY(xrc_sm_compute(sms[v]));
  }
  return XRC_OK;
}

// ---- the rest of the tile-sharded step without NCCL: barrier and all-gather by xchg_kernel over the peer mappings ----
static void rc_fill_xchg(xrc_rc* rc, XchgArgs* x)
{
  memset(x, 0, sizeof(*x));
  x->n_ranks = rc->peer_n;
  x->rank = rc->peer_rank;
  x->epoch = ++rc->xchg_epoch;
  for (uint32_t r = 0; r < rc->peer_n; ++r)
    x->blk[r] = reinterpret_cast<unsigned char*>(rc->peer_bufs[r]) + rc->xchg_off;
  x->host_status = reinterpret_cast<uint32_t*>(rc->h_gather + rc->max_projs);
  x->timeout_ns = 30ull * 1000ull * 1000ull * 1000ull;
}

// the unit range rank r owns: the camera-major list cut into n contiguous balanced chunks (the first N % n take one more),
// the same cut as the ray-casting kernel's owner() and regi.unit_chunks
static void rank_units(uint32_t n_total, uint32_t n_ranks, uint32_t r, uint32_t& u0, uint32_t& u1)
{
  const uint32_t base = n_total / n_ranks, extra = n_total % n_ranks;
  u0 = r * base + std::min(r, extra);
  u1 = u0 + base + (r < extra ? 1u : 0u);
}

int xrc_rc_peer_barrier(xrc_rc* rc)
{
  XRC_CHECK_ARG(rc, "null ray caster");
  XRC_CHECK_ARG(rc->allocated && rc->peer_n >= 1 && rc->d_buf_own, "xrc_rc_peer_barrier: allocate and attach first");
  XRC_TRY(use_device(rc->ctx));
  XchgArgs x;
  rc_fill_xchg(rc, &x);
  return launch_xchg(x, rc->ctx->stream);
}

int xrc_obj_fn_tiles_enqueue_gather(xrc_rc* rc, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses)
{
  XRC_CHECK_ARG(rc && sms && n_views > 0, "xrc_obj_fn_tiles_enqueue_gather: bad argument");
  XRC_CHECK_ARG(rc->allocated && rc->peer_n >= 1 && rc->d_buf_own, "xrc_obj_fn_tiles_enqueue_gather: allocate and attach first");
  XRC_CHECK_ARG((uint64_t)n_views * n_poses <= rc->max_projs, "xrc_obj_fn_tiles_enqueue_gather: population exceeds the allocated projections");
  XRC_CHECK_ARG(n_views <= kXchgMaxSeg, "xrc_obj_fn_tiles_enqueue_gather: too many views");
  uint32_t u0, u1;
  rank_units(n_views * n_poses, rc->peer_n, rc->peer_rank, u0, u1);
  XRC_TRY(xrc_obj_fn_units_enqueue_metrics(rc, sms, n_views, n_poses, u0, u1 - u0));
  XRC_TRY(use_device(rc->ctx));
  XchgArgs x;
  rc_fill_xchg(rc, &x);
  for (uint32_t v = 0; v < n_views; ++v)
  {
    uint32_t p0, p1;
    unit_range(u0, u1, v, n_poses, p0, p1);
    if (p1 == p0)
      continue;
    x.seg_src[x.n_seg] = sms[v]->d_sims;
    x.seg_first[x.n_seg] = v * n_poses + p0;
    x.seg_count[x.n_seg] = p1 - p0;
    ++x.n_seg;
  }
  x.n_units_total = n_views * n_poses;
  x.host_out = rc->h_gather;
  return launch_xchg(x, rc->ctx->stream);
}

int xrc_obj_fn_tiles_finish(xrc_rc* rc, uint32_t n_views, uint32_t n_poses, float* sims_out, float* per_view_out)
{
  XRC_CHECK_ARG(rc && sims_out && n_views > 0, "xrc_obj_fn_tiles_finish: bad argument");
  XRC_CHECK_ARG(rc->allocated && rc->h_gather && (uint64_t)n_views * n_poses <= rc->max_projs, "xrc_obj_fn_tiles_finish: bad state");
  XRC_TRY(use_device(rc->ctx));
  XRC_CUDA(cudaStreamSynchronize(rc->ctx->stream));
  uint32_t* status = reinterpret_cast<uint32_t*>(rc->h_gather + rc->max_projs);
  if (*status)
  {
    *status = 0;
    XRC_FAIL(XRC_ERR_CUDA, "tile-sharded objective: a peer rank did not reach the barrier within 30 s");
  }
  if (per_view_out)
    memcpy(per_view_out, rc->h_gather, sizeof(float) * (size_t)n_views * n_poses);
  combine_mean(rc->h_gather, n_views, n_poses, sims_out);
  return XRC_OK;
}

int xrc_obj_fn_tiles(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses,
                     const float* cam_to_phys, float* sims_out, float* per_view_out)
{
  XRC_CHECK_ARG(sims_out, "xrc_obj_fn_tiles: null output");
  if (!n_poses)
    return XRC_OK;
  XRC_TRY(xrc_obj_fn_tiles_enqueue_drr(rc, vol_idx, n_views, n_poses, cam_to_phys));
  XRC_TRY(xrc_rc_peer_barrier(rc));   // every rank's tiles have landed in their owners' buffers
  XRC_TRY(xrc_obj_fn_tiles_enqueue_gather(rc, sms, n_views, n_poses));
  return xrc_obj_fn_tiles_finish(rc, n_views, n_poses, sims_out, per_view_out);
}

int xrc_obj_fn_units_enqueue(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses,
                             const float* cam_to_phys, uint32_t first_unit, uint32_t n_units)
{
  XRC_CHECK_ARG(rc && sms && cam_to_phys && n_views > 0, "xrc_obj_fn_units: bad argument");
  XRC_CHECK_ARG((uint64_t)n_views * n_poses < (1ull << 32) && (uint64_t)first_unit + n_units <= (uint64_t)n_views * n_poses,
                "xrc_obj_fn_units: unit range outside the n_views x n_poses projection list");
  if (!n_units)
    return XRC_OK;
  const uint32_t u0 = first_unit, u1 = first_unit + n_units;
  XRC_TRY(obj_fn_enqueue_units(rc, vol_idx, sms, n_views, n_poses, cam_to_phys, u0, u1));
  for (uint32_t v = 0; v < n_views; ++v)
  {
    uint32_t p0, p1;
    unit_range(u0, u1, v, n_poses, p0, p1);
    if (p1 > p0)
      XRC_TRY(xrc_sm_compute(sms[v]));
  }
  return XRC_OK;
}

int xrc_obj_fn_units(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses,
                     const float* cam_to_phys, uint32_t first_unit, uint32_t n_units, float* unit_sims_out)
{
  XRC_CHECK_ARG(unit_sims_out, "xrc_obj_fn_units: null output");
  XRC_TRY(xrc_obj_fn_units_enqueue(rc, vol_idx, sms, n_views, n_poses, cam_to_phys, first_unit, n_units));
  if (!n_units)
    return XRC_OK;
  return obj_fn_finish_units(rc, sms, n_views, n_poses, first_unit, first_unit + n_units, unit_sims_out, first_unit);
}

int xrc_obj_fn_multi(uint32_t n_dev, xrc_rc* const* rcs, xrc_sm* const* sms, uint32_t vol_idx, uint32_t n_views,
                     uint32_t n_poses, const float* cam_to_phys, float* sims_out, float* per_view_out)
{
  XRC_CHECK_ARG(n_dev > 0 && rcs && sms && cam_to_phys && sims_out && n_views > 0, "xrc_obj_fn_multi: bad argument");
  if (!n_poses)
    return XRC_OK;
  XRC_CHECK_ARG((uint64_t)n_views * n_poses < (1ull << 32), "xrc_obj_fn_multi: too many projections");
  std::vector<float> tmp;
  float* pv = per_view_out;
  if (!pv)
  {
    tmp.resize((size_t)n_views * n_poses);
    pv = tmp.data();
  }
  // the n_views x n_poses projection list (camera-major) is cut into contiguous balanced chunks, which may
  // straddle views; the first n_units % n_dev devices take one unit more.  One view: 100 poses on 8 devices -> 13 13 13 13
  // 12 12 12 12.  Three views, one pose (the BOBYQA regime of a multi-view registration): one view per device.
  std::vector<uint32_t> begin(n_dev + 1, 0);
  for (uint32_t d = 0; d < n_dev; ++d)
    unit_chunk(n_views * n_poses, n_dev, d, begin[d], begin[d + 1]);
  // enqueue everything first (asynchronous launches: the devices run concurrently), then collect.  The ray casts of
  // all devices go out before any metric kernel, so that the last device starts after n_dev launches, not 4 n_dev.
  int status = XRC_OK;
  uint32_t enqueued = 0;
  for (uint32_t d = 0; d < n_dev && status == XRC_OK; ++d, ++enqueued)
    if (begin[d + 1] > begin[d])
      status = obj_fn_enqueue_units(rcs[d], vol_idx, sms + (size_t)d * n_views, n_views, n_poses, cam_to_phys, begin[d],
                                    begin[d + 1]);
  for (uint32_t d = 0; d < n_dev && status == XRC_OK; ++d)
    for (uint32_t v = 0; v < n_views && status == XRC_OK; ++v)
    {
      uint32_t p0, p1;
      unit_range(begin[d], begin[d + 1], v, n_poses, p0, p1);
      if (p1 > p0)
        status = xrc_sm_compute(sms[(size_t)d * n_views + v]);
    }
  for (uint32_t d = 0; d < enqueued; ++d)
  {
    if (begin[d + 1] == begin[d] || !rcs[d])
      continue;
    const int s2 = obj_fn_finish_units(rcs[d], sms + (size_t)d * n_views, n_views, n_poses, begin[d], begin[d + 1], pv, 0);
    if (status == XRC_OK)
      status = s2;
  }
  XRC_TRY(status);
  combine_mean(pv, n_views, n_poses, sims_out);
  return XRC_OK;
}

// ---- SE(3) magnitude penalty (Regi2D3DPenaltyFnSE3Mag + FoldNormDist), host arithmetic in f32 in the reference's order
namespace
{
// FoldNormDist (lib/basic_math/xregFoldNormDist.cpp:30-70)
struct FoldNorm
{
  float m, s, two_s_sq, norm_const, log_norm_const;
  FoldNorm(float m_arg, float s_arg)
      : m(m_arg), s(s_arg), two_s_sq(2 * s_arg * s_arg), norm_const(std::sqrt(two_s_sq * 3.141592653589793f)),
        log_norm_const(std::log(norm_const))
  {
  }
  float exp_helper(float x) const
  {
    const float x_minus_m = x - m, x_plus_m = x + m;
    return std::exp((x_minus_m * x_minus_m) / -two_s_sq) + std::exp((x_plus_m * x_plus_m) / -two_s_sq);
  }
  float log_density(float x) const
  {
    if (x >= 0)
    {
      const float un_norm_prob = exp_helper(x);
      return (un_norm_prob > 1.0e-14f) ? (std::log(un_norm_prob) - log_norm_const) : std::numeric_limits<float>::lowest();
    }
    return -std::numeric_limits<float>::infinity();
  }
};

// ComputeRotAngTransMag (lib/transforms/xregRigidUtils.cpp:247-251) with LogSO3ToPt (xregRotUtils.cpp:107-126)
void rot_ang_trans_mag(const float T[12], float* rot, float* trans)
{
  const float theta = std::acos((T[0] + T[5] + T[10] - 1) / 2);
  float x[3] = {0.f, 0.f, 0.f};
  if (std::abs(theta) > 1.0e-14f)
  {
    x[0] = T[9] - T[6];
    x[1] = T[2] - T[8];
    x[2] = T[4] - T[1];
    const float k = theta / (2 * std::sin(theta));
    for (int i = 0; i < 3; ++i)
      x[i] *= k;
  }
  *rot = std::sqrt((x[0] * x[0] + x[1] * x[1]) + x[2] * x[2]);
  *trans = std::sqrt((T[3] * T[3] + T[7] * T[7]) + T[11] * T[11]);
}
}  // namespace

int xrc_se3_mag_penalty(const xrc_se3_penalty* pen, uint32_t n, const float* cam_wrt_obj, float* reg_vals_out)
{
  XRC_CHECK_ARG(pen && reg_vals_out && (n == 0 || cam_wrt_obj), "xrc_se3_mag_penalty: null argument");
  XRC_CHECK_ARG(pen->rot_std > 0.f && pen->trans_std > 0.f, "xrc_se3_mag_penalty: standard deviations must be positive");
  const FoldNorm rot_pdf(pen->rot_mean, pen->rot_std), trans_pdf(pen->trans_mean, pen->trans_std);
  // xregRegi2D3DPenaltyFnSE3Mag.cpp:70-77
  float inter_inv[12], init_vol_to_cam[12], init_X_to_inter[12];
  affine_inverse_f32(pen->inter_frame, inter_inv);
  affine_inverse_f32(pen->init_cam_to_vol, init_vol_to_cam);
  affine_mul(inter_inv, pen->inter_wrt_vol ? pen->init_cam_to_vol : init_vol_to_cam, init_X_to_inter);
  for (uint32_t p = 0; p < n; ++p)
  {
    // :86-103
    const float* cur = cam_wrt_obj + 12 * (size_t)p;
    float tmp[12], cur_inter_to_X[12], M[12];
    if (pen->inter_wrt_vol)
    {
      affine_inverse_f32(cur, tmp);
      affine_mul(tmp, pen->inter_frame, cur_inter_to_X);
    }
    else
      affine_mul(cur, pen->inter_frame, cur_inter_to_X);
    affine_mul(init_X_to_inter, cur_inter_to_X, M);
    float rot_err, trans_err;
    rot_ang_trans_mag(M, &rot_err, &trans_err);
    const float rot_lp = rot_pdf.log_density(rot_err), trans_lp = trans_pdf.log_density(trans_err);
    reg_vals_out[p] = 0.0f + (rot_pdf.log_norm_const - rot_lp + trans_pdf.log_norm_const - trans_lp);
  }
  return XRC_OK;
}

static void compose_se3_poses(uint32_t n_poses, const float* params, const float* pre12, const float* post12, float* poses)
{
  for (uint32_t p = 0; p < n_poses; ++p)
  {
    float* T = poses + 12 * (size_t)p;
    xrc_exp_se3(params + 6 * (size_t)p, T);
    if (pre12)
      affine_mul(pre12, T, T);
    if (post12)
      affine_mul(T, post12, T);
  }
}

int xrc_obj_fn_se3_pen(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses,
                       const float* params, const float* pre12, const float* post12, const xrc_se3_penalty* pen,
                       float* sims_out, float* per_view_out, float* penalty_out)
{
  XRC_CHECK_ARG(params && sims_out, "xrc_obj_fn_se3_pen: null argument");
  if (!pen)
    return xrc_obj_fn_se3(rc, vol_idx, sms, n_views, n_poses, params, pre12, post12, sims_out, per_view_out);
  if (!n_poses)
    return XRC_OK;
  std::vector<float> poses((size_t)n_poses * 12), reg(n_poses), tmp;
  compose_se3_poses(n_poses, params, pre12, post12, poses.data());
  // the device works on the projections while the host evaluates the regulariser of the same poses
  XRC_TRY(obj_fn_enqueue(rc, vol_idx, sms, n_views, n_poses, poses.data()));
  const int ps = xrc_se3_mag_penalty(pen, n_poses, poses.data(), reg.data());
  float* pv = per_view_out;
  if (!pv)
  {
    tmp.resize((size_t)n_views * n_poses);
    pv = tmp.data();
  }
  XRC_TRY(obj_fn_finish(rc, sms, n_views, n_poses, pv, n_poses));
  XRC_TRY(ps);
  combine_mean(pv, n_views, n_poses, sims_out);
  // Intensity2D3DRegi::obj_fn, xregIntensity2D3DRegi.cpp:653-688
  for (uint32_t p = 0; p < n_poses; ++p)
  {
    if (penalty_out)
      penalty_out[p] = reg[p];
    float sv = sims_out[p], rv = reg[p];
    if (pen->use_coeffs)
    {
      sv *= pen->img_sim_coeff;
      rv *= pen->penalty_coeff;
    }
    sims_out[p] = sv + rv;
  }
  return XRC_OK;
}

int xrc_obj_fn_se3(xrc_rc* rc, uint32_t vol_idx, xrc_sm* const* sms, uint32_t n_views, uint32_t n_poses,
                   const float* params, const float* pre12, const float* post12, float* sims_out, float* per_view_out)
{
  XRC_CHECK_ARG(params, "xrc_obj_fn_se3: null parameters");
  std::vector<float> poses((size_t)n_poses * 12);
  compose_se3_poses(n_poses, params, pre12, post12, poses.data());
  return xrc_obj_fn(rc, vol_idx, sms, n_views, n_poses, poses.data(), sims_out, per_view_out);
}

}  // extern "C"
