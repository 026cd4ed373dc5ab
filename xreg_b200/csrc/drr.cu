// Batched line-integral DRR ray casting for sm_100a.
//
// Replaces xregLineIntegralKernel (lib/ray_cast/xregRayCastLineIntOCL.cpp:39-117)
// with the arithmetic of the CPU class (lib/ray_cast/xregRayCastLineIntCPU.cpp:105-349):
// per-pixel ray / volume-box clipping in f32 with IEEE-rounded, uncontracted
// operations (so clip masks and step counts are bit-identical to the CPU path),
// then fixed-step marching with manual trilinear interpolation.
//
// Design notes (DESIGN.md has the numbers):
//  * one thread per detector pixel, one warp = 8x4 pixel patch, one CTA = 16x16
//    pixels of ONE projection; CTAs of the same detector tile for all poses of
//    the population are adjacent in launch order so that their beams share L2.
//  * the volume is repacked once at set_volumes() time into a layout whose
//    per-sample fetch is 2 x 128-bit loads (XY-quad records) instead of 8
//    scattered 32-bit loads; other layouts are kept for measurement.
//  * no tensor cores: the path is a gather + lerp, not a contraction.
#include "common.h"

namespace xrc
{

// ----------------------------------------------------------------------------
// exact f32 helpers: never contracted into FMAs, IEEE division / sqrt
// ----------------------------------------------------------------------------
__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float fsqrt(float a) { return __fsqrt_rn(a); }

__device__ __forceinline__ float dot3(float a0, float a1, float a2, float b0, float b1, float b2)
{
  return fadd(fadd(fmul(a0, b0), fmul(a1, b1)), fmul(a2, b2));
}

__device__ __forceinline__ float norm3(float x, float y, float z)
{
  return fsqrt(fadd(fadd(fmul(x, x), fmul(y, y)), fmul(z, z)));
}

struct Ray
{
  bool hit;
  float x, y, z;     // first sample (continuous index)
  float sx, sy, sz;  // step vector
  uint32_t nsamples; // num_steps + 1
};

// Per-projection constants, computed once per CTA into shared memory:
// X = phys_to_idx o pose (xregRayCastLineIntCPU.cpp:207-208), p = X * pinhole (:211)
struct ProjConst
{
  float X[12];
  float p[3];
};

__device__ __forceinline__ void compute_proj_const(const DrrArgs& a, const xrc_cam& cam,
                                                   const float* __restrict__ pose, ProjConst* pc)
{
  // executed by threads 0..11 then 0..2 of the CTA (see callers)
  const int t = threadIdx.x;
  if (t < 12)
  {
    const int r = t >> 2, c = t & 3;
    const float* A = a.phys_to_idx;
    float v = dot3(A[4 * r], A[4 * r + 1], A[4 * r + 2], pose[c], pose[4 + c], pose[8 + c]);
    if (c == 3)
      v = fadd(v, A[4 * r + 3]);
    pc->X[t] = v;
  }
  __syncthreads();
  if (t < 3)
  {
    const float* X = pc->X;
    pc->p[t] = fadd(dot3(X[4 * t], X[4 * t + 1], X[4 * t + 2], cam.pinhole[0], cam.pinhole[1], cam.pinhole[2]),
                    X[4 * t + 3]);
  }
  __syncthreads();
}

// xregRayCastLineIntCPU.cpp:176-268 for one pixel
__device__ __forceinline__ Ray setup_ray(const xrc_cam& cam, const ProjConst& pc, float step_size,
                                         int nx, int ny, int nz, uint32_t row, uint32_t col)
{
  Ray ray;
  ray.hit = false;
  ray.nsamples = 0;
  ray.x = ray.y = ray.z = ray.sx = ray.sy = ray.sz = 0.f;

  // CameraModel::ind_pt_to_phys_det_pt (xregPerspectiveXform.cpp:391-414)
  const float det_z = fmul((cam.frame_type == 1) ? -1.0f : 1.0f, cam.focal_len);
  const float i0 = fmul(det_z, (float)col), i1 = fmul(det_z, (float)row), i2 = fmul(det_z, 1.0f);
  const float* Ki = cam.intrins_inv;
  float c0 = dot3(Ki[0], Ki[1], Ki[2], i0, i1, i2);
  float c1 = dot3(Ki[3], Ki[4], Ki[5], i0, i1, i2);
  float c2 = dot3(Ki[6], Ki[7], Ki[8], i0, i1, i2);
  if (cam.frame_type == 2)
  {
    c0 = fadd(c0, 0.0f);
    c1 = fadd(c1, 0.0f);
    c2 = fadd(c2, -cam.focal_len);
  }
  const float* E = cam.extrins_inv;
  const float d0 = fadd(dot3(E[0], E[1], E[2], c0, c1, c2), E[3]);
  const float d1 = fadd(dot3(E[4], E[5], E[6], c0, c1, c2), E[7]);
  const float d2 = fadd(dot3(E[8], E[9], E[10], c0, c1, c2), E[11]);

  const float* X = pc.X;
  const float px = pc.p[0], py = pc.p[1], pz = pc.p[2];
  const float dx = fsub(fadd(dot3(X[0], X[1], X[2], d0, d1, d2), X[3]), px);
  const float dy = fsub(fadd(dot3(X[4], X[5], X[6], d0, d1, d2), X[7]), py);
  const float dz = fsub(fadd(dot3(X[8], X[9], X[10], d0, d1, d2), X[11]), pz);

  // RayRectIntersect, limit_to_segment = true (xregSpatialPrimitives.cpp:175-222)
  float t0 = 0.f, t1 = 1.f;
  bool hit = true;
  const float pp[3] = {px, py, pz};
  const float dd[3] = {dx, dy, dz};
  const float mx[3] = {(float)(nx - 1), (float)(ny - 1), (float)(nz - 1)};
#pragma unroll
  for (int k = 0; k < 3; ++k)
  {
    if (hit)
    {
      if (fabsf(dd[k]) > 1.0e-8f)
      {
        const float inv = fdiv(1.0f, dd[k]);
        float ta = fmul(fsub(0.0f, pp[k]), inv);
        float tb = fmul(fsub(mx[k], pp[k]), inv);
        if (tb < ta)
        {
          const float tmp = ta;
          ta = tb;
          tb = tmp;
        }
        t0 = (t0 < ta) ? ta : t0;
        t1 = (tb < t1) ? tb : t1;
        if (t0 > t1)
          hit = false;
      }
      else if ((pp[k] < 0.0f) || (pp[k] > mx[k]))
      {
        hit = false;
      }
    }
  }

  const float tol = 1.0e-3f;  // kVOL_BB_STEP_INC_TOL, xregRayCastBaseCPU.h:37
  if (hit && (fsub(t1, t0) > fmul(2.0f, tol)))
  {
    t0 = fadd(t0, tol);
    t1 = fsub(t1, tol);
    ray.x = fadd(px, fmul(t0, dx));
    ray.y = fadd(py, fmul(t0, dy));
    ray.z = fadd(pz, fmul(t0, dz));
    const float L = norm3(dx, dy, dz);
    const float len = fmul(fsub(t1, t0), L);
    float ux = fsub(d0, cam.pinhole[0]), uy = fsub(d1, cam.pinhole[1]), uz = fsub(d2, cam.pinhole[2]);
    const float dn = norm3(ux, uy, uz);
    ux = fmul(fdiv(ux, dn), step_size);
    uy = fmul(fdiv(uy, dn), step_size);
    uz = fmul(fdiv(uz, dn), step_size);
    const float s0 = dot3(X[0], X[1], X[2], ux, uy, uz);
    const float s1 = dot3(X[4], X[5], X[6], ux, uy, uz);
    const float s2 = dot3(X[8], X[9], X[10], ux, uy, uz);
    const float step_len = norm3(s0, s1, s2);
    const unsigned long long num_steps = __float2ull_rz(fdiv(len, step_len));
    const float scale = fdiv(step_len, L);
    ray.sx = fmul(dx, scale);
    ray.sy = fmul(dy, scale);
    ray.sz = fmul(dz, scale);
    ray.nsamples = (uint32_t)((num_steps > 0xFFFFFFFEull) ? 0xFFFFFFFEull : num_steps) + 1u;
    ray.hit = true;
  }
  return ray;
}

// ----------------------------------------------------------------------------
// trilinear sample for the different HBM layouts.  All produce
//   vx00 + ... lerps in f32 with the ITK weight / neighbour rules (clamping the
// coordinate to [0, n-1] reproduces ITK's start-index clamp, "distance <= 0" and
// "neighbour beyond end index" branches exactly; DESIGN.md has the argument).
// floor() uses the round-down add trick (full-rate FADD.RM instead of F2I/I2F).
// ----------------------------------------------------------------------------
struct Cell
{
  int ix, iy, iz;
  float wx, wy, wz;
};

__device__ __forceinline__ void split_coord(float x, float hi, int& i, float& w)
{
  const float c = fminf(fmaxf(x, 0.0f), hi);
  const float t = __fadd_rd(c, 8388608.0f);  // 2^23: mantissa now holds floor(c)
  const float b = t - 8388608.0f;            // exact
  w = c - b;                                 // exact (Sterbenz-like: same binade or below)
  i = __float_as_int(t) - 0x4B000000;
}

__device__ __forceinline__ float lerp(float a, float b, float w) { return fmaf(w, b - a, a); }

__device__ __forceinline__ float trilerp(float v000, float v100, float v010, float v110, float v001,
                                         float v101, float v011, float v111, float wx, float wy, float wz)
{
  const float vx00 = lerp(v000, v100, wx);
  const float vx10 = lerp(v010, v110, wx);
  const float vx01 = lerp(v001, v101, wx);
  const float vx11 = lerp(v011, v111, wx);
  const float vxx0 = lerp(vx00, vx10, wy);
  const float vxx1 = lerp(vx01, vx11, wy);
  return lerp(vxx0, vxx1, wz);
}

template <int LAYOUT>
struct Sampler;

// padded linear volume: (nx+1) x (ny+1) x (nz+1), edge replicated
template <>
struct Sampler<XRC_LAYOUT_LINEAR>
{
  const float* __restrict__ v;
  int sy, sz;
  float hx, hy, hz;
  __device__ Sampler(const DrrArgs& a)
      : v((const float*)a.vol), sy(a.nx + 1), sz((a.nx + 1) * (a.ny + 1)), hx((float)(a.nx - 1)),
        hy((float)(a.ny - 1)), hz((float)(a.nz - 1))
  {
  }
  __device__ __forceinline__ float operator()(float x, float y, float z) const
  {
    int ix, iy, iz;
    float wx, wy, wz;
    split_coord(x, hx, ix, wx);
    split_coord(y, hy, iy, wy);
    split_coord(z, hz, iz, wz);
    const float* p = v + ((size_t)iz * sz + (size_t)iy * sy + ix);
    const float v000 = __ldg(p), v100 = __ldg(p + 1);
    const float v010 = __ldg(p + sy), v110 = __ldg(p + sy + 1);
    const float v001 = __ldg(p + sz), v101 = __ldg(p + sz + 1);
    const float v011 = __ldg(p + sz + sy), v111 = __ldg(p + sz + sy + 1);
    return trilerp(v000, v100, v010, v110, v001, v101, v011, v111, wx, wy, wz);
  }
};

// XY-quad records: float4 {v(x,y), v(x+1,y), v(x,y+1), v(x+1,y+1)} per voxel,
// nx x ny x (nz+1) records (last plane replicated)
template <>
struct Sampler<XRC_LAYOUT_QUAD>
{
  const float4* __restrict__ v;
  int sy, sz;
  float hx, hy, hz;
  __device__ Sampler(const DrrArgs& a)
      : v((const float4*)a.vol), sy(a.nx), sz(a.nx * a.ny), hx((float)(a.nx - 1)), hy((float)(a.ny - 1)),
        hz((float)(a.nz - 1))
  {
  }
  __device__ __forceinline__ float operator()(float x, float y, float z) const
  {
    int ix, iy, iz;
    float wx, wy, wz;
    split_coord(x, hx, ix, wx);
    split_coord(y, hy, iy, wy);
    split_coord(z, hz, iz, wz);
    const float4* p = v + ((size_t)iz * sz + (size_t)iy * sy + ix);
    const float4 q0 = __ldg(p);
    const float4 q1 = __ldg(p + sz);
    return trilerp(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, wx, wy, wz);
  }
};

// full 2x2x2 corner records: 2 x float4 per voxel (32-byte sector aligned)
template <>
struct Sampler<XRC_LAYOUT_OCT>
{
  const float4* __restrict__ v;
  int sy, sz;
  float hx, hy, hz;
  __device__ Sampler(const DrrArgs& a)
      : v((const float4*)a.vol), sy(a.nx), sz(a.nx * a.ny), hx((float)(a.nx - 1)), hy((float)(a.ny - 1)),
        hz((float)(a.nz - 1))
  {
  }
  __device__ __forceinline__ float operator()(float x, float y, float z) const
  {
    int ix, iy, iz;
    float wx, wy, wz;
    split_coord(x, hx, ix, wx);
    split_coord(y, hy, iy, wy);
    split_coord(z, hz, iz, wz);
    const float4* p = v + 2 * ((size_t)iz * sz + (size_t)iy * sy + ix);
    const float4 q0 = __ldg(p);
    const float4 q1 = __ldg(p + 1);
    return trilerp(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, wx, wy, wz);
  }
};

// 3D float texture, point sampled, clamp addressing: 8 fetches
template <>
struct Sampler<XRC_LAYOUT_TEX>
{
  cudaTextureObject_t t;
  float hx, hy, hz;
  __device__ Sampler(const DrrArgs& a)
      : t(a.tex), hx((float)(a.nx - 1)), hy((float)(a.ny - 1)), hz((float)(a.nz - 1))
  {
  }
  __device__ __forceinline__ float operator()(float x, float y, float z) const
  {
    int ix, iy, iz;
    float wx, wy, wz;
    split_coord(x, hx, ix, wx);
    split_coord(y, hy, iy, wy);
    split_coord(z, hz, iz, wz);
    const float fx = (float)ix + 0.5f, fy = (float)iy + 0.5f, fz = (float)iz + 0.5f;
    const float v000 = tex3D<float>(t, fx, fy, fz), v100 = tex3D<float>(t, fx + 1.f, fy, fz);
    const float v010 = tex3D<float>(t, fx, fy + 1.f, fz), v110 = tex3D<float>(t, fx + 1.f, fy + 1.f, fz);
    const float v001 = tex3D<float>(t, fx, fy, fz + 1.f), v101 = tex3D<float>(t, fx + 1.f, fy, fz + 1.f);
    const float v011 = tex3D<float>(t, fx, fy + 1.f, fz + 1.f), v111 = tex3D<float>(t, fx + 1.f, fy + 1.f, fz + 1.f);
    return trilerp(v000, v100, v010, v110, v001, v101, v011, v111, wx, wy, wz);
  }
};

// 3D float4 texture of XY-quad records, point sampled: 2 fetches
template <>
struct Sampler<XRC_LAYOUT_TEX_QUAD>
{
  cudaTextureObject_t t;
  float hx, hy, hz;
  __device__ Sampler(const DrrArgs& a)
      : t(a.tex), hx((float)(a.nx - 1)), hy((float)(a.ny - 1)), hz((float)(a.nz - 1))
  {
  }
  __device__ __forceinline__ float operator()(float x, float y, float z) const
  {
    int ix, iy, iz;
    float wx, wy, wz;
    split_coord(x, hx, ix, wx);
    split_coord(y, hy, iy, wy);
    split_coord(z, hz, iz, wz);
    const float fx = (float)ix + 0.5f, fy = (float)iy + 0.5f, fz = (float)iz + 0.5f;
    const float4 q0 = tex3D<float4>(t, fx, fy, fz);
    const float4 q1 = tex3D<float4>(t, fx, fy, fz + 1.f);
    return trilerp(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, wx, wy, wz);
  }
};

// ----------------------------------------------------------------------------
// main kernel
// ----------------------------------------------------------------------------
constexpr int kTileW = 16;
constexpr int kTileH = 16;
constexpr int kThreads = kTileW * kTileH;

__device__ __forceinline__ void cta_coords(const DrrArgs& a, uint32_t& proj, uint32_t& tile)
{
  const uint32_t b = blockIdx.x;
  if (a.order == 0)
  {
    proj = b % a.n_projs;
    tile = b / a.n_projs;
  }
  else
  {
    const uint32_t nt = a.tiles_x * a.tiles_y;
    tile = b % nt;
    proj = b / nt;
  }
}

__device__ __forceinline__ void thread_pixel(const DrrArgs& a, uint32_t tile, uint32_t& row, uint32_t& col)
{
  // warp = 8 (cols) x 4 (rows) pixel patch; 2 x 4 warps per 16 x 16 tile
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tx = tile % a.tiles_x, ty = tile / a.tiles_x;
  col = tx * kTileW + (warp & 1) * 8 + (lane & 7);
  row = ty * kTileH + (warp >> 1) * 4 + (lane >> 3);
}

template <int LAYOUT, int KERNEL_ID>
__global__ void __launch_bounds__(kThreads) drr_kernel(const DrrArgs a)
{
  __shared__ ProjConst pc;
  __shared__ xrc_cam cam_s;
  __shared__ unsigned long long cta_samples;

  uint32_t proj, tile;
  cta_coords(a, proj, tile);

  const uint32_t ci = a.cam_idx[proj];
  {
    // stage the camera (25 words) in shared memory
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.cams + ci);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&cam_s);
    if (threadIdx.x < sizeof(xrc_cam) / 4)
      dst[threadIdx.x] = src[threadIdx.x];
    if (threadIdx.x == 0)
      cta_samples = 0ull;
  }
  __syncthreads();
  compute_proj_const(a, cam_s, a.poses + 12 * (size_t)proj, &pc);

  uint32_t row, col;
  thread_pixel(a, tile, row, col);
  const bool in_img = (row < a.rows) && (col < a.cols);

  Ray ray;
  ray.hit = false;
  ray.nsamples = 0;
  if (in_img)
    ray = setup_ray(cam_s, pc, a.step_size, a.nx, a.ny, a.nz, row, col);

  float sum = (KERNEL_ID == XRC_KERNEL_MAX) ? -3.402823466e+38f : 0.0f;
  if (ray.hit)
  {
    const Sampler<LAYOUT> smp(a);
    float x = ray.x, y = ray.y, z = ray.z;
    const float sx = ray.sx, sy = ray.sy, sz = ray.sz;
    const uint32_t n = ray.nsamples;
#pragma unroll 4
    for (uint32_t s = 0; s < n; ++s)
    {
      const float v = smp(x, y, z);
      if (KERNEL_ID == XRC_KERNEL_MAX)
        sum = fmaxf(sum, v);
      else
        sum = fadd(sum, v);
      x = fadd(x, sx);
      y = fadd(y, sy);
      z = fadd(z, sz);
    }
    sum = fmul(sum, a.step_size);  // xregRayCastLineIntCPU.cpp:279
  }

  if (in_img)
  {
    const size_t npix = (size_t)a.rows * a.cols;
    const size_t o = (size_t)proj * npix + (size_t)row * a.cols + col;
    float base;
    if (a.init_mode == 0)
      base = a.default_bg;
    else if (a.init_mode == 1)
      base = __ldg(a.bg + (size_t)ci * npix + (size_t)row * a.cols + col);
    else
      base = a.out[o];
    const float aa = fadd(0.0f, fmul(sum, 1.0f));  // :282
    a.out[o] = (KERNEL_ID == XRC_KERNEL_MAX) ? fmaxf(base, aa) : fadd(base, aa);  // :285
  }

  if (a.sample_counter)
  {
    unsigned long long n = ray.nsamples;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0)
      atomicAdd(&cta_samples, n);
    __syncthreads();
    if (threadIdx.x == 0)
      atomicAdd(a.sample_counter, cta_samples);
  }
}

// same ray set-up, emits the clip mask and sample counts (parity instrumentation)
__global__ void __launch_bounds__(kThreads) ray_info_kernel(const DrrArgs a)
{
  __shared__ ProjConst pc;
  __shared__ xrc_cam cam_s;
  uint32_t proj, tile;
  cta_coords(a, proj, tile);
  const uint32_t ci = a.cam_idx[proj];
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.cams + ci);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&cam_s);
    if (threadIdx.x < sizeof(xrc_cam) / 4)
      dst[threadIdx.x] = src[threadIdx.x];
  }
  __syncthreads();
  compute_proj_const(a, cam_s, a.poses + 12 * (size_t)proj, &pc);
  uint32_t row, col;
  thread_pixel(a, tile, row, col);
  if ((row < a.rows) && (col < a.cols))
  {
    const Ray ray = setup_ray(cam_s, pc, a.step_size, a.nx, a.ny, a.nz, row, col);
    const size_t o = (size_t)proj * a.rows * a.cols + (size_t)row * a.cols + col;
    if (a.ray_mask)
      a.ray_mask[o] = ray.hit ? 1 : 0;
    if (a.ray_steps)
      a.ray_steps[o] = ray.nsamples;
    if (a.sample_counter && ray.nsamples)
      atomicAdd(a.sample_counter, (unsigned long long)ray.nsamples);
  }
}

template <int LAYOUT>
static int launch_layout(const DrrArgs& a, int kernel_id, cudaStream_t st)
{
  const uint32_t nblocks = a.n_projs * a.tiles_x * a.tiles_y;
  if (kernel_id == XRC_KERNEL_SUM)
    drr_kernel<LAYOUT, XRC_KERNEL_SUM><<<nblocks, kThreads, 0, st>>>(a);
  else
    drr_kernel<LAYOUT, XRC_KERNEL_MAX><<<nblocks, kThreads, 0, st>>>(a);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

int launch_drr(const DrrArgs& a_in, int layout, int kernel_id, cudaStream_t st)
{
  DrrArgs a = a_in;
  a.tiles_x = (a.cols + kTileW - 1) / kTileW;
  a.tiles_y = (a.rows + kTileH - 1) / kTileH;
  if (!a.n_projs)
    return XRC_OK;
  switch (layout)
  {
    case XRC_LAYOUT_LINEAR: return launch_layout<XRC_LAYOUT_LINEAR>(a, kernel_id, st);
    case XRC_LAYOUT_QUAD: return launch_layout<XRC_LAYOUT_QUAD>(a, kernel_id, st);
    case XRC_LAYOUT_OCT: return launch_layout<XRC_LAYOUT_OCT>(a, kernel_id, st);
    case XRC_LAYOUT_TEX: return launch_layout<XRC_LAYOUT_TEX>(a, kernel_id, st);
    case XRC_LAYOUT_TEX_QUAD: return launch_layout<XRC_LAYOUT_TEX_QUAD>(a, kernel_id, st);
    default: XRC_FAIL(XRC_ERR_INVALID, "unknown volume layout");
  }
}

int launch_ray_info(const DrrArgs& a_in, cudaStream_t st)
{
  DrrArgs a = a_in;
  a.tiles_x = (a.cols + kTileW - 1) / kTileW;
  a.tiles_y = (a.rows + kTileH - 1) / kTileH;
  if (!a.n_projs)
    return XRC_OK;
  ray_info_kernel<<<a.n_projs * a.tiles_x * a.tiles_y, kThreads, 0, st>>>(a);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

// ----------------------------------------------------------------------------
// volume repacking (once per set_volumes)
// ----------------------------------------------------------------------------
__global__ void repack_linear_kernel(const float* __restrict__ src, float* __restrict__ dst, int nx, int ny, int nz)
{
  const size_t n = (size_t)(nx + 1) * (ny + 1) * (nz + 1);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const int x = (int)(i % (nx + 1));
    const int y = (int)((i / (nx + 1)) % (ny + 1));
    const int z = (int)(i / ((size_t)(nx + 1) * (ny + 1)));
    dst[i] = src[((size_t)min(z, nz - 1) * ny + min(y, ny - 1)) * nx + min(x, nx - 1)];
  }
}

__global__ void repack_quad_kernel(const float* __restrict__ src, float4* __restrict__ dst, int nx, int ny, int nz,
                                   int nz_out)
{
  const size_t n = (size_t)nx * ny * nz_out;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const int x = (int)(i % nx);
    const int y = (int)((i / nx) % ny);
    const int z = min((int)(i / ((size_t)nx * ny)), nz - 1);
    const int x1 = min(x + 1, nx - 1), y1 = min(y + 1, ny - 1);
    const float* p = src + (size_t)z * nx * ny;
    dst[i] = make_float4(p[(size_t)y * nx + x], p[(size_t)y * nx + x1], p[(size_t)y1 * nx + x], p[(size_t)y1 * nx + x1]);
  }
}

__global__ void repack_oct_kernel(const float* __restrict__ src, float4* __restrict__ dst, int nx, int ny, int nz)
{
  const size_t n = (size_t)nx * ny * nz;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const int x = (int)(i % nx);
    const int y = (int)((i / nx) % ny);
    const int z = (int)(i / ((size_t)nx * ny));
    const int x1 = min(x + 1, nx - 1), y1 = min(y + 1, ny - 1), z1 = min(z + 1, nz - 1);
    const float* p = src + (size_t)z * nx * ny;
    const float* q = src + (size_t)z1 * nx * ny;
    dst[2 * i] = make_float4(p[(size_t)y * nx + x], p[(size_t)y * nx + x1], p[(size_t)y1 * nx + x], p[(size_t)y1 * nx + x1]);
    dst[2 * i + 1] = make_float4(q[(size_t)y * nx + x], q[(size_t)y * nx + x1], q[(size_t)y1 * nx + x], q[(size_t)y1 * nx + x1]);
  }
}

void free_volume(DeviceVolume* v)
{
  if (v->tex)
    cudaDestroyTextureObject(v->tex);
  if (v->array)
    cudaFreeArray(v->array);
  if (v->data)
    cudaFree(v->data);
  v->tex = 0;
  v->array = nullptr;
  v->data = nullptr;
  v->bytes = 0;
}

static int make_texture(DeviceVolume* v, const void* d_src, bool quad, cudaStream_t st)
{
  const size_t nx = v->dims[0], ny = v->dims[1], nz = v->dims[2];
  cudaChannelFormatDesc desc = quad ? cudaCreateChannelDesc<float4>() : cudaCreateChannelDesc<float>();
  const size_t esz = quad ? sizeof(float4) : sizeof(float);
  XRC_CUDA(cudaMalloc3DArray(&v->array, &desc, make_cudaExtent(nx, ny, nz)));
  cudaMemcpy3DParms cp = {};
  cp.srcPtr = make_cudaPitchedPtr(const_cast<void*>(d_src), nx * esz, nx, ny);
  cp.dstArray = v->array;
  cp.extent = make_cudaExtent(nx, ny, nz);
  cp.kind = cudaMemcpyDeviceToDevice;
  XRC_CUDA(cudaMemcpy3DAsync(&cp, st));
  cudaResourceDesc rd = {};
  rd.resType = cudaResourceTypeArray;
  rd.res.array.array = v->array;
  cudaTextureDesc td = {};
  td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeClamp;
  td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeElementType;
  td.normalizedCoords = 0;
  XRC_CUDA(cudaCreateTextureObject(&v->tex, &rd, &td, nullptr));
  v->bytes = nx * ny * nz * esz;
  return XRC_OK;
}

int repack_volume(const float* d_linear, DeviceVolume* v, int layout, cudaStream_t st)
{
  const int nx = (int)v->dims[0], ny = (int)v->dims[1], nz = (int)v->dims[2];
  const int grid = 148 * 8, block = 256;
  v->layout = layout;
  switch (layout)
  {
    case XRC_LAYOUT_LINEAR:
    {
      v->bytes = sizeof(float) * (size_t)(nx + 1) * (ny + 1) * (nz + 1);
      XRC_CUDA(cudaMalloc(&v->data, v->bytes));
      repack_linear_kernel<<<grid, block, 0, st>>>(d_linear, (float*)v->data, nx, ny, nz);
      count_launch();
      break;
    }
    case XRC_LAYOUT_QUAD:
    {
      v->bytes = sizeof(float4) * (size_t)nx * ny * (nz + 1);
      XRC_CUDA(cudaMalloc(&v->data, v->bytes));
      repack_quad_kernel<<<grid, block, 0, st>>>(d_linear, (float4*)v->data, nx, ny, nz, nz + 1);
      count_launch();
      break;
    }
    case XRC_LAYOUT_OCT:
    {
      v->bytes = 2 * sizeof(float4) * (size_t)nx * ny * nz;
      XRC_CUDA(cudaMalloc(&v->data, v->bytes));
      repack_oct_kernel<<<grid, block, 0, st>>>(d_linear, (float4*)v->data, nx, ny, nz);
      count_launch();
      break;
    }
    case XRC_LAYOUT_TEX:
      XRC_TRY(make_texture(v, d_linear, false, st));
      break;
    case XRC_LAYOUT_TEX_QUAD:
    {
      float4* tmp = nullptr;
      XRC_CUDA(cudaMalloc(&tmp, sizeof(float4) * (size_t)nx * ny * nz));
      repack_quad_kernel<<<grid, block, 0, st>>>(d_linear, tmp, nx, ny, nz, nz);
      count_launch();
      const int s = make_texture(v, tmp, true, st);
      cudaStreamSynchronize(st);
      cudaFree(tmp);
      XRC_TRY(s);
      break;
    }
    default: XRC_FAIL(XRC_ERR_INVALID, "unknown volume layout");
  }
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

}  // namespace xrc
