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
#include <algorithm>

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

__device__ __forceinline__ uint32_t proj_cam_index(const DrrArgs& a, uint32_t proj)
{
  // host-supplied indices are validated by the ABI; a caller's device array (xrc_rc_set_poses_device) is not, so clamp
  return min(a.use_inline ? a.inl_cam[proj] : a.cam_idx[proj], a.n_cams - 1u);
}

__device__ __forceinline__ void compute_proj_const(const DrrArgs& a, const xrc_cam& cam, uint32_t proj, ProjConst* pc)
{
  // executed by threads 0..11 then 0..2 of the CTA (see callers)
  const int t = threadIdx.x;
  if (t < 12)
  {
    const int r = t >> 2, c = t & 3;
    const float* A = a.phys_to_idx;
    float p0, p1, p2;  // column c of the pose
    if (a.use_inline)
    {
      p0 = a.inl_poses[12 * proj + c], p1 = a.inl_poses[12 * proj + 4 + c], p2 = a.inl_poses[12 * proj + 8 + c];
    }
    else
    {
      const float* __restrict__ pose = a.poses + 12 * (size_t)proj;
      p0 = pose[c], p1 = pose[4 + c], p2 = pose[8 + c];
    }
    float v = dot3(A[4 * r], A[4 * r + 1], A[4 * r + 2], p0, p1, p2);
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
                                         int nx, int ny, int nz, uint32_t row, uint32_t col, bool limit_to_segment = true)
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

  // RayRectIntersect (xregSpatialPrimitives.cpp:175-222); the line integral limits the ray to the source-detector segment,
  // the depth ray caster does not (xregRayCastDepthCPU.cpp:118-121)
  float t0 = 0.f, t1 = limit_to_segment ? 1.f : __int_as_float(0x7f800000);
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

// Nearest-neighbour interpolation (RayCaster::kRAY_CAST_INTERP_NN, xregRayCastLineIntCPU.cpp:128-130): the voxel at
// ITK's ConvertContinuousIndexToNearestIndex = floor(x + 0.5) per axis (evaluated exactly, in double), clamped to the
// volume.  No arithmetic on the values: sums are bit-identical to the oracle's.  Reads the padded f32 copy or the first
// component of a record (every record layout keeps v(ix, iy, iz) there unchanged), whichever payload the volume has.
template <>
struct Sampler<kLayoutNN>
{
  const float* __restrict__ v;
  size_t s0, s1, s2, off;
  int hx, hy, hz;
  __device__ Sampler(const DrrArgs& a)
      : v(a.nn_base), s0(a.nn_s[0]), s1(a.nn_s[1]), s2(a.nn_s[2]), off(a.nn_off), hx(a.nx - 1), hy(a.ny - 1), hz(a.nz - 1)
  {
  }
  __device__ __forceinline__ float operator()(float x, float y, float z) const
  {
    const int ix = min(max((int)floor((double)x + 0.5), 0), hx);
    const int iy = min(max((int)floor((double)y + 0.5), 0), hy);
    const int iz = min(max((int)floor((double)z + 0.5), 0), hz);
    return __ldg(v + (off + (size_t)ix * s0 + (size_t)iy * s1 + (size_t)iz * s2));
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
  uint32_t t;
  if (a.order == 0)
  {
    proj = b % a.n_projs;
    t = b / a.n_projs;
  }
  else
  {
    t = b % a.tile_count;
    proj = b / a.tile_count;
  }
  tile = a.tile_first + t * a.tile_stride;
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

  const uint32_t ci = proj_cam_index(a, proj);
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
  compute_proj_const(a, cam_s, proj, &pc);

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


// ----------------------------------------------------------------------------
// Depth ray caster (RayCasterDepthCPU, lib/ray_cast/xregRayCastDepthCPU.cpp:42-272): per ray the
// first sample whose interpolated value reaches the collision threshold, refined by step halvings, as the distance of
// that point from the pinhole in the camera frame; stored with min.  Every decision (value >= threshold) must be the
// CPU class's, so the interpolation here is ITK's arithmetic itself -- f64 lerps of the f32 corners with f32 weights,
// uncontracted -- read from whatever payload the volume has (like the nearest-neighbour sampler).  Not a throughput
// kernel: a depth image is rendered once in a while, not per optimiser iteration.
// ----------------------------------------------------------------------------
struct VoxelReader
{
  const float* __restrict__ v;
  size_t s0, s1, s2, off;
  int hx, hy, hz;
  __device__ VoxelReader(const DrrArgs& a)
      : v(a.nn_base), s0(a.nn_s[0]), s1(a.nn_s[1]), s2(a.nn_s[2]), off(a.nn_off), hx(a.nx - 1), hy(a.ny - 1), hz(a.nz - 1)
  {
  }
  __device__ __forceinline__ double at(int ix, int iy, int iz) const
  {
    return (double)__ldg(v + (off + (size_t)ix * s0 + (size_t)iy * s1 + (size_t)iz * s2));
  }
  __device__ __forceinline__ float nearest(float x, float y, float z) const
  {
    const int ix = min(max((int)floor((double)x + 0.5), 0), hx);
    const int iy = min(max((int)floor((double)y + 0.5), 0), hy);
    const int iz = min(max((int)floor((double)z + 0.5), 0), hz);
    return (float)at(ix, iy, iz);
  }
  // itk::LinearInterpolateImageFunction::EvaluateOptimized(Dispatch<3>) as the oracle restates it (xo_interp_linear)
  __device__ __forceinline__ float linear(float x, float y, float z) const
  {
    const float xs[3] = {x, y, z};
    const int h[3] = {hx, hy, hz};
    int b[3], n[3];
    double d[3];
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
      int bk = (int)floorf(xs[k]);
      bk = min(max(bk, 0), h[k]);
      float dist = __fsub_rn(xs[k], (float)bk);
      int nk = bk + 1;
      if (dist <= 0.0f)
      {
        dist = 0.0f;
        nk = bk;
      }
      if (nk > h[k])
      {
        nk = bk;
        dist = 0.0f;
      }
      b[k] = bk;
      n[k] = nk;
      d[k] = (double)dist;
    }
    const double v000 = at(b[0], b[1], b[2]), v100 = at(n[0], b[1], b[2]);
    const double v010 = at(b[0], n[1], b[2]), v110 = at(n[0], n[1], b[2]);
    const double v001 = at(b[0], b[1], n[2]), v101 = at(n[0], b[1], n[2]);
    const double v011 = at(b[0], n[1], n[2]), v111 = at(n[0], n[1], n[2]);
    auto lerp64 = [](double a, double bb, double w) { return __dadd_rn(a, __dmul_rn(__dsub_rn(bb, a), w)); };
    const double vx00 = lerp64(v000, v100, d[0]);
    const double vx10 = lerp64(v010, v110, d[0]);
    const double vxx0 = lerp64(vx00, vx10, d[1]);
    const double vx01 = lerp64(v001, v101, d[0]);
    const double vx11 = lerp64(v011, v111, d[0]);
    const double vxx1 = lerp64(vx01, vx11, d[1]);
    return (float)lerp64(vxx0, vxx1, d[2]);
  }
};

template <bool NN>
__global__ void __launch_bounds__(256) depth_kernel(const DrrArgs a)
{
  __shared__ ProjConst pc;
  __shared__ xrc_cam cam_s;
  __shared__ float Xinv[12];

  const uint32_t tiles_x = (a.cols + 15u) / 16u, tiles_y = (a.rows + 15u) / 16u;
  const uint32_t proj = blockIdx.x % a.n_projs, tile = blockIdx.x / a.n_projs;
  const uint32_t ci = proj_cam_index(a, proj);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.cams + ci);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&cam_s);
    if (threadIdx.x < sizeof(xrc_cam) / 4)
      dst[threadIdx.x] = src[threadIdx.x];
  }
  __syncthreads();
  compute_proj_const(a, cam_s, proj, &pc);
  if (threadIdx.x == 0)
  {
    // Transform<float,3,Affine>::inverse() in the oracle's convention (xo_affine_inverse): cofactor inverse of the linear
    // part, det = (cof00 m00 + cof10 m10) + cof20 m20, translation -(inv . t)
    const float* X = pc.X;
    auto m = [X](int i, int j) { return X[4 * i + j]; };
    auto cof = [&m](int i, int j) {
      const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      return fsub(fmul(m(i1, j1), m(i2, j2)), fmul(m(i1, j2), m(i2, j1)));
    };
    const float c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
    const float det = fadd(fadd(fmul(c0, m(0, 0)), fmul(c1, m(1, 0))), fmul(c2, m(2, 0)));
    const float invdet = fdiv(1.0f, det);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        Xinv[4 * i + j] = fmul(cof(j, i), invdet);
    for (int i = 0; i < 3; ++i)
      Xinv[4 * i + 3] = -dot3(Xinv[4 * i], Xinv[4 * i + 1], Xinv[4 * i + 2], X[3], X[7], X[11]);
  }
  __syncthreads();
  (void)tiles_y;

  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t tx = tile % tiles_x, ty = tile / tiles_x;
  const uint32_t col = tx * 16u + (warp & 1) * 8u + (lane & 7), row = ty * 16u + (warp >> 1) * 4u + (lane >> 3);
  if (row >= a.rows || col >= a.cols)
    return;

  const size_t npix = (size_t)a.rows * a.cols;
  const size_t o = (size_t)proj * npix + (size_t)row * a.cols + col;
  float base;   // RayCasterCPU::pre_compute
  if (a.init_mode == 0)
    base = a.default_bg;
  else if (a.init_mode == 1)
    base = __ldg(a.bg + (size_t)ci * npix + (size_t)row * a.cols + col);
  else
    base = a.out[o];

  const Ray ray = setup_ray(cam_s, pc, a.step_size, a.nx, a.ny, a.nz, row, col, false);
  float out = base;
  if (ray.hit)
  {
    const VoxelReader vr(a);
    float x = ray.x, y = ray.y, z = ray.z;
    float sx = ray.sx, sy = ray.sy, sz = ray.sz;
    for (uint32_t s = 0; s < ray.nsamples; ++s)
    {
      float v = NN ? vr.nearest(x, y, z) : vr.linear(x, y, z);
      if (v >= a.depth_thresh)
      {
        for (uint32_t bt = 0; bt < a.depth_backtrack; ++bt)   // xregRayCastDepthCPU.cpp:206-216
        {
          sx = fmul(sx, 0.5f);
          sy = fmul(sy, 0.5f);
          sz = fmul(sz, 0.5f);
          const bool back = v >= a.depth_thresh;
          x = back ? fsub(x, sx) : fsub(x, -sx);
          y = back ? fsub(y, sy) : fsub(y, -sy);
          z = back ? fsub(z, sz) : fsub(z, -sz);
          v = NN ? vr.nearest(x, y, z) : vr.linear(x, y, z);
        }
        const float cx = fadd(dot3(Xinv[0], Xinv[1], Xinv[2], x, y, z), Xinv[3]);
        const float cy = fadd(dot3(Xinv[4], Xinv[5], Xinv[6], x, y, z), Xinv[7]);
        const float cz = fadd(dot3(Xinv[8], Xinv[9], Xinv[10], x, y, z), Xinv[11]);
        const float depth = norm3(fsub(cx, cam_s.pinhole[0]), fsub(cy, cam_s.pinhole[1]), fsub(cz, cam_s.pinhole[2]));
        out = (depth < base) ? depth : base;   // std::min(buf, depth), :226
        break;
      }
      x = fadd(x, sx);
      y = fadd(y, sy);
      z = fadd(z, sz);
    }
  }
  a.out[o] = out;
}

int launch_depth(const DrrArgs& a, int nearest, cudaStream_t st)
{
  if (!a.n_projs)
    return XRC_OK;
  const uint32_t tiles = ((a.cols + 15u) / 16u) * ((a.rows + 15u) / 16u);
  if (nearest)
    depth_kernel<true><<<a.n_projs * tiles, 256, 0, st>>>(a);
  else
    depth_kernel<false><<<a.n_projs * tiles, 256, 0, st>>>(a);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

// ----------------------------------------------------------------------------
// XRC_LAYOUT_PAX: principal-axis stacks (the default layout)
//
// Three padded XY-quad record stacks, one per principal ray axis k.  In stack k
// the slow axis c is volume axis k, the fast axis a = (k+1)%3 and the mid axis
// b = (k+2)%3; record (ia, ib, ic), ia in [-1, na-1], ib in [-1, nb-1],
// ic in [-1, nc], is float4 {v(ia,ib,ic), v(ia+1,ib,ic), v(ia,ib+1,ic),
// v(ia+1,ib+1,ic)} with indices clamped to the volume (replicated border).
// A CTA picks the stack of the axis its central ray is most parallel to, so
// that whatever the view direction
//   * the 8x4-pixel warp footprint lies in the (a, b) plane: the 8 lanes of a
//     quarter-warp read 4-5 consecutive 16-byte records (1-2 L1 wavefronts per
//     quarter instead of 8 when rays run along the record axis), and
//   * consecutive samples walk along c, so plane c+1 of one step is plane c of
//     the next (L1 reuse).
// The replicated one-record border makes coordinate clamping unnecessary: a
// sample that f32 drift carries to x in (-1, 0) or (n-1, n) blends two copies
// of the edge voxel, which is exactly ITK's "clamp base index / drop neighbour
// beyond the end" result (xregRayCastLineIntCPU.cpp:273).  Rays for
// which drift < 1/2 voxel cannot be proven take the clamped loop.
// Inner loop (packed): FADD2 / FFMA2 (sm_100 f32x2) for the x,y position,
// floor and the b- and c-lerps; 32-bit record index with the 1.5*2^23 magic
// constants folded into one precomputed offset.
// ----------------------------------------------------------------------------
constexpr float kMagic = 12582912.0f;  // 1.5 * 2^23: float_as_int(x + kMagic) - 0x4B400000 == floor(x), x in (-2^22, 2^22)

__device__ __forceinline__ float2 mk2(float a, float b) { return make_float2(a, b); }

struct PaxStack
{
  const float4* __restrict__ base;
  uint32_t Sb, Sc, K;
  float ha, hb, hc;
};

// &base[rec] without touching the FMA pipe: the compiler's choice for base + rec * 16 is IMAD.WIDE, which
// competes with the lerps for the pipe that bounds the loop once the loads hit L1; LEA + LEA.HI.X run on the ALU pipe.
__device__ __forceinline__ float4 pax_load(const float4* __restrict__ base, uint32_t rec)
{
  unsigned long long addr;
  asm("{\n\t"
      ".reg .u32 blo, bhi, alo, ahi;\n\t"
      "mov.b64 {blo, bhi}, %1;\n\t"
      "shf.l.wrap.b32 ahi, %2, 0, 4;\n\t"
      "shl.b32 alo, %2, 4;\n\t"
      "add.cc.u32 alo, alo, blo;\n\t"
      "addc.u32 ahi, ahi, bhi;\n\t"
      "mov.b64 %0, {alo, ahi};\n\t"
      "}"
      : "=l"(addr)
      : "l"(base), "r"(rec));
  return __ldg(reinterpret_cast<const float4*>(addr));
}

// record index + interpolation weights of one sample
template <bool PACK, bool CLAMP>
__device__ __forceinline__ void pax_cell(const PaxStack& st, float xa, float xb, float xc, uint32_t& rec, float& wa,
                                         float& wb, float& wc)
{
  if (CLAMP)
  {
    xa = fminf(fmaxf(xa, 0.0f), st.ha);
    xb = fminf(fmaxf(xb, 0.0f), st.hb);
    xc = fminf(fmaxf(xc, 0.0f), st.hc);
  }
  if (PACK)
  {
    const float2 ab = mk2(xa, xb);
    const float2 t = __fadd2_rd(ab, mk2(kMagic, kMagic));
    const float tc = __fadd_rd(xc, kMagic);
    rec = __float_as_uint(tc) * st.Sc + __float_as_uint(t.y) * st.Sb + __float_as_uint(t.x) + st.K;
    const float2 nb = __fadd2_rn(mk2(kMagic, kMagic), mk2(-t.x, -t.y));  // -(floor) exactly
    const float2 w = __fadd2_rn(ab, nb);                                  // x - floor(x), exact
    wa = w.x;
    wb = w.y;
    wc = xc - (tc - kMagic);
  }
  else
  {
    const float ta = __fadd_rd(xa, kMagic), tb = __fadd_rd(xb, kMagic), tc = __fadd_rd(xc, kMagic);
    rec = __float_as_uint(tc) * st.Sc + __float_as_uint(tb) * st.Sb + __float_as_uint(ta) + st.K;
    wa = xa - (ta - kMagic);
    wb = xb - (tb - kMagic);
    wc = xc - (tc - kMagic);
  }
}

// q0 = plane c, q1 = plane c+1, each in difference form {v00, dA, dB, dAB}:
//   dA = v10 - v00, dB = v01 - v00, dAB = v11 - v10 - v01 + v00   (computed in f64, rounded once, at repack time)
// so that the bilinear value of a plane is v00 + wa dA + wb (dB + wa dAB): 3 FMAs instead of 3 lerps (6 ops);
// with the c-lerp a sample costs 8 FP32 operations instead of 14.  Each coefficient carries <= 1/2 ulp of its
// own magnitude, the value <= ~1.5 ulp(max |v|): the same order as the rounding of the f32 lerp chain itself.
template <bool PACK>
__device__ __forceinline__ float pax_lerp(const float4& q0, const float4& q1, float wa, float wb, float wc)
{
  const float p0 = fmaf(wa, fmaf(wb, q0.w, q0.y), fmaf(wb, q0.z, q0.x));
  const float p1 = fmaf(wa, fmaf(wb, q1.w, q1.y), fmaf(wb, q1.z, q1.x));
  return fmaf(wc, p1 - p0, p0);
}

template <bool PACK>
__device__ __forceinline__ void pax_advance(float& pa, float& pb, float& pc, float sa, float sb, float sc)
{
  if (PACK)
  {
    const float2 p = __fadd2_rn(mk2(pa, pb), mk2(sa, sb));
    pa = p.x;
    pb = p.y;
  }
  else
  {
    pa = __fadd_rn(pa, sa);
    pb = __fadd_rn(pb, sb);
  }
  pc = __fadd_rn(pc, sc);
}

// Software-pipelined marching loop.  Samples are processed in groups of BATCH; the
// 2*BATCH loads of group g+1 are issued BEFORE the lerps of group g and consumed one
// iteration later (values carried across the loop back-edge, so the assembler cannot
// sink the loads next to their uses).  The kernel is bound by the L1 data pipe
// (97 % busy in ncu, profiles/); the extra loads in flight per thread keep that pipe
// fed through L1 misses.  Positions and the sum advance in sample order, exactly like
// the sequential reference loop (xregRayCastLineIntCPU.cpp:270-277).
template <int KERNEL_ID, bool PACK, bool CLAMP, int BATCH>
__device__ __forceinline__ float pax_march(const PaxStack& st, float& pa, float& pb, float& pc, float sa, float sb,
                                           float sc, uint32_t n, float sum0)
{
  // sum0: the running value (0 / -FLT_MAX at the start of a ray; the sum so far when a ray is marched in segments);
  // pa, pb, pc are left at the position after the n samples
  float sum = sum0;
  const uint32_t ng = n / BATCH;
  if (ng > 0)
  {
    float wa[BATCH], wb[BATCH], wc[BATCH];
    float4 q0[BATCH], q1[BATCH];
#pragma unroll
    for (int j = 0; j < BATCH; ++j)
    {
      uint32_t rec;
      pax_cell<PACK, CLAMP>(st, pa, pb, pc, rec, wa[j], wb[j], wc[j]);
      q0[j] = pax_load(st.base, rec);
      q1[j] = pax_load(st.base, rec + st.Sc);
      pax_advance<PACK>(pa, pb, pc, sa, sb, sc);
    }
#pragma unroll 2
    for (uint32_t g = 1; g < ng; ++g)
    {
      float nwa[BATCH], nwb[BATCH], nwc[BATCH];
      float4 n0[BATCH], n1[BATCH];
#pragma unroll
      for (int j = 0; j < BATCH; ++j)
      {
        uint32_t rec;
        pax_cell<PACK, CLAMP>(st, pa, pb, pc, rec, nwa[j], nwb[j], nwc[j]);
        n0[j] = pax_load(st.base, rec);
        n1[j] = pax_load(st.base, rec + st.Sc);
        pax_advance<PACK>(pa, pb, pc, sa, sb, sc);
      }
#pragma unroll
      for (int j = 0; j < BATCH; ++j)
      {
        const float v = pax_lerp<PACK>(q0[j], q1[j], wa[j], wb[j], wc[j]);
        sum = (KERNEL_ID == XRC_KERNEL_MAX) ? fmaxf(sum, v) : __fadd_rn(sum, v);
        q0[j] = n0[j];
        q1[j] = n1[j];
        wa[j] = nwa[j];
        wb[j] = nwb[j];
        wc[j] = nwc[j];
      }
    }
#pragma unroll
    for (int j = 0; j < BATCH; ++j)
    {
      const float v = pax_lerp<PACK>(q0[j], q1[j], wa[j], wb[j], wc[j]);
      sum = (KERNEL_ID == XRC_KERNEL_MAX) ? fmaxf(sum, v) : __fadd_rn(sum, v);
    }
  }
  for (uint32_t s = ng * BATCH; s < n; ++s)
  {
    uint32_t rec;
    float wa, wb, wc;
    pax_cell<PACK, CLAMP>(st, pa, pb, pc, rec, wa, wb, wc);
    const float4 q0 = pax_load(st.base, rec);
    const float4 q1 = pax_load(st.base, rec + st.Sc);
    const float v = pax_lerp<PACK>(q0, q1, wa, wb, wc);
    sum = (KERNEL_ID == XRC_KERNEL_MAX) ? fmaxf(sum, v) : __fadd_rn(sum, v);
    pax_advance<PACK>(pa, pb, pc, sa, sb, sc);
  }
  return sum;
}


// ----------------------------------------------------------------------------
// Empty-space trimming (exact).  A sample whose 8 corner voxels are all zero adds
// +0 to the sequential f32 sum, so not fetching it changes nothing, bit for bit.
// The volume carries a map with one bit per 8^3-voxel block, set when the block or
// any of its 26 neighbours holds a non-zero voxel (built once in set_volumes).
// If the bit at a sample's position is clear, every voxel within 8 voxels of it
// (per axis) is zero, hence so are all samples whose ideal position lies within
// (8 - 2) voxels: 1 voxel for the interpolation neighbour, 1/2 voxel for the f32
// drift of the position chain (bounded by the `safe` test), and slack.  A ray
// therefore checks one bit per m = floor(6 / max|step_axis|) samples, walks in from
// both ends while the bits are clear, and marches only [s0, s1).  The warp then
// marches the union of its lanes' ranges so that its lanes stay on the same stack
// planes (coalescing is what the kernel lives on).  Sample positions still come
// from the same chain of f32 additions as in the reference loop
// (xregRayCastLineIntCPU.cpp:270-277): the skipped samples only lose their fetch.
// ----------------------------------------------------------------------------
constexpr int kOccShift = 3;             // 8^3-voxel blocks
constexpr float kOccReach = 6.0f;        // (1 << kOccShift) - 2

// Map bit of the block holding (x, y, z) and, when it is clear, the number m >= 1 of consecutive samples starting at
// this one (walking with velocity (vx, vy, vz) voxels per sample) that the bit vouches for: block b guarantees zeros
// for voxel indices [8b - 8, 8b + 15] per axis, i.e. for samples whose ideal position stays inside [8b - 6.5, 8b + 14.5)
// (1 voxel for the interpolation neighbour, 1/2 for the drift of the f32 position chain).  iv* = 1 / max(|v*|, 1e-6).
__device__ __forceinline__ bool occ_clear(const DrrArgs& a, float x, float y, float z, float vx, float vy, float vz,
                                          float ivx, float ivy, float ivz, uint32_t& m)
{
  const int bx = min(max(__float2int_rd(x), 0), a.nx - 1) >> kOccShift;
  const int by = min(max(__float2int_rd(y), 0), a.ny - 1) >> kOccShift;
  const int bz = min(max(__float2int_rd(z), 0), a.nz - 1) >> kOccShift;
  const uint32_t w = __ldg(a.occ + ((size_t)((uint32_t)bz * a.occ_ny + (uint32_t)by) * a.occ_wx + ((uint32_t)bx >> 5)));
  const float ox = (float)(bx << kOccShift), oy = (float)(by << kOccShift), oz = (float)(bz << kOccShift);
  // distance to the end of the vouched interval in the walking direction, 0.1 voxel of slack
  const float dx = (vx > 0.f) ? (ox + 14.4f) - x : x - (ox - 6.4f);
  const float dy = (vy > 0.f) ? (oy + 14.4f) - y : y - (oy - 6.4f);
  const float dz = (vz > 0.f) ? (oz + 14.4f) - z : z - (oz - 6.4f);
  const float k = fminf(fminf(dx * ivx, dy * ivy), fminf(dz * ivz, 63.0f));
  m = 1u + (uint32_t)fmaxf(k, 0.0f);
  return ((w >> (bx & 31)) & 1u) == 0u;
}

// [s0, s1) = the samples of this ray that may be non-zero
__device__ __forceinline__ void trim_ray(const DrrArgs& a, const Ray& ray, uint32_t& s0, uint32_t& s1)
{
  const uint32_t n = ray.nsamples;
  s0 = 0;
  s1 = n;
  const float ax = fabsf(ray.sx), ay = fabsf(ray.sy), az = fabsf(ray.sz);
  if (!(fmaxf(ax, fmaxf(ay, az)) <= kOccReach))
    return;
  // 1. clip against the bounding box of the non-zero voxels (grown by 2 voxels: interpolation neighbour + drift):
  //    samples whose ideal position is outside cannot touch a non-zero voxel
  {
    float tmin = 0.0f, tmax = (float)n;
    const float p[3] = {ray.x, ray.y, ray.z}, v[3] = {ray.sx, ray.sy, ray.sz};
#pragma unroll
    for (int k = 0; k < 3; ++k)
    {
      if (fabsf(v[k]) > 1.0e-6f)
      {
        const float inv = __frcp_rn(v[k]);
        const float t1 = (a.occ_lo[k] - p[k]) * inv, t2 = (a.occ_hi[k] - p[k]) * inv;
        tmin = fmaxf(tmin, fminf(t1, t2));
        tmax = fminf(tmax, fmaxf(t1, t2));
      }
      else if (p[k] < a.occ_lo[k] || p[k] > a.occ_hi[k])
        tmax = -1.0f;
    }
    if (!(tmin <= tmax))
    {
      s0 = s1 = n;  // the ray misses everything that is not zero
      return;
    }
    s0 = (uint32_t)fmaxf(floorf(tmin) - 1.0f, 0.0f);
    s1 = min(n, (uint32_t)(ceilf(tmax) + 2.0f));
    if (s0 >= s1)
    {
      s0 = s1 = n;
      return;
    }
  }
  // 2. walk in from both ends while the map bits are clear
  const float ivx = __frcp_rn(fmaxf(ax, 1.0e-6f)), ivy = __frcp_rn(fmaxf(ay, 1.0e-6f)), ivz = __frcp_rn(fmaxf(az, 1.0e-6f));
  while (s0 < s1)
  {
    const float f = (float)s0;
    uint32_t m;
    if (!occ_clear(a, fmaf(f, ray.sx, ray.x), fmaf(f, ray.sy, ray.y), fmaf(f, ray.sz, ray.z), ray.sx, ray.sy, ray.sz, ivx,
                   ivy, ivz, m))
      break;
    s0 += m;
  }
  if (s0 >= s1)
  {
    s0 = s1 = n;  // nothing but air
    return;
  }
  while (s1 > s0)
  {
    const float f = (float)(s1 - 1u);
    uint32_t m;
    if (!occ_clear(a, fmaf(f, ray.sx, ray.x), fmaf(f, ray.sy, ray.y), fmaf(f, ray.sz, ray.z), -ray.sx, -ray.sy, -ray.sz,
                   ivx, ivy, ivz, m))
      break;
    s1 = (s1 - s0 > m) ? s1 - m : s0;
  }
}

__device__ __forceinline__ float sel3(int k, float v0, float v1, float v2) { return (k == 0) ? v0 : ((k == 1) ? v1 : v2); }

// ---- pieces shared by the two PAX kernels ----------------------------------------------------------
struct PaxCta
{
  ProjConst pc;
  xrc_cam cam;
  unsigned long long cta_samples;
  int axis, swap;
};

// Camera + pose constants to shared memory, then the stack (principal axis of the ray through pixel
// (cr, cc), the tile centre) and the warp orientation.  Any choice gives the same samples; it only
// decides the memory access pattern, so plain (contracted) arithmetic is fine here.  Returns the camera index.
__device__ __forceinline__ uint32_t pax_cta_prologue(const DrrArgs& a, PaxCta& sh, uint32_t proj, uint32_t cr, uint32_t cc)
{
  const uint32_t ci = proj_cam_index(a, proj);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.cams + ci);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sh.cam);
    if (threadIdx.x < sizeof(xrc_cam) / 4)
      dst[threadIdx.x] = src[threadIdx.x];
    if (threadIdx.x == 0)
      sh.cta_samples = 0ull;
  }
  __syncthreads();
  compute_proj_const(a, sh.cam, proj, &sh.pc);

  if (threadIdx.x == 0)
  {
    const float det_z = ((sh.cam.frame_type == 1) ? -1.0f : 1.0f) * sh.cam.focal_len;
    const float* Ki = sh.cam.intrins_inv;
    const float* E = sh.cam.extrins_inv;
    const float* X = sh.pc.X;
    // camera-frame images of (col, row, 1), d/dcol and d/drow
    float cv[3][3];
#pragma unroll
    for (int r = 0; r < 3; ++r)
    {
      cv[0][r] = det_z * (Ki[3 * r] * (float)cc + Ki[3 * r + 1] * (float)cr + Ki[3 * r + 2]);
      cv[1][r] = det_z * Ki[3 * r];
      cv[2][r] = det_z * Ki[3 * r + 1];
    }
    if (sh.cam.frame_type == 2)
      cv[0][2] -= sh.cam.focal_len;
    float iv[3][3];  // index-space: centre detector point, d/dcol, d/drow
#pragma unroll
    for (int q = 0; q < 3; ++q)
    {
      float w[3];
#pragma unroll
      for (int r = 0; r < 3; ++r)
        w[r] = E[4 * r] * cv[q][0] + E[4 * r + 1] * cv[q][1] + E[4 * r + 2] * cv[q][2] + ((q == 0) ? E[4 * r + 3] : 0.0f);
#pragma unroll
      for (int r = 0; r < 3; ++r)
        iv[q][r] = X[4 * r] * w[0] + X[4 * r + 1] * w[1] + X[4 * r + 2] * w[2] + ((q == 0) ? X[4 * r + 3] : 0.0f);
    }
    const float dx = fabsf(iv[0][0] - sh.pc.p[0]), dy = fabsf(iv[0][1] - sh.pc.p[1]), dz = fabsf(iv[0][2] - sh.pc.p[2]);
    int k = (dz >= dx && dz >= dy) ? 2 : ((dy >= dx) ? 1 : 0);
    // stacks are built on demand from a host-side estimate of the axes needed: a missing one is replaced by any
    // built one (the samples are the same, only the access pattern is worse)
    if (!a.pax[k])
    {
      if (a.pax_want)
        *(volatile uint32_t*)(a.pax_want + k) = 1u;   // host-mapped: the next compute() builds stack k
      k = a.pax[(k + 1) % 3] ? (k + 1) % 3 : (k + 2) % 3;
    }
    const int ka = (k == 2) ? 0 : k + 1;
    sh.axis = k;
    // quarter-warps (8 consecutive lanes) run along the detector direction that moves fastest along
    // axis a, so that their 8 records are consecutive in memory (1-2 L1 wavefronts per quarter)
    sh.swap = (fabsf(sel3(ka, iv[2][0], iv[2][1], iv[2][2])) > fabsf(sel3(ka, iv[1][0], iv[1][1], iv[1][2]))) ? 1 : 0;
    if (a.variant & 16)
      sh.swap = 0;
  }
  __syncthreads();
  return ci;
}

// Per-lane marching state of one ray in the chosen stack's (a, b, c) axes.
struct PaxLane
{
  PaxStack st;
  float a0, b0, c0, sa, sb, sc;
  bool hit, safe;
  uint32_t n;       // num_steps + 1
  uint32_t s0, s1;  // samples to fetch: [s0, s1) (empty-space trimming; the whole warp shares s0)
  float rx, ry, rz, rsx, rsy, rsz;   // first sample and step in volume axes (interior-gap tests)
};

template <int KERNEL_ID>
__device__ __forceinline__ PaxLane pax_lane_setup(const DrrArgs& a, const PaxCta& sh, uint32_t row, uint32_t col, bool in_img)
{
  PaxLane L;
  Ray ray;
  ray.hit = false;
  ray.nsamples = 0;
  ray.x = ray.y = ray.z = ray.sx = ray.sy = ray.sz = 0.f;
  if (in_img)
    ray = setup_ray(sh.cam, sh.pc, a.step_size, a.nx, a.ny, a.nz, row, col);
  L.hit = ray.hit;
  L.n = ray.nsamples;
  L.rx = ray.x, L.ry = ray.y, L.rz = ray.z, L.rsx = ray.sx, L.rsy = ray.sy, L.rsz = ray.sz;
  L.safe = false;
  L.st.base = nullptr;
  L.st.Sb = L.st.Sc = L.st.K = 0u;
  L.st.ha = L.st.hb = L.st.hc = 0.f;
  L.a0 = L.b0 = L.c0 = L.sa = L.sb = L.sc = 0.f;
  if (ray.hit)
  {
    const int k = sh.axis, ka = (k == 2) ? 0 : k + 1, kb = (ka == 2) ? 0 : ka + 1;
    PaxStack& st = L.st;
    st.base = (const float4*)((k == 0) ? a.pax[0] : ((k == 1) ? a.pax[1] : a.pax[2]));
    st.Sb = (k == 0) ? a.pax_sb[0] : ((k == 1) ? a.pax_sb[1] : a.pax_sb[2]);
    st.Sc = (k == 0) ? a.pax_sc[0] : ((k == 1) ? a.pax_sc[1] : a.pax_sc[2]);
    // rec = (ic+1)*Sc + (ib+1)*Sb + (ia+1) with i* = float_as_int(t*) - 0x4B400000
    st.K = (st.Sc + st.Sb + 1u) - 0x4B400000u * (st.Sc + st.Sb + 1u);
    const float hx = (float)(a.nx - 1), hy = (float)(a.ny - 1), hz = (float)(a.nz - 1);
    st.ha = sel3(ka, hx, hy, hz);
    st.hb = sel3(kb, hx, hy, hz);
    st.hc = sel3(k, hx, hy, hz);
    L.a0 = sel3(ka, ray.x, ray.y, ray.z), L.b0 = sel3(kb, ray.x, ray.y, ray.z), L.c0 = sel3(k, ray.x, ray.y, ray.z);
    L.sa = sel3(ka, ray.sx, ray.sy, ray.sz), L.sb = sel3(kb, ray.sx, ray.sy, ray.sz), L.sc = sel3(k, ray.sx, ray.sy, ray.sz);
    // drift bound: every add rounds by <= ulp(h)/2 <= h * 2^-24  ->  n * h < 2^23 keeps the
    // accumulated error below 1/2 voxel; the end points themselves lie within [-1/2, h + 1/2]
    const float fn = (float)ray.nsamples;
    const float ea = fmaf(fn, L.sa, L.a0), eb = fmaf(fn, L.sb, L.b0), ec = fmaf(fn, L.sc, L.c0);
    const float hmax = fmaxf(st.ha, fmaxf(st.hb, st.hc)) + 1.0f;
    L.safe = !(a.variant & 2) && (fn * hmax < 8388608.0f) && (fminf(L.a0, ea) > -0.5f) &&
             (fmaxf(L.a0, ea) < st.ha + 0.5f) && (fminf(L.b0, eb) > -0.5f) && (fmaxf(L.b0, eb) < st.hb + 0.5f) &&
             (fminf(L.c0, ec) > -0.5f) && (fmaxf(L.c0, ec) < st.hc + 0.5f);
  }

  // empty-space trimming (sum kernel only: a zero sample is neutral for +, not for max)
  L.s0 = 0;
  L.s1 = ray.nsamples;
  if (KERNEL_ID == XRC_KERNEL_SUM && a.occ)
  {
    if (ray.hit && L.safe)
      trim_ray(a, ray, L.s0, L.s1);
    const uint32_t lo = __reduce_min_sync(0xffffffffu, (ray.hit && L.s1 > L.s0) ? L.s0 : 0xffffffffu);
    const uint32_t hi = __reduce_max_sync(0xffffffffu, (ray.hit && L.s1 > L.s0) ? L.s1 : 0u);
    if (ray.hit && L.s1 > L.s0)
    {
      L.s0 = lo;
      L.s1 = min(hi, ray.nsamples);
    }
  }
  return L;
}

// RayCasterCPU::pre_compute + the store of ComputeLineInts (xregRayCastBaseCPU.cpp:128-158, xregRayCastLineIntCPU.cpp:279-285)
// projection buffer that holds projection `proj`: this device's, or -- tile-sharded over several GPUs -- the owner's
__device__ __forceinline__ float* proj_out(const DrrArgs& a, uint32_t proj)
{
  if (!a.peer_n)
    return a.out;
  const uint32_t big = a.peer_extra * (a.peer_base + 1u);   // projections held by the ranks that take one more
  const uint32_t owner = (proj < big) ? proj / (a.peer_base + 1u) : a.peer_extra + (proj - big) / a.peer_base;
  return a.peer_out[owner];
}

template <int KERNEL_ID>
__device__ __forceinline__ void pax_store(const DrrArgs& a, uint32_t proj, uint32_t ci, uint32_t row, uint32_t col, float sum)
{
  const size_t npix = (size_t)a.rows * a.cols;
  const size_t o = (size_t)proj * npix + (size_t)row * a.cols + col;
  float* __restrict__ out = proj_out(a, proj);
  float base_v;
  if (a.init_mode == 0)
    base_v = a.default_bg;
  else if (a.init_mode == 1)
    base_v = __ldg(a.bg + (size_t)ci * npix + (size_t)row * a.cols + col);
  else
    base_v = out[o];
  const float aa = fadd(0.0f, fmul(sum, 1.0f));  // :282
  out[o] = (KERNEL_ID == XRC_KERNEL_MAX) ? fmaxf(base_v, aa) : fadd(base_v, aa);  // :285
}

// instrumentation: add this thread's fetched-sample count to the global counter (all threads of the CTA call it)
__device__ __forceinline__ void pax_count(const DrrArgs& a, PaxCta& sh, unsigned long long n, uint32_t tile)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0)
    atomicAdd(&sh.cta_samples, n);
  __syncthreads();
  if (threadIdx.x == 0)
  {
    atomicAdd(a.sample_counter, sh.cta_samples);
    if (a.tile_counter)
      atomicAdd(a.tile_counter + tile, sh.cta_samples);
  }
}

// ---- one thread per pixel, CTA = 16 x 16 pixels of one projection ---------------------------------------
// Interior gaps (GAPS, sum kernel, on request: xrc_rc_set_skip_empty(rc, 3); for volumes whose non-zero structures lie far
// apart along the view direction):
// the range [s0, s1) left by the trimming still crosses the air between two bones.  The warp therefore marches in
// segments of kGapSeg samples and asks the empty-space map before each one: when every lane's bit is clear, the warp
// skips what the bits vouch for (minimum over the lanes) -- the skipped samples keep their three position FADDs, lose
// their fetch, and a sample whose 8 corners are zero adds +0 to the sequential sum, so no bit of the result changes
// (tests/test_gpu_skip_empty.py).  Control flow is warp-uniform: the lanes stay on the same stack planes.
constexpr uint32_t kGapSeg = 16;   // samples marched between two looks at the map
constexpr uint32_t kGapMin = 4;    // shortest run worth skipping

template <int KERNEL_ID, bool PACK, int BATCH, int MINB, bool GAPS = false>
__global__ void __launch_bounds__(kThreads, MINB) drr_pax_kernel(const DrrArgs a)
{
  __shared__ PaxCta sh;

  uint32_t proj, tile;
  cta_coords(a, proj, tile);
  const uint32_t tx = tile % a.tiles_x, ty = tile / a.tiles_x;
  const uint32_t ci = pax_cta_prologue(a, sh, proj, min(ty * kTileH + kTileH / 2, a.rows - 1),
                                       min(tx * kTileW + kTileW / 2, a.cols - 1));

  uint32_t row, col;
  {
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (sh.swap)
    {
      // warp = 4 (cols) x 8 (rows); 4 x 2 warps per 16 x 16 tile
      col = tx * kTileW + (warp >> 1) * 4 + (lane >> 3);
      row = ty * kTileH + (warp & 1) * 8 + (lane & 7);
    }
    else
    {
      col = tx * kTileW + (warp & 1) * 8 + (lane & 7);
      row = ty * kTileH + (warp >> 1) * 4 + (lane >> 3);
    }
  }
  const bool in_img = (row < a.rows) && (col < a.cols);
  PaxLane L = pax_lane_setup<KERNEL_ID>(a, sh, row, col, in_img);

  float sum = (KERNEL_ID == XRC_KERNEL_MAX) ? -3.402823466e+38f : 0.0f;
  unsigned long long fetched = 0ull;
  if (GAPS && KERNEL_ID == XRC_KERNEL_SUM)
  {
    const bool work = L.hit && L.s1 > L.s0;
    const uint32_t lo = __reduce_min_sync(0xffffffffu, work ? L.s0 : 0xffffffffu);
    const uint32_t hi = __reduce_max_sync(0xffffffffu, work ? L.s1 : 0u);
    if (work && !a.count_only)
      for (uint32_t i = 0; i < L.s0; ++i)
        pax_advance<PACK>(L.a0, L.b0, L.c0, L.sa, L.sb, L.sc);
    const float ivx = __frcp_rn(fmaxf(fabsf(L.rsx), 1.0e-6f)), ivy = __frcp_rn(fmaxf(fabsf(L.rsy), 1.0e-6f)),
                ivz = __frcp_rn(fmaxf(fabsf(L.rsz), 1.0e-6f));
    const bool can_skip = a.occ && (fmaxf(fabsf(L.rsx), fmaxf(fabsf(L.rsy), fabsf(L.rsz))) <= kOccReach);
    uint32_t sidx = lo;   // warp-uniform
    while (sidx < hi)
    {
      const bool mine = work && sidx < L.s1;    // (work lanes share s0 == lo)
      uint32_t skip = 0xffffu;
      if (mine)
      {
        skip = 0u;
        if (L.safe && can_skip)
        {
          const float f = (float)sidx;
          uint32_t m;
          if (occ_clear(a, fmaf(f, L.rsx, L.rx), fmaf(f, L.rsy, L.ry), fmaf(f, L.rsz, L.rz), L.rsx, L.rsy, L.rsz, ivx, ivy, ivz, m))
            skip = m;
        }
      }
      const uint32_t wskip = __reduce_min_sync(0xffffffffu, skip);
      if (wskip >= kGapMin)
      {
        const uint32_t k = min(wskip, hi - sidx);
        if (mine && !a.count_only)
        {
          const uint32_t kk = min(k, L.s1 - sidx);
          for (uint32_t i = 0; i < kk; ++i)
            pax_advance<PACK>(L.a0, L.b0, L.c0, L.sa, L.sb, L.sc);
        }
        sidx += k;
        continue;
      }
      const uint32_t g = min(kGapSeg, hi - sidx);
      if (mine)
      {
        const uint32_t n = min(g, L.s1 - sidx);
        fetched += n;
        if (!a.count_only)
        {
          if (L.safe)
            sum = pax_march<KERNEL_ID, PACK, false, BATCH>(L.st, L.a0, L.b0, L.c0, L.sa, L.sb, L.sc, n, sum);
          else
            sum = pax_march<KERNEL_ID, PACK, true, 1>(L.st, L.a0, L.b0, L.c0, L.sa, L.sb, L.sc, n, sum);
        }
      }
      sidx += g;
    }
  }
  else
  {
    if (L.hit && L.s1 > L.s0 && !a.count_only)
    {
      for (uint32_t i = 0; i < L.s0; ++i)
        pax_advance<PACK>(L.a0, L.b0, L.c0, L.sa, L.sb, L.sc);
      if (L.safe)
        sum = pax_march<KERNEL_ID, PACK, false, BATCH>(L.st, L.a0, L.b0, L.c0, L.sa, L.sb, L.sc, L.s1 - L.s0, sum);
      else
        sum = pax_march<KERNEL_ID, PACK, true, 1>(L.st, L.a0, L.b0, L.c0, L.sa, L.sb, L.sc, L.s1 - L.s0, sum);
    }
    fetched = L.hit ? (unsigned long long)(L.s1 - L.s0) : 0ull;
  }
  if (L.hit)
    sum = fmul(sum, a.step_size);  // xregRayCastLineIntCPU.cpp:279

  if (in_img && !a.count_only)
    pax_store<KERNEL_ID>(a, proj, ci, row, col, sum);
  if (a.sample_counter)
    pax_count(a, sh, fetched, tile);  // samples actually fetched
}

template <int KERNEL_ID>
static int launch_pax_k(const DrrArgs& a, cudaStream_t st)
{
  const uint32_t nblocks = a.n_projs * a.tile_count;
  if (!nblocks)
    return XRC_OK;
  // variant (measurement only): bit0 packed f32x2 position / floor arithmetic, bit1 force clamped loop, bits 2-3 batch id, bit4 no lane swap,
  // bits 5-7 min CTAs per SM (register budget)
  const bool packed = (a.variant & 1) != 0;
  const int batch = (a.variant >> 2) & 3;
  const int minb = (a.variant >> 5) & 7;
#define XRC_PAX_LAUNCH(B, M) drr_pax_kernel<KERNEL_ID, true, B, M><<<nblocks, kThreads, 0, st>>>(a)
  if (KERNEL_ID != XRC_KERNEL_SUM)
    drr_pax_kernel<KERNEL_ID, false, 1, 5><<<nblocks, kThreads, 0, st>>>(a);
  else if (!packed && !batch)
  {
    // Default: scalar FP32 (FFMA / FADD issue to both FMA pipes; the packed f32x2 forms save issue slots the
    // loop does not need and measured 1 % slower).  Throughput regime (a CMA-ES population: more CTAs than fit
    // at once): one sample group in flight ahead, 5 CTAs / 40 warps per SM.  Few CTAs (one wave): spend the
    // idle registers on a deeper software pipeline.  All variants add the same values in the same order.
    // at most one CTA per SM (a 192 x 192 detector, one pose): every ray is a chain of L2 / HBM-latency load groups
    // and registers are free: 8 sample groups in flight (population 1 at 192^2: 102 -> 98 us per evaluation)
    static const int deep = getenv("XRC_PAX_DEEP") ? atoi(getenv("XRC_PAX_DEEP")) : 8;   // 4: the round-1 pipeline (measurement)
    const bool gaps = a.occ && a.gaps;   // on request: interior gaps too (see GAPS above)
#define XRC_PAX_GO(B, M)                                                             \
  do                                                                                 \
  {                                                                                  \
    if (gaps)                                                                        \
      drr_pax_kernel<KERNEL_ID, false, B, M, true><<<nblocks, kThreads, 0, st>>>(a); \
    else                                                                             \
      drr_pax_kernel<KERNEL_ID, false, B, M><<<nblocks, kThreads, 0, st>>>(a);       \
  } while (0)
    if (deep == 8 && nblocks <= 160u)
      XRC_PAX_GO(8, 1);
    else if (nblocks <= 148u * 2u)
      XRC_PAX_GO(4, 2);
    else if (nblocks <= 148u * 4u)
      XRC_PAX_GO(2, 4);
    else
      XRC_PAX_GO(1, 5);
#undef XRC_PAX_GO
  }
  else if (!packed)
    drr_pax_kernel<KERNEL_ID, false, 1, 5><<<nblocks, kThreads, 0, st>>>(a);
  else if (batch == 2)
  {
    if (minb == 3) XRC_PAX_LAUNCH(2, 3);
    else XRC_PAX_LAUNCH(2, 4);
  }
  else if (batch == 3)
    XRC_PAX_LAUNCH(4, 3);
  else
  {
    if (minb == 4) XRC_PAX_LAUNCH(1, 4);
    else XRC_PAX_LAUNCH(1, 5);
  }
#undef XRC_PAX_LAUNCH
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

static int launch_pax(const DrrArgs& a, int kernel_id, cudaStream_t st)
{
  return (kernel_id == XRC_KERNEL_SUM) ? launch_pax_k<XRC_KERNEL_SUM>(a, st) : launch_pax_k<XRC_KERNEL_MAX>(a, st);
}

// same ray set-up, emits the clip mask and sample counts (parity instrumentation)
__global__ void __launch_bounds__(kThreads) ray_info_kernel(const DrrArgs a)
{
  __shared__ ProjConst pc;
  __shared__ xrc_cam cam_s;
  uint32_t proj, tile;
  cta_coords(a, proj, tile);
  const uint32_t ci = proj_cam_index(a, proj);
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(a.cams + ci);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&cam_s);
    if (threadIdx.x < sizeof(xrc_cam) / 4)
      dst[threadIdx.x] = src[threadIdx.x];
  }
  __syncthreads();
  compute_proj_const(a, cam_s, proj, &pc);
  uint32_t row, col;
  thread_pixel(a, tile, row, col);
  unsigned long long n = 0ull;
  if ((row < a.rows) && (col < a.cols))
  {
    const Ray ray = setup_ray(cam_s, pc, a.step_size, a.nx, a.ny, a.nz, row, col);
    const size_t o = (size_t)proj * a.rows * a.cols + (size_t)row * a.cols + col;
    if (a.ray_mask)
      a.ray_mask[o] = ray.hit ? 1 : 0;
    if (a.ray_steps)
      a.ray_steps[o] = ray.nsamples;
    n = ray.nsamples;
  }
  if (a.sample_counter)
  {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      n += __shfl_xor_sync(0xffffffffu, n, o);
    if ((threadIdx.x & 31) == 0 && n)
      atomicAdd(a.sample_counter, n);
  }
}

template <int LAYOUT>
static int launch_layout(const DrrArgs& a, int kernel_id, cudaStream_t st)
{
  if (a.peer_n || a.tile_count != a.tiles_x * a.tiles_y)
    XRC_FAIL(XRC_ERR_UNSUPPORTED, "tile-sharded launches need the default (principal-axis stack) layout");
  const uint32_t nblocks = a.n_projs * a.tiles_x * a.tiles_y;
  if (kernel_id == XRC_KERNEL_SUM)
    drr_kernel<LAYOUT, XRC_KERNEL_SUM><<<nblocks, kThreads, 0, st>>>(a);
  else
    drr_kernel<LAYOUT, XRC_KERNEL_MAX><<<nblocks, kThreads, 0, st>>>(a);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

// tile_stride == 0 on entry: the whole detector; else the caller's (tile_first, tile_stride, tile_count), clipped
static void fill_tiles(DrrArgs& a)
{
  a.tiles_x = (a.cols + kTileW - 1) / kTileW;
  a.tiles_y = (a.rows + kTileH - 1) / kTileH;
  const uint32_t nt = a.tiles_x * a.tiles_y;
  if (!a.tile_stride)
  {
    a.tile_first = 0;
    a.tile_stride = 1;
    a.tile_count = nt;
  }
  else if (a.tile_first >= nt)
    a.tile_count = 0;
  else
    a.tile_count = std::min(a.tile_count, (nt - a.tile_first + a.tile_stride - 1) / a.tile_stride);
}

int launch_drr(const DrrArgs& a_in, int layout, int kernel_id, cudaStream_t st)
{
  DrrArgs a = a_in;
  fill_tiles(a);
  if (!a.n_projs)
    return XRC_OK;
  switch (layout)
  {
    case XRC_LAYOUT_LINEAR: return launch_layout<XRC_LAYOUT_LINEAR>(a, kernel_id, st);
    case XRC_LAYOUT_QUAD: return launch_layout<XRC_LAYOUT_QUAD>(a, kernel_id, st);
    case XRC_LAYOUT_OCT: return launch_layout<XRC_LAYOUT_OCT>(a, kernel_id, st);
    case XRC_LAYOUT_TEX: return launch_layout<XRC_LAYOUT_TEX>(a, kernel_id, st);
    case XRC_LAYOUT_TEX_QUAD: return launch_layout<XRC_LAYOUT_TEX_QUAD>(a, kernel_id, st);
    case XRC_LAYOUT_PAX: return launch_pax(a, kernel_id, st);
    case kLayoutNN: return launch_layout<kLayoutNN>(a, kernel_id, st);
    default: XRC_FAIL(XRC_ERR_INVALID, "unknown volume layout");
  }
}

int launch_ray_info(const DrrArgs& a_in, cudaStream_t st)
{
  DrrArgs a = a_in;
  fill_tiles(a);
  if (!a.n_projs || !a.tile_count)
    return XRC_OK;
  ray_info_kernel<<<a.n_projs * a.tile_count, kThreads, 0, st>>>(a);
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


// stack k of XRC_LAYOUT_PAX: A x B x C records, A >= n[a]+1, B >= n[b]+1 (row / plane pitch, see pax_pitch), C = n[c]+2;
// records beyond n[a] / n[b] are never addressed (they hold replicated edge values like the border)

// The f32 volume read back out of a built stack: record (ia, ib, ic) of stack j keeps v(ia, ib, ic) unchanged in .x, so
// a further stack can be built from any existing one and the f32 copy need not stay resident.
struct PaxSource
{
  const float* lin;      // x-fastest f32 volume, or null
  const float4* stack;   // else: stack j
  int j;
  uint32_t sb, sc;
};

__device__ __forceinline__ float pax_source_voxel(const PaxSource& s, int x, int y, int z, int nx, int ny)
{
  if (s.lin)
    return s.lin[((size_t)z * ny + y) * nx + x];
  const int i[3] = {x, y, z};
  const int ja = (s.j + 1) % 3, jb = (s.j + 2) % 3;
  return s.stack[(size_t)(i[s.j] + 1) * s.sc + (size_t)(i[jb] + 1) * s.sb + (size_t)(i[ja] + 1)].x;
}

__global__ void repack_pax_from_kernel(const PaxSource src, float4* __restrict__ dst, int nx, int ny, int nz, int k, int A, int B)
{
  const int n[3] = {nx, ny, nz};
  const int ka = (k + 1) % 3, kb = (k + 2) % 3;
  const int C = n[k] + 2;
  const size_t total = (size_t)A * B * C;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
  {
    const int ia = (int)(i % A) - 1;
    const int ib = (int)((i / A) % B) - 1;
    const int ic = (int)(i / ((size_t)A * B)) - 1;
    int p0[3], p1[3];
    p0[ka] = min(max(ia, 0), n[ka] - 1), p1[ka] = min(max(ia + 1, 0), n[ka] - 1);
    p0[kb] = min(max(ib, 0), n[kb] - 1), p1[kb] = min(max(ib + 1, 0), n[kb] - 1);
    p0[k] = p1[k] = min(max(ic, 0), n[k] - 1);
    int q[3] = {p0[0], p0[1], p0[2]};
    const double v00 = pax_source_voxel(src, q[0], q[1], q[2], nx, ny);
    q[ka] = p1[ka];
    const double v10 = pax_source_voxel(src, q[0], q[1], q[2], nx, ny);
    q[kb] = p1[kb];
    const double v11 = pax_source_voxel(src, q[0], q[1], q[2], nx, ny);
    q[ka] = p0[ka];
    const double v01 = pax_source_voxel(src, q[0], q[1], q[2], nx, ny);
    dst[i] = make_float4((float)v00, (float)(v10 - v00), (float)(v01 - v00), (float)((v11 - v10) - (v01 - v00)));
  }
}

// Row pitch A (records) and plane pitch A * B of a PAX stack.  Measurement knob: XRC_PAX_PITCH="ra,rb" rounds the
// row pitch up to ra (mod 8 records = mod one 128-byte line) and B so that the plane pitch is rb (mod 8).
static void pax_pitch(size_t& A, size_t& B)
{
  const char* e = getenv("XRC_PAX_PITCH");
  int ra = -1, rb = -1;
  if (!e || sscanf(e, "%d,%d", &ra, &rb) != 2 || ra < 0 || ra > 7 || rb < 0 || rb > 7)
    return;
  while ((A & 7) != (size_t)ra)
    ++A;
  for (int i = 0; i < 8 && ((A * B) & 7) != (size_t)rb; ++i)
    ++B;
}


// HUToLinAttFilter::GenerateData (lib/image/xregHUToLinAtt.cpp:45-69), in place: f64 arithmetic in the reference's
// operation order (no contraction), rounded to f32 once; mu_water = 0.02683, mu_air = 0.02485e-4 (xregHUToLinAtt.h:73-74)
__global__ void hu_to_lin_att_kernel(float* __restrict__ v, size_t n, double hu_lower)
{
  const double mu_water = 0.02683 * 1.0, mu_air = 0.02485 * 0.0001;
  const double hu_scale = __dmul_rn(__dsub_rn(mu_water, mu_air), 1.0e-3);
  const double mu_lower = __dadd_rn(__dmul_rn(hu_lower, hu_scale), mu_water);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
  {
    const double a = __dsub_rn(__dadd_rn(__dmul_rn((double)v[i], hu_scale), mu_water), mu_lower);
    v[i] = (float)fmax(a, 0.0);
  }
}

void launch_hu_to_lin_att(float* d_vol, size_t n, float hu_lower, cudaStream_t st)
{
  hu_to_lin_att_kernel<<<148 * 8, 256, 0, st>>>(d_vol, n, (double)hu_lower);
  count_launch();
}

// ---- empty-space map: bit (bx, by, bz) = any non-zero voxel in blocks [bx-1, bx+1] x [by-1, by+1] x [bz-1, bz+1]
__global__ void occ_raw_kernel(const float* __restrict__ src, uint8_t* __restrict__ raw, int nx, int ny, int nz, int gx,
                               int gy)
{
  // one CTA (8 x 8 x 8 threads) per block
  const int bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;
  const int x = (bx << kOccShift) + (threadIdx.x & 7), y = (by << kOccShift) + ((threadIdx.x >> 3) & 7),
            z = (bz << kOccShift) + (threadIdx.x >> 6);
  int nonzero = 0;
  if (x < nx && y < ny && z < nz)
    nonzero = (__float_as_uint(src[((size_t)z * ny + y) * nx + x]) & 0x7fffffffu) != 0u;  // +0 and -0 are empty
  nonzero = __syncthreads_or(nonzero);
  if (threadIdx.x == 0)
    raw[((size_t)bz * gy + by) * gx + bx] = (uint8_t)(nonzero != 0);
}

// bounding box (in blocks) of the raw flags: bb = {min x, y, z, max x, y, z}, initialised to {INT_MAX.., -1..}
__global__ void occ_bbox_kernel(const uint8_t* __restrict__ raw, int* __restrict__ bb, int gx, int gy, int gz)
{
  const size_t total = (size_t)gx * gy * gz;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
  {
    if (raw[i])
    {
      const int x = (int)(i % gx), y = (int)((i / gx) % gy), z = (int)(i / ((size_t)gx * gy));
      atomicMin(bb + 0, x);
      atomicMin(bb + 1, y);
      atomicMin(bb + 2, z);
      atomicMax(bb + 3, x);
      atomicMax(bb + 4, y);
      atomicMax(bb + 5, z);
    }
  }
}

__global__ void occ_dilate_kernel(const uint8_t* __restrict__ raw, uint32_t* __restrict__ occ, int gx, int gy, int gz,
                                  int wx)
{
  const size_t total = (size_t)wx * gy * gz;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x)
  {
    const int w = (int)(i % wx), by = (int)((i / wx) % gy), bz = (int)(i / ((size_t)wx * gy));
    uint32_t bits = 0;
    for (int b = 0; b < 32; ++b)
    {
      const int bx = 32 * w + b;
      if (bx >= gx)
        break;
      int any = 0;
      for (int dz = -1; dz <= 1; ++dz)
        for (int dy = -1; dy <= 1; ++dy)
          for (int dx = -1; dx <= 1; ++dx)
          {
            const int qx = bx + dx, qy = by + dy, qz = bz + dz;
            if (qx >= 0 && qx < gx && qy >= 0 && qy < gy && qz >= 0 && qz < gz)
              any |= raw[((size_t)qz * gy + qy) * gx + qx];
          }
      bits |= (uint32_t)(any != 0) << b;
    }
    occ[i] = bits;
  }
}

int build_occupancy(const float* d_linear, DeviceVolume* v, cudaStream_t st)
{
  const int nx = (int)v->dims[0], ny = (int)v->dims[1], nz = (int)v->dims[2];
  const int E = 1 << kOccShift;
  const int gx = (nx + E - 1) / E, gy = (ny + E - 1) / E, gz = (nz + E - 1) / E;
  const int wx = (gx + 31) / 32;
  if (gy > 65535 || gz > 65535)
    return XRC_OK;  // no map: every sample is marched
  uint8_t* raw = nullptr;
  XRC_CUDA(cudaMalloc(&raw, (size_t)gx * gy * gz));
  if (cudaMalloc(&v->occ, sizeof(uint32_t) * (size_t)wx * gy * gz) != cudaSuccess)
  {
    cudaFree(raw);
    XRC_FAIL(XRC_ERR_NOMEM, "build_occupancy: out of device memory");
  }
  int* d_bb = nullptr;
  if (cudaMalloc(&d_bb, 6 * sizeof(int)) != cudaSuccess)
  {
    cudaFree(raw);
    XRC_FAIL(XRC_ERR_NOMEM, "build_occupancy: out of device memory");
  }
  int h_bb[6] = {0x7fffffff, 0x7fffffff, 0x7fffffff, -1, -1, -1};
  cudaMemcpyAsync(d_bb, h_bb, sizeof(h_bb), cudaMemcpyHostToDevice, st);
  occ_raw_kernel<<<dim3(gx, gy, gz), 512, 0, st>>>(d_linear, raw, nx, ny, nz, gx, gy);
  occ_dilate_kernel<<<148 * 2, 256, 0, st>>>(raw, v->occ, gx, gy, gz, wx);
  occ_bbox_kernel<<<148 * 2, 256, 0, st>>>(raw, d_bb, gx, gy, gz);
  count_launch(3);
  cudaMemcpyAsync(h_bb, d_bb, sizeof(h_bb), cudaMemcpyDeviceToHost, st);
  std::vector<uint8_t> h_raw((size_t)gx * gy * gz);
  cudaMemcpyAsync(h_raw.data(), raw, h_raw.size(), cudaMemcpyDeviceToHost, st);
  const cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(raw);
  cudaFree(d_bb);
  XRC_CUDA(e);
  XRC_CUDA(cudaGetLastError());
  // ideal sample positions that can touch a non-zero voxel: voxel range of the flagged blocks, grown by 2
  const int dims[3] = {nx, ny, nz};
  for (int k = 0; k < 3; ++k)
  {
    if (h_bb[3 + k] < 0)
    {
      v->occ_lo[k] = 1.0f;  // all-zero volume: an empty box
      v->occ_hi[k] = 0.0f;
    }
    else
    {
      v->occ_lo[k] = (float)(h_bb[k] * E) - 2.0f;
      v->occ_hi[k] = (float)std::min(h_bb[3 + k] * E + E - 1, dims[k] - 1) + 2.0f;
    }
  }
  v->occ_wx = (uint32_t)wx;
  v->occ_ny = (uint32_t)gy;
  v->occ_nz = (uint32_t)gz;
  // how full is the box of the non-zero voxels (blocks with a non-zero voxel / blocks of the box)?  A body in air: 0.55
  // and more; a bone-masked CT: a small fraction -- then the rays' trimmed ranges still cross air between the bones and
  // the kernel looks for interior gaps too
  v->occ_fill = 1.0f;
  if (h_bb[3] >= 0)
  {
    unsigned long long set = 0;
    for (uint8_t b : h_raw)
      set += b ? 1u : 0u;
    const double box = (double)(h_bb[3] - h_bb[0] + 1) * (double)(h_bb[4] - h_bb[1] + 1) * (double)(h_bb[5] - h_bb[2] + 1);
    v->occ_fill = (float)std::min(1.0, (double)set / std::max(box, 1.0));
  }
  return XRC_OK;
}

int build_pax_stack(DeviceVolume* v, int k, cudaStream_t st)
{
  if (v->pax[k])
    return XRC_OK;
  PaxSource src;
  src.lin = v->src;
  src.stack = nullptr;
  src.j = 0;
  src.sb = src.sc = 0;
  if (!v->src)
  {
    // no f32 copy any more: read the voxels back out of a stack that exists
    for (int j = 0; j < 3 && !src.stack; ++j)
      if (v->pax[j])
      {
        src.stack = (const float4*)v->pax[j];
        src.j = j;
        src.sb = v->pax_sb[j];
        src.sc = v->pax_sc[j];
      }
    if (!src.stack)
      XRC_FAIL(XRC_ERR_INVALID, "build_pax_stack: the volume has neither its f32 source nor a stack");
  }
  const int n[3] = {(int)v->dims[0], (int)v->dims[1], (int)v->dims[2]};
  size_t A = (size_t)n[(k + 1) % 3] + 1, B = (size_t)n[(k + 2) % 3] + 1;
  const size_t C = (size_t)n[k] + 2;
  pax_pitch(A, B);
  XRC_CUDA(cudaMalloc(&v->pax[k], sizeof(float4) * A * B * C));
  v->pax_sb[k] = (uint32_t)A;
  v->pax_sc[k] = (uint32_t)(A * B);
  v->bytes += sizeof(float4) * A * B * C;
  repack_pax_from_kernel<<<148 * 8, 256, 0, st>>>(src, (float4*)v->pax[k], n[0], n[1], n[2], k, (int)A, (int)B);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  if (v->src)
  {
    // the first stack exists: further ones are built from it, the f32 copy has served its purpose
    XRC_CUDA(cudaStreamSynchronize(st));
    v->bytes -= sizeof(float) * (size_t)n[0] * n[1] * n[2];
    cudaFree(v->src);
    v->src = nullptr;
  }
  return XRC_OK;
}

void free_volume(DeviceVolume* v)
{
  if (v->src)
    cudaFree(v->src);
  v->src = nullptr;
  if (v->h_want)
    cudaFreeHost(v->h_want);
  v->h_want = nullptr;
  if (v->occ)
    cudaFree(v->occ);
  v->occ = nullptr;
  for (int k = 0; k < 3; ++k)
  {
    if (v->pax[k])
      cudaFree(v->pax[k]);
    v->pax[k] = nullptr;
  }
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
    case XRC_LAYOUT_PAX:
    {
      // the stacks themselves are built on demand (build_pax_stack) from a device copy of the f32 volume
      const int n[3] = {nx, ny, nz};
      for (int k = 0; k < 3; ++k)
      {
        size_t A = (size_t)n[(k + 1) % 3] + 1, B = (size_t)n[(k + 2) % 3] + 1;
        const size_t C = (size_t)n[k] + 2;
        pax_pitch(A, B);
        if (A * B * C >= (1ull << 32))
          XRC_FAIL(XRC_ERR_UNSUPPORTED, "volume too large for the PAX layout (record index must fit 32 bits)");
      }
      const size_t nb = sizeof(float) * (size_t)nx * ny * nz;
      XRC_CUDA(cudaMalloc(&v->src, nb));
      XRC_CUDA(cudaMemcpyAsync(v->src, d_linear, nb, cudaMemcpyDeviceToDevice, st));
      v->bytes = nb;
      XRC_CUDA(cudaHostAlloc(&v->h_want, 3 * sizeof(uint32_t), cudaHostAllocMapped | cudaHostAllocPortable));
      v->h_want[0] = v->h_want[1] = v->h_want[2] = 0u;
      break;
    }
    default: XRC_FAIL(XRC_ERR_INVALID, "unknown volume layout");
  }
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

}  // namespace xrc
