/*
 * xreg_oracle.c -- CPU restatement of the xReg DRR + similarity-metric hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see xreg_oracle.h).  The reference has no tests for
 * this path.  xo_drr is pinned to the reference's own source lines compiled over
 * stand-in types (oracle/ref_pin/, tests/test_oracle_ref_slice.py: bit for bit),
 * and so are xo_patch_weights / xo_patch_ncc, xo_ncc (unmasked: up to the Eigen
 * reduction convention), xo_ssd, xo_grad_ncc / xo_patch_grad_ncc (OpenCV filters as
 * call-outs) and xo_hu_to_lin_att; the Gaussian / Sobel filter arithmetic itself is
 * PARITY UNPINNED by the reference and pinned by OpenCV's Python binding instead.
 *
 * Build: gcc -O2 -ffp-contract=off -fopenmp -fPIC -shared  (no -march, no -ffast-math)
 * OpenMP stands in for tbb::parallel_for at exactly the reference's parallel
 * loops (xregRayCastLineIntCPU.cpp:289, xregImgSimMetric2DNCCCPU.cpp:208,
 * xregImgSimMetric2DPatchNCCCPU.cpp:260).
 */
#include "xreg_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int xo_num_threads(void)
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Overrides OMP_NUM_THREADS (launchers such as torchrun export OMP_NUM_THREADS=1 to every rank). */
void xo_set_num_threads(int n)
{
#ifdef _OPENMP
  if (n > 0)
    omp_set_num_threads(n);
#else
  (void)n;
#endif
}

static int resolve_threads(int n_threads)
{
  const int mx = xo_num_threads();
  return (n_threads <= 0 || n_threads > mx) ? mx : n_threads;
}

/* ------------------------------------------------------------------------ */
/* f32 geometry, frozen order: dot products left to right, no FMA           */
/* ------------------------------------------------------------------------ */

static inline float dot3(float a0, float a1, float a2, float b0, float b1, float b2)
{
  return ((a0 * b0) + (a1 * b1)) + (a2 * b2);
}

/* y = A * x for a row-major 3x4 affine: (R x) + t */
static inline void affine_apply(const float a[12], const float x[3], float y[3])
{
  for (int r = 0; r < 3; ++r)
  {
    y[r] = dot3(a[4 * r], a[4 * r + 1], a[4 * r + 2], x[0], x[1], x[2]) + a[4 * r + 3];
  }
}

static inline void mat3_apply(const float m[9], const float x[3], float y[3])
{
  for (int r = 0; r < 3; ++r)
  {
    y[r] = dot3(m[3 * r], m[3 * r + 1], m[3 * r + 2], x[0], x[1], x[2]);
  }
}

/* 3x3 inverse by cofactors: inv(i,j) = cof<j,i> / det with
 * det = (cof<0,0> m00 + cof<1,0> m10) + cof<2,0> m20
 * (Eigen 3.3 compute_inverse_size3 shape; call sites xregRayCastLineIntCPU.cpp:309,
 * xregPerspectiveXform.cpp:247). */
static inline float cof3(const float m[9], int i, int j)
{
  const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
  return (m[3 * i1 + j1] * m[3 * i2 + j2]) - (m[3 * i1 + j2] * m[3 * i2 + j1]);
}

void xo_mat3_inverse(const float m[9], float out[9])
{
  const float c0 = cof3(m, 0, 0), c1 = cof3(m, 1, 0), c2 = cof3(m, 2, 0);
  const float det = ((c0 * m[0]) + (c1 * m[3])) + (c2 * m[6]);
  const float invdet = 1.0f / det;
  for (int i = 0; i < 3; ++i)
  {
    for (int j = 0; j < 3; ++j)
    {
      out[3 * i + j] = cof3(m, j, i) * invdet;
    }
  }
}

/* Transform<float,3,Affine>::inverse(): linear^-1 and -(linear^-1) t */
void xo_affine_inverse(const float a[12], float out[12])
{
  float m[9], mi[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      m[3 * r + c] = a[4 * r + c];
  xo_mat3_inverse(m, mi);
  const float t[3] = {a[3], a[7], a[11]};
  for (int r = 0; r < 3; ++r)
  {
    for (int c = 0; c < 3; ++c)
      out[4 * r + c] = mi[3 * r + c];
    out[4 * r + 3] = -dot3(mi[3 * r], mi[3 * r + 1], mi[3 * r + 2], t[0], t[1], t[2]);
  }
}

/* out = a * b (affine * affine): linear = A B, translation = (A b_t) + a_t
 * (xregRayCastLineIntCPU.cpp:207-208) */
void xo_affine_compose(const float a[12], const float b[12], float out[12])
{
  float tmp[12];
  for (int r = 0; r < 3; ++r)
  {
    for (int c = 0; c < 3; ++c)
    {
      tmp[4 * r + c] = dot3(a[4 * r], a[4 * r + 1], a[4 * r + 2], b[c], b[4 + c], b[8 + c]);
    }
    tmp[4 * r + 3] = dot3(a[4 * r], a[4 * r + 1], a[4 * r + 2], b[3], b[7], b[11]) + a[4 * r + 3];
  }
  memcpy(out, tmp, sizeof(tmp));
}

/* lib/transforms/xregPerspectiveXform.cpp:200-254 */
void xo_cam_setup_naive(xo_cam* cam, float focal_len, uint32_t rows, uint32_t cols,
                        float row_spacing, float col_spacing, int32_t frame_type)
{
  float K[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  K[0] = focal_len / col_spacing;
  K[4] = focal_len / row_spacing;
  if (frame_type == 1)
  {
    K[0] *= -1;
    K[4] *= -1;
  }
  /* (num_cols - 1) * 0.5 is evaluated in double then narrowed to CoordScalar */
  K[2] = (float)((double)(cols - 1) * 0.5);
  K[5] = (float)((double)(rows - 1) * 0.5);

  memset(cam, 0, sizeof(*cam));
  cam->rows = rows;
  cam->cols = cols;
  cam->focal_len = focal_len;
  cam->frame_type = frame_type;
  xo_mat3_inverse(K, cam->intrins_inv);
  cam->extrins_inv[0] = cam->extrins_inv[5] = cam->extrins_inv[10] = 1.0f;
  /* pinhole_pt = Pt3::Zero() for every frame type (xregPerspectiveXform.cpp:253),
   * honoured literally. */
}

/* lib/transforms/xregPerspectiveXform.cpp:302-334 */
void xo_cam_setup(xo_cam* cam, const float intrins[9], const float extrins[16],
                  uint32_t rows, uint32_t cols, float row_spacing, float col_spacing,
                  int32_t frame_type)
{
  memset(cam, 0, sizeof(*cam));
  cam->rows = rows;
  cam->cols = cols;
  cam->frame_type = frame_type;
  xo_mat3_inverse(intrins, cam->intrins_inv);
  /* FocalLenFromIntrins (xregPerspectiveXform.cpp:186-190) */
  cam->focal_len = (fabsf(intrins[0] * col_spacing) +
                    fabsf(intrins[4] * ((row_spacing < 0) ? col_spacing : row_spacing))) / 2.0f;
  /* SE3Inv (lib/transforms/xregRigidUtils.cpp:29-38): R^T, -1 * R^T * t */
  float rt[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c)
      rt[3 * r + c] = extrins[4 * c + r];
  const float t[3] = {extrins[3], extrins[7], extrins[11]};
  for (int r = 0; r < 3; ++r)
  {
    for (int c = 0; c < 3; ++c)
      cam->extrins_inv[4 * r + c] = rt[3 * r + c];
    cam->extrins_inv[4 * r + 3] =
        dot3(-1.0f * rt[3 * r], -1.0f * rt[3 * r + 1], -1.0f * rt[3 * r + 2], t[0], t[1], t[2]);
  }
  if (frame_type == 0 || frame_type == 1)
  {
    cam->pinhole[0] = cam->extrins_inv[3];
    cam->pinhole[1] = cam->extrins_inv[7];
    cam->pinhole[2] = cam->extrins_inv[11];
  }
  else
  {
    const float ph[3] = {0.0f, 0.0f, cam->focal_len};
    affine_apply(cam->extrins_inv, ph, cam->pinhole);
  }
}

/* lib/ray_cast/xregRayCastInterface.cpp:97-114: camera-major replication */
void xo_distribute_xforms(const float* poses, uint32_t n_poses, uint32_t n_cams,
                          float* out_poses, uint32_t* out_cam_idx)
{
  uint32_t g = 0;
  for (uint32_t c = 0; c < n_cams; ++c)
  {
    for (uint32_t p = 0; p < n_poses; ++p, ++g)
    {
      memcpy(out_poses + 12 * (size_t)g, poses + 12 * (size_t)p, 12 * sizeof(float));
      out_cam_idx[g] = c;
    }
  }
}

/* lib/ray_cast/xregRayCastBaseCPU.cpp:128-158 */
void xo_pre_compute(float* buf, uint32_t n_projs, uint32_t rows, uint32_t cols,
                    const uint32_t* cam_idx, const float* const* bg_projs,
                    int store_method, float default_bg)
{
  const size_t npix = (size_t)rows * cols;
  if (bg_projs)
  {
    for (uint32_t p = 0; p < n_projs; ++p)
    {
      memcpy(buf + p * npix, bg_projs[cam_idx[p]], sizeof(float) * npix);
    }
  }
  if (store_method == XO_STORE_REPLACE && !bg_projs)
  {
    const size_t tot = npix * n_projs;
    for (size_t i = 0; i < tot; ++i)
      buf[i] = default_bg;
  }
}

/* ITK 5.1.1 LinearInterpolateImageFunction<Image<float,3>,float>::EvaluateOptimized
 * (Dispatch<3>), call site xregRayCastLineIntCPU.cpp:273-274.  ITK itself is an
 * un-vendored dependency (README.md:57-69, ITK 5.1.1); this restates its
 * published algorithm: base index = floor clamped to the start index, f32
 * distances, f64 lerps x then y then z, neighbours beyond the end index dropped
 * and axes whose distance is <= 0 not interpolated.  Dropping a neighbour /
 * skipping an axis is arithmetically identical to a lerp with weight 0, which
 * is how it is written here. */
double xo_interp_linear(const float* vol, const uint64_t dims[3], const float x[3])
{
  int64_t b[3], b1[3];
  float w[3];
  for (int k = 0; k < 3; ++k)
  {
    int64_t bk = (int64_t)floorf(x[k]);
    if (bk < 0)
      bk = 0;
    if (bk > (int64_t)dims[k] - 1) /* never read out of bounds (ITK would) */
      bk = (int64_t)dims[k] - 1;
    float d = x[k] - (float)bk;
    int64_t nk = bk + 1;
    if (d <= 0.0f)
    {
      d = 0.0f;
      nk = bk;
    }
    if (nk > (int64_t)dims[k] - 1)
    {
      nk = bk;
      d = 0.0f;
    }
    b[k] = bk;
    b1[k] = nk;
    w[k] = d;
  }
  const size_t sx = 1, sy = dims[0], sz = dims[0] * dims[1];
#define V(i, j, k) ((double)vol[(size_t)(i) * sx + (size_t)(j) * sy + (size_t)(k) * sz])
  const double v000 = V(b[0], b[1], b[2]), v100 = V(b1[0], b[1], b[2]);
  const double v010 = V(b[0], b1[1], b[2]), v110 = V(b1[0], b1[1], b[2]);
  const double v001 = V(b[0], b[1], b1[2]), v101 = V(b1[0], b[1], b1[2]);
  const double v011 = V(b[0], b1[1], b1[2]), v111 = V(b1[0], b1[1], b1[2]);
#undef V
  const double d0 = w[0], d1 = w[1], d2 = w[2];
  const double vx00 = v000 + (v100 - v000) * d0;
  const double vx10 = v010 + (v110 - v010) * d0;
  const double vxx0 = vx00 + (vx10 - vx00) * d1;
  const double vx01 = v001 + (v101 - v001) * d0;
  const double vx11 = v011 + (v111 - v011) * d0;
  const double vxx1 = vx01 + (vx11 - vx01) * d1;
  return vxx0 + (vxx1 - vxx0) * d2;
}

/* ITK 5.1.1 NearestNeighborInterpolateImageFunction<Image<float,3>,float>::EvaluateAtContinuousIndex: the pixel at
 * ConvertContinuousIndexToNearestIndex(x), i.e. Index::CopyWithRound -> Math::RoundHalfIntegerUp per axis (the other
 * interpolator RayCasterLineIntCPU can select, xregRayCastLineIntCPU.cpp:128-130).  ITK is un-vendored: restated as
 * floor(x + 0.5) evaluated exactly (in double; ITK's SSE form 2x + 0.5 -> cvt -> >> 1 agrees except within one f32 ulp
 * of a half-integer at coordinates >= 2^22).  ITK would read out of bounds for an index outside the image; the ray
 * caster's 1e-3 nudge keeps x inside [0, n - 1] up to f32 drift, and the index is clamped here. */
double xo_interp_nn(const float* vol, const uint64_t dims[3], const float x[3])
{
  int64_t b[3];
  for (int k = 0; k < 3; ++k)
  {
    int64_t bk = (int64_t)floor((double)x[k] + 0.5);
    if (bk < 0)
      bk = 0;
    if (bk > (int64_t)dims[k] - 1)
      bk = (int64_t)dims[k] - 1;
    b[k] = bk;
  }
  return (double)vol[(size_t)b[0] + (size_t)b[1] * dims[0] + (size_t)b[2] * dims[0] * dims[1]];
}

/* lib/spatial/xregSpatialPrimitives.cpp:175-222, limit_to_segment = true */
static int ray_rect_intersect_ex(const float mn[3], const float mx[3], const float p[3],
                                 const float d[3], int limit_to_segment, float* t_start, float* t_stop);
static int ray_rect_intersect(const float mn[3], const float mx[3], const float p[3],
                              const float d[3], float* t_start, float* t_stop)
{
  return ray_rect_intersect_ex(mn, mx, p, d, 1, t_start, t_stop);
}
static int ray_rect_intersect_ex(const float mn[3], const float mx[3], const float p[3],
                                 const float d[3], int limit_to_segment, float* t_start, float* t_stop)
{
  float t0 = 0.0f, t1 = limit_to_segment ? 1.0f : INFINITY;
  int hit = 1;
  for (int k = 0; k < 3; ++k)
  {
    if (fabsf(d[k]) > 1.0e-8f) /* double literal compare in the reference; same set of floats */
    {
      const float inv = 1.0f / d[k];
      float a = (mn[k] - p[k]) * inv;
      float b = (mx[k] - p[k]) * inv;
      if (b < a)
      {
        const float tmp = a;
        a = b;
        b = tmp;
      }
      t0 = (t0 < a) ? a : t0; /* std::max(t_start, t[0]) */
      t1 = (b < t1) ? b : t1; /* std::min(t_stop, t[1]) */
      if (t0 > t1)
      {
        hit = 0;
        break;
      }
    }
    else if ((p[k] < mn[k]) || (p[k] > mx[k]))
    {
      hit = 0;
      break;
    }
  }
  *t_start = t0;
  *t_stop = t1;
  return hit;
}

static inline float norm3(const float v[3])
{
  return sqrtf(((v[0] * v[0]) + (v[1] * v[1])) + (v[2] * v[2]));
}

#define XO_VOL_BB_STEP_INC_TOL 1.0e-3f /* xregRayCastBaseCPU.h:37 */

/* lib/ray_cast/xregRayCastLineIntCPU.cpp:105-349 */
int xo_drr(const float* vol, const uint64_t dims[3], const float idx_to_phys[12],
           const xo_cam* cams, uint32_t n_cams,
           const float* poses, const uint32_t* cam_idx, uint32_t n_projs,
           float step_size, int kernel_id,
           float* buf, uint8_t* hit_mask, uint32_t* num_steps_out,
           uint64_t* total_samples, int n_threads)
{
  return xo_drr_interp(vol, dims, idx_to_phys, cams, n_cams, poses, cam_idx, n_projs, step_size, kernel_id, XO_INTERP_LINEAR,
                       buf, hit_mask, num_steps_out, total_samples, n_threads);
}

int xo_drr_interp(const float* vol, const uint64_t dims[3], const float idx_to_phys[12],
                  const xo_cam* cams, uint32_t n_cams,
                  const float* poses, const uint32_t* cam_idx, uint32_t n_projs,
                  float step_size, int kernel_id, int interp,
                  float* buf, uint8_t* hit_mask, uint32_t* num_steps_out,
                  uint64_t* total_samples, int n_threads)
{
  if (!n_cams || !n_projs)
    return 0;
  if (interp != XO_INTERP_LINEAR && interp != XO_INTERP_NN)
    return -3; /* sinc / B-spline: not restated */
  const uint32_t rows = cams[0].rows, cols = cams[0].cols;
  for (uint32_t c = 1; c < n_cams; ++c)
  {
    if (cams[c].rows != rows || cams[c].cols != cols)
      return -1; /* xregRayCastBaseCPU.cpp:60-70 */
  }
  for (uint32_t p = 0; p < n_projs; ++p)
  {
    if (cam_idx[p] >= n_cams)
      return -2;
  }

  /* xregITKBasicImageUtils.h:54-75 */
  const float aabb_min[3] = {0.0f, 0.0f, 0.0f};
  const float aabb_max[3] = {(float)(dims[0] - 1), (float)(dims[1] - 1), (float)(dims[2] - 1)};

  float phys_to_idx[12];
  xo_affine_inverse(idx_to_phys, phys_to_idx); /* :307-309 */

  const size_t npix = (size_t)rows * cols;
  const int64_t n_rays = (int64_t)npix * n_projs;
  uint64_t S = 0;
  const int nt = resolve_threads(n_threads);
  (void)nt;

#pragma omp parallel for schedule(dynamic, 1024) num_threads(nt) reduction(+ : S)
  for (int64_t i = 0; i < n_rays; ++i)
  {
    const size_t proj = (size_t)i / npix;
    const size_t off = (size_t)i - npix * proj;
    const size_t row = off / cols;
    const size_t col = off - (size_t)cols * row;
    const xo_cam* cam = &cams[cam_idx[proj]];

    /* CameraModel::ind_pt_to_phys_det_pt (xregPerspectiveXform.cpp:391-414) */
    const float det_z = ((cam->frame_type == 1) ? -1.0f : 1.0f) * cam->focal_len;
    const float ind[3] = {det_z * (float)col, det_z * (float)row, det_z * 1.0f};
    float tmp3[3], det[3];
    mat3_apply(cam->intrins_inv, ind, tmp3);
    if (cam->frame_type == 2)
    {
      tmp3[0] = tmp3[0] + 0.0f;
      tmp3[1] = tmp3[1] + 0.0f;
      tmp3[2] = tmp3[2] + (-cam->focal_len);
    }
    affine_apply(cam->extrins_inv, tmp3, det);

    float X[12];
    xo_affine_compose(phys_to_idx, poses + 12 * proj, X); /* :207-208 */

    float p[3], xd[3], d[3];
    affine_apply(X, cam->pinhole, p); /* :211 */
    affine_apply(X, det, xd);         /* :215-216 */
    d[0] = xd[0] - p[0];
    d[1] = xd[1] - p[1];
    d[2] = xd[2] - p[2];

    float t0 = 0.0f, t1 = 0.0f;
    const int hit = ray_rect_intersect(aabb_min, aabb_max, p, d, &t0, &t1);

    float sum = (kernel_id == XO_KERNEL_MAX) ? -FLT_MAX : 0.0f;
    uint8_t m = 0;
    uint32_t ns = 0;

    if (hit && ((t1 - t0) > (2.0f * XO_VOL_BB_STEP_INC_TOL))) /* :233 */
    {
      m = 1;
      t0 += XO_VOL_BB_STEP_INC_TOL;
      t1 -= XO_VOL_BB_STEP_INC_TOL;

      float x[3] = {p[0] + (t0 * d[0]), p[1] + (t0 * d[1]), p[2] + (t0 * d[2])}; /* :240-241 */
      const float L = norm3(d);                                                  /* :243-244 */
      const float len = (t1 - t0) * L;                                           /* :245-246 */

      /* :249-251  (X.linear * (normalized(det - pinhole) * step)).norm() */
      float dir[3] = {det[0] - cam->pinhole[0], det[1] - cam->pinhole[1], det[2] - cam->pinhole[2]};
      const float dn = norm3(dir);
      dir[0] = (dir[0] / dn) * step_size;
      dir[1] = (dir[1] / dn) * step_size;
      dir[2] = (dir[2] / dn) * step_size;
      float sv[3];
      for (int r = 0; r < 3; ++r)
        sv[r] = dot3(X[4 * r], X[4 * r + 1], X[4 * r + 2], dir[0], dir[1], dir[2]);
      const float step_len = norm3(sv);

      const uint64_t num_steps = (uint64_t)(len / step_len); /* :253-254 */
      const float scale = step_len / L;                      /* :263-264 */
      const float stepv[3] = {d[0] * scale, d[1] * scale, d[2] * scale};

      for (uint64_t s = 0; s <= num_steps; ++s) /* :270-277 */
      {
        const float v = (interp == XO_INTERP_NN) ? (float)xo_interp_nn(vol, dims, x) : (float)xo_interp_linear(vol, dims, x);
        if (kernel_id == XO_KERNEL_MAX)
          sum = (sum < v) ? v : sum;
        else
          sum = sum + v;
        x[0] += stepv[0];
        x[1] += stepv[1];
        x[2] += stepv[2];
      }
      sum *= step_size; /* :279 */
      ns = (uint32_t)(num_steps + 1);
      S += num_steps + 1;
    }
    const float aa_sum = 0.0f + (sum * 1.0f); /* :282 */
    if (kernel_id == XO_KERNEL_MAX)
      buf[i] = (buf[i] < aa_sum) ? aa_sum : buf[i];
    else
      buf[i] = buf[i] + aa_sum; /* :285 */
    if (hit_mask)
      hit_mask[i] = m;
    if (num_steps_out)
      num_steps_out[i] = ns;
  }
  if (total_samples)
    *total_samples = S;
  return 0;
}

/* RayCasterDepthCPU::compute / RayCastDepthFn::operator() (lib/ray_cast/xregRayCastDepthCPU.cpp:42-272): per ray, walk
 * the volume from the source side in steps of step_size; at the first sample whose interpolated value is >=
 * collision_thresh refine the crossing by num_backtracking_steps halvings of the step (:206-216), take the distance of
 * that point from the pinhole in the camera frame (through the INVERSE of the camera -> index transform, :97, :228) and
 * store buf = min(buf, depth) (:226).  Rays that never reach the threshold leave buf alone (the class initialises it
 * with kRAY_CAST_MAX_DEPTH = 1e37 through pre_compute; call xo_pre_compute with that default first).  The ray is NOT
 * limited to the source-detector segment (RayRectIntersect(..., false), :118-121).  Eigen's Transform::inverse() is
 * the stated convention of xo_affine_inverse. */
int xo_depth(const float* vol, const uint64_t dims[3], const float idx_to_phys[12],
             const xo_cam* cams, uint32_t n_cams,
             const float* poses, const uint32_t* cam_idx, uint32_t n_projs,
             float step_size, int interp, float collision_thresh, uint32_t num_backtracking_steps,
             float* buf, int n_threads)
{
  if (!n_cams || !n_projs)
    return 0;
  if (interp != XO_INTERP_LINEAR && interp != XO_INTERP_NN)
    return -3;
  const uint32_t rows = cams[0].rows, cols = cams[0].cols;
  for (uint32_t c = 1; c < n_cams; ++c)
    if (cams[c].rows != rows || cams[c].cols != cols)
      return -1;
  for (uint32_t p = 0; p < n_projs; ++p)
    if (cam_idx[p] >= n_cams)
      return -2;
  const float aabb_min[3] = {0.0f, 0.0f, 0.0f};
  const float aabb_max[3] = {(float)(dims[0] - 1), (float)(dims[1] - 1), (float)(dims[2] - 1)};
  float phys_to_idx[12];
  xo_affine_inverse(idx_to_phys, phys_to_idx); /* :245-249 */
  const size_t npix = (size_t)rows * cols;
  const int64_t n_rays = (int64_t)npix * n_projs;
  const int nt = resolve_threads(n_threads);
  (void)nt;

#pragma omp parallel for schedule(dynamic, 1024) num_threads(nt)
  for (int64_t i = 0; i < n_rays; ++i)
  {
    const size_t proj = (size_t)i / npix;
    const size_t off = (size_t)i - npix * proj;
    const size_t row = off / cols;
    const size_t col = off - (size_t)cols * row;
    const xo_cam* cam = &cams[cam_idx[proj]];

    /* CameraModel::ind_pt_to_phys_det_pt (xregPerspectiveXform.cpp:391-414), :88-89 */
    const float det_z = ((cam->frame_type == 1) ? -1.0f : 1.0f) * cam->focal_len;
    const float ind[3] = {det_z * (float)col, det_z * (float)row, det_z * 1.0f};
    float tmp3[3], det[3];
    mat3_apply(cam->intrins_inv, ind, tmp3);
    if (cam->frame_type == 2)
    {
      tmp3[0] = tmp3[0] + 0.0f;
      tmp3[1] = tmp3[1] + 0.0f;
      tmp3[2] = tmp3[2] + (-cam->focal_len);
    }
    affine_apply(cam->extrins_inv, tmp3, det);

    float X[12], Xinv[12];
    xo_affine_compose(phys_to_idx, poses + 12 * proj, X); /* :93 */
    xo_affine_inverse(X, Xinv);                           /* :95 */

    float p[3], xd[3], d[3];
    affine_apply(X, cam->pinhole, p); /* :98 */
    affine_apply(X, det, xd);         /* :101 */
    d[0] = xd[0] - p[0];
    d[1] = xd[1] - p[1];
    d[2] = xd[2] - p[2];

    float t0 = 0.0f, t1 = 0.0f;
    const int hit = ray_rect_intersect_ex(aabb_min, aabb_max, p, d, 0, &t0, &t1); /* :118-121 */
    if (!(hit && ((t1 - t0) > (2.0f * XO_VOL_BB_STEP_INC_TOL))))                  /* :125 */
      continue;
    t0 += XO_VOL_BB_STEP_INC_TOL;
    t1 -= XO_VOL_BB_STEP_INC_TOL;
    float x[3] = {p[0] + (t0 * d[0]), p[1] + (t0 * d[1]), p[2] + (t0 * d[2])}; /* :131 */
    const float L = norm3(d);                                                  /* :133 */
    const float len = (t1 - t0) * L;                                           /* :134 */
    float dir[3] = {det[0] - cam->pinhole[0], det[1] - cam->pinhole[1], det[2] - cam->pinhole[2]};
    const float dn = norm3(dir);
    dir[0] = (dir[0] / dn) * step_size;
    dir[1] = (dir[1] / dn) * step_size;
    dir[2] = (dir[2] / dn) * step_size;
    float sv0[3];
    for (int r = 0; r < 3; ++r)
      sv0[r] = dot3(X[4 * r], X[4 * r + 1], X[4 * r + 2], dir[0], dir[1], dir[2]);
    const float step_len = norm3(sv0);                     /* :137 */
    const uint64_t num_steps = (uint64_t)(len / step_len); /* :139 */
    const float scale = step_len / L;                      /* :147 */
    float sv[3] = {d[0] * scale, d[1] * scale, d[2] * scale};

    for (uint64_t s = 0; s <= num_steps; ++s) /* :155 */
    {
      float v = (interp == XO_INTERP_NN) ? (float)xo_interp_nn(vol, dims, x) : (float)xo_interp_linear(vol, dims, x);
      if (v >= collision_thresh)
      {
        for (uint32_t b = 0; b < num_backtracking_steps; ++b) /* :162-171 */
        {
          sv[0] *= 0.5f;
          sv[1] *= 0.5f;
          sv[2] *= 0.5f;
          if (v >= collision_thresh)
          {
            x[0] = x[0] - sv[0];
            x[1] = x[1] - sv[1];
            x[2] = x[2] - sv[2];
          }
          else
          {
            x[0] = x[0] - (-sv[0]);
            x[1] = x[1] - (-sv[1]);
            x[2] = x[2] - (-sv[2]);
          }
          v = (interp == XO_INTERP_NN) ? (float)xo_interp_nn(vol, dims, x) : (float)xo_interp_linear(vol, dims, x);
        }
        float c[3];
        affine_apply(Xinv, x, c); /* :180-181 */
        const float e[3] = {c[0] - cam->pinhole[0], c[1] - cam->pinhole[1], c[2] - cam->pinhole[2]};
        const float depth = norm3(e);
        buf[i] = (depth < buf[i]) ? depth : buf[i]; /* std::min(buf, depth) */
        break;
      }
      x[0] += sv[0];
      x[1] += sv[1];
      x[2] += sv[2];
    }
  }
  return 0;
}

/* ---- log remap of a projection (pre-processing) -------------------------------------------------
 * ImageIntensLogTransFilter::GenerateData (lib/image/xregImageIntensLogTrans.cpp:55-144) is restated line by line
 * (xo_log_remap) and pinned to the reference's own lines (tests/test_oracle_ref_slice.py).  Its default path takes I0 from
 * the maximum of the image smoothed by itk::DiscreteGaussianImageFilter (variance 2, :97-105): that filter is ITK 5.1.1, an
 * un-vendored dependency -- its published algorithm is restated below and is PARITY UNPINNED (nothing of ITK can be run
 * here): itk::GaussianOperator::GenerateCoefficients (discrete Gaussian e^-t I_n(t) from the modified Bessel functions, the
 * Abramowitz & Stegun 9.8.1-9.8.4 polynomials and the downward recurrence with accuracy 40, summed until 1 - max error,
 * at most 32 wide, normalised), applied per dimension (y first, then x: the filter assigns the operators in reverse
 * order) with a zero-flux Neumann boundary, double accumulation in kernel order, a float image between the passes. */
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
  const double ACCURACY = 40.0;
  if (y == 0.0)
    return 0.0;
  const double toy = 2.0 / fabs(y);
  double qip = 0.0, acc = 0.0, qi = 1.0;
  for (int j = 2 * (n + (int)sqrt(ACCURACY * n)); j > 0; j--)
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

/* coeffs[0 .. 2 radius]: the symmetric kernel; returns the radius (<= max_width) */
int xo_itk_gaussian_coeffs(double variance, double max_error, int max_width, double* coeffs)
{
  double half[80];
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
    if (half[i] <= 0.0)
      break;
    if (n > max_width || n >= 79)
      break;
  }
  for (int i = 0; i < n; ++i)
    half[i] /= sum;
  const int r = n - 1;
  for (int i = 0; i <= r; ++i)
  {
    coeffs[r + i] = half[i];
    coeffs[r - i] = half[i];
  }
  return r;
}

void xo_itk_discrete_gaussian_2d(const float* img, uint32_t rows, uint32_t cols, double variance, float* out)
{
  double k[160];
  const int r = xo_itk_gaussian_coeffs(variance, 0.01, 32, k);
  float* tmp = (float*)malloc(sizeof(float) * (size_t)rows * cols);
  for (int64_t y = 0; y < (int64_t)rows; ++y)   /* first filter: along y */
    for (int64_t x = 0; x < (int64_t)cols; ++x)
    {
      double s = 0.0;
      for (int i = 0; i <= 2 * r; ++i)
      {
        int64_t yy = y + i - r;
        yy = yy < 0 ? 0 : (yy > (int64_t)rows - 1 ? (int64_t)rows - 1 : yy);
        s += k[i] * (double)img[(size_t)yy * cols + x];
      }
      tmp[(size_t)y * cols + x] = (float)s;
    }
  for (int64_t y = 0; y < (int64_t)rows; ++y)   /* last filter: along x */
    for (int64_t x = 0; x < (int64_t)cols; ++x)
    {
      double s = 0.0;
      for (int i = 0; i <= 2 * r; ++i)
      {
        int64_t xx = x + i - r;
        xx = xx < 0 ? 0 : (xx > (int64_t)cols - 1 ? (int64_t)cols - 1 : xx);
        s += k[i] * (double)tmp[(size_t)y * cols + xx];
      }
      out[(size_t)y * cols + x] = (float)s;
    }
  free(tmp);
}

/* ImageIntensLogTransFilter::GenerateData (lib/image/xregImageIntensLogTrans.cpp:55-144).  smoothed: the output of the
 * DiscreteGaussianImageFilter call (:97-103) when use_max_intensity_as_I0 and not normalize_zero_one, else unused (NULL:
 * computed with the restatement above).  I0_used (optional) returns the I0 of the log map. */
void xo_log_remap(const float* img, uint32_t rows, uint32_t cols, int normalize_zero_one, int use_max_intensity_as_I0,
                  float I0, const float* smoothed, float* out, float* I0_used)
{
  const float eps = 1.0e-6f;
  const size_t n = (size_t)rows * cols;
  float I0_to_use = I0;
  const float* src = img;
  if (normalize_zero_one)
  {
    float mx = img[0];
    for (size_t i = 1; i < n; ++i)
      mx = (mx < img[i]) ? img[i] : mx; /* std::max_element */
    const float scale = 1.0f / mx;
    for (size_t i = 0; i < n; ++i)
      out[i] = img[i] * scale;
    src = out;
    if (use_max_intensity_as_I0)
      I0_to_use = 1.0f;
  }
  else if (use_max_intensity_as_I0)
  {
    float* own = NULL;
    if (!smoothed)
    {
      own = (float*)malloc(sizeof(float) * n);
      xo_itk_discrete_gaussian_2d(img, rows, cols, 2.0, own);
      smoothed = own;
    }
    float mx = smoothed[0];
    for (size_t i = 1; i < n; ++i)
      mx = (mx < smoothed[i]) ? smoothed[i] : mx;
    I0_to_use = mx;
    free(own);
  }
  float min_pos = 0.0f;
  int found = 0;
  for (size_t i = 0; i < n; ++i)
  {
    if (src[i] > eps)
    {
      if (found)
      {
        if (min_pos > src[i])
          min_pos = src[i];
      }
      else
      {
        min_pos = src[i];
        found = 1;
      }
    }
  }
  const float out_max_val = -logf(min_pos / I0_to_use);
  for (size_t i = 0; i < n; ++i)
  {
    const float x = src[i];
    out[i] = (x > eps) ? -logf(x / I0_to_use) : out_max_val;
  }
  if (I0_used)
    *I0_used = I0_to_use;
}

/* ---- down-sampling of a projection image (pre-processing) ------------------------------------------
 * DownsampleImage (lib/itk/xregITKResampleUtils.h:49-112, the B-spline default of :181-188), what DownsampleProjData
 * (lib/image/xregProjData.cpp:40-99) applies to the fixed image of every registration level: smooth with
 * itk::DiscreteGaussianImageFilter (variance (0.5 / factor)^2, image spacing ignored) when factor < 1, then
 * itk::ResampleImageFilter with the identity transform, the same origin / direction, spacing / factor, size
 * (unsigned long)(size * factor + 0.5), and a cubic itk::BSplineInterpolateImageFunction.  The control flow is the
 * reference's; ALL of the arithmetic is ITK 5.1.1's (un-vendored): restated from its published algorithms and PARITY
 * UNPINNED -- the discrete Gaussian above; BSplineDecompositionImageFilter (Unser's recursive prefilter, pole sqrt(3) - 2,
 * gain (1 - z)(1 - 1/z), causal initialisation truncated at tolerance 1e-10 or the full mirror sum, anti-causal
 * initialisation z / (z^2 - 1) (z c[N-2] + c[N-1]), dimension 0 first, double coefficients); evaluation at the continuous
 * input index i / factor (identity transform, same origin and direction) with the cubic weights
 * w3 = w^3 / 6, w0 = 1/6 + w (w - 1) / 2 - w3, w2 = w + w0 - 2 w3, w1 = 1 - w0 - w2 - w3, support floor(x) - 1 .. + 2,
 * mirror (whole-sample) boundary, sum over the 16 taps with dimension 0 fastest in double; pixels whose index is outside
 * [-0.5, size - 0.5) take the default value 0; the result is cast to float.  An independent implementation of the same
 * published algorithm (SciPy's spline_filter / map_coordinates, mode 'mirror') agrees to rounding
 * (tests/test_oracle_metrics.py). */
static void bspline3_prefilter_line(double* c, int64_t n)
{
  if (n == 1)
    return;
  const double z = sqrt(3.0) - 2.0;
  const double c0 = (1.0 - z) * (1.0 - 1.0 / z);
  for (int64_t i = 0; i < n; ++i)
    c[i] *= c0;
  /* causal initialisation (mirror boundaries) */
  {
    double zn = z, sum;
    const int64_t horizon = (int64_t)ceil(log(1.0e-10) / log(fabs(z)));
    if (horizon < n)
    {
      sum = c[0];
      for (int64_t i = 1; i < horizon; ++i)
      {
        sum += zn * c[i];
        zn *= z;
      }
      c[0] = sum;
    }
    else
    {
      const double iz = 1.0 / z;
      double z2n = pow(z, (double)(n - 1));
      sum = c[0] + z2n * c[n - 1];
      z2n *= z2n * iz;
      for (int64_t i = 1; i <= n - 2; ++i)
      {
        sum += (zn + z2n) * c[i];
        zn *= z;
        z2n *= iz;
      }
      c[0] = sum / (1.0 - zn * zn);
    }
  }
  for (int64_t i = 1; i < n; ++i)
    c[i] += z * c[i - 1];
  c[n - 1] = (z / (z * z - 1.0)) * (z * c[n - 2] + c[n - 1]);
  for (int64_t i = n - 2; i >= 0; --i)
    c[i] = z * (c[i + 1] - c[i]);
}

static int64_t mirror_index(int64_t i, int64_t n)
{
  if (n == 1)
    return 0;
  if (i < 0)
    i = -i;
  if (i >= n)
    i = (n - 1) - (i - (n - 1));
  return i;
}

void xo_downsample_size(uint32_t rows, uint32_t cols, double factor, uint32_t* out_rows, uint32_t* out_cols)
{
  *out_cols = (uint32_t)(unsigned long)((double)cols * factor + 0.5);
  *out_rows = (uint32_t)(unsigned long)((double)rows * factor + 0.5);
}

void xo_downsample_image(const float* img, uint32_t rows, uint32_t cols, double factor, double sigma, float* out)
{
  uint32_t orows, ocols;
  xo_downsample_size(rows, cols, factor, &orows, &ocols);
  const size_t n = (size_t)rows * cols;
  float* sm = NULL;
  const float* src = img;
  if ((factor < 1.0) && (fabs(sigma) > 1.0e-6))   /* xregITKResampleUtils.h:68-86 */
  {
    const double s = (sigma < 0.0) ? (0.5 / factor) : sigma;
    sm = (float*)malloc(sizeof(float) * n);
    xo_itk_discrete_gaussian_2d(img, rows, cols, s * s, sm);
    src = sm;
  }
  double* c = (double*)malloc(sizeof(double) * n);
  for (size_t i = 0; i < n; ++i)
    c[i] = (double)src[i];
  for (int64_t y = 0; y < (int64_t)rows; ++y)   /* dimension 0 (x) first */
    bspline3_prefilter_line(c + (size_t)y * cols, (int64_t)cols);
  double* line = (double*)malloc(sizeof(double) * rows);
  for (int64_t x = 0; x < (int64_t)cols; ++x)
  {
    for (int64_t y = 0; y < (int64_t)rows; ++y)
      line[y] = c[(size_t)y * cols + x];
    bspline3_prefilter_line(line, (int64_t)rows);
    for (int64_t y = 0; y < (int64_t)rows; ++y)
      c[(size_t)y * cols + x] = line[y];
  }
  free(line);
  for (int64_t oy = 0; oy < (int64_t)orows; ++oy)
    for (int64_t ox = 0; ox < (int64_t)ocols; ++ox)
    {
      const double xs[2] = {(double)ox / factor, (double)oy / factor};
      const int64_t len[2] = {(int64_t)cols, (int64_t)rows};
      float v = 0.0f;
      if (xs[0] >= -0.5 && xs[0] < (double)cols - 0.5 && xs[1] >= -0.5 && xs[1] < (double)rows - 0.5)
      {
        double w[2][4];
        int64_t idx[2][4];
        for (int d = 0; d < 2; ++d)
        {
          const int64_t i0 = (int64_t)floor((float)xs[d]) - 1;
          const double t = xs[d] - (double)(i0 + 1);
          w[d][3] = (1.0 / 6.0) * t * t * t;
          w[d][0] = (1.0 / 6.0) + 0.5 * t * (t - 1.0) - w[d][3];
          w[d][2] = t + w[d][0] - 2.0 * w[d][3];
          w[d][1] = 1.0 - w[d][0] - w[d][2] - w[d][3];
          for (int k = 0; k < 4; ++k)
            idx[d][k] = mirror_index(i0 + k, len[d]);
        }
        double acc = 0.0;
        for (int p = 0; p < 16; ++p)   /* dimension 0 fastest */
        {
          const int kx = p & 3, ky = p >> 2;
          acc += (w[0][kx] * w[1][ky]) * c[(size_t)idx[1][ky] * cols + (size_t)idx[0][kx]];
        }
        v = (float)acc;
      }
      out[(size_t)oy * ocols + ox] = v;
    }
  free(c);
  free(sm);
}

/* HUToLinAttFilter::GenerateData (lib/image/xregHUToLinAtt.cpp:45-69; constants xregHUToLinAtt.h:73-76) */
void xo_hu_to_lin_att(const float* hu, float* att, uint64_t n, float hu_lower)
{
  const double mu_water = 0.02683 * 1.0, mu_air = 0.02485 * 0.0001;
  const double hu_scale = (mu_water - mu_air) * 1.0e-3;
  const double mu_lower = ((double)hu_lower * hu_scale) + mu_water;
  for (uint64_t i = 0; i < n; ++i)
  {
    const double a = (hu[i] * hu_scale) + mu_water - mu_lower;
    att[i] = (float)((a > 0.0) ? a : 0.0);
  }
}

/* ------------------------------------------------------------------------ */
/* NCC                                                                      */
/* ------------------------------------------------------------------------ */

/* Eigen 3.3 linear vectorised reduction (SSE Packet4f, two packet
 * accumulators, then predux (a0+a2)+(a1+a3), scalar tail), as used by
 * .mean() / .sum() / .dot() at xregImgSimMetric2DNCCCPU.cpp:58-61,72,182.
 * kind 0: sum a[i]; 1: sum (a[i]-c)^2; 2: sum a[i]*b[i]; 3: sum (a[i]-b[i])^2
 * (xregImgSimMetric2DSSDCPU.cpp:81).  Buffers are taken as aligned (alignedStart = 0). */
static float eigen_redux(const float* a, const float* b, float c, size_t n, int kind)
{
#define ELEM(i) ((kind == 0) ? a[i] : (kind == 1) ? ((a[i] - c) * (a[i] - c)) : (kind == 2) ? (a[i] * b[i]) : ((a[i] - b[i]) * (a[i] - b[i])))
  const size_t ps = 4;
  const size_t aligned2 = (n / (2 * ps)) * (2 * ps);
  const size_t aligned = (n / ps) * ps;
  float res;
  if (aligned)
  {
    float p0[4], p1[4];
    for (size_t l = 0; l < 4; ++l)
      p0[l] = ELEM(l);
    if (aligned > ps)
    {
      for (size_t l = 0; l < 4; ++l)
        p1[l] = ELEM(ps + l);
      for (size_t idx = 2 * ps; idx < aligned2; idx += 2 * ps)
      {
        for (size_t l = 0; l < 4; ++l)
        {
          p0[l] = p0[l] + ELEM(idx + l);
          p1[l] = p1[l] + ELEM(idx + ps + l);
        }
      }
      for (size_t l = 0; l < 4; ++l)
        p0[l] = p0[l] + p1[l];
      if (aligned > aligned2)
      {
        for (size_t l = 0; l < 4; ++l)
          p0[l] = p0[l] + ELEM(aligned2 + l);
      }
    }
    res = (p0[0] + p0[2]) + (p0[1] + p0[3]);
    for (size_t idx = aligned; idx < n; ++idx)
      res = res + ELEM(idx);
  }
  else
  {
    res = ELEM(0);
    for (size_t idx = 1; idx < n; ++idx)
      res = res + ELEM(idx);
  }
#undef ELEM
  return res;
}

/* ComputeImage2DMeanStdDev (:52-64) */
static void img_mean_std(const float* a, size_t n, float* mean, float* sd)
{
  const float mu = eigen_redux(a, NULL, 0.0f, n, 0) / (float)n;
  const float ss = eigen_redux(a, NULL, mu, n, 1);
  /* std::sqrt(float / size_t-int) : Scalar / Index -> float */
  const float s = sqrtf(ss / (float)(n - 1));
  *mean = mu;
  *sd = (s < 1.0e-6f) ? 1.0e-6f : s;
}

/* ComputeImage2DMeanStdDevWithMask (:78-115) */
static void img_mean_std_mask(const float* a, const uint8_t* mask, size_t n, size_t len,
                              float* mean, float* sd)
{
  float mu = 0.0f;
  for (size_t i = 0; i < n; ++i)
    if (mask[i])
      mu += a[i];
  mu /= (float)len;
  float s = 0.0f;
  for (size_t i = 0; i < n; ++i)
  {
    if (mask[i])
    {
      const float t = a[i] - mu;
      s += t * t;
    }
  }
  s = sqrtf(s / (float)(len - 1));
  *mean = mu;
  *sd = (s < 1.0e-6f) ? 1.0e-6f : s;
}

/* ImgSimMetric2DSSDCPU::compute / process_mask (xregImgSimMetric2DSSDCPU.cpp:62-110): fixed and moving
 * images zeroed outside the mask (the moving buffer IS modified), sum of squared differences / num pixels */
void xo_ssd(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols,
            float* mov, uint32_t n_imgs, float* sims, int n_threads)
{
  const size_t n = (size_t)rows * cols;
  float* f = (float*)malloc(sizeof(float) * n);
  memcpy(f, fixed, sizeof(float) * n);
  if (mask)
    for (size_t i = 0; i < n; ++i)
      if (!mask[i])
        f[i] = 0.0f;
  const int nt = resolve_threads(n_threads);
  (void)nt;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
  for (int64_t k = 0; k < (int64_t)n_imgs; ++k)
  {
    float* m = mov + (size_t)k * n;
    if (mask)
      for (size_t i = 0; i < n; ++i)
        if (!mask[i])
          m[i] = 0.0f;
    sims[k] = eigen_redux(f, m, 0.0f, n, 3) / (float)n; /* :81 */
  }
  free(f);
}

void xo_ncc(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols,
            float* mov, uint32_t n_imgs, float* sims, int n_threads)
{
  const size_t n = (size_t)rows * cols;
  float* f0 = (float*)malloc(sizeof(float) * n);
  memcpy(f0, fixed, sizeof(float) * n);
  float f_mean, f_sd;
  size_t mask_len = 0;
  /* process_mask (:211-236) */
  if (!mask)
  {
    img_mean_std(f0, n, &f_mean, &f_sd);
  }
  else
  {
    for (size_t i = 0; i < n; ++i)
      mask_len += mask[i] ? 1 : 0;
    img_mean_std_mask(f0, mask, n, mask_len, &f_mean, &f_sd);
  }
  for (size_t i = 0; i < n; ++i)
    f0[i] -= f_mean;

  const int nt = resolve_threads(n_threads);
  (void)nt;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nt)
  for (int64_t k = 0; k < (int64_t)n_imgs; ++k)
  {
    float* m = mov + (size_t)k * n;
    float m_mean, m_sd, sim;
    if (!mask)
    {
      img_mean_std(m, n, &m_mean, &m_sd);
      for (size_t i = 0; i < n; ++i)
        m[i] -= m_mean;
      /* size_t * float * float -> float (:182) */
      sim = eigen_redux(f0, m, 0.0f, n, 2) / (((float)n * f_sd) * m_sd);
    }
    else
    {
      img_mean_std_mask(m, mask, n, mask_len, &m_mean, &m_sd);
      for (size_t i = 0; i < n; ++i)
        m[i] -= m_mean;
      sim = 0.0f;
      for (size_t i = 0; i < n; ++i)
        if (mask[i])
          sim += f0[i] * m[i];
      sim /= ((float)mask_len * f_sd) * m_sd;
    }
    sims[k] = (1.0f - sim) * 0.5f; /* :205 */
  }
  free(f0);
}

/* ------------------------------------------------------------------------ */
/* Gaussian blur + Sobel (OpenCV 3.4.12, un-vendored; README.md:57-69)       */
/* ------------------------------------------------------------------------ */

static inline int reflect101(int i, int n)
{
  if (n == 1)
    return 0;
  while (i < 0 || i >= n)
  {
    if (i < 0)
      i = -i;
    else
      i = 2 * (n - 1) - i;
  }
  return i;
}

/* cv::getGaussianKernel(n, sigma <= 0, CV_32F): fixed tables for n = 1,3,5,7,
 * otherwise sigma = 0.3((n-1)*0.5 - 1) + 0.8 sampled and normalised. */
int xo_gauss_kernel(int width, float* cf)
{
  static const float t1[] = {1.f};
  static const float t3[] = {0.25f, 0.5f, 0.25f};
  static const float t5[] = {0.0625f, 0.25f, 0.375f, 0.25f, 0.0625f};
  static const float t7[] = {0.03125f, 0.109375f, 0.21875f, 0.28125f, 0.21875f, 0.109375f, 0.03125f};
  if (width <= 0 || !(width & 1))
    return -1;
  const float* tab = (width == 1) ? t1 : (width == 3) ? t3 : (width == 5) ? t5 : (width == 7) ? t7 : NULL;
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
    const double t = exp(scale2x * x * x);
    cf[i] = (float)t;
    sum += cf[i];
  }
  sum = 1.0 / sum;
  for (int i = 0; i < width; ++i)
    cf[i] = (float)(cf[i] * sum);
  return 0;
}

/* Separable symmetric filter, rows then columns, f32, evaluated in OpenCV's
 * symmetric form  k[c]*x0 + sum_j k[c+j]*(x[-j] + x[+j])  with reflect-101
 * borders (cv::GaussianBlur call sites xregImgSimMetric2DGradImgCPU.cpp:54-55,93-94). */
void xo_gauss_blur(const float* img, uint32_t rows, uint32_t cols, int width, float* out)
{
  const size_t n = (size_t)rows * cols;
  if (width <= 1)
  {
    memcpy(out, img, sizeof(float) * n);
    return;
  }
  float cf[64];
  if (width > 63 || xo_gauss_kernel(width, cf))
  {
    memcpy(out, img, sizeof(float) * n);
    return;
  }
  const int h = width / 2;
  float* tmp = (float*)malloc(sizeof(float) * n);
  for (int r = 0; r < (int)rows; ++r)
  {
    const float* src = img + (size_t)r * cols;
    for (int c = 0; c < (int)cols; ++c)
    {
      float s = cf[h] * src[c];
      for (int j = 1; j <= h; ++j)
      {
        s = s + cf[h + j] * (src[reflect101(c - j, (int)cols)] + src[reflect101(c + j, (int)cols)]);
      }
      tmp[(size_t)r * cols + c] = s;
    }
  }
  for (int r = 0; r < (int)rows; ++r)
  {
    for (int c = 0; c < (int)cols; ++c)
    {
      float s = cf[h] * tmp[(size_t)r * cols + c];
      for (int j = 1; j <= h; ++j)
      {
        s = s + cf[h + j] * (tmp[(size_t)reflect101(r - j, (int)rows) * cols + c] +
                             tmp[(size_t)reflect101(r + j, (int)rows) * cols + c]);
      }
      out[(size_t)r * cols + c] = s;
    }
  }
  free(tmp);
}

/* cv::Sobel(ddepth -1, dx/dy, ksize 3, scale 1, delta 0, BORDER_REFLECT_101)
 * (xregImgSimMetric2DGradImgCPU.cpp:62-65,97-100) */
void xo_sobel(const float* img, uint32_t rows, uint32_t cols, float* gx, float* gy)
{
  const int R = (int)rows, C = (int)cols;
#define P(r, c) img[(size_t)reflect101((r), R) * cols + reflect101((c), C)]
  for (int r = 0; r < R; ++r)
  {
    for (int c = 0; c < C; ++c)
    {
      /* dx: row pass [-1 0 1], column pass [1 2 1] */
      const float dm = P(r - 1, c + 1) - P(r - 1, c - 1);
      const float d0 = P(r, c + 1) - P(r, c - 1);
      const float dp = P(r + 1, c + 1) - P(r + 1, c - 1);
      gx[(size_t)r * cols + c] = (dm + dp) + 2.0f * d0;
      /* dy: row pass [1 2 1], column pass [-1 0 1] */
      const float sm = (P(r - 1, c - 1) + P(r - 1, c + 1)) + 2.0f * P(r - 1, c);
      const float sp = (P(r + 1, c - 1) + P(r + 1, c + 1)) + 2.0f * P(r + 1, c);
      gy[(size_t)r * cols + c] = sp - sm;
    }
  }
#undef P
}

void xo_grad_imgs(const float* img, uint32_t rows, uint32_t cols, int gauss_width,
                  float* gx, float* gy)
{
  if (gauss_width)
  {
    float* tmp = (float*)malloc(sizeof(float) * (size_t)rows * cols);
    xo_gauss_blur(img, rows, cols, gauss_width, tmp);
    xo_sobel(tmp, rows, cols, gx, gy);
    free(tmp);
  }
  else
  {
    xo_sobel(img, rows, cols, gx, gy);
  }
}

/* lib/regi/sim_metrics_2d/xregImgSimMetric2DGradNCCCPU.cpp:29-65 */
void xo_grad_ncc(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols,
                 int gauss_width, const float* mov, uint32_t n_imgs, float* sims,
                 int n_threads)
{
  const size_t n = (size_t)rows * cols;
  float* fgx = (float*)malloc(sizeof(float) * n);
  float* fgy = (float*)malloc(sizeof(float) * n);
  float* mgx = (float*)malloc(sizeof(float) * n * n_imgs);
  float* mgy = (float*)malloc(sizeof(float) * n * n_imgs);
  float* sx = (float*)malloc(sizeof(float) * n_imgs);
  float* sy = (float*)malloc(sizeof(float) * n_imgs);
  xo_grad_imgs(fixed, rows, cols, gauss_width, fgx, fgy);
  /* serial loop over images as at xregImgSimMetric2DGradImgCPU.cpp:86 */
  for (uint32_t k = 0; k < n_imgs; ++k)
    xo_grad_imgs(mov + (size_t)k * n, rows, cols, gauss_width, mgx + (size_t)k * n, mgy + (size_t)k * n);
  xo_ncc(fgx, mask, rows, cols, mgx, n_imgs, sx, n_threads);
  xo_ncc(fgy, mask, rows, cols, mgy, n_imgs, sy, n_threads);
  for (uint32_t k = 0; k < n_imgs; ++k)
    sims[k] = (float)(0.5 * (sx[k] + sy[k])); /* :61-62, 0.5 is a double literal */
  free(fgx);
  free(fgy);
  free(mgx);
  free(mgy);
  free(sx);
  free(sy);
}

/* ------------------------------------------------------------------------ */
/* Patch NCC                                                                */
/* ------------------------------------------------------------------------ */

static uint32_t n_centres(uint32_t dim, uint32_t r, uint32_t stride)
{
  if (2 * r + 1 > dim)
    return 0;
  return (dim - 1 - 2 * r) / stride + 1;
}

uint64_t xo_num_patches(uint32_t rows, uint32_t cols, uint32_t radius, uint32_t stride)
{
  return (uint64_t)n_centres(rows, radius, stride) * n_centres(cols, radius, stride);
}

/* xregImgSimMetric2DPatchCommon.cpp:309-410 (random-patch distribution excluded) */
void xo_patch_weights(uint32_t rows, uint32_t cols, const xo_patch_opts* o,
                      const uint8_t* mask, const float* wgt_img, float* weights)
{
  const uint32_t r = o->radius, st = o->stride, d = 2 * r + 1;
  const uint32_t ncr = n_centres(rows, r, st), ncc = n_centres(cols, r, st);
  const size_t np = (size_t)ncr * ncc;
  for (size_t k = 0; k < np; ++k)
    weights[k] = 1.0f;
  const int use_mask_w = o->use_mask_for_weighting && mask;
  if (!wgt_img && !use_mask_w)
    return;
  size_t k = 0;
  for (uint32_t i = 0; i < ncr; ++i)
  {
    for (uint32_t j = 0; j < ncc; ++j, ++k)
    {
      const uint32_t cr = r + i * st, cc = r + j * st;
      if (wgt_img)
      {
        weights[k] = wgt_img[(size_t)cr * cols + cc];
        if (use_mask_w && !mask[(size_t)cr * cols + cc])
          weights[k] = 0.0f;
      }
      else
      {
        size_t cnt = 0;
        for (uint32_t pr = cr - r; pr <= cr + r; ++pr)
          for (uint32_t pc = cc - r; pc <= cc + r; ++pc)
            cnt += mask[(size_t)pr * cols + pc] ? 1 : 0;
        weights[k] = (float)cnt / (float)((size_t)d * d);
      }
    }
  }
  if (o->normalize_weights_as_prob)
  {
    float ws = 0.0f;
    for (k = 0; k < np; ++k)
      ws += weights[k];
    for (k = 0; k < np; ++k)
      weights[k] /= ws;
  }
}

/* detail::ComputePatchMeanStdDev (xregImgSimMetric2DPatchNCCCPU.cpp:558-619) */
static void patch_mean_std(const float* img, const uint8_t* mask, uint32_t cols,
                           uint32_t r0, uint32_t c0, uint32_t d, int use_mask_for_stats,
                           float* mean_out, float* sd_out, size_t* n_out)
{
  size_t cnt = 0;
  float mean = 0.0f;
  for (uint32_t r = 0; r < d; ++r)
  {
    const float* row = img + (size_t)(r0 + r) * cols + c0;
    for (uint32_t c = 0; c < d; ++c)
    {
      if (!use_mask_for_stats || (!mask || mask[(size_t)(r0 + r) * cols + c0 + c]))
      {
        mean += row[c];
        ++cnt;
      }
    }
  }
  float var = 0.0f;
  if (cnt > 1)
  {
    mean /= (float)cnt;
    for (uint32_t r = 0; r < d; ++r)
    {
      const float* row = img + (size_t)(r0 + r) * cols + c0;
      for (uint32_t c = 0; c < d; ++c)
      {
        if (!use_mask_for_stats || (!mask || mask[(size_t)(r0 + r) * cols + c0 + c]))
        {
          const float t = row[c] - mean;
          var += t * t;
        }
      }
    }
    var /= (float)cnt - 1.0f;
  }
  const float s = sqrtf(var);
  *mean_out = mean;
  *sd_out = (s < 1.0e-6f) ? 1.0e-6f : s;
  *n_out = cnt;
}

void xo_patch_ncc(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols,
                  const xo_patch_opts* o, const float* weights,
                  const float* mov, uint32_t n_imgs, float* sims, float* patch_sims,
                  int n_threads)
{
  const uint32_t r = o->radius, st = o->stride, d = 2 * r + 1;
  const uint32_t ncr = n_centres(rows, r, st), ncc = n_centres(cols, r, st);
  const size_t np = (size_t)ncr * ncc;
  const size_t n = (size_t)rows * cols;
  const int nt = resolve_threads(n_threads);
  (void)nt;

  /* fixed patch statistics (:398-410).  The reference materialises
   * F'_k[p] = (f[p] - mu_fk) / (sigma_fk * n_k); only mu, sigma*n are kept
   * here and F' is re-evaluated with the identical f32 expression on use. */
  float* f_mean = (float*)malloc(sizeof(float) * np);
  float* f_den = (float*)malloc(sizeof(float) * np);
#pragma omp parallel for schedule(static) num_threads(nt)
  for (int64_t k = 0; k < (int64_t)np; ++k)
  {
    const uint32_t r0 = (uint32_t)(k / ncc) * st, c0 = (uint32_t)(k % ncc) * st;
    float mu, sd;
    size_t cnt;
    patch_mean_std(fixed, mask, cols, r0, c0, d, o->use_mask_for_patch_stats, &mu, &sd, &cnt);
    f_mean[k] = mu;
    f_den[k] = sd * (float)cnt; /* tmp_std_dev * tmp_num_patch_elems_for_stats */
  }

  float* vals = (float*)malloc(sizeof(float) * np);
  for (uint32_t mi = 0; mi < n_imgs; ++mi) /* serial over images (:103) */
  {
    const float* m = mov + (size_t)mi * n;
#pragma omp parallel for schedule(dynamic, 256) num_threads(nt)
    for (int64_t k = 0; k < (int64_t)np; ++k)
    {
      vals[k] = 0.0f;
      const float w = weights ? weights[k] : 1.0f;
      if (!o->weight_patch_sims || (fabsf(w) > 1.0e-6f)) /* :214 */
      {
        const uint32_t r0 = (uint32_t)(k / ncc) * st, c0 = (uint32_t)(k % ncc) * st;
        float mu, sd;
        size_t cnt;
        patch_mean_std(m, mask, cols, r0, c0, d, o->use_mask_for_patch_stats, &mu, &sd, &cnt);
        float acc = 0.0f;
        for (uint32_t pr = 0; pr < d; ++pr)
        {
          const size_t base = (size_t)(r0 + pr) * cols + c0;
          for (uint32_t pc = 0; pc < d; ++pc)
          {
            if (!mask || mask[base + pc])
            {
              const float fp = (fixed[base + pc] - f_mean[k]) / f_den[k];
              acc += ((m[base + pc] - mu) / sd) * fp; /* :241 */
            }
          }
        }
        const float s = 1.0f - acc;
        if (patch_sims)
          patch_sims[(size_t)mi * np + k] = s;
        vals[k] = (o->weight_patch_sims ? w : 1.0f) * s;
      }
      else if (patch_sims)
      {
        patch_sims[(size_t)mi * np + k] = 0.0f;
      }
    }
    float sum = 0.0f;
    for (size_t k = 0; k < np; ++k)
      sum += vals[k]; /* :262-266 */
    if (o->compute_mean_of_patch_sims)
    {
      sum /= (float)np;
    }
    else if (o->weight_patch_sims)
    {
      float tw = 0.0f;
      for (size_t k = 0; k < np; ++k)
        tw += weights ? weights[k] : 1.0f;
      sum /= tw;
    }
    sims[mi] = sum;
  }
  free(vals);
  free(f_mean);
  free(f_den);
}

/* Patch subsets: ImgSimMetric2DPatchCommon::set_patches_to_use / the random patches of patch_indices_to_use
 * (xregImgSimMetric2DPatchCommon.cpp:231-241, 413-493).  ImgSimMetric2DPatchNCCCPU::compute then loops over the LOCAL
 * index list patch_inds_to_use_ (:97-101, :204): the value of local patch j is that of global patch subset[j]
 * (weights are those of the whole grid), the sequential f32 sum runs in subset order (:262-266), the mean divides by
 * the subset size (:268-271, num_patches() = patch_inds_to_use_.size()) and the weighted combine by the sequential f32
 * sum of the subset's weights (:272-285).  Indices may repeat. */
void xo_patch_ncc_subset(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols,
                         const xo_patch_opts* o, const float* weights, const uint64_t* subset, uint64_t n_subset,
                         const float* mov, uint32_t n_imgs, float* sims, int n_threads)
{
  const size_t np = (size_t)n_centres(rows, o->radius, o->stride) * n_centres(cols, o->radius, o->stride);
  float* ps = (float*)malloc(sizeof(float) * np * (n_imgs ? n_imgs : 1));
  float* all = (float*)malloc(sizeof(float) * (n_imgs ? n_imgs : 1));
  xo_patch_ncc(fixed, mask, rows, cols, o, weights, mov, n_imgs, all, ps, n_threads);
  for (uint32_t mi = 0; mi < n_imgs; ++mi)
  {
    float sum = 0.0f;
    for (uint64_t j = 0; j < n_subset; ++j)
    {
      const size_t k = (size_t)subset[j];
      const float w = weights ? weights[k] : 1.0f;
      float v = 0.0f;
      if (!o->weight_patch_sims || (fabsf(w) > 1.0e-6f))
        v = (o->weight_patch_sims ? w : 1.0f) * ps[(size_t)mi * np + k];
      sum += v;
    }
    if (o->compute_mean_of_patch_sims)
    {
      sum /= (float)n_subset;
    }
    else if (o->weight_patch_sims)
    {
      float tw = 0.0f;
      for (uint64_t j = 0; j < n_subset; ++j)
        tw += weights ? weights[subset[j]] : 1.0f;
      sum /= tw;
    }
    sims[mi] = sum;
  }
  free(ps);
  free(all);
}

/* xregImgSimMetric2DPatchGradNCCCPU.cpp:126-135: the same subset for both directions */
void xo_patch_grad_ncc_subset(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols, int gauss_width,
                              const xo_patch_opts* o, const float* weights, const uint64_t* subset, uint64_t n_subset,
                              const float* mov, uint32_t n_imgs, float* sims, int n_threads)
{
  const size_t n = (size_t)rows * cols;
  float* fgx = (float*)malloc(sizeof(float) * n);
  float* fgy = (float*)malloc(sizeof(float) * n);
  float* mgx = (float*)malloc(sizeof(float) * n * n_imgs);
  float* mgy = (float*)malloc(sizeof(float) * n * n_imgs);
  float* sx = (float*)malloc(sizeof(float) * n_imgs);
  float* sy = (float*)malloc(sizeof(float) * n_imgs);
  xo_grad_imgs(fixed, rows, cols, gauss_width, fgx, fgy);
  for (uint32_t k = 0; k < n_imgs; ++k)
    xo_grad_imgs(mov + (size_t)k * n, rows, cols, gauss_width, mgx + (size_t)k * n, mgy + (size_t)k * n);
  xo_patch_ncc_subset(fgx, mask, rows, cols, o, weights, subset, n_subset, mgx, n_imgs, sx, n_threads);
  xo_patch_ncc_subset(fgy, mask, rows, cols, o, weights, subset, n_subset, mgy, n_imgs, sy, n_threads);
  for (uint32_t k = 0; k < n_imgs; ++k)
    sims[k] = (float)(0.5 * (sx[k] + sy[k]));
  free(fgx);
  free(fgy);
  free(mgx);
  free(mgy);
  free(sx);
  free(sy);
}

/* lib/regi/sim_metrics_2d/xregImgSimMetric2DPatchGradNCCCPU.cpp:34-253 */
void xo_patch_grad_ncc(const float* fixed, const uint8_t* mask, uint32_t rows, uint32_t cols,
                       int gauss_width, const xo_patch_opts* o, const float* weights,
                       const float* mov, uint32_t n_imgs, float* sims, int n_threads)
{
  const size_t n = (size_t)rows * cols;
  float* fgx = (float*)malloc(sizeof(float) * n);
  float* fgy = (float*)malloc(sizeof(float) * n);
  float* mgx = (float*)malloc(sizeof(float) * n * n_imgs);
  float* mgy = (float*)malloc(sizeof(float) * n * n_imgs);
  float* sx = (float*)malloc(sizeof(float) * n_imgs);
  float* sy = (float*)malloc(sizeof(float) * n_imgs);
  xo_grad_imgs(fixed, rows, cols, gauss_width, fgx, fgy);
  for (uint32_t k = 0; k < n_imgs; ++k)
    xo_grad_imgs(mov + (size_t)k * n, rows, cols, gauss_width, mgx + (size_t)k * n, mgy + (size_t)k * n);
  xo_patch_ncc(fgx, mask, rows, cols, o, weights, mgx, n_imgs, sx, NULL, n_threads);
  xo_patch_ncc(fgy, mask, rows, cols, o, weights, mgy, n_imgs, sy, NULL, n_threads);
  for (uint32_t k = 0; k < n_imgs; ++k)
    sims[k] = (float)(0.5 * (sx[k] + sy[k])); /* :222 */
  free(fgx);
  free(fgy);
  free(mgx);
  free(mgy);
  free(sx);
  free(sy);
}

/* lib/regi/sim_metrics_2d/xregImgSimMetric2DCombine.cpp:67-86 */
void xo_combine_mean(const float* view_sims, uint32_t n_views, uint32_t n_poses, float* out)
{
  for (uint32_t p = 0; p < n_poses; ++p)
    out[p] = 0.0f;
  for (uint32_t v = 0; v < n_views; ++v)
    for (uint32_t p = 0; p < n_poses; ++p)
      out[p] += view_sims[(size_t)v * n_poses + p];
  for (uint32_t p = 0; p < n_poses; ++p)
    out[p] /= (float)n_views;
}
