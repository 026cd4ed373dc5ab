// C++ parity test of the host mirror (include/xreg_cuda.hpp) against the CPU oracle (oracle/xreg_oracle.h).
//
// Written the way a test inside an xReg checkout would drive the reference interfaces: build a scene, create a
// RayCaster and ImgSimMetric2D objects, follow the call-order contract, compare with the CPU classes'
// arithmetic -- here restated by the oracle, which this TEST links (the product library never does).
//
//   host_mirror_test --no-gpu   host logic only: CameraModel set-up, transform algebra, patch grid / weights, error types
//   host_mirror_test            whole path on cuda:0: DRR parity (bit-exact clip masks and sample counts, <= 1e-4 relative),
//                               the five metrics (<= 1e-5 absolute), multi-view distribution + CombineMean, the one-call
//                               objective, store methods / background projections, error behaviour
// Exit code 0 = all checks passed.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "xreg_cuda.hpp"
#include "xreg_oracle.h"

using namespace xreg_b200;

static int g_failed = 0;

#define CHECK(cond)                                                          \
  do                                                                         \
  {                                                                          \
    if (!(cond))                                                             \
    {                                                                        \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      ++g_failed;                                                            \
    }                                                                        \
  } while (0)

template <class E>
static bool Throws(const std::function<void()>& f)
{
  try
  {
    f();
  }
  catch (const E&)
  {
    return true;
  }
  catch (...)
  {
    return false;
  }
  return false;
}

// deterministic LCG (no <random>: identical streams on every libstdc++)
struct Rng
{
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed * 6364136223846793005ull + 1442695040888963407ull) {}
  double uniform()
  {
    s = s * 6364136223846793005ull + 1442695040888963407ull;
    return static_cast<double>((s >> 11) & ((1ull << 53) - 1)) / static_cast<double>(1ull << 53);
  }
  double normal()
  {
    const double u1 = uniform() + 1e-300, u2 = uniform();
    return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586 * u2);
  }
};

struct Scene
{
  std::vector<float> voxels;
  Volume vol;
  CameraModel cam;
  FrameTransform nominal;
};

// ellipsoidal "body" (0.02 + noise) in air with a few denser blobs; anisotropic spacing; centred on the origin
static Scene MakeScene(const int nx, const int ny, const int nz, const size_type det_rows, const size_type det_cols)
{
  Scene s;
  s.voxels.assign(static_cast<size_t>(nx) * ny * nz, 0.0f);
  Rng rng(20211009);
  struct Blob
  {
    double c[3], r[3];
  };
  std::vector<Blob> blobs(6);
  const double dims[3] = {double(nx), double(ny), double(nz)};
  for (auto& b : blobs)
  {
    for (int k = 0; k < 3; ++k)
    {
      b.c[k] = (0.3 + 0.4 * rng.uniform()) * dims[k];
      b.r[k] = (0.05 + 0.1 * rng.uniform()) * dims[k];
    }
  }
  for (int z = 0; z < nz; ++z)
    for (int y = 0; y < ny; ++y)
      for (int x = 0; x < nx; ++x)
      {
        const double p[3] = {double(x), double(y), double(z)};
        double e = 0;
        for (int k = 0; k < 3; ++k)
        {
          const double t = (p[k] - 0.5 * (dims[k] - 1)) / (0.45 * dims[k]);
          e += t * t;
        }
        float v = 0;
        if (e <= 1.0)
        {
          v = 0.02f + 0.005f * static_cast<float>(rng.uniform());
          for (const auto& b : blobs)
          {
            double q = 0;
            for (int k = 0; k < 3; ++k)
            {
              const double t = (p[k] - b.c[k]) / b.r[k];
              q += t * t;
            }
            if (q <= 1.0)
            {
              v = 0.05f;
            }
          }
        }
        s.voxels[(static_cast<size_t>(z) * ny + y) * nx + x] = v;
      }
  s.vol.data = s.voxels.data();
  s.vol.size[0] = nx;
  s.vol.size[1] = ny;
  s.vol.size[2] = nz;
  s.vol.spacing[0] = 0.9;
  s.vol.spacing[1] = 1.1;
  s.vol.spacing[2] = 1.3;
  for (int k = 0; k < 3; ++k)
  {
    s.vol.origin[k] = -0.5 * (dims[k] - 1.0) * s.vol.spacing[k];
  }
  s.cam.coord_frame_type = CameraModel::kORIGIN_AT_FOCAL_PT_DET_NEG_Z;
  s.cam.setup(400.0f, det_rows, det_cols, 1.6f, 1.5f);
  // camera looks along +y of the volume, detector rows along -z, volume centre 250 mm from the source
  const float R[9] = {1, 0, 0, 0, 0, 1, 0, -1, 0};
  for (int r = 0; r < 3; ++r)
  {
    for (int c = 0; c < 3; ++c)
    {
      s.nominal(r, c) = R[3 * r + c];
    }
    s.nominal(r, 3) = -(R[3 * r + 2] * -250.0f);
  }
  return s;
}

// nominal pose perturbed about the volume centre (the origin): exp(x) * nominal
static FrameTransformList MakePoses(const Scene& s, const size_type n, const uint64_t seed, const double rot_deg = 5.0,
                                    const double trans_mm = 5.0)
{
  Rng rng(seed);
  FrameTransformList out;
  for (size_type i = 0; i < n; ++i)
  {
    float x[6];
    for (int k = 0; k < 3; ++k)
    {
      x[k] = static_cast<float>(rng.normal() * rot_deg * 3.141592653589793 / 180.0);
      x[3 + k] = static_cast<float>(rng.normal() * trans_mm);
    }
    out.push_back(ExpSE3(x) * s.nominal);
  }
  return out;
}

static xo_cam ToOracleCam(const CameraModel& c)
{
  const xrc_cam x = c.to_xrc();
  xo_cam o;
  static_assert(sizeof(xo_cam) == sizeof(xrc_cam), "POD layouts agree");
  std::memcpy(&o, &x, sizeof(o));
  return o;
}

struct OracleDrr
{
  std::vector<float> img;
  std::vector<uint8_t> mask;
  std::vector<uint32_t> steps;
  uint64_t total = 0;
};

static OracleDrr OracleRayCast(const Scene& s, const std::vector<CameraModel>& cams, const FrameTransformList& xforms,
                               const std::vector<uint32_t>& cam_idx, const float step = 1.0f, const int kernel = XO_KERNEL_SUM,
                               const int store = XO_STORE_REPLACE, const float bg = 0.0f, const float* const* bg_projs = nullptr,
                               const std::vector<float>* prev = nullptr)
{
  const size_type n = xforms.size(), npix = cams[0].num_det_rows * cams[0].num_det_cols;
  std::vector<xo_cam> ocams;
  for (const auto& c : cams)
  {
    ocams.push_back(ToOracleCam(c));
  }
  std::vector<float> poses(12 * n);
  for (size_type i = 0; i < n; ++i)
  {
    xforms[i].to3x4(&poses[12 * i]);
  }
  float i2p[12];
  s.vol.idx_to_phys(i2p);
  OracleDrr o;
  o.img = prev ? *prev : std::vector<float>(n * npix, 0.0f);
  o.mask.resize(n * npix);
  o.steps.resize(n * npix);
  xo_pre_compute(o.img.data(), static_cast<uint32_t>(n), static_cast<uint32_t>(cams[0].num_det_rows),
                 static_cast<uint32_t>(cams[0].num_det_cols), cam_idx.data(), bg_projs, store, bg);
  const int rc = xo_drr(s.vol.data, s.vol.size, i2p, ocams.data(), static_cast<uint32_t>(ocams.size()), poses.data(), cam_idx.data(),
                        static_cast<uint32_t>(n), step, kernel, o.img.data(), o.mask.data(), o.steps.data(), &o.total, 0);
  CHECK(rc == 0);
  return o;
}

// north_star tolerances: clip masks and indexing bit-exact, per-pixel DRR relative error <= 1e-4
static double CheckDrr(const float* got, const OracleDrr& ref)
{
  double worst = 0;
  bool exact_outside = true;
  for (size_t i = 0; i < ref.img.size(); ++i)
  {
    if (!ref.mask[i])
    {
      exact_outside = exact_outside && (got[i] == ref.img[i]);
    }
    else if (ref.img[i] != 0.0f)
    {
      worst = std::fmax(worst, std::fabs(double(got[i]) - double(ref.img[i])) / std::fabs(double(ref.img[i])));
    }
  }
  CHECK(exact_outside);
  CHECK(worst <= 1.0e-4);
  return worst;
}

static double MaxAbsDiff(const std::vector<float>& a, const std::vector<float>& b)
{
  CHECK(a.size() == b.size());
  double w = 0;
  for (size_t i = 0; i < a.size() && i < b.size(); ++i)
  {
    w = std::fmax(w, std::fabs(double(a[i]) - double(b[i])));
  }
  return w;
}

static const double kSIM_TOL = 1.0e-5;  // north_star: similarity values within 1e-5 absolute

// ---------------------------------------------------------------------------------------------------------------
static void TestHostLogic()
{
  // CameraModel::setup against the oracle's restatement of xregPerspectiveXform.cpp:200-254,302-334 (bit-equal)
  for (int frame = 0; frame < 3; ++frame)
  {
    CameraModel c;
    c.coord_frame_type = static_cast<CameraModel::CameraCoordFrame>(frame);
    c.setup(1020.0f, 480, 472, 0.62f, 0.61f);
    xo_cam o;
    xo_cam_setup_naive(&o, 1020.0f, 480, 472, 0.62f, 0.61f, frame);
    const xo_cam m = ToOracleCam(c);
    CHECK(std::memcmp(&m, &o, sizeof(o)) == 0);

    // general set-up: rotated / translated extrinsic
    const float x[6] = {0.1f, -0.6f, 0.25f, 12.0f, -7.5f, 30.0f};
    const FrameTransform ext = ExpSE3(x);
    CameraModel g;
    g.coord_frame_type = c.coord_frame_type;
    g.setup(c.intrins, ext, 480, 472, 0.62f, 0.61f);
    xo_cam og;
    xo_cam_setup(&og, c.intrins, ext.m, 480, 472, 0.62f, 0.61f, frame);
    const xo_cam mg = ToOracleCam(g);
    CHECK(std::memcmp(&mg, &og, sizeof(og)) == 0);
  }
  // transform algebra against the oracle's frozen f32 order
  {
    const float xa[6] = {0.3f, 0.2f, -0.4f, 5.0f, 6.0f, -7.0f}, xb[6] = {-0.1f, 0.5f, 0.2f, -1.0f, 2.5f, 3.0f};
    const FrameTransform A = ExpSE3(xa), B = ExpSE3(xb);
    float a[12], b[12], ref[12], got[12];
    A.to3x4(a);
    B.to3x4(b);
    xo_affine_compose(a, b, ref);
    (A * B).to3x4(got);
    CHECK(std::memcmp(ref, got, sizeof(ref)) == 0);
    // rigid inverse really inverts
    (A * A.rigid_inverse()).to3x4(got);
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 4; ++c)
        CHECK(std::fabs(got[4 * r + c] - (r == c ? 1.0f : 0.0f)) < 1e-5f);
  }
  // Volume::idx_to_phys: double metadata narrowed once (xregITKBasicImageUtils.h:157-165)
  {
    Volume v;
    v.spacing[0] = 0.8;
    v.spacing[1] = 0.7;
    v.spacing[2] = 1.25;
    v.origin[0] = -100.1;
    v.origin[1] = 3.3;
    v.origin[2] = 7.0;
    const double c30 = std::cos(0.5235987755982988), s30 = std::sin(0.5235987755982988);
    const double D[9] = {c30, -s30, 0, s30, c30, 0, 0, 0, 1};
    std::memcpy(v.direction, D, sizeof(D));
    float t[12];
    v.idx_to_phys(t);
    CHECK(t[0] == static_cast<float>(c30 * 0.8) && t[1] == static_cast<float>(-s30 * 0.7) && t[10] == 1.25f);
    CHECK(t[3] == static_cast<float>(-100.1) && t[11] == 7.0f);
  }
  // patch grid and weights against the oracle's restatement of xregImgSimMetric2DPatchCommon.cpp:256-410
  {
    struct Probe : ImgSimMetric2DPatchCommon
    {
    } p;
    const size_type rows = 41, cols = 37;
    std::vector<uint8_t> mask(rows * cols);
    std::vector<float> wimg(rows * cols);
    Rng rng(7);
    for (size_t i = 0; i < mask.size(); ++i)
    {
      mask[i] = rng.uniform() < 0.7 ? 255 : 0;
      wimg[i] = static_cast<float>(rng.uniform());
    }
    for (const size_type radius : {size_type(1), size_type(3), size_type(6)})
      for (const size_type stride : {size_type(1), size_type(2), size_type(5)})
        for (int mode = 0; mode < 4; ++mode)  // bit 0: weight image, bit 1: normalise
        {
          p.set_patch_radius(radius);
          p.set_patch_stride(stride);
          p.set_normalize_weights_as_prob((mode & 2) != 0);
          p.set_wgt_img((mode & 1) ? ImgSimMetric2DPatchCommon::WgtImg(wimg.data(), rows, cols) : ImgSimMetric2DPatchCommon::WgtImg());
          CHECK(ImgSimMetric2DPatchCommon::NumPatches(rows, cols, radius, stride) ==
                xo_num_patches(rows, cols, static_cast<uint32_t>(radius), static_cast<uint32_t>(stride)));
          std::vector<float> got;
          CHECK(p.compute_weights(rows, cols, mask.data(), &got));
          xo_patch_opts o = {static_cast<uint32_t>(radius), static_cast<uint32_t>(stride), 0, 1, 1, 0, (mode & 2) ? 1 : 0};
          std::vector<float> ref(xo_num_patches(rows, cols, o.radius, o.stride));
          xo_patch_weights(rows, cols, &o, mask.data(), (mode & 1) ? wimg.data() : nullptr, ref.data());
          CHECK(got.size() == ref.size());
          CHECK(got.size() == ref.size() && std::memcmp(got.data(), ref.data(), ref.size() * sizeof(float)) == 0);
        }
    // no mask, no weight image: weights stay 1 -> nothing to hand down
    std::vector<float> none;
    p.set_wgt_img(ImgSimMetric2DPatchCommon::WgtImg());
    CHECK(!p.compute_weights(rows, cols, nullptr, &none));
    CHECK(Throws<UnsupportedOperationException>([&] { p.set_choose_rand_patches(true); }));
  }
  // CameraModel::setup rejects nonsense like the reference's asserts
  CHECK(Throws<XregCudaError>([] {
    CameraModel c;
    c.setup(0.0f, 10, 10, 1.0f, 1.0f);
  }));
}

// ---------------------------------------------------------------------------------------------------------------
static void TestRayCaster(Context& ctx, const Scene& s)
{
  const size_type rows = s.cam.num_det_rows, cols = s.cam.num_det_cols, npix = rows * cols;
  RayCasterLineIntCUDA ray_caster(ctx);
  RayCaster& rc = ray_caster;  // drive it through the reference's base interface

  // call-order contract: compute() before allocate_resources() is an assertion failure (xregRayCastLineIntCPU.cpp:296)
  CHECK(Throws<XregCudaError>([&] { rc.compute(); }));

  rc.set_volume(s.vol);
  rc.set_camera_model(s.cam);
  rc.set_num_projs(6);
  rc.allocate_resources();
  CHECK(rc.max_num_projs() == 6 && rc.num_projs() == 6);
  CHECK(rc.max_num_projs_possible() > 6);

  const FrameTransformList poses = MakePoses(s, 6, 1, 10.0, 6.0);
  rc.set_xforms_cam_to_itk_phys(poses);
  rc.compute();

  const std::vector<uint32_t> cam0(6, 0);
  const OracleDrr ref = OracleRayCast(s, {s.cam}, poses, cam0);
  std::vector<uint8_t> mask(6 * npix);
  std::vector<uint32_t> steps(6 * npix);
  const uint64_t total = ray_caster.ray_info(mask.data(), steps.data());
  CHECK(mask == ref.mask);    // bit-exact ray / volume-box intersection masks
  CHECK(steps == ref.steps);  // identical sample counts per ray
  CHECK(total == ref.total);
  const double e = CheckDrr(rc.raw_host_pixel_buf(), ref);
  std::printf("  drr: 6 poses, %zu x %zu, S = %llu, max rel err %.3g\n", rows, cols, (unsigned long long)total, e);

  // proj(i): view into the host buffer, image-major (xregRayCastBaseCPU.h:149-154)
  const RayCaster::Proj p3 = rc.proj(3);
  CHECK(p3.rows == rows && p3.cols == cols && p3.data == rc.raw_host_pixel_buf() + 3 * npix);

  // fewer projections after allocation (xregRayCastInterface.cpp:131-139); more than the capacity asserts
  rc.set_num_projs(2);
  rc.set_xforms_cam_to_itk_phys({poses[4], poses[1]});
  rc.compute();
  {
    const OracleDrr r2 = OracleRayCast(s, {s.cam}, {poses[4], poses[1]}, {0, 0});
    CheckDrr(rc.raw_host_pixel_buf(), r2);
  }
  CHECK(Throws<XregCudaError>([&] { rc.set_num_projs(7); }));
  rc.set_num_projs(6);
  rc.set_xforms_cam_to_itk_phys(poses);

  // mutable pose reference + pre/post multiplication
  {
    const float dx[6] = {0, 0, 0.05f, 2.0f, 0, 0};
    const FrameTransform d = ExpSE3(dx);
    rc.xform_cam_to_itk_phys(2) = d * poses[2];
    rc.post_multiply_all_xforms(d);
    FrameTransformList expect = poses;
    expect[2] = d * poses[2];
    for (auto& t : expect)
    {
      t = t * d;
    }
    rc.compute();
    const OracleDrr r3 = OracleRayCast(s, {s.cam}, expect, cam0);
    CheckDrr(rc.raw_host_pixel_buf(), r3);
    rc.set_xforms_cam_to_itk_phys(poses);
  }

  // step size, max kernel, default background value
  {
    rc.set_ray_step_size(0.5f);
    ray_caster.set_kernel_id(kRAY_CAST_LINE_INT_MAX_KERNEL);
    rc.set_default_bg_pixel_val(0.25f);
    rc.compute();
    const OracleDrr r4 = OracleRayCast(s, {s.cam}, poses, cam0, 0.5f, XO_KERNEL_MAX, XO_STORE_REPLACE, 0.25f);
    CheckDrr(rc.raw_host_pixel_buf(), r4);
    rc.set_ray_step_size(1.0f);
    ray_caster.set_kernel_id(kRAY_CAST_LINE_INT_SUM_KERNEL);
    rc.set_default_bg_pixel_val(0.0f);
  }

  // ACCUM on top of the previous projections; background projections (RayCasterCPU::pre_compute)
  {
    rc.compute();
    std::vector<float> first(rc.raw_host_pixel_buf(), rc.raw_host_pixel_buf() + 6 * npix);
    rc.use_proj_store_accum_method();
    rc.compute();
    const OracleDrr r5 = OracleRayCast(s, {s.cam}, poses, cam0, 1.0f, XO_KERNEL_SUM, XO_STORE_ACCUM, 0.0f, nullptr, &first);
    CheckDrr(rc.raw_host_pixel_buf(), r5);
    rc.use_proj_store_replace_method();

    std::vector<float> bg(npix);
    for (size_type i = 0; i < npix; ++i)
    {
      bg[i] = 0.001f * static_cast<float>(i % 97);
    }
    rc.set_bg_proj(RayCaster::Proj(bg.data(), rows, cols));
    rc.compute();
    const float* bgp[1] = {bg.data()};
    const OracleDrr r6 = OracleRayCast(s, {s.cam}, poses, cam0, 1.0f, XO_KERNEL_SUM, XO_STORE_REPLACE, 0.0f, bgp);
    CheckDrr(rc.raw_host_pixel_buf(), r6);
    rc.set_use_bg_projs(false);
  }

  // caller-owned host buffer (use_external_host_pixel_buf, xregRayCastBaseCPU.cpp:120-126)
  {
    std::vector<float> ext(6 * npix, -1.0f);
    rc.use_external_host_pixel_buf(ext.data());
    rc.compute();
    CHECK(rc.raw_host_pixel_buf() == ext.data());
    CheckDrr(ext.data(), ref);
    rc.use_external_host_pixel_buf(nullptr);
  }

  // linear and nearest-neighbour interpolation; sinc / B-spline are unsupported (the OpenCL backend is linear only,
  // xregRayCastBaseOCL.cpp:338-341)
  rc.use_sinc_interp();
  CHECK(Throws<UnsupportedOperationException>([&] { rc.compute(); }));
  rc.use_bspline_interp();
  CHECK(Throws<UnsupportedOperationException>([&] { rc.compute(); }));
  rc.use_nn_interp();
  rc.compute();
  {
    // nearest neighbour: no arithmetic on the voxel values -> bit-identical to the CPU class (xo_drr_interp)
    std::vector<float> o(6 * npix, 0.0f);
    std::vector<float> p12(12 * poses.size());
    std::vector<uint32_t> ci(poses.size(), 0u);
    for (size_type i = 0; i < poses.size(); ++i)
      poses[i].to3x4(&p12[12 * i]);
    float i2p[12];
    s.vol.idx_to_phys(i2p);
    const xo_cam oc = ToOracleCam(s.cam);
    CHECK(xo_drr_interp(s.vol.data, s.vol.size, i2p, &oc, 1, p12.data(), ci.data(), static_cast<uint32_t>(poses.size()), 1.0f,
                        XO_KERNEL_SUM, XO_INTERP_NN, o.data(), nullptr, nullptr, nullptr, 0) == 0);
    CHECK(std::memcmp(o.data(), rc.raw_host_pixel_buf(), sizeof(float) * o.size()) == 0);
  }
  rc.use_linear_interp();
  rc.compute();
  CheckDrr(rc.raw_host_pixel_buf(), ref);

  // pose-list size must match num_projs (xregASSERT in the reference)
  CHECK(Throws<XregCudaError>([&] { rc.set_xforms_cam_to_itk_phys(FrameTransformList(5)); }));
}

// ---------------------------------------------------------------------------------------------------------------
static void TestMetrics(Context& ctx, const Scene& s)
{
  const size_type rows = s.cam.num_det_rows, cols = s.cam.num_det_cols, npix = rows * cols, n = 5;
  RayCasterLineIntCUDA rc(ctx);
  rc.set_volume(s.vol);
  rc.set_camera_model(s.cam);
  rc.set_num_projs(n);
  rc.allocate_resources();

  // fixed image: DRR at a held-out pose plus deterministic noise
  const FrameTransformList held = MakePoses(s, 1, 99, 1.0, 1.0);
  rc.set_num_projs(1);
  rc.set_xforms_cam_to_itk_phys(held);
  rc.compute();
  std::vector<float> fixed(rc.raw_host_pixel_buf(), rc.raw_host_pixel_buf() + npix);
  {
    float mx = 0;
    for (const float v : fixed)
    {
      mx = std::fmax(mx, v);
    }
    Rng rng(5);
    for (float& v : fixed)
    {
      v += static_cast<float>(rng.normal() * 0.01 * mx);
    }
  }
  std::vector<uint8_t> mask(npix);
  for (size_type r = 0; r < rows; ++r)
    for (size_type c = 0; c < cols; ++c)
    {
      const double dr = double(r) - 0.5 * (rows - 1), dc = double(c) - 0.5 * (cols - 1);
      mask[r * cols + c] = (dr * dr + dc * dc <= 0.2 * rows * rows) ? 1 : 0;
    }

  rc.set_num_projs(n);
  const FrameTransformList poses = MakePoses(s, n, 3);
  rc.set_xforms_cam_to_itk_phys(poses);
  rc.compute();
  // the oracle's metrics run on the oracle's own DRRs: the whole chain is compared, not just the metric stage
  const OracleDrr drr = OracleRayCast(s, {s.cam}, poses, std::vector<uint32_t>(n, 0));

  const ImgSimMetric2D::Image fixed_img(fixed.data(), rows, cols);
  const ImgSimMetric2D::ImageMask mask_img(mask.data(), rows, cols);

  auto prepare = [&](ImgSimMetric2D& sm, const bool with_mask) {
    sm.set_num_moving_images(n);
    sm.set_fixed_image(fixed_img);
    sm.set_mov_imgs_buf_from_ray_caster(&rc);
    if (with_mask)
    {
      sm.set_mask(mask_img);
    }
    sm.allocate_resources();
  };

  for (int with_mask = 0; with_mask < 2; ++with_mask)
  {
    const uint8_t* m = with_mask ? mask.data() : nullptr;
    std::vector<float> ref(n);
    {
      ImgSimMetric2DSSDCUDA sm(ctx);
      prepare(sm, with_mask);
      sm.compute();
      std::vector<float> mov = drr.img;
      xo_ssd(fixed.data(), m, rows, cols, mov.data(), n, ref.data(), 0);
      // SSD is not normalised: compare relative to the value
      double w = 0;
      for (size_type i = 0; i < n; ++i)
      {
        w = std::fmax(w, std::fabs(double(sm.sim_val(i)) - double(ref[i])) / std::fmax(1e-12, std::fabs(double(ref[i]))));
      }
      CHECK(w <= 1e-4);
    }
    {
      ImgSimMetric2DNCCCUDA sm(ctx);
      prepare(sm, with_mask);
      sm.compute();
      std::vector<float> mov = drr.img;
      xo_ncc(fixed.data(), m, rows, cols, mov.data(), n, ref.data(), 0);
      const double d = MaxAbsDiff(sm.sim_vals(), ref);
      std::printf("  ncc (mask %d): max |diff| %.3g\n", with_mask, d);
      CHECK(d <= kSIM_TOL);
    }
    {
      ImgSimMetric2DGradNCCCUDA sm(ctx);
      CHECK(sm.smooth_img_before_sobel_kernel_radius() == 5);
      sm.set_smooth_img_before_sobel_kernel_radius(3);
      prepare(sm, with_mask);
      sm.compute();
      xo_grad_ncc(fixed.data(), m, rows, cols, 3, drr.img.data(), n, ref.data(), 0);
      const double d = MaxAbsDiff(sm.sim_vals(), ref);
      std::printf("  grad-ncc (mask %d): max |diff| %.3g\n", with_mask, d);
      CHECK(d <= kSIM_TOL);
      // reached through the parameter mix-in, as the reference apps do with dynamic_cast (pelvis...main.cpp:259-270)
      ImgSimMetric2D* base = &sm;
      CHECK(dynamic_cast<ImgSimMetric2DGradImgParamInterface*>(base) != nullptr);
    }
    {
      ImgSimMetric2DPatchNCCCUDA sm(ctx);
      sm.set_patch_radius(4);
      sm.set_patch_stride(2);
      prepare(sm, with_mask);
      sm.compute();
      xo_patch_opts o = {4, 2, 0, 1, 1, 0, 1};
      std::vector<float> w(xo_num_patches(rows, cols, 4, 2));
      const float* wp = nullptr;
      if (m)
      {
        xo_patch_weights(rows, cols, &o, m, nullptr, w.data());
        wp = w.data();
      }
      xo_patch_ncc(fixed.data(), m, rows, cols, &o, wp, drr.img.data(), n, ref.data(), nullptr, 0);
      const double d = MaxAbsDiff(sm.sim_vals(), ref);
      std::printf("  patch-ncc (mask %d): max |diff| %.3g\n", with_mask, d);
      CHECK(d <= kSIM_TOL);
    }
    {
      ImgSimMetric2DPatchGradNCCCUDA sm(ctx);
      sm.set_patch_radius(5);
      sm.set_smooth_img_before_sobel_kernel_radius(5);
      prepare(sm, with_mask);
      sm.compute();
      xo_patch_opts o = {5, 1, 0, 1, 1, 0, 1};
      std::vector<float> w(xo_num_patches(rows, cols, 5, 1));
      const float* wp = nullptr;
      if (m)
      {
        xo_patch_weights(rows, cols, &o, m, nullptr, w.data());
        wp = w.data();
      }
      xo_patch_grad_ncc(fixed.data(), m, rows, cols, 5, &o, wp, drr.img.data(), n, ref.data(), 0);
      const double d = MaxAbsDiff(sm.sim_vals(), ref);
      std::printf("  patch-grad-ncc (mask %d): max |diff| %.3g\n", with_mask, d);
      CHECK(d <= kSIM_TOL);
      ImgSimMetric2D* base = &sm;
      CHECK(dynamic_cast<ImgSimMetric2DPatchCommon*>(base) != nullptr);
      CHECK(dynamic_cast<ImgSimMetric2DGradImgParamInterface*>(base) != nullptr);

      // the mask may change between computes (process_updated_mask): remove it again
      if (with_mask)
      {
        sm.set_mask(ImgSimMetric2D::ImageMask());
        sm.compute();
        xo_patch_grad_ncc(fixed.data(), nullptr, rows, cols, 5, &o, nullptr, drr.img.data(), n, ref.data(), 0);
        CHECK(MaxAbsDiff(sm.sim_vals(), ref) <= kSIM_TOL);
      }
    }
  }

  // moving images from a caller-owned host buffer (set_mov_imgs_host_buf) with a projection offset
  {
    ImgSimMetric2DNCCCUDA sm(ctx);
    sm.set_num_moving_images(2);
    sm.set_fixed_image(fixed_img);
    std::vector<float> host = drr.img;
    sm.set_mov_imgs_host_buf(host.data(), 3);
    sm.allocate_resources();
    sm.compute();
    std::vector<float> ref(n);
    std::vector<float> mov = drr.img;
    xo_ncc(fixed.data(), nullptr, rows, cols, mov.data(), n, ref.data(), 0);
    CHECK(std::fabs(sm.sim_val(0) - ref[3]) <= kSIM_TOL && std::fabs(sm.sim_val(1) - ref[4]) <= kSIM_TOL);
  }

  // error behaviour: compute before allocate; a mask of the wrong size
  {
    ImgSimMetric2DNCCCUDA sm(ctx);
    sm.set_num_moving_images(n);
    sm.set_fixed_image(fixed_img);
    CHECK(Throws<XregCudaError>([&] { sm.compute(); }));
    CHECK(Throws<XregCudaError>([&] { sm.set_mask(ImgSimMetric2D::ImageMask(mask.data(), rows - 1, cols)); }));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// two views: distribute_xforms_among_cam_models (camera-major), one metric per view at offset v * pop, CombineMean --
// the wiring of Intensity2D3DRegi::setup / obj_fn (xregIntensity2D3DRegi.cpp:43-133,571-696) -- and the same through
// the one-call objective.
static void TestMultiViewObjective(Context& ctx, const Scene& s)
{
  const size_type rows = s.cam.num_det_rows, cols = s.cam.num_det_cols, npix = rows * cols, pop = 4, views = 2;
  // second view: the C-arm rotated 35 degrees about the camera-frame y axis through the isocentre
  CameraModel cam2;
  cam2.coord_frame_type = s.cam.coord_frame_type;
  {
    const double a = 35.0 * 3.141592653589793 / 180.0;
    FrameTransform C, Ci, R;
    C(2, 3) = -250.0f;
    Ci(2, 3) = 250.0f;
    R(0, 0) = static_cast<float>(std::cos(a));
    R(0, 2) = static_cast<float>(-std::sin(a));
    R(2, 0) = static_cast<float>(std::sin(a));
    R(2, 2) = static_cast<float>(std::cos(a));
    cam2.setup(s.cam.intrins, C * R * Ci, rows, cols, s.cam.det_row_spacing, s.cam.det_col_spacing);
  }
  const std::vector<CameraModel> cams = {s.cam, cam2};

  RayCasterLineIntCUDA rc(ctx);
  rc.set_volume(s.vol);
  rc.set_camera_models(cams);
  rc.set_num_projs(pop * views);
  rc.allocate_resources();

  // fixed images: the two views of a held-out pose
  const FrameTransformList held = MakePoses(s, 1, 1234, 1.0, 1.0);
  rc.set_num_projs(views);
  rc.distribute_xform_among_cam_models(held[0]);
  CHECK(rc.camera_model_proj_associations() == (RayCaster::CamModelAssocList{0, 1}));
  rc.compute();
  std::vector<float> fixed(rc.raw_host_pixel_buf(), rc.raw_host_pixel_buf() + views * npix);
  rc.set_num_projs(pop * views);

  const FrameTransformList poses = MakePoses(s, pop, 77);
  rc.distribute_xforms_among_cam_models(poses);
  for (size_type g = 0; g < pop * views; ++g)
  {
    CHECK(rc.camera_model_proj_associations()[g] == g / pop);  // camera-major
  }
  rc.compute();

  // oracle: same distribution, DRRs, per-view Grad-NCC, mean over views
  std::vector<float> p12(12 * pop), dist(12 * pop * views);
  std::vector<uint32_t> cam_idx(pop * views);
  for (size_type i = 0; i < pop; ++i)
  {
    poses[i].to3x4(&p12[12 * i]);
  }
  xo_distribute_xforms(p12.data(), pop, views, dist.data(), cam_idx.data());
  FrameTransformList dist_x;
  for (size_type g = 0; g < pop * views; ++g)
  {
    dist_x.push_back(FrameTransform::From3x4(&dist[12 * g]));
  }
  const OracleDrr drr = OracleRayCast(s, cams, dist_x, cam_idx);
  CheckDrr(rc.raw_host_pixel_buf(), drr);
  std::vector<float> ref_views(views * pop), ref(pop);
  for (size_type v = 0; v < views; ++v)
  {
    xo_grad_ncc(fixed.data() + v * npix, nullptr, rows, cols, 5, drr.img.data() + v * pop * npix, pop, ref_views.data() + v * pop, 0);
  }
  xo_combine_mean(ref_views.data(), views, pop, ref.data());

  ImgSimMetric2DGradNCCCUDA sm0(ctx), sm1(ctx);
  ImgSimMetric2D* sms[2] = {&sm0, &sm1};
  for (size_type v = 0; v < views; ++v)
  {
    sms[v]->set_num_moving_images(pop);
    sms[v]->set_fixed_image(ImgSimMetric2D::Image(fixed.data() + v * npix, rows, cols));
    sms[v]->set_mov_imgs_buf_from_ray_caster(&rc, pop * v);
    sms[v]->allocate_resources();
    sms[v]->compute();
  }
  ImgSimMetric2DCombineMean combine;
  combine.set_num_sim_metrics(views);
  combine.set_num_projs_per_sim_metric(pop);
  combine.allocate_resources();
  combine.set_sim_metric(0, &sm0);
  combine.set_sim_metric(1, &sm1);
  combine.compute();
  const double d = MaxAbsDiff(combine.sim_vals(), ref);
  std::printf("  2 views x %zu poses, grad-ncc, CombineMean: max |diff| %.3g\n", pop, d);
  CHECK(d <= kSIM_TOL);

  // a metric cannot be re-bound to a different ray caster (xregImgSimMetric2DCPU.cpp:45-70)
  {
    RayCasterLineIntCUDA other(ctx);
    CHECK(Throws<XregCudaError>([&] { sm0.set_mov_imgs_buf_from_ray_caster(&other, 0); }));
  }

  // the same evaluation as ONE call (xrc_obj_fn), then a smaller population, then the full one again
  {
    RayCasterLineIntCUDA rc2(ctx);
    rc2.set_volume(s.vol);
    rc2.set_camera_models(cams);
    ImgSimMetric2DGradNCCCUDA a(ctx), b(ctx);
    a.set_fixed_image(ImgSimMetric2D::Image(fixed.data(), rows, cols));
    b.set_fixed_image(ImgSimMetric2D::Image(fixed.data() + npix, rows, cols));
    Intensity2D3DObjFn obj(&rc2, {&a, &b}, pop);
    const std::vector<float> full = obj(poses);
    CHECK(full == combine.sim_vals());  // bitwise: same kernels, same order
    CHECK(a.sim_vals() == sm0.sim_vals() && b.sim_vals() == sm1.sim_vals());
    const std::vector<float> two = obj({poses[2], poses[0]});
    CHECK(two.size() == 2 && two[0] == full[2] && two[1] == full[0]);  // a pose's value does not depend on its batch
    CHECK(obj(poses) == full);
    CHECK(Throws<XregCudaError>([&] { obj(FrameTransformList(pop + 1)); }));
  }
}

// ---------------------------------------------------------------------------------------------------------------
// the depth ray caster behind the same base class, and the projection pre-processing
static void TestDepthAndPreProc(Context& ctx, const Scene& s)
{
  const size_type rows = s.cam.num_det_rows, cols = s.cam.num_det_cols, npix = rows * cols, n = 3;
  const FrameTransformList poses = MakePoses(s, n, 11, 10.0, 6.0);
  float vmax = 0.0f;
  const size_t nvox = static_cast<size_t>(s.vol.size[0]) * s.vol.size[1] * s.vol.size[2];
  for (size_t i = 0; i < nvox; ++i)
  {
    vmax = std::max(vmax, s.vol.data[i]);
  }
  std::vector<float> p12(12 * n);
  std::vector<uint32_t> ci(n, 0u);
  for (size_type i = 0; i < n; ++i)
  {
    poses[i].to3x4(&p12[12 * i]);
  }
  float i2p[12];
  s.vol.idx_to_phys(i2p);
  const xo_cam oc = ToOracleCam(s.cam);

  RayCasterDepthCUDA depth(ctx);
  RayCaster& rc = depth;   // through the base class, as the factories hand it out
  rc.set_volume(s.vol);
  rc.set_camera_model(s.cam);
  rc.set_num_projs(n);
  rc.allocate_resources();
  rc.set_xforms_cam_to_itk_phys(poses);
  CHECK(rc.default_bg_pixel_val() == kRAY_CAST_MAX_DEPTH);
  size_t surfaces = 0;
  for (int pass = 0; pass < 2; ++pass)
  {
    const float thr = (pass ? 0.7f : 0.4f) * vmax;
    const size_type nb = pass ? 6 : 0;
    depth.set_render_thresh(thr);
    depth.set_num_backtracking_steps(nb);
    rc.compute();
    std::vector<float> want(n * npix, XO_RAY_CAST_MAX_DEPTH);
    CHECK(xo_depth(s.vol.data, s.vol.size, i2p, &oc, 1, p12.data(), ci.data(), static_cast<uint32_t>(n), 1.0f, XO_INTERP_LINEAR,
                   thr, static_cast<uint32_t>(nb), want.data(), 0) == 0);
    CHECK(std::memcmp(want.data(), rc.raw_host_pixel_buf(), sizeof(float) * want.size()) == 0);
    for (const float d : want)
    {
      surfaces += (d < 1.0e36f) ? 1 : 0;
    }
  }
  CHECK(surfaces > 1000);
  std::printf("  depth ray caster: bit-identical, %zu surface pixels\n", surfaces);

  // log remap and down-sampling of a projection (a line-integral DRR turned into an intensity image)
  RayCasterLineIntCUDA lin(ctx);
  lin.set_volume(s.vol);
  lin.set_camera_model(s.cam);
  lin.set_num_projs(1);
  lin.allocate_resources();
  lin.set_xforms_cam_to_itk_phys(FrameTransformList(1, poses[0]));
  lin.compute();
  std::vector<float> img(lin.raw_host_pixel_buf(), lin.raw_host_pixel_buf() + npix);
  for (float& v : img)
  {
    v = 4000.0f * std::exp(-v);
  }
  img[5] = 0.0f;
  {
    std::vector<float> got(npix), want(npix);
    const float i0 = LogRemap(ctx, Image2D<const float>(img.data(), rows, cols), got.data());
    float want_i0 = 0.0f;
    xo_log_remap(img.data(), static_cast<uint32_t>(rows), static_cast<uint32_t>(cols), 0, 1, 1.0f, nullptr, want.data(), &want_i0);
    CHECK(i0 == want_i0);
    double worst = 0.0;
    for (size_type i = 0; i < npix; ++i)
    {
      worst = std::max(worst, static_cast<double>(std::fabs(got[i] - want[i])) / std::max(1.0e-30, static_cast<double>(std::fabs(want[i]))));
    }
    CHECK(worst <= 2.5e-7);   // 1 ulp: correctly rounded logarithm vs glibc's logf
    std::printf("  log remap: I0 %.6g bit-equal, max rel diff %.2e\n", static_cast<double>(i0), worst);
  }
  for (const double f : {0.5, 0.25})
  {
    std::vector<float> got;
    size_type r = 0, c = 0;
    DownsampleImage(ctx, Image2D<const float>(img.data(), rows, cols), f, &got, &r, &c);
    uint32_t wr = 0, wc = 0;
    xo_downsample_size(static_cast<uint32_t>(rows), static_cast<uint32_t>(cols), f, &wr, &wc);
    std::vector<float> want(static_cast<size_t>(wr) * wc);
    xo_downsample_image(img.data(), static_cast<uint32_t>(rows), static_cast<uint32_t>(cols), f, -1.0, want.data());
    CHECK(r == wr && c == wc && got.size() == want.size());
    CHECK(std::memcmp(got.data(), want.data(), sizeof(float) * want.size()) == 0);
  }
  std::printf("  down-sampling: bit-identical to the oracle\n");
}

int main(int argc, char** argv)
{
  const bool no_gpu = (argc > 1 && std::strcmp(argv[1], "--no-gpu") == 0);
  try
  {
    TestHostLogic();
    if (no_gpu)
    {
      // no CPU fallback: without a device the context cannot be created and nothing computes
      int n_dev_ok = 0;
      try
      {
        Context c(0);
        n_dev_ok = 1;
      }
      catch (const XregCudaError& e)
      {
        CHECK(e.status() == XRC_ERR_CUDA);
      }
      (void)n_dev_ok;
      std::printf("host_mirror_test (no gpu): %s\n", g_failed ? "FAILED" : "ok");
      return g_failed ? 1 : 0;
    }
    Context ctx(0);
    const Scene s = MakeScene(64, 56, 48, 80, 96);
    TestRayCaster(ctx, s);
    TestMetrics(ctx, s);
    TestMultiViewObjective(ctx, s);
    TestDepthAndPreProc(ctx, s);
  }
  catch (const std::exception& e)
  {
    std::fprintf(stderr, "unexpected exception: %s\n", e.what());
    return 2;
  }
  std::printf("host_mirror_test: %s\n", g_failed ? "FAILED" : "ok");
  return g_failed ? 1 : 0;
}
