// TEST INFRASTRUCTURE.  The xReg adapter classes (adapters/xreg: RayCasterLineIntCUDA,
// ImgSimMetric2D*CUDA) LINKED with the reference's own base-class sources (xregRayCastInterface.cpp,
// xregRayCastSyncBuf.cpp, xregImgSimMetric2D.cpp, xregImgSimMetric2DPatchCommon.cpp, xregImgSimMetric2DCombine.cpp,
// CameraModel::setup cut from xregPerspectiveXform.cpp -- compiled where they lie under /root/reference, see
// build_link.py) and driven through the reference's interfaces only -- xreg::RayCaster*, xreg::ImgSimMetric2D*, the
// parameter mix-ins reached by dynamic_cast, ImgSimMetric2DCombineMean -- in the order Intensity2D3DRegi::setup() and
// ::obj_fn() use them (lib/regi/interfaces_2d_3d/xregIntensity2D3DRegi.cpp:43-133, 571-696), with the static-volume
// background set up like MultiLevelMultiObjRegi::run (xregMultiObjMultiLevel2D3DRegi.cpp:363-412).
//
//   xreg_adapter_driver <input.bin> <output.bin>
//
// Input / output layouts: tests/test_gpu_adapters_run.py (the only reader / writer).  Eigen / ITK / OpenCV are the
// functional stand-ins of tests/xreg_link/third_party; everything between them and libxreg_cuda.so is the reference's
// code and the adapters.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "xregImgSimMetric2DCUDA.h"
#include "xregImgSimMetric2DCombine.h"
#include "xregRayCastLineIntCUDA.h"

namespace
{

struct Reader
{
  FILE* f;
  template <class T>
  T get()
  {
    T v;
    if (fread(&v, sizeof(T), 1, f) != 1)
      throw std::runtime_error("short input");
    return v;
  }
  template <class T>
  void get(T* dst, size_t n)
  {
    if (n && fread(dst, sizeof(T), n, f) != n)
      throw std::runtime_error("short input");
  }
};

template <class T>
void put(FILE* f, const T* src, size_t n)
{
  if (n && fwrite(src, sizeof(T), n, f) != n)
    throw std::runtime_error("short write");
}

using Vol = xreg::RayCaster::Vol;
using VolPtr = xreg::RayCaster::VolPtr;
using Img = xreg::ImgSimMetric2D::Image;
using ImgPtr = xreg::ImgSimMetric2D::ImagePtr;
using Mask = xreg::ImgSimMetric2D::ImageMask;
using MaskPtr = xreg::ImgSimMetric2D::ImageMaskPtr;

VolPtr make_volume(const uint32_t n[3], const double sp[3], const double org[3], const double dir[9], std::vector<float>& data)
{
  VolPtr v = Vol::New();
  Vol::RegionType reg;
  for (unsigned d = 0; d < 3; ++d)
  {
    reg.SetIndex(d, 0);
    reg.SetSize(d, n[d]);
  }
  v->SetRegions(reg);
  auto px = Vol::PixelContainer::New();
  px->SetImportPointer(data.data(), data.size(), false);
  v->SetPixelContainer(px);
  v->SetSpacing(sp);
  v->SetOrigin(org);
  Vol::DirectionType D;
  for (unsigned r = 0; r < 3; ++r)
    for (unsigned c = 0; c < 3; ++c)
      D(r, c) = dir[3 * r + c];
  v->SetDirection(D);
  return v;
}

template <class TImg, class T>
typename TImg::Pointer make_image(uint32_t rows, uint32_t cols, std::vector<T>& data)
{
  typename TImg::Pointer im = TImg::New();
  typename TImg::RegionType reg;
  reg.SetIndex(0, 0);
  reg.SetIndex(1, 0);
  reg.SetSize(0, cols);
  reg.SetSize(1, rows);
  im->SetRegions(reg);
  auto px = TImg::PixelContainer::New();
  px->SetImportPointer(data.data(), data.size(), false);
  im->SetPixelContainer(px);
  return im;
}

std::shared_ptr<xreg::ImgSimMetric2D> make_metric(xrc_ctx* ctx, uint32_t kind)
{
  // the "cuda" branches of the backend factories (xregImgSimMetric2DProgOpts.cpp:71-94)
  switch (kind)
  {
    case 0: return std::make_shared<xreg::ImgSimMetric2DNCCCUDA>(ctx);
    case 1: return std::make_shared<xreg::ImgSimMetric2DGradNCCCUDA>(ctx);
    case 2: return std::make_shared<xreg::ImgSimMetric2DPatchNCCCUDA>(ctx);
    case 3: return std::make_shared<xreg::ImgSimMetric2DPatchGradNCCCUDA>(ctx);
    case 4: return std::make_shared<xreg::ImgSimMetric2DSSDCUDA>(ctx);
    default: throw std::runtime_error("bad metric kind");
  }
}

}  // namespace

int main(int argc, char** argv)
{
  if (argc != 3)
  {
    fprintf(stderr, "usage: %s input.bin output.bin\n", argv[0]);
    return 2;
  }
  try
  {
    FILE* fin = fopen(argv[1], "rb");
    if (!fin)
      throw std::runtime_error("cannot open input");
    Reader in{fin};
    char magic[4];
    in.get(magic, 4);
    if (memcmp(magic, "XRLK", 4) != 0)
      throw std::runtime_error("bad magic");
    const uint32_t n_views = in.get<uint32_t>(), pop = in.get<uint32_t>(), n_evals = in.get<uint32_t>();
    uint32_t dims[3];
    in.get(dims, 3);
    const uint32_t rows = in.get<uint32_t>(), cols = in.get<uint32_t>();
    const uint32_t kind = in.get<uint32_t>(), patch_radius = in.get<uint32_t>(), patch_stride = in.get<uint32_t>(),
                   gauss = in.get<uint32_t>(), use_mask = in.get<uint32_t>(), n_moving = in.get<uint32_t>(),
                   has_static = in.get<uint32_t>(), weight_flags = in.get<uint32_t>(), n_subset = in.get<uint32_t>();
    const float step = in.get<float>();
    double sp[3], org[3], dir[9];
    in.get(sp, 3);
    in.get(org, 3);
    in.get(dir, 9);
    const size_t nvox = (size_t)dims[0] * dims[1] * dims[2], npix = (size_t)rows * cols;
    const uint32_t n_vols = n_moving + has_static;
    std::vector<std::vector<float>> vol_data(n_vols, std::vector<float>(nvox));
    for (auto& v : vol_data)
      in.get(v.data(), nvox);

    std::vector<xreg::CameraModel> cams(n_views);
    for (uint32_t v = 0; v < n_views; ++v)
    {
      float K[9], E[16], rs, cs;
      in.get(K, 9);
      in.get(E, 16);
      rs = in.get<float>();
      cs = in.get<float>();
      const uint32_t frame = in.get<uint32_t>();
      xreg::Mat3x3 intrins;
      xreg::Mat4x4 extrins;
      for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c)
          intrins(r, c) = K[3 * r + c];
      for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c)
          extrins(r, c) = E[4 * r + c];
      cams[v].coord_frame_type = static_cast<xreg::CameraModel::CameraCoordFrame>(frame);
      cams[v].setup(intrins, extrins, rows, cols, rs, cs);   // the reference's own CameraModel::setup
    }
    std::vector<std::vector<float>> fixed(n_views, std::vector<float>(npix));
    for (auto& f : fixed)
      in.get(f.data(), npix);
    std::vector<std::vector<unsigned char>> masks(use_mask ? n_views : 0, std::vector<unsigned char>(npix));
    for (auto& m : masks)
      in.get(m.data(), npix);
    std::vector<uint64_t> subset(n_subset);
    in.get(subset.data(), n_subset);
    std::vector<float> static_pose(16);
    if (has_static)
      in.get(static_pose.data(), 16);
    // poses[eval][moving object][pose] as 4x4 row-major
    std::vector<float> poses((size_t)n_evals * n_moving * pop * 16);
    in.get(poses.data(), poses.size());
    fclose(fin);

    xrc_ctx* ctx = nullptr;
    if (xrc_ctx_create(0, &ctx) != XRC_OK)
      throw std::runtime_error(std::string("xrc_ctx_create: ") + xrc_last_error());
    FILE* fout = fopen(argv[2], "wb");
    if (!fout)
      throw std::runtime_error("cannot open output");
    {
      // ---- what the apps do before the registration object exists (e.g. pelvis...main.cpp:230-270)
      std::shared_ptr<xreg::RayCaster> ray_caster = std::make_shared<xreg::RayCasterLineIntCUDA>(ctx);
      if (auto* li = dynamic_cast<xreg::RayCastLineIntParamInterface*>(ray_caster.get()))
        li->set_kernel_id(xreg::kRAY_CAST_LINE_INT_SUM_KERNEL);
      xreg::RayCaster::VolList vols;
      for (uint32_t i = 0; i < n_vols; ++i)
        vols.push_back(make_volume(dims, sp, org, dir, vol_data[i]));
      ray_caster->set_volumes(vols);
      ray_caster->set_camera_models(cams);
      ray_caster->use_linear_interp();
      ray_caster->set_ray_step_size(step);

      std::vector<std::shared_ptr<xreg::ImgSimMetric2D>> sim_metrics;
      std::vector<ImgPtr> fixed_imgs;
      std::vector<MaskPtr> mask_imgs;
      for (uint32_t v = 0; v < n_views; ++v)
      {
        auto sm = make_metric(ctx, kind);
        if (auto* patch = dynamic_cast<xreg::ImgSimMetric2DPatchCommon*>(sm.get()))
        {
          patch->set_patch_radius(patch_radius);
          patch->set_patch_stride(patch_stride);
          patch->set_compute_mean_of_patch_sims((weight_flags & 1) != 0);
          patch->set_weight_patch_sims_in_combine((weight_flags & 2) != 0);
          patch->set_use_mask_for_patch_weighting((weight_flags & 4) != 0);
          patch->set_use_mask_for_patch_stats((weight_flags & 8) != 0);
          if (n_subset)
          {
            xreg::ImgSimMetric2DPatchCommon::PatchIndexList inds(subset.begin(), subset.end());
            patch->set_patches_to_use(inds);
          }
        }
        if (auto* grad = dynamic_cast<xreg::ImgSimMetric2DGradImgParamInterface*>(sm.get()))
          grad->set_smooth_img_before_sobel_kernel_radius(gauss);
        fixed_imgs.push_back(make_image<Img>(rows, cols, fixed[v]));
        sm->set_fixed_image(fixed_imgs.back());
        if (use_mask)
        {
          mask_imgs.push_back(make_image<Mask>(rows, cols, masks[v]));
          sm->set_mask(mask_imgs.back());
        }
        sim_metrics.push_back(sm);
      }

      // ---- static volume: background projections (xregMultiObjMultiLevel2D3DRegi.cpp:363-412)
      std::vector<std::vector<float>> bg_store;
      if (has_static)
      {
        xreg::FrameTransform T;
        for (int r = 0; r < 4; ++r)
          for (int c = 0; c < 4; ++c)
            T.matrix()(r, c) = static_pose[4 * r + c];
        // the ray caster is allocated once, for the population; the static pass shrinks num_projs
        ray_caster->set_num_projs((xreg::size_type)pop * n_views);
        ray_caster->allocate_resources();
        ray_caster->set_num_projs(n_views);
        ray_caster->set_use_bg_projs(false);
        ray_caster->use_proj_store_replace_method();
        ray_caster->distribute_xform_among_cam_models(T);
        ray_caster->compute(n_moving);   // the static volume is the last one
        xreg::RayCaster::ProjList bg_imgs(n_views);
        bg_store.assign(n_views, std::vector<float>(npix));
        for (uint32_t v = 0; v < n_views; ++v)
        {
          xreg::RayCaster::ProjPtr p = ray_caster->proj(v);   // deep copy, as ITKImageDeepCopy there
          memcpy(bg_store[v].data(), p->GetBufferPointer(), npix * sizeof(float));
          bg_imgs[v] = make_image<xreg::RayCaster::Proj>(rows, cols, bg_store[v]);
        }
        ray_caster->set_use_bg_projs(true);
        ray_caster->set_bg_projs(bg_imgs);
      }

      // ---- Intensity2D3DRegi::setup() (xregIntensity2D3DRegi.cpp:43-133)
      const xreg::size_type num_projs_per_view = pop;
      const xreg::size_type tot_num_projs = num_projs_per_view * n_views;
      ray_caster->set_num_projs(tot_num_projs);
      if (!has_static)   // need_to_alloc_ray_caster_
        ray_caster->allocate_resources();
      for (uint32_t v = 0; v < n_views; ++v)
      {
        sim_metrics[v]->set_save_aux_info(false);
        sim_metrics[v]->set_num_moving_images(num_projs_per_view);
        // view-major ordering of projections in memory
        sim_metrics[v]->set_mov_imgs_buf_from_ray_caster(ray_caster.get(), num_projs_per_view * v);
      }
      for (uint32_t v = 0; v < n_views; ++v)
        sim_metrics[v]->allocate_resources();
      auto combiner = std::make_shared<xreg::ImgSimMetric2DCombineMean>();
      combiner->set_num_sim_metrics(n_views);
      combiner->set_num_projs_per_sim_metric(num_projs_per_view);
      combiner->allocate_resources();
      for (uint32_t v = 0; v < n_views; ++v)
        combiner->set_sim_metric(v, sim_metrics[v].get());

      // ---- Intensity2D3DRegi::obj_fn() (xregIntensity2D3DRegi.cpp:571-696), n_evals times
      std::vector<float> sim_vals(pop), per_view((size_t)n_views * pop);
      for (uint32_t e = 0; e < n_evals; ++e)
      {
        const bool orig_use_bg = ray_caster->use_bg_projs();
        if (has_static)
          ray_caster->set_use_bg_projs(true);
        ray_caster->use_proj_store_replace_method();
        for (uint32_t obj = 0; obj < n_moving; ++obj)
        {
          xreg::FrameTransformList xforms(pop);
          for (uint32_t p = 0; p < pop; ++p)
          {
            const float* m = poses.data() + (((size_t)e * n_moving + obj) * pop + p) * 16;
            for (int r = 0; r < 4; ++r)
              for (int c = 0; c < 4; ++c)
                xforms[p].matrix()(r, c) = m[4 * r + c];
          }
          ray_caster->distribute_xforms_among_cam_models(xforms);
          ray_caster->compute(obj);
          if (has_static)
            ray_caster->set_use_bg_projs(false);
          ray_caster->use_proj_store_accum_method();
        }
        for (uint32_t v = 0; v < n_views; ++v)
          sim_metrics[v]->compute();
        combiner->compute();
        for (uint32_t p = 0; p < pop; ++p)
          sim_vals[p] = combiner->sim_val(p);
        if (has_static)
          ray_caster->set_use_bg_projs(orig_use_bg);
        for (uint32_t v = 0; v < n_views; ++v)
          for (uint32_t p = 0; p < pop; ++p)
            per_view[(size_t)v * pop + p] = sim_metrics[v]->sim_val(p);
        put(fout, sim_vals.data(), pop);
        put(fout, per_view.data(), per_view.size());
      }
      // projections of the last evaluation through the host accessors
      put(fout, ray_caster->raw_host_pixel_buf(), (size_t)tot_num_projs * npix);
      xreg::RayCaster::ProjPtr p0 = ray_caster->proj(tot_num_projs - 1);
      put(fout, p0->GetBufferPointer(), npix);
      const float spacing_out[2] = {(float)p0->GetSpacing()[0], (float)p0->GetSpacing()[1]};
      put(fout, spacing_out, 2);
      cv::Mat m0 = ray_caster->proj_ocv(0);
      put(fout, m0.ptr<float>(0), npix);
      // patch grid as the reference's PatchCommon set it up (host logic of the reference, run inside the adapter)
      uint64_t n_patches = 0;
      if (auto* patch = dynamic_cast<xreg::ImgSimMetric2DPatchCommon*>(sim_metrics[0].get()))
        n_patches = patch->num_patches();
      put(fout, &n_patches, 1);
    }
    fclose(fout);
    xrc_ctx_destroy(ctx);
  }
  catch (const std::exception& e)
  {
    fprintf(stderr, "xreg_adapter_driver: %s\n", e.what());
    return 1;
  }
  return 0;
}
