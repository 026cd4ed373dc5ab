/* Plain-C client of the drop-in boundary (include/xreg_cuda.h): what the C++ adapters of INTEGRATION.md do, without
 * Python.  Builds with `gcc -std=c99` (the header must be valid C), links against libxreg_cuda.so.
 *
 *   abi_smoke           -> on a GPU box: a 24^3 constant volume, one axis-aligned camera, checks the line integral of the
 *                          central ray against the analytic value (constant volume, ray along an axis),
 *                          then NCC of the DRR with itself (= (1 - (N-1)/N) / 2) through xrc_obj_fn.
 *   abi_smoke --no-gpu  -> only the calls that need no device: version, exp map, error path of a bad argument.
 * Exit code 0 = all checks passed. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "xreg_cuda.h"

#define CHECK(cond)                                                        \
  do                                                                       \
  {                                                                        \
    if (!(cond))                                                           \
    {                                                                      \
      fprintf(stderr, "FAILED %s:%d: %s (last error: %s)\n", __FILE__, __LINE__, #cond, xrc_last_error()); \
      return 1;                                                            \
    }                                                                      \
  } while (0)

#define OK(expr) CHECK((expr) == XRC_OK)

int main(int argc, char** argv)
{
  const int no_gpu = (argc > 1 && strcmp(argv[1], "--no-gpu") == 0);

  CHECK(xrc_version() == XRC_VERSION);
  {
    /* ExpSE3 of a pure translation and of a rotation by pi/2 about z */
    const float t[6] = {0, 0, 0, 1.5f, -2.0f, 3.0f};
    float m[12];
    xrc_exp_se3(t, m);
    CHECK(m[0] == 1.0f && m[5] == 1.0f && m[10] == 1.0f && m[3] == 1.5f && m[7] == -2.0f && m[11] == 3.0f);
    const float r[6] = {0, 0, 1.57079632679f, 0, 0, 0};
    xrc_exp_se3(r, m);
    CHECK(fabsf(m[0]) < 1e-6f && fabsf(m[1] + 1.0f) < 1e-6f && fabsf(m[4] - 1.0f) < 1e-6f && fabsf(m[10] - 1.0f) < 1e-6f);
  }
  /* error convention: status code + thread-local message, no exceptions across the boundary */
  CHECK(xrc_rc_create(NULL, NULL) == XRC_ERR_INVALID);
  CHECK(strlen(xrc_last_error()) > 0);
  if (no_gpu)
  {
    printf("abi_smoke (no gpu): ok\n");
    return 0;
  }

  xrc_ctx* ctx = NULL;
  OK(xrc_ctx_create(0, &ctx));
  xrc_rc* rc = NULL;
  OK(xrc_rc_create(ctx, &rc));

  /* 24^3 volume of constant attenuation 0.02 / mm, 1 mm voxels, centred on the origin */
  enum { N = 24 };
  float* vol = (float*)malloc(sizeof(float) * N * N * N);
  for (int i = 0; i < N * N * N; ++i)
    vol[i] = 0.02f;
  const float* vols[1] = {vol};
  const uint64_t dims[1][3] = {{N, N, N}};
  const float half = 0.5f * (N - 1);
  const float i2p[1][12] = {{1, 0, 0, -half, 0, 1, 0, -half, 0, 0, 1, -half}};
  OK(xrc_rc_set_volumes(rc, 1, vols, dims, i2p));

  /* 33 x 33 detector, 1 mm pixels, focal length 500 mm, origin at the focal point, detector at z = +500 */
  xrc_cam cam;
  memset(&cam, 0, sizeof(cam));
  cam.rows = cam.cols = 33;
  cam.focal_len = 500.0f;
  cam.frame_type = 0;
  /* K = [f 0 16; 0 f 16; 0 0 1] in pixel units (pixel = 1 mm)  ->  K^-1 */
  cam.intrins_inv[0] = 1.0f / 500.0f;
  cam.intrins_inv[2] = -16.0f / 500.0f;
  cam.intrins_inv[4] = 1.0f / 500.0f;
  cam.intrins_inv[5] = -16.0f / 500.0f;
  cam.intrins_inv[8] = 1.0f;
  cam.extrins_inv[0] = cam.extrins_inv[5] = cam.extrins_inv[10] = 1.0f;
  OK(xrc_rc_set_cameras(rc, 1, &cam));
  OK(xrc_rc_allocate(rc, 2));

  /* camera -> volume: the volume centre sits 250 mm down the optical axis */
  const float poses[2][12] = {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, -250.0f}, {1, 0, 0, 2.0f, 0, 1, 0, 0, 0, 0, 1, -250.0f}};
  OK(xrc_rc_set_poses(rc, 2, &poses[0][0], NULL));
  OK(xrc_rc_compute(rc, 0));
  float* drr = (float*)malloc(sizeof(float) * 2 * 33 * 33);
  OK(xrc_rc_read_projs(rc, 0, 2, drr));
  /* central ray: along z, 23 voxels of material; the 1e-3 nudge is in units of the 500 mm source-detector segment, so
   * 0.5 mm goes at either end: 22 steps, 23 samples of 0.02 at step 1 (constant volume, axis-aligned ray) */
  const float centre = drr[16 * 33 + 16];
  CHECK(fabsf(centre - 23.0f * 0.02f) < 1e-5f);
  /* a denser 8 x 8 column through the middle (re-setting the volume re-packs it): the central ray now sums 0.05's,
   * and the projection is no longer constant, which the similarity test below needs */
  for (int z = 0; z < N; ++z)
    for (int y = 8; y < 16; ++y)
      for (int x = 8; x < 16; ++x)
        vol[(z * N + y) * N + x] = 0.05f;
  OK(xrc_rc_set_volumes(rc, 1, vols, dims, i2p));
  OK(xrc_rc_compute(rc, 0));
  OK(xrc_rc_read_projs(rc, 0, 2, drr));
  CHECK(fabsf(drr[16 * 33 + 16] - 23.0f * 0.05f) < 1e-5f);
  CHECK(fabsf(drr[0] - 23.0f * 0.02f) < 1e-5f);
  uint64_t total = 0, fetched = 0;
  OK(xrc_rc_ray_info(rc, 0, NULL, NULL, &total));
  OK(xrc_rc_fetched_samples(rc, 0, &fetched));
  CHECK(total > 0 && fetched <= total);

  /* NCC of the first DRR against itself and against the shifted one, through the one-call objective */
  xrc_sm* sm = NULL;
  OK(xrc_sm_create(ctx, XRC_SM_NCC, &sm));
  OK(xrc_sm_set_fixed(sm, drr, 33, 33));
  OK(xrc_sm_bind_ray_caster(sm, rc, 0));
  OK(xrc_sm_allocate(sm, 2));
  float sims[2] = {-1.0f, -1.0f};
  xrc_sm* sms[1];
  sms[0] = sm;
  OK(xrc_obj_fn(rc, 0, sms, 1, 2, &poses[0][0], sims, NULL));
  const float n_pix = 33.0f * 33.0f;
  CHECK(fabsf(sims[0] - 0.5f * (1.0f - (n_pix - 1.0f) / n_pix)) < 2e-6f);
  CHECK(sims[1] > sims[0]);
  /* call-order errors are reported, not crashed on */
  CHECK(xrc_rc_set_num_projs(rc, 3) == XRC_ERR_INVALID);

  OK(xrc_sm_destroy(sm));
  OK(xrc_rc_destroy(rc));
  OK(xrc_ctx_destroy(ctx));
  free(vol);
  free(drr);
  printf("abi_smoke: ok (central line integral %.6f, ncc sims %.7f %.7f, %llu of %llu samples fetched, %llu launches)\n", centre,
         sims[0], sims[1], (unsigned long long)fetched, (unsigned long long)total, (unsigned long long)xrc_launch_count());
  return 0;
}
