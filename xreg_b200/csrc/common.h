// Shared host-side declarations of libxreg_cuda.so (not part of the C ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>
#include <vector>

#include "../../include/xreg_cuda.h"

namespace xrc
{

void set_error(const std::string& msg);
extern std::atomic<uint64_t> g_launch_count;

inline void count_launch(uint64_t n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define XRC_FAIL(code, msg)  \
  do                         \
  {                          \
    ::xrc::set_error((msg)); \
    return (code);           \
  } while (0)

#define XRC_CHECK_ARG(cond, msg)        \
  do                                    \
  {                                     \
    if (!(cond))                        \
      XRC_FAIL(XRC_ERR_INVALID, (msg)); \
  } while (0)

#define XRC_CUDA(expr)                                                                   \
  do                                                                                     \
  {                                                                                      \
    cudaError_t e__ = (expr);                                                            \
    if (e__ != cudaSuccess)                                                              \
    {                                                                                    \
      ::xrc::set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));             \
      return (e__ == cudaErrorMemoryAllocation) ? XRC_ERR_NOMEM : XRC_ERR_CUDA;          \
    }                                                                                    \
  } while (0)

#define XRC_TRY(expr)     \
  do                      \
  {                       \
    int s__ = (expr);     \
    if (s__ != XRC_OK)    \
      return s__;         \
  } while (0)

// Volume as held on the device.  Exactly one of the layout payloads is populated
// (plus `linear` kept for repacking / debugging when cheap).
struct DeviceVolume
{
  uint64_t dims[3] = {0, 0, 0};
  float idx_to_phys[12];
  float phys_to_idx[12];
  int layout = XRC_LAYOUT_LINEAR;
  void* data = nullptr;             // cudaMalloc'd payload (LINEAR padded / QUAD / OCT)
  cudaArray_t array = nullptr;      // TEX / TEX_QUAD
  cudaTextureObject_t tex = 0;
  size_t bytes = 0;
  // XRC_LAYOUT_PAX: one padded XY-quad record stack per principal ray axis k
  // (slow axis c = k, fast a = (k+1)%3, mid b = (k+2)%3); see drr.cu
  // Stacks are built ON DEMAND (build_pax_stack): only the principal axes the current cameras x poses can select
  // exist -- the first from the f32 copy `src` (dropped right after), further ones from an existing stack, whose
  // records keep the voxel values unchanged (a single-view registration lives on one stack: 4x the volume instead of
  // 12x).  A CTA whose preferred stack is missing uses any built one (same samples).
  void* pax[3] = {nullptr, nullptr, nullptr};
  uint32_t pax_sb[3] = {0, 0, 0};   // record strides of the mid / slow axis
  uint32_t pax_sc[3] = {0, 0, 0};
  float* src = nullptr;             // x-fastest f32 volume on the device (PAX only; until the first stack is built)
  uint32_t* h_want = nullptr;       // 3 words of host-mapped pinned memory: a CTA that had to fall back from stack k
                                    // stores 1 to word k; the next compute() builds that stack (drr.cu, api.cu)
  // empty-space map (drr.cu, "empty-space trimming"): one bit per 8^3-voxel block, set when any voxel of the
  // block or of its 26 neighbours is non-zero; x-fastest, 32 blocks per word
  uint32_t* occ = nullptr;
  uint32_t occ_wx = 0, occ_ny = 0, occ_nz = 0;   // words per block row, block rows, block slices
  float occ_lo[3] = {0, 0, 0}, occ_hi[3] = {0, 0, 0};  // box of sample positions that can touch a non-zero voxel
  float occ_fill = 1.0f;            // fraction of the map's bits that are set inside that box (sparse volumes: interior gaps)
};

constexpr int kLayoutNN = 100;      // launch_drr: nearest-neighbour sampling of any payload (not an XRC_LAYOUT_* of the ABI)
constexpr uint32_t kInlinePoses = 8;
constexpr uint32_t kMaxPeers = 8;   // ranks of a tile-sharded job (one box)

// Arguments of the DRR kernels (passed by value).
struct DrrArgs
{
  const void* vol;            // layout payload
  cudaTextureObject_t tex;
  int nx, ny, nz;             // volume dims
  float phys_to_idx[12];
  const xrc_cam* cams;        // device
  uint32_t n_cams;            // camera indices read from device memory are clamped to n_cams - 1
  const float* poses;         // device, n_projs x 12
  const uint32_t* cam_idx;    // device
  uint32_t n_projs;
  uint32_t rows, cols;
  uint32_t tiles_x, tiles_y;
  float step_size;
  float* out;                 // n_projs x rows x cols
  int init_mode;              // 0: default_bg + val, 1: bg[cam] + val, 2: out + val (ACCUM)
  float default_bg;
  const float* bg;            // n_cams x rows x cols
  unsigned long long* sample_counter;  // optional
  int order;                  // 0: projection fastest over CTAs, 1: tile fastest
  const void* pax[3];         // XRC_LAYOUT_PAX stacks (nullptr: not built yet)
  uint32_t* pax_want;         // host-mapped: word k := 1 when a CTA wanted the missing stack k
  uint32_t pax_sb[3], pax_sc[3];
  int variant;                // tuning: bit0 = scalar (non-packed) FP32 math in the PAX kernel
  uint8_t* ray_mask;          // ray-info kernel only
  uint32_t* ray_steps;        // ray-info kernel only
  const uint32_t* occ;        // empty-space map of the volume (nullptr = march every sample)
  uint32_t occ_wx, occ_ny;
  float occ_lo[3], occ_hi[3];
  int count_only;             // instrumentation: count the samples the kernel would fetch, do not march / store
  int gaps;                   // sparse volume: also skip runs of empty samples INSIDE a ray's trimmed range (drr.cu, GAPS)
  // tile subset of this launch: tiles tile_first + i * tile_stride, i < tile_count (tile_stride == 0 on entry: all tiles)
  uint32_t tile_first, tile_stride, tile_count;
  unsigned long long* tile_counter;   // optional, with sample_counter: fetched samples per detector tile (tile planning)
  // multi-GPU tile sharding (xrc_rc_compute_tiles): projection `proj` is stored at its global index in the buffer of
  // the rank that owns it -- peer_out[owner(proj)], a peer-mapped (NVLink) address unless the owner is this rank --
  // where the n_projs projections are cut into peer_n contiguous balanced chunks (peer_base = n_projs / peer_n, the
  // first peer_extra chunks take one more).  peer_n == 0: single device, `out`.
  uint32_t peer_n, peer_base, peer_extra;
  float* peer_out[kMaxPeers];
  // nearest-neighbour interpolation (kLayoutNN): voxel (ix, iy, iz) is the float at nn_base[nn_off + ix * nn_s[0] +
  // iy * nn_s[1] + iz * nn_s[2]] -- the padded f32 copy, or the first component of a record of whatever stack exists
  const float* nn_base;
  uint32_t nn_s[3], nn_off;
  // depth ray caster (launch_depth): collision threshold and number of step-halving refinements
  float depth_thresh;
  uint32_t depth_backtrack;
  // small populations (latency regime): poses travel in the kernel parameters instead of an H2D copy
  int use_inline;
  float inl_poses[kInlinePoses * 12];
  uint32_t inl_cam[kInlinePoses];
};

int repack_volume(const float* d_linear, DeviceVolume* v, int layout, cudaStream_t st);
int build_pax_stack(DeviceVolume* v, int k, cudaStream_t st);   // no-op when stack k exists
int build_occupancy(const float* d_linear, DeviceVolume* v, cudaStream_t st);
void launch_hu_to_lin_att(float* d_vol, size_t n, float hu_lower, cudaStream_t st);
void free_volume(DeviceVolume* v);
int launch_drr(const DrrArgs& a, int layout, int kernel_id, cudaStream_t st);
int launch_ray_info(const DrrArgs& a, cudaStream_t st);
int launch_depth(const DrrArgs& a, int nearest, cudaStream_t st);   // needs the nn_* payload description

void affine_inverse_f32(const float a[12], float out[12]);

// Cross-GPU exchange block at the tail of every rank's projection buffer (xchg.cu)
constexpr uint32_t kXchgThreads = 256;
constexpr uint32_t kXchgMaxSeg = 16;                    // views a rank's unit range can touch
constexpr uint32_t kXchgSimsOff = 1024;                 // flags (8 x 32 B) first, then the gathered values
constexpr uint32_t kXchgBlockBytes = kXchgSimsOff;      // + 4 * max_projs
struct XchgArgs
{
  uint32_t n_ranks, rank, epoch;
  unsigned char* blk[kMaxPeers];         // every rank's exchange block (own included; peers' over CUDA IPC)
  uint32_t n_seg;                        // this rank's values: seg_count[s] floats at seg_src[s] are units seg_first[s]...
  const float* seg_src[kXchgMaxSeg];
  uint32_t seg_first[kXchgMaxSeg], seg_count[kXchgMaxSeg];
  uint32_t n_units_total;
  float* host_out;                       // after the barrier: all n_units_total gathered values, host-mapped (or null)
  uint32_t* host_status;                 // set to 1 when a peer did not arrive within timeout_ns
  unsigned long long timeout_ns;
};
int launch_xchg(const XchgArgs& a, cudaStream_t st);

}  // namespace xrc
