// Cross-GPU exchange for the tile-sharded objective (one process per GPU): a barrier and an all-gather
// of the similarity scalars done by ONE tiny kernel per rank over peer-mapped memory (CUDA IPC mappings of the ranks'
// projection buffers, whose tail holds an exchange block) -- NVLink stores and system-scope flags instead of two
// NCCL collectives per objective evaluation (~20 us each at 8 GPUs, of a 1.6 ms step).
//   block layout (kXchgBlockBytes + 4 * max_projs bytes at xchg_off of every rank's buffer):
//     u32 flag[kMaxPeers] at 32-byte pitch: flag[r] = the last epoch rank r has reached (written by rank r)
//     f32 sims[max_projs] at kXchgSimsOff: the gathered values, unit u = view * n_poses + pose
// Epochs only grow, so one flag set serves every barrier of a step: a rank that is ahead has passed all earlier ones.
#include "common.h"

namespace xrc
{

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v)
{
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p)
{
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__global__ void __launch_bounds__(kXchgThreads) xchg_kernel(const XchgArgs a)
{
  const uint32_t tid = threadIdx.x;
  // 1. my values into every rank's gathered vector (their global unit index)
  for (uint32_t s = 0; s < a.n_seg; ++s)
    for (uint32_t i = tid; i < a.seg_count[s]; i += kXchgThreads)
    {
      const float v = __ldcg(a.seg_src[s] + i);
      for (uint32_t r = 0; r < a.n_ranks; ++r)
        reinterpret_cast<float*>(a.blk[r] + kXchgSimsOff)[a.seg_first[s] + i] = v;
    }
  __threadfence_system();
  __syncthreads();
  // 2. tell every rank that I am here (everything this stream did before is visible to whoever sees the flag), then
  // wait for every rank's flag in my own block
  __shared__ int timed_out;
  if (tid == 0)
    timed_out = 0;
  __syncthreads();
  if (tid < a.n_ranks)
  {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(a.blk[tid] + 32u * a.rank), a.epoch);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(a.blk[a.rank] + 32u * tid);
    const unsigned long long t0 = global_ns();
    while ((int32_t)(ld_acquire_sys(mine) - a.epoch) < 0)
    {
      if (global_ns() - t0 > a.timeout_ns)
      {
        timed_out = 1;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  if (a.host_status && tid == 0 && timed_out)
    *a.host_status = 1u;
  // 3. all units -> host-mapped memory
  if (a.host_out)
  {
    const volatile float* src = reinterpret_cast<const volatile float*>(a.blk[a.rank] + kXchgSimsOff);
    for (uint32_t i = tid; i < a.n_units_total; i += kXchgThreads)
      a.host_out[i] = src[i];
  }
}

int launch_xchg(const XchgArgs& a, cudaStream_t st)
{
  xchg_kernel<<<1, kXchgThreads, 0, st>>>(a);
  count_launch();
  XRC_CUDA(cudaGetLastError());
  return XRC_OK;
}

}  // namespace xrc
