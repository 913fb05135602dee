// b200_shard.cu -- batch sharding of has_conv_fwd_t::run_fwd over the GPUs of one box, behind the C ABI (`b200_shard_*`, include/boda_b200.h).
// The reference has no multi-GPU path; north_star asks for "a single NCCL broadcast of weights and gather of logits
// over NVLink" with a C++ host side. One process per GPU; this object holds that process's communicator and buffers:
//   * weights:  ncclBroadcast of ONE flat device buffer (the NCCL C API, resolved from the already-loaded libnccl.so.2 at run time, so the
//               library keeps no link-time dependency); the caller slices it into the parameter vars device-to-device -- no host round trip;
//   * logits:   every rank owns a gather buffer [world][bytes_per_rank] that its peers map through CUDA IPC. After its forward a rank runs
//               ONE small kernel (`shard_push_kernel`, one CTA per peer) that writes its logits straight into slot [step parity][rank] of every
//               peer's buffer over NVLink (128-bit stores on peer-mapped pointers) and then raises its step counter there with a system-scope
//               release; `shard_wait_kernel` (one CTA) acquires the step counters of all ranks in the LOCAL buffer. No NCCL kernel spins on
//               SMs beside the persistent contraction kernels (which want all 148 SMs, one CTA each): the round-1 per-step ncclAllGather did,
//               and 8-GPU efficiency fell to 0.53. The NCCL all-gather is kept as `b200_shard_all_gather_nccl` for A/B runs.
// Rendezvous bytes (the 128-byte ncclUniqueId, the 64-byte IPC handles) are plain memory the caller moves between the processes with
// whatever it has (bench.py: torch.distributed object collectives; a Boda driver: files or MPI).
#include "b200_shard.h"
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <cstdio>
#include <cstring>

namespace boda {

#define SH_CHK(x) do { cudaError_t const e_ = (x); if (e_ != cudaSuccess) { rt_err(string("CUDA error: ") + cudaGetErrorString(e_) + " in " #x " at " + __FILE__ + ":" + std::to_string(__LINE__)); } } while (0)

namespace {

constexpr int kMaxWorld = 16;
constexpr uint64_t kFlagBytes = 256;  // [world] step counters in front of the data, in the same allocation (one IPC handle)

struct nccl_api_t {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

nccl_api_t &nccl() {
  static nccl_api_t api;
  if (api.lib) { return api; }
  // the copy the process already holds (torch's bundled libnccl.so.2, same soname) is returned by dlopen; else the system one is loaded
  void *lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { rt_err(string("b200_shard: libnccl.so.2 is not loadable: ") + dlerror()); }
  auto sym = [&](char const *n) { void *p = dlsym(lib, n); if (!p) { rt_err(string("b200_shard: libnccl.so.2 lacks ") + n); } return p; };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.Broadcast = reinterpret_cast<decltype(api.Broadcast)>(sym("ncclBroadcast"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  api.lib = lib;
  return api;
}
void nccl_chk(ncclResult_t r, char const *what) { if (r != ncclSuccess) { rt_err(string("NCCL error in ") + what + ": " + nccl().GetErrorString(r)); } }

struct peer_ptrs_t { unsigned char *base[kMaxWorld]; };

// One CTA per destination rank: copy `bytes` (a multiple of 16) of this rank's output into slot [rank] of that rank's gather buffer -- a
// peer-mapped pointer, the stores travel over NVLink -- then publish: all of the CTA's stores are ordered before a system-scope release
// store of the step counter at [rank] of the destination's flag block.
__global__ void __launch_bounds__(256) shard_push_kernel(peer_ptrs_t peers, unsigned char const *__restrict__ src, uint64_t bytes, int rank, int world, unsigned int step) {
  unsigned char *dst_base = peers.base[blockIdx.x];
  uint4 const *s = reinterpret_cast<uint4 const *>(src);
  // two halves, by step parity: the logits of step i stay readable while step i+1 is being written (the consumer is at most one step behind)
  uint4 *d = reinterpret_cast<uint4 *>(dst_base + kFlagBytes + (static_cast<uint64_t>(step & 1u) * world + static_cast<uint64_t>(rank)) * bytes);
  uint64_t const n = bytes >> 4;
  for (uint64_t i = threadIdx.x; i < n; i += 4 * 256) {  // four 128-bit loads in flight per thread
    uint4 v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (i + k * 256 < n) { v[k] = s[i + k * 256]; } }
#pragma unroll
    for (int k = 0; k < 4; ++k) { if (i + k * 256 < n) { d[i + k * 256] = v[k]; } }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int *flag = reinterpret_cast<unsigned int *>(dst_base) + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(step) : "memory");
  }
}

// One CTA, one thread per rank: wait until that rank's step counter in the LOCAL flag block has reached `step` (its logits of that step have
// landed in the local gather buffer). Bounded: a lost peer traps with a message instead of hanging the GPU.
__global__ void shard_wait_kernel(unsigned char *local_base, int world, unsigned int step) {
  if (static_cast<int>(threadIdx.x) < world) {
    unsigned int const *flag = reinterpret_cast<unsigned int const *>(local_base) + threadIdx.x;
    unsigned int v = 0, spins = 0;
    while (true) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      if (static_cast<int>(v - step) >= 0) { break; }
      __nanosleep(100);
      if (++spins > (1u << 25)) { printf("b200_shard: rank %d never published step %u (have %u)\n", (int)threadIdx.x, step, v); __trap(); }
    }
  }
}

// Both in one launch (the per-step form): CTAs 0 .. world-1 push this step's logits as shard_push_kernel does, CTA `world` waits for
// `wait_step` as shard_wait_kernel does (wait_step == 0: nothing to wait for). Launched with programmatic stream serialization: the launch
// and the wait overlap the tail of the forward; the pushing CTAs execute griddepcontrol.wait before they read the logits.
__global__ void __launch_bounds__(1024) shard_push_wait_kernel(peer_ptrs_t peers, unsigned char const *__restrict__ src, uint64_t bytes, int rank, int world, unsigned int step,
                                                              unsigned char *local_base, unsigned int wait_step) {
  if (static_cast<int>(blockIdx.x) == world) {
    if (wait_step != 0u && static_cast<int>(threadIdx.x) < world) {
      unsigned int const *flag = reinterpret_cast<unsigned int const *>(local_base) + threadIdx.x;
      unsigned int v = 0, spins = 0;
      while (true) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
        if (static_cast<int>(v - wait_step) >= 0) { break; }
        __nanosleep(100);
        if (++spins > (1u << 25)) { printf("b200_shard: rank %d never published step %u (have %u)\n", (int)threadIdx.x, wait_step, v); __trap(); }
      }
    }
    return;
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");  // the forward that wrote `src` is complete and visible
  unsigned char *dst_base = peers.base[blockIdx.x];
  uint4 const *s = reinterpret_cast<uint4 const *>(src);
  uint4 *d = reinterpret_cast<uint4 *>(dst_base + kFlagBytes + (static_cast<uint64_t>(step & 1u) * world + static_cast<uint64_t>(rank)) * bytes);
  uint64_t const n = bytes >> 4;
  for (uint64_t i = threadIdx.x; i < n; i += 8 * 1024) {  // 1024 threads x eight 128-bit loads in flight: 128 KB of logits in one round trip
    uint4 v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (i + k * 1024 < n) { v[k] = s[i + k * 1024]; } }
#pragma unroll
    for (int k = 0; k < 8; ++k) { if (i + k * 1024 < n) { d[i + k * 1024] = v[k]; } }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned int *flag = reinterpret_cast<unsigned int *>(dst_base) + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(step) : "memory");
  }
}

}  // namespace

struct b200_shard_impl_t {
  ncclComm_t comm = nullptr;
  unsigned char *local = nullptr;       // [kFlagBytes of step counters][2 (step parity)][world][bytes_per_rank]
  uint64_t bytes_per_rank = 0;
  peer_ptrs_t peers;
  bool imported = false;
  unsigned int step = 0;
};

b200_shard_t::b200_shard_t(int device_, int rank_, int world_) : device(device_), rank(rank_), world(world_), impl(new b200_shard_impl_t) {
  if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world) { delete impl; impl = nullptr; rt_err("b200_shard: bad rank / world " + str(rank) + " / " + str(world)); }
  memset(&impl->peers, 0, sizeof(impl->peers));
  int ndev = 0;
  cudaError_t const e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { delete impl; impl = nullptr; rt_err(string("b200_shard needs a CUDA device and there is no CPU fallback: ") + cudaGetErrorString(e)); }
}
b200_shard_t::~b200_shard_t() {
  if (!impl) { return; }
  cudaSetDevice(device);
  if (impl->imported) { for (int r = 0; r < world; ++r) { if (r != rank && impl->peers.base[r]) { cudaIpcCloseMemHandle(impl->peers.base[r]); } } }
  if (impl->local) { cudaFree(impl->local); }
  if (impl->comm) { nccl().CommDestroy(impl->comm); }
  delete impl;
}

void b200_shard_t::nccl_unique_id(void *id_out) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  nccl_chk(nccl().GetUniqueId(&id), "ncclGetUniqueId");
  memcpy(id_out, &id, sizeof(id));
}
void b200_shard_t::nccl_init(void const *id_bytes) {
  if (impl->comm) { rt_err("b200_shard: NCCL communicator already initialised"); }
  SH_CHK(cudaSetDevice(device));
  ncclUniqueId id;
  memcpy(&id, id_bytes, sizeof(id));
  nccl_chk(nccl().CommInitRank(&impl->comm, world, id, rank), "ncclCommInitRank");
}
void b200_shard_t::broadcast(void *dev_buf, uint64_t bytes, int root, void *stream) {
  if (!impl->comm) { rt_err("b200_shard: broadcast before nccl_init"); }
  SH_CHK(cudaSetDevice(device));
  nccl_chk(nccl().Broadcast(dev_buf, dev_buf, bytes, ncclChar, root, impl->comm, static_cast<cudaStream_t>(stream)), "ncclBroadcast");
}
void b200_shard_t::all_gather_nccl(void const *dev_src, void *dev_dst, uint64_t bytes_per_rank, void *stream) {
  if (!impl->comm) { rt_err("b200_shard: all_gather before nccl_init"); }
  SH_CHK(cudaSetDevice(device));
  nccl_chk(nccl().AllGather(dev_src, dev_dst, bytes_per_rank, ncclChar, impl->comm, static_cast<cudaStream_t>(stream)), "ncclAllGather");
}

void b200_shard_t::gather_export(uint64_t bytes_per_rank, void *ipc_handle_out) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  if (impl->local) { rt_err("b200_shard: gather buffer already allocated"); }
  if (bytes_per_rank == 0 || (bytes_per_rank & 15)) { rt_err("b200_shard: bytes_per_rank must be a positive multiple of 16"); }
  if (static_cast<uint64_t>(world) * 4 > 32 * 4) { rt_err("b200_shard: world too large for the flag block"); }  // (uint32 index 32 is the device-side step counter)
  SH_CHK(cudaSetDevice(device));
  uint64_t const total = kFlagBytes + 2 * static_cast<uint64_t>(world) * bytes_per_rank;
  SH_CHK(cudaMalloc(&impl->local, total));
  SH_CHK(cudaMemset(impl->local, 0, total));
  impl->bytes_per_rank = bytes_per_rank;
  cudaIpcMemHandle_t h;
  SH_CHK(cudaIpcGetMemHandle(&h, impl->local));
  memcpy(ipc_handle_out, &h, sizeof(h));
}
void b200_shard_t::gather_import(void const *ipc_handles) {
  if (!impl->local) { rt_err("b200_shard: gather_import before gather_export"); }
  if (impl->imported) { rt_err("b200_shard: peers already imported"); }
  SH_CHK(cudaSetDevice(device));
  for (int r = 0; r < world; ++r) {
    if (r == rank) { impl->peers.base[r] = impl->local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<unsigned char const *>(ipc_handles) + static_cast<size_t>(r) * sizeof(h), sizeof(h));
    void *p = nullptr;
    SH_CHK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    impl->peers.base[r] = static_cast<unsigned char *>(p);
  }
  impl->imported = true;
}
uint32_t b200_shard_t::gather_push(void const *dev_src, void *stream) {
  if (!impl->imported) { rt_err("b200_shard: gather_push before gather_import"); }
  SH_CHK(cudaSetDevice(device));
  ++impl->step;
  shard_push_kernel<<<world, 256, 0, static_cast<cudaStream_t>(stream)>>>(impl->peers, static_cast<unsigned char const *>(dev_src), impl->bytes_per_rank, rank, world, impl->step);
  SH_CHK(cudaGetLastError());
  ++n_launches;
  return impl->step;
}
void b200_shard_t::gather_wait(uint32_t step, void *stream) {
  if (!impl->imported) { rt_err("b200_shard: gather_wait before gather_import"); }
  SH_CHK(cudaSetDevice(device));
  shard_wait_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(impl->local, world, step);
  SH_CHK(cudaGetLastError());
  ++n_launches;
}
uint32_t b200_shard_t::gather_push_wait(void const *dev_src, uint32_t wait_step, void *stream) {
  if (!impl->imported) { rt_err("b200_shard: gather_push_wait before gather_import"); }
  SH_CHK(cudaSetDevice(device));
  ++impl->step;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(world + 1);
  cfg.blockDim = dim3(1024);
  cfg.stream = static_cast<cudaStream_t>(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SH_CHK(cudaLaunchKernelEx(&cfg, shard_push_wait_kernel, impl->peers, static_cast<unsigned char const *>(dev_src), impl->bytes_per_rank, rank, world, impl->step, impl->local, wait_step));
  ++n_launches;
  return impl->step;
}
void *b200_shard_t::gather_ptr(uint32_t step) const { return impl->local ? impl->local + kFlagBytes + static_cast<uint64_t>(step & 1u) * world * impl->bytes_per_rank : nullptr; }
uint32_t b200_shard_t::step() const { return impl->step; }
void b200_shard_t::gather_desc(b200_gather_desc_t &d) {
  if (!impl->imported) { rt_err("b200_shard: gather_desc before gather_import"); }
  SH_CHK(cudaSetDevice(device));
  memset(&d, 0, sizeof(d));
  for (int r = 0; r < world; ++r) { d.peer_base[r] = impl->peers.base[r]; }
  d.local_base = impl->local;
  d.bytes_per_rank = impl->bytes_per_rank;
  d.flag_bytes = kFlagBytes;
  d.rank = rank; d.world = world;
  SH_CHK(cudaDeviceSynchronize());
  SH_CHK(cudaMemcpy(impl->local + 32 * 4, &impl->step, 4, cudaMemcpyHostToDevice));  // device-side step counter := the host's
}
uint32_t b200_shard_t::step_from_device() {
  SH_CHK(cudaSetDevice(device));
  SH_CHK(cudaDeviceSynchronize());
  SH_CHK(cudaMemcpy(&impl->step, impl->local + 32 * 4, 4, cudaMemcpyDeviceToHost));
  return impl->step;
}

}  // namespace boda
