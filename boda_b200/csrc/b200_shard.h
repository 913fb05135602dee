// b200_shard.h -- batch sharding of has_conv_fwd_t::run_fwd over the GPUs of one box: the per-process communicator object behind the
// `b200_shard_*` C ABI (include/boda_b200.h). The reference has no multi-GPU path; see b200_shard.cu.
#pragma once
#include "boda_base.h"

namespace boda {

struct b200_shard_impl_t;

// What a kernel needs to do the gather itself (fcchain.cuh: the last layer's CTAs store the logits into every rank's buffer and the last CTA
// publishes the step): the peer-mapped buffers, the local one, and where the counters live. `step_ctr` (device) holds the last step published
// from this rank; a kernel that gathers reads it, publishes step_ctr + 1 and stores that back, so a CUDA graph can replay it.
struct b200_gather_desc_t {
  unsigned char *peer_base[16];
  unsigned char *local_base;
  uint64_t bytes_per_rank, flag_bytes;  // a buffer: [flag_bytes: uint32 step counters, one per rank; step_ctr at uint32 index 32][2 (step parity)][world][bytes_per_rank]
  int rank, world;
};

struct b200_shard_t {
  int device, rank, world;
  uint64_t n_launches = 0;  // kernels of this module launched so far (claimed in bench.py's gpu_launches)
  b200_shard_t(int device, int rank, int world);
  ~b200_shard_t();
  b200_shard_t(b200_shard_t const &) = delete;

  // NCCL (weights): rank 0 makes the 128-byte unique id, the caller hands it to every rank, every rank joins; then one in-place broadcast
  void nccl_unique_id(void *id_out_128);
  void nccl_init(void const *id_128);
  void broadcast(void *dev_buf, uint64_t bytes, int root, void *stream);
  void all_gather_nccl(void const *dev_src, void *dev_dst, uint64_t bytes_per_rank, void *stream);  // A/B baseline for the peer-memory gather

  // peer-memory gather (logits): export the local buffer's 64-byte IPC handle, import everyone's, then per step push + wait
  void gather_export(uint64_t bytes_per_rank, void *ipc_handle_out_64);
  void gather_import(void const *ipc_handles_world_x_64);
  uint32_t gather_push(void const *dev_src, void *stream);  // returns the step number it published
  void gather_wait(uint32_t step, void *stream);
  uint32_t gather_push_wait(void const *dev_src, uint32_t wait_step, void *stream);  // both in ONE launch (wait_step 0: push only); returns the step it published
  void *gather_ptr(uint32_t step) const;  // device pointer of the local [world][bytes_per_rank] result of `step` (two halves, by step parity)
  uint32_t step() const;
  // in-kernel gather: the descriptor (also writes the host's step counter to the device-side one), and the way back (device -> host counter)
  void gather_desc(b200_gather_desc_t &d);
  uint32_t step_from_device();

  b200_shard_impl_t *impl;
};

}  // namespace boda
