// pdl.cuh -- programmatic dependent launch (griddepcontrol) for the kernel chain of one forward pass.
// Every kernel of this back-end is launched with cudaLaunchAttributeProgrammaticStreamSerialization (b200_compute.cu: launch_k), so kernel
// N+1 may be scheduled while kernel N drains: its launch latency, barrier / TMEM setup and tensor-map prefetch overlap N's tail.
// The contract that makes this safe: EVERY kernel executes pdl_wait() before its first access to global memory (griddepcontrol.wait
// returns once all prerequisite grids have completed and their writes are visible), and nothing is written before it.
#pragma once

namespace b200 {

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// kernels without a prologue worth overlapping: let the next grid start launching, then wait for the previous one
__device__ __forceinline__ void pdl_prologue() { pdl_launch_dependents(); pdl_wait(); }

}  // namespace b200
