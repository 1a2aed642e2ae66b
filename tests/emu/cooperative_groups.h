// tests/emu/cooperative_groups.h -- the two group barriers the kernels use, on the CPU interpreter (see cuda_runtime.h).
#pragma once
#include "cuda_runtime.h"
namespace cooperative_groups {
struct grid_group { void sync() const { emu::grid_barrier(); } };
struct cluster_group {
  void sync() const { emu::cluster_barrier(); }
  unsigned block_rank() const { return blockIdx.x % emu::g_cta.cluster_size; }
  unsigned num_blocks() const { return emu::g_cta.cluster_size; }
};
struct thread_block { void sync() const { emu::cta_barrier(); } };
static inline grid_group this_grid() { return grid_group{}; }
static inline cluster_group this_cluster() { return cluster_group{}; }
static inline thread_block this_thread_block() { return thread_block{}; }
}  // namespace cooperative_groups
