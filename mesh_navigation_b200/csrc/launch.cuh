// Launch / dynamic-shared-memory macros shared by the kernel headers and the host code.
#pragma once
// Kernel launches go through one macro: the CPU interpreter behind the `-m "not gpu"` logic tests (tests/emu/) compiles
// this very file with g++, which has no <<<>>>.  MNB_EMU_ACTIVE is only ever defined by tests/emu/cuda_runtime.h; the
// shipped library is built by nvcc without it and contains no host execution path for any kernel.
#ifdef MNB_EMU_ACTIVE
#define MNB_LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch(kern, (unsigned)(grid), (unsigned)(block), (size_t)(smem), 1u, false, __VA_ARGS__)
#define MNB_DYNAMIC_SMEM(name) unsigned char* name = emu::g_cta.dyn_smem
#else
#define MNB_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define MNB_DYNAMIC_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif
