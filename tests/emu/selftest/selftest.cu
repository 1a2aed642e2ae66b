// Self-test of the CPU interpreter (tests/emu/cuda_runtime.h): the warp / block / grid primitives must behave as the CUDA
// programming guide documents them, otherwise the replayed parity suite proves nothing.  Built by g++ only.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <vector>
namespace cg = cooperative_groups;

struct Out { unsigned v[16]; };

__global__ void k_warp(Out* out) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Out& o = out[threadIdx.x];
  o.v[0] = __shfl_sync(0xffffffffu, lane * 10u, 5);                 // broadcast of lane 5
  o.v[1] = __shfl_sync(0xffffffffu, lane, 3, 8);                    // lane 3 of the own 8-lane segment
  o.v[2] = __shfl_xor_sync(0xffffffffu, lane, 1);
  o.v[3] = __shfl_xor_sync(0xffffffffu, lane, 4, 8);
  o.v[4] = __shfl_up_sync(0xffffffffu, lane, 2);                    // lanes 0,1 keep their own value
  o.v[5] = __shfl_down_sync(0xffffffffu, lane, 3);                  // lanes 29..31 keep their own value
  o.v[6] = __ballot_sync(0xffffffffu, (lane % 3) == 0);
  o.v[7] = (unsigned)__any_sync(0xffffffffu, lane == 17) * 2u + (unsigned)__all_sync(0xffffffffu, lane < 32);
  o.v[8] = __reduce_min_sync(0xffffffffu, 100u - lane) + 1000u * __reduce_add_sync(0xffffffffu, 1u);
  const unsigned long long big = ((unsigned long long)lane << 40) | warp;
  o.v[9] = (unsigned)(__shfl_sync(0xffffffffu, big, 31) >> 40);     // 64-bit payload
  const double d = 0.5 * lane;
  o.v[10] = (unsigned)(__shfl_xor_sync(0xffffffffu, d, 16) * 2.0);
  __shared__ unsigned acc[4];
  if (threadIdx.x < 4) acc[threadIdx.x] = 0;
  __syncthreads();
  atomicAdd(&acc[warp & 3], lane);
  atomicMin(&acc[3], 7u);                                           // acc[3] may be 0 already: stays <= 7
  __syncthreads();
  o.v[11] = acc[warp & 3] == 0 ? 0u : acc[0];                       // every warp added 0+..+31 = 496 to its slot
  o.v[12] = blockIdx.x * 1000u + blockDim.x + gridDim.x * 100000u;
}

// lanes that return early must not deadlock a later full-mask collective of the remaining lanes' warps, and a block
// barrier must count only live threads
__global__ void k_exit(unsigned* out) {
  if (threadIdx.x >= 40) return;                                    // warp 1 keeps 8 lanes, warps 2.. exit completely
  __syncthreads();
  const unsigned b = __ballot_sync(0xffffffffu, 1);
  out[threadIdx.x] = b;
}

__global__ void k_grid(unsigned* counter, unsigned* out) {
  cg::grid_group g = cg::this_grid();
  for (int it = 0; it < 5; ++it) {
    if (threadIdx.x == 0) atomicAdd(counter, 1u);
    g.sync();
    const unsigned seen = *(volatile unsigned*)counter;             // every CTA must see all increments of this phase
    if (threadIdx.x == 0 && seen != (unsigned)(it + 1) * gridDim.x) atomicAdd(&out[0], 1u);
    g.sync();
  }
  if (threadIdx.x == 0) atomicAdd(&out[1], 1u);
}

#define CHECK(c) do { if (!(c)) { std::printf("interpreter self-test FAILED: %s (line %d)\n", #c, __LINE__); return 1; } } while (0)

int main() {
  Out* d_out; cudaMalloc((void**)&d_out, sizeof(Out) * 128);
  emu::launch(k_warp, 2u, 128u, (size_t)0, 1u, false, d_out);
  std::vector<Out> h(128); cudaMemcpy(h.data(), d_out, sizeof(Out) * 128, cudaMemcpyDeviceToHost);
  for (unsigned t = 0; t < 128; ++t) {
    const unsigned lane = t & 31; const Out& o = h[t];
    CHECK(o.v[0] == 50u);
    CHECK(o.v[1] == (lane & ~7u) + 3u);
    CHECK(o.v[2] == (lane ^ 1u));
    CHECK(o.v[3] == (lane ^ 4u));
    CHECK(o.v[4] == (lane >= 2 ? lane - 2 : lane));
    CHECK(o.v[5] == (lane + 3 < 32 ? lane + 3 : lane));
    unsigned bal = 0; for (unsigned l = 0; l < 32; ++l) if (l % 3 == 0) bal |= 1u << l;
    CHECK(o.v[6] == bal);
    CHECK(o.v[7] == 3u);
    CHECK(o.v[8] == 69u + 32000u);
    CHECK(o.v[9] == 31u);
    CHECK(o.v[10] == (lane ^ 16u));
    CHECK(o.v[11] == 496u || o.v[11] == 0u);
    CHECK(o.v[12] == 1000u + 128u + 200000u);      // last CTA run sequentially leaves blockIdx 1 in its own records only
  }
  unsigned* d_u; cudaMalloc((void**)&d_u, sizeof(unsigned) * 64); cudaMemset(d_u, 0, sizeof(unsigned) * 64);
  emu::launch(k_exit, 1u, 128u, (size_t)0, 1u, false, d_u);
  std::vector<unsigned> hu(64); cudaMemcpy(hu.data(), d_u, sizeof(unsigned) * 64, cudaMemcpyDeviceToHost);
  for (unsigned t = 0; t < 40; ++t) { if (hu[t] != (t < 32 ? 0xffffffffu : 0x000000ffu)) std::printf("t=%u got %08x\n", t, hu[t]); CHECK(hu[t] == (t < 32 ? 0xffffffffu : 0x000000ffu)); }
  // cooperative launch: one process per CTA, grid barrier + global atomics through the shared arena
  unsigned* d_c; cudaMalloc((void**)&d_c, sizeof(unsigned) * 4); cudaMemset(d_c, 0, sizeof(unsigned) * 4);
  CHECK(emu::launch(k_grid, 6u, 64u, (size_t)0, 1u, true, d_c, d_c + 1) == cudaSuccess);
  unsigned hc[4]; cudaMemcpy(hc, d_c, sizeof(hc), cudaMemcpyDeviceToHost);
  CHECK(hc[0] == 30u && hc[1] == 0u && hc[2] == 6u);
  std::printf("interpreter self-test ok\n");
  return 0;
}
