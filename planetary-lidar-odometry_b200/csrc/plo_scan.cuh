// plo_scan.cuh — block-level scan / order-preserving compaction helpers shared by the
// index build and the pair compaction (internal linkage: included by several .cu files).
#pragma once

#include <cstdlib>

#include "plo_internal.cuh"

namespace {

// Programmatic dependent launch (index build, source upload): these kernels are short (2-25 us) and strictly chained, so
// much of what separates them is launch latency.  A kernel launched through plo_launch_chained() may have its blocks
// scheduled while the previous grid is still draining; PLO_CHAIN_ENTER() at its very top waits until that grid has
// completed and its writes are visible (griddepcontrol.wait -- the ordering a plain stream launch gives, nothing is read
// or written before it), then lets the NEXT kernel of the chain be scheduled behind this one.  Launched the ordinary way
// both instructions do nothing.
#ifdef PLO_CHAIN_EARLY   // experiment: release the successor before waiting (a deeper cascade of resident, waiting grids)
#define PLO_CHAIN_ENTER()                                         \
  do {                                                            \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
    asm volatile("griddepcontrol.wait;" ::: "memory");            \
  } while (0)
#else
#define PLO_CHAIN_ENTER()                                         \
  do {                                                            \
    asm volatile("griddepcontrol.wait;" ::: "memory");            \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); \
  } while (0)
#endif

template <typename... KArgs, typename... Args>
static inline cudaError_t plo_launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool plain = getenv("PLO_NO_CHAIN") != nullptr;   // debugging / A-B: ordinary stream launches
  cfg.numAttrs = plain ? 0 : 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

constexpr int kTile = 1024;   // elements per block in the compaction kernels (256 threads x 4)

// in-place exclusive scan of a small int array by one block; total -> *total_out
__global__ void __launch_bounds__(1024) k_scan_exclusive(int* __restrict__ data, int n, int* __restrict__ total_out) {
  PLO_CHAIN_ENTER();
  __shared__ int s_warp[32];
  __shared__ int s_carry;
  constexpr int kPer = 8;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 1024 * kPer) {
    const int i0 = base + threadIdx.x * kPer;
    int v[kPer];
    int sum = 0;
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      v[j] = (i0 + j < n) ? data[i0 + j] : 0;
      sum += v[j];
    }
    int inc = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(PLO_FULL_MASK, inc, o);
      if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      int w = s_warp[lane];
      int winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(PLO_FULL_MASK, winc, o);
        if (lane >= o) winc += t;
      }
      s_warp[lane] = winc - w;  // exclusive over warps
    }
    __syncthreads();
    const int carry = s_carry;
    int run = carry + s_warp[warp] + (inc - sum);
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
      if (i0 + j < n) data[i0 + j] = run;
      run += v[j];
    }
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = run;
    __syncthreads();
  }
  if (threadIdx.x == 0 && total_out != nullptr) *total_out = s_carry;
}

// order-preserving rank of each finite point inside its 1024-tile (i = base + j*256 + t)
__device__ __forceinline__ void tile_ranks(const bool fin[kTile / 256], int rank[kTile / 256]) {
  __shared__ int s_c[(kTile / 256) * 8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned bal[kTile / 256];
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j) {
    bal[j] = __ballot_sync(PLO_FULL_MASK, fin[j]);
    if (lane == 0) s_c[j * 8 + warp] = __popc(bal[j]);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
#pragma unroll
    for (int q = 0; q < (kTile / 256) * 8; ++q) { int t = s_c[q]; s_c[q] = run; run += t; }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j) rank[j] = s_c[j * 8 + warp] + __popc(bal[j] & ((1u << lane) - 1u));
}

}  // namespace
