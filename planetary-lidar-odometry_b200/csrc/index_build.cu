// index_build.cu — map index build (replaces IMLSICPMatcher::setTargetPointCloud's
// Nabo::NNSearchD::createKDTreeLinearHeap, src/imls_icp.cpp:80-103) and source upload
// (setSourcePointCloud, src/imls_icp.cpp:74-78), including the non-finite strip of
// RemoveNANandINFData (src/imls_icp.cpp:58-72).
//
// Index = "curve-sorted wide BVH": 39-bit Hilbert key per point (over the
// cloud's bounding cube, 13 bit/axis) -> LSD radix sort (4 x 10 bit, stable) -> leaves of 32
// consecutive points -> levels of 32 consecutive nodes, one AABB per node.  Everything
// is sized by the number of uploaded points, so no host synchronisation is needed:
// non-finite points get the maximal key and +inf coordinates and sink to the tail.
//
// Algorithmic bytes per map point: read 24 B (xyz+normal) + write 24 B
// reordered + 4 B index = 52 B.  Roofline: HBM.
#include <math_constants.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <string>

#include "plo_internal.cuh"
#include "plo_scan.cuh"

namespace {

#ifndef PLO_SORT_ITEMS
#define PLO_SORT_ITEMS 16   // measured at 1 M keys (profiles/r2x_ab_index_build.txt): 16 -> 0.229 ms, 8 -> 0.231, 4 -> 0.282
#endif
constexpr int kSortItems = PLO_SORT_ITEMS;      // keys per thread in the radix-sort kernels
constexpr int kSortTile = 256 * kSortItems;     // keys per block
constexpr int kRadixBits = 10;    // 4 passes over the 39-bit key
constexpr int kRadix = 1 << kRadixBits;
#ifndef PLO_AXIS_BITS
#define PLO_AXIS_BITS 13
#define PLO_KEY_BITS 40
#endif
constexpr int kAxisBits = PLO_AXIS_BITS;     // Hilbert cells per axis = 2^13 (3.7 cm on a 300 m cube; ties keep input order)
constexpr int kKeyBits = PLO_KEY_BITS;
constexpr int kPasses = kKeyBits / kRadixBits;

__device__ __forceinline__ unsigned f2ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  unsigned u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
  return __uint_as_float(u);
}
__device__ __forceinline__ bool finite3(float x, float y, float z) {
  return isfinite(x) && isfinite(y) && isfinite(z);
}

// bbox[0..2] = min (ordered uint), bbox[3..5] = max
__global__ void k_init_bbox(unsigned* bbox) {
  PLO_CHAIN_ENTER();
  if (threadIdx.x < 3) bbox[threadIdx.x] = 0xffffffffu;
  else if (threadIdx.x < 6) bbox[threadIdx.x] = 0u;
}

// records -> float4 point (w = 1 if xyz finite) + float4 normal; finite count per block;
// bounding box of the finite points.
__global__ void __launch_bounds__(256) k_unpack_count(const char* __restrict__ rec, int stride, int n, int vec16,
                                                      float4* __restrict__ praw, float4* __restrict__ nraw,
                                                      int* __restrict__ blockcnt, unsigned* __restrict__ bbox) {
  PLO_CHAIN_ENTER();
  __shared__ int s_cnt[8];
  const int base = blockIdx.x * kTile;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int cnt = 0;
  float lo[3] = {CUDART_INF_F, CUDART_INF_F, CUDART_INF_F}, hi[3] = {-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F};
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j) {
    const int i = base + j * 256 + threadIdx.x;
    if (i < n) {
      float x, y, z, nx, ny, nz;
      if (vec16) {   // 16-byte aligned records: two 128-bit loads per point
        const float4 a = *reinterpret_cast<const float4*>(rec + (size_t)i * stride);
        const float4 b = *reinterpret_cast<const float4*>(rec + (size_t)i * stride + 16);
        x = a.x; y = a.y; z = a.z; nx = b.x; ny = b.y; nz = b.z;
      } else {
        const float* r = reinterpret_cast<const float*>(rec + (size_t)i * stride);
        x = r[0]; y = r[1]; z = r[2]; nx = r[4]; ny = r[5]; nz = r[6];
      }
      const bool fin = finite3(x, y, z);
      praw[i] = make_float4(x, y, z, fin ? 1.f : 0.f);
      nraw[i] = make_float4(nx, ny, nz, 0.f);
      if (fin) {
        ++cnt;
        lo[0] = fminf(lo[0], x); lo[1] = fminf(lo[1], y); lo[2] = fminf(lo[2], z);
        hi[0] = fmaxf(hi[0], x); hi[1] = fmaxf(hi[1], y); hi[2] = fmaxf(hi[2], z);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(PLO_FULL_MASK, cnt, o);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(PLO_FULL_MASK, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(PLO_FULL_MASK, hi[a], o));
    }
  }
  __shared__ float s_lo[8][3], s_hi[8][3];
  if (lane == 0) {
    s_cnt[warp] = cnt;
#pragma unroll
    for (int a = 0; a < 3; ++a) { s_lo[warp][a] = lo[a]; s_hi[warp][a] = hi[a]; }
  }
  __syncthreads();
  // one atomic per block and bound (per warp they were 47 k atomics on six addresses: the kernel's critical path)
  if (bbox != nullptr && threadIdx.x < 6) {
    const int a = threadIdx.x % 3;
    const bool is_max = threadIdx.x >= 3;
    float v = is_max ? -CUDART_INF_F : CUDART_INF_F;
    bool any = false;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
      if (s_cnt[w] > 0) { any = true; v = is_max ? fmaxf(v, s_hi[w][a]) : fminf(v, s_lo[w][a]); }
    }
    if (any) { if (is_max) atomicMax(&bbox[3 + a], f2ord(v)); else atomicMin(&bbox[a], f2ord(v)); }
  }
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_cnt[w];
    blockcnt[blockIdx.x] = t;
  }
}

__device__ __forceinline__ unsigned long long spread16(unsigned v) {
  unsigned long long x = v & 0xffffu;
  x = (x | (x << 16)) & 0x0000ff0000ffull;
  x = (x | (x << 8)) & 0x00f00f00f00full;
  x = (x | (x << 4)) & 0x0c30c30c30c3ull;
  x = (x | (x << 2)) & 0x249249249249ull;
  return x;
}

// Hilbert index of a kAxisBits-per-axis cell (Skilling's transpose algorithm).  Unlike the Z-order
// curve the Hilbert curve has no jumps: 32 (or 1024) consecutive points always form one compact blob,
// so leaf and node boxes overlap far less (measured: -27 % leaves, -34 % level-1 nodes per query).
__device__ __forceinline__ unsigned long long hilbert_key(unsigned x0, unsigned x1, unsigned x2) {
  unsigned X[3] = {x0, x1, x2};
#pragma unroll
  for (unsigned Q = 1u << (kAxisBits - 1); Q > 1u; Q >>= 1) {
    const unsigned P = Q - 1u;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      if (X[i] & Q) X[0] ^= P;
      else { const unsigned t = (X[0] ^ X[i]) & P; X[0] ^= t; X[i] ^= t; }
    }
  }
  X[1] ^= X[0];
  X[2] ^= X[1];
  unsigned t = 0u;
#pragma unroll
  for (unsigned Q = 1u << (kAxisBits - 1); Q > 1u; Q >>= 1)
    if (X[2] & Q) t ^= Q - 1u;
  X[0] ^= t; X[1] ^= t; X[2] ^= t;
  return (spread16(X[0]) << 2) | (spread16(X[1]) << 1) | spread16(X[2]);
}

// stripped-cloud index of every raw point + space-filling-curve key; vals = raw index (iota)
__global__ void __launch_bounds__(256) k_keys(const float4* __restrict__ praw, int n, const int* __restrict__ blockoff,
                                              const unsigned* __restrict__ bbox, int* __restrict__ cidx,
                                              unsigned long long* __restrict__ keys, int* __restrict__ vals) {
  PLO_CHAIN_ENTER();
  const int base = blockIdx.x * kTile;
  bool fin[kTile / 256];
  float4 p[kTile / 256];
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j) {
    const int i = base + j * 256 + threadIdx.x;
    p[j] = (i < n) ? praw[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    fin[j] = p[j].w != 0.f;
  }
  int rank[kTile / 256];
  tile_ranks(fin, rank);
  const float lox = ord2f(bbox[0]), loy = ord2f(bbox[1]), loz = ord2f(bbox[2]);
  const float ext = fmaxf(fmaxf(ord2f(bbox[3]) - lox, ord2f(bbox[4]) - loy), ord2f(bbox[5]) - loz);
  constexpr unsigned kCellMax = (1u << kAxisBits) - 1u;
  const float scale = (ext > 0.f && isfinite(ext)) ? (float)kCellMax / ext : 0.f;
  const int off = blockoff[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j) {
    const int i = base + j * 256 + threadIdx.x;
    if (i >= n) continue;
    unsigned long long key = (1ull << kKeyBits) - 1ull;
    int ci = -1;
    if (fin[j]) {
      ci = off + rank[j];
      const unsigned qx = min(kCellMax, (unsigned)fmaxf(0.f, (p[j].x - lox) * scale));
      const unsigned qy = min(kCellMax, (unsigned)fmaxf(0.f, (p[j].y - loy) * scale));
      const unsigned qz = min(kCellMax, (unsigned)fmaxf(0.f, (p[j].z - loz) * scale));
      key = hilbert_key(qx, qy, qz);
    }
    cidx[i] = ci;
    keys[i] = key;
    vals[i] = i;
  }
}

// order-preserving compaction of the source into float4 point / normal arrays
__global__ void __launch_bounds__(256) k_compact_source(const float4* __restrict__ praw, const float4* __restrict__ nraw, int n,
                                                        const int* __restrict__ blockoff, float4* __restrict__ sp,
                                                        float4* __restrict__ sn) {
  PLO_CHAIN_ENTER();
  const int base = blockIdx.x * kTile;
  bool fin[kTile / 256];
  float4 p[kTile / 256];
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j) {
    const int i = base + j * 256 + threadIdx.x;
    p[j] = (i < n) ? praw[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    fin[j] = p[j].w != 0.f;
  }
  int rank[kTile / 256];
  tile_ranks(fin, rank);
  const int off = blockoff[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kTile / 256; ++j) {
    const int i = base + j * 256 + threadIdx.x;
    if (i < n && fin[j]) {
      sp[off + rank[j]] = p[j];
      sn[off + rank[j]] = nraw[i];
    }
  }
}

// ---- LSD radix sort, one 8-bit digit per pass -------------------------------------

__global__ void __launch_bounds__(256) k_sort_hist(const unsigned long long* __restrict__ keys, int n, int shift,
                                                   int nb, int* __restrict__ hist, int* __restrict__ digit_total) {
  PLO_CHAIN_ENTER();
  __shared__ int s_h[kRadix];
  for (int d = threadIdx.x; d < kRadix; d += 256) s_h[d] = 0;
  __syncthreads();
  const int base = blockIdx.x * kSortTile;
  // every key of the thread first, then the counting: left in one loop the compiler keeps load -> atomic -> load -> ... in
  // order (the loads are predicated, the shared atomics are not moved across), i.e. 16 DRAM round trips one after the
  // other: 9 us per launch for 8 MB, measured.
  unsigned dig[kSortItems];
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const int i = base + j * 256 + threadIdx.x;
    dig[j] = (i < n) ? ((unsigned)(__ldg(&keys[i]) >> shift) & (kRadix - 1)) : (unsigned)kRadix;
  }
  // (consecutive points have similar keys, so a warp's digits collapse to a few bins; aggregating them with match.any
  // before the atomic -- -DPLO_HIST_MATCH -- was measured SLOWER than the plain shared-memory atomics: 0.241 vs 0.229 ms)
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const unsigned d = dig[j];
    const bool valid = d < (unsigned)kRadix;
#ifdef PLO_HIST_MATCH
    const unsigned peers = __match_any_sync(PLO_FULL_MASK, d);
    if (valid && (threadIdx.x & 31) == (__ffs(peers) - 1)) atomicAdd(&s_h[d], __popc(peers));
#else
    if (valid) atomicAdd(&s_h[d], 1);
#endif
  }
  __syncthreads();
  for (int d = threadIdx.x; d < kRadix; d += 256) {
    const int v = s_h[d];
    hist[d * nb + blockIdx.x] = v;
    if (v) atomicAdd(&digit_total[d], v);
  }
}

// one block per digit: base = sum of the totals of all smaller digits, then an exclusive scan of the
// digit's row (one entry per sort block) — replaces a single-block scan of the whole table
__global__ void __launch_bounds__(256) k_sort_scan(int* __restrict__ hist, int nb, const int* __restrict__ digit_total) {
  PLO_CHAIN_ENTER();
  __shared__ int s_warp[8];
  __shared__ int s_carry;
  const int d = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int part = 0;
  for (int j = threadIdx.x; j < d; j += 256) part += digit_total[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(PLO_FULL_MASK, part, o);
  if (lane == 0) s_warp[warp] = part;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_warp[w];
    s_carry = t;
  }
  __syncthreads();
  int* row = hist + (size_t)d * nb;
  for (int base = 0; base < nb; base += 256) {
    const int i = base + threadIdx.x;
    const int v = (i < nb) ? row[i] : 0;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(PLO_FULL_MASK, inc, o);
      if (lane >= o) inc += t;
    }
    __syncthreads();   // s_warp reuse
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int woff = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) woff += (w < warp) ? s_warp[w] : 0;
    const int carry = s_carry;
    if (i < nb) row[i] = carry + woff + inc - v;
    __syncthreads();
    if (threadIdx.x == 255) s_carry = carry + woff + inc;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(256) k_sort_scatter(const unsigned long long* __restrict__ keys_in,
                                                      const int* __restrict__ vals_in,
                                                      unsigned long long* __restrict__ keys_out,
                                                      int* __restrict__ vals_out, int n, int shift, int nb,
                                                      const int* __restrict__ hist_scanned) {
  PLO_CHAIN_ENTER();
  __shared__ int s_w[8][kRadix];   // per-warp running digit counts -> exclusive over warps
  __shared__ int s_g[kRadix];      // global base of (digit, this block)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int q = threadIdx.x; q < 8 * kRadix; q += 256) (&s_w[0][0])[q] = 0;
  for (int d = threadIdx.x; d < kRadix; d += 256) s_g[d] = hist_scanned[d * nb + blockIdx.x];
  __syncthreads();
  const int base = blockIdx.x * kSortTile + warp * (kSortItems * 32);
  unsigned long long key[kSortItems];
  int val[kSortItems];
  int rank[kSortItems];
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {   // all loads in flight before the (serial) ranking, see k_sort_hist
    const int i = base + j * 32 + lane;
    key[j] = (i < n) ? *reinterpret_cast<const volatile unsigned long long*>(&keys_in[i]) : 0ull;   // volatile: ptxas keeps them here
  }
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const int i = base + j * 32 + lane;
    val[j] = (i < n) ? *reinterpret_cast<const volatile int*>(&vals_in[i]) : 0;
  }
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const int i = base + j * 32 + lane;
    const bool valid = i < n;
    const unsigned d = valid ? ((unsigned)(key[j] >> shift) & (kRadix - 1)) : kRadix;  // invalid lanes group apart
    const unsigned peers = __match_any_sync(PLO_FULL_MASK, d);
    int pre = 0;
    if (valid) pre = s_w[warp][d];
    __syncwarp();
    if (valid && lane == (__ffs(peers) - 1)) s_w[warp][d] = pre + __popc(peers);
    __syncwarp();
    rank[j] = pre + __popc(peers & ((1u << lane) - 1u));
  }
  __syncthreads();
  for (int d = threadIdx.x; d < kRadix; d += 256) {
    int run = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) { int t = s_w[w][d]; s_w[w][d] = run; run += t; }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < kSortItems; ++j) {
    const int i = base + j * 32 + lane;
    if (i < n) {
      const unsigned d = (unsigned)(key[j] >> shift) & (kRadix - 1);
      const int pos = s_g[d] + s_w[warp][d] + rank[j];
      keys_out[pos] = key[j];
      vals_out[pos] = val[j];
    }
  }
}

// ---- leaves and levels --------------------------------------------------------------

// one warp per leaf: gather the sorted points/normals, emit the leaf AABB
__global__ void __launch_bounds__(256) k_gather_leaves(const int* __restrict__ vals, const float4* __restrict__ praw,
                                                       const float4* __restrict__ nraw, const int* __restrict__ cidx,
                                                       int n_raw, int n_pad, int n_leaf_pad, float4* __restrict__ pts,
                                                       float4* __restrict__ nrm, int* __restrict__ pos_of_cidx,
                                                       float4* __restrict__ lo0, float4* __restrict__ hi0) {
  PLO_CHAIN_ENTER();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int leaf = blockIdx.x * wpb + (threadIdx.x >> 5); leaf < n_leaf_pad; leaf += gridDim.x * wpb) {
    const int j = leaf * PLO_LEAF + lane;
    float4 p = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, __int_as_float(-1));
    float4 nn = make_float4(0.f, 0.f, 0.f, 0.f);
    bool fin = false;
    if (j < n_raw) {
      const int raw = vals[j];
      const float4 pr = praw[raw];
      fin = pr.w != 0.f;
      if (fin) {
        const int ci = cidx[raw];
        p = make_float4(pr.x, pr.y, pr.z, __int_as_float(ci));
        nn = nraw[raw];
        pos_of_cidx[ci] = j;
      }
    }
    if (j < n_pad) {   // leaves beyond the last real one only get an (empty) box
      pts[j] = p;
      nrm[j] = nn;
    }
    float l0 = fin ? p.x : CUDART_INF_F, l1 = fin ? p.y : CUDART_INF_F, l2 = fin ? p.z : CUDART_INF_F;
    float h0 = fin ? p.x : -CUDART_INF_F, h1 = fin ? p.y : -CUDART_INF_F, h2 = fin ? p.z : -CUDART_INF_F;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l0 = fminf(l0, __shfl_xor_sync(PLO_FULL_MASK, l0, o));
      l1 = fminf(l1, __shfl_xor_sync(PLO_FULL_MASK, l1, o));
      l2 = fminf(l2, __shfl_xor_sync(PLO_FULL_MASK, l2, o));
      h0 = fmaxf(h0, __shfl_xor_sync(PLO_FULL_MASK, h0, o));
      h1 = fmaxf(h1, __shfl_xor_sync(PLO_FULL_MASK, h1, o));
      h2 = fmaxf(h2, __shfl_xor_sync(PLO_FULL_MASK, h2, o));
    }
    if (lane == 0) {
      lo0[leaf] = make_float4(l0, l1, l2, 0.f);
      hi0[leaf] = make_float4(h0, h1, h2, 0.f);
    }
  }
}

// one warp per parent node: union of its 32 children
__global__ void __launch_bounds__(256) k_build_level(const float4* __restrict__ lo_c, const float4* __restrict__ hi_c,
                                                     int n_child_pad, float4* __restrict__ lo_p,
                                                     float4* __restrict__ hi_p, int n_par_pad) {
  PLO_CHAIN_ENTER();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int par = blockIdx.x * wpb + (threadIdx.x >> 5); par < n_par_pad; par += gridDim.x * wpb) {
    const int ch = par * PLO_FANOUT + lane;
    float4 l = make_float4(CUDART_INF_F, CUDART_INF_F, CUDART_INF_F, 0.f);
    float4 h = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, 0.f);
    if (ch < n_child_pad) { l = lo_c[ch]; h = hi_c[ch]; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      l.x = fminf(l.x, __shfl_xor_sync(PLO_FULL_MASK, l.x, o));
      l.y = fminf(l.y, __shfl_xor_sync(PLO_FULL_MASK, l.y, o));
      l.z = fminf(l.z, __shfl_xor_sync(PLO_FULL_MASK, l.z, o));
      h.x = fmaxf(h.x, __shfl_xor_sync(PLO_FULL_MASK, h.x, o));
      h.y = fmaxf(h.y, __shfl_xor_sync(PLO_FULL_MASK, h.y, o));
      h.z = fmaxf(h.z, __shfl_xor_sync(PLO_FULL_MASK, h.z, o));
    }
    if (lane == 0) { lo_p[par] = l; hi_p[par] = h; }
  }
}

// ---- local map: TransformToEnd of src/laser_odometry.cpp:88-114 on the device ----

struct MapPose { double m[12]; };   // rows of [R t]: x_prev = R x_cur + t (rPose of the registration just done)

// kept frames: p' = R^T (p - t) in double, float32 store (and n' = R^T n when asked), written to the front of `out`
__global__ void __launch_bounds__(256) k_map_advance(const float4* __restrict__ in, int64_t n, float4* __restrict__ out,
                                                     const DevState* __restrict__ st, MapPose P, int identity,
                                                     int transform_normals) {
  __shared__ double T[12];
  if (threadIdx.x < 12) T[threadIdx.x] = st ? st->rPose[threadIdx.x] : P.m[threadIdx.x];
  __syncthreads();
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    float4 a = in[2 * i], b = in[2 * i + 1];
    if (!identity) {
      const double dx = __dsub_rn((double)a.x, T[3]), dy = __dsub_rn((double)a.y, T[7]), dz = __dsub_rn((double)a.z, T[11]);
      a.x = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[0], dx), __dmul_rn(T[4], dy)), __dmul_rn(T[8], dz)));
      a.y = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[1], dx), __dmul_rn(T[5], dy)), __dmul_rn(T[9], dz)));
      a.z = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[2], dx), __dmul_rn(T[6], dy)), __dmul_rn(T[10], dz)));
      if (transform_normals) {
        const double u = (double)b.x, v = (double)b.y, w = (double)b.z;
        b.x = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[0], u), __dmul_rn(T[4], v)), __dmul_rn(T[8], w)));
        b.y = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[1], u), __dmul_rn(T[5], v)), __dmul_rn(T[9], w)));
        b.z = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[2], u), __dmul_rn(T[6], v)), __dmul_rn(T[10], w)));
      }
    }
    out[2 * i] = a;
    out[2 * i + 1] = b;
  }
}

// new frame: caller's records -> packed 32-byte records behind the kept ones
__global__ void __launch_bounds__(256) k_map_append(const char* __restrict__ rec, int stride, int64_t n, int vec16,
                                                    float4* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < n; i += (int64_t)gridDim.x * 256) {
    float4 a, b;
    if (vec16) {
      a = *reinterpret_cast<const float4*>(rec + (size_t)i * stride);
      b = *reinterpret_cast<const float4*>(rec + (size_t)i * stride + 16);
    } else {
      const float* r = reinterpret_cast<const float*>(rec + (size_t)i * stride);
      a = make_float4(r[0], r[1], r[2], 0.f);
      b = make_float4(r[4], r[5], r[6], 0.f);
    }
    a.w = 0.f;
    b.w = 0.f;
    out[2 * i] = a;
    out[2 * i + 1] = b;
  }
}

inline int64_t round_up(int64_t v, int64_t m) { return (v + m - 1) / m * m; }

}  // namespace

#define LAUNCH_CHECK(c)                          \
  do {                                           \
    (c)->launches++;                             \
    PLO_CUDA((c), cudaGetLastError());           \
  } while (0)

int plo_build_index(plo_ctx* c, const void* dev_records, int64_t n, int32_t stride) {
  c->have_target = false;
  c->pca_valid = false;
  c->projected = false;
  c->prev_valid = false;
  c->n_raw_t = n;
  c->n_levels = 0;
  DevCounts* dc = c->counts.as<DevCounts>();
  if (n == 0) {
    PLO_CUDA(c, cudaMemsetAsync(&dc->n_target, 0, sizeof(int), c->stream));
    c->n_pad_t = 0;
    c->have_target = true;
    return PLO_OK;
  }
  if (n > (int64_t)1 << 30) return plo_fail(c, PLO_ERR_UNSUPPORTED, "target larger than 2^30 points");
  // level geometry (host): cnt[l] nodes, padded to a multiple of 32
  int64_t cnt[PLO_MAX_LEVELS], pad[PLO_MAX_LEVELS];
  cnt[0] = (n + PLO_LEAF - 1) / PLO_LEAF;
  pad[0] = round_up(cnt[0], PLO_FANOUT);
  int L = 1;
  while (cnt[L - 1] > PLO_FANOUT) {
    if (L == PLO_MAX_LEVELS) return plo_fail(c, PLO_ERR_UNSUPPORTED, "too many index levels");
    cnt[L] = pad[L - 1] / PLO_FANOUT;
    pad[L] = round_up(cnt[L], PLO_FANOUT);
    ++L;
  }
  const int64_t n_pad = cnt[0] * PLO_LEAF;
  c->n_pad_t = n_pad;
  const int nb = (int)((n + kTile - 1) / kTile);
  const int nbs = (int)((n + kSortTile - 1) / kSortTile);
  PLO_CUDA(c, c->t_praw.reserve(sizeof(float4) * n));
  PLO_CUDA(c, c->t_nraw.reserve(sizeof(float4) * n));
  PLO_CUDA(c, c->t_cidx.reserve(sizeof(int) * n));
  PLO_CUDA(c, c->blockcnt.reserve(sizeof(int) * (size_t)(nb + 1)));
  PLO_CUDA(c, c->bbox.reserve(sizeof(unsigned) * 8));
  for (int a = 0; a < 2; ++a) {
    PLO_CUDA(c, c->keys[a].reserve(sizeof(unsigned long long) * n));
    PLO_CUDA(c, c->vals[a].reserve(sizeof(int) * n));
  }
  PLO_CUDA(c, c->hist.reserve(sizeof(int) * (size_t)kRadix * nbs));
  PLO_CUDA(c, c->pts_sorted.reserve(sizeof(float4) * n_pad));
  PLO_CUDA(c, c->nrm_sorted.reserve(sizeof(float4) * n_pad));
  PLO_CUDA(c, c->pos_of_cidx.reserve(sizeof(int) * n));
  for (int l = 0; l < L; ++l) {
    PLO_CUDA(c, c->lvl_lo[l].reserve(sizeof(float4) * pad[l]));
    PLO_CUDA(c, c->lvl_hi[l].reserve(sizeof(float4) * pad[l]));
  }
  cudaStream_t s = c->stream;
  const int vec16 = (stride >= 32 && stride % 16 == 0 && reinterpret_cast<uintptr_t>(dev_records) % 16 == 0) ? 1 : 0;
  if (c->ev[0]) cudaEventRecord(c->ev[0], s);
  PLO_CUDA(c, plo_launch_chained(k_init_bbox, dim3(1), dim3(32), s, c->bbox.as<unsigned>()));
  LAUNCH_CHECK(c);
  PLO_CUDA(c, plo_launch_chained(k_unpack_count, dim3(nb), dim3(256), s, static_cast<const char*>(dev_records), stride, (int)n, vec16, c->t_praw.as<float4>(),
                                    c->t_nraw.as<float4>(), c->blockcnt.as<int>(), c->bbox.as<unsigned>()));
  LAUNCH_CHECK(c);
  PLO_CUDA(c, plo_launch_chained(k_scan_exclusive, dim3(1), dim3(1024), s, c->blockcnt.as<int>(), nb, &dc->n_target));
  LAUNCH_CHECK(c);
  PLO_CUDA(c, plo_launch_chained(k_keys, dim3(nb), dim3(256), s, c->t_praw.as<float4>(), (int)n, c->blockcnt.as<int>(), c->bbox.as<unsigned>(),
                            c->t_cidx.as<int>(), c->keys[0].as<unsigned long long>(), c->vals[0].as<int>()));
  LAUNCH_CHECK(c);
  int cur = 0;
  PLO_CUDA(c, c->digit_total.reserve(sizeof(int) * kPasses * kRadix));
  PLO_CUDA(c, cudaMemsetAsync(c->digit_total.p, 0, sizeof(int) * kPasses * kRadix, s));
  for (int pass = 0; pass < kPasses; ++pass) {
    const int shift = pass * kRadixBits;
    int* tot = c->digit_total.as<int>() + pass * kRadix;
    PLO_CUDA(c, plo_launch_chained(k_sort_hist, dim3(nbs), dim3(256), s, c->keys[cur].as<unsigned long long>(), (int)n, shift, nbs,
                                   c->hist.as<int>(), tot));
    LAUNCH_CHECK(c);
    PLO_CUDA(c, plo_launch_chained(k_sort_scan, dim3(kRadix), dim3(256), s, c->hist.as<int>(), nbs, tot));
    LAUNCH_CHECK(c);
    PLO_CUDA(c, plo_launch_chained(k_sort_scatter, dim3(nbs), dim3(256), s, c->keys[cur].as<unsigned long long>(), c->vals[cur].as<int>(),
                                   c->keys[cur ^ 1].as<unsigned long long>(), c->vals[cur ^ 1].as<int>(), (int)n, shift, nbs,
                                   c->hist.as<int>()));
    LAUNCH_CHECK(c);
    cur ^= 1;
  }
  {
    const int64_t warps = pad[0];
    const int blocks = (int)std::min<int64_t>((warps + 7) / 8, (int64_t)plo_grid(c, 16));
    PLO_CUDA(c, plo_launch_chained(k_gather_leaves, dim3(blocks), dim3(256), s, c->vals[cur].as<int>(), c->t_praw.as<float4>(), c->t_nraw.as<float4>(),
                                           c->t_cidx.as<int>(), (int)n, (int)n_pad, (int)pad[0], c->pts_sorted.as<float4>(),
                                           c->nrm_sorted.as<float4>(), c->pos_of_cidx.as<int>(),
                                           c->lvl_lo[0].as<float4>(), c->lvl_hi[0].as<float4>()));
    LAUNCH_CHECK(c);
  }
  for (int l = 1; l < L; ++l) {
    const int blocks = (int)std::min<int64_t>((pad[l] + 7) / 8, (int64_t)plo_grid(c, 16));
    PLO_CUDA(c, plo_launch_chained(k_build_level, dim3(blocks), dim3(256), s, c->lvl_lo[l - 1].as<float4>(), c->lvl_hi[l - 1].as<float4>(), (int)pad[l - 1],
                                         c->lvl_lo[l].as<float4>(), c->lvl_hi[l].as<float4>(), (int)pad[l]));
    LAUNCH_CHECK(c);
  }
  if (c->ev[1]) { cudaEventRecord(c->ev[1], s); c->ev_index_pending = true; }
  c->n_levels = L;
  for (int l = 0; l < L; ++l) c->level_nodes[l] = pad[l];
  c->have_target = true;
  return PLO_OK;
}

// Stable LSD radix sort of n (u64 key, i32 value) pairs on the context's stream: `passes` 10-bit digits
// starting at bit 0.  keys[0]/vals[0] hold the input; returns the index (0/1) of the buffers that hold
// the sorted output.  hist must hold kRadix * ceil(n / 4096) ints, digit_total `passes` * kRadix ints.
int plo_sort_pairs(plo_ctx* c, unsigned long long* keys[2], int* vals[2], int64_t n, int passes, int* hist, int* digit_total,
                   int* out_which, int first_shift) {
  const int nbs = (int)((n + kSortTile - 1) / kSortTile);
  cudaStream_t s = c->stream;
  PLO_CUDA(c, cudaMemsetAsync(digit_total, 0, sizeof(int) * (size_t)passes * kRadix, s));
  int cur = 0;
  for (int pass = 0; pass < passes; ++pass) {
    const int shift = first_shift + pass * kRadixBits;
    int* tot = digit_total + pass * kRadix;
    k_sort_hist<<<nbs, 256, 0, s>>>(keys[cur], (int)n, shift, nbs, hist, tot);
    LAUNCH_CHECK(c);
    k_sort_scan<<<kRadix, 256, 0, s>>>(hist, nbs, tot);
    LAUNCH_CHECK(c);
    k_sort_scatter<<<nbs, 256, 0, s>>>(keys[cur], vals[cur], keys[cur ^ 1], vals[cur ^ 1], (int)n, shift, nbs, hist);
    LAUNCH_CHECK(c);
    cur ^= 1;
  }
  *out_which = cur;
  return PLO_OK;
}

size_t plo_sort_hist_ints(int64_t n) { return (size_t)kRadix * (size_t)((n + kSortTile - 1) / kSortTile); }
size_t plo_sort_total_ints(int passes) { return (size_t)passes * kRadix; }

int plo_upload_source(plo_ctx* c, const void* dev_records, int64_t n, int32_t stride) {
  c->have_source = false;
  c->projected = false;
  c->prev_valid = false;
  c->m_raw = n;
  DevCounts* dc = c->counts.as<DevCounts>();
  if (n == 0) {
    PLO_CUDA(c, cudaMemsetAsync(&dc->n_source, 0, sizeof(int), c->stream));
    c->have_source = true;
    return PLO_OK;
  }
  if (n > (int64_t)1 << 30) return plo_fail(c, PLO_ERR_UNSUPPORTED, "source larger than 2^30 points");
  const int nb = (int)((n + kTile - 1) / kTile);
  PLO_CUDA(c, c->s_praw.reserve(sizeof(float4) * n));
  PLO_CUDA(c, c->s_nraw.reserve(sizeof(float4) * n));
  PLO_CUDA(c, c->s_p.reserve(sizeof(float4) * n));
  PLO_CUDA(c, c->s_n.reserve(sizeof(float4) * n));
  PLO_CUDA(c, c->blockcnt.reserve(sizeof(int) * (size_t)(nb + 1)));
  cudaStream_t s = c->stream;
  const int vec16 = (stride >= 32 && stride % 16 == 0 && reinterpret_cast<uintptr_t>(dev_records) % 16 == 0) ? 1 : 0;
  PLO_CUDA(c, plo_launch_chained(k_unpack_count, dim3(nb), dim3(256), s, static_cast<const char*>(dev_records), stride, (int)n, vec16, c->s_praw.as<float4>(),
                                    c->s_nraw.as<float4>(), c->blockcnt.as<int>(), nullptr));
  LAUNCH_CHECK(c);
  PLO_CUDA(c, plo_launch_chained(k_scan_exclusive, dim3(1), dim3(1024), s, c->blockcnt.as<int>(), nb, &dc->n_source));
  LAUNCH_CHECK(c);
  PLO_CUDA(c, plo_launch_chained(k_compact_source, dim3(nb), dim3(256), s, c->s_praw.as<float4>(), c->s_nraw.as<float4>(), (int)n, c->blockcnt.as<int>(),
                                      c->s_p.as<float4>(), c->s_n.as<float4>()));
  LAUNCH_CHECK(c);
  c->have_source = true;
  return PLO_OK;
}

// Local map push (accumulateTargetCloud, src/laser_odometry.cpp:116-136, WITH the TransformToEnd step the reference
// left commented out at :118-124): the queued frames move into the new frame's coordinates, the oldest frames beyond
// max_queue drop out, the new frame is appended, and the index is rebuilt over the result — all on the device, no
// host synchronisation.  The pose comes from the host (T) or straight from the device-resident loop state of the
// registration just done (pose_from_device), which keeps register -> push -> register free of host round trips.
int plo_map_push_records(plo_ctx* c, const void* dev_records, int64_t n, int32_t stride, const double* T_host_or_null,
                         bool pose_from_device, int32_t max_queue, bool transform_normals) {
  if (max_queue < 1) max_queue = 1;
  // the new queue is worked out in locals and committed only after every step that can fail has succeeded
  std::vector<int64_t> frames = c->map_frames;
  int64_t total = 0;
  for (int64_t f : frames) total += f;
  int64_t drop = 0;
  while ((int64_t)frames.size() + 1 > max_queue && !frames.empty()) {
    drop += frames.front();
    frames.erase(frames.begin());
  }
  const int64_t keep = total - drop;
  const int64_t new_total = keep + n;
  if (new_total > (int64_t)1 << 30) return plo_fail(c, PLO_ERR_UNSUPPORTED, "local map larger than 2^30 points");
  DevBuf& src = c->map_rec[c->map_cur];
  DevBuf& dst = c->map_rec[c->map_cur ^ 1];
  PLO_CUDA(c, dst.reserve(sizeof(float4) * 2 * (size_t)std::max<int64_t>(new_total, 1)));
  cudaStream_t s = c->stream;
  if (keep > 0) {
    MapPose P;
    const bool identity = !pose_from_device && T_host_or_null == nullptr;
    for (int i = 0; i < 12; ++i) P.m[i] = T_host_or_null ? T_host_or_null[i] : ((i % 5 == 0) ? 1.0 : 0.0);
    const int blocks = (int)std::min<int64_t>((keep + 255) / 256, (int64_t)plo_grid(c, 8));
    k_map_advance<<<blocks, 256, 0, s>>>(src.as<float4>() + 2 * drop, keep, dst.as<float4>(),
                                         pose_from_device ? c->state.as<DevState>() : nullptr, P, identity ? 1 : 0,
                                         transform_normals ? 1 : 0);
    LAUNCH_CHECK(c);
  }
  if (n > 0) {
    const int vec16 = (stride >= 32 && stride % 16 == 0 && reinterpret_cast<uintptr_t>(dev_records) % 16 == 0) ? 1 : 0;
    const int blocks = (int)std::min<int64_t>((n + 255) / 256, (int64_t)plo_grid(c, 8));
    k_map_append<<<blocks, 256, 0, s>>>(static_cast<const char*>(dev_records), stride, n, vec16, dst.as<float4>() + 2 * keep);
    LAUNCH_CHECK(c);
  }
  const int rc = plo_build_index(c, dst.p, new_total, 32);
  if (rc != PLO_OK) {
    // the old buffer is untouched, but the index over it is gone: the map is reset rather than left half-described
    c->map_frames.clear();
    c->err += " (plo_map_push: the local map was reset)";
    return rc;
  }
  frames.push_back(n);
  c->map_frames = frames;
  c->map_cur ^= 1;
  return PLO_OK;
}
