// knn_project.cu — the matcher's hot kernel: per-iteration source transform
// (src/laser_odometry.cpp:527-549) + IMLSICPMatcher::ProjSourcePtToSurface
// (src/imls_icp.cpp:496-745) with ImplicitMLSFunction (:301-483) fused in, and the
// per-map-point PCA normal pass (ComputeNormal, :753-794, call sites :411-433,:647-669).
//
// One warp per query, warp-uniform control flow over the curve-sorted wide BVH of
// index_build.cu (32 children per node, 32 points per leaf: every node test is one
// coalesced float4 pair per lane, every leaf one coalesced 512-byte load).  Exact k-NN in
// three phases:
//   A  bound: an upper bound D of the k-th neighbour distance (squared).  Triangle inequality
//      on a reference whose k-th distance is known — the same query in the previous ICP
//      iteration (temporal) or the previous query of the warp's chunk (carry; LiDAR clouds
//      arrive in scan order) — D = (sqrt(kd2_ref) + |x - x_ref|)^2, inflated by 1e-9; for the
//      carry reference also the farthest of ITS k neighbours as seen from x (k distinct points
//      within that distance: adjacent scan points share most neighbours).  Without
//      a useful reference: greedy descent to the nearest leaves and the k-th smallest
//      rounded-up distance among their points (bitonic sort / merge across the warp).
//   B  collect (fp32, conservative): depth-first walk; boxes and points are tested with
//      float arithmetic in directed rounding (lower bounds of the true distances) against D
//      rounded up, so the buffered set is a SUPERSET of {p : d2(p) <= D}.  Candidates are
//      appended to a per-warp shared-memory buffer at ballot/popc offsets — no serial
//      dependency between candidates.  If the buffer fills, the bound shrinks to the k-th
//      smallest rounded-up distance buffered so far and the buffer is compacted.
//   C  select (fp64, exact): d2 = ((dx*dx + dy*dy) + dz*dz) in double without FMA for the
//      buffered candidates only; libnabo's acceptance rule (d2 <= r*r, self-match epsilon);
//      each candidate computes its rank under the (d2, index) order (D3: ties by index);
//      ranks < k are scattered to their slot: lane j then holds the j-th neighbour.
// Exactness: a point is only ever skipped when a LOWER bound of its distance exceeds an
// UPPER bound of the k-th distance; the final order is decided on the exact fp64 values.
//
// Candidate cache (k_project only; opt-in at compile time, -DPLO_CACHE: it makes the late projections
// 12 % faster and the first three slower, a net loss on a 7-iteration registration -- see DESIGN.md 3.2).
// ICP re-projects the SAME source against the SAME map ~7 times and
// after the second iteration a query moves by millimetres, so phase B mostly re-discovers the
// leaves it found last time.  A walk therefore leaves behind, per query, up to kCacheN candidate
// positions, the query position x_ref it was made from and a radius e2 such that EVERY map point
// outside the cache has d2(x_ref, p) > e2.  A later projection with a proven k-th-distance bound D
// (phase A, temporal) and displacement delta = |x - x_ref| may skip the walk when
// sqrt(D) + delta <= sqrt(e2) (all in float, rounded against the claim): a point outside the cache
// is then farther than sqrt(e2) - delta >= sqrt(D) from x, i.e. not a candidate, so scanning the
// cached positions with the same lower-bound test yields the same superset the walk would, and
// phase C decides on exact fp64 values as always.  When the test fails the walk runs (with the
// bound inflated so that the cache gets a margin) and refreshes the cache.
//
// The 1-NN of :601-609 (no self match) is the first list entry with d2 > DBL_EPSILON;
// only if the list is full of coincident points is a second (k=1) search needed.
//
// Algorithmic bytes per source point per iteration: 24 B query + k * 24 B
// neighbours (+ 24 B pair written) = 504 / 528 B at k = 20.  Roofline: HBM (in practice
// L2: a 1 M-point map is 32 MB and stays L2-resident).
#include <float.h>
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>

#include "plo_internal.cuh"

namespace {

#ifndef PLO_KCAP
#define PLO_KCAP 120   // 16 warps x this scratch must fit the 48 KB of static shared memory
#endif
constexpr int kCap = PLO_KCAP;     // candidate buffer entries per warp
#ifndef PLO_WARPS_PER_BLOCK
#define PLO_WARPS_PER_BLOCK 16   // 2 blocks of 16 warps per SM: with block-local source ranges more warps share a neighbourhood in L1
#endif                           // (4 x 8: 2.27 ms per registration, 8 x 4: 2.40 ms, 2 x 16: 2.22 ms)
constexpr int kWarpsPerBlock = PLO_WARPS_PER_BLOCK;
// Candidate cache: measured, NOT in the default build -- compile with -DPLO_CACHE to enable
#ifdef PLO_CACHE
constexpr bool kUseCache = true;
#else
constexpr bool kUseCache = false;
#endif
constexpr int kCacheN = 64;        // cached candidate positions per query (two per lane)
#ifndef PLO_BLOCK_RANGES
#define PLO_BLOCK_RANGES 85        // percent of the source handed out as one contiguous range per block once the pose settles
                                   // (0 = global counter only; 70 / 85 / 92 / 100 measured, profiles/r1j_ab_experiments.txt)
#endif
#ifndef PLO_CACHE_INFLATE
#define PLO_CACHE_INFLATE 3.2f     // refresh walk: bound (squared) = this x the k-th distance of the bound's reference
#endif
#ifndef PLO_GREEDY_LEAVES
#define PLO_GREEDY_LEAVES 2
#endif
constexpr int kGreedyLeaves = PLO_GREEDY_LEAVES;   // leaves examined by the greedy phase-A bound
#ifndef PLO_MINB
#define PLO_MINB 2
#endif

// neighbour list: lane j (< k) holds the j-th best entry
struct TopK {
  double d2;
  int idx;   // stripped-cloud index (tie-break key, reported to the caller)
  int pos;   // position in the sorted arrays (for gathers)
};

// traversal statistics; only carried by the hooks instantiation of the kernel
struct SearchStats {
  int n_leaf, n_node, n_cand;
  bool on;
  __device__ __forceinline__ void leaf() { if (on) n_leaf++; }
  __device__ __forceinline__ void node() { if (on) n_node++; }
};

struct WarpScratch {
  double d2[kCap];   // phase C: exact distances
  int idx[kCap];     // phase C: stripped-cloud indices
  int pos[kCap];     // phase B: positions of the buffered candidates
  float lo[kCap];    // phase B: lower bounds of their squared distances
  double od2[PLO_MAX_K];
  int oidx[PLO_MAX_K];
  int opos[PLO_MAX_K];
  float new_Df;   // out-parameter of shrink_buffer (kept out of registers / local memory)
};

// ---- conservative fp32 geometry (directed rounding) ------------------------------------

__device__ __forceinline__ float box_lo2(float qx, float qy, float qz, const float4 lo, const float4 hi) {
  const float ex = fmaxf(fmaxf(__fsub_rd(lo.x, qx), __fsub_rd(qx, hi.x)), 0.f);
  const float ey = fmaxf(fmaxf(__fsub_rd(lo.y, qy), __fsub_rd(qy, hi.y)), 0.f);
  const float ez = fmaxf(fmaxf(__fsub_rd(lo.z, qz), __fsub_rd(qz, hi.z)), 0.f);
  return __fadd_rd(__fadd_rd(__fmul_rd(ex, ex), __fmul_rd(ey, ey)), __fmul_rd(ez, ez));
}

__device__ __forceinline__ float dist_lo2(float qx, float qy, float qz, const float4 p) {
  const float ax = fabsf(__fsub_rz(qx, p.x)), ay = fabsf(__fsub_rz(qy, p.y)), az = fabsf(__fsub_rz(qz, p.z));
  return __fadd_rd(__fadd_rd(__fmul_rd(ax, ax), __fmul_rd(ay, ay)), __fmul_rd(az, az));
}

// upper bound of the true squared distance from its lower bound (rel. gap of the rd chain < 1e-6)
__device__ __forceinline__ float hi_from_lo(float lo) { return __fmul_ru(lo, 1.000001f); }

// D (double) -> float threshold for lower-bound tests, with a safety margin
__device__ __forceinline__ float bound_f(double D) { return __fmul_ru(__double2float_ru(D), 1.000001f); }

// triangle-inequality bound, everything rounded up: the k points nearest to x_ref (k-th squared distance
// <= kref) are all within sqrt(kref) + |x - x_ref| of x; returns the float threshold for lower-bound tests
__device__ __forceinline__ float tri_bound(float kref, float x, float y, float z, float rx, float ry, float rz) {
  const float ax = fmaxf(fabsf(__fsub_ru(x, rx)), fabsf(__fsub_rd(x, rx)));
  const float ay = fmaxf(fabsf(__fsub_ru(y, ry)), fabsf(__fsub_rd(y, ry)));
  const float az = fmaxf(fabsf(__fsub_ru(z, rz)), fabsf(__fsub_rd(z, rz)));
  const float s2 = __fadd_ru(__fadd_ru(__fmul_ru(ax, ax), __fmul_ru(ay, ay)), __fmul_ru(az, az));
  const float rad = __fadd_ru(__fsqrt_ru(kref), __fsqrt_ru(s2));
  return __fmul_ru(__fmul_ru(rad, rad), 1.000002f);
}

__device__ __forceinline__ double dist2_exact(double qx, double qy, double qz, const float4 p) {
  const double dx = __dsub_rn(qx, (double)p.x), dy = __dsub_rn(qy, (double)p.y), dz = __dsub_rn(qz, (double)p.z);
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__device__ __forceinline__ unsigned sort32_asc(unsigned key, int lane) {
#pragma unroll
  for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
    for (int j = k2 >> 1; j > 0; j >>= 1) {
      const unsigned other = __shfl_xor_sync(PLO_FULL_MASK, key, j);
      const bool up = (lane & k2) == 0;
      const bool lower = (lane & j) == 0;
      key = (lower == up) ? min(key, other) : max(key, other);
    }
  }
  return key;
}

// 32 smallest of an ascending run `best` and an ascending run `row`, ascending
__device__ __forceinline__ unsigned merge32_low(unsigned best, unsigned row, int lane) {
  best = min(best, __shfl_sync(PLO_FULL_MASK, row, 31 - lane));   // bitonic
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const unsigned other = __shfl_xor_sync(PLO_FULL_MASK, best, j);
    best = ((lane & j) == 0) ? min(best, other) : max(best, other);
  }
  return best;
}

// sortable key of a candidate for bound purposes: rounded-up distance if the point certainly
// satisfies libnabo's acceptance rule, else +inf
__device__ __forceinline__ unsigned bound_key(float lo, float r2f_lo, bool allow_self) {
  const float hi = hi_from_lo(lo);
  const bool certain = (hi <= r2f_lo) && (allow_self || lo > 2.3e-16f);
  return certain ? __float_as_uint(hi) : 0xffffffffu;
}

// ---- phase A: greedy bound -------------------------------------------------------------

template <int LEVEL>
struct Greedy {
  // 32 smallest bound keys (ascending over the lanes) among the points of the kGreedyLeaves
  // leaves nearest to q below the greedily chosen path
  static __device__ __forceinline__ unsigned run(const MapView& m, int node, float qx, float qy, float qz, float r2f_lo,
                                                 bool allow_self, SearchStats& st, int lane) {
    const int child = node * PLO_FANOUT + lane;
    st.node();
    // heuristic score (any choice is valid): squared distance to the box centre
    const float4 lo = __ldg(&m.lo[LEVEL - 1][child]), hi = __ldg(&m.hi[LEVEL - 1][child]);
    const float cx = qx - 0.5f * (lo.x + hi.x), cy = qy - 0.5f * (lo.y + hi.y), cz = qz - 0.5f * (lo.z + hi.z);
    const float sc = cx * cx + cy * cy + cz * cz;
    unsigned key = (lo.x <= hi.x && sc == sc) ? __float_as_uint(fminf(sc, 3.0e38f)) : 0xffffffffu;
    if constexpr (LEVEL == 1) {
      unsigned best = 0xffffffffu;
#pragma unroll 1
      for (int t = 0; t < kGreedyLeaves; ++t) {
        const unsigned mn = __reduce_min_sync(PLO_FULL_MASK, key);
        if (mn >= 0x7f800000u) break;   // no (more) non-empty leaves
        const int c = __ffs(__ballot_sync(PLO_FULL_MASK, key == mn)) - 1;
        if (lane == c) key = 0xffffffffu;
        const float4 p = __ldg(&m.pts[(node * PLO_FANOUT + c) * PLO_LEAF + lane]);
        st.leaf();
        const unsigned row = sort32_asc(bound_key(dist_lo2(qx, qy, qz, p), r2f_lo, allow_self), lane);
        best = (t == 0) ? row : merge32_low(best, row, lane);
      }
      return best;
    } else {
      const unsigned mn = __reduce_min_sync(PLO_FULL_MASK, key);
      if (mn >= 0x7f800000u) return 0xffffffffu;   // only empty boxes below
      const int c = __ffs(__ballot_sync(PLO_FULL_MASK, key == mn)) - 1;
      return Greedy<LEVEL - 1>::run(m, node * PLO_FANOUT + c, qx, qy, qz, r2f_lo, allow_self, st, lane);
    }
  }
};

// ---- phase C: exact distances + rank selection -----------------------------------------

// ranks of the C candidates in ws.d2/ws.idx under the (d2, index) order; ranks < k are scattered:
// afterwards ws.o*[j] is the j-th best (d2 = +inf where fewer than k are acceptable)
__device__ __forceinline__ void rank_select(WarpScratch& ws, int C, int k, int lane) {
  ws.od2[lane] = CUDART_INF;
  ws.oidx[lane] = -1;
  ws.opos[lane] = -1;
  __syncwarp();
#ifndef PLO_NO_FASTRANK
  if (C <= 32) {
    // the common case, one candidate per lane: rank by distance alone (one broadcast load and one
    // compare per candidate); the index order is only consulted when two distances are bit-equal
    const bool own = lane < C;
    const double d = own ? ws.d2[lane] : CUDART_INF;
    const int x = own ? ws.idx[lane] : 0x7fffffff;
    int r = 0;
#pragma unroll 4
    for (int j = 0; j < C; ++j) r += (ws.d2[j] < d) ? 1 : 0;
    const unsigned same = __match_any_sync(PLO_FULL_MASK, __double_as_longlong(d));   // executed by all 32 lanes
    const bool tied = own && d < CUDART_INF && __popc(same) > 1;
    if (__any_sync(PLO_FULL_MASK, tied)) {
      if (tied)
        for (int j = 0; j < C; ++j) r += (ws.d2[j] == d && ws.idx[j] < x) ? 1 : 0;
    }
    if (own && r < k && d < CUDART_INF) { ws.od2[r] = d; ws.oidx[r] = x; ws.opos[r] = ws.pos[lane]; }
    __syncwarp();
    return;
  }
#endif
  for (int base = 0; base < C; base += 64) {
    const int i0 = base + lane, i1 = base + 32 + lane;
    const bool own0 = i0 < C, own1 = i1 < C;
    const double d0 = own0 ? ws.d2[i0] : CUDART_INF, d1 = own1 ? ws.d2[i1] : CUDART_INF;
    const int x0 = own0 ? ws.idx[i0] : 0x7fffffff, x1 = own1 ? ws.idx[i1] : 0x7fffffff;
    int r0 = 0, r1 = 0;
    if (C - base > 32) {
      for (int j = 0; j < C; ++j) {
        const double dj = ws.d2[j];
        const int ij = ws.idx[j];
        r0 += (dj < d0 || (dj == d0 && ij < x0)) ? 1 : 0;
        r1 += (dj < d1 || (dj == d1 && ij < x1)) ? 1 : 0;
      }
    } else {
      for (int j = 0; j < C; ++j) {
        const double dj = ws.d2[j];
        const int ij = ws.idx[j];
        r0 += (dj < d0 || (dj == d0 && ij < x0)) ? 1 : 0;
      }
    }
    if (own0 && r0 < k && d0 < CUDART_INF) { ws.od2[r0] = d0; ws.oidx[r0] = x0; ws.opos[r0] = ws.pos[i0]; }
    if (own1 && r1 < k && d1 < CUDART_INF) { ws.od2[r1] = d1; ws.oidx[r1] = x1; ws.opos[r1] = ws.pos[i1]; }
  }
  __syncwarp();
}

// exact fp64 distances of the buffered candidates, libnabo's acceptance rule (unacceptable -> +inf)
__device__ __forceinline__ void exact_distances(const MapView& m, WarpScratch& ws, int count, float qx, float qy, float qz,
                                                double r2, bool allow_self, int lane) {
  const double dqx = (double)qx, dqy = (double)qy, dqz = (double)qz;
  for (int base = 0; base < count; base += 32) {
    const int i = base + lane;
    if (i < count) {
      const float4 p = __ldg(&m.pts[ws.pos[i]]);
      const double d2 = dist2_exact(dqx, dqy, dqz, p);
      const bool ok = (d2 <= r2) && (allow_self || d2 > DBL_EPSILON);
      ws.d2[i] = ok ? d2 : CUDART_INF;
      ws.idx[i] = __float_as_int(p.w);
    }
  }
  __syncwarp();
}

// massive ties at the bound (more than kCap - 32 coincident points): keep the exact k best of the buffer.
// Out of line: pathological inputs only.
__device__ __noinline__ float exact_shrink(const MapView& m, WarpScratch* ws, int count, float qx, float qy, float qz, double r2,
                                           int allow_self, int k, float Df) {
  const int lane = threadIdx.x & 31;
  exact_distances(m, *ws, count, qx, qy, qz, r2, allow_self != 0, lane);
  rank_select(*ws, count, k, lane);
  if (lane < k) {
    ws->pos[lane] = ws->opos[lane];
    ws->lo[lane] = (ws->od2[lane] < CUDART_INF) ? __double2float_rd(ws->od2[lane]) : CUDART_INF_F;
  }
  const double kd = ws->od2[k - 1];
  __syncwarp();
  return (kd < CUDART_INF) ? fminf(Df, bound_f(kd)) : Df;
}

// ---- phase B: conservative collect -----------------------------------------------------

struct Collector {
  float Df;       // current float threshold for lower bounds (warp-uniform)
  int count;      // buffered candidates
  int appended;   // statistics
  int shrinks;
};

// buffer full: Df <- k-th smallest rounded-up distance among the buffered candidates that are
// certainly acceptable; buffer compacted to lo <= Df.  Out of line: rare.
__device__ __noinline__ int shrink_buffer(WarpScratch* ws, int count, float Df, float r2f_lo, int allow_self, int k) {
  const int lane = threadIdx.x & 31;
  unsigned best = 0xffffffffu;
  for (int base = 0; base < count; base += 32) {
    const int i = base + lane;
    const unsigned key = (i < count) ? bound_key(ws->lo[i], r2f_lo, allow_self != 0) : 0xffffffffu;
    const unsigned row = sort32_asc(key, lane);
    best = (base == 0) ? row : merge32_low(best, row, lane);
  }
  const unsigned kth = __shfl_sync(PLO_FULL_MASK, best, k - 1);
  if (kth < 0x7f800000u) Df = fminf(Df, __fmul_ru(__uint_as_float(kth), 1.000001f));
  int kept = 0;
  for (int base = 0; base < count; base += 32) {   // in-place stable compaction (o <= i)
    const int i = base + lane;
    float l = 0.f;
    int ps = 0;
    bool keepit = false;
    if (i < count) { l = ws->lo[i]; ps = ws->pos[i]; keepit = l <= Df; }
    const unsigned b = __ballot_sync(PLO_FULL_MASK, keepit);
    __syncwarp();
    if (keepit) {
      const int o = kept + __popc(b & ((1u << lane) - 1u));
      ws->lo[o] = l;
      ws->pos[o] = ps;
    }
    kept += __popc(b);
    __syncwarp();
  }
  if (lane == 0) ws->new_Df = Df;
  __syncwarp();
  return kept;
}

__device__ __forceinline__ void collect_leaf(const MapView& m, int leaf, float qx, float qy, float qz, float r2f_lo, double r2,
                                             bool allow_self, int k, WarpScratch& ws, Collector& col, SearchStats& st,
                                             int lane) {
  const float4 p = __ldg(&m.pts[leaf * PLO_LEAF + lane]);
  st.leaf();
  const float lo = dist_lo2(qx, qy, qz, p);
  bool pass = lo <= col.Df;
  unsigned b = __ballot_sync(PLO_FULL_MASK, pass);
  if (b == 0u) return;
  if (col.count + __popc(b) > kCap) {
    col.count = shrink_buffer(&ws, col.count, col.Df, r2f_lo, allow_self ? 1 : 0, k);
    col.Df = ws.new_Df;
    col.shrinks++;
    pass = pass && (lo <= col.Df);
    b = __ballot_sync(PLO_FULL_MASK, pass);
    if (col.count + __popc(b) > kCap) {   // still full: > kCap - 32 candidates tie at the bound
      col.Df = exact_shrink(m, &ws, col.count, qx, qy, qz, r2, allow_self ? 1 : 0, k, col.Df);
      col.count = k;   // entries with d2 = +inf among them are dropped again by phase C
      col.shrinks += 1000;
      pass = pass && (lo <= col.Df);
      b = __ballot_sync(PLO_FULL_MASK, pass);
    }
    if (b == 0u) return;
  }
  if (pass) {
    const int o = col.count + __popc(b & ((1u << lane) - 1u));
    ws.lo[o] = lo;
    ws.pos[o] = leaf * PLO_LEAF + lane;
  }
  col.count += __popc(b);
  col.appended += __popc(b);
}

// ORDERED: visit the children nearest-centre-first (any order is exact).  Worth its cost only when
// the bound is loose (no reference): the densest neighbourhood of q then tightens it early.
template <int LEVEL, bool ORDERED>
struct Collect {
  static __device__ __forceinline__ void run(const MapView& m, int node, float qx, float qy, float qz, float r2f_lo, double r2,
                                             bool allow_self, int k, WarpScratch& ws, Collector& col, SearchStats& st,
                                             int lane) {
    const int child = node * PLO_FANOUT + lane;
    st.node();
    const float4 lo = __ldg(&m.lo[LEVEL - 1][child]), hi = __ldg(&m.hi[LEVEL - 1][child]);
    const float bd = box_lo2(qx, qy, qz, lo, hi);
    unsigned key = 0u;
    if constexpr (ORDERED) {
      const float cx = qx - 0.5f * (lo.x + hi.x), cy = qy - 0.5f * (lo.y + hi.y), cz = qz - 0.5f * (lo.z + hi.z);
      key = __float_as_uint(fminf(cx * cx + cy * cy + cz * cz, 3.0e38f));
    }
    unsigned mask = __ballot_sync(PLO_FULL_MASK, bd <= col.Df);
    int shrinks_seen = col.shrinks;
    while (mask != 0u) {
      int c = __ffs(mask) - 1;
      if constexpr (ORDERED) {
        if ((mask & (mask - 1)) != 0u) {   // more than one child left
          const unsigned mn = __reduce_min_sync(PLO_FULL_MASK, ((mask >> lane) & 1u) ? key : 0xffffffffu);
          c = __ffs(__ballot_sync(PLO_FULL_MASK, ((mask >> lane) & 1u) && key == mn)) - 1;
        }
      }
      mask &= ~(1u << c);
      if constexpr (LEVEL == 1) collect_leaf(m, node * PLO_FANOUT + c, qx, qy, qz, r2f_lo, r2, allow_self, k, ws, col, st, lane);
      else Collect<LEVEL - 1, ORDERED>::run(m, node * PLO_FANOUT + c, qx, qy, qz, r2f_lo, r2, allow_self, k, ws, col, st, lane);
      if (col.shrinks != shrinks_seen) {   // the bound shrank below: re-test the remaining children
        shrinks_seen = col.shrinks;
        mask &= __ballot_sync(PLO_FULL_MASK, bd <= col.Df);
      }
    }
  }
};

// per-query candidate cache in global memory (see the file header)
struct QueryCache {
  int* pos;      // [kCacheN] positions in the sorted arrays, -1 = empty
  float4* cx;    // x_ref.xyz, w = e2 (<= 0: no valid cache)
  bool read;     // a previous projection of the same clouds wrote it
  bool widen;    // the pose is settling: a refresh walk may look farther than it must (margin for the next moves)
};

// upper bound of |a - b|
__device__ __forceinline__ float dist_hi(float x, float y, float z, float rx, float ry, float rz) {
  const float ax = fmaxf(fabsf(__fsub_ru(x, rx)), fabsf(__fsub_rd(x, rx)));
  const float ay = fmaxf(fabsf(__fsub_ru(y, ry)), fabsf(__fsub_rd(y, ry)));
  const float az = fmaxf(fabsf(__fsub_ru(z, rz)), fabsf(__fsub_rd(z, rz)));
  return __fsqrt_ru(__fadd_ru(__fadd_ru(__fmul_ru(ax, ax), __fmul_ru(ay, ay)), __fmul_ru(az, az)));
}

// phase B from the cache: the cached positions against the bound, same lower-bound test as a leaf scan
__device__ __forceinline__ void collect_cached(const MapView& m, const int* __restrict__ cpos, float qx, float qy, float qz,
                                               WarpScratch& ws, Collector& col, int lane) {
#pragma unroll
  for (int h = 0; h < kCacheN / 32; ++h) {
    const int ps = cpos[h * 32 + lane];
    bool pass = false;
    float lo = 0.f;
    if (ps >= 0) {
      lo = dist_lo2(qx, qy, qz, __ldg(&m.pts[ps]));
      pass = lo <= col.Df;
    }
    const unsigned b = __ballot_sync(PLO_FULL_MASK, pass);
    if (pass) {
      const int o = col.count + __popc(b & ((1u << lane) - 1u));
      ws.lo[o] = lo;
      ws.pos[o] = ps;
    }
    col.count += __popc(b);
  }
  col.appended += col.count;
}

// after a walk: the buffer holds every point with lo <= Df.  Keep at most kCacheN of them (threshold t <= Df
// lowered until they fit): every point outside then has d2 >= lo > t.
__device__ __forceinline__ void store_cache(const QueryCache& qc, WarpScratch& ws, const Collector& col, float qx, float qy,
                                            float qz, int lane) {
  __syncwarp();
  float t = col.Df;
  int cnt = col.count;
  for (int pass = 0; cnt > kCacheN && pass < 24; ++pass) {   // 0.8^24 < 0.005: beyond that (ties at zero distance) no cache
    t = __fmul_rd(t, 0.8f);
    cnt = 0;
    for (int base = 0; base < col.count; base += 32) {
      const int i = base + lane;
      cnt += __popc(__ballot_sync(PLO_FULL_MASK, i < col.count && ws.lo[i] <= t));
    }
  }
  int o = 0;
  for (int base = 0; base < col.count; base += 32) {
    const int i = base + lane;
    const bool keep = cnt <= kCacheN && i < col.count && ws.lo[i] <= t;
    const unsigned b = __ballot_sync(PLO_FULL_MASK, keep);
    if (keep) qc.pos[o + __popc(b & ((1u << lane) - 1u))] = ws.pos[i];
    o += __popc(b);
  }
  for (int i = o + lane; i < kCacheN; i += 32) qc.pos[i] = -1;
  // a bound that met the massive-tie fallback (exact_shrink) no longer describes the buffer: no cache
  const bool valid = cnt <= kCacheN && col.shrinks < 1000 && t > 0.f && t < CUDART_INF_F;
  if (lane == 0) *qc.cx = make_float4(qx, qy, qz, valid ? t : -1.f);
}

// keep the buffered candidates with lo <= t (stable, in place)
__device__ __forceinline__ int filter_buffer(WarpScratch& ws, int count, float t, int lane) {
  int kept = 0;
  for (int base = 0; base < count; base += 32) {
    const int i = base + lane;
    float l = 0.f;
    int ps = 0;
    bool keepit = false;
    if (i < count) { l = ws.lo[i]; ps = ws.pos[i]; keepit = l <= t; }
    const unsigned b = __ballot_sync(PLO_FULL_MASK, keepit);
    __syncwarp();
    if (keepit) {
      const int o = kept + __popc(b & ((1u << lane) - 1u));
      ws.lo[o] = l;
      ws.pos[o] = ps;
    }
    kept += __popc(b);
    __syncwarp();
  }
  return kept;
}

// exact k-NN of q (float32 coordinates, as the reference stores the transformed point).
// Df0: float threshold derived from a proven upper bound of the k-th distance (squared), or +inf;
// with `refine` the greedy bound is evaluated as well and the walk is ordered.
// Result: lane j holds neighbour j (d2 = +inf where not filled).
template <int LEVELS, bool CACHE = false>
__device__ __forceinline__ void knn_topk(const MapView& m, float qx, float qy, float qz, float Df0, bool refine, double r2,
                                         int k, bool allow_self, WarpScratch& ws, TopK& tk, SearchStats& st, int lane,
                                         const QueryCache* qc = nullptr, float ref_kf = 0.f) {
  st.n_leaf = st.n_node = st.n_cand = 0;
  tk.d2 = CUDART_INF;
  tk.idx = -1;
  tk.pos = -1;
  if (!(isfinite(qx) && isfinite(qy) && isfinite(qz))) return;
  const float r2f_lo = __double2float_rd(r2);   // "certainly within the radius" threshold
  Collector col;
  col.Df = fminf(Df0, bound_f(r2));
  col.count = 0;
  col.appended = 0;
  col.shrinks = 0;
  bool walked = true;
  if constexpr (CACHE) {
    if (qc->read && !refine) {
      const float4 cx = *qc->cx;
      // sqrt(D) + |x - x_ref| <= sqrt(e2), rounded against the claim (NaN compares false)
      if (cx.w > 0.f && __fadd_ru(__fsqrt_ru(col.Df), dist_hi(qx, qy, qz, cx.x, cx.y, cx.z)) <= __fsqrt_rd(cx.w)) {
        collect_cached(m, qc->pos, qx, qy, qz, ws, col, lane);
        walked = false;
      }
    }
  }
  if (walked) {
    if (refine) {
      const unsigned best = Greedy<LEVELS>::run(m, 0, qx, qy, qz, r2f_lo, allow_self, st, lane);
      const unsigned kth = __shfl_sync(PLO_FULL_MASK, best, k - 1);
      if (kth < 0x7f800000u) col.Df = fminf(col.Df, __fmul_ru(__uint_as_float(kth), 1.000001f));
      Collect<LEVELS, true>::run(m, 0, qx, qy, qz, r2f_lo, r2, allow_self, k, ws, col, st, lane);
      if constexpr (CACHE) store_cache(*qc, ws, col, qx, qy, qz, lane);
    } else if constexpr (CACHE) {
      // refresh walk: a wider ball, so that the cache outlives the next small moves of the query
      // (PLO_CACHE_INFLATE x the reference's k-th distance, unless the proven bound is already looser)
      const float tight = col.Df;
      if (qc->widen) col.Df = fminf(fmaxf(tight, __fmul_ru(ref_kf, PLO_CACHE_INFLATE)), bound_f(r2));
      Collect<LEVELS, false>::run(m, 0, qx, qy, qz, r2f_lo, r2, allow_self, k, ws, col, st, lane);
      store_cache(*qc, ws, col, qx, qy, qz, lane);
      if (tight < col.Df) {   // back to the proven bound for phase C
        __syncwarp();
        col.count = filter_buffer(ws, col.count, tight, lane);
        col.Df = tight;
      }
    } else {
      Collect<LEVELS, false>::run(m, 0, qx, qy, qz, r2f_lo, r2, allow_self, k, ws, col, st, lane);
    }
  }
  __syncwarp();
  if (st.on) st.n_cand = col.appended + 100000 * col.shrinks + (walked ? 0 : 50000);
  if (col.count > 32) {
    // more than a warp's worth of candidates: a float k-th bound drops most of the surplus before
    // the O(C^2 / 32) exact ranking
    col.count = shrink_buffer(&ws, col.count, col.Df, r2f_lo, allow_self ? 1 : 0, k);
  }
  exact_distances(m, ws, col.count, qx, qy, qz, r2, allow_self, lane);
  rank_select(ws, col.count, k, lane);
  tk.d2 = ws.od2[lane];
  tk.idx = ws.oidx[lane];
  tk.pos = ws.opos[lane];
  __syncwarp();
}

// the rare second search of the 1-NN rule (k = 1, no self match), out of line; result in ws.o*[0]
template <int LEVELS>
__device__ __noinline__ void knn1_noself(const MapView& m, float qx, float qy, float qz, double r2, WarpScratch* ws) {
  TopK tk;
  SearchStats st;
  st.on = false;
  knn_topk<LEVELS>(m, qx, qy, qz, CUDART_INF_F, true, r2, 1, false, *ws, tk, st, threadIdx.x & 31);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PLO_FULL_MASK, v, o);
  return v;
}

__device__ __forceinline__ bool finite3d(double a, double b, double c) { return isfinite(a) && isfinite(b) && isfinite(c); }

// `angle > thr` of src/imls_icp.cpp:444-451 / :683-692, NaN => false (point kept).
// A float estimate of the cosine (rel. error < 1e-6) decides when it is more than 1e-5 away
// from cos(thr); otherwise the reference's own formula (sqrt, divide, acos in double) does.
__device__ __forceinline__ bool angle_exceeds(double ax, double ay, double az, double bx, double by, double bz,
                                              const DevParams& P) {
  const double dot = __dadd_rn(__dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by)), __dmul_rn(az, bz));
  const double na2 = __dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az));
  const double nb2 = __dadd_rn(__dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by)), __dmul_rn(bz, bz));
  const float prod = (float)(na2 * nb2);
  if (prod > 1e-30f && prod < 1e30f) {
    const float ce = (float)dot * rsqrtf(prod);
    const float ct = (float)P.cos_thr;
    // clearly inside the cone: angle <= thr, or cos > 1 by rounding (acos = NaN, compares false): keep
    if (ce > ct + 1e-5f) return false;
    // clearly outside (and not at cos < -1, where the reference's acos is NaN and the point is kept)
    if (ce < ct - 1e-5f && ce > -0.999f) return true;
  }
  const double c = dot / (sqrt(na2) * sqrt(nb2));
  if (!(c == c)) return false;
  const double angle = acos(c) * 180.0 / 3.14159265358979323846;
  return angle > P.angle_thr;
}

struct ProjectOut {
  float4* qx;        // transformed source point (float32), w = bits of status
  float4* qy;        // projected point y (float32)
  float4* qn;        // normal of the 1-NN (float32)
  int* status;
  float* kd2f;       // k-th neighbour distance (squared, rounded up) of this projection; +inf if the list is not full
  int* cache_pos;    // [M * kCacheN] candidate cache (file header)
  float4* cache_cx;  // [M] x_ref + e2 of the cache
  // hooks
  double* height;
  int* nn1_idx;
  double* nn1_d2;
  int* nn_idx;
  double* nn_d2;
  int* search_stats;   // [M*3] leaves scanned, nodes expanded, candidates buffered (+ 100000 * shrinks)
};

template <bool PCA, int LEVELS, bool HOOKS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, PLO_MINB) k_project(const __grid_constant__ MapView m,
                                                                          const float4* __restrict__ sp,
                                                                          const float4* __restrict__ sn,
                                                                          const DevCounts* __restrict__ counts,
                                                                          const DevState* __restrict__ st, DevParams P,
                                                                          ProjectOut out, int chunk_arg,
                                                                          int* __restrict__ chunk_counter) {
  if (st->done) return;
  __shared__ WarpScratch s_ws[kWarpsPerBlock];
  const int lane = threadIdx.x & 31;
  WarpScratch& ws = s_ws[threadIdx.x >> 5];
  const int use_prev = st->use_prev;
  const int warm = st->warm;
  const int chunk = chunk_arg > 0 ? chunk_arg : st->chunk;
  const int n_src = counts->n_source;
  const int n_tgt = m.n_raw > 0 ? counts->n_target : 0;
  // rPose rows (src/laser_odometry.cpp:530-535), kept in shared memory: 24 registers less per thread
  __shared__ double T[12];
  if (threadIdx.x < 12) T[threadIdx.x] = st->rPose[threadIdx.x];
#if PLO_BLOCK_RANGES > 0
  __shared__ int s_next;
  if (threadIdx.x == 0) s_next = 0;
#endif
  __syncthreads();

  // Each warp walks chunks of `chunk` consecutive source points (handed out dynamically, one
  // atomic per chunk: per-query cost varies a lot).  Correctness never depends on the order of
  // the source points, only the quality of the carry bound does.
#if PLO_BLOCK_RANGES > 0
  // Locality: the first PLO_BLOCK_RANGES % of the source is cut into one contiguous range per block, whose warps
  // take chunks from a shared-memory counter -- the eight warps of a block then work on neighbouring scan points
  // and share leaves and boxes in L1; the rest is handed out through the global counter and evens out the tail.
  // Only once the pose is settling (short chunks, even cost per query): with the long chunks and the uneven cost of
  // the first projections the ranges unbalance the blocks (measured: second projection 0.41 -> 0.49 ms).
  const int n_static = warm ? (int)((long long)n_src * PLO_BLOCK_RANGES / 100) / chunk * chunk : 0;
  const int per_block = ((n_static + (int)gridDim.x - 1) / (int)gridDim.x + chunk - 1) / chunk * chunk;
  const int b0 = min((int)blockIdx.x * per_block, n_static), b1 = min(b0 + per_block, n_static);
  bool own_range = true;
#endif
  while (true) {
   int c0 = 0, c_end = n_src;
#if PLO_BLOCK_RANGES > 0
   if (own_range) {
     if (lane == 0) c0 = b0 + atomicAdd(&s_next, 1) * chunk;
     c0 = __shfl_sync(PLO_FULL_MASK, c0, 0);
     c_end = b1;
     if (c0 >= b1) own_range = false;
   }
   if (!own_range) {
     if (lane == 0) c0 = n_static + atomicAdd(chunk_counter, 1) * chunk;
     c0 = __shfl_sync(PLO_FULL_MASK, c0, 0);
     c_end = n_src;
   }
#else
   if (lane == 0) c0 = atomicAdd(chunk_counter, 1) * chunk;
   c0 = __shfl_sync(PLO_FULL_MASK, c0, 0);
#endif
   if (c0 >= n_src) break;
   float carry_kf = CUDART_INF_F, carry_x = 0.f, carry_y = 0.f, carry_z = 0.f;
   int carry_pos = -1;   // lane j < k: position of the previous query's j-th neighbour (valid while carry_kf is finite)
   const int c1 = min(c0 + chunk, c_end);
   for (int qi = c0; qi < c1; ++qi) {
    const float4 p = __ldg(&sp[qi]);
    const float4 nf = __ldg(&sn[qi]);
    const double px = (double)p.x, py = (double)p.y, pz = (double)p.z;
    // p' = rPose * [p;1] in double, stored as float32 (:537-539)
    const float xf = __double2float_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[0], px), __dmul_rn(T[1], py)), __dmul_rn(T[2], pz)), T[3]));
    const float yf = __double2float_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[4], px), __dmul_rn(T[5], py)), __dmul_rn(T[6], pz)), T[7]));
    const float zf = __double2float_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[8], px), __dmul_rn(T[9], py)), __dmul_rn(T[10], pz)), T[11]));
    float nxf = nf.x, nyf = nf.y, nzf = nf.z;
    if (P.transform_normal) {   // :541-548
      const double a = (double)nf.x, b = (double)nf.y, cc = (double)nf.z;
      nxf = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[0], a), __dmul_rn(T[1], b)), __dmul_rn(T[2], cc)));
      nyf = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[4], a), __dmul_rn(T[5], b)), __dmul_rn(T[6], cc)));
      nzf = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[8], a), __dmul_rn(T[9], b)), __dmul_rn(T[10], cc)));
    }
    const double qx = (double)xf, qy = (double)yf, qz = (double)zf;        // imls_icp.cpp:556
    const double xnx = (double)nxf, xny = (double)nyf, xnz = (double)nzf;   // :557

    // bounds of the k-th distance by the triangle inequality: the k points that were nearest to a
    // reference position x_ref are all within sqrt(kd2_ref) + |x - x_ref| of x.  Reference =
    // this query in the previous projection (temporal) and the previous query of the chunk (carry).
    // All in float with upward rounding (conservative); kd2 references are stored rounded up.
    float Df0 = CUDART_INF_F, ref_kf = CUDART_INF_F;
    if (use_prev) {
      const float kprev = out.kd2f[qi];
      if (kprev < CUDART_INF_F) {
        const float4 xp = out.qx[qi];
        Df0 = tri_bound(kprev, xf, yf, zf, xp.x, xp.y, xp.z);
        ref_kf = kprev;
      }
    }
    if (carry_kf < CUDART_INF_F) {
      const float Dc = tri_bound(carry_kf, xf, yf, zf, carry_x, carry_y, carry_z);
      if (Dc < Df0) { Df0 = Dc; ref_kf = carry_kf; }
#ifndef PLO_NO_CARRY_LIST
      // the previous query's k neighbours are k distinct map points: the farthest of them from x bounds the
      // k-th distance of x as well (adjacent scan points share most neighbours: far tighter than the triangle)
      float hi = 0.f;
      if (lane < P.k) hi = hi_from_lo(dist_lo2(xf, yf, zf, __ldg(&m.pts[carry_pos])));
      const float Dn = __uint_as_float(__reduce_max_sync(PLO_FULL_MASK, __float_as_uint(hi)));
      if (Dn < Df0 && Dn <= __double2float_rd(P.r2)) { Df0 = __fmul_ru(Dn, 1.000001f); ref_kf = Df0; }
#endif
    }
    if (!(Df0 == Df0)) Df0 = CUDART_INF_F;
    // a bound more than 2.5x (in distance) above its reference would buffer > 6x k candidates:
    // evaluate the greedy bound as well (and walk nearest-first)
    const bool refine = !(Df0 < CUDART_INF_F) || !(Df0 <= 6.25f * ref_kf);

    TopK tk;
    SearchStats ss;
    ss.on = HOOKS;
    QueryCache qc;
    qc.pos = out.cache_pos + (size_t)qi * kCacheN;
    qc.cx = out.cache_cx + qi;
    qc.read = use_prev != 0;
    qc.widen = warm != 0;
    if (n_tgt > 0) knn_topk<LEVELS, kUseCache>(m, xf, yf, zf, Df0, refine, P.r2, P.k, true, ws, tk, ss, lane, &qc, ref_kf);   // :372-375 ALLOW_SELF_MATCH
    else { tk.d2 = CUDART_INF; tk.idx = -1; tk.pos = -1; ss.n_leaf = ss.n_node = ss.n_cand = 0; }
    const bool has = (lane < P.k) && (tk.d2 < CUDART_INF);
    const double kd2_now = __shfl_sync(PLO_FULL_MASK, tk.d2, P.k - 1);
    const float kd2f_now = (kd2_now < CUDART_INF) ? __double2float_ru(kd2_now) : CUDART_INF_F;
    carry_kf = kd2f_now; carry_x = xf; carry_y = yf; carry_z = zf;
    carry_pos = tk.pos;

    // ---- 1-NN without self match (:601-609) ----
    int i1 = -1, pos1 = -1;
    double d1 = CUDART_INF;
    {
      const unsigned nz = __ballot_sync(PLO_FULL_MASK, has && tk.d2 > DBL_EPSILON);
      if (nz) {
        const int j1 = __ffs(nz) - 1;
        i1 = __shfl_sync(PLO_FULL_MASK, tk.idx, j1);
        pos1 = __shfl_sync(PLO_FULL_MASK, tk.pos, j1);
        d1 = __shfl_sync(PLO_FULL_MASK, tk.d2, j1);
      } else if (__popc(__ballot_sync(PLO_FULL_MASK, has)) == P.k) {
        // the list is full of points coincident with the query: search again, k = 1, no self match
        knn1_noself<LEVELS>(m, xf, yf, zf, P.r2, &ws);
        const double dd = ws.od2[0];
        if (dd < CUDART_INF) { d1 = dd; i1 = ws.oidx[0]; pos1 = ws.opos[0]; }
        __syncwarp();
      }
    }

    // ---- per-neighbour data for the IMLS sum (one neighbour per lane) ----
    double pnx = 0.0, pny = 0.0, pnz = 0.0, ddx = 0.0, ddy = 0.0, ddz = 0.0;
    bool keep = false;
    if (has) {
      const float4 pp = __ldg(&m.pts[tk.pos]);
      if (PCA) { pnx = m.nrm_pca[3 * (size_t)tk.pos]; pny = m.nrm_pca[3 * (size_t)tk.pos + 1]; pnz = m.nrm_pca[3 * (size_t)tk.pos + 2]; }
      else { const float4 nn = __ldg(&m.nrm[tk.pos]); pnx = (double)nn.x; pny = (double)nn.y; pnz = (double)nn.z; }
      ddx = __dsub_rn(qx, (double)pp.x); ddy = __dsub_rn(qy, (double)pp.y); ddz = __dsub_rn(qz, (double)pp.z);
      keep = finite3d(pnx, pny, pnz);                                           // :436-440 (:396-400 holds by construction)
      if (keep && P.angle_constraint) keep = !angle_exceeds(xnx, xny, xnz, pnx, pny, pnz, P);   // :442-451
    }

    int status = PLO_PT_OK;
    double height = CUDART_NAN;
    double n0x = CUDART_NAN, n0y = CUDART_NAN, n0z = CUDART_NAN;
    if (i1 < 0) status = PLO_PT_NO_NORMAL;                 // :612-617
    else if (d1 > P.h2) status = PLO_PT_TOO_FAR;           // :620-625
    else {
      if (PCA) { n0x = m.nrm_pca[3 * (size_t)pos1]; n0y = m.nrm_pca[3 * (size_t)pos1 + 1]; n0z = m.nrm_pca[3 * (size_t)pos1 + 2]; }
      else { const float4 nn = __ldg(&m.nrm[pos1]); n0x = (double)nn.x; n0y = (double)nn.y; n0z = (double)nn.z; }   // :630-633
      if (!finite3d(n0x, n0y, n0z)) status = PLO_PT_INVALID_NORMAL;            // :673-679
      else if (P.angle_constraint && angle_exceeds(xnx, xny, xnz, n0x, n0y, n0z, P)) status = PLO_PT_NORMAL_CONSTRAINT;   // :681-692
    }
    if (status == PLO_PT_OK) {   // warp-uniform
      const int cnt = __popc(__ballot_sync(PLO_FULL_MASK, keep));
      if (cnt < 3) status = PLO_PT_MLS_FAIL;               // :463-466, :696-701
      else {
        // :468 — the bandwidth h_max = sqrt(d2[cnt-1]) / 3 indexes the UNFILTERED sorted distance
        // list with the filtered count; -d2 / h_max / h_max == -9 * d2 / d2[cnt-1]
        const double cinv = -9.0 / __shfl_sync(PLO_FULL_MASK, tk.d2, cnt - 1);
        double w = 0.0, pr = 0.0;
        if (keep) {
          w = exp(tk.d2 * cinv);                           // :474-475 (diff_norm == d2, same arithmetic)
          pr = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(w, ddx), pnx), __dmul_rn(__dmul_rn(w, ddy), pny)), __dmul_rn(__dmul_rn(w, ddz), pnz));   // :476
        }
        const double wsum = warp_sum(w), psum = warp_sum(pr);
        height = psum / (wsum + 1e-5);                     // :480
        if (!isfinite(height)) status = PLO_PT_NAN_INF_HEIGHT;   // :703-717
      }
    }
    if (lane == 0) {
      float4 ox = make_float4(xf, yf, zf, __int_as_float(status));
      float4 oy = make_float4(0.f, 0.f, 0.f, 0.f), on = make_float4(0.f, 0.f, 0.f, 0.f);
      if (status == PLO_PT_OK) {   // :719-731
        oy.x = __double2float_rn(__dsub_rn(qx, __dmul_rn(height, n0x)));
        oy.y = __double2float_rn(__dsub_rn(qy, __dmul_rn(height, n0y)));
        oy.z = __double2float_rn(__dsub_rn(qz, __dmul_rn(height, n0z)));
        on.x = __double2float_rn(n0x); on.y = __double2float_rn(n0y); on.z = __double2float_rn(n0z);
      }
      out.qx[qi] = ox; out.qy[qi] = oy; out.qn[qi] = on;
      out.status[qi] = status;
      out.kd2f[qi] = kd2f_now;
    }
    if constexpr (HOOKS) {
      if (lane < P.k) {
        out.nn_idx[(size_t)qi * P.k + lane] = has ? tk.idx : -1;
        out.nn_d2[(size_t)qi * P.k + lane] = has ? tk.d2 : CUDART_INF;
      }
      if (lane == 0) {
        out.height[qi] = height;
        out.nn1_idx[qi] = i1;
        out.nn1_d2[qi] = d1;
        out.search_stats[3 * (size_t)qi] = ss.n_leaf;
        out.search_stats[3 * (size_t)qi + 1] = ss.n_node;
        out.search_stats[3 * (size_t)qi + 2] = ss.n_cand;
      }
    }
   }
  }
}

#include "knn_project_tile.cuh"

// ---- PCA normals: IMLSICPMatcher::ComputeNormal (src/imls_icp.cpp:753-794) -------------

// cyclic Jacobi on a symmetric 3x3; returns the unit eigenvector of the smallest eigenvalue
__device__ void smallest_eigvec3(double a00, double a01, double a02, double a11, double a12, double a22, double v[3]) {
  double A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
  double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int sweep = 0; sweep < 32; ++sweep) {
    const double off = A[0][1] * A[0][1] + A[0][2] * A[0][2] + A[1][2] * A[1][2];
    if (off == 0.0) break;
#pragma unroll
    for (int pq = 0; pq < 3; ++pq) {
      const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
      const double apq = A[p][q];
      if (apq == 0.0) continue;
      const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
      const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      const double cs = 1.0 / sqrt(t * t + 1.0), sn = t * cs;
#pragma unroll
      for (int k = 0; k < 3; ++k) { const double akp = A[k][p], akq = A[k][q]; A[k][p] = cs * akp - sn * akq; A[k][q] = sn * akp + cs * akq; }
#pragma unroll
      for (int k = 0; k < 3; ++k) { const double apk = A[p][k], aqk = A[q][k]; A[p][k] = cs * apk - sn * aqk; A[q][k] = sn * apk + cs * aqk; }
#pragma unroll
      for (int k = 0; k < 3; ++k) { const double vkp = V[k][p], vkq = V[k][q]; V[k][p] = cs * vkp - sn * vkq; V[k][q] = sn * vkp + cs * vkq; }
    }
  }
  int best = 0;
  if (A[1][1] < A[best][best]) best = 1;
  if (A[2][2] < A[best][best]) best = 2;
  v[0] = V[0][best]; v[1] = V[1][best]; v[2] = V[2][best];
}

// one warp per map point (sorted position): k_normal nearest within r_normal, no self
// match (flags = SORT_RESULTS only, :414-416); D1: a normal exists iff all slots filled.
constexpr int kPcaWarps = 8;   // its own block size: one search per map point, no block-local ranges to share

template <int LEVELS>
__global__ void __launch_bounds__(kPcaWarps * 32) k_pca_normals(const __grid_constant__ MapView m, DevParams P,
                                                                    double* __restrict__ nrm_pca, int n_pad) {
  __shared__ WarpScratch s_ws[kPcaWarps];
  WarpScratch& ws = s_ws[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int pos = blockIdx.x * wpb + (threadIdx.x >> 5); pos < n_pad; pos += gridDim.x * wpb) {
    const float4 p = __ldg(&m.pts[pos]);
    double nx = CUDART_INF, ny = CUDART_INF, nz = CUDART_INF;   // :418-421
    if (isfinite(p.x)) {
      TopK tk;
      SearchStats ss;
      ss.on = false;
      knn_topk<LEVELS>(m, p.x, p.y, p.z, CUDART_INF_F, true, P.r_normal2, P.k_normal, false, ws, tk, ss, lane);
      const bool has = (lane < P.k_normal) && (tk.d2 < CUDART_INF);
      const int cnt = __popc(__ballot_sync(PLO_FULL_MASK, has));
      if (cnt == P.k_normal) {
        double x = 0.0, y = 0.0, z = 0.0;
        if (has) { const float4 pp = __ldg(&m.pts[tk.pos]); x = (double)pp.x; y = (double)pp.y; z = (double)pp.z; }
        const double nd = (double)cnt;
        const double mx = warp_sum(x) / nd, my = warp_sum(y) / nd, mz = warp_sum(z) / nd;   // :758-763
        const double dx = has ? x - mx : 0.0, dy = has ? y - my : 0.0, dz = has ? z - mz : 0.0;
        const double c00 = warp_sum(dx * dx) / nd, c01 = warp_sum(dx * dy) / nd, c02 = warp_sum(dx * dz) / nd;   // :766-771
        const double c11 = warp_sum(dy * dy) / nd, c12 = warp_sum(dy * dz) / nd, c22 = warp_sum(dz * dz) / nd;
        double v[3];
        smallest_eigvec3(c00, c01, c02, c11, c12, c22, v);   // :776-778
        const double nn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        if (nn > 0.0) { v[0] /= nn; v[1] /= nn; v[2] /= nn; }   // :791
        if (v[2] < 0.0) { v[0] = -v[0]; v[1] = -v[1]; v[2] = -v[2]; }   // D2
        nx = v[0]; ny = v[1]; nz = v[2];
      }
    }
    if (lane == 0) {
      nrm_pca[3 * (size_t)pos] = nx;
      nrm_pca[3 * (size_t)pos + 1] = ny;
      nrm_pca[3 * (size_t)pos + 2] = nz;
    }
  }
}

}  // namespace

int plo_reserve_query_buffers(plo_ctx* c, bool hooks) {
  const size_t m = (size_t)(c->m_raw > 0 ? c->m_raw : 1);
  PLO_CUDA(c, c->q_x.reserve(sizeof(float4) * m));
  PLO_CUDA(c, c->q_y.reserve(sizeof(float4) * m));
  PLO_CUDA(c, c->q_n.reserve(sizeof(float4) * m));
  PLO_CUDA(c, c->q_status.reserve(sizeof(int) * m));
  PLO_CUDA(c, c->q_kd2.reserve(sizeof(double) * m));
  if (kUseCache) {
    PLO_CUDA(c, c->q_cache_pos.reserve(sizeof(int) * kCacheN * m));
    PLO_CUDA(c, c->q_cache_cx.reserve(sizeof(float4) * m));
  }
  if (hooks) {
    PLO_CUDA(c, c->q_height.reserve(sizeof(double) * m));
    PLO_CUDA(c, c->q_nn1_idx.reserve(sizeof(int) * m));
    PLO_CUDA(c, c->q_nn1_d2.reserve(sizeof(double) * m));
    PLO_CUDA(c, c->q_nn_idx.reserve(sizeof(int) * m * c->prm.search_number));
    PLO_CUDA(c, c->q_nn_d2.reserve(sizeof(double) * m * c->prm.search_number));
    PLO_CUDA(c, c->q_stats.reserve(sizeof(int) * 3 * m));
  }
  return PLO_OK;
}

int plo_launch_pca_normals(plo_ctx* c) {
  if (c->pca_valid || c->n_raw_t == 0) { c->pca_valid = true; return PLO_OK; }
  PLO_CUDA(c, c->nrm_pca.reserve(sizeof(double) * 3 * (size_t)c->n_pad_t));
  const int grid = plo_grid(c, 4);
  const MapView mv = c->map_view();
  double* out = c->nrm_pca.as<double>();
  switch (c->n_levels) {
    case 1: k_pca_normals<1><<<grid, kPcaWarps * 32, 0, c->stream>>>(mv, c->dprm, out, (int)c->n_pad_t); break;
    case 2: k_pca_normals<2><<<grid, kPcaWarps * 32, 0, c->stream>>>(mv, c->dprm, out, (int)c->n_pad_t); break;
    case 3: k_pca_normals<3><<<grid, kPcaWarps * 32, 0, c->stream>>>(mv, c->dprm, out, (int)c->n_pad_t); break;
    case 4: k_pca_normals<4><<<grid, kPcaWarps * 32, 0, c->stream>>>(mv, c->dprm, out, (int)c->n_pad_t); break;
    case 5: k_pca_normals<5><<<grid, kPcaWarps * 32, 0, c->stream>>>(mv, c->dprm, out, (int)c->n_pad_t); break;
    default: k_pca_normals<6><<<grid, kPcaWarps * 32, 0, c->stream>>>(mv, c->dprm, out, (int)c->n_pad_t); break;
  }
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  c->pca_valid = true;
  return PLO_OK;
}

namespace {
template <bool PCA, bool HOOKS>
void launch_project_levels(plo_ctx* c, int blocks, const ProjectOut& out, int chunk) {
  const MapView mv = c->map_view();
  const float4* sp = c->s_p.as<float4>();
  const float4* sn = c->s_n.as<float4>();
  const DevCounts* dc = c->counts.as<DevCounts>();
  const DevState* st = c->state.as<DevState>();
  int* cc = c->chunk_counter.as<int>();
  const int T = kWarpsPerBlock * 32;
  switch (c->n_levels) {   // an empty map (n_levels == 0) never walks the tree: any instantiation does
    case 0:
    case 1: k_project<PCA, 1, HOOKS><<<blocks, T, 0, c->stream>>>(mv, sp, sn, dc, st, c->dprm, out, chunk, cc); break;
    case 2: k_project<PCA, 2, HOOKS><<<blocks, T, 0, c->stream>>>(mv, sp, sn, dc, st, c->dprm, out, chunk, cc); break;
    case 3: k_project<PCA, 3, HOOKS><<<blocks, T, 0, c->stream>>>(mv, sp, sn, dc, st, c->dprm, out, chunk, cc); break;
    case 4: k_project<PCA, 4, HOOKS><<<blocks, T, 0, c->stream>>>(mv, sp, sn, dc, st, c->dprm, out, chunk, cc); break;
    case 5: k_project<PCA, 5, HOOKS><<<blocks, T, 0, c->stream>>>(mv, sp, sn, dc, st, c->dprm, out, chunk, cc); break;
    default: k_project<PCA, 6, HOOKS><<<blocks, T, 0, c->stream>>>(mv, sp, sn, dc, st, c->dprm, out, chunk, cc); break;
  }
}
template <bool PCA, int LEVELS, bool HOOKS>
cudaError_t launch_tile_one(plo_ctx* c, const ProjectOut& out) {
  auto kern = k_project_tile<PCA, LEVELS, HOOKS>;
  const size_t smem = tile_warp_bytes(c->dprm.k) * kTileWarps;
  static size_t smem_set = 0;   // per instantiation
  if (smem > smem_set) {
    const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    smem_set = smem;
  }
  int per_sm = 0;
  cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kTileWarps * 32, smem);
  if (e != cudaSuccess) return e;
  const int64_t tiles = (c->m_raw + 31) / 32;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((tiles + kTileWarps - 1) / kTileWarps,
                                                                (int64_t)c->sm_count * std::max(per_sm, 1)));
  kern<<<blocks, kTileWarps * 32, smem, c->stream>>>(c->map_view(), c->s_p.as<float4>(), c->s_n.as<float4>(), c->s_order.as<int>(),
                                                      c->counts.as<DevCounts>(), c->state.as<DevState>(), c->dprm, out,
                                                      c->chunk_counter.as<int>());
  return cudaGetLastError();
}

template <bool PCA, bool HOOKS>
cudaError_t launch_tile_levels(plo_ctx* c, const ProjectOut& out) {
  switch (c->n_levels) {
    case 0:
    case 1: return launch_tile_one<PCA, 1, HOOKS>(c, out);
    case 2: return launch_tile_one<PCA, 2, HOOKS>(c, out);
    case 3: return launch_tile_one<PCA, 3, HOOKS>(c, out);
    case 4: return launch_tile_one<PCA, 4, HOOKS>(c, out);
    case 5: return launch_tile_one<PCA, 5, HOOKS>(c, out);
    default: return launch_tile_one<PCA, 6, HOOKS>(c, out);
  }
}
}  // namespace

int plo_launch_project(plo_ctx* c, bool hooks) {
  if (c->m_raw == 0) return PLO_OK;
  ProjectOut out;
  out.qx = c->q_x.as<float4>(); out.qy = c->q_y.as<float4>(); out.qn = c->q_n.as<float4>();
  out.status = c->q_status.as<int>();
  out.kd2f = c->q_kd2.as<float>();
  out.cache_pos = c->q_cache_pos.as<int>(); out.cache_cx = c->q_cache_cx.as<float4>();
  out.height = c->q_height.as<double>(); out.nn1_idx = c->q_nn1_idx.as<int>(); out.nn1_d2 = c->q_nn1_d2.as<double>();
  out.nn_idx = c->q_nn_idx.as<int>(); out.nn_d2 = c->q_nn_d2.as<double>();
  out.search_stats = c->q_stats.as<int>();
  if (c->tile_mode) {   // lane-per-query kernel over Hilbert-ordered tiles of the source (knn_project_tile.cuh)
    PLO_CUDA(c, c->chunk_counter.reserve(sizeof(int)));
    PLO_CUDA(c, cudaMemsetAsync(c->chunk_counter.p, 0, sizeof(int), c->stream));
    if (c->dprm.use_pca_normals) PLO_CUDA(c, (hooks ? launch_tile_levels<true, true>(c, out) : launch_tile_levels<true, false>(c, out)));
    else PLO_CUDA(c, (hooks ? launch_tile_levels<false, true>(c, out) : launch_tile_levels<false, false>(c, out)));
    c->prev_valid = true;
    c->launches++;
    return PLO_OK;
  }
  // persistent grid (PLO_MINB blocks per SM); chunks of consecutive source points are fetched through
  // an atomic counter.  The chunk length comes from the device-side loop state (see k_solve_update)
  // unless the cloud is too small to fill the GPU (then 1) or the tuning knob overrides it.
  const int64_t slots = (int64_t)plo_grid(c, PLO_MINB) * kWarpsPerBlock;
  int chunk = (c->m_raw < 16 * slots) ? 1 : 0;
  if (const char* e = getenv("PLO_CHUNK")) chunk = std::max(0, atoi(e));   // tuning knob (0 = device-side policy)
  const int64_t warps = c->m_raw;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((warps + kWarpsPerBlock - 1) / kWarpsPerBlock, (int64_t)plo_grid(c, PLO_MINB)));
  PLO_CUDA(c, c->chunk_counter.reserve(sizeof(int)));
  PLO_CUDA(c, cudaMemsetAsync(c->chunk_counter.p, 0, sizeof(int), c->stream));
  if (c->dprm.use_pca_normals) {
    if (hooks) launch_project_levels<true, true>(c, blocks, out, chunk);
    else launch_project_levels<true, false>(c, blocks, out, chunk);
  } else {
    if (hooks) launch_project_levels<false, true>(c, blocks, out, chunk);
    else launch_project_levels<false, false>(c, blocks, out, chunk);
  }
  c->prev_valid = true;   // later projections of the same clouds may use this one's k-th distances
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  return PLO_OK;
}
