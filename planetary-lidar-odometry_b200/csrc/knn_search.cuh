// knn_search.cuh — exact k-NN over the curve-sorted wide BVH of index_build.cu, one warp per query: the three-phase
// search described in the header of knn_project.cu (A bound, B conservative fp32 collect, C exact fp64 select).
// Shared by the projection kernels and the PCA-normal kernel.  Internal linkage.
#pragma once

#include <float.h>
#include <math_constants.h>

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
#ifndef PLO_BLOCK_RANGES
#define PLO_BLOCK_RANGES 85        // percent of the source handed out as one contiguous range per block once the pose settles
                                   // (0 = global counter only; 70 / 85 / 92 / 100 measured, profiles/r1j_ab_experiments.txt)
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

struct __align__(16) WarpScratch {   // (16: the settled path overlays its own layout, with float4 members, on the same slice)
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

// ---- per-query candidate tiles (written here, consumed by tile_query in knn_project.cu) -----------
//
// ICP re-projects the SAME source against the SAME map several times and once the pose settles a query moves by
// millimetres.  A walk can therefore leave behind, per query, a TILE: up to kTileSlots candidate points (coordinates
// + position in the sorted arrays, 16 B each, contiguous: the next projection reads them with two coalesced 512-byte
// loads instead of walking the tree), the query position x_ref it was made from and a squared radius e2 such that
// EVERY map point outside the tile has d2(x_ref, p) > e2.  A later projection with a proven k-th-distance bound D and
// displacement delta = |x - x_ref| may use the tile instead of the tree when sqrt(D) + delta <= sqrt(e2) (in float,
// rounded against the claim): a point outside the tile is then farther than sqrt(e2) - delta >= sqrt(D) from x, i.e.
// not a candidate, so the tile filtered by the same lower-bound test is the same superset the walk would buffer, and
// phase C decides on exact fp64 values as always.  The statement is about x_ref and the map only, so a tile stays true
// for as long as the map does.
constexpr int kTileSlots = 64;
#ifndef PLO_TILE_INFLATE
#define PLO_TILE_INFLATE 2.6f   // store-mode walk of a query whose tile failed: bound (squared) = this x the k-th distance of the bound's reference
#endif
#ifndef PLO_TILE_INFLATE_ALL
#define PLO_TILE_INFLATE_ALL 2.6f   // the same for the projection that writes every query's first tile
#endif

struct TileSink {
  float4* pts;    // [kTileSlots] this query's slots: x, y, z, w = bits of the position (-1 = empty)
  float4* meta;   // x_ref.xyz, w = e2 (<= 0: no valid tile)
  bool store;     // the pose is settling: leave a tile behind (and look a little farther than needed to give it a margin)
};

// after a walk the buffer holds every point with lo <= Df.  Keep at most kTileSlots of them (threshold t <= Df
// lowered until they fit): every point outside then has d2 >= lo > t.  Then the buffer goes back to the proven
// bound `tight` for phase C; returns its new count.  Out of line: only the projections made while the pose is
// settling come here, the walk of the first projections keeps its instruction footprint.
__device__ __noinline__ int store_tile(const MapView& m, const TileSink* sink, WarpScratch* wsp, int count, float Df, int shrinks,
                                       float tight, float qx, float qy, float qz) {
  WarpScratch& ws = *wsp;
  const int lane = threadIdx.x & 31;
  __syncwarp();
  float t = Df;
  int cnt = count;
  for (int pass = 0; cnt > kTileSlots && pass < 32; ++pass) {   // 0.85^32 < 0.006: beyond that (ties at zero distance) no tile
    t = __fmul_rd(t, 0.85f);
    cnt = 0;
    for (int base = 0; base < count; base += 32) {
      const int i = base + lane;
      cnt += __popc(__ballot_sync(PLO_FULL_MASK, i < count && ws.lo[i] <= t));
    }
  }
  const bool fits = cnt <= kTileSlots;
  int o = 0;
  for (int base = 0; base < count; base += 32) {
    const int i = base + lane;
    const bool keep = fits && i < count && ws.lo[i] <= t;
    const unsigned b = __ballot_sync(PLO_FULL_MASK, keep);
    if (keep) {
      const int ps = ws.pos[i];
      const float4 p = __ldg(&m.pts[ps]);
      __stcs(&sink->pts[o + __popc(b & ((1u << lane) - 1u))], make_float4(p.x, p.y, p.z, __int_as_float(ps)));
    }
    o += __popc(b);
  }
  for (int i = o + lane; i < kTileSlots; i += 32) __stcs(&sink->pts[i], make_float4(0.f, 0.f, 0.f, __int_as_float(-1)));
  // a bound that met the massive-tie fallback (exact_shrink) no longer describes the buffer: no tile
  const bool valid = fits && shrinks < 1000 && t > 0.f && t < CUDART_INF_F;
  if (lane == 0) *sink->meta = make_float4(qx, qy, qz, valid ? t : -1.f);
  if (tight < Df) {
    __syncwarp();
    count = filter_buffer(ws, count, tight, lane);
  }
  return count;
}

// exact k-NN of q (float32 coordinates, as the reference stores the transformed point).
// Df0: float threshold derived from a proven upper bound of the k-th distance (squared), or +inf;
// with `refine` the greedy bound is evaluated as well and the walk is ordered.
// Result: lane j holds neighbour j (d2 = +inf where not filled).
template <int LEVELS, bool TILE = false>
__device__ __forceinline__ void knn_topk(const MapView& m, float qx, float qy, float qz, float Df0, bool refine, double r2,
                                         int k, bool allow_self, WarpScratch& ws, TopK& tk, SearchStats& st, int lane,
                                         const TileSink* sink = nullptr, float ref_kf = 0.f, float inflate = 1.f) {
  st.n_leaf = st.n_node = st.n_cand = 0;
  tk.d2 = CUDART_INF;
  tk.idx = -1;
  tk.pos = -1;
  if (!(isfinite(qx) && isfinite(qy) && isfinite(qz))) {
    if constexpr (TILE) { if (sink->store && lane == 0) *sink->meta = make_float4(0.f, 0.f, 0.f, -1.f); }
    return;
  }
  const float r2f_lo = __double2float_rd(r2);   // "certainly within the radius" threshold
  Collector col;
  col.Df = fminf(Df0, bound_f(r2));
  col.count = 0;
  col.appended = 0;
  col.shrinks = 0;
  float tight = CUDART_INF_F;   // the proven bound, when the walk below looks farther than it (store mode)
  if (refine) {
    const unsigned best = Greedy<LEVELS>::run(m, 0, qx, qy, qz, r2f_lo, allow_self, st, lane);
    const unsigned kth = __shfl_sync(PLO_FULL_MASK, best, k - 1);
    if (kth < 0x7f800000u) col.Df = fminf(col.Df, __fmul_ru(__uint_as_float(kth), 1.000001f));
    Collect<LEVELS, true>::run(m, 0, qx, qy, qz, r2f_lo, r2, allow_self, k, ws, col, st, lane);
  } else {
    if constexpr (TILE) {
      if (sink->store) {
        // a wider ball than the proven bound asks for, so that the tile outlives the next small moves of the query
        // (`inflate` x the reference's k-th distance, unless the proven bound is already looser)
        tight = col.Df;
        col.Df = fminf(fmaxf(tight, __fmul_ru(ref_kf, inflate)), bound_f(r2));
      }
    }
    Collect<LEVELS, false>::run(m, 0, qx, qy, qz, r2f_lo, r2, allow_self, k, ws, col, st, lane);
  }
  if constexpr (TILE) {
    if (sink->store) {
      col.count = store_tile(m, sink, &ws, col.count, col.Df, col.shrinks, tight, qx, qy, qz);
      col.Df = fminf(col.Df, tight);
    }
  }
  __syncwarp();
  if (st.on) st.n_cand = col.appended + 100000 * col.shrinks;
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

}  // namespace
