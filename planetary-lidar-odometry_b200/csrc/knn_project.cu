// knn_project.cu — the matcher's hot kernels: per-iteration source transform
// (src/laser_odometry.cpp:527-549) + IMLSICPMatcher::ProjSourcePtToSurface
// (src/imls_icp.cpp:496-745) with ImplicitMLSFunction (:301-483) fused in,
// and the per-map-point PCA normal pass (ComputeNormal, :753-794, call sites :411-433,:647-669).
//
// One projection (project_phase: the body of k_project, and of every iteration of k_register_loop) has two paths that
// share one per-query tail (gates, IMLS sum, output):
//
// the tree walk (cold_query) — one warp per query, warp-uniform control flow over the curve-sorted wide BVH of
// index_build.cu (32 children per node, 32 points per leaf: every node test is one coalesced float4 pair
// per lane, every leaf one coalesced 512-byte load).  Exact k-NN in three phases (knn_search.cuh):
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
// Once the pose is settling (small last step) the walk also leaves a candidate TILE per query behind
// (knn_search.cuh: 64 candidate points, x_ref, the radius e2 it is complete for).
//
// the tile path (tile_query) — the streaming form of the settled iterations: a warp takes a GROUP of up to 32 consecutive
// queries; lane = query for the per-query scalars (transform, temporal bound, tile validity
// sqrt(D) + |x - x_ref| <= sqrt(e2)), then warp = query for each valid one: the tile arrives with two coalesced
// 512-byte loads (evict-first: 135 MB per projection stream through, the map stays in L2), is filtered by the same
// lower-bound test as a leaf, exact fp64 distances and ranks decide as in phase C — no tree, no dependent loads.
// Queries whose tile does not cover the bound go to a self-flagging miss list that warps out of groups work off through
// the tree walk while the others still stream tiles; every per-query result is bitwise the same whichever path produced it.
//
// The normal equations of the pairs are reduced and solved behind a grid barrier inside k_register_loop, or by
// k_reduce_solve (p2plane_solve.cu), one launch behind k_project, in the graph / enqueue-all forms of the loop.  (A variant
// fused into the projection itself -- per-group partial sums handed from warp to warp with release tickets, the
// last block solving -- was built and measured: every device-scope fence stalls the whole SM's memory pipe, 17 k of
// them per projection cost 0.1 ms, more than the launch they save.)
//
// The 1-NN of :601-609 (no self match) is the first list entry with d2 > DBL_EPSILON;
// only if the list is full of coincident points is a second (k=1) search needed.
//
// Algorithmic bytes per source point per iteration: 24 B query + k * 24 B
// neighbours (+ 24 B pair written) = 504 / 528 B at k = 20.  Roofline: HBM (tree walk: in practice
// L2, a 1 M-point map is 32 MB and stays L2-resident; tile path: the tiles stream from HBM).
#include <float.h>
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>

#include "knn_search.cuh"
#include "p2plane_device.cuh"
#include "plo_internal.cuh"

namespace {

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


// upper bound of |a - b|
__device__ __forceinline__ float dist_hi(float x, float y, float z, float rx, float ry, float rz) {
  const float ax = fmaxf(fabsf(__fsub_ru(x, rx)), fabsf(__fsub_rd(x, rx)));
  const float ay = fmaxf(fabsf(__fsub_ru(y, ry)), fabsf(__fsub_rd(y, ry)));
  const float az = fmaxf(fabsf(__fsub_ru(z, rz)), fabsf(__fsub_rd(z, rz)));
  return __fsqrt_ru(__fadd_ru(__fadd_ru(__fmul_ru(ax, ax), __fmul_ru(ay, ay)), __fmul_ru(az, az)));
}

struct ProjectOut {
  float4* qx;        // transformed source point (float32), w = bits of status
  float4* qy;        // projected point y (float32)
  float4* qn;        // normal of the 1-NN (float32)
  int* status;
  float* kd2f;       // k-th neighbour distance (squared, rounded up) of this projection; +inf if the list is not full
  float4* tile_pts;  // [M * kTileSlots] candidate tiles (knn_search.cuh)
  float4* tile_meta; // [M] x_ref + e2 of the tile
  // hooks
  double* height;
  int* nn1_idx;
  double* nn1_d2;
  int* nn_idx;
  double* nn_d2;
  int* search_stats;   // [M*3] leaves scanned, nodes expanded, candidates buffered (+ 100000 * shrinks, + 50000: from the tile)
};

// device-side counters of one projection (settled + cold launch); all zero between projections
enum { CNT_CHUNK = 0, CNT_BLOCKS_DONE = 1, CNT_N_MISS = 2, CNT_MISS_CURSOR = 3, CNT_SETTLED_GROUP = 4, CNT_GROUPS_DONE = 5, CNT_N = 8 };

#ifndef PLO_TAIL_PCT
#define PLO_TAIL_PCT 12     // last per cent of the cloud handed out in short chunks (tree walk with long chunks)
#endif
#ifndef PLO_TAIL_CHUNK
#define PLO_TAIL_CHUNK 2
#endif
constexpr int kGroup = 32;   // most queries a warp takes at a time on the settled path (fewer when the cloud is small, see plo_launch_project)

struct LoopSync {
  int* counters;    // [CNT_N]
  int* miss_list;   // [M] queries the settled path hands to the tree walk; -1 = empty slot
};

// ---- the per-query tail shared by both kernels: 1-NN gates, per-neighbour filters, IMLS sum, output --------
// tk: lane j holds neighbour j; pp: that neighbour's coordinates.  Returns false (nothing written) only when SETTLED
// and the query needs the tree after all (list full of points coincident with the query: second search, :601-609).
template <bool PCA, int LEVELS, bool HOOKS, bool SETTLED>
__device__ __forceinline__ bool query_tail(const MapView& m, const DevParams& P, const ProjectOut& out, int qi, float xf, float yf,
                                           float zf, float nxf, float nyf, float nzf, const TopK& tk, const float4 pp,
                                           WarpScratch* ws, const SearchStats& ss, int lane, float& kd2f_out) {
  const double qx = (double)xf, qy = (double)yf, qz = (double)zf;        // imls_icp.cpp:556
  const double xnx = (double)nxf, xny = (double)nyf, xnz = (double)nzf;   // :557
  const bool has = (lane < P.k) && (tk.d2 < CUDART_INF);
  const double kd2_now = __shfl_sync(PLO_FULL_MASK, tk.d2, P.k - 1);
  const float kd2f_now = (kd2_now < CUDART_INF) ? __double2float_ru(kd2_now) : CUDART_INF_F;
  kd2f_out = kd2f_now;

  // ---- 1-NN without self match (:601-609) ----
  int i1 = -1, pos1 = -1, j1 = -1;   // j1: its place in the list (-1: it came from the second search)
  double d1 = CUDART_INF;
  {
    const unsigned nz = __ballot_sync(PLO_FULL_MASK, has && tk.d2 > DBL_EPSILON);
    if (nz) {
      j1 = __ffs(nz) - 1;
      i1 = __shfl_sync(PLO_FULL_MASK, tk.idx, j1);
      pos1 = __shfl_sync(PLO_FULL_MASK, tk.pos, j1);
      d1 = __shfl_sync(PLO_FULL_MASK, tk.d2, j1);
    } else if (__popc(__ballot_sync(PLO_FULL_MASK, has)) == P.k) {
      // the list is full of points coincident with the query: search again, k = 1, no self match
      if constexpr (SETTLED) {
        return false;
      } else {
        knn1_noself<LEVELS>(m, xf, yf, zf, P.r2, ws);
        const double dd = ws->od2[0];
        if (dd < CUDART_INF) { d1 = dd; i1 = ws->oidx[0]; pos1 = ws->opos[0]; }
        __syncwarp();
      }
    }
  }

  // ---- per-neighbour data for the IMLS sum (one neighbour per lane) ----
  double pnx = 0.0, pny = 0.0, pnz = 0.0, ddx = 0.0, ddy = 0.0, ddz = 0.0;
  bool fin = false, keep = false;
  if (has) {
    if (PCA) { pnx = m.nrm_pca[3 * (size_t)tk.pos]; pny = m.nrm_pca[3 * (size_t)tk.pos + 1]; pnz = m.nrm_pca[3 * (size_t)tk.pos + 2]; }
    else { const float4 nn = __ldg(&m.nrm[tk.pos]); pnx = (double)nn.x; pny = (double)nn.y; pnz = (double)nn.z; }
    ddx = __dsub_rn(qx, (double)pp.x); ddy = __dsub_rn(qy, (double)pp.y); ddz = __dsub_rn(qz, (double)pp.z);
    fin = finite3d(pnx, pny, pnz);                                            // :436-440 (:396-400 holds by construction)
    keep = fin;
    if (keep && P.angle_constraint) keep = !angle_exceeds(xnx, xny, xnz, pnx, pny, pnz, P);   // :442-451
  }
  const unsigned fin_mask = __ballot_sync(PLO_FULL_MASK, fin), keep_mask = __ballot_sync(PLO_FULL_MASK, keep);

  int status = PLO_PT_OK;
  double height = CUDART_NAN;
  double n0x = CUDART_NAN, n0y = CUDART_NAN, n0z = CUDART_NAN;
  if (i1 < 0) status = PLO_PT_NO_NORMAL;                 // :612-617
  else if (d1 > P.h2) status = PLO_PT_TOO_FAR;           // :620-625
  else {
    if (PCA) { n0x = m.nrm_pca[3 * (size_t)pos1]; n0y = m.nrm_pca[3 * (size_t)pos1 + 1]; n0z = m.nrm_pca[3 * (size_t)pos1 + 2]; }
    else { const float4 nn = __ldg(&m.nrm[pos1]); n0x = (double)nn.x; n0y = (double)nn.y; n0z = (double)nn.z; }   // :630-633
    if (j1 >= 0) {
      // the 1-NN is neighbour j1 of the list: its normal went through the same two tests on lane j1 (same operands)
      if (!((fin_mask >> j1) & 1u)) status = PLO_PT_INVALID_NORMAL;            // :673-679
      else if (!((keep_mask >> j1) & 1u)) status = PLO_PT_NORMAL_CONSTRAINT;   // :681-692
    } else {
      if (!finite3d(n0x, n0y, n0z)) status = PLO_PT_INVALID_NORMAL;
      else if (P.angle_constraint && angle_exceeds(xnx, xny, xnz, n0x, n0y, n0z, P)) status = PLO_PT_NORMAL_CONSTRAINT;
    }
  }
  if (status == PLO_PT_OK) {   // warp-uniform
    const int cnt = __popc(keep_mask);
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
      // both butterfly sums (offsets 16, 8, 4, 2, 1 -- the order the oracle follows) in one: after the first exchange the
      // lower half-warp carries the weights, the upper half the projections; same operand pairs, same bits
      const bool lower = lane < 16;
      double acc = (lower ? w : pr) + __shfl_xor_sync(PLO_FULL_MASK, lower ? pr : w, 16);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) acc += __shfl_xor_sync(PLO_FULL_MASK, acc, o);
      const double wsum = __shfl_sync(PLO_FULL_MASK, acc, 0), psum = __shfl_sync(PLO_FULL_MASK, acc, 16);
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
  return true;
}

// p' = rPose * [p;1] in double, stored as float32 (src/laser_odometry.cpp:530-539); normals rotated on request (:541-548)
__device__ __forceinline__ void transform_query(const double* __restrict__ T, const float4 p, const float4 nf, int transform_normal,
                                                float& xf, float& yf, float& zf, float& nxf, float& nyf, float& nzf) {
  const double px = (double)p.x, py = (double)p.y, pz = (double)p.z;
  xf = __double2float_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[0], px), __dmul_rn(T[1], py)), __dmul_rn(T[2], pz)), T[3]));
  yf = __double2float_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[4], px), __dmul_rn(T[5], py)), __dmul_rn(T[6], pz)), T[7]));
  zf = __double2float_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[8], px), __dmul_rn(T[9], py)), __dmul_rn(T[10], pz)), T[11]));
  nxf = nf.x; nyf = nf.y; nzf = nf.z;
  if (transform_normal) {
    const double a = (double)nf.x, b = (double)nf.y, cc = (double)nf.z;
    nxf = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[0], a), __dmul_rn(T[1], b)), __dmul_rn(T[2], cc)));
    nyf = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[4], a), __dmul_rn(T[5], b)), __dmul_rn(T[6], cc)));
    nzf = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[8], a), __dmul_rn(T[9], b)), __dmul_rn(T[10], cc)));
  }
}

// ---- the tree walk of one query (cold path) -----------------------------------------------------------------

// carry reference: the previous query of the warp's chunk
struct Carry {
  float kf, x, y, z;
  int pos;   // lane j < k: position of the previous query's j-th neighbour (valid while kf is finite)
};

template <bool PCA, int LEVELS, bool HOOKS>
__device__ __forceinline__ void cold_query(const MapView& m, const float4* __restrict__ sp, const float4* __restrict__ sn,
                                           const DevParams& P, const ProjectOut& out, const double* __restrict__ T, int qi,
                                           int use_prev, int n_tgt, bool store, float inflate, WarpScratch& ws, Carry& cy, int lane) {
  const float4 p = __ldg(&sp[qi]);
  const float4 nf = __ldg(&sn[qi]);
  float xf, yf, zf, nxf, nyf, nzf;
  transform_query(T, p, nf, P.transform_normal, xf, yf, zf, nxf, nyf, nzf);

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
  if (cy.kf < CUDART_INF_F) {
    const float Dc = tri_bound(cy.kf, xf, yf, zf, cy.x, cy.y, cy.z);
    if (Dc < Df0) { Df0 = Dc; ref_kf = cy.kf; }
#ifndef PLO_NO_CARRY_LIST
    // the previous query's k neighbours are k distinct map points: the farthest of them from x bounds the
    // k-th distance of x as well (adjacent scan points share most neighbours: far tighter than the triangle)
    float hi = 0.f;
    if (lane < P.k) hi = hi_from_lo(dist_lo2(xf, yf, zf, __ldg(&m.pts[cy.pos])));
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
  TileSink sink;
  sink.pts = out.tile_pts + (size_t)qi * kTileSlots;
  sink.meta = out.tile_meta + qi;
  sink.store = store;
  if (n_tgt > 0) knn_topk<LEVELS, true>(m, xf, yf, zf, Df0, refine, P.r2, P.k, true, ws, tk, ss, lane, &sink, ref_kf, inflate);   // :372-375 ALLOW_SELF_MATCH
  else {
    tk.d2 = CUDART_INF; tk.idx = -1; tk.pos = -1; ss.n_leaf = ss.n_node = ss.n_cand = 0;
    if (store && lane == 0) *sink.meta = make_float4(0.f, 0.f, 0.f, -1.f);
  }
  float4 pp = make_float4(0.f, 0.f, 0.f, 0.f);
  if (lane < P.k && tk.d2 < CUDART_INF) pp = __ldg(&m.pts[tk.pos]);
  float kd2f_now;
  query_tail<PCA, LEVELS, HOOKS, false>(m, P, out, qi, xf, yf, zf, nxf, nyf, nzf, tk, pp, &ws, ss, lane, kd2f_now);
  cy.kf = kd2f_now; cy.x = xf; cy.y = yf; cy.z = zf;
  cy.pos = tk.pos;
}

// ---- a query answered from its candidate tile (settled path) -----------------------------------------------

struct __align__(16) TileScratch {
  double d2[kTileSlots];    // exact distances of the candidates that passed the bound (+inf: not acceptable)
  float4 xyz[kTileSlots];   // their coordinates, w = bits of the position in the sorted arrays
  int idx[kTileSlots];      // stripped-cloud indices (fetched only when two distances tie, or for the hooks)
  double od2[PLO_MAX_K];    // rank r -> distance
  int oslot[PLO_MAX_K];     // rank r -> candidate slot
};
static_assert(sizeof(TileScratch) <= sizeof(WarpScratch), "the two per-warp scratch layouts share one shared-memory slice");

// ranks of the C candidates under the (d2, index) order; ranks < k are scattered: ts.od2[r] / ts.oslot[r]
__device__ __forceinline__ void rank_tile(const MapView& m, TileScratch& ts, int C, int k, int lane) {
  ts.od2[lane] = CUDART_INF;
  ts.oslot[lane] = -1;
  __syncwarp();
  const bool own0 = lane < C, own1 = lane + 32 < C;
  const double d0 = own0 ? ts.d2[lane] : CUDART_INF, d1 = own1 ? ts.d2[lane + 32] : CUDART_INF;
  int r0 = 0, r1 = 0;
  bool need_idx = C > 32;
  if (!need_idx) {
    // the common case, one candidate per lane: rank by distance alone; the index order is only
    // consulted when two distances are bit-equal
#pragma unroll 4
    for (int j = 0; j < C; ++j) r0 += (ts.d2[j] < d0) ? 1 : 0;
    const unsigned same = __match_any_sync(PLO_FULL_MASK, __double_as_longlong(d0));
    need_idx = __any_sync(PLO_FULL_MASK, own0 && d0 < CUDART_INF && __popc(same) > 1);
  }
  if (need_idx) {
    for (int i = lane; i < C; i += 32) ts.idx[i] = __float_as_int(__ldg(&m.pts[__float_as_int(ts.xyz[i].w)]).w);
    __syncwarp();
    const int x0 = own0 ? ts.idx[lane] : 0x7fffffff, x1 = own1 ? ts.idx[lane + 32] : 0x7fffffff;
    r0 = 0;
    for (int j = 0; j < C; ++j) {
      const double dj = ts.d2[j];
      const int ij = ts.idx[j];
      r0 += (dj < d0 || (dj == d0 && ij < x0)) ? 1 : 0;
      r1 += (dj < d1 || (dj == d1 && ij < x1)) ? 1 : 0;
    }
  }
  if (own0 && r0 < k && d0 < CUDART_INF) { ts.od2[r0] = d0; ts.oslot[r0] = lane; }
  if (own1 && r1 < k && d1 < CUDART_INF) { ts.od2[r1] = d1; ts.oslot[r1] = lane + 32; }
  __syncwarp();
}

// the eight 128-byte lines of a tile into L2, one line per lane
__device__ __forceinline__ void prefetch_tile(const float4* tp, int lane) {
  if (lane < kTileSlots / 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(tp + lane * 8));
}

// One query whose tile covers its bound bD: the tile (two coalesced 512-byte loads, evict-first: the tiles stream
// through, the map stays in L2) filtered by the same lower-bound test as a leaf, exact fp64 distances, ranks, tail.
// Returns false when the query needs the tree after all (query_tail).
template <bool PCA, bool HOOKS>
__device__ __forceinline__ bool tile_query(const MapView& m, const DevParams& P, const ProjectOut& out, int qj, float bx, float by,
                                           float bz, float bnx, float bny, float bnz, float bD, TileScratch& ts, int lane) {
  const unsigned lt = (1u << lane) - 1u;
  const float4* tp = out.tile_pts + (size_t)qj * kTileSlots;
  const float4 c0 = __ldcs(tp + lane), c1 = __ldcs(tp + 32 + lane);
  const bool pass0 = __float_as_int(c0.w) >= 0 && dist_lo2(bx, by, bz, c0) <= bD;
  const bool pass1 = __float_as_int(c1.w) >= 0 && dist_lo2(bx, by, bz, c1) <= bD;
  const unsigned b0 = __ballot_sync(PLO_FULL_MASK, pass0), b1 = __ballot_sync(PLO_FULL_MASK, pass1);
  const int C = __popc(b0) + __popc(b1);
  const double dqx = (double)bx, dqy = (double)by, dqz = (double)bz;
  if (pass0) {   // exact fp64 distance, libnabo's acceptance rule (:372-375: ALLOW_SELF_MATCH)
    const int o = __popc(b0 & lt);
    const double d2 = dist2_exact(dqx, dqy, dqz, c0);
    ts.d2[o] = (d2 <= P.r2) ? d2 : CUDART_INF;
    ts.xyz[o] = c0;
  }
  if (pass1) {
    const int o = __popc(b0) + __popc(b1 & lt);
    const double d2 = dist2_exact(dqx, dqy, dqz, c1);
    ts.d2[o] = (d2 <= P.r2) ? d2 : CUDART_INF;
    ts.xyz[o] = c1;
  }
  __syncwarp();
  rank_tile(m, ts, C, P.k, lane);
  TopK tk;
  tk.d2 = ts.od2[lane];
  const int slot = ts.oslot[lane];
  float4 pp = make_float4(0.f, 0.f, 0.f, 0.f);
  tk.pos = -1;
  tk.idx = -1;
  if (lane < P.k && tk.d2 < CUDART_INF) {
    pp = ts.xyz[slot];
    tk.pos = __float_as_int(pp.w);
    tk.idx = HOOKS ? __float_as_int(__ldg(&m.pts[tk.pos]).w) : 0;   // only its sign matters without the hooks
  }
  __syncwarp();
  SearchStats ss;
  ss.on = HOOKS;
  ss.n_leaf = 0; ss.n_node = 0; ss.n_cand = C + 50000;
  float kd2f_now;
  return query_tail<PCA, 1, HOOKS, true>(m, P, out, qj, bx, by, bz, bnx, bny, bnz, tk, pp, nullptr, ss, lane, kd2f_now);
}

// ---- k_project ----------------------------------------------------------------------------------------------
//
// ONE launch per projection, persistent grid (2 blocks of 16 warps per SM).
// * Without tiles (first projections of a registration): every warp walks chunks of consecutive source points
//   through the tree.
// * With tiles (left behind by a projection made while the pose was settling): a warp takes GROUPS of 32 consecutive
//   queries -- lane = query for the per-query scalars (transform, temporal bound, tile validity), then warp = query
//   for each valid one (tile_query).  Queries whose tile does not cover the bound are appended to a miss list; warps
//   that run out of groups work that list off with the tree walk WHILE the other warps still stream tiles (an entry
//   is its own ready flag: -1 until written, reset by its consumer), so the latency-bound walks of the few misses
//   hide behind the streaming of the many hits.
// Every per-query result is bitwise the same whichever path produced it.
// the loop-state values one projection works with (read once per projection, warp-uniform)
struct ProjState {
  int use_prev;   // q_x / q_kd2 hold a previous projection of the same clouds
  bool store;     // the pose is settling: short chunks, block-local ranges, tiles left behind
  bool tiles;     // every query has a tile: settled path first, the tree walk only for its misses
  int chunk;
};

// One projection by the calling block (all blocks of the persistent grid call it): see k_project.
// T: rPose rows in shared memory; s_next: the block's chunk counter in shared memory, zero on entry.
template <bool PCA, int LEVELS, bool HOOKS>
__device__ __forceinline__ void project_phase(const MapView& m, const float4* __restrict__ sp, const float4* __restrict__ sn,
                                              const DevParams& P, const ProjectOut& out, const LoopSync& L, const ProjState ps,
                                              const double* __restrict__ T, int* s_next_p, int n_src, int n_tgt, int group,
                                              WarpScratch& ws, int lane) {
  const int use_prev = ps.use_prev;
  const bool store = ps.store;
  const bool tiles = ps.tiles;
  const int chunk = ps.chunk;
  int& s_next = *s_next_p;
  if (tiles) {
    TileScratch& ts = reinterpret_cast<TileScratch&>(ws);
    const int G = (n_src + group - 1) / group;
    const unsigned lt = (1u << lane) - 1u;
    const float r2f = bound_f(P.r2);
    // ---- groups of `group` (<= 32) queries from their tiles ----
    while (true) {
      int g = 0;
      if (lane == 0) g = atomicAdd(&L.counters[CNT_SETTLED_GROUP], 1);
      g = __shfl_sync(PLO_FULL_MASK, g, 0);
      if (g >= G) break;
      // lane = query: transform, temporal bound, tile validity
      const int qi = g * group + lane;
      const bool act = lane < group && qi < n_src;
      float xf = 0.f, yf = 0.f, zf = 0.f, nxf = 0.f, nyf = 0.f, nzf = 0.f, Df = CUDART_INF_F;
      bool valid = false;
      if (act) {
        transform_query(T, __ldg(&sp[qi]), __ldg(&sn[qi]), P.transform_normal, xf, yf, zf, nxf, nyf, nzf);
        const float kprev = out.kd2f[qi];
        const float4 meta = out.tile_meta[qi];
        if (kprev < CUDART_INF_F && meta.w > 0.f) {
          const float4 xp = out.qx[qi];
          Df = fminf(tri_bound(kprev, xf, yf, zf, xp.x, xp.y, xp.z), r2f);
          // sqrt(D) + |x - x_ref| <= sqrt(e2), rounded against the claim (NaN compares false)
          valid = __fadd_ru(__fsqrt_ru(Df), dist_hi(xf, yf, zf, meta.x, meta.y, meta.z)) <= __fsqrt_rd(meta.w);
        }
      }
      const unsigned hit = __ballot_sync(PLO_FULL_MASK, valid);
      const unsigned miss = __ballot_sync(PLO_FULL_MASK, act && !valid);
      if (miss) {
        int base = 0;
        if (lane == 0) base = atomicAdd(&L.counters[CNT_N_MISS], __popc(miss));
        base = __shfl_sync(PLO_FULL_MASK, base, 0);
        if (act && !valid) __stcg(&L.miss_list[base + __popc(miss & lt)], qi);
      }
      // warp = query, for every query of the group whose tile covers its bound.  The tiles come from HBM (they stream
      // through, 135 MB per projection): the next query's tile is prefetched into L2 while this one is worked on
      if (hit) prefetch_tile(out.tile_pts + (size_t)(g * group + __ffs(hit) - 1) * kTileSlots, lane);
      for (unsigned rem = hit; rem; rem &= rem - 1u) {
        const int j = __ffs(rem) - 1;
        const int qj = g * group + j;
        if (rem & (rem - 1u)) prefetch_tile(out.tile_pts + (size_t)(g * group + __ffs(rem & (rem - 1u)) - 1) * kTileSlots, lane);
        const float bx = __shfl_sync(PLO_FULL_MASK, xf, j), by = __shfl_sync(PLO_FULL_MASK, yf, j), bz = __shfl_sync(PLO_FULL_MASK, zf, j);
        const float bnx = __shfl_sync(PLO_FULL_MASK, nxf, j), bny = __shfl_sync(PLO_FULL_MASK, nyf, j), bnz = __shfl_sync(PLO_FULL_MASK, nzf, j);
        const float bD = __shfl_sync(PLO_FULL_MASK, Df, j);
        const bool ok = tile_query<PCA, HOOKS>(m, P, out, qj, bx, by, bz, bnx, bny, bnz, bD, ts, lane);
        if (!ok && lane == 0) __stcg(&L.miss_list[atomicAdd(&L.counters[CNT_N_MISS], 1)], qj);   // needs the tree after all
      }
      __syncwarp();
      if (lane == 0) atomicAdd(&L.counters[CNT_GROUPS_DONE], 1);   // after this group's last append
    }
  }
  // ---- tree walk: every query in chunks (no tiles), or the settled path's misses as they come in ----
  // Each warp walks chunks of `chunk` consecutive source points (handed out dynamically, one
  // atomic per chunk: per-query cost varies a lot).  Correctness never depends on the order of
  // the source points, only the quality of the carry bound does.
  // Locality: once the pose is settling the first PLO_BLOCK_RANGES % of the source is cut into one contiguous range
  // per block, whose warps take chunks from a shared-memory counter -- the warps of a block then work on neighbouring
  // scan points and share leaves and boxes in L1; the rest is handed out through the global counter and evens out
  // the tail.  With the long chunks and the uneven cost of the first projections the ranges unbalance the blocks
  // (measured: second projection 0.41 -> 0.49 ms).
  {
    const float inflate = tiles ? PLO_TILE_INFLATE : PLO_TILE_INFLATE_ALL;
    const int G = (n_src + group - 1) / group;
    const int n_static = (!tiles && store && PLO_BLOCK_RANGES > 0) ? (int)((long long)n_src * PLO_BLOCK_RANGES / 100) / chunk * chunk : 0;
    const int per_block = ((n_static + (int)gridDim.x - 1) / (int)gridDim.x + chunk - 1) / chunk * chunk;
    const int b0 = min((int)blockIdx.x * per_block, n_static), b1 = min(b0 + per_block, n_static);
    bool own_range = n_static > 0;
    const int tail_chunk = chunk > PLO_TAIL_CHUNK ? PLO_TAIL_CHUNK : chunk;
    const int n_long = (chunk > tail_chunk) ? (int)((long long)(n_src - n_static) * (100 - PLO_TAIL_PCT) / 100) / chunk : 0x3fffffff;   // long chunks handed out
    const int tail0 = (chunk > tail_chunk) ? n_static + n_long * chunk : 0x7fffffff;   // first query of the short-chunk tail
    while (true) {
      int c0 = 0, c1 = 0;
      if (tiles) {
        if (lane == 0) {
          const int i = atomicAdd(&L.counters[CNT_MISS_CURSOR], 1);
          volatile int* vl = L.miss_list;
          volatile int* vc = L.counters;
          unsigned backoff = 64;
          while (true) {
            if (i >= n_src) { c0 = -2; break; }                  // the list cannot grow this long: nothing left for this warp
            c0 = vl[i];
            if (c0 >= 0) { vl[i] = -1; break; }                 // consumed: the slot is free for the next projection
            if (vc[CNT_GROUPS_DONE] >= G && i >= vc[CNT_N_MISS]) { c0 = -2; break; }   // every group is through and the list is shorter
            __nanosleep(backoff);                                // idle warps poll ever more rarely: the issue slots belong to the others
            backoff = min(backoff * 2u, 2048u);
          }
        }
        c0 = __shfl_sync(PLO_FULL_MASK, c0, 0);
        if (c0 < 0) break;
        c1 = c0 + 1;
      } else {
        int c_end = n_src;
        if (own_range) {
          if (lane == 0) c0 = b0 + atomicAdd(&s_next, 1) * chunk;
          c0 = __shfl_sync(PLO_FULL_MASK, c0, 0);
          c_end = b1;
          if (c0 >= b1) own_range = false;
        }
        int len = chunk;
        if (!own_range) {
          // long chunks for most of the cloud, short ones for its last PLO_TAIL_PCT %: the projection ends when the
          // last warp finishes its last chunk, and a long chunk picked up late is 25-50 us of tail
          if (lane == 0) {
            const int u = atomicAdd(&L.counters[CNT_CHUNK], 1);
            c0 = (u < n_long) ? n_static + u * chunk : tail0 + (u - n_long) * tail_chunk;
          }
          c0 = __shfl_sync(PLO_FULL_MASK, c0, 0);
          c_end = n_src;
          if (c0 >= tail0) len = tail_chunk;
        }
        if (c0 >= n_src) break;
        c1 = min(c0 + len, c_end);
      }
      Carry cy;
      cy.kf = CUDART_INF_F; cy.x = cy.y = cy.z = 0.f; cy.pos = -1;
      for (int qi = c0; qi < c1; ++qi) cold_query<PCA, LEVELS, HOOKS>(m, sp, sn, P, out, T, qi, use_prev, n_tgt, store, inflate, ws, cy, lane);
    }
  }

}

template <bool PCA, int LEVELS, bool HOOKS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, PLO_MINB) k_project(const __grid_constant__ MapView m,
                                                                          const float4* __restrict__ sp,
                                                                          const float4* __restrict__ sn,
                                                                          const DevCounts* __restrict__ counts,
                                                                          DevState* __restrict__ st,
                                                                          const __grid_constant__ DevParams P,
                                                                          const __grid_constant__ ProjectOut out,
                                                                          const __grid_constant__ LoopSync L, int chunk_arg, int group) {
  if (st->done) return;
  __shared__ WarpScratch s_ws[kWarpsPerBlock];
  // lane and the warp's scratch offset are made opaque: left to itself the compiler rematerialises them from the
  // special registers ~40 times per query (S2R + shifts: 6 % of the kernel's instructions)
  int lane = threadIdx.x & 31;
  unsigned ws_ofs = (threadIdx.x >> 5) * (unsigned)sizeof(WarpScratch);
  asm volatile("" : "+r"(lane), "+r"(ws_ofs));
  WarpScratch& ws = *reinterpret_cast<WarpScratch*>(reinterpret_cast<char*>(s_ws) + ws_ofs);
  ProjState ps;
  ps.use_prev = st->use_prev;
  ps.store = st->warm != 0;
  ps.tiles = ps.use_prev && st->tiles_ready;
  ps.chunk = chunk_arg > 0 ? chunk_arg : (chunk_arg < 0 ? min(st->chunk, -chunk_arg) : st->chunk);   // > 0: forced, < 0: cap
  const int n_src = counts->n_source;
  const int n_tgt = m.n_raw > 0 ? counts->n_target : 0;
  // rPose rows (src/laser_odometry.cpp:530-535), kept in shared memory: 24 registers less per thread
  __shared__ double T[12];
  __shared__ int s_next;
  if (threadIdx.x < 12) T[threadIdx.x] = st->rPose[threadIdx.x];
  if (threadIdx.x == 0) s_next = 0;
  __syncthreads();
  project_phase<PCA, LEVELS, HOOKS>(m, sp, sn, P, out, L, ps, T, &s_next, n_src, n_tgt, group, ws, lane);

  // ---- epilogue: the last block to get here resets the projection's counters ----
  __syncthreads();
  if (threadIdx.x != 0) return;
  __threadfence();
  if (atomicAdd(&L.counters[CNT_BLOCKS_DONE], 1) != (int)gridDim.x - 1) return;
  __threadfence();
  const int it = st->iters;
  if (it < 32) st->miss_hist[it] = ps.tiles ? ((volatile int*)L.counters)[CNT_N_MISS] : -1;
  if (ps.store) st->tiles_ready = 1;   // every query handled by the tree walk since the reset left a tile (or an invalid mark) behind
  for (int i = 0; i < CNT_N; ++i) L.counters[i] = 0;
}

// ---- k_register_loop: the whole ICP loop of one registration in ONE launch ------------------------------------
//
// Cooperative persistent grid (every block resident).  Per iteration: projection (project_phase) | grid barrier |
// every block reduces its fixed share of the pairs into a partial (reduce_pairs_block, the same partition and order as
// k_reduce_solve) | grid barrier | every block sums the partials in the same fixed order and runs the 6x6 solve and
// the loop tail on its own shared-memory copy of the loop state -- all copies stay bitwise identical, no third barrier,
// no launch, no graph node, no host between the iterations (src/laser_odometry.cpp:524-647 resident on the device).
// Data written by other SMs is read after a barrier whose fence invalidates the SM's L1; loads of such data never use
// the read-only path.
// One arrival counter that only grows (k_init_state zeroes it before every registration): barrier number j of a launch is
// passed when the counter reaches j * n_blocks, so there is no reset, no generation word and no second atomic on the path of
// the last arriver (measured with PLO_LOOP_TIMING behind the balanced reduce phase: 4.0 -> 3.0 us per barrier).  `target` = arrivals that complete
// THIS barrier; every block runs through the same sequence of barriers and counts it for itself.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();   // release: this block's writes (ordered before by the bar.sync above) are visible device-wide
    atomicAdd(bar, 1u);
    volatile unsigned* vb = bar;
    while (vb[0] < target) __nanosleep(20);
    __threadfence();   // acquire side: also drops this SM's L1 lines of data other SMs have rewritten
  }
  __syncthreads();
}

template <bool PCA, int LEVELS>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, PLO_MINB) k_register_loop(const __grid_constant__ MapView m,
                                                                                const float4* __restrict__ sp,
                                                                                const float4* __restrict__ sn,
                                                                                const DevCounts* __restrict__ counts,
                                                                                DevState* __restrict__ st,
                                                                                const __grid_constant__ DevParams P,
                                                                                const __grid_constant__ ProjectOut out,
                                                                                const __grid_constant__ LoopSync L, int chunk_arg,
                                                                                int group, double* __restrict__ partials,
                                                                                unsigned* __restrict__ bar) {
  __shared__ WarpScratch s_ws[kWarpsPerBlock];
  __shared__ DevState s_st;   // this block's copy of the loop state (all copies evolve identically)
  __shared__ double T[12];
  __shared__ int s_next;
  int lane = threadIdx.x & 31;
  unsigned ws_ofs = (threadIdx.x >> 5) * (unsigned)sizeof(WarpScratch);
  asm volatile("" : "+r"(lane), "+r"(ws_ofs));
  WarpScratch& ws = *reinterpret_cast<WarpScratch*>(reinterpret_cast<char*>(s_ws) + ws_ofs);
  // the reduce / solve phases reuse the search scratch: [warps][PLO_NSUM] warp sums, then the PLO_NSUM totals
  double(*s_red)[PLO_NSUM] = reinterpret_cast<double(*)[PLO_NSUM]>(s_ws);
  double* s_sum = reinterpret_cast<double*>(s_ws) + kWarpsPerBlock * PLO_NSUM;
  static_assert(sizeof(double) * (kWarpsPerBlock + 1) * PLO_NSUM <= sizeof(WarpScratch) * kWarpsPerBlock, "scratch reuse");
  for (int i = threadIdx.x; i < (int)(sizeof(DevState) / sizeof(int)); i += blockDim.x)
    reinterpret_cast<int*>(&s_st)[i] = reinterpret_cast<const int*>(st)[i];
  __syncthreads();
  const int n_src = counts->n_source;
  const int n_tgt = m.n_raw > 0 ? counts->n_target : 0;
#ifdef PLO_LOOP_TIMING
  unsigned long long* dbg = reinterpret_cast<unsigned long long*>(partials + (size_t)gridDim.x * PLO_NSUM);   // scratch behind the partials
  auto stamp = [&](int it, int ph) { if (blockIdx.x == 0 && threadIdx.x == 0 && it < 16) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); dbg[it * 8 + ph] = t; } };
#else
  auto stamp = [](int, int) {};
#endif
  unsigned bar_target = 0u;   // arrivals that complete the next grid barrier (see grid_barrier)
  while (!s_st.done) {
    const int dbg_it = s_st.iters;
    stamp(dbg_it, 0);
    ProjState ps;
    ps.use_prev = s_st.use_prev;
    ps.store = s_st.warm != 0;
    ps.tiles = ps.use_prev && s_st.tiles_ready;
    ps.chunk = chunk_arg > 0 ? chunk_arg : (chunk_arg < 0 ? min(s_st.chunk, -chunk_arg) : s_st.chunk);
    if (threadIdx.x < 12) T[threadIdx.x] = s_st.rPose[threadIdx.x];
    if (threadIdx.x == 0) s_next = 0;
    __syncthreads();
    project_phase<PCA, LEVELS, false>(m, sp, sn, P, out, L, ps, T, &s_next, n_src, n_tgt, group, ws, lane);
    stamp(dbg_it, 1);
    grid_barrier(bar, bar_target += gridDim.x);
    stamp(dbg_it, 2);
    // every query is projected.  Block 0 keeps the books while everybody reduces.
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      const int it = s_st.iters;
      if (it < 32) st->miss_hist[it] = ps.tiles ? ((volatile int*)L.counters)[CNT_N_MISS] : -1;
      for (int i = 0; i < CNT_N; ++i) L.counters[i] = 0;
    }
    reduce_pairs_block<kWarpsPerBlock>(out.qx, out.qy, out.qn, n_src, P, partials + (size_t)blockIdx.x * PLO_NSUM, s_red);
    stamp(dbg_it, 3);
    grid_barrier(bar, bar_target += gridDim.x);
    stamp(dbg_it, 4);
    sum_block_partials(partials, (int)gridDim.x, s_sum, s_red);
    stamp(dbg_it, 5);
    if (threadIdx.x == 0 && ps.store) s_st.tiles_ready = 1;
    if (threadIdx.x < 32) solve_from_sums(s_sum, &s_st, P, 1, (cudaGraphConditionalHandle)0, 0, 0);   // warp 0, collectively
    __syncthreads();
    stamp(dbg_it, 6);
  }
  if (blockIdx.x == 0) {
    // miss_hist was kept in global memory by block 0; everything else comes from the shared copy
    __syncthreads();
    for (int i = threadIdx.x; i < 32; i += blockDim.x) s_st.miss_hist[i] = st->miss_hist[i];
    __syncthreads();
    for (int i = threadIdx.x; i < (int)(sizeof(DevState) / sizeof(int)); i += blockDim.x)
      reinterpret_cast<int*>(st)[i] = reinterpret_cast<const int*>(&s_st)[i];
  }
}

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


// ---- stand-alone matcher entry points of the reference's surface ------------------------------------------------

// bool IMLSICPMatcher::ImplicitMLSFunction(PointType& x, double& height), src/imls_icp.cpp:301-483, for a batch of
// points: k-NN within r (:372-375), per-neighbour filters (:380-460), cnt >= 3 (:463-466), the bandwidth quirk (:468)
// and the IMLS sum (:470-480) -- without the 1-NN gates of ProjSourcePtToSurface.  One warp per point.
template <bool PCA, int LEVELS>
__global__ void __launch_bounds__(kPcaWarps * 32) k_imls_height(const __grid_constant__ MapView m, DevParams P,
                                                                    const DevCounts* __restrict__ counts, const float* __restrict__ pts6,
                                                                    int n, double* __restrict__ height, int* __restrict__ ok) {
  __shared__ WarpScratch s_ws[kPcaWarps];
  WarpScratch& ws = s_ws[threadIdx.x >> 5];
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int n_tgt = m.n_raw > 0 ? counts->n_target : 0;
  for (int i = blockIdx.x * wpb + (threadIdx.x >> 5); i < n; i += gridDim.x * wpb) {
    const float xf = pts6[6 * (size_t)i], yf = pts6[6 * (size_t)i + 1], zf = pts6[6 * (size_t)i + 2];
    const double qx = (double)xf, qy = (double)yf, qz = (double)zf;
    const double xnx = (double)pts6[6 * (size_t)i + 3], xny = (double)pts6[6 * (size_t)i + 4], xnz = (double)pts6[6 * (size_t)i + 5];
    TopK tk;
    tk.d2 = CUDART_INF; tk.idx = -1; tk.pos = -1;
    SearchStats ss;
    ss.on = false;
    if (n_tgt > 0) knn_topk<LEVELS>(m, xf, yf, zf, CUDART_INF_F, true, P.r2, P.k, true, ws, tk, ss, lane);
    const bool has = (lane < P.k) && (tk.d2 < CUDART_INF);
    double pnx = 0.0, pny = 0.0, pnz = 0.0, ddx = 0.0, ddy = 0.0, ddz = 0.0;
    bool keep = false;
    if (has) {
      const float4 pp = __ldg(&m.pts[tk.pos]);
      if (PCA) { pnx = m.nrm_pca[3 * (size_t)tk.pos]; pny = m.nrm_pca[3 * (size_t)tk.pos + 1]; pnz = m.nrm_pca[3 * (size_t)tk.pos + 2]; }
      else { const float4 nn = __ldg(&m.nrm[tk.pos]); pnx = (double)nn.x; pny = (double)nn.y; pnz = (double)nn.z; }
      ddx = __dsub_rn(qx, (double)pp.x); ddy = __dsub_rn(qy, (double)pp.y); ddz = __dsub_rn(qz, (double)pp.z);
      keep = finite3d(pnx, pny, pnz);
      if (keep && P.angle_constraint) keep = !angle_exceeds(xnx, xny, xnz, pnx, pny, pnz, P);
    }
    const int cnt = __popc(__ballot_sync(PLO_FULL_MASK, keep));
    double h = CUDART_NAN;
    if (cnt >= 3) {
      const double cinv = -9.0 / __shfl_sync(PLO_FULL_MASK, tk.d2, cnt - 1);
      double w = 0.0, pr = 0.0;
      if (keep) {
        w = exp(tk.d2 * cinv);
        pr = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(w, ddx), pnx), __dmul_rn(__dmul_rn(w, ddy), pny)), __dmul_rn(__dmul_rn(w, ddz), pnz));
      }
      const double wsum = warp_sum(w), psum = warp_sum(pr);
      h = psum / (wsum + 1e-5);
    }
    if (lane == 0) { height[i] = h; ok[i] = cnt >= 3 ? 1 : 0; }
    __syncwarp();
  }
}

// Eigen::Vector3d IMLSICPMatcher::ComputeNormal(std::vector<Eigen::Vector3d>&), src/imls_icp.cpp:753-794: mean,
// population covariance, eigenvector of the smallest eigenvalue, normalised, no sign disambiguation.  One thread,
// sums in the reference's order.
__global__ void k_compute_normal(const double* __restrict__ p, int n, double* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double mx = 0.0, my = 0.0, mz = 0.0;
  for (int i = 0; i < n; ++i) { mx += p[3 * i]; my += p[3 * i + 1]; mz += p[3 * i + 2]; }   // :758-763
  const double nd = (double)n;
  mx /= nd; my /= nd; mz /= nd;
  double c00 = 0, c01 = 0, c02 = 0, c11 = 0, c12 = 0, c22 = 0;
  for (int i = 0; i < n; ++i) {   // :766-771
    const double dx = p[3 * i] - mx, dy = p[3 * i + 1] - my, dz = p[3 * i + 2] - mz;
    c00 += dx * dx; c01 += dx * dy; c02 += dx * dz; c11 += dy * dy; c12 += dy * dz; c22 += dz * dz;
  }
  double v[3];
  smallest_eigvec3(c00 / nd, c01 / nd, c02 / nd, c11 / nd, c12 / nd, c22 / nd, v);   // :776-778
  const double nn = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
  if (nn > 0.0) { v[0] /= nn; v[1] /= nn; v[2] /= nn; }   // :791
  out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
}

}  // namespace

namespace {
// device buffers that must read zero when first used (tickets / counters are left at zero by every projection)
int reserve_zeroed(plo_ctx* c, DevBuf& b, size_t bytes) {
  const void* before = b.p;
  PLO_CUDA(c, b.reserve(bytes));
  if (b.p != before) PLO_CUDA(c, cudaMemsetAsync(b.p, 0, b.cap, c->stream));
  return PLO_OK;
}
}  // namespace

int plo_reserve_query_buffers(plo_ctx* c, bool hooks) {
  const size_t m = (size_t)(c->m_raw > 0 ? c->m_raw : 1);
  PLO_CUDA(c, c->q_x.reserve(sizeof(float4) * m));
  PLO_CUDA(c, c->q_y.reserve(sizeof(float4) * m));
  PLO_CUDA(c, c->q_n.reserve(sizeof(float4) * m));
  PLO_CUDA(c, c->q_status.reserve(sizeof(int) * m));
  PLO_CUDA(c, c->q_kd2.reserve(sizeof(double) * m));
  PLO_CUDA(c, c->q_tile_pts.reserve(sizeof(float4) * kTileSlots * m));
  PLO_CUDA(c, c->q_tile_meta.reserve(sizeof(float4) * m));
  PLO_TRY(reserve_zeroed(c, c->sync_counters, sizeof(int) * CNT_N));
  {
    const void* before = c->miss_list.p;
    PLO_CUDA(c, c->miss_list.reserve(sizeof(int) * m));
    if (c->miss_list.p != before) PLO_CUDA(c, cudaMemsetAsync(c->miss_list.p, 0xff, c->miss_list.cap, c->stream));   // -1 = empty slot
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
struct ProjectLaunch {
  MapView mv;
  const float4 *sp, *sn;
  const DevCounts* dc;
  DevState* st;
  ProjectOut out;
  LoopSync sync;
  int blocks, chunk, group;
};

template <bool PCA, bool HOOKS>
void launch_project_levels(plo_ctx* c, const ProjectLaunch& a) {
  const int T = kWarpsPerBlock * 32;
  switch (c->n_levels) {   // an empty map (n_levels == 0) never walks the tree: any instantiation does
    case 0:
    case 1: k_project<PCA, 1, HOOKS><<<a.blocks, T, 0, c->stream>>>(a.mv, a.sp, a.sn, a.dc, a.st, c->dprm, a.out, a.sync, a.chunk, a.group); break;
    case 2: k_project<PCA, 2, HOOKS><<<a.blocks, T, 0, c->stream>>>(a.mv, a.sp, a.sn, a.dc, a.st, c->dprm, a.out, a.sync, a.chunk, a.group); break;
    case 3: k_project<PCA, 3, HOOKS><<<a.blocks, T, 0, c->stream>>>(a.mv, a.sp, a.sn, a.dc, a.st, c->dprm, a.out, a.sync, a.chunk, a.group); break;
    case 4: k_project<PCA, 4, HOOKS><<<a.blocks, T, 0, c->stream>>>(a.mv, a.sp, a.sn, a.dc, a.st, c->dprm, a.out, a.sync, a.chunk, a.group); break;
    case 5: k_project<PCA, 5, HOOKS><<<a.blocks, T, 0, c->stream>>>(a.mv, a.sp, a.sn, a.dc, a.st, c->dprm, a.out, a.sync, a.chunk, a.group); break;
    default: k_project<PCA, 6, HOOKS><<<a.blocks, T, 0, c->stream>>>(a.mv, a.sp, a.sn, a.dc, a.st, c->dprm, a.out, a.sync, a.chunk, a.group); break;
  }
}
}  // namespace

namespace {
void fill_project_launch(plo_ctx* c, ProjectLaunch& a) {
  a.mv = c->map_view();
  a.sp = c->s_p.as<float4>();
  a.sn = c->s_n.as<float4>();
  a.dc = c->counts.as<DevCounts>();
  a.st = c->state.as<DevState>();
  ProjectOut& out = a.out;
  out.qx = c->q_x.as<float4>(); out.qy = c->q_y.as<float4>(); out.qn = c->q_n.as<float4>();
  out.status = c->q_status.as<int>();
  out.kd2f = c->q_kd2.as<float>();
  out.tile_pts = c->q_tile_pts.as<float4>(); out.tile_meta = c->q_tile_meta.as<float4>();
  out.height = c->q_height.as<double>(); out.nn1_idx = c->q_nn1_idx.as<int>(); out.nn1_d2 = c->q_nn1_d2.as<double>();
  out.nn_idx = c->q_nn_idx.as<int>(); out.nn_d2 = c->q_nn_d2.as<double>();
  out.search_stats = c->q_stats.as<int>();
  LoopSync& L = a.sync;
  L.counters = c->sync_counters.as<int>();
  L.miss_list = c->miss_list.as<int>();
  // persistent grids (PLO_MINB blocks per SM); chunks of consecutive source points are fetched through
  // an atomic counter.  The chunk length comes from the device-side loop state (see finish_iteration)
  // unless the cloud is too small to fill the GPU (then 1) or the tuning knob overrides it.
  const int64_t slots = (int64_t)plo_grid(c, PLO_MINB) * kWarpsPerBlock;
  // the device-side policy (8 / 4 / 1) is capped so that every warp of the grid gets at least ~2.5 chunks: chunks of
  // consecutive points buy the carry bound (no greedy descent per query), too few of them unbalance the grid
  int cap = 8;
  while (cap > 1 && c->m_raw * 2 < (int64_t)cap * 5 * slots) cap >>= 1;
  a.chunk = -cap;
  if (c->tune_chunk >= 0) a.chunk = c->tune_chunk;   // tuning knob: > 0 forces a length, 0 = device-side policy uncapped
  // settled path: a warp takes `group` consecutive queries at a time (their per-query scalars one per lane); small
  // clouds get small groups so that every warp of the grid has work (a 2 000-point source in groups of 32 would keep
  // 63 warps busy, each answering 32 queries one after the other)
  a.group = kGroup;
  while (a.group > 1 && c->m_raw < (int64_t)a.group * 2 * slots) a.group >>= 1;
  if (c->tune_group > 0) a.group = std::min(c->tune_group, kGroup);
  const int64_t warps = std::max<int64_t>(c->m_raw, 1);
  a.blocks = (int)std::max<int64_t>(1, std::min<int64_t>((warps + kWarpsPerBlock - 1) / kWarpsPerBlock, (int64_t)plo_grid(c, PLO_MINB)));
}
}  // namespace

// One projection = one k_project launch.
int plo_launch_project(plo_ctx* c, bool hooks) {
  if (c->m_raw == 0) return PLO_OK;
  ProjectLaunch a;
  fill_project_launch(c, a);
  if (c->dprm.use_pca_normals) {
    if (hooks) launch_project_levels<true, true>(c, a);
    else launch_project_levels<true, false>(c, a);
  } else {
    if (hooks) launch_project_levels<false, true>(c, a);
    else launch_project_levels<false, false>(c, a);
  }
  c->prev_valid = true;   // later projections of the same clouds may use this one's k-th distances
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  return PLO_OK;
}

int plo_launch_imls_height(plo_ctx* c, const float* d_pts6, int n, double* d_height, int* d_ok) {
  if (n <= 0) return PLO_OK;
  const int grid = std::max(1, std::min((n + kPcaWarps - 1) / kPcaWarps, plo_grid(c, 4)));
  const MapView mv = c->map_view();
  const DevCounts* dc = c->counts.as<DevCounts>();
#define PLO_IMLS_CASE(L)                                                                                                       \
  case L:                                                                                                                      \
    if (c->dprm.use_pca_normals) k_imls_height<true, L><<<grid, kPcaWarps * 32, 0, c->stream>>>(mv, c->dprm, dc, d_pts6, n, d_height, d_ok); \
    else k_imls_height<false, L><<<grid, kPcaWarps * 32, 0, c->stream>>>(mv, c->dprm, dc, d_pts6, n, d_height, d_ok);                     \
    break;
  switch (std::max(c->n_levels, 1)) {
    PLO_IMLS_CASE(1) PLO_IMLS_CASE(2) PLO_IMLS_CASE(3) PLO_IMLS_CASE(4) PLO_IMLS_CASE(5)
    default:
      if (c->dprm.use_pca_normals) k_imls_height<true, 6><<<grid, kPcaWarps * 32, 0, c->stream>>>(mv, c->dprm, dc, d_pts6, n, d_height, d_ok);
      else k_imls_height<false, 6><<<grid, kPcaWarps * 32, 0, c->stream>>>(mv, c->dprm, dc, d_pts6, n, d_height, d_ok);
  }
#undef PLO_IMLS_CASE
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  return PLO_OK;
}

int plo_launch_compute_normal(plo_ctx* c, const double* d_pts3, int n, double* d_out) {
  k_compute_normal<<<1, 32, 0, c->stream>>>(d_pts3, n, d_out);
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  return PLO_OK;
}

int plo_loop_blocks(const plo_ctx* c) {
  const int64_t warps = std::max<int64_t>(c->m_raw, 1);
  return (int)std::max<int64_t>(1, std::min<int64_t>((warps + kWarpsPerBlock - 1) / kWarpsPerBlock, (int64_t)plo_grid(c, PLO_MINB)));
}

namespace {
template <bool PCA, int LEVELS>
cudaError_t launch_loop_one(plo_ctx* c, ProjectLaunch& a, double* partials, unsigned* bar) {
  auto kern = k_register_loop<PCA, LEVELS>;
  void* args[] = {&a.mv, &a.sp, &a.sn, &a.dc, &a.st, &c->dprm, &a.out, &a.sync, &a.chunk, &a.group, &partials, &bar};
  return cudaLaunchCooperativeKernel((const void*)kern, dim3(a.blocks), dim3(kWarpsPerBlock * 32), args, 0, c->stream);
}
template <bool PCA>
cudaError_t launch_loop_levels(plo_ctx* c, ProjectLaunch& a, double* partials, unsigned* bar) {
  switch (c->n_levels) {
    case 0:
    case 1: return launch_loop_one<PCA, 1>(c, a, partials, bar);
    case 2: return launch_loop_one<PCA, 2>(c, a, partials, bar);
    case 3: return launch_loop_one<PCA, 3>(c, a, partials, bar);
    case 4: return launch_loop_one<PCA, 4>(c, a, partials, bar);
    case 5: return launch_loop_one<PCA, 5>(c, a, partials, bar);
    default: return launch_loop_one<PCA, 6>(c, a, partials, bar);
  }
}
}  // namespace

// The whole weighted-LS ICP loop of one registration as ONE cooperative launch (k_register_loop).
int plo_launch_register_loop(plo_ctx* c) {
  ProjectLaunch a;
  fill_project_launch(c, a);
  PLO_CUDA(c, c->partials.reserve(sizeof(double) * PLO_NSUM * (size_t)plo_grid(c, PLO_MINB) + 8 * 8 * 16));   // (+ the PLO_LOOP_TIMING stamps)
  if (!c->loop_barrier.p) {
    PLO_CUDA(c, c->loop_barrier.reserve(sizeof(unsigned) * 2));
  }
  const cudaError_t e = c->dprm.use_pca_normals ? launch_loop_levels<true>(c, a, c->partials.as<double>(), c->loop_barrier.as<unsigned>())
                                               : launch_loop_levels<false>(c, a, c->partials.as<double>(), c->loop_barrier.as<unsigned>());
  if (e != cudaSuccess) return plo_fail(c, PLO_ERR_CUDA, std::string("k_register_loop: ") + cudaGetErrorString(e));
  c->prev_valid = true;
  c->launches++;
  return PLO_OK;
}

#ifdef PLO_LOOP_TIMING
extern "C" __attribute__((visibility("default"))) int plo_debug_solve_stamps(unsigned long long* out8) {
  return (int)cudaMemcpyFromSymbol(out8, g_solve_stamp, sizeof(unsigned long long) * 8);
}
#endif
