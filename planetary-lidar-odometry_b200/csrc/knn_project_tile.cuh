// knn_project_tile.cuh — the projection kernel for large sources: one LANE per query, one warp per
// tile of 32 spatially adjacent queries (the source is ordered along its own Hilbert curve at upload,
// plo_upload_source).  Included by knn_project.cu inside its anonymous namespace; same semantics and
// outputs as k_project (transform of src/laser_odometry.cpp:527-549, ProjSourcePtToSurface
// src/imls_icp.cpp:496-745 with ImplicitMLSFunction :301-483), different mapping to the machine:
//
//   * the warp walks the wide BVH ONCE for its 32 queries.  Internal nodes: lane c tests child c
//     against the tile's bounding box inflated by the largest per-lane search radius (box-to-box
//     lower bound, directed rounding).  Leaves: every lane tests the leaf box against its own ball;
//     a leaf that no ball reaches is skipped.
//   * a visited leaf is loaded once (one coalesced 512-byte read, prefetched while the previous leaf is
//     scanned), staged in shared memory, and every lane scans its 32 points against its own query
//     (broadcast reads, no cross-lane traffic): conservative fp32 lower bounds first, as one 32-bit
//     pass mask; then, for the few survivors, the exact fp64 distance
//     (d2 = ((dx*dx + dy*dy) + dz*dz), no FMA) and libnabo's acceptance rule.
//   * each lane keeps its k best in an UNSORTED shared-memory column with the current maximum tracked
//     in registers: a survivor either fills a free slot or replaces the maximum (then one rescan of the
//     column); the lane's search radius follows the maximum.  One insertion sort per lane at the end.
//   * bounds: the temporal triangle-inequality bound of k_project; a lane without a useful one first
//     descends greedily to its own nearest leaf and takes the k-th distance among that leaf's points.
//   * the IMLS tail (normal gates, weights, height) runs per lane over its sorted list, sequentially
//     in list order (the summation order of the reference).
//
// Compared with the warp-per-query kernel a leaf fetched from L2 serves 32 queries instead of one and
// no instruction is spent on warp-wide ranking.
// Exactness argument unchanged: a point is skipped only when a LOWER bound of its distance exceeds an
// UPPER bound of the lane's k-th distance; order and ties are decided on exact fp64 values and indices.

constexpr int kTileWarps = 4;
constexpr size_t kTileScratchBytes = (sizeof(WarpScratch) + 127) / 128 * 128;   // 1-NN fallback (rare)

__host__ __device__ inline size_t tile_warp_bytes(int k) {
  return sizeof(float4) * 32 + kTileScratchBytes + (sizeof(double) + sizeof(int)) * 32 * (size_t)k;
}

__device__ __forceinline__ unsigned f2ord_u(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f_u(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// lower bound of the squared distance between two boxes (a: tile box, b: node box); +inf for an empty b
__device__ __forceinline__ float boxbox_lo2(float alx, float aly, float alz, float ahx, float ahy, float ahz, const float4 lo,
                                            const float4 hi) {
  const float ex = fmaxf(fmaxf(__fsub_rd(lo.x, ahx), __fsub_rd(alx, hi.x)), 0.f);
  const float ey = fmaxf(fmaxf(__fsub_rd(lo.y, ahy), __fsub_rd(aly, hi.y)), 0.f);
  const float ez = fmaxf(fmaxf(__fsub_rd(lo.z, ahz), __fsub_rd(alz, hi.z)), 0.f);
  return __fadd_rd(__fadd_rd(__fmul_rd(ex, ex), __fmul_rd(ey, ey)), __fmul_rd(ez, ez));
}

struct TileLane {
  float x, y, z;   // query (float32, as stored by the reference)
  float thr;       // float threshold for lower-bound tests; < 0 for an idle lane
  int n;           // list length
  int seed;        // leaf already consumed by the greedy bound (-1: none)
  double maxd2;    // full list: its maximum under the (d2, index) order ...
  int maxslot;     // ... and where it sits
};

struct TileCtx {
  const float4* pts;
  float4* stage;
  double* ld2;    // lane column (stride 32)
  int* lpos;      // lane column (stride 32)
  double r2;
  int k;
  int n_leaf, n_node;   // statistics (hooks): per tile
  int n_cand, n_repl;   // per lane: fp32 survivors, replacements of the maximum
};

__device__ __forceinline__ int tile_idx_of(const TileCtx& tc, int pos) { return __float_as_int(__ldg(&tc.pts[pos]).w); }

// maximum of the full column under the (d2, index) order; ties (equal d2) are rare and cost two loads
__device__ __forceinline__ void tile_rescan_max(const TileCtx& tc, TileLane& q) {
  double md = tc.ld2[0];
  int ms = 0;
  for (int j = 1; j < tc.k; ++j) {
    const double dj = tc.ld2[j * 32];
    bool gt = dj > md;
    if (dj == md) gt = tile_idx_of(tc, tc.lpos[j * 32]) > tile_idx_of(tc, tc.lpos[ms * 32]);
    if (gt) { md = dj; ms = j; }
  }
  q.maxd2 = md;
  q.maxslot = ms;
  q.thr = fminf(q.thr, bound_f(md));
}

// exact test of one fp32 survivor (ALLOW_SELF_MATCH search: d2 <= r2 is the whole acceptance rule)
__device__ __forceinline__ void tile_consider(TileCtx& tc, TileLane& q, const float4 p, int pos) {
  const double d2 = dist2_exact((double)q.x, (double)q.y, (double)q.z, p);
  if (!(d2 <= tc.r2)) return;
  if (q.n < tc.k) {
    tc.ld2[q.n * 32] = d2;
    tc.lpos[q.n * 32] = pos;
    q.n++;
    if (q.n == tc.k) tile_rescan_max(tc, q);
    return;
  }
  if (d2 > q.maxd2) return;
  if (d2 == q.maxd2 && __float_as_int(p.w) > tile_idx_of(tc, tc.lpos[q.maxslot * 32])) return;
  tc.ld2[q.maxslot * 32] = d2;
  tc.lpos[q.maxslot * 32] = pos;
  tc.n_repl++;
  tile_rescan_max(tc, q);
}

// all lanes scan the 32 points of a leaf (`pj`: this lane's point of it); a lane takes candidates iff `mine`
__device__ __forceinline__ void tile_scan_leaf(TileCtx& tc, TileLane& q, int leaf, const float4 pj, bool mine, int lane) {
  __syncwarp();
  tc.stage[lane] = pj;
  __syncwarp();
  tc.n_leaf++;
  const float thr = mine ? q.thr : -1.f;
  unsigned pm = 0u;
#pragma unroll
  for (int i = 0; i < PLO_LEAF; ++i) {
    const float4 p = tc.stage[i];
    pm |= (dist_lo2(q.x, q.y, q.z, p) <= thr ? 1u : 0u) << i;
  }
  tc.n_cand += __popc(pm);
  while (pm != 0u) {
    const int i = __ffs(pm) - 1;
    pm &= pm - 1u;
    const float4 p = tc.stage[i];
    if (dist_lo2(q.x, q.y, q.z, p) <= q.thr) tile_consider(tc, q, p, leaf * PLO_LEAF + i);   // the radius may have shrunk
  }
}

// in-place insertion sort of the lane's column by (d2, index)
__device__ __forceinline__ void tile_sort_list(const TileCtx& tc, int n) {
  for (int a = 1; a < n; ++a) {
    const double d = tc.ld2[a * 32];
    const int ps = tc.lpos[a * 32];
    int j = a;
    while (j > 0) {
      const double dj = tc.ld2[(j - 1) * 32];
      bool less = d < dj;
      if (d == dj) less = tile_idx_of(tc, ps) < tile_idx_of(tc, tc.lpos[(j - 1) * 32]);
      if (!less) break;
      tc.ld2[j * 32] = dj;
      tc.lpos[j * 32] = tc.lpos[(j - 1) * 32];
      --j;
    }
    tc.ld2[j * 32] = d;
    tc.lpos[j * 32] = ps;
  }
}

template <int LEVEL>
struct TileWalk {
  static __device__ __forceinline__ void run(const MapView& m, int node, TileCtx& tc, TileLane& q, float alx, float aly, float alz,
                                             float ahx, float ahy, float ahz, float& Rf, int lane) {
    const int child = node * PLO_FANOUT + lane;
    tc.n_node++;
    const float4 lo = __ldg(&m.lo[LEVEL - 1][child]), hi = __ldg(&m.hi[LEVEL - 1][child]);
    const float bd = boxbox_lo2(alx, aly, alz, ahx, ahy, ahz, lo, hi);
    unsigned mask = __ballot_sync(PLO_FULL_MASK, bd <= Rf);
    [[maybe_unused]] float4 pre = make_float4(0.f, 0.f, 0.f, 0.f);   // LEVEL 1: points of leaf `pre_c`, fetched ahead
    [[maybe_unused]] int pre_c = -1;
    if constexpr (LEVEL == 1) {
      if (mask != 0u) {
        pre_c = __ffs(mask) - 1;
        pre = __ldg(&m.pts[(node * PLO_FANOUT + pre_c) * PLO_LEAF + lane]);
      }
    }
    while (mask != 0u) {
      const int c = __ffs(mask) - 1;
      mask &= mask - 1u;
      if constexpr (LEVEL == 1) {
        const int leaf = node * PLO_FANOUT + c;
        const float4 cur = (pre_c == c) ? pre : __ldg(&m.pts[leaf * PLO_LEAF + lane]);
        if (mask != 0u) {   // next candidate leaf: its load overlaps this leaf's scan
          pre_c = __ffs(mask) - 1;
          pre = __ldg(&m.pts[(node * PLO_FANOUT + pre_c) * PLO_LEAF + lane]);
        }
        const float4 bl = make_float4(__shfl_sync(PLO_FULL_MASK, lo.x, c), __shfl_sync(PLO_FULL_MASK, lo.y, c),
                                      __shfl_sync(PLO_FULL_MASK, lo.z, c), 0.f);
        const float4 bh = make_float4(__shfl_sync(PLO_FULL_MASK, hi.x, c), __shfl_sync(PLO_FULL_MASK, hi.y, c),
                                      __shfl_sync(PLO_FULL_MASK, hi.z, c), 0.f);
        const bool mine = (box_lo2(q.x, q.y, q.z, bl, bh) <= q.thr) && (leaf != q.seed);
        if (!__any_sync(PLO_FULL_MASK, mine)) continue;
        tile_scan_leaf(tc, q, leaf, cur, mine, lane);
      } else {
        TileWalk<LEVEL - 1>::run(m, node * PLO_FANOUT + c, tc, q, alx, aly, alz, ahx, ahy, ahz, Rf, lane);
      }
      // radii only shrink: re-test the remaining children against the current largest one
      const float Rn = __uint_as_float(__reduce_max_sync(PLO_FULL_MASK, __float_as_uint(fmaxf(q.thr, 0.f))));
      if (Rn < Rf) {
        Rf = Rn;
        mask &= __ballot_sync(PLO_FULL_MASK, bd <= Rf);
      }
    }
  }
};

// per-lane greedy descent: the leaf whose ancestors' box centres are nearest to the query
template <int LEVELS>
__device__ __forceinline__ int tile_seed_leaf(const MapView& m, float qx, float qy, float qz) {
  int node = 0;
#pragma unroll
  for (int level = LEVELS - 1; level >= 0; --level) {
    float best = CUDART_INF_F;
    int bc = -1;
    for (int c = 0; c < PLO_FANOUT; ++c) {
      const float4 lo = __ldg(&m.lo[level][node * PLO_FANOUT + c]), hi = __ldg(&m.hi[level][node * PLO_FANOUT + c]);
      if (!(lo.x <= hi.x)) continue;   // empty box (padding)
      const float cx = qx - 0.5f * (lo.x + hi.x), cy = qy - 0.5f * (lo.y + hi.y), cz = qz - 0.5f * (lo.z + hi.z);
      const float sc = cx * cx + cy * cy + cz * cz;
      if (sc < best) { best = sc; bc = c; }
    }
    if (bc < 0) return -1;
    node = node * PLO_FANOUT + bc;
  }
  return node;
}

template <bool PCA, int LEVELS, bool HOOKS>
__global__ void __launch_bounds__(kTileWarps * 32, 4) k_project_tile(const __grid_constant__ MapView m,
                                                                     const float4* __restrict__ sp, const float4* __restrict__ sn,
                                                                     const int* __restrict__ s_order,
                                                                     const DevCounts* __restrict__ counts,
                                                                     const DevState* __restrict__ st, DevParams P, ProjectOut out,
                                                                     int* __restrict__ tile_counter) {
  if (st->done) return;
  extern __shared__ __align__(16) unsigned char s_raw[];
  __shared__ double T[12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int k = P.k;
  unsigned char* wbase = s_raw + (size_t)warp * tile_warp_bytes(k);
  TileCtx tc;
  tc.pts = m.pts;
  tc.stage = reinterpret_cast<float4*>(wbase);
  WarpScratch* ws = reinterpret_cast<WarpScratch*>(wbase + sizeof(float4) * 32);   // 1-NN fallback only
  double* ld2_all = reinterpret_cast<double*>(wbase + sizeof(float4) * 32 + kTileScratchBytes);
  int* lpos_all = reinterpret_cast<int*>(wbase + sizeof(float4) * 32 + kTileScratchBytes + sizeof(double) * 32 * (size_t)k);
  tc.ld2 = ld2_all + lane;
  tc.lpos = lpos_all + lane;
  tc.r2 = P.r2;
  tc.k = k;
  const int use_prev = st->use_prev;
  const int n_src = counts->n_source;
  const int n_tgt = m.n_raw > 0 ? counts->n_target : 0;
  if (threadIdx.x < 12) T[threadIdx.x] = st->rPose[threadIdx.x];
  __syncthreads();

  while (true) {
    int t0 = 0;
    if (lane == 0) t0 = atomicAdd(tile_counter, 1) * 32;
    t0 = __shfl_sync(PLO_FULL_MASK, t0, 0);
    if (t0 >= n_src) break;
    const int qi = (t0 + lane < n_src) ? __ldg(&s_order[t0 + lane]) : -1;
    const bool active = qi >= 0;
    TileLane q;
    q.x = q.y = q.z = 0.f;
    q.thr = -1.f;
    q.n = 0;
    q.seed = -1;
    q.maxd2 = CUDART_INF;
    q.maxslot = 0;
    tc.n_leaf = tc.n_node = tc.n_cand = tc.n_repl = 0;
    float nxf = 0.f, nyf = 0.f, nzf = 0.f;
    bool refine = false;
    if (active) {
      const float4 p = __ldg(&sp[qi]);
      const float4 nf = __ldg(&sn[qi]);
      const double px = (double)p.x, py = (double)p.y, pz = (double)p.z;
      // p' = rPose * [p;1] in double, stored as float32 (src/laser_odometry.cpp:530-539)
      q.x = __double2float_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[0], px), __dmul_rn(T[1], py)), __dmul_rn(T[2], pz)), T[3]));
      q.y = __double2float_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[4], px), __dmul_rn(T[5], py)), __dmul_rn(T[6], pz)), T[7]));
      q.z = __double2float_rn(__dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[8], px), __dmul_rn(T[9], py)), __dmul_rn(T[10], pz)), T[11]));
      nxf = nf.x; nyf = nf.y; nzf = nf.z;
      if (P.transform_normal) {   // :541-548
        const double a = (double)nf.x, b = (double)nf.y, cc = (double)nf.z;
        nxf = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[0], a), __dmul_rn(T[1], b)), __dmul_rn(T[2], cc)));
        nyf = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[4], a), __dmul_rn(T[5], b)), __dmul_rn(T[6], cc)));
        nzf = __double2float_rn(__dadd_rn(__dadd_rn(__dmul_rn(T[8], a), __dmul_rn(T[9], b)), __dmul_rn(T[10], cc)));
      }
    }
    const bool searching = active && n_tgt > 0 && isfinite(q.x) && isfinite(q.y) && isfinite(q.z);
    if (searching) {
      // temporal bound: the k points nearest to this query's previous position are all within
      // sqrt(kd2_prev) + |x - x_prev| of x (float, rounded up; see tri_bound)
      float Df0 = CUDART_INF_F, ref_kf = CUDART_INF_F;
      if (use_prev) {
        const float kprev = out.kd2f[qi];
        if (kprev < CUDART_INF_F) {
          const float4 xp = out.qx[qi];
          Df0 = tri_bound(kprev, q.x, q.y, q.z, xp.x, xp.y, xp.z);
          ref_kf = kprev;
        }
      }
      if (!(Df0 == Df0)) Df0 = CUDART_INF_F;
      refine = !(Df0 < CUDART_INF_F) || !(Df0 <= 6.25f * ref_kf);
      q.thr = fminf(Df0, bound_f(P.r2));
    }

    if (__any_sync(PLO_FULL_MASK, searching)) {
      // ---- greedy bound for the lanes without a tight one: k-th distance within the own nearest leaf ----
      if (__any_sync(PLO_FULL_MASK, refine)) {
        if (refine) q.seed = tile_seed_leaf<LEVELS>(m, q.x, q.y, q.z);
        unsigned todo = __ballot_sync(PLO_FULL_MASK, q.seed >= 0);
        while (todo != 0u) {
          const int leaf = __shfl_sync(PLO_FULL_MASK, q.seed, __ffs(todo) - 1);
          const bool mine = q.seed == leaf;
          tile_scan_leaf(tc, q, leaf, __ldg(&m.pts[leaf * PLO_LEAF + lane]), mine, lane);
          todo &= ~__ballot_sync(PLO_FULL_MASK, mine);
        }
      }
      // ---- tile box and the largest radius ----
      const unsigned ux = f2ord_u(q.x), uy = f2ord_u(q.y), uz = f2ord_u(q.z);
      const float alx = ord2f_u(__reduce_min_sync(PLO_FULL_MASK, searching ? ux : 0xffffffffu));
      const float aly = ord2f_u(__reduce_min_sync(PLO_FULL_MASK, searching ? uy : 0xffffffffu));
      const float alz = ord2f_u(__reduce_min_sync(PLO_FULL_MASK, searching ? uz : 0xffffffffu));
      const float ahx = ord2f_u(__reduce_max_sync(PLO_FULL_MASK, searching ? ux : 0u));
      const float ahy = ord2f_u(__reduce_max_sync(PLO_FULL_MASK, searching ? uy : 0u));
      const float ahz = ord2f_u(__reduce_max_sync(PLO_FULL_MASK, searching ? uz : 0u));
      float Rf = __uint_as_float(__reduce_max_sync(PLO_FULL_MASK, __float_as_uint(fmaxf(q.thr, 0.f))));
      TileWalk<LEVELS>::run(m, 0, tc, q, alx, aly, alz, ahx, ahy, ahz, Rf, lane);
      tile_sort_list(tc, q.n);
    }
    __syncwarp();

    // ---- 1-NN without self match (src/imls_icp.cpp:601-609): first list entry with d2 > DBL_EPSILON ----
    int j1 = -1;
    for (int j = 0; j < q.n; ++j)
      if (tc.ld2[j * 32] > DBL_EPSILON) { j1 = j; break; }
    int i1 = -1, pos1 = -1;
    double d1 = CUDART_INF;
    if (j1 >= 0) {
      d1 = tc.ld2[j1 * 32];
      pos1 = tc.lpos[j1 * 32];
      i1 = __float_as_int(__ldg(&m.pts[pos1]).w);
    }
    {
      // the list is full of points coincident with the query: search again, k = 1, no self match (warp-wide, rare)
      unsigned fb = __ballot_sync(PLO_FULL_MASK, searching && j1 < 0 && q.n == k);
      while (fb != 0u) {
        const int l = __ffs(fb) - 1;
        fb &= fb - 1u;
        const float fx = __shfl_sync(PLO_FULL_MASK, q.x, l), fy = __shfl_sync(PLO_FULL_MASK, q.y, l), fz = __shfl_sync(PLO_FULL_MASK, q.z, l);
        knn1_noself<LEVELS>(m, fx, fy, fz, P.r2, ws);
        if (lane == l && ws->od2[0] < CUDART_INF) { d1 = ws->od2[0]; i1 = ws->oidx[0]; pos1 = ws->opos[0]; }
        __syncwarp();
      }
    }

    if (active) {
      const double qx = (double)q.x, qy = (double)q.y, qz = (double)q.z;           // imls_icp.cpp:556
      const double xnx = (double)nxf, xny = (double)nyf, xnz = (double)nzf;        // :557
      const float kd2f_now = (q.n == k) ? __double2float_ru(tc.ld2[(k - 1) * 32]) : CUDART_INF_F;
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
      if (status == PLO_PT_OK) {
        // ImplicitMLSFunction (:376-480): neighbours with a finite normal inside the angle cone
        unsigned keepmask = 0u;
        int cnt = 0;
        for (int j = 0; j < q.n; ++j) {
          const int pos = tc.lpos[j * 32];
          double pnx, pny, pnz;
          if (PCA) { pnx = m.nrm_pca[3 * (size_t)pos]; pny = m.nrm_pca[3 * (size_t)pos + 1]; pnz = m.nrm_pca[3 * (size_t)pos + 2]; }
          else { const float4 nn = __ldg(&m.nrm[pos]); pnx = (double)nn.x; pny = (double)nn.y; pnz = (double)nn.z; }
          bool keep = finite3d(pnx, pny, pnz);                                           // :436-440
          if (keep && P.angle_constraint) keep = !angle_exceeds(xnx, xny, xnz, pnx, pny, pnz, P);   // :442-451
          if (keep) { keepmask |= 1u << j; ++cnt; }
        }
        if (cnt < 3) status = PLO_PT_MLS_FAIL;               // :463-466, :696-701
        else {
          // :468 — h_max = sqrt(d2[cnt-1]) / 3 indexes the UNFILTERED sorted list with the filtered count
          const double cinv = -9.0 / tc.ld2[(cnt - 1) * 32];
          double wsum = 0.0, psum = 0.0;
          for (int j = 0; j < q.n; ++j) {
            if (!((keepmask >> j) & 1u)) continue;
            const int pos = tc.lpos[j * 32];
            const float4 pp = __ldg(&m.pts[pos]);
            double pnx, pny, pnz;
            if (PCA) { pnx = m.nrm_pca[3 * (size_t)pos]; pny = m.nrm_pca[3 * (size_t)pos + 1]; pnz = m.nrm_pca[3 * (size_t)pos + 2]; }
            else { const float4 nn = __ldg(&m.nrm[pos]); pnx = (double)nn.x; pny = (double)nn.y; pnz = (double)nn.z; }
            const double ddx = __dsub_rn(qx, (double)pp.x), ddy = __dsub_rn(qy, (double)pp.y), ddz = __dsub_rn(qz, (double)pp.z);
            const double w = exp(tc.ld2[j * 32] * cinv);   // :474-475
            const double pr = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(w, ddx), pnx), __dmul_rn(__dmul_rn(w, ddy), pny)), __dmul_rn(__dmul_rn(w, ddz), pnz));   // :476
            wsum = __dadd_rn(wsum, w);
            psum = __dadd_rn(psum, pr);
          }
          height = psum / (wsum + 1e-5);                     // :480
          if (!isfinite(height)) status = PLO_PT_NAN_INF_HEIGHT;   // :703-717
        }
      }
      float4 ox = make_float4(q.x, q.y, q.z, __int_as_float(status));
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
      if constexpr (HOOKS) {
        for (int j = 0; j < k; ++j) {
          const bool has = j < q.n;
          out.nn_idx[(size_t)qi * k + j] = has ? __float_as_int(__ldg(&m.pts[tc.lpos[j * 32]]).w) : -1;
          out.nn_d2[(size_t)qi * k + j] = has ? tc.ld2[j * 32] : CUDART_INF;
        }
        out.height[qi] = height;
        out.nn1_idx[qi] = i1;
        out.nn1_d2[qi] = d1;
        out.search_stats[3 * (size_t)qi] = tc.n_leaf;
        out.search_stats[3 * (size_t)qi + 1] = tc.n_node;
        out.search_stats[3 * (size_t)qi + 2] = tc.n_cand + 100000 * tc.n_repl;
      }
    }
    __syncwarp();
  }
}
