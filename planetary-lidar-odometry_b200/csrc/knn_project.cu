// knn_project.cu — the matcher's hot kernel: per-iteration source transform
// (src/laser_odometry.cpp:527-549) + IMLSICPMatcher::ProjSourcePtToSurface
// (src/imls_icp.cpp:496-745) with ImplicitMLSFunction (:301-483) fused in, and the
// per-map-point PCA normal pass (ComputeNormal, :753-794, call sites :411-433,:647-669).
//
// One warp per query.  The warp walks the Morton-sorted wide BVH of index_build.cu with
// warp-uniform control flow: at every internal node the 32 lanes test the 32 child boxes
// (one coalesced float4 pair per lane), children are visited nearest-first
// (REDUX min over the box distances); a leaf is one coalesced 512-byte load of 32
// points, one exact fp64 distance per lane.  The k best neighbours live one per lane
// in registers, ordered by (d2, index); insertion is a ballot/popc rank plus one
// shuffle-up.  libnabo's knn semantics are reproduced exactly:
//   accept iff d2 <= r*r and (allow_self || d2 > DBL_EPSILON), k best ascending,
//   d2 = ((dx*dx + dy*dy) + dz*dz) in double without FMA, ties by index (D3).
// A node is pruned iff boxd2 > min(r2, current k-th d2); boxd2 is computed with the same
// operation order as d2, so by monotonicity of IEEE rounding it never exceeds the d2 of a
// point inside the box and the strict comparison keeps the result exact even for ties.
//
// The 1-NN of :601-609 (no self match) is the first list entry with d2 > DBL_EPSILON;
// only if the list is full of coincident points is a second (k=1) search needed.
//
// Algorithmic bytes per source point per iteration: 24 B query + k * 24 B
// neighbours (+ 24 B pair written) = 504 / 528 B at k = 20.  Roofline: HBM (in practice
// L2: a 1 M-point map is 32 MB and stays L2-resident).
#include <float.h>
#include <math_constants.h>

#include "plo_internal.cuh"

namespace {

struct Query {
  double x, y, z;
};

// neighbour list: lane j (< k) holds the j-th best entry
struct TopK {
  double d2;
  int idx;   // stripped-cloud index (tie-break key, reported to the caller)
  int pos;   // position in the sorted arrays (for gathers)
};

struct Search {
  double r2;
  double kd2;     // (d2, idx) of the current k-th entry, (+inf, INT_MAX) while not full
  int kidx;
  int k;
  unsigned kmask;
  bool allow_self;
  int n_leaf, n_node, n_ins;   // traversal statistics (reported through the hooks)
};

__device__ __forceinline__ double dist2(const Query& q, float px, float py, float pz) {
  const double dx = __dsub_rn(q.x, (double)px), dy = __dsub_rn(q.y, (double)py), dz = __dsub_rn(q.z, (double)pz);
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

__device__ __forceinline__ double box_dist2(const Query& q, const float4 lo, const float4 hi) {
  const double ex = fmax(fmax(__dsub_rn((double)lo.x, q.x), __dsub_rn(q.x, (double)hi.x)), 0.0);
  const double ey = fmax(fmax(__dsub_rn((double)lo.y, q.y), __dsub_rn(q.y, (double)hi.y)), 0.0);
  const double ez = fmax(fmax(__dsub_rn((double)lo.z, q.z), __dsub_rn(q.z, (double)hi.z)), 0.0);
  return __dadd_rn(__dadd_rn(__dmul_rn(ex, ex), __dmul_rn(ey, ey)), __dmul_rn(ez, ez));
}

__device__ __forceinline__ bool better(double d2, int idx, double kd2, int kidx) {
  return d2 < kd2 || (d2 == kd2 && idx < kidx);
}

__device__ __forceinline__ void visit_leaf(const MapView& m, int leaf, const Query& q, Search& s, TopK& tk, int lane) {
  const float4 p = __ldg(&m.pts[leaf * PLO_LEAF + lane]);
  s.n_leaf++;
  const double d2 = dist2(q, p.x, p.y, p.z);
  const int cidx = __float_as_int(p.w);
  const bool pass = (d2 <= s.r2) && (s.allow_self || d2 > DBL_EPSILON) && better(d2, cidx, s.kd2, s.kidx);
  unsigned cand = __ballot_sync(PLO_FULL_MASK, pass);
  while (cand) {
    const int src = __ffs(cand) - 1;
    cand &= cand - 1;
    const double cd2 = __shfl_sync(PLO_FULL_MASK, d2, src);
    const int ci = __shfl_sync(PLO_FULL_MASK, cidx, src);
    if (!better(cd2, ci, s.kd2, s.kidx)) continue;   // warp-uniform
    s.n_ins++;
    const bool less = (tk.d2 < cd2) || (tk.d2 == cd2 && tk.idx < ci);
    const int at = __popc(__ballot_sync(PLO_FULL_MASK, less) & s.kmask);
    const double ud2 = __shfl_up_sync(PLO_FULL_MASK, tk.d2, 1);
    const int uidx = __shfl_up_sync(PLO_FULL_MASK, tk.idx, 1);
    const int upos = __shfl_up_sync(PLO_FULL_MASK, tk.pos, 1);
    if (lane == at) { tk.d2 = cd2; tk.idx = ci; tk.pos = leaf * PLO_LEAF + src; }
    else if (lane > at) { tk.d2 = ud2; tk.idx = uidx; tk.pos = upos; }
    s.kd2 = __shfl_sync(PLO_FULL_MASK, tk.d2, s.k - 1);
    s.kidx = __shfl_sync(PLO_FULL_MASK, tk.idx, s.k - 1);
  }
}

template <int LEVEL>
struct Visit {
  // `node` is a node of level LEVEL (or the virtual root); its children live in level LEVEL-1
  static __device__ __forceinline__ void run(const MapView& m, int node, const Query& q, Search& s, TopK& tk, int lane) {
    const int child = node * PLO_FANOUT + lane;
    s.n_node++;
    const double bd = box_dist2(q, __ldg(&m.lo[LEVEL - 1][child]), __ldg(&m.hi[LEVEL - 1][child]));
    // positive floats order like their bit patterns; rounding down keeps the order weakly
    const unsigned key = __float_as_uint(__double2float_rd(bd));
    unsigned pending = PLO_FULL_MASK;
    while (true) {
      const double bound = fmin(s.kd2, s.r2);
      const bool ok = ((pending >> lane) & 1u) && (bd <= bound);
      const unsigned live = __ballot_sync(PLO_FULL_MASK, ok);
      if (live == 0u) break;
      const unsigned mn = __reduce_min_sync(PLO_FULL_MASK, ok ? key : 0xffffffffu);
      const int c = __ffs(__ballot_sync(PLO_FULL_MASK, ok && key == mn)) - 1;
      pending &= live;            // boxes that failed once can never pass later (bound only shrinks)
      pending &= ~(1u << c);
      if (LEVEL == 1) visit_leaf(m, node * PLO_FANOUT + c, q, s, tk, lane);
      else Visit<(LEVEL > 1 ? LEVEL - 1 : 1)>::run(m, node * PLO_FANOUT + c, q, s, tk, lane);
    }
  }
};

__device__ __forceinline__ void knn_search(const MapView& m, const Query& q, Search& s, TopK& tk, int lane) {
  tk.d2 = CUDART_INF;
  tk.idx = 0x7fffffff;
  tk.pos = -1;
  s.kd2 = CUDART_INF;
  s.kidx = 0x7fffffff;
  s.n_leaf = s.n_node = s.n_ins = 0;
  if (!(isfinite(q.x) && isfinite(q.y) && isfinite(q.z))) return;
  switch (m.n_levels) {   // warp-uniform
    case 1: Visit<1>::run(m, 0, q, s, tk, lane); break;
    case 2: Visit<2>::run(m, 0, q, s, tk, lane); break;
    case 3: Visit<3>::run(m, 0, q, s, tk, lane); break;
    case 4: Visit<4>::run(m, 0, q, s, tk, lane); break;
    case 5: Visit<5>::run(m, 0, q, s, tk, lane); break;
    case 6: Visit<6>::run(m, 0, q, s, tk, lane); break;
    default: break;
  }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(PLO_FULL_MASK, v, o);
  return v;
}

__device__ __forceinline__ bool finite3d(double a, double b, double c) { return isfinite(a) && isfinite(b) && isfinite(c); }

// `angle > thr` of src/imls_icp.cpp:444-451 / :683-692, NaN => false (point kept).
// Decided on the cosine when it is clearly away from the threshold, by acos otherwise.
__device__ __forceinline__ bool angle_exceeds(double ax, double ay, double az, double bx, double by, double bz,
                                              const DevParams& P) {
  const double dot = __dadd_rn(__dadd_rn(__dmul_rn(ax, bx), __dmul_rn(ay, by)), __dmul_rn(az, bz));
  const double na = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(ax, ax), __dmul_rn(ay, ay)), __dmul_rn(az, az)));
  const double nb = sqrt(__dadd_rn(__dadd_rn(__dmul_rn(bx, bx), __dmul_rn(by, by)), __dmul_rn(bz, bz)));
  const double c = dot / (na * nb);
  if (!(c == c)) return false;
  if (fabs(c - P.cos_thr) > 1e-9 && fabs(c) <= 1.0) return c < P.cos_thr;
  const double angle = acos(c) * 180.0 / 3.14159265358979323846;
  return angle > P.angle_thr;
}

struct ProjectOut {
  float4* qx;        // transformed source point (float32), w = bits of status
  float4* qy;        // projected point y (float32)
  float4* qn;        // normal of the 1-NN (float32)
  int* status;
  // hooks
  double* height;
  int* nn1_idx;
  double* nn1_d2;
  int* nn_idx;
  double* nn_d2;
  int* search_stats;   // [M*3] leaves, internal nodes, insertions of the k-NN search
};

template <bool PCA>
__global__ void __launch_bounds__(256) k_project(MapView m, const float4* __restrict__ sp, const float4* __restrict__ sn,
                                                 const DevCounts* __restrict__ counts, const DevState* __restrict__ st,
                                                 DevParams P, ProjectOut out, int hooks) {
  if (st->done) return;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const int n_src = counts->n_source;
  const int n_tgt = m.n_raw > 0 ? counts->n_target : 0;
  // rPose rows (src/laser_odometry.cpp:530-535)
  double T[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) T[i] = st->rPose[i];

  for (int qi = blockIdx.x * wpb + (threadIdx.x >> 5); qi < n_src; qi += gridDim.x * wpb) {
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
    Query q{(double)xf, (double)yf, (double)zf};           // imls_icp.cpp:556
    const double xnx = (double)nxf, xny = (double)nyf, xnz = (double)nzf;   // :557

    Search s;
    s.r2 = P.r2;
    s.k = P.k;
    s.kmask = (P.k >= 32) ? 0xffffffffu : ((1u << P.k) - 1u);
    s.allow_self = true;   // :372-375 ALLOW_SELF_MATCH
    TopK tk;
    if (n_tgt > 0) knn_search(m, q, s, tk, lane);
    else { tk.d2 = CUDART_INF; tk.idx = 0x7fffffff; tk.pos = -1; s.n_leaf = s.n_node = s.n_ins = 0; }
    const bool has = (lane < P.k) && (tk.d2 < CUDART_INF);

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
        Search s1;
        s1.r2 = P.r2; s1.k = 1; s1.kmask = 1u; s1.allow_self = false;
        TopK t1;
        knn_search(m, q, s1, t1, lane);
        const double dd = __shfl_sync(PLO_FULL_MASK, t1.d2, 0);
        if (dd < CUDART_INF) {
          d1 = dd;
          i1 = __shfl_sync(PLO_FULL_MASK, t1.idx, 0);
          pos1 = __shfl_sync(PLO_FULL_MASK, t1.pos, 0);
        }
      }
    }

    // ---- per-neighbour data for the IMLS sum (one neighbour per lane) ----
    double pnx = 0.0, pny = 0.0, pnz = 0.0, ddx = 0.0, ddy = 0.0, ddz = 0.0;
    bool keep = false;
    if (has) {
      const float4 pp = __ldg(&m.pts[tk.pos]);
      if (PCA) { pnx = m.nrm_pca[3 * (size_t)tk.pos]; pny = m.nrm_pca[3 * (size_t)tk.pos + 1]; pnz = m.nrm_pca[3 * (size_t)tk.pos + 2]; }
      else { const float4 nn = __ldg(&m.nrm[tk.pos]); pnx = (double)nn.x; pny = (double)nn.y; pnz = (double)nn.z; }
      ddx = __dsub_rn(q.x, (double)pp.x); ddy = __dsub_rn(q.y, (double)pp.y); ddz = __dsub_rn(q.z, (double)pp.z);
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
        // :468 — the bandwidth indexes the UNFILTERED sorted distance list with the filtered count
        const double hmax = sqrt(__shfl_sync(PLO_FULL_MASK, tk.d2, cnt - 1)) / 3.0;
        double w = 0.0, pr = 0.0;
        if (keep) {
          w = exp(-tk.d2 / hmax / hmax);                   // :474-475 (diff_norm == d2, same arithmetic)
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
        oy.x = __double2float_rn(__dsub_rn(q.x, __dmul_rn(height, n0x)));
        oy.y = __double2float_rn(__dsub_rn(q.y, __dmul_rn(height, n0y)));
        oy.z = __double2float_rn(__dsub_rn(q.z, __dmul_rn(height, n0z)));
        on.x = __double2float_rn(n0x); on.y = __double2float_rn(n0y); on.z = __double2float_rn(n0z);
      }
      out.qx[qi] = ox; out.qy[qi] = oy; out.qn[qi] = on;
      out.status[qi] = status;
    }
    if (hooks) {
      if (lane < P.k) {
        out.nn_idx[(size_t)qi * P.k + lane] = has ? tk.idx : -1;
        out.nn_d2[(size_t)qi * P.k + lane] = has ? tk.d2 : CUDART_INF;
      }
      if (lane == 0) {
        out.height[qi] = height;
        out.nn1_idx[qi] = i1;
        out.nn1_d2[qi] = d1;
        out.search_stats[3 * (size_t)qi] = s.n_leaf;
        out.search_stats[3 * (size_t)qi + 1] = s.n_node;
        out.search_stats[3 * (size_t)qi + 2] = s.n_ins;
      }
    }
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
__global__ void __launch_bounds__(256) k_pca_normals(MapView m, DevParams P, double* __restrict__ nrm_pca, int n_pad) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int pos = blockIdx.x * wpb + (threadIdx.x >> 5); pos < n_pad; pos += gridDim.x * wpb) {
    const float4 p = __ldg(&m.pts[pos]);
    double nx = CUDART_INF, ny = CUDART_INF, nz = CUDART_INF;   // :418-421
    if (isfinite(p.x)) {
      Query q{(double)p.x, (double)p.y, (double)p.z};
      Search s;
      s.r2 = P.r_normal2; s.k = P.k_normal; s.kmask = (P.k_normal >= 32) ? 0xffffffffu : ((1u << P.k_normal) - 1u);
      s.allow_self = false;
      TopK tk;
      knn_search(m, q, s, tk, lane);
      const bool has = (lane < P.k_normal) && (tk.d2 < CUDART_INF);
      const int cnt = __popc(__ballot_sync(PLO_FULL_MASK, has));
      if (cnt == P.k_normal) {
        double x = 0.0, y = 0.0, z = 0.0;
        if (has) { const float4 pp = __ldg(&m.pts[tk.pos]); x = (double)pp.x; y = (double)pp.y; z = (double)pp.z; }
        const double inv = 1.0 / (double)cnt;
        const double mx = warp_sum(x) * inv, my = warp_sum(y) * inv, mz = warp_sum(z) * inv;   // :758-763
        const double dx = has ? x - mx : 0.0, dy = has ? y - my : 0.0, dz = has ? z - mz : 0.0;
        const double c00 = warp_sum(dx * dx) * inv, c01 = warp_sum(dx * dy) * inv, c02 = warp_sum(dx * dz) * inv;   // :766-771
        const double c11 = warp_sum(dy * dy) * inv, c12 = warp_sum(dy * dz) * inv, c22 = warp_sum(dz * dz) * inv;
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
  k_pca_normals<<<plo_grid(c, 8), 256, 0, c->stream>>>(c->map_view(), c->dprm, c->nrm_pca.as<double>(), (int)c->n_pad_t);
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  c->pca_valid = true;
  return PLO_OK;
}

int plo_launch_project(plo_ctx* c, bool hooks) {
  if (c->m_raw == 0) return PLO_OK;
  ProjectOut out;
  out.qx = c->q_x.as<float4>(); out.qy = c->q_y.as<float4>(); out.qn = c->q_n.as<float4>();
  out.status = c->q_status.as<int>();
  out.height = c->q_height.as<double>(); out.nn1_idx = c->q_nn1_idx.as<int>(); out.nn1_d2 = c->q_nn1_d2.as<double>();
  out.nn_idx = c->q_nn_idx.as<int>(); out.nn_d2 = c->q_nn_d2.as<double>();
  out.search_stats = c->q_stats.as<int>();
  const int64_t warps = c->m_raw;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>((warps + 7) / 8, (int64_t)plo_grid(c, 8)));
  if (c->dprm.use_pca_normals)
    k_project<true><<<blocks, 256, 0, c->stream>>>(c->map_view(), c->s_p.as<float4>(), c->s_n.as<float4>(),
                                                   c->counts.as<DevCounts>(), c->state.as<DevState>(), c->dprm, out, hooks ? 1 : 0);
  else
    k_project<false><<<blocks, 256, 0, c->stream>>>(c->map_view(), c->s_p.as<float4>(), c->s_n.as<float4>(),
                                                    c->counts.as<DevCounts>(), c->state.as<DevState>(), c->dprm, out, hooks ? 1 : 0);
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  return PLO_OK;
}
