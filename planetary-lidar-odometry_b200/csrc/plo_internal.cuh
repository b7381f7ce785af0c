// plo_internal.cuh — shared declarations of the CUDA implementation behind plo_c_api.h.
// Compiled for sm_100a only.  The whole library is built with -fmad=false: the distance,
// box-bound and transform arithmetic must round exactly like the reference's double code
// (no FMA contraction) so that neighbour sets are bit-exact.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "plo/plo_c_api.h"

#define PLO_MAX_LEVELS 6          // 32 points/leaf * 32^5 nodes > 1e9 points
#define PLO_LEAF 32               // points per leaf == warp size
#define PLO_FANOUT 32             // children per internal node == warp size
#define PLO_MAX_K 32              // neighbour list lives one entry per lane
#define PLO_NSUM 36               // 21 H + 6 g + sw + swbb + count + 6 drop counters
#define PLO_FULL_MASK 0xffffffffu
#ifndef PLO_CHUNK_COLD
#define PLO_CHUNK_COLD 8   // no temporal bound (first projection, or after a large pose step)
#endif
#ifndef PLO_CHUNK_MID
#define PLO_CHUNK_MID 4    // previous projection available, step size unknown
#endif
#ifndef PLO_CHUNK_WARM
#define PLO_CHUNK_WARM 1   // small pose step: temporal bound is tight
#endif

// ---------------------------------------------------------------------------------
// device-side views
// ---------------------------------------------------------------------------------

// Spatial index over the target ("map"): points sorted by 48-bit Morton key, cut into
// leaves of 32 consecutive points; level l+1 groups 32 consecutive nodes of level l.
// Every level stores one AABB per node as two float4 (lo.xyz, hi.xyz); arrays are padded
// to a multiple of 32 nodes with empty boxes (lo=+inf, hi=-inf).
struct MapView {
  const float4* pts;      // [n_pad] sorted: x,y,z, w = int bits of the stripped-cloud index
  const float4* nrm;      // [n_pad] sorted: delivered normals (float), w unused
  const double* nrm_pca;  // [3*n_pad] sorted: PCA normals (double) when !is_get_normals
  const float4* lo[PLO_MAX_LEVELS];
  const float4* hi[PLO_MAX_LEVELS];
  int n_levels;           // stored levels; level n_levels-1 has <= 32 nodes
  int n_raw;              // points uploaded (before strip); 0 => empty map
};

// Device-resident loop state: written by the solve kernel, read by the next projection.
struct DevState {
  double rPose[16];
  double delta[16];
  double H[21];
  double g[6];
  double x0[6];   // trimmed LS: solution of the untrimmed first pass (src/solver.cpp:107)
  double Tbest[16];   // RANSAC: best hypothesis (src/solver.cpp:317-320), identity otherwise
  double U[36];       // DRPM: eigenvectors of the weighted information matrix (columns), ascending eigenvalues
  double ev[6];
  double probs[6];    // DRPM: non-degeneracy probabilities (include/degeneracy.h:74-105)
  long long ransac_best;   // inliers of the best hypothesis
  int ransac_iters;        // hypotheses evaluated
  int pad0;
  double sw, swbb;
  double rms, delta_dist, delta_angle;
  long long pairs;
  long long dropped[6];
  int iters;      // solves performed
  int status;     // 0 while running, else plo_reg_status
  int rank;
  int done;       // kernels of later iterations return immediately when set
  int use_prev;   // q_x / q_kd2 hold a previous projection of the same clouds (temporal bound usable)
  int chunk;      // consecutive source points per warp in the next projection (carry bound vs. balance)
  int warm;       // the last pose step was small: the next projection's queries barely move (short chunks, tiles are left behind)
  int miss_hist[32];   // per projection of this registration: queries the tile path handed to the tree (-1: it did not run)
  int tiles_ready; // every query has a candidate tile (or an invalid mark) from a projection of the same clouds: k_project_settled runs
};

struct DevCounts {
  int n_target;   // finite target points
  int n_source;   // finite source points
  int n_pairs;    // compaction result of plo_get_pairs
  int pad;
};

// kernel-side copy of the parameters that matter on the device
struct DevParams {
  double h2, r2, r_normal2;
  double angle_thr, cos_thr;
  double delta_dist_thr, delta_angle_thr;
  double ransac_dist_thr, huber_thr2;
  int k, k_normal;
  int use_pca_normals, angle_constraint, transform_normal;
  int correspond_number, weight_mode, iterations;
  int solver;            // plo_solver
  double ls_threshold;   // LS.threshold (trim fraction at either end)
  int ransac_max_iterations, ransac_final;
  double ransac_min_inliers_pct;
  double drpm_threshold, drpm_sp2, drpm_sn2;
  unsigned long long ransac_seed;
  int ext_weights;       // the pair weights come from the caller (host-vector solver entry points): no Huber/exp evaluation, no normalisation
  int pad_ext;
};

// ---------------------------------------------------------------------------------
// host-side context
// ---------------------------------------------------------------------------------

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t bytes);
  void release();
  template <typename T> T* as() const { return reinterpret_cast<T*>(p); }
};

struct plo_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  plo_params prm;
  DevParams dprm;
  std::string err;
  int64_t launches = 0;

  // target
  int64_t n_raw_t = 0;
  int64_t n_pad_t = 0;
  bool have_target = false;
  bool pca_valid = false;
  int n_levels = 0;
  int64_t level_nodes[PLO_MAX_LEVELS] = {0};
  DevBuf t_stage, t_stage2, t_praw, t_nraw, t_cidx, blockcnt, bbox;
  DevBuf keys[2], vals[2], hist, digit_total;
  DevBuf pts_sorted, nrm_sorted, nrm_pca, pos_of_cidx;
  DevBuf lvl_lo[PLO_MAX_LEVELS], lvl_hi[PLO_MAX_LEVELS];

  // source
  int64_t m_raw = 0;
  bool have_source = false;
  DevBuf s_stage, s_stage2, s_praw, s_nraw, s_p, s_n;

  // device-resident local map (plo_map_push): frames back to back as 32-byte records {x,y,z,-,nx,ny,nz,-},
  // all expressed in the frame of the most recent push; two buffers, swapped on every push
  DevBuf map_rec[2];
  int map_cur = 0;
  std::vector<int64_t> map_frames;   // points per queued frame, oldest first

  // front-end (plo_frontend): scratch and the filtered cloud of the last run, device-resident
  DevBuf fe_stage, fe_counts, fe_blockcnt, fe_kp, fe_ring, fe_inten, fe_rp, fe_rsrc, fe_nn[2], fe_status, fe_nrm, fe_ev;
  DevBuf fe_rec, fe_ev3, fe_cand, fe_src, fe_keys[2], fe_vals[2], fe_hist, fe_tot;
  int64_t fe_n_in = 0;
  bool fe_valid = false;

  // per-query results of the last projection
  DevBuf q_x, q_y, q_n, q_status, q_kd2;
  DevBuf q_tile_pts, q_tile_meta;   // per-query candidate tiles (knn_search.cuh): 64 x float4 + x_ref / e2
  // one projection's device-side bookkeeping (knn_project.cu): counters, the settled kernel's miss list
  DevBuf sync_counters, miss_list;
  bool prev_valid = false;   // q_x / q_kd2 hold the previous projection of the SAME clouds and k, r
  bool hooks_valid = false;
  DevBuf q_height, q_nn1_idx, q_nn1_d2, q_nn_idx, q_nn_d2, q_stats;
  bool projected = false;

  // reduction / solve
  DevBuf partials, state, counts, scratch, reduce_ticket, loop_barrier;
  DevBuf ls_keys[2], ls_vals[2], ls_hist, ls_tot, ls_mask;   // trimmed-LS selection
  DevBuf ransac_mind, partials2;                              // RANSAC hypothesis table (RansacScratch), DRPM noise partials
  DevBuf h_src, h_ref, h_nrm, h_w;   // plo_solve_wls_host staging
  DevBuf h_wext, counts_saved;       // host-vector LS / RANSAC / DRPM entry points: caller weights, the context's own counts
  const double* host_w = nullptr;    // != nullptr while such a call runs with caller weights
  bool host_drpm_only = false;       // plo_solve_drpm_host: DRPM tail without the RANSAC front
  DevState* h_state = nullptr;       // pinned
  DevCounts* h_counts = nullptr;     // pinned

  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  float ms_index = 0.f, ms_register = 0.f;
  bool ev_index_pending = false, ev_reg_pending = false;
  // resident loop as a CUDA graph: a WHILE conditional node whose body is one ICP iteration
  cudaGraph_t loop_graph = nullptr;
  cudaGraphExec_t loop_exec = nullptr;
  std::vector<unsigned long long> loop_sig;
  // batched mode: host->device copies of pair i+1 overlap the registration of pair i
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copied[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr}, ev_batch_start = nullptr;
  bool stage_used[2] = {false, false};   // target / source staging buffer has a pending consumer event
  DevBuf batch_slots;                    // plo_register_batch: per-unit result slots (device), pinned host mirror below
  DevState* h_batch_slots = nullptr;
  size_t h_batch_cap = 0;
  int body_launches = 2;   // kernels per loop iteration of the captured body
  // tuning knobs: environment read once at plo_create (PLO_CHUNK, PLO_NO_GRAPH), or plo_set_tuning
  int tune_group = 0;        // > 0: queries per warp and fetch on the settled path (power of two <= 32)
  int tune_chunk = -1;       // >= 0: chunk length of the tree walk (0 = device-side policy)
  bool tune_no_graph = false;   // enqueue-all loop instead of the conditional graph (ncu cannot profile kernel nodes of such graphs)
  bool tune_loop_kernel = true; // resident weighted-LS loop as ONE cooperative launch (k_register_loop); PLO_LOOP_KERNEL=0: graph of launches
  bool tune_fuse = true;        // resident weighted-LS loop: reduce + solve + loop tail in one launch (PLO_FUSE=0: the two stand-alone kernels)
  bool tune_force_warm = false; // projections start in the settled regime (tiles stored / used) -- parity tests of k_project_settled
  bool graph_launched = false;   // the last enqueue_register went through the graph
  bool loop_kernel_launched = false;   // ... through k_register_loop (one launch)
  bool coop_ok = true;           // the device supports cooperative launches
  bool graph_ok = true;      // cleared if the driver rejects conditional nodes: falls back to enqueue-all
  bool profiling = false;
  std::vector<cudaEvent_t> ev_proj;   // 2 per loop iteration when profiling
  float ms_project_mean = 0.f;
  float ms_project_each[64] = {0.f};   // per ICP iteration (profiling mode)
  int n_project = 0;

  MapView map_view() const;
};

int plo_fail(plo_ctx* c, int code, const std::string& msg);
#define PLO_CUDA(c, expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return plo_fail((c), PLO_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)
#define PLO_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != PLO_OK) return _r; \
  } while (0)

// launch geometry: persistent grids sized from the SM count
static inline int plo_grid(const plo_ctx* c, int blocks_per_sm) { return c->sm_count * blocks_per_sm; }

// ---- index_build.cu ---------------------------------------------------------------
int plo_build_index(plo_ctx* c, const void* dev_records, int64_t n, int32_t stride);
int plo_upload_source(plo_ctx* c, const void* dev_records, int64_t n, int32_t stride);
int plo_sort_pairs(plo_ctx* c, unsigned long long* keys[2], int* vals[2], int64_t n, int passes, int* hist, int* digit_total,
                   int* out_which, int first_shift = 0);
size_t plo_sort_hist_ints(int64_t n);
size_t plo_sort_total_ints(int passes);
int plo_map_push_records(plo_ctx* c, const void* dev_records, int64_t n, int32_t stride, const double* T_host_or_null,
                         bool pose_from_device, int32_t max_queue, bool transform_normals);
// ---- frontend.cu ------------------------------------------------------------------
int plo_frontend_run(plo_ctx* c, const void* dev_records, int64_t n, int32_t stride, const plo_frontend_params* fp);
int plo_frontend_fetch_counts(plo_ctx* c, int64_t out7[7]);
// ---- knn_project.cu ---------------------------------------------------------------
int plo_launch_pca_normals(plo_ctx* c);
int plo_launch_project(plo_ctx* c, bool hooks);
int plo_launch_register_loop(plo_ctx* c);   // the whole weighted-LS loop in one cooperative launch
int plo_loop_blocks(const plo_ctx* c);       // blocks of the persistent projection / loop grids
int plo_reserve_query_buffers(plo_ctx* c, bool hooks);
int plo_launch_imls_height(plo_ctx* c, const float* d_pts6, int n, double* d_height, int* d_ok);
int plo_launch_compute_normal(plo_ctx* c, const double* d_pts3, int n, double* d_out);
// ---- p2plane_solve.cu -------------------------------------------------------------
int plo_launch_reduce_solve(plo_ctx* c, bool advance_loop, unsigned long long cond_handle = 0ull);
int plo_reserve_solver_buffers(plo_ctx* c);
int plo_launch_reduce_solve_host_pairs(plo_ctx* c, const double* d_src, const double* d_ref,
                                       const double* d_nrm, const double* d_w, int64_t n);
int plo_launch_init_state(plo_ctx* c, const double* T0_host_or_null);
int plo_launch_compact_pairs(plo_ctx* c, float* d_src, float* d_ref, float* d_nrm, int32_t* d_idx);
