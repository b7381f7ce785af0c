// plo_api.cu — implementation of the C ABI declared in include/plo/plo_c_api.h.
// Host-side plumbing only: context, parameter flattening, staging copies, launch
// sequencing of the resident loop (src/laser_odometry.cpp:524-647).  No arithmetic of the
// hot path happens on the host and there is no CPU fallback.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <new>
#include <vector>

#include "plo_internal.cuh"

static std::string g_create_error;
extern "C" {
static void destroy_loop_graph(plo_ctx* c);
}

cudaError_t DevBuf::reserve(size_t bytes) {
  if (bytes <= cap) return cudaSuccess;
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    want = bytes;
    e = cudaMalloc(&p, want);
  }
  if (e == cudaSuccess) cap = want;
  return e;
}

void DevBuf::release() {
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
}

int plo_fail(plo_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_error = msg;
  return code;
}

MapView plo_ctx::map_view() const {
  MapView m;
  m.pts = pts_sorted.as<float4>();
  m.nrm = nrm_sorted.as<float4>();
  m.nrm_pca = nrm_pca.as<double>();
  for (int l = 0; l < PLO_MAX_LEVELS; ++l) {
    m.lo[l] = lvl_lo[l].as<float4>();
    m.hi[l] = lvl_hi[l].as<float4>();
  }
  m.n_levels = n_levels;
  m.n_raw = (int)n_raw_t;
  return m;
}

static void flatten_params(plo_ctx* c) {
  const plo_params& p = c->prm;
  DevParams& d = c->dprm;
  d.h2 = p.h * p.h;                 // src/imls_icp.cpp:620
  d.r2 = p.r * p.r;                 // libnabo: maxRadius2 = maxRadius * maxRadius
  d.r_normal2 = p.r_normal * p.r_normal;
  d.angle_thr = p.angle_diff_threshold;
  // cosine shortcut of the `angle > thr` test; thresholds outside [0,180) never / always fire
  if (!(p.angle_diff_threshold < 180.0)) d.cos_thr = -2.0;
  else if (p.angle_diff_threshold < 0.0) d.cos_thr = 2.0;
  else d.cos_thr = cos(p.angle_diff_threshold * 3.14159265358979323846 / 180.0);
  d.delta_dist_thr = p.delta_dist_threshold;
  d.delta_angle_thr = p.delta_angle_threshold;
  d.ransac_dist_thr = p.ransac_distance_threshold;
  d.huber_thr2 = p.huber_threshold * p.ransac_distance_threshold;   // src/solver.cpp:339
  d.k = p.search_number;
  d.k_normal = p.search_number_normal;
  d.use_pca_normals = p.is_get_normals ? 0 : 1;
  d.angle_constraint = p.normal_angle_constraint ? 1 : 0;
  d.transform_normal = p.transform_normal ? 1 : 0;
  d.correspond_number = p.correspond_number;
  d.weight_mode = p.weight_mode;
  d.iterations = p.iterations;
  d.solver = p.solver;
  d.ls_threshold = p.ls_threshold;
  d.ransac_max_iterations = p.ransac_max_iterations;
  d.ransac_final = p.ransac_final;
  d.ransac_min_inliers_pct = p.ransac_min_inliers_percentage;
  d.drpm_threshold = p.drpm_threshold;
  d.drpm_sp2 = p.drpm_stdev_points * p.drpm_stdev_points;     // include/degeneracy.h:49
  d.drpm_sn2 = p.drpm_stdev_normals * p.drpm_stdev_normals;   // src/solver.cpp:486-497
  d.ransac_seed = p.ransac_seed ? p.ransac_seed : 1ull;
  d.ext_weights = 0;
  d.pad_ext = 0;
}

extern "C" {

int plo_version(void) { return 100; }

void plo_default_params(plo_params* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->iterations = 30;                      // config.json:134
  p->h = 1.0;                              // :91
  p->r = 3.0;                              // :92
  p->r_normal = 1.0;                       // :103
  p->is_get_normals = 1;                   // :102
  p->search_number_normal = 10;            // :104
  p->search_number = 20;                   // :116
  p->normal_angle_constraint = 1;          // :111
  p->angle_diff_threshold = 30.0;          // :112
  p->transform_normal = 0;                 // :85
  p->correspond_number = 6;                // :89
  p->delta_dist_threshold = 0.001;         // :135
  p->delta_angle_threshold = 0.0001745353; // :136
  p->weight_mode = PLO_W_UNIT;
  p->ransac_distance_threshold = 0.8;      // :146
  p->huber_threshold = 0.648;              // :148
  p->solver = PLO_SOLVER_WLS;
  p->ls_threshold = 0.02;                  // :141
  p->ransac_max_iterations = 5000;         // :145
  p->ransac_min_inliers_percentage = 0.95; // :147
  p->ransac_final = PLO_FINAL_DRPM;        // :149
  p->drpm_threshold = 0.05;                // :151
  p->drpm_stdev_points = 0.02;             // :152
  p->drpm_stdev_normals = 0.05;            // :153
  p->ransac_seed = 1;
}

int plo_create(int device, plo_ctx** out) {
  if (!out) return plo_fail(nullptr, PLO_ERR_INVALID_ARG, "plo_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return plo_fail(nullptr, PLO_ERR_NO_DEVICE,
                    std::string("plo_create: no CUDA device (") + cudaGetErrorString(e) + "); there is no CPU fallback");
  if (device < 0 || device >= ndev) return plo_fail(nullptr, PLO_ERR_INVALID_ARG, "plo_create: bad device ordinal");
  plo_ctx* c = new (std::nothrow) plo_ctx();
  if (!c) return plo_fail(nullptr, PLO_ERR_STATE, "plo_create: out of host memory");
  c->device = device;
  plo_default_params(&c->prm);
  flatten_params(c);
#define CREATE_CUDA(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      plo_fail(nullptr, PLO_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
      plo_destroy(c);                                                                       \
      return PLO_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)
  CREATE_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  CREATE_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) {
    plo_fail(nullptr, PLO_ERR_UNSUPPORTED, "plo_create: device is not sm_100 class (library is built for sm_100a only)");
    plo_destroy(c);
    return PLO_ERR_UNSUPPORTED;
  }
  c->sm_count = prop.multiProcessorCount;
  c->coop_ok = prop.cooperativeLaunch != 0;
  // tuning knobs from the environment, read once (plo_set_tuning overrides them later)
  if (const char* e = getenv("PLO_CHUNK")) c->tune_chunk = std::max(0, atoi(e));
  if (const char* e = getenv("PLO_NO_GRAPH")) c->tune_no_graph = atoi(e) != 0;
  if (const char* e = getenv("PLO_FUSE")) c->tune_fuse = atoi(e) != 0;
  if (const char* e = getenv("PLO_LOOP_KERNEL")) c->tune_loop_kernel = atoi(e) != 0;
  CREATE_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->own_stream = true;
  for (int i = 0; i < 4; ++i) CREATE_CUDA(cudaEventCreate(&c->ev[i]));
  CREATE_CUDA(c->state.reserve(sizeof(DevState)));
  CREATE_CUDA(c->counts.reserve(sizeof(DevCounts)));
  CREATE_CUDA(cudaMemsetAsync(c->counts.p, 0, sizeof(DevCounts), c->stream));
  CREATE_CUDA(cudaMemsetAsync(c->state.p, 0, sizeof(DevState), c->stream));
  CREATE_CUDA(cudaMallocHost(&c->h_state, sizeof(DevState)));
  CREATE_CUDA(cudaMallocHost(&c->h_counts, sizeof(DevCounts)));
  CREATE_CUDA(cudaStreamSynchronize(c->stream));
#undef CREATE_CUDA
  *out = c;
  return PLO_OK;
}

void plo_destroy(plo_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  DevBuf* bufs[] = {&c->t_stage, &c->t_stage2, &c->s_stage2, &c->t_praw, &c->t_nraw, &c->t_cidx, &c->blockcnt, &c->bbox, &c->keys[0], &c->keys[1],
                    &c->vals[0], &c->vals[1], &c->hist, &c->digit_total, &c->pts_sorted, &c->nrm_sorted, &c->nrm_pca, &c->pos_of_cidx,
                    &c->s_stage, &c->s_praw, &c->s_nraw, &c->s_p, &c->s_n, &c->map_rec[0], &c->map_rec[1], &c->fe_stage, &c->fe_counts, &c->fe_blockcnt, &c->fe_kp, &c->fe_ring, &c->fe_inten, &c->fe_rp, &c->fe_rsrc, &c->fe_nn[0], &c->fe_nn[1], &c->fe_status, &c->fe_nrm, &c->fe_ev, &c->fe_rec, &c->fe_ev3, &c->fe_cand, &c->fe_src, &c->fe_keys[0], &c->fe_keys[1], &c->fe_vals[0], &c->fe_vals[1], &c->fe_hist, &c->fe_tot, &c->q_x, &c->q_y, &c->q_n, &c->q_status, &c->q_kd2, &c->q_tile_pts, &c->q_tile_meta, &c->sync_counters, &c->miss_list, &c->reduce_ticket, &c->loop_barrier, &c->h_wext, &c->counts_saved, &c->batch_slots,
                    &c->q_height, &c->q_nn1_idx, &c->q_nn1_d2, &c->q_nn_idx, &c->q_nn_d2, &c->q_stats, &c->partials, &c->state,
                    &c->counts, &c->scratch, &c->ls_keys[0], &c->ls_keys[1], &c->ls_vals[0], &c->ls_vals[1], &c->ls_hist, &c->ls_tot, &c->ls_mask, &c->ransac_mind, &c->partials2, &c->h_src, &c->h_ref, &c->h_nrm, &c->h_w};
  for (DevBuf* b : bufs) b->release();
  for (int l = 0; l < PLO_MAX_LEVELS; ++l) { c->lvl_lo[l].release(); c->lvl_hi[l].release(); }
  if (c->h_state) cudaFreeHost(c->h_state);
  if (c->h_counts) cudaFreeHost(c->h_counts);
  if (c->h_batch_slots) cudaFreeHost(c->h_batch_slots);
  for (int i = 0; i < 4; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
  for (cudaEvent_t e : c->ev_proj) cudaEventDestroy(e);
  destroy_loop_graph(c);
  for (int i = 0; i < 2; ++i) { if (c->ev_copied[i]) cudaEventDestroy(c->ev_copied[i]); if (c->ev_consumed[i]) cudaEventDestroy(c->ev_consumed[i]); }
  if (c->ev_batch_start) cudaEventDestroy(c->ev_batch_start);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

const char* plo_last_error(const plo_ctx* c) { return c ? c->err.c_str() : g_create_error.c_str(); }

int plo_set_stream(plo_ctx* c, void* cuda_stream) {
  if (!c) return PLO_ERR_INVALID_ARG;
  PLO_CUDA(c, cudaSetDevice(c->device));
  PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  destroy_loop_graph(c);
  c->stream = static_cast<cudaStream_t>(cuda_stream);
  c->own_stream = false;
  return PLO_OK;
}

int plo_set_tuning(plo_ctx* c, const char* name, int32_t value) {
  if (!c || !name) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_set_tuning: NULL argument");
  const std::string n(name);
  if (n == "chunk") c->tune_chunk = value;
  else if (n == "group") c->tune_group = (value >= 1 && value <= 32 && (value & (value - 1)) == 0) ? value : 0;
  else if (n == "no_graph") c->tune_no_graph = value != 0;
  else if (n == "force_warm") c->tune_force_warm = value != 0;
  else if (n == "fuse") c->tune_fuse = value != 0;
  else if (n == "loop_kernel") c->tune_loop_kernel = value != 0;
  else return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_set_tuning: unknown knob '" + n + "' (chunk, group, no_graph, force_warm, fuse, loop_kernel)");
  return PLO_OK;
}

int plo_stream_wait_event(plo_ctx* c, void* cuda_event) {
  if (!c || !cuda_event) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_stream_wait_event: NULL argument");
  PLO_CUDA(c, cudaSetDevice(c->device));
  PLO_CUDA(c, cudaStreamWaitEvent(c->stream, static_cast<cudaEvent_t>(cuda_event), 0));
  return PLO_OK;
}

int plo_synchronize(plo_ctx* c) {
  if (!c) return PLO_ERR_INVALID_ARG;
  PLO_CUDA(c, cudaSetDevice(c->device));
  PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->copy_stream) PLO_CUDA(c, cudaStreamSynchronize(c->copy_stream));
  return PLO_OK;
}

int plo_set_params(plo_ctx* c, const plo_params* p) {
  if (!c || !p) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_set_params: NULL argument");
  if (p->search_number < 1 || p->search_number > PLO_MAX_K)
    return plo_fail(c, PLO_ERR_UNSUPPORTED, "plo_set_params: search_number must be in [1, 32]");
  if (!p->is_get_normals && (p->search_number_normal < 1 || p->search_number_normal > PLO_MAX_K))
    return plo_fail(c, PLO_ERR_UNSUPPORTED, "plo_set_params: search_number_normal must be in [1, 32]");
  if (p->iterations < 0) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_set_params: iterations < 0");
  if (!(p->r >= 0.0) || !(p->h >= 0.0) || !(p->r_normal >= 0.0))
    return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_set_params: h, r, r_normal must be >= 0");
  if (p->weight_mode != PLO_W_UNIT && p->weight_mode != PLO_W_HUBER_EXP)
    return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_set_params: unknown weight_mode");
  if (p->solver != PLO_SOLVER_WLS && p->solver != PLO_SOLVER_LS && p->solver != PLO_SOLVER_RANSAC)
    return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_set_params: unknown solver");
  if (p->solver == PLO_SOLVER_RANSAC && p->ransac_final != PLO_FINAL_LS && p->ransac_final != PLO_FINAL_WLS &&
      p->ransac_final != PLO_FINAL_DRPM)
    return plo_fail(c, PLO_ERR_UNSUPPORTED, "plo_set_params: RANSAC final_solve_method must be LS, Weighted LS or DRPM");
  const bool trims = p->solver == PLO_SOLVER_LS || (p->solver == PLO_SOLVER_RANSAC && p->ransac_final == PLO_FINAL_LS);
  if (trims && !(p->ls_threshold >= 0.0 && p->ls_threshold < 0.5))
    return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_set_params: ls_threshold must be in [0, 0.5)");
  const bool pca_changed = c->prm.r_normal != p->r_normal || c->prm.search_number_normal != p->search_number_normal ||
                           c->prm.is_get_normals != p->is_get_normals;
  c->prm = *p;
  flatten_params(c);
  if (pca_changed) c->pca_valid = false;
  c->hooks_valid = false;
  c->prev_valid = false;
  return PLO_OK;
}

// copy stream + the events that order it against the main stream (created on first use)
static int ensure_copy_stream(plo_ctx* c) {
  if (!c->copy_stream) PLO_CUDA(c, cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  for (int b = 0; b < 2; ++b) {
    if (!c->ev_copied[b]) PLO_CUDA(c, cudaEventCreateWithFlags(&c->ev_copied[b], cudaEventDisableTiming));
    if (!c->ev_consumed[b]) PLO_CUDA(c, cudaEventCreateWithFlags(&c->ev_consumed[b], cudaEventDisableTiming));
  }
  if (!c->ev_batch_start) PLO_CUDA(c, cudaEventCreateWithFlags(&c->ev_batch_start, cudaEventDisableTiming));
  return PLO_OK;
}

// Host records -> staging buffer on the COPY stream, unpack / index build on the main stream behind an event: the
// upload of a cloud overlaps whatever the main stream is still doing (set_target then set_source: the source crosses
// the bus while the index of the target is being built).  slot 0 = target staging, 1 = source staging.
static int stage_and_run(plo_ctx* c, const void* host_pts, int64_t n, int32_t stride, DevBuf& stage, bool target) {
  if (n < 0 || (n > 0 && !host_pts)) return plo_fail(c, PLO_ERR_INVALID_ARG, "set cloud: bad pointer / count");
  if (stride < 28 || (stride % 4) != 0) return plo_fail(c, PLO_ERR_INVALID_ARG, "set cloud: stride must be >= 28 and a multiple of 4");
  PLO_CUDA(c, cudaSetDevice(c->device));
  if (n > 0) {
    PLO_TRY(ensure_copy_stream(c));
    const int b = target ? 0 : 1;
    const void* before = stage.p;
    PLO_CUDA(c, stage.reserve((size_t)n * stride));
    // the previous user of this staging buffer (main stream) must be done with it; a re-allocation frees synchronously
    if (stage.p == before && c->stage_used[b]) PLO_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[b], 0));
    PLO_CUDA(c, cudaMemcpyAsync(stage.p, host_pts, (size_t)n * stride, cudaMemcpyHostToDevice, c->copy_stream));
    PLO_CUDA(c, cudaEventRecord(c->ev_copied[b], c->copy_stream));
    PLO_CUDA(c, cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0));
  }
  const int rc = target ? plo_build_index(c, stage.p, n, stride) : plo_upload_source(c, stage.p, n, stride);
  if (n > 0 && rc == PLO_OK) {
    PLO_CUDA(c, cudaEventRecord(c->ev_consumed[target ? 0 : 1], c->stream));
    c->stage_used[target ? 0 : 1] = true;
  }
  return rc;
}

int plo_set_target(plo_ctx* c, const void* host_pts, int64_t n, int32_t stride) {
  if (!c) return PLO_ERR_INVALID_ARG;
  return stage_and_run(c, host_pts, n, stride, c->t_stage, true);
}

int plo_set_source(plo_ctx* c, const void* host_pts, int64_t n, int32_t stride) {
  if (!c) return PLO_ERR_INVALID_ARG;
  return stage_and_run(c, host_pts, n, stride, c->s_stage, false);
}

int plo_set_target_device(plo_ctx* c, const void* dev_pts, int64_t n, int32_t stride) {
  if (!c) return PLO_ERR_INVALID_ARG;
  if (n < 0 || (n > 0 && !dev_pts)) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_set_target_device: bad pointer / count");
  if (stride < 28 || (stride % 4) != 0) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_set_target_device: bad stride");
  PLO_CUDA(c, cudaSetDevice(c->device));
  return plo_build_index(c, dev_pts, n, stride);
}

int plo_set_source_device(plo_ctx* c, const void* dev_pts, int64_t n, int32_t stride) {
  if (!c) return PLO_ERR_INVALID_ARG;
  if (n < 0 || (n > 0 && !dev_pts)) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_set_source_device: bad pointer / count");
  if (stride < 28 || (stride % 4) != 0) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_set_source_device: bad stride");
  PLO_CUDA(c, cudaSetDevice(c->device));
  return plo_upload_source(c, dev_pts, n, stride);
}

int plo_map_reset(plo_ctx* c) {
  if (!c) return PLO_ERR_INVALID_ARG;
  c->map_frames.clear();
  return PLO_OK;
}

static int map_push_common(plo_ctx* c, const void* dev_records, int64_t n, int32_t stride, const double* T_last_curr,
                           int32_t pose_from_last_register, int32_t max_queue, int32_t transform_normals) {
  if (pose_from_last_register && !c->projected)
    return plo_fail(c, PLO_ERR_STATE, "plo_map_push: pose_from_last_register without a registration on this context");
  return plo_map_push_records(c, dev_records, n, stride, pose_from_last_register ? nullptr : T_last_curr,
                              pose_from_last_register != 0, max_queue, transform_normals != 0);
}

int plo_map_push(plo_ctx* c, const void* host_pts, int64_t n, int32_t stride, const double* T_last_curr,
                 int32_t pose_from_last_register, int32_t max_queue, int32_t transform_normals) {
  if (!c) return PLO_ERR_INVALID_ARG;
  if (n < 0 || (n > 0 && !host_pts)) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_map_push: bad pointer / count");
  if (stride < 28 || (stride % 4) != 0) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_map_push: stride must be >= 28 and a multiple of 4");
  PLO_CUDA(c, cudaSetDevice(c->device));
  if (n > 0) {
    PLO_CUDA(c, c->t_stage.reserve((size_t)n * stride));
    PLO_CUDA(c, cudaMemcpyAsync(c->t_stage.p, host_pts, (size_t)n * stride, cudaMemcpyHostToDevice, c->stream));
  }
  return map_push_common(c, c->t_stage.p, n, stride, T_last_curr, pose_from_last_register, max_queue, transform_normals);
}

int plo_map_push_device(plo_ctx* c, const void* dev_pts, int64_t n, int32_t stride, const double* T_last_curr,
                        int32_t pose_from_last_register, int32_t max_queue, int32_t transform_normals) {
  if (!c) return PLO_ERR_INVALID_ARG;
  if (n < 0 || (n > 0 && !dev_pts)) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_map_push_device: bad pointer / count");
  if (stride < 28 || (stride % 4) != 0) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_map_push_device: bad stride");
  PLO_CUDA(c, cudaSetDevice(c->device));
  return map_push_common(c, dev_pts, n, stride, T_last_curr, pose_from_last_register, max_queue, transform_normals);
}

int plo_map_info(plo_ctx* c, int32_t* frames, int64_t* points) {
  if (!c) return PLO_ERR_INVALID_ARG;
  int64_t total = 0;
  for (int64_t f : c->map_frames) total += f;
  if (frames) *frames = (int32_t)c->map_frames.size();
  if (points) *points = total;
  return PLO_OK;
}

int plo_map_get(plo_ctx* c, float* records8, int64_t cap) {
  if (!c || !records8) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_map_get: NULL argument");
  int64_t total = 0;
  for (int64_t f : c->map_frames) total += f;
  if (cap < total) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_map_get: buffer too small");
  if (total == 0) return PLO_OK;
  PLO_CUDA(c, cudaSetDevice(c->device));
  PLO_CUDA(c, cudaMemcpyAsync(records8, c->map_rec[c->map_cur].p, sizeof(float) * 8 * (size_t)total, cudaMemcpyDeviceToHost, c->stream));
  PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  return PLO_OK;
}

void plo_frontend_default_params(plo_frontend_params* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->n_scans = 64;
  p->min_range = 0.5f;                 // src/scan_registration.cpp:62
  p->max_range = 120.0f;               // :63
  p->scan_period = 0.1f;               // :55
  p->window_size = 3;                  // config.json:9
  p->iter_step = 1;                    // :10
  p->knn_distance_threshold = 10.0f;   // :11
  p->plane_distance_threshold = 0.02f; // :14
  p->valid_points_threshold = 0.8f;    // :15
  p->use_all_points = 1;               // :78
  p->planarity_threshold = 0.05f;      // :39
}

static int frontend_common(plo_ctx* c, const void* dev_pts, int64_t n, int32_t stride, const plo_frontend_params* p,
                           plo_frontend_stats* stats) {
  if (!p) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_frontend: params is NULL");
  if (p->n_scans != 16 && p->n_scans != 32 && p->n_scans != 64)
    return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_frontend: n_scans must be 16, 32 or 64 (\"wrong scan number\", scan_registration.cpp:1011)");
  if (p->window_size < 0 || p->iter_step < 1 || (2 * p->window_size) / p->iter_step + 1 > 64)
    return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_frontend: bad window_size / iter_step");
  PLO_TRY(plo_frontend_run(c, dev_pts, n, stride, p));
  if (stats) {
    int64_t v[7];
    PLO_TRY(plo_frontend_fetch_counts(c, v));
    stats->n_out = v[0]; stats->gated = v[1]; stats->ringed = v[2];
    stats->pca_failures = v[3]; stats->plane_failures = v[4]; stats->candidates = v[5];
  }
  return PLO_OK;
}

int plo_frontend(plo_ctx* c, const void* host_pts, int64_t n, int32_t stride, const plo_frontend_params* p, plo_frontend_stats* stats) {
  if (!c) return PLO_ERR_INVALID_ARG;
  if (n < 0 || (n > 0 && !host_pts)) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_frontend: bad pointer / count");
  if (stride < 12 || (stride % 4) != 0) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_frontend: stride must be >= 12 and a multiple of 4");
  PLO_CUDA(c, cudaSetDevice(c->device));
  if (n > 0) {
    PLO_CUDA(c, c->fe_stage.reserve((size_t)n * stride));
    PLO_CUDA(c, cudaMemcpyAsync(c->fe_stage.p, host_pts, (size_t)n * stride, cudaMemcpyHostToDevice, c->stream));
  }
  return frontend_common(c, c->fe_stage.p, n, stride, p, stats);
}

int plo_frontend_device(plo_ctx* c, const void* dev_pts, int64_t n, int32_t stride, const plo_frontend_params* p,
                        plo_frontend_stats* stats) {
  if (!c) return PLO_ERR_INVALID_ARG;
  if (n < 0 || (n > 0 && !dev_pts)) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_frontend_device: bad pointer / count");
  if (stride < 12 || (stride % 4) != 0) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_frontend_device: bad stride");
  PLO_CUDA(c, cudaSetDevice(c->device));
  return frontend_common(c, dev_pts, n, stride, p, stats);
}

int plo_frontend_get(plo_ctx* c, float* records12, float* eigenvalues3, uint8_t* candidate, int32_t* src_index, int64_t cap) {
  if (!c) return PLO_ERR_INVALID_ARG;
  if (!c->fe_valid) return plo_fail(c, PLO_ERR_STATE, "plo_frontend_get: call plo_frontend first");
  PLO_CUDA(c, cudaSetDevice(c->device));
  int64_t v[7];
  PLO_TRY(plo_frontend_fetch_counts(c, v));
  const size_t m = (size_t)v[0];
  if ((int64_t)m > cap) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_frontend_get: buffers too small");
  if (m == 0) return PLO_OK;
  if (records12) PLO_CUDA(c, cudaMemcpyAsync(records12, c->fe_rec.p, sizeof(float) * 12 * m, cudaMemcpyDeviceToHost, c->stream));
  if (eigenvalues3) PLO_CUDA(c, cudaMemcpyAsync(eigenvalues3, c->fe_ev3.p, sizeof(float) * 3 * m, cudaMemcpyDeviceToHost, c->stream));
  if (candidate) PLO_CUDA(c, cudaMemcpyAsync(candidate, c->fe_cand.p, m, cudaMemcpyDeviceToHost, c->stream));
  if (src_index) PLO_CUDA(c, cudaMemcpyAsync(src_index, c->fe_src.p, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, c->stream));
  PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  return PLO_OK;
}

int plo_frontend_device_records(plo_ctx* c, const void** dev_records, int64_t* n) {
  if (!c || !dev_records || !n) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_frontend_device_records: NULL argument");
  if (!c->fe_valid) return plo_fail(c, PLO_ERR_STATE, "plo_frontend_device_records: call plo_frontend first");
  PLO_CUDA(c, cudaSetDevice(c->device));
  int64_t v[7];
  PLO_TRY(plo_frontend_fetch_counts(c, v));
  *dev_records = c->fe_rec.p;
  *n = v[0];
  return PLO_OK;
}

static int fetch_counts(plo_ctx* c) {
  PLO_CUDA(c, cudaSetDevice(c->device));
  PLO_CUDA(c, cudaMemcpyAsync(c->h_counts, c->counts.p, sizeof(DevCounts), cudaMemcpyDeviceToHost, c->stream));
  PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  return PLO_OK;
}

int64_t plo_target_size(plo_ctx* c) {
  if (!c || !c->have_target) return -1;
  if (fetch_counts(c) != PLO_OK) return -1;
  return c->h_counts->n_target;
}

int64_t plo_source_size(plo_ctx* c) {
  if (!c || !c->have_source) return -1;
  if (fetch_counts(c) != PLO_OK) return -1;
  return c->h_counts->n_source;
}

static int require_clouds(plo_ctx* c, const char* who) {
  if (!c->have_target) return plo_fail(c, PLO_ERR_STATE, std::string(who) + ": no target cloud set");
  if (!c->have_source) return plo_fail(c, PLO_ERR_STATE, std::string(who) + ": no source cloud set");
  return PLO_OK;
}

static int fetch_state(plo_ctx* c) {
  PLO_CUDA(c, cudaMemcpyAsync(c->h_state, c->state.p, sizeof(DevState), cudaMemcpyDeviceToHost, c->stream));
  PLO_CUDA(c, cudaMemcpyAsync(c->h_counts, c->counts.p, sizeof(DevCounts), cudaMemcpyDeviceToHost, c->stream));
  PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  return PLO_OK;
}

int plo_project(plo_ctx* c, const double T[16], int32_t hooks, plo_proj_stats* stats) {
  if (!c || !T) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_project: NULL argument");
  PLO_TRY(require_clouds(c, "plo_project"));
  PLO_CUDA(c, cudaSetDevice(c->device));
  PLO_TRY(plo_reserve_query_buffers(c, hooks != 0));
  if (c->dprm.use_pca_normals) PLO_TRY(plo_launch_pca_normals(c));
  PLO_TRY(plo_launch_init_state(c, T));
  PLO_TRY(plo_launch_project(c, hooks != 0));
  c->projected = true;
  c->hooks_valid = hooks != 0;
  if (stats) {
    // the drop counters come out of the same reduction the solver uses
    const int saved_solver = c->dprm.solver;
    c->dprm.solver = PLO_SOLVER_WLS;
    const int rrc = plo_launch_reduce_solve(c, false);
    c->dprm.solver = saved_solver;
    PLO_TRY(rrc);
    PLO_TRY(fetch_state(c));
    stats->n_source = c->h_counts->n_source;
    stats->n_pairs = c->h_state->pairs;
    for (int i = 0; i < 6; ++i) stats->dropped[i] = c->h_state->dropped[i];
  }
  return PLO_OK;
}

int plo_get_pairs(plo_ctx* c, float* src_xyz, float* ref_xyz, float* ref_n, int32_t* src_idx, int64_t cap, int64_t* n) {
  if (!c || !n) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_get_pairs: NULL argument");
  if (!c->projected) return plo_fail(c, PLO_ERR_STATE, "plo_get_pairs: call plo_project first");
  PLO_CUDA(c, cudaSetDevice(c->device));
  const size_t m = (size_t)std::max<int64_t>(c->m_raw, 1);
  PLO_CUDA(c, c->h_src.reserve(sizeof(float) * 3 * m));
  PLO_CUDA(c, c->h_ref.reserve(sizeof(float) * 3 * m));
  PLO_CUDA(c, c->h_nrm.reserve(sizeof(float) * 3 * m));
  PLO_CUDA(c, c->h_w.reserve(sizeof(int32_t) * m));
  PLO_TRY(plo_launch_compact_pairs(c, c->h_src.as<float>(), c->h_ref.as<float>(), c->h_nrm.as<float>(), c->h_w.as<int32_t>()));
  PLO_TRY(fetch_counts(c));
  const int64_t np = c->h_counts->n_pairs;
  *n = np;
  if (np > cap && (src_xyz || ref_xyz || ref_n || src_idx))
    return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_get_pairs: capacity too small");
  if (np > 0) {
    if (src_xyz) PLO_CUDA(c, cudaMemcpyAsync(src_xyz, c->h_src.p, sizeof(float) * 3 * np, cudaMemcpyDeviceToHost, c->stream));
    if (ref_xyz) PLO_CUDA(c, cudaMemcpyAsync(ref_xyz, c->h_ref.p, sizeof(float) * 3 * np, cudaMemcpyDeviceToHost, c->stream));
    if (ref_n) PLO_CUDA(c, cudaMemcpyAsync(ref_n, c->h_nrm.p, sizeof(float) * 3 * np, cudaMemcpyDeviceToHost, c->stream));
    if (src_idx) PLO_CUDA(c, cudaMemcpyAsync(src_idx, c->h_w.p, sizeof(int32_t) * np, cudaMemcpyDeviceToHost, c->stream));
    PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return PLO_OK;
}

int plo_get_neighbors(plo_ctx* c, int32_t* nn_idx, double* nn_d2, int32_t* nn1_idx, double* nn1_d2) {
  if (!c) return PLO_ERR_INVALID_ARG;
  if (!c->projected || !c->hooks_valid) return plo_fail(c, PLO_ERR_STATE, "plo_get_neighbors: last plo_project had hooks == 0");
  PLO_CUDA(c, cudaSetDevice(c->device));
  PLO_TRY(fetch_counts(c));
  const size_t m = (size_t)c->h_counts->n_source;
  const size_t k = (size_t)c->prm.search_number;
  if (m == 0) return PLO_OK;
  if (nn_idx) PLO_CUDA(c, cudaMemcpyAsync(nn_idx, c->q_nn_idx.p, sizeof(int32_t) * m * k, cudaMemcpyDeviceToHost, c->stream));
  if (nn_d2) PLO_CUDA(c, cudaMemcpyAsync(nn_d2, c->q_nn_d2.p, sizeof(double) * m * k, cudaMemcpyDeviceToHost, c->stream));
  if (nn1_idx) PLO_CUDA(c, cudaMemcpyAsync(nn1_idx, c->q_nn1_idx.p, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, c->stream));
  if (nn1_d2) PLO_CUDA(c, cudaMemcpyAsync(nn1_d2, c->q_nn1_d2.p, sizeof(double) * m, cudaMemcpyDeviceToHost, c->stream));
  PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  return PLO_OK;
}

int plo_get_search_stats(plo_ctx* c, int32_t* stats3) {
  if (!c || !stats3) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_get_search_stats: NULL argument");
  if (!c->projected || !c->hooks_valid) return plo_fail(c, PLO_ERR_STATE, "plo_get_search_stats: last plo_project had hooks == 0");
  PLO_CUDA(c, cudaSetDevice(c->device));
  PLO_TRY(fetch_counts(c));
  const size_t m = (size_t)c->h_counts->n_source;
  if (m == 0) return PLO_OK;
  PLO_CUDA(c, cudaMemcpyAsync(stats3, c->q_stats.p, sizeof(int32_t) * 3 * m, cudaMemcpyDeviceToHost, c->stream));
  PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  return PLO_OK;
}

int plo_get_query_results(plo_ctx* c, int32_t* status, double* height) {
  if (!c) return PLO_ERR_INVALID_ARG;
  if (!c->projected) return plo_fail(c, PLO_ERR_STATE, "plo_get_query_results: call plo_project first");
  if (height && !c->hooks_valid) return plo_fail(c, PLO_ERR_STATE, "plo_get_query_results: heights need hooks != 0");
  PLO_CUDA(c, cudaSetDevice(c->device));
  PLO_TRY(fetch_counts(c));
  const size_t m = (size_t)c->h_counts->n_source;
  if (m == 0) return PLO_OK;
  if (status) PLO_CUDA(c, cudaMemcpyAsync(status, c->q_status.p, sizeof(int32_t) * m, cudaMemcpyDeviceToHost, c->stream));
  if (height) PLO_CUDA(c, cudaMemcpyAsync(height, c->q_height.p, sizeof(double) * m, cudaMemcpyDeviceToHost, c->stream));
  PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  return PLO_OK;
}

}  // extern "C"

// gather of the target normals into stripped-cloud order (device helper kernel)
__global__ void k_export_normals(const float4* __restrict__ nrm, const double* __restrict__ nrm_pca,
                                 const int* __restrict__ pos_of_cidx, const DevCounts* __restrict__ counts, int use_pca,
                                 double* __restrict__ out) {
  const int n = counts->n_target;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int pos = pos_of_cidx[i];
    if (use_pca) {
      out[3 * (size_t)i] = nrm_pca[3 * (size_t)pos];
      out[3 * (size_t)i + 1] = nrm_pca[3 * (size_t)pos + 1];
      out[3 * (size_t)i + 2] = nrm_pca[3 * (size_t)pos + 2];
    } else {
      const float4 v = nrm[pos];
      out[3 * (size_t)i] = (double)v.x;
      out[3 * (size_t)i + 1] = (double)v.y;
      out[3 * (size_t)i + 2] = (double)v.z;
    }
  }
}

extern "C" {

int plo_imls_height(plo_ctx* c, const float* xyz_normal6, int64_t n, double* height, int32_t* ok) {
  if (!c || n < 0 || (n > 0 && (!xyz_normal6 || !height || !ok))) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_imls_height: bad argument");
  if (!c->have_target) return plo_fail(c, PLO_ERR_STATE, "plo_imls_height: no target cloud set");
  if (n == 0) return PLO_OK;
  if (n > (int64_t)1 << 28) return plo_fail(c, PLO_ERR_UNSUPPORTED, "plo_imls_height: too many points");
  PLO_CUDA(c, cudaSetDevice(c->device));
  if (c->dprm.use_pca_normals) PLO_TRY(plo_launch_pca_normals(c));
  const size_t in_b = sizeof(float) * 6 * (size_t)n, h_b = sizeof(double) * (size_t)n, ok_b = sizeof(int) * (size_t)n;
  PLO_CUDA(c, c->scratch.reserve(in_b + h_b + ok_b + 64));
  char* base = c->scratch.as<char>();
  double* d_h = reinterpret_cast<double*>(base);
  float* d_in = reinterpret_cast<float*>(base + h_b);
  int* d_ok = reinterpret_cast<int*>(base + h_b + in_b);
  PLO_CUDA(c, cudaMemcpyAsync(d_in, xyz_normal6, in_b, cudaMemcpyHostToDevice, c->stream));
  PLO_TRY(plo_launch_imls_height(c, d_in, (int)n, d_h, d_ok));
  PLO_CUDA(c, cudaMemcpyAsync(height, d_h, h_b, cudaMemcpyDeviceToHost, c->stream));
  PLO_CUDA(c, cudaMemcpyAsync(ok, d_ok, ok_b, cudaMemcpyDeviceToHost, c->stream));
  PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  return PLO_OK;
}

int plo_compute_normal(plo_ctx* c, const double* pts3, int64_t n, double normal[3]) {
  if (!c || !normal || n < 0 || (n > 0 && !pts3)) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_compute_normal: bad argument");
  if (n > 1 << 20) return plo_fail(c, PLO_ERR_UNSUPPORTED, "plo_compute_normal: more than 2^20 points");
  PLO_CUDA(c, cudaSetDevice(c->device));
  PLO_CUDA(c, c->scratch.reserve(sizeof(double) * (3 * (size_t)std::max<int64_t>(n, 1) + 4)));
  double* d_p = c->scratch.as<double>() + 4;
  if (n > 0) PLO_CUDA(c, cudaMemcpyAsync(d_p, pts3, sizeof(double) * 3 * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  PLO_TRY(plo_launch_compute_normal(c, d_p, (int)n, c->scratch.as<double>()));
  PLO_CUDA(c, cudaMemcpyAsync(normal, c->scratch.p, sizeof(double) * 3, cudaMemcpyDeviceToHost, c->stream));
  PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  return PLO_OK;
}

int plo_get_target_normals(plo_ctx* c, double* out) {
  if (!c || !out) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_get_target_normals: NULL argument");
  if (!c->have_target) return plo_fail(c, PLO_ERR_STATE, "plo_get_target_normals: no target cloud set");
  PLO_CUDA(c, cudaSetDevice(c->device));
  if (c->n_raw_t == 0) return PLO_OK;
  if (c->dprm.use_pca_normals) PLO_TRY(plo_launch_pca_normals(c));
  PLO_CUDA(c, c->scratch.reserve(sizeof(double) * 3 * (size_t)c->n_raw_t));
  k_export_normals<<<plo_grid(c, 4), 256, 0, c->stream>>>(c->nrm_sorted.as<float4>(), c->nrm_pca.as<double>(),
                                                          c->pos_of_cidx.as<int>(), c->counts.as<DevCounts>(),
                                                          c->dprm.use_pca_normals, c->scratch.as<double>());
  c->launches++;
  PLO_CUDA(c, cudaGetLastError());
  PLO_TRY(fetch_counts(c));
  const size_t n = (size_t)c->h_counts->n_target;
  if (n > 0) {
    PLO_CUDA(c, cudaMemcpyAsync(out, c->scratch.p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  return PLO_OK;
}

static int solve_on_pairs(plo_ctx* c, double delta[16], int32_t* rank, int solver, const char* who) {
  if (!c || !delta) return plo_fail(c, PLO_ERR_INVALID_ARG, std::string(who) + ": NULL argument");
  if (!c->projected) return plo_fail(c, PLO_ERR_STATE, std::string(who) + ": call plo_project first");
  PLO_CUDA(c, cudaSetDevice(c->device));
  const int saved = c->dprm.solver;
  c->dprm.solver = solver;
  const int rc = plo_launch_reduce_solve(c, false);
  c->dprm.solver = saved;
  PLO_TRY(rc);
  PLO_TRY(fetch_state(c));
  memcpy(delta, c->h_state->delta, sizeof(double) * 16);
  if (rank) *rank = c->h_state->rank;
  return PLO_OK;
}

int plo_solve_wls(plo_ctx* c, double delta[16], int32_t* rank) { return solve_on_pairs(c, delta, rank, PLO_SOLVER_WLS, "plo_solve_wls"); }

int plo_solve_ls(plo_ctx* c, double delta[16], int32_t* rank) { return solve_on_pairs(c, delta, rank, PLO_SOLVER_LS, "plo_solve_ls"); }

int plo_solve_ransac(plo_ctx* c, double delta[16], double probs[6], int64_t* inliers, int32_t* hypotheses) {
  PLO_TRY(solve_on_pairs(c, delta, nullptr, PLO_SOLVER_RANSAC, "plo_solve_ransac"));
  if (probs) memcpy(probs, c->h_state->probs, sizeof(double) * 6);
  if (inliers) *inliers = c->h_state->ransac_best;
  if (hypotheses) *hypotheses = c->h_state->ransac_iters;
  return PLO_OK;
}

int plo_solve_wls_host(plo_ctx* c, const double* src, const double* ref, const double* nrm, const double* w, int64_t n,
                       double delta[16], int32_t* rank) {
  if (!c || !delta || n < 0 || (n > 0 && (!src || !ref || !nrm)))
    return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_solve_wls_host: bad argument");
  PLO_CUDA(c, cudaSetDevice(c->device));
  const size_t bytes = sizeof(double) * 3 * (size_t)std::max<int64_t>(n, 1);
  PLO_CUDA(c, c->h_src.reserve(bytes));
  PLO_CUDA(c, c->h_ref.reserve(bytes));
  PLO_CUDA(c, c->h_nrm.reserve(bytes));
  if (w) PLO_CUDA(c, c->h_w.reserve(sizeof(double) * (size_t)std::max<int64_t>(n, 1)));
  if (n > 0) {
    PLO_CUDA(c, cudaMemcpyAsync(c->h_src.p, src, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    PLO_CUDA(c, cudaMemcpyAsync(c->h_ref.p, ref, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    PLO_CUDA(c, cudaMemcpyAsync(c->h_nrm.p, nrm, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, c->stream));
    if (w) PLO_CUDA(c, cudaMemcpyAsync(c->h_w.p, w, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
  }
  PLO_TRY(plo_launch_reduce_solve_host_pairs(c, c->h_src.as<double>(), c->h_ref.as<double>(), c->h_nrm.as<double>(),
                                             w ? c->h_w.as<double>() : nullptr, n));
  PLO_TRY(fetch_state(c));
  memcpy(delta, c->h_state->delta, sizeof(double) * 16);
  if (rank) *rank = c->h_state->rank;
  return PLO_OK;
}

// ---- reference-shaped LS / RANSAC / DRPM: host vectors in, 4x4 out (include/solver.h:84-90, :100-114, :129-139) ----
//
// The device solvers work on the float32 pairs a projection leaves behind (what getXYZ / getNormals promote to
// double, include/common.h:51-75).  The host entry points stage caller pairs in the same arrays, so the coordinates
// must be float32-representable -- which the reference's own call site guarantees (src/laser_odometry.cpp:595-599).
namespace {
struct HostPairs {
  std::vector<float4> x, y, n;
  std::vector<int> status;
};

int stage_host_pairs(plo_ctx* c, const char* who, const double* src, const double* ref, const double* nrm, int64_t n, HostPairs& hp) {
  if (n < 0 || (n > 0 && (!src || !ref || !nrm))) return plo_fail(c, PLO_ERR_INVALID_ARG, std::string(who) + ": bad argument");
  if (n > (int64_t)1 << 30) return plo_fail(c, PLO_ERR_UNSUPPORTED, std::string(who) + ": more than 2^30 pairs");
  hp.x.resize((size_t)n); hp.y.resize((size_t)n); hp.n.resize((size_t)n); hp.status.assign((size_t)n, PLO_PT_OK);
  bool exact = true;
  auto f = [&exact](double v) { const float q = (float)v; exact = exact && ((double)q == v || v != v); return q; };
  for (int64_t i = 0; i < n; ++i) {
    float4 a = {f(src[3 * i]), f(src[3 * i + 1]), f(src[3 * i + 2]), 0.f};   // w = bits of PLO_PT_OK
    hp.x[(size_t)i] = a;
    hp.y[(size_t)i] = float4{f(ref[3 * i]), f(ref[3 * i + 1]), f(ref[3 * i + 2]), 0.f};
    hp.n[(size_t)i] = float4{f(nrm[3 * i]), f(nrm[3 * i + 1]), f(nrm[3 * i + 2]), 0.f};
  }
  if (!exact)
    return plo_fail(c, PLO_ERR_UNSUPPORTED, std::string(who) + ": coordinates must be float32-representable (the pairs of getXYZ / "
                    "getNormals, include/common.h:51-75); use plo_solve_wls_host for arbitrary doubles");
  return PLO_OK;
}

// runs `solver` on caller pairs through the projection's own arrays; the context's clouds stay, its last projection
// (pairs, temporal bounds, tiles) is gone afterwards
int solve_host_pairs(plo_ctx* c, const char* who, const double* src, const double* ref, const double* nrm, const double* w,
                     int64_t n, const plo_params& prm, bool drpm_only, double delta[16]) {
  if (!c || !delta) return plo_fail(c, PLO_ERR_INVALID_ARG, std::string(who) + ": NULL argument");
  HostPairs hp;
  PLO_TRY(stage_host_pairs(c, who, src, ref, nrm, n, hp));
  PLO_CUDA(c, cudaSetDevice(c->device));
  const plo_params saved_prm = c->prm;
  const DevParams saved_dprm = c->dprm;
  const int64_t saved_m = c->m_raw;
  c->prm = prm;
  flatten_params(c);
  c->m_raw = n;
  int rc = plo_reserve_query_buffers(c, false);
  if (rc == PLO_OK) rc = plo_reserve_solver_buffers(c);
  cudaError_t e = cudaSuccess;
  if (rc == PLO_OK) {
    e = c->counts_saved.reserve(sizeof(DevCounts));
    if (e == cudaSuccess) e = cudaMemcpyAsync(c->counts_saved.p, c->counts.p, sizeof(DevCounts), cudaMemcpyDeviceToDevice, c->stream);
    const int n32 = (int)n;
    if (e == cudaSuccess) e = cudaMemcpyAsync(&c->counts.as<DevCounts>()->n_source, &n32, sizeof(int), cudaMemcpyHostToDevice, c->stream);
    if (e == cudaSuccess && n > 0) {
      e = cudaMemcpyAsync(c->q_x.p, hp.x.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(c->q_y.p, hp.y.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(c->q_n.p, hp.n.data(), sizeof(float4) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess) e = cudaMemcpyAsync(c->q_status.p, hp.status.data(), sizeof(int) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
      if (e == cudaSuccess && w) {
        e = c->h_wext.reserve(sizeof(double) * (size_t)n);
        if (e == cudaSuccess) e = cudaMemcpyAsync(c->h_wext.p, w, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream);
      }
    }
    if (e != cudaSuccess) rc = plo_fail(c, PLO_ERR_CUDA, std::string(who) + ": " + cudaGetErrorString(e));
  }
  if (rc == PLO_OK) {
    c->host_drpm_only = drpm_only;
    c->host_w = (drpm_only && w && n > 0) ? c->h_wext.as<double>() : nullptr;
    rc = plo_launch_init_state(c, nullptr);
    if (rc == PLO_OK) rc = plo_launch_reduce_solve(c, false);
    c->host_drpm_only = false;
    c->host_w = nullptr;
  }
  // the staging copies above are from pageable host memory: done when they return.  Put the context's own counts back.
  if (c->counts_saved.p) cudaMemcpyAsync(c->counts.p, c->counts_saved.p, sizeof(DevCounts), cudaMemcpyDeviceToDevice, c->stream);
  if (rc == PLO_OK) rc = fetch_state(c);
  c->prm = saved_prm;
  c->dprm = saved_dprm;
  c->m_raw = saved_m;
  c->projected = false;
  c->prev_valid = false;
  c->hooks_valid = false;
  if (rc == PLO_OK) memcpy(delta, c->h_state->delta, sizeof(double) * 16);
  return rc;
}
}  // namespace

int plo_solve_ls_host(plo_ctx* c, const double* src, const double* ref, const double* nrm, int64_t n, double threshold,
                      double delta[16], int32_t* rank) {
  if (!c) return PLO_ERR_INVALID_ARG;
  if (!(threshold >= 0.0 && threshold < 0.5)) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_solve_ls_host: threshold must be in [0, 0.5)");
  plo_params p = c->prm;
  p.solver = PLO_SOLVER_LS;
  p.ls_threshold = threshold;
  PLO_TRY(solve_host_pairs(c, "plo_solve_ls_host", src, ref, nrm, nullptr, n, p, false, delta));
  if (rank) *rank = c->h_state->rank;
  return PLO_OK;
}

int plo_solve_ransac_host(plo_ctx* c, const double* src, const double* ref, const double* nrm, int64_t n, const plo_params* ransac,
                          double delta[16], double probs[6], int64_t* inliers, int32_t* hypotheses) {
  if (!c || !ransac) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_solve_ransac_host: NULL argument");
  plo_params p = *ransac;
  p.solver = PLO_SOLVER_RANSAC;
  if (p.ransac_final != PLO_FINAL_LS && p.ransac_final != PLO_FINAL_WLS && p.ransac_final != PLO_FINAL_DRPM)
    return plo_fail(c, PLO_ERR_UNSUPPORTED, "plo_solve_ransac_host: final_solve_method must be LS, Weighted LS or DRPM");
  if (p.ransac_final == PLO_FINAL_LS && !(p.ls_threshold >= 0.0 && p.ls_threshold < 0.5))
    return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_solve_ransac_host: ls_threshold must be in [0, 0.5)");
  PLO_TRY(solve_host_pairs(c, "plo_solve_ransac_host", src, ref, nrm, nullptr, n, p, false, delta));
  if (probs) memcpy(probs, c->h_state->probs, sizeof(double) * 6);
  if (inliers) *inliers = c->h_state->ransac_best;
  if (hypotheses) *hypotheses = c->h_state->ransac_iters;
  return PLO_OK;
}

int plo_solve_drpm_host(plo_ctx* c, const double* src, const double* ref, const double* nrm, const double* w, int64_t n,
                        double threshold, double stdev_points, double stdev_normals, double delta[16], double probs[6]) {
  if (!c) return PLO_ERR_INVALID_ARG;
  plo_params p = c->prm;
  p.solver = PLO_SOLVER_WLS;
  p.weight_mode = PLO_W_UNIT;
  p.drpm_threshold = threshold;
  p.drpm_stdev_points = stdev_points;
  p.drpm_stdev_normals = stdev_normals;
  PLO_TRY(solve_host_pairs(c, "plo_solve_drpm_host", src, ref, nrm, w, n, p, true, delta));
  if (probs) memcpy(probs, c->h_state->probs, sizeof(double) * 6);
  return PLO_OK;
}

int plo_get_normal_equations(plo_ctx* c, double H21[21], double g6[6], double* sw, double* swbb, int64_t* count) {
  if (!c) return PLO_ERR_INVALID_ARG;
  PLO_CUDA(c, cudaSetDevice(c->device));
  PLO_TRY(fetch_state(c));
  if (H21) memcpy(H21, c->h_state->H, sizeof(double) * 21);
  if (g6) memcpy(g6, c->h_state->g, sizeof(double) * 6);
  if (sw) *sw = c->h_state->sw;
  if (swbb) *swbb = c->h_state->swbb;
  if (count) *count = c->h_state->pairs;
  return PLO_OK;
}

// everything a captured iteration depends on; the loop graph is rebuilt when any of it changes
static std::vector<unsigned long long> loop_signature(const plo_ctx* c) {
  std::vector<unsigned long long> v;
  auto add = [&v](const void* p) { v.push_back((unsigned long long)(uintptr_t)p); };
  add(c->stream);
  add(c->pts_sorted.p); add(c->nrm_sorted.p); add(c->nrm_pca.p);
  for (int l = 0; l < PLO_MAX_LEVELS; ++l) { add(c->lvl_lo[l].p); add(c->lvl_hi[l].p); }
  add(c->s_p.p); add(c->s_n.p); add(c->q_x.p); add(c->q_y.p); add(c->q_n.p); add(c->q_status.p); add(c->q_kd2.p); add(c->q_tile_pts.p); add(c->q_tile_meta.p);
  add(c->partials.p); add(c->state.p); add(c->counts.p); add(c->sync_counters.p); add(c->miss_list.p); add(c->reduce_ticket.p);
  v.push_back((unsigned long long)c->tune_chunk); v.push_back(c->tune_fuse ? 1ull : 0ull); v.push_back((unsigned long long)c->tune_group);
  add(c->ls_keys[0].p); add(c->ls_keys[1].p); add(c->ls_vals[0].p); add(c->ls_vals[1].p); add(c->ls_hist.p); add(c->ls_tot.p); add(c->ls_mask.p);
  add(c->ransac_mind.p); add(c->partials2.p); add(c->h_src.p); add(c->h_ref.p); add(c->h_nrm.p); add(c->h_w.p); add(c->blockcnt.p);
  v.push_back((unsigned long long)c->n_levels);
  v.push_back((unsigned long long)c->n_raw_t);   // MapView.n_raw is a kernel argument
  v.push_back((unsigned long long)c->m_raw);     // launch geometry derives from it
  const unsigned char* pb = reinterpret_cast<const unsigned char*>(&c->dprm);
  for (size_t i = 0; i + 8 <= sizeof(DevParams); i += 8) { unsigned long long w; memcpy(&w, pb + i, 8); v.push_back(w); }
  return v;
}

static void destroy_loop_graph(plo_ctx* c) {
  if (c->loop_exec) cudaGraphExecDestroy(c->loop_exec);
  if (c->loop_graph) cudaGraphDestroy(c->loop_graph);
  c->loop_exec = nullptr;
  c->loop_graph = nullptr;
  c->loop_sig.clear();
}

// WHILE conditional node whose body is one ICP iteration: k_project,
// then k_reduce_solve, whose last block solves and sets the condition from the device-side state (weighted LS); for
// the other solvers the stand-alone reduce / solve kernels of p2plane_solve.cu follow and the last of them sets it.
static int build_loop_graph(plo_ctx* c) {
  destroy_loop_graph(c);
  cudaGraph_t g = nullptr;
  if (cudaGraphCreate(&g, 0) != cudaSuccess) return -1;
  cudaGraphConditionalHandle handle;
  if (cudaGraphConditionalHandleCreate(&handle, g, 1, cudaGraphCondAssignDefault) != cudaSuccess) { cudaGraphDestroy(g); return -1; }
  cudaGraphNodeParams np = {};
  np.type = cudaGraphNodeTypeConditional;
  np.conditional.handle = handle;
  np.conditional.type = cudaGraphCondTypeWhile;
  np.conditional.size = 1;
  cudaGraphNode_t node;
  if (cudaGraphAddNode(&node, g, nullptr, 0, &np) != cudaSuccess) { cudaGraphDestroy(g); return -1; }
  cudaGraph_t body = np.conditional.phGraph_out[0];
  if (cudaStreamBeginCaptureToGraph(c->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
    cudaGraphDestroy(g);
    return -1;
  }
  const int64_t launches_before = c->launches;
  int rc = plo_launch_project(c, false);
  if (rc == PLO_OK) rc = plo_launch_reduce_solve(c, true, (unsigned long long)handle);
  c->body_launches = (int)(c->launches - launches_before);   // kernels per loop iteration
  c->launches = launches_before;   // capture is not execution
  cudaGraph_t captured = nullptr;
  const cudaError_t e = cudaStreamEndCapture(c->stream, &captured);
  if (rc != PLO_OK || e != cudaSuccess) { cudaGraphDestroy(g); cudaGetLastError(); return -1; }
  cudaGraphExec_t exec = nullptr;
  if (cudaGraphInstantiate(&exec, g, 0) != cudaSuccess) { cudaGraphDestroy(g); cudaGetLastError(); return -1; }
  c->loop_graph = g;
  c->loop_exec = exec;
  c->loop_sig = loop_signature(c);
  return 0;
}

static int enqueue_register(plo_ctx* c, const double* T0) {
  PLO_TRY(plo_reserve_query_buffers(c, false));
  PLO_TRY(plo_reserve_solver_buffers(c));
  if (c->dprm.use_pca_normals) PLO_TRY(plo_launch_pca_normals(c));
  PLO_TRY(plo_launch_init_state(c, T0));
  if (c->prm.iterations <= 0) { c->hooks_valid = false; return PLO_OK; }
  // Product path: the loop is a CUDA graph with a device-evaluated WHILE condition — exactly the
  // iterations that are needed run, with no host round trip.  With per-launch profiling (or if the
  // driver refuses conditional nodes) every iteration is enqueued up front instead; kernels of
  // iterations after convergence then see the device-side `done` flag and return at once.
  // PLO_NO_GRAPH=1: enqueue-all path (ncu cannot profile kernel nodes of graphs with conditional nodes)
  // Product path for the weighted-LS solver: the whole loop is ONE cooperative launch (k_register_loop), grid barriers
  // between the phases of an iteration, no launch / graph node / host between iterations.
  if (!c->profiling && !c->tune_no_graph && c->tune_loop_kernel && c->tune_fuse && c->dprm.solver == PLO_SOLVER_WLS && c->m_raw > 0 &&
      c->coop_ok) {
    PLO_TRY(plo_launch_register_loop(c));
    c->projected = true;
    c->hooks_valid = false;
    c->graph_launched = false;
    c->loop_kernel_launched = true;
    return PLO_OK;
  }
  c->loop_kernel_launched = false;
  if (!c->profiling && c->graph_ok && !c->tune_no_graph && c->m_raw > 0) {
    if (!c->loop_exec || c->loop_sig != loop_signature(c)) {
      if (build_loop_graph(c) != 0) c->graph_ok = false;
    }
    if (c->loop_exec) {
      const bool prev = c->prev_valid;
      (void)prev;
      PLO_CUDA(c, cudaGraphLaunch(c->loop_exec, c->stream));
      c->prev_valid = true;
      c->projected = true;
      c->hooks_valid = false;
      c->graph_launched = true;
      return PLO_OK;
    }
  }
  c->graph_launched = false;
  if (c->profiling) {
    while ((int)c->ev_proj.size() < 2 * c->prm.iterations) {
      cudaEvent_t e;
      PLO_CUDA(c, cudaEventCreate(&e));
      c->ev_proj.push_back(e);
    }
  }
  for (int it = 0; it < c->prm.iterations; ++it) {
    if (c->profiling) cudaEventRecord(c->ev_proj[2 * it], c->stream);
    PLO_TRY(plo_launch_project(c, false));
    if (c->profiling) cudaEventRecord(c->ev_proj[2 * it + 1], c->stream);
    PLO_TRY(plo_launch_reduce_solve(c, true));
  }
  c->projected = true;
  c->hooks_valid = false;
  return PLO_OK;
}

// kernels a graph-launched loop executed (the host did not enqueue them one by one)
static void count_graph_launches(plo_ctx* c, const DevState* s) {
  if (!c->graph_launched) return;
  const int bodies = s->iters + (s->status == PLO_REG_TOO_FEW_PAIRS ? 1 : 0);
  c->launches += (int64_t)c->body_launches * bodies;
  c->graph_launched = false;
}

static void fill_reg_stats(const DevState* s, plo_reg_stats* st, int iterations) {
  st->status = s->status ? s->status : (iterations == 0 ? PLO_REG_MAX_ITERS : s->status);
  st->iters = s->iters;
  st->pairs = s->pairs;
  st->rms = s->rms;
  for (int i = 0; i < 6; ++i) st->dropped[i] = s->dropped[i];
  st->delta_dist = s->delta_dist;
  st->delta_angle = s->delta_angle;
  st->rank = s->rank;
  st->reserved = 0;
}

int plo_register(plo_ctx* c, const double T0[16], double T[16], plo_reg_stats* stats) {
  if (!c || !T) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_register: NULL argument");
  PLO_TRY(require_clouds(c, "plo_register"));
  PLO_CUDA(c, cudaSetDevice(c->device));
  cudaEventRecord(c->ev[2], c->stream);
  PLO_TRY(enqueue_register(c, T0));
  cudaEventRecord(c->ev[3], c->stream);
  c->ev_reg_pending = true;
  PLO_TRY(fetch_state(c));
  count_graph_launches(c, c->h_state);
  memcpy(T, c->h_state->rPose, sizeof(double) * 16);
  if (stats) fill_reg_stats(c->h_state, stats, c->prm.iterations);
  if (c->profiling) {
    // launches that did work: one per solve, plus the one that found too few pairs
    int n = c->h_state->iters + (c->h_state->status == PLO_REG_TOO_FEW_PAIRS ? 1 : 0);
    n = std::min(n, c->prm.iterations);
    float total = 0.f;
    for (int it = 0; it < n; ++it) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, c->ev_proj[2 * it], c->ev_proj[2 * it + 1]);
      total += ms;
      if (it < 64) c->ms_project_each[it] = ms;
    }
    c->n_project = n;
    c->ms_project_mean = n > 0 ? total / n : 0.f;
  }
  return PLO_OK;
}

int plo_set_profiling(plo_ctx* c, int32_t enabled) {
  if (!c) return PLO_ERR_INVALID_ARG;
  c->profiling = enabled != 0;
  return PLO_OK;
}

int plo_last_kernel_timings(plo_ctx* c, float* ms_project_mean, int32_t* n_project) {
  if (!c) return PLO_ERR_INVALID_ARG;
  if (ms_project_mean) *ms_project_mean = c->ms_project_mean;
  if (n_project) *n_project = c->n_project;
  return PLO_OK;
}

int plo_last_project_times(plo_ctx* c, float* ms_each, int32_t cap, int32_t* n_project) {
  if (!c) return PLO_ERR_INVALID_ARG;
  const int n = std::min(c->n_project, 64);
  if (ms_each)
    for (int i = 0; i < n && i < cap; ++i) ms_each[i] = c->ms_project_each[i];
  if (n_project) *n_project = n;
  return PLO_OK;
}

// copies the loop state of one finished registration into a device-side result slot
__global__ void k_store_result(const DevState* __restrict__ st, DevState* __restrict__ slot) {
  if (threadIdx.x == 0) *slot = *st;
}

int plo_register_batch(plo_ctx* c, int32_t count, const void* const* sources, const int64_t* n_src,
                       const void* const* targets, const int64_t* n_tgt, int32_t stride, int32_t on_device,
                       double* T_out, plo_reg_stats* stats_out) {
  if (!c || count < 0 || (count > 0 && (!sources || !n_src || !targets || !n_tgt || !T_out)))
    return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_register_batch: bad argument");
  if (count == 0) return PLO_OK;
  PLO_CUDA(c, cudaSetDevice(c->device));
  // result slots (device + pinned host) live in the context and only ever grow: no allocation in a timed region
  PLO_CUDA(c, c->batch_slots.reserve(sizeof(DevState) * (size_t)count));
  if ((size_t)count > c->h_batch_cap) {
    if (c->h_batch_slots) cudaFreeHost(c->h_batch_slots);
    c->h_batch_slots = nullptr;
    c->h_batch_cap = 0;
    const size_t want = std::max<size_t>((size_t)count + (size_t)count / 2, 64);
    PLO_CUDA(c, cudaMallocHost(&c->h_batch_slots, sizeof(DevState) * want));
    c->h_batch_cap = want;
  }
  DevBuf& slots = c->batch_slots;
  DevState* h_slots = c->h_batch_slots;
  int rc = PLO_OK;
  int graph_units = 0;
  // Host inputs: double-buffered staging on a second stream, so that the upload of pair i+1 overlaps
  // the registration of pair i (pinned host memory makes the copies truly asynchronous).
  DevBuf* stage_t[2] = {&c->t_stage, &c->t_stage2};
  DevBuf* stage_s[2] = {&c->s_stage, &c->s_stage2};
  if (!on_device) {
    if (stride < 28 || (stride % 4) != 0) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_register_batch: bad stride");
    int64_t max_t = 1, max_s = 1;
    for (int i = 0; i < count; ++i) {
      if (n_tgt[i] < 0 || n_src[i] < 0 || (n_tgt[i] > 0 && !targets[i]) || (n_src[i] > 0 && !sources[i]))
        return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_register_batch: bad pointer / count");
      max_t = std::max(max_t, n_tgt[i]);
      max_s = std::max(max_s, n_src[i]);
    }
    PLO_TRY(ensure_copy_stream(c));
    // earlier work on either stream may still use the staging buffers: drain the copy stream's view of them first
    PLO_CUDA(c, cudaStreamSynchronize(c->copy_stream));
    for (int b = 0; b < 2; ++b) {
      PLO_CUDA(c, stage_t[b]->reserve((size_t)max_t * stride));
      PLO_CUDA(c, stage_s[b]->reserve((size_t)max_s * stride));
    }
    c->stage_used[0] = c->stage_used[1] = false;   // this call orders the buffers with its own events
    PLO_CUDA(c, cudaEventRecord(c->ev_batch_start, c->stream));
    PLO_CUDA(c, cudaStreamWaitEvent(c->copy_stream, c->ev_batch_start, 0));
  }
  for (int i = 0; i < count && rc == PLO_OK; ++i) {
    if (on_device) {
      rc = plo_set_target_device(c, targets[i], n_tgt[i], stride);
      if (rc == PLO_OK) rc = plo_set_source_device(c, sources[i], n_src[i], stride);
    } else {
      const int b = i & 1;
      cudaError_t e = cudaSuccess;
      if (i >= 2) e = cudaStreamWaitEvent(c->copy_stream, c->ev_consumed[b], 0);   // pair i-2 has unpacked this buffer
      if (e == cudaSuccess && n_tgt[i] > 0) e = cudaMemcpyAsync(stage_t[b]->p, targets[i], (size_t)n_tgt[i] * stride, cudaMemcpyHostToDevice, c->copy_stream);
      if (e == cudaSuccess && n_src[i] > 0) e = cudaMemcpyAsync(stage_s[b]->p, sources[i], (size_t)n_src[i] * stride, cudaMemcpyHostToDevice, c->copy_stream);
      if (e == cudaSuccess) e = cudaEventRecord(c->ev_copied[b], c->copy_stream);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(c->stream, c->ev_copied[b], 0);
      if (e != cudaSuccess) { rc = plo_fail(c, PLO_ERR_CUDA, std::string("plo_register_batch: ") + cudaGetErrorString(e)); break; }
      rc = plo_build_index(c, stage_t[b]->p, n_tgt[i], stride);
      if (rc == PLO_OK) rc = plo_upload_source(c, stage_s[b]->p, n_src[i], stride);
      if (rc == PLO_OK && cudaEventRecord(c->ev_consumed[b], c->stream) != cudaSuccess) rc = plo_fail(c, PLO_ERR_CUDA, "plo_register_batch: event record failed");
    }
    if (rc == PLO_OK) rc = enqueue_register(c, nullptr);
    if (rc == PLO_OK) {
      if (c->graph_launched) { graph_units++; c->graph_launched = false; }
      k_store_result<<<1, 32, 0, c->stream>>>(c->state.as<DevState>(), slots.as<DevState>() + i);
      c->launches++;
      if (cudaGetLastError() != cudaSuccess) rc = plo_fail(c, PLO_ERR_CUDA, "plo_register_batch: k_store_result launch failed");
    }
  }
  if (rc == PLO_OK) {
    cudaError_t e = cudaMemcpyAsync(h_slots, slots.p, sizeof(DevState) * (size_t)count, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) rc = plo_fail(c, PLO_ERR_CUDA, std::string("plo_register_batch: ") + cudaGetErrorString(e));
  } else {
    cudaStreamSynchronize(c->stream);
  }
  if (!on_device && c->copy_stream) cudaStreamSynchronize(c->copy_stream);
  if (rc == PLO_OK) {
    for (int i = 0; i < count; ++i) {
      if (graph_units > 0) c->launches += (int64_t)c->body_launches * (h_slots[i].iters + (h_slots[i].status == PLO_REG_TOO_FEW_PAIRS ? 1 : 0));
      memcpy(T_out + 16 * (size_t)i, h_slots[i].rPose, sizeof(double) * 16);
      if (stats_out) fill_reg_stats(&h_slots[i], &stats_out[i], c->prm.iterations);
    }
  }
  return rc;
}

int64_t plo_launch_count(const plo_ctx* c) { return c ? c->launches : 0; }

int plo_last_timings(plo_ctx* c, float* ms_index_build, float* ms_register) {
  if (!c) return PLO_ERR_INVALID_ARG;
  PLO_CUDA(c, cudaSetDevice(c->device));
  PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->ev_index_pending) { cudaEventElapsedTime(&c->ms_index, c->ev[0], c->ev[1]); c->ev_index_pending = false; }
  if (c->ev_reg_pending) { cudaEventElapsedTime(&c->ms_register, c->ev[2], c->ev[3]); c->ev_reg_pending = false; }
  if (ms_index_build) *ms_index_build = c->ms_index;
  if (ms_register) *ms_register = c->ms_register;
  return PLO_OK;
}

int plo_last_tile_misses(plo_ctx* c, int32_t* misses, int32_t cap, int32_t* n_project) {
  if (!c) return PLO_ERR_INVALID_ARG;
  const int n = std::min(c->h_state->iters + (c->h_state->status == PLO_REG_TOO_FEW_PAIRS ? 1 : 0), 32);
  if (misses)
    for (int i = 0; i < n && i < cap; ++i) misses[i] = c->h_state->miss_hist[i];
  if (n_project) *n_project = n;
  return PLO_OK;
}

// PLO_LOOP_TIMING builds only: globaltimer stamps of block 0 at the phase boundaries of k_register_loop (ns, 7 per iteration)
int plo_debug_loop_stamps(plo_ctx* c, unsigned long long* out128) {
  if (!c || !out128) return PLO_ERR_INVALID_ARG;
  PLO_CUDA(c, cudaMemcpy(out128, c->partials.as<double>() + (size_t)plo_loop_blocks(c) * PLO_NSUM, 8 * 8 * 16, cudaMemcpyDeviceToHost));
  return PLO_OK;
}

int plo_time_project_kernel(plo_ctx* c, const double T[16], int32_t reps, float* ms_mean) {
  if (!c || !T || !ms_mean || reps < 1) return plo_fail(c, PLO_ERR_INVALID_ARG, "plo_time_project_kernel: bad argument");
  PLO_TRY(require_clouds(c, "plo_time_project_kernel"));
  PLO_CUDA(c, cudaSetDevice(c->device));
  PLO_TRY(plo_reserve_query_buffers(c, false));
  if (c->dprm.use_pca_normals) PLO_TRY(plo_launch_pca_normals(c));
  PLO_TRY(plo_launch_init_state(c, T));
  PLO_TRY(plo_launch_project(c, false));   // warm-up; leaves the k-th distances for the temporal bound
  PLO_TRY(plo_launch_init_state(c, T));
  cudaEvent_t a, b;
  PLO_CUDA(c, cudaEventCreate(&a));
  PLO_CUDA(c, cudaEventCreate(&b));
  cudaEventRecord(a, c->stream);
  for (int i = 0; i < reps; ++i) PLO_TRY(plo_launch_project(c, false));
  cudaEventRecord(b, c->stream);
  PLO_CUDA(c, cudaStreamSynchronize(c->stream));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, a, b);
  cudaEventDestroy(a);
  cudaEventDestroy(b);
  *ms_mean = ms / reps;
  c->projected = true;
  c->hooks_valid = false;
  return PLO_OK;
}

}  // extern "C"
