// capi.cu — the extern "C" boundary (include/glio_b200.h) and the context that owns device state.
// No CPU fallback: without a CUDA device glio_create fails.
#include <array>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#if defined(__linux__)
#include <pthread.h>
#include <sched.h>
#endif
#include <chrono>
#include <deque>
#include <map>
#include <memory>

#include "common.cuh"
#include "hostmath.h"
#include "solver.h"
#include "marg.h"

using namespace glio;

namespace {

struct Slot {
  int64_t Q = 0;
  int stride = 3;
  DevBuf<float> scan;            // device copy of the scan (stride floats per point)
  const float* scan_ptr = nullptr;  // == scan.p, or a caller-owned device pointer (GLIO_DEVICE input)
  DevBuf<float4> m_cpw, m_nsd;   // compacted matches, scan order
  DevBuf<int32_t> m_src;
  int64_t n_match = 0;
  DevBuf<float4> s_cpw, s_nsd;   // selected subset
  int64_t n_sel = -1;            // -1: no selection -> all matches are active
  DevBuf<float4> e_cps, e_pa, e_pb;  // edge correspondences (caller-provided)
  int64_t n_edge = 0;
  // debug (keep_debug)
  DevBuf<uint8_t> dbg_status;
  DevBuf<int32_t> dbg_idx5;
  DevBuf<float> dbg_sqd5;
  DevBuf<float4> dbg_pm;
  DevBuf<double> dbg_plane;
  int64_t dbg_Q = 0;
  void release() {
    scan.release(); m_cpw.release(); m_nsd.release(); m_src.release(); s_cpw.release(); s_nsd.release(); e_cps.release(); e_pa.release(); e_pb.release();
    dbg_status.release(); dbg_idx5.release(); dbg_sqd5.release(); dbg_pm.release(); dbg_plane.release();
  }
};

}  // namespace

// One marginalisation whose host half (A/b assembly, Schur, eigen-decomposition, prior) runs on the context's worker thread while
// the caller goes on (typically: hands the next window's association to the GPU).  The device half was queued by the caller.
struct glio_marg_job {
  glio_ctx* ctx = nullptr;
  int W = 0;
  double eps = 1e-8;
  unsigned int epoch = 0;
  int submit_cpu = -1;                 // CPU the submitting thread ran on (the worker keeps off it)
  std::vector<double> A, b;            // host-side accumulations (IMU / previous prior), marginalisation ordering
  std::vector<double> x0_pose, x0_sb;  // linearisation point of the kept states
  std::vector<double> out;             // snapshot of the device result (W x 28)
  std::atomic<int> copied{0};          // the worker no longer needs the context's result buffer
  std::atomic<int> done{0};
  int rc = 0;
  std::string err;
  glio_marg_prior* prior = nullptr;
};

struct glio_ctx {
  int device = 0;
  // worker thread of glio_window_marginalize_async (started lazily, joined by glio_destroy)
  std::thread marg_thread;
  std::mutex marg_mu;
  std::condition_variable marg_cv;
  glio_marg_job* marg_queued = nullptr;      // handed to the worker (guarded by marg_mu)
  bool marg_quit = false;
  glio_marg_job* marg_pending = nullptr;     // the job in flight, caller-thread view (one at a time)
  cudaStream_t st = nullptr;
  glio_params prm{};
  std::string err;
  LaunchCounter lc;

  int eval_slots = 296;        // resident evaluation blocks: 2 per SM (set from the device at creation)
  float pts_per_cell = 8.0f;   // target points per occupied grid cell (tuning hook: env GLIO_PTS_PER_CELL)
  GridBuild map;
  bool has_map = false;
  DevBuf<float> map_stage;
  // map prefetch (glio_map_prefetch): the next window's local map travels on the copy stream while the current window is marginalised
  DevBuf<float> map_stage2; const float* pf_ptr = nullptr; int64_t pf_M = 0; int pf_stride = 0; cudaEvent_t ev_map = nullptr;

  std::vector<std::unique_ptr<Slot>> slots;

  // ---- batch (scan-to-multiscan) state
  struct Frame {
    int64_t Q = 0; int stride = 3;
    DevBuf<float> scan; const float* scan_ptr = nullptr;
    double pose[7] = {0, 0, 0, 1, 0, 0, 0};
    GridBuild grid; bool grid_valid = false;
  };
  struct Pair {
    int cur = 0, oth = 0;
    DevBuf<float4> m_cpw; DevBuf<double> m_nc; DevBuf<int32_t> m_src; int64_t n_match = 0;
    DevBuf<float4> s_cpw; DevBuf<double> s_nc; int64_t n_sel = -1;
  };
  std::map<int, std::unique_ptr<Frame>> frames;
  std::vector<std::unique_ptr<Pair>> pairs;
  std::map<std::pair<int, int>, int> pair_index;
  bool bin_dirty = true; int bin_K = -1; int n_bin_items = 0;
  glio_allreduce_fn allreduce = nullptr; void* allreduce_user = nullptr;
  DevBuf<BinItem> d_bin_items; DevBuf<int> d_pair_item_start, d_kf_inc_start; DevBuf<BinIncidence> d_inc;
  DevBuf<double> d_bin_partials, d_pair_sums, d_bin_diag, d_bin_off, d_bin_poses;
  PinnedBuf<double> h_bin_diag, h_bin_off, h_bin_poses;

  // association workspace
  DevBuf<float4> w_pm, w_pmq, w_nsd;
  DevBuf<uint16_t> w_seg;
  DevBuf<uint8_t> w_status;
  DevBuf<float> w_weight;
  DevBuf<double> w_nc, w_plane;
  DevBuf<int32_t> w_idx5;
  DevBuf<float> w_sqd5;
  DevBuf<int32_t> w_knn_idx;
  DevBuf<uint32_t> w_deferred;
  int tile_rings = 64;         // tuning hook: env GLIO_TILE_RINGS (>= 32: the tile pass scans all rings itself)
  int grow_mode = 1;           // growth slabs of the box search: 1 = batched row bounds + row culling, 0 = row after row (env GLIO_KNN_GROW)
  int knn_mode = 2;            // 2: per-thread box growth (default, fastest measured); 3: same from the own cell; 1: per-thread ring growth; 0: warp-cooperative tile pass; 4-7: the measured alternatives of knn_tile.cu (env GLIO_KNN_MODE)
  DevBuf<float> w_knn_sqd;
  DevBuf<unsigned long long> d_stats;
  DevBuf<int> w_flags, w_pos, cell_count, cell_pos, scan_tmp;
  DevBuf<SegDesc> d_segs;
  DevBuf<CompactDst> d_dst;
  DevBuf<int> d_counts;
  PinnedBuf<int> h_counts;
  DevBuf<int> d_bad;
  DevBuf<int32_t> d_keep;
  // front-end feature extraction work space
  DevBuf<float> f_cloud, f_curv; DevBuf<int8_t> f_label, f_picked; DevBuf<int32_t> f_scan, f_ring, f_lf, f_counts, f_out; DevBuf<float4> f_ds, f_outds;
  // K2e scratch (kept in the context: no cudaMalloc / cudaFree per call, nothing to leak when a call throws)
  DevBuf<EdgeItem> e_items; DevBuf<int> e_start; DevBuf<double> e_part, e_out, e_poses;

  // evaluation
  DevBuf<EvalItem> d_items;
  DevBuf<int> d_kf_item_start;
  DevBuf<double> d_partials, d_out, d_poses, d_r, d_J;
  PinnedBuf<double> h_out, h_poses;
  // local map maintenance: world-frame keyframe clouds in deque order, concatenation, voxel filter work space
  struct LmFrame { DevBuf<float4> pts; int64_t n = 0; };
  std::deque<std::unique_ptr<LmFrame>> lm_frames;
  std::vector<std::unique_ptr<LmFrame>> lm_spare;      // buffers of popped frames, reused by the next push (no cudaFree / cudaMalloc per keyframe)
  DevBuf<float> lm_stage; DevBuf<float4> lm_cat; VoxelWork lm_vox; int64_t map_n = 0;
  cudaStream_t st_copy = nullptr;       // window scans travel on their own stream so the upload of the next window can overlap a running solve
  cudaEvent_t ev_scans = nullptr, ev_main = nullptr;
  bool scans_pending = false;
  PinnedBuf<unsigned int> h_flag;      // completion epoch written by the last block of k_eval_unary
  unsigned int eval_epoch = 0;
  DevBuf<unsigned int> d_ticket;
  bool items_dirty = true;
  int items_W = -1;
  int n_items = 0;

  Slot& slot(int k) {
    GLIO_REQUIRE(k >= 0 && k < 4096, GLIO_ERR_ARG, "slot index out of range");
    if ((size_t)k >= slots.size()) slots.resize((size_t)k + 1);
    if (!slots[k]) slots[k].reset(new Slot());
    return *slots[k];
  }
};

namespace {

// GLIO_TRACE=1: wall-clock marks of the association call (each mark synchronises the stream first)
struct Trace {
  bool on; cudaStream_t st; std::chrono::steady_clock::time_point t0; const char* what;
  Trace(cudaStream_t s, const char* w) : on(getenv("GLIO_TRACE") != nullptr), st(s), t0(std::chrono::steady_clock::now()), what(w) {}
  void mark(const char* label) { if (!on) return; cudaStreamSynchronize(st); fprintf(stderr, "[glio trace] %s %-18s %8.1f us\n", what, label, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count()); }
};

thread_local std::string tl_err;

template <class F>
int guarded(glio_ctx* ctx, F&& f) {
  try {
    if (ctx) GLIO_CUDA_TRY(cudaSetDevice(ctx->device));
    f();
    return GLIO_OK;
  } catch (const Error& e) {
    if (ctx) ctx->err = e.msg;
    tl_err = e.msg; set_global_error(e.msg);
    return e.code;
  } catch (const std::exception& e) {
    if (ctx) ctx->err = e.what();
    tl_err = e.what(); set_global_error(e.what());
    return GLIO_ERR_STATE;
  }
}

// bring `n` points (stride floats each) to the device; returns the device pointer to read from
const float* stage_points(glio_ctx* c, DevBuf<float>& buf, const float* xyz, int64_t n, int stride, int mem) {
  GLIO_REQUIRE(xyz != nullptr && n > 0, GLIO_ERR_ARG, "null or empty point array");
  GLIO_REQUIRE(stride >= 3 && stride <= 64, GLIO_ERR_ARG, "stride_floats must be in [3,64]");
  if (mem == GLIO_DEVICE) return xyz;
  buf.reserve((size_t)n * stride);
  GLIO_CUDA_TRY(cudaMemcpyAsync(buf.p, xyz, (size_t)n * stride * sizeof(float), cudaMemcpyHostToDevice, c->st));
  return buf.p;
}

void ensure_work(glio_ctx* c, int64_t Qt, bool pair) {
  c->w_pm.reserve((size_t)Qt); c->w_pmq.reserve((size_t)Qt); c->w_seg.reserve((size_t)Qt);
  c->w_status.reserve((size_t)Qt); c->w_weight.reserve((size_t)Qt);
  c->w_flags.reserve((size_t)Qt + 1); c->w_pos.reserve((size_t)Qt + 1);
  c->w_knn_idx.reserve((size_t)Qt * 5); c->w_knn_sqd.reserve((size_t)Qt * 5); c->w_deferred.reserve((size_t)2 * Qt);
  if (!c->d_stats.p) { c->d_stats.reserve(4); GLIO_CUDA_TRY(cudaMemsetAsync(c->d_stats.p, 0, 4 * sizeof(unsigned long long), c->st)); }
  if (pair) c->w_nc.reserve((size_t)Qt * 6); else c->w_nsd.reserve((size_t)Qt);
  if (c->prm.keep_debug) { c->w_idx5.reserve((size_t)Qt * 5); c->w_sqd5.reserve((size_t)Qt * 5); c->w_plane.reserve((size_t)Qt * 4); }
}

AssocWork make_work(glio_ctx* c, int64_t Qt, bool pair) {
  AssocWork w{};
  w.Qt = Qt; w.pm = c->w_pm.p; w.pmq = c->w_pmq.p; w.seg = c->w_seg.p; w.status = c->w_status.p;
  w.knn_idx = c->w_knn_idx.p; w.knn_sqd = c->w_knn_sqd.p; w.n_fallback = c->d_stats.p;
  w.knn_mode = c->knn_mode; w.tile_rings = c->tile_rings; w.grow_mode = c->grow_mode; w.deferred = c->w_deferred.p; w.n_deferred = reinterpret_cast<unsigned int*>(c->d_stats.p + 2);
  w.nsd = pair ? nullptr : c->w_nsd.p; w.weight = c->w_weight.p; w.normal_cent = pair ? c->w_nc.p : nullptr;
  if (c->prm.keep_debug) { w.idx5 = c->w_idx5.p; w.sqd5 = c->w_sqd5.p; w.plane = c->w_plane.p; }
  return w;
}

// run scan-to-map association for the given slots (one launch), compact, fetch counts
void associate_slots(glio_ctx* c, const std::vector<int>& ids, const std::vector<std::array<double, 7>>& lidar_poses,
                     int64_t* n_match) {
  GLIO_REQUIRE(c->has_map, GLIO_ERR_STATE, "glio_set_map must be called before association");
  if (c->scans_pending) { GLIO_CUDA_TRY(cudaStreamWaitEvent(c->st, c->ev_scans, 0)); c->scans_pending = false; }
  Trace tr(c->st, "associate");
  const int nseg = (int)ids.size();
  std::vector<SegDesc> segs(nseg);
  std::vector<CompactDst> dst(nseg);
  int64_t Qt = 0;
  for (int s = 0; s < nseg; ++s) {
    Slot& sl = c->slot(ids[s]);
    GLIO_REQUIRE(sl.Q > 0 && sl.scan_ptr, GLIO_ERR_STATE, "slot has no scan");
    segs[s].src = sl.scan_ptr; segs[s].stride = sl.stride; segs[s].count = sl.Q; segs[s].offset = Qt;
    for (int k = 0; k < 3; ++k) segs[s].t[k] = lidar_poses[s][k];
    for (int k = 0; k < 4; ++k) segs[s].q[k] = lidar_poses[s][3 + k];
    sl.m_cpw.reserve((size_t)sl.Q); sl.m_nsd.reserve((size_t)sl.Q); sl.m_src.reserve((size_t)sl.Q);
    dst[s].cpw = sl.m_cpw.p; dst[s].nsd = sl.m_nsd.p; dst[s].nc = nullptr; dst[s].src = sl.m_src.p;
    sl.n_sel = -1;
    Qt += sl.Q;
  }
  ensure_work(c, Qt, false);
  c->d_segs.reserve(nseg); c->d_dst.reserve(nseg); c->d_counts.reserve(nseg); c->h_counts.reserve(nseg);
  GLIO_CUDA_TRY(cudaMemcpyAsync(c->d_segs.p, segs.data(), nseg * sizeof(SegDesc), cudaMemcpyHostToDevice, c->st));
  GLIO_CUDA_TRY(cudaMemcpyAsync(c->d_dst.p, dst.data(), nseg * sizeof(CompactDst), cudaMemcpyHostToDevice, c->st));
  AssocWork w = make_work(c, Qt, false);
  AssocGates gates{c->prm.kd_max_radius, c->prm.surf_dist_thres, c->prm.weight_min};
  tr.mark("setup+h2d");
  assoc_run(c->map, c->d_segs.p, nseg, w, gates, nullptr, 0, c->cell_count, c->cell_pos, c->scan_tmp, c->st, c->lc);
  tr.mark("assoc_run");
  compact_run(w, c->d_segs.p, nseg, c->w_flags.p, c->w_pos.p, c->scan_tmp, c->d_dst.p, c->d_counts.p, c->st, c->lc);
  tr.mark("compact");
  GLIO_CUDA_TRY(cudaMemcpyAsync(c->h_counts.p, c->d_counts.p, nseg * sizeof(int), cudaMemcpyDeviceToHost, c->st));
  if (c->prm.keep_debug) {
    for (int s = 0; s < nseg; ++s) {
      Slot& sl = c->slot(ids[s]);
      const int64_t o = segs[s].offset, q = sl.Q;
      sl.dbg_status.reserve(q); sl.dbg_idx5.reserve(5 * q); sl.dbg_sqd5.reserve(5 * q); sl.dbg_pm.reserve(q); sl.dbg_plane.reserve(4 * q);
      GLIO_CUDA_TRY(cudaMemcpyAsync(sl.dbg_status.p, w.status + o, q, cudaMemcpyDeviceToDevice, c->st));
      GLIO_CUDA_TRY(cudaMemcpyAsync(sl.dbg_idx5.p, w.idx5 + 5 * o, 5 * q * sizeof(int32_t), cudaMemcpyDeviceToDevice, c->st));
      GLIO_CUDA_TRY(cudaMemcpyAsync(sl.dbg_sqd5.p, w.sqd5 + 5 * o, 5 * q * sizeof(float), cudaMemcpyDeviceToDevice, c->st));
      GLIO_CUDA_TRY(cudaMemcpyAsync(sl.dbg_pm.p, w.pm + o, q * sizeof(float4), cudaMemcpyDeviceToDevice, c->st));
      GLIO_CUDA_TRY(cudaMemcpyAsync(sl.dbg_plane.p, w.plane + 4 * o, 4 * q * sizeof(double), cudaMemcpyDeviceToDevice, c->st));
      sl.dbg_Q = q;
    }
  }
  GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
  for (int s = 0; s < nseg; ++s) {
    Slot& sl = c->slot(ids[s]);
    sl.n_match = c->h_counts.p[s];
    if (n_match) n_match[s] = sl.n_match;
  }
  c->items_dirty = true;
  tr.mark("done");
}

void build_items(glio_ctx* c, int W) {
  if (!c->items_dirty && c->items_W == W) return;
  std::vector<EvalItem> items;
  std::vector<int> start(W + 1, 0);
  // item size: the whole launch should be about one wave of 2 resident blocks per SM (less per-item reduction
  // overhead, no wave-quantisation tail), but never below GLIO_ITEM_MAX residuals per item
  int64_t n_total = 0;
  for (int k = 0; k < W && (size_t)k < c->slots.size(); ++k) if (c->slots[k]) n_total += c->slots[k]->n_sel >= 0 ? c->slots[k]->n_sel : c->slots[k]->n_match;
  int64_t item_size = (n_total + c->eval_slots - 1) / std::max(c->eval_slots, 1);
  item_size = std::max<int64_t>(GLIO_ITEM_MAX, ((item_size + 255) / 256) * 256);
  for (int k = 0; k < W; ++k) {
    start[k] = (int)items.size();
    if ((size_t)k >= c->slots.size() || !c->slots[k]) continue;
    Slot& sl = *c->slots[k];
    const bool sel = sl.n_sel >= 0;
    const int64_t n = sel ? sl.n_sel : sl.n_match;
    const float4* cpw = sel ? sl.s_cpw.p : sl.m_cpw.p;
    const float4* nsd = sel ? sl.s_nsd.p : sl.m_nsd.p;
    const int64_t parts = std::max<int64_t>(1, (n + item_size - 1) / item_size);
    const int64_t per = ((n + parts - 1) / parts + 255) / 256 * 256;      // equal, 256-aligned parts
    for (int64_t o = 0; o < n; o += per) {
      EvalItem it; it.cpw = cpw + o; it.nsd = nsd + o; it.count = (int32_t)std::min<int64_t>(per, n - o); it.kf = k;
      items.push_back(it);
    }
  }
  start[W] = (int)items.size();
  c->n_items = (int)items.size();
  c->d_items.reserve(items.size() + 1); c->d_kf_item_start.reserve(W + 1);
  c->d_partials.reserve((items.size() + 1) * GLIO_NACC);
  c->d_out.reserve((size_t)W * GLIO_NACC); c->h_out.reserve((size_t)W * GLIO_NACC);
  c->d_poses.reserve((size_t)W * 7); c->h_poses.reserve((size_t)W * 7);
  if (!c->d_ticket.p) { c->d_ticket.reserve(1); GLIO_CUDA_TRY(cudaMemsetAsync(c->d_ticket.p, 0, sizeof(unsigned int), c->st)); }
  if (!items.empty())
    GLIO_CUDA_TRY(cudaMemcpyAsync(c->d_items.p, items.data(), items.size() * sizeof(EvalItem), cudaMemcpyHostToDevice, c->st));
  GLIO_CUDA_TRY(cudaMemcpyAsync(c->d_kf_item_start.p, start.data(), (W + 1) * sizeof(int), cudaMemcpyHostToDevice, c->st));
  GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));   // items/start are stack vectors
  c->items_dirty = false;
  c->items_W = W;
}

EvalParams eval_params(const glio_ctx* c) {
  EvalParams ep;
  for (int k = 0; k < 4; ++k) ep.q_lb[k] = c->prm.q_lb[k];
  for (int k = 0; k < 3; ++k) ep.t_lb[k] = c->prm.t_lb[k];
  ep.lidar_const = c->prm.lidar_const;
  ep.huber_delta = c->prm.huber_delta;
  ep.unit_score = c->prm.unit_score != 0 ? 1 : 0;
  return ep;
}

void fill_summary(const SolverSummary& S, int n, glio_solver_summary* summary, glio_iteration* iter_log, int iter_cap, double* step_log, int64_t step_cap) {
  if (summary) {
      memset(summary, 0, sizeof(*summary));
      summary->termination = S.termination; summary->num_iterations = (int)S.iterations.size();
      summary->num_successful_steps = S.num_successful_steps; summary->num_unsuccessful_steps = S.num_unsuccessful_steps;
      summary->num_evaluations = S.num_evaluations; summary->num_jacobian_evaluations = S.num_jacobian_evaluations;
      summary->num_linear_solves = S.num_linear_solves; summary->num_valid_steps = (int)(S.steps.size() / (size_t)n);
      summary->initial_cost = S.initial_cost; summary->final_cost = S.final_cost;
      summary->eval_seconds = S.eval_seconds; summary->linear_solver_seconds = S.linear_solver_seconds; summary->total_seconds = S.total_seconds;
      snprintf(summary->message, sizeof(summary->message), "%s", S.message.c_str());
    }
    if (iter_log) for (int i = 0; i < (int)S.iterations.size() && i < iter_cap; ++i) {
      const IterationRecord& r = S.iterations[i];
      glio_iteration& q = iter_log[i];
      q.iteration = r.iteration; q.step_is_valid = r.step_is_valid; q.step_is_successful = r.step_is_successful; q.reserved = 0;
      q.cost = r.cost; q.cost_change = r.cost_change; q.gradient_max_norm = r.gradient_max_norm; q.gradient_norm = r.gradient_norm;
      q.step_norm = r.step_norm; q.relative_decrease = r.relative_decrease; q.trust_region_radius = r.trust_region_radius; q.mu = r.mu;
    }
    if (step_log) memcpy(step_log, S.steps.data(), sizeof(double) * (size_t)std::min<int64_t>(step_cap, (int64_t)S.steps.size()));
}

// device evaluation of all active unary residuals -> c->h_out (W x 28: 21 upper-tri H, 6 g, 1 cost), synchronised
// launch half: returns the epoch to wait for (0: nothing was launched, h_out already holds zeros)
unsigned int eval_unary_launch(glio_ctx* c, int W, const double* poses_body, int jac_kind, bool want_jac) {
  if (c->marg_pending) {            // an asynchronous marginalisation still owns the result buffer until its worker has copied it
    while (!c->marg_pending->copied.load(std::memory_order_acquire)) {
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
  }
  build_items(c, W);
  // Zero-copy round trip: the W poses travel in the kernel parameters, the last block writes the W x 28 result into
  // pinned host memory and raises an epoch flag the host spins on.  This replaces H2D copy + launch + D2H copy +
  // stream synchronise on the per-iteration path, and lets the caller evaluate its host factors while the kernel runs.
  memcpy(c->h_poses.p, poses_body, (size_t)W * 7 * sizeof(double));
  if (c->n_items <= 0) {
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    memset(c->h_out.p, 0, (size_t)W * GLIO_NACC * sizeof(double));
    return 0u;
  }
  if (!c->h_flag.p) { c->h_flag.reserve(4); memset(c->h_flag.p, 0, 4 * sizeof(unsigned int)); }
  const unsigned int epoch = ++c->eval_epoch;
  if (W > EV_MAXW) GLIO_CUDA_TRY(cudaMemcpyAsync(c->d_poses.p, c->h_poses.p, (size_t)W * 7 * sizeof(double), cudaMemcpyHostToDevice, c->st));
  eval_unary_run(c->d_items.p, c->n_items, W, c->d_poses.p, eval_params(c), jac_kind, want_jac, c->d_partials.p, c->h_out.p,
                 c->d_kf_item_start.p, c->d_ticket.p, c->st, c->lc, c->h_flag.p, epoch, c->h_poses.p);
  (void)cudaStreamQuery(c->st);            // pushes the launch to the device now (otherwise it may sit in the driver's queue until the next API call)
  return epoch;
}

void eval_unary_wait(glio_ctx* c, unsigned int epoch) {
  if (!epoch) return;
  volatile unsigned int* flag = c->h_flag.p;
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned long long spins = 0; *flag != epoch; ++spins) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
    if ((spins & 0x3ffull) == 0x3ffull) {
      const cudaError_t q = cudaStreamQuery(c->st);
      if (q != cudaSuccess && q != cudaErrorNotReady) GLIO_CUDA_TRY(q);
      if (q == cudaSuccess && *flag != epoch) throw Error{GLIO_ERR_CUDA, "eval_unary: kernel finished without raising the completion flag"};
      if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 30.0) throw Error{GLIO_ERR_CUDA, "eval_unary: timed out waiting for the device"};
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
}

// device evaluation of all active unary residuals -> c->h_out (W x 28: 21 upper-tri H, 6 g, 1 cost), synchronised
void eval_unary_blocks(glio_ctx* c, int W, const double* poses_body, int jac_kind, bool want_jac) {
  eval_unary_wait(c, eval_unary_launch(c, W, poses_body, jac_kind, want_jac));
}

}  // namespace

extern "C" {

void glio_default_params(glio_params* p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  p->kd_max_radius = 1.5; p->surf_dist_thres = 0.18; p->lidar_const = 7.5; p->weight_min = 0.3; p->huber_delta = 1.0;
  p->q_lb[0] = 1.0; p->t_lb[2] = 0.28;          // GLIO/config/config_urban_hk.yaml:90-97
  p->batch_max_radius = 1.5; p->batch_dist_thres = 0.18; p->batch_score = 2.5;
  p->cell_size = 0.f; p->keep_debug = 0; p->unit_score = 0;
}

int glio_create(int device, const glio_params* params, glio_ctx** out) {
  if (!out) return GLIO_ERR_ARG;
  *out = nullptr;
  glio_ctx* c = nullptr;
  int rc = guarded(nullptr, [&] {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) throw Error{GLIO_ERR_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e) + " (glio_b200 has no CPU fallback)"};
    GLIO_REQUIRE(device >= 0 && device < ndev, GLIO_ERR_ARG, "device index out of range");
    GLIO_CUDA_TRY(cudaSetDevice(device));
    c = new glio_ctx();
    c->device = device;
    if (params) c->prm = *params; else glio_default_params(&c->prm);
    if (const char* e = getenv("GLIO_KNN_MODE")) { const int v = atoi(e); if (v >= 0 && v <= 7) c->knn_mode = v; }
    if (const char* e = getenv("GLIO_KNN_GROW")) { const int v = atoi(e); if (v >= 0 && v <= 1) c->grow_mode = v; }
    if (const char* e = getenv("GLIO_TILE_RINGS")) { const int v = atoi(e); if (v >= 1 && v < 64) c->tile_rings = v; }
    if (const char* e = getenv("GLIO_PTS_PER_CELL")) { const float v = (float)atof(e); if (v > 0.1f && v < 1000.f) c->pts_per_cell = v; }
    c->map.build_pairs = c->knn_mode == 4;
    GLIO_CUDA_TRY(cudaStreamCreateWithFlags(&c->st, cudaStreamNonBlocking));
    GLIO_CUDA_TRY(cudaStreamCreateWithFlags(&c->st_copy, cudaStreamNonBlocking));
    GLIO_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_scans, cudaEventDisableTiming));
    GLIO_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    { int sms = 148; if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) c->eval_slots = 2 * sms; }
  });
  if (rc != GLIO_OK) { delete c; return rc; }
  *out = c;
  return GLIO_OK;
}

void glio_destroy(glio_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->marg_thread.joinable()) {
    if (c->marg_pending) { glio_marg_prior* dropped = nullptr; glio_marg_job_wait(c->marg_pending, &dropped); if (dropped) glio_marg_prior_destroy(dropped); }
    { std::lock_guard<std::mutex> lk(c->marg_mu); c->marg_quit = true; }
    c->marg_cv.notify_one();
    c->marg_thread.join();
  }
  if (c->st_copy) cudaStreamSynchronize(c->st_copy);
  if (c->st) cudaStreamSynchronize(c->st);
  c->map.release(); c->map_stage.release(); c->map_stage2.release(); if (c->ev_map) cudaEventDestroy(c->ev_map);
  for (auto& f : c->lm_frames) f->pts.release();
  for (auto& f : c->lm_spare) f->pts.release();
  c->lm_frames.clear(); c->lm_spare.clear(); c->lm_stage.release(); c->lm_cat.release(); c->lm_vox.release();
  for (auto& f : c->frames) { f.second->scan.release(); f.second->grid.release(); }
  for (auto& p : c->pairs) { p->m_cpw.release(); p->m_nc.release(); p->m_src.release(); p->s_cpw.release(); p->s_nc.release(); }
  c->d_bin_items.release(); c->d_pair_item_start.release(); c->d_kf_inc_start.release(); c->d_inc.release(); c->d_bin_partials.release();
  c->d_pair_sums.release(); c->d_bin_diag.release(); c->d_bin_off.release(); c->d_bin_poses.release(); c->h_bin_diag.release(); c->h_bin_off.release(); c->h_bin_poses.release();
  for (auto& s : c->slots) if (s) s->release();
  c->w_pm.release(); c->w_pmq.release(); c->w_nsd.release(); c->w_seg.release(); c->w_status.release(); c->w_weight.release();
  c->w_nc.release(); c->w_plane.release(); c->w_idx5.release(); c->w_sqd5.release(); c->w_flags.release(); c->w_pos.release();
  c->cell_count.release(); c->cell_pos.release(); c->scan_tmp.release(); c->d_segs.release(); c->d_dst.release(); c->d_counts.release();
  c->h_counts.release(); c->d_bad.release(); c->d_keep.release(); c->d_items.release(); c->d_kf_item_start.release();
  c->d_partials.release(); c->d_out.release(); c->d_poses.release(); c->d_r.release(); c->d_J.release(); c->h_out.release();
  c->h_poses.release(); c->h_flag.release(); c->d_ticket.release(); c->w_knn_idx.release(); c->w_knn_sqd.release(); c->d_stats.release(); c->w_deferred.release();
  c->f_cloud.release(); c->f_curv.release(); c->f_label.release(); c->f_picked.release(); c->f_scan.release(); c->f_ring.release(); c->f_lf.release(); c->f_counts.release(); c->f_out.release(); c->f_ds.release(); c->f_outds.release();
  c->e_items.release(); c->e_start.release(); c->e_part.release(); c->e_out.release(); c->e_poses.release();
  for (auto& r : c->lc.recs) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  for (auto& e : c->lc.pool) cudaEventDestroy(e);
  c->lc.recs.clear(); c->lc.pool.clear();
  if (c->ev_scans) cudaEventDestroy(c->ev_scans);
  if (c->ev_main) cudaEventDestroy(c->ev_main);
  if (c->st_copy) cudaStreamDestroy(c->st_copy);
  if (c->st) cudaStreamDestroy(c->st);
  delete c;
}

const char* glio_last_error(const glio_ctx* c) { return c ? c->err.c_str() : global_error(); }

int glio_synchronize(glio_ctx* c) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] { GLIO_CUDA_TRY(cudaStreamSynchronize(c->st_copy)); GLIO_CUDA_TRY(cudaStreamSynchronize(c->st)); });
}
void* glio_stream(glio_ctx* c) { return c ? (void*)c->st : nullptr; }
int64_t glio_launch_count(const glio_ctx* c) { return c ? c->lc.n : 0; }

int glio_profile_enable(glio_ctx* c, int on) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    for (auto& r : c->lc.recs) { c->lc.pool.push_back(r.a); c->lc.pool.push_back(r.b); }
    c->lc.recs.clear();
    c->lc.prof = on != 0;
    // events are created up front: cudaEventCreate inside the timed region would distort what is being measured
    if (on) while (c->lc.pool.size() < 16384) { cudaEvent_t e; GLIO_CUDA_TRY(cudaEventCreate(&e)); c->lc.pool.push_back(e); }
  });
}

int glio_profile_get(glio_ctx* c, const char* kernel_name, double* ms_total, int64_t* launches) {
  if (!c || !kernel_name) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    double tot = 0; int64_t cnt = 0;
    for (auto& r : c->lc.recs) {
      if (strcmp(r.name, kernel_name) != 0) continue;
      float ms = 0.f;
      if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { tot += ms; ++cnt; }
    }
    if (ms_total) *ms_total = tot;
    if (launches) *launches = cnt;
  });
}

int glio_get_stats(glio_ctx* c, int64_t* knn_fallback_queries, int reset) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    unsigned long long h[4] = {0, 0, 0, 0};
    if (c->d_stats.p) {
      GLIO_CUDA_TRY(cudaMemcpyAsync(h, c->d_stats.p, sizeof(h), cudaMemcpyDeviceToHost, c->st));
      GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
      if (reset) GLIO_CUDA_TRY(cudaMemsetAsync(c->d_stats.p, 0, sizeof(h), c->st));
    }
    if (knn_fallback_queries) *knn_fallback_queries = (int64_t)h[0];
  });
}

int glio_get_params(const glio_ctx* c, glio_params* out) {
  if (!c || !out) return GLIO_ERR_ARG;
  *out = c->prm;
  return GLIO_OK;
}

// direct upload of a slot's match list (association done elsewhere, e.g. by the unmodified Estimator): the Ceres shim
// uses this to move LidarPlaneNormFactor residual blocks onto the device.
int glio_set_matches(glio_ctx* c, int slot, const float* cp, const float* nsd, const float* weight, int64_t n) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    Slot& sl = c->slot(slot);
    sl.n_match = 0; sl.n_sel = -1; c->items_dirty = true;
    if (n <= 0) return;
    GLIO_REQUIRE(cp && nsd && weight, GLIO_ERR_ARG, "null match arrays");
    std::vector<float4> h0(n), h1(n);
    for (int64_t i = 0; i < n; ++i) {
      h0[i] = make_float4(cp[3 * i], cp[3 * i + 1], cp[3 * i + 2], weight[i]);
      h1[i] = make_float4(nsd[4 * i], nsd[4 * i + 1], nsd[4 * i + 2], nsd[4 * i + 3]);
    }
    sl.m_cpw.reserve(n); sl.m_nsd.reserve(n); sl.m_src.reserve(n);
    GLIO_CUDA_TRY(cudaMemcpyAsync(sl.m_cpw.p, h0.data(), n * sizeof(float4), cudaMemcpyHostToDevice, c->st));
    GLIO_CUDA_TRY(cudaMemcpyAsync(sl.m_nsd.p, h1.data(), n * sizeof(float4), cudaMemcpyHostToDevice, c->st));
    GLIO_CUDA_TRY(cudaMemsetAsync(sl.m_src.p, 0, n * sizeof(int32_t), c->st));
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    sl.n_match = n;
  });
}

int glio_get_match_counts(glio_ctx* c, int W, int64_t* n_match, int64_t* n_active) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    for (int k = 0; k < W; ++k) {
      Slot& sl = c->slot(k);
      if (n_match) n_match[k] = sl.n_match;
      if (n_active) n_active[k] = sl.n_sel >= 0 ? sl.n_sel : sl.n_match;
    }
  });
}

void glio_lidar_pose(const glio_params* prm, const double pose_body[7], double t2[3], double q2[4]) {
  lidar_pose_in_map(prm->q_lb, prm->t_lb, pose_body, t2, q2);
}

int glio_set_map(glio_ctx* c, const float* xyz, int64_t M, int stride, int mem) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(M >= 5, GLIO_ERR_ARG, "map needs at least 5 points");
    const float* d;
    if (mem != GLIO_DEVICE && c->pf_ptr == xyz && c->pf_M == M && c->pf_stride == stride) {
      // the upload was started by glio_map_prefetch: wait for it on the main stream, no second copy
      GLIO_CUDA_TRY(cudaStreamWaitEvent(c->st, c->ev_map, 0));
      std::swap(c->map_stage, c->map_stage2);
      d = c->map_stage.p;
    } else d = stage_points(c, c->map_stage, xyz, M, stride, mem);
    c->pf_ptr = nullptr;
    grid_build(c->map, d, stride, M, nullptr, nullptr, c->prm.cell_size, c->pts_per_cell, c->st, c->lc);
    c->has_map = true; c->map_n = M;
  });
}

// Start the host->device copy of the NEXT local map on the copy stream (returns at once).  A following glio_set_map with the same
// pointer, count and stride uses the staged copy.  The caller must not modify the host buffer until that glio_set_map.
int glio_map_prefetch(glio_ctx* c, const float* xyz, int64_t M, int stride) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(xyz != nullptr && M >= 5 && stride >= 3 && stride <= 64, GLIO_ERR_ARG, "glio_map_prefetch: bad arguments");
    if (!c->ev_map) GLIO_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_map, cudaEventDisableTiming));
    c->map_stage2.reserve((size_t)M * stride);
    GLIO_CUDA_TRY(cudaMemcpyAsync(c->map_stage2.p, xyz, (size_t)M * stride * sizeof(float), cudaMemcpyHostToDevice, c->st_copy));
    GLIO_CUDA_TRY(cudaEventRecord(c->ev_map, c->st_copy));
    c->pf_ptr = xyz; c->pf_M = M; c->pf_stride = stride;
  });
}

// ---- local map maintenance on the device (Estimator.cpp:3529-3631 buildLocalMapWithLandMark, :3615-3618 downSampleCloud) ----
int glio_localmap_clear(glio_ctx* c) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    for (auto& f : c->lm_frames) f->pts.release();
    c->lm_frames.clear();
    for (auto& f : c->lm_spare) f->pts.release();
    c->lm_spare.clear();
  });
}

int glio_localmap_push(glio_ctx* c, int at_front, const float* cloud_xyz, int64_t n, int stride, int mem, const double t[3], const double q[4]) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(t && q, GLIO_ERR_ARG, "null pose");
    const float* d = stage_points(c, c->lm_stage, cloud_xyz, n, stride, mem);
    std::unique_ptr<glio_ctx::LmFrame> f;
    if (!c->lm_spare.empty()) { f = std::move(c->lm_spare.back()); c->lm_spare.pop_back(); }
    else f = std::make_unique<glio_ctx::LmFrame>();
    f->pts.reserve((size_t)n); f->n = n;
    localmap_transform(d, stride, n, t, q, f->pts.p, c->st, c->lc);        // transformCloud: float(q*double(p)+t)
    if (mem != GLIO_DEVICE) GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));    // the staging buffer is reused by the next push
    if (at_front) c->lm_frames.push_front(std::move(f)); else c->lm_frames.push_back(std::move(f));
  });
}

int glio_localmap_pop_front(glio_ctx* c) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(!c->lm_frames.empty(), GLIO_ERR_STATE, "local map holds no keyframe cloud");
    // the stream is ordered: anything still reading the popped buffer was enqueued before whatever reuses it
    c->lm_spare.push_back(std::move(c->lm_frames.front()));
    c->lm_frames.pop_front();
  });
}

int glio_localmap_size(glio_ctx* c, int* n_frames, int64_t* n_points) {
  if (!c) return GLIO_ERR_ARG;
  int64_t np = 0;
  for (auto& f : c->lm_frames) np += f->n;
  if (n_frames) *n_frames = (int)c->lm_frames.size();
  if (n_points) *n_points = np;
  return GLIO_OK;
}

int glio_localmap_build(glio_ctx* c, float leaf, int64_t* n_map) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    int64_t N = 0;
    for (auto& f : c->lm_frames) N += f->n;
    GLIO_REQUIRE(N >= 5, GLIO_ERR_STATE, "local map needs at least 5 points");
    c->lm_cat.reserve((size_t)N);
    int64_t o = 0;
    for (auto& f : c->lm_frames) {                                          // *surf_local_map += *recent_surf_keyframes[i]
      GLIO_CUDA_TRY(cudaMemcpyAsync(c->lm_cat.p + o, f->pts.p, (size_t)f->n * sizeof(float4), cudaMemcpyDeviceToDevice, c->st));
      o += f->n;
    }
    const int64_t m = leaf > 0.f ? voxel_filter_run(c->lm_vox, c->lm_cat.p, N, leaf, c->st, c->lc) : -1;   // ds_filter_surf_map.filter
    if (m >= 0) {
      GLIO_REQUIRE(m >= 5, GLIO_ERR_STATE, "down-sampled local map has fewer than 5 points");
      grid_build(c->map, c->lm_vox.out_xyz.p, 3, m, nullptr, nullptr, c->prm.cell_size, c->pts_per_cell, c->st, c->lc);   // kd_tree_surf_local_map->setInputCloud
      c->map_n = m;
    } else {                                                                // leaf <= 0, or PCL's pass-through when the voxel grid would overflow
      grid_build(c->map, reinterpret_cast<const float*>(c->lm_cat.p), 4, N, nullptr, nullptr, c->prm.cell_size, c->pts_per_cell, c->st, c->lc);
      c->map_n = N;
    }
    c->has_map = true;
    if (n_map) *n_map = c->map_n;
  });
}

int glio_get_map(glio_ctx* c, int64_t capacity, float* xyz, int64_t* n_out) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(c->has_map, GLIO_ERR_STATE, "no map set");
    const int64_t n = std::min<int64_t>(capacity, c->map_n);
    if (n_out) *n_out = c->map_n;
    if (n <= 0 || !xyz) return;
    std::vector<float> h((size_t)n * 4);
    GLIO_CUDA_TRY(cudaMemcpyAsync(h.data(), c->map.tmp4.p, (size_t)n * sizeof(float4), cudaMemcpyDeviceToHost, c->st));
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    for (int64_t i = 0; i < n; ++i) { xyz[3 * i] = h[4 * i]; xyz[3 * i + 1] = h[4 * i + 1]; xyz[3 * i + 2] = h[4 * i + 2]; }
  });
}

int glio_assoc_scan_to_map(glio_ctx* c, int slot, const float* scan_xyz, int64_t Q, int stride, int mem,
                           const double t[3], const double q[4], int64_t* n_match) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(t && q, GLIO_ERR_ARG, "null pose");
    Slot& sl = c->slot(slot);
    sl.scan_ptr = stage_points(c, sl.scan, scan_xyz, Q, stride, mem);
    sl.Q = Q; sl.stride = stride;
    std::vector<std::array<double, 7>> lp(1);
    for (int k = 0; k < 3; ++k) lp[0][k] = t[k];
    for (int k = 0; k < 4; ++k) lp[0][3 + k] = q[k];
    associate_slots(c, std::vector<int>{slot}, lp, n_match);
  });
}

int glio_window_set_scans(glio_ctx* c, int W, const float* const* scans, const int64_t* Q, int stride, int mem) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(W > 0 && scans && Q, GLIO_ERR_ARG, "bad window arguments");
    GLIO_REQUIRE(stride >= 3 && stride <= 64, GLIO_ERR_ARG, "stride_floats must be in [3,64]");
    // Host scans are uploaded on the copy stream: the call returns at once and the next association waits for the
    // event.  The matches of the previous association stay valid (they hold their own copies of the points), so the
    // scans of the NEXT window can be handed over while the current window is still being solved.
    if (mem != GLIO_DEVICE) {
      GLIO_CUDA_TRY(cudaEventRecord(c->ev_main, c->st));
      GLIO_CUDA_TRY(cudaStreamWaitEvent(c->st_copy, c->ev_main, 0));      // nothing queued on the main stream may still read the slot buffers
    }
    for (int k = 0; k < W; ++k) {
      Slot& sl = c->slot(k);
      GLIO_REQUIRE(scans[k] != nullptr && Q[k] > 0, GLIO_ERR_ARG, "null or empty point array");
      if (mem == GLIO_DEVICE) sl.scan_ptr = scans[k];
      else {
        sl.scan.reserve((size_t)Q[k] * stride);
        GLIO_CUDA_TRY(cudaMemcpyAsync(sl.scan.p, scans[k], (size_t)Q[k] * stride * sizeof(float), cudaMemcpyHostToDevice, c->st_copy));
        sl.scan_ptr = sl.scan.p;
      }
      sl.Q = Q[k]; sl.stride = stride;
    }
    if (mem != GLIO_DEVICE) { GLIO_CUDA_TRY(cudaEventRecord(c->ev_scans, c->st_copy)); c->scans_pending = true; }
  });
}

// One keyframe's scan into one slot (copy stream for host buffers, like glio_window_set_scans): what a sliding window needs per
// new keyframe - the other W-1 scans are already resident.
int glio_window_set_scan(glio_ctx* c, int slot, const float* scan, int64_t Q, int stride, int mem) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(scan != nullptr && Q > 0, GLIO_ERR_ARG, "null or empty point array");
    GLIO_REQUIRE(stride >= 3 && stride <= 64, GLIO_ERR_ARG, "stride_floats must be in [3,64]");
    Slot& sl = c->slot(slot);
    if (mem == GLIO_DEVICE) sl.scan_ptr = scan;
    else {
      GLIO_CUDA_TRY(cudaEventRecord(c->ev_main, c->st));
      GLIO_CUDA_TRY(cudaStreamWaitEvent(c->st_copy, c->ev_main, 0));
      sl.scan.reserve((size_t)Q * stride);
      GLIO_CUDA_TRY(cudaMemcpyAsync(sl.scan.p, scan, (size_t)Q * stride * sizeof(float), cudaMemcpyHostToDevice, c->st_copy));
      sl.scan_ptr = sl.scan.p;
      GLIO_CUDA_TRY(cudaEventRecord(c->ev_scans, c->st_copy)); c->scans_pending = true;
    }
    sl.Q = Q; sl.stride = stride;
  });
}

// slideWindow (GLIO/src/Estimator.cpp: the oldest keyframe leaves, every other one moves down one slot): O(1) per slot, no
// copies; the matches of the slots move with them.  The new keyframe's scan then goes into slot W-1 with glio_window_set_scan.
int glio_window_slide(glio_ctx* c, int W) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(W >= 2 && W <= 4096, GLIO_ERR_ARG, "bad window size");
    c->slot(W - 1);
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st_copy));
    std::unique_ptr<Slot> first = std::move(c->slots[0]);
    for (int k = 0; k + 1 < W; ++k) c->slots[k] = std::move(c->slots[k + 1]);
    c->slots[W - 1] = std::move(first);                       // its buffers are reused by the next upload
    if (c->slots[W - 1]) { c->slots[W - 1]->Q = 0; c->slots[W - 1]->scan_ptr = nullptr; c->slots[W - 1]->n_match = 0; c->slots[W - 1]->n_sel = -1; }
    c->items_dirty = true;
  });
}

int glio_window_associate(glio_ctx* c, int W, const double* poses_body, int64_t* n_match) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(W > 0 && poses_body, GLIO_ERR_ARG, "bad window arguments");
    std::vector<int> ids(W);
    std::vector<std::array<double, 7>> lp(W);
    for (int k = 0; k < W; ++k) {
      ids[k] = k;
      lidar_pose_in_map(c->prm.q_lb, c->prm.t_lb, poses_body + 7 * k, &lp[k][0], &lp[k][3]);   // Estimator.cpp:2216-2217
    }
    associate_slots(c, ids, lp, n_match);
  });
}

int glio_get_matches(glio_ctx* c, int slot, int64_t capacity, float* cp, float* nsd, float* weight, int32_t* src, int64_t* n_out) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    Slot& sl = c->slot(slot);
    const int64_t n = sl.n_match;
    if (n_out) *n_out = n;
    GLIO_REQUIRE(capacity >= n, GLIO_ERR_ARG, "capacity smaller than the match count");
    if (n == 0) return;
    std::vector<float4> h(n);
    if (cp || weight) {
      GLIO_CUDA_TRY(cudaMemcpyAsync(h.data(), sl.m_cpw.p, n * sizeof(float4), cudaMemcpyDeviceToHost, c->st));
      GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
      for (int64_t i = 0; i < n; ++i) {
        if (cp) { cp[3 * i] = h[i].x; cp[3 * i + 1] = h[i].y; cp[3 * i + 2] = h[i].z; }
        if (weight) weight[i] = h[i].w;
      }
    }
    if (nsd) {
      GLIO_CUDA_TRY(cudaMemcpyAsync(nsd, sl.m_nsd.p, n * sizeof(float4), cudaMemcpyDeviceToHost, c->st));
      GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    }
    if (src) {
      GLIO_CUDA_TRY(cudaMemcpyAsync(src, sl.m_src.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
      GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    }
  });
}

int glio_get_assoc_debug(glio_ctx* c, int slot, int64_t Q, uint8_t* status, int32_t* idx5, float* sqd5, float* pm, double* plane) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(c->prm.keep_debug, GLIO_ERR_STATE, "context was created without keep_debug");
    Slot& sl = c->slot(slot);
    GLIO_REQUIRE(sl.dbg_Q == Q && Q > 0, GLIO_ERR_ARG, "Q does not match the slot's last association");
    if (status) GLIO_CUDA_TRY(cudaMemcpyAsync(status, sl.dbg_status.p, Q, cudaMemcpyDeviceToHost, c->st));
    if (idx5) GLIO_CUDA_TRY(cudaMemcpyAsync(idx5, sl.dbg_idx5.p, 5 * Q * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
    if (sqd5) GLIO_CUDA_TRY(cudaMemcpyAsync(sqd5, sl.dbg_sqd5.p, 5 * Q * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    if (plane) GLIO_CUDA_TRY(cudaMemcpyAsync(plane, sl.dbg_plane.p, 4 * Q * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    std::vector<float4> h;
    if (pm) { h.resize(Q); GLIO_CUDA_TRY(cudaMemcpyAsync(h.data(), sl.dbg_pm.p, Q * sizeof(float4), cudaMemcpyDeviceToHost, c->st)); }
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    if (pm) for (int64_t i = 0; i < Q; ++i) { pm[3 * i] = h[i].x; pm[3 * i + 1] = h[i].y; pm[3 * i + 2] = h[i].z; }
  });
}

int glio_select(glio_ctx* c, int slot, const int32_t* keep, int64_t n) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    Slot& sl = c->slot(slot);
    c->items_dirty = true;
    if (n < 0) { sl.n_sel = -1; return; }
    GLIO_REQUIRE(n == 0 || keep, GLIO_ERR_ARG, "null selection list");
    sl.n_sel = n;
    if (n == 0) return;
    c->d_keep.reserve(n); c->d_bad.reserve(1);
    sl.s_cpw.reserve(n); sl.s_nsd.reserve(n);
    GLIO_CUDA_TRY(cudaMemcpyAsync(c->d_keep.p, keep, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->st));
    GLIO_CUDA_TRY(cudaMemsetAsync(c->d_bad.p, 0, sizeof(int), c->st));
    gather_selection(c->d_keep.p, n, sl.n_match, sl.m_cpw.p, sl.m_nsd.p, nullptr, sl.s_cpw.p, sl.s_nsd.p, nullptr, c->d_bad.p, c->st, c->lc);
    int bad = 0;
    GLIO_CUDA_TRY(cudaMemcpyAsync(&bad, c->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    if (bad) { sl.n_sel = -1; throw Error{GLIO_ERR_ARG, "selection index out of range"}; }
  });
}

int glio_eval_unary(glio_ctx* c, int W, const double* poses_body, int jac_kind, double* H, double* g, double* cost) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(W > 0 && W <= 4096 && poses_body, GLIO_ERR_ARG, "bad arguments");
    GLIO_REQUIRE(jac_kind == 0 || jac_kind == 1, GLIO_ERR_ARG, "jac_kind must be 0 or 1");
    const bool want_jac = (H != nullptr) || (g != nullptr);
    eval_unary_blocks(c, W, poses_body, jac_kind, want_jac);
    for (int k = 0; k < W; ++k) {
      const double* o = c->h_out.p + (size_t)k * GLIO_NACC;
      if (H) {
        int idx = 0;
        for (int p = 0; p < 6; ++p) for (int q = p; q < 6; ++q) { H[36 * k + 6 * p + q] = o[idx]; H[36 * k + 6 * q + p] = o[idx]; ++idx; }
      }
      if (g) for (int p = 0; p < 6; ++p) g[6 * k + p] = o[21 + p];
      if (cost) cost[k] = o[27];
    }
  });
}

// ---- K3: marginalisation of KF0 (MarginalizationInfo::PreMarginalize + Marginalize, MarginalizationFactor.cpp:107-202)
}  // extern "C"

namespace {

// host half: the device blocks (W x 28 doubles: 21 upper-triangular H, 6 g, cost) join A/b, then Schur + decomposition + prior
glio_marg_prior* marg_finish_host(int W, double eps, std::vector<double>& A, std::vector<double>& b, const double* hout,
                                  const double* x0_pose, const double* x0_sb) {
  const int N = 6 * W + 18, m = 15, n = N - m;
  static const int ut[6][6] = {{0, 1, 2, 3, 4, 5}, {1, 6, 7, 8, 9, 10}, {2, 7, 11, 12, 13, 14}, {3, 8, 12, 15, 16, 17}, {4, 9, 13, 16, 18, 19}, {5, 10, 14, 17, 19, 20}};
  for (int k = 0; k < W; ++k) {
    const double* o = hout + (size_t)k * GLIO_NACC;
    const int base = marg_index_t(k);                       // t at base, q at base + 3 in every keyframe of this ordering
    for (int p = 0; p < 6; ++p) { b[base + p] += o[21 + p]; for (int q = 0; q < 6; ++q) A[(size_t)(base + p) * N + base + q] += o[ut[p][q]]; }
  }
  std::vector<double> LJ((size_t)n * n), lr(n);
  if (marginalize_dense(A.data(), b.data(), N, m, eps, LJ.data(), lr.data()) != GLIO_OK) return nullptr;
  return glio_marg_prior_create(W, LJ.data(), lr.data(), x0_pose, x0_sb);
}

// device half + the caller's host factors: queues the LiDAR re-evaluation of every keyframe with the ambient x,y,z quaternion
// columns (Estimator.cpp:2538-2576) and evaluates the host callback while the kernel runs
unsigned int marg_begin(glio_ctx* c, int W, const double* poses, const double* speed_bias, glio_host_marg_fn host_marg, void* user,
                        std::vector<double>& A, std::vector<double>& b) {
  const int N = 6 * W + 18;
  const unsigned int epoch = eval_unary_launch(c, W, poses, 1, true);
  A.assign((size_t)N * N, 0.0); b.assign(N, 0.0);
  if (host_marg) GLIO_REQUIRE(host_marg(user, W, poses, speed_bias, A.data(), b.data()) == 0, GLIO_ERR_STATE, "host_marg callback failed");
  return epoch;
}

void marg_worker(glio_ctx* c) {
  cudaSetDevice(c->device);
#if defined(__linux__)
  cpu_set_t base; CPU_ZERO(&base);
  const bool have_base = pthread_getaffinity_np(pthread_self(), sizeof(base), &base) == 0;
  int excluded = -1;
#endif
  for (;;) {
    glio_marg_job* j = nullptr;
    {
      std::unique_lock<std::mutex> lk(c->marg_mu);
      c->marg_cv.wait(lk, [&] { return c->marg_queued != nullptr || c->marg_quit; });
      if (c->marg_quit && !c->marg_queued) return;
      j = c->marg_queued; c->marg_queued = nullptr;
    }
#if defined(__linux__)
    // The submitting thread is about to queue the next window's kernels: a worker woken onto ITS core would preempt it for the
    // whole host half and the GPU would idle meanwhile.  Keep off that core (the rest of the process mask stays allowed).
    if (have_base && j->submit_cpu >= 0 && j->submit_cpu != excluded && CPU_COUNT(&base) > 1) {
      cpu_set_t m = base; CPU_CLR(j->submit_cpu, &m);
      if (CPU_COUNT(&m) > 0 && pthread_setaffinity_np(pthread_self(), sizeof(m), &m) == 0) excluded = j->submit_cpu;
    }
#endif
    try {
      eval_unary_wait(c, j->epoch);
      j->out.assign(c->h_out.p, c->h_out.p + (size_t)j->W * GLIO_NACC);
      j->copied.store(1, std::memory_order_release);
      j->prior = marg_finish_host(j->W, j->eps, j->A, j->b, j->out.data(), j->x0_pose.data(), j->x0_sb.data());
      if (!j->prior) { j->rc = GLIO_ERR_STATE; j->err = "marginalisation failed (decomposition or prior)"; }
    } catch (const Error& e) { j->rc = e.code; j->err = e.msg; }
    catch (const std::exception& e) { j->rc = GLIO_ERR_STATE; j->err = e.what(); }
    j->copied.store(1, std::memory_order_release);
    j->done.store(1, std::memory_order_release);
  }
}

}  // namespace

extern "C" {

int glio_window_marginalize(glio_ctx* c, int W, const double* poses, const double* speed_bias, glio_host_marg_fn host_marg, void* user,
                            double eps, glio_marg_prior** prior_out) {
  if (!c || !prior_out) return GLIO_ERR_ARG;
  *prior_out = nullptr;
  return guarded(c, [&] {
    GLIO_REQUIRE(W >= 2 && W <= 4096 && poses && speed_bias, GLIO_ERR_ARG, "glio_window_marginalize: need W >= 2, poses and speed_bias");
    std::vector<double> A, b;
    const unsigned int epoch = marg_begin(c, W, poses, speed_bias, host_marg, user, A, b);
    eval_unary_wait(c, epoch);
    glio_marg_prior* made = marg_finish_host(W, eps, A, b, c->h_out.p, poses + 7, speed_bias + 9);
    GLIO_REQUIRE(made != nullptr, GLIO_ERR_STATE, "marginalisation failed (decomposition or prior)");
    *prior_out = made;
  });
}

// The same pass with its host half on the context's worker thread: returns as soon as the device half is queued and the host
// callback has run; the Schur complement, the decomposition and the prior are computed while the caller hands the next window's
// association to the GPU.  One job at a time per context; glio_marg_job_wait joins it (and must be called before the next solve
// needs the prior).  Results are identical to glio_window_marginalize.
int glio_window_marginalize_async(glio_ctx* c, int W, const double* poses, const double* speed_bias, glio_host_marg_fn host_marg, void* user,
                                  double eps, glio_marg_job** job_out) {
  if (!c || !job_out) return GLIO_ERR_ARG;
  *job_out = nullptr;
  return guarded(c, [&] {
    GLIO_REQUIRE(W >= 2 && W <= 4096 && poses && speed_bias, GLIO_ERR_ARG, "glio_window_marginalize_async: need W >= 2, poses and speed_bias");
    GLIO_REQUIRE(c->marg_pending == nullptr, GLIO_ERR_STATE, "glio_window_marginalize_async: the previous job has not been waited for");
    std::unique_ptr<glio_marg_job> j(new glio_marg_job());
    j->ctx = c; j->W = W; j->eps = eps;
    j->x0_pose.assign(poses + 7, poses + 7 * (size_t)W); j->x0_sb.assign(speed_bias + 9, speed_bias + 18);
    j->epoch = marg_begin(c, W, poses, speed_bias, host_marg, user, j->A, j->b);
#if defined(__linux__)
    j->submit_cpu = sched_getcpu();
#endif
    if (!c->marg_thread.joinable()) c->marg_thread = std::thread(marg_worker, c);
    { std::lock_guard<std::mutex> lk(c->marg_mu); c->marg_queued = j.get(); }
    c->marg_cv.notify_one();
    c->marg_pending = j.get();
    *job_out = j.release();
  });
}

int glio_marg_job_wait(glio_marg_job* j, glio_marg_prior** prior_out) {
  if (!j) return GLIO_ERR_ARG;
  if (prior_out) *prior_out = nullptr;
  for (unsigned long long spins = 0; !j->done.load(std::memory_order_acquire); ++spins) {
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
    if ((spins & 0xfffull) == 0xfffull) std::this_thread::yield();
  }
  glio_ctx* c = j->ctx;
  if (c && c->marg_pending == j) c->marg_pending = nullptr;
  const int rc = j->rc;
  if (rc != GLIO_OK) { if (c) c->err = j->err; if (j->prior) glio_marg_prior_destroy(j->prior); }
  else if (prior_out) *prior_out = j->prior;
  else if (j->prior) glio_marg_prior_destroy(j->prior);
  delete j;
  return rc;
}

// ---- front-end feature extraction: Preprocessing::cloudHandler, GLIO/src/Preprocessing.cpp:529-655
int glio_extract_features(glio_ctx* c, const float* cloud_xyzi, int64_t n, int stride, int intensity_offset, int mem, int n_scans, const int32_t* scan_start,
                          const int32_t* scan_end, int ds_rate, double edge_thres, double surf_thres, float ds_v, float* curvature, int8_t* label,
                          int32_t* sharp, int64_t* n_sharp, int32_t* less_sharp, int64_t* n_less_sharp, int32_t* flat, int64_t* n_flat,
                          int32_t* less_flat, int64_t* n_less_flat, float* less_flat_ds, int64_t* n_less_flat_ds) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(cloud_xyzi && n > 10 && n < ((int64_t)1 << 30), GLIO_ERR_ARG, "glio_extract_features: bad cloud");
    GLIO_REQUIRE(stride >= 4 && stride <= 64 && intensity_offset >= 3 && intensity_offset < stride, GLIO_ERR_ARG, "glio_extract_features: stride_floats must be >= 4 and 3 <= intensity_offset < stride");
    GLIO_REQUIRE(n_scans > 0 && n_scans <= 1024 && scan_start && scan_end && ds_rate >= 1 && ds_v > 0.f, GLIO_ERR_ARG, "glio_extract_features: bad ring table");
    for (int r = 0; r < n_scans; ++r) GLIO_REQUIRE(scan_start[r] >= 5 && scan_end[r] < n - 5 && scan_end[r] - scan_start[r] < n, GLIO_ERR_ARG, "glio_extract_features: ring range outside [5, n-6]");
    FeatArgs a{};
    const float* d_cloud = cloud_xyzi;
    if (mem != GLIO_DEVICE) { c->f_cloud.reserve((size_t)n * stride); GLIO_CUDA_TRY(cudaMemcpyAsync(c->f_cloud.p, cloud_xyzi, (size_t)n * stride * sizeof(float), cudaMemcpyHostToDevice, c->st)); d_cloud = c->f_cloud.p; }
    c->f_curv.reserve(n); c->f_label.reserve(n); c->f_picked.reserve(n); c->f_scan.reserve((size_t)2 * n_scans);
    c->f_ring.reserve((size_t)n_scans * (FEAT_MAX_SHARP + FEAT_MAX_LESS_SHARP + FEAT_MAX_FLAT)); c->f_lf.reserve(n); c->f_ds.reserve(n);
    c->f_counts.reserve((size_t)5 * n_scans + (size_t)5 * (n_scans + 1) + 4);
    c->f_out.reserve((size_t)n_scans * (FEAT_MAX_SHARP + FEAT_MAX_LESS_SHARP + FEAT_MAX_FLAT) + n); c->f_outds.reserve(n);
    GLIO_CUDA_TRY(cudaMemcpyAsync(c->f_scan.p, scan_start, n_scans * sizeof(int32_t), cudaMemcpyHostToDevice, c->st));
    GLIO_CUDA_TRY(cudaMemcpyAsync(c->f_scan.p + n_scans, scan_end, n_scans * sizeof(int32_t), cudaMemcpyHostToDevice, c->st));
    a.cloud = d_cloud; a.stride = stride; a.ioff = intensity_offset; a.n = n; a.n_scans = n_scans; a.scan_start = c->f_scan.p; a.scan_end = c->f_scan.p + n_scans; a.ds_rate = ds_rate;
    a.edge_thres = edge_thres; a.surf_thres = surf_thres; a.ds_v = ds_v;
    a.curv = c->f_curv.p; a.label = c->f_label.p; a.picked = c->f_picked.p;
    a.ring_sharp = c->f_ring.p; a.ring_less_sharp = a.ring_sharp + (size_t)n_scans * FEAT_MAX_SHARP; a.ring_flat = a.ring_less_sharp + (size_t)n_scans * FEAT_MAX_LESS_SHARP;
    a.ring_less_flat = c->f_lf.p; a.ring_ds = c->f_ds.p;
    a.counts = c->f_counts.p; a.offsets = c->f_counts.p + 5 * n_scans; a.err = c->f_counts.p + 5 * n_scans + 5 * (n_scans + 1);
    a.out_sharp = c->f_out.p; a.out_less_sharp = a.out_sharp + (size_t)n_scans * FEAT_MAX_SHARP; a.out_flat = a.out_less_sharp + (size_t)n_scans * FEAT_MAX_LESS_SHARP;
    a.out_less_flat = a.out_flat + (size_t)n_scans * FEAT_MAX_FLAT; a.out_ds = c->f_outds.p;
    features_run(a, c->st, c->lc);
    std::vector<int32_t> tot((size_t)5 * (n_scans + 1) + 1);
    GLIO_CUDA_TRY(cudaMemcpyAsync(tot.data(), a.offsets, tot.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    const int32_t err = tot[(size_t)5 * (n_scans + 1)];
    GLIO_REQUIRE(err == 0, GLIO_ERR_ARG, err == 1 ? "glio_extract_features: a ring sector has more than 4096 points" : "glio_extract_features: a ring has more than 8192 less-flat points");
    const int64_t cnt[5] = {tot[0 * (n_scans + 1) + n_scans], tot[1 * (n_scans + 1) + n_scans], tot[2 * (n_scans + 1) + n_scans], tot[3 * (n_scans + 1) + n_scans], tot[4 * (n_scans + 1) + n_scans]};
    if (n_sharp) *n_sharp = cnt[0];
    if (n_less_sharp) *n_less_sharp = cnt[1];
    if (n_flat) *n_flat = cnt[2];
    if (n_less_flat) *n_less_flat = cnt[3];
    if (n_less_flat_ds) *n_less_flat_ds = cnt[4];
    if (curvature) GLIO_CUDA_TRY(cudaMemcpyAsync(curvature, a.curv, n * sizeof(float), cudaMemcpyDeviceToHost, c->st));
    if (label) GLIO_CUDA_TRY(cudaMemcpyAsync(label, a.label, n, cudaMemcpyDeviceToHost, c->st));
    if (sharp && cnt[0]) GLIO_CUDA_TRY(cudaMemcpyAsync(sharp, a.out_sharp, cnt[0] * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
    if (less_sharp && cnt[1]) GLIO_CUDA_TRY(cudaMemcpyAsync(less_sharp, a.out_less_sharp, cnt[1] * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
    if (flat && cnt[2]) GLIO_CUDA_TRY(cudaMemcpyAsync(flat, a.out_flat, cnt[2] * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
    if (less_flat && cnt[3]) GLIO_CUDA_TRY(cudaMemcpyAsync(less_flat, a.out_less_flat, cnt[3] * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
    if (less_flat_ds && cnt[4]) GLIO_CUDA_TRY(cudaMemcpyAsync(less_flat_ds, a.out_ds, cnt[4] * sizeof(float4), cudaMemcpyDeviceToHost, c->st));
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
  });
}

void glio_default_solver_options(glio_solver_options* o) {
  if (!o) return;
  memset(o, 0, sizeof(*o));
  o->max_num_iterations = 15; o->dogleg_type = 0; o->use_nonmonotonic_steps = 0; o->max_consecutive_nonmonotonic_steps = 5;
  o->initial_trust_region_radius = 1e4; o->max_trust_region_radius = 1e16; o->min_trust_region_radius = 1e-32;
  o->min_relative_decrease = 1e-3; o->min_lm_diagonal = 1e-6; o->max_lm_diagonal = 1e32; o->max_num_consecutive_invalid_steps = 5;
  o->jacobi_scaling = 1; o->function_tolerance = 1e-6; o->gradient_tolerance = 1e-10; o->parameter_tolerance = 1e-8;
  o->fuse_candidate_jacobian = 1;
}

static int window_solve_impl(glio_ctx* c, int W, double* poses, double* speed_bias, glio_host_factors_fn host_factors,
                             glio_host_factors_band_fn host_band, int hb_hint, void* user,
                             const glio_solver_options* options, glio_solver_summary* summary, glio_iteration* iter_log, int iter_cap,
                             double* step_log, int64_t step_cap) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(W > 0 && W <= 4096 && poses, GLIO_ERR_ARG, "bad arguments");
    glio_solver_options o; if (options) o = *options; else glio_default_solver_options(&o);
    const bool sb = speed_bias != nullptr;
    const int nt = sb ? 15 : 6, na = sb ? 16 : 7, n = W * nt;
    std::vector<ParamBlock> blocks;
    for (int k = 0; k < W; ++k) {
      blocks.push_back(ParamBlock{na * k, 3, nt * k, 3, false, nullptr});
      blocks.push_back(ParamBlock{na * k + 3, 4, nt * k + 3, 3, true, nullptr});
      if (sb) blocks.push_back(ParamBlock{na * k + 7, 9, nt * k + 6, 9, false, nullptr});
    }
    SolverOptions so;
    so.max_num_iterations = o.max_num_iterations; so.dogleg_type = o.dogleg_type; so.use_nonmonotonic_steps = o.use_nonmonotonic_steps != 0;
    so.max_consecutive_nonmonotonic_steps = o.max_consecutive_nonmonotonic_steps;
    so.initial_trust_region_radius = o.initial_trust_region_radius; so.max_trust_region_radius = o.max_trust_region_radius;
    so.min_trust_region_radius = o.min_trust_region_radius; so.min_relative_decrease = o.min_relative_decrease;
    so.min_lm_diagonal = o.min_lm_diagonal; so.max_lm_diagonal = o.max_lm_diagonal;
    so.max_num_consecutive_invalid_steps = o.max_num_consecutive_invalid_steps; so.jacobi_scaling = o.jacobi_scaling != 0;
    so.function_tolerance = o.function_tolerance; so.gradient_tolerance = o.gradient_tolerance; so.parameter_tolerance = o.parameter_tolerance;
    so.fuse_candidate_jacobian = o.fuse_candidate_jacobian != 0;
    GLIO_REQUIRE(o.trust_region_strategy == 0 || o.trust_region_strategy == 1, GLIO_ERR_ARG, "trust_region_strategy must be 0 (DOGLEG) or 1 (LEVENBERG_MARQUARDT)");
    so.trust_region_strategy = o.trust_region_strategy;
    TrustRegionDogleg solver(blocks, so);
    std::vector<double> x((size_t)W * na), pz((size_t)W * 7), sz((size_t)W * 9);
    for (int k = 0; k < W; ++k) {
      for (int i = 0; i < 7; ++i) x[(size_t)na * k + i] = poses[7 * k + i];
      if (sb) for (int i = 0; i < 9; ++i) x[(size_t)na * k + 7 + i] = speed_bias[9 * k + i];
    }
    std::vector<double> Hd;       // dense scratch for the host callback's n x n interface
    int hb_fixed = -1;
    EvalFn eval = [&](const double* xa, bool want_jac, double* cost, BandMat* H, double* g) -> bool {
      for (int k = 0; k < W; ++k) {
        for (int i = 0; i < 7; ++i) pz[7 * k + i] = xa[(size_t)na * k + i];
        if (sb) for (int i = 0; i < 9; ++i) sz[9 * k + i] = xa[(size_t)na * k + 7 + i];
      }
      // the device evaluates the LiDAR residuals while the host evaluates its own factors
      const unsigned int epoch = eval_unary_launch(c, W, pz.data(), 0, want_jac);
      double ct = 0;
      if (want_jac) std::fill(g, g + n, 0.0);
      auto add_device = [&]() {
        eval_unary_wait(c, epoch);
        for (int k = 0; k < W; ++k) ct += c->h_out.p[(size_t)k * GLIO_NACC + 27];
        if (want_jac) for (int k = 0; k < W; ++k) for (int p = 0; p < 6; ++p) g[nt * k + p] += c->h_out.p[(size_t)k * GLIO_NACC + 21 + p];
      };
      if (host_band) {
        // host factors accumulate straight into the band matrix (no dense n x n scratch)
        if (want_jac) H->reset(n, hb_hint >= 5 ? hb_hint : n - 1);
        const int hrc = host_band(user, W, pz.data(), sb ? sz.data() : nullptr, want_jac ? 1 : 0, want_jac ? H->a.data() : nullptr, want_jac ? H->hb : 0, g, &ct);
        add_device();
        if (hrc != 0) return false;
        if (want_jac) for (int k = 0; k < W; ++k) {
          const double* ob = c->h_out.p + (size_t)k * GLIO_NACC;
          int idx = 0;
          for (int p = 0; p < 6; ++p) for (int q = p; q < 6; ++q) { H->at(nt * k + q, nt * k + p) += ob[idx]; ++idx; }
        }
        *cost = ct;
        return std::isfinite(ct);
      }
      int hrc = 0;
      if (host_factors) {
        if (want_jac) Hd.assign((size_t)n * n, 0.0);
        hrc = host_factors(user, W, pz.data(), sb ? sz.data() : nullptr, want_jac ? 1 : 0, want_jac ? Hd.data() : nullptr, g, &ct);
      }
      add_device();
      if (hrc != 0) return false;
      if (want_jac) {
        // band of J^T J from what the callback actually wrote, re-derived at EVERY Jacobian evaluation (an entry that turns
        // non-zero later must not be dropped); the LiDAR blocks need 5
        {
          int w = 5;
          if (host_factors) for (int i = 0; i < n; ++i) for (int j = 0; j < i - w; ++j) if (Hd[(size_t)i * n + j] != 0.0 || Hd[(size_t)j * n + i] != 0.0) { w = i - j; break; }
          hb_fixed = std::max(hb_fixed, w);
        }
        H->reset(n, hb_fixed);
        const int hb = H->hb;
        if (host_factors) for (int i = 0; i < n; ++i) for (int j = std::max(0, i - hb); j <= i; ++j) H->at(i, j) = Hd[(size_t)i * n + j];
        for (int k = 0; k < W; ++k) {
          const double* ob = c->h_out.p + (size_t)k * GLIO_NACC;
          int idx = 0;
          for (int p = 0; p < 6; ++p) for (int q = p; q < 6; ++q) { H->at(nt * k + q, nt * k + p) += ob[idx]; ++idx; }
        }
      }
      *cost = ct;
      return std::isfinite(ct);
    };
    SolverSummary S;
    solver.solve(x.data(), eval, &S);
    for (int k = 0; k < W; ++k) {
      for (int i = 0; i < 7; ++i) poses[7 * k + i] = x[(size_t)na * k + i];
      if (sb) for (int i = 0; i < 9; ++i) speed_bias[9 * k + i] = x[(size_t)na * k + 7 + i];
    }
    fill_summary(S, n, summary, iter_log, iter_cap, step_log, step_cap);
  });
}

int glio_window_solve(glio_ctx* c, int W, double* poses, double* speed_bias, glio_host_factors_fn host_factors, void* user,
                      const glio_solver_options* options, glio_solver_summary* summary, glio_iteration* iter_log, int iter_cap,
                      double* step_log, int64_t step_cap) {
  return window_solve_impl(c, W, poses, speed_bias, host_factors, nullptr, -1, user, options, summary, iter_log, iter_cap, step_log, step_cap);
}

int glio_window_solve_band(glio_ctx* c, int W, double* poses, double* speed_bias, glio_host_factors_band_fn host_factors, int half_bandwidth,
                           void* user, const glio_solver_options* options, glio_solver_summary* summary, glio_iteration* iter_log,
                           int iter_cap, double* step_log, int64_t step_cap) {
  return window_solve_impl(c, W, poses, speed_bias, nullptr, host_factors, half_bandwidth, user, options, summary, iter_log, iter_cap, step_log, step_cap);
}

int glio_eval_unary_residuals(glio_ctx* c, int slot, const double pose_body[7], int jac_kind, int64_t capacity, double* r, double* J, int64_t* n_out) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    Slot& sl = c->slot(slot);
    const bool sel = sl.n_sel >= 0;
    const int64_t n = sel ? sl.n_sel : sl.n_match;
    if (n_out) *n_out = n;
    GLIO_REQUIRE(capacity >= n, GLIO_ERR_ARG, "capacity smaller than the residual count");
    if (n == 0) return;
    c->d_r.reserve(n); c->d_J.reserve(6 * n); c->d_poses.reserve(7); c->h_poses.reserve(7);
    memcpy(c->h_poses.p, pose_body, 7 * sizeof(double));
    GLIO_CUDA_TRY(cudaMemcpyAsync(c->d_poses.p, c->h_poses.p, 7 * sizeof(double), cudaMemcpyHostToDevice, c->st));
    eval_unary_residuals_run(sel ? sl.s_cpw.p : sl.m_cpw.p, sel ? sl.s_nsd.p : sl.m_nsd.p, n, c->d_poses.p, eval_params(c), jac_kind,
                             c->d_r.p, c->d_J.p, c->st, c->lc);
    if (r) GLIO_CUDA_TRY(cudaMemcpyAsync(r, c->d_r.p, n * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    if (J) GLIO_CUDA_TRY(cudaMemcpyAsync(J, c->d_J.p, 6 * n * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
  });
}


// ------------------------------------------------------------------------------------------------------------
// batch path: frames, pair association (K1b), binary evaluation (K2b)
// ------------------------------------------------------------------------------------------------------------
int glio_batch_clear(glio_ctx* c) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    for (auto& f : c->frames) { f.second->scan.release(); f.second->grid.release(); }
    for (auto& p : c->pairs) { p->m_cpw.release(); p->m_nc.release(); p->m_src.release(); p->s_cpw.release(); p->s_nc.release(); }
    c->frames.clear(); c->pairs.clear(); c->pair_index.clear(); c->bin_dirty = true;
  });
}

int glio_batch_set_frame(glio_ctx* c, int frame, const float* scan_xyz, int64_t Q, int stride, int mem, const double pose[7]) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(frame >= 0 && pose, GLIO_ERR_ARG, "bad frame arguments");
    auto& fp = c->frames[frame];
    if (!fp) fp.reset(new glio_ctx::Frame());
    glio_ctx::Frame& f = *fp;
    f.scan_ptr = stage_points(c, f.scan, scan_xyz, Q, stride, mem);
    f.Q = Q; f.stride = stride; f.grid_valid = false;
    for (int i = 0; i < 7; ++i) f.pose[i] = pose[i];
  });
}

// pose-only update (association always uses the poses registered here — the reference uses pose_info_keyframe, quirk Q8)
int glio_batch_set_pose(glio_ctx* c, int frame, const double pose[7]) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    auto it = c->frames.find(frame);
    GLIO_REQUIRE(it != c->frames.end() && pose, GLIO_ERR_ARG, "unknown frame");
    for (int i = 0; i < 7; ++i) it->second->pose[i] = pose[i];
    it->second->grid_valid = false;
  });
}

int glio_batch_associate_pairs(glio_ctx* c, const int32_t* pairs_cur, const int32_t* pairs_oth, int64_t n_pairs, int64_t* n_match) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(n_pairs > 0 && pairs_cur && pairs_oth, GLIO_ERR_ARG, "bad pair list");
    // group the requested pairs by the searched frame: one grid, one launch per group
    std::map<int, std::vector<int64_t>> by_oth;
    for (int64_t i = 0; i < n_pairs; ++i) {
      GLIO_REQUIRE(c->frames.count(pairs_cur[i]) && c->frames.count(pairs_oth[i]) && pairs_cur[i] != pairs_oth[i], GLIO_ERR_ARG, "pair references an unknown frame");
      by_oth[pairs_oth[i]].push_back(i);
    }
    AssocGates gates{c->prm.batch_max_radius, c->prm.batch_dist_thres, c->prm.weight_min};
    for (auto& grp : by_oth) {
      glio_ctx::Frame& fo = *c->frames[grp.first];
      if (!fo.grid_valid) {
        // world cloud of the searched frame (body pose applied directly to the stored points, quirk Q7) + its grid
        fo.grid.build_pairs = c->knn_mode == 4;
        grid_build(fo.grid, fo.scan_ptr, fo.stride, fo.Q, fo.pose, fo.pose + 3, c->prm.cell_size, c->pts_per_cell, c->st, c->lc);
        fo.grid_valid = true;
      }
      const int nseg = (int)grp.second.size();
      std::vector<SegDesc> segs(nseg);
      int64_t Qt = 0;
      for (int s = 0; s < nseg; ++s) {
        glio_ctx::Frame& fc = *c->frames[pairs_cur[grp.second[s]]];
        segs[s].src = fc.scan_ptr; segs[s].stride = fc.stride; segs[s].count = fc.Q; segs[s].offset = Qt;
        for (int k = 0; k < 3; ++k) segs[s].t[k] = fc.pose[k];
        for (int k = 0; k < 4; ++k) segs[s].q[k] = fc.pose[3 + k];
        Qt += fc.Q;
      }
      ensure_work(c, Qt, true);
      c->d_segs.reserve(nseg); c->d_dst.reserve(nseg); c->d_counts.reserve(nseg); c->h_counts.reserve(nseg);
      GLIO_CUDA_TRY(cudaMemcpyAsync(c->d_segs.p, segs.data(), nseg * sizeof(SegDesc), cudaMemcpyHostToDevice, c->st));
      AssocWork w = make_work(c, Qt, true);
      assoc_run(fo.grid, c->d_segs.p, nseg, w, gates, fo.scan_ptr, fo.stride, c->cell_count, c->cell_pos, c->scan_tmp, c->st, c->lc);
      compact_count(w, c->d_segs.p, nseg, c->w_flags.p, c->w_pos.p, c->scan_tmp, c->d_counts.p, c->st, c->lc);
      GLIO_CUDA_TRY(cudaMemcpyAsync(c->h_counts.p, c->d_counts.p, nseg * sizeof(int), cudaMemcpyDeviceToHost, c->st));
      GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
      std::vector<CompactDst> dst(nseg);
      for (int s = 0; s < nseg; ++s) {
        const int64_t i = grp.second[s];
        const std::pair<int, int> key(pairs_cur[i], pairs_oth[i]);
        auto pit = c->pair_index.find(key);
        int pi;
        if (pit == c->pair_index.end()) { pi = (int)c->pairs.size(); c->pairs.emplace_back(new glio_ctx::Pair()); c->pair_index[key] = pi; }
        else pi = pit->second;
        glio_ctx::Pair& pr = *c->pairs[pi];
        pr.cur = key.first; pr.oth = key.second; pr.n_match = c->h_counts.p[s]; pr.n_sel = -1;
        const size_t nm = (size_t)std::max<int64_t>(pr.n_match, 1);
        pr.m_cpw.reserve(nm); pr.m_nc.reserve(6 * nm); pr.m_src.reserve(nm);
        dst[s].cpw = pr.m_cpw.p; dst[s].nsd = nullptr; dst[s].nc = pr.m_nc.p; dst[s].src = pr.m_src.p;
        if (n_match) n_match[i] = pr.n_match;
      }
      GLIO_CUDA_TRY(cudaMemcpyAsync(c->d_dst.p, dst.data(), nseg * sizeof(CompactDst), cudaMemcpyHostToDevice, c->st));
      compact_scatter(w, c->d_segs.p, nseg, c->w_pos.p, c->d_dst.p, c->st, c->lc);
      GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));   // dst/segs are stack vectors
    }
    c->bin_dirty = true;
  });
}

int glio_batch_associate(glio_ctx* c, int cur, const int32_t* oth, int n_oth, int64_t* n_match) {
  if (!c || !oth || n_oth <= 0) return GLIO_ERR_ARG;
  std::vector<int32_t> cc(n_oth, cur);
  return glio_batch_associate_pairs(c, cc.data(), oth, n_oth, n_match);
}

int glio_batch_pair_list(glio_ctx* c, int64_t capacity, int32_t* cur, int32_t* oth, int64_t* n_pairs) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    const int64_t n = (int64_t)c->pairs.size();
    if (n_pairs) *n_pairs = n;
    GLIO_REQUIRE(capacity >= n || (!cur && !oth), GLIO_ERR_ARG, "capacity smaller than the pair count");
    for (int64_t i = 0; i < n; ++i) { if (cur) cur[i] = c->pairs[i]->cur; if (oth) oth[i] = c->pairs[i]->oth; }
  });
}

int glio_batch_get_matches(glio_ctx* c, int cur, int oth, int64_t capacity, float* cp, float* weight, double* normal_cent, int32_t* src, int64_t* n_out) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    auto it = c->pair_index.find(std::make_pair(cur, oth));
    GLIO_REQUIRE(it != c->pair_index.end(), GLIO_ERR_ARG, "pair was not associated");
    glio_ctx::Pair& pr = *c->pairs[it->second];
    const int64_t n = pr.n_match;
    if (n_out) *n_out = n;
    GLIO_REQUIRE(capacity >= n, GLIO_ERR_ARG, "capacity smaller than the match count");
    if (n == 0) return;
    std::vector<float4> h;
    if (cp || weight) { h.resize(n); GLIO_CUDA_TRY(cudaMemcpyAsync(h.data(), pr.m_cpw.p, n * sizeof(float4), cudaMemcpyDeviceToHost, c->st)); }
    if (normal_cent) GLIO_CUDA_TRY(cudaMemcpyAsync(normal_cent, pr.m_nc.p, 6 * n * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    if (src) GLIO_CUDA_TRY(cudaMemcpyAsync(src, pr.m_src.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, c->st));
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    for (int64_t i = 0; i < n && !h.empty(); ++i) {
      if (cp) { cp[3 * i] = h[i].x; cp[3 * i + 1] = h[i].y; cp[3 * i + 2] = h[i].z; }
      if (weight) weight[i] = h[i].w;
    }
  });
}

int glio_batch_select(glio_ctx* c, int cur, int oth, const int32_t* keep, int64_t n) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    auto it = c->pair_index.find(std::make_pair(cur, oth));
    GLIO_REQUIRE(it != c->pair_index.end(), GLIO_ERR_ARG, "pair was not associated");
    glio_ctx::Pair& pr = *c->pairs[it->second];
    c->bin_dirty = true;
    if (n < 0) { pr.n_sel = -1; return; }
    GLIO_REQUIRE(n == 0 || keep, GLIO_ERR_ARG, "null selection list");
    pr.n_sel = n;
    if (n == 0) return;
    c->d_keep.reserve(n); c->d_bad.reserve(1); pr.s_cpw.reserve(n); pr.s_nc.reserve(6 * n);
    GLIO_CUDA_TRY(cudaMemcpyAsync(c->d_keep.p, keep, n * sizeof(int32_t), cudaMemcpyHostToDevice, c->st));
    GLIO_CUDA_TRY(cudaMemsetAsync(c->d_bad.p, 0, sizeof(int), c->st));
    gather_selection(c->d_keep.p, n, pr.n_match, pr.m_cpw.p, nullptr, pr.m_nc.p, pr.s_cpw.p, nullptr, pr.s_nc.p, c->d_bad.p, c->st, c->lc);
    int bad = 0;
    GLIO_CUDA_TRY(cudaMemcpyAsync(&bad, c->d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    if (bad) { pr.n_sel = -1; throw Error{GLIO_ERR_ARG, "selection index out of range"}; }
  });
}

}  // extern "C"

namespace {

// work items + incidence lists of the binary evaluation (rebuilt when associations / selections change)
void build_bin_items(glio_ctx* c, int K) {
  if (!c->bin_dirty && c->bin_K == K) return;
  const int P = (int)c->pairs.size();
  std::vector<BinItem> items;
  std::vector<int> pstart(P + 1, 0);
  std::vector<std::vector<BinIncidence>> inc(K);
  for (int p = 0; p < P; ++p) {
    glio_ctx::Pair& pr = *c->pairs[p];
    GLIO_REQUIRE(pr.cur < K && pr.oth < K, GLIO_ERR_ARG, "pair references a keyframe >= K");
    pstart[p] = (int)items.size();
    const bool sel = pr.n_sel >= 0;
    const int64_t n = sel ? pr.n_sel : pr.n_match;
    const float4* cpw = sel ? pr.s_cpw.p : pr.m_cpw.p;
    const double* nc = sel ? pr.s_nc.p : pr.m_nc.p;
    for (int64_t o = 0; o < n; o += GLIO_ITEM_MAX) {
      BinItem it; it.cpw = cpw + o; it.nc = nc + 6 * o; it.count = (int32_t)std::min<int64_t>(GLIO_ITEM_MAX, n - o); it.kf_c = pr.cur; it.kf_o = pr.oth;
      items.push_back(it);
    }
    inc[pr.cur].push_back(BinIncidence{p, 0});
    inc[pr.oth].push_back(BinIncidence{p, 1});
  }
  pstart[P] = (int)items.size();
  std::vector<int> kstart(K + 1, 0);
  std::vector<BinIncidence> flat;
  for (int k = 0; k < K; ++k) { kstart[k] = (int)flat.size(); flat.insert(flat.end(), inc[k].begin(), inc[k].end()); }
  kstart[K] = (int)flat.size();
  c->n_bin_items = (int)items.size();
  c->d_bin_items.reserve(items.size() + 1); c->d_pair_item_start.reserve(P + 1); c->d_kf_inc_start.reserve(K + 1); c->d_inc.reserve(flat.size() + 1);
  c->d_bin_partials.reserve((items.size() + 1) * GLIO_NACC_BIN); c->d_pair_sums.reserve((size_t)(P + 1) * GLIO_NACC_BIN);
  c->d_bin_diag.reserve((size_t)K * GLIO_NACC); c->d_bin_off.reserve((size_t)(P + 1) * 36); c->d_bin_poses.reserve((size_t)K * 7);
  c->h_bin_diag.reserve((size_t)K * GLIO_NACC); c->h_bin_off.reserve((size_t)(P + 1) * 36); c->h_bin_poses.reserve((size_t)K * 7);
  if (!items.empty()) GLIO_CUDA_TRY(cudaMemcpyAsync(c->d_bin_items.p, items.data(), items.size() * sizeof(BinItem), cudaMemcpyHostToDevice, c->st));
  GLIO_CUDA_TRY(cudaMemcpyAsync(c->d_pair_item_start.p, pstart.data(), (P + 1) * sizeof(int), cudaMemcpyHostToDevice, c->st));
  GLIO_CUDA_TRY(cudaMemcpyAsync(c->d_kf_inc_start.p, kstart.data(), (K + 1) * sizeof(int), cudaMemcpyHostToDevice, c->st));
  if (!flat.empty()) GLIO_CUDA_TRY(cudaMemcpyAsync(c->d_inc.p, flat.data(), flat.size() * sizeof(BinIncidence), cudaMemcpyHostToDevice, c->st));
  GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
  c->bin_dirty = false; c->bin_K = K;
}

// device evaluation of all active binary residuals -> c->h_bin_diag (K x 28), c->h_bin_off (P x 36): everything is queued
// (kernels, the all-reduce of a sharded run, the copies to pinned host memory); eval_binary_wait makes the result visible
void eval_binary_launch(glio_ctx* c, int K, const double* poses, bool want_jac, double huber) {
  build_bin_items(c, K);
  const int P = (int)c->pairs.size();
  memcpy(c->h_bin_poses.p, poses, (size_t)K * 7 * sizeof(double));
  GLIO_CUDA_TRY(cudaMemcpyAsync(c->d_bin_poses.p, c->h_bin_poses.p, (size_t)K * 7 * sizeof(double), cudaMemcpyHostToDevice, c->st));
  eval_binary_run(c->d_bin_items.p, c->n_bin_items, c->d_pair_item_start.p, P, K, c->d_bin_poses.p, c->prm.batch_score, huber, want_jac,
                  c->d_bin_partials.p, c->d_pair_sums.p, c->d_kf_inc_start.p, c->d_inc.p, c->d_bin_diag.p, c->d_bin_off.p, nullptr, c->st, c->lc);
  if (c->allreduce) {
    // keyframe-sharded evaluation: every rank holds the same pair list but only its own pairs carry residuals;
    // one sum over ranks of the pose-block buffers per evaluation (SURVEY 8e "pose-block allreduce")
    GLIO_REQUIRE(c->allreduce(c->allreduce_user, c->d_bin_diag.p, (int64_t)K * GLIO_NACC, (void*)c->st) == 0, GLIO_ERR_NCCL, "allreduce hook failed (diag)");
    if (want_jac && P > 0)
      GLIO_REQUIRE(c->allreduce(c->allreduce_user, c->d_bin_off.p, (int64_t)P * 36, (void*)c->st) == 0, GLIO_ERR_NCCL, "allreduce hook failed (off)");
  }
  GLIO_CUDA_TRY(cudaMemcpyAsync(c->h_bin_diag.p, c->d_bin_diag.p, (size_t)K * GLIO_NACC * sizeof(double), cudaMemcpyDeviceToHost, c->st));
  if (want_jac && P > 0) GLIO_CUDA_TRY(cudaMemcpyAsync(c->h_bin_off.p, c->d_bin_off.p, (size_t)P * 36 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
}
void eval_binary_wait(glio_ctx* c) { GLIO_CUDA_TRY(cudaStreamSynchronize(c->st)); }
void eval_binary_blocks(glio_ctx* c, int K, const double* poses, bool want_jac, double huber) {
  eval_binary_launch(c, K, poses, want_jac, huber);
  eval_binary_wait(c);
}

}  // namespace

extern "C" {

int glio_eval_binary(glio_ctx* c, int K, const double* poses, double* Hdiag, double* Hoff, double* g, double* cost) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(K > 0 && poses, GLIO_ERR_ARG, "bad arguments");
    const bool want_jac = Hdiag || Hoff || g;
    eval_binary_blocks(c, K, poses, want_jac, /*no loss on batch LiDAR factors, Estimator.cpp:2768*/ 0.0);
    double ct = 0;
    for (int k = 0; k < K; ++k) {
      const double* o = c->h_bin_diag.p + (size_t)k * GLIO_NACC;
      if (Hdiag) { int idx = 0; for (int p = 0; p < 6; ++p) for (int q = p; q < 6; ++q) { Hdiag[36 * k + 6 * p + q] = o[idx]; Hdiag[36 * k + 6 * q + p] = o[idx]; ++idx; } }
      if (g) for (int p = 0; p < 6; ++p) g[6 * k + p] = o[21 + p];
      ct += o[27];
    }
    if (Hoff) memcpy(Hoff, c->h_bin_off.p, c->pairs.size() * 36 * sizeof(double));
    if (cost) *cost = ct;
  });
}


int glio_set_allreduce(glio_ctx* c, glio_allreduce_fn fn, void* user) {
  if (!c) return GLIO_ERR_ARG;
  c->allreduce = fn; c->allreduce_user = user;
  return GLIO_OK;
}

int glio_batch_declare_pairs(glio_ctx* c, const int32_t* pairs_cur, const int32_t* pairs_oth, int64_t n_pairs) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    for (int64_t i = 0; i < n_pairs; ++i) {
      const std::pair<int, int> key(pairs_cur[i], pairs_oth[i]);
      if (c->pair_index.count(key)) continue;
      c->pair_index[key] = (int)c->pairs.size();
      c->pairs.emplace_back(new glio_ctx::Pair());
      c->pairs.back()->cur = key.first; c->pairs.back()->oth = key.second;
    }
    c->bin_dirty = true;
  });
}

int glio_batch_set_pair_matches(glio_ctx* c, int cur, int oth, const float* cp, const double* normal_cent, const float* weight, int64_t n) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(cur >= 0 && oth >= 0 && cur != oth, GLIO_ERR_ARG, "bad pair");
    GLIO_REQUIRE(n >= 0 && (n == 0 || (cp && normal_cent && weight)), GLIO_ERR_ARG, "null match arrays");
    const std::pair<int, int> key(cur, oth);
    auto it = c->pair_index.find(key);
    int pi;
    if (it == c->pair_index.end()) {
      pi = (int)c->pairs.size(); c->pair_index[key] = pi;
      c->pairs.emplace_back(new glio_ctx::Pair()); c->pairs.back()->cur = cur; c->pairs.back()->oth = oth;
    } else pi = it->second;
    glio_ctx::Pair& pr = *c->pairs[pi];
    pr.n_match = n; pr.n_sel = -1;
    if (n > 0) {
      std::vector<float> cpw((size_t)n * 4);
      std::vector<int32_t> src((size_t)n);
      for (int64_t i = 0; i < n; ++i) { cpw[4 * i] = cp[3 * i]; cpw[4 * i + 1] = cp[3 * i + 1]; cpw[4 * i + 2] = cp[3 * i + 2]; cpw[4 * i + 3] = weight[i]; src[i] = (int32_t)i; }
      pr.m_cpw.reserve((size_t)n); pr.m_nc.reserve((size_t)n * 6); pr.m_src.reserve((size_t)n);
      GLIO_CUDA_TRY(cudaMemcpyAsync(pr.m_cpw.p, cpw.data(), (size_t)n * sizeof(float4), cudaMemcpyHostToDevice, c->st));
      GLIO_CUDA_TRY(cudaMemcpyAsync(pr.m_nc.p, normal_cent, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice, c->st));
      GLIO_CUDA_TRY(cudaMemcpyAsync(pr.m_src.p, src.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, c->st));
      GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    }
    c->bin_dirty = true;
  });
}

int glio_batch_solve(glio_ctx* c, int K, double* poses, double* speed_bias, glio_host_factors_band_fn host_factors, void* user,
                     const glio_solver_options* options, glio_solver_summary* summary, glio_iteration* iter_log, int iter_cap,
                     double* step_log, int64_t step_cap) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(K > 0 && poses, GLIO_ERR_ARG, "bad arguments");
    glio_solver_options o;
    if (options) o = *options;
    else { glio_default_solver_options(&o); o.dogleg_type = 1; o.use_nonmonotonic_steps = 1; o.max_num_iterations = 100; }  // Estimator.cpp:3276-3281
    const bool sb = speed_bias != nullptr;
    const int nt = sb ? 15 : 6, na = sb ? 16 : 7, n = K * nt;
    int span = 1;
    for (auto& p : c->pairs) span = std::max(span, std::abs(p->cur - p->oth));
    const int hb = (span + 1) * nt - 1;
    std::vector<ParamBlock> blocks;
    for (int k = 0; k < K; ++k) {
      blocks.push_back(ParamBlock{na * k, 3, nt * k, 3, false, nullptr});
      blocks.push_back(ParamBlock{na * k + 3, 4, nt * k + 3, 3, true, nullptr});
      if (sb) blocks.push_back(ParamBlock{na * k + 7, 9, nt * k + 6, 9, false, nullptr});
    }
    SolverOptions so;
    so.max_num_iterations = o.max_num_iterations; so.dogleg_type = o.dogleg_type; so.use_nonmonotonic_steps = o.use_nonmonotonic_steps != 0;
    so.max_consecutive_nonmonotonic_steps = o.max_consecutive_nonmonotonic_steps;
    so.initial_trust_region_radius = o.initial_trust_region_radius; so.max_trust_region_radius = o.max_trust_region_radius;
    so.min_trust_region_radius = o.min_trust_region_radius; so.min_relative_decrease = o.min_relative_decrease;
    so.min_lm_diagonal = o.min_lm_diagonal; so.max_lm_diagonal = o.max_lm_diagonal;
    so.max_num_consecutive_invalid_steps = o.max_num_consecutive_invalid_steps; so.jacobi_scaling = o.jacobi_scaling != 0;
    so.function_tolerance = o.function_tolerance; so.gradient_tolerance = o.gradient_tolerance; so.parameter_tolerance = o.parameter_tolerance;
    so.fuse_candidate_jacobian = o.fuse_candidate_jacobian != 0;
    GLIO_REQUIRE(o.trust_region_strategy == 0 || o.trust_region_strategy == 1, GLIO_ERR_ARG, "trust_region_strategy must be 0 (DOGLEG) or 1 (LEVENBERG_MARQUARDT)");
    so.trust_region_strategy = o.trust_region_strategy;
    TrustRegionDogleg solver(blocks, so);
    std::vector<double> x((size_t)K * na), pz((size_t)K * 7), sz((size_t)K * 9);
    for (int k = 0; k < K; ++k) {
      for (int i = 0; i < 7; ++i) x[(size_t)na * k + i] = poses[7 * k + i];
      if (sb) for (int i = 0; i < 9; ++i) x[(size_t)na * k + 7 + i] = speed_bias[9 * k + i];
    }
    EvalFn eval = [&](const double* xa, bool want_jac, double* cost, BandMat* H, double* g) -> bool {
      for (int k = 0; k < K; ++k) {
        for (int i = 0; i < 7; ++i) pz[7 * k + i] = xa[(size_t)na * k + i];
        if (sb) for (int i = 0; i < 9; ++i) sz[9 * k + i] = xa[(size_t)na * k + 7 + i];
      }
      // the device evaluates the LiDAR pairs (and a sharded run sums them over the ranks) while the host evaluates its own
      // factors into the freshly reset band; the device blocks are added when they have arrived
      eval_binary_launch(c, K, pz.data(), want_jac, 0.0);
      double ct = 0;
      int hrc = 0;
      if (want_jac) { H->reset(n, hb); std::fill(g, g + n, 0.0); }
      if (host_factors) hrc = host_factors(user, K, pz.data(), sb ? sz.data() : nullptr, want_jac ? 1 : 0, want_jac ? H->a.data() : nullptr, want_jac ? H->hb : 0, g, &ct);
      eval_binary_wait(c);
      if (hrc != 0) return false;
      for (int k = 0; k < K; ++k) ct += c->h_bin_diag.p[(size_t)k * GLIO_NACC + 27];
      if (want_jac) {
        const int w = H->hb + 1;
        double* a = H->a.data();
        for (int k = 0; k < K; ++k) {
          const double* ob = c->h_bin_diag.p + (size_t)k * GLIO_NACC;
          int idx = 0;
          for (int p = 0; p < 6; ++p) for (int q = p; q < 6; ++q) { H->at(nt * k + q, nt * k + p) += ob[idx]; ++idx; }
          for (int p = 0; p < 6; ++p) g[nt * k + p] += ob[21 + p];
        }
        for (size_t pi = 0; pi < c->pairs.size(); ++pi) {
          const int kc = c->pairs[pi]->cur, ko = c->pairs[pi]->oth;
          const double* ob = c->h_bin_off.p + pi * 36;                 // d2 cost / d x_cur[p] d x_oth[q]
          if (kc > ko) {                                               // rows of cur, six contiguous band entries each
            for (int p = 0; p < 6; ++p) { double* row = a + (size_t)(nt * kc + p) * w + (nt * ko - (nt * kc + p) + H->hb); for (int q = 0; q < 6; ++q) row[q] += ob[6 * p + q]; }
          } else if (kc < ko) {                                        // stored transposed: rows of oth
            for (int q = 0; q < 6; ++q) { double* row = a + (size_t)(nt * ko + q) * w + (nt * kc - (nt * ko + q) + H->hb); for (int p = 0; p < 6; ++p) row[p] += ob[6 * p + q]; }
          } else {
            for (int p = 0; p < 6; ++p) for (int q = 0; q < 6; ++q) H->add_sym(nt * kc + p, nt * ko + q, ob[6 * p + q]);
          }
        }
      }
      *cost = ct;
      return std::isfinite(ct);
    };
    SolverSummary S;
    solver.solve(x.data(), eval, &S);
    for (int k = 0; k < K; ++k) {
      for (int i = 0; i < 7; ++i) poses[7 * k + i] = x[(size_t)na * k + i];
      if (sb) for (int i = 0; i < 9; ++i) speed_bias[9 * k + i] = x[(size_t)na * k + 7 + i];
    }
    fill_summary(S, n, summary, iter_log, iter_cap, step_log, step_cap);
  });
}


// ---- K2e: edge correspondences are an input (the reference has no edge association)
int glio_set_edges(glio_ctx* c, int slot, const float* cp, const float* pa, const float* pb, const double* s, int64_t n) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    Slot& sl = c->slot(slot);
    sl.n_edge = 0;
    if (n <= 0) return;
    GLIO_REQUIRE(cp && pa && pb && s, GLIO_ERR_ARG, "null edge arrays");
    std::vector<float4> h0(n), h1(n), h2(n);
    for (int64_t i = 0; i < n; ++i) {
      h0[i] = make_float4(cp[3 * i], cp[3 * i + 1], cp[3 * i + 2], (float)s[i]);
      h1[i] = make_float4(pa[3 * i], pa[3 * i + 1], pa[3 * i + 2], 0.f);
      h2[i] = make_float4(pb[3 * i], pb[3 * i + 1], pb[3 * i + 2], 0.f);
    }
    sl.e_cps.reserve(n); sl.e_pa.reserve(n); sl.e_pb.reserve(n);
    GLIO_CUDA_TRY(cudaMemcpyAsync(sl.e_cps.p, h0.data(), n * sizeof(float4), cudaMemcpyHostToDevice, c->st));
    GLIO_CUDA_TRY(cudaMemcpyAsync(sl.e_pa.p, h1.data(), n * sizeof(float4), cudaMemcpyHostToDevice, c->st));
    GLIO_CUDA_TRY(cudaMemcpyAsync(sl.e_pb.p, h2.data(), n * sizeof(float4), cudaMemcpyHostToDevice, c->st));
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    sl.n_edge = n;
  });
}

int glio_eval_edge(glio_ctx* c, int W, const double* poses_body, double* H, double* g, double* cost) {
  if (!c) return GLIO_ERR_ARG;
  return guarded(c, [&] {
    GLIO_REQUIRE(W > 0 && W <= 4096 && poses_body, GLIO_ERR_ARG, "bad arguments");
    std::vector<EdgeItem> items; std::vector<int> start(W + 1, 0);
    for (int k = 0; k < W; ++k) {
      start[k] = (int)items.size();
      if ((size_t)k >= c->slots.size() || !c->slots[k]) continue;
      Slot& sl = *c->slots[k];
      for (int64_t o = 0; o < sl.n_edge; o += GLIO_ITEM_MAX) {
        EdgeItem it; it.cps = sl.e_cps.p + o; it.pa = sl.e_pa.p + o; it.pb = sl.e_pb.p + o;
        it.count = (int32_t)std::min<int64_t>(GLIO_ITEM_MAX, sl.n_edge - o); it.kf = k; items.push_back(it);
      }
    }
    start[W] = (int)items.size();
    DevBuf<EdgeItem>& d_items = c->e_items; DevBuf<int>& d_start = c->e_start; DevBuf<double>& d_part = c->e_part; DevBuf<double>& d_out = c->e_out; DevBuf<double>& d_poses = c->e_poses;
    d_items.reserve(items.size() + 1); d_start.reserve(W + 1); d_part.reserve((items.size() + 1) * GLIO_NACC); d_out.reserve((size_t)W * GLIO_NACC); d_poses.reserve((size_t)W * 7);
    if (!c->d_ticket.p) { c->d_ticket.reserve(1); GLIO_CUDA_TRY(cudaMemsetAsync(c->d_ticket.p, 0, sizeof(unsigned int), c->st)); }
    if (!items.empty()) GLIO_CUDA_TRY(cudaMemcpyAsync(d_items.p, items.data(), items.size() * sizeof(EdgeItem), cudaMemcpyHostToDevice, c->st));
    GLIO_CUDA_TRY(cudaMemcpyAsync(d_start.p, start.data(), (W + 1) * sizeof(int), cudaMemcpyHostToDevice, c->st));
    GLIO_CUDA_TRY(cudaMemcpyAsync(d_poses.p, poses_body, (size_t)W * 7 * sizeof(double), cudaMemcpyHostToDevice, c->st));
    const bool want_jac = H || g;
    eval_edge_run(d_items.p, (int)items.size(), W, d_poses.p, eval_params(c), want_jac, d_part.p, d_out.p, d_start.p, c->d_ticket.p, c->st, c->lc);
    std::vector<double> ho((size_t)W * GLIO_NACC);
    GLIO_CUDA_TRY(cudaMemcpyAsync(ho.data(), d_out.p, ho.size() * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    GLIO_CUDA_TRY(cudaStreamSynchronize(c->st));
    for (int k = 0; k < W; ++k) {
      const double* o = ho.data() + (size_t)k * GLIO_NACC;
      if (H) { int idx = 0; for (int p = 0; p < 6; ++p) for (int q = p; q < 6; ++q) { H[36 * k + 6 * p + q] = o[idx]; H[36 * k + 6 * q + p] = o[idx]; ++idx; } }
      if (g) for (int p = 0; p < 6; ++p) g[6 * k + p] = o[21 + p];
      if (cost) cost[k] = o[27];
    }
  });
}

}  // extern "C"
