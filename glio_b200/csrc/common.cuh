// common.cuh — shared device/host declarations of the glio_b200 CUDA library (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/glio_b200.h"

namespace glio {

// ----------------------------------------------------------------------------------------------
// error plumbing: every CUDA call is checked; failures become a status + message, never abort.
// ----------------------------------------------------------------------------------------------
struct Error {
  int code;
  std::string msg;
};
void set_global_error(const std::string& m);
const char* global_error();

#define GLIO_CUDA_TRY(expr)                                                                   \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      char _b[512];                                                                           \
      snprintf(_b, sizeof(_b), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      throw ::glio::Error{GLIO_ERR_CUDA, _b};                                                 \
    }                                                                                         \
  } while (0)

#define GLIO_REQUIRE(cond, code, text)                      \
  do {                                                      \
    if (!(cond)) throw ::glio::Error{(code), (text)};       \
  } while (0)

// grow-only device buffer
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  void reserve(size_t n) {
    if (n <= cap) return;
    if (p) GLIO_CUDA_TRY(cudaFree(p));
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 64;
    GLIO_CUDA_TRY(cudaMalloc((void**)&p, want * sizeof(T)));
    cap = want;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

template <class T>
struct PinnedBuf {
  T* p = nullptr;
  size_t cap = 0;
  void reserve(size_t n) {
    if (n <= cap) return;
    if (p) GLIO_CUDA_TRY(cudaFreeHost(p));
    p = nullptr; cap = 0;
    size_t want = n + n / 8 + 64;
    GLIO_CUDA_TRY(cudaMallocHost((void**)&p, want * sizeof(T)));
    cap = want;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// ----------------------------------------------------------------------------------------------
// Uniform grid over a point cloud (K0 output).  Points are counting-sorted by cell with x the
// fastest-varying cell coordinate, so the three x-adjacent cells of a (y,z) row are ONE contiguous range.
// pts[k] = (x, y, z, bitcast(original index)).
// ----------------------------------------------------------------------------------------------
// Pair layout of the cell-sorted map for the staged tile search (knn_tile.cu): record k holds sorted points 2k and 2k+1
// as (x0,x1,y0,y1 | z0,z1 | idx0,idx1) so that one 16-byte + one 8-byte shared-memory load feed the packed fp32x2
// distance of two candidates; a row of cells is a contiguous, 32-byte aligned run of records (bulk-copy friendly).
struct __align__(32) PairRec { float x0, x1, y0, y1, z0, z1; int i0, i1; };
constexpr float PAIR_FAR = 1.0e18f;      // coordinate of a filler half-record: its squared distance (1e36) is finite and huge

struct GridDesc {
  float ox, oy, oz;     // origin (min corner of cell (0,0,0))
  float cell, inv_cell;
  int nx, ny, nz;
  int64_t npts;
  const int* cell_start;   // [nx*ny*nz + 1]
  const float4* pts;       // [npts] sorted by cell
};

// One association segment = one scan transformed by one pose, queried against the launch's grid.
struct SegDesc {
  const float* src;     // scan points, stride floats apart
  int stride;
  int64_t count;
  int64_t offset;       // first global query index of the segment
  double t[3];
  double q[4];
};

struct AssocGates {
  double max_radius;    // compared with the squared 5th distance (quirk Q1)
  double dist_thres;
  double weight_min;
};

// workspace views handed to the association kernels (global query index g in [0, Qt))
struct AssocWork {
  int64_t Qt;
  float4* pm;           // transformed query (x,y,z, _)
  float4* pmq;          // the same, cell-sorted order, .w = bitcast scan index (written by k_order_scatter)
  uint16_t* seg;        // segment id of the query
  int32_t* knn_idx;     // [5][Qt] neighbour indices, sorted order (K1a -> K1b)
  float* knn_sqd;       // [5][Qt] squared distances, sorted order
  unsigned long long* n_fallback;  // statistics: queries deferred from the tile pass to the single-query pass
  int tile_rings;       // rings scanned by the tile pass before deferring (>= 32: never defer)
  int grow_mode;        // box growth slabs: see SearchArgs::grow_mode
  int knn_mode;         // 0: warp-cooperative tile pass, 1: per-thread ring growth, 2: per-thread box growth from 3x3x3, 3: box growth from the own cell, 4: staged tile search + team pass, 5: start box / growth in two launches, 6: far queries in a warp pass, 7: cell-by-cell start box (knn_tile.cu)
  uint32_t* deferred;   // [Qt] sorted positions of deferred queries
  unsigned int* n_deferred;
  uint8_t* status;
  float4* nsd;          // weight*n, weight*d   (scan-to-map)
  float* weight;
  double* normal_cent;  // 6 per query (pair mode) or nullptr
  // debug (nullable)
  int32_t* idx5;
  float* sqd5;
  double* plane;
};

// destination of one segment's compacted matches
struct CompactDst {
  float4* cpw;
  float4* nsd;
  double* nc;
  int32_t* src;
};

// per-block work item of the evaluation kernels: a contiguous run of residuals of one keyframe
struct EvalItem {
  const float4* cpw;    // (cp.xyz, weight)
  const float4* nsd;    // (n_s.xyz, d_s)
  int32_t count;
  int32_t kf;
};

// Launch bookkeeping: counts every kernel launch (bench "gpu_launches") and, when profiling is enabled, brackets
// each launch with CUDA events on the launching stream so per-kernel device time can be read back
// (bench.py roofline: measured live inside the timed region).
struct LaunchCounter {
  int64_t n = 0;
  bool prof = false;
  struct Rec { const char* name; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get_event() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void begin(const char* name, cudaStream_t st) {
    ++n;
    if (prof) { Rec r{name, get_event(), get_event()}; cudaEventRecord(r.a, st); recs.push_back(r); }
  }
  void end(cudaStream_t st) { if (prof && !recs.empty()) cudaEventRecord(recs.back().b, st); }
};

// ---- K0 (grid.cu)
struct GridBuild {
  DevBuf<int> cell_start;
  DevBuf<float4> pts;
  DevBuf<float4> tmp4;     // unsorted (transformed) points, original order: (x,y,z,idx)
  DevBuf<int> sorted_pos;  // original index -> position in pts (cell-sorted)
  DevBuf<PairRec> pairs;   // pair layout of pts (built when build_pairs; knn_mode 4)
  bool build_pairs = false;
  DevBuf<int> fill;        // scatter cursors
  DevBuf<int> scan_tmp;
  DevBuf<int> bounds;      // 6 order-preserving ints on device
  GridDesc desc{};
  void release() { cell_start.release(); pts.release(); tmp4.release(); pairs.release(); sorted_pos.release(); fill.release(); scan_tmp.release(); bounds.release(); }
};
void grid_build(GridBuild& gb, const float* d_xyz, int stride, int64_t n, const double* t, const double* q,
                float cell_size_hint, float pts_per_cell, cudaStream_t st, LaunchCounter& lc);
void exclusive_scan_i32(const int* in, int* out, int64_t n, DevBuf<int>& tmp, cudaStream_t st, LaunchCounter& lc);

// ---- local map maintenance (grid.cu): world-frame keyframe clouds -> pcl::VoxelGrid-style down-sampling
struct VoxelWork {
  DevBuf<float4> tmp4;
  DevBuf<int> bounds, count, start, opos, order, sorted, scan_tmp, out_vox;
  DevBuf<float> out_xyz;      // filtered points, 3 floats each, ascending voxel index
  void release() { tmp4.release(); bounds.release(); count.release(); start.release(); opos.release(); order.release(); sorted.release(); scan_tmp.release(); out_vox.release(); out_xyz.release(); }
};
void localmap_transform(const float* d_xyz, int stride, int64_t n, const double* t, const double* q, float4* d_out, cudaStream_t st, LaunchCounter& lc);
// returns the number of filtered points (in w.out_xyz / w.out_vox), or -1 when PCL would pass the input through unchanged
int64_t voxel_filter_run(VoxelWork& w, const float4* d_in, int64_t n, float leaf, cudaStream_t st, LaunchCounter& lc);

// ---- K1 / K1b (assoc.cu)
void assoc_run(const GridBuild& gb, const SegDesc* d_segs, int nseg, const AssocWork& w, const AssocGates& gates,
               const float* oth_local, int oth_stride, DevBuf<int>& cell_count, DevBuf<int>& cell_pos,
               DevBuf<int>& scan_tmp, cudaStream_t st, LaunchCounter& lc);
void compact_count(const AssocWork& w, const SegDesc* d_segs, int nseg, int* d_flags, int* d_pos, DevBuf<int>& scan_tmp,
                   int* d_counts, cudaStream_t st, LaunchCounter& lc);
void compact_scatter(const AssocWork& w, const SegDesc* d_segs, int nseg, const int* d_pos, const void* d_dst, cudaStream_t st,
                     LaunchCounter& lc);
void compact_run(const AssocWork& w, const SegDesc* d_segs, int nseg, int* d_flags, int* d_pos, DevBuf<int>& scan_tmp,
                 const void* d_dst /*CompactDst[nseg]*/, int* d_counts, cudaStream_t st, LaunchCounter& lc);
void gather_selection(const int32_t* d_keep, int64_t n, int64_t n_match, const float4* cpw, const float4* nsd, const double* nc,
                      float4* o_cpw, float4* o_nsd, double* o_nc, int* d_bad, cudaStream_t st, LaunchCounter& lc);

// ---- K2 / K2b / K2e (eval.cu)
struct EvalParams {
  double q_lb[4], t_lb[3];
  double lidar_const;
  double huber_delta;
  int unit_score;       // 1: score = lidar_const for every match (front-end factor), 0: lidar_const * weight
};
constexpr int GLIO_NACC = 28;        // 21 upper-triangular H + 6 g + 1 cost
constexpr int EV_MAXW = 64;          // poses are passed in the kernel parameters up to this many keyframes
constexpr int GLIO_ITEM_MAX = 2048;  // residuals per evaluation work item
void eval_unary_run(const EvalItem* d_items, int nitems, int W, const double* d_poses, const EvalParams& ep, int jac_kind,
                    bool want_jac, double* d_partials, double* d_out, const int* d_kf_item_start, unsigned int* d_ticket,
                    cudaStream_t st, LaunchCounter& lc, unsigned int* done_flag = nullptr, unsigned int epoch = 0,
                    const double* h_poses = nullptr);   // h_poses: host copy of the poses, passed by value when W <= EV_MAXW
void eval_unary_residuals_run(const float4* cpw, const float4* nsd, int64_t n, const double* d_pose, const EvalParams& ep,
                              int jac_kind, double* d_r, double* d_J, cudaStream_t st, LaunchCounter& lc);


// ---- K2e (eval.cu): edge factors
struct EdgeItem {
  const float4* cps;   // (cp.xyz, s)
  const float4* pa;    // line point a
  const float4* pb;    // line point b
  int32_t count;
  int32_t kf;
};
void eval_edge_run(const EdgeItem* d_items, int nitems, int W, const double* d_poses, const EvalParams& ep, bool want_jac,
                   double* d_partials, double* d_out, const int* d_kf_item_start, unsigned int* d_ticket, cudaStream_t st, LaunchCounter& lc);

// ---- K2b (eval.cu): binary plane factors
struct BinItem {
  const float4* cpw;          // (cp.xyz, weight)
  const double* nc;           // 6 per residual: local-frame unit normal, local-frame centroid
  int32_t count;
  int32_t kf_c, kf_o;
};
constexpr int GLIO_NACC_BIN = 55;   // uu(6) uv(9) vv(6) uz(9) vz(9) zz(6) | r*u(3) r*v(3) r*z(3) | cost
// incidence list of a keyframe for the deterministic block assembly: (pair index, role 0 = cur / 1 = oth)
struct BinIncidence { int32_t pair; int32_t role; };
void eval_binary_run(const BinItem* d_items, int nitems, const int* d_pair_item_start, int n_pairs, int K, const double* d_poses,
                     double score_scale, double huber_delta, bool want_jac, double* d_partials, double* d_pair_sums,
                     const int* d_kf_inc_start, const BinIncidence* d_inc, double* d_diag /*K*28*/, double* d_off /*n_pairs*36*/,
                     double* d_cost, cudaStream_t st, LaunchCounter& lc);

// ---- front-end feature extraction (features.cu; Preprocessing.cpp:529-655)
constexpr int FEAT_MAX_SHARP = 12, FEAT_MAX_LESS_SHARP = 60, FEAT_MAX_FLAT = 24;   // per ring: 6 sectors x (2, 10, 4)
struct FeatArgs {
  const float* cloud; int stride; int ioff; int64_t n;     // ioff: float offset of the intensity (3 packed, 4 in pcl::PointXYZI)
  int n_scans; const int32_t* scan_start; const int32_t* scan_end; int ds_rate;
  double edge_thres, surf_thres; float ds_v;
  float* curv; int8_t* label; int8_t* picked;
  int32_t* ring_sharp; int32_t* ring_less_sharp; int32_t* ring_flat;   // [n_scans][MAX] slots
  int32_t* ring_less_flat;                                             // [n] ring r's list starts at scan_start[r]
  float4* ring_ds;                                                     // [n] ring r's down-sampled points start at scan_start[r]
  int32_t* counts;                                                     // [5][n_scans]: sharp, less_sharp, flat, less_flat, ds
  int32_t* offsets;                                                    // [5][n_scans + 1]
  int32_t* err;                                                        // != 0: a sector / ring exceeded the shared-memory capacity
  int32_t* out_sharp; int32_t* out_less_sharp; int32_t* out_flat; int32_t* out_less_flat; float4* out_ds;
};

void features_run(FeatArgs& a, cudaStream_t st, LaunchCounter& lc);

}  // namespace glio
