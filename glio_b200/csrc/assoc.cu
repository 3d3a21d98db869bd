// assoc.cu — K1 / K1b: per-point LiDAR surf association on a uniform grid.
//
// Replaces Estimator::findCorrespondingSurfFeatures (GLIO/src/Estimator.cpp:3633-3708) and
// findGlobalCorrespondingSurfFeatures[Add]_Batch (Estimator.cpp:3710-3892):
//   transform -> exact 5-NN (FLANN L2_Simple<float> distances, ties by index) -> gate on the squared 5th
//   distance -> 5x3 column-pivoted Householder LS plane fit (double) -> validity -> weight -> outputs.
// Exactness of the kNN: rings of grid cells are searched until the 5th best distance is provably inside the
// searched cube, or the cube already covers the gate radius (then anything unseen fails the gate anyway).
//
// This translation unit is compiled with -fmad=false (see Makefile) in addition to the explicit *_rn
// intrinsics of devmath.cuh: kNN indices and the valid mask must be bit-exact w.r.t. the no-FMA reference.
#include "common.cuh"
#include "devmath.cuh"

namespace glio {

__device__ __forceinline__ int find_seg(const SegDesc* __restrict__ segs, int nseg, int64_t g) {
  int lo = 0, hi = nseg - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (segs[mid].offset <= g) lo = mid; else hi = mid - 1;
  }
  return lo;
}

// ---- pass 1: transform queries, histogram them by (clamped) grid cell --------------------------------
__global__ void __launch_bounds__(256) k_transform_hist(GridDesc grid, const SegDesc* __restrict__ segs, int nseg,
                                                        int64_t Qt, float4* __restrict__ pm, uint16_t* __restrict__ segid,
                                                        int* __restrict__ cell_count) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Qt) return;
  const int s = find_seg(segs, nseg, g);
  const SegDesc& sd = segs[s];
  const float* src = sd.src + (g - sd.offset) * sd.stride;
  float p[3] = {src[0], src[1], src[2]};
  PoseD P;
  P.t[0] = sd.t[0]; P.t[1] = sd.t[1]; P.t[2] = sd.t[2];
  P.q[0] = sd.q[0]; P.q[1] = sd.q[1]; P.q[2] = sd.q[2]; P.q[3] = sd.q[3];
  transform_point_f(P, p, p);
  pm[g] = make_float4(p[0], p[1], p[2], 0.f);
  segid[g] = (uint16_t)s;
  atomicAdd(&cell_count[cell_of_clamped(grid, p[0], p[1], p[2])], 1);
}

// ---- pass 2: scatter query ids in cell order (locality: a warp's queries share grid cells) -----------
__global__ void __launch_bounds__(256) k_order_scatter(GridDesc grid, int64_t Qt, const float4* __restrict__ pm,
                                                       const int* __restrict__ cell_start, int* __restrict__ fill,
                                                       uint32_t* __restrict__ order) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Qt) return;
  const float4 p = pm[g];
  const int c = cell_of_clamped(grid, p.x, p.y, p.z);
  const int pos = cell_start[c] + atomicAdd(&fill[c], 1);
  order[pos] = (uint32_t)g;
}

// ---- top-5 by (distance, index) -------------------------------------------------------------------
struct Top5 {
  float d0, d1, d2, d3, d4;
  int i0, i1, i2, i3, i4;
};
__device__ __forceinline__ bool lex_less(float da, int ia, float db, int ib) { return da < db || (da == db && ia < ib); }
#define GLIO_CSWAP(DA, IA, DB, IB) \
  if (lex_less(DB, IB, DA, IA)) { float _d = DA; DA = DB; DB = _d; int _i = IA; IA = IB; IB = _i; }
__device__ __forceinline__ void top5_push(Top5& t, float d, int id) {
  if (lex_less(d, id, t.d4, t.i4)) {
    t.d4 = d; t.i4 = id;
    GLIO_CSWAP(t.d3, t.i3, t.d4, t.i4)
    GLIO_CSWAP(t.d2, t.i2, t.d3, t.i3)
    GLIO_CSWAP(t.d1, t.i1, t.d2, t.i2)
    GLIO_CSWAP(t.d0, t.i0, t.d1, t.i1)
  }
}

__device__ __forceinline__ void scan_range(const float4* __restrict__ pts, int s, int e, float qx, float qy, float qz, Top5& t) {
  for (int k = s; k < e; ++k) {
    const float4 p = __ldg(&pts[k]);
    top5_push(t, l2_simple(qx, qy, qz, p.x, p.y, p.z), __float_as_int(p.w));
  }
}

// Exact 5-NN of (qx,qy,qz) in the grid.  Guarantee: if the true 5th squared distance is < gate (+margin) the
// returned five are exact; otherwise t.d4 >= gate (the caller's radius gate fails either way).
__device__ __forceinline__ void knn5_grid(const GridDesc& g, float qx, float qy, float qz, float gate_sq, Top5& t) {
  const float INF = __int_as_float(0x7f800000);
  t.d0 = t.d1 = t.d2 = t.d3 = t.d4 = INF;
  t.i0 = t.i1 = t.i2 = t.i3 = t.i4 = 0x7fffffff;
  const int cx = cell_coord(qx, g.ox, g.inv_cell), cy = cell_coord(qy, g.oy, g.inv_cell), cz = cell_coord(qz, g.oz, g.inv_cell);
  // number of rings that covers sqrt(gate) with a safety margin (mis-binning by float rounding < 1e-3 m)
  const int rmax = (int)ceilf((sqrtf(gate_sq) + 2e-3f) * g.inv_cell) + 1;
  // far outside the grid: nothing within the gate
  if (cx < -rmax || cy < -rmax || cz < -rmax || cx >= g.nx + rmax || cy >= g.ny + rmax || cz >= g.nz + rmax) return;
  const int* __restrict__ cs = g.cell_start;
  for (int r = 1; r <= rmax; ++r) {
    const int z0 = max(cz - r, 0), z1 = min(cz + r, g.nz - 1);
    const int y0 = max(cy - r, 0), y1 = min(cy + r, g.ny - 1);
    const int xa = cx - r, xb = cx + r;
    const int x0 = max(xa, 0), x1 = min(xb, g.nx - 1);
    for (int z = z0; z <= z1; ++z) {
      const bool zshell = (z == cz - r) || (z == cz + r);
      for (int y = y0; y <= y1; ++y) {
        const int row = (z * g.ny + y) * g.nx;
        const bool shell = r == 1 || zshell || (y == cy - r) || (y == cy + r);
        if (shell) {
          if (x0 <= x1) scan_range(g.pts, __ldg(&cs[row + x0]), __ldg(&cs[row + x1 + 1]), qx, qy, qz, t);
        } else {
          if (xa >= 0 && xa < g.nx) scan_range(g.pts, __ldg(&cs[row + xa]), __ldg(&cs[row + xa + 1]), qx, qy, qz, t);
          if (xb >= 0 && xb < g.nx) scan_range(g.pts, __ldg(&cs[row + xb]), __ldg(&cs[row + xb + 1]), qx, qy, qz, t);
        }
      }
    }
    // distance from the query to the faces of the searched cube (faces clipped by the grid are infinitely far:
    // nothing lives outside the grid)
    float b = INF;
    if (cx - r > 0)        b = fminf(b, qx - (g.ox + (float)(cx - r) * g.cell));
    if (cx + r < g.nx - 1) b = fminf(b, (g.ox + (float)(cx + r + 1) * g.cell) - qx);
    if (cy - r > 0)        b = fminf(b, qy - (g.oy + (float)(cy - r) * g.cell));
    if (cy + r < g.ny - 1) b = fminf(b, (g.oy + (float)(cy + r + 1) * g.cell) - qy);
    if (cz - r > 0)        b = fminf(b, qz - (g.oz + (float)(cz - r) * g.cell));
    if (cz + r < g.nz - 1) b = fminf(b, (g.oz + (float)(cz + r + 1) * g.cell) - qz);
    if (b == INF) return;                       // the cube covers the whole grid
    const float bs = b * 0.999f - 2e-3f;        // safety: float rounding of cell assignment / face positions
    if (bs > 0.f && t.d4 <= bs * bs) return;    // 5th neighbour provably inside the cube
  }
}

struct KnnArgs {
  GridDesc grid;
  const SegDesc* segs;
  int64_t Qt;
  const float4* pm;
  const uint16_t* segid;
  const uint32_t* order;
  AssocGates gates;
  // outputs
  uint8_t* status;
  float4* nsd;
  float* weight;
  double* normal_cent;
  int32_t* idx5;
  float* sqd5;
  double* plane;
  // pair mode: local-frame points of the searched frame (original order)
  const float* oth_local;
  int oth_stride;
  // original-order world points of the searched cloud are recovered from grid.pts via idx? no: see pts_by_idx
  const float4* pts_by_idx;   // unsorted float4 copy (x,y,z,idx) of the searched cloud, original order
};

template <bool PAIR>
__global__ void __launch_bounds__(128) k_knn_plane(KnnArgs a) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.Qt) return;
  const int64_t g = a.order[p];
  const float4 q4 = a.pm[g];
  Top5 t;
  knn5_grid(a.grid, q4.x, q4.y, q4.z, (float)a.gates.max_radius, t);
  int id[5] = {t.i0, t.i1, t.i2, t.i3, t.i4};
  uint8_t st;
  float w = 0.f;
  double n[3] = {0, 0, 0}, d = 0;
  double nl[3] = {0, 0, 0}, cl[3] = {0, 0, 0};
  if (t.i4 != 0x7fffffff && (double)t.d4 < a.gates.max_radius) {          // Estimator.cpp:3651 / :3751
    double A[3][5];
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float4 m = __ldg(&a.pts_by_idx[id[j]]);
      A[0][j] = (double)m.x; A[1][j] = (double)m.y; A[2][j] = (double)m.z;
    }
    double Aw[3][5];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int j = 0; j < 5; ++j) Aw[c][j] = A[c][j];
    double x[3];
    plane_solve5(Aw, x);                                                    // :3661
    plane_from_solution(x, n, d);                                           // :3662-3663
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 5; ++j) {                                           // :3667-3674
      const double v = dadd(dadd(dadd(dmul(n[0], A[0][j]), dmul(n[1], A[1][j])), dmul(n[2], A[2][j])), d);
      if (fabs(v) > a.gates.dist_thres) ok = false;
    }
    if (ok) {
      w = weight_of(n, d, q4.x, q4.y, q4.z);                                // :3678-3679
      st = ((double)w > a.gates.weight_min) ? GLIO_MATCH_VALID : GLIO_MATCH_FAIL_WEIGHT;  // :3681
    } else st = GLIO_MATCH_FAIL_PLANE;
    if (PAIR && st == GLIO_MATCH_VALID) {
      // local-frame fit on the same five indices (Estimator.cpp:3752-3772)
      double Al[3][5];
      double sx = 0.0, sy = 0.0, sz = 0.0;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const float* lp = a.oth_local + (int64_t)id[j] * a.oth_stride;
        Al[0][j] = (double)lp[0]; Al[1][j] = (double)lp[1]; Al[2][j] = (double)lp[2];
        sx = dadd(sx, Al[0][j]); sy = dadd(sy, Al[1][j]); sz = dadd(sz, Al[2][j]);
      }
      cl[0] = sx / 5.0; cl[1] = sy / 5.0; cl[2] = sz / 5.0;
      double xl[3], dl;
      plane_solve5(Al, xl);
      plane_from_solution(xl, nl, dl);
    }
  } else st = GLIO_MATCH_FAIL_RADIUS;

  a.status[g] = st;
  if (!PAIR) {
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (st == GLIO_MATCH_VALID) {                                            // :3683-3687 float stores
      o.x = (float)dmul((double)w, n[0]); o.y = (float)dmul((double)w, n[1]);
      o.z = (float)dmul((double)w, n[2]); o.w = (float)dmul((double)w, d);
    }
    a.nsd[g] = o;
  } else {
    double* nc = a.normal_cent + 6 * g;
    const bool v = st == GLIO_MATCH_VALID;
    nc[0] = v ? nl[0] : 0.0; nc[1] = v ? nl[1] : 0.0; nc[2] = v ? nl[2] : 0.0;
    nc[3] = v ? cl[0] : 0.0; nc[4] = v ? cl[1] : 0.0; nc[5] = v ? cl[2] : 0.0;
  }
  a.weight[g] = (st == GLIO_MATCH_VALID || st == GLIO_MATCH_FAIL_WEIGHT) ? w : 0.f;
  if (a.idx5) {
    const float sd[5] = {t.d0, t.d1, t.d2, t.d3, t.d4};
#pragma unroll
    for (int j = 0; j < 5; ++j) { a.idx5[5 * g + j] = (id[j] == 0x7fffffff) ? -1 : id[j]; a.sqd5[5 * g + j] = sd[j]; }
    a.plane[4 * g] = n[0]; a.plane[4 * g + 1] = n[1]; a.plane[4 * g + 2] = n[2]; a.plane[4 * g + 3] = d;
  }
}

void assoc_run(const GridBuild& gb, const SegDesc* d_segs, int nseg, const AssocWork& w, const AssocGates& gates,
               const float* oth_local, int oth_stride, DevBuf<int>& cell_count, DevBuf<int>& cell_pos,
               DevBuf<int>& scan_tmp, cudaStream_t st, LaunchCounter& lc) {
  const GridDesc& grid = gb.desc;
  const int64_t Qt = w.Qt;
  if (Qt <= 0) return;
  GLIO_REQUIRE(Qt < ((int64_t)1 << 32), GLIO_ERR_ARG, "assoc_run: too many queries in one launch");
  GLIO_REQUIRE(nseg > 0 && nseg < 65536, GLIO_ERR_ARG, "assoc_run: bad segment count");
  const int64_t ncell = (int64_t)grid.nx * grid.ny * grid.nz;
  cell_count.reserve((size_t)ncell + 2);
  cell_pos.reserve((size_t)ncell + 2);
  GLIO_CUDA_TRY(cudaMemsetAsync(cell_count.p, 0, (size_t)(ncell + 1) * sizeof(int), st));
  const unsigned nb = (unsigned)((Qt + 255) / 256);
  k_transform_hist<<<nb, 256, 0, st>>>(grid, d_segs, nseg, Qt, w.pm, w.seg, cell_count.p); lc.n++;
  exclusive_scan_i32(cell_count.p, cell_pos.p, ncell + 1, scan_tmp, st, lc);
  GLIO_CUDA_TRY(cudaMemsetAsync(cell_count.p, 0, (size_t)(ncell + 1) * sizeof(int), st));
  k_order_scatter<<<nb, 256, 0, st>>>(grid, Qt, w.pm, cell_pos.p, cell_count.p, w.order); lc.n++;
  KnnArgs a;
  a.grid = grid; a.segs = d_segs; a.Qt = Qt; a.pm = w.pm; a.segid = w.seg; a.order = w.order; a.gates = gates;
  a.status = w.status; a.nsd = w.nsd; a.weight = w.weight; a.normal_cent = w.normal_cent;
  a.idx5 = w.idx5; a.sqd5 = w.sqd5; a.plane = w.plane;
  a.oth_local = oth_local; a.oth_stride = oth_stride; a.pts_by_idx = gb.tmp4.p;
  const unsigned nk = (unsigned)((Qt + 127) / 128);
  if (w.normal_cent) { k_knn_plane<true><<<nk, 128, 0, st>>>(a); }
  else { k_knn_plane<false><<<nk, 128, 0, st>>>(a); }
  lc.n++;
  GLIO_CUDA_TRY(cudaGetLastError());
}

// ---- stream compaction of the valid matches, per segment, in scan order ------------------------------
__global__ void __launch_bounds__(256) k_flags(const uint8_t* __restrict__ status, int64_t Qt, int* __restrict__ flags) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < Qt) flags[g] = status[g] == GLIO_MATCH_VALID ? 1 : 0;
  if (g == Qt) flags[g] = 0;
}

__global__ void __launch_bounds__(256) k_compact(const SegDesc* __restrict__ segs, int64_t Qt, const uint16_t* __restrict__ segid,
                                                 const uint8_t* __restrict__ status, const int* __restrict__ pos,
                                                 const float4* __restrict__ nsd, const float* __restrict__ weight,
                                                 const double* __restrict__ normal_cent, const CompactDst* __restrict__ dst) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= Qt) return;
  if (status[g] != GLIO_MATCH_VALID) return;
  const int s = segid[g];
  const SegDesc& sd = segs[s];
  const int lp = pos[g] - pos[sd.offset];
  const int64_t li = g - sd.offset;
  const float* src = sd.src + li * sd.stride;
  const CompactDst d = dst[s];
  d.cpw[lp] = make_float4(src[0], src[1], src[2], weight[g]);
  if (d.nsd) d.nsd[lp] = nsd[g];
  if (d.nc) {
#pragma unroll
    for (int k = 0; k < 6; ++k) d.nc[6 * (int64_t)lp + k] = normal_cent[6 * g + k];
  }
  d.src[lp] = (int32_t)li;
}

__global__ void k_seg_counts(const SegDesc* __restrict__ segs, int nseg, const int* __restrict__ pos, int* __restrict__ counts) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nseg) counts[s] = pos[segs[s].offset + segs[s].count] - pos[segs[s].offset];
}

void compact_run(const AssocWork& w, const SegDesc* d_segs, int nseg, int* d_flags, int* d_pos, DevBuf<int>& scan_tmp,
                 const void* d_dst /*CompactDst[nseg]*/, int* d_counts, cudaStream_t st, LaunchCounter& lc) {
  const int64_t Qt = w.Qt;
  const unsigned nb = (unsigned)((Qt + 1 + 255) / 256);
  k_flags<<<nb, 256, 0, st>>>(w.status, Qt, d_flags); lc.n++;
  exclusive_scan_i32(d_flags, d_pos, Qt + 1, scan_tmp, st, lc);
  k_seg_counts<<<(nseg + 127) / 128, 128, 0, st>>>(d_segs, nseg, d_pos, d_counts); lc.n++;
  k_compact<<<nb, 256, 0, st>>>(d_segs, Qt, w.seg, w.status, d_pos, w.nsd, w.weight, w.normal_cent, (const CompactDst*)d_dst); lc.n++;
  GLIO_CUDA_TRY(cudaGetLastError());
}

// gather (feature selection as an input index list)
__global__ void __launch_bounds__(256) k_gather_sel(const int32_t* __restrict__ keep, int64_t n, int64_t n_match,
                                                    const float4* __restrict__ cpw, const float4* __restrict__ nsd,
                                                    const double* __restrict__ nc, float4* __restrict__ o_cpw,
                                                    float4* __restrict__ o_nsd, double* __restrict__ o_nc, int* __restrict__ bad) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int k = keep[i];
  if (k < 0 || k >= n_match) { atomicAdd(bad, 1); return; }
  o_cpw[i] = cpw[k];
  if (nsd) o_nsd[i] = nsd[k];
  if (nc) {
#pragma unroll
    for (int c = 0; c < 6; ++c) o_nc[6 * i + c] = nc[6 * (int64_t)k + c];
  }
}

void gather_selection(const int32_t* d_keep, int64_t n, int64_t n_match, const float4* cpw, const float4* nsd, const double* nc,
                      float4* o_cpw, float4* o_nsd, double* o_nc, int* d_bad, cudaStream_t st, LaunchCounter& lc) {
  if (n <= 0) return;
  k_gather_sel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_keep, n, n_match, cpw, nsd, nc, o_cpw, o_nsd, o_nc, d_bad); lc.n++;
  GLIO_CUDA_TRY(cudaGetLastError());
}

}  // namespace glio
