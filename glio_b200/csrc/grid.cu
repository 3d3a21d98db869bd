// grid.cu — K0: uniform-grid build over a point cloud (replaces pcl::KdTreeFLANN::setInputCloud,
// GLIO/src/Estimator.cpp:2056,3729-3731,3821-3823) + the shared exclusive-scan primitive.
//
// Layout: points are counting-sorted by cell id ((cz*ny + cy)*nx + cx, x fastest) into float4
// (x, y, z, bitcast(original index)).  HBM traffic: read 4*stride*n, write 16 n + the cell table.
#include "common.cuh"
#include "devmath.cuh"

namespace glio {

static std::string g_err;
void set_global_error(const std::string& m) { g_err = m; }
const char* global_error() { return g_err.c_str(); }

// ------------------------------------------------------------------------------------------------
// exclusive scan (int32), 3-phase: per-block scan of 4096 items + scan of block sums (recursive) + add.
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_T = 1024;
constexpr int SCAN_IPT = 4;
constexpr int SCAN_B = SCAN_T * SCAN_IPT;

__global__ void __launch_bounds__(SCAN_T) k_scan_block(const int* __restrict__ in, int* __restrict__ out,
                                                        int* __restrict__ block_sums, int64_t n) {
  __shared__ int warp_sums[32];
  const int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_IPT;
  int v[SCAN_IPT];
  int tsum = 0;
#pragma unroll
  for (int k = 0; k < SCAN_IPT; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    tsum += v[k];
  }
  // inclusive warp scan of thread sums
  int x = tsum;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  if (lane == 31) warp_sums[wid] = x;
  __syncthreads();
  if (wid == 0) {
    int w = warp_sums[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int y = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += y;
    }
    warp_sums[lane] = w;
  }
  __syncthreads();
  int excl = x - tsum + (wid > 0 ? warp_sums[wid - 1] : 0);
#pragma unroll
  for (int k = 0; k < SCAN_IPT; ++k) {
    if (base + k < n) out[base + k] = excl;
    excl += v[k];
  }
  if (threadIdx.x == SCAN_T - 1 && block_sums) block_sums[blockIdx.x] = excl;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_add(int* __restrict__ out, const int* __restrict__ block_offs, int64_t n) {
  const int add = block_offs[blockIdx.x];
  const int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_IPT;
#pragma unroll
  for (int k = 0; k < SCAN_IPT; ++k)
    if (base + k < n) out[base + k] += add;
}

// second half of a two-launch scan: every block sums the block totals in front of it by itself (nb <= 4096 totals: one
// coalesced read + a block reduction) instead of waiting for a third launch to scan them
__global__ void __launch_bounds__(SCAN_T) k_scan_add2(int* __restrict__ out, const int* __restrict__ block_sums, int64_t n) {
  __shared__ int warp_sums[32];
  int part = 0;
  for (int i = threadIdx.x; i < (int)blockIdx.x; i += SCAN_T) part += block_sums[i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = part;
  __syncthreads();
  int add = 0;
#pragma unroll
  for (int w = 0; w < SCAN_T / 32; ++w) add += warp_sums[w];
  if (blockIdx.x == 0) return;
  const int64_t base = (int64_t)blockIdx.x * SCAN_B + (int64_t)threadIdx.x * SCAN_IPT;
#pragma unroll
  for (int k = 0; k < SCAN_IPT; ++k)
    if (base + k < n) out[base + k] += add;
}

static void scan_rec(const int* in, int* out, int64_t n, int* tmp, int64_t tmp_cap, cudaStream_t st, LaunchCounter& lc) {
  const int64_t nb = (n + SCAN_B - 1) / SCAN_B;
  if (nb <= 1) {
    lc.begin("k_scan_block", st); k_scan_block<<<1, SCAN_T, 0, st>>>(in, out, nullptr, n); lc.end(st);
    return;
  }
  GLIO_REQUIRE(nb <= tmp_cap, GLIO_ERR_STATE, "scan scratch too small");
  lc.begin("k_scan_block", st); k_scan_block<<<(unsigned)nb, SCAN_T, 0, st>>>(in, out, tmp, n); lc.end(st);
  if (nb <= 4096) {
    lc.begin("k_scan_add", st); k_scan_add2<<<(unsigned)nb, SCAN_T, 0, st>>>(out, tmp, n); lc.end(st);
    return;
  }
  scan_rec(tmp, tmp, nb, tmp + nb, tmp_cap - nb, st, lc);
  lc.begin("k_scan_add", st); k_scan_add<<<(unsigned)nb, SCAN_T, 0, st>>>(out, tmp, n); lc.end(st);
}

void exclusive_scan_i32(const int* in, int* out, int64_t n, DevBuf<int>& tmp, cudaStream_t st, LaunchCounter& lc) {
  if (n <= 0) return;
  int64_t need = 0;
  for (int64_t m = (n + SCAN_B - 1) / SCAN_B; m > 1; m = (m + SCAN_B - 1) / SCAN_B) need += m;
  need += 8;
  tmp.reserve((size_t)need);
  scan_rec(in, out, n, tmp.p, (int64_t)tmp.cap, st, lc);
  GLIO_CUDA_TRY(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------
// grid build kernels
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__host__ __device__ __forceinline__ float ordered_to_float_h(int i) {
  int j = i ^ ((i >> 31) & 0x7fffffff);
#ifdef __CUDA_ARCH__
  return __int_as_float(j);
#else
  float f; memcpy(&f, &j, 4); return f;
#endif
}

// load (and optionally transform: world = float(q*double(p)+t), Estimator.cpp:1517-1545) points into float4 + index,
// and reduce the bounding box.
__global__ void __launch_bounds__(256) k_load_bounds(const float* __restrict__ xyz, int stride, int64_t n, PoseD pose, int has_pose,
                                                     float4* __restrict__ out, int* __restrict__ bounds /*6 ordered ints*/) {
  float mn[3] = {3.0e38f, 3.0e38f, 3.0e38f}, mx[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float p[3] = {xyz[i * stride], xyz[i * stride + 1], xyz[i * stride + 2]};
    if (has_pose) transform_point_f(pose, p, p);
    out[i] = make_float4(p[0], p[1], p[2], __int_as_float((int)i));
#pragma unroll
    for (int d = 0; d < 3; ++d) { mn[d] = fminf(mn[d], p[d]); mx[d] = fmaxf(mx[d], p[d]); }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
      mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
    }
  }
  // one set of global atomics per block (8 warps -> shared memory -> thread 0), not per warp
  __shared__ float smn[8][3], smx[8][3];
  const int wid = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) { smn[wid][d] = mn[d]; smx[wid][d] = mx[d]; }
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    const int d = threadIdx.x;
    float a = smn[0][d], b = smx[0][d];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { a = fminf(a, smn[w][d]); b = fmaxf(b, smx[w][d]); }
    atomicMin(&bounds[d], float_to_ordered(a));
    atomicMax(&bounds[3 + d], float_to_ordered(b));
  }
}

__global__ void k_init_bounds(int* bounds) {
  if (threadIdx.x < 3) bounds[threadIdx.x] = 0x7fffffff;
  else if (threadIdx.x < 6) bounds[threadIdx.x] = (int)0x80000000;
}

__global__ void __launch_bounds__(256) k_cell_hist(const float4* __restrict__ pts, int64_t n, GridDesc g, int* __restrict__ count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = pts[i];
    atomicAdd(&count[cell_of_clamped(g, p.x, p.y, p.z)], 1);
  }
}

__global__ void __launch_bounds__(256) k_cell_scatter(const float4* __restrict__ pts, int64_t n, GridDesc g,
                                                      const int* __restrict__ cell_start, int* __restrict__ fill,
                                                      float4* __restrict__ sorted, int* __restrict__ sorted_pos) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = pts[i];
    int c = cell_of_clamped(g, p.x, p.y, p.z);
    int pos = cell_start[c] + atomicAdd(&fill[c], 1);
    sorted[pos] = p;
    sorted_pos[i] = pos;          // original index -> position in the cell-sorted copy (K1b gathers neighbours from there)
  }
}

// pair layout of the sorted map (see PairRec); the odd tail half-record is a far-away filler
__global__ void __launch_bounds__(256) k_make_pairs(const float4* __restrict__ sorted, int64_t n, PairRec* __restrict__ pairs) {
  const int64_t np = (n + 1) >> 1;
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < np; k += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = sorted[2 * k];
    float4 b = make_float4(PAIR_FAR, 0.f, 0.f, __int_as_float(0x7fffffff));
    if (2 * k + 1 < n) b = sorted[2 * k + 1];
    PairRec r;
    r.x0 = a.x; r.x1 = b.x; r.y0 = a.y; r.y1 = b.y; r.z0 = a.z; r.z1 = b.z; r.i0 = __float_as_int(a.w); r.i1 = __float_as_int(b.w);
    pairs[k] = r;
  }
}

void grid_build(GridBuild& gb, const float* d_xyz, int stride, int64_t n, const double* t, const double* q,
                float cell_size_hint, float pts_per_cell, cudaStream_t st, LaunchCounter& lc) {
  GLIO_REQUIRE(n > 0 && n < (int64_t)1 << 31, GLIO_ERR_ARG, "grid_build: point count out of range");
  gb.tmp4.reserve((size_t)n);
  gb.pts.reserve((size_t)n);
  gb.sorted_pos.reserve((size_t)n);
  gb.bounds.reserve(8);
  PoseD pose{};
  int has_pose = 0;
  if (t && q) { has_pose = 1; for (int k = 0; k < 3; ++k) pose.t[k] = t[k]; for (int k = 0; k < 4; ++k) pose.q[k] = q[k]; }
  int* d_bounds = (int*)gb.bounds.p;
  lc.begin("k_init_bounds", st); k_init_bounds<<<1, 32, 0, st>>>(d_bounds); lc.end(st);
  const int nb = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  lc.begin("k_load_bounds", st); k_load_bounds<<<nb, 256, 0, st>>>(d_xyz, stride, n, pose, has_pose, gb.tmp4.p, d_bounds); lc.end(st);
  int hb[6];
  GLIO_CUDA_TRY(cudaMemcpyAsync(hb, d_bounds, sizeof(hb), cudaMemcpyDeviceToHost, st));
  GLIO_CUDA_TRY(cudaStreamSynchronize(st));
  float mn[3], mx[3];
  for (int d = 0; d < 3; ++d) { mn[d] = ordered_to_float_h(hb[d]); mx[d] = ordered_to_float_h(hb[3 + d]); }
  for (int d = 0; d < 3; ++d)
    GLIO_REQUIRE(std::isfinite(mn[d]) && std::isfinite(mx[d]), GLIO_ERR_ARG, "grid_build: non-finite point coordinates");
  // cell size: caller hint, else from the bounding-box surface area (LiDAR clouds are surfaces): about
  // pts_per_cell points per occupied cell.
  float L[3] = {mx[0] - mn[0] + 1e-3f, mx[1] - mn[1] + 1e-3f, mx[2] - mn[2] + 1e-3f};
  float cell = cell_size_hint;
  if (!(cell > 0)) {
    double area = 2.0 * ((double)L[0] * L[1] + (double)L[1] * L[2] + (double)L[0] * L[2]);
    cell = (float)std::sqrt(pts_per_cell * area / (double)n);
    if (cell < 0.05f) cell = 0.05f;
    if (cell > 4.0f) cell = 4.0f;
  }
  const double max_cells = 48.0e6;
  for (;;) {
    double nc = std::ceil(L[0] / cell + 1) * std::ceil(L[1] / cell + 1) * std::ceil(L[2] / cell + 1);
    if (nc <= max_cells) break;
    cell *= 1.26f;
  }
  GridDesc& g = gb.desc;
  g.cell = cell; g.inv_cell = 1.0f / cell;
  g.ox = mn[0] - 0.5f * cell; g.oy = mn[1] - 0.5f * cell; g.oz = mn[2] - 0.5f * cell;
  g.nx = (int)std::floor((mx[0] - g.ox) / cell) + 2;
  g.ny = (int)std::floor((mx[1] - g.oy) / cell) + 2;
  g.nz = (int)std::floor((mx[2] - g.oz) / cell) + 2;
  g.npts = n;
  const int64_t ncell = (int64_t)g.nx * g.ny * g.nz;
  gb.cell_start.reserve((size_t)ncell + 2);
  gb.fill.reserve((size_t)ncell + 2);
  GLIO_CUDA_TRY(cudaMemsetAsync(gb.fill.p, 0, (size_t)(ncell + 1) * sizeof(int), st));
  lc.begin("k_cell_hist", st); k_cell_hist<<<nb, 256, 0, st>>>(gb.tmp4.p, n, g, gb.fill.p); lc.end(st);
  exclusive_scan_i32(gb.fill.p, gb.cell_start.p, ncell + 1, gb.scan_tmp, st, lc);
  GLIO_CUDA_TRY(cudaMemsetAsync(gb.fill.p, 0, (size_t)(ncell + 1) * sizeof(int), st));
  lc.begin("k_cell_scatter", st); k_cell_scatter<<<nb, 256, 0, st>>>(gb.tmp4.p, n, g, gb.cell_start.p, gb.fill.p, gb.pts.p, gb.sorted_pos.p); lc.end(st);
  if (gb.build_pairs) {
    gb.pairs.reserve((size_t)((n + 1) / 2 + 1));
    lc.begin("k_make_pairs", st); k_make_pairs<<<nb, 256, 0, st>>>(gb.pts.p, n, gb.pairs.p); lc.end(st);
  }
  GLIO_CUDA_TRY(cudaGetLastError());
  g.cell_start = gb.cell_start.p;
  g.pts = gb.pts.p;
}

// =====================================================================================================================
// Local map maintenance on the device: keyframe clouds moved to the world frame when they are pushed
// (transformCloud, Estimator.cpp:1517-1545), concatenated, then pcl::VoxelGrid down-sampling as GLIO configures it
// (ds_filter_surf_map, leaf 0.4 m, Estimator.cpp:854, 3617-3618).  The voxel filter is a counting sort by PCL's voxel
// index (float min/max, min_b = floor(min*inv_leaf), idx = ijk0 + ijk1*div0 + ijk2*div0*div1) followed by a per-voxel float
// sum in INPUT order and a float division by the count; output in ascending voxel index (PCL's output order).
// PCL sums the points of a voxel in whatever order std::sort leaves them; summing in input order is the deterministic
// member of that family (the stable order).
// =====================================================================================================================
__global__ void __launch_bounds__(256) k_lm_transform(const float* __restrict__ xyz, int stride, int64_t n, PoseD pose, float4* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float p[3] = {xyz[i * stride], xyz[i * stride + 1], xyz[i * stride + 2]};
    transform_point_f(pose, p, p);
    out[i] = make_float4(p[0], p[1], p[2], 0.f);
  }
}

void localmap_transform(const float* d_xyz, int stride, int64_t n, const double* t, const double* q, float4* d_out, cudaStream_t st, LaunchCounter& lc) {
  PoseD pose;
  for (int k = 0; k < 3; ++k) pose.t[k] = t[k];
  for (int k = 0; k < 4; ++k) pose.q[k] = q[k];
  const int nb = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  lc.begin("k_lm_transform", st); k_lm_transform<<<nb, 256, 0, st>>>(d_xyz, stride, n, pose, d_out); lc.end(st);
  GLIO_CUDA_TRY(cudaGetLastError());
}

struct VoxelDesc { float inv; int min_b[3]; int div0, div01; };
__device__ __forceinline__ int voxel_of(const VoxelDesc& v, float x, float y, float z) {
  const int i0 = (int)__fsub_rn(floorf(__fmul_rn(x, v.inv)), (float)v.min_b[0]);
  const int i1 = (int)__fsub_rn(floorf(__fmul_rn(y, v.inv)), (float)v.min_b[1]);
  const int i2 = (int)__fsub_rn(floorf(__fmul_rn(z, v.inv)), (float)v.min_b[2]);
  return i0 + i1 * v.div0 + i2 * v.div01;
}
__global__ void __launch_bounds__(256) k_vox_hist(const float4* __restrict__ pts, int64_t n, VoxelDesc v, int* __restrict__ count) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    atomicAdd(&count[voxel_of(v, p.x, p.y, p.z)], 1);
  }
}
__global__ void __launch_bounds__(256) k_vox_scatter(const float4* __restrict__ pts, int64_t n, VoxelDesc v, const int* __restrict__ start,
                                                     int* __restrict__ fill, int* __restrict__ order) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 p = pts[i];
    const int c = voxel_of(v, p.x, p.y, p.z);
    order[start[c] + atomicAdd(&fill[c], 1)] = (int)i;
  }
}
// Order the point indices of every voxel ascending (= input order): the atomic scatter leaves them in arrival order.
// One warp per voxel: bitonic sort of 32*K keys held K per lane (shuffles for partner distances < 32, register swaps
// above); voxels with more than 256 points fall back to counting ranks.
template <int K>
__device__ __forceinline__ void warp_bitonic_sort(int (&v)[K], int lane) {
  constexpr int N = 32 * K;
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32) {
        const int jr = j >> 5;
#pragma unroll
        for (int r = 0; r < K; ++r) {
          const int pr = r ^ jr;
          if (pr > r) {
            const bool asc = (((r << 5) + lane) & k) == 0;
            const int a = v[r], b = v[pr];
            const bool sw = asc ? (a > b) : (a < b);
            v[r] = sw ? b : a; v[pr] = sw ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int r = 0; r < K; ++r) {
          const bool asc = (((r << 5) + lane) & k) == 0;
          const int o = __shfl_xor_sync(0xffffffffu, v[r], j);
          const bool lower = (lane & j) == 0;
          v[r] = (lower == asc) ? min(v[r], o) : max(v[r], o);
        }
      }
    }
  }
}

template <int K>
__device__ __forceinline__ void sort_segment(const int* __restrict__ order, int* __restrict__ sorted, int s, int n, int lane) {
  int v[K];
#pragma unroll
  for (int r = 0; r < K; ++r) { const int i = (r << 5) + lane; v[r] = i < n ? order[s + i] : 0x7fffffff; }
  warp_bitonic_sort<K>(v, lane);
#pragma unroll
  for (int r = 0; r < K; ++r) { const int i = (r << 5) + lane; if (i < n) sorted[s + i] = v[r]; }
}

__global__ void __launch_bounds__(256) k_vox_sort(const int* __restrict__ start, int64_t nvox, const int* __restrict__ order, int* __restrict__ sorted) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t c0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 32; c0 < nvox; c0 += nwarps * 32) {
    // the 32 lanes look at 32 consecutive voxels, then the warp serves the non-empty ones in turn
    const int64_t c = c0 + lane;
    int s = 0, n = 0;
    if (c < nvox) { s = start[c]; n = start[c + 1] - s; }
    unsigned m = __ballot_sync(0xffffffffu, n > 0);
    while (m) {
      const int l = __ffs(m) - 1; m &= m - 1;
      const int ss = __shfl_sync(0xffffffffu, s, l), nn = __shfl_sync(0xffffffffu, n, l);
      if (nn == 1) { if (lane == 0) sorted[ss] = order[ss]; }
      else if (nn <= 32) sort_segment<1>(order, sorted, ss, nn, lane);
      else if (nn <= 64) sort_segment<2>(order, sorted, ss, nn, lane);
      else if (nn <= 128) sort_segment<4>(order, sorted, ss, nn, lane);
      else if (nn <= 256) sort_segment<8>(order, sorted, ss, nn, lane);
      else {
        for (int i = lane; i < nn; i += 32) {
          const int key = order[ss + i];
          int r = 0;
          for (int k = 0; k < nn; ++k) r += order[ss + k] < key ? 1 : 0;
          sorted[ss + r] = key;
        }
      }
    }
  }
}
__global__ void __launch_bounds__(256) k_vox_flags(const int* __restrict__ start, int64_t nvox, int* __restrict__ flags) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < nvox) flags[c] = start[c + 1] > start[c] ? 1 : 0;
  if (c == nvox) flags[c] = 0;
}
__global__ void __launch_bounds__(256) k_vox_centroid(const float4* __restrict__ pts, const int* __restrict__ start, const int* __restrict__ sorted,
                                                      const int* __restrict__ opos, int64_t nvox, float* __restrict__ out_xyz, int* __restrict__ out_vox) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nvox) return;
  const int s = start[c], e = start[c + 1];
  if (e <= s) return;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int k = s; k < e; ++k) { const float4 p = pts[sorted[k]]; sx = __fadd_rn(sx, p.x); sy = __fadd_rn(sy, p.y); sz = __fadd_rn(sz, p.z); }
  const float cnt = (float)(e - s);
  const int o = opos[c];
  out_xyz[3 * (size_t)o] = __fdiv_rn(sx, cnt); out_xyz[3 * (size_t)o + 1] = __fdiv_rn(sy, cnt); out_xyz[3 * (size_t)o + 2] = __fdiv_rn(sz, cnt);
  if (out_vox) out_vox[o] = (int)c;
}

int64_t voxel_filter_run(VoxelWork& w, const float4* d_in, int64_t n, float leaf, cudaStream_t st, LaunchCounter& lc) {
  GLIO_REQUIRE(n > 0 && n < ((int64_t)1 << 31), GLIO_ERR_ARG, "voxel_filter: bad point count");
  GLIO_REQUIRE(leaf > 0.f, GLIO_ERR_ARG, "voxel_filter: leaf size must be positive");
  w.tmp4.reserve((size_t)n); w.bounds.reserve(8);
  lc.begin("k_init_bounds", st); k_init_bounds<<<1, 32, 0, st>>>(w.bounds.p); lc.end(st);
  const int nb = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  PoseD none{};
  lc.begin("k_load_bounds", st); k_load_bounds<<<nb, 256, 0, st>>>(reinterpret_cast<const float*>(d_in), 4, n, none, 0, w.tmp4.p, w.bounds.p); lc.end(st);
  int hb[6];
  GLIO_CUDA_TRY(cudaMemcpyAsync(hb, w.bounds.p, sizeof(hb), cudaMemcpyDeviceToHost, st));
  GLIO_CUDA_TRY(cudaStreamSynchronize(st));
  float mn[3], mx[3];
  for (int d = 0; d < 3; ++d) { mn[d] = ordered_to_float_h(hb[d]); mx[d] = ordered_to_float_h(hb[3 + d]); }
  for (int d = 0; d < 3; ++d) GLIO_REQUIRE(std::isfinite(mn[d]) && std::isfinite(mx[d]), GLIO_ERR_ARG, "voxel_filter: non-finite point coordinates");
  const float inv = 1.0f / leaf;                               // setLeafSize: Array4f::Ones() / leaf
  const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1, dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)0x7fffffff) return -1;           // PCL: "leaf size too small", input passed through
  VoxelDesc v; v.inv = inv;
  int div[3];
  for (int d = 0; d < 3; ++d) { v.min_b[d] = (int)std::floor(mn[d] * inv); div[d] = (int)std::floor(mx[d] * inv) - v.min_b[d] + 1; }
  v.div0 = div[0]; v.div01 = div[0] * div[1];
  const int64_t nvox = (int64_t)div[0] * div[1] * div[2];
  GLIO_REQUIRE(nvox <= (int64_t)256 * 1024 * 1024, GLIO_ERR_ARG, "voxel_filter: more than 2^28 voxels in the bounding box (dense voxel table)");
  w.count.reserve((size_t)nvox + 2); w.start.reserve((size_t)nvox + 2); w.opos.reserve((size_t)nvox + 2);
  w.order.reserve((size_t)n); w.sorted.reserve((size_t)n); w.out_xyz.reserve((size_t)n * 3); w.out_vox.reserve((size_t)n);
  GLIO_CUDA_TRY(cudaMemsetAsync(w.count.p, 0, (size_t)(nvox + 1) * sizeof(int), st));
  lc.begin("k_vox_hist", st); k_vox_hist<<<nb, 256, 0, st>>>(w.tmp4.p, n, v, w.count.p); lc.end(st);
  exclusive_scan_i32(w.count.p, w.start.p, nvox + 1, w.scan_tmp, st, lc);
  GLIO_CUDA_TRY(cudaMemsetAsync(w.count.p, 0, (size_t)(nvox + 1) * sizeof(int), st));
  lc.begin("k_vox_scatter", st); k_vox_scatter<<<nb, 256, 0, st>>>(w.tmp4.p, n, v, w.start.p, w.count.p, w.order.p); lc.end(st);
  lc.begin("k_vox_sort", st); k_vox_sort<<<148 * 8, 256, 0, st>>>(w.start.p, nvox, w.order.p, w.sorted.p); lc.end(st);
  const unsigned nbv = (unsigned)((nvox + 1 + 255) / 256);
  lc.begin("k_vox_flags", st); k_vox_flags<<<nbv, 256, 0, st>>>(w.start.p, nvox, w.count.p); lc.end(st);
  exclusive_scan_i32(w.count.p, w.opos.p, nvox + 1, w.scan_tmp, st, lc);
  lc.begin("k_vox_centroid", st); k_vox_centroid<<<nbv, 256, 0, st>>>(w.tmp4.p, w.start.p, w.sorted.p, w.opos.p, nvox, w.out_xyz.p, w.out_vox.p); lc.end(st);
  int m = 0;
  GLIO_CUDA_TRY(cudaMemcpyAsync(&m, w.opos.p + nvox, sizeof(int), cudaMemcpyDeviceToHost, st));
  GLIO_CUDA_TRY(cudaStreamSynchronize(st));
  GLIO_CUDA_TRY(cudaGetLastError());
  return (int64_t)m;
}


}  // namespace glio
