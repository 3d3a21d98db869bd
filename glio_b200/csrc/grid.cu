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

static void scan_rec(const int* in, int* out, int64_t n, int* tmp, int64_t tmp_cap, cudaStream_t st, LaunchCounter& lc) {
  const int64_t nb = (n + SCAN_B - 1) / SCAN_B;
  if (nb <= 1) {
    lc.begin("k_scan_block", st); k_scan_block<<<1, SCAN_T, 0, st>>>(in, out, nullptr, n); lc.end(st);
    return;
  }
  GLIO_REQUIRE(nb <= tmp_cap, GLIO_ERR_STATE, "scan scratch too small");
  lc.begin("k_scan_block", st); k_scan_block<<<(unsigned)nb, SCAN_T, 0, st>>>(in, out, tmp, n); lc.end(st);
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
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      atomicMin(&bounds[d], float_to_ordered(mn[d]));
      atomicMax(&bounds[3 + d], float_to_ordered(mx[d]));
    }
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
                                                      float4* __restrict__ sorted) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float4 p = pts[i];
    int c = cell_of_clamped(g, p.x, p.y, p.z);
    int pos = cell_start[c] + atomicAdd(&fill[c], 1);
    sorted[pos] = p;
  }
}

void grid_build(GridBuild& gb, const float* d_xyz, int stride, int64_t n, const double* t, const double* q,
                float cell_size_hint, float pts_per_cell, cudaStream_t st, LaunchCounter& lc) {
  GLIO_REQUIRE(n > 0 && n < (int64_t)1 << 31, GLIO_ERR_ARG, "grid_build: point count out of range");
  gb.tmp4.reserve((size_t)n);
  gb.pts.reserve((size_t)n);
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
  lc.begin("k_cell_scatter", st); k_cell_scatter<<<nb, 256, 0, st>>>(gb.tmp4.p, n, g, gb.cell_start.p, gb.fill.p, gb.pts.p); lc.end(st);
  GLIO_CUDA_TRY(cudaGetLastError());
  g.cell_start = gb.cell_start.p;
  g.pts = gb.pts.p;
}

}  // namespace glio
