// glio_oracle_features.cpp — CPU restatement of the front end's feature extraction,
// Preprocessing::cloudHandler (GLIO/src/Preprocessing.cpp:529-655): LOAM-style curvature, per-ring / per-sector selection of
// sharp / less-sharp edge points and flat surf points with neighbour suppression, the remaining points as "less flat",
// down-sampled ring by ring with pcl::VoxelGrid (leaf ds_v).  TEST INFRASTRUCTURE ONLY (see glio_oracle.h).
//
// Input = `laserCloud` of the reference: the points of the scan lines concatenated ring after ring, x,y,z,intensity per
// point, with scanStartInd / scanEndInd per ring (:529-534).  Everything up to there (ring assignment from the vertical
// angle, IMU undistortion) is sensor plumbing and stays on the host.
// std::sort with `cloudCurvature[i] < cloudCurvature[j]` (:16, :554) leaves the order of equal curvatures to the
// implementation: order_mode 0 = literal std::sort, 1 = ties by index (the deterministic member the CUDA path reproduces).
// PCL (VoxelGrid<PointXYZI>, all fields averaged) is absent from /root/reference: restated as in glio_oracle.cpp.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "glio_oracle.h"

namespace {

// pcl::VoxelGrid<PointXYZI>::applyFilter on one ring cloud (x,y,z,intensity), stable order inside a voxel when order_mode = 1
int64_t voxel_filter_xyzi(const std::vector<float>& in /*4 per point*/, float leaf, int order_mode, std::vector<float>& out) {
  const int64_t n = (int64_t)in.size() / 4;
  if (n <= 0) return 0;
  const float inv = 1.0f / leaf;
  float mn[3] = {in[0], in[1], in[2]}, mx[3] = {in[0], in[1], in[2]};
  for (int64_t i = 1; i < n; ++i) for (int d = 0; d < 3; ++d) { mn[d] = std::min(mn[d], in[4 * i + d]); mx[d] = std::max(mx[d], in[4 * i + d]); }
  const int64_t dx = (int64_t)((mx[0] - mn[0]) * inv) + 1, dy = (int64_t)((mx[1] - mn[1]) * inv) + 1, dz = (int64_t)((mx[2] - mn[2]) * inv) + 1;
  if (dx * dy * dz > (int64_t)std::numeric_limits<int32_t>::max()) { out.insert(out.end(), in.begin(), in.end()); return n; }   // PCL: input passed through
  int min_b[3], max_b[3], div_b[3];
  for (int d = 0; d < 3; ++d) { min_b[d] = (int)std::floor(mn[d] * inv); max_b[d] = (int)std::floor(mx[d] * inv); div_b[d] = max_b[d] - min_b[d] + 1; }
  const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  struct Cpi { unsigned int idx; unsigned int pt; bool operator<(const Cpi& o) const { return idx < o.idx; } };
  std::vector<Cpi> iv; iv.reserve((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    const int i0 = (int)(std::floor(in[4 * i] * inv) - (float)min_b[0]);
    const int i1 = (int)(std::floor(in[4 * i + 1] * inv) - (float)min_b[1]);
    const int i2 = (int)(std::floor(in[4 * i + 2] * inv) - (float)min_b[2]);
    iv.push_back(Cpi{(unsigned int)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), (unsigned int)i});
  }
  if (order_mode == 0) std::sort(iv.begin(), iv.end()); else std::stable_sort(iv.begin(), iv.end());
  int64_t m = 0;
  for (size_t a = 0; a < iv.size();) {
    size_t b = a + 1;
    while (b < iv.size() && iv[b].idx == iv[a].idx) ++b;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    for (size_t k = a; k < b; ++k) for (int d = 0; d < 4; ++d) s[d] += in[4 * (size_t)iv[k].pt + d];
    const float cnt = (float)(b - a);
    for (int d = 0; d < 4; ++d) out.push_back(s[d] / cnt);
    ++m; a = b;
  }
  return m;
}

}  // namespace

extern "C" int go_extract_features(const float* cloud, int64_t n, int n_scans, const int32_t* scan_start, const int32_t* scan_end, int ds_rate,
                                   double edge_thres, double surf_thres, float ds_v, int order_mode, float* curvature, int8_t* label,
                                   int32_t* sharp, int64_t* n_sharp, int32_t* less_sharp, int64_t* n_less_sharp, int32_t* flat, int64_t* n_flat,
                                   int32_t* less_flat, int64_t* n_less_flat, float* less_flat_ds, int64_t* n_less_flat_ds, int32_t* ring_ds_count) {
  auto X = [&](int64_t i) { return cloud[4 * i]; };
  auto Y = [&](int64_t i) { return cloud[4 * i + 1]; };
  auto Z = [&](int64_t i) { return cloud[4 * i + 2]; };
  std::vector<float> curv((size_t)n, 0.f); std::vector<int> sortInd((size_t)n, 0), picked((size_t)n, 0), lab((size_t)n, 0);
  for (int64_t i = 5; i < n - 5; ++i) {                                                     // :537-546
    const float dX = X(i - 5) + X(i - 4) + X(i - 3) + X(i - 2) + X(i - 1) - 10 * X(i) + X(i + 1) + X(i + 2) + X(i + 3) + X(i + 4) + X(i + 5);
    const float dY = Y(i - 5) + Y(i - 4) + Y(i - 3) + Y(i - 2) + Y(i - 1) - 10 * Y(i) + Y(i + 1) + Y(i + 2) + Y(i + 3) + Y(i + 4) + Y(i + 5);
    const float dZ = Z(i - 5) + Z(i - 4) + Z(i - 3) + Z(i - 2) + Z(i - 1) - 10 * Z(i) + Z(i + 1) + Z(i + 2) + Z(i + 3) + Z(i + 4) + Z(i + 5);
    curv[i] = dX * dX + dY * dY + dZ * dZ; sortInd[i] = (int)i;
  }
  int64_t ns = 0, nls = 0, nf = 0, nlf = 0, nds = 0;
  auto mark = [&](int ind) {                                                                 // :584-603 / :622-641
    for (int l = 1; l <= 5; ++l) {
      const float dx = X(ind + l) - X(ind + l - 1), dy = Y(ind + l) - Y(ind + l - 1), dz = Z(ind + l) - Z(ind + l - 1);
      if (dx * dx + dy * dy + dz * dz > 0.05) break;
      picked[ind + l] = 1;
    }
    for (int l = -1; l >= -5; --l) {
      const float dx = X(ind + l) - X(ind + l + 1), dy = Y(ind + l) - Y(ind + l + 1), dz = Z(ind + l) - Z(ind + l + 1);
      if (dx * dx + dy * dy + dz * dz > 0.05) break;
      picked[ind + l] = 1;
    }
  };
  for (int i = 0; i < n_scans; ++i) {
    if (ring_ds_count) ring_ds_count[i] = 0;
    if (scan_end[i] - scan_start[i] < 6 || i % ds_rate != 0) continue;                      // :546-547
    std::vector<float> ring_less_flat;
    for (int j = 0; j < 6; ++j) {
      const int sp = scan_start[i] + (scan_end[i] - scan_start[i]) * j / 6;
      const int ep = scan_start[i] + (scan_end[i] - scan_start[i]) * (j + 1) / 6 - 1;
      if (order_mode == 0) std::sort(sortInd.begin() + sp, sortInd.begin() + ep + 1, [&](int a, int b) { return curv[a] < curv[b]; });
      else std::sort(sortInd.begin() + sp, sortInd.begin() + ep + 1, [&](int a, int b) { return curv[a] < curv[b] || (curv[a] == curv[b] && a < b); });
      int largest = 0;
      for (int k = ep; k >= sp; --k) {                                                      // :557-605
        const int ind = sortInd[k];
        if (picked[ind] == 0 && curv[ind] > edge_thres) {
          ++largest;
          if (largest <= 2) { lab[ind] = 2; sharp[ns++] = ind; less_sharp[nls++] = ind; }
          else if (largest <= 10) { lab[ind] = 1; less_sharp[nls++] = ind; }
          else break;
          picked[ind] = 1;
          mark(ind);
        }
      }
      int smallest = 0;
      for (int k = sp; k <= ep; ++k) {                                                      // :607-643
        const int ind = sortInd[k];
        if (X(ind) * X(ind) + Y(ind) * Y(ind) + Z(ind) * Z(ind) < 0.25) continue;
        if (picked[ind] == 0 && curv[ind] < surf_thres) {
          lab[ind] = -1; flat[nf++] = ind;
          ++smallest;
          if (smallest >= 4) break;
          picked[ind] = 1;
          mark(ind);
        }
      }
      for (int k = sp; k <= ep; ++k) {                                                      // :645-650
        if (X(k) * X(k) + Y(k) * Y(k) + Z(k) * Z(k) < 0.25) continue;
        if (lab[k] <= 0) { less_flat[nlf++] = k; for (int d = 0; d < 4; ++d) ring_less_flat.push_back(cloud[4 * (int64_t)k + d]); }
      }
    }
    std::vector<float> ds;                                                                  // :653-659
    const int64_t m = voxel_filter_xyzi(ring_less_flat, ds_v, order_mode, ds);
    if (less_flat_ds) std::memcpy(less_flat_ds + 4 * nds, ds.data(), sizeof(float) * ds.size());
    if (ring_ds_count) ring_ds_count[i] = (int32_t)m;
    nds += m;
  }
  if (curvature) std::memcpy(curvature, curv.data(), sizeof(float) * (size_t)n);
  if (label) for (int64_t k = 0; k < n; ++k) label[k] = (int8_t)lab[k];
  *n_sharp = ns; *n_less_sharp = nls; *n_flat = nf; *n_less_flat = nlf; *n_less_flat_ds = nds;
  return 0;
}
