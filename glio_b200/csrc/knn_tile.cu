// knn_tile.cu — K1a, staged tile search (GLIO_KNN_MODE=4): exact 5-NN over the local map for
// Estimator::findCorrespondingSurfFeatures (GLIO/src/Estimator.cpp:3643-3651, kd_tree->nearestKSearch(point, 5, ...)).
//
// Design (the north_star's "TMA / shared-memory staging of local-map tiles"):
//   * queries arrive sorted by grid cell, so the 32 queries of a warp sit in a short run of x-adjacent cells of one (y,z)
//     cell row.  The warp owns that TILE: the 3x3 cell rows around it, each row ONE contiguous, 32-byte aligned run of the
//     pair-layout map (PairRec), are bulk-copied global -> shared by the copy engine (cp.async.bulk + mbarrier
//     complete_tx; SASS UBLKCP), nine copies issued by nine lanes, no register staging, no per-thread gathers.
//   * every lane then scans the SAME staged candidate list with uniform trip counts (broadcast LDS, no divergence):
//       pass A  packed fp32x2 distances -> 8 interleaved running minima; the 5th smallest of them bounds the 5th
//               neighbour distance from above (five distinct candidates are at least that close)
//       pass B  packed distances again, candidates under the bound (x 1+1e-6: the packed arithmetic is fused) set a bit in
//               a per-lane survivor mask (16 candidates per word, parked in shared memory; about 6-9 survive of ~190)
//       flush   the mask bits are compacted into a dense per-lane list (cheap divergent loop), then the survivors are
//               re-evaluated with the exact FLANN L2_Simple<float> order and inserted in the (distance, index) top-5
//               -> the only place a result is decided, so indices and distances stay bit-exact
//   * a lane whose 5th distance is provably inside the scanned box is final; the rest (sparse map, query displaced from the
//     surface by the pose error, oversized tile) are DEFERRED with their 5th distance as a proven search radius.
//   * second pass (k_knn_tile2): the deferred queries, compacted IN SORTED ORDER, form tiles again; a tile's box is the union
//     of its lanes' search spheres (so every lane is final afterwards), staged nine rows at a time; no pass A is needed
//     because the radius is already a bound.  What does not fit the staging buffer falls back to
//   * the team pass (k_knn_team): 8 lanes per query split the cell rows of the sphere's bounding box, private top-5 each,
//     merged with three shuffle rounds (half-cleaner + 5-sorter).  Normally (almost) empty.
// Compiled with -fmad=false like assoc.cu; the packed helpers are pre-filters only (knn_common.cuh).
#include "knn_common.cuh"

namespace glio {

constexpr int TK_WARPS = 4;          // warps (= tiles in flight) per CTA
constexpr int TK_CAP_PAIRS = 224;    // staged pair records per tile (448 map points, 7 KB); larger tiles are split / deferred
constexpr int TK_WORDS = TK_CAP_PAIRS / 8 + 1;
constexpr int TK_LIST = 16;          // per-lane survivor list capacity (more survivors are pushed straight from the mask loop)
constexpr int TK_SPAN = 12;          // max x-extent (cells) of the queries sharing one tile (first pass)
constexpr int TK_SPAN2 = 8;          // same, second pass
constexpr float TK_MARGIN = 1.000002f;
constexpr float TK_WIDE_CELLS = 2.5f;  // second pass: search radius (in cells) above which a query is handed to the team pass
constexpr int TK_SMEM_WARP = (TK_CAP_PAIRS + 8) * 32 + TK_WORDS * 64 + TK_LIST * 64;   // bytes per warp
constexpr int TK_SMEM = TK_WARPS * TK_SMEM_WARP + TK_WARPS * 8;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void cmpx(float& a, float& b) { const float lo = fminf(a, b), hi = fmaxf(a, b); a = lo; b = hi; }
// 5th smallest of eight values (optimal 19-comparator network; unused outputs are dead code)
__device__ __forceinline__ float fifth_of_8(float m0, float m1, float m2, float m3, float m4, float m5, float m6, float m7) {
  cmpx(m0, m2); cmpx(m1, m3); cmpx(m4, m6); cmpx(m5, m7);
  cmpx(m0, m4); cmpx(m1, m5); cmpx(m2, m6); cmpx(m3, m7);
  cmpx(m0, m1); cmpx(m2, m3); cmpx(m4, m5); cmpx(m6, m7);
  cmpx(m2, m4); cmpx(m3, m5);
  cmpx(m1, m4); cmpx(m3, m6);
  cmpx(m1, m2); cmpx(m3, m4); cmpx(m5, m6);
  return m4;
}

// per-warp staging area
struct WarpTile {
  PairRec* tile;              // [TK_CAP_PAIRS + 8]
  uint16_t* mask;             // [TK_WORDS][32], already offset by the lane
  uint16_t* list;             // [TK_LIST][32], already offset by the lane
  unsigned long long* bar;
  unsigned parity;
};

__device__ __forceinline__ WarpTile warp_tile_init(unsigned char* smem, int wid, int lane) {
  WarpTile wt;
  unsigned char* base = smem + (size_t)wid * TK_SMEM_WARP;
  wt.tile = reinterpret_cast<PairRec*>(base);
  wt.mask = reinterpret_cast<uint16_t*>(base + (TK_CAP_PAIRS + 8) * 32) + lane;
  wt.list = reinterpret_cast<uint16_t*>(base + (TK_CAP_PAIRS + 8) * 32 + TK_WORDS * 64) + lane;
  wt.bar = reinterpret_cast<unsigned long long*>(smem + (size_t)TK_WARPS * TK_SMEM_WARP) + wid;
  wt.parity = 0;
  if (lane == 0) {
    mbar_init(wt.bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  fence_proxy_async();
  __syncwarp();
  return wt;
}

// Stage up to nine cell rows: lanes 0..8 hold one row's point range [s, e) of the sorted map each (s == e: nothing).  All 32
// lanes must call.  Returns the padded number of pair records staged (a multiple of 8; 0: nothing to scan), or -1 when the
// rows do not fit the buffer (nothing staged).
__device__ __forceinline__ int stage_rows(WarpTile& wt, const PairRec* __restrict__ gpairs, int lane, int s, int e) {
  const int ps = s >> 1;
  const int np = e > s ? ((e + 1) >> 1) - ps : 0;
  int off = np;
#pragma unroll
  for (int o = 1; o < 16; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, off, o); if (lane >= o) off += v; }
  const int total = __shfl_sync(0xffffffffu, off, 15);
  off -= np;
  if (total > TK_CAP_PAIRS) return -1;
  if (total == 0) return 0;
  fence_proxy_async();              // earlier generic-proxy accesses of this buffer (previous tile) before the async writes
  __syncwarp();
  if (lane == 0) mbar_expect_tx(wt.bar, (unsigned)total * 32u);
  __syncwarp();
  if (np > 0) bulk_g2s(wt.tile + off, gpairs + ps, (unsigned)np * 32u, wt.bar);
  mbar_wait(wt.bar, wt.parity);
  wt.parity ^= 1u;
  // the pair run of a row may start / end with a half-record of the neighbouring cell: neutralise it
  if (np > 0) {
    if (s & 1) { wt.tile[off].x0 = PAIR_FAR; wt.tile[off].i0 = 0x7fffffff; }
    if (e & 1) { wt.tile[off + np - 1].x1 = PAIR_FAR; wt.tile[off + np - 1].i1 = 0x7fffffff; }
  }
  const int npad = (total + 7) & ~7;
  if (lane < npad - total) {
    PairRec f; f.x0 = PAIR_FAR; f.x1 = PAIR_FAR; f.y0 = 0.f; f.y1 = 0.f; f.z0 = 0.f; f.z1 = 0.f; f.i0 = 0x7fffffff; f.i1 = 0x7fffffff;
    wt.tile[total + lane] = f;
  }
  __syncwarp();
  return npad;
}

// packed squared distances of the query to the two points of a pair record (fused arithmetic: a pre-filter, not a result)
__device__ __forceinline__ void pair_dist(const ulonglong2* T2, int k, f32x2_t nqx2, f32x2_t nqy2, f32x2_t nqz2, float& d0, float& d1) {
  const f32x2_t one2 = 0x3f8000003f800000ull;
  const ulonglong2 xy = T2[2 * k];
  const unsigned long long zz = *reinterpret_cast<const unsigned long long*>(&T2[2 * k + 1]);
  // p - q through the FMA pipe (p * 1 + (-q)); the ALU pipe is the busy one in this kernel
  const f32x2_t dx = fma2(xy.x, one2, nqx2), dy = fma2(xy.y, one2, nqy2), dz = fma2(zz, one2, nqz2);
  unpack2(fma2(dz, dz, fma2(dy, dy, mul2(dx, dx))), d0, d1);
}

struct QueryRegs { float qx, qy, qz; f32x2_t nqx2, nqy2, nqz2; };
__device__ __forceinline__ QueryRegs make_query(float qx, float qy, float qz) {
  QueryRegs q; q.qx = qx; q.qy = qy; q.qz = qz; q.nqx2 = pack2(-qx, -qx); q.nqy2 = pack2(-qy, -qy); q.nqz2 = pack2(-qz, -qz); return q;
}

// exact distance of staged candidate j to the query, pushed into the top-5
__device__ __forceinline__ void push_exact(const WarpTile& wt, int j, const QueryRegs& q, Top5& tt) {
  const float* rec = reinterpret_cast<const float*>(wt.tile) + (j >> 1) * 8 + (j & 1);
  const int id = __float_as_int(rec[6]);
  top5_push(tt, l2_simple(q.qx, q.qy, q.qz, rec[0], rec[2], rec[4]), id);
}

// pass B + flush for the calling lane (no warp-level synchronisation inside: callers may be a subset of the warp)
__device__ __forceinline__ void filter_and_flush(const WarpTile& wt, int npad, const QueryRegs& q, float thr, Top5& tt) {
  const ulonglong2* T2 = reinterpret_cast<const ulonglong2*>(wt.tile);
  const int nwords = npad >> 3;
#pragma unroll 1
  for (int w = 0; w < nwords; ++w) {
    unsigned m = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      float d0, d1;
      pair_dist(T2, 8 * w + u, q.nqx2, q.nqy2, q.nqz2, d0, d1);
      if (d0 <= thr) m |= 1u << (2 * u);
      if (d1 <= thr) m |= 2u << (2 * u);
    }
    wt.mask[w * 32] = (uint16_t)m;
  }
  // mask bits -> dense per-lane list (short divergent loop), then a dense exact pass
  int cnt = 0;
#pragma unroll 1
  for (int w = 0; w < nwords; ++w) {
    unsigned m = wt.mask[w * 32];
    while (m) {
      const int j = 16 * w + __ffs(m) - 1;
      m &= m - 1;
      if (cnt < TK_LIST) { wt.list[cnt * 32] = (uint16_t)j; ++cnt; }
      else push_exact(wt, j, q, tt);                       // list full (rare: loose bound): straight from the mask loop
    }
  }
#pragma unroll 1
  for (int i = 0; i < cnt; ++i) push_exact(wt, wt.list[i * 32], q, tt);
}

// ---------------------------------------------------------------------------------------------------------------------
// first pass: one warp per 32 consecutive (cell-sorted) queries
// ---------------------------------------------------------------------------------------------------------------------
struct TileArgs {
  SearchArgs sa;
  const PairRec* pairs;
  unsigned int* dmask;      // [nwarps] lanes deferred by each warp of the first pass
  int* dcount;              // [nwarps + 1] their number (scanned into dpos)
  const int* dpos;          // [nwarps + 1] exclusive scan of dcount; dpos[nwarps] = number of deferred queries
  int nwarps;
  uint32_t* fb_list;        // fallback list of the second pass (team pass input)
  unsigned int* n_fb;
};

__global__ void __launch_bounds__(32 * TK_WARPS) k_knn_tile(TileArgs ta) {
  extern __shared__ __align__(128) unsigned char tk_smem[];
  const SearchArgs& a = ta.sa;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int gw = blockIdx.x * TK_WARPS + wid;
  const int64_t p = (int64_t)gw * 32 + lane;
  const bool active = p < a.Qt;
  const GridDesc& G = a.grid;
  const float INF = __int_as_float(0x7f800000);
  WarpTile wt = warp_tile_init(tk_smem, wid, lane);

  float qx = 0.f, qy = 0.f, qz = 0.f;
  if (active) { const float4 q4 = a.pmq[p]; qx = q4.x; qy = q4.y; qz = q4.z; }
  const QueryRegs q = make_query(qx, qy, qz);
  const int cx = cell_coord(qx, G.ox, G.inv_cell), cy = cell_coord(qy, G.oy, G.inv_cell), cz = cell_coord(qz, G.oz, G.inv_cell);
  const float gate_r = sqrtf(a.gate_sq) + 2e-3f;
  const int rmax = (int)ceilf(gate_r * G.inv_cell) + 1;
  // far outside the grid: nothing within the gate radius
  const bool far_out = cx < -rmax || cy < -rmax || cz < -rmax || cx >= G.nx + rmax || cy >= G.ny + rmax || cz >= G.nz + rmax;
  if (active && far_out) { Top5 t; top5_init(t); store_top5(a, p, t); }
  bool done = !active || far_out;
  bool deferred = false;
  const float gate_cap = a.gate_sq * TK_MARGIN;
  const int* __restrict__ cs = G.cell_start;
  int span_lim = TK_SPAN;

  for (;;) {
    const unsigned pending = __ballot_sync(0xffffffffu, !done);
    if (!pending) break;
    const int leader = __ffs(pending) - 1;
    const int lcx = __shfl_sync(0xffffffffu, cx, leader), lcy = __shfl_sync(0xffffffffu, cy, leader), lcz = __shfl_sync(0xffffffffu, cz, leader);
    const bool part = !done && cy == lcy && cz == lcz && cx >= lcx && cx < lcx + span_lim;
    const int hix = __reduce_max_sync(0xffffffffu, part ? cx : lcx);
    const int xa = lcx - 1, xb = hix + 1;                       // scanned cell range in x
    const int x0 = max(xa, 0), x1 = min(xb, G.nx - 1);
    // ---- the nine rows of the tile: bounds (lanes 0..8), staging
    int s = 0, e = 0;
    if (lane < 9) {
      const int z = lcz + lane / 3 - 1, y = lcy + lane % 3 - 1;
      if (z >= 0 && z < G.nz && y >= 0 && y < G.ny && x0 <= x1) {
        const int row = (z * G.ny + y) * G.nx;
        s = __ldg(&cs[row + x0]); e = __ldg(&cs[row + x1 + 1]);
      }
    }
    const int npad = stage_rows(wt, ta.pairs, lane, s, e);
    if (npad < 0 && hix > lcx) { span_lim = max(1, (hix - lcx + 1) >> 1); continue; }   // too many candidates: narrower tile
    span_lim = TK_SPAN;
    bool final_ok = false;
    Top5 tt;
    top5_init(tt);
    if (npad > 0 && part) {
      const ulonglong2* T2 = reinterpret_cast<const ulonglong2*>(wt.tile);
      // ---- pass A: eight interleaved running minima of the packed distances
      float m0 = INF, m1 = INF, m2 = INF, m3 = INF, m4 = INF, m5 = INF, m6 = INF, m7 = INF;
#pragma unroll 1
      for (int k = 0; k < npad; k += 4) {
        float d[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) pair_dist(T2, k + u, q.nqx2, q.nqy2, q.nqz2, d[2 * u], d[2 * u + 1]);
        m0 = fminf(m0, d[0]); m1 = fminf(m1, d[1]); m2 = fminf(m2, d[2]); m3 = fminf(m3, d[3]);
        m4 = fminf(m4, d[4]); m5 = fminf(m5, d[5]); m6 = fminf(m6, d[6]); m7 = fminf(m7, d[7]);
      }
      // five distinct candidates lie within the 5th smallest minimum: an upper bound of the exact 5th distance once the
      // fused packed arithmetic (<= 4e-7 relative from the unfused order) is covered by the margin; nothing beyond the
      // radius gate can matter (Estimator.cpp:3651)
      const float bound = fminf(fifth_of_8(m0, m1, m2, m3, m4, m5, m6, m7) * TK_MARGIN, gate_cap);
      filter_and_flush(wt, npad, q, bound * TK_MARGIN, tt);
      // distance to the faces of the scanned box; faces clipped by the grid are infinitely far (nothing lives outside)
      float b = INF;
      if (xa > 0)               b = fminf(b, qx - (G.ox + (float)xa * G.cell));
      if (xb < G.nx - 1)        b = fminf(b, (G.ox + (float)(xb + 1) * G.cell) - qx);
      if (lcy - 1 > 0)          b = fminf(b, qy - (G.oy + (float)(lcy - 1) * G.cell));
      if (lcy + 1 < G.ny - 1)   b = fminf(b, (G.oy + (float)(lcy + 2) * G.cell) - qy);
      if (lcz - 1 > 0)          b = fminf(b, qz - (G.oz + (float)(lcz - 1) * G.cell));
      if (lcz + 1 < G.nz - 1)   b = fminf(b, (G.oz + (float)(lcz + 2) * G.cell) - qz);
      const float bs = b * 0.999f - 2e-3f;   // safety: float rounding of cell assignment / face positions
      final_ok = (b == INF) || (bs > 0.f && key_dist(tt.k4) <= bs * bs) || bs >= gate_r;
    }
    if (part) {
      done = true;
      store_top5(a, p, tt);                       // deferred lanes: the second pass reads the 5th distance as its search radius
      deferred = !final_ok;
    }
  }
  const unsigned dm = __ballot_sync(0xffffffffu, deferred);
  if (lane == 0) { ta.dmask[gw] = dm; ta.dcount[gw] = __popc(dm); }
}

// deferred queries in sorted order: warp w's lanes land at dpos[w] ...
__global__ void __launch_bounds__(256) k_defer_scatter(TileArgs ta) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w == 0) { const unsigned nd = (unsigned)ta.dpos[ta.nwarps]; *ta.sa.n_deferred = nd; if (ta.sa.n_fallback) atomicAdd(ta.sa.n_fallback, (unsigned long long)nd); }
  if (w >= ta.nwarps) return;
  unsigned m = ta.dmask[w];
  int o = ta.dpos[w];
  while (m) { const int b = __ffs(m) - 1; m &= m - 1; ta.sa.deferred[o++] = (uint32_t)(w * 32 + b); }
}

// ---------------------------------------------------------------------------------------------------------------------
// second pass: tiles of deferred queries; a tile's box is the union of its lanes' search spheres
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32 * TK_WARPS) k_knn_tile2(TileArgs ta) {
  extern __shared__ __align__(128) unsigned char tk_smem[];
  const SearchArgs& a = ta.sa;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const GridDesc& G = a.grid;
  const int* __restrict__ cs = G.cell_start;
  const float gate_cap = a.gate_sq * TK_MARGIN;
  const unsigned int nd = (unsigned int)ta.dpos[ta.nwarps];
  const unsigned int nwarps = gridDim.x * TK_WARPS;
  WarpTile wt = warp_tile_init(tk_smem, wid, lane);
  const int BIG = 0x3fffffff;
  for (unsigned int base = (blockIdx.x * TK_WARPS + wid) * 32u; base < nd; base += nwarps * 32u) {
    const unsigned int item = base + lane;
    const bool valid = item < nd;
    int64_t p = 0;
    float qx = 0.f, qy = 0.f, qz = 0.f, T = 0.f;
    if (valid) {
      p = a.deferred[item];
      const float4 q4 = a.pmq[p];
      qx = q4.x; qy = q4.y; qz = q4.z;
      T = fminf(a.knn_sqd[4 * a.Qt + p], gate_cap);       // exact 5th distance inside the first pass's box (a proven bound), or the gate
    }
    const QueryRegs q = make_query(qx, qy, qz);
    const float r = sqrtf(T) * 1.001f + 2e-3f;            // margin: float rounding of the cell assignment
    const int cx = cell_coord(qx, G.ox, G.inv_cell), cy = cell_coord(qy, G.oy, G.inv_cell), cz = cell_coord(qz, G.oz, G.inv_cell);
    const int xlo = cell_coord(qx - r, G.ox, G.inv_cell), xhi = cell_coord(qx + r, G.ox, G.inv_cell);
    const int ylo = cell_coord(qy - r, G.oy, G.inv_cell), yhi = cell_coord(qy + r, G.oy, G.inv_cell);
    const int zlo = cell_coord(qz - r, G.oz, G.inv_cell), zhi = cell_coord(qz + r, G.oz, G.inv_cell);
    const float thr = T * TK_MARGIN;
    // a wide sphere would blow up the box of every lane sharing its tile: those queries go straight to the team pass
    const bool wide = valid && r > TK_WIDE_CELLS * G.cell;
    {
      const unsigned wm = __ballot_sync(0xffffffffu, wide);
      if (wm) {
        unsigned int fb = 0;
        if (lane == 0) fb = atomicAdd(ta.n_fb, (unsigned int)__popc(wm));
        fb = __shfl_sync(0xffffffffu, fb, 0);
        if (wide) ta.fb_list[fb + __popc(wm & ((1u << lane) - 1u))] = (uint32_t)p;
      }
    }
    bool done = !valid || wide;
    for (;;) {
      const unsigned pending = __ballot_sync(0xffffffffu, !done);
      if (!pending) break;
      const int leader = __ffs(pending) - 1;
      const int lcx = __shfl_sync(0xffffffffu, cx, leader), lcy = __shfl_sync(0xffffffffu, cy, leader), lcz = __shfl_sync(0xffffffffu, cz, leader);
      const bool part = !done && cy == lcy && cz == lcz && cx >= lcx && cx < lcx + TK_SPAN2;
      const int X0 = max(__reduce_min_sync(0xffffffffu, part ? xlo : BIG), 0), X1 = min(__reduce_max_sync(0xffffffffu, part ? xhi : -BIG), G.nx - 1);
      const int Y0 = max(__reduce_min_sync(0xffffffffu, part ? ylo : BIG), 0), Y1 = min(__reduce_max_sync(0xffffffffu, part ? yhi : -BIG), G.ny - 1);
      const int Z0 = max(__reduce_min_sync(0xffffffffu, part ? zlo : BIG), 0), Z1 = min(__reduce_max_sync(0xffffffffu, part ? zhi : -BIG), G.nz - 1);
      const int nyr = Y1 - Y0 + 1;
      const int nrows = (X0 <= X1 && nyr > 0 && Z1 >= Z0) ? nyr * (Z1 - Z0 + 1) : 0;
      Top5 tt;
      top5_init(tt);
      bool overflow = false;
      for (int g = 0; g < nrows; g += 9) {
        int s = 0, e = 0;
        if (lane < 9 && g + lane < nrows) {
          const int ri = g + lane;
          const int row = ((Z0 + ri / nyr) * G.ny + (Y0 + ri % nyr)) * G.nx;
          s = __ldg(&cs[row + X0]); e = __ldg(&cs[row + X1 + 1]);
        }
        const int npad = stage_rows(wt, ta.pairs, lane, s, e);
        if (npad < 0) { overflow = true; break; }
        if (npad > 0 && part) filter_and_flush(wt, npad, q, thr, tt);
        __syncwarp();
      }
      if (part) {
        done = true;
        if (!overflow) store_top5(a, p, tt);
      }
      const unsigned fm = __ballot_sync(0xffffffffu, part && overflow);
      if (fm) {
        unsigned int fb = 0;
        if (lane == 0) fb = atomicAdd(ta.n_fb, (unsigned int)__popc(fm));
        fb = __shfl_sync(0xffffffffu, fb, 0);
        if (part) ta.fb_list[fb + __popc(fm & ((1u << lane) - 1u))] = (uint32_t)p;
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// team pass (fallback): 8 lanes per query
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long shfl_xor_u64(unsigned long long v, int m) {
  const unsigned lo = __shfl_xor_sync(0xffffffffu, (unsigned)v, m), hi = __shfl_xor_sync(0xffffffffu, (unsigned)(v >> 32), m);
  return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long kmin(unsigned long long x, unsigned long long y) { return x < y ? x : y; }

__global__ void __launch_bounds__(128, 12) k_knn_team(SearchArgs a, const uint32_t* __restrict__ qlist, const unsigned int* __restrict__ n_list) {
  const int lane = threadIdx.x & 31, team = lane >> 3, tl = lane & 7;
  const unsigned int nwarps = gridDim.x * (blockDim.x >> 5);
  const unsigned int nd = *n_list;
  const GridDesc& G = a.grid;
  const int* __restrict__ cs = G.cell_start;
  const float gate_cap = a.gate_sq * TK_MARGIN;
  for (unsigned int base = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4u; base < nd; base += nwarps * 4u) {
    const unsigned int item = base + team;
    const bool valid = item < nd;
    int64_t p = 0;
    float qx = 0.f, qy = 0.f, qz = 0.f, T = 0.f;
    if (valid) {
      p = qlist[item];
      const float4 q4 = a.pmq[p];
      qx = q4.x; qy = q4.y; qz = q4.z;
      T = fminf(a.knn_sqd[4 * a.Qt + p], gate_cap);      // the first pass's 5th distance inside its box, or +inf
    }
    Top5 t;
    top5_init(t);
    float d4f = T;
    if (valid) {
      const float r = sqrtf(T) * 1.001f + 2e-3f;
      const int ylo = max(cell_coord(qy - r, G.oy, G.inv_cell), 0), yhi = min(cell_coord(qy + r, G.oy, G.inv_cell), G.ny - 1);
      const int zlo = max(cell_coord(qz - r, G.oz, G.inv_cell), 0), zhi = min(cell_coord(qz + r, G.oz, G.inv_cell), G.nz - 1);
      const int nyr = yhi - ylo + 1, nzr = zhi - zlo + 1;
      const int nrows = (nyr > 0 && nzr > 0) ? nyr * nzr : 0;
      for (int ri = tl; ri < nrows; ri += 8) {
        const int z = zlo + ri / nyr, y = ylo + ri % nyr;
        // squared distance from the query to the (y,z) rectangle of this cell row, shrunk by the rounding margin
        const float yl = G.oy + (float)y * G.cell, zl = G.oz + (float)z * G.cell;
        const float dy = fmaxf(fmaxf(yl - qy, qy - (yl + G.cell)), 0.f), dz = fmaxf(fmaxf(zl - qz, qz - (zl + G.cell)), 0.f);
        const float dys = fmaxf(dy * 0.999f - 2e-3f, 0.f), dzs = fmaxf(dz * 0.999f - 2e-3f, 0.f);
        const float rem = d4f - (dys * dys + dzs * dzs);
        if (rem < 0.f) continue;
        const float rx = sqrtf(rem) * 1.001f + 2e-3f;
        const int xl = max(cell_coord(qx - rx, G.ox, G.inv_cell), 0), xh = min(cell_coord(qx + rx, G.ox, G.inv_cell), G.nx - 1);
        if (xl > xh) continue;
        const int row = (z * G.ny + y) * G.nx;
        const int s = __ldg(&cs[row + xl]), e = __ldg(&cs[row + xh + 1]);
        for (int k = s; k < e; k += 4) {
          // four independent loads in flight per lane (the loop is latency-bound otherwise); the tail is padded with far fillers
          float4 c[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) c[u] = (k + u < e) ? __ldg(&G.pts[k + u]) : make_float4(PAIR_FAR, 0.f, 0.f, __int_as_float(0x7fffffff));
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float d = l2_simple(qx, qy, qz, c[u].x, c[u].y, c[u].z);
            if (d <= d4f) {
              const unsigned long long key = make_key(d, __float_as_int(c[u].w));
              if (key < t.k4) {
                t.k4 = key; GLIO_KSWAP(t.k3, t.k4) GLIO_KSWAP(t.k2, t.k3) GLIO_KSWAP(t.k1, t.k2) GLIO_KSWAP(t.k0, t.k1)
                d4f = fminf(d4f, key_dist(t.k4));
              }
            }
          }
        }
      }
    }
    // merge the eight private lists: the five smallest of two sorted 5-lists are min(a_i, b_{4-i}); re-sort; three rounds
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) {
      const unsigned long long b0 = shfl_xor_u64(t.k0, m), b1 = shfl_xor_u64(t.k1, m), b2 = shfl_xor_u64(t.k2, m), b3 = shfl_xor_u64(t.k3, m), b4 = shfl_xor_u64(t.k4, m);
      t.k0 = kmin(t.k0, b4); t.k1 = kmin(t.k1, b3); t.k2 = kmin(t.k2, b2); t.k3 = kmin(t.k3, b1); t.k4 = kmin(t.k4, b0);
      GLIO_KSWAP(t.k0, t.k1) GLIO_KSWAP(t.k3, t.k4) GLIO_KSWAP(t.k2, t.k4) GLIO_KSWAP(t.k2, t.k3) GLIO_KSWAP(t.k1, t.k4)
      GLIO_KSWAP(t.k0, t.k3) GLIO_KSWAP(t.k0, t.k2) GLIO_KSWAP(t.k1, t.k3) GLIO_KSWAP(t.k1, t.k2)
    }
    if (valid && tl == 0) store_top5(a, p, t);
  }
}

void knn_tile_run(const SearchArgs& sa, const PairRec* d_pairs, DevBuf<int>& work, DevBuf<int>& scan_tmp, cudaStream_t st, LaunchCounter& lc) {
  GLIO_REQUIRE(d_pairs != nullptr, GLIO_ERR_STATE, "knn_tile_run: the grid has no pair layout (build_pairs)");
  static bool attr_done = false;
  if (!attr_done) {
    GLIO_CUDA_TRY(cudaFuncSetAttribute(k_knn_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, TK_SMEM));
    GLIO_CUDA_TRY(cudaFuncSetAttribute(k_knn_tile2, cudaFuncAttributeMaxDynamicSharedMemorySize, TK_SMEM));
    attr_done = true;
  }
  const unsigned nb = (unsigned)((sa.Qt + 32 * TK_WARPS - 1) / (32 * TK_WARPS));
  const int nw = (int)nb * TK_WARPS;
  // work layout: dmask[nw] | dcount[nw + 1] | dpos[nw + 1] | n_fb
  work.reserve((size_t)3 * nw + 8);
  TileArgs ta;
  ta.sa = sa; ta.pairs = d_pairs; ta.nwarps = nw;
  ta.dmask = reinterpret_cast<unsigned int*>(work.p); ta.dcount = work.p + nw; ta.dpos = work.p + 2 * nw + 1;
  ta.n_fb = reinterpret_cast<unsigned int*>(work.p + 3 * nw + 2); ta.fb_list = sa.deferred + sa.Qt;     // second half of the 2*Qt list buffer
  GLIO_CUDA_TRY(cudaMemsetAsync(ta.dcount + nw, 0, sizeof(int), st));
  GLIO_CUDA_TRY(cudaMemsetAsync(ta.n_fb, 0, sizeof(unsigned int), st));
  lc.begin("k_knn_tile", st); k_knn_tile<<<nb, 32 * TK_WARPS, TK_SMEM, st>>>(ta); lc.end(st);
  exclusive_scan_i32(ta.dcount, work.p + 2 * nw + 1, (int64_t)nw + 1, scan_tmp, st, lc);
  lc.begin("k_defer_scatter", st); k_defer_scatter<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(ta); lc.end(st);
  lc.begin("k_knn_tile2", st); k_knn_tile2<<<148 * 4, 32 * TK_WARPS, TK_SMEM, st>>>(ta); lc.end(st);
  lc.begin("k_knn_team", st); k_knn_team<<<148 * 4, 128, 0, st>>>(sa, ta.fb_list, ta.n_fb); lc.end(st);
  GLIO_CUDA_TRY(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------------
// GLIO_KNN_MODE=5: the per-thread box search of assoc.cu (k_knn_box) split in two launches.  In the fused kernel the face
// growth runs at about 5 of 32 lanes (one query in ten needs it, each warp waits for its few growing lanes) and costs two
// fifths of all instructions; here the first launch stops after the 3x3x3 start box, the queries with an open face are
// compacted IN SORTED ORDER (k_defer_scatter), and the second launch grows them with every lane busy.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 6) k_knn_box_start(TileArgs ta) {
  __shared__ int sbnd[18 * 128];
  const SearchArgs& a = ta.sa;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const GridDesc& g = a.grid;
  bool open = false;
  if (p < a.Qt) {
    const float4 q4 = a.pmq[p];
    const float qx = q4.x, qy = q4.y, qz = q4.z;
    Top5 t; top5_init(t);
    const int cx = cell_coord(qx, g.ox, g.inv_cell), cy = cell_coord(qy, g.oy, g.inv_cell), cz = cell_coord(qz, g.oz, g.inv_cell);
    const float gate_r = sqrtf(a.gate_sq) + 2e-3f;
    const int rmax = (int)ceilf(gate_r * g.inv_cell) + 1;
    const bool far_out = cx < -rmax || cy < -rmax || cz < -rmax || cx >= g.nx + rmax || cy >= g.ny + rmax || cz >= g.nz + rmax;
    float d4f = __int_as_float(0x7f800000);
    if (!far_out) {
      thread_rings(g, qx, qy, qz, cx, cy, cz, 1, 1, rmax, t, d4f, sbnd + threadIdx.x);
#pragma unroll 1
      for (int f = 0; f < 6 && !open; ++f) open = box_face_open(g, f, qx, qy, qz, gate_r, cx - 1, cx + 1, cy - 1, cy + 1, cz - 1, cz + 1, t);
    }
    // the second launch rebuilds the keys from all five (index, distance) pairs
    a.knn_idx[0 * a.Qt + p] = key_idx(t.k0); a.knn_idx[1 * a.Qt + p] = key_idx(t.k1); a.knn_idx[2 * a.Qt + p] = key_idx(t.k2);
    a.knn_idx[3 * a.Qt + p] = key_idx(t.k3); a.knn_idx[4 * a.Qt + p] = key_idx(t.k4);
    a.knn_sqd[4 * a.Qt + p] = key_dist(t.k4);
    if (open || a.store_all_sqd) {
      a.knn_sqd[0 * a.Qt + p] = key_dist(t.k0); a.knn_sqd[1 * a.Qt + p] = key_dist(t.k1); a.knn_sqd[2 * a.Qt + p] = key_dist(t.k2);
      a.knn_sqd[3 * a.Qt + p] = key_dist(t.k3);
    }
  }
  const unsigned dm = __ballot_sync(0xffffffffu, open);
  const int gw = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (lane == 0) { ta.dmask[gw] = dm; ta.dcount[gw] = __popc(dm); }
}

__global__ void __launch_bounds__(128, 6) k_knn_grow(TileArgs ta) {
  const SearchArgs& a = ta.sa;
  const GridDesc& g = a.grid;
  const unsigned int nd = (unsigned int)ta.dpos[ta.nwarps];
  const float gate_r = sqrtf(a.gate_sq) + 2e-3f;
  const int rmax = (int)ceilf(gate_r * g.inv_cell) + 1;
  for (unsigned int item = blockIdx.x * blockDim.x + threadIdx.x; item < nd; item += gridDim.x * blockDim.x) {
    const int64_t p = a.deferred[item];
    const float4 q4 = a.pmq[p];
    const float qx = q4.x, qy = q4.y, qz = q4.z;
    const int cx = cell_coord(qx, g.ox, g.inv_cell), cy = cell_coord(qy, g.oy, g.inv_cell), cz = cell_coord(qz, g.oz, g.inv_cell);
    Top5 t;
    t.k0 = make_key(a.knn_sqd[0 * a.Qt + p], a.knn_idx[0 * a.Qt + p]); t.k1 = make_key(a.knn_sqd[1 * a.Qt + p], a.knn_idx[1 * a.Qt + p]);
    t.k2 = make_key(a.knn_sqd[2 * a.Qt + p], a.knn_idx[2 * a.Qt + p]); t.k3 = make_key(a.knn_sqd[3 * a.Qt + p], a.knn_idx[3 * a.Qt + p]);
    t.k4 = make_key(a.knn_sqd[4 * a.Qt + p], a.knn_idx[4 * a.Qt + p]);
    float d4f = key_dist(t.k4);
    box_grow(g, qx, qy, qz, gate_r, rmax, cx - 1, cx + 1, cy - 1, cy + 1, cz - 1, cz + 1, t, d4f);
    store_top5(a, p, t);
  }
}

void knn_box2_run(const SearchArgs& sa, DevBuf<int>& work, DevBuf<int>& scan_tmp, cudaStream_t st, LaunchCounter& lc) {
  const unsigned nb = (unsigned)((sa.Qt + 127) / 128);
  const int nw = (int)nb * 4;
  work.reserve((size_t)3 * nw + 8);
  TileArgs ta;
  ta.sa = sa; ta.pairs = nullptr; ta.nwarps = nw;
  ta.dmask = reinterpret_cast<unsigned int*>(work.p); ta.dcount = work.p + nw; ta.dpos = work.p + 2 * nw + 1;
  ta.n_fb = nullptr; ta.fb_list = nullptr;
  GLIO_CUDA_TRY(cudaMemsetAsync(ta.dcount + nw, 0, sizeof(int), st));
  lc.begin("k_knn_box_start", st); k_knn_box_start<<<nb, 128, 0, st>>>(ta); lc.end(st);
  exclusive_scan_i32(ta.dcount, work.p + 2 * nw + 1, (int64_t)nw + 1, scan_tmp, st, lc);
  lc.begin("k_defer_scatter", st); k_defer_scatter<<<(unsigned)((nw + 255) / 256), 256, 0, st>>>(ta); lc.end(st);
  lc.begin("k_knn_grow", st); k_knn_grow<<<148 * 12, 128, 0, st>>>(ta); lc.end(st);
  GLIO_CUDA_TRY(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------------
// GLIO_KNN_MODE=6: k_knn_box, except that a query with fewer than five map points in its 3x3x3 start box (it lies far from
// every surface: its result needs the whole gate sphere, about 700 cells scanned one row after the other by one thread) is
// handed to the team pass, where several lanes share that scan.
// ---------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128, 6) k_knn_box_far(SearchArgs a) {
  __shared__ int sbnd[18 * 128];
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const GridDesc& g = a.grid;
  bool far = false;
  if (p < a.Qt) {
    const float4 q4 = a.pmq[p];
    const float qx = q4.x, qy = q4.y, qz = q4.z;
    Top5 t; top5_init(t);
    const int cx = cell_coord(qx, g.ox, g.inv_cell), cy = cell_coord(qy, g.oy, g.inv_cell), cz = cell_coord(qz, g.oz, g.inv_cell);
    const float gate_r = sqrtf(a.gate_sq) + 2e-3f;
    const int rmax = (int)ceilf(gate_r * g.inv_cell) + 1;
    const bool far_out = cx < -rmax || cy < -rmax || cz < -rmax || cx >= g.nx + rmax || cy >= g.ny + rmax || cz >= g.nz + rmax;
    float d4f = __int_as_float(0x7f800000);
    if (!far_out) {
      thread_rings(g, qx, qy, qz, cx, cy, cz, 1, 1, rmax, t, d4f, sbnd + threadIdx.x);
      far = key_idx(t.k4) == 0x7fffffff;
      if (far) {
        bool open = false;
#pragma unroll 1
        for (int f = 0; f < 6 && !open; ++f) open = box_face_open(g, f, qx, qy, qz, gate_r, cx - 1, cx + 1, cy - 1, cy + 1, cz - 1, cz + 1, t);
        far = open;                                   // a box already clipped by the grid on every side is final as it is
      }
      if (!far) box_grow(g, qx, qy, qz, gate_r, rmax, cx - 1, cx + 1, cy - 1, cy + 1, cz - 1, cz + 1, t, d4f);
    }
    store_top5(a, p, t);                              // far: knn_sqd[4] = +inf -> the team pass searches the gate sphere
  }
  const unsigned fm = __ballot_sync(0xffffffffu, far);
  if (fm) {
    unsigned int base = 0;
    if (lane == 0) { base = atomicAdd(a.n_deferred, (unsigned int)__popc(fm)); if (a.n_fallback) atomicAdd(a.n_fallback, (unsigned long long)__popc(fm)); }
    base = __shfl_sync(0xffffffffu, base, 0);
    if (far) a.deferred[base + __popc(fm & ((1u << lane) - 1u))] = (uint32_t)p;
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// k_knn_far: one WARP per far query (fewer than five map points in its start box, so the whole gate sphere has to be
// searched).  Lanes own the cell rows of the sphere's bounding box for the bounds (two loads per row, 32 rows at a time,
// each row clipped to the circle the sphere cuts out of it), then the candidates of those rows are FLATTENED over the lanes
// (warp scan of the row lengths, per-lane binary search by shuffle) so every lane evaluates one candidate per step whatever
// the rows look like.  Pass A: per-lane minimum of the exact distances; the 5th smallest of the 32 lane minima bounds the 5th
// neighbour distance (five distinct candidates).  Pass B: candidates under the bound are collected (ballot + prefix) into a
// small shared list; the five smallest (distance, index) keys of the list are extracted with warp reductions.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int FAR_WARPS = 4;
constexpr int FAR_ROWS = 512;        // rows of the bounding box kept per warp ((2R+1)^2 with R = ceil(gate radius / cell) <= 11)
constexpr int FAR_LIST = 128;        // survivor list capacity

__device__ __forceinline__ unsigned long long warp_min_key(unsigned long long k) {
  const unsigned hi = __reduce_min_sync(0xffffffffu, (unsigned)(k >> 32));
  const unsigned lo = __reduce_min_sync(0xffffffffu, (unsigned)(k >> 32) == hi ? (unsigned)k : 0xffffffffu);
  return ((unsigned long long)hi << 32) | lo;
}

__global__ void __launch_bounds__(32 * FAR_WARPS) k_knn_far(SearchArgs a, const uint32_t* __restrict__ qlist, const unsigned int* __restrict__ n_list) {
  __shared__ int s_rs[FAR_WARPS][FAR_ROWS];            // first point of the row's range
  __shared__ int s_rl[FAR_WARPS][FAR_ROWS];            // its length
  __shared__ unsigned long long s_keys[FAR_WARPS][FAR_LIST];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned int nwarps = gridDim.x * FAR_WARPS;
  const unsigned int nd = *n_list;
  const GridDesc& G = a.grid;
  const int* __restrict__ cs = G.cell_start;
  const float INF = __int_as_float(0x7f800000);
  const float T0 = a.gate_sq * TK_MARGIN;              // nothing beyond the radius gate can matter (Estimator.cpp:3651)
  int* rs = s_rs[wid]; int* rl = s_rl[wid];
  unsigned long long* keys = s_keys[wid];
  for (unsigned int item = blockIdx.x * FAR_WARPS + wid; item < nd; item += nwarps) {
    const int64_t p = qlist[item];
    const float4 q4 = a.pmq[p];
    const float qx = q4.x, qy = q4.y, qz = q4.z;
    const float r = sqrtf(T0) * 1.001f + 2e-3f;
    const int ylo = max(cell_coord(qy - r, G.oy, G.inv_cell), 0), yhi = min(cell_coord(qy + r, G.oy, G.inv_cell), G.ny - 1);
    const int zlo = max(cell_coord(qz - r, G.oz, G.inv_cell), 0), zhi = min(cell_coord(qz + r, G.oz, G.inv_cell), G.nz - 1);
    const int nyr = yhi - ylo + 1, nzr = zhi - zlo + 1;
    const int nrows_all = (nyr > 0 && nzr > 0) ? nyr * nzr : 0;
    float bound = T0;
    unsigned long long best[5] = {KEY_EMPTY, KEY_EMPTY, KEY_EMPTY, KEY_EMPTY, KEY_EMPTY};
    int nlist = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {
      float lane_min = INF;
#pragma unroll 1
      for (int row0 = 0; row0 < nrows_all; row0 += FAR_ROWS) {          // (one window unless the cells are tiny)
        const int nrows = min(nrows_all - row0, FAR_ROWS);
        // ---- bounds of every row of the window, clipped to the circle the sphere cuts out of the row
        __syncwarp();
        for (int ri = lane; ri < nrows; ri += 32) {
          const int z = zlo + (row0 + ri) / nyr, y = ylo + (row0 + ri) % nyr;
          const float yl = G.oy + (float)y * G.cell, zl = G.oz + (float)z * G.cell;
          const float dy = fmaxf(fmaxf(yl - qy, qy - (yl + G.cell)), 0.f), dz = fmaxf(fmaxf(zl - qz, qz - (zl + G.cell)), 0.f);
          const float dys = fmaxf(dy * 0.999f - 2e-3f, 0.f), dzs = fmaxf(dz * 0.999f - 2e-3f, 0.f);
          const float rem = T0 - (dys * dys + dzs * dzs);
          int s = 0, len = 0;
          if (rem >= 0.f) {
            const float rx = sqrtf(rem) * 1.001f + 2e-3f;
            const int xl = max(cell_coord(qx - rx, G.ox, G.inv_cell), 0), xh = min(cell_coord(qx + rx, G.ox, G.inv_cell), G.nx - 1);
            if (xl <= xh) { const int row = (z * G.ny + y) * G.nx; s = __ldg(&cs[row + xl]); len = __ldg(&cs[row + xh + 1]) - s; }
          }
          rs[ri] = s; rl[ri] = len;
        }
        __syncwarp();
        for (int rb = 0; rb < nrows; rb += 32) {
          const int ri = rb + lane;
          const int len = ri < nrows ? rl[ri] : 0, s = ri < nrows ? rs[ri] : 0;
          int incl = len;
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
          const int total = __shfl_sync(0xffffffffu, incl, 31);
          for (int base = 0; base < total; base += 32) {
            const int c = base + lane;
            // row of flat candidate c: the first lane whose inclusive count exceeds c (binary search over the lanes by shuffle)
            int lo = 0;
#pragma unroll
            for (int step = 16; step > 0; step >>= 1) { const int v = __shfl_sync(0xffffffffu, incl, lo + step - 1); if (v <= c) lo += step; }
            const int src_incl = __shfl_sync(0xffffffffu, incl, lo), src_len = __shfl_sync(0xffffffffu, len, lo), src_s = __shfl_sync(0xffffffffu, s, lo);
            float d = INF; int id = 0x7fffffff;
            if (c < total) {
              const float4 pt = __ldg(&G.pts[src_s + (c - (src_incl - src_len))]);
              d = l2_simple(qx, qy, qz, pt.x, pt.y, pt.z); id = __float_as_int(pt.w);
            }
            if (pass == 0) lane_min = fminf(lane_min, d);
            else {
              const bool keep = d <= bound;
              const unsigned bm = __ballot_sync(0xffffffffu, keep);
              if (keep) { const int slot = nlist + __popc(bm & ((1u << lane) - 1u)); if (slot < FAR_LIST) keys[slot] = make_key(d, id); else top5_push(*reinterpret_cast<Top5*>(best), d, id); }
              nlist = min(nlist + __popc(bm), FAR_LIST + 64);
            }
          }
        }
      }
      if (pass == 0) {
        // 5th smallest of the 32 lane minima (each is a distinct candidate); fewer than five finite minima leave the gate as the bound
        unsigned mine = __float_as_uint(lane_min), fifth = 0x7f800000u;
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const unsigned m = __reduce_min_sync(0xffffffffu, mine);
          fifth = m;
          const unsigned who = __ballot_sync(0xffffffffu, mine == m);
          if (lane == __ffs(who) - 1) mine = 0x7f800000u;
        }
        bound = fminf(__uint_as_float(fifth), T0);
      }
    }
    __syncwarp();
    // ---- the five smallest keys of the list (+ anything that overflowed into lane-private `best`, folded in by every lane)
    Top5 t; top5_init(t);
    const int nl = min(nlist, FAR_LIST);
    unsigned long long mykeys[FAR_LIST / 32];
#pragma unroll
    for (int u = 0; u < FAR_LIST / 32; ++u) { const int i = u * 32 + lane; mykeys[u] = i < nl ? keys[i] : KEY_EMPTY; }
    // overflow candidates live in one lane's `best`: bring them into the pool as well
    unsigned long long ov[5] = {best[0], best[1], best[2], best[3], best[4]};
    unsigned long long res[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) {
      unsigned long long m = KEY_EMPTY;
#pragma unroll
      for (int u = 0; u < FAR_LIST / 32; ++u) m = m < mykeys[u] ? m : mykeys[u];
#pragma unroll
      for (int u = 0; u < 5; ++u) m = m < ov[u] ? m : ov[u];
      const unsigned long long g = warp_min_key(m);
      res[k] = g;
      // remove one copy of it (keys are distinct points, so at most one lane slot holds it)
#pragma unroll
      for (int u = 0; u < FAR_LIST / 32; ++u) if (mykeys[u] == g) mykeys[u] = KEY_EMPTY;
#pragma unroll
      for (int u = 0; u < 5; ++u) if (ov[u] == g) ov[u] = KEY_EMPTY;
    }
    t.k0 = res[0]; t.k1 = res[1]; t.k2 = res[2]; t.k3 = res[3]; t.k4 = res[4];
    if (lane == 0) store_top5(a, p, t);
    __syncwarp();
  }
}

void knn_box_far_run(const SearchArgs& sa, cudaStream_t st, LaunchCounter& lc) {
  lc.begin("k_knn_box_far", st); k_knn_box_far<<<(unsigned)((sa.Qt + 127) / 128), 128, 0, st>>>(sa); lc.end(st);
  lc.begin("k_knn_far", st); k_knn_far<<<148 * 8, 32 * FAR_WARPS, 0, st>>>(sa, sa.deferred, sa.n_deferred); lc.end(st);
  GLIO_CUDA_TRY(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------------------
// GLIO_KNN_MODE=7: the default box search with a CELL-BY-CELL start box.  At cfg-2 density the 5th neighbour sits 0.13 - 0.2 m
// away while a cell is ~0.3 m wide, so a query's search sphere touches a handful of the 27 cells of its start box.  The four
// cell boundaries of each of the nine rows are fetched up front (36 independent loads, parked in the thread's shared-memory
// column); the own cell is scanned first, then the 6 face, 12 edge and 8 corner cells, each only if it can still hold something
// nearer than the current 5th distance (per-axis gaps to the own cell's walls with the face-test margins, strict inequality: a
// skipped cell holds only points strictly farther than the final 5th distance, so the box counts as scanned for the growth
// proof).  Bit-exact - and 3.4x SLOWER than the nine-row start box (2.1 ms, profiles/r02_sweep_start_box.txt): a warp's
// queries share a cell but not a pruning pattern, so the warp executes the union of the lanes' cells - all 27, with 27 loop
// prologues instead of 9 - and the 27 inlined scan loops are 8 600 SASS instructions (137 KB, beyond the instruction cache).
// Kept as the checkable form of that negative result.
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float wall_gap2(float d) {
  const float m = d * 0.999f - 2e-3f;
  return m > 0.f ? m * m : 0.f;
}
#define GLIO_CELL(I, K, G2) \
  if (!(d4f <= (G2))) scan_range_keys(g.pts, sb[((I) * 4 + (K)) * 128], sb[((I) * 4 + (K) + 1) * 128], qx, qy, qz, t, d4f);
__device__ __forceinline__ void start_box_cells(const GridDesc& g, float qx, float qy, float qz, int cx, int cy, int cz, Top5& t, float& d4f, int* __restrict__ sb) {
  const int* __restrict__ cs = g.cell_start;
  const int xk0 = min(max(cx - 1, 0), g.nx), xk1 = min(max(cx, 0), g.nx), xk2 = min(max(cx + 1, 0), g.nx), xk3 = min(max(cx + 2, 0), g.nx);
#pragma unroll
  for (int dz = 0; dz < 3; ++dz) {
    int b[12];
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
      const int z = cz + dz - 1, y = cy + dy - 1;
      const bool ok = z >= 0 && z < g.nz && y >= 0 && y < g.ny;
      const int row = (z * g.ny + y) * g.nx;
      b[4 * dy + 0] = ok ? __ldg(&cs[row + xk0]) : 0; b[4 * dy + 1] = ok ? __ldg(&cs[row + xk1]) : 0;
      b[4 * dy + 2] = ok ? __ldg(&cs[row + xk2]) : 0; b[4 * dy + 3] = ok ? __ldg(&cs[row + xk3]) : 0;
    }
#pragma unroll
    for (int j = 0; j < 12; ++j) sb[(dz * 12 + j) * 128] = b[j];
  }
  const float gxm = wall_gap2(qx - (g.ox + (float)cx * g.cell)), gxp = wall_gap2((g.ox + (float)(cx + 1) * g.cell) - qx);
  const float gym = wall_gap2(qy - (g.oy + (float)cy * g.cell)), gyp = wall_gap2((g.oy + (float)(cy + 1) * g.cell) - qy);
  const float gzm = wall_gap2(qz - (g.oz + (float)cz * g.cell)), gzp = wall_gap2((g.oz + (float)(cz + 1) * g.cell) - qz);
  // row I = (dz+1)*3 + (dy+1), cell K = dx+1
  GLIO_CELL(4, 1, -1.f)
  GLIO_CELL(4, 0, gxm) GLIO_CELL(4, 2, gxp) GLIO_CELL(3, 1, gym) GLIO_CELL(5, 1, gyp) GLIO_CELL(1, 1, gzm) GLIO_CELL(7, 1, gzp)
  GLIO_CELL(3, 0, gym + gxm) GLIO_CELL(3, 2, gym + gxp) GLIO_CELL(5, 0, gyp + gxm) GLIO_CELL(5, 2, gyp + gxp)
  GLIO_CELL(1, 0, gzm + gxm) GLIO_CELL(1, 2, gzm + gxp) GLIO_CELL(7, 0, gzp + gxm) GLIO_CELL(7, 2, gzp + gxp)
  GLIO_CELL(0, 1, gzm + gym) GLIO_CELL(2, 1, gzm + gyp) GLIO_CELL(6, 1, gzp + gym) GLIO_CELL(8, 1, gzp + gyp)
  GLIO_CELL(0, 0, gzm + gym + gxm) GLIO_CELL(0, 2, gzm + gym + gxp) GLIO_CELL(2, 0, gzm + gyp + gxm) GLIO_CELL(2, 2, gzm + gyp + gxp)
  GLIO_CELL(6, 0, gzp + gym + gxm) GLIO_CELL(6, 2, gzp + gym + gxp) GLIO_CELL(8, 0, gzp + gyp + gxm) GLIO_CELL(8, 2, gzp + gyp + gxp)
}
#undef GLIO_CELL

__global__ void __launch_bounds__(128, 6) k_knn_box_cells(SearchArgs a) {
  __shared__ int sbnd[36 * 128];
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.Qt) return;
  const GridDesc& g = a.grid;
  const float4 q4 = a.pmq[p];
  const float qx = q4.x, qy = q4.y, qz = q4.z;
  Top5 t; top5_init(t);
  const int cx = cell_coord(qx, g.ox, g.inv_cell), cy = cell_coord(qy, g.oy, g.inv_cell), cz = cell_coord(qz, g.oz, g.inv_cell);
  const float gate_r = sqrtf(a.gate_sq) + 2e-3f;
  const int rmax = (int)ceilf(gate_r * g.inv_cell) + 1;
  const bool far_out = cx < -rmax || cy < -rmax || cz < -rmax || cx >= g.nx + rmax || cy >= g.ny + rmax || cz >= g.nz + rmax;
  float d4f = __int_as_float(0x7f800000);
  if (!far_out) {
    start_box_cells(g, qx, qy, qz, cx, cy, cz, t, d4f, sbnd + threadIdx.x);
    box_grow(g, qx, qy, qz, gate_r, rmax, cx - 1, cx + 1, cy - 1, cy + 1, cz - 1, cz + 1, t, d4f, a.grow_mode ? sbnd + threadIdx.x : nullptr);
  }
  store_top5(a, p, t);
}

void knn_box_cells_run(const SearchArgs& sa, cudaStream_t st, LaunchCounter& lc) {
  lc.begin("k_knn_box_cells", st); k_knn_box_cells<<<(unsigned)((sa.Qt + 127) / 128), 128, 0, st>>>(sa); lc.end(st);
  GLIO_CUDA_TRY(cudaGetLastError());
}

}  // namespace glio
