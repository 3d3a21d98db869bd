// knn_common.cuh — pieces shared by the K1a search kernels (assoc.cu, knn_tile.cu): the top-5 register list, the search
// arguments and the packed fp32x2 helpers.
#pragma once
#include "common.cuh"
#include "devmath.cuh"

namespace glio {

// ---- top-5 by (distance, index) -------------------------------------------------------------------
// One 64-bit key per neighbour: (bits of the non-negative float distance) << 32 | index.  Unsigned key order ==
// lexicographic (distance, index) order, so ties are broken by index exactly like a stable sort by (distance, index).
struct Top5 {
  unsigned long long k0, k1, k2, k3, k4;
};
__device__ __forceinline__ unsigned long long make_key(float d, int id) { return ((unsigned long long)__float_as_uint(d) << 32) | (unsigned int)id; }
__device__ __forceinline__ float key_dist(unsigned long long k) { return __uint_as_float((unsigned int)(k >> 32)); }
__device__ __forceinline__ int key_idx(unsigned long long k) { return (int)(unsigned int)(k & 0xffffffffull); }
constexpr unsigned long long KEY_EMPTY = 0x7f8000007fffffffull;   // (+inf, INT_MAX)
__device__ __forceinline__ void top5_init(Top5& t) { t.k0 = t.k1 = t.k2 = t.k3 = t.k4 = KEY_EMPTY; }
// compare-exchange of two keys: one 64-bit compare + select for the smaller key, the larger one by XOR (the plain
// two-select form made ptxas emit a second, mirrored compare: 8 instead of 6 instructions per exchange in the hottest
// block of K1a)
__device__ __forceinline__ void key_cswap(unsigned long long& a, unsigned long long& b) {
  const unsigned long long lo = a < b ? a : b;
  b = a ^ b ^ lo;
  a = lo;
}
#define GLIO_KSWAP(A, B) key_cswap((A), (B));
__device__ __forceinline__ void top5_push(Top5& t, float d, int id) {
  const unsigned long long k = make_key(d, id);
  if (k < t.k4) {
    t.k4 = k;
    GLIO_KSWAP(t.k3, t.k4)
    GLIO_KSWAP(t.k2, t.k3)
    GLIO_KSWAP(t.k1, t.k2)
    GLIO_KSWAP(t.k0, t.k1)
  }
}

// ---- packed fp32x2 arithmetic (Blackwell FADD2/FMUL2/FFMA2) for candidate PRE-FILTERS only.
// ptxas contracts packed mul+add into FFMA2 even for .rn operands and under -fmad=false, so packed results may
// differ from the reference's unfused ((dx*dx)+dy*dy)+dz*dz by a few ulps.  They are therefore used only to reject
// candidates that are clearly farther than a threshold (with a 1e-6 relative margin); every candidate that might
// matter is re-evaluated with the exact scalar l2_simple() before it can enter a top-5.
typedef unsigned long long f32x2_t;
__device__ __forceinline__ f32x2_t pack2(float a, float b) { f32x2_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack2(f32x2_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2_t sub2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t mul2(f32x2_t a, f32x2_t b) { f32x2_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2_t fma2(f32x2_t a, f32x2_t b, f32x2_t c) { f32x2_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }

struct SearchArgs {
  GridDesc grid;
  int64_t Qt;
  const float4* pm;
  const uint32_t* order;
  float gate_sq;
  int32_t* knn_idx;     // [5][Qt] sorted order
  float* knn_sqd;       // [5][Qt] sorted order; rows 0..3 are written only when store_all_sqd (the gate needs the 5th only)
  bool store_all_sqd;
  unsigned long long* n_fallback;   // statistics: queries deferred to the second pass
  int tile_rings;                   // (mode 0) rings the tile pass may scan before it defers a query
  uint32_t* deferred;               // sorted positions of deferred queries
  unsigned int* n_deferred;
};

__device__ __forceinline__ void store_top5(const SearchArgs& a, int64_t p, const Top5& t) {
  a.knn_idx[0 * a.Qt + p] = key_idx(t.k0); a.knn_idx[1 * a.Qt + p] = key_idx(t.k1); a.knn_idx[2 * a.Qt + p] = key_idx(t.k2);
  a.knn_idx[3 * a.Qt + p] = key_idx(t.k3); a.knn_idx[4 * a.Qt + p] = key_idx(t.k4);
  a.knn_sqd[4 * a.Qt + p] = key_dist(t.k4);
  if (a.store_all_sqd) {
    a.knn_sqd[0 * a.Qt + p] = key_dist(t.k0); a.knn_sqd[1 * a.Qt + p] = key_dist(t.k1); a.knn_sqd[2 * a.Qt + p] = key_dist(t.k2);
    a.knn_sqd[3 * a.Qt + p] = key_dist(t.k3);
  }
}

// K1a, staged tile search + team pass (knn_tile.cu, GLIO_KNN_MODE=4)
void knn_tile_run(const SearchArgs& sa, const PairRec* d_pairs, DevBuf<int>& work, DevBuf<int>& scan_tmp, cudaStream_t st, LaunchCounter& lc);

}  // namespace glio
