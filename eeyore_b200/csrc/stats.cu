// On-device chain diagnostics: mean, sample covariance, initial-sequence Monte Carlo covariance (INSE), multivariate
// ESS and autocorrelation, batched over chains.
//
// Replaces (reference paths):
//   eeyore/stats/cov.py:5-15           sample covariance (n-1 denominator)
//   eeyore/linalg/is_pos_def.py:3-11   "exactly symmetric and Cholesky succeeds"
//   eeyore/stats/inse_mc_cov.py:9-83   INSE estimator (adjust=False): the python double loop of torch.ger outer products
//   eeyore/stats/multi_ess.py:6-14     n (det cov / det inse)^(1/p)
// ACF is builder-defined (kanga is absent from the reference snapshot).
//
// Mapping (round 2; the first version gave a whole CTA to one chain and spent its time in CTA barriers and in serial
// factorisations by one warp while seven waited):
//   * ONE WARP PER CHAIN, no CTA-wide synchronisation anywhere; chains are independent, so a warp walks through its own
//     data-dependent INSE loop (inse_mc_cov.py:20-73) at its own pace.
//   * The chain is STREAMED, never staged whole: every pass (column means; autocorrelation; lags 0 and 1; one pass per
//     further lag pair) pulls the rows through a 64-row shared-memory ring, 16-row chunks fetched by TMA 1-D bulk copies
//     (cp.async.bulk + mbarrier, issued by one lane three chunks ahead) when a chain's rows are contiguous -- the
//     [C, n, P] layout the samplers write with sample_layout = "cnp" -- and by plain loads for any other strides.  Rows
//     are centred once, on arrival; rows past the end read as zero, so lagged sums need no tail handling.  Any n works.
//   * Lagged cross-products sum_i x_i (x) x_{i+l}: the warp is two 16-lane row slices (8 consecutive rows of a chunk each),
//     each lane owns a TE x TE register tile of the P x P product (P = 20: 5 x 5, all 32 lanes busy).  Consecutive rows
//     share their partner rows, so the B-side values slide through registers: lags 0 and 1 together cost 2 TE loads per
//     row for 2 TE^2 FMAs.  INSE only needs gamma_{2m} + gamma_{2m+1}, so for m >= 1 ONE pass multiplies x_i with the pair
//     sum x_{i+2m} + x_{i+2m+1} (formed from the sliding window): half the FMAs of two separate lags.
//   * Cholesky (the is_pos_def test) and LU with partial pivoting (torch.det) work on a shared-memory copy, one matrix row
//     per lane, the pivot row read as a broadcast; compact rolled loops (a fully unrolled register version was tried first:
//     its straight-line code made instruction-cache misses the kernel's top stall).
// Passes over a chain when everything is asked for: lags 0 + 1 (on rows shifted by a sample of the chain, with the means as a
// by-product), autocorrelation, one per further lag pair.  HBM traffic: each pass reads the chain once (n P sizeof(T) bytes
// per chain).  FLOPs: 2 n P^2 per lag, n P^2 per lag pair.
#include <cuda_runtime.h>
#include <string>
#include "common.cuh"
#include "../../include/eeyore_b200.h"

namespace eb {

constexpr int kStatWarps = 4;          // warps (= chains in flight) per CTA
constexpr int kRB = 16;                // rows per chunk; a half-warp (row slice) works on 8 consecutive rows of it
constexpr int kHB = kRB / 2;
constexpr int kNBuf = 4;               // chunks in the ring
constexpr int kW = kRB * kNBuf;        // ring rows (64)
constexpr int kNearLag = (kNBuf - 2) * kRB - 1;   // largest lag whose partner rows are still in the ring (31)
constexpr int kAcfGroup = 16;          // autocorrelation lags handled per pass
constexpr unsigned kFull = 0xffffffffu;

template <typename T> struct StatsArgs {
  const T* x;
  long s_iter, s_chain, s_param;
  int n, P;
  long C;
  T* out_mean;   // [C,P]
  T* out_cov;    // [C,P,P]
  T* out_inse;   // [C,P,P]
  T* out_ess;    // [C]
  int* out_status;  // [C] 0 ok, 1 = not enough samples (inse_mc_cov.py:44-45)
  int* out_lags;    // [C,2] (sn, last accepted m)
  int max_lag;
  T* out_acf;    // [C, max_lag+1, P]
  const long* index;   // optional [C]: the chains to process (inputs and outputs are addressed by index[k]); NULL = 0 .. C-1
  int defer_m;         // >= 0: a chain whose estimate is not positive definite up to lag pair defer_m is left with status 3
};

// ---- mbarrier / bulk copy (one barrier per ring slot, owned by a warp) ---------------------------------------------------
__device__ __forceinline__ uint32_t st_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void st_mbar_init(uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(st_smem_u32(bar)));
}
__device__ __forceinline__ void st_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(st_smem_u32(bar)), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(st_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(st_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void st_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_%=;\n\t}" ::"r"(st_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void st_fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- the ring: rows of ONE chain streamed through shared memory by the warp that owns it ---------------------------------
// PC = the number of parameters when it is a multiple of four (compile-time strides: every shared-memory offset of the hot
// loops is an immediate), 0 = run-time P.
template <typename T, int PC> struct Ring {
  const T* base;        // first element of the chain
  long s_iter, s_param;
  int n, p_rt;
  bool bulk;            // rows contiguous and 16-byte aligned: chunks come by cp.async.bulk
  T* X;                 // [kW][P] centred rows
  const T* mean;        // [32] subtracted on arrival
  uint64_t* bar;        // [kNBuf]
  uint32_t* phase;      // parity bit per slot (shared by every view of this warp's ring)
  // A view streams rows shift, shift + 1, ... in chunks of kRB rows through `nslots` (a power of two) slots from slot0 on:
  // the whole ring (shift 0, 4 slots) for lags whose partner rows stay within the window, or two half rings for far lags --
  // the rows themselves through slots 0-1 and their partners (shift = lag) through slots 2-3.
  int shift, slot0, nslots;
  int next_issue, next_ready, last_chunk;

  __device__ __forceinline__ int P() const { return PC ? PC : p_rt; }
  __device__ __forceinline__ void view(int shift_, int slot0_, int nslots_) { shift = shift_; slot0 = slot0_; nslots = nslots_; }
  __device__ __forceinline__ int slot_of(int k) const { return slot0 + (k & (nslots - 1)); }
  __device__ __forceinline__ int rows_of(int k) const {
    const int r = n - shift - k * kRB;
    return r < 0 ? 0 : (r > kRB ? kRB : r);
  }
  __device__ __forceinline__ const T* src_of(int k) const { return base + ((long)k * kRB + shift) * s_iter; }
  __device__ __forceinline__ bool by_bulk(int k) const {
    const int rows = rows_of(k);
    return bulk && rows > 0 && ((rows * P() * (int)sizeof(T)) & 15) == 0 && (reinterpret_cast<uintptr_t>(src_of(k)) & 15) == 0;
  }
  // row g of the chain (g >= shift, within the window)
  __device__ __forceinline__ const T* row(int g) const {
    const int q = g - shift;
    return X + ((slot0 + ((q / kRB) & (nslots - 1))) * kRB + (q & (kRB - 1))) * P();
  }
  // all lanes; the slot's previous contents are dead (the callers' fence + __syncwarp order the last accesses before this)
  __device__ __forceinline__ void issue(int k, int lane) {
    const int rows = rows_of(k);
    if (rows <= 0) return;
    T* dst = X + slot_of(k) * kRB * P();
    const T* src = src_of(k);
    if (by_bulk(k)) {
      if (lane == 0) st_bulk_load(dst, src, (uint32_t)(rows * P() * sizeof(T)), bar + slot_of(k));
    } else {
      for (int e = lane; e < rows * P(); e += 32) {
        const int i = e / P(), j = e - i * P();
        dst[e] = src[i * s_iter + j * s_param];
      }
    }
  }
  __device__ __forceinline__ void begin(int last, int lane) {
    next_issue = 0; next_ready = 0; last_chunk = last;
    st_fence_proxy_async();   // this lane's generic-proxy accesses to the ring memory, before the async-proxy (TMA) writes
    __syncwarp();
    for (; next_issue < nslots && next_issue <= last_chunk; ++next_issue) issue(next_issue, lane);
  }
  // chunks up to k arrived, centred, zero-filled past the end
  __device__ __forceinline__ void ready(int k, int lane) {
    while (next_ready <= k) {
      const int kk = next_ready++, slot = slot_of(kk), rows = rows_of(kk);
      if (by_bulk(kk)) {
        st_mbar_wait(bar + slot, (*phase >> slot) & 1u);
        *phase ^= 1u << slot;
      }
      __syncwarp();
      T* dst = X + slot * kRB * P();
      if (PC > 0 && rows == kRB) {          // full chunk, compile-time P: kRB P / 32 = P / 2 elements per lane
        constexpr int NE = PC > 0 ? PC / 2 : 1;
        T v[NE], mv[NE];
#pragma unroll
        for (int q = 0; q < NE; ++q) { v[q] = dst[lane + 32 * q]; mv[q] = mean[(lane + 32 * q) % (PC > 0 ? PC : 1)]; }
#pragma unroll
        for (int q = 0; q < NE; ++q) dst[lane + 32 * q] = v[q] - mv[q];
      } else {
        for (int e = lane; e < kRB * P(); e += 32) {
          const int i = e / P(), j = e - i * P();
          dst[e] = (i < rows) ? dst[e] - mean[j] : T(0);
        }
      }
      __syncwarp();
    }
  }
  // chunk `done` has been consumed: its slot takes the next chunk
  __device__ __forceinline__ void release(int done, int lane) {
    st_fence_proxy_async();
    __syncwarp();
    if (next_issue <= last_chunk && next_issue == done + nslots) { issue(next_issue, lane); ++next_issue; }
  }
  // centred value of row g, column j straight from global memory (autocorrelation lags beyond the window)
  __device__ __forceinline__ T far(int g, int j) const {
    return (g < n) ? base[(long)g * s_iter + (long)j * s_param] - mean[j] : T(0);
  }
};

template <typename T> __device__ __forceinline__ T shfl_t(T v, int src) { return __shfl_sync(kFull, v, src); }
template <typename T> __device__ __forceinline__ T shfl_xor_t(T v, int m) { return __shfl_xor_sync(kFull, v, m); }

// ---- factorisations of a P x P matrix in shared memory (row-major, row stride ld odd), one row per lane -------------------
// Compact rolled loops in real functions: straight-line register versions (one copy per call site, ~80 KB of code each for
// P = 20) made instruction-cache misses the top stall of the kernel.  Lane r updates row r; the pivot row is read by
// every lane at the same address (a broadcast), so nothing goes through shuffles but the pivot search.
// torch.linalg.cholesky succeeds <=> every pivot is positive (is_pos_def.py:5-9).  Right-looking; w is destroyed.
template <typename T> __device__ __noinline__ bool warp_chol_ok(T* w, int P, int ld, int lane) {
  T* my = w + (lane < P ? lane : 0) * ld;
  for (int k = 0; k < P; ++k) {
    const T d = w[k * ld + k];
    if (!(d > T(0))) return false;              // uniform: every lane reads the same element
    const T rinv = T(1) / sqrt_t<T>(d);
    const bool below = lane > k && lane < P;
    const T lk = below ? my[k] * rinv : T(0);   // L[r][k]
    if (below) my[k] = lk;
    __syncwarp();
    if (below)
      for (int c = k + 1; c <= lane; ++c) my[c] = fma_t<T>(-lk, w[c * ld + k], my[c]);   // lower triangle only
    __syncwarp();
  }
  return true;
}

// determinant by LU with partial pivoting (torch.det; inse_mc_cov.py:47,66, multi_ess.py:9-12).  Rows stay where they are:
// `pos` is the logical position of a lane's row, a pivot swaps positions only.  w is destroyed.
template <typename T> __device__ __noinline__ T warp_det_lu(T* w, int P, int ld, int lane) {
  bool done = lane >= P;
  int pos = lane;
  T det = T(1);
  T* my = w + (lane < P ? lane : 0) * ld;
  for (int k = 0; k < P; ++k) {
    const T ak = done ? T(0) : my[k];
    // NaN sorts like +inf: it becomes the pivot and ends the loop uniformly (a NaN key would break the total order and the
    // lanes would disagree about the pivot lane)
    T v = done ? T(-1) : ((ak != ak) ? T(INFINITY) : fabs(ak));
    int vp = pos, vl = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {   // largest |a[k]|; ties -> lowest logical position (LAPACK idamax order)
      const T v2 = shfl_xor_t(v, o);
      const int p2 = __shfl_xor_sync(kFull, vp, o), l2 = __shfl_xor_sync(kFull, vl, o);
      const bool take = (v2 > v) || (v2 == v && p2 < vp);
      if (take) { v = v2; vp = p2; vl = l2; }
    }
    const T* pr = w + vl * ld;
    const T pv = pr[k];
    if (vp != k) {                        // row interchange: the row at logical position k takes the pivot row's position
      det = -det;
      if (!done && pos == k) pos = vp;
    }
    if (lane == vl) pos = k;
    det *= pv;
    if (pv == T(0) || pv != pv) break;    // uniform
    if (!done && lane != vl) {
      const T f = ak * (T(1) / pv);
      for (int c = k + 1; c < P; ++c) my[c] = fma_t<T>(-f, pr[c], my[c]);
    }
    if (lane == vl) done = true;
    __syncwarp();
  }
  return det;
}

// ---- lagged cross-products ---------------------------------------------------------------------------------------------------
// lane = (h, ta, tb): row slice h (8 consecutive rows of every 16-row chunk), tile (ta, tb) of TE x TE entries with STRIDED
// ownership: the lane holds entries (ta + 4 r, tb + 4 c) -- for a fixed r the four ta lanes read four consecutive values.
template <typename T, int TE> struct Tile {
  T v[TE][TE];
  __device__ __forceinline__ void zero() {
#pragma unroll
    for (int r = 0; r < TE; ++r)
#pragma unroll
      for (int c = 0; c < TE; ++c) v[r][c] = T(0);
  }
  __device__ __forceinline__ void rank1(const T (&a)[TE], const T (&b)[TE]) {
#pragma unroll
    for (int r = 0; r < TE; ++r)
#pragma unroll
      for (int c = 0; c < TE; ++c) v[r][c] = fma_t<T>(a[r], b[c], v[r][c]);
  }
  // sum of the two row slices, then the h == 0 lanes store the tile into a row-major [4 TE][ld] matrix
  __device__ __forceinline__ void store(T* mat, int ld, int ta, int tb, int h) {
#pragma unroll
    for (int r = 0; r < TE; ++r)
#pragma unroll
      for (int c = 0; c < TE; ++c) {
        const T s = v[r][c] + shfl_xor_t(v[r][c], 16);
        if (h == 0) mat[(ta + 4 * r) * ld + tb + 4 * c] = s;
      }
  }
};

// out[r] = row[t0 + 4 r] (zero past the last column when P is not a multiple of four)
template <typename T, int TE, int PC> __device__ __forceinline__ void load_cols(const T* row, int t0, int P, T (&out)[TE]) {
#pragma unroll
  for (int r = 0; r < TE; ++r) out[r] = (PC > 0 || t0 + 4 * r < P) ? row[t0 + 4 * r] : T(0);
}

// column means: mean[j] = (1/n) sum_i x[i][j]   (the ring's own mean must be zero during this pass).
// Returns true when some column never changes (a chain that never moved, or a frozen parameter): its mean is then set to
// that value exactly, the centred column is exactly zero, so is its row / column of every lagged product, and no Sigma_m can
// be positive definite -- the caller reports 'Not enough samples' (inse_mc_cov.py:44-45) without walking through n / 2 lags.
template <typename T, int PC> __device__ bool pass_mean(Ring<T, PC>& rg, T* mean_out, int lane) {
  const int n = rg.n, P = rg.P(), nci = (n + kRB - 1) / kRB;
  rg.begin(nci - 1, lane);
  T s0 = T(0), s1 = T(0), first = T(0);
  bool same = true;
  for (int ci = 0; ci < nci; ++ci) {
    rg.ready(ci, lane);
    if (lane < P) {
      const T* x0 = rg.X + (ci & (kNBuf - 1)) * kRB * P + lane;
      if (ci == 0) first = x0[0];
      const int rows = rg.rows_of(ci);
#pragma unroll
      for (int t = 0; t < kRB; t += 2) {
        const T v0 = x0[t * P], v1 = x0[(t + 1) * P];
        s0 += v0; s1 += v1;
        same = same && (t >= rows || v0 == first) && (t + 1 >= rows || v1 == first);
      }
    }
    rg.release(ci, lane);
  }
  __syncwarp();
  const bool frozen = lane < P && same;
  mean_out[lane] = (lane < P) ? (frozen ? first : (s0 + s1) * (T(1) / T(n))) : T(0);
  __syncwarp();
  return __any_sync(kFull, frozen);
}

// A0 = sum_i x_i (x) x_i and A1 = sum_i x_i (x) x_{i+1} in one pass (DUAL), or one of them (lag l = 0 / 1) for tiles too large
// to keep two in registers.  The B-side columns of row i + 1 are row i + 1's own B-side columns one step later: a sliding
// register window, 2 TE shared-memory loads per row for up to 2 TE^2 FMAs.  csum (lanes with tb == 0; lag-0 pass only)
// collects the column sums of the rows as they stand in the ring: the pass runs on rows shifted by a sample of the chain
// instead of centred rows, and the caller corrects the products with the mean of the shifted rows -- one pass over the
// chain less than "means first".  The row loop is unrolled by two only: with eight rows unrolled the loop bodies of the
// passes (8 KB each) evicted each other from the instruction caches (a fifth of all stall samples).
template <typename T, int TE, int PC, bool DUAL>
__device__ void pass_lag01(Ring<T, PC>& rg, int l, Tile<T, TE>& a0, Tile<T, TE>& a1, T (&csum)[TE], int lane) {
  const int n = rg.n, P = rg.P(), nci = (n + kRB - 1) / kRB;
  const int h = lane >> 4, ta = (lane >> 2) & 3, tb = lane & 3;
  const bool next_row = DUAL || l == 1;
  const bool sums = tb == 0 && (DUAL || l == 0);
  rg.begin(next_row ? nci : nci - 1, lane);     // row n (a zero row) is the partner of row n - 1
  a0.zero();
  if (DUAL) a1.zero();
#pragma unroll
  for (int r = 0; r < TE; ++r) csum[r] = T(0);
  for (int ci = 0; ci < nci; ++ci) {
    rg.ready(next_row ? ci + 1 : ci, lane);
    const int r0 = ci * kRB + kHB * h;
    const T* xr = rg.row(r0);                   // this slice's 8 rows are contiguous in the ring
    const T* xlast = rg.row(r0 + kHB);          // the row after them may sit in the next slot (or wrap)
    T pb[TE];
    load_cols<T, TE, PC>(xr, tb, P, pb);
#pragma unroll 2
    for (int t = 0; t < kHB; ++t) {
      T xa[TE], nb[TE];
      load_cols<T, TE, PC>(xr + t * P, ta, P, xa);
      if (next_row || t + 1 < kHB) load_cols<T, TE, PC>((t + 1 < kHB) ? xr + (t + 1) * P : xlast, tb, P, nb);
      if (DUAL) { a0.rank1(xa, pb); a1.rank1(xa, nb); }
      else a0.rank1(xa, l == 0 ? pb : nb);
      if (sums) {
#pragma unroll
        for (int r = 0; r < TE; ++r) csum[r] += xa[r];
      }
#pragma unroll
      for (int c = 0; c < TE; ++c) pb[c] = nb[c];
    }
    rg.release(ci, lane);
  }
}

// B = sum_i x_i (x) (x_{i+l} + x_{i+l+1}): the pair sum is formed from the sliding window (TE adds per row).  Lags beyond the
// ring window (slowly mixing chains, and chains whose estimate never becomes positive definite) stream the partner rows
// through the second half of the ring (rp: a view shifted by l) -- the same loop at the same speed.
template <typename T, int TE, int PC> __device__ void pass_pair(Ring<T, PC>& rg, int l, Tile<T, TE>& b, int lane) {
  const int n = rg.n, P = rg.P(), nci = (n + kRB - 1) / kRB;
  const int h = lane >> 4, ta = (lane >> 2) & 3, tb = lane & 3;
  const bool near = l <= kNearLag;
  Ring<T, PC> rp = rg;
  if (near) {
    rg.view(0, 0, kNBuf);
    rg.begin((nci * kRB + l) / kRB, lane);
  } else {
    rg.view(0, 0, kNBuf / 2);
    rp.view(l, kNBuf / 2, kNBuf / 2);
    rg.begin(nci - 1, lane);
    rp.begin(nci, lane);                 // partner rows of chunk ci: l + 16 ci ... l + 16 ci + 16 (the last one in chunk ci + 1)
  }
  b.zero();
  for (int ci = 0; ci < nci; ++ci) {
    if (near) {
      rg.ready((ci * kRB + kRB + l) / kRB, lane);
    } else {
      rg.ready(ci, lane);
      rp.ready(ci + 1, lane);
    }
    const Ring<T, PC>& rb = near ? rg : rp;
    const int r0 = ci * kRB + kHB * h;
    const T* xr = rg.row(r0);
    T pb[TE];
    load_cols<T, TE, PC>(rb.row(r0 + l), tb, P, pb);
#pragma unroll 2
    for (int t = 0; t < kHB; ++t) {
      T xa[TE], nb[TE], yb[TE];
      load_cols<T, TE, PC>(xr + t * P, ta, P, xa);
      load_cols<T, TE, PC>(rb.row(r0 + l + t + 1), tb, P, nb);
#pragma unroll
      for (int c = 0; c < TE; ++c) { yb[c] = pb[c] + nb[c]; pb[c] = nb[c]; }
      b.rank1(xa, yb);
    }
    rg.release(ci, lane);
    if (!near) rp.release(ci, lane);
  }
  rg.view(0, 0, kNBuf);
}

// acc[k] = sum_i x_i[j] x_{i+k0+k}[j] for lane j, k < kAcfGroup: the 16 + 16 partner values of a chunk sit in registers
// (one shared-memory load per row and lane for 16 FMAs)
template <typename T, int PC> __device__ void pass_acf(Ring<T, PC>& rg, int k0, T (&acc)[kAcfGroup], int lane) {
  const int n = rg.n, P = rg.P(), nci = (n + kRB - 1) / kRB;
  const bool near = k0 + 2 * kAcfGroup - 1 < kW - kRB;    // the partner rows of a chunk (up to i0 + k0 + 31) are still in the ring
  const int ahead = (k0 + 2 * kAcfGroup - 1) / kRB;        // chunks beyond ci that hold partner rows
  rg.begin(near ? nci - 1 + ahead : nci - 1, lane);
#pragma unroll
  for (int k = 0; k < kAcfGroup; ++k) acc[k] = T(0);
  const int j = lane < P ? lane : 0;
  static_assert(kAcfGroup == kRB, "one partner window per chunk");
  for (int ci = 0; ci < nci; ++ci) {
    rg.ready(near ? ci + ahead : ci, lane);
    const int i0 = ci * kRB;
    T w[2 * kAcfGroup];
#pragma unroll
    for (int t = 0; t < 2 * kAcfGroup; ++t) w[t] = near ? rg.row(i0 + k0 + t)[j] : rg.far(i0 + k0 + t, j);
#pragma unroll
    for (int t = 0; t < kRB; ++t) {
      const T xi = (k0 == 0) ? w[t] : rg.row(i0 + t)[j];
#pragma unroll
      for (int k = 0; k < kAcfGroup; ++k) acc[k] = fma_t<T>(xi, w[t + k], acc[k]);
    }
    rg.release(ci, lane);
  }
}

// resident CTAs per SM: three (170 registers) while one lane's tiles are small, two (255 registers) beyond
template <typename T, int TE> constexpr int stats_min_blocks() { return (TE * TE * (int)sizeof(T) <= 25 * 8) ? 3 : 2; }

template <typename T, int TE, bool EXACT>
__global__ void __launch_bounds__(kStatWarps * 32, stats_min_blocks<T, TE>()) chain_stats_kernel(const StatsArgs<T> a) {
  constexpr int PT = 4 * TE, LD = PT + 1, PC = EXACT ? PT : 0;
  constexpr bool DUAL = stats_min_blocks<T, TE>() == 3;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = a.n, P = PC ? PC : a.P;
  // per-warp carve (bytes, every piece 16-byte aligned)
  const size_t ring_bytes = ((size_t)kW * P * sizeof(T) + 15) & ~size_t(15);
  const size_t mat_bytes = ((size_t)PT * LD * sizeof(T) + 15) & ~size_t(15);
  const size_t scratch_bytes = ring_bytes > 3 * mat_bytes ? ring_bytes : 3 * mat_bytes;   // the ring, then A0 | A1 | work copy
  const size_t per_warp = scratch_bytes + mat_bytes + 64 * sizeof(T) + kNBuf * sizeof(uint64_t);
  unsigned char* my = smem_raw + (size_t)warp * per_warp;
  T* X = reinterpret_cast<T*>(my);
  T* M0 = reinterpret_cast<T*>(my);                       // product matrices alias the (dead) ring between passes
  T* M1 = reinterpret_cast<T*>(my + mat_bytes);
  T* Wk = reinterpret_cast<T*>(my + 2 * mat_bytes);
  T* Sg = reinterpret_cast<T*>(my + scratch_bytes);        // running INSE estimate, row-major [PT][LD]
  T* mean = reinterpret_cast<T*>(my + scratch_bytes + mat_bytes);
  T* aux = mean + 32;                                       // column sums of the shifted rows
  uint64_t* bar = reinterpret_cast<uint64_t*>(my + scratch_bytes + mat_bytes + 64 * sizeof(T));
  if (lane == 0) {
#pragma unroll
    for (int b = 0; b < kNBuf; ++b) st_mbar_init(bar + b);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  Ring<T, PC> rg;
  rg.s_iter = a.s_iter; rg.s_param = a.s_param; rg.n = n; rg.p_rt = P;
  uint32_t ring_phase = 0;
  rg.X = X; rg.mean = mean; rg.bar = bar; rg.phase = &ring_phase;
  rg.view(0, 0, kNBuf);
  const int h = lane >> 4, ta = (lane >> 2) & 3, tb = lane & 3;
  const T inv_n = T(1) / T(n);
  const bool want_second = a.out_cov || a.out_inse || a.out_ess;

  for (long ck = (long)blockIdx.x * kStatWarps + warp; ck < a.C; ck += (long)gridDim.x * kStatWarps) {
    const long c = a.index ? a.index[ck] : ck;
    rg.base = a.x + c * a.s_chain;
    rg.bulk = a.s_param == 1 && a.s_iter == P && (reinterpret_cast<uintptr_t>(rg.base) & 15) == 0;
    auto autocorrelation = [&]() {   // builder-defined; needs the true mean in `mean`
      const int K = a.max_lag;
      T den = T(1);
      for (int k0 = 0; k0 <= K; k0 += kAcfGroup) {
        T acc[kAcfGroup];
        pass_acf<T, PC>(rg, k0, acc, lane);
        if (k0 == 0) den = acc[0];
        if (lane < P) {
#pragma unroll
          for (int k = 0; k < kAcfGroup; ++k)
            if (k0 + k <= K) a.out_acf[(c * (K + 1) + k0 + k) * P + lane] = acc[k] / den;
        }
      }
    };
    if (!want_second) {
      // ---- mean (and autocorrelation) only --------------------------------------------------------------------------------
      mean[lane] = T(0);
      __syncwarp();
      pass_mean<T, PC>(rg, mean, lane);
      if (a.out_mean && lane < P) a.out_mean[c * P + lane] = mean[lane];
      if (a.out_acf) autocorrelation();
      continue;
    }

    // ---- lags 0 and 1 on SHIFTED rows u_i = x_i - s, s = the middle row of the chain (no separate pass for the means) ---------
    // With S = sum_i u_i, d = S / n:   sum_i (u_i - d)(u_i - d)^T = A0u - n d d^T,
    //   sum_{i<n-1} (u_i - d)(u_{i+1} - d)^T = A1u - d (S - u_0)^T - (S - u_{n-1}) d^T + (n - 1) d d^T;  mean = s + d.
    // s is a sample of the chain, so |d| is of the order of the chain's spread: no cancellation to speak of.
    mean[lane] = lane < P ? rg.base[(long)(n / 2) * a.s_iter + (long)lane * a.s_param] : T(0);
    __syncwarp();
    {
      T csum[TE];
      if constexpr (DUAL) {
        Tile<T, TE> a0, a1;
        pass_lag01<T, TE, PC, true>(rg, 0, a0, a1, csum, lane);
        __syncwarp();                       // the ring is dead: its memory takes the two product matrices
        a0.store(M0, LD, ta, tb, h);
        a1.store(M1, LD, ta, tb, h);
      } else {                              // large tiles: one lag per pass; lag 0 waits in Sg (free until Sigma_0 is formed)
        Tile<T, TE> acc;
        T unused[TE];
        pass_lag01<T, TE, PC, false>(rg, 0, acc, acc, csum, lane);
        __syncwarp();
        acc.store(Sg, LD, ta, tb, h);
        __syncwarp();
        pass_lag01<T, TE, PC, false>(rg, 1, acc, acc, unused, lane);
        __syncwarp();
        acc.store(M1, LD, ta, tb, h);
        __syncwarp();
        for (int e = lane; e < PT * LD; e += 32) M0[e] = Sg[e];
      }
#pragma unroll
      for (int r = 0; r < TE; ++r) {        // column sums: the two row slices, then the owner lanes publish them
        const T cs = csum[r] + shfl_xor_t(csum[r], 16);
        if (h == 0 && tb == 0) aux[ta + 4 * r] = cs;
      }
      __syncwarp();
    }
    // a column whose shifted values are all exactly zero never moved: no Sigma_m can be positive definite (its row and column
    // of every lagged product are exactly zero) -- 'Not enough samples' (inse_mc_cov.py:44-45) without walking through n / 2 lags
    const bool frozen = __any_sync(kFull, lane < P && M0[lane * LD + lane] == T(0));
    {
      const int jj = lane < P ? lane : 0;
      const T sj = mean[jj], Sj = aux[jj], dj = Sj * inv_n;
      const T u0 = rg.base[(long)jj * a.s_param] - sj;
      const T uL = rg.base[(long)(n - 1) * a.s_iter + (long)jj * a.s_param] - sj;
      for (int q = 0; q < P; ++q) {
        const T dq = shfl_t(dj, q), Sq = shfl_t(Sj, q), u0q = shfl_t(u0, q);
        if (lane < P) {
          const T dd = dj * dq;
          M0[lane * LD + q] -= T(n) * dd;
          M1[lane * LD + q] += T(n - 1) * dd - dj * (Sq - u0q) - (Sj - uL) * dq;
        }
      }
      __syncwarp();
      if (lane < P) mean[lane] = sj + dj;
      __syncwarp();
      if (a.out_mean && lane < P) a.out_mean[c * P + lane] = mean[lane];
    }
    // Wk: work copy for the factorisations (the third matrix of the dead ring's memory)
    T det_cov;
    {
      if (lane < P) {
        for (int q = 0; q < P; ++q) {
          const T a0rq = M0[lane * LD + q];
          const T cv = a0rq / T(n - 1);                                      // cov.py:13-15
          Wk[lane * LD + q] = cv;
          if (a.out_cov) a.out_cov[(c * P + lane) * P + q] = cv;
          // Gam = sym(gam0 + gam1), inse_mc_cov.py:32-33: pure additions before the scaling, so that elements (r, q) and
          // (q, r) are bitwise equal (the reference's is_pos_def demands exact symmetry); Sigma_0 = -gam0 + 2 Gam (:35-36)
          const T se = a0rq + M1[lane * LD + q];
          const T st = M0[q * LD + lane] + M1[q * LD + lane];
          const T gam = (se + st) * (T(0.5) * inv_n);
          Sg[lane * LD + q] = T(2) * gam - a0rq * inv_n;
        }
      }
      __syncwarp();
      det_cov = warp_det_lu<T>(Wk, P, LD, lane);
    }
    if (a.out_acf) autocorrelation();       // the ring takes its memory back (M0 / M1 have been consumed; Sg lives outside)
    if (!a.out_inse && !a.out_ess) continue;

    // ---- INSE (inse_mc_cov.py:20-73) ---------------------------------------------------------------------------------------
    // Sg holds Sigma_{m-1} (the last accepted estimate in phase 2); the candidate Sigma_m = Sg + 2 Gam_m is formed element by
    // element wherever it is needed (identical arithmetic each time)
    const int ub = n / 2;
    int sn = ub, m_last = -1;
    T last_det = T(0);
    bool phase2 = false;
    auto candidate = [&](T* dst) {   // dst = Sg + 2 sym(B) / n for m >= 1 (:37-38, :62); Sg itself for m = 0
      if (lane < P)
        for (int q = 0; q < P; ++q) {
          const T gam = (M0[lane * LD + q] + M0[q * LD + lane]) * (T(0.5) * inv_n);
          dst[lane * LD + q] = Sg[lane * LD + q] + T(2) * gam;
        }
      __syncwarp();
    };
    auto copy_sg = [&](T* dst) {
      if (lane < P)
        for (int q = 0; q < P; ++q) dst[lane * LD + q] = Sg[lane * LD + q];
      __syncwarp();
    };
    bool deferred = false;
    for (int m = 0; m < (frozen ? 0 : ub); ++m) {
      if (!phase2 && a.defer_m >= 0 && m > a.defer_m) { deferred = true; break; }
      if (m > 0) {
        Tile<T, TE> b;
        pass_pair<T, TE, PC>(rg, 2 * m, b, lane);
        __syncwarp();
        b.store(M0, LD, ta, tb, h);
        __syncwarp();
      }
      if (!phase2) {
        if (m > 0) candidate(Sg);                                          // Sigma_m is kept whether or not it is PD yet
        copy_sg(Wk);
        if (warp_chol_ok<T>(Wk, P, LD, lane)) {                            // is_pos_def(Sig), :40
          __syncwarp();
          copy_sg(Wk);
          last_det = warp_det_lu<T>(Wk, P, LD, lane);                      // last_dtm = det(Sig), :47
          sn = m; m_last = m; phase2 = true;
        }
      } else {
        candidate(Wk);
        const T cur = warp_det_lu<T>(Wk, P, LD, lane);                     // :66
        if (!(cur > last_det)) break;                                      // current_dtm <= last_dtm -> break, :68-69
        candidate(Sg);
        last_det = cur;
        m_last = m;
      }
      __syncwarp();
    }
    // 1 = 'Not enough samples' (:44-45); 3 = undecided after defer_m lag pairs: the caller finishes these chains in a second
    // launch (index list), where all of them run side by side instead of each holding up the warp that met it
    const int status = deferred ? 3 : (sn > ub - 1 ? 1 : 0);
    __syncwarp();
    if (a.out_inse && lane < P)
      for (int q = 0; q < P; ++q) a.out_inse[(c * P + lane) * P + q] = frozen ? qnan<T>() : Sg[lane * LD + q];
    // ---- multi-ESS (multi_ess.py:9-14) --------------------------------------------------------------------------------------
    if (lane == 0) {
      const double ratio = (double)det_cov / (double)last_det;
      const T ess = (T)((double)n * pow(ratio, 1.0 / (double)P));
      if (a.out_ess) a.out_ess[c] = status == 0 ? ess : qnan<T>();
      if (a.out_status) a.out_status[c] = status;
      if (a.out_lags) { a.out_lags[2 * c] = sn; a.out_lags[2 * c + 1] = m_last; }
    }
    __syncwarp();
  }
}

template <typename T> size_t stats_smem_bytes(int P, int TE) {
  const int PT = 4 * TE, LD = PT + 1;
  const size_t ring_bytes = ((size_t)kW * P * sizeof(T) + 15) & ~size_t(15);
  const size_t mat_bytes = ((size_t)PT * LD * sizeof(T) + 15) & ~size_t(15);
  const size_t scratch_bytes = ring_bytes > 3 * mat_bytes ? ring_bytes : 3 * mat_bytes;
  return kStatWarps * (scratch_bytes + mat_bytes + 64 * sizeof(T) + kNBuf * sizeof(uint64_t));
}

template <typename T, int TE, bool EXACT> cudaError_t launch_stats_te(const StatsArgs<T>& a, cudaStream_t st) {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t smem = stats_smem_bytes<T>(a.P, TE);
  auto kern = chain_stats_kernel<T, TE, EXACT>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const long want = (a.C + kStatWarps - 1) / kStatWarps;
  const long cap = (long)sms * 3 * 4;        // persistent-ish: a few CTAs per resident slot, warps stride over the chains
  const long grid = want < cap ? want : cap;
  kern<<<(unsigned)grid, kStatWarps * 32, smem, st>>>(a);
  return cudaGetLastError();
}

template <typename T> cudaError_t launch_stats(const StatsArgs<T>& a, cudaStream_t st) {
  switch ((a.P + 3) / 4) {
    case 1: return a.P == 4 ? launch_stats_te<T, 1, true>(a, st) : launch_stats_te<T, 1, false>(a, st);
    case 2: return a.P == 8 ? launch_stats_te<T, 2, true>(a, st) : launch_stats_te<T, 2, false>(a, st);
    case 3: return a.P == 12 ? launch_stats_te<T, 3, true>(a, st) : launch_stats_te<T, 3, false>(a, st);
    case 4: return a.P == 16 ? launch_stats_te<T, 4, true>(a, st) : launch_stats_te<T, 4, false>(a, st);
    case 5: return a.P == 20 ? launch_stats_te<T, 5, true>(a, st) : launch_stats_te<T, 5, false>(a, st);
    case 6: return a.P == 24 ? launch_stats_te<T, 6, true>(a, st) : launch_stats_te<T, 6, false>(a, st);
    case 7: return a.P == 28 ? launch_stats_te<T, 7, true>(a, st) : launch_stats_te<T, 7, false>(a, st);
    case 8: return a.P == 32 ? launch_stats_te<T, 8, true>(a, st) : launch_stats_te<T, 8, false>(a, st);
  }
  return cudaErrorInvalidValue;
}

}  // namespace eb

using namespace eb;

extern "C" {

// defined in capi.cu
int eeyore_b200_set_error_(int code, const char* msg);

int eeyore_b200_chain_stats(int dtype, int64_t n_chains, int64_t n_samples, int n_params, const void* samples,
                            int64_t ss_iter, int64_t ss_chain, int64_t ss_param, void* out_mean, void* out_cov,
                            void* out_inse, void* out_ess, int32_t* out_status, int32_t* out_lags, int max_lag,
                            void* out_acf, const int64_t* chain_index, int defer_after, void* stream) {
  if (!samples || n_chains < 1 || n_samples < 2 || n_params < 1)
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "chain_stats: bad sizes or null samples");
  if (n_params > 32) return eeyore_b200_set_error_(EEYORE_B200_EUNSUPPORTED, "chain_stats: at most 32 parameters per chain");
  if (n_samples >= (1LL << 30)) return eeyore_b200_set_error_(EEYORE_B200_EUNSUPPORTED, "chain_stats: at most 2^30 samples per chain");
  if (out_acf && (max_lag < 0 || max_lag >= n_samples))
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "acf: max_lag must be in [0, n_samples)");
  cudaError_t e;
  if (dtype == EEYORE_B200_F64) {
    StatsArgs<double> a{(const double*)samples, ss_iter, ss_chain, ss_param, (int)n_samples, n_params, n_chains,
                        (double*)out_mean, (double*)out_cov, (double*)out_inse, (double*)out_ess, out_status, out_lags,
                        max_lag, (double*)out_acf, (const long*)chain_index, defer_after};
    e = launch_stats<double>(a, (cudaStream_t)stream);
  } else if (dtype == EEYORE_B200_F32) {
    StatsArgs<float> a{(const float*)samples, ss_iter, ss_chain, ss_param, (int)n_samples, n_params, n_chains,
                       (float*)out_mean, (float*)out_cov, (float*)out_inse, (float*)out_ess, out_status, out_lags,
                       max_lag, (float*)out_acf, (const long*)chain_index, defer_after};
    e = launch_stats<float>(a, (cudaStream_t)stream);
  } else {
    return eeyore_b200_set_error_(EEYORE_B200_EINVAL, "dtype must be f32 or f64");
  }
  if (e != cudaSuccess) return eeyore_b200_set_error_(EEYORE_B200_ECUDA, cudaGetErrorString(e));
  return EEYORE_B200_OK;
}

}  // extern "C"
