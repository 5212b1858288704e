// Simplified manifold MALA with the expected-Fisher metric, one WARP per chain.
//
// The reference snapshot contains no SMMALA: this kernel is builder-defined and follows
// the reference's MALA structure (eeyore/samplers/mala.py:46-82) with a MultivariateNormalKernel(loc, scale_tril)
// proposal (eeyore/kernels/multivariate_normal_kernel.py:11-19):
//   G(theta) = T [ sum_i p_i (1 - p_i) J_i J_i^T + diag(1 / scale^2) ],  J_i = d a_L,i / d theta   (binary head)
//   G = R R^T (Cholesky, R lower);  mean(theta) = theta + step/2 G^-1 grad;  theta' = mean + sqrt(step) R^-T z
//   log q(b | a) = -(P/2) log(2 pi step) + sum_j log R_jj(a) - |R(a)^T (b - mean(a))|^2 / (2 step)
//   accept <=> Cholesky of G(theta') succeeded and log u < lt' - lt - log q(theta'|theta) + log q(theta|theta')
// Parity is checked against tests/golden/smmala_*.npz (runs assembled from the reference's own pieces, oracle/make_golden.py:
// smmala_goldens) and, at larger sizes, against oracle/samplers.py:smmala_run (itself pinned by those goldens).
//
// Work split inside the warp, per batch of 32 data rows:
//   phase A (lane = data row): forward pass, Jacobian row J_i, log-lik term, gradient += (y_i - p_i) J_i;
//                              J_i and w_i = p_i (1 - p_i) go to shared memory;
//   phase B (lane = 4x4 block of the lower triangle of G, two slices of the staged rows when the blocks fit 16 lanes):
//                              G_block += w_r J_r[a] J_r[b]; five 128-bit shared-memory loads per 16 FMAs (2x4 tiles fed by
//                              seven 64-bit loads per 8 FMAs kept the shared-memory pipe at 64 % of its peak).
// Then a warp-level in-place Cholesky, two triangular solves (lane = vector element) and the accept test.
#pragma once
#include "chain_kernels.cuh"

namespace eb {

// chains (warps) per block.  The kernel is latency-bound (serial factorisation / solves per chain), so resident warps are
// what counts: one 16-warp CTA per SM (the fp64 math tables cost 32 KB per CTA whatever its size; 11.3 KB per warp with the two
// factors stored as packed triangles; 128 registers per thread, no spills).
#ifndef EB_SM_WARPS
#define EB_SM_WARPS 16
#endif
constexpr int kSmWarps = EB_SM_WARPS;

template <class NET> struct SmGeom {
  static constexpr int P = NET::P;
  static constexpr int CQ = (P + 3) / 4;       // column / row quads
  static constexpr int PSV = 4 * CQ + 2;       // staged row: J (zero padded), w, one pad: rows are 16-byte aligned (LDS.128)
  static constexpr int LD = 0;                 // the P x P factors are stored as packed lower triangles (tri_at)
  static constexpr int MATN = P * (P + 1) / 2;
  static constexpr int NB = CQ * (CQ + 1) / 2; // 4x4 blocks of the lower triangle
  static constexpr int SLICES = NB <= 16 ? 2 : 1;
  // fp64: the metric is accumulated by FP64 tensor-core MMAs (DMMA m8n8k4) over 8 x 8 tiles of the lower triangle.  Row P of
  // the padded matrix carries the gradient (see smmala_eval), hence P + 1 rows.  Staged row: 8 NT columns, then w; the row
  // stride is = 4 (mod 8) doubles so that the 4 data rows x 4 columns a half-warp loads fall into 32 distinct banks.
  static constexpr int NT = (P + 1 + 7) / 8;
  static constexpr int PSV_D = ((8 * NT + 1 - 4 + 7) / 8) * 8 + 4;
  static_assert(PSV_D >= 8 * NT + 1 && PSV_D % 8 == 4, "staged row stride of the DMMA path");
  static_assert(P <= 32, "one lane per vector element");
  static_assert(NB <= 32, "one 4x4 metric block per lane");
};

template <typename T> struct SmIsF64 { static constexpr bool value = false; };
template <> struct SmIsF64<double> { static constexpr bool value = true; };

template <typename T, class NET> struct alignas(16) SmWarpMem {
  using Geo = SmGeom<NET>;
  static constexpr int PSV = SmIsF64<T>::value ? Geo::PSV_D : Geo::PSV;
  T v[32 * PSV];
  T mat[2][Geo::MATN + (Geo::MATN & 1)];   // metric / Cholesky factors, packed lower triangles: [cur], [proposal] (roles swap on accept)
  T dinv[2][32];                // 1 / R_jj
  T vec[32];                    // scratch vector (lane-indexed)
  T th[32];                     // the parameter vector under evaluation (read as broadcasts: no per-lane register copy)
};

template <typename T> struct SmemVec {   // read-only vector in shared memory with the [] of the register arrays
  const T* p;
  EB_HD T operator[](int j) const { return p[j]; }
};

// d a_L / d theta for one row (binary head, linear last pre-activation): back-propagation with seed 1.
template <typename T, class NET, class TH>
EB_HD void jacobian_row(const TH& th, const T* xr, T (&J)[NET::P], T& a_out) {
#pragma unroll
  for (int j = 0; j < NET::P; ++j) J[j] = T(0);
  T h0[NET::D0];
#pragma unroll
  for (int i = 0; i < NET::D0; ++i) h0[i] = xr[i];
  T h1[NET::D1];
  dense_fwd<T, NET::D0, NET::D1, NET::OFF0, true>(th, h0, h1);
  T one[1] = {T(1)};
  if constexpr (NET::NL == 2) {
    T a[NET::D2];
    dense_fwd<T, NET::D1, NET::D2, NET::OFF1, false>(th, h1, a);
    a_out = a[0];
    T d1[NET::D1], d0[NET::D0];
    dense_bwd<T, NET::D1, NET::D2, NET::OFF1, true>(th, h1, one, J, d1);
    dense_bwd<T, NET::D0, NET::D1, NET::OFF0, false>(th, h0, d1, J, d0);
  } else {
    T h2[NET::D2];
    dense_fwd<T, NET::D1, NET::D2, NET::OFF1, true>(th, h1, h2);
    T a[NET::DL];
    dense_fwd<T, NET::D2, NET::DL, NET::OFF2, false>(th, h2, a);
    a_out = a[0];
    T d2[NET::D2], d1[NET::D1], d0[NET::D0];
    dense_bwd<T, NET::D2, NET::DL, NET::OFF2, true>(th, h2, one, J, d2);
    dense_bwd<T, NET::D1, NET::D2, NET::OFF1, true>(th, h1, d2, J, d1);
    dense_bwd<T, NET::D0, NET::D1, NET::OFF0, false>(th, h0, d1, J, d0);
  }
}

// D (8 x 8, fp64) += A (8 x 4) B (4 x 8) on the FP64 tensor cores.  Fragments: lane l holds A[l / 4][l % 4], B[l % 4][l / 4]
// and D[l / 4][2 (l % 4) + {0, 1}].  On B200 a DMMA.8x8x4 keeps the FP64 pipe busy for 16 cycles (512 FLOP: the DFMA rate,
// tools/dmma_probe.cu) but costs ONE issue slot and two operand registers instead of 8 DFMAs fed by shared-memory loads.
EB_D void dmma884(double& d0, double& d1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

// Element (i, j), j <= i, of a lower-triangular P x P matrix in shared memory: LD > 0 = full rows with leading dimension LD,
// LD == 0 = packed rows (row i starts at i (i + 1) / 2; half the memory -- SMMALA keeps two factors per warp, and the packed
// form is what lets 16 warps share an SM).  Row starts i (i + 1) / 2 of consecutive lanes fall into distinct banks.
template <int LD> EB_HD constexpr int tri_at(int i, int j) { return LD > 0 ? i * LD + j : i * (i + 1) / 2 + j; }

template <typename T> EB_D T warp_sum(T v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(0xffffffffu, v, m);
  return v;
}

// In-place Cholesky of the lower triangle of m (P x P, leading dimension LD); dinv[j] = 1 / R_jj.
// Returns false if a pivot is not positive / not finite (linalg/is_pos_def.py:5-9 semantics); logdet = sum log R_jj.
// Right-looking (outer-product) form, lane = row: after column j is scaled, every lane subtracts L_ij L_cj from its own row
// entries c = j + 1 .. i -- independent updates, no serial dot products (the left-looking form spent sum_j 2 j dependent
// LDS + FMA pairs, 45 % of the kernel's stall samples).  Element (i, c) receives the same FMAs in the same order k = 0 .. c - 1 as
// before, so the factor is bit-identical.
template <typename T> EB_D T chol_log(T l) { return log_t<T>(l); }
template <> EB_D double chol_log<double>(double l) { return log_pos_normal(l); }   // l = sqrt(d), d > 0 finite: a normal number
// pivot l = sqrt(d) and its reciprocal; fp64: one coupled Newton chain for both inside the safe exponent range
template <typename T> EB_D void chol_pivot(T d, T& l, T& li) { l = sqrt_t<T>(d); li = T(1) / l; }
template <> EB_D void chol_pivot<double>(double d, double& l, double& li) {
  if (d > 1e-290 && d < 1e290) {            // uniform over the warp: every lane holds the same pivot
    sqrt_rsqrt_pos(d, &l, &li);
  } else {
    l = sqrt(d);
    li = 1.0 / l;
  }
}
template <typename T, int P, int LD> EB_D bool warp_chol_inplace(T* m, T* dinv, T& logdet) {
  const int lane = threadIdx.x & 31;
  bool ok = true;
  T ld = T(0);
  T* my = m + tri_at<LD>(lane < P ? lane : 0, 0);
  for (int j = 0; j < P; ++j) {
    const T d = m[tri_at<LD>(j, j)];
    if (!(d > T(0)) || !(d < T(INFINITY))) { ok = false; break; }     // uniform: every lane reads the same element
    T l, li;
    chol_pivot<T>(d, l, li);
    ld += chol_log<T>(l);
    const bool below = lane > j && lane < P;
    const T lij = below ? my[j] * li : T(0);
    if (below) my[j] = lij;
    if (lane == j) { my[j] = l; dinv[j] = li; }
    __syncwarp();
    if (below) {   // four entries per trip, loads first: the compiler cannot tell that column j and the row entries never alias
      int c = j + 1;
      for (; c + 3 <= lane; c += 4) {
        const T l0 = m[tri_at<LD>(c, j)], l1 = m[tri_at<LD>(c + 1, j)], l2 = m[tri_at<LD>(c + 2, j)], l3 = m[tri_at<LD>(c + 3, j)];
        const T a0 = my[c], a1 = my[c + 1], a2 = my[c + 2], a3 = my[c + 3];
        my[c] = fma_t<T>(-lij, l0, a0); my[c + 1] = fma_t<T>(-lij, l1, a1);
        my[c + 2] = fma_t<T>(-lij, l2, a2); my[c + 3] = fma_t<T>(-lij, l3, a3);
      }
      for (; c <= lane; ++c) my[c] = fma_t<T>(-lij, m[tri_at<LD>(c, j)], my[c]);
    }
    __syncwarp();
  }
  __syncwarp();
  logdet = ld;
  return ok;
}

// R y = b (forward substitution); lane i holds b_i on entry and y_i on return (i < P).
template <typename T, int P, int LD> EB_D T warp_solve_lower(const T* R, const T* dinv, T b) {
  const int lane = threadIdx.x & 31;
  const T* my = R + tri_at<LD>(lane < P ? lane : 0, 0);
  // the dependent chain is shuffle -> multiply -> FMA per step; the matrix entries and reciprocal pivots of four steps are
  // fetched together so that no shared-memory latency sits inside it
#pragma unroll 1
  for (int j0 = 0; j0 < P; j0 += 4) {
    T r[4], di[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int j = j0 + k < P ? j0 + k : P - 1; r[k] = my[j]; di[k] = dinv[j]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = j0 + k;
      if (j < P) {
        const T yj = __shfl_sync(0xffffffffu, b, j) * di[k];
        if (lane == j) b = yj;
        else if (lane > j && lane < P) b = fma_t<T>(-r[k], yj, b);
      }
    }
  }
  return b;
}

// R^T x = y (backward substitution); lane i holds y_i on entry and x_i on return.
template <typename T, int P, int LD> EB_D T warp_solve_upper_t(const T* R, const T* dinv, T y) {
  const int lane = threadIdx.x & 31;
#pragma unroll 1
  for (int j0 = P - 1; j0 >= 0; j0 -= 4) {
    T r[4], di[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) { const int j = j0 - k >= 0 ? j0 - k : 0; r[k] = R[tri_at<LD>(j, lane <= j ? lane : 0)]; di[k] = dinv[j]; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = j0 - k;
      if (j >= 0) {
        const T xj = __shfl_sync(0xffffffffu, y, j) * di[k];
        if (lane == j) y = xj;
        else if (lane < j) y = fma_t<T>(-r[k], xj, y);
      }
    }
  }
  return y;
}

// |R^T d|^2 with lane i holding d_i.
template <typename T, int P, int LD> EB_D T warp_rt_norm2(const T* R, T d, T* scratch) {
  const int lane = threadIdx.x & 31;
  if (lane < P) scratch[lane] = d;
  __syncwarp();
  T w = T(0);
  if (lane < P)
    for (int j = lane; j < P; ++j) w = fma_t<T>(R[tri_at<LD>(j, lane)], scratch[j], w);
  __syncwarp();
  return warp_sum<T>(w * w);
}

// log_target, gradient (replicated in every lane) and the metric's Cholesky factor (shared memory) at theta.
// theta is in sm.th (lane j wrote element j); on return lane j < P holds element j of the gradient in g_l.
template <typename T, class NET>
EB_D bool smmala_eval(const DataView<T>& d, SmWarpMem<T, NET>& sm, int buf, int tile_rp, int tile_cq, T& lt, T& g_l,
                      T& logdet) {
  using Geo = SmGeom<NET>;
  constexpr int P = NET::P, PSV = SmWarpMem<T, NET>::PSV, LD = Geo::LD;
  constexpr bool F64 = SmIsF64<T>::value;
  const int lane = threadIdx.x & 31;
  const SmemVec<T> th{sm.th};
  const int jl = lane < P ? lane : 0;
  T ll = T(0);
  g_l = T(0);
  T* G = sm.mat[buf];

  if constexpr (F64) {
    // ---- fp64: metric AND gradient on the FP64 tensor cores ---------------------------------------------------------------
    // Staged row r: [J_r (P) | delta_r | 0 .. | w_r].  With A = rows of (w J | delta) and B = (J | delta), the padded product
    // sum_r A_r B_r^T holds the metric in its leading P x P block and, in row P, sum_r delta_r J_r = the gradient of the
    // log-likelihood: no separate gradient pass over the staged rows.  Per four data rows a lane loads NT values + w from
    // shared memory (bank-conflict free, see PSV_D) and issues NT (NT + 1) / 2 DMMAs.
    constexpr int NT = Geo::NT, NTT = NT * (NT + 1) / 2;
    const int fk = lane & 3, fn = lane >> 2;
    double acc[NTT][2];
#pragma unroll
    for (int q = 0; q < NTT; ++q) acc[q][0] = acc[q][1] = 0.0;
    for (int base = 0; base < d.n_rows; base += 32) {
      const int i = base + lane;
      T* vrow = sm.v + lane * PSV;
      if (i < d.n_rows) {
        T J[P], a;
        jacobian_row<T, NET>(th, d.x + i * NET::D0, J, a);   // th: broadcast reads of sm.th
        T al[1] = {a}, dl[1], p;
        ll += head_loss<T, NET>(al, d.y[i], 0, dl, &p);
#pragma unroll
        for (int j = 0; j < P; ++j) vrow[j] = J[j];
        vrow[P] = dl[0];                       // d loglik_i / d a_L,i
#pragma unroll
        for (int j = P + 1; j < 8 * NT; ++j) vrow[j] = T(0);
        vrow[8 * NT] = p * (T(1) - p);
      } else {
#pragma unroll
        for (int j = 0; j <= 8 * NT; ++j) vrow[j] = T(0);
      }
      __syncwarp();
      const int rows = min(32, d.n_rows - base);
#pragma unroll 2
      for (int k0 = 0; k0 < rows; k0 += 4) {   // rows beyond the data were staged as zeros
        const T* vr = sm.v + (k0 + fk) * PSV;
        const T w = vr[8 * NT];
        T av[NT], bv[NT];
#pragma unroll
        for (int t = 0; t < NT; ++t) bv[t] = vr[8 * t + fn];
#pragma unroll
        for (int t = 0; t < NT; ++t) av[t] = bv[t] * ((8 * t + fn == P) ? T(1) : w);
        int q = 0;
#pragma unroll
        for (int ti = 0; ti < NT; ++ti)
#pragma unroll
          for (int tj = 0; tj <= ti; ++tj, ++q) dmma884(acc[q][0], acc[q][1], av[ti], bv[tj]);
      }
      __syncwarp();
    }
    // row P of the padded product -> the gradient, lane j = element j
    {
      constexpr int TG = P / 8;
      int q = TG * (TG + 1) / 2;
#pragma unroll
      for (int tj = 0; tj <= TG; ++tj, ++q)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int col = 8 * tj + 2 * fk + e;
          if (fn == P % 8 && col < P) sm.vec[col] = acc[q][e];
        }
      __syncwarp();
      g_l = sm.vec[jl];
      __syncwarp();
    }
    // the lower triangle of the metric
    {
      int q = 0;
#pragma unroll
      for (int ti = 0; ti < NT; ++ti)
#pragma unroll
        for (int tj = 0; tj <= ti; ++tj, ++q)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int row = 8 * ti + fn, col = 8 * tj + 2 * fk + e;
            if (row < P && col <= row) {
              T val = acc[q][e];
              if (row == col) val += d.pivar[row];
              if (d.has_temperature) val *= d.temperature;
              G[tri_at<LD>(row, col)] = val;
            }
          }
    }
  } else {
  const int slice = (Geo::SLICES == 2 && lane >= Geo::NB) ? 1 : 0;   // tile_rp / tile_cq: this lane's block (row quad, col quad)
  T acc[4][4];
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[r][c] = T(0);

  for (int base = 0; base < d.n_rows; base += 32) {
    const int i = base + lane;
    T* vrow = sm.v + lane * PSV;
    if (i < d.n_rows) {
      T J[P], a;
      jacobian_row<T, NET>(th, d.x + i * NET::D0, J, a);   // th: broadcast reads of sm.th
      T al[1] = {a}, dl[1], p;
      ll += head_loss<T, NET>(al, d.y[i], 0, dl, &p);
#pragma unroll
      for (int j = 0; j < P; ++j) vrow[j] = J[j];
#pragma unroll
      for (int j = P; j < 4 * Geo::CQ; ++j) vrow[j] = T(0);
      vrow[4 * Geo::CQ] = p * (T(1) - p);
      vrow[4 * Geo::CQ + 1] = dl[0];          // d loglik_i / d a_L,i: the gradient is J^T delta over the staged rows
    } else {
#pragma unroll
      for (int j = 0; j < PSV; ++j) vrow[j] = T(0);
    }
    __syncwarp();
    const int rows = min(32, d.n_rows - base);
    {   // gradient: lane j sums delta_r J_r[j] over the staged rows (replaces 20 per-lane accumulators + 20 warp reductions)
      T s0 = T(0), s1 = T(0);
#pragma unroll 4
      for (int r = 0; r + 1 < 32; r += 2) {
        s0 = fma_t<T>(sm.v[r * PSV + 4 * Geo::CQ + 1], sm.v[r * PSV + jl], s0);
        s1 = fma_t<T>(sm.v[(r + 1) * PSV + 4 * Geo::CQ + 1], sm.v[(r + 1) * PSV + jl], s1);
      }
      g_l += s0 + s1;                          // rows beyond the data were staged as zeros
    }
    if (tile_rp >= 0) {
#pragma unroll 2
      for (int r = slice; r < rows; r += Geo::SLICES) {
        const T* vr = sm.v + r * PSV;
        const T w = vr[4 * Geo::CQ];
        T av[4], bv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { av[q] = vr[4 * tile_rp + q]; bv[q] = vr[4 * tile_cq + q]; }
#pragma unroll
        for (int q = 0; q < 4; ++q) av[q] *= w;
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[q][c] = fma_t<T>(av[q], bv[c], acc[q][c]);
      }
    }
    __syncwarp();
  }
  // assemble the lower triangle of G in shared memory
#pragma unroll
  for (int r = 0; r < 4; ++r)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      T val = acc[r][c];
      if constexpr (Geo::SLICES == 2) val += __shfl_down_sync(0xffffffffu, val, Geo::NB);   // the other slice of the rows
      const int row = 4 * tile_rp + r, col = 4 * tile_cq + c;
      if (tile_rp >= 0 && slice == 0 && row < P && col <= row) {
        if (row == col) val += d.pivar[row];
        if (d.has_temperature) val *= d.temperature;
        G[tri_at<LD>(row, col)] = val;
      }
    }
  }
  ll = warp_sum<T>(ll);
  // prior: lane j owns parameter j; the value is summed in parameter order by every lane (as before: identical in all lanes)
  T lp = d.lp_const;
#pragma unroll
  for (int j = 0; j < P; ++j) {
    const T dd = th[j] - d.ploc[j];
    lp = fma_t<T>(-(dd * dd), T(0.5) * d.pivar[j], lp);
  }
  g_l = fma_t<T>(-(th[jl] - d.ploc[jl]), d.pivar[jl], g_l);
  if (d.has_temperature) { ll *= d.temperature; lp *= d.temperature; g_l *= d.temperature; }
  if (lane >= P) g_l = T(0);
  lt = ll + lp;
  __syncwarp();
  return warp_chol_inplace<T, P, LD>(G, sm.dinv[buf], logdet);
}

template <typename T, class NET>
__global__ void __launch_bounds__(kSmWarps * 32, 1) smmala_kernel(const ChainArgs<T> a) {
  using Geo = SmGeom<NET>;
  constexpr int P = NET::P, LD = Geo::LD;

  extern __shared__ __align__(16) unsigned char smem[];
  const SmemLayout<T, NET> lay(a.n_rows, kSmWarps, false);
  const DataView<T> d = stage_data<T, NET>(smem, lay, a);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  SmWarpMem<T, NET>& sm = reinterpret_cast<SmWarpMem<T, NET>*>(smem + align16(lay.total))[warp];
  long chain = (long)blockIdx.x * kSmWarps + warp;
  const bool live = chain < a.n_chains;
  if (!live) chain = a.n_chains - 1;

  // lane -> 4x4 block (row quad, column quad <= row quad) of the lower triangle; with two slices lanes NB .. 2 NB - 1 own the
  // same blocks for the odd staged rows
  int tile_rp = -1, tile_cq = -1;
  {
    int t = 0;
    const int want = lane < Geo::NB ? lane : (Geo::SLICES == 2 && lane < 2 * Geo::NB ? lane - Geo::NB : -1);
    for (int rq = 0; rq < Geo::CQ; ++rq)
      for (int cq = 0; cq <= rq; ++cq, ++t)
        if (t == want) { tile_rp = rq; tile_cq = cq; }
  }

  const T step = a.step, half_step = T(0.5) * step, sq_step = sqrt_t<T>(step);
  const T qconst = T(-0.5) * T(P) * log_t<T>(T(6.283185307179586) * step);
  const uint32_t gchain = a.chain0 + (uint32_t)chain;

  // current state: lane j holds element j of the sample, its gradient and the proposal mean
  T thc_l = T(0), gc_l = T(0), lt_c, logdet_c;
  int cur = 0;
  bool ok_c;
  {
    if (lane < P) thc_l = a.theta[chain * a.st_c + lane * a.st_p];
    sm.th[lane] = thc_l;
    __syncwarp();
    ok_c = smmala_eval<T, NET>(d, sm, cur, tile_rp, tile_cq, lt_c, gc_l, logdet_c);
  }
  T mean_c = thc_l + half_step * warp_solve_upper_t<T, P, LD>(sm.mat[cur], sm.dinv[cur],
                                                            warp_solve_lower<T, P, LD>(sm.mat[cur], sm.dinv[cur], gc_l));
  uint32_t n_acc = 0;

  for (long t = 0; t < a.n_iters; ++t) {
    // ---- noise: lane j owns z_j ------------------------------------------------------------------------------
    T zl = T(0), u;
    if (a.rng_mode == 0) {
      constexpr int PER = sizeof(T) == 8 ? 2 : 4;
      const int blk = lane / PER;
      if (lane < P) {
        U4 w = philox4x32_10(U4{(uint32_t)blk, a.iter0 + (uint32_t)t, gchain, 0u}, a.key.k0, a.key.k1);
        T v[4];
        if constexpr (sizeof(T) == 8) {
          box_muller<T>(Uni<double>::from(w.x, w.y), Uni<double>::from(w.z, w.w), &v[0], &v[1]);
          v[2] = v[3] = T(0);
        } else {
          box_muller<T>(Uni<float>::from(w.x), Uni<float>::from(w.y), &v[0], &v[1]);
          box_muller<T>(Uni<float>::from(w.z), Uni<float>::from(w.w), &v[2], &v[3]);
        }
        const int k = lane % PER;
        zl = k == 0 ? v[0] : (k == 1 ? v[1] : (k == 2 ? v[2] : v[3]));
      }
      u = philox_uniform<T>(a.key, gchain, a.iter0 + (uint32_t)t);
    } else {
      if (lane < P) zl = a.z_tape[((size_t)t * a.n_chains + chain) * P + lane];
      u = a.u_tape[(size_t)t * a.n_chains + chain];
    }
    // ---- proposal theta' = mean_c + sqrt(step) R^-T z ---------------------------------------------------------
    const T propl = mean_c + sq_step * warp_solve_upper_t<T, P, LD>(sm.mat[cur], sm.dinv[cur], zl);
    const int nxt = cur ^ 1;
    T lt_p, logdet_p, gpl = T(0);
    bool ok_p;
    __syncwarp();
    sm.th[lane] = propl;
    __syncwarp();
    ok_p = smmala_eval<T, NET>(d, sm, nxt, tile_rp, tile_cq, lt_p, gpl, logdet_p);
    bool acc = false;
    T mean_p = T(0);
    if (ok_p && ok_c) {
      mean_p = propl + half_step * warp_solve_upper_t<T, P, LD>(sm.mat[nxt], sm.dinv[nxt],
                                                              warp_solve_lower<T, P, LD>(sm.mat[nxt], sm.dinv[nxt], gpl));
      // forward density: theta' - mean_c = sqrt(step) R_c^-T z, so |R_c^T (theta' - mean_c)|^2 = step |z|^2 -- no product with R_c
      const T q_f = qconst + logdet_c - T(0.5) * warp_sum<T>(lane < P ? zl * zl : T(0));
      const T q_b = qconst + logdet_p - warp_rt_norm2<T, P, LD>(sm.mat[nxt], thc_l - mean_p, sm.vec) / (T(2) * step);
      const T log_rate = lt_p - lt_c - q_f + q_b;
      acc = log_t<T>(u) < log_rate;
    }
    if (acc) {
      ++n_acc;
      cur = nxt;
      lt_c = lt_p; logdet_c = logdet_p; mean_c = mean_p; thc_l = propl; gc_l = gpl;
    }
    if (t >= a.n_burnin && (t - a.n_burnin) % a.thin == 0 && live && lane < P) {
      const long s = (t - a.n_burnin) / a.thin;
      if (a.out_samples) a.out_samples[s * a.ss_i + chain * a.ss_c + lane * a.ss_p] = thc_l;
      if (a.out_grad) a.out_grad[s * a.ss_i + chain * a.ss_c + lane * a.ss_p] = gc_l;
      if (lane == 0) {
        if (a.out_target) a.out_target[s * a.n_chains + chain] = lt_c;
        if (a.out_acc) a.out_acc[s * a.n_chains + chain] = acc ? 1 : 0;
      }
    }
  }
  if (live) {
    if (lane < P) {
      a.theta[chain * a.st_c + lane * a.st_p] = thc_l;
      a.grad[chain * a.st_c + lane * a.st_p] = gc_l;
    }
    if (lane == 0) {
      a.target[chain] = lt_c;
      if (a.acc_count) a.acc_count[chain] += n_acc;
    }
  }
}

template <typename T, class NET> cudaError_t launch_smmala(const ChainArgs<T>& a, cudaStream_t st) {
  const SmemLayout<T, NET> lay(a.n_rows, kSmWarps, false);
  const size_t smem = align16(lay.total) + kSmWarps * sizeof(SmWarpMem<T, NET>);
  auto kern = smmala_kernel<T, NET>;
  cudaError_t e = reserve_smem(kern, smem);
  if (e != cudaSuccess) return e;
  const long blocks = (a.n_chains + kSmWarps - 1) / kSmWarps;
  kern<<<(unsigned)blocks, kSmWarps * 32, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace eb
