// Batched one-sided Jacobi SVD with truncation, the "decompose" half of
// cuquantum.tensornet.experimental.contract_decompose as the reference uses it for MPS
// gates (mps_utils.py:32-37, 78-84; SVDMethod semantics in SURVEY.md Appendix B).
//
// Formulation.  For W (n x m, n <= m, row-major) plane rotations from the left make the ROWS
// mutually orthogonal:  J W = diag(sigma) V^H  (up to order), so sigma_i = |row_i| and
// V^H_i = row_i / sigma_i.  No left factor is accumulated: the caller recovers
// U diag(f(sigma)) = W_in (row_i)^H f(sigma_i) / sigma_i^2 with one GEMM, which is exact
// for the product of the two emitted factors whatever the conditioning
// (A B = W_in V V^H).  A matrix with more rows than columns is handed in transposed.
//
// Blocking.  Rows are grouped in blocks of `block` rows; a CTA loads one PAIR of blocks into
// shared memory, rotates the block x block cross pairs there (one warp per row pair: three dot
// products by warp-shuffle reduction, then the rotation) and writes them back.
// A sweep = round-robin over block pairs plus one step for the pairs inside each block (one
// launch per step, all matrices of the batch in the same launch).  Matrices that fit one block pair run to convergence inside a
// single launch.  Sweeps stop when no rotation above `tol` happened (host reads one
// counter per matrix per sweep).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>

#include "qtn_internal.h"

namespace {

template <typename R> struct cx_of;
template <> struct cx_of<float> { using type = float2; };
template <> struct cx_of<double> { using type = double2; };
template <typename R> using cx = typename cx_of<R>::type;

constexpr int kWarps = 8;
constexpr int kSmemBudget = 160 * 1024;
constexpr int kWarpsWide = 16;                 // "svd_wide": 512-thread CTAs, one per SM, twice the block rows
constexpr int kSmemBudgetWide = 200 * 1024;

// Round-robin tournament ("circle method") on n players, n even: step in [0, n-1), p in [0, n/2).
__host__ __device__ inline void circle(int step, int p, int n, int &a, int &b) {
  const int m = n - 1;
  if (p == 0) { a = step % m; b = m; }
  else { a = (step + p) % m; b = (step - p + m) % m; }
}

template <typename R>
__device__ __forceinline__ R warp_sum(R v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sums of four values over the warp with 6 shuffles instead of 20: the halves of the warp trade two of
// the four values, then quarters trade one.  On return lane 0 holds sum(a), lane 8 sum(b), lane 16 sum(c),
// lane 24 sum(d) (in `a`); the other lanes hold copies of one of them.
template <typename R>
__device__ __forceinline__ R warp_sum4(R a, R b, R c, R d, int lane) {
  const bool h16 = lane & 16;
  R k0 = h16 ? c : a, k1 = h16 ? d : b;
  const R s0 = h16 ? a : c, s1 = h16 ? b : d;
  k0 += __shfl_xor_sync(0xffffffffu, s0, 16);
  k1 += __shfl_xor_sync(0xffffffffu, s1, 16);
  const bool h8 = lane & 8;
  R k = h8 ? k1 : k0;
  const R sx = h8 ? k0 : k1;
  k += __shfl_xor_sync(0xffffffffu, sx, 8);
  k += __shfl_xor_sync(0xffffffffu, k, 4);
  k += __shfl_xor_sync(0xffffffffu, k, 2);
  k += __shfl_xor_sync(0xffffffffu, k, 1);
  return k;
}

// Rotation that orthogonalises rows x, y with |x|^2 = a, |y|^2 = b, x.y^H = g:  x' = c x - sp y,
// y' = conj(sp) x + c y, the smaller of the two angles.  cos 2t = |d| / r with d = b - a, r = sqrt(d^2 + 4|g|^2);
// c = sqrt((1 + cos 2t) / 2), sp = sign(d) g / (r c): two rsqrt and no division (a double-precision sqrt or
// division costs ~20 FP64 instructions, which every lane of the warp issues).
// `shift` = |g|^2 / (r c^2) = 2|g|^2 / (r + |d|), signed like d: the rotated rows have |x'|^2 = a - shift and
// |y'|^2 = b + shift (the larger row grows).
__device__ __forceinline__ void jacobi_rotation(double a, double b, double gr, double gi, double &c, double &spr,
                                                double &spi, double &shift) {
  const double d = b - a;
  const double g2 = fma(gr, gr, gi * gi);
  const double r2 = fma(d, d, 4.0 * g2);
  const double inv_r = rsqrt(r2);
  const double c2 = fma(0.5 * fabs(d), inv_r, 0.5);
  const double inv_c = rsqrt(c2);
  c = c2 * inv_c;
  const double k = copysign(inv_r * inv_c, d);
  spr = gr * k;
  spi = gi * k;
  shift = g2 * k * inv_c;
}
__device__ __forceinline__ void jacobi_rotation(double a, double b, double gr, double gi, double &c, double &spr,
                                                double &spi) {
  double shift;
  jacobi_rotation(a, b, gr, gi, c, spr, spi, shift);
}

// Items whose cross steps run in jacobi_cross_reg_kernel (chosen by the host in choose_block).
constexpr int kRegBlockRows = 8;
__host__ __device__ inline bool reg_mode(const qtn_svd_item &it, int reg_cap) {
  return reg_cap > 0 && it.nblocks > 1 && it.block == kRegBlockRows && it.m <= reg_cap;
}

// NW warps per CTA.  When a round has fewer row pairs than warps, each pair is shared by wpp = NW / pairs
// warps: every warp takes a contiguous 1/wpp of the columns, the partial dot products meet in shared
// memory behind a named barrier of the pair's warps, and all of them derive the same rotation.
template <typename R, int NW>
__global__ void __launch_bounds__(32 * NW)
jacobi_step_kernel(const qtn_svd_item *__restrict__ items, int step, int sweep, int max_inner, R tol,
                   const int *__restrict__ rot_prev, int *__restrict__ rot_cur, int *__restrict__ changed,
                   long long ld, int gstep, int steps_per_sweep, int reg_cap) {
  using T = cx<R>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *rows = reinterpret_cast<T *>(smem_raw);
  const int item_id = blockIdx.y;
  if (sweep > 0 && rot_prev[item_id] == 0) return;          // this matrix has converged
  const qtn_svd_item it = items[item_id];
  const int n = it.n, m = it.m, b = it.block, nb = it.nblocks;
  if (n < 2) return;
  const int nbe = nb + (nb & 1);
  if (reg_mode(it, reg_cap) && step != nbe - 1) return;     // its cross steps run in jacobi_cross_reg_kernel
  // steps 0 .. nbe-2: block PAIRS (I, J), only the b*b cross pairs are rotated;
  // step nbe-1: every block alone, its b(b-1)/2 internal pairs.  One sweep = all of them.
  const int nsteps = nb == 1 ? 1 : nbe;
  if (step >= nsteps) return;
  int I = 0, J = -1;
  if (nb == 1) {
    if (blockIdx.x > 0) return;
  } else if (step == nbe - 1) {
    if ((int)blockIdx.x >= nb) return;
    I = blockIdx.x;
  } else {
    if ((int)blockIdx.x >= nbe / 2) return;
    circle(step, blockIdx.x, nbe, I, J);
    if (I >= nb || J >= nb) return;                           // paired with the dummy block: idle
  }
  // Skip the visit when neither block has been rotated since this very pair was last checked (one
  // sweep ago, same step): its rows are unchanged, hence still mutually orthogonal.  In the late
  // sweeps of a graded matrix only the small rows still move, and most visits end here.
  int *chg = changed + (long long)item_id * ld;
  if (nb > 1 && sweep > 0) {
    const int t_prev = gstep - steps_per_sweep;
    if (chg[I] < t_prev && (J < 0 || chg[J] < t_prev)) return;
  }
  const int nr = J < 0 ? b : 2 * b;                            // rows held by this CTA (even)
  T *W = (T *)it.w;
  auto grow = [&](int lr) { return lr < b ? I * b + lr : J * b + (lr - b); };
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  __shared__ R s_part[NW][4];
  for (int lr = warp; lr < nr; lr += NW) {
    const int g = grow(lr);
    T *dst = rows + (size_t)lr * m;
    if (g < n) { const T *src = W + (size_t)g * m; for (int c = lane; c < m; c += 32) dst[c] = src[c]; }
    else { T z; z.x = 0; z.y = 0; for (int c = lane; c < m; c += 32) dst[c] = z; }
  }
  __syncthreads();
  const int inner_max = nb == 1 ? max_inner : 1;
  int rotated_any = 0;
  for (int inner = 0; inner < inner_max; ++inner) {
    int rotated = 0;
    const int nrounds = J < 0 ? nr - 1 : b;
    const int npairs = J < 0 ? nr / 2 : b;
    // warps per pair (power of two; block sizes are powers of two), columns per warp
    int wpp = 1;
    while (wpp * 2 * npairs <= NW && (m / (wpp * 2)) >= 64) wpp <<= 1;
    const int sub = warp & (wpp - 1), group = warp / wpp, ngroups = NW / wpp;
    const int cchunk = (m + wpp - 1) / wpp;
    const int c0 = sub * cchunk, c1 = min(m, c0 + cchunk);
    for (int round = 0; round < nrounds; ++round) {
      for (int q = group; q < npairs; q += ngroups) {
        int x, y;
        if (J < 0) circle(round, q, nr, x, y);                  // all pairs inside one block
        else { x = q; y = b + ((q + round) & (b - 1)); }        // row q of I with row (q+round)%b of J
        if (grow(x) >= n || grow(y) >= n) continue;
        T *X = rows + (size_t)x * m, *Y = rows + (size_t)y * m;
        R a = 0, bb = 0, gr = 0, gi = 0;
        for (int c = c0 + lane; c < c1; c += 32) {
          const T u = X[c], v = Y[c];
          a = fma(u.x, u.x, fma(u.y, u.y, a));
          bb = fma(v.x, v.x, fma(v.y, v.y, bb));
          gr = fma(u.x, v.x, fma(u.y, v.y, gr));               // g = sum u * conj(v)
          gi = fma(u.y, v.x, fma(-u.x, v.y, gi));
        }
        a = warp_sum<R>(a); bb = warp_sum<R>(bb); gr = warp_sum<R>(gr); gi = warp_sum<R>(gi);
        if (wpp > 1) {
          if (lane == 0) { s_part[warp][0] = a; s_part[warp][1] = bb; s_part[warp][2] = gr; s_part[warp][3] = gi; }
          asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(32 * wpp) : "memory");
          a = 0; bb = 0; gr = 0; gi = 0;
          for (int w = group * wpp; w < (group + 1) * wpp; ++w) {   // same order in every warp
            a += s_part[w][0]; bb += s_part[w][1]; gr += s_part[w][2]; gi += s_part[w][3];
          }
          asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(32 * wpp) : "memory");   // s_part is reused
        }
        const R g2 = gr * gr + gi * gi;
        // relative criterion at every scale (rows far below sigma_max are orthogonalised as
        // accurately as the leading ones); a row whose squared norm underflowed is left alone
        if (!(g2 > tol * tol * a * bb) || !(fmin(a, bb) > (R)0)) continue;
        rotated = 1;
        double cd, sprd, spid;
        jacobi_rotation((double)a, (double)bb, (double)gr, (double)gi, cd, sprd, spid);
        const R c = (R)cd, spr = (R)sprd, spi = (R)spid;                // s * e^{i phi}
        for (int col = c0 + lane; col < c1; col += 32) {
          const T u = X[col], v = Y[col];
          T xn, yn;
          // x' = c x - s e^{i phi} y ;  y' = s e^{-i phi} x + c y
          xn.x = c * u.x - (spr * v.x - spi * v.y);
          xn.y = c * u.y - (spr * v.y + spi * v.x);
          yn.x = c * v.x + (spr * u.x + spi * u.y);
          yn.y = c * v.y + (spr * u.y - spi * u.x);
          X[col] = xn; Y[col] = yn;
        }
      }
      __syncthreads();
    }
    rotated = __syncthreads_or(rotated);
    rotated_any = rotated;
    if (!rotated) break;
  }
  for (int lr = warp; lr < nr; lr += NW) {
    const int g = grow(lr);
    if (g >= n) continue;
    const T *src = rows + (size_t)lr * m;
    T *dst = W + (size_t)g * m;
    for (int c = lane; c < m; c += 32) dst[c] = src[c];
  }
  if (t == 0 && rotated_any) {
    atomicAdd(&rot_cur[item_id], 1);
    chg[I] = gstep;
    if (J >= 0) chg[J] = gstep;
  }
}

// ---------------------------------------------------------------------------------------------
// Register-resident cross step.  The shared-memory kernel above moves 48 KB through shared memory per
// 512-column complex128 row pair (two reads + one write of both rows) against ~10 k DFMAs: it is bound
// by shared-memory bandwidth, not by the FP64 pipe.  Here a block pair is 8 + 8 rows; row q of block I
// lives in the REGISTERS of warp group q (WPP warps, CPL columns per lane each) for the whole visit and
// goes global -> registers -> global without touching shared memory; only block J is staged in shared
// memory, read and written once per pair (16 KB).  The pairs of a block (step nbe-1) and matrices that
// fit one block pair stay with jacobi_step_kernel.
// ---------------------------------------------------------------------------------------------
constexpr int kRegBlock = kRegBlockRows;

template <typename R, int CPL, int WPP>
__global__ void __launch_bounds__(32 * kRegBlock * WPP)
jacobi_cross_reg_kernel(const qtn_svd_item *__restrict__ items, int step, int sweep, R tol,
                        const int *__restrict__ rot_prev, int *__restrict__ rot_cur, int *__restrict__ changed,
                        long long ld, int gstep, int steps_per_sweep, int reg_cap) {
  using T = cx<R>;
  constexpr int b = kRegBlock;
  constexpr int NW = b * WPP;
  constexpr int MP = 32 * WPP * CPL;                         // padded row length held by a warp group
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *ys = reinterpret_cast<T *>(smem_raw);                  // b x MP
  __shared__ R s_part[NW][4];
  const int item_id = blockIdx.y;
  if (sweep > 0 && rot_prev[item_id] == 0) return;
  const qtn_svd_item it = items[item_id];
  if (!reg_mode(it, reg_cap)) return;
  const int n = it.n, m = it.m, nb = it.nblocks;
  const int nbe = nb + (nb & 1);
  if (step >= nbe - 1 || (int)blockIdx.x >= nbe / 2) return;
  int I, J;
  circle(step, blockIdx.x, nbe, I, J);
  if (I >= nb || J >= nb) return;
  int *chg = changed + (long long)item_id * ld;
  if (sweep > 0) {
    const int t_prev = gstep - steps_per_sweep;
    if (chg[I] < t_prev && chg[J] < t_prev) return;
  }
  T *W = (T *)it.w;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int group = warp / WPP, sub = warp % WPP;
  const int cbase = sub * 32 * CPL + lane;                   // my columns: cbase + 32 k
  T xr[CPL];
  {
    const int gx = I * b + group, gy = J * b + group;
    const T *sx = W + (size_t)gx * m, *sy = W + (size_t)gy * m;
    T *dy = ys + (size_t)group * MP;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int c = cbase + 32 * k;
      T z; z.x = 0; z.y = 0;
      xr[k] = (gx < n && c < m) ? sx[c] : z;
      dy[c] = (gy < n && c < m) ? sy[c] : z;
    }
  }
  __syncthreads();
  int rotated = 0;
  for (int round = 0; round < b; ++round) {
    T *Y = ys + (size_t)((group + round) & (b - 1)) * MP;
    T yr[CPL];
    R a = 0, bb = 0, gr = 0, gi = 0;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      yr[k] = Y[cbase + 32 * k];
      const T u = xr[k], v = yr[k];
      a = fma(u.x, u.x, fma(u.y, u.y, a));
      bb = fma(v.x, v.x, fma(v.y, v.y, bb));
      gr = fma(u.x, v.x, fma(u.y, v.y, gr));                   // g = sum u * conj(v)
      gi = fma(u.y, v.x, fma(-u.x, v.y, gi));
    }
    if (WPP > 1) {
      const R part = warp_sum4<R>(a, bb, gr, gi, lane);
      if ((lane & 7) == 0) s_part[warp][lane >> 3] = part;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(32 * WPP) : "memory");
      a = 0; bb = 0; gr = 0; gi = 0;
#pragma unroll
      for (int w = 0; w < WPP; ++w) {                          // same order in every warp of the group
        a += s_part[group * WPP + w][0]; bb += s_part[group * WPP + w][1];
        gr += s_part[group * WPP + w][2]; gi += s_part[group * WPP + w][3];
      }
    } else {
      a = warp_sum<R>(a); bb = warp_sum<R>(bb); gr = warp_sum<R>(gr); gi = warp_sum<R>(gi);
    }
    const R g2 = gr * gr + gi * gi;
    if ((g2 > tol * tol * a * bb) && (fmin(a, bb) > (R)0)) {
      rotated = 1;
      double cd, sprd, spid;
      jacobi_rotation((double)a, (double)bb, (double)gr, (double)gi, cd, sprd, spid);
      const R c = (R)cd, spr = (R)sprd, spi = (R)spid;             // s * e^{i phi}
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const T u = xr[k], v = yr[k];
        T xn, yn;
        xn.x = c * u.x - (spr * v.x - spi * v.y);
        xn.y = c * u.y - (spr * v.y + spi * v.x);
        yn.x = c * v.x + (spr * u.x + spi * u.y);
        yn.y = c * v.y + (spr * u.y - spi * u.x);
        xr[k] = xn;
        Y[cbase + 32 * k] = yn;
      }
    }
    __syncthreads();                                           // the Y rows change hands; s_part is reused
  }
  {
    const int gx = I * b + group, gy = J * b + group;
    T *dx = W + (size_t)gx * m, *dyg = W + (size_t)gy * m;
    const T *sy = ys + (size_t)group * MP;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int c = cbase + 32 * k;
      if (gx < n && c < m) dx[c] = xr[k];
      if (gy < n && c < m) dyg[c] = sy[c];
    }
  }
  rotated = __syncthreads_or(rotated);
  if (t == 0 && rotated) {
    atomicAdd(&rot_cur[item_id], 1);
    chg[I] = gstep;
    chg[J] = gstep;
  }
}

// ---------------------------------------------------------------------------------------------
// Streaming variant of the register-resident cross step: a persistent CTA walks the (matrix, block pair)
// visits of one step; while a visit is being rotated the rows of the next one arrive in shared memory by
// cp.async (X rows into a staging buffer, Y rows into the second of two Y buffers), so the DRAM/L2 latency
// of a visit start is hidden and the tail of one visit overlaps the head of the next.  Row norms are
// computed once per visit (when the rows arrive) and carried through the 8 rounds with the exact update
// |x'|^2 = a - shift, |y'|^2 = b + shift, which halves the dot-product work per pair; a sweep without
// rotations -- the only kind that ends the iteration -- uses freshly computed norms only.
// ---------------------------------------------------------------------------------------------
template <int BYTES>
__device__ __forceinline__ void cp_async_zfill(void *smem_dst, const void *gsrc, bool valid) {
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int src_bytes = valid ? BYTES : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2, %3;" ::"r"(dst), "l"(gsrc), "n"(BYTES), "r"(src_bytes)
               : "memory");
}

template <typename R, int CPL, int WPP>
__global__ void __launch_bounds__(32 * kRegBlock * WPP, 1)
jacobi_cross_stream_kernel(const qtn_svd_item *__restrict__ items, int step, int sweep, R tol,
                           const int *__restrict__ rot_prev, int *__restrict__ rot_cur, int *__restrict__ changed,
                           long long ld, int gstep, int steps_per_sweep, int reg_cap, int pairs_max, int batch) {
  using T = cx<R>;
  constexpr int b = kRegBlock;
  constexpr int NW = b * WPP;
  constexpr int MP = 32 * WPP * CPL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *xbuf = reinterpret_cast<T *>(smem_raw);                 // b x MP: X rows of the coming visit
  T *ybufs = xbuf + (size_t)b * MP;                          // 2 x (b x MP): Y rows, current / coming
  __shared__ R s_part[NW][2];
  __shared__ R s_yn[b][WPP];                                 // |y row|^2 as WPP partial sums
  __shared__ R s_xn[NW];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int group = warp / WPP, sub = warp % WPP;
  const int cbase = sub * 32 * CPL + lane;
  const int total = pairs_max * batch;

  // next visit at or after virtual index w that has work (same answer in every thread)
  auto next_work = [&](int w, int &item_id, int &I, int &J) {
    for (; w < total; w += (int)gridDim.x) {
      item_id = w / pairs_max;
      const int p = w - item_id * pairs_max;
      if (sweep > 0 && rot_prev[item_id] == 0) continue;
      const qtn_svd_item it = items[item_id];
      if (!reg_mode(it, reg_cap)) continue;
      const int nb = it.nblocks, nbe = nb + (nb & 1);
      if (step >= nbe - 1 || p >= nbe / 2) continue;
      circle(step, p, nbe, I, J);
      if (I >= nb || J >= nb) continue;
      if (sweep > 0) {
        const int *chg = changed + (long long)item_id * ld;
        const int t_prev = gstep - steps_per_sweep;
        if (chg[I] < t_prev && chg[J] < t_prev) continue;
      }
      return w;
    }
    return total;
  };
  auto prefetch = [&](int item_id, int I, int J, T *ydst) {
    const qtn_svd_item it = items[item_id];
    const T *W = (const T *)it.w;
    const int n = it.n, m = it.m;
    const int gx = I * b + group, gy = J * b + group;
    const T *sx = W + (size_t)gx * m, *sy = W + (size_t)gy * m;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int c = cbase + 32 * k;
      const bool vx = gx < n && c < m, vy = gy < n && c < m;
      cp_async_zfill<(int)sizeof(T)>(xbuf + (size_t)group * MP + c, vx ? (const void *)(sx + c) : (const void *)W, vx);
      cp_async_zfill<(int)sizeof(T)>(ydst + (size_t)group * MP + c, vy ? (const void *)(sy + c) : (const void *)W, vy);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  int item_id = 0, I = 0, J = 0, ycur = 0;
  int w = next_work((int)blockIdx.x, item_id, I, J);
  if (w < total) prefetch(item_id, I, J, ybufs);
  while (w < total) {
    T *Yb = ybufs + (size_t)ycur * b * MP;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    T xr[CPL];
    {
      R pa = 0, pb = 0;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int c = cbase + 32 * k;
        xr[k] = xbuf[(size_t)group * MP + c];
        const T v = Yb[(size_t)group * MP + c];
        pa = fma(xr[k].x, xr[k].x, fma(xr[k].y, xr[k].y, pa));
        pb = fma(v.x, v.x, fma(v.y, v.y, pb));
      }
      pa = warp_sum<R>(pa); pb = warp_sum<R>(pb);
      if (lane == 0) { s_xn[warp] = pa; s_yn[group][sub] = pb; }
    }
    __syncthreads();                                           // xbuf is free, the norms are visible
    int n_item, nI, nJ;
    const int wn = next_work(w + (int)gridDim.x, n_item, nI, nJ);
    if (wn < total) prefetch(n_item, nI, nJ, ybufs + (size_t)(ycur ^ 1) * b * MP);
    R ax = 0;
#pragma unroll
    for (int q = 0; q < WPP; ++q) ax += s_xn[group * WPP + q];
    int rotated = 0;
    for (int round = 0; round < b; ++round) {
      const int yrow = (group + round) & (b - 1);
      T *Y = Yb + (size_t)yrow * MP;
      T yr[CPL];
#pragma unroll
      for (int k = 0; k < CPL; ++k) yr[k] = Y[cbase + 32 * k];
      R gr = 0, gi = 0, gr1 = 0, gi1 = 0, gr2 = 0, gi2 = 0, gr3 = 0, gi3 = 0;   // four chains each: DFMA latency
#pragma unroll
      for (int k = 0; k < CPL; k += 2) {
        const T u = xr[k], v = yr[k], u1 = xr[k + 1], v1 = yr[k + 1];
        gr = fma(u.x, v.x, gr); gr1 = fma(u.y, v.y, gr1);       // g = sum u * conj(v)
        gi = fma(u.y, v.x, gi); gi1 = fma(-u.x, v.y, gi1);
        gr2 = fma(u1.x, v1.x, gr2); gr3 = fma(u1.y, v1.y, gr3);
        gi2 = fma(u1.y, v1.x, gi2); gi3 = fma(-u1.x, v1.y, gi3);
      }
      gr = (gr + gr1) + (gr2 + gr3);
      gi = (gi + gi1) + (gi2 + gi3);
      {
        const bool h16 = lane & 16;
        R k = h16 ? gi : gr;
        const R sx = h16 ? gr : gi;
        k += __shfl_xor_sync(0xffffffffu, sx, 16);
        k += __shfl_xor_sync(0xffffffffu, k, 8);
        k += __shfl_xor_sync(0xffffffffu, k, 4);
        k += __shfl_xor_sync(0xffffffffu, k, 2);
        k += __shfl_xor_sync(0xffffffffu, k, 1);
        if ((lane & 15) == 0) s_part[warp][lane >> 4] = k;
      }
      R bb = 0;                                                // read before the barrier: a partner warp that
#pragma unroll                                                 // is ahead rewrites s_yn[yrow] after its rotation
      for (int q = 0; q < WPP; ++q) bb += s_yn[yrow][q];
      asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(32 * WPP) : "memory");
      gr = 0; gi = 0;
#pragma unroll
      for (int q = 0; q < WPP; ++q) {                          // same order in every warp of the group
        gr += s_part[group * WPP + q][0]; gi += s_part[group * WPP + q][1];
      }
      const R g2 = gr * gr + gi * gi;
      if ((g2 > tol * tol * ax * bb) && (fmin(ax, bb) > (R)0)) {
        rotated = 1;
        double cd, sprd, spid, shift;
        jacobi_rotation((double)ax, (double)bb, (double)gr, (double)gi, cd, sprd, spid, shift);
        const R c = (R)cd, spr = (R)sprd, spi = (R)spid;           // s * e^{i phi}
#pragma unroll
        for (int k = 0; k < CPL; ++k) {
          const T u = xr[k], v = yr[k];
          T xn, yn;
          xn.x = c * u.x - (spr * v.x - spi * v.y);
          xn.y = c * u.y - (spr * v.y + spi * v.x);
          yn.x = c * v.x + (spr * u.x + spi * u.y);
          yn.y = c * v.y + (spr * u.y - spi * u.x);
          xr[k] = xn;
          Y[cbase + 32 * k] = yn;
        }
        ax = (R)fmax((double)ax - shift, 0.0);
        if (lane == 0) s_yn[yrow][sub] = sub == 0 ? (R)fmax((double)bb + shift, 0.0) : (R)0;
      }
      __syncthreads();                                         // the Y rows change hands; s_part, s_yn reused
    }
    {
      const qtn_svd_item it = items[item_id];
      T *W = (T *)it.w;
      const int n = it.n, m = it.m;
      const int gx = I * b + group, gy = J * b + group;
      T *dx = W + (size_t)gx * m, *dyg = W + (size_t)gy * m;
      const T *sy = Yb + (size_t)group * MP;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int c = cbase + 32 * k;
        if (gx < n && c < m) dx[c] = xr[k];
        if (gy < n && c < m) dyg[c] = sy[c];
      }
    }
    rotated = __syncthreads_or(rotated);                       // also: Yb may be overwritten from here on
    if (t == 0 && rotated) {
      atomicAdd(&rot_cur[item_id], 1);
      int *chg = changed + (long long)item_id * ld;
      chg[I] = gstep;
      chg[J] = gstep;
    }
    w = wn; item_id = n_item; I = nI; J = nJ; ycur ^= 1;
  }
}

// ---------------------------------------------------------------------------------------------
// Two-pass variant for TWO resident CTAs per SM.  With every warp of a CTA walking the same phases
// (shared-memory read, dot, shuffle reduction, scalar rotation set-up, rotation, shared-memory write) the
// FP64 pipe and the shared-memory pipe take turns idling (ncu: both ~35-40 % busy).  Here the Y row is not
// held in registers between the dot product and the rotation but read from shared memory twice, which keeps
// the kernel within 64 registers so that two CTAs (two independent block pairs) share an SM and fill each
// other's gaps; the price is 24 KB instead of 16 KB of shared-memory traffic per row pair.  One visit per CTA
// (the second CTA hides the load/store latency), norms carried as in the streaming kernel.
// ---------------------------------------------------------------------------------------------
template <typename R, int CPL, int WPP>
__global__ void __launch_bounds__(32 * kRegBlock * WPP, 2)
jacobi_cross_dual_kernel(const qtn_svd_item *__restrict__ items, int step, int sweep, R tol,
                         const int *__restrict__ rot_prev, int *__restrict__ rot_cur, int *__restrict__ changed,
                         long long ld, int gstep, int steps_per_sweep, int reg_cap) {
  using T = cx<R>;
  constexpr int b = kRegBlock;
  constexpr int NW = b * WPP;
  constexpr int MP = 32 * WPP * CPL;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *ys = reinterpret_cast<T *>(smem_raw);                  // b x MP
  __shared__ R s_part[NW][2];
  __shared__ R s_yn[b][WPP];
  __shared__ R s_xn[NW];
  const int item_id = blockIdx.y;
  if (sweep > 0 && rot_prev[item_id] == 0) return;
  const qtn_svd_item it = items[item_id];
  if (!reg_mode(it, reg_cap)) return;
  const int n = it.n, m = it.m, nb = it.nblocks;
  const int nbe = nb + (nb & 1);
  if (step >= nbe - 1 || (int)blockIdx.x >= nbe / 2) return;
  int I, J;
  circle(step, blockIdx.x, nbe, I, J);
  if (I >= nb || J >= nb) return;
  int *chg = changed + (long long)item_id * ld;
  if (sweep > 0) {
    const int t_prev = gstep - steps_per_sweep;
    if (chg[I] < t_prev && chg[J] < t_prev) return;
  }
  T *W = (T *)it.w;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int group = warp / WPP, sub = warp % WPP;
  const int cbase = sub * 32 * CPL + lane;                   // my columns: cbase + 32 k
  T xr[CPL];
  {
    const int gx = I * b + group, gy = J * b + group;
    const T *sx = W + (size_t)gx * m, *sy = W + (size_t)gy * m;
    T *dy = ys + (size_t)group * MP;
    R pa = 0, pb = 0;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int c = cbase + 32 * k;
      T z; z.x = 0; z.y = 0;
      xr[k] = (gx < n && c < m) ? sx[c] : z;
      const T v = (gy < n && c < m) ? sy[c] : z;
      dy[c] = v;
      pa = fma(xr[k].x, xr[k].x, fma(xr[k].y, xr[k].y, pa));
      pb = fma(v.x, v.x, fma(v.y, v.y, pb));
    }
    pa = warp_sum<R>(pa); pb = warp_sum<R>(pb);
    if (lane == 0) { s_xn[warp] = pa; s_yn[group][sub] = pb; }
  }
  __syncthreads();
  R ax = 0;
#pragma unroll
  for (int q = 0; q < WPP; ++q) ax += s_xn[group * WPP + q];
  int rotated = 0;
  for (int round = 0; round < b; ++round) {
    const int yrow = (group + round) & (b - 1);
    T *Y = ys + (size_t)yrow * MP + cbase;
    R gr = 0, gi = 0, gr1 = 0, gi1 = 0;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const T u = xr[k], v = Y[32 * k];
      gr = fma(u.x, v.x, gr); gr1 = fma(u.y, v.y, gr1);         // g = sum u * conj(v)
      gi = fma(u.y, v.x, gi); gi1 = fma(-u.x, v.y, gi1);
    }
    gr += gr1; gi += gi1;
    R bb = 0;                                                  // before the barrier (see the streaming kernel)
#pragma unroll
    for (int q = 0; q < WPP; ++q) bb += s_yn[yrow][q];
    {
      const bool h16 = lane & 16;
      R k = h16 ? gi : gr;
      const R sx = h16 ? gr : gi;
      k += __shfl_xor_sync(0xffffffffu, sx, 16);
      k += __shfl_xor_sync(0xffffffffu, k, 8);
      k += __shfl_xor_sync(0xffffffffu, k, 4);
      k += __shfl_xor_sync(0xffffffffu, k, 2);
      k += __shfl_xor_sync(0xffffffffu, k, 1);
      if (WPP > 1) {
        if ((lane & 15) == 0) s_part[warp][lane >> 4] = k;
      } else {                                                 // one warp per pair: no exchange
        gr = __shfl_sync(0xffffffffu, k, 0);
        gi = __shfl_sync(0xffffffffu, k, 16);
      }
    }
    if (WPP > 1) {
      asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "r"(32 * WPP) : "memory");
      gr = 0; gi = 0;
#pragma unroll
      for (int q = 0; q < WPP; ++q) { gr += s_part[group * WPP + q][0]; gi += s_part[group * WPP + q][1]; }
    }
    const R g2 = gr * gr + gi * gi;
    if ((g2 > tol * tol * ax * bb) && (fmin(ax, bb) > (R)0)) {
      rotated = 1;
      double cd, sprd, spid, shift;
      jacobi_rotation((double)ax, (double)bb, (double)gr, (double)gi, cd, sprd, spid, shift);
      const R c = (R)cd, spr = (R)sprd, spi = (R)spid;             // s * e^{i phi}
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const T u = xr[k], v = Y[32 * k];
        T xn, yn;
        xn.x = c * u.x - (spr * v.x - spi * v.y);
        xn.y = c * u.y - (spr * v.y + spi * v.x);
        yn.x = c * v.x + (spr * u.x + spi * u.y);
        yn.y = c * v.y + (spr * u.y - spi * u.x);
        xr[k] = xn;
        Y[32 * k] = yn;
      }
      ax = (R)fmax((double)ax - shift, 0.0);
      if (lane == 0) s_yn[yrow][sub] = sub == 0 ? (R)fmax((double)bb + shift, 0.0) : (R)0;
    }
    __syncthreads();                                           // the Y rows change hands; s_part, s_yn reused
  }
  {
    const int gx = I * b + group, gy = J * b + group;
    T *dx = W + (size_t)gx * m, *dyg = W + (size_t)gy * m;
    const T *sy = ys + (size_t)group * MP;
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int c = cbase + 32 * k;
      if (gx < n && c < m) dx[c] = xr[k];
      if (gy < n && c < m) dyg[c] = sy[c];
    }
  }
  rotated = __syncthreads_or(rotated);
  if (t == 0 && rotated) {
    atomicAdd(&rot_cur[item_id], 1);
    chg[I] = gstep;
    chg[J] = gstep;
  }
}

// Row norms -> singular values sorted descending (+ permutation), truncation rank, S normalisation.
template <typename R>
__global__ void __launch_bounds__(256)
finalize_kernel(const qtn_svd_item *__restrict__ items, qtn_svd_trunc tr, double *__restrict__ scratch,
                double *__restrict__ sigma, int32_t *__restrict__ perm, long long ld,
                int32_t *__restrict__ rank, double *__restrict__ scale) {
  using T = cx<R>;
  const qtn_svd_item it = items[blockIdx.x];
  const int n = it.n, m = it.m;
  const T *W = (const T *)it.w;
  double *raw = scratch + (long long)blockIdx.x * ld;
  double *sg = sigma + (long long)blockIdx.x * ld;
  int32_t *pm = perm + (long long)blockIdx.x * ld;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int r = warp; r < n; r += kWarps) {
    double s = 0;
    for (int c = lane; c < m; c += 32) { const T v = W[(size_t)r * m + c]; s += (double)v.x * v.x + (double)v.y * v.y; }
    s = warp_sum<double>(s);
    if (lane == 0) raw[r] = sqrt(s);
  }
  __syncthreads();
  for (int i = t; i < n; i += 256) {
    const double si = raw[i];
    int rk = 0;
    for (int j = 0; j < n; ++j) { const double sj = raw[j]; rk += (sj > si) || (sj == si && j < i); }
    sg[rk] = si; pm[rk] = i;
  }
  __syncthreads();
  if (t == 0) {
    // Noise floor: the left factor is recovered as theta * V / sigma, which is only meaningful for
    // sigma well above the rounding level of the working precision; directions below
    // 16 eps sigma_max are numerical noise of the input and are dropped whatever the cutoffs say
    // (their total weight is below the precision of the state itself).
    const double eps = sizeof(R) == 8 ? 1.1e-16 : 6.0e-8;
    int keep = 0;
    const double thr = fmax(fmax(tr.abs_cutoff, tr.rel_cutoff * sg[0]), 16.0 * eps * sg[0]);
    while (keep < n && sg[keep] > thr) ++keep;
    if (tr.discarded_weight_cutoff > 0) {
      double total = 0;
      for (int i = 0; i < n; ++i) total += sg[i] * sg[i];
      double tail = 0;                                          // weight of sg[keep-1:] kept so far
      while (keep > 1) {
        tail += sg[keep - 1] * sg[keep - 1];
        if (tail / total < tr.discarded_weight_cutoff) --keep; else break;
      }
    }
    if (tr.max_extent > 0 && keep > tr.max_extent) keep = tr.max_extent;
    if (keep < 1) keep = 1;
    double sc = 1.0;
    if (tr.normalization == 1) { double s1 = 0; for (int i = 0; i < keep; ++i) s1 += sg[i]; sc = s1 > 0 ? 1.0 / s1 : 1.0; }
    else if (tr.normalization == 2) { double s2 = 0; for (int i = 0; i < keep; ++i) s2 += sg[i] * sg[i]; sc = s2 > 0 ? 1.0 / sqrt(s2) : 1.0; }
    else if (tr.normalization == 3) { sc = sg[0] > 0 ? 1.0 / sg[0] : 1.0; }
    rank[blockIdx.x] = keep;
    scale[blockIdx.x] = sc;
  }
}

// out1[c][:] = f1_c * W[perm[c]][:],  out2[c][:] = f2_c * W[perm[c]][:]   for c < k
template <typename R>
__global__ void __launch_bounds__(256)
emit_kernel(const cx<R> *__restrict__ W, int m, const int32_t *__restrict__ perm,
            const double *__restrict__ sigma, double scale, int partition, int flags,
            cx<R> *__restrict__ out1, cx<R> *__restrict__ out2) {
  const int transposed = flags & 1;
  const double cj = (flags & 2) ? -1.0 : 1.0;       // bit 1: out1 is conjugated
  using T = cx<R>;
  const int c = blockIdx.x;
  const double s = sigma[c], sn = s * scale;
  double fa = 1, fb = 1;                 // what the left / right factor absorbs
  if (partition == 1) fa = sn;
  else if (partition == 2) fb = sn;
  else if (partition == 3) { fa = sqrt(sn); fb = fa; }
  double f1 = 0, f2 = 0;
  if (s > 0) {
    if (!transposed) { f1 = fb / s; f2 = fa / (s * s); }
    else { f1 = fa / s; f2 = fb / (s * s); }
  }
  const T *src = W + (size_t)perm[c] * m;
  for (int col = threadIdx.x; col < m; col += 256) {
    const T v = src[col];
    T o1, o2;
    o1.x = (R)(f1 * v.x); o1.y = (R)(cj * f1 * v.y);
    o2.x = (R)(f2 * v.x); o2.y = (R)(f2 * v.y);
    out1[(size_t)c * m + col] = o1;
    out2[(size_t)c * m + col] = o2;
  }
}

// ---------------------------------------------------------------------------------------------
// LQ preconditioner (Drmac-Veselic style, without pivoting): W = L Q by right-looking modified
// Gram-Schmidt on the rows, panel by panel; only L is kept, as X = L^H (n x n, upper triangular).
// Row-Jacobi on X converges in ~3x fewer sweeps than on a strongly graded W (the rows of X are
// graded and nearly orthogonal), and theta = L Q = Z Sigma (J Q) has the same singular values and
// LEFT vectors Z: the caller takes the left factor from the rows of the converged X and the right
// factor from one GEMM with theta (the mirror image of the un-preconditioned path).
// MGS computes L as accurately as Householder QR would (Bjorck): what is lost is the orthogonality of
// Q, which is discarded.
// ---------------------------------------------------------------------------------------------
constexpr int kLqPanel = 16;
constexpr int kLqRowsPerCta = 32;

template <typename R>
__device__ __forceinline__ void lq_project(cx<R> *w, const cx<R> *q, int m, int lane, R &cr, R &ci) {
  // c = sum w * conj(q);  w -= c q
  R ar = 0, ai = 0;
  for (int c = lane; c < m; c += 32) {
    const cx<R> u = w[c], v = q[c];
    ar = fma(u.x, v.x, fma(u.y, v.y, ar));
    ai = fma(u.y, v.x, fma(-u.x, v.y, ai));
  }
  ar = warp_sum<R>(ar); ai = warp_sum<R>(ai);
  for (int c = lane; c < m; c += 32) {
    const cx<R> v = q[c];
    cx<R> u = w[c];
    u.x -= ar * v.x - ai * v.y;
    u.y -= ar * v.y + ai * v.x;
    w[c] = u;
  }
  cr = ar; ci = ai;
}

// One CTA per matrix: orthonormalise the rows of panel p among themselves (they were already
// projected against every earlier panel), record the diagonal block of L.
template <typename R>
__global__ void __launch_bounds__(256)
lq_panel_kernel(const qtn_svd_item *__restrict__ items, const qtn_svd_item *__restrict__ xitems, int p, int pb) {
  using T = cx<R>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *rows = reinterpret_cast<T *>(smem_raw);
  __shared__ R s_norm;
  const qtn_svd_item it = items[blockIdx.x];
  const int n = it.n, m = it.m;
  const int r0 = p * pb;
  if (r0 >= n) return;
  const int nr = min(pb, n - r0);
  T *W = (T *)it.w;
  T *X = (T *)xitems[blockIdx.x].w;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int lr = warp; lr < nr; lr += kWarps)
    for (int c = lane; c < m; c += 32) rows[(size_t)lr * m + c] = W[(size_t)(r0 + lr) * m + c];
  __syncthreads();
  for (int i = 0; i < nr; ++i) {
    T *qi = rows + (size_t)i * m;
    if (warp == 0) {
      R a = 0;
      for (int c = lane; c < m; c += 32) { const T v = qi[c]; a = fma(v.x, v.x, fma(v.y, v.y, a)); }
      a = warp_sum<R>(a);
      if (lane == 0) s_norm = sqrt(a);
    }
    __syncthreads();
    const R nrm = s_norm;
    const R inv = nrm > (R)0 ? (R)1 / nrm : (R)0;
    for (int c = t; c < m; c += 256) { T v = qi[c]; v.x *= inv; v.y *= inv; qi[c] = v; }
    if (t == 0) { T d; d.x = nrm; d.y = 0; X[(size_t)(r0 + i) * n + (r0 + i)] = d; }
    __syncthreads();
    for (int j = i + 1 + warp; j < nr; j += kWarps) {
      R cr, ci;
      lq_project<R>(rows + (size_t)j * m, qi, m, lane, cr, ci);
      if (lane == 0) { T d; d.x = cr; d.y = -ci; X[(size_t)(r0 + i) * n + (r0 + j)] = d; }   // X = L^H
    }
    __syncthreads();
  }
  for (int lr = warp; lr < nr; lr += kWarps)
    for (int c = lane; c < m; c += 32) W[(size_t)(r0 + lr) * m + c] = rows[(size_t)lr * m + c];
}

// Project the rows behind panel p against its (now orthonormal) rows, one after the other (true MGS
// order); a warp owns a trailing row, kept in shared memory while the 16 projections run.
template <typename R>
__global__ void __launch_bounds__(256)
lq_update_kernel(const qtn_svd_item *__restrict__ items, const qtn_svd_item *__restrict__ xitems, int p, int pb) {
  using T = cx<R>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const qtn_svd_item it = items[blockIdx.y];
  const int n = it.n, m = it.m;
  const int r0 = p * pb;
  const int t0 = r0 + pb;                                   // first trailing row
  const int first = t0 + (int)blockIdx.x * kLqRowsPerCta;
  if (r0 >= n || first >= n) return;
  T *q = reinterpret_cast<T *>(smem_raw);                   // pb x m
  T *mine = q + (size_t)pb * m;                             // kWarps x m
  T *W = (T *)it.w;
  T *X = (T *)xitems[blockIdx.y].w;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int lr = warp; lr < pb; lr += kWarps)
    for (int c = lane; c < m; c += 32) q[(size_t)lr * m + c] = W[(size_t)(r0 + lr) * m + c];
  __syncthreads();
  T *w = mine + (size_t)warp * m;
  const int last = min(n, first + kLqRowsPerCta);
  for (int j = first + warp; j < last; j += kWarps) {
    for (int c = lane; c < m; c += 32) w[c] = W[(size_t)j * m + c];
    __syncwarp();
    for (int i = 0; i < pb; ++i) {
      R cr, ci;
      lq_project<R>(w, q + (size_t)i * m, m, lane, cr, ci);
      if (lane == 0) { T d; d.x = cr; d.y = -ci; X[(size_t)(r0 + i) * n + j] = d; }
      __syncwarp();
    }
    for (int c = lane; c < m; c += 32) W[(size_t)j * m + c] = w[c];
    __syncwarp();
  }
}

// Register variant of lq_update_kernel for rows of at most 32 * CPL columns: the trailing row lives in the
// registers of its warp for the 16 projections, each panel row is read from shared memory once per projection
// (kept in registers between the dot product and the update): a third of the shared-memory traffic, and 16
// warps per CTA since the rows no longer need a shared-memory slot each.
// (8 warps when the two register rows need more than 64 registers, i.e. rows of 257..512 complex128 columns)
template <typename R, int CPL> constexpr int lq_warps_reg() { return CPL * (int)sizeof(cx<R>) / 4 * 2 > 64 ? 8 : 16; }
template <typename R, int CPL>
__global__ void __launch_bounds__(32 * lq_warps_reg<R, CPL>())
lq_update_reg_kernel(const qtn_svd_item *__restrict__ items, const qtn_svd_item *__restrict__ xitems, int p, int pb) {
  using T = cx<R>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const qtn_svd_item it = items[blockIdx.y];
  const int n = it.n, m = it.m;
  const int r0 = p * pb;
  const int t0 = r0 + pb;                                   // first trailing row
  const int first = t0 + (int)blockIdx.x * kLqRowsPerCta;
  if (r0 >= n || first >= n) return;
  constexpr int MP = 32 * CPL;
  constexpr int kLqWarpsReg = lq_warps_reg<R, CPL>();
  T *q = reinterpret_cast<T *>(smem_raw);                   // pb x MP, zero-padded
  T *W = (T *)it.w;
  T *X = (T *)xitems[blockIdx.y].w;
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  for (int lr = warp; lr < pb; lr += kLqWarpsReg)
    for (int c = lane; c < MP; c += 32) {
      T z; z.x = 0; z.y = 0;
      q[(size_t)lr * MP + c] = (r0 + lr < n && c < m) ? W[(size_t)(r0 + lr) * m + c] : z;
    }
  __syncthreads();
  const int last = min(n, first + kLqRowsPerCta);
  for (int j = first + warp; j < last; j += kLqWarpsReg) {
    T w[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int c = lane + 32 * k;
      T z; z.x = 0; z.y = 0;
      w[k] = c < m ? W[(size_t)j * m + c] : z;
    }
    for (int i = 0; i < pb; ++i) {
      const T *qi = q + (size_t)i * MP + lane;
      T v[CPL];
      R ar = 0, ai = 0, ar1 = 0, ai1 = 0;
#pragma unroll
      for (int k = 0; k < CPL; ++k) {                        // c = sum w * conj(q)
        v[k] = qi[32 * k];
        ar = fma(w[k].x, v[k].x, ar); ar1 = fma(w[k].y, v[k].y, ar1);
        ai = fma(w[k].y, v[k].x, ai); ai1 = fma(-w[k].x, v[k].y, ai1);
      }
      ar += ar1; ai += ai1;
      {
        const bool h16 = lane & 16;
        R kk = h16 ? ai : ar;
        const R sx = h16 ? ar : ai;
        kk += __shfl_xor_sync(0xffffffffu, sx, 16);
        kk += __shfl_xor_sync(0xffffffffu, kk, 8);
        kk += __shfl_xor_sync(0xffffffffu, kk, 4);
        kk += __shfl_xor_sync(0xffffffffu, kk, 2);
        kk += __shfl_xor_sync(0xffffffffu, kk, 1);
        ar = __shfl_sync(0xffffffffu, kk, 0);
        ai = __shfl_sync(0xffffffffu, kk, 16);
      }
#pragma unroll
      for (int k = 0; k < CPL; ++k) {                        // w -= c q
        w[k].x -= ar * v[k].x - ai * v[k].y;
        w[k].y -= ar * v[k].y + ai * v[k].x;
      }
      if (lane == 0) { T d; d.x = ar; d.y = -ai; X[(size_t)(r0 + i) * n + j] = d; }
    }
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int c = lane + 32 * k;
      if (c < m) W[(size_t)j * m + c] = w[k];
    }
  }
}

template <typename R, int CPL>
cudaError_t launch_lq_update_reg(dim3 grid, cudaStream_t st, const qtn_svd_item *d_items, const qtn_svd_item *d_x,
                                 int p, int pb) {
  const size_t smem = (size_t)pb * 32 * CPL * sizeof(cx<R>);
  static size_t set = 0;
  if (smem > set) {
    cudaError_t e = cudaFuncSetAttribute(lq_update_reg_kernel<R, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    set = smem;
  }
  lq_update_reg_kernel<R, CPL><<<grid, 32 * lq_warps_reg<R, CPL>(), smem, st>>>(d_items, d_x, p, pb);
  return cudaGetLastError();
}

template <typename R>
int svd_lq_impl(int batch, const qtn_svd_item *items, const qtn_svd_item *xitems, unsigned char *ws, cudaStream_t st) {
  const int esz = (int)sizeof(cx<R>);
  int max_n = 1, max_m = 1;
  for (int i = 0; i < batch; ++i) {
    if (items[i].n < 1 || items[i].m < 1 || !items[i].w || !xitems[i].w || xitems[i].n != items[i].n ||
        xitems[i].m != items[i].n)
      return qtn_fail(QTN_ERR_ARG, "qtn_svd_lq: bad item");
    max_n = std::max(max_n, (int)items[i].n);
    max_m = std::max(max_m, (int)items[i].m);
  }
  int pb = kLqPanel;
  while (pb > 1 && (size_t)(pb + kWarps) * max_m * esz > (size_t)200 * 1024) pb >>= 1;
  if ((size_t)(pb + kWarps) * max_m * esz > (size_t)200 * 1024)
    return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_svd_lq: rows too long for shared memory");
  qtn_svd_item *d_items = (qtn_svd_item *)ws;
  qtn_svd_item *d_x = d_items + batch;
  QTN_CUDA_CHECK(cudaMemcpyAsync(d_items, items, (size_t)batch * sizeof(qtn_svd_item), cudaMemcpyHostToDevice, st));
  QTN_CUDA_CHECK(cudaMemcpyAsync(d_x, xitems, (size_t)batch * sizeof(qtn_svd_item), cudaMemcpyHostToDevice, st));
  for (int i = 0; i < batch; ++i)
    QTN_CUDA_CHECK(cudaMemsetAsync(xitems[i].w, 0, (size_t)items[i].n * items[i].n * esz, st));
  const size_t smem_panel = (size_t)pb * max_m * esz, smem_upd = (size_t)(pb + kWarps) * max_m * esz;
  static size_t set_panel = 0, set_upd = 0;
  if (smem_panel > set_panel) {
    QTN_CUDA_CHECK(cudaFuncSetAttribute(lq_panel_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_panel));
    set_panel = smem_panel;
  }
  if (smem_upd > set_upd) {
    QTN_CUDA_CHECK(cudaFuncSetAttribute(lq_update_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_upd));
    set_upd = smem_upd;
  }
  const int panels = (max_n + pb - 1) / pb;
  for (int p = 0; p < panels; ++p) {
    lq_panel_kernel<R><<<batch, 256, smem_panel, st>>>(d_items, d_x, p, pb);
    QTN_LAUNCH_CHECK();
    const int trailing = max_n - (p + 1) * pb;
    if (trailing > 0) {
      dim3 grid((unsigned)((trailing + kLqRowsPerCta - 1) / kLqRowsPerCta), (unsigned)batch);
      // rows of <= 512 (complex128) / 1024 (complex64) columns: register variant (64 registers of row data)
      constexpr int kCplMax = sizeof(R) == 8 ? 16 : 32;
      if (max_m <= 32 * kCplMax) {
        cudaError_t e = max_m <= 16 * kCplMax ? launch_lq_update_reg<R, kCplMax / 2>(grid, st, d_items, d_x, p, pb)
                                              : launch_lq_update_reg<R, kCplMax>(grid, st, d_items, d_x, p, pb);
        if (e != cudaSuccess) return qtn_fail(QTN_ERR_CUDA, cudaGetErrorString(e));
      } else {
        lq_update_kernel<R><<<grid, 256, smem_upd, st>>>(d_items, d_x, p, pb);
        QTN_LAUNCH_CHECK();
      }
    }
  }
  return QTN_OK;
}

int choose_block(int n, int m, int esz, int budget, int *block, int *nblocks) {
  int b = 32;
  while (b >= 1 && (long long)2 * b * m * esz > budget) b >>= 1;
  if (b < 1) return -1;
  if (n <= 2 * b) {                       // the whole matrix is one resident block
    int p = 2;
    while (p < n) p <<= 1;
    *block = p; *nblocks = 1;
    return 0;
  }
  *block = b; *nblocks = (n + b - 1) / b;
  return 0;
}

template <typename R, int CPL, int WPP>
cudaError_t launch_cross_reg(dim3 grid, cudaStream_t st, const qtn_svd_item *d_items, int step, int sweep, R tol,
                             const int *rot_prev, int *rot_cur, int *changed, long long ld, int gstep,
                             int steps_per_sweep, int reg_cap) {
  constexpr size_t smem = (size_t)kRegBlock * 32 * WPP * CPL * sizeof(cx<R>);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(jacobi_cross_reg_kernel<R, CPL, WPP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  jacobi_cross_reg_kernel<R, CPL, WPP><<<grid, 32 * kRegBlock * WPP, smem, st>>>(
      d_items, step, sweep, tol, rot_prev, rot_cur, changed, ld, gstep, steps_per_sweep, reg_cap);
  return cudaGetLastError();
}

template <typename R, int CPL, int WPP>
cudaError_t launch_cross_dual(dim3 grid, cudaStream_t st, const qtn_svd_item *d_items, int step, int sweep, R tol,
                              const int *rot_prev, int *rot_cur, int *changed, long long ld, int gstep,
                              int steps_per_sweep, int reg_cap) {
  constexpr size_t smem = (size_t)kRegBlock * 32 * WPP * CPL * sizeof(cx<R>);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(jacobi_cross_dual_kernel<R, CPL, WPP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  jacobi_cross_dual_kernel<R, CPL, WPP><<<grid, 32 * kRegBlock * WPP, smem, st>>>(
      d_items, step, sweep, tol, rot_prev, rot_cur, changed, ld, gstep, steps_per_sweep, reg_cap);
  return cudaGetLastError();
}

template <typename R, int CPL, int WPP>
cudaError_t launch_cross_stream(int pairs_max, int batch, cudaStream_t st, const qtn_svd_item *d_items, int step,
                                int sweep, R tol, const int *rot_prev, int *rot_cur, int *changed, long long ld,
                                int gstep, int steps_per_sweep, int reg_cap) {
  constexpr size_t smem = (size_t)3 * kRegBlock * 32 * WPP * CPL * sizeof(cx<R>);
  static int sms = 0;
  if (!sms) {
    cudaError_t e = cudaFuncSetAttribute(jacobi_cross_stream_kernel<R, CPL, WPP>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  }
  const int grid = std::max(1, std::min(sms, pairs_max * batch));
  jacobi_cross_stream_kernel<R, CPL, WPP><<<grid, 32 * kRegBlock * WPP, smem, st>>>(
      d_items, step, sweep, tol, rot_prev, rot_cur, changed, ld, gstep, steps_per_sweep, reg_cap, pairs_max, batch);
  return cudaGetLastError();
}

// register-kernel capacity (columns) for a dtype: 64 registers of row data per thread at most
template <typename R> constexpr int reg_cap_of() { return sizeof(R) == 8 ? 512 : 1024; }

template <typename R, int NW>
int svd_rows_impl(int batch, qtn_svd_item *items, double tol, int max_sweeps, const qtn_svd_trunc *trunc,
                  double *sigma, int32_t *perm, int64_t ld, int32_t *rank_host, double *scale_host,
                  int32_t *sweeps_host, unsigned char *ws, cudaStream_t st, int reg_variant) {
  const int esz = (int)sizeof(cx<R>);
  constexpr int kBudget = NW > kWarps ? kSmemBudgetWide : kSmemBudget;
  const int reg_cap = reg_variant > 0 ? reg_cap_of<R>() : 0;
  int max_steps = 1, max_pairs = 1, reg_max_m = 0, reg_cross_steps = 0, reg_pairs = 1;
  size_t smem = 0;
  std::vector<char> need_smem_kernel;                     // per step: does jacobi_step_kernel have work?
  for (int i = 0; i < batch; ++i) {
    qtn_svd_item &it = items[i];
    if (it.n < 1 || it.m < 1 || !it.w) return qtn_fail(QTN_ERR_ARG, "qtn_svd_rows: bad item");
    if (it.n > ld) return qtn_fail(QTN_ERR_ARG, "qtn_svd_rows: ld smaller than a row count");
    if (choose_block(it.n, it.m, esz, kBudget, &it.block, &it.nblocks))
      return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_svd_rows: rows too long for shared memory");
    if (reg_cap > 0 && it.nblocks > 1 && it.m <= reg_cap) {
      it.block = kRegBlock;
      it.nblocks = (it.n + kRegBlock - 1) / kRegBlock;
    }
    const int nbe = it.nblocks + (it.nblocks & 1);
    const int nsteps = it.nblocks == 1 ? 1 : nbe;
    max_steps = std::max(max_steps, nsteps);
    if ((int)need_smem_kernel.size() < nsteps) need_smem_kernel.resize(nsteps, 0);
    const bool reg = reg_mode(it, reg_cap);
    if (reg) {
      reg_max_m = std::max(reg_max_m, (int)it.m);
      reg_cross_steps = std::max(reg_cross_steps, nbe - 1);
      reg_pairs = std::max(reg_pairs, nbe / 2);
      need_smem_kernel[nbe - 1] = 1;
      max_pairs = std::max(max_pairs, (int)it.nblocks);
      smem = std::max(smem, (size_t)it.block * it.m * esz);
    } else {
      for (int sidx = 0; sidx < nsteps; ++sidx) need_smem_kernel[sidx] = 1;
      max_pairs = std::max(max_pairs, it.nblocks == 1 ? 1 : std::max(nbe / 2, it.nblocks));
      const int nr = it.nblocks == 1 ? it.block : 2 * it.block;
      smem = std::max(smem, (size_t)nr * it.m * esz);
    }
  }
  // workspace: items | (unused) | rot[2] | rank | scale | scratch sigma | last-rotation step per block
  qtn_svd_item *d_items = (qtn_svd_item *)ws;
  size_t off = ((size_t)batch * sizeof(qtn_svd_item) + 255) & ~(size_t)255;
  off += ((size_t)batch * 8 + 255) & ~(size_t)255;
  int *d_rot = (int *)(ws + off); off += ((size_t)batch * 8 + 255) & ~(size_t)255;
  int32_t *d_rank = (int32_t *)(ws + off); off += ((size_t)batch * 4 + 255) & ~(size_t)255;
  double *d_scale = (double *)(ws + off); off += ((size_t)batch * 8 + 255) & ~(size_t)255;
  double *d_scratch = (double *)(ws + off); off += ((size_t)batch * ld * 8 + 255) & ~(size_t)255;
  int *d_changed = (int *)(ws + off);                         // [batch][ld]: step of the last rotation per block
  QTN_CUDA_CHECK(cudaMemsetAsync(d_changed, 0x7f, (size_t)batch * ld * sizeof(int), st));
  QTN_CUDA_CHECK(cudaMemcpyAsync(d_items, items, (size_t)batch * sizeof(qtn_svd_item), cudaMemcpyHostToDevice, st));
  static size_t smem_set = 0;
  if (smem > smem_set) {
    QTN_CUDA_CHECK(cudaFuncSetAttribute(jacobi_step_kernel<R, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)std::max(smem, (size_t)kBudget)));
    smem_set = std::max(smem, (size_t)kBudget);
  }
  // Convergence is read back ONE SWEEP LATE: the counters of sweep s travel to pinned host memory behind an event
  // while sweep s + 1 is already being launched (its kernels return at once for matrices whose previous sweep made
  // no rotation), so the device never waits for the host between sweeps.
  static thread_local int *h_rot = nullptr;
  static thread_local size_t h_cap = 0;
  static thread_local cudaEvent_t ev[2] = {nullptr, nullptr};
  if (h_cap < (size_t)2 * batch) {
    if (h_rot) cudaFreeHost(h_rot);
    h_cap = (size_t)2 * batch + 64;
    QTN_CUDA_CHECK(cudaHostAlloc((void **)&h_rot, h_cap * sizeof(int), cudaHostAllocDefault));
  }
  for (int e = 0; e < 2; ++e)
    if (!ev[e]) QTN_CUDA_CHECK(cudaEventCreateWithFlags(&ev[e], cudaEventDisableTiming));
  auto clean = [&](int sw) {                     // did sweep `sw` leave every matrix without a rotation?
    const int *r = h_rot + (size_t)(sw & 1) * batch;
    for (int i = 0; i < batch; ++i)
      if (r[i] != 0) return false;
    return true;
  };
  int sweep = 0;
  bool converged = false;
  for (; sweep < max_sweeps; ++sweep) {
    int *rot_cur = d_rot + (sweep & 1) * batch, *rot_prev = d_rot + ((sweep & 1) ^ 1) * batch;
    QTN_CUDA_CHECK(cudaMemsetAsync(rot_cur, 0, (size_t)batch * sizeof(int), st));
    dim3 grid((unsigned)max_pairs, (unsigned)batch);
    dim3 rgrid((unsigned)reg_pairs, (unsigned)batch);
    for (int step = 0; step < max_steps; ++step) {
      const int gstep = sweep * max_steps + step + 1;
      if (step < reg_cross_steps) {
        cudaError_t e;
#define QTN_REG(CPL, WPP)                                                                                    \
  launch_cross_reg<R, CPL, WPP>(rgrid, st, d_items, step, sweep, (R)tol, rot_prev, rot_cur, d_changed,        \
                                (long long)ld, gstep, max_steps, reg_cap)
#define QTN_STREAM(CPL, WPP)                                                                                 \
  launch_cross_stream<R, CPL, WPP>(reg_pairs, batch, st, d_items, step, sweep, (R)tol, rot_prev, rot_cur,     \
                                   d_changed, (long long)ld, gstep, max_steps, reg_cap)
#define QTN_DUAL(CPL, WPP)                                                                                   \
  launch_cross_dual<R, CPL, WPP>(rgrid, st, d_items, step, sweep, (R)tol, rot_prev, rot_cur, d_changed,       \
                                 (long long)ld, gstep, max_steps, reg_cap)
        if constexpr (sizeof(R) == 8) {
          if (reg_variant == 5) e = reg_max_m <= 256 ? QTN_DUAL(8, 1) : QTN_DUAL(16, 1);
          else if (reg_variant == 4) e = reg_max_m <= 256 ? QTN_DUAL(4, 2) : QTN_DUAL(8, 2);
          else if (reg_variant == 3) e = reg_max_m <= 256 ? QTN_STREAM(4, 2) : QTN_STREAM(8, 2);
          else if (reg_variant == 2) e = reg_max_m <= 256 ? QTN_REG(2, 4) : QTN_REG(4, 4);
          else e = reg_max_m <= 256 ? QTN_REG(4, 2) : QTN_REG(8, 2);
        } else {
          if (reg_variant == 5) e = reg_max_m <= 512 ? QTN_DUAL(16, 1) : QTN_DUAL(32, 1);
          else if (reg_variant == 4) e = reg_max_m <= 512 ? QTN_DUAL(8, 2) : QTN_DUAL(16, 2);
          else if (reg_variant == 3) e = reg_max_m <= 512 ? QTN_STREAM(8, 2) : QTN_STREAM(16, 2);
          else if (reg_variant == 2) e = reg_max_m <= 512 ? QTN_REG(4, 4) : QTN_REG(8, 4);
          else e = reg_max_m <= 512 ? QTN_REG(8, 2) : QTN_REG(16, 2);
        }
#undef QTN_REG
#undef QTN_STREAM
#undef QTN_DUAL
        if (e != cudaSuccess) return qtn_fail(QTN_ERR_CUDA, cudaGetErrorString(e));
      }
      if (need_smem_kernel[step]) {
        jacobi_step_kernel<R, NW><<<grid, 32 * NW, smem, st>>>(d_items, step, sweep, max_sweeps, (R)tol, rot_prev,
                                                               rot_cur, d_changed, (long long)ld, gstep, max_steps,
                                                               reg_cap);
        QTN_LAUNCH_CHECK();
      }
    }
    QTN_CUDA_CHECK(cudaMemcpyAsync(h_rot + (size_t)(sweep & 1) * batch, rot_cur, (size_t)batch * sizeof(int),
                                   cudaMemcpyDeviceToHost, st));
    QTN_CUDA_CHECK(cudaEventRecord(ev[sweep & 1], st));
    if (sweep >= 1) {
      QTN_CUDA_CHECK(cudaEventSynchronize(ev[(sweep - 1) & 1]));
      if (clean(sweep - 1)) { converged = true; break; }      // sweep `sweep` was launched for nothing: all early exits
    }
  }
  if (!converged) {                              // ran out of sweeps (or a single sweep): the last one may be the clean one
    const int last = sweep - 1;
    QTN_CUDA_CHECK(cudaEventSynchronize(ev[last & 1]));
    if (clean(last)) converged = true;
  }
  // sweeps executed, the clean one included (with the one-sweep lag `sweep` already is that count on a break)
  if (sweeps_host) *sweeps_host = sweep;
  finalize_kernel<R><<<batch, 256, 0, st>>>(d_items, *trunc, d_scratch, sigma, perm, ld, d_rank, d_scale);
  QTN_LAUNCH_CHECK();
  QTN_CUDA_CHECK(cudaMemcpyAsync(rank_host, d_rank, (size_t)batch * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  QTN_CUDA_CHECK(cudaMemcpyAsync(scale_host, d_scale, (size_t)batch * sizeof(double), cudaMemcpyDeviceToHost, st));
  QTN_CUDA_CHECK(cudaStreamSynchronize(st));
  return QTN_OK;
}

}  // namespace

extern "C" int64_t qtn_svd_workspace_bytes(int batch, int64_t ld) {
  if (batch < 1 || ld < 1) return 0;
  int64_t off = ((int64_t)batch * (int64_t)sizeof(qtn_svd_item) + 255) & ~(int64_t)255;
  off += 4 * (((int64_t)batch * 8 + 255) & ~(int64_t)255);
  off += (((int64_t)batch * ld * 8 + 255) & ~(int64_t)255) + (int64_t)batch * ld * 4 + 512;
  return off;
}

extern "C" int qtn_svd_rows(int dtype, int batch, qtn_svd_item *items, double tol, int max_sweeps,
                            const qtn_svd_trunc *trunc, double *sigma, int32_t *perm, int64_t ld,
                            int32_t *rank_host, double *scale_host, int32_t *sweeps_host, void *workspace,
                            int64_t workspace_bytes, void *stream) {
  return qtn_svd_rows_opts(dtype, batch, items, tol, max_sweeps, trunc, sigma, perm, ld, rank_host, scale_host,
                           sweeps_host, workspace, workspace_bytes, nullptr, stream);
}

extern "C" int qtn_svd_rows_opts(int dtype, int batch, qtn_svd_item *items, double tol, int max_sweeps,
                                 const qtn_svd_trunc *trunc, double *sigma, int32_t *perm, int64_t ld,
                                 int32_t *rank_host, double *scale_host, int32_t *sweeps_host, void *workspace,
                                 int64_t workspace_bytes, const qtn_options *given, void *stream) {
  qtn_options opt;
  qtn_options_resolve(given, &opt);
  if (dtype != QTN_C64 && dtype != QTN_C128) return qtn_fail(QTN_ERR_ARG, "qtn_svd_rows: bad dtype");
  if (batch < 1 || !items || !trunc || !sigma || !perm || !rank_host || !scale_host || !workspace)
    return qtn_fail(QTN_ERR_ARG, "qtn_svd_rows: null argument");
  if (workspace_bytes < qtn_svd_workspace_bytes(batch, ld))
    return qtn_fail(QTN_ERR_WORKSPACE, "qtn_svd_rows: workspace too small (qtn_svd_workspace_bytes)");
  if (max_sweeps < 1) max_sweeps = 30;
  if (!(tol > 0)) {            // LAPACK xGESVJ-style: sqrt(row length) * machine epsilon (x8)
    int max_m = 1;
    for (int i = 0; i < batch; ++i) max_m = std::max(max_m, (int)items[i].m);
    tol = 8.0 * std::sqrt((double)max_m) * (dtype == QTN_C64 ? 6.0e-8 : 1.1e-16);
  }
  cudaStream_t st = (cudaStream_t)stream;
  // "svd_group_mb": matrices are swept in groups whose rows fit that many MiB, so that a group stays in L2 over
  // the ~n/b launches of a sweep (0 = the whole batch at once)
  const int64_t esz = dtype == QTN_C64 ? 8 : 16;
  const int64_t group_bytes = (int64_t)opt.svd_group_mb << 20;
  const bool wide = opt.svd_wide != 0;
  const int reg_variant = opt.svd_reg;
  int sweeps_max = 0;
  for (int g0 = 0; g0 < batch;) {
    int g1 = g0;
    int64_t bytes = 0;
    while (g1 < batch) {
      const int64_t add = (int64_t)items[g1].n * items[g1].m * esz;
      if (g1 > g0 && group_bytes > 0 && bytes + add > group_bytes) break;
      bytes += add; ++g1;
    }
    int32_t sw = 0;
    const int nb = g1 - g0;
    int rc;
#define QTN_SVD_CALL(R, NW)                                                                                      \
  svd_rows_impl<R, NW>(nb, items + g0, tol, max_sweeps, trunc, sigma + (int64_t)g0 * ld, perm + (int64_t)g0 * ld, \
                       ld, rank_host + g0, scale_host + g0, &sw, (unsigned char *)workspace, st, reg_variant)
    if (dtype == QTN_C64) rc = wide ? QTN_SVD_CALL(float, kWarpsWide) : QTN_SVD_CALL(float, kWarps);
    else rc = wide ? QTN_SVD_CALL(double, kWarpsWide) : QTN_SVD_CALL(double, kWarps);
#undef QTN_SVD_CALL
    if (rc != QTN_OK) return rc;
    sweeps_max = std::max(sweeps_max, (int)sw);
    g0 = g1;
  }
  if (sweeps_host) *sweeps_host = sweeps_max;
  return QTN_OK;
}

extern "C" int64_t qtn_svd_lq_workspace_bytes(int batch) {
  return batch < 1 ? 0 : 2 * (int64_t)batch * (int64_t)sizeof(qtn_svd_item) + 256;
}

extern "C" int qtn_svd_lq(int dtype, int batch, const qtn_svd_item *items, const qtn_svd_item *xitems,
                          void *workspace, int64_t workspace_bytes, void *stream) {
  if (dtype != QTN_C64 && dtype != QTN_C128) return qtn_fail(QTN_ERR_ARG, "qtn_svd_lq: bad dtype");
  if (batch < 1 || !items || !xitems || !workspace) return qtn_fail(QTN_ERR_ARG, "qtn_svd_lq: null argument");
  if (workspace_bytes < qtn_svd_lq_workspace_bytes(batch))
    return qtn_fail(QTN_ERR_WORKSPACE, "qtn_svd_lq: workspace too small (qtn_svd_lq_workspace_bytes)");
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == QTN_C64 ? svd_lq_impl<float>(batch, items, xitems, (unsigned char *)workspace, st)
                          : svd_lq_impl<double>(batch, items, xitems, (unsigned char *)workspace, st);
}

extern "C" int qtn_svd_emit(int dtype, const void *w, int32_t m, const int32_t *perm, const double *sigma,
                            double scale, int32_t k, int32_t partition, int32_t transposed, void *out1,
                            void *out2, void *stream) {
  if (dtype != QTN_C64 && dtype != QTN_C128) return qtn_fail(QTN_ERR_ARG, "qtn_svd_emit: bad dtype");
  if (!w || !perm || !sigma || !out1 || !out2 || m < 1 || k < 1 || partition < 0 || partition > 3 || transposed < 0 || transposed > 3)
    return qtn_fail(QTN_ERR_ARG, "qtn_svd_emit: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == QTN_C64)
    emit_kernel<float><<<k, 256, 0, st>>>((const float2 *)w, m, perm, sigma, scale, partition, transposed,
                                          (float2 *)out1, (float2 *)out2);
  else
    emit_kernel<double><<<k, 256, 0, st>>>((const double2 *)w, m, perm, sigma, scale, partition, transposed,
                                           (double2 *)out1, (double2 *)out2);
  QTN_LAUNCH_CHECK();
  return QTN_OK;
}
