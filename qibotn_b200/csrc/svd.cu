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

template <typename R>
__global__ void __launch_bounds__(256)
jacobi_step_kernel(const qtn_svd_item *__restrict__ items, int step, int sweep, int max_inner, R tol,
                   const int *__restrict__ rot_prev, int *__restrict__ rot_cur, int *__restrict__ changed,
                   long long ld, int gstep, int steps_per_sweep) {
  using T = cx<R>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *rows = reinterpret_cast<T *>(smem_raw);
  const int item_id = blockIdx.y;
  if (sweep > 0 && rot_prev[item_id] == 0) return;          // this matrix has converged
  const qtn_svd_item it = items[item_id];
  const int n = it.n, m = it.m, b = it.block, nb = it.nblocks;
  if (n < 2) return;
  const int nbe = nb + (nb & 1);
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
  for (int lr = warp; lr < nr; lr += kWarps) {
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
    for (int round = 0; round < nrounds; ++round) {
      for (int q = warp; q < npairs; q += kWarps) {
        int x, y;
        if (J < 0) circle(round, q, nr, x, y);                  // all pairs inside one block
        else { x = q; y = b + ((q + round) & (b - 1)); }        // row q of I with row (q+round)%b of J
        if (grow(x) >= n || grow(y) >= n) continue;
        T *X = rows + (size_t)x * m, *Y = rows + (size_t)y * m;
        R a = 0, bb = 0, gr = 0, gi = 0;
        for (int c = lane; c < m; c += 32) {
          const T u = X[c], v = Y[c];
          a = fma(u.x, u.x, fma(u.y, u.y, a));
          bb = fma(v.x, v.x, fma(v.y, v.y, bb));
          gr = fma(u.x, v.x, fma(u.y, v.y, gr));               // g = sum u * conj(v)
          gi = fma(u.y, v.x, fma(-u.x, v.y, gi));
        }
        a = warp_sum<R>(a); bb = warp_sum<R>(bb); gr = warp_sum<R>(gr); gi = warp_sum<R>(gi);
        const R g2 = gr * gr + gi * gi;
        // relative criterion at every scale (rows far below sigma_max are orthogonalised as
        // accurately as the leading ones); a row whose squared norm underflowed is left alone
        if (!(g2 > tol * tol * a * bb) || !(fmin(a, bb) > (R)0)) continue;
        rotated = 1;
        const double gabs = sqrt((double)g2);
        const double zeta = ((double)bb - (double)a) / (2.0 * gabs);
        const double tt = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
        const double cd = 1.0 / sqrt(1.0 + tt * tt), sd = cd * tt;
        const R c = (R)cd;
        const R spr = (R)(sd * gr / gabs), spi = (R)(sd * gi / gabs);   // s * e^{i phi}
        for (int col = lane; col < m; col += 32) {
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
  for (int lr = warp; lr < nr; lr += kWarps) {
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
      lq_update_kernel<R><<<grid, 256, smem_upd, st>>>(d_items, d_x, p, pb);
      QTN_LAUNCH_CHECK();
    }
  }
  return QTN_OK;
}

int choose_block(int n, int m, int esz, int *block, int *nblocks) {
  int b = 32;
  while (b >= 1 && (long long)2 * b * m * esz > kSmemBudget) b >>= 1;
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

template <typename R>
int svd_rows_impl(int batch, qtn_svd_item *items, double tol, int max_sweeps, const qtn_svd_trunc *trunc,
                  double *sigma, int32_t *perm, int64_t ld, int32_t *rank_host, double *scale_host,
                  int32_t *sweeps_host, unsigned char *ws, cudaStream_t st) {
  const int esz = (int)sizeof(cx<R>);
  int max_steps = 1, max_pairs = 1;
  size_t smem = 0;
  for (int i = 0; i < batch; ++i) {
    qtn_svd_item &it = items[i];
    if (it.n < 1 || it.m < 1 || !it.w) return qtn_fail(QTN_ERR_ARG, "qtn_svd_rows: bad item");
    if (it.n > ld) return qtn_fail(QTN_ERR_ARG, "qtn_svd_rows: ld smaller than a row count");
    if (choose_block(it.n, it.m, esz, &it.block, &it.nblocks))
      return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_svd_rows: rows too long for shared memory");
    const int nbe = it.nblocks + (it.nblocks & 1);
    max_steps = std::max(max_steps, it.nblocks == 1 ? 1 : nbe);
    max_pairs = std::max(max_pairs, it.nblocks == 1 ? 1 : std::max(nbe / 2, it.nblocks));
    const int nr = it.nblocks == 1 ? it.block : 2 * it.block;
    smem = std::max(smem, (size_t)nr * it.m * esz);
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
    QTN_CUDA_CHECK(cudaFuncSetAttribute(jacobi_step_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)std::max(smem, (size_t)kSmemBudget)));
    smem_set = std::max(smem, (size_t)kSmemBudget);
  }
  std::vector<int> rot(batch);
  int sweep = 0;
  for (; sweep < max_sweeps; ++sweep) {
    int *rot_cur = d_rot + (sweep & 1) * batch, *rot_prev = d_rot + ((sweep & 1) ^ 1) * batch;
    QTN_CUDA_CHECK(cudaMemsetAsync(rot_cur, 0, (size_t)batch * sizeof(int), st));
    dim3 grid((unsigned)max_pairs, (unsigned)batch);
    for (int step = 0; step < max_steps; ++step) {
      jacobi_step_kernel<R><<<grid, 256, smem, st>>>(d_items, step, sweep, max_sweeps, (R)tol, rot_prev, rot_cur,
                                                    d_changed, (long long)ld, sweep * max_steps + step + 1,
                                                    max_steps);
      QTN_LAUNCH_CHECK();
    }
    QTN_CUDA_CHECK(cudaMemcpyAsync(rot.data(), rot_cur, (size_t)batch * sizeof(int), cudaMemcpyDeviceToHost, st));
    QTN_CUDA_CHECK(cudaStreamSynchronize(st));
    bool any = false;
    for (int i = 0; i < batch; ++i) any = any || rot[i] != 0;
    if (!any) { ++sweep; break; }
  }
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
  return dtype == QTN_C64
             ? svd_rows_impl<float>(batch, items, tol, max_sweeps, trunc, sigma, perm, ld, rank_host, scale_host,
                                    sweeps_host, (unsigned char *)workspace, st)
             : svd_rows_impl<double>(batch, items, tol, max_sweeps, trunc, sigma, perm, ld, rank_host, scale_host,
                                     sweeps_host, (unsigned char *)workspace, st);
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
