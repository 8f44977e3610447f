// Dense complex building blocks of the MPS path (arbitrary extents, row-major):
//   qtn_gemm          C = op(A) * op(B) (+ C)     -- A.B two-site contraction, transfer-matrix
//                                                    sweeps, dense read-out, U = theta * V / S
//   qtn_mps_apply_1q  "ipj,qp->iqj" in place       (reference mps_utils.py:64-69)
//   qtn_mps_apply_2q  gate on the (p,q) legs of a two-site block stored as the (2l x 2r)
//                     matrix [(i,p),(q,k)]         (the gate part of mps_utils.py:78-84; with the
//                                                    SWAP gate it is the exchange of mps_utils.py:32-37)
// The GEMM is a shared-memory tiled FFMA/DFMA kernel (64x64x16 tile, 4x4 outputs per thread);
// MPS bonds are not powers of two, so the bit-tensor kernels of gett*.cu do not apply here.
#include <algorithm>
#include <cstdint>

#include "qtn_internal.h"

namespace {

template <typename R> struct cx_of;
template <> struct cx_of<float> { using type = float2; };
template <> struct cx_of<double> { using type = double2; };
template <typename R> using cx = typename cx_of<R>::type;

template <typename R>
__device__ __forceinline__ void cfma(cx<R> &acc, const cx<R> &a, const cx<R> &b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}

constexpr int GT = 64;   // tile edge
constexpr int GK = 16;   // k step

// op: bit 0 = conjugate, bit 1 = transpose (stored matrix is the transpose of the operand)
template <typename R>
__device__ __forceinline__ void gemm_tile(int opa, int opb, long long M, long long N, long long K,
                                          const cx<R> *__restrict__ A, long long lda, const cx<R> *__restrict__ B,
                                          long long ldb, cx<R> *__restrict__ C, long long ldc, int accumulate,
                                          long long m0, long long n0) {
  using T = cx<R>;
  __shared__ T sA[GK][GT + 1];
  __shared__ T sB[GK][GT + 1];
  const int t = threadIdx.x;
  const int tm = (t >> 4) * 4, tn = (t & 15) * 4;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) { acc[i][j].x = 0; acc[i][j].y = 0; }
  const bool ta = opa & 2, tb = opb & 2;
  const R ca = (opa & 1) ? R(-1) : R(1), cb = (opb & 1) ? R(-1) : R(1);
  for (long long k0 = 0; k0 < K; k0 += GK) {
    // A tile: GT x GK values, 4 per thread; the fastest thread index follows the stored
    // matrix's contiguous dimension
    for (int e = t; e < GT * GK; e += 256) {
      int mm, kk;
      if (ta) { mm = e % GT; kk = e / GT; } else { kk = e % GK; mm = e / GK; }
      const long long gm = m0 + mm, gk = k0 + kk;
      T v; v.x = 0; v.y = 0;
      if (gm < M && gk < K) { v = ta ? A[gk * lda + gm] : A[gm * lda + gk]; v.y *= ca; }
      sA[kk][mm] = v;
    }
    for (int e = t; e < GT * GK; e += 256) {
      int nn, kk;
      if (tb) { kk = e % GK; nn = e / GK; } else { nn = e % GT; kk = e / GT; }
      const long long gn = n0 + nn, gk = k0 + kk;
      T v; v.x = 0; v.y = 0;
      if (gn < N && gk < K) { v = tb ? B[gn * ldb + gk] : B[gk * ldb + gn]; v.y *= cb; }
      sB[kk][nn] = v;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < GK; ++kk) {
      T a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = sA[kk][tm + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = sB[kk][tn + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) cfma<R>(acc[i][j], a[i], b[j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const long long gm = m0 + tm + i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const long long gn = n0 + tn + j;
      if (gn >= N) continue;
      T v = acc[i][j];
      if (accumulate) { const T o = C[gm * ldc + gn]; v.x += o.x; v.y += o.y; }
      C[gm * ldc + gn] = v;
    }
  }
}

template <typename R>
__global__ void __launch_bounds__(256)
gemm_kernel(int opa, int opb, long long M, long long N, long long K, const cx<R> *__restrict__ A,
            long long lda, const cx<R> *__restrict__ B, long long ldb, cx<R> *__restrict__ C,
            long long ldc, int accumulate) {
  gemm_tile<R>(opa, opb, M, N, K, A, lda, B, ldb, C, ldc, accumulate, (long long)blockIdx.y * GT,
               (long long)blockIdx.x * GT);
}

// Up to kGemmBatch independent products in one launch (blockIdx.z = product): the 62 small GEMMs of an MPS
// layer fill 64 of 148 SMs each when launched one by one.
constexpr int kGemmBatch = 32;
struct GemmBatch {
  int count;
  qtn_gemm_item it[kGemmBatch];
};

template <typename R>
__global__ void __launch_bounds__(256)
gemm_batched_kernel(const __grid_constant__ GemmBatch batch, int accumulate) {
  const qtn_gemm_item &g = batch.it[blockIdx.z];
  const long long m0 = (long long)blockIdx.y * GT, n0 = (long long)blockIdx.x * GT;
  if (m0 >= g.m || n0 >= g.n) return;
  gemm_tile<R>(g.opa, g.opb, g.m, g.n, g.k, (const cx<R> *)g.a, g.lda, (const cx<R> *)g.b, g.ldb, (cx<R> *)g.c,
               g.ldc, accumulate, m0, n0);
}

// site[i][q][j] = sum_p G[q][p] site[i][p][j]
template <typename R>
__global__ void __launch_bounds__(256)
apply_1q_kernel(long long chi_l, long long chi_r, cx<R> *__restrict__ site, const cx<R> *__restrict__ gate) {
  using T = cx<R>;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= chi_l * chi_r) return;
  const long long i = e / chi_r, j = e % chi_r;
  T *p0 = site + (i * 2) * chi_r + j, *p1 = p0 + chi_r;
  const T x0 = *p0, x1 = *p1;
  const T g00 = gate[0], g01 = gate[1], g10 = gate[2], g11 = gate[3];
  T y0, y1;
  y0.x = 0; y0.y = 0; y1.x = 0; y1.y = 0;
  cfma<R>(y0, g00, x0); cfma<R>(y0, g01, x1);
  cfma<R>(y1, g10, x0); cfma<R>(y1, g11, x1);
  *p0 = y0; *p1 = y1;
}

// theta[(i,r),(s,k)] = sum_{p,q} G[r][s][p][q] * theta[(i,p),(q,k)], theta is (2*chi_l) x (2*chi_r)
template <typename R>
__global__ void __launch_bounds__(256)
apply_2q_kernel(long long chi_l, long long chi_r, cx<R> *__restrict__ theta, const cx<R> *__restrict__ gate) {
  using T = cx<R>;
  __shared__ T g[16];
  if (threadIdx.x < 16) g[threadIdx.x] = gate[threadIdx.x];
  __syncthreads();
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= chi_l * chi_r) return;
  const long long i = e / chi_r, k = e % chi_r;
  const long long ld = 2 * chi_r;
  T *base = theta + (2 * i) * ld + k;
  T x[4];
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int q = 0; q < 2; ++q) x[p * 2 + q] = base[p * ld + q * chi_r];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      T y; y.x = 0; y.y = 0;
#pragma unroll
      for (int pq = 0; pq < 4; ++pq) cfma<R>(y, g[(r * 2 + s) * 4 + pq], x[pq]);
      base[r * ld + s * chi_r] = y;
    }
}

// One site of perfect sampling from an MPS (see qtn_mps_sample_step): a warp per shot.
// w_p[s,:] = v[s,:] A^p, z_p[s,:] = w_p[s,:] R with R the right environment of the next bond, so that
// q_p = Re <z_p[s,:], w_p[s,:]> is the weight of outcome p given the bits already drawn.
template <typename R>
__global__ void __launch_bounds__(256)
sample_step_kernel(long long nshots, long long r, const cx<R> *__restrict__ w0, const cx<R> *__restrict__ w1,
                   const cx<R> *__restrict__ z0, const cx<R> *__restrict__ z1, const double *__restrict__ u,
                   uint8_t *__restrict__ bits, long long bit_stride, double *__restrict__ ptotal,
                   cx<R> *__restrict__ vnext) {
  const long long s = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (s >= nshots) return;
  const cx<R> *a0 = w0 + s * r, *a1 = w1 + s * r, *b0 = z0 + s * r, *b1 = z1 + s * r;
  double q0 = 0, q1 = 0;
  for (long long c = lane; c < r; c += 32) {
    q0 += (double)b0[c].x * a0[c].x + (double)b0[c].y * a0[c].y;
    q1 += (double)b1[c].x * a1[c].x + (double)b1[c].y * a1[c].y;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    q0 += __shfl_xor_sync(0xffffffffu, q0, o);
    q1 += __shfl_xor_sync(0xffffffffu, q1, o);
  }
  q0 = fmax(q0, 0.0); q1 = fmax(q1, 0.0);
  const double tot = q0 + q1;
  // outcome 1 iff u >= P(0); a shot whose weight vanished (numerically impossible prefix) reads 0
  const int bit = tot > 0 ? (u[s] * tot >= q0 ? 1 : 0) : 0;
  const double qb = bit ? q1 : q0;
  const cx<R> *src = bit ? a1 : a0;
  const double scale = qb > 0 ? rsqrt(qb) : 0.0;             // keeps v R v^H = 1 along the chain
  for (long long c = lane; c < r; c += 32) {
    cx<R> v;
    v.x = (R)(src[c].x * scale); v.y = (R)(src[c].y * scale);
    vnext[s * r + c] = v;
  }
  if (lane == 0) {
    bits[s * bit_stride] = (uint8_t)bit;
    ptotal[s] *= tot > 0 ? qb / tot : 0.0;
  }
}

template <typename R>
int gemm_launch(int opa, int opb, long long m, long long n, long long k, const void *a, long long lda,
                const void *b, long long ldb, void *c, long long ldc, int accumulate, cudaStream_t st) {
  dim3 grid((unsigned)((n + GT - 1) / GT), (unsigned)((m + GT - 1) / GT));
  gemm_kernel<R><<<grid, 256, 0, st>>>(opa, opb, m, n, k, (const cx<R> *)a, lda, (const cx<R> *)b, ldb,
                                       (cx<R> *)c, ldc, accumulate);
  QTN_LAUNCH_CHECK();
  return QTN_OK;
}

}  // namespace

extern "C" int qtn_gemm(int dtype, int opa, int opb, int64_t m, int64_t n, int64_t k, const void *a,
                        int64_t lda, const void *b, int64_t ldb, void *c, int64_t ldc, int accumulate,
                        void *stream) {
  if (dtype != QTN_C64 && dtype != QTN_C128) return qtn_fail(QTN_ERR_ARG, "qtn_gemm: bad dtype");
  if (m < 0 || n < 0 || k < 0 || (opa & ~3) || (opb & ~3)) return qtn_fail(QTN_ERR_ARG, "qtn_gemm: bad argument");
  if (m == 0 || n == 0) return QTN_OK;
  if (!a || !b || !c) return qtn_fail(QTN_ERR_ARG, "qtn_gemm: null pointer");
  if ((m + GT - 1) / GT > 65535) return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_gemm: m too large");
  cudaStream_t st = (cudaStream_t)stream;
  return dtype == QTN_C64 ? gemm_launch<float>(opa, opb, m, n, k, a, lda, b, ldb, c, ldc, accumulate, st)
                          : gemm_launch<double>(opa, opb, m, n, k, a, lda, b, ldb, c, ldc, accumulate, st);
}

extern "C" int qtn_mps_apply_1q(int dtype, int64_t chi_l, int64_t chi_r, void *site, const void *gate,
                                void *stream) {
  if (dtype != QTN_C64 && dtype != QTN_C128) return qtn_fail(QTN_ERR_ARG, "qtn_mps_apply_1q: bad dtype");
  if (chi_l <= 0 || chi_r <= 0 || !site || !gate) return qtn_fail(QTN_ERR_ARG, "qtn_mps_apply_1q: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((chi_l * chi_r + 255) / 256);
  if (dtype == QTN_C64)
    apply_1q_kernel<float><<<grid, 256, 0, st>>>(chi_l, chi_r, (float2 *)site, (const float2 *)gate);
  else
    apply_1q_kernel<double><<<grid, 256, 0, st>>>(chi_l, chi_r, (double2 *)site, (const double2 *)gate);
  QTN_LAUNCH_CHECK();
  return QTN_OK;
}

extern "C" int qtn_mps_apply_2q(int dtype, int64_t chi_l, int64_t chi_r, void *theta, const void *gate,
                                void *stream) {
  if (dtype != QTN_C64 && dtype != QTN_C128) return qtn_fail(QTN_ERR_ARG, "qtn_mps_apply_2q: bad dtype");
  if (chi_l <= 0 || chi_r <= 0 || !theta || !gate) return qtn_fail(QTN_ERR_ARG, "qtn_mps_apply_2q: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((chi_l * chi_r + 255) / 256);
  if (dtype == QTN_C64)
    apply_2q_kernel<float><<<grid, 256, 0, st>>>(chi_l, chi_r, (float2 *)theta, (const float2 *)gate);
  else
    apply_2q_kernel<double><<<grid, 256, 0, st>>>(chi_l, chi_r, (double2 *)theta, (const double2 *)gate);
  QTN_LAUNCH_CHECK();
  return QTN_OK;
}

extern "C" int qtn_mps_sample_step(int dtype, int64_t nshots, int64_t r, const void *w0, const void *w1,
                                   const void *z0, const void *z1, const double *u, uint8_t *bits,
                                   int64_t bit_stride, double *ptotal, void *vnext, void *stream) {
  if (dtype != QTN_C64 && dtype != QTN_C128) return qtn_fail(QTN_ERR_ARG, "qtn_mps_sample_step: bad dtype");
  if (nshots <= 0 || r <= 0 || bit_stride <= 0 || !w0 || !w1 || !z0 || !z1 || !u || !bits || !ptotal || !vnext)
    return qtn_fail(QTN_ERR_ARG, "qtn_mps_sample_step: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)((nshots + 7) / 8);
  if (dtype == QTN_C64)
    sample_step_kernel<float><<<grid, 256, 0, st>>>(nshots, r, (const float2 *)w0, (const float2 *)w1,
                                                    (const float2 *)z0, (const float2 *)z1, u, bits, bit_stride,
                                                    ptotal, (float2 *)vnext);
  else
    sample_step_kernel<double><<<grid, 256, 0, st>>>(nshots, r, (const double2 *)w0, (const double2 *)w1,
                                                     (const double2 *)z0, (const double2 *)z1, u, bits, bit_stride,
                                                     ptotal, (double2 *)vnext);
  QTN_LAUNCH_CHECK();
  return QTN_OK;
}

extern "C" int qtn_gemm_batched(int dtype, int count, const qtn_gemm_item *items, int accumulate, void *stream) {
  if (dtype != QTN_C64 && dtype != QTN_C128) return qtn_fail(QTN_ERR_ARG, "qtn_gemm_batched: bad dtype");
  if (count < 0 || (count > 0 && !items)) return qtn_fail(QTN_ERR_ARG, "qtn_gemm_batched: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  for (int base = 0; base < count; base += kGemmBatch) {
    GemmBatch batch;
    batch.count = std::min(kGemmBatch, count - base);
    long long tm = 0, tn = 0;
    int live = 0;
    for (int i = 0; i < batch.count; ++i) {
      const qtn_gemm_item &g = items[base + i];
      if (g.m < 0 || g.n < 0 || g.k < 0 || (g.opa & ~3) || (g.opb & ~3))
        return qtn_fail(QTN_ERR_ARG, "qtn_gemm_batched: bad item");
      if (g.m == 0 || g.n == 0) continue;
      if (!g.a || !g.b || !g.c) return qtn_fail(QTN_ERR_ARG, "qtn_gemm_batched: null pointer");
      batch.it[live++] = g;
      tm = std::max(tm, ((long long)g.m + GT - 1) / GT);
      tn = std::max(tn, ((long long)g.n + GT - 1) / GT);
    }
    if (!live) continue;
    if (tm > 65535) return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_gemm_batched: m too large");
    batch.count = live;
    dim3 grid((unsigned)tn, (unsigned)tm, (unsigned)live);
    if (dtype == QTN_C64) gemm_batched_kernel<float><<<grid, 256, 0, st>>>(batch, accumulate);
    else gemm_batched_kernel<double><<<grid, 256, 0, st>>>(batch, accumulate);
    QTN_LAUNCH_CHECK();
  }
  return QTN_OK;
}
