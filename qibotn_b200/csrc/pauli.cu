// Fused Pauli-string expectation on a dense state: <psi| P |psi> in ONE pass over psi, without
// materialising P|psi> (north_star item d).  It serves the expectation drivers
// (reference eval.py:455-477) when the state over the active qubits fits in HBM: the circuit is
// contracted once to a dense vector and every Hamiltonian term costs one streaming read.
//
// P = prod_q sigma_q acts on a basis state as  P|x> = i^{nY} (-1)^{popcount(x & zmask)} |x ^ xmask>
// (xmask = qubits carrying X or Y, zmask = qubits carrying Z or Y), so
//     <psi|P|psi> = i^{nY} sum_x (-1)^{popcount(x & zmask)} conj(psi[x ^ xmask]) psi[x].
// Both reads are coalesced (x ^ xmask permutes addresses inside aligned blocks).
// Algorithmic traffic: 2 * 2^n * sizeof(element) bytes per term (half when xmask == 0).
#include <algorithm>
#include <cstdint>
#include <vector>

#include "qtn_internal.h"

namespace {

template <typename R> struct cx_of;
template <> struct cx_of<float> { using type = float2; };
template <> struct cx_of<double> { using type = double2; };

template <typename R>
__global__ void __launch_bounds__(256)
pauli_expectation_kernel(int nqubits, const typename cx_of<R>::type *__restrict__ psi,
                         const uint64_t *__restrict__ xmask, const uint64_t *__restrict__ zmask,
                         const int32_t *__restrict__ ny, double2 *__restrict__ out) {
  using T = typename cx_of<R>::type;
  __shared__ double red[2][8];
  const int term = blockIdx.y;
  const uint64_t fx = xmask[term], fz = zmask[term];
  const uint64_t total = 1ull << nqubits;
  double sr = 0, si = 0;
  for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total;
       x += (uint64_t)gridDim.x * blockDim.x) {
    const T a = psi[x];
    const T b = fx ? psi[x ^ fx] : a;
    // conj(b) * a
    double pr = (double)b.x * a.x + (double)b.y * a.y;
    double pi = (double)b.x * a.y - (double)b.y * a.x;
    if (__popcll(x & fz) & 1) { pr = -pr; pi = -pi; }
    sr += pr; si += pi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sr += __shfl_xor_sync(0xffffffffu, sr, o);
    si += __shfl_xor_sync(0xffffffffu, si, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sr; red[1][threadIdx.x >> 5] = si; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double r = 0, i = 0;
    for (int w = 0; w < 8; ++w) { r += red[0][w]; i += red[1][w]; }
    // times i^ny
    double fr = r, fi = i;
    switch (ny[term] & 3) {
      case 1: fr = -i; fi = r; break;
      case 2: fr = -r; fi = -i; break;
      case 3: fr = i; fi = -r; break;
      default: break;
    }
    atomicAdd(&out[term].x, fr);
    atomicAdd(&out[term].y, fi);
  }
}

// Multi-term pass: every term of a group carries the SAME xmask, so the pair (psi[x], psi[x ^ xmask]) and the
// product conj(psi[x ^ xmask]) psi[x] are read / formed once and only the sign (-1)^{popcount(x & zmask_t)}
// differs per term: one streaming pass over psi serves up to kGroup terms (a sum of ZZ terms, xmask = 0,
// costs ceil(nterms / 16) reads of psi instead of nterms).
constexpr int kGroup = 16;
struct PauliGroup {
  uint64_t xmask;
  uint64_t zmask[kGroup];
  int32_t ny[kGroup];
  int32_t index[kGroup];      // position of the term in the caller's list
  int32_t count;
};

template <typename R>
__global__ void __launch_bounds__(256)
pauli_group_kernel(int nqubits, const typename cx_of<R>::type *__restrict__ psi, const __grid_constant__ PauliGroup G,
                   double2 *__restrict__ out) {
  using T = typename cx_of<R>::type;
  __shared__ double red[8][2 * kGroup];
  const uint64_t fx = G.xmask;
  const uint64_t total = 1ull << nqubits;
  double acc[2 * kGroup];
#pragma unroll
  for (int j = 0; j < 2 * kGroup; ++j) acc[j] = 0.0;
  for (uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < total;
       x += (uint64_t)gridDim.x * blockDim.x) {
    const T a = psi[x];
    const T b = fx ? psi[x ^ fx] : a;
    const double pr = (double)b.x * a.x + (double)b.y * a.y;      // conj(b) * a
    const double pi = (double)b.x * a.y - (double)b.y * a.x;
#pragma unroll
    for (int j = 0; j < kGroup; ++j) {
      if (j < G.count) {
        const bool neg = __popcll(x & G.zmask[j]) & 1;
        acc[2 * j] += neg ? -pr : pr;
        acc[2 * j + 1] += neg ? -pi : pi;
      }
    }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 2 * kGroup; ++j) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < G.count) {
    const int j = threadIdx.x;
    double r = 0, i = 0;
    for (int w = 0; w < 8; ++w) { r += red[w][2 * j]; i += red[w][2 * j + 1]; }
    double fr = r, fi = i;                                        // times i^ny
    switch (G.ny[j] & 3) {
      case 1: fr = -i; fi = r; break;
      case 2: fr = -r; fi = -i; break;
      case 3: fr = i; fi = -r; break;
      default: break;
    }
    atomicAdd(&out[G.index[j]].x, fr);
    atomicAdd(&out[G.index[j]].y, fi);
  }
}

}  // namespace

extern "C" int qtn_pauli_expectation_host(int dtype, int nqubits, const void *state, int nterms,
                                          const uint64_t *xmask, const uint64_t *zmask, const int32_t *ny,
                                          void *out, void *stream) {
  if (dtype != QTN_C64 && dtype != QTN_C128) return qtn_fail(QTN_ERR_ARG, "qtn_pauli_expectation_host: bad dtype");
  if (nqubits < 0 || nqubits > 40 || nterms < 1 || nterms > (1 << 24) || !state || !xmask || !zmask || !ny || !out)
    return qtn_fail(QTN_ERR_ARG, "qtn_pauli_expectation_host: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  QTN_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)nterms * sizeof(double2), st));
  const uint64_t total = 1ull << nqubits;
  uint64_t blocks = (total + 256 * 8 - 1) / (256 * 8);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  // group the terms by xmask (stable: the order inside a group is the caller's)
  std::vector<int> order(nterms);
  for (int i = 0; i < nterms; ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return xmask[a] < xmask[b]; });
  for (int i = 0; i < nterms;) {
    PauliGroup G{};
    G.xmask = xmask[order[i]];
    while (i < nterms && G.count < kGroup && xmask[order[i]] == G.xmask) {
      G.zmask[G.count] = zmask[order[i]];
      G.ny[G.count] = ny[order[i]];
      G.index[G.count] = order[i];
      ++G.count; ++i;
    }
    if (dtype == QTN_C64)
      pauli_group_kernel<float><<<(unsigned)blocks, 256, 0, st>>>(nqubits, (const float2 *)state, G, (double2 *)out);
    else
      pauli_group_kernel<double><<<(unsigned)blocks, 256, 0, st>>>(nqubits, (const double2 *)state, G, (double2 *)out);
    QTN_LAUNCH_CHECK();
  }
  return QTN_OK;
}

extern "C" int qtn_pauli_expectation(int dtype, int nqubits, const void *state, int nterms,
                                     const uint64_t *xmask, const uint64_t *zmask, const int32_t *ny,
                                     void *out, void *stream) {
  if (dtype != QTN_C64 && dtype != QTN_C128) return qtn_fail(QTN_ERR_ARG, "qtn_pauli_expectation: bad dtype");
  if (nqubits < 0 || nqubits > 40 || nterms < 1 || nterms > 65535 || !state || !xmask || !zmask || !ny || !out)
    return qtn_fail(QTN_ERR_ARG, "qtn_pauli_expectation: bad argument");
  cudaStream_t st = (cudaStream_t)stream;
  QTN_CUDA_CHECK(cudaMemsetAsync(out, 0, (size_t)nterms * sizeof(double2), st));
  const uint64_t total = 1ull << nqubits;
  uint64_t blocks = (total + 256 * 8 - 1) / (256 * 8);          // ~8 elements per thread
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 16) blocks = 148 * 16;
  dim3 grid((unsigned)blocks, (unsigned)nterms);
  if (dtype == QTN_C64)
    pauli_expectation_kernel<float><<<grid, 256, 0, st>>>(nqubits, (const float2 *)state, xmask, zmask, ny,
                                                          (double2 *)out);
  else
    pauli_expectation_kernel<double><<<grid, 256, 0, st>>>(nqubits, (const double2 *)state, xmask, zmask, ny,
                                                           (double2 *)out);
  QTN_LAUNCH_CHECK();
  return QTN_OK;
}
