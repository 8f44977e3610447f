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
  auto add = [&](uint64_t x, const T a, const T b) {
    // conj(b) * a
    double pr = (double)b.x * a.x + (double)b.y * a.y;
    double pi = (double)b.x * a.y - (double)b.y * a.x;
    if (__popcll(x & fz) & 1) { pr = -pr; pi = -pi; }
    sr += pr; si += pi;
  };
  constexpr int U = 4;                         // independent loads in flight per thread
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t x0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; x0 + (U - 1) * stride < total; x0 += U * stride) {
    T a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = psi[x0 + u * stride];
#pragma unroll
    for (int u = 0; u < U; ++u) b[u] = fx ? psi[(x0 + u * stride) ^ fx] : a[u];
#pragma unroll
    for (int u = 0; u < U; ++u) add(x0 + u * stride, a[u], b[u]);
  }
  for (; x0 < total; x0 += stride) {
    const T a = psi[x0];
    add(x0, a, fx ? psi[x0 ^ fx] : a);
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

// Sign of a term on basis state x as a mask for the SIGN BIT of a double: parity of popcount(x & zmask) moved to
// bit 31 of the high word (no select, no conversion; the accumulation is then one integer XOR and one DADD).
template <bool W32>
__device__ __forceinline__ uint32_t sign_word(uint64_t x, uint64_t z) {
  if constexpr (W32) return (uint32_t)__popc((uint32_t)x & (uint32_t)z) << 31;
  else return (uint32_t)__popcll(x & z) << 31;
}
__device__ __forceinline__ double flip_sign(double v, uint32_t sw) {
  return __hiloint2double(__double2hiint(v) ^ (int)sw, __double2loint(v));
}

// ZONLY: xmask == 0, the pair product is |psi[x]|^2 (real): half the accumulators.  W32: at most 32 qubits, masks
// and basis states fit 32-bit words.  U independent 16-byte loads per thread are in flight before the first use
// (the first version of this kernel had one 8-byte load in flight per thread and ran at 0.47 TB/s:
// profiles/r02_ncu_hbm_kernels.txt).
template <typename R, bool ZONLY, bool W32>
__global__ void __launch_bounds__(256)
pauli_group_kernel(int nqubits, const typename cx_of<R>::type *__restrict__ psi, const __grid_constant__ PauliGroup G,
                   double2 *__restrict__ out) {
  using T = typename cx_of<R>::type;
  constexpr int NACC = ZONLY ? kGroup : 2 * kGroup;
  constexpr int U = 4;
  __shared__ double red[8][NACC];
  const uint64_t fx = G.xmask;
  const uint64_t total = 1ull << nqubits;
  double acc[NACC];
#pragma unroll
  for (int j = 0; j < NACC; ++j) acc[j] = 0.0;
  auto add = [&](uint64_t x, const T a, const T b) {
    const double pr = (double)b.x * a.x + (double)b.y * a.y;      // conj(b) * a
    double pi = 0.0;
    if constexpr (!ZONLY) pi = (double)b.x * a.y - (double)b.y * a.x;
#pragma unroll
    for (int j = 0; j < kGroup; ++j) {
      if (j < G.count) {
        const uint32_t sw = sign_word<W32>(x, G.zmask[j]);
        if constexpr (ZONLY) acc[j] += flip_sign(pr, sw);
        else { acc[2 * j] += flip_sign(pr, sw); acc[2 * j + 1] += flip_sign(pi, sw); }
      }
    }
  };
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  uint64_t x0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; x0 + (U - 1) * stride < total; x0 += U * stride) {
    T a[U], b[U];
#pragma unroll
    for (int u = 0; u < U; ++u) a[u] = psi[x0 + u * stride];
    if constexpr (!ZONLY) {
#pragma unroll
      for (int u = 0; u < U; ++u) b[u] = psi[(x0 + u * stride) ^ fx];
    }
#pragma unroll
    for (int u = 0; u < U; ++u) add(x0 + u * stride, a[u], ZONLY ? a[u] : b[u]);
  }
  for (; x0 < total; x0 += stride) {
    const T a = psi[x0];
    add(x0, a, ZONLY ? a : psi[x0 ^ fx]);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < NACC; ++j) {
    double v = acc[j];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) red[warp][j] = v;
  }
  __syncthreads();
  if (threadIdx.x < G.count) {
    const int j = threadIdx.x;
    double r = 0, i = 0;
    for (int w = 0; w < 8; ++w) {
      if constexpr (ZONLY) r += red[w][j];
      else { r += red[w][2 * j]; i += red[w][2 * j + 1]; }
    }
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

template <typename R>
void launch_group(unsigned blocks, cudaStream_t st, int nqubits, const void *state, const PauliGroup &G, double2 *out) {
  using T = typename cx_of<R>::type;
  const bool z = G.xmask == 0, w32 = nqubits <= 32;
  if (z && w32) pauli_group_kernel<R, true, true><<<blocks, 256, 0, st>>>(nqubits, (const T *)state, G, out);
  else if (z) pauli_group_kernel<R, true, false><<<blocks, 256, 0, st>>>(nqubits, (const T *)state, G, out);
  else if (w32) pauli_group_kernel<R, false, true><<<blocks, 256, 0, st>>>(nqubits, (const T *)state, G, out);
  else pauli_group_kernel<R, false, false><<<blocks, 256, 0, st>>>(nqubits, (const T *)state, G, out);
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
  uint64_t blocks = (total + 256 * 16 - 1) / (256 * 16);
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 6) blocks = 148 * 6;
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
    if (dtype == QTN_C64) launch_group<float>((unsigned)blocks, st, nqubits, state, G, (double2 *)out);
    else launch_group<double>((unsigned)blocks, st, nqubits, state, G, (double2 *)out);
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
