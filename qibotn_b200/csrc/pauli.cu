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
#include <cstdint>

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

}  // namespace

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
