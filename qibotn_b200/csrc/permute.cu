// Tensor permutes and small device utilities.
//
// qtn_permute_bits: permutation of a bit tensor (every extent 2).  A tile is the
// set of labels that occupy the low 6 address bits of the input or of the
// output (<= 12 labels, 4096 elements).  The tile is read in ascending input
// address order, parked in shared memory at its output-order index (XOR-swizzled
// against bank conflicts) and written in ascending output address order, so both
// sides of the copy are coalesced.  HBM-bound: 2 * 2^rank * sizeof(elem) bytes.
#include <algorithm>
#include <cstdint>
#include <vector>

#include "qtn_internal.h"

namespace {

struct BitPermParams {
  int nt;                 // tile bits
  int nhi;                // tile-index bits
  uint8_t in_pos[12];     // tile labels sorted by input position
  uint16_t in_sidx[12];   // staging index bit of that label (1 << rank in output order)
  uint8_t out_pos[12];    // tile labels sorted by output position
  uint8_t hi_in[64], hi_out[64];
  int accumulate;
};

__device__ __forceinline__ int swz(int s) { return s ^ ((s >> 6) & 63); }

template <typename T>
__global__ void __launch_bounds__(256)
permute_bits_kernel(const __grid_constant__ BitPermParams P, const T *__restrict__ in,
                    T *__restrict__ out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *stage = reinterpret_cast<T *>(smem_raw);
  __shared__ long long it_in[16], it_out[16];
  __shared__ int it_s[16];
  const int t = threadIdx.x;
  if (t < 16) {
    long long gi = 0, go = 0; int s = 0;
    for (int j = 8; j < P.nt; ++j)
      if ((t >> (j - 8)) & 1) { gi |= 1ll << P.in_pos[j]; s += P.in_sidx[j]; go |= 1ll << P.out_pos[j]; }
    it_in[t] = gi; it_out[t] = go; it_s[t] = s;
  }
  long long ti = 0, to = 0; int ts = 0;
  for (int j = 0; j < 8 && j < P.nt; ++j)
    if ((t >> j) & 1) { ti |= 1ll << P.in_pos[j]; ts += P.in_sidx[j]; to |= 1ll << P.out_pos[j]; }
  long long base_in = 0, base_out = 0;
  {
    unsigned long long id = blockIdx.x;
    for (int i = 0; i < P.nhi; ++i)
      if ((id >> i) & 1ull) { base_in |= 1ll << P.hi_in[i]; base_out |= 1ll << P.hi_out[i]; }
  }
  __syncthreads();
  const int count = 1 << P.nt;
  for (int e = t, it = 0; e < count; e += 256, ++it)
    stage[swz(ts + it_s[it])] = in[base_in + (ti | it_in[it])];
  __syncthreads();
  for (int e = t, it = 0; e < count; e += 256, ++it) {
    T v = stage[swz(e)];
    const long long g = base_out + (to | it_out[it]);
    if (P.accumulate) { const T o = out[g]; v.x += o.x; v.y += o.y; }
    out[g] = v;
  }
}

// Vectorised variant for full tiles (12 labels, 4096 elements) of dense layouts, where the label at input
// address bit 0 and the label at output address bit 0 are tile labels by construction: every global access
// moves 16 bytes (two complex64 values / one complex128 value), all 8 loads of a thread are in flight before
// the first shared-memory store, and the write side reads its pairs with one 16-byte shared-memory load.
// Staging index = rank in output order, so an output pair is adjacent in the stage; the XOR swizzle leaves
// bit 0 alone to keep it so.
__device__ __forceinline__ int swz2(int s) { return s ^ (((s >> 6) & 31) << 1); }

template <typename T>
__global__ void __launch_bounds__(256)
permute_bits_vec_kernel(const __grid_constant__ BitPermParams P, const T *__restrict__ in, T *__restrict__ out) {
  constexpr int PER = sizeof(T) == 8 ? 2 : 1;          // elements per 16-byte access
  constexpr int NV = 4096 / PER / 256;                 // 16-byte accesses per thread
  struct __align__(16) V { T e[PER]; };
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T *stage = reinterpret_cast<T *>(smem_raw);
  const int t = threadIdx.x;
  constexpr int LOW = PER == 2 ? 1 : 0;                // enumeration starts above the pair bit
  long long ti = 0, to = 0; int ts = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if ((t >> j) & 1) { ti |= 1ll << P.in_pos[j + LOW]; ts += P.in_sidx[j + LOW]; to |= 1ll << P.out_pos[j + LOW]; }
  long long base_in = 0, base_out = 0;
  {
    unsigned long long id = blockIdx.x;
    for (int i = 0; i < P.nhi; ++i)
      if ((id >> i) & 1ull) { base_in |= 1ll << P.hi_in[i]; base_out |= 1ll << P.hi_out[i]; }
  }
  V v[NV];
#pragma unroll
  for (int it = 0; it < NV; ++it) {
    long long gi = 0;
#pragma unroll
    for (int j = 8 + LOW; j < 12; ++j)
      if ((it >> (j - 8 - LOW)) & 1) gi |= 1ll << P.in_pos[j];
    v[it] = *reinterpret_cast<const V *>(in + base_in + (ti | gi));
  }
  const int pair_s = PER == 2 ? P.in_sidx[0] : 0;      // staging step of the label at input bit 0
#pragma unroll
  for (int it = 0; it < NV; ++it) {
    int s = ts;
#pragma unroll
    for (int j = 8 + LOW; j < 12; ++j)
      if ((it >> (j - 8 - LOW)) & 1) s += P.in_sidx[j];
    stage[swz2(s)] = v[it].e[0];
    if constexpr (PER == 2) stage[swz2(s + pair_s)] = v[it].e[1];
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < NV; ++it) {
    long long go = 0;
#pragma unroll
    for (int j = 8 + LOW; j < 12; ++j)
      if ((it >> (j - 8 - LOW)) & 1) go |= 1ll << P.out_pos[j];
    const int e = (t + it * 256) * PER;                // staging index of the (first) element
    V w = *reinterpret_cast<const V *>(stage + swz2(e));
    V *dst = reinterpret_cast<V *>(out + base_out + (to | go));
    if (P.accumulate) {
      const V o = *dst;
#pragma unroll
      for (int q = 0; q < PER; ++q) { w.e[q].x += o.e[q].x; w.e[q].y += o.e[q].y; }
    }
    *dst = w;
  }
}

// Rank-2 transpose (the MPS path's only permute: theta <-> theta^T, factor transposes): 32 x 32 tiles through
// padded shared memory, both sides coalesced -- the general kernel below gathers with a div/mod chain per element.
template <typename T>
__global__ void __launch_bounds__(256)
transpose_kernel(long long rows, long long cols, const T *__restrict__ in, T *__restrict__ out) {
  __shared__ T tile[32][33];
  const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const long long r = r0 + ty + k, c = c0 + tx;
    if (r < rows && c < cols) tile[ty + k][tx] = in[r * cols + c];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const long long c = c0 + ty + k, r = r0 + tx;
    if (r < rows && c < cols) out[c * rows + r] = tile[tx][ty + k];
  }
}

struct PermParams {
  int rank;
  long long n;
  long long out_extent[8];
  long long in_stride[8];   // stride in the input of output axis i
};

template <typename T>
__global__ void __launch_bounds__(256)
permute_general_kernel(const __grid_constant__ PermParams P, const T *__restrict__ in,
                       T *__restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < P.n;
       i += (long long)gridDim.x * blockDim.x) {
    long long r = i, off = 0;
    for (int a = P.rank - 1; a >= 0; --a) {
      const long long q = r / P.out_extent[a];
      off += (r - q * P.out_extent[a]) * P.in_stride[a];
      r = q;
    }
    out[i] = in[off];
  }
}

template <typename T>
__global__ void axpy_kernel(long long n, T alpha, const T *__restrict__ x, T *__restrict__ y,
                            int accumulate) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const T v = x[i];
    T r;
    r.x = alpha.x * v.x - alpha.y * v.y;
    r.y = alpha.x * v.y + alpha.y * v.x;
    if (accumulate) { r.x += y[i].x; r.y += y[i].y; }
    y[i] = r;
  }
}

__global__ void set_i32_kernel(int32_t *dst, int32_t v) { *dst = v; }
__global__ void add_i32_kernel(int32_t *dst, int32_t d) { *dst += d; }

}  // namespace

extern "C" int qtn_permute_bits(int dtype, int rank, const uint8_t *pos_in, const uint8_t *pos_out,
                                const void *in, void *out, int accumulate, void *stream) {
  if (rank < 0 || rank > 62 || !in || !out || (rank && (!pos_in || !pos_out)))
    return qtn_fail(QTN_ERR_ARG, "qtn_permute_bits: bad argument");
  if (dtype != QTN_C64 && dtype != QTN_C128) return qtn_fail(QTN_ERR_ARG, "qtn_permute_bits: bad dtype");
  BitPermParams P{};
  P.accumulate = accumulate;
  std::vector<int> tile, hi;
  int c = 6;
  for (;;) {
    tile.clear(); hi.clear();
    for (int i = 0; i < rank; ++i) (pos_in[i] < c || pos_out[i] < c ? tile : hi).push_back(i);
    if ((int)tile.size() <= 12) break;
    --c;   // only reachable with gapped layouts
  }
  // grow the tile with the next-lowest labels while there is room (fewer, fatter CTAs)
  std::sort(hi.begin(), hi.end(), [&](int x, int y) {
    return std::min(pos_in[x], pos_out[x]) < std::min(pos_in[y], pos_out[y]);
  });
  while ((int)tile.size() < 12 && !hi.empty()) { tile.push_back(hi.front()); hi.erase(hi.begin()); }
  P.nt = (int)tile.size();
  P.nhi = (int)hi.size();
  if (P.nhi > 31) return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_permute_bits: tensor too large");
  std::vector<int> by_in = tile, by_out = tile;
  std::sort(by_in.begin(), by_in.end(), [&](int x, int y) { return pos_in[x] < pos_in[y]; });
  std::sort(by_out.begin(), by_out.end(), [&](int x, int y) { return pos_out[x] < pos_out[y]; });
  for (int r = 0; r < P.nt; ++r) {
    P.in_pos[r] = pos_in[by_in[r]];
    P.out_pos[r] = pos_out[by_out[r]];
    const int orank = (int)(std::find(by_out.begin(), by_out.end(), by_in[r]) - by_out.begin());
    P.in_sidx[r] = (uint16_t)(1 << orank);
  }
  for (int i = 0; i < P.nhi; ++i) { P.hi_in[i] = pos_in[hi[i]]; P.hi_out[i] = pos_out[hi[i]]; }
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = 1u << P.nhi;
  // full tiles of dense layouts: 16-byte accesses on both sides (the pair bit is the label at address bit 0)
  const bool vec = P.nt == 12 && P.in_pos[0] == 0 && P.out_pos[0] == 0 &&
                   ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
  if (vec) {
    // in the vector kernel a thread's 16-byte output access covers staging indices e, e + 1: the enumeration of
    // the write side is by output order, whose bit 0 is the label at output address bit 0
    if (dtype == QTN_C64) {
      permute_bits_vec_kernel<float2><<<grid, 256, 4096 * sizeof(float2), st>>>(P, (const float2 *)in, (float2 *)out);
    } else {
      static bool attr_v = false;
      if (!attr_v) {
        QTN_CUDA_CHECK(cudaFuncSetAttribute(permute_bits_vec_kernel<double2>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, 4096 * (int)sizeof(double2)));
        attr_v = true;
      }
      permute_bits_vec_kernel<double2><<<grid, 256, 4096 * sizeof(double2), st>>>(P, (const double2 *)in, (double2 *)out);
    }
    QTN_LAUNCH_CHECK();
    return QTN_OK;
  }
  if (dtype == QTN_C64) {
    permute_bits_kernel<float2><<<grid, 256, 4096 * sizeof(float2), st>>>(P, (const float2 *)in,
                                                                        (float2 *)out);
  } else {
    static bool attr = false;
    if (!attr) {
      QTN_CUDA_CHECK(cudaFuncSetAttribute(permute_bits_kernel<double2>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          4096 * (int)sizeof(double2)));
      attr = true;
    }
    permute_bits_kernel<double2><<<grid, 256, 4096 * sizeof(double2), st>>>(P, (const double2 *)in,
                                                                          (double2 *)out);
  }
  QTN_LAUNCH_CHECK();
  return QTN_OK;
}

extern "C" int qtn_permute(int dtype, int rank, const int64_t *extent_in, const int32_t *perm,
                           const void *in, void *out, void *stream) {
  if (rank < 1 || rank > 8 || !extent_in || !perm || !in || !out)
    return qtn_fail(QTN_ERR_ARG, "qtn_permute: bad argument");
  if (dtype != QTN_C64 && dtype != QTN_C128) return qtn_fail(QTN_ERR_ARG, "qtn_permute: bad dtype");
  long long stride_in[8];
  long long s = 1;
  for (int a = rank - 1; a >= 0; --a) { stride_in[a] = s; s *= extent_in[a]; }
  PermParams P{};
  P.rank = rank;
  P.n = s;
  bool seen[8] = {false};
  for (int i = 0; i < rank; ++i) {
    if (perm[i] < 0 || perm[i] >= rank || seen[perm[i]]) return qtn_fail(QTN_ERR_ARG, "qtn_permute: bad perm");
    seen[perm[i]] = true;
    P.out_extent[i] = extent_in[perm[i]];
    P.in_stride[i] = stride_in[perm[i]];
  }
  if (P.n == 0) return QTN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (rank == 2 && perm[0] == 1 && perm[1] == 0 && extent_in[1] <= 0x7fffffffLL * 32 && (extent_in[0] + 31) / 32 <= 65535) {
    const dim3 tgrid((unsigned)((extent_in[1] + 31) / 32), (unsigned)((extent_in[0] + 31) / 32));
    if (dtype == QTN_C64)
      transpose_kernel<float2><<<tgrid, 256, 0, st>>>(extent_in[0], extent_in[1], (const float2 *)in, (float2 *)out);
    else
      transpose_kernel<double2><<<tgrid, 256, 0, st>>>(extent_in[0], extent_in[1], (const double2 *)in, (double2 *)out);
    QTN_LAUNCH_CHECK();
    return QTN_OK;
  }
  const unsigned grid = (unsigned)std::min<long long>((P.n + 255) / 256, 148 * 16);
  if (dtype == QTN_C64)
    permute_general_kernel<float2><<<grid, 256, 0, st>>>(P, (const float2 *)in, (float2 *)out);
  else
    permute_general_kernel<double2><<<grid, 256, 0, st>>>(P, (const double2 *)in, (double2 *)out);
  QTN_LAUNCH_CHECK();
  return QTN_OK;
}

extern "C" int qtn_axpy(int dtype, int64_t n, double alpha_re, double alpha_im, const void *x, void *y,
                        int accumulate, void *stream) {
  if (n < 0 || !x || !y) return qtn_fail(QTN_ERR_ARG, "qtn_axpy: bad argument");
  if (n == 0) return QTN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned grid = (unsigned)std::min<long long>((n + 255) / 256, 148 * 16);
  if (dtype == QTN_C64) {
    float2 a; a.x = (float)alpha_re; a.y = (float)alpha_im;
    axpy_kernel<float2><<<grid, 256, 0, st>>>(n, a, (const float2 *)x, (float2 *)y, accumulate);
  } else if (dtype == QTN_C128) {
    double2 a; a.x = alpha_re; a.y = alpha_im;
    axpy_kernel<double2><<<grid, 256, 0, st>>>(n, a, (const double2 *)x, (double2 *)y, accumulate);
  } else {
    return qtn_fail(QTN_ERR_ARG, "qtn_axpy: bad dtype");
  }
  QTN_LAUNCH_CHECK();
  return QTN_OK;
}

extern "C" int qtn_set_i32(int32_t *dst, int32_t value, void *stream) {
  if (!dst) return qtn_fail(QTN_ERR_ARG, "qtn_set_i32: null");
  set_i32_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dst, value);
  QTN_LAUNCH_CHECK();
  return QTN_OK;
}

extern "C" int qtn_add_i32(int32_t *dst, int32_t delta, void *stream) {
  if (!dst) return qtn_fail(QTN_ERR_ARG, "qtn_add_i32: null");
  add_i32_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(dst, delta);
  QTN_LAUNCH_CHECK();
  return QTN_OK;
}
