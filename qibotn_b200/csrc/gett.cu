// Pairwise contraction of bit tensors (every extent 2) with the permutation
// fused into the loads and stores: a GETT-style kernel driven by bit-deposit
// address algebra instead of a separate transpose pass.
//
// Address model: each label of a tensor owns one *address bit*.  A CTA works on
// a tile whose labels (m_lo, n_lo, k_lo) were chosen on the host so that they
// cover the low address bits of A, B and C; the remaining labels index tiles
// (m_hi, n_hi, batch) or loop iterations (k_hi).  Inside a tile, elements are
// enumerated in ascending *address* order of the operand being moved, so global
// traffic is coalesced whatever the logical (GEMM) order of the labels is.
//
// Arithmetic: complex FMA on the FP32 / FP64 pipes (4 real FMA per complex MAC),
// 4x4 complex outputs per thread, 256 threads -> 4096 outputs per CTA.
#include <cstdint>
#include <type_traits>

#include "qtn_internal.h"

namespace {

template <typename R> struct cx_of;
template <> struct cx_of<float> { using type = float2; };
template <> struct cx_of<double> { using type = double2; };
template <typename R> using cx = typename cx_of<R>::type;

template <typename R> __device__ __forceinline__ cx<R> cx_zero() { cx<R> z; z.x = 0; z.y = 0; return z; }

template <typename R>
__device__ __forceinline__ void cfma(cx<R> &acc, const cx<R> &a, const cx<R> &b) {
  acc.x = fma(a.x, b.x, acc.x);
  acc.x = fma(-a.y, b.y, acc.x);
  acc.y = fma(a.x, b.y, acc.y);
  acc.y = fma(a.y, b.x, acc.y);
}

__device__ __forceinline__ long long deposit(unsigned long long idx, const uint8_t *pos, int n) {
  long long off = 0;
  for (int i = 0; i < n; ++i) off |= (long long)((idx >> i) & 1ull) << pos[i];
  return off;
}

// the same for label lists in which position 255 marks a label the operand does not carry
__device__ __forceinline__ long long deposit_present(unsigned long long idx, const uint8_t *pos, int n) {
  long long off = 0;
  for (int i = 0; i < n; ++i)
    if (pos[i] != 255) off |= (long long)((idx >> i) & 1ull) << pos[i];
  return off;
}

__device__ __forceinline__ long long slice_offset(const int32_t *slice_id, int n, const uint8_t *bit,
                                                  const uint8_t *pos) {
  if (n == 0) return 0;
  const unsigned sid = (unsigned)(*slice_id);
  long long off = 0;
  for (int i = 0; i < n; ++i) off |= (long long)((sid >> bit[i]) & 1u) << pos[i];
  return off;
}

// --------------------------------------------------------------------------
// Tiled kernel
// --------------------------------------------------------------------------
// cp.async of one element (8 bytes complex64 / 16 bytes complex128); `valid` false = zero fill (phantom rows)
template <int BYTES>
__device__ __forceinline__ void cp_async_elem(void *dst_smem, const void *src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int n = valid ? BYTES : 0;
  if constexpr (BYTES == 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
  else
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

template <typename R>
__global__ void __launch_bounds__(256, sizeof(R) == 4 ? 3 : 2)
gett_tile_kernel(const __grid_constant__ qtn_pair_plan_t P, const cx<R> *__restrict__ A,
                 const cx<R> *__restrict__ B, cx<R> *__restrict__ C, cx<R> *__restrict__ ws,
                 const int32_t *__restrict__ slice_id) {
  using T = cx<R>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // two stages of [A tile | B tile]; the epilogue's staging buffer (4096 outputs) aliases the stage that was
  // consumed last when it fits there (complex64: the other stage already receives the next tile), else both
  const int stage_elems = (P.sa_stride + P.sb_stride) << P.tkb;
  T *stages = reinterpret_cast<T *>(smem_raw);
  const bool xtile = stage_elems >= 4096;      // cross-tile prefetch: the staging buffer fits one stage
  // per-iteration parts of the tile enumerations (bits 8.. of the element index)
  __shared__ long long a_it_g[16], b_it_g[16], c_it_g[16];
  __shared__ int a_it_s[16], b_it_s[16];
  // Staging swizzle.  A thread owns outputs m = tm + i TM/4, n = tn + j TN/4; for fixed (i, j) the 16 lanes of a
  // half warp store to staging indices that differ in the bits their lane bits own, wherever the C layout puts
  // them -- often far apart (stride 4..1024 elements: up to 16-way bank conflicts, which made this kernel
  // shared-memory bound: ncu, round 2).  The low nibble of the staging index (the 8-byte slot inside a 128-byte
  // row) is XORed with a linear function G of the bits above it, chosen so that the four lane-owned bits land on
  // four different nibble bits: conflict-free stores, and the read-out (consecutive indices per half warp, their
  // high bits equal) stays conflict-free under any such G.
  __shared__ unsigned char swz_col[12];
  __shared__ unsigned char c_it_swz[16];

  const int t = threadIdx.x;
  const int TK = 1 << P.tkb;

  // ---- enumeration tables (the same for every tile) -----------------------------
  if (t == 32) {
    int pos[4];
    unsigned used = 0;
    for (int b = 0; b < 4; ++b) {
      const int sidx = b < P.tnb - 2 ? P.c_n_sidx[b] : P.c_m_sidx[b - (P.tnb - 2)];
      pos[b] = __ffs(sidx) - 1;
      if (pos[b] >= 0 && pos[b] < 4) used |= 1u << pos[b];
    }
    for (int b = 0; b < 12; ++b) swz_col[b] = 0;
    for (int b = 0; b < 4; ++b)
      if (pos[b] >= 4 && pos[b] < 12) {
        int u = 0;
        while (u < 4 && ((used >> u) & 1u)) ++u;
        if (u < 4) { swz_col[pos[b]] = (unsigned char)(1u << u); used |= 1u << u; }
      }
  }
  if (t < 16) {
    long long g = 0; int s = 0;
    for (int j = 8; j < P.na_lo; ++j)
      if ((t >> (j - 8)) & 1) {
        if (P.a_lo_pos[j] == 255) g = -1; else if (g >= 0) g |= 1ll << P.a_lo_pos[j];
        s += P.a_lo_sstr[j];
      }
    a_it_g[t] = g; a_it_s[t] = s;
    g = 0; s = 0;
    for (int j = 8; j < P.nb_lo; ++j)
      if ((t >> (j - 8)) & 1) {
        if (P.b_lo_pos[j] == 255) g = -1; else if (g >= 0) g |= 1ll << P.b_lo_pos[j];
        s += P.b_lo_sstr[j];
      }
    b_it_g[t] = g; b_it_s[t] = s;
    g = 0;
    for (int j = 8; j < P.nc_lo; ++j)
      if ((t >> (j - 8)) & 1) g |= 1ll << P.c_lo_pos[j];
    c_it_g[t] = g;
  }
  long long a_t_g = 0, b_t_g = 0, c_t_g = 0;
  int a_t_s = 0, b_t_s = 0;
  for (int j = 0; j < 8; ++j) {
    if (!((t >> j) & 1)) continue;
    if (j < P.na_lo) {
      if (P.a_lo_pos[j] == 255) a_t_g = -1; else if (a_t_g >= 0) a_t_g |= 1ll << P.a_lo_pos[j];
      a_t_s += P.a_lo_sstr[j];
    }
    if (j < P.nb_lo) {
      if (P.b_lo_pos[j] == 255) b_t_g = -1; else if (b_t_g >= 0) b_t_g |= 1ll << P.b_lo_pos[j];
      b_t_s += P.b_lo_sstr[j];
    }
    if (j < P.nc_lo) c_t_g |= 1ll << P.c_lo_pos[j];
  }
  const bool a_act = P.na_lo >= 8 || t < (1 << P.na_lo);
  const bool b_act = P.nb_lo >= 8 || t < (1 << P.nb_lo);
  const int a_nit = P.na_lo > 8 ? 1 << (P.na_lo - 8) : 1;
  const int b_nit = P.nb_lo > 8 ? 1 << (P.nb_lo - 8) : 1;

  // ---- thread's 4x4 outputs: staging indices -----------------------------------
  const int tn = t & ((1 << (P.tnb - 2)) - 1);
  const int tm = t >> (P.tnb - 2);
  const int qm = 1 << (P.tmb - 2), qn = 1 << (P.tnb - 2);          // a thread's rows / columns are qm / qn apart
  int mi[4], nj[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = tm + i * qm, n = tn + i * qn;
    int sm = 0, sn = 0;
    for (int b = 0; b < P.tmb; ++b) if ((m >> b) & 1) sm += P.c_m_sidx[b];
    for (int b = 0; b < P.tnb; ++b) if ((n >> b) & 1) sn += P.c_n_sidx[b];
    mi[i] = sm; nj[i] = sn;
  }
  const int c_count = 1 << P.nc_lo;
  const int k_iters_bits = P.n_khi - P.split_bits;
  const unsigned long long k_iters = 1ull << k_iters_bits;
  const R sgn_a = P.conj_a ? (R)-1 : (R)1, sgn_b = P.conj_b ? (R)-1 : (R)1;
  const bool any_conj = P.conj_a || P.conj_b;
  const long long a_slice = slice_offset(slice_id, P.n_slice_a, P.slice_bit_a, P.slice_pos_a);
  const long long b_slice = slice_offset(slice_id, P.n_slice_b, P.slice_bit_b, P.slice_pos_b);
  const bool acc_out = P.accumulate && !P.split_bits;

  // ---- a persistent CTA walks tiles blockIdx.x, + gridDim.x, ... -------------------
  // Copies (cp.async, one per element, ascending address order of the operand, zero fill for phantom rows) run
  // one k step ahead of the multiplication -- across tile boundaries too, so the first step of the next tile
  // arrives while this tile is multiplied and stored.  Conjugation flags are applied when the values are read.
  long long a_k = 0, b_k = 0;                  // operand offsets of the step to be COPIED next
  auto enter = [&](unsigned long long tile) {
    unsigned long long id = tile;
    const unsigned split = (unsigned)(id & ((1ull << P.split_bits) - 1));
    id >>= P.split_bits;
    const unsigned long long mh = id & ((1ull << P.n_mhi) - 1);
    id >>= P.n_mhi;
    const unsigned long long nh = id & ((1ull << P.n_nhi) - 1);
    id >>= P.n_nhi;
    const unsigned long long kh0 = (unsigned long long)split << k_iters_bits;
    a_k = a_slice | deposit(mh, P.a_mhi, P.n_mhi) | deposit(id, P.a_bat, P.n_batch) | deposit(kh0, P.a_khi, P.n_khi);
    b_k = b_slice | deposit(nh, P.b_nhi, P.n_nhi) | deposit(id, P.b_bat, P.n_batch) | deposit(kh0, P.b_khi, P.n_khi);
  };
  auto issue = [&](int buf) {                  // this thread's copies of the step at (a_k, b_k) into stage `buf`
    T *sA = stages + (size_t)buf * stage_elems;
    T *sB = sA + (P.sa_stride << P.tkb);
    if (a_act) {
      for (int it = 0; it < a_nit; ++it) {
        const long long g = a_it_g[it];
        const bool ok = a_t_g >= 0 && g >= 0;
        cp_async_elem<(int)sizeof(T)>(sA + a_t_s + a_it_s[it], A + (ok ? a_k + (a_t_g | g) : 0), ok);
      }
    }
    if (b_act) {
      for (int it = 0; it < b_nit; ++it) {
        const long long g = b_it_g[it];
        const bool ok = b_t_g >= 0 && g >= 0;
        cp_async_elem<(int)sizeof(T)>(sB + b_t_s + b_it_s[it], B + (ok ? b_k + (b_t_g | g) : 0), ok);
      }
    }
  };
  auto next_k = [&](unsigned long long kk) {   // flip the trailing bits that change between kk and kk+1
    const unsigned long long flip = kk ^ (kk + 1);
    const int nflip = 64 - __clzll(flip);
    for (int i = 0; i < nflip && i < k_iters_bits; ++i) {
      a_k ^= 1ll << P.a_khi[i];
      b_k ^= 1ll << P.b_khi[i];
    }
  };

  const unsigned long long ntiles = (unsigned long long)P.grid;
  unsigned long long tile = blockIdx.x;
  __syncthreads();                             // the iteration tables are in shared memory
  auto swz_of = [&](int x) {                   // G(x): XOR of the columns of the set bits 4..11
    int g = 0;
    for (int b = 4; b < 12; ++b) if ((x >> b) & 1) g ^= swz_col[b];
    return g;
  };
#pragma unroll
  for (int i = 0; i < 4; ++i) {                // staging index (12 bits) and its swizzle nibble share a register
    mi[i] |= swz_of(mi[i]) << 12;
    nj[i] |= swz_of(nj[i]) << 12;
  }
  const int g_t = swz_of(t);
  if (t < 16) c_it_swz[t] = (unsigned char)swz_of(t << 8);
  __syncthreads();
  if (tile < ntiles) { enter(tile); issue(0); }
  asm volatile("cp.async.commit_group;" ::: "memory");
  int buf = 0;                                 // stage holding the step to be multiplied next
  for (; tile < ntiles; tile += gridDim.x) {
    T acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = cx_zero<R>();
    const unsigned long long next_tile = tile + gridDim.x;
    bool pending_next = false;                 // the next tile's first step still has to be issued (no cross-tile room)
    for (unsigned long long kk = 0; kk < k_iters; ++kk) {
      // one group per step, possibly empty: "all but the newest group complete" = this step has landed
      if (kk + 1 < k_iters) { next_k(kk); issue(buf ^ 1); }
      else if (next_tile < ntiles) {
        if (xtile) { enter(next_tile); issue(buf ^ 1); } else pending_next = true;
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 1;" ::: "memory");
      __syncthreads();
      const T *pa = stages + (size_t)buf * stage_elems + tm;
      const T *pb = stages + (size_t)buf * stage_elems + (P.sa_stride << P.tkb) + tn;
      if (any_conj) {
#pragma unroll 4
        for (int k = 0; k < TK; ++k) {
          T a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) { a[i] = pa[i * qm]; a[i].y *= sgn_a; }
#pragma unroll
          for (int j = 0; j < 4; ++j) { b[j] = pb[j * qn]; b[j].y *= sgn_b; }
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) cfma<R>(acc[i][j], a[i], b[j]);
          pa += P.sa_stride;
          pb += P.sb_stride;
        }
      } else {
#pragma unroll 4
        for (int k = 0; k < TK; ++k) {
          T a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = pa[i * qm];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = pb[j * qn];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) cfma<R>(acc[i][j], a[i], b[j]);
          pa += P.sa_stride;
          pb += P.sb_stride;
        }
      }
      __syncthreads();
      buf ^= 1;
    }

    // ---- epilogue: registers -> staging (C address order) -> global ---------------
    // the stage consumed last (buf ^ 1) is free; without cross-tile prefetch nothing is in flight and the
    // staging buffer may span both stages
    T *sC = xtile ? stages + (size_t)(buf ^ 1) * stage_elems : stages;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int idx = (mi[i] & 0xfff) + (nj[j] & 0xfff);
        if (idx < c_count) sC[idx ^ ((mi[i] ^ nj[j]) >> 12)] = acc[i][j];
      }
    __syncthreads();
    {
      unsigned long long id = tile;
      const unsigned split = (unsigned)(id & ((1ull << P.split_bits) - 1));
      id >>= P.split_bits;
      const unsigned long long mh = id & ((1ull << P.n_mhi) - 1);
      id >>= P.n_mhi;
      const unsigned long long nh = id & ((1ull << P.n_nhi) - 1);
      id >>= P.n_nhi;
      const long long c_base = deposit(mh, P.c_mhi, P.n_mhi) | deposit(nh, P.c_nhi, P.n_nhi) |
                               deposit(id, P.c_bat, P.n_batch);
      T *out = C;
      long long out_off = c_base;
      if (P.split_bits) { out = ws; out_off += (long long)split * P.c_elems; }
#pragma unroll 4
      for (int e = t, it = 0; e < c_count; e += 256, ++it) {
        const long long g = out_off + (c_t_g | c_it_g[it]);
        T v = sC[e ^ g_t ^ c_it_swz[it]];
        if (acc_out) { const T o = out[g]; v.x += o.x; v.y += o.y; }
        out[g] = v;
      }
    }
    __syncthreads();                           // the staging buffer becomes a stage again
    if (pending_next) {                        // complex128: the next tile's first step, into the stage `buf`
      enter(next_tile);
      issue(buf);
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// C (+)= sum over split-K partials, elementwise (C dense: c_elems contiguous elements)
template <typename R>
__global__ void __launch_bounds__(256)
splitk_sum_kernel(cx<R> *__restrict__ C, const cx<R> *__restrict__ ws, long long c_elems, int nsplit,
                  int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c_elems) return;
  cx<R> s = cx_zero<R>();
  for (int k = 0; k < nsplit; ++k) { const cx<R> v = ws[k * c_elems + i]; s.x += v.x; s.y += v.y; }
  if (accumulate) { const cx<R> o = C[i]; s.x += o.x; s.y += o.y; }
  C[i] = s;
}

// --------------------------------------------------------------------------
// K-parallel reduction kernel: outputs <= 16, long contracted extent.
// Each thread walks k = base + t + 256*it (logical k order = ascending
// min(address bit in A, address bit in B)), accumulating all M*N outputs.
// --------------------------------------------------------------------------
template <typename R, int MB, int NB>
__global__ void __launch_bounds__(256)
gett_reduce_kernel(const __grid_constant__ qtn_pair_plan_t P, const cx<R> *__restrict__ A,
                   const cx<R> *__restrict__ B, cx<R> *__restrict__ ws,
                   const int32_t *__restrict__ slice_id) {
  using T = cx<R>;
  constexpr int M = 1 << MB, N = 1 << NB;
  __shared__ T red[8][M * N];
  const int t = threadIdx.x;
  const int chunk = P.red_chunk_bits;
  long long a_m[M], b_n[N];
#pragma unroll
  for (int m = 0; m < M; ++m) a_m[m] = deposit(m, P.a_mhi, MB);
#pragma unroll
  for (int n = 0; n < N; ++n) b_n[n] = deposit(n, P.b_nhi, NB);
  const long long a_base = slice_offset(slice_id, P.n_slice_a, P.slice_bit_a, P.slice_pos_a);
  const long long b_base = slice_offset(slice_id, P.n_slice_b, P.slice_bit_b, P.slice_pos_b);
  const int tb = chunk < 8 ? chunk : 8;   // bits of k carried by the thread id
  const unsigned long long k0 = ((unsigned long long)blockIdx.x << chunk) | (unsigned)t;
  long long a_k = a_base | deposit(k0, P.a_khi, P.n_khi);
  long long b_k = b_base | deposit(k0, P.b_khi, P.n_khi);
  T acc[M * N];
#pragma unroll
  for (int i = 0; i < M * N; ++i) acc[i] = cx_zero<R>();
  const int it_bits = chunk - tb;
  const unsigned iters = (t < (1 << tb)) ? (1u << it_bits) : 0u;
  for (unsigned it = 0; it < iters; ++it) {
    T a[M], b[N];
#pragma unroll
    for (int m = 0; m < M; ++m) { a[m] = A[a_k | a_m[m]]; if (P.conj_a) a[m].y = -a[m].y; }
#pragma unroll
    for (int n = 0; n < N; ++n) { b[n] = B[b_k | b_n[n]]; if (P.conj_b) b[n].y = -b[n].y; }
#pragma unroll
    for (int m = 0; m < M; ++m)
#pragma unroll
      for (int n = 0; n < N; ++n) cfma<R>(acc[m * N + n], a[m], b[n]);
    const unsigned flip = it ^ (it + 1);
    const int nflip = 32 - __clz(flip);
    for (int i = 0; i < nflip && i < it_bits; ++i) {
      a_k ^= 1ll << P.a_khi[tb + i];
      b_k ^= 1ll << P.b_khi[tb + i];
    }
  }
#pragma unroll
  for (int i = 0; i < M * N; ++i) {
    R x = acc[i].x, y = acc[i].y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      x += __shfl_xor_sync(0xffffffffu, x, o);
      y += __shfl_xor_sync(0xffffffffu, y, o);
    }
    if ((t & 31) == 0) { red[t >> 5][i].x = x; red[t >> 5][i].y = y; }
  }
  __syncthreads();
  if (t < M * N) {
    T s = cx_zero<R>();
#pragma unroll
    for (int w = 0; w < 8; ++w) { s.x += red[w][t].x; s.y += red[w][t].y; }
    ws[(long long)blockIdx.x * 16 + t] = s;
  }
}

// second stage: one CTA sums the per-CTA partials and writes the (<=16) outputs
template <typename R>
__global__ void __launch_bounds__(256)
reduce_final_kernel(const __grid_constant__ qtn_pair_plan_t P, cx<R> *__restrict__ C,
                    const cx<R> *__restrict__ ws, long long nparts) {
  using T = cx<R>;
  __shared__ T red[8][16];
  const int t = threadIdx.x;
  const int MN = 1 << (P.n_mhi + P.n_nhi);
  T s[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) s[i] = cx_zero<R>();
  for (long long p = t; p < nparts; p += 256)
#pragma unroll
    for (int i = 0; i < 16; ++i)
      if (i < MN) { const T v = ws[p * 16 + i]; s[i].x += v.x; s[i].y += v.y; }
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    R x = s[i].x, y = s[i].y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      x += __shfl_xor_sync(0xffffffffu, x, o);
      y += __shfl_xor_sync(0xffffffffu, y, o);
    }
    if ((t & 31) == 0) { red[t >> 5][i].x = x; red[t >> 5][i].y = y; }
  }
  __syncthreads();
  if (t < MN) {
    T v = cx_zero<R>();
    for (int w = 0; w < 8; ++w) { v.x += red[w][t].x; v.y += red[w][t].y; }
    // acc index = m * N + n
    const int n = t & ((1 << P.n_nhi) - 1), m = t >> P.n_nhi;
    const long long g = deposit(m, P.c_mhi, P.n_mhi) | deposit(n, P.c_nhi, P.n_nhi);
    if (P.accumulate) { const T o = C[g]; v.x += o.x; v.y += o.y; }
    C[g] = v;
  }
}

// --------------------------------------------------------------------------
// K-reduction kernel: at most 16 x 16 outputs, long contracted extent (inner-product-like
// nodes: the closing contraction of two big halves, expectation sandwiches).
// A CTA streams its share of K in chunks of 2^tkb (<= 64) values: each chunk of A (16 x chunk)
// and B is gathered in ascending address order into shared memory (phantom rows = 0); thread
// (kg, sb) accumulates the 4 x 4 output block sb over the k values kg, kg+16, ... of the chunk.
// At the end the 16 k-groups are summed through shared memory and the CTA writes ONE partial
// C (256 values) to the workspace; qtn_launch_splitk_sum adds the partials.
// HBM-bound by design: 16 KiB of operands per 64K real FMAs.
// --------------------------------------------------------------------------
constexpr int KRED_ROW = 16;        // rows of the A / B chunk in shared memory: element (k, row) of a chunk sits at
                                    // k * 16 + (row ^ 2 (k & 7)).  The XOR spreads lanes that walk k over the banks
                                    // (a padded row stride of 18 left 3 of 4 store wavefronts conflicted: ncu, round 1)
                                    // and keeps row pairs adjacent and 16-byte aligned for the 128-bit loads.
__device__ __forceinline__ int kred_swz(int idx) { return idx ^ (((idx >> 4) & 7) << 1); }

template <typename R>
__global__ void __launch_bounds__(256, 2)
gett_kred_kernel(const __grid_constant__ qtn_pair_plan_t P, const cx<R> *__restrict__ A,
                 const cx<R> *__restrict__ B, cx<R> *__restrict__ ws,
                 const int32_t *__restrict__ slice_id) {
  using T = cx<R>;
  constexpr int NIT = sizeof(R) == 4 ? 8 : 4;                // loads per thread, operand and chunk
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int KC = 1 << P.tkb;                                 // k values per chunk
  const int QI = P.kred_inner, BL = 1 << QI;                 // inner batch values: chunk row = (k << QI) + bl
  const int ROWS = KC << QI;
  T *sA = reinterpret_cast<T *>(smem_raw);                   // [2][ROWS][16]
  T *sB = sA + 2 * ROWS * KRED_ROW;                          // [2][ROWS][16]
  const int t = threadIdx.x;
  const long long a_base = slice_offset(slice_id, P.n_slice_a, P.slice_bit_a, P.slice_pos_a) |
                           deposit_present(blockIdx.y, P.a_bat, P.n_batch);   // grid.y = batch (hyper) value
  const long long b_base = slice_offset(slice_id, P.n_slice_b, P.slice_bit_b, P.slice_pos_b) |
                           deposit_present(blockIdx.y, P.b_bat, P.n_batch);
  // enumeration: element index e = t + 256 * u.  Bits 0-7 come from the thread (a_t_*), bits 8-10
  // from u: their global / shared strides are three per-operand constants, so the offset of
  // iteration u is a sum of at most three registers (u is a compile-time constant below).
  long long a_t_g = 0, b_t_g = 0;
  int a_t_s = 0, b_t_s = 0;
  for (int j = 0; j < 8; ++j) {
    if (!((t >> j) & 1)) continue;
    if (j < P.na_lo) { a_t_g |= 1ll << P.a_lo_pos[j]; a_t_s += P.a_lo_sstr[j]; }
    if (j < P.nb_lo) { b_t_g |= 1ll << P.b_lo_pos[j]; b_t_s += P.b_lo_sstr[j]; }
  }
  long long ag[3], bg[3];
  int as_[3], bs_[3];
  for (int j = 0; j < 3; ++j) {
    ag[j] = 8 + j < P.na_lo ? 1ll << P.a_lo_pos[8 + j] : 0;
    as_[j] = 8 + j < P.na_lo ? P.a_lo_sstr[8 + j] : 0;
    bg[j] = 8 + j < P.nb_lo ? 1ll << P.b_lo_pos[8 + j] : 0;
    bs_[j] = 8 + j < P.nb_lo ? P.b_lo_sstr[8 + j] : 0;
  }
  const bool a_act = P.na_lo >= 8 || t < (1 << P.na_lo);
  const bool b_act = P.nb_lo >= 8 || t < (1 << P.nb_lo);
  const int a_nit = P.na_lo > 8 ? 1 << (P.na_lo - 8) : 1;     // <= NIT
  const int b_nit = P.nb_lo > 8 ? 1 << (P.nb_lo - 8) : 1;

  // The CTA walks GROUPS of 2^IB consecutive chunks: group g = blockIdx.x, + gridDim.x, ... so that
  // CTAs in flight stay on neighbouring addresses; inside a group the chunk offset advances by
  // flipping address bits, the full bit-deposit is paid once per group.
  const int IB = P.n_khi < 3 ? P.n_khi : 3;
  const unsigned long long ngroups = 1ull << (P.n_khi - IB);
  unsigned long long group = blockIdx.x;
  unsigned inner = 0;
  long long a_k = 0, b_k = 0;
  auto enter_group = [&]() {
    a_k = a_base | deposit(group, P.a_khi + IB, P.n_khi - IB);
    b_k = b_base | deposit(group, P.b_khi + IB, P.n_khi - IB);
    inner = 0;
  };
  auto advance = [&]() {                        // -> next chunk of the stream (or past the end)
    if (inner + 1 == (1u << IB)) { group += gridDim.x; if (group < ngroups) enter_group(); return; }
    const unsigned flip = inner ^ (inner + 1);
    const int nflip = 32 - __clz(flip);
    for (int i = 0; i < nflip; ++i) { a_k ^= 1ll << P.a_khi[i]; b_k ^= 1ll << P.b_khi[i]; }
    ++inner;
  };

  // thread = (4 x 4 output block sb, inner batch value bl, k group kg): 16 >> QI k groups share a chunk
  const int sb = t & 15, bl = (t >> 4) & (BL - 1), kg = t >> (4 + QI);
  const int kgroups = 16 >> QI;
  const int mi = (sb >> 2) * 4, ni = (sb & 3) * 4;
  T acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = cx_zero<R>();

  T va[NIT], vb[NIT];
  auto load = [&]() {
    const T *pa_g = A + (a_k | a_t_g), *pb_g = B + (b_k | b_t_g);
#pragma unroll
    for (int u = 0; u < NIT; ++u) {
      va[u] = cx_zero<R>(); vb[u] = cx_zero<R>();
      if (a_act && u < a_nit) va[u] = pa_g[((u & 1) ? ag[0] : 0) + ((u & 2) ? ag[1] : 0) + ((u & 4) ? ag[2] : 0)];
      if (b_act && u < b_nit) vb[u] = pb_g[((u & 1) ? bg[0] : 0) + ((u & 2) ? bg[1] : 0) + ((u & 4) ? bg[2] : 0)];
    }
  };
  auto stash = [&](int buf) {
    T *da = sA + buf * ROWS * KRED_ROW, *db = sB + buf * ROWS * KRED_ROW;
#pragma unroll
    for (int u = 0; u < NIT; ++u) {
      if (a_act && u < a_nit) {
        T v = va[u]; if (P.conj_a) v.y = -v.y;
        da[kred_swz(a_t_s + ((u & 1) ? as_[0] : 0) + ((u & 2) ? as_[1] : 0) + ((u & 4) ? as_[2] : 0))] = v;
      }
      if (b_act && u < b_nit) {
        T v = vb[u]; if (P.conj_b) v.y = -v.y;
        db[kred_swz(b_t_s + ((u & 1) ? bs_[0] : 0) + ((u & 2) ? bs_[1] : 0) + ((u & 4) ? bs_[2] : 0))] = v;
      }
    }
  };
  bool have = group < ngroups;
  if (have) { enter_group(); load(); advance(); stash(0); }
  __syncthreads();
  for (int buf = 0; have; buf ^= 1) {
    const bool more = group < ngroups;            // the stream position after the current chunk
    if (more) { load(); advance(); }              // next chunk in flight while this one is used
    const T *pa = sA + buf * ROWS * KRED_ROW, *pb = sB + buf * ROWS * KRED_ROW;
    for (int kq = kg; kq < KC; kq += kgroups) {
      T a[4], b[4];
      const int k = (kq << QI) + bl;            // chunk row of (k value, this thread's inner batch value)
      const int sk = (k & 7) << 1;              // this row's XOR (kred_swz)
      if constexpr (sizeof(R) == 4) {           // two complex64 per 128-bit shared-memory load
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float4 u = *reinterpret_cast<const float4 *>(pa + k * KRED_ROW + ((mi + 2 * i) ^ sk));
          const float4 v = *reinterpret_cast<const float4 *>(pb + k * KRED_ROW + ((ni + 2 * i) ^ sk));
          a[2 * i].x = u.x; a[2 * i].y = u.y; a[2 * i + 1].x = u.z; a[2 * i + 1].y = u.w;
          b[2 * i].x = v.x; b[2 * i].y = v.y; b[2 * i + 1].x = v.z; b[2 * i + 1].y = v.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = pa[k * KRED_ROW + ((mi + i) ^ sk)];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = pb[k * KRED_ROW + ((ni + j) ^ sk)];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) cfma<R>(acc[i][j], a[i], b[j]);
    }
    if (more) stash(buf ^ 1);
    __syncthreads();
    have = more;
  }
  // sum the k groups of every inner batch value: red[kg * BL + bl][m * 16 + n] aliases the chunk buffers
  T *red = reinterpret_cast<T *>(smem_raw);
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[(t >> 4) * 256 + (mi + i) * 16 + (ni + j)] = acc[i][j];
  __syncthreads();
  {
    const int m = t >> 4, n = t & 15;
    const bool live = m < (1 << P.m_real) && n < (1 << P.n_real);
    const long long off = deposit(m, P.c_mhi, P.m_real) | deposit(n, P.c_nhi, P.n_real) |
                          deposit(blockIdx.y, P.c_bat, P.n_batch);
    for (int b = 0; b < BL; ++b) {
      T s = cx_zero<R>();
      for (int g = 0; g < kgroups; ++g) { const T v = red[(g * BL + b) * 256 + t]; s.x += v.x; s.y += v.y; }
      if (live) ws[(long long)blockIdx.x * P.c_elems + (off | deposit(b, P.kred_c_in, QI))] = s;
    }
  }
}

template <typename R>
int launch_tile(const qtn_pair_plan_t &P, const void *a, const void *b, void *c, void *ws,
                const int32_t *slice_id, cudaStream_t st) {
  using T = cx<R>;
  static int ctas_max = 0;
  if (!ctas_max) {
    QTN_CUDA_CHECK(cudaFuncSetAttribute(gett_tile_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        2 * 4608 * (int)sizeof(float2)));
    int dev = 0, sms = 0;
    QTN_CUDA_CHECK(cudaGetDevice(&dev));
    QTN_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    ctas_max = (sizeof(R) == 4 ? 3 : 2) * sms; // resident CTAs per SM (registers, ~68 KB of shared memory each)
  }
  if (P.grid <= 0 || P.grid > 0x7fffffffLL) return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_pair_run: grid too large");
  // persistent CTAs: the tile-independent set-up is paid once per CTA and copies run ahead across tiles
  const long long ctas = P.grid < ctas_max ? P.grid : ctas_max;
  gett_tile_kernel<R><<<(unsigned)ctas, 256, P.smem_bytes, st>>>(
      P, (const T *)a, (const T *)b, (T *)c, (T *)ws, slice_id);
  QTN_LAUNCH_CHECK();
  if (P.split_bits) {
    const long long n = P.c_elems;
    splitk_sum_kernel<R><<<(unsigned)((n + 255) / 256), 256, 0, st>>>((T *)c, (const T *)ws, n,
                                                                     1 << P.split_bits, P.accumulate);
    QTN_LAUNCH_CHECK();
  }
  return QTN_OK;
}

template <typename R>
int launch_kred(const qtn_pair_plan_t &P, const void *a, const void *b, void *c, void *ws,
                const int32_t *slice_id, cudaStream_t st) {
  using T = cx<R>;
  static bool attr_set = false;
  if (!attr_set) {
    QTN_CUDA_CHECK(cudaFuncSetAttribute(gett_kred_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        80 * 1024));
    attr_set = true;
  }
  gett_kred_kernel<R><<<dim3((unsigned)P.grid, 1u << P.n_batch), 256, P.smem_bytes, st>>>(
      P, (const T *)a, (const T *)b, (T *)ws, slice_id);
  QTN_LAUNCH_CHECK();
  return qtn_launch_splitk_sum(P.dtype, c, ws, P.c_elems, (int)P.grid, P.accumulate, st);
}

template <typename R, int MB, int NB>
int launch_reduce_mn(const qtn_pair_plan_t &P, const void *a, const void *b, void *c, void *ws,
                     const int32_t *slice_id, cudaStream_t st) {
  using T = cx<R>;
  gett_reduce_kernel<R, MB, NB><<<(unsigned)P.grid, 256, 0, st>>>(P, (const T *)a, (const T *)b,
                                                                  (T *)ws, slice_id);
  QTN_LAUNCH_CHECK();
  reduce_final_kernel<R><<<1, 256, 0, st>>>(P, (T *)c, (const T *)ws, P.grid);
  QTN_LAUNCH_CHECK();
  return QTN_OK;
}

template <typename R>
int launch_reduce(const qtn_pair_plan_t &P, const void *a, const void *b, void *c, void *ws,
                  const int32_t *slice_id, cudaStream_t st) {
  const int key = P.n_mhi * 8 + P.n_nhi;
  switch (key) {
#define QTN_CASE(MB, NB) \
  case MB * 8 + NB: return launch_reduce_mn<R, MB, NB>(P, a, b, c, ws, slice_id, st);
    QTN_CASE(0, 0) QTN_CASE(1, 0) QTN_CASE(2, 0) QTN_CASE(3, 0) QTN_CASE(4, 0)
    QTN_CASE(1, 1) QTN_CASE(2, 1) QTN_CASE(3, 1) QTN_CASE(2, 2)
#undef QTN_CASE
    default: return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_pair_run: reduce kernel needs m >= n, m + n <= 4");
  }
}

}  // namespace

int qtn_launch_splitk_sum(int dtype, void *c, const void *ws, long long n, int nsplit, int accumulate,
                          cudaStream_t st) {
  const unsigned grid = (unsigned)((n + 255) / 256);
  if (dtype == QTN_C64)
    splitk_sum_kernel<float><<<grid, 256, 0, st>>>((float2 *)c, (const float2 *)ws, n, nsplit, accumulate);
  else
    splitk_sum_kernel<double><<<grid, 256, 0, st>>>((double2 *)c, (const double2 *)ws, n, nsplit, accumulate);
  QTN_LAUNCH_CHECK();
  return QTN_OK;
}

// max(|re|, |im|) over a complex64 vector as an fp32 bit pattern, atomically merged into *amax
// (non-negative floats order like unsigned integers).  Streaming, 16-byte loads, grid-stride.
__global__ void __launch_bounds__(256) absmax_kernel(const float4 *__restrict__ x, long long n4, const float2 *__restrict__ tail,
                                                     int ntail, uint32_t *__restrict__ amax) {
  uint32_t m = 0;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = __ldg(x + i);
    m = max(max(m, __float_as_uint(v.x) & 0x7fffffffu), max(__float_as_uint(v.y) & 0x7fffffffu,
            max(__float_as_uint(v.z) & 0x7fffffffu, __float_as_uint(v.w) & 0x7fffffffu)));
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < ntail) {
    const float2 v = tail[threadIdx.x];
    m = max(m, max(__float_as_uint(v.x) & 0x7fffffffu, __float_as_uint(v.y) & 0x7fffffffu));
  }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicMax(amax, m);
}

int qtn_launch_absmax(const void *x, long long n, uint32_t *amax, cudaStream_t st) {
  if (n <= 0) return QTN_OK;
  const long long n4 = n / 2;                       // float4 = two complex64 values
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  absmax_kernel<<<(unsigned)blocks, 256, 0, st>>>((const float4 *)x, n4, (const float2 *)x + 2 * n4, (int)(n - 2 * n4), amax);
  QTN_LAUNCH_CHECK();
  return QTN_OK;
}

extern "C" int qtn_absmax(int dtype, int64_t n, const void *x, uint32_t *amax, void *stream) {
  if (dtype != QTN_C64) return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_absmax: complex64 only");
  if (n < 0 || (n > 0 && !x) || !amax) return qtn_fail(QTN_ERR_ARG, "qtn_absmax: bad argument");
  return qtn_launch_absmax(x, n, amax, (cudaStream_t)stream);
}

extern "C" int qtn_pair_run_scaled(const qtn_pair_plan_t *plan, const void *a, const void *b, void *c,
                                   void *workspace, const int32_t *slice_id, const uint32_t *amax_a,
                                   const uint32_t *amax_b, uint32_t *amax_c, void *stream) {
  if (!plan || !a || !b || !c) return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: null argument");
  if (plan->workspace_bytes > 0 && !workspace)
    return qtn_fail(QTN_ERR_WORKSPACE, "qtn_pair_run: plan needs a workspace");
  if ((plan->n_slice_a || plan->n_slice_b) && !slice_id)
    return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: plan has sliced labels but slice_id is NULL");
  cudaStream_t st = (cudaStream_t)stream;
  if (plan->swapped) {
    const void *tmp = a; a = b; b = tmp;
    const uint32_t *tm = amax_a; amax_a = amax_b; amax_b = tm;
  }
  const bool f32 = plan->dtype == QTN_C64;
  if (!f32) amax_c = nullptr;                       // magnitude words are a complex64 device
  if (plan->kernel == QTN_KERNEL_TC) {
    if (plan->tc_kind == 1 && (!amax_a || !amax_b)) {
      // a caller without magnitude words (plain qtn_pair_run): measure both operands here, into the two
      // words the planner reserved at the end of the workspace -- correct, at the price of one extra
      // pass over A and B
      uint32_t *words = reinterpret_cast<uint32_t *>((unsigned char *)workspace + plan->workspace_bytes - 256);
      QTN_CUDA_CHECK(cudaMemsetAsync(words, 0, 8, st));
      long long ea = 1, eb = 1;
      ea <<= (plan->tmb + plan->tkb + plan->n_mhi + plan->n_khi + plan->n_slice_a);
      eb <<= (plan->tnb + plan->tkb + plan->n_nhi + plan->n_khi + plan->n_slice_b);
      int rc = qtn_launch_absmax(a, ea, words, st);
      if (rc != QTN_OK) return rc;
      rc = qtn_launch_absmax(b, eb, words + 1, st);
      if (rc != QTN_OK) return rc;
      amax_a = words; amax_b = words + 1;
    }
    return qtn_launch_tc(*plan, a, b, c, workspace, slice_id, amax_a, amax_b, amax_c, st);
  }
  int rc;
  if (plan->kernel == QTN_KERNEL_TILE)
    rc = f32 ? launch_tile<float>(*plan, a, b, c, workspace, slice_id, st)
             : launch_tile<double>(*plan, a, b, c, workspace, slice_id, st);
  else if (plan->kernel == QTN_KERNEL_KRED)
    rc = f32 ? launch_kred<float>(*plan, a, b, c, workspace, slice_id, st)
             : launch_kred<double>(*plan, a, b, c, workspace, slice_id, st);
  else if (plan->kernel == QTN_KERNEL_REDUCE)
    rc = f32 ? launch_reduce<float>(*plan, a, b, c, workspace, slice_id, st)
             : launch_reduce<double>(*plan, a, b, c, workspace, slice_id, st);
  else
    return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: unknown kernel id");
  if (rc != QTN_OK) return rc;
  if (amax_c) return qtn_launch_absmax(c, plan->c_elems, amax_c, st);
  return QTN_OK;
}

extern "C" int qtn_pair_run(const qtn_pair_plan_t *plan, const void *a, const void *b, void *c,
                            void *workspace, const int32_t *slice_id, void *stream) {
  return qtn_pair_run_scaled(plan, a, b, c, workspace, slice_id, nullptr, nullptr, nullptr, stream);
}
