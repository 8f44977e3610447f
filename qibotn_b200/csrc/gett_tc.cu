// Pairwise contraction of complex64 bit tensors on the 5th-generation tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM), permutation fused into the loads
// and stores exactly as in gett.cu.
//
// Arithmetic.  A complex GEMM C = A * B is four real GEMMs on separate re/im planes
//     Cr = Ar*Br - Ai*Bi        Ci = Ar*Bi + Ai*Br
// (the minus sign is the MMA's negate-A bit), and every fp32 operand x is split as
// x = hi + lo with hi = tf32_round_nearest(x), lo = x - hi (exact in fp32), keeping
// hi*hi + hi*lo + lo*hi: the classic error-compensated "3xTF32" product, whose
// relative error per term is ~2^-21 (dropped lo*lo and the truncation of lo to tf32).
// So one complex MAC costs 12 tf32 MACs; the 8 real flops it is credited with
//  do not depend on that.
//
// Structure of a CTA (288 threads):
//   warps 0-7  producers: gather one K-chunk of the A and B tiles from global memory in
//              ascending *address* order (coalesced whatever the label order), split to
//              the 4+4 planes and scatter them into shared memory in the tensor core's
//              canonical K-major no-swizzle layout (8-row x 16-byte core matrices);
//              fence.proxy.async + mbarrier arrive per stage;  afterwards the same
//              warps are the epilogue: tcgen05.ld of Cr/Ci, interleave, coalesced stores;
//   warp 8     allocates TMEM; lane 0 issues the tcgen05.mma's of each stage and
//              releases the stage with tcgen05.commit.
// C's layout is chosen by the planner (m bits 0-4 | n bits | m bits 5-6 | tile index),
// which makes the epilogue's stores contiguous straight out of the TMEM registers.
#include <cuda_fp16.h>

#include <cstdint>

#include "qtn_internal.h"

namespace {

constexpr int kProducerThreads = 256;
constexpr int kThreads = 288;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug must surface as a launch error, never as a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  unsigned spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) __trap();
  }
}

// one arrival that also announces `bytes` of asynchronous (bulk-copy) traffic on the barrier
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// global -> shared bulk copy (the TMA engine, no tensor map): completes `bytes` on the barrier
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst_smem, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// Ampere-style asynchronous copy, 8 bytes, global -> shared (LDGSTS): no register staging
__device__ __forceinline__ void cp_async8(uint32_t dst_smem, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// arrive on `bar` once all cp.async copies this thread has issued so far have landed (no pending-count increment:
// the arrival is one of the barrier's expected count)
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
constexpr int kRawDepth = 6;   // chunks of A gathers in flight per CTA in the staggered kernel (16 KB each)

__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem),
               "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS)
               : "memory");
}

// One lane of a fully converged warp.  Code under `if (elect_one())` inside warp-uniform control
// flow lets ptxas keep descriptors in uniform registers and issue the MMAs back to back; under a
// `lane == 0` branch it wraps EVERY tcgen05.mma in an ELECT / BRA.U.ANY loop (~6 extra
// instructions and a branch per MMA), which made the issuing thread the bottleneck of the kernel.
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// D[tmem] (+)= A[smem desc] * B[smem desc]^T, tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrive when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}

// Shared-memory matrix descriptor, K-major, no swizzle: core matrix = 8 rows x 16 bytes,
// stored as 128 contiguous bytes; LBO = byte step between core matrices along K,
// SBO = byte step between 8-row groups.  (PTX ISA "tcgen05 shared memory descriptor";
// bit 46 = descriptor version 1 of sm_100.)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) |
         ((uint64_t)(sbo >> 4) << 32) | (1ull << 46);
}

// kind::tf32 instruction descriptor: D fp32, A/B tf32, both K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int n, int neg_a) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)neg_a << 13) | ((uint32_t)(n >> 3) << 17) |
         ((128u >> 4) << 24);
}

template <int N> struct tmem_ld;
template <> struct tmem_ld<16> {
  __device__ static __forceinline__ void run(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
  }
};
template <> struct tmem_ld<8> {
  __device__ static __forceinline__ void run(uint32_t taddr, uint32_t *r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7])
        : "r"(taddr));
  }
};
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// x ~ hi + lo, both representable in tf32 (10 explicit mantissa bits): hi = round-to-nearest(x),
// lo = round-to-nearest(x - hi).  x - hi is exact in fp32 but can carry 13 significant bits, which
// the tensor core would TRUNCATE; rounding it here keeps every operand exactly representable.
// The rounding is done on the bit pattern (add half an ulp of the 13 dropped bits, clear them:
// ties away from zero, like cvt.rna) -- two integer instructions, where ptxas expands
// cvt.rna.tf32.f32 into a ~15-instruction sequence on sm_100a.
__device__ __forceinline__ float round_tf32(float x) {
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split_tf32(float x, float &hi, float &lo) {
  hi = round_tf32(x);
  lo = round_tf32(x - hi);
}

// running max of |re|, |im| as fp32 bit patterns (non-negative floats order like unsigned integers)
__device__ __forceinline__ void amax_acc(uint32_t &m, float2 v) {
  m = max(m, max(__float_as_uint(v.x) & 0x7fffffffu, __float_as_uint(v.y) & 0x7fffffffu));
}
// publish a warp's running maximum (see qtn_pair_run_scaled)
__device__ __forceinline__ void amax_publish(uint32_t *amax, uint32_t m) {
  if (amax == nullptr) return;
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicMax(amax, m);
}

__device__ __forceinline__ long long deposit(unsigned long long idx, const uint8_t *pos, int n) {
  long long off = 0;
  for (int i = 0; i < n; ++i) off |= (long long)((idx >> i) & 1ull) << pos[i];
  return off;
}

// The same bit deposit computed by a whole (converged) warp: lane l places bit l, one OR-reduction
// combines them.  Used once per tile, where a per-thread loop over ~20 labels with 64-bit variable
// shifts was a third of the producers' instructions.
__device__ __forceinline__ long long deposit_warp(unsigned long long idx, const uint8_t *pos, int n) {
  const int lane = threadIdx.x & 31;
  unsigned lo = 0, hi = 0;
  for (int base = 0; base < n; base += 32) {
    const int i = base + lane;
    if (i < n && ((idx >> i) & 1ull)) {
      const unsigned long long b = 1ull << pos[i];
      lo |= (unsigned)b;
      hi |= (unsigned)(b >> 32);
    }
  }
  lo = __reduce_or_sync(0xffffffffu, lo);
  hi = __reduce_or_sync(0xffffffffu, hi);
  return (long long)(((unsigned long long)hi << 32) | lo);
}

__device__ __forceinline__ long long slice_offset(const int32_t *slice_id, int n, const uint8_t *bit,
                                                  const uint8_t *pos) {
  if (n == 0) return 0;
  const unsigned sid = (unsigned)(*slice_id);
  long long off = 0;
  for (int i = 0; i < n; ++i) off |= (long long)((sid >> bit[i]) & 1u) << pos[i];
  return off;
}

template <int TNB, int TKB>
__global__ void __launch_bounds__(kThreads, (TNB <= 6 ? 2 : 1))
gett_tc_kernel(const __grid_constant__ qtn_pair_plan_t P, const float2 *__restrict__ A,
               const float2 *__restrict__ B, float2 *__restrict__ C, float2 *__restrict__ ws,
               const int32_t *__restrict__ slice_id) {
  constexpr int TN = 1 << TNB, KT = 1 << TKB;
  constexpr int PLANE_A = 128 * KT;                 // floats per A plane
  constexpr int PLANE_B = TN * KT;                  // floats per B plane
  constexpr int STAGE = 4 * (PLANE_A + PLANE_B);    // floats per stage
  constexpr int NA = PLANE_A / kProducerThreads;    // A elements per producer thread and stage
  constexpr int NB = (PLANE_B + kProducerThreads - 1) / kProducerThreads;
  constexpr int TMEM_COLS = 2 * TN < 32 ? 32 : 2 * TN;
  constexpr uint32_t LBO = 128, SBO = KT * 32;      // bytes
  constexpr uint32_t IDESC = umma_idesc_tf32(TN, 0), IDESC_NEG = umma_idesc_tf32(TN, 1);

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float *stages = reinterpret_cast<float *>(smem_raw);
  const int S = P.tc_stages;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)S * STAGE * sizeof(float));
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * S + 1);
  __shared__ long long a_it_g[8], b_it_g[16];
  __shared__ int a_it_s[8], b_it_s[16];

  const int t = threadIdx.x;
  const int warp = t >> 5, lane = t & 31;
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  const uint32_t accum_bar = bar0 + 8u * (2 * S);

  // ---- which tile ----------------------------------------------------------------
  unsigned long long id = blockIdx.x;
  const unsigned split = (unsigned)(id & ((1ull << P.split_bits) - 1));
  id >>= P.split_bits;
  // n tiles vary fastest: CTAs that share an A tile run together and the second read hits L2
  const unsigned long long nh = id & ((1ull << P.n_nhi) - 1);
  id >>= P.n_nhi;
  const unsigned long long mh = id;
  const long long a_base = deposit(mh, P.a_mhi, P.n_mhi) |
                           slice_offset(slice_id, P.n_slice_a, P.slice_bit_a, P.slice_pos_a);
  const long long b_base = deposit(nh, P.b_nhi, P.n_nhi) |
                           slice_offset(slice_id, P.n_slice_b, P.slice_bit_b, P.slice_pos_b);
  const long long c_base = deposit(mh, P.c_mhi, P.n_mhi) | deposit(nh, P.c_nhi, P.n_nhi);
  const int k_iters_bits = P.n_khi - P.split_bits;
  const unsigned long long k_iters = 1ull << k_iters_bits;
  const unsigned long long kh0 = (unsigned long long)split << k_iters_bits;

  // ---- one-time setup ----------------------------------------------------------------
  if (t < 16) {
    long long g = 0;
    int s = 0;
    if (t < 8) {
      for (int j = 8; j < P.na_lo; ++j)
        if ((t >> (j - 8)) & 1) { g |= 1ll << P.a_lo_pos[j]; s += P.a_lo_sstr[j]; }
      a_it_g[t] = g; a_it_s[t] = s;
    }
    g = 0; s = 0;
    for (int j = 8; j < P.nb_lo; ++j)
      if ((t >> (j - 8)) & 1) { g |= 1ll << P.b_lo_pos[j]; s += P.b_lo_sstr[j]; }
    b_it_g[t] = g; b_it_s[t] = s;
  }
  if (t == 32) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), kProducerThreads);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    fence_proxy_async_smem();
    fence_mbar_init();
  }
  if (warp == 8) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ================= producers =================
    long long a_t_g = 0, b_t_g = 0;
    int a_t_s = 0, b_t_s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (!((t >> j) & 1)) continue;
      if (j < P.na_lo) { a_t_g |= 1ll << P.a_lo_pos[j]; a_t_s += P.a_lo_sstr[j]; }
      if (j < P.nb_lo) { b_t_g |= 1ll << P.b_lo_pos[j]; b_t_s += P.b_lo_sstr[j]; }
    }
    const bool b_act = P.nb_lo >= 8 || t < (1 << P.nb_lo);
    const float sgn_a = P.conj_a ? -1.f : 1.f, sgn_b = P.conj_b ? -1.f : 1.f;
    long long a_k = a_base | deposit(kh0, P.a_khi, P.n_khi);
    long long b_k = b_base | deposit(kh0, P.b_khi, P.n_khi);
    int s = 0;
    uint32_t ph = 0;
    auto load = [&](float2(&va)[NA], float2(&vb)[NB]) {
#pragma unroll
      for (int i = 0; i < NA; ++i) va[i] = __ldg(A + (a_k | a_t_g | a_it_g[i]));
#pragma unroll
      for (int i = 0; i < NB; ++i)
        vb[i] = b_act ? __ldg(B + (b_k | b_t_g | b_it_g[i])) : make_float2(0.f, 0.f);
    };
    auto advance = [&](unsigned long long kk) {          // offsets of k chunk kk -> kk + 1
      const unsigned long long flip = kk ^ (kk + 1);
      const int nflip = 64 - __clzll(flip);
      for (int i = 0; i < nflip && i < k_iters_bits; ++i) {
        a_k ^= 1ll << P.a_khi[i];
        b_k ^= 1ll << P.b_khi[i];
      }
    };
    auto store = [&](const float2(&va)[NA], const float2(&vb)[NB]) {
      mbar_wait(empty_bar(s), ph ^ 1u);
      float *sA = stages + (size_t)s * STAGE;
      float *sB = sA + 4 * PLANE_A;
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        const int o = a_t_s + a_it_s[i];
        float h, l;
        split_tf32(va[i].x, h, l);
        sA[o] = h; sA[PLANE_A + o] = l;
        split_tf32(va[i].y * sgn_a, h, l);
        sA[2 * PLANE_A + o] = h; sA[3 * PLANE_A + o] = l;
      }
      if (b_act) {
#pragma unroll
        for (int i = 0; i < NB; ++i) {
          const int o = b_t_s + b_it_s[i];
          float h, l;
          split_tf32(vb[i].x, h, l);
          sB[o] = h; sB[PLANE_B + o] = l;
          split_tf32(vb[i].y * sgn_b, h, l);
          sB[2 * PLANE_B + o] = h; sB[3 * PLANE_B + o] = l;
        }
      }
      fence_proxy_async_smem();          // generic-proxy stores -> visible to the tensor core
      mbar_arrive(full_bar(s));
      if (++s == S) { s = 0; ph ^= 1u; }
    };
    // register double buffering: the loads of chunk kk+1 are in flight while chunk kk is split
    // and stored (and before its pipeline slot is even free)
    float2 va0[NA], vb0[NB], va1[NA], vb1[NB];
    load(va0, vb0);
    advance(0);
    for (unsigned long long kk = 0; kk < k_iters; kk += 2) {
      const bool more1 = kk + 1 < k_iters;
      if (more1) { load(va1, vb1); advance(kk + 1); }
      store(va0, vb0);
      if (more1) {
        if (kk + 2 < k_iters) { load(va0, vb0); advance(kk + 2); }
        store(va1, vb1);
      }
    }

    // ================= epilogue =================
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int q = warp & 3, half = warp >> 2;
    const int r = q * 32 + lane;                     // tile row = TMEM lane
    float2 *out = P.split_bits ? ws + (long long)split * P.c_elems : C;
    out += c_base + (r & 31) + ((long long)(r >> 5) << (5 + TNB));
    constexpr int HALF = TN / 2;
    constexpr int CH = HALF < 16 ? HALF : 16;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
    for (int c0 = half * HALF; c0 < (half + 1) * HALF; c0 += CH) {
      uint32_t re[CH], im[CH];
      tmem_ld<CH>::run(trow + (uint32_t)c0, re);
      tmem_ld<CH>::run(trow + (uint32_t)(TN + c0), im);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < CH; ++j)
        out[(long long)(c0 + j) << 5] = make_float2(__uint_as_float(re[j]), __uint_as_float(im[j]));
    }
  } else if (lane == 0) {
    // ================= MMA issuer (one thread) =================
    int s = 0;
    uint32_t ph = 0;
    for (unsigned long long kk = 0; kk < k_iters; ++kk) {
      mbar_wait(full_bar(s), ph);
      tc_fence_after();
      const uint32_t sa = smem_u32(stages + (size_t)s * STAGE);
      const uint32_t sb = sa + 16u * PLANE_A;
#pragma unroll
      for (int ks = 0; ks < KT / 8; ++ks) {
        const uint32_t ko = (uint32_t)ks * 2u * LBO;
        const uint64_t arh = umma_desc(sa + ko, LBO, SBO);
        const uint64_t arl = umma_desc(sa + 4u * PLANE_A + ko, LBO, SBO);
        const uint64_t aih = umma_desc(sa + 8u * PLANE_A + ko, LBO, SBO);
        const uint64_t ail = umma_desc(sa + 12u * PLANE_A + ko, LBO, SBO);
        const uint64_t brh = umma_desc(sb + ko, LBO, SBO);
        const uint64_t brl = umma_desc(sb + 4u * PLANE_B + ko, LBO, SBO);
        const uint64_t bih = umma_desc(sb + 8u * PLANE_B + ko, LBO, SBO);
        const uint64_t bil = umma_desc(sb + 12u * PLANE_B + ko, LBO, SBO);
        const uint32_t first = (kk | (unsigned)ks) ? 1u : 0u;
        const uint32_t cr = tmem_base, ci = tmem_base + TN;
        // small cross terms first, the hi*hi products last
        umma_tf32(cr, arl, brh, IDESC, first);
        umma_tf32(cr, arh, brl, IDESC, 1u);
        umma_tf32(cr, ail, bih, IDESC_NEG, 1u);
        umma_tf32(cr, aih, bil, IDESC_NEG, 1u);
        umma_tf32(cr, arh, brh, IDESC, 1u);
        umma_tf32(cr, aih, bih, IDESC_NEG, 1u);
        umma_tf32(ci, arl, bih, IDESC, first);
        umma_tf32(ci, arh, bil, IDESC, 1u);
        umma_tf32(ci, ail, brh, IDESC, 1u);
        umma_tf32(ci, aih, brl, IDESC, 1u);
        umma_tf32(ci, arh, bih, IDESC, 1u);
        umma_tf32(ci, aih, brh, IDESC, 1u);
      }
      umma_commit(empty_bar(s));            // frees the stage once these MMAs have read it
      if (++s == S) { s = 0; ph ^= 1u; }
    }
    umma_commit(accum_bar);                 // accumulators complete
  }

  // ---- teardown ------------------------------------------------------------------------
  tc_fence_before();
  __syncthreads();
  if (warp == 8) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}


// --------------------------------------------------------------------------------------------
// Persistent variant: one CTA per SM walks tiles blockIdx.x, blockIdx.x + gridDim.x, ...
// Roles are fully decoupled so that every phase of tile i overlaps another phase of tile i+1:
//   warps 0-7   producers  (global gathers -> split -> shared memory), a flat stream over
//               (tile, k chunk) that runs ahead across tile boundaries;
//   warps 8-11  epilogue   (TMEM -> registers -> global), one TMEM lane quarter each;
//   warp 12     MMA issuer (+ TMEM allocation).
// Accumulators are double-buffered in TMEM when two fit (2^TNB <= 128).
// --------------------------------------------------------------------------------------------
constexpr int kThreadsP = 416;
constexpr int kEpilogueThreads = 128;
// the staggered kernel: 8 producer warps, 8 epilogue warps, 2 MMA-issue warps (one per half; the first one is the
// forwarder in the peer CTA)
constexpr int kThreadsP2 = 576;
constexpr int kEpilogueThreads2 = 256;
constexpr int kMmaWarp2 = 16;

template <int TNB, int TKB, bool SEP, bool PRE>
__global__ void __launch_bounds__(kThreadsP, 1)
gett_tc_persistent_kernel(const __grid_constant__ qtn_pair_plan_t P, const float2 *__restrict__ A,
                          const float2 *__restrict__ B, float2 *__restrict__ C,
                          float2 *__restrict__ ws, const int32_t *__restrict__ slice_id,
                          uint32_t *__restrict__ amax_c) {
  constexpr int TN = 1 << TNB, KT = 1 << TKB;
  constexpr int PLANE_A = 128 * KT, PLANE_B = TN * KT;
  constexpr int STAGE = 4 * (PLANE_A + PLANE_B);
  constexpr int NA = PLANE_A / kProducerThreads;
  constexpr int NB = (PLANE_B + kProducerThreads - 1) / kProducerThreads;
  // SEP: the hi*hi products and the small cross terms accumulate in SEPARATE TMEM regions and are
  // added in the epilogue.  The tensor core truncates (does not round) at every accumulation, a
  // bias proportional to the number of MMAs that hit the accumulator holding the large sum: 2 per
  // k step this way instead of 6 (measured on the full C4 amplitude: see DESIGN.md section 3.1).
  constexpr int ACC = SEP ? 4 : 2;                   // TMEM regions of TN columns per buffer
  static_assert(ACC * TN <= 512, "accumulators do not fit in TMEM");
  constexpr int NBUF = 2 * ACC * TN <= 512 ? 2 : 1;
  constexpr int TMEM_COLS = NBUF * ACC * TN < 32 ? 32 : NBUF * ACC * TN;
  constexpr uint32_t LBO = 128, SBO = KT * 32;
  constexpr uint32_t IDESC = umma_idesc_tf32(TN, 0), IDESC_NEG = umma_idesc_tf32(TN, 1);

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float *stages = reinterpret_cast<float *>(smem_raw);
  const int S = P.tc_stages;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)S * STAGE * sizeof(float));
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 2 * S + 2 * NBUF);
  const int t = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0), lane = t & 31;   // provably warp-uniform
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (2 * S + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (2 * S + NBUF + b); };

  const unsigned long long ntiles = (unsigned long long)P.grid;     // tiles incl. split-K parts
  const int k_iters_bits = P.n_khi - P.split_bits;
  const unsigned long long k_iters = 1ull << k_iters_bits;
  const long long a_slice = slice_offset(slice_id, P.n_slice_a, P.slice_bit_a, P.slice_pos_a);
  const long long b_slice = slice_offset(slice_id, P.n_slice_b, P.slice_bit_b, P.slice_pos_b);

  if (t == 32) {
    // with a pre-packed B one producer thread arrives twice: early with the byte count of the bulk
    // copy, and with everybody else once its A values are stored
    for (int s = 0; s < S; ++s) { mbar_init(full_bar(s), kProducerThreads + (PRE ? 1 : 0)); mbar_init(empty_bar(s), 1); }
    for (int b = 0; b < NBUF; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), kEpilogueThreads); }
    fence_proxy_async_smem();
    fence_mbar_init();
  }
  if (warp == 12) tmem_alloc<TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ================= producers: flat stream over (tile, k chunk) =================
    long long a_t_g = 0, b_t_g = 0;
    int a_t_s = 0, b_t_s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (!((t >> j) & 1)) continue;
      if (j < P.na_lo) { a_t_g |= 1ll << P.a_lo_pos[j]; a_t_s += P.a_lo_sstr[j]; }
      if (j < P.nb_lo) { b_t_g |= 1ll << P.b_lo_pos[j]; b_t_s += P.b_lo_sstr[j]; }
    }
    const bool b_act = P.nb_lo >= 8 || t < (1 << P.nb_lo);
    const float sgn_a = P.conj_a ? -1.f : 1.f, sgn_b = P.conj_b ? -1.f : 1.f;
    unsigned long long tile = blockIdx.x, kk = 0;
    long long a_k = 0, b_k = 0;
    constexpr bool pre = PRE;                      // B arrives as ready-made stage images
    constexpr uint32_t kImageBytes = 16u * PLANE_B;
    const unsigned char *images = reinterpret_cast<const unsigned char *>(ws) + P.tc_prepack_off;
    unsigned long long image0 = 0;                 // image index of this tile's first k chunk
    auto enter_tile = [&]() {
      unsigned long long id = tile;
      const unsigned split = (unsigned)(id & ((1ull << P.split_bits) - 1));
      id >>= P.split_bits;
      const unsigned long long nh = id & ((1ull << P.n_nhi) - 1);
      const unsigned long long mh = id >> P.n_nhi;
      const unsigned long long kh0 = (unsigned long long)split << k_iters_bits;
      a_k = a_slice | deposit_warp(mh, P.a_mhi, P.n_mhi) | deposit_warp(kh0, P.a_khi, P.n_khi);
      if constexpr (!PRE) b_k = b_slice | deposit_warp(nh, P.b_nhi, P.n_nhi) | deposit_warp(kh0, P.b_khi, P.n_khi);
      image0 = (nh << P.n_khi) | kh0;
      kk = 0;
    };
    auto advance = [&]() {                        // next k chunk, or the first one of the next tile
      if (kk + 1 == k_iters) { tile += gridDim.x; if (tile < ntiles) enter_tile(); return; }
      const unsigned long long flip = kk ^ (kk + 1);
      const int nflip = 64 - __clzll(flip);
      for (int i = 0; i < nflip && i < k_iters_bits; ++i) {
        a_k ^= 1ll << P.a_khi[i];
        b_k ^= 1ll << P.b_khi[i];
      }
      ++kk;
    };
    constexpr int NBR = PRE ? 1 : NB;              // B registers (none needed when pre-packed)
    auto load = [&](float2(&va)[NA], float2(&vb)[NBR], unsigned long long &img) {
#pragma unroll
      for (int i = 0; i < NA; ++i) va[i] = __ldg(A + (a_k | a_t_g | P.tc_a_it_g[i]));
      img = image0 + kk;
      if constexpr (!pre) {
#pragma unroll
        for (int i = 0; i < NB; ++i)
          vb[i] = b_act ? __ldg(B + (b_k | b_t_g | P.tc_b_it_g[i])) : make_float2(0.f, 0.f);
      }
    };
    int s = 0;
    uint32_t ph = 0;
    auto store = [&](const float2(&va)[NA], const float2(&vb)[NBR], unsigned long long img) {
      mbar_wait(empty_bar(s), ph ^ 1u);
      float *sA = stages + (size_t)s * STAGE;
      float *sB = sA + 4 * PLANE_A;
      if (pre && t == 0) {                          // the whole B stage: one bulk copy
        mbar_arrive_expect_tx(full_bar(s), kImageBytes);
        bulk_copy_g2s(smem_u32(sB), images + img * kImageBytes, kImageBytes, full_bar(s));
      }
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        const int o = a_t_s + P.tc_a_it_s[i];
        float h, l;
        split_tf32(va[i].x, h, l);
        sA[o] = h; sA[PLANE_A + o] = l;
        split_tf32(va[i].y * sgn_a, h, l);
        sA[2 * PLANE_A + o] = h; sA[3 * PLANE_A + o] = l;
      }
      if constexpr (!pre) {
        if (b_act) {
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            const int o = b_t_s + P.tc_b_it_s[i];
            float h, l;
            split_tf32(vb[i].x, h, l);
            sB[o] = h; sB[PLANE_B + o] = l;
            split_tf32(vb[i].y * sgn_b, h, l);
            sB[2 * PLANE_B + o] = h; sB[3 * PLANE_B + o] = l;
          }
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(full_bar(s));
      if (++s == S) { s = 0; ph ^= 1u; }
    };
    // ring of D register buffers: D-1 chunks of global loads are in flight while one chunk is
    // split and stored; deeper for the small tiles whose nodes are HBM-bound
    constexpr int D = (NA + (PRE ? 0 : NB)) <= 12 ? 4 : 2;
    float2 va[D][NA], vb[D][NBR];
    unsigned long long img[D];
    bool have[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      have[d] = tile < ntiles;
      img[d] = 0;
      if (have[d]) { if (d == 0) enter_tile(); load(va[d], vb[d], img[d]); advance(); }
    }
    bool done = false;
    while (!done) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if (done) continue;
        if (!have[d]) { done = true; continue; }
        store(va[d], vb[d], img[d]);
        have[d] = tile < ntiles;
        if (have[d]) { load(va[d], vb[d], img[d]); advance(); }
      }
    }
  } else if (warp < 12) {
    // ================= epilogue: TMEM -> global, tile after tile =================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const long long row_off = (r & 31) + ((long long)(r >> 5) << (5 + TNB));
    constexpr int CH = TN < 16 ? TN : 16;
    uint32_t cmax = 0;
    unsigned long long it = 0;
    for (unsigned long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int buf = (int)(it % NBUF);
      const uint32_t tph = (uint32_t)((it / NBUF) & 1);
      unsigned long long id = tile;
      const unsigned split = (unsigned)(id & ((1ull << P.split_bits) - 1));
      id >>= P.split_bits;
      const unsigned long long nh = id & ((1ull << P.n_nhi) - 1);
      const unsigned long long mh = id >> P.n_nhi;
      const long long c_base = deposit_warp(mh, P.c_mhi, P.n_mhi) | deposit_warp(nh, P.c_nhi, P.n_nhi);
      float2 *out = (P.split_bits ? ws + (long long)split * P.c_elems : C) + c_base + row_off;
      mbar_wait(tfull_bar(buf), tph);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * ACC * TN);
#pragma unroll 1
      for (int c0 = 0; c0 < TN; c0 += CH) {
        uint32_t re[CH], im[CH];
        tmem_ld<CH>::run(trow + (uint32_t)c0, re);
        tmem_ld<CH>::run(trow + (uint32_t)(TN + c0), im);
        if constexpr (SEP) {
          uint32_t xr[CH], xi[CH];
          tmem_ld<CH>::run(trow + (uint32_t)(2 * TN + c0), xr);
          tmem_ld<CH>::run(trow + (uint32_t)(3 * TN + c0), xi);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const float2 v = make_float2(__uint_as_float(re[j]) + __uint_as_float(xr[j]),
                                         __uint_as_float(im[j]) + __uint_as_float(xi[j]));
            amax_acc(cmax, v);
            out[(long long)(c0 + j) << 5] = v;
          }
        } else {
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const float2 v = make_float2(__uint_as_float(re[j]), __uint_as_float(im[j]));
            amax_acc(cmax, v);
            out[(long long)(c0 + j) << 5] = v;
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty_bar(buf));                 // accumulator buffer may be overwritten
    }
    if (P.split_bits == 0) amax_publish(amax_c, cmax);
  } else {
    // ================= MMA issuer: the whole warp walks the loops, one elected lane issues =================
    int s = 0;
    uint32_t ph = 0;
    unsigned long long it = 0;
    const uint32_t tmem_b = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t stage0 = smem_u32(stages);
    for (unsigned long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      const int buf = (int)(it % NBUF);
      const uint32_t tph = (uint32_t)((it / NBUF) & 1);
      mbar_wait(tempty_bar(buf), tph ^ 1u);         // epilogue has drained this buffer
      tc_fence_after();
      const uint32_t cr = tmem_b + (uint32_t)(buf * ACC * TN), ci = cr + TN;
      const uint32_t xr = SEP ? cr + 2 * TN : cr, xi = SEP ? cr + 3 * TN : ci;   // cross-term targets
      for (unsigned long long kk = 0; kk < k_iters; ++kk) {
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = stage0 + (uint32_t)s * (uint32_t)(STAGE * sizeof(float));
          const uint32_t sb = sa + 16u * PLANE_A;
#pragma unroll
          for (int ks = 0; ks < KT / 8; ++ks) {
            const uint32_t ko = (uint32_t)ks * 2u * LBO;
            const uint64_t arh = umma_desc(sa + ko, LBO, SBO);
            const uint64_t arl = umma_desc(sa + 4u * PLANE_A + ko, LBO, SBO);
            const uint64_t aih = umma_desc(sa + 8u * PLANE_A + ko, LBO, SBO);
            const uint64_t ail = umma_desc(sa + 12u * PLANE_A + ko, LBO, SBO);
            const uint64_t brh = umma_desc(sb + ko, LBO, SBO);
            const uint64_t brl = umma_desc(sb + 4u * PLANE_B + ko, LBO, SBO);
            const uint64_t bih = umma_desc(sb + 8u * PLANE_B + ko, LBO, SBO);
            const uint64_t bil = umma_desc(sb + 12u * PLANE_B + ko, LBO, SBO);
            const uint32_t acc = (kk | (unsigned)ks) ? 1u : 0u;
            const uint32_t accx = SEP ? acc : 1u;      // SEP: the cross region starts from zero too
            if constexpr (!SEP) umma_tf32(cr, arh, brh, IDESC, acc);      // hi*hi first, then the rest
            umma_tf32(xr, arl, brh, IDESC, accx);
            umma_tf32(xr, arh, brl, IDESC, 1u);
            umma_tf32(xr, ail, bih, IDESC_NEG, 1u);
            umma_tf32(xr, aih, bil, IDESC_NEG, 1u);
            if constexpr (SEP) umma_tf32(cr, arh, brh, IDESC, acc);
            umma_tf32(cr, aih, bih, IDESC_NEG, 1u);
            if constexpr (!SEP) umma_tf32(ci, arh, bih, IDESC, acc);
            umma_tf32(xi, arl, bih, IDESC, accx);
            umma_tf32(xi, arh, bil, IDESC, 1u);
            umma_tf32(xi, ail, brh, IDESC, 1u);
            umma_tf32(xi, aih, brl, IDESC, 1u);
            if constexpr (SEP) umma_tf32(ci, arh, bih, IDESC, acc);
            umma_tf32(ci, aih, brh, IDESC, 1u);
            if (P.tc_terms >= 4) {                   // lo*lo as well
              umma_tf32(xr, arl, brl, IDESC, 1u);
              umma_tf32(xr, ail, bil, IDESC_NEG, 1u);
              umma_tf32(xi, arl, bil, IDESC, 1u);
              umma_tf32(xi, ail, brl, IDESC, 1u);
            }
          }
          umma_commit(empty_bar(s));
          if (kk + 1 == k_iters) umma_commit(tfull_bar(buf));
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 12) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc<TMEM_COLS>(tmem_base);
  }
}


// --------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2): two CTAs of one cluster, resident on the two SMs of a
// TPC, work on two tiles that share the same B tile (m tile indices 2j and 2j+1).  Each CTA stages
// its own 128 rows of A and only HALF of the B rows; the leader's MMA thread issues one
// 256 x TN x 8 instruction that reads both shared memories and writes each CTA's 128 accumulator
// lanes into that CTA's TMEM.  Per SM and per MMA this is 4 KB (A) + 2 KB (B) of operand reads
// instead of 4 + 4 -- the single-CTA M = 128, N = 128 tf32 MMA needs the whole 128 B/clk of
// shared-memory bandwidth by itself, leaving nothing for the producers' stores -- and half the
// bulk-copy traffic for B.  Always split accumulators (SEP) and pre-packed B (PRE).
//
// Protocol (per CTA unless noted):
//   full[s]    256 producer arrivals + 1 expect_tx (the four half-plane bulk copies)
//   pfull[s]   (leader's) 1 arrival: the peer's forwarder thread waits the peer's full[s] and
//              arrives remotely -- the leader issues once both are complete
//   empty[s]   1 arrival in BOTH CTAs: tcgen05.commit ... multicast::cluster
//   tfull[b]   1 arrival in BOTH CTAs by the same multicast commit
//   tempty[b]  (leader's) 256 arrivals: the 128 epilogue threads of each CTA
// --------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same shared-memory offset in CTA `cta` of this cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  unsigned spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins > (1u << 24)) __trap();
  }
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t slot_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "n"(COLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void umma_tf32_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in both CTAs once the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}
__host__ __device__ constexpr uint32_t umma_idesc_tf32_m256(int n, int neg_a) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)neg_a << 13) | ((uint32_t)(n >> 3) << 17) |
         ((256u >> 4) << 24);
}

template <int TNB, int TKB>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsP, 1)
gett_tc_pair_kernel(const __grid_constant__ qtn_pair_plan_t P, const float2 *__restrict__ A,
                    float2 *__restrict__ C, float2 *__restrict__ ws, const int32_t *__restrict__ slice_id,
                    uint32_t *__restrict__ amax_c) {
  constexpr int TN = 1 << TNB, KT = 1 << TKB;
  constexpr int PLANE_A = 128 * KT, PLANE_B = TN * KT, PLANE_BH = PLANE_B / 2;
  constexpr int STAGE = 4 * (PLANE_A + PLANE_BH);
  constexpr int NA = PLANE_A / kProducerThreads;
  constexpr int ACC = 4;
  static_assert(TN >= 32 && ACC * TN <= 512, "tile does not fit the pair kernel");
  constexpr int NBUF = 2 * ACC * TN <= 512 ? 2 : 1;
  constexpr int TMEM_COLS = NBUF * ACC * TN;
  constexpr uint32_t LBO = 128, SBO = KT * 32;
  constexpr uint32_t IDESC = umma_idesc_tf32_m256(TN, 0), IDESC_NEG = umma_idesc_tf32_m256(TN, 1);

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float *stages = reinterpret_cast<float *>(smem_raw);
  const int S = P.tc_stages;
  uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)S * STAGE * sizeof(float));
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * S + 2 * NBUF);
  const int t = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0), lane = t & 31;   // provably warp-uniform
  const uint32_t rank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto pfull_bar = [&](int s) { return bar0 + 8u * (2 * S + s); };
  auto tfull_bar = [&](int b) { return bar0 + 8u * (3 * S + b); };
  auto tempty_bar = [&](int b) { return bar0 + 8u * (3 * S + NBUF + b); };

  // a "super tile" = the two tiles (mh = 2j, 2j+1) of one (split, nh, j); CTA pairs stride over them
  const unsigned long long nsuper = (unsigned long long)P.grid >> 1;
  const unsigned long long pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int k_iters_bits = P.n_khi - P.split_bits;
  const unsigned long long k_iters = 1ull << k_iters_bits;
  const long long a_slice = slice_offset(slice_id, P.n_slice_a, P.slice_bit_a, P.slice_pos_a);
  auto decode = [&](unsigned long long u, unsigned &split, unsigned long long &nh, unsigned long long &mh) {
    split = (unsigned)(u & ((1ull << P.split_bits) - 1));
    u >>= P.split_bits;
    nh = u & ((1ull << P.n_nhi) - 1);
    mh = ((u >> P.n_nhi) << 1) | rank;
  };

  if (t == 32) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), kProducerThreads + 1);
      mbar_init(empty_bar(s), 1);
      mbar_init(pfull_bar(s), 1);
    }
    for (int b = 0; b < NBUF; ++b) { mbar_init(tfull_bar(b), 1); mbar_init(tempty_bar(b), 2 * kEpilogueThreads); }
    fence_proxy_async_smem();
    fence_mbar_init();
  }
  if (warp == 12) tmem_alloc_pair<TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  cluster_sync_all();                              // barriers of both CTAs are initialised
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8) {
    // ================= producers: this CTA's A rows, its half of the B rows =================
    long long a_t_g = 0;
    int a_t_s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (!((t >> j) & 1)) continue;
      if (j < P.na_lo) { a_t_g |= 1ll << P.a_lo_pos[j]; a_t_s += P.a_lo_sstr[j]; }
    }
    const float sgn_a = P.conj_a ? -1.f : 1.f;
    unsigned long long u = pair0, kk = 0;
    long long a_k = 0;
    constexpr uint32_t kImageBytes = 16u * PLANE_B, kHalfBytes = 4u * PLANE_BH;
    const unsigned char *images = reinterpret_cast<const unsigned char *>(ws) + P.tc_prepack_off;
    unsigned long long image0 = 0;
    auto enter_tile = [&]() {
      unsigned split;
      unsigned long long nh, mh;
      decode(u, split, nh, mh);
      const unsigned long long kh0 = (unsigned long long)split << k_iters_bits;
      a_k = a_slice | deposit_warp(mh, P.a_mhi, P.n_mhi) | deposit_warp(kh0, P.a_khi, P.n_khi);
      image0 = (nh << P.n_khi) | kh0;
      kk = 0;
    };
    auto advance = [&]() {
      if (kk + 1 == k_iters) { u += npairs; if (u < nsuper) enter_tile(); return; }
      const unsigned long long flip = kk ^ (kk + 1);
      const int nflip = 64 - __clzll(flip);
      for (int i = 0; i < nflip && i < k_iters_bits; ++i) a_k ^= 1ll << P.a_khi[i];
      ++kk;
    };
    auto load = [&](float2(&va)[NA], unsigned long long &img) {
#pragma unroll
      for (int i = 0; i < NA; ++i) va[i] = __ldg(A + (a_k | a_t_g | P.tc_a_it_g[i]));
      img = image0 + kk;
    };
    int s = 0;
    uint32_t ph = 0;
    auto store = [&](const float2(&va)[NA], unsigned long long img) {
      mbar_wait(empty_bar(s), ph ^ 1u);
      float *sA = stages + (size_t)s * STAGE;
      float *sB = sA + 4 * PLANE_A;
      if (t == 0) {                                 // rows [rank*TN/2, (rank+1)*TN/2) of the four B planes
        mbar_arrive_expect_tx(full_bar(s), 4u * kHalfBytes);
        const unsigned char *src = images + img * kImageBytes + (size_t)rank * kHalfBytes;
#pragma unroll
        for (int pl = 0; pl < 4; ++pl)
          bulk_copy_g2s(smem_u32(sB + pl * PLANE_BH), src + (size_t)pl * 4u * PLANE_B, kHalfBytes, full_bar(s));
      }
#pragma unroll
      for (int i = 0; i < NA; ++i) {
        const int o = a_t_s + P.tc_a_it_s[i];
        float h, l;
        split_tf32(va[i].x, h, l);
        sA[o] = h; sA[PLANE_A + o] = l;
        split_tf32(va[i].y * sgn_a, h, l);
        sA[2 * PLANE_A + o] = h; sA[3 * PLANE_A + o] = l;
      }
      fence_proxy_async_smem();
      mbar_arrive(full_bar(s));
      if (++s == S) { s = 0; ph ^= 1u; }
    };
    constexpr int D = 4;
    float2 va[D][NA];
    unsigned long long img[D];
    bool have[D];
#pragma unroll
    for (int d = 0; d < D; ++d) {
      have[d] = u < nsuper;
      img[d] = 0;
      if (have[d]) { if (d == 0) enter_tile(); load(va[d], img[d]); advance(); }
    }
    bool done = false;
    while (!done) {
#pragma unroll
      for (int d = 0; d < D; ++d) {
        if (done) continue;
        if (!have[d]) { done = true; continue; }
        store(va[d], img[d]);
        have[d] = u < nsuper;
        if (have[d]) { load(va[d], img[d]); advance(); }
      }
    }
  } else if (warp < 12) {
    // ================= epilogue: this CTA's 128 accumulator lanes -> its tile of C =================
    const int q = warp & 3;
    const int r = q * 32 + lane;
    const long long row_off = (r & 31) + ((long long)(r >> 5) << (5 + TNB));
    constexpr int CH = 16;
    uint32_t cmax = 0;
    unsigned long long it = 0;
    for (unsigned long long u = pair0; u < nsuper; u += npairs, ++it) {
      const int buf = (int)(it % NBUF);
      const uint32_t tph = (uint32_t)((it / NBUF) & 1);
      unsigned split;
      unsigned long long nh, mh;
      decode(u, split, nh, mh);
      const long long c_base = deposit_warp(mh, P.c_mhi, P.n_mhi) | deposit_warp(nh, P.c_nhi, P.n_nhi);
      float2 *out = (P.split_bits ? ws + (long long)split * P.c_elems : C) + c_base + row_off;
      mbar_wait(tfull_bar(buf), tph);
      tc_fence_after();
      const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * ACC * TN);
#pragma unroll 1
      for (int c0 = 0; c0 < TN; c0 += CH) {
        uint32_t re[CH], im[CH], xr[CH], xi[CH];
        tmem_ld<CH>::run(trow + (uint32_t)c0, re);
        tmem_ld<CH>::run(trow + (uint32_t)(TN + c0), im);
        tmem_ld<CH>::run(trow + (uint32_t)(2 * TN + c0), xr);
        tmem_ld<CH>::run(trow + (uint32_t)(3 * TN + c0), xi);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < CH; ++j) {
          const float2 v = make_float2(__uint_as_float(re[j]) + __uint_as_float(xr[j]),
                                       __uint_as_float(im[j]) + __uint_as_float(xi[j]));
          amax_acc(cmax, v);
          out[(long long)(c0 + j) << 5] = v;
        }
      }
      tc_fence_before();
      if (rank == 0) mbar_arrive(tempty_bar(buf));
      else mbar_arrive_remote(tempty_bar(buf), 0);
    }
    if (P.split_bits == 0) amax_publish(amax_c, cmax);
  } else if (rank == 0) {
    // ================= MMA issuer (leader CTA): whole warp walks the loops, one elected lane issues =================
    int s = 0;
    uint32_t ph = 0;
    unsigned long long it = 0;
    const uint32_t tmem_b = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t stage0 = smem_u32(stages);
    for (unsigned long long u = pair0; u < nsuper; u += npairs, ++it) {
      const int buf = (int)(it % NBUF);
      const uint32_t tph = (uint32_t)((it / NBUF) & 1);
      mbar_wait_cluster(tempty_bar(buf), tph ^ 1u);
      tc_fence_after();
      const uint32_t cr = tmem_b + (uint32_t)(buf * ACC * TN), ci = cr + TN;
      const uint32_t xr = cr + 2 * TN, xi = cr + 3 * TN;
      for (unsigned long long kk = 0; kk < k_iters; ++kk) {
        mbar_wait(full_bar(s), ph);
        mbar_wait_cluster(pfull_bar(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t sa = stage0 + (uint32_t)s * (uint32_t)(STAGE * sizeof(float));
          const uint32_t sb = sa + 16u * PLANE_A;
#pragma unroll
          for (int ks = 0; ks < KT / 8; ++ks) {
            const uint32_t ko = (uint32_t)ks * 2u * LBO;
            const uint64_t arh = umma_desc(sa + ko, LBO, SBO);
            const uint64_t arl = umma_desc(sa + 4u * PLANE_A + ko, LBO, SBO);
            const uint64_t aih = umma_desc(sa + 8u * PLANE_A + ko, LBO, SBO);
            const uint64_t ail = umma_desc(sa + 12u * PLANE_A + ko, LBO, SBO);
            const uint64_t brh = umma_desc(sb + ko, LBO, SBO);
            const uint64_t brl = umma_desc(sb + 4u * PLANE_BH + ko, LBO, SBO);
            const uint64_t bih = umma_desc(sb + 8u * PLANE_BH + ko, LBO, SBO);
            const uint64_t bil = umma_desc(sb + 12u * PLANE_BH + ko, LBO, SBO);
            const uint32_t acc = (kk | (unsigned)ks) ? 1u : 0u;
            umma_tf32_pair(xr, arl, brh, IDESC, acc);
            umma_tf32_pair(xr, arh, brl, IDESC, 1u);
            umma_tf32_pair(xr, ail, bih, IDESC_NEG, 1u);
            umma_tf32_pair(xr, aih, bil, IDESC_NEG, 1u);
            umma_tf32_pair(cr, arh, brh, IDESC, acc);
            umma_tf32_pair(cr, aih, bih, IDESC_NEG, 1u);
            umma_tf32_pair(xi, arl, bih, IDESC, acc);
            umma_tf32_pair(xi, arh, bil, IDESC, 1u);
            umma_tf32_pair(xi, ail, brh, IDESC, 1u);
            umma_tf32_pair(xi, aih, brl, IDESC, 1u);
            umma_tf32_pair(ci, arh, bih, IDESC, acc);
            umma_tf32_pair(ci, aih, brh, IDESC, 1u);
            if (P.tc_terms >= 4) {
              umma_tf32_pair(xr, arl, brl, IDESC, 1u);
              umma_tf32_pair(xr, ail, bil, IDESC_NEG, 1u);
              umma_tf32_pair(xi, arl, bil, IDESC, 1u);
              umma_tf32_pair(xi, ail, brl, IDESC, 1u);
            }
          }
          umma_commit_pair(empty_bar(s));
          if (kk + 1 == k_iters) umma_commit_pair(tfull_bar(buf));
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  } else {
    // ================= forwarder (peer CTA): "my stage is full" -> the leader =================
    int s = 0;
    uint32_t ph = 0;
    for (unsigned long long u = pair0; u < nsuper; u += npairs) {
      for (unsigned long long kk = 0; kk < k_iters; ++kk) {
        mbar_wait(full_bar(s), ph);
        if (lane == 0) mbar_arrive_remote(pfull_bar(s), 0);
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();                              // nobody leaves while the other still signals / reads it
  if (warp == 12) {
    tc_fence_after();
    tmem_dealloc_pair<TMEM_COLS>(tmem_base);
  }
}


// --------------------------------------------------------------------------------------------
// Staggered CTA-pair variant ("tc_pair" plan value 2, the default for paired tiles; written for TN = 128
// columns below, the same with TN / 2-column halves and TN / 4-row chunks for 64 and 32).
//
// A 128 x 128 tile with split accumulators fills TMEM (4 regions x 128 columns), so the pair kernel
// above cannot double-buffer it: the tensor pipe idles while the epilogue drains 256 KB of TMEM per
// tile (ncu, round 1: tensor pipe active 61.7 %, = exactly the MMA cycles / elapsed cycles).  Here the
// tile is treated as two 64-column HALVES that share every A stage but own their accumulators:
//     TMEM columns  [h*256, h*256+128)      HH_h : hi*hi sums         (half h = n columns h*64 .. h*64+63)
//                   [h*256+128, h*256+256)  XX_h : the cross terms
// and the MMA warp walks the flat (tile, k chunk) stream with TWO cursors: half 0 runs ahead of half 1
// by up to the depth of the shared-memory ring.  When half 0 finishes a tile its accumulators drain
// while the tensor pipe still has the lagging half-1 chunks of that tile to do, and half 0 starts the
// next tile while half 1 drains -- the drain of one half hides behind the MMAs of the other.
//
// Re and Im of a product are produced by ONE MMA of 128 columns: the B operand of half h is the row
// concatenation [Br | Bi] (for Ar) or [Bi | -Br] (for Ai, issued with the negate-A bit), so the 64
// columns of a half cost the same 12 MMAs of N = 128 per 8 contracted values as before and A is not
// read from shared memory more often.  In cta_group::2 CTA r supplies rows [r*64, r*64+64) of every
// B operand; with n = h*64 + r*32 + j (j < 32) a region's 128 columns are
//     [ re(r=0) 32 | im(r=0) 32 | re(r=1) 32 | im(r=1) 32 ].
// B images (tc_prepack_b2_kernel) per (n tile, k chunk) and per rank r:
//     half h: hi sequence [Br_hi | Bi_hi | -Br_hi], lo sequence [Br_lo | Bi_lo | -Br_lo], 32 rows each,
// 384 rows x KT per rank = one cp.async.bulk per stage and CTA.
// --------------------------------------------------------------------------------------------
// Non-blocking probes (mbarrier.test_wait): the MMA warp of the staggered kernel polls several barriers
// and must not be put to sleep on one of them (try_wait suspends the thread for a system-dependent
// time when the phase is not complete) while the other half has chunks ready to issue.
// Relaxed semantics on purpose: an .acquire.cluster probe makes ptxas emit CCTL.IVALL (an L1 flush) on
// every success and a .release.cluster arrive emits MEMBAR.ALL.GPU + ERRBAR, which stalls the arriving
// thread until all its global stores are performed (ncu, round 2: 7 % of the epilogue's and 16 % of
// the forwarder's samples, and ~500 cycles per pass of the issue loop).  What the barriers order here
// is tensor-core traffic: shared-memory operands are published by fence.proxy.async + the local full
// barrier before the forwarder arrives, TMEM reads are closed by tcgen05.wait::ld +
// tcgen05.fence::before_thread_sync before the epilogue arrives, and the MMA warp runs
// tcgen05.fence::after_thread_sync after every successful probe.
__device__ __forceinline__ bool mbar_test_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {     // acquire at CTA scope
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote_relaxed(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar), "r"(cta)
      : "memory");
}

// ---- scaled 3xFP16 split (plan tc_kind = 1) ------------------------------------------------------
// x * 2^s = h1 + h2 / 2^11 (+ error <= 2^-24 |x| 2^s), h1 = fp16(x 2^s), h2 = fp16((x 2^s - h1) 2^11),
// with s = 14 - exponent(amax) so that the largest magnitude of the tensor lands in [2^14, 2^15): every
// value within 2^-28 of the largest keeps 22 significant bits (fp16 normals span 2^-14 .. 2^15, the
// residual is rescaled into the same range), smaller ones degrade gradually instead of vanishing.
// HH accumulates h1*h1, XX accumulates h1*h2 + h2*h1 (scaled by 2^11); the epilogue forms
// (HH + XX / 2^11) / 2^(sa+sb).  fp16 carries the same 11 significant bits as tf32, and kind::f16
// MMAs contract 16 values in the cycles kind::tf32 needs for 8: half the tensor-pipe work.
__device__ __forceinline__ int amax_scale_exp(const uint32_t *amax) {
  const uint32_t bits = amax ? *reinterpret_cast<const volatile uint32_t *>(amax) : 0u;
  int e = (int)((bits >> 23) & 0xffu) - 127;
  if ((bits & 0x7fffffffu) == 0u) e = 0;
  int s = 14 - e;
  return s < -100 ? -100 : (s > 100 ? 100 : s);
}
__device__ __forceinline__ float exp2_int(int s) { return __uint_as_float((uint32_t)(s + 127) << 23); }
__device__ __forceinline__ void split_f16x2(float x0, float x1, uint32_t &h1, uint32_t &h2) {
  const __half2 a = __floats2half2_rn(x0, x1);                       // x0 -> low half (lower k)
  const float2 af = __half22float2(a);
  const __half2 b = __floats2half2_rn((x0 - af.x) * 2048.f, (x1 - af.y) * 2048.f);
  h1 = *reinterpret_cast<const uint32_t *>(&a);
  h2 = *reinterpret_cast<const uint32_t *>(&b);
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// kind::f16 instruction descriptor: D fp32, A/B fp16 (format 0), both K-major, M = 256
__host__ __device__ constexpr uint32_t umma_idesc_f16_m256(int n, int neg_a) {
  return (1u << 4) | ((uint32_t)neg_a << 13) | ((uint32_t)(n >> 3) << 17) | ((256u >> 4) << 24);
}

template <int TNB, int TKB, bool F16>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreadsP2, 1)
gett_tc_pair2_kernel(const __grid_constant__ qtn_pair_plan_t P, const float2 *__restrict__ A,
                     float2 *__restrict__ C, float2 *__restrict__ ws, const int32_t *__restrict__ slice_id,
                     const uint32_t *__restrict__ amax_a, const uint32_t *__restrict__ amax_b,
                     uint32_t *__restrict__ amax_c) {
  constexpr int TN = 1 << TNB, KT = 1 << TKB;
  constexpr int Q = TN / 4;                             // B rows per chunk = columns of one [re] or [im] block
  constexpr int EB = F16 ? 2 : 4;                       // bytes per operand-plane element
  constexpr int PLANE_A = 128 * KT * EB;                // bytes
  constexpr int CHUNK = Q * KT * EB, SEQ = 3 * CHUNK, HALF = 2 * SEQ, RANK_B = 2 * HALF;
  constexpr int STAGE = 4 * PLANE_A + RANK_B;
  constexpr int NA = 128 * KT / kProducerThreads;
  constexpr int TMEM_COLS = 4 * TN;                     // two halves x (HH, XX) x TN columns
  static_assert(TNB >= 5 && TNB <= 7, "staggered pair tiles have 32, 64 or 128 columns");
  constexpr int KSTEPS = KT * EB / 32;                  // MMAs of one operand pair per stage (32 bytes of K each)
  constexpr uint32_t LBO = 128, SBO = KT * EB * 8;
  constexpr uint32_t IDESC = F16 ? umma_idesc_f16_m256(TN, 0) : umma_idesc_tf32_m256(TN, 0);
  constexpr uint32_t IDESC_NEG = F16 ? umma_idesc_f16_m256(TN, 1) : umma_idesc_tf32_m256(TN, 1);
  static_assert(!F16 || TKB == 4, "the fp16 split packs k pairs: 16 values per stage");

  // A travels global -> shared memory with cp.async (8 bytes per thread and element) into a ring of kRawDepth
  // raw chunks, and only then through registers into the operand planes: the gathers of kRawDepth chunks
  // (96 KB per SM) are in flight without holding registers, where a register ring of 3 chunks (48 KB)
  // left the producers waiting on DRAM for a third of their time (ncu, round 2: long-scoreboard stalls at
  // the first use of the loaded values; read latency behind the 16 GB of C writes is ~3 us).
  constexpr int RAW = NA * kProducerThreads * 8;        // bytes of one raw chunk (the tile's A elements, 8 B each)

  extern __shared__ __align__(1024) unsigned char smem_raw[];
  const int S = P.tc_stages;
  unsigned char *raw_ring = smem_raw + (size_t)S * STAGE;
  // (a pre-packed A needs no raw ring: the plan gives those bytes to more stages)
  uint64_t *bars = reinterpret_cast<uint64_t *>(raw_ring + (P.tc_prepack_a ? (size_t)0 : (size_t)kRawDepth * RAW));
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 3 * S + 4);
  unsigned long long *img_ring = reinterpret_cast<unsigned long long *>(bars + 3 * S + 4 + 2);   // thread 0 only
  uint64_t *raw_bars = reinterpret_cast<uint64_t *>(img_ring + kRawDepth);   // "raw chunk landed" (tc_gather)
  const int t = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0), lane = t & 31;   // provably warp-uniform
  const uint32_t rank = __shfl_sync(0xffffffffu, cluster_ctarank(), 0);
  const uint32_t bar0 = smem_u32(bars);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (S + s); };
  auto pfull_bar = [&](int s) { return bar0 + 8u * (2 * S + s); };
  auto tfull_bar = [&](int h) { return bar0 + 8u * (3 * S + h); };
  auto tempty_bar = [&](int h) { return bar0 + 8u * (3 * S + 2 + h); };

  const unsigned long long nsuper = (unsigned long long)P.grid >> 1;
  const unsigned long long pair0 = blockIdx.x >> 1, npairs = gridDim.x >> 1;
  const int k_iters_bits = P.n_khi - P.split_bits;
  const unsigned long long k_iters = 1ull << k_iters_bits;
  const long long a_slice = slice_offset(slice_id, P.n_slice_a, P.slice_bit_a, P.slice_pos_a);
  auto decode = [&](unsigned long long u, unsigned &split, unsigned long long &nh, unsigned long long &mh) {
    split = (unsigned)(u & ((1ull << P.split_bits) - 1));
    u >>= P.split_bits;
    nh = u & ((1ull << P.n_nhi) - 1);
    mh = ((u >> P.n_nhi) << 1) | rank;
  };

  if (t == 32) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), P.tc_prepack_a ? 1 : kProducerThreads + 1);
      mbar_init(empty_bar(s), 2);                    // one commit per MMA-issuing warp (half)
      mbar_init(pfull_bar(s), 1);
    }
    for (int h = 0; h < 2; ++h) { mbar_init(tfull_bar(h), 1); mbar_init(tempty_bar(h), 2 * kEpilogueThreads2); }
    for (int d = 0; d < kRawDepth; ++d) mbar_init(smem_u32(raw_bars + d), P.tc_gather == 2 ? 1 : kProducerThreads);
    fence_proxy_async_smem();
    fence_mbar_init();
  }
  if (warp == kMmaWarp2) tmem_alloc_pair<TMEM_COLS>(smem_u32(tmem_slot));
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 8 && P.tc_prepack_a) {
    // ================= producer (one thread): A and B arrive as pre-packed stage images, two bulk copies per stage ======
    if constexpr (F16) {
      if (t == 0) {
        constexpr uint32_t kRankBytes = RANK_B, kImageBytes = 2u * kRankBytes, kImageA = 4u * PLANE_A;
        const unsigned char *images = reinterpret_cast<const unsigned char *>(ws) + P.tc_prepack_off + (size_t)rank * kRankBytes;
        const unsigned char *images_a = reinterpret_cast<const unsigned char *>(ws) + P.tc_prepack_a_off;
        int s = 0;
        uint32_t ph = 0;
        for (unsigned long long u = pair0; u < nsuper; u += npairs) {
          unsigned split;
          unsigned long long nh, mh;
          decode(u, split, nh, mh);
          const unsigned long long kh0 = (unsigned long long)split << k_iters_bits;
          const unsigned long long bt = P.n_batch ? mh >> (P.n_mhi - P.n_batch) : 0ull;
          const unsigned long long img_b0 = ((((bt << P.n_nhi) | nh) << P.n_khi)) | kh0;
          const unsigned long long img_a0 = (mh << P.n_khi) | kh0;
          for (unsigned long long kk = 0; kk < k_iters; ++kk) {
            mbar_wait(empty_bar(s), ph ^ 1u);
            unsigned char *sA = smem_raw + (size_t)s * STAGE;
            mbar_arrive_expect_tx(full_bar(s), kImageA + kRankBytes);
            bulk_copy_g2s(smem_u32(sA), images_a + (img_a0 + kk) * kImageA, kImageA, full_bar(s));
            bulk_copy_g2s(smem_u32(sA + 4 * PLANE_A), images + (img_b0 + kk) * kImageBytes, kRankBytes, full_bar(s));
            if (++s == S) { s = 0; ph ^= 1u; }
          }
        }
      }
    }
  } else if (warp < 8) {
    // ================= producers: this CTA's A rows (split in registers), its B image (one bulk copy) =================
    long long a_t_g = 0;
    int a_t_s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (!((t >> j) & 1)) continue;
      if (j < P.na_lo) { a_t_g |= 1ll << P.a_lo_pos[j]; a_t_s += P.a_lo_sstr[j]; }
    }
    // tc_gather: the copies of a chunk are issued in ascending ADDRESS order of the tile's labels (bit r of the
    // element index e = i * 256 + t walks the label of rank r), element e lands at raw[e]; the thread that
    // converts an element finds it at the index built from the ranks of its own enumeration bits.
    const bool gather = P.tc_gather != 0;
    // tc_gather == 2: the tile's labels own a contiguous run of 2^tc_run_bits elements at the bottom of A's address
    // bits, so a chunk is 2^(n - run_bits) contiguous runs: lane q of warp 0 fetches run q with ONE bulk copy (the
    // TMA engine: no LSU gather at all) into the raw slot, element e of the address order again at raw[e]
    const bool bulk = P.tc_gather == 2;
    const int run_bits = P.tc_run_bits;
    const int n_runs = bulk ? 1 << (P.na_lo - run_bits) : 0;
    long long run_g = 0;                              // global offset of this lane's run inside the chunk
    if (bulk && t < n_runs) {
      for (int j = run_bits; j < P.na_lo; ++j)
        if ((t >> (j - run_bits)) & 1) run_g |= 1ll << P.tc_a_gbit[j];
    }
    long long g_t_g = 0;
    int r_t = 0;
    if (gather) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (!((t >> j) & 1)) continue;
        g_t_g |= 1ll << P.tc_a_gbit[j];
        r_t += 1 << P.tc_a_rank[j];
      }
    }
    const float sgn_a = P.conj_a ? -1.f : 1.f;
    const float sc_re = F16 ? exp2_int(amax_scale_exp(amax_a)) : 1.f, sc_im = sc_re * sgn_a;
    unsigned long long u = pair0, kk = 0;
    long long a_k = 0;
    constexpr uint32_t kRankBytes = RANK_B, kImageBytes = 2u * kRankBytes;
    const unsigned char *images = reinterpret_cast<const unsigned char *>(ws) + P.tc_prepack_off + (size_t)rank * kRankBytes;
    unsigned long long image0 = 0;
    auto enter_tile = [&]() {
      unsigned split;
      unsigned long long nh, mh;
      decode(u, split, nh, mh);
      const unsigned long long kh0 = (unsigned long long)split << k_iters_bits;
      a_k = a_slice | deposit_warp(mh, P.a_mhi, P.n_mhi) | deposit_warp(kh0, P.a_khi, P.n_khi);
      // batch labels are the top n_batch bits of the m-tile index: one set of B images per batch value
      const unsigned long long bt = P.n_batch ? mh >> (P.n_mhi - P.n_batch) : 0ull;
      image0 = ((((bt << P.n_nhi) | nh) << P.n_khi)) | kh0;
      kk = 0;
    };
    auto advance = [&]() {
      if (kk + 1 == k_iters) { u += npairs; if (u < nsuper) enter_tile(); return; }
      const unsigned long long flip = kk ^ (kk + 1);
      const int nflip = 64 - __clzll(flip);
      for (int i = 0; i < nflip && i < k_iters_bits; ++i) a_k ^= 1ll << P.a_khi[i];
      ++kk;
    };
    const uint32_t raw_t = smem_u32(raw_ring) + (uint32_t)t * 8u;
    auto load = [&](int slot) {                       // this thread's NA elements of the current chunk -> raw slot
      const uint32_t dst = raw_t + (uint32_t)slot * (uint32_t)RAW;
      if (bulk) {
        if (warp == 0) {
          const uint32_t bar = smem_u32(raw_bars + slot);
          fence_proxy_async_smem();                   // the slot's earlier (generic-proxy) reads precede the TMA writes
          if (t == 0) mbar_arrive_expect_tx(bar, (uint32_t)RAW);
          __syncwarp();
          if (t < n_runs)
            bulk_copy_g2s(smem_u32(raw_ring) + (uint32_t)slot * (uint32_t)RAW + ((uint32_t)t << (run_bits + 3)),
                          A + (a_k | run_g), 8u << run_bits, bar);
        }
      } else if (gather) {
#pragma unroll
        for (int i = 0; i < NA; ++i)
          cp_async8(dst + (uint32_t)i * (kProducerThreads * 8u), A + (a_k | g_t_g | P.tc_a_it_gg[i]));
      } else {
#pragma unroll
        for (int i = 0; i < NA; ++i)
          cp_async8(dst + (uint32_t)i * (kProducerThreads * 8u), A + (a_k | a_t_g | P.tc_a_it_g[i]));
      }
      if (t == 0) img_ring[slot] = image0 + kk;
    };
    int s = 0;
    uint32_t ph = 0;
    auto store = [&](int slot) {
      float2 va[NA];
      const unsigned long long img = t == 0 ? img_ring[slot] : 0ull;   // before a refill of the slot replaces it
      if (gather) {
        const float2 *src = reinterpret_cast<const float2 *>(raw_ring + (size_t)slot * RAW) + r_t;
#pragma unroll
        for (int i = 0; i < NA; ++i) va[i] = src[P.tc_a_it_raw[i]];
        // The loads must have RETURNED before this thread reports the slot free: a shared-memory load that is merely
        // issued can still be overtaken by another thread's copy into the slot (seen on the full C4 contraction with
        // bulk copies: relative error 2.7e-4).  A value that depends on every loaded element, pinned before the
        // barrier, makes the barrier wait for the data.
        {
          float chk = 0.f;
#pragma unroll
          for (int i = 0; i < NA; ++i) chk += va[i].x;
          asm volatile("" ::"f"(chk) : "memory");
        }
        // every producer has taken its elements out of the slot: it may be refilled (by other threads' copies)
        asm volatile("bar.sync 1, %0;" ::"n"(kProducerThreads) : "memory");
        if (u < nsuper) { load(slot); advance(); }
        if (!bulk) cp_async_mbar_arrive(smem_u32(raw_bars + slot));
      } else {
        const float2 *src = reinterpret_cast<const float2 *>(raw_ring + (size_t)slot * RAW) + t;
#pragma unroll
        for (int i = 0; i < NA; ++i) va[i] = src[i * kProducerThreads];
      }
      mbar_wait(empty_bar(s), ph ^ 1u);
      unsigned char *sA = smem_raw + (size_t)s * STAGE;
      if (t == 0) {
        mbar_arrive_expect_tx(full_bar(s), kRankBytes);
        bulk_copy_g2s(smem_u32(sA + 4 * PLANE_A), images + img * kImageBytes, kRankBytes, full_bar(s));
      }
      if constexpr (F16) {
        // elements 2p and 2p+1 of a thread differ in the lowest k bit: one packed 32-bit store per plane
        uint32_t *w = reinterpret_cast<uint32_t *>(sA);
        constexpr int PW = PLANE_A / 4;
#pragma unroll
        for (int p = 0; p < NA / 2; ++p) {
          const int o = (a_t_s + P.tc_a_it_s[2 * p]) >> 1;
          uint32_t h1, h2;
          split_f16x2(va[2 * p].x * sc_re, va[2 * p + 1].x * sc_re, h1, h2);
          w[o] = h1; w[PW + o] = h2;
          split_f16x2(va[2 * p].y * sc_im, va[2 * p + 1].y * sc_im, h1, h2);
          w[2 * PW + o] = h1; w[3 * PW + o] = h2;
        }
      } else {
        float *f = reinterpret_cast<float *>(sA);
        constexpr int PF = PLANE_A / 4;
#pragma unroll
        for (int i = 0; i < NA; ++i) {
          const int o = a_t_s + P.tc_a_it_s[i];
          float h, l;
          split_tf32(va[i].x, h, l);
          f[o] = h; f[PF + o] = l;
          split_tf32(va[i].y * sgn_a, h, l);
          f[2 * PF + o] = h; f[3 * PF + o] = l;
        }
      }
      fence_proxy_async_smem();
      mbar_arrive(full_bar(s));
      if (++s == S) { s = 0; ph ^= 1u; }
    };
    // every ring slot owns one cp.async group per lap (an empty one once the stream has ended), so that
    // "all but the kRawDepth - 1 youngest groups are complete" always means "the oldest slot has landed"
    const unsigned long long my_tiles = pair0 < nsuper ? (nsuper - pair0 + npairs - 1) / npairs : 0;
    unsigned long long remaining = my_tiles * k_iters;   // chunks still to be stored
    if (u < nsuper) enter_tile();
    if (gather) {
      // every thread arrives on a slot's barrier once per lap, when its own copies into the slot (possibly none)
      // have landed: 256 arrivals = the chunk is complete, whoever copied which element
#pragma unroll 1
      for (int d = 0; d < kRawDepth; ++d) {
        if (u < nsuper) { load(d); advance(); }
        if (!bulk) cp_async_mbar_arrive(smem_u32(raw_bars + d));
      }
      int slot = 0;
      uint32_t rph = 0;
#pragma unroll 1
      while (remaining) {
        mbar_wait(smem_u32(raw_bars + slot), rph);
        store(slot);                                     // refills the slot itself, behind the named barrier
        if (++slot == kRawDepth) { slot = 0; rph ^= 1u; }
        --remaining;
      }
    } else {
#pragma unroll 1
      for (int d = 0; d < kRawDepth; ++d) {
        if (u < nsuper) { load(d); advance(); }
        cp_async_commit();
      }
      int slot = 0;
#pragma unroll 1
      while (remaining) {
        cp_async_wait<kRawDepth - 1>();
        store(slot);
        if (u < nsuper) { load(slot); advance(); }
        cp_async_commit();
        if (++slot == kRawDepth) slot = 0;
        --remaining;
      }
    }
  } else if (warp < kMmaWarp2) {
    // ================= epilogue: half 0 then half 1 of every tile, each as soon as its MMAs are complete =================
    // EIGHT warps: a warp reads the TMEM lanes of its quarter (warp & 3) only, so two warps share a quarter and take
    // one rank block of columns each.  With four warps the drain of a half tile (TMEM load latency + 64 stores per
    // thread) took longer than the other half's MMAs: ncu with both operands pre-packed -- no producer work left --
    // still showed the tensor pipe at 52 % (profiles/r02_ncu_pair2_prepack_a_m21n10k7.txt).
    const int q = warp & 3;
    const int eh = (warp - 8) >> 2;                     // which rank block of every half this warp drains
    const int r = q * 32 + lane;
    const long long row_off = (r & 31) + ((long long)(r >> 5) << (5 + P.n_real));   // n_real < TNB: phantom column bit
    const bool store_h1 = P.n_real == TNB;              // the phantom half is multiplied (zeros) but never stored
    constexpr int CH = Q < 16 ? Q : 16, NCH = Q / CH;  // column chunks of one [re] block
    // fp16 split: undo the operand scales (two factors: their product may leave the fp32 range)
    const float inv_a = F16 ? exp2_int(-amax_scale_exp(amax_a)) : 1.f;
    const float inv_b = F16 ? exp2_int(-amax_scale_exp(amax_b)) : 1.f;
    const float xs = F16 ? (1.f / 2048.f) : 1.f;
    uint32_t cmax = 0;
    unsigned long long it = 0;
    for (unsigned long long u = pair0; u < nsuper; u += npairs, ++it) {
      const uint32_t tph = (uint32_t)(it & 1);
      unsigned split;
      unsigned long long nh, mh;
      decode(u, split, nh, mh);
      const long long c_base = deposit_warp(mh, P.c_mhi, P.n_mhi) | deposit_warp(nh, P.c_nhi, P.n_nhi);
      float2 *out = (P.split_bits ? ws + (long long)split * P.c_elems : C) + c_base + row_off;
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        mbar_wait(tfull_bar(h), tph);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * 2 * TN);
#pragma unroll 2
        for (int rc = eh * NCH; rc < (eh + 1) * NCH && (h == 0 || store_h1); ++rc) {
          const int rr = rc / NCH, cc = (rc % NCH) * CH;    // rank block, first column inside it
          const uint32_t col = (uint32_t)(rr * 2 * Q + cc);
          uint32_t re[CH], im[CH], xr[CH], xi[CH];
          tmem_ld<CH>::run(trow + col, re);
          tmem_ld<CH>::run(trow + col + (uint32_t)Q, im);
          tmem_ld<CH>::run(trow + (uint32_t)TN + col, xr);
          tmem_ld<CH>::run(trow + (uint32_t)(TN + Q) + col, xi);
          tmem_ld_wait();
          float2 *o = out + ((long long)(h * (TN / 2) + rr * Q + cc) << 5);
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            float2 v;
            if constexpr (F16) {
              v.x = fmaf(__uint_as_float(xr[j]), xs, __uint_as_float(re[j])) * inv_a * inv_b;
              v.y = fmaf(__uint_as_float(xi[j]), xs, __uint_as_float(im[j])) * inv_a * inv_b;
            } else {
              v.x = __uint_as_float(re[j]) + __uint_as_float(xr[j]);
              v.y = __uint_as_float(im[j]) + __uint_as_float(xi[j]);
            }
            amax_acc(cmax, v);
            o[(long long)j << 5] = v;
          }
        }
        tc_fence_before();
        if (rank == 0) mbar_arrive_relaxed(tempty_bar(h));
        else mbar_arrive_remote_relaxed(tempty_bar(h), 0);
      }
    }
    if (P.split_bits == 0) amax_publish(amax_c, cmax);
  } else if (rank == 0) {
    // ================= MMA issuers (leader CTA): one warp per half =================
    // Warp 16 issues the MMAs of half 0, warp 17 those of half 1, each in a plain blocking loop over the flat
    // (tile, k chunk) stream: wait for the half's accumulators at a tile start, wait for the stage, six MMAs, commit.
    // A stage is released when BOTH warps have committed on its empty barrier (count 2).  The halves stagger by
    // themselves: the epilogue drains half 0 first, so half 0 resumes first and half 1 works on the remaining
    // chunks of the tile while half 0 waits for its next drain.  The first version walked both cursors in ONE warp
    // that polled up to four barriers per pass and rebuilt eight descriptors per chunk: ~150 dependent
    // instructions, ~500 cycles per pass for 1.26 chunk-halves on average -- the issue loop, not the producers or
    // the epilogue, kept the tensor pipe at 48-52 % (ncu source view; pre-packing A and doubling the epilogue
    // warps changed nothing, a one-thread loop and a fixed-lag single-warp loop were slower).
    const int h = warp - kMmaWarp2;                     // the half this warp issues
    const uint32_t tmem_b = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t stage0 = smem_u32(smem_raw);
    constexpr uint64_t D_HI = ((uint64_t)(SBO >> 4) << 32) | (1ull << 46) | ((uint64_t)(LBO >> 4) << 16);
    constexpr uint32_t STEP = (uint32_t)STAGE >> 4;      // stage s: address field + s * STEP
    constexpr uint32_t KO = (2u * LBO) >> 4;             // k step inside a stage
    auto field = [&](uint32_t addr) { return (addr & 0x3FFFFu) >> 4; };
    const uint32_t a_rh = field(stage0), a_rl = field(stage0 + PLANE_A), a_ih = field(stage0 + 2u * PLANE_A),
                   a_il = field(stage0 + 3u * PLANE_A);
    const uint32_t bh = field(stage0 + 4u * PLANE_A + (uint32_t)h * (uint32_t)HALF);
    constexpr uint32_t F_CHUNK = (uint32_t)CHUNK >> 4, F_SEQ = (uint32_t)SEQ >> 4;
    const uint32_t hh = tmem_b + (uint32_t)(h * 2 * TN), xx = hh + (uint32_t)TN;
    auto all_lanes = [](bool v) { return __all_sync(0xffffffffu, v) != 0; };
    auto wait_relaxed = [&](uint32_t bar, uint32_t parity) {      // barriers with arrivals from the peer CTA
      unsigned spins = 0;
      while (!all_lanes(mbar_test_relaxed(bar, parity)))
        if (++spins > (1u << 26)) __trap();
    };
    int s = 0;
    uint32_t ph = 0;
    unsigned long long it = 0;
    for (unsigned long long u = pair0; u < nsuper; u += npairs, ++it) {
      wait_relaxed(tempty_bar(h), (uint32_t)((it & 1) ^ 1));       // this half's accumulators are drained
      for (unsigned long long kk = 0; kk < k_iters; ++kk) {
        mbar_wait(full_bar(s), ph);
        wait_relaxed(pfull_bar(s), ph);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t so = (uint32_t)s * STEP;
#pragma unroll
          for (int ks = 0; ks < KSTEPS; ++ks) {
            const uint32_t o = so + (uint32_t)ks * KO;
            const uint64_t arh = D_HI | (a_rh + o), arl = D_HI | (a_rl + o);
            const uint64_t aih = D_HI | (a_ih + o), ail = D_HI | (a_il + o);
            const uint64_t bh0 = D_HI | (bh + o), bh1 = D_HI | (bh + F_CHUNK + o);      // [Br_hi | Bi_hi], [Bi_hi | -Br_hi]
            const uint64_t bl0 = D_HI | (bh + F_SEQ + o), bl1 = D_HI | (bh + F_SEQ + F_CHUNK + o);
            const uint32_t acc = (kk | (unsigned)ks) ? 1u : 0u;
            if constexpr (F16) {
              umma_f16_pair(xx, arl, bh0, IDESC, acc);
              umma_f16_pair(xx, ail, bh1, IDESC_NEG, 1u);
              umma_f16_pair(xx, arh, bl0, IDESC, 1u);
              umma_f16_pair(xx, aih, bl1, IDESC_NEG, 1u);
              umma_f16_pair(hh, arh, bh0, IDESC, acc);
              umma_f16_pair(hh, aih, bh1, IDESC_NEG, 1u);
            } else {
              umma_tf32_pair(xx, arl, bh0, IDESC, acc);
              umma_tf32_pair(xx, ail, bh1, IDESC_NEG, 1u);
              umma_tf32_pair(xx, arh, bl0, IDESC, 1u);
              umma_tf32_pair(xx, aih, bl1, IDESC_NEG, 1u);
              umma_tf32_pair(hh, arh, bh0, IDESC, acc);
              umma_tf32_pair(hh, aih, bh1, IDESC_NEG, 1u);
            }
          }
          umma_commit_pair(empty_bar(s));               // this half has read the stage (the barrier counts both)
          if (kk + 1 == k_iters) umma_commit_pair(tfull_bar(h));
        }
        __syncwarp();
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  } else if (warp == kMmaWarp2) {
    // ================= forwarder (peer CTA): "my stage is full" -> the leader =================
    int s = 0;
    uint32_t ph = 0;
    for (unsigned long long u = pair0; u < nsuper; u += npairs) {
      for (unsigned long long kk = 0; kk < k_iters; ++kk) {
        mbar_wait(full_bar(s), ph);
        if (lane == 0) mbar_arrive_remote_relaxed(pfull_bar(s), 0);
        if (++s == S) { s = 0; ph ^= 1u; }
      }
    }
  }

  __syncwarp();
  tc_fence_before();
  cluster_sync_all();
  if (warp == kMmaWarp2) {
    tc_fence_after();
    tmem_dealloc_pair<TMEM_COLS>(tmem_base);
  }
}

// B images of the staggered kernel: see the layout above.  One CTA per (n tile, k chunk); the tile
// enumeration (global address, logical position) is the planner's, the logical (n, k) of an element
// is decoded from its offset in the canonical plane layout (tf32: 8 rows x 4 values per core matrix,
// fp16: 8 rows x 8 values).
template <int TNB, int TKB, bool F16>
__global__ void __launch_bounds__(kProducerThreads)
tc_prepack_b2_kernel(const __grid_constant__ qtn_pair_plan_t P, const float2 *__restrict__ B,
                     unsigned char *__restrict__ images, const int32_t *__restrict__ slice_id,
                     const uint32_t *__restrict__ amax_b) {
  constexpr int TN = 1 << TNB, KT = 1 << TKB, Q = TN / 4;
  constexpr int RG = KT * 8;                                      // elements per 8-row group of a plane
  constexpr int CHUNK = Q * KT, SEQ = 3 * CHUNK, HALF = 2 * SEQ, RANK_B = 2 * HALF;   // elements
  constexpr int NB = (TN * KT + kProducerThreads - 1) / kProducerThreads;
  const int t = threadIdx.x;
  const unsigned long long nb = (unsigned long long)blockIdx.x >> P.n_khi;       // (batch, n tile)
  const unsigned long long nh = nb & ((1ull << P.n_nhi) - 1), bt = nb >> P.n_nhi;
  const unsigned long long kh = (unsigned long long)blockIdx.x & ((1ull << P.n_khi) - 1);
  const long long b_k = slice_offset(slice_id, P.n_slice_b, P.slice_bit_b, P.slice_pos_b) |
                        deposit(nh, P.b_nhi, P.n_nhi) | deposit(kh, P.b_khi, P.n_khi) |
                        deposit(bt, P.b_bat, P.n_batch);
  long long b_t_g = 0;
  int b_t_s = 0;
#pragma unroll
  bool b_phantom = false;
  for (int j = 0; j < 8; ++j)
    if (((t >> j) & 1) && j < P.nb_lo) {
      if (P.b_lo_pos[j] == 255) b_phantom = true; else b_t_g |= 1ll << P.b_lo_pos[j];
      b_t_s += P.b_lo_sstr[j];
    }
  const float sgn_b = P.conj_b ? -1.f : 1.f;
  const float sc = F16 ? exp2_int(amax_scale_exp(amax_b)) : 1.f;
  if (!(P.nb_lo >= 8 || t < (1 << P.nb_lo))) return;
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    // a phantom column bit (position 255: 16-column batch nodes run as 32-column tiles) selects rows of zeros
    const bool phantom = b_phantom || P.tc_b_it_g[i] < 0;
    const float2 v = phantom ? make_float2(0.f, 0.f) : __ldg(B + (b_k | b_t_g | P.tc_b_it_g[i]));
    const int o = b_t_s + P.tc_b_it_s[i];                       // offset in the canonical 128 x KT plane
    int k, n, inner;
    if constexpr (F16) {
      k = (o & 7) | (((o % RG) >> 6) << 3);
      n = ((o >> 3) & 7) | ((o / RG) << 3);
      inner = ((n % Q) >> 3) * RG + (k >> 3) * 64 + (n & 7) * 8 + (k & 7);
    } else {
      k = (o & 3) | (((o % RG) >> 5) << 2);
      n = ((o >> 2) & 7) | ((o / RG) << 3);
      inner = ((n % Q) >> 3) * RG + (k >> 2) * 32 + (n & 7) * 4 + (k & 3);
    }
    const size_t at = (size_t)blockIdx.x * 2 * RANK_B + ((n / Q) & 1) * RANK_B + (n / (2 * Q)) * HALF + inner;
    if constexpr (F16) {
      __half *dst = reinterpret_cast<__half *>(images) + at;
      float x = v.x * sc;
      __half h1 = __float2half_rn(x), h2 = __float2half_rn((x - __half2float(h1)) * 2048.f);
      dst[0] = h1; dst[2 * CHUNK] = __hneg(h1);
      dst[SEQ] = h2; dst[SEQ + 2 * CHUNK] = __hneg(h2);
      x = v.y * sgn_b * sc;
      h1 = __float2half_rn(x); h2 = __float2half_rn((x - __half2float(h1)) * 2048.f);
      dst[CHUNK] = h1; dst[SEQ + CHUNK] = h2;
    } else {
      float *dst = reinterpret_cast<float *>(images) + at;
      float h, l;
      split_tf32(v.x, h, l);
      dst[0] = h; dst[2 * CHUNK] = -h;
      dst[SEQ] = l; dst[SEQ + 2 * CHUNK] = -l;
      split_tf32(v.y * sgn_b, h, l);
      dst[CHUNK] = h; dst[SEQ + CHUNK] = l;
    }
  }
}

// A images of the staggered fp16 kernel ("tc_prepack_a"): one CTA per (m tile, k chunk) gathers the 128 x 16 values of
// that stage with the producers' own enumeration (a_lo_pos / tc_a_it_g: the same addresses, the same plane offsets),
// scales, splits and writes the four planes [re h1 | re h2 | im h1 | im h2] as one 16 KB image.
template <int TKB>
__global__ void __launch_bounds__(kProducerThreads)
tc_prepack_a2_kernel(const __grid_constant__ qtn_pair_plan_t P, const float2 *__restrict__ A,
                     unsigned char *__restrict__ images, const int32_t *__restrict__ slice_id,
                     const uint32_t *__restrict__ amax_a) {
  constexpr int KT = 1 << TKB;
  constexpr int PLANE_A = 128 * KT * 2;                 // bytes
  constexpr int NA = 128 * KT / kProducerThreads;
  const int t = threadIdx.x;
  const unsigned long long mh = (unsigned long long)blockIdx.x >> P.n_khi;
  const unsigned long long kh = (unsigned long long)blockIdx.x & ((1ull << P.n_khi) - 1);
  const long long a_k = slice_offset(slice_id, P.n_slice_a, P.slice_bit_a, P.slice_pos_a) |
                        deposit(mh, P.a_mhi, P.n_mhi) | deposit(kh, P.a_khi, P.n_khi);
  long long a_t_g = 0;
  int a_t_s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (((t >> j) & 1) && j < P.na_lo) { a_t_g |= 1ll << P.a_lo_pos[j]; a_t_s += P.a_lo_sstr[j]; }
  const float sc_re = exp2_int(amax_scale_exp(amax_a)), sc_im = P.conj_a ? -sc_re : sc_re;
  float2 va[NA];
#pragma unroll
  for (int i = 0; i < NA; ++i) va[i] = __ldg(A + (a_k | a_t_g | P.tc_a_it_g[i]));
  uint32_t *w = reinterpret_cast<uint32_t *>(images + (size_t)blockIdx.x * 4 * PLANE_A);
  constexpr int PW = PLANE_A / 4;
#pragma unroll
  for (int p = 0; p < NA / 2; ++p) {                    // elements 2p, 2p+1 differ in the lowest k bit: one packed word
    const int o = (a_t_s + P.tc_a_it_s[2 * p]) >> 1;
    uint32_t h1, h2;
    split_f16x2(va[2 * p].x * sc_re, va[2 * p + 1].x * sc_re, h1, h2);
    w[o] = h1; w[PW + o] = h2;
    split_f16x2(va[2 * p].y * sc_im, va[2 * p + 1].y * sc_im, h1, h2);
    w[2 * PW + o] = h1; w[3 * PW + o] = h2;
  }
}

// B pre-pack: one CTA per (n tile, k chunk) writes that B stage -- split into hi/lo re/im planes
// and laid out exactly as the main kernel's shared-memory stage -- to the workspace.
template <int TNB, int TKB>
__global__ void __launch_bounds__(kProducerThreads)
tc_prepack_b_kernel(const __grid_constant__ qtn_pair_plan_t P, const float2 *__restrict__ B,
                    float *__restrict__ images, const int32_t *__restrict__ slice_id) {
  constexpr int TN = 1 << TNB, KT = 1 << TKB;
  constexpr int PLANE_B = TN * KT;
  constexpr int NB = (PLANE_B + kProducerThreads - 1) / kProducerThreads;
  const int t = threadIdx.x;
  const unsigned long long nh = (unsigned long long)blockIdx.x >> P.n_khi;
  const unsigned long long kh = (unsigned long long)blockIdx.x & ((1ull << P.n_khi) - 1);
  const long long b_k = slice_offset(slice_id, P.n_slice_b, P.slice_bit_b, P.slice_pos_b) |
                        deposit(nh, P.b_nhi, P.n_nhi) | deposit(kh, P.b_khi, P.n_khi);
  long long b_t_g = 0;
  int b_t_s = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j)
    if (((t >> j) & 1) && j < P.nb_lo) { b_t_g |= 1ll << P.b_lo_pos[j]; b_t_s += P.b_lo_sstr[j]; }
  if (!(P.nb_lo >= 8 || t < (1 << P.nb_lo))) return;
  const float sgn_b = P.conj_b ? -1.f : 1.f;
  float *dst = images + (size_t)blockIdx.x * 4 * PLANE_B;
#pragma unroll
  for (int i = 0; i < NB; ++i) {
    const float2 v = __ldg(B + (b_k | b_t_g | P.tc_b_it_g[i]));
    const int o = b_t_s + P.tc_b_it_s[i];
    float h, l;
    split_tf32(v.x, h, l);
    dst[o] = h; dst[PLANE_B + o] = l;
    split_tf32(v.y * sgn_b, h, l);
    dst[2 * PLANE_B + o] = h; dst[3 * PLANE_B + o] = l;
  }
}

template <int TNB, int TKB>
int launch_tc_inst(const qtn_pair_plan_t &P, const void *a, const void *b, void *c, void *ws,
                   const int32_t *slice_id, const uint32_t *amax_a, const uint32_t *amax_b, uint32_t *amax_c,
                   bool &amax_done, cudaStream_t st) {
  amax_done = P.split_bits == 0;                // the persistent kernels publish max |C| from their epilogue
  if (P.tc_reserved > 0) {                      // persistent variant: tc_reserved = CTAs to launch
    const long long ctas = P.grid < P.tc_reserved ? P.grid : P.tc_reserved;
    if (P.tc_pair == 2) {                         // staggered CTA pairs: tiles of 32, 64 or 128 columns
      if constexpr (TNB >= 5 && TNB <= 7) {
        const long long images = 1ll << (P.n_nhi + P.n_khi + P.n_batch);
        if (!P.tc_pad || !P.tc_prepack || (P.grid & 1) || (ctas & 1) || images > 0x7fffffffLL)
          return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: malformed staggered CTA-pair plan");
        auto go2 = [&](auto prepack, auto kernel, int &smem_set) -> int {
          prepack<<<(unsigned)images, kProducerThreads, 0, st>>>(P, (const float2 *)b,
                                                                 (unsigned char *)ws + P.tc_prepack_off, slice_id, amax_b);
          QTN_LAUNCH_CHECK();
          if (P.tc_prepack_a) {
            if constexpr (TKB == 4) {
              const long long a_images = 1ll << (P.n_mhi + P.n_khi);
              if (P.tc_kind != 1 || a_images > 0x7fffffffLL)
                return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: malformed A pre-pack plan");
              tc_prepack_a2_kernel<TKB><<<(unsigned)a_images, kProducerThreads, 0, st>>>(
                  P, (const float2 *)a, (unsigned char *)ws + P.tc_prepack_a_off, slice_id, amax_a);
              QTN_LAUNCH_CHECK();
            } else {
              return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: the A pre-pack needs 16 contracted values per stage");
            }
          }
          if (smem_set < P.smem_bytes) {
            QTN_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes));
            smem_set = P.smem_bytes;
          }
          kernel<<<(unsigned)ctas, kThreadsP2, P.smem_bytes, st>>>(P, (const float2 *)a, (float2 *)c, (float2 *)ws, slice_id,
                                                                  amax_a, amax_b, amax_c);
          QTN_LAUNCH_CHECK();
          return QTN_OK;
        };
        static int set_t = 0, set_h = 0;
        if (P.tc_kind == 1) {
          if constexpr (TKB == 4) {
            if (!amax_a || !amax_b)
              return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: an fp16-split plan needs the operands' magnitude words");
            return go2(tc_prepack_b2_kernel<TNB, TKB, true>, gett_tc_pair2_kernel<TNB, TKB, true>, set_h);
          } else {
            return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: the fp16 split needs 16 contracted values per stage");
          }
        }
        return go2(tc_prepack_b2_kernel<TNB, TKB, false>, gett_tc_pair2_kernel<TNB, TKB, false>, set_t);
      } else {
        return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: staggered CTA pairs need a tile of 32 to 128 columns");
      }
    }
    if (P.n_batch) return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: batch labels need the staggered CTA-pair kernel");
    if (P.tc_prepack) {
      const long long images = 1ll << (P.n_nhi + P.n_khi);
      if (images > 0x7fffffffLL) return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_pair_run: too many B images");
      tc_prepack_b_kernel<TNB, TKB><<<(unsigned)images, kProducerThreads, 0, st>>>(
          P, (const float2 *)b, (float *)((unsigned char *)ws + P.tc_prepack_off), slice_id);
      QTN_LAUNCH_CHECK();
    }
    auto go = [&](auto kernel, int &smem_set) -> int {
      if (smem_set < P.smem_bytes) {
        QTN_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes));
        smem_set = P.smem_bytes;
      }
      kernel<<<(unsigned)ctas, kThreadsP, P.smem_bytes, st>>>(P, (const float2 *)a, (const float2 *)b, (float2 *)c,
                                                             (float2 *)ws, slice_id, amax_c);
      QTN_LAUNCH_CHECK();
      return QTN_OK;
    };
    static int set_sp = 0, set_s = 0, set_p = 0, set_0 = 0;
    if (P.tc_pair) {                              // CTA pairs (cta_group::2): tc_reserved is even
      if constexpr (TNB >= 5 && TNB <= 7) {
        if (!P.tc_pad || !P.tc_prepack || (P.grid & 1) || (ctas & 1))
          return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: malformed CTA-pair plan");
        static int set_pair = 0;
        if (set_pair < P.smem_bytes) {
          QTN_CUDA_CHECK(cudaFuncSetAttribute(gett_tc_pair_kernel<TNB, TKB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              P.smem_bytes));
          set_pair = P.smem_bytes;
        }
        gett_tc_pair_kernel<TNB, TKB><<<(unsigned)ctas, kThreadsP, P.smem_bytes, st>>>(P, (const float2 *)a, (float2 *)c,
                                                                                    (float2 *)ws, slice_id, amax_c);
        QTN_LAUNCH_CHECK();
        return QTN_OK;
      } else {
        return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: CTA pairs need a tile of 32 to 128 columns");
      }
    }
    if (P.tc_pad) {                               // separate cross-term accumulators (TN <= 128)
      if constexpr (TNB <= 7) {
        return P.tc_prepack ? go(gett_tc_persistent_kernel<TNB, TKB, true, true>, set_sp)
                            : go(gett_tc_persistent_kernel<TNB, TKB, true, false>, set_s);
      } else {
        return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: split accumulators need a tile of at most 128 columns");
      }
    }
    return P.tc_prepack ? go(gett_tc_persistent_kernel<TNB, TKB, false, true>, set_p)
                        : go(gett_tc_persistent_kernel<TNB, TKB, false, false>, set_0);
  }
  if (P.n_batch) return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: batch labels need the staggered CTA-pair kernel");
  amax_done = false;                            // one CTA per tile: no magnitude tracking in the kernel
  static int smem_set = 0;
  if (smem_set < P.smem_bytes) {
    QTN_CUDA_CHECK(cudaFuncSetAttribute(gett_tc_kernel<TNB, TKB>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, P.smem_bytes));
    smem_set = P.smem_bytes;
  }
  gett_tc_kernel<TNB, TKB><<<(unsigned)P.grid, kThreads, P.smem_bytes, st>>>(
      P, (const float2 *)a, (const float2 *)b, (float2 *)c, (float2 *)ws, slice_id);
  QTN_LAUNCH_CHECK();
  return QTN_OK;
}

}  // namespace

int qtn_launch_tc(const qtn_pair_plan_t &P, const void *a, const void *b, void *c, void *ws,
                  const int32_t *slice_id, const uint32_t *amax_a, const uint32_t *amax_b, uint32_t *amax_c,
                  cudaStream_t st) {
  if (P.dtype != QTN_C64) return qtn_fail(QTN_ERR_UNSUPPORTED, "tensor-core kernel is complex64 only");
  if (P.grid <= 0 || P.grid > 0x7fffffffLL) return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_pair_run: grid too large");
  if (P.tmb != 7 || P.tc_stages < 1 || P.tc_stages > 8)
    return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: malformed tensor-core plan");
  int rc;
  bool amax_done = false;
  switch (P.tnb * 8 + P.tkb) {
#define QTN_TC_CASE(NB, KB) \
  case NB * 8 + KB: rc = launch_tc_inst<NB, KB>(P, a, b, c, ws, slice_id, amax_a, amax_b, amax_c, amax_done, st); break;
    QTN_TC_CASE(4, 3) QTN_TC_CASE(4, 4) QTN_TC_CASE(5, 3) QTN_TC_CASE(5, 4) QTN_TC_CASE(6, 3)
    QTN_TC_CASE(6, 4) QTN_TC_CASE(7, 3) QTN_TC_CASE(7, 4) QTN_TC_CASE(8, 3) QTN_TC_CASE(8, 4)
#undef QTN_TC_CASE
    default: return qtn_fail(QTN_ERR_ARG, "qtn_pair_run: unsupported tensor-core tile");
  }
  if (rc != QTN_OK) return rc;
  if (P.split_bits) {
    rc = qtn_launch_splitk_sum(QTN_C64, c, ws, P.c_elems, 1 << P.split_bits, P.accumulate, st);
    if (rc != QTN_OK) return rc;
  }
  if (amax_c && !amax_done) return qtn_launch_absmax(c, P.c_elems, amax_c, st);
  return QTN_OK;
}
