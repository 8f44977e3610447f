/*
 * qibotn_b200 -- C ABI of the B200 (sm_100a) tensor-network engine.
 *
 * The reference (qiboteam/qibotn v0.0.7) has NO C ABI or FFI of its own: its
 * arithmetic happens inside cuQuantum, reached through six Python library
 * calls.  Each entry point below names the reference call
 * site(s) it takes over; paths are relative to /root/reference/src/qibotn/.
 *
 * Conventions
 *  - plain C types only; device pointers are raw `void*`; `stream` is a
 *    `cudaStream_t` passed as `void*` (NULL = legacy default stream);
 *  - every function returns 0 on success, a negative QTN_ERR_* otherwise;
 *    `qtn_last_error()` gives a thread-local message;
 *  - the caller owns all buffers (inputs, outputs, workspace); nothing is
 *    allocated or freed behind its back except in the *_host convenience
 *    entry points, which say so;
 *  - complex numbers are interleaved (re, im): QTN_C64 = 2 x float,
 *    QTN_C128 = 2 x double;
 *  - a "bit tensor" is a tensor whose every extent is 2 (all circuit tensors
 *    are).  Its layout is given per label as an *address bit position*: the
 *    element with label values v_l sits at offset sum_l v_l << pos_l (in
 *    elements).  Dense row-major with the first label slowest is
 *    pos = rank-1, rank-2, ..., 0.  Positions may have gaps (a sliced label
 *    simply leaves its bit out and is added to the base offset).
 */
#ifndef QIBOTN_B200_H
#define QIBOTN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QTN_C64 0
#define QTN_C128 1

#define QTN_OK 0
#define QTN_ERR_ARG (-1)
#define QTN_ERR_CUDA (-2)
#define QTN_ERR_UNSUPPORTED (-3)
#define QTN_ERR_WORKSPACE (-4)

#define QTN_MAX_RANK 48   /* labels per bit tensor */
#define QTN_MAX_SLICE 24  /* sliced labels touching one operand */

const char *qtn_version(void);
const char *qtn_last_error(void);
/* Device facts the host planner sizes things with (SM count, max opt-in smem). */
int qtn_device_info(int device, int *sm_count, int *max_smem_optin, int *cc_major, int *cc_minor);

/* ------------------------------------------------------------------------
 * Pairwise contraction of two bit tensors, permutation fused into the loads
 * and stores (no separate transpose pass).
 *
 * Takes over: every pairwise step inside cuquantum.tensornet.contract /
 * Network.contract (eval.py:265,297,313,375,444,476) -- K1, K2, K5 of SURVEY.md.
 *
 * Label roles are derived from membership: in A,B,C = batch (hyper index);
 * in A,B only = contracted; in A,C = free-left; in B,C = free-right.
 * ------------------------------------------------------------------------ */
typedef struct {
  int32_t dtype;                 /* QTN_C64 | QTN_C128 */
  int32_t rank_a, rank_b, rank_c;
  int32_t label_a[QTN_MAX_RANK];
  uint8_t pos_a[QTN_MAX_RANK];
  int32_t label_b[QTN_MAX_RANK];
  uint8_t pos_b[QTN_MAX_RANK];
  int32_t label_c[QTN_MAX_RANK]; /* order irrelevant when c_layout_free */
  uint8_t pos_c[QTN_MAX_RANK];   /* in: prescribed layout; out (c_layout_free): chosen layout */
  int32_t c_layout_free;         /* 1: let the planner pick C's layout (dense, rank_c bits) */
  int32_t conj_a, conj_b;        /* conjugate an operand on load */
  int32_t accumulate;            /* C += A*B instead of C = A*B (slice accumulation) */
  /* Slicing: bit `slice_bit_x[i]` of the run-time slice id selects the half of
   * operand X at address bit `slice_pos_x[i]` (label fixed, absent from label_x). */
  int32_t n_slice_a, n_slice_b;
  uint8_t slice_bit_a[QTN_MAX_SLICE], slice_pos_a[QTN_MAX_SLICE];
  uint8_t slice_bit_b[QTN_MAX_SLICE], slice_pos_b[QTN_MAX_SLICE];
} qtn_pair_desc;

#define QTN_KERNEL_TILE 0   /* shared-memory tiled FFMA/DFMA kernel */
#define QTN_KERNEL_REDUCE 1 /* tiny output, long contracted extent: K-parallel reduction */

#define QTN_LO_MAX 16
#define QTN_HI_MAX 40

/* Launch-ready form of one pairwise contraction.  Produced on the host by
 * qtn_pair_plan (pure function, no GPU needed), consumed by qtn_pair_run.
 * Public so that tests can check the address algebra without a GPU. */
typedef struct {
  int32_t kernel;                /* QTN_KERNEL_* */
  int32_t dtype;
  int32_t swapped;               /* operands exchanged so that the left free extent >= right */
  int32_t tmb, tnb, tkb;         /* log2 tile extents (logical, incl. phantom bits) */
  int32_t m_real, n_real, k_real;/* real (non-phantom) bits among tmb/tnb/tkb */
  int32_t n_mhi, n_nhi, n_khi, n_batch;
  int32_t split_bits;            /* split-K: log2 partial sums, summed by a second launch */
  int32_t conj_a, conj_b, accumulate;
  /* global -> shared enumeration, sorted by ascending address bit (phantoms last, pos 255) */
  int32_t na_lo, nb_lo;
  uint8_t a_lo_pos[QTN_LO_MAX];
  uint16_t a_lo_sstr[QTN_LO_MAX];/* shared-memory stride (elements) of that label */
  uint8_t b_lo_pos[QTN_LO_MAX];
  uint16_t b_lo_sstr[QTN_LO_MAX];
  int32_t sa_stride, sb_stride;  /* row (k) strides of the A/B tiles in shared memory */
  /* epilogue: staging index of logical m/n bits, and C positions in ascending order */
  int32_t nc_lo;                 /* real C tile bits */
  uint8_t c_lo_pos[QTN_LO_MAX];
  uint16_t c_m_sidx[QTN_LO_MAX];
  uint16_t c_n_sidx[QTN_LO_MAX];
  /* tile / loop index bit i -> address bit */
  uint8_t a_mhi[QTN_HI_MAX], c_mhi[QTN_HI_MAX];
  uint8_t b_nhi[QTN_HI_MAX], c_nhi[QTN_HI_MAX];
  uint8_t a_khi[QTN_HI_MAX], b_khi[QTN_HI_MAX];
  uint8_t a_bat[QTN_LO_MAX], b_bat[QTN_LO_MAX], c_bat[QTN_LO_MAX];
  int32_t n_slice_a, n_slice_b;
  uint8_t slice_bit_a[QTN_MAX_SLICE], slice_pos_a[QTN_MAX_SLICE];
  uint8_t slice_bit_b[QTN_MAX_SLICE], slice_pos_b[QTN_MAX_SLICE];
  /* REDUCE kernel: logical k bit i -> address bits; m/n (<= 4 bits together) */
  int32_t red_chunk_bits;        /* k values per CTA = 2^red_chunk_bits */
  /* launch */
  int64_t grid;                  /* CTAs of the main launch */
  int32_t block;
  int32_t smem_bytes;
  int64_t workspace_bytes;       /* split-K / reduction partials; 0 = none */
  int64_t c_elems;               /* 2^rank_c */
  /* accounting (SURVEY.md section 8d conventions) */
  double flops;                  /* 8 * 2^(m+n+k+batch) */
  double bytes;                  /* elem * (|A| + |B| + |C|) */
} qtn_pair_plan_t;

/* Host-only: derive roles, tile shape, coalescing-driven label placement and,
 * when desc->c_layout_free, C's layout (written back to desc->pos_c).
 * `sm_count` sizes split-K (pass 148 for B200; <=0 means 148). */
int qtn_pair_plan(qtn_pair_desc *desc, int sm_count, qtn_pair_plan_t *plan);

/* Launch.  `slice_id` is a DEVICE pointer to one int32 (may be NULL when the
 * plan has no sliced labels): keeping it on the device lets a captured CUDA
 * graph be replayed for every slice.  `workspace` must hold plan->workspace_bytes. */
int qtn_pair_run(const qtn_pair_plan_t *plan, const void *a, const void *b, void *c,
                 void *workspace, const int32_t *slice_id, void *stream);

/* ------------------------------------------------------------------------
 * Bit-tensor permute: out[label -> pos_out] = in[label -> pos_in].
 * Shared-memory staged so that both the read and the write are coalesced.
 * Takes over the implicit transposes cuTENSOR does inside `contract` and the
 * final re-ordering to the requested output modes (circuit_convertor.py:52-57).
 * Algorithmic traffic: 2 * 2^rank * sizeof(element) bytes.
 * ------------------------------------------------------------------------ */
int qtn_permute_bits(int dtype, int rank, const uint8_t *pos_in, const uint8_t *pos_out,
                     const void *in, void *out, int accumulate, void *stream);

/* General strided permute of an arbitrary-extent tensor (rank <= 8), used for
 * MPS site tensors whose bonds are not powers of two.  perm[i] = input axis
 * that becomes output axis i; both tensors dense row-major. */
int qtn_permute(int dtype, int rank, const int64_t *extent_in, const int32_t *perm,
                const void *in, void *out, void *stream);

/* y = alpha * x (+ y) on complex vectors; increments slice id: small utilities
 * that keep whole slice loops on the device. */
int qtn_axpy(int dtype, int64_t n, double alpha_re, double alpha_im, const void *x, void *y,
             int accumulate, void *stream);
int qtn_set_i32(int32_t *dst, int32_t value, void *stream);
int qtn_add_i32(int32_t *dst, int32_t delta, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QIBOTN_B200_H */
