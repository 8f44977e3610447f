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
#define QTN_KERNEL_TC 2     /* complex64, large tiles: tcgen05 TF32x3 MMAs, accumulators in TMEM */
#define QTN_KERNEL_KRED 3   /* <= 16 x 16 outputs, long contracted extent: streaming K reduction */

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
  uint8_t a_bat[QTN_HI_MAX], b_bat[QTN_HI_MAX], c_bat[QTN_HI_MAX];   /* batch (hyper) labels */
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
  /* QTN_KERNEL_TC only: shared-memory pipeline depth and the byte sizes of one A / one B
   * operand plane (a stage holds 4 A planes then 4 B planes: re_hi, re_lo, im_hi, im_lo) */
  int32_t tc_stages;
  int32_t tc_plane_a, tc_plane_b;
  int32_t tc_reserved;           /* > 0: persistent variant, number of CTAs to launch */
  /* iteration part of the tile enumerations (bits 8.. of the element index): global offset and
   * shared-memory offset of iteration i; kernel parameters, i.e. constant-bank operands */
  int64_t tc_a_it_g[8], tc_b_it_g[16];
  int32_t tc_a_it_s[8], tc_b_it_s[16];
  int32_t tc_terms;              /* 3: hi*hi + hi*lo + lo*hi;  4: + lo*lo ("tc_terms" option) */
  int32_t tc_pad;                /* 1: cross terms accumulate in their own TMEM region ("tc_accum_split") */
  /* B pre-packing: when many m tiles share each B tile, a small pre-kernel writes B once, already
   * split and laid out as shared-memory stage images, into the workspace at byte offset
   * tc_prepack_off; the main kernel then fetches a whole B stage with one cp.async.bulk. */
  int32_t tc_prepack;
  int32_t tc_pair;               /* 1: CTA-pair kernel (tcgen05 cta_group::2), tc_reserved even ("tc_pair");
                                  * 2: staggered CTA-pair kernel for 128-column tiles ("tc_stagger") */
  int64_t tc_prepack_off;
  /* 0: error-compensated 3xTF32 split (12 kind::tf32 MMAs per 8 contracted complex values);
   * 1: scaled 3xFP16 split (12 kind::f16 MMAs per 16 values: half the tensor-pipe cycles, the same
   *    22 significant bits): operands are scaled by a power of two taken from the run-time magnitude
   *    bound of each tensor, so the plan must be run with qtn_pair_run_scaled ("tc_fp16"). */
  int32_t tc_kind;
  /* 1: the staggered kernel gathers A in ascending ADDRESS order of the tile's labels (a warp's 32 copies of
   * 8 bytes cover the 5 lowest address bits among them: runs of up to 256 contiguous bytes) into the raw
   * shared-memory ring and the conversion to operand planes reads the ring across threads ("tc_gather");
   * 0: every thread gathers the elements it converts itself (lanes own the bank-selecting plane offsets,
   * whatever their addresses).  tc_a_gbit[r] = address bit of the tile label of rank r in address order,
   * tc_a_rank[j] = that rank for enumeration bit j of a_lo_pos; tc_a_it_gg / tc_a_it_raw = global offset and
   * raw-ring element index of iteration i in the gather order / of the converted element. */
  int32_t tc_gather;
  uint8_t tc_a_gbit[QTN_LO_MAX], tc_a_rank[QTN_LO_MAX];
  int64_t tc_a_it_gg[8];
  int32_t tc_a_it_raw[8];
  /* QTN_KERNEL_TILE: shared-memory stages of (A tile, B tile); 2 = the next k step is copied (cp.async) while the
   * current one is multiplied */
  int32_t tile_stages;
  /* 1: the staggered fp16 kernel reads A as pre-packed stage images too ("tc_prepack_a"): a pre-kernel gathers,
   * scales and splits A once per launch into images laid out exactly as the four A planes of a shared-memory stage
   * (one image per m tile and k chunk, 128 x 16 complex values = 16 KB: as large as A itself), at workspace offset
   * tc_prepack_a_off; the main kernel's producers shrink to one thread issuing two bulk copies per stage.  Chosen
   * when >= 4 n tiles read every A tile (A would otherwise be gathered and converted once per n tile). */
  /* QTN_KERNEL_KRED: number of INNER batch labels (batch labels at low address bits, enumerated inside the CTA as
   * a second row dimension of the shared-memory chunk instead of by grid.y) and their C positions */
  int32_t kred_inner;
  uint8_t kred_c_in[4];
  int32_t tc_prepack_a;
  int32_t tc_run_bits;           /* tc_gather == 2: log2 of the contiguous run the tile's labels own at the bottom of A */
  int64_t tc_prepack_a_off;
} qtn_pair_plan_t;

/* Per-call options.  Every switch below can be handed to the planner / the SVD driver explicitly
 * (qtn_pair_plan_opts, qtn_svd_rows_opts): nothing a caller plans or runs that way reads process-wide state,
 * so concurrent callers with different settings do not interfere.  qtn_options_default fills in the library
 * defaults; the plain entry points (qtn_pair_plan, qtn_svd_rows) use the process-wide copy that qtn_set_option
 * edits (a convenience for tests and A/B runs: set it before threads start). */
typedef struct {
  int32_t tc_enable, tc_min_log2, tc_persistent, tc_terms, tc_accum_split, tc_prepack, tc_pair, tc_stagger, tc_fp16;
  int32_t svd_wide, svd_group_mb, svd_reg;
  int32_t tc_gather;
  int32_t tc_prepack_a;
  int32_t tc_gather_min_run;     /* tc_gather = 2: smallest run (log2 elements) fetched by bulk copies; default 10 = 8 KB */
  int32_t reserved[1];
} qtn_options;
void qtn_options_default(qtn_options *opt);     /* the built-in defaults */
void qtn_options_current(qtn_options *opt);     /* the process-wide copy (defaults + qtn_set_option edits) */

/* Library-wide switches (thread-unsafe, meant for tests and A/B measurements):
 *   "tc_enable"    1 (default) lets qtn_pair_plan choose QTN_KERNEL_TC, 0 forbids it;
 *   "tc_min_log2"  smallest m+n+k (log2 of the MAC count) for which TC is chosen (default 18);
 *   "tc_terms"     3 (default) or 4: products kept by the TF32 split (4 adds lo*lo, +33% MMAs);
 *   "tc_accum_split" 1 (default): hi*hi and the cross terms use separate TMEM accumulators, added in
 *                  the epilogue (3x fewer truncating accumulations on the large sum; tile <= 128 columns);
 *   "tc_prepack"   1 (default): pre-pack the small operand once per launch (see tc_prepack above);
 *   "tc_pair"      5 (default): tiles of 2^5 .. 128 columns with a pre-packed B run on CTA pairs
 *                  (tcgen05 cta_group::2: two m tiles share one B tile, each SM stages half of it);
 *                  n = 6, 7 raises the smallest paired tile to 2^n columns, 0 turns pairs off;
 *   "tc_stagger"   1 (default): 128-column paired tiles run as two 64-column halves with their own accumulators
 *                  that share every A stage (plan value tc_pair = 2): the TMEM drain of one half overlaps the
 *                  MMAs of the other; 0: the round-1 pair kernel (tile and accumulators as one unit);
 *   "tc_fp16"      1 (default): staggered tiles split their operands into scaled fp16 pieces (tc_kind = 1,
 *                  half the tensor-pipe work of the TF32 split; run through qtn_pair_run_scaled); 0: TF32 split;
 *   "tc_gather"    0: every producer thread gathers the elements it converts;
 *                  2 (default): when the labels of a tile own at least the "tc_gather_min_run" (default 10) lowest
 *                  address bits of A (runs of >= 8 KB: A is often the output tile layout of an earlier node), a
 *                  chunk is fetched into the raw ring by one cp.async.bulk per run (no LSU gather) and converted
 *                  across threads -- a gain on long runs (m25 n6 k6: 9.0 -> 8.1 ms), neutral at 2-4 KB, a loss on
 *                  shorter ones (profiles/r02_gather_bulk_ab.jsonl), hence the threshold;
 *                  1: A is gathered in ascending address order (contiguous runs per warp) and the raw elements
 *                  change hands across threads through shared memory (an mbarrier per raw chunk + one named
 *                  barrier per chunk).  Measured 5-20 % slower (profiles/r02_gather_ab.jsonl): fewer LSU
 *                  wavefronts, but the producers are bound by their serial wait -> convert -> store chain, which
 *                  the extra hand-off lengthens;
 *   "tc_prepack_a" 1 (default): staggered fp16 plans whose A tiles are read by >= 4 n tiles pre-pack A into stage
 *                  images as well (see tc_prepack_a above); 0: A is always gathered by the producers; 2: always
 *                  pre-packed (measurement switch: the extra pass over A does not pay with fewer n tiles);
 *   "tc_persistent" 1 (default): one CTA per SM streams over the tiles with loads, MMAs and
 *                  epilogue of neighbouring tiles overlapped; 0: one CTA per tile.
 *   "svd_wide"     1 (default): qtn_svd_rows runs 512-thread CTAs holding twice the block rows (200 KB of
 *                  shared memory, row pairs split across warps); 0: 256-thread CTAs, 160 KB.
 *   "svd_reg"      which kernel rotates the block pairs (8 + 8 rows) of qtn_svd_rows matrices whose rows have
 *                  <= 512 (complex128) / 1024 (complex64) columns:
 *                  4 (default) two-pass kernel within 64 registers, two CTAs per SM, 2 warps per row pair;
 *                  5 the same with one warp per row pair; 3 persistent streaming kernel (cp.async prefetch of the
 *                  next block pair); 1 / 2 first register-resident kernel with 2 / 4 warps per row pair;
 *                  0 shared-memory kernel only (also used for longer rows, the pairs inside a block and
 *                  matrices that fit one block pair).  All give the same singular values (tests/test_gpu_mps.py).
 *   "svd_group_mb" qtn_svd_rows sweeps the batch in groups of at most this many MiB of rows, so that a
 *                  group stays L2-resident over the launches of a sweep; 0 = the whole batch at once.
 * Returns QTN_ERR_ARG for an unknown name. */
int qtn_set_option(const char *name, int value);
int qtn_get_option(const char *name, int *value);
/* sizeof() of a public struct by name ("qtn_pair_desc", "qtn_pair_plan_t", "qtn_svd_item",
 * "qtn_svd_trunc", "qtn_gemm_item", "qtn_options"), -1 if unknown: lets a foreign-language binding verify its mirror. */
int64_t qtn_sizeof(const char *struct_name);

/* Host-only: derive roles, tile shape, coalescing-driven label placement and,
 * when desc->c_layout_free, C's layout (written back to desc->pos_c).
 * `sm_count` sizes split-K (pass 148 for B200; <=0 means 148). */
int qtn_pair_plan(qtn_pair_desc *desc, int sm_count, qtn_pair_plan_t *plan);
/* The same with explicit options (NULL = the process-wide copy). */
int qtn_pair_plan_opts(qtn_pair_desc *desc, int sm_count, const qtn_options *opt, qtn_pair_plan_t *plan);

/* Launch.  `slice_id` is a DEVICE pointer to one int32 (may be NULL when the
 * plan has no sliced labels): keeping it on the device lets a captured CUDA
 * graph be replayed for every slice.  `workspace` must hold plan->workspace_bytes. */
int qtn_pair_run(const qtn_pair_plan_t *plan, const void *a, const void *b, void *c,
                 void *workspace, const int32_t *slice_id, void *stream);

/* qtn_pair_run with magnitude tracking.  Every tensor may carry a DEVICE word `amax`: the bit pattern
 * of a non-negative float that bounds max(|re|, |im|) over the tensor (0 = unknown/empty).  Plans with
 * tc_kind == 1 need amax_a and amax_b (the operands are scaled into the fp16 range by
 * 2^(14 - exponent(amax))); any plan updates *amax_c (atomic max over the words it was given) when
 * amax_c is not NULL -- inside the epilogue of the tensor-core kernels, by one extra streaming pass over C
 * for the others.  complex128 plans ignore all three.  The caller zeroes a word before its tensor is
 * (re)computed; a stale larger bound only costs precision below 2^-28 of that bound. */
int qtn_pair_run_scaled(const qtn_pair_plan_t *plan, const void *a, const void *b, void *c, void *workspace,
                        const int32_t *slice_id, const uint32_t *amax_a, const uint32_t *amax_b,
                        uint32_t *amax_c, void *stream);
/* *amax = max(*amax, bits(max over x of max(|re|, |im|))) for a complex64 vector of n elements. */
int qtn_absmax(int dtype, int64_t n, const void *x, uint32_t *amax, void *stream);

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


/* ------------------------------------------------------------------------
 * MPS path (arbitrary extents, dense row-major tensors).
 * Site tensors are (chi_l, 2, chi_r) as in mps_utils.py:48-50; a two-site block
 * theta is the (2*chi_l) x (2*chi_r) matrix with rows (i,p) and columns (q,k).
 * ------------------------------------------------------------------------ */

/* C[m x n] = opA(A) * opB(B) (+ C).  op bit 0 = conjugate, bit 1 = the stored matrix is the
 * transpose of the operand (so 0 = N, 1 = conj, 2 = T, 3 = H).  lda/ldb/ldc in elements.
 * Takes over the contraction half of contract_decompose (mps_utils.py:78-84: A.B), the chain
 * contractions of mps_contraction_helper.py:115-118 and the transfer-matrix sweeps of :32-51,73-113. */
#define QTN_OP_N 0
#define QTN_OP_C 1
#define QTN_OP_T 2
#define QTN_OP_H 3
int qtn_gemm(int dtype, int opa, int opb, int64_t m, int64_t n, int64_t k, const void *a, int64_t lda,
             const void *b, int64_t ldb, void *c, int64_t ldc, int accumulate, void *stream);

/* `count` independent products C_i = opA_i(A_i) * opB_i(B_i) (+ C_i) in ceil(count / 32) launches: the A.B
 * products and the factor recoveries of one brickwork layer of two-site gates (mps_utils.py:78-84 is called once
 * per gate by the reference; disjoint bonds of a layer are independent).  `items` is a HOST array. */
typedef struct {
  const void *a, *b;
  void *c;
  int32_t m, n, k;
  int32_t lda, ldb, ldc;
  int32_t opa, opb;
} qtn_gemm_item;
int qtn_gemm_batched(int dtype, int count, const qtn_gemm_item *items, int accumulate, void *stream);

/* site[i][q][j] = sum_p gate[q][p] site[i][p][j] in place; gate = 4 device elements.
 * Takes over contract("ipj,qp->iqj") (mps_utils.py:64-69). */
int qtn_mps_apply_1q(int dtype, int64_t chi_l, int64_t chi_r, void *site, const void *gate, void *stream);

/* theta[(i,r),(s,k)] = sum_pq gate[r][s][p][q] theta[(i,p),(q,k)] in place; gate = 16 device
 * elements.  With the SWAP matrix this is the leg exchange "ipj,jqk->iqj,jpk" (mps_utils.py:32-37). */
int qtn_mps_apply_2q(int dtype, int64_t chi_l, int64_t chi_r, void *theta, const void *gate, void *stream);

/* One site of perfect sampling from an MPS, all shots at once -- what `nshots` of the quimb platform's
 * execute_circuit asks of the state (backends/quimb.py:172-186: Counter(circ.sample(nshots)), then
 * |amplitude|^2 of the most frequent strings) and what result.py:50-57 `frequencies()` hands out, without
 * the dense vector.  For the site with tensor A (chi_l, 2, r) the caller has computed, with qtn_gemm,
 * w_p = v A^p (nshots x r; v = the shots' left boundary vectors, normalised so that v R_prev v^H = 1) and
 * z_p = w_p R (R = right environment of the bond to the right, r x r).  The kernel forms the outcome weights
 * q_p[s] = Re <z_p[s,:], w_p[s,:]>, draws bit[s] = (u[s] (q_0 + q_1) >= q_0), writes it to bits[s * bit_stride],
 * multiplies ptotal[s] by the conditional probability q_bit / (q_0 + q_1) and emits the next boundary
 * vectors vnext[s,:] = w_bit[s,:] / sqrt(q_bit).  u: nshots uniforms in [0, 1) (device, double). */
int qtn_mps_sample_step(int dtype, int64_t nshots, int64_t r, const void *w0, const void *w1, const void *z0,
                        const void *z1, const double *u, uint8_t *bits, int64_t bit_stride, double *ptotal,
                        void *vnext, void *stream);

/* Batched one-sided Jacobi SVD with truncation: the decompose half of contract_decompose
 * (mps_utils.py:32,78; cuTensorNet SVDMethod semantics: abs_cutoff, rel_cutoff,
 * discarded_weight_cutoff, max_extent, normalization, partition). */
typedef struct {
  void *w;                /* device, n x m row-major, n <= m advised; rows orthogonalised IN PLACE */
  int32_t n, m;
  int32_t block, nblocks; /* out: blocking chosen by the library */
} qtn_svd_item;

typedef struct {
  double abs_cutoff, rel_cutoff, discarded_weight_cutoff;
  int32_t max_extent;     /* 0 = unlimited */
  int32_t normalization;  /* 0 none, 1 L1, 2 L2, 3 LInf (applied to the kept singular values) */
} qtn_svd_trunc;

int64_t qtn_svd_workspace_bytes(int batch, int64_t ld);
/* LQ preconditioner for qtn_svd_rows (the QR preconditioning of Drmac-Veselic's Jacobi SVD, which
 * cuSOLVER's gesvdj-class routines behind contract_decompose use as well): right-looking modified
 * Gram-Schmidt over the ROWS of each items[i].w (n x m, n <= m, overwritten), keeping only
 * X = L^H (n x n) in xitems[i].w (xitems[i].n = xitems[i].m = items[i].n).  theta = L Q and L share
 * singular values and left vectors, so qtn_svd_rows is then run on `xitems`: its rows converge to
 * sigma_i z_i^H with Z the LEFT singular vectors of the input -- see qtn_svd_emit. */
int64_t qtn_svd_lq_workspace_bytes(int batch);
int qtn_svd_lq(int dtype, int batch, const qtn_svd_item *items, const qtn_svd_item *xitems,
               void *workspace, int64_t workspace_bytes, void *stream);
/* On return every W holds J*W_in with mutually orthogonal rows; sigma[i*ld + c] is the c-th
 * largest row norm (= singular value) of matrix i and perm[i*ld + c] the row that carries it;
 * rank_host[i] = number of singular values kept, scale_host[i] = factor the kept values are
 * multiplied by (normalization).  `items` is a HOST array (its block fields are filled in);
 * sigma/perm/workspace are device buffers.  Synchronises `stream` (one read-back per sweep). */
int qtn_svd_rows(int dtype, int batch, qtn_svd_item *items, double tol, int max_sweeps,
                 const qtn_svd_trunc *trunc, double *sigma, int32_t *perm, int64_t ld,
                 int32_t *rank_host, double *scale_host, int32_t *sweeps_host, void *workspace,
                 int64_t workspace_bytes, void *stream);
/* qtn_svd_rows with explicit options (svd_wide, svd_reg, svd_group_mb; NULL = the process-wide copy). */
int qtn_svd_rows_opts(int dtype, int batch, qtn_svd_item *items, double tol, int max_sweeps,
                      const qtn_svd_trunc *trunc, double *sigma, int32_t *perm, int64_t ld,
                      int32_t *rank_host, double *scale_host, int32_t *sweeps_host, void *workspace,
                      int64_t workspace_bytes, const qtn_options *opt, void *stream);
/* Scaled copies of the k leading rows (sorted order) of an orthogonalised W (m columns):
 * with s = sigma[c], sn = s*scale and (fa, fb) the share of sn absorbed by the left / right factor
 * (partition 0 none: (1,1); 1 "U": (sn,1); 2 "V": (1,sn); 3 "UV": (sqrt sn, sqrt sn)):
 *   transposed == 0 (W_in = theta):    out1 = fb/s * row   (the right factor, k x m)
 *                                      out2 = fa/s^2 * row (left factor = theta * out2^H)
 *   transposed == 1 (W_in = theta^T):  out1 = fa/s * row   (left factor transposed, k x m)
 *                                      out2 = fb/s^2 * row (right factor = conj(out2) * theta)
 * `transposed` + 2 additionally conjugates out1 (used with qtn_svd_lq, where the rows are
 * sigma * z^H of the LEFT vectors).
 * sigma/perm point at this matrix's row of the arrays qtn_svd_rows filled. */
int qtn_svd_emit(int dtype, const void *w, int32_t m, const int32_t *perm, const double *sigma,
                 double scale, int32_t k, int32_t partition, int32_t transposed, void *out1, void *out2,
                 void *stream);

/* ------------------------------------------------------------------------
 * Fused Pauli-string expectation on a dense state (qubit 0 = most significant address bit):
 * out[t] = <psi| P_t |psi> (complex128 pairs, device) for nterms strings given as bit masks over
 * the ADDRESS bits: xmask = qubits with X or Y, zmask = qubits with Z or Y, ny = number of Y.
 * One streaming pass over psi per term, P|psi> is never stored.  Serves the per-term loop of
 * expectation_tn (eval.py:473-476) when the dense state fits in memory.
 * ------------------------------------------------------------------------ */
int qtn_pauli_expectation(int dtype, int nqubits, const void *state, int nterms, const uint64_t *xmask,
                          const uint64_t *zmask, const int32_t *ny, void *out, void *stream);
/* The same with the masks in HOST memory (xmask, zmask, ny: host arrays; state, out: device): terms that share
 * an xmask are evaluated up to 16 at a time in ONE pass over psi (the pair product conj(psi[x ^ xmask]) psi[x]
 * is formed once, only the Z sign differs per term), so a sum of ZZ terms reads psi ceil(nterms / 16) times
 * instead of nterms times.  Algorithmic traffic per pass as above. */
int qtn_pauli_expectation_host(int dtype, int nqubits, const void *state, int nterms, const uint64_t *xmask,
                               const uint64_t *zmask, const int32_t *ny, void *out, void *stream);

/* ------------------------------------------------------------------------
 * Measurement helper (bench.py): achieved rate of the FP32 (is_double = 0) or FP64 FMA pipe with 16
 * independent FMA chains per thread, 2 flops per FMA, in TFLOP/s -- the roofline denominator of the kernels
 * that run on those pipes (Jacobi rotations, FFMA/DFMA contractions).  `scratch`: >= 148*8*256*8 device
 * bytes.  Synchronises `stream`.
 * ------------------------------------------------------------------------ */
int qtn_microbench_fma(int is_double, int iters, void *scratch, int64_t scratch_bytes, double *tflops,
                       void *stream);

#ifdef __cplusplus
}
#endif
#endif /* QIBOTN_B200_H */
