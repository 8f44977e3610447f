// Host-side planning of one pairwise bit-tensor contraction: label roles,
// tile shape, which labels live inside a tile (chosen so that global loads and
// stores walk contiguous addresses), C's layout when the caller leaves it free,
// split-K, launch geometry.  Pure function of the descriptor: no CUDA calls,
// so it is unit-tested on CPU-only machines.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "qtn_internal.h"

namespace {

struct Lab {
  int label;
  int pa = -1, pb = -1, pc = -1;  // address bit positions (-1: absent)
};

int find_label(const int32_t *labels, int n, int label) {
  for (int i = 0; i < n; ++i)
    if (labels[i] == label) return i;
  return -1;
}

constexpr int kPhantom = 255;
// Largest set of pre-packed B images (bytes of the 3xTF32 form; the fp16 form is 3/4 of it): B is re-laid once per
// launch at 1.5-2x its size, which pays whenever >= 8 m tiles read every image -- also for a B of many GB (the deep
// hyper-indexed nodes of QAOA light cones: both operands 2^31 elements).  The workspace is sized by the plan.
constexpr int64_t kMaxImageBytes = (int64_t)1 << 35;

}  // namespace

extern "C" int qtn_pair_plan(qtn_pair_desc *d, int sm_count, qtn_pair_plan_t *p) {
  return qtn_pair_plan_opts(d, sm_count, nullptr, p);
}

extern "C" int qtn_pair_plan_opts(qtn_pair_desc *d, int sm_count, const qtn_options *given, qtn_pair_plan_t *p) {
  if (!d || !p) return qtn_fail(QTN_ERR_ARG, "qtn_pair_plan: null argument");
  qtn_options opt;
  qtn_options_resolve(given, &opt);
  if (d->dtype != QTN_C64 && d->dtype != QTN_C128)
    return qtn_fail(QTN_ERR_ARG, "qtn_pair_plan: dtype must be QTN_C64 or QTN_C128");
  if (d->rank_a < 0 || d->rank_b < 0 || d->rank_c < 0 || d->rank_a > QTN_MAX_RANK ||
      d->rank_b > QTN_MAX_RANK || d->rank_c > QTN_MAX_RANK)
    return qtn_fail(QTN_ERR_ARG, "qtn_pair_plan: rank out of range");
  if (d->n_slice_a < 0 || d->n_slice_b < 0 || d->n_slice_a > QTN_MAX_SLICE ||
      d->n_slice_b > QTN_MAX_SLICE)
    return qtn_fail(QTN_ERR_ARG, "qtn_pair_plan: too many sliced labels on one operand");
  if (sm_count <= 0) sm_count = 148;
  std::memset(p, 0, sizeof(*p));

  // ---- swap so that the left free extent is the larger one -----------------
  int free_a = 0, free_b = 0;
  for (int i = 0; i < d->rank_a; ++i)
    if (find_label(d->label_b, d->rank_b, d->label_a[i]) < 0) ++free_a;
  for (int i = 0; i < d->rank_b; ++i)
    if (find_label(d->label_a, d->rank_a, d->label_b[i]) < 0) ++free_b;
  const bool swap = free_b > free_a;
  const int32_t *la = swap ? d->label_b : d->label_a;
  const uint8_t *pa = swap ? d->pos_b : d->pos_a;
  const int ra = swap ? d->rank_b : d->rank_a;
  const int32_t *lb = swap ? d->label_a : d->label_b;
  const uint8_t *pb = swap ? d->pos_a : d->pos_b;
  const int rb = swap ? d->rank_a : d->rank_b;
  p->swapped = swap;
  p->dtype = d->dtype;
  p->conj_a = swap ? d->conj_b : d->conj_a;
  p->conj_b = swap ? d->conj_a : d->conj_b;
  p->accumulate = d->accumulate;
  p->n_slice_a = swap ? d->n_slice_b : d->n_slice_a;
  p->n_slice_b = swap ? d->n_slice_a : d->n_slice_b;
  std::memcpy(p->slice_bit_a, swap ? d->slice_bit_b : d->slice_bit_a, QTN_MAX_SLICE);
  std::memcpy(p->slice_pos_a, swap ? d->slice_pos_b : d->slice_pos_a, QTN_MAX_SLICE);
  std::memcpy(p->slice_bit_b, swap ? d->slice_bit_a : d->slice_bit_b, QTN_MAX_SLICE);
  std::memcpy(p->slice_pos_b, swap ? d->slice_pos_a : d->slice_pos_b, QTN_MAX_SLICE);

  // ---- classify labels ------------------------------------------------------
  std::vector<Lab> M, N, K, Bt;
  for (int i = 0; i < ra; ++i) {
    Lab l;
    l.label = la[i];
    l.pa = pa[i];
    for (int j = 0; j < i; ++j)
      if (la[j] == la[i]) return qtn_fail(QTN_ERR_ARG, "qtn_pair_plan: repeated label in an operand");
    int jb = find_label(lb, rb, l.label);
    int jc = find_label(d->label_c, d->rank_c, l.label);
    if (jb >= 0) l.pb = pb[jb];
    if (jc >= 0) l.pc = d->c_layout_free ? -1 : d->pos_c[jc];
    if (jb >= 0 && jc >= 0) Bt.push_back(l);
    else if (jb >= 0) K.push_back(l);
    else if (jc >= 0) M.push_back(l);
    else return qtn_fail(QTN_ERR_ARG, "qtn_pair_plan: label of A is neither in B nor in C");
  }
  for (int i = 0; i < rb; ++i) {
    if (find_label(la, ra, lb[i]) >= 0) continue;
    Lab l;
    l.label = lb[i];
    l.pb = pb[i];
    int jc = find_label(d->label_c, d->rank_c, l.label);
    if (jc < 0) return qtn_fail(QTN_ERR_ARG, "qtn_pair_plan: label of B is neither in A nor in C");
    l.pc = d->c_layout_free ? -1 : d->pos_c[jc];
    N.push_back(l);
  }
  if ((int)(M.size() + N.size() + Bt.size()) != d->rank_c)
    return qtn_fail(QTN_ERR_ARG, "qtn_pair_plan: C has labels found in neither operand");
  const int mbits = (int)M.size(), nbits = (int)N.size(), kbits = (int)K.size(),
            bbits = (int)Bt.size();
  const int esz = d->dtype == QTN_C64 ? 8 : 16;
  p->flops = 8.0 * std::ldexp(1.0, mbits + nbits + kbits + bbits);
  p->bytes = esz * (std::ldexp(1.0, ra) + std::ldexp(1.0, rb) + std::ldexp(1.0, d->rank_c));
  p->c_elems = (int64_t)1 << d->rank_c;
  p->block = 256;

  std::vector<std::pair<int, int>> assigned;   // (label, chosen C position) when C is free
  // positions below this many bits are "forced" into a tile for coalescing
  const int cforce = d->dtype == QTN_C64 ? 4 : 3;

  // Batch (hyper) labels on the tensor cores: only the staggered CTA-pair kernel with pre-packed B images
  // carries them (a batch value = one more group of m tiles for A and C, one more set of B images), so the
  // plan must be one that lands there: >= 8 m tiles per B image, 32..128 columns, images within 1 GiB.
  auto tc_batch_ok = [&]() {
    if (!(opt.tc_pair && opt.tc_prepack && opt.tc_persistent && opt.tc_accum_split && opt.tc_stagger &&
          opt.tc_terms == 3))
      return false;
    // 16 columns (n = 4) run as a 32-column tile whose upper half is a phantom: zero rows in the B images, never stored
    const int tnb = std::min(std::max(nbits, 5), 7), tkb = std::min(kbits, 4);
    if (mbits < 10 || tnb < opt.tc_pair || opt.tc_pair > 5 || mbits - 7 + bbits > QTN_HI_MAX) return false;
    const int64_t images = (int64_t)1 << std::min(40, (nbits - tnb) + (kbits - tkb) + bbits);
    return images * (4 * (int64_t)(4 << (tnb + tkb))) <= kMaxImageBytes;
  };

  // ---- K-parallel reduction for tiny outputs --------------------------------
  if (bbits == 0 && mbits + nbits <= 4 && kbits >= 10) {
    p->kernel = QTN_KERNEL_REDUCE;
    std::sort(K.begin(), K.end(), [](const Lab &x, const Lab &y) {
      return std::min(x.pa, x.pb) != std::min(y.pa, y.pb) ? std::min(x.pa, x.pb) < std::min(y.pa, y.pb)
                                                          : x.label < y.label;
    });
    std::sort(M.begin(), M.end(), [](const Lab &x, const Lab &y) { return x.pa < y.pa; });
    std::sort(N.begin(), N.end(), [](const Lab &x, const Lab &y) { return x.pb < y.pb; });
    int next_c = 0;
    // C layout when free: n bits lowest, then m bits
    for (auto &l : N) if (d->c_layout_free) { l.pc = next_c++; assigned.push_back({l.label, l.pc}); }
    for (auto &l : M) if (d->c_layout_free) { l.pc = next_c++; assigned.push_back({l.label, l.pc}); }
    p->n_mhi = mbits; p->n_nhi = nbits; p->n_khi = kbits;
    for (int i = 0; i < mbits; ++i) { p->a_mhi[i] = M[i].pa; p->c_mhi[i] = M[i].pc; }
    for (int i = 0; i < nbits; ++i) { p->b_nhi[i] = N[i].pb; p->c_nhi[i] = N[i].pc; }
    for (int i = 0; i < kbits; ++i) { p->a_khi[i] = K[i].pa; p->b_khi[i] = K[i].pb; }
    p->m_real = mbits; p->n_real = nbits; p->k_real = kbits;
    // 2^chunk k values per CTA; aim for >= 4 CTAs per SM but at least 2^11 per CTA
    int chunk = std::max(11, kbits - (int)std::ceil(std::log2((double)sm_count * 8)));
    chunk = std::min(chunk, kbits);
    p->red_chunk_bits = chunk;
    p->grid = (int64_t)1 << (kbits - chunk);
    p->smem_bytes = 0;
    p->workspace_bytes = p->grid * 16 * esz;
  } else if (mbits <= 5 && nbits <= 5 && kbits >= 10 &&
             bbits + std::max(0, mbits - 4) + std::max(0, nbits - 4) <= 15 && kbits >= 7) {
    // ---- streaming K reduction (<= 16 x 16 outputs) ----------------------------
    p->kernel = QTN_KERNEL_KRED;
    // Batch labels that own LOW address bits of an operand would leave every CTA with runs of 8-16 bytes (one batch
    // value per CTA): up to three of them (lowest address first, those below bit 6) become INNER labels -- a second
    // row dimension of the shared-memory chunk, handled by different thread groups of the same CTA -- so that the
    // CTA's gathers stay contiguous.  The chunk keeps its size: 2^q inner values x 2^(tkb) k values.
    std::sort(Bt.begin(), Bt.end(), [](const Lab &x, const Lab &y) {
      return std::min(x.pa, x.pb) != std::min(y.pa, y.pb) ? std::min(x.pa, x.pb) < std::min(y.pa, y.pb) : x.label < y.label;
    });
    std::vector<Lab> Bin;
    while (Bin.size() < 3 && !Bt.empty() && std::min(Bt.front().pa, Bt.front().pb) < 6) {
      Bin.push_back(Bt.front());
      Bt.erase(Bt.begin());
    }
    const int q = (int)Bin.size();
    const int tkb = std::min(kbits, (d->dtype == QTN_C64 ? 7 : 6) - q);   // chunk: 128 / 64 (k value, inner batch value) rows
    p->tmb = 4; p->tnb = 4; p->tkb = tkb;
    p->kred_inner = q;
    std::sort(M.begin(), M.end(), [](const Lab &x, const Lab &y) { return x.pa < y.pa; });
    std::sort(N.begin(), N.end(), [](const Lab &x, const Lab &y) { return x.pb < y.pb; });
    // a fifth m (n) label -- the one with the highest address -- is enumerated by grid.y next to the batch labels:
    // it is absent from B (A), which is simply read once per value of it
    std::vector<Lab> Mx, Nx;
    while ((int)M.size() > 4) { Mx.push_back(M.back()); M.pop_back(); }
    while ((int)N.size() > 4) { Nx.push_back(N.back()); N.pop_back(); }
    const int mt = (int)M.size(), nt = (int)N.size();
    const int ybits = (int)Bt.size() + (int)Mx.size() + (int)Nx.size();      // labels enumerated by grid.y
    p->m_real = mt; p->n_real = nt; p->k_real = tkb;
    // chunk labels = the K labels with the lowest addresses (in either operand): contiguous runs
    std::sort(K.begin(), K.end(), [](const Lab &x, const Lab &y) {
      return std::min(x.pa, x.pb) != std::min(y.pa, y.pb) ? std::min(x.pa, x.pb) < std::min(y.pa, y.pb)
                                                          : x.label < y.label;
    });
    std::vector<Lab> klo(K.begin(), K.begin() + tkb), khi(K.begin() + tkb, K.end());
    std::sort(khi.begin(), khi.end(), [](const Lab &x, const Lab &y) {
      return std::max(x.pa, x.pb) < std::max(y.pa, y.pb);
    });
    if ((int)khi.size() > QTN_HI_MAX) return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_pair_plan: rank too large");
    if (d->c_layout_free) {
      int next_c = 0;
      for (auto &l : N) { l.pc = next_c++; assigned.push_back({l.label, l.pc}); }
      for (auto &l : M) { l.pc = next_c++; assigned.push_back({l.label, l.pc}); }
      for (auto *v : {&Bin, &Nx, &Mx, &Bt})
        for (auto &l : *v) { l.pc = next_c++; assigned.push_back({l.label, l.pc}); }
    }
    for (int i = 0; i < q; ++i) p->kred_c_in[i] = (uint8_t)Bin[i].pc;
    for (int i = 0; i < mt; ++i) { p->a_mhi[i] = M[i].pa; p->c_mhi[i] = M[i].pc; }
    for (int i = 0; i < nt; ++i) { p->b_nhi[i] = N[i].pb; p->c_nhi[i] = N[i].pc; }
    // grid.y: batch (hyper) labels -- one independent reduction per value -- and the surplus m / n labels
    // (position 255 = the operand does not carry the label)
    p->n_batch = ybits;
    {
      int i = 0;
      for (auto &l : Bt) { p->a_bat[i] = l.pa; p->b_bat[i] = l.pb; p->c_bat[i] = l.pc; ++i; }
      for (auto &l : Mx) { p->a_bat[i] = l.pa; p->b_bat[i] = kPhantom; p->c_bat[i] = l.pc; ++i; }
      for (auto &l : Nx) { p->a_bat[i] = kPhantom; p->b_bat[i] = l.pb; p->c_bat[i] = l.pc; ++i; }
    }
    p->n_mhi = 0; p->n_nhi = 0; p->n_khi = (int)khi.size();
    for (int i = 0; i < p->n_khi; ++i) { p->a_khi[i] = khi[i].pa; p->b_khi[i] = khi[i].pb; }
    // shared-memory chunk: element (k, inner batch value bl, row) at ((k << q) + bl) * 16 + row before the kernel's
    // bank swizzle; enumerations in address order
    struct E { int pos; int sstr; };
    std::vector<E> ea, eb;
    for (int i = 0; i < mt; ++i) ea.push_back({M[i].pa, 1 << i});
    for (int i = 0; i < q; ++i) ea.push_back({Bin[i].pa, 16 << i});
    for (int i = 0; i < tkb; ++i) ea.push_back({klo[i].pa, (16 << q) << i});
    for (int i = 0; i < nt; ++i) eb.push_back({N[i].pb, 1 << i});
    for (int i = 0; i < q; ++i) eb.push_back({Bin[i].pb, 16 << i});
    for (int i = 0; i < tkb; ++i) eb.push_back({klo[i].pb, (16 << q) << i});
    auto by_pos = [](const E &x, const E &y) { return x.pos < y.pos; };
    std::sort(ea.begin(), ea.end(), by_pos);
    std::sort(eb.begin(), eb.end(), by_pos);
    p->na_lo = (int)ea.size(); p->nb_lo = (int)eb.size();
    for (int i = 0; i < p->na_lo; ++i) { p->a_lo_pos[i] = (uint8_t)ea[i].pos; p->a_lo_sstr[i] = (uint16_t)ea[i].sstr; }
    for (int i = 0; i < p->nb_lo; ++i) { p->b_lo_pos[i] = (uint8_t)eb[i].pos; p->b_lo_sstr[i] = (uint16_t)eb[i].sstr; }
    // two resident CTAs per SM stride over groups of chunks together (group = blockIdx.x + i * grid);
    // every CTA leaves one partial C in the workspace
    const int64_t ngroups = (int64_t)1 << std::max(0, p->n_khi - 3);   // groups of 8 chunks
    p->split_bits = 0;
    p->grid = std::min<int64_t>(ngroups, std::max<int64_t>(1, ((int64_t)2 * sm_count) >> ybits));
    p->workspace_bytes = ((int64_t)esz << d->rank_c) * p->grid;
    // max(2 buffers x (A + B) chunk, 16 x 256 partials)
    p->smem_bytes = std::max(2 * 2 * (1 << (tkb + q)) * 16 * esz, 16 * 256 * esz);
  } else if (d->dtype == QTN_C64 && d->c_layout_free && (bbits == 0 || tc_batch_ok()) && mbits >= 7 && nbits >= 4 &&
             kbits >= 3 && mbits + nbits + kbits >= opt.tc_min_log2 && opt.tc_enable) {
    // ---- tensor-core kernel (gett_tc.cu) ----------------------------------------
    // Tile = 128 rows of A x 2^tnb rows of B, 2^tkb contracted values per pipeline stage.
    p->kernel = QTN_KERNEL_TC;
    // separate accumulators for the cross terms need 4 TMEM regions per tile: 128 columns at most
    const int sep = opt.tc_accum_split && opt.tc_persistent;
    // batch nodes with 16 columns: a phantom fifth column bit (position 255 in B, none in C) makes the tile 32 wide,
    // the narrowest the staggered kernel takes; its half 1 multiplies zeros and is not stored
    const bool phantom_n = bbits > 0 && nbits == 4;
    if (phantom_n) { Lab ph; ph.label = -1; ph.pb = kPhantom; N.push_back(ph); }
    const int tmb = 7, tnb = std::min(nbits + (phantom_n ? 1 : 0), sep ? 7 : 8), tkb = std::min(kbits, 4);
    const int KT = 1 << tkb, TN = 1 << tnb;
    p->tmb = tmb; p->tnb = tnb; p->tkb = tkb;
    p->block = 288;
    auto pick = [&](std::vector<Lab> &v, int want, auto key, auto forced) {
      std::stable_sort(v.begin(), v.end(), [&](const Lab &x, const Lab &y) {
        bool fx = forced(x), fy = forced(y);
        if (fx != fy) return fx;
        if (key(x) != key(y)) return key(x) < key(y);
        return x.label < y.label;
      });
      std::vector<Lab> lo(v.begin(), v.begin() + want), hi(v.begin() + want, v.end());
      return std::make_pair(lo, hi);
    };
    auto m_sel = pick(M, tmb, [](const Lab &l) { return l.pa; }, [&](const Lab &l) { return l.pa < cforce; });
    auto n_sel = pick(N, tnb, [](const Lab &l) { return l.pb; }, [&](const Lab &l) { return l.pb < cforce; });
    auto k_sel = pick(K, tkb, [](const Lab &l) { return std::min(l.pa, l.pb); },
                      [&](const Lab &l) { return l.pa < cforce || l.pb < cforce; });
    std::vector<Lab> &mlo = m_sel.first, &mhi = m_sel.second;
    std::vector<Lab> &nlo = n_sel.first, &nhi = n_sel.second;
    std::vector<Lab> &klo = k_sel.first, &khi = k_sel.second;
    // logical bit j of the row / k index inside the tile = j-th label in ascending address order
    // of the operand it is read from: the lowest-address labels then land on the 8 rows x 4
    // values of one core matrix, which spreads a warp's shared-memory stores over the banks
    std::sort(mlo.begin(), mlo.end(), [](const Lab &x, const Lab &y) { return x.pa < y.pa; });
    std::sort(nlo.begin(), nlo.end(), [](const Lab &x, const Lab &y) { return x.pb < y.pb; });
    std::sort(klo.begin(), klo.end(), [](const Lab &x, const Lab &y) { return x.pa < y.pa; });
    std::sort(mhi.begin(), mhi.end(), [](const Lab &x, const Lab &y) { return x.pa < y.pa; });
    std::sort(nhi.begin(), nhi.end(), [](const Lab &x, const Lab &y) { return x.pb < y.pb; });
    std::sort(khi.begin(), khi.end(), [](const Lab &x, const Lab &y) {
      return std::max(x.pa, x.pb) < std::max(y.pa, y.pb);
    });
    p->m_real = tmb; p->n_real = tnb - (phantom_n ? 1 : 0); p->k_real = tkb;
    // batch labels: the top bits of the m-tile index for A and C (a_mhi / c_mhi), separate deposits for B
    // (b_bat, read by the image pre-pack); n_batch tells the kernel how many of the m-hi bits they are
    const int true_mhi = (int)mhi.size();
    std::sort(Bt.begin(), Bt.end(), [](const Lab &x, const Lab &y) { return x.pa < y.pa; });
    for (auto &l : Bt) mhi.push_back(l);
    p->n_mhi = (int)mhi.size(); p->n_nhi = (int)nhi.size(); p->n_khi = (int)khi.size();
    p->n_batch = bbits;
    if (p->n_mhi > QTN_HI_MAX || p->n_nhi > QTN_HI_MAX || p->n_khi > QTN_HI_MAX)
      return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_pair_plan: rank too large");
    // C: m bits 0-4 | n bits | m bits 5-6 | n-hi | m-hi  (the epilogue's store order)
    {
      int next_c = 0;
      auto give = [&](Lab &l) { l.pc = next_c++; assigned.push_back({l.label, l.pc}); };
      for (int j = 0; j < 5; ++j) give(mlo[j]);
      for (auto &l : nlo) if (l.label >= 0) give(l);
      for (int j = 5; j < tmb; ++j) give(mlo[j]);
      for (auto &l : nhi) give(l);
      for (auto &l : mhi) give(l);
    }
    // offset (in floats inside one operand plane) of row bit j / k bit i in the canonical
    // K-major no-swizzle layout: (r&7)*4 + (k&3) + (r>>3)*(KT*8) + (k>>2)*32
    auto row_off = [&](int j) { return j < 3 ? 4 << j : (KT * 8) << (j - 3); };
    auto k_off = [&](int i) { return i < 2 ? 1 << i : 32 << (i - 2); };
    struct E { int pos; int sstr; };
    std::vector<E> ea, eb;
    for (int j = 0; j < tmb; ++j) ea.push_back({mlo[j].pa, row_off(j)});
    for (int i = 0; i < tkb; ++i) ea.push_back({klo[i].pa, k_off(i)});
    for (int j = 0; j < tnb; ++j) eb.push_back({nlo[j].pb, row_off(j)});
    for (int i = 0; i < tkb; ++i) eb.push_back({klo[i].pb, k_off(i)});
    // Enumeration order = which label each bit of (lane, warp, iteration) walks.  The five lane
    // bits are the labels that own the five bank-selecting offsets of a core matrix (row bits
    // 0-2 -> 4, 8, 16 floats; k bits 0-1 -> 1, 2 floats): a warp's 32 stores then hit 32
    // different banks.  Those labels are the lowest-address ones of their kind, so the two
    // lowest address bits are always among them and global loads keep >= 32-byte runs.
    auto lane_first = [](std::vector<E> &v) {
      auto is_lane = [](const E &x) { return x.sstr < 32; };
      std::stable_sort(v.begin(), v.end(), [&](const E &x, const E &y) {
        if (is_lane(x) != is_lane(y)) return is_lane(x);
        return x.pos < y.pos;
      });
    };
    lane_first(ea);
    lane_first(eb);
    p->na_lo = (int)ea.size(); p->nb_lo = (int)eb.size();
    for (int i = 0; i < p->na_lo; ++i) { p->a_lo_pos[i] = (uint8_t)ea[i].pos; p->a_lo_sstr[i] = (uint16_t)ea[i].sstr; }
    for (int i = 0; i < p->nb_lo; ++i) { p->b_lo_pos[i] = (uint8_t)eb[i].pos; p->b_lo_sstr[i] = (uint16_t)eb[i].sstr; }
    for (int it = 0; it < 8; ++it) {
      int64_t g = 0; int32_t so = 0;
      for (int j = 8; j < p->na_lo; ++j)
        if ((it >> (j - 8)) & 1) { g |= (int64_t)1 << p->a_lo_pos[j]; so += p->a_lo_sstr[j]; }
      p->tc_a_it_g[it] = g; p->tc_a_it_s[it] = so;
    }
    for (int it = 0; it < 16; ++it) {
      int64_t g = 0; int32_t so = 0;
      for (int j = 8; j < p->nb_lo; ++j)
        if ((it >> (j - 8)) & 1) {
          if (p->b_lo_pos[j] == kPhantom) g = -1; else if (g >= 0) g |= (int64_t)1 << p->b_lo_pos[j];
          so += p->b_lo_sstr[j];
        }
      p->tc_b_it_g[it] = g; p->tc_b_it_s[it] = so;
    }
    p->nc_lo = tmb + tnb;
    for (int j = 0; j < tmb; ++j) p->c_lo_pos[j] = (uint8_t)mlo[j].pc;          // m bits, then n bits
    for (int j = 0; j < tnb; ++j) p->c_lo_pos[tmb + j] = nlo[j].label >= 0 ? (uint8_t)nlo[j].pc : (uint8_t)kPhantom;
    for (int i = 0; i < p->n_mhi; ++i) { p->a_mhi[i] = mhi[i].pa; p->c_mhi[i] = mhi[i].pc; }
    for (int i = 0; i < p->n_nhi; ++i) { p->b_nhi[i] = nhi[i].pb; p->c_nhi[i] = nhi[i].pc; }
    for (int i = 0; i < p->n_khi; ++i) { p->a_khi[i] = khi[i].pa; p->b_khi[i] = khi[i].pb; }
    for (int i = 0; i < bbits; ++i) {
      p->a_bat[i] = mhi[true_mhi + i].pa; p->b_bat[i] = mhi[true_mhi + i].pb; p->c_bat[i] = mhi[true_mhi + i].pc;
    }
    // split-K when there are fewer tiles than SMs (one CTA owns a whole SM here)
    const double tiles = std::ldexp(1.0, p->n_mhi + p->n_nhi);
    int split = 0;
    while (split < p->n_khi - 1 && tiles * (1 << split) < (double)sm_count) ++split;
    p->split_bits = split;
    p->grid = ((int64_t)1 << (p->n_mhi + p->n_nhi)) << split;
    p->workspace_bytes = split ? ((int64_t)esz << (d->rank_c + split)) : 0;
    p->tc_plane_a = 128 * KT * 4;
    p->tc_plane_b = TN * KT * 4;
    const int stage_bytes = 4 * (p->tc_plane_a + p->tc_plane_b);
    // pipeline depth: never deeper than the k loop; small tiles leave room for 2 CTAs per SM
    // (110 KB each) so that one CTA's epilogue overlaps the other's loads
    const int64_t k_iters = (int64_t)1 << (p->n_khi - split);
    const int budget = (tnb <= 6 && 2 * stage_bytes <= 110 * 1024) ? 110 * 1024 : 200 * 1024;
    int stages = std::max(1, std::min(4, budget / stage_bytes));
    if ((int64_t)stages > k_iters) stages = (int)k_iters;
    p->tc_terms = opt.tc_terms;
    p->tc_pad = sep;
    p->tc_reserved = 0;
    if (opt.tc_persistent) {
      // persistent kernel: one CTA per SM streams over tiles, so the pipeline may be deeper than
      // one tile's k loop; tc_reserved carries the number of CTAs to launch
      stages = std::max(2, std::min(6, (200 * 1024) / stage_bytes));
      p->tc_reserved = sm_count;
    }
    p->tc_stages = stages;
    p->smem_bytes = stages * stage_bytes + 256;
    // pre-pack B when at least 8 m tiles reuse each B tile and the image (2 x |B| bytes) is small
    p->tc_prepack = 0;
    p->tc_prepack_off = 0;
    {
      const int64_t images = (int64_t)1 << (p->n_nhi + p->n_khi + bbits);
      const int64_t image_bytes = 4 * (int64_t)p->tc_plane_b;
      if (opt.tc_prepack && p->tc_reserved > 0 && true_mhi >= 3 && images * image_bytes <= kMaxImageBytes) {
        p->tc_prepack = 1;
        p->tc_prepack_off = (p->workspace_bytes + 255) & ~(int64_t)255;
        p->workspace_bytes = p->tc_prepack_off + images * image_bytes;
      }
    }
    // CTA pairs (cta_group::2): two m tiles that share a B tile run on the two SMs of a TPC, each
    // staging half of B -- for the tiles whose MMAs saturate shared-memory bandwidth on one SM
    p->tc_pair = 0;
    if (opt.tc_pair && p->tc_prepack && sep && tnb >= opt.tc_pair && tnb <= 7 && true_mhi >= 1) {
      p->tc_pair = 1;
      const int pair_stage = 4 * (p->tc_plane_a + p->tc_plane_b / 2);
      p->tc_stages = std::max(2, std::min(6, (200 * 1024) / pair_stage));
      p->smem_bytes = p->tc_stages * pair_stage + 256;
      p->tc_reserved = sm_count & ~1;
      // The staggered variant: two half-width halves with their own accumulators share every A stage; the
      // drain of one half hides behind the MMAs of the other.  Per CTA a stage holds the 4 A planes and
      // 3 * TN rows of B (2 halves x {hi, lo} x [Br | Bi | -Br] x TN / 4 rows); an image holds both ranks' rows.
      if (opt.tc_stagger && p->tc_terms == 3) {
        const int rank_b = 3 * TN * KT * 4;
        const int stage2 = 4 * p->tc_plane_a + rank_b;
        const int64_t images = (int64_t)1 << (p->n_nhi + p->n_khi + bbits);
        const int64_t base = (p->tc_prepack_off);
        p->tc_pair = 2;
        // shared memory: operand stages + a ring of 6 raw A chunks (128 x KT elements of 8 bytes: the
        // cp.async gathers in flight) + barriers
        const int raw_ring = 6 * 128 * KT * 8;
        p->tc_stages = std::max(2, std::min(8, (227 * 1024 - 512 - raw_ring) / stage2));
        p->smem_bytes = p->tc_stages * stage2 + raw_ring + 512;
        p->workspace_bytes = base + images * 2 * rank_b;
        // Scaled 3xFP16 split (tc_kind = 1): the same stage geometry in 2-byte elements.  Operand planes are
        // K-major fp16 core matrices (8 rows x 8 values): offset in halves of row bit j / k bit i =
        // (r&7)*8 + (k&7) + (r>>3)*(KT*8) + (k>>3)*64.  A producer thread packs two k-adjacent values
        // into one 32-bit store, so the lowest k bit is the first ITERATION bit of the enumeration and the
        // five lane bits own the five bank-selecting word offsets (k bits 1-2, row bits 0-2).
        if (opt.tc_fp16 && tkb == 4) {
          p->tc_kind = 1;
          auto row_off16 = [&](int j) { return j < 3 ? 8 << j : (KT * 8) << (j - 3); };
          auto k_off16 = [&](int i) { return i < 3 ? 1 << i : 64 << (i - 3); };
          std::vector<E> fa, fb;
          for (int j = 0; j < tmb; ++j) fa.push_back({mlo[j].pa, row_off16(j)});
          for (int i = 0; i < tkb; ++i) fa.push_back({klo[i].pa, k_off16(i)});
          for (int j = 0; j < tnb; ++j) fb.push_back({nlo[j].pb, row_off16(j)});
          for (int i = 0; i < tkb; ++i) fb.push_back({klo[i].pb, k_off16(i)});
          auto cls = [](const E &x) { return x.sstr == 1 ? 2 : (x.sstr <= 32 ? 0 : 1); };   // lane, other, pair bit
          std::stable_sort(fa.begin(), fa.end(), [&](const E &x, const E &y) {
            if ((cls(x) == 0) != (cls(y) == 0)) return cls(x) == 0;
            return x.pos < y.pos;
          });
          // the pair bit (k bit 0) becomes enumeration bit 8 = iteration bit 0
          for (size_t q = 0; q < fa.size(); ++q)
            if (fa[q].sstr == 1) { E e = fa[q]; fa.erase(fa.begin() + q); fa.insert(fa.begin() + 8, e); break; }
          std::stable_sort(fb.begin(), fb.end(), [&](const E &x, const E &y) { return x.pos < y.pos; });
          for (int i = 0; i < p->na_lo; ++i) { p->a_lo_pos[i] = (uint8_t)fa[i].pos; p->a_lo_sstr[i] = (uint16_t)fa[i].sstr; }
          for (int i = 0; i < p->nb_lo; ++i) { p->b_lo_pos[i] = (uint8_t)fb[i].pos; p->b_lo_sstr[i] = (uint16_t)fb[i].sstr; }
          for (int it = 0; it < 8; ++it) {
            int64_t g = 0; int32_t so = 0;
            for (int j = 8; j < p->na_lo; ++j)
              if ((it >> (j - 8)) & 1) { g |= (int64_t)1 << p->a_lo_pos[j]; so += p->a_lo_sstr[j]; }
            p->tc_a_it_g[it] = g; p->tc_a_it_s[it] = so;
          }
          for (int it = 0; it < 16; ++it) {
            int64_t g = 0; int32_t so = 0;
            for (int j = 8; j < p->nb_lo; ++j)
              if ((it >> (j - 8)) & 1) {
                if (p->b_lo_pos[j] == kPhantom) g = -1; else if (g >= 0) g |= (int64_t)1 << p->b_lo_pos[j];
                so += p->b_lo_sstr[j];
              }
            p->tc_b_it_g[it] = g; p->tc_b_it_s[it] = so;
          }
          p->tc_plane_a = 128 * KT * 2;
          p->tc_plane_b = TN * KT * 2;
          const int rank_h = 3 * TN * KT * 2;
          const int stage_h = 4 * p->tc_plane_a + rank_h;
          p->tc_stages = std::max(2, std::min(8, (227 * 1024 - 512 - raw_ring) / stage_h));
          p->smem_bytes = p->tc_stages * stage_h + raw_ring + 512;
          // + 256 bytes: two magnitude words for callers that come through plain qtn_pair_run
          p->workspace_bytes = base + images * 2 * rank_h + 256;
        }
      }
    }
    // A as pre-packed stage images (see tc_prepack_a in the header): worth one more pass over A when every A tile
    // is read by at least four n tiles
    p->tc_prepack_a = 0;
    p->tc_prepack_a_off = 0;
    if (p->tc_pair == 2 && p->tc_kind == 1 && opt.tc_prepack_a && (p->n_nhi >= 2 || opt.tc_prepack_a >= 2)) {
      const int64_t a_images = (int64_t)1 << (p->n_mhi + p->n_khi);
      const int64_t a_bytes = a_images * 4 * (int64_t)p->tc_plane_a;
      if (a_bytes <= kMaxImageBytes && a_images <= 0x7fffffffLL) {
        p->tc_prepack_a = 1;
        p->tc_prepack_a_off = (p->workspace_bytes + 255) & ~(int64_t)255;
        p->workspace_bytes = p->tc_prepack_a_off + a_bytes + 256;   // the magnitude words stay the last 256 bytes
        // no raw ring for gathered A: the shared memory goes to a deeper operand ring (the halves may drift further
        // apart, which hides more of an accumulator drain)
        const int stage_h = 4 * p->tc_plane_a + 3 * (1 << p->tnb) * (1 << p->tkb) * 2;
        p->tc_stages = std::max(2, std::min(8, (227 * 1024 - 512) / stage_h));
        p->smem_bytes = p->tc_stages * stage_h + 512;
      }
    }
    // address-ordered gather of A for the staggered kernel (see tc_gather in the header)
    p->tc_gather = 0;
    if (p->tc_pair == 2 && opt.tc_gather && !p->tc_prepack_a) {
      const int n = p->na_lo;
      std::vector<int> idx(n);
      for (int j = 0; j < n; ++j) idx[j] = j;
      std::sort(idx.begin(), idx.end(), [&](int x, int y) { return p->a_lo_pos[x] < p->a_lo_pos[y]; });
      for (int r = 0; r < n; ++r) { p->tc_a_gbit[r] = p->a_lo_pos[idx[r]]; p->tc_a_rank[idx[r]] = (uint8_t)r; }
      for (int it = 0; it < 8; ++it) {
        int64_t g = 0; int32_t raw = 0;
        for (int j = 8; j < n; ++j)
          if ((it >> (j - 8)) & 1) { g |= (int64_t)1 << p->tc_a_gbit[j]; raw += 1 << p->tc_a_rank[j]; }
        p->tc_a_it_gg[it] = g; p->tc_a_it_raw[it] = raw;
      }
      p->tc_gather = 1;
      if (opt.tc_gather == 2) {
        // bulk mode: the run of address bits 0, 1, 2, ... owned by the tile.  Measured on the C4 plan
        // (profiles/r02_gather_bulk_ab.jsonl): runs of 8 KB (2 copies per 16 KB chunk) win (m25 n6 k6: 9.0 -> 8.1 ms),
        // 2-4 KB are neutral, 512 B - 1 KB (16-32 copies per chunk) lose up to 40 %: tiles below the threshold
        // ("tc_gather_min_run", default 10 bits) keep the per-thread gather
        int r = 0;
        while (r < n && p->tc_a_gbit[r] == r) ++r;
        if (r >= opt.tc_gather_min_run && n - r <= 5) { p->tc_gather = 2; p->tc_run_bits = r; }
        else p->tc_gather = 0;
      }
    }
    p->sa_stride = 0; p->sb_stride = 0;
    if (bbits > 0 && p->tc_pair != 2)
      return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_pair_plan: batch labels need the staggered CTA-pair kernel");
  } else {
    // ---- shared-memory tiled kernel -----------------------------------------
    p->kernel = QTN_KERNEL_TILE;
    int tmb, tnb;
    if (nbits >= 6 && mbits >= 6) { tmb = 6; tnb = 6; }
    else if (nbits < 6) { tnb = std::max(nbits, 2); tmb = 12 - tnb; }
    else { tmb = std::max(mbits, 2); tnb = 12 - tmb; }   // unreachable after swap, kept for safety
    const int TM = 1 << tmb, TN = 1 << tnb;
    // two stages of (A, B) tiles (the next k step is copied while this one is multiplied) within 72 KB of shared
    // memory, so that three CTAs share an SM: 4224 elements per stage in complex64 (16 k values per step on a
    // 256 x 16 tile were measured slower than 8), 2304 in complex128
    int tkb = 4;
    while (tkb > 0 && (1 << tkb) * (TM + TN + 4) > (d->dtype == QTN_C64 ? 4224 : 2304)) --tkb;
    tkb = std::min(tkb, kbits);
    p->tmb = tmb; p->tnb = tnb; p->tkb = tkb;
    p->sa_stride = TM + 2; p->sb_stride = TN + 2;

    // pick which labels sit inside the tile
    auto pick = [&](std::vector<Lab> &v, int want, auto key, auto forced) {
      std::stable_sort(v.begin(), v.end(), [&](const Lab &x, const Lab &y) {
        bool fx = forced(x), fy = forced(y);
        if (fx != fy) return fx;
        if (key(x) != key(y)) return key(x) < key(y);
        return x.label < y.label;
      });
      int take = std::min<int>(want, (int)v.size());
      std::vector<Lab> lo(v.begin(), v.begin() + take), hi(v.begin() + take, v.end());
      return std::make_pair(lo, hi);
    };
    auto m_sel = pick(M, tmb, [](const Lab &l) { return l.pa; },
                      [&](const Lab &l) { return l.pa < cforce || (l.pc >= 0 && l.pc < cforce); });
    auto n_sel = pick(N, tnb, [](const Lab &l) { return l.pb; },
                      [&](const Lab &l) { return l.pb < cforce || (l.pc >= 0 && l.pc < cforce); });
    auto k_sel = pick(K, tkb, [](const Lab &l) { return std::min(l.pa, l.pb); },
                      [&](const Lab &l) { return l.pa < cforce || l.pb < cforce; });
    std::vector<Lab> &mlo = m_sel.first, &mhi = m_sel.second;
    std::vector<Lab> &nlo = n_sel.first, &nhi = n_sel.second;
    std::vector<Lab> &klo = k_sel.first, &khi = k_sel.second;
    // logical bit order inside the tile = ascending address bit of the operand read
    std::sort(mlo.begin(), mlo.end(), [](const Lab &x, const Lab &y) { return x.pa < y.pa; });
    std::sort(nlo.begin(), nlo.end(), [](const Lab &x, const Lab &y) { return x.pb < y.pb; });
    std::sort(klo.begin(), klo.end(), [](const Lab &x, const Lab &y) { return x.pa < y.pa; });
    std::sort(mhi.begin(), mhi.end(), [](const Lab &x, const Lab &y) { return x.pa < y.pa; });
    std::sort(nhi.begin(), nhi.end(), [](const Lab &x, const Lab &y) { return x.pb < y.pb; });
    std::sort(khi.begin(), khi.end(), [](const Lab &x, const Lab &y) {
      return std::max(x.pa, x.pb) < std::max(y.pa, y.pb);
    });
    std::sort(Bt.begin(), Bt.end(), [](const Lab &x, const Lab &y) { return x.pa < y.pa; });
    p->m_real = (int)mlo.size(); p->n_real = (int)nlo.size(); p->k_real = (int)klo.size();
    p->n_mhi = (int)mhi.size(); p->n_nhi = (int)nhi.size(); p->n_khi = (int)khi.size();
    p->n_batch = bbits;
    if (bbits > QTN_HI_MAX) return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_pair_plan: too many batch labels");

    // C layout when free: tile bits lowest (n then m), then n-hi, m-hi, batch
    if (d->c_layout_free) {
      int next_c = 0;
      for (auto *v : {&nlo, &mlo, &nhi, &mhi, &Bt})
        for (auto &l : *v) { l.pc = next_c++; assigned.push_back({l.label, l.pc}); }
    }

    // A tile enumeration: (m_lo, k_lo) labels sorted by A position; phantoms last
    struct E { int pos; int sstr; };
    std::vector<E> ea, eb;
    for (int i = 0; i < tmb; ++i)
      ea.push_back({i < (int)mlo.size() ? mlo[i].pa : kPhantom, 1 << i});
    for (int i = 0; i < tkb; ++i) ea.push_back({klo[i].pa, p->sa_stride << i});
    for (int i = 0; i < tnb; ++i)
      eb.push_back({i < (int)nlo.size() ? nlo[i].pb : kPhantom, 1 << i});
    for (int i = 0; i < tkb; ++i) eb.push_back({klo[i].pb, p->sb_stride << i});
    auto by_pos = [](const E &x, const E &y) { return x.pos < y.pos; };
    std::stable_sort(ea.begin(), ea.end(), by_pos);
    std::stable_sort(eb.begin(), eb.end(), by_pos);
    p->na_lo = (int)ea.size(); p->nb_lo = (int)eb.size();
    for (int i = 0; i < p->na_lo; ++i) { p->a_lo_pos[i] = (uint8_t)ea[i].pos; p->a_lo_sstr[i] = (uint16_t)ea[i].sstr; }
    for (int i = 0; i < p->nb_lo; ++i) { p->b_lo_pos[i] = (uint8_t)eb[i].pos; p->b_lo_sstr[i] = (uint16_t)eb[i].sstr; }

    // epilogue: real tile labels of C sorted by C position; staging index = rank in that order
    struct CE { int pos; bool is_m; int bit; };
    std::vector<CE> ec;
    for (int i = 0; i < (int)mlo.size(); ++i) ec.push_back({mlo[i].pc, true, i});
    for (int i = 0; i < (int)nlo.size(); ++i) ec.push_back({nlo[i].pc, false, i});
    std::sort(ec.begin(), ec.end(), [](const CE &x, const CE &y) { return x.pos < y.pos; });
    p->nc_lo = (int)ec.size();
    for (int r = 0; r < p->nc_lo; ++r) {
      p->c_lo_pos[r] = (uint8_t)ec[r].pos;
      if (ec[r].is_m) p->c_m_sidx[ec[r].bit] = (uint16_t)(1 << r);
      else p->c_n_sidx[ec[r].bit] = (uint16_t)(1 << r);
    }
    // phantom m/n bits: staging index beyond the real range (never stored)
    {
      int r = p->nc_lo;
      for (int i = (int)mlo.size(); i < tmb; ++i) p->c_m_sidx[i] = (uint16_t)(1 << r++);
      for (int i = (int)nlo.size(); i < tnb; ++i) p->c_n_sidx[i] = (uint16_t)(1 << r++);
    }
    if ((int)mhi.size() > QTN_HI_MAX || (int)nhi.size() > QTN_HI_MAX || (int)khi.size() > QTN_HI_MAX)
      return qtn_fail(QTN_ERR_UNSUPPORTED, "qtn_pair_plan: rank too large");
    for (int i = 0; i < (int)mhi.size(); ++i) { p->a_mhi[i] = mhi[i].pa; p->c_mhi[i] = mhi[i].pc; }
    for (int i = 0; i < (int)nhi.size(); ++i) { p->b_nhi[i] = nhi[i].pb; p->c_nhi[i] = nhi[i].pc; }
    for (int i = 0; i < (int)khi.size(); ++i) { p->a_khi[i] = khi[i].pa; p->b_khi[i] = khi[i].pb; }
    for (int i = 0; i < bbits; ++i) { p->a_bat[i] = Bt[i].pa; p->b_bat[i] = Bt[i].pb; p->c_bat[i] = Bt[i].pc; }

    // split-K when the tile grid cannot fill the machine
    const double tiles = std::ldexp(1.0, p->n_mhi + p->n_nhi + p->n_batch);
    int split = 0;
    while (split < p->n_khi - 1 && tiles * (1 << split) < 2.0 * sm_count) ++split;
    p->split_bits = split;
    int64_t grid_tiles = (int64_t)1 << (p->n_mhi + p->n_nhi + p->n_batch);
    p->grid = grid_tiles << split;
    p->workspace_bytes = split ? ((int64_t)esz << (d->rank_c + split)) : 0;
    int tile_bytes = ((1 << tkb) * (p->sa_stride + p->sb_stride)) * esz;
    int stage_bytes = 4096 * esz;
    p->tile_stages = 2;                                        // copies run one k step ahead, across tiles too
    p->smem_bytes = std::max(p->tile_stages * tile_bytes, stage_bytes);
  }

  // write the chosen layout of C back to the descriptor
  for (const auto &lp : assigned) {
    int jc = find_label(d->label_c, d->rank_c, lp.first);
    d->pos_c[jc] = (uint8_t)lp.second;
  }
  return QTN_OK;
}
