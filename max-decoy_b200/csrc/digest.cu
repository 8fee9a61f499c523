// digest.cu -- K1: in-silico tryptic digestion + integer peptide masses on the GPU.
//
// Replaces FastaDigester::process_file -> DigestEnzym::digest -> Peptide::new
// (utility/input_file_digester/fasta_digester.rs:69-140; models/enzyms/digest_enzym.rs:61-86;
// models/enzyms/trypsin.rs:29; models/peptides/peptide.rs:27-37; peptide_interface.rs:22-28)
// and the UNIQUE(aa_sequence, weight) constraint of table `peptides` (db/schema.sql:41).
//
// Layout: all proteins are one residue buffer in HBM.  Pass 1 flags piece starts (protein start, or
// previous residue in {K,R} and this one != P); the start list is compacted; each piece emits up to
// MC+1 peptides (length filter); every occurrence gets (weight, hash64) over its generalized bytes;
// a stable two-pass radix sort orders occurrences by (weight, hash64, emission order); duplicates
// are removed by exact byte comparison inside equal-key groups; survivors are the peptide table.
#include "cubx.cuh"

namespace {

__global__ void k_mark_protein_starts(const uint64_t* __restrict__ off, uint32_t n_prot, uint8_t* __restrict__ pstart) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_prot && off[i] < off[i + 1]) pstart[off[i]] = 1;
}

// trypsin.rs:29  (?<=[KR])(?!P)
__global__ void k_piece_flags(const uint8_t* __restrict__ res, uint64_t n, const uint8_t* __restrict__ pstart, uint8_t* __restrict__ flag) {
  uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint8_t f = pstart[p];
  if (!f && p > 0) {
    uint8_t a = res[p - 1];
    f = ((a == 'K' || a == 'R') && res[p] != 'P') ? 1 : 0;
  }
  flag[p] = f;
}

__global__ void k_piece_protein(const uint32_t* __restrict__ starts, uint32_t n_pieces, const uint64_t* __restrict__ off, uint32_t n_prot,
                                uint32_t* __restrict__ piece_prot) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pieces) return;
  uint64_t pos = starts[i];
  uint32_t lo = 0, hi = n_prot + 1;  // first index with off[idx] > pos
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (off[mid] <= pos) lo = mid + 1; else hi = mid;
  }
  piece_prot[i] = lo - 1;
}

// digest_enzym.rs:64-86: for piece i, concatenations with 0..MC following pieces of the same protein
template <bool EMIT>
__global__ void k_occurrences(const uint32_t* __restrict__ starts, const uint32_t* __restrict__ piece_prot, uint32_t n_pieces, uint64_t n_res,
                              uint32_t mc_max, uint32_t min_len, uint32_t max_len, uint32_t* __restrict__ count,
                              const uint32_t* __restrict__ occ_off, uint32_t* __restrict__ occ_begin, uint8_t* __restrict__ occ_len,
                              uint8_t* __restrict__ occ_mc, uint32_t* __restrict__ occ_prot) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_pieces) return;
  uint32_t b = starts[i], prot = piece_prot[i], c = 0;
  uint32_t o = EMIT ? occ_off[i] : 0;
  for (uint32_t mc = 0; mc <= mc_max; mc++) {
    uint32_t j = i + mc;
    if (j >= n_pieces || piece_prot[j] != prot) break;
    uint64_t e = (j + 1 < n_pieces) ? (uint64_t)starts[j + 1] : n_res;
    uint32_t len = (uint32_t)(e - b);
    if (len > max_len) break;
    if (len >= min_len) {
      if (EMIT) { occ_begin[o + c] = b; occ_len[o + c] = (uint8_t)len; occ_mc[o + c] = (uint8_t)mc; occ_prot[o + c] = prot; }
      c++;
    }
  }
  if (!EMIT) count[i] = c;
}

__device__ __forceinline__ uint8_t generalize(uint8_t c) { return (c == 'I' || c == 'L') ? (uint8_t)'J' : c; }  // amino_acid.rs:139-141

struct MassTable { int64_t m[MD_NCODES]; };

// Peptide::new (peptide.rs:27-37): weight = H2O + sum (amino_acid.rs:130-136); plus the ordering hash
__global__ void k_occ_keys(const uint8_t* __restrict__ res, const uint32_t* __restrict__ occ_begin, const uint8_t* __restrict__ occ_len, uint32_t n_occ,
                           MassTable T, int64_t* __restrict__ weight, uint64_t* __restrict__ hash) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_occ) return;
  const uint8_t* s = res + occ_begin[i];
  uint32_t len = occ_len[i];
  int64_t w = MD_WATER_UDA;
  uint64_t h = md_hash_init();
  for (uint32_t k = 0; k < len; k++) {
    uint8_t c = generalize(s[k]);
    w += T.m[md_code_of(c)];
    h = md_hash_step(h, c);
  }
  weight[i] = w;
  hash[i] = md_hash_fin(h, len);
}

__global__ void k_iota(uint32_t* v, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] = i;
}

template <class T>
__global__ void k_gather(const T* __restrict__ src, const uint32_t* __restrict__ idx, T* __restrict__ dst, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}

// head[k] = k if sorted element k starts a new (weight, hash) group else 0  (max-scan -> group start)
__global__ void k_group_heads(const int64_t* __restrict__ w_sorted, const uint64_t* __restrict__ hash, const uint32_t* __restrict__ sidx, uint32_t n,
                              uint32_t* __restrict__ head) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  bool h = (k == 0) || w_sorted[k] != w_sorted[k - 1] || hash[sidx[k]] != hash[sidx[k - 1]];
  head[k] = h ? k : 0;
}

__device__ __forceinline__ bool same_sequence(const uint8_t* __restrict__ res, uint32_t ba, uint32_t bb, uint32_t len) {
  for (uint32_t k = 0; k < len; k++)
    if (generalize(res[ba + k]) != generalize(res[bb + k])) return false;
  return true;
}

// rep[k] = first element of k's group with the same generalized bytes (k itself if none): exact dedupe
__global__ void k_find_rep(const uint8_t* __restrict__ res, const uint32_t* __restrict__ sidx, const uint32_t* __restrict__ gstart,
                           const uint32_t* __restrict__ occ_begin, const uint8_t* __restrict__ occ_len, uint32_t n, uint32_t* __restrict__ rep,
                           uint8_t* __restrict__ is_rep) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint32_t me = sidx[k], g = gstart[k], r = k;
  uint32_t len = occ_len[me];
  for (uint32_t j = g; j < k; j++) {
    uint32_t other = sidx[j];
    if (occ_len[other] == len && same_sequence(res, occ_begin[other], occ_begin[me], len)) { r = j; break; }
  }
  rep[k] = r;
  is_rep[k] = (r == k) ? 1 : 0;
}

__global__ void k_u8_to_u32(const uint8_t* in, uint32_t* out, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

// per representative: peptide row scalars; every occurrence contributes to min(mc) and one (peptide, protein) key
__global__ void k_peptide_scalars(const uint32_t* __restrict__ sidx, const uint32_t* __restrict__ rep, const uint32_t* __restrict__ pep_ord,
                                  const uint8_t* __restrict__ is_rep, const uint8_t* __restrict__ occ_len, const uint8_t* __restrict__ occ_mc,
                                  const uint32_t* __restrict__ occ_prot, const int64_t* __restrict__ w_sorted, const uint64_t* __restrict__ hash,
                                  uint32_t n, uint8_t* __restrict__ pep_len, uint32_t* __restrict__ pep_len32, int64_t* __restrict__ pep_weight,
                                  uint64_t* __restrict__ pep_hash, uint32_t* __restrict__ pep_mc32, uint32_t* __restrict__ pep_first,
                                  uint64_t* __restrict__ assoc_key) {
  uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  uint32_t me = sidx[k];
  uint32_t p = pep_ord[rep[k]];
  if (is_rep[k]) {
    pep_len[p] = occ_len[me]; pep_len32[p] = occ_len[me];
    pep_weight[p] = w_sorted[k]; pep_hash[p] = hash[me]; pep_first[p] = me;
  }
  atomicMin(&pep_mc32[p], (uint32_t)occ_mc[me]);
  assoc_key[k] = ((uint64_t)p << 32) | occ_prot[me];
}

// peptide bytes (generalized) and the 21 counts (peptide_interface.rs:22-28)
__global__ void k_peptide_bytes(const uint8_t* __restrict__ res, const uint32_t* __restrict__ occ_begin, const uint32_t* __restrict__ pep_first,
                                const uint32_t* __restrict__ seq_off, const uint8_t* __restrict__ pep_len, const uint32_t* __restrict__ pep_mc32,
                                uint32_t n_pep, uint8_t* __restrict__ seq, int16_t* __restrict__ counts, uint8_t* __restrict__ pep_mc) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pep) return;
  const uint8_t* s = res + occ_begin[pep_first[p]];
  uint8_t* d = seq + seq_off[p];
  uint32_t len = pep_len[p];
  int16_t c[MD_ALPHABET_SIZE];
#pragma unroll
  for (int a = 0; a < MD_ALPHABET_SIZE; a++) c[a] = 0;
  for (uint32_t k = 0; k < len; k++) {
    uint8_t ch = generalize(s[k]);
    d[k] = ch;
    int a = md_alpha_of_code(md_code_of(ch));
#pragma unroll
    for (int q = 0; q < MD_ALPHABET_SIZE; q++) c[q] += (q == a);
  }
#pragma unroll
  for (int a = 0; a < MD_ALPHABET_SIZE; a++) counts[(size_t)p * MD_ALPHABET_SIZE + a] = c[a];
  pep_mc[p] = (uint8_t)pep_mc32[p];
}

__global__ void k_assoc_offsets(const uint64_t* __restrict__ keys, uint32_t n_keys, uint32_t n_pep, uint32_t* __restrict__ assoc_off) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p > n_pep) return;
  uint64_t target = (uint64_t)p << 32;
  uint32_t lo = 0, hi = n_keys;
  while (lo < hi) {
    uint32_t mid = (lo + hi) >> 1;
    if (keys[mid] < target) lo = mid + 1; else hi = mid;
  }
  assoc_off[p] = lo;
}

__global__ void k_low32(const uint64_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)keys[i];
}

// membership table of Decoy::is_peptide (decoy.rs:49-60)
__global__ void k_ht_insert(const uint64_t* __restrict__ pep_hash, uint32_t n_pep, unsigned long long* __restrict__ ht_key, uint32_t* __restrict__ ht_val,
                            uint32_t mask) {
  uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_pep) return;
  unsigned long long h = pep_hash[p];
  if (h == 0) h = 1;
  uint32_t slot = (uint32_t)h & mask;
  for (;;) {
    unsigned long long prev = atomicCAS(&ht_key[slot], 0ULL, h);
    if (prev == 0ULL) { ht_val[slot] = p; return; }
    slot = (slot + 1) & mask;
  }
}

inline uint32_t blocks(uint64_t n, uint32_t bs = 256) { return (uint32_t)((n + bs - 1) / bs); }

}  // namespace

void digest_run(md_ctx* ctx, const uint8_t* residues, const uint64_t* off, uint32_t n_prot, const md_digest_params& p) {
  MD_REQUIRE(p.max_len <= MD_MAX_PEPTIDE_LEN, MD_ERR_INVALID, "md_digest: max_len > 60 (tasks/digestion.rs:101)");
  MD_REQUIRE(p.max_missed_cleavages <= 60, MD_ERR_INVALID, "md_digest: max_missed_cleavages > 60 (tasks/digestion.rs:74)");
  MD_REQUIRE(p.min_len >= 1 && p.min_len <= p.max_len, MD_ERR_INVALID, "md_digest: need 1 <= min_len <= max_len");
  for (uint32_t i = 0; i < n_prot; i++) MD_REQUIRE(off[i + 1] >= off[i], MD_ERR_INVALID, "md_digest: protein_offsets not monotone");
  const uint64_t N = n_prot ? off[n_prot] - off[0] : 0;
  MD_REQUIRE(N < 0xFFFF0000ull, MD_ERR_UNSUPPORTED, "md_digest: more than 4 Gi residues per call");
  PeptideStore& P = ctx->peps;
  P.ready = false; ctx->index.ready = false;
  P.n = 0; P.seq_bytes = 0; P.n_assoc = 0;
  cudaStream_t st = ctx->stream;
  if (N == 0) {  // nothing to digest: an empty, valid table
    P.seq_off.need(1); MD_CUDA(cudaMemsetAsync(P.seq_off.p, 0, sizeof(uint32_t), st));
    P.assoc_off.need(1); MD_CUDA(cudaMemsetAsync(P.assoc_off.p, 0, sizeof(uint32_t), st));
    P.ht_key.need(2); P.ht_val.need(2); P.ht_mask = 1; MD_CUDA(cudaMemsetAsync(P.ht_key.p, 0, 2 * sizeof(uint64_t), st));
    MD_CUDA(cudaStreamSynchronize(st));
    P.ready = true;
    return;
  }
  // normalise offsets to start at 0
  std::vector<uint64_t> off0(n_prot + 1);
  for (uint32_t i = 0; i <= n_prot; i++) off0[i] = off[i] - off[0];

  DevBuf<uint8_t> d_res, d_pstart, d_flag;
  DevBuf<uint64_t> d_off;
  DevBuf<uint32_t> d_scalar;  // [0] generic count
  d_res.need(N); d_pstart.need(N); d_flag.need(N); d_off.need(n_prot + 1); d_scalar.need(4);
  MD_CUDA(cudaMemcpyAsync(d_res.p, residues + off[0], N, cudaMemcpyHostToDevice, st));
  MD_CUDA(cudaMemcpyAsync(d_off.p, off0.data(), (n_prot + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  MD_CUDA(cudaMemsetAsync(d_pstart.p, 0, N, st));
  MD_LAUNCH(ctx, k_mark_protein_starts, blocks(n_prot), 256, 0, d_off.p, n_prot, d_pstart.p);
  MD_LAUNCH(ctx, k_piece_flags, blocks(N), 256, 0, d_res.p, N, d_pstart.p, d_flag.p);

  DevBuf<uint32_t> d_starts;
  d_starts.need(N);
  cubx_select_flagged_index(ctx, d_flag.p, d_starts.p, d_scalar.p, N);
  const uint32_t n_pieces = d2h_scalar(ctx, d_scalar.p);
  d_pstart.release(); d_flag.release();

  DevBuf<uint32_t> d_piece_prot, d_cnt, d_occ_off;
  d_piece_prot.need(n_pieces); d_cnt.need(n_pieces + 1); d_occ_off.need(n_pieces + 1);
  MD_LAUNCH(ctx, k_piece_protein, blocks(n_pieces), 256, 0, d_starts.p, n_pieces, d_off.p, n_prot, d_piece_prot.p);
  MD_CUDA(cudaMemsetAsync(d_cnt.p, 0, (n_pieces + 1) * sizeof(uint32_t), st));
  MD_LAUNCH(ctx, k_occurrences<false>, blocks(n_pieces), 256, 0, d_starts.p, d_piece_prot.p, n_pieces, N, p.max_missed_cleavages, p.min_len, p.max_len,
            d_cnt.p, nullptr, nullptr, nullptr, nullptr, nullptr);
  cubx_exclusive_sum(ctx, d_cnt.p, d_occ_off.p, n_pieces + 1);
  const uint32_t n_occ = d2h_scalar(ctx, d_occ_off.p + n_pieces);
  MD_REQUIRE(n_occ < 0x7FFFFFFFu, MD_ERR_UNSUPPORTED, "md_digest: too many peptide occurrences for one call");

  if (n_occ == 0) {
    P.seq_off.need(1); MD_CUDA(cudaMemsetAsync(P.seq_off.p, 0, sizeof(uint32_t), st));
    P.assoc_off.need(1); MD_CUDA(cudaMemsetAsync(P.assoc_off.p, 0, sizeof(uint32_t), st));
    P.ht_key.need(2); P.ht_val.need(2); P.ht_mask = 1; MD_CUDA(cudaMemsetAsync(P.ht_key.p, 0, 2 * sizeof(uint64_t), st));
    MD_CUDA(cudaStreamSynchronize(st));
    P.ready = true;
    return;
  }

  DevBuf<uint32_t> d_occ_begin, d_occ_prot;
  DevBuf<uint8_t> d_occ_len, d_occ_mc;
  d_occ_begin.need(n_occ); d_occ_prot.need(n_occ); d_occ_len.need(n_occ); d_occ_mc.need(n_occ);
  MD_LAUNCH(ctx, k_occurrences<true>, blocks(n_pieces), 256, 0, d_starts.p, d_piece_prot.p, n_pieces, N, p.max_missed_cleavages, p.min_len, p.max_len,
            nullptr, d_occ_off.p, d_occ_begin.p, d_occ_len.p, d_occ_mc.p, d_occ_prot.p);
  d_cnt.release(); d_piece_prot.release(); d_starts.release(); d_occ_off.release();

  MassTable T;
  for (int c = 0; c < MD_NCODES; c++) T.m[c] = kResidueMassByCode[c];
  DevBuf<int64_t> d_w, d_w2;
  DevBuf<uint64_t> d_h, d_h2;
  DevBuf<uint32_t> d_i0, d_i1;
  d_w.need(n_occ); d_w2.need(n_occ); d_h.need(n_occ); d_h2.need(n_occ); d_i0.need(n_occ); d_i1.need(n_occ);
  MD_LAUNCH(ctx, k_occ_keys, blocks(n_occ), 256, 0, d_res.p, d_occ_begin.p, d_occ_len.p, n_occ, T, d_w.p, d_h.p);
  MD_LAUNCH(ctx, k_iota, blocks(n_occ), 256, 0, d_i0.p, n_occ);
  // stable sort by hash, then stable sort by weight  ->  (weight, hash, emission order)
  cubx_sort_pairs(ctx, d_h.p, d_h2.p, d_i0.p, d_i1.p, n_occ);
  MD_LAUNCH(ctx, k_gather<int64_t>, blocks(n_occ), 256, 0, d_w.p, d_i1.p, d_w2.p, n_occ);
  DevBuf<int64_t> d_ws;  // weights in final sorted order
  d_ws.need(n_occ);
  cubx_sort_pairs(ctx, d_w2.p, d_ws.p, d_i1.p, d_i0.p, n_occ);  // d_i0 = sorted occurrence ids
  d_w2.release(); d_h2.release(); d_i1.release();
  uint32_t* sidx = d_i0.p;

  DevBuf<uint32_t> d_head, d_gstart, d_rep, d_ord, d_rep32;
  DevBuf<uint8_t> d_isrep;
  d_head.need(n_occ); d_gstart.need(n_occ); d_rep.need(n_occ); d_ord.need(n_occ + 1); d_rep32.need(n_occ + 1); d_isrep.need(n_occ);
  MD_LAUNCH(ctx, k_group_heads, blocks(n_occ), 256, 0, d_ws.p, d_h.p, sidx, n_occ, d_head.p);
  cubx_inclusive_max_u32(ctx, d_head.p, d_gstart.p, n_occ);
  MD_LAUNCH(ctx, k_find_rep, blocks(n_occ), 256, 0, d_res.p, sidx, d_gstart.p, d_occ_begin.p, d_occ_len.p, n_occ, d_rep.p, d_isrep.p);
  MD_CUDA(cudaMemsetAsync(d_rep32.p, 0, (n_occ + 1) * sizeof(uint32_t), st));
  MD_LAUNCH(ctx, k_u8_to_u32, blocks(n_occ), 256, 0, d_isrep.p, d_rep32.p, n_occ);
  cubx_exclusive_sum(ctx, d_rep32.p, d_ord.p, n_occ + 1);
  const uint32_t n_pep = d2h_scalar(ctx, d_ord.p + n_occ);
  d_head.release(); d_gstart.release(); d_rep32.release();

  P.len.need(n_pep); P.mc.need(n_pep); P.weight.need(n_pep); P.hash.need(n_pep); P.seq_off.need(n_pep + 1);
  P.counts.need((size_t)n_pep * MD_ALPHABET_SIZE);
  DevBuf<uint32_t> d_len32, d_mc32, d_first;
  DevBuf<uint64_t> d_akey, d_akey2;
  d_len32.need(n_pep + 1); d_mc32.need(n_pep); d_first.need(n_pep); d_akey.need(n_occ); d_akey2.need(n_occ);
  MD_CUDA(cudaMemsetAsync(d_len32.p, 0, (n_pep + 1) * sizeof(uint32_t), st));
  MD_CUDA(cudaMemsetAsync(d_mc32.p, 0xFF, n_pep * sizeof(uint32_t), st));
  MD_LAUNCH(ctx, k_peptide_scalars, blocks(n_occ), 256, 0, sidx, d_rep.p, d_ord.p, d_isrep.p, d_occ_len.p, d_occ_mc.p, d_occ_prot.p, d_ws.p, d_h.p, n_occ,
            P.len.p, d_len32.p, P.weight.p, P.hash.p, d_mc32.p, d_first.p, d_akey.p);
  cubx_exclusive_sum(ctx, d_len32.p, P.seq_off.p, n_pep + 1);
  const uint32_t seq_bytes = d2h_scalar(ctx, P.seq_off.p + n_pep);
  P.seq.need(seq_bytes + 16);
  MD_LAUNCH(ctx, k_peptide_bytes, blocks(n_pep), 256, 0, d_res.p, d_occ_begin.p, d_first.p, P.seq_off.p, P.len.p, d_mc32.p, n_pep, P.seq.p, P.counts.p, P.mc.p);

  // peptides_proteins links: unique (peptide, protein) pairs, CSR by peptide
  cubx_sort_keys(ctx, d_akey.p, d_akey2.p, n_occ);
  cubx_unique_u64(ctx, d_akey2.p, d_akey.p, d_scalar.p, n_occ);
  const uint32_t n_assoc = d2h_scalar(ctx, d_scalar.p);
  P.assoc_off.need(n_pep + 1); P.assoc_protein.need(n_assoc);
  MD_LAUNCH(ctx, k_assoc_offsets, blocks(n_pep + 1), 256, 0, d_akey.p, n_assoc, n_pep, P.assoc_off.p);
  MD_LAUNCH(ctx, k_low32, blocks(n_assoc), 256, 0, d_akey.p, n_assoc, P.assoc_protein.p);

  // membership table
  uint32_t cap = 16;
  while (cap < 2u * n_pep + 2u) cap <<= 1;
  P.ht_key.need(cap); P.ht_val.need(cap); P.ht_mask = cap - 1;
  MD_CUDA(cudaMemsetAsync(P.ht_key.p, 0, (size_t)cap * sizeof(uint64_t), st));
  MD_LAUNCH(ctx, k_ht_insert, blocks(n_pep), 256, 0, P.hash.p, n_pep, (unsigned long long*)P.ht_key.p, P.ht_val.p, P.ht_mask);

  MD_CUDA(cudaStreamSynchronize(st));
  P.n = n_pep; P.seq_bytes = seq_bytes; P.n_assoc = n_assoc;
  P.ready = true;
}

namespace {
template <class T>
T* host_copy(md_ctx* ctx, const T* dev, size_t n) {
  T* h = (T*)malloc((n ? n : 1) * sizeof(T));
  MD_REQUIRE(h != nullptr, MD_ERR_NOMEM, "out of host memory");
  if (n) MD_CUDA(cudaMemcpyAsync(h, dev, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  return h;
}
}  // namespace

void digest_export(md_ctx* ctx, md_peptide_table* out) {
  PeptideStore& P = ctx->peps;
  MD_REQUIRE(P.ready, MD_ERR_STATE, "md_peptides_export: no digest yet");
  memset(out, 0, sizeof(*out));
  out->n = P.n; out->seq_bytes = P.seq_bytes; out->n_assoc = P.n_assoc;
  out->seq = host_copy(ctx, P.seq.p, P.seq_bytes);
  out->missed_cleavages = host_copy(ctx, P.mc.p, P.n);
  out->weight = host_copy(ctx, P.weight.p, P.n);
  out->counts = host_copy(ctx, P.counts.p, P.n * MD_ALPHABET_SIZE);
  out->assoc_protein = host_copy(ctx, P.assoc_protein.p, P.n_assoc);
  uint32_t* so = host_copy(ctx, P.seq_off.p, P.n + 1);
  uint32_t* ao = host_copy(ctx, P.assoc_off.p, P.n + 1);
  MD_CUDA(cudaStreamSynchronize(ctx->stream));
  out->seq_off = (uint64_t*)malloc((P.n + 1) * sizeof(uint64_t));
  out->assoc_off = (uint64_t*)malloc((P.n + 1) * sizeof(uint64_t));
  for (uint64_t i = 0; i <= P.n; i++) { out->seq_off[i] = so[i]; out->assoc_off[i] = ao[i]; }
  free(so); free(ao);
}
