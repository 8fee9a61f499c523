// decoy.cu -- K3: mass-constrained decoy generation on the GPU.
//
// REFERENCE_RANDOM restates DecoyGenerator::generate_decoys' worker loop (utility/decoy_generator.rs:127-188)
// and ModifiedPeptide::swap_amino_acids_to_hit_mass_tolerance (models/peptides/modified_peptide.rs:451-508):
// grow a random sequence until its weight exceeds the upper limit, then repair it by greedy single-residue
// substitutions (<= 100 passes, one random kick per pass), testing the window / variable modifications after
// every applied substitution.  One GPU lane runs one attempt at a time and fetches the next attempt from a
// global queue the moment it finishes, so lanes never idle behind a slow neighbour.  Attempts are keyed by
// (seed, spectrum_id, attempt) through a Philox counter RNG, so the decoys of a spectrum are the first n distinct
// non-peptide successes in attempt order, independent of scheduling, batching and sharding.
// PERMUTE_TARGET restates DecoyGenerator::vary_targets (decoy_generator.rs:265-296).
#include "cubx.cuh"
#include "modpep.cuh"
#include "decoyutil.cuh"

#include <algorithm>

namespace {

inline uint32_t blocks(uint64_t n, uint32_t bs = 256) { return (uint32_t)((n + bs - 1) / bs); }

constexpr int kThreads = 128;          // lanes per CTA of the attempt kernels
constexpr int kMaxRoundAttempts = 4096;  // per spectrum per round (bounds the selection kernel's shared memory)

// letter tables in alphabet-index space, built once per call on the host
struct DecoyTables {
  int64_t mprime[32];      // residue mass + fixed delta, by alphabet index (0..20)
  int64_t sorted_m[32];    // mprime sorted ascending (ties by alphabet index), padded with INT64_MAX
  uint8_t sorted_a[32];    // alphabet index at each sorted position
  uint8_t run_min_a[32];   // lowest alphabet index among entries of equal mass
  uint8_t code_of_a[32];   // residue code of alphabet index
  uint8_t has_var_a[32];
  uint8_t has_fix_a[32];
  int64_t var_a[32];
  // improving-substitution filter: letter a has a substitution that strictly reduces |d| iff its gap to the next lighter
  // (d > 0) / heavier (d < 0) distinct mass is < 2|d|.  Gaps sorted ascending + the letter set of every prefix.
  uint32_t gapb_sorted[32], gapa_sorted[32];   // padded with 0xFFFFFFFF
  uint32_t maskb_prefix[33], maska_prefix[33];
};

struct TSeq {  // a lane's working sequence in shared memory (alphabet indices), transposed for conflict-free access
  uint8_t* base;
  __device__ __forceinline__ uint8_t& at(uint32_t i) const { return base[i * kThreads]; }
};
struct TSeqCode {  // view as residue codes for md_try_variable
  TSeq s; const uint8_t* code_of_a;
  __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return code_of_a[s.at(i)]; }
};

struct AttemptOut {
  uint8_t* rows; uint8_t* len; uint64_t* mask; int64_t* w; uint64_t* hash;
};

// write one finished attempt (len == 0 -> failure)
__device__ void store_attempt(const AttemptOut& O, uint64_t slot, const TSeq& seq, uint32_t L, uint64_t mask, int64_t w, const DecoyTables& T,
                              const PeptideView& PV) {
  if (L == 0) { O.len[slot] = 0; return; }
  uint8_t ascii[MD_MAX_PEPTIDE_LEN];
  uint64_t h = md_hash_init();
  uint8_t* row = O.rows + slot * MD_DECOY_ROW;
  for (uint32_t i = 0; i < L; i++) {
    uint8_t code = T.code_of_a[seq.at(i)];
    uint8_t ch = md_letter_of(code);
    ascii[i] = ch; row[i] = code;
    h = md_hash_step(h, ch);
  }
  for (uint32_t i = L; i < ((L + 15u) & ~15u); i++) row[i] = MD_CODE_OTHER;   // pad the last 16-byte chunk
  { const uint32_t pad = MD_CODE_OTHER * 0x01010101u;
    for (uint32_t c = (L + 15u) >> 4; c < MD_DECOY_ROW / 16; c++) reinterpret_cast<uint4*>(row)[c] = make_uint4(pad, pad, pad, pad); }
  h = md_hash_fin(h, L);
  if (is_peptide(ascii, L, h, PV.ht_key, PV.ht_val, PV.ht_mask, PV.seq, PV.off, PV.len)) { O.len[slot] = 0; return; }
  O.len[slot] = (uint8_t)L; O.mask[slot] = mask; O.w[slot] = w; O.hash[slot] = h;
}

// work item -> (list entry, attempt ordinal) by binary search on the round's prefix of attempt counts
__device__ __forceinline__ uint32_t find_entry(const uint32_t* __restrict__ att_off, uint32_t n_list, uint32_t w) {
  uint32_t lo = 0, hi = n_list;
  while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (att_off[mid] <= w) lo = mid; else hi = mid; }
  return lo;
}

// Philox4x32-10 output stream of one attempt, buffered in a per-lane shared-memory ring so that the block function runs
// for the whole warp at once (every kRngPeriod steps of the kernel loop) instead of as a divergent tail behind whichever
// lane happens to run dry.  The stream is exactly Philox4::next()'s: blocks c0 = 0, 1, 2, ... in order, four words each.
struct RngRing {
  uint32_t* buf;                 // 8 words, stride kThreads
  uint32_t k0, k1, c0, c1, c2;   // key, next block, attempt, spectrum
  uint32_t head, count;
  __device__ __forceinline__ void start(uint64_t seed, uint32_t spectrum_id, uint32_t attempt) {
    k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32); c0 = 0; c1 = attempt; c2 = spectrum_id; head = 0; count = 0;
  }
  __device__ __forceinline__ void produce() {   // one block -> 4 words into the ring (needs count <= 4)
    uint32_t a = c0, b = c1, c = c2, d = MD_TAG_RANDOM, x = k0, y = k1;
#pragma unroll
    for (int r = 0; r < 10; r++) {
      const uint64_t p0 = (uint64_t)0xD2511F53u * a, p1 = (uint64_t)0xCD9E8D57u * c;
      const uint32_t n0 = (uint32_t)(p1 >> 32) ^ b ^ x, n1 = (uint32_t)p1;
      const uint32_t n2 = (uint32_t)(p0 >> 32) ^ d ^ y, n3 = (uint32_t)p0;
      a = n0; b = n1; c = n2; d = n3;
      x += 0x9E3779B9u; y += 0xBB67AE85u;
    }
    const uint32_t t = head + count;
    buf[((t + 0) & 7) * kThreads] = a; buf[((t + 1) & 7) * kThreads] = b; buf[((t + 2) & 7) * kThreads] = c; buf[((t + 3) & 7) * kThreads] = d;
    c0++; count += 4;
  }
  __device__ __forceinline__ uint32_t next() {
    if (count == 0) produce();
    const uint32_t v = buf[(head & 7) * kThreads];
    head++; count--;
    return v;
  }
  __device__ __forceinline__ uint32_t below(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 32); }
};
constexpr uint32_t kRngPeriod = 6;

// One lane = one attempt at a time, run as a flat state machine: every pass of the kernel loop does ONE greedy step for
// every lane that has an attempt (kSpec positions evaluated side by side against the current residual: independent
// search chains; the first improving substitution is applied and the positions behind it are looked at again in the
// next step), so the lanes of a warp stay in the same code although their attempts are at different tries / positions
// / lengths (nested try/position loops would make 31 finished lanes wait for the one that runs all 100 tries).  Lanes
// whose attempt ends (hit, or 100 tries without one) park until a quarter of the warp is free, then fetch and grow new
// attempts together.  All residual arithmetic is 32-bit: |w - P| stays far below 2^30 uDa after the grow phase
// (checked; reported via `overflow`).
__global__ void __launch_bounds__(kThreads) k_decoy_random(const md_precursor* __restrict__ prec, const uint32_t* __restrict__ list,
                                                           const uint32_t* __restrict__ att_off, const uint32_t* __restrict__ att_base, uint32_t n_list,
                                                           uint32_t total, uint32_t* __restrict__ queue, uint64_t seed, const __grid_constant__ ModTables M,
                                                           const __grid_constant__ DecoyTables T, AttemptOut O, PeptideView PV, int* __restrict__ overflow) {
  __shared__ uint8_t sseq[MD_MAX_PEPTIDE_LEN * kThreads];
  __shared__ uint32_t s_rng[8 * kThreads];
  __shared__ uint64_t s_pm[MD_ALPHABET_SIZE * kThreads];   // per lane and letter: the positions holding that letter
  __shared__ int32_t s_sorted[32];      // (mass + fixed delta) ascending, padded with INT32_MAX
  __shared__ int32_t s_mprime[32];      // by alphabet index
  __shared__ int32_t s_var[32];
  __shared__ uint8_t s_runmin[32];
  __shared__ uint32_t s_gapb[32], s_gapa[32], s_maskb[33], s_maska[33];
  if (threadIdx.x < 33) { s_maskb[threadIdx.x] = T.maskb_prefix[threadIdx.x]; s_maska[threadIdx.x] = T.maska_prefix[threadIdx.x]; }
  if (threadIdx.x < 32) {
    s_gapb[threadIdx.x] = T.gapb_sorted[threadIdx.x]; s_gapa[threadIdx.x] = T.gapa_sorted[threadIdx.x];
    const int64_t sm = T.sorted_m[threadIdx.x];
    s_sorted[threadIdx.x] = sm > 0x3FFFFFFF ? INT32_MAX : (int32_t)sm;
    s_mprime[threadIdx.x] = (int32_t)T.mprime[threadIdx.x]; s_var[threadIdx.x] = (int32_t)T.var_a[threadIdx.x];
    s_runmin[threadIdx.x] = T.run_min_a[threadIdx.x];
  }
  __syncthreads();
  TSeq seq{sseq + threadIdx.x};
  const bool any_var = M.nvar > 0;
  // one variable letter without a fixed modification (e.g. Met oxidation): its positions are tracked in `vpos`, and
  // try_variable_modifications needs no walk over the sequence
  const int va = M.var_simple_code >= 0 ? md_alpha_of_code((uint32_t)M.var_simple_code) : -1;

  bool busy = false, drained = false, fm_stale = true;
  uint32_t pending = 0;           // 1 = hit, 2 = gave up: the result is written when the free lanes refill together
  uint32_t fm = 0, present = 0;   // letters with an improving substitution at the current d / letters in the sequence
  uint64_t* pm = s_pm + threadIdx.x;
  auto add_letter = [&](uint32_t a, uint32_t i) { pm[a * kThreads] |= 1ULL << i; present |= 1u << a; };
  auto del_letter = [&](uint32_t a, uint32_t i) { const uint64_t v = pm[a * kThreads] & ~(1ULL << i); pm[a * kThreads] = v; if (v == 0) present &= ~(1u << a); };
  uint32_t wi = 0, L = 0, pos = 0, tries = 0;
  int64_t w = 0, P = 0, lo = 0, hi = 0;
  int32_t d = 0;
  uint64_t mask = 0, vpos = 0;
  RngRing rng; rng.buf = s_rng + threadIdx.x; rng.start(seed, 0, 0);

  for (uint32_t it = 0;; it++) {
    // ---- refill: when at least 8 lanes are free (or nobody works), they fetch and grow new attempts together
    const uint32_t free_m = __ballot_sync(0xffffffffu, !busy && (!drained || pending));
    const uint32_t busy_m = __ballot_sync(0xffffffffu, busy);
    if (free_m && (__popc(free_m) >= 8 || busy_m == 0)) {
      if (!busy && pending) { store_attempt(O, wi, seq, pending == 1 ? L : 0, mask, w, T, PV); pending = 0; }
      if (!busy && !drained) {
        wi = atomicAdd(queue, 1u);
        if (wi >= total) drained = true;
        else {
          const uint32_t li = find_entry(att_off, n_list, wi);
          const md_precursor pr = prec[list[li]];
          P = pr.mass; lo = pr.lo; hi = pr.hi;
          rng.start(seed, pr.spectrum_id, att_base[li] + (wi - att_off[li]));
          // grow (decoy_generator.rs:142-159): uniform letters until the weight exceeds the upper limit
          w = MD_WATER_UDA; L = 0; mask = 0; vpos = 0; present = 0; bool dead = false;
          for (int a = 0; a < MD_ALPHABET_SIZE; a++) pm[a * kThreads] = 0;
          for (;;) {
            const uint32_t a = rng.below(MD_ALPHABET_SIZE);
            if (L >= MD_MAX_PEPTIDE_LEN) { dead = true; break; }  // > 60 residues: VARCHAR(60) would reject it
            if ((int)a == va) vpos |= 1ULL << L;
            add_letter(a, L); seq.at(L++) = (uint8_t)a;
            w += s_mprime[a];
            if (w > hi) break;
          }
          const int64_t dd = w - P;
          if (!dead && (dd > 0x3FFFFFFF || dd < -0x3FFFFFFF)) { *overflow = 2; dead = true; }
          if (dead) store_attempt(O, wi, seq, 0, 0, 0, T, PV);
          else { busy = true; tries = 0; pos = 0; d = (int32_t)dd; fm_stale = true; }
        }
      }
    }
    if (__ballot_sync(0xffffffffu, busy) == 0) {
      if (__ballot_sync(0xffffffffu, !drained || pending) == 0) break;
      continue;
    }
    // ---- keep the random-number rings topped up, all lanes together
    if (it % kRngPeriod == 0) { if (busy && rng.count <= 4) rng.produce(); }
    // ---- one greedy step (modified_peptide.rs:454-487) of every busy lane.  Which LETTERS have a substitution that
    //      strictly reduces |d| follows from d alone (`fm`, see DecoyTables); per-letter position masks then give the
    //      first position of the pass that can improve, and only there the full nearest-mass search runs.
    if (busy) {
      if (fm_stale) {
        const uint32_t x = 2u * (uint32_t)(d < 0 ? -d : d);
        const uint32_t* g = d > 0 ? s_gapb : s_gapa;
        uint32_t k = 0;
        if (g[k + 15] < x) k += 16;
        if (g[k + 7] < x) k += 8;
        if (g[k + 3] < x) k += 4;
        if (g[k + 1] < x) k += 2;
        if (g[k] < x) k += 1;
        fm = d == 0 ? 0u : (d > 0 ? s_maskb[k] : s_maska[k]);
        fm_stale = false;
      }
      // first position at or behind `pos` whose letter can improve (none: the rest of the pass is a no-op): OR of the
      // position masks of the qualifying letters -- or, when most letters qualify (right after a kick), the complement
      // of the OR over the few that do not (every position holds exactly one of the present letters).
      uint32_t found = L;
      {
        const uint32_t m0 = fm & present, n0 = present & ~fm;
        const bool inv = __popc(n0) < __popc(m0);
        uint64_t acc = 0;
        for (uint32_t m = inv ? n0 : m0; m; m &= m - 1) acc |= pm[(__ffs(m) - 1) * kThreads];
        uint64_t cand = inv ? ~acc & ((1ULL << L) - 1ULL) : acc;      // L <= 60
        cand &= ~0ULL << pos;                                          // pos < L
        if (cand) found = (uint32_t)__ffsll((long long)cand) - 1;
      }
      bool hit = false;
      if (found < L) {
        pos = found;
        const uint32_t cur = seq.at(pos);
        const int32_t best = d < 0 ? -d : d;
        // best single substitution = letter whose (mass+fixed) is closest to mprime[cur] - d; strict improvement,
        // ties by alphabet order (the reference follows HashMap order there)
        const int32_t target = s_mprime[cur] - d;
        uint32_t k = 0;
        if (s_sorted[k + 15] < target) k += 16;
        if (s_sorted[k + 7] < target) k += 8;
        if (s_sorted[k + 3] < target) k += 4;
        if (s_sorted[k + 1] < target) k += 2;
        if (s_sorted[k] < target) k += 1;
        // candidates: sorted[k-1] (< target) and sorted[k] (>= target)
        const int64_t dist_hi = (k < MD_ALPHABET_SIZE) ? (int64_t)s_sorted[k] - target : INT64_MAX;
        const int64_t dist_lo = (k > 0) ? (int64_t)target - s_sorted[k - 1] : INT64_MAX;
        const uint32_t a_hi = (k < MD_ALPHABET_SIZE) ? s_runmin[k] : 255u;
        const uint32_t a_lo = (k > 0) ? s_runmin[k - 1] : 255u;
        int64_t cd; uint32_t a0;
        if (dist_lo < dist_hi || (dist_lo == dist_hi && a_lo < a_hi)) { cd = dist_lo; a0 = a_lo; } else { cd = dist_hi; a0 = a_hi; }
        if (cd < (int64_t)best && a0 != cur) {
          // remove_modification_at + swap + fixed mod of the new letter (:470-482)
          if ((mask >> pos) & 1) { w -= s_var[cur]; mask &= ~(1ULL << pos); }
          w += s_mprime[a0] - s_mprime[cur];
          seq.at(pos) = (uint8_t)a0; del_letter(cur, pos); add_letter(a0, pos);
          vpos = (vpos & ~(1ULL << pos)) | ((int)a0 == va ? 1ULL << pos : 0ULL);
          if (md_in_window(w, lo, hi)) hit = true;
          else if (any_var) {
            if (va >= 0) {
              if (vpos) hit = md_try_variable_simple(M, vpos, w - (int64_t)__popcll(mask) * M.var[M.var_simple_code], w, mask, lo, hi);
            } else {
              TSeqCode sc{seq, T.code_of_a};
              if (md_try_variable(M, sc, L, w, mask, lo, hi, overflow)) hit = true;
            }
          }
          const int64_t dd = w - P;
          if (dd > 0x3FFFFFFF || dd < -0x3FFFFFFF) *overflow = 2;
          d = (int32_t)dd; fm_stale = true;
        } else {
          *overflow = 3;   // the letter filter and the search disagree: cannot happen (reported as an internal error)
        }
        pos += 1;
      } else {
        pos = L;
      }
      if (hit) { pending = 1; busy = false; }
      else if (pos >= L) {
        // random kick (:489-505)
        const uint32_t i = rng.below(L);
        const uint32_t c = rng.below(MD_ALPHABET_SIZE);
        const uint32_t old = seq.at(i);
        if ((mask >> i) & 1) { w -= s_var[old]; mask &= ~(1ULL << i); }
        w += s_mprime[c] - s_mprime[old];
        seq.at(i) = (uint8_t)c; del_letter(old, i); add_letter(c, i);
        vpos = (vpos & ~(1ULL << i)) | ((int)c == va ? 1ULL << i : 0ULL);
        const int64_t dd = w - P;
        if (dd > 0x3FFFFFFF || dd < -0x3FFFFFFF) *overflow = 2;
        d = (int32_t)dd; fm_stale = true;
        pos = 0;
        if (++tries == 100) { pending = 2; busy = false; }
      }
    }
  }
}

// vary_targets (decoy_generator.rs:265-296), counter-based: attempt a shuffles target (a mod T) of the spectrum
__global__ void __launch_bounds__(kThreads) k_decoy_permute(const md_precursor* __restrict__ prec, const uint32_t* __restrict__ list,
                                                            const uint32_t* __restrict__ att_off, const uint32_t* __restrict__ att_base, uint32_t n_list,
                                                            uint32_t total, uint64_t seed, const __grid_constant__ DecoyTables T,
                                                            const uint64_t* __restrict__ cand_off, const uint64_t* __restrict__ cand_desc,
                                                            const uint64_t* __restrict__ cand_mask, const int64_t* __restrict__ cand_w,
                                                            const uint8_t* __restrict__ idx_rows, AttemptOut O, PeptideView PV) {
  __shared__ uint8_t sseq[MD_MAX_PEPTIDE_LEN * kThreads];
  TSeq seq{sseq + threadIdx.x};
  const uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x;
  if (wi >= total) return;
  const uint32_t li = find_entry(att_off, n_list, wi);
  const uint32_t s = list[li];
  const md_precursor pr = prec[s];
  const uint32_t attempt = att_base[li] + (wi - att_off[li]);
  const uint64_t t0 = cand_off[s], nt = cand_off[s + 1] - t0;
  if (nt == 0) { O.len[wi] = 0; return; }
  const uint64_t c = t0 + attempt % nt;
  // the reference tests the shuffled sequence with fixed modifications only (:278-279)
  if (cand_mask[c] != 0 || !md_in_window(cand_w[c], pr.lo, pr.hi)) { O.len[wi] = 0; return; }
  const uint64_t d = cand_desc[c];
  const uint8_t* row = idx_rows + (d & 0xFFFFFFFFFFull) * 16;
  const uint32_t L = (uint32_t)(d >> 40) & 0xFF;
  bool ok = true;
  for (uint32_t i = 0; i < L; i++) {
    int a = md_alpha_of_code(row[i]);
    if (a < 0) ok = false;  // letters outside the decoy alphabet cannot be written back as a decoy row
    seq.at(i) = (uint8_t)(a < 0 ? 0 : a);
  }
  if (!ok) { O.len[wi] = 0; return; }
  Philox4 rng; rng.init(seed, pr.spectrum_id, attempt, MD_TAG_PERMUTE);
  for (uint32_t i = L; i > 1; i--) {
    uint32_t j = rng.below(i);
    uint8_t a = seq.at(i - 1); seq.at(i - 1) = seq.at(j); seq.at(j) = a;
  }
  store_attempt(O, wi, seq, L, 0, cand_w[c], T, PV);
}

// One CTA per listed spectrum: keep the successes of this round that are new (not equal to an accepted decoy or to an
// earlier success), in attempt order, until the spectrum has n_per decoys (HashSet<Decoy>, decoy_generator.rs:40,164).
// Linear time: accepted decoys and successes go into a shared-memory hash set keyed by the 64-bit sequence hash, each
// entry remembering the lowest ordinal (accepted decoys first, then attempts in order) that carried the key; a success
// is kept iff it is that first carrier.  (Sequences are identified by their 64-bit hash here.)
__global__ void __launch_bounds__(256) k_decoy_select(const uint32_t* __restrict__ list, const uint32_t* __restrict__ att_off,
                                                      const uint32_t* __restrict__ att_base, uint32_t n_per, uint32_t slots, AttemptOut A,
                                                      uint8_t* __restrict__ dec_rows, uint8_t* __restrict__ dec_len, uint64_t* __restrict__ dec_mask,
                                                      int64_t* __restrict__ dec_w, uint64_t* __restrict__ dec_hash, uint32_t* __restrict__ dec_attempt,
                                                      uint32_t* __restrict__ dec_count) {
  extern __shared__ __align__(16) unsigned long long s_key[];   // slots (0 = empty)
  uint32_t* s_ord = reinterpret_cast<uint32_t*>(s_key + slots);   // slots
  uint8_t* s_keep = reinterpret_cast<uint8_t*>(s_ord + slots);    // kMaxRoundAttempts
  __shared__ uint32_t s_scan[256];
  __shared__ uint32_t s_base;
  const uint32_t li = blockIdx.x, s = list[li];
  const uint32_t a0 = att_off[li], na = att_off[li + 1] - a0;
  const uint32_t have = dec_count[s];
  const uint64_t dbase = (uint64_t)s * n_per;
  const uint32_t smask = slots - 1;
  for (uint32_t i = threadIdx.x; i < slots; i += blockDim.x) { s_key[i] = 0ULL; s_ord[i] = 0xFFFFFFFFu; }
  __syncthreads();
  auto insert = [&](unsigned long long h, uint32_t ord) {
    if (h == 0ULL) h = 1ULL;
    uint32_t slot = (uint32_t)h & smask;
    for (;;) {
      const unsigned long long prev = atomicCAS(&s_key[slot], 0ULL, h);
      if (prev == 0ULL || prev == h) { atomicMin(&s_ord[slot], ord); return; }
      slot = (slot + 1) & smask;
    }
  };
  for (uint32_t j = threadIdx.x; j < have; j += blockDim.x) insert(dec_hash[dbase + j], j);
  for (uint32_t a = threadIdx.x; a < na; a += blockDim.x) if (A.len[a0 + a]) insert(A.hash[a0 + a], have + a);
  __syncthreads();
  for (uint32_t a = threadIdx.x; a < na; a += blockDim.x) {
    bool keep = A.len[a0 + a] != 0;
    if (keep) {
      unsigned long long h = A.hash[a0 + a];
      if (h == 0ULL) h = 1ULL;
      uint32_t slot = (uint32_t)h & smask;
      while (s_key[slot] != h) slot = (slot + 1) & smask;
      keep = s_ord[slot] == have + a;
    }
    s_keep[a] = keep ? 1 : 0;
  }
  __syncthreads();
  // ordered compaction: thread t owns the contiguous chunk [t*per, (t+1)*per)
  const uint32_t per = (na + blockDim.x - 1) / blockDim.x;
  const uint32_t b = threadIdx.x * per, e = min(b + per, na);
  uint32_t c = 0;
  for (uint32_t a = b; a < e; a++) c += s_keep[a];
  s_scan[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (uint32_t t = 0; t < blockDim.x; t++) { uint32_t v = s_scan[t]; s_scan[t] = run; run += v; }
    s_base = run;
  }
  __syncthreads();
  uint32_t o = have + s_scan[threadIdx.x];
  for (uint32_t a = b; a < e; a++) {
    if (!s_keep[a]) continue;
    if (o < n_per) {
      const uint64_t src = a0 + a, dst = dbase + o;
      const uint4* sr = reinterpret_cast<const uint4*>(A.rows + src * MD_DECOY_ROW);
      uint4* dr = reinterpret_cast<uint4*>(dec_rows + dst * MD_DECOY_ROW);
      dr[0] = sr[0]; dr[1] = sr[1]; dr[2] = sr[2]; dr[3] = sr[3];
      dec_len[dst] = A.len[src]; dec_mask[dst] = A.mask[src]; dec_w[dst] = A.w[src]; dec_hash[dst] = A.hash[src];
      dec_attempt[dst] = att_base[li] + a;
    }
    o++;
  }
  __syncthreads();
  if (threadIdx.x == 0) dec_count[s] = min(n_per, have + s_base);
}

// Fallback of k_decoy_select for tables that do not fit shared memory (quadratic scan, exact sequence compare).
// One CTA per listed spectrum: keep the successes of this round that are new (not equal to an accepted decoy or to an
// earlier success), in attempt order, until the spectrum has n_per decoys (HashSet<Decoy>, decoy_generator.rs:40,164).
__global__ void __launch_bounds__(256) k_decoy_select_n2(const uint32_t* __restrict__ list, const uint32_t* __restrict__ att_off,
                                                      const uint32_t* __restrict__ att_base, uint32_t n_per, AttemptOut A, uint8_t* __restrict__ dec_rows,
                                                      uint8_t* __restrict__ dec_len, uint64_t* __restrict__ dec_mask, int64_t* __restrict__ dec_w,
                                                      uint64_t* __restrict__ dec_hash, uint32_t* __restrict__ dec_attempt, uint32_t* __restrict__ dec_count) {
  __shared__ uint64_t s_hash[kMaxRoundAttempts];
  __shared__ uint8_t s_keep[kMaxRoundAttempts];
  __shared__ uint32_t s_scan[256];
  __shared__ uint32_t s_base;
  const uint32_t li = blockIdx.x, s = list[li];
  const uint32_t a0 = att_off[li], na = att_off[li + 1] - a0;
  const uint32_t have = dec_count[s];
  const uint64_t dbase = (uint64_t)s * n_per;
  for (uint32_t a = threadIdx.x; a < na; a += blockDim.x) s_hash[a] = A.len[a0 + a] ? A.hash[a0 + a] : 0ULL;
  __syncthreads();
  for (uint32_t a = threadIdx.x; a < na; a += blockDim.x) {
    const uint32_t L = A.len[a0 + a];
    bool keep = L != 0;
    if (keep) {
      const uint64_t h = s_hash[a];
      const uint8_t* mine = A.rows + (uint64_t)(a0 + a) * MD_DECOY_ROW;
      for (uint32_t j = 0; j < have && keep; j++) {
        if (dec_hash[dbase + j] == h && dec_len[dbase + j] == L) {
          const uint8_t* o = dec_rows + (dbase + j) * MD_DECOY_ROW; bool eq = true;
          for (uint32_t i = 0; i < L; i++) if (o[i] != mine[i]) { eq = false; break; }
          if (eq) keep = false;
        }
      }
      for (uint32_t j = 0; j < a && keep; j++) {
        if (s_hash[j] == h && A.len[a0 + j] == L) {
          const uint8_t* o = A.rows + (uint64_t)(a0 + j) * MD_DECOY_ROW; bool eq = true;
          for (uint32_t i = 0; i < L; i++) if (o[i] != mine[i]) { eq = false; break; }
          if (eq) keep = false;
        }
      }
    }
    s_keep[a] = keep ? 1 : 0;
  }
  __syncthreads();
  // ordered compaction: thread t owns the contiguous chunk [t*per, (t+1)*per)
  const uint32_t per = (na + blockDim.x - 1) / blockDim.x;
  const uint32_t b = threadIdx.x * per, e = min(b + per, na);
  uint32_t c = 0;
  for (uint32_t a = b; a < e; a++) c += s_keep[a];
  s_scan[threadIdx.x] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (uint32_t t = 0; t < blockDim.x; t++) { uint32_t v = s_scan[t]; s_scan[t] = run; run += v; }
    s_base = run;
  }
  __syncthreads();
  uint32_t o = have + s_scan[threadIdx.x];
  for (uint32_t a = b; a < e; a++) {
    if (!s_keep[a]) continue;
    if (o < n_per) {
      const uint64_t src = a0 + a, dst = dbase + o;
      const uint4* sr = reinterpret_cast<const uint4*>(A.rows + src * MD_DECOY_ROW);
      uint4* dr = reinterpret_cast<uint4*>(dec_rows + dst * MD_DECOY_ROW);
      dr[0] = sr[0]; dr[1] = sr[1]; dr[2] = sr[2]; dr[3] = sr[3];
      dec_len[dst] = A.len[src]; dec_mask[dst] = A.mask[src]; dec_w[dst] = A.w[src]; dec_hash[dst] = s_hash[a];
      dec_attempt[dst] = att_base[li] + a;
    }
    o++;
  }
  __syncthreads();
  if (threadIdx.x == 0) dec_count[s] = min(n_per, have + s_base);
}

DecoyTables make_tables(const ModTables& M) {
  DecoyTables T;
  memset(&T, 0, sizeof(T));
  const char* alpha = MD_ALPHABET;
  struct E { int64_t m; int a; };
  std::vector<E> v;
  for (int a = 0; a < MD_ALPHABET_SIZE; a++) {
    uint32_t code = md_code_of((uint8_t)alpha[a]);
    T.code_of_a[a] = (uint8_t)code;
    T.has_fix_a[a] = M.has_fix[code]; T.has_var_a[a] = M.has_var[code]; T.var_a[a] = M.var[code];
    T.mprime[a] = M.mass[code] + (M.has_fix[code] ? M.fix[code] : 0);
    v.push_back({T.mprime[a], a});
  }
  std::sort(v.begin(), v.end(), [](const E& x, const E& y) { return x.m != y.m ? x.m < y.m : x.a < y.a; });
  for (int k = 0; k < 32; k++) { T.sorted_m[k] = INT64_MAX; T.sorted_a[k] = 255; T.run_min_a[k] = 255; }
  for (int k = 0; k < MD_ALPHABET_SIZE; k++) { T.sorted_m[k] = v[k].m; T.sorted_a[k] = (uint8_t)v[k].a; }
  for (int k = 0; k < MD_ALPHABET_SIZE; k++) {
    int j = k; while (j > 0 && v[j - 1].m == v[k].m) j--;  // first of the equal-mass run has the lowest alphabet index
    T.run_min_a[k] = (uint8_t)v[j].a;
  }
  // gap tables
  {
    struct G { uint32_t g; int a; };
    std::vector<G> gb, ga;
    for (int a = 0; a < MD_ALPHABET_SIZE; a++) {
      int64_t below = -1, above = -1;
      for (int c = 0; c < MD_ALPHABET_SIZE; c++) {
        if (T.mprime[c] < T.mprime[a] && (below < 0 || T.mprime[c] > below)) below = T.mprime[c];
        if (T.mprime[c] > T.mprime[a] && (above < 0 || T.mprime[c] < above)) above = T.mprime[c];
      }
      gb.push_back({below < 0 ? 0xFFFFFFFFu : (uint32_t)std::min<int64_t>(T.mprime[a] - below, 0xFFFFFFFEll), a});
      ga.push_back({above < 0 ? 0xFFFFFFFFu : (uint32_t)std::min<int64_t>(above - T.mprime[a], 0xFFFFFFFEll), a});
    }
    auto by_gap = [](const G& x, const G& y) { return x.g != y.g ? x.g < y.g : x.a < y.a; };
    std::sort(gb.begin(), gb.end(), by_gap); std::sort(ga.begin(), ga.end(), by_gap);
    for (int k = 0; k < 32; k++) { T.gapb_sorted[k] = 0xFFFFFFFFu; T.gapa_sorted[k] = 0xFFFFFFFFu; }
    T.maskb_prefix[0] = 0; T.maska_prefix[0] = 0;
    for (int k = 0; k < MD_ALPHABET_SIZE; k++) {
      T.gapb_sorted[k] = gb[k].g; T.maskb_prefix[k + 1] = T.maskb_prefix[k] | (gb[k].g == 0xFFFFFFFFu ? 0u : 1u << gb[k].a);
      T.gapa_sorted[k] = ga[k].g; T.maska_prefix[k + 1] = T.maska_prefix[k] | (ga[k].g == 0xFFFFFFFFu ? 0u : 1u << ga[k].a);
    }
    for (int k = MD_ALPHABET_SIZE + 1; k < 33; k++) { T.maskb_prefix[k] = T.maskb_prefix[MD_ALPHABET_SIZE]; T.maska_prefix[k] = T.maska_prefix[MD_ALPHABET_SIZE]; }
  }
  return T;
}

}  // namespace

void decoys_exhaustive_dev(md_ctx* ctx, uint32_t n, uint32_t n_per);  // exhaustive.cu

void decoys_generate_dev(md_ctx* ctx, uint32_t n, uint32_t n_per, int mode, uint64_t seed) {
  IdentifyWorkspace& W = ctx->ws; PeptideStore& P = ctx->peps;
  const size_t slots = (size_t)n * n_per;
  W.dec_rows.need(slots * MD_DECOY_ROW + 64); W.dec_len.need(slots + 1); W.dec_mask.need(slots + 1); W.dec_w.need(slots + 1);
  W.dec_hash.need(slots + 1); W.dec_attempt.need(slots + 1); W.dec_count.need(n + 1);
  MD_CUDA(cudaMemsetAsync(W.dec_count.p, 0, (n + 1) * sizeof(uint32_t), ctx->stream));
  if (!n || !n_per) return;
  if (mode == MD_DECOY_EXHAUSTIVE) { decoys_exhaustive_dev(ctx, n, n_per); return; }
  MD_REQUIRE(mode == MD_DECOY_REFERENCE_RANDOM || mode == MD_DECOY_PERMUTE_TARGET, MD_ERR_INVALID, "unknown decoy mode");

  const DecoyTables T = make_tables(ctx->mods);
  if (mode == MD_DECOY_REFERENCE_RANDOM)
    for (int a = 0; a < MD_ALPHABET_SIZE; a++)
      MD_REQUIRE(T.mprime[a] > 0 && T.mprime[a] < (1 << 29) && T.var_a[a] > -(1 << 29) && T.var_a[a] < (1 << 29), MD_ERR_UNSUPPORTED,
                 "decoy generation works in 32-bit residuals: residue and modification masses must stay below 536 Da");
  PeptideView PV{(const unsigned long long*)P.ht_key.p, P.ht_val.p, P.ht_mask, P.seq.p, P.seq_off.p, P.len.p};
  std::vector<uint64_t> h_cand_off;
  if (mode == MD_DECOY_PERMUTE_TARGET) {
    h_cand_off.resize(n + 1);
    MD_CUDA(cudaMemcpyAsync(h_cand_off.data(), W.cand_off.p, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    MD_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  // per spectrum bookkeeping on the host
  std::vector<uint32_t> used(n, 0), count(n, 0), cap(n);
  // spectra in ascending precursor mass: neighbouring attempts grow sequences of similar length (lockstep passes)
  std::vector<uint32_t> by_mass(n);
  {
    std::vector<md_precursor> hp(n);
    MD_CUDA(cudaMemcpyAsync(hp.data(), W.prec.p, n * sizeof(md_precursor), cudaMemcpyDeviceToHost, ctx->stream));
    MD_CUDA(cudaStreamSynchronize(ctx->stream));
    for (uint32_t s = 0; s < n; s++) by_mass[s] = s;
    std::stable_sort(by_mass.begin(), by_mass.end(), [&](uint32_t a, uint32_t b) { return hp[a].mass < hp[b].mass; });
  }
  for (uint32_t s = 0; s < n; s++) {
    uint64_t c = md_attempt_cap(n_per);
    if (mode == MD_DECOY_PERMUTE_TARGET) c = std::min<uint64_t>(c, (h_cand_off[s + 1] - h_cand_off[s]) * 1000ull);
    cap[s] = (uint32_t)c;
  }
  DevBuf<uint32_t>& d_list = W.t_list; DevBuf<uint32_t>& d_off = W.t_off; DevBuf<uint32_t>& d_base = W.t_base; DevBuf<uint32_t>& d_queue = W.t_queue;
  DevBuf<int>& d_ovf = W.t_ovf;
  d_list.need(n + 1); d_off.need(n + 2); d_base.need(n + 1); d_queue.need(1); d_ovf.need(1);
  MD_CUDA(cudaMemsetAsync(d_ovf.p, 0, sizeof(int), ctx->stream));
  std::vector<uint32_t> list, off, base;
  for (int round = 0; round < 64; round++) {
    list.clear(); off.assign(1, 0); base.clear();
    for (uint32_t si = 0; si < n; si++) {
      const uint32_t s = by_mass[si];
      if (count[s] >= n_per || used[s] >= cap[s]) continue;
      uint32_t want;
      if (used[s] == 0) want = n_per + n_per / 4 + 32;
      else {
        double yield = std::max(0.02, (double)count[s] / (double)used[s]);
        want = (uint32_t)((double)(n_per - count[s]) / yield * 1.15) + 8;
      }
      want = std::min<uint32_t>({want, (uint32_t)kMaxRoundAttempts, cap[s] - used[s]});
      if (off.back() + (uint64_t)want > 0x7FFFFFFFull) break;  // the rest waits for the next round
      list.push_back(s); base.push_back(used[s]); off.push_back(off.back() + want);
    }
    if (list.empty()) break;
    const uint32_t n_list = (uint32_t)list.size(), total = off.back();
    W.att_rows.need((size_t)total * MD_DECOY_ROW + 64); W.att_len.need(total + 1); W.att_mask.need(total + 1); W.att_w.need(total + 1); W.att_hash.need(total + 1);
    AttemptOut O{W.att_rows.p, W.att_len.p, W.att_mask.p, W.att_w.p, W.att_hash.p};
    MD_CUDA(cudaMemcpyAsync(d_list.p, list.data(), n_list * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    MD_CUDA(cudaMemcpyAsync(d_off.p, off.data(), (n_list + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    MD_CUDA(cudaMemcpyAsync(d_base.p, base.data(), n_list * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    MD_CUDA(cudaEventRecord(ctx->ev[4], ctx->stream));
    if (mode == MD_DECOY_REFERENCE_RANDOM) {
      MD_CUDA(cudaMemsetAsync(d_queue.p, 0, sizeof(uint32_t), ctx->stream));
      uint32_t grid = std::min<uint32_t>((total + kThreads - 1) / kThreads, (uint32_t)ctx->n_sm * 8u);
      MD_LAUNCH(ctx, k_decoy_random, grid, kThreads, 0, W.prec.p, d_list.p, d_off.p, d_base.p, n_list, total, d_queue.p, seed, ctx->mods, T, O, PV, d_ovf.p);
    } else {
      MD_LAUNCH(ctx, k_decoy_permute, blocks(total, kThreads), kThreads, 0, W.prec.p, d_list.p, d_off.p, d_base.p, n_list, total, seed, T, W.cand_off.p,
                W.cand_desc.p, W.cand_mask.p, W.cand_w.p, ctx->index.rows.p, O, PV);
    }
    MD_CUDA(cudaEventRecord(ctx->ev[5], ctx->stream));
    ctx->mark("  attempts");
    {
      uint32_t need = 1;                                    // most (accepted + attempted) of one spectrum in this round
      for (uint32_t i = 0; i < n_list; i++) need = std::max(need, count[list[i]] + (off[i + 1] - off[i]));
      uint32_t slots = 1024; while (slots < 2 * need) slots <<= 1;
      const size_t smem = (size_t)slots * 12 + kMaxRoundAttempts;
      if (smem <= 200 * 1024) {
        MD_CUDA(cudaFuncSetAttribute(k_decoy_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        MD_LAUNCH(ctx, k_decoy_select, n_list, 256, smem, d_list.p, d_off.p, d_base.p, n_per, slots, O, W.dec_rows.p, W.dec_len.p, W.dec_mask.p, W.dec_w.p,
                  W.dec_hash.p, W.dec_attempt.p, W.dec_count.p);
      } else {
        MD_LAUNCH(ctx, k_decoy_select_n2, n_list, 256, 0, d_list.p, d_off.p, d_base.p, n_per, O, W.dec_rows.p, W.dec_len.p, W.dec_mask.p, W.dec_w.p,
                  W.dec_hash.p, W.dec_attempt.p, W.dec_count.p);
      }
    }
    MD_CUDA(cudaMemcpyAsync(count.data(), W.dec_count.p, n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    MD_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->mark("  select");
    for (uint32_t i = 0; i < n_list; i++) used[list[i]] += off[i + 1] - off[i];
    { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]); ctx->acc_ms_kdecoy += ms; ctx->acc_attempts += total; }
  }
  const int ovf = d2h_scalar(ctx, d_ovf.p);
  MD_REQUIRE(ovf != 3, MD_ERR_DEVICE, "decoy generation: internal error (substitution filter disagrees with the mass search)");
  MD_REQUIRE(ovf != 2, MD_ERR_UNSUPPORTED, "decoy generation: |weight - precursor| left the 32-bit range (residue or modification masses above ~1000 Da?)");
  MD_REQUIRE(!ovf, MD_ERR_UNSUPPORTED, "variable-modification placement enumeration exceeds 2^22 subsets for one decoy");
}

void decoys_export(md_ctx* ctx, uint32_t n, uint32_t n_per, md_decoy_table* out) {
  IdentifyWorkspace& W = ctx->ws;
  memset(out, 0, sizeof(*out));
  const size_t slots = (size_t)n * n_per;
  std::vector<uint8_t> rows(slots * MD_DECOY_ROW + 1), len(slots + 1);
  std::vector<uint64_t> mask(slots + 1); std::vector<int64_t> w(slots + 1); std::vector<uint32_t> att(slots + 1), count(n + 1);
  if (slots) {
    MD_CUDA(cudaMemcpyAsync(rows.data(), W.dec_rows.p, slots * MD_DECOY_ROW, cudaMemcpyDeviceToHost, ctx->stream));
    MD_CUDA(cudaMemcpyAsync(len.data(), W.dec_len.p, slots, cudaMemcpyDeviceToHost, ctx->stream));
    MD_CUDA(cudaMemcpyAsync(mask.data(), W.dec_mask.p, slots * 8, cudaMemcpyDeviceToHost, ctx->stream));
    MD_CUDA(cudaMemcpyAsync(w.data(), W.dec_w.p, slots * 8, cudaMemcpyDeviceToHost, ctx->stream));
    MD_CUDA(cudaMemcpyAsync(att.data(), W.dec_attempt.p, slots * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  if (n) MD_CUDA(cudaMemcpyAsync(count.data(), W.dec_count.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  MD_CUDA(cudaStreamSynchronize(ctx->stream));
  uint64_t total = 0, bytes = 0;
  for (uint32_t s = 0; s < n; s++) for (uint32_t j = 0; j < count[s]; j++) { total++; bytes += len[(size_t)s * n_per + j]; }
  out->n_spectra = n; out->n = total; out->seq_bytes = bytes;
  out->off = (uint64_t*)malloc((n + 1) * 8); out->seq = (uint8_t*)malloc(bytes + 1); out->seq_off = (uint64_t*)malloc((total + 1) * 8);
  out->var_mask = (uint64_t*)malloc((total + 1) * 8); out->weight = (int64_t*)malloc((total + 1) * 8); out->mod_weight = (int64_t*)malloc((total + 1) * 8);
  out->attempt = (uint32_t*)malloc((total + 1) * 4);
  uint64_t k = 0, b = 0;
  out->off[0] = 0; out->seq_off[0] = 0;
  for (uint32_t s = 0; s < n; s++) {
    for (uint32_t j = 0; j < count[s]; j++) {
      size_t slot = (size_t)s * n_per + j;
      uint32_t L = len[slot];
      int64_t wu = MD_WATER_UDA;
      for (uint32_t i = 0; i < L; i++) { uint8_t code = rows[slot * MD_DECOY_ROW + i]; out->seq[b + i] = md_letter_of(code); wu += kResidueMassByCode[code]; }
      b += L;
      out->seq_off[k + 1] = b; out->var_mask[k] = mask[slot]; out->weight[k] = wu; out->mod_weight[k] = w[slot]; out->attempt[k] = att[slot];
      k++;
    }
    out->off[s + 1] = k;
  }
}
