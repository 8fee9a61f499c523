// score.cu -- K4: spectrum binning + the fused fragment-and-score kernel + per-spectrum top-k.
//
// The reference has no scorer: identification_task writes <spectrum>.fasta / .comet.params and an external Comet
// binary computes b/y cross-correlation (tasks/identification.rs:323-368; utility/comet_parameter.rs:6-124;
// run_splitup_and_identification.sh:47-60).  This file does that step in place, with a Comet-style fast xcorr
// consistent with the emitted parameters (b/y ions, monoisotopic fragments, fragment_bin_tol = tolerance,
// fragment_bin_offset = 0, theoretical_fragment_ions = 1 (no flanking), max_fragment_charge = 3) but defined in
// exact integers so that it is reproducible bit for bit (see oracle/maxdecoy_oracle.cpp, same definition):
//   bin(m/z)   = floor(m/z[uDa] / w) + 1,  w = fragment tolerance in uDa
//   y[bin]     = max over peaks of sqrt(I), scaled to 50 per tenth of the m/z range, peaks <= 5 % of the base peak dropped,
//                quantised to Q16
//   T[b]       = 151*y[b] - sum_{j=b-75..b+75} y[j]          (= 150 * 2^16 * fast_xcorr[b])
//   raw score  = sum over b/y fragments and fragment charges of T[bin];   score = 0.005 * raw / (150 * 2^16)
//
// Kernel design (B200): one persistent CTA of 1024 threads per SM (spectra from a global queue).  Per spectrum the CTA
// stages the sparse binned spectrum (bin, y, prefix sums of y) in shared memory and expands it, tile by tile, into a
// dense int32 table tile of 49152 bins (192 KiB): vectorised zero fill, then one warp per window edge paints the
// piecewise-constant -sum(y) segments with coalesced stores (no atomics: segment values come from the prefix sums),
// then one thread per peak adds the 151*y spike.  Every thread then scores one candidate at a time: the row (residue
// codes, 16-byte padded) comes in with 128-bit loads, the per-letter packed (quotient, remainder) mass table lives in
// registers and is read with one warp shuffle per residue, fragment bins follow from a division-free running (q, r)
// sum, and each fragment costs exactly one shared-memory gather.  Partial scores of a
// candidate chunk stay in shared memory across tiles; top-k packs (score, ordinal) into one 64-bit key and selects with
// warp REDUX max.  Nothing but the PSM rows is written to HBM unless the caller asks for all scores.
#include "cubx.cuh"

namespace {

inline uint32_t blocks(uint64_t n, uint32_t bs = 256) { return (uint32_t)((n + bs - 1) / bs); }

constexpr int kXcorrOffset = 75;
#ifndef MD_SCORE_THREADS
#define MD_SCORE_THREADS 768
#endif
constexpr int kScoreThreads = MD_SCORE_THREADS;
constexpr uint32_t kTileBins = 49152;     // 192 KiB of int32 per CTA
constexpr uint32_t kCandChunk = 1536;     // candidates whose partial scores stay in shared memory across tiles
constexpr uint32_t kPeakCap = 1024;       // binned peaks staged in shared memory (larger spectra read them from HBM)
constexpr uint32_t kMaxBins = 1u << 26;   // table bins per spectrum (byte offsets 4*bin must stay far below kStop4)
constexpr uint32_t kMaxTopK = 128;
constexpr uint32_t kTileCache = 32;
constexpr uint32_t kFastTopK = 8;

// ------------------------------------------------------------------------------------------------
// precursor windows: tasks/identification.rs:203-211 (utility/mod.rs:9-11; models/mass/mod.rs:6-8,14-16)
// every operation rounded separately (no FMA), truncating conversion
// ------------------------------------------------------------------------------------------------
__global__ void k_precursors(const double* __restrict__ pmz, const uint8_t* __restrict__ charge, const uint32_t* __restrict__ sid, uint32_t n,
                             int64_t lppm, int64_t uppm, int64_t abs_lo, int64_t abs_hi, uint32_t id_base, md_precursor* __restrict__ out) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const double mz = pmz[s], zc = (double)charge[s], H = 1.007276;
  const double tl = __dmul_rn(__ddiv_rn(mz, 1000000.0), (double)lppm);
  const double tu = __dmul_rn(__ddiv_rn(mz, 1000000.0), (double)uppm);
  const double b = __dmul_rn(H, zc);
  int64_t P = (int64_t)__dmul_rn(__dsub_rn(__dmul_rn(mz, zc), b), 1000000.0);
  int64_t lo = (int64_t)__dmul_rn(__dsub_rn(__dmul_rn(__dsub_rn(mz, tl), zc), b), 1000000.0);
  int64_t hi = (int64_t)__dmul_rn(__dsub_rn(__dmul_rn(__dadd_rn(mz, tu), zc), b), 1000000.0);
  if (abs_lo != 0 || abs_hi != 0) { lo = P - abs_lo; hi = P + abs_hi; }
  md_precursor pr;
  pr.mass = P; pr.lo = lo; pr.hi = hi; pr.charge = charge[s]; pr.spectrum_id = sid ? sid[s] : id_base + s;
  out[s] = pr;
}

// ------------------------------------------------------------------------------------------------
// K4a: bin one spectrum per warp -> sorted unique (bin, yq) list
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_max_d(double v) {
  for (int o = 16; o; o >>= 1) { double t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
  for (int o = 16; o; o >>= 1) { int t = __shfl_xor_sync(0xffffffffu, v, o); v = t > v ? t : v; }
  return v;
}

struct PeakEval { bool valid; int32_t bin; double raw; };
__device__ __forceinline__ PeakEval eval_peak(double mz, float I, int64_t P, int64_t w) {
  PeakEval r; r.valid = false; r.bin = 0; r.raw = 0.0;
  if (!(I > 0.0f) || !(mz > 0.0) || !(mz < 1.0e7)) return r;
  int64_t mzint = (int64_t)__dmul_rn(mz, 1000000.0);
  if (!(mzint > 0) || !(mzint < P + 50000000LL)) return r;
  r.valid = true; r.bin = (int32_t)(mzint / w) + 1; r.raw = sqrt((double)I);
  return r;
}

__global__ void k_bin_spectra(const uint64_t* __restrict__ peak_off, const double* __restrict__ peak_mz, const float* __restrict__ peak_int,
                              const md_precursor* __restrict__ prec, uint32_t n, int64_t w, uint32_t min_peaks, int32_t* __restrict__ pk_bin,
                              int32_t* __restrict__ pk_yq, uint32_t* __restrict__ pk_pre, uint32_t* __restrict__ pk_count, int32_t* __restrict__ pk_hbin, int* __restrict__ unsorted) {
  const uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= n) return;
  const uint64_t p0 = peak_off[s], p1 = peak_off[s + 1];
  const int64_t P = prec[s].mass;
  // pass A: base peak, highest bin, number of usable peaks, sortedness
  double gmax = 0.0; int hbin = 0; uint32_t nvalid = 0; bool bad = false;
  for (uint64_t i = p0 + lane; i < p1; i += 32) {
    double mz = peak_mz[i];
    if (i > p0 && mz < peak_mz[i - 1]) bad = true;
    PeakEval e = eval_peak(mz, peak_int[i], P, w);
    if (e.valid) { nvalid++; if (e.raw > gmax) gmax = e.raw; if (e.bin > hbin) hbin = e.bin; }
  }
  gmax = warp_max_d(gmax); hbin = warp_max_i(hbin);
  for (int o = 16; o; o >>= 1) nvalid += __shfl_xor_sync(0xffffffffu, nvalid, o);
  if (__any_sync(0xffffffffu, bad)) { if (lane == 0) *unsorted = 1; }
  if (nvalid < min_peaks || nvalid == 0) {
    if (lane == 0) { pk_count[s] = 0; pk_hbin[s] = -1; }  // not scored (Comet minimum_peaks, comet_parameter.rs:62)
    return;
  }
  // pass B: maximum per tenth of the bin range
  const int wsize = hbin / 10 + 1;
  double winmax[10];
#pragma unroll
  for (int k = 0; k < 10; k++) winmax[k] = 0.0;
  for (uint64_t i = p0 + lane; i < p1; i += 32) {
    PeakEval e = eval_peak(peak_mz[i], peak_int[i], P, w);
    if (e.valid) {
      int k = e.bin / wsize;
#pragma unroll
      for (int q = 0; q < 10; q++) if (q == k && e.raw > winmax[q]) winmax[q] = e.raw;
    }
  }
#pragma unroll
  for (int k = 0; k < 10; k++) winmax[k] = warp_max_d(winmax[k]);
  // pass C: emit runs of equal bins (peaks are sorted by m/z, so bins are non-decreasing)
  const double thr = __dmul_rn(0.05, gmax);
  int carry_bin = 0; uint32_t carry_cnt = 0;
  for (uint64_t base = p0; base < p1; base += 32) {
    uint64_t i = base + lane;
    PeakEval e; e.valid = false; e.bin = 0; e.raw = 0.0;
    if (i < p1) e = eval_peak(peak_mz[i], peak_int[i], P, w);
    const bool kept = e.valid && e.raw > thr;
    int32_t yq = 0;
    if (kept) {
      int k = e.bin / wsize; double wm = 0.0;
#pragma unroll
      for (int q = 0; q < 10; q++) if (q == k) wm = winmax[q];
      double y = __dmul_rn(e.raw, __ddiv_rn(50.0, wm));
      yq = (int32_t)__dadd_rn(__dmul_rn(y, 65536.0), 0.5);
    }
    // bin of the previous kept peak = exclusive prefix max over kept bins (0 = none)
    int v = kept ? e.bin : 0, incl = v;
    for (int o = 1; o < 32; o <<= 1) { int t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o && t > incl) incl = t; }
    int prev = __shfl_up_sync(0xffffffffu, incl, 1);
    if (lane == 0) prev = 0;
    if (carry_bin > prev) prev = carry_bin;
    const bool head = kept && e.bin != prev;
    const uint32_t hb = __ballot_sync(0xffffffffu, head);
    const uint32_t rank = __popc(hb & ((2u << lane) - 1));  // heads at or before this lane
    if (kept) {
      uint64_t o = p0 + carry_cnt + rank - 1;
      if (head) pk_bin[o] = e.bin;
      atomicMax(&pk_yq[o], yq);
    }
    carry_cnt += __popc(hb);
    int last = __shfl_sync(0xffffffffu, incl, 31);
    if (last > carry_bin) carry_bin = last;
  }
  if (lane == 0) { pk_count[s] = carry_cnt; pk_hbin[s] = hbin; }
  // exclusive prefix sums of y (mod 2^32: window sums are < 2^31, differences stay exact); carry_cnt + 1 entries at p0 + s
  __syncwarp();
  uint32_t run = 0;
  uint32_t* pre = pk_pre + p0 + s;
  for (uint32_t base = 0; base < carry_cnt; base += 32) {
    const uint32_t i = base + lane;
    const uint32_t v = i < carry_cnt ? (uint32_t)__ldcg(&pk_yq[p0 + i]) : 0u;
    uint32_t incl = v;
    for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
    if (i < carry_cnt) pre[i] = run + incl - v;
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) pre[carry_cnt] = run;
}

// ------------------------------------------------------------------------------------------------
// K4: score
// ------------------------------------------------------------------------------------------------
struct ScoreConst {
  uint32_t w;                 // bin width, uDa
  uint32_t qp, rp;            // proton  = qp*w + rp
  uint32_t q2p, r2p;          // 2*proton
  uint32_t tq4[32], tr[32];   // per residue code: (mass + fixed delta) = q*w + r; tq4 = 4*q (table byte offsets)
  uint32_t vq4[32], vr[32];   // per residue code: (mass + fixed + variable delta)
  uint32_t max_frag_charge;
  uint32_t top_k, n_per;
};

struct ScoreArgs {
  const md_precursor* prec; uint32_t n_spec;
  const uint64_t* peak_off; const int32_t* pk_bin; const int32_t* pk_yq; const uint32_t* pk_pre; const uint32_t* pk_count; const int32_t* pk_hbin;
  const uint64_t* cand_off; const uint64_t* cand_desc; const uint64_t* cand_mask; const int64_t* cand_w; const uint32_t* cand_pep;
  const uint8_t* idx_rows;
  const uint8_t* dec_rows; const uint8_t* dec_len; const uint64_t* dec_mask; const int64_t* dec_w; const uint32_t* dec_count;
  int64_t* tscore; int64_t* dscore;   // raw score of every candidate, or NULL (PSM rows only)
  md_psm* psm;
  uint32_t* work;
  unsigned long long* stat64;  // [0] pairs scored, [1] algorithmic bytes (14 + len per pair)
  int* error;                  // set when a spectrum needs more table bins than kMaxBins
  unsigned long long* timing;  // MD_SCORE_TIMING=1: per-phase SM cycles summed over CTAs (thread 0's clock), else NULL
};

__device__ __forceinline__ uint32_t div3(uint32_t x) { return __umulhi(x, 0xAAAAAAABu) >> 1; }

// Raw score of one candidate against the table tile [t0, t0+tn).
//   b ion of split k, charge c: floor((B_k + c*proton) / (c*w)) + 1;  y ion: floor((modw - B_k + c*proton) / (c*w)) + 1
// with X = B_k + proton kept as (Q, R), X = Q*w + R, so no fragment needs a division.
// one table gather, branch-free: the byte offset is clamped to the always-zero slot behind the tile
__device__ __forceinline__ int32_t lds_s32(uint32_t addr) {
  int32_t v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void acc_wide(int64_t& acc, int32_t v) {  // acc += v as one IMAD.WIDE
  asm("mad.wide.s32 %0, %1, 1, %0;" : "+l"(acc) : "r"(v));
}
#define MD_GATHER(acc, off4) acc_wide(acc, lds_s32(tab_s + min((uint32_t)(off4), tn4)))
struct LaneTab { uint32_t q4, r, vq4, vr; };   // lane = residue code
constexpr uint32_t kStop4 = 1u << 30;           // added to the running offset at the last residue: every later bin misses

struct CandRef { const uint4* row; uint32_t len; uint64_t mask; int64_t modw; };

// Raw scores of NC candidates of one thread against the table tile [t0, t0+tn).
//   b ion of split k, charge c: floor((B_k + c*proton) / (c*w)) + 1;  y ion: floor((modw - B_k + c*proton) / (c*w)) + 1
// with X = B_k + proton kept as (Q, R), X = Q*w + R, R < w, so no fragment needs a division.  Q is carried as the byte
// offset of the charge-1 b bin in the tile (QB = 4*(Q + 1 - t0), wrapping); bins outside the tile wrap or overshoot and
// are clamped onto the zero slot tab[tn] by one unsigned min.  The last residue is never a prefix: at position len-1
// kStop4 is added to QB, which throws every later b and y bin of every charge out of the tile -- no per-residue
// length predicate.  The loop runs in 4-residue words up to the longest candidate of the warp.
template <int NCH, bool HASVAR, int NC>
__device__ __forceinline__ void score_multi(const CandRef (&cr)[NC], uint32_t maxlen, uint32_t tab_s, uint32_t t0, uint32_t tn, const ScoreConst& C,
                                            const LaneTab& L, int64_t (&out)[NC]) {
  const uint32_t w = C.w, tn4 = 4u * tn, off4 = 4u * (1u - t0);
  const uint32_t K2 = 4u * C.qp - off4, K3 = 4u * C.q2p - off4;
  const uint32_t nword = maxlen > 1 ? (maxlen - 1 + 3) >> 2 : 0;  // warp-uniform
  uint32_t Y1[NC], Rt1[NC], Y2[NC], Rt2[NC], Y3[NC], Rt3[NC], QB[NC], R1[NC], nsplit[NC];
  int64_t accb[NC], accy[NC];
#pragma unroll
  for (int i = 0; i < NC; i++) {
    // T_c = modw + (c+1)*proton  ->  (Qt, Rt)
    const uint64_t T1 = (uint64_t)cr[i].modw + 2ull * MD_PROTON_UDA;
    uint32_t Qt1 = (uint32_t)(T1 / w); Rt1[i] = (uint32_t)(T1 - (uint64_t)Qt1 * w);
    uint32_t Qt2 = Qt1 + C.qp; Rt2[i] = Rt1[i] + C.rp; if (Rt2[i] >= w) { Rt2[i] -= w; Qt2++; }
    uint32_t Qt3 = Qt2 + C.qp; Rt3[i] = Rt2[i] + C.rp; if (Rt3[i] >= w) { Rt3[i] -= w; Qt3++; }
    Y1[i] = 4u * Qt1 + 2u * off4;               // y1 offset = Y1 - QB - 4*borrow
    Y2[i] = 4u * Qt2 + off4;                    // 4*(Qt2 - Q - borrow) = Y2 - QB - 4*borrow, then halved
    Y3[i] = 4u * Qt3 + off4;
    QB[i] = 4u * C.qp + off4; R1[i] = C.rp;     // X = B_k + proton
    nsplit[i] = cr[i].len > 0 ? cr[i].len - 1 : 0;   // residues 0..len-2 are followed by a split
    if (nsplit[i] == 0) QB[i] += kStop4;
    accb[i] = 0; accy[i] = 0;
  }
  for (uint32_t c = 0; c * 4 < nword; c++) {
    uint32_t words[NC][4];
#pragma unroll
    for (int i = 0; i < NC; i++) { const uint4 v = __ldg(cr[i].row + c); words[i][0] = v.x; words[i][1] = v.y; words[i][2] = v.z; words[i][3] = v.w; }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (c * 4 + k >= nword) break;
#pragma unroll
      for (int j = 0; j < 4; j++) {
#pragma unroll
        for (int i = 0; i < NC; i++) {
          const uint32_t pos = c * 16 + k * 4 + j;
          const uint32_t code = words[i][k] >> (8 * j);            // (shfl takes the source lane modulo 32; codes are < 32)
          uint32_t q4 = __shfl_sync(0xffffffffu, L.q4, code), r = __shfl_sync(0xffffffffu, L.r, code);
          if (HASVAR) {
            const uint32_t q4v = __shfl_sync(0xffffffffu, L.vq4, code), rv = __shfl_sync(0xffffffffu, L.vr, code);
            if ((cr[i].mask >> pos) & 1) { q4 = q4v; r = rv; }
          }
          QB[i] += q4; R1[i] += r;
          if (R1[i] >= w) { R1[i] -= w; QB[i] += 4u; }
          {  // fragment charge 1
            const uint32_t yb = Y1[i] - QB[i] - (Rt1[i] < R1[i] ? 4u : 0u);
            MD_GATHER(accb[i], QB[i]); MD_GATHER(accy[i], yb);
          }
          if (NCH >= 2) {
            const uint32_t x4 = QB[i] + K2 + (R1[i] + C.rp >= w ? 4u : 0u);
            const uint32_t bb = ((x4 >> 1) & ~3u) + off4;
            const uint32_t y4 = Y2[i] - QB[i] - (Rt2[i] < R1[i] ? 4u : 0u);
            const uint32_t yb = ((y4 >> 1) & ~3u) + off4;
            MD_GATHER(accb[i], bb); MD_GATHER(accy[i], yb);
          }
          if (NCH >= 3) {
            const uint32_t x4 = QB[i] + K3 + (R1[i] + C.r2p >= w ? 4u : 0u);
            const uint32_t bb = ((__umulhi(x4, 0xAAAAAAABu) >> 1) & ~3u) + off4;   // 4*floor(x/3) from 4*x
            const uint32_t y4 = Y3[i] - QB[i] - (Rt3[i] < R1[i] ? 4u : 0u);
            const uint32_t yb = ((__umulhi(y4, 0xAAAAAAABu) >> 1) & ~3u) + off4;
            MD_GATHER(accb[i], bb); MD_GATHER(accy[i], yb);
          }
          if (pos + 1 == nsplit[i]) QB[i] += kStop4;   // the next residue is the last one
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NC; i++) out[i] = accb[i] + accy[i];
}

__device__ __forceinline__ CandRef cand_ref(const ScoreArgs& A, uint32_t s, uint32_t v, uint32_t nt, uint64_t t0c, uint32_t n_per) {
  CandRef r;
  if (v < nt) {
    const uint64_t c = t0c + v, d = A.cand_desc[c];
    r.row = reinterpret_cast<const uint4*>(A.idx_rows + (d & 0xFFFFFFFFFFull) * 16);
    r.len = (uint32_t)(d >> 40) & 0xFF; r.mask = A.cand_mask[c]; r.modw = A.cand_w[c];
  } else {
    const uint64_t j = (uint64_t)s * n_per + (v - nt);
    r.row = reinterpret_cast<const uint4*>(A.dec_rows + j * MD_DECOY_ROW);
    r.len = A.dec_len[j]; r.mask = A.dec_mask[j]; r.modw = A.dec_w[j];
  }
  return r;
}

// PSM order: raw score descending, candidate ordinal ascending  <=>  key descending
constexpr int64_t kKeyBias = 1ll << 39;
__device__ __forceinline__ unsigned long long psm_key(int64_t score, uint32_t v) {
  return ((unsigned long long)(score + kKeyBias) << 24) | (unsigned long long)(0xFFFFFFu - v);
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long k) {
  const uint32_t hi = (uint32_t)(k >> 32);
  const uint32_t mh = __reduce_max_sync(0xffffffffu, hi);
  const uint32_t ml = __reduce_max_sync(0xffffffffu, hi == mh ? (uint32_t)k : 0u);
  return ((unsigned long long)mh << 32) | ml;
}

__device__ __forceinline__ void write_psm_row(const ScoreArgs& A, const ScoreConst& C, const md_precursor& pr, uint32_t s, uint32_t r, unsigned long long k,
                                              uint32_t nt, uint32_t nd, uint64_t t0c) {
  md_psm row;
  row.spectrum_id = pr.spectrum_id; row.rank = 0; row.is_decoy = 0; row.charge = (uint8_t)pr.charge; row.candidate = 0; row.var_mask = 0;
  row.mod_weight = 0; row.raw_score = 0; row.score = 0.0f; row.n_targets = nt; row.n_decoys = nd; row._pad = 0;
  if (k != 0ull) {
    const uint32_t bv = 0xFFFFFFu - (uint32_t)(k & 0xFFFFFFull);
    const int64_t bs = (int64_t)(k >> 24) - kKeyBias;
    row.rank = (uint16_t)(r + 1); row.raw_score = bs;
    row.score = (float)(0.005 * (double)bs / (150.0 * 65536.0));
    if (bv < nt) { row.is_decoy = 0; row.candidate = (uint64_t)A.cand_pep[t0c + bv] + 1; row.var_mask = A.cand_mask[t0c + bv]; row.mod_weight = A.cand_w[t0c + bv]; }
    else { const uint64_t j = (uint64_t)s * C.n_per + (bv - nt); row.is_decoy = 1; row.candidate = bv - nt; row.var_mask = A.dec_mask[j]; row.mod_weight = A.dec_w[j]; }
  }
  A.psm[(uint64_t)s * C.top_k + r] = row;
}

// first index in [0, n) with a[i] >= x
__device__ __forceinline__ uint32_t lower_bound_i32(const int32_t* a, uint32_t n, int32_t x) {
  uint32_t l = 0, h = n;
  while (l < h) { const uint32_t m = (l + h) >> 1; if (a[m] < x) l = m + 1; else h = m; }
  return l;
}

// One warp's share of a tile pass: units of 32 candidates in length-descending order (s_order), fetched from a
// shared counter, so that warps stay busy until the chunk is done and every warp runs candidates of one length.
template <int NCH, bool HASVAR>
__device__ __forceinline__ void score_units(const ScoreArgs& A, const ScoreConst& C, uint32_t s, uint32_t c0, uint32_t cn, uint32_t nt, uint64_t t0c,
                                            uint32_t tab_s, uint32_t t0, uint32_t tn, const LaneTab& L, int64_t* s_score,
                                            const uint16_t* s_order, uint32_t* s_unit) {
  const uint32_t lane = threadIdx.x & 31, units = (cn + 31) >> 5;
  for (;;) {
    uint32_t u = 0;
    if (lane == 0) u = atomicAdd(s_unit, 1u);
    u = __shfl_sync(0xffffffffu, u, 0);
    if (u >= units) break;
    const uint32_t slot = u * 32 + lane;
    CandRef cr[1];
    cr[0].row = reinterpret_cast<const uint4*>(A.idx_rows); cr[0].len = 0; cr[0].mask = 0; cr[0].modw = 0;
    uint32_t v = 0;
    if (slot < cn) {
      v = s_order[slot];
      cr[0] = cand_ref(A, s, c0 + v, nt, t0c, C.n_per);
    }
    const uint32_t maxlen = __reduce_max_sync(0xffffffffu, cr[0].len);
    int64_t part[1];
    score_multi<NCH, HASVAR, 1>(cr, maxlen, tab_s, t0, tn, C, L, part);
    if (slot < cn) s_score[v] += part[0];
  }
}

struct SpecShared {
  uint32_t work[2];
  unsigned long long wkey[2][kScoreThreads / 32];
  unsigned long long top[kMaxTopK];
  long long tacc[8];
  unsigned long long wtop[kScoreThreads / 32][kFastTopK];  // per-warp best keys, descending
  uint32_t hist[64];        // candidates per length (counting sort of the chunk)
  uint32_t unit;            // next unit of the current tile pass
  uint32_t tile_pa[kTileCache], tile_pb[kTileCache];  // peaks that can reach each of the first tiles
};

template <bool HASVAR>
__global__ void __launch_bounds__(kScoreThreads, 1) k_score(const __grid_constant__ ScoreArgs A, const __grid_constant__ ScoreConst C) {
  extern __shared__ __align__(16) int32_t tab[];                     // kTileBins
  int64_t* s_score = reinterpret_cast<int64_t*>(tab + kTileBins + 4); // kCandChunk (tab[kTileBins] = always-zero sentinel slot)
  int32_t* s_bin = reinterpret_cast<int32_t*>(s_score + kCandChunk); // kPeakCap
  int32_t* s_yq = s_bin + kPeakCap;                                  // kPeakCap
  uint32_t* s_pre = reinterpret_cast<uint32_t*>(s_yq + kPeakCap);    // kPeakCap + 4
  uint16_t* s_order = reinterpret_cast<uint16_t*>(s_pre + kPeakCap + 4);  // kCandChunk: chunk slots in length-descending order
  __shared__ SpecShared sh;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t NW = kScoreThreads / 32;
  // per-letter packed (q, r) tables in registers: lane = residue code
  const LaneTab L{C.tq4[lane], C.tr[lane], C.vq4[lane], C.vr[lane]};
  uint32_t my_pairs = 0, my_bytes = 0;   // per thread, summed at the end
  long long t_last = A.timing ? clock64() : 0;
  if (tid < 8) sh.tacc[tid] = 0;
#define MD_TICK(k) do { if (A.timing && tid == 0) { const long long now_ = clock64(); sh.tacc[k] += now_ - t_last; t_last = now_; } } while (0)

  const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(tab);
  if (tid == 0) sh.work[0] = atomicAdd(A.work, 1u);
  __syncthreads();
  for (uint32_t it = 0;; it++) {
    const uint32_t s = sh.work[it & 1];
    if (s >= A.n_spec) break;
    if (tid == 0) sh.work[(it + 1) & 1] = atomicAdd(A.work, 1u);     // next spectrum, read after this one's barriers
    const md_precursor pr = A.prec[s];
    const uint64_t t0c = A.cand_off[s];
    const uint32_t nt = (uint32_t)(A.cand_off[s + 1] - t0c);
    const uint32_t nd = A.dec_count ? A.dec_count[s] : 0;
    const uint32_t ncand = nt + nd;
    const int32_t hbin = A.pk_hbin[s];
    const uint32_t npk = A.pk_count[s];
    const uint64_t pk0 = A.peak_off[s];
    const uint32_t K = C.top_k;
    uint32_t nch = pr.charge > 1 ? pr.charge - 1 : 1;
    if (nch > C.max_frag_charge) nch = C.max_frag_charge;
    if (nch < 1) nch = 1;
    const uint32_t NB = hbin >= 0 ? (uint32_t)hbin + kXcorrOffset + 1 : 0;  // table bins [0, NB)
    bool scored = hbin >= 0;
    if (NB > kMaxBins || ncand > 0xFFFFFFu) { if (tid == 0) *A.error = 1; scored = false; }

    // ---- stage the binned spectrum
    const int32_t* pbin = A.pk_bin + pk0; const int32_t* pyq = A.pk_yq + pk0; const uint32_t* ppre = A.pk_pre + pk0 + s;
    if (scored && npk <= kPeakCap) {
      for (uint32_t i = tid; i < npk; i += kScoreThreads) { s_bin[i] = pbin[i]; s_yq[i] = pyq[i]; }
      for (uint32_t i = tid; i <= npk; i += kScoreThreads) s_pre[i] = ppre[i];
      pbin = s_bin; pyq = s_yq; ppre = s_pre;
    }
    for (uint32_t r = tid; r < K; r += kScoreThreads) sh.top[r] = 0ull;
    __syncthreads();
    // peaks whose window edges can reach each tile: bin in [t0 - 230, t0 + tn + 76)
    if (scored && tid < kTileCache && (uint64_t)tid * kTileBins < NB) {
      const uint32_t t0 = tid * kTileBins, tn = min(kTileBins, NB - t0);
      sh.tile_pa[tid] = lower_bound_i32(pbin, npk, (int32_t)t0 - 230);
      sh.tile_pb[tid] = lower_bound_i32(pbin, npk, (int32_t)(t0 + tn) + kXcorrOffset + 1);
    }
    MD_TICK(0);

    const bool fast = K <= kFastTopK && ncand <= kCandChunk;   // warp-local top-k, merged by warp 0 while the others move on
    for (uint32_t c0 = 0; c0 < ncand || c0 == 0; c0 += kCandChunk) {
      const uint32_t cn = min(kCandChunk, ncand - c0);
      for (uint32_t v = tid; v < cn; v += kScoreThreads) s_score[v] = 0;
      // counting sort of the chunk by peptide length, longest first
      if (tid < 64) sh.hist[tid] = 0;
      __syncthreads();
      if (scored) {
        for (uint32_t v = tid; v < cn; v += kScoreThreads) {
          const uint32_t u = c0 + v;
          const uint32_t len = u < nt ? (uint32_t)(A.cand_desc[t0c + u] >> 40) & 0xFF : A.dec_len[(uint64_t)s * C.n_per + (u - nt)];
          atomicAdd(&sh.hist[63u - min(len, 63u)], 1u);
          my_pairs++; my_bytes += 14 + len;
        }
      }
      __syncthreads();
      if (warp == 0) {  // exclusive prefix over the 64 buckets
        const uint32_t h0 = sh.hist[2 * lane], h1 = sh.hist[2 * lane + 1];
        uint32_t incl = h0 + h1;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
        sh.hist[2 * lane] = incl - h0 - h1; sh.hist[2 * lane + 1] = incl - h1;
      }
      __syncthreads();
      if (scored) {
        for (uint32_t v = tid; v < cn; v += kScoreThreads) {
          const uint32_t u = c0 + v;
          const uint32_t len = u < nt ? (uint32_t)(A.cand_desc[t0c + u] >> 40) & 0xFF : A.dec_len[(uint64_t)s * C.n_per + (u - nt)];
          s_order[atomicAdd(&sh.hist[63u - min(len, 63u)], 1u)] = (uint16_t)v;
        }
      }
      // (visible to the scoring warps after the table-build barriers)
      if (scored && cn) {
        for (uint32_t t0 = 0; t0 < NB; t0 += kTileBins) {
          const uint32_t tn = min(kTileBins, NB - t0);
          const uint32_t ti = t0 / kTileBins;
          uint32_t pa, pb;
          if (ti < kTileCache) { pa = sh.tile_pa[ti]; pb = sh.tile_pb[ti]; }
          else { pa = lower_bound_i32(pbin, npk, (int32_t)t0 - 230); pb = lower_bound_i32(pbin, npk, (int32_t)(t0 + tn) + kXcorrOffset + 1); }
          if (pa == pb) continue;  // an all-zero tile adds nothing
          // (1) zero the tile
          {
            uint4* z = reinterpret_cast<uint4*>(tab);
            const uint32_t n4 = (tn + 4) >> 2;            // incl. the zero slot tab[tn] every out-of-tile gather lands on
            for (uint32_t i = tid; i < n4; i += kScoreThreads) z[i] = make_uint4(0, 0, 0, 0);
            if (tid == 0) sh.unit = 0;
          }
          __syncthreads();
          MD_TICK(1);
          // (2) paint -S[b], S[b] = sum of y over bins [b-75, b+75]: piecewise constant between window edges (no atomics:
          //     the value of every segment comes from the prefix sums of y).
          //     Edge 2k   = entry of peak pa+k at x = bin-75: S = pre[p+1] - pre[lo], lo = first peak with bin >= x-75
          //     Edge 2k+1 = exit  of peak pa+k at x = bin+76: S = pre[hi] - pre[p+1], hi = first peak with bin > x+75
          //     and the segment runs to the next edge of either kind.  One warp per edge, coalesced stores.
          for (uint32_t e = warp; e < 2 * (pb - pa); e += NW) {
            const uint32_t p = pa + (e >> 1);
            const int32_t bp = pbin[p];
            int32_t xa, xb; uint32_t S;
            if ((e & 1) == 0) {
              uint32_t lo = p;
              while (lo > 0 && pbin[lo - 1] >= bp - 2 * kXcorrOffset) lo--;
              S = ppre[p + 1] - ppre[lo];
              xa = bp - kXcorrOffset;
              xb = pbin[lo] + kXcorrOffset + 1;                                  // next exit
              if (p + 1 < npk) xb = min(xb, pbin[p + 1] - kXcorrOffset);         // next entry
            } else {
              uint32_t hi = p + 1;
              while (hi < npk && pbin[hi] <= bp + 2 * kXcorrOffset + 1) hi++;
              S = ppre[hi] - ppre[p + 1];
              xa = bp + kXcorrOffset + 1;
              xb = hi < npk ? pbin[hi] - kXcorrOffset : INT32_MAX;               // next entry
              if (p + 1 < npk) xb = min(xb, pbin[p + 1] + kXcorrOffset + 1);     // next exit
            }
            if (S == 0) continue;
            const int32_t a = max(xa, (int32_t)t0), b = min(xb, (int32_t)(t0 + tn));
            const int32_t val = -(int32_t)S;
            for (int32_t x = a + (int32_t)lane; x < b; x += 32) tab[x - (int32_t)t0] = val;
          }
          __syncthreads();
          MD_TICK(2);
          // (3) the peak's own bin: + 151*y
          for (uint32_t i = tid; i < pb - pa; i += kScoreThreads) {
            const uint32_t x = (uint32_t)pbin[pa + i] - t0;
            if (x < tn) tab[x] += 151 * pyq[pa + i];
          }
          __syncthreads();
          MD_TICK(3);
          // (4) score the chunk against the tile
          switch (nch) {
            case 1: score_units<1, HASVAR>(A, C, s, c0, cn, nt, t0c, tab_s, t0, tn, L, s_score, s_order, &sh.unit); break;
            case 2: score_units<2, HASVAR>(A, C, s, c0, cn, nt, t0c, tab_s, t0, tn, L, s_score, s_order, &sh.unit); break;
            default: score_units<3, HASVAR>(A, C, s, c0, cn, nt, t0c, tab_s, t0, tn, L, s_score, s_order, &sh.unit); break;
          }
          __syncthreads();
          MD_TICK(4);
        }
      }
      // ---- raw scores of every candidate (on request)
      if (A.tscore) {
        for (uint32_t v = tid; v < cn; v += kScoreThreads) {
          const uint32_t u = c0 + v;
          if (u < nt) A.tscore[t0c + u] = s_score[v]; else A.dscore[(uint64_t)s * C.n_per + (u - nt)] = s_score[v];
        }
      }
      // ---- top-k of (this chunk's candidates) U (best of the earlier chunks): K rounds of block-wide max
      if (!fast && scored && cn && K) {
        // scores -> keys, in place; slots cn..cn+K-1 (conceptually) hold the running list, owned by threads < K
        for (uint32_t v = tid; v < cn; v += kScoreThreads) s_score[v] = (int64_t)psm_key(s_score[v], c0 + v);
        unsigned long long carry = tid < K ? sh.top[tid] : 0ull;   // this thread's entry of the running list
        __syncthreads();
        unsigned long long prev = ~0ull;
        for (uint32_t r = 0; r < K; r++) {
          unsigned long long best = (carry < prev) ? carry : 0ull;
          for (uint32_t v = tid; v < cn; v += kScoreThreads) {
            const unsigned long long k = (unsigned long long)s_score[v];
            if (k < prev && k > best) best = k;
          }
          best = warp_max_u64(best);
          if (lane == 0) sh.wkey[r & 1][warp] = best;
          __syncthreads();
          unsigned long long m = lane < NW ? sh.wkey[r & 1][lane] : 0ull;
          m = warp_max_u64(m);
          if (tid == 0) sh.top[r] = m;
          prev = m;
          if (m == 0ull) break;   // fewer candidates than rows (uniform: every thread sees the same m)
        }
        __syncthreads();
        MD_TICK(5);
      }
    }
    // ---- PSM rows
    if (fast) {
      if (scored && K) {
        unsigned long long k0 = tid < ncand ? psm_key(s_score[tid], tid) : 0ull;
        unsigned long long k1 = tid + kScoreThreads < ncand ? psm_key(s_score[tid + kScoreThreads], tid + kScoreThreads) : 0ull;
        for (uint32_t r = 0; r < K; r++) {
          const unsigned long long wm = warp_max_u64(k0 > k1 ? k0 : k1);
          if (wm != 0ull) { if (k0 == wm) k0 = 0ull; else if (k1 == wm) k1 = 0ull; }
          if (lane == 0) sh.wtop[warp][r] = wm;
        }
      }
      __syncthreads();
      MD_TICK(5);
      if (warp == 0) {
        unsigned long long mine = 0ull;      // lane r ends up with the r-th best key of the spectrum
        if (scored && K) {
          uint32_t idx = 0;
          for (uint32_t r = 0; r < K; r++) {
            const unsigned long long head = (lane < NW && idx < K) ? sh.wtop[lane][idx] : 0ull;
            const unsigned long long wm = warp_max_u64(head);
            if (wm != 0ull && head == wm) idx++;
            if (lane == r) mine = wm;
          }
        }
        if (lane < K) write_psm_row(A, C, pr, s, lane, mine, nt, nd, t0c);
      }
    } else {
      for (uint32_t r = tid; r < K; r += kScoreThreads) write_psm_row(A, C, pr, s, r, scored ? sh.top[r] : 0ull, nt, nd, t0c);
      __syncthreads();
      MD_TICK(6);
    }
  }
  if (A.timing && tid == 0) for (int k = 0; k < 8; k++) atomicAdd(&A.timing[k], (unsigned long long)sh.tacc[k]);
  unsigned long long wp = my_pairs, wb = my_bytes;
  for (int o = 16; o; o >>= 1) { wp += __shfl_xor_sync(0xffffffffu, wp, o); wb += __shfl_xor_sync(0xffffffffu, wb, o); }
  if (lane == 0 && wp) { atomicAdd(&A.stat64[0], wp); atomicAdd(&A.stat64[1], wb); }
}

void split_qr(int64_t m, uint32_t w, uint32_t* q, uint32_t* r) {
  if (m < 0) m = 0;
  *q = (uint32_t)(m / w); *r = (uint32_t)(m % w);
}

}  // namespace

void precursors_dev(md_ctx* ctx, const SpectraDev& S, const md_search_params& p, uint32_t id_base) {
  ctx->ws.prec.need(S.n + 1);
  if (!S.n) return;
  MD_LAUNCH(ctx, k_precursors, blocks(S.n), 256, 0, S.pmz, S.charge, S.sid, S.n, p.lower_ppm, p.upper_ppm, p.abs_lower_uda, p.abs_upper_uda, id_base, ctx->ws.prec.p);
}

void score_run_dev(md_ctx* ctx, const SpectraDev& S, uint64_t n_peaks, const md_search_params& p, uint32_t n_per, md_psm* psm_dev, bool want_all) {
  IdentifyWorkspace& W = ctx->ws;
  const uint32_t n = S.n;
  if (!n) return;
  const int64_t w = (int64_t)llround(p.fragment_tolerance * 1000000.0);
  const uint32_t mfc = p.max_fragment_charge ? p.max_fragment_charge : 3;
  MD_REQUIRE(mfc <= 3, MD_ERR_UNSUPPORTED, "max_fragment_charge > 3 (the reference fixes it to 3: comet_parameter.rs:55)");
  MD_REQUIRE(p.top_k <= kMaxTopK, MD_ERR_UNSUPPORTED, "top_k > 128");
  // ---- K4a
  W.pk_bin.need(n_peaks + 1); W.pk_yq.need(n_peaks + 1); W.pk_pre.need(n_peaks + n + 2); W.pk_count.need(n + 1); W.pk_hbin.need(n + 1);
  DevBuf<int>& d_flag = W.t_unsorted; d_flag.need(2);
  MD_CUDA(cudaMemsetAsync(d_flag.p, 0, 2 * sizeof(int), ctx->stream));
  MD_CUDA(cudaMemsetAsync(W.pk_yq.p, 0, (n_peaks + 1) * sizeof(int32_t), ctx->stream));
  MD_LAUNCH(ctx, k_bin_spectra, blocks((uint64_t)n * 32, 128), 128, 0, S.peak_off, S.peak_mz, S.peak_int, W.prec.p, n, w, p.min_peaks, W.pk_bin.p, W.pk_yq.p,
            W.pk_pre.p, W.pk_count.p, W.pk_hbin.p, d_flag.p);
  // ---- K4
  ScoreConst C;
  memset(&C, 0, sizeof(C));
  C.w = (uint32_t)w; C.max_frag_charge = mfc; C.top_k = p.top_k; C.n_per = n_per;
  split_qr(MD_PROTON_UDA, C.w, &C.qp, &C.rp);
  split_qr(2 * MD_PROTON_UDA, C.w, &C.q2p, &C.r2p);
  bool has_var = false;
  for (int c = 0; c < 32; c++) {
    int64_t m = c < MD_NCODES ? ctx->mods.mass[c] : 0;
    int64_t f = (c < MD_NCODES && ctx->mods.has_fix[c]) ? ctx->mods.fix[c] : 0;
    int64_t v = (c < MD_NCODES && ctx->mods.has_var[c]) ? ctx->mods.var[c] : 0;
    uint32_t q, r;
    split_qr(m + f, C.w, &q, &r);
    MD_REQUIRE(q < (1u << 22), MD_ERR_UNSUPPORTED, "fragment_tolerance too small for this residue mass");
    C.tq4[c] = 4u * q; C.tr[c] = r;
    split_qr(m + f + v, C.w, &q, &r);
    MD_REQUIRE(q < (1u << 22), MD_ERR_UNSUPPORTED, "fragment_tolerance too small for this residue mass");
    C.vq4[c] = 4u * q; C.vr[c] = r;
    if (c < MD_NCODES && ctx->mods.has_var[c]) has_var = true;
  }
  uint64_t n_targets = 0;
  if (want_all) {
    MD_CUDA(cudaMemcpyAsync(&n_targets, W.cand_off.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
    MD_CUDA(cudaStreamSynchronize(ctx->stream));
    W.tscore.need(n_targets + 1); W.dscore.need((size_t)n * n_per + 1);
  }
  DevBuf<uint32_t>& work = W.counters; work.need(4);
  MD_CUDA(cudaMemsetAsync(work.p, 0, 4 * sizeof(uint32_t), ctx->stream));
  W.stat64.need(16);
  MD_CUDA(cudaMemsetAsync(W.stat64.p, 0, 16 * sizeof(unsigned long long), ctx->stream));
  const bool timing = getenv("MD_SCORE_TIMING") != nullptr;
  ScoreArgs A;
  A.prec = W.prec.p; A.n_spec = n; A.peak_off = S.peak_off; A.pk_bin = W.pk_bin.p; A.pk_yq = W.pk_yq.p; A.pk_pre = W.pk_pre.p; A.pk_count = W.pk_count.p; A.pk_hbin = W.pk_hbin.p;
  A.cand_off = W.cand_off.p; A.cand_desc = W.cand_desc.p; A.cand_mask = W.cand_mask.p; A.cand_w = W.cand_w.p; A.cand_pep = W.cand_pep.p;
  A.idx_rows = ctx->index.rows.p;
  A.dec_rows = W.dec_rows.p; A.dec_len = W.dec_len.p; A.dec_mask = W.dec_mask.p; A.dec_w = W.dec_w.p; A.dec_count = n_per ? W.dec_count.p : nullptr;
  A.tscore = want_all ? W.tscore.p : nullptr; A.dscore = want_all ? W.dscore.p : nullptr; A.psm = psm_dev; A.work = work.p; A.stat64 = W.stat64.p;
  A.error = d_flag.p + 1; A.timing = timing ? W.stat64.p + 8 : nullptr;
  const size_t smem = ((size_t)kTileBins + 4) * sizeof(int32_t) + (size_t)kCandChunk * sizeof(int64_t) + (size_t)kPeakCap * 8 + ((size_t)kPeakCap + 4) * 4 + (size_t)kCandChunk * 2;
  const uint32_t grid = std::min<uint32_t>(n, (uint32_t)ctx->n_sm);
  MD_CUDA(cudaEventRecord(ctx->ev[6], ctx->stream));
  if (has_var) {
    MD_CUDA(cudaFuncSetAttribute(k_score<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MD_LAUNCH(ctx, k_score<true>, grid, kScoreThreads, smem, A, C);
  } else {
    MD_CUDA(cudaFuncSetAttribute(k_score<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MD_LAUNCH(ctx, k_score<false>, grid, kScoreThreads, smem, A, C);
  }
  MD_CUDA(cudaEventRecord(ctx->ev[7], ctx->stream));
  unsigned long long h_stat[2] = {0, 0};
  int h_flag[2] = {0, 0};
  MD_CUDA(cudaMemcpyAsync(h_stat, W.stat64.p, sizeof(h_stat), cudaMemcpyDeviceToHost, ctx->stream));
  MD_CUDA(cudaMemcpyAsync(h_flag, d_flag.p, sizeof(h_flag), cudaMemcpyDeviceToHost, ctx->stream));
  MD_CUDA(cudaStreamSynchronize(ctx->stream));
  { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]); ctx->acc_ms_kscore += ms; ctx->acc_pairs += h_stat[0]; ctx->acc_score_bytes += h_stat[1]; }
  if (timing) {
    unsigned long long t[8];
    MD_CUDA(cudaMemcpy(t, W.stat64.p + 8, sizeof(t), cudaMemcpyDeviceToHost));
    double tot = 0; for (int k = 0; k < 6; k++) tot += (double)t[k];
    static const char* names[6] = {"stage", "zero", "paint", "spike", "score", "topk"};
    fprintf(stderr, "[md_score_timing] grid=%u", grid);
    for (int k = 0; k < 6; k++) fprintf(stderr, " %s=%.1f%%", names[k], 100.0 * (double)t[k] / tot);
    fprintf(stderr, " cycles/CTA=%.0f\n", tot / grid);
  }
  MD_REQUIRE(!h_flag[0], MD_ERR_INVALID, "spectra: peaks of a spectrum must be sorted by m/z");
  MD_REQUIRE(!h_flag[1], MD_ERR_UNSUPPORTED, "a spectrum needs more than 2^26 fragment bins or has more than 2^24 candidates");
}
