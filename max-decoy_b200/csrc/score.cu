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
// Kernel design (B200): one persistent CTA of 768 threads per SM (spectra from a global queue).  T[b] is non-zero only
// within 75 bins of a peak, i.e. on a few hundred 64-bin blocks of the 100k+ bins a 0.02-Da table spans, so the CTA keeps
// a BLOCK-COMPRESSED table in shared memory: a 16-bit block map (bin >> 6 -> compressed block << 6 | 63, 0 = the all-zero block)
// plus up to 719 dense 64-bin blocks (180 KiB of int32).  A centroided MS2 spectrum (~100-400 peaks above the 5 %
// threshold) fits whole, so every candidate is walked ONCE; denser spectra fall back to several tiles of compressed
// blocks.  Build: peaks mark their blocks in a bitmap (shared-memory atomicOr), a warp scan numbers the blocks, the
// occupied blocks are zeroed with vectorised stores, every peak drops its window as a difference pair (+y at bin-75, -y
// at bin+76) and its spike as (-151y, +151y) with shared-memory atomics, and one block-wide inclusive scan straight over
// the compressed blocks turns the differences into -T (the running sum is 0 at the end of every run of occupied blocks,
// so the scan needs no notion of where the runs are).  Scoring: the
// candidates are counting-sorted by length into 32-candidate units that warps pull from a shared counter; a thread
// scores one candidate: 128-bit row loads, the per-letter (q, r) mass table in registers read by warp shuffle, a
// division-free running (Q, R) prefix, two shared-memory loads per fragment (block map, then table), a stop offset at
// the last residue instead of a per-residue length predicate, IMAD.WIDE accumulation.  Partial scores of a candidate
// chunk stay in shared memory across tiles; top-k packs (score, ordinal) into one 64-bit key, warp-local REDUX max,
// merged by warp 0 while the other warps already stage the next spectrum.  Only PSM rows go to HBM unless the
// caller asks for all scores.
#include "cubx.cuh"

namespace {

inline uint32_t blocks(uint64_t n, uint32_t bs = 256) { return (uint32_t)((n + bs - 1) / bs); }

constexpr int kXcorrOffset = 75;
#ifndef MD_SCORE_THREADS
#define MD_SCORE_THREADS 768
#endif
constexpr int kScoreThreads = MD_SCORE_THREADS;
constexpr uint32_t kBlkShift = 6, kBlk = 1u << kBlkShift;      // table block = 64 bins
constexpr uint32_t kTileBins = 46080;      // 180 KiB of int32 per CTA = 720 blocks, block 0 is the all-zero block
constexpr uint32_t kTileBlocks = kTileBins / kBlk - 1;          // usable blocks per tile
static_assert(kTileBlocks < (1u << (16 - kBlkShift)), "a block-map entry holds the compressed block in its upper 10 bits");
constexpr uint32_t kMapCap = 6144;         // block-map entries staged in shared memory (393k bins; larger tables use an HBM map)
constexpr uint32_t kCandChunk = 1536;      // candidates whose partial scores stay in shared memory across tiles
constexpr uint32_t kPeakCap = 512;         // binned peaks staged in shared memory, double-buffered (larger spectra read them from HBM)
constexpr uint32_t kMaxBins = 1u << 26;    // table bins per spectrum (bins must stay far below kStop)
constexpr uint32_t kMaxTopK = 128;
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
                              int32_t* __restrict__ pk_yq, uint32_t* __restrict__ pk_count, int32_t* __restrict__ pk_hbin, int* __restrict__ unsorted) {
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
}

// ------------------------------------------------------------------------------------------------
// K4: score
// ------------------------------------------------------------------------------------------------
struct ScoreConst {
  uint32_t w;                 // bin width, uDa
  float rcp_w;                // ~ 1/w
  uint32_t qp, rp;            // proton  = qp*w + rp
  uint32_t q2p, r2p;          // 2*proton
  uint32_t tq[32], tr[32];    // per residue code: (mass + fixed delta) = q*w + r
  uint32_t vq[32], vr[32];    // per residue code: (mass + fixed + variable delta)
  uint32_t max_frag_charge;
  uint32_t top_k, n_per;
};

struct ScoreArgs {
  const md_precursor* prec; uint32_t n_spec;
  const uint64_t* peak_off; const int32_t* pk_bin; const int32_t* pk_yq; const uint32_t* pk_count; const int32_t* pk_hbin;
  const uint64_t* cand_off; const uint64_t* cand_desc; const uint64_t* cand_mask; const int64_t* cand_w; const uint32_t* cand_pep;
  const uint8_t* idx_rows;
  uint64_t dec_slots;   // decoy slots of the batch: the second halves of the decoy rows start dec_slots * 32 bytes in (md_dec_byte)
  const uint8_t* dec_rows; const uint8_t* dec_len; const uint64_t* dec_mask; const int64_t* dec_w; const uint32_t* dec_count;
  int64_t* tscore; int64_t* dscore;   // raw score of every candidate, or NULL (PSM rows only)
  md_psm* psm;
  uint32_t* work;
  unsigned long long* stat64;  // [0] pairs scored, [1] algorithmic bytes (14 + len per pair)
  int* error;                  // set when a spectrum needs more table bins than kMaxBins
  unsigned long long* timing;  // MD_SCORE_TIMING=1: per-phase SM cycles summed over CTAs (thread 0's clock), else NULL
  // spectra with very many candidates (open searches) are split into `parts` work items, one contiguous range of candidate
  // chunks each; a part leaves its K best keys in part_top and the last part to finish (parts_done) writes the PSM rows
  uint32_t parts; unsigned long long* part_top; uint32_t* parts_done;
  uint16_t* gmap; uint32_t* gbits; uint32_t gmap_stride;   // per-CTA block map / block bitmap in HBM for tables beyond kMapCap blocks
};

__device__ __forceinline__ uint32_t div3(uint32_t x) { return __umulhi(x, 0xAAAAAAABu) >> 1; }

__device__ __forceinline__ int32_t lds_s32(uint32_t addr) {
  int32_t v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u16(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void acc_wide(int64_t& acc, int32_t v) {  // acc += v as one IMAD.WIDE
  asm("mad.wide.s32 %0, %1, 1, %0;" : "+l"(acc) : "r"(v));
}
// The table as the scoring loop sees it: block map (shared memory address, or an HBM pointer when MAPG) + dense blocks.
struct TableView { uint32_t tab_s, map_s, nblk; const uint16_t* gmap; };
// one fragment: bin -> block (clamped onto the map's zero entry) -> compressed block -> table entry; branch-free
template <bool MAPG>
__device__ __forceinline__ int32_t gather(uint32_t bin, const TableView& V) {
  const uint32_t blk = min(bin >> kBlkShift, V.nblk);
  const uint32_t e = MAPG ? (uint32_t)V.gmap[blk] : lds_u16(V.map_s + 2u * blk);
  // e = (compressed block << 6) | 63, or 0: one AND yields the table index, and every fragment that misses reads the
  // same word (entry 0 of the all-zero block), which the shared-memory pipe serves as one broadcast instead of a
  // bank-conflicting random access
  const uint32_t t = e & (bin | ~(kBlk - 1u));
  return lds_s32(V.tab_s + (t << 2));
}
struct LaneTab { uint32_t q, r, vq, vr; };      // lane = residue code
constexpr uint32_t kStop = 1u << 30;            // added to the running bin at the last residue: every later bin misses

struct CandRef { const uint4* row; const uint4* row_hi; uint32_t len; uint64_t mask; int64_t modw; };   // row: bytes 0..31, row_hi: bytes 32..63

// Raw score of one candidate against the (tile of the) block-compressed table.
//   b ion of split k, charge c: floor((B_k + c*proton) / (c*w)) + 1;  y ion: floor((modw - B_k + c*proton) / (c*w)) + 1
// with X = B_k + proton kept as (Q, R), X = Q*w + R, R < w, so no fragment needs a division.  QB = Q + 1 is the charge-1
// b bin itself; bins past the table, and the wrapped "negative" ones, are clamped onto the map's zero entry by one
// unsigned min.  The last residue is never a prefix: at position len-1 kStop is added to QB, which throws every later
// b and y bin of every charge out of the table -- no per-residue length predicate.  The loop runs in 4-residue words up
// to the longest candidate of the warp.
template <int NCH, bool HASVAR, bool MAPG>
__device__ __forceinline__ int64_t score_one(const CandRef& cr, uint32_t maxlen, const TableView& V, const ScoreConst& C, const LaneTab& L) {
  const uint32_t w = C.w;
  const uint32_t K2 = C.qp - 1u, K3 = C.q2p - 1u;
  const uint32_t nword = maxlen > 1 ? (maxlen - 1 + 3) >> 2 : 0;  // warp-uniform
  // T_c = modw + (c+1)*proton  ->  (Qt, Rt)
  const uint64_t T1 = (uint64_t)cr.modw + 2ull * MD_PROTON_UDA;
  // T1 = Qt1*w + Rt1 without the 64-bit division routine (~100 instructions per candidate): float estimate of the
  // quotient (T1 < 2^34, so it is off by a few units at most), exact remainder, stepwise correction
  uint32_t Qt1 = (uint32_t)__fmul_rz((float)T1, C.rcp_w);
  int64_t rem = (int64_t)T1 - (int64_t)((uint64_t)Qt1 * w);
  while (rem < 0) { Qt1--; rem += w; }
  while (rem >= (int64_t)w) { Qt1++; rem -= w; }
  const uint32_t Rt1 = (uint32_t)rem;
  uint32_t Qt2 = Qt1 + C.qp, Rt2 = Rt1 + C.rp; if (Rt2 >= w) { Rt2 -= w; Qt2++; }
  uint32_t Qt3 = Qt2 + C.qp, Rt3 = Rt2 + C.rp; if (Rt3 >= w) { Rt3 -= w; Qt3++; }
  const uint32_t Y1 = Qt1 + 2u;               // y1 bin = Y1 - QB - borrow
  const uint32_t Y2 = Qt2 + 1u, Y3 = Qt3 + 1u;  // Qt_c - Q - borrow = Y_c - QB - borrow, then halved / divided by 3, + 1
  uint32_t QB = C.qp + 1u, R1 = C.rp;         // X = B_k + proton
  const uint32_t nsplit = cr.len > 0 ? cr.len - 1 : 0;   // residues 0..len-2 are followed by a split
  if (nsplit == 0) QB += kStop;
  int64_t acc = 0;
  for (uint32_t c = 0; c * 4 < nword; c++) {
    const uint4 v = __ldg(c < 2 ? cr.row + c : cr.row_hi + (c - 2));
    const uint32_t words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; k++) {
      if (c * 4 + k >= nword) break;
#pragma unroll
      for (int j = 0; j < 4; j++) {
        const uint32_t pos = c * 16 + k * 4 + j;
        const uint32_t code = words[k] >> (8 * j);            // (shfl takes the source lane modulo 32; codes are < 32)
        uint32_t q = __shfl_sync(0xffffffffu, L.q, code), r = __shfl_sync(0xffffffffu, L.r, code);
        if (HASVAR) {
          const uint32_t qv = __shfl_sync(0xffffffffu, L.vq, code), rv = __shfl_sync(0xffffffffu, L.vr, code);
          if ((cr.mask >> pos) & 1) { q = qv; r = rv; }
        }
        QB += q; R1 += r;
        if (R1 >= w) { R1 -= w; QB++; }
        // |T| <= 151 * 50 * 2^16 < 2^29: up to four entries add up in 32 bits, then one IMAD.WIDE into the 64-bit score
        int32_t s4 = gather<MAPG>(QB, V) + gather<MAPG>(Y1 - QB - (Rt1 < R1 ? 1u : 0u), V);                // fragment charge 1
        if (NCH >= 2) {
          s4 += gather<MAPG>(((QB + K2 + (R1 + C.rp >= w ? 1u : 0u)) >> 1) + 1u, V);
          s4 += gather<MAPG>(((Y2 - QB - (Rt2 < R1 ? 1u : 0u)) >> 1) + 1u, V);
        }
        acc_wide(acc, s4);
        if (NCH >= 3) {
          acc_wide(acc, gather<MAPG>(div3(QB + K3 + (R1 + C.r2p >= w ? 1u : 0u)) + 1u, V) +
                            gather<MAPG>(div3(Y3 - QB - (Rt3 < R1 ? 1u : 0u)) + 1u, V));
        }
        if (pos + 1 == nsplit) QB += kStop;   // the next residue is the last one
      }
    }
  }
  return acc;
}

// (without variable modifications every mask is 0: not read)
template <bool HASVAR>
__device__ __forceinline__ CandRef cand_ref(const ScoreArgs& A, uint32_t s, uint32_t v, uint32_t nt, uint64_t t0c, uint32_t n_per) {
  CandRef r;
  if (v < nt) {
    const uint64_t c = t0c + v, d = A.cand_desc[c];
    r.row = reinterpret_cast<const uint4*>(A.idx_rows + (d & 0xFFFFFFFFFFull) * 16); r.row_hi = r.row + 2;
    r.len = (uint32_t)(d >> 40) & 0xFF; r.mask = HASVAR ? A.cand_mask[c] : 0ull; r.modw = A.cand_w[c];
  } else {
    const uint64_t j = (uint64_t)s * n_per + (v - nt);
    r.row = reinterpret_cast<const uint4*>(A.dec_rows + j * MD_DECOY_HALF); r.row_hi = reinterpret_cast<const uint4*>(A.dec_rows + (A.dec_slots + j) * MD_DECOY_HALF);
    r.len = A.dec_len[j]; r.mask = HASVAR ? A.dec_mask[j] : 0ull; r.modw = A.dec_w[j];
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


// One warp's share of a tile pass: units of 32 candidates in length-descending order (s_order), fetched from a
// shared counter, so that warps stay busy until the chunk is done and every warp runs candidates of one length.
template <int NCH, bool HASVAR, bool MAPG>
__device__ __forceinline__ void score_units(const ScoreArgs& A, const ScoreConst& C, uint32_t s, uint32_t c0, uint32_t cn, uint32_t nt, uint64_t t0c,
                                            const TableView& V, const LaneTab& L, int64_t* s_score, const uint16_t* s_order, uint32_t* s_unit) {
  const uint32_t lane = threadIdx.x & 31, units = (cn + 31) >> 5;
  for (;;) {
    uint32_t u = 0;
    if (lane == 0) u = atomicAdd(s_unit, 1u);
    u = __shfl_sync(0xffffffffu, u, 0);
    if (u >= units) break;
    const uint32_t slot = u * 32 + lane;
    CandRef cr;
    cr.row = reinterpret_cast<const uint4*>(A.idx_rows); cr.row_hi = cr.row; cr.len = 0; cr.mask = 0; cr.modw = 0;
    uint32_t v = 0;
    if (slot < cn) {
      v = s_order[slot];
      cr = cand_ref<HASVAR>(A, s, c0 + v, nt, t0c, C.n_per);
    }
    const uint32_t maxlen = __reduce_max_sync(0xffffffffu, cr.len);
    const int64_t part = score_one<NCH, HASVAR, MAPG>(cr, maxlen, V, C, L);
    if (slot < cn) s_score[v] += part;
  }
}

// What the prefetch warp leaves for the next spectrum of the CTA (double-buffered with the staged peaks, the block bitmap
// and the sorted candidate order).
struct SpecMeta {
  md_precursor pr;
  uint64_t t0c, pk0;
  uint32_t s, nt, nd, npk, nact;
  uint32_t part, c_lo, c_hi;         // this work item's part of the spectrum: candidates [c_lo, c_hi)
  int32_t hbin;
  uint32_t pre_blocks;               // peaks, bitmap and block numbering are already in shared memory
};

struct FinRecord { md_precursor pr; uint64_t t0c; uint32_t s, nt, nd, pending, ranked, part; };

struct SpecShared {
  SpecMeta meta[2];
  unsigned long long wkey[2][kScoreThreads / 32];
  unsigned long long top[kMaxTopK];
  unsigned long long wtop[2][kScoreThreads / 32][kFastTopK];  // per-warp best keys of a spectrum, descending
  FinRecord fin[2];         // what the finish warp needs to write that spectrum's PSM rows
  long long tacc[8];
  uint32_t hist[64];        // candidates per length (counting sort of a chunk by the whole CTA)
  uint32_t unit;            // next unit of the current tile pass
  uint32_t nact;            // occupied table blocks of the spectrum
  uint32_t x0; int32_t cin; // first bin of the current tile, running sum in front of it
  uint32_t wsum[kScoreThreads / 32];
  uint32_t bits[2][kMapCap / 32], bpre[2][kMapCap / 32];   // occupied-block bitmap and its exclusive popcount prefix
};

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ uint32_t cand_len(const ScoreArgs& A, uint32_t s, uint32_t u, uint32_t nt, uint64_t t0c, uint32_t n_per) {
  return u < nt ? (uint32_t)(A.cand_desc[t0c + u] >> 40) & 0xFF : A.dec_len[(uint64_t)s * n_per + (u - nt)];
}

// Run by ONE warp while the others score: take the CTA's next spectrum from the queue and leave in shared memory what
// its iteration would otherwise wait for -- the precursor and candidate ranges, the binned peaks (cp.async) and the
// occupied-block bitmap with its numbering.
__device__ __noinline__ void prefetch_spectrum(const ScoreArgs& A, SpecShared& sh, uint32_t slot, int32_t* s_bin, int32_t* s_yq) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t vn = 0;
  if (lane == 0) vn = atomicAdd(A.work, 1u);
  vn = __shfl_sync(0xffffffffu, vn, 0);
  const uint32_t sn = vn / A.parts;          // work item -> (spectrum, part); sn >= n_spec ends the CTA
  SpecMeta m;
  m.part = vn % A.parts; m.c_lo = 0; m.c_hi = 0;
  m.s = sn; m.pre_blocks = 0; m.nact = 0; m.nt = 0; m.nd = 0; m.npk = 0; m.hbin = -1; m.t0c = 0; m.pk0 = 0;
  m.pr.mass = 0; m.pr.lo = 0; m.pr.hi = 0; m.pr.charge = 0; m.pr.spectrum_id = 0;
  if (sn < A.n_spec) {
    m.pr = A.prec[sn];
    m.t0c = A.cand_off[sn];
    m.nt = (uint32_t)(A.cand_off[sn + 1] - m.t0c);
    m.nd = A.dec_count ? A.dec_count[sn] : 0;
    m.hbin = A.pk_hbin[sn]; m.npk = A.pk_count[sn]; m.pk0 = A.peak_off[sn];
    const uint32_t NB = m.hbin >= 0 ? (uint32_t)m.hbin + kXcorrOffset + 1 : 0;
    const uint32_t nblk = (NB + kBlk - 1) >> kBlkShift, nwords = (nblk + 31) >> 5;
    const uint32_t ncand = m.nt + m.nd;
    {   // this part's contiguous range of candidate chunks
      const uint32_t nchunks = (ncand + kCandChunk - 1) / kCandChunk, per = (nchunks + A.parts - 1) / A.parts;
      m.c_lo = min(m.part * per * kCandChunk, ncand); m.c_hi = min((m.part + 1) * per * kCandChunk, ncand);
    }
    const bool scored = m.hbin >= 0 && NB <= kMaxBins && ncand <= 0xFFFFFFu;
    if (scored && m.npk <= kPeakCap && nblk <= kMapCap) {
      const int32_t* gb = A.pk_bin + m.pk0; const int32_t* gy = A.pk_yq + m.pk0;
      int32_t* sb = s_bin + slot * kPeakCap; int32_t* sy = s_yq + slot * kPeakCap;
      for (uint32_t i = lane; i < m.npk; i += 32) { cp_async4(sb + i, gb + i); cp_async4(sy + i, gy + i); }
      uint32_t* bits = sh.bits[slot]; uint32_t* bpre = sh.bpre[slot];
      for (uint32_t i = lane; i < nwords; i += 32) bits[i] = 0;
      cp_async_wait_all();
      __syncwarp();
      for (uint32_t p = lane; p < m.npk; p += 32) {     // every bin within 75 of a peak, and the bin behind
        const int32_t b = sb[p];
        const uint32_t b0 = (uint32_t)max(b - kXcorrOffset, 0) >> kBlkShift, b1 = min((uint32_t)(b + kXcorrOffset + 1), NB - 1) >> kBlkShift;
        for (uint32_t k = b0; k <= b1; k++) atomicOr(&bits[k >> 5], 1u << (k & 31));
      }
      __syncwarp();
      uint32_t run = 0;
      for (uint32_t base = 0; base < nwords; base += 32) {   // exclusive popcount prefix over the bitmap words
        const uint32_t i = base + lane;
        const uint32_t c = i < nwords ? (uint32_t)__popc(bits[i]) : 0u;
        uint32_t incl = c;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
        if (i < nwords) bpre[i] = run + incl - c;
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
      m.nact = run; m.pre_blocks = 1;
    }
  }
  __syncwarp();
  if (lane == 0) sh.meta[slot] = m;
}

// Run by ONE warp while the others score the next spectrum: merge the per-warp top-k lists of spectrum fin[slot] (lane r
// ends up with its r-th best key) and write its PSM rows.
__device__ __noinline__ void finish_spectrum(const ScoreArgs& A, const ScoreConst& C, SpecShared& sh, uint32_t slot) {
  const FinRecord f = sh.fin[slot];
  if (!f.pending) return;
  const uint32_t lane = threadIdx.x & 31, K = C.top_k;
  constexpr uint32_t NW = kScoreThreads / 32;
  unsigned long long mine = 0ull;
  if (f.ranked) {
    uint32_t idx = 0;
    for (uint32_t r = 0; r < K; r++) {
      const unsigned long long head = (lane < NW && idx < K) ? sh.wtop[slot][lane][idx] : 0ull;
      const unsigned long long wm = warp_max_u64(head);
      if (wm != 0ull && head == wm) idx++;
      if (lane == r) mine = wm;
    }
  }
  if (A.parts > 1) {
    // leave this part's keys; the last part of the spectrum to arrive merges all of them
    unsigned long long* pt = A.part_top + (size_t)f.s * A.parts * kFastTopK;
    if (lane < kFastTopK) pt[f.part * kFastTopK + lane] = lane < K ? mine : 0ull;
    __threadfence();
    uint32_t arrived = 0;
    if (lane == 0) arrived = atomicAdd(&A.parts_done[f.s], 1u);
    arrived = __shfl_sync(0xffffffffu, arrived, 0);
    if (arrived + 1 != A.parts) return;
    __threadfence();
    static_assert(8 * kFastTopK <= 64, "a lane merges at most two keys of the parts' lists");
    unsigned long long k0 = lane < A.parts * kFastTopK ? __ldcg(pt + lane) : 0ull;
    unsigned long long k1 = lane + 32 < A.parts * kFastTopK ? __ldcg(pt + lane + 32) : 0ull;
    mine = 0ull;
    for (uint32_t r = 0; r < K; r++) {
      const unsigned long long wm = warp_max_u64(k0 > k1 ? k0 : k1);
      if (wm != 0ull) { if (k0 == wm) k0 = 0ull; else if (k1 == wm) k1 = 0ull; }
      if (lane == r) mine = wm;
    }
  }
  if (lane < K) write_psm_row(A, C, f.pr, f.s, lane, mine, f.nt, f.nd, f.t0c);
}

// MODE 0: no spectrum of the batch has more candidates than one chunk (the usual 10-ppm search: the chunk loop runs once,
// known at compile time); 1: general; 2: general, and spectra are divided into parts (ScoreArgs::parts > 1)
template <bool HASVAR, int MODE>
__global__ void __launch_bounds__(kScoreThreads, 1) k_score(const __grid_constant__ ScoreArgs A, const __grid_constant__ ScoreConst C) {
  extern __shared__ __align__(16) int32_t tab[];                     // kTileBins
  int64_t* s_score = reinterpret_cast<int64_t*>(tab + kTileBins);    // kCandChunk
  int32_t* s_bin = reinterpret_cast<int32_t*>(s_score + kCandChunk); // 2 x kPeakCap
  int32_t* s_yq = s_bin + 2 * kPeakCap;                              // 2 x kPeakCap
  uint16_t* s_order = reinterpret_cast<uint16_t*>(s_yq + 2 * kPeakCap);   // kCandChunk: chunk slots in length-descending order
  uint16_t* s_map = s_order + kCandChunk;                            // kMapCap + 2
  __shared__ SpecShared sh;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t NW = kScoreThreads / 32;
  // two warps start every scoring phase with a side job: one fetches the CTA's next spectrum, the other merges the
  // per-warp top-k lists of the previous spectrum and writes its PSM rows -- both off the critical path
  constexpr uint32_t kPrefetchWarp = NW - 1, kFinishWarp = NW - 2;
  const LaneTab L{C.tq[lane], C.tr[lane], C.vq[lane], C.vr[lane]};   // per-letter (q, r) tables in registers: lane = residue code
  uint32_t my_pairs = 0, my_bytes = 0;   // per thread, summed at the end
  long long t_last = A.timing ? clock64() : 0;
  if (tid < 8) sh.tacc[tid] = 0;
#define MD_TICK(k) do { if (A.timing && tid == 0) { const long long now_ = clock64(); sh.tacc[k] += now_ - t_last; t_last = now_; } } while (0)

  if (warp == kPrefetchWarp) prefetch_spectrum(A, sh, 0, s_bin, s_yq);
  if (tid < 2) sh.fin[tid].pending = 0;
  for (uint32_t it = 0;; it++) {
    const uint32_t slot = it & 1;
    __syncthreads();                  // this spectrum's prefetch is visible; the other slot's buffers are free again
    const SpecMeta mt = sh.meta[slot];
    const uint32_t s = mt.s;
    if (s >= A.n_spec) {
      if (warp == kFinishWarp && it > 0) finish_spectrum(A, C, sh, slot ^ 1u);     // the CTA's last spectrum
      break;
    }
    bool side_done = false;           // (helper warps) this iteration's side job is done
    const md_precursor pr = mt.pr;
    const uint64_t t0c = mt.t0c;
    const uint32_t nt = mt.nt, nd = mt.nd;
    const uint32_t ncand = nt + nd;
    const int32_t hbin = mt.hbin;
    const uint32_t npk = mt.npk;
    const uint64_t pk0 = mt.pk0;
    const uint32_t K = C.top_k;
    uint32_t nch = pr.charge > 1 ? pr.charge - 1 : 1;
    if (nch > C.max_frag_charge) nch = C.max_frag_charge;
    if (nch < 1) nch = 1;
    const uint32_t NB = hbin >= 0 ? (uint32_t)hbin + kXcorrOffset + 1 : 0;  // table bins [0, NB)
    const uint32_t nblk = (NB + kBlk - 1) >> kBlkShift;
    bool scored = hbin >= 0;
    const bool mapg = nblk > kMapCap;
    if (NB > kMaxBins || ncand > 0xFFFFFFu || (mapg && (!A.gmap || nblk + 1 > A.gmap_stride))) { if (tid == 0) *A.error = 1; scored = false; }
    uint16_t* map = mapg ? A.gmap + (size_t)blockIdx.x * A.gmap_stride : s_map;
    uint32_t* bits = mapg ? A.gbits + (size_t)blockIdx.x * 2 * (A.gmap_stride / 32 + 1) : sh.bits[slot];
    uint32_t* bpre = mapg ? bits + (A.gmap_stride / 32 + 1) : sh.bpre[slot];
    const uint32_t nwords = (nblk + 31) >> 5;
    const int32_t* pbin = A.pk_bin + pk0; const int32_t* pyq = A.pk_yq + pk0;
    if (npk <= kPeakCap && mt.pre_blocks) { pbin = s_bin + slot * kPeakCap; pyq = s_yq + slot * kPeakCap; }
    for (uint32_t r = tid; r < K; r += kScoreThreads) sh.top[r] = 0ull;

    if (scored && !mt.pre_blocks) {
      // ---- (spectra the prefetch warp left alone: more peaks or table blocks than the shared-memory staging holds)
      for (uint32_t i = tid; i < nwords; i += kScoreThreads) bits[i] = 0;
      __syncthreads();
      for (uint32_t p = tid; p < npk; p += kScoreThreads) {   // every bin within 75 of a peak, and the bin behind
        const int32_t b = pbin[p];
        const uint32_t b0 = (uint32_t)max(b - kXcorrOffset, 0) >> kBlkShift, b1 = min((uint32_t)(b + kXcorrOffset + 1), NB - 1) >> kBlkShift;
        for (uint32_t k = b0; k <= b1; k++) atomicOr(&bits[k >> 5], 1u << (k & 31));
      }
      __syncthreads();
      if (warp == 0) {   // exclusive popcount prefix over the bitmap words
        uint32_t run = 0;
        for (uint32_t base = 0; base < nwords; base += 32) {
          const uint32_t i = base + lane;
          const uint32_t c = i < nwords ? (uint32_t)__popc(bits[i]) : 0u;
          uint32_t incl = c;
          for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
          if (i < nwords) bpre[i] = run + incl - c;
          run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) sh.nact = run;
      }
      __syncthreads();
    }
    const uint32_t nact = scored ? (mt.pre_blocks ? mt.nact : sh.nact) : 0;
    MD_TICK(0);

    const bool fast = K <= kFastTopK;   // per-warp running top-k lists, merged by the finish warp while the CTA scores the next spectrum
    // the table of one tile of occupied blocks [cb0, cb0 + cbn) (block-uniform: every thread calls it)
    auto build_tile = [&](const uint32_t cb0, const uint32_t cbn) {
      // (1) block map of the tile (0 = the all-zero block), zero the occupied blocks
      for (uint32_t k = tid; k <= nblk; k += kScoreThreads) {
        uint32_t m = 0;
        if (k < nblk) {
          const uint32_t wd = bits[k >> 5];
          if ((wd >> (k & 31)) & 1u) {
            const uint32_t c = bpre[k >> 5] + (uint32_t)__popc(wd & ((1u << (k & 31)) - 1u));
            if (c >= cb0 && c < cb0 + cbn) m = c - cb0 + 1u;
          }
        }
        map[k] = (uint16_t)(m ? (m << kBlkShift) | (kBlk - 1u) : 0u);
        if (m == 1u) sh.x0 = k << kBlkShift;                              // first bin of the tile
      }
      {
        uint4* z = reinterpret_cast<uint4*>(tab);
        const uint32_t n4 = (cbn + 1u) * (kBlk / 4);
        for (uint32_t i = tid; i < n4; i += kScoreThreads) z[i] = make_uint4(0, 0, 0, 0);
        if (tid == 0) { sh.unit = 0; sh.cin = 0; }
      }
      __syncthreads();
      MD_TICK(2);
      // (2) T as a difference array: every peak adds +y at the first bin of its window (bin-75), -y behind its last
      //     (bin+76), and its own 151*y spike as -151*y at bin / +151*y at bin+1, so that the running sum over the
      //     bins is R[b] = S[b] - 151*y[b] = -T[b].  All four bins lie in occupied blocks (the marking covers
      //     [bin-75, bin+76]), R is 0 at the end of every run of occupied blocks, and the compressed blocks are in
      //     bin order: one inclusive scan straight over the compressed table yields -T.  Integer adds commute, so
      //     the shared-memory atomics keep the table exact.  A later tile starts from the sum of the differences
      //     at the bins in front of it (sh.cin).
      {
        const uint32_t X0 = sh.x0;
        int32_t cin = 0;
        for (uint32_t p = tid; p < npk; p += kScoreThreads) {
          const int32_t bp = pbin[p], y = pyq[p];
          const uint32_t xa = (uint32_t)max(bp - kXcorrOffset, 0), xe = (uint32_t)(bp + kXcorrOffset + 1), xs = (uint32_t)bp;
          const int32_t spike = 151 * y;
          uint32_t m;
          m = (uint32_t)map[xa >> kBlkShift] >> kBlkShift; if (m) atomicAdd(&tab[(m << kBlkShift) + (xa & (kBlk - 1u))], y);
          // (xe == NB for the last peak: the bins of the last block behind the table's end must read 0 too)
          if (xe < (nblk << kBlkShift)) { m = (uint32_t)map[xe >> kBlkShift] >> kBlkShift; if (m) atomicAdd(&tab[(m << kBlkShift) + (xe & (kBlk - 1u))], -y); }
          m = (uint32_t)map[xs >> kBlkShift] >> kBlkShift; if (m) atomicAdd(&tab[(m << kBlkShift) + (xs & (kBlk - 1u))], -spike);
          m = (uint32_t)map[(xs + 1u) >> kBlkShift] >> kBlkShift; if (m) atomicAdd(&tab[(m << kBlkShift) + ((xs + 1u) & (kBlk - 1u))], spike);
          if (cb0 > 0) cin += (xa < X0 ? y : 0) - (xe < X0 ? y : 0) - (xs < X0 ? spike : 0) + (xs + 1u < X0 ? spike : 0);
        }
        if (cb0 > 0 && cin != 0) atomicAdd(&sh.cin, cin);
      }
      __syncthreads();
      MD_TICK(3);
      // (3) the scan, in place, negated: every thread owns a contiguous range of 16-byte words (an odd number of them:
      //     the 128-bit accesses of a quarter warp then fall into distinct banks); pass A adds up the ranges, a warp scan
      //     and the per-warp sums give every thread its start value, pass B rescans the range
      {
        uint4* t4 = reinterpret_cast<uint4*>(tab) + kBlk / 4;            // entry 0 = first occupied block
        const uint32_t N4 = cbn * (kBlk / 4);
        const uint32_t c4 = ((N4 + kScoreThreads - 1) / kScoreThreads) | 1u;
        const uint32_t s0 = min(tid * c4, N4), s1 = min(s0 + c4, N4);
        uint32_t part = 0;
        for (uint32_t i = s0; i < s1; i++) { const uint4 v = t4[i]; part += v.x + v.y + v.z + v.w; }
        uint32_t incl = part;
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += t; }
        if (lane == 31) sh.wsum[warp] = incl;
        __syncthreads();
        uint32_t run = __reduce_add_sync(0xffffffffu, lane < warp ? sh.wsum[lane] : 0u) + (uint32_t)sh.cin + incl - part;
        for (uint32_t i = s0; i < s1; i++) {
          uint4 v = t4[i];
          v.x += run; v.y += v.x; v.z += v.y; v.w += v.z; run = v.w;
          t4[i] = make_uint4(0u - v.x, 0u - v.y, 0u - v.z, 0u - v.w);
        }
      }
      __syncthreads();
      MD_TICK(6);
    };
    const bool one_tile = nact <= kTileBlocks;   // the usual case: the table is built once and serves every chunk of candidates
    const uint32_t c_lo = MODE == 2 ? mt.c_lo : 0u, c_hi = MODE == 2 ? mt.c_hi : ncand;     // this work item's candidates (the whole spectrum unless it was split)
    for (uint32_t c0 = c_lo; MODE == 0 ? c0 == 0 : (c0 < c_hi || c0 == c_lo); c0 += kCandChunk) {
      const uint32_t cn = c0 < c_hi ? min(kCandChunk, c_hi - c0) : 0u;
      for (uint32_t v = tid; v < cn; v += kScoreThreads) s_score[v] = 0;
      // counting sort of the chunk by peptide length, longest first
      {
        if (tid < 64) sh.hist[tid] = 0;
        __syncthreads();
        static_assert(kCandChunk <= 2 * kScoreThreads, "a thread sorts at most two candidates of a chunk");
        uint32_t len2[2] = {0, 0};
        if (scored) {
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const uint32_t v = tid + h * kScoreThreads;
            if (v < cn) {
              const uint32_t len = cand_len(A, s, c0 + v, nt, t0c, C.n_per);
              len2[h] = len;
              atomicAdd(&sh.hist[63u - min(len, 63u)], 1u);
              my_pairs++; my_bytes += 14 + len;
            }
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
#pragma unroll
          for (int h = 0; h < 2; h++) {
            const uint32_t v = tid + h * kScoreThreads;
            if (v < cn) s_order[atomicAdd(&sh.hist[63u - min(len2[h], 63u)], 1u)] = (uint16_t)v;
          }
        }
      }
      // (visible to the scoring warps after the table-build barriers)
      MD_TICK(1);
      if (scored && cn) {
        for (uint32_t cb0 = 0; cb0 < nact; cb0 += kTileBlocks) {       // tiles of occupied blocks (usually one)
          const uint32_t cbn = min(kTileBlocks, nact - cb0);
          if (!one_tile || c0 == c_lo) build_tile(cb0, cbn);
          else { if (tid == 0) sh.unit = 0; __syncthreads(); }               // same table; the chunk's order and zeroed scores are visible
          // (4) score the chunk against the tile; one warp first fetches the CTA's next spectrum
          if (!side_done) {
            if (warp == kPrefetchWarp) prefetch_spectrum(A, sh, slot ^ 1u, s_bin, s_yq);
            if (warp == kFinishWarp && it > 0) finish_spectrum(A, C, sh, slot ^ 1u);
            side_done = true;
          }
          {
            TableView V;
            V.tab_s = (uint32_t)__cvta_generic_to_shared(tab); V.map_s = (uint32_t)__cvta_generic_to_shared(s_map); V.nblk = nblk; V.gmap = map;
            switch (nch * 2 + (mapg ? 1 : 0)) {
              case 2: score_units<1, HASVAR, false>(A, C, s, c0, cn, nt, t0c, V, L, s_score, s_order, &sh.unit); break;
              case 3: score_units<1, HASVAR, true>(A, C, s, c0, cn, nt, t0c, V, L, s_score, s_order, &sh.unit); break;
              case 4: score_units<2, HASVAR, false>(A, C, s, c0, cn, nt, t0c, V, L, s_score, s_order, &sh.unit); break;
              case 5: score_units<2, HASVAR, true>(A, C, s, c0, cn, nt, t0c, V, L, s_score, s_order, &sh.unit); break;
              case 6: score_units<3, HASVAR, false>(A, C, s, c0, cn, nt, t0c, V, L, s_score, s_order, &sh.unit); break;
              default: score_units<3, HASVAR, true>(A, C, s, c0, cn, nt, t0c, V, L, s_score, s_order, &sh.unit); break;
            }
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
      // ---- top-k, K <= 8: every warp keeps the K best keys it has seen so far (lane r holds its r-th best across chunks)
      if (fast && scored && K) {   // (also without a candidate: the lists must not keep an earlier spectrum's keys)
        unsigned long long k0 = tid < cn ? psm_key(s_score[tid], c0 + tid) : 0ull;
        unsigned long long k1 = tid + kScoreThreads < cn ? psm_key(s_score[tid + kScoreThreads], c0 + tid + kScoreThreads) : 0ull;
        if (MODE == 0 || c_hi - c_lo <= kCandChunk) {      // one chunk (the usual case): nothing to carry over
          for (uint32_t r = 0; r < K; r++) {
            const unsigned long long wm = warp_max_u64(k0 > k1 ? k0 : k1);
            if (wm != 0ull) { if (k0 == wm) k0 = 0ull; else if (k1 == wm) k1 = 0ull; }
            if (lane == 0) sh.wtop[slot][warp][r] = wm;
          }
        } else {
          unsigned long long carry = (c0 > c_lo && lane < K) ? sh.wtop[slot][warp][lane] : 0ull, mine = 0ull;
          for (uint32_t r = 0; r < K; r++) {
            const unsigned long long b01 = k0 > k1 ? k0 : k1;
            const unsigned long long wm = warp_max_u64(b01 > carry ? b01 : carry);
            if (wm != 0ull) { if (k0 == wm) k0 = 0ull; else if (k1 == wm) k1 = 0ull; else if (carry == wm) carry = 0ull; }
            if (lane == r) mine = wm;
          }
          __syncwarp();
          if (lane < K) sh.wtop[slot][warp][lane] = mine;
        }
      }
      // ---- top-k of (this chunk's candidates) U (best of the earlier chunks), K > 8: K rounds of block-wide max
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
      }
    }
    if (!side_done) {   // (nothing was scored)
      if (warp == kPrefetchWarp) prefetch_spectrum(A, sh, slot ^ 1u, s_bin, s_yq);
      if (warp == kFinishWarp && it > 0) finish_spectrum(A, C, sh, slot ^ 1u);
    }
    // ---- PSM rows
    if (fast) {
      // the per-warp top-k lists are complete; the finish warp merges them and writes the rows during the next scoring phase
      if (tid == 0) {
        FinRecord f; f.pr = pr; f.t0c = t0c; f.s = s; f.nt = nt; f.nd = nd; f.pending = 1; f.ranked = (scored && K) ? 1u : 0u; f.part = mt.part;
        sh.fin[slot] = f;
      }
      MD_TICK(5);
    } else {
      for (uint32_t r = tid; r < K; r += kScoreThreads) write_psm_row(A, C, pr, s, r, scored ? sh.top[r] : 0ull, nt, nd, t0c);
      if (tid == 0) sh.fin[slot].pending = 0;
      __syncthreads();
      MD_TICK(5);
    }
  }
  if (A.timing && tid == 0) for (int k = 0; k < 8; k++) atomicAdd(&A.timing[k], (unsigned long long)sh.tacc[k]);
  unsigned long long wp = my_pairs, wb = my_bytes;
  for (int o = 16; o; o >>= 1) { wp += __shfl_xor_sync(0xffffffffu, wp, o); wb += __shfl_xor_sync(0xffffffffu, wb, o); }
  if (lane == 0 && wp) { atomicAdd(&A.stat64[0], wp); atomicAdd(&A.stat64[1], wb); }
}

__global__ void k_max_i32(const int32_t* __restrict__ v, uint32_t n, int32_t* __restrict__ out) {
  int32_t m = INT32_MIN;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, v[i]);
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
}

// most candidates (targets + decoys) of one spectrum of the batch
__global__ void k_max_candidates(const uint64_t* __restrict__ cand_off, const uint32_t* __restrict__ dec_count, uint32_t n, int32_t* __restrict__ out) {
  int32_t m = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint64_t c = cand_off[i + 1] - cand_off[i] + (dec_count ? dec_count[i] : 0u);
    m = max(m, (int32_t)min(c, (uint64_t)0x7FFFFFFF));
  }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0) atomicMax(out, m);
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
  W.pk_bin.need(n_peaks + 1); W.pk_yq.need(n_peaks + 1); W.pk_count.need(n + 1); W.pk_hbin.need(n + 1);
  DevBuf<int>& d_flag = W.t_unsorted; d_flag.need(4);
  MD_CUDA(cudaMemsetAsync(d_flag.p, 0, 4 * sizeof(int), ctx->stream));
  MD_CUDA(cudaMemsetAsync(W.pk_yq.p, 0, (n_peaks + 1) * sizeof(int32_t), ctx->stream));
  MD_LAUNCH(ctx, k_bin_spectra, blocks((uint64_t)n * 32, 128), 128, 0, S.peak_off, S.peak_mz, S.peak_int, W.prec.p, n, w, p.min_peaks, W.pk_bin.p, W.pk_yq.p,
            W.pk_count.p, W.pk_hbin.p, d_flag.p);
  // the largest table of the batch decides whether the block maps fit shared memory
  MD_LAUNCH(ctx, k_max_i32, std::min<uint32_t>(blocks(n), 64), 256, 0, W.pk_hbin.p, n, d_flag.p + 2);
  MD_LAUNCH(ctx, k_max_candidates, std::min<uint32_t>(blocks(n), 64), 256, 0, W.cand_off.p, n_per ? W.dec_count.p : nullptr, n, d_flag.p + 3);
  int h_pre[4] = {0, 0, 0, 0};
  MD_CUDA(cudaMemcpyAsync(h_pre, d_flag.p, sizeof(h_pre), cudaMemcpyDeviceToHost, ctx->stream));
  // ---- K4
  ScoreConst C;
  memset(&C, 0, sizeof(C));
  C.w = (uint32_t)w; C.rcp_w = 1.0f / (float)w; C.max_frag_charge = mfc; C.top_k = p.top_k; C.n_per = n_per;
  split_qr(MD_PROTON_UDA, C.w, &C.qp, &C.rp);
  split_qr(2 * MD_PROTON_UDA, C.w, &C.q2p, &C.r2p);
  bool has_var = false;
  for (int c = 0; c < 32; c++) {
    int64_t m = c < MD_NCODES ? ctx->mods.mass[c] : 0;
    int64_t f = (c < MD_NCODES && ctx->mods.has_fix[c]) ? ctx->mods.fix[c] : 0;
    int64_t v = (c < MD_NCODES && ctx->mods.has_var[c]) ? ctx->mods.var[c] : 0;
    split_qr(m + f, C.w, &C.tq[c], &C.tr[c]);
    split_qr(m + f + v, C.w, &C.vq[c], &C.vr[c]);
    MD_REQUIRE(C.tq[c] < (1u << 22) && C.vq[c] < (1u << 22), MD_ERR_UNSUPPORTED, "fragment_tolerance too small for this residue mass");
    if (c < MD_NCODES && ctx->mods.has_var[c]) has_var = true;
  }
  uint64_t n_targets = 0;
  MD_CUDA(cudaMemcpyAsync(&n_targets, W.cand_off.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  MD_CUDA(cudaStreamSynchronize(ctx->stream));
  MD_REQUIRE(!h_pre[0], MD_ERR_INVALID, "spectra: peaks of a spectrum must be sorted by m/z");
  if (want_all) { W.tscore.need(n_targets + 1); W.dscore.need((size_t)n * n_per + 1); }
  // Few spectra with very many candidates each (open searches) would leave SMs idle and make one CTA walk hundreds of
  // chunks: split every spectrum into `parts` work items (contiguous ranges of candidate chunks; each builds the table
  // itself, which is cheap beside the chunks), about two items per SM.  MD_SCORE_SPLIT_MIN = candidates per spectrum
  // (batch average) from which that is done (tests lower it).
  uint32_t parts = 1;
  {
    uint64_t n_cand = (uint64_t)n * n_per, split_min = 16ull * kCandChunk;
    n_cand += n_targets;
    if (const char* env = getenv("MD_SCORE_SPLIT_MIN")) split_min = std::max<long long>(1, atoll(env));
    if (n < (uint32_t)ctx->n_sm && p.top_k <= kFastTopK && n_cand / n >= split_min)
      parts = std::min<uint32_t>(std::min<uint32_t>(8u, (2u * (uint32_t)ctx->n_sm + n - 1) / n), (uint32_t)((n_cand / n + kCandChunk - 1) / kCandChunk));
    parts = std::max(parts, 1u);
  }
  const uint32_t grid = std::min<uint32_t>(n * parts, (uint32_t)ctx->n_sm);
  const uint32_t max_nblk = h_pre[2] >= 0 ? ((uint32_t)h_pre[2] + kXcorrOffset + 1 + kBlk - 1) / kBlk : 0;
  uint32_t gstride = 0;
  if (max_nblk > kMapCap && max_nblk <= (kMaxBins >> kBlkShift)) {   // block maps of this batch do not fit shared memory
    gstride = (max_nblk + 64) & ~31u;
    W.gmap.need((size_t)grid * gstride); W.gbits.need((size_t)grid * 2 * (gstride / 32 + 1));
  }
  DevBuf<uint32_t>& work = W.counters; work.need(4);
  MD_CUDA(cudaMemsetAsync(work.p, 0, 4 * sizeof(uint32_t), ctx->stream));
  W.stat64.need(16);
  MD_CUDA(cudaMemsetAsync(W.stat64.p, 0, 16 * sizeof(unsigned long long), ctx->stream));
  const bool timing = getenv("MD_SCORE_TIMING") != nullptr;
  ScoreArgs A;
  A.prec = W.prec.p; A.n_spec = n; A.peak_off = S.peak_off; A.pk_bin = W.pk_bin.p; A.pk_yq = W.pk_yq.p; A.pk_count = W.pk_count.p; A.pk_hbin = W.pk_hbin.p;
  A.cand_off = W.cand_off.p; A.cand_desc = W.cand_desc.p; A.cand_mask = W.cand_mask.p; A.cand_w = W.cand_w.p; A.cand_pep = W.cand_pep.p;
  A.idx_rows = ctx->index.rows.p;
  A.dec_slots = (uint64_t)n * n_per; A.dec_rows = W.dec_rows.p; A.dec_len = W.dec_len.p; A.dec_mask = W.dec_mask.p; A.dec_w = W.dec_w.p; A.dec_count = n_per ? W.dec_count.p : nullptr;
  A.tscore = want_all ? W.tscore.p : nullptr; A.dscore = want_all ? W.dscore.p : nullptr; A.psm = psm_dev; A.work = work.p; A.stat64 = W.stat64.p;
  A.error = d_flag.p + 1; A.timing = timing ? W.stat64.p + 8 : nullptr;
  A.gmap = gstride ? W.gmap.p : nullptr; A.gbits = gstride ? W.gbits.p : nullptr; A.gmap_stride = gstride;
  A.parts = parts; A.part_top = nullptr; A.parts_done = nullptr;
  if (parts > 1) {
    W.part_top.need((size_t)n * parts * kFastTopK); W.parts_done.need(n);
    MD_CUDA(cudaMemsetAsync(W.parts_done.p, 0, n * sizeof(uint32_t), ctx->stream));
    A.part_top = W.part_top.p; A.parts_done = W.parts_done.p;
  }
  const size_t smem = (size_t)kTileBins * sizeof(int32_t) + (size_t)kCandChunk * sizeof(int64_t) + 2 * (size_t)kPeakCap * 8 + (size_t)kCandChunk * 2 +
                      ((size_t)kMapCap + 2) * 2;
  MD_CUDA(cudaEventRecord(ctx->ev[6], ctx->stream));
  auto launch = [&](auto kernel) {
    MD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    MD_LAUNCH(ctx, kernel, grid, kScoreThreads, smem, A, C);
  };
  const bool single = parts == 1 && h_pre[3] <= (int)kCandChunk;     // every spectrum fits one chunk of candidates
  if (has_var) { if (parts > 1) launch(k_score<true, 2>); else if (single) launch(k_score<true, 0>); else launch(k_score<true, 1>); }
  else { if (parts > 1) launch(k_score<false, 2>); else if (single) launch(k_score<false, 0>); else launch(k_score<false, 1>); }
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
    double tot = 0; for (int k = 0; k < 7; k++) tot += (double)t[k];
    static const char* names[7] = {"stage+blocks", "sort", "map+zero", "differences", "score", "topk", "scan"};
    fprintf(stderr, "[md_score_timing] grid=%u", grid);
    for (int k = 0; k < 7; k++) fprintf(stderr, " %s=%.1f%%", names[k], 100.0 * (double)t[k] / tot);
    fprintf(stderr, " cycles/CTA=%.0f\n", tot / grid);
  }
  MD_REQUIRE(!h_flag[1], MD_ERR_UNSUPPORTED, "a spectrum needs more than 2^26 fragment bins or has more than 2^24 candidates");
}
