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
// Kernel design (B200): one persistent CTA per SM; per spectrum the CTA expands the sparse binned spectrum into a
// dense int32 table tile in shared memory (zero + shared-memory atomic scatter of +-75-bin windows), then every
// thread scores one candidate at a time: the row (residue codes, 16-byte padded) comes in with 128-bit coalesced
// loads, the per-letter (quotient, remainder) mass table lives in registers and is read with warp shuffles, fragment
// bins follow from a division-free running (q, r) sum, and each fragment costs exactly one shared-memory gather.
// Spectra wider than one tile are processed tile by tile; partial scores are kept in HBM (L2-resident).
#include "cubx.cuh"

namespace {

inline uint32_t blocks(uint64_t n, uint32_t bs = 256) { return (uint32_t)((n + bs - 1) / bs); }

constexpr int kXcorrOffset = 75;
constexpr int kScoreThreads = 512;
constexpr uint32_t kTileBins = 49152;  // 192 KiB of int32 per CTA

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
  uint32_t qp, rp;            // proton  = qp*w + rp
  uint32_t q2p, r2p;          // 2*proton
  uint32_t tq[32], tr[32];    // per residue code: (mass + fixed delta) = tq*w + tr
  uint32_t vq[32], vr[32];    // per residue code: (mass + variable delta)
  uint32_t max_frag_charge;
  uint32_t top_k, n_per;
};

struct ScoreArgs {
  const md_precursor* prec; uint32_t n_spec;
  const uint64_t* peak_off; const int32_t* pk_bin; const int32_t* pk_yq; const uint32_t* pk_count; const int32_t* pk_hbin;
  const uint64_t* cand_off; const uint64_t* cand_desc; const uint64_t* cand_mask; const int64_t* cand_w; const uint32_t* cand_pep;
  const uint8_t* idx_rows;
  const uint8_t* dec_rows; const uint8_t* dec_len; const uint64_t* dec_mask; const int64_t* dec_w; const uint32_t* dec_count;
  int64_t* tscore; int64_t* dscore;
  md_psm* psm;
  uint32_t* work;
  unsigned long long* stat64;  // [0] pairs scored, [1] algorithmic bytes (14 + len per pair)
};

__device__ __forceinline__ uint32_t div3(uint32_t x) { return __umulhi(x, 0xAAAAAAABu) >> 1; }

// partial raw score of one candidate against the table tile [t0, t0+tn)
template <int NCH, bool HASVAR>
__device__ __forceinline__ int64_t score_one(const uint4* __restrict__ row, uint32_t len, uint64_t mask, int64_t modw, uint32_t maxlen,
                                             const int32_t* __restrict__ tab, uint32_t t0, uint32_t tn, const ScoreConst& C, uint32_t lq, uint32_t lr,
                                             uint32_t lvq, uint32_t lvr) {
  const uint32_t w = C.w;
  // T_c = modw + (c+1)*proton  ->  (Qt, Rt)
  uint64_t T1 = (uint64_t)modw + 2ull * MD_PROTON_UDA;
  uint32_t Qt1 = (uint32_t)(T1 / w), Rt1 = (uint32_t)(T1 - (uint64_t)Qt1 * w);
  uint32_t Rt2 = Rt1 + C.rp, Qt2 = Qt1 + C.qp; if (Rt2 >= w) { Rt2 -= w; Qt2++; }
  uint32_t Rt3 = Rt2 + C.rp, Qt3 = Qt2 + C.qp; if (Rt3 >= w) { Rt3 -= w; Qt3++; }
  uint32_t Q1 = C.qp, R1 = C.rp;  // X1 = B_k + proton
  int64_t acc = 0;
  const uint32_t nsplit = len > 0 ? len - 1 : 0;            // residues 0..len-2 are followed by a split
  const uint32_t nchunk = maxlen > 1 ? (maxlen - 1 + 15) >> 4 : 0;  // warp-uniform
  for (uint32_t c = 0; c < nchunk; c++) {
    const uint4 v = __ldg(row + c);
    const uint32_t words[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 16; j++) {
      const uint32_t i = c * 16 + j;
      const uint32_t code = (words[j >> 2] >> (8 * (j & 3))) & 31u;
      uint32_t q = __shfl_sync(0xffffffffu, lq, code), r = __shfl_sync(0xffffffffu, lr, code);
      if (HASVAR) {
        uint32_t q2 = __shfl_sync(0xffffffffu, lvq, code), r2 = __shfl_sync(0xffffffffu, lvr, code);
        if ((mask >> i) & 1) { q = q2; r = r2; }
      }
      const bool on = i < nsplit;
      Q1 += q; R1 += r; if (R1 >= w) { R1 -= w; Q1++; }
      {  // fragment charge 1
        uint32_t bb = Q1 + 1 - t0;
        uint32_t yb = Qt1 - Q1 - (Rt1 < R1 ? 1u : 0u) + 1 - t0;
        if (on && bb < tn) acc += tab[bb];
        if (on && yb < tn) acc += tab[yb];
      }
      if (NCH >= 2) {
        uint32_t r2 = R1 + C.rp, q2 = Q1 + C.qp + (r2 >= w ? 1u : 0u);
        uint32_t bb = (q2 >> 1) + 1 - t0;
        uint32_t yb = ((Qt2 - Q1 - (Rt2 < R1 ? 1u : 0u)) >> 1) + 1 - t0;
        if (on && bb < tn) acc += tab[bb];
        if (on && yb < tn) acc += tab[yb];
      }
      if (NCH >= 3) {
        uint32_t r3 = R1 + C.r2p, q3 = Q1 + C.q2p + (r3 >= w ? 1u : 0u);
        uint32_t bb = div3(q3) + 1 - t0;
        uint32_t yb = div3(Qt3 - Q1 - (Rt3 < R1 ? 1u : 0u)) + 1 - t0;
        if (on && bb < tn) acc += tab[bb];
        if (on && yb < tn) acc += tab[yb];
      }
    }
  }
  return acc;
}

struct CandRef { const uint4* row; uint32_t len; uint64_t mask; int64_t modw; int64_t* score; };

__device__ __forceinline__ CandRef cand_ref(const ScoreArgs& A, uint32_t s, uint32_t v, uint32_t nt, uint64_t t0c, uint32_t n_per) {
  CandRef r;
  if (v < nt) {
    const uint64_t c = t0c + v, d = A.cand_desc[c];
    r.row = reinterpret_cast<const uint4*>(A.idx_rows + (d & 0xFFFFFFFFFFull) * 16);
    r.len = (uint32_t)(d >> 40) & 0xFF; r.mask = A.cand_mask[c]; r.modw = A.cand_w[c]; r.score = A.tscore + c;
  } else {
    const uint64_t j = (uint64_t)s * n_per + (v - nt);
    r.row = reinterpret_cast<const uint4*>(A.dec_rows + j * MD_DECOY_ROW);
    r.len = A.dec_len[j]; r.mask = A.dec_mask[j]; r.modw = A.dec_w[j]; r.score = A.dscore + j;
  }
  return r;
}

template <int NCH, bool HASVAR>
__device__ void score_tile(const ScoreArgs& A, const ScoreConst& C, uint32_t s, uint32_t nt, uint32_t nd, uint64_t t0c, const int32_t* tab, uint32_t t0,
                           uint32_t tn, bool first, uint32_t lq, uint32_t lr, uint32_t lvq, uint32_t lvr) {
  const uint32_t ncand = nt + nd;
  const uint32_t rounds = (ncand + kScoreThreads - 1) / kScoreThreads;
  for (uint32_t it = 0; it < rounds; it++) {
    const uint32_t v = it * kScoreThreads + threadIdx.x;
    const bool live = v < ncand;
    CandRef r; r.row = reinterpret_cast<const uint4*>(A.idx_rows); r.len = 0; r.mask = 0; r.modw = 0; r.score = nullptr;
    if (live) r = cand_ref(A, s, v, nt, t0c, C.n_per);
    const uint32_t maxlen = __reduce_max_sync(0xffffffffu, r.len);
    int64_t part = score_one<NCH, HASVAR>(r.row, r.len, r.mask, r.modw, maxlen, tab, t0, tn, C, lq, lr, lvq, lvr);
    if (live) *r.score = first ? part : *r.score + part;
  }
}

// ordering of PSMs: raw score descending, candidate ordinal ascending
__device__ __forceinline__ bool psm_better(int64_t sa, uint32_t va, int64_t sb, uint32_t vb) { return sa > sb || (sa == sb && va < vb); }

template <bool HASVAR>
__global__ void __launch_bounds__(kScoreThreads, 1) k_score(const __grid_constant__ ScoreArgs A, const __grid_constant__ ScoreConst C) {
  extern __shared__ __align__(16) int32_t tab[];
  __shared__ uint32_t s_work;
  __shared__ int64_t s_rs[kScoreThreads / 32];
  __shared__ uint32_t s_rv[kScoreThreads / 32];
  __shared__ int64_t s_best_s; __shared__ uint32_t s_best_v;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // per-letter (q, r) tables in registers: lane = residue code
  const uint32_t lq = C.tq[lane], lr = C.tr[lane], lvq = C.vq[lane], lvr = C.vr[lane];
  unsigned long long my_pairs = 0, my_bytes = 0;

  for (;;) {
    if (threadIdx.x == 0) s_work = atomicAdd(A.work, 1u);
    __syncthreads();
    const uint32_t s = s_work;
    __syncthreads();
    if (s >= A.n_spec) break;
    const md_precursor pr = A.prec[s];
    const uint64_t t0c = A.cand_off[s];
    const uint32_t nt = (uint32_t)(A.cand_off[s + 1] - t0c);
    const uint32_t nd = A.dec_count ? A.dec_count[s] : 0;
    const uint32_t ncand = nt + nd;
    const int32_t hbin = A.pk_hbin[s];
    const bool scored = hbin >= 0;
    uint32_t nch = pr.charge > 1 ? pr.charge - 1 : 1;
    if (nch > C.max_frag_charge) nch = C.max_frag_charge;
    if (nch < 1) nch = 1;

    if (!scored || ncand == 0) {
      for (uint32_t v = threadIdx.x; v < ncand; v += kScoreThreads) {
        CandRef r = cand_ref(A, s, v, nt, t0c, C.n_per);
        *r.score = 0;
      }
    } else {
      for (uint32_t v = threadIdx.x; v < ncand; v += kScoreThreads) {
        const uint32_t len = v < nt ? (uint32_t)(A.cand_desc[t0c + v] >> 40) & 0xFF : A.dec_len[(uint64_t)s * C.n_per + (v - nt)];
        my_pairs++; my_bytes += 14 + len;
      }
      const uint32_t NB = (uint32_t)hbin + kXcorrOffset + 1;  // table bins [0, NB)
      const uint64_t pk0 = A.peak_off[s]; const uint32_t npk = A.pk_count[s];
      bool first = true;
      for (uint32_t t0 = 0; t0 < NB; t0 += kTileBins) {
        const uint32_t tn = min(kTileBins, NB - t0);
        // peaks that touch this tile: bin in [t0 - 75, t0 + tn + 75)
        uint32_t pa, pb;
        {
          const int32_t lo_bin = (int32_t)t0 - kXcorrOffset, hi_bin = (int32_t)(t0 + tn) + kXcorrOffset;
          uint32_t l = 0, h = npk;
          while (l < h) { uint32_t m = (l + h) >> 1; if (A.pk_bin[pk0 + m] < lo_bin) l = m + 1; else h = m; }
          pa = l; h = npk;
          while (l < h) { uint32_t m = (l + h) >> 1; if (A.pk_bin[pk0 + m] < hi_bin) l = m + 1; else h = m; }
          pb = l;
        }
        if (pa == pb) {  // an all-zero tile adds nothing
          if (first) {
            for (uint32_t v = threadIdx.x; v < ncand; v += kScoreThreads) { CandRef r = cand_ref(A, s, v, nt, t0c, C.n_per); *r.score = 0; }
            first = false;
          }
          continue;
        }
        // zero the tile
        {
          uint4* z = reinterpret_cast<uint4*>(tab);
          const uint32_t n4 = (tn + 3) >> 2;
          for (uint32_t i = threadIdx.x; i < n4; i += kScoreThreads) z[i] = make_uint4(0, 0, 0, 0);
        }
        __syncthreads();
        // expand: T[b] += 150*y at the peak bin, -= y at the 150 neighbours
        for (uint32_t p = pa + warp; p < pb; p += kScoreThreads / 32) {
          const int32_t bin = A.pk_bin[pk0 + p], yq = A.pk_yq[pk0 + p];
          for (int o = (int)lane - kXcorrOffset; o <= kXcorrOffset; o += 32) {
            const uint32_t idx = (uint32_t)(bin + o) - t0;
            if (idx < tn) atomicAdd(&tab[idx], o == 0 ? 150 * yq : -yq);
          }
        }
        __syncthreads();
        switch (nch) {
          case 1: score_tile<1, HASVAR>(A, C, s, nt, nd, t0c, tab, t0, tn, first, lq, lr, lvq, lvr); break;
          case 2: score_tile<2, HASVAR>(A, C, s, nt, nd, t0c, tab, t0, tn, first, lq, lr, lvq, lvr); break;
          default: score_tile<3, HASVAR>(A, C, s, nt, nd, t0c, tab, t0, tn, first, lq, lr, lvq, lvr); break;
        }
        first = false;
        __syncthreads();
      }
    }
    __syncthreads();
    // ---- per-spectrum top-k (PSM rows) ----
    const uint32_t K = C.top_k;
    int64_t prev_s = INT64_MAX; uint32_t prev_v = 0; bool have_prev = false;
    for (uint32_t r = 0; r < K; r++) {
      int64_t bs = INT64_MIN; uint32_t bv = 0xFFFFFFFFu;
      if (scored) {
        for (uint32_t v = threadIdx.x; v < ncand; v += kScoreThreads) {
          const int64_t sc = v < nt ? A.tscore[t0c + v] : A.dscore[(uint64_t)s * C.n_per + (v - nt)];
          if (have_prev && !psm_better(prev_s, prev_v, sc, v)) continue;  // already reported
          if (bv == 0xFFFFFFFFu || psm_better(sc, v, bs, bv)) { bs = sc; bv = v; }
        }
      }
      for (int o = 16; o; o >>= 1) {
        int64_t os = __shfl_xor_sync(0xffffffffu, bs, o); uint32_t ov = __shfl_xor_sync(0xffffffffu, bv, o);
        if (ov != 0xFFFFFFFFu && (bv == 0xFFFFFFFFu || psm_better(os, ov, bs, bv))) { bs = os; bv = ov; }
      }
      if (lane == 0) { s_rs[warp] = bs; s_rv[warp] = bv; }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int q = 1; q < kScoreThreads / 32; q++)
          if (s_rv[q] != 0xFFFFFFFFu && (bv == 0xFFFFFFFFu || psm_better(s_rs[q], s_rv[q], bs, bv))) { bs = s_rs[q]; bv = s_rv[q]; }
        s_best_s = bs; s_best_v = bv;
        md_psm row;
        row.spectrum_id = pr.spectrum_id; row.rank = 0; row.is_decoy = 0; row.charge = (uint8_t)pr.charge; row.candidate = 0; row.var_mask = 0;
        row.mod_weight = 0; row.raw_score = 0; row.score = 0.0f; row.n_targets = nt; row.n_decoys = nd; row._pad = 0;
        if (bv != 0xFFFFFFFFu) {
          row.rank = (uint16_t)(r + 1); row.raw_score = bs;
          row.score = (float)(0.005 * (double)bs / (150.0 * 65536.0));
          if (bv < nt) { row.is_decoy = 0; row.candidate = (uint64_t)A.cand_pep[t0c + bv] + 1; row.var_mask = A.cand_mask[t0c + bv]; row.mod_weight = A.cand_w[t0c + bv]; }
          else { const uint64_t j = (uint64_t)s * C.n_per + (bv - nt); row.is_decoy = 1; row.candidate = bv - nt; row.var_mask = A.dec_mask[j]; row.mod_weight = A.dec_w[j]; }
        }
        A.psm[(uint64_t)s * K + r] = row;
      }
      __syncthreads();
      prev_s = s_best_s; prev_v = s_best_v; have_prev = prev_v != 0xFFFFFFFFu;
      if (!have_prev) {  // fewer candidates than rows: the remaining rows are empty
        if (threadIdx.x == 0) {
          for (uint32_t r2 = r + 1; r2 < K; r2++) {
            md_psm row;
            row.spectrum_id = pr.spectrum_id; row.rank = 0; row.is_decoy = 0; row.charge = (uint8_t)pr.charge; row.candidate = 0; row.var_mask = 0;
            row.mod_weight = 0; row.raw_score = 0; row.score = 0.0f; row.n_targets = nt; row.n_decoys = nd; row._pad = 0;
            A.psm[(uint64_t)s * K + r2] = row;
          }
        }
        break;
      }
    }
    __syncthreads();
  }
  for (int o = 16; o; o >>= 1) { my_pairs += __shfl_xor_sync(0xffffffffu, my_pairs, o); my_bytes += __shfl_xor_sync(0xffffffffu, my_bytes, o); }
  if (lane == 0 && my_pairs) { atomicAdd(&A.stat64[0], my_pairs); atomicAdd(&A.stat64[1], my_bytes); }
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

void score_run_dev(md_ctx* ctx, const SpectraDev& S, uint64_t n_peaks, const md_search_params& p, uint32_t n_per, md_psm* psm_dev) {
  IdentifyWorkspace& W = ctx->ws;
  const uint32_t n = S.n;
  if (!n) return;
  const int64_t w = (int64_t)llround(p.fragment_tolerance * 1000000.0);
  const uint32_t mfc = p.max_fragment_charge ? p.max_fragment_charge : 3;
  MD_REQUIRE(mfc <= 3, MD_ERR_UNSUPPORTED, "max_fragment_charge > 3 (the reference fixes it to 3: comet_parameter.rs:55)");
  // ---- K4a
  W.pk_bin.need(n_peaks + 1); W.pk_yq.need(n_peaks + 1); W.pk_count.need(n + 1); W.pk_hbin.need(n + 1);
  DevBuf<int> d_flag; d_flag.need(1);
  MD_CUDA(cudaMemsetAsync(d_flag.p, 0, sizeof(int), ctx->stream));
  MD_CUDA(cudaMemsetAsync(W.pk_yq.p, 0, (n_peaks + 1) * sizeof(int32_t), ctx->stream));
  MD_LAUNCH(ctx, k_bin_spectra, blocks((uint64_t)n * 32, 128), 128, 0, S.peak_off, S.peak_mz, S.peak_int, W.prec.p, n, w, p.min_peaks, W.pk_bin.p, W.pk_yq.p,
            W.pk_count.p, W.pk_hbin.p, d_flag.p);
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
    split_qr(m + f, C.w, &C.tq[c], &C.tr[c]);
    split_qr(m + f + v, C.w, &C.vq[c], &C.vr[c]);
    if (c < MD_NCODES && ctx->mods.has_var[c]) has_var = true;
  }
  uint64_t n_targets = 0;
  MD_CUDA(cudaMemcpyAsync(&n_targets, W.cand_off.p + n, sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream));
  MD_CUDA(cudaStreamSynchronize(ctx->stream));
  W.tscore.need(n_targets + 1); W.dscore.need((size_t)n * n_per + 1);
  DevBuf<uint32_t>& work = W.counters; work.need(4);
  MD_CUDA(cudaMemsetAsync(work.p, 0, 4 * sizeof(uint32_t), ctx->stream));
  W.stat64.need(2);
  MD_CUDA(cudaMemsetAsync(W.stat64.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
  ScoreArgs A;
  A.prec = W.prec.p; A.n_spec = n; A.peak_off = S.peak_off; A.pk_bin = W.pk_bin.p; A.pk_yq = W.pk_yq.p; A.pk_count = W.pk_count.p; A.pk_hbin = W.pk_hbin.p;
  A.cand_off = W.cand_off.p; A.cand_desc = W.cand_desc.p; A.cand_mask = W.cand_mask.p; A.cand_w = W.cand_w.p; A.cand_pep = W.cand_pep.p;
  A.idx_rows = ctx->index.rows.p;
  A.dec_rows = W.dec_rows.p; A.dec_len = W.dec_len.p; A.dec_mask = W.dec_mask.p; A.dec_w = W.dec_w.p; A.dec_count = n_per ? W.dec_count.p : nullptr;
  A.tscore = W.tscore.p; A.dscore = W.dscore.p; A.psm = psm_dev; A.work = work.p; A.stat64 = W.stat64.p;
  const size_t smem = (size_t)kTileBins * sizeof(int32_t);
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
  MD_CUDA(cudaMemcpyAsync(h_stat, W.stat64.p, sizeof(h_stat), cudaMemcpyDeviceToHost, ctx->stream));
  const int unsorted = d2h_scalar(ctx, d_flag.p);
  { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]); ctx->acc_ms_kscore += ms; ctx->acc_pairs += h_stat[0]; ctx->acc_score_bytes += h_stat[1]; }
  MD_REQUIRE(!unsorted, MD_ERR_INVALID, "spectra: peaks of a spectrum must be sorted by m/z");
}
