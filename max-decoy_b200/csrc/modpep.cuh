// modpep.cuh -- device restatement of the ModifiedPeptide mass-window logic
// (models/peptides/modified_peptide.rs:157-159 hits_mass_tolerance, :369-401 remove_all_variable_modifications,
//  :512-543 try_variable_modifications; utility/combinations/n_choose_k.rs:12-49 subset order).
// A working peptide is (sequence, weight, mask of residues carrying their variable modification).  Which residues carry
// their letter's fixed modification, and which can take the variable one, follows from letter and position alone
// (md_fix_applies / md_can_var in common.cuh: anywhere, or the matching terminus; a slot that holds the fixed
// modification never takes the variable one, :355, :532).
#pragma once
#include "common.cuh"

#define MD_VAR_ENUM_CAP (1u << 22)  // subsets tried per peptide before giving up (the reference would spin)

__device__ __forceinline__ bool md_in_window(int64_t w, int64_t lo, int64_t hi) { return lo <= w && w <= hi; }

// previous k-subset bit mask in descending numeric order; 0 when m was the smallest (2^k - 1)
__host__ __device__ __forceinline__ uint64_t md_prev_combination(uint64_t m) {
  uint64_t t = m & (~m + 1);
  if (t != 1) return (m & ~t) | (t >> 1);
  uint64_t ones = m & ~(m + 1);
  uint64_t rest = m & ~ones;
  if (rest == 0) return 0;
  uint64_t low = rest & (~rest + 1);
#ifdef __CUDA_ARCH__
  int j = __popcll(ones);
#else
  int j = __builtin_popcountll(ones);
#endif
  uint64_t moved = low >> 1;
  uint64_t packed = ((1ULL << j) - 1) * (moved >> j);
  return (rest & ~low) | moved | packed;
}

// try_variable_modifications when exactly one letter has a variable modification and no fixed one (M.var_simple_code):
// the weight depends on the subset size only, and the first subset of size n in NChooseK order is the first n
// positions.  `allpos` = positions holding that letter, `base` = weight without any variable modification.
// Same contract as md_try_variable (which calls this): on failure (w, mask) = the last subset tried.
__device__ __forceinline__ bool md_try_variable_simple(const ModTables& M, uint64_t allpos, int64_t base, int64_t& w, uint64_t& mask, int64_t lo,
                                                       int64_t hi) {
  const uint32_t d = (uint32_t)__popcll(allpos);
  const uint32_t nmax = M.nvar < d ? M.nvar : d;
  const int64_t delta = M.var[M.var_simple_code];
  uint64_t first = 0, rest = allpos;
  for (uint32_t n = 1; n <= nmax; n++) {
    first |= rest & (~rest + 1); rest &= rest - 1;
    const int64_t wn = base + (int64_t)n * delta;
    if (md_in_window(wn, lo, hi)) { w = wn; mask = first; return true; }
  }
  // last subset tried: size nmax, the nmax last positions
  uint64_t last = allpos;
  for (uint32_t k = d; k > nmax; k--) last &= last - 1;  // drop the d-nmax lowest positions
  w = base + (int64_t)nmax * delta; mask = last;
  return false;
}

// Seq: functor uint32_t operator()(uint32_t i) -> residue code.
// On return true: (w, mask) = the first configuration inside [lo,hi].  On return false: (w, mask) = the last
// configuration tried (the reference leaves it applied, which the decoy repair loop then sees), or unchanged
// if nothing was tried.  *overflow is set when the enumeration cap is hit.
template <class Seq>
__device__ bool md_try_variable(const ModTables& M, Seq seq, uint32_t len, int64_t& w, uint64_t& mask, int64_t lo, int64_t hi, int* overflow) {
  if (M.nvar == 0) return false;
  uint64_t allpos = 0, eff = 0;
  for (uint32_t i = 0; i < len; i++) {
    uint32_t c = seq(i);
    if (M.has_var[c]) { allpos |= 1ULL << i; if (md_can_var(M, c, i, len)) eff |= 1ULL << i; }
  }
  const uint32_t d = (uint32_t)__popcll(allpos);
  if (d == 0) return false;
  int64_t base = w;
  for (uint64_t m = mask; m; m &= m - 1) base -= M.var[seq((uint32_t)__ffsll((long long)m) - 1)];
  const uint32_t nmax = M.nvar < d ? M.nvar : d;
  if (M.var_simple_code >= 0) return md_try_variable_simple(M, allpos, base, w, mask, lo, hi);
  uint32_t tried = 0;
  uint64_t last_sel = 0; int64_t last_w = w; bool any = false;
  for (uint32_t n = 1; n <= nmax; n++) {
    uint64_t cm = (d == 64 ? ~0ULL : ((1ULL << d) - 1)) ^ ((1ULL << (d - n)) - 1);
    while (cm) {
      // compressed bit (d-1-b) <-> b-th position (ascending) of allpos
      uint64_t sel = 0, rest = allpos; int64_t wn = base;
      for (uint32_t b = 0; b < d; b++) {
        uint64_t bit = rest & (~rest + 1); rest &= rest - 1;
        if ((cm >> (d - 1 - b)) & 1) {
          if (bit & eff) { sel |= bit; wn += M.var[seq((uint32_t)__ffsll((long long)bit) - 1)]; }
        }
      }
      if (md_in_window(wn, lo, hi)) { w = wn; mask = sel; return true; }
      last_sel = sel; last_w = wn; any = true;
      if (++tried >= MD_VAR_ENUM_CAP) { if (overflow) *overflow = 1; w = last_w; mask = last_sel; return false; }
      cm = md_prev_combination(cm);
    }
  }
  if (any) { w = last_w; mask = last_sel; }
  return false;
}
