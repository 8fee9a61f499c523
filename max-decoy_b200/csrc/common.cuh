// common.cuh -- shared definitions of the sm_100a MaxDecoy hot-path library.
// Reference citations are relative to /root/reference/src/proteomic/.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/maxdecoy.h"

#define MD_NSM_FALLBACK 148

// ----------------------------------------------------------------------------------------
// error handling: nothing throws across the C ABI
// ----------------------------------------------------------------------------------------
struct MdError {
  int code;
  std::string msg;
};

#define MD_CUDA(expr)                                                                        \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      throw MdError{MD_ERR_DEVICE, std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + \
                                       __FILE__ + ":" + std::to_string(__LINE__) + ")"};     \
  } while (0)

#define MD_REQUIRE(cond, code, text)        \
  do {                                      \
    if (!(cond)) throw MdError{(code), (text)}; \
  } while (0)

// ----------------------------------------------------------------------------------------
// device buffer that only grows (the ctx keeps its workspaces between calls)
// ----------------------------------------------------------------------------------------
template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
  // contents are NOT preserved on growth
  T* need(size_t n) {
    if (n > cap) {
      release();
      size_t want = n + n / 8 + 64;
      MD_CUDA(cudaMalloc((void**)&p, want * sizeof(T)));
      cap = want;
    }
    return p;
  }
  size_t bytes() const { return cap * sizeof(T); }
};

// ----------------------------------------------------------------------------------------
// residue codes and masses
// ----------------------------------------------------------------------------------------
// code = letter - 'A' for 'A'..'Z', 26 for anything else (mass 0 = X, amino_acid.rs:115)
#define MD_NCODES 27
#define MD_CODE_OTHER 26

__host__ __device__ inline uint32_t md_code_of(uint8_t c) {
  return (c >= 'A' && c <= 'Z') ? (uint32_t)(c - 'A') : (uint32_t)MD_CODE_OTHER;
}
__host__ __device__ inline uint8_t md_letter_of(uint32_t code) {
  return code < 26 ? (uint8_t)('A' + code) : (uint8_t)'?';
}

// int(mono * 1e6) of amino_acid.rs:7-35, evaluated once on the host in double (mass/mod.rs:6-8);
// the literals below are the resulting integers (E and K carry the truncation quirk).
static const int64_t kResidueMassByCode[MD_NCODES] = {
    /*A*/ 71037110,  /*B*/ 114534950, /*C*/ 103009190, /*D*/ 115026940, /*E*/ 129042589,
    /*F*/ 147068410, /*G*/ 57021460,  /*H*/ 137058910, /*I*/ 113084060, /*J*/ 113084060,
    /*K*/ 128094959, /*L*/ 113084060, /*M*/ 131040490, /*N*/ 114042930, /*O*/ 109052800,
    /*P*/ 97052760,  /*Q*/ 128058580, /*R*/ 156101110, /*S*/ 87032030,  /*T*/ 101047680,
    /*U*/ 150953630, /*V*/ 99068410,  /*W*/ 186079310, /*X*/ 0,         /*Y*/ 163063330,
    /*Z*/ 128550590, /*other*/ 0};

// alphabet index (0..20, MD_ALPHABET order) of a code, or -1
__host__ __device__ inline int md_alpha_of_code(uint32_t code) {
  // MD_ALPHABET = "ARNDCEQGHJKMFPOSTUVWY"
  switch (code) {
    case 0: return 0;    // A
    case 17: return 1;   // R
    case 13: return 2;   // N
    case 3: return 3;    // D
    case 2: return 4;    // C
    case 4: return 5;    // E
    case 16: return 6;   // Q
    case 6: return 7;    // G
    case 7: return 8;    // H
    case 9: return 9;    // J
    case 10: return 10;  // K
    case 12: return 11;  // M
    case 5: return 12;   // F
    case 15: return 13;  // P
    case 14: return 14;  // O
    case 18: return 15;  // S
    case 19: return 16;  // T
    case 20: return 17;  // U
    case 21: return 18;  // V
    case 22: return 19;  // W
    case 24: return 20;  // Y
    default: return -1;
  }
}

// hash of a generalized sequence: part of the canonical peptide order (must equal the oracle's)
__host__ __device__ inline uint64_t md_hash_init() { return 0xcbf29ce484222325ULL; }
__host__ __device__ inline uint64_t md_hash_step(uint64_t h, uint8_t b) { return (h ^ b) * 0x100000001b3ULL; }
__host__ __device__ inline uint64_t md_hash_fin(uint64_t h, uint32_t len) {
  h ^= len;
  h ^= h >> 33; h *= 0xff51afd7ed558ccdULL; h ^= h >> 33; h *= 0xc4ceb9fe1a85ec53ULL; h ^= h >> 33;
  return h;
}

// ----------------------------------------------------------------------------------------
// modification set as the kernels see it (passed by value / __constant__-free)
// ----------------------------------------------------------------------------------------
struct ModTables {
  // per residue code
  int64_t mass[MD_NCODES];    // unmodified residue mass
  int64_t fix[MD_NCODES];     // fixed delta (0 if none)
  int64_t var[MD_NCODES];     // variable delta (0 if none)
  uint8_t has_fix[MD_NCODES];
  uint8_t has_var[MD_NCODES];
  // sorted modifiable letters (identification.rs:173-178): alphabet index and merged delta
  int32_t n_letters;
  int32_t letter_alpha[MD_ALPHABET_SIZE];
  int64_t letter_delta[MD_ALPHABET_SIZE];  // variable overrides fixed (identification.rs:190-196)
  int64_t letter_mass[MD_ALPHABET_SIZE];   // residue mass + merged delta (divisor of K_a)
  uint32_t nvar;                           // -n
  // fast path of try_variable_modifications: exactly one variable letter, without a fixed mod
  int32_t var_simple_code;                 // code of that letter, or -1
  // position of the letter's fixed / variable modification (modification.rs:24-33): MD_POS_A anywhere, MD_POS_N / MD_POS_C
  // terminus.  A terminal modification sits on the first / last residue only, and only when that residue is its letter
  // (add_modification_at, modified_peptide.rs:421-447; set_variable_modification_at, :339-367).
  uint8_t fix_pos[MD_NCODES];
  uint8_t var_pos[MD_NCODES];
  uint32_t has_terminal;                   // any modification with position N or C
};
#define MD_POS_A 0
#define MD_POS_N 1
#define MD_POS_C 2
__host__ __device__ inline bool md_pos_at(uint32_t pos, uint32_t i, uint32_t L) {
  return pos == MD_POS_A || (pos == MD_POS_N && i == 0) || (pos == MD_POS_C && i + 1 == L);
}
// does residue i of a sequence of L residues carry its letter's fixed modification?
__host__ __device__ inline bool md_fix_applies(const ModTables& M, uint32_t code, uint32_t i, uint32_t L) {
  return M.has_fix[code] && md_pos_at(M.fix_pos[code], i, L);
}
// can it take its letter's variable modification?  (the slot -- side chain, N-terminus, C-terminus -- must exist at this
// position and must not hold the fixed modification: AlreadyFixModificationInPlace, modified_peptide.rs:311,327,350)
__host__ __device__ inline bool md_can_var(const ModTables& M, uint32_t code, uint32_t i, uint32_t L) {
  return M.has_var[code] && md_pos_at(M.var_pos[code], i, L) && !(M.has_fix[code] && M.fix_pos[code] == M.var_pos[code]);
}

// Philox4x32-10 (must equal the oracle's)
struct Philox4 {
  uint32_t k0, k1, c0, c1, c2, c3;
  uint32_t o[4];
  int have;
  __host__ __device__ inline void init(uint64_t seed, uint32_t spectrum_id, uint32_t attempt, uint32_t tag) {
    k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
    c0 = 0; c1 = attempt; c2 = spectrum_id; c3 = tag; have = 0;
  }
  __host__ __device__ inline void refill() {
    uint32_t a = c0, b = c1, c = c2, d = c3, x = k0, y = k1;
#pragma unroll
    for (int r = 0; r < 10; r++) {
      uint64_t p0 = (uint64_t)0xD2511F53u * a, p1 = (uint64_t)0xCD9E8D57u * c;
      uint32_t n0 = (uint32_t)(p1 >> 32) ^ b ^ x, n1 = (uint32_t)p1;
      uint32_t n2 = (uint32_t)(p0 >> 32) ^ d ^ y, n3 = (uint32_t)p0;
      a = n0; b = n1; c = n2; d = n3;
      x += 0x9E3779B9u; y += 0xBB67AE85u;
    }
    o[0] = a; o[1] = b; o[2] = c; o[3] = d;
    c0++; have = 4;
  }
  __host__ __device__ inline uint32_t next() {
    if (!have) refill();
    uint32_t v = o[4 - have];
    have--;
    return v;
  }
  __host__ __device__ inline uint32_t below(uint32_t n) { return (uint32_t)(((uint64_t)next() * n) >> 32); }
};
#define MD_TAG_RANDOM 0x4D444543u
#define MD_TAG_PERMUTE 0x4D445045u

__host__ __device__ inline uint32_t md_attempt_cap(uint32_t n) { return 16u * n + 1024u; }

// score rows: residue codes, padded to 16 bytes; decoy rows are fixed 64-byte slots
#define MD_DECOY_ROW 64
// Accepted decoys live in TWO PLANES of 32-byte half rows (first halves of all slots, then second halves): the score
// kernel reads a row in 16-byte chunks only as far as the candidate is long, most decoys end inside the first half, and
// with this layout the DRAM bursts that bring one first half bring its neighbours' first halves, not unused padding.
#define MD_DECOY_HALF 32
__host__ __device__ inline size_t md_dec_byte(size_t n_slots, size_t slot, uint32_t i) {
  return i < MD_DECOY_HALF ? slot * MD_DECOY_HALF + i : (n_slots + slot) * MD_DECOY_HALF + (i - MD_DECOY_HALF);
}
