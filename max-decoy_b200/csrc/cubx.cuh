// cubx.cuh -- thin wrappers over CUB device primitives (scan / select / radix sort) that use the
// ctx's growing temp buffer and stream.  CUB is plumbing here; the hot kernels are hand-written.
#pragma once
#include <cub/cub.cuh>

#include "ctx.h"

struct MaxOpU32 {
  __host__ __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

template <class InIt, class OutIt>
inline void cubx_exclusive_sum(md_ctx* ctx, InIt in, OutIt out, size_t n) {
  size_t bytes = 0;
  MD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, in, out, (int)n, ctx->stream));
  void* tmp = ctx->ws.cub_tmp.need(bytes + 16);
  MD_CUDA(cub::DeviceScan::ExclusiveSum(tmp, bytes, in, out, (int)n, ctx->stream));
  ctx->cub_calls++;
}

inline void cubx_inclusive_max_u32(md_ctx* ctx, const uint32_t* in, uint32_t* out, size_t n) {
  size_t bytes = 0;
  MD_CUDA(cub::DeviceScan::InclusiveScan(nullptr, bytes, in, out, MaxOpU32(), (int)n, ctx->stream));
  void* tmp = ctx->ws.cub_tmp.need(bytes + 16);
  MD_CUDA(cub::DeviceScan::InclusiveScan(tmp, bytes, in, out, MaxOpU32(), (int)n, ctx->stream));
  ctx->cub_calls++;
}

// out = indices i in [0,n) with flags[i] != 0; *d_count = how many
inline void cubx_select_flagged_index(md_ctx* ctx, const uint8_t* flags, uint32_t* out, uint32_t* d_count, size_t n) {
  cub::CountingInputIterator<uint32_t> it(0);
  size_t bytes = 0;
  MD_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, it, flags, out, d_count, (int)n, ctx->stream));
  void* tmp = ctx->ws.cub_tmp.need(bytes + 16);
  MD_CUDA(cub::DeviceSelect::Flagged(tmp, bytes, it, flags, out, d_count, (int)n, ctx->stream));
  ctx->cub_calls++;
}

inline void cubx_unique_u64(md_ctx* ctx, const uint64_t* in, uint64_t* out, uint32_t* d_count, size_t n) {
  size_t bytes = 0;
  MD_CUDA(cub::DeviceSelect::Unique(nullptr, bytes, in, out, d_count, (int)n, ctx->stream));
  void* tmp = ctx->ws.cub_tmp.need(bytes + 16);
  MD_CUDA(cub::DeviceSelect::Unique(tmp, bytes, in, out, d_count, (int)n, ctx->stream));
  ctx->cub_calls++;
}

// stable LSD radix sort of (key, value) pairs; result lands in keys_out / vals_out
template <class K, class V>
inline void cubx_sort_pairs(md_ctx* ctx, const K* keys_in, K* keys_out, const V* vals_in, V* vals_out, size_t n) {
  size_t bytes = 0;
  MD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, (int)sizeof(K) * 8, ctx->stream));
  void* tmp = ctx->ws.cub_tmp.need(bytes + 16);
  MD_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0, (int)sizeof(K) * 8, ctx->stream));
  ctx->cub_calls++;
}

template <class K>
inline void cubx_sort_keys(md_ctx* ctx, const K* keys_in, K* keys_out, size_t n) {
  size_t bytes = 0;
  MD_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, bytes, keys_in, keys_out, (int)n, 0, (int)sizeof(K) * 8, ctx->stream));
  void* tmp = ctx->ws.cub_tmp.need(bytes + 16);
  MD_CUDA(cub::DeviceRadixSort::SortKeys(tmp, bytes, keys_in, keys_out, (int)n, 0, (int)sizeof(K) * 8, ctx->stream));
  ctx->cub_calls++;
}

template <class T>
inline T d2h_scalar(md_ctx* ctx, const T* d) {
  T v;
  MD_CUDA(cudaMemcpyAsync(&v, d, sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
  MD_CUDA(cudaStreamSynchronize(ctx->stream));
  return v;
}
