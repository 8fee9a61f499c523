// comm.cu -- multi-GPU entry points of the C ABI: one rank per GPU, spectra sharded, index replicated; the only exchange
// of the path is the all-gather of the fixed-width PSM tables (tasks/identification.rs:201: every iteration of the
// spectrum loop is independent).  NCCL is loaded at run time (dlopen) so that a single-GPU host needs no NCCL at all and
// the library binds to whichever libnccl.so.2 the process already uses (e.g. the one a PyTorch host brought along).
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>

#include "ctx.h"

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};

NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) { api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (api.handle) break; }
    if (!api.handle) { api.error = std::string("NCCL not found (dlopen libnccl.so.2): ") + (dlerror() ? dlerror() : ""); return; }
    auto sym = [&](const char* n) { void* p = dlsym(api.handle, n); if (!p && api.error.empty()) api.error = std::string("NCCL symbol missing: ") + n; return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  });
  return api;
}

thread_local std::string g_comm_err;
int comm_fail(md_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg;
  g_comm_err = msg;
  return code;
}
#define MD_NCCL(ctx, api, expr)                                                                                      \
  do {                                                                                                               \
    ncclResult_t r_ = (expr);                                                                                        \
    if (r_ != ncclSuccess) return comm_fail(ctx, MD_ERR_DEVICE, std::string(#expr) + ": " + (api).GetErrorString(r_)); \
  } while (0)

bool is_device_pointer(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

}  // namespace

void comm_sync(md_ctx* ctx) {
  if (ctx && ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
}

void comm_wait_for_buffer(md_ctx* ctx, const void* buf) {
  if (!ctx || !ctx->comm_stream || !buf) return;
  for (auto& g : ctx->pending)
    if (g.local == buf && g.done) cudaStreamWaitEvent(ctx->stream, g.done, 0);
}

void comm_release(md_ctx* ctx) {
  if (!ctx) return;
  if (ctx->comm_stream) cudaStreamSynchronize(ctx->comm_stream);
  if (ctx->comm) {
    NcclApi& api = nccl_api();
    if (api.CommDestroy) api.CommDestroy((ncclComm_t)ctx->comm);
  }
  for (auto& g : ctx->pending) { if (g.done) cudaEventDestroy(g.done); g.done = nullptr; g.local = nullptr; }
  if (ctx->ev_rows) { cudaEventDestroy(ctx->ev_rows); ctx->ev_rows = nullptr; }
  if (ctx->comm_stream) { cudaStreamDestroy(ctx->comm_stream); ctx->comm_stream = nullptr; }
  ctx->comm = nullptr; ctx->rank = 0; ctx->nranks = 1; ctx->n_gathers = 0;
}

extern "C" {

int md_comm_unique_id(uint8_t* id) {
  if (!id) return comm_fail(nullptr, MD_ERR_INVALID, "md_comm_unique_id: id is NULL");
  NcclApi& api = nccl_api();
  if (!api.error.empty()) return comm_fail(nullptr, MD_ERR_UNSUPPORTED, api.error);
  static_assert(sizeof(ncclUniqueId) == MD_COMM_ID_BYTES, "MD_COMM_ID_BYTES must equal sizeof(ncclUniqueId)");
  ncclUniqueId u;
  MD_NCCL(nullptr, api, api.GetUniqueId(&u));
  memcpy(id, &u, sizeof(u));
  return MD_OK;
}

int md_comm_init(md_ctx* ctx, int32_t rank, int32_t nranks, const uint8_t* id) {
  if (!ctx) return comm_fail(ctx, MD_ERR_INVALID, "md_comm_init: null ctx");
  if (nranks < 1 || rank < 0 || rank >= nranks) return comm_fail(ctx, MD_ERR_INVALID, "md_comm_init: rank / nranks out of range");
  comm_release(ctx);
  if (nranks == 1) { ctx->rank = 0; ctx->nranks = 1; return MD_OK; }
  if (!id) return comm_fail(ctx, MD_ERR_INVALID, "md_comm_init: id is NULL");
  NcclApi& api = nccl_api();
  if (!api.error.empty()) return comm_fail(ctx, MD_ERR_UNSUPPORTED, api.error);
  if (cudaSetDevice(ctx->device) != cudaSuccess) return comm_fail(ctx, MD_ERR_DEVICE, "md_comm_init: cudaSetDevice failed");
  ncclUniqueId u;
  memcpy(&u, id, sizeof(u));
  ncclComm_t comm = nullptr;
  MD_NCCL(ctx, api, api.CommInitRank(&comm, nranks, u, rank));
  ctx->comm = comm; ctx->rank = rank; ctx->nranks = nranks;
  if (cudaStreamCreateWithFlags(&ctx->comm_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->ev_rows, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->pending[0].done, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&ctx->pending[1].done, cudaEventDisableTiming) != cudaSuccess)
    return comm_fail(ctx, MD_ERR_DEVICE, "md_comm_init: cannot create the gather stream");
  return MD_OK;
}

int md_gather_psms(md_ctx* ctx, const md_psm* local, uint64_t rows, md_psm* all) {
  if (!ctx || (rows && (!local || !all))) return comm_fail(ctx, MD_ERR_INVALID, "md_gather_psms: null argument");
  if (!rows) return MD_OK;
  try {
    MD_CUDA(cudaSetDevice(ctx->device));
    const size_t bytes = rows * sizeof(md_psm);
    const bool ldev = is_device_pointer(local), adev = is_device_pointer(all);
    cudaStream_t st = ctx->stream;
    if (ctx->nranks == 1 || !ctx->comm) {
      if ((const void*)local != (const void*)all) MD_CUDA(cudaMemcpyAsync(all, local, bytes, cudaMemcpyDefault, st));
      if (!(ldev && adev)) MD_CUDA(cudaStreamSynchronize(st));
      return MD_OK;
    }
    NcclApi& api = nccl_api();
    if (ldev && adev) {
      // rows -> (event) -> gather on the comm stream; the ctx stream is free for the next batch
      MD_CUDA(cudaEventRecord(ctx->ev_rows, st));
      MD_CUDA(cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_rows, 0));
      MD_NCCL(ctx, api, api.AllGather(local, all, bytes, ncclUint8, (ncclComm_t)ctx->comm, ctx->comm_stream));
      md_ctx::PendingGather& g = ctx->pending[ctx->n_gathers++ & 1u];
      g.local = local;
      MD_CUDA(cudaEventRecord(g.done, ctx->comm_stream));
      return MD_OK;
    }
    const md_psm* send = local; md_psm* recv = all;
    if (!ldev) { send = ctx->gat_send.need(rows); MD_CUDA(cudaMemcpyAsync(ctx->gat_send.p, local, bytes, cudaMemcpyHostToDevice, st)); }
    if (!adev) recv = ctx->gat_recv.need(rows * (size_t)ctx->nranks);
    MD_NCCL(ctx, api, api.AllGather(send, recv, bytes, ncclUint8, (ncclComm_t)ctx->comm, st));
    if (!adev) MD_CUDA(cudaMemcpyAsync(all, recv, bytes * (size_t)ctx->nranks, cudaMemcpyDeviceToHost, st));
    if (!(ldev && adev)) MD_CUDA(cudaStreamSynchronize(st));
    return MD_OK;
  } catch (const MdError& e) {
    return comm_fail(ctx, e.code, e.msg);
  }
}

int md_comm_destroy(md_ctx* ctx) {
  if (!ctx) return comm_fail(ctx, MD_ERR_INVALID, "md_comm_destroy: null ctx");
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  comm_release(ctx);
  return MD_OK;
}

}  // extern "C"
