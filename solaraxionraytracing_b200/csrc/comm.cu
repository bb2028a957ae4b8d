// comm.cu — the one collective of the path: the sum of the detector images and counters of all GPUs.
//
// Rays shard over GPUs by global ray index (Philox counter), with no exchange while tracing; what has
// to be merged afterwards is additive: image [M][256][256], the sum-of-squares image and the counters. The handle keeps
// image and sum-of-squares image in ONE allocation followed by a small f64 area; sart_allreduce writes the counters into
// that area as f64 (exact below 2^53) and issues ONE ncclAllReduce(sum, f64) over the whole block into a separate
// "merged" block, so a rank's own accumulators stay untouched (no reset between steps) and the call can be repeated.
//
// NCCL is resolved with dlopen at first use: libsart.so has no link-time dependency on it, a process that already
// loaded NCCL (PyTorch) shares that copy, and a host without NCCL can use everything else.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>

#include "kernels.h"
#include "sart_internal.h"

namespace sart {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      api.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (api.lib) break;
    }
    if (api.lib) {
#define SART_SYM(field, sym) api.field = reinterpret_cast<decltype(api.field)>(dlsym(api.lib, sym))
      SART_SYM(GetUniqueId, "ncclGetUniqueId");
      SART_SYM(CommInitRank, "ncclCommInitRank");
      SART_SYM(CommInitAll, "ncclCommInitAll");
      SART_SYM(AllReduce, "ncclAllReduce");
      SART_SYM(GroupStart, "ncclGroupStart");
      SART_SYM(GroupEnd, "ncclGroupEnd");
      SART_SYM(CommDestroy, "ncclCommDestroy");
      SART_SYM(GetErrorString, "ncclGetErrorString");
#undef SART_SYM
      if (!api.GetUniqueId || !api.CommInitRank || !api.CommInitAll || !api.AllReduce || !api.GroupStart || !api.GroupEnd ||
          !api.CommDestroy || !api.GetErrorString) {
        dlclose(api.lib);
        api.lib = nullptr;
      }
    }
  }
  return api.lib ? &api : nullptr;
}

static int nccl_fail(ncclResult_t r, const char* what) {
  NcclApi* n = nccl();
  return fail(SART_ERR_CUDA, "%s: %s", what, n ? n->GetErrorString(r) : "NCCL unavailable");
}
#define SART_NCCL(call)                                      \
  do {                                                       \
    ncclResult_t r__ = (call);                               \
    if (r__ != ncclSuccess) return nccl_fail(r__, #call);    \
  } while (0)

// counters -> f64 words behind the two images (pack), one thread per (mass, word)
__global__ void k_pack_counters(const sart_counters_t* __restrict__ c, int nMasses, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nMasses * kCounterWords) return;
  const int m = i / kCounterWords, w = i % kCounterWords;
  const unsigned long long* raw = reinterpret_cast<const unsigned long long*>(c + m);
  constexpr int nInt = int(offsetof(sart_counters_t, sum_w) / 8);
  out[i] = w < nInt ? double(raw[w]) : reinterpret_cast<const double*>(raw)[w];
}

}  // namespace sart

using namespace sart;

extern "C" {

int sart_comm_unique_id(char id[SART_COMM_ID_BYTES]) {
  if (!id) return fail(SART_ERR_ARG, "sart_comm_unique_id: id is NULL");
  NcclApi* n = nccl();
  if (!n) return fail(SART_ERR_CONFIG, "NCCL (libnccl.so.2) could not be loaded: %s", dlerror() ? dlerror() : "not found");
  static_assert(sizeof(ncclUniqueId) == SART_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId u;
  SART_NCCL(n->GetUniqueId(&u));
  std::memcpy(id, &u, sizeof u);
  return SART_OK;
}

int sart_comm_init_rank(sart_handle_t* h, int n_ranks, int rank, const char id[SART_COMM_ID_BYTES]) {
  if (!h || !id) return fail(SART_ERR_ARG, "sart_comm_init_rank: NULL argument");
  if (n_ranks < 1 || rank < 0 || rank >= n_ranks) return fail(SART_ERR_ARG, "sart_comm_init_rank: rank %d of %d", rank, n_ranks);
  if (h->comm) return fail(SART_ERR_ARG, "sart_comm_init_rank: the handle already has a communicator");
  NcclApi* n = nccl();
  if (!n) return fail(SART_ERR_CONFIG, "NCCL (libnccl.so.2) could not be loaded");
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  ncclUniqueId u;
  std::memcpy(&u, id, sizeof u);
  ncclComm_t c = nullptr;
  const ncclResult_t r = n->CommInitRank(&c, n_ranks, u, rank);
  cudaSetDevice(prev);
  if (r != ncclSuccess) return nccl_fail(r, "ncclCommInitRank");
  h->comm = c; h->comm_ranks = n_ranks; h->comm_rank = rank;
  return SART_OK;
}

int sart_comm_init_all(sart_handle_t* const* handles, int n) {
  if (!handles || n < 1 || n > 64) return fail(SART_ERR_ARG, "sart_comm_init_all: need 1..64 handles");
  int devs[64];
  for (int i = 0; i < n; ++i) {
    if (!handles[i]) return fail(SART_ERR_ARG, "sart_comm_init_all: handle %d is NULL", i);
    if (handles[i]->comm) return fail(SART_ERR_ARG, "sart_comm_init_all: handle %d already has a communicator", i);
    devs[i] = handles[i]->device;
    for (int j = 0; j < i; ++j)
      if (devs[j] == devs[i]) return fail(SART_ERR_ARG, "sart_comm_init_all: handles %d and %d share device %d", j, i, devs[i]);
  }
  NcclApi* api = nccl();
  if (!api) return fail(SART_ERR_CONFIG, "NCCL (libnccl.so.2) could not be loaded");
  ncclComm_t comms[64];
  SART_NCCL(api->CommInitAll(comms, n, devs));
  for (int i = 0; i < n; ++i) { handles[i]->comm = comms[i]; handles[i]->comm_ranks = n; handles[i]->comm_rank = i; }
  return SART_OK;
}

void sart_comm_destroy(sart_handle_t* h) {
  if (!h || !h->comm) return;
  NcclApi* n = nccl();
  if (n) {
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    n->CommDestroy(static_cast<ncclComm_t>(h->comm));
  }
  h->comm = nullptr; h->comm_ranks = 0; h->comm_rank = 0;
}

int sart_allreduce(sart_handle_t* const* handles, int n) {
  if (!handles || n < 1) return fail(SART_ERR_ARG, "sart_allreduce: need at least one handle");
  NcclApi* api = nccl();
  if (!api) return fail(SART_ERR_CONFIG, "NCCL (libnccl.so.2) could not be loaded");
  for (int i = 0; i < n; ++i) {
    sart_handle* h = handles[i];
    if (!h || !h->comm) return fail(SART_ERR_ARG, "sart_allreduce: handle %d has no communicator (sart_comm_init_rank / _all)", i);
    if (!h->d_image) return fail(SART_ERR_NOMEM, "sart_allreduce: handle %d has no image buffers", i);
    if (h->n_masses != handles[0]->n_masses) return fail(SART_ERR_ARG, "sart_allreduce: handles differ in their number of axion masses");
  }
  int prev = 0;
  cudaGetDevice(&prev);
  // the counters of every handle as f64 words behind its two images, then one all-reduce per handle inside one group
  for (int i = 0; i < n; ++i) {
    sart_handle* h = handles[i];
    cudaSetDevice(h->device);
    const size_t len = size_t(h->n_masses) * SART_IMAGE_BINS * SART_IMAGE_BINS;
    const size_t words = merged_words(h->n_masses);
    if (h->merged_masses != h->n_masses) {
      cudaFree(h->d_merged);
      h->d_merged = nullptr; h->merged_masses = 0;
      cudaError_t e = cudaMalloc(&h->d_merged, words * sizeof(double));
      if (e != cudaSuccess) { cudaSetDevice(prev); return fail(SART_ERR_CUDA, "cudaMalloc(merged): %s", cudaGetErrorString(e)); }
      h->merged_masses = h->n_masses;
    }
    const int nw = h->n_masses * kCounterWords;
    k_pack_counters<<<(nw + 127) / 128, 128, 0, h->stream>>>(h->d_counters, h->n_masses, h->d_image + 2 * len);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { cudaSetDevice(prev); return fail(SART_ERR_CUDA, "k_pack_counters: %s", cudaGetErrorString(e)); }
  }
  ncclResult_t r = api->GroupStart();
  for (int i = 0; i < n && r == ncclSuccess; ++i) {
    sart_handle* h = handles[i];
    cudaSetDevice(h->device);
    r = api->AllReduce(h->d_image, h->d_merged, merged_words(h->n_masses), ncclDouble, ncclSum, static_cast<ncclComm_t>(h->comm),
                       h->stream);
  }
  const ncclResult_t r2 = api->GroupEnd();
  cudaSetDevice(prev);
  if (r != ncclSuccess) return nccl_fail(r, "ncclAllReduce");
  if (r2 != ncclSuccess) return nccl_fail(r2, "ncclGroupEnd");
  for (int i = 0; i < n; ++i) handles[i]->merged_valid = 1;
  return SART_OK;
}

int sart_read_merged(sart_handle_t* h, double* image, double* image_w2, sart_counters_t* counters) {
  if (!h) return fail(SART_ERR_ARG, "handle is NULL");
  if (!h->d_merged || !h->merged_valid || h->merged_masses != h->n_masses)
    return fail(SART_ERR_ARG, "sart_read_merged: no merged result (call sart_allreduce first)");
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(h->device);
  const size_t len = size_t(h->n_masses) * SART_IMAGE_BINS * SART_IMAGE_BINS;
  cudaError_t e = cudaSuccess;
  if (image) e = cudaMemcpyAsync(image, h->d_merged, len * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess && image_w2) e = cudaMemcpyAsync(image_w2, h->d_merged + len, len * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  double words[SART_MAX_MASSES * kCounterWords];
  if (e == cudaSuccess && counters)
    e = cudaMemcpyAsync(words, h->d_merged + 2 * len, size_t(h->n_masses) * kCounterWords * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaSetDevice(prev);
  if (e != cudaSuccess) return fail(SART_ERR_CUDA, "sart_read_merged: %s", cudaGetErrorString(e));
  if (counters) {
    constexpr int nInt = int(offsetof(sart_counters_t, sum_w) / 8);
    for (int m = 0; m < h->n_masses; ++m) {
      unsigned long long* raw = reinterpret_cast<unsigned long long*>(counters + m);
      for (int w = 0; w < kCounterWords; ++w) {
        const double v = words[m * kCounterWords + w];
        if (w < nInt) raw[w] = (unsigned long long)(v + 0.5);
        else reinterpret_cast<double*>(raw)[w] = v;
      }
    }
  }
  return SART_OK;
}

}  // extern "C"
