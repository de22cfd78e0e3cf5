// ehmc_comm: peer-memory mailboxes for the in-kernel all-reduce of the fused ensemble run (k_small_ens).
// One process per GPU; every rank allocates one small mailbox with cudaMalloc, exports it as a CUDA IPC handle, the
// caller exchanges the 64-byte handles (torch.distributed all_gather in the Python binding) and every rank maps all
// peers' mailboxes.  Kernels then store into / poll those mappings directly over NVLink / NVSwitch.
#include <cstring>
#include <new>

#include "host_defs.h"
#include "k_small_ens.cuh"

using namespace ehmc;

extern "C" int ehmc_comm_create(ehmc_ctx* ctx, int rank, int worldSize, ehmc_comm** out) {
  if (!ctx || !out) return fail(EHMC_ERR_INVALID, "ehmc_comm_create: NULL argument");
  *out = nullptr;
  if (worldSize < 1 || worldSize > EHMC_COMM_MAX_RANKS || rank < 0 || rank >= worldSize)
    return fail(EHMC_ERR_INVALID, "ehmc_comm_create: rank %d of %d (at most %d ranks)", rank, worldSize, EHMC_COMM_MAX_RANKS);
  CUDA_TRY(cudaSetDevice(ctx->device));
  ehmc_comm* m = new (std::nothrow) ehmc_comm();
  if (!m) return fail(EHMC_ERR_NOMEM, "out of host memory");
  m->ctx = ctx;
  m->rank = rank;
  m->world = worldSize;
  m->bytes = sizeof(double) * 2 * (size_t)worldSize * ENS_MB_STRIDE;
  if (cudaMalloc(&m->mailbox, m->bytes) != cudaSuccess || cudaMemset(m->mailbox, 0, m->bytes) != cudaSuccess ||
      cudaMalloc(&m->peers_dev, sizeof(double*) * worldSize) != cudaSuccess) {
    cudaGetLastError();
    ehmc_comm_destroy(m);
    return fail(EHMC_ERR_NOMEM, "ehmc_comm_create: device allocation failed");
  }
  m->peers[rank] = static_cast<double*>(m->mailbox);
  if (worldSize == 1) {
    CUDA_TRY(cudaMemcpy(m->peers_dev, m->peers, sizeof(double*), cudaMemcpyHostToDevice));
    m->connected = true;
  }
  CUDA_TRY(cudaDeviceSynchronize());
  *out = m;
  return EHMC_OK;
}

extern "C" int ehmc_comm_handle(ehmc_comm* m, void* out) {
  if (!m || !out) return fail(EHMC_ERR_INVALID, "ehmc_comm_handle: NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) <= EHMC_COMM_HANDLE_BYTES, "IPC handle size");
  CUDA_TRY(cudaSetDevice(m->ctx->device));
  cudaIpcMemHandle_t h;
  CUDA_TRY(cudaIpcGetMemHandle(&h, m->mailbox));
  memset(out, 0, EHMC_COMM_HANDLE_BYTES);
  memcpy(out, &h, sizeof(h));
  return EHMC_OK;
}

extern "C" int ehmc_comm_connect(ehmc_comm* m, const void* handles) {
  if (!m || !handles) return fail(EHMC_ERR_INVALID, "ehmc_comm_connect: NULL argument");
  if (m->connected) return EHMC_OK;
  CUDA_TRY(cudaSetDevice(m->ctx->device));
  for (int r = 0; r < m->world; ++r) {
    if (r == m->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const char*>(handles) + (size_t)r * EHMC_COMM_HANDLE_BYTES, sizeof(h));
    void* p = nullptr;
    const cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(EHMC_ERR_CUDA, "ehmc_comm_connect: cudaIpcOpenMemHandle of rank %d failed: %s (peer access / IPC "
                  "between the GPUs of this job is required for the fused ensemble run)", r, cudaGetErrorString(e));
    }
    m->peers[r] = static_cast<double*>(p);
    m->opened[r] = true;
  }
  CUDA_TRY(cudaMemcpy(m->peers_dev, m->peers, sizeof(double*) * m->world, cudaMemcpyHostToDevice));
  m->connected = true;
  return EHMC_OK;
}

extern "C" int ehmc_comm_destroy(ehmc_comm* m) {
  if (!m) return EHMC_OK;
  if (m->ctx) cudaSetDevice(m->ctx->device);
  for (int r = 0; r < m->world; ++r)
    if (m->opened[r] && m->peers[r]) cudaIpcCloseMemHandle(m->peers[r]);
  if (m->peers_dev) cudaFree(m->peers_dev);
  if (m->mailbox) cudaFree(m->mailbox);
  cudaGetLastError();
  delete m;
  return EHMC_OK;
}
