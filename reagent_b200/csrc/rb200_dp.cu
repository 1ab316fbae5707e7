// reagent_b200 -- peer-memory plumbing of the fused data-parallel optimizer step
// (rb200_adam_soft_update with dp_world > 1, rb200_optim.cu): IPC-exportable device buffers
// and their mapping into peer processes (one process per GPU, NVLink / NVSwitch P2P).
//
// The reference has no collective on this path; docs/distributed.rst:12-22 documents the
// intent (synchronous data parallelism, gradient all-reduce per step).
#include <string.h>

#include "rb200_common.cuh"

using namespace rb200;

static_assert(sizeof(cudaIpcMemHandle_t) == RB200_IPC_HANDLE_BYTES, "IPC handle size");

extern "C" int rb200_dp_alloc(int64_t bytes, void** out_ptr) {
  if (bytes <= 0 || !out_ptr) { set_last_error("rb200_dp_alloc: bad argument"); return RB200_E_INVALID; }
  void* p = nullptr;
  if (int rc = check_cuda(cudaMalloc(&p, (size_t)bytes), "cudaMalloc(dp buffer)")) return rc;
  if (int rc = check_cuda(cudaMemset(p, 0, (size_t)bytes), "cudaMemset(dp buffer)")) { cudaFree(p); return rc; }
  if (int rc = check_cuda(cudaDeviceSynchronize(), "cudaDeviceSynchronize(dp buffer)")) { cudaFree(p); return rc; }
  *out_ptr = p;
  return RB200_OK;
}

extern "C" int rb200_dp_free(void* ptr) {
  if (!ptr) return RB200_OK;
  return check_cuda(cudaFree(ptr), "cudaFree(dp buffer)");
}

extern "C" int rb200_dp_ipc_handle(void* ptr, unsigned char* handle_out_host) {
  if (!ptr || !handle_out_host) { set_last_error("rb200_dp_ipc_handle: null argument"); return RB200_E_INVALID; }
  cudaIpcMemHandle_t h;
  if (int rc = check_cuda(cudaIpcGetMemHandle(&h, ptr), "cudaIpcGetMemHandle")) return rc;
  memcpy(handle_out_host, &h, sizeof(h));
  return RB200_OK;
}

extern "C" int rb200_dp_ipc_open(const unsigned char* handle_host, void** out_ptr) {
  if (!handle_host || !out_ptr) { set_last_error("rb200_dp_ipc_open: null argument"); return RB200_E_INVALID; }
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_host, sizeof(h));
  void* p = nullptr;
  if (int rc = check_cuda(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return rc;
  *out_ptr = p;
  return RB200_OK;
}

extern "C" int rb200_dp_ipc_close(void* ptr) {
  if (!ptr) return RB200_OK;
  return check_cuda(cudaIpcCloseMemHandle(ptr), "cudaIpcCloseMemHandle");
}
