// C ABI: peer buffers for the fused evaluation + gather over NVLink (include/fastcheb.h, "multi-GPU").
//
// One process per GPU.  The gathering rank allocates the result buffer with fastcheb_peer_alloc and ships the
// 64-byte CUDA IPC handle to the other ranks (any transport: torch.distributed object broadcast here); they map it
// with fastcheb_peer_open and pass `mapped + their slice offset` as the OUTPUT pointer of fastcheb_eval_async /
// fastcheb_eval_jac_async.  The evaluation kernel's own result stores then travel over NVLink/NVSwitch straight into
// the gatherer's HBM: no staging copy, no separate collective kernel — the gather is fused into the evaluation.
#include <cstring>
#include "host_util.h"

using namespace fastcheb;

extern "C" {

int fastcheb_peer_alloc(void** dev_ptr, size_t bytes, void* ipc_handle_out, int device) {
  if (!dev_ptr || !ipc_handle_out || bytes == 0) return fail(FASTCHEB_EINVAL, "bad peer_alloc argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == FASTCHEB_IPC_HANDLE_BYTES, "IPC handle size");
  *dev_ptr = nullptr;
  if (!device_info(device)) return FASTCHEB_ECUDA;
  DeviceGuard dg(device);
  void* p = nullptr;
  FC_CUDA(cudaMalloc(&p, bytes));
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(FASTCHEB_ECUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  }
  memcpy(ipc_handle_out, &h, sizeof h);
  *dev_ptr = p;
  return FASTCHEB_OK;
}

int fastcheb_peer_free(void* dev_ptr, int device) {
  if (!dev_ptr) return FASTCHEB_OK;
  DeviceGuard dg(device);
  FC_CUDA(cudaFree(dev_ptr));
  return FASTCHEB_OK;
}

int fastcheb_peer_open(void** mapped_ptr, const void* ipc_handle, int device) {
  if (!mapped_ptr || !ipc_handle) return fail(FASTCHEB_EINVAL, "null argument");
  *mapped_ptr = nullptr;
  if (!device_info(device)) return FASTCHEB_ECUDA;
  DeviceGuard dg(device);
  cudaIpcMemHandle_t h;
  memcpy(&h, ipc_handle, sizeof h);
  FC_CUDA(cudaIpcOpenMemHandle(mapped_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return FASTCHEB_OK;
}

int fastcheb_peer_close(void* mapped_ptr, int device) {
  if (!mapped_ptr) return FASTCHEB_OK;
  DeviceGuard dg(device);
  FC_CUDA(cudaIpcCloseMemHandle(mapped_ptr));
  return FASTCHEB_OK;
}

int fastcheb_peer_read(void* host_dst, const void* dev_src, size_t bytes, int device) {
  if (!host_dst || !dev_src) return fail(FASTCHEB_EINVAL, "null argument");
  DeviceGuard dg(device);
  FC_CUDA(cudaMemcpy(host_dst, dev_src, bytes, cudaMemcpyDeviceToHost));
  return FASTCHEB_OK;
}

}  // extern "C"
