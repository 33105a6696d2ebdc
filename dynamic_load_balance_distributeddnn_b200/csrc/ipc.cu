// Symmetric-memory fallback allocator: cudaMalloc + CUDA IPC handles.
//
// Preferred allocator is torch.distributed._symmetric_memory (cuMem VMM + NVLS multicast binding);
// when that is unavailable this gives every rank a peer-mapped view of every other rank's buffer with
// plain CUDA IPC (no multicast => the NVLS algorithm is disabled, one-/two-shot P2P still work).
// Handle exchange happens in Python over torch.distributed (parallel/symm.py).
#include "common.cuh"
#include <string.h>

DLB_API int dlb_ipc_alloc(unsigned long long nbytes, unsigned long long* out_ptr) {
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, nbytes);
  if (e != cudaSuccess) return (int)e;
  e = cudaMemset(p, 0, nbytes);
  *out_ptr = (unsigned long long)p;
  return (int)e;
}

DLB_API int dlb_ipc_free(unsigned long long ptr) { return (int)cudaFree((void*)ptr); }

DLB_API int dlb_ipc_get_handle(unsigned long long ptr, unsigned char* out64) {
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, (void*)ptr);
  if (e != cudaSuccess) return (int)e;
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out64, &h, 64);
  return 0;
}

DLB_API int dlb_ipc_open(const unsigned char* in64, unsigned long long* out_ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, in64, 64);
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  *out_ptr = (unsigned long long)p;
  return (int)e;
}

DLB_API int dlb_ipc_close(unsigned long long ptr) { return (int)cudaIpcCloseMemHandle((void*)ptr); }

DLB_API int dlb_enable_peer_access(int dev, int peer) {
  int can = 0;
  cudaDeviceCanAccessPeer(&can, dev, peer);
  if (!can) return -1;
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
  return (int)e;
}
