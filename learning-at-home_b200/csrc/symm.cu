// symm.cu — symmetric heap plumbing: a cudaMalloc'ed arena per rank whose CUDA-IPC handle is exchanged through
// torch.distributed (python side, parallel/symmetric.py) and mapped into every peer process, so kernels can issue
// plain ld/st/red to peer HBM over NVLink 5 / NVSwitch.  (The reference moves the same bytes through pickled TCP
// messages: /root/reference/lib/utils/connection.py:22-51.)
#include <cuda_runtime.h>
#include <string.h>

extern "C" {

int lah_symm_alloc(unsigned long long bytes, void** out) {
    void* p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return -(int)e;
    e = cudaMemset(p, 0, bytes);
    if (e != cudaSuccess) return -(int)e;
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) return -(int)e;
    *out = p;
    return 0;
}

int lah_symm_free(void* p) { return -(int)cudaFree(p); }

int lah_symm_get_handle(void* p, char* out64) {
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) return -(int)e;
    static_assert(sizeof(h) == 64, "ipc handle size");
    memcpy(out64, &h, 64);
    return 0;
}

int lah_symm_open_handle(const char* in64, void** out) {
    cudaIpcMemHandle_t h;
    memcpy(&h, in64, 64);
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) return -(int)e;
    *out = p;
    return 0;
}

int lah_symm_close_handle(void* p) { return -(int)cudaIpcCloseMemHandle(p); }

int lah_device_sync() { return -(int)cudaDeviceSynchronize(); }

}  // extern "C"
