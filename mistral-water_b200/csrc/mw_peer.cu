// mw_peer.cu -- peer-memory plumbing of the multi-GPU tile set (the reference has no counterpart:
// nothing in Scripts/FFTMesh.cs couples two meshes, and it runs on one device).
//
// One process per GPU.  Every rank exports the buffer its peers should write its gathered slots into (CUDA IPC), opens
// the peers' buffers FROM ITS OWN DEVICE (cudaIpcMemLazyEnablePeerAccess maps the exporter's memory for direct NVLink
// access by the opening device), and the all-gather of the final float buffers becomes one asynchronous copy per
// peer: the copy engines move the slot over NVLink / NVSwitch while the SMs run the next frame.
#include <string.h>

#include "mw_common.cuh"

extern "C" int mw_peer_export(const void* dev_ptr, void* handle64, uint64_t* offset)
{
    if (!dev_ptr || !handle64 || !offset) { mw_set_error("mw_peer_export: null argument"); return MW_E_INVALID_ARG; }
    static_assert(sizeof(cudaIpcMemHandle_t) == MW_PEER_HANDLE_BYTES, "handle size");
    cudaPointerAttributes at;
    MW_CUDA(cudaPointerGetAttributes(&at, dev_ptr));
    if (at.type != cudaMemoryTypeDevice) { mw_set_error("mw_peer_export: not a device pointer"); return MW_E_INVALID_ARG; }
    MW_CUDA(cudaSetDevice(at.device));
    // the handle names the whole allocation the pointer lives in: report where inside it the pointer is
    CUdeviceptr base = 0;
    size_t size = 0;
    typedef CUresult (*RangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    static RangeFn range = nullptr;
    if (!range) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        MW_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &p, cudaEnableDefault, &q));
        if (!p || q != cudaDriverEntryPointSuccess) { mw_set_error("cuMemGetAddressRange not available"); return MW_E_CUDA; }
        range = (RangeFn)p;
    }
    if (range(&base, &size, (CUdeviceptr)dev_ptr) != CUDA_SUCCESS) { mw_set_error("cuMemGetAddressRange failed"); return MW_E_CUDA; }
    cudaIpcMemHandle_t h;
    MW_CUDA(cudaIpcGetMemHandle(&h, (void*)base));
    memcpy(handle64, &h, sizeof(h));
    *offset = (uint64_t)((CUdeviceptr)dev_ptr - base);
    return MW_OK;
}

extern "C" int mw_peer_open(int device, const void* handle64, void** base)
{
    if (!handle64 || !base) { mw_set_error("mw_peer_open: null argument"); return MW_E_INVALID_ARG; }
    *base = nullptr;
    MW_CUDA(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, sizeof(h));
    MW_CUDA(cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess));
    return MW_OK;
}

extern "C" int mw_peer_close(int device, void* base)
{
    if (!base) return MW_OK;
    MW_CUDA(cudaSetDevice(device));
    MW_CUDA(cudaIpcCloseMemHandle(base));
    return MW_OK;
}

extern "C" int mw_peer_copy(void* dst, const void* src, uint64_t bytes, void* cuda_stream)
{
    if (!dst || !src) { mw_set_error("mw_peer_copy: null argument"); return MW_E_INVALID_ARG; }
    MW_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)cuda_stream));
    return MW_OK;
}
