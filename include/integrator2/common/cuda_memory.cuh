// Typed wrappers over cudaMalloc / cudaMemcpyAsync with the reference's names
// (/root/reference/src/common/cuda_memory.cuh); the CLI and user code call copy_d2h etc. directly.
#ifndef CUDA_MEMORY_CUH
#define CUDA_MEMORY_CUH

#include "cuda_helper.cuh"

template <class T> void allocate_device(T **ptr, size_t count) {
    if (count) checkCudaErrors(cudaMalloc(ptr, count * sizeof(T)));
    else printf("Zero device memory allocation requested\n");
}
template <class T> void allocate_host(T **ptr, size_t count) {
    if (count) checkCudaErrors(cudaMallocHost(ptr, count * sizeof(T)));
    else printf("Zero host memory allocation requested\n");
}
template <class T> void free_device(T *ptr) { if (ptr) checkCudaErrors(cudaFree(ptr)); }
template <class T> void free_host(T *ptr) { if (ptr) checkCudaErrors(cudaFreeHost(ptr)); }
template <class T> void zero_value_device(T *ptr, size_t count, cudaStream_t stream = nullptr) {
    checkCudaErrors(cudaMemsetAsync(ptr, 0, count * sizeof(T), stream));
}

namespace i2host {
template <class T> void copy_any(const T *src, const T *dst, size_t count, cudaMemcpyKind kind, const char *what, cudaStream_t stream) {
    if (count) checkCudaErrors(cudaMemcpyAsync((void *)dst, (const void *)src, count * sizeof(T), kind, stream));
    else printf("Zero %s memory copy requested\n", what);
}
}  // namespace i2host

template <class T> void copy_h2d(const T *src, const T *dst, size_t count, cudaStream_t stream = nullptr) {
    i2host::copy_any(src, dst, count, cudaMemcpyHostToDevice, "host-to-device", stream);
}
// (the reference prints "host-to-device" for a zero-size d2h copy as well: cuda_memory.cuh:109)
template <class T> void copy_d2h(const T *src, const T *dst, size_t count, cudaStream_t stream = nullptr) {
    i2host::copy_any(src, dst, count, cudaMemcpyDeviceToHost, "host-to-device", stream);
}
template <class T> void copy_d2d(const T *src, const T *dst, size_t count, cudaStream_t stream = nullptr) {
    i2host::copy_any(src, dst, count, cudaMemcpyDeviceToDevice, "device-to-device", stream);
}

#endif  // CUDA_MEMORY_CUH
