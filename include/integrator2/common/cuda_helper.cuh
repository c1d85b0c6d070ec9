// Error-check macro and launch-size helpers with the reference's names and behaviour
// (/root/reference/src/common/cuda_helper.cuh:17-54): a failing CUDA call prints file:line and exits.
#ifndef CUDA_HELPER_CUH
#define CUDA_HELPER_CUH

#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

const int gpuThreads = 256;
const int gpuThreadsMax = 1024;
const int gpuThreads2D = 16;

inline unsigned int blocksForSize(unsigned int n, unsigned int maxThreads = gpuThreads) { return (n + maxThreads - 1) / maxThreads; }

template <typename T>
void check(T result, char const *const func, const char *const file, int const line) {
    if (result) {
        fprintf(stderr, "CUDA error at %s:%d code=%d(%s) \"%s\" \n", file, line, static_cast<unsigned int>(result),
                cudaGetErrorName((cudaError_t)result), func);
        exit(EXIT_FAILURE);
    }
}
#define checkCudaErrors(val) check((val), #val, __FILE__, __LINE__)

// i2_* C-ABI calls return 0 / cudaError_t / negative code: same print-and-exit policy
void checkI2(int rc, const char *what, const char *file, int line);
#define checkI2Errors(val) checkI2((val), #val, __FILE__, __LINE__)

// prints "GPU memory usage: ..." like the reference (src/common/cuda_helper.cu:24-31)
size_t requestFreeDeviceMemoryAmount();

#endif  // CUDA_HELPER_CUH
