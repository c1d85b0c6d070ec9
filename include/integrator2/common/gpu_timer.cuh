// GpuTimer: cudaEvent pair printing "Time for <msg>: %6.3f ms" (/root/reference/src/common/gpu_timer.cuh:14-77).
#ifndef GPU_TIMER_CUH
#define GPU_TIMER_CUH

#include "cuda_helper.cuh"

class GpuTimer {
public:
    GpuTimer() { checkCudaErrors(cudaEventCreate(&t0)); checkCudaErrors(cudaEventCreate(&t1)); }
    ~GpuTimer() { if (t0) cudaEventDestroy(t0); if (t1) cudaEventDestroy(t1); }
    void start(cudaStream_t stream = nullptr) { checkCudaErrors(cudaEventRecord(t0, stream)); }
    float stop(const char *message = nullptr, cudaStream_t stream = nullptr) {
        checkCudaErrors(cudaEventRecord(t1, stream));
        checkCudaErrors(cudaEventSynchronize(t1));
        float ms = 0.f;
        checkCudaErrors(cudaEventElapsedTime(&ms, t0, t1));
        if (message) printf("Time for %s: %6.3f ms\n", message, ms);
        return ms;
    }
private:
    cudaEvent_t t0 = nullptr, t1 = nullptr;
};

#endif  // GPU_TIMER_CUH
