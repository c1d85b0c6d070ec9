#include "../../include/integrator2/common/cuda_helper.cuh"
#include "host_context.h"

namespace i2host {
i2_context *context() {
    static i2_context *ctx = nullptr;
    if (!ctx) {
        int dev = 0;
        checkCudaErrors(cudaGetDevice(&dev));
        checkI2Errors(i2_create(&ctx, dev));
        // the reference runs everything on the legacy default stream; so do the drop-in classes, which keeps
        // GpuTimer (events on stream 0) and the CLI's direct copy_d2h calls ordered with the kernels
        checkI2Errors(i2_set_stream(ctx, nullptr));
    }
    return ctx;
}
}  // namespace i2host

void checkI2(int rc, const char *what, const char *file, int line) {
    if (rc) {
        fprintf(stderr, "CUDA error at %s:%d code=%d(%s) \"%s\" \n", file, line, rc, i2_error_string(rc), what);
        exit(EXIT_FAILURE);
    }
}

size_t requestFreeDeviceMemoryAmount() {
    size_t freeBytes = 0, totalBytes = 0;
    checkCudaErrors(cudaMemGetInfo(&freeBytes, &totalBytes));
    printf("GPU memory usage: %5.1f MBytes free out of total %5.1f MBytes\n", freeBytes / 1048576.0f, totalBytes / 1048576.0f);
    return freeBytes;
}
