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

int gpus() {
    static int n = -1;
    if (n < 0) {
        n = 1;
        if (const char *e = getenv("I2_GPUS")) {
            int have = 1;
            checkCudaErrors(cudaGetDeviceCount(&have));
            n = atoi(e);
            if (n < 1) n = 1;
            if (n > have) {
                fprintf(stderr, "I2_GPUS=%d but only %d device(s) visible: using %d\n", n, have, have);
                n = have;
            }
        }
    }
    return n;
}

i2_mgpu *mgpu() {
    static i2_mgpu *mg = nullptr;
    if (!mg) checkI2Errors(i2_mgpu_create_local(&mg, gpus(), nullptr));
    return mg;
}

bool gatherToDevice0() {
    const char *e = getenv("I2_GATHER");
    return e && atoi(e) != 0;
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
