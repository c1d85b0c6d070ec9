// Process-wide C-ABI handles shared by the drop-in host classes (the reference is single-device with
// process-global constant memory, so one context per process reproduces its lifetime rules).
#pragma once
#include "../../include/i2_abi.h"

namespace i2host {
i2_context *context();   // created on first use on the current CUDA device (device 0 by default)
// Multi-GPU mode of the drop-in classes: env I2_GPUS=N (N > 1, clamped to the visible devices) makes Evaluator3D::runAllPairs
// shard the task lists over N GPUs through i2_mgpu_* (one process, NCCL inside the library).  1 = the reference's behaviour.
int gpus();
i2_mgpu *mgpu();         // created on first use (i2_mgpu_create_local over devices 0 .. gpus()-1)
bool gatherToDevice0();  // env I2_GATHER=1: after a multi-GPU run collect tasks and results in the evaluator's device vectors
}
