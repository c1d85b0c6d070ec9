// Process-wide C-ABI context shared by the drop-in host classes (the reference is single-device with
// process-global constant memory, so one context per process reproduces its lifetime rules).
#pragma once
#include "../../include/i2_abi.h"

namespace i2host {
i2_context *context();   // created on first use on the current CUDA device (device 0 by default)
}
