// Stand-in for <cuda_runtime.h> on the emulation include path (tests/cuda_emul): TEST INFRASTRUCTURE ONLY.
#pragma once
#include "../cuda_emul.h"
#include "../cuda_runtime_emul.h"
