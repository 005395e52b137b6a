// Host-simulation stand-in for <cuda_runtime.h> (tools/sim: TEST TOOLING ONLY; see simt.h)
#pragma once
#include "../simt.h"
