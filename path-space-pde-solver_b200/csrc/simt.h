// simt.h -- the kernels are written against plain CUDA C++.  When PSPDE_EMULATE is defined (tests/emu only,
// never in the product build) the same source is compiled for the host against a fiber-based SIMT emulator so
// that index logic and barrier placement can be exercised without a GPU.
#pragma once
#if defined(PSPDE_EMULATE)
#include "simt_emul.h"
#else
#include <cuda_runtime.h>
#define PSPDE_DYN_SMEM(name) extern __shared__ float4 name[]
#endif
