// Kernel-launch macro: real <<<>>> launch under nvcc; under -DDFVO_HOSTSIM (CPU test build, see
// tests/hostsim/cuda_hostsim.h) the same kernel body runs inside a fiber-based emulator.
#pragma once
#ifdef DFVO_HOSTSIM
#include "cuda_hostsim.h"
#else
#include <cuda_runtime.h>
#include <atomic>
namespace dfvo { extern std::atomic<long long> g_launch_count; }
#define DFVO_LAUNCH(kern, grid, block, smem, stream, ...) \
  do { ++dfvo::g_launch_count; kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); } while (0)
#define DFVO_DYN_SMEM(type, name) extern __shared__ __align__(16) unsigned char _dyn_smem_raw[]; \
  type* name = reinterpret_cast<type*>(_dyn_smem_raw)
#endif
