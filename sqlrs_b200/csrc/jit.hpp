// sqlrs_b200 — run-time specialisation of the hand-written kernel skeletons (csrc/jit/*.cuh).
//
// A SQL operator's expressions are only known when the plan arrives, so the skeletons are
// compiled per (operator shape, input schema, expressions) with NVRTC for sm_100a, -lineinfo,
// into a cubin that is loaded through the driver entry points (resolved with
// cudaGetDriverEntryPoint — the library does not link libcuda, so it also loads on a box
// without a GPU).  Compiled modules are cached in memory by source text and on disk under
// $SQLRS_B200_JIT_CACHE (default: <library dir>/jit_cache) so that repeated plans and later
// processes skip NVRTC.
#pragma once
#include <string>

#include "common.hpp"

namespace sq {

struct JitKernel;  // opaque: one loaded kernel function

// `generated`: the row program + glue emitted by codegen; `skeleton`: name of an embedded
// skeleton (e.g. "agg"); `kernel_name`: extern "C" __global__ symbol inside it.
JitKernel* jit_get(const std::string& skeleton, const std::string& generated, const std::string& kernel_name);
// compile only (no GPU needed): returns the cubin; used by the build check and the CPU tests
std::string jit_compile_to_cubin(const std::string& skeleton, const std::string& generated, std::string* log);
std::string jit_full_source(const std::string& skeleton, const std::string& generated);
void jit_launch(JitKernel* k, unsigned grid, unsigned block, size_t dyn_smem, cudaStream_t stream, void** args);
int jit_max_blocks_per_sm(JitKernel* k, int block, size_t dyn_smem);

int device_sm_count(int device);

// embedded skeleton sources (generated into embedded_sources.cpp by the build)
const char* embedded_source(const std::string& name);

}  // namespace sq
