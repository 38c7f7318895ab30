// common.cuh -- error plumbing shared by the translation units of libsprl_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <cstdarg>
#include <cstdio>
#include <string>

#include "../../include/sprl_b200.h"

namespace sprl {

// Last error message of the calling host thread (sprl_last_error()).
std::string& last_error_ref();
int fail(int code, const char* fmt, ...);

#define SPRL_CUDA(expr)                                                                  \
    do {                                                                                 \
        cudaError_t err__ = (expr);                                                      \
        if (err__ != cudaSuccess)                                                        \
            return ::sprl::fail(SPRL_E_CUDA, "%s failed at %s:%d: %s", #expr, __FILE__,  \
                                __LINE__, cudaGetErrorString(err__));                    \
    } while (0)

// Selects the device; fails loudly when there is none (there is no CPU fallback).
int use_device(int device);

template <typename T>
struct DeviceBuf {
    T* p = nullptr;
    size_t n = 0;
    DeviceBuf() {}
    DeviceBuf(const DeviceBuf&) = delete;
    DeviceBuf& operator=(const DeviceBuf&) = delete;
    ~DeviceBuf() { release(); }
    cudaError_t alloc(size_t count) {
        release();
        n = count;
        if (count == 0) return cudaSuccess;
        return cudaMalloc((void**)&p, count * sizeof(T));
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
};

// NVTX range of a host entry point (the reference has a stopwatch only; here every C ABI call that
// enqueues or waits for device work shows up as a named range in a timeline profiler; header-only nvtx3, free when no
// profiler is attached).
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace sprl
