// common.h — error handling + small RAII helpers shared by all translation units.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/whisper_b200.h"

struct WbError : std::runtime_error {
    int code;
    WbError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void wb_set_error(const std::string& msg);

#define WB_THROW(code, ...)                                   \
    do {                                                      \
        char _b[512];                                         \
        snprintf(_b, sizeof(_b), __VA_ARGS__);                \
        throw WbError((code), _b);                            \
    } while (0)

#define CUDA_CHECK(expr)                                                                   \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess)                                                             \
            WB_THROW(WB_ECUDA, "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, \
                     __LINE__, cudaGetErrorString(_e));                                    \
    } while (0)

#define WB_REQUIRE(cond, code, ...)              \
    do {                                         \
        if (!(cond)) WB_THROW((code), __VA_ARGS__); \
    } while (0)

// Grow-only device buffer.
template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    void reserve(size_t n) {
        if (n <= cap) return;
        release();
        CUDA_CHECK(cudaMalloc(&p, n * sizeof(T)));
        cap = n;
    }
    void reserve_zero(size_t n) {
        if (n <= cap) return;
        reserve(n);
        CUDA_CHECK(cudaMemset(p, 0, n * sizeof(T)));
        CUDA_CHECK(cudaStreamSynchronize(cudaStreamLegacy));      // the users run on non-blocking streams
    }
};

// WB_BLOCKING_SYNC=1: host waits yield the core instead of spinning (many contexts per host, e.g. 8 GPUs
// x 4 batches in flight on a box with fewer cores than waiting threads).
inline bool wb_blocking_sync() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("WB_BLOCKING_SYNC"); v = (e && e[0] == '1') ? 1 : 0; }
    return v == 1;
}
struct CudaEvent {
    cudaEvent_t e = nullptr;
    CudaEvent() { cudaEventCreateWithFlags(&e, wb_blocking_sync() ? cudaEventBlockingSync : cudaEventDefault); }
    ~CudaEvent() { if (e) cudaEventDestroy(e); }
    CudaEvent(const CudaEvent&) = delete;
    CudaEvent& operator=(const CudaEvent&) = delete;
};

// stream synchronisation that honours WB_BLOCKING_SYNC
inline cudaError_t wb_stream_sync(cudaStream_t st) {
    if (!wb_blocking_sync()) return cudaStreamSynchronize(st);
    CudaEvent ev;
    cudaError_t e = cudaEventRecord(ev.e, st);
    return e != cudaSuccess ? e : cudaEventSynchronize(ev.e);
}
static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }
