// Stand-in for the CUDA runtime -- TEST INFRASTRUCTURE (tests/emu/ec_hostbuild.cpp), never shipped, never seen by nvcc.
//
// Just enough of the API for gkr-mimc_b200/csrc/ec/ec.cu's host driver to run on the CPU: "device" memory is host memory with
// poisoned contents and guard bands that are checked on every synchronize and free (an out-of-bounds WRITE of a kernel body or of a
// copy aborts the test with a message), streams are immediate, events count launches.  One fake device.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotReady = 600, cudaErrorInvalidValue = 1 };
typedef struct FakeStream* cudaStream_t;
typedef struct FakeEvent* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1 };

struct FakeStream {
    int dummy;
};
struct FakeEvent {
    double t;
};

namespace fakecuda {
constexpr size_t GUARD = 256;
inline std::map<void*, size_t>& live() {
    static std::map<void*, size_t> m;
    return m;
}
inline double& clock_ms() {
    static double t = 0.0;
    return t;
}
inline void check_guards(const char* when) {
    for (auto& kv : live()) {
        const unsigned char* p = (const unsigned char*)kv.first;
        for (size_t i = 0; i < GUARD; i++)
            if (p[-(ptrdiff_t)GUARD + (ptrdiff_t)i] != 0xC3 || p[kv.second + i] != 0xC3) {
                fprintf(stderr, "fake CUDA runtime: guard band of a %zu-byte allocation overwritten (%s, offset %zd)\n", kv.second, when,
                        p[-(ptrdiff_t)GUARD + (ptrdiff_t)i] != 0xC3 ? (ptrdiff_t)i - (ptrdiff_t)GUARD : (ptrdiff_t)(kv.second + i));
                abort();
            }
    }
}
inline cudaError_t alloc(void** out, size_t bytes) {
    unsigned char* raw = (unsigned char*)aligned_alloc(256, ((bytes + 2 * GUARD + 255) / 256) * 256 + 256);
    if (!raw) return cudaErrorMemoryAllocation;
    memset(raw, 0xC3, GUARD);
    memset(raw + GUARD, 0xA5, bytes);  // poisoned: nothing may rely on zero-initialised device memory
    memset(raw + GUARD + bytes, 0xC3, GUARD);
    *out = raw + GUARD;
    live()[*out] = bytes;
    return cudaSuccess;
}
inline cudaError_t release(void* p) {
    if (!p) return cudaSuccess;
    check_guards("cudaFree");
    auto it = live().find(p);
    if (it == live().end()) {
        fprintf(stderr, "fake CUDA runtime: cudaFree of an unknown pointer\n");
        abort();
    }
    live().erase(it);
    free((unsigned char*)p - GUARD);
    return cudaSuccess;
}
}  // namespace fakecuda

inline cudaError_t cudaGetDeviceCount(int* n) {
    *n = 1;
    return cudaSuccess;
}
inline cudaError_t cudaSetDevice(int d) { return d == 0 ? cudaSuccess : cudaErrorInvalidValue; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) {
    *s = new FakeStream{0};
    return cudaSuccess;
}
inline cudaError_t cudaStreamDestroy(cudaStream_t s) {
    delete s;
    return cudaSuccess;
}
inline cudaError_t cudaStreamSynchronize(cudaStream_t) {
    fakecuda::check_guards("cudaStreamSynchronize");
    return cudaSuccess;
}
inline cudaError_t cudaMalloc(void** p, size_t bytes) { return fakecuda::alloc(p, bytes); }
inline cudaError_t cudaFree(void* p) { return fakecuda::release(p); }
inline cudaError_t cudaMallocHost(void** p, size_t bytes) { return fakecuda::alloc(p, bytes); }
inline cudaError_t cudaFreeHost(void* p) { return fakecuda::release(p); }
inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind, cudaStream_t) {
    memmove(dst, src, bytes);
    return cudaSuccess;
}
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t) {
    memset(p, v, bytes);
    return cudaSuccess;
}
inline cudaError_t cudaEventCreate(cudaEvent_t* e) {
    *e = new FakeEvent{0.0};
    return cudaSuccess;
}
inline cudaError_t cudaEventDestroy(cudaEvent_t e) {
    delete e;
    return cudaSuccess;
}
inline cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    e->t = fakecuda::clock_ms();
    return cudaSuccess;
}
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b) {
    *ms = (float)(b->t - a->t);
    return cudaSuccess;
}
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "fake CUDA runtime error"; }
