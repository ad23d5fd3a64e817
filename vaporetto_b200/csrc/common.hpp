// vaporetto_b200 — shared host-side definitions (error plumbing, small utilities).
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>

namespace vpt {

// Status codes of the C ABI (include/vaporetto_b200.h).  0..6 mirror the variants of
// the reference's VaporettoError (vaporetto/src/errors.rs:15-38).
enum Status : int {
    kOk = 0,
    kInvalidModel = 1,
    kInvalidArgument = 2,
    kInvalidSentence = 3,
    kDecodeError = 4,
    kIoError = 5,
    kCudaError = 16,
    kUnsupported = 17,
    kInternal = 18,
};

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& m);
const char* last_error();

static inline int32_t wrapping_add(int32_t a, int32_t b) { return int32_t(uint32_t(a) + uint32_t(b)); }

// 64-bit finalizer (splitmix64 / murmur3 fmix style); shared by the host table builder and the kernels.
#if defined(__CUDACC__)
#define VPT_HD __host__ __device__ __forceinline__
#else
#define VPT_HD inline
#endif

VPT_HD uint64_t mix64(uint64_t x) {
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

}  // namespace vpt
