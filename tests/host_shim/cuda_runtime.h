/* TEST INFRASTRUCTURE ONLY — a minimal stand-in for <cuda_runtime.h> that lets g++ compile the DEVICE functions of
 * ncollide_b200/csrc (one pair per thread, no cross-lane operations) for the host, so that their logic and f32 operation
 * order can be checked against the oracle in the GPU-less container (tests/test_device_source_on_host.py).  Nothing in the
 * product includes this file; the product path still needs the CUDA library and a GPU. */
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __launch_bounds__(...)
#define __constant__

struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct uint2 { uint32_t x, y; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }

template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline uint32_t __float_as_uint(float f) { uint32_t i; std::memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline float __uint_as_float(uint32_t i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline uint32_t atomicAdd(uint32_t* p, uint32_t v) { uint32_t o = *p; *p += v; return o; }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
using std::isinf;
using std::isnan;
using std::signbit;

typedef int cudaError_t;
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum { cudaSuccess = 0 };
static inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::malloc(n); return cudaSuccess; }
static inline cudaError_t cudaFree(void* p) { std::free(p); return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t*) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
