// Shared device-side helpers for the RHCCQ hot-path kernels (sm_100a).
//
// Every kernel in this directory is written against a small vocabulary
// (threadIdx/blockIdx, __syncthreads, shared-memory atomics, the block scan and
// reduce helpers below, IEEE double ops with explicit rounding).  The product
// build compiles it with nvcc for sm_100a.  tests/emu/ compiles the same
// sources with g++ and -DRHCCQ_HOST_EMU: one host "thread" per CTA, blocks run
// one after another.  That build exists only so that the CPU test tier can
// check the kernels' logic against the oracle without a GPU; it is never
// loaded by the package (roibasedimagecompression_b200/_lib.py loads only the
// nvcc-built library and raises when it is missing).
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifdef RHCCQ_HOST_EMU
// ---------------------------------------------------------------- host emulation shims
#include <string.h>
#include <stdlib.h>
#include <math.h>
struct rhccq_emu_dim3 { unsigned x, y, z; };
struct int2 { int x, y; };
struct uint4 { unsigned x, y, z, w; };
struct float4 { float x, y, z, w; };
extern rhccq_emu_dim3 threadIdx, blockIdx, blockDim, gridDim;
extern unsigned char* rhccq_emu_dyn_smem;
void rhccq_emu_prepare_smem(size_t bytes);
typedef void* cudaStream_t;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __shared__ static
#define __restrict__
#define __launch_bounds__(...)
#define RHCCQ_DYN_SMEM(name) unsigned char* name = rhccq_emu_dyn_smem
static inline void __syncthreads() {}
static inline void __syncwarp() {}
static inline void __threadfence() {}
static inline void __threadfence_block() {}
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }
template <class T> static inline T atomicCAS(T* p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __fma_rn(double a, double b, double c) { return fma(a, b, c); }
static inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
static inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fsub_rn(float a, float b) { return a - b; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline unsigned __vabsdiffu4(unsigned a, unsigned b) {
    unsigned r = 0;
    for (int s = 0; s < 32; s += 8) {
        int x = (a >> s) & 255, y = (b >> s) & 255;
        r |= (unsigned)(x > y ? x - y : y - x) << s;
    }
    return r;
}
static inline unsigned __dp4a(unsigned a, unsigned b, unsigned c) {
    for (int s = 0; s < 32; s += 8) c += ((a >> s) & 255) * ((b >> s) & 255);
    return c;
}
#define RHCCQ_EMU_BYTEWISE(name, expr)                                        \
    static inline unsigned name(unsigned a, unsigned b) {                     \
        unsigned r = 0;                                                       \
        for (int s = 0; s < 32; s += 8) {                                     \
            int x = (a >> s) & 255, y = (b >> s) & 255;                       \
            r |= ((unsigned)(expr) & 255u) << s;                              \
        }                                                                     \
        return r;                                                             \
    }
RHCCQ_EMU_BYTEWISE(__vsubus4, x > y ? x - y : 0)
RHCCQ_EMU_BYTEWISE(__vminu4, x < y ? x : y)
RHCCQ_EMU_BYTEWISE(__vmaxu4, x > y ? x : y)
RHCCQ_EMU_BYTEWISE(__vadd4, x + y)
#define RHCCQ_LAUNCH(kern, grid, block, smem, stream, ...)                          \
    do {                                                                            \
        gridDim.x = (unsigned)(grid); blockDim.x = 1; threadIdx.x = 0;              \
        for (unsigned rhccq_b_ = 0; rhccq_b_ < (unsigned)(grid); ++rhccq_b_) {      \
            blockIdx.x = rhccq_b_;                                                  \
            rhccq_emu_prepare_smem((size_t)(smem));                                 \
            kern(__VA_ARGS__);                                                      \
        }                                                                           \
    } while (0)
#else
// ---------------------------------------------------------------- real CUDA
#include <cuda_runtime.h>
#define RHCCQ_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define RHCCQ_LAUNCH(kern, grid, block, smem, stream, ...) \
    kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

#ifdef RHCCQ_HOST_EMU
#define RHCCQ_LANE 0
#define RHCCQ_WARP 0
#define RHCCQ_NWARPS 1
#define RHCCQ_WARP_SIZE 1
// warp votes / shuffles of a one-lane warp
static inline unsigned rhccq_ballot(int pred) { return pred ? 1u : 0u; }
static inline int rhccq_any(int pred) { return pred; }
template <class T> static inline T rhccq_shfl(T v, int) { return v; }
template <class T> static inline T rhccq_shfl_xor(T v, int) { return v; }
static inline unsigned rhccq_lanemask_lt() { return 0u; }
#else
#define RHCCQ_LANE ((int)(threadIdx.x & 31))
#define RHCCQ_WARP ((int)(threadIdx.x >> 5))
#define RHCCQ_NWARPS ((int)(blockDim.x >> 5))
#define RHCCQ_WARP_SIZE 32
__device__ __forceinline__ unsigned rhccq_ballot(int pred) { return __ballot_sync(0xffffffffu, pred); }
__device__ __forceinline__ int rhccq_any(int pred) { return __any_sync(0xffffffffu, pred); }
template <class T> __device__ __forceinline__ T rhccq_shfl(T v, int src) { return __shfl_sync(0xffffffffu, v, src); }
template <class T> __device__ __forceinline__ T rhccq_shfl_xor(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ unsigned rhccq_lanemask_lt() { return (1u << (threadIdx.x & 31)) - 1u; }
#endif

#define RHCCQ_PAR_FOR(i, n) for (int i = (int)threadIdx.x; i < (int)(n); i += (int)blockDim.x)

#define RHCCQ_MAX_WARPS 32

// ---------------------------------------------------------------- colour keys
// A colour is the 24-bit key R<<16|G<<8|B: ascending key order is the
// lexicographic (R,G,B) order np.unique(axis=0) produces
// (/root/reference/encoder/compression/clustering.py:21-23).
__device__ __forceinline__ uint32_t rhccq_pack_rgb(unsigned r, unsigned g, unsigned b) {
    return (r << 16) | (g << 8) | b;
}
__device__ __forceinline__ int rhccq_key_r(uint32_t k) { return (int)((k >> 16) & 255u); }
__device__ __forceinline__ int rhccq_key_g(uint32_t k) { return (int)((k >> 8) & 255u); }
__device__ __forceinline__ int rhccq_key_b(uint32_t k) { return (int)(k & 255u); }

// Squared euclidean distance of two packed colours (exact, <= 195075).
__device__ __forceinline__ int rhccq_d2(uint32_t a, uint32_t b) {
    unsigned d = __vabsdiffu4(a, b);
    return (int)__dp4a(d, d, 0u);
}

__device__ __forceinline__ int rhccq_isqrt(int v) {       // floor(sqrt(v)), v >= 0
    int r = 0;
    while ((long long)(r + 1) * (r + 1) <= (long long)v) ++r;
    return r;
}

__device__ __forceinline__ int rhccq_next_pow2(int v) {
    int p = 1;
    while (p < v) p <<= 1;
    return p;
}

// ---------------------------------------------------------------- block collectives
// All threads of the CTA must call these (they contain barriers).  `scratch`
// holds RHCCQ_MAX_WARPS elements of T in shared memory.
#ifdef RHCCQ_HOST_EMU
template <class T> static inline T rhccq_block_sum(T v, T*) { return v; }
template <class T> static inline T rhccq_block_min(T v, T*) { return v; }
template <class T> static inline T rhccq_block_max(T v, T*) { return v; }
static inline int rhccq_block_or(int v, int*) { return v; }
// exclusive scan of one value per thread; *total receives the block sum
template <class T> static inline T rhccq_block_excl_scan(T v, T* total, T*) { *total = v; return (T)0; }
#else
template <class T> __device__ __forceinline__ T rhccq_shfl_down(T v, int d) {
    return __shfl_down_sync(0xffffffffu, v, d);
}
template <class T> __device__ __forceinline__ T rhccq_shfl_up(T v, int d) {
    return __shfl_up_sync(0xffffffffu, v, d);
}
template <class T, class Op>
__device__ __forceinline__ T rhccq_block_reduce(T v, T* scratch, Op op, T identity) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v = op(v, rhccq_shfl_down(v, d));
    __syncthreads();                       // scratch may still be read by a previous call
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0) {
        v = lane < nwarp ? scratch[lane] : identity;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v = op(v, rhccq_shfl_down(v, d));
        if (lane == 0) scratch[0] = v;
    }
    __syncthreads();
    return scratch[0];
}
template <class T> struct rhccq_op_sum { __device__ T operator()(T a, T b) const { return a + b; } };
template <class T> struct rhccq_op_min { __device__ T operator()(T a, T b) const { return b < a ? b : a; } };
template <class T> struct rhccq_op_max { __device__ T operator()(T a, T b) const { return b > a ? b : a; } };
template <class T> __device__ __forceinline__ T rhccq_block_sum(T v, T* s) {
    return rhccq_block_reduce(v, s, rhccq_op_sum<T>(), (T)0);
}
template <class T> __device__ __forceinline__ T rhccq_block_min(T v, T* s) {
    return rhccq_block_reduce(v, s, rhccq_op_min<T>(), v);
}
template <class T> __device__ __forceinline__ T rhccq_block_max(T v, T* s) {
    return rhccq_block_reduce(v, s, rhccq_op_max<T>(), v);
}
__device__ __forceinline__ int rhccq_block_or(int v, int* s) {
    return __syncthreads_or(v);
}
template <class T>
__device__ __forceinline__ T rhccq_block_excl_scan(T v, T* total, T* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nwarp = (blockDim.x + 31) >> 5;
    T incl = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = rhccq_shfl_up(incl, d);
        if (lane >= d) incl += o;
    }
    __syncthreads();
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        T w = lane < nwarp ? scratch[lane] : (T)0;
        T wi = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            T o = rhccq_shfl_up(wi, d);
            if (lane >= d) wi += o;
        }
        scratch[lane] = wi - w;             // exclusive warp offsets; lane 31 of the last warp gives the total below
        if (lane == 31) scratch[32] = wi;   // block total (scratch holds RHCCQ_MAX_WARPS + 1)
    }
    __syncthreads();
    T r = scratch[warp] + incl - v;
    *total = scratch[32];
    return r;
}
#endif

// Exclusive prefix sum over an array in (shared or global) memory, in place:
// a[i] <- sum_{j<i} a[j]; returns the total.  `scratch` holds RHCCQ_MAX_WARPS+1
// elements.  Every thread processes a contiguous chunk so that the order of
// additions is fixed (integer types only, so the order is immaterial anyway).
template <class T>
__device__ __forceinline__ T rhccq_block_excl_scan_array(T* a, int n, T* scratch) {
    const int nt = (int)blockDim.x, t = (int)threadIdx.x;
    const int per = (n + nt - 1) / nt;
    const int lo = t * per < n ? t * per : n;
    const int hi = lo + per < n ? lo + per : n;
    T s = 0;
    for (int i = lo; i < hi; ++i) s += a[i];
    T total;
    T base = rhccq_block_excl_scan<T>(s, &total, scratch);
    for (int i = lo; i < hi; ++i) { T v = a[i]; a[i] = base; base += v; }
    __syncthreads();
    return total;
}

// ---------------------------------------------------------------- block bitonic sort
// Ascending sort of a[0..np2) (np2 a power of two; pad with the maximum value).
// a may live in shared or global memory; every thread of the CTA must call.
template <class T>
__device__ __forceinline__ void rhccq_block_bitonic_sort(T* a, int np2) {
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = (int)threadIdx.x; t < (np2 >> 1); t += (int)blockDim.x) {
                const int i = 2 * t - (t & (j - 1));
                const int l = i + j;
                const bool up = (i & k) == 0;
                const T x = a[i], y = a[l];
                if ((y < x) == up) { a[i] = y; a[l] = x; }
            }
            __syncthreads();
        }
    }
}

// Carve typed arrays out of a byte workspace (shared or global), 16-byte aligned.
struct rhccq_carver {
    unsigned char* p;
    __device__ __forceinline__ explicit rhccq_carver(unsigned char* base) : p(base) {}
    template <class T> __device__ __forceinline__ T* take(size_t count) {
        T* r = reinterpret_cast<T*>(p);
        p += (count * sizeof(T) + 15) & ~(size_t)15;
        return r;
    }
};
__host__ __device__ static inline size_t rhccq_carve_bytes(size_t count, size_t elem) { return (count * elem + 15) & ~(size_t)15; }
