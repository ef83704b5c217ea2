// gw_common.cuh -- shared device/host helpers of libgenewalk_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/genewalk_b200.h"

namespace gw {

// ---- error plumbing ------------------------------------------------------------------------
void set_error(const char *fmt, ...);

#define GW_CUDA_OK(expr)                                                                       \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            gw::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return GW_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

// after every kernel launch: count it (gw_launch_count) and surface launch-configuration errors
void count_launch();
#define GW_LAUNCH_OK()                       \
    do {                                     \
        gw::count_launch();                  \
        GW_CUDA_OK(cudaGetLastError());      \
    } while (0)

#define GW_REQUIRE(cond, ...)                  \
    do {                                       \
        if (!(cond)) {                         \
            gw::set_error(__VA_ARGS__);        \
            return GW_ERR_INVALID;             \
        }                                      \
    } while (0)

#define GW_TRY(expr)              \
    do {                          \
        int _rc = (expr);         \
        if (_rc != GW_OK) return _rc; \
    } while (0)

// The default memory pool hands unused memory back to the OS at every synchronisation point
// (release threshold 0), and that release waits for the whole device like cudaFree does: with
// several replicates in flight on different streams, one stream's host synchronisation would then
// stall until all other streams have drained (measured: 6-10 s bubbles in the replicate pipeline).
// Keep the pool's memory instead; it is reused by the next call's temporaries.
inline void retain_pool_memory()
{
    static unsigned long long done_mask = 0;   // bit per device; a benign race repeats the call
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    if ((done_mask >> dev) & 1ULL) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        unsigned long long threshold = ~0ULL;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
    }
    done_mask |= 1ULL << dev;
}

// stream-ordered temporary (cudaMallocAsync); freed on the same stream by the destructor
struct Temp {
    void *ptr = nullptr;
    cudaStream_t stream = nullptr;
    cudaError_t alloc(size_t bytes, cudaStream_t s)
    {
        retain_pool_memory();
        stream = s;
        return cudaMallocAsync(&ptr, bytes ? bytes : 1, s);
    }
    template <class T> T *as() const { return reinterpret_cast<T *>(ptr); }
    ~Temp()
    {
        if (ptr) cudaFreeAsync(ptr, stream);
    }
    Temp() = default;
    Temp(const Temp &) = delete;
    Temp &operator=(const Temp &) = delete;
};

inline int num_sms()
{
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess)
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms;
}

template <class T> __host__ __device__ constexpr T ceil_div(T a, T b) { return (a + b - 1) / b; }

// ---- Philox4x32-10 (Salmon et al. SC'11) -----------------------------------------------------
struct u32x4 {
    uint32_t x, y, z, w;
};

__host__ __device__ __forceinline__ u32x4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2,
                                                        uint32_t c3, uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
#ifdef __CUDA_ARCH__
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        const uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    return u32x4{c0, c1, c2, c3};
}

constexpr uint32_t kStreamWalk = 0x57u;    // 'W'
constexpr uint32_t kStreamSgns = 0x53u;    // 'S'
constexpr uint32_t kStreamConfig = 0x43u;  // 'C'

// ---- device-wide primitives implemented in gw_scan_sort.cu -----------------------------------
// exclusive prefix sums; out may alias in; total (optional, device pointer) receives the grand total
int scan_exclusive_i64(const int64_t *in, int64_t *out, int64_t count, int64_t *total,
                       cudaStream_t s);
int scan_exclusive_u32(const uint32_t *in, uint32_t *out, int64_t count, uint32_t *total,
                       cudaStream_t s);
int scan_exclusive_f64(const double *in, double *out, int64_t count, double *total,
                       cudaStream_t s);
// inclusive running minimum from the right: out[i] = min(in[i..count))
int scan_suffix_min_f64(const double *in, double *out, int64_t count, cudaStream_t s);

// segmented variant: (value, segment id) pairs, segment ids non-decreasing with the index
struct SegVal {
    double v;
    long long seg;
};
int scan_segmented_suffix_min(const SegVal *in, SegVal *out, int64_t count, cudaStream_t s);

// stable LSD radix sorts (ascending).  `bits` = number of low key bits that can differ.
// keys/vals are sorted in place (ping-pong temporaries are allocated internally).
int radix_sort_u64(uint64_t *keys, uint32_t *vals /*nullable*/, int64_t count, int bits,
                   cudaStream_t s);
int radix_sort_u32(uint32_t *keys, uint32_t *vals /*nullable*/, int64_t count, int bits,
                   cudaStream_t s);

}  // namespace gw
