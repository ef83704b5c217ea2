// gw_scan_sort.cu -- device-wide scan and stable LSD radix sort used by the graph builder, the
// configuration-model randomizer, the alias-table builder and the statistics kernels.
// These are the small supporting primitives of the hot path (sizes: 2E stubs, A adjacency entries,
// nreps*A null similarities); the dominant kernels live in gw_walk.cu / gw_sgns.cu.
#include <math.h>
#include <stdarg.h>

#include <atomic>

#include "gw_common.cuh"

namespace gw {

// ---- error string (thread-local) -------------------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char *last_error() { return g_err; }

// kernels launched by this library since load (a statistic for bench.py's gpu_launches)
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long long launches() { return g_launches.load(std::memory_order_relaxed); }

// =============================================================================================
// scan: tiles of 256 threads x 8 items, recursive over the tile totals
// =============================================================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

struct OpSum {
    template <class T> __device__ __forceinline__ T operator()(T a, T b) const { return a + b; }
};
struct OpMin {
    template <class T> __device__ __forceinline__ T operator()(T a, T b) const
    {
        return b < a ? b : a;
    }
};

template <class T> __device__ __forceinline__ T shfl_up_t(T v, int delta)
{
    static_assert(sizeof(T) % 4 == 0, "shuffle helper moves 32-bit words");
    T r;
    const uint32_t *src = reinterpret_cast<const uint32_t *>(&v);
    uint32_t *dst = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (int k = 0; k < (int)(sizeof(T) / 4); ++k) dst[k] = __shfl_up_sync(0xffffffffu, src[k], delta);
    return r;
}

// Scans one tile; writes the tile-local scan to out and the tile total to tile_totals[blockIdx.x].
// kReverse scans from the right end (logical index i -> count-1-i); kInclusive selects inclusive.
template <class T, class Op, bool kReverse, bool kInclusive>
__global__ void __launch_bounds__(kScanThreads)
scan_tile_kernel(const T *in, T *out, int64_t count, T *tile_totals,  // in may alias out
                 T identity, Op op)
{
    __shared__ T warp_tot[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    T item[kScanItems];
    T run = identity;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k;
        T v = identity;
        if (i < count) v = in[kReverse ? count - 1 - i : i];
        run = op(run, v);
        item[k] = run;  // inclusive within the thread
    }
    // block-wide exclusive scan of the thread totals
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T incl = run;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
        T o = shfl_up_t(incl, dlt);
        if (lane >= dlt) incl = op(o, incl);
    }
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    T warp_off = identity;
    for (int w = 0; w < warp; ++w) warp_off = op(warp_off, warp_tot[w]);
    T excl = shfl_up_t(incl, 1);
    if (lane == 0) excl = identity;
    const T thread_off = op(warp_off, excl);
    if (threadIdx.x == kScanThreads - 1 && tile_totals) tile_totals[blockIdx.x] = op(warp_off, incl);
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + k;
        if (i < count) {
            T v;
            if (kInclusive) v = op(thread_off, item[k]);
            else v = (k == 0) ? thread_off : op(thread_off, item[k - 1]);
            out[kReverse ? count - 1 - i : i] = v;
        }
    }
}

template <class T, class Op, bool kReverse>
__global__ void __launch_bounds__(kScanThreads)
scan_add_offsets_kernel(T *out, int64_t count, const T *tile_offsets, Op op)
{
    const T off = tile_offsets[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * kScanTile;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        const int64_t i = base + (int64_t)k * kScanThreads + threadIdx.x;
        if (i < count) {
            const int64_t j = kReverse ? count - 1 - i : i;
            out[j] = op(off, out[j]);
        }
    }
}

template <class T, class Op>
__global__ void store_total_kernel(const T *offsets_last, const T *totals_last, T *total, Op op)
{
    *total = op(*offsets_last, *totals_last);
}

template <class T, class Op, bool kReverse, bool kInclusive>
static int scan_impl(const T *in, T *out, int64_t count, T *total, T identity, Op op,
                     cudaStream_t s)
{
    if (count <= 0) {
        if (total) GW_CUDA_OK(cudaMemsetAsync(total, 0, sizeof(T), s));  // sum scans only
        return GW_OK;
    }
    const int64_t tiles = ceil_div<int64_t>(count, kScanTile);
    Temp totals;
    GW_CUDA_OK(totals.alloc(sizeof(T) * (size_t)tiles * 2, s));
    T *tot = totals.as<T>();
    T *off = tot + tiles;
    scan_tile_kernel<T, Op, kReverse, kInclusive>
        <<<(unsigned)tiles, kScanThreads, 0, s>>>(in, out, count, tot, identity, op);
    GW_LAUNCH_OK();
    if (tiles > 1) {
        // exclusive scan of the tile totals (recursion depth <= 3 for any count < 2^33)
        GW_TRY((scan_impl<T, Op, false, false>(tot, off, tiles, nullptr, identity, op, s)));
        scan_add_offsets_kernel<T, Op, kReverse><<<(unsigned)tiles, kScanThreads, 0, s>>>(out, count, off, op);
        GW_LAUNCH_OK();
        if (total) {
            store_total_kernel<T, Op><<<1, 1, 0, s>>>(off + tiles - 1, tot + tiles - 1, total, op);
            GW_LAUNCH_OK();
        }
    } else if (total) {
        GW_CUDA_OK(cudaMemcpyAsync(total, tot, sizeof(T), cudaMemcpyDeviceToDevice, s));
    }
    return GW_OK;
}

int scan_exclusive_i64(const int64_t *in, int64_t *out, int64_t count, int64_t *total, cudaStream_t s)
{
    return scan_impl<int64_t, OpSum, false, false>(in, out, count, total, (int64_t)0, OpSum(), s);
}
int scan_exclusive_u32(const uint32_t *in, uint32_t *out, int64_t count, uint32_t *total, cudaStream_t s)
{
    return scan_impl<uint32_t, OpSum, false, false>(in, out, count, total, 0u, OpSum(), s);
}
int scan_exclusive_f64(const double *in, double *out, int64_t count, double *total, cudaStream_t s)
{
    return scan_impl<double, OpSum, false, false>(in, out, count, total, 0.0, OpSum(), s);
}
int scan_suffix_min_f64(const double *in, double *out, int64_t count, cudaStream_t s)
{
    return scan_impl<double, OpMin, true, true>(in, out, count, nullptr, HUGE_VAL, OpMin(), s);
}

// segmented running minimum from the right: elements carry their (non-decreasing) segment id;
// scanning right-to-left, the accumulator restarts whenever the segment id changes.
struct OpSegMin {
    __device__ __forceinline__ SegVal operator()(SegVal acc, SegVal x) const
    {
        // acc = combination of elements further right (earlier in scan order), x = next element
        if (x.seg < 0) return acc;  // identity element
        SegVal r;
        r.seg = x.seg;
        r.v = (acc.seg == x.seg && acc.v < x.v) ? acc.v : x.v;
        return r;
    }
};
int scan_segmented_suffix_min(const SegVal *in, SegVal *out, int64_t count, cudaStream_t s)
{
    SegVal id;
    id.v = HUGE_VAL;
    id.seg = -1;
    return scan_impl<SegVal, OpSegMin, true, true>(in, out, count, nullptr, id, OpSegMin(), s);
}

// =============================================================================================
// radix sort: stable LSD, 8-bit digits; per pass: tile histograms -> scan -> ranked scatter
// =============================================================================================
constexpr int kSortThreads = 256;
constexpr int kSortRounds = 16;
constexpr int kSortTile = kSortThreads * kSortRounds;  // 4096 keys per CTA

template <class K>
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const K *__restrict__ keys, int64_t count, int shift, uint32_t *hist /*[256][tiles]*/,
                  int64_t tiles)
{
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int r = 0; r < kSortRounds; ++r) {
        const int64_t i = base + (int64_t)r * kSortThreads + threadIdx.x;
        if (i < count) atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(int64_t)threadIdx.x * tiles + blockIdx.x] = h[threadIdx.x];
}

template <class K, bool kHasVals>
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const K *__restrict__ keys_in, const uint32_t *__restrict__ vals_in,
                     K *__restrict__ keys_out, uint32_t *__restrict__ vals_out, int64_t count,
                     int shift, const uint32_t *__restrict__ hist_scanned, int64_t tiles)
{
    __shared__ uint32_t base[256];                       // next output offset per digit
    __shared__ uint32_t wcount[kSortThreads / 32][256];  // per-warp digit counts -> warp offsets
    __shared__ uint32_t round_tot[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    base[threadIdx.x] = hist_scanned[(int64_t)threadIdx.x * tiles + blockIdx.x];
    const int64_t tile0 = (int64_t)blockIdx.x * kSortTile;
    for (int r = 0; r < kSortRounds; ++r) {
#pragma unroll
        for (int w = 0; w < kSortThreads / 32; ++w) wcount[w][threadIdx.x] = 0;
        __syncthreads();
        const int64_t i = tile0 + (int64_t)r * kSortThreads + threadIdx.x;
        const bool valid = i < count;
        K key = 0;
        uint32_t digit = 0, rank = 0;
        const unsigned vmask = __ballot_sync(0xffffffffu, valid);
        if (valid) {
            key = keys_in[i];
            digit = (uint32_t)(key >> shift) & 255u;
            const unsigned peers = __match_any_sync(vmask, digit);
            rank = __popc(peers & ((1u << lane) - 1u));
            if (rank == 0) wcount[warp][digit] = __popc(peers);
        }
        __syncthreads();
        {   // thread d turns the per-warp counts of digit d into exclusive offsets
            uint32_t run = 0;
#pragma unroll
            for (int w = 0; w < kSortThreads / 32; ++w) {
                const uint32_t c = wcount[w][threadIdx.x];
                wcount[w][threadIdx.x] = run;
                run += c;
            }
            round_tot[threadIdx.x] = run;
        }
        __syncthreads();
        if (valid) {
            const uint32_t dst = base[digit] + wcount[warp][digit] + rank;
            keys_out[dst] = key;
            if (kHasVals) vals_out[dst] = vals_in[i];
        }
        __syncthreads();
        base[threadIdx.x] += round_tot[threadIdx.x];
        // (the zeroing at the top of the next round is ordered by the barrier that follows it)
        if (tile0 + (int64_t)(r + 1) * kSortThreads >= count) break;
    }
}

template <class K>
static int radix_sort_impl(K *keys, uint32_t *vals, int64_t count, int bits, cudaStream_t s)
{
    if (count <= 1) return GW_OK;
    if (count >= ((int64_t)1 << 32)) {
        set_error("radix sort: count %lld exceeds 2^32-1", (long long)count);
        return GW_ERR_UNSUPPORTED;
    }
    if (bits <= 0) return GW_OK;
    if (bits > (int)sizeof(K) * 8) bits = (int)sizeof(K) * 8;
    const int passes = ceil_div(bits, 8);
    const int64_t tiles = ceil_div<int64_t>(count, kSortTile);
    Temp tk, tv, th;
    GW_CUDA_OK(tk.alloc(sizeof(K) * (size_t)count, s));
    if (vals) GW_CUDA_OK(tv.alloc(sizeof(uint32_t) * (size_t)count, s));
    GW_CUDA_OK(th.alloc(sizeof(uint32_t) * 256 * (size_t)tiles, s));
    K *kin = keys, *kout = tk.as<K>();
    uint32_t *vin = vals, *vout = tv.as<uint32_t>();
    uint32_t *hist = th.as<uint32_t>();
    for (int p = 0; p < passes; ++p) {
        const int shift = p * 8;
        radix_hist_kernel<K><<<(unsigned)tiles, kSortThreads, 0, s>>>(kin, count, shift, hist, tiles);
        GW_LAUNCH_OK();
        GW_TRY(scan_exclusive_u32(hist, hist, 256 * tiles, nullptr, s));
        if (vals)
            radix_scatter_kernel<K, true><<<(unsigned)tiles, kSortThreads, 0, s>>>(
                kin, vin, kout, vout, count, shift, hist, tiles);
        else
            radix_scatter_kernel<K, false><<<(unsigned)tiles, kSortThreads, 0, s>>>(
                kin, nullptr, kout, nullptr, count, shift, hist, tiles);
        GW_LAUNCH_OK();
        K *t = kin; kin = kout; kout = t;
        uint32_t *u = vin; vin = vout; vout = u;
    }
    if (kin != keys) {
        GW_CUDA_OK(cudaMemcpyAsync(keys, kin, sizeof(K) * (size_t)count, cudaMemcpyDeviceToDevice, s));
        if (vals)
            GW_CUDA_OK(cudaMemcpyAsync(vals, vin, sizeof(uint32_t) * (size_t)count,
                                       cudaMemcpyDeviceToDevice, s));
    }
    return GW_OK;
}

int radix_sort_u64(uint64_t *keys, uint32_t *vals, int64_t count, int bits, cudaStream_t s)
{
    return radix_sort_impl<uint64_t>(keys, vals, count, bits, s);
}
int radix_sort_u32(uint32_t *keys, uint32_t *vals, int64_t count, int bits, cudaStream_t s)
{
    return radix_sort_impl<uint32_t>(keys, vals, count, bits, s);
}

}  // namespace gw

// ---- C ABI: housekeeping --------------------------------------------------------------------
extern "C" int gw_version(void) { return GW_VERSION; }
extern "C" long long gw_launch_count(void) { return gw::launches(); }
extern "C" const char *gw_last_error(void) { return gw::last_error(); }
extern "C" int gw_device_count(void)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        gw::set_error("cudaGetDeviceCount: %s", cudaGetErrorString(e));
        return e == cudaErrorNoDevice ? GW_ERR_NO_DEVICE : GW_ERR_CUDA;
    }
    return n;
}
