// gw_stats.cu -- gene-GO similarity statistics (replaces the numeric core of
// /root/reference/genewalk/perform_statistics.py:93-124,247-283, null_distributions.py:33-48 and
// the pooled sort at cli.py:238).
#include <math.h>

#include "gw_common.cuh"

namespace gw {

// ---- cosine similarities ----------------------------------------------------------------------
// KeyedVectors.similarity: dot(unitvec(a), unitvec(b)); unitvec scales by float32(1/|a|).
__global__ void cosine_pairs_kernel(int32_t d, const float *__restrict__ vec, int64_t m,
                                    const int32_t *__restrict__ a_idx,
                                    const int32_t *__restrict__ b_idx, float *sims)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const float *a = vec + (int64_t)a_idx[k] * d;
    const float *b = vec + (int64_t)b_idx[k] * d;
    float na = 0.0f, nb = 0.0f;
    for (int i = 0; i < d; ++i) {
        na = __fadd_rn(na, __fmul_rn(a[i], a[i]));
        nb = __fadd_rn(nb, __fmul_rn(b[i], b[i]));
    }
    na = __fsqrt_rn(na);
    nb = __fsqrt_rn(nb);
    const float ia = na > 0.0f ? (float)(1.0 / (double)na) : 1.0f;
    const float ib = nb > 0.0f ? (float)(1.0 / (double)nb) : 1.0f;
    float dot = 0.0f;
    for (int i = 0; i < d; ++i) dot = __fadd_rn(dot, __fmul_rn(__fmul_rn(a[i], ia), __fmul_rn(b[i], ib)));
    sims[k] = dot;
}

__global__ void row_norms_kernel(int32_t n, int32_t d, const float *__restrict__ vec, float *norms)
{
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n) return;
    const float *a = vec + (int64_t)v * d;
    float s = 0.0f;
    for (int i = 0; i < d; ++i) s = __fadd_rn(s, __fmul_rn(a[i], a[i]));
    norms[v] = __fsqrt_rn(s);
}

__global__ void null_flags_kernel(int32_t n, const int32_t *__restrict__ row_ptr,
                                  const int32_t *__restrict__ col_idx, int32_t *row_of_entry,
                                  int64_t *flag)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t v = warp; v < n; v += nwarps) {
        const int32_t b = row_ptr[v], e = row_ptr[v + 1];
        for (int32_t k = b + lane; k < e; k += 32) {
            row_of_entry[k] = (int32_t)v;
            flag[k] = col_idx[k] != (int32_t)v ? 1 : 0;
        }
    }
}

// KeyedVectors.distances: 1 - dot(V[others], v)/(|v| * |V[others]|); the caller takes 1 - that.
__global__ void null_sims_kernel(int64_t A, int32_t d, const float *__restrict__ vec,
                                 const float *__restrict__ norms,
                                 const int32_t *__restrict__ row_of_entry,
                                 const int32_t *__restrict__ col_idx,
                                 const int64_t *__restrict__ pos, float *sims)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= A) return;
    const int32_t u = row_of_entry[k], v = col_idx[k];
    if (u == v) return;
    const float *a = vec + (int64_t)u * d;
    const float *b = vec + (int64_t)v * d;
    float dot = 0.0f;
    for (int i = 0; i < d; ++i) dot = __fadd_rn(dot, __fmul_rn(b[i], a[i]));
    const float sim = __fdiv_rn(dot, __fmul_rn(norms[u], norms[v]));
    const float dist = __fsub_rn(1.0f, sim);
    sims[pos[k]] = __fsub_rn(1.0f, dist);
}

// ---- float sort ---------------------------------------------------------------------------------
__global__ void f32_to_sortable_kernel(int64_t n, uint32_t *x)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t b = x[i];
        x[i] = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    }
}
__global__ void sortable_to_f32_kernel(int64_t n, uint32_t *x)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        const uint32_t b = x[i];
        x[i] = (b & 0x80000000u) ? (b & 0x7fffffffu) : ~b;
    }
}

// ---- empirical p-values ---------------------------------------------------------------------------
__global__ void psim_kernel(const float *__restrict__ null_sorted, int64_t n_null,
                            const float *__restrict__ sims, int64_t m, int64_t *ranks, double *pvals)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    const float x = sims[k];
    int64_t lo = 0, hi = n_null;  // np.searchsorted(side='left'): first index with null[i] >= x
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(null_sorted + mid) < x) lo = mid + 1; else hi = mid;
    }
    if (ranks) ranks[k] = lo;
    double p = 1.0 - (double)lo / (double)n_null;
    if (p < 1e-16) p = 1e-16;
    pvals[k] = p;
}

// ---- Benjamini-Hochberg -----------------------------------------------------------------------------
__global__ void bh_keys_kernel(int64_t m, const double *__restrict__ p, uint64_t *keys, uint32_t *idx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m) {
        keys[i] = (uint64_t)__double_as_longlong(p[i]);  // p >= 0: bit pattern is order-preserving
        idx[i] = (uint32_t)i;
    }
}

__global__ void bh_segment_keys_kernel(int64_t m, const uint32_t *__restrict__ idx,
                                       const int64_t *__restrict__ seg_ptr, int64_t nseg,
                                       uint32_t *seg_keys)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int64_t i = idx[j];
    int64_t lo = 0, hi = nseg;  // last s with seg_ptr[s] <= i
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (seg_ptr[mid] <= i) lo = mid; else hi = mid;
    }
    seg_keys[j] = (uint32_t)lo;
}

// sorted position j (segment s, ascending p inside the segment) -> p_(i) / (i / n_s)
__global__ void bh_raw_kernel(int64_t m, const double *__restrict__ p,
                              const uint32_t *__restrict__ idx, const uint32_t *__restrict__ seg_keys,
                              const int64_t *__restrict__ seg_ptr, SegVal *raw)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const int64_t s = seg_keys[j];
    const int64_t b = seg_ptr[s], ns = seg_ptr[s + 1] - b;
    const double ecdf = (double)(j - b + 1) / (double)ns;
    raw[j].v = p[idx[j]] / ecdf;
    raw[j].seg = s;
}

__global__ void bh_store_kernel(int64_t m, const SegVal *__restrict__ scanned,
                                const uint32_t *__restrict__ idx, double *q)
{
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= m) return;
    const double v = scanned[j].v;
    q[idx[j]] = v > 1.0 ? 1.0 : v;
}

// ---- log_stats ---------------------------------------------------------------------------------------
__global__ void log_stats_kernel(const double *__restrict__ vals, int64_t rows, int32_t nreps, double *out)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const double eps = 1e-16;
    const double *v = vals + r * nreps;
    double mean = 0.0;
    for (int i = 0; i < nreps; ++i) mean += log(v[i] + eps);
    mean /= (double)nreps;
    const double g_mean = exp(mean) - eps;
    double g_std = eps;
    if (nreps > 1) {
        double ss = 0.0;
        for (int i = 0; i < nreps; ++i) {
            const double dlt = log(v[i] + eps) - mean;
            ss += dlt * dlt;
        }
        g_std = exp(sqrt(ss / (double)(nreps - 1)));
    }
    const double e = 1.96 / sqrt((double)nreps);
    out[3 * r] = g_mean;
    out[3 * r + 1] = g_mean * pow(g_std, -e);
    out[3 * r + 2] = g_mean * pow(g_std, e);
}

// mean / SEM of the per-replicate similarities (perform_statistics.py:180-184): float32 arithmetic
// like np.mean / np.std on a float32 array, SEM = std(ddof=0) / sqrt(nreps) in float64
__global__ void mean_sem_kernel(const float *__restrict__ sims, int64_t rows, int32_t nreps,
                                float *mean, double *sem)
{
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const float *v = sims + r * nreps;
    float sum = 0.0f;
    for (int i = 0; i < nreps; ++i) sum = __fadd_rn(sum, v[i]);
    const float mu = __fdiv_rn(sum, (float)nreps);
    float ss = 0.0f;
    for (int i = 0; i < nreps; ++i) {
        const float dlt = __fsub_rn(v[i], mu);
        ss = __fadd_rn(ss, __fmul_rn(dlt, dlt));
    }
    const float sd = __fsqrt_rn(__fdiv_rn(ss, (float)nreps));
    mean[r] = mu;
    sem[r] = (double)sd / sqrt((double)nreps);
}

}  // namespace gw

extern "C" int gw_mean_sem_f32(const float *sims, int64_t rows, int32_t nreps, float *mean,
                               double *sem, gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(rows >= 0 && nreps >= 1, "gw_mean_sem_f32: bad sizes");
    if (rows == 0) return GW_OK;
    GW_REQUIRE(sims && mean && sem, "gw_mean_sem_f32: null argument");
    mean_sem_kernel<<<(unsigned)ceil_div<int64_t>(rows, 128), 128, 0, s>>>(sims, rows, nreps, mean, sem);
    GW_LAUNCH_OK();
    return GW_OK;
}

extern "C" int gw_cosine_pairs(int32_t d, const float *vectors, int64_t m, const int32_t *a_idx,
                               const int32_t *b_idx, float *sims, gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(d > 0 && m >= 0, "gw_cosine_pairs: bad sizes");
    if (m == 0) return GW_OK;
    GW_REQUIRE(vectors && a_idx && b_idx && sims, "gw_cosine_pairs: null argument");
    cosine_pairs_kernel<<<(unsigned)ceil_div<int64_t>(m, 256), 256, 0, s>>>(d, vectors, m, a_idx, b_idx, sims);
    GW_LAUNCH_OK();
    return GW_OK;
}

extern "C" int gw_null_similarities(int32_t n, const int32_t *row_ptr, const int32_t *col_idx,
                                    int32_t d, const float *vectors, float *sims_out,
                                    int64_t *count_out_host, gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(n > 0 && d > 0 && row_ptr && col_idx && vectors && sims_out && count_out_host,
               "gw_null_similarities: null argument");
    int32_t A32 = 0;
    GW_CUDA_OK(cudaMemcpyAsync(&A32, row_ptr + n, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    GW_CUDA_OK(cudaStreamSynchronize(s));
    const int64_t A = A32;
    if (A == 0) {
        *count_out_host = 0;
        return GW_OK;
    }
    Temp t_rows, t_flag, t_norm, t_cnt;
    GW_CUDA_OK(t_rows.alloc(sizeof(int32_t) * (size_t)A, s));
    GW_CUDA_OK(t_flag.alloc(sizeof(int64_t) * (size_t)A, s));
    GW_CUDA_OK(t_norm.alloc(sizeof(float) * (size_t)n, s));
    GW_CUDA_OK(t_cnt.alloc(sizeof(int64_t), s));
    null_flags_kernel<<<num_sms() * 8, 256, 0, s>>>(n, row_ptr, col_idx, t_rows.as<int32_t>(), t_flag.as<int64_t>());
    GW_LAUNCH_OK();
    GW_TRY(scan_exclusive_i64(t_flag.as<int64_t>(), t_flag.as<int64_t>(), A, t_cnt.as<int64_t>(), s));
    row_norms_kernel<<<ceil_div(n, 256), 256, 0, s>>>(n, d, vectors, t_norm.as<float>());
    GW_LAUNCH_OK();
    null_sims_kernel<<<(unsigned)ceil_div<int64_t>(A, 256), 256, 0, s>>>(
        A, d, vectors, t_norm.as<float>(), t_rows.as<int32_t>(), col_idx, t_flag.as<int64_t>(), sims_out);
    GW_LAUNCH_OK();
    GW_CUDA_OK(cudaMemcpyAsync(count_out_host, t_cnt.ptr, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    GW_CUDA_OK(cudaStreamSynchronize(s));
    return GW_OK;
}

extern "C" int gw_sort_f32(float *keys, int64_t count, gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(count >= 0 && (keys || count == 0), "gw_sort_f32: bad arguments");
    if (count <= 1) return GW_OK;
    const unsigned blocks = (unsigned)ceil_div<int64_t>(count, 256);
    uint32_t *x = reinterpret_cast<uint32_t *>(keys);
    f32_to_sortable_kernel<<<blocks, 256, 0, s>>>(count, x);
    GW_LAUNCH_OK();
    GW_TRY(radix_sort_u32(x, nullptr, count, 32, s));
    sortable_to_f32_kernel<<<blocks, 256, 0, s>>>(count, x);
    GW_LAUNCH_OK();
    return GW_OK;
}

extern "C" int gw_psim(const float *null_sorted, int64_t n_null, const float *sims, int64_t m,
                       int64_t *ranks, double *pvals, gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(m >= 0 && n_null > 0 && null_sorted, "gw_psim: empty null distribution");
    if (m == 0) return GW_OK;
    GW_REQUIRE(sims && pvals, "gw_psim: null argument");
    psim_kernel<<<(unsigned)ceil_div<int64_t>(m, 256), 256, 0, s>>>(null_sorted, n_null, sims, m, ranks, pvals);
    GW_LAUNCH_OK();
    return GW_OK;
}

extern "C" int gw_bh_segmented(const double *pvals, int64_t m, const int64_t *seg_ptr, int64_t nseg,
                               double *qvals, gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(m >= 0 && nseg >= 0, "gw_bh_segmented: bad sizes");
    if (m == 0) return GW_OK;
    GW_REQUIRE(pvals && seg_ptr && qvals && nseg >= 1, "gw_bh_segmented: null argument");
    GW_REQUIRE(m < ((int64_t)1 << 32) && nseg < ((int64_t)1 << 32), "gw_bh_segmented: too many rows");
    Temp t_keys, t_idx, t_seg, t_raw;
    GW_CUDA_OK(t_keys.alloc(sizeof(uint64_t) * (size_t)m, s));
    GW_CUDA_OK(t_idx.alloc(sizeof(uint32_t) * (size_t)m, s));
    GW_CUDA_OK(t_seg.alloc(sizeof(uint32_t) * (size_t)m, s));
    GW_CUDA_OK(t_raw.alloc(sizeof(SegVal) * (size_t)m, s));
    const unsigned blocks = (unsigned)ceil_div<int64_t>(m, 256);
    bh_keys_kernel<<<blocks, 256, 0, s>>>(m, pvals, t_keys.as<uint64_t>(), t_idx.as<uint32_t>());
    GW_LAUNCH_OK();
    GW_TRY(radix_sort_u64(t_keys.as<uint64_t>(), t_idx.as<uint32_t>(), m, 64, s));
    bh_segment_keys_kernel<<<blocks, 256, 0, s>>>(m, t_idx.as<uint32_t>(), seg_ptr, nseg, t_seg.as<uint32_t>());
    GW_LAUNCH_OK();
    int bits = 1;
    while (bits < 32 && ((int64_t)1 << bits) < nseg) ++bits;
    if (nseg > 1) GW_TRY(radix_sort_u32(t_seg.as<uint32_t>(), t_idx.as<uint32_t>(), m, bits, s));
    bh_raw_kernel<<<blocks, 256, 0, s>>>(m, pvals, t_idx.as<uint32_t>(), t_seg.as<uint32_t>(), seg_ptr,
                                         t_raw.as<SegVal>());
    GW_LAUNCH_OK();
    GW_TRY(scan_segmented_suffix_min(t_raw.as<SegVal>(), t_raw.as<SegVal>(), m, s));
    bh_store_kernel<<<blocks, 256, 0, s>>>(m, t_raw.as<SegVal>(), t_idx.as<uint32_t>(), qvals);
    GW_LAUNCH_OK();
    return GW_OK;
}

extern "C" int gw_log_stats(const double *vals, int64_t rows, int32_t nreps, double *out,
                            gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(rows >= 0 && nreps >= 1, "gw_log_stats: bad sizes");
    if (rows == 0) return GW_OK;
    GW_REQUIRE(vals && out, "gw_log_stats: null argument");
    log_stats_kernel<<<(unsigned)ceil_div<int64_t>(rows, 128), 128, 0, s>>>(vals, rows, nreps, out);
    GW_LAUNCH_OK();
    return GW_OK;
}
