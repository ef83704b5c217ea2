// gw_graph.cu -- device graph construction: configuration-model randomizer and the distinct-
// neighbour CSR builder (replaces nx.configuration_model + the networkx adjacency it yields,
// /root/reference/genewalk/null_distributions.py:23-29).
#include "gw_common.cuh"

namespace gw {

__global__ void widen_degrees_kernel(int32_t n, const int32_t *__restrict__ deg, int64_t *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = deg[i];
}

// one warp per node: stub k in [start[i], start[i]+deg[i]) is owned by node i and gets its shuffle key
__global__ void stub_keys_kernel(int32_t n, const int32_t *__restrict__ deg,
                                 const int64_t *__restrict__ start, uint32_t key0, uint32_t key1,
                                 uint64_t *keys, uint32_t *owner)
{
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp; i < n; i += nwarps) {
        const int64_t b = start[i], e = b + deg[i];
        for (int64_t k = b + lane; k < e; k += 32) {
            const u32x4 r = philox4x32_10((uint32_t)(uint64_t)k, (uint32_t)((uint64_t)k >> 32), 0u,
                                          kStreamConfig, key0, key1);
            keys[k] = ((uint64_t)r.x << 32) | r.y;
            owner[k] = (uint32_t)i;
        }
    }
}

__global__ void pair_halves_kernel(int64_t half, const uint32_t *__restrict__ shuffled,
                                   int32_t *edge_u, int32_t *edge_v)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < half) {
        edge_u[k] = (int32_t)shuffled[k];
        edge_v[k] = (int32_t)shuffled[k + half];
    }
}

__global__ void edge_keys_kernel(int64_t m, int64_t n, const int32_t *__restrict__ u,
                                 const int32_t *__restrict__ v, uint64_t *keys)
{
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k < m) {
        const uint64_t a = (uint64_t)u[k], b = (uint64_t)v[k];
        keys[2 * k] = a * (uint64_t)n + b;
        keys[2 * k + 1] = b * (uint64_t)n + a;
    }
}

__global__ void unique_flags_kernel(int64_t count, const uint64_t *__restrict__ keys, int64_t *flag)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) flag[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// unique sorted keys -> col_idx (compacted) and row_ptr (row boundaries, empty rows included)
__global__ void emit_csr_kernel(int64_t count, int64_t n, const uint64_t *__restrict__ keys,
                                const int64_t *__restrict__ pos, const int64_t *__restrict__ nnz,
                                int32_t *row_ptr, int32_t *col_idx)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const bool first = (i == 0) || keys[i] != keys[i - 1];
    if (!first) return;
    const int64_t row = (int64_t)(keys[i] / (uint64_t)n);
    const int64_t p = pos[i];
    col_idx[p] = (int32_t)(keys[i] - (uint64_t)row * (uint64_t)n);
    // rows (prev_row, row] start at p
    int64_t prev_row = -1;
    if (i > 0) prev_row = (int64_t)(keys[i - 1] / (uint64_t)n);
    for (int64_t r = prev_row + 1; r <= row; ++r) row_ptr[r] = (int32_t)p;
    if (p == *nnz - 1)  // the last unique entry also closes every remaining (empty) row
        for (int64_t r = row + 1; r <= n; ++r) row_ptr[r] = (int32_t)(*nnz);
}

}  // namespace gw

extern "C" int gw_config_model(int32_t n, const int32_t *degrees, int64_t S, uint64_t seed,
                               int32_t *edge_u, int32_t *edge_v, gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(n > 0 && degrees && edge_u && edge_v, "gw_config_model: null argument");
    GW_REQUIRE(S >= 0 && S < ((int64_t)1 << 32), "gw_config_model: stub count %lld out of range", (long long)S);
    GW_REQUIRE((S & 1) == 0, "Invalid degree sequence: sum of degrees must be even, not odd");
    if (S == 0) return GW_OK;
    Temp t_start, t_keys, t_owner;
    GW_CUDA_OK(t_start.alloc(sizeof(int64_t) * (size_t)n, s));
    GW_CUDA_OK(t_keys.alloc(sizeof(uint64_t) * (size_t)S, s));
    GW_CUDA_OK(t_owner.alloc(sizeof(uint32_t) * (size_t)S, s));
    const int blocks = ceil_div(n, 256);
    widen_degrees_kernel<<<blocks, 256, 0, s>>>(n, degrees, t_start.as<int64_t>());
    GW_LAUNCH_OK();
    GW_TRY(scan_exclusive_i64(t_start.as<int64_t>(), t_start.as<int64_t>(), n, nullptr, s));
    stub_keys_kernel<<<num_sms() * 8, 256, 0, s>>>(n, degrees, t_start.as<int64_t>(), (uint32_t)seed,
                                                  (uint32_t)(seed >> 32), t_keys.as<uint64_t>(),
                                                  t_owner.as<uint32_t>());
    GW_LAUNCH_OK();
    GW_TRY(radix_sort_u64(t_keys.as<uint64_t>(), t_owner.as<uint32_t>(), S, 64, s));
    const int64_t half = S / 2;
    pair_halves_kernel<<<(unsigned)ceil_div<int64_t>(half, 256), 256, 0, s>>>(half, t_owner.as<uint32_t>(),
                                                                            edge_u, edge_v);
    GW_LAUNCH_OK();
    return GW_OK;
}

extern "C" int gw_csr_from_edges(int32_t n, int64_t m, const int32_t *u, const int32_t *v,
                                 int32_t *row_ptr, int32_t *col_idx, int64_t *nnz_out_host,
                                 gw_stream_t stream)
{
    using namespace gw;
    cudaStream_t s = (cudaStream_t)stream;
    GW_REQUIRE(n > 0 && row_ptr && col_idx && nnz_out_host, "gw_csr_from_edges: null argument");
    GW_REQUIRE(m >= 0 && 2 * m < ((int64_t)1 << 31), "gw_csr_from_edges: m=%lld out of range", (long long)m);
    if (m == 0) {
        GW_CUDA_OK(cudaMemsetAsync(row_ptr, 0, sizeof(int32_t) * ((size_t)n + 1), s));
        GW_CUDA_OK(cudaStreamSynchronize(s));
        *nnz_out_host = 0;
        return GW_OK;
    }
    GW_REQUIRE(u && v, "gw_csr_from_edges: null edge arrays");
    const int64_t count = 2 * m;
    Temp t_keys, t_flag, t_nnz;
    GW_CUDA_OK(t_keys.alloc(sizeof(uint64_t) * (size_t)count, s));
    GW_CUDA_OK(t_flag.alloc(sizeof(int64_t) * (size_t)count, s));
    GW_CUDA_OK(t_nnz.alloc(sizeof(int64_t), s));
    edge_keys_kernel<<<(unsigned)ceil_div<int64_t>(m, 256), 256, 0, s>>>(m, n, u, v, t_keys.as<uint64_t>());
    GW_LAUNCH_OK();
    int bits = 1;
    while (bits < 64 && ((uint64_t)1 << bits) < (uint64_t)n * (uint64_t)n) ++bits;
    GW_TRY(radix_sort_u64(t_keys.as<uint64_t>(), nullptr, count, bits, s));
    const unsigned blocks = (unsigned)ceil_div<int64_t>(count, 256);
    unique_flags_kernel<<<blocks, 256, 0, s>>>(count, t_keys.as<uint64_t>(), t_flag.as<int64_t>());
    GW_LAUNCH_OK();
    GW_TRY(scan_exclusive_i64(t_flag.as<int64_t>(), t_flag.as<int64_t>(), count, t_nnz.as<int64_t>(), s));
    emit_csr_kernel<<<blocks, 256, 0, s>>>(count, n, t_keys.as<uint64_t>(), t_flag.as<int64_t>(),
                                          t_nnz.as<int64_t>(), row_ptr, col_idx);
    GW_LAUNCH_OK();
    GW_CUDA_OK(cudaMemcpyAsync(nnz_out_host, t_nnz.ptr, sizeof(int64_t), cudaMemcpyDeviceToHost, s));
    GW_CUDA_OK(cudaStreamSynchronize(s));
    return GW_OK;
}
