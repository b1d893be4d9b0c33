// dSCC (Spearman correlation of the i<j wish distances and reconstructed distances) at map sizes where neither the
// N x N distance matrix nor the P = N(N-1)/2 gathered vectors of the reference can exist.
//
// Reference: scipy.stats.spearmanr(dist_truth, dist_out) on the triu_indices gathers, every iteration in
// HiC_GAT_generalize_directly.py:210-242 and once per model in HiC-GNN_main.py:135-139 -- 16 B of int64 indices per pair,
// two P-vectors and an O(P log P) sort: 20 GB + minutes at 50k loci.
//
// Here: Spearman = Pearson of the average ranks.  Ranks are taken from two FINE HISTOGRAMS (B bins each) built in one
// streaming pass over the target rows (d is recomputed from the coordinates, never stored): the rank of a value is the
// mid-rank of its bin, i.e. the bin is treated as one tie group.  That is exact for the big tie group of a Hi-C wish matrix
// (every zero-contact pair has t == 1.0 and falls into one bin) and perturbs every other rank by at most half a bin population
// (relative 1/(2B) ~ 1e-7 at B = 2^22), far inside the 1e-3 the parity contract asks of the final dSCC.  A second streaming pass
// accumulates sum_k rank_t(k) * rank_d(k) in f64.  Both passes are row-block local: a sharded caller all-reduces the histograms
// and the cross sum.
#include "common.cuh"

namespace hicgat {
namespace {

constexpr int kRows = 32;   // rows per block tile
constexpr int kThreadsM = 256;

__device__ __forceinline__ int bin_of(float v, float scale, int nbins) {
    int b = (int)(v * scale);
    return b < 0 ? 0 : (b >= nbins ? nbins - 1 : b);
}

// grid = (column tiles of 256, row tiles of 32); upper triangle only (i < j)
template <bool CROSS>
__global__ void __launch_bounds__(kThreadsM) rank_pass_kernel(const float* __restrict__ coords, const float* __restrict__ target, int64_t pitch, int n, int r0, int r1,
                                                              float d_scale, float t_scale, int nbins, unsigned long long* __restrict__ hist_d,
                                                              unsigned long long* __restrict__ hist_t, const double* __restrict__ rank_d,
                                                              const double* __restrict__ rank_t, double* __restrict__ cross) {
    __shared__ float s_x[kRows][3];
    __shared__ double s_red[kThreadsM / 32];
    const int i0 = r0 + blockIdx.y * kRows;
    const int j = blockIdx.x * kThreadsM + threadIdx.x;
    if ((int)(blockIdx.x + 1) * kThreadsM - 1 <= i0) return;  // the whole tile lies on or below the diagonal
    if (threadIdx.x < kRows * 3) {
        const int r = min(i0 + threadIdx.x / 3, n - 1);
        s_x[threadIdx.x / 3][threadIdx.x % 3] = coords[(size_t)r * 3 + threadIdx.x % 3];
    }
    __syncthreads();
    double acc = 0.0;
    if (j < n) {
        const float xj = coords[(size_t)j * 3], yj = coords[(size_t)j * 3 + 1], zj = coords[(size_t)j * 3 + 2];
        const int rows = min(kRows, r1 - i0);
        for (int u = 0; u < rows; ++u) {
            const int i = i0 + u;
            if (i >= j) break;
            const float dx = xj - s_x[u][0], dy = yj - s_x[u][1], dz = zj - s_x[u][2];
            const float d = sqrtf(dx * dx + dy * dy + dz * dz);
            const float t = __ldg(target + (size_t)(i - r0) * pitch + j);
            const int bd = bin_of(d, d_scale, nbins), bt = bin_of(t, t_scale, nbins);
            if constexpr (CROSS) {
                acc += rank_d[bd] * rank_t[bt];
            } else {
                atomicAdd(hist_d + bd, 1ull);
                atomicAdd(hist_t + bt, 1ull);
            }
        }
    }
    if constexpr (CROSS) {
        acc = warp_sum(acc);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
        __syncthreads();
        if (threadIdx.x == 0) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kThreadsM / 32; ++w) s += s_red[w];
            if (s != 0.0) atomicAdd(cross, s);
        }
    }
}

// Implicit (sparse) target: every pair i<j of rows [r0,r1) goes into hist_d (no target stream).
__global__ void __launch_bounds__(kThreadsM) dist_hist_kernel(const float* __restrict__ coords, int n, int r0, int r1, float d_scale, int nbins,
                                                              unsigned long long* __restrict__ hist_d) {
    __shared__ float s_x[kRows][3];
    const int i0 = r0 + blockIdx.y * kRows;
    const int j = blockIdx.x * kThreadsM + threadIdx.x;
    if ((int)(blockIdx.x + 1) * kThreadsM - 1 <= i0) return;
    if (threadIdx.x < kRows * 3) {
        const int r = min(i0 + threadIdx.x / 3, n - 1);
        s_x[threadIdx.x / 3][threadIdx.x % 3] = coords[(size_t)r * 3 + threadIdx.x % 3];
    }
    __syncthreads();
    if (j >= n) return;
    const float xj = coords[(size_t)j * 3], yj = coords[(size_t)j * 3 + 1], zj = coords[(size_t)j * 3 + 2];
    const int rows = min(kRows, r1 - i0);
    for (int u = 0; u < rows; ++u) {
        if (i0 + u >= j) break;
        const float dx = xj - s_x[u][0], dy = yj - s_x[u][1], dz = zj - s_x[u][2];
        atomicAdd(hist_d + bin_of(sqrtf(dx * dx + dy * dy + dz * dz), d_scale, nbins), 1ull);
    }
}

// bins of the reconstructed distance of the stored pairs (i < j) of CSR rows [r0, r1): one thread per stored entry
__global__ void __launch_bounds__(256) edge_dist_bins_kernel(const float* __restrict__ coords, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                             int r0, int r1, float d_scale, int nbins, int32_t* __restrict__ bins) {
    const int i = r0 + blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= r1) return;
    const float xi = coords[(size_t)i * 3], yi = coords[(size_t)i * 3 + 1], zi = coords[(size_t)i * 3 + 2];
    for (int k = rowptr[i] + lane; k < rowptr[i + 1]; k += 32) {
        const int j = col[k];
        const float dx = coords[(size_t)j * 3] - xi, dy = coords[(size_t)j * 3 + 1] - yi, dz = coords[(size_t)j * 3 + 2] - zi;
        bins[k] = j > i ? bin_of(sqrtf(dx * dx + dy * dy + dz * dz), d_scale, nbins) : -1;
    }
}

}  // namespace
}  // namespace hicgat

using namespace hicgat;

static int rank_pass(bool cross, const float* coords, const float* target, int64_t pitch, int64_t n, int64_t r0, int64_t r1, float d_scale, float t_scale,
                     int nbins, unsigned long long* hist_d, unsigned long long* hist_t, const double* rank_d, const double* rank_t, double* cross_out,
                     cudaStream_t stream) {
    HICGAT_REQUIRE(n > 0 && n < (1ll << 30) && r0 >= 0 && r1 >= r0 && r1 <= n && nbins > 0, "hicgat_rank_*: bad n/r0/r1/nbins");
    HICGAT_REQUIRE(coords && (target || r0 == r1) && pitch >= n, "hicgat_rank_*: null pointer / bad pitch");
    if (r1 == r0) return HICGAT_OK;
    dim3 grid((unsigned)((n + kThreadsM - 1) / kThreadsM), (unsigned)((r1 - r0 + kRows - 1) / kRows));
    if (cross) rank_pass_kernel<true><<<grid, kThreadsM, 0, stream>>>(coords, target, pitch, (int)n, (int)r0, (int)r1, d_scale, t_scale, nbins, nullptr, nullptr, rank_d, rank_t, cross_out);
    else rank_pass_kernel<false><<<grid, kThreadsM, 0, stream>>>(coords, target, pitch, (int)n, (int)r0, (int)r1, d_scale, t_scale, nbins, hist_d, hist_t, nullptr, nullptr, nullptr);
    HICGAT_CHECK_LAUNCH("rank_pass_kernel");
    return HICGAT_OK;
}

extern "C" int hicgat_rank_histograms(const float* coords, const float* target, int64_t pitch, int64_t n, int64_t r0, int64_t r1, float d_scale,
                                      float t_scale, int nbins, uint64_t* hist_d, uint64_t* hist_t, hicgat_stream_t stream) {
    HICGAT_REQUIRE(hist_d && hist_t, "hicgat_rank_histograms: null histogram");
    return rank_pass(false, coords, target, pitch, n, r0, r1, d_scale, t_scale, nbins, reinterpret_cast<unsigned long long*>(hist_d),
                     reinterpret_cast<unsigned long long*>(hist_t), nullptr, nullptr, nullptr, static_cast<cudaStream_t>(stream));
}

extern "C" int hicgat_rank_cross_sum(const float* coords, const float* target, int64_t pitch, int64_t n, int64_t r0, int64_t r1, float d_scale,
                                     float t_scale, int nbins, const double* rank_d, const double* rank_t, double* cross, hicgat_stream_t stream) {
    HICGAT_REQUIRE(rank_d && rank_t && cross, "hicgat_rank_cross_sum: null pointer");
    return rank_pass(true, coords, target, pitch, n, r0, r1, d_scale, t_scale, nbins, nullptr, nullptr, rank_d, rank_t, cross, static_cast<cudaStream_t>(stream));
}

extern "C" int hicgat_dist_histogram(const float* coords, int64_t n, int64_t r0, int64_t r1, float d_scale, int nbins, uint64_t* hist_d,
                                     hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(coords && hist_d && n > 0 && n < (1ll << 30) && r0 >= 0 && r1 >= r0 && r1 <= n && nbins > 0, "hicgat_dist_histogram: bad arguments");
    if (r1 == r0) return HICGAT_OK;
    dim3 grid((unsigned)((n + kThreadsM - 1) / kThreadsM), (unsigned)((r1 - r0 + kRows - 1) / kRows));
    dist_hist_kernel<<<grid, kThreadsM, 0, stream>>>(coords, (int)n, (int)r0, (int)r1, d_scale, nbins, reinterpret_cast<unsigned long long*>(hist_d));
    HICGAT_CHECK_LAUNCH("dist_hist_kernel");
    return HICGAT_OK;
}

extern "C" int hicgat_edge_dist_bins(const float* coords, const int32_t* rowptr, const int32_t* col, int64_t n, int64_t r0, int64_t r1, float d_scale,
                                     int nbins, int32_t* bins, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(coords && rowptr && bins && n > 0 && r0 >= 0 && r1 >= r0 && r1 <= n && nbins > 0, "hicgat_edge_dist_bins: bad arguments");
    if (r1 == r0) return HICGAT_OK;
    edge_dist_bins_kernel<<<(unsigned)((r1 - r0 + 7) / 8), 256, 0, stream>>>(coords, rowptr, col, (int)r0, (int)r1, d_scale, nbins, bins);
    HICGAT_CHECK_LAUNCH("edge_dist_bins_kernel");
    return HICGAT_OK;
}
