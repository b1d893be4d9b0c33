// One-off tensor builders that feed the training step:
//   * wish-distance matrix     -- utils.cont2dist (utils.py:75-80) + the per-iteration
//                                 truth.float() cast (HiC-GNN_main.py:127)
//   * symmetric CSR graph      -- graph half of utils.load_input (utils.py:33-71)
//   * SAGEConv edge weights    -- layers.py:41-54 (adjust_weights)
// All HBM-bound streaming / scan work; integer results are bit-exact against the oracle.
#include <math_constants.h>

#include "common.cuh"

namespace hicgat {
namespace {

// (1/a)^factor with ATen's pow special cases (exponent 1 -> copy, 0.5 -> sqrt, 2 -> square).
// Factors 1 and 2 are bit-exact against torch's CPU result; 0.5 is the correctly rounded IEEE
// sqrt (torch's AVX-512 CPU sqrt is Sleef u05 and can be 1-2 ulp off it); generic factors use
// CUDA's f64 pow (<= 2 ulp).
__device__ __forceinline__ double inv_pow(double a, double factor, int fkind) {
    const double r = 1.0 / a;
    switch (fkind) {
        case 1: return r;
        case 2: return sqrt(r);
        case 3: return r * r;
        default: return pow(r, factor);
    }
}
int factor_kind(double f) { return f == 1.0 ? 1 : f == 0.5 ? 2 : f == 2.0 ? 3 : 0; }

// Both passes are pure streams with a long f64 division chain per element: what limits them is the
// number of bytes a thread keeps in flight, so each thread owns TWO adjacent columns (one 16-byte load)
// and walks TWO rows per iteration (4 independent divisions, 32 B outstanding per thread, 64 KB per SM).
__device__ __forceinline__ void max_update(double& m, double a, double factor, int fkind, bool skip) {
    const double v = inv_pow(a, factor, fkind);
    // torch.max(nan_to_num(dist, posinf=0)): NaN -> 0, +inf -> 0 (utils.py:78)
    if (!skip && v > m && v < CUDART_INF) m = v;
}

__global__ void __launch_bounds__(256) cont2dist_max_kernel(const double* __restrict__ adj, int64_t ld, int n, int r0, int r1,
                                                            double factor, int fkind, unsigned long long* max_bits) {
    double m = 0.0;
    const bool vec = ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(adj) & 15) == 0);
    for (int i = r0 + 2 * blockIdx.x; i < r1; i += 2 * gridDim.x) {
        const double* row0 = adj + (size_t)(i - r0) * ld;
        const bool two = i + 1 < r1;
        const double* row1 = two ? row0 + ld : row0;
        if (vec) {
            for (int j = 2 * threadIdx.x; j < n; j += 2 * blockDim.x) {
                double2 a0, a1;
                if (j + 1 < n) {
                    a0 = __ldg(reinterpret_cast<const double2*>(row0 + j));
                    a1 = __ldg(reinterpret_cast<const double2*>(row1 + j));
                } else {
                    a0 = make_double2(row0[j], 0.0);
                    a1 = make_double2(row1[j], 0.0);
                }
                // No contact (+0.0: ~99 % of a 50k-locus map) gives (1/0)^f = +inf, which the max ignores.  When ALL elements this
                // warp holds are +0.0 -- everything away from the diagonal band -- the f64 division chains are skipped as a whole
                // (a warp-uniform branch: the four divisions of the generic path keep their instruction-level parallelism).
                const bool zero4 = (__double_as_longlong(a0.x) | __double_as_longlong(a0.y) | __double_as_longlong(a1.x) | __double_as_longlong(a1.y)) == 0ll;
                if (factor > 0.0 && __all_sync(__activemask(), zero4)) continue;
                max_update(m, a0.x, factor, fkind, j == i);
                max_update(m, a1.x, factor, fkind, !two || j == i + 1);
                if (j + 1 < n) {
                    max_update(m, a0.y, factor, fkind, j + 1 == i);
                    max_update(m, a1.y, factor, fkind, !two || j + 1 == i + 1);
                }
            }
        } else {
            for (int j = threadIdx.x; j < n; j += blockDim.x) {
                max_update(m, row0[j], factor, fkind, j == i);
                max_update(m, row1[j], factor, fkind, !two || j == i + 1);
            }
        }
    }
    __shared__ double s[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) m = fmax(m, s[w]);
        // non-negative doubles order like their bit patterns
        atomicMax(max_bits, (unsigned long long)__double_as_longlong(m));
    }
}

__device__ __forceinline__ double wish_value(double a, bool diag, double factor, int fkind, double mx) {
    double v = diag ? 0.0 : inv_pow(a, factor, fkind);   // utils.py:76-77
    if (v != v) v = 0.0;                                  // nan_to_num: NaN -> 0
    else if (v == CUDART_INF) v = mx;                     // posinf -> max (utils.py:79)
    else if (v == -CUDART_INF) v = -1.7976931348623157e308;
    return v / mx;                                        // utils.py:80
}

// thread = two adjacent columns, rows strided by gridDim.y, two rows per iteration
__global__ void __launch_bounds__(256) cont2dist_apply_kernel(const double* __restrict__ adj, int64_t ld, int n, int r0, int r1,
                                                              double factor, int fkind, const double* __restrict__ max_in,
                                                              double* __restrict__ o64, int64_t ld64, float* __restrict__ o32, int64_t p32) {
    const double mx = *max_in;
    const bool fast_ok = factor > 0.0 && mx > 0.0 && mx < CUDART_INF;
    const int j = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const bool vin = ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(adj) & 15) == 0);
    const bool v64 = o64 && ((ld64 & 1) == 0) && ((reinterpret_cast<uintptr_t>(o64) & 15) == 0);
    const bool v32 = o32 && ((p32 & 1) == 0) && ((reinterpret_cast<uintptr_t>(o32) & 7) == 0);
    constexpr int kR = 4;  // rows per iteration: all loads are issued before the first division (64 B in flight per thread)
    for (int i = r0 + kR * blockIdx.y; i < r1; i += kR * gridDim.y) {
        double2 av[kR];
#pragma unroll
        for (int u = 0; u < kR; ++u) {
            const int r = i + u;
            av[u] = make_double2(0.0, 0.0);
            if (r < r1 && j + 1 < n) {
                const double* row = adj + (size_t)(r - r0) * ld;
                av[u] = vin ? __ldg(reinterpret_cast<const double2*>(row + j)) : make_double2(row[j], row[j + 1]);
            }
        }
#pragma unroll
        for (int u = 0; u < kR; ++u) {
            const int r = i + u;
            if (r >= r1) break;
            const double* row = adj + (size_t)(r - r0) * ld;
            if (j + 1 < n) {
                const double2 a = av[u];
                // +0.0 off the diagonal: (1/0)^f = +inf -> max -> max / max = exactly 1.0 (finite non-zero max).  Warp-uniform
                // shortcut for warps that hold nothing else: no f64 divisions, the kernel runs at HBM speed away from the band.
                const bool zero2 = (__double_as_longlong(a.x) | __double_as_longlong(a.y)) == 0ll && j != r && j + 1 != r;
                double w0, w1;
                if (fast_ok && __all_sync(__activemask(), zero2)) {
                    w0 = w1 = 1.0;
                } else {
                    w0 = wish_value(a.x, j == r, factor, fkind, mx);
                    w1 = wish_value(a.y, j + 1 == r, factor, fkind, mx);
                }
                if (o64) {
                    double* d = o64 + (size_t)(r - r0) * ld64 + j;
                    if (v64) *reinterpret_cast<double2*>(d) = make_double2(w0, w1);
                    else { d[0] = w0; d[1] = w1; }
                }
                if (o32) {
                    float* d = o32 + (size_t)(r - r0) * p32 + j;
                    if (v32) *reinterpret_cast<float2*>(d) = make_float2((float)w0, (float)w1);
                    else { d[0] = (float)w0; d[1] = (float)w1; }
                }
            } else {
                if (j < n) {
                    const double w0 = wish_value(row[j], j == r, factor, fkind, mx);
                    if (o64) o64[(size_t)(r - r0) * ld64 + j] = w0;
                    if (o32) o32[(size_t)(r - r0) * p32 + j] = (float)w0;
                }
                // keep the 16-byte pitch padding defined
                if (o32) {
                    for (int c = max(j, n); c < min((int64_t)j + 2, p32); ++c) o32[(size_t)(r - r0) * p32 + c] = 0.f;
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------- CSR build
// Edge rule of utils.load_input (networkx edge walk + SparseTensor.to_symmetric), closed form:
// {i,j}, i != j, exists iff A[i,j] != 0 or A[j,i] != 0; weight = A[max,min] if non-zero else
// A[min,max].  One warp per row; the column A[:,i] is read strided (adjacent rows share sectors).
__device__ __forceinline__ bool edge_at(const double* __restrict__ adj, int64_t ld, int i, int j, int self_loops, float& w) {
    if (j == i) {
        w = 1.f;
        return self_loops != 0;
    }
    const double a_ij = adj[(size_t)i * ld + j], a_ji = adj[(size_t)j * ld + i];
    const double lower = i > j ? a_ij : a_ji, upper = i > j ? a_ji : a_ij;
    w = (float)(lower != 0.0 ? lower : upper);
    return (a_ij != 0.0) || (a_ji != 0.0);
}

__global__ void __launch_bounds__(256) csr_count_kernel(const double* __restrict__ adj, int64_t ld, int n, int self_loops, int64_t* __restrict__ rowcount) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    int cnt = 0;
    for (int base = 0; base < n; base += 32) {
        const int j = base + lane;
        float w;
        const bool e = j < n && edge_at(adj, ld, i, j, self_loops, w);
        cnt += __popc(__ballot_sync(0xffffffffu, e));
    }
    if (lane == 0) rowcount[i] = cnt;
}

__global__ void __launch_bounds__(256) csr_fill_kernel(const double* __restrict__ adj, int64_t ld, int n, int self_loops,
                                                       const int64_t* __restrict__ rowptr, int64_t* __restrict__ col, float* __restrict__ val) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    int64_t pos = rowptr[i];
    for (int base = 0; base < n; base += 32) {
        const int j = base + lane;
        float w = 0.f;
        const bool e = j < n && edge_at(adj, ld, i, j, self_loops, w);
        const unsigned m = __ballot_sync(0xffffffffu, e);
        if (e) {
            const int64_t p = pos + __popc(m & ((1u << lane) - 1u));
            col[p] = j;
            val[p] = w;
        }
        pos += __popc(m);
    }
}

// exclusive scan of int64 counts, single CTA (n is at most a few 1e5; one-off)
__global__ void __launch_bounds__(1024) scan_i64_kernel(const int64_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
    __shared__ int64_t s_warp[32];
    __shared__ int64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int64_t base = 0; base < n; base += 1024) {
        const int64_t idx = base + threadIdx.x;
        const int64_t v = idx < n ? in[idx] : 0;
        int64_t x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            int64_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int64_t carry = s_carry;
        const int64_t incl = x + (warp > 0 ? s_warp[warp - 1] : 0) + carry;
        if (idx < n) out[idx] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) out[n] = s_carry;
}

__global__ void pack_i32_kernel(const int64_t* __restrict__ rowptr, const int64_t* __restrict__ col, int64_t n, int64_t nnz,
                                int32_t* __restrict__ rowptr32, int32_t* __restrict__ col32) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += stride) col32[k] = (int32_t)col[k];
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k <= n; k += stride) rowptr32[k] = (int32_t)rowptr[k];
}

// SAGEConv.adjust_weights (layers.py:41-54): s = adj_t.sum(dim=0) accumulated in storage
// order (f32); norm = val / s[row].  The pattern and the values are symmetric, so the column
// sum of j visits exactly the values of row j in the same (ascending) order: a sequential
// per-row sum is bit-identical to torch-sparse's scatter-add.
__global__ void sage_norm_kernel(const int32_t* __restrict__ rowptr, const float* __restrict__ val, int n,
                                 float* __restrict__ colsum, float* __restrict__ norm_val) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = rowptr[i], e = rowptr[i + 1];
    float s = 0.f;
    for (int k = b; k < e; ++k) s += val[k];
    colsum[i] = s;
    const float inv = 1.0f / s;  // torch.divide(ones, sum_vec)
    for (int k = b; k < e; ++k) norm_val[k] = inv * val[k];
}

// perm[k] = index of entry (j, i) for entry k = (i, j): binary search in row j (columns ascending)
__global__ void transpose_perm_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n, int32_t* __restrict__ perm) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    for (int k = rowptr[i] + lane; k < rowptr[i + 1]; k += 32) {
        const int j = col[k];
        int lo = rowptr[j], hi = rowptr[j + 1] - 1, found = -1;
        while (lo <= hi) {
            const int mid = (lo + hi) >> 1;
            const int c = col[mid];
            if (c == i) { found = mid; break; }
            if (c < i) lo = mid + 1; else hi = mid - 1;
        }
        perm[k] = found;
    }
}


// torch-sparse set_diag for a pattern WITHOUT diagonal entries (load_input removes them,
// utils.py:33,59-63): row i gains (i,i) at its sorted position.  One warp per row.
__global__ void __launch_bounds__(256) add_self_loops_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, int n,
                                                             int32_t* __restrict__ out_rowptr, int32_t* __restrict__ out_col) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i > n) return;
    if (i == n) {
        if (lane == 0) out_rowptr[n] = rowptr[n] + n;
        return;
    }
    const int rs = rowptr[i], re = rowptr[i + 1];
    if (lane == 0) out_rowptr[i] = rs + i;
    int below = 0;  // entries with column < i
    for (int k = rs + lane; k < re; k += 32) {
        const int c = col[k];
        const int lt = c < i;
        out_col[k + i + (lt ? 0 : 1)] = c;
        below += lt;
    }
    below = __reduce_add_sync(0xffffffffu, below);
    if (lane == 0) out_col[rs + i + below] = i;
}

}  // namespace
}  // namespace hicgat

using namespace hicgat;

extern "C" size_t hicgat_cont2dist_workspace_bytes(int64_t, int64_t, int64_t) { return 256; }

extern "C" int hicgat_cont2dist_max_f64(const double* adj, int64_t ld, int64_t n, int64_t r0, int64_t r1, double factor,
                                        double* max_out, void* workspace, size_t workspace_bytes, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    (void)workspace; (void)workspace_bytes;
    HICGAT_REQUIRE(adj && max_out, "hicgat_cont2dist_max_f64: null pointer");
    HICGAT_REQUIRE(n > 0 && n < (1ll << 30) && r0 >= 0 && r1 >= r0 && r1 <= n && ld >= n, "hicgat_cont2dist_max_f64: bad shape");
    HICGAT_CUDA(cudaMemsetAsync(max_out, 0, sizeof(double), stream));
    if (r1 == r0) return HICGAT_OK;
    const int64_t row_pairs = (r1 - r0 + 1) / 2;
    const int grid = (int)(row_pairs < 148 * 8 ? row_pairs : 148 * 8);
    cont2dist_max_kernel<<<grid, 256, 0, stream>>>(adj, ld, (int)n, (int)r0, (int)r1, factor, factor_kind(factor),
                                                   reinterpret_cast<unsigned long long*>(max_out));
    HICGAT_CHECK_LAUNCH("cont2dist_max_kernel");
    return HICGAT_OK;
}

extern "C" int hicgat_cont2dist_apply_f64(const double* adj, int64_t ld, int64_t n, int64_t r0, int64_t r1, double factor,
                                          const double* max_in, double* out_f64, int64_t ld_f64, float* out_f32,
                                          int64_t pitch_f32, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(adj && max_in && (out_f64 || out_f32), "hicgat_cont2dist_apply_f64: null pointer");
    HICGAT_REQUIRE(n > 0 && n < (1ll << 30) && r0 >= 0 && r1 >= r0 && r1 <= n && ld >= n, "hicgat_cont2dist_apply_f64: bad shape");
    HICGAT_REQUIRE(!out_f64 || ld_f64 >= n, "hicgat_cont2dist_apply_f64: ld_f64 < n");
    HICGAT_REQUIRE(!out_f32 || pitch_f32 >= n, "hicgat_cont2dist_apply_f64: pitch_f32 < n");
    if (r1 == r0) return HICGAT_OK;
    const int64_t width = out_f32 ? pitch_f32 : n;
    // ~16 CTAs per SM in total; a thread owns two columns and walks row pairs with a stride of gridDim.y
    // (a CTA per (column block, row) would be 6.4 M one-element CTAs at 50k loci)
    const int64_t xblocks = (width + 511) / 512;
    const int64_t row_pairs = (r1 - r0 + 3) / 4;  // row groups of 4
    int64_t yblocks = (148 * 16 + xblocks - 1) / xblocks;
    if (yblocks > row_pairs) yblocks = row_pairs;
    if (yblocks < 1) yblocks = 1;
    dim3 grid((unsigned)xblocks, (unsigned)yblocks);
    cont2dist_apply_kernel<<<grid, 256, 0, stream>>>(adj, ld, (int)n, (int)r0, (int)r1, factor, factor_kind(factor), max_in,
                                                     out_f64, ld_f64, out_f32, pitch_f32);
    HICGAT_CHECK_LAUNCH("cont2dist_apply_kernel");
    return HICGAT_OK;
}

extern "C" int hicgat_csr_count_f64(const double* adj, int64_t ld, int64_t n, int with_self_loops, int64_t* rowcount, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(adj && rowcount && n > 0 && n < (1ll << 30) && ld >= n, "hicgat_csr_count_f64: bad arguments");
    csr_count_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(adj, ld, (int)n, with_self_loops, rowcount);
    HICGAT_CHECK_LAUNCH("csr_count_kernel");
    return HICGAT_OK;
}

extern "C" int hicgat_csr_scan_i64(const int64_t* rowcount, int64_t n, int64_t* rowptr, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(rowcount && rowptr && n >= 0, "hicgat_csr_scan_i64: bad arguments");
    scan_i64_kernel<<<1, 1024, 0, stream>>>(rowcount, n, rowptr);
    HICGAT_CHECK_LAUNCH("scan_i64_kernel");
    return HICGAT_OK;
}

extern "C" int hicgat_csr_fill_f64(const double* adj, int64_t ld, int64_t n, int with_self_loops, const int64_t* rowptr,
                                   int64_t* col, float* val, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(adj && rowptr && col && val && n > 0 && n < (1ll << 30) && ld >= n, "hicgat_csr_fill_f64: bad arguments");
    csr_fill_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(adj, ld, (int)n, with_self_loops, rowptr, col, val);
    HICGAT_CHECK_LAUNCH("csr_fill_kernel");
    return HICGAT_OK;
}

extern "C" int hicgat_csr_pack_i32(const int64_t* rowptr, const int64_t* col, int64_t n, int64_t nnz, int32_t* rowptr32,
                                   int32_t* col32, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(rowptr && rowptr32 && (nnz == 0 || (col && col32)) && n >= 0 && nnz >= 0 && nnz < (1ll << 31), "hicgat_csr_pack_i32: bad arguments");
    pack_i32_kernel<<<148 * 4, 256, 0, stream>>>(rowptr, col, n, nnz, rowptr32, col32);
    HICGAT_CHECK_LAUNCH("pack_i32_kernel");
    return HICGAT_OK;
}

extern "C" int hicgat_sage_norm_values(const int32_t* rowptr, const int32_t* col, const float* val, int64_t n, float* colsum,
                                       float* norm_val, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    (void)col;
    HICGAT_REQUIRE(rowptr && val && colsum && norm_val && n > 0, "hicgat_sage_norm_values: bad arguments");
    sage_norm_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(rowptr, val, (int)n, colsum, norm_val);
    HICGAT_CHECK_LAUNCH("sage_norm_kernel");
    return HICGAT_OK;
}

extern "C" int hicgat_csr_transpose_perm(const int32_t* rowptr, const int32_t* col, int64_t n, int32_t* perm, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(rowptr && col && perm && n > 0, "hicgat_csr_transpose_perm: bad arguments");
    transpose_perm_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(rowptr, col, (int)n, perm);
    HICGAT_CHECK_LAUNCH("transpose_perm_kernel");
    return HICGAT_OK;
}

extern "C" int hicgat_csr_add_self_loops_i32(const int32_t* rowptr, const int32_t* col, int64_t n, int32_t* out_rowptr,
                                             int32_t* out_col, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(rowptr && col && out_rowptr && out_col && n > 0 && n < (1ll << 24), "hicgat_csr_add_self_loops_i32: bad arguments");
    add_self_loops_kernel<<<(unsigned)((n + 1 + 7) / 8), 256, 0, stream>>>(rowptr, col, (int)n, out_rowptr, out_col);
    HICGAT_CHECK_LAUNCH("add_self_loops_kernel");
    return HICGAT_OK;
}
