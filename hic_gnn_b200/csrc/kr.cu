// Knight-Ruiz matrix balancing on the GPU: the two O(N^2) pieces of KRnorm (r_utils.R:1-93, run by
// normalize.R:1-11 as an R subprocess from e.g. HiC-GNN_main.py:85) -- SURVEY.md section 8 row f-1.
//   * y = A x            (r_utils.R:24,38,66: every inner CG step and every outer step)
//   * out = round(x_i A_ij x_j, 6)   (r_utils.R:75,90)
// Both are HBM-bound f64 streams (8 B per matrix element); the O(N) vector algebra and the scalar
// control flow of the Newton-CG iteration stay on the host side (hic_gnn_b200/kr.py).
#include "common.cuh"

namespace hicgat {
namespace {

// one warp per row, lanes stride the row with 16-byte loads; fixed reduction order => reproducible
__global__ void __launch_bounds__(256) gemv_f64_kernel(const double* __restrict__ A, int64_t ld, int n, const double* __restrict__ x,
                                                       double* __restrict__ y) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const double* row = A + (size_t)i * ld;
    double s0 = 0.0, s1 = 0.0;
    const bool vec = ((ld & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0);
    if (vec) {
        const int n2 = n >> 1;
        for (int j = lane; j < n2; j += 32) {
            const double2 a = __ldg(reinterpret_cast<const double2*>(row) + j);
            const double2 b = __ldg(reinterpret_cast<const double2*>(x) + j);
            s0 = fma(a.x, b.x, s0);
            s1 = fma(a.y, b.y, s1);
        }
        if ((n & 1) && lane == 0) s0 = fma(row[n - 1], x[n - 1], s0);
    } else {
        for (int j = lane; j < n; j += 32) s0 = fma(row[j], x[j], s0);
    }
    const double s = warp_sum(s0 + s1);
    if (lane == 0) y[i] = s;
}

__global__ void __launch_bounds__(256) kr_scale_round_kernel(const double* __restrict__ A, int64_t ld, int n, const double* __restrict__ x,
                                                             double* __restrict__ out, int64_t ldo, double scale) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const double xj = x[j];
    for (int i = blockIdx.y; i < n; i += gridDim.y) {
        const double v = (x[i] * A[(size_t)i * ld + j]) * xj;          // t(t(x*A)*x), r_utils.R:75
        out[(size_t)i * ldo + j] = rint(v * scale) / scale;            // round(., 6): half-to-even on v * 1e6, like numpy / R
    }
}

// CSR counterpart of gemv_f64_kernel for the list -> graph path that never builds the dense matrix (row f-2): y = A x with
// A given by its symmetric CSR pattern and f64 values.  One warp per row, fixed lane-strided order => reproducible.
__global__ void __launch_bounds__(256) spmv_csr_f64_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                                           const double* __restrict__ val, int n, const double* __restrict__ x, double* __restrict__ y) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    double s = 0.0;
    for (int k = rowptr[i] + lane; k < rowptr[i + 1]; k += 32) s = fma(val[k], x[col[k]], s);
    s = warp_sum(s);
    if (lane == 0) y[i] = s;
}

}  // namespace
}  // namespace hicgat

using namespace hicgat;

extern "C" int hicgat_spmv_csr_f64(const int32_t* rowptr, const int32_t* col, const double* val, int64_t n, const double* x, double* y,
                                   hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(rowptr && x && y && n > 0 && n < (1ll << 30), "hicgat_spmv_csr_f64: bad arguments");
    spmv_csr_f64_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(rowptr, col, val, (int)n, x, y);
    HICGAT_CHECK_LAUNCH("spmv_csr_f64_kernel");
    return HICGAT_OK;
}

extern "C" int hicgat_gemv_f64(const double* A, int64_t ld, int64_t n, const double* x, double* y, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(A && x && y && n > 0 && n < (1ll << 30) && ld >= n, "hicgat_gemv_f64: bad arguments");
    HICGAT_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0, "hicgat_gemv_f64: x must be 16-byte aligned");
    gemv_f64_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(A, ld, (int)n, x, y);
    HICGAT_CHECK_LAUNCH("gemv_f64_kernel");
    return HICGAT_OK;
}

extern "C" int hicgat_kr_scale_round_f64(const double* A, int64_t ld, int64_t n, const double* x, double* out, int64_t ldo,
                                         int decimals, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(A && x && out && n > 0 && n < (1ll << 30) && ld >= n && ldo >= n && decimals >= 0 && decimals <= 15, "hicgat_kr_scale_round_f64: bad arguments");
    double scale = 1.0;
    for (int d = 0; d < decimals; ++d) scale *= 10.0;
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)(n < 4096 ? n : 4096));
    kr_scale_round_kernel<<<grid, 256, 0, stream>>>(A, ld, (int)n, x, out, ldo, scale);
    HICGAT_CHECK_LAUNCH("kr_scale_round_kernel");
    return HICGAT_OK;
}
