// Support for NON-symmetric targets.
//
// The fused loss kernels (pairloss.cu) use the column-side sum g_j = sum_i w_ij (x_j - x_i) as the
// complete gradient, which holds only for t_ij = t_ji.  The reference does not symmetrise what it is
// given: utils.load_input keeps `y` as passed (utils.py:29-35), MSELoss runs over the full matrix
// (HiC-GNN_main.py:127), and utils.convert_to_matrix's `triu + tril(mat.T, 1)` (utils.py:21) is
// asymmetric on the first off-diagonal when a list carries lower-triangle records.  Autograd then gives
//     dL/dx_k = (2/N^2) [ sum_i (d_ik - t_ik)(x_k - x_i)/d_ik  +  sum_j (d_kj - t_kj)(x_k - x_j)/d_kj ]
//              = column-side term (what the fused kernel accumulates)  +  row-side term (this file).
// Two entry points:
//   hicgat_asymmetry_f32 / _f64   max |M_ij - M_ji| over rows [r0, r1) of a full matrix: run ONCE when a
//                                 target is built, so that the symmetric fast path is never taken blindly;
//   hicgat_pairloss_rowside_add   grad[i] += c * sum_j (d_ij - t_ij)/d_ij (x_i - x_j) for i in [r0, r1):
//                                 the row-side term for targets that failed the check (a second pass over
//                                 the row block: asymmetric targets are the exception, not the hot path).
#include "common.cuh"

namespace hicgat {
namespace {

// 32 x 32 tiles; block (bx, by) compares tile (rows r0+32*by.., cols 32*bx..) with its mirror image.
template <typename T>
__global__ void __launch_bounds__(256) asymmetry_kernel(const T* __restrict__ m, int64_t ld, int n, int r0, int r1, unsigned long long* __restrict__ out) {
    __shared__ T tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int row0 = r0 + blockIdx.y * 32, col0 = blockIdx.x * 32;
    // mirror tile: rows col0.., cols row0..
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = col0 + ty + 8 * k, c = row0 + tx;
        tile[ty + 8 * k][tx] = (r < n && c < r1) ? m[(size_t)r * ld + c] : T(0);
    }
    __syncthreads();
    double worst = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int r = row0 + ty + 8 * k, c = col0 + tx;
        if (r < r1 && c < n) {
            const double a = (double)m[(size_t)r * ld + c], b = (double)tile[tx][ty + 8 * k];
            const double d = fabs(a - b);
            // NaN on one side only, or different infinities, count as asymmetric; NaN on both sides does not
            const bool an = a != a, bn = b != b;
            if (an != bn) worst = 1.0 / 0.0;
            else if (!an && a != b) worst = fmax(worst, d != d ? 1.0 / 0.0 : d);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) worst = fmax(worst, __shfl_xor_sync(0xffffffffu, worst, o));
    if (tx == 0 && worst > 0.0) atomicMax(out, (unsigned long long)__double_as_longlong(worst));  // non-negative doubles order like their bit patterns
}

template <typename T>
int asymmetry_impl(const T* m, int64_t ld, int64_t n, int64_t r0, int64_t r1, double* out, cudaStream_t stream, const char* what) {
    HICGAT_REQUIRE(m && out && n > 0 && n < (1ll << 30) && ld >= n && r0 >= 0 && r1 >= r0 && r1 <= n, "%s: bad arguments", what);
    HICGAT_CUDA(cudaMemsetAsync(out, 0, sizeof(double), stream));
    if (r1 == r0) return HICGAT_OK;
    dim3 grid((unsigned)((n + 31) / 32), (unsigned)((r1 - r0 + 31) / 32));
    asymmetry_kernel<T><<<grid, 256, 0, stream>>>(m, ld, (int)n, (int)r0, (int)r1, reinterpret_cast<unsigned long long*>(out));
    HICGAT_CHECK_LAUNCH("asymmetry_kernel");
    return HICGAT_OK;
}

// warp per row i; lanes stride over the columns with 128-bit loads.  Same distance arithmetic as the fused
// kernel (d^2 + 1e-30 in the FMA chain, rsqrt.approx): w = c (d - t) / d, zero on the diagonal (dx = 0).
__global__ void __launch_bounds__(256) pairloss_rowside_kernel(const float* __restrict__ coords, const float* __restrict__ target, int64_t pitch, int n,
                                                               int r0, int r1, float c, float* __restrict__ grad) {
    const int i = r0 + blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= r1) return;
    const float xi = coords[(size_t)i * 3], yi = coords[(size_t)i * 3 + 1], zi = coords[(size_t)i * 3 + 2];
    const float* trow = target + (size_t)(i - r0) * pitch;
    float gx = 0.f, gy = 0.f, gz = 0.f;
    for (int j0 = lane * 4; j0 < n; j0 += 128) {
        const float4 t4 = ldg_stream_f4(trow + j0);  // pitch is a multiple of 4 and >= n: in bounds, padding masked below
        const float tv[4] = {t4.x, t4.y, t4.z, t4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int j = min(j0 + k, n - 1);
            const float dx = xi - __ldg(coords + (size_t)j * 3), dy = yi - __ldg(coords + (size_t)j * 3 + 1), dz = zi - __ldg(coords + (size_t)j * 3 + 2);
            const float d2 = fmaf(dx, dx, fmaf(dy, dy, fmaf(dz, dz, 1e-30f)));
            const float rs = rsqrt_approx(d2);
            const float w = (j0 + k < n) ? c * (d2 * rs - tv[k]) * rs : 0.f;
            gx = fmaf(w, dx, gx); gy = fmaf(w, dy, gy); gz = fmaf(w, dz, gz);
        }
    }
    gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
    if (lane == 0) { grad[(size_t)i * 3] += gx; grad[(size_t)i * 3 + 1] += gy; grad[(size_t)i * 3 + 2] += gz; }
}

}  // namespace
}  // namespace hicgat

using namespace hicgat;

extern "C" int hicgat_asymmetry_f32(const float* mat, int64_t ld, int64_t n, int64_t r0, int64_t r1, double* out_max, hicgat_stream_t stream) {
    return asymmetry_impl<float>(mat, ld, n, r0, r1, out_max, static_cast<cudaStream_t>(stream), "hicgat_asymmetry_f32");
}

extern "C" int hicgat_asymmetry_f64(const double* mat, int64_t ld, int64_t n, int64_t r0, int64_t r1, double* out_max, hicgat_stream_t stream) {
    return asymmetry_impl<double>(mat, ld, n, r0, r1, out_max, static_cast<cudaStream_t>(stream), "hicgat_asymmetry_f64");
}

extern "C" int hicgat_pairloss_rowside_add(const float* coords, const float* target, int64_t pitch, int64_t n, int64_t r0, int64_t r1, float c,
                                           float* grad, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(n > 0 && n < (1ll << 30) && r0 >= 0 && r1 >= r0 && r1 <= n, "hicgat_pairloss_rowside_add: bad n/r0/r1");
    HICGAT_REQUIRE(coords && grad && (target || r0 == r1), "hicgat_pairloss_rowside_add: null pointer");
    HICGAT_REQUIRE(pitch >= n && (pitch % 4) == 0 && aligned16(target), "hicgat_pairloss_rowside_add: pitch must be >= n and a multiple of 4, target 16-byte aligned");
    if (r1 == r0) return HICGAT_OK;
    pairloss_rowside_kernel<<<(unsigned)((r1 - r0 + 7) / 8), 256, 0, stream>>>(coords, target, pitch, (int)n, (int)r0, (int)r1, c, grad);
    HICGAT_CHECK_LAUNCH("pairloss_rowside_kernel");
    return HICGAT_OK;
}
