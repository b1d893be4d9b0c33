// Fused  y = relu(LayerNorm(x; gamma, beta)) [+ residual]  and its backward for the MLP heads of the GAT net:
// `F.relu(self.norm_a(self.densea(x))) + x_initial`, `F.relu(self.norm1(self.dense1(x))) + x_initial`,
// `F.relu(self.norm2(self.dense2(x)))`  (models.py:670-690).  The Linear layers stay cuBLAS; what is fused here
// is the N x C elementwise / normalisation glue between them, which ATen runs as LayerNorm + clamp + add forward
// and threshold_backward + layer_norm_grad_input + GammaBetaBackward (380 us per call at 50k loci) backward:
// at N = 49 850 that glue was 1.5 ms of a 14.9 ms training step.  HBM-bound: forward 8..12 B, backward 12 B
// per element.  One warp per row, lanes own fixed columns, so the gamma / beta gradients accumulate in
// registers over the rows a warp visits and are combined in a fixed order (bit-reproducible).
#include "common.cuh"

namespace hicgat {
namespace {

constexpr int kLnWarps = 8;

// V = C / 32 columns per lane, moved W = min(V, 4) at a time: column(k, e) = k * 32 * W + lane * W + e
template <int V>
struct LnLane {
    static constexpr int W = V >= 4 ? 4 : V;
    static constexpr int K = V / W;
    __device__ static __forceinline__ void load(const float* __restrict__ row, int lane, float (&v)[V]) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const float* p = row + k * 32 * W + lane * W;
            if constexpr (W == 4) {
                const float4 t = *reinterpret_cast<const float4*>(p);
                v[k * 4] = t.x; v[k * 4 + 1] = t.y; v[k * 4 + 2] = t.z; v[k * 4 + 3] = t.w;
            } else if constexpr (W == 2) {
                const float2 t = *reinterpret_cast<const float2*>(p);
                v[k * 2] = t.x; v[k * 2 + 1] = t.y;
            } else {
                v[k] = *p;
            }
        }
    }
    __device__ static __forceinline__ void store(float* __restrict__ row, int lane, const float (&v)[V]) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float* p = row + k * 32 * W + lane * W;
            if constexpr (W == 4) *reinterpret_cast<float4*>(p) = make_float4(v[k * 4], v[k * 4 + 1], v[k * 4 + 2], v[k * 4 + 3]);
            else if constexpr (W == 2) *reinterpret_cast<float2*>(p) = make_float2(v[k * 2], v[k * 2 + 1]);
            else *p = v[k];
        }
    }
    __device__ static __forceinline__ int column(int lane, int idx) { return (idx / W) * 32 * W + lane * W + (idx % W); }
};

template <int V>
__global__ void __launch_bounds__(kLnWarps * 32) ln_relu_add_fwd_kernel(const float* __restrict__ x, const float* __restrict__ res,
                                                                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                                        int n, float* __restrict__ y, float* __restrict__ mean,
                                                                        float* __restrict__ rstd) {
    constexpr int C = V * 32;
    const int lane = threadIdx.x & 31, row = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
    if (row >= n) return;
    float v[V], g[V], b[V];
    LnLane<V>::load(x + (size_t)row * C, lane, v);
    LnLane<V>::load(gamma, lane, g);
    LnLane<V>::load(beta, lane, b);
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) s += v[k];
    const float mu = warp_sum(s) * (1.0f / C);
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < V; ++k) {
        v[k] -= mu;
        ss = fmaf(v[k], v[k], ss);
    }
    const float r = rsqrtf(warp_sum(ss) * (1.0f / C) + eps);   // biased variance, like torch.nn.LayerNorm
    float o[V];
    if (res) LnLane<V>::load(res + (size_t)row * C, lane, o);
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const float pre = fmaf(v[k] * r, g[k], b[k]);
        o[k] = fmaxf(pre, 0.f) + (res ? o[k] : 0.f);
    }
    LnLane<V>::store(y + (size_t)row * C, lane, o);
    if (lane == 0) {
        mean[row] = mu;
        rstd[row] = r;
    }
}

// dx for the rows of this warp; per-CTA partials of d gamma / d beta -> part[cta][2C]
template <int V>
__global__ void __launch_bounds__(kLnWarps * 32) ln_relu_add_bwd_kernel(const float* __restrict__ gy, const float* __restrict__ x,
                                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                        const float* __restrict__ gamma, const float* __restrict__ beta, int n,
                                                                        float* __restrict__ dx, float* __restrict__ part) {
    constexpr int C = V * 32;
    __shared__ float s_part[kLnWarps][2 * C];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float g[V], b[V], dg[V], db[V];
    LnLane<V>::load(gamma, lane, g);
    LnLane<V>::load(beta, lane, b);
#pragma unroll
    for (int k = 0; k < V; ++k) dg[k] = db[k] = 0.f;
    for (int row = blockIdx.x * kLnWarps + warp; row < n; row += gridDim.x * kLnWarps) {
        float v[V], go[V];
        LnLane<V>::load(x + (size_t)row * C, lane, v);
        LnLane<V>::load(gy + (size_t)row * C, lane, go);
        const float mu = mean[row], r = rstd[row];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < V; ++k) {
            v[k] = (v[k] - mu) * r;                                   // x_hat
            const float pre = fmaf(v[k], g[k], b[k]);
            go[k] = pre > 0.f ? go[k] : 0.f;                          // relu backward: grad where the OUTPUT is > 0
            db[k] += go[k];
            dg[k] = fmaf(go[k], v[k], dg[k]);
            go[k] *= g[k];                                            // d x_hat
            s1 += go[k];
            s2 = fmaf(go[k], v[k], s2);
        }
        s1 = warp_sum(s1) * (1.0f / C);
        s2 = warp_sum(s2) * (1.0f / C);
#pragma unroll
        for (int k = 0; k < V; ++k) go[k] = r * (go[k] - s1 - v[k] * s2);
        LnLane<V>::store(dx + (size_t)row * C, lane, go);
    }
#pragma unroll
    for (int k = 0; k < V; ++k) {
        const int c = LnLane<V>::column(lane, k);
        s_part[warp][c] = dg[k];
        s_part[warp][C + c] = db[k];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += kLnWarps * 32) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < kLnWarps; ++w) s += s_part[w][c];
        part[(size_t)blockIdx.x * (2 * C) + c] = s;
    }
}

// d gamma [C] | d beta [C]: 32 columns per block, 32 threads per column each adding the partials blk = ty, ty + 32, ...
// in order, then the 32 sub-sums in order (fixed association => bit-reproducible)
__global__ void __launch_bounds__(1024) ln_param_reduce_kernel(const float* __restrict__ part, int nblk, int c2, int C, float* __restrict__ dgamma,
                                                               float* __restrict__ dbeta) {
    __shared__ float s_sum[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float s = 0.f;
    if (c < c2) {
        for (int b = ty; b < nblk; b += 32) s += part[(size_t)b * c2 + c];
    }
    s_sum[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && c < c2) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < 32; ++k) t += s_sum[k][tx];
        if (c < C) dgamma[c] = t;
        else dbeta[c - C] = t;
    }
}

int bwd_blocks(int64_t n) {
    const int64_t need = (n + kLnWarps - 1) / kLnWarps;
    const int64_t cap = 148 * 4;  // 4 CTAs of 8 warps per SM keep ~48 rows per SM in flight
    return (int)(need < cap ? need : cap);
}

}  // namespace
}  // namespace hicgat

using namespace hicgat;

#define HICGAT_LN_DISPATCH(c, CALL)                                                      \
    switch (c) {                                                                         \
        case 32: { constexpr int V = 1; CALL; } break;                                   \
        case 64: { constexpr int V = 2; CALL; } break;                                   \
        case 128: { constexpr int V = 4; CALL; } break;                                  \
        case 256: { constexpr int V = 8; CALL; } break;                                  \
        case 512: { constexpr int V = 16; CALL; } break;                                 \
        default:                                                                         \
            set_error("hicgat_ln_relu_add: unsupported width %d (32, 64, 128, 256, 512)", (int)(c)); \
            return HICGAT_ERR_INVALID;                                                   \
    }

extern "C" int hicgat_ln_relu_add_fwd(const float* x, const float* residual, const float* gamma, const float* beta, float eps, int64_t n,
                                      int32_t c, float* y, float* mean, float* rstd, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(x && gamma && beta && y && mean && rstd && n > 0 && n < (1ll << 30), "hicgat_ln_relu_add_fwd: bad arguments");
    HICGAT_REQUIRE(aligned16(x) && aligned16(residual) && aligned16(gamma) && aligned16(beta) && aligned16(y), "hicgat_ln_relu_add_fwd: pointers must be 16-byte aligned");
    const unsigned grid = (unsigned)((n + kLnWarps - 1) / kLnWarps);
    HICGAT_LN_DISPATCH(c, (ln_relu_add_fwd_kernel<V><<<grid, kLnWarps * 32, 0, stream>>>(x, residual, gamma, beta, eps, (int)n, y, mean, rstd)));
    HICGAT_CHECK_LAUNCH("ln_relu_add_fwd_kernel");
    return HICGAT_OK;
}

extern "C" size_t hicgat_ln_relu_add_bwd_workspace_bytes(int64_t n, int32_t c) {
    if (n <= 0 || c <= 0) return 0;
    return sizeof(float) * 2 * (size_t)c * (size_t)bwd_blocks(n);
}

extern "C" int hicgat_ln_relu_add_bwd(const float* grad_y, const float* x, const float* mean, const float* rstd, const float* gamma,
                                      const float* beta, int64_t n, int32_t c, float* dx, float* dgamma, float* dbeta, void* workspace,
                                      size_t workspace_bytes, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(grad_y && x && mean && rstd && gamma && beta && dx && dgamma && dbeta && workspace && n > 0 && n < (1ll << 30), "hicgat_ln_relu_add_bwd: bad arguments");
    HICGAT_REQUIRE(aligned16(grad_y) && aligned16(x) && aligned16(gamma) && aligned16(beta) && aligned16(dx), "hicgat_ln_relu_add_bwd: pointers must be 16-byte aligned");
    const int nblk = bwd_blocks(n);
    if (workspace_bytes < hicgat_ln_relu_add_bwd_workspace_bytes(n, c)) {
        set_error("hicgat_ln_relu_add_bwd: workspace %zu < required %zu", workspace_bytes, hicgat_ln_relu_add_bwd_workspace_bytes(n, c));
        return HICGAT_ERR_WORKSPACE;
    }
    float* part = static_cast<float*>(workspace);
    HICGAT_LN_DISPATCH(c, (ln_relu_add_bwd_kernel<V><<<nblk, kLnWarps * 32, 0, stream>>>(grad_y, x, mean, rstd, gamma, beta, (int)n, dx, part)));
    HICGAT_CHECK_LAUNCH("ln_relu_add_bwd_kernel");
    ln_param_reduce_kernel<<<(2 * c + 31) / 32, 1024, 0, stream>>>(part, nblk, 2 * c, c, dgamma, dbeta);
    HICGAT_CHECK_LAUNCH("ln_param_reduce_kernel");
    return HICGAT_OK;
}
