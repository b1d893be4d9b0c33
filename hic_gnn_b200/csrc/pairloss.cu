// Fused pairwise-distance loss (forward + backward in one pass over the target block).
//
// Replaces torch.cdist + MSELoss / L1 / Pearson-moment glue of the reference loops
// (models.py:39, HiC-GNN_main.py:127, HiC_GAT_generalize_directly.py:210-225,
// train_and_test_same_res_GAT_node2vec.py:131-134) without materialising the N x N
// prediction.  HBM-bound: 4 B per ordered pair (one f32 target element), d = 3.
//
// Work decomposition
//   CTA  = 8 warps x (128 columns) x (RB rows); grid = (column strips, row chunks).
//   lane = 4 consecutive columns j (one 128-bit streaming load per row), kept as two packed
//          f32x2 pairs so the arithmetic runs on FADD2/FMUL2/FFMA2;
//   warp = row groups of U=8 rows, interleaved over the CTA's 8 warps (8 loads in flight/lane);
//   x_j and the column-side gradient accumulators live in registers for the whole CTA
//   lifetime, x_i (warp-uniform) is broadcast from shared memory.
// Because the target is symmetric the column-side sum  g_j = sum_i w_ij (x_j - x_i)  is the
// complete gradient: no row-side reduction, no atomics.  Row chunks are combined by the last
// CTA to finish each column strip (fixed summation order => bit-reproducible).
#include "common.cuh"

namespace hicgat {
namespace {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kCols = 128;  // columns per CTA strip (32 lanes x 4)
constexpr int kU = 8;       // rows per group (loads in flight per lane)
constexpr int kNM = HICGAT_PAIR_NMOM;

struct Acc {
    f2 gx[2], gy[2], gz[2];          // column-side gradient, 2 column pairs
    f2 see;                          // sum (d-t)^2, all pairs
    f2 sd, sdd, st, stt, sdt, seu;   // upper-triangle moments
    float sabs0, sabs1;              // upper-triangle sum |d-t|
};

template <uint32_t MODE>
__device__ __forceinline__ f2 grad_weight(f2 e, f2 rs, float c_mse, float c_l1) {
    if constexpr ((MODE & 3u) == HICGAT_PAIR_GRAD_MSE) {
        return f2_mul(e, rs);
    } else if constexpr ((MODE & 3u) == HICGAT_PAIR_GRAD_L1) {
        float ea, eb, ra, rb;
        f2_unpack(e, ea, eb);
        f2_unpack(rs, ra, rb);
        return f2_pack(copysignf(ra, ea), copysignf(rb, eb));
    } else if constexpr ((MODE & 3u) == 3u) {
        float ea, eb;
        f2_unpack(e, ea, eb);
        f2 s = f2_pack(copysignf(c_l1, ea), copysignf(c_l1, eb));
        return f2_mul(f2_fma(e, f2_pack(c_mse, c_mse), s), rs);
    } else {
        return 0ull;
    }
}

// One row x one column pair, no masks.  UPPER: this (row, pair) lies strictly above the diagonal.
template <uint32_t MODE, bool UPPER>
__device__ __forceinline__ void pair_fast(Acc& a, int p, f2 xjx, f2 xjy, f2 xjz, f2 xix, f2 xiy,
                                          f2 xiz, f2 t, float c_mse, float c_l1) {
    f2 dx = f2_sub(xjx, xix), dy = f2_sub(xjy, xiy), dz = f2_sub(xjz, xiz);
    f2 d2 = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
    float d2a, d2b;
    f2_unpack(d2, d2a, d2b);
    f2 rs = f2_pack(rsqrt_approx(fmaxf(d2a, 1e-30f)), rsqrt_approx(fmaxf(d2b, 1e-30f)));
    f2 d = f2_mul(d2, rs);
    f2 e = f2_sub(d, t);
    a.see = f2_fma(e, e, a.see);
    if constexpr ((MODE & 3u) != 0) {
        f2 w = grad_weight<MODE>(e, rs, c_mse, c_l1);
        a.gx[p] = f2_fma(w, dx, a.gx[p]);
        a.gy[p] = f2_fma(w, dy, a.gy[p]);
        a.gz[p] = f2_fma(w, dz, a.gz[p]);
    }
    if constexpr (UPPER && (MODE & HICGAT_PAIR_MOMENTS)) {
        float ea, eb;
        f2_unpack(e, ea, eb);
        a.sabs0 += fabsf(ea);
        a.sabs1 += fabsf(eb);
        a.sd = f2_add(a.sd, d);
        a.sdd = f2_add(a.sdd, d2);
        a.st = f2_add(a.st, t);
        a.stt = f2_fma(t, t, a.stt);
        a.sdt = f2_fma(d, t, a.sdt);
        a.seu = f2_fma(e, e, a.seu);
    }
}

// Masked variant for edge strips (columns >= n) and diagonal-crossing groups.
// mv: 1 for valid columns; mu: 1 where additionally row < col.
template <uint32_t MODE>
__device__ __forceinline__ void pair_masked(Acc& a, int p, f2 xjx, f2 xjy, f2 xjz, f2 xix, f2 xiy,
                                            f2 xiz, f2 t, f2 mv, f2 mu, float c_mse, float c_l1) {
    f2 dx = f2_sub(xjx, xix), dy = f2_sub(xjy, xiy), dz = f2_sub(xjz, xiz);
    f2 d2 = f2_fma(dz, dz, f2_fma(dy, dy, f2_mul(dx, dx)));
    float d2a, d2b;
    f2_unpack(d2, d2a, d2b);
    f2 rs = f2_pack(rsqrt_approx(fmaxf(d2a, 1e-30f)), rsqrt_approx(fmaxf(d2b, 1e-30f)));
    f2 d = f2_mul(d2, rs);
    f2 e = f2_mul(f2_sub(d, t), mv);
    a.see = f2_fma(e, e, a.see);
    if constexpr ((MODE & 3u) != 0) {
        f2 w = f2_mul(grad_weight<MODE>(e, rs, c_mse, c_l1), mv);
        a.gx[p] = f2_fma(w, dx, a.gx[p]);
        a.gy[p] = f2_fma(w, dy, a.gy[p]);
        a.gz[p] = f2_fma(w, dz, a.gz[p]);
    }
    if constexpr ((MODE & HICGAT_PAIR_MOMENTS) != 0) {
        f2 eu = f2_mul(e, mu), du = f2_mul(d, mu), tu = f2_mul(t, mu);
        float ea, eb;
        f2_unpack(eu, ea, eb);
        a.sabs0 += fabsf(ea);
        a.sabs1 += fabsf(eb);
        a.sd = f2_add(a.sd, du);
        a.sdd = f2_fma(du, du, a.sdd);
        a.st = f2_add(a.st, tu);
        a.stt = f2_fma(tu, tu, a.stt);
        a.sdt = f2_fma(du, tu, a.sdt);
        a.seu = f2_fma(eu, eu, a.seu);
    }
}

struct Params {
    const float* coords;
    const float* target;  // points at row r0
    int64_t pitch;
    int n, r0, r1, rb, nstrips, nchunks;
    float c_mse, c_l1;
    double* moments;
    float* grad;
    double* grad64;        // optional f64 copy of grad (packed all-reduce buffer)
    float* gpart;          // [nchunks][nstrips][384]
    double* mpart;         // [nchunks*nstrips][kNM]
    unsigned* strip_count; // [nstrips]
    unsigned* done_count;  // [1]
};

template <uint32_t MODE>
__global__ void __launch_bounds__(kThreads, 2) pairloss_kernel(const Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* s_xy = reinterpret_cast<float4*>(smem_raw);                  // [rb] (x,x,y,y)
    float2* s_z = reinterpret_cast<float2*>(smem_raw + sizeof(float4) * P.rb);  // [rb] (z,z)
    __shared__ float s_g[kWarps][kCols * 3 + 4];
    __shared__ double s_m[kWarps][kNM];
    __shared__ unsigned s_ticket[2];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int strip = blockIdx.x, chunk = blockIdx.y;
    const int n = P.n;
    const int col0 = strip * kCols + lane * 4;
    const int row_begin = P.r0 + chunk * P.rb;
    const int row_end = min(row_begin + P.rb, P.r1);
    const int nrows = row_end - row_begin;
    const bool edge = (strip + 1) * kCols > n;

    // stage this chunk's row coordinates, duplicated for packed broadcast
    for (int r = threadIdx.x; r < nrows; r += kThreads) {
        const float* c = P.coords + (size_t)(row_begin + r) * 3;
        float x = c[0], y = c[1], z = c[2];
        s_xy[r] = make_float4(x, x, y, y);
        s_z[r] = make_float2(z, z);
    }
    // this lane's 4 columns
    float cj[4][3];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int c = min(col0 + k, n - 1);
#pragma unroll
        for (int q = 0; q < 3; ++q) cj[k][q] = P.coords[(size_t)c * 3 + q];
    }
    f2 xjx[2], xjy[2], xjz[2], mv[2];
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        xjx[p] = f2_pack(cj[2 * p][0], cj[2 * p + 1][0]);
        xjy[p] = f2_pack(cj[2 * p][1], cj[2 * p + 1][1]);
        xjz[p] = f2_pack(cj[2 * p][2], cj[2 * p + 1][2]);
        mv[p] = f2_pack(col0 + 2 * p < n ? 1.f : 0.f, col0 + 2 * p + 1 < n ? 1.f : 0.f);
    }
    Acc a;
#pragma unroll
    for (int p = 0; p < 2; ++p) a.gx[p] = a.gy[p] = a.gz[p] = 0ull;
    a.see = a.sd = a.sdd = a.st = a.stt = a.sdt = a.seu = 0ull;
    a.sabs0 = a.sabs1 = 0.f;
    __syncthreads();

    const bool can_load = col0 + 3 < P.pitch;  // pitch is a multiple of 4
    const float* tbase = P.target + (size_t)(row_begin - P.r0) * P.pitch + col0;
    const int strip_lo = strip * kCols, strip_hi = strip_lo + kCols - 1;
    const int ngroups = (nrows + kU - 1) / kU;

    for (int g = warp; g < ngroups; g += kWarps) {
        const int rl = g * kU;                      // local row of the group
        const int rows_here = min(kU, nrows - rl);
        const int rg = row_begin + rl;              // global row
        float4 t[kU];
        if (rows_here == kU && can_load) {
#pragma unroll
            for (int u = 0; u < kU; ++u) t[u] = ldg_stream_f4(tbase + (size_t)(rl + u) * P.pitch);
        } else {
#pragma unroll
            for (int u = 0; u < kU; ++u)
                t[u] = (u < rows_here && can_load) ? ldg_stream_f4(tbase + (size_t)(rl + u) * P.pitch)
                                                   : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const bool fast = !edge && rows_here == kU;
        if (fast && rg + kU - 1 < strip_lo) {  // strictly above the diagonal
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                float4 xy = s_xy[rl + u];
                float2 zz = s_z[rl + u];
                f2 xix = f2_pack(xy.x, xy.y), xiy = f2_pack(xy.z, xy.w), xiz = f2_pack(zz.x, zz.y);
                pair_fast<MODE, true>(a, 0, xjx[0], xjy[0], xjz[0], xix, xiy, xiz, f2_pack(t[u].x, t[u].y), P.c_mse, P.c_l1);
                pair_fast<MODE, true>(a, 1, xjx[1], xjy[1], xjz[1], xix, xiy, xiz, f2_pack(t[u].z, t[u].w), P.c_mse, P.c_l1);
            }
        } else if (fast && rg > strip_hi) {    // strictly below
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                float4 xy = s_xy[rl + u];
                float2 zz = s_z[rl + u];
                f2 xix = f2_pack(xy.x, xy.y), xiy = f2_pack(xy.z, xy.w), xiz = f2_pack(zz.x, zz.y);
                pair_fast<MODE, false>(a, 0, xjx[0], xjy[0], xjz[0], xix, xiy, xiz, f2_pack(t[u].x, t[u].y), P.c_mse, P.c_l1);
                pair_fast<MODE, false>(a, 1, xjx[1], xjy[1], xjz[1], xix, xiy, xiz, f2_pack(t[u].z, t[u].w), P.c_mse, P.c_l1);
            }
        } else {
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                if (u < rows_here) {
                    const int r = rg + u;
                    float4 xy = s_xy[rl + u];
                    float2 zz = s_z[rl + u];
                    f2 xix = f2_pack(xy.x, xy.y), xiy = f2_pack(xy.z, xy.w), xiz = f2_pack(zz.x, zz.y);
                    f2 mu0 = f2_pack((col0 + 0 < n && r < col0 + 0) ? 1.f : 0.f, (col0 + 1 < n && r < col0 + 1) ? 1.f : 0.f);
                    f2 mu1 = f2_pack((col0 + 2 < n && r < col0 + 2) ? 1.f : 0.f, (col0 + 3 < n && r < col0 + 3) ? 1.f : 0.f);
                    pair_masked<MODE>(a, 0, xjx[0], xjy[0], xjz[0], xix, xiy, xiz, f2_pack(t[u].x, t[u].y), mv[0], mu0, P.c_mse, P.c_l1);
                    pair_masked<MODE>(a, 1, xjx[1], xjy[1], xjz[1], xix, xiy, xiz, f2_pack(t[u].z, t[u].w), mv[1], mu1, P.c_mse, P.c_l1);
                }
            }
        }
    }

    // ---- CTA combine: gradients (fixed warp order) and moments (f64) ----
    if constexpr ((MODE & 3u) != 0) {
#pragma unroll
        for (int p = 0; p < 2; ++p) {
            float x0, x1, y0, y1, z0, z1;
            f2_unpack(a.gx[p], x0, x1);
            f2_unpack(a.gy[p], y0, y1);
            f2_unpack(a.gz[p], z0, z1);
            float* dst = &s_g[warp][(lane * 4 + 2 * p) * 3];
            dst[0] = x0; dst[1] = y0; dst[2] = z0;
            dst[3] = x1; dst[4] = y1; dst[5] = z1;
        }
    }
    {
        double m[kNM];
        m[0] = (double)f2_hsum(a.see);
        if constexpr ((MODE & HICGAT_PAIR_MOMENTS) != 0) {
            m[1] = (double)a.sabs0 + (double)a.sabs1;
            m[2] = (double)f2_hsum(a.sd);
            m[3] = (double)f2_hsum(a.sdd);
            m[4] = (double)f2_hsum(a.st);
            m[5] = (double)f2_hsum(a.stt);
            m[6] = (double)f2_hsum(a.sdt);
            m[7] = (double)f2_hsum(a.seu);
        } else {
#pragma unroll
            for (int k = 1; k < kNM; ++k) m[k] = 0.0;
        }
#pragma unroll
        for (int k = 0; k < kNM; ++k) {
            if (k == 0 || (MODE & HICGAT_PAIR_MOMENTS)) m[k] = warp_sum(m[k]);
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < kNM; ++k) s_m[warp][k] = m[k];
        }
    }
    __syncthreads();
    const int cta = chunk * P.nstrips + strip;
    if constexpr ((MODE & 3u) != 0) {
        float* gp = P.gpart + (size_t)cta * (kCols * 3);
        for (int i = threadIdx.x; i < kCols * 3; i += kThreads) {
            float s = 0.f;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += s_g[w][i];
            __stcg(gp + i, s);
        }
    }
    if (threadIdx.x < kNM) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += s_m[w][threadIdx.x];
        __stcg(P.mpart + (size_t)cta * kNM + threadIdx.x, s);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        s_ticket[0] = atomicAdd(P.strip_count + strip, 1u);
        s_ticket[1] = atomicAdd(P.done_count, 1u);
    }
    __syncthreads();
    const bool last_in_strip = s_ticket[0] == (unsigned)(P.nchunks - 1);
    const bool last_overall = s_ticket[1] == (unsigned)(P.nchunks * P.nstrips - 1);
    if (!last_in_strip && !last_overall) return;
    __threadfence();
    if constexpr ((MODE & 3u) != 0) {
        if (last_in_strip) {
            const float scale = ((MODE & 3u) == HICGAT_PAIR_GRAD_MSE) ? P.c_mse
                                : ((MODE & 3u) == HICGAT_PAIR_GRAD_L1) ? P.c_l1 : 1.f;
            for (int i = threadIdx.x; i < kCols * 3; i += kThreads) {
                double s = 0.0;
                for (int c = 0; c < P.nchunks; ++c)
                    s += (double)__ldcg(P.gpart + ((size_t)c * P.nstrips + strip) * (kCols * 3) + i);
                const int col = strip * kCols + i / 3;
                if (col < n) {
                    const double v = s * (double)scale;
                    if (P.grad) P.grad[(size_t)strip * kCols * 3 + i] = (float)v;
                    if (P.grad64) P.grad64[(size_t)strip * kCols * 3 + i] = (double)(float)v;
                }
            }
        }
    }
    if (last_overall) {
        // fixed-order f64 reduction of all CTA moment partials
        __shared__ double s_red[kThreads / 32][kNM];
        double m[kNM];
#pragma unroll
        for (int k = 0; k < kNM; ++k) m[k] = 0.0;
        const int total = P.nchunks * P.nstrips;
        for (int c = threadIdx.x; c < total; c += kThreads) {
#pragma unroll
            for (int k = 0; k < kNM; ++k) m[k] += __ldcg(P.mpart + (size_t)c * kNM + k);
        }
#pragma unroll
        for (int k = 0; k < kNM; ++k) m[k] = warp_sum(m[k]);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < kNM; ++k) s_red[warp][k] = m[k];
        }
        __syncthreads();
        if (threadIdx.x < kNM) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < kWarps; ++w) s += s_red[w][threadIdx.x];
            P.moments[threadIdx.x] = s;
        }
    }
}

// ------------------------------------------------------------------ materialising variant
__global__ void pairdist_fwd_kernel(const float* __restrict__ coords, int n, float* __restrict__ dist, int64_t pitch) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i0 = blockIdx.y * 16;
    if (j >= n) return;
    const float xj = coords[j * 3], yj = coords[j * 3 + 1], zj = coords[j * 3 + 2];
    for (int i = i0; i < min(i0 + 16, n); ++i) {
        const float dx = xj - __ldg(coords + i * 3), dy = yj - __ldg(coords + i * 3 + 1), dz = zj - __ldg(coords + i * 3 + 2);
        dist[(size_t)i * pitch + j] = sqrtf(dx * dx + dy * dy + dz * dz);
    }
}

// grad_coords[i] = sum_j (G[i,j] + G[j,i]) (x_i - x_j)/d_ij ; one warp per locus i
__global__ void pairdist_bwd_kernel(const float* __restrict__ coords, int n, const float* __restrict__ G, int64_t pitch, float* __restrict__ gc) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n) return;
    const float xi = coords[i * 3], yi = coords[i * 3 + 1], zi = coords[i * 3 + 2];
    float gx = 0.f, gy = 0.f, gz = 0.f;
    for (int j = lane; j < n; j += 32) {
        const float dx = xi - coords[j * 3], dy = yi - coords[j * 3 + 1], dz = zi - coords[j * 3 + 2];
        const float d2 = dx * dx + dy * dy + dz * dz;
        if (d2 > 0.f) {
            const float w = (G[(size_t)i * pitch + j] + G[(size_t)j * pitch + i]) * rsqrtf(d2);
            gx += w * dx; gy += w * dy; gz += w * dz;
        }
    }
    gx = warp_sum(gx); gy = warp_sum(gy); gz = warp_sum(gz);
    if (lane == 0) { gc[i * 3] = gx; gc[i * 3 + 1] = gy; gc[i * 3 + 2] = gz; }
}

int g_rows_per_cta = 0;

int pick_rows_per_cta(int64_t nrows, int nstrips) {
    if (g_rows_per_cta > 0) return g_rows_per_cta;
    // largest chunk that still gives >= ~6 CTAs per SM (148 SMs); bounds the partial buffers
    const int64_t want = 148 * 6;
    int rb = 1024;
    while (rb > 64 && (int64_t)nstrips * ((nrows + rb - 1) / rb) < want) rb >>= 1;
    return rb;
}

struct Layout {
    int nstrips, rb, nchunks;
    size_t off_counts, off_gpart, off_mpart, total;
};

Layout make_layout(int64_t n, int64_t r0, int64_t r1) {
    Layout L;
    L.nstrips = (int)((n + kCols - 1) / kCols);
    const int64_t nrows = r1 - r0;
    L.rb = pick_rows_per_cta(nrows, L.nstrips);
    L.nchunks = (int)((nrows + L.rb - 1) / L.rb);
    if (L.nchunks < 1) L.nchunks = 1;
    L.off_counts = 0;
    L.off_mpart = align_up(sizeof(unsigned) * (size_t)(L.nstrips + 1), 256);
    L.off_gpart = L.off_mpart + align_up(sizeof(double) * kNM * (size_t)L.nstrips * L.nchunks, 256);
    L.total = L.off_gpart + sizeof(float) * (size_t)kCols * 3 * L.nstrips * L.nchunks;
    return L;
}

}  // namespace
}  // namespace hicgat

using namespace hicgat;

extern "C" int hicgat_pairloss_set_tuning(int rows_per_cta, int variant) {
    (void)variant;
    if (rows_per_cta != 0 && (rows_per_cta < 8 || rows_per_cta > 4096 || (rows_per_cta % 8) != 0)) {
        set_error("hicgat_pairloss_set_tuning: rows_per_cta must be 0 or a multiple of 8 in [8,4096]");
        return HICGAT_ERR_INVALID;
    }
    g_rows_per_cta = rows_per_cta;
    return HICGAT_OK;
}

extern "C" size_t hicgat_pairloss_workspace_bytes(int64_t n, int64_t r0, int64_t r1) {
    if (n <= 0 || r0 < 0 || r1 < r0 || r1 > n) return 0;
    return make_layout(n, r0, r1).total;  // for the CURRENT tuning; re-query after set_tuning
}

static int pairloss_impl(const float* coords, const float* target, int64_t pitch, int64_t n,
                         int64_t r0, int64_t r1, uint32_t mode, float c_mse, float c_l1,
                         double* moments, float* grad, double* grad64, void* workspace, size_t workspace_bytes,
                         hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(n > 0 && n < (1ll << 30) && r0 >= 0 && r1 >= r0 && r1 <= n, "hicgat_pairloss_fwd_bwd: bad n/r0/r1 (%lld,%lld,%lld)", (long long)n, (long long)r0, (long long)r1);
    // an empty row block (r0 == r1, a trailing rank of a sharded run) has no target rows to point at
    HICGAT_REQUIRE(coords && (target || r0 == r1) && moments && workspace, "hicgat_pairloss_fwd_bwd: null pointer");
    HICGAT_REQUIRE(pitch >= n && (pitch % 4) == 0, "hicgat_pairloss_fwd_bwd: pitch %lld must be >= n and a multiple of 4", (long long)pitch);
    HICGAT_REQUIRE(aligned16(target), "hicgat_pairloss_fwd_bwd: target must be 16-byte aligned");  // NULL passes
    HICGAT_REQUIRE((mode & ~7u) == 0, "hicgat_pairloss_fwd_bwd: unknown mode bits 0x%x", mode);
    HICGAT_REQUIRE(!(mode & 3u) || grad || grad64, "hicgat_pairloss_fwd_bwd: grad is NULL but a gradient mode is set");
    const Layout L = make_layout(n, r0, r1);
    if (workspace_bytes < L.total) {
        set_error("hicgat_pairloss_fwd_bwd: workspace %zu < required %zu", workspace_bytes, L.total);
        return HICGAT_ERR_WORKSPACE;
    }
    unsigned char* ws = static_cast<unsigned char*>(workspace);
    HICGAT_CUDA(cudaMemsetAsync(ws + L.off_counts, 0, sizeof(unsigned) * (size_t)(L.nstrips + 1), stream));
    if (r1 == r0) {  // empty row block: contributes nothing
        HICGAT_CUDA(cudaMemsetAsync(moments, 0, sizeof(double) * kNM, stream));
        if (grad) HICGAT_CUDA(cudaMemsetAsync(grad, 0, sizeof(float) * 3 * (size_t)n, stream));
        if (grad64) HICGAT_CUDA(cudaMemsetAsync(grad64, 0, sizeof(double) * 3 * (size_t)n, stream));
        return HICGAT_OK;
    }
    Params P;
    P.coords = coords; P.target = target; P.pitch = pitch;
    P.n = (int)n; P.r0 = (int)r0; P.r1 = (int)r1; P.rb = L.rb; P.nstrips = L.nstrips; P.nchunks = L.nchunks;
    P.c_mse = c_mse; P.c_l1 = c_l1; P.moments = moments; P.grad = grad; P.grad64 = grad64;
    P.strip_count = reinterpret_cast<unsigned*>(ws + L.off_counts);
    P.done_count = P.strip_count + L.nstrips;
    P.mpart = reinterpret_cast<double*>(ws + L.off_mpart);
    P.gpart = reinterpret_cast<float*>(ws + L.off_gpart);
    dim3 grid(L.nstrips, L.nchunks);
    const size_t smem = (sizeof(float4) + sizeof(float2)) * (size_t)L.rb;
    switch (mode) {
#define HICGAT_CASE(M) case M: pairloss_kernel<M><<<grid, kThreads, smem, stream>>>(P); break;
        HICGAT_CASE(0u) HICGAT_CASE(1u) HICGAT_CASE(2u) HICGAT_CASE(3u)
        HICGAT_CASE(4u) HICGAT_CASE(5u) HICGAT_CASE(6u) HICGAT_CASE(7u)
#undef HICGAT_CASE
    }
    HICGAT_CHECK_LAUNCH("pairloss_kernel");
    return HICGAT_OK;
}

extern "C" int hicgat_pairloss_fwd_bwd(const float* coords, const float* target, int64_t pitch, int64_t n,
                                       int64_t r0, int64_t r1, uint32_t mode, float c_mse, float c_l1,
                                       double* moments, float* grad, void* workspace, size_t workspace_bytes,
                                       hicgat_stream_t stream) {
    return pairloss_impl(coords, target, pitch, n, r0, r1, mode, c_mse, c_l1, moments, grad, nullptr, workspace, workspace_bytes, stream);
}

extern "C" int hicgat_pairloss_fwd_bwd_packed(const float* coords, const float* target, int64_t pitch, int64_t n,
                                              int64_t r0, int64_t r1, uint32_t mode, float c_mse, float c_l1,
                                              double* packed, void* workspace, size_t workspace_bytes,
                                              hicgat_stream_t stream) {
    if (!packed) {
        set_error("hicgat_pairloss_fwd_bwd_packed: null pointer");
        return HICGAT_ERR_INVALID;
    }
    return pairloss_impl(coords, target, pitch, n, r0, r1, mode, c_mse, c_l1, packed, nullptr, packed + HICGAT_PAIR_NMOM, workspace, workspace_bytes, stream);
}

extern "C" int hicgat_pairdist_fwd(const float* coords, int64_t n, float* dist, int64_t pitch, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(coords && dist && n > 0 && pitch >= n && n < (1ll << 30), "hicgat_pairdist_fwd: bad arguments");
    dim3 grid((unsigned)((n + 255) / 256), (unsigned)((n + 15) / 16));
    pairdist_fwd_kernel<<<grid, 256, 0, stream>>>(coords, (int)n, dist, pitch);
    HICGAT_CHECK_LAUNCH("pairdist_fwd_kernel");
    return HICGAT_OK;
}

extern "C" int hicgat_pairdist_bwd(const float* coords, int64_t n, const float* grad_dist, int64_t pitch,
                                   float* grad_coords, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(coords && grad_dist && grad_coords && n > 0 && pitch >= n && n < (1ll << 30), "hicgat_pairdist_bwd: bad arguments");
    pairdist_bwd_kernel<<<(unsigned)((n + 7) / 8), 256, 0, stream>>>(coords, (int)n, grad_dist, pitch, grad_coords);
    HICGAT_CHECK_LAUNCH("pairdist_bwd_kernel");
    return HICGAT_OK;
}
