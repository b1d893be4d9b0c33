// fp32-accurate Linear GEMMs on the 5th-generation tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM,
// operands staged by TMA) for the MLP heads and the GATConv projection of the training step
// (models.py:23-55, 634-691, 1020-1047; torch.nn.Linear = ATen addmm in the reference).
//
// Why not plain TF32: the parity contract is 1e-5 on losses and gradients, TF32 alone carries 2^-11.  The classic 3xTF32
// split  a = a_hi + a_lo  (a_hi = tf32(a), a_lo = tf32(a - a_hi)),  a.b ~ a_hi b_hi + a_hi b_lo + a_lo b_hi  keeps ~2^-21 and
// needs no special kernel once the three partial products are laid out along K:
//        D[M,N] = A'[M,3K] . B'[N,3K]^T     with  A' = [A_hi | A_hi | A_lo],  B' = [B_hi | B_lo | B_hi]
// is ONE tensor-core GEMM with an fp32 accumulator.  hicgat_split_tf32 writes A' / B' (optionally transposed, for the
// weight-gradient GEMM whose reduction runs over the rows of the activations); hicgat_gemm_tf32_tn is the GEMM.
//
// Kernel (one CTA per 128 x BN output tile and K split; 6 warps):
//   warp 0   TMA producer: cp.async.bulk.tensor.2d boxes of 128 rows x 32 tf32 (128 B, SWIZZLE_128B) of A' and BN x 32 of B'
//            into a 4-stage shared-memory ring, completion on the stage's `full` mbarrier;
//   warp 1   allocates 2 x BN TMEM columns (main product / cross terms, see the kernel); one elected lane issues 4 x tcgen05.mma.cta_group::1.kind::tf32 (128 x BN x 8) per
//            stage from shared-memory matrix descriptors, tcgen05.commit frees the stage (`empty` mbarrier) and, after the last
//            k-block, signals the accumulator (`acc` mbarrier);
//   warps 2-5 epilogue: tcgen05.ld 32 lanes x 32 columns per instruction -> registers -> (+ bias) -> coalesced 128-byte row
//            segments in global memory (or the split-K partial buffer).
// HBM/L2-bound for these shapes (K' <= 1536, N <= 512): arithmetic intensity of a 128 x 256 tile is 43 flop/B.
#include <cuda.h>

#include <algorithm>
#include <atomic>

#include "common.cuh"

namespace hicgat {
namespace {

constexpr int kBM = 128;        // rows of D per CTA = TMEM lanes
constexpr int kBK = 32;         // tf32 elements per k-block = 128 bytes = one swizzle atom row
constexpr int kUK = 8;          // K of one tcgen05.mma kind::tf32
constexpr int kGemmThreads = 192;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "GEMM_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra GEMM_DONE;\n"
        "bra GEMM_WAIT;\n"
        "GEMM_DONE:\n"
        "}\n" ::"r"(smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_addr(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(smem_addr(bar))
                 : "memory");
}
// shared-memory matrix descriptor, K-major operand, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart
__device__ __forceinline__ uint64_t umma_desc_k_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // start address           bits [ 0,14)
    d |= (uint64_t)1 << 16;                    // leading byte offset (unused for swizzled K-major) bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;          // stride byte offset       bits [32,46)
    d |= (uint64_t)1 << 46;                    // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                    // layout type: SWIZZLE_128B
    return d;
}
// instruction descriptor: D = f32, A = B = tf32, both K-major, M x N
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]),
          "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
          "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

struct GemmParams {
    float* d;            // [M, ldd] (ksplit == 1) or partials [ksplit][M][ldd]
    const float* bias;   // [N] or nullptr (added when ksplit == 1; otherwise by the reduce kernel)
    int m, n, kblocks, kb_per_split, ldd;
    int kparts;          // 3: k-block kb is a main (hi.hi) block iff kb % 3 == 0, else a cross-term block -> accumulator 1; 1: plain GEMM
};

template <int BN, int STAGES>
__global__ void __launch_bounds__(kGemmThreads, 1) gemm_tf32_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                                                       const GemmParams P) {
    extern __shared__ unsigned char gsm_raw[];
    unsigned char* tiles = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gsm_raw) + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B: 1024-byte aligned
    constexpr int kABytes = kBM * kBK * 4, kBBytes = BN * kBK * 4, kStage = kABytes + kBBytes;
    __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], acc_bar;
    __shared__ uint32_t tmem_base_smem;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile_n = blockIdx.x, tile_m = blockIdx.y, split = blockIdx.z;
    const int kb0 = split * P.kb_per_split;
    const int kb1 = min(kb0 + P.kb_per_split, P.kblocks);
    const int nkb = kb1 - kb0;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Two accumulators.  The tensor core adds every MMA result to the fp32 accumulator with TRUNCATION (measured: the error of one
    // accumulator fed all 3K/8 steps of a 3xTF32 product is (3K/8) * 2^-24 / 2 = 5.8e-6 at K = 512, five times cuBLAS' fp32 SIMT
    // error).  The two cross terms a_hi b_lo + a_lo b_hi are 2^-11 of the main product but cost 2/3 of those steps, so they get an
    // accumulator of their own (its ulp is 2^-11 smaller: no visible error) and the epilogue adds the two in registers (round to nearest).
    const bool use_main = nkb > 0, use_lo = P.kparts == 3 && nkb > 1;  // split ranges start at a multiple of 3 (plan_gemm)
    if (warp == 1) {  // one full warp allocates 2 * BN TMEM columns (power of two >= 32) and gives the permit back
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tmem_base_smem)), "n"(2 * BN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == 0) {
        if (lane == 0) {  // ---- TMA producer
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES, use = i / STAGES;
                if (use > 0) mbar_wait(&empty_bar[s], (uint32_t)((use - 1) & 1));
                mbar_expect_tx(&full_bar[s], (uint32_t)kStage);
                unsigned char* st = tiles + (size_t)s * kStage;
                tma_load_2d(st, &map_a, (kb0 + i) * kBK, tile_m * kBM, &full_bar[s]);
                tma_load_2d(st + kABytes, &map_b, (kb0 + i) * kBK, tile_n * BN, &full_bar[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer
            constexpr uint32_t idesc = umma_idesc_tf32(kBM, BN);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES, use = i / STAGES;
                mbar_wait(&full_bar[s], (uint32_t)(use & 1));
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t a0 = smem_addr(tiles + (size_t)s * kStage), b0 = a0 + kABytes;
                const uint64_t da = umma_desc_k_sw128(a0), db = umma_desc_k_sw128(b0);
                const bool lo = P.kparts == 3 && ((kb0 + i) % 3) != 0;
                const bool first = lo ? (i == 1) : (i == 0);  // kb0 is a multiple of 3: block 0 is main, block 1 the first cross-term block
#pragma unroll
                for (int k = 0; k < kBK / kUK; ++k) {
                    // advance along K inside the 128-byte swizzle atom: + k * 32 bytes on the start address (4 LSB dropped)
                    umma_tf32(tmem_base + (lo ? (uint32_t)BN : 0u), da + (uint64_t)(k * kUK * 4 >> 4), db + (uint64_t)(k * kUK * 4 >> 4), idesc,
                              (first && k == 0) ? 0u : 1u);
                }
                umma_commit(&empty_bar[s]);   // arrives when the MMAs above have finished reading this stage
            }
            umma_commit(&acc_bar);            // accumulator complete
        }
    } else {
        // ---- epilogue warps 2..5: TMEM lanes 32 * (warp % 4) .. + 31 = rows of the tile
        const int q = warp & 3;
        const int row = tile_m * kBM + q * 32 + lane;
        mbar_wait(&acc_bar, 0u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        float* out = P.d + (size_t)split * P.m * P.ldd + (size_t)row * P.ldd + tile_n * BN;
        const bool add_bias = P.bias != nullptr && gridDim.z == 1;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = 0.f;
            if (use_main) tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c, v);
            if (use_lo) {
                float u[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(BN + c), u);
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] += u[i];
            }
            if (row < P.m) {
                const int col0 = tile_n * BN + c;
                if (col0 + 32 <= P.n && (P.ldd & 3) == 0) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        float4 o = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                        if (add_bias) {
                            o.x += __ldg(P.bias + col0 + i); o.y += __ldg(P.bias + col0 + i + 1);
                            o.z += __ldg(P.bias + col0 + i + 2); o.w += __ldg(P.bias + col0 + i + 3);
                        }
                        *reinterpret_cast<float4*>(out + c + i) = o;
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (col0 + i < P.n) out[c + i] = v[i] + (add_bias ? __ldg(P.bias + col0 + i) : 0.f);
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    }
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
    }
}

// sum of the split-K partials in split order (+ bias): deterministic
__global__ void __launch_bounds__(256) gemm_splitk_reduce_kernel(const float* __restrict__ part, int ksplit, int m, int n, int ldd, const float* __restrict__ bias,
                                                                 float* __restrict__ d, int ldo) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (int64_t)m * n) return;
    const int r = (int)(i / n), c = (int)(i - (int64_t)r * n);
    double s = 0.0;
    for (int z = 0; z < ksplit; ++z) s += (double)part[((size_t)z * m + r) * ldd + c];
    d[(size_t)r * ldo + c] = (float)s + (bias ? bias[c] : 0.f);
}

__device__ __forceinline__ float to_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}

// dst = the three parts s0, s1, s2 (s in {hi, lo} chosen by `pattern` bits: bit p set = lo for part p) INTERLEAVED per k-block of 32
// along the reduction dimension: [s0(0:32) | s1(0:32) | s2(0:32) | s0(32:64) | ...], so that every run of 3 k-blocks of the GEMM holds one
// main (hi.hi) block and its two cross-term blocks.
//   transpose = 0: src [rows, cols] -> dst [rows, 3 * kpad]        (reduction over cols; kpad = cols rounded up to 32, padding zero)
//   transpose = 1: src [rows, cols] -> dst [cols, 3 * kpad]        (reduction over rows; kpad = rows rounded up to 32)
__global__ void __launch_bounds__(256) split_tf32_kernel(const float* __restrict__ src, int rows, int cols, int64_t lds, float* __restrict__ dst, int kpad,
                                                         int pattern, int transpose) {
    __shared__ float tile[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    const int r0 = blockIdx.y * 32, c0 = blockIdx.x * 32;
    const int64_t ldd = 3 * (int64_t)kpad;
    if (!transpose) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = r0 + ty + 8 * k, c = c0 + tx;
            if (r >= rows || c >= kpad) continue;
            const float a = c < cols ? src[(size_t)r * lds + c] : 0.f;
            const float hi = to_tf32(a), lo = to_tf32(a - hi);
#pragma unroll
            for (int p = 0; p < 3; ++p) dst[(size_t)r * ldd + (size_t)((c >> 5) * 3 + p) * 32 + (c & 31)] = ((pattern >> p) & 1) ? lo : hi;
        }
    } else {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = r0 + ty + 8 * k, c = c0 + tx;
            tile[ty + 8 * k][tx] = (r < rows && c < cols) ? src[(size_t)r * lds + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = c0 + ty + 8 * k, r = r0 + tx;  // output row = source column c, output column = source row r
            if (c >= cols || r >= kpad) continue;
            const float a = tile[tx][ty + 8 * k];
            const float hi = to_tf32(a), lo = to_tf32(a - hi);
#pragma unroll
            for (int p = 0; p < 3; ++p) dst[(size_t)c * ldd + (size_t)((r >> 5) * 3 + p) * 32 + (r & 31)] = ((pattern >> p) & 1) ? lo : hi;
        }
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn gemm_encode_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// operand [rows, k] row-major (ld elements), box = 32 (k) x box_rows, 128-byte swizzle, out-of-bounds rows / columns read as zero
bool make_operand_map(CUtensorMap* map, const float* ptr, int64_t rows, int64_t k, int64_t ld, int box_rows) {
    EncodeTiledFn fn = gemm_encode_fn();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {(cuuint64_t)k, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    const cuuint32_t estride[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct GemmPlan {
    bool wide;
    int64_t ksplit, kb_per_split, kblocks, ldp;
};
// 128 x 256 tiles when there are enough of them to fill the chip twice over, else 128 x 128; split K when the tiles alone
// leave SMs idle (the weight-gradient GEMMs: M, N <= 512 with K = 3 x loci)
GemmPlan plan_gemm(int64_t m, int64_t n, int64_t k, int kparts) {
    GemmPlan g;
    g.kblocks = (k + kBK - 1) / kBK;
    const int64_t tm = (m + kBM - 1) / kBM, t128 = tm * ((n + 127) / 128);
    g.wide = n > 128 && t128 >= 296;
    const int64_t tiles = g.wide ? tm * ((n + 255) / 256) : t128;
    int64_t ks = 1;
    if (tiles < 148) ks = std::min<int64_t>(std::max<int64_t>(148 / tiles, 1), std::max<int64_t>(g.kblocks / 16, 1));
    g.kb_per_split = (g.kblocks + ks - 1) / ks;
    if (kparts == 3) {
        // The tensor core truncates on every accumulation step: at most 16 main k-blocks (64 MMAs, ~2e-6) per accumulator; longer
        // reductions (the weight-gradient GEMMs: K = 3 x loci) become more split-K partials, summed in f64 by the reduce kernel.
        g.kb_per_split = (g.kb_per_split + 2) / 3 * 3;
        if (g.kb_per_split > 48) g.kb_per_split = 48;
    }
    g.ksplit = (g.kblocks + g.kb_per_split - 1) / g.kb_per_split;
    g.ldp = (n + 3) / 4 * 4;
    return g;
}

template <int BN, int STAGES>
cudaError_t launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, const GemmParams& P, dim3 grid, cudaStream_t stream) {
    constexpr int smem = STAGES * (kBM * kBK * 4 + BN * kBK * 4) + 1024;
    static std::atomic<uint64_t> attr_set{0};
    int dev = 0;
    cudaGetDevice(&dev);
    const uint64_t bit = 1ull << (dev & 63);
    if (dev >= 64 || !(attr_set.load(std::memory_order_relaxed) & bit)) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tf32_tn_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return e;
        attr_set.fetch_or(bit, std::memory_order_relaxed);
    }
    gemm_tf32_tn_kernel<BN, STAGES><<<grid, kGemmThreads, smem, stream>>>(ma, mb, P);
    return cudaGetLastError();
}

}  // namespace
}  // namespace hicgat

using namespace hicgat;

extern "C" int hicgat_split_tf32(const float* src, int64_t rows, int64_t cols, int64_t ld, float* dst, int pattern, int transpose, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(src && dst && rows > 0 && cols > 0 && ld >= cols && rows < (1ll << 30) && cols < (1ll << 30) && pattern >= 0 && pattern < 8,
                   "hicgat_split_tf32: bad arguments");
    const int64_t red = transpose ? rows : cols;
    const int kpad = (int)((red + kBK - 1) / kBK * kBK);
    // grid covers the padded reduction extent so that the padding is written as zero
    const int64_t gr = transpose ? kpad : rows, gc = transpose ? cols : kpad;
    dim3 grid((unsigned)((gc + 31) / 32), (unsigned)((gr + 31) / 32));
    split_tf32_kernel<<<grid, 256, 0, stream>>>(src, (int)rows, (int)cols, ld, dst, kpad, pattern, transpose ? 1 : 0);
    HICGAT_CHECK_LAUNCH("split_tf32_kernel");
    return HICGAT_OK;
}

extern "C" size_t hicgat_gemm_tf32_workspace_bytes(int64_t m, int64_t n, int64_t k, int kparts) {
    if (m <= 0 || n <= 0 || k <= 0) return 0;
    const GemmPlan g = plan_gemm(m, n, k, kparts);
    return g.ksplit > 1 ? (size_t)g.ksplit * m * g.ldp * sizeof(float) : 0;
}

extern "C" int hicgat_gemm_tf32_tn(const float* a, int64_t lda, const float* b, int64_t ldb, int64_t m, int64_t n, int64_t k, int kparts, const float* bias,
                                   float* d, int64_t ldd, void* workspace, size_t workspace_bytes, hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(a && b && d && m > 0 && n > 0 && k > 0 && m < (1ll << 30) && n < (1ll << 30) && k < (1ll << 30), "hicgat_gemm_tf32_tn: bad shape");
    HICGAT_REQUIRE(lda >= k && ldb >= k && ldd >= n && (lda % 4) == 0 && (ldb % 4) == 0 && aligned16(a) && aligned16(b) && aligned16(d),
                   "hicgat_gemm_tf32_tn: operands must be 16-byte aligned with leading dimensions that are multiples of 4");
    HICGAT_REQUIRE((kparts == 1 || kparts == 3) && (k % (kBK * kparts)) == 0, "hicgat_gemm_tf32_tn: k must be a multiple of 32 * kparts, kparts 1 or 3");
    const GemmPlan g = plan_gemm(m, n, k, kparts);
    const int bn = g.wide ? 256 : 128;
    CUtensorMap ma, mb;
    if (!make_operand_map(&ma, a, m, k, lda, kBM) || !make_operand_map(&mb, b, n, k, ldb, bn)) {
        set_error("hicgat_gemm_tf32_tn: cuTensorMapEncodeTiled failed");
        return HICGAT_ERR_CUDA;
    }
    GemmParams P;
    P.bias = bias; P.m = (int)m; P.n = (int)n; P.kblocks = (int)g.kblocks; P.kb_per_split = (int)g.kb_per_split;
    P.kparts = kparts;
    if (g.ksplit > 1) {
        HICGAT_REQUIRE(workspace && workspace_bytes >= (size_t)g.ksplit * m * g.ldp * sizeof(float) && aligned16(workspace), "hicgat_gemm_tf32_tn: workspace too small");
        P.d = static_cast<float*>(workspace);
        P.ldd = (int)g.ldp;
    } else {
        P.d = d;
        P.ldd = (int)ldd;
    }
    dim3 grid((unsigned)((n + bn - 1) / bn), (unsigned)((m + kBM - 1) / kBM), (unsigned)g.ksplit);
    cudaError_t e = g.wide ? launch_gemm<256, 4>(ma, mb, P, grid, stream) : launch_gemm<128, 4>(ma, mb, P, grid, stream);
    if (e != cudaSuccess) {
        set_error("gemm_tf32_tn_kernel: launch failed: %s", cudaGetErrorString(e));
        return HICGAT_ERR_CUDA;
    }
    count_launch();
    if (g.ksplit > 1) {
        const int64_t total = m * n;
        gemm_splitk_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(P.d, (int)g.ksplit, (int)m, (int)n, (int)g.ldp, bias, d, (int)ldd);
        HICGAT_CHECK_LAUNCH("gemm_splitk_reduce_kernel");
    }
    return HICGAT_OK;
}
