// One-shot all-reduce of the row-sharded loss partials over NVLink peer memory.
//
// The reference has no multi-device code; this is the exchange step of the row-sharded
// pairwise loss.  Every rank's fused loss kernel leaves its partial
// [8 x f64 moments | 3n x f32 gradient contribution] in a SYMMETRIC buffer (the same allocation
// mapped into every rank's address space over NVLink / NVSwitch).  One kernel per rank then
//   1. signals "my partial for epoch e is complete" into every peer's signal pad
//      (release at system scope) and waits until all peers have signalled epoch e;
//   2. reads all `world` partials straight from peer memory (ld.relaxed.sys, 128-bit) and adds
//      them in rank order 0..world-1 (gradient: f32 partials accumulated in f64, rounded once),
//      so every rank computes bit-identical sums;
//   3. writes moments (f64, + an optional constant vector) and the gradient as f32 [n,3].
// There is no trailing barrier: the caller alternates between two packed buffers (epoch parity),
// and a rank can only reach epoch e+1's barrier after finishing epoch e's reads, so a buffer is
// never rewritten (epoch e+2's loss kernel) while a peer still reads it.
// Signal slots hold monotonically increasing epochs, so they are never reset.
#include "common.cuh"

namespace hicgat {
namespace {

constexpr int kMaxWorld = 16;

struct PeerTable {
    const unsigned char* buf[kMaxWorld];  // peer r's partial for this epoch parity: [8 x f64 moments | 3n x f32 gradient]
    uint32_t* pad[kMaxWorld];             // peer r's signal pad (uint32 slots)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_f32x4(const void* p) {
    float4 v;
    asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys_f32(const void* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const void* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) allreduce_partials_kernel(const PeerTable T, int rank, int world, int slot_base, uint32_t epoch,
                                                                 int64_t nfloat, const double* __restrict__ moment_const,
                                                                 double* __restrict__ out_moments, float* __restrict__ out_grad) {
    // ---- 1. cross-GPU barrier on this epoch (block 0 signals; every block polls its own pad)
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();  // the loss kernel's writes (previous launch on this stream) -> visible to peers
        st_release_sys(T.pad[threadIdx.x] + slot_base + rank, epoch);
    }
    if (threadIdx.x < world) {
        const uint32_t* mine = T.pad[rank] + slot_base + threadIdx.x;
        // epochs only grow; the signed distance handles the 2^32 wrap
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) { __nanosleep(20); }
    }
    __syncthreads();

    // ---- 2. moments: f64, rank order
    if (blockIdx.x == 0 && threadIdx.x < HICGAT_PAIR_NMOM) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += ld_relaxed_sys_f64(T.buf[r] + 8 * threadIdx.x);
        out_moments[threadIdx.x] = s + (moment_const ? moment_const[threadIdx.x] : 0.0);
    }
    // ---- 3. gradient: f32 partials, accumulated in f64 in rank order, rounded once
    const int64_t quads = nfloat >> 2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < quads; i += (int64_t)gridDim.x * blockDim.x) {
        double sx = 0.0, sy = 0.0, sz = 0.0, sw = 0.0;
        for (int r = 0; r < world; ++r) {
            const float4 v = ld_relaxed_sys_f32x4(T.buf[r] + 64 + 16 * i);
            sx += (double)v.x; sy += (double)v.y; sz += (double)v.z; sw += (double)v.w;
        }
        reinterpret_cast<float4*>(out_grad)[i] = make_float4((float)sx, (float)sy, (float)sz, (float)sw);
    }
    if (blockIdx.x == 0 && threadIdx.x < (nfloat & 3)) {
        const int64_t e = (quads << 2) + threadIdx.x;
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += (double)ld_relaxed_sys_f32(T.buf[r] + 64 + 4 * e);
        out_grad[e] = (float)s;
    }
}

}  // namespace
}  // namespace hicgat

using namespace hicgat;

extern "C" int hicgat_allreduce_partials_p2p(const uint64_t* peer_bufs_host, const uint64_t* signal_pads_host, int rank, int world,
                                             int64_t n, int64_t buf_offset_bytes, int slot_base, uint32_t epoch,
                                             const double* moment_const, double* out_moments, float* out_grad,
                                             hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(peer_bufs_host && signal_pads_host && out_moments && out_grad, "hicgat_allreduce_partials_p2p: null pointer");
    HICGAT_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "hicgat_allreduce_partials_p2p: bad rank/world (%d/%d)", rank, world);
    HICGAT_REQUIRE(n > 0 && n < (1ll << 30) && buf_offset_bytes >= 0 && (buf_offset_bytes % 16) == 0 && slot_base >= 0, "hicgat_allreduce_partials_p2p: bad n/offset");
    HICGAT_REQUIRE(aligned16(out_grad), "hicgat_allreduce_partials_p2p: out_grad must be 16-byte aligned");
    PeerTable T;
    for (int r = 0; r < world; ++r) {
        HICGAT_REQUIRE(peer_bufs_host[r] && signal_pads_host[r], "hicgat_allreduce_partials_p2p: null peer pointer for rank %d", r);
        HICGAT_REQUIRE((peer_bufs_host[r] % 16) == 0, "hicgat_allreduce_partials_p2p: peer buffer %d not 16-byte aligned", r);
        T.buf[r] = reinterpret_cast<const unsigned char*>(peer_bufs_host[r] + (uint64_t)buf_offset_bytes);
        T.pad[r] = reinterpret_cast<uint32_t*>(signal_pads_host[r]);
    }
    const int64_t nfloat = 3 * n;
    int grid = (int)((nfloat / 4 + 255) / 256);
    if (grid > 148) grid = 148;  // all blocks co-resident: every block polls the barrier
    if (grid < 1) grid = 1;
    allreduce_partials_kernel<<<grid, 256, 0, stream>>>(T, rank, world, slot_base, epoch, nfloat, moment_const, out_moments, out_grad);
    HICGAT_CHECK_LAUNCH("allreduce_partials_kernel");
    return HICGAT_OK;
}
