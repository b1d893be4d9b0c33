// One-shot all-reduce of the row-sharded loss partials over NVLink peer memory.
//
// The reference has no multi-device code; this is the exchange step of the row-sharded
// pairwise loss.  Every rank's fused loss kernel leaves
// packed f64[8 + 3n] = [moments | gradient contribution] in a SYMMETRIC buffer (the same
// allocation mapped into every rank's address space over NVLink / NVSwitch).  One kernel per
// rank then
//   1. signals "my partial for epoch e is complete" into every peer's signal pad
//      (release at system scope) and waits until all peers have signalled epoch e;
//   2. reads all `world` partials straight from peer memory (ld.relaxed.sys, 128-bit) and adds
//      them in rank order 0..world-1, so every rank computes bit-identical sums;
//   3. writes moments (f64, + an optional constant vector) and the gradient as f32 [n,3]:
//      the unpack/convert of the NCCL path is fused in.
// There is no trailing barrier: the caller alternates between two packed buffers (epoch parity),
// and a rank can only reach epoch e+1's barrier after finishing epoch e's reads, so a buffer is
// never rewritten (epoch e+2's loss kernel) while a peer still reads it.
// Signal slots hold monotonically increasing epochs, so they are never reset.
#include "common.cuh"

namespace hicgat {
namespace {

constexpr int kMaxWorld = 16;

struct PeerTable {
    const double* buf[kMaxWorld];   // peer r's packed buffer for this epoch parity
    uint32_t* pad[kMaxWorld];       // peer r's signal pad (uint32 slots)
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double2 ld_relaxed_sys_f64x2(const double* p) {
    double2 v;
    asm volatile("ld.relaxed.sys.global.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double* p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) allreduce_packed_kernel(const PeerTable T, int rank, int world, int slot_base, uint32_t epoch,
                                                               int64_t count, const double* __restrict__ moment_const,
                                                               double* __restrict__ out_moments, float* __restrict__ out_grad) {
    // ---- 1. cross-GPU barrier on this epoch (block 0 signals; every block polls its own copy)
    if (blockIdx.x == 0 && threadIdx.x < world) {
        __threadfence_system();  // the loss kernel's writes (previous launch on this stream) -> visible to peers
        st_release_sys(T.pad[threadIdx.x] + slot_base + rank, epoch);
    }
    if (threadIdx.x < world) {
        const uint32_t* mine = T.pad[rank] + slot_base + threadIdx.x;
        // epochs only grow; signed distance handles the 2^32 wrap
        while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) { __nanosleep(20); }
    }
    __syncthreads();

    // ---- 2./3. rank-ordered sum of the peers' partials, two f64 per thread per step
    const int64_t pairs = count >> 1;  // count = 8 + 3n: handle the odd tail separately
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < pairs; i += (int64_t)gridDim.x * blockDim.x) {
        double2 s = make_double2(0.0, 0.0);
        for (int r = 0; r < world; ++r) {
            const double2 v = ld_relaxed_sys_f64x2(T.buf[r] + 2 * i);
            s.x += v.x;
            s.y += v.y;
        }
        const int64_t e0 = 2 * i;
        if (e0 < HICGAT_PAIR_NMOM) {
            out_moments[e0] = s.x + (moment_const ? moment_const[e0] : 0.0);
            out_moments[e0 + 1] = s.y + (moment_const ? moment_const[e0 + 1] : 0.0);
        } else {
            out_grad[e0 - HICGAT_PAIR_NMOM] = (float)s.x;
            out_grad[e0 + 1 - HICGAT_PAIR_NMOM] = (float)s.y;
        }
    }
    if ((count & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const int64_t e0 = count - 1;
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += ld_relaxed_sys_f64(T.buf[r] + e0);
        out_grad[e0 - HICGAT_PAIR_NMOM] = (float)s;
    }
}

}  // namespace
}  // namespace hicgat

using namespace hicgat;

extern "C" int hicgat_allreduce_packed_p2p(const uint64_t* peer_bufs_host, const uint64_t* signal_pads_host, int rank, int world,
                                           int64_t n, int64_t buf_offset_bytes, int slot_base, uint32_t epoch,
                                           const double* moment_const, double* out_moments, float* out_grad,
                                           hicgat_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    HICGAT_REQUIRE(peer_bufs_host && signal_pads_host && out_moments && out_grad, "hicgat_allreduce_packed_p2p: null pointer");
    HICGAT_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "hicgat_allreduce_packed_p2p: bad rank/world (%d/%d)", rank, world);
    HICGAT_REQUIRE(n > 0 && n < (1ll << 30) && buf_offset_bytes >= 0 && (buf_offset_bytes % 16) == 0 && slot_base >= 0, "hicgat_allreduce_packed_p2p: bad n/offset");
    PeerTable T;
    for (int r = 0; r < world; ++r) {
        HICGAT_REQUIRE(peer_bufs_host[r] && signal_pads_host[r], "hicgat_allreduce_packed_p2p: null peer pointer for rank %d", r);
        HICGAT_REQUIRE((peer_bufs_host[r] % 16) == 0, "hicgat_allreduce_packed_p2p: peer buffer %d not 16-byte aligned", r);
        T.buf[r] = reinterpret_cast<const double*>(peer_bufs_host[r] + (uint64_t)buf_offset_bytes);
        T.pad[r] = reinterpret_cast<uint32_t*>(signal_pads_host[r]);
    }
    const int64_t count = HICGAT_PAIR_NMOM + 3 * n;
    int grid = (int)((count / 2 + 255) / 256);
    if (grid > 148) grid = 148;  // all blocks co-resident: every block polls the barrier
    if (grid < 1) grid = 1;
    allreduce_packed_kernel<<<grid, 256, 0, stream>>>(T, rank, world, slot_base, epoch, count, moment_const, out_moments, out_grad);
    HICGAT_CHECK_LAUNCH("allreduce_packed_kernel");
    return HICGAT_OK;
}
