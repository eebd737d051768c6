// p2p_allreduce.cu -- gradient all-reduce over NVLink peer memory with CTAs small enough to run BESIDE the ROIAlign grids.
//
// The one collective near this path is the adaptation-training gradient all-reduce of the reference's DDP wrap
// (coin/engine/trainer.py:66-72; BASELINE.json configs[3]). NCCL's kernels never overlap this path's step: the ROIAlign grids
// hold every SM's register file and a NCCL CTA does not fit a retiring ROIAlign slot (tools/ar_overlap.py: the all-reduce adds
// its full stand-alone time). This all-reduce is built from CTAs of 224 threads and <= 40 registers - the size of the slot one
// retiring ROIAlign CTA frees - launched on a high-priority stream, so that it trickles into the machine while the step runs.
//
// Two-shot over peer-mapped buffers (every rank's gradient buffer and flag words are mapped into every rank through CUDA IPC;
// one process per GPU, <= 8 GPUs of one NVSwitch domain):
//   barrier A      every rank's input is complete (stream order) and visible
//   reduce-scatter rank r sums slice r of all ranks' buffers (peers read over NVLink, fixed order 0..n-1) into its own buffer
//   barrier B
//   all-gather     rank r copies the reduced slices of the other ranks into its own buffer
//   barrier C      nobody overwrites a buffer a peer is still reading
// Barriers are single-warp kernels: lane p stores the call's epoch into peer p's flag word for this rank (st.release.sys) and
// spins on its own flag word for peer p (ld.acquire.sys); epochs only grow, so flags are never reset. A spin that exceeds
// ~2^27 polls sets an error word and gives up instead of hanging the device.
#include "common.cuh"

namespace coin {

constexpr int kP2PMax = 8;
constexpr int kP2PThreads = 224;

struct PeerPtrs {
    float* data[kP2PMax];
    uint32_t* flags[kP2PMax];      // [3 barriers][kP2PMax] words per rank
};

__global__ void p2p_barrier_kernel(const PeerPtrs p, int rank, int n, int slot, uint32_t epoch, int32_t* __restrict__ err) {
    const int t = threadIdx.x;
    if (t >= n) return;
    __threadfence_system();
    uint32_t* dst = p.flags[t] + slot * kP2PMax + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(epoch) : "memory");
    const uint32_t* src = p.flags[rank] + slot * kP2PMax + t;
    uint32_t v, spins = 0;
    while (true) {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(src) : "memory");
        if ((int32_t)(v - epoch) >= 0) break;
        if (++spins > (1u << 27)) { atomicExch(err, 1 + slot); break; }
        __nanosleep(64);
    }
}

__global__ void __launch_bounds__(kP2PThreads)
p2p_reduce_scatter_kernel(const PeerPtrs p, float* __restrict__ own_data, int rank, int n, size_t off4, size_t n4_slice) {
    const size_t base = off4 + (size_t)rank * n4_slice;
    float4* __restrict__ own = reinterpret_cast<float4*>(own_data) + base;
    for (size_t i = blockIdx.x * (size_t)kP2PThreads + threadIdx.x; i < n4_slice; i += (size_t)gridDim.x * kP2PThreads) {
        float4 v[kP2PMax];
#pragma unroll
        for (int q = 0; q < kP2PMax; ++q)        // all ranks' loads in flight (peers over NVLink: uncached loads)
            v[q] = q < n ? __ldcv(reinterpret_cast<const float4*>(p.data[q]) + base + i) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 acc = v[0];
#pragma unroll
        for (int q = 1; q < kP2PMax; ++q)        // fixed order 0..n-1: every slice is summed by exactly one rank
            if (q < n) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
        own[i] = acc;
    }
}

__global__ void __launch_bounds__(kP2PThreads)
p2p_all_gather_kernel(const PeerPtrs p, float* __restrict__ own_data, int rank, int n, size_t off4, size_t n4_slice) {
    float4* __restrict__ own = reinterpret_cast<float4*>(own_data) + off4;
#pragma unroll
    for (int q = 0; q < kP2PMax; ++q) {          // the reduced slice of every other rank
        if (q >= n || q == rank) continue;
        const float4* __restrict__ src = reinterpret_cast<const float4*>(p.data[q]) + off4 + (size_t)q * n4_slice;
        float4* __restrict__ dst = own + (size_t)q * n4_slice;
        for (size_t i = blockIdx.x * (size_t)kP2PThreads + threadIdx.x; i < n4_slice; i += (size_t)gridDim.x * kP2PThreads)
            dst[i] = __ldcv(src + i);
    }
}

}  // namespace coin
using namespace coin;

// data_ptrs / flag_ptrs: HOST arrays of n device pointers (entry r = rank r's buffer as mapped into THIS process; entry `rank`
// is the local allocation). The reduced range is [offset, offset + nelem) floats of every data buffer; offset and nelem must be
// multiples of 4 * n. flags: 3 * 8 uint32 per rank, zero-initialised once. epoch: a value that grows with every call on the
// group (1, 2, 3, ...; the same on every rank). err: device int32, set non-zero when a barrier gave up.
extern "C" int coin_p2p_all_reduce(void* const* data_ptrs, void* const* flag_ptrs, int rank, int n, int64_t offset, int64_t nelem,
                                   uint32_t epoch, int32_t* err, int max_ctas, coin_stream_t stream) {
    COIN_REQUIRE(data_ptrs && flag_ptrs && err, "p2p_all_reduce: null pointer");
    COIN_REQUIRE(n >= 1 && n <= kP2PMax && rank >= 0 && rank < n, "p2p_all_reduce: rank %d of %d (max %d)", rank, n, kP2PMax);
    COIN_REQUIRE(offset >= 0 && nelem >= 0 && offset % (4 * n) == 0 && nelem % (4 * n) == 0,
                 "p2p_all_reduce: offset and nelem must be multiples of 4 * n");
    if (nelem == 0 || n == 1) return COIN_OK;
    PeerPtrs p;
    for (int q = 0; q < kP2PMax; ++q) {
        p.data[q] = static_cast<float*>(data_ptrs[q < n ? q : 0]);
        p.flags[q] = static_cast<uint32_t*>(flag_ptrs[q < n ? q : 0]);
        COIN_REQUIRE(p.data[q] && p.flags[q], "p2p_all_reduce: null peer pointer");
    }
    cudaStream_t s = as_stream(stream);
    const size_t n4_slice = (size_t)nelem / 4 / n, off4 = (size_t)offset / 4;
    const unsigned ctas = (unsigned)std::max<int64_t>(1, std::min<int64_t>(ceil_div((int64_t)n4_slice, kP2PThreads * 4),
                                                                          max_ctas > 0 ? max_ctas : 4 * kNumSMs));
    p2p_barrier_kernel<<<1, 32, 0, s>>>(p, rank, n, 0, epoch, err);
    if (int rc = check_launch("p2p_barrier_kernel")) return rc;
    p2p_reduce_scatter_kernel<<<ctas, kP2PThreads, 0, s>>>(p, p.data[rank], rank, n, off4, n4_slice);
    if (int rc = check_launch("p2p_reduce_scatter_kernel")) return rc;
    p2p_barrier_kernel<<<1, 32, 0, s>>>(p, rank, n, 1, epoch, err);
    if (int rc = check_launch("p2p_barrier_kernel")) return rc;
    p2p_all_gather_kernel<<<ctas, kP2PThreads, 0, s>>>(p, p.data[rank], rank, n, off4, n4_slice);
    if (int rc = check_launch("p2p_all_gather_kernel")) return rc;
    p2p_barrier_kernel<<<1, 32, 0, s>>>(p, rank, n, 2, epoch, err);
    return check_launch("p2p_barrier_kernel");
}
