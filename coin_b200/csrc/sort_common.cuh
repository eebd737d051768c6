// sort_common.cuh -- the single-CTA bitonic sort and the 64-bit (score, index) key shared by the NMS pipeline (nms.cu) and the
// RPN proposal selection (rpn_proposals.cu).
#pragma once
#include "common.cuh"

namespace coin {

constexpr int kSmallSort = 4096;  // single-CTA bitonic sort up to this many boxes

// descending score, ascending index  ->  ascending 64-bit key
__device__ __forceinline__ uint64_t sort_key(float s, uint32_t idx) {
    s = s + 0.0f;  // -0.0 -> +0.0 so that signed zeros tie like they do on the CPU
    uint32_t u = __float_as_uint(s);
    if (s != s) u = 0x7fc00000u;                       // NaN sorts first, as torch's descending sort
    u ^= (u >> 31) ? 0xffffffffu : 0x80000000u;        // ascending-orderable
    return ((uint64_t)(~u) << 32) | idx;
}

// In-place ascending bitonic sort of npow (a power of two) 64-bit keys in shared memory by a CTA of NT threads.
// Every thread owns a compare-exchange in every stage (pair index t -> elements i = t with a zero inserted at bit
// log2(j), and i | j). Stages with j <= 32 stay inside an aligned 64-key window that the same warp owns in every such
// stage, so they are separated by __syncwarp only; the CTA barrier is paid where a stage crosses windows (j >= 64).
template <int NT>
__device__ __forceinline__ void bitonic_sort_smem(uint64_t* __restrict__ k, int npow) {
    const int half = npow >> 1;
    for (int kk = 2; kk <= npow; kk <<= 1) {
        for (int j = kk >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < half; t += NT) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int p = i | j;
                const uint64_t a = k[i], b = k[p];
                const bool up = (i & kk) == 0;
                if ((a > b) == up) { k[i] = b; k[p] = a; }
            }
            const int next_j = j > 1 ? (j >> 1) : kk;     // first stage of the next merge level has j = kk
            if (j > 32 || next_j > 32) __syncthreads(); else __syncwarp();
        }
    }
    __syncthreads();
}

// 512 threads x 24 registers = 12288 registers: fits the register slot of one retiring ROIAlign CTA (16128) inside the step
constexpr int kSortThreads = 512;
#define COIN_SORT_BOUNDS __maxnreg__(24)

constexpr int kChunk = 4096;      // keys per CTA of the chunked sorts (each chunk sorted on its own, then ranked across chunks)

}  // namespace coin
