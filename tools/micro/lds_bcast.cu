// Micro-benchmark: shared-memory wavefronts of LDS.32/.64/.128 under broadcast patterns (read with ncu:
// l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum / smsp__inst_executed_op_shared_ld.sum).
#include <cstdio>
#include <cuda_runtime.h>

template <int VEC, int MODE>
__global__ void k(float* out, int iters) {
    __shared__ __align__(16) float s[4096];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) s[i] = (float)i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    int cell;
    if (MODE == 0) cell = lane;            // all distinct
    else if (MODE == 1) cell = lane >> 1;  // pairs of lanes share (16 distinct)
    else if (MODE == 2) cell = lane >> 2;  // 8 distinct
    else cell = (lane * 11 / 32) ;         // ~11 distinct, monotone
    const int pitch = VEC == 4 ? 68 : (VEC == 2 ? 66 : 65);
    const float* p = s + cell * pitch;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        const int c = (it * VEC) & 63;
        if (VEC == 1) { float v; asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(p + c))); acc += v; }
        if (VEC == 2) { float2 v; asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((unsigned)__cvta_generic_to_shared(p + c))); acc += v.x + v.y; }
        if (VEC == 4) { float4 v; asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"((unsigned)__cvta_generic_to_shared(p + c))); acc += v.x + v.y + v.z + v.w; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    float* out; cudaMalloc(&out, 1 << 20);
#define RUN(V, M) k<V, M><<<1, 32>>>(out, 1024);
    RUN(1, 0) RUN(1, 1) RUN(1, 2) RUN(1, 3)
    RUN(2, 0) RUN(2, 1) RUN(2, 2) RUN(2, 3)
    RUN(4, 0) RUN(4, 1) RUN(4, 2) RUN(4, 3)
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
