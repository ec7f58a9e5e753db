// Latency probe (dependent chains, one warp): REDUX, SHFL, ballot, LDS, bar.sync -- cycles per operation on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void probe(unsigned long long *out, int iters) {
    __shared__ uint32_t sm[1024];
    const int lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = (i * 7 + 3) & 1023;
    __syncthreads();
    uint32_t v = lane * 2654435761u;
    long long t0, t1;
    if (threadIdx.x < 32) {
        t0 = clock64();
        for (int i = 0; i < iters; ++i) v = __reduce_max_sync(0xffffffffu, v ^ (uint32_t)i) + lane;
        t1 = clock64();
        if (lane == 0) out[0] = (t1 - t0);
        t0 = clock64();
        for (int i = 0; i < iters; ++i) v = __shfl_xor_sync(0xffffffffu, v, 1) + (uint32_t)i;
        t1 = clock64();
        if (lane == 0) out[1] = (t1 - t0);
        t0 = clock64();
        for (int i = 0; i < iters; ++i) v = __ballot_sync(0xffffffffu, (v + i) & 1u) + lane;
        t1 = clock64();
        if (lane == 0) out[2] = (t1 - t0);
        t0 = clock64();
        for (int i = 0; i < iters; ++i) v = sm[v & 1023];
        t1 = clock64();
        if (lane == 0) out[3] = (t1 - t0);
        t0 = clock64();
        for (int i = 0; i < iters; ++i) v = __reduce_add_sync(0xffffffffu, v & 1u) + i;
        t1 = clock64();
        if (lane == 0) out[5] = (t1 - t0);
    }
    __syncthreads();
    t0 = clock64();
    for (int i = 0; i < iters; ++i) { sm[threadIdx.x & 1023] = v + i; __syncthreads(); v += sm[(threadIdx.x + 32) & 1023]; }
    t1 = clock64();
    if (threadIdx.x == 0) out[4] = (t1 - t0);
    if (v == 0xdeadbeef) out[7] = v;
}
int main() {
    unsigned long long *d, h[8] = {0};
    cudaMalloc(&d, 64); cudaMemset(d, 0, 64);
    const int iters = 2000;
    for (int threads : {512, 1024}) {
        probe<<<1, threads>>>(d, iters); cudaDeviceSynchronize();
        probe<<<1, threads>>>(d, iters); cudaDeviceSynchronize();
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
        printf("threads %d: redux.max %.1f  shfl %.1f  ballot %.1f  lds %.1f  redux.add %.1f  sts+bar+lds %.1f cycles/op\n", threads,
               h[0] / (double)iters, h[1] / (double)iters, h[2] / (double)iters, h[3] / (double)iters, h[5] / (double)iters, h[4] / (double)iters);
    }
    return 0;
}
