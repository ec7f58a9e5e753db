// Cluster synchronisation latency probe: barrier.cluster arrive/wait vs remote-mbarrier signalling, 4-CTA cluster, 512 threads.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void csync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t s32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __cluster_dims__(4, 1, 1) probe(unsigned long long *out, int iters) {
    __shared__ uint64_t bar[2];
    __shared__ uint32_t rec[2][64][4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t rank = cta_rank();
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&bar[i])), "r"(64));   // 4 CTAs x 16 warps
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    csync();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) csync();
    long long t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    // remote-arrive protocol: every warp writes a record into every CTA's table and arrives on that CTA's mbarrier; everyone waits locally
    csync();
    t0 = clock64();
    uint32_t ph = 0;
    for (int i = 0; i < iters; ++i) {
        const int par = i & 1;
        if (lane < 4) {
            uint32_t ra, rb;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(s32(&rec[par][rank * 16 + warp][0])), "r"(lane));
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(s32(&bar[par])), "r"(lane));
            asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(ra), "r"(i), "r"(warp), "r"(lane), "r"(0) : "memory");
            asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(rb) : "memory");
        }
        __syncwarp();
        uint32_t done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(s32(&bar[par])), "r"((ph >> par) & 1u) : "memory");
        }
        ph ^= (1u << par);
    }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[1] = t1 - t0;
    csync();
    // st.async protocol: the data store itself completes transaction bytes on the consumer's mbarrier (no fence, no cluster barrier)
    __shared__ uint64_t xbar[2];
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&xbar[i])), "r"(1));
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&xbar[i])), "r"(64 * 16) : "memory");
        }
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    csync();
    t0 = clock64();
    ph = 0;
    for (int i = 0; i < iters; ++i) {
        const int par = i & 1;
        if (lane < 4) {
            uint32_t ra, rb;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(s32(&rec[par][rank * 16 + warp][0])), "r"(lane));
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(rb) : "r"(s32(&xbar[par])), "r"(lane));
            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
                         ::"r"(ra), "r"(i), "r"(warp), "r"(lane), "r"(0), "r"(rb) : "memory");
        }
        __syncwarp();
        uint32_t done = 0;
        while (!done) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                         : "=r"(done) : "r"(s32(&xbar[par])), "r"((ph >> par) & 1u) : "memory");
        }
        ph ^= (1u << par);
        volatile uint32_t *rr = &rec[par][lane][0];
        uint32_t chk = rr[0];
        if (chk != (uint32_t)i) out[3] = 1;       // stale record seen
        __syncthreads();                            // every warp of this CTA has read the records of this phase
        if (tid == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&xbar[par])), "r"(64 * 16) : "memory");
    }
    t1 = clock64();
    if (tid == 0 && blockIdx.x == 0) out[2] = t1 - t0;
    csync();
}
int main() {
    unsigned long long *d, h[4];
    cudaMalloc(&d, 32); cudaMemset(d, 0, 32);
    const int iters = 2000;
    for (int rep = 0; rep < 2; ++rep) { probe<<<4, 512>>>(d, iters); cudaError_t e = cudaDeviceSynchronize(); if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; } }
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("cluster(4) x 512 threads: barrier.cluster arrive+wait %.0f cycles;  remote store + remote mbarrier arrive + local wait %.0f cycles;"
           "  st.async complete_tx + local wait + __syncthreads re-arm %.0f cycles (stale=%llu)\n",
           h[0] / (double)iters, h[1] / (double)iters, h[2] / (double)iters, h[3]);
    return 0;
}
