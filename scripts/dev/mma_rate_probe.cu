// Dev probe (not part of libspsk): issue rate / throughput of tcgen05.mma kind::f16, cta_group::1, both operands in shared
// memory in the canonical K-major NO-swizzle layout used by sa_mma.cu, for several shapes and row-group strides (SBO).
//   nvcc -arch=sm_100a -o mma_rate_probe mma_rate_probe.cu
#include "../../spsnet_b200/csrc/mma_ptx.cuh"
#include <vector>
using namespace spsk;
namespace spsk { void set_error(const char *, ...) {} int cuda_fail(cudaError_t, const char *) { return -3; } void count_launch() {} }

__global__ void __launch_bounds__(128, 1) probe(int n_mma, int N, uint32_t sbo_a, uint32_t sbo_b, int kblocks, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + 64);
    uint8_t *A = smem + 1024, *B = smem + 1024 + 98304;
    for (int i = threadIdx.x; i < (196608) / 4; i += 128) reinterpret_cast<uint32_t *>(smem + 1024)[i] = 0x3C003C00u;   // fp16 1.0
    if (threadIdx.x == 0) { mbar_init(smem_u32(bar), 1); mbar_init_fence(); }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc(smem_u32(slot), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *slot;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (threadIdx.x < 32) {
        const bool lead = elect_one();
        const uint32_t idesc = umma_idesc(128, N);
        const uint32_t a_lo = umma_desc_lo(smem_u32(A), 128u), a_hi = umma_desc_hi(sbo_a);
        const uint32_t b_lo = umma_desc_lo(smem_u32(B), 128u), b_hi = umma_desc_hi(sbo_b);
        __syncwarp();
        t0 = clock64();
        if (lead) {
            for (int i = 0; i < n_mma; ++i) {
                const uint32_t kb = (uint32_t)(i % kblocks) * 16u;
                umma_f16_lohi(tmem + (uint32_t)((i & 1) * 256), a_lo + kb, a_hi, b_lo + kb, b_hi, idesc, 1u);
            }
            umma_commit(smem_u32(bar));
        }
        __syncwarp();
        t1 = clock64();
        mbar_wait(smem_u32(bar), 0u);
        t2 = clock64();
        if (lead && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
    long long *d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 196608);
    const int n = 2000;
    struct Cfg { int N; uint32_t sa, sb; int kb; const char *what; } cfgs[] = {
        {128, 1024, 1024, 4, "N=128 SBO 1024/1024 (64-wide K tiles)"},
        {128, 8192, 1024, 4, "N=128 A SBO 8192 (512-wide activations), B SBO 1024"},
        {128, 1024, 4096, 4, "N=128 A SBO 1024, B SBO 4096 (last layer: B = activations)"},
        {256, 1024, 1024, 4, "N=256 SBO 1024/1024"},
        {256, 4096, 1024, 4, "N=256 A SBO 4096, B 1024"},
        {64, 1024, 1024, 4, "N=64"},
        {32, 1024, 1024, 4, "N=32"},
    };
    for (auto &c : cfgs) {
        for (int grid : {1, 148}) {
            probe<<<grid, 128, 1024 + 196608>>>(n, c.N, c.sa, c.sb, c.kb, d);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
            printf("%-62s grid %3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA (%s)\n", c.what, grid, (double)h[0] / n, (double)h[1] / n, cudaGetErrorString(e));
        }
    }
    return 0;
}
