// Dev probe (not part of libspsk): what does ONE issuing thread pay per weight tile in the sa_mma issue loop?
// Patterns per tile of NM tcgen05.mma (128 x 128 x 16, kind::f16, operands in shared memory, no-swizzle K-major):
//   0: NM MMAs                                  (pipe rate)
//   1: NM MMAs + commit
//   2: try_wait(completed barrier) + NM MMAs + commit            (the round-1 loop)
//   3: (NM-1) MMAs + try_wait + 1 MMA + commit                   (acquire the next stage in the shadow of this tile's MMAs)
//   4: as 2 plus tcgen05.fence::after_thread_sync per tile
//   5: as 3 plus the fence after the wait
//   6: try_wait + NM MMAs + 2 commits
//   nvcc -gencode arch=compute_100a,code=sm_100a -o build/issue_loop_probe issue_loop_probe.cu
#include "../../spsnet_b200/csrc/mma_ptx.cuh"
#include <vector>
using namespace spsk;
namespace spsk { void set_error(const char *, ...) {} int cuda_fail(cudaError_t, const char *) { return -3; } void count_launch() {} }

__global__ void __launch_bounds__(128, 1) probe(int n_tiles, int nm, int pattern, long long *out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);   // [0] done-wait target (completed phase 0), [1] commit sink, [2] final
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + 64);
    uint8_t *A = smem + 1024, *B = smem + 1024 + 98304;
    for (int i = threadIdx.x; i < (196608) / 4; i += 128) reinterpret_cast<uint32_t *>(smem + 1024)[i] = 0x3C003C00u;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(bars + 0), 1); mbar_init(smem_u32(bars + 1), 1u << 20); mbar_init(smem_u32(bars + 2), 1);
        mbar_init_fence();
        mbar_arrive(smem_u32(bars + 0));   // phase 0 of bars[0] is complete from here on
    }
    fence_proxy_async();
    if (threadIdx.x < 32) tmem_alloc(smem_u32(slot), 512);
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *slot;
    long long t0 = 0, t1 = 0, t2 = 0;
    if (threadIdx.x < 32) {
        const bool lead = elect_one();
        const uint32_t idesc = umma_idesc(128, 128);
        const uint32_t a_lo = umma_desc_lo(smem_u32(A), 128u), a_hi = umma_desc_hi(1024u);
        const uint32_t b_lo = umma_desc_lo(smem_u32(B), 128u), b_hi = umma_desc_hi(1024u);
        const uint32_t done = smem_u32(bars + 0), sink = smem_u32(bars + 1), fin = smem_u32(bars + 2);
        __syncwarp();
        t0 = clock64();
        for (int t = 0; t < n_tiles; ++t) {
            const uint32_t d = tmem + (uint32_t)((t & 3) * 128);
            if (pattern == 2 || pattern == 4 || pattern == 6) { mbar_wait(done, 0u); if (pattern == 4) tc_fence_after(); }
            const int first = (pattern == 3 || pattern == 5) ? nm - 1 : nm;
            if (lead) for (int j = 0; j < first; ++j) umma_f16_lohi(d, a_lo + 16u * (j & 3), a_hi, b_lo + 16u * (j & 3), b_hi, idesc, 1u);
            if (pattern == 3 || pattern == 5) {
                mbar_wait(done, 0u);
                if (pattern == 5) tc_fence_after();
                if (lead) umma_f16_lohi(d, a_lo + 48u, a_hi, b_lo + 48u, b_hi, idesc, 1u);
            }
            if (lead && pattern >= 1) umma_commit(sink);
            if (lead && pattern == 6) umma_commit(sink);
            __syncwarp();
        }
        if (lead) umma_commit(fin);
        __syncwarp();
        t1 = clock64();
        mbar_wait(fin, 0u);
        t2 = clock64();
        if (lead && blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    tc_fence_before(); __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tmem, 512);
}

int main() {
    long long *d; cudaMalloc(&d, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 1024 + 196608);
    const int n = 500;
    const char *names[] = {"MMAs only", "MMAs + commit", "wait + MMAs + commit (round-1 loop)", "MMAs-1 + wait + MMA + commit (early acquire)",
                           "wait + fence + MMAs + commit", "MMAs-1 + wait + fence + MMA + commit", "wait + MMAs + 2 commits"};
    for (int nm : {4, 8}) for (int p = 0; p < 7; ++p) for (int grid : {1, 148}) {
        probe<<<grid, 128, 1024 + 196608>>>(n, nm, p, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%d MMAs/tile  %-48s grid %3d: issue %.1f cyc/tile (%.1f / MMA), complete %.1f cyc/tile (%s)\n", nm, names[p], grid,
               (double)h[0] / n, (double)h[0] / n / nm, (double)h[1] / n, cudaGetErrorString(e));
    }
    return 0;
}
