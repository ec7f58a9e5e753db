// Dev experiment (not part of libspsk): validates the cta_group::2 tcgen05 mechanics on a tiny GEMM before they go
// into sa_mma.cu.  D[256 rows, 128 couts] = X[256, 64] . W[128, 64]^T, rows split over the CTA pair (A operand, M = 256),
// couts split over the pair (B operand: each CTA stages N/2 = 64 couts).   nvcc -arch=sm_100a -o pair_test pair_mma_test.cu
#include "../../spsnet_b200/csrc/mma_ptx.cuh"
#include <vector>
#include <cstdlib>
#include <cmath>
using namespace spsk;

namespace spsk { void set_error(const char *, ...) {} int cuda_fail(cudaError_t, const char *) { return -3; } void count_launch() {} }

__device__ __forceinline__ uint32_t cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(128, 1) pair_kernel(const __half *X, const __half *W, float *D, int mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem);        // [0]: acc full
    uint32_t *slot = reinterpret_cast<uint32_t *>(smem + 64);
    uint8_t *xs = smem + 1024;            // 128 rows x 64 k  (16 KB)
    uint8_t *ws = xs + 16384;             // 64 couts x 64 k  (8 KB)
    const int tid = threadIdx.x, warp = tid >> 5;
    const uint32_t rank = cta_rank();
    // stage operands: canonical K-major no-swizzle, SBO = 64*16 = 1024
    {
        const int r = tid;   // row of this CTA
        const __half *src = X + (size_t)(rank * 128 + r) * 64;
        for (int g = 0; g < 8; ++g)
            *reinterpret_cast<uint4 *>(xs + (r >> 3) * 1024 + g * 128 + (r & 7) * 16) = *reinterpret_cast<const uint4 *>(src + g * 8);
        if (tid < 64) {
            const int c = tid;   // cout of this CTA's half
            const __half *wsrc = W + (size_t)(rank * 64 + c) * 64;
            for (int g = 0; g < 8; ++g)
                *reinterpret_cast<uint4 *>(ws + (c >> 3) * 1024 + g * 128 + (c & 7) * 16) = *reinterpret_cast<const uint4 *>(wsrc + g * 8);
        }
    }
    if (tid == 0) { mbar_init(smem_u32(bar), 1); mbar_init_fence(); }
    fence_proxy_async();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    tc_fence_after();
    const uint32_t tmem = *slot;
    if (rank == 0 && warp == 0) {
        if (elect_one()) {
            const uint32_t idesc = umma_idesc(256, 128);
            for (int j = 0; j < 4; ++j) {
                const uint64_t ad = umma_desc(smem_u32(xs) + j * 256, 128, 1024);
                const uint64_t bd = umma_desc(smem_u32(ws) + j * 256, 128, 1024);
                const uint32_t acc = j ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n\t}"
                    ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(acc), "r"(0u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                         ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
        }
        __syncwarp();
    }
    mbar_wait(smem_u32(bar), 0u);
    tc_fence_after();
    {
        const int r = tid;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < 128; c0 += 16) {
            float v[16];
            tmem_ld16(taddr + c0, v);
            for (int i = 0; i < 16; ++i) D[(size_t)(rank * 128 + r) * 128 + c0 + i] = v[i];
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(128));
}

int main() {
    const int M = 256, N = 128, K = 64;
    std::vector<__half> hx(M * K), hw(N * K);
    std::vector<float> fx(M * K), fw(N * K);
    srand(1);
    for (int i = 0; i < M * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hx[i] = __float2half(v); fx[i] = __half2float(hx[i]); }
    for (int i = 0; i < N * K; ++i) { float v = (rand() % 2001 - 1000) / 1000.f; hw[i] = __float2half(v); fw[i] = __half2float(hw[i]); }
    __half *dx, *dw; float *dd;
    cudaMalloc(&dx, M * K * 2); cudaMalloc(&dw, N * K * 2); cudaMalloc(&dd, M * N * 4);
    cudaMemcpy(dx, hx.data(), M * K * 2, cudaMemcpyHostToDevice); cudaMemcpy(dw, hw.data(), N * K * 2, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xFF, M * N * 4);
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(2); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 1024 + 16384 + 8192;
    cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, pair_kernel, (const __half *)dx, (const __half *)dw, dd, 0);
    printf("launch: %s\n", cudaGetErrorString(e));
    e = cudaDeviceSynchronize();
    printf("sync: %s\n", cudaGetErrorString(e));
    std::vector<float> hd(M * N);
    cudaMemcpy(hd.data(), dd, M * N * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0; int bad = 0;
    for (int r = 0; r < M; ++r) for (int c = 0; c < N; ++c) {
        double s = 0; for (int k = 0; k < K; ++k) s += (double)fx[r * K + k] * fw[c * K + k];
        const double d = fabs(s - hd[r * N + c]); if (!(d <= 1e-3)) { if (bad < 5) printf("mismatch r=%d c=%d want %f got %f\n", r, c, s, hd[r * N + c]); ++bad; }
        if (d > maxerr) maxerr = d;
    }
    printf("pair MMA: max err %g, mismatches %d / %d\n", maxerr, bad, M * N);
    return bad != 0;
}
