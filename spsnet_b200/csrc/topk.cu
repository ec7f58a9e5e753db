// topk.cu -- score-based down-sampling (IA-SSD ctr/cls-aware, SPSNet stability-aware) in one launch.
//
// Replaces the torch-op chains of PointnetSAModuleMSG_WithSampling.forward:
//   ctr/cls-aware (pointnet2_modules.py:287-291):  max(dim=-1) -> sigmoid -> topk(npoint) -> .int()
//   'ss'/'sss'   (pointnet2_modules.py:293-303):   sigmoid(max cls) * (1 - sigmoid(stds/8 - 3)) -> topk -> .int()
// (4 resp. ~8 tiny kernels + a radix-select/sort in torch.topk).  One CTA per scene: scores are
// computed with the same fp32 op sequence torch uses (each torch op rounds once: mul by 1/8 is exact,
// sub, sigmoid = 1/(1+expf(-x)) with IEEE add/div, rsub, mul), packed as (score_bits, ~index) 64-bit
// keys and sorted descending by an in-shared-memory bitonic network; the first npoint keys are the
// answer.  Order: descending score, ties by ascending point index (torch.topk leaves tie order
// unspecified; see DESIGN.md "top-k parity").
#include "common.cuh"

namespace spsk {

constexpr int TOPK_THREADS = 1024;

__device__ __forceinline__ float sigmoid_like_torch(float x) {
    // ATen sigmoid (CUDA): one / (one + std::exp(-a)) in fp32
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

__global__ void __launch_bounds__(TOPK_THREADS, 1)
score_topk_kernel(int n, int n_pad, int num_class, int npoint, const float *__restrict__ cls,
                  const float *__restrict__ stds, int *__restrict__ idx, float *__restrict__ scores) {
    extern __shared__ unsigned long long keys[];
    const int b = blockIdx.x;
    const float *cb = cls + (size_t)b * n * num_class;
    const float *sb = stds ? stds + (size_t)b * n : nullptr;
    for (int i = threadIdx.x; i < n_pad; i += TOPK_THREADS) {
        unsigned long long key = 0ull;
        if (i < n) {
            float mx = __ldg(cb + (size_t)i * num_class);
            for (int c = 1; c < num_class; ++c) {   // torch.max propagates NaN (fmaxf would drop it); a NaN score sorts first, as in torch.topk
                const float v = __ldg(cb + (size_t)i * num_class + c);
                mx = (v != v || v > mx) ? v : mx;
            }
            float score = sigmoid_like_torch(mx);
            if (sb) {
                const float t = __fsub_rn(__fmul_rn(__ldg(sb + i), 0.125f), 3.0f);
                const float sta = __fsub_rn(1.0f, sigmoid_like_torch(t));
                score = __fmul_rn(score, sta);
            }
            key = ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
        }
        keys[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= n_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n_pad; i += TOPK_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = keys[i], c = keys[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < c) : (a > c)) { keys[i] = c; keys[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (int j = threadIdx.x; j < npoint; j += TOPK_THREADS) {
        const unsigned long long key = keys[j];
        idx[(size_t)b * npoint + j] = (int)(0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull));
        if (scores) scores[(size_t)b * npoint + j] = __uint_as_float((uint32_t)(key >> 32));
    }
}

}  // namespace spsk

extern "C" int spsk_score_topk(int b, int n, int num_class, int npoint, const float *cls, const float *stds,
                               int *idx, float *scores, spsk_stream_t stream) {
    using namespace spsk;
    SPSK_REQUIRE(b >= 0 && n >= 1 && num_class >= 1 && npoint >= 0, SPSK_ERR_INVALID_ARG,
                 "score_topk: bad sizes b=%d n=%d num_class=%d npoint=%d", b, n, num_class, npoint);
    SPSK_REQUIRE(npoint <= n, SPSK_ERR_INVALID_ARG, "score_topk: npoint=%d > n=%d", npoint, n);
    SPSK_REQUIRE(n <= SPSK_TOPK_MAX_N, SPSK_ERR_UNSUPPORTED, "score_topk: n=%d > %d", n, SPSK_TOPK_MAX_N);
    if (b == 0 || npoint == 0) return SPSK_OK;
    SPSK_REQUIRE(cls && idx, SPSK_ERR_INVALID_ARG, "score_topk: null pointer");
    int n_pad = 2;
    while (n_pad < n) n_pad <<= 1;
    const size_t smem = sizeof(unsigned long long) * (size_t)n_pad;
    if (smem + 2048 > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(score_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(score_topk_kernel)");
    }
    score_topk_kernel<<<b, TOPK_THREADS, smem, as_stream(stream)>>>(n, n_pad, num_class, npoint, cls, stds, idx, scores);
    SPSK_LAUNCH_CHECK("score_topk_kernel");
    return SPSK_OK;
}
