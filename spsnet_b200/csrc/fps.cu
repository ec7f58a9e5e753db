// fps.cu -- farthest point sampling (D-FPS on xyz, F-FPS on a precomputed distance matrix).
//
// Replaces farthest_point_sampling_kernel / furthest_point_sampling_with_dist_kernel of the
// reference (src/sampling_gpu.cu:93-209, 256-371).  Same arithmetic, same winners (bit-exact,
// including the block-tree tie-break, see common.cuh::fps_rank), different machine mapping:
//
//   reference                                     here
//   -----------------------------------------     ------------------------------------------------
//   running minima `temp` in global memory,       running minima in REGISTERS (P per thread), xyz
//   xyz re-read from global every iteration       staged once into shared memory (SoA) and, for
//                                                 P <= 8, also held in registers
//   11 __syncthreads per iteration (smem tree)    1 __syncthreads per iteration: REDUX.MAX warp
//                                                 arg-max -> one smem slot per warp (double
//                                                 buffered) -> REDUX.MAX again in every warp
//
// FPS is latency bound: m-1 strictly sequential arg-max steps per scene.  One CTA per scene.
#include "common.cuh"

namespace spsk {

// order-preserving float -> uint map (handles negatives; -0 is canonicalised by the caller)
__device__ __forceinline__ uint32_t ordered_bits(float f) {
    uint32_t u = __float_as_uint(f);
    return u ^ ((u >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// Block-wide arg-max of (value, ~rank): returns the winning point index to every thread.
// slots: W uint2 entries for this iteration's parity.  One barrier.
__device__ __forceinline__ uint32_t block_argmax(uint32_t u, uint32_t inv_rank, uint2 *slots, int nwarps,
                                                 uint32_t s_mask, uint32_t s_log2) {
    const uint32_t lane = lane_id();
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t m = __reduce_max_sync(0xFFFFFFFFu, u);
    const uint32_t c = (u == m) ? inv_rank : 0u;
    const uint32_t r = __reduce_max_sync(0xFFFFFFFFu, c);
    if (lane == 0) slots[warp] = make_uint2(m, r);
    __syncthreads();
    uint2 v = make_uint2(0u, 0u);
    if ((int)lane < nwarps) v = slots[lane];
    const uint32_t m2 = __reduce_max_sync(0xFFFFFFFFu, v.x);
    const uint32_t c2 = (v.x == m2) ? v.y : 0u;
    const uint32_t r2 = __reduce_max_sync(0xFFFFFFFFu, c2);
    // invert rank(k) = brev(k & s_mask) | (k >> s_log2)
    const uint32_t rank = ~r2;
    const uint32_t low_mask = s_log2 ? ((1u << (32u - s_log2)) - 1u) : 0xFFFFFFFFu;
    return (__brev(rank) & s_mask) | ((rank & low_mask) << s_log2);
}

// Main kernel, n >= 1024 (reference block size S = 1024 = T).  All points of one thread (k = tid + p*T)
// share k mod S and rank(k) grows with p, so "first strict maximum in p order" is exactly the
// reference's in-thread rule AND its tree tie-break restricted to this thread.  T is a compile-time
// constant and padding points carry tmp = -1 (fminf(d, -1) = -1 never beats best >= -1), so the
// unrolled body is branch-free straight-line code: 3 LDS (or registers) + 10 ALU per point.
template <int P, bool REGXYZ, bool DISTMAT>
__global__ void __launch_bounds__(1024, 1)
fps_kernel_1024(int n, int m, const float *__restrict__ src, float *__restrict__ temp, int *__restrict__ idx) {
    constexpr int T = 1024;
    constexpr uint32_t s_mask = 1023u, s_log2 = 10u;
    extern __shared__ float smem[];
    __shared__ uint2 slots[2][32];
    const int tid = threadIdx.x;
    const size_t scene = blockIdx.x;
    constexpr int NP = P * T;  // padded point count
    float *sx = smem, *sy = smem + NP, *sz = smem + 2 * NP;
    const float *base = DISTMAT ? src + scene * (size_t)n * n : src + scene * (size_t)n * 3;
    if (temp) temp += scene * (size_t)n;
    idx += scene * (size_t)m;

    if (!DISTMAT) {
        for (int i = tid; i < 3 * NP; i += T) smem[i] = 0.f;
        __syncthreads();
        for (int i = tid; i < 3 * n; i += T) {
            const float v = __ldg(base + i);
            const int k = i / 3, c = i - 3 * k;
            smem[c * NP + k] = v;
        }
        __syncthreads();
    }

    float tmp[P];
    float px[REGXYZ ? P : 1], py[REGXYZ ? P : 1], pz[REGXYZ ? P : 1];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int k = tid + p * T;
        tmp[p] = (k < n) ? (temp ? temp[k] : 1e10f) : -1.f;
        if (REGXYZ) { px[p] = sx[k]; py[p] = sy[k]; pz[p] = sz[k]; }
    }

    int old = 0;
    if (tid == 0) idx[0] = 0;

    for (int j = 1; j < m; ++j) {
        float x1 = 0.f, y1 = 0.f, z1 = 0.f;
        const float *drow = nullptr;
        if (DISTMAT) {
            drow = base + (size_t)old * n;
        } else {
            x1 = sx[old]; y1 = sy[old]; z1 = sz[old];
        }
        float best = -1.f;
        int bp = 0;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int k = tid + p * T;
            float d;
            if (DISTMAT) d = (k < n) ? __ldg(drow + k) : 0.f;
            else if (REGXYZ) d = sqdist3(px[p], py[p], pz[p], x1, y1, z1);
            else d = sqdist3(sx[k], sy[k], sz[k], x1, y1, z1);
            const float t = fminf(d, tmp[p]);
            tmp[p] = t;
            const bool gt = t > best;
            best = gt ? t : best;
            bp = gt ? p : bp;
        }
        const uint32_t kb = (uint32_t)(tid + bp * T);
        const uint32_t u = ordered_bits(__fadd_rn(best, 0.f));
        const uint32_t inv_rank = ~fps_rank(kb, s_mask, s_log2);
        old = (int)block_argmax(u, inv_rank, slots[j & 1], T / 32, s_mask, s_log2);
        if (tid == 0) idx[j] = old;
    }

    if (temp) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int k = tid + p * T;
            if (k < n) temp[k] = tmp[p];
        }
    }
}

// Small scenes, n < 1024: the reference block has S = 2^floor(log2 n) < 1024 threads; T = blockDim.x =
// max(S, 32) is a runtime multiple of S and every thread owns at most P = 2 points (same in-thread rule).
template <int P, bool REGXYZ, bool DISTMAT>
__global__ void __launch_bounds__(1024, 1)
fps_kernel(int n, int m, uint32_t s_mask, uint32_t s_log2, const float *__restrict__ src,
           float *__restrict__ temp, int *__restrict__ idx) {
    extern __shared__ float smem[];
    __shared__ uint2 slots[2][32];

    const int T = blockDim.x;
    const int tid = threadIdx.x;
    const int nwarps = T >> 5;
    const size_t scene = blockIdx.x;
    float *sx = smem, *sy = smem + n, *sz = smem + 2 * (size_t)n;
    const float *base = DISTMAT ? src + scene * (size_t)n * n : src + scene * (size_t)n * 3;
    if (temp) temp += scene * (size_t)n;
    idx += scene * (size_t)m;

    if (!DISTMAT) {
        for (int i = tid; i < 3 * n; i += T) {
            const float v = base[i];
            const int k = i / 3, c = i - 3 * k;
            smem[c * n + k] = v;
        }
        __syncthreads();
    }

    float tmp[P];
    float px[REGXYZ ? P : 1], py[REGXYZ ? P : 1], pz[REGXYZ ? P : 1];
#pragma unroll
    for (int p = 0; p < P; ++p) {
        const int k = tid + p * T;
        tmp[p] = (k < n) ? (temp ? temp[k] : 1e10f) : 0.f;
        if (REGXYZ) {
            px[p] = (k < n) ? sx[k] : 0.f;
            py[p] = (k < n) ? sy[k] : 0.f;
            pz[p] = (k < n) ? sz[k] : 0.f;
        }
    }

    int old = 0;
    if (tid == 0) idx[0] = 0;
    const bool has_point = tid < n;

    for (int j = 1; j < m; ++j) {
        float x1 = 0.f, y1 = 0.f, z1 = 0.f;
        const float *drow = nullptr;
        if (DISTMAT) {
            drow = base + (size_t)old * n;
        } else {
            x1 = sx[old]; y1 = sy[old]; z1 = sz[old];
        }
        float best = -1.f;
        int bp = 0;
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int k = tid + p * T;
            if (k < n) {
                float d;
                if (DISTMAT) d = __ldg(drow + k);
                else if (REGXYZ) d = sqdist3(px[p], py[p], pz[p], x1, y1, z1);
                else d = sqdist3(sx[k], sy[k], sz[k], x1, y1, z1);
                const float t = fminf(d, tmp[p]);
                tmp[p] = t;
                if (t > best) { best = t; bp = p; }
            }
        }
        const uint32_t kb = (uint32_t)(tid + bp * T);
        const uint32_t u = has_point ? ordered_bits(__fadd_rn(best, 0.f)) : 0u;
        const uint32_t inv_rank = has_point ? ~fps_rank(kb, s_mask, s_log2) : 0u;
        old = (int)block_argmax(u, inv_rank, slots[j & 1], nwarps, s_mask, s_log2);
        if (tid == 0) idx[j] = old;
    }

    if (temp) {
#pragma unroll
        for (int p = 0; p < P; ++p) {
            const int k = tid + p * T;
            if (k < n) temp[k] = tmp[p];
        }
    }
}

// Any-n fallback: running minima in the caller's `temp` (global/L2), xyz from global.  Same winners.
template <bool DISTMAT>
__global__ void __launch_bounds__(1024, 1)
fps_generic_kernel(int n, int m, uint32_t s_mask, uint32_t s_log2, const float *__restrict__ src,
                   float *__restrict__ temp, int *__restrict__ idx) {
    __shared__ uint2 slots[2][32];
    const int T = blockDim.x;
    const int tid = threadIdx.x;
    const size_t scene = blockIdx.x;
    const float *base = DISTMAT ? src + scene * (size_t)n * n : src + scene * (size_t)n * 3;
    temp += scene * (size_t)n;
    idx += scene * (size_t)m;
    int old = 0;
    if (tid == 0) idx[0] = 0;
    const bool has_point = tid < n;
    for (int j = 1; j < m; ++j) {
        float x1 = 0.f, y1 = 0.f, z1 = 0.f;
        const float *drow = nullptr;
        if (DISTMAT) drow = base + (size_t)old * n;
        else { x1 = __ldg(base + old * 3); y1 = __ldg(base + old * 3 + 1); z1 = __ldg(base + old * 3 + 2); }
        float best = -1.f;
        int bk = tid;
        for (int k = tid; k < n; k += T) {
            float d;
            if (DISTMAT) d = __ldg(drow + k);
            else d = sqdist3(__ldg(base + k * 3), __ldg(base + k * 3 + 1), __ldg(base + k * 3 + 2), x1, y1, z1);
            const float t = fminf(d, temp[k]);
            temp[k] = t;
            if (t > best) { best = t; bk = k; }
        }
        const uint32_t u = has_point ? ordered_bits(__fadd_rn(best, 0.f)) : 0u;
        const uint32_t inv_rank = has_point ? ~fps_rank((uint32_t)bk, s_mask, s_log2) : 0u;
        old = (int)block_argmax(u, inv_rank, slots[j & 1], T >> 5, s_mask, s_log2);
        if (tid == 0) idx[j] = old;
    }
}

template <int P, bool REGXYZ, bool DISTMAT>
static int launch_fps(int b, int n, int m, int threads, uint32_t s_mask, uint32_t s_log2, const float *src,
                      float *temp, int *idx, cudaStream_t st) {
    const size_t smem = DISTMAT ? 0 : sizeof(float) * 3 * (size_t)n;
    auto kern = fps_kernel<P, REGXYZ, DISTMAT>;
    if (smem + 2048 > 48 * 1024) {  // dynamic + the kernel's static smem must stay under the 48 KB default
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(fps_kernel)");
    }
    kern<<<b, threads, smem, st>>>(n, m, s_mask, s_log2, src, temp, idx);
    SPSK_LAUNCH_CHECK("fps_kernel");
    return SPSK_OK;
}

template <int P, bool REGXYZ, bool DISTMAT>
static int launch_fps_1024(int b, int n, int m, const float *src, float *temp, int *idx, cudaStream_t st) {
    const size_t smem = DISTMAT ? 0 : sizeof(float) * 3 * (size_t)P * 1024;
    auto kern = fps_kernel_1024<P, REGXYZ, DISTMAT>;
    if (smem + 2048 > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(fps_kernel_1024)");
    }
    kern<<<b, 1024, smem, st>>>(n, m, src, temp, idx);
    SPSK_LAUNCH_CHECK("fps_kernel_1024");
    return SPSK_OK;
}


template <bool DISTMAT>
static int fps_dispatch(int b, int n, int m, const float *src, float *temp, int *idx, cudaStream_t st) {
    SPSK_REQUIRE(b >= 0 && n >= 1 && m >= 0, SPSK_ERR_INVALID_ARG, "fps: bad sizes b=%d n=%d m=%d", b, n, m);
    SPSK_REQUIRE(src && idx, SPSK_ERR_INVALID_ARG, "fps: null pointer");
    if (b == 0 || m == 0) return SPSK_OK;
    const int S = ref_block_threads(n);
    uint32_t s_log2 = 0;
    while ((1 << s_log2) < S) ++s_log2;
    const uint32_t s_mask = (uint32_t)S - 1u;
    const int threads = S < 32 ? 32 : S;  // multiple of S, >= one warp; 1024 for n >= 1024
    const int per_thread = (n + threads - 1) / threads;
    const bool fits = per_thread <= 16;  // 16 x 1024 padded points of SoA xyz = 192 KB of shared memory
    if (!fits) {
        SPSK_REQUIRE(temp != nullptr, SPSK_ERR_UNSUPPORTED,
                     "fps: n=%d exceeds the on-chip variant (<=16384); pass a (b,n) `temp` scratch filled with 1e10", n);
        fps_generic_kernel<DISTMAT><<<b, threads, 0, st>>>(n, m, s_mask, s_log2, src, temp, idx);
        SPSK_LAUNCH_CHECK("fps_generic_kernel");
        return SPSK_OK;
    }
    if (threads == 1024) {
        if (per_thread <= 1) return launch_fps_1024<1, !DISTMAT, DISTMAT>(b, n, m, src, temp, idx, st);
        if (per_thread <= 2) return launch_fps_1024<2, !DISTMAT, DISTMAT>(b, n, m, src, temp, idx, st);
        if (per_thread <= 4) return launch_fps_1024<4, !DISTMAT, DISTMAT>(b, n, m, src, temp, idx, st);
        if (per_thread <= 8) return launch_fps_1024<8, !DISTMAT, DISTMAT>(b, n, m, src, temp, idx, st);
        return launch_fps_1024<16, false, DISTMAT>(b, n, m, src, temp, idx, st);
    }
    if (per_thread <= 1) return launch_fps<1, !DISTMAT, DISTMAT>(b, n, m, threads, s_mask, s_log2, src, temp, idx, st);
    return launch_fps<2, !DISTMAT, DISTMAT>(b, n, m, threads, s_mask, s_log2, src, temp, idx, st);
}

}  // namespace spsk

extern "C" int spsk_farthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx,
                                            spsk_stream_t stream) {
    return spsk::fps_dispatch<false>(b, n, m, xyz, temp, idx, spsk::as_stream(stream));
}

extern "C" int spsk_furthest_point_sampling_with_dist(int b, int n, int m, const float *dist, float *temp,
                                                      int *idx, spsk_stream_t stream) {
    return spsk::fps_dispatch<true>(b, n, m, dist, temp, idx, spsk::as_stream(stream));
}
