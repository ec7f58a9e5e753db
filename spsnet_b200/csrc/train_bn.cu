// train_bn.cu -- the small kernels around the batch-statistics passes of spsk_sa_mma_forward (training-mode BatchNorm of the
// grouped shared MLP, reference pointnet2_modules.py:203-211 in train()):
//
//   spsk_sa_pack_layer      conv weight (cout, cin) fp32 [x per-cout BN scale] -> the layer's fp16 weight tiles in the canonical
//                           UMMA layout (include/spsk.h: wtiles), layer-0 row permutation and hi/lo split included.  Training
//                           re-packs every layer every step; one launch per layer instead of a few dozen torch ops.
//   spsk_bn_stats_reduce    per-CTA partial sums of a statistics pass -> (c, 2) fp64 [sum z, sum z^2] (+ the row count in slot
//                           2c, so that ONE all-reduce of 2c + 1 doubles synchronises the statistics across ranks: SyncBatchNorm)
//   spsk_bn_stats_finalize  sums -> mean / biased variance -> folded scale g = gamma / sqrt(var + eps) and bias beta - mean * g
//                           for the next pass, and the running-statistics update torch's BatchNorm does (momentum, unbiased
//                           variance).
// All three are a few microseconds; they exist so that a training step stays launch-light (no host arithmetic, no sync).
#include "common.cuh"
#include <cuda_fp16.h>

namespace spsk {

// one thread per element (kv, c) of the packed matrix W'[vk, cpad]
__global__ void __launch_bounds__(256)
sa_pack_layer_kernel(const float *__restrict__ w, int cout, int cin, const float *__restrict__ scale, int first, int c_feat, int use_xyz,
                     int kpad, int cpad, int split, __half *__restrict__ out) {
    const int vk = split ? 2 * kpad : kpad;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= (long long)vk * cpad) return;
    const int c = (int)(e / vk), kv = (int)(e - (long long)c * vk);
    const int k = kv >= kpad ? kv - kpad : kv;   // split: rows [0, kpad) = Wh, [kpad, 2 kpad) = Wl
    // source column of the reference weight (reference input order of layer 0: [x y z | features], pointnet2_utils.py:315)
    int src = -1;
    if (first) {
        const int xr = use_xyz ? 3 : 0;
        const int xo = split ? (c_feat ? 8 : 0) : (c_feat + 7) / 8 * 8;
        if (k < c_feat) src = xr + k;
        else if (use_xyz && k >= xo && k < xo + 3) src = k - xo;
    } else if (k < cin) {
        src = k;
    }
    float v = 0.f;
    if (src >= 0 && c < cout) {
        v = __ldg(w + (size_t)c * cin + src);
        if (scale) v *= __ldg(scale + c);
    }
    __half h = __float2half_rn(v);
    if (split && kv >= kpad) h = __float2half_rn(v - __half2float(h));
    // position: cout chunk cc (128 wide, the last one narrower), k tile kc (64 wide, the last one narrower), canonical layout
    const int cc = c >> 7, r = c & 127;
    const int ncols = min(128, cpad - cc * 128);
    const int kc = kv >> 6, kk = kv & 63;
    const int kw = min(64, vk - kc * 64);
    const size_t tile = (size_t)128 * cc * vk + (size_t)ncols * 64 * kc;                          // halfs
    const size_t off = (size_t)(r >> 3) * (kw * 8) + (size_t)(kk >> 3) * 64 + (size_t)(r & 7) * 8 + (kk & 7);
    out[tile + off] = h;
}

// block = 32 channels x RED_PY part-lanes: thread (cx, py) sums parts py, py + RED_PY, ... of channel cx (consecutive cx read
// consecutive 16-byte cells: coalesced), then lane py = 0 adds the RED_PY partial sums in a fixed order -- reproducible, and
// nparts / RED_PY dependent loads deep instead of nparts (one thread per channel walking 444 slices took 80-110 us).
constexpr int RED_PY = 16;
__global__ void __launch_bounds__(32 * RED_PY)
bn_stats_reduce_kernel(const double *__restrict__ parts, int nparts, int cpad, int c, double count, double *__restrict__ sums) {
    __shared__ double2 sh[RED_PY][32];
    const int cx = threadIdx.x, py = threadIdx.y;
    const int ch = blockIdx.x * 32 + cx;
    if (ch == 0 && py == 0) sums[2 * c] = count;
    double s = 0.0, q = 0.0;
    if (ch < c) {
#pragma unroll 4
        for (int p = py; p < nparts; p += RED_PY) {
            const double2 v = *reinterpret_cast<const double2 *>(parts + ((size_t)p * cpad + ch) * 2);
            s += v.x;
            q += v.y;
        }
    }
    sh[py][cx] = make_double2(s, q);
    __syncthreads();
    if (py == 0 && ch < c) {
        s = 0.0; q = 0.0;
#pragma unroll
        for (int i = 0; i < RED_PY; ++i) { s += sh[i][cx].x; q += sh[i][cx].y; }
        sums[2 * ch] = s;
        sums[2 * ch + 1] = q;
    }
}

// sums -> (scale, bias) of the BN fold + torch's running-statistics update, for one channel
__device__ __forceinline__ void bn_finalize_channel(int ch, double s, double q, double total, const float *gamma, const float *beta, float eps,
                                                    float momentum, float *running_mean, float *running_var, float *scale, float *bias,
                                                    double *moments) {
    const double mean = s / total;
    double var = q / total - mean * mean;
    var = var > 0.0 ? var : 0.0;
    const double g = (double)__ldg(gamma + ch) / sqrt(var + (double)eps);
    scale[ch] = (float)g;
    bias[ch] = (float)((double)__ldg(beta + ch) - mean * g);
    if (moments) { moments[2 * ch] = mean; moments[2 * ch + 1] = var; }
    if (running_mean && momentum >= 0.f) {   // torch: running = (1 - momentum) * running + momentum * batch, variance unbiased
        const double unbiased = var * (total / (total > 1.0 ? total - 1.0 : 1.0));
        running_mean[ch] = (float)((1.0 - (double)momentum) * (double)running_mean[ch] + (double)momentum * mean);
        running_var[ch] = (float)((1.0 - (double)momentum) * (double)running_var[ch] + (double)momentum * unbiased);
    }
}

// reduce + finalize in one launch (statistics local to this rank: nothing to all-reduce in between); also bumps num_batches_tracked
__global__ void __launch_bounds__(32 * RED_PY)
bn_stats_reduce_finalize_kernel(const double *__restrict__ parts, int nparts, int cpad, int c, double count, const float *__restrict__ gamma,
                                const float *__restrict__ beta, float eps, float momentum, float *__restrict__ running_mean,
                                float *__restrict__ running_var, long long *__restrict__ num_batches_tracked, float *__restrict__ scale,
                                float *__restrict__ bias, double *__restrict__ sums) {
    __shared__ double2 sh[RED_PY][32];
    const int cx = threadIdx.x, py = threadIdx.y;
    const int ch = blockIdx.x * 32 + cx;
    if (ch == 0 && py == 0) {
        if (sums) sums[2 * c] = count;
        if (num_batches_tracked) *num_batches_tracked += 1;
    }
    double s = 0.0, q = 0.0;
    if (ch < c) {
#pragma unroll 4
        for (int p = py; p < nparts; p += RED_PY) {
            const double2 v = *reinterpret_cast<const double2 *>(parts + ((size_t)p * cpad + ch) * 2);
            s += v.x;
            q += v.y;
        }
    }
    sh[py][cx] = make_double2(s, q);
    __syncthreads();
    if (py == 0 && ch < c) {
        s = 0.0; q = 0.0;
#pragma unroll
        for (int i = 0; i < RED_PY; ++i) { s += sh[i][cx].x; q += sh[i][cx].y; }
        if (sums) { sums[2 * ch] = s; sums[2 * ch + 1] = q; }
        bn_finalize_channel(ch, s, q, count, gamma, beta, eps, momentum, running_mean, running_var, scale, bias, nullptr);
    }
}

__global__ void __launch_bounds__(128)
bn_stats_finalize_kernel(const double *__restrict__ sums, int c, const float *__restrict__ gamma, const float *__restrict__ beta, float eps,
                         float momentum, float *__restrict__ running_mean, float *__restrict__ running_var, float *__restrict__ scale,
                         float *__restrict__ bias, double *__restrict__ moments) {
    const int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= c) return;
    bn_finalize_channel(ch, sums[2 * ch], sums[2 * ch + 1], sums[2 * c], gamma, beta, eps, momentum, running_mean, running_var, scale, bias, moments);
}

}  // namespace spsk

extern "C" int spsk_sa_pack_layer(const float *w, int cout, int cin, const float *scale, int first, int c_feat, int use_xyz, int kpad, int cpad,
                                  int split, void *wtiles, spsk_stream_t stream) {
    using namespace spsk;
    SPSK_REQUIRE(w && wtiles, SPSK_ERR_INVALID_ARG, "sa_pack_layer: null pointer");
    SPSK_REQUIRE(cout >= 1 && cin >= 1 && kpad >= 16 && kpad % 16 == 0 && cpad >= cout && cpad % 16 == 0, SPSK_ERR_INVALID_ARG,
                 "sa_pack_layer: bad sizes cout=%d cin=%d kpad=%d cpad=%d", cout, cin, kpad, cpad);
    if (first) {
        const int need = split ? 16 : ((c_feat + 7) / 8 * 8 + (use_xyz ? 8 : 0));
        SPSK_REQUIRE(c_feat >= 0 && cin == c_feat + (use_xyz ? 3 : 0) && kpad >= need && (!split || (c_feat <= 8 && kpad == 16)), SPSK_ERR_INVALID_ARG,
                     "sa_pack_layer: layer 0 takes cin = c_feat + 3*use_xyz = %d inputs (got %d) in kpad >= %d (got %d)", c_feat + (use_xyz ? 3 : 0),
                     cin, need, kpad);
    } else {
        SPSK_REQUIRE(cin <= kpad, SPSK_ERR_INVALID_ARG, "sa_pack_layer: cin=%d > kpad=%d", cin, kpad);
    }
    const long long total = (long long)(split ? 2 * kpad : kpad) * cpad;
    sa_pack_layer_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(w, cout, cin, scale, first ? 1 : 0, c_feat, use_xyz ? 1 : 0, kpad,
                                                                                         cpad, split ? 1 : 0, reinterpret_cast<__half *>(wtiles));
    SPSK_LAUNCH_CHECK("sa_pack_layer_kernel");
    return SPSK_OK;
}

extern "C" int spsk_bn_stats_reduce(const double *parts, int nparts, int cpad, int c, double count, double *sums, spsk_stream_t stream) {
    using namespace spsk;
    SPSK_REQUIRE(parts && sums && nparts >= 1 && c >= 1 && c <= cpad && count >= 1.0, SPSK_ERR_INVALID_ARG, "bn_stats_reduce: bad arguments");
    SPSK_REQUIRE((reinterpret_cast<uintptr_t>(parts) & 15) == 0, SPSK_ERR_INVALID_ARG, "bn_stats_reduce: parts must be 16-byte aligned");
    bn_stats_reduce_kernel<<<(c + 31) / 32, dim3(32, RED_PY), 0, as_stream(stream)>>>(parts, nparts, cpad, c, count, sums);
    SPSK_LAUNCH_CHECK("bn_stats_reduce_kernel");
    return SPSK_OK;
}

extern "C" int spsk_bn_stats_finalize(const double *sums, int c, const float *gamma, const float *beta, float eps, float momentum,
                                      float *running_mean, float *running_var, float *scale, float *bias, double *moments, spsk_stream_t stream) {
    using namespace spsk;
    SPSK_REQUIRE(sums && gamma && beta && scale && bias && c >= 1, SPSK_ERR_INVALID_ARG, "bn_stats_finalize: null pointer");
    SPSK_REQUIRE((running_mean == nullptr) == (running_var == nullptr), SPSK_ERR_INVALID_ARG, "bn_stats_finalize: running_mean / running_var go together");
    bn_stats_finalize_kernel<<<(c + 127) / 128, 128, 0, as_stream(stream)>>>(sums, c, gamma, beta, eps, momentum, running_mean, running_var, scale, bias,
                                                                            moments);
    SPSK_LAUNCH_CHECK("bn_stats_finalize_kernel");
    return SPSK_OK;
}

extern "C" int spsk_bn_stats_reduce_finalize(const double *parts, int nparts, int cpad, int c, double count, const float *gamma, const float *beta,
                                             float eps, float momentum, float *running_mean, float *running_var, long long *num_batches_tracked,
                                             float *scale, float *bias, double *sums, spsk_stream_t stream) {
    using namespace spsk;
    SPSK_REQUIRE(parts && gamma && beta && scale && bias && nparts >= 1 && c >= 1 && c <= cpad && count >= 1.0, SPSK_ERR_INVALID_ARG,
                 "bn_stats_reduce_finalize: bad arguments");
    SPSK_REQUIRE((reinterpret_cast<uintptr_t>(parts) & 15) == 0, SPSK_ERR_INVALID_ARG, "bn_stats_reduce_finalize: parts must be 16-byte aligned");
    SPSK_REQUIRE((running_mean == nullptr) == (running_var == nullptr), SPSK_ERR_INVALID_ARG, "bn_stats_reduce_finalize: running_mean / running_var go together");
    bn_stats_reduce_finalize_kernel<<<(c + 31) / 32, dim3(32, RED_PY), 0, as_stream(stream)>>>(parts, nparts, cpad, c, count, gamma, beta, eps, momentum,
                                                                                              running_mean, running_var, num_batches_tracked, scale, bias, sums);
    SPSK_LAUNCH_CHECK("bn_stats_reduce_finalize_kernel");
    return SPSK_OK;
}
