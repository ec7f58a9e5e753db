// scatter_grad.cu -- atomic-free, bit-reproducible backward of gather / group / three_interpolate.
//
// The reference's backward kernels scatter with atomicAdd (src/sampling_gpu.cu:46-63, src/group_points_gpu.cu:14-31,
// src/interpolate_gpu.cu:127-149): the order in which the contributions of one source point arrive is arbitrary, so two
// runs of the same training step differ in the last bits.  All three are the same operation,
//     grad_points[b, c, idx[b, l]] += w[b, l] * grad_out[b, c, l / div]        l = 0 .. L-1
// (gather: L = npoints, div = 1, w = 1;  group: L = npoint * nsample, div = 1, w = 1;  interpolate: L = 3 n, div = 3,
// w = weight), done here as a sorted segment reduction:
//   1. keys (scene * N + target) -> a STABLE radix sort (cub::DeviceRadixSort, only the bits that can be set) carrying the
//      entry number l: every target's contributions become one contiguous run, in ascending l;
//   2. one thread per (scene, target) finds where its run starts (binary search, once, shared by all channels);
//   3. one thread per (scene, target, channel) sums its run in that fixed order -- a plain store per output element, no
//      atomics, no pre-zeroed output needed.
// Same values as the reference up to fp32 summation order (the reference's own order is not defined).
#include <algorithm>

#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace spsk {

__global__ void sg_keys_kernel(int b, int n, long long l_per_scene, const int *__restrict__ idx, unsigned int *__restrict__ keys,
                               unsigned int *__restrict__ vals) {
    const long long total = (long long)b * l_per_scene;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long scene = i / l_per_scene;
        int t = __ldg(idx + i);
        t = min(max(t, 0), n - 1);   // the reference would write out of bounds; clamp instead
        keys[i] = (unsigned int)(scene * n + t);
        vals[i] = (unsigned int)(i - scene * l_per_scene);
    }
}

__device__ __forceinline__ long long sg_lower_bound(const unsigned int *__restrict__ keys, long long lo, long long hi, unsigned int key) {
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(keys + mid) < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// start[scene * n + t] = first sorted position whose key is >= scene * n + t   (start[b * n] = b * l_per_scene)
__global__ void sg_bounds_kernel(long long bn, int n, long long l_per_scene, const unsigned int *__restrict__ keys,
                                 unsigned int *__restrict__ start) {
    for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k <= bn; k += (long long)gridDim.x * blockDim.x) {
        if (k == bn) { start[k] = (unsigned int)((bn / n) * l_per_scene); continue; }
        const long long scene = k / n;
        start[k] = (unsigned int)sg_lower_bound(keys, scene * l_per_scene, (scene + 1) * l_per_scene, (unsigned int)k);
    }
}

// grid (ceil(n / 128), c, b): thread = one target of one channel of one scene
__global__ void __launch_bounds__(128)
sg_reduce_kernel(int n, int c, long long l_per_scene, int cols, int div, const unsigned int *__restrict__ start,
                 const unsigned int *__restrict__ vals, const float *__restrict__ weight, const float *__restrict__ grad_out,
                 float *__restrict__ grad_points) {
    const int t = blockIdx.x * 128 + threadIdx.x;
    if (t >= n) return;
    const int ch = blockIdx.y, scene = blockIdx.z;
    const size_t k = (size_t)scene * n + t;
    const unsigned int lo = __ldg(start + k);
    // the next target's start, except across a scene boundary (keys of the next scene start at its own offset)
    const unsigned int hi = (t + 1 < n) ? __ldg(start + k + 1) : (unsigned int)((long long)(scene + 1) * l_per_scene);
    const float *g = grad_out + ((size_t)scene * c + ch) * cols;
    const float *w = weight ? weight + (size_t)scene * l_per_scene : nullptr;
    float acc = 0.0f;
    for (unsigned int p = lo; p < hi; ++p) {
        const unsigned int l = __ldg(vals + p);
        const float v = __ldg(g + l / (unsigned int)div);
        acc = w ? __fmaf_rn(v, __ldg(w + l), acc) : __fadd_rn(acc, v);
    }
    grad_points[((size_t)scene * c + ch) * n + t] = acc;
}

static size_t sg_sort_temp_bytes(long long total, int end_bit) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned int *)nullptr, (unsigned int *)nullptr,
                                    (const unsigned int *)nullptr, (unsigned int *)nullptr, (int)total, 0, end_bit);
    return bytes;
}

static int sg_end_bit(long long bn) {
    int bits = 1;
    while ((1ll << bits) < bn && bits < 32) ++bits;
    return bits;
}

static size_t sg_align(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace spsk

using namespace spsk;

extern "C" {

SPSK_API long long spsk_scatter_grad_workspace_bytes(int b, int n, long long l_per_scene) {
    if (b <= 0 || n <= 0 || l_per_scene <= 0) return 0;
    const long long total = (long long)b * l_per_scene;
    if (total > 0x7fffffffLL || (long long)b * n > 0xffffffffLL) return -1;
    return (long long)(4 * sg_align((size_t)total * 4) + sg_align(((size_t)b * n + 1) * 4) +
                       sg_align(sg_sort_temp_bytes(total, sg_end_bit((long long)b * n))) + 256);
}

/* grad_points[b, ch, idx[b, l]] += weight[b, l] * grad_out[b, ch, l / div]; grad_out (b, c, cols), idx (b, l_per_scene),
 * weight (b, l_per_scene) or NULL, grad_points (b, c, n) fully written. */
SPSK_API int spsk_scatter_grad(int b, int c, int n, long long l_per_scene, int cols, int div, const float *grad_out,
                               const int *idx, const float *weight, float *grad_points, void *workspace,
                               long long workspace_bytes, spsk_stream_t stream) {
    SPSK_REQUIRE(b >= 0 && c >= 0 && n >= 1 && l_per_scene >= 0 && div >= 1 && cols >= 0, SPSK_ERR_INVALID_ARG,
                 "spsk_scatter_grad: bad sizes b=%d c=%d n=%d l=%lld cols=%d div=%d", b, c, n, l_per_scene, cols, div);
    if (b == 0 || c == 0) return SPSK_OK;
    SPSK_REQUIRE(grad_points, SPSK_ERR_INVALID_ARG, "spsk_scatter_grad: null grad_points");
    cudaStream_t st = as_stream(stream);
    if (l_per_scene == 0) {
        cudaError_t e = cudaMemsetAsync(grad_points, 0, sizeof(float) * (size_t)b * c * n, st);
        if (e != cudaSuccess) return cuda_fail(e, "spsk_scatter_grad memset");
        return SPSK_OK;
    }
    SPSK_REQUIRE(grad_out && idx, SPSK_ERR_INVALID_ARG, "spsk_scatter_grad: null pointer");
    SPSK_REQUIRE((l_per_scene + div - 1) / div <= cols, SPSK_ERR_INVALID_ARG, "spsk_scatter_grad: grad_out has %d columns, entries need %lld",
                 cols, (l_per_scene + div - 1) / div);
    SPSK_REQUIRE(c <= 65535 && b <= 65535, SPSK_ERR_UNSUPPORTED, "spsk_scatter_grad: c=%d / b=%d exceed the grid limits", c, b);
    const long long need = spsk_scatter_grad_workspace_bytes(b, n, l_per_scene);
    SPSK_REQUIRE(need >= 0, SPSK_ERR_UNSUPPORTED, "spsk_scatter_grad: problem too large for 32-bit keys");
    SPSK_REQUIRE(workspace && workspace_bytes >= need, SPSK_ERR_WORKSPACE, "spsk_scatter_grad: workspace %lld < %lld bytes",
                 workspace_bytes, need);
    const long long total = (long long)b * l_per_scene;
    char *ws = reinterpret_cast<char *>(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    const size_t seg = sg_align((size_t)total * 4);
    unsigned int *keys_in = reinterpret_cast<unsigned int *>(ws), *vals_in = reinterpret_cast<unsigned int *>(ws + seg);
    unsigned int *keys_out = reinterpret_cast<unsigned int *>(ws + 2 * seg), *vals_out = reinterpret_cast<unsigned int *>(ws + 3 * seg);
    unsigned int *start = reinterpret_cast<unsigned int *>(ws + 4 * seg);
    void *tmp = ws + 4 * seg + sg_align(((size_t)b * n + 1) * 4);
    const int end_bit = sg_end_bit((long long)b * n);
    size_t tmp_bytes = sg_sort_temp_bytes(total, end_bit);
    const int blocks = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    sg_keys_kernel<<<blocks, 256, 0, st>>>(b, n, l_per_scene, idx, keys_in, vals_in);
    SPSK_LAUNCH_CHECK("sg_keys_kernel");
    cudaError_t e = cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_in, keys_out, vals_in, vals_out, (int)total, 0, end_bit, st);
    if (e != cudaSuccess) return cuda_fail(e, "cub::DeviceRadixSort::SortPairs");
    count_launch();
    const long long bn = (long long)b * n;
    sg_bounds_kernel<<<(int)std::min<long long>((bn + 256) / 256, 148 * 16), 256, 0, st>>>(bn, n, l_per_scene, keys_out, start);
    SPSK_LAUNCH_CHECK("sg_bounds_kernel");
    dim3 grid((n + 127) / 128, c, b);
    sg_reduce_kernel<<<grid, 128, 0, st>>>(n, c, l_per_scene, cols, div, start, vals_out, weight, grad_out, grad_points);
    SPSK_LAUNCH_CHECK("sg_reduce_kernel");
    return SPSK_OK;
}

}  // extern "C"
