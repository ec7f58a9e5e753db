// gather_group.cu -- index gathers (forward) and their scatter-add backwards.
//
// Replaces gather_points_kernel_fast / gather_points_grad_kernel_fast (src/sampling_gpu.cu:8-24,46-63)
// and group_points_kernel_fast / group_points_grad_kernel_fast (src/group_points_gpu.cu:53-72,14-31).
// Pure copies: bit-exact.  HBM/L2-bound; one thread per output element along the contiguous output
// dimension (coalesced writes, idx read once per thread and reused across a channel loop).
#include "common.cuh"

namespace spsk {

constexpr int GG_THREADS = 256;
constexpr int GG_CH_PER_BLOCK = 8;  // channels handled by one CTA (idx reuse)

// out[b,c,j] = points[b,c,idx[b,j]]   with j < cols, cols = npoints (gather) or npoints*nsample (group)
__global__ void __launch_bounds__(GG_THREADS)
gather_cols_kernel(int c, int n, int cols, const float *__restrict__ points, const int *__restrict__ idx,
                   float *__restrict__ out) {
    const int b = blockIdx.z;
    const int j = blockIdx.x * GG_THREADS + threadIdx.x;
    if (j >= cols) return;
    const int src = __ldg(idx + (size_t)b * cols + j);
    const int c0 = blockIdx.y * GG_CH_PER_BLOCK;
    const int c1 = min(c, c0 + GG_CH_PER_BLOCK);
    const float *pb = points + (size_t)b * c * n;
    float *ob = out + (size_t)b * c * cols;
#pragma unroll 4
    for (int ci = c0; ci < c1; ++ci) ob[(size_t)ci * cols + j] = __ldg(pb + (size_t)ci * n + src);
}

// grad_points[b,c,idx[b,j]] += grad_out[b,c,j]
__global__ void __launch_bounds__(GG_THREADS)
scatter_cols_kernel(int c, int n, int cols, const float *__restrict__ grad_out, const int *__restrict__ idx,
                    float *__restrict__ grad_points) {
    const int b = blockIdx.z;
    const int j = blockIdx.x * GG_THREADS + threadIdx.x;
    if (j >= cols) return;
    const int dst = __ldg(idx + (size_t)b * cols + j);
    const int c0 = blockIdx.y * GG_CH_PER_BLOCK;
    const int c1 = min(c, c0 + GG_CH_PER_BLOCK);
    const float *gb = grad_out + (size_t)b * c * cols;
    float *pb = grad_points + (size_t)b * c * n;
    for (int ci = c0; ci < c1; ++ci) atomicAdd(pb + (size_t)ci * n + dst, __ldg(gb + (size_t)ci * cols + j));
}

// out[b,j,:] = in[b,idx[b,j],:]  -- point-major row gather (new_xyz = xyz[sample_idx]); replaces the
// transpose -> gather_operation -> transpose -> contiguous chain of pointnet2_modules.py:261,424.
__global__ void __launch_bounds__(GG_THREADS)
gather_rows_kernel(int n, int m, int c, const float *__restrict__ in, const int *__restrict__ idx, float *__restrict__ out) {
    const int b = blockIdx.y;
    const int e = blockIdx.x * GG_THREADS + threadIdx.x;  // element of the (m, c) output slab
    if (e >= m * c) return;
    const int j = e / c, ch = e - j * c;
    const int src = __ldg(idx + (size_t)b * m + j);
    out[(size_t)b * m * c + e] = __ldg(in + ((size_t)b * n + src) * c + ch);
}

static int launch_cols(bool forward, int b, int c, int n, long long cols, const float *a, const int *idx, float *o,
                       cudaStream_t st, const char *what) {
    SPSK_REQUIRE(b >= 0 && c >= 0 && n >= 0 && cols >= 0, SPSK_ERR_INVALID_ARG, "%s: bad sizes", what);
    SPSK_REQUIRE(cols <= 0x7FFFFFFFLL, SPSK_ERR_UNSUPPORTED, "%s: npoints*nsample overflows int32", what);
    if (b == 0 || c == 0 || cols == 0) return SPSK_OK;
    SPSK_REQUIRE(a && idx && o, SPSK_ERR_INVALID_ARG, "%s: null pointer", what);
    SPSK_REQUIRE(b <= 65535 && (c + GG_CH_PER_BLOCK - 1) / GG_CH_PER_BLOCK <= 65535, SPSK_ERR_UNSUPPORTED,
                 "%s: b or c too large for the launch grid", what);
    dim3 grid((unsigned)((cols + GG_THREADS - 1) / GG_THREADS), (c + GG_CH_PER_BLOCK - 1) / GG_CH_PER_BLOCK, b);
    if (forward) gather_cols_kernel<<<grid, GG_THREADS, 0, st>>>(c, n, (int)cols, a, idx, o);
    else scatter_cols_kernel<<<grid, GG_THREADS, 0, st>>>(c, n, (int)cols, a, idx, o);
    SPSK_LAUNCH_CHECK(what);
    return SPSK_OK;
}

}  // namespace spsk

extern "C" int spsk_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out,
                                  spsk_stream_t stream) {
    return spsk::launch_cols(true, b, c, n, npoints, points, idx, out, spsk::as_stream(stream), "gather_points");
}
extern "C" int spsk_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx,
                                       float *grad_points, spsk_stream_t stream) {
    return spsk::launch_cols(false, b, c, n, npoints, grad_out, idx, grad_points, spsk::as_stream(stream), "gather_points_grad");
}
extern "C" int spsk_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx,
                                 float *out, spsk_stream_t stream) {
    return spsk::launch_cols(true, b, c, n, (long long)npoints * nsample, points, idx, out, spsk::as_stream(stream), "group_points");
}
extern "C" int spsk_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out,
                                      const int *idx, float *grad_points, spsk_stream_t stream) {
    return spsk::launch_cols(false, b, c, n, (long long)npoints * nsample, grad_out, idx, grad_points, spsk::as_stream(stream), "group_points_grad");
}

extern "C" int spsk_gather_rows(int b, int n, int m, int c, const float *in, const int *idx, float *out,
                                spsk_stream_t stream) {
    using namespace spsk;
    SPSK_REQUIRE(b >= 0 && n >= 0 && m >= 0 && c >= 0 && b <= 65535, SPSK_ERR_INVALID_ARG, "gather_rows: bad sizes");
    SPSK_REQUIRE((long long)m * c <= 0x7FFFFFFFLL, SPSK_ERR_UNSUPPORTED, "gather_rows: m*c overflows int32");
    if (b == 0 || m == 0 || c == 0) return SPSK_OK;
    SPSK_REQUIRE(in && idx && out, SPSK_ERR_INVALID_ARG, "gather_rows: null pointer");
    dim3 grid((m * c + GG_THREADS - 1) / GG_THREADS, b);
    gather_rows_kernel<<<grid, GG_THREADS, 0, as_stream(stream)>>>(n, m, c, in, idx, out);
    SPSK_LAUNCH_CHECK("gather_rows_kernel");
    return SPSK_OK;
}
