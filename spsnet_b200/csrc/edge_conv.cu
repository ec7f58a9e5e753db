// edge_conv.cu -- SPSNet's "surface feature" extractor (DenseEdgeConv, USE_SURFACE: True in SPSNet.yaml:48) as two
// fp32 kernels per convolution unit.
//
// Replaces, per unit of FeatureExtraction (pcdet/ops/pointnet2/pointnet2_batch/surface_feature.py:118-187):
//     x = transforms[i](x)                                   FCLayer                     (:7-26)
//     y = convs[i](x, pos)                                   DenseEdgeConv.forward        (:45-115)
// which in the reference materialises (B, N, K, 72) / (B, N, K, 36) / (B, N, K, 48) / (B, N, K, 60) tensors
// (K = 16 neighbours; 1.2 GB for the first one at B = 16, N = 16384) through ~15 torch kernels.
//
// Algebra.  Every row of the grouped tensor is [x_i, x_j, x_j - x_i] (or x_j - x_i for the first unit) with x_i the
// centre and x_j a neighbour, and the dense connections only ever append x_i again, so with the first layer
// W1 = [W1a | W1b | W1c], the middle layer W2 = [W2a | W2b] and the last layer W3 = [W3a | W3b | W3c]:
//     l1 = relu(P_i + Q_j)                   P = (W1a - W1c) x + b1,  Q = (W1b + W1c) x      (first unit: P = b1 - W1 x, Q = W1 x)
//     l2 = relu(W2a l1 + R2_i)               R2 = W2b x + b2
//     l3 =      W3a l2 + W3b l1 + R3_i       R3 = W3c x + b3
//     out_i = [max_j l3, max_j l2, max_j l1, x_i]          (max over the K neighbours; x_i is constant over j)
// P, Q, R2, R3 are per-POINT (kernel 1, together with the transform FC), only two 12 x 12 and one 24 x 12 products
// remain per (point, neighbour) ROW (kernel 2): 432 instead of 1872 multiply-adds per row, nothing materialised.
// Same fp32 arithmetic up to summation order (tests: <= 1e-5 of range against the reference modules with the
// neighbour lists teacher-forced, <= 1e-3 end to end).
//
// Weights travel BY VALUE in the launch parameters (__grid_constant__ structs): every FFMA reads its weight operand
// straight from the constant bank, all threads of a warp the same address.  HBM traffic per unit: x in, t (24) + u (48)
// + out (60) floats per point, 16 index reads and 16 gathered 48-byte Q rows per point (L2 resident: u is 12.6 MB at
// B = 16, N = 16384).  Bound: fp32 issue (7.5 k FFMA per point), not HBM.
#include "common.cuh"

namespace spsk {

constexpr int EC_C = SPSK_EDGE_CH;    // 24: width of the transformed features
constexpr int EC_G = SPSK_EDGE_GROW;  // 12: growth rate

template <int CIN>
__global__ void __launch_bounds__(128)
edge_point_kernel(const __grid_constant__ spsk_edge_point_weights w, int rows, const float *__restrict__ x, int ldx,
                  float *__restrict__ t, float *__restrict__ u) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const int cin = CIN > 0 ? CIN : w.cin;
    float tv[EC_C];
#pragma unroll
    for (int o = 0; o < EC_C; ++o) tv[o] = w.bt[o];
    const float *xr = x + (size_t)r * ldx;
    if (CIN > 0) {
#pragma unroll
        for (int k = 0; k < (CIN > 0 ? CIN : 1); ++k) {
            const float xv = __ldg(xr + k);
#pragma unroll
            for (int o = 0; o < EC_C; ++o) tv[o] = fmaf(w.wt[k * EC_C + o], xv, tv[o]);
        }
    } else {
        for (int k = 0; k < cin; ++k) {
            const float xv = __ldg(xr + k);
#pragma unroll
            for (int o = 0; o < EC_C; ++o) tv[o] = fmaf(w.wt[k * EC_C + o], xv, tv[o]);
        }
    }
    if (w.relu) {
#pragma unroll
        for (int o = 0; o < EC_C; ++o) tv[o] = fmaxf(tv[o], 0.0f);
    }
    float4 *tp = reinterpret_cast<float4 *>(t + (size_t)r * EC_C);
#pragma unroll
    for (int o = 0; o < EC_C / 4; ++o) tp[o] = make_float4(tv[4 * o], tv[4 * o + 1], tv[4 * o + 2], tv[4 * o + 3]);
    float4 *up = reinterpret_cast<float4 *>(u + (size_t)r * 4 * EC_G);
#pragma unroll
    for (int q = 0; q < 4 * EC_G / 4; ++q) {
        float acc[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int o = 4 * q + e;
            float a = w.c[o];
#pragma unroll
            for (int k = 0; k < EC_C; ++k) a = fmaf(w.m[k * 4 * EC_G + o], tv[k], a);
            acc[e] = a;
        }
        up[q] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
}

__global__ void __launch_bounds__(128)
edge_aggr_kernel(const __grid_constant__ spsk_edge_aggr_weights w, int b, int n, int K, const int *__restrict__ idx,
                 const float *__restrict__ t, const float *__restrict__ u, float *__restrict__ out, int ldo) {
    const long long r = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= (long long)b * n) return;
    const long long base = (r / n) * n;
    float P[EC_G], R2[EC_G], R3[EC_G], m1[EC_G], m2[EC_G], m3[EC_G];
    {
        const float4 *ur = reinterpret_cast<const float4 *>(u + (size_t)r * 4 * EC_G);
#pragma unroll
        for (int q = 0; q < EC_G / 4; ++q) {
            const float4 p = __ldg(ur + q), r2 = __ldg(ur + 2 * (EC_G / 4) + q), r3 = __ldg(ur + 3 * (EC_G / 4) + q);
            P[4 * q] = p.x; P[4 * q + 1] = p.y; P[4 * q + 2] = p.z; P[4 * q + 3] = p.w;
            R2[4 * q] = r2.x; R2[4 * q + 1] = r2.y; R2[4 * q + 2] = r2.z; R2[4 * q + 3] = r2.w;
            R3[4 * q] = r3.x; R3[4 * q + 1] = r3.y; R3[4 * q + 2] = r3.z; R3[4 * q + 3] = r3.w;
        }
    }
#pragma unroll
    for (int o = 0; o < EC_G; ++o) m1[o] = m2[o] = m3[o] = -3.402823466e38f;
    const int *ir = idx + (size_t)r * K;
    int prev = -1;
    for (int k = 0; k < K; ++k) {
        const int j = __ldg(ir + k);
        if (j == prev) continue;  // first-hit padding repeats an index: max is idempotent
        prev = j;
        const float4 *qr = reinterpret_cast<const float4 *>(u + (size_t)(base + j) * 4 * EC_G + EC_G);
        float l1[EC_G], l2[EC_G], l3[EC_G];
#pragma unroll
        for (int q = 0; q < EC_G / 4; ++q) {
            const float4 v = __ldg(qr + q);
            l1[4 * q] = fmaxf(P[4 * q] + v.x, 0.0f);
            l1[4 * q + 1] = fmaxf(P[4 * q + 1] + v.y, 0.0f);
            l1[4 * q + 2] = fmaxf(P[4 * q + 2] + v.z, 0.0f);
            l1[4 * q + 3] = fmaxf(P[4 * q + 3] + v.w, 0.0f);
        }
#pragma unroll
        for (int o = 0; o < EC_G; ++o) {
            float a = R2[o];
#pragma unroll
            for (int i = 0; i < EC_G; ++i) a = fmaf(w.w2a[i * EC_G + o], l1[i], a);
            l2[o] = fmaxf(a, 0.0f);
        }
#pragma unroll
        for (int o = 0; o < EC_G; ++o) {
            float a = R3[o];
#pragma unroll
            for (int i = 0; i < EC_G; ++i) a = fmaf(w.w3a[i * EC_G + o], l2[i], a);
#pragma unroll
            for (int i = 0; i < EC_G; ++i) a = fmaf(w.w3b[i * EC_G + o], l1[i], a);
            l3[o] = a;
        }
#pragma unroll
        for (int o = 0; o < EC_G; ++o) {
            m1[o] = fmaxf(m1[o], l1[o]);
            m2[o] = fmaxf(m2[o], l2[o]);
            m3[o] = fmaxf(m3[o], l3[o]);
        }
    }
    float4 *op = reinterpret_cast<float4 *>(out + (size_t)r * ldo);
#pragma unroll
    for (int q = 0; q < EC_G / 4; ++q) {
        op[q] = make_float4(m3[4 * q], m3[4 * q + 1], m3[4 * q + 2], m3[4 * q + 3]);
        op[EC_G / 4 + q] = make_float4(m2[4 * q], m2[4 * q + 1], m2[4 * q + 2], m2[4 * q + 3]);
        op[2 * (EC_G / 4) + q] = make_float4(m1[4 * q], m1[4 * q + 1], m1[4 * q + 2], m1[4 * q + 3]);
    }
    const float4 *tr = reinterpret_cast<const float4 *>(t + (size_t)r * EC_C);
#pragma unroll
    for (int q = 0; q < EC_C / 4; ++q) op[3 * (EC_G / 4) + q] = __ldg(tr + q);
}

}  // namespace spsk

using namespace spsk;

extern "C" {

SPSK_API int spsk_edge_conv_point(const spsk_edge_point_weights *w, int rows, const float *x, int ldx, float *t,
                                  float *u, spsk_stream_t stream) {
    SPSK_REQUIRE(w && x && t && u, SPSK_ERR_INVALID_ARG, "spsk_edge_conv_point: null pointer");
    SPSK_REQUIRE(rows >= 0 && w->cin > 0 && w->cin <= SPSK_EDGE_MAX_CIN && ldx >= w->cin, SPSK_ERR_INVALID_ARG,
                 "spsk_edge_conv_point: rows=%d cin=%d (max %d) ldx=%d", rows, w->cin, SPSK_EDGE_MAX_CIN, ldx);
    if (rows == 0) return SPSK_OK;
    const int grid = (rows + 127) / 128;
    cudaStream_t st = as_stream(stream);
    if (w->cin == 3)
        edge_point_kernel<3><<<grid, 128, 0, st>>>(*w, rows, x, ldx, t, u);
    else if (w->cin == 60)
        edge_point_kernel<60><<<grid, 128, 0, st>>>(*w, rows, x, ldx, t, u);
    else
        edge_point_kernel<0><<<grid, 128, 0, st>>>(*w, rows, x, ldx, t, u);
    SPSK_LAUNCH_CHECK("edge_point_kernel");
    return SPSK_OK;
}

SPSK_API int spsk_edge_conv_aggregate(const spsk_edge_aggr_weights *w, int b, int n, int k, const int *idx,
                                      const float *t, const float *u, float *out, int ldo, spsk_stream_t stream) {
    SPSK_REQUIRE(w && idx && t && u && out, SPSK_ERR_INVALID_ARG, "spsk_edge_conv_aggregate: null pointer");
    SPSK_REQUIRE(b >= 0 && n >= 0 && k > 0 && ldo >= SPSK_EDGE_CH + 3 * SPSK_EDGE_GROW && ldo % 4 == 0, SPSK_ERR_INVALID_ARG,
                 "spsk_edge_conv_aggregate: b=%d n=%d k=%d ldo=%d (need >= %d, multiple of 4)", b, n, k, ldo,
                 SPSK_EDGE_CH + 3 * SPSK_EDGE_GROW);
    const long long rows = (long long)b * n;
    if (rows == 0) return SPSK_OK;
    SPSK_REQUIRE((rows + 127) / 128 <= 0x7fffffffLL, SPSK_ERR_UNSUPPORTED, "spsk_edge_conv_aggregate: too many rows");
    edge_aggr_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, as_stream(stream)>>>(*w, b, n, k, idx, t, u, out, ldo);
    SPSK_LAUNCH_CHECK("edge_aggr_kernel");
    return SPSK_OK;
}

}  // extern "C"
