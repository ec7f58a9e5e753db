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

constexpr int EP_ROWS = 128;                  // points per CTA
constexpr int EP_OST = EC_C + 4 * EC_G + 1;   // 73: odd stride of the staged [t | u] rows (conflict-free row walks)

// One thread per point, but every global access goes through a shared-memory tile so that it is COALESCED: a thread walking
// its own 240-byte row touches 32 different lines per warp instruction (measured 0.29 ms at 262144 points, 8 % of the HBM
// rate); the tile of 128 rows is one contiguous block of the input and of each output.
template <int CIN>
__global__ void __launch_bounds__(EP_ROWS)
edge_point_kernel(const __grid_constant__ spsk_edge_point_weights w, int rows, const float *__restrict__ x, int ldx,
                  float *__restrict__ t, float *__restrict__ u) {
    __shared__ float sm[EP_ROWS * EP_OST];
    const int cin = CIN > 0 ? CIN : w.cin;
    const int sin = cin | 1;                  // odd input stride
    const int tid = threadIdx.x;
    const int r0 = blockIdx.x * EP_ROWS;
    const int nr = min(EP_ROWS, rows - r0);
    for (int i = tid; i < nr * cin; i += EP_ROWS) {
        const int row = i / cin, k = i - row * cin;
        sm[row * sin + k] = __ldg(x + (size_t)(r0 + row) * ldx + k);
    }
    __syncthreads();
    float tv[EC_C];
#pragma unroll
    for (int o = 0; o < EC_C; ++o) tv[o] = w.bt[o];
    if (tid < nr) {
        const float *xr = sm + tid * sin;
        if (CIN > 0) {
#pragma unroll
            for (int k = 0; k < (CIN > 0 ? CIN : 1); ++k) {
                const float xv = xr[k];
#pragma unroll
                for (int o = 0; o < EC_C; ++o) tv[o] = fmaf(w.wt[k * EC_C + o], xv, tv[o]);
            }
        } else {
            for (int k = 0; k < cin; ++k) {
                const float xv = xr[k];
#pragma unroll
                for (int o = 0; o < EC_C; ++o) tv[o] = fmaf(w.wt[k * EC_C + o], xv, tv[o]);
            }
        }
        if (w.relu) {
#pragma unroll
            for (int o = 0; o < EC_C; ++o) tv[o] = fmaxf(tv[o], 0.0f);
        }
    }
    __syncthreads();   // every thread is done with the input tile: reuse the buffer for the outputs
    if (tid < nr) {
        float *orow = sm + tid * EP_OST;
#pragma unroll
        for (int o = 0; o < EC_C; ++o) orow[o] = tv[o];
#pragma unroll
        for (int o = 0; o < 4 * EC_G; ++o) {
            float a = w.c[o];
#pragma unroll
            for (int k = 0; k < EC_C; ++k) a = fmaf(w.m[k * 4 * EC_G + o], tv[k], a);
            orow[EC_C + o] = a;
        }
    }
    __syncthreads();
    float *tg = t + (size_t)r0 * EC_C;
    for (int i = tid; i < nr * EC_C; i += EP_ROWS) {
        const int row = i / EC_C, k = i - row * EC_C;
        tg[i] = sm[row * EP_OST + k];
    }
    float *ug = u + (size_t)r0 * 4 * EC_G;
    for (int i = tid; i < nr * 4 * EC_G; i += EP_ROWS) {
        const int row = i / (4 * EC_G), k = i - row * (4 * EC_G);
        ug[i] = sm[row * EP_OST + EC_C + k];
    }
}

__global__ void __launch_bounds__(128)
edge_aggr_kernel(const __grid_constant__ spsk_edge_aggr_weights w, int b, int n, int K, const int *__restrict__ idx,
                 const float *__restrict__ t, const float *__restrict__ u, float *__restrict__ out, int ldo) {
    const long long total = (long long)b * n;
    const long long r0 = (long long)blockIdx.x * blockDim.x;
    const bool live = r0 + threadIdx.x < total;
    const long long r = live ? r0 + threadIdx.x : total - 1;   // idle threads of the last CTA shadow the last point
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
    // Two neighbours per step: every weight fetched from the constant bank feeds two FFMAs.  The lists are padded with
    // their first entry (ball query semantics), so once an entry repeats the first one the rest is padding; an odd tail
    // pairs the last neighbour with itself (max is idempotent).
    const int j_first = __ldg(ir);
    for (int k = 0; k < K; k += 2) {
        const int ja = __ldg(ir + k);
        const int jb = k + 1 < K ? __ldg(ir + k + 1) : ja;
        if (k > 0 && ja == j_first) break;             // padding from here on
        const float4 *qa = reinterpret_cast<const float4 *>(u + (size_t)(base + ja) * 4 * EC_G + EC_G);
        const float4 *qb = reinterpret_cast<const float4 *>(u + (size_t)(base + jb) * 4 * EC_G + EC_G);
        float l1a[EC_G], l1b[EC_G], l2a[EC_G], l2b[EC_G];
#pragma unroll
        for (int q = 0; q < EC_G / 4; ++q) {
            const float4 va = __ldg(qa + q), vb = __ldg(qb + q);
            l1a[4 * q] = fmaxf(P[4 * q] + va.x, 0.0f);         l1b[4 * q] = fmaxf(P[4 * q] + vb.x, 0.0f);
            l1a[4 * q + 1] = fmaxf(P[4 * q + 1] + va.y, 0.0f); l1b[4 * q + 1] = fmaxf(P[4 * q + 1] + vb.y, 0.0f);
            l1a[4 * q + 2] = fmaxf(P[4 * q + 2] + va.z, 0.0f); l1b[4 * q + 2] = fmaxf(P[4 * q + 2] + vb.z, 0.0f);
            l1a[4 * q + 3] = fmaxf(P[4 * q + 3] + va.w, 0.0f); l1b[4 * q + 3] = fmaxf(P[4 * q + 3] + vb.w, 0.0f);
        }
#pragma unroll
        for (int o = 0; o < EC_G; ++o) {
            float a0 = R2[o], a1 = R2[o];
#pragma unroll
            for (int i = 0; i < EC_G; ++i) {
                const float wv = w.w2a[i * EC_G + o];
                a0 = fmaf(wv, l1a[i], a0);
                a1 = fmaf(wv, l1b[i], a1);
            }
            l2a[o] = fmaxf(a0, 0.0f);
            l2b[o] = fmaxf(a1, 0.0f);
        }
#pragma unroll
        for (int o = 0; o < EC_G; ++o) {
            float a0 = R3[o], a1 = R3[o];
#pragma unroll
            for (int i = 0; i < EC_G; ++i) {
                const float wv = w.w3a[i * EC_G + o];
                a0 = fmaf(wv, l2a[i], a0);
                a1 = fmaf(wv, l2b[i], a1);
            }
#pragma unroll
            for (int i = 0; i < EC_G; ++i) {
                const float wv = w.w3b[i * EC_G + o];
                a0 = fmaf(wv, l1a[i], a0);
                a1 = fmaf(wv, l1b[i], a1);
            }
            m3[o] = fmaxf(m3[o], fmaxf(a0, a1));
        }
#pragma unroll
        for (int o = 0; o < EC_G; ++o) {
            m1[o] = fmaxf(m1[o], fmaxf(l1a[o], l1b[o]));
            m2[o] = fmaxf(m2[o], fmaxf(l2a[o], l2b[o]));
        }
    }
    // stage the 60-float rows through shared memory so the global stores are coalesced (see edge_point_kernel)
    __shared__ float so[128 * 61];
    float *orow = so + threadIdx.x * 61;
#pragma unroll
    for (int o = 0; o < EC_G; ++o) {
        orow[o] = m3[o];
        orow[EC_G + o] = m2[o];
        orow[2 * EC_G + o] = m1[o];
    }
    {
        const float4 *tr = reinterpret_cast<const float4 *>(t + (size_t)r * EC_C);
#pragma unroll
        for (int q = 0; q < EC_C / 4; ++q) {
            const float4 v = __ldg(tr + q);
            orow[3 * EC_G + 4 * q] = v.x; orow[3 * EC_G + 4 * q + 1] = v.y; orow[3 * EC_G + 4 * q + 2] = v.z; orow[3 * EC_G + 4 * q + 3] = v.w;
        }
    }
    __syncthreads();
    const int nr = (int)min((long long)blockDim.x, total - r0);
    constexpr int W = EC_C + 3 * EC_G;   // 60
    for (int i = threadIdx.x; i < nr * W; i += blockDim.x) {
        const int row = i / W, k = i - row * W;
        out[(size_t)(r0 + row) * ldo + k] = so[row * 61 + k];
    }
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
