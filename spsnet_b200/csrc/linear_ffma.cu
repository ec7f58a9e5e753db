// linear_ffma.cu -- fp32 CUDA-core (FFMA) shared-MLP layers: grouped (gather -> 1x1 conv -> BN -> ReLU
// [-> pool over nsample]) and point-wise (Conv1d 1x1 on channel-major tensors).
//
// Replaces, per layer, the reference's chain grouping_operation x2 + subtract + torch.cat
// (pointnet2_utils.py:307-315) + Conv2d(1x1, bias=False) + BatchNorm2d(eval) + ReLU
// (pointnet2_modules.py:204-211) [+ F.max_pool2d over nsample, :433-436], and the Conv1d+BN1d+ReLU
// aggregation / confidence / vote layers (:216-243, 485-500).  The grouped (B,3+C,npoint,nsample)
// tensor is never materialised: the gather is the A-operand loader of the GEMM, BN is folded into the
// weights by the host, bias + ReLU + max-pool run in the epilogue.
//
// This is the exact-fp32 path: it serves layers whose channel width is too small to be a real dense
// contraction and is the numerical reference (on the GPU) for the tensor-core path (sa_mma.cu).
// Tile: 128 rows x 64 channels x 16 k, 256 threads, 8x4 accumulators per thread.
#include "common.cuh"

namespace spsk {

constexpr int LF_BM = 128, LF_BN = 64, LF_BK = 16, LF_THREADS = 256;
constexpr int LF_TM = 8, LF_TN = 4;

enum { A_ROWS = 0, A_GATHER = 1, A_CHMAJOR = 2 };
enum { O_ROWS = 0, O_POOL_MAX = 1, O_POOL_AVG = 2, O_CHMAJOR = 3 };

struct LinArgs {
    long long rows;      // total rows (b*m*nsample or b*m)
    int c_in, c_out;
    const float *wt;     // (c_in, c_out)
    const float *bias;   // (c_out) or NULL
    int relu;
    // A operand
    const float *in;     // A_ROWS: (rows, c_in); A_CHMAJOR: (b, c_in, m)
    spsk_group_desc g;   // A_GATHER
    int xyz_ch;          // 3 if use_xyz else 0
    // output
    float *out;          // O_ROWS: (rows, c_out); O_CHMAJOR: (b, c_out, m); O_POOL_*: (b, c_total, m)
    int m, nsample, c_total, co_off;
};

template <int AMODE, int OMODE>
__global__ void __launch_bounds__(LF_THREADS)
linear_ffma_kernel(LinArgs a) {
    __shared__ __align__(16) float As[LF_BK][LF_BM + 4];
    __shared__ __align__(16) float Ws[LF_BK][LF_BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const long long row0 = (long long)blockIdx.x * LF_BM;
    const int col0 = blockIdx.y * LF_BN;

    // A-loader mapping: each thread owns one tile row and 8 consecutive k
    const int lrow = tid >> 1;
    const int lk0 = (tid & 1) * 8;
    const long long grow = row0 + lrow;
    const bool row_ok = grow < a.rows;
    // per-row constants for the gather / channel-major loaders
    int src_j = 0;
    float cx = 0.f, cy = 0.f, cz = 0.f;
    const float *feat_b = nullptr, *xyz_b = nullptr, *in_b = nullptr;
    if (AMODE == A_GATHER && row_ok) {
        const long long q = grow / a.g.nsample;  // b*m + p
        const int bb = (int)(q / a.g.m);
        src_j = __ldg(a.g.idx + grow);
        if (a.xyz_ch) {
            const float *c = a.g.new_xyz + q * 3;
            cx = __ldg(c); cy = __ldg(c + 1); cz = __ldg(c + 2);
            xyz_b = a.g.xyz + ((size_t)bb * a.g.n + src_j) * 3;
        }
        if (a.g.c_feat) feat_b = a.g.features + (size_t)bb * a.g.c_feat * a.g.n + src_j;
    }
    if (AMODE == A_CHMAJOR && row_ok) {
        const int bb = (int)(grow / a.m);
        const int p = (int)(grow - (long long)bb * a.m);
        in_b = a.in + (size_t)bb * a.c_in * a.m + p;
    }

    float acc[LF_TM][LF_TN];
#pragma unroll
    for (int i = 0; i < LF_TM; ++i)
#pragma unroll
        for (int j = 0; j < LF_TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < a.c_in; k0 += LF_BK) {
        // ---- stage A tile (k-major in smem)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int k = k0 + lk0 + i;
            float v = 0.f;
            if (row_ok && k < a.c_in) {
                if (AMODE == A_ROWS) {
                    v = __ldg(a.in + grow * a.c_in + k);
                } else if (AMODE == A_CHMAJOR) {
                    v = __ldg(in_b + (size_t)k * a.m);
                } else {
                    if (k < a.xyz_ch) {
                        const float ctr = (k == 0) ? cx : (k == 1 ? cy : cz);
                        v = __fsub_rn(__ldg(xyz_b + k), ctr);  // grouped_xyz -= new_xyz (pointnet2_utils.py:310)
                    } else {
                        v = __ldg(feat_b + (size_t)(k - a.xyz_ch) * a.g.n);
                    }
                }
            }
            As[lk0 + i][lrow] = v;
        }
        // ---- stage W tile
        {
            const int wk = tid >> 4;            // 0..15
            const int wc = (tid & 15) * 4;      // 0..60
            const int k = k0 + wk;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int c = col0 + wc + j;
                Ws[wk][wc + j] = (k < a.c_in && c < a.c_out) ? __ldg(a.wt + (size_t)k * a.c_out + c) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < LF_BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4 *>(&As[kk][ty * LF_TM]);
            const float4 a1 = *reinterpret_cast<const float4 *>(&As[kk][ty * LF_TM + 4]);
            const float4 w = *reinterpret_cast<const float4 *>(&Ws[kk][tx * LF_TN]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int i = 0; i < LF_TM; ++i)
#pragma unroll
                for (int j = 0; j < LF_TN; ++j) acc[i][j] = fmaf(av[i], wv[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue: bias (folded BN shift), ReLU, store / pool
#pragma unroll
    for (int j = 0; j < LF_TN; ++j) {
        const int c = col0 + tx * LF_TN + j;
        if (c >= a.c_out) continue;
        const float bv = a.bias ? __ldg(a.bias + c) : 0.f;
        long long cur_q = -1;
        float cur = 0.f;
#pragma unroll
        for (int i = 0; i < LF_TM; ++i) {
            const long long r = row0 + ty * LF_TM + i;
            if (r >= a.rows) break;
            float v = acc[i][j] + bv;
            if (a.relu) v = fmaxf(v, 0.f);
            if (OMODE == O_ROWS) {
                a.out[r * a.c_out + c] = v;
            } else if (OMODE == O_CHMAJOR) {
                const long long bb = r / a.m;
                const long long p = r - bb * a.m;
                a.out[((size_t)bb * a.c_out + c) * a.m + p] = v;
            } else {
                const long long q = r / a.nsample;
                if (q != cur_q) {
                    if (cur_q >= 0) {
                        const long long bb = cur_q / a.m, p = cur_q - bb * a.m;
                        float *dst = a.out + ((size_t)bb * a.c_total + a.co_off + c) * a.m + p;
                        if (OMODE == O_POOL_MAX) atomicMax(reinterpret_cast<int *>(dst), __float_as_int(cur));
                        else atomicAdd(dst, cur / (float)a.nsample);
                    }
                    cur_q = q;
                    cur = v;
                } else {
                    cur = (OMODE == O_POOL_MAX) ? fmaxf(cur, v) : cur + v;
                }
            }
        }
        if (OMODE == O_POOL_MAX || OMODE == O_POOL_AVG) {
            if (cur_q >= 0) {
                const long long bb = cur_q / a.m, p = cur_q - bb * a.m;
                float *dst = a.out + ((size_t)bb * a.c_total + a.co_off + c) * a.m + p;
                if (OMODE == O_POOL_MAX) atomicMax(reinterpret_cast<int *>(dst), __float_as_int(cur));
                else atomicAdd(dst, cur / (float)a.nsample);
            }
        }
    }
}

template <int AMODE>
static int launch_linear(const LinArgs &a, int omode, cudaStream_t st) {
    const long long tiles = (a.rows + LF_BM - 1) / LF_BM;
    SPSK_REQUIRE(tiles <= 0x7FFFFFFFLL, SPSK_ERR_UNSUPPORTED, "linear: too many rows (%lld)", a.rows);
    dim3 grid((unsigned)tiles, (a.c_out + LF_BN - 1) / LF_BN);
    switch (omode) {
        case O_ROWS: linear_ffma_kernel<AMODE, O_ROWS><<<grid, LF_THREADS, 0, st>>>(a); break;
        case O_POOL_MAX: linear_ffma_kernel<AMODE, O_POOL_MAX><<<grid, LF_THREADS, 0, st>>>(a); break;
        case O_POOL_AVG: linear_ffma_kernel<AMODE, O_POOL_AVG><<<grid, LF_THREADS, 0, st>>>(a); break;
        default: linear_ffma_kernel<AMODE, O_CHMAJOR><<<grid, LF_THREADS, 0, st>>>(a); break;
    }
    SPSK_LAUNCH_CHECK("linear_ffma_kernel");
    return SPSK_OK;
}

}  // namespace spsk

extern "C" int spsk_grouped_linear(const spsk_group_desc *g, int gather, const float *in_rows, int c_in,
                                   const float *wt, const float *bias, int c_out, int relu, int pool,
                                   float *out_rows, float *out_pooled, int c_total, int co_off,
                                   spsk_stream_t stream) {
    using namespace spsk;
    SPSK_REQUIRE(g != nullptr, SPSK_ERR_INVALID_ARG, "grouped_linear: null descriptor");
    SPSK_REQUIRE(g->b >= 0 && g->m >= 0 && g->nsample >= 1 && g->n >= 0 && c_in >= 1 && c_out >= 1, SPSK_ERR_INVALID_ARG,
                 "grouped_linear: bad sizes b=%d m=%d nsample=%d c_in=%d c_out=%d", g->b, g->m, g->nsample, c_in, c_out);
    SPSK_REQUIRE(wt != nullptr, SPSK_ERR_INVALID_ARG, "grouped_linear: null weights");
    SPSK_REQUIRE(pool >= 0 && pool <= 2, SPSK_ERR_INVALID_ARG, "grouped_linear: pool=%d", pool);
    LinArgs a{};
    a.rows = (long long)g->b * g->m * g->nsample;
    if (a.rows == 0) return SPSK_OK;
    a.c_in = c_in; a.c_out = c_out; a.wt = wt; a.bias = bias; a.relu = relu;
    a.m = g->m; a.nsample = g->nsample; a.c_total = c_total; a.co_off = co_off;
    a.g = *g;
    if (pool) {
        SPSK_REQUIRE(out_pooled != nullptr, SPSK_ERR_INVALID_ARG, "grouped_linear: pool without out_pooled");
        SPSK_REQUIRE(pool != 1 || relu, SPSK_ERR_INVALID_ARG, "grouped_linear: max-pool epilogue needs relu (non-negative values)");
        SPSK_REQUIRE(co_off >= 0 && co_off + c_out <= c_total, SPSK_ERR_INVALID_ARG, "grouped_linear: channel window [%d,%d) outside c_total=%d", co_off, co_off + c_out, c_total);
        a.out = out_pooled;
    } else {
        SPSK_REQUIRE(out_rows != nullptr, SPSK_ERR_INVALID_ARG, "grouped_linear: null out_rows");
        a.out = out_rows;
    }
    if (gather) {
        a.xyz_ch = g->use_xyz ? 3 : 0;
        SPSK_REQUIRE(c_in == a.xyz_ch + g->c_feat, SPSK_ERR_INVALID_ARG, "grouped_linear: c_in=%d != %d xyz + %d feature channels", c_in, a.xyz_ch, g->c_feat);
        SPSK_REQUIRE(g->idx && (!a.xyz_ch || (g->xyz && g->new_xyz)) && (!g->c_feat || g->features), SPSK_ERR_INVALID_ARG, "grouped_linear: null gather source");
        return launch_linear<A_GATHER>(a, pool, as_stream(stream));
    }
    SPSK_REQUIRE(in_rows != nullptr, SPSK_ERR_INVALID_ARG, "grouped_linear: null in_rows");
    a.in = in_rows;
    return launch_linear<A_ROWS>(a, pool, as_stream(stream));
}

extern "C" int spsk_pointwise_linear(int b, int m, const float *in, int c_in, const float *wt, const float *bias,
                                     int c_out, int relu, float *out, spsk_stream_t stream) {
    using namespace spsk;
    SPSK_REQUIRE(b >= 0 && m >= 0 && c_in >= 1 && c_out >= 1, SPSK_ERR_INVALID_ARG, "pointwise_linear: bad sizes");
    if (b == 0 || m == 0) return SPSK_OK;
    SPSK_REQUIRE(in && wt && out, SPSK_ERR_INVALID_ARG, "pointwise_linear: null pointer");
    LinArgs a{};
    a.rows = (long long)b * m;
    a.c_in = c_in; a.c_out = c_out; a.wt = wt; a.bias = bias; a.relu = relu;
    a.in = in; a.out = out; a.m = m; a.nsample = 1;
    return launch_linear<A_CHMAJOR>(a, O_CHMAJOR, as_stream(stream));
}
