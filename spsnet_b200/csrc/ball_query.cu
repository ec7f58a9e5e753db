// ball_query.cu -- ball query (single radius, dilated shell, and fused multi-scale).
//
// Replaces ball_query_kernel_fast / ball_query_dilated_kernel_fast of the reference
// (src/ball_query_gpu.cu:9-45, 70-117): one THREAD per centre walking all n points serially, every
// lane of a warp re-reading the same xyz[k] from global, divergent early exit, one pass per radius.
//
// Here: one WARP per centre.  The CTA stages xyz tiles into shared memory once (SoA, conflict-free),
// each lane tests one point of a 32-point chunk against every radius of the MSG layer in the same
// pass, a ballot + prefix-popcount keeps index order, and the warp stops as soon as every scale has
// its nsample neighbours.  Results are identical to the reference (same fp32 expression via
// common.cuh::sqdist3, strict compare against the fp32 product radius*radius, first-hit padding).
#include "common.cuh"

namespace spsk {

constexpr int BQ_WARPS = 8;      // centres per CTA
constexpr int BQ_TILE = 1024;    // points staged per tile (12 KB)

struct BqScales {
    float r2[SPSK_MAX_SCALES];      // radius*radius (fp32 product, like the reference)
    float r2_min[SPSK_MAX_SCALES];  // dilated only: min_radius^2
    int nsample[SPSK_MAX_SCALES];
    int *idx[SPSK_MAX_SCALES];
};

// DILATED: the reference's two-clause predicate (d2 == 0) and (r_min^2 <= d2 < r_max^2); a point can
// contribute two entries.  WRITE_EMPTY: rows without any hit are written as zeros (fused path) instead
// of being left untouched (reference ABI; the caller pre-zeroes).
template <int NS, bool DILATED, bool WRITE_EMPTY>
__global__ void __launch_bounds__(BQ_WARPS * 32)
ball_query_kernel(int n, int m, BqScales sc, const float *__restrict__ new_xyz, const float *__restrict__ xyz) {
    __shared__ float sx[BQ_TILE], sy[BQ_TILE], sz[BQ_TILE];
    const int b = blockIdx.y;
    const int warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();
    const int p = blockIdx.x * BQ_WARPS + warp;
    const bool active = p < m;
    const float *pts = xyz + (size_t)b * n * 3;

    float cx = 0.f, cy = 0.f, cz = 0.f;
    if (active) {
        const float *c = new_xyz + ((size_t)b * m + p) * 3;
        cx = __ldg(c); cy = __ldg(c + 1); cz = __ldg(c + 2);
    }
    int cnt[NS], first[NS];
    int *out[NS];
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        cnt[s] = 0;
        first[s] = 0;
        out[s] = sc.idx[s] + ((size_t)b * m + (active ? p : 0)) * sc.nsample[s];
    }
    bool done = !active;

    for (int t0 = 0; t0 < n; t0 += BQ_TILE) {
        const int tn = min(BQ_TILE, n - t0);
        // (the __syncthreads_and at the bottom of the previous iteration guards tile reuse)
        for (int i = threadIdx.x; i < tn * 3; i += BQ_WARPS * 32) {
            const float v = __ldg(pts + (size_t)t0 * 3 + i);
            const int k = i / 3, c = i - 3 * k;
            (c == 0 ? sx : (c == 1 ? sy : sz))[k] = v;
        }
        __syncthreads();
        if (!done) {
            for (int c0 = 0; c0 < tn; c0 += 32) {
                const int kl = c0 + (int)lane;
                const bool valid = kl < tn;
                const float d2 = valid ? sqdist3(cx, cy, cz, sx[kl], sy[kl], sz[kl]) : 0.f;
                const int k = t0 + kl;
                bool all_full = true;
#pragma unroll
                for (int s = 0; s < NS; ++s) {
                    const int ns = sc.nsample[s];
                    if (cnt[s] < ns) {
                        if (!DILATED) {
                            const bool hit = valid && (d2 < sc.r2[s]);
                            const uint32_t bal = __ballot_sync(0xFFFFFFFFu, hit);
                            if (bal) {
                                if (cnt[s] == 0) first[s] = t0 + c0 + __ffs(bal) - 1;
                                const int pos = cnt[s] + __popc(bal & ((1u << lane) - 1u));
                                if (hit && pos < ns) out[s][pos] = k;
                                cnt[s] += __popc(bal);
                            }
                        } else {
                            const bool h0 = valid && (d2 == 0.f);
                            const bool h1 = valid && (d2 >= sc.r2_min[s]) && (d2 < sc.r2[s]);
                            const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, h0);
                            const uint32_t b1 = __ballot_sync(0xFFFFFFFFu, h1);
                            if (b0 | b1) {
                                if (cnt[s] == 0) first[s] = t0 + c0 + __ffs(b0 | b1) - 1;
                                const uint32_t lt = (1u << lane) - 1u;
                                int pos = cnt[s] + __popc(b0 & lt) + __popc(b1 & lt);
                                if (h0) { if (pos < ns) out[s][pos] = k; ++pos; }
                                if (h1 && pos < ns) out[s][pos] = k;
                                cnt[s] += __popc(b0) + __popc(b1);
                            }
                        }
                    }
                    all_full = all_full && (cnt[s] >= ns);
                }
                if (all_full) { done = true; break; }
            }
        }
        if (__syncthreads_and(done)) break;
    }

    if (!active) return;
    __syncwarp();
#pragma unroll
    for (int s = 0; s < NS; ++s) {
        const int ns = sc.nsample[s];
        const int c = min(cnt[s], ns);
        if (c > 0) {
            for (int l = c + (int)lane; l < ns; l += 32) out[s][l] = first[s];  // first-hit padding
        } else if (WRITE_EMPTY) {
            for (int l = (int)lane; l < ns; l += 32) out[s][l] = 0;
        }
    }
}

template <bool DILATED, bool WRITE_EMPTY>
static int launch_bq(int b, int n, int m, int nscales, const BqScales &sc, const float *new_xyz, const float *xyz,
                     cudaStream_t st) {
    dim3 grid((m + BQ_WARPS - 1) / BQ_WARPS, b);
    dim3 block(BQ_WARPS * 32);
    switch (nscales) {
        case 1: ball_query_kernel<1, DILATED, WRITE_EMPTY><<<grid, block, 0, st>>>(n, m, sc, new_xyz, xyz); break;
        case 2: ball_query_kernel<2, DILATED, WRITE_EMPTY><<<grid, block, 0, st>>>(n, m, sc, new_xyz, xyz); break;
        case 3: ball_query_kernel<3, DILATED, WRITE_EMPTY><<<grid, block, 0, st>>>(n, m, sc, new_xyz, xyz); break;
        default: ball_query_kernel<4, DILATED, WRITE_EMPTY><<<grid, block, 0, st>>>(n, m, sc, new_xyz, xyz); break;
    }
    SPSK_LAUNCH_CHECK("ball_query_kernel");
    return SPSK_OK;
}

static int check_bq(int b, int n, int m, const float *new_xyz, const float *xyz) {
    SPSK_REQUIRE(b >= 0 && n >= 0 && m >= 0, SPSK_ERR_INVALID_ARG, "ball_query: bad sizes b=%d n=%d m=%d", b, n, m);
    SPSK_REQUIRE(b <= 65535, SPSK_ERR_UNSUPPORTED, "ball_query: b=%d > 65535", b);
    SPSK_REQUIRE(new_xyz && xyz, SPSK_ERR_INVALID_ARG, "ball_query: null pointer");
    return SPSK_OK;
}

}  // namespace spsk

extern "C" int spsk_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz,
                               const float *xyz, int *idx, spsk_stream_t stream) {
    using namespace spsk;
    if (int rc = check_bq(b, n, m, new_xyz, xyz)) return rc;
    SPSK_REQUIRE(idx && nsample >= 1, SPSK_ERR_INVALID_ARG, "ball_query: nsample=%d / null idx", nsample);
    if (b == 0 || m == 0) return SPSK_OK;
    BqScales sc{};
    sc.r2[0] = radius * radius;
    sc.nsample[0] = nsample;
    sc.idx[0] = idx;
    return launch_bq<false, false>(b, n, m, 1, sc, new_xyz, xyz, as_stream(stream));
}

extern "C" int spsk_ball_query_dilated(int b, int n, int m, float max_radius, float min_radius, int nsample,
                                       const float *new_xyz, const float *xyz, int *idx, spsk_stream_t stream) {
    using namespace spsk;
    if (int rc = check_bq(b, n, m, new_xyz, xyz)) return rc;
    SPSK_REQUIRE(idx && nsample >= 1, SPSK_ERR_INVALID_ARG, "ball_query_dilated: nsample=%d / null idx", nsample);
    if (b == 0 || m == 0) return SPSK_OK;
    BqScales sc{};
    sc.r2[0] = max_radius * max_radius;
    sc.r2_min[0] = min_radius * min_radius;
    sc.nsample[0] = nsample;
    sc.idx[0] = idx;
    return launch_bq<true, false>(b, n, m, 1, sc, new_xyz, xyz, as_stream(stream));
}

extern "C" int spsk_ball_query_msg(int b, int n, int m, int nscales, const float *radius, const int *nsample,
                                   const float *new_xyz, const float *xyz, int *const *idx, spsk_stream_t stream) {
    using namespace spsk;
    if (int rc = check_bq(b, n, m, new_xyz, xyz)) return rc;
    SPSK_REQUIRE(nscales >= 1 && nscales <= SPSK_MAX_SCALES && radius && nsample && idx, SPSK_ERR_INVALID_ARG,
                 "ball_query_msg: nscales=%d (1..%d) / null pointer", nscales, SPSK_MAX_SCALES);
    if (b == 0 || m == 0) return SPSK_OK;
    BqScales sc{};
    for (int s = 0; s < nscales; ++s) {
        SPSK_REQUIRE(nsample[s] >= 1 && idx[s], SPSK_ERR_INVALID_ARG, "ball_query_msg: scale %d nsample=%d / null idx", s, nsample[s]);
        sc.r2[s] = radius[s] * radius[s];
        sc.nsample[s] = nsample[s];
        sc.idx[s] = idx[s];
    }
    return launch_bq<false, true>(b, n, m, nscales, sc, new_xyz, xyz, as_stream(stream));
}
