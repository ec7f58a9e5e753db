/*
 * oracle.c -- TEST INFRASTRUCTURE ONLY (checker + reported CPU baseline).  Never shipped, never on
 * the product path: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.
 *
 * A plain-C, CPU restatement of the algorithms of the reference's pointnet2_batch CUDA ops
 * (AlanLiangC/SPSNet, pcdet/ops/pointnet2/pointnet2_batch/src/).  Nothing here is copied from the
 * reference: each function re-states, in scalar C, what the cited kernel computes, INCLUDING the
 * fp32 rounding sequence the reference's nvcc build produces (verified on the SASS of the
 * reference objects rebuilt for sm_100a, see DESIGN.md "numerics"):
 *
 *      d2 = fmaf(dz, dz, fmaf(dx, dx, dy * dy))          (FMUL on the y term, then two FFMA)
 *
 * and the block-level arg-max tie-break of the FPS kernels, which is simulated literally
 * (per-thread strided scan + pairwise tree merge), not through a derived closed form, so that it
 * is an independent check of the closed-form key the CUDA kernels use.
 *
 * PARITY PIN: the reference ships no tests / golden vectors for this path.  The
 * pin is therefore (1) tests/golden/ fixtures produced by RUNNING the rebuilt, unmodified
 * reference CUDA ops (oracle/_ref) on a B200 with tests/golden/make_golden.py, and (2) the -m gpu
 * tests, which run reference, oracle and candidate side by side on the same inputs.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* reference: src/cuda_utils.h:10-14  opt_n_threads(): largest power of two <= n, clamped to [1,1024].
 * (restated with integer arithmetic; the survey verified the reference's log()-based form returns
 * the same value for every n in [1, 2^20].) */
ORC_API int orc_opt_n_threads(int n) {
    int p = 1;
    while (p * 2 <= n && p < 1024) p *= 2;
    return p;
}

/* squared distance with the reference build's rounding sequence (see header). */
static inline float sqdist3(float ax, float ay, float az, float bx, float by, float bz) {
    float dx = ax - bx, dy = ay - by, dz = az - bz;
    float t = dy * dy;
    t = fmaf(dx, dx, t);
    return fmaf(dz, dz, t);
}

/* ---- D-FPS ---------------------------------------------------------------------------------
 * reference: src/sampling_gpu.cu:93-209 farthest_point_sampling_kernel<block_size>, launched with
 * one block of S = opt_n_threads(n) threads per scene (src/sampling_gpu.cu:211-253).
 *   idx[0] = 0; for j in 1..m-1: every point k: d = |p_k - p_old|^2; temp[k] = min(d, temp[k]);
 *   thread t keeps the first strict maximum over k = t, t+S, ...; a pairwise tree (t vs t+h,
 *   h = S/2..1) keeps the left entry unless the right one is strictly greater.
 * xyz (b,n,3), temp (b,n) in/out (caller pre-fills 1e10, pointnet2_utils.py:26), idx (b,m). */
static void fps_scene(int n, int m, const float *xyz, float *temp, int *idx, int S,
                      float *dists, int *dists_i, const float *dist_matrix) {
    if (m <= 0) return;
    int old = 0;
    idx[0] = 0;
    for (int j = 1; j < m; ++j) {
        float x1 = 0.f, y1 = 0.f, z1 = 0.f;
        if (!dist_matrix) { x1 = xyz[old * 3 + 0]; y1 = xyz[old * 3 + 1]; z1 = xyz[old * 3 + 2]; }
        for (int t = 0; t < S; ++t) {
            int besti = 0;
            float best = -1.f;
            for (int k = t; k < n; k += S) {
                float d;
                if (dist_matrix) d = dist_matrix[(size_t)old * n + k];
                else d = sqdist3(xyz[k * 3 + 0], xyz[k * 3 + 1], xyz[k * 3 + 2], x1, y1, z1);
                float d2 = fminf(d, temp[k]);
                temp[k] = d2;
                besti = d2 > best ? k : besti;
                best = d2 > best ? d2 : best;
            }
            dists[t] = best;
            dists_i[t] = besti;
        }
        for (int h = S / 2; h >= 1; h /= 2) {
            for (int t = 0; t < h; ++t) {
                float v1 = dists[t], v2 = dists[t + h];
                int i1 = dists_i[t], i2 = dists_i[t + h];
                dists[t] = fmaxf(v1, v2);
                dists_i[t] = v2 > v1 ? i2 : i1;
            }
        }
        old = dists_i[0];
        idx[j] = old;
    }
}

ORC_API int orc_farthest_point_sampling(int b, int n, int m, const float *xyz, float *temp, int *idx) {
    int S = orc_opt_n_threads(n);
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < b; ++bi) {
        float *dists = (float *)malloc(sizeof(float) * S);
        int *dists_i = (int *)malloc(sizeof(int) * S);
        fps_scene(n, m, xyz + (size_t)bi * n * 3, temp + (size_t)bi * n, idx + (size_t)bi * m, S, dists, dists_i, NULL);
        free(dists);
        free(dists_i);
    }
    return 1;
}

/* ---- F-FPS over a precomputed (b,n,n) distance matrix ------------------------------------------
 * reference: src/sampling_gpu.cu:256-371 furthest_point_sampling_with_dist_kernel (same loop, d is
 * read from row `old` of the matrix). */
ORC_API int orc_furthest_point_sampling_with_dist(int b, int n, int m, const float *dist, float *temp, int *idx) {
    int S = orc_opt_n_threads(n);
#pragma omp parallel for schedule(dynamic, 1)
    for (int bi = 0; bi < b; ++bi) {
        float *dists = (float *)malloc(sizeof(float) * S);
        int *dists_i = (int *)malloc(sizeof(int) * S);
        fps_scene(n, m, NULL, temp + (size_t)bi * n, idx + (size_t)bi * m, S, dists, dists_i, dist + (size_t)bi * n * n);
        free(dists);
        free(dists_i);
    }
    return 2; /* the reference wrapper returns 2 (src/sampling.cpp:46-56) */
}

/* ---- gather: out[b,c,j] = points[b,c,idx[b,j]] ------------------------------------------------
 * reference: src/sampling_gpu.cu:8-24 gather_points_kernel_fast. */
ORC_API int orc_gather_points(int b, int c, int n, int npoints, const float *points, const int *idx, float *out) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * n;
            float *dst = out + ((size_t)bi * c + ci) * npoints;
            const int *ix = idx + (size_t)bi * npoints;
            for (int j = 0; j < npoints; ++j) dst[j] = src[ix[j]];
        }
    return 1;
}

/* reference: src/sampling_gpu.cu:46-63 gather_points_grad_kernel_fast (atomicAdd scatter; summation
 * order on the GPU is unspecified, here it is index order). grad_points must be pre-zeroed. */
ORC_API int orc_gather_points_grad(int b, int c, int n, int npoints, const float *grad_out, const int *idx, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * npoints;
            float *dst = grad_points + ((size_t)bi * c + ci) * n;
            const int *ix = idx + (size_t)bi * npoints;
            for (int j = 0; j < npoints; ++j) dst[ix[j]] += g[j];
        }
    return 1;
}

/* ---- ball query ----------------------------------------------------------------------------------
 * reference: src/ball_query_gpu.cu:9-45 ball_query_kernel_fast.  Scan k = 0..n-1 in index order,
 * keep the first nsample with d2 < radius^2 (radius^2 is an fp32 product, strict compare); on the
 * first hit all nsample slots are filled with that index; a centre without any hit leaves its row
 * untouched (the python wrapper pre-zeroes idx, pointnet2_utils.py:246). */
ORC_API int orc_ball_query(int b, int n, int m, float radius, int nsample, const float *new_xyz, const float *xyz, int *idx) {
    float radius2 = radius * radius;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int p = 0; p < m; ++p) {
            const float *c = new_xyz + ((size_t)bi * m + p) * 3;
            const float *pts = xyz + (size_t)bi * n * 3;
            int *o = idx + ((size_t)bi * m + p) * nsample;
            int cnt = 0;
            for (int k = 0; k < n; ++k) {
                float d2 = sqdist3(c[0], c[1], c[2], pts[k * 3 + 0], pts[k * 3 + 1], pts[k * 3 + 2]);
                if (d2 < radius2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[l] = k;
                    o[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;
                }
            }
        }
    return 1;
}

/* reference: src/ball_query_gpu.cu:70-117 ball_query_dilated_kernel_fast.  Two independent clauses
 * per point: (d2 == 0) and (min_r^2 <= d2 < max_r^2); both fire for a coincident point when
 * min_radius == 0, which inserts that index twice (quirk preserved). */
ORC_API int orc_ball_query_dilated(int b, int n, int m, float max_radius, float min_radius, int nsample,
                                   const float *new_xyz, const float *xyz, int *idx) {
    float r_max2 = max_radius * max_radius;
    float r_min2 = min_radius * min_radius;
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int p = 0; p < m; ++p) {
            const float *c = new_xyz + ((size_t)bi * m + p) * 3;
            const float *pts = xyz + (size_t)bi * n * 3;
            int *o = idx + ((size_t)bi * m + p) * nsample;
            int cnt = 0;
            for (int k = 0; k < n; ++k) {
                float d2 = sqdist3(c[0], c[1], c[2], pts[k * 3 + 0], pts[k * 3 + 1], pts[k * 3 + 2]);
                if (d2 == 0.f) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[l] = k;
                    o[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;
                }
                if (d2 >= r_min2 && d2 < r_max2) {
                    if (cnt == 0)
                        for (int l = 0; l < nsample; ++l) o[l] = k;
                    o[cnt] = k;
                    ++cnt;
                    if (cnt >= nsample) break;
                }
            }
        }
    return 1;
}

/* ---- group: out[b,c,p,s] = points[b,c,idx[b,p,s]] ---------------------------------------------
 * reference: src/group_points_gpu.cu:53-72 group_points_kernel_fast. */
ORC_API int orc_group_points(int b, int c, int n, int npoints, int nsample, const float *points, const int *idx, float *out) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * n;
            float *dst = out + ((size_t)bi * c + ci) * npoints * nsample;
            const int *ix = idx + (size_t)bi * npoints * nsample;
            for (int j = 0; j < npoints * nsample; ++j) dst[j] = src[ix[j]];
        }
    return 1;
}

/* reference: src/group_points_gpu.cu:14-31 group_points_grad_kernel_fast (atomicAdd scatter). */
ORC_API int orc_group_points_grad(int b, int c, int n, int npoints, int nsample, const float *grad_out, const int *idx, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * npoints * nsample;
            float *dst = grad_points + ((size_t)bi * c + ci) * n;
            const int *ix = idx + (size_t)bi * npoints * nsample;
            for (int j = 0; j < npoints * nsample; ++j) dst[ix[j]] += g[j];
        }
    return 1;
}

/* ---- three_nn -----------------------------------------------------------------------------------
 * reference: src/interpolate_gpu.cu:16-59 three_nn_kernel_fast.  float d compared against DOUBLE
 * running bests initialised to 1e40; strict '<' so the first index wins ties; outputs the three
 * squared distances cast to float (the python wrapper takes sqrt, pointnet2_utils.py:126). */
ORC_API void orc_three_nn(int b, int n, int m, const float *unknown, const float *known, float *dist2, int *idx) {
#pragma omp parallel for collapse(2) schedule(static)
    for (int bi = 0; bi < b; ++bi)
        for (int p = 0; p < n; ++p) {
            const float *u = unknown + ((size_t)bi * n + p) * 3;
            const float *kn = known + (size_t)bi * m * 3;
            double best1 = 1e40, best2 = 1e40, best3 = 1e40;
            int i1 = 0, i2 = 0, i3 = 0;
            for (int k = 0; k < m; ++k) {
                float d = sqdist3(u[0], u[1], u[2], kn[k * 3 + 0], kn[k * 3 + 1], kn[k * 3 + 2]);
                if (d < best1) {
                    best3 = best2; i3 = i2;
                    best2 = best1; i2 = i1;
                    best1 = d; i1 = k;
                } else if (d < best2) {
                    best3 = best2; i3 = i2;
                    best2 = d; i2 = k;
                } else if (d < best3) {
                    best3 = d; i3 = k;
                }
            }
            float *od = dist2 + ((size_t)bi * n + p) * 3;
            int *oi = idx + ((size_t)bi * n + p) * 3;
            od[0] = (float)best1; od[1] = (float)best2; od[2] = (float)best3;
            oi[0] = i1; oi[1] = i2; oi[2] = i3;
        }
}

/* ---- three_interpolate --------------------------------------------------------------------------
 * reference: src/interpolate_gpu.cu:84-104 three_interpolate_kernel_fast:
 *   out[b,c,n] = w0*p[i0] + w1*p[i1] + w2*p[i2], contracted by nvcc into
 *   fmaf(w2, p2, fmaf(w0, p0, w1*p1))   (FMUL on the index-1 term, then FFMA with the index-0 term, then
 *   FFMA with the index-2 term: read off the SASS of the rebuilt reference object, DESIGN.md "numerics"). */
ORC_API void orc_three_interpolate(int b, int c, int m, int n, const float *points, const int *idx, const float *weight, float *out) {
#pragma omp parallel for collapse(2)
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *src = points + ((size_t)bi * c + ci) * m;
            float *dst = out + ((size_t)bi * c + ci) * n;
            for (int p = 0; p < n; ++p) {
                const int *ix = idx + ((size_t)bi * n + p) * 3;
                const float *w = weight + ((size_t)bi * n + p) * 3;
                float t = w[1] * src[ix[1]];
                t = fmaf(w[0], src[ix[0]], t);
                dst[p] = fmaf(w[2], src[ix[2]], t);
            }
        }
}

/* reference: src/interpolate_gpu.cu:127-149 three_interpolate_grad_kernel_fast (3 atomicAdds). */
ORC_API void orc_three_interpolate_grad(int b, int c, int n, int m, const float *grad_out, const int *idx, const float *weight, float *grad_points) {
    for (int bi = 0; bi < b; ++bi)
        for (int ci = 0; ci < c; ++ci) {
            const float *g = grad_out + ((size_t)bi * c + ci) * n;
            float *dst = grad_points + ((size_t)bi * c + ci) * m;
            for (int p = 0; p < n; ++p) {
                const int *ix = idx + ((size_t)bi * n + p) * 3;
                const float *w = weight + ((size_t)bi * n + p) * 3;
                dst[ix[0]] += g[p] * w[0];
                dst[ix[1]] += g[p] * w[1];
                dst[ix[2]] += g[p] * w[2];
            }
        }
}

/* ---- closed-form FPS tie-break key, exported so tests can check it against the literal simulation
 * above.  Among points whose min-distance equals the maximum, the reference block picks the one
 * minimising (bit_reverse_{log2 S}(k mod S), k); rank(k) below orders exactly that. */
ORC_API uint32_t orc_fps_rank(int k, int S) {
    uint32_t L = 0;
    while ((1 << L) < S) ++L;
    uint32_t low = (uint32_t)k & (uint32_t)(S - 1), rev = 0;
    for (uint32_t i = 0; i < L; ++i) rev |= ((low >> i) & 1u) << (31 - i);
    return rev | ((uint32_t)k >> L);
}

/* =====================================================================================================
 * rotated BEV IoU / 3-D IoU / NMS of pcdet/ops/iou3d_nms, restated literally
 * (per-pair trigonometry, atan2 inside every comparison of the bubble sort, host-style greedy loop), i.e.
 * deliberately NOT organised like the CUDA kernels it checks.
 * PARITY PIN: tests/golden/iou3d_reference_cpu.npz holds outputs of the reference's own CPU implementation
 * (src/iou3d_cpu.cpp, compiled unmodified into oracle/_ref and run in the build container by
 * tests/golden/make_golden_iou3d.py); tests/test_oracle_cpu.py requires bit-equality with it.
 * ===================================================================================================== */
typedef struct { float x, y; } orc_pt;

static float orc_cross3(orc_pt p1, orc_pt p2, orc_pt p0) { /* iou3d_nms_kernel.cu:39-41 */
    return (p1.x - p0.x) * (p2.y - p0.y) - (p2.x - p0.x) * (p1.y - p0.y);
}
static float orc_fmin(float a, float b) { return a > b ? b : a; } /* iou3d_cpu.cpp:30-36 */
static float orc_fmax(float a, float b) { return a > b ? a : b; }

static int orc_in_box(const float *box, orc_pt p) { /* iou3d_nms_kernel.cu:48-58 */
    const float margin = 1e-2f;
    float c = cosf(-box[6]), s = sinf(-box[6]);
    float rx = (p.x - box[0]) * c + (p.y - box[1]) * (-s);
    float ry = (p.x - box[0]) * s + (p.y - box[1]) * c;
    return fabsf(rx) < box[3] / 2 + margin && fabsf(ry) < box[4] / 2 + margin;
}

static int orc_seg_x(orc_pt p1, orc_pt p0, orc_pt q1, orc_pt q0, orc_pt *ans) { /* iou3d_nms_kernel.cu:60-96 */
    const float eps = 1e-8f;
    int hit = orc_fmin(p0.x, p1.x) <= orc_fmax(q0.x, q1.x) && orc_fmin(q0.x, q1.x) <= orc_fmax(p0.x, p1.x) &&
              orc_fmin(p0.y, p1.y) <= orc_fmax(q0.y, q1.y) && orc_fmin(q0.y, q1.y) <= orc_fmax(p0.y, p1.y);
    if (!hit) return 0;
    float s1 = orc_cross3(q0, p1, p0), s2 = orc_cross3(p1, q1, p0);
    float s3 = orc_cross3(p0, q1, q0), s4 = orc_cross3(q1, p1, q0);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return 0;
    float s5 = orc_cross3(q1, p1, p0);
    if (fabsf(s5 - s1) > eps) {
        ans->x = (s5 * q0.x - s1 * q1.x) / (s5 - s1);
        ans->y = (s5 * q0.y - s1 * q1.y) / (s5 - s1);
    } else {
        float a0 = p0.y - p1.y, b0 = p1.x - p0.x, c0 = p0.x * p1.y - p1.x * p0.y;
        float a1 = q0.y - q1.y, b1 = q1.x - q0.x, c1 = q0.x * q1.y - q1.x * q0.y;
        float D = a0 * b1 - a1 * b0;
        ans->x = (b0 * c1 - b1 * c0) / D;
        ans->y = (a1 * c0 - a0 * c1) / D;
    }
    return 1;
}

static void orc_corners(const float *box, orc_pt *out /*5*/) { /* iou3d_nms_kernel.cu:112-158 */
    float hx = box[3] / 2, hy = box[4] / 2;
    float x1 = box[0] - hx, y1 = box[1] - hy, x2 = box[0] + hx, y2 = box[1] + hy;
    float c = cosf(box[6]), s = sinf(box[6]);
    float px[4] = {x1, x2, x2, x1}, py[4] = {y1, y1, y2, y2};
    for (int k = 0; k < 4; ++k) {
        out[k].x = (px[k] - box[0]) * c + (py[k] - box[1]) * (-s) + box[0];
        out[k].y = (px[k] - box[0]) * s + (py[k] - box[1]) * c + box[1];
    }
    out[4] = out[0];
}

static float orc_overlap1(const float *a, const float *b) { /* box_overlap, iou3d_nms_kernel.cu:108-216 */
    orc_pt ca[5], cb[5], pts[16], ctr = {0.f, 0.f};
    int cnt = 0;
    orc_corners(a, ca);
    orc_corners(b, cb);
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            orc_pt t;
            if (cnt < 16 && orc_seg_x(ca[i + 1], ca[i], cb[j + 1], cb[j], &t)) {
                pts[cnt++] = t;
                ctr.x = ctr.x + t.x;
                ctr.y = ctr.y + t.y;
            }
        }
    for (int k = 0; k < 4; ++k) {
        if (cnt < 16 && orc_in_box(a, cb[k])) { ctr.x = ctr.x + cb[k].x; ctr.y = ctr.y + cb[k].y; pts[cnt++] = cb[k]; }
        if (cnt < 16 && orc_in_box(b, ca[k])) { ctr.x = ctr.x + ca[k].x; ctr.y = ctr.y + ca[k].y; pts[cnt++] = ca[k]; }
    }
    if (cnt == 0) return 0.f; /* the reference divides 0/0 here and then sums nothing: area 0 */
    ctr.x /= cnt;
    ctr.y /= cnt;
    for (int j = 0; j < cnt - 1; ++j)
        for (int i = 0; i < cnt - j - 1; ++i)
            if (atan2f(pts[i].y - ctr.y, pts[i].x - ctr.x) > atan2f(pts[i + 1].y - ctr.y, pts[i + 1].x - ctr.x)) {
                orc_pt t = pts[i]; pts[i] = pts[i + 1]; pts[i + 1] = t;
            }
    float area = 0.f;
    for (int k = 0; k < cnt - 1; ++k) {
        float ax = pts[k].x - pts[0].x, ay = pts[k].y - pts[0].y;
        float bx = pts[k + 1].x - pts[0].x, by = pts[k + 1].y - pts[0].y;
        area += ax * by - ay * bx;
    }
    return fabsf(area) / 2.0f;
}

static float orc_iou_bev1(const float *a, const float *b) { /* iou3d_nms_kernel.cu:218-225 */
    float sa = a[3] * a[4], sb = b[3] * b[4], so = orc_overlap1(a, b);
    return so / fmaxf(sa + sb - so, 1e-8f);
}

static float orc_iou_normal1(const float *a, const float *b) { /* iou3d_nms_kernel.cu:321-333 */
    float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
    float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
    float w = fmaxf(right - left, 0.f), h = fmaxf(bottom - top, 0.f);
    float inter = w * h, sa = a[3] * a[4], sb = b[3] * b[4];
    return inter / fmaxf(sa + sb - inter, 1e-8f);
}

/* mode 0: BEV overlap, 1: BEV IoU, 2: 3-D IoU (iou3d_nms_utils.py:48-81, one fp32 rounding per torch op) */
ORC_API void orc_boxes_matrix(int na, const float *A, int nb, const float *B, float *out, int mode) {
#pragma omp parallel for schedule(dynamic, 4)
    for (int i = 0; i < na; ++i)
        for (int j = 0; j < nb; ++j) {
            const float *a = A + (size_t)i * 7, *b = B + (size_t)j * 7;
            float r;
            if (mode == 0) r = orc_overlap1(a, b);
            else if (mode == 1) r = orc_iou_bev1(a, b);
            else {
                float ov = orc_overlap1(a, b);
                float amax = a[2] + a[5] / 2, amin = a[2] - a[5] / 2, bmax = b[2] + b[5] / 2, bmin = b[2] - b[5] / 2;
                float oh = fmaxf(fminf(amax, bmax) - fmaxf(amin, bmin), 0.f);
                float o3 = ov * oh;
                float va = a[3] * a[4] * a[5], vb = b[3] * b[4] * b[5];
                r = o3 / fmaxf(va + vb - o3, 1e-6f);
            }
            out[(size_t)i * nb + j] = r;
        }
}

/* nms_gpu / nms_normal_gpu (iou3d_nms.cpp:90-188): boxes sorted by descending score; a box survives unless an
 * earlier SURVIVOR overlaps it by more than thresh.  Returns the number kept; keep[] = their positions. */
ORC_API int orc_nms(int n, const float *boxes, float thresh, int normal, long long *keep) {
    unsigned char *dead = (unsigned char *)calloc((size_t)(n > 0 ? n : 1), 1);
    int nk = 0;
    for (int i = 0; i < n; ++i) {
        if (dead[i]) continue;
        keep[nk++] = i;
        const float *a = boxes + (size_t)i * 7;
#pragma omp parallel for schedule(static)
        for (int j = i + 1; j < n; ++j) {
            if (dead[j]) continue;
            const float *b = boxes + (size_t)j * 7;
            float v = normal ? orc_iou_normal1(a, b) : orc_iou_bev1(a, b);
            if (v > thresh) dead[j] = 1;
        }
    }
    free(dead);
    return nk;
}

ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
