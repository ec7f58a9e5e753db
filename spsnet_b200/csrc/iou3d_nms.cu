// iou3d_nms.cu -- rotated BEV overlap / IoU, 3-D IoU and NMS, entirely on the device.
//
// Replaces pcdet/ops/iou3d_nms:
//   boxes_overlap_bev_gpu / boxes_iou_bev_gpu  (src/iou3d_nms.cpp:48-88  -> src/iou3d_nms_kernel.cu:236-265)
//   boxes_iou3d_gpu                            (iou3d_nms_utils.py:48-81: 1 kernel + ~14 torch ops -> one launch)
//   nms_gpu / nms_normal_gpu                   (src/iou3d_nms.cpp:90-188 -> src/iou3d_nms_kernel.cu:267-365)
//
// What is different from the reference (results are the same):
//   * the reference NMS allocates the N x N/64 bit mask with cudaMalloc, copies it to the HOST, runs the greedy
//     suppression loop on the CPU and returns the count by value (a device synchronisation per scene).  Here the
//     mask stays in a caller-provided workspace, a second kernel performs the greedy pass on the device (64 boxes
//     resolved per step from the diagonal words, the surviving rows OR-ed into the running "removed" words by the
//     whole CTA) and writes keep indices + count to device memory: no allocation, no sync, CUDA-graph capturable,
//     and batched (blockIdx.y = scene, per-scene box counts read from device memory);
//   * only the upper triangle of the mask is computed (the reference also fills the tiles below the diagonal, which
//     its host loop never reads, iou3d_nms.cpp:123-126);
//   * per-box quantities the reference recomputes for every PAIR (rotated corners, cos/sin of +-heading: 20 trig calls
//     per pair, iou3d_nms_kernel.cu:52-53,141-142) are computed once per box per tile and staged in shared memory;
//   * a conservative bounding-circle test skips the polygon clipping for pairs that cannot touch (their overlap is
//     exactly 0 in the reference as well: no edge crossing and no corner within the 1e-2 margin).
//
// Numerics: the polygon clipping keeps the reference's sequence of fp32 operations (same expressions, same libm
// calls cosf/sinf/atan2f, nvcc's default FMA contraction on both sides), so IoU values agree to the last bits and the
// keep lists are identical on the parity tests (tests/test_gpu_det.py compares against the rebuilt reference kernels).
#include "common.cuh"

namespace spsk {

constexpr float IOU_EPS = 1e-8f;     // iou3d_nms_kernel.cu:14
constexpr float IOU_MARGIN = 1e-2f;  // iou3d_nms_kernel.cu:50
constexpr int MAXP = 16;             // iou3d_nms_kernel.cu:136 (cross_points[16])

struct BoxP {             // 16 floats (+1 pad): everything box_overlap needs that depends on ONE box only
    float cx, cy;         // centre
    float hxm, hym;       // dx/2 + MARGIN, dy/2 + MARGIN (check_in_box2d)
    float cn, sn;         // cos(-heading), sin(-heading) (check_in_box2d)
    float area;           // dx*dy
    float rad;            // bounding-circle radius (+ margin slack), for the early-out only
    float px[4], py[4];   // rotated corners
    float pad;            // 17-float stride: consecutive boxes fall in distinct shared-memory banks
};

__device__ __forceinline__ void box_prep(const float *__restrict__ b, BoxP &o) {
    const float x = b[0], y = b[1], dx = b[3], dy = b[4], ang = b[6];
    const float hx = dx / 2, hy = dy / 2;
    const float x1 = x - hx, y1 = y - hy, x2 = x + hx, y2 = y + hy;
    const float c = cosf(ang), s = sinf(ang);
    const float qx[4] = {x1, x2, x2, x1};
    const float qy[4] = {y1, y1, y2, y2};
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // rotate_around_center, iou3d_nms_kernel.cu:98-102
        const float nx = (qx[k] - x) * c + (qy[k] - y) * (-s) + x;
        const float ny = (qx[k] - x) * s + (qy[k] - y) * c + y;
        o.px[k] = nx;
        o.py[k] = ny;
    }
    o.cx = x;
    o.cy = y;
    o.hxm = dx / 2 + IOU_MARGIN;
    o.hym = dy / 2 + IOU_MARGIN;
    o.cn = cosf(-ang);
    o.sn = sinf(-ang);
    o.area = dx * dy;
    o.rad = 0.5f * sqrtf(dx * dx + dy * dy) + 0.0625f;
}

__device__ __forceinline__ float cross3(float p1x, float p1y, float p2x, float p2y, float p0x, float p0y) {
    return (p1x - p0x) * (p2y - p0y) - (p2x - p0x) * (p1y - p0y);  // iou3d_nms_kernel.cu:39-41
}

__device__ __forceinline__ bool in_box(const BoxP &b, float px, float py) {  // iou3d_nms_kernel.cu:48-58
    const float rot_x = (px - b.cx) * b.cn + (py - b.cy) * (-b.sn);
    const float rot_y = (px - b.cx) * b.sn + (py - b.cy) * b.cn;
    return fabsf(rot_x) < b.hxm && fabsf(rot_y) < b.hym;
}

// segment p0->p1 against q0->q1 (iou3d_nms_kernel.cu:60-96)
__device__ __forceinline__ bool seg_intersect(float p1x, float p1y, float p0x, float p0y, float q1x, float q1y,
                                              float q0x, float q0y, float &ox, float &oy) {
    const bool rect = fminf(p0x, p1x) <= fmaxf(q0x, q1x) && fminf(q0x, q1x) <= fmaxf(p0x, p1x) &&
                      fminf(p0y, p1y) <= fmaxf(q0y, q1y) && fminf(q0y, q1y) <= fmaxf(p0y, p1y);
    if (!rect) return false;
    const float s1 = cross3(q0x, q0y, p1x, p1y, p0x, p0y);
    const float s2 = cross3(p1x, p1y, q1x, q1y, p0x, p0y);
    const float s3 = cross3(p0x, p0y, q1x, q1y, q0x, q0y);
    const float s4 = cross3(q1x, q1y, p1x, p1y, q0x, q0y);
    if (!(s1 * s2 > 0 && s3 * s4 > 0)) return false;
    const float s5 = cross3(q1x, q1y, p1x, p1y, p0x, p0y);
    if (fabsf(s5 - s1) > IOU_EPS) {
        ox = (s5 * q0x - s1 * q1x) / (s5 - s1);
        oy = (s5 * q0y - s1 * q1y) / (s5 - s1);
    } else {
        const float a0 = p0y - p1y, b0 = p1x - p0x, c0 = p0x * p1y - p1x * p0y;
        const float a1 = q0y - q1y, b1 = q1x - q0x, c1 = q0x * q1y - q1x * q0y;
        const float D = a0 * b1 - a1 * b0;
        ox = (b0 * c1 - b1 * c0) / D;
        oy = (a1 * c0 - a0 * c1) / D;
    }
    return true;
}

// box_overlap (iou3d_nms_kernel.cu:108-216) on prepared boxes
__device__ float overlap_bev(const BoxP &a, const BoxP &b) {
    {   // bounding circles apart (with slack for the 1e-2 corner margin): the reference finds no point, area 0
        const float ddx = a.cx - b.cx, ddy = a.cy - b.cy, r = a.rad + b.rad;
        if (ddx * ddx + ddy * ddy > r * r) return 0.0f;
    }
    float ptx[MAXP], pty[MAXP], pang[MAXP];
    int cnt = 0;
    float sx = 0.0f, sy = 0.0f;
    // rolled on purpose: one copy of the segment test instead of 16 keeps the kernel inside the instruction cache (ncu
    // showed no-instruction stalls on the unrolled version); the corners are read from shared memory with dynamic indices
#pragma unroll 1
    for (int i = 0; i < 4; ++i) {
        const float a1x = a.px[(i + 1) & 3], a1y = a.py[(i + 1) & 3], a0x = a.px[i], a0y = a.py[i];
#pragma unroll 1
        for (int j = 0; j < 4; ++j) {
            float ox, oy;
            if (seg_intersect(a1x, a1y, a0x, a0y, b.px[(j + 1) & 3], b.py[(j + 1) & 3], b.px[j], b.py[j], ox, oy)) {
                if (cnt < MAXP) {
                    sx = sx + ox;
                    sy = sy + oy;
                    ptx[cnt] = ox;
                    pty[cnt] = oy;
                    ++cnt;
                }
            }
        }
    }
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        if (in_box(a, b.px[k], b.py[k]) && cnt < MAXP) {
            sx = sx + b.px[k];
            sy = sy + b.py[k];
            ptx[cnt] = b.px[k];
            pty[cnt] = b.py[k];
            ++cnt;
        }
        if (in_box(b, a.px[k], a.py[k]) && cnt < MAXP) {
            sx = sx + a.px[k];
            sy = sy + a.py[k];
            ptx[cnt] = a.px[k];
            pty[cnt] = a.py[k];
            ++cnt;
        }
    }
    if (cnt < 3) return 0.0f;  // the fan below is empty (cnt <= 1) or degenerate (cnt == 2: cross(0, v) = 0)
    sx /= cnt;
    sy /= cnt;
    // The reference bubble-sorts with point_cmp = atan2(a - c) > atan2(b - c) (strict, hence stable): any stable
    // sort by that angle yields the same order.  Angles once per point, insertion sort.
    for (int k = 0; k < cnt; ++k) pang[k] = atan2f(pty[k] - sy, ptx[k] - sx);
    for (int k = 1; k < cnt; ++k) {
        const float ka = pang[k], kx = ptx[k], ky = pty[k];
        int m = k - 1;
        while (m >= 0 && pang[m] > ka) {
            pang[m + 1] = pang[m];
            ptx[m + 1] = ptx[m];
            pty[m + 1] = pty[m];
            --m;
        }
        pang[m + 1] = ka;
        ptx[m + 1] = kx;
        pty[m + 1] = ky;
    }
    float area = 0.0f;
    const float x0 = ptx[0], y0 = pty[0];
    for (int k = 0; k < cnt - 1; ++k) {
        const float ax = ptx[k] - x0, ay = pty[k] - y0, bx = ptx[k + 1] - x0, by = pty[k + 1] - y0;
        area += ax * by - ay * bx;
    }
    return fabsf(area) / 2.0f;
}

__device__ __forceinline__ float iou_bev_p(const BoxP &a, const BoxP &b) {  // iou3d_nms_kernel.cu:218-225
    const float s_overlap = overlap_bev(a, b);
    return s_overlap / fmaxf(a.area + b.area - s_overlap, IOU_EPS);
}

__device__ __forceinline__ float iou_normal_raw(const float *a, const float *b) {  // iou3d_nms_kernel.cu:321-333
    const float left = fmaxf(a[0] - a[3] / 2, b[0] - b[3] / 2), right = fminf(a[0] + a[3] / 2, b[0] + b[3] / 2);
    const float top = fmaxf(a[1] - a[4] / 2, b[1] - b[4] / 2), bottom = fminf(a[1] + a[4] / 2, b[1] + b[4] / 2);
    const float width = fmaxf(right - left, 0.f), height = fmaxf(bottom - top, 0.f);
    const float interS = width * height;
    const float Sa = a[3] * a[4];
    const float Sb = b[3] * b[4];
    return interS / fmaxf(Sa + Sb - interS, IOU_EPS);
}

// ---- pairwise matrices ---------------------------------------------------------------------------------------
// mode 0: BEV overlap area, 1: BEV IoU, 2: 3-D IoU (iou3d_nms_utils.py:59-81, torch ops rounded one by one).
constexpr int MT = 16;
__global__ void __launch_bounds__(MT *MT)
boxes_matrix_kernel(int num_a, const float *__restrict__ boxes_a, int num_b, const float *__restrict__ boxes_b,
                    float *__restrict__ out, int mode) {
    __shared__ BoxP sa[MT], sb[MT];
    const int tx = threadIdx.x & (MT - 1), ty = threadIdx.x / MT;
    const int a0 = blockIdx.y * MT, b0 = blockIdx.x * MT;
    if (threadIdx.x < MT) {
        if (a0 + threadIdx.x < num_a) box_prep(boxes_a + (size_t)(a0 + threadIdx.x) * 7, sa[threadIdx.x]);
    } else if (threadIdx.x < 2 * MT) {
        const int t = threadIdx.x - MT;
        if (b0 + t < num_b) box_prep(boxes_b + (size_t)(b0 + t) * 7, sb[t]);
    }
    __syncthreads();
    const int ai = a0 + ty, bi = b0 + tx;
    if (ai >= num_a || bi >= num_b) return;
    float r;
    if (mode == 0) {
        r = overlap_bev(sa[ty], sb[tx]);
    } else if (mode == 1) {
        r = iou_bev_p(sa[ty], sb[tx]);
    } else {
        const float *A = boxes_a + (size_t)ai * 7, *B = boxes_b + (size_t)bi * 7;
        const float ov = overlap_bev(sa[ty], sb[tx]);
        const float a_max = __fadd_rn(A[2], __fdiv_rn(A[5], 2.0f)), a_min = __fsub_rn(A[2], __fdiv_rn(A[5], 2.0f));
        const float b_max = __fadd_rn(B[2], __fdiv_rn(B[5], 2.0f)), b_min = __fsub_rn(B[2], __fdiv_rn(B[5], 2.0f));
        const float oh = fmaxf(__fsub_rn(fminf(a_max, b_max), fmaxf(a_min, b_min)), 0.0f);
        const float o3 = __fmul_rn(ov, oh);
        const float va = __fmul_rn(__fmul_rn(A[3], A[4]), A[5]), vb = __fmul_rn(__fmul_rn(B[3], B[4]), B[5]);
        r = __fdiv_rn(o3, fmaxf(__fsub_rn(__fadd_rn(va, vb), o3), 1e-6f));
    }
    out[(size_t)ai * num_b + bi] = r;
}

// ---- NMS: suppression bit mask (upper triangle) -------------------------------------------------------------
// One CTA = one 32 x 64 tile of (row box, column box) pairs of one scene; 256 threads: 8 lanes per row, lane q takes
// columns q, q+8, ... (conflict-free shared reads of consecutive prepared boxes) and the 8 partial words are OR-ed with
// three shuffles.  Bit j of mask[row][col_blk] = IoU(row, 64*col_blk + j) > thresh, only for columns after the row.
constexpr int NT = 64;   // columns per tile = bits per mask word
constexpr int NR = 32;   // rows per tile
template <bool NORMAL>
__global__ void __launch_bounds__(256)
nms_mask_kernel(int n, const float *__restrict__ boxes, const int *__restrict__ counts, float thresh,
                unsigned long long *__restrict__ mask) {
    const int cb = (n + NT - 1) / NT;
    const int row_blk = blockIdx.x / cb, col_blk = blockIdx.x % cb;
    const int row0 = row_blk * NR, col0 = col_blk * NT;
    if (col0 + NT - 1 <= row0) return;   // every column of the tile precedes (or is) every row: never read
    const int scene = blockIdx.y;
    const int nb = counts ? min(counts[scene], n) : n;
    if (row0 >= nb || col0 >= nb) return;
    const float *bx = boxes + (size_t)scene * n * 7;
    unsigned long long *mk = mask + (size_t)scene * n * cb;
    const int row_size = min(nb - row0, NR), col_size = min(nb - col0, NT);

    __shared__ float raw[NORMAL ? (NR + NT) * 7 : 1];
    __shared__ BoxP prep[NORMAL ? 1 : NR + NT];
    if (NORMAL) {
        for (int i = threadIdx.x; i < (NR + NT) * 7; i += 256) {
            const int bi = i / 7, c = i - bi * 7;
            const bool is_col = bi >= NR;
            const int local = is_col ? bi - NR : bi, lim = is_col ? col_size : row_size, base = is_col ? col0 : row0;
            raw[i] = local < lim ? bx[(size_t)(base + local) * 7 + c] : 0.0f;
        }
    } else if (threadIdx.x < NR + NT) {
        const bool is_col = threadIdx.x >= NR;
        const int local = is_col ? threadIdx.x - NR : threadIdx.x, lim = is_col ? col_size : row_size, base = is_col ? col0 : row0;
        if (local < lim) box_prep(bx + (size_t)(base + local) * 7, prep[threadIdx.x]);
    }
    __syncthreads();
    const int r = threadIdx.x >> 3, q = threadIdx.x & 7;
    unsigned long long t = 0ull;
    if (r < row_size) {
        const int gr = row0 + r;
        if (NORMAL) {
            for (int c = q; c < col_size; c += 8)
                if (col0 + c > gr && iou_normal_raw(raw + r * 7, raw + (NR + c) * 7) > thresh) t |= 1ull << c;
        } else {
            const BoxP &a = prep[r];
            for (int c = q; c < col_size; c += 8)
                if (col0 + c > gr && iou_bev_p(a, prep[NR + c]) > thresh) t |= 1ull << c;
        }
    }
    t |= __shfl_xor_sync(0xffffffffu, t, 1);
    t |= __shfl_xor_sync(0xffffffffu, t, 2);
    t |= __shfl_xor_sync(0xffffffffu, t, 4);
    if (q == 0 && r < row_size) mk[(size_t)(row0 + r) * cb + col_blk] = t;
}

// Small scenes (n <= 1024; IA-SSD post-processing has 256 boxes per scene): the kernel above is bound by the LATENCY of its
// longest thread (8 polygon clippings in sequence at low occupancy; ncu: 74 us for 16 x 256 boxes at 14 % warps active).
// Here every thread owns exactly ONE pair: a tile is 4 rows x 64 columns, warp w handles row w/2, columns 32 (w%2) + lane,
// and the mask word is assembled from two ballots.
constexpr int FR = 4;
template <bool NORMAL>
__global__ void __launch_bounds__(256)
nms_mask_fine_kernel(int n, const float *__restrict__ boxes, const int *__restrict__ counts, float thresh,
                     unsigned long long *__restrict__ mask) {
    const int cb = (n + NT - 1) / NT;
    const int row_blk = blockIdx.x / cb, col_blk = blockIdx.x % cb;
    const int row0 = row_blk * FR, col0 = col_blk * NT;
    if (col0 + NT - 1 <= row0) return;
    const int scene = blockIdx.y;
    const int nb = counts ? min(counts[scene], n) : n;
    if (row0 >= nb || col0 >= nb) return;
    const float *bx = boxes + (size_t)scene * n * 7;
    unsigned long long *mk = mask + (size_t)scene * n * cb;
    const int row_size = min(nb - row0, FR), col_size = min(nb - col0, NT);
    __shared__ float raw[NORMAL ? (FR + NT) * 7 : 1];
    __shared__ BoxP prep[NORMAL ? 1 : FR + NT];
    __shared__ uint32_t part[FR][2];
    if (NORMAL) {
        for (int i = threadIdx.x; i < (FR + NT) * 7; i += 256) {
            const int bi = i / 7, c = i - bi * 7;
            const bool is_col = bi >= FR;
            const int local = is_col ? bi - FR : bi, lim = is_col ? col_size : row_size, base = is_col ? col0 : row0;
            raw[i] = local < lim ? bx[(size_t)(base + local) * 7 + c] : 0.0f;
        }
    } else if (threadIdx.x < FR + NT) {
        const bool is_col = threadIdx.x >= FR;
        const int local = is_col ? threadIdx.x - FR : threadIdx.x, lim = is_col ? col_size : row_size, base = is_col ? col0 : row0;
        if (local < lim) box_prep(bx + (size_t)(base + local) * 7, prep[threadIdx.x]);
    }
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = w >> 1, c = (w & 1) * 32 + lane;
    bool hit = false;
    if (r < row_size && c < col_size && col0 + c > row0 + r) {
        if (NORMAL)
            hit = iou_normal_raw(raw + r * 7, raw + (FR + c) * 7) > thresh;
        else
            hit = iou_bev_p(prep[r], prep[FR + c]) > thresh;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, hit);
    if (lane == 0) part[r][w & 1] = m;
    __syncthreads();
    if (threadIdx.x < row_size)
        mk[(size_t)(row0 + threadIdx.x) * cb + col_blk] = (unsigned long long)part[threadIdx.x][0] | ((unsigned long long)part[threadIdx.x][1] << 32);
}

// ---- NMS: greedy pass on the device --------------------------------------------------------------------------
// Optional emission of the final detections (spsk_detect_postprocess): rank r < post_max of the keep list maps through
// order[] to the original centre of the scene and copies its box / score / label into padded outputs.
struct NmsEmit {
    const int *order;        // (batch, n) sorted position -> original row within the scene (NULL = no emission)
    const float *box_preds;  // (batch*m, 7)
    const float *scores;     // (batch*m)
    const int *labels;       // (batch*m)
    int m;                   // rows per scene in box_preds / scores / labels
    int post_max;
    float *out_boxes;        // (batch, post_max, 7)
    float *out_scores;       // (batch, post_max)
    long long *out_labels;   // (batch, post_max)
    long long *out_index;    // (batch, post_max) original row within the scene
    int *out_count;          // (batch)
};

// survivor `srt` (position in the sorted order) is the rank-th kept box of its scene
__device__ __forceinline__ void emit_kept(const NmsEmit &em, long long *kp, int scene, int n, int srt, int rank) {
    if (kp) kp[rank] = srt;
    if (em.order && rank < em.post_max) {
        const int orig = em.order[(size_t)scene * n + srt];
        const size_t row = (size_t)scene * em.m + orig, o = (size_t)scene * em.post_max + rank;
#pragma unroll
        for (int c = 0; c < 7; ++c) em.out_boxes[o * 7 + c] = em.box_preds[row * 7 + c];
        em.out_scores[o] = em.scores[row];
        em.out_labels[o] = em.labels[row];
        em.out_index[o] = orig;
    }
}

__device__ __forceinline__ void emit_tail(const NmsEmit &em, int *num_keep, int scene, int total, int nthreads) {
    if (threadIdx.x == 0 && num_keep) num_keep[scene] = total;
    if (em.order) {
        const int cnt = min(total, em.post_max);
        if (threadIdx.x == 0) em.out_count[scene] = cnt;
        for (int i = cnt * 7 + threadIdx.x; i < em.post_max * 7; i += nthreads) em.out_boxes[(size_t)scene * em.post_max * 7 + i] = 0.0f;
        for (int i = cnt + threadIdx.x; i < em.post_max; i += nthreads) {
            const size_t o = (size_t)scene * em.post_max + i;
            em.out_scores[o] = 0.0f;
            em.out_labels[o] = 0;
            em.out_index[o] = -1;
        }
    }
}

// Scenes of at most 1024 boxes (IA-SSD: 256 centres on KITTI, 1024 on Waymo): the whole mask fits in shared memory, one
// coalesced load, then ONE WARP walks the rows in order with lane j owning "removed" word j: per row a shuffle to fetch
// the owner's word, a bit test, and for survivors one shared-memory OR per lane.  No global latency inside the
// sequential part (the block-wise kernel below pays ~1 us of L2 latency per 64 rows).
constexpr int RS_MAXN = 1024, RS_MAXW = RS_MAXN / NT;
__global__ void __launch_bounds__(256)
nms_reduce_small_kernel(int n, const int *__restrict__ counts, const unsigned long long *__restrict__ mask,
                        long long *__restrict__ keep, int *__restrict__ num_keep, NmsEmit em) {
    extern __shared__ unsigned long long sm[];   // n x cb words (8 KB at n = 256, 128 KB at n = 1024)
    __shared__ unsigned long long skb[RS_MAXW];
    __shared__ int spre[RS_MAXW + 1];
    const int cb = (n + NT - 1) / NT;            // <= RS_MAXW <= 32: one "removed" word per lane
    const int scene = blockIdx.x;
    const int nb = counts ? min(counts[scene], n) : n;
    const int cbb = (nb + NT - 1) / NT;
    const unsigned long long *mk = mask + (size_t)scene * n * cb;
    for (int i = threadIdx.x; i < nb * cb; i += 256) {
        const int row = i / cb, j = i - row * cb;
        sm[i] = (j >= (row >> 6) && j < cbb) ? mk[i] : 0ull;   // words below the diagonal are never written
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        unsigned long long remv = 0ull, kb = 0ull;
        for (int t = 0; t < nb; ++t) {
            const int blk = t >> 6, bit = t & 63;
            const unsigned long long cur = __shfl_sync(0xffffffffu, remv, blk);
            if (!((cur >> bit) & 1ull)) {
                if (lane == blk) kb |= 1ull << bit;
                if (lane < cb) remv |= sm[t * cb + lane];
            }
        }
        if (lane < RS_MAXW) skb[lane] = lane < cb ? kb : 0ull;
        int c = lane < cb ? __popcll(kb) : 0, incl = c;
#pragma unroll
        for (int o = 1; o < RS_MAXW; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane < RS_MAXW) spre[lane + 1] = incl;
        if (lane == 0) spre[0] = 0;
    }
    __syncthreads();
    long long *kp = keep ? keep + (size_t)scene * n : nullptr;
    for (int t = threadIdx.x; t < nb; t += 256) {
        const unsigned long long kb = skb[t >> 6];
        if ((kb >> (t & 63)) & 1ull) emit_kept(em, kp, scene, n, t, spre[t >> 6] + __popcll(kb & ((1ull << (t & 63)) - 1ull)));
    }
    emit_tail(em, num_keep, scene, spre[RS_MAXW], 256);
}

constexpr int RT = 256;
__global__ void __launch_bounds__(RT)
nms_reduce_kernel(int n, const int *__restrict__ counts, const unsigned long long *__restrict__ mask,
                  long long *__restrict__ keep, int *__restrict__ num_keep, NmsEmit em) {
    extern __shared__ unsigned long long remv[];  // cb words
    __shared__ unsigned long long diag[NT];
    __shared__ unsigned long long kept_bits;
    const int cb = (n + NT - 1) / NT;
    const int scene = blockIdx.x;
    const int nb = counts ? min(counts[scene], n) : n;
    const int cbb = (nb + NT - 1) / NT;
    const unsigned long long *mk = mask + (size_t)scene * n * cb;
    long long *kp = keep ? keep + (size_t)scene * n : nullptr;
    for (int j = threadIdx.x; j < cbb; j += RT) remv[j] = 0ull;
    __syncthreads();
    int total = 0;
    for (int blk = 0; blk < cbb; ++blk) {
        const int rows = min(nb - blk * NT, NT);
        if (threadIdx.x < NT) diag[threadIdx.x] = threadIdx.x < rows ? mk[(size_t)(blk * NT + threadIdx.x) * cb + blk] : 0ull;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long cur = remv[blk], kb = 0ull;
            for (int t = 0; t < rows; ++t) {
                if (!((cur >> t) & 1ull)) {
                    kb |= 1ull << t;
                    cur |= diag[t];
                }
            }
            kept_bits = kb;
        }
        __syncthreads();
        const unsigned long long kb = kept_bits;
        for (int j = blk + 1 + threadIdx.x; j < cbb; j += RT) {
            unsigned long long acc = remv[j], rest = kb;
            while (rest) {
                const int t = __ffsll((long long)rest) - 1;
                rest &= rest - 1;
                acc |= mk[(size_t)(blk * NT + t) * cb + j];
            }
            remv[j] = acc;
        }
        if (threadIdx.x < NT && ((kb >> threadIdx.x) & 1ull))
            emit_kept(em, kp, scene, n, blk * NT + threadIdx.x, total + __popcll(kb & ((1ull << threadIdx.x) - 1ull)));
        total += __popcll(kb);
        __syncthreads();
    }
    emit_tail(em, num_keep, scene, total, RT);
}

static int nms_launch(int batch, int n, const float *boxes, const int *counts, float thresh, int normal,
                      long long *keep, int *num_keep, void *workspace, long long ws_bytes, const NmsEmit &em,
                      cudaStream_t st) {
    const int cb = (n + NT - 1) / NT;
    const long long need = (long long)batch * n * cb * 8;
    SPSK_REQUIRE(workspace && ws_bytes >= need, SPSK_ERR_WORKSPACE, "nms: workspace %lld < %lld bytes", ws_bytes, need);
    SPSK_REQUIRE(n <= SPSK_NMS_MAX_N, SPSK_ERR_UNSUPPORTED, "nms: n=%d > %d", n, SPSK_NMS_MAX_N);
    SPSK_REQUIRE((long long)((n + NR - 1) / NR) * cb <= 0x7fffffffLL && batch <= 65535, SPSK_ERR_UNSUPPORTED, "nms: grid too large");
    unsigned long long *mask = static_cast<unsigned long long *>(workspace);
    if (n <= 1024) {   // latency-bound regime: one pair per thread
        dim3 grid(((n + FR - 1) / FR) * cb, batch);
        if (normal)
            nms_mask_fine_kernel<true><<<grid, 256, 0, st>>>(n, boxes, counts, thresh, mask);
        else
            nms_mask_fine_kernel<false><<<grid, 256, 0, st>>>(n, boxes, counts, thresh, mask);
    } else {
        dim3 grid(((n + NR - 1) / NR) * cb, batch);
        if (normal)
            nms_mask_kernel<true><<<grid, 256, 0, st>>>(n, boxes, counts, thresh, mask);
        else
            nms_mask_kernel<false><<<grid, 256, 0, st>>>(n, boxes, counts, thresh, mask);
    }
    SPSK_LAUNCH_CHECK("nms_mask_kernel");
    if (n <= RS_MAXN) {
        const int smem = n * cb * 8;
        static SmemAttrOnce attr;
        if (smem > 48 * 1024)
            if (int rc = attr.ensure((const void *)nms_reduce_small_kernel, RS_MAXN * RS_MAXW * 8, "cudaFuncSetAttribute(nms_reduce_small_kernel)")) return rc;
        nms_reduce_small_kernel<<<batch, 256, smem, st>>>(n, counts, mask, keep, num_keep, em);
    }
    else
        nms_reduce_kernel<<<batch, RT, cb * 8, st>>>(n, counts, mask, keep, num_keep, em);
    SPSK_LAUNCH_CHECK("nms_reduce_kernel");
    return SPSK_OK;
}

// ---- fused head decode + score sort (spsk_detect_postprocess, stage 1) ----------------------------------------
// One CTA per scene.  Per centre: label = argmax_c logits (first maximum), score = sigmoid(max logit)
// (detector3d_template.py:225,259), box = PointResidual_BinOri_Coder.decode_torch (box_coder_utils.py:279-319) with
// torch's one-rounding-per-op arithmetic; then (score >= thresh) centres are ordered by descending score (ties: lower
// row first; model_nms_utils.py:7-16 mask + topk) and their boxes written in that order for the mask kernel.
constexpr int DT = 256;
__global__ void __launch_bounds__(DT)
detect_sort_kernel(spsk_detect_desc d, int m_pad, float *__restrict__ sorted_boxes, int *__restrict__ order,
                   int *__restrict__ nvalid) {
    extern __shared__ unsigned long long keys[];
    __shared__ int n_ok;
    const int scene = blockIdx.x;
    if (threadIdx.x == 0) n_ok = 0;
    __syncthreads();
    const float bin_inter = (float)(2.0 * 3.14159265358979323846 / d.bin_size);
    const float bin_half = (float)(2.0 * 3.14159265358979323846 / d.bin_size / 2.0);
    const float neg_pi = (float)3.14159265358979323846;
    int local_ok = 0;
    for (int i = threadIdx.x; i < m_pad; i += DT) {
        unsigned long long key = 0ull;
        if (i < d.m) {
            const size_t row = (size_t)scene * d.m + i;
            float score = 0.0f;
            int lab;
            if (d.cls) {
                const float *cl = d.cls + row * d.ld_cls;
                float mx = cl[0];
                lab = 0;
                for (int c = 1; c < d.num_class; ++c) {
                    const float v = cl[c];
                    if (v > mx) { mx = v; lab = c; }
                }
                score = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-mx)));
                d.scores[row] = score;
                d.labels[row] = lab + 1;
            } else {  // decode only, classes given (box_coder.decode_torch(..., pred_classes))
                lab = min(max(d.labels[row] - 1, 0), d.num_class - 1);
            }
            if (d.reg) {
                const float *rg = d.reg + row * d.ld_reg;
                const float *ct = d.centers + row * d.ld_centers;
                float bxv[7];
                if (d.mean_size) {
                    const float dxa = d.mean_size[lab * 3 + 0], dya = d.mean_size[lab * 3 + 1], dza = d.mean_size[lab * 3 + 2];
                    const float diag = __fsqrt_rn(__fadd_rn(__fmul_rn(dxa, dxa), __fmul_rn(dya, dya)));
                    bxv[0] = __fadd_rn(__fmul_rn(rg[0], diag), ct[0]);
                    bxv[1] = __fadd_rn(__fmul_rn(rg[1], diag), ct[1]);
                    bxv[2] = __fadd_rn(__fmul_rn(rg[2], dza), ct[2]);
                    bxv[3] = __fmul_rn(expf(rg[3]), dxa);
                    bxv[4] = __fmul_rn(expf(rg[4]), dya);
                    bxv[5] = __fmul_rn(expf(rg[5]), dza);
                } else {
                    bxv[0] = __fadd_rn(rg[0], ct[0]);
                    bxv[1] = __fadd_rn(rg[1], ct[1]);
                    bxv[2] = __fadd_rn(rg[2], ct[2]);
                    bxv[3] = expf(rg[3]);
                    bxv[4] = expf(rg[4]);
                    bxv[5] = expf(rg[5]);
                }
                int bin = 0;
                float bmx = rg[6];
                for (int k = 1; k < d.bin_size; ++k) {
                    const float v = rg[6 + k];
                    if (v > bmx) { bmx = v; bin = k; }
                }
                const float res = __fadd_rn(rg[6 + d.bin_size + bin], 0.0f);  // sum(bin_res * one_hot)
                float ang = __fadd_rn(__fsub_rn(__fmul_rn((float)bin, bin_inter), neg_pi), bin_half);
                ang = __fadd_rn(ang, __fmul_rn(res, bin_half));
                bxv[6] = ang;
#pragma unroll
                for (int c = 0; c < 7; ++c) d.box_preds[row * 7 + c] = bxv[c];
            }
            if (score >= d.score_thresh) {
                ++local_ok;
                key = ((unsigned long long)__float_as_uint(score) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
            }
        }
        if (sorted_boxes) keys[i] = key;
    }
    if (!sorted_boxes) return;  // decode-only call (uniform for the grid)
    if (local_ok) atomicAdd(&n_ok, local_ok);
    __syncthreads();   // also orders this CTA's box_preds stores before the sorted copy below
    for (int k = 2; k <= m_pad; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < m_pad; i += DT) {
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
    const int nv = min(n_ok, d.pre_max);
    if (threadIdx.x == 0) nvalid[scene] = nv;
    for (int r = threadIdx.x; r < d.m; r += DT) {
        int orig = -1;
        if (r < nv) {
            orig = (int)(0xFFFFFFFFu - (uint32_t)(keys[r] & 0xFFFFFFFFull));
            const size_t row = (size_t)scene * d.m + orig;
#pragma unroll
            for (int c = 0; c < 7; ++c) sorted_boxes[((size_t)scene * d.m + r) * 7 + c] = d.box_preds[row * 7 + c];
        }
        order[(size_t)scene * d.m + r] = orig;
    }
}

}  // namespace spsk

using namespace spsk;

static int matrix_call(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *out, int mode,
                       spsk_stream_t stream, const char *what) {
    SPSK_REQUIRE(num_a >= 0 && num_b >= 0, SPSK_ERR_INVALID_ARG, "%s: negative box count", what);
    if (num_a == 0 || num_b == 0) return SPSK_OK;
    SPSK_REQUIRE(boxes_a && boxes_b && out, SPSK_ERR_INVALID_ARG, "%s: null pointer", what);
    dim3 grid((num_b + MT - 1) / MT, (num_a + MT - 1) / MT);
    SPSK_REQUIRE(grid.y <= 65535, SPSK_ERR_UNSUPPORTED, "%s: num_a=%d too large", what, num_a);
    boxes_matrix_kernel<<<grid, MT * MT, 0, as_stream(stream)>>>(num_a, boxes_a, num_b, boxes_b, out, mode);
    SPSK_LAUNCH_CHECK(what);
    return SPSK_OK;
}

extern "C" {

SPSK_API int spsk_boxes_overlap_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b,
                                    float *ans_overlap, spsk_stream_t stream) {
    return matrix_call(num_a, boxes_a, num_b, boxes_b, ans_overlap, 0, stream, "spsk_boxes_overlap_bev");
}

SPSK_API int spsk_boxes_iou_bev(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans_iou,
                                spsk_stream_t stream) {
    return matrix_call(num_a, boxes_a, num_b, boxes_b, ans_iou, 1, stream, "spsk_boxes_iou_bev");
}

SPSK_API int spsk_boxes_iou3d(int num_a, const float *boxes_a, int num_b, const float *boxes_b, float *ans_iou,
                              spsk_stream_t stream) {
    return matrix_call(num_a, boxes_a, num_b, boxes_b, ans_iou, 2, stream, "spsk_boxes_iou3d");
}

SPSK_API long long spsk_nms_workspace_bytes(int batch, int n) {
    if (batch < 0 || n < 0) return 0;
    const long long cb = (n + NT - 1) / NT;
    return (long long)batch * n * cb * 8;
}

SPSK_API int spsk_nms(int batch, int n, const float *boxes, const int *counts, float thresh, int normal,
                      long long *keep, int *num_keep, void *workspace, long long workspace_bytes,
                      spsk_stream_t stream) {
    SPSK_REQUIRE(batch >= 0 && n >= 0, SPSK_ERR_INVALID_ARG, "spsk_nms: negative size");
    SPSK_REQUIRE(num_keep, SPSK_ERR_INVALID_ARG, "spsk_nms: null num_keep");
    if (batch == 0) return SPSK_OK;
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(num_keep, 0, sizeof(int) * batch, as_stream(stream));
        if (e != cudaSuccess) return cuda_fail(e, "spsk_nms memset");
        return SPSK_OK;
    }
    SPSK_REQUIRE(boxes && keep, SPSK_ERR_INVALID_ARG, "spsk_nms: null pointer");
    NmsEmit em = {};
    return nms_launch(batch, n, boxes, counts, thresh, normal, keep, num_keep, workspace, workspace_bytes, em,
                      as_stream(stream));
}

SPSK_API long long spsk_detect_workspace_bytes(int batch, int m) {
    if (batch < 0 || m < 0) return 0;
    // sorted boxes (batch,m,7) f32 | order (batch,m) i32 | nvalid (batch) i32 (padded to 8) | mask
    const long long a = (long long)batch * m * 7 * 4, b = (long long)batch * m * 4, c = (((long long)batch * 4 + 7) / 8) * 8;
    return ((a + b + 7) / 8) * 8 + c + spsk_nms_workspace_bytes(batch, m);
}

SPSK_API int spsk_detect_postprocess(const spsk_detect_desc *dp, spsk_stream_t stream) {
    SPSK_REQUIRE(dp, SPSK_ERR_INVALID_ARG, "spsk_detect_postprocess: null descriptor");
    const spsk_detect_desc d = *dp;
    SPSK_REQUIRE(d.batch >= 0 && d.m > 0 && d.num_class > 0 && d.bin_size > 0, SPSK_ERR_INVALID_ARG,
                 "spsk_detect_postprocess: bad sizes b=%d m=%d classes=%d bins=%d", d.batch, d.m, d.num_class, d.bin_size);
    SPSK_REQUIRE(d.m <= SPSK_DETECT_MAX_M, SPSK_ERR_UNSUPPORTED, "spsk_detect_postprocess: m=%d > %d", d.m, SPSK_DETECT_MAX_M);
    SPSK_REQUIRE(d.box_preds && d.labels && (d.cls || d.reg), SPSK_ERR_INVALID_ARG,
                 "spsk_detect_postprocess: null box_preds / labels, or neither cls nor reg given");
    SPSK_REQUIRE(!d.cls || (d.scores && d.ld_cls >= d.num_class), SPSK_ERR_INVALID_ARG,
                 "spsk_detect_postprocess: cls given but scores null or ld_cls < num_class");
    SPSK_REQUIRE(!d.reg || (d.centers && d.ld_reg >= 6 + 2 * d.bin_size && d.ld_centers >= 3), SPSK_ERR_INVALID_ARG,
                 "spsk_detect_postprocess: reg given but centers null or row strides too small");
    const bool decode_only = d.out_boxes == nullptr;
    SPSK_REQUIRE(decode_only || d.cls, SPSK_ERR_INVALID_ARG, "spsk_detect_postprocess: NMS needs class logits");
    SPSK_REQUIRE(decode_only || (d.post_max > 0 && d.pre_max > 0), SPSK_ERR_INVALID_ARG,
                 "spsk_detect_postprocess: pre/post max must be > 0");
    SPSK_REQUIRE(decode_only || (d.out_scores && d.out_labels && d.out_index && d.out_count), SPSK_ERR_INVALID_ARG,
                 "spsk_detect_postprocess: null output pointer");
    if (d.batch == 0) return SPSK_OK;
    if (decode_only) {
        detect_sort_kernel<<<d.batch, DT, 16, as_stream(stream)>>>(d, d.m, nullptr, nullptr, nullptr);
        SPSK_LAUNCH_CHECK("detect_sort_kernel(decode)");
        return SPSK_OK;
    }
    const long long need = spsk_detect_workspace_bytes(d.batch, d.m);
    SPSK_REQUIRE(d.workspace && d.workspace_bytes >= need, SPSK_ERR_WORKSPACE,
                 "spsk_detect_postprocess: workspace %lld < %lld bytes", d.workspace_bytes, need);
    char *ws = static_cast<char *>(d.workspace);
    const long long a = (long long)d.batch * d.m * 7 * 4, b = (long long)d.batch * d.m * 4;
    float *sorted_boxes = reinterpret_cast<float *>(ws);
    int *order = reinterpret_cast<int *>(ws + a);
    const long long off_nv = ((a + b + 7) / 8) * 8;
    int *nvalid = reinterpret_cast<int *>(ws + off_nv);
    const long long off_mask = off_nv + (((long long)d.batch * 4 + 7) / 8) * 8;
    int m_pad = 2;
    while (m_pad < d.m) m_pad *= 2;
    cudaStream_t st = as_stream(stream);
    detect_sort_kernel<<<d.batch, DT, m_pad * 8, st>>>(d, m_pad, sorted_boxes, order, nvalid);
    SPSK_LAUNCH_CHECK("detect_sort_kernel");
    NmsEmit em;
    em.order = order;
    em.box_preds = d.box_preds;
    em.scores = d.scores;
    em.labels = d.labels;
    em.m = d.m;
    em.post_max = d.post_max;
    em.out_boxes = d.out_boxes;
    em.out_scores = d.out_scores;
    em.out_labels = d.out_labels;
    em.out_index = d.out_index;
    em.out_count = d.out_count;
    return nms_launch(d.batch, d.m, sorted_boxes, nvalid, d.nms_thresh, d.nms_normal, nullptr, nullptr, ws + off_mask,
                      d.workspace_bytes - off_mask, em, st);
}

}  // extern "C"
