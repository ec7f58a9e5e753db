// sa_mma_common.cuh -- types and epilogue helpers shared by the fused set-abstraction kernels
// (sa_mma.cu: one CTA per tile;  sa_mma_pair.cu: a CTA pair per 256-row tile, tcgen05 cta_group::2).
#pragma once
#include "mma_ptx.cuh"

namespace spsk {

constexpr int MM_ROWS = 128;          // grouped rows per tile
constexpr int MM_THREADS = 192;       // G = 1: 4 gather/epilogue warps + producer + mma;  G = 2: 8 + 2 warps = 320 threads
constexpr int MM_STAGE_BYTES = 16384; // largest weight tile: [128 cout][64 k] fp16
constexpr int MM_MAX_LAYERS = 4;
constexpr int MM_MAX_STAGES = 8;
constexpr int MM_HDR = 1024;          // barriers + TMEM slot
constexpr int MM_SCHED_MAX = 104;     // tabulated weight tiles (+ ring padding entries) per row tile, 2 x 16 bytes each in the kernel parameters
// flags of SaArgs::sched[e].w (bits 0..3 = MMAs of the entry, K / 16 = 1..8; 0 for a padding entry;  bits 16..26 = descriptor
// offset, in 16-byte units, of the entry's SECOND 64-wide k tile inside its ring slot)
constexpr uint32_t SCH_FIRST_KC = 1u << 4;      // first weight tile of a job: take an accumulator
constexpr uint32_t SCH_LAST_KC = 1u << 5;       // last weight tile of a job: commit the accumulator
constexpr uint32_t SCH_LRING_FIRST = 1u << 11;  // first entry of the tile on the last layer's overlay ring: the producer waits for HID_DONE
constexpr uint32_t SCH_HID_DONE = 1u << 12;     // last MMAs reading the activation buffer the overlay ring lives in
constexpr uint32_t SCH_LAST_LAYER = 1u << 13;   // orientation B (weights are the A operand)
constexpr int MM_MAX_XC = 16;         // 64-wide K chunks per activation buffer (K <= 1024)

struct SaLayer {
    int kpad;      // true input width, multiple of 16
    int cpad;      // output width: multiple of 16 (hidden) / 128 (last)
    int n_cc;      // ceil(cpad / 128)
    int xw;        // activation buffer width in halfs: kpad (plain) or 2*kpad (split: [hi | lo])
    int vk;        // K the MMAs run over: kpad (plain) or 3*kpad (split: Xh.Wh + Xl.Wh + Xh.Wl)
    int wk;        // K of the packed weights: kpad (plain) or 2*kpad (split: [Wh ; Wl], Wh is read by two of the three products)
    int n_kc;      // ceil(wk / 64)  weight tiles per cout chunk
    int n_xc;      // ceil(xw / 64)  readiness chunks of the activation buffer
    int w_off;     // byte offset of this layer's tiles in the packed weights
    int bias_off;  // float offset into the bias array
};

struct SaArgs {
    int nlayers;
    SaLayer L[MM_MAX_LAYERS];
    int b, n, m, nsample, ns_log2;
    int c_feat, cpad8, ldtwin, use_xyz, split;
    long long rows;  // b*m*nsample
    int ntiles;
    int nstages, resident, w_total, tmem_cols, nbuf, nbuf_log2;
    int l0_fused, l0_off;   // split chains: layer 0 (K <= 11 real inputs) is evaluated in fp32 by the gather threads from the
                            // [16][cpad0] fp32 weights + bias appended to the resident weights at byte l0_off; the MMA chain starts at layer 1
    int rot_last;    // rotate the last layer's cout-chunk order by blockIdx (de-synchronises the CTAs' weight streams)
    int sched_n;     // > 0: streaming chain whose per-tile MMA schedule (sched_n weight tiles) is tabulated in shared memory
    int narrow;      // resident chain with one job per layer: the MMA warp runs the register-resident fast loop
    int lstages;     // > 0: the last layer streams its weights through `lstages` extra 16 KB slots overlaid on the activation
                     // buffer that is dead while it runs (the input of layer nlayers-2)
    int xa_bytes, xb_bytes;
    const float *xyz, *new_xyz, *feat32;
    const __half *twin;   // (b, n, ldtwin)
    const int *idx;       // (b, m, nsample)
    const uint8_t *wtiles;
    const float *bias;
    float *out;           // (b, c_total, m) or null
    int c_total, co_off, cout_last;
    __half *out16;        // (b*m, ld16) or null
    int ld16, co16, n16, o16lo;
    unsigned long long *prof;   // optional per-role wait/work cycle counters (spsk_sa_mma_set_profile), null = off
    unsigned int *ovf;          // fp16 range guard word of this device (may be null) and this call's tag bit
    unsigned int ovf_bit;
    int abl;                    // measurement aid, honoured by the PROFILING kernel variants only (SPSK_SA_ABL bit mask; results are then
                                // WRONG by design): 1 = the producer signals weight stages without copying after the first tile,
                                // 2 = hidden epilogues skip their shared-memory stores, 8 = one MMA per weight tile instead of kw/16
    double *stats;              // batch-statistics pass (training-mode BN): per-(CTA, epilogue group) partial sums, see include/spsk.h
    uint4 sched[MM_SCHED_MAX];  // streaming chains: the per-tile MMA schedule (sched_n entries), see sa_mma.cu::build_schedule
    uint4 ring[MM_SCHED_MAX];   // ... and, per entry, the static ring slot / barriers / phase parities / global source of its weight tile
};

// byte offset of weight tile (cc, kc) inside a layer: chunks of 128 couts are contiguous (cc-major), inside a chunk
// the tiles follow each other along K; a tile is ncols x kw fp16 in canonical layout with SBO = kw*16
__device__ __forceinline__ int wtile_off(const SaLayer &Ly, int cc, int kc, int ncols) {
    return Ly.w_off + (128 * cc * Ly.wk + ncols * 64 * kc) * 2;
}

// 16 accumulator columns -> + bias -> ReLU -> fp16 -> two 16-byte stores into the next operand (K-major).
// `mx` collects the largest value seen (fp16 range guard: the caller flags > 65504 once per job; 3-input FMNMX, half an
// instruction per value).
__device__ __forceinline__ void store_hidden16(const float *v, const float *bias16, uint8_t *dst, float &mx) {
    const float4 *b4 = reinterpret_cast<const float4 *>(bias16);
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 bb = __ldg(b4 + i);
        const float s0 = v[4 * i] + bb.x, s1 = v[4 * i + 1] + bb.y, s2 = v[4 * i + 2] + bb.z, s3 = v[4 * i + 3] + bb.w;
        mx = fmaxf(fmaxf(mx, s0), s1);
        mx = fmaxf(fmaxf(mx, s2), s3);
        h[2 * i] = pack_h2_relu(s0, s1);
        h[2 * i + 1] = pack_h2_relu(s2, s3);
    }
    *reinterpret_cast<uint4 *>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(h[4], h[5], h[6], h[7]);
}
// hi/lo split of 8 fp32 values: hi = fp16(x), lo = fp16(x - hi)
__device__ __forceinline__ void split8(const float *y, uint4 &hi, uint4 &lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hh = __floats2half2_rn(y[2 * i], y[2 * i + 1]);
        const float2 hf = __half22float2(hh);
        h[i] = *reinterpret_cast<const uint32_t *>(&hh);
        l[i] = pack_h2(y[2 * i] - hf.x, y[2 * i + 1] - hf.y);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void store_hidden16_split(const float *v, const float *bias16, uint8_t *dst_hi, uint8_t *dst_lo, float &mx) {
    const float4 *b4 = reinterpret_cast<const float4 *>(bias16);
    float y[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 bb = __ldg(b4 + i);
        y[4 * i] = fmaxf(v[4 * i] + bb.x, 0.f); y[4 * i + 1] = fmaxf(v[4 * i + 1] + bb.y, 0.f);
        y[4 * i + 2] = fmaxf(v[4 * i + 2] + bb.z, 0.f); y[4 * i + 3] = fmaxf(v[4 * i + 3] + bb.w, 0.f);
        mx = fmaxf(fmaxf(mx, y[4 * i]), y[4 * i + 1]);
        mx = fmaxf(fmaxf(mx, y[4 * i + 2]), y[4 * i + 3]);
    }
    uint4 h0, l0, h1, l1;
    split8(y, h0, l0);
    split8(y + 8, h1, l1);
    *reinterpret_cast<uint4 *>(dst_hi) = h0;
    *reinterpret_cast<uint4 *>(dst_hi + 128) = h1;
    *reinterpret_cast<uint4 *>(dst_lo) = l0;
    *reinterpret_cast<uint4 *>(dst_lo + 128) = l1;
}

// Last-layer epilogue of one 128-cout chunk: thread = cout; max over each centre's NS consecutive columns,
// + bias, ReLU, store.  The output cursors advance incrementally with the centre.
struct PoolOut {
    float mx16 = 0.f;   // largest value written as fp16 (range guard)
    float bv;
    bool w32, w16;
    float *outc;
    __half *out16;
    int ld16, o16lo;
    long long q, qmax;
    int m, p;
    size_t scene_stride;
    __device__ __forceinline__ void emit(float run) {
        const float y = fmaxf(run + bv, 0.f);
        if (q < qmax) {
            if (w32) *outc = y;
            if (w16) {
                mx16 = fmaxf(mx16, y);
                const __half h = __float2half_rn(y);
                *out16 = h;
                if (o16lo > 0) out16[o16lo] = __float2half_rn(y - __half2float(h));
            }
        }
        ++q;
        ++outc;
        out16 += ld16;
        if (++p == m) { p = 0; outc += scene_stride - (size_t)m; }
    }
};
// NS = 16 / 32 (every shipped IA-SSD / SPSNet config): compile-time group boundaries
template <int NS, int NCOLS = MM_ROWS>
__device__ __forceinline__ void pool_chunk(uint32_t taddr, PoolOut &o) {
    float run = -3.0e38f;
#pragma unroll 1
    for (int c0 = 0; c0 < NCOLS; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            run = fmaxf(run, v[i]);
            if (((i + 1) % NS) == 0) { o.emit(run); run = -3.0e38f; }
        }
    }
}
// any power of two <= 128
static __device__ __noinline__ void pool_chunk_any(uint32_t taddr, PoolOut &o, int ns, int ncols = MM_ROWS) {
    float run = -3.0e38f;
#pragma unroll 1
    for (int c0 = 0; c0 < ncols; c0 += 16) {
        float v[16];
        tmem_ld16(taddr + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            run = fmaxf(run, v[i]);
            if (((c0 + i + 1) & (ns - 1)) == 0) { o.emit(run); run = -3.0e38f; }
        }
    }
}

static_assert(sizeof(SaArgs) <= 4096, "SaArgs travels in the kernel parameters: keep it within the classic 4 KB limit");

// Batch-statistics pass: thread = cout, columns = the tile's rows.  Sum and sum of squares of the raw accumulators over the
// `nv` real rows of the tile (fp32 over <= 128 values), added in fp64 to the cell this thread owns (no other thread of the grid
// touches it: plain read-modify-write, reproducible bit for bit).
static __device__ __forceinline__ void stats_chunk(uint32_t taddr, int nv, double *cell) {
    float s = 0.f, q = 0.f;
#pragma unroll 1
    for (int c0 = 0; c0 < nv; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const float x = (c0 + i < nv) ? v[i] : 0.f;
            s += x;
            q = fmaf(x, x, q);
        }
    }
    double2 acc = *reinterpret_cast<double2 *>(cell);
    acc.x += (double)s;
    acc.y += (double)q;
    *reinterpret_cast<double2 *>(cell) = acc;
}

// ---- optional role profiling: cycles spent per wait / work category, summed over CTAs ------------------------
enum { PF_MMA_TOTAL = 0, PF_MMA_ACC_EMPTY, PF_MMA_W_FULL, PF_MMA_XR, PF_PROD_W_EMPTY, PF_PROD_HID, PF_EPI_TOTAL, PF_EPI_GATHER,
       PF_EPI_WAIT_HID, PF_EPI_WORK_HID, PF_EPI_WAIT_POOL, PF_EPI_WORK_POOL, PF_MMA_ISSUE, PF_MMA_COMMIT, PF_COUNT };
// Phase counters.  ON = false compiles to nothing: the MMA-issue and epilogue loops are issue-bound, and even predicated-off
// clock reads and adds cost their slots (the same lesson as in fps.cu: -23 % there).  The kernels are instantiated both ways
// and the counting variant is launched only while spsk_sa_mma_set_profile() has a buffer installed.
template <bool ON>
struct ProfT {
    unsigned long long acc[PF_COUNT];
    bool on;
    __device__ __forceinline__ void init(bool enable) {
        on = enable;
#pragma unroll
        for (int i = 0; i < PF_COUNT; ++i) acc[i] = 0ull;
    }
    __device__ __forceinline__ long long now() const { return on ? clock64() : 0ll; }
    __device__ __forceinline__ void add(int k, long long t0) { if (on) acc[k] += (unsigned long long)(clock64() - t0); }
    __device__ __forceinline__ void flush(unsigned long long *dst) const {
        if (!on) return;
        for (int i = 0; i < PF_COUNT; ++i)
            if (acc[i]) atomicAdd(dst + i, acc[i]);
    }
};
template <>
struct ProfT<false> {
    __device__ __forceinline__ void init(bool) {}
    __device__ __forceinline__ long long now() const { return 0ll; }
    __device__ __forceinline__ void add(int, long long) {}
    __device__ __forceinline__ void flush(unsigned long long *) const {}
};
using Prof = ProfT<true>;


}  // namespace spsk
