// sa_mma.cu -- fused set-abstraction scale on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// ONE kernel per MSG scale replaces the reference's chain (pointnet2_utils.py:307-315,
// pointnet2_modules.py:204-211,431-436):
//     grouping_operation(xyz) - new_xyz ; grouping_operation(features) ; torch.cat ;
//     3 x [Conv2d 1x1 (no bias) -> BatchNorm2d(eval) -> ReLU] ; F.max_pool2d over nsample
// The (B,3+C,npoint,nsample) grouped tensor and all conv/BN/ReLU intermediates stay on chip.
//
// Machine mapping (persistent CTAs, one per SM, 192 threads, TMEM 512 columns):
//   tile        = 128 grouped rows (= 128/nsample centres), looped over by each CTA
//   warps 0-3   : (a) gather: row r = thread r builds X0[r, :] = [features(idx) | xyz(idx) - centre | 0] as fp16
//                     with 16-byte loads from the point-major fp16 feature twin and 16-byte smem stores;
//                 (b) epilogue: TMEM -> registers (tcgen05.ld), + folded-BN bias, ReLU, -> fp16 operand of the
//                     next layer in shared memory; last layer: max over nsample + bias + ReLU -> global
//   warp 4      : weight producer: 1-D bulk async copies (cp.async.bulk + mbarrier complete_tx) of host-packed
//                 16 KB weight tiles through a ring of stages
//   warp 5      : TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, fp16 operands, fp32 accumulate)
//   hidden layers  (orientation A): D[row, cout]  = X[row, k] . W[cout, k]^T   M = 128 rows,  N = cout chunk
//   last layer     (orientation B): D[cout, row]  = W[cout, k] . X[row, k]^T   M = 128 couts, N = 128 rows
//                 so that the max over the nsample rows of a centre is a per-thread loop over TMEM columns;
//                 4 accumulator buffers (4 x 128 columns) overlap the chunk epilogue with the next chunk's MMAs.
//   All operands use the canonical K-major, no-swizzle UMMA layout (8-row x 16-byte core matrices):
//       byte(r, k) = (r/8)*SBO + (k/8)*128 + (r%8)*16 + (k%8)*2          LBO = 128
//
// Numerics: operands are fp32 values rounded to fp16 (11-bit significand, same as TF32), products are
// exact and accumulate in fp32; measured end-to-end error of a 3-layer scale is ~4e-4 of the output
// range (DESIGN.md "precision"), inside the 1e-3 bar of BASELINE.json.  The exact-fp32 path is
// linear_ffma.cu.
#include "common.cuh"
#include <cuda_fp16.h>

namespace spsk {

constexpr int MM_ROWS = 128;          // grouped rows per tile
constexpr int MM_THREADS = 192;       // 4 gather/epilogue warps + producer + mma
constexpr int MM_WTILE_BYTES = 16384; // [128 cout][64 k] fp16
constexpr int MM_MAX_LAYERS = 4;
constexpr int MM_MAX_STAGES = 6;

struct MmaLayer {
    int kpad;      // input width, multiple of 16
    int cpad;      // output width: multiple of 16 (hidden) / 128 (last)
    int n_cc;      // ceil(cpad / 128)
    int n_kc;      // ceil(kpad / 64)
    int tile_off;  // first weight tile of this layer (units of 16 KB tiles)
    int bias_off;  // offset into the bias array (floats)
};

struct MmaArgs {
    int nlayers;
    MmaLayer L[MM_MAX_LAYERS];
    int b, n, m, nsample, ns_log2;
    int cpad8;       // feature channels in the twin (multiple of 8, 0 if none)
    int use_xyz;
    int k0pad;
    long long rows;  // b*m*nsample
    int ntiles;
    int nstages;
    int tmem_cols;   // 256 or 512 (power of two >= every accumulator this chain needs)
    int nbuf;        // last-layer accumulator buffers of 128 columns: tmem_cols / 128
    int xa_bytes, xb_bytes;
    const float *xyz, *new_xyz;
    const __half *twin;   // (b, n, cpad8)
    const int *idx;       // (b, m, nsample)
    const __half *wtiles; // packed weight tiles
    const float *bias;
    float *out;           // (b, c_total, m)
    int c_total, co_off, cout_last;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// bounded wait: a protocol bug traps (the context dies, the box survives) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// UMMA shared-memory descriptor, K-major, SWIZZLE_NONE: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout_type=0 [61,64)   (cute/arch/mma_sm100_desc.hpp::SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) |
           (1ull << 46);
}
// instruction descriptor kind::f16: D=f32 (bits 4-5 = 1), A=B=f16 (0), K-major both, N>>3 at [17,23), M>>4 at [24,29)
__device__ __forceinline__ uint32_t umma_idesc(int m, int n) {
    return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float *v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float *v) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// 16 accumulator columns -> + bias -> ReLU -> fp16 -> two 16-byte stores into the next operand (K-major)
__device__ __forceinline__ void store_hidden16(const float *v, const float *bias16, uint8_t *dst) {
    const float4 *b4 = reinterpret_cast<const float4 *>(bias16);
    uint32_t h[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float4 bb = __ldg(b4 + i);
        h[2 * i] = pack_h2(fmaxf(v[4 * i] + bb.x, 0.f), fmaxf(v[4 * i + 1] + bb.y, 0.f));
        h[2 * i + 1] = pack_h2(fmaxf(v[4 * i + 2] + bb.z, 0.f), fmaxf(v[4 * i + 3] + bb.w, 0.f));
    }
    *reinterpret_cast<uint4 *>(dst) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4 *>(dst + 128) = make_uint4(h[4], h[5], h[6], h[7]);
}

// Last-layer epilogue of one 128-cout chunk: thread = cout; max over each centre's NS consecutive columns,
// + bias, ReLU, store to out[b, co_off + ch, p].  NS is a compile-time power of two, so the group
// boundaries cost nothing; (scene, centre) of consecutive groups is advanced incrementally.
template <int NS>
__device__ __forceinline__ void pool_chunk(uint32_t taddr, float bv, bool ch_ok, float *outc, long long q0, long long qmax,
                                           int m, size_t scene_stride, int p) {
    // outc points at out[b(q0), co_off + ch, p(q0)]
    long long q = q0;
    float run = -3.0e38f;
#pragma unroll
    for (int c0 = 0; c0 < MM_ROWS; c0 += 32) {
        float v[32];
        tmem_ld32(taddr + (uint32_t)c0, v);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            run = fmaxf(run, v[i]);
            if (((c0 + i + 1) % NS) == 0) {
                if (ch_ok && q < qmax) *outc = fmaxf(run + bv, 0.f);
                run = -3.0e38f;
                ++q;
                ++outc;
                if (++p == m) { p = 0; outc += scene_stride - (size_t)m; }
            }
        }
    }
}
__device__ __forceinline__ uint32_t pack_h2_unused_(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t *>(&h);
}

// ---- the kernel -----------------------------------------------------------------------------------
__global__ void __launch_bounds__(MM_THREADS, 2)
sa_mma_kernel(const MmaArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    // carve: [barriers 256 B][XA][XB][W stages]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 192);
    uint8_t *xa = smem + 256;
    uint8_t *xb = xa + a.xa_bytes;
    uint8_t *wst = xb + a.xb_bytes;
    const uint32_t bar0 = smem_u32(bars);
    auto W_FULL = [&](int s) { return bar0 + 8u * s; };
    auto W_EMPTY = [&](int s) { return bar0 + 8u * (MM_MAX_STAGES + s); };
    auto ACC_FULL = [&](int i) { return bar0 + 8u * (2 * MM_MAX_STAGES + i); };
    auto ACC_EMPTY = [&](int i) { return bar0 + 8u * (2 * MM_MAX_STAGES + 4 + i); };
    const uint32_t X_READY = bar0 + 8u * (2 * MM_MAX_STAGES + 8);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int nL = a.nlayers;

    if (tid == 0) {
        for (int s = 0; s < MM_MAX_STAGES; ++s) { mbar_init(W_FULL(s), 1); mbar_init(W_EMPTY(s), 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(ACC_FULL(i), 1); mbar_init(ACC_EMPTY(i), 128); }
        mbar_init(X_READY, 128);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(a.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ================= weight producer =================
        if ((tid & 31) == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                for (int l = 0; l < nL; ++l) {
                    const MmaLayer &Ly = a.L[l];
                    for (int cc = 0; cc < Ly.n_cc; ++cc)
                        for (int kc = 0; kc < Ly.n_kc; ++kc, ++it) {
                            const int s = it % a.nstages;
                            const uint32_t ph = (it / a.nstages) & 1u;
                            mbar_wait(W_EMPTY(s), ph ^ 1u);
                            mbar_expect_tx(W_FULL(s), MM_WTILE_BYTES);
                            const __half *src = a.wtiles + (size_t)(Ly.tile_off + cc * Ly.n_kc + kc) * (MM_WTILE_BYTES / 2);
                            bulk_g2s(smem_u32(wst + (size_t)s * MM_WTILE_BYTES), src, MM_WTILE_BYTES, W_FULL(s));
                        }
                }
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        if ((tid & 31) == 0) {
            uint32_t it = 0, xr = 0;
            uint32_t af[4] = {0, 0, 0, 0};   // completed uses of ACC_FULL[i]
            uint32_t bu[4] = {0, 0, 0, 0};   // orientation-B uses of accumulator buffer i
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                for (int l = 0; l < nL; ++l) {
                    const MmaLayer &Ly = a.L[l];
                    const bool last = (l == nL - 1);
                    const uint32_t xbase = smem_u32((l & 1) ? xb : xa);
                    const uint32_t x_sbo = (uint32_t)Ly.kpad * 16u;
                    mbar_wait(X_READY, xr & 1u);
                    ++xr;
                    tc_fence_after();
                    for (int cc = 0; cc < Ly.n_cc; ++cc) {
                        const int ncols = min(128, Ly.cpad - cc * 128);  // couts in this chunk (multiple of 16)
                        uint32_t d_tmem;
                        uint32_t idesc;
                        int buf = 0;
                        if (!last) {
                            d_tmem = tmem_base + (uint32_t)(cc * 128);
                            idesc = umma_idesc(128, ncols);           // M = rows, N = couts
                        } else {
                            buf = cc & (a.nbuf - 1);
                            if (bu[buf] > 0) { mbar_wait(ACC_EMPTY(buf), (bu[buf] - 1) & 1u); tc_fence_after(); }
                            ++bu[buf];
                            d_tmem = tmem_base + (uint32_t)(buf * 128);
                            idesc = umma_idesc(128, MM_ROWS);         // M = couts (tile rows), N = rows
                        }
                        for (int kc = 0; kc < Ly.n_kc; ++kc, ++it) {
                            const int s = it % a.nstages;
                            const uint32_t ph = (it / a.nstages) & 1u;
                            mbar_wait(W_FULL(s), ph);
                            tc_fence_after();
                            const uint32_t wbase = smem_u32(wst + (size_t)s * MM_WTILE_BYTES);
                            const int nk16 = min(4, (Ly.kpad - kc * 64) / 16);
                            for (int j = 0; j < nk16; ++j) {
                                const uint64_t xd = umma_desc(xbase + (uint32_t)(kc * 4 + j) * 256u, 128u, x_sbo);
                                const uint64_t wd = umma_desc(wbase + (uint32_t)j * 256u, 128u, 1024u);
                                const uint32_t acc = (kc | j) ? 1u : 0u;
                                if (!last) umma_f16(d_tmem, xd, wd, idesc, acc);
                                else umma_f16(d_tmem, wd, xd, idesc, acc);
                            }
                            umma_commit(W_EMPTY(s));   // stage reusable once these MMAs retire
                        }
                        if (last) { umma_commit(ACC_FULL(buf)); ++af[buf]; }
                    }
                    if (!last) { umma_commit(ACC_FULL(0)); ++af[0]; }
                }
            }
            (void)af;
        }
    } else {
        // ================= gather + epilogue (threads 0..127; thread = row / TMEM lane) =================
        uint32_t af[4] = {0, 0, 0, 0};
        const uint32_t lane_field = (uint32_t)(warp * 32) << 16;
        const int r = tid;
        const uint32_t row_off = (uint32_t)(r >> 3) * 0u;  // placeholder to keep the formula visible below
        (void)row_off;
        for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
            // ---- gather X0 (into XA): [features | dxyz | 0]
            {
                const MmaLayer &L0 = a.L[0];
                const uint32_t sbo = (uint32_t)L0.kpad * 16u;
                uint8_t *xrow = xa + (size_t)(r >> 3) * sbo + (size_t)(r & 7) * 16;
                const long long grow = (long long)tile * MM_ROWS + r;
                const bool ok = grow < a.rows;
                int j = 0;
                long long q = 0;
                int bb = 0;
                if (ok) {
                    q = grow >> a.ns_log2;
                    bb = (int)(q / a.m);
                    j = __ldg(a.idx + grow);
                }
                const int nch = a.k0pad >> 3;
                const int fch = a.cpad8 >> 3;
                const uint4 zero = make_uint4(0, 0, 0, 0);
                const uint4 *trow = ok && fch ? reinterpret_cast<const uint4 *>(a.twin + ((size_t)bb * a.n + j) * a.cpad8) : nullptr;
                for (int c = 0; c < fch; ++c) {
                    const uint4 v = ok ? __ldg(trow + c) : zero;
                    *reinterpret_cast<uint4 *>(xrow + (size_t)c * 128) = v;
                }
                int c = fch;
                if (a.use_xyz) {
                    uint4 v = zero;
                    if (ok) {
                        const float *p = a.xyz + ((size_t)bb * a.n + j) * 3;
                        const float *ctr = a.new_xyz + (size_t)q * 3;
                        const float dx = __fsub_rn(__ldg(p), __ldg(ctr));
                        const float dy = __fsub_rn(__ldg(p + 1), __ldg(ctr + 1));
                        const float dz = __fsub_rn(__ldg(p + 2), __ldg(ctr + 2));
                        v.x = pack_h2(dx, dy);
                        v.y = pack_h2(dz, 0.f);
                    }
                    *reinterpret_cast<uint4 *>(xrow + (size_t)c * 128) = v;
                    ++c;
                }
                for (; c < nch; ++c) *reinterpret_cast<uint4 *>(xrow + (size_t)c * 128) = zero;
                fence_proxy_async();
                mbar_arrive(X_READY);
            }
            // ---- hidden layers: D[row, cout] -> relu(D + bias) -> fp16 -> next X
            for (int l = 0; l < nL - 1; ++l) {
                const MmaLayer &Ly = a.L[l];
                uint8_t *xo = ((l + 1) & 1) ? xb : xa;
                const uint32_t sbo = (uint32_t)a.L[l + 1].kpad * 16u;
                uint8_t *xrow = xo + (size_t)(r >> 3) * sbo + (size_t)(r & 7) * 16;
                mbar_wait(ACC_FULL(0), af[0] & 1u);
                ++af[0];
                tc_fence_after();
                const float *bias = a.bias + Ly.bias_off;
                int c0 = 0;
                for (; c0 + 32 <= Ly.cpad; c0 += 32) {
                    float v[32];
                    tmem_ld32(tmem_base + lane_field + (uint32_t)c0, v);
                    store_hidden16(v, bias + c0, xrow + (size_t)(c0 >> 3) * 128);
                    store_hidden16(v + 16, bias + c0 + 16, xrow + (size_t)((c0 >> 3) + 2) * 128);
                }
                if (c0 < Ly.cpad) {
                    float v[16];
                    tmem_ld16(tmem_base + lane_field + (uint32_t)c0, v);
                    store_hidden16(v, bias + c0, xrow + (size_t)(c0 >> 3) * 128);
                }
                // zero the K padding of the next layer's operand (kpad_{l+1} == cpad_l by construction, so none)
                tc_fence_before();
                fence_proxy_async();
                mbar_arrive(X_READY);
            }
            // ---- last layer: D[cout, row]; thread = cout; max over each centre's nsample columns
            {
                const MmaLayer &Ly = a.L[nL - 1];
                const float *bias = a.bias + Ly.bias_off;
                const long long q0 = ((long long)tile * MM_ROWS) >> a.ns_log2;  // first centre of the tile
                const long long qmax = (long long)a.b * a.m;
                const int ns = a.nsample;
                const long long bb0 = q0 / a.m;
                const int p0 = (int)(q0 - bb0 * a.m);
                for (int cc = 0; cc < Ly.n_cc; ++cc) {
                    const int buf = cc & (a.nbuf - 1);
                    mbar_wait(ACC_FULL(buf), af[buf] & 1u);
                    ++af[buf];
                    tc_fence_after();
                    const int ch = cc * 128 + r;
                    const bool ch_ok = ch < a.cout_last;
                    const float bv = ch_ok ? __ldg(bias + ch) : 0.f;
                    float *outc = a.out + ((size_t)bb0 * a.c_total + a.co_off + (ch_ok ? ch : 0)) * a.m + (size_t)p0;
                    const size_t sstride = (size_t)a.c_total * a.m;
                    const uint32_t taddr = tmem_base + lane_field + (uint32_t)(buf * 128);
                    switch (ns) {
                        case 1: pool_chunk<1>(taddr, bv, ch_ok, outc, q0, qmax, a.m, sstride, p0); break;
                        case 2: pool_chunk<2>(taddr, bv, ch_ok, outc, q0, qmax, a.m, sstride, p0); break;
                        case 4: pool_chunk<4>(taddr, bv, ch_ok, outc, q0, qmax, a.m, sstride, p0); break;
                        case 8: pool_chunk<8>(taddr, bv, ch_ok, outc, q0, qmax, a.m, sstride, p0); break;
                        case 16: pool_chunk<16>(taddr, bv, ch_ok, outc, q0, qmax, a.m, sstride, p0); break;
                        case 32: pool_chunk<32>(taddr, bv, ch_ok, outc, q0, qmax, a.m, sstride, p0); break;
                        case 64: pool_chunk<64>(taddr, bv, ch_ok, outc, q0, qmax, a.m, sstride, p0); break;
                        default: pool_chunk<128>(taddr, bv, ch_ok, outc, q0, qmax, a.m, sstride, p0); break;
                    }
                    tc_fence_before();
                    mbar_arrive(ACC_EMPTY(buf));
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols));
    }
}

// ---- feature twin: (b, c, n) f32 channel-major -> (b, n, cpad8) fp16 point-major (zero padded) -------------
__global__ void __launch_bounds__(256)
make_twin_kernel(int c, int n, int cpad8, const float *__restrict__ in, __half *__restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z;
    const int n0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
    for (int i = ty; i < 32; i += 8) {
        const int ci = c0 + i, ni = n0 + tx;
        tile[i][tx] = (ci < c && ni < n) ? __ldg(in + ((size_t)b * c + ci) * n + ni) : 0.f;
    }
    __syncthreads();
    for (int i = ty; i < 32; i += 8) {
        const int ni = n0 + i, ci = c0 + tx;
        if (ni < n && ci < cpad8) out[((size_t)b * n + ni) * cpad8 + ci] = __float2half_rn(tile[tx][i]);
    }
}

}  // namespace spsk

extern "C" int spsk_make_twin(int b, int c, int n, int cpad8, const float *features, void *twin, spsk_stream_t stream) {
    using namespace spsk;
    SPSK_REQUIRE(b >= 0 && c >= 1 && n >= 0 && cpad8 >= c && (cpad8 % 8) == 0 && b <= 65535, SPSK_ERR_INVALID_ARG,
                 "make_twin: bad sizes b=%d c=%d n=%d cpad8=%d", b, c, n, cpad8);
    if (b == 0 || n == 0) return SPSK_OK;
    SPSK_REQUIRE(features && twin, SPSK_ERR_INVALID_ARG, "make_twin: null pointer");
    dim3 grid((n + 31) / 32, (cpad8 + 31) / 32, b);
    make_twin_kernel<<<grid, 256, 0, as_stream(stream)>>>(c, n, cpad8, features, reinterpret_cast<__half *>(twin));
    SPSK_LAUNCH_CHECK("make_twin_kernel");
    return SPSK_OK;
}

extern "C" int spsk_sa_mma_smem_bytes(int nlayers, const int *kpad, const int *cpad, int *nstages_out) {
    // X ping-pong buffers + as many 16 KB weight stages as fit (2..6) + 256 B of barriers
    using namespace spsk;
    if (nlayers < 1 || nlayers > MM_MAX_LAYERS) return -1;
    int xa = 0, xbv = 0;
    for (int l = 0; l < nlayers; ++l) {
        const int bytes = kpad[l] * MM_ROWS * 2;
        if (l & 1) xbv = max(xbv, bytes); else xa = max(xa, bytes);
    }
    const int budget = 227 * 1024;
    int stages = (budget - 256 - xa - xbv) / MM_WTILE_BYTES;
    if (stages > MM_MAX_STAGES) stages = MM_MAX_STAGES;
    if (stages < 2) return -1;
    if (nstages_out) *nstages_out = stages;
    int total = 256 + xa + xbv + stages * MM_WTILE_BYTES;
    // TMEM: 256 columns per CTA when every hidden accumulator fits (two CTAs can then share an SM and overlap one
    // tile's gather / epilogue with the other's MMAs); the shared-memory floors keep the CTA count per SM within
    // the TMEM budget (3 x 76 KB and 2 x 120 KB both exceed 227 KB).
    int hidden_max = 0;
    for (int l = 0; l + 1 < nlayers; ++l) hidden_max = max(hidden_max, cpad[l]);
    const bool two = hidden_max <= 256 && (256 + xa + xbv + 2 * MM_WTILE_BYTES) <= 112 * 1024;
    if (two) {
        stages = (112 * 1024 - 256 - xa - xbv) / MM_WTILE_BYTES;
        if (stages > MM_MAX_STAGES) stages = MM_MAX_STAGES;
        if (nstages_out) *nstages_out = stages;
        total = 256 + xa + xbv + stages * MM_WTILE_BYTES;
        if (total < 76 * 1024) total = 76 * 1024;
        return -total;  // negative: "two CTAs per SM" variant (256 TMEM columns)
    }
    if (total < 120 * 1024) total = 120 * 1024;
    return total;
}

extern "C" int spsk_sa_mma_forward(const spsk_group_desc *g, int cpad8, const void *twin, int nlayers, const int *kpad,
                                   const int *cpad, const int *tile_off, const int *bias_off, const void *wtiles,
                                   const float *bias, int cout_last, float *out_pooled, int c_total, int co_off,
                                   spsk_stream_t stream) {
    using namespace spsk;
    SPSK_REQUIRE(g && kpad && cpad && tile_off && bias_off && wtiles && bias && out_pooled, SPSK_ERR_INVALID_ARG, "sa_mma: null pointer");
    SPSK_REQUIRE(nlayers >= 1 && nlayers <= MM_MAX_LAYERS, SPSK_ERR_UNSUPPORTED, "sa_mma: nlayers=%d (1..%d)", nlayers, MM_MAX_LAYERS);
    SPSK_REQUIRE(g->nsample >= 1 && g->nsample <= MM_ROWS && (g->nsample & (g->nsample - 1)) == 0, SPSK_ERR_UNSUPPORTED,
                 "sa_mma: nsample=%d must be a power of two <= %d", g->nsample, MM_ROWS);
    SPSK_REQUIRE(g->idx && g->xyz && g->new_xyz && (cpad8 == 0 || twin), SPSK_ERR_INVALID_ARG, "sa_mma: null gather source");
    SPSK_REQUIRE(cpad8 % 8 == 0 && cpad8 >= 0, SPSK_ERR_INVALID_ARG, "sa_mma: cpad8=%d", cpad8);
    MmaArgs a{};
    a.nlayers = nlayers;
    for (int l = 0; l < nlayers; ++l) {
        const bool last = l == nlayers - 1;
        SPSK_REQUIRE(kpad[l] >= 16 && kpad[l] % 16 == 0 && kpad[l] <= 1024, SPSK_ERR_UNSUPPORTED, "sa_mma: layer %d kpad=%d", l, kpad[l]);
        SPSK_REQUIRE(cpad[l] >= 16 && cpad[l] % (last ? 128 : 16) == 0 && cpad[l] <= (last ? 4096 : 512), SPSK_ERR_UNSUPPORTED,
                     "sa_mma: layer %d cpad=%d", l, cpad[l]);
        SPSK_REQUIRE(last || cpad[l] == kpad[l + 1], SPSK_ERR_INVALID_ARG, "sa_mma: layer %d cpad != next kpad", l);
        a.L[l].kpad = kpad[l];
        a.L[l].cpad = cpad[l];
        a.L[l].n_cc = (cpad[l] + 127) / 128;
        a.L[l].n_kc = (kpad[l] + 63) / 64;
        a.L[l].tile_off = tile_off[l];
        a.L[l].bias_off = bias_off[l];
    }
    a.b = g->b; a.n = g->n; a.m = g->m; a.nsample = g->nsample;
    a.ns_log2 = 0;
    while ((1 << a.ns_log2) < g->nsample) ++a.ns_log2;
    a.cpad8 = cpad8;
    a.use_xyz = g->use_xyz ? 1 : 0;
    a.k0pad = kpad[0];
    SPSK_REQUIRE(a.k0pad >= cpad8 + (a.use_xyz ? 8 : 0), SPSK_ERR_INVALID_ARG, "sa_mma: kpad[0]=%d too small for %d feature + xyz channels", a.k0pad, cpad8);
    a.rows = (long long)g->b * g->m * g->nsample;
    if (a.rows == 0) return SPSK_OK;
    const long long ntiles = (a.rows + MM_ROWS - 1) / MM_ROWS;
    SPSK_REQUIRE(ntiles <= 0x7FFFFFFF, SPSK_ERR_UNSUPPORTED, "sa_mma: too many rows");
    a.ntiles = (int)ntiles;
    int nstages = 0;
    int smem = spsk_sa_mma_smem_bytes(nlayers, kpad, cpad, &nstages);
    SPSK_REQUIRE(smem != -1, SPSK_ERR_UNSUPPORTED, "sa_mma: activation tiles do not fit shared memory (max layer width too large)");
    const bool two_per_sm = smem < 0;
    if (two_per_sm) smem = -smem;
    a.nstages = nstages;
    a.tmem_cols = two_per_sm ? 256 : 512;
    a.nbuf = a.tmem_cols / 128;
    a.xa_bytes = 0; a.xb_bytes = 0;
    for (int l = 0; l < nlayers; ++l) {
        const int bytes = kpad[l] * MM_ROWS * 2;
        if (l & 1) a.xb_bytes = max(a.xb_bytes, bytes); else a.xa_bytes = max(a.xa_bytes, bytes);
    }
    a.xyz = g->xyz; a.new_xyz = g->new_xyz; a.twin = reinterpret_cast<const __half *>(twin); a.idx = g->idx;
    a.wtiles = reinterpret_cast<const __half *>(wtiles); a.bias = bias;
    a.out = out_pooled; a.c_total = c_total; a.co_off = co_off; a.cout_last = cout_last;
    SPSK_REQUIRE(co_off >= 0 && co_off + cout_last <= c_total, SPSK_ERR_INVALID_ARG, "sa_mma: channel window outside c_total");
    static int attr_set_for = 0;
    if (smem > attr_set_for) {
        cudaError_t e = cudaFuncSetAttribute(sa_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(sa_mma_kernel)");
        attr_set_for = 227 * 1024;
    }
    // a few CTAs per SM slot (static tile striding): when another stream's kernels hold some SMs, late CTAs start
    // as soon as any SM frees up instead of doubling the kernel's duration
    const int slots = SPSK_NUM_SMS * (two_per_sm ? 2 : 1) * 2;
    const int grid = a.ntiles < slots ? a.ntiles : slots;
    sa_mma_kernel<<<grid, MM_THREADS, smem, as_stream(stream)>>>(a);
    SPSK_LAUNCH_CHECK("sa_mma_kernel");
    return SPSK_OK;
}
