// pw_mma.cu -- point-wise (1x1 Conv1d + folded BN [+ ReLU]) layer on the 5th-gen tensor cores, sm_100a.
//
// Replaces the aggregation / confidence / vote Conv1d stacks of the reference
// (pointnet2_modules.py:216-243, 447-458, 485-500), which run as cuDNN/cuBLAS GEMMs over the channel-major
// (B, C, npoint) tensor.  Here activations travel between the fused kernels as POINT-MAJOR fp16 rows
// (rows = B*npoint, K contiguous), which is directly the K-major A operand of tcgen05.mma:
//
//     Y[row, n] = act( sum_k X[row, k] * W[n, k] + bias[n] )        fp16 operands, fp32 accumulate
//
//   grid   = (ceil(rows/128), ceil(n/128)); one 128 x <=128 output tile per CTA, K streamed in 64-wide chunks
//   warps 0-3 : (a) A loader: 16-byte cp.async (LDGSTS) from the row-major activations straight into the
//                   canonical K-major no-swizzle UMMA layout (8 rows x 16 B core matrices), 2 chunks in flight;
//               (b) epilogue: tcgen05.ld -> + bias -> ReLU -> up to three views of the result:
//                   fp32 channel-major (the reference's (B, C, npoint) tensor), fp16 point-major (input of the
//                   next fused kernel / gather twin), fp32 point-major (class logits (B, npoint, num_class))
//   warp 4    : W producer: host-packed 16 KB weight tiles by 1-D bulk async copy (mbarrier complete_tx)
//   warp 5    : TMEM allocator + single-thread tcgen05.mma issuer
#include "mma_ptx.cuh"

namespace spsk {

constexpr int PW_ROWS = 128;
constexpr int PW_THREADS = 192;
constexpr int PW_TILE_BYTES = 16384;                       // 128 rows x 64 k fp16
// <stages, lag, split>: ring depth, how many chunks of cp.async stay in flight behind the one being issued, arithmetic.
// plain: a stage = [X chunk | W tile] = 32 KB; short K <3, 2> (96 KB, two CTAs per SM), long K <6, 4> (192 KB).
// split: a stage = [Xh | Xl | Wh | Wl] = 64 KB (<3, 2>, 192 KB): each 64-wide chunk of the TRUE K is staged once and
//        multiplied three ways (Xh.Wh + Xl.Wh + Xh.Wl) -- 2/3 of the shared-memory / L2 traffic of streaming the
//        K-concatenated [Xh | Xl | Xh] . [Wh ; Wh ; Wl] product, and a third of the pipeline steps.
constexpr int pw_smem(int stages, bool split) { return 256 + stages * (split ? 4 : 2) * PW_TILE_BYTES; }

struct PwArgs {
    int rows, k, ldx, n, npad, n_kc, relu;
    int split, xlo;              // split: x rows are [hi (k) ... lo (k) at column xlo]
    int o16lo;                   // > 0: out16 also receives the fp16 residual of every value at column o16lo + c
    const __half *x;
    const __half *wtiles;
    const float *bias;
    float *out_cm; int m, c_total, co_off;
    __half *out16; int ld16, n16;
    float *out_pm; int ldpm;
    unsigned int *ovf; unsigned int ovf_bit;   // fp16 range guard (api.cu)
};

template <int PW_STAGES, int PW_LAG, bool SPLIT>
__global__ void __launch_bounds__(PW_THREADS, (PW_STAGES <= 3 && !SPLIT) ? 2 : 1)
pw_mma_kernel(const PwArgs a) {
    constexpr int STAGE_BYTES = (SPLIT ? 4 : 2) * PW_TILE_BYTES;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + 192);
    uint8_t *stage0 = smem + 256;
    const uint32_t bar0 = smem_u32(bars);
    auto FULL = [&](int s) { return bar0 + 8u * s; };
    auto EMPTY = [&](int s) { return bar0 + 8u * (PW_STAGES + s); };
    const uint32_t ACC_FULL = bar0 + 8u * (2 * PW_STAGES);
    auto XS = [&](int s) { return stage0 + (size_t)s * STAGE_BYTES; };                                        // [Xh | Xl]
    auto WS = [&](int s) { return stage0 + (size_t)s * STAGE_BYTES + (SPLIT ? 2 : 1) * PW_TILE_BYTES; };       // [Wh | Wl]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tile = blockIdx.x, cc = blockIdx.y;
    const int ncols = min(128, a.npad - cc * 128);   // multiple of 16

    if (tid == 0) {
        for (int s = 0; s < PW_STAGES; ++s) { mbar_init(FULL(s), 128 + 1); mbar_init(EMPTY(s), 1); }
        mbar_init(ACC_FULL, 1);
        mbar_init_fence();
    }
    if (warp == 5) tmem_alloc(smem_u32(tmem_slot), 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        const bool leader = elect_one();
        for (int kc = 0; kc < a.n_kc; ++kc) {
            const int s = kc % PW_STAGES;
            const uint32_t ph = (kc / PW_STAGES) & 1u;
            mbar_wait(EMPTY(s), ph ^ 1u);
            if (leader) {
                constexpr uint32_t WB = (SPLIT ? 2 : 1) * PW_TILE_BYTES;   // split: [Wh tile | Wl tile] are adjacent in the packing
                mbar_expect_tx(FULL(s), WB);
                bulk_g2s(smem_u32(WS(s)), a.wtiles + (size_t)(cc * a.n_kc + kc) * (WB / 2), WB, FULL(s));
            }
            __syncwarp();
        }
    } else if (warp == 5) {
        const bool leader = elect_one();
        const uint32_t idesc = umma_idesc(128, ncols);
        for (int kc = 0; kc < a.n_kc; ++kc) {
            const int s = kc % PW_STAGES;
            const uint32_t ph = (kc / PW_STAGES) & 1u;
            mbar_wait(FULL(s), ph);
            tc_fence_after();
            const uint32_t xb = smem_u32(XS(s)), wb = smem_u32(WS(s));
            const int nk16 = min(4, (a.k - kc * 64) / 16);
            if (leader) {
                for (int j = 0; j < nk16; ++j) {
                    const uint64_t xh = umma_desc(xb + (uint32_t)j * 256u, 128u, 1024u), wh = umma_desc(wb + (uint32_t)j * 256u, 128u, 1024u);
                    umma_f16(tmem_base, xh, wh, idesc, (kc | j) ? 1u : 0u);
                    if (SPLIT) {
                        umma_f16(tmem_base, umma_desc(xb + PW_TILE_BYTES + (uint32_t)j * 256u, 128u, 1024u), wh, idesc, 1u);   // Xl . Wh
                        umma_f16(tmem_base, xh, umma_desc(wb + PW_TILE_BYTES + (uint32_t)j * 256u, 128u, 1024u), idesc, 1u);   // Xh . Wl
                    }
                }
                umma_commit(EMPTY(s));
                if (kc == a.n_kc - 1) umma_commit(ACC_FULL);
            }
            __syncwarp();
        }
    } else {
        // ---------------- A loader: quarter-warp = 8 consecutive rows x one 16-byte k-group --------------
        // smem: conflict-free 128-byte runs; global: 64 contiguous bytes per row per instruction pair
        const int r8 = lane & 7, gq = lane >> 3;
        for (int kc = 0; kc < a.n_kc; ++kc) {
            const int s = kc % PW_STAGES;
            const uint32_t ph = (kc / PW_STAGES) & 1u;
            mbar_wait(EMPTY(s), ph ^ 1u);
            const uint32_t xs = smem_u32(XS(s));
            const int ng = min(8, (a.k - kc * 64) / 8);   // valid 16-byte groups in this chunk
#pragma unroll
            for (int it = 0; it < 4; ++it) {
                const int r = warp * 32 + it * 8 + r8;
                const long long row = (long long)tile * PW_ROWS + r;
                const bool ok = row < a.rows;
                const __half *src = a.x + (size_t)(ok ? row : 0) * a.ldx;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const int g = gq + 4 * j;
                    if (g < ng) {
                        const int col = kc * 64 + g * 8;
                        const uint32_t dst = xs + (uint32_t)(r >> 3) * 1024u + (uint32_t)g * 128u + (uint32_t)(r & 7) * 16u;
                        cp_async16(dst, src + col, ok ? 16u : 0u);
                        if (SPLIT) cp_async16(dst + PW_TILE_BYTES, src + a.xlo + col, ok ? 16u : 0u);
                    }
                }
            }
            cp_async_commit();
            if (kc >= PW_LAG) {
                cp_async_wait<PW_LAG>();
                fence_proxy_async();
                mbar_arrive(FULL((kc - PW_LAG) % PW_STAGES));
            }
        }
        cp_async_wait<0>();
        fence_proxy_async();
        for (int kc = max(a.n_kc - PW_LAG, 0); kc < a.n_kc; ++kc) mbar_arrive(FULL(kc % PW_STAGES));

        // ---------------- epilogue: thread = row (TMEM lane) ----------------------------------------------
        mbar_wait(ACC_FULL, 0u);
        tc_fence_after();
        const int r = tid;
        const long long row = (long long)tile * PW_ROWS + r;
        const bool ok = row < a.rows;
        const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const int ch0 = cc * 128;
        long long bb = 0;
        int p = 0;
        if (a.out_cm && ok) { bb = row / a.m; p = (int)(row - bb * a.m); }
        float mx16 = 0.f;   // largest magnitude stored as fp16 by this thread
        for (int c0 = 0; c0 < ncols; c0 += 16) {
            float v[16];
            __syncwarp();
            tmem_ld16(taddr + (uint32_t)c0, v);   // warp-collective: executed by every lane, guards come after
            const float4 *b4 = reinterpret_cast<const float4 *>(a.bias + ch0 + c0);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 bv = __ldg(b4 + i);
                v[4 * i] += bv.x; v[4 * i + 1] += bv.y; v[4 * i + 2] += bv.z; v[4 * i + 3] += bv.w;
            }
            if (a.relu) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
            }
            if (ok && a.out_cm) {
                float *o = a.out_cm + ((size_t)bb * a.c_total + a.co_off + ch0 + c0) * a.m + p;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (ch0 + c0 + i < a.n) o[(size_t)i * a.m] = v[i];
            }
            if (ok && a.out16) {
                __half *o = a.out16 + (size_t)row * a.ld16 + ch0 + c0;
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (ch0 + c0 + 8 * q < a.n16) {
                        uint32_t h[4], l[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            mx16 = fmaxf(fmaxf(mx16, fabsf(v[8 * q + 2 * i])), fabsf(v[8 * q + 2 * i + 1]));
                            const __half2 hh = __floats2half2_rn(v[8 * q + 2 * i], v[8 * q + 2 * i + 1]);
                            const float2 hf = __half22float2(hh);
                            h[i] = *reinterpret_cast<const uint32_t *>(&hh);
                            l[i] = pack_h2(v[8 * q + 2 * i] - hf.x, v[8 * q + 2 * i + 1] - hf.y);
                        }
                        *reinterpret_cast<uint4 *>(o + 8 * q) = make_uint4(h[0], h[1], h[2], h[3]);
                        if (a.o16lo > 0) *reinterpret_cast<uint4 *>(o + a.o16lo + 8 * q) = make_uint4(l[0], l[1], l[2], l[3]);
                    }
            }
            if (ok && a.out_pm) {
                float *o = a.out_pm + (size_t)row * a.ldpm + ch0 + c0;
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    if (ch0 + c0 + i < a.n) o[i] = v[i];
            }
        }
        if (mx16 > FP16_MAX && a.ovf) atomicOr(a.ovf, a.ovf_bit);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 128);
    }
}

}  // namespace spsk

extern "C" int spsk_pw_mma_forward(const spsk_pw_desc *d, spsk_stream_t stream) {
    using namespace spsk;
    SPSK_REQUIRE(d, SPSK_ERR_INVALID_ARG, "pw_mma: null descriptor");
    SPSK_REQUIRE(d->rows >= 0 && d->n >= 1 && d->k >= 16 && d->k % 16 == 0 && d->ldx >= d->k && d->ldx % 8 == 0, SPSK_ERR_INVALID_ARG,
                 "pw_mma: bad sizes rows=%d k=%d ldx=%d n=%d (k multiple of 16, ldx multiple of 8)", d->rows, d->k, d->ldx, d->n);
    if (d->rows == 0) return SPSK_OK;
    SPSK_REQUIRE(d->x && d->wtiles && d->bias, SPSK_ERR_INVALID_ARG, "pw_mma: null input pointer");
    SPSK_REQUIRE(d->out_cm || d->out16 || d->out_pm, SPSK_ERR_INVALID_ARG, "pw_mma: no output requested");
    SPSK_REQUIRE((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->wtiles) & 15) == 0, SPSK_ERR_INVALID_ARG,
                 "pw_mma: x / wtiles must be 16-byte aligned");
    PwArgs a{};
    a.rows = d->rows; a.k = d->k; a.ldx = d->ldx; a.n = d->n; a.relu = d->relu ? 1 : 0;
    a.npad = (d->n + 15) / 16 * 16;
    a.split = d->split ? 1 : 0;
    a.xlo = d->xlo;
    a.n_kc = (d->k + 63) / 64;
    if (a.split)
        SPSK_REQUIRE(d->xlo >= d->k && d->xlo % 8 == 0 && d->xlo + d->k <= d->ldx, SPSK_ERR_INVALID_ARG,
                     "pw_mma: split input needs the lo part at a column xlo >= k, multiple of 8, inside ldx");
    a.x = reinterpret_cast<const __half *>(d->x);
    a.wtiles = reinterpret_cast<const __half *>(d->wtiles);
    a.bias = d->bias;
    a.out_cm = d->out_cm; a.m = d->m; a.c_total = d->c_total; a.co_off = d->co_off;
    if (a.out_cm)
        SPSK_REQUIRE(d->m >= 1 && d->rows % d->m == 0 && d->co_off >= 0 && d->co_off + d->n <= d->c_total, SPSK_ERR_INVALID_ARG,
                     "pw_mma: channel-major output needs rows %% m == 0 and a channel window inside c_total");
    a.out16 = reinterpret_cast<__half *>(d->out16); a.ld16 = d->ld16; a.n16 = d->n16; a.o16lo = d->o16lo;
    if (a.out16 && d->o16lo)
        SPSK_REQUIRE(d->o16lo >= d->n16 && d->o16lo % 8 == 0 && d->o16lo + d->n16 <= d->ld16, SPSK_ERR_INVALID_ARG,
                     "pw_mma: residual columns [o16lo, o16lo + n16) must follow the values inside ld16");
    if (a.out16)
        SPSK_REQUIRE(d->ld16 % 8 == 0 && d->n16 % 8 == 0 && d->n16 >= d->n && d->n16 <= d->ld16 &&
                         (reinterpret_cast<uintptr_t>(d->out16) & 15) == 0,
                     SPSK_ERR_INVALID_ARG, "pw_mma: fp16 output needs ld16, n16 multiples of 8, n <= n16 <= ld16, 16-byte alignment");
    a.out_pm = d->out_pm; a.ldpm = d->ldpm;
    a.ovf = fp16_overflow_word();
    a.ovf_bit = 1u << (d->ovf_tag & 31);
    if (a.out_pm) SPSK_REQUIRE(d->ldpm >= d->n, SPSK_ERR_INVALID_ARG, "pw_mma: ldpm < n");
    const bool deep = !a.split && a.n_kc >= 12;
    static SmemAttrOnce attr_s, attr_d, attr_x;
    if (a.split) {
        if (int rc = attr_x.ensure(reinterpret_cast<const void *>(pw_mma_kernel<3, 2, true>), pw_smem(3, true), "pw_mma_kernel<3,2,split>")) return rc;
    } else if (deep) {
        if (int rc = attr_d.ensure(reinterpret_cast<const void *>(pw_mma_kernel<6, 4, false>), pw_smem(6, false), "pw_mma_kernel<6,4>")) return rc;
    } else {
        if (int rc = attr_s.ensure(reinterpret_cast<const void *>(pw_mma_kernel<3, 2, false>), pw_smem(3, false), "pw_mma_kernel<3,2>")) return rc;
    }
    // n16 may extend past npad (zero columns up to the consumer's K padding): cover them with column tiles
    const int ncover = a.out16 ? max(a.npad, a.n16) : a.npad;
    a.npad = (ncover + 15) / 16 * 16;
    dim3 grid((unsigned)((d->rows + PW_ROWS - 1) / PW_ROWS), (unsigned)((a.npad + 127) / 128));
    if (a.split) pw_mma_kernel<3, 2, true><<<grid, PW_THREADS, pw_smem(3, true), as_stream(stream)>>>(a);
    else if (deep) pw_mma_kernel<6, 4, false><<<grid, PW_THREADS, pw_smem(6, false), as_stream(stream)>>>(a);
    else pw_mma_kernel<3, 2, false><<<grid, PW_THREADS, pw_smem(3, false), as_stream(stream)>>>(a);
    SPSK_LAUNCH_CHECK("pw_mma_kernel");
    return SPSK_OK;
}
