// sa_mma_pair.cu -- the fused set-abstraction scale for WIDE chains on a CTA PAIR (tcgen05 cta_group::2), sm_100a.
//
// Same computation and operand layouts as sa_mma.cu; different machine mapping.  On one CTA a wide chain (IA-SSD
// layers 2 and 5: 256..1024 channels) is bounded by (a) single-thread MMA issue -- ~140 cycles of issue overhead per
// 64-cycle 128x128x16 MMA -- and (b) weight streaming: the 128-row activations fill shared memory, leaving two 16 KB
// weight stages, and every CTA re-reads the whole chain's weights (1.4 MB for layer 5) per 128 rows.  A pair of CTAs
// (one cluster of 2 SMs) works on a 256-row tile instead:
//
//   hidden layers (orientation A):  D[256 rows, <=256 couts] = X . W^T      M = 256: each CTA supplies its 128 rows of X
//                                   (A operand) from its own shared memory and HALF of the cout chunk (B operand)
//   last layer    (orientation B):  D[256 couts, 256 rows]   = W . X^T      each CTA supplies its 128 couts of W (A) and
//                                   its 128 rows of X (B); its TMEM receives 128 couts x 256 rows, pooled per thread
//
// so each weight byte staged by a CTA serves 256 rows x 2 (4x less weight traffic per row per SM, and the same 16 KB
// stage now covers 256 couts), and one tcgen05.mma.cta_group::2 does the work of four single-CTA MMAs (4x fewer issues).
//
// Protocol.  Only the leader CTA (cluster rank 0) issues MMAs and commits; `tcgen05.commit ... multicast::cluster`
// arrives on the SAME barrier in both CTAs (weight stage empty, accumulator full, hidden-done), so producers and
// epilogues of both CTAs run exactly as in the single-CTA kernel.  What the leader must additionally know about its
// peer -- peer weight stage landed, peer activation chunk written, peer accumulator drained -- is forwarded by the
// peer's otherwise idle MMA warp: it walks the same schedule, waits on the peer-local barrier and performs ONE remote
// `mbarrier.arrive.release.cluster` on the leader's mirror barrier.
#include "sa_mma_common.cuh"
#include <stdlib.h>

namespace spsk {

constexpr int PR_HDR = 2048;        // barriers + TMEM slot
constexpr int PR_ROWS = 256;        // rows per pair tile
constexpr int PR_THREADS = 320;     // 2 epilogue warpgroups + producer + MMA / forwarder

__device__ __forceinline__ uint32_t pr_cta_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void pr_cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// one arrival on the barrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_bar), "r"(cta));
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(ra) : "memory");
}
// wait with cluster-scope acquire (the arrival came from the other CTA)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
        if (done) return;
    }
    __trap();
}
__device__ __forceinline__ void umma2_f16_lohi(uint32_t d_tmem, uint32_t alo, uint32_t ahi, uint32_t blo, uint32_t bhi, uint32_t idesc,
                                               uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %2};\n\t"
        "mov.b64 db, {%3, %4};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, {%7, %7, %7, %7, %7, %7, %7, %7}, p;\n\t}"
        ::"r"(d_tmem), "r"(alo), "r"(ahi), "r"(blo), "r"(bhi), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
}
// completion of every MMA issued so far -> the barrier at this offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma2_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}

// byte offset of CTA `rank`'s weight tile (cc, kc) of a layer in the PAIR packing: 256-wide cout chunks contiguous; inside a
// chunk rank 0's rows then rank 1's; inside a rank the tiles follow each other along K; a tile is `rows x kw` fp16
__device__ __forceinline__ int ptile_off(const SaLayer &Ly, int cc, int rank, int kc, int rows) {
    return Ly.w_off + (256 * cc * Ly.vk + rank * rows * Ly.vk + rows * 64 * kc) * 2;
}

__global__ void __launch_bounds__(PR_THREADS, 1)
sa_mma_pair_kernel(const __grid_constant__ SaArgs a) {
    constexpr int W_PROD = 8, W_MMA = 9;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + PR_HDR - 16);
    uint8_t *xa = smem + PR_HDR;
    uint8_t *xb = xa + a.xa_bytes;
    uint8_t *wst = xb + a.xb_bytes;
    const uint32_t bar0 = smem_u32(bars);
    // local barriers (both CTAs)
    auto W_FULL = [&](int s) { return bar0 + 8u * s; };
    auto W_EMPTY = [&](int s) { return bar0 + 8u * (8 + s); };
    auto WL_FULL = [&](int s) { return bar0 + 8u * (16 + s); };
    auto WL_EMPTY = [&](int s) { return bar0 + 8u * (24 + s); };
    auto ACC_FULL = [&](int i) { return bar0 + 8u * (32 + i); };
    auto ACC_EMPTY = [&](int i) { return bar0 + 8u * (34 + i); };
    auto XR = [&](int buf, int c) { return bar0 + 8u * (36 + buf * MM_MAX_XC + c); };
    const uint32_t HID_DONE = bar0 + 8u * 68;
    // mirrors of the PEER's barriers, used in the leader only (one forwarded arrival each)
    auto P_W_FULL = [&](int s) { return bar0 + 8u * (72 + s); };
    auto P_WL_FULL = [&](int s) { return bar0 + 8u * (80 + s); };
    auto P_ACC_EMPTY = [&](int i) { return bar0 + 8u * (88 + i); };
    auto P_XR = [&](int buf, int c) { return bar0 + 8u * (90 + buf * MM_MAX_XC + c); };

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int nL = a.nlayers;
    const uint32_t rank = pr_cta_rank();
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

    if (tid == 0) {
        for (int s = 0; s < 8; ++s) {
            mbar_init(W_FULL(s), 1); mbar_init(W_EMPTY(s), 1); mbar_init(WL_FULL(s), 1); mbar_init(WL_EMPTY(s), 1);
            mbar_init(P_W_FULL(s), 1); mbar_init(P_WL_FULL(s), 1);
        }
        for (int i = 0; i < 2; ++i) { mbar_init(ACC_FULL(i), 1); mbar_init(ACC_EMPTY(i), 128); mbar_init(P_ACC_EMPTY(i), 1); }
        for (int i = 0; i < 2 * MM_MAX_XC; ++i) { mbar_init(XR(0, i), 128); mbar_init(P_XR(0, i), 1); }
        mbar_init(HID_DONE, 1);
        mbar_init_fence();
    }
    if (warp == W_MMA) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    pr_cluster_sync();   // both CTAs' barriers are initialised before anyone signals across the pair
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    uint8_t *lst = ((nL - 2) & 1) ? xb : xa;   // overlay slots of the last layer's ring (lstages > 0)

    if (warp == W_PROD) {
        // ================= weight producer: this CTA's half of every tile =================
        const bool leader = elect_one();
        uint32_t tcount = 0;
        int ws = 0, ls = 0;
        uint32_t wph = 0u, lph = 0u;
        for (int tile = pair; tile < a.ntiles; tile += npairs, ++tcount) {
            for (int l = 0; l < nL; ++l) {
                const SaLayer &Ly = a.L[l];
                const bool last = l == nL - 1;
                const bool lring = a.lstages > 0 && last;
                if (lring) mbar_wait_cluster(HID_DONE, tcount & 1u);   // the overlaid activation buffer is dead from here on
                for (int cc = 0; cc < Ly.n_cc; ++cc) {
                    const int cw = min(256, Ly.cpad - cc * 256);
                    const int rows = last ? 128 : (cw >> 1);
                    for (int kc = 0; kc < Ly.n_kc; ++kc) {
                        const int kw = min(64, Ly.vk - kc * 64);
                        const uint32_t bytes = (uint32_t)(rows * kw * 2);
                        const uint8_t *src = a.wtiles + ptile_off(Ly, cc, (int)rank, kc, rows);
                        uint32_t full, empty, dst, ph;
                        if (!lring) {
                            ph = wph;
                            full = W_FULL(ws); empty = W_EMPTY(ws); dst = smem_u32(wst + (size_t)ws * MM_STAGE_BYTES);
                            if (++ws == a.nstages) { ws = 0; wph ^= 1u; }
                        } else {
                            ph = lph;
                            full = WL_FULL(ls); empty = WL_EMPTY(ls); dst = smem_u32(lst + (size_t)ls * MM_STAGE_BYTES);
                            if (++ls == a.lstages) { ls = 0; lph ^= 1u; }
                        }
                        mbar_wait_cluster(empty, ph ^ 1u);   // freed by the leader's multicast commit
                        if (leader) {
                            mbar_expect_tx(full, bytes);
                            bulk_g2s(dst, src, bytes, full);
                        }
                        __syncwarp();
                    }
                }
            }
        }
    } else if (warp == W_MMA) {
        // ================= leader: MMA issuer;  peer: forwarder of its barrier completions =================
        const bool leader = elect_one();
        const bool is_lead_cta = rank == 0;
        Prof pf;   // leader CTA: [mma_wait_w] local weights, [prod_wait_stage] peer weights, [mma_wait_x] local X, [prod_wait_hid] peer X
        pf.init(a.prof != nullptr && leader && is_lead_cta);
        const long long t_start = pf.now();
        uint32_t job = 0;
        int ws = 0, ls = 0;
        uint32_t wph = 0u, lph = 0u;
        uint32_t xph[2] = {0u, 0u};
        for (int tile = pair; tile < a.ntiles; tile += npairs) {
            for (int l = 0; l < nL; ++l) {
                const SaLayer &Ly = a.L[l];
                const bool last = (l == nL - 1);
                const bool lring = a.lstages > 0 && last;
                const int xbuf = l & 1;
                const uint32_t x_lo0 = umma_desc_lo(smem_u32(xbuf ? xb : xa), 128u);
                const uint32_t x_hi = umma_desc_hi((uint32_t)Ly.xw * 16u);
                int xwait = 0;
                for (int cc = 0; cc < Ly.n_cc; ++cc, ++job) {
                    const int cw = min(256, Ly.cpad - cc * 256);
                    const int buf = (int)(job & 1u);
                    const uint32_t use = job >> 1;
                    if (use > 0) {
                        const long long t0 = pf.now();
                        mbar_wait(ACC_EMPTY(buf), (use - 1) & 1u);
                        if (is_lead_cta) mbar_wait_cluster(P_ACC_EMPTY(buf), (use - 1) & 1u);
                        else if (leader) mbar_arrive_remote(P_ACC_EMPTY(buf), 0u);
                        pf.add(PF_MMA_ACC_EMPTY, t0);
                        tc_fence_after();
                    }
                    const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 256);
                    const uint32_t idesc = last ? umma_idesc(256, PR_ROWS) : umma_idesc(256, cw);
                    for (int kc = 0; kc < Ly.n_kc; ++kc) {
                        const int kw = min(64, Ly.vk - kc * 64);
                        uint32_t wbase, wempty;
                        if (lring) {
                            { const long long t0 = pf.now(); mbar_wait(WL_FULL(ls), lph); pf.add(PF_MMA_W_FULL, t0); }
                            if (is_lead_cta) { const long long t0 = pf.now(); mbar_wait_cluster(P_WL_FULL(ls), lph); pf.add(PF_PROD_W_EMPTY, t0); }
                            else if (leader) mbar_arrive_remote(P_WL_FULL(ls), 0u);
                            wbase = smem_u32(lst + (size_t)ls * MM_STAGE_BYTES);
                            wempty = WL_EMPTY(ls);
                            if (++ls == a.lstages) { ls = 0; lph ^= 1u; }
                        } else {
                            { const long long t0 = pf.now(); mbar_wait(W_FULL(ws), wph); pf.add(PF_MMA_W_FULL, t0); }
                            if (is_lead_cta) { const long long t0 = pf.now(); mbar_wait_cluster(P_W_FULL(ws), wph); pf.add(PF_PROD_W_EMPTY, t0); }
                            else if (leader) mbar_arrive_remote(P_W_FULL(ws), 0u);
                            wbase = smem_u32(wst + (size_t)ws * MM_STAGE_BYTES);
                            wempty = W_EMPTY(ws);
                            if (++ws == a.nstages) { ws = 0; wph ^= 1u; }
                        }
                        const int nk16 = kw >> 4;
                        if (cc == 0) {
                            const int need = (kc * 4 + nk16 - 1) >> 2;
                            while (xwait <= need) {
                                const uint32_t par = (xph[xbuf] >> xwait) & 1u;
                                { const long long t0 = pf.now(); mbar_wait(XR(xbuf, xwait), par); pf.add(PF_MMA_XR, t0); }
                                if (is_lead_cta) { const long long t0 = pf.now(); mbar_wait_cluster(P_XR(xbuf, xwait), par); pf.add(PF_PROD_HID, t0); }
                                else if (leader) mbar_arrive_remote(P_XR(xbuf, xwait), 0u);
                                xph[xbuf] ^= (1u << xwait);
                                ++xwait;
                            }
                        }
                        tc_fence_after();
                        if (is_lead_cta && leader) {
                            const long long t_i = pf.now();
                            const uint32_t w_lo = umma_desc_lo(wbase, 128u);
                            const uint32_t w_hi = umma_desc_hi((uint32_t)kw * 16u);
                            const uint32_t x_lo = x_lo0 + (uint32_t)kc * 64u;
                            for (int j = 0; j < nk16; ++j) {
                                const uint32_t acc = (kc | j) ? 1u : 0u;
                                if (!last) umma2_f16_lohi(d_tmem, x_lo + 16u * j, x_hi, w_lo + 16u * j, w_hi, idesc, acc);
                                else umma2_f16_lohi(d_tmem, w_lo + 16u * j, w_hi, x_lo + 16u * j, x_hi, idesc, acc);
                            }
                            pf.add(PF_MMA_ISSUE, t_i);
                            const long long t_c = pf.now();
                            umma2_commit(wempty);
                            if (kc == Ly.n_kc - 1) {
                                umma2_commit(ACC_FULL(buf));
                                if (a.lstages > 0 && l == nL - 2 && cc == Ly.n_cc - 1) umma2_commit(HID_DONE);
                            }
                            pf.add(PF_MMA_COMMIT, t_c);
                        }
                        __syncwarp();
                    }
                }
            }
        }
        pf.add(PF_MMA_TOTAL, t_start);
        pf.flush(a.prof);
    } else {
        // ================= gather + epilogue: two warpgroups, alternate jobs; thread = row / TMEM lane =================
        Prof pf;
        pf.init(a.prof != nullptr && tid == 0 && rank == 0);
        const long long t_start = pf.now();
        uint32_t job = 0;
        const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;
        const int r = tid & 127;
        const uint32_t grp = (uint32_t)(tid >> 7);
        const uint32_t row_off = (uint32_t)(r & 7) * 16u;
        uint32_t prev_last_job = 0;
        bool have_prev = false;
        int jn = 0;
        if (grp == 0 && pair < a.ntiles) {
            const long long grow = (long long)pair * PR_ROWS + rank * 128 + r;
            if (grow < a.rows) jn = __ldg(a.idx + grow);
        }
        for (int tile = pair; tile < a.ntiles; tile += npairs) {
            // ---- gather X0 (into XA) for this CTA's 128 rows of the pair tile -- warpgroup 0
            const long long t_g = pf.now();
            if (grp == 0 && have_prev) mbar_wait_cluster(ACC_FULL((int)(prev_last_job & 1u)), (prev_last_job >> 1) & 1u);
            if (grp == 0) {
                const SaLayer &L0 = a.L[0];
                const uint32_t sbo = (uint32_t)L0.xw * 16u;
                uint8_t *xrow = xa + (size_t)(r >> 3) * sbo + row_off;
                const long long grow = (long long)tile * PR_ROWS + rank * 128 + r;
                const bool ok = grow < a.rows;
                const int j = jn;
                {
                    const long long gnext = grow + (long long)npairs * PR_ROWS;
                    if (gnext < a.rows) jn = __ldg(a.idx + gnext);
                }
                long long q = 0;
                int bb = 0;
                if (ok) { q = grow >> a.ns_log2; bb = (int)(q / a.m); }
                float dx = 0.f, dy = 0.f, dz = 0.f;
                if (a.use_xyz && ok) {
                    const float *p = a.xyz + ((size_t)bb * a.n + j) * 3;
                    const float *ctr = a.new_xyz + (size_t)q * 3;
                    dx = __fsub_rn(__ldg(p), __ldg(ctr));
                    dy = __fsub_rn(__ldg(p + 1), __ldg(ctr + 1));
                    dz = __fsub_rn(__ldg(p + 2), __ldg(ctr + 2));
                }
                const uint4 zero = make_uint4(0, 0, 0, 0);
                const int nch = L0.kpad >> 3;
                const int fch = a.cpad8 >> 3;
                const uint4 *trow = (ok && fch) ? reinterpret_cast<const uint4 *>(a.twin + ((size_t)bb * a.n + j) * a.ldtwin) : nullptr;
                int c = 0;
                for (; c + 8 <= fch; c += 8) {
                    uint4 t[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) t[u] = ok ? __ldg(trow + c + u) : zero;
#pragma unroll
                    for (int u = 0; u < 8; ++u) *reinterpret_cast<uint4 *>(xrow + (size_t)(c + u) * 128) = t[u];
                }
                for (; c < fch; ++c) *reinterpret_cast<uint4 *>(xrow + (size_t)c * 128) = ok ? __ldg(trow + c) : zero;
                if (a.use_xyz) {
                    *reinterpret_cast<uint4 *>(xrow + (size_t)c * 128) = make_uint4(pack_h2(dx, dy), pack_h2(dz, 0.f), 0u, 0u);
                    ++c;
                }
                for (; c < nch; ++c) *reinterpret_cast<uint4 *>(xrow + (size_t)c * 128) = zero;
                fence_proxy_async();
                for (int ch = 0; ch < L0.n_xc; ++ch) mbar_arrive(XR(0, ch));
            }
            pf.add(PF_EPI_GATHER, t_g);
            // ---- hidden layers: D[row, <=256 couts] -> relu(D + bias) -> fp16 -> next X, 64 columns at a time
            for (int l = 0; l < nL - 1; ++l) {
                const SaLayer &Ly = a.L[l];
                const SaLayer &Ln = a.L[l + 1];
                const int obuf = (l + 1) & 1;
                uint8_t *xo = obuf ? xb : xa;
                const uint32_t sbo = (uint32_t)Ln.xw * 16u;
                uint8_t *xrow = xo + (size_t)(r >> 3) * sbo + row_off;
                const float *bias = a.bias + Ly.bias_off;
                for (int cc = 0; cc < Ly.n_cc; ++cc, ++job) {
                    if ((job & 1u) != grp) continue;
                    const int cw = min(256, Ly.cpad - cc * 256);
                    const int buf = (int)(job & 1u);
                    { const long long t0 = pf.now(); mbar_wait_cluster(ACC_FULL(buf), (job >> 1) & 1u); pf.add(PF_EPI_WAIT_HID, t0); }
                    tc_fence_after();
                    const long long t_w = pf.now();
                    const uint32_t taddr = tmem_base + lane_field + (uint32_t)(buf * 256);
                    float hmx = 0.f;   // fp16 range guard (see sa_mma.cu)
                    for (int h0 = 0; h0 < cw; h0 += 64) {
                        const int hend = min(cw, h0 + 64);
                        int c0 = h0;
                        for (; c0 + 32 <= hend; c0 += 32) {
                            float v[32];
                            tmem_ld32(taddr + (uint32_t)c0, v);
                            const int col = cc * 256 + c0;
                            store_hidden16(v, bias + col, xrow + (size_t)(col >> 3) * 128, hmx);
                            store_hidden16(v + 16, bias + col + 16, xrow + (size_t)((col >> 3) + 2) * 128, hmx);
                        }
                        if (c0 < hend) {
                            float v[16];
                            tmem_ld16(taddr + (uint32_t)c0, v);
                            const int col = cc * 256 + c0;
                            store_hidden16(v, bias + col, xrow + (size_t)(col >> 3) * 128, hmx);
                        }
                        fence_proxy_async();
                        mbar_arrive(XR(obuf, (cc * 256 + h0) >> 6));
                    }
                    if (hmx > FP16_MAX && a.ovf) atomicOr(a.ovf, a.ovf_bit);
                    tc_fence_before();
                    mbar_arrive(ACC_EMPTY(buf));
                    pf.add(PF_EPI_WORK_HID, t_w);
                }
            }
            // ---- last layer: D[cout, 256 rows]; thread = one of this CTA's 128 couts; max over each centre's nsample columns
            {
                const SaLayer &Ly = a.L[nL - 1];
                const float *bias = a.bias + Ly.bias_off;
                const long long q0 = ((long long)tile * PR_ROWS) >> a.ns_log2;
                const long long qmax = (long long)a.b * a.m;
                const int ns = a.nsample;
                const long long bb0 = q0 / a.m;
                const int p0 = (int)(q0 - bb0 * a.m);
                const size_t sstride = (size_t)a.c_total * a.m;
                for (int cc = 0; cc < Ly.n_cc; ++cc, ++job) {
                    if ((job & 1u) != grp) continue;
                    const int buf = (int)(job & 1u);
                    { const long long t0 = pf.now(); mbar_wait_cluster(ACC_FULL(buf), (job >> 1) & 1u); pf.add(PF_EPI_WAIT_POOL, t0); }
                    tc_fence_after();
                    const long long t_w = pf.now();
                    const int ch = cc * 256 + (int)rank * 128 + r;
                    const bool w32 = a.out != nullptr && ch < a.cout_last;
                    const bool w16 = a.out16 != nullptr && ch < a.n16;
                    PoolOut o;
                    o.bv = __ldg(bias + ch); o.w32 = w32; o.w16 = w16;
                    o.outc = a.out + ((size_t)bb0 * a.c_total + a.co_off + (w32 ? ch : 0)) * a.m + (size_t)p0;
                    o.out16 = a.out16 + (size_t)q0 * a.ld16 + a.co16 + (w16 ? ch : 0);
                    o.ld16 = a.ld16; o.o16lo = a.o16lo; o.q = q0; o.qmax = qmax; o.m = a.m; o.p = p0; o.scene_stride = sstride;
                    const uint32_t taddr = tmem_base + lane_field + (uint32_t)(buf * 256);
                    if (ns == 32) pool_chunk<32, PR_ROWS>(taddr, o);
                    else if (ns == 16) pool_chunk<16, PR_ROWS>(taddr, o);
                    else pool_chunk_any(taddr, o, ns, PR_ROWS);
                    if (o.mx16 > FP16_MAX && a.ovf) atomicOr(a.ovf, a.ovf_bit);
                    tc_fence_before();
                    mbar_arrive(ACC_EMPTY(buf));
                    pf.add(PF_EPI_WORK_POOL, t_w);
                }
            }
            prev_last_job = job - 1u;
            have_prev = true;
        }
        pf.add(PF_EPI_TOTAL, t_start);
        pf.flush(a.prof);
    }

    tc_fence_before();
    __syncthreads();
    pr_cluster_sync();   // the peer's shared memory / TMEM stay alive until every MMA of the pair has retired
    if (warp == W_MMA) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    }
}

}  // namespace spsk

// Pair launch.  Called by spsk_sa_mma_forward when the descriptor asks for the pair packing (d->pair != 0).
int spsk_sa_mma_pair_launch(const spsk::SaArgs &a, int smem_bytes, cudaStream_t st) {
    using namespace spsk;
    static SmemAttrOnce attr;
    if (int rc = attr.ensure(reinterpret_cast<const void *>(sa_mma_pair_kernel), 227 * 1024, "sa_mma_pair_kernel")) return rc;
    const int max_pairs = (SPSK_NUM_SMS / 2) * 2;   // two pairs per SM-pair slot (static striding, as in sa_mma.cu)
    const int npairs = a.ntiles < max_pairs ? a.ntiles : max_pairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * npairs));
    cfg.blockDim = dim3(PR_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, sa_mma_pair_kernel, a);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelEx(sa_mma_pair_kernel)");
    SPSK_LAUNCH_CHECK("sa_mma_pair_kernel");
    return SPSK_OK;
}
