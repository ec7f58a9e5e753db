// sa_mma.cu -- fused set-abstraction scale on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
//
// ONE kernel per MSG scale replaces the reference's chain (pointnet2_utils.py:307-315,
// pointnet2_modules.py:204-211,431-436):
//     grouping_operation(xyz) - new_xyz ; grouping_operation(features) ; torch.cat ;
//     3 x [Conv2d 1x1 (no bias) -> BatchNorm2d(eval) -> ReLU] ; F.max_pool2d over nsample
// The (B,3+C,npoint,nsample) grouped tensor and all conv/BN/ReLU intermediates stay on chip.
//
// Machine mapping (persistent CTAs: 192 threads x 1-3 per SM, 160 threads x 4 per SM for resident narrow chains, 320 threads x 1
// for the wide streaming chains, depending on the chain's footprint):
//   tile        = 128 grouped rows (= 128/nsample centres), looped over by each CTA
//   job         = (layer, 128-wide cout chunk): one accumulator of <=128 TMEM columns; jobs of a tile run in
//                 layer order through a ring of `nbuf` accumulators, so the MMAs of job j+1.. overlap the
//                 epilogue of job j
//   warps 0-3   : (a) gather: row r = thread r builds X0[r, :] = [features(idx) | xyz(idx) - centre | 0] as fp16
//                     with 16-byte loads from the point-major fp16 feature twin and 16-byte smem stores;
//                 (b) hidden-layer epilogue: TMEM -> registers (tcgen05.ld), + folded-BN bias, ReLU, -> fp16 operand
//                     of the next layer in shared memory, signalled per 64-wide K chunk so the next layer's MMAs
//                     start before the whole activation is written;
//                 (c) last layer: max over nsample + bias + ReLU -> global (fp32 channel-major and/or fp16 point-major)
//   warp 4      : weight producer: 1-D bulk async copies (cp.async.bulk + mbarrier complete_tx) of host-packed
//                 weight tiles, either once (chain resident in shared memory) or through a ring of 16 / 32 KB slots
//                 following the host-built static schedule (SaArgs::ring)
//   warp 5      : TMEM allocator + single-thread tcgen05.mma issuer (kind::f16, fp16 operands, fp32 accumulate); three
//                 loops: register-resident descriptors (narrow resident chains), the tabulated static-ring schedule
//                 (streaming chains), a general fallback
//   statistics mode (SaArgs::stats, training-mode BatchNorm): the last layer's epilogue sums z and z^2 per cout instead
//                 of bias + ReLU + max-pool (include/spsk.h, train_fused.py)
//   hidden layers  (orientation A): D[row, cout]  = X[row, k] . W[cout, k]^T   M = 128 rows,  N = cout chunk
//   last layer     (orientation B): D[cout, row]  = W[cout, k] . X[row, k]^T   M = 128 couts, N = 128 rows
//                 so that the max over the nsample rows of a centre is a per-thread loop over TMEM columns.
//   All operands use the canonical K-major, no-swizzle UMMA layout (8-row x 16-byte core matrices):
//       byte(r, k) = (r/8)*SBO + (k/8)*128 + (r%8)*16 + (k%8)*2          LBO = 128
//
// Numerics.  Default: operands are fp32 values rounded to fp16 (11-bit significand, same as TF32), products are
// exact and accumulate in fp32; measured end-to-end error of a 3-layer scale is ~5e-4 of the output range.
// `split` mode (narrow chains, every K <= 64, i.e. IA-SSD layer 0 where the error does not average out): every
// operand is carried as hi + lo fp16 halves and each product is evaluated as Xh.Wh + Xl.Wh + Xh.Wl by
// walking K three times ([Xh | Xl | Xh] against the packed [Wh ; Wl], Wh serving two products) -- fp32-grade results (~1e-6) for 3x MMA work that
// these layers do not notice.  The exact-fp32 CUDA-core path is linear_ffma.cu.
#include "sa_mma_common.cuh"
#include <stdlib.h>

namespace spsk {

// ---- the kernel -----------------------------------------------------------------------------------
// G = epilogue warpgroups.  G = 1 (192 threads, up to 3 CTAs per SM) for chains whose concurrency comes from co-resident
// CTAs; G = 2 (320 threads, one CTA per SM) for the wide chains: the two warpgroups take alternate jobs, so two
// accumulators drain concurrently and every scheduler holds two epilogue warps to hide TMEM / shared-memory latency.
// NP = no producer warp (resident chains only: the one bulk load of the chain is issued by the MMA warp): 160-thread CTAs, FOUR
// per SM at 96 registers.  The narrow chains are bound by the latency of their gather -> MMA -> epilogue hand-offs, i.e. by
// how many tiles an SM keeps in flight; a fourth CTA is a fourth tile.
template <int G, bool PROF, bool NP>
__global__ void __launch_bounds__(128 * G + 64 - (NP ? 32 : 0), G == 1 ? (NP ? 4 : 3) : 1)
sa_mma_kernel(const __grid_constant__ SaArgs a) {
    constexpr int W_PROD = NP ? -1 : 4 * G, W_MMA = NP ? 4 * G : 4 * G + 1;
    extern __shared__ __align__(128) uint8_t smem[];
    // carve: [header: barriers + tmem slot][XA][XB][weights]
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + MM_HDR - 16);
    uint8_t *xa = smem + MM_HDR;
    uint8_t *xb = xa + a.xa_bytes;
    uint8_t *wst = xb + a.xb_bytes;
    const uint32_t bar0 = smem_u32(bars);
    auto W_FULL = [&](int s) { return bar0 + 8u * s; };
    auto W_EMPTY = [&](int s) { return bar0 + 8u * (MM_MAX_STAGES + s); };
    auto ACC_FULL = [&](int i) { return bar0 + 8u * (2 * MM_MAX_STAGES + i); };
    auto ACC_EMPTY = [&](int i) { return bar0 + 8u * (2 * MM_MAX_STAGES + 4 + i); };
    auto XR = [&](int buf, int c) { return bar0 + 8u * (2 * MM_MAX_STAGES + 8 + buf * MM_MAX_XC + c); };
    auto WL_FULL = [&](int s) { return bar0 + 8u * (2 * MM_MAX_STAGES + 8 + 2 * MM_MAX_XC + s); };
    auto WL_EMPTY = [&](int s) { return bar0 + 8u * (3 * MM_MAX_STAGES + 8 + 2 * MM_MAX_XC + s); };
    const uint32_t HID_DONE = bar0 + 8u * (4 * MM_MAX_STAGES + 8 + 2 * MM_MAX_XC);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int nL = a.nlayers;
    // The last layer's cout chunks are independent jobs: every CTA walks them from a different starting chunk, so that
    // at any moment the CTAs stream DIFFERENT weight tiles (148 SMs requesting the same 16 KB in lock-step serialise on
    // the L2 slices holding those lines).
    const int rot_last = a.rot_last ? (int)(blockIdx.x % (unsigned)a.L[nL - 1].n_cc) : 0;
    auto chunk_of = [&](int l, int cci, int n_cc) { int c = cci + (l == nL - 1 ? rot_last : 0); return c >= n_cc ? c - n_cc : c; };

    if (tid == 0) {
        for (int s = 0; s < MM_MAX_STAGES; ++s) { mbar_init(W_FULL(s), 1); mbar_init(W_EMPTY(s), 1); }
        for (int i = 0; i < 4; ++i) { mbar_init(ACC_FULL(i), 1); mbar_init(ACC_EMPTY(i), 128); }
        for (int i = 0; i < 2 * MM_MAX_XC; ++i) mbar_init(XR(0, i), 128);
        for (int s = 0; s < MM_MAX_STAGES; ++s) { mbar_init(WL_FULL(s), 1); mbar_init(WL_EMPTY(s), 1); }
        mbar_init(HID_DONE, 1);
        mbar_init_fence();
    }
    if (warp == W_MMA) tmem_alloc(smem_u32(tmem_slot), (uint32_t)a.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == W_PROD) {
        // ================= weight producer (warp-uniform loops, one elected lane issues) =================
        const bool leader = elect_one();
        if (a.resident) {
            // the whole packed chain becomes resident: shared-memory image == global image
            if (leader) {
                mbar_expect_tx(W_FULL(0), (uint32_t)a.w_total);
                for (int off = 0; off < a.w_total; off += MM_STAGE_BYTES) {
                    const int bytes = min(MM_STAGE_BYTES, a.w_total - off);
                    bulk_g2s(smem_u32(wst + off), a.wtiles + off, (uint32_t)bytes, W_FULL(0));
                }
            }
        } else if (a.sched_n > 0) {
            // table-driven (see the tabulated issue loop): stage, barriers and phase of every weight tile are static per entry
            ProfT<PROF> pf;
            pf.init(a.prof != nullptr && leader);
            const int nent = a.sched_n;
            uint32_t tpar = 0u, tcount = 0u;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, tpar ^= 1u, ++tcount) {
                for (int e = 0; e < nent; ++e) {
                    const uint4 R = a.ring[e];
                    if (a.sched[e].w & SCH_LRING_FIRST) { const long long t0 = pf.now(); mbar_wait(HID_DONE, tpar); pf.add(PF_PROD_HID, t0); }
                    { const long long t0 = pf.now(); mbar_wait(bar0 + ((R.z >> 10) & 1023u), ((R.z >> 20) ^ ((R.z >> 21) & tpar) ^ 1u) & 1u); pf.add(PF_PROD_W_EMPTY, t0); }
                    if (leader) {
                        const uint32_t full = bar0 + (R.z & 1023u), bytes = R.y & 0xFFFFu;
                        if (bytes == 0u || (PROF && (a.abl & 1) && tcount > 0u)) {
                            mbar_arrive(full);   // padding entry (keeps the ring position static), or ablation: stage "lands" without a copy
                        } else {
                            mbar_expect_tx(full, bytes);
                            bulk_g2s(bar0 + ((R.y >> 16) << 4), a.wtiles + R.x, bytes, full);
                        }
                    }
                    __syncwarp();
                }
            }
            pf.flush(a.prof);
        } else {
            ProfT<PROF> pf;
            pf.init(a.prof != nullptr && leader);
            uint32_t tcount = 0;
            int ws = 0, ls = 0;
            uint32_t wph = 0u, lph = 0u;
            uint8_t *lst = ((nL - 2) & 1) ? xb : xa;   // overlay slots of the last layer's ring (lstages > 0)
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, ++tcount) {
                for (int l = 0; l < nL; ++l) {
                    const SaLayer &Ly = a.L[l];
                    const bool lring = a.lstages > 0 && l == nL - 1;
                    if (lring) { const long long t0 = pf.now(); mbar_wait(HID_DONE, tcount & 1u); pf.add(PF_PROD_HID, t0); }   // the overlaid activation buffer is dead from here on
                    for (int cci = 0; cci < Ly.n_cc; ++cci) {
                        const int cc = chunk_of(l, cci, Ly.n_cc);
                        const int ncols = min(128, Ly.cpad - cc * 128);
                        for (int kc = 0; kc < Ly.n_kc; ++kc) {
                            const int kw = min(64, Ly.wk - kc * 64);
                            const uint32_t bytes = (uint32_t)(ncols * kw * 2);
                            const uint8_t *src = a.wtiles + wtile_off(Ly, cc, kc, ncols);
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
                            { const long long t0 = pf.now(); mbar_wait(empty, ph ^ 1u); pf.add(PF_PROD_W_EMPTY, t0); }
                            if (leader) {
                                mbar_expect_tx(full, bytes);
                                bulk_g2s(dst, src, bytes, full);
                            }
                            __syncwarp();
                        }
                    }
                }
            }
            pf.flush(a.prof);
        }
    } else if (warp == W_MMA) {
        // ================= MMA issuer (warp-uniform loops, one elected lane issues) =================
        if (a.narrow) {
            // Narrow chains (weights resident, one job per layer): every descriptor is tile-invariant, so they are built ONCE
            // into registers (loops over layers fully unrolled -> static indexing) and the per-job path is just: wait for the
            // accumulator, wait for the activation chunks, issue, commit.  The general loop below spends ~2 k cycles of
            // address arithmetic and parameter loads per job, which is the critical path of these latency-bound chains.
            const bool leader = elect_one();
            if (NP && leader) {   // no producer warp: the resident chain is loaded from here
                mbar_expect_tx(W_FULL(0), (uint32_t)a.w_total);
                for (int off = 0; off < a.w_total; off += MM_STAGE_BYTES) {
                    const int bytes = min(MM_STAGE_BYTES, a.w_total - off);
                    bulk_g2s(smem_u32(wst + off), a.wtiles + off, (uint32_t)bytes, W_FULL(0));
                }
            }
            uint32_t x_lo[MM_MAX_LAYERS], x_hi[MM_MAX_LAYERS], w_lo[MM_MAX_LAYERS], idesc[MM_MAX_LAYERS], tstride[MM_MAX_LAYERS];
            int nv[MM_MAX_LAYERS], nk1[MM_MAX_LAYERS], nk2[MM_MAX_LAYERS], nxc[MM_MAX_LAYERS], wkl[MM_MAX_LAYERS];
#pragma unroll
            for (int l = 0; l < MM_MAX_LAYERS; ++l) {
                const SaLayer &Ly = a.L[l < nL ? l : 0];
                const bool last = (l == nL - 1);
                x_lo[l] = umma_desc_lo(smem_u32((l & 1) ? xb : xa), 128u);
                x_hi[l] = umma_desc_hi((uint32_t)Ly.xw * 16u);
                w_lo[l] = umma_desc_lo(smem_u32(wst + Ly.w_off), 128u);
                tstride[l] = (uint32_t)(Ly.cpad * 64 * 2) >> 4;            // one full 64-k tile of this layer, in descriptor units
                idesc[l] = last ? umma_idesc(128, MM_ROWS) : umma_idesc(128, Ly.cpad);
                nv[l] = Ly.vk >> 4;
                nk1[l] = a.split ? (Ly.kpad >> 4) : 0x7FFFFFFF;       // split: K blocks [0, nk1) = Xh.Wh, [nk1, nk2) = Xl.Wh, [nk2, nv) = Xh.Wl
                nk2[l] = a.split ? 2 * (Ly.kpad >> 4) : 0x7FFFFFFF;
                nxc[l] = Ly.n_xc;
                wkl[l] = Ly.wk;
            }
            mbar_wait(W_FULL(0), 0u);   // the resident weights have landed
            uint32_t job = 0;
            uint32_t xph[2] = {0u, 0u};
            const uint32_t nbmask = (uint32_t)a.nbuf - 1u;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
#pragma unroll
                for (int l = 0; l < MM_MAX_LAYERS; ++l) {
                    if (l < nL && l >= a.l0_fused) {
                        const bool last = (l == nL - 1);
                        const int buf = (int)(job & nbmask);
                        const uint32_t use = job >> a.nbuf_log2;
                        if (use > 0) mbar_wait(ACC_EMPTY(buf), (use - 1) & 1u);
                        for (int c = 0; c < nxc[l]; ++c) {
                            mbar_wait(XR(l & 1, c), (xph[l & 1] >> c) & 1u);
                            xph[l & 1] ^= (1u << c);
                        }
                        tc_fence_after();
                        if (leader) {
                            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 128);
                            for (int v = 0; v < nv[l]; ++v) {
                                const int xblk = v >= nk2[l] ? v - nk2[l] : v;       // [hi | lo | hi]
                                const int wblk = v >= nk1[l] ? v - nk1[l] : v;       // [Wh | Wh | Wl] out of the packed [Wh ; Wl]
                                const int kc = wblk >> 2, jj = wblk & 3;
                                const uint32_t kw = (uint32_t)min(64, wkl[l] - kc * 64);
                                const uint32_t wl = w_lo[l] + (uint32_t)kc * tstride[l] + 16u * (uint32_t)jj;
                                const uint32_t wh = umma_desc_hi(kw * 16u);
                                const uint32_t xl = x_lo[l] + 16u * (uint32_t)xblk;
                                if (!last) umma_f16_lohi(d_tmem, xl, x_hi[l], wl, wh, idesc[l], v ? 1u : 0u);
                                else umma_f16_lohi(d_tmem, wl, wh, xl, x_hi[l], idesc[l], v ? 1u : 0u);
                            }
                            umma_commit(ACC_FULL(buf));
                        }
                        __syncwarp();
                        ++job;
                    }
                }
            }
        } else if (a.sched_n > 0) {
            // Streaming chains: the per-tile schedule is identical for every tile and every CTA, so the HOST tabulates it into
            // the kernel parameters (SaArgs::sched / ::ring, see build_schedule) and this loop only interprets it.  What the
            // ablations of round 2 showed (profiles/r02_sa_mma_ablations.txt): with the weight copies removed, or the hidden
            // epilogues, the kernel time did not move, and with THREE OF FOUR MMAs removed it dropped by 15 % -- the chain was
            // bound by this single warp's instruction latency per weight tile (ring bookkeeping, barrier addresses, phase bits
            // kept in local memory, flag decoding: ~450 cycles around four 66-cycle MMAs), not by the tensor pipe.  So everything
            // that can be static IS static: the ring slot of every weight tile is fixed by padding each ring's entries per tile
            // to a multiple of its depth (padding entries: the producer arrives without copying, this loop only releases the
            // slot), which makes the slot address, both barrier addresses and the phase parity table constants -- parity =
            // static bit ^ (tile parity & "odd number of uses per tile" bit).  Parameters are read with uniform loads straight
            // into the uniform registers tcgen05.mma takes its operands from.
            //   sched[e]: .x activation descriptor lo of the entry's first K block, relative to the CTA's shared memory (16 B units)
            //             .y instruction descriptor   .z activation desc hi | weight desc hi << 16   .w SCH_* flags | MMAs | 2nd-tile offset
            //   ring[e] : .x byte offset of the weight tile(s) in the packed weights   .y bytes (0 = padding) | slot offset / 16 << 16
            //             .z full barrier | empty barrier << 10 | parity << 20 | parity tile-dependent << 21 | up to two activation
            //                waits: valid / parity / tile-dependent at bits 22..24 and 25..27   .w their barriers (10 bits each; offsets from bar0)
            // An entry covers ONE 64-wide k tile (16 KB slot) or, where the ring has room for >= 2 slots of 32 KB, TWO consecutive
            // ones (one bulk copy, 8 MMAs): the waits, the commit and this loop's latency are paid per entry, the MMAs are not.
            const bool leader = elect_one();
            ProfT<PROF> pf;
            pf.init(a.prof != nullptr && leader);
            const long long t_start = pf.now();
            uint32_t job = 0, tpar = 0u;
            const uint32_t nbmask = (uint32_t)a.nbuf - 1u;
            const int nent = a.sched_n;
            const uint32_t smem_base16 = bar0 >> 4;   // descriptor units
            uint4 E = a.sched[0], R = a.ring[0];
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x, tpar ^= 1u) {
                int buf = 0;
                uint32_t d_tmem = tmem_base;
                for (int e = 0; e < nent; ++e) {
                    const int en = e + 1 < nent ? e + 1 : 0;
                    const uint4 En = a.sched[en], Rn = a.ring[en];   // next entry: its constant-bank latency hides under this entry's waits
                    const uint32_t f = E.w;
                    if (f & SCH_FIRST_KC) {
                        buf = (int)(job & nbmask);
                        const uint32_t use = job >> a.nbuf_log2;
                        if (use > 0) { const long long t0 = pf.now(); mbar_wait(ACC_EMPTY(buf), (use - 1) & 1u); pf.add(PF_MMA_ACC_EMPTY, t0); }
                        d_tmem = tmem_base + (uint32_t)(buf * 128);
                    }
                    { const long long t0 = pf.now(); mbar_wait(bar0 + (R.z & 1023u), ((R.z >> 20) ^ ((R.z >> 21) & tpar)) & 1u); pf.add(PF_MMA_W_FULL, t0); }
                    if (R.z & (1u << 22)) {
                        const long long t0 = pf.now();
                        mbar_wait(bar0 + (R.w & 1023u), ((R.z >> 23) ^ ((R.z >> 24) & tpar)) & 1u);
                        if (R.z & (1u << 25)) mbar_wait(bar0 + (R.w >> 10), ((R.z >> 26) ^ ((R.z >> 27) & tpar)) & 1u);
                        pf.add(PF_MMA_XR, t0);
                    }
                    tc_fence_after();
                    if (leader) {
                        const long long t_i = pf.now();
                        const int nk16 = (PROF && (a.abl & 8) && (f & 15u)) ? 1 : (int)(f & 15u);   // 0 for padding entries
                        if (nk16) {
                            const uint32_t x_lo = E.x + smem_base16;
                            const uint32_t w_lo = (((smem_base16 + (R.y >> 16)) & 0x3FFFu) | ((128u >> 4) << 16));   // umma_desc_lo(slot, LBO 128)
                            const uint32_t w_lo2 = w_lo + ((f >> 16) & 2047u) - 64u;                                  // second k tile, minus its j offset
                            const uint32_t x_hi = E.z & 0xFFFFu, w_hi = E.z >> 16;
                            const uint32_t acc0 = (f & SCH_FIRST_KC) ? 0u : 1u;
                            const int n1 = nk16 < 4 ? nk16 : 4;
                            if (!(f & SCH_LAST_LAYER)) {
                                umma_f16_lohi(d_tmem, x_lo, x_hi, w_lo, w_hi, E.y, acc0);
                                for (int j = 1; j < n1; ++j) umma_f16_lohi(d_tmem, x_lo + 16u * j, x_hi, w_lo + 16u * j, w_hi, E.y, 1u);
                                for (int j = 4; j < nk16; ++j) umma_f16_lohi(d_tmem, x_lo + 16u * j, x_hi, w_lo2 + 16u * j, w_hi, E.y, 1u);
                            } else {
                                umma_f16_lohi(d_tmem, w_lo, w_hi, x_lo, x_hi, E.y, acc0);
                                for (int j = 1; j < n1; ++j) umma_f16_lohi(d_tmem, w_lo + 16u * j, w_hi, x_lo + 16u * j, x_hi, E.y, 1u);
                                for (int j = 4; j < nk16; ++j) umma_f16_lohi(d_tmem, w_lo2 + 16u * j, w_hi, x_lo + 16u * j, x_hi, E.y, 1u);
                            }
                        }
                        pf.add(PF_MMA_ISSUE, t_i);
                        const long long t_c = pf.now();
                        umma_commit(bar0 + ((R.z >> 10) & 1023u));   // slot reusable once these MMAs retire
                        if (f & SCH_LAST_KC) {
                            umma_commit(ACC_FULL(buf));
                            if (f & SCH_HID_DONE) umma_commit(HID_DONE);
                        }
                        pf.add(PF_MMA_COMMIT, t_c);
                    }
                    __syncwarp();
                    if (f & SCH_LAST_KC) ++job;
                    E = En; R = Rn;
                }
            }
            pf.add(PF_MMA_TOTAL, t_start);
            pf.flush(a.prof);
        } else {
            const bool leader = elect_one();
            ProfT<PROF> pf;
            pf.init(a.prof != nullptr && leader);
            const long long t_start = pf.now();
            uint32_t job = 0;
            int ws = 0, ls = 0;             // ring positions (hidden / last-layer ring) ...
            uint32_t wph = 0u, lph = 0u;    // ... and their phase parities
            uint32_t xph[2] = {0u, 0u};     // phase parity bit per readiness chunk of XA / XB
            bool first = true;
            uint8_t *lst = ((nL - 2) & 1) ? xb : xa;
            const uint32_t nbmask = (uint32_t)a.nbuf - 1u;
            for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
                for (int l = a.l0_fused; l < nL; ++l) {
                    const SaLayer &Ly = a.L[l];
                    const bool last = (l == nL - 1);
                    const bool lring = a.lstages > 0 && last;
                    const int xbuf = l & 1;
                    const uint32_t x_lo0 = umma_desc_lo(smem_u32(xbuf ? xb : xa), 128u);
                    const uint32_t x_hi = umma_desc_hi((uint32_t)Ly.xw * 16u);
                    const int nk1 = Ly.kpad >> 4;         // split: weight blocks [0, nk1) = Wh (times Xh and Xl), [nk1, 2 nk1) = Wl (times Xh)
                    int xwait = 0;                        // readiness chunks of this layer's input already waited for
                    for (int cci = 0; cci < Ly.n_cc; ++cci, ++job) {
                        const int cc = chunk_of(l, cci, Ly.n_cc);
                        const int ncols = min(128, Ly.cpad - cc * 128);  // couts in this chunk (multiple of 16)
                        const int buf = (int)(job & nbmask);
                        const uint32_t use = job >> a.nbuf_log2;
                        if (use > 0) { const long long t0 = pf.now(); mbar_wait(ACC_EMPTY(buf), (use - 1) & 1u); pf.add(PF_MMA_ACC_EMPTY, t0); tc_fence_after(); }
                        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * 128);
                        const uint32_t idesc = last ? umma_idesc(128, MM_ROWS) : umma_idesc(128, ncols);
                        for (int kc = 0; kc < Ly.n_kc; ++kc) {
                            const int kw = min(64, Ly.wk - kc * 64);
                            uint32_t wbase, wempty = 0u;
                            if (a.resident) {
                                if (first) { mbar_wait(W_FULL(0), 0u); first = false; }
                                wbase = smem_u32(wst + wtile_off(Ly, cc, kc, ncols));
                            } else if (lring) {
                                { const long long t0 = pf.now(); mbar_wait(WL_FULL(ls), lph); pf.add(PF_MMA_W_FULL, t0); }
                                wbase = smem_u32(lst + (size_t)ls * MM_STAGE_BYTES);
                                wempty = WL_EMPTY(ls);
                                if (++ls == a.lstages) { ls = 0; lph ^= 1u; }
                            } else {
                                { const long long t0 = pf.now(); mbar_wait(W_FULL(ws), wph); pf.add(PF_MMA_W_FULL, t0); }
                                wbase = smem_u32(wst + (size_t)ws * MM_STAGE_BYTES);
                                wempty = W_EMPTY(ws);
                                if (++ws == a.nstages) { ws = 0; wph ^= 1u; }
                            }
                            const int nk16 = kw >> 4;
                            if (cci == 0) {
                                // the activation chunks this weight tile touches must have landed (epilogue of layer l-1 / gather)
                                int need = (kc * 4 + nk16 - 1) >> 2;
                                if (a.split) need = Ly.n_xc - 1;   // hi | lo | hi re-read: simply wait for the whole (narrow) operand
                                while (xwait <= need) {
                                    const long long t0 = pf.now();
                                    mbar_wait(XR(xbuf, xwait), (xph[xbuf] >> xwait) & 1u);
                                    pf.add(PF_MMA_XR, t0);
                                    xph[xbuf] ^= (1u << xwait);
                                    ++xwait;
                                }
                            }
                            tc_fence_after();
                            if (leader) {
                                const long long t_i = pf.now();
                                const uint32_t w_lo = umma_desc_lo(wbase, 128u);
                                const uint32_t w_hi = umma_desc_hi((uint32_t)kw * 16u);
                                const uint32_t x_lo = x_lo0 + (uint32_t)kc * 64u;   // 4 K blocks of 256 bytes (>>4: 16 each)
                                if (nk16 == 4 && !a.split) {
                                    if (!last) {
#pragma unroll
                                        for (int j = 0; j < 4; ++j)
                                            umma_f16_lohi(d_tmem, x_lo + 16u * j, x_hi, w_lo + 16u * j, w_hi, idesc, (kc | j) ? 1u : 0u);
                                    } else {
#pragma unroll
                                        for (int j = 0; j < 4; ++j)
                                            umma_f16_lohi(d_tmem, w_lo + 16u * j, w_hi, x_lo + 16u * j, x_hi, idesc, (kc | j) ? 1u : 0u);
                                    }
                                } else {
                                    for (int j = 0; j < nk16; ++j) {
                                        const int wb = kc * 4 + j;   // K block of the packed weights
                                        const uint32_t wl = w_lo + 16u * (uint32_t)j;
                                        auto issue = [&](int xblk, uint32_t acc) {
                                            const uint32_t xl = x_lo0 + 16u * (uint32_t)xblk;
                                            if (!last) umma_f16_lohi(d_tmem, xl, x_hi, wl, w_hi, idesc, acc);
                                            else umma_f16_lohi(d_tmem, wl, w_hi, xl, x_hi, idesc, acc);
                                        };
                                        if (!a.split) issue(wb, wb ? 1u : 0u);
                                        else if (wb < nk1) { issue(wb, wb ? 1u : 0u); issue(nk1 + wb, 1u); }   // Xh.Wh + Xl.Wh: one Wh block, two products
                                        else issue(wb - nk1, 1u);                                              // Xh.Wl
                                    }
                                }
                                pf.add(PF_MMA_ISSUE, t_i);
                                const long long t_c = pf.now();
                                if (!a.resident) umma_commit(wempty);   // stage reusable once these MMAs retire
                                if (kc == Ly.n_kc - 1) {
                                    umma_commit(ACC_FULL(buf));
                                    if (a.lstages > 0 && l == nL - 2 && cci == Ly.n_cc - 1) umma_commit(HID_DONE);   // its input buffer may now hold weight tiles
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
        }
    } else {
        // ================= gather + epilogue (threads 0..127; thread = row / TMEM lane) =================
        ProfT<PROF> pf;
        pf.init(a.prof != nullptr && tid == 0);
        const long long t_start = pf.now();
        uint32_t job = 0;
        const uint32_t lane_field = (uint32_t)((warp & 3) * 32) << 16;   // a warp reads the TMEM lane quarter warp % 4
        const int r = tid & 127;
        const uint32_t grp = (uint32_t)(tid >> 7);                       // epilogue warpgroup: takes jobs with job % G == grp
        const uint32_t row_off = (uint32_t)(r & 7) * 16u;
        const int first_tile = blockIdx.x;
        uint32_t prev_last_job = 0;   // G = 2: the previous tile's last job (its MMAs are the last readers of XA)
        bool have_prev = false;
        if (a.l0_fused) mbar_wait(W_FULL(0), 0u);   // layer 0's fp32 weights live in the resident weight image
        int jn = 0;   // neighbour index of this thread's row in the NEXT tile (prefetched one tile ahead)
        if (grp == 0 && first_tile < a.ntiles) {
            const long long grow = (long long)first_tile * MM_ROWS + r;
            if (grow < a.rows) jn = __ldg(a.idx + grow);
        }
        for (int tile = first_tile; tile < a.ntiles; tile += gridDim.x) {
            // ---- gather X0 (into XA): [features | dxyz | 0]   (split: [hi(k0) | lo(k0)])   -- warpgroup 0
            const long long t_g = pf.now();
            if (G == 2 && grp == 0 && have_prev) {
                // the other warpgroup may own the previous tile's last job: peek at its accumulator barrier so that every
                // MMA reading the buffer this gather overwrites has retired (waits do not consume the phase)
                const int pb = (int)(prev_last_job & ((uint32_t)a.nbuf - 1u));
                mbar_wait(ACC_FULL(pb), (prev_last_job >> a.nbuf_log2) & 1u);
            }
            if (grp == 0) {
                const SaLayer &L0 = a.L[0];
                const uint32_t sbo = (uint32_t)L0.xw * 16u;
                uint8_t *xrow = xa + (size_t)(r >> 3) * sbo + row_off;
                const long long grow = (long long)tile * MM_ROWS + r;
                const bool ok = grow < a.rows;
                const int j = jn;
                {
                    const long long gnext = grow + (long long)gridDim.x * MM_ROWS;
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
                if (!a.split) {
                    const int nch = L0.kpad >> 3;
                    const int fch = a.cpad8 >> 3;
                    // the row's feature channels go global -> shared with 16-byte cp.async (LDGSTS): every load of the row is in
                    // flight at once and no register is staged (the register version moved them in batches of 8, one L2 round
                    // trip per batch: 4 round trips for the 256-channel rows of layer 5); rows past the end are zero-filled
                    // (rows of <= 8 loads -- one batch either way -- keep the register path: cp.async + wait + proxy fence measured
                    // 5-8 % slower on the 64-channel chains of layer 1)
                    const uint4 *trow = fch ? reinterpret_cast<const uint4 *>(a.twin + (ok ? ((size_t)bb * a.n + j) * a.ldtwin : 0)) : nullptr;
                    int c = 0;
                    if (fch > 8) {
                        const uint32_t xdst = smem_u32(xrow);
                        for (; c < fch; ++c) cp_async16(xdst + (uint32_t)c * 128u, trow + c, ok ? 16u : 0u);
                        cp_async_commit();
                    } else {
                        uint4 t[8];
#pragma unroll
                        for (int u = 0; u < 8; ++u) t[u] = (ok && u < fch) ? __ldg(trow + u) : zero;
#pragma unroll
                        for (int u = 0; u < 8; ++u)
                            if (u < fch) *reinterpret_cast<uint4 *>(xrow + (size_t)u * 128) = t[u];
                        c = fch;
                    }
                    if (a.use_xyz) {
                        *reinterpret_cast<uint4 *>(xrow + (size_t)c * 128) = make_uint4(pack_h2(dx, dy), pack_h2(dz, 0.f), 0u, 0u);
                        ++c;
                    }
                    for (; c < nch; ++c) *reinterpret_cast<uint4 *>(xrow + (size_t)c * 128) = zero;
                } else {
                    // exact inputs: features from the fp32 channel-major tensor (c_feat <= 8), k order [f0..f7 | x y z 0..]
                    float y[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) y[i] = 0.f;
                    if (ok) {
#pragma unroll
                        for (int c = 0; c < 8; ++c)
                            if (c < a.c_feat) y[c] = __ldg(a.feat32 + ((size_t)bb * a.c_feat + c) * a.n + j);
                    }
                    if (a.c_feat) { y[8] = dx; y[9] = dy; y[10] = dz; }
                    else { y[0] = dx; y[1] = dy; y[2] = dz; }
                    if (a.l0_fused) {
                        // layer 0 right here, in fp32 on the CUDA cores (<= 11 real inputs x <= 32 outputs per row): saves one
                        // MMA job + TMEM round trip + hand-off per tile; the result goes straight into layer 1's [hi | lo] operand
                        const float *w0 = reinterpret_cast<const float *>(wst + a.l0_off);
                        const int c0n = L0.cpad;   // 16 or 32
                        float h[32];
#pragma unroll
                        for (int c = 0; c < 32; c += 4) {
                            const float4 bv = c < c0n ? *reinterpret_cast<const float4 *>(w0 + 16 * c0n + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                            h[c] = bv.x; h[c + 1] = bv.y; h[c + 2] = bv.z; h[c + 3] = bv.w;
                        }
                        const int xo = a.c_feat ? 8 : 0;
#pragma unroll
                        for (int k = 0; k < 16; ++k) {
                            if (k < a.c_feat || (k >= xo && k < xo + 3)) {   // warp-uniform: rows that are not padding
                                const float yk = y[k];
#pragma unroll
                                for (int c = 0; c < 32; c += 4) {
                                    if (c < c0n) {
                                        const float4 wv = *reinterpret_cast<const float4 *>(w0 + k * c0n + c);
                                        h[c] = fmaf(yk, wv.x, h[c]); h[c + 1] = fmaf(yk, wv.y, h[c + 1]);
                                        h[c + 2] = fmaf(yk, wv.z, h[c + 2]); h[c + 3] = fmaf(yk, wv.w, h[c + 3]);
                                    }
                                }
                            }
                        }
                        const SaLayer &L1 = a.L[1];
                        uint8_t *x1 = xb + (size_t)(r >> 3) * ((uint32_t)L1.xw * 16u) + row_off;
                        uint8_t *x1lo = x1 + (size_t)(L1.kpad >> 3) * 128;
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            if (8 * g < c0n) {
                                float t8[8];
#pragma unroll
                                for (int i = 0; i < 8; ++i) t8[i] = fmaxf(h[8 * g + i], 0.f);
                                uint4 hi, lo;
                                split8(t8, hi, lo);
                                *reinterpret_cast<uint4 *>(x1 + (size_t)g * 128) = hi;
                                *reinterpret_cast<uint4 *>(x1lo + (size_t)g * 128) = lo;
                            }
                        }
                        fence_proxy_async();
                        for (int c = 0; c < L1.n_xc; ++c) mbar_arrive(XR(1, c));
                    } else {
                    const int nch = L0.kpad >> 3;   // 1 or 2 groups of 8
                    uint8_t *xlo = xrow + (size_t)nch * 128;
                    uint4 hi, lo;
                    split8(y, hi, lo);
                    *reinterpret_cast<uint4 *>(xrow) = hi;
                    *reinterpret_cast<uint4 *>(xlo) = lo;
                    if (nch > 1) {
                        split8(y + 8, hi, lo);
                        *reinterpret_cast<uint4 *>(xrow + 128) = hi;
                        *reinterpret_cast<uint4 *>(xlo + 128) = lo;
                    }
                    }
                }
                if (!a.l0_fused) {
                    if (!a.split && (a.cpad8 >> 3) > 8) cp_async_wait<0>();
                    fence_proxy_async();
                    for (int c = 0; c < L0.n_xc; ++c) mbar_arrive(XR(0, c));
                }
            }
            pf.add(PF_EPI_GATHER, t_g);
            // ---- hidden layers: D[row, cout] -> relu(D + bias) -> fp16 -> next X, chunk by chunk
            for (int l = a.l0_fused; l < nL - 1; ++l) {
                const SaLayer &Ly = a.L[l];
                const SaLayer &Ln = a.L[l + 1];
                const int obuf = (l + 1) & 1;
                uint8_t *xo = obuf ? xb : xa;
                const uint32_t sbo = (uint32_t)Ln.xw * 16u;
                uint8_t *xrow = xo + (size_t)(r >> 3) * sbo + row_off;
                const float *bias = a.bias + Ly.bias_off;
                for (int cc = 0; cc < Ly.n_cc; ++cc, ++job) {
                    if (G == 2 && (job & 1u) != grp) continue;
                    const int ncols = min(128, Ly.cpad - cc * 128);
                    const int buf = (int)(job & ((uint32_t)a.nbuf - 1u));
                    { const long long t0 = pf.now(); mbar_wait(ACC_FULL(buf), (job >> a.nbuf_log2) & 1u); pf.add(PF_EPI_WAIT_HID, t0); }
                    tc_fence_after();
                    const long long t_w = pf.now();
                    const uint32_t taddr = tmem_base + lane_field + (uint32_t)(buf * 128);
                    float hmx = 0.f;   // fp16 range guard: largest hidden activation this thread stores in this job
                    if (!a.split) {
                        for (int h0 = 0; h0 < ncols; h0 += 64) {   // one 64-wide K chunk of the next operand at a time
                            const int hend = min(ncols, h0 + 64);
                            int c0 = (PROF && (a.abl & 2)) ? hend : h0;   // ablation: no TMEM loads / stores, only the hand-off
                            for (; c0 + 32 <= hend; c0 += 32) {
                                float v[32];
                                tmem_ld32(taddr + (uint32_t)c0, v);
                                const int col = cc * 128 + c0;
                                store_hidden16(v, bias + col, xrow + (size_t)(col >> 3) * 128, hmx);
                                store_hidden16(v + 16, bias + col + 16, xrow + (size_t)((col >> 3) + 2) * 128, hmx);
                            }
                            if (c0 < hend) {
                                float v[16];
                                tmem_ld16(taddr + (uint32_t)c0, v);
                                const int col = cc * 128 + c0;
                                store_hidden16(v, bias + col, xrow + (size_t)(col >> 3) * 128, hmx);
                            }
                            fence_proxy_async();
                            mbar_arrive(XR(obuf, (cc * 128 + h0) >> 6));
                        }
                    } else {
                        uint8_t *xlo = xrow + (size_t)(Ln.kpad >> 3) * 128;
                        int c0 = 0;
                        for (; c0 + 32 <= ncols; c0 += 32) {
                            float v[32];
                            tmem_ld32(taddr + (uint32_t)c0, v);
                            store_hidden16_split(v, bias + c0, xrow + (size_t)(c0 >> 3) * 128, xlo + (size_t)(c0 >> 3) * 128, hmx);
                            store_hidden16_split(v + 16, bias + c0 + 16, xrow + (size_t)((c0 >> 3) + 2) * 128, xlo + (size_t)((c0 >> 3) + 2) * 128, hmx);
                        }
                        if (c0 < ncols) {
                            float v[16];
                            tmem_ld16(taddr + (uint32_t)c0, v);
                            store_hidden16_split(v, bias + c0, xrow + (size_t)(c0 >> 3) * 128, xlo + (size_t)(c0 >> 3) * 128, hmx);
                        }
                        fence_proxy_async();
                        for (int c = 0; c < Ln.n_xc; ++c) mbar_arrive(XR(obuf, c));
                    }
                    if (hmx > FP16_MAX && a.ovf) atomicOr(a.ovf, a.ovf_bit);
                    tc_fence_before();
                    mbar_arrive(ACC_EMPTY(buf));
                    pf.add(PF_EPI_WORK_HID, t_w);
                }
            }
            // ---- batch-statistics pass (training-mode BN): last layer D[cout, row], thread = cout; no bias / ReLU / pool, the raw
            //      accumulators of the tile's real rows are summed per cout into this (CTA, group)'s slice of a.stats
            if (a.stats) {
                const SaLayer &Ly = a.L[nL - 1];
                const long long left = a.rows - (long long)tile * MM_ROWS;
                const int nv = left < MM_ROWS ? (int)left : MM_ROWS;
                double *slice = a.stats + (size_t)(blockIdx.x * G + (G == 2 ? grp : 0u)) * (size_t)Ly.cpad * 2;
                for (int cci = 0; cci < Ly.n_cc; ++cci, ++job) {
                    if (G == 2 && (job & 1u) != grp) continue;
                    const int cc = chunk_of(nL - 1, cci, Ly.n_cc);
                    const int buf = (int)(job & ((uint32_t)a.nbuf - 1u));
                    mbar_wait(ACC_FULL(buf), (job >> a.nbuf_log2) & 1u);
                    tc_fence_after();
                    stats_chunk(tmem_base + lane_field + (uint32_t)(buf * 128), nv, slice + (size_t)(cc * 128 + r) * 2);
                    tc_fence_before();
                    mbar_arrive(ACC_EMPTY(buf));
                }
            } else
            // ---- last layer: D[cout, row]; thread = cout; max over each centre's nsample columns
            {
                const SaLayer &Ly = a.L[nL - 1];
                const float *bias = a.bias + Ly.bias_off;
                const long long q0 = ((long long)tile * MM_ROWS) >> a.ns_log2;  // first centre of the tile
                const long long qmax = (long long)a.b * a.m;
                const int ns = a.nsample;
                const long long bb0 = q0 / a.m;
                const int p0 = (int)(q0 - bb0 * a.m);
                const size_t sstride = (size_t)a.c_total * a.m;
                for (int cci = 0; cci < Ly.n_cc; ++cci, ++job) {
                    if (G == 2 && (job & 1u) != grp) continue;
                    const int cc = chunk_of(nL - 1, cci, Ly.n_cc);
                    const int buf = (int)(job & ((uint32_t)a.nbuf - 1u));
                    { const long long t0 = pf.now(); mbar_wait(ACC_FULL(buf), (job >> a.nbuf_log2) & 1u); pf.add(PF_EPI_WAIT_POOL, t0); }
                    tc_fence_after();
                    const long long t_w = pf.now();
                    const int ch = cc * 128 + r;
                    const bool w32 = a.out != nullptr && ch < a.cout_last;
                    const bool w16 = a.out16 != nullptr && ch < a.n16;
                    const float bv = __ldg(bias + ch);   // bias is padded to cpad
                    PoolOut o;
                    o.bv = bv; o.w32 = w32; o.w16 = w16;
                    o.outc = a.out + ((size_t)bb0 * a.c_total + a.co_off + (w32 ? ch : 0)) * a.m + (size_t)p0;
                    o.out16 = a.out16 + (size_t)q0 * a.ld16 + a.co16 + (w16 ? ch : 0);
                    o.ld16 = a.ld16; o.o16lo = a.o16lo; o.q = q0; o.qmax = qmax; o.m = a.m; o.p = p0; o.scene_stride = sstride;
                    const uint32_t taddr = tmem_base + lane_field + (uint32_t)(buf * 128);
                    if (ns == 32) pool_chunk<32>(taddr, o);
                    else if (ns == 16) pool_chunk<16>(taddr, o);
                    else pool_chunk_any(taddr, o, ns);
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
    if (warp == W_MMA) {
        tc_fence_after();
        tmem_dealloc(tmem_base, (uint32_t)a.tmem_cols);
    }
}

// ---- feature twin: (b, c, n) f32 channel-major -> (b, n, ld) fp16 point-major (zero padded) -------------
__global__ void __launch_bounds__(256)
make_twin_kernel(int c, int n, int cpad8, const float *__restrict__ in, __half *__restrict__ out, unsigned int *ovf) {
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
        if (ni < n && ci < cpad8) {
            out[((size_t)b * n + ni) * cpad8 + ci] = __float2half_rn(tile[tx][i]);
            if (fabsf(tile[tx][i]) > FP16_MAX && ovf) atomicOr(ovf, 1u);   // untagged (bit 0): the caller's input does not fit fp16
        }
    }
}

static unsigned long long *g_sa_prof = nullptr;

// ---- tuning / A-B knobs: environment variables, read ONCE per process (not on every launch) -------------------------------
struct SaTuning {
    int max_stages, max_ctas, grid_mult;
    int abl;
    bool no_lring, no_sched, rot, no_narrow, one_group, no_double;
    SaTuning() {
        auto geti = [](const char *name, int dflt) { const char *e = getenv(name); return e ? atoi(e) : dflt; };
        max_stages = max(2, min(MM_MAX_STAGES, geti("SPSK_SA_MAX_STAGES", MM_MAX_STAGES)));   // sensitivity measurements
        max_ctas = max(1, min(4, geti("SPSK_SA_MAX_CTAS", 4)));
        grid_mult = max(1, min(8, geti("SPSK_SA_GRID_MULT", 1)));   // persistent CTAs, exactly one resident set (2 measured 1 % slower with 8 batches in flight)
        abl = geti("SPSK_SA_ABL", 0);                                 // ablations of the profiling kernels (see SaArgs::abl)
        no_lring = getenv("SPSK_SA_NO_LRING") != nullptr;
        no_sched = getenv("SPSK_SA_NO_SCHED") != nullptr;
        rot = getenv("SPSK_SA_ROT") != nullptr;              // opt-in: measured neutral on B200 (the weight stream is not L2 hot-line bound)
        no_narrow = getenv("SPSK_SA_NO_NARROW") != nullptr;
        no_double = getenv("SPSK_SA_NO_DOUBLE") != nullptr;   // A/B: one 64-wide k tile per schedule entry everywhere
        one_group = getenv("SPSK_SA_ONE_GROUP") != nullptr;
    }
};
static const SaTuning &sa_tuning() {
    static const SaTuning t;
    return t;
}

// ---- host-side planning ---------------------------------------------------------------------------------
struct SaPlan {
    SaLayer L[MM_MAX_LAYERS];
    int xa_bytes, xb_bytes, w_total, l0_off, resident, nstages, lstages, ctas, tmem_cols, nbuf, smem, sched_n;
};

constexpr int PR_HDR_BYTES = 2048;   // header of the pair kernel (sa_mma_pair.cu)

static int sa_plan(const spsk_sa_mma_desc *d, SaPlan *P) {
    const int nL = d->nlayers;
    const bool pair = d->pair != 0;
    SPSK_REQUIRE(nL >= 1 && nL <= MM_MAX_LAYERS, SPSK_ERR_UNSUPPORTED, "sa_mma: nlayers=%d (1..%d)", nL, MM_MAX_LAYERS);
    const int split = d->split ? 1 : 0;
    int w_off = 0, b_off = 0;
    P->xa_bytes = P->xb_bytes = 0;
    for (int l = 0; l < nL; ++l) {
        const bool last = l == nL - 1;
        const int kpad = d->kpad[l], cpad = d->cpad[l];
        SPSK_REQUIRE(kpad >= 16 && kpad % 16 == 0 && kpad <= 1024, SPSK_ERR_UNSUPPORTED, "sa_mma: layer %d kpad=%d", l, kpad);
        SPSK_REQUIRE(cpad >= 16 && cpad % (last ? 128 : 16) == 0 && cpad <= (last ? 4096 : 1024), SPSK_ERR_UNSUPPORTED,
                     "sa_mma: layer %d cpad=%d", l, cpad);
        SPSK_REQUIRE(last || cpad == d->kpad[l + 1], SPSK_ERR_INVALID_ARG, "sa_mma: layer %d cpad != next kpad", l);
        SPSK_REQUIRE(!split || (kpad <= 64 && (last || cpad <= 64)), SPSK_ERR_UNSUPPORTED, "sa_mma: split arithmetic needs every width <= 64");
        SPSK_REQUIRE(!pair || (!split && (!last || cpad % 256 == 0)), SPSK_ERR_UNSUPPORTED, "sa_mma: pair kernel needs plain arithmetic and a last cpad multiple of 256");
        SaLayer &Ly = P->L[l];
        Ly.kpad = kpad; Ly.cpad = cpad;
        Ly.n_cc = pair ? (cpad + 255) / 256 : (cpad + 127) / 128;
        Ly.xw = split ? 2 * kpad : kpad;
        Ly.vk = split ? 3 * kpad : kpad;
        Ly.wk = split ? 2 * kpad : kpad;
        Ly.n_kc = (Ly.wk + 63) / 64;
        Ly.n_xc = (Ly.xw + 63) / 64;
        Ly.w_off = w_off; Ly.bias_off = b_off;
        w_off += Ly.wk * cpad * 2;
        b_off += cpad;
        const int xbytes = Ly.xw * MM_ROWS * 2;
        if (l & 1) P->xb_bytes = max(P->xb_bytes, xbytes); else P->xa_bytes = max(P->xa_bytes, xbytes);
    }
    P->w_total = w_off;
    P->l0_off = 0;
    if (d->l0_fused) {
        SPSK_REQUIRE(split && nL >= 2 && d->cpad[0] <= 32 && !pair, SPSK_ERR_UNSUPPORTED, "sa_mma: layer-0 fusion needs a split chain of >= 2 layers with cpad[0] <= 32");
        P->l0_off = w_off;
        P->w_total = w_off + (16 * d->cpad[0] + d->cpad[0]) * 4;
        P->w_total = (P->w_total + 15) & ~15;
    }
    const int xtot = (pair ? PR_HDR_BYTES : MM_HDR) + P->xa_bytes + P->xb_bytes;
    const int sm_bytes = 228 * 1024;   // per SM; every resident CTA also reserves 1 KB
    // CTAs per SM: as many (<= 3) as shared memory allows with the chain resident or >= 2 ring stages; the TMEM
    // share (512 / ctas rounded down to a power of two) must hold at least one 128-column accumulator
    auto fits = [&](int ctas, int *resident, int *nstages, int *smem) {
        const int per = (sm_bytes / ctas - 1024) & ~127;
        if (xtot + P->w_total <= per) { *resident = 1; *nstages = 1; *smem = xtot + P->w_total; return true; }
        const int st = (per - xtot) / MM_STAGE_BYTES;
        if (st < 2) return false;
        const int cap = sa_tuning().max_stages;
        *resident = 0; *nstages = st > cap ? cap : st; *smem = xtot + *nstages * MM_STAGE_BYTES;
        return true;
    };
    int ctas = 0;
    const int cmax = pair ? 1 : sa_tuning().max_ctas;
    bool one_job_per_layer = true;   // the register-resident "narrow" issue loop (the only one the 4-CTA shape runs)
    for (int l = 0; l < nL; ++l) one_job_per_layer = one_job_per_layer && P->L[l].n_cc == 1;
    for (int c = cmax; c >= 1; --c) {
        int res, st, sm;
        if (!fits(c, &res, &st, &sm)) continue;
        if (c == 4 && !(res && one_job_per_layer && !d->l0_fused)) continue;   // four CTAs per SM: resident narrow chains only (no producer warp)
        ctas = c; P->resident = res; P->nstages = st; P->smem = sm;
        break;
    }
    SPSK_REQUIRE(ctas >= 1, SPSK_ERR_UNSUPPORTED, "sa_mma: activation tiles do not fit shared memory (layer widths too large)");
    P->ctas = ctas;
    // a starved ring (big activations leave < 4 stages): give the last layer -- most of the chain's weight bytes -- its own
    // ring overlaid on the activation buffer that is dead while it runs
    P->lstages = 0;
    if (pair && P->resident) {   // the pair kernel always streams: turn the resident plan into a ring
        P->resident = 0;
        P->nstages = min(MM_MAX_STAGES, max(2, (P->smem - xtot) / MM_STAGE_BYTES));
        P->smem = xtot + P->nstages * MM_STAGE_BYTES;
    }
    if (!P->resident && P->nstages < 4 && nL >= 2 && !sa_tuning().no_lring) {
        const int other = ((nL - 2) & 1) ? P->xb_bytes : P->xa_bytes;
        const int e = other / MM_STAGE_BYTES;
        if (e >= 3) P->lstages = e > MM_MAX_STAGES ? MM_MAX_STAGES : e;
    }
    P->tmem_cols = ctas == 1 ? 512 : (ctas == 2 ? 256 : 128);   // 3 or 4 CTAs: 128 columns each
    P->nbuf = P->tmem_cols / 128;
    // streaming chains: tabulate the per-tile MMA schedule (one 16-byte entry per weight tile) into the kernel parameters
    P->sched_n = 0;
    if (!pair && !P->resident && !split && !sa_tuning().no_sched) {
        int ntab = 0;
        for (int l = 0; l < nL; ++l) ntab += P->L[l].n_cc * P->L[l].n_kc;
        if (ntab <= MM_SCHED_MAX) P->sched_n = ntab;
    }
    // shared-memory floor so that no more than `ctas` CTAs land on an SM (their TMEM allocations would not fit)
    const int floor_bytes = (sm_bytes / (ctas + 1) - 1024 + 256) & ~127;
    if (P->smem < floor_bytes) P->smem = floor_bytes;
    return SPSK_OK;
}

// The per-tile schedule of a streaming chain (see the tabulated issue loop and the table-driven producer of sa_mma_kernel).
// Shared-memory offsets are relative to the CTA's dynamic shared memory: [header MM_HDR: barriers][XA][XB][weight ring]; the
// barrier offsets mirror the W_FULL / W_EMPTY / XR / WL_FULL / WL_EMPTY lambdas of the kernel.  Returns the number of entries
// (weight-tile groups + ring padding), or 0 when it does not fit MM_SCHED_MAX.
static int build_schedule(SaArgs &a, bool allow_double) {
    const int nL = a.nlayers;
    const bool lring = a.lstages > 0;
    auto w_full = [](int s) { return 8u * (uint32_t)s; };
    auto w_empty = [](int s) { return 8u * (uint32_t)(MM_MAX_STAGES + s); };
    auto xr = [](int buf, int c) { return 8u * (uint32_t)(2 * MM_MAX_STAGES + 8 + buf * MM_MAX_XC + c); };
    auto wl_full = [](int s) { return 8u * (uint32_t)(2 * MM_MAX_STAGES + 8 + 2 * MM_MAX_XC + s); };
    auto wl_empty = [](int s) { return 8u * (uint32_t)(3 * MM_MAX_STAGES + 8 + 2 * MM_MAX_XC + s); };
    // a ring whose 16 KB stages can be regrouped into >= 2 slots of 32 KB takes two consecutive k tiles per entry
    const int stages16[2] = {a.nstages, lring ? a.lstages : 0};
    bool dbl[2];
    int depth[2];
    for (int r = 0; r < 2; ++r) {
        dbl[r] = allow_double && stages16[r] >= 4;
        depth[r] = stages16[r] ? (dbl[r] ? stages16[r] / 2 : stages16[r]) : 1;
    }
    auto ring_of = [&](int l) { return (lring && l == nL - 1) ? 1 : 0; };
    // how many k tiles entry (layer l, first tile kc) covers: 2 when the ring is doubled and both tiles are full 64-wide ones
    auto span = [&](int l, int kc) {
        const SaLayer &Ly = a.L[l];
        return (dbl[ring_of(l)] && kc + 1 < Ly.n_kc && Ly.wk - (kc + 1) * 64 >= 64) ? 2 : 1;
    };
    int n_ring[2] = {0, 0};
    for (int l = 0; l < nL; ++l) {
        int per_chunk = 0;
        for (int kc = 0; kc < a.L[l].n_kc; kc += span(l, kc)) ++per_chunk;
        n_ring[ring_of(l)] += a.L[l].n_cc * per_chunk;
    }
    int pad[2], uses[2];
    for (int r = 0; r < 2; ++r) {
        pad[r] = (depth[r] - n_ring[r] % depth[r]) % depth[r];
        uses[r] = (n_ring[r] + pad[r]) / depth[r];
    }
    if (n_ring[0] + n_ring[1] + pad[0] + pad[1] > MM_SCHED_MAX) return 0;
    // completions per tile of every activation-chunk barrier (a buffer serves layers l, l+2, ...)
    int n_compl[2][MM_MAX_XC] = {};
    for (int l = 0; l < nL; ++l)
        for (int c = 0; c < a.L[l].n_xc; ++c) ++n_compl[l & 1][c];
    const uint32_t ring_base[2] = {(uint32_t)(MM_HDR + a.xa_bytes + a.xb_bytes),
                                   (uint32_t)MM_HDR + ((((nL - 2) & 1) && nL >= 2) ? (uint32_t)a.xa_bytes : 0u)};   // overlay: input buffer of layer nL-2
    int e = 0, pos[2] = {0, 0};
    int seen[2][MM_MAX_XC] = {};   // completions of each activation-chunk barrier before the layer being emitted
    auto ring_words = [&](int r, uint32_t src, uint32_t bytes, uint4 &R) {
        const int slot = pos[r] % depth[r], k = pos[r] / depth[r];
        ++pos[r];
        const uint32_t slot_off = ring_base[r] + (uint32_t)slot * (dbl[r] ? 2u : 1u) * MM_STAGE_BYTES;
        const uint32_t full = r ? wl_full(slot) : w_full(slot), empty = r ? wl_empty(slot) : w_empty(slot);
        R.x = src;
        R.y = bytes | ((slot_off >> 4) << 16);
        R.z = full | (empty << 10) | ((uint32_t)(k & 1) << 20) | ((uint32_t)(uses[r] & 1) << 21);
        R.w = 0u;
    };
    auto pad_ring = [&](int r) {
        for (int i = 0; i < pad[r]; ++i, ++e) {
            uint4 R;
            ring_words(r, 0u, 0u, R);
            a.sched[e] = make_uint4(0u, 0u, 0u, 0u);   // no MMAs, no job flags: the issuer only releases the slot
            a.ring[e] = R;
        }
    };
    for (int l = 0; l < nL; ++l) {
        const SaLayer &Ly = a.L[l];
        const bool last = (l == nL - 1);
        const int r = ring_of(l);
        if (r == 1) pad_ring(0);   // ring 0 is complete for this tile before the overlay ring starts
        const uint32_t x_off = (uint32_t)MM_HDR + ((l & 1) ? (uint32_t)a.xa_bytes : 0u);
        const uint32_t x_lo0 = (x_off >> 4) | ((128u >> 4) << 16);                  // umma_desc_lo(base + x_off, LBO 128) - (base >> 4)
        const uint32_t x_hi = (((uint32_t)Ly.xw * 16u) >> 4) | (1u << 14);           // umma_desc_hi(SBO)
        for (int cci = 0; cci < Ly.n_cc; ++cci) {
            const int ncols = (Ly.cpad - cci * 128) < 128 ? (Ly.cpad - cci * 128) : 128;
            for (int kc = 0; kc < Ly.n_kc; ++e) {
                const int nt = span(l, kc);
                const int kw = (Ly.wk - kc * 64) < 64 ? (Ly.wk - kc * 64) : 64;   // width of the entry's first tile (the second, if any, is 64)
                const int nk16 = (kw >> 4) + (nt == 2 ? 4 : 0);
                uint32_t f = (uint32_t)nk16 | ((uint32_t)(ncols * 8) << 16);      // second tile: ncols x 64 halfs further = ncols * 8 units
                if (kc == 0) f |= SCH_FIRST_KC;
                if (kc + nt >= Ly.n_kc) f |= SCH_LAST_KC;
                if (r == 1 && cci == 0 && kc == 0) f |= SCH_LRING_FIRST;
                if (lring && l == nL - 2 && cci == Ly.n_cc - 1 && kc + nt >= Ly.n_kc) f |= SCH_HID_DONE;
                if (last) f |= SCH_LAST_LAYER;
                const int n_idesc = last ? MM_ROWS : ncols;
                const uint32_t idesc = (1u << 4) | ((uint32_t)(n_idesc >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);   // umma_idesc(128, n)
                const uint32_t w_hi = (((uint32_t)kw * 16u) >> 4) | (1u << 14);
                a.sched[e] = make_uint4(x_lo0 + (uint32_t)kc * 64u, idesc, x_hi | (w_hi << 16), f);
                uint4 R;
                ring_words(r, (uint32_t)(Ly.w_off + (128 * cci * Ly.wk + ncols * 64 * kc) * 2), (uint32_t)(ncols * (kw + (nt == 2 ? 64 : 0)) * 2), R);
                if (cci == 0) {
                    // the 64-wide activation chunks these weight tiles read become ready during the first cout chunk of the layer
                    // (chunk kc: written by the gather / the previous layer's epilogue); later cout chunks re-read what is already there
                    for (int t = 0; t < nt; ++t) {
                        const int c = kc + t;
                        R.z |= ((1u | ((uint32_t)(seen[l & 1][c] & 1) << 1) | ((uint32_t)(n_compl[l & 1][c] & 1) << 2)) << (22 + 3 * t));
                        R.w |= xr(l & 1, c) << (10 * t);
                    }
                }
                a.ring[e] = R;
                kc += nt;
            }
        }
        for (int c = 0; c < Ly.n_xc; ++c) ++seen[l & 1][c];
    }
    pad_ring(lring ? 1 : 0);
    return e;
}

}  // namespace spsk

int spsk_sa_mma_pair_launch(const spsk::SaArgs &a, int smem_bytes, cudaStream_t st);   // sa_mma_pair.cu

extern "C" int spsk_make_twin(int b, int c, int n, int cpad8, const float *features, void *twin, spsk_stream_t stream) {
    using namespace spsk;
    SPSK_REQUIRE(b >= 0 && c >= 1 && n >= 0 && cpad8 >= c && (cpad8 % 8) == 0 && b <= 65535, SPSK_ERR_INVALID_ARG,
                 "make_twin: bad sizes b=%d c=%d n=%d cpad8=%d", b, c, n, cpad8);
    if (b == 0 || n == 0) return SPSK_OK;
    SPSK_REQUIRE(features && twin, SPSK_ERR_INVALID_ARG, "make_twin: null pointer");
    dim3 grid((n + 31) / 32, (cpad8 + 31) / 32, b);
    make_twin_kernel<<<grid, 256, 0, as_stream(stream)>>>(c, n, cpad8, features, reinterpret_cast<__half *>(twin), fp16_overflow_word());
    SPSK_LAUNCH_CHECK("make_twin_kernel");
    return SPSK_OK;
}

extern "C" int spsk_sa_mma_set_profile(unsigned long long *counters) {
    spsk::g_sa_prof = counters;
    return SPSK_OK;
}

extern "C" int spsk_sa_mma_config(const spsk_sa_mma_desc *d, int *smem_bytes, int *ctas_per_sm, int *nstages, int *resident) {
    using namespace spsk;
    SPSK_REQUIRE(d, SPSK_ERR_INVALID_ARG, "sa_mma: null descriptor");
    SaPlan P;
    if (int rc = sa_plan(d, &P)) return rc;
    if (smem_bytes) *smem_bytes = P.smem;
    if (ctas_per_sm) *ctas_per_sm = P.ctas;
    if (nstages) *nstages = P.nstages;
    if (resident) *resident = P.resident;
    return SPSK_OK;
}

// Host-only: the static schedule a streaming chain would run with (test / documentation aid; no device work).
//   words: capacity >= 8 * SPSK_SA_SCHED_MAX uint32; entry e = words[8e .. 8e+3] (sched: activation descriptor lo, instruction
//   descriptor, descriptor hi words, flags) + words[8e+4 .. 8e+7] (ring: source offset, bytes | slot offset/16 << 16, barriers and
//   parities, activation-chunk barriers) -- the layout documented at the tabulated issue loop of sa_mma_kernel.
//   info[0] = hidden-ring 16 KB stages, [1] = overlay-ring 16 KB stages (0 = none), [2] = resident, [3] = shared-memory offset of
//   the hidden ring, [4] = of the overlay ring, [5] = total packed weight bytes.  *n = 0 when the chain is not tabulated.
extern "C" int spsk_sa_mma_schedule(const spsk_sa_mma_desc *d, int *n, unsigned int *words, int *info) {
    using namespace spsk;
    SPSK_REQUIRE(d && n && words && info, SPSK_ERR_INVALID_ARG, "sa_mma_schedule: null pointer");
    SaPlan P;
    if (int rc = sa_plan(d, &P)) return rc;
    static SaArgs a;   // ~4 KB: keep it off the stack; host-only helper, not thread-safe by design (tests)
    a = SaArgs{};
    a.nlayers = d->nlayers;
    for (int l = 0; l < d->nlayers; ++l) a.L[l] = P.L[l];
    a.nstages = P.nstages; a.lstages = P.lstages; a.resident = P.resident;
    a.xa_bytes = P.xa_bytes; a.xb_bytes = P.xb_bytes;
    a.sched_n = (P.sched_n > 0 && !d->pair) ? build_schedule(a, !sa_tuning().no_double) : 0;
    *n = a.sched_n;
    for (int e = 0; e < a.sched_n; ++e) {
        const uint4 E = a.sched[e], R = a.ring[e];
        unsigned int *w = words + 8 * e;
        w[0] = E.x; w[1] = E.y; w[2] = E.z; w[3] = E.w; w[4] = R.x; w[5] = R.y; w[6] = R.z; w[7] = R.w;
    }
    info[0] = P.nstages; info[1] = P.lstages; info[2] = P.resident;
    info[3] = MM_HDR + P.xa_bytes + P.xb_bytes;
    info[4] = MM_HDR + ((((d->nlayers - 2) & 1) && d->nlayers >= 2) ? P.xa_bytes : 0);
    info[5] = P.w_total;
    return SPSK_OK;
}

extern "C" int spsk_sa_mma_stats_parts(const spsk_sa_mma_desc *d, int *nparts) {
    using namespace spsk;
    SPSK_REQUIRE(d && nparts, SPSK_ERR_INVALID_ARG, "sa_mma: null descriptor");
    SPSK_REQUIRE(!d->pair, SPSK_ERR_UNSUPPORTED, "sa_mma: the batch-statistics pass runs on the single-CTA kernel (pair == 0)");
    SaPlan P;
    if (int rc = sa_plan(d, &P)) return rc;
    const long long rows = (long long)d->b * d->m * d->nsample;
    const long long ntiles = (rows + MM_ROWS - 1) / MM_ROWS;
    const long long slots = (long long)SPSK_NUM_SMS * P.ctas * sa_tuning().grid_mult;
    const long long grid = ntiles < slots ? ntiles : slots;
    const bool two_groups = P.ctas == 1 && !sa_tuning().one_group;
    *nparts = (int)(grid * (two_groups ? 2 : 1));
    return SPSK_OK;
}

extern "C" int spsk_sa_mma_forward(const spsk_sa_mma_desc *d, spsk_stream_t stream) {
    using namespace spsk;
    SPSK_REQUIRE(d, SPSK_ERR_INVALID_ARG, "sa_mma: null descriptor");
    SPSK_REQUIRE(d->wtiles && d->bias && (d->out_cm || d->out16 || d->stats), SPSK_ERR_INVALID_ARG, "sa_mma: null weights / no output");
    SPSK_REQUIRE(!d->stats || !d->pair, SPSK_ERR_UNSUPPORTED, "sa_mma: the batch-statistics pass runs on the single-CTA kernel (pair == 0)");
    SPSK_REQUIRE(d->nsample >= 1 && d->nsample <= MM_ROWS && (d->nsample & (d->nsample - 1)) == 0, SPSK_ERR_UNSUPPORTED,
                 "sa_mma: nsample=%d must be a power of two <= %d", d->nsample, MM_ROWS);
    SPSK_REQUIRE(d->b >= 0 && d->n >= 1 && d->m >= 0 && d->c_feat >= 0, SPSK_ERR_INVALID_ARG, "sa_mma: bad sizes");
    SPSK_REQUIRE(d->idx && d->xyz && d->new_xyz, SPSK_ERR_INVALID_ARG, "sa_mma: null gather source");
    SaPlan P;
    if (int rc = sa_plan(d, &P)) return rc;
    SaArgs a{};
    a.nlayers = d->nlayers;
    for (int l = 0; l < d->nlayers; ++l) a.L[l] = P.L[l];
    a.b = d->b; a.n = d->n; a.m = d->m; a.nsample = d->nsample;
    a.ns_log2 = 0;
    while ((1 << a.ns_log2) < d->nsample) ++a.ns_log2;
    a.c_feat = d->c_feat;
    a.use_xyz = d->use_xyz ? 1 : 0;
    a.split = d->split ? 1 : 0;
    if (a.split) {
        SPSK_REQUIRE(d->c_feat <= 8 && (d->c_feat == 0 || d->features), SPSK_ERR_UNSUPPORTED, "sa_mma: split arithmetic takes <= 8 fp32 feature channels");
        SPSK_REQUIRE(d->kpad[0] == 16, SPSK_ERR_INVALID_ARG, "sa_mma: split kpad[0]=%d must be 16 ([f0..f7 | x y z 0..])", d->kpad[0]);
        a.feat32 = d->features;
        a.cpad8 = d->c_feat ? 8 : 0;
    } else {
        a.cpad8 = (d->c_feat + 7) / 8 * 8;
        SPSK_REQUIRE(d->c_feat == 0 || (d->twin && d->ldtwin >= a.cpad8 && d->ldtwin % 8 == 0 && (reinterpret_cast<uintptr_t>(d->twin) & 15) == 0),
                     SPSK_ERR_INVALID_ARG, "sa_mma: twin must be a 16-byte aligned (b, n, ldtwin) fp16 tensor with ldtwin %% 8 == 0 and >= c_feat");
        SPSK_REQUIRE(d->kpad[0] >= a.cpad8 + (a.use_xyz ? 8 : 0), SPSK_ERR_INVALID_ARG, "sa_mma: kpad[0]=%d too small for %d feature + xyz channels",
                     d->kpad[0], a.cpad8);
        a.twin = reinterpret_cast<const __half *>(d->twin);
        a.ldtwin = d->ldtwin;
    }
    a.rows = (long long)d->b * d->m * d->nsample;
    if (a.rows == 0) return SPSK_OK;
    const long long ntiles = (a.rows + MM_ROWS - 1) / MM_ROWS;
    SPSK_REQUIRE(ntiles <= 0x7FFFFFFF, SPSK_ERR_UNSUPPORTED, "sa_mma: too many rows");
    a.ntiles = (int)ntiles;
    a.lstages = P.lstages;
    a.sched_n = P.sched_n;
    a.rot_last = (P.L[d->nlayers - 1].n_cc > 1 && sa_tuning().rot) ? 1 : 0;
    a.l0_fused = d->l0_fused ? 1 : 0;
    a.l0_off = P.l0_off;
    SPSK_REQUIRE(!a.l0_fused || P.resident, SPSK_ERR_UNSUPPORTED, "sa_mma: layer-0 fusion needs the chain resident in shared memory");
    a.narrow = P.resident && !sa_tuning().no_narrow;
    for (int l = 0; l < d->nlayers; ++l)
        if (P.L[l].n_cc != 1) a.narrow = 0;
    a.nstages = P.nstages; a.resident = P.resident; a.w_total = P.w_total; a.tmem_cols = P.tmem_cols; a.nbuf = P.nbuf;
    a.nbuf_log2 = P.nbuf == 4 ? 2 : (P.nbuf == 2 ? 1 : 0);
    a.xa_bytes = P.xa_bytes; a.xb_bytes = P.xb_bytes;
    a.xyz = d->xyz; a.new_xyz = d->new_xyz; a.idx = d->idx;
    a.wtiles = reinterpret_cast<const uint8_t *>(d->wtiles); a.bias = d->bias;
    SPSK_REQUIRE((reinterpret_cast<uintptr_t>(d->wtiles) & 15) == 0, SPSK_ERR_INVALID_ARG, "sa_mma: wtiles must be 16-byte aligned");
    a.cout_last = d->cout_last;
    SPSK_REQUIRE(d->cout_last >= 1 && d->cout_last <= d->cpad[d->nlayers - 1], SPSK_ERR_INVALID_ARG, "sa_mma: cout_last outside the last layer");
    a.out = d->out_cm; a.c_total = d->c_total; a.co_off = d->co_off;
    if (a.out) SPSK_REQUIRE(d->co_off >= 0 && d->co_off + d->cout_last <= d->c_total, SPSK_ERR_INVALID_ARG, "sa_mma: channel window outside c_total");
    a.out16 = reinterpret_cast<__half *>(d->out16); a.ld16 = d->ld16; a.co16 = d->co16; a.n16 = d->n16; a.o16lo = d->o16lo;
    if (a.out16 && d->o16lo)
        SPSK_REQUIRE(d->o16lo > 0 && d->o16lo + d->co16 + d->n16 <= d->ld16, SPSK_ERR_INVALID_ARG, "sa_mma: residual columns outside ld16");
    if (a.out16)
        SPSK_REQUIRE(d->n16 >= d->cout_last && d->n16 <= d->cpad[d->nlayers - 1] && d->co16 >= 0 && d->co16 + d->n16 <= d->ld16, SPSK_ERR_INVALID_ARG,
                     "sa_mma: fp16 output window [co16, co16 + n16) outside ld16 or wider than the last layer");
    a.prof = g_sa_prof;
    a.ovf = fp16_overflow_word();
    a.ovf_bit = 1u << (d->ovf_tag & 31);
    a.stats = d->stats;
    a.abl = sa_tuning().abl;
    if (d->pair) {
        a.ntiles = (int)((a.rows + 255) / 256);   // 256-row tiles, one per CTA pair
        return spsk_sa_mma_pair_launch(a, P.smem, as_stream(stream));
    }
    const int mult = sa_tuning().grid_mult;
    const int slots = SPSK_NUM_SMS * P.ctas * mult;
    const int grid = a.ntiles < slots ? a.ntiles : slots;
    if (a.sched_n > 0) {
        a.sched_n = build_schedule(a, !sa_tuning().no_double);   // 0: does not fit the table, the general issue loop takes over
        if (a.sched_n > 0) a.rot_last = 0;   // the table fixes the chunk order
    }
    const bool two_groups = P.ctas == 1 && !sa_tuning().one_group;
    if (a.stats) {
        SPSK_REQUIRE((reinterpret_cast<uintptr_t>(d->stats) & 15) == 0, SPSK_ERR_INVALID_ARG, "sa_mma: stats must be 16-byte aligned");
        SPSK_REQUIRE(d->stats_parts >= grid * (two_groups ? 2 : 1), SPSK_ERR_INVALID_ARG, "sa_mma: stats holds %d parts, this launch writes %d",
                     d->stats_parts, grid * (two_groups ? 2 : 1));
        a.out = nullptr; a.out16 = nullptr;
        // the slices this launch accumulates into start from zero (stream-ordered; slices beyond them are not touched)
        const size_t bytes = (size_t)grid * (two_groups ? 2 : 1) * (size_t)P.L[d->nlayers - 1].cpad * 2 * sizeof(double);
        cudaError_t e = cudaMemsetAsync(d->stats, 0, bytes, as_stream(stream));
        if (e != cudaSuccess) return cuda_fail(e, "sa_mma: cudaMemsetAsync(stats)");
    }
#define SPSK_SA_LAUNCH(GV, PV)                                                                                               \
    do {                                                                                                                    \
        static SmemAttrOnce attr;                                                                                           \
        if (int rc = attr.ensure(reinterpret_cast<const void *>(sa_mma_kernel<GV, PV, false>), 227 * 1024, "sa_mma_kernel")) return rc; \
        sa_mma_kernel<GV, PV, false><<<grid, 128 * GV + 64, P.smem, as_stream(stream)>>>(a);                                    \
    } while (0)
    if (P.ctas == 4) {   // resident narrow chain, no producer warp
        SPSK_REQUIRE(a.narrow && a.resident, SPSK_ERR_UNSUPPORTED, "sa_mma: the 4-CTA shape needs the resident narrow issue loop (unset SPSK_SA_NO_NARROW)");
        static SmemAttrOnce attr4, attr4p;
        if (a.prof) {
            if (int rc = attr4p.ensure(reinterpret_cast<const void *>(sa_mma_kernel<1, true, true>), 227 * 1024, "sa_mma_kernel<np,prof>")) return rc;
            sa_mma_kernel<1, true, true><<<grid, 160, P.smem, as_stream(stream)>>>(a);
        } else {
            if (int rc = attr4.ensure(reinterpret_cast<const void *>(sa_mma_kernel<1, false, true>), 227 * 1024, "sa_mma_kernel<np>")) return rc;
            sa_mma_kernel<1, false, true><<<grid, 160, P.smem, as_stream(stream)>>>(a);
        }
    } else if (two_groups) {
        if (a.prof) SPSK_SA_LAUNCH(2, true); else SPSK_SA_LAUNCH(2, false);
    } else {
        if (a.prof) SPSK_SA_LAUNCH(1, true); else SPSK_SA_LAUNCH(1, false);
    }
#undef SPSK_SA_LAUNCH
    SPSK_LAUNCH_CHECK("sa_mma_kernel");
    return SPSK_OK;
}
