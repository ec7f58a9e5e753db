"""The static-ring schedule of the streaming sa_mma chains (csrc/sa_mma.cu::build_schedule), checked on the
host: the table the issue loop and the weight producer interpret is replayed for several tiles against a DYNAMIC model of the
rings and barriers (use counters per slot, completion counters per activation chunk) -- every tabulated phase parity must equal
what the dynamic bookkeeping of the previous kernel version would have computed, the ring slots must be visited round-robin, and
the entries must cover the packed weights exactly once with the right number of MMAs per job."""
import ctypes as C

import pytest
import torch

from spsnet_b200 import pointnet2_utils as pu
from spsnet_b200._lib import lib

SCH_FIRST_KC, SCH_LAST_KC, SCH_LRING_FIRST, SCH_HID_DONE, SCH_LAST_LAYER = 1 << 4, 1 << 5, 1 << 11, 1 << 12, 1 << 13
MM_MAX_STAGES, MM_MAX_XC, STAGE = 8, 16, 16384


def xr_off(buf, c):
    return 8 * (2 * MM_MAX_STAGES + 8 + buf * MM_MAX_XC + c)


CHAINS = [   # (c_feat, widths): the KITTI / Waymo streaming chains + odd shapes (ragged k tiles, ragged cout chunks, 2 and 4 layers)
    (64, [64, 96, 128]), (128, [128, 128, 256]), (128, [128, 256, 256]), (256, [256, 256, 512]), (256, [256, 512, 1024]),
    (100, [200, 304, 500]), (300, [384, 640]), (256, [256, 256, 256, 512]), (400, [256, 768]), (64, [512]),
]


def _schedule(c_feat, widths):
    cin = c_feat + 3
    chain = []
    for co in widths:
        chain.append((torch.zeros(cin, co), torch.zeros(co), True))
        cin = co
    pk = pu.MmaChain(chain, c_feat, True, pair=False)
    assert pk.ok and not pk.split
    d = pk._desc()
    n = C.c_int(0)
    words = (C.c_uint32 * (8 * 104))()
    info = (C.c_int * 6)()
    assert lib.spsk_sa_mma_schedule(C.byref(d), C.byref(n), words, info) == 0
    ents = [tuple(words[8 * e:8 * e + 8]) for e in range(n.value)]
    return pk, ents, list(info)


@pytest.mark.parametrize("c_feat,widths", CHAINS, ids=[f"c{c}-" + "x".join(map(str, w)) for c, w in CHAINS])
def test_static_schedule_replays_like_the_dynamic_rings(c_feat, widths):
    pk, ents, (nstages, lstages, resident, ring0_off, ring1_off, w_total) = _schedule(c_feat, widths)
    if resident:
        assert ents == []          # resident chains run the register-resident narrow loop or the general loop, no table
        return
    assert ents, "a streaming chain of this size must be tabulated"
    nL = len(widths)
    n_cc = [(cp + 127) // 128 for cp in pk.cpad]
    n_xc = [(k + 63) // 64 for k in pk.kpad]
    jobs_per_layer_end = [sum(n_cc[:l + 1]) for l in range(nL)]
    uses, completions = {}, {}
    slot_of = {}
    for t in range(5):                                   # five tiles: both tile parities, several wraps of every ring
        tpar = t & 1
        job, layer, k16_in_job, src_expect = 0, 0, 0, 0
        waited = set()
        ring_seq = {0: [], 1: []}
        flags_seen = {"lring_first": 0, "hid_done": 0}
        mmas = 0
        for (x_lo, idesc, hi, f, src, by, rz, rw) in ents:
            nbytes, slot_off = by & 0xFFFF, (by >> 16) * 16
            full, empty = rz & 1023, (rz >> 10) & 1023
            ring = 0 if slot_off >= ring0_off else 1
            ring_seq[ring].append(full)
            assert slot_of.setdefault(full, (slot_off, empty, ring)) == (slot_off, empty, ring)   # a barrier pair belongs to one slot
            # ---- the phase parity the dynamic bookkeeping would use: number of earlier uses of this slot
            want = uses.get(full, 0) & 1
            got = ((rz >> 20) ^ ((rz >> 21) & tpar)) & 1
            assert got == want, f"tile {t}: slot parity {got} != dynamic {want}"
            uses[full] = uses.get(full, 0) + 1
            nk16 = f & 15
            if nk16 == 0:                                # padding entry: releases the slot, nothing else
                assert nbytes == 0 and not (f & (SCH_FIRST_KC | SCH_LAST_KC | SCH_LAST_LAYER | SCH_HID_DONE)) and not (rz & (1 << 22))
                continue
            # ---- weights: every byte exactly once, in packing order; the slot holds what is copied
            assert src == src_expect
            src_expect += nbytes
            cap = (2 if ((ring == 0 and nstages >= 4) or (ring == 1 and lstages >= 4)) else 1) * STAGE
            assert nbytes <= cap and slot_off + nbytes <= 227 * 1024
            # ---- jobs and MMAs
            if f & SCH_FIRST_KC:
                assert k16_in_job == 0
            k16_in_job += nk16
            mmas += nk16
            assert bool(f & SCH_LAST_LAYER) == (layer == nL - 1)
            flags_seen["lring_first"] += bool(f & SCH_LRING_FIRST)
            flags_seen["hid_done"] += bool(f & SCH_HID_DONE)
            # ---- activation-chunk waits: parity = completions of that barrier so far
            for i in range(2):
                if rz & (1 << (22 + 3 * i)):
                    bar = (rw >> (10 * i)) & 1023
                    assert bar in [xr_off(layer & 1, c) for c in range(n_xc[layer])] and (layer, bar) not in waited
                    waited.add((layer, bar))
                    want = completions.get(bar, 0) & 1
                    got = ((rz >> (23 + 3 * i)) ^ ((rz >> (24 + 3 * i)) & tpar)) & 1
                    assert got == want, f"tile {t} layer {layer}: chunk parity {got} != dynamic {want}"
            if f & SCH_LAST_KC:
                assert k16_in_job == pk.kpad[layer] // 16, "a job issues exactly K / 16 MMAs"
                k16_in_job = 0
                job += 1
                if job == jobs_per_layer_end[layer]:
                    # the layer is done: every input chunk was waited for once; its barriers have completed once more
                    assert {b for (l, b) in waited if l == layer} == {xr_off(layer & 1, c) for c in range(n_xc[layer])}
                    for c in range(n_xc[layer]):
                        completions[xr_off(layer & 1, c)] = completions.get(xr_off(layer & 1, c), 0) + 1
                    layer += 1
        assert layer == nL and src_expect == w_total
        assert mmas == sum(n_cc[l] * (pk.kpad[l] // 16) for l in range(nL))
        assert flags_seen["lring_first"] == flags_seen["hid_done"] == (1 if lstages else 0)
        for ring, seq in ring_seq.items():               # round-robin over the ring's slots, whole turns per tile
            if not seq:
                continue
            depth = len(set(seq))
            assert len(seq) % depth == 0 and all(seq[i] == seq[i % depth] for i in range(len(seq)))
    # slots of a ring do not overlap each other or the other ring
    spans = sorted((off, off + (2 * STAGE if ((r == 0 and nstages >= 4) or (r == 1 and lstages >= 4)) else STAGE)) for off, _, r in slot_of.values())
    assert all(a1 <= b0 for (_, a1), (b0, _) in zip(spans, spans[1:]))


def test_double_tiles_where_the_ring_has_room():
    """Layer 5 scale 2 (259 -> 256 -> 512 -> 1024): 2 hidden stages (single tiles), overlay ring of 4 stages = 2 slots of 32 KB
    (two k tiles per entry, also in the first cout chunk, whose entries then carry two activation-chunk waits): 26 + 32 entries
    instead of 90."""
    pk, ents, info = _schedule(256, [256, 512, 1024])
    assert info[0] == 2 and info[1] == 4 and not info[2]
    real = [e for e in ents if e[3] & 15]
    assert len(real) == 26 + 32 and max(e[3] & 15 for e in real) == 8
    assert sum(1 for e in real if (e[6] >> 25) & 1) == 4    # the four double entries of the last layer's first chunk wait for two chunks
    assert sum(1 for e in ents if not (e[3] & 15)) <= 2     # ring padding
