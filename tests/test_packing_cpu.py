"""Host logic of the tensor-core path on CPU: the packed weight images must follow the layouts include/spsk.h documents
(canonical K-major no-swizzle UMMA tiles: half(r, k) = (r/8)*(kw*8) + (k/8)*64 + (r%8)*8 + (k%8)), for the plain, split
(hi + lo) and CTA-pair packings, and the launch plans the library picks for the IA-SSD chains are the documented ones."""
import numpy as np
import pytest
import torch

from spsnet_b200 import pointnet2_utils as pu


def _chain(c_feat, widths, seed=0):
    g = torch.Generator().manual_seed(seed)
    cin, chain = c_feat + 3, []
    for w in widths:
        chain.append((torch.randn(cin, w, generator=g), torch.randn(w, generator=g), True))
        cin = w
    return chain


def _untile(flat, rows, kw):
    """inverse of the canonical layout: flat fp16 (rows*kw) -> (rows, kw)"""
    a = flat.reshape(rows // 8, kw // 8, 8, 8)          # (rg, kg, r, k)
    return a.permute(0, 2, 1, 3).reshape(rows, kw)


def _expected_w(pk, chain, l):
    """zero-padded (kin, cpad) fp32 weight of layer l in the kernel's k order"""
    wt = chain[l][0]
    W = torch.zeros(pk.kpad[l], pk.cpad[l])
    if l == 0:
        c_feat, xr = pk.c_feat, 3
        xo = (8 if c_feat else 0) if pk.split else pk.cpad8
        if c_feat:
            W[0:c_feat, :wt.shape[1]] = wt[xr:xr + c_feat]
        W[xo:xo + 3, :wt.shape[1]] = wt[0:3]
    else:
        W[:wt.shape[0], :wt.shape[1]] = wt
    return W


@pytest.mark.parametrize("c_feat,widths,pair", [(64, [64, 96, 128], False), (256, [256, 512, 1024], False), (1, [32, 32, 64], False),
                                                (0, [16, 32], False), (256, [256, 512, 1024], True), (128, [128, 256, 256], True)])
def test_mma_chain_packing_layout(c_feat, widths, pair):
    chain = _chain(c_feat, widths)
    pk = pu.MmaChain(chain, c_feat, True, pair=pair)
    assert pk.ok and pk.pair == pair
    flat = pk.wtiles[: sum((2 if pk.split else 1) * k * c for k, c in zip(pk.kpad, pk.cpad))]
    off = 0
    for l in range(pk.nlayers):
        W = _expected_w(pk, chain, l)
        if pk.split:
            Wh = W.half()
            Wv = torch.cat([Wh, (W - Wh.float()).half()], 0)
        else:
            Wv = W.half()
        vk, cp = Wv.shape
        last = l == pk.nlayers - 1
        chunk = 256 if pair else 128
        for c0 in range(0, cp, chunk):
            cw = min(chunk, cp - c0)
            parts = [(c0 + rk * (cw // 2), cw // 2) for rk in range(2)] if pair else [(c0, cw)]
            for r0, nr in parts:
                for k0 in range(0, vk, 64):
                    kw = min(64, vk - k0)
                    got = _untile(flat[off:off + nr * kw], nr, kw)
                    want = Wv[k0:k0 + kw, r0:r0 + nr].t()
                    assert torch.equal(got, want), f"layer {l} chunk {c0} k {k0}"
                    off += nr * kw
        if pair and last:
            assert cp % 256 == 0
    assert off == flat.numel()
    # biases: per layer cpad floats, zero padded
    boff = 0
    for l, (wt, b, _) in enumerate(chain):
        assert torch.equal(pk.bias[boff:boff + wt.shape[1]], b) and torch.all(pk.bias[boff + wt.shape[1]:boff + pk.cpad[l]] == 0)
        boff += pk.cpad[l]


def test_split_weights_reconstruct_fp32():
    chain = _chain(1, [32, 32, 64], seed=3)
    pk = pu.MmaChain(chain, 1, True)
    assert pk.split and pk.kpad == [16, 32, 32] and pk.cpad == [32, 32, 128]
    W = _expected_w(pk, chain, 1)
    kp, cp = pk.kpad[1], pk.cpad[1]
    off = 2 * pk.kpad[0] * pk.cpad[0]
    tiles = pk.wtiles[off:off + 2 * kp * cp]
    # packed K = 64 -> one k tile of cp rows = [Wh ; Wl] (Wh is stored once and read by two of the three products)
    Wv = _untile(tiles[: cp * 64], cp, 64).t().float()          # (64, cp)
    assert (Wv[:kp] + Wv[kp:] - W).abs().max() <= 2.0 ** -20 * W.abs().max()


@pytest.mark.parametrize("c_in,c_out,split", [(96, 64, False), (1536, 512, True), (256, 3, True), (16, 200, False)])
def test_pw_layer_packing_layout(c_in, c_out, split):
    g = torch.Generator().manual_seed(c_in)
    wt, b = torch.randn(c_in, c_out, generator=g), torch.randn(c_out, generator=g)
    L = pu.PwLayer(wt, b, True, split=split)
    k64 = (L.k + 63) // 64 * 64
    ncov = L.bias.numel()
    assert ncov % 128 == 0 and ncov >= max(c_out, L.n16) and L.k % 16 == 0
    Wt = torch.zeros(ncov, k64)
    Wt[:c_out, :c_in] = wt.t()
    Wh = Wt.half()
    parts = [Wh, (Wt - Wh.float()).half()] if split else [Wh]
    per = 128 * 64
    off = 0
    for cc in range(ncov // 128):
        for kc in range(k64 // 64):
            for P in parts:
                got = _untile(L.wtiles[off:off + per], 128, 64)
                assert torch.equal(got, P[cc * 128:(cc + 1) * 128, kc * 64:(kc + 1) * 64])
                off += per
    assert off == L.wtiles.numel()
    assert torch.equal(L.bias[:c_out], b) and torch.all(L.bias[c_out:] == 0)


def test_launch_plans_of_the_iassd_chains():
    """spsk_sa_mma_config (host-only): narrow chains resident with 3 CTAs per SM and split arithmetic; wide chains stream."""
    plans = {}
    for name, (cf, w) in {"l0s2": (1, [32, 32, 64]), "l1s1": (64, [64, 64, 128]), "l2s2": (128, [128, 256, 256]),
                          "l5s2": (256, [256, 512, 1024])}.items():
        pk = pu.MmaChain(_chain(cf, w), cf, True)
        assert pk.ok
        plans[name] = (pk.split, pk.resident, pk.ctas_per_sm, pk.nstages)
    assert plans["l0s2"][:3] == (True, 1, 4)   # 160-thread no-producer CTAs, four per SM
    assert plans["l1s1"][:3] == (False, 1, 3)
    assert plans["l2s2"][:3] == (False, 0, 1) and plans["l2s2"][3] >= 6
    assert plans["l5s2"][:3] == (False, 0, 1) and plans["l5s2"][3] >= 2
    # a chain whose activations cannot fit is refused, not mis-launched
    big = pu.MmaChain(_chain(256, [1024, 1024, 1024]), 256, True)
    assert not big.ok
