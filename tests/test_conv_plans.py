"""Host-side tiling of the TMA-fed convolution family (conv_tma.cu), no GPU: sweep the layer shapes of the four
networks (plus random ones) through `ccb_debug_conv_plan` and check the invariants the kernels rely on: shared-memory
and TMEM budgets, complete coverage of K / channels / taps, TMA box limits, 16-byte aligned slab rows."""
import ctypes as C
import itertools
import random

import __graft_entry__ as ge

SMEM_MAX = 227 * 1024
FPROP, DGRAD, WGRAD = 0, 1, 2


def _lib():
    from cc_b200 import _lib as L
    lib = C.CDLL(ge.build())
    lib.ccb_debug_conv_plan.restype = C.c_int
    lib.ccb_debug_conv_plan.argtypes = [C.POINTER(L.ConvDesc), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int)]
    return lib, L


def _desc(L, B, Ci, H, W, Co, k, s, p):
    d = L.ConvDesc()
    d.B, d.Ci, d.Hi, d.Wi, d.Co = B, Ci, H, W, Co
    d.Ho, d.Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
    d.kh = d.kw = k
    d.stride, d.pad, d.act, d.slope, d.impl = s, p, 0, 0.0, 0
    return d


def _plan(lib, d, op, py=0, px=0):
    out = (C.c_int * 16)()
    assert lib.ccb_debug_conv_plan(C.byref(d), op, py, px, out) == 0
    return list(out)


# (B, Ci, H, W, Co, k, s, p): every distinct conv of DispResNet6 / PoseNetB6 / MaskNet6 / Back2Future at 256x832, b4
NET_SHAPES = [
    (4, 3, 256, 832, 32, 7, 2, 3), (4, 32, 128, 416, 32, 7, 1, 3), (4, 32, 128, 416, 64, 3, 2, 1), (4, 64, 64, 208, 64, 3, 1, 1),
    (4, 32, 128, 416, 64, 1, 2, 0), (4, 64, 64, 208, 128, 3, 2, 1), (4, 128, 32, 104, 128, 3, 1, 1), (4, 128, 32, 104, 256, 3, 2, 1),
    (4, 256, 16, 52, 256, 3, 1, 1), (4, 256, 16, 52, 512, 3, 2, 1), (4, 512, 8, 26, 512, 3, 1, 1), (4, 512, 8, 26, 512, 3, 2, 1),
    (4, 512, 4, 13, 512, 3, 1, 1), (4, 512, 4, 13, 512, 3, 2, 1), (4, 512, 2, 7, 512, 3, 1, 1), (4, 1024, 4, 13, 512, 3, 1, 1),
    (4, 1024, 8, 26, 512, 3, 1, 1), (4, 512, 16, 52, 256, 3, 1, 1), (4, 256, 32, 104, 128, 3, 1, 1), (4, 129, 64, 208, 64, 3, 1, 1),
    (4, 65, 128, 416, 32, 3, 1, 1), (4, 17, 256, 832, 16, 3, 1, 1), (4, 16, 256, 832, 1, 3, 1, 1), (4, 32, 128, 416, 1, 3, 1, 1),
    (4, 64, 64, 208, 1, 3, 1, 1), (4, 128, 32, 104, 1, 3, 1, 1), (4, 15, 256, 832, 16, 7, 2, 3), (4, 16, 128, 416, 32, 5, 2, 2),
    (4, 32, 64, 208, 64, 3, 2, 1), (4, 64, 32, 104, 128, 3, 2, 1), (4, 128, 16, 52, 256, 3, 2, 1), (4, 256, 8, 26, 256, 3, 2, 1),
    (4, 256, 4, 13, 256, 3, 2, 1), (4, 256, 1, 4, 24, 1, 1, 0), (4, 65, 128, 416, 32, 1, 1, 0), (4, 17, 256, 832, 16, 1, 1, 0),
    (1, 6, 256, 832, 16, 3, 2, 1), (1, 115, 32, 104, 128, 3, 1, 1), (1, 81, 64, 208, 96, 3, 1, 1), (1, 565, 8, 26, 128, 3, 1, 1),
]


def _random_shapes(n, seed=3):
    rnd = random.Random(seed)
    out = []
    while len(out) < n:
        k = rnd.choice([1, 3, 5, 7])
        s = rnd.choice([1, 2])
        p = rnd.choice([0, k // 2])
        B, Ci, Co = rnd.randint(1, 6), rnd.randint(1, 300), rnd.randint(1, 300)
        H, W = rnd.randint(k, 120), rnd.randint(max(k, 8), 240)
        if (H + 2 * p - k) // s + 1 >= 1 and (W + 2 * p - k) // s + 1 >= 1:
            out.append((B, Ci, H, W, Co, k, s, p))
    return out


def _check_gather_plan(v, Cc, N, in_stride, what):
    kind = v[0]
    ntaps, sx, sy = v[13], v[14], v[15]
    if kind == 2:      # slab
        _, cs, cblocks, kt_full, ktiles, SW, SH, slab_bytes, nslab, nst, nbox, smem, mt = v[:13]
        stage = 2 * 16384 + 2 * nbox * 128
        assert cs * cblocks >= Cc and cs * (cblocks - 1) < Cc, what
        assert kt_full * 32 >= cs * max(ntaps, 1), what
        tail = Cc - (cblocks - 1) * cs
        assert ktiles == (cblocks - 1) * kt_full + -(-tail * max(ntaps, 1) // 32), what
        assert SW % 4 == 0 and SW >= 31 * in_stride + sx + 1 + 3 and SW <= 256, what
        assert SH == (4 * mt - 1) * in_stride + sy + 1 and SH <= 256 and cs <= 256, what
        assert slab_bytes >= cs * SH * SW * 4 and nslab == (2 if cblocks > 1 else 1), what
        assert nst >= 2 and nst <= 6 and nst * stage + nslab * slab_bytes <= smem <= SMEM_MAX, what
        assert nbox % 16 == 0 and nbox >= min(N, 128) and 2 * nbox <= 256, what
        assert 1 <= mt <= 4 and mt * 3 * nbox <= 512 and (mt == 1 or cblocks == 1), what
    elif kind == 4:    # direct
        _, CC, nchunks, NG, TH, SW, SH, slab_bytes, _, _, n8, smem, nblocks = v[:13]
        assert CC * nchunks >= Cc and CC * (nchunks - 1) < Cc, what
        assert NG * TH == 32 and n8 == NG * 8 and n8 * nblocks >= N, what
        assert SW % 4 == 0 and SW >= 31 * in_stride + sx + 1 + 3 and SW <= 256, what
        assert SH == (TH - 1) * in_stride + sy + 1 and SH <= 256 and CC <= 256, what
        assert slab_bytes >= CC * SH * SW * 4 and smem >= slab_bytes + ntaps * CC * n8 * 4 and smem <= 100 * 1024, what
    else:
        assert kind in (3, -1), what


def test_plans_cover_the_problem():
    lib, L = _lib()
    kinds = set()
    for shp in NET_SHAPES + _random_shapes(300):
        B, Ci, H, W, Co, k, s, p = shp
        d = _desc(L, *shp)
        v = _plan(lib, d, FPROP)
        _check_gather_plan(v, Ci, Co, s, ('fprop', shp))
        kinds.add(v[0])
        for py, px in itertools.product(range(min(s, H)), range(min(s, W))):
            v = _plan(lib, d, DGRAD, py, px)
            _check_gather_plan(v, Co, Ci, 1, ('dgrad', py, px, shp))
        v = _plan(lib, d, WGRAD)
        if v[0] == 5:
            _, cwid, cblocks, tpt, tgroups, SW, SH, slab_bytes, _, nst, nbox, smem, splits, KK, dx0, stages = v
            what = ('wgrad', shp)
            assert cwid * cblocks >= Ci and cwid <= 128 and tpt * cwid <= 128 and tpt * tgroups >= KK, what
            assert SW % 4 == 0 and SW >= 31 * s + k + dx0 and SW <= 256 and 0 <= dx0 < 4 and (dx0 - (-p)) % 4 == 0, what
            assert 1 <= SH <= k and slab_bytes >= cwid * SH * SW * 4, what
            stage = 2 * (16384 + nbox * 128) + slab_bytes
            assert nst >= 2 and nst * stage <= smem <= SMEM_MAX and 3 * nbox <= 512, what
            assert stages == B * d.Ho * (-(-d.Wo // 32)) and 1 <= splits <= max(1, stages), what
        else:
            assert v[0] == -1
    assert {2, 3, 4} <= kinds            # the sweep reaches the slab, the aligned and the direct kernels


def test_channels_last_plans():
    """The channels-last slab kernel's tiling (conv_nhwc.cu) over the same sweep: slab geometry, shared-memory / TMEM
    budgets of the one- and two-CTA-per-SM configurations, TMA box limits, K coverage."""
    lib, L = _lib()
    taken = 0
    for shp in NET_SHAPES + _random_shapes(300, seed=5):
        B, Ci, H, W, Co, k, s, p = shp
        d = _desc(L, *shp)
        cases = [(16 + FPROP, 0, 0, Ci, Co)] if s == 1 else []
        cases += [(16 + DGRAD, py, px, Co, Ci) for py, px in itertools.product(range(min(s, H)), range(min(s, W)))]
        for op, py, px, Cc, N in cases:
            v = _plan(lib, d, op, py, px)
            if v[0] != 6:
                assert v[0] == -1
                continue
            taken += 1
            what = (op, py, px, shp)
            _, cblocks, SH, SW, mt, slab_bytes, slab_tx, nslab, nst, nbox, ntile_w, tmem_cols, smem, Kp, spans, ctas = v
            sx, sy = spans >> 16, spans & 0xFFFF
            assert Cc >= 32 and N >= 16 and cblocks * 32 >= Cc > (cblocks - 1) * 32, what
            assert SW == 8 + sx <= 16 and SH == 16 * mt + sy <= 256, what            # TMA box (32 ch, SW px, SH rows)
            assert slab_tx == SH * SW * 128 and slab_bytes >= slab_tx and slab_bytes % 1024 == 0, what
            assert ntile_w in (64, 128) and tmem_cols == (256 if ntile_w == 64 else 512), what
            assert nbox % 16 == 0 and min(N, ntile_w) <= nbox <= ntile_w and mt * 3 * nbox <= tmem_cols, what
            assert 1 <= mt <= 4 and nslab in (1, 2) and (nslab == 1 or cblocks > 1), what
            assert 2 <= nst <= 8 and (nst >= 3 or ntile_w == 128), what
            assert smem == 2 * nslab * slab_bytes + nst * 2 * nbox * 128 + 3072, what
            assert smem <= (113 * 1024 if ntile_w == 64 else SMEM_MAX), what            # two CTAs per SM need <= half the SM
            assert Kp % 32 == 0 and Kp >= cblocks * 32 and ctas >= 1, what
    assert taken > 100
