"""Invariants of the tensor encode kernel's host-side work plan (column groups, CTA ranges, shared memory), checked on
the CPU through rb_debug_tensor_plan for every instantiated subvector width and a sweep of shapes.  A plan that broke
one of these would hang or corrupt the persistent kernel, which is why it is tested without a GPU."""
import ctypes as C

import pytest

from reductive_b200._cabi import lib

DSUBS = [2, 4, 6, 8, 10, 12, 16, 20, 24, 30, 32]
SMEM_LIMIT = 227 * 1024


def _plan(M, dsub, n_tiles, sms=148):
    out = (C.c_longlong * 6)()
    starts = (C.c_ushort * 160)()
    lib.rb_debug_tensor_plan.argtypes = [C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.POINTER(C.c_longlong),
                                         C.POINTER(C.c_ushort)]
    ok = lib.rb_debug_tensor_plan(M, dsub, n_tiles, sms, out, starts)
    if not ok:
        return None
    gm, n_groups, stages, pitch_f, ctas, smem = (int(v) for v in out)
    return dict(gm=gm, n_groups=n_groups, stages=stages, pitch_f=pitch_f, ctas=ctas, smem=smem,
                starts=[int(starts[g]) for g in range(n_groups + 1)])


@pytest.mark.parametrize("dsub", DSUBS)
def test_plan_invariants(dsub):
    seen = 0
    for M in list(range(1, 41)) + [48, 64, 96, 100, 128, 150]:
        for n_tiles in (1, 2, 3, 7, 37, 147, 148, 149, 1000, 15625, 97657):
            p = _plan(M, dsub, n_tiles)
            if p is None:
                continue
            seen += 1
            gm, ng = p["gm"], p["n_groups"]
            assert 1 <= gm <= 16 and ng == -(-M // gm) and ng <= 148
            assert p["stages"] in (2, 3, 4)                   # A-operand ring depth the barriers are sized for
            assert p["smem"] <= SMEM_LIMIT - 1024
            assert p["pitch_f"] <= 256 and (p["pitch_f"] * 4) % 16 == 0 and ((p["pitch_f"] * 4) // 16) % 2 == 1
            assert p["pitch_f"] >= 2 * dsub * -(-gm // 2)      # room for pairs of subvectors
            if ng > 1:
                assert (gm * dsub) % 4 == 0                   # 16-byte aligned TMA box start
            s = p["starts"]
            assert s[0] == 0 and s[-1] == p["ctas"] <= 148
            assert all(b > a for a, b in zip(s, s[1:]))       # every group owns at least one CTA
            assert all(b - a <= n_tiles for a, b in zip(s, s[1:]))  # no CTA without a tile
            # balance: the busiest CTA is within 35% of the ideal share once there is enough work (with fewer than four
            # CTAs per group the static split is coarser: up to 3 vs 2 CTAs for equal groups)
            if n_tiles >= 1000:
                widths = [gm] * (ng - 1) + [M - gm * (ng - 1)]
                span = max(-(-n_tiles // (b - a)) * w for a, b, w in zip(s, s[1:], widths))
                slack = 1.35 if ng <= 37 else (1.6 if ng <= 50 else 2.1)  # 75 groups on 148 CTAs: 2 vs 1 CTA per group
                assert span <= slack * n_tiles * M / 148 + max(widths), (M, n_tiles, p)
    assert seen > 100


def test_plan_rejects_what_the_kernel_cannot_do():
    assert _plan(3, 5, 100) is None          # odd widths cannot be packed in pairs; the exact kernel serves them
    assert _plan(10, 10, 0) is None          # nothing to do
