"""Work lists of the 2-CTA GEMM (stream-K tail, half-width tail, supertile raster) checked on the HOST: the enumerator the
three roles of a CTA walk on the device (`pair_next_item` in csrc/gemm.cu) is compiled for the host as well, and
`fgb_gemm_schedule_check` walks it for every cluster of a launch — every (tile, K-block, column half) exactly once, one owner
per split tile, the owner's wait list equal to the clusters that dump a partial, split tiles last. A wrong list is a wrong
result at best and a spinning owner at worst, so this runs without a GPU in the CPU suite."""
import ctypes

import pytest

from fairygen_b200 import _lib

STEP_SHAPES = [(9216, 3072), (3072, 3072), (14336, 3072), (3072, 14336)]     # q|k|v, o / cross, FFN1, FFN2 (n, k)


def check(m, n, k, sms=148, ws=1, min_k=6144):
    lib = _lib.lib()
    splits, halves = ctypes.c_int32(-1), ctypes.c_int32(-1)
    rc = lib.fgb_gemm_schedule_check(m, n, k, sms, ws, min_k, ctypes.byref(splits), ctypes.byref(halves))
    assert rc == 0, (m, n, k, sms, ws, min_k, lib.fgb_last_error().decode())
    return splits.value, halves.value


@pytest.mark.parametrize("rows", [27280, 13640, 6820, 3410, 8190, 4095, 2048, 1024, 700, 256])
def test_step_shapes_at_every_rank_size(rows):
    for n, k in STEP_SHAPES:
        for ws in (0, 1):
            check(rows, n, k, ws=ws)
            check(rows, n, k, ws=ws, min_k=0)       # what the GPU tests force: short-K tiles split too


def test_the_documented_cases():
    # FFN2 on an SP4 rank: 324 tiles = 4 waves of 74 + 28 -> stream-K (K = 14336 >= 6144); the K = 3072 projections of the same
    # rank take the half-width tail instead (28 tiles -> 56 items); without a workspace FFN2 does, too
    assert check(6820, 3072, 14336) == (28, 0)
    assert check(6820, 3072, 3072) == (0, 56)
    assert check(6820, 3072, 14336, ws=0) == (0, 56)
    # headline rows: 1284 tiles = 17 waves + 26; q|k|v 3852 = 52 waves + 4
    assert check(27280, 3072, 3072) == (0, 52)
    assert check(27280, 9216, 3072) == (0, 8)
    assert check(27280, 3072, 14336)[0] == 26
    # a tail that fills more than 0.9 of the clusters (FFN1: 5992 = 80 waves + 72) stays whole
    assert check(27280, 14336, 3072, min_k=0) == (0, 0)


def test_sweep_of_shapes_and_machine_sizes():
    n_checked = 0
    for sms in (148, 132, 64, 8, 2):
        for m in (256, 257, 511, 1000, 2561, 4100, 6820, 9999, 19000, 27280):
            for n in (8, 200, 256, 768, 1088, 3072, 9216):
                for k in (64, 200, 512, 1536, 4096, 14336):
                    for ws, min_k in ((0, 6144), (1, 6144), (1, 0)):
                        check(m, n, k, sms=sms, ws=ws, min_k=min_k)
                        n_checked += 1
    assert n_checked == 5 * 10 * 7 * 6 * 3


def check_attn(s_q, s_kv, heads, sms=148, ws=1):
    lib = _lib.lib()
    split, n_split = ctypes.c_int32(-1), ctypes.c_int32(-1)
    rc = lib.fgb_attn_schedule_check(s_q, s_kv, heads, sms, ws, ctypes.byref(split), ctypes.byref(n_split))
    assert rc == 0, (s_q, s_kv, heads, sms, ws, lib.fgb_last_error().decode())
    return split.value, n_split.value


def test_attention_work_lists():
    """The persistent attention kernel's list [whole units | split units x key chunks] (plan_split + decode_item): every
    (256-query unit, KV tile) once, for the headline shape on 1 / 2 / 4 / 8-way head splits and a sweep of ragged shapes."""
    # headline self-attention: 107 pairs x 24 heads = 2568 units = 17 waves of 148 + 52 -> the tail is cut along the keys
    split, n = check_attn(27280, 27280, 24)
    assert n == 2568 % 148 and split > 1
    # 3 heads per rank (Ulysses SP8): 321 units = 2 waves + 25
    split, n = check_attn(27280, 27280, 3)
    assert n == 321 % 148 and split >= 2
    assert check_attn(27280, 512, 24) == (1, 0)              # cross-attention: 4 KV tiles, nothing to split
    assert check_attn(27280, 27280, 24, ws=0) == (1, 0)       # no workspace, no split
    for sms in (148, 132, 16, 1):
        for s_q in (1, 255, 256, 257, 4100, 8190, 27280):
            for s_kv in (1, 127, 128, 129, 512, 2047, 2048, 4100, 27280):
                for heads in (1, 2, 3, 6, 12, 24):
                    for ws in (0, 1):
                        check_attn(s_q, s_kv, heads, sms=sms, ws=ws)
