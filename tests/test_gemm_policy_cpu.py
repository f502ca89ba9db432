"""The host-side tile policy of mobi_gemm, asked through the C ABI's mobi_gemm_plan (validates and decides exactly as a launch
would, builds no tensor maps, launches nothing: no GPU needed; without a device the library assumes the B200's 148 SMs).
The expected choices are the measured ones (profiles/r02/kbench_gemm_tiles.log, DESIGN.md sections 4 and 7)."""
import ctypes as C

import pytest

from mobi_b200 import _lib as L

FAKE = 0x10000  # non-null, 16-byte aligned, never dereferenced on the host


def plan(M, N, K, *, epilogue=L.EPI_PLAIN, out_f32=True, residual=False, conv=None, tile_n=0, pair=0, heads=0, head_dim=0,
         tokens=0, batch=1):
    a = L.GemmArgs()
    a.A = a.B = a.out = FAKE
    a.out2 = a.out3 = FAKE if epilogue not in (L.EPI_PLAIN, L.EPI_GEGLU2) else 0
    a.residual = FAKE if residual else 0
    a.M, a.N, a.K = M, N, K
    a.lda, a.ldb, a.ldo = K, K, N // 2 if epilogue == L.EPI_GEGLU2 else N
    a.out_dtype = L.DT_F32 if out_f32 else L.DT_BF16
    a.res_dtype = L.DT_F32
    a.epilogue, a.tile_n, a.pair, a.batch = epilogue, tile_n, pair, batch
    a.heads, a.head_dim, a.tokens = heads, head_dim, tokens
    a.rows_per_group = 1
    if conv is not None:
        n, h, w, c = conv
        a.conv, a.n_img, a.H, a.W, a.C, a.KH, a.KW, a.pad_h, a.pad_w = 1, n, h, w, c, 3, 3, 1, 1
        a.lda = c
    t, p, pers = C.c_int32(-1), C.c_int32(-1), C.c_int32(-1)
    rc = L.load().mobi_gemm_plan(C.byref(a), C.byref(t), C.byref(p), C.byref(pers))
    return rc, t.value, p.value, pers.value


def test_convolutions_with_n_a_multiple_of_320_take_wide_pair_tiles_once_they_fill_the_machine():
    # 32 UNet rows (8 joint samples x CFG): level-0 ResBlock conv, its 960 -> 320 skip variant, level 1 and level 2
    assert plan(32 * 4096, 320, 9 * 320, conv=(32, 64, 64, 320)) == (0, 160, 3, 1)
    assert plan(32 * 4096, 320, 9 * 960, conv=(32, 64, 64, 960)) == (0, 160, 3, 1)
    assert plan(32 * 1024, 640, 9 * 640, conv=(32, 32, 32, 640)) == (0, 160, 3, 1)
    assert plan(32 * 256, 1280, 9 * 1280, conv=(32, 16, 16, 1280)) == (0, 160, 3, 1)
    # 8 x 8 level at 32 rows: 8 x 4 = 32 wide tiles < 74 CTA pairs -> measured slower, not chosen (and too few 160-wide pair
    # tiles for a wave of pairs either)
    rc, t, p, pers = plan(32 * 64, 1280, 9 * 1280, conv=(32, 8, 8, 1280))
    assert (rc, t, pers) == (0, 160, 1) and p != 3
    # ... but chosen at 128 rows (config 3's micro-batch)
    assert plan(128 * 64, 1280, 9 * 1280, conv=(128, 8, 8, 1280))[2] == 3
    # a width that is no multiple of 320 (VAE: 128 / 256 / 512 channels) never is
    assert plan(8 * 65536, 128, 9 * 128, conv=(8, 256, 256, 128))[2] in (0, 1)
    assert plan(8 * 4096, 512, 9 * 512, conv=(8, 64, 64, 512))[2] in (0, 1)


def test_bf16_output_epilogues_take_256_wide_pair_tiles_when_n_fills_them():
    M = 32 * 4096
    assert plan(M, 2560, 320, epilogue=L.EPI_GEGLU2, out_f32=False) == (0, 256, 1, 1)
    assert plan(M, 960, 320, epilogue=L.EPI_QKV_ROW, out_f32=False, heads=8, head_dim=40, tokens=4096) == (0, 256, 1, 1)
    assert plan(32 * 256, 10240, 1280, epilogue=L.EPI_GEGLU2, out_f32=False) == (0, 256, 1, 1)
    # N = 320 would waste 37 % of two 256-wide tiles: stays on 160
    rc, t, p, pers = plan(M // 2, 320, 320, epilogue=L.EPI_HEADS, out_f32=False, heads=8, head_dim=40, tokens=4096)
    assert (rc, t, pers) == (0, 160, 1)


def test_f32_outputs_keep_160_wide_tiles_and_pair_up_only_for_long_k():
    M = 32 * 4096
    assert plan(M, 320, 320, residual=True) == (0, 160, 0, 1)            # to_out + residual: HBM / epilogue bound
    assert plan(M, 320, 1280, residual=True) == (0, 160, 1, 1)           # ff2 + residual: 20 k-blocks
    assert plan(4, 1280, 320)[0] == 0                                    # the time-embedding MLP: tiny M is fine


def test_forced_modes_are_honoured_or_refused_never_replaced():
    M = 32 * 4096
    assert plan(M, 320, 2880, conv=(32, 64, 64, 320), tile_n=160, pair=1) == (0, 160, 1, 1)
    assert plan(M, 320, 2880, conv=(32, 64, 64, 320), tile_n=160, pair=2) == (0, 160, 2, 1)
    assert plan(M, 320, 2880, conv=(32, 64, 64, 320), pair=-1)[2] == 0
    assert plan(M, 320, 2880, conv=(32, 64, 64, 320), tile_n=128, pair=3)[0] != 0        # wide pairs are 2 x 160 columns
    assert b"pair = 3" in L.load().mobi_last_error()
    assert plan(M, 480, 320, tile_n=160, pair=2)[0] != 0                                 # 3 n-tiles: no two pairs per cluster
    assert plan(M, 2560, 320, epilogue=L.EPI_GEGLU2, out_f32=False, tile_n=160, pair=2)[0] != 0   # PLAIN only
    assert plan(0, 320, 320)[0] != 0 and plan(M, 320, 321)[0] != 0                       # the usual argument checks still apply
