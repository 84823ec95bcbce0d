"""GPU parity tests for K3 (LZ4HC hash-chain, levels 3..9) through the C-ABI.

north_star asks for exact round trip + ratio beside the reference's; this kernel goes further and is
byte-identical to the oracle's restatement of compressHashChain (reference src/lz4hc.zig:976-1064)."""
import numpy as np
import pytest

import corpus

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("level", [3, 5, 8, 9])
@pytest.mark.parametrize("name,data", corpus.block_cases() + corpus.compat_cases())
def test_hc_bytes_equal_oracle(z, oracle, name, data, level):
    for b in ([data[i:i + 65536] for i in range(0, len(data), 65536)] or [b""]):
        want = oracle.compress_hc(b, level)
        got = z.lz4hc.compressHC(b, level)
        assert got == want, (name, level, len(got), len(want))
        assert z.lz4.decompressSafe(got, len(b)) == b


def test_hc_level_routing(z, oracle):
    rep = b"ABCD" * 500
    for lvl in (0, 1, -3):                                   # < 2 -> 9 (src/lz4hc.zig:1445)
        assert z.lz4hc.compressHC(rep, lvl) == oracle.compress_hc(rep, 9)
    for lvl in (2, 10, 11, 12, 40):                          # LZ4MID / optimal parser: outside the path
        with pytest.raises(z.B2Error) as e:
            z.lz4hc.compressHC(rep, lvl)
        assert e.value.name == "b2lz4.UnsupportedLevel"
    assert z.lz4hc.compressHC(b"", 9) == b""
    with pytest.raises(z.B2Error) as e:
        z.lz4hc.compressHC(rep, 9, dst_capacity=0)
    assert e.value.name == "lz4.OutputTooSmall"


def test_hc_synthetic_classes_and_big_blocks(z, oracle):
    from zig_lz4_b200 import datagen
    for mode in (0, 1, 2, 3):
        for n in (65536, 262144, 4096):
            d = datagen.generate(n, mode=mode, seed=mode * 7 + n).tobytes()
            want = oracle.compress_hc(d, 9)
            got = z.lz4hc.compressHC(d, 9)
            assert got == want, (mode, n, len(got), len(want))
            assert z.lz4.decompressSafe(got, n) == d
    d = corpus.f8_hazard_input()                              # guarded in oracle and kernel
    got = z.lz4hc.compressHC(d, 9)
    assert got == oracle.compress_hc(d, 9)
    assert z.lz4.decompressSafe(got, len(d)) == d


def test_hc_limited_output(z, oracle):
    from zig_lz4_b200 import datagen
    d = datagen.generate(20000, mode=0).tobytes()
    full = oracle.compress_hc(d, 9)
    for cap in (len(full) + 16, len(full) + 5, len(full), len(full) - 1, len(full) // 2):
        try:
            want = (0, oracle.compress_hc(d, 9, cap=cap))
        except oracle.OracleError as e:
            want = (e.code, None)
        try:
            got = (0, z.lz4hc.compressHC(d, 9, dst_capacity=cap))
        except z.B2Error as e:
            got = (e.code, None)
        assert got == want, cap


def test_hc_frames_and_records(z, oracle, ctx):
    """configs[4]: HC level 9 on 256 KiB blocks, and a batch of 4 KiB records"""
    from zig_lz4_b200 import datagen
    data = datagen.generate((2 << 20) + 999, mode=0).tobytes()
    zp = z.lz4f.Preferences(blockSizeID=5, blockMode=1, compressionLevel=9, blockChecksumFlag=1)
    op = oracle.make_prefs(block_size_id=5, block_mode=1, compression_level=9, block_checksum=1)
    f = z.lz4f.compressFrame(data, zp)
    assert f == oracle.compress_frame(data, op, threads=8)
    assert z.lz4f.decompressFrame(f, len(data)) == data
    recs = datagen.generate(4096 * 64, mode=0, seed=11).tobytes()
    offs = [i * 4096 for i in range(64)]
    lens = [4096] * 64
    caps = [int(z.lz4.compressBound(4096))] * 64
    doffs = [i * caps[0] for i in range(64)]
    dst, ol, st = ctx.compress_hc_batch(recs, offs, lens, caps[0] * 64, doffs, caps, level=9)
    for i in range(64):
        want = oracle.compress_hc(recs[offs[i]:offs[i] + 4096], 9)
        assert st[i] == 0 and dst[doffs[i]:doffs[i] + int(ol[i])].tobytes() == want, i


def test_hc_levels_interleaved_on_one_workspace(z, oracle):
    """The HC work areas persist between calls (epoch-tagged tables, nothing re-zeroed): every kernel variant — the
    single-chain walk of levels <= 6, the jump-table walk of levels 7..9, before and after a block has switched to jump
    mode — must leave them in a state the next one reads correctly.  Text blocks have long, far-apart chains (jump mode),
    binary ones long near chains (stay in chain mode), mixed ones switch part-way."""
    from zig_lz4_b200 import datagen
    blocks = [datagen.generate(262144, mode=m, seed=s).tobytes() for m, s in ((0, 11), (1, 12), (4, 13), (2, 14), (0, 15))]
    for rnd in range(2):
        for level in (9, 3, 7, 6, 9, 8, 4, 9):
            for b in blocks:
                assert z.lz4hc.compressHC(b, level) == oracle.compress_hc(b, level), (rnd, level)
    # a batch of many blocks at once (several work areas), then single blocks again on the same areas
    big = datagen.generate(16 << 20, mode=0, seed=21).tobytes()
    zp = z.lz4f.Preferences(blockSizeID=5, blockMode=1, compressionLevel=9)
    op = oracle.make_prefs(block_size_id=5, block_mode=1, compression_level=9)
    assert z.lz4f.compressFrame(big, zp) == oracle.compress_frame(big, op, threads=8)
    for level in (5, 9):
        assert z.lz4hc.compressHC(blocks[0], level) == oracle.compress_hc(blocks[0], level)
