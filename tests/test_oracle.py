"""CPU tests (no GPU): pin the oracle.

The reference holds no golden vectors, so the oracle is pinned by
  (1) the second-source vectors of SURVEY.md §8(c) (an independent restatement, stock-decoder checked),
  (2) the reference's own assertions restated: round trips, size inequalities, error kinds,
  (3) stock decoders (liblz4.so.1 / pyarrow) accepting its output — the in-container stand-in for the
      `lz4` CLI the reference shells out to (src/test_compat.zig:141-254),
  (4) XXH32 against libxxhash / python-xxhash.
"""
import ctypes as C
import os

import pytest

import corpus


def _liblz4():
    try:
        return C.CDLL("liblz4.so.1")
    except OSError:
        return None


def stock_block_decode(comp, n):
    l = _liblz4()
    if l is None:
        pytest.skip("liblz4.so.1 not present")
    dst = (C.c_char * max(1, n))()
    r = l.LZ4_decompress_safe(comp, dst, len(comp), n)
    return r, bytes(dst[:max(r, 0)])


def stock_frame_decode(frame, n):
    import pyarrow as pa
    return pa.decompress(frame, decompressed_size=n, codec="lz4").to_pybytes()


# ---- (1) second-source vectors, SURVEY.md §8(c) ----
def test_second_source_block_vectors(oracle):
    o = oracle
    assert o.compress_fast(b"AAAA").hex() == "4041414141"
    assert o.compress_fast(b"A" * 160).hex() == "2f41410100865041414141" + "41"
    assert o.compress_fast(b"ABCDEFGH" * 125).hex() == "9f41424344454647484108" + "00ffffffca504445464748"
    assert o.compress_fast(b"Hello World!") == b"\xc0Hello World!"
    s = b"Hello, World! This is a test of the LZ4 compression algorithm."
    c = o.compress_fast(s)
    assert (len(s), len(c), o.xxh32(c)) == (62, 64, 0x6D410CFE)
    c = o.compress_fast(corpus.LOREM)
    assert (len(corpus.LOREM), len(c), o.xxh32(c)) == (191, 191, 0xF827CCF4)
    c = o.compress_fast(bytes(i % 256 for i in range(10000)))
    assert (len(c), o.xxh32(c)) == (306, 0xB77363DF)
    big = bytes(i % 256 for i in range(100000))
    c = o.compress_fast(big[:65536])
    assert (len(c), o.xxh32(c)) == (523, 0x3FE92CD7)
    c = o.compress_fast(big[65536:])
    assert (len(c), o.xxh32(c)) == (402, 0x09E24334)
    c = o.compress_fast(bytes(65536))
    assert (len(c), o.xxh32(c)) == (268, 0x3468134A)
    c = o.compress_fast(bytes(4 << 20))
    assert (len(c), o.xxh32(c)) == (16460, 0x043B484A)


def test_second_source_frame_vectors(oracle):
    o = oracle
    assert o.compress_frame(b"").hex() == "04224d184040c000000000"
    assert o.xxh32(b"") == 0x02CC5D05
    # header-checksum bytes (FLG, BD) -> HC
    for flg, bd, hc in ((0x40, 0x40, 0xC0), (0x64, 0x40, 0xA7), (0x74, 0x70, 0x8E)):
        assert (o.xxh32(bytes([flg, bd])) >> 8) & 0xFF == hc


# ---- (4) XXH32 ----
def test_xxh32_against_xxhash(oracle):
    import xxhash
    rnd = os.urandom(70000)
    for n in list(range(0, 70)) + [255, 256, 1000, 4095, 65536, 70000]:
        for seed in (0, 1, 0x9E3779B1):
            assert oracle.xxh32(rnd[:n], seed) == xxhash.xxh32(rnd[:n], seed=seed).intdigest()


# ---- (2)+(3) reference test inputs: round trip, stock decoder ----
@pytest.mark.parametrize("name,data", corpus.block_cases() + corpus.compat_cases())
def test_block_roundtrip_and_stock_decoder(oracle, name, data):
    c = oracle.compress_fast(data)
    assert len(c) <= oracle.compress_bound(len(data))
    assert oracle.decompress_safe(c, len(data)) == data          # src/test.zig round trips
    if data:
        r, out = stock_block_decode(c, len(data))
        assert r == len(data) and out == data
    else:
        assert c == b""                                           # src/test.zig:182-206: empty -> 0


@pytest.mark.parametrize("name,data", corpus.block_cases())
@pytest.mark.parametrize("level", [3, 6, 8, 9])
def test_hc_roundtrip_and_stock_decoder(oracle, name, data, level):
    c = oracle.compress_hc(data, level)
    assert oracle.decompress_safe(c, len(data)) == data
    if data:
        r, out = stock_block_decode(c, len(data))
        assert r == len(data) and out == data


def test_hc_reference_assertions(oracle):
    o = oracle
    rep = b"ABCD" * 500
    assert len(o.compress_hc(rep, 9)) < len(rep) // 10                      # test_lz4hc.zig:62-95
    rnd = dict(corpus.block_cases())["random1000"]
    assert len(o.compress_hc(rnd, 9)) >= len(rnd)                            # test_lz4hc.zig:123-153
    for lvl in (0, 1, -5):                                                   # level < 2 -> 9 (lz4hc.zig:1445)
        assert o.compress_hc(rep, lvl) == o.compress_hc(rep, 9)
    for lvl in (2, 10, 11, 12, 99):                                          # not restated (out of scope)
        with pytest.raises(o.OracleError) as e:
            o.compress_hc(rep, lvl)
        assert e.value.code == o.UnsupportedLevel
    # pattern analysis (level 9) never loses to level 8 on 1/2/4-byte patterns (test_lz4hc.zig:271-325)
    for pat in (b"A", b"AB", b"ABCD"):
        d = pat * 1000
        assert len(o.compress_hc(d, 9)) <= len(o.compress_hc(d, 8)) + 4


def test_hc_f8_guard(oracle):
    d = corpus.f8_hazard_input()
    before = oracle.lib().b2o_hc_f8_guard_hits()
    c = oracle.compress_hc(d, 9)
    assert oracle.decompress_safe(c, len(d)) == d
    assert oracle.lib().b2o_hc_f8_guard_hits() >= before   # guard may or may not fire; must never crash


def test_decompress_error_kinds(oracle):
    o = oracle
    good = o.compress_fast(b"A" * 160)
    for bad, code in (
        (b"\xf0", o.CorruptedData),                       # LL extension runs off the input (:125)
        (b"\x40AB", o.CorruptedData),                     # literal run longer than input (:136)
        (b"\x10A\x00\x00", o.CorruptedData),              # offset 0 (:154)
        (b"\x10A\x05\x00", o.CorruptedData),              # offset > op without dict (:231/:183)
        (b"\x1fA\x01\x00", o.CorruptedData),              # ML extension runs off (:162)
        (b"\x10A\x01", o.CorruptedData),                  # truncated offset (:149)
    ):
        with pytest.raises(o.OracleError) as e:
            o.decompress_safe(bad, 1000)
        assert e.value.code == code, bad
    with pytest.raises(o.OracleError) as e:
        o.decompress_safe(good, 100)                       # OutputTooSmall (:137/:174)
    assert e.value.code == o.OutputTooSmall
    assert o.decompress_safe(good, 0) == b""               # dst.len == 0 -> 0, no error (:98)
    assert o.decompress_safe(b"", 10) == b""               # src.len == 0 -> 0 (:97)
    with pytest.raises(o.OracleError) as e:
        o.compress_fast(b"A" * 160, cap=11)                # one byte short
    assert e.value.code == o.OutputTooSmall
    assert o.compress_fast(b"A" * 160, cap=12) == good


def test_dictionary_decode(oracle):
    """decompressSafeUsingDict semantics (src/lz4.zig:180-228,960-964): hand-built blocks whose matches
    start in the dictionary, end in it, or straddle into dst[0..]."""
    o = oracle
    d = b"0123456789abcdef"
    # 2 literals "XY", match offset 10 (op=2 -> 8 bytes back into dict), length 4+2=6, then final literals
    blk = bytes([0x22]) + b"XY" + bytes([10, 0]) + bytes([0x30]) + b"end"
    # op=2: offset 10 -> starts 8 before dst start = dict[-8:] = "89abcdef", take 6 -> "89abcd"
    assert o.decompress_safe(blk, 100, dict=d) == b"XY" + b"89abcd" + b"end"
    # straddle: offset 4 at op=2 -> 2 from dict ("ef") then continues at dst[0..] "XY" and overlaps
    blk = bytes([0x24]) + b"XY" + bytes([4, 0]) + bytes([0x10]) + b"!"
    assert o.decompress_safe(blk, 100, dict=d) == b"XY" + b"efXYefXY" + b"!"
    with pytest.raises(o.OracleError) as e:
        o.decompress_safe(blk, 100)                        # same block without the dict: CorruptedData
    assert e.value.code == o.CorruptedData
    blk = bytes([0x20]) + b"XY" + bytes([30, 0])           # offset beyond dict start (:190)
    with pytest.raises(o.OracleError) as e:
        o.decompress_safe(blk, 100, dict=d)
    assert e.value.code == o.CorruptedData


# ---- frames ----
PREFS = [
    dict(),
    dict(block_mode=1),
    dict(content_checksum=1),
    dict(block_checksum=1),
    dict(block_mode=1, block_checksum=1, content_checksum=1, content_size=12345),
    dict(block_size_id=5, block_mode=1), dict(block_size_id=6), dict(block_size_id=7, content_checksum=1),
    dict(dict_id=0xCAFE),
    dict(compression_level=9, block_mode=1), dict(compression_level=3, block_checksum=1),
]


@pytest.mark.parametrize("kw", PREFS)
@pytest.mark.parametrize("name,data", corpus.compat_cases())
def test_frame_roundtrip_and_stock(oracle, kw, name, data):
    o = oracle
    if "content_size" in kw:
        kw = dict(kw, content_size=len(data))   # the stock decoder enforces it; the reference does not
    p = o.make_prefs(**kw)
    f = o.compress_frame(data, p)
    assert len(f) <= o.compress_frame_bound(len(data), p)
    assert f[:4] == b"\x04\x22\x4d\x18"                                       # test_lz4f.zig:50-51
    assert o.decompress_frame(f, len(data) + 16) == data
    assert o.compress_frame(data, p, threads=4) == f
    assert o.decompress_frame(f, len(data) + 16, threads=4) == data
    if kw.get("dict_id", 0) == 0:
        assert stock_frame_decode(f, len(data)) == data                        # test_compat.zig G1


def test_frame_multi_block_and_errors(oracle):
    o = oracle
    data = corpus.multi_block_1mib()                                           # test_lz4f.zig:94-131
    p = o.make_prefs(block_mode=1, content_checksum=1, block_checksum=1)
    f = o.compress_frame(data, p)
    assert o.decompress_frame(f, len(data)) == data
    assert stock_frame_decode(f, len(data)) == data
    bad = bytearray(f); bad[-1] ^= 0xFF                                        # test_lz4f.zig:167-179
    with pytest.raises(o.OracleError) as e:
        o.decompress_frame(bytes(bad), len(data))
    assert e.value.code == o.F_BASE + 17                                       # ContentChecksumInvalid
    bad = bytearray(f); bad[40] ^= 0x01                                        # payload of block 0
    with pytest.raises(o.OracleError) as e:
        o.decompress_frame(bytes(bad), len(data))
    assert e.value.code == o.F_BASE + 6                                        # BlockChecksumInvalid
    with pytest.raises(o.OracleError) as e:
        o.decompress_frame(f[:len(f) // 2], len(data))
    assert e.value.code == o.F_BASE + 13                                       # FrameSizeWrong
    with pytest.raises(o.OracleError) as e:
        o.decompress_frame(f, len(data) - 1)
    assert e.value.code in (o.F_BASE + 15, o.F_BASE + 10)                      # DecompressionFailed / DstMaxSizeTooSmall
    with pytest.raises(o.OracleError) as e:
        o.compress_frame(data, p, cap=o.compress_frame_bound(len(data), p) - 1)
    assert e.value.code == o.F_BASE + 10                                       # DstMaxSizeTooSmall
    for hdr, code in ((b"\x04\x22\x4d", 11), (b"\x00\x00\x00\x00\x40\x40\xc0", 12), (b"\x04\x22\x4d\x18\x80\x40\xc0", 5),
                      (b"\x04\x22\x4d\x18\x42\x40\xc0", 7), (b"\x04\x22\x4d\x18\x40\x41\xc0", 7),
                      (b"\x04\x22\x4d\x18\x40\x10\xc0", 1), (b"\x04\x22\x4d\x18\x40\x40\xc1", 16)):
        with pytest.raises(o.OracleError) as e:
            o.decompress_frame(hdr, 10)
        assert e.value.code == o.F_BASE + code, hdr
    assert o.header_size(f) == 7
    assert o.header_size(b"\x50\x2a\x4d\x18\x00") == 8                         # skippable magic (lz4f.zig:459-462)


def test_frame_accepts_stock_encoder_output(oracle):
    """src/test_compat.zig G2 (:203-254): frames written by the stock implementation decode here."""
    import pyarrow as pa
    for name, data in corpus.compat_cases():
        f = pa.compress(data, codec="lz4", asbytes=True)
        if name == "large":
            # pyarrow writes 64 KiB *linked* blocks (FLG 0x40); the reference decoder gives every block a
            # fresh window, so a genuinely linked second block fails exactly like this.
            # (The reference's own G2 test uses the `lz4` CLI, whose default 4 MiB independent blocks keep
            # the 100 000-byte case in one block.)
            with pytest.raises(oracle.OracleError) as e:
                oracle.decompress_frame(f, len(data) + 16)
            assert e.value.code == oracle.F_BASE + 15
            continue
        assert oracle.decompress_frame(f, len(data) + 16) == data
