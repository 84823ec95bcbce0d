"""GPU parity tests for the block API (K1 compress, K2 decompress), through the C-ABI.

Bar: integer/byte work -> bit-exact.  Compressed bytes must equal the oracle's (the restatement of
lz4.compressFast, reference src/lz4.zig:292-447); decoded bytes and error kinds must equal
decompressGeneric's (src/lz4.zig:89-251)."""
import numpy as np
import pytest

import corpus

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,data", corpus.block_cases() + corpus.compat_cases())
def test_compress_default_bytes_equal_oracle(z, oracle, name, data):
    blocks = [data[i:i + 65536] for i in range(0, len(data), 65536)] or [b""]   # configs[0]: 64 KiB blocks
    for b in blocks:
        want = oracle.compress_fast(b)
        got = z.lz4.compressDefault(b)
        assert got == want, (name, len(b), len(got), len(want))
        assert z.lz4.decompressSafe(got, len(b)) == b                              # src/test.zig round trip
        assert z.compressDefault(b) == want                                         # flat re-export (root.zig:11)


@pytest.mark.parametrize("accel", [1, 2, 5, 33, 63, 64, 65, 100, 1000, 65537, 1 << 20, 0])
def test_compress_fast_acceleration(z, oracle, accel):
    from zig_lz4_b200 import datagen
    for mode in (0, 1, 2, 3):
        d = datagen.generate(40000, mode=mode).tobytes()
        assert z.lz4.compressFast(d, accel) == oracle.compress_fast(d, accel), (mode, accel)


def test_compress_sizes_sweep(z, oracle):
    """every size 0..300 and block-boundary sizes; bytes with short periods to hit all emit paths"""
    rng = np.random.default_rng(7)
    base = rng.integers(0, 4, size=70000, dtype=np.uint8).tobytes()
    for n in list(range(0, 300)) + [4095, 4096, 4097, 65535, 65536, 65537, 69999]:
        d = base[:n]
        assert z.lz4.compressDefault(d) == oracle.compress_fast(d), n


def test_large_blocks_u32_table(z, oracle):
    """blocks > 64 KiB use the wide table; long literal runs (> 64 KiB) and long matches"""
    from zig_lz4_b200 import datagen
    for mode, n in ((0, 300000), (2, 1 << 20), (3, 200000), (4, 700001)):
        d = datagen.generate(n, mode=mode, seed=n).tobytes()
        c = z.lz4.compressDefault(d)
        assert c == oracle.compress_fast(d), (mode, n)
        assert z.lz4.decompressSafe(c, n) == d
    zeros = bytes(4 << 20)
    c = z.lz4.compressDefault(zeros)
    assert (len(c), oracle.xxh32(c)) == (16460, 0x043B484A)                        # SURVEY §8c vector
    assert z.lz4.decompressSafe(c, len(zeros)) == zeros


def test_output_too_small_is_exact(z, oracle):
    d = b"A" * 160
    good = oracle.compress_fast(d)
    assert z.lz4.compressDefault(d, dst_capacity=len(good)) == good
    with pytest.raises(z.B2Error) as e:
        z.lz4.compressDefault(d, dst_capacity=len(good) - 1)
    assert e.value.name == "lz4.OutputTooSmall"
    from zig_lz4_b200 import datagen
    t = datagen.generate(30000, mode=0).tobytes()
    want = oracle.compress_fast(t)
    for cap in (len(want), len(want) - 1, len(want) // 2, 1, 0):
        if cap >= len(want):
            assert z.lz4.compressDefault(t, dst_capacity=cap) == want
        else:
            with pytest.raises(z.B2Error) as e:
                z.lz4.compressDefault(t, dst_capacity=cap)
            assert e.value.name == "lz4.OutputTooSmall"


def test_decompress_error_kinds(z, oracle):
    good = oracle.compress_fast(b"A" * 160)
    cases = [b"\xf0", b"\x40AB", b"\x10A\x00\x00", b"\x10A\x05\x00", b"\x1fA\x01\x00", b"\x10A\x01",
             b"\xf0" + b"\xff" * 100, b"\x1f" + b"A" + b"\x01\x00" + b"\xff" * 40]
    for bad in cases:
        with pytest.raises(oracle.OracleError) as eo:
            oracle.decompress_safe(bad, 1000)
        with pytest.raises(z.B2Error) as eg:
            z.lz4.decompressSafe(bad, 1000)
        assert eg.value.code == eo.value.code, bad
    with pytest.raises(z.B2Error) as e:
        z.lz4.decompressSafe(good, 100)
    assert e.value.name == "lz4.OutputTooSmall"
    assert z.lz4.decompressSafe(good, 0) == b""             # dst.len == 0 -> 0 (src/lz4.zig:98)
    assert z.lz4.decompressSafe(b"", 10) == b""             # src.len == 0 -> 0 (src/lz4.zig:97)
    assert z.lz4.decompressSafe(good, 160) == b"A" * 160
    assert z.lz4.decompressSafe(good, 1000) == b"A" * 160   # larger dst is fine


def test_decompress_random_garbage_matches_oracle(z, oracle):
    """fuzz: random / mutated streams must give the oracle's status and bytes"""
    rng = np.random.default_rng(99)
    from zig_lz4_b200 import datagen
    base = oracle.compress_fast(datagen.generate(5000, mode=0).tobytes())
    streams = [rng.integers(0, 256, size=int(rng.integers(1, 200)), dtype=np.uint8).tobytes() for _ in range(60)]
    for _ in range(60):
        b = bytearray(base)
        for _ in range(int(rng.integers(1, 4))):
            b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        streams.append(bytes(b[:int(rng.integers(1, len(b) + 1))]))
    for s in streams:
        for cap in (6000, 300):
            try:
                want = (0, oracle.decompress_safe(s, cap))
            except oracle.OracleError as e:
                want = (e.code, None)
            try:
                got = (0, z.lz4.decompressSafe(s, cap))
            except z.B2Error as e:
                got = (e.code, None)
            assert got == want, (s.hex(), cap)


def test_decompress_stock_encoder_blocks(z):
    """blocks produced by stock liblz4 (different parser: long offsets, overlapping matches) decode"""
    import pyarrow as pa
    from zig_lz4_b200 import datagen
    for mode in (0, 1, 2, 3):
        d = datagen.generate(200000, mode=mode).tobytes()
        c = pa.compress(d, codec="lz4_raw", asbytes=True)
        assert z.lz4.decompressSafe(c, len(d)) == d
    for d in (b"ab" * 40000, b"abc" * 30000, bytes(100000), b"0123456789abcdefg" * 5000):
        c = pa.compress(d, codec="lz4_raw", asbytes=True)
        assert z.lz4.decompressSafe(c, len(d)) == d


def test_dictionary_decode(z, oracle):
    d = b"0123456789abcdef"
    blocks = [bytes([0x22]) + b"XY" + bytes([10, 0]) + bytes([0x30]) + b"end",
              bytes([0x24]) + b"XY" + bytes([4, 0]) + bytes([0x10]) + b"!",
              bytes([0x20]) + b"XY" + bytes([30, 0]),
              bytes([0x2f]) + b"XY" + bytes([18, 0]) + bytes([200]) + bytes([0x10]) + b"!"]
    for blk in blocks:
        for dic in (d, b"", d * 5000):
            try:
                want = (0, oracle.decompress_safe(blk, 1000, dict=dic))
            except oracle.OracleError as e:
                want = (e.code, None)
            try:
                got = (0, z.lz4.decompressSafeUsingDict(blk, 1000, dic))
            except z.B2Error as e:
                got = (e.code, None)
            assert got == want, (blk.hex(), len(dic))


def test_batch_host_api(z, oracle, ctx):
    """b2lz4_compress_fast_batch / decompress_safe_batch with ragged blocks incl. empty and tiny ones"""
    from zig_lz4_b200 import datagen
    data = datagen.generate(3 << 20, mode=4).tobytes()
    lens = [0, 1, 12, 13, 14, 100, 65536, 65536, 70000, 200000, 5, 4096, 4096, 33333] * 3
    offs, pos = [], 0
    for l in lens:
        offs.append(pos); pos += l
    caps = [int(z.lz4.compressBound(l)) for l in lens]
    doffs = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.uint64)
    dst, ol, st = ctx.compress_fast_batch(data, offs, lens, int(sum(caps)), doffs, caps)
    comp = []
    for i, l in enumerate(lens):
        want = oracle.compress_fast(data[offs[i]:offs[i] + l])
        assert st[i] == 0 and ol[i] == len(want), i
        got = dst[int(doffs[i]):int(doffs[i]) + int(ol[i])].tobytes()
        assert got == want, i
        comp.append(got)
    blob = b"".join(comp)
    coffs = np.concatenate([[0], np.cumsum([len(c) for c in comp])[:-1]]).astype(np.uint64)
    ucap = [l for l in lens]
    uoffs = np.array(offs, dtype=np.uint64)
    out, ol2, st2 = ctx.decompress_safe_batch(blob, coffs, [len(c) for c in comp], len(data), uoffs, ucap)
    assert (st2 == 0).all()
    assert (ol2 == np.array(lens)).all()
    assert out[:pos].tobytes() == data[:pos]


def test_xxh32_matches(z, oracle):
    import os
    rnd = os.urandom(300000)
    for n in list(range(0, 40)) + [255, 4096, 65535, 65536, 100001, 300000]:
        for seed in (0, 12345):
            assert z.lz4.xxh32(rnd[:n], seed) == oracle.xxh32(rnd[:n], seed), n
    assert z.lz4.xxh32(rnd[1:100000]) == oracle.xxh32(rnd[1:100000])


def _batch_decode_vs_oracle(ctx, oracle, streams, caps, dic=None):
    blob = b"".join(streams)
    coffs = np.concatenate([[0], np.cumsum([len(s) for s in streams])[:-1]]).astype(np.uint64)
    uoffs = np.concatenate([[0], np.cumsum(caps)[:-1]]).astype(np.uint64)
    out, ol, st = ctx.decompress_safe_batch(blob, coffs, [len(s) for s in streams], int(sum(caps)) + 1, uoffs, caps, dict=dic)
    for i, s in enumerate(streams):
        try:
            want = (0, oracle.decompress_safe(s, caps[i], dict=dic))
        except oracle.OracleError as e:
            want = (e.code, None)
        assert int(st[i]) == want[0], (i, len(s), caps[i], int(st[i]), want[0])
        if want[0] == 0:
            assert int(ol[i]) == len(want[1]), i
            assert out[int(uoffs[i]):int(uoffs[i]) + int(ol[i])].tobytes() == want[1], i


def test_decode_synthetic_streams_dense_dependencies(ctx, oracle):
    """hand-assembled streams: neighbouring sequences that feed each other (offsets inside the last few
    bytes), overlapping matches, long literal runs / matches in the middle of short ones, every batch
    fill from 1 to > 32 sequences — the cases the lane-per-sequence decoder has to order correctly"""
    rng = np.random.default_rng(2024)
    streams, caps = [], []
    for nseq in list(range(1, 70)) + [100, 257, 1000, 4000]:
        for near, p_near in ((1, 1.0), (3, 0.9), (8, 0.7), (40, 0.5), (300, 0.5), (64, 0.0)):
            s, d = corpus.synth_lz4_stream(rng, nseq, near=near, p_near=p_near,
                                           ll_max=int(rng.integers(1, 30)), ml_max=int(rng.integers(4, 40)),
                                           p_long=float(rng.choice([0.0, 0.02, 0.2])), tail=int(rng.integers(0, 40)))
            assert oracle.decompress_safe(s, len(d)) == d
            streams.append(s); caps.append(len(d))
    _batch_decode_vs_oracle(ctx, oracle, streams, caps)


def test_decode_synthetic_streams_errors_and_capacity(ctx, oracle):
    """the same streams truncated, with short / oversized outputs and flipped bytes: status must be the oracle's"""
    rng = np.random.default_rng(77)
    streams, caps = [], []
    for _ in range(150):
        s, d = corpus.synth_lz4_stream(rng, int(rng.integers(1, 200)), near=int(rng.integers(1, 100)), p_near=0.6,
                                       p_long=0.05, tail=int(rng.integers(0, 20)))
        kind = int(rng.integers(0, 5))
        if kind == 0:
            s = s[:int(rng.integers(1, len(s) + 1))]
            cap = len(d)
        elif kind == 1:
            cap = int(rng.integers(0, len(d) + 1))
        elif kind == 2:
            cap = len(d) + int(rng.integers(1, 100))
        elif kind == 3:
            b = bytearray(s)
            for _ in range(int(rng.integers(1, 4))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            s = bytes(b); cap = len(d) + 50
        else:
            cap = len(d)
        streams.append(s); caps.append(cap)
    _batch_decode_vs_oracle(ctx, oracle, streams, caps)


def test_decode_synthetic_streams_with_dictionary(ctx, oracle):
    """records whose early matches reach into a shared dictionary (decompressSafeUsingDict, src/lz4.zig:960)"""
    rng = np.random.default_rng(5)
    dic = rng.integers(0, 256, size=5000, dtype=np.uint8).tobytes()
    streams, caps = [], []
    for _ in range(60):
        s, d = corpus.synth_lz4_stream(rng, int(rng.integers(1, 120)), near=int(rng.integers(1, 100)), p_near=0.5)
        # re-base: decode with the dictionary as history by rewriting nothing — offsets stay valid; then add
        # streams whose first offsets point before dst
        streams.append(s); caps.append(len(d))
        b = bytearray(s)
        ll = b[0] >> 4
        if ll < 15 and len(b) > ll + 3:
            off = int(rng.integers(ll + 1, ll + 4000))
            b[1 + ll] = off & 255; b[2 + ll] = off >> 8
            streams.append(bytes(b)); caps.append(len(d) + 10)
    _batch_decode_vs_oracle(ctx, oracle, streams, caps, dic=dic)


def test_dictionary_compression_round_trips_and_pays(z, oracle, ctx):
    """encode side of decompressSafeUsingDict (SURVEY §8f rank 3 / configs[4] record case).  There is no reference
    output to equal (the reference never matches into a dictionary): the bar is the reference's own
    dictionary *decoder* (oracle) and K2 reproducing the input, compressFast's bytes without a dictionary, and a
    better ratio on small records that share a vocabulary."""
    from zig_lz4_b200 import datagen
    vocab = datagen.generate(1 << 20, mode=0, seed=31).tobytes()
    dic = vocab[:65536]
    recs = vocab[200000:200000 + 64 * 4096]
    rng = np.random.default_rng(3)
    # single-block entry point, various dictionary sizes (incl. > 64 KiB: only the tail counts) and accelerations
    for d in (b"", dic[:7], dic[:100], dic[:4096], dic, vocab[:100000]):
        for accel in (1, 4):
            for blk in (recs[:4096], recs[:13], recs[:12], b"", recs[:70000], dic[-3000:] + recs[:500], bytes(5000)):
                c = z.lz4.compressFastUsingDict(blk, d, accel)
                assert oracle.decompress_safe(c, len(blk), dict=d) == blk
                assert z.lz4.decompressSafeUsingDict(c, len(blk), d) == blk
                if len(d) == 0:
                    assert c == oracle.compress_fast(blk, accel)
    # a block that IS a piece of the dictionary compresses to almost nothing
    c = z.lz4.compressFastUsingDict(dic[1000:5000], dic)
    assert len(c) < 64 and oracle.decompress_safe(c, 4000, dict=dic) == dic[1000:5000]
    # batch of 4 KiB records: ratio with the dictionary beats the dictionary-blind one
    nrec = 64
    offs = np.arange(nrec, dtype=np.uint64) * 4096
    lens = np.full(nrec, 4096, dtype=np.uint32)
    caps = np.full(nrec, int(z.lz4.compressBound(4096)), dtype=np.uint32)
    doffs = np.arange(nrec, dtype=np.uint64) * int(caps[0])
    dst, ol, st = ctx.compress_fast_dict_batch(recs, offs, lens, int(caps.sum()), doffs, caps, dic)
    assert (st == 0).all()
    dst0, ol0, st0 = ctx.compress_fast_batch(recs, offs, lens, int(caps.sum()), doffs, caps)
    assert int(ol.sum()) < 0.9 * int(ol0.sum()), (int(ol.sum()), int(ol0.sum()))   # measured: 0.82 (4096-entry single-probe table)
    blob = b"".join(dst[int(doffs[i]):int(doffs[i]) + int(ol[i])].tobytes() for i in range(nrec))
    coffs = np.concatenate([[0], np.cumsum(ol)[:-1]]).astype(np.uint64)
    out, ol2, st2 = ctx.decompress_safe_batch(blob, coffs, ol, nrec * 4096, offs, lens, dict=dic)
    assert (st2 == 0).all() and out[:nrec * 4096].tobytes() == recs
    for i in range(0, nrec, 9):
        ci = dst[int(doffs[i]):int(doffs[i]) + int(ol[i])].tobytes()
        assert oracle.decompress_safe(ci, 4096, dict=dic) == recs[i * 4096:(i + 1) * 4096]
    # capacity errors are reported per record
    tiny = np.full(nrec, 16, dtype=np.uint32)
    _, ol3, st3 = ctx.compress_fast_dict_batch(recs, offs, lens, int(caps.sum()), doffs, tiny, dic)
    assert (st3 == 1).all() and (ol3 == 0).all()
