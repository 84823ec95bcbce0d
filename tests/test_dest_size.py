"""compressDestSize (reference src/lz4.zig:551-616): the oracle's literal restatement on the CPU, and the device
bisection (one K1 launch per probe over all blocks) against it through the C-ABI.

Bar: the pair (consumed, compressed size) is the reference's for every (input, capacity); dst[0..size) is
compressDefault(src[0..consumed)).  The reference itself leaves its LAST probe in dst, which is that stream only when
the last probe fitted — `test_reference_dst_is_its_last_probe` pins that observation on the oracle."""
import numpy as np
import pytest

import corpus


def _inputs():
    from zig_lz4_b200 import datagen
    rng = np.random.default_rng(11)
    out = [("A200", b"AAAAAAAAAA" * 20),                                   # src/test_dictionary.zig:81
           ("short", b"Hello World!"),
           ("tiny", b"abc"),
           ("period8", b"ABCDEFGH" * 125),
           ("lowent", rng.integers(0, 4, size=5000, dtype=np.uint8).tobytes())]
    for mode in (0, 1, 2, 3):
        out.append(("class%d" % mode, datagen.generate(20000, mode=mode, seed=3 + mode).tobytes()))
    return out


def _caps(n, full):
    caps = {0, 1, 2, 5, 12, 13, 14, 17, 50, n // 3, n // 2, n - 1, n, n + 1, full - 1, full, full + 7}
    return sorted(c for c in caps if c >= 0)


@pytest.mark.parametrize("name,data", _inputs())
def test_oracle_dest_size_properties(oracle, name, data):
    full = oracle.compress_bound(len(data))
    for cap in _caps(len(data), full):
        used, size, dst = oracle.compress_dest_size(data, cap)
        assert used <= len(data) and size <= cap
        # the sizes describe compressDefault of the chosen prefix
        assert size == len(oracle.compress_fast(data[:used])) if used else size == 0
        if cap >= full:
            assert used == len(data)


def test_reference_own_case(oracle):
    """src/test_dictionary.zig:78-103: 200 x 'A' into 50 bytes -> consumes something, decodes back"""
    data = b"AAAAAAAAAA" * 20
    used, size, dst = oracle.compress_dest_size(data, 50)
    assert 0 < used <= len(data) and size <= 50
    assert oracle.decompress_safe(dst[:size], used) == data[:used]


def test_reference_dst_is_its_last_probe(oracle):
    """when the last probe did not fit, the reference's dst is not the stream of the prefix it reports"""
    from zig_lz4_b200 import datagen
    data = datagen.generate(20000, mode=0, seed=3).tobytes()
    diverged = 0
    for cap in (500, 1000, 3000, 7000):
        used, size, dst = oracle.compress_dest_size(data, cap)
        want = oracle.compress_fast(data[:used])
        assert len(want) == size
        diverged += dst[:size] != want
    assert diverged > 0


def test_src_size_argument_limits_the_input(oracle):
    data = b"ABCDEFGH" * 125
    used, size, _ = oracle.compress_dest_size(data, 4096, src_size=100)
    assert used == 100 and size == len(oracle.compress_fast(data[:100]))
    assert oracle.compress_dest_size(data, 10, src_size=0)[:2] == (0, 0)


# ------------------------------------------------------------------ device
@pytest.mark.gpu
@pytest.mark.parametrize("name,data", _inputs())
def test_dest_size_equals_oracle(z, oracle, name, data):
    full = oracle.compress_bound(len(data))
    for cap in _caps(len(data), full):
        used, size, _ = oracle.compress_dest_size(data, cap)
        got, got_used = z.lz4.compressDestSize(data, cap)
        assert (got_used, len(got)) == (used, size), (name, cap)
        assert got == (oracle.compress_fast(data[:used]) if used else b""), (name, cap)
        if used:
            assert z.lz4.decompressSafe(got, used) == data[:used]


@pytest.mark.gpu
def test_dest_size_src_size_and_empty(z, oracle):
    data = b"ABCDEFGH" * 125
    got, used = z.lz4.compressDestSize(data, 4096, src_size=100)
    assert used == 100 and got == oracle.compress_fast(data[:100])
    assert z.lz4.compressDestSize(data, 10, src_size=0) == (b"", 0)
    assert z.lz4.compressDestSize(b"", 10) == (b"", 0)


@pytest.mark.gpu
def test_dest_size_batch_mixed_blocks(ctx, oracle):
    """many blocks of different classes, lengths and capacities in one call; large blocks take the wide table"""
    from zig_lz4_b200 import datagen
    rng = np.random.default_rng(5)
    blocks, caps = [], []
    for i in range(96):
        n = int(rng.integers(0, 9000)) if i % 7 else int(rng.integers(60000, 140000))
        b = datagen.generate(n, mode=i % 4, seed=100 + i).tobytes() if n else b""
        blocks.append(b)
        full = oracle.compress_bound(n)
        caps.append(int([0, 7, n // 4, n // 2, n, full, max(0, full - 1)][i % 7]))
    src = b"".join(blocks)
    src_len = np.array([len(b) for b in blocks], dtype=np.uint32)
    src_off = np.concatenate([[0], np.cumsum(src_len[:-1], dtype=np.uint64)]).astype(np.uint64)
    dst_cap = np.array(caps, dtype=np.uint32)
    dst_off = np.concatenate([[0], np.cumsum(dst_cap[:-1] + 3, dtype=np.uint64)]).astype(np.uint64)
    total = int(dst_off[-1] + dst_cap[-1]) + 8
    dst, used, out_len, status = ctx.compress_dest_size_batch(src, src_off, src_len, total, dst_off, dst_cap)
    for i, b in enumerate(blocks):
        w_used, w_size, _ = oracle.compress_dest_size(b, caps[i])
        assert status[i] == 0
        assert (int(used[i]), int(out_len[i])) == (w_used, w_size), (i, len(b), caps[i])
        got = dst[int(dst_off[i]):int(dst_off[i]) + int(out_len[i])].tobytes()
        assert got == (oracle.compress_fast(b[:w_used]) if w_used else b""), i
    # nothing written outside the blocks' own output (3 guard bytes after every capacity stay zero)
    mask = np.ones(total, dtype=bool)
    for i in range(len(blocks)):
        mask[int(dst_off[i]):int(dst_off[i]) + int(out_len[i])] = False
    assert not dst[mask].any()
