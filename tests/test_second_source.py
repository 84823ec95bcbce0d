"""Parity pin: the reference holds no golden vectors and cannot be built here (pure
Zig, no toolchain), so the C oracle is pinned by a SECOND, independently written restatement of the Zig —
tests/second_source/zlz4_second.py (Python, written from the .zig text, not from oracle/*.c) — whose outputs over a
fixed corpus are committed as tests/golden/second_source_vectors.json (generator: tests/second_source/make_vectors.py).

CPU: oracle == committed vectors (fast accel 1/3/70, HC 3/6/9 incl. the F8 input and > 64 KiB pattern runs, 208
frames over every preference combination, 49 decoder exit cases) and oracle == the restatement run live on small
inputs.  GPU: the CUDA path through the C-ABI == the same vectors."""
import hashlib
import json
import os
import random
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "second_source"))

import inputs  # noqa: E402
import zlz4_second as s2  # noqa: E402

with open(os.path.join(HERE, "golden", "second_source_vectors.json")) as _f:
    V = json.load(_f)


@pytest.fixture(scope="module")
def blocks():
    return inputs.block_inputs()


@pytest.fixture(scope="module")
def frames():
    return inputs.frame_inputs()


def _check(e, comp):
    tag = "%s %s" % (e["name"], {k: e[k] for k in ("kind", "accel", "level", "prefs") if k in e})
    assert len(comp) == e["csize"], tag
    assert hashlib.sha1(comp).hexdigest() == e["sha1_out"], tag
    if "out_hex" in e:
        assert comp.hex() == e["out_hex"], tag


def _oprefs(oracle, kw):
    return oracle.make_prefs(**kw)


def _zprefs(z, kw):
    return z.lz4f.Preferences(blockSizeID=kw["block_size_id"], blockMode=kw["block_mode"],
                              contentChecksumFlag=kw["content_checksum"], blockChecksumFlag=kw["block_checksum"],
                              contentSize=kw["content_size"], dictID=kw["dict_id"], compressionLevel=kw["compression_level"])


def test_vector_inputs_are_reproducible(blocks, frames):
    for e in V["blocks"]:
        assert hashlib.sha1(blocks[e["name"]]).hexdigest() == e["sha1_in"], e["name"]
    for e in V["frames"]:
        assert hashlib.sha1(frames[e["name"]]).hexdigest() == e["sha1_in"], e["name"]


def test_oracle_blocks_equal_second_source(oracle, blocks):
    for e in V["blocks"]:
        data = blocks[e["name"]]
        if e["kind"] == "fast":
            _check(e, oracle.compress_fast(data, e["accel"]))
        else:
            _check(e, oracle.compress_hc(data, e["level"]))
    assert V["f8_guard_hits"] >= 1          # the F8 hazard input really reaches src/lz4hc.zig:636 with matchIndex == 0


def test_oracle_frames_equal_second_source(oracle, frames):
    for e in V["frames"]:
        data = frames[e["name"]]
        f = oracle.compress_frame(data, _oprefs(oracle, e["prefs"]))
        _check(e, f)
        assert oracle.decompress_frame(f, len(data) + 3) == data


def _oracle_error_name(oracle, exc):
    return oracle.status_name(exc.code).split(".")[-1]


def test_oracle_decoder_exits_equal_second_source(oracle):
    want = {(d["kind"], d["name"]): d for d in V["decode_errors"]}
    for name, blob, cap, dic in inputs.hostile_blocks():
        d = want[("block", name)]
        try:
            out = oracle.decompress_safe(blob, cap, dic)
            assert d["ok"] and len(out) == d["n_out"] and hashlib.sha1(out).hexdigest() == d["sha1_out"], name
        except oracle.OracleError as e:
            assert not d["ok"] and _oracle_error_name(oracle, e) == d["error"], (name, str(e), d)
    for name, blob, cap in inputs.hostile_frames(s2):
        d = want[("frame", name)]
        try:
            out = oracle.decompress_frame(blob, cap)
            assert d["ok"] and len(out) == d["n_out"] and hashlib.sha1(out).hexdigest() == d["sha1_out"], name
        except oracle.OracleError as e:
            assert not d["ok"] and _oracle_error_name(oracle, e) == d["error"], (name, str(e), d)


def test_second_source_live_vs_oracle_small(oracle):
    """The restatement itself, run here on small random structured inputs, against the oracle (all entry points)."""
    rnd = random.Random(7)
    for it in range(120):
        n = rnd.choice((0, 1, 5, 12, 13, 14, 40, 200, 1000, 3000))
        alpha = rnd.choice((2, 4, 16, 256))
        data = bytearray(rnd.randrange(alpha) for _ in range(n))
        for _ in range(rnd.randrange(0, 6)):                     # paste repeats: matches at assorted distances
            if n > 8:
                a, b, ln = rnd.randrange(n), rnd.randrange(n), rnd.randrange(4, 60)
                data[b:b + ln] = data[a:a + ln]
        data = bytes(data[:n])
        accel = rnd.choice((1, 1, 2, 9, 65, 70000))
        c = s2.compress_fast(data, accel)
        assert c == oracle.compress_fast(data, accel)
        assert s2.decompress_safe(c, n) == data
        level = rnd.choice((3, 4, 5, 6, 7, 8, 9))
        h = s2.compress_hc(data, level)
        assert h == oracle.compress_hc(data, level)
        assert oracle.decompress_safe(h, n) == data if n else h == b""
        assert s2.xxh32(data) == oracle.xxh32(data)
        kw = dict(block_size_id=rnd.choice((0, 4, 5)), block_mode=rnd.randrange(2), block_checksum=rnd.randrange(2),
                  content_checksum=rnd.randrange(2), content_size=rnd.choice((0, n)), dict_id=rnd.choice((0, 9)),
                  compression_level=rnd.choice((0, 0, 3, 9)))
        f = s2.compress_frame(data, s2.Prefs(**kw))
        assert f == oracle.compress_frame(data, _oprefs(oracle, kw))
        assert s2.decompress_frame(f, n) == data
        assert s2.frame_bound(n, s2.Prefs(**kw)) == oracle.compress_frame_bound(n, _oprefs(oracle, kw))
        assert s2.header_size(f) == oracle.header_size(f)
        # truncated / capacity-starved decodes: same exit on both sides
        for cap, blob in ((max(n - 1, 0), c), (n, c[:max(len(c) - 2, 0)]), (n, c[:len(c) // 2])):
            try:
                a = ("ok", s2.decompress_safe(blob, cap))
            except s2.Lz4Error as e:
                a = ("err", e.kind)
            try:
                b = ("ok", oracle.decompress_safe(blob, cap))
            except oracle.OracleError as e:
                b = ("err", _oracle_error_name(oracle, e))
            assert a == b, (it, cap, len(blob))
        cap = compress_cap = rnd.choice((len(c), max(len(c) - 1, 0), len(c) // 2))
        try:
            a = ("ok", s2.compress_fast(data, accel, cap=cap))
        except s2.Lz4Error as e:
            a = ("err", e.kind)
        try:
            b = ("ok", oracle.compress_fast(data, accel, cap=compress_cap))
        except oracle.OracleError as e:
            b = ("err", _oracle_error_name(oracle, e))
        if n:
            assert a == b, (it, "limited output", cap, len(c))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_cuda_blocks_equal_second_source(z, blocks):
    for e in V["blocks"]:
        data = blocks[e["name"]]
        if e["kind"] == "fast":
            c = z.lz4.compressFast(data, e["accel"])
        else:
            c = z.lz4hc.compressHC(data, e["level"])
        _check(e, c)
        if e.get("accel", 1) == 1 and e.get("level", 9) == 9:
            assert z.lz4.decompressSafe(c, e["n"]) == data, e["name"]


@pytest.mark.gpu
def test_cuda_frames_equal_second_source(z, frames):
    for e in V["frames"]:
        data = frames[e["name"]]
        f = z.lz4f.compressFrame(data, _zprefs(z, e["prefs"]))
        _check(e, f)
        assert z.lz4f.decompressFrame(f, e["n"] + 3) == data, e["name"]


@pytest.mark.gpu
def test_cuda_decoder_exits_equal_second_source(z):
    want = {(d["kind"], d["name"]): d for d in V["decode_errors"]}
    for name, blob, cap, dic in inputs.hostile_blocks():
        d = want[("block", name)]
        try:
            out = z.lz4.decompressSafe(blob, cap) if dic is None else z.lz4.decompressSafeUsingDict(blob, cap, dic)
            assert d["ok"] and len(out) == d["n_out"] and hashlib.sha1(out).hexdigest() == d["sha1_out"], name
        except z.B2Error as e:
            assert not d["ok"] and e.name.split(".")[-1] == d["error"], (name, e.name, d)
    for name, blob, cap in inputs.hostile_frames(s2):
        d = want[("frame", name)]
        try:
            out = z.lz4f.decompressFrame(blob, cap)
            assert d["ok"] and len(out) == d["n_out"] and hashlib.sha1(out).hexdigest() == d["sha1_out"], name
        except z.B2Error as e:
            assert not d["ok"] and e.name.split(".")[-1] == d["error"], (name, e.name, d)
