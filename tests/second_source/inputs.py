"""Deterministic inputs for the second-source vectors (shared by make_vectors.py and tests/test_second_source.py)."""
import collections
import os
import random
import struct
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import corpus  # noqa: E402
import zig_lz4_b200  # noqa: E402,F401  (registers the package alias)
from zig_lz4_b200 import datagen  # noqa: E402

CLASSES = (("text", 0), ("binary", 1), ("redundant", 2), ("random", 3))


def _runs_200k():
    """long single-byte runs separated by noise, > 64 KiB: drives the HC pattern analysis (src/lz4hc.zig:626-678)
    with candidates on both sides of the 65536 chain-table wrap."""
    rnd = random.Random(20261018)
    b = bytearray()
    while len(b) < 200000:
        b += bytes([rnd.choice(b"AB\0\xff")]) * rnd.choice((3, 5, 9, 40, 300, 2000, 70000 if len(b) < 1000 else 17))
        b += bytes(rnd.getrandbits(8) for _ in range(rnd.choice((0, 1, 2, 7))))
    return bytes(b[:200000])


def block_inputs():
    d = collections.OrderedDict()
    for name, data in corpus.block_cases():
        d[name] = data
    for name, data in corpus.compat_cases():
        d["compat_" + name] = data
    for cname, mode in CLASSES:
        for n in (4096, 65536, 262144):
            d["%s_%d" % (cname, n)] = datagen.generate(n, mode=mode).tobytes()
    d["mixed_65537"] = datagen.generate(65537, mode=datagen.MIXED, span=4096).tobytes()
    d["mixed_65535"] = datagen.generate(65535, mode=datagen.MIXED, span=4096).tobytes()
    d["f8_hazard"] = corpus.f8_hazard_input()
    d["runs_200k"] = _runs_200k()
    d["mod16_256k"] = corpus.multi_block_1mib()[:262144]
    d["zeros_70000"] = bytes(70000)
    return d


def frame_inputs():
    d = collections.OrderedDict()
    d["mixed_300k"] = datagen.generate(300000, mode=datagen.MIXED, span=16384).tobytes()
    for name, data in corpus.compat_cases():
        d["compat_" + name] = data
    d["mod16_1mib"] = corpus.multi_block_1mib()
    d["f8_hazard"] = corpus.f8_hazard_input()
    d["text_600k"] = datagen.generate(600000, mode=datagen.TEXT).tobytes()
    return d


def hostile_blocks():
    """(name, stream, capacity, dictionary|None) — every exit of decompressGeneric, src/lz4.zig:111-248."""
    lit = lambda b: bytes([len(b) << 4]) + b  # noqa: E731
    seq = lambda l, off, mlc: bytes([(len(l) << 4) | mlc]) + l + struct.pack("<H", off)  # noqa: E731
    dic = bytes(range(200, 256)) * 4
    return [
        ("literals_only", lit(b"hello"), 5, None),
        ("cap_too_small_literals", lit(b"hello"), 4, None),
        ("cap_zero", lit(b"hello"), 0, None),
        ("literal_run_past_input", bytes([0x50]) + b"abc", 16, None),
        ("ll_ext_runs_off", bytes([0xF0, 255, 255]), 4096, None),
        ("ll_ext_ok", bytes([0xF0, 3]) + bytes(18), 18, None),
        ("offset_truncated", bytes([0x10]) + b"a" + b"\x01", 64, None),
        ("offset_zero", seq(b"abcd", 0, 0) + lit(b"12345"), 64, None),
        ("offset_before_start", seq(b"abcd", 5, 0) + lit(b"12345"), 64, None),
        ("offset_exactly_start", seq(b"abcd", 4, 0) + lit(b"12345"), 64, None),
        ("match_overflows_cap", seq(b"abcd", 4, 6) + lit(b"12345"), 13, None),
        ("match_fits_exactly", seq(b"abcd", 4, 6), 14, None),
        ("ml_ext_runs_off", seq(b"abcd", 1, 15) + bytes([255, 255]), 4096, None),
        ("ml_ext_ok_overlap", seq(b"ab", 1, 15) + bytes([255, 7]) + lit(b"12345"), 4096, None),
        ("ends_after_offset_len", seq(b"abcd", 2, 3), 64, None),
        ("token_then_nothing", bytes([0x00]), 64, None),
        ("dict_match_inside", seq(b"", 10, 2) + lit(b"12345"), 64, dic),
        ("dict_match_spans", seq(b"xy", 6, 9) + lit(b"12345"), 64, dic),
        ("dict_match_spans_overlap", seq(b"x", 3, 15) + bytes([40]) + lit(b"12345"), 128, dic),
        ("dict_offset_too_far", seq(b"xy", 2 + len(dic) + 1, 0) + lit(b"12345"), 64, dic),
        ("dict_offset_at_edge", seq(b"xy", 2 + len(dic), 0) + lit(b"12345"), 64, dic),
    ]


def hostile_frames(s2):
    """(name, frame bytes, capacity) — exits of parseFrameHeader / decompressFrame, src/lz4f.zig:483-638.
    Frames are built with the second source `s2` (tests/second_source/zlz4_second.py)."""
    data = datagen.generate(150000, mode=datagen.MIXED, span=8192).tobytes()
    P = s2.Prefs
    good = s2.compress_frame(data, P(block_mode=1, block_checksum=1, content_checksum=1, content_size=len(data)))
    plain = s2.compress_frame(data, P())
    rnd = datagen.generate(70000, mode=datagen.RANDOM).tobytes()
    rawf = s2.compress_frame(rnd, P(block_mode=1))

    def flip(b, i, x=0x01):
        b = bytearray(b)
        b[i] ^= x
        return bytes(b)

    out = [
        ("good", good, len(data)),
        ("good_cap_plus", good, len(data) + 100),
        ("cap_short_by_one", good, len(data) - 1),
        ("cap_zero", good, 0),
        ("plain_default_prefs", plain, len(data)),
        ("truncated_header_3", good[:3], 10),
        ("truncated_header_6", good[:6], 10),
        ("truncated_in_content_size", good[:10], 10),
        ("bad_magic", flip(good, 0), len(data)),
        ("skippable_magic", struct.pack("<II", 0x184D2A53, 4) + b"abcd", 16),
        ("bad_version", flip(good, 4, 0x80), len(data)),
        ("reserved_flg_bit", flip(good, 4, 0x02), len(data)),
        ("reserved_bd_bit", flip(good, 5, 0x01), len(data)),
        ("bd_size_invalid", flip(plain, 5, 0x70), len(data)),
        ("header_checksum", flip(good, 14), len(data)),
        ("block_payload_flip", flip(good, 40), len(data)),
        ("block_checksum_flip", flip(good, len(good) - 9), len(data)),
        ("content_checksum_flip", flip(good, len(good) - 1), len(data)),
        ("no_content_checksum_bytes", good[:-4], len(data)),
        ("truncated_mid_block", good[:5000], len(data)),
        ("truncated_block_header", good[:17], len(data)),
        ("no_end_mark", plain[:-4], len(data)),
        ("plain_payload_flip_offset0", flip(plain, 7 + 4 + 1 + 20), len(data)),
        ("raw_blocks", rawf, len(rnd)),
        ("raw_block_no_room", rawf, 65535),
        ("raw_second_block_no_room", rawf, 65536 + 10),
        ("empty_frame", s2.compress_frame(b"", P()), 0),
        ("empty_frame_content_checksum", s2.compress_frame(b"", P(content_checksum=1)), 8),
    ]
    return out
