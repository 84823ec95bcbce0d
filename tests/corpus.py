"""Parity corpus: the inputs of the reference's own tests, restated, plus synthetic
classes.  Each entry cites the reference test it comes from."""
import random
import struct

import numpy as np


def xoshiro256pp_bytes(seed, n):
    """Zig std.Random.DefaultPrng (Xoshiro256++ seeded through SplitMix64) `bytes()` as best restated
    from memory of Zig std 0.15 — the reference case (src/test_compat.zig:32-36, seed 12345) only relies on
    the bytes being incompressible, so exact equality with Zig's stream is not required."""
    M = (1 << 64) - 1

    def sm(x):
        x = (x + 0x9E3779B97F4A7C15) & M
        z = x
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & M
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & M
        return x, z ^ (z >> 31)

    s = []
    x = seed
    for _ in range(4):
        x, v = sm(x)
        s.append(v)

    def rotl(v, k):
        return ((v << k) | (v >> (64 - k))) & M

    out = bytearray()
    while len(out) < n:
        r = (rotl((s[0] + s[3]) & M, 23) + s[0]) & M
        t = (s[1] << 17) & M
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45)
        out += struct.pack("<Q", r)
    return bytes(out[:n])


LOREM = (b"Lorem ipsum dolor sit amet, consectetur adipiscing elit.\n"
         b"Sed do eiusmod tempor incididunt ut labore et dolore magna aliqua.\n"
         b"Ut enim ad minim veniam, quis nostrud exercitation ullamco laboris.")          # test_compat.zig:12-16


def compat_cases():
    """src/test_compat.zig:25-56 — the "test-compat corpus" of BASELINE.json configs[0]."""
    return [
        ("small", b"Hello World!"),
        ("repeated", b"ABCDEFGH" * 125),
        ("text", LOREM),
        ("random", xoshiro256pp_bytes(12345, 256)),
        ("empty", b""),
        ("large", bytes(i % 256 for i in range(100000))),
    ]


def block_cases():
    """src/test.zig inputs (:19,:58,:127,:182,:208,:239-246) + lz4hc / lz4f test inputs."""
    rnd = random.Random(54321)
    cases = [
        ("AAAA", b"AAAA"),                                                                  # test.zig:19
        ("sentence", b"Hello, World! This is a test of the LZ4 compression algorithm."),   # test.zig:58
        ("A160", b"A" * 160),                                                               # test.zig:127
        ("empty", b""),                                                                     # test.zig:182
        ("ABC", b"ABC"),                                                                    # test.zig:208
        ("mod256_10000", bytes(i % 256 for i in range(10000))),                             # test.zig:239-246
        ("ABCDx500", b"ABCD" * 500),                                                        # test_lz4hc.zig:62-71
        ("random1000", bytes(rnd.getrandbits(8) for _ in range(1000))),                     # test_lz4hc.zig:123-153
        ("TestDatax", b"TestData" * 200),                                                   # test_lz4hc.zig:234
        ("pat1", b"A" * 1000), ("pat2", b"AB" * 1000), ("pat4", b"ABCD" * 1000),            # test_lz4hc.zig:280-282
        ("frame_hello", b"Hello, World! This is a test of LZ4 frame compression. " * 10),   # lz4f.zig:678
        ("A1000", b"A" * 1000),                                                             # test_lz4f.zig:219
        ("Hello100", b"Hello " * 100),                                                      # test_lz4f.zig:260
    ]
    for size in (1, 10, 100, 1000, 10000):                                                  # test_lz4hc.zig:191-227
        half = size // 2
        cases.append(("halfX_%d" % size, b"X" * half + bytes(rnd.getrandbits(8) for _ in range(size - half))))
    for n in range(0, 40):                                                                  # boundary sizes around MFLIMIT
        cases.append(("tiny%d" % n, bytes((i * 7) % 5 + 65 for i in range(n))))
    return cases


def multi_block_1mib():
    """src/test_lz4f.zig:94-131 — 1 MiB of (i/16)%256"""
    return bytes((i // 16) % 256 for i in range(1 << 20))


def stream_1mib():
    """src/test_lz4hc_stream.zig:352-405 — 1 MiB of (i/256)%256"""
    return bytes((i // 256) % 256 for i in range(1 << 20))


def f8_hazard_input():
    """128 KiB random + "XXXXX"@10000,@70000 + 9 x 'Y' @65533 — drives the reference's u32
    underflow at src/lz4hc.zig:636 (guarded in oracle and kernel)."""
    rnd = random.Random(8)
    b = bytearray(rnd.getrandbits(8) for _ in range(128 * 1024))
    b[10000:10005] = b"XXXXX"
    b[70000:70005] = b"XXXXX"
    b[65533:65542] = b"Y" * 9
    return bytes(b)


def synth_lz4_stream(rng, nseq, ll_max=20, ml_max=20, near=64, p_near=0.5, p_long=0.02, tail=5):
    """A hand-assembled valid LZ4 block with `nseq` sequences and a literal tail: random literal-run and
    match lengths, offsets biased to the last `near` bytes (dense dependencies between neighbouring
    sequences, overlapping matches with offset < length) and a few long runs.  Returns (stream, decoded)."""
    out = bytearray()
    s = bytearray()

    def put_len(v):
        while v >= 255:
            s.append(255); v -= 255
        s.append(v)

    for _ in range(nseq):
        ll = int(rng.integers(0, ll_max + 1))
        ml = int(rng.integers(4, ml_max + 1))
        if rng.random() < p_long:
            ll = int(rng.integers(15, 700))
        if rng.random() < p_long:
            ml = int(rng.integers(19, 3000))
        if not out and ll == 0:
            ll = 1
        lits = rng.integers(0, 256, size=ll, dtype=np.uint8).tobytes()
        pos = len(out) + ll
        if rng.random() < p_near:
            off = int(rng.integers(1, min(near, pos) + 1))
        else:
            off = int(rng.integers(1, min(65535, pos) + 1))
        s.append((min(ll, 15) << 4) | min(ml - 4, 15))
        if ll >= 15:
            put_len(ll - 15)
        s += lits
        s += bytes([off & 255, off >> 8])
        if ml - 4 >= 15:
            put_len(ml - 4 - 15)
        out += lits
        for i in range(ml):
            out.append(out[pos - off + i])
    lits = rng.integers(0, 256, size=tail, dtype=np.uint8).tobytes()
    s.append(min(tail, 15) << 4)
    if tail >= 15:
        put_len(tail - 15)
    s += lits
    out += lits
    return bytes(s), bytes(out)
