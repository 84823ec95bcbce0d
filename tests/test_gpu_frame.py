"""GPU parity tests for the frame path (K1/K2 + K4 xxh32 + K5 scan + K6 assembly + K7 walk) through the
C-ABI: frames must be byte-identical to the oracle's restatement of lz4f.compressFrame
(reference src/lz4f.zig:354-446), decodable by the stock decoder, and decompressFrame must reproduce
decompressFrame's results and error kinds (src/lz4f.zig:541-638)."""
import ctypes as C

import numpy as np
import pytest

import corpus

pytestmark = pytest.mark.gpu

PREFS = [
    dict(),
    dict(block_mode=1),
    dict(content_checksum=1),
    dict(block_checksum=1),
    dict(block_mode=1, block_checksum=1, content_checksum=1, content_size=-1),
    dict(block_size_id=5, block_mode=1), dict(block_size_id=6), dict(block_size_id=7, content_checksum=1),
    dict(dict_id=0xCAFE),
]


def both_prefs(z, oracle, kw, n):
    kw = dict(kw)
    if kw.get("content_size") == -1:
        kw["content_size"] = n
    zp = z.lz4f.Preferences(blockSizeID=kw.get("block_size_id", 0), blockMode=kw.get("block_mode", 0),
                            contentChecksumFlag=kw.get("content_checksum", 0), contentSize=kw.get("content_size", 0),
                            dictID=kw.get("dict_id", 0), blockChecksumFlag=kw.get("block_checksum", 0),
                            compressionLevel=kw.get("compression_level", 0))
    return zp, oracle.make_prefs(**kw)


def stock_frame_decode(frame, n):
    import pyarrow as pa
    return pa.decompress(frame, decompressed_size=n, codec="lz4").to_pybytes()


@pytest.mark.parametrize("kw", PREFS)
@pytest.mark.parametrize("name,data", corpus.compat_cases() + [("hello10", b"Hello, World! This is a test of LZ4 frame compression. " * 10)])
def test_frame_bytes_equal_oracle(z, oracle, kw, name, data):
    zp, op = both_prefs(z, oracle, kw, len(data))
    want = oracle.compress_frame(data, op)
    got = z.lz4f.compressFrame(data, zp)
    assert got == want, (name, kw, len(got), len(want))
    assert z.lz4f.decompressFrame(got, len(data) + 64) == data
    assert z.lz4f.decompressFrame(got, len(data)) == data
    if not kw.get("dict_id"):
        assert stock_frame_decode(got, len(data)) == data           # test_compat.zig G1 stand-in


def test_default_prefs_none(z, oracle):
    data = corpus.LOREM * 7
    assert z.lz4f.compressFrame(data) == oracle.compress_frame(data)
    assert z.lz4f.compressFrame(b"") == bytes.fromhex("04224d184040c000000000")


def test_multi_block_frames(z, oracle):
    for data in (corpus.multi_block_1mib(), corpus.stream_1mib()):
        for kw in (dict(), dict(block_mode=1, content_checksum=1, block_checksum=1), dict(block_size_id=5, block_checksum=1)):
            zp, op = both_prefs(z, oracle, kw, len(data))
            f = z.lz4f.compressFrame(data, zp)
            assert f == oracle.compress_frame(data, op, threads=8)
            assert z.lz4f.decompressFrame(f, len(data)) == data
            assert stock_frame_decode(f, len(data)) == data


@pytest.mark.parametrize("bsid,mib", [(4, 48), (7, 64), (6, 33)])
def test_synthetic_mixed_frame(z, oracle, ctx, bsid, mib):
    """mixed-entropy classes, raw-stored blocks included (class 3), ragged tail"""
    from zig_lz4_b200 import datagen
    n = (mib << 20) + 12345
    span = 65536 if bsid == 4 else (4 << 20)
    data = datagen.generate(n, mode=4, span=span)
    kw = dict(block_size_id=bsid, block_mode=1, block_checksum=1, content_checksum=1, content_size=-1)
    zp, op = both_prefs(z, oracle, kw, n)
    f = ctx.compress_frame(data, zp)
    want = oracle.compress_frame(data, op, threads=oracle.hardware_threads())
    assert len(f) == len(want)
    assert f == want
    out = ctx.decompress_frame(f, n)
    assert out == data.tobytes()
    assert oracle.decompress_frame(f, n, threads=oracle.hardware_threads()) == data.tobytes()


def test_frame_error_kinds(z, oracle):
    data = corpus.multi_block_1mib()
    zp, op = both_prefs(z, oracle, dict(block_mode=1, content_checksum=1, block_checksum=1), len(data))
    f = z.lz4f.compressFrame(data, zp)

    def same_error(frame, cap):
        try:
            want = (0, oracle.decompress_frame(frame, cap))
        except oracle.OracleError as e:
            want = (e.code, None)
        try:
            got = (0, z.lz4f.decompressFrame(frame, cap))
        except z.B2Error as e:
            got = (e.code, None)
        assert got[0] == want[0], (oracle.status_name(want[0]), z.status_name(got[0]))
        if want[0] == 0:
            assert got[1] == want[1]
        return want[0]

    bad = bytearray(f); bad[-1] ^= 0xFF
    assert same_error(bytes(bad), len(data)) == 117             # ContentChecksumInvalid (test_lz4f.zig:167-179)
    bad = bytearray(f); bad[40] ^= 0x01
    assert same_error(bytes(bad), len(data)) == 106             # BlockChecksumInvalid
    bad = bytearray(f); bad[len(f) // 2] ^= 0x55
    same_error(bytes(bad), len(data))
    assert same_error(f[:len(f) // 2], len(data)) == 113        # FrameSizeWrong
    assert same_error(f[:-4], len(data)) == 113                 # content checksum missing
    assert same_error(f[:-8], len(data)) in (0, 113)            # end mark missing (loop just ends), then checksum
    same_error(f, len(data) - 1)
    same_error(f, len(data) - 65536)
    same_error(f, 65536)
    same_error(f, 0)
    for hdr in (b"\x04\x22\x4d", b"\x00\x00\x00\x00\x40\x40\xc0", b"\x04\x22\x4d\x18\x80\x40\xc0", b"\x04\x22\x4d\x18\x42\x40\xc0",
                b"\x04\x22\x4d\x18\x40\x41\xc0", b"\x04\x22\x4d\x18\x40\x10\xc0", b"\x04\x22\x4d\x18\x40\x40\xc1", b""):
        same_error(hdr, 10)
    # no checksums: corrupt a token -> DecompressionFailed or silently different data, same as the oracle
    zp2, op2 = both_prefs(z, oracle, dict(), len(data))
    g = bytearray(z.lz4f.compressFrame(data, zp2))
    rng = np.random.default_rng(3)
    for _ in range(12):
        h = bytearray(g)
        h[int(rng.integers(7, len(h)))] ^= int(rng.integers(1, 256))
        same_error(bytes(h), len(data) + 70000)
    with pytest.raises(z.B2Error) as e:
        z.lz4f.compressFrame(data, zp, dst_capacity=z.lz4f.compressFrameBound(len(data), zp) - 1)
    assert e.value.name == "lz4f.DstMaxSizeTooSmall"            # src/lz4f.zig:363-366


def test_foreign_frames(z, oracle):
    """frames from the stock encoder (G2, test_compat.zig:203-254), including a linked multi-block frame
    that the reference decoder rejects and short non-final blocks (general layout path)."""
    import pyarrow as pa
    for name, data in corpus.compat_cases():
        f = pa.compress(data, codec="lz4", asbytes=True)
        try:
            want = (0, oracle.decompress_frame(f, len(data) + 16))
        except oracle.OracleError as e:
            want = (e.code, None)
        try:
            got = (0, z.lz4f.decompressFrame(f, len(data) + 16))
        except z.B2Error as e:
            got = (e.code, None)
        assert got == want, name
    # hand-assembled frame with short, empty-ish and raw blocks in the middle
    parts = [b"A" * 1000, b"", b"xyz" * 11, bytes(range(256)) * 300, b"tail"]
    frame = bytearray(bytes.fromhex("04224d186040" + "82"))
    for i, p in enumerate(parts):
        if not p:
            continue
        c = oracle.compress_fast(p)
        if i % 2 == 0 and len(c) < len(p):
            frame += len(c).to_bytes(4, "little") + c
        else:
            frame += (len(p) | 0x80000000).to_bytes(4, "little") + p
    frame += b"\0\0\0\0"
    total = b"".join(parts)
    assert oracle.decompress_frame(bytes(frame), len(total) + 5) == total
    assert z.lz4f.decompressFrame(bytes(frame), len(total) + 5) == total
    assert z.lz4f.decompressFrame(bytes(frame), len(total)) == total
    for cap in (len(total) - 1, 1500, 1000, 999, 0):
        try:
            want = (0, oracle.decompress_frame(bytes(frame), cap))
        except oracle.OracleError as e:
            want = (e.code, None)
        try:
            got = (0, z.lz4f.decompressFrame(bytes(frame), cap))
        except z.B2Error as e:
            got = (e.code, None)
        assert got == want, cap


def test_streaming_trio_equals_one_shot(z, oracle):
    """README.md:98-122 API: begin ++ update* ++ end == compressFrame for the same prefs."""
    from zig_lz4_b200 import datagen
    data = datagen.generate((3 << 20) + 777, mode=4).tobytes()
    rng = np.random.default_rng(5)
    for kw in (dict(), dict(block_mode=1, content_checksum=1, block_checksum=1), dict(block_size_id=5, content_checksum=1)):
        zp, op = both_prefs(z, oracle, kw, len(data))
        want = oracle.compress_frame(data, op, threads=8)
        for pattern in ("whole", "small", "ragged"):
            cctx = z.lz4f.createCompressionContext()
            out = z.lz4f.compressBegin(cctx, zp)
            pos = 0
            while pos < len(data):
                step = {"whole": len(data), "small": 10007, "ragged": int(rng.integers(1, 300000))}[pattern]
                out += z.lz4f.compressUpdate(cctx, data[pos:pos + step])
                pos += step
            out += z.lz4f.compressEnd(cctx)
            z.lz4f.freeCompressionContext(cctx)
            assert out == want, (kw, pattern)
    cctx = z.lz4f.createCompressionContext()
    with pytest.raises(z.B2Error) as e:
        z.lz4f.compressUpdate(cctx, b"abc")
    assert e.value.name == "lz4f.CompressionStateUninitialized"
    out = z.lz4f.compressBegin(cctx, None) + z.lz4f.compressEnd(cctx)
    assert out == bytes.fromhex("04224d184040c000000000")


def test_device_pointer_frame_and_shards(z, oracle, ctx):
    """the *_dev entry points on torch-owned device memory + the multi-GPU shard decomposition:
    header ++ body_0 ++ body_1 ++ endmark ++ checksum == one-shot frame."""
    import torch
    from zig_lz4_b200 import datagen
    n = (9 << 20) + 4321
    data = datagen.generate(n, mode=4)
    kw = dict(block_mode=1, block_checksum=1, content_checksum=1, content_size=-1)
    zp, op = both_prefs(z, oracle, kw, n)
    want = oracle.compress_frame(data, op, threads=8)
    src = torch.from_numpy(data).cuda()
    cap = z.lz4f.compressFrameBound(n, zp)
    dst = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    before = z.kernel_launch_count()
    sz = ctx.compress_frame_dev(src.data_ptr(), n, dst.data_ptr(), cap, zp, s)
    assert z.kernel_launch_count() > before
    assert dst[:sz].cpu().numpy().tobytes() == want
    back = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    m = ctx.decompress_frame_dev(dst.data_ptr(), sz, back.data_ptr(), n, s)
    assert m == n and torch.equal(back[:n], src)
    # shards: split at a block boundary
    cut = (n // 65536 // 2) * 65536
    b0 = torch.empty(cap, dtype=torch.uint8, device="cuda")
    b1 = torch.empty(cap, dtype=torch.uint8, device="cuda")
    s0 = ctx.compress_blocks_dev(src.data_ptr(), cut, b0.data_ptr(), cap, zp, s)
    s1 = ctx.compress_blocks_dev(src.data_ptr() + cut, n - cut, b1.data_ptr(), cap, zp, s)
    L = z.lib()
    from zig_lz4_b200._native import XxhState
    st = XxhState()
    L.b2lz4_xxh32_state_init(C.byref(st), 0)
    assert L.b2lz4_xxh32_state_update_dev(ctx.handle, C.byref(st), src.data_ptr(), cut, s) == 0       # rank 0
    assert L.b2lz4_xxh32_state_update_dev(ctx.handle, C.byref(st), src.data_ptr() + cut, n - cut, s) == 0  # rank 1
    csum = L.b2lz4_xxh32_state_final(C.byref(st))
    frame = (z.lz4f.writeFrameHeader(zp) + b0[:s0].cpu().numpy().tobytes() + b1[:s1].cpu().numpy().tobytes() + b"\0\0\0\0" +
             csum.to_bytes(4, "little"))
    assert frame == want
    # shard decode
    o0 = torch.empty(cut + 64, dtype=torch.uint8, device="cuda")
    assert ctx.decompress_blocks_dev(b0.data_ptr(), s0, o0.data_ptr(), cut, 65536, True, s) == cut
    assert torch.equal(o0[:cut], src[:cut])


def _same_frame_result(z, oracle, frame, cap):
    try:
        want = (0, oracle.decompress_frame(frame, cap))
    except oracle.OracleError as e:
        want = (e.code, None)
    try:
        got = (0, z.lz4f.decompressFrame(frame, cap))
    except z.B2Error as e:
        got = (e.code, None)
    assert got[0] == want[0], (oracle.status_name(want[0]), z.status_name(got[0]))
    if want[0] == 0:
        assert got[1] == want[1]
    return want[0]


def test_parallel_index_adversarial_payloads(z, oracle):
    """the parallel block index (k_index.cu) tests every byte position as a possible header: payloads made of
    plausible header words (small u32s, raw flags, zeros) must not derail it, and a stored block larger than
    the frame's block size (accepted by the reference, src/lz4f.zig:578-587) must take the serial fallback"""
    rng = np.random.default_rng(11)
    hdr64k = bytes.fromhex("04224d186040" + "82")              # independent blocks, 64 KiB, no checksums
    hdr_bc = bytes.fromhex("04224d187040") + b"\x00"            # + block checksums; header checksum fixed below
    hdr_bc = hdr_bc[:6] + bytes([(oracle.xxh32(hdr_bc[4:6]) >> 8) & 0xFF])

    def raw_block(payload, bc=False):
        rec = (len(payload) | 0x80000000).to_bytes(4, "little") + payload
        if bc:
            rec += oracle.xxh32(payload).to_bytes(4, "little")
        return rec

    small_words = rng.integers(1, 3000, size=15000, dtype=np.uint32).tobytes()            # every word looks like a header
    flagged = (rng.integers(0, 70000, size=15000, dtype=np.uint32) | 0x80000000).astype(np.uint32).tobytes()
    zeros = bytes(60000)
    mixed = b"".join([small_words[:20000], zeros[:5000], flagged[:20000], b"\x01\x00\x00\x00" * 3000])
    for bc, hdr in ((False, hdr64k), (True, hdr_bc)):
        payloads = [small_words[:60000], flagged[:60000], zeros, mixed[:65536], b"\x04\x00\x00\x00\x00\x00\x00\x00" * 4000]
        frame = bytearray(hdr)
        total = b""
        for pl in payloads * 3:
            c = oracle.compress_fast(pl)
            if len(c) < len(pl) and len(total) % 2 == 0:
                frame += len(c).to_bytes(4, "little") + c
                if bc:
                    frame += oracle.xxh32(c).to_bytes(4, "little")
            else:
                frame += raw_block(pl, bc)
            total += pl
        frame += b"\0\0\0\0"
        assert _same_frame_result(z, oracle, bytes(frame), len(total)) == 0
        _same_frame_result(z, oracle, bytes(frame[:-4]), len(total))          # no end mark: loop runs off the end
        _same_frame_result(z, oracle, bytes(frame[:-9]), len(total))          # truncated last record
        _same_frame_result(z, oracle, bytes(frame[:len(frame) // 2]), len(total))
    # a stored block above the block size: the candidate filter rejects it, the serial walk accepts it
    big = rng.integers(0, 256, size=70000, dtype=np.uint8).tobytes()
    frame = hdr64k + raw_block(b"abc" * 100) + raw_block(big) + raw_block(b"tail") + b"\0\0\0\0"
    assert _same_frame_result(z, oracle, frame, 80000) == 0
    # empty body, body that is only an end mark, garbage after the header
    for body in (b"", b"\0\0\0\0", b"\0\0\0", b"\xff\xff\xff\x7f", b"\x05\x00\x00\x80abc"):
        _same_frame_result(z, oracle, hdr64k + body, 100)


def test_parallel_index_equals_serial_walk(z, oracle, ctx):
    """same frame through the parallel index and (B2_SERIAL_WALK=1, separate process) the serial header chase"""
    import os
    import subprocess
    import sys
    from zig_lz4_b200 import datagen
    data = datagen.generate(24 << 20, mode=4).tobytes()
    zp, op = both_prefs(z, oracle, dict(block_mode=1, block_checksum=1, content_checksum=1), len(data))
    f = z.lz4f.compressFrame(data, zp)
    assert z.lz4f.decompressFrame(f, len(data)) == data
    code = ("import sys; sys.path.insert(0, %r); import zig_lz4_b200 as z; z.debug_tune('serial_walk', 1); "
            "f = open(sys.argv[1], 'rb').read(); "
            "d = z.lz4f.decompressFrame(f, int(sys.argv[2])); import hashlib; print(hashlib.sha1(d).hexdigest())")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import hashlib
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".lz4") as tf:
        tf.write(f); tf.flush()
        out = subprocess.run([sys.executable, "-c", code % root, tf.name, str(len(data))], capture_output=True,
                             text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == hashlib.sha1(data).hexdigest()


@pytest.fixture
def small_pipeline_chunks(z):
    """makes the host-pointer frame calls pipeline (upload | kernels | download on three streams) with 24-block
    chunks, so that frames of a few MiB go through the chunked path"""
    z.debug_tune("pipe_blocks", 24)
    yield
    z.debug_tune("pipe_blocks", 0)


@pytest.mark.parametrize("kw", PREFS + [dict(compression_level=9, block_checksum=1, content_checksum=1)])
def test_pipelined_host_frames(z, oracle, small_pipeline_chunks, kw):
    """chunked, double-buffered host path == one-shot oracle frame, for every preference combination"""
    from zig_lz4_b200 import datagen
    bsid = kw.get("block_size_id", 0)
    n = {0: 9, 5: 30, 6: 90, 7: 300}[bsid] * (1 << 20) + 12345
    if kw.get("compression_level"):
        n = 7 * (1 << 20) + 99
    data = datagen.generate(n, mode=4, seed=n).tobytes()
    zp, op = both_prefs(z, oracle, kw, len(data))
    want = oracle.compress_frame(data, op, threads=8)
    got = z.lz4f.compressFrame(data, zp)
    assert len(got) == len(want)
    assert got == want
    assert z.lz4f.decompressFrame(got, len(data)) == data
    assert z.lz4f.decompressFrame(got, len(data) + 777) == data


def test_pipelined_host_decode_falls_back_exactly(z, oracle, small_pipeline_chunks):
    """anything unusual inside a chunk (corruption, short output, short blocks) must give the oracle's result.  9 MiB = 144
    blocks against 24-block chunks: long enough for the decode call to start before the header chain is walked to its end
    (>= 4.4 chunks), so corruption found by the late part of the walk — after the first chunks are already in flight —
    is covered too."""
    from zig_lz4_b200 import datagen
    data = datagen.generate(9 << 20, mode=4, seed=5).tobytes()
    zp, op = both_prefs(z, oracle, dict(block_mode=1, block_checksum=1, content_checksum=1), len(data))
    f = z.lz4f.compressFrame(data, zp)
    rng = np.random.default_rng(8)
    for _ in range(6):
        bad = bytearray(f)
        bad[int(rng.integers(7, len(bad)))] ^= int(rng.integers(1, 256))
        _same_frame_result(z, oracle, bytes(bad), len(data))
    # block headers damaged in the last third of the frame (found by the late part of the header walk): a size beyond
    # the block size, a size running off the frame, a premature end mark
    pos, offs = 7, []              # (header: magic, FLG, BD, HC — no contentSize in these preferences)
    while True:
        h = int.from_bytes(f[pos:pos + 4], "little")
        if h == 0:
            break
        offs.append(pos)
        pos += 4 + (h & 0x7FFFFFFF) + 4
    for k, word in ((len(offs) - 3, 0x00020000), (len(offs) - 10, 0x7FFFFFF0), (len(offs) - 20, 0)):
        bad = bytearray(f)
        bad[offs[k]:offs[k] + 4] = word.to_bytes(4, "little")
        _same_frame_result(z, oracle, bytes(bad), len(data))
    for cap in (len(data) - 1, len(data) - 65536, 3 << 20, 65536 * 30 + 1):
        _same_frame_result(z, oracle, f, cap)
    _same_frame_result(z, oracle, f[:-4], len(data))
    _same_frame_result(z, oracle, f[:len(f) - 70000], len(data))
    # foreign layout: short blocks in the middle of a long frame
    zp2, op2 = both_prefs(z, oracle, dict(block_mode=1), len(data))
    parts = [data[i:i + 50000] for i in range(0, 4 << 20, 50000)]
    frame = bytearray(bytes.fromhex("04224d186040" + "82"))
    for pt in parts:
        cpt = oracle.compress_fast(pt)
        if len(cpt) < len(pt):
            frame += len(cpt).to_bytes(4, "little") + cpt
        else:
            frame += (len(pt) | 0x80000000).to_bytes(4, "little") + pt
    frame += b"\0\0\0\0"
    assert _same_frame_result(z, oracle, bytes(frame), 5 << 20) == 0


def test_calls_from_several_host_threads(z, oracle):
    """SURVEY §8b threading: every entry point is callable from several host threads at once — on the shared default
    context (calls serialise on its lock) and on one context per thread (calls overlap on the device)."""
    import threading
    from zig_lz4_b200 import datagen
    inputs = [datagen.generate(300000 + 7919 * i, mode=i % 4, seed=40 + i).tobytes() for i in range(6)]
    wants = [oracle.compress_frame(d, None) for d in inputs]
    errors = []

    def shared(i):
        try:
            for _ in range(3):
                f = z.lz4f.compressFrame(inputs[i])
                assert f == wants[i]
                assert z.lz4f.decompressFrame(f, len(inputs[i])) == inputs[i]
                b = inputs[i][:50000]
                assert z.lz4.decompressSafe(z.lz4.compressDefault(b), len(b)) == b
        except BaseException as e:          # noqa: surfaced below
            errors.append(("shared", i, repr(e)))

    def own(i):
        try:
            c = z.Context(0)
            for _ in range(3):
                f = c.compress_frame(inputs[i])
                assert f == wants[i]
                assert c.decompress_frame(f, cap=len(inputs[i])) == inputs[i]
            c.close()
        except BaseException as e:          # noqa
            errors.append(("own", i, repr(e)))

    threads = [threading.Thread(target=shared, args=(i,)) for i in range(6)] + [threading.Thread(target=own, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=120)
    assert not errors, errors
    assert not any(t.is_alive() for t in threads)
