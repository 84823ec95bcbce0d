"""BASELINE.json configurations at (or near) their full sizes, through size-independent properties: the frame is
compared with the oracle's by length + SHA-1 (the oracle compresses on all host threads), the round trip by
equality on the device, and the stock decoder must accept a sample of the blocks."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _roundtrip(z, oracle, ctx, n, zp, op, mode_span):
    import torch
    from zig_lz4_b200 import datagen
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    datagen.fill_ptr(host.data_ptr(), n, mode=datagen.MIXED, span=mode_span)
    src = host.to("cuda")
    cap = z.lz4f.compressFrameBound(n, zp)
    comp = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
    back = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    cs = ctx.compress_frame_dev(src.data_ptr(), n, comp.data_ptr(), cap, zp, 0)
    m = ctx.decompress_frame_dev(comp.data_ptr(), cs, back.data_ptr(), n, 0)
    assert m == n and torch.equal(back[:n], src)
    got = comp[:cs].cpu().numpy()
    want = oracle.compress_frame(host.numpy(), op, threads=oracle.hardware_threads())
    assert len(want) == cs
    assert hashlib.sha1(got.tobytes()).hexdigest() == hashlib.sha1(bytes(want)).hexdigest()
    return host, got


def test_config1_full_size_frame_equals_oracle(z, oracle, ctx):
    """configs[1]: 1 GiB mixed-entropy, 64 KiB independent blocks, default fast mode"""
    n = 1 << 30
    zp = z.lz4f.Preferences(blockSizeID=4, blockMode=1)
    op = oracle.make_prefs(block_size_id=4, block_mode=1)
    host, frame = _roundtrip(z, oracle, ctx, n, zp, op, 65536)
    # the stock decoder accepts the frame's first blocks (a frame cut after 64 blocks + end mark)
    import pyarrow as pa
    pos, blocks = 7, 0
    while blocks < 64:
        sz = int.from_bytes(frame[pos:pos + 4].tobytes(), "little") & 0x7FFFFFFF
        pos += 4 + sz; blocks += 1
    part = frame[:pos].tobytes() + b"\0\0\0\0"
    assert pa.decompress(part, decompressed_size=64 * 65536, codec="lz4").to_pybytes() == host[:64 * 65536].numpy().tobytes()


def test_config3_shape_4mib_blocks_checksums(z, oracle, ctx):
    """configs[2] shape at 1 GiB: one frame, 4 MiB independent blocks, block + content checksums, contentSize"""
    n = 1 << 30
    zp = z.lz4f.Preferences(blockSizeID=7, blockMode=1, blockChecksumFlag=1, contentChecksumFlag=1, contentSize=n)
    op = oracle.make_prefs(block_size_id=7, block_mode=1, block_checksum=1, content_checksum=1, content_size=n)
    _roundtrip(z, oracle, ctx, n, zp, op, 4 << 20)


def test_host_pointer_pipeline_full_size(z, oracle, ctx):
    """the chunked three-stream host path on 768 MiB (6 chunks): frame equals the oracle's, round trip exact"""
    from zig_lz4_b200 import datagen
    n = 768 << 20
    data = datagen.generate(n, mode=datagen.MIXED, span=65536, seed=99)
    zp = z.lz4f.Preferences(blockSizeID=4, blockMode=1, contentChecksumFlag=0)
    op = oracle.make_prefs(block_size_id=4, block_mode=1)
    f = ctx.compress_frame(data, zp)
    want = oracle.compress_frame(data, op, threads=oracle.hardware_threads())
    assert len(f) == len(want) and hashlib.sha1(bytes(f)).hexdigest() == hashlib.sha1(bytes(want)).hexdigest()
    back = np.empty(n, dtype=np.uint8)
    m = ctx.decompress_frame(f, dst=back)
    assert m == n and (back == data).all()


def test_ordered_launch_ragged_tail_and_single_class(z, oracle, ctx):
    """The expensive-first launch order (estimate pass + explicit block sets + scatter of sizes and status words: frames of
    at least 8288 blocks of 64 KiB): a frame whose last block is 321 bytes, block checksums on, and a single-class (text)
    input where the estimate puts every block into one bucket — both equal the oracle's frames; the same frame with the
    ordering switched off (b2lz4_debug_tune spare3) has the same bytes."""
    import torch
    from zig_lz4_b200 import datagen
    n = 8300 * 65536 + 321
    zp = z.lz4f.Preferences(blockSizeID=4, blockMode=1, blockChecksumFlag=1)
    op = oracle.make_prefs(block_size_id=4, block_mode=1, block_checksum=1)
    host, frame = _roundtrip(z, oracle, ctx, n, zp, op, 65536)
    old = z.debug_tune("spare3", 1)
    try:
        src = host.to("cuda")
        cap = z.lz4f.compressFrameBound(n, zp)
        comp = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
        cs = ctx.compress_frame_dev(src.data_ptr(), n, comp.data_ptr(), cap, zp, 0)
        assert cs == len(frame) and np.array_equal(comp[:cs].cpu().numpy(), frame)
    finally:
        z.debug_tune("spare3", old)
    # one class only
    text = torch.empty(n, dtype=torch.uint8).pin_memory()
    datagen.fill_ptr(text.data_ptr(), n, mode=0, span=65536)
    src = text.to("cuda")
    cs = ctx.compress_frame_dev(src.data_ptr(), n, comp.data_ptr(), cap, zp, 0)
    want = oracle.compress_frame(text.numpy(), op, threads=oracle.hardware_threads())
    assert cs == len(want) and hashlib.sha1(comp[:cs].cpu().numpy().tobytes()).hexdigest() == hashlib.sha1(bytes(want)).hexdigest()
