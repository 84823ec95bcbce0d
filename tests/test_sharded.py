"""Multi-GPU frame sharding (zig_lz4_b200/sharded.py): pure host logic, the world_size-2 gloo run of the
whole compress / decompress protocol on CPU (codec = the oracle engine of tests/oracle_engine.py), and the product
engine on one GPU."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_block_ranges_partition_every_block_once():
    from zig_lz4_b200 import sharded
    for world in (1, 2, 3, 4, 8):
        for nblocks in (0, 1, 2, 7, 8, 9, 16384, 2560):
            ranges = [sharded.block_range(r, world, nblocks) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == nblocks
            for a, b in zip(ranges, ranges[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


def test_byte_ranges_and_layout():
    from zig_lz4_b200 import sharded
    n, bs = (5 << 16) + 1234, 65536
    r0, r1 = sharded.byte_range(0, 2, n, bs), sharded.byte_range(1, 2, n, bs)
    assert r0 == (0, 3 << 16) and r1 == (3 << 16, n)
    assert sharded.byte_range(0, 2, 0, bs) == (0, 0) and sharded.byte_range(1, 2, 100, bs) == (0, 100)
    lay = sharded.frame_layout(15, [100, 0, 250], True)
    assert lay.body_offsets == [15, 115, 115] and lay.end_mark_pos == 365 and lay.total == 373
    assert sharded.frame_layout(7, [10], False).total == 21


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_ranks_gloo_frame_equals_one_shot(tmp_path, oracle):
    """rank 0 + rank 1 over gloo: sharded compress == the oracle's one-shot frame, sharded decode == the input,
    corrupt frames raise the reference's error on both ranks"""
    port, out = _free_port(), str(tmp_path / "result")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_sharded_worker.py"), str(r), "2", str(port), out],
                              env=env) for r in range(2)]
    try:
        for p in procs:
            p.wait(timeout=240)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r in range(2):
        with open("%s.%d" % (out, r)) as f:
            assert f.read() == "ok"


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.gpu
@pytest.mark.skipif(_gpu_count() < 2, reason="needs two GPUs (the driver's multi-GPU box; bench.py --gpus N runs the same "
                                             "checks as a pre-flight)")
def test_two_ranks_nccl_cuda_engine(tmp_path, z, oracle):
    """two processes, one GPU each, NCCL, the product engine: gathered frame == the oracle's one-shot frame, sharded decode
    == each rank's input, a corrupted content checksum raises ContentChecksumInvalid on both ranks"""
    port, out = _free_port(), str(tmp_path / "result")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "_sharded_nccl_worker.py"), str(r), "2", str(port), out],
                              env=env) for r in range(2)]
    try:
        for p in procs:
            p.wait(timeout=600)
    finally:
        for p in procs:
            if p.poll() is None:
                p.kill()
    for r in range(2):
        with open("%s.%d" % (out, r)) as f:
            assert f.read() == "ok"


@pytest.mark.gpu
def test_cuda_engine_single_rank(z, oracle):
    """world_size 1 through the product engine: same frame as the one-shot call, decode through index + cuts"""
    import torch
    from zig_lz4_b200 import datagen, sharded
    eng = sharded.CudaEngine(0)
    n = (9 << 20) + 4321
    data = datagen.generate(n, mode=4)
    for bc, cc in ((1, 1), (0, 0)):
        prefs = z.lz4f.Preferences(blockSizeID=4, blockMode=1, blockChecksumFlag=bc, contentChecksumFlag=cc, contentSize=n)
        src = torch.from_numpy(data).cuda()
        frame, layout, body = sharded.compress_frame_sharded(eng, src, prefs)
        want = oracle.compress_frame(data, oracle.make_prefs(4, 1, cc, n, 0, bc, 0), threads=8)
        assert frame.cpu().numpy().tobytes() == want
        out, (lo, hi), total = sharded.decompress_frame_sharded(eng, frame)
        assert total == n and (lo, hi) == (0, (n + 65535) // 65536)
        assert torch.equal(out, src)
        idx = eng.index(frame)
        assert idx["nblocks"] == hi and idx["terminal"] == 0 and idx["block_size"] == 65536
        assert idx["end_pos"] == len(want) - (4 if cc else 0)
        if cc:
            bad = frame.clone()
            bad[-1] ^= 0x55
            with pytest.raises(z.B2Error) as e:
                sharded.decompress_frame_sharded(eng, bad)
            assert e.value.name == "lz4f.ContentChecksumInvalid"


@pytest.mark.gpu
@pytest.mark.parametrize("bsid,bc,cc", [(4, 1, 1), (4, 0, 0), (5, 1, 0)])
def test_single_process_multi_gpu_calls(z, oracle, bsid, bc, cc):
    """b2lz4f_compress_frame_mgpu / _decompress_frame_mgpu: same bytes, sizes and errors as the one-GPU calls, for any
    ngpus (clamped to the devices present: on a one-GPU box this runs the one-GPU path, on a multi-GPU box the sharded one)"""
    import torch
    from zig_lz4_b200 import datagen
    n = (24 << 20) + 777
    data = datagen.generate(n, mode=4, seed=9).tobytes()
    prefs = z.lz4f.Preferences(blockSizeID=bsid, blockMode=1, blockChecksumFlag=bc, contentChecksumFlag=cc, contentSize=n)
    want = oracle.compress_frame(data, oracle.make_prefs(bsid, 1, cc, n, 0, bc, 0), threads=8)
    for ngpus in (1, 2, 8):
        f = z.lz4f.compressFrameMultiGPU(data, prefs, ngpus=ngpus)
        assert f == want, (ngpus, len(f), len(want))
        assert z.lz4f.decompressFrameMultiGPU(f, n, ngpus=ngpus) == data
        assert z.lz4f.decompressFrameMultiGPU(f, n + 1000, ngpus=ngpus) == data
    ngpus = max(2, torch.cuda.device_count())
    # errors keep the one-GPU kinds
    with pytest.raises(z.B2Error) as e:
        z.lz4f.decompressFrameMultiGPU(want, n - 1, ngpus=ngpus)
    assert e.value.name == _one_gpu_error(z, want, n - 1)
    bad = bytearray(want)
    bad[len(bad) // 2] ^= 0x40
    assert _mgpu_result(z, bytes(bad), n, ngpus) == _one_gpu_result(z, bytes(bad), n)
    if cc:
        bad = bytearray(want)
        bad[-1] ^= 1
        with pytest.raises(z.B2Error) as e:
            z.lz4f.decompressFrameMultiGPU(bytes(bad), n, ngpus=ngpus)
        assert e.value.name == "lz4f.ContentChecksumInvalid"
    with pytest.raises(z.B2Error) as e:
        z.lz4f.compressFrameMultiGPU(data, prefs, ngpus=ngpus, dst=bytearray(len(want)))
    assert e.value.name == "lz4f.DstMaxSizeTooSmall"
    # short frames and empty input take the one-GPU path
    small = data[:100000]
    sp = z.lz4f.Preferences(blockSizeID=bsid, blockMode=1, blockChecksumFlag=bc, contentChecksumFlag=cc)
    assert z.lz4f.compressFrameMultiGPU(small, sp, ngpus=ngpus) == z.lz4f.compressFrame(small, sp)
    assert z.lz4f.compressFrameMultiGPU(b"", sp, ngpus=ngpus) == z.lz4f.compressFrame(b"", sp)


@pytest.mark.gpu
def test_single_process_multi_gpu_hc(z, oracle):
    """compressHC level 9 through the multi-GPU call: every device keeps its own chain tables"""
    from zig_lz4_b200 import datagen
    n = (6 << 20) + 11
    data = datagen.generate(n, mode=1, seed=2).tobytes()
    prefs = z.lz4f.Preferences(blockSizeID=4, blockMode=1, blockChecksumFlag=1, compressionLevel=9)
    want = oracle.compress_frame(data, oracle.make_prefs(4, 1, 0, 0, 0, 1, 9), threads=8)
    f = z.lz4f.compressFrameMultiGPU(data, prefs, ngpus=8)
    assert f == want
    assert z.lz4f.decompressFrameMultiGPU(f, n, ngpus=8) == data


def _one_gpu_result(z, frame, cap):
    try:
        return (0, z.lz4f.decompressFrame(frame, cap))
    except z.B2Error as e:
        return (e.code, None)


def _mgpu_result(z, frame, cap, ngpus):
    try:
        return (0, z.lz4f.decompressFrameMultiGPU(frame, cap, ngpus=ngpus))
    except z.B2Error as e:
        return (e.code, None)


def _one_gpu_error(z, frame, cap):
    try:
        z.lz4f.decompressFrame(frame, cap)
    except z.B2Error as e:
        return e.name
    return "ok"
