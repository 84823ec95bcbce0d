"""One rank of the world_size-2 NCCL test of zig_lz4_b200.sharded with the PRODUCT engine (CudaEngine, libb2lz4.so) —
spawned by tests/test_sharded.py::test_two_ranks_nccl_cuda_engine.  argv: rank world port outfile.
The oracle is used as the checker only (gathered frame == the oracle's one-shot frame)."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main(rank, world, port):
    import torch
    import torch.distributed as dist
    import b2oracle as o
    import zig_lz4_b200 as z
    from zig_lz4_b200 import datagen, sharded

    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world, device_id=dev)
    eng = sharded.CudaEngine(rank)
    before = z.kernel_launch_count()
    cases = [
        ((37 << 20) + 4321, dict(block_size_id=4, block_mode=1, content_checksum=1, block_checksum=1), True),
        ((24 << 20) + 5, dict(block_size_id=7, block_mode=1, content_checksum=1, block_checksum=1), True),
        ((9 << 20), dict(block_size_id=5, block_mode=1, content_checksum=0, block_checksum=0), False),
        (65536, dict(block_size_id=4, block_mode=1, content_checksum=1, block_checksum=0), False),     # one block: rank 0 idle
    ]
    for n, kw, csz in cases:
        data = datagen.generate(n, mode=4, seed=n + 3)
        bs = {4: 65536, 5: 262144, 7: 4 << 20}[kw["block_size_id"]]
        prefs = z.lz4f.Preferences(blockSizeID=kw["block_size_id"], blockMode=1, contentChecksumFlag=kw["content_checksum"],
                                   blockChecksumFlag=kw["block_checksum"], contentSize=n if csz else 0)
        lo, hi = sharded.byte_range(rank, world, n, bs)
        shard = torch.from_numpy(data[lo:hi].copy()).to(dev) if hi > lo else eng.empty(0)
        frame, layout, body = sharded.compress_frame_sharded(eng, shard, prefs, gather_to=0)
        want = o.compress_frame(data, o.make_prefs(kw["block_size_id"], 1, kw["content_checksum"], n if csz else 0, 0,
                                                   kw["block_checksum"], 0), threads=4)
        assert layout.total == len(want), (layout.total, len(want))
        if rank == 0:
            assert frame.cpu().numpy().tobytes() == want, "NCCL-gathered frame differs from the oracle's one-shot frame"
        else:
            assert frame is None
        part, (blo, bhi), total = sharded.decompress_frame_sharded(eng, frame, src=0)
        assert total == n and (blo, bhi) == sharded.block_range(rank, world, (n + bs - 1) // bs)
        assert torch.equal(part, shard), "sharded decode differs from this rank's input range"
        whole, _, _ = sharded.decompress_frame_sharded(eng, frame, src=0, gather_to=1)
        if rank == 1:
            assert whole.cpu().numpy().tobytes() == data.tobytes()
        if kw["content_checksum"]:
            bad = None
            if rank == 0:
                bad = frame.clone()
                bad[-1] ^= 0x55
            try:
                sharded.decompress_frame_sharded(eng, bad, src=0)
                raise AssertionError("corrupt content checksum accepted")
            except z.B2Error as e:
                assert e.code == sharded.ERR_CONTENT_CHECKSUM_INVALID, e.code
    assert z.kernel_launch_count() > before, "no kernel of libb2lz4.so was launched"
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    try:
        main(rank, world, port)
        msg = "ok"
    except BaseException:
        msg = traceback.format_exc()
    with open("%s.%d" % (out, rank), "w") as f:
        f.write(msg)
