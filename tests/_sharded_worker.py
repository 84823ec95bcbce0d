"""One rank of the world_size-2 gloo test of zig_lz4_b200.sharded (spawned by tests/test_sharded.py).
argv: rank world port outfile.  Writes 'ok' to outfile.rank on success, the traceback otherwise."""
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def main(rank, world, port, out):
    import numpy as np
    import torch
    import torch.distributed as dist
    import b2oracle as o
    from oracle_engine import OracleEngine
    from zig_lz4_b200 import datagen, sharded
    from zig_lz4_b200._native import B2Error, Prefs

    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    eng = OracleEngine()
    cases = [
        ((5 << 16) + 1234, dict(block_size_id=4, block_mode=1, content_checksum=1, block_checksum=1, content_size=1)),
        ((3 << 18) + 77, dict(block_size_id=5, block_mode=1, content_checksum=0, block_checksum=0, content_size=0)),
        (65536, dict(block_size_id=4, block_mode=1, content_checksum=1, block_checksum=0, content_size=0)),   # one block: rank 0 idle
        (0, dict(block_size_id=4, block_mode=1, content_checksum=1, block_checksum=1, content_size=0)),       # empty input
    ]
    for n, kw in cases:
        data = datagen.generate(n, mode=4, seed=n + 1).tobytes() if n else b""
        prefs = Prefs(kw["block_size_id"], kw["block_mode"], kw["content_checksum"], 0, n if kw["content_size"] else 0, 0,
                      kw["block_checksum"], 0, 0, 0)
        bs = {4: 65536, 5: 262144}[kw["block_size_id"]]
        lo, hi = sharded.byte_range(rank, world, n, bs)
        shard = torch.frombuffer(bytearray(data[lo:hi]), dtype=torch.uint8) if hi > lo else eng.empty(0)
        frame, layout, body = sharded.compress_frame_sharded(eng, shard, prefs, gather_to=0)
        want = o.compress_frame(data, o.make_prefs(kw["block_size_id"], kw["block_mode"], kw["content_checksum"],
                                                   n if kw["content_size"] else 0, 0, kw["block_checksum"], 0))
        assert layout.total == len(want), (layout, len(want))
        if rank == 0:
            assert frame.numpy().tobytes() == want, "sharded frame differs from the one-shot frame"
        else:
            assert frame is None
        # sharded layout without a gather: this rank's body sits at its offset of the one-shot frame
        _, layout2, body2 = sharded.compress_frame_sharded(eng, shard, prefs, gather_to=None)
        assert body2.numpy().tobytes() == want[layout2.body_offsets[rank]:layout2.body_offsets[rank] + layout2.body_sizes[rank]]
        # decode: frame on rank 0, every rank decodes its block range; gathered on rank 1
        whole, (blo, bhi), total = sharded.decompress_frame_sharded(eng, frame, src=0, gather_to=1)
        assert total == n
        nblocks = (n + bs - 1) // bs
        assert (blo, bhi) == sharded.block_range(rank, world, nblocks)
        if rank == 1:
            assert whole.numpy().tobytes() == data
        part, _, _ = sharded.decompress_frame_sharded(eng, frame, src=0, gather_to=None)
        assert part.numpy().tobytes() == data[blo * bs:min(bhi * bs, n)]
        # errors surface on every rank with the reference's kind
        if n:
            bad = None
            if rank == 0:
                b = bytearray(want)
                if kw["content_checksum"]:
                    b[-1] ^= 0x55
                else:
                    b = b[:len(b) - 9]                                     # cut inside the last record
                bad = torch.frombuffer(b, dtype=torch.uint8)
            try:
                sharded.decompress_frame_sharded(eng, bad, src=0)
                raise AssertionError("corrupt frame accepted")
            except B2Error as e:
                assert e.code == (117 if kw["content_checksum"] else 113), e.code
    # the size exchange itself
    assert sharded.exchange_sizes(100 + rank, eng.device) == [100 + r for r in range(world)]
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    rank, world, port, out = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    try:
        main(rank, world, port, out)
        msg = "ok"
    except BaseException:
        msg = traceback.format_exc()
    with open("%s.%d" % (out, rank), "w") as f:
        f.write(msg)
