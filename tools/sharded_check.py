"""Multi-GPU check of zig_lz4_b200/sharded.py over NCCL (run under torchrun, one rank per GPU):
sharded compress (gathered on rank 0) == the one-shot frame of the same bytes; sharded decode == the input; a corrupted
content checksum raises on every rank.  Prints one JSON line per case on rank 0.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/sharded_check.py"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
import zig_lz4_b200 as z
from zig_lz4_b200 import datagen, sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
eng = sharded.CudaEngine(local)
for mib, bsid, bc, cc in ((256, 4, 1, 1), (1024, 4, 0, 0), (1024, 7, 1, 0), (3, 4, 1, 1)):
    n = (mib << 20) + 12345
    bs = z.lz4f.BlockSizeID.toBlockSize(bsid)
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    datagen.fill_ptr(host.data_ptr(), n, mode=datagen.MIXED, span=bs)          # every rank generates the same bytes
    lo, hi = sharded.byte_range(rank, world, n, bs)
    shard = host[lo:hi].to(dev)
    prefs = z.lz4f.Preferences(blockSizeID=bsid, blockMode=1, blockChecksumFlag=bc, contentChecksumFlag=cc, contentSize=n)
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    frame, layout, body = sharded.compress_frame_sharded(eng, shard, prefs, gather_to=0)
    torch.cuda.synchronize(); dist.barrier(); t1 = time.perf_counter()
    same = None
    if rank == 0:
        whole = host.to(dev)
        cap = z.lz4f.compressFrameBound(n, prefs)
        one = torch.empty(cap + 64, dtype=torch.uint8, device=dev)
        sz = eng.ctx.compress_frame_dev(whole.data_ptr(), n, one.data_ptr(), cap, prefs, torch.cuda.current_stream().cuda_stream)
        same = sz == frame.numel() and torch.equal(one[:sz], frame)
        del whole, one
    torch.cuda.synchronize(); dist.barrier(); t2 = time.perf_counter()
    out, (blo, bhi), total = sharded.decompress_frame_sharded(eng, frame, src=0, gather_to=None)
    torch.cuda.synchronize(); dist.barrier(); t3 = time.perf_counter()
    ok_dec = total == n and torch.equal(out, shard)
    flags = torch.tensor([1 if ok_dec else 0], dtype=torch.int64, device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    raised = None
    if cc:
        bad = None
        if rank == 0:
            bad = frame.clone(); bad[-1] ^= 0x55
        try:
            sharded.decompress_frame_sharded(eng, bad, src=0)
            raised = "nothing"
        except z.B2Error as e:
            raised = e.name
    if rank == 0:
        print(json.dumps({"world": world, "mib": mib, "block_size": bs, "block_checksum": bc, "content_checksum": cc,
                          "frame_equals_one_shot": bool(same), "decode_equals_input_all_ranks": bool(flags.item()),
                          "body_sizes": layout.body_sizes, "compress_gather_ms": round((t1 - t0) * 1e3, 2),
                          "decode_ms": round((t3 - t2) * 1e3, 2), "corrupt_checksum_raises": raised}), flush=True)
dist.barrier()
dist.destroy_process_group()
