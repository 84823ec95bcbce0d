"""BASELINE.json configs 3-5 on one GPU, device-resident, one JSON line each (sizes reduced with --scale).
config 3: lz4f frame, 4 MiB independent blocks, block + content checksums (reported with and without the serial content chain)
config 4: decompress-only per class (tools/class_probe.py prints these)
config 5: compressHC level 9 on 256 KiB blocks; 4 KiB records (batch API), ratio beside the oracle's on a sample"""
import argparse, os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np
import torch
import zig_lz4_b200 as z
from zig_lz4_b200 import datagen

ap = argparse.ArgumentParser()
ap.add_argument("--c3-mib", type=int, default=2048)
ap.add_argument("--c5-mib", type=int, default=256)
ap.add_argument("--skip", default="")
a = ap.parse_args()
ctx = z.Context(0)
ctx.set_timing(True)


def timed(fn, reps=2):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    return r, best


if "c3" not in a.skip:
    n = a.c3_mib << 20
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    datagen.fill_ptr(host.data_ptr(), n, mode=datagen.MIXED, span=4 << 20)
    src = host.to("cuda")
    for label, cc in (("codec+block checksums", 0), ("+content checksum", 1)):
        zp = z.lz4f.Preferences(blockSizeID=7, blockMode=1, blockChecksumFlag=1, contentChecksumFlag=cc, contentSize=n)
        cap = z.lz4f.compressFrameBound(n, zp)
        comp = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
        back = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
        cs, tc = timed(lambda: ctx.compress_frame_dev(src.data_ptr(), n, comp.data_ptr(), cap, zp, 0))
        pc = ctx.last_phase_ms()
        m, td = timed(lambda: ctx.decompress_frame_dev(comp.data_ptr(), cs, back.data_ptr(), n, 0))
        pd = ctx.last_phase_ms()
        assert m == n and torch.equal(back[:n], src)
        print(json.dumps({"config": 3, "what": label, "mib": a.c3_mib, "blocks": n >> 22, "ratio": round(n / cs, 3),
                          "compress_gbs": round(n / tc / 1e9, 2), "decompress_gbs": round(n / td / 1e9, 2),
                          "compress_phase_ms": [round(x, 2) for x in pc], "decompress_phase_ms": [round(x, 2) for x in pd]}), flush=True)

if "c5" not in a.skip:
    import b2oracle as o
    n = a.c5_mib << 20
    for cname, mode in (("text", 0), ("binary", 1)):
        host = torch.empty(n, dtype=torch.uint8).pin_memory()
        datagen.fill_ptr(host.data_ptr(), n, mode=mode, span=65536)
        src = host.to("cuda")
        zp = z.lz4f.Preferences(blockSizeID=5, blockMode=1, compressionLevel=9)
        cap = z.lz4f.compressFrameBound(n, zp)
        comp = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
        back = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
        cs, tc = timed(lambda: ctx.compress_frame_dev(src.data_ptr(), n, comp.data_ptr(), cap, zp, 0), reps=1)
        pc = ctx.last_phase_ms()
        m, td = timed(lambda: ctx.decompress_frame_dev(comp.data_ptr(), cs, back.data_ptr(), n, 0))
        assert m == n and torch.equal(back[:n], src)
        # oracle ratio on a 4 MiB sample (16 blocks), and byte equality of those blocks
        sample = host[:4 << 20].numpy().tobytes()
        osz = sum(len(o.compress_hc(sample[i:i + (256 << 10)], 9)) for i in range(0, len(sample), 256 << 10))
        fast = z.lz4f.Preferences(blockSizeID=5, blockMode=1)
        cs0 = ctx.compress_frame_dev(src.data_ptr(), n, comp.data_ptr(), cap, fast, 0)
        print(json.dumps({"config": 5, "what": "HC level 9, 256 KiB blocks, " + cname, "mib": a.c5_mib, "ratio_hc9": round(n / cs, 4),
                          "ratio_oracle_hc9_sample": round(len(sample) / osz, 4), "ratio_fast": round(n / cs0, 4),
                          "compress_gbs": round(n / tc / 1e9, 3), "k3_ms": round(pc[0], 1), "decompress_gbs": round(n / td / 1e9, 2)}), flush=True)
    # 4 KiB records through the batch API (host arrays; timing includes the copies)
    nrec = 1 << 16
    recs = datagen.generate(nrec * 4096, mode=0).tobytes()
    offs = np.arange(nrec, dtype=np.uint64) * 4096
    lens = np.full(nrec, 4096, dtype=np.uint32)
    caps = np.full(nrec, int(z.lz4.compressBound(4096)), dtype=np.uint32)
    doffs = np.arange(nrec, dtype=np.uint64) * int(caps[0])
    ctx.compress_hc_batch(recs, offs, lens, int(caps.sum()), doffs, caps, level=9)     # first call: workspace + pinned staging
    t0 = time.perf_counter()
    dst, ol, st = ctx.compress_hc_batch(recs, offs, lens, int(caps.sum()), doffs, caps, level=9)
    t1 = time.perf_counter()
    assert (st == 0).all()
    want = sum(len(o.compress_hc(recs[i * 4096:(i + 1) * 4096], 9)) for i in range(256))
    got = int(ol[:256].sum())
    print(json.dumps({"config": 5, "what": "HC level 9, 4 KiB records (batch API, host arrays incl. copies)", "records": nrec,
                      "ratio": round(nrec * 4096 / int(ol.sum()), 4), "first256_bytes_equal_oracle_total": got == want,
                      "gbs_incl_copies": round(nrec * 4096 / (t1 - t0) / 1e9, 3)}), flush=True)
