"""Small fixed workload for ncu captures: one compress + one decompress of --mib MiB of one data class (device-resident)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import zig_lz4_b200 as z
from zig_lz4_b200 import datagen

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=256)
ap.add_argument("--mode", type=int, default=0)          # 0 text 1 binary 2 redundant 3 random 4 mixed
ap.add_argument("--block-id", type=int, default=4)
ap.add_argument("--level", type=int, default=0)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--tune", default="")
a = ap.parse_args()
for kv in filter(None, a.tune.split(",")):
    k, v = kv.split("=")
    z.debug_tune(k, int(v))
n = a.mib << 20
ctx = z.Context(0)
zp = z.lz4f.Preferences(blockSizeID=a.block_id, blockMode=1, compressionLevel=a.level)
cap = z.lz4f.compressFrameBound(n, zp)
host = torch.empty(n, dtype=torch.uint8).pin_memory()
datagen.fill_ptr(host.data_ptr(), n, mode=a.mode, span=65536)
src = host.to("cuda")
comp = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
back = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
for _ in range(a.reps):
    cs = ctx.compress_frame_dev(src.data_ptr(), n, comp.data_ptr(), cap, zp, 0)
    m = ctx.decompress_frame_dev(comp.data_ptr(), cs, back.data_ptr(), n, 0)
assert m == n and torch.equal(back[:n], src)
print("ok", n, cs)
