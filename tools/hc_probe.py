"""K3 (compressHC level 9) on N MiB of one data class, 256 KiB blocks, device-resident; prints kernel ms."""
import argparse, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import zig_lz4_b200 as z
from zig_lz4_b200 import datagen
ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=256)
ap.add_argument("--mode", type=int, default=0)
ap.add_argument("--level", type=int, default=9)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--tune", default="")
a = ap.parse_args()
for kv in filter(None, a.tune.split(",")):
    k, v = kv.split("=")
    z.debug_tune(k, int(v))
n = a.mib << 20
ctx = z.Context(0); ctx.set_timing(True)
zp = z.lz4f.Preferences(blockSizeID=5, blockMode=1, compressionLevel=a.level)
cap = z.lz4f.compressFrameBound(n, zp)
host = torch.empty(n, dtype=torch.uint8).pin_memory()
datagen.fill_ptr(host.data_ptr(), n, mode=a.mode, span=65536)
src = host.to("cuda")
comp = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
for _ in range(a.reps):
    cs = ctx.compress_frame_dev(src.data_ptr(), n, comp.data_ptr(), cap, zp, 0)
    ph = ctx.last_phase_ms()
print(json.dumps({"tune": a.tune, "mib": a.mib, "mode": a.mode, "level": a.level, "ratio": round(n / cs, 4), "k3_ms": round(ph[0], 1), "gbs": round(n / ph[0] / 1e6, 3)}))
