"""Per data class kernel times (SURVEY.md 8d classes): compress + decompress of N bytes of text / binary / redundant /
random / mixed data in 64 KiB (or --block-id) blocks, device-resident, CUDA-event phase times from the library."""
import argparse, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import zig_lz4_b200 as z
from zig_lz4_b200 import datagen

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--block-id", type=int, default=4)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--classes", default="text,binary,redundant,random,mixed")
ap.add_argument("--tune", default="", help="comma-separated b2lz4_debug_tune knobs, e.g. k2_variant=1,k2_occ=8")
a = ap.parse_args()
for kv in filter(None, a.tune.split(",")):
    k, v = kv.split("=")
    z.debug_tune(k, int(v))
n = a.mib << 20
ctx = z.Context(0)
ctx.set_timing(True)
zp = z.lz4f.Preferences(blockSizeID=a.block_id, blockMode=z.lz4f.BlockMode.independent)
cap = z.lz4f.compressFrameBound(n, zp)
host = torch.empty(n, dtype=torch.uint8).pin_memory()
comp = torch.empty(cap + 64, dtype=torch.uint8, device="cuda")
back = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
for name, mode in (("text", 0), ("binary", 1), ("redundant", 2), ("random", 3), ("mixed", 4)):
    if name not in a.classes.split(","):
        continue
    datagen.fill_ptr(host.data_ptr(), n, mode=mode, span=65536)
    src = host.to("cuda")
    best_c = best_d = best_i = 1e9
    for _ in range(a.reps):
        cs = ctx.compress_frame_dev(src.data_ptr(), n, comp.data_ptr(), cap, zp, 0)
        pc = ctx.last_phase_ms()
        m = ctx.decompress_frame_dev(comp.data_ptr(), cs, back.data_ptr(), n, 0)
        pd = ctx.last_phase_ms()
        best_c = min(best_c, pc[0]); best_d = min(best_d, pd[0]); best_i = min(best_i, pd[2])
    assert m == n and torch.equal(back[:n], src)
    print(json.dumps({"tune": a.tune, "class": name, "mib": a.mib, "ratio": round(n / cs, 3), "k1_ms": round(best_c, 3),
                      "k1_gbs": round(n / best_c / 1e6, 1), "k2_ms": round(best_d, 3), "k2_gbs": round(n / best_d / 1e6, 1),
                      "index_ms": round(best_i, 3)}), flush=True)
