"""Host-pointer round trip through the single-process multi-GPU calls (b2lz4f_compress_frame_mgpu /
b2lz4f_decompress_frame_mgpu): configs[1] data, pinned host buffers, ngpus = 1, 2, 4, 8 (as many as the box has)."""
import argparse, ctypes as C, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import zig_lz4_b200 as z
from zig_lz4_b200 import datagen
ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--reps", type=int, default=4)
a = ap.parse_args()
n = a.mib << 20
zp = z.lz4f.Preferences(blockSizeID=4, blockMode=1)
host = torch.empty(n, dtype=torch.uint8).pin_memory()
datagen.fill_ptr(host.data_ptr(), n, mode=datagen.MIXED, span=65536)
cap = z.lz4f.compressFrameBound(n, zp)
hcomp = torch.empty(cap, dtype=torch.uint8).pin_memory()
hback = torch.empty(n, dtype=torch.uint8).pin_memory()
def arr(t, k): return np.frombuffer((C.c_uint8 * k).from_address(t.data_ptr()), dtype=np.uint8)
hs, hd, hb = arr(host, n), arr(hcomp, cap), arr(hback, n)
ref = None
for g in (1, 2, 4, 8):
    if g > torch.cuda.device_count():
        break
    bc = bd = 1e9
    for _ in range(a.reps):
        t0 = time.perf_counter(); cs = z.lz4f.compressFrameMultiGPU(hs, zp, ngpus=g, dst=hd); t1 = time.perf_counter()
        m = z.lz4f.decompressFrameMultiGPU(hd[:cs], ngpus=g, dst=hb); t2 = time.perf_counter()
        bc, bd = min(bc, t1 - t0), min(bd, t2 - t1)
    assert m == n and (hb == hs).all()
    frame = bytes(hd[:cs])
    if ref is None:
        ref = frame
    print(json.dumps({"ngpus": g, "mib": a.mib, "frame_equals_one_gpu_frame": frame == ref, "compress_ms": round(bc * 1e3, 2),
                      "decompress_ms": round(bd * 1e3, 2), "round_trip_gbs": round(n / (bc + bd) / 1e9, 2),
                      "compress_gbs": round(n / bc / 1e9, 2), "decompress_gbs": round(n / bd / 1e9, 2)}), flush=True)
