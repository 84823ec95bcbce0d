"""Host-pointer round trip (configs[1]) split into its two calls, with the pipeline on/off and the raw PCIe copy rates."""
import sys, time, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import zig_lz4_b200 as z
from zig_lz4_b200 import datagen
n = 1<<30
ctx = z.Context(0)
zp = z.lz4f.Preferences(blockSizeID=z.lz4f.BlockSizeID.max64KB, blockMode=z.lz4f.BlockMode.independent)
host = torch.empty(n, dtype=torch.uint8).pin_memory()
datagen.fill_ptr(host.data_ptr(), n, mode=datagen.MIXED, span=65536)
cap = z.lz4f.compressFrameBound(n, zp)
hcomp = torch.empty(cap, dtype=torch.uint8).pin_memory()
hback = torch.empty(n, dtype=torch.uint8).pin_memory()
import ctypes as C
def arr(t, k): return np.frombuffer((C.c_uint8 * k).from_address(t.data_ptr()), dtype=np.uint8)
hs, hd, hb = arr(host, n), arr(hcomp, cap), arr(hback, n)
for label, knobs in (("pipelined", {}), ("blocks1024", {"pipe_blocks": 1024}), ("blocks512", {"pipe_blocks": 512}), ("blocks4096", {"pipe_blocks": 4096}), ("oneshot", {"no_pipeline": 1})):
    for k in ("pipe_blocks", "no_pipeline"): z.debug_tune(k, knobs.get(k, 0))
    for it in range(3):
        t0 = time.perf_counter(); cs = ctx.compress_frame(hs, zp, dst=hd); t1 = time.perf_counter()
        m = ctx.decompress_frame(hd[:cs], dst=hb); t2 = time.perf_counter()
    print(label, "compress %.1f ms  decompress %.1f ms  total %.1f ms -> %.1f GB/s" % ((t1-t0)*1e3, (t2-t1)*1e3, (t2-t0)*1e3, n/(t2-t0)/1e9), flush=True)
    assert m == n and (hb == hs).all()
# raw copy speeds
d = torch.empty(n, dtype=torch.uint8, device='cuda')
for _ in range(2):
    torch.cuda.synchronize(); t0=time.perf_counter(); d.copy_(host, non_blocking=True); torch.cuda.synchronize(); t1=time.perf_counter()
    hback.copy_(d, non_blocking=True); torch.cuda.synchronize(); t2=time.perf_counter()
print("H2D %.1f GB/s  D2H %.1f GB/s" % (n/(t1-t0)/1e9, n/(t2-t1)/1e9))
