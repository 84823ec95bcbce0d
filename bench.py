#!/usr/bin/env python
"""bench.py — the hot path's measurements.

Headline (the JSON line's metric / value / e2e / roofline): BASELINE.json configs[1] — 1 GiB of synthetic mixed-entropy
data per GPU (four classes rotating every 64 KiB), one lz4f frame with 64 KiB independent
blocks, default fast mode.  One step = compress the buffer into a frame and decompress that frame back.  `value` =
uncompressed bytes / (t_compress + t_decompress) with everything resident in HBM, CUDA-event timed on the launching
stream, max over ranks.  `e2e` = the same round trip through the host-pointer C-ABI (pinned host buffers, H2D + D2H
inside the timed region; a pageable-buffer figure — what a Zig caller's slices are — is reported next to it).  N > 1:
weak scaling, one rank per GPU, 1 GiB per rank, the only exchange is an all-gather of the per-rank body sizes.

Extra sub-records of the same line (each with its own roofline fraction and CPU baseline; `--no-extra` skips them):
  config3  BASELINE configs[2]: ONE 10 GiB frame, 4 MiB independent blocks, block + content checksums, strong-scaled over
           the ranks (frame gathered to rank 0, index-cut decode, checksum state hand-off inside the timed region),
           timed without and with the serial content checksum
  config4  BASELINE configs[3]: decompress-only, per data class, 64 KiB and 4 MiB blocks, index build timed separately
  config5  BASELINE configs[4]: compressHC level 9 on 256 KiB blocks and on 4 KiB records (+ the shared-dictionary
           fast-mode ratio), GPU and CPU GB/s and both ratios

--impl reference times the CPU oracle (C restatement of the Zig reference; no zig toolchain exists in this image),
rebuilt -O3 -march=native on the box, on all host threads, for the headline metric/config.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "LZ4 compress & decompress GB/s (uncompressed)"
GIB = 1 << 30
SEED = 0x4C5A3442


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bytes", type=int, default=GIB, help="uncompressed bytes per GPU of the headline workload (default 1 GiB)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the config3/4/5 sub-records")
    ap.add_argument("--config3-gib", type=int, default=10)
    ap.add_argument("--config4-gib", type=int, default=4, help="content per class (BASELINE names 16; default sized for the step budget)")
    ap.add_argument("--config5-gib", type=int, default=1)
    return ap.parse_args()


def headline_config(nbytes):
    """The `config` object — byte-identical in the GPU arm and the reference arm."""
    return {"workload": "configs[1]: synthetic mixed-entropy (text/binary/redundant/random rotating every 64 KiB), one lz4f "
                        "frame, 64 KiB independent blocks, default fast mode, compress + decompress",
            "bytes_per_gpu": int(nbytes), "block_size": 65536}


def library_version():
    import zig_lz4_b200 as z
    return z.lib().b2lz4_version().decode()


def measured_traffic(nbytes):
    """DRAM bytes per launch of the two codec kernels from the committed ncu capture of this workload
    (profiles/traffic.json).  Refused (None + reason) unless it was captured from this very build of the kernels."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
    except Exception:
        return None, None, "profiles/traffic.json missing"
    if int(t.get("bytes_per_gpu", 0)) != int(nbytes):
        return None, None, "traffic.json is for another workload size"
    if t.get("b2lz4_version") != library_version():
        return None, None, "traffic.json was captured from build %r, this is %r" % (t.get("b2lz4_version"), library_version())
    return t["k_compress_fast"], t["k_decompress"], None


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU legs (oracle)
def oracle_frame_roundtrip(data, prefs_kw, threads, reps):
    """The CPU oracle on `threads` host threads over `data` (numpy uint8): best-of-`reps` compress and decompress."""
    import ctypes as C
    import numpy as np
    import b2oracle as o
    L = o.lib()
    n = data.nbytes
    p = o.make_prefs(**prefs_kw)
    cap = o.compress_frame_bound(n, p)
    dst = np.empty(cap, dtype=np.uint8)
    back = np.empty(n, dtype=np.uint8)
    out = C.c_size_t(0)
    tc, td, csize = [], [], 0
    for _ in range(reps):
        t0 = time.perf_counter()
        rc = L.b2o_compress_frame_mt(data.ctypes.data, n, dst.ctypes.data, cap, C.byref(p), C.byref(out), threads)
        t1 = time.perf_counter()
        assert rc == 0, rc
        csize = out.value
        rc = L.b2o_decompress_frame_mt(dst.ctypes.data, csize, back.ctypes.data, n, C.byref(out), threads)
        t2 = time.perf_counter()
        assert rc == 0 and out.value == n, (rc, out.value)
        tc.append(t1 - t0); td.append(t2 - t1)
    assert (back == data).all()
    return {"compress_gbs": n / min(tc) / 1e9, "decompress_gbs": n / min(td) / 1e9, "ratio": n / csize,
            "tc": tc, "td": td, "csize": csize}


def cpu_arm(nbytes, threads, steps, warmup):
    """Headline workload on the CPU oracle: compress + decompress, `steps` timed passes after `warmup`."""
    from zig_lz4_b200 import datagen
    data = datagen.generate(nbytes, mode=datagen.MIXED, span=65536)
    r = oracle_frame_roundtrip(data, dict(block_size_id=4, block_mode=1), threads, warmup + steps)
    tc, td = r["tc"][warmup:], r["td"][warmup:]
    best = min(a + b for a, b in zip(tc, td))
    return {"compress_gbs": nbytes / min(tc) / 1e9, "decompress_gbs": nbytes / min(td) / 1e9,
            "roundtrip_gbs": nbytes * len(tc) / (sum(tc) + sum(td)) / 1e9, "roundtrip_best_gbs": nbytes / best / 1e9,
            "ratio": r["ratio"], "ms_per_step": (sum(tc) + sum(td)) / len(tc) * 1e3}


def run_reference(args, rank, world):
    if rank != 0:
        return
    import b2oracle as o
    flags = o.prefer_native()
    threads = o.hardware_threads()
    sample = min(args.bytes, GIB)
    r = cpu_arm(sample, threads, args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": round(r["roundtrip_gbs"], 4), "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(r["ms_per_step"], 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": headline_config(args.bytes),
        "compress_gbs": round(r["compress_gbs"], 4), "decompress_gbs": round(r["decompress_gbs"], 4), "ratio": round(r["ratio"], 4),
        "best_step_gbs": round(r["roundtrip_best_gbs"], 4),
        "cpu_baseline": {"value": round(r["roundtrip_gbs"], 4), "unit": "GB/s", "cores": threads, "kind": "port",
                         "sample": "%d MiB of the same workload per step; C restatement of the Zig reference (oracle/, gcc %s "
                                   "built on this box), one block per task on all host threads, parallel frame assembly"
                                   % (sample >> 20, flags)},
        "e2e": {"value": round(r["roundtrip_gbs"], 4), "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm helpers
class Env:
    """what every sub-benchmark needs"""
    pass


def gen_to_device(E, nbytes, mode, span, unit0):
    """`nbytes` of the synthetic corpus starting at 64 KiB unit `unit0`, generated on the host in <= 1 GiB pieces
    (pinned staging) and copied to this rank's GPU."""
    import torch
    from zig_lz4_b200 import datagen
    out = torch.empty(nbytes + 64, dtype=torch.uint8, device=E.dev)
    piece = min(nbytes, GIB)
    if E.stage is None or E.stage.numel() < piece:
        E.stage = torch.empty(max(piece, 1), dtype=torch.uint8).pin_memory()
    pos = 0
    while pos < nbytes:
        k = min(piece, nbytes - pos)
        # the class of a unit depends on (byte position / span), so pieces must start on a multiple of 4 spans
        datagen.fill_ptr(E.stage.data_ptr(), k, seed=SEED + unit0 + pos // 65536, mode=mode, span=span)
        out[pos:pos + k].copy_(E.stage[:k], non_blocking=False)
        pos += k
    return out


def max_over_ranks(E, vals):
    import torch
    import torch.distributed as dist
    if E.world == 1:
        return [float(v) for v in vals]
    t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device=E.dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def sum_over_ranks(E, vals):
    import torch
    import torch.distributed as dist
    if E.world == 1:
        return [float(v) for v in vals]
    t = torch.tensor([float(v) for v in vals], dtype=torch.float64, device=E.dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()]


def timed(E, fn):
    """fn() bracketed by barrier + synchronize and CUDA events on the launching stream; returns (result, ms max over ranks)."""
    import torch
    E.barrier()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(E.stream)
    r = fn()
    e1.record(E.stream)
    e1.synchronize()
    ms = e0.elapsed_time(e1)
    E.barrier()
    return r, max_over_ranks(E, [ms])[0]


def bench_config3(E, args):
    """One frame of `config3_gib` GiB, 4 MiB independent blocks, block checksums (+ content checksum), strong-scaled."""
    import torch
    import zig_lz4_b200 as z
    from zig_lz4_b200 import datagen, sharded
    total = args.config3_gib * GIB
    bs = 4 << 20
    lo, hi = sharded.byte_range(E.rank, E.world, total, bs)
    n = hi - lo
    src = gen_to_device(E, n, datagen.MIXED, bs, lo // 65536)[:n]
    out = {"workload": "configs[2]: one %d GiB lz4f frame, 4 MiB independent blocks, xxh32 block checksums (+ content checksum), "
                       "synthetic mixed-entropy (class rotating every 4 MiB), %s" % (args.config3_gib,
                       "one GPU" if E.world == 1 else "sharded by block range over %d GPUs, frame gathered to rank 0, index-cut "
                       "decode, checksum state handed rank to rank" % E.world),
           "bytes_total": total, "block_size": bs, "n_gpus": E.world, "scaling": "strong"}
    for cc in (0, 1):
        zp = z.lz4f.Preferences(blockSizeID=7, blockMode=1, blockChecksumFlag=1, contentChecksumFlag=cc, contentSize=total)
        reps = 2 if cc == 0 else 1
        best_c = best_d = 1e18
        k1 = k2 = 0.0
        csize = 0
        if E.world == 1:
            cap = z.lz4f.compressFrameBound(n, zp)
            comp = torch.empty(cap + 64, dtype=torch.uint8, device=E.dev)
            back = torch.empty(n + 64, dtype=torch.uint8, device=E.dev)
            for it in range(reps + 1):     # first pass allocates the workspace
                csize, ms_c = timed(E, lambda: E.ctx.compress_frame_dev(src.data_ptr(), n, comp.data_ptr(), cap, zp, E.s))
                ph_c = E.ctx.last_phase_ms()
                m, ms_d = timed(E, lambda: E.ctx.decompress_frame_dev(comp.data_ptr(), csize, back.data_ptr(), n, E.s))
                ph_d = E.ctx.last_phase_ms()
                assert m == n
                if it > 0 or reps == 1:
                    if ms_c < best_c: best_c, k1 = ms_c, ph_c[0]
                    if ms_d < best_d: best_d, k2 = ms_d, ph_d[0]
            assert torch.equal(back[:n], src), "config3 round trip mismatch"
            del comp, back
        else:
            for it in range(reps + 1):
                (frame, layout, body), ms_c = timed(E, lambda: sharded.compress_frame_sharded(E.engine, src, zp, gather_to=0))
                csize = layout.total
                (dec, _, tot), ms_d = timed(E, lambda: sharded.decompress_frame_sharded(E.engine, frame, src=0))
                assert tot == total
                if it > 0 or reps == 1:
                    best_c = min(best_c, ms_c); best_d = min(best_d, ms_d)
            assert torch.equal(dec, src), "config3 sharded round trip mismatch"
            del frame, dec, body
        rec = {"compress_gbs": round(total / best_c / 1e6, 3), "decompress_gbs": round(total / best_d / 1e6, 3),
               "compress_ms": round(best_c, 3), "decompress_ms": round(best_d, 3), "ratio": round(total / csize, 4)}
        if E.world == 1 and k1 > 0 and k2 > 0:
            rec["roofline"] = {"bound": "hbm", "kernel": "k_compress_fast<u32 tables>", "achieved": round((n + csize) / k1 / 1e6, 2),
                               "peak": E.peak, "unit": "GB/s", "frac": round((n + csize) / k1 / 1e6 / E.peak, 5), "traffic": None,
                               "kernel_ms": round(k1, 3)}
            rec["roofline_decompress"] = {"bound": "hbm", "kernel": "k_decompress", "achieved": round((n + csize) / k2 / 1e6, 2),
                                          "peak": E.peak, "unit": "GB/s", "frac": round((n + csize) / k2 / 1e6 / E.peak, 5),
                                          "traffic": None, "kernel_ms": round(k2, 3)}
        out["block_checksums" if cc == 0 else "block_and_content_checksums"] = rec
    if E.world == 1 and E.rank == 0 and not args.no_cpu_baseline:
        import b2oracle as o
        sample = GIB
        data = datagen.generate(sample, mode=datagen.MIXED, span=bs)
        threads = o.hardware_threads()
        cb = {}
        for cc in (0, 1):
            r = oracle_frame_roundtrip(data, dict(block_size_id=7, block_mode=1, block_checksum=1, content_checksum=cc,
                                                  content_size=sample), threads, 2)
            cb["block_checksums" if cc == 0 else "block_and_content_checksums"] = {
                "compress_gbs": round(r["compress_gbs"], 3), "decompress_gbs": round(r["decompress_gbs"], 3)}
        out["cpu_baseline"] = {"value": cb, "unit": "GB/s", "cores": threads, "kind": "port",
                               "sample": "1 GiB of the same generator (256 blocks of 4 MiB), best of 2; content checksum on its "
                                         "own thread next to the block workers"}
    del src
    torch.cuda.empty_cache()
    return out


def bench_config4(E, args):
    """Decompress-only, per class, 64 KiB and 4 MiB blocks; frames pre-compressed by K1 (untimed); strong-scaled total."""
    import torch
    import zig_lz4_b200 as z
    from zig_lz4_b200 import datagen, sharded
    total = args.config4_gib * GIB
    out = {"workload": "configs[3]: decompress-only, %d GiB of content per class (BASELINE names 16 GiB; --config4-gib 16 runs "
                       "it), frames of independent blocks pre-compressed by this library, %s"
                       % (args.config4_gib, "one GPU" if E.world == 1 else "block ranges over %d GPUs, each rank decodes the "
                          "frame of its range" % E.world),
           "bytes_per_class": total, "n_gpus": E.world, "scaling": "strong", "classes": {}}
    cpu = {}
    for cname, mode in (("text", datagen.TEXT), ("binary", datagen.BINARY), ("redundant", datagen.REDUNDANT)):
        lo, hi = sharded.byte_range(E.rank, E.world, total, 4 << 20)
        n = hi - lo
        src = gen_to_device(E, n, mode, 65536, lo // 65536)[:n]
        back = torch.empty(n + 64, dtype=torch.uint8, device=E.dev)
        rec = {}
        for bname, sid in (("64KiB", 4), ("4MiB", 7)):
            zp = z.lz4f.Preferences(blockSizeID=sid, blockMode=1)
            cap = z.lz4f.compressFrameBound(n, zp)
            comp = torch.empty(cap + 64, dtype=torch.uint8, device=E.dev)
            csize = E.ctx.compress_frame_dev(src.data_ptr(), n, comp.data_ptr(), cap, zp, E.s)
            best, k2, idx = 1e18, 0.0, 0.0
            for _ in range(3):
                m, ms = timed(E, lambda: E.ctx.decompress_frame_dev(comp.data_ptr(), csize, back.data_ptr(), n, E.s))
                ph = E.ctx.last_phase_ms()
                assert m == n
                if ms < best: best, k2, idx = ms, ph[0], ph[2]
            assert torch.equal(back[:n], src), "config4 mismatch"
            csum = sum_over_ranks(E, [csize])[0]
            k2m, idxm = max_over_ranks(E, [k2, idx])
            rec[bname] = {"decompress_gbs": round(total / best / 1e6, 2), "kernel_gbs": round(total / k2m / 1e6, 2),
                          "index_ms": round(idxm, 3), "ratio": round(total / csum, 4),
                          "roofline": {"bound": "hbm", "kernel": "k_decompress", "achieved": round((total + csum) / k2m / 1e6, 2),
                                       "peak": E.peak * E.world, "unit": "GB/s",
                                       "frac": round((total + csum) / k2m / 1e6 / (E.peak * E.world), 5), "traffic": None}}
            del comp
        out["classes"][cname] = rec
        if E.world == 1 and E.rank == 0 and not args.no_cpu_baseline:
            import b2oracle as o
            sample = 512 << 20
            data = datagen.generate(sample, mode=mode, span=65536)
            threads = o.hardware_threads()
            cpu[cname] = {b: round(oracle_frame_roundtrip(data, dict(block_size_id=sid, block_mode=1), threads, 2)["decompress_gbs"], 3)
                          for b, sid in (("64KiB", 4), ("4MiB", 7))}
        del src, back
        torch.cuda.empty_cache()
    if cpu:
        import b2oracle as o
        out["cpu_baseline"] = {"value": cpu, "unit": "GB/s (decompress)", "cores": o.hardware_threads(), "kind": "port",
                               "sample": "512 MiB per class, best of 2, serial header walk + one block per task"}
    return out


def bench_config5(E, args):
    """compressHC level 9: 256 KiB blocks (text, binary) and 4 KiB records; GPU and CPU GB/s, both ratios."""
    import ctypes as C
    import numpy as np
    import torch
    import zig_lz4_b200 as z
    from zig_lz4_b200 import datagen, sharded
    total = args.config5_gib * GIB
    out = {"workload": "configs[4]: compressHC level 9, %d GiB per class in 256 KiB independent blocks (text-like, binary) and "
                       "4 KiB records; reference ratio for the dictionary case is dictionary-blind (SURVEY F6)" % args.config5_gib,
           "bytes_per_class": total, "n_gpus": E.world, "scaling": "strong", "blocks_256KiB": {}}
    cpu = {}
    want_cpu = E.world == 1 and E.rank == 0 and not args.no_cpu_baseline
    if want_cpu:
        import b2oracle as o
        threads = o.hardware_threads()
    for cname, mode in (("text", datagen.TEXT), ("binary", datagen.BINARY)):
        lo, hi = sharded.byte_range(E.rank, E.world, total, 256 << 10)
        n = hi - lo
        src = gen_to_device(E, n, mode, 65536, lo // 65536)[:n]
        zp = z.lz4f.Preferences(blockSizeID=5, blockMode=1, compressionLevel=9)
        cap = z.lz4f.compressFrameBound(n, zp)
        comp = torch.empty(cap + 64, dtype=torch.uint8, device=E.dev)
        back = torch.empty(n + 64, dtype=torch.uint8, device=E.dev)
        best, k3, csize = 1e18, 0.0, 0
        for it in range(2):              # first pass allocates the HC tables
            csize, ms = timed(E, lambda: E.ctx.compress_frame_dev(src.data_ptr(), n, comp.data_ptr(), cap, zp, E.s))
            ph = E.ctx.last_phase_ms()
            if it > 0 and ms < best: best, k3 = ms, ph[0]
        m = E.ctx.decompress_frame_dev(comp.data_ptr(), csize, back.data_ptr(), n, E.s)
        assert m == n and torch.equal(back[:n], src), "config5 HC round trip mismatch"
        csum = sum_over_ranks(E, [csize])[0]
        k3m = max_over_ranks(E, [k3])[0]
        rec = {"compress_gbs": round(total / best / 1e6, 3), "ratio": round(total / csum, 4),
               "roofline": {"bound": "hbm", "kernel": "k_compress_hc", "achieved": round((total + csum) / k3m / 1e6, 3),
                            "peak": E.peak * E.world, "unit": "GB/s",
                            "frac": round((total + csum) / k3m / 1e6 / (E.peak * E.world), 6), "traffic": None}}
        if want_cpu:
            sample = 256 << 20
            data = datagen.generate(sample, mode=mode, span=65536)
            r = oracle_frame_roundtrip(data, dict(block_size_id=5, block_mode=1, compression_level=9), threads, 1)
            rec["cpu_gbs"] = round(r["compress_gbs"], 4)
            rec["cpu_ratio"] = round(r["ratio"], 4)
            fr = oracle_frame_roundtrip(data[:64 << 20], dict(block_size_id=5, block_mode=1), threads, 1)
            rec["fast_mode_ratio"] = round(fr["ratio"], 4)
        out["blocks_256KiB"][cname] = rec
        del src, comp, back
        torch.cuda.empty_cache()
    # ---- 4 KiB records (device batch API), text-like; dictionary-blind HC-9 == the reference's behaviour ----
    if E.world == 1:
        nrec = total // 4096
        src = gen_to_device(E, total, datagen.TEXT, 65536, 1 << 20)[:total]
        rcap = z.lz4.compressBound(4096)
        ar = torch.arange(nrec, dtype=torch.int64, device=E.dev)
        soff = (ar * 4096).contiguous()
        doff = (ar * rcap).contiguous()
        slen = torch.full((nrec,), 4096, dtype=torch.int32, device=E.dev)
        dcap = torch.full((nrec,), rcap, dtype=torch.int32, device=E.dev)
        olen = torch.zeros(nrec, dtype=torch.int32, device=E.dev)
        stat = torch.zeros(nrec, dtype=torch.int32, device=E.dev)
        dst = torch.empty(nrec * rcap + 64, dtype=torch.uint8, device=E.dev)
        best = 1e18
        for it in range(2):
            _, ms = timed(E, lambda: E.ctx.compress_hc_batch_dev(src.data_ptr(), soff.data_ptr(), slen.data_ptr(), dst.data_ptr(),
                                                                 doff.data_ptr(), dcap.data_ptr(), olen.data_ptr(), stat.data_ptr(),
                                                                 nrec, 9, E.s))
            if it > 0: best = min(best, ms)
        assert int(stat.abs().sum().item()) == 0
        csum = int(olen.to(torch.int64).sum().item())
        rec = {"records": nrec, "compress_gbs": round(total / best / 1e6, 3), "ratio": round(total / csum, 4),
               "roofline": {"bound": "hbm", "kernel": "k_compress_hc", "achieved": round((total + csum) / best / 1e6, 3),
                            "peak": E.peak, "unit": "GB/s", "frac": round((total + csum) / best / 1e6 / E.peak, 6), "traffic": None}}
        # shared 64 KiB dictionary, fast mode with real dictionary matches (k_compress_dict.cu) on a 65 536-record sample
        ns = 65536
        hs = src[:ns * 4096].cpu().numpy()
        dic = datagen.generate(65536, seed=SEED + 77, mode=datagen.TEXT).tobytes()
        so = np.arange(ns, dtype=np.uint64) * 4096
        do = np.arange(ns, dtype=np.uint64) * rcap
        sl = np.full(ns, 4096, dtype=np.uint32)
        dc = np.full(ns, rcap, dtype=np.uint32)
        res = E.ctx.compress_fast_dict_batch(hs, so, sl, ns * rcap, do, dc, dic)
        ol_d = np.asarray(res[1], dtype=np.int64)
        res0 = E.ctx.compress_fast_batch(hs, so, sl, ns * rcap, do, dc)
        ol_0 = np.asarray(res0[1], dtype=np.int64)
        rec["fast_mode_ratio"] = round(ns * 4096 / float(ol_0.sum()), 4)
        rec["fast_mode_shared_dictionary_ratio"] = round(ns * 4096 / float(ol_d.sum()), 4)
        if want_cpu:
            import b2oracle as o
            L = o.lib()
            cdst = np.empty(ns * rcap, dtype=np.uint8)
            col = np.zeros(ns, dtype=np.uint32)
            cst = np.zeros(ns, dtype=np.int32)
            t0 = time.perf_counter()
            L.b2o_batch(2, 9, hs.ctypes.data, so.ctypes.data, sl.ctypes.data, cdst.ctypes.data, do.ctypes.data, dc.ctypes.data,
                        col.ctypes.data, cst.ctypes.data, ns, threads)
            dt = time.perf_counter() - t0
            assert int(np.abs(cst).sum()) == 0
            rec["cpu_gbs"] = round(ns * 4096 / dt / 1e9, 4)
            rec["cpu_ratio"] = round(ns * 4096 / float(col.astype(np.int64).sum()), 4)
            gpu_sample = olen[:ns].cpu().numpy()
            rec["bytes_equal_cpu_on_sample"] = bool((gpu_sample == col.astype(np.int32)).all())
        out["records_4KiB"] = rec
        del src, dst
        torch.cuda.empty_cache()
    if want_cpu:
        out["cpu_baseline"] = {"value": {k: v.get("cpu_gbs") for k, v in list(out["blocks_256KiB"].items()) +
                                         [("records_4KiB", out.get("records_4KiB", {}))]},
                               "unit": "GB/s (compress)", "cores": threads, "kind": "port",
                               "sample": "256 MiB per class in 256 KiB blocks / 65 536 records, one pass, one block per task"}
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import zig_lz4_b200 as z
    from zig_lz4_b200 import datagen, sharded

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.bytes
    bs = 65536
    zp = z.lz4f.Preferences(blockSizeID=z.lz4f.BlockSizeID.max64KB, blockMode=z.lz4f.BlockMode.independent)
    # N > 1: the frame is sharded by block range, one process per GPU (zig-lz4_b200/sharded.py)
    engine = sharded.CudaEngine(local) if world > 1 else None
    ctx = engine.ctx if engine else z.Context(local)
    # a real (non-default) stream: the library maps stream 0 to the context's own stream, and the CUDA events below
    # must sit on the stream the kernels are launched on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    s = stream.cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    peak, peak_src = peaks()
    E = Env()
    E.rank, E.world, E.dev, E.ctx, E.engine, E.stream, E.s, E.barrier, E.peak, E.stage = rank, world, dev, ctx, engine, stream, s, barrier, peak, None

    if world > 1:
        # pre-flight of the multi-rank product path over NCCL (CudaEngine): gathered frame decodes back, a corrupted
        # content checksum is refused on every rank
        small = gen_to_device(E, 8 << 20, datagen.MIXED, 65536, rank * 128)[:8 << 20]
        zq = z.lz4f.Preferences(blockMode=1, blockChecksumFlag=1, contentChecksumFlag=1, contentSize=(8 << 20) * world)
        frame, layout, _ = sharded.compress_frame_sharded(engine, small, zq, gather_to=0)
        dec, _, tot = sharded.decompress_frame_sharded(engine, frame, src=0)
        assert tot == (8 << 20) * world and torch.equal(dec, small), "sharded pre-flight mismatch"
        if rank == 0:
            frame[layout.total - 1] ^= 1
        try:
            sharded.decompress_frame_sharded(engine, frame, src=0)
            raise SystemExit("sharded pre-flight: corrupted content checksum was accepted")
        except z.B2Error as e:
            assert e.code == sharded.ERR_CONTENT_CHECKSUM_INVALID, e
        del small, frame, dec

    # ---- synthetic shard of this rank (rank r holds blocks [r*B, (r+1)*B) of the world-sized frame) ----
    host = torch.empty(n, dtype=torch.uint8).pin_memory()
    datagen.fill_ptr(host.data_ptr(), n, seed=SEED + rank * (n // 65536), mode=datagen.MIXED, span=65536)
    src = host.to(dev, non_blocking=False)
    cap = z.lz4f.compressFrameBound(n, zp)
    comp = torch.empty(cap + 64, dtype=torch.uint8, device=dev)
    back = torch.empty(n + 64, dtype=torch.uint8, device=dev)

    def step():
        """compress the shard, exchange body sizes (N>1), decompress it back.  Returns (csize, tc_ms, td_ms, phases)."""
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e2 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        if world == 1:
            csize = ctx.compress_frame_dev(src.data_ptr(), n, comp.data_ptr(), cap, zp, s)
        else:
            # body of this rank's blocks + the one all_gather of body sizes that fixes the frame layout
            _, layout, body = sharded.compress_frame_sharded(engine, src, zp, gather_to=None)
            csize = body.numel()
        ph_c = ctx.last_phase_ms()
        e1.record(stream)
        if world == 1:
            m = ctx.decompress_frame_dev(comp.data_ptr(), csize, back.data_ptr(), n, s)
        else:
            m = ctx.decompress_blocks_dev(body.data_ptr(), csize, back.data_ptr(), n, bs, False, s)
        ph_d = ctx.last_phase_ms()
        e2.record(stream)
        e2.synchronize()
        assert m == n
        return csize, e0.elapsed_time(e1), e1.elapsed_time(e2), ph_c, ph_d

    ctx.set_timing(True)
    for _ in range(max(args.warmup, 3)):
        step()
    assert torch.equal(back[:n], src), "round trip mismatch"

    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = z.kernel_launch_count()
    t_all0 = torch.cuda.Event(enable_timing=True); t_all1 = torch.cuda.Event(enable_timing=True)
    t_all0.record(stream)
    tc = td = 0.0
    kc = kd = walk = 0.0
    csize = 0
    for _ in range(args.steps):
        csize, a, b, ph_c, ph_d = step()
        tc += a; td += b
        kc += ph_c[0]; kd += ph_d[0]; walk += ph_d[2]
    t_all1.record(stream)
    barrier()
    launches = z.kernel_launch_count() - launches0
    total_ms = t_all0.elapsed_time(t_all1)
    if world > 1:
        total_ms, tc, td = max_over_ranks(E, [total_ms, tc, td])
        launches = int(sum_over_ranks(E, [launches])[0])
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e9
    comp_gbs = world * n / (tc / args.steps * 1e-3) / 1e9
    dec_gbs = world * n / (td / args.steps * 1e-3) / 1e9

    # ---- roofline of the dominant kernel (K1, the compressor) and of the decompressor (K2) ----
    kc_ms = kc / args.steps
    kd_ms = kd / args.steps
    roof_c = (n + csize) / (kc_ms * 1e-3) / 1e9
    roof_d = (n + csize) / (kd_ms * 1e-3) / 1e9
    tr_c, tr_d, tr_why = measured_traffic(n)

    # ---- e2e: host-pointer C-ABI, H2D and D2H inside the timed region; pinned buffers, then pageable ones ----
    e2e = None
    if not args.no_e2e:
        def e2e_run(hsrc, hdst, hbk, steps):
            def one():
                cs = ctx.compress_frame(hsrc, zp, dst=hdst)
                m = ctx.decompress_frame(hdst[:cs], dst=hbk)
                assert m == n
                return cs
            for _ in range(2):
                cs = one()
            assert (hbk == hsrc).all()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                cs = one()
            barrier()
            dt = time.perf_counter() - t0
            if world > 1:
                dt = max_over_ranks(E, [dt])[0]
            return dt, cs

        hcomp = torch.empty(cap, dtype=torch.uint8).pin_memory()
        hback = torch.empty(n, dtype=torch.uint8).pin_memory()
        hsrc = np.frombuffer((C_ubyte_array(host.data_ptr(), n)), dtype=np.uint8)
        hdst = np.frombuffer((C_ubyte_array(hcomp.data_ptr(), cap)), dtype=np.uint8)
        hbk = np.frombuffer((C_ubyte_array(hback.data_ptr(), n)), dtype=np.uint8)
        dt, cs = e2e_run(hsrc, hdst, hbk, args.steps)
        e2e = {"value": round(world * n * args.steps / dt / 1e9, 4), "unit": "GB/s", "h2d_bytes_per_step": world * (n + cs),
               "d2h_bytes_per_step": world * (cs + n), "ms_per_step": round(dt / args.steps * 1e3, 3),
               "pcie_gbs_per_direction": round((n + cs) * args.steps / dt / 1e9, 2),
               "api": "b2lz4f_compress_frame_ctx + b2lz4f_decompress_frame_ctx (host pointers, pinned)"}
        # the link itself, for the reader: raw pinned copies of 512 MiB, each direction alone and both at once
        try:
            pn = 512 << 20
            d_a = torch.empty(pn, dtype=torch.uint8, device="cuda"); d_b = torch.empty(pn, dtype=torch.uint8, device="cuda")
            s_a, s_b = torch.cuda.Stream(), torch.cuda.Stream()

            def wall(fn):
                best = 1e9
                for _ in range(3):
                    torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
                    best = min(best, time.perf_counter() - t0)
                return best

            def up():
                with torch.cuda.stream(s_a):
                    d_a.copy_(host[:pn], non_blocking=True)

            def down():
                with torch.cuda.stream(s_b):
                    hback[:pn].copy_(d_b, non_blocking=True)

            e2e["pcie_measured_gbs"] = {"h2d": round(pn / wall(up) / 1e9, 1), "d2h": round(pn / wall(down) / 1e9, 1),
                                        "each_way_when_both_run": round(pn / wall(lambda: (up(), down())) / 1e9, 1),
                                        "floor_ms_per_step": None}
            both = e2e["pcie_measured_gbs"]["each_way_when_both_run"]
            # two synchronous calls: the upload of N bytes bounds the first, the download of N bytes the second
            e2e["pcie_measured_gbs"]["floor_ms_per_step"] = round(2 * n / both / 1e6, 1)
            del d_a, d_b
        except Exception as ex:      # never let the side measurement cost the line
            e2e["pcie_measured_gbs"] = {"error": str(ex)[:80]}
        # what a caller with ordinary (pageable) slices gets: same calls, numpy-owned buffers
        psrc = np.array(hsrc, copy=True)
        pdst = np.empty(cap, dtype=np.uint8)
        pbk = np.empty(n, dtype=np.uint8)
        psteps = max(1, min(args.steps, 3))
        dtp, _ = e2e_run(psrc, pdst, pbk, psteps)
        e2e["pageable"] = {"value": round(world * n * psteps / dtp / 1e9, 4), "unit": "GB/s",
                           "ms_per_step": round(dtp / psteps * 1e3, 3),
                           "note": "same calls with pageable host buffers (a Zig caller's slices)"}
        del hcomp, hback, psrc, pdst, pbk

    clocks = sampler.stop() if rank == 0 else None      # sampled over both timed regions (device-resident and e2e)
    del comp, back, src
    torch.cuda.empty_cache()

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        import b2oracle as o
        o.prefer_native()                                 # every CPU leg below runs the -march=native build of the oracle
    extra = {}
    if not args.no_extra:
        for name, fn in (("config3", bench_config3), ("config4", bench_config4), ("config5", bench_config5)):
            t0 = time.perf_counter()
            try:
                extra[name] = fn(E, args)
                extra[name]["bench_seconds"] = round(time.perf_counter() - t0, 1)
            except Exception as e:   # a sub-record must not take the headline down; the failure is reported in its place
                import traceback
                extra[name] = {"error": "%s: %s" % (type(e).__name__, e), "trace": traceback.format_exc()[-600:]}
                if world > 1:
                    raise

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 3), "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": headline_config(n),
            "notes": {"sharding": "block range per rank, all-gather of body sizes",
                      "l2": "inputs (1 GiB raw, ~0.55 GiB compressed per GPU) are larger than the 126 MB L2; no flush needed",
                      "library": library_version()},
            "compress_gbs": round(comp_gbs, 3), "decompress_gbs": round(dec_gbs, 3), "ratio": round(n / csize, 4),
            "roofline": {"bound": "hbm", "kernel": "k_compress_fast", "achieved": round(roof_c, 2), "peak": peak, "unit": "GB/s",
                         "frac": round(roof_c / peak, 5), "traffic": tr_c["traffic_bytes"] if tr_c else None, "peak_source": peak_src,
                         "algorithmic_bytes": n + csize, "kernel_ms": round(kc_ms, 4),
                         "issue_slots_busy_pct_ncu": tr_c["issue_active_pct"] if tr_c else None,
                         "traffic_note": tr_why or "profiles/traffic.json (ncu --set full capture of this workload and build)",
                         "note": "byte-serial LZ77 per block: bound by instruction issue and dependent-load latency, not by HBM "
                                 "(DESIGN.md section 4)"},
            "roofline_decompress": {"bound": "hbm", "kernel": "k_decompress", "achieved": round(roof_d, 2), "peak": peak,
                                    "unit": "GB/s", "frac": round(roof_d / peak, 5), "traffic": tr_d["traffic_bytes"] if tr_d else None,
                                    "algorithmic_bytes": n + csize, "kernel_ms": round(kd_ms, 4),
                                    "issue_slots_busy_pct_ncu": tr_d["issue_active_pct"] if tr_d else None,
                                    "index_walk_ms": round(walk / args.steps, 4)},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        if not args.no_cpu_baseline and world == 1:
            import b2oracle as o
            flags = o.prefer_native()
            threads = o.hardware_threads()
            sample = min(n, GIB)
            r = cpu_arm(sample, threads, 5, 1)
            line["cpu_baseline"] = {"value": round(r["roundtrip_best_gbs"], 4), "unit": "GB/s", "cores": threads, "kind": "port",
                                    "compress_gbs": round(r["compress_gbs"], 4), "decompress_gbs": round(r["decompress_gbs"], 4),
                                    "sample": "%d MiB of the same workload, 1 warm-up + best of 5 passes; C restatement of the Zig "
                                              "reference (oracle/, gcc %s built on this box), one block per task on all host "
                                              "threads, parallel frame assembly" % (sample >> 20, flags)}
        else:
            line["cpu_baseline"] = None
        line.update(extra)
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def C_ubyte_array(ptr, n):
    import ctypes as C
    return (C.c_uint8 * n).from_address(ptr)


if __name__ == "__main__":
    main()
