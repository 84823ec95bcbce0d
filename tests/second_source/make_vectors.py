#!/usr/bin/env python
"""Writes tests/golden/second_source_vectors.json from the independent Python restatement
(tests/second_source/zlz4_second.py) — NOT from the C oracle.  Run from the repo root:

    python tests/second_source/make_vectors.py

Inputs are regenerated deterministically by tests/second_source/inputs.py (the bench's corpus generator for
the four SURVEY §8d classes + the reference's own test inputs + the F8 hazard input), so only hashes and
sizes are stored (the bytes themselves for outputs of <= 48 bytes)."""
import hashlib
import itertools
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), HERE]

import inputs  # noqa: E402
import zlz4_second as s2  # noqa: E402


def entry(name, data, comp, **kw):
    e = dict(name=name, n=len(data), sha1_in=hashlib.sha1(data).hexdigest(), csize=len(comp),
             sha1_out=hashlib.sha1(comp).hexdigest())
    if len(comp) <= 48:
        e["out_hex"] = comp.hex()
    e.update(kw)
    return e


def main():
    t0 = time.time()
    V = {"_about": "outputs of tests/second_source/zlz4_second.py (independent Python restatement of the Zig reference); "
                   "written by tests/second_source/make_vectors.py", "blocks": [], "frames": [], "decode_errors": []}
    blocks = inputs.block_inputs()
    for name, data in blocks.items():
        c = s2.compress_fast(data, 1)
        assert s2.decompress_safe(c, len(data)) == data
        V["blocks"].append(entry(name, data, c, kind="fast", accel=1))
        if len(data) <= 65536 and not name.startswith("tiny"):
            for accel in (3, 70):
                V["blocks"].append(entry(name, data, s2.compress_fast(data, accel), kind="fast", accel=accel))
        for level in (3, 6, 9):
            if name.startswith("tiny") and level != 9:
                continue
            c = s2.compress_hc(data, level)
            assert s2.decompress_safe(c, len(data)) == data
            V["blocks"].append(entry(name, data, c, kind="hc", level=level))
        print("block %-22s %7d B  %.1fs" % (name, len(data), time.time() - t0), flush=True)
    V["f8_guard_hits"] = s2.F8_HITS

    cache = {}
    frames = inputs.frame_inputs()
    # every combination of the preferences that change the bytes, fast mode, on the multi-block input
    combos = list(itertools.product((0, 4, 5, 6, 7), (0, 1), (0, 1), (0, 1), (0, 1), (0, 77)))
    for sid, mode, bchk, cchk, csz, did in combos:
        data = frames["mixed_300k"]
        kw = dict(block_size_id=sid, block_mode=mode, block_checksum=bchk, content_checksum=cchk,
                  content_size=len(data) if csz else 0, dict_id=did, compression_level=0)
        f = s2.compress_frame(data, s2.Prefs(**kw), block_cache=cache)
        assert s2.decompress_frame(f, len(data)) == data
        V["frames"].append(entry("mixed_300k", data, f, prefs=kw))
    # HC frames + the other inputs on a reduced set of combinations
    for name, data in frames.items():
        for sid, bchk, cchk, level in ((0, 0, 0, 0), (4, 1, 1, 0), (5, 1, 1, 9), (7, 0, 1, 3), (6, 1, 0, 6)):
            if name == "mixed_300k" and level == 0:
                continue
            kw = dict(block_size_id=sid, block_mode=1, block_checksum=bchk, content_checksum=cchk,
                      content_size=len(data), dict_id=0, compression_level=level)
            f = s2.compress_frame(data, s2.Prefs(**kw), block_cache=cache)
            assert s2.decompress_frame(f, len(data) + 5) == data
            V["frames"].append(entry(name, data, f, prefs=kw))
        print("frame %-22s %7d B  %.1fs" % (name, len(data), time.time() - t0), flush=True)

    # decoder error kinds (src/lz4.zig:111-248, src/lz4f.zig:541-638): hostile inputs, expected member name
    for name, blob, cap, dic in inputs.hostile_blocks():
        try:
            out = s2.decompress_safe(blob, cap, dic)
            res = dict(ok=True, n_out=len(out), sha1_out=hashlib.sha1(out).hexdigest())
        except s2.Lz4Error as e:
            res = dict(ok=False, error=e.kind)
        V["decode_errors"].append(dict(name=name, kind="block", cap=cap, **res))
    for name, blob, cap in inputs.hostile_frames(s2):
        try:
            out = s2.decompress_frame(blob, cap)
            res = dict(ok=True, n_out=len(out), sha1_out=hashlib.sha1(out).hexdigest())
        except s2.Lz4Error as e:
            res = dict(ok=False, error=e.kind)
        V["decode_errors"].append(dict(name=name, kind="frame", cap=cap, **res))

    path = os.path.join(ROOT, "tests", "golden", "second_source_vectors.json")
    with open(path, "w") as f:
        json.dump(V, f, indent=0, sort_keys=True)
    print("wrote %s: %d blocks, %d frames, %d decode cases, F8 guard hits %d, %.0fs"
          % (path, len(V["blocks"]), len(V["frames"]), len(V["decode_errors"]), V["f8_guard_hits"], time.time() - t0))


if __name__ == "__main__":
    main()
