#!/usr/bin/env python
"""Writes tests/golden/vectors.json: compressed bytes for the reference's own test inputs.

The reference (pure Zig) cannot be built in this image, so these vectors are produced by the C restatement in
oracle/ (each function cites the Zig lines it follows) AFTER it reproduced every second-source vector of
SURVEY.md §8(c) and was accepted by the stock decoders (tests/test_oracle.py).  They pin today's oracle
behaviour: a change in oracle/ or in the kernels that alters a single output byte fails the golden tests.
Run:  python tests/golden/make_golden.py"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import b2oracle as o
import corpus


def entry(name, data, comp):
    e = {"name": name, "n": len(data), "csize": len(comp), "sha1_in": hashlib.sha1(data).hexdigest(),
         "sha1_out": hashlib.sha1(comp).hexdigest()}
    if len(data) <= 256:
        e["in_hex"] = data.hex()
    if len(comp) <= 512:
        e["out_hex"] = comp.hex()
    return e


def main():
    out = {"blocks_fast": [], "blocks_hc9": [], "frames": []}
    cases = corpus.block_cases() + corpus.compat_cases()
    for name, data in cases:
        for i in range(0, max(len(data), 1), 65536):
            b = data[i:i + 65536]
            out["blocks_fast"].append(entry("%s@%d" % (name, i), b, o.compress_fast(b)))
            if len(b) <= 20000:
                out["blocks_hc9"].append(entry("%s@%d" % (name, i), b, o.compress_hc(b, 9)))
    prefs = [dict(), dict(block_mode=1, content_checksum=1), dict(block_mode=1, block_checksum=1, content_checksum=1),
             dict(block_size_id=5, block_mode=1), dict(block_mode=1, compression_level=9)]
    for name, data in corpus.compat_cases():
        for kw in prefs:
            f = o.compress_frame(data, o.make_prefs(**kw))
            e = entry(name, data, f)
            e["prefs"] = kw
            out["frames"].append(e)
    with open(os.path.join(HERE, "vectors.json"), "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)
    print("wrote", sum(len(v) for v in out.values()), "vectors")


if __name__ == "__main__":
    main()
