"""Golden fixtures (tests/golden/vectors.json, written by tests/golden/make_golden.py): compressed bytes of the
reference's own test inputs.  CPU: the oracle must still produce them.  GPU: the CUDA path must produce them
through the C-ABI (blocks: K1 / K3, frames: the whole frame writer) and decode them back."""
import hashlib
import json
import os

import pytest

import corpus

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "vectors.json")) as _f:
    G = json.load(_f)


def _inputs():
    d = {}
    for name, data in corpus.block_cases() + corpus.compat_cases():
        for i in range(0, max(len(data), 1), 65536):
            d["%s@%d" % (name, i)] = data[i:i + 65536]
    return d


def _check(e, comp):
    assert len(comp) == e["csize"], e["name"]
    assert hashlib.sha1(comp).hexdigest() == e["sha1_out"], e["name"]
    if "out_hex" in e:
        assert comp.hex() == e["out_hex"], e["name"]


def test_golden_inputs_are_the_corpus():
    ins = _inputs()
    for e in G["blocks_fast"]:
        assert hashlib.sha1(ins[e["name"]]).hexdigest() == e["sha1_in"]


def test_oracle_reproduces_golden(oracle):
    ins = _inputs()
    for e in G["blocks_fast"]:
        _check(e, oracle.compress_fast(ins[e["name"]]))
    for e in G["blocks_hc9"]:
        _check(e, oracle.compress_hc(ins[e["name"]], 9))
    full = dict(corpus.compat_cases())
    for e in G["frames"]:
        _check(e, oracle.compress_frame(full[e["name"]], oracle.make_prefs(**e["prefs"])))


@pytest.mark.gpu
def test_cuda_blocks_reproduce_golden(z):
    ins = _inputs()
    for e in G["blocks_fast"]:
        c = z.lz4.compressDefault(ins[e["name"]])
        _check(e, c)
        assert z.lz4.decompressSafe(c, e["n"]) == ins[e["name"]]
    for e in G["blocks_hc9"]:
        c = z.lz4hc.compressHC(ins[e["name"]], 9)
        _check(e, c)
        assert z.lz4.decompressSafe(c, e["n"]) == ins[e["name"]]


@pytest.mark.gpu
def test_cuda_frames_reproduce_golden(z):
    full = dict(corpus.compat_cases())
    for e in G["frames"]:
        kw = e["prefs"]
        zp = z.lz4f.Preferences(blockSizeID=kw.get("block_size_id", 0), blockMode=kw.get("block_mode", 0),
                                contentChecksumFlag=kw.get("content_checksum", 0), blockChecksumFlag=kw.get("block_checksum", 0),
                                compressionLevel=kw.get("compression_level", 0))
        f = z.lz4f.compressFrame(full[e["name"]], zp)
        _check(e, f)
        assert z.lz4f.decompressFrame(f, e["n"] + 8) == full[e["name"]]
