"""CPU tests (no GPU): the C-ABI library loads, exports every symbol include/b2lz4.h declares, its
host-side pieces (bounds, header codec, status names) agree with the oracle, and compute entry points
fail loudly instead of falling back when there is no CUDA device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "b2lz4.h")).read()
    return sorted(set(re.findall(r"B2LZ4_API\s+[^;(]*?\b(b2lz4f?_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(z):
    names = declared_symbols()
    assert len(names) >= 40
    L = C.CDLL(z.library_path())
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    from zig_lz4_b200 import _native
    assert sorted(_native.EXPORTS) == names


def test_no_oracle_in_product():
    """The product tree must not reference the oracle (it is test infrastructure)."""
    pkg = os.path.join(ROOT, "zig-lz4_b200")
    for dp, dn, fn in os.walk(pkg):
        if "build" in dp:
            continue
        for f in fn:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")):
                s = open(os.path.join(dp, f), errors="ignore").read()
                assert "b2oracle" not in s and "b2o_" not in s and "libb2oracle" not in s, f


def test_bounds_match_oracle(z, oracle):
    for n in (0, 1, 12, 13, 254, 255, 256, 4096, 65535, 65536, 262144, 1 << 20, 4 << 20, 0x7E000000, 0x7E000001):
        assert z.lz4.compressBound(n) == oracle.compress_bound(n)
    assert z.lz4.compressBound(65536) == 65809 and z.lz4.compressBound(4096) == 4128
    for kw in (dict(), dict(block_size_id=7, block_checksum=1, content_checksum=1), dict(block_size_id=5)):
        zp = z.lz4f.Preferences(blockSizeID=kw.get("block_size_id", 0), blockChecksumFlag=kw.get("block_checksum", 0),
                                contentChecksumFlag=kw.get("content_checksum", 0))
        op = oracle.make_prefs(**kw)
        for n in (0, 1, 65536, 65537, 100000, 10 << 20):
            assert z.lz4f.compressFrameBound(n, zp) == oracle.compress_frame_bound(n, op)
    assert z.lz4f.compressFrameBound(100000) == oracle.compress_frame_bound(100000)


def test_header_codec_matches_oracle(z, oracle):
    L = oracle.lib()
    for kw in (dict(), dict(block_mode=1), dict(content_checksum=1, block_checksum=1), dict(content_size=12345),
               dict(dict_id=77, content_size=1 << 40, block_size_id=7, block_mode=1)):
        zp = z.lz4f.Preferences(blockSizeID=kw.get("block_size_id", 0), blockMode=kw.get("block_mode", 0),
                                contentChecksumFlag=kw.get("content_checksum", 0), contentSize=kw.get("content_size", 0),
                                dictID=kw.get("dict_id", 0), blockChecksumFlag=kw.get("block_checksum", 0))
        hz = z.lz4f.writeFrameHeader(zp)
        buf = (C.c_uint8 * 32)()
        out = C.c_size_t(0)
        op = oracle.make_prefs(**kw)
        assert L.b2o_write_frame_header(buf, 32, C.byref(op), C.byref(out)) == 0
        assert hz == bytes(buf[:out.value])
        info, size = z.lz4f.parseFrameHeader(hz + b"\0\0\0\0")
        assert size == len(hz) == z.lz4f.headerSize(hz)
        assert info.content_size == kw.get("content_size", 0) and info.dict_id == kw.get("dict_id", 0)
        assert info.block_checksum == kw.get("block_checksum", 0)
    assert z.lz4f.writeFrameHeader(None) == bytes.fromhex("04224d184040c0")                # SURVEY a-12
    for hdr, name in ((b"\x04\x22\x4d", "lz4f.FrameHeaderIncomplete"), (b"\0\0\0\0\x40\x40\xc0", "lz4f.FrameTypeUnknown"),
                      (b"\x04\x22\x4d\x18\x80\x40\xc0", "lz4f.HeaderVersionWrong"),
                      (b"\x04\x22\x4d\x18\x42\x40\xc0", "lz4f.ReservedFlagSet"),
                      (b"\x04\x22\x4d\x18\x40\x10\xc0", "lz4f.MaxBlockSizeInvalid"),
                      (b"\x04\x22\x4d\x18\x40\x40\xc1", "lz4f.HeaderChecksumInvalid")):
        with pytest.raises(z.B2Error) as e:
            z.lz4f.parseFrameHeader(hdr)
        assert e.value.name == name
    assert z.lz4f.headerSize(b"\x50\x2a\x4d\x18\x00") == 8


def test_status_names(z, oracle):
    L = z.lib()
    for code in list(range(0, 7)) + list(range(100, 123)):
        assert L.b2lz4_status_name(code).decode() == oracle.status_name(code)


def test_compute_fails_loudly_without_gpu(z):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(z.B2Error) as e:
        z.lz4.compressDefault(b"hello world hello world")
    assert e.value.name == "b2lz4.CudaError"
    with pytest.raises(z.B2Error):
        z.lz4f.compressFrame(b"x" * 100)
    with pytest.raises(z.B2Error):
        z.Context(0)


def _build_abi_check(tmp_path):
    import subprocess
    exe = str(tmp_path / "abi_check")
    libdir = os.path.join(ROOT, "zig-lz4_b200")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi", "abi_check.c"), "-o", exe,
                           "-L", libdir, "-l:libb2lz4.so", "-Wl,-rpath," + libdir])
    return exe


def test_header_compiles_as_c_and_calls_fail_loudly_without_gpu(z, tmp_path):
    """include/b2lz4.h through a C compiler (-Wall -Wextra -Werror): the prototypes the Zig shim binds are checked by
    gcc, not only resolved by name; without a device every compute call returns B2LZ4_ERR_CUDA."""
    import subprocess
    import torch
    exe = _build_abi_check(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == ("abi ok gpu" if torch.cuda.is_available() else "abi ok nogpu")


@pytest.mark.gpu
def test_c_program_round_trips_on_gpu(z, tmp_path):
    import subprocess
    exe = _build_abi_check(tmp_path)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    assert out.stdout.strip() == "abi ok gpu"
