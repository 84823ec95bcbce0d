"""K1 / K2 / K3 device code executed on the host (tools/warp_emu: every lane a coroutine, every warp collective a rendezvous)
and compared with the oracle: the kernels' LOGIC — window walk, static match chain, batched emit, forward word ring,
chunked / serial decode front ends, exact tier, error kinds — is checked here, where there is no GPU.  The GPU suite
checks the compiled kernels; this one makes a logic regression visible in the CPU tier already.  Inputs and outputs sit
between inaccessible pages (emu::Guarded): a read or write outside the 16-byte granules that hold the buffers' own bytes
— which no tool reports on the GPU box, compute-sanitizer being closed there — is a crash here.
Reference semantics: /root/reference/src/lz4.zig:292-447 (compressFast), :89-259 (decompressGeneric)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tools", "warp_emu")


@pytest.fixture(scope="module")
def emu_built():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "libb2oracle.so"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", EMU, "all"], stdout=subprocess.DEVNULL)
    return EMU


def run(binary, *args):
    r = subprocess.run([os.path.join(EMU, binary)] + [str(a) for a in args], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert " 0 failed" in r.stdout, r.stdout + r.stderr
    return r.stdout


@pytest.mark.parametrize("cls", [0, 1, 2, 3, 4])
def test_k1_fast_path_equals_oracle(emu_built, cls):
    """64 KiB blocks of every data class + edge sizes, capacities and unaligned starts (u16 and u32 tables), acceleration 1"""
    out = run("emu_k1", cls, 6)
    if cls in (0, 4):
        assert "batched 0)" not in out          # the static chain really ran


@pytest.mark.parametrize("accel", [2, 7, 70, 65537])
def test_k1_general_path_equals_oracle(emu_built, accel):
    run("emu_k1", 4, 1, 65536, accel)


def test_k1_large_block_u32_tables_deferred_flush(emu_built):
    """large blocks: u32 tables, the variant that writes a window's batch under the next window's candidate reads — 1 MiB of
    text, 1 MiB of binary records, and a 4 MiB block of all four classes (positions beyond 2^16 and 2^21)"""
    run("emu_k1", 0, 1, 1048576)
    run("emu_k1", 1, 1, 1048576)
    run("emu_k1", 4, 1, 4194304)


@pytest.mark.parametrize("cls", [0, 1, 2, 3, 4])
def test_k2_all_tiers_equal_oracle(emu_built, cls):
    """chunked front end, serial front end and exact tier: bytes, sizes and error kinds on intact, truncated and damaged streams"""
    run("emu_k2", cls, 4)


@pytest.mark.parametrize("cls", [0, 1, 2, 4])
def test_k3_hash_chain_equals_oracle(emu_built, cls):
    """compressHC (src/lz4hc.zig:976-1064 with insertAndGetWiderMatch :538-681): levels 3 / 6 / 9, the hop-by-hop walk and
    the jump-table walk, blocks compressed one after the other on ONE work area (epoch base, lazily built jump levels),
    small inputs and limited-output exits — bytes and status codes equal the oracle's"""
    run("emu_k3", cls, 2)


def test_k3_blocks_beyond_the_chain_window(emu_built):
    """64 KiB blocks of text and binary records, and 256 KiB blocks (config 5's block size): positions beyond the 65 536-entry
    chain table, distances clamped at the window, chains of up to 256 candidates, the pattern step"""
    run("emu_k3", 0, 1, 65536)
    run("emu_k3", 1, 1, 65536)
    run("emu_k3", 0, 1, 262144)
    run("emu_k3", 2, 1, 262144)


def test_reference_test_inputs_through_all_three_kernels(emu_built, tmp_path):
    """the reference's own test inputs (tests/corpus.py: src/test_compat.zig:25-56, src/test.zig, test_lz4hc.zig, test_lz4f.zig
    inputs, the F8 hazard input of SURVEY F8) through K1 (as one block and in 64 KiB blocks, accelerations 1 and 3, both
    table widths), K2 (oracle streams of fast mode and HC 9, all three tiers) and K3 (levels 3 / 6 / 9, both walks)"""
    from corpus import block_cases, compat_cases, f8_hazard_input
    files = []
    for i, (name, data) in enumerate(compat_cases() + block_cases() + [("f8", f8_hazard_input())]):
        f = tmp_path / ("%03d_%s.bin" % (i, name))
        f.write_bytes(data)
        files.append("@" + str(f))
    for binary in ("emu_k1", "emu_k2", "emu_k3"):
        out = run(binary, *files)
        assert "files:" in out
