"""A CPU codec engine for zig_lz4_b200.sharded built on the oracle — TEST INFRASTRUCTURE: it lets the multi-process
host logic (partition, size exchange, frame layout, checksum hand-off, cuts, gathers) run under gloo with world_size 2
in a container without a GPU.  The product engine is sharded.CudaEngine."""
import ctypes as C

import torch

import b2oracle as o
from zig_lz4_b200._native import B2Error


def _oprefs(prefs):
    if prefs is None:
        return o.make_prefs()
    return o.make_prefs(prefs.block_size_id, prefs.block_mode, prefs.content_checksum, prefs.content_size, prefs.dict_id,
                        prefs.block_checksum, prefs.compression_level)


def _block_size(block_size_id):
    return {0: 65536, 4: 65536, 5: 262144, 6: 1 << 20, 7: 4 << 20}[block_size_id]


class OracleEngine:
    STATE_BYTES = C.sizeof(o.XxhState)
    device = torch.device("cpu")

    def empty(self, n):
        return torch.zeros(int(n), dtype=torch.uint8)

    def header(self, prefs):
        buf = (C.c_uint8 * 32)()
        out = C.c_size_t(0)
        p = _oprefs(prefs)
        rc = o.lib().b2o_write_frame_header(buf, 32, C.byref(p), C.byref(out))
        assert rc == 0
        return bytes(buf[:out.value])

    def compress_body(self, shard, prefs):
        """records of the shard's blocks = the oracle's frame of the shard without header, end mark and content checksum"""
        p = _oprefs(prefs)
        p.content_checksum = 0
        p.content_size = 0
        data = shard.numpy().tobytes()
        frame = o.compress_frame(data, p)
        hs = o.header_size(frame[:19])
        return torch.frombuffer(bytearray(frame[hs:len(frame) - 4]), dtype=torch.uint8) if len(frame) - 4 > hs else self.empty(0)

    def checksum_init(self):
        return bytes(o.xxh32_state_init(0))

    def checksum_update(self, state, data):
        st = o.XxhState.from_buffer_copy(state)
        o.xxh32_state_update(st, data.numpy().tobytes())
        return bytes(st)

    def checksum_final(self, state):
        return o.xxh32_state_final(o.XxhState.from_buffer_copy(state))

    def index(self, frame):
        """the serial header chain, reference src/lz4f.zig:563-589"""
        raw = frame.numpy().tobytes()
        info = o.make_prefs()
        size = C.c_size_t(0)
        hb = raw[:19]
        rc = o.lib().b2o_parse_frame_header(C.cast(C.c_char_p(hb), C.c_void_p), len(hb), C.byref(info), C.byref(size))
        if rc:
            raise B2Error(rc)
        bs, bc = _block_size(info.block_size_id), info.block_checksum
        pos, n, off, terminal = size.value, len(raw), [], 1
        while pos < n:
            if pos + 4 > n:
                terminal = 2
                break
            h = int.from_bytes(raw[pos:pos + 4], "little")
            pos += 4
            if h == 0:
                terminal = 0
                break
            sz = h & 0x7FFFFFFF
            if pos + sz > n or (bc and pos + sz + 4 > n):
                terminal = 2
                break
            off.append(pos)
            pos += sz + (4 if bc else 0)
        return {"nblocks": len(off), "end_pos": pos, "terminal": terminal, "header_size": size.value, "block_size": bs,
                "block_checksum": bc, "content_checksum": info.content_checksum, "content_size": info.content_size,
                "off": torch.tensor(off, dtype=torch.int64)}

    def decode_body(self, body, cap, block_size, block_checksum):
        raw = body.numpy().tobytes()
        pos, out = 0, bytearray()
        while pos < len(raw):
            h = int.from_bytes(raw[pos:pos + 4], "little")
            pos += 4
            sz = h & 0x7FFFFFFF
            payload = raw[pos:pos + sz]
            pos += sz
            if block_checksum:
                if int.from_bytes(raw[pos:pos + 4], "little") != o.xxh32(payload):
                    raise B2Error(106)                                     # BlockChecksumInvalid
                pos += 4
            if h & 0x80000000:
                out += payload
            else:
                try:
                    out += o.decompress_safe(payload, cap - len(out))
                except o.OracleError:
                    raise B2Error(115)                                     # DecompressionFailed
        return torch.frombuffer(out, dtype=torch.uint8) if out else self.empty(0)
