"""One lz4f frame over several GPUs: one process per GPU, the frame sharded by contiguous block ranges.

Blocks are independent in both directions (reference src/lz4f.zig:379-430 compresses them one by one, :563-614 decodes
them one by one), so rank k of G owns blocks [k*B/G, (k+1)*B/G):

  compress    each rank encodes its range into a *body* (the block records exactly as they sit in the frame); ONE
              all_gather of a single 64-bit size gives every rank the frame layout
                  header (rank 0) | body_0 | ... | body_{G-1} | end mark | content checksum (last rank);
              the content checksum is one serial XXH32 chain, handed rank k -> k+1 as a 40-byte state;
              bodies are gathered to one rank with point-to-point sends of their exact sizes, or left sharded.
  decompress  the block index of the frame (the parallel replacement of the serial header chain) gives the
              byte position of every block record; the frame is cut at G+1 of them, every rank decodes its cut; the
              content checksum is verified by the same state hand-off over the decoded ranges.

The data path has no collective besides that size exchange.  The transport is `torch.distributed` (NCCL over NVLink on
the GPU box, gloo in the CPU tests of the host logic); the codec is an *engine* object — `CudaEngine` below drives
libb2lz4.so and needs a CUDA device (there is no CPU engine in the product; tests/ supplies one built on the oracle to
exercise this file's partitioning, layout and hand-off logic with world_size 2 on CPU)."""
import ctypes as C
from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import _native, lz4f
from ._native import B2Error, XxhState

ERR_FRAME_SIZE_WRONG = 113            # lz4f.Error.FrameSizeWrong, reference src/lz4f.zig:31-55
ERR_CONTENT_CHECKSUM_INVALID = 117    # lz4f.Error.ContentChecksumInvalid


# ------------------------------------------------------------------ partition and layout (pure host logic)
def block_range(rank, world, nblocks):
    """blocks [lo, hi) of rank `rank`: floor(k*B/G) .. floor((k+1)*B/G)"""
    return rank * nblocks // world, (rank + 1) * nblocks // world


def byte_range(rank, world, n, block_size):
    """raw bytes [lo, hi) of rank `rank` for an input of n bytes cut into block_size blocks"""
    nblocks = (n + block_size - 1) // block_size
    lo, hi = block_range(rank, world, nblocks)
    return min(lo * block_size, n), min(hi * block_size, n)


@dataclass
class FrameLayout:
    header_size: int
    body_sizes: list        # per rank
    body_offsets: list      # per rank, position of body_k in the frame
    end_mark_pos: int
    total: int              # frame size including end mark and content checksum


def frame_layout(header_size, body_sizes, content_checksum):
    offs, pos = [], header_size
    for s in body_sizes:
        offs.append(pos)
        pos += s
    return FrameLayout(header_size, list(body_sizes), offs, pos, pos + 4 + (4 if content_checksum else 0))


def _rank_world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def exchange_sizes(size, device, group=None):
    """the one collective of the compress path: all_gather of one int64 per rank"""
    rank, world = _rank_world(group)
    if world == 1:
        return [int(size)]
    mine = torch.tensor([int(size)], dtype=torch.int64, device=device)
    every = torch.zeros(world, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(every, mine, group=group)
    return [int(x) for x in every.tolist()]


def _bytes_to_tensor(b, device):
    return torch.frombuffer(bytearray(b), dtype=torch.uint8).to(device)


def _tensor_to_bytes(t):
    return t.cpu().numpy().tobytes()


def _chain_checksum(engine, data, group):
    """content checksum over the ranks' data in rank order: receive the running state from rank-1, absorb this
    rank's bytes, pass it on; the last rank finishes it.  Returns the checksum on the last rank, None elsewhere.
    The state is `engine.STATE_BYTES` of plain data (four lanes, the <16-byte tail, the byte count)."""
    rank, world = _rank_world(group)
    if rank == 0:
        state = engine.checksum_init()
    else:
        buf = torch.zeros(engine.STATE_BYTES, dtype=torch.uint8, device=engine.device)
        dist.recv(buf, src=rank - 1, group=group)
        state = _tensor_to_bytes(buf)
    state = engine.checksum_update(state, data)
    if rank + 1 < world:
        dist.send(_bytes_to_tensor(state, engine.device), dst=rank + 1, group=group)
        return None
    return engine.checksum_final(state)


# ------------------------------------------------------------------ compress
def compress_frame_sharded(engine, shard, prefs, group=None, gather_to=0):
    """`shard`: uint8 tensor with this rank's byte_range() of the input (prefs.content_size, if set, is the total).
    Returns (frame, layout, body): `frame` is the whole frame on rank `gather_to` (None elsewhere, or everywhere when
    gather_to is None: the frame stays sharded as `body` at layout.body_offsets[rank])."""
    rank, world = _rank_world(group)
    body = engine.compress_body(shard, prefs)
    sizes = exchange_sizes(body.numel(), engine.device, group)
    header = engine.header(prefs)
    cc = prefs is not None and prefs.content_checksum == 1
    layout = frame_layout(len(header), sizes, cc)
    csum = _chain_checksum(engine, shard, group) if cc else None
    if gather_to is None:
        return None, layout, body
    last = world - 1
    trailer = None
    if rank == last:
        trailer = b"\0\0\0\0" + (csum.to_bytes(4, "little") if cc else b"")
    frame = None
    if rank == gather_to:
        frame = engine.empty(layout.total)
        frame[:len(header)] = _bytes_to_tensor(header, engine.device)
        for r in range(world):
            dstv = frame[layout.body_offsets[r]:layout.body_offsets[r] + sizes[r]]
            if r == rank:
                dstv.copy_(body)
            elif sizes[r]:
                dist.recv(dstv, src=r, group=group)
        tail = frame[layout.end_mark_pos:]
        if rank == last:
            tail.copy_(_bytes_to_tensor(trailer, engine.device))
        else:
            dist.recv(tail, src=last, group=group)
    else:
        if body.numel():
            dist.send(body.contiguous(), dst=gather_to, group=group)
        if rank == last:
            dist.send(_bytes_to_tensor(trailer, engine.device), dst=gather_to, group=group)
    return frame, layout, body


# ------------------------------------------------------------------ decompress
_META = 10  # nblocks, block_size, block_checksum, content_checksum, stored checksum, terminal, header_size, end_pos, status, spare


def decompress_frame_sharded(engine, frame, group=None, src=0, gather_to=None):
    """`frame`: uint8 tensor holding the whole frame on rank `src` (ignored elsewhere).  Every rank decodes its block
    range and returns (out, (block_lo, block_hi), total): `out` is this rank's decoded bytes, or the whole output on rank
    `gather_to` when that is given (None on the other ranks).  Raises the reference's lz4f error on every rank."""
    rank, world = _rank_world(group)
    meta = torch.zeros(_META + world + 1, dtype=torch.int64, device=engine.device)
    if rank == src:
        status = 0
        try:
            idx = engine.index(frame)
        except B2Error as e:
            status, idx = e.code, None
        if idx is not None:
            nb = idx["nblocks"]
            if idx["terminal"] == 2:
                status = ERR_FRAME_SIZE_WRONG                               # truncated record chain, src/lz4f.zig:565,582,591
            stored = 0
            if status == 0 and idx["content_checksum"]:
                if idx["end_pos"] + 4 > frame.numel():
                    status = ERR_FRAME_SIZE_WRONG                           # :626
                else:
                    stored = int.from_bytes(_tensor_to_bytes(frame[idx["end_pos"]:idx["end_pos"] + 4]), "little")
            # record k starts 4 bytes before its payload; the chain ends where the end mark starts (or at the end of
            # the input when there is none: terminal 1)
            chain_end = idx["end_pos"] - (4 if idx["terminal"] == 0 else 0)
            cuts = []
            for r in range(world + 1):
                b = r * nb // world
                cuts.append(int(idx["off"][b]) - 4 if b < nb else chain_end)
            vals = [nb, idx["block_size"], idx["block_checksum"], idx["content_checksum"], stored, idx["terminal"],
                    idx["header_size"], idx["end_pos"], status, 0] + cuts
        else:
            vals = [0] * (_META + world + 1)
            vals[8] = status
        meta.copy_(torch.tensor(vals, dtype=torch.int64).to(engine.device))
    if world > 1:
        dist.broadcast(meta, src=src, group=group)
    m = [int(x) for x in meta.tolist()]
    nb, block_size, bc, cc, stored, terminal, status = m[0], m[1], m[2], m[3], m[4], m[5], m[8]
    if status:
        raise B2Error(status)
    cuts = m[_META:]
    lo, hi = block_range(rank, world, nb)
    # the cut of every rank travels from `src` with point-to-point sends of its exact size
    if rank == src:
        for r in range(world):
            if r != rank and cuts[r + 1] > cuts[r]:
                dist.send(frame[cuts[r]:cuts[r + 1]].contiguous(), dst=r, group=group)
        body = frame[cuts[rank]:cuts[rank + 1]]
    else:
        body = engine.empty(cuts[rank + 1] - cuts[rank])
        if body.numel():
            dist.recv(body, src=src, group=group)
    err = 0
    out = engine.empty(0)
    try:
        if hi > lo:
            out = engine.decode_body(body, (hi - lo) * block_size, block_size, bool(bc))
    except B2Error as e:
        err = e.code
    # the first failing block decides the error, as in the serial loop: lowest rank with a failure wins
    errs = exchange_sizes(err, engine.device, group)
    first = next((e for e in errs if e), 0)
    if first:
        raise B2Error(first)
    if cc:
        got = _chain_checksum(engine, out, group)
        verdict = torch.zeros(1, dtype=torch.int64, device=engine.device)
        if rank == world - 1:
            verdict[0] = 0 if got == stored else ERR_CONTENT_CHECKSUM_INVALID
        if world > 1:
            dist.broadcast(verdict, src=world - 1, group=group)
        if int(verdict.item()):
            raise B2Error(int(verdict.item()))
    sizes = exchange_sizes(out.numel(), engine.device, group)
    total = sum(sizes)
    if gather_to is None:
        return out, (lo, hi), total
    if rank == gather_to:
        whole = engine.empty(total)
        pos = 0
        for r in range(world):
            dstv = whole[pos:pos + sizes[r]]
            if r == rank:
                dstv.copy_(out)
            elif sizes[r]:
                dist.recv(dstv, src=r, group=group)
            pos += sizes[r]
        return whole, (lo, hi), total
    if out.numel():
        dist.send(out.contiguous(), dst=gather_to, group=group)
    return None, (lo, hi), total


# ------------------------------------------------------------------ the product engine
class CudaEngine:
    """The codec of one rank: a b2lz4 Context on this rank's GPU.  Tensors are CUDA uint8; every call goes through
    the C-ABI (`b2lz4f_compress_blocks_dev`, `b2lz4f_index_frame_dev`, `b2lz4f_decompress_blocks_dev`,
    `b2lz4_xxh32_state_*`).  Raises without a CUDA device: there is no CPU fallback."""

    def __init__(self, device_index):
        if not torch.cuda.is_available():
            raise RuntimeError("CudaEngine needs a CUDA device; the product has no CPU codec")
        self.device = torch.device("cuda", device_index)
        self.ctx = _native.Context(device_index)
        self._body = None
        self._idx = None

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def empty(self, n):
        return torch.empty(int(n), dtype=torch.uint8, device=self.device)

    def header(self, prefs):
        return lz4f.writeFrameHeader(prefs)

    def compress_body(self, shard, prefs):
        n = shard.numel()
        cap = lz4f.compressFrameBound(n, prefs)
        if self._body is None or self._body.numel() < cap + 64:
            self._body = self.empty(cap + 64)
        size = self.ctx.compress_blocks_dev(shard.data_ptr(), n, self._body.data_ptr(), cap, prefs, self._stream())
        return self._body[:size]

    STATE_BYTES = C.sizeof(XxhState)   # struct b2lz4_xxh32_state, 40 bytes

    def checksum_init(self):
        st = XxhState()
        _native.lib().b2lz4_xxh32_state_init(C.byref(st), 0)
        return bytes(st)

    def checksum_update(self, state, data):
        st = XxhState.from_buffer_copy(state)
        if data.numel():
            self.ctx.xxh32_state_update_dev(st, data.data_ptr(), data.numel(), self._stream())
        return bytes(st)

    def checksum_final(self, state):
        st = XxhState.from_buffer_copy(state)
        return _native.lib().b2lz4_xxh32_state_final(C.byref(st))

    def index(self, frame):
        n = frame.numel()
        info = self.ctx.index_frame_dev(frame.data_ptr(), n, 0, 0, 0, self._stream())
        nb = int(info.nblocks)
        off = torch.empty(max(1, nb), dtype=torch.int64, device=self.device)
        hdr = torch.empty(max(1, nb), dtype=torch.int32, device=self.device)
        if nb:
            info = self.ctx.index_frame_dev(frame.data_ptr(), n, off.data_ptr(), hdr.data_ptr(), nb, self._stream())
        torch.cuda.synchronize(self.device)
        return {"nblocks": nb, "end_pos": int(info.end_pos), "terminal": int(info.terminal), "header_size": int(info.header_size),
                "block_size": int(info.block_size), "block_checksum": int(info.block_checksum),
                "content_checksum": int(info.content_checksum), "content_size": int(info.content_size), "off": off[:nb].cpu()}

    def decode_body(self, body, cap, block_size, block_checksum):
        out = self.empty(cap + 64)
        m = self.ctx.decompress_blocks_dev(body.data_ptr(), body.numel(), out.data_ptr(), cap, block_size, block_checksum,
                                           self._stream())
        return out[:m]
