"""SECOND SOURCE (test infrastructure only) — an independent Python restatement of the reference's hot path.

Written from the Zig text of /root/reference/src/{lz4,lz4hc,lz4f}.zig, *not* from oracle/*.c: different
language, different author pass, different structure (Python ints and bytes, no pointer arithmetic).  The
only job of this file is to pin the C oracle (and through it the CUDA path) to the reference's algorithm by
a second, independently written derivation — the reference holds no golden vectors and there is
no zig toolchain to run it.  tests/second_source/make_vectors.py runs it over a fixed corpus and commits the
result hashes; tests/test_second_source.py checks oracle == vectors (CPU) and CUDA == vectors (GPU).

Scope: lz4.compressFast / decompressSafe(+UsingDict) (src/lz4.zig:89-259,283-519,960-964),
lz4hc.compressHC levels 3..9 (src/lz4hc.zig:129-131,170-264,308-386,491-681,976-1064,1394-1489),
lz4f.compressFrame / decompressFrame / header codec (src/lz4f.zig:138-638), XXH32 (Zig std, standard XXH32).

Two places where this restatement cannot "do what the reference does" because the reference's behaviour is
undefined there; both are written down in DESIGN.md §2 and are the same choices the oracle makes:
  * src/lz4hc.zig:636 `matchIndex - 1` on a u32 with matchIndex == 0: treated as "no pattern
    candidate" (the wrapped index fails `>= lowestMatchIndex`/bounds in every sane reading); counted in
    `F8_HITS`.
  * src/lz4hc.zig:1037 final-literal head-room ignores the length bytes: unreachable with a
    compressBound-sized destination (the only capacity the vectors use).
"""
import struct

M32 = 0xFFFFFFFF

MINMATCH = 4
LASTLITERALS = 5
MFLIMIT = 12
RUN_MASK = 15
ML_MASK = 15
DIST_MAX = 65535
MAX_INPUT = 0x7E000000
ACCEL_MAX = 65537
GOLDEN = 2654435761


class Lz4Error(Exception):
    """kind is the reference's error member name (lz4.Error src/lz4.zig:48-55, lz4f.Error src/lz4f.zig:31-55)."""

    def __init__(self, kind):
        Exception.__init__(self, kind)
        self.kind = kind


def u32(b, i):
    return b[i] | (b[i + 1] << 8) | (b[i + 2] << 16) | (b[i + 3] << 24)


def compress_bound(n):                                            # src/lz4.zig:80-83
    return 0 if n > MAX_INPUT else n + n // 255 + 16


# ------------------------------------------------------------------------------------------------ XXH32
P1, P2, P3, P4, P5 = 2654435761, 2246822519, 3266489917, 668265263, 374761393


def _rotl(x, r):
    return ((x << r) | (x >> (32 - r))) & M32


def xxh32(data, seed=0):
    n = len(data)
    i = 0
    if n >= 16:
        acc = [(seed + P1 + P2) & M32, (seed + P2) & M32, seed & M32, (seed - P1) & M32]
        while i + 16 <= n:
            for k in range(4):
                acc[k] = (_rotl((acc[k] + u32(data, i + 4 * k) * P2) & M32, 13) * P1) & M32
            i += 16
        h = (_rotl(acc[0], 1) + _rotl(acc[1], 7) + _rotl(acc[2], 12) + _rotl(acc[3], 18)) & M32
    else:
        h = (seed + P5) & M32
    h = (h + n) & M32
    while i + 4 <= n:
        h = (_rotl((h + u32(data, i) * P3) & M32, 17) * P4) & M32
        i += 4
    while i < n:
        h = (_rotl((h + data[i] * P5) & M32, 11) * P1) & M32
        i += 1
    h ^= h >> 15
    h = (h * P2) & M32
    h ^= h >> 13
    h = (h * P3) & M32
    h ^= h >> 16
    return h


# ------------------------------------------------------------------------------------------------ fast path
def _len_bytes(v):
    """The 255-run that follows a saturated nibble: value v = length - 15."""
    out = bytearray()
    while v >= 255:
        out.append(255)
        v -= 255
    out.append(v)
    return out


def _literal_only(src, start, out, cap):
    """compressAsLiterals / finishCompression, src/lz4.zig:449-519 (the capacity tests there all reduce to
    "does the finished block fit", every step only moves forward)."""
    ll = len(src) - start
    if ll == 0:
        return bytes(out)
    if ll >= RUN_MASK:
        out.append(RUN_MASK << 4)
        out += _len_bytes(ll - RUN_MASK)
    else:
        out.append(ll << 4)
    out += src[start:]
    if len(out) > cap:
        raise Lz4Error("OutputTooSmall")
    return bytes(out)


def compress_fast(src, accel=1, cap=None):
    """lz4.compressFast, src/lz4.zig:292-447."""
    n = len(src)
    if n > MAX_INPUT:
        raise Lz4Error("InputTooLarge")
    if n == 0:
        return b""
    if cap is None:
        cap = compress_bound(n)
    out = bytearray()
    if n < MFLIMIT + 1:
        if cap < 1:
            raise Lz4Error("OutputTooSmall")
        return _literal_only(src, 0, out, cap)
    table = [0] * 4096
    lim = n - MFLIMIT
    mlim = n - LASTLITERALS
    a = min(max(accel, 1), ACCEL_MAX)
    ip = 1
    anchor = 0
    while ip < lim:
        step = a
        nb = a
        fwd = ip
        while True:
            ip = fwd
            fwd += step
            step = nb >> 6
            nb += 1
            if fwd > lim:
                return _literal_only(src, anchor, out, cap)
            s = u32(src, ip)
            h = ((s * GOLDEN) & M32) >> 20
            m = table[h]
            ok = m > 0 and m < ip and m + DIST_MAX >= ip and u32(src, m) == s
            table[h] = ip
            if ok:
                break
        ll = ip - anchor
        tok = len(out)
        out.append(0)
        if len(out) >= cap:                                       # :364
            raise Lz4Error("OutputTooSmall")
        if ll >= RUN_MASK:
            out[tok] = RUN_MASK << 4
            out += _len_bytes(ll - RUN_MASK)
        else:
            out[tok] = ll << 4
        out += src[anchor:ip]
        off = ip - m
        out.append(off & 255)
        out.append(off >> 8)
        ip += MINMATCH
        m += MINMATCH
        ml = 0
        while ip < mlim and src[ip] == src[m]:
            ip += 1
            m += 1
            ml += 1
        if ml >= ML_MASK:
            out[tok] |= ML_MASK
            out += _len_bytes(ml - ML_MASK)
        else:
            out[tok] |= ml
        if len(out) > cap:                                        # every check :365-429 is "would this byte fit"
            raise Lz4Error("OutputTooSmall")
        anchor = ip
        if ip < lim:
            s = u32(src, ip)
            table[((s * GOLDEN) & M32) >> 20] = ip
            ip += 1
    return _literal_only(src, anchor, out, cap)


def decompress_safe(src, cap, dictionary=None):
    """lz4.decompressSafe / decompressSafeUsingDict, src/lz4.zig:89-259,960-964.  Returns the decoded bytes."""
    n = len(src)
    if n == 0 or cap == 0:
        return b""
    out = bytearray()
    ip = 0
    while True:
        if ip >= n:
            break
        tok = src[ip]
        ip += 1
        ll = tok >> 4
        if ll == RUN_MASK:
            while True:
                if ip >= n:
                    raise Lz4Error("CorruptedData")
                s = src[ip]
                ip += 1
                ll += s
                if s != 255:
                    break
        if ll > 0:
            if ip + ll > n:
                raise Lz4Error("CorruptedData")
            if len(out) + ll > cap:
                raise Lz4Error("OutputTooSmall")
            out += src[ip:ip + ll]
            ip += ll
        if ip >= n:
            break
        if ip + 2 > n:
            raise Lz4Error("CorruptedData")
        off = src[ip] | (src[ip + 1] << 8)
        ip += 2
        if off == 0:
            raise Lz4Error("CorruptedData")
        ml = tok & ML_MASK
        if ml == ML_MASK:
            while True:
                if ip >= n:
                    raise Lz4Error("CorruptedData")
                s = src[ip]
                ip += 1
                ml += s
                if s != 255:
                    break
        ml += MINMATCH
        op = len(out)
        if op + ml > cap:
            raise Lz4Error("OutputTooSmall")
        if off > op:
            if dictionary is None:
                raise Lz4Error("CorruptedData")
            if off > op + len(dictionary):
                raise Lz4Error("CorruptedData")
            back = off - op
            dpos = len(dictionary) - back
            if ml <= back:
                out += dictionary[dpos:dpos + ml]
            else:
                out += dictionary[dpos:]
                for i in range(ml - back):                         # rest continues at dst[0..], may overlap itself
                    out.append(out[i])
        else:
            start = op - off
            for i in range(ml):
                out.append(out[start + i])
    return bytes(out)


# ------------------------------------------------------------------------------------------------ HC levels 3..9
NB_SEARCHES = {3: 4, 4: 8, 5: 16, 6: 32, 7: 64, 8: 128, 9: 256}      # src/lz4hc.zig:72-86
F8_HITS = 0


def _common(src, a, b, limit):
    """lz4Count (src/lz4hc.zig:234-264): equal bytes of src[a..] and src[b..], a stops at `limit`."""
    k = 0
    room = limit - a
    while k + 32 <= room and src[a + k:a + k + 32] == src[b + k:b + k + 32]:
        k += 32
    while k < room and src[a + k] == src[b + k]:
        k += 1
    return k if room > 0 else 0


def _count_pattern(src, start, end, pattern):
    """countPattern (src/lz4hc.zig:170-199).  Only ever called with a pattern of four equal bytes
    (isRepetitivePattern :225-228), so the 8-byte stride and the rotating tail byte both compare against
    that byte."""
    b = pattern & 255
    p = start
    while p < end and src[p] == b:
        p += 1
    return p - start


def _reverse_count_pattern(src, start, low, pattern):
    """reverseCountPattern (src/lz4hc.zig:202-222), same remark."""
    b = pattern & 255
    p = start
    while p > low and src[p - 1] == b:
        p -= 1
    return start - p


def _hc_emit(out, src, anchor, ip, ml, off, cap):
    """encodeSequence with limitedOutput (src/lz4hc.zig:308-386)."""
    ll = ip - anchor
    if len(out) + ll // 255 + ll + 8 > cap:                       # :320-325
        raise Lz4Error("OutputTooSmall")
    tok = len(out)
    out.append(0)
    if ll >= RUN_MASK:
        out[tok] = RUN_MASK << 4
        out += _len_bytes(ll - RUN_MASK)
    else:
        out[tok] = ll << 4
    out += src[anchor:ip]
    out.append(off & 255)
    out.append(off >> 8)
    code = ml - MINMATCH
    if len(out) + code // 255 + 6 > cap:                           # :355-359
        raise Lz4Error("OutputTooSmall")
    if code >= ML_MASK:
        out[tok] += ML_MASK
        out += _len_bytes(code - ML_MASK)                          # :364-376 writes the same bytes two at a time
    else:
        out[tok] += code


def compress_hc(src, level=9, cap=None):
    """lz4hc.compressHC (src/lz4hc.zig:1440-1489) for the hash-chain levels, greedy."""
    global F8_HITS
    n = len(src)
    if n > MAX_INPUT:
        raise Lz4Error("InputTooLarge")
    if n == 0:
        return b""
    if level < 2:                                                 # :1445
        level = 9
    if level > 12:
        level = 12
    if level not in NB_SEARCHES:
        raise Lz4Error("UnsupportedLevel")                         # levels 2, 10-12: other strategies, out of scope
    if cap is None:
        cap = compress_bound(n)
    if cap == 0:
        raise Lz4Error("OutputTooSmall")                           # :1461
    attempts_max = NB_SEARCHES[level]
    pattern_analysis = attempts_max > 128                          # :983
    out = bytearray()
    if n < MFLIMIT + 1:                                            # :995-998 -> encodeLiterals :1394-1425
        if cap < n + 1 + n // 255:
            raise Lz4Error("OutputTooSmall")
        if n >= RUN_MASK:
            out.append(RUN_MASK << 4)
            out += _len_bytes(n - RUN_MASK)
        else:
            out.append(n << 4)
        out += src
        return bytes(out)

    head = [0] * 32768                                            # hashTable, 0 = empty
    chain = [0] * 65536                                           # chainTable (one-shot path zero-fills it, :405-408)
    next_insert = 0
    mflimit = n - MFLIMIT
    mlimit = n - LASTLITERALS
    ip = 0
    anchor = 0
    while ip <= mflimit:
        # insertHC :491-510 — every position below ip, in order
        while next_insert < ip:
            h = ((u32(src, next_insert) * GOLDEN) & M32) >> 17
            prev = head[h]
            d = DIST_MAX + 1 if prev > next_insert else next_insert - prev
            chain[next_insert & 0xFFFF] = min(d, DIST_MAX)
            head[h] = next_insert
            next_insert += 1
        # insertAndGetWiderMatch :538-681 with iLowLimit == ip, longest = 3
        lowest = 0 if ip < 65536 else ip - DIST_MAX               # :552-553 with lowLimit == 0
        pat = u32(src, ip)
        best_len, best_off = MINMATCH - 1, 0
        m = head[((pat * GOLDEN) & M32) >> 17]
        if m != 0:
            left = attempts_max
            while m > 0 and left > 0:
                if m > ip or ip - m > DIST_MAX:
                    break
                left -= 1
                if m >= lowest and u32(src, m) == pat:
                    length = MINMATCH + _common(src, ip + MINMATCH, m + MINMATCH, mlimit)
                    if length > best_len:
                        best_len, best_off = length, ip - m
                        if length > attempts_max:
                            break
                d = chain[m & 0xFFFF]
                if d == 0 or d > m:
                    break
                m -= d
            if pattern_analysis and chain[m & 0xFFFF] == 1:       # :626-631 (result.len > 0 always holds)
                if (pat & 0xFFFF) == (pat >> 16) and (pat & 0xFF) == (pat >> 24):
                    src_run = _count_pattern(src, ip + 4, mlimit, pat) + 4
                    if m == 0:
                        F8_HITS += 1                               # u32 underflow in the reference (:636), see header
                    else:
                        cand = m - 1
                        if cand >= lowest and u32(src, cand) == pat:
                            fwd = _count_pattern(src, cand + 4, mlimit, pat) + 4
                            back = _reverse_count_pattern(src, cand, 0, pat)
                            back = cand - max(cand - back, lowest)
                            seg = back + fwd
                            cap_ml = min(seg, src_run)
                            if seg >= src_run and fwd <= src_run:
                                new_m = cand + fwd - src_run
                            else:
                                new_m = cand - back
                            if cap_ml > best_len and ip - new_m <= DIST_MAX:
                                best_len, best_off = cap_ml, ip - new_m
        if best_len < MINMATCH or best_off == 0:
            ip += 1
            continue
        _hc_emit(out, src, anchor, ip, best_len, best_off, cap)
        ip += best_len
        anchor = ip
    last = n - anchor
    if last > 0:
        if len(out) + last + 1 > cap:                              # :1037
            raise Lz4Error("OutputTooSmall")
        if last >= RUN_MASK:
            out.append(RUN_MASK << 4)
            out += _len_bytes(last - RUN_MASK)
        else:
            out.append(last << 4)
        out += src[anchor:]
    return bytes(out)


# ------------------------------------------------------------------------------------------------ frames
BLOCK_BYTES = {0: 65536, 4: 65536, 5: 262144, 6: 1 << 20, 7: 4 << 20}   # src/lz4f.zig:73-80
MAGIC = 0x184D2204


class Prefs(object):
    """lz4f.Preferences + FrameInfo, src/lz4f.zig:106-122 (fields that are ever read)."""

    def __init__(self, block_size_id=0, block_mode=0, content_checksum=0, block_checksum=0, content_size=0,
                 dict_id=0, compression_level=0):
        self.block_size_id = block_size_id
        self.block_mode = block_mode            # 0 linked (default!), 1 independent — flag only
        self.content_checksum = content_checksum
        self.block_checksum = block_checksum
        self.content_size = content_size
        self.dict_id = dict_id
        self.compression_level = compression_level


def frame_bound(n, p):                                            # src/lz4f.zig:274-301
    bs = BLOCK_BYTES[p.block_size_id]
    nblocks = (n + bs - 1) // bs
    per = 4 + compress_bound(bs) + (4 if p.block_checksum else 0)
    return 19 + nblocks * per + 4 + (4 if p.content_checksum else 0)


def frame_header(p):                                              # src/lz4f.zig:304-351,152-184,224-232,138-141
    flg = 0x40
    if p.block_mode == 1:
        flg |= 0x20
    if p.block_checksum:
        flg |= 0x10
    if p.content_size != 0:
        flg |= 0x08
    if p.content_checksum:
        flg |= 0x04
    if p.dict_id != 0:
        flg |= 0x01
    bd = {0: 4, 4: 4, 5: 5, 6: 6, 7: 7}[p.block_size_id] << 4
    desc = bytearray([flg, bd])
    if p.content_size != 0:
        desc += struct.pack("<Q", p.content_size)
    if p.dict_id != 0:
        desc += struct.pack("<I", p.dict_id)
    return struct.pack("<I", MAGIC) + bytes(desc) + bytes([(xxh32(bytes(desc)) >> 8) & 255])


def compress_frame(src, p, cap=None, block_cache=None):
    """lz4f.compressFrame, src/lz4f.zig:354-446.  block_cache: optional dict reused between calls by the
    vector generator (same block bytes + level -> same compressed block; pure memoisation)."""
    n = len(src)
    if cap is not None and cap < frame_bound(n, p):
        raise Lz4Error("DstMaxSizeTooSmall")
    out = bytearray(frame_header(p))
    bs = BLOCK_BYTES[p.block_size_id]
    pos = 0
    while pos < n:
        block = src[pos:pos + bs]
        key = (p.compression_level, block)
        comp = block_cache.get(key) if block_cache is not None else None
        if comp is None:
            if p.compression_level > 0:
                comp = compress_hc(block, p.compression_level)
            else:
                comp = compress_fast(block, 1)
            if block_cache is not None:
                block_cache[key] = comp
        if len(comp) >= len(block):
            out += struct.pack("<I", len(block) | 0x80000000)
            stored = block
        else:
            out += struct.pack("<I", len(comp))
            stored = comp
        out += stored
        if p.block_checksum:
            out += struct.pack("<I", xxh32(stored))
        pos += len(block)
    out += b"\0\0\0\0"
    if p.content_checksum:
        out += struct.pack("<I", xxh32(src))
    return bytes(out)


def header_size(src):                                             # src/lz4f.zig:451-480
    if len(src) < 5:
        raise Lz4Error("FrameHeaderIncomplete")
    magic = u32(src, 0)
    if magic != MAGIC:
        if (magic & 0xFFFFFFF0) == 0x184D2A50:
            return 8
        raise Lz4Error("FrameTypeUnknown")
    return 7 + (8 if src[4] & 0x08 else 0) + (4 if src[4] & 0x01 else 0)


def parse_header(src):                                            # src/lz4f.zig:483-538,187-221,235-249
    if len(src) < 7:
        raise Lz4Error("FrameHeaderIncomplete")
    if u32(src, 0) != MAGIC:
        raise Lz4Error("FrameTypeUnknown")
    flg = src[4]
    if (flg >> 6) & 3 != 1:
        raise Lz4Error("HeaderVersionWrong")
    if flg & 0x02:
        raise Lz4Error("ReservedFlagSet")
    bd = src[5]
    if bd & 0x8F:
        raise Lz4Error("ReservedFlagSet")
    sid = (bd >> 4) & 7
    if sid not in (0, 4, 5, 6, 7):
        raise Lz4Error("MaxBlockSizeInvalid")
    pos = 6
    info = {"block_size_id": 4 if sid == 0 else sid, "block_mode": (flg >> 5) & 1, "block_checksum": (flg >> 4) & 1,
            "content_checksum": (flg >> 2) & 1, "content_size": 0, "dict_id": 0}
    if flg & 0x08:
        if len(src) < pos + 8:
            raise Lz4Error("FrameHeaderIncomplete")
        info["content_size"] = struct.unpack_from("<Q", src, pos)[0]
        pos += 8
    if flg & 0x01:
        if len(src) < pos + 4:
            raise Lz4Error("FrameHeaderIncomplete")
        info["dict_id"] = u32(src, pos)
        pos += 4
    if len(src) < pos + 1:
        raise Lz4Error("FrameHeaderIncomplete")
    if src[pos] != (xxh32(bytes(src[4:pos])) >> 8) & 255:
        raise Lz4Error("HeaderChecksumInvalid")
    return info, pos + 1


def decompress_frame(src, cap):
    """lz4f.decompressFrame, src/lz4f.zig:541-638: sequential dstPos, every block decoded into what is left of dst."""
    info, pos = parse_header(src)
    out = bytearray()
    n = len(src)
    while pos < n:
        if pos + 4 > n:
            raise Lz4Error("FrameSizeWrong")
        hdr = u32(src, pos)
        pos += 4
        if hdr == 0:
            break
        raw = bool(hdr & 0x80000000)
        size = hdr & 0x7FFFFFFF
        if pos + size > n:
            raise Lz4Error("FrameSizeWrong")
        data = src[pos:pos + size]
        pos += size
        if info["block_checksum"]:
            if pos + 4 > n:
                raise Lz4Error("FrameSizeWrong")
            if u32(src, pos) != xxh32(data):
                raise Lz4Error("BlockChecksumInvalid")
            pos += 4
        if raw:
            if len(out) + size > cap:
                raise Lz4Error("DstMaxSizeTooSmall")
            out += data
        else:
            try:
                out += decompress_safe(data, cap - len(out))
            except Lz4Error:
                raise Lz4Error("DecompressionFailed")
    if info["content_checksum"]:
        if pos + 4 > n:
            raise Lz4Error("FrameSizeWrong")
        if u32(src, pos) != xxh32(bytes(out)):
            raise Lz4Error("ContentChecksumInvalid")
        pos += 4
    return bytes(out)
