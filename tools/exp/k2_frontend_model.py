"""CPU model of K2's chunked front end (k_decompress.cu: decode_block_fast) — used while developing without a GPU:
checks that the per-position delta table + halting walk + per-lane token decode visit exactly the sequences of a plain
serial parse, for every source alignment, and that everything the fast tier refuses is left to the exact tier."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]

CHUNK, MARGIN, SLOTS, SPAN = 256, 32, 544, 18


def serial_parse(src):
    """reference parse (src/lz4.zig:111-171): list of (tokpos, lit, LL, ML or None)"""
    n = len(src); ip = 0; out = []
    while ip < n:
        tp = ip; t = src[ip]; ip += 1
        LL = t >> 4
        if LL == 15:
            while True:
                s = src[ip]; ip += 1; LL += s
                if s != 255: break
        lit = ip; ip += LL
        if ip >= n:
            out.append((tp, lit, LL, None)); break
        ip += 2
        ML = t & 15
        if ML == 15:
            while True:
                s = src[ip]; ip += 1; ML += s
                if s != 255: break
        out.append((tp, lit, LL, ML + 4))
    return out


def model(src, align):
    """returns list of events: ('batch', [(tokpos, lit, LL, ML)...]) or ('exact', ip) ; align = address of src mod 8"""
    n = len(src)
    events = []
    ip = 0
    if n <= SPAN:
        return events, ip
    isafe = n - SPAN
    while ip < isafe:
        sh = (align + ip) & 7
        cb = ip - sh
        B = np.zeros(CHUNK + MARGIN, dtype=np.int64)
        for w in range(32 + MARGIN // 8):
            if cb + 8 * w < n:
                for j in range(8):
                    a = cb + 8 * w + j
                    B[8 * w + j] = src[a] if 0 <= a < n else 0xAA   # bytes outside the stream inside a mapped word: garbage
        D = np.zeros(SLOTS, dtype=np.int64)
        lim = isafe - cb
        for p in range(CHUNK):
            if p >= lim: continue
            t = int(B[p]); hi, lo = t >> 4, t & 15
            d = 3 + hi + (1 if lo == 15 else 0)
            if hi == 15: d += 1 + int(B[p + 1])
            D[p] = 2 * d
        p2 = 2 * sh; pos = []
        for k in range(32):
            pos.append(p2); p2 += int(D[p2 // 2])
        pos.append(p2)
        assert max(pos) // 2 < SLOTS
        live = [D[x // 2] != 0 for x in pos[:32]]
        k = sum(live)
        assert all(live[:k]) and not any(live[k:])
        fields = []
        bad = []
        for lane in range(32):
            if not live[lane]:
                fields.append(None); bad.append(False); continue
            mp = pos[lane] // 2
            t = int(B[mp]); q = mp + 1; LL = t >> 4; bd = False
            if LL == 15:
                x = int(B[q]); LL += x; q += 1; bd = x == 255
            lit = cb + q; q += LL + 2; ML = t & 15
            if ML == 15:
                staged = q < CHUNK + MARGIN
                y = int(B[q if staged else 0]); ML += y; q += 1
                bd = bd or (not staged) or y == 255
            bd = bd or cb + q > n
            fields.append((cb + mp, lit, LL, ML + 4)); bad.append(bd)
        cut = any(bad)
        if cut: k = bad.index(True)
        ipn = cb + pos[k] // 2
        if k:
            events.append(("batch", fields[:k]))
        ip = ipn
        if cut:
            events.append(("exact", ip))
            t = src[ip]; q = ip + 1; LL = t >> 4
            if LL == 15:
                while True:
                    s = src[q]; q += 1; LL += s
                    if s != 255: break
            q += LL
            if q >= n:
                return events, n
            q += 2
            if (t & 15) == 15:
                while True:
                    s = src[q]; q += 1
                    if s != 255: break
            ip = q
        continue
    return events, ip
    isafe = n - SPAN
    while ip < isafe:
        sh = (align + ip) & 7
        cb = ip - sh
        B = np.zeros(CHUNK + MARGIN, dtype=np.int64)
        for w in range(32 + MARGIN // 8):
            if cb + 8 * w < n:
                for j in range(8):
                    a = cb + 8 * w + j
                    B[8 * w + j] = src[a] if 0 <= a < n else 0xAA   # bytes outside the stream inside a mapped word: garbage
        D = np.zeros(SLOTS, dtype=np.int64)
        lim = isafe - cb
        for p in range(CHUNK):
            if p >= lim: continue
            t = int(B[p]); hi, lo = t >> 4, t & 15
            d = 3 + hi + (1 if lo == 15 else 0)
            if hi == 15: d += 1 + int(B[p + 1])
            D[p] = 2 * d
        p2 = 2 * sh; pos = []
        for k in range(32):
            pos.append(p2); p2 += int(D[p2 // 2])
        pos.append(p2)
        assert max(pos) // 2 < SLOTS
        live = [D[x // 2] != 0 for x in pos[:32]]
        k = sum(live)
        assert all(live[:k]) and not any(live[k:])
        fields = []
        bad = []
        for lane in range(32):
            if not live[lane]:
                fields.append(None); bad.append(False); continue
            mp = pos[lane] // 2
            t = int(B[mp]); q = mp + 1; LL = t >> 4; bd = False
            if LL == 15:
                x = int(B[q]); LL += x; q += 1; bd = x == 255
            lit = cb + q; q += LL + 2; ML = t & 15
            if ML == 15:
                staged = q < CHUNK + MARGIN
                y = int(B[q if staged else 0]); ML += y; q += 1
                bd = bd or (not staged) or y == 255
            bd = bd or cb + q > n
            fields.append((cb + mp, lit, LL, ML + 4)); bad.append(bd)
        cut = any(bad)
        if cut: k = bad.index(True)
        ipn = cb + pos[k] // 2
        if k:
            events.append(("batch", fields[:k]))
        ip = ipn
        if cut:
            events.append(("exact", ip))
            t = src[ip]; q = ip + 1; LL = t >> 4
            if LL == 15:
                while True:
                    s = src[q]; q += 1; LL += s
                    if s != 255: break
            q += LL
            if q >= n:
                return events, n
            q += 2
            if (t & 15) == 15:
                while True:
                    s = src[q]; q += 1
                    if s != 255: break
            ip = q
        continue
        seqs = []
        for lane in range(k):
            mp = pos[lane] // 2
            t = int(B[mp]); q = mp + 1; LL = t >> 4
            if LL == 15: LL += int(B[q]); q += 1
            lit = cb + q; ML = t & 15
            if ML == 15:
                y = int(B[q + LL + 2]); ML += y
                if y == 255: ML += int(B[q + LL + 3])
            seqs.append((cb + mp, lit, LL, ML + 4))
        events.append(("batch", seqs))
        ip = cb + p2 // 2
    return events, ip


def check(src, align):
    ref = serial_parse(src)
    ev, ip_end = model(src, align)
    i = 0
    nb = nx = 0
    for e in ev:
        if e[0] == "batch":
            for s in e[1]:
                assert ref[i] == s, (i, ref[i], s, align)
                assert s[1] + s[2] + 2 <= len(src)
                i += 1
            nb += len(e[1])
        else:
            assert ref[i][0] == e[1], (ref[i], e)
            i += 1; nx += 1
    if i < len(ref):
        assert ref[i][0] == ip_end, (ref[i], ip_end, len(src))
    else:
        assert ip_end >= len(src) - SPAN
    return nb, nx, len(ref) - i


if __name__ == "__main__":
    import b2oracle as o, corpus, zig_lz4_b200
    from zig_lz4_b200 import datagen
    o.lib()
    rng = np.random.default_rng(5)
    tot = [0, 0, 0]
    for mode in range(4):
        for n in (4096, 65536):
            d = datagen.generate(n, mode=mode).tobytes()
            for c in (o.compress_fast(d), o.compress_hc(d, 9)):
                for al in (0, 3, 7):
                    r = check(c, al)
                    for j in range(3): tot[j] += r[j]
        print("class", mode, tot, flush=True)
    for it in range(60):
        s, _ = corpus.synth_lz4_stream(rng, int(rng.integers(1, 400)), ll_max=int(rng.choice([3, 20, 40])),
                                       ml_max=int(rng.choice([8, 20, 300])), p_long=float(rng.choice([0, 0.02, 0.3])))
        for al in range(8):
            r = check(s, al)
            for j in range(3): tot[j] += r[j]
    for name, data in corpus.block_cases() + corpus.compat_cases():
        for c in (o.compress_fast(data[:65536]), o.compress_hc(data[:65536], 9)):
            if c:
                for al in (0, 5):
                    check(c, al)
    print("sequences in batches %d, via exact step %d, left to the exact tier at block ends %d" % tuple(tot))
