"""LZ4 block surgery for the certificate tests (test infrastructure): parse a block into its
sequences, change the PARSE without changing what it decodes to, emit it again, and splice the
result into a ZFS-send stream.  Every mutation yields a valid block of the public LZ4 format that is
NOT what the declared encoder (oracle/lz4_zfs.c) emits for those bytes -- RECOMPRESS must therefore
re-encode it, and K3c (kernels_lz4.cuh warp_lz4_certify) must refuse to certify it."""
import numpy as np


def parse(block):
    """-> ([(literals: bytes, offset, matchlen)], closing literals: bytes)"""
    b = bytes(block)
    ip, seqs = 0, []
    while True:
        tok = b[ip]; ip += 1
        ll = tok >> 4
        if ll == 15:
            while True:
                s = b[ip]; ip += 1
                ll += s
                if s != 255:
                    break
        lits = b[ip:ip + ll]; ip += ll
        if ip == len(b):
            return seqs, lits
        off = b[ip] | (b[ip + 1] << 8); ip += 2
        ml = tok & 15
        if ml == 15:
            while True:
                s = b[ip]; ip += 1
                ml += s
                if s != 255:
                    break
        seqs.append((lits, off, ml + 4))


def _lenext(v):
    out = bytearray()
    while v >= 255:
        out.append(255); v -= 255
    out.append(v)
    return bytes(out)


def emit(seqs, tail, last_nibble=0):
    out = bytearray()
    for lits, off, ml in seqs:
        ll, m = len(lits), ml - 4
        out.append((min(ll, 15) << 4) | min(m, 15))
        if ll >= 15:
            out += _lenext(ll - 15)
        out += lits
        out += bytes((off & 255, off >> 8))
        if m >= 15:
            out += _lenext(m - 15)
    out.append((min(len(tail), 15) << 4) | (last_nibble & 15))
    if len(tail) >= 15:
        out += _lenext(len(tail) - 15)
    out += tail
    return bytes(out)


def decode(seqs, tail):
    d = bytearray()
    for lits, off, ml in seqs:
        d += lits
        for _ in range(ml):
            d.append(d[-off])
    d += tail
    return bytes(d)


MUTATIONS = ("split", "shorten", "shift", "offset", "unmatch", "nibble")


def mutate(block, kind, rng):
    """One parse change of `kind` somewhere in the block; None when the block offers no place for it."""
    seqs, tail = parse(block)
    if not seqs:
        return None
    if kind == "nibble":
        return emit(seqs, tail, last_nibble=1 + int(rng.integers(15)))
    order = list(rng.permutation(len(seqs)))
    data = decode(seqs, tail)
    starts, pos = [], 0
    for lits, off, ml in seqs:
        pos += len(lits); starts.append(pos); pos += ml
    for i in order:
        lits, off, ml = seqs[i]
        m = starts[i]
        nxt = seqs[i + 1] if i + 1 < len(seqs) else None
        if kind == "split" and ml >= 8:
            k = 4 + int(rng.integers(ml - 7))
            new = seqs[:i] + [(lits, off, k), (b"", off, ml - k)] + seqs[i + 1:]
            return emit(new, tail)
        if kind == "shorten" and ml >= 5:
            last = data[m + ml - 1:m + ml]
            new = list(seqs)
            new[i] = (lits, off, ml - 1)
            if nxt is not None:
                new[i + 1] = (last + nxt[0], nxt[1], nxt[2])
                return emit(new, tail)
            return emit(new, last + tail)
        if kind == "shift" and ml >= 5:
            new = list(seqs)
            new[i] = (lits + data[m:m + 1], off, ml - 1)
            return emit(new, tail)
        if kind == "offset":
            want = data[m:m + ml]
            lo = max(0, m - 65535)
            j = data.rfind(want, lo, m + ml - 1)
            while j != -1 and (m - j == off or j >= m):
                j = data.rfind(want, lo, j + ml - 1) if j > lo else -1
            if j != -1 and 0 < m - j <= 65535:
                new = list(seqs)
                new[i] = (lits, m - j, ml)
                return emit(new, tail)
        if kind == "unmatch":
            body = lits + data[m:m + ml]
            if nxt is not None:
                new = seqs[:i] + [(body + nxt[0], nxt[1], nxt[2])] + seqs[i + 2:]
                return emit(new, tail)
            return emit(seqs[:i], body + tail)
    return None


def splice(oracle, stream, changes):
    """`stream`: a `zfs send -c`-shaped stream.  changes: {write-record ordinal: new LZ4 block bytes}.
    Returns the stream with those records' frames replaced (BE32 length, sector padding, header
    sizes) and every checksum re-stamped."""
    cnt, offs = oracle.stream_index(stream)
    parts, w = [], 0
    for k in range(cnt):
        o = int(offs[k]); e = int(offs[k + 1]) if k + 1 < cnt else stream.size
        h = stream[o:o + 312].copy()
        pay = stream[o + 312:e]
        if int(h[0]) == 3:
            if w in changes and int(h[50]) == 15:
                blk = np.frombuffer(changes[w], dtype=np.uint8)
                m = blk.size
                ps = (m + 4 + 511) & ~511
                fr = np.zeros(ps, dtype=np.uint8)
                fr[0:4] = [m >> 24, (m >> 16) & 255, (m >> 8) & 255, m & 255]
                fr[4:4 + m] = blk
                h[96:104] = np.array([ps], dtype=np.uint64).view(np.uint8)
                pay = fr
            w += 1
        parts += [h, pay]
    out = np.ascontiguousarray(np.concatenate(parts))
    assert oracle.stream_restamp(out)[0] == 0
    return out


def blocks_of(oracle, stream):
    """[(write-record ordinal, LZ4 block bytes)] of the compressed WRITE records of a stream"""
    cnt, offs = oracle.stream_index(stream)
    out, w = [], 0
    for k in range(cnt):
        o = int(offs[k]); e = int(offs[k + 1]) if k + 1 < cnt else stream.size
        if int(stream[o]) == 3:
            if int(stream[o + 50]) == 15:
                pay = stream[o + 312:e]
                clen = (int(pay[0]) << 24) | (int(pay[1]) << 16) | (int(pay[2]) << 8) | int(pay[3])
                out.append((w, bytes(pay[4:4 + clen])))
            w += 1
    return out
