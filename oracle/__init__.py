"""ctypes binding of the CPU ORACLE (oracle/libmtz_oracle.so).

TEST INFRASTRUCTURE ONLY -- see oracle/mtz_oracle.h.  Importable from tests/,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py``; never from ``manatee_b200`` (tests/test_layout.py enforces it).
"parity unpinned": the reference pins no checksum / codec result for this path.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmtz_oracle.so")

OK, EINVAL, EFORMAT, ECKSUM, ECODEC, ENOSPC = 0, -1, -4, -5, -6, -7
DRR_HDR = 312
PAYLOAD_PCG, PAYLOAD_PGPAGE, PAYLOAD_ZERO = 0, 1, 2


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


class Cksum(C.Structure):
    _fields_ = [("w", C.c_uint64 * 4)]

    def tuple(self):
        return tuple(int(x) for x in self.w)


class Partial(C.Structure):
    _fields_ = [("n", C.c_uint64), ("a", C.c_uint64), ("b", C.c_uint64),
                ("c", C.c_uint64), ("d", C.c_uint64)]

    def tuple(self):
        return (int(self.n), int(self.a), int(self.b), int(self.c), int(self.d))


class StreamStats(C.Structure):
    _fields_ = [("records", C.c_uint64), ("write_records", C.c_uint64),
                ("bytes_in", C.c_uint64), ("bytes_out", C.c_uint64),
                ("bad_record", C.c_uint64), ("lz4_in", C.c_uint64),
                ("lz4_out", C.c_uint64), ("end_cksum", Cksum)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, sz, i32, i64, u64, u32 = C.c_void_p, C.c_size_t, C.c_int, C.c_int64, C.c_uint64, C.c_uint32
        L.orc_fletcher4_incremental.argtypes = [vp, sz, C.POINTER(Cksum)]
        L.orc_fletcher4_native.argtypes = [vp, sz, C.POINTER(Cksum)]
        L.orc_fletcher4_partial.argtypes = [vp, sz, C.POINTER(Partial)]
        L.orc_fletcher4_partial_simd.argtypes = [vp, sz, C.POINTER(Partial), i32]
        L.orc_fletcher4_simd_lanes.argtypes = [i32]
        L.orc_fletcher4_simd_lanes.restype = i32
        L.orc_mt_set_lanes.argtypes = [i32]
        L.orc_mt_set_lanes.restype = i32
        L.orc_fletcher4_apply.argtypes = [C.POINTER(Cksum), C.POINTER(Partial)]
        L.orc_partial_concat.argtypes = [C.POINTER(Partial)] * 3
        L.orc_tri2.argtypes = [u64]; L.orc_tri2.restype = u64
        L.orc_tri3.argtypes = [u64]; L.orc_tri3.restype = u64
        L.orc_drr_payload_len.argtypes = [vp]; L.orc_drr_payload_len.restype = i64
        L.orc_stream_index.argtypes = [vp, sz, vp, sz]; L.orc_stream_index.restype = i64
        L.orc_stream_verify.argtypes = [vp, sz, C.POINTER(StreamStats)]
        for name in ("compress", "decompress", "recompress"):
            f = getattr(L, "orc_stream_" + name)
            f.argtypes = [vp, sz, vp, sz, C.POINTER(sz), C.POINTER(StreamStats)]
            f.restype = i32
        L.orc_stream_restamp.argtypes = [vp, sz, C.POINTER(Cksum)]
        L.orc_lz4_compress_block.argtypes = [vp, i32, vp, i32]
        L.orc_lz4_decompress_block.argtypes = [vp, i32, vp, i32]
        L.orc_zfs_lz4_compress.argtypes = [vp, sz, vp]; L.orc_zfs_lz4_compress.restype = sz
        L.orc_zfs_lz4_decompress.argtypes = [vp, sz, vp, sz]
        L.orc_gen_payload.argtypes = [i32, u64, vp, sz]; L.orc_gen_payload.restype = None
        L.orc_synth_stream_size.argtypes = [u64, u32]; L.orc_synth_stream_size.restype = sz
        L.orc_synth_stream.argtypes = [vp, sz, C.POINTER(sz), u64, u32, i32, u64, i32]
        L.orc_synth_shard_size.argtypes = [u64, u32, i32]; L.orc_synth_shard_size.restype = sz
        L.orc_synth_shard_fill.argtypes = [vp, sz, u64, u32, i32, u64, i32, vp, i32]
        L.orc_synth_shard_stamp.argtypes = [vp, u64, u32, i32, vp, C.POINTER(Cksum)]
        L.orc_mt_verify.argtypes = [vp, sz, i32, C.POINTER(C.c_double), C.POINTER(StreamStats)]
        L.orc_mt_recompress.argtypes = [vp, sz, vp, sz, C.POINTER(sz), i32,
                                        C.POINTER(C.c_double), C.POINTER(StreamStats)]
        _lib = L
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _u8(buf):
    if isinstance(buf, np.ndarray):
        a = buf if buf.dtype == np.uint8 else buf.view(np.uint8)
        return np.ascontiguousarray(a)
    return np.frombuffer(bytes(buf), dtype=np.uint8)


def fletcher4(buf, state=None):
    a = _u8(buf)
    ck = Cksum()
    if state is not None:
        for i in range(4):
            ck.w[i] = state[i]
    lib().orc_fletcher4_incremental(_ptr(a), a.size, C.byref(ck))
    return ck.tuple()


def fletcher4_partial(buf):
    a = _u8(buf)
    p = Partial()
    lib().orc_fletcher4_partial(_ptr(a), a.size, C.byref(p))
    return p.tuple()


def fletcher4_apply(state, partial):
    ck = Cksum()
    for i in range(4):
        ck.w[i] = state[i]
    p = Partial(*partial)
    lib().orc_fletcher4_apply(C.byref(ck), C.byref(p))
    return ck.tuple()


def partial_concat(x, y):
    px, py, po = Partial(*x), Partial(*y), Partial()
    lib().orc_partial_concat(C.byref(px), C.byref(py), C.byref(po))
    return po.tuple()


def synth_stream(nwrites, recsize=131072, kind=PAYLOAD_PCG, first_rec=0, nthreads=0, out=None):
    """Seeded BEGIN/OBJECT/WRITE*n/END stream as a numpy uint8 array."""
    L = lib()
    need = L.orc_synth_stream_size(nwrites, recsize)
    if out is None:
        out = np.empty(need, dtype=np.uint8)
    n = C.c_size_t(0)
    rc = L.orc_synth_stream(_ptr(out), out.size, C.byref(n), nwrites, recsize, kind, first_rec,
                            nthreads or (os.cpu_count() or 1))
    if rc != OK:
        raise RuntimeError("orc_synth_stream rc=%d" % rc)
    return out[:n.value]


def synth_shard_fill(nwrites, recsize, kind, first_rec, flags, out=None, nthreads=0):
    """Step 1 of a record-index shard (see synth.c): returns (bytes, ppay)."""
    L = lib()
    need = L.orc_synth_shard_size(nwrites, recsize, flags)
    if out is None:
        out = np.empty(need, dtype=np.uint8)
    ppay = np.zeros((nwrites + 1, 5), dtype=np.uint64)
    rc = L.orc_synth_shard_fill(_ptr(out), out.size, nwrites, recsize, kind, first_rec, flags,
                                _ptr(ppay), nthreads or (os.cpu_count() or 1))
    if rc != OK:
        raise RuntimeError("orc_synth_shard_fill rc=%d" % rc)
    return out[:need], ppay


def synth_shard_stamp(buf, nwrites, recsize, flags, ppay, state):
    """Step 2: stamp checksums from running `state`; returns the state after the shard."""
    ck = Cksum()
    for i in range(4):
        ck.w[i] = state[i]
    rc = lib().orc_synth_shard_stamp(_ptr(buf), nwrites, recsize, flags, _ptr(ppay), C.byref(ck))
    if rc != OK:
        raise RuntimeError("orc_synth_shard_stamp rc=%d" % rc)
    return ck.tuple()


def stream_index(stream):
    a = _u8(stream)
    cnt = lib().orc_stream_index(_ptr(a), a.size, None, 0)
    if cnt < 0:
        return cnt, None
    offs = np.empty(cnt, dtype=np.uint64)
    lib().orc_stream_index(_ptr(a), a.size, _ptr(offs), cnt)
    return cnt, offs


def stream_verify(stream):
    a = _u8(stream)
    st = StreamStats()
    rc = lib().orc_stream_verify(_ptr(a), a.size, C.byref(st))
    return rc, st


def _xform(name, stream, cap=None):
    a = _u8(stream)
    if cap is None:
        cap = a.size * 2 + (1 << 20)
        if name == "decompress":
            cap = a.size * 16 + (1 << 20)
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t(0)
    st = StreamStats()
    rc = getattr(lib(), "orc_stream_" + name)(_ptr(a), a.size, _ptr(out), out.size, C.byref(n),
                                               C.byref(st))
    return rc, out[:n.value], st


def stream_compress(stream, cap=None):
    return _xform("compress", stream, cap)


WIRE_MAGIC = bytes.fromhex("4d545a4c5a345731")          # "MTZLZ4W1", mtz_oracle.h ORC_WIRE_MAGIC
WIRE_PRE_BYTES = 32


def wire_strip(stream):
    """The COMPRESS output without the lz4-stage-v1 preambles (32 bytes in front of every BEGIN):
    the plain compressed send stream underneath, i.e. what `zfs send -c` would carry."""
    a = _u8(stream)
    L = lib()
    L.orc_drr_payload_len.restype = C.c_int64
    L.orc_drr_payload_len.argtypes = [C.c_void_p]
    parts, pos, start = [], 0, 0
    n = a.size
    while pos < n:
        if n - pos >= WIRE_PRE_BYTES and a[pos:pos + 8].tobytes() == WIRE_MAGIC:
            if pos > start:
                parts.append(a[start:pos])
            pos += WIRE_PRE_BYTES
            start = pos
            continue
        if n - pos < 312:
            break
        pl = L.orc_drr_payload_len(a[pos:].ctypes.data)
        if pl < 0:
            break
        pos += 312 + pl
    parts.append(a[start:])
    return np.ascontiguousarray(np.concatenate(parts)) if len(parts) > 1 else np.ascontiguousarray(parts[0])


def stream_compress_plain(stream, cap=None):
    """COMPRESS without the wire framing: the compressed send stream itself (what the device-level
    pipeline produces and what RECOMPRESS / VERIFY take) -- stream_compress() minus its preambles."""
    rc, c, st = stream_compress(stream, cap)
    return rc, (wire_strip(c) if rc == 0 else c), st


def stream_decompress(stream, cap=None):
    return _xform("decompress", stream, cap)


def stream_recompress(stream, cap=None):
    return _xform("recompress", stream, cap)


def stream_restamp(stream):
    ck = Cksum()
    rc = lib().orc_stream_restamp(_ptr(stream), stream.size, C.byref(ck))
    return rc, ck.tuple()


def lz4_compress_block(src, osize=None):
    a = _u8(src)
    osize = osize if osize is not None else a.size + a.size // 255 + 64
    dst = np.empty(osize + 16, dtype=np.uint8)
    n = lib().orc_lz4_compress_block(_ptr(a), a.size, _ptr(dst), osize)
    return dst[:n].copy()


def lz4_decompress_block(src, maxout):
    a = _u8(src)
    dst = np.empty(maxout + 16, dtype=np.uint8)
    n = lib().orc_lz4_decompress_block(_ptr(a), a.size, _ptr(dst), maxout)
    return n, (dst[:n].copy() if n >= 0 else None)


def zfs_lz4_compress(src):
    a = _u8(src)
    dst = np.empty(a.size + 1024, dtype=np.uint8)
    ps = lib().orc_zfs_lz4_compress(_ptr(a), a.size, _ptr(dst))
    return ps, (dst[:ps].copy() if ps < a.size else None)


def zfs_lz4_decompress(src, lsize):
    a = _u8(src)
    dst = np.empty(lsize + 16, dtype=np.uint8)
    rc = lib().orc_zfs_lz4_decompress(_ptr(a), a.size, _ptr(dst), lsize)
    return rc, dst[:lsize]


def gen_payload(kind, recidx, length):
    dst = np.empty(length, dtype=np.uint8)
    lib().orc_gen_payload(kind, recidx, _ptr(dst), length)
    return dst


def fletcher4_partial_simd(buf, lanes=-1):
    """lane-parallel form (baseline only); lanes: -1 best, 0 scalar, 4 avx2, 8 avx512f"""
    a = _u8(buf)
    p = Partial()
    lib().orc_fletcher4_partial_simd(_ptr(a), a.size, C.byref(p), lanes)
    return p.tuple()


def simd_lanes(force=-1):
    return lib().orc_fletcher4_simd_lanes(force)


def mt_release():
    """free mt_recompress's cached scratch (as large as the logical stream)"""
    lib().orc_mt_release()


def mt_set_lanes(lanes):
    """Fletcher-4 flavour of mt_verify / mt_recompress; returns the lanes in use"""
    return lib().orc_mt_set_lanes(lanes)


def mt_verify(stream, nthreads):
    a = _u8(stream)
    st = StreamStats()
    secs = C.c_double(0)
    rc = lib().orc_mt_verify(_ptr(a), a.size, nthreads, C.byref(secs), C.byref(st))
    return rc, secs.value, st


def mt_recompress(stream, nthreads, cap=None):
    a = _u8(stream)
    cap = cap or a.size * 3 + (1 << 20)
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t(0)
    st = StreamStats()
    secs = C.c_double(0)
    rc = lib().orc_mt_recompress(_ptr(a), a.size, _ptr(out), out.size, C.byref(n), nthreads,
                                 C.byref(secs), C.byref(st))
    return rc, secs.value, out[:n.value], st
