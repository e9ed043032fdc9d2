"""GPU parity for the re-encoding stage modes (COMPRESS / DECOMPRESS / RECOMPRESS):
whole-stream output must equal the CPU oracle's byte for byte (headers, LZ4
frames, every re-stamped checksum), through every entry point of the C ABI."""
import hashlib
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _mixed_stream(oracle, n=40, recsize=131072):
    """pg-page records with a few incompressible and all-zero ones mixed in."""
    s = oracle.synth_stream(n, recsize=recsize, kind=oracle.PAYLOAD_PGPAGE).copy()
    cnt, offs = oracle.stream_index(s)
    for k in range(2, cnt - 1):
        o = int(offs[k]) + 312
        if k % 7 == 3:
            s[o:o + recsize] = oracle.gen_payload(oracle.PAYLOAD_PCG, k, recsize)
        elif k % 11 == 5:
            s[o:o + recsize] = 0
    rc, _ = oracle.stream_restamp(s)
    assert rc == 0
    return s


def _gpu(mode, src, cap=None, **kw):
    from manatee_b200 import GpuSnapshotStage
    out = np.zeros(cap or (src.size * 2 + (1 << 20)), dtype=np.uint8)
    with GpuSnapshotStage(mode, **kw) as g:
        n = g.process_host(src, out)
        return out[:n].copy(), g.stats(), g.end_checksum()


@pytest.mark.parametrize("batch", [0, 1 << 20, 5 << 20])
def test_compress_matches_oracle(oracle, batch):
    s = _mixed_stream(oracle)
    rc, want, st = oracle.stream_compress(s)
    assert rc == 0
    got, gs, end = _gpu("compress", s, batch_bytes=batch, n_slots=3)
    assert got.size == want.size
    assert np.array_equal(got, want)
    assert gs["lz4_encoded"] == st.lz4_out
    assert end == st.end_cksum.tuple()
    assert oracle.stream_verify(oracle.wire_strip(got))[0] == 0      # the send stream under the wire framing


@pytest.mark.parametrize("recsize", [4096, 65536, 131072, 1 << 20])
def test_transport_identity_compress_then_decompress(oracle, recsize):
    """The one contract the reference guarantees: bytes into zfs recv == bytes out of zfs send."""
    s = _mixed_stream(oracle, n=24, recsize=recsize)
    c, _, _ = _gpu("compress", s)
    rc, want_d, st = oracle.stream_decompress(c)
    assert rc == 0
    d, gs, end = _gpu("decompress", c, cap=s.size + (1 << 20))
    assert np.array_equal(d, s), "DECOMPRESS(COMPRESS(x)) != x"
    assert np.array_equal(d, want_d)
    assert gs["lz4_decoded"] == st.lz4_in


def test_recompress_matches_oracle_and_is_idempotent(oracle):
    s = _mixed_stream(oracle, n=30)
    rc, c, _ = oracle.stream_compress_plain(s)
    # a stream compressed by a DIFFERENT encoder: rebuild frames with liblz4
    import ctypes as C
    lz = C.CDLL("liblz4.so.1")
    lz.LZ4_compress_default.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    cnt, offs = oracle.stream_index(s)
    parts = []
    for k in range(cnt):
        o = int(offs[k]); e = int(offs[k + 1]) if k + 1 < cnt else s.size
        h = s[o:o + 312].copy()
        if int(h[0]) == 3:
            p = s[o + 312:e]
            buf = np.empty(p.size + 4096, dtype=np.uint8)
            m = lz.LZ4_compress_default(p.ctypes.data, buf.ctypes.data, p.size, buf.size)
            ps = (m + 4 + 511) & ~511
            if ps < p.size - (p.size >> 3):
                fr = np.zeros(ps, dtype=np.uint8)
                fr[0:4] = [m >> 24, (m >> 16) & 255, (m >> 8) & 255, m & 255]
                fr[4:4 + m] = buf[:m]
                h[50] = 15
                h[96:104] = np.array([ps], dtype=np.uint64).view(np.uint8)
                parts += [h, fr]
                continue
        if int(h[0]) == 0:
            vi = int(h[16:24].view(np.uint64)[0]) | (((1 << 22) | (1 << 17)) << 2)
            h[16:24] = np.array([vi], dtype=np.uint64).view(np.uint8)
        parts += [h, s[o + 312:e]]
    foreign = np.concatenate(parts)
    assert oracle.stream_restamp(foreign)[0] == 0
    rc, want, st = oracle.stream_recompress(foreign)
    assert rc == 0
    got, gs, end = _gpu("recompress", foreign, cap=s.size + (1 << 20), batch_bytes=3 << 20)
    assert np.array_equal(got, want)
    assert end == st.end_cksum.tuple()
    # idempotence: recompressing our own output changes nothing
    again, _, _ = _gpu("recompress", got, cap=s.size + (1 << 20))
    assert np.array_equal(again, got)
    # (not equal to COMPRESS(raw): COMPRESS marks BEGIN for its DECOMPRESS peer, RECOMPRESS does not)
    assert got.size == c.size


def test_incompressible_stream_passes_through(oracle):
    s = oracle.synth_stream(12, kind=oracle.PAYLOAD_PCG)
    rc, want, st = oracle.stream_compress(s)
    got, gs, _ = _gpu("compress", s)
    assert np.array_equal(got, want) and gs["lz4_encoded"] == 0
    assert got.size == s.size + oracle.WIRE_PRE_BYTES     # stored raw; only the wire preamble is added


def test_streaming_compress_and_decompress(oracle):
    from manatee_b200 import GpuSnapshotStage
    s = _mixed_stream(oracle, n=50)
    rc, want, _ = oracle.stream_compress(s)

    def pump(mode, data, chunk):
        out = bytearray()
        with GpuSnapshotStage(mode, ring_bytes=16 << 20, batch_bytes=4 << 20, n_slots=3,
                              out_ring_bytes=8 << 20) as g:
            def prod():
                for i in range(0, len(data), chunk):
                    g.write(np.frombuffer(data[i:i + chunk], dtype=np.uint8))
                g.flush()
            t = threading.Thread(target=prod); t.start()
            while True:
                b = g.read(1 << 20)
                if b is None:
                    break
                out += b
            t.join()
        return bytes(out)

    c = pump("compress", s.tobytes(), 65536)
    assert hashlib.sha256(c).hexdigest() == hashlib.sha256(want.tobytes()).hexdigest()
    d = pump("decompress", c, 100003)
    assert d == s.tobytes()


def test_device_api_recompress_subbatched(oracle):
    import torch
    from manatee_b200 import GpuSnapshotStage, index_host
    s = _mixed_stream(oracle, n=64)
    rc, c, _ = oracle.stream_compress_plain(s)
    rc, want, st = oracle.stream_recompress(c)
    recs, used = index_host(c)
    d_in = torch.from_numpy(c).cuda()
    d_recs = torch.from_numpy(recs.view(np.uint8).copy()).cuda()
    d_out = torch.zeros(s.size + (1 << 20), dtype=torch.uint8, device="cuda")
    with GpuSnapshotStage("recompress") as g:
        g.dev_submit(d_in.data_ptr(), c.size, d_recs.data_ptr(), len(recs), d_out.data_ptr(), d_out.numel())
        ob, carry, carry_out = g.dev_finish()
        got = d_out[:ob].cpu().numpy()
        assert np.array_equal(got, want)
        assert g.end_checksum() == st.end_cksum.tuple()
        assert carry_out == oracle.fletcher4(want)


def test_corrupt_frame_is_ecodec_at_the_oracles_record(oracle):
    from manatee_b200._native import MtzError, ECODEC
    s = _mixed_stream(oracle, n=20)
    rc, c, _ = oracle.stream_compress(s)
    c = c.copy()
    W = oracle.WIRE_PRE_BYTES                              # the lz4-stage-v1 preamble in front of BEGIN
    cnt, offs = oracle.stream_index(c[W:])
    o = W + int(offs[9])
    c[o + 312 + 6] = 0xff; c[o + 312 + 7] = 0xff          # first match offset -> beyond block start
    assert oracle.stream_restamp(c[W:])[0] == 0            # checksums valid, frame is not
    rc, _, st = oracle.stream_decompress(c)
    assert rc == oracle.ECODEC
    with pytest.raises(MtzError) as ei:
        _gpu("decompress", c, cap=s.size + (1 << 20))
    assert ei.value.code == ECODEC


def test_mode_preconditions_like_the_oracle(oracle):
    from manatee_b200._native import MtzError, EINVAL
    s = _mixed_stream(oracle, n=4)
    rc, c, _ = oracle.stream_compress_plain(s)             # an already compressed send stream
    assert oracle.stream_compress(c)[0] == oracle.EINVAL
    assert oracle.stream_decompress(s)[0] == oracle.EINVAL
    with pytest.raises(MtzError) as ei:
        _gpu("compress", c)
    assert ei.value.code == EINVAL
    with pytest.raises(MtzError) as ei:
        _gpu("decompress", s)
    assert ei.value.code == EINVAL


def test_golden_fixtures_on_gpu(oracle):
    """The committed golden vectors (tests/golden/, generated by make_golden.py)."""
    import json
    import os
    from manatee_b200 import GpuSnapshotStage
    from manatee_b200._native import MtzError
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    meta = json.load(open(os.path.join(gold, "golden.json")))
    s = np.fromfile(os.path.join(gold, "stream_small.bin"), dtype=np.uint8)
    c = np.fromfile(os.path.join(gold, "stream_small_lz4.bin"), dtype=np.uint8)
    with GpuSnapshotStage("verify") as g:
        g.process_host(s)
        assert ["%016x" % x for x in g.end_checksum()] == meta["stream_small"]["end_cksum"]
    got, gs, end = _gpu("compress", s)
    assert np.array_equal(got, c)
    assert ["%016x" % x for x in end] == meta["stream_small_lz4"]["end_cksum"]
    back, _, _ = _gpu("decompress", c)
    assert np.array_equal(back, s)
    bad = np.fromfile(os.path.join(gold, "stream_small_corrupt_5.bin"), dtype=np.uint8)
    with GpuSnapshotStage("verify") as g:
        with pytest.raises(MtzError):
            g.process_host(bad)
        assert g.stats()["bad_record"] == meta["stream_small_corrupt_5"]["bad_record"]


def _all_types_stream(oracle, seed=7):
    """A stream exercising every DRR record type and odd sizes (the oracle generator only
    emits BEGIN/OBJECT/WRITE/END): built by hand, stamped by the oracle."""
    rng = np.random.default_rng(seed)

    def hdr(t, f=None):
        f = f or {}
        h = np.zeros(312, dtype=np.uint8)
        h[0:4] = np.array([t], dtype=np.uint32).view(np.uint8)
        for off, (val, width) in f.items():
            h[off:off + width] = np.array([val], dtype={4: np.uint32, 8: np.uint64}[width]).view(np.uint8)
        return h

    parts = []
    b = hdr(0, {8: (0x2F5bacbac, 8), 16: (1 | (0x4 << 2), 8), 4: (24, 4)})
    parts += [b, rng.integers(0, 256, 24, dtype=np.uint8)]                       # BEGIN with a 24-byte payload
    parts += [hdr(1, {8: (8, 8), 28: (13, 4)}), rng.integers(0, 256, 16, dtype=np.uint8)]   # OBJECT bonus 13 -> 16
    parts += [hdr(2, {8: (9, 8), 16: (100, 8)})]                               # FREEOBJECTS
    for k, ls in enumerate([512, 1024, 4096, 131072, 16384, 2 << 20]):
        pay = oracle.gen_payload(oracle.PAYLOAD_PGPAGE if k % 2 == 0 else oracle.PAYLOAD_PCG, 50 + k, ls)
        w = hdr(3, {8: (8, 8), 24: (k * (1 << 21), 8), 32: (ls, 8)})
        w[48] = 7
        parts += [w, pay]
        if k == 2:
            parts += [hdr(4, {8: (8, 8), 16: (1 << 30, 8), 24: (4096, 8)})]   # FREE
            parts += [hdr(7, {8: (8, 8), 16: (520, 8)}), rng.integers(0, 256, 520, dtype=np.uint8)]  # SPILL
        if k == 4:
            parts += [hdr(6, {8: (8, 8), 16: (0, 8), 24: (4096, 8)})]          # WRITE_BYREF
            e = hdr(8, {8: (8, 8), 16: (0, 8), 24: (512, 8), 48: (512, 4), 52: (21, 4)})
            parts += [e, rng.integers(0, 256, 24, dtype=np.uint8)]               # WRITE_EMBEDDED psize 21 -> 24
    parts += [hdr(5)]
    s = np.concatenate(parts)
    rc, _ = oracle.stream_restamp(s)
    assert rc == 0 and oracle.stream_verify(s)[0] == 0
    return s


def test_every_record_type_all_modes(oracle):
    from manatee_b200 import GpuSnapshotStage
    s = _all_types_stream(oracle)
    rc, st = oracle.stream_verify(s)
    with GpuSnapshotStage("verify", batch_bytes=1 << 20) as g:
        g.process_host(s)
        assert g.end_checksum() == st.end_cksum.tuple() and g.stats()["records"] == st.records
    rc, want, st = oracle.stream_compress(s)
    assert rc == 0
    got, gs, end = _gpu("compress", s, batch_bytes=1 << 20)
    assert np.array_equal(got, want) and end == st.end_cksum.tuple()
    back, _, _ = _gpu("decompress", got)
    assert np.array_equal(back, s)
    plain = oracle.wire_strip(got)                  # RECOMPRESS takes the send stream, not the stage wire
    rc, want_r, _ = oracle.stream_recompress(plain)
    got_r, _, _ = _gpu("recompress", plain)
    assert np.array_equal(got_r, want_r)


def test_empty_and_tiny_streams(oracle):
    from manatee_b200 import GpuSnapshotStage
    empty = np.zeros(0, dtype=np.uint8)
    for mode in ("verify", "compress"):
        with GpuSnapshotStage(mode) as g:
            out = np.zeros(1024, dtype=np.uint8)
            assert g.process_host(empty, out) == 0
    s = oracle.synth_stream(0)                      # BEGIN, OBJECT, END only
    got, gs, end = _gpu("compress", s)
    rc, want, st = oracle.stream_compress(s)
    assert np.array_equal(got, want) and gs["records"] == 3


def test_sixteen_mib_record(oracle):
    """largest ZFS block: K1 chunk loop (T3 weights overflow guard) and the u32-table encoder"""
    s = oracle.synth_stream(2, recsize=16 << 20, kind=oracle.PAYLOAD_PGPAGE)
    rc, want, st = oracle.stream_compress(s)
    got, gs, end = _gpu("compress", s, batch_bytes=8 << 20)
    assert np.array_equal(got, want) and end == st.end_cksum.tuple()
    back, _, _ = _gpu("decompress", got, cap=s.size + (1 << 20), batch_bytes=8 << 20)
    assert np.array_equal(back, s)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_randomized_streams_all_modes(oracle, seed):
    """Seeded fuzz: random record mix / sizes / payload kinds / batch sizes; every mode must
    reproduce the oracle bit for bit, and DECOMPRESS(COMPRESS(x)) == x."""
    from manatee_b200 import GpuSnapshotStage
    rng = np.random.default_rng(1000 + seed)

    def u(h, off, val, width):
        h[off:off + width] = np.array([val], dtype={4: np.uint32, 8: np.uint64}[width]).view(np.uint8)

    parts = []
    b = np.zeros(312, dtype=np.uint8)
    u(b, 8, 0x2F5bacbac, 8); u(b, 16, 1 | (0x4 << 2), 8)
    parts.append(b)
    for k in range(int(rng.integers(20, 60))):
        kind = int(rng.integers(0, 10))
        h = np.zeros(312, dtype=np.uint8)
        if kind <= 5:                                   # DRR_WRITE, sizes 512 B .. 256 KiB
            ls = int(rng.choice([512, 1024, 2048, 8192, 16384, 65536, 131072, 262144]))
            pk = int(rng.integers(0, 3))
            pay = oracle.gen_payload([oracle.PAYLOAD_PGPAGE, oracle.PAYLOAD_PCG, oracle.PAYLOAD_ZERO][pk],
                                     int(rng.integers(0, 1 << 20)), ls)
            if pk == 0 and rng.integers(0, 2):
                pay = pay.copy(); pay[ls // 3:] = rng.integers(0, 256, ls - ls // 3, dtype=np.uint8)
            u(h, 0, 3, 4); u(h, 8, 8, 8); u(h, 24, k << 18, 8); u(h, 32, ls, 8); h[48] = 7
            parts += [h, pay]
        elif kind == 6:
            u(h, 0, 1, 4); bl = int(rng.integers(0, 300)); u(h, 28, bl, 4)
            parts += [h, rng.integers(0, 256, (bl + 7) & ~7, dtype=np.uint8)]
        elif kind == 7:
            u(h, 0, 4, 4); u(h, 8, 8, 8); u(h, 16, k << 20, 8); u(h, 24, 4096, 8)
            parts.append(h)
        elif kind == 8:
            u(h, 0, 7, 4); ln = int(rng.integers(1, 64)) * 8; u(h, 16, ln, 8)
            parts += [h, rng.integers(0, 256, ln, dtype=np.uint8)]
        else:
            u(h, 0, 2, 4); u(h, 8, k, 8); u(h, 16, 5, 8)
            parts.append(h)
    e = np.zeros(312, dtype=np.uint8); u(e, 0, 5, 4)
    parts.append(e)
    s = np.concatenate(parts)
    assert oracle.stream_restamp(s)[0] == 0
    rc, vst = oracle.stream_verify(s)
    assert rc == 0
    batch = int(rng.choice([1 << 19, 1 << 20, 3 << 20, 0]))
    with GpuSnapshotStage("verify", batch_bytes=batch) as g:
        g.process_host(s)
        assert g.end_checksum() == vst.end_cksum.tuple()
    rc, want_c, cst = oracle.stream_compress(s)
    got_c, gs, end = _gpu("compress", s, batch_bytes=batch, n_slots=int(rng.integers(1, 5)))
    assert np.array_equal(got_c, want_c) and end == cst.end_cksum.tuple()
    got_d, _, _ = _gpu("decompress", got_c, cap=s.size + (1 << 20), batch_bytes=batch)
    assert np.array_equal(got_d, s)
    plain = oracle.wire_strip(got_c)
    rc, want_r, _ = oracle.stream_recompress(plain)
    got_r, _, _ = _gpu("recompress", plain, cap=s.size + (1 << 20), batch_bytes=batch)
    assert np.array_equal(got_r, want_r)


def test_codec_shards_with_deferred_chain(oracle):
    """Record-index sharding of RECOMPRESS (what two ranks do): K2/K3/assemble per shard in
    any order, input verdict after the 40-byte aggregate exchange, output stamp chain hopping
    shard to shard with the 32-byte checksum.  Concatenated output == oracle on the whole."""
    import torch
    from manatee_b200 import GpuSnapshotStage, index_host
    from manatee_b200 import shard as SH
    from manatee_b200._native import FLAG_DEFER_VERIFY
    s = _mixed_stream(oracle, n=48)
    rc, c, _ = oracle.stream_compress_plain(s)
    rc, want, st = oracle.stream_recompress(c)
    recs, used = index_host(c)
    cut = 23
    cut_off = int(recs["off"][cut])
    shards = [(0, cut_off, recs[:cut].copy()), (cut_off, c.size - cut_off, recs[cut:].copy())]
    shards[1][2]["off"] -= cut_off
    d_in = torch.from_numpy(c).cuda()
    stages, outs, aggs = [], [], []
    for (o, n, r) in reversed(shards):                      # submit order is irrelevant
        g = GpuSnapshotStage("recompress", flags=FLAG_DEFER_VERIFY)
        d_r = torch.from_numpy(r.view(np.uint8).copy()).cuda()
        d_o = torch.zeros(int(r["lsize"].sum()) + 312 * len(r) + (1 << 20), dtype=torch.uint8, device="cuda")
        g.dev_submit(d_in.data_ptr() + o, n, d_r.data_ptr(), len(r), d_o.data_ptr(), d_o.numel())
        stages.insert(0, g); outs.insert(0, (d_o, d_r)); aggs.insert(0, g.dev_aggregate())
    try:
        carry_out = (0, 0, 0, 0)
        pieces = []
        for k, g in enumerate(stages):
            ob, _, carry_out = g.dev_finish(carry_in=SH.carry_before(k, aggs), carry_out_in=carry_out)
            pieces.append(outs[k][0][:ob].cpu().numpy())
        got = np.concatenate(pieces)
        assert np.array_equal(got, want)
        assert carry_out == oracle.fletcher4(want)
        assert stages[1].end_checksum() == st.end_cksum.tuple()
    finally:
        for g in stages:
            g.close()


def test_size_independent_properties_at_2gib(oracle):
    """Properties that need no oracle pass over the data (so they scale to BASELINE sizes;
    bench.py checks the same ones at 16 GiB / 64 GiB): END checksum of a GPU-compressed stream
    verifies on the GPU; DECOMPRESS(COMPRESS(x)) == x; RECOMPRESS is idempotent; one flipped
    bit anywhere is reported at the record that follows it."""
    import torch
    from manatee_b200 import GpuSnapshotStage, index_host
    from manatee_b200._native import MtzError, ECKSUM
    nw = (2 << 30) // 131384
    s = oracle.synth_stream(nw, kind=oracle.PAYLOAD_PGPAGE)
    recs, used = index_host(s)
    d_in = torch.from_numpy(s).cuda()
    d_recs = torch.from_numpy(recs.view(np.uint8).copy()).cuda()
    d_c = torch.empty(s.size + (1 << 20), dtype=torch.uint8, device="cuda")
    with GpuSnapshotStage("compress") as g:
        g.dev_submit(d_in.data_ptr(), s.size, d_recs.data_ptr(), len(recs), d_c.data_ptr(), d_c.numel())
        nc, _, _ = g.dev_finish()
        end_c = g.end_checksum()
        assert g.stats()["lz4_encoded"] == nw and nc < s.size // 2
    # verify the compressed stream with a GPU-built table
    d_rc = torch.empty((len(recs) + 16) * 32, dtype=torch.uint8, device="cuda")
    with GpuSnapshotStage("verify") as g:
        n_idx, used_c = g.dev_index(d_c.data_ptr(), nc, d_rc.data_ptr(), len(recs) + 16)
        assert n_idx == len(recs) and used_c == nc
        g.dev_submit(d_c.data_ptr(), nc, d_rc.data_ptr(), n_idx)
        g.dev_finish()
        assert g.end_checksum() == end_c
    # recompress: idempotent; decompress: identity
    d_o = torch.empty(s.size + (1 << 20), dtype=torch.uint8, device="cuda")
    with GpuSnapshotStage("recompress") as g:
        g.dev_submit(d_c.data_ptr(), nc, d_rc.data_ptr(), n_idx, d_o.data_ptr(), d_o.numel())
        nr, _, _ = g.dev_finish()
        assert nr == nc and torch.equal(d_o[:nr], d_c[:nc])
    with GpuSnapshotStage("decompress") as g:
        g.dev_submit(d_c.data_ptr(), nc, d_rc.data_ptr(), n_idx, d_o.data_ptr(), d_o.numel())
        nd, _, _ = g.dev_finish()
        assert nd == s.size and torch.equal(d_o[:nd], d_in[:s.size])
    # one flipped bit in record k's payload -> ECKSUM at record k+1
    kbad = 9000
    pos = int(recs["off"][kbad]) + 312 + 77777
    d_in[pos] ^= 0x20
    with GpuSnapshotStage("verify") as g:
        g.dev_submit(d_in.data_ptr(), s.size, d_recs.data_ptr(), len(recs))
        with pytest.raises(MtzError) as ei:
            g.dev_finish()
        assert ei.value.code == ECKSUM and g.stats()["bad_record"] == kbad + 1


def test_wire_preamble_versioning(oracle):
    """the lz4-stage-v1 wire is framed -- a 32-byte preamble (magic, version, capability
    word) in front of every DRR_BEGIN of a COMPRESS output, outside the stream checksum.  DECOMPRESS
    speaks exactly the versions / capabilities it knows and inverts only what the stage produced;
    the other modes never see a preamble (the raw wire stays the reference's byte stream)."""
    from manatee_b200 import GpuSnapshotStage
    from manatee_b200._native import MtzError, EFORMAT, EINVAL
    two = np.concatenate([_mixed_stream(oracle, n=6, recsize=16384), _mixed_stream(oracle, n=3, recsize=8192)])
    rc, want, st = oracle.stream_compress(two)
    assert rc == 0
    W = oracle.WIRE_PRE_BYTES
    got, _, _ = _gpu("compress", two, batch_bytes=1 << 20)
    assert np.array_equal(got, want)
    # one preamble per sub-stream, each directly in front of a BEGIN (magic 0x2F5bacbac at +8 of it)
    at = [i for i in range(0, got.size - 8, 8) if got[i:i + 8].tobytes() == oracle.WIRE_MAGIC]
    assert len(at) == 2 and at[0] == 0
    for a in at:
        assert int.from_bytes(got[a + 8:a + 12].tobytes(), "little") == 1          # version
        assert int.from_bytes(got[a + W + 8:a + W + 16].tobytes(), "little") == 0x2F5bacbac
    back, _, _ = _gpu("decompress", got, cap=two.size + (1 << 20), batch_bytes=1 << 20)
    assert np.array_equal(back, two)
    cases = []
    v2 = got.copy(); v2[8] = 2
    cases.append((v2, EFORMAT))                                   # a newer wire version
    cap = got.copy(); cap[at[1] + 13] = 0x80
    cases.append((cap, EFORMAT))                                  # a capability bit this side does not know
    cases.append((oracle.wire_strip(got), EINVAL))                # not produced by the COMPRESS stage
    lone = np.concatenate([got[:W], got[W + 312:]])               # preamble not followed by BEGIN
    cases.append((lone, EFORMAT))
    for bad, code in cases:
        assert oracle.stream_decompress(bad)[0] == code
        with pytest.raises(MtzError) as ei:
            _gpu("decompress", bad, cap=two.size + (1 << 20))
        assert ei.value.code == code
        with GpuSnapshotStage("decompress", ring_bytes=4 << 20, out_ring_bytes=4 << 20, batch_bytes=1 << 20) as g:
            with pytest.raises(MtzError) as ei:
                g.write(bad)
                g.flush()
                while g.read(1 << 20) is not None:
                    pass
            assert ei.value.code == code
    # the wire framing is for the decompressing peer only
    for mode in ("verify", "recompress", "compress"):
        with pytest.raises(MtzError) as ei:
            _gpu(mode, got)
        assert ei.value.code == EFORMAT


def test_unaligned_logical_size_never_reaches_a_kernel(oracle):
    """lsize = 1001 on an LZ4 record: MTZ_EFORMAT from the bulk call, the ring API and
    the GPU-side parser -- never a misaligned store in k_assemble."""
    import torch
    from manatee_b200 import GpuSnapshotStage
    from manatee_b200._native import MtzError, EFORMAT
    from manatee_b200.stage import REC_DTYPE
    s = oracle.synth_stream(6, recsize=4096, kind=oracle.PAYLOAD_PGPAGE)
    rc, c, _ = oracle.stream_compress_plain(s)
    cnt, offs = oracle.stream_index(c)
    bad = c.copy()
    o = int(offs[3])
    bad[o + 32:o + 40] = np.frombuffer((1001).to_bytes(8, "little"), dtype=np.uint8)
    with pytest.raises(MtzError) as ei:
        _gpu("recompress", bad)
    assert ei.value.code == EFORMAT
    with GpuSnapshotStage("recompress", ring_bytes=4 << 20, out_ring_bytes=4 << 20, batch_bytes=1 << 20) as g:
        with pytest.raises(MtzError) as ei:
            g.write(bad); g.flush()
            while g.read(1 << 20) is not None:
                pass
        assert ei.value.code == EFORMAT
    d = torch.from_numpy(bad).cuda()
    d_recs = torch.zeros((cnt + 8) * 32, dtype=torch.uint8, device="cuda")
    with GpuSnapshotStage("verify") as g:
        with pytest.raises(MtzError) as ei:
            g.dev_index(d.data_ptr(), bad.size, d_recs.data_ptr(), cnt + 8)
        assert ei.value.code == EFORMAT
