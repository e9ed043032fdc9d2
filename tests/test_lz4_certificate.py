"""RECOMPRESS's certificate (kernels_lz4.cuh warp_lz4_certify, K3c): a record whose incoming LZ4
block is PROVEN to be what the declared encoder (oracle/lz4_zfs.c) emits for the decoded bytes is
passed through; every other record is re-encoded by the serial matcher.  Either way the output must
be the oracle's RECOMPRESS, bit for bit.  The dangerous direction is a false certificate (a block
that is valid LZ4, decodes to the same bytes, but is not the encoder's parse): tests/lz4_frames.py
makes such blocks by changing one parse decision at a time -- a split match, a match one byte
shorter or starting one byte later, another valid offset, a match turned into literals, a closing
token with a stray low nibble -- and the counts must show that none of them was certified while
every untouched record was.  CPU: the whole library on the SIMT emulator; `-m gpu`: the product."""
import os
import subprocess
import sys

import numpy as np
import pytest

import lz4_frames as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emul_library(emul_so):
    from manatee_b200 import _native as N
    saved = (N.SO_PATH, N._lib)
    N.SO_PATH, N._lib = emul_so, None
    try:
        yield N.lib()
    finally:
        N.SO_PATH, N._lib = saved


def _canonical(oracle, n, recsize, kind=None):
    raw = oracle.synth_stream(n, recsize, oracle.PAYLOAD_PGPAGE if kind is None else kind)
    rc, c, _ = oracle.stream_compress_plain(raw)
    assert rc == 0
    return np.ascontiguousarray(c), raw.size


def _recompress(src, cap, **kw):
    from manatee_b200 import GpuSnapshotStage
    with GpuSnapshotStage("recompress", **kw) as g:
        out = np.empty(cap, dtype=np.uint8)
        n = g.process_host(src, out)
        return out[:n].copy(), g.stats(), g.end_checksum()


def _check(oracle, src, cap, want_certified, **kw):
    rc, want, st = oracle.stream_recompress(src)
    assert rc == 0
    got, gs, end = _recompress(src, cap, **kw)
    assert np.array_equal(got, want)
    assert end == st.end_cksum.tuple()
    if want_certified is not None:
        assert gs["lz4_certified"] == want_certified, gs
    return gs


def _mutated(oracle, c, kinds, rng, every=2):
    blocks = F.blocks_of(oracle, c)
    changes = {}
    for i, (w, blk) in enumerate(blocks):
        if i % every:
            continue
        kind = kinds[int(rng.integers(len(kinds)))]
        nb = F.mutate(blk, kind, rng)
        if nb is not None and nb != blk:
            assert F.decode(*F.parse(nb)) == F.decode(*F.parse(blk))
            changes[w] = nb
    return F.splice(oracle, c, changes), len(blocks) - len(changes), len(changes)


def _all_cases(oracle, run, small=False):
    from manatee_b200 import _native as N
    # every record of an encoder-made stream is certified, in all three table flavours
    for n, recsize in (((3, 131072), (8, 16384), (2, 65536), (2, 262144)) if small else
                       ((5, 131072), (12, 16384), (4, 65536), (3, 262144))):
        c, cap = _canonical(oracle, n, recsize)
        gs = run(oracle, c, cap + (1 << 20), n)
        assert gs["lz4_encoded"] == n
    # MTZ_FLAG_REENCODE_ALL: nothing certified, same bytes
    nf = 2 if small else 6
    c, cap = _canonical(oracle, nf, 131072)
    gs = run(oracle, c, cap + (1 << 20), 0, flags=N.FLAG_REENCODE_ALL)
    assert gs["lz4_encoded"] == nf
    # one parse decision changed in every second record
    rng = np.random.default_rng(20260921)
    for kind in F.MUTATIONS:
        for n, recsize in (((2, 131072), (6, 16384)) if small else ((6, 131072), (10, 16384))):
            c, cap = _canonical(oracle, n, recsize)
            m, untouched, touched = _mutated(oracle, c, (kind,), rng)
            assert touched > 0, kind
            run(oracle, m, cap + (1 << 20), untouched)


def _shaped(oracle, n, recsize, seed):
    """pg-page records with stretches that stress the replay: tens of kilobytes of noise (literal runs
    of thousands of attempts, step > 1, 64-attempt rounds and their clashes), long zero runs (match
    lengths with many 255 extension bytes, overlapping offset 1), short periodic patterns."""
    rng = np.random.default_rng(seed)
    raw = oracle.synth_stream(n, recsize, oracle.PAYLOAD_PGPAGE).copy()
    cnt, offs = oracle.stream_index(raw)
    for k in range(cnt):
        o = int(offs[k])
        if int(raw[o]) != 3:
            continue
        pay = raw[o + 312:o + 312 + recsize]
        a = int(rng.integers(0, recsize // 4)); ln = int(rng.integers(recsize // 16, recsize // 3))
        pay[a:a + ln] = rng.integers(0, 256, min(ln, recsize - a), dtype=np.uint8)
        b = int(rng.integers(recsize // 2, recsize - recsize // 8)); lz = int(rng.integers(300, recsize // 8))
        pay[b:b + lz] = 0
        c = int(rng.integers(0, recsize - 700))
        pay[c:c + 600] = np.resize(np.frombuffer(b"abcdefg", dtype=np.uint8), 600)
    assert oracle.stream_restamp(raw)[0] == 0
    rc, cs, _ = oracle.stream_compress_plain(raw)
    assert rc == 0
    return np.ascontiguousarray(cs), raw.size


def _shaped_cases(oracle, run, sizes):
    rng = np.random.default_rng(5)
    for n, recsize in sizes:
        c, cap = _shaped(oracle, n, recsize, seed=recsize + n)
        nblk = len(F.blocks_of(oracle, c))
        assert nblk >= n - 1                                   # (a record may end up stored raw)
        run(oracle, c, cap + (1 << 20), nblk)
        m, untouched, touched = _mutated(oracle, c, F.MUTATIONS, rng)
        assert touched > 0
        run(oracle, m, cap + (1 << 20), untouched)


def test_certificate_on_the_emulated_library(emul_library, oracle):
    _all_cases(oracle, _check, small=True)


def test_long_literal_runs_and_long_matches_on_the_emulated_library(emul_library, oracle):
    _shaped_cases(oracle, _check, ((4, 131072), (6, 32768)))


@pytest.mark.parametrize("seed", [1, 2])
def test_certificate_fuzz_on_the_emulated_library(emul_library, oracle, seed):
    """random parse changes, several per block, small records so that thousands of decisions are covered"""
    rng = np.random.default_rng(seed)
    c, cap = _canonical(oracle, 48, 8192)
    blocks = F.blocks_of(oracle, c)
    changes = {}
    for w, blk in blocks:
        if rng.integers(4) == 0:
            continue
        nb = blk
        for _ in range(1 + int(rng.integers(3))):
            t = F.mutate(nb, F.MUTATIONS[int(rng.integers(len(F.MUTATIONS)))], rng)
            nb = t if t is not None else nb
        if nb != blk:
            changes[w] = nb
    m = F.splice(oracle, c, changes)
    _check(oracle, m, cap + (1 << 20), len(blocks) - len(changes), batch_bytes=64 << 10)


def test_incompressible_and_foreign_records_take_the_encoder(emul_library, oracle):
    """records stored raw have no block to certify; a literals-only block is valid LZ4 but not what
    the encoder emits for compressible bytes"""
    raw = oracle.synth_stream(3, 131072, oracle.PAYLOAD_PCG)
    rc, c, _ = oracle.stream_compress_plain(raw)
    _check(oracle, np.ascontiguousarray(c), raw.size + (1 << 20), 0)
    c, cap = _canonical(oracle, 4, 16384)
    blocks = F.blocks_of(oracle, c)
    w, blk = blocks[1]
    data = F.decode(*F.parse(blk))
    m = F.splice(oracle, c, {w: F.emit([], data)})
    # (the literals-only frame is bigger than the record: the header keeps drr_compressiontype lz4)
    _check(oracle, m, cap + (1 << 20), len(blocks) - 1)


def _many_tiny_sequences(oracle, recsize):
    """a record whose incoming block has more matches than K2's parse table holds (lsize / 8): one
    literal + a 4-byte match, over and over -- valid LZ4, decodes fine, must simply not be certified"""
    seqs = [(b"abcdefgh", 8, 4)]
    pos = 12
    while pos + 5 + 16 <= recsize:
        seqs.append((bytes([65 + (pos * 7) % 23]), 5, 4))
        pos += 5
    tail = bytes(range(48, 48 + recsize - pos))
    data = F.decode(seqs, tail)
    assert len(data) == recsize and len(seqs) > recsize // 8
    raw = oracle.synth_stream(3, recsize, oracle.PAYLOAD_PGPAGE).copy()
    cnt, offs = oracle.stream_index(raw)
    writes = [int(offs[k]) for k in range(cnt) if int(raw[int(offs[k])]) == 3]
    raw[writes[1] + 312:writes[1] + 312 + recsize] = np.frombuffer(data, dtype=np.uint8)
    assert oracle.stream_restamp(raw)[0] == 0
    rc, c, _ = oracle.stream_compress_plain(raw)
    c = np.ascontiguousarray(c)
    blocks = F.blocks_of(oracle, c)
    assert [w for w, _ in blocks] == [0, 1, 2]
    return F.splice(oracle, c, {1: F.emit(seqs, tail)}), raw.size, len(blocks)


def test_parse_table_overflow_is_not_certified(emul_library, oracle):
    for recsize in (16384, 131072):
        m, cap, nblk = _many_tiny_sequences(oracle, recsize)
        _check(oracle, m, cap + (1 << 20), nblk - 1)


def test_certificate_off_gives_the_same_bytes(emul_so, oracle, tmp_path):
    """MTZ_CERTIFY=0: every record re-encoded, same output, nothing certified"""
    c, cap = _canonical(oracle, 6, 16384)
    rc, want, st = oracle.stream_recompress(c)
    p = tmp_path / "in.bin"
    c.tofile(p)
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from manatee_b200 import _native as N; "
            "N.SO_PATH, N._lib = %r, None; from manatee_b200 import GpuSnapshotStage; "
            "src = np.fromfile(%r, dtype=np.uint8); out = np.empty(%d, dtype=np.uint8); "
            "g = GpuSnapshotStage('recompress'); n = g.process_host(src, out); s = g.stats(); g.close(); "
            "out[:n].tofile(%r); print(s['lz4_certified'], s['lz4_encoded'])"
            % (ROOT, emul_so, str(p), cap + (1 << 20), str(tmp_path / "out.bin")))
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, MTZ_CERTIFY="0"),
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert r.stdout.split() == ["0", "6"]
    assert np.array_equal(np.fromfile(tmp_path / "out.bin", dtype=np.uint8), want)


@pytest.mark.gpu
def test_certificate_on_the_gpu(oracle):
    _all_cases(oracle, _check)


@pytest.mark.gpu
def test_long_literal_runs_and_long_matches_on_the_gpu(oracle):
    _shaped_cases(oracle, _check, ((96, 131072), (64, 32768), (8, 1 << 20)))


@pytest.mark.gpu
def test_certificate_fuzz_on_the_gpu(oracle):
    rng = np.random.default_rng(7)
    c, cap = _canonical(oracle, 256, 131072)
    m, untouched, touched = _mutated(oracle, c, F.MUTATIONS, rng, every=3)
    assert touched > 50
    _check(oracle, m, cap + (1 << 20), untouched, batch_bytes=8 << 20)
