"""CPU: size-independent properties of the oracle's stream walk over randomized streams
that use every DRR record type (the same hand-built stream the GPU parity tests use, with
varying seeds, plus multi-sub-stream concatenations).  These pin the checker itself:
what a corrupted byte must do, transport identity, idempotence, threaded == sequential."""
import numpy as np
import pytest

from test_gpu_codec import _all_types_stream      # builder only; nothing GPU is touched


def _records(oracle, s):
    cnt, offs = oracle.stream_index(s)
    assert cnt > 0
    return [int(o) for o in offs] + [int(s.size)]


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_every_corrupted_byte_is_caught_at_the_right_record(oracle, seed):
    s = _all_types_stream(oracle, seed=seed)
    offs = _records(oracle, s)
    rng = np.random.default_rng(100 + seed)
    assert oracle.stream_verify(s)[0] == 0
    for _ in range(60):
        pos = int(rng.integers(0, s.size))
        j = max(i for i in range(len(offs) - 1) if offs[i] <= pos)
        bad = s.copy()
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        rc, st = oracle.stream_verify(bad)
        inside = pos - offs[j]
        if rc == oracle.EFORMAT:
            assert inside < 312             # a header field that sizes the record
            continue
        assert rc == oracle.ECKSUM, (pos, j, inside, rc)
        # a record's header -- including its own checksum field -- is judged at that record
        # (BEGIN carries no checksum: caught by the next record); its payload by the next one
        own = inside < 312 and j != 0
        assert st.bad_record == (j if own else j + 1), (pos, j, inside, st.bad_record)


@pytest.mark.parametrize("seed", [4, 5])
def test_transport_identity_idempotence_and_threads(oracle, seed):
    s = _all_types_stream(oracle, seed=seed)
    rc, c, st = oracle.stream_compress(s)
    z = oracle.wire_strip(c)                 # the compressed send stream under the lz4-stage-v1 framing
    assert rc == 0 and oracle.stream_verify(z)[0] == 0 and c.size < s.size
    rc, d, _ = oracle.stream_decompress(c)
    assert rc == 0 and np.array_equal(d, s)
    rc, r, _ = oracle.stream_recompress(z)
    assert rc == 0 and np.array_equal(r, z)
    for nt in (1, 3, 8):
        rc, secs, g, _ = oracle.mt_recompress(z, nt)
        assert rc == 0 and np.array_equal(g, z)
        assert oracle.mt_verify(s, nt)[0] == 0 and oracle.mt_verify(z, nt)[0] == 0
    # the modes are only defined where they make sense: DECOMPRESS wants the stage's wire preamble,
    # COMPRESS refuses an already compressed stream (the host picks VERIFY for those, f2)
    assert oracle.stream_decompress(s)[0] == oracle.EINVAL
    assert oracle.stream_compress(z)[0] == oracle.EINVAL


def test_sub_streams_restart_the_checksum(oracle):
    a = _all_types_stream(oracle, seed=8)
    b = oracle.synth_stream(5, recsize=8192, kind=oracle.PAYLOAD_PGPAGE)
    two = np.concatenate([a, b, a])
    rc, st = oracle.stream_verify(two)
    assert rc == 0 and st.records == 2 * oracle.stream_index(a)[0] + oracle.stream_index(b)[0]
    assert st.end_cksum.tuple() == oracle.stream_verify(a)[1].end_cksum.tuple()
    rc, c, _ = oracle.stream_compress(two)
    rc2, d, _ = oracle.stream_decompress(c)
    assert rc == 0 and rc2 == 0 and np.array_equal(d, two)
    # a flipped bit in the middle stream is reported with the global record index
    bad = two.copy()
    pos = a.size + 312 + 312 + 100                      # inside b's OBJECT/first WRITE area
    bad[pos] ^= 4
    rc, st = oracle.stream_verify(bad)
    na, nb = oracle.stream_index(a)[0], oracle.stream_index(b)[0]
    assert rc == oracle.ECKSUM and na <= st.bad_record < na + nb
    assert oracle.mt_verify(bad, 4)[0] == oracle.ECKSUM
