"""Committed digests (tests/golden/golden.json, section "digests", made by
tests/golden/make_golden.py): hashes of seeded streams that are too large to commit as files --
every record type, mixed compressible / incompressible / all-zero payloads, both LZ4 table
flavours -- and of what the oracle makes of them in every mode.
  CPU: the oracle still reproduces every digest (pins the checker against its own history:
       the vector Fletcher-4 and the faster decoder copies went in without moving one).
  GPU: the CUDA path reproduces the same digests through the C ABI, so parity is also held
       against COMMITTED values, not only against an oracle run in the same process."""
import hashlib
import json
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")
sys.path.insert(0, GOLD)


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_oracle_reproduces_committed_digests(oracle):
    import make_golden
    want = json.load(open(os.path.join(GOLD, "golden.json")))["digests"]
    got = make_golden.digests()
    assert sorted(got["streams"]) == sorted(want["streams"]) and sorted(got["blocks"]) == sorted(want["blocks"])
    for k in want["streams"]:
        assert got["streams"][k] == want["streams"][k], k
    for k in want["blocks"]:
        assert got["blocks"][k] == want["blocks"][k], k


CASE_NAMES = ["all_types_seed7", "all_types_seed31", "mixed_24x128k", "mixed_40x16k", "pcg_12x64k"]


def check_case_through_the_stage(name):
    """one digest case through VERIFY / COMPRESS / RECOMPRESS / DECOMPRESS of whatever library the
    stage is bound to (the GPU one here; tests/test_emul_library.py points it at the emulated one)"""
    import make_golden
    from test_gpu_codec import _gpu
    from manatee_b200 import GpuSnapshotStage
    w = json.load(open(os.path.join(GOLD, "golden.json")))["digests"]["streams"][name]
    s = dict(make_golden.digest_cases())[name]()
    assert _sha(s) == w["sha256"], name                         # same input as when the digest was made
    with GpuSnapshotStage("verify", batch_bytes=1 << 20) as g:
        g.process_host(s)
        assert ["%016x" % x for x in g.end_checksum()] == w["end_cksum"], name
        assert g.stats()["records"] == w["records"]
    c, gs, end = _gpu("compress", s, batch_bytes=1 << 20)
    assert _sha(c) == w["compress_sha256"] and c.size == w["compress_bytes"], name
    assert gs["lz4_encoded"] == w["compress_lz4"] and ["%016x" % x for x in end] == w["compress_end_cksum"]
    import oracle as O
    r, _, _ = _gpu("recompress", O.wire_strip(c))               # the send stream under the wire framing
    assert _sha(r) == w["recompress_sha256"], name
    d, _, _ = _gpu("decompress", c, cap=s.size + (1 << 20))
    assert _sha(d) == w["sha256"], name


def test_digest_case_list_matches_the_generator():
    import make_golden
    assert [n for n, _ in make_golden.digest_cases()] == CASE_NAMES


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASE_NAMES)
def test_gpu_reproduces_committed_digests(oracle, name):
    check_case_through_the_stage(name)
