"""Generates the committed golden fixtures under tests/golden/.

The reference (manatee 2.1.1) holds NO golden stream, checksum vector or
known-answer test for this path: these fixtures are produced by
the CPU oracle and pinned by hand-computable known answers (tests/test_oracle.py).
Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def digest_cases():
    """(name, builder) of seeded inputs too large to commit as files: only digests are stored.
    Shared with tests/test_zz_golden_digests.py, which rebuilds them and compares."""
    sys.path.insert(0, os.path.dirname(HERE))
    from test_gpu_codec import _all_types_stream, _mixed_stream
    return [
        ("all_types_seed7", lambda: _all_types_stream(O, seed=7)),
        ("all_types_seed31", lambda: _all_types_stream(O, seed=31)),
        ("mixed_24x128k", lambda: _mixed_stream(O, n=24, recsize=131072)),
        ("mixed_40x16k", lambda: _mixed_stream(O, n=40, recsize=16384)),
        ("pcg_12x64k", lambda: O.synth_stream(12, recsize=65536, kind=O.PAYLOAD_PCG)),
    ]


BLOCK_CASES = [(kind, idx, n) for kind in ("PAYLOAD_PGPAGE", "PAYLOAD_ZERO") for idx, n in
               [(1, 1024), (2, 8192), (3, 65546), (4, 65547), (5, 131072), (6, 1 << 20)]]


def digests():
    """Everything the oracle says about each case, as hashes: raw stream, END checksum,
    COMPRESS output (+ its END checksum and LZ4 record count), RECOMPRESS of that; and the
    ZFS-LZ4 frame of single blocks on both sides of every table-flavour boundary."""
    d = {"streams": {}, "blocks": {}}
    for name, build in digest_cases():
        s = build()
        rc, st = O.stream_verify(s)
        assert rc == 0
        rc, c, cst = O.stream_compress(s)
        assert rc == 0
        rc, r, rst = O.stream_recompress(O.wire_strip(c))     # of the send stream under the wire framing
        assert rc == 0
        d["streams"][name] = {
            "bytes": int(s.size), "records": int(st.records), "sha256": sha(s),
            "end_cksum": ["%016x" % x for x in st.end_cksum.tuple()],
            "compress_sha256": sha(c), "compress_bytes": int(c.size), "compress_lz4": int(cst.lz4_out),
            "compress_end_cksum": ["%016x" % x for x in cst.end_cksum.tuple()],
            "recompress_sha256": sha(r)}
    for kind, idx, n in BLOCK_CASES:
        p = O.gen_payload(getattr(O, kind), idx, n)
        ps, frame = O.zfs_lz4_compress(p)
        d["blocks"]["%s_%d_%d" % (kind, idx, n)] = {
            "payload_sha256": sha(p), "psize": int(ps), "frame_sha256": sha(frame[:ps] if ps < n else p),
            "fletcher4": ["%016x" % x for x in O.fletcher4(p)]}
    return d


def main():
    meta = {}
    s = O.synth_stream(8, recsize=4096, kind=O.PAYLOAD_PGPAGE)
    s.tofile(os.path.join(HERE, "stream_small.bin"))
    rc, st = O.stream_verify(s)
    meta["stream_small"] = {"bytes": int(s.size), "records": int(st.records),
                            "end_cksum": ["%016x" % x for x in st.end_cksum.tuple()],
                            "sha256": hashlib.sha256(s.tobytes()).hexdigest()}
    rc, c, st = O.stream_compress(s)
    c.tofile(os.path.join(HERE, "stream_small_lz4.bin"))
    meta["stream_small_lz4"] = {"bytes": int(c.size), "lz4_records": int(st.lz4_out),
                                "end_cksum": ["%016x" % x for x in st.end_cksum.tuple()],
                                "sha256": hashlib.sha256(c.tobytes()).hexdigest()}
    bad = s.copy()
    cnt, offs = O.stream_index(bad)
    bad[int(offs[5]) + 312 + 100] ^= 0x40
    bad.tofile(os.path.join(HERE, "stream_small_corrupt_5.bin"))
    rc, st = O.stream_verify(bad)
    meta["stream_small_corrupt_5"] = {"rc": rc, "bad_record": int(st.bad_record)}
    # one 128 KiB pg-page block and its frame (the 4096-slot / u32 table flavour)
    p = O.gen_payload(O.PAYLOAD_PGPAGE, 42, 131072)
    ps, frame = O.zfs_lz4_compress(p)
    meta["block_pgpage_42"] = {"psize": int(ps), "payload_sha256": hashlib.sha256(p.tobytes()).hexdigest(),
                               "frame_sha256": hashlib.sha256(frame.tobytes()).hexdigest(),
                               "fletcher4": ["%016x" % x for x in O.fletcher4(p)]}
    meta["digests"] = digests()
    json.dump(meta, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
