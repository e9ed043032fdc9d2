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
    json.dump(meta, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
    print(json.dumps(meta, indent=1))


if __name__ == "__main__":
    main()
