#!/usr/bin/env python
"""Fake `zfs` for plumbing tests and the CPU plumbing baseline (SURVEY.md 4, 8d
config 1): selected through the reference's own `zfsPath` knob
(lib/backupSender.js:177, lib/zfsClient.js:793) and put first on PATH for the bare
`zfs list` of lib/backupSender.js:253.

  zfs list -t snapshot -H -d 1 -S name -o name <ds>   -> <ds>@<13 digits> lines
  zfs send -v -P <snap>    -> stream from $FAKE_ZFS_STREAM on stdout, the
                              `full/size/HH:MM:SS` progress protocol on stderr
  zfs recv -v -u <ds>      -> drains stdin, writes sha256 + byte count to $FAKE_ZFS_RECV_OUT
"""
import hashlib
import os
import sys
import time


def main():
    a = sys.argv[1:]
    if not a:
        return 2
    if a[0] == "list":
        ds = a[-1]
        if os.environ.get("FAKE_ZFS_NO_SNAPSHOTS"):
            return 0
        sys.stdout.write("%s@operator-made\n%s@1405378955344\n%s@1405378000000\n" % (ds, ds, ds))
        return 0
    if a[0] == "send":
        snap = a[-1]
        if os.environ.get("FAKE_ZFS_SEND_COUNT"):
            with open(os.environ["FAKE_ZFS_SEND_COUNT"], "a") as f:
                f.write("send\n")
        path = os.environ["FAKE_ZFS_STREAM"]
        size = os.path.getsize(path)
        sys.stderr.write("full\t%s\t%d\nsize\t%d\n" % (snap, size, size))
        sys.stderr.flush()
        sent = 0
        fail_at = int(os.environ.get("FAKE_ZFS_SEND_FAIL_AT", "-1"))
        out = sys.stdout.buffer
        with open(path, "rb") as f:
            while True:
                buf = f.read(1 << 20)
                if not buf:
                    break
                out.write(buf)
                sent += len(buf)
                sys.stderr.write("%s\t%d\t%s\n" % (time.strftime("%H:%M:%S"), sent, snap))
                sys.stderr.flush()
                if 0 <= fail_at <= sent:
                    sys.stderr.write("internal error: fake failure\n")
                    return 1
        out.flush()
        return 0
    if a[0] in ("recv", "receive"):
        h = hashlib.sha256()
        n = 0
        inp = sys.stdin.buffer
        while True:
            buf = inp.read(1 << 20)
            if not buf:
                break
            h.update(buf)
            n += len(buf)
        with open(os.environ["FAKE_ZFS_RECV_OUT"], "w") as f:
            f.write("%s %d\n" % (h.hexdigest(), n))
        sys.stderr.write("received %d bytes\n" % n)
        return 0
    return 0


if __name__ == "__main__":
    sys.exit(main())
