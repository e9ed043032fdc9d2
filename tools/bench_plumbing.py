#!/usr/bin/env python
"""BASELINE configs[0] / BASELINE.md row B0': the restore plumbing end to end over
loopback TCP with a fake `zfs` -- POST /backup, `zfs send` child -> [stage] -> socket ->
[stage] -> `zfs recv` child, job polling -- through the Python mirror of the reference's
modules (Node is not installed, so the reference's own JS cannot be timed here).

  python tools/bench_plumbing.py [GiB] [off|verify|compress|cpump]

`cpump` is row B0' proper: tools/pump_pair.c, the two pipes of the reference restated in C (read <= 64 KiB
-> write, one thread per side, loopback TCP) between the same fake `zfs send` / `zfs recv` children
-- the ceiling of the reference's plumbing on a box without Node.js, labelled "restated".
"""
import hashlib, json, os, socket, stat, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import oracle as O
from manatee_b200.host import BackupSender, BackupServer, ZfsClient


def free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def cpump(zfs, env, s, want):
    """`zfs send` | pump_pair send  ==TCP==>  pump_pair recv | `zfs recv`"""
    import subprocess
    exe = os.path.join(ROOT, "tools", "pump_pair")
    if not os.path.exists(exe):
        subprocess.check_call(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "tools", "pump_pair.c")])
    port = free_port()
    t0 = time.perf_counter()
    recv = subprocess.Popen([exe, "recv", str(port)], stdout=subprocess.PIPE, env=env)
    zrecv = subprocess.Popen([zfs, "recv", "-v", "-u", "zones/y/data/manatee"], stdin=recv.stdout, env=env,
                             stderr=subprocess.DEVNULL)
    zsend = subprocess.Popen([zfs, "send", "-v", "-P", "zones/x/data/manatee@1"], stdout=subprocess.PIPE, env=env,
                             stderr=subprocess.DEVNULL)
    send = subprocess.Popen([exe, "send", "127.0.0.1", str(port)], stdin=zsend.stdout, env=env)
    rcs = [p.wait() for p in (zsend, send, recv, zrecv)]
    dt = time.perf_counter() - t0
    digest, n = open(env["FAKE_ZFS_RECV_OUT"]).read().split()
    return {"gib": round(s.size / 2**30, 3), "seconds": round(dt, 3), "stream_gibs": round(s.size / 2**30 / dt, 3),
            "identity": digest == want and int(n) == s.size, "exit_codes": rcs,
            "what": "restated plumbing (tools/pump_pair.c), NOT the reference's Node.js; includes child start-up"}


def main():
    gib = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    modes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["off"]
    d = tempfile.mkdtemp()
    zfs = os.path.join(d, "zfs")
    open(zfs, "w").write("#!/bin/sh\nexec %s %s \"$@\"\n" % (sys.executable, os.path.join(ROOT, "tools", "fake_zfs.py")))
    os.chmod(zfs, os.stat(zfs).st_mode | stat.S_IEXEC)
    nw = int(gib * 2**30) // 131384
    s = O.synth_stream(nw, kind=O.PAYLOAD_PGPAGE)
    sp = os.path.join(d, "stream.bin"); s.tofile(sp)
    want = hashlib.sha256(s.tobytes()).hexdigest()
    env = dict(os.environ, PATH=d + os.pathsep + os.environ["PATH"], FAKE_ZFS_STREAM=sp,
               FAKE_ZFS_RECV_OUT=os.path.join(d, "recv.out"))
    out = {}
    for mode in modes:
        if mode == "cpump":
            out[mode] = cpump(zfs, env, s, want)
            continue
        sg = rg = None
        if mode == "verify":
            sg = rg = {"mode": "verify"}
        elif mode == "compress":
            sg, rg = {"mode": "compress"}, {"mode": "decompress"}
        srv = BackupServer.start({"log": None, "port": 0, "host": "127.0.0.1"})
        snd = BackupSender.start({"log": None, "dataset": "zones/x/data/manatee", "zfsPath": zfs,
                                  "queue": srv.getQueue(), "gpu": sg, "env": env})
        cli = ZfsClient({"log": None, "dataset": "zones/y/data/manatee", "dbUser": "postgres",
                         "mountpoint": "/m", "pollInterval": 100, "zfsHost": "127.0.0.1", "zfsPath": zfs,
                         "zfsPort": free_port(), "gpu": rg, "env": env, "zfsBin": zfs, "zfsEnv": env})
        res = {}
        t0 = time.perf_counter()
        cli.restore("http://127.0.0.1:%d" % srv.port, lambda err, old: res.update(err=err))
        dt = time.perf_counter() - t0
        snd.join(10); srv.close()
        digest, n = open(env["FAKE_ZFS_RECV_OUT"]).read().split()
        out[mode] = {"gib": round(s.size / 2**30, 3), "seconds": round(dt, 3),
                     "stream_gibs": round(s.size / 2**30 / dt, 3), "identity": digest == want,
                     "err": str(res.get("err")) if res.get("err") else None,
                     "wire_bytes": (cli._restoreObject or {}).get("gpu", {}).get("bytes_out", s.size)}
    print(json.dumps({"plumbing": out, "note": "Python mirror of backupServer/backupSender/zfsClient, "
                      "fake zfs children, loopback TCP, includes job polling (100 ms) and child start-up"}))


if __name__ == "__main__":
    main()
