"""the steps either side of the bulk-data path -- snapshot cadence
and GC racing a long send (lib/snapShotter.js), the receiver's dataset lifecycle
(lib/zfsClient.js:115-221, 514-624 over lib/common.js zfs helpers), GET /restore
(lib/statusServer.js:112-121) and the rebuild progress consumer (lib/adm.js:1550-1678).
All against tools/fake_zfs.py's pool model; no GPU."""
import hashlib
import json
import os
import re
import socket
import stat
import subprocess
import sys
import threading
import time
import urllib.error
import urllib.request

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DS = "zones/y/data/manatee"


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.fixture()
def pool(tmp_path, oracle):
    z = tmp_path / "bin"
    z.mkdir()
    zfs = z / "zfs"
    zfs.write_text("#!/bin/sh\nexec %s %s \"$@\"\n" % (sys.executable, os.path.join(ROOT, "tools", "fake_zfs.py")))
    zfs.chmod(zfs.stat().st_mode | stat.S_IEXEC)
    stream = oracle.synth_stream(24, recsize=131072, kind=oracle.PAYLOAD_PGPAGE)
    sp = tmp_path / "stream.bin"
    stream.tofile(str(sp))
    env = {"PATH": str(z) + os.pathsep + os.environ.get("PATH", ""),
           "FAKE_ZFS_STREAM": str(sp), "FAKE_ZFS_RECV_OUT": str(tmp_path / "recv.out"),
           "FAKE_ZFS_STATE": str(tmp_path / "pool.json"), "FAKE_ZFS_SEND_COUNT": str(tmp_path / "sends")}

    class P(object):
        pass
    p = P()
    p.zfs, p.env, p.stream, p.tmp = str(zfs), env, stream, tmp_path

    def z_(*args, check=True, env_extra=None):
        e = dict(env)
        e.update(env_extra or {})
        r = subprocess.run([p.zfs] + list(args), env=e, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        if check:
            assert r.returncode == 0, r.stderr
        return r
    p.z = z_
    def state_():
        import fcntl
        with open(env["FAKE_ZFS_STATE"], "a+") as f:     # the fake rewrites it under the same lock
            fcntl.flock(f, fcntl.LOCK_EX)
            f.seek(0)
            raw = f.read()
        return json.loads(raw) if raw.strip() else {"datasets": {}, "held": [], "clock": 0}
    p.state = state_
    p.opts = lambda **kw: dict(kw, zfs=p.zfs, env=env)
    return p


def _client(pool, **over):
    from manatee_b200.host import ZfsClient
    o = {"log": None, "dataset": DS, "dbUser": "postgres", "mountpoint": "/manatee/pg", "pollInterval": 50,
         "zfsHost": "127.0.0.1", "zfsPath": pool.zfs, "zfsPort": _free_port(), "env": pool.env,
         "zfsBin": pool.zfs, "zfsEnv": pool.env}
    o.update(over)
    return ZfsClient(o)


def _serve(pool, env_extra=None, **sender_over):
    from manatee_b200.host import BackupSender, BackupServer
    env = dict(pool.env)
    env.update(env_extra or {})
    srv = BackupServer.start({"log": None, "port": 0, "host": "127.0.0.1"})
    o = {"log": None, "dataset": "zones/x/data/manatee", "zfsPath": pool.zfs, "queue": srv.getQueue(), "env": env}
    o.update(sender_over)
    sender = BackupSender.start(o)
    return srv, sender


# ------------------------------------------------------------------ lib/common.js helpers
def test_zfs_cmd_helpers_and_error_wrapping(pool):
    from manatee_b200.host import zfs_cmd as Z
    assert Z.zfsExists(pool.opts(dataset="zones/a")) is False
    Z.zfsCreate(pool.opts(dataset="zones/a", props={"canmount": "noauto"}))
    assert Z.zfsExists(pool.opts(dataset="zones/a")) is True
    Z.zfsSet(pool.opts(dataset="zones/a", property="mountpoint", value="/x"))
    assert Z.zfsGet(pool.opts(dataset="zones/a", property="mountpoint")) == "/x"
    assert Z.zfsGet(pool.opts(dataset="zones/a", property="mounted")) == "no"
    Z.zfsMount(pool.opts(dataset="zones/a"))
    assert Z.zfsGet(pool.opts(dataset="zones/a", property="mounted")) == "yes"
    Z.zfsInherit(pool.opts(dataset="zones/a", property="mountpoint"))
    assert Z.zfsGet(pool.opts(dataset="zones/a", property="mountpoint")) == "-"
    Z.zfsSnapshot(pool.opts(dataset="zones/a", snapshot="1405378955344"))
    Z.zfsUnmount(pool.opts(dataset="zones/a"))
    Z.zfsRename(pool.opts(dataset="zones/a", target="zones/iso/old/a", parents=True))
    assert Z.zfsExists(pool.opts(dataset="zones/iso/old/a")) and not Z.zfsExists(pool.opts(dataset="zones/a"))
    assert pool.state()["datasets"]["zones/iso/old/a"]["snapshots"][0][0] == "1405378955344"
    # failures carry the operation, then the cause with exit status and stderr (VError chain)
    with pytest.raises(Z.ZfsError) as ei:
        Z.zfsSet(pool.opts(dataset="zones/nope", property="canmount", value="off"))
    m = str(ei.value)
    assert m.startswith('set property "canmount" to "off" on dataset "zones/nope": exec ') and \
        "exited with status 1" in m and "does not exist" in m
    with pytest.raises(Z.ZfsError) as ei:
        Z.zfsRename(pool.opts(dataset="zones/iso/old/a", target="zones/q/r/a", parents=False))
    assert str(ei.value).startswith('rename dataset "zones/iso/old/a" to "zones/q/r/a": ')
    with pytest.raises(Z.ZfsError) as ei:
        Z.zfsDestroy(pool.opts(dataset="zones/iso/old/a"))              # has a snapshot, no -r
    assert str(ei.value).startswith('destroy dataset "zones/iso/old/a": ')
    Z.zfsDestroy(pool.opts(dataset="zones/iso/old/a", recursive=True))
    assert not Z.zfsExists(pool.opts(dataset="zones/iso/old/a"))
    with pytest.raises(AssertionError):
        Z.zfsSet(pool.opts(dataset="zones/a", property="canmount"))      # opts.value (string) is required
    with pytest.raises(Z.ZfsError):
        Z.zfsExists({"dataset": "zones/a", "zfs": "/nonexistent/zfs"})


# ------------------------------------------------------------------ restore() lifecycle
def test_restore_into_empty_pool_sets_properties_mounts_and_snapshots(pool):
    srv, sender = _serve(pool)
    cli = _client(pool)
    res = {}
    cli.restore("http://127.0.0.1:%d" % srv.port, lambda err, old: res.update(err=err, old=old))
    sender.join(10); srv.close()
    assert res["err"] is None and res["old"] is None           # nothing to isolate
    d = pool.state()["datasets"][DS]
    assert d["props"]["canmount"] == "noauto" and d["props"]["mountpoint"] == "/manatee/pg"
    assert "snapdir" not in d["props"] and d["mounted"] is True
    names = [s[0] for s in d["snapshots"]]
    assert len(names) == 2 and all(re.match(r"^\d{13}$", n) for n in names)   # received + initial
    assert abs(int(names[-1]) - time.time() * 1000) < 60000
    digest, n = open(pool.env["FAKE_ZFS_RECV_OUT"]).read().split()
    assert digest == hashlib.sha256(pool.stream.tobytes()).hexdigest()


def test_restore_isolates_the_existing_dataset_first(pool):
    pool.z("create", "-o", "mountpoint=/manatee/pg", DS)
    pool.z("snapshot", DS + "@1400000000000")
    assert pool.state()["datasets"][DS]["mounted"] is True
    srv, sender = _serve(pool)
    cli = _client(pool)
    res = {}
    cli.restore("http://127.0.0.1:%d" % srv.port, lambda err, old: res.update(err=err, old=old))
    sender.join(10); srv.close()
    assert res["err"] is None
    assert re.match(r"^zones/y/data/isolated/autorebuild-\d{4}-\d\d-\d\dT\d\d:\d\d:\d\d\.\d{3}Z$", res["old"])
    st = pool.state()["datasets"]
    old = st[res["old"]]
    assert old["props"]["canmount"] == "off" and old["mounted"] is False and "mountpoint" not in old["props"]
    assert [s[0] for s in old["snapshots"]] == ["1400000000000"]         # preserved, not destroyed
    assert "zones/y/data/isolated" in st                                 # rename -p made the holding area
    assert st[DS]["mounted"] is True and st[DS]["props"]["canmount"] == "noauto"


def test_busy_dataset_fails_before_any_byte_is_requested(pool):
    pool.z("create", DS)
    st = pool.state(); st["datasets"][DS]["busy"] = True
    json.dump(st, open(pool.env["FAKE_ZFS_STATE"], "w"))
    srv, sender = _serve(pool)
    cli = _client(pool)
    res = {}
    cli.restore("http://127.0.0.1:%d" % srv.port, lambda err, old: res.update(err=err, old=old))
    srv.close()
    m = str(res["err"])
    assert m.startswith('receiving snapshot from "http://127.0.0.1:') and 'preserving dataset "%s"' % DS in m \
        and "Device busy" in m
    assert res["old"] is None
    assert not os.path.exists(pool.env["FAKE_ZFS_SEND_COUNT"])           # no POST, no zfs send
    assert pool.state()["datasets"][DS]["mounted"] is True               # left as it was


def test_mounted_check_after_canmount_off(pool):
    """isolateDataset refuses to rename when `mounted` is not "no" (lib/zfsClient.js:576-587)"""
    from manatee_b200.host import zfs_cmd as Z
    pool.z("create", DS)
    cli = _client(pool)
    real = Z.zfsGet
    try:
        Z.zfsGet = lambda opts: "yes"
        with pytest.raises(Z.ZfsError) as ei:
            cli.isolateDataset({"prefix": "rebuild"})
    finally:
        Z.zfsGet = real
    assert 'wanted "no" but found "yes" for property "mounted"' in str(ei.value)
    assert DS in pool.state()["datasets"]                                 # not renamed


def test_failed_receive_still_reports_the_isolated_dataset(pool):
    pool.z("create", DS)
    srv, sender = _serve(pool, env_extra={"FAKE_ZFS_SEND_FAIL_AT": str(1 << 20)})
    cli = _client(pool)
    res = {}
    cli.restore("http://127.0.0.1:%d" % srv.port, lambda err, old: res.update(err=err, old=old))
    sender.join(10); srv.close()
    assert res["err"] is not None and res["old"] is not None and res["old"].startswith("zones/y/data/isolated/")
    assert res["old"] in pool.state()["datasets"]                         # the operator can roll back


# ------------------------------------------------------------------ /restore + progress consumer
def test_status_server_restore_and_watcher(pool):
    from manatee_b200.host import StatusServer, RestoreWatcher
    srv, sender = _serve(pool, env_extra={"FAKE_ZFS_SEND_DELAY": "0.15"})
    cli = _client(pool)
    health = {"healthy": False}
    ss = StatusServer.start({"log": None, "port": 0, "zfsClient": cli, "ping": lambda: dict(health)})
    base = "http://127.0.0.1:%d" % ss.port
    try:
        assert json.loads(urllib.request.urlopen(base + "/restore").read()) == {"restore": None}
        assert urllib.request.urlopen(base + "/").read().decode().split() == ["/", "/ping", "/restore"]
        with pytest.raises(urllib.error.HTTPError) as ei:
            urllib.request.urlopen(base + "/ping")
        assert ei.value.code == 503
        health["healthy"] = True
        assert json.loads(urllib.request.urlopen(base + "/ping").read())["healthy"] is True
        res = {}
        t = threading.Thread(target=cli.restore, args=("http://127.0.0.1:%d" % srv.port,
                                                       lambda err, old: res.update(err=err, old=old)))
        t.start()
        w = RestoreWatcher()
        last = w.watch(base, until=lambda r: r.get("done") is True, interval=0.05, timeout=60)
        t.join(60)
        assert res["err"] is None
        size = pool.stream.size
        assert last["size"] == str(size) and last["completed"] == str(size)      # strings, like the reference
        # the first poll can land before `zfs send -v` printed the size: a nosize bar, as in the reference
        assert w.events[0] in (("bar", size), ("bar", None)) and w.events[-1] == ("end", size)
        assert w.throughput is not None and w.restoreTry == 1
    finally:
        sender.join(10); srv.close(); ss.close()


def test_watcher_counts_restore_attempts():
    from manatee_b200.host import RestoreWatcher
    from manatee_b200.host.status_server import RESTORE_RETRIES
    w = RestoreWatcher()
    assert w.observe({}) is None and w.observe({"restore": None}) is None
    w.observe({"restore": {"uuid": "a", "dataset": "d", "done": False}}, now=0.0)
    assert w.events == [("bar", None)]                                       # nosize bar
    w.observe({"restore": {"uuid": "a", "dataset": "d", "done": 0, "size": "100", "completed": "40"}}, now=1.0)
    w.observe({"restore": {"uuid": "a", "dataset": "d", "done": 0, "size": "100", "completed": "70",
                           "gpu": {"records": 3}}}, now=2.0)
    assert w.bar["done_bytes"] == 70 and w.throughput == 30.0 and w.gpu == {"records": 3}
    w.observe({"restore": {"uuid": "a", "dataset": "d", "done": True, "size": "100", "completed": "100"}}, now=3.0)
    assert w.events[-1] == ("end", 100) and w.bar is None
    for k in range(RESTORE_RETRIES - 1):                                     # sitter keeps restarting the restore
        w.observe({"restore": {"uuid": "b%d" % k, "dataset": "d", "done": False}})
    assert w.restoreTry == RESTORE_RETRIES and [e for e in w.events if e[0] == "retry"][-1] == ("retry", 1)
    with pytest.raises(RuntimeError) as ei:
        w.observe({"restore": {"uuid": "zz", "dataset": "d", "done": False}})
    assert "not an active peer after %d restore attempts" % RESTORE_RETRIES in str(ei.value)


# ------------------------------------------------------------------ SnapShotter
def _snaps(pool, ds):
    return [s[0] for s in sorted(pool.state()["datasets"][ds]["snapshots"], key=lambda s: s[1])]


def _shotter(pool, ds, **over):
    from manatee_b200.host import SnapShotter
    o = {"log": None, "dataset": ds, "zfsBin": pool.zfs, "zfsEnv": pool.env}
    o.update(over)
    return SnapShotter(o)


def test_snapshotter_names_and_retention(pool):
    ds = "zones/x/data/manatee"
    s = _shotter(pool, ds, snapshotNumber=3)
    assert s._cleanupOnce() is None and s.lastCleanup["snapshots"] == 0      # no dataset yet: quiet
    s._createOnce()                                                          # failure is swallowed
    pool.z("create", ds)
    pool.z("snapshot", ds + "@operator-made")
    for k in range(5):
        assert s.createSnapshot(str(1405378955000 + k)) is None
    assert s._createOnce() is True
    names = _snaps(pool, ds)
    assert len(names) == 7 and re.match(r"^\d{13}$", names[-1]) and abs(int(names[-1]) - time.time() * 1000) < 60000
    assert s._cleanupOnce() is None
    c = s.lastCleanup
    assert c["snapshots"] == 6 and c["ignored"] == 1 and len(c["deleted"]) == 3 and not c["stuck"]
    left = _snaps(pool, ds)
    assert left[0] == "operator-made" and left[1:] == names[4:]               # oldest three epoch names went
    assert s._cleanupOnce() is None and s.lastCleanup["deleted"] == []       # at the threshold: nothing to do


def test_snapshot_being_sent_is_stuck_and_the_next_oldest_goes(pool):
    ds = "zones/x/data/manatee"
    pool.z("create", ds)
    s = _shotter(pool, ds, snapshotNumber=2)
    for k in range(5):
        s.createSnapshot(str(1405378955000 + k))
    st = pool.state(); st["held"] = [ds + "@1405378955000"]                  # a `zfs send` has it open
    json.dump(st, open(pool.env["FAKE_ZFS_STATE"], "w"))
    assert s._cleanupOnce() is None
    assert s.lastCleanup["stuck"] == [ds + "@1405378955000"]
    assert _snaps(pool, ds) == ["1405378955000", "1405378955004"]            # 3 others deleted instead
    # stuck count reaches the number that had to go -> error, emitted by the daemon loop
    s.createSnapshot("1405378955005")
    err = s._cleanupOnce()
    assert err is not None and "dataset is busy" in str(err) and s.lastCleanup["deleted"] == []
    # every snapshot stuck
    st = pool.state(); st["held"] = [ds + "@" + n for n in _snaps(pool, ds)]
    json.dump(st, open(pool.env["FAKE_ZFS_STATE"], "w"))
    s2 = _shotter(pool, ds, snapshotNumber=1)
    assert s2._cleanupOnce() is not None and len(s2.lastCleanup["stuck"]) >= 2


def test_snapshotter_daemon_health_gate_and_errors(pool):
    from manatee_b200.host import StatusServer
    ds = "zones/x/data/manatee"
    pool.z("create", ds)
    health = {"healthy": False}
    ss = StatusServer.start({"log": None, "port": 0, "ping": lambda: dict(health)})
    s = _shotter(pool, ds, pollInterval=40, snapshotNumber=3, healthUrl="http://127.0.0.1:%d" % ss.port)
    errs = []
    s.on("error", errs.append)
    try:
        started = []
        s.start(lambda: started.append(1))
        assert started == [1]
        time.sleep(0.5)
        assert _snaps(pool, ds) == []                                        # unhealthy: no snapshots
        health["healthy"] = True
        t_end = time.time() + 20
        while time.time() < t_end and not s.lastCleanup["deleted"]:
            time.sleep(0.05)
        assert s.lastCleanup["deleted"], "retention never kicked in"
    finally:
        s.close(); ss.close()
    names = _snaps(pool, ds)
    assert 1 <= len(names) <= 5 and all(re.match(r"^\d{13}$", n) for n in names) and not errs


def test_sender_ships_the_snapshotters_newest_and_gc_cannot_take_it_mid_send(pool):
    """f3 end to end: SnapShotter names feed _getLatestSnapshot (newest 13-digit name wins
    over operator snapshots); while that send runs the snapshot is held, cleanup skips it."""
    ds = "zones/x/data/manatee"
    pool.z("create", ds)
    s = _shotter(pool, ds, snapshotNumber=1)
    for k in range(3):
        s.createSnapshot(str(1405378955000 + k))
    pool.z("snapshot", ds + "@zzz-operator")
    srv, sender = _serve(pool, env_extra={"FAKE_ZFS_SEND_DELAY": "1.5"})
    cli = _client(pool)
    res = {}
    t = threading.Thread(target=cli.restore, args=("http://127.0.0.1:%d" % srv.port,
                                                   lambda err, old: res.update(err=err, old=old)))
    t.start()
    newest = ds + "@1405378955002"
    t_end = time.time() + 20
    while time.time() < t_end and newest not in pool.state()["held"]:
        time.sleep(0.02)
    assert newest in pool.state()["held"], "send never started"
    # excess = 2: the two older ones go, the held newest is never tried
    assert s._cleanupOnce() is None and s.lastCleanup["stuck"] == []
    s.createSnapshot("1405378955003"); s.createSnapshot("1405378955004")
    # now the held one is the oldest: it is stuck, and the retention count is met by
    # destroying the next oldest ones instead (lib/snapShotter.js:282-330)
    assert s._cleanupOnce() is None
    assert s.lastCleanup["stuck"] == [newest]
    assert s.lastCleanup["deleted"] == [ds + "@1405378955003", ds + "@1405378955004"]
    t.join(60)
    sender.join(10); srv.close()
    assert res["err"] is None
    assert cli._restoreObject["done"] is True
    assert newest not in pool.state()["held"]
    s.createSnapshot("1405378955005")
    assert s._cleanupOnce() is None and s.lastCleanup["deleted"] == [newest]  # released after the send


# ------------------------------------------------------------------ daemon entry points
def _spawn(module, cfg_path, env):
    p = subprocess.Popen([sys.executable, "-m", module, "-f", cfg_path], cwd=ROOT, env=env,
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    line = p.stdout.readline()                       # one JSON line once it is up
    return p, (json.loads(line) if line.strip() else None)


def test_backupserver_entry_point_serves_a_restore(pool):
    """backupserver.js mirror: `-f <config>` with the reference's config shape
    (test/etc/backupserver.json) starts REST server + sender on one queue."""
    cfg = {"backupServerCfg": {"port": 0, "host": "127.0.0.1"},
           "backupSenderCfg": {"zfsPath": pool.zfs, "dataset": "zones/x/data/manatee", "env": pool.env}}
    path = str(pool.tmp / "backupserver.json")
    json.dump(cfg, open(path, "w"))
    env = dict(os.environ, **pool.env)
    p, hello = _spawn("manatee_b200.host.backupserver", path, env)
    try:
        assert hello and hello["name"] == "manatee-backupserver" and hello["port"] > 0, p.stderr.read()
        cli = _client(pool)
        res = {}
        cli.restore("http://127.0.0.1:%d" % hello["port"], lambda err, old: res.update(err=err, old=old))
        assert res["err"] is None, res
        digest, n = open(pool.env["FAKE_ZFS_RECV_OUT"]).read().split()
        assert digest == hashlib.sha256(pool.stream.tobytes()).hexdigest()
        assert cli._restoreObject["done"] is True and "gpu" not in cli._restoreObject    # gpu key absent: legacy
    finally:
        p.terminate(); p.wait(10)
    # unreadable configuration is fatal, like process.abort() in the reference
    bad = str(pool.tmp / "broken.json")
    open(bad, "w").write("{ not json")
    r = subprocess.run([sys.executable, "-m", "manatee_b200.host.backupserver", "-f", bad], cwd=ROOT, env=env,
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=60)
    assert r.returncode == 134 and "Unable to read/parse configuration file" in r.stderr


def test_snapshotter_entry_point(pool):
    ds = "zones/x/data/manatee"
    pool.z("create", ds)
    cfg = {"dataset": ds, "pollInterval": 50, "snapshotNumber": 3, "zfsBin": pool.zfs, "zfsEnv": pool.env}
    path = str(pool.tmp / "snapshotter.json")
    json.dump(cfg, open(path, "w"))
    p, hello = _spawn("manatee_b200.host.snapshotter", path, dict(os.environ, **pool.env))
    try:
        assert hello and hello["dataset"] == ds, p.stderr.read()
        t_end = time.time() + 30
        seen_max = 0
        while time.time() < t_end:
            n = len(_snaps(pool, ds))
            seen_max = max(seen_max, n)
            if seen_max >= 4 and n <= 4:             # retention has had to delete something
                break
            time.sleep(0.1)
        assert seen_max >= 3
    finally:
        p.terminate(); p.wait(10)
    names = _snaps(pool, ds)
    assert names and all(re.match(r"^\d{13}$", n) for n in names) and len(names) <= 6
