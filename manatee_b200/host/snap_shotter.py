"""SnapShotter -- mirror of lib/snapShotter.js (the step before the path).

It produces the snapshot names BackupSender._getLatestSnapshot picks from
(lib/backupSender.js:244-288: newest `<dataset>@<13 digits>`), and its cleanup races a
long `zfs send`: a snapshot that is being sent cannot be destroyed.

Kept from the reference:
  * every pollInterval ms (default 1000) `zfs snapshot <dataset>@<epoch ms>`, the first
    one immediately (:105-151); if healthUrl is set, GET <healthUrl>/ping first and skip
    the snapshot when the request fails or reports healthy == false (:110-131);
    a failed `zfs snapshot` is logged and swallowed (createSnapshot, :445-470)
  * cleanup (:175-433), immediately and then pollInterval after each run ends:
      dataset missing -> nothing, no error;
      list `-t snapshot -H -d 1 -s creation -o name`, keep only `^\\d{13}$` names
      (operator-made snapshots are never touched);
      if count >= snapshotNumber (default 10) destroy the oldest `count - snapshotNumber`;
      a destroy that fails marks the snapshot stuck and the NEXT oldest is tried
      instead; it is an error (emitted as 'error') once the stuck ones reach the
      number that had to go, or when every snapshot is stuck.
Additive: close() (the reference daemon never stops), and the synchronous
_createOnce()/_cleanupOnce() the timers call, so tests can step it.
"""
import json
import re
import threading
import time
import urllib.request

from . import zfs_cmd

RE_SNAPSHOT = re.compile(r"^([^@]+)@([^@]+)$")
RE_EPOCH_MS = re.compile(r"^\d{13}$")


class SnapShotter(object):
    def __init__(self, options):
        assert isinstance(options, dict), "options (object) is required"
        assert isinstance(options.get("dataset"), str), "options.dataset (string) is required"
        for k in ("pollInterval", "snapshotNumber"):
            assert options.get(k) is None or isinstance(options[k], (int, float)), \
                "options.%s (number) is optional" % k
        assert options.get("healthUrl") is None or isinstance(options["healthUrl"], str)
        self._zfsRuns = 0
        self._pollInterval = options.get("pollInterval") or 1 * 1000
        self._dataset = options["dataset"]
        self._snapshotNumber = options.get("snapshotNumber") or 10
        self._healthUrl = options.get("healthUrl")
        self._zfsBin = options.get("zfsBin") or zfs_cmd.ZFS_BIN
        self._zfsEnv = options.get("zfsEnv") or {}
        self._handlers = {}
        self._stop = threading.Event()
        self._threads = []
        self.lastCleanup = None                   # additive: summary of the last cleanup pass

    def on(self, event, fn):
        self._handlers.setdefault(event, []).append(fn)
        return self

    def emit(self, event, *args):
        for fn in self._handlers.get(event, []):
            fn(*args)

    # -- lib/snapShotter.js:577-612
    def _execZfs(self, opts):
        assert isinstance(opts.get("label"), str) and all(isinstance(a, str) for a in opts["args"])
        self._zfsRuns += 1
        t0 = time.monotonic()
        err, info = None, {}
        try:
            info = zfs_cmd.zfsExecCommon({"zfs": self._zfsBin, "env": self._zfsEnv}, opts["args"])
        except zfs_cmd.ZfsError as e:
            err, info = e, dict(e.info)
        info["duration_ms"] = int(round((time.monotonic() - t0) * 1000))
        return err, info

    # -- lib/snapShotter.js:445-470: errors are logged, never returned
    def createSnapshot(self, name):
        assert isinstance(name, str), "name (string) is required"
        return self._writeSnapshot(self._dataset + "@" + name)

    def _writeSnapshot(self, snapshot):
        assert RE_SNAPSHOT.match(snapshot), "invalid snapshot: " + snapshot
        err, _ = self._execZfs({"label": "write snapshot", "args": ["snapshot", snapshot]})
        return err

    def _deleteSnapshot(self, snapshot):
        assert RE_SNAPSHOT.match(snapshot), "invalid snapshot: " + snapshot
        err, _ = self._execZfs({"label": "delete snapshot", "args": ["destroy", snapshot]})
        return err

    # -- create(), lib/snapShotter.js:105-151
    def _createOnce(self):
        if self._healthUrl:
            try:
                with urllib.request.urlopen(self._healthUrl.rstrip("/") + "/ping", timeout=10) as r:
                    obj = json.loads(r.read().decode() or "null")
                if isinstance(obj, dict) and not obj.get("healthy"):
                    return False
            except Exception:                                 # noqa: BLE001  (503 included)
                return False
        self.createSnapshot(str(int(time.time() * 1000)))
        return True

    # -- cleanup(), lib/snapShotter.js:175-433 -> error or None
    def _cleanupOnce(self):
        summary = {"snapshots": 0, "ignored": 0, "deleted": [], "stuck": []}
        self.lastCleanup = summary
        try:
            if not zfs_cmd.zfsExists({"zfs": self._zfsBin, "env": self._zfsEnv, "dataset": self._dataset}):
                return None                                   # no_dataset: not an error
        except zfs_cmd.ZfsError as e:
            return e
        assert not RE_SNAPSHOT.match(self._dataset), self._dataset + " should not be a snapshot"
        err, info = self._execZfs({"label": "list snapshots for cleanup",
                                   "args": ["list", "-t", "snapshot", "-H", "-d", "1", "-s", "creation",
                                            "-o", "name", self._dataset]})
        if err:
            return err
        snapshots = []
        for line in info["stdout"].split("\n"):
            t = RE_SNAPSHOT.match(line)
            if not t:
                continue
            if not RE_EPOCH_MS.match(t.group(2)):
                summary["ignored"] += 1
                continue
            snapshots.append(line)
        summary["snapshots"] = len(snapshots)
        if len(snapshots) < self._snapshotNumber:
            return None
        excess = len(snapshots) - self._snapshotNumber
        if excess <= 0:
            return None
        deleted = 0
        for s in snapshots:                                   # oldest first
            if deleted >= excess:
                break
            derr = self._deleteSnapshot(s)
            if derr:
                summary["stuck"].append(s)
                nstuck = len(summary["stuck"])
                if nstuck >= len(snapshots):                  # nothing can be deleted at all
                    return derr
                if nstuck >= excess:                          # over the threshold and staying there
                    return derr
                continue
            deleted += 1
            summary["deleted"].append(s)
        return None

    # -- start(), lib/snapShotter.js:95-441
    def start(self, callback=None):
        def create_loop():
            while not self._stop.is_set():
                try:
                    self._createOnce()
                except Exception as e:                        # noqa: BLE001
                    self.emit("error", e)
                self._stop.wait(self._pollInterval / 1000.0)

        def cleanup_loop():
            while not self._stop.is_set():
                err = self._cleanupOnce()
                if err:
                    self.emit("error", err)
                self._stop.wait(self._pollInterval / 1000.0)

        for fn in (create_loop, cleanup_loop):
            t = threading.Thread(target=fn, daemon=True)
            t.start()
            self._threads.append(t)
        if callback:
            callback()

    def close(self):
        self._stop.set()
        for t in self._threads:
            t.join(10)
