"""StatusServer `/restore` + the rebuild progress consumer.

StatusServer: mirror of lib/statusServer.js:30-121 restricted to what the path feeds:
    GET /         the endpoint list, text
    GET /restore  {"restore": <ZfsClient._restoreObject>}  ({} before any restore) :112-121
    GET /ping     200 <status> when healthy, 503 otherwise                          :82-103
`/state` (the ZooKeeper state machine's debug dump) is out of scope.  The reference
reaches the client as options.shard._pg._zfsClient; here the ZfsClient (anything with
a `_restoreObject`) is passed directly, and `ping` is an optional callable returning
the PostgreSQL manager's status dict.

RestoreWatcher: the consumer side, `manatee-adm rebuild`'s _watchSitter
(lib/adm.js:1550-1678): polls /restore once a second, sizes a progress bar from the
STRING `size`, advances it by the delta of the STRING `completed`, ends it on `done`,
and counts a changed job uuid as a new restore attempt (at most RESTORE_RETRIES).
Additive: throughput from consecutive polls and the stage's `gpu` stats object when
the job carries one -- fields the reference never reads.
"""
import json
import threading
import time
import urllib.request
from http.server import BaseHTTPRequestHandler, ThreadingHTTPServer

RESTORE_RETRIES = 5                       # lib/adm.js RESTORE_RETRIES


class StatusServer(object):
    def __init__(self, options):
        assert isinstance(options, dict), "options (object) is required"
        assert isinstance(options.get("port"), int), "options.port (number) is required"
        self._zfsClient = options.get("zfsClient")
        self._ping = options.get("ping")
        outer = self

        class H(BaseHTTPRequestHandler):
            def log_message(self, *a):                        # quiet
                pass

            def _send(self, code, obj, text=False):
                body = obj.encode() if text else json.dumps(obj, default=str).encode()
                self.send_response(code)
                self.send_header("Content-Type", "text/plain" if text else "application/json")
                self.send_header("Content-Length", str(len(body)))
                self.end_headers()
                self.wfile.write(body)

            def do_GET(self):
                path = self.path.split("?")[0]
                if path == "/":
                    return self._send(200, "/\n/ping\n/restore\n", text=True)
                if path == "/restore":
                    stat = {}
                    if outer._zfsClient is not None:
                        stat["restore"] = outer._zfsClient._restoreObject
                    return self._send(200, stat)
                if path == "/ping":
                    if outer._ping is None:
                        return self._send(503, "PG not inited")
                    stat = outer._ping()
                    return self._send(200 if stat.get("healthy") else 503, stat)
                return self._send(404, {"code": "ResourceNotFound", "message": path + " does not exist"})

        self._server = ThreadingHTTPServer((options.get("host", "127.0.0.1"), options["port"]), H)
        self.port = self._server.server_address[1]
        self._thread = threading.Thread(target=self._server.serve_forever, daemon=True)
        self._thread.start()

    @staticmethod
    def start(cfg):
        return StatusServer(cfg)

    def close(self):
        self._server.shutdown()
        self._server.server_close()


class RestoreWatcher(object):
    """State machine of lib/adm.js:1563-1666 without the terminal: feed it /restore
    bodies with observe(), read .events / .bar."""

    def __init__(self):
        self.bar = None                   # {'filename', 'size'|None, 'done_bytes'}
        self.lastByte = 0
        self.restoreTry = 1
        self.lastRestore = None
        self.events = []
        self.throughput = None            # additive: bytes/s between the last two polls
        self.gpu = None                   # additive: job.gpu
        self._lastT = None

    def observe(self, obj, now=None):
        now = time.monotonic() if now is None else now
        restore = obj.get("restore") if isinstance(obj, dict) else None
        if not restore:
            return None
        if self.lastRestore and "uuid" in self.lastRestore and self.lastRestore["uuid"] != restore.get("uuid"):
            if self.restoreTry >= RESTORE_RETRIES:
                raise RuntimeError("This Manatee instance is not an active peer after %d restore attempts.  "
                                   "Check sitter logs." % RESTORE_RETRIES)
            self.events.append(("retry", RESTORE_RETRIES - self.restoreTry))
            self.restoreTry += 1
        if self.bar is None and not restore.get("done"):
            size = restore.get("size")
            self.bar = {"filename": restore.get("dataset"), "size": int(size, 10) if size else None,
                        "done_bytes": 0}
            self.events.append(("bar", self.bar["size"]))
        if self.bar is not None and restore.get("completed"):
            completed = int(restore["completed"], 10)
            advance = completed - (self.lastByte or 0)
            if self._lastT is not None and now > self._lastT and advance >= 0:
                self.throughput = advance / (now - self._lastT)
            self.lastByte = completed
            self.bar["done_bytes"] += advance
            self._lastT = now
        if "gpu" in restore:
            self.gpu = restore["gpu"]
        if self.bar is not None and restore.get("done"):
            self.events.append(("end", self.bar["done_bytes"]))
            self.bar, self.lastByte = None, None
        self.lastRestore = restore
        return restore

    def watch(self, url, until, interval=1.0, timeout=60.0):
        """poll <url>/restore until `until(restore)` is true (the reference waits for the
        peer to come online, which is outside this path)"""
        t_end = time.monotonic() + timeout
        while time.monotonic() < t_end:
            with urllib.request.urlopen(url.rstrip("/") + "/restore", timeout=10) as r:
                restore = self.observe(json.loads(r.read().decode()))
            if restore is not None and until(restore):
                return restore
            time.sleep(interval)
        raise TimeoutError("restore did not finish")
