"""BackupQueue -- mirror of lib/backupQueue.js (array + 'push' event, never evicts)."""
import threading


class BackupQueue(object):
    """push() emits 'push' (lib/backupQueue.js:56-67); get(uuid) is a linear scan
    (:96-110); pop() exists but nothing calls it, so jobs are never evicted
    (:78-83) -- the poller and `manatee-adm rebuild` rely on old jobs staying
    readable."""

    def __init__(self, options=None):
        self._queue = []
        self._listeners = {}
        self._lock = threading.Lock()

    def on(self, event, fn):
        self._listeners.setdefault(event, []).append(fn)
        return self

    def emit(self, event, *args):
        for fn in list(self._listeners.get(event, [])):
            fn(*args)

    def push(self, obj):
        with self._lock:
            self._queue.append(obj)
        self.emit("push", obj)

    def pop(self, callback):
        with self._lock:
            obj = self._queue.pop() if self._queue else None
        return callback(obj)

    def get(self, uuid, callback):
        job = None
        with self._lock:
            for j in self._queue:
                if j.get("uuid") == uuid:
                    job = j
                    break
        return callback(job)
