"""Host-side mirror of the reference's peer-bootstrap pipeline surface.

Node.js is not installed in this image (`node --version`: not found), so the host
side above the C ABI is mirrored in Python with the same names, argument meaning,
job-object fields and error behaviour as the reference modules:

    backup_queue.BackupQueue     lib/backupQueue.js
    backup_server.BackupServer   lib/backupServer.js   (POST /backup/, GET /backup/:uuid)
    backup_sender.BackupSender   lib/backupSender.js   (_send, _getLatestSnapshot)
    zfs_client.ZfsClient         lib/zfsClient.js      (_receive, _postRestoreRequest,
                                                        _pollRestoreCompletion, restore,
                                                        isolateDataset, snapshotDataset)
    backupserver (module main)   backupserver.js       (-f <config>, same config file shape)
    snapshotter  (module main)   snapshotter.js        (-f <config>)
    zfs_cmd                      lib/common.js         (zfsSet/Get/Inherit/Rename/Mount/...)
    snap_shotter.SnapShotter     lib/snapShotter.js    (8f f3: snapshot cadence + GC)
    status_server.StatusServer   lib/statusServer.js   (8f f4: GET /restore, /ping)
    status_server.RestoreWatcher lib/adm.js:1550-1678  (8f f4: rebuild progress consumer)

``js/`` holds the Node sources a maintainer ships; both splice the same
``GpuSnapshotStage`` into the two ``.pipe()`` calls.
"""
from .backup_queue import BackupQueue  # noqa: F401
from .backup_server import BackupServer  # noqa: F401
from .backup_sender import BackupSender  # noqa: F401
from .zfs_client import ZfsClient  # noqa: F401
from .snap_shotter import SnapShotter  # noqa: F401
from .status_server import StatusServer, RestoreWatcher  # noqa: F401
