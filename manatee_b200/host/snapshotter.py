"""snapshotter -- mirror of the reference's snapshotter.js entry point.

    python -m manatee_b200.host.snapshotter -f etc/snapshotter.json [-v]

Same command line and configuration file as the reference daemon (snapshotter.js,
test/etc/snapshotter.json): { "dataset": ..., "pollInterval": 2000, "snapshotNumber": 5,
"healthUrl": ... }.  Errors from the SnapShotter are logged, never fatal
(snapshotter.js:113-115).
"""
import json
import logging
import sys
import threading

from .backupserver import parseOptions, readConfig
from .snap_shotter import SnapShotter

NAME = "manatee-snapshotter"


def main(argv=None):
    options = parseOptions(sys.argv[1:] if argv is None else argv)
    logging.basicConfig(level=logging.DEBUG if options.get("verbose") else logging.INFO,
                        format="%(asctime)s %(name)s %(levelname)s %(message)s")
    log = logging.getLogger(NAME)
    config = readConfig(options)
    config["log"] = log
    snapShotter = SnapShotter(config)
    snapShotter.on("error", lambda err: log.error("got error from snapshotter: %s", err))
    snapShotter.start(lambda: log.info("snapshotter started"))
    print(json.dumps({"name": NAME, "dataset": config["dataset"]}), flush=True)
    threading.Event().wait()


if __name__ == "__main__":
    main()
