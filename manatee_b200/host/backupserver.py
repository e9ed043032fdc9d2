"""backupserver -- mirror of the reference's backupserver.js entry point.

    python -m manatee_b200.host.backupserver -f etc/backupserver.json [-v]

Same command line (`-f <file>`, `-v`) and the same configuration file as the reference
daemon (backupserver.js:46-125, test/etc/backupserver.json):

    { "backupServerCfg": { "port": 12345 },
      "backupSenderCfg": { "zfsPath": "/usr/sbin/zfs", "dataset": "zones/<uuid>/data/manatee",
                           "gpu": { "mode": "verify" }, "coalesceMs": 0 } }

`gpu` and `coalesceMs` are the additive keys; without them the process
behaves like the reference: REST server + sender sharing one queue, identity pipe.
An unreadable / unparsable configuration file is fatal (the reference aborts).
"""
import getopt
import json
import logging
import sys
import threading

from .backup_sender import BackupSender
from .backup_server import BackupServer

NAME = "manatee-backupserver"


def parseOptions(argv):
    opts = {}
    try:
        got, _ = getopt.getopt(argv, "vf:", ["file="])
    except getopt.GetoptError as e:
        logging.getLogger(NAME).critical("Unsupported option: %s", e)
        raise SystemExit(2)
    for o, a in got:
        if o in ("-f", "--file"):
            opts["file"] = a
        elif o == "-v":
            opts["verbose"] = opts.get("verbose", 0) + 1
    return opts


def readConfig(options):
    try:
        with open(options["file"], "r") as f:
            cfg = json.load(f)
    except Exception as e:                                   # noqa: BLE001
        logging.getLogger(NAME).critical("Unable to read/parse configuration file %r: %s",
                                         options.get("file"), e)
        raise SystemExit(134)                                # process.abort()
    cfg.update(options)
    return cfg


def start(config):
    """-> (BackupServer, BackupSender): server and sender share the same queue"""
    log = logging.getLogger(NAME)
    config["backupServerCfg"]["log"] = log
    config["backupSenderCfg"]["log"] = log
    backupServer = BackupServer.start(config["backupServerCfg"])
    config["backupSenderCfg"]["queue"] = backupServer.getQueue()
    sender = BackupSender.start(config["backupSenderCfg"])
    sender.on("err", lambda e: log.error("unable to send backup: %s", e))
    sender.on("done", lambda j: log.info("successfully sent backup %s", j.get("uuid")))
    log.info("backupserver started")
    return backupServer, sender


def main(argv=None):
    options = parseOptions(sys.argv[1:] if argv is None else argv)
    logging.basicConfig(level=logging.DEBUG if options.get("verbose") else logging.INFO,
                        format="%(asctime)s %(name)s %(levelname)s %(message)s")
    config = readConfig(options)
    server, _ = start(config)
    print(json.dumps({"name": NAME, "port": server.port}), flush=True)     # for supervisors/tests
    threading.Event().wait()                                 # the daemon never exits on its own


if __name__ == "__main__":
    main()
