"""zfs(1M) metadata helpers -- mirror of lib/common.js:148-460.

The reference forks `/sbin/zfs` with an EMPTY environment (locale-independent
output, lib/common.js:153-163) and wraps every failure in a VError whose message
names the operation ("set property ... on dataset ...: <cause>").  Same here: each
helper raises ZfsError("<operation>: <cause>"), the cause carrying the exit status
and stderr.  `opts['zfs']` overrides the binary (the tests point it at
tools/fake_zfs.py; the reference hard-codes the path) and `opts['env']` the
environment handed to it (default: empty, like the reference).

None of this touches stream bytes; it is the dataset lifecycle either side of the
bulk-data path.
"""
import subprocess

ZFS_BIN = "/sbin/zfs"
MAX_BUFFER = 2 * 1024 * 1024            # lib/common.js:160


class ZfsError(RuntimeError):
    def __init__(self, msg, info=None, cause=None):
        RuntimeError.__init__(self, msg)
        self.info = info or {}
        self.cause = cause


def _wrap(cause, fmt, *args):
    return ZfsError("%s: %s" % (fmt % args, cause), info=getattr(cause, "info", None), cause=cause)


def zfsExecCommon(opts, argv):
    """lib/common.js:148-171 -> {'stdout','stderr','status'}; raises ZfsError on exit != 0"""
    assert all(isinstance(a, str) for a in argv), "args (arrayOfString) is required"
    exe = opts.get("zfs") or ZFS_BIN
    try:
        p = subprocess.run([exe] + list(argv), env=opts.get("env") or {}, stdout=subprocess.PIPE,
                           stderr=subprocess.PIPE, timeout=opts.get("timeout", 300))
    except OSError as e:
        raise ZfsError('exec "%s %s": %s' % (exe, " ".join(argv), e))
    info = {"stdout": p.stdout[:MAX_BUFFER].decode("utf-8", "replace"),
            "stderr": p.stderr[:MAX_BUFFER].decode("utf-8", "replace"), "status": p.returncode}
    if p.returncode != 0:
        raise ZfsError('exec "%s %s": exited with status %d: %s' %
                       (exe, " ".join(argv), p.returncode, info["stderr"].strip()), info=info)
    return info


def _req(opts, *names):
    assert isinstance(opts, dict), "opts (object) is required"
    for n in names:
        assert isinstance(opts.get(n), str), "opts.%s (string) is required" % n


def zfsSet(opts):                                             # lib/common.js:177-199
    _req(opts, "dataset", "property", "value")
    try:
        zfsExecCommon(opts, ["set", opts["property"] + "=" + opts["value"], opts["dataset"]])
    except ZfsError as e:
        raise _wrap(e, 'set property "%s" to "%s" on dataset "%s"', opts["property"], opts["value"],
                    opts["dataset"])


def zfsInherit(opts):                                         # lib/common.js:204-224
    _req(opts, "dataset", "property")
    try:
        zfsExecCommon(opts, ["inherit", opts["property"], opts["dataset"]])
    except ZfsError as e:
        raise _wrap(e, 'clear property "%s" on dataset "%s"', opts["property"], opts["dataset"])


def zfsGet(opts):                                             # lib/common.js:229-259
    _req(opts, "dataset", "property")
    try:
        info = zfsExecCommon(opts, ["get", "-Hp", opts["property"], opts["dataset"]])
    except ZfsError as e:
        raise _wrap(e, 'get property "%s" from dataset "%s"', opts["property"], opts["dataset"])
    t = info["stdout"].split("\t")
    if len(t) != 4 or t[0] != opts["dataset"] or t[1] != opts["property"]:
        raise ZfsError('zfs get "%s" "%s": invalid line: %s' %
                       (opts["property"], opts["dataset"], info["stdout"].strip()))
    return t[2]


def zfsSnapshot(opts):                                        # lib/common.js:264-284
    _req(opts, "dataset", "snapshot")
    try:
        zfsExecCommon(opts, ["snapshot", opts["dataset"] + "@" + opts["snapshot"]])
    except ZfsError as e:
        raise _wrap(e, 'snapshot dataset "%s" as "%s"', opts["dataset"], opts["snapshot"])


def zfsCreate(opts):                                          # lib/common.js:289-314
    _req(opts, "dataset")
    args = ["create"]
    for k, v in sorted((opts.get("props") or {}).items()):
        args += ["-o", "%s=%s" % (k, v)]
    args.append(opts["dataset"])
    try:
        zfsExecCommon(opts, args)
    except ZfsError as e:
        raise _wrap(e, 'create dataset "%s"', opts["dataset"])


def zfsRename(opts):                                          # lib/common.js:319-345
    _req(opts, "dataset", "target")
    assert isinstance(opts.get("parents"), bool), "opts.parents (bool) is required"
    args = ["rename"] + (["-p"] if opts["parents"] else []) + [opts["dataset"], opts["target"]]
    try:
        zfsExecCommon(opts, args)
    except ZfsError as e:
        raise _wrap(e, 'rename dataset "%s" to "%s"', opts["dataset"], opts["target"])


def zfsMount(opts):                                           # lib/common.js:350-366
    _req(opts, "dataset")
    try:
        zfsExecCommon(opts, ["mount", opts["dataset"]])
    except ZfsError as e:
        raise _wrap(e, 'mount dataset "%s"', opts["dataset"])


def zfsUnmount(opts):                                         # lib/common.js:401-426
    _req(opts, "dataset")
    args = ["unmount"] + (["-f"] if opts.get("force") else []) + [opts["dataset"]]
    try:
        zfsExecCommon(opts, args)
    except ZfsError as e:
        raise _wrap(e, 'unmount dataset "%s"', opts["dataset"])


def zfsDestroy(opts):                                         # lib/common.js:371-396
    _req(opts, "dataset")
    args = ["destroy"] + (["-r"] if opts.get("recursive") else []) + [opts["dataset"]]
    try:
        zfsExecCommon(opts, args)
    except ZfsError as e:
        raise _wrap(e, '%sdestroy dataset "%s"', "recursively " if opts.get("recursive") else "",
                    opts["dataset"])


def zfsExists(opts):                                          # lib/common.js:431-450
    _req(opts, "dataset")
    try:
        info = zfsExecCommon(opts, ["list", "-Hp", "-o", "name"])
    except ZfsError as e:
        raise _wrap(e, 'check for dataset "%s"', opts["dataset"])
    return opts["dataset"] in info["stdout"].split("\n")
