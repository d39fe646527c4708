"""Resident sketch-table server: SURVEY.md 8f rank 3 ("persistent GPU table"), built as the form that
measurably wins (profiles/r02_f3_cache_bench.json).

What it removes.  HYMET forks `mash screen` three times per sample
(/root/reference/run_hymet_cami.sh:85-96 -> scripts/mash.sh:14) and every process pays the fixed
cost again: interpreter start, CUDA context creation, `.msh` parse, table build -- about a second at
C2, of which the screen itself is 0.1-0.2 s.  A cache file of the built table does not help (reading
2 GB back costs what rebuilding costs).  Keeping the context AND the tables alive does: this daemon
owns one GPU, holds the tables of the sketch files it has been asked about (LRU), and serves
`mash screen` / the fused three-file stage over a Unix socket.  `bin/mash` attaches to it when
HYMET_SCREEN_SERVER is set (hymet_b200/cli.py) and otherwise runs in-process as before, so the
drop-in contract -- argv in, TSV on stdout, exit code -- is unchanged.

    bin/hymet-screen-server start [--device N] [--socket PATH] [--max-dbs K] [--idle-exit SECONDS]
    bin/hymet-screen-server status | stop
    HYMET_SCREEN_SERVER=1     bin/mash screen ...   # use the daemon if it is there, else in-process
    HYMET_SCREEN_SERVER=auto  bin/mash screen ...   # start it (detached) on first use

Wire format (both directions): frames of  tag (1 byte) | length (4 bytes, little endian) | payload.
Request: one 'J' frame with a JSON object.  Reply: any number of 'E' (stderr text) and 'O' (stdout
bytes) frames, then one 'X' frame with the exit status as ASCII.  The socket is created mode 0600 in
a directory only its owner can enter; there is no other authentication.
"""
from __future__ import annotations

import json
import os
import socket
import struct
import sys
import time
from typing import Dict, List, Optional, Tuple

PROTOCOL = 1


def default_socket_path() -> str:
    if os.environ.get("HYMET_SCREEN_SOCKET"):
        return os.environ["HYMET_SCREEN_SOCKET"]
    base = os.environ.get("XDG_RUNTIME_DIR") or "/tmp"
    return os.path.join(base, "hymet-screen-%d" % os.getuid(), "gpu%s.sock" % os.environ.get("HYMET_SCREEN_DEVICE", "0"))


# ---------------------------------------------------------------- framing (shared with the client in cli.py)
def send_frame(sock: socket.socket, tag: bytes, payload: bytes) -> None:
    sock.sendall(tag + struct.pack("<I", len(payload)) + payload)


def recv_exact(sock: socket.socket, n: int) -> bytes:
    buf = bytearray()
    while len(buf) < n:
        chunk = sock.recv(min(1 << 20, n - len(buf)))
        if not chunk:
            raise ConnectionError("peer closed the connection")
        buf += chunk
    return bytes(buf)


def recv_frame(sock: socket.socket) -> Tuple[bytes, bytes]:
    head = recv_exact(sock, 5)
    (n,) = struct.unpack("<I", head[1:])
    return head[:1], recv_exact(sock, n) if n else b""


class _SockWriter:
    """File-like that turns writes into frames (stdout of a served command)."""

    def __init__(self, sock, tag):
        self.sock, self.tag, self.buf = sock, tag, []
        self.n = 0

    def write(self, s):
        b = s.encode("utf-8", "surrogateescape") if isinstance(s, str) else bytes(s)
        self.buf.append(b)
        self.n += len(b)
        if self.n >= (1 << 16):
            self.flush()
        return len(s)

    def flush(self):
        if self.buf:
            send_frame(self.sock, self.tag, b"".join(self.buf))
            self.buf, self.n = [], 0


# ---------------------------------------------------------------- client side
def request(sock_path: str, req: dict, stdout=None, stderr=None, timeout: Optional[float] = None) -> int:
    """Send one request, relay the reply's frames to stdout/stderr, return the exit status.
    Raises ConnectionError/OSError when no server answers (the caller then works in-process)."""
    stdout = stdout or sys.stdout
    stderr = stderr or sys.stderr
    s = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    s.settimeout(timeout)
    try:
        s.connect(sock_path)
        send_frame(s, b"J", json.dumps(req).encode())
        out = getattr(stdout, "buffer", None)
        while True:
            tag, payload = recv_frame(s)
            if tag == b"O":
                if out is not None:
                    out.write(payload)
                else:
                    stdout.write(payload.decode("utf-8", "surrogateescape"))
            elif tag == b"E":
                stderr.write(payload.decode("utf-8", "replace"))
            elif tag == b"X":
                if out is not None:
                    out.flush()
                stdout.flush()
                return int(payload.decode() or "1")
            else:
                raise ConnectionError("unexpected frame %r" % tag)
    finally:
        s.close()


def spawn_detached(sock_path: str, device: int) -> None:
    """Start the daemon in its own session; returns when the socket answers or after ~60 s."""
    import subprocess
    exe = [sys.executable, "-m", "hymet_b200.server", "serve", "--socket", sock_path, "--device", str(device)]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PYTHONPATH=root + os.pathsep + os.environ.get("PYTHONPATH", ""))
    os.makedirs(os.path.dirname(sock_path), mode=0o700, exist_ok=True)
    log = open(sock_path + ".log", "ab")
    subprocess.Popen(exe, stdin=subprocess.DEVNULL, stdout=log, stderr=log, start_new_session=True, env=env, cwd="/")
    t_end = time.time() + 60
    while time.time() < t_end:
        try:
            if request(sock_path, {"op": "ping"}, stdout=_Null(), stderr=_Null(), timeout=2.0) == 0:
                return
        except OSError:
            time.sleep(0.05)
    raise ConnectionError("the screen server did not come up (see %s.log)" % sock_path)


class _Null:
    def write(self, s):
        return len(s)

    def flush(self):
        pass


# ---------------------------------------------------------------- server side
class Server:
    def __init__(self, sock_path: str, device: int, max_dbs: int = 4, idle_exit: float = 0.0):
        self.sock_path, self.device, self.max_dbs, self.idle_exit = sock_path, device, max_dbs, idle_exit
        self.dbs: Dict[tuple, dict] = {}        # key -> {"db": LiteDb, "scr": LiteScreen | None, "used": t, "paths": [...]}
        self.n_served = 0
        self.t_start = time.time()
        from . import _lite as hs
        self.hs = hs
        hs._abi.init(device)                      # CUDA context: paid once, here

    # -- tables ------------------------------------------------------------
    @staticmethod
    def _key(paths: List[str]) -> tuple:
        k = []
        for p in paths:
            try:
                st = os.stat(p)
                k.append((os.path.realpath(p), st.st_mtime_ns, st.st_size))
            except OSError:
                k.append((os.path.realpath(p), -1, -1))     # the library reports the missing file; nothing is cached
        return tuple(k)

    def table(self, paths: List[str], tolerate: bool = False):
        """The resident table over `paths` (one file, or several for the fused stage), building it on
        first use.  A changed file (mtime / size) is a different key: it is rebuilt, the old one ages out."""
        hs = self.hs
        present = [p for p in paths if os.path.exists(p)] if tolerate else paths
        key = (self._key(present), tuple(paths), tolerate)
        ent = self.dbs.get(key)
        if ent is None:
            if len(paths) == 1 and not tolerate:
                db = hs.LiteDb(paths[0], self.device)
            else:
                db = hs.LiteDb(paths, self.device, tolerate=tolerate)
            ent = {"db": db, "scr": None, "used": 0.0}
            self.dbs[key] = ent
            while len(self.dbs) > self.max_dbs:
                old = min((k for k in self.dbs if k != key), key=lambda k: self.dbs[k]["used"])
                self._drop(old)
        ent["used"] = time.time()
        return ent

    def _drop(self, key):
        ent = self.dbs.pop(key)
        if ent["scr"] is not None:
            ent["scr"].close()
        ent["db"].close()

    def screen_for(self, ent):
        """One long-lived screen per table: its arena, pinned ring and sparse buffers stay allocated."""
        hs = self.hs
        if ent["scr"] is None:
            ent["scr"] = hs.LiteScreen(ent["db"])
            ent["scr"].set_option("file_readers", 8)            # a long-lived handle amortises pinning a bigger ring
            ent["scr"].set_option("file_block_bytes", 16 << 20)
        else:
            ent["scr"].reset()
        return ent["scr"]

    # -- requests ----------------------------------------------------------
    def handle(self, conn: socket.socket) -> bool:
        """Serve one connection.  Returns False when asked to shut down."""
        tag, payload = recv_frame(conn)
        if tag != b"J":
            raise ConnectionError("expected a request frame")
        req = json.loads(payload.decode())
        out, err = _SockWriter(conn, b"O"), _SockWriter(conn, b"E")
        op = req.get("op")
        status, keep = 1, True
        try:
            if op == "ping":
                status = 0
            elif op == "shutdown":
                status, keep = 0, False
            elif op == "status":
                info = {"protocol": PROTOCOL, "pid": os.getpid(), "device": self.device, "uptime_s": time.time() - self.t_start,
                        "served": self.n_served,
                        "tables": [{"files": list(k[1]), "refs": e["db"].n_refs, "hbm_mb": e["db"].info.device_bytes / 1e6}
                                   for k, e in self.dbs.items()]}
                out.write(json.dumps(info) + "\n")
                status = 0
            elif op == "screen":
                from . import cli
                cwd = req.get("cwd") or "/"
                argv = req["argv"]
                status = cli.screen_main(argv, stdout=out, stderr=err, cwd=cwd, server=self)
                self.n_served += 1
            elif op == "stage":
                from . import stage
                status = stage.main(req["argv"], stdout=out, stderr=err, cwd=req.get("cwd") or "/", server=self)
                self.n_served += 1
            else:
                err.write("ERROR: unknown request %r\n" % op)
        except Exception as e:                                    # noqa: BLE001  (a request must never take the daemon down)
            err.write("ERROR: %s: %s\n" % (type(e).__name__, e))
            status = 1
        out.flush()
        err.flush()
        send_frame(conn, b"X", str(status).encode())
        return keep

    def serve_forever(self) -> None:
        d = os.path.dirname(self.sock_path)
        os.makedirs(d, mode=0o700, exist_ok=True)
        if os.path.exists(self.sock_path):
            try:            # a live server already owns the socket: leave it alone
                if request(self.sock_path, {"op": "ping"}, stdout=_Null(), stderr=_Null(), timeout=2.0) == 0:
                    sys.stderr.write("hymet-screen-server: already running on %s\n" % self.sock_path)
                    return
            except OSError:
                pass
            os.unlink(self.sock_path)
        srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
        old = os.umask(0o177)
        try:
            srv.bind(self.sock_path)
        finally:
            os.umask(old)
        srv.listen(16)
        srv.settimeout(self.idle_exit if self.idle_exit > 0 else None)
        sys.stderr.write("hymet-screen-server: device %d, socket %s, pid %d\n" % (self.device, self.sock_path, os.getpid()))
        sys.stderr.flush()
        try:
            while True:
                try:
                    conn, _ = srv.accept()
                except socket.timeout:
                    sys.stderr.write("hymet-screen-server: idle for %.0f s, exiting\n" % self.idle_exit)
                    break
                try:
                    conn.settimeout(None)
                    if not self.handle(conn):
                        break
                except (ConnectionError, OSError, ValueError) as e:
                    sys.stderr.write("hymet-screen-server: dropped a connection: %s\n" % e)
                finally:
                    conn.close()
        finally:
            srv.close()
            try:
                os.unlink(self.sock_path)
            except OSError:
                pass
            for k in list(self.dbs):
                self._drop(k)


def main(argv: Optional[List[str]] = None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        sys.stdout.write(__doc__)
        return 0
    cmd, rest = argv[0], argv[1:]
    opts = {"--socket": None, "--device": os.environ.get("HYMET_SCREEN_DEVICE", "0"), "--max-dbs": "4", "--idle-exit": "0"}
    i = 0
    while i < len(rest):
        if rest[i] in opts and i + 1 < len(rest):
            opts[rest[i]] = rest[i + 1]
            i += 2
        else:
            sys.stderr.write("hymet-screen-server: unknown argument %s\n" % rest[i])
            return 2
    device = int(opts["--device"])
    os.environ["HYMET_SCREEN_DEVICE"] = str(device)
    sock_path = opts["--socket"] or default_socket_path()
    if cmd == "serve":                      # foreground
        Server(sock_path, device, int(opts["--max-dbs"]), float(opts["--idle-exit"])).serve_forever()
        return 0
    if cmd == "start":                      # detached
        try:
            spawn_detached(sock_path, device)
        except ConnectionError as e:
            sys.stderr.write("ERROR: %s\n" % e)
            return 1
        sys.stdout.write("%s\n" % sock_path)
        return 0
    if cmd in ("status", "stop"):
        try:
            return request(sock_path, {"op": "status" if cmd == "status" else "shutdown"}, timeout=10.0)
        except OSError:
            sys.stderr.write("hymet-screen-server: not running (%s)\n" % sock_path)
            return 1
    sys.stderr.write("hymet-screen-server: unknown command %s\n" % cmd)
    return 2


if __name__ == "__main__":
    sys.exit(main())
