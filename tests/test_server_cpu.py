"""Resident table server: protocol, client fallback and argument handling on a box without a GPU.
(The served screen itself is a GPU test: tests/test_gpu_server.py.)"""
import json
import os
import socket
import subprocess
import sys
import threading

import pytest

from hymet_b200 import server

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_framing_round_trip():
    a, b = socket.socketpair()
    try:
        server.send_frame(a, b"J", json.dumps({"op": "ping"}).encode())
        server.send_frame(a, b"O", b"")
        server.send_frame(a, b"E", b"x" * 200_000)
        assert server.recv_frame(b) == (b"J", b'{"op": "ping"}')
        assert server.recv_frame(b) == (b"O", b"")
        tag, payload = server.recv_frame(b)
        assert tag == b"E" and payload == b"x" * 200_000
        a.close()
        with pytest.raises(ConnectionError):
            server.recv_frame(b)
    finally:
        b.close()


def test_client_relays_frames_and_exit_status(tmp_path):
    """A stand-in server (no GPU) answers a request: stdout bytes, stderr text and the status reach the client."""
    path = str(tmp_path / "s.sock")
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    srv.bind(path)
    srv.listen(1)
    seen = {}

    def serve():
        conn, _ = srv.accept()
        tag, payload = server.recv_frame(conn)
        seen["req"] = json.loads(payload.decode())
        w = server._SockWriter(conn, b"O")
        w.write("0.99957\t991/1000\t3\t0\tGCF_x\tc\n")
        w.flush()
        server.send_frame(conn, b"E", b"Loading db.msh...\n")
        server.send_frame(conn, b"X", b"7")
        conn.close()

    th = threading.Thread(target=serve)
    th.start()

    class Buf:
        def __init__(self): self.s = ""
        def write(self, x): self.s += x; return len(x)
        def flush(self): pass
    out, err = Buf(), Buf()
    rc = server.request(path, {"op": "screen", "argv": ["db.msh", "q.fna"], "cwd": "/w"}, stdout=out, stderr=err, timeout=10)
    th.join()
    srv.close()
    assert rc == 7 and out.s.startswith("0.99957\t991/1000") and err.s == "Loading db.msh...\n"
    assert seen["req"] == {"op": "screen", "argv": ["db.msh", "q.fna"], "cwd": "/w"}


def test_mash_falls_back_in_process_when_no_server_answers(tmp_path):
    """HYMET_SCREEN_SERVER=1 with nothing listening: the drop-in behaves exactly as without the variable
    (here, without a GPU: the in-process path fails loudly with exit 1 -- never a silent CPU result)."""
    env = dict(os.environ, HYMET_SCREEN_SERVER="1", HYMET_SCREEN_SOCKET=str(tmp_path / "none.sock"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "mash"), "screen", "x.txt", "q.fna"], capture_output=True, text=True, env=env)
    assert r.returncode == 1 and "does not look like a sketch" in r.stderr and r.stdout == ""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "mash"), "screen", "-h"], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "mash screen" in r.stdout


def test_server_cli_status_without_daemon(tmp_path):
    env = dict(os.environ, HYMET_SCREEN_SOCKET=str(tmp_path / "none.sock"))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "hymet-screen-server"), "status"], capture_output=True, text=True, env=env)
    assert r.returncode == 1 and "not running" in r.stderr
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bin", "hymet-screen-server"), "--help"], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "HYMET_SCREEN_SERVER" in r.stdout
