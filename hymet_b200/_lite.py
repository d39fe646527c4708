"""numpy-free binding for the one-shot `mash screen` process (hymet_b200/cli.py).

A `mash screen` invocation on C2 is about one second of wall clock, 0.1-0.4 s of which was
`import numpy`.  The CLI needs none of it: the four result columns are plain C arrays and the
TSV formatter walks them once.  Same C ABI, same reporting rules (S15/S16) as
hymet_b200/screen.py + hymet_b200/tsv.py; tests/test_gpu_parity.py compares the CLI's bytes with
the oracle CLI's.
"""
from __future__ import annotations

import ctypes as C
import threading
from typing import Iterator, List

from . import _abi
from ._abi import HsError, check  # noqa: F401  (re-exported)


class LiteDb:
    def __init__(self, path: str, device: int):
        """Parse the .msh on a helper thread (host only) while this thread creates the CUDA context."""
        L = _abi.load()
        m, box = C.c_void_p(), {}

        def parse():
            try:
                check(L.hs_msh_open(path.encode(), C.byref(m)))   # hs_last_error() is thread local: check here
            except Exception as e:                                 # noqa: BLE001
                box["err"] = e

        th = threading.Thread(target=parse)
        th.start()
        try:
            _abi.init(device)
        finally:
            th.join()
        if "err" in box:
            raise box["err"]
        self._h = C.c_void_p()
        try:
            check(L.hs_db_from_msh(m, C.byref(self._h)))
        finally:
            L.hs_msh_free(m)
        self.info = _abi.DbInfo()
        check(L.hs_db_info(self._h, C.byref(self.info)))
        self.n_refs = int(self.info.n_refs)
        self.n_distinct = int(self.info.n_distinct)

    def ref(self, i: int):
        nm, cm = C.c_char_p(), C.c_char_p()
        ln, nh = C.c_uint64(), C.c_uint64()
        check(_abi.load().hs_db_ref(self._h, i, C.byref(nm), C.byref(cm), C.byref(ln), C.byref(nh)))
        return ((nm.value or b"").decode("utf-8", "replace"), (cm.value or b"").decode("utf-8", "replace"), ln.value, nh.value)


class LiteScreen:
    def __init__(self, db: LiteDb, probe_filter: bool = True):
        self.db = db
        self._h = C.c_void_p()
        check(_abi.load().hs_screen_new(db._h, C.byref(self._h)))
        if not probe_filter:
            check(_abi.load().hs_screen_set_option(self._h, b"filter", 0))

    def feed_fasta(self, path: str, threads: int):
        check(_abi.load().hs_screen_feed_fasta(self._h, path.encode(), threads))

    def flush(self):
        check(_abi.load().hs_screen_flush(self._h))

    def stats(self) -> dict:
        st = _abi.Stats()
        check(_abi.load().hs_screen_stats(self._h, C.byref(st)))
        return st.asdict()

    def finish_lines(self, wta: bool, min_identity: float, max_pvalue: float) -> Iterator[str]:
        """Rows a11-a16: reduce, then the TSV lines in sketch order (S15: keep iff (shared > 0 or -i < 0)
        and identity >= -i and p <= -v; numbers as C `%g`)."""
        n = max(self.db.n_refs, 1)
        shared = (C.c_uint64 * n)()
        median = (C.c_uint32 * n)()
        identity = (C.c_double * n)()
        pvalue = (C.c_double * n)()
        check(_abi.load().hs_screen_finish(self._h, int(wta), shared, median, identity, pvalue, None))
        all_rows = min_identity < 0.0
        for i in range(self.db.n_refs):
            sh = shared[i]
            if not sh and not all_rows:
                continue
            ident, p = identity[i], pvalue[i]
            if ident < min_identity or p > max_pvalue:
                continue
            name, comment, _, size = self.db.ref(i)
            yield "%s\t%d/%d\t%d\t%s\t%s\t%s\n" % ("%g" % ident, sh, size, median[i], "%g" % p, name, comment)


def screen_lines(db_path: str, inputs: List[str], threads: int, wta: bool, min_identity: float, max_pvalue: float,
                 device: int, probe_filter: bool = True):
    """Convenience used by tests: the whole screen as a list of TSV lines."""
    db = LiteDb(db_path, device)
    scr = LiteScreen(db, probe_filter)
    for p in inputs:
        scr.feed_fasta(p, threads)
    return list(scr.finish_lines(wta, min_identity, max_pvalue))
