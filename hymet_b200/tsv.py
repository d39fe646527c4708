"""`mash screen` wire format (S15/S16): the TSV that scripts/mash.sh:15-55,
scripts/limit_candidates.py:97-122 and scripts/downloadDB.py:106-111 consume.

identity \\t shared/size \\t median-multiplicity \\t p-value \\t query-ID \\t query-comment
Numbers are C++ ``ostream << double`` = C ``%g`` with precision 6; Python's ``%g`` is the
same correctly-rounded conversion.
"""
from __future__ import annotations

from typing import Iterable, List, Sequence

import numpy as np


def fmt_g(x: float) -> str:
    return "%g" % x


def screen_lines(shared: Sequence[int], sizes: Sequence[int], median: Sequence[int], identity: Sequence[float],
                 pvalue: Sequence[float], names: Sequence[str], comments: Sequence[str],
                 min_identity: float = 0.0, max_pvalue: float = 1.0) -> List[str]:
    """S15: sketch order; keep iff (shared>0 or -i<0) and identity >= -i and p <= -v."""
    out = []
    shared = np.asarray(shared); identity = np.asarray(identity); pvalue = np.asarray(pvalue)
    keep = ((shared != 0) | (min_identity < 0.0)) & ~(identity < min_identity) & ~(pvalue > max_pvalue)
    for i in np.nonzero(keep)[0]:
        out.append("%s\t%d/%d\t%d\t%s\t%s\t%s\n" % (fmt_g(float(identity[i])), int(shared[i]), int(sizes[i]),
                                                  int(median[i]), fmt_g(float(pvalue[i])), names[i], comments[i]))
    return out


def screen_lines_db(res, db, min_identity: float = 0.0, max_pvalue: float = 1.0, begin: int = 0, end: int = None) -> List[str]:
    """Same lines for references [begin, end) of a hymet_b200.screen.Database, looking up names and
    sketch sizes only for the references that are reported (S15 keeps few of 50 000+)."""
    end = db.n_refs if end is None else end
    shared = np.asarray(res.shared[begin:end]); identity = np.asarray(res.identity[begin:end])
    pvalue = np.asarray(res.pvalue[begin:end])
    keep = ((shared != 0) | (min_identity < 0.0)) & ~(identity < min_identity) & ~(pvalue > max_pvalue)
    out = []
    for i in np.nonzero(keep)[0]:
        name, comment, _, size = db.ref(begin + int(i))
        out.append("%s\t%d/%d\t%d\t%s\t%s\t%s\n" % (fmt_g(float(identity[i])), int(shared[i]), size,
                                                  int(res.median[begin + i]), fmt_g(float(pvalue[i])), name, comment))
    return out


def write_screen(fh, lines: Iterable[str]) -> None:
    for ln in lines:
        fh.write(ln)
