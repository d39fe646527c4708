"""CPU restatement of the classifier's weighted-LCA vote -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/scripts/classification_cami.py:251-308 (`_weighted_lca`, `_process_one`) on the
integer-coded form the CUDA kernel takes (hymet_b200/csrc/lca_kernels.cu): plain Python floats are IEEE
doubles, dicts keep insertion order, `max` returns the first maximal item -- the three properties the
kernel has to reproduce bit for bit.  Pinned against the reference itself: tests/golden/make_lca_golden.py
imports the reference module in the build container and commits its inputs and outputs.
Nothing under hymet_b200/ imports this module.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple


def weighted_lca(q_off: Sequence[int], tax: Sequence[int], w: Sequence[float], names: Sequence[Sequence[int]],
                 n_ranks: int = 8) -> List[Tuple[List[int], float, bool]]:
    """Per query: (chosen name id per resolved rank, confidence, any alignment had a taxid)."""
    out = []
    for q in range(len(q_off) - 1):
        tw = {}
        any_hit = False
        for j in range(int(q_off[q]), int(q_off[q + 1])):          # classification_cami.py:291-299
            t = int(tax[j])
            if t < 0:
                continue
            any_hit = True
            tw[t] = tw.get(t, 0.0) + float(w[j])
        chosen: List[int] = []
        conf = 1.0
        if any_hit and sum(tw.values()) > 0:                        # :259-261
            for r in range(n_ranks):                                # :266-284
                name_w = {}
                denom = 0.0
                for t, wt in tw.items():
                    nm = int(names[t][r])
                    if nm:
                        name_w[nm] = name_w.get(nm, 0.0) + wt
                        denom += wt
                if denom <= 0 or not name_w:
                    break
                best, best_w = max(name_w.items(), key=lambda kv: kv[1])
                chosen.append(best)
                conf *= best_w / denom
        out.append((chosen, min(conf, 1.0) if chosen else 0.0, any_hit))
    return out
