#!/usr/bin/env python
"""Per-pipe instruction histogram and hot-loop listing of one kernel from an ncu report.

    python tools/sass_hist.py REPORT.ncu-rep UNITS [--hot FRACTION] [--sass OUT.sass] [--json OUT.json]

Reads `ncu --page source --csv` (per SASS line: executed warp instructions, stall samples), weighs every
instruction by how often it ran, and reports warp instructions per UNIT (for k_stream: per valid k-mer;
pass the launch's k-mer count) broken down by issue pipe.  --sass writes the hot lines (those executed at
least FRACTION x the most-executed line) as a listing with their counts and stall samples.
Pipe classes follow the Blackwell SASS mnemonics: ALU (logic/shift/add/compare/select/move), FMA-class
integer multiply-add (IMAD*), LSU/MIO (LDS/LDG/STS/ATOM/RED/SHFL/VOTE...), the uniform datapath, control,
conversions.
"""
import argparse
import csv
import json
import re
import subprocess

ALU = {"LOP3", "SHF", "IADD3", "IADD", "ISETP", "SEL", "PRMT", "MOV", "LEA", "IMNMX", "VIMNMX", "SGXT", "BMSK", "POPC", "FLO",
       "BREV", "PLOP3", "ICMP", "IABS", "LOP", "SHL", "SHR", "P2R", "R2P", "CS2R", "VABSDIFF", "FSEL", "FSETP", "FMNMX", "IADD32I",
       "LOP32I", "MOV32I", "VIADD", "VIADDMNMX", "IDP"}
FMA = {"IMAD", "FFMA", "FMUL", "FADD", "HFMA2", "IMUL", "FMA", "IMAD32I"}
FP64 = {"DFMA", "DMUL", "DADD", "DSETP", "DMNMX"}
LSU = {"LDS", "STS", "LDG", "STG", "LD", "ST", "LDL", "STL", "ATOM", "ATOMG", "ATOMS", "RED", "LDSM", "LDC", "LDGSTS", "MEMBAR",
       "ERRBAR", "CCTL", "MATCH", "SHFL", "VOTE", "QSPC", "UBLKCP", "UTMALDG", "SYNCS", "FENCE", "LDGDEPBAR", "DEPBAR", "ELECT",
       "REDUX", "S2R", "LEPC", "NANOSLEEP", "B2R", "R2B"}
CTRL = {"BRA", "BSSY", "BSYNC", "EXIT", "RET", "CALL", "WARPSYNC", "BAR", "BRX", "JMP", "NOP", "YIELD", "BMOV", "BREAK", "BPT", "KILL",
        "ACQBULK", "ENDCOLLECTIVE"}
XU = {"MUFU", "I2F", "F2I", "I2I", "F2F", "I2FP", "F2FP", "FRND", "I2IP"}


def pipe_of(op: str) -> str:
    base = op.split(".")[0]
    if base.startswith("U") and base not in ("UBLKCP", "UTMALDG"):
        return "uniform"
    for name, group in (("alu", ALU), ("fma_imad", FMA), ("fp64", FP64), ("lsu_mio", LSU), ("control", CTRL), ("xu_conv", XU)):
        if base in group:
            return name
    return "other"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("units", type=float, help="units processed by the profiled launch (k-mers, probes, ...)")
    ap.add_argument("--hot", type=float, default=0.2)
    ap.add_argument("--sass")
    ap.add_argument("--json")
    ap.add_argument("--unit-name", default="k-mer")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    kernel = rows[0][1] if rows and rows[0] and rows[0][0] == "Kernel Name" else "?"
    h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[h]
    ci = {n: hdr.index(n) for n in ("Source", "Instructions Executed", "Thread Instructions Executed", "# Samples")}
    insts = []
    for r in rows[h + 1:]:
        if len(r) <= ci["Instructions Executed"]:
            continue
        src = r[ci["Source"]].strip()
        m = re.match(r"(@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", src)
        if not m:
            continue
        insts.append({"addr": r[0], "sass": src, "op": m.group(2), "n": int(r[ci["Instructions Executed"]] or 0),
                      "threads": int(r[ci["Thread Instructions Executed"]] or 0), "samples": int(r[ci["# Samples"]] or 0)})
    total = sum(i["n"] for i in insts)
    # the launch's own counter is the authority for the total (the source page's per-line counts add up to a
    # few % more: it also attributes replayed issues); the per-line counts give the SHARES
    raw_total = None
    try:
        r2 = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
        rr = list(csv.reader(r2.splitlines()))
        raw_total = float(rr[-1][rr[0].index("smsp__inst_executed.sum")])
    except Exception:
        pass
    scale = (raw_total / total) if raw_total else 1.0
    by_pipe, by_op = {}, {}
    for i in insts:
        p = pipe_of(i["op"])
        by_pipe[p] = by_pipe.get(p, 0) + i["n"]
        base = i["op"].split(".")[0]
        by_op[base] = by_op.get(base, 0) + i["n"]
    out = {"kernel": kernel, "report": a.report, "units": a.units, "unit": a.unit_name,
           "warp_instructions": raw_total or total, "warp_instructions_source": "smsp__inst_executed.sum" if raw_total else "source page",
           "source_page_line_total": total,
           "warp_inst_per_unit": (raw_total or total) / a.units,
           "thread_inst_per_unit_x32": 32 * (raw_total or total) / a.units,
           "per_pipe_warp_inst_per_unit": {k: scale * v / a.units for k, v in sorted(by_pipe.items(), key=lambda kv: -kv[1])},
           "per_pipe_share": {k: v / total for k, v in sorted(by_pipe.items(), key=lambda kv: -kv[1])},
           "top_opcodes_warp_inst_per_unit": {k: scale * v / a.units for k, v in sorted(by_op.items(), key=lambda kv: -kv[1])[:24]},
           "stall_samples_total": sum(i["samples"] for i in insts)}
    print(json.dumps(out, indent=1))
    if a.json:
        json.dump(out, open(a.json, "w"), indent=1)
    if a.sass:
        top = max(i["n"] for i in insts)
        with open(a.sass, "w") as fh:
            fh.write("// %s\n// hot lines of %s: executed >= %.0f %% of the most-executed line (%d warp instructions)\n"
                     "// columns: warp instructions executed | stall samples | pipe | SASS\n" % (a.report, kernel, 100 * a.hot, top))
            for i in insts:
                if i["n"] >= a.hot * top:
                    fh.write("%12d %6d  %-9s %s\n" % (i["n"], i["samples"], pipe_of(i["op"]), i["sass"]))


if __name__ == "__main__":
    main()
