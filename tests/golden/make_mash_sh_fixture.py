#!/usr/bin/env python
"""Test fixture: the verbatim bytes of the reference's scripts/mash.sh, gzip + base64 inside
mash_sh_fixture.json (run in the build container: needs /root/reference).

Why: the drop-in claim is "the UNMODIFIED script works with our `mash` first on PATH".  The GPU box has
no /root/reference, so the script travels as test DATA -- never imported, never on a product path --
and tests/test_gpu_parity.py::test_full_reference_mash_sh_with_gpu_mash unpacks it into a temporary
directory, together with the `bc` stand-in of tests/test_stage_cpu.py (this image has no bc)."""
import base64
import gzip
import hashlib
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/scripts/mash.sh"
raw = open(SRC, "rb").read()
json.dump({"provenance": SRC + " (jorgeMFS/HYMET), verbatim; test fixture only",
           "sha256": hashlib.sha256(raw).hexdigest(), "n_bytes": len(raw), "n_lines": raw.count(b"\n"),
           "gzip_base64": base64.b64encode(gzip.compress(raw, 9, mtime=0)).decode()},
          open(os.path.join(HERE, "mash_sh_fixture.json"), "w"), indent=1)
print("wrote mash_sh_fixture.json:", len(raw), "bytes")
