"""Mash sketch files (``.msh``): Cap'n Proto reader and writer in NumPy.

HYMET reads three such files, ``data/sketch1.msh`` .. ``sketch3.msh``
(/root/reference/run_hymet_cami.sh:52,85-96, main.pl:44-46), through the
external ``mash`` binary (/root/reference/scripts/mash.sh:14).  There is no
libcapnp / pycapnp in this image, so the format (standard
unpacked Cap'n Proto stream framing + Mash's ``MinHash`` schema) is decoded by
hand.  The product's loader is the C++ one in ``csrc/msh_capnp.cpp`` (behind
``hs_db_load_msh``); this module is

* the **writer** (``mash sketch`` is never run by HYMET and is absent here, so
  test/bench databases are fabricated with it), able to emit multi-segment
  files with single- and double-far pointers like large real sketch files, and
* an independent **reader** used by tests and by ``python -m hymet_b200.msh``
  (a verifier that dumps the header of any user-supplied ``.msh``).

Schema slots (Cap'n Proto layout rules applied to Mash's MinHash.capnp)::

    MinHash        data 3 words / 4 pointers
      kmerSize u32 @byte0   windowSize u32 @4   minHashesPerWindow u32 @8
      concatenated bit96    noncanonical bit97  preserveCase bit98
      error f32 @16         hashSeed u32 @20 (stored XOR 42)
      ptr0 referenceListOld ptr1 locusList  ptr2 alphabet  ptr3 referenceList
    ReferenceList  data 0 / 1 pointer: references List(Reference) (composite)
    Reference      data 3 words / 7 pointers
      length u32 @0  length64 u64 @8  numValidKmers u64 @16
      ptr0 sequence ptr1 quality ptr2 name ptr3 comment
      ptr4 hashes32 ptr5 hashes64 ptr6 counts32
"""
from __future__ import annotations

import struct
import sys
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

U64 = np.uint64
_MH_D, _MH_P = 3, 4
_REF_D, _REF_P = 3, 7


def use64(k: int, alphabet_size: int = 4) -> bool:
    """S1: hashes are 64-bit iff alphabet^k > 2^32 (nucleotides: k >= 17)."""
    return float(alphabet_size) ** k > 2.0 ** 32


@dataclass
class SketchDB:
    """Host-side view of a sketch database (what ``mash info`` would show)."""

    k: int
    s: int
    seed: int = 42
    names: List[str] = field(default_factory=list)
    comments: List[str] = field(default_factory=list)
    lengths: np.ndarray = field(default_factory=lambda: np.zeros(0, U64))
    offsets: np.ndarray = field(default_factory=lambda: np.zeros(1, U64))
    hashes: np.ndarray = field(default_factory=lambda: np.zeros(0, U64))  # ascending per reference
    alphabet: str = "ACGT"
    noncanonical: bool = False
    preserve_case: bool = False

    @property
    def n_refs(self) -> int:
        return len(self.offsets) - 1

    @property
    def use64(self) -> bool:
        return use64(self.k, len(self.alphabet) or 4)

    def ref_hashes(self, i: int) -> np.ndarray:
        return self.hashes[int(self.offsets[i]):int(self.offsets[i + 1])]


# --------------------------------------------------------------------------
# writer
# --------------------------------------------------------------------------
def _struct_ptr(off: int, dwords: int, pwords: int) -> int:
    return ((off & 0x3FFFFFFF) << 2) | (dwords << 32) | (pwords << 48)


def _list_ptr(off: int, esize: int, count: int) -> int:
    return 1 | ((off & 0x3FFFFFFF) << 2) | (esize << 32) | (count << 35)


def _far_ptr(seg: int, word: int, double: bool = False) -> int:
    return 2 | (4 if double else 0) | (word << 3) | (seg << 32)


class _Arena:
    def __init__(self, seg_cap_words: int):
        self.cap = seg_cap_words
        self.segs: List[List[np.ndarray]] = [[]]
        self.len: List[int] = [0]

    def _place(self, arr: np.ndarray, new_seg_ok: bool = True):
        if new_seg_ok and self.len[-1] and self.len[-1] + len(arr) > self.cap:
            self.segs.append([])
            self.len.append(0)
        s = len(self.segs) - 1
        w = self.len[s]
        self.segs[s].append(arr)
        self.len[s] += len(arr)
        return s, w

    def alloc(self, arr: np.ndarray, slot, kind_bits, double_far: bool = False):
        """Place ``arr`` and write the pointer for it into ``slot``.

        slot = (holder_array, index_in_holder, holder_seg, holder_abs_word_of_index)
        kind_bits(off) -> pointer word for a same-segment pointer with that offset.
        """
        holder, idx, hseg, hword = slot
        if double_far:
            # object in its own fresh segment, 2-word landing pad in yet another
            self.segs.append([]); self.len.append(0)
            oseg, oword = self._place(arr, new_seg_ok=False)
            self.segs.append([]); self.len.append(0)
            pad = np.zeros(2, U64)
            pseg, pword = self._place(pad, new_seg_ok=False)
            pad[0] = U64(_far_ptr(oseg, oword))
            pad[1] = U64(kind_bits(0))
            holder[idx] = U64(_far_ptr(pseg, pword, double=True))
            self.segs.append([]); self.len.append(0)
            return oseg, oword
        cur = len(self.segs) - 1
        will_move = self.len[cur] and self.len[cur] + len(arr) + 1 > self.cap
        tseg = cur + 1 if will_move else cur
        if tseg == hseg:
            s, w = self._place(arr)
            holder[idx] = U64(kind_bits(w - hword - 1))
            return s, w
        pad = np.zeros(1, U64)
        if will_move:
            self.segs.append([]); self.len.append(0)
        pseg, pword = self._place(pad, new_seg_ok=False)
        s, w = self._place(arr, new_seg_ok=False)
        pad[0] = U64(kind_bits(0))
        holder[idx] = U64(_far_ptr(pseg, pword))
        return s, w


def _text_words(t: str) -> (np.ndarray, int):
    b = t.encode("utf-8") + b"\0"
    n = len(b)
    b += b"\0" * (-n % 8)
    return np.frombuffer(b, dtype="<u8").copy(), n


def write_msh(path: str, db: SketchDB, *, seg_cap_words: int = 1 << 27,
              write_alphabet: bool = True, use_old_list: bool = False,
              double_far_refs: Sequence[int] = ()) -> None:
    """Serialise ``db`` the way ``mash sketch -o`` would (unpacked stream)."""
    n = db.n_refs
    u64 = db.use64
    ar = _Arena(seg_cap_words)
    root = np.zeros(1, U64)
    ar._place(root)
    mh = np.zeros(_MH_D + _MH_P, U64)
    mseg, mword = ar.alloc(mh, (root, 0, 0, 0), lambda o: _struct_ptr(o, _MH_D, _MH_P))
    d = mh[:_MH_D].view(np.uint8)
    d[0:4] = np.frombuffer(struct.pack("<I", db.k), np.uint8)
    d[4:8] = np.frombuffer(struct.pack("<I", db.k), np.uint8)  # windowSize (unused by screen)
    d[8:12] = np.frombuffer(struct.pack("<I", db.s), np.uint8)
    bits = (1 if False else 0) | (2 if db.noncanonical else 0) | (4 if db.preserve_case else 0)
    d[12] = bits
    d[20:24] = np.frombuffer(struct.pack("<I", (db.seed ^ 42) & 0xFFFFFFFF), np.uint8)
    if write_alphabet:
        tw, tn = _text_words(db.alphabet)
        ar.alloc(tw, (mh, _MH_D + 2, mseg, mword + _MH_D + 2), lambda o: _list_ptr(o, 2, tn))
    rl = np.zeros(1, U64)
    pidx = 0 if use_old_list else 3
    rlseg, rlword = ar.alloc(rl, (mh, _MH_D + pidx, mseg, mword + _MH_D + pidx),
                             lambda o: _struct_ptr(o, 0, 1))
    stride = _REF_D + _REF_P
    refs = np.zeros(1 + n * stride, U64)
    refs[0] = U64(_struct_ptr(n, _REF_D, _REF_P))  # composite tag: offset field = element count
    rseg, rword = ar.alloc(refs, (rl, 0, rlseg, rlword), lambda o: _list_ptr(o, 7, n * stride))
    body = refs[1:].reshape(n, stride) if n else refs[1:].reshape(0, stride)
    lengths = np.asarray(db.lengths, dtype=U64)
    if n:
        lo = np.where(lengths < (1 << 32), lengths, 0).astype(U64)
        body[:, 0] = lo                      # length u32 @0 (upper half of the word stays 0)
        body[:, 1] = lengths                 # length64 @8
    dfar = set(int(i) for i in double_far_refs)
    for i in range(n):
        base = rword + 1 + i * stride + _REF_D
        row = body[i]
        tw, tn = _text_words(db.names[i] if db.names else "")
        ar.alloc(tw, (row, _REF_D + 2, rseg, base + 2), lambda o, tn=tn: _list_ptr(o, 2, tn))
        c = db.comments[i] if db.comments else ""
        if c:
            tw, tn = _text_words(c)
            ar.alloc(tw, (row, _REF_D + 3, rseg, base + 3), lambda o, tn=tn: _list_ptr(o, 2, tn))
        h = np.ascontiguousarray(db.ref_hashes(i), dtype=U64)
        cnt = len(h)
        if cnt == 0:
            continue
        if u64:
            ar.alloc(h.copy(), (row, _REF_D + 5, rseg, base + 5),
                     lambda o, cnt=cnt: _list_ptr(o, 5, cnt), double_far=i in dfar)
        else:
            h32 = h.astype(np.uint32)
            if cnt % 2:
                h32 = np.concatenate([h32, np.zeros(1, np.uint32)])
            ar.alloc(h32.view(U64).copy(), (row, _REF_D + 4, rseg, base + 4),
                     lambda o, cnt=cnt: _list_ptr(o, 4, cnt), double_far=i in dfar)
    # drop empty trailing segments
    while len(ar.segs) > 1 and ar.len[-1] == 0:
        ar.segs.pop(); ar.len.pop()
    nseg = len(ar.segs)
    hdr = struct.pack("<I", nseg - 1) + b"".join(struct.pack("<I", L) for L in ar.len)
    hdr += b"\0" * (-len(hdr) % 8)
    with open(path, "wb") as f:
        f.write(hdr)
        for seg in ar.segs:
            for a in seg:
                f.write(a.tobytes())


# --------------------------------------------------------------------------
# reader (independent of the C++ loader; same spec)
# --------------------------------------------------------------------------
class _Msg:
    def __init__(self, buf: bytes):
        nseg = struct.unpack_from("<I", buf, 0)[0] + 1
        sizes = struct.unpack_from("<%dI" % nseg, buf, 4)
        off = (4 + 4 * nseg + 7) // 8 * 8
        self.segs = []
        for sz in sizes:
            self.segs.append(np.frombuffer(buf, dtype="<u8", count=sz, offset=off))
            off += sz * 8
        if off > len(buf):
            raise ValueError("segments exceed file size")

    def resolve(self, seg: int, word: int):
        """Follow the pointer stored at (seg, word) -> dict describing the target."""
        for _ in range(4):
            p = int(self.segs[seg][word])
            if p == 0:
                return None
            t = p & 3
            if t == 2:
                dbl = (p >> 2) & 1
                off = (p >> 3) & 0x1FFFFFFF
                tseg = p >> 32
                if not dbl:
                    seg, word = tseg, off
                    continue
                far = int(self.segs[tseg][off]); tag = int(self.segs[tseg][off + 1])
                if far & 3 != 2:
                    raise ValueError("bad double-far landing pad")
                return self._describe(tag, far >> 32, (far >> 3) & 0x1FFFFFFF)
            off = (p >> 2) & 0x3FFFFFFF
            if off & 0x20000000:
                off -= 1 << 30
            return self._describe(p, seg, word + 1 + off)
        raise ValueError("far pointer chain too long")

    def _describe(self, p: int, seg: int, word: int):
        t = p & 3
        if t == 0:
            return dict(kind="struct", seg=seg, word=word, d=(p >> 32) & 0xFFFF, p=p >> 48)
        if t == 1:
            es = (p >> 32) & 7
            cnt = p >> 35
            if es == 7:
                tag = int(self.segs[seg][word])
                return dict(kind="list", es=7, seg=seg, word=word + 1, count=(tag >> 2) & 0x3FFFFFFF,
                            d=(tag >> 32) & 0xFFFF, p=tag >> 48)
            return dict(kind="list", es=es, seg=seg, word=word, count=cnt)
        raise ValueError("unexpected pointer type")

    def data(self, st, nbytes_off: int, fmt: str):
        size = struct.calcsize(fmt)
        if nbytes_off + size > st["d"] * 8:
            return 0
        raw = self.segs[st["seg"]][st["word"]:st["word"] + st["d"]].tobytes()
        return struct.unpack_from(fmt, raw, nbytes_off)[0]

    def ptr(self, st, idx: int):
        if idx >= st["p"]:
            return None
        return self.resolve(st["seg"], st["word"] + st["d"] + idx)

    def text(self, st, idx: int) -> str:
        t = self.ptr(st, idx)
        if not t or t["kind"] != "list" or t["es"] != 2 or t["count"] == 0:
            return ""
        nw = (t["count"] + 7) // 8
        raw = self.segs[t["seg"]][t["word"]:t["word"] + nw].tobytes()[:t["count"] - 1]
        return raw.decode("utf-8", "replace")


def read_msh(path: str) -> SketchDB:
    with open(path, "rb") as f:
        buf = f.read()
    m = _Msg(buf)
    root = m.resolve(0, 0)
    if not root or root["kind"] != "struct":
        raise ValueError("bad root pointer")
    k = m.data(root, 0, "<I")
    s = m.data(root, 8, "<I")
    seed = m.data(root, 20, "<I") ^ 42
    flags = m.data(root, 12, "<B")
    alphabet = m.text(root, 2) or "ACGT"
    db = SketchDB(k=k, s=s, seed=seed, alphabet=alphabet,
                  noncanonical=bool(flags & 2), preserve_case=bool(flags & 4))
    refs = None
    for pidx in (3, 0):
        rl = m.ptr(root, pidx)
        if rl and rl["kind"] == "struct":
            cand = m.ptr(rl, 0)
            if cand and cand["kind"] == "list" and cand["count"] > 0:
                refs = cand
                break
    names, comments, lengths, chunks, offs = [], [], [], [], [0]
    if refs is not None:
        if refs["es"] != 7:
            raise ValueError("reference list is not composite")
        stride = refs["d"] + refs["p"]
        for i in range(refs["count"]):
            st = dict(kind="struct", seg=refs["seg"], word=refs["word"] + i * stride, d=refs["d"], p=refs["p"])
            l64 = m.data(st, 8, "<Q")
            lengths.append(l64 if l64 else m.data(st, 0, "<I"))
            names.append(m.text(st, 2))
            comments.append(m.text(st, 3))
            hl = m.ptr(st, 5 if db.use64 else 4)
            if hl and hl["count"]:
                if db.use64:
                    h = m.segs[hl["seg"]][hl["word"]:hl["word"] + hl["count"]].astype(U64)
                else:
                    nw = (hl["count"] + 1) // 2
                    h = m.segs[hl["seg"]][hl["word"]:hl["word"] + nw].view(np.uint32)[:hl["count"]].astype(U64)
            else:
                h = np.zeros(0, U64)
            chunks.append(h)
            offs.append(offs[-1] + len(h))
    db.names, db.comments = names, comments
    db.lengths = np.asarray(lengths, dtype=U64)
    db.offsets = np.asarray(offs, dtype=U64)
    db.hashes = np.concatenate(chunks) if chunks else np.zeros(0, U64)
    db._root_shape = (root["d"], root["p"])  # a valid Mash file decodes to (3, 4)
    return db


def _main(argv: Optional[List[str]] = None) -> int:
    """``python -m hymet_b200.msh file.msh`` -- dump header + first references."""
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 1:
        print("usage: python -m hymet_b200.msh <sketch.msh>", file=sys.stderr)
        return 2
    db = read_msh(argv[0])
    print(f"root struct shape (dataWords, ptrWords) = {db._root_shape}  [expect (3, 4)]")
    print(f"k={db.k} s={db.s} seed={db.seed} alphabet={db.alphabet!r} "
          f"noncanonical={db.noncanonical} preserveCase={db.preserve_case} use64={db.use64}")
    print(f"references={db.n_refs} entries={len(db.hashes)}")
    for i in range(min(5, db.n_refs)):
        h = db.ref_hashes(i)
        asc = bool(np.all(h[1:] > h[:-1])) if len(h) > 1 else True
        print(f"  [{i}] {db.names[i]!r} len={int(db.lengths[i])} n={len(h)} ascending={asc} "
              f"first={[hex(int(x)) for x in h[:3]]} comment={db.comments[i][:60]!r}")
    return 0


if __name__ == "__main__":
    sys.exit(_main())
