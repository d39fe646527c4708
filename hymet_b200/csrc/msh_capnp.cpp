// msh_capnp.cpp -- see msh_capnp.h.  Host code, no CUDA.
#include "msh_capnp.h"

#include <fcntl.h>
#include <math.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>

#include "../../include/hymet_screen.h"

namespace hs {
namespace {

struct Obj {
    enum Kind { Null, Struct, List } kind = Null;
    uint32_t seg = 0;
    uint64_t word = 0;        // first word of the content
    uint32_t dwords = 0, pwords = 0;  // struct shape (or composite element shape)
    int esize = 0;            // list element size code
    uint64_t count = 0;       // list element count
};

class Msg {
  public:
    bool init(const uint8_t *base, size_t size, std::string &err)
    {
        if (size < 8) { err = "file too small to be a sketch"; return false; }
        uint32_t nseg_m1;
        memcpy(&nseg_m1, base, 4);
        const uint64_t nseg = (uint64_t)nseg_m1 + 1;
        if (nseg > (1u << 22) || 4 + 4 * nseg > size) { err = "bad segment table"; return false; }
        uint64_t off = (4 + 4 * nseg + 7) / 8 * 8;
        segs_.resize(nseg);
        for (uint64_t i = 0; i < nseg; i++) {
            uint32_t w;
            memcpy(&w, base + 4 + 4 * i, 4);
            if (off + (uint64_t)w * 8 > size) { err = "segment exceeds file size (truncated .msh?)"; return false; }
            segs_[i] = {reinterpret_cast<const uint64_t *>(base + off), w};
            off += (uint64_t)w * 8;
        }
        return true;
    }

    // Decode the pointer stored at (seg, word).
    bool resolve(uint32_t seg, uint64_t word, Obj &o, std::string &err) const
    {
        for (int hop = 0; hop < 3; hop++) {
            if (!in(seg, word, 1)) { err = "pointer outside its segment"; return false; }
            const uint64_t p = segs_[seg].w[word];
            if (p == 0) { o = Obj(); return true; }
            const int type = (int)(p & 3);
            if (type == 2) {  // far pointer
                const bool dbl = (p >> 2) & 1;
                const uint64_t pad = (p >> 3) & 0x1FFFFFFFull;
                const uint32_t tseg = (uint32_t)(p >> 32);
                if (!dbl) { seg = tseg; word = pad; continue; }
                if (!in(tseg, pad, 2)) { err = "double-far landing pad out of range"; return false; }
                const uint64_t far = segs_[tseg].w[pad], tag = segs_[tseg].w[pad + 1];
                if ((far & 3) != 2 || ((far >> 2) & 1)) { err = "malformed double-far landing pad"; return false; }
                return describe(tag, (uint32_t)(far >> 32), (far >> 3) & 0x1FFFFFFFull, o, err);
            }
            if (type == 3) { err = "capability pointer in a sketch file"; return false; }
            int64_t off = (int64_t)((p >> 2) & 0x3FFFFFFFull);
            if (off & 0x20000000ll) off -= 0x40000000ll;
            const int64_t target = (int64_t)word + 1 + off;
            if (target < 0) { err = "negative pointer target"; return false; }
            return describe(p, seg, (uint64_t)target, o, err);
        }
        err = "far-pointer chain too long";
        return false;
    }

    bool ptr(const Obj &st, uint32_t idx, Obj &o, std::string &err) const
    {
        if (idx >= st.pwords) { o = Obj(); return true; }  // older/smaller struct: field absent
        return resolve(st.seg, st.word + st.dwords + idx, o, err);
    }
    template <class T> T data(const Obj &st, uint32_t byte_off) const
    {
        T v = 0;
        if ((uint64_t)byte_off + sizeof(T) <= (uint64_t)st.dwords * 8)
            memcpy(&v, reinterpret_cast<const uint8_t *>(segs_[st.seg].w + st.word) + byte_off, sizeof(T));
        return v;
    }
    bool text(const Obj &st, uint32_t idx, std::string &out, std::string &err) const
    {
        Obj t;
        out.clear();
        if (!ptr(st, idx, t, err)) return false;
        if (t.kind != Obj::List || t.esize != 2 || t.count == 0) return true;
        const char *c = reinterpret_cast<const char *>(segs_[t.seg].w + t.word);
        out.assign(c, strnlen(c, t.count - 1));
        return true;
    }
    const uint64_t *words(const Obj &o) const { return segs_[o.seg].w + o.word; }

  private:
    struct Seg { const uint64_t *w; uint32_t n; };
    std::vector<Seg> segs_;
    bool in(uint32_t seg, uint64_t word, uint64_t n) const
    {
        return seg < segs_.size() && word + n <= segs_[seg].n && word + n >= word;
    }
    bool describe(uint64_t p, uint32_t seg, uint64_t word, Obj &o, std::string &err) const
    {
        o = Obj();
        o.seg = seg; o.word = word;
        const int type = (int)(p & 3);
        if (type == 0) {
            o.kind = Obj::Struct;
            o.dwords = (uint32_t)((p >> 32) & 0xFFFF); o.pwords = (uint32_t)(p >> 48);
            if (!in(seg, word, (uint64_t)o.dwords + o.pwords)) { err = "struct outside its segment"; return false; }
            return true;
        }
        if (type != 1) { err = "unexpected pointer type"; return false; }
        o.kind = Obj::List;
        o.esize = (int)((p >> 32) & 7);
        o.count = p >> 35;
        uint64_t nwords;
        if (o.esize == 7) {  // composite: `count` is the word count, the tag has the element count
            if (!in(seg, word, 1)) { err = "composite tag outside its segment"; return false; }
            nwords = o.count;
            const uint64_t tag = segs_[seg].w[word];
            o.count = (tag >> 2) & 0x3FFFFFFFull;
            o.dwords = (uint32_t)((tag >> 32) & 0xFFFF); o.pwords = (uint32_t)(tag >> 48);
            o.word = word + 1;
            if ((uint64_t)o.count * (o.dwords + o.pwords) > nwords) { err = "composite list overruns its words"; return false; }
        } else {
            static const int bits[8] = {0, 1, 8, 16, 32, 64, 64, 0};
            nwords = (o.count * (uint64_t)bits[o.esize] + 63) / 64;
        }
        if (!in(seg, o.word, nwords)) { err = "list outside its segment"; return false; }
        return true;
    }
};

struct Mapping {
    void *p = MAP_FAILED;
    size_t n = 0;
    ~Mapping() { if (p != MAP_FAILED) munmap(p, n); }
};

}  // namespace

int msh_read(const std::string &path, MshData &out, std::string &err)
{
    const auto t0 = std::chrono::steady_clock::now();
    int fd = open(path.c_str(), O_RDONLY);
    if (fd < 0) { err = "could not open " + path; return HS_EIO; }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size <= 0) { close(fd); err = path + ": empty or unreadable"; return HS_EIO; }
    Mapping map;
    map.n = (size_t)sb.st_size;
    map.p = mmap(nullptr, map.n, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map.p == MAP_FAILED) { err = "mmap failed for " + path; return HS_EIO; }

    Msg m;
    if (!m.init(static_cast<const uint8_t *>(map.p), map.n, err)) { err = path + ": " + err; return HS_EFORMAT; }
    Obj root;
    if (!m.resolve(0, 0, root, err) || root.kind != Obj::Struct) {
        err = path + ": bad root pointer" + (err.empty() ? "" : " (" + err + ")");
        return HS_EFORMAT;
    }
    out = MshData();
    out.k = m.data<uint32_t>(root, 0);
    out.s = m.data<uint32_t>(root, 8);
    out.seed = m.data<uint32_t>(root, 20) ^ 42u;  // Cap'n Proto stores value XOR default
    const uint8_t flags = m.data<uint8_t>(root, 12);
    std::string alphabet;
    if (!m.text(root, 2, alphabet, err)) { err = path + ": " + err; return HS_EFORMAT; }
    if (out.k == 0 || out.k > 32) { err = path + ": k-mer size " + std::to_string(out.k) + " unsupported (1..32)"; return HS_EUNSUPPORTED; }
    if ((flags & 2) || (flags & 4) || (!alphabet.empty() && alphabet != "ACGT")) {
        // S22: refuse rather than silently mis-screen
        err = path + ": unsupported sketch type (noncanonical / preserveCase / non-nucleotide alphabet)";
        return HS_EUNSUPPORTED;
    }
    out.use64 = pow(4.0, (double)out.k) > pow(2.0, 32.0);  // S1

    Obj refs;
    bool have = false;
    for (uint32_t pidx : {3u, 0u}) {  // referenceList, else referenceListOld
        Obj rl, cand;
        if (!m.ptr(root, pidx, rl, err)) { err = path + ": " + err; return HS_EFORMAT; }
        if (rl.kind != Obj::Struct) continue;
        if (!m.ptr(rl, 0, cand, err)) { err = path + ": " + err; return HS_EFORMAT; }
        if (cand.kind == Obj::List && cand.count > 0) { refs = cand; have = true; break; }
    }
    out.offsets.assign(1, 0);
    if (have) {
        if (refs.esize != 7) { err = path + ": reference list is not a composite list"; return HS_EFORMAT; }
        const uint64_t n = refs.count, stride = (uint64_t)refs.dwords + refs.pwords;
        out.names.resize(n); out.comments.resize(n); out.lengths.resize(n); out.offsets.resize(n + 1);
        std::vector<Obj> lists(n);
        const uint32_t hidx = out.use64 ? 5u : 4u;
        for (uint64_t i = 0; i < n; i++) {
            Obj r; r.kind = Obj::Struct; r.seg = refs.seg; r.word = refs.word + i * stride;
            r.dwords = refs.dwords; r.pwords = refs.pwords;
            const uint64_t l64 = m.data<uint64_t>(r, 8);
            out.lengths[i] = l64 ? l64 : m.data<uint32_t>(r, 0);
            if (!m.text(r, 2, out.names[i], err) || !m.text(r, 3, out.comments[i], err) ||
                !m.ptr(r, hidx, lists[i], err)) { err = path + ": " + err; return HS_EFORMAT; }
            uint64_t cnt = 0;
            if (lists[i].kind == Obj::List) {
                if (lists[i].esize != (out.use64 ? 5 : 4)) { err = path + ": hash list has the wrong element size"; return HS_EFORMAT; }
                cnt = lists[i].count;
            }
            out.offsets[i + 1] = out.offsets[i] + cnt;
        }
        out.hashes.resize(out.offsets[n]);
        for (uint64_t i = 0; i < n; i++) {
            const uint64_t cnt = out.offsets[i + 1] - out.offsets[i];
            if (!cnt) continue;
            uint64_t *dst = out.hashes.data() + out.offsets[i];
            if (out.use64) {
                memcpy(dst, m.words(lists[i]), cnt * 8);
            } else {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(m.words(lists[i]));
                for (uint64_t j = 0; j < cnt; j++) dst[j] = src[j];
            }
        }
    }
    out.t_parse_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return HS_OK;
}

}  // namespace hs
