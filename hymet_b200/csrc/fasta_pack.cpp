// fasta_pack.cpp -- see fasta_pack.h.  Host code, no CUDA.
#include "fasta_pack.h"

#include <string.h>
#include <zlib.h>

namespace hs {

namespace {

struct Lut {
    uint8_t v[256];
    Lut()
    {
        for (int i = 0; i < 256; i++) v[i] = 4;  // not in the alphabet (S4) -> invalid position
        v['A'] = v['a'] = 0;                     // S3: case folded
        v['C'] = v['c'] = 1;
        v['G'] = v['g'] = 2;
        v['T'] = v['t'] = 3;
    }
};
const Lut kLut;

struct Writer {
    uint64_t *seq;
    uint32_t *inv;
    uint64_t w = 0, acc = 0;
    uint32_t iacc = 0;
    int cnt = 0;
    uint64_t positions = 0;
    inline void push(uint32_t code)
    {
        acc = (acc << 2) | (code & 3u);
        iacc = (iacc << 1) | (code >> 2);
        if (++cnt == 32) {
            seq[w] = acc; inv[w] = iacc; w++;
            cnt = 0; acc = 0; iacc = 0;
        }
        positions++;
    }
    inline void finish()
    {
        if (cnt) {
            const int pad = 32 - cnt;
            seq[w] = acc << (2 * pad);
            inv[w] = (iacc << pad) | ((pad >= 32) ? ~0u : ((1u << pad) - 1u));
            w++;
            cnt = 0; acc = 0; iacc = 0;
        }
    }
};

}  // namespace

uint64_t pack_text_span(const char *t, size_t n, uint64_t *seq, uint32_t *inv, PackStats *st)
{
    Writer wr{seq, inv};
    enum { SEEK_HDR, IN_SEQ } state = SEEK_HDR;
    bool fastq = false;
    uint64_t rec_len = 0;
    size_t i = 0;
    while (i < n) {
        const char *nl = (const char *)memchr(t + i, '\n', n - i);
        size_t j = nl ? (size_t)(nl - t) : n;  // line = [i, j)
        const char c0 = t[i];
        if (c0 == '>' || c0 == '@') {
            // header line: one invalid position keeps k-mers from bridging records (S6)
            state = IN_SEQ;
            fastq = (c0 == '@');
            rec_len = 0;
            wr.push(4);
            if (st) st->n_records++;
            i = j + 1;
            continue;
        }
        if (state == SEEK_HDR) { i = j + 1; continue; }
        if (c0 == '+') {
            i = j + 1;  // the '+' line
            if (fastq) {  // quality block: as many symbols as the sequence had
                uint64_t q = 0;
                while (i < n && q < rec_len) {
                    if (t[i] != '\n' && t[i] != '\r') q++;
                    i++;
                }
                nl = i < n ? (const char *)memchr(t + i, '\n', n - i) : nullptr;
                i = nl ? (size_t)(nl - t) + 1 : n;
            }
            state = SEEK_HDR;
            continue;
        }
        size_t L = j - i;
        if (L && t[i + L - 1] == '\r') L--;  // kseq: one trailing CR dropped, everything else kept
        const unsigned char *p = (const unsigned char *)t + i;
        for (size_t q = 0; q < L; q++) wr.push(kLut.v[p[q]]);
        rec_len += L;
        if (st) st->n_seq_bases += L;
        i = j + 1;
    }
    wr.finish();
    if (st) st->n_positions += wr.positions;
    return wr.w;
}

std::vector<std::pair<size_t, size_t>> split_records(const char *t, size_t n, int parts, size_t min_span)
{
    std::vector<std::pair<size_t, size_t>> out;
    size_t first = 0;
    while (first < n && (t[first] == '\n' || t[first] == '\r')) first++;
    if (parts < 1) parts = 1;
    if (first < n && t[first] != '>') parts = 1;  // FASTQ ('@' also occurs in quality lines): do not split
    size_t span = n / (size_t)parts + 1;
    if (span < min_span) span = min_span;
    size_t beg = 0;
    while (beg < n) {
        size_t end = beg + span;
        if (end >= n || parts == 1) {
            end = n;
        } else {
            // advance to the next line that starts a record
            for (;;) {
                const char *nl = (const char *)memchr(t + end, '\n', n - end);
                if (!nl) { end = n; break; }
                end = (size_t)(nl - t) + 1;
                if (end >= n || t[end] == '>') break;
            }
        }
        out.emplace_back(beg, end);
        beg = end;
    }
    return out;
}

bool slurp_file(const std::string &path, std::vector<char> &out, std::string &err)
{
    gzFile f = path == "-" ? gzdopen(0, "rb") : gzopen(path.c_str(), "rb");
    if (!f) { err = "could not open " + path; return false; }
    gzbuffer(f, 1 << 20);
    size_t len = out.size();
    for (;;) {
        if (out.size() - len < (1u << 22)) out.resize(out.size() ? out.size() * 2 : (1u << 24));
        int r = gzread(f, out.data() + len, 1u << 22);
        if (r < 0) { err = "read error on " + path; gzclose(f); return false; }
        if (r == 0) break;
        len += (size_t)r;
    }
    gzclose(f);
    out.resize(len);
    return true;
}

}  // namespace hs
