// fasta_pack.cpp -- see fasta_pack.h.  Host code, no CUDA.
//
// The packer is the end-to-end limiter (the GPU consumes > 100 Gbp/s), so sequence
// lines are converted 32 characters at a time with AVX2 when the host has it
// (runtime dispatch; the scalar loop is the same arithmetic one character at a time).
#include "fasta_pack.h"

#include <string.h>
#include <zlib.h>

#if defined(__x86_64__)
#include <immintrin.h>
#define HS_X86 1
#endif

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

// Left-aligned accumulator: base i of the pending word sits at bits [62-2i, 63-2i],
// its invalid flag at bit (31-i).
struct Writer {
    uint64_t *seq;
    uint32_t *inv;
    uint64_t w = 0, acc = 0;
    uint32_t iacc = 0;
    int cnt = 0;
    uint64_t positions = 0;

    inline void push(uint32_t code)
    {
        acc |= (uint64_t)(code & 3u) << (62 - 2 * cnt);
        iacc |= (code >> 2) << (31 - cnt);
        if (++cnt == 32) {
            seq[w] = acc; inv[w] = iacc; w++;
            cnt = 0; acc = 0; iacc = 0;
        }
        positions++;
    }
    // n (1..32) bases: codes left-aligned in c, flags left-aligned in m, unused low bits zero
    inline void append(uint64_t c, uint32_t m, int n)
    {
        if (cnt) { acc |= c >> (2 * cnt); iacc |= m >> cnt; } else { acc = c; iacc = m; }
        if (cnt + n >= 32) {
            seq[w] = acc; inv[w] = iacc; w++;
            const int used = 32 - cnt;
            if (used < 32) { acc = c << (2 * used); iacc = m << used; } else { acc = 0; iacc = 0; }
            cnt = cnt + n - 32;
        } else {
            cnt += n;
        }
        positions += (uint64_t)n;
    }
    inline void finish()
    {
        if (cnt) {
            seq[w] = acc;
            inv[w] = iacc | ((1u << (32 - cnt)) - 1u);  // padding is invalid
            w++;
            cnt = 0; acc = 0; iacc = 0;
        }
    }
};

void line_scalar(Writer &wr, const unsigned char *p, size_t L)
{
    for (size_t q = 0; q < L; q++) wr.push(kLut.v[p[q]]);
}

#ifdef HS_X86
// 32 characters -> 64 bits of 2-bit codes (first base most significant) + 32 invalid flags.
__attribute__((target("avx2"))) inline void convert32(__m256i v, uint64_t &codes, uint32_t &bad)
{
    const __m256i up = _mm256_and_si256(v, _mm256_set1_epi8((char)0xDF));            // fold case (S3)
    const __m256i c1 = _mm256_and_si256(_mm256_srli_epi16(up, 1), _mm256_set1_epi8(3));  // A0 C1 T2 G3
    const __m256i code = _mm256_xor_si256(c1, _mm256_and_si256(_mm256_srli_epi16(c1, 1), _mm256_set1_epi8(1)));  // A0 C1 G2 T3
    const __m256i letters = _mm256_setr_epi8('A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
                                             'A', 'C', 'G', 'T', 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i ok = _mm256_cmpeq_epi8(up, _mm256_shuffle_epi8(letters, code));   // S4: exactly A/C/G/T
    // reverse byte order inside each 128-bit half so movemask puts base 0 at the top
    const __m256i rev = _mm256_setr_epi8(15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0,
                                         15, 14, 13, 12, 11, 10, 9, 8, 7, 6, 5, 4, 3, 2, 1, 0);
    const uint32_t mm = ~(uint32_t)_mm256_movemask_epi8(_mm256_shuffle_epi8(ok, rev));
    bad = (mm << 16) | (mm >> 16);
    // four codes -> one byte (first base in the top two bits), then 8 bytes -> big-endian u64
    const __m256i w16 = _mm256_maddubs_epi16(_mm256_and_si256(code, ok),  // invalid positions carry code 0
                                             _mm256_set1_epi32(0x01041040));  // (c0*64+c1*16), (c2*4+c3)
    const __m256i w32 = _mm256_madd_epi16(w16, _mm256_set1_epi16(1));
    const __m256i b16 = _mm256_packus_epi32(w32, w32);
    const __m256i b8 = _mm256_packus_epi16(b16, b16);
    const uint64_t lo = (uint32_t)_mm256_extract_epi32(b8, 0), hi = (uint32_t)_mm256_extract_epi32(b8, 4);
    codes = __builtin_bswap64(lo | (hi << 32));
}

__attribute__((target("avx2"))) void line_avx2(Writer &wr, const unsigned char *p, size_t L, const unsigned char *safe_end)
{
    size_t off = 0;
    if (L >= 32) {
        // full vectors: the bit offset inside the pending word is constant along the line,
        // so the accumulator stays in registers and the loop has no branches
        uint64_t acc = wr.acc, w = wr.w;
        uint32_t iacc = wr.iacc;
        uint64_t *seq = wr.seq;
        uint32_t *inv = wr.inv;
        const int r = wr.cnt;
        if (r == 0) {
            for (; off + 32 <= L; off += 32, w++) {
                uint64_t c; uint32_t b;
                convert32(_mm256_loadu_si256((const __m256i *)(p + off)), c, b);
                seq[w] = c; inv[w] = b;
            }
        } else {
            for (; off + 32 <= L; off += 32, w++) {
                uint64_t c; uint32_t b;
                convert32(_mm256_loadu_si256((const __m256i *)(p + off)), c, b);
                seq[w] = acc | (c >> (2 * r)); inv[w] = iacc | (b >> r);
                acc = c << (64 - 2 * r); iacc = b << (32 - r);
            }
        }
        wr.acc = acc; wr.iacc = iacc; wr.w = w;
        wr.positions += off;
    }
    const int n = (int)(L - off);
    if (n) {
        __m256i v;
        if (p + off + 32 <= safe_end) {
            v = _mm256_loadu_si256((const __m256i *)(p + off));
        } else {  // never read past the caller's buffer
            unsigned char tmp[32] = {0};
            memcpy(tmp, p + off, (size_t)n);
            v = _mm256_loadu_si256((const __m256i *)tmp);
        }
        uint64_t c; uint32_t b;
        convert32(v, c, b);
        wr.append(c & (~0ull << (64 - 2 * n)), b & (~0u << (32 - n)), n);
    }
}
#endif

}  // namespace

uint64_t pack_text_span(const char *t, size_t n, uint64_t *seq, uint32_t *inv, PackStats *st)
{
    Writer wr{seq, inv};
#ifdef HS_X86
    static const bool have_avx2 = __builtin_cpu_supports("avx2");
#else
    const bool have_avx2 = false;
#endif
    const unsigned char *safe_end = (const unsigned char *)t + n;
    enum { SEEK_HDR, IN_SEQ } state = SEEK_HDR;
    bool fastq = false;
    uint64_t rec_len = 0, n_records = 0, n_seq = 0;
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
            n_records++;
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
#ifdef HS_X86
        if (have_avx2) line_avx2(wr, p, L, safe_end); else
#endif
            line_scalar(wr, p, L);
        rec_len += L;
        n_seq += L;
        i = j + 1;
    }
    wr.finish();
    if (st) { st->n_records += n_records; st->n_seq_bases += n_seq; st->n_positions += wr.positions; }
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
