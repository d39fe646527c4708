/*
 * oracle/mash_screen_oracle.c  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, CPU-only restatement of the `mash screen` containment pass that
 * HYMET runs at /root/reference/scripts/mash.sh:14
 *     mash screen -p 8 -v 0.9 "$MASH_SCREEN" "$INPUT_DIR"/ *.fna > "$SCREEN_TAB"
 * and whose TSV is consumed at scripts/mash.sh:15-55,
 * scripts/limit_candidates.py:97-122 and scripts/downloadDB.py:106-111.
 *
 * PARITY STATUS: **parity unpinned by the reference**.  The arithmetic of this
 * path lives in the third-party `mash` binary (marbl/Mash, bioconda, version
 * unpinned at /root/reference/environment.yml:9 and run_hymet_cami.sh:72;
 * target semantics: Mash v2.3).  Its source is not under /root/reference and
 * no test or fixture of the reference pins its output (tests/test_cli.py is
 * dry-run only).  This file therefore restates Mash's *published* algorithm
 * (Ondov et al. 2016 "Mash", Ondov et al. 2019 "Mash Screen"; rules S1-S22 of
 * SURVEY.md Appendix A) and is pinned against
 *   - the public MurmurHash3_x64_128 known-answer vectors (SURVEY Appendix C;
 *     cross-checked in tests against Appleby's reference build when present),
 *   - the identity / p-value known answers (Appendix C; p-values re-derived
 *     with mpmath at 50 digits, tests/golden/make_golden.py),
 *   - an independent pure-Python restatement (oracle/py_micro_oracle.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this code, and only as the checker or as the
 * timed CPU baseline.  The product (hymet_b200/) never links or calls it.
 *
 * Deliberately written the way the CPU tool works (ASCII strings, a
 * reverse-complement copy, memcmp for the canonical choice, a CPU hash map)
 * and NOT the way the CUDA path works (2-bit rolling words), so that the two
 * are independent statements of the same rules.
 *
 * Build: see oracle/Makefile  (gcc -O3 -pthread ... -lz -lm)
 */
#define _GNU_SOURCE
#include <ctype.h>
#include <errno.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <zlib.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------ */
/* S2: MurmurHash3_x64_128 (public-domain algorithm by Austin Appleby,       */
/* restated).  Mash's getHash(): seed from the sketch (default 42), keep the  */
/* first 8 output bytes (h1) in the 64-bit regime, the low 32 bits of h1      */
/* when alphabet^k <= 2^32 (S1).                                              */
/* ------------------------------------------------------------------------ */
static inline uint64_t rotl64(uint64_t x, int r) { return (x << r) | (x >> (64 - r)); }

static inline uint64_t fmix64(uint64_t k)
{
    k ^= k >> 33;
    k *= 0xff51afd7ed558ccdULL;
    k ^= k >> 33;
    k *= 0xc4ceb9fe1a85ec53ULL;
    k ^= k >> 33;
    return k;
}

ORC_API void orc_murmur3_x64_128(const void *key, int len, uint32_t seed, uint64_t out[2])
{
    const uint8_t *data = (const uint8_t *)key;
    const int nblocks = len / 16;
    uint64_t h1 = seed, h2 = seed;
    const uint64_t c1 = 0x87c37b91114253d5ULL, c2 = 0x4cf5ad432745937fULL;

    for (int i = 0; i < nblocks; i++) {
        uint64_t k1, k2;
        memcpy(&k1, data + 16 * i, 8); /* little-endian host assumed (x86-64) */
        memcpy(&k2, data + 16 * i + 8, 8);
        k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
        h1 = rotl64(h1, 27); h1 += h2; h1 = h1 * 5 + 0x52dce729;
        k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2;
        h2 = rotl64(h2, 31); h2 += h1; h2 = h2 * 5 + 0x38495ab5;
    }
    const uint8_t *tail = data + nblocks * 16;
    uint64_t k1 = 0, k2 = 0;
    switch (len & 15) {
    case 15: k2 ^= (uint64_t)tail[14] << 48; /* fallthrough */
    case 14: k2 ^= (uint64_t)tail[13] << 40; /* fallthrough */
    case 13: k2 ^= (uint64_t)tail[12] << 32; /* fallthrough */
    case 12: k2 ^= (uint64_t)tail[11] << 24; /* fallthrough */
    case 11: k2 ^= (uint64_t)tail[10] << 16; /* fallthrough */
    case 10: k2 ^= (uint64_t)tail[9] << 8;   /* fallthrough */
    case 9:  k2 ^= (uint64_t)tail[8];
             k2 *= c2; k2 = rotl64(k2, 33); k2 *= c1; h2 ^= k2; /* fallthrough */
    case 8:  k1 ^= (uint64_t)tail[7] << 56; /* fallthrough */
    case 7:  k1 ^= (uint64_t)tail[6] << 48; /* fallthrough */
    case 6:  k1 ^= (uint64_t)tail[5] << 40; /* fallthrough */
    case 5:  k1 ^= (uint64_t)tail[4] << 32; /* fallthrough */
    case 4:  k1 ^= (uint64_t)tail[3] << 24; /* fallthrough */
    case 3:  k1 ^= (uint64_t)tail[2] << 16; /* fallthrough */
    case 2:  k1 ^= (uint64_t)tail[1] << 8;  /* fallthrough */
    case 1:  k1 ^= (uint64_t)tail[0];
             k1 *= c1; k1 = rotl64(k1, 31); k1 *= c2; h1 ^= k1;
    }
    h1 ^= (uint64_t)len; h2 ^= (uint64_t)len;
    h1 += h2; h2 += h1;
    h1 = fmix64(h1); h2 = fmix64(h2);
    h1 += h2; h2 += h1;
    out[0] = h1; out[1] = h2;
}

/* S1: 64-bit hashes iff 4^k > 2^32, i.e. k >= 17 for nucleotides. */
ORC_API int orc_use64(uint32_t k) { return pow(4.0, (double)k) > pow(2.0, 32.0); }

static inline uint64_t mash_hash(const char *kmer, uint32_t k, uint32_t seed, int use64)
{
    uint64_t h[2];
    orc_murmur3_x64_128(kmer, (int)k, seed, h);
    return use64 ? h[0] : (h[0] & 0xffffffffULL);
}

/* ------------------------------------------------------------------------ */
/* S3-S5: hashSequence.  Upper-case, reverse-complement copy, every window    */
/* of k alphabet-only characters, canonical = memcmp(fwd, rc) <= 0 ? fwd : rc */
/* ------------------------------------------------------------------------ */
static inline char comp_base(char c)
{
    switch (c) {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default:  return c; /* non-alphabet characters never enter a hashed window */
    }
}
static inline int is_acgt(char c) { return c == 'A' || c == 'C' || c == 'G' || c == 'T'; }

/* Per-position hashes of one record (for K1 parity tests).
 * out_hash[i], out_valid[i] for window starting at i, i in [0, n-k]. */
ORC_API int orc_hash_sequence(const char *seq, uint64_t n, uint32_t k, uint32_t seed,
                              uint64_t *out_hash, uint8_t *out_valid)
{
    if (n < k) return 0;
    const int use64 = orc_use64(k);
    char *fwd = (char *)malloc(n), *rc = (char *)malloc(n);
    if (!fwd || !rc) { free(fwd); free(rc); return -1; }
    for (uint64_t i = 0; i < n; i++) fwd[i] = (char)toupper((unsigned char)seq[i]);
    for (uint64_t i = 0; i < n; i++) rc[i] = comp_base(fwd[n - 1 - i]);
    for (uint64_t i = 0; i + k <= n; i++) {
        int ok = 1;
        for (uint32_t j = 0; j < k; j++) if (!is_acgt(fwd[i + j])) { ok = 0; break; }
        out_valid[i] = (uint8_t)ok;
        out_hash[i] = 0;
        if (!ok) continue;
        const char *f = fwd + i, *r = rc + (n - i - k);
        const char *kmer = memcmp(f, r, k) <= 0 ? f : r;
        out_hash[i] = mash_hash(kmer, k, seed, use64);
    }
    free(fwd); free(rc);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Input: FASTA / FASTQ, plain or gzip, multi-line (S6).                      */
/* ------------------------------------------------------------------------ */
typedef struct { uint64_t off, len; } rec_t;
typedef struct {
    char *seq;       /* all record sequences back to back (raw case) */
    uint64_t nseq, cap;
    rec_t *recs; uint64_t nrec, rcap;
} seqset_t;

static int ss_push_rec(seqset_t *ss, uint64_t off, uint64_t len)
{
    if (ss->nrec == ss->rcap) {
        ss->rcap = ss->rcap ? ss->rcap * 2 : 1024;
        ss->recs = (rec_t *)realloc(ss->recs, ss->rcap * sizeof(rec_t));
        if (!ss->recs) return -1;
    }
    ss->recs[ss->nrec].off = off; ss->recs[ss->nrec].len = len; ss->nrec++;
    return 0;
}
static int ss_reserve(seqset_t *ss, uint64_t extra)
{
    if (ss->nseq + extra > ss->cap) {
        uint64_t nc = ss->cap ? ss->cap : (1u << 20);
        while (nc < ss->nseq + extra) nc *= 2;
        ss->seq = (char *)realloc(ss->seq, nc);
        if (!ss->seq) return -1;
        ss->cap = nc;
    }
    return 0;
}

/* Parse a whole text buffer.  Record = header line starting with '>' or '@';
 * sequence lines are joined; for '@' records a '+' line starts a quality
 * block of the same length which is skipped. */
static int parse_text(seqset_t *ss, const char *t, uint64_t n)
{
    uint64_t i = 0;
    while (i < n) {
        /* find a header */
        while (i < n && t[i] != '>' && t[i] != '@') { while (i < n && t[i] != '\n') i++; i++; }
        if (i >= n) break;
        const int fastq = t[i] == '@';
        while (i < n && t[i] != '\n') i++; /* skip header line */
        i++;
        const uint64_t off = ss->nseq;
        while (i < n && t[i] != '>' && t[i] != '@' && t[i] != '+') {
            uint64_t j = i;
            while (j < n && t[j] != '\n') j++;
            uint64_t L = j - i;
            if (ss_reserve(ss, L)) return -1;
            /* kseq semantics: keep every byte of the line (blanks included: they are
             * simply non-alphabet characters), drop only one trailing '\r' */
            if (L && t[i + L - 1] == '\r') L--;
            memcpy(ss->seq + ss->nseq, t + i, L);
            ss->nseq += L;
            i = j + 1;
        }
        const uint64_t len = ss->nseq - off;
        if (ss_push_rec(ss, off, len)) return -1;
        if (fastq && i < n && t[i] == '+') {
            while (i < n && t[i] != '\n') i++;
            i++;
            uint64_t q = 0;
            while (i < n && q < len) { if (t[i] != '\n' && t[i] != '\r') q++; i++; }
            while (i < n && t[i] != '\n') i++;
            i++;
        }
    }
    return 0;
}

static int slurp_gz(const char *path, char **out, uint64_t *n)
{
    gzFile f = strcmp(path, "-") == 0 ? gzdopen(0, "rb") : gzopen(path, "rb");
    if (!f) return -1;
    gzbuffer(f, 1 << 20);
    uint64_t cap = 1 << 22, len = 0;
    char *buf = (char *)malloc(cap);
    for (;;) {
        if (cap - len < (1 << 20)) { cap *= 2; buf = (char *)realloc(buf, cap); }
        if (!buf) { gzclose(f); return -1; }
        int r = gzread(f, buf + len, 1 << 20);
        if (r < 0) { free(buf); gzclose(f); return -1; }
        if (r == 0) break;
        len += (uint64_t)r;
    }
    gzclose(f);
    *out = buf; *n = len;
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Bottom-s distinct hashes (S9) -- mixture sketch and `mash sketch`.          */
/* A candidate buffer cut back to the s smallest distinct values whenever it  */
/* fills; `thr` is the largest kept value once s values are held.             */
/* ------------------------------------------------------------------------ */
typedef struct { uint64_t *v; uint64_t n, cap, s; uint64_t thr; int full; } bottom_t;

static int cmp_u64(const void *a, const void *b)
{
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}
static void bottom_init(bottom_t *b, uint64_t s)
{
    b->s = s; b->cap = 4 * s + 64; b->n = 0; b->full = 0; b->thr = ~0ULL;
    b->v = (uint64_t *)malloc(b->cap * sizeof(uint64_t));
}
static void bottom_compact(bottom_t *b)
{
    qsort(b->v, b->n, sizeof(uint64_t), cmp_u64);
    uint64_t m = 0;
    for (uint64_t i = 0; i < b->n; i++)
        if (m == 0 || b->v[i] != b->v[m - 1]) b->v[m++] = b->v[i];
    if (m > b->s) m = b->s;
    b->n = m;
    if (m == b->s && m > 0) { b->full = 1; b->thr = b->v[m - 1]; }
}
static inline void bottom_try(bottom_t *b, uint64_t h)
{
    if (b->full && h >= b->thr) return; /* cannot enter (equal => duplicate or not smaller) */
    b->v[b->n++] = h;
    if (b->n == b->cap) bottom_compact(b);
}
static void bottom_free(bottom_t *b) { free(b->v); b->v = NULL; }

/* ------------------------------------------------------------------------ */
/* Sketch DB (S7) and its CPU hash map  hash -> dense id.                      */
/* ------------------------------------------------------------------------ */
typedef struct orc_db {
    uint32_t k, s, seed;
    int use64;
    uint64_t n_refs, n_entries, n_distinct;
    uint64_t *offsets;  /* n_refs + 1 */
    uint64_t *hashes;   /* n_entries, ascending within each reference */
    uint64_t *lengths;  /* n_refs (S18) */
    char **names, **comments; /* may be NULL */
    /* map */
    uint64_t cap;       /* power of two */
    uint64_t *keys;
    uint32_t *ids;      /* dense id + 1; 0 = empty slot */
    uint32_t *entry_id; /* n_entries: dense id of every stored hash */
    /* inverted lists: distinct id -> reference indices, ascending */
    uint64_t *inv_off;  /* n_distinct + 1 */
    uint32_t *inv_ref;
} orc_db;

static inline uint64_t mix64(uint64_t x)
{
    x ^= x >> 32; x *= 0x9E3779B97F4A7C15ULL; x ^= x >> 29;
    return x;
}

static int db_build_map(orc_db *db)
{
    uint64_t cap = 16;
    while (cap < 2 * db->n_entries + 16) cap *= 2;
    db->cap = cap;
    db->keys = (uint64_t *)calloc(cap, sizeof(uint64_t));
    db->ids = (uint32_t *)calloc(cap, sizeof(uint32_t));
    db->entry_id = (uint32_t *)malloc((db->n_entries + 1) * sizeof(uint32_t));
    if (!db->keys || !db->ids || !db->entry_id) return -1;
    uint64_t d = 0;
    for (uint64_t e = 0; e < db->n_entries; e++) {
        uint64_t h = db->hashes[e], p = mix64(h) & (cap - 1);
        while (db->ids[p] && db->keys[p] != h) p = (p + 1) & (cap - 1);
        if (!db->ids[p]) { db->keys[p] = h; db->ids[p] = (uint32_t)(++d); }
        db->entry_id[e] = db->ids[p] - 1;
    }
    db->n_distinct = d;
    db->inv_off = (uint64_t *)calloc(d + 2, sizeof(uint64_t));
    db->inv_ref = (uint32_t *)malloc((db->n_entries + 1) * sizeof(uint32_t));
    if (!db->inv_off || !db->inv_ref) return -1;
    for (uint64_t e = 0; e < db->n_entries; e++) db->inv_off[db->entry_id[e] + 1]++;
    for (uint64_t i = 0; i < d; i++) db->inv_off[i + 1] += db->inv_off[i];
    uint64_t *cur = (uint64_t *)malloc((d + 1) * sizeof(uint64_t));
    memcpy(cur, db->inv_off, (d + 1) * sizeof(uint64_t));
    for (uint64_t r = 0; r < db->n_refs; r++)
        for (uint64_t e = db->offsets[r]; e < db->offsets[r + 1]; e++)
            db->inv_ref[cur[db->entry_id[e]]++] = (uint32_t)r;
    free(cur);
    return 0;
}

static inline int64_t db_lookup(const orc_db *db, uint64_t h)
{
    uint64_t p = mix64(h) & (db->cap - 1);
    while (db->ids[p]) {
        if (db->keys[p] == h) return (int64_t)db->ids[p] - 1;
        p = (p + 1) & (db->cap - 1);
    }
    return -1;
}

ORC_API orc_db *orc_db_from_arrays(uint32_t k, uint32_t s, uint32_t seed, uint64_t n_refs,
                                   const uint64_t *offsets, const uint64_t *hashes,
                                   const uint64_t *lengths)
{
    orc_db *db = (orc_db *)calloc(1, sizeof(orc_db));
    db->k = k; db->s = s; db->seed = seed; db->use64 = orc_use64(k);
    db->n_refs = n_refs; db->n_entries = offsets[n_refs];
    db->offsets = (uint64_t *)malloc((n_refs + 1) * sizeof(uint64_t));
    db->hashes = (uint64_t *)malloc((db->n_entries + 1) * sizeof(uint64_t));
    db->lengths = (uint64_t *)malloc((n_refs + 1) * sizeof(uint64_t));
    memcpy(db->offsets, offsets, (n_refs + 1) * sizeof(uint64_t));
    memcpy(db->hashes, hashes, db->n_entries * sizeof(uint64_t));
    memcpy(db->lengths, lengths, n_refs * sizeof(uint64_t));
    if (db_build_map(db)) return NULL;
    return db;
}

ORC_API void orc_db_free(orc_db *db)
{
    if (!db) return;
    if (db->names) for (uint64_t i = 0; i < db->n_refs; i++) free(db->names[i]);
    if (db->comments) for (uint64_t i = 0; i < db->n_refs; i++) free(db->comments[i]);
    free(db->names); free(db->comments);
    free(db->offsets); free(db->hashes); free(db->lengths);
    free(db->keys); free(db->ids); free(db->entry_id); free(db->inv_off); free(db->inv_ref);
    free(db);
}
ORC_API uint64_t orc_db_n_refs(const orc_db *db) { return db->n_refs; }
ORC_API uint64_t orc_db_n_entries(const orc_db *db) { return db->n_entries; }
ORC_API uint64_t orc_db_n_distinct(const orc_db *db) { return db->n_distinct; }
ORC_API uint32_t orc_db_k(const orc_db *db) { return db->k; }
ORC_API uint32_t orc_db_s(const orc_db *db) { return db->s; }
ORC_API uint32_t orc_db_seed(const orc_db *db) { return db->seed; }
ORC_API const char *orc_db_name(const orc_db *db, uint64_t i) { return db->names ? db->names[i] : ""; }
ORC_API const char *orc_db_comment(const orc_db *db, uint64_t i) { return db->comments ? db->comments[i] : ""; }
ORC_API uint64_t orc_db_size(const orc_db *db, uint64_t i) { return db->offsets[i + 1] - db->offsets[i]; }
ORC_API uint64_t orc_db_length(const orc_db *db, uint64_t i) { return db->lengths[i]; }
ORC_API const uint64_t *orc_db_hashes(const orc_db *db, uint64_t i) { return db->hashes + db->offsets[i]; }

/* ------------------------------------------------------------------------ */
/* .msh reader: Cap'n Proto stream framing + the MinHash schema (SURVEY       */
/* Appendix B).  Own minimal decoder -- no libcapnp in this image.             */
/* ------------------------------------------------------------------------ */
typedef struct { const uint64_t **seg; uint32_t *seg_words; uint32_t nseg; } cp_msg;
typedef struct { uint32_t seg; uint64_t word; int kind; /*0 struct,1 list,-1 null*/
                 uint32_t dwords, pwords;  /* struct shape, or composite element shape */
                 int esize; uint64_t count; } cp_obj;

static int cp_resolve(const cp_msg *m, uint32_t seg, uint64_t pword, cp_obj *o)
{
    for (int hop = 0; hop < 4; hop++) {
        if (seg >= m->nseg || pword >= m->seg_words[seg]) return -1;
        uint64_t p = m->seg[seg][pword];
        if (p == 0) { o->kind = -1; return 0; }
        int type = (int)(p & 3);
        if (type == 2) { /* far pointer */
            int dbl = (int)((p >> 2) & 1);
            uint64_t off = (p >> 3) & 0x1fffffffULL;
            uint32_t tseg = (uint32_t)(p >> 32);
            if (!dbl) { seg = tseg; pword = off; continue; }
            if (tseg >= m->nseg || off + 1 >= m->seg_words[tseg]) return -1;
            uint64_t far = m->seg[tseg][off], tag = m->seg[tseg][off + 1];
            if ((far & 3) != 2) return -1;
            uint32_t cseg = (uint32_t)(far >> 32);
            uint64_t cword = (far >> 3) & 0x1fffffffULL;
            o->seg = cseg;
            int ttype = (int)(tag & 3);
            if (ttype == 0) { o->kind = 0; o->word = cword; o->dwords = (uint32_t)((tag >> 32) & 0xffff); o->pwords = (uint32_t)(tag >> 48); return 0; }
            if (ttype == 1) { o->kind = 1; o->word = cword; o->esize = (int)((tag >> 32) & 7); o->count = tag >> 35; goto list_tag; }
            return -1;
        }
        {
            int32_t off = (int32_t)(uint32_t)(p & 0xffffffffULL) >> 2;
            uint64_t target = (uint64_t)((int64_t)pword + 1 + off);
            o->seg = seg; o->word = target;
            if (type == 0) { o->kind = 0; o->dwords = (uint32_t)((p >> 32) & 0xffff); o->pwords = (uint32_t)(p >> 48); return 0; }
            if (type == 1) { o->kind = 1; o->esize = (int)((p >> 32) & 7); o->count = p >> 35; goto list_tag; }
            return -1;
        }
    list_tag:
        if (o->esize == 7) { /* composite: tag word first */
            if (o->seg >= m->nseg || o->word >= m->seg_words[o->seg]) return -1;
            uint64_t tag = m->seg[o->seg][o->word];
            o->count = (tag >> 2) & 0x3fffffffULL;
            o->dwords = (uint32_t)((tag >> 32) & 0xffff);
            o->pwords = (uint32_t)(tag >> 48);
            o->word += 1;
        }
        return 0;
    }
    return -1;
}
static inline const uint64_t *cp_words(const cp_msg *m, uint32_t seg, uint64_t w) { return m->seg[seg] + w; }

static uint32_t cp_u32(const cp_msg *m, const cp_obj *s, uint32_t byte_off)
{
    if (byte_off + 4 > s->dwords * 8) return 0;
    uint32_t v; memcpy(&v, (const char *)cp_words(m, s->seg, s->word) + byte_off, 4); return v;
}
static uint64_t cp_u64(const cp_msg *m, const cp_obj *s, uint32_t byte_off)
{
    if (byte_off + 8 > s->dwords * 8) return 0;
    uint64_t v; memcpy(&v, (const char *)cp_words(m, s->seg, s->word) + byte_off, 8); return v;
}
static int cp_bit(const cp_msg *m, const cp_obj *s, uint32_t bit)
{
    if (bit / 8 >= s->dwords * 8) return 0;
    return (((const uint8_t *)cp_words(m, s->seg, s->word))[bit / 8] >> (bit & 7)) & 1;
}
static int cp_ptr(const cp_msg *m, const cp_obj *s, uint32_t idx, cp_obj *o)
{
    if (idx >= s->pwords) { o->kind = -1; return 0; }
    return cp_resolve(m, s->seg, s->word + s->dwords + idx, o);
}
static char *cp_text(const cp_msg *m, const cp_obj *s, uint32_t idx)
{
    cp_obj t;
    if (cp_ptr(m, s, idx, &t) || t.kind != 1 || t.esize != 2 || t.count == 0) return strdup("");
    char *r = (char *)malloc(t.count);
    memcpy(r, cp_words(m, t.seg, t.word), t.count);
    r[t.count - 1] = 0;
    return r;
}

ORC_API orc_db *orc_db_load_msh(const char *path, char *err, size_t errlen)
{
#define FAIL(msg) do { snprintf(err, errlen, "%s: %s", path, msg); goto fail; } while (0)
    FILE *f = fopen(path, "rb");
    uint64_t *buf = NULL; orc_db *db = NULL; cp_msg m; memset(&m, 0, sizeof m);
    if (!f) { snprintf(err, errlen, "could not open %s", path); return NULL; }
    fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
    buf = (uint64_t *)malloc((size_t)sz + 8);
    if (!buf || fread(buf, 1, (size_t)sz, f) != (size_t)sz) FAIL("read error");
    fclose(f); f = NULL;
    if (sz < 8) FAIL("truncated");
    const uint32_t *h32 = (const uint32_t *)buf;
    m.nseg = h32[0] + 1;
    if (m.nseg > 1u << 20 || (uint64_t)(4 + 4 * (uint64_t)m.nseg) > (uint64_t)sz) FAIL("bad segment table");
    uint64_t hdr_words = (1 + (uint64_t)m.nseg + 1) / 2; /* (4 + 4*nseg) bytes padded to 8 */
    m.seg = (const uint64_t **)malloc(m.nseg * sizeof(*m.seg));
    m.seg_words = (uint32_t *)malloc(m.nseg * sizeof(uint32_t));
    uint64_t w = hdr_words;
    for (uint32_t i = 0; i < m.nseg; i++) {
        m.seg_words[i] = h32[1 + i];
        m.seg[i] = buf + w;
        w += m.seg_words[i];
    }
    if (w * 8 > (uint64_t)sz) FAIL("segments exceed file size");

    cp_obj root;
    if (cp_resolve(&m, 0, 0, &root) || root.kind != 0) FAIL("bad root pointer");
    db = (orc_db *)calloc(1, sizeof(orc_db));
    db->k = cp_u32(&m, &root, 0);
    db->s = cp_u32(&m, &root, 8);
    db->seed = cp_u32(&m, &root, 20) ^ 42u;
    int noncanonical = cp_bit(&m, &root, 97), preserve_case = cp_bit(&m, &root, 98);
    char *alphabet = cp_text(&m, &root, 2);
    if (noncanonical || preserve_case || (alphabet[0] && strcmp(alphabet, "ACGT") != 0)) {
        free(alphabet); FAIL("unsupported sketch type (S22: noncanonical/preserveCase/non-nucleotide)");
    }
    free(alphabet);
    db->use64 = orc_use64(db->k);

    cp_obj rl, refs;
    if (cp_ptr(&m, &root, 3, &rl)) FAIL("bad referenceList pointer");
    int have = 0;
    if (rl.kind == 0 && cp_ptr(&m, &rl, 0, &refs) == 0 && refs.kind == 1 && refs.count > 0) have = 1;
    if (!have) {
        if (cp_ptr(&m, &root, 0, &rl) || rl.kind != 0) FAIL("no reference list");
        if (cp_ptr(&m, &rl, 0, &refs) || refs.kind != 1) FAIL("no references");
    }
    if (refs.esize != 7) FAIL("reference list is not a composite list");
    db->n_refs = refs.count;
    db->offsets = (uint64_t *)calloc(db->n_refs + 1, sizeof(uint64_t));
    db->lengths = (uint64_t *)calloc(db->n_refs + 1, sizeof(uint64_t));
    db->names = (char **)calloc(db->n_refs + 1, sizeof(char *));
    db->comments = (char **)calloc(db->n_refs + 1, sizeof(char *));
    const uint32_t stride = refs.dwords + refs.pwords;
    /* pass 1: sizes */
    for (uint64_t i = 0; i < db->n_refs; i++) {
        cp_obj r = refs; r.kind = 0; r.word = refs.word + i * stride;
        cp_obj hl;
        uint64_t cnt = 0;
        if (cp_ptr(&m, &r, db->use64 ? 5 : 4, &hl)) FAIL("bad hash list pointer");
        if (hl.kind == 1) cnt = hl.count;
        db->offsets[i + 1] = db->offsets[i] + cnt;
    }
    db->n_entries = db->offsets[db->n_refs];
    db->hashes = (uint64_t *)malloc((db->n_entries + 1) * sizeof(uint64_t));
    for (uint64_t i = 0; i < db->n_refs; i++) {
        cp_obj r = refs; r.kind = 0; r.word = refs.word + i * stride;
        uint64_t l64 = cp_u64(&m, &r, 8);
        db->lengths[i] = l64 ? l64 : cp_u32(&m, &r, 0);
        db->names[i] = cp_text(&m, &r, 2);
        db->comments[i] = cp_text(&m, &r, 3);
        cp_obj hl;
        cp_ptr(&m, &r, db->use64 ? 5 : 4, &hl);
        uint64_t cnt = db->offsets[i + 1] - db->offsets[i];
        if (!cnt) continue;
        if (db->use64) {
            if (hl.esize != 5) FAIL("hashes64 element size");
            memcpy(db->hashes + db->offsets[i], cp_words(&m, hl.seg, hl.word), cnt * 8);
        } else {
            if (hl.esize != 4) FAIL("hashes32 element size");
            const uint32_t *p = (const uint32_t *)cp_words(&m, hl.seg, hl.word);
            for (uint64_t j = 0; j < cnt; j++) db->hashes[db->offsets[i] + j] = p[j];
        }
    }
    free(m.seg); free(m.seg_words); free(buf);
    if (db_build_map(db)) { orc_db_free(db); snprintf(err, errlen, "out of memory"); return NULL; }
    return db;
fail:
    if (f) fclose(f);
    free(m.seg); free(m.seg_words); free(buf);
    orc_db_free(db);
    return NULL;
#undef FAIL
}

/* ------------------------------------------------------------------------ */
/* S13 identity, S14 p-value.                                                 */
/* ------------------------------------------------------------------------ */
ORC_API double orc_identity(uint64_t shared, uint64_t size, uint32_t k)
{
    if (shared == size) return 1.0;
    if (shared == 0) return 0.0;
    return pow((double)shared / (double)size, 1.0 / (double)k);
}

/* Regularized incomplete beta I_x(a,b) by the Lentz continued fraction, in
 * long double so that the oracle itself is good to ~1e-16 relative and the
 * 1e-12 parity tolerance is spent on the device code, not here.  (Mash calls
 * gsl_cdf_binomial_Q(x-1, r, n) == I_r(x, n-x+1); a Boost build calls
 * cdf(complement(binomial(n, r), x-1)) -- the same function.) */
static long double betacf_l(long double a, long double b, long double x)
{
    const long double tiny = 1e-4900L, eps = 1e-19L;
    long double qab = a + b, qap = a + 1, qam = a - 1, c = 1, d = 1 - qab * x / qap;
    if (fabsl(d) < tiny) d = tiny;
    d = 1 / d;
    long double h = d;
    for (int m = 1; m <= 100000; m++) {
        long double m2 = 2.0L * m, aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1 + aa * d; if (fabsl(d) < tiny) d = tiny;
        c = 1 + aa / c; if (fabsl(c) < tiny) c = tiny;
        d = 1 / d; h *= d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1 + aa * d; if (fabsl(d) < tiny) d = tiny;
        c = 1 + aa / c; if (fabsl(c) < tiny) c = tiny;
        d = 1 / d;
        long double del = d * c;
        h *= del;
        if (fabsl(del - 1) < eps) break;
    }
    return h;
}
static long double betai_l(long double a, long double b, long double x)
{
    if (x <= 0) return 0;
    if (x >= 1) return 1;
    long double lbt = lgammal(a + b) - lgammal(a) - lgammal(b) + a * logl(x) + b * log1pl(-x);
    if (x < (a + 1) / (a + b + 2)) return expl(lbt) * betacf_l(a, b, x) / a;
    return 1 - expl(lbt) * betacf_l(b, a, 1 - x) / b;
}

ORC_API double orc_pvalue(uint64_t x, uint64_t set_size, double kmer_space, uint64_t sketch_size)
{
    if (x == 0) return 1.0;
    double r = 1.0 / (1.0 + kmer_space / (double)set_size);
    if (x > sketch_size) return 0.0;
    return (double)betai_l((long double)x, (long double)(sketch_size - x + 1), (long double)r);
}

/* S10 */
ORC_API uint64_t orc_set_size(const uint64_t *bottom_sorted, uint64_t n, int use64)
{
    if (n == 0) return 0;
    double est = pow(2.0, use64 ? 64.0 : 32.0) * (double)n / (double)bottom_sorted[n - 1];
    if (!(est < 18446744073709551615.0)) return ~0ULL; /* top == 0 etc.: saturate (documented) */
    return (uint64_t)est;
}

/* ------------------------------------------------------------------------ */
/* Streaming pass (S6, S8, S9): threads take records from a shared cursor.    */
/* ------------------------------------------------------------------------ */
typedef struct {
    const orc_db *db; const seqset_t *ss;
    uint32_t *counts;           /* n_distinct, atomically incremented */
    uint64_t *cursor;           /* shared record cursor */
    bottom_t bottom;
    uint64_t n_kmers;
    uint32_t k, seed, s; int use64; int count_hits;
} worker_t;

static void hash_record(worker_t *w, const char *raw, uint64_t n)
{
    const uint32_t k = w->k;
    if (n < k) return; /* S6: records shorter than k are skipped */
    char *fwd = (char *)malloc(n), *rc = (char *)malloc(n);
    for (uint64_t i = 0; i < n; i++) fwd[i] = (char)toupper((unsigned char)raw[i]);
    for (uint64_t i = 0; i < n; i++) rc[i] = comp_base(fwd[n - 1 - i]);
    uint64_t run = 0; /* number of consecutive alphabet characters ending at i */
    for (uint64_t i = 0; i < n; i++) {
        run = is_acgt(fwd[i]) ? run + 1 : 0;
        if (run < k) continue;
        const uint64_t st = i + 1 - k;
        const char *f = fwd + st, *r = rc + (n - st - k);
        const char *kmer = memcmp(f, r, k) <= 0 ? f : r;
        const uint64_t h = mash_hash(kmer, k, w->seed, w->use64);
        w->n_kmers++;
        bottom_try(&w->bottom, h);
        if (w->count_hits) {
            int64_t id = db_lookup(w->db, h);
            if (id >= 0) __atomic_fetch_add(&w->counts[id], 1u, __ATOMIC_RELAXED);
        }
    }
    free(fwd); free(rc);
}

static void *worker_main(void *arg)
{
    worker_t *w = (worker_t *)arg;
    for (;;) {
        uint64_t r = __atomic_fetch_add(w->cursor, 1, __ATOMIC_RELAXED);
        if (r >= w->ss->nrec) break;
        hash_record(w, w->ss->seq + w->ss->recs[r].off, w->ss->recs[r].len);
    }
    bottom_compact(&w->bottom);
    return NULL;
}

/* Large records would serialise the pool; split them into overlapping pieces
 * (k-1 overlap) -- the multiset of k-mers is unchanged (S19). */
static void split_long_records(seqset_t *ss, uint32_t k, uint64_t piece)
{
    uint64_t n0 = ss->nrec;
    for (uint64_t r = 0; r < n0; r++) {
        uint64_t off = ss->recs[r].off, len = ss->recs[r].len;
        if (len <= piece || len < k) continue;
        ss->recs[r].len = piece;
        uint64_t pos = piece - (k - 1);
        while (pos < len) {
            uint64_t l = len - pos < piece ? len - pos : piece;
            ss_push_rec(ss, off + pos, l);
            if (pos + l >= len) break;
            pos += l - (k - 1);
        }
    }
}

static int run_stream(const orc_db *db, uint32_t k, uint32_t s, uint32_t seed, seqset_t *ss,
                      int threads, uint32_t *counts, uint64_t *bottom_out, uint64_t *n_bottom,
                      uint64_t *n_kmers)
{
    if (threads < 1) threads = 1;
    split_long_records(ss, k, 1u << 20);
    worker_t *w = (worker_t *)calloc((size_t)threads, sizeof(worker_t));
    pthread_t *th = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    uint64_t cursor = 0;
    for (int t = 0; t < threads; t++) {
        w[t].db = db; w[t].ss = ss; w[t].counts = counts; w[t].cursor = &cursor;
        w[t].k = k; w[t].s = s; w[t].seed = seed; w[t].use64 = orc_use64(k);
        w[t].count_hits = counts != NULL;
        bottom_init(&w[t].bottom, s);
        pthread_create(&th[t], NULL, worker_main, &w[t]);
    }
    bottom_t all; bottom_init(&all, s);
    uint64_t nk = 0;
    for (int t = 0; t < threads; t++) {
        pthread_join(th[t], NULL);
        nk += w[t].n_kmers;
        for (uint64_t i = 0; i < w[t].bottom.n; i++) bottom_try(&all, w[t].bottom.v[i]);
        bottom_free(&w[t].bottom);
    }
    bottom_compact(&all);
    memcpy(bottom_out, all.v, all.n * sizeof(uint64_t));
    *n_bottom = all.n; *n_kmers = nk;
    bottom_free(&all); free(w); free(th);
    return 0;
}

/* `mash sketch` semantics for one genome (Appendix B): pool all records,
 * keep the s smallest distinct canonical k-mer hashes, length = total bases. */
ORC_API int orc_sketch_text(const char *text, uint64_t n, uint32_t k, uint32_t s, uint32_t seed,
                            int threads, uint64_t *out_hashes, uint64_t *out_n, uint64_t *out_len)
{
    seqset_t ss; memset(&ss, 0, sizeof ss);
    if (parse_text(&ss, text, n)) return -1;
    uint64_t total = 0, nk;
    for (uint64_t r = 0; r < ss.nrec; r++) total += ss.recs[r].len;
    run_stream(NULL, k, s, seed, &ss, threads, NULL, out_hashes, out_n, &nk);
    *out_len = total;
    free(ss.seq); free(ss.recs);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* The screen proper (S7-S18).                                                */
/* ------------------------------------------------------------------------ */
typedef struct {
    uint64_t set_size, n_bases, n_records, n_kmers, n_mixture;
    double t_stream, t_reduce;
} orc_stats;

static double now_s(void)
{
    struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}
static int cmp_u32(const void *a, const void *b)
{
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : x > y;
}

static int screen_seqset(const orc_db *db, seqset_t *ss, int threads, int wta,
                         uint64_t *shared, uint32_t *median, double *identity, double *pvalue,
                         uint32_t *counts_out, uint64_t *mixture_out, orc_stats *st)
{
    const uint64_t N = db->n_refs, D = db->n_distinct;
    uint32_t *counts = (uint32_t *)calloc(D + 1, sizeof(uint32_t));
    uint64_t *bottom = (uint64_t *)malloc((db->s + 1) * sizeof(uint64_t));
    uint64_t nb = 0, nk = 0, total = 0;
    for (uint64_t r = 0; r < ss->nrec; r++) total += ss->recs[r].len;
    st->n_records = ss->nrec; st->n_bases = total;
    double t0 = now_s();
    run_stream(db, db->k, db->s, db->seed, ss, threads, counts, bottom, &nb, &nk);
    double t1 = now_s();
    st->n_kmers = nk; st->n_mixture = nb;
    st->set_size = orc_set_size(bottom, nb, db->use64);
    if (mixture_out) memcpy(mixture_out, bottom, nb * sizeof(uint64_t));
    if (counts_out) /* per stored entry, so callers need not know dense ids */
        for (uint64_t e = 0; e < db->n_entries; e++) counts_out[e] = counts[db->entry_id[e]];

    /* S11: shared + depths, via each reference's own hash list */
    uint32_t *depth = (uint32_t *)malloc((db->n_entries + 1) * sizeof(uint32_t));
    uint64_t *fill = (uint64_t *)calloc(N + 1, sizeof(uint64_t));
    for (uint64_t i = 0; i < N; i++) {
        uint64_t sh = 0;
        for (uint64_t e = db->offsets[i]; e < db->offsets[i + 1]; e++) {
            uint32_t c = counts[db->entry_id[e]];
            if (c) depth[db->offsets[i] + sh++] = c;
        }
        shared[i] = sh;
    }
    if (wta) { /* S17 */
        double *score = (double *)malloc((N + 1) * sizeof(double));
        for (uint64_t i = 0; i < N; i++) {
            score[i] = orc_identity(shared[i], db->offsets[i + 1] - db->offsets[i], db->k);
            shared[i] = 0;
        }
        for (uint64_t d = 0; d < D; d++) {
            if (!counts[d]) continue;
            uint64_t best = ~0ULL;
            for (uint64_t p = db->inv_off[d]; p < db->inv_off[d + 1]; p++) {
                uint64_t r = db->inv_ref[p];
                if (best == ~0ULL || score[r] > score[best] ||
                    (score[r] == score[best] && db->lengths[r] >= db->lengths[best]))
                    best = r; /* full tie: highest reference index wins (documented rule) */
            }
            depth[db->offsets[best] + fill[best]++] = counts[d];
        }
        for (uint64_t i = 0; i < N; i++) shared[i] = fill[i];
        free(score);
    }
    /* S12-S14 */
    const double kmer_space = pow(4.0, (double)db->k);
    for (uint64_t i = 0; i < N; i++) {
        uint64_t sz = db->offsets[i + 1] - db->offsets[i];
        qsort(depth + db->offsets[i], shared[i], sizeof(uint32_t), cmp_u32);
        median[i] = shared[i] ? depth[db->offsets[i] + shared[i] / 2] : 0;
        identity[i] = orc_identity(shared[i], sz, db->k);
        pvalue[i] = orc_pvalue(shared[i], st->set_size, kmer_space, sz);
    }
    st->t_stream = t1 - t0; st->t_reduce = now_s() - t1;
    free(depth); free(fill); free(counts); free(bottom);
    return 0;
}

/* Screen FASTA/FASTQ text held in memory (all inputs pooled into one mixture). */
ORC_API int orc_screen_text(const orc_db *db, const char *text, uint64_t n, int threads, int wta,
                            uint64_t *shared, uint32_t *median, double *identity, double *pvalue,
                            uint32_t *counts_per_entry, uint64_t *mixture, orc_stats *st)
{
    seqset_t ss; memset(&ss, 0, sizeof ss);
    if (parse_text(&ss, text, n)) return -1;
    int rc = screen_seqset(db, &ss, threads, wta, shared, median, identity, pvalue,
                           counts_per_entry, mixture, st);
    free(ss.seq); free(ss.recs);
    return rc;
}

ORC_API int orc_screen_files(const orc_db *db, const char **paths, int n_paths, int threads, int wta,
                             uint64_t *shared, uint32_t *median, double *identity, double *pvalue,
                             orc_stats *st, char *err, size_t errlen)
{
    seqset_t ss; memset(&ss, 0, sizeof ss);
    for (int i = 0; i < n_paths; i++) {
        char *buf; uint64_t n;
        if (slurp_gz(paths[i], &buf, &n)) { snprintf(err, errlen, "could not open %s", paths[i]); free(ss.seq); free(ss.recs); return -2; }
        int rc = parse_text(&ss, buf, n);
        free(buf);
        if (rc) return -1;
    }
    if (ss.nrec == 0) { snprintf(err, errlen, "Did not find sequence records in inputs"); free(ss.seq); free(ss.recs); return -3; }
    int rc = screen_seqset(db, &ss, threads, wta, shared, median, identity, pvalue, NULL, NULL, st);
    free(ss.seq); free(ss.recs);
    return rc;
}

/* ------------------------------------------------------------------------ */
/* CLI:  oracle_mash screen [-p N] [-w] [-i I] [-v P] db.msh in.fa [in.fa..] */
/* (S15, S16, S20, S21).  Used as the timed CPU baseline.                     */
/* ------------------------------------------------------------------------ */
#ifdef ORC_MAIN
int main(int argc, char **argv)
{
    if (argc < 2 || strcmp(argv[1], "screen") != 0) {
        fprintf(stderr, "usage: oracle_mash screen [-p N] [-w] [-i I] [-v P] <queries>.msh <mixture> [<mixture>] ...\n");
        return argc < 2 ? 0 : 2;
    }
    int threads = 1, wta = 0; double imin = 0.0, pmax = 1.0;
    int a = 2;
    for (; a < argc && argv[a][0] == '-' && argv[a][1]; a++) {
        if (!strcmp(argv[a], "-w")) wta = 1;
        else if (!strcmp(argv[a], "-h")) { a = argc; break; }
        else if (a + 1 < argc && !strcmp(argv[a], "-p")) threads = atoi(argv[++a]);
        else if (a + 1 < argc && !strcmp(argv[a], "-i")) imin = atof(argv[++a]);
        else if (a + 1 < argc && !strcmp(argv[a], "-v")) pmax = atof(argv[++a]);
        else { fprintf(stderr, "ERROR: unknown option %s\n", argv[a]); return 1; }
    }
    if (argc - a < 2) {
        fprintf(stderr, "usage: oracle_mash screen [options] <queries>.msh <mixture> [<mixture>] ...\n");
        return 0;
    }
    const char *dbp = argv[a];
    size_t L = strlen(dbp);
    if (L < 4 || strcmp(dbp + L - 4, ".msh") != 0) {
        fprintf(stderr, "ERROR: %s does not look like a sketch (.msh)\n", dbp);
        return 1;
    }
    char err[512];
    double t0 = now_s();
    fprintf(stderr, "Loading %s...\n", dbp);
    orc_db *db = orc_db_load_msh(dbp, err, sizeof err);
    if (!db) { fprintf(stderr, "ERROR: %s\n", err); return 1; }
    fprintf(stderr, "   %llu distinct hashes.\n", (unsigned long long)db->n_distinct);
    double t1 = now_s();
    uint64_t N = db->n_refs;
    uint64_t *shared = (uint64_t *)calloc(N + 1, 8);
    uint32_t *median = (uint32_t *)calloc(N + 1, 4);
    double *identity = (double *)calloc(N + 1, 8), *pv = (double *)calloc(N + 1, 8);
    orc_stats st; memset(&st, 0, sizeof st);
    fprintf(stderr, "Streaming from %d inputs...\n", argc - a - 1);
    int rc = orc_screen_files(db, (const char **)(argv + a + 1), argc - a - 1, threads, wta,
                              shared, median, identity, pv, &st, err, sizeof err);
    if (rc) { fprintf(stderr, "ERROR: %s\n", err); return 1; }
    fprintf(stderr, "   Estimated distinct k-mers in mixture: %llu\n", (unsigned long long)st.set_size);
    fprintf(stderr, "Writing output...\n");
    for (uint64_t i = 0; i < N; i++) {
        if (!(shared[i] != 0 || imin < 0.0)) continue;
        if (identity[i] < imin) continue;
        if (pv[i] > pmax) continue;
        printf("%g\t%llu/%llu\t%u\t%g\t%s\t%s\n", identity[i], (unsigned long long)shared[i],
               (unsigned long long)orc_db_size(db, i), median[i], pv[i], orc_db_name(db, i),
               orc_db_comment(db, i));
    }
    fprintf(stderr, "[oracle timing] load %.3f s, stream %.3f s (%llu bases, %.1f Mbp/s, %d threads), reduce %.3f s\n",
            t1 - t0, st.t_stream, (unsigned long long)st.n_bases,
            st.t_stream > 0 ? 1e-6 * (double)st.n_bases / st.t_stream : 0.0, threads, st.t_reduce);
    orc_db_free(db);
    return 0;
}
#endif
