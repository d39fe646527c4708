// tests/host_emul/fuzz_asan.cpp -- TEST INFRASTRUCTURE.  Built with -fsanitize=address,undefined by
// tests/test_abi_cpu.py: the two host-side parsers of untrusted input (the .msh reader and the FASTA
// packer / record splitter) must survive corrupted and random input without touching memory they do
// not own.  Exact-size heap buffers make every overrun visible to the sanitizer.
//   fuzz_asan msh file...     parse each file, print "<ok> <rejected>"
//   fuzz_asan pack <rounds>   random text through pack_text_span / split_records
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../hymet_b200/csrc/fasta_pack.h"
#include "../../hymet_b200/csrc/msh_capnp.h"

int main(int argc, char **argv)
{
    if (argc >= 2 && !strcmp(argv[1], "msh")) {
        int ok = 0, bad = 0;
        for (int i = 2; i < argc; i++) {
            hs::MshData d;
            std::string err;
            if (hs::msh_read(argv[i], d, err)) bad++; else ok++;
        }
        printf("%d %d\n", ok, bad);
        return 0;
    }
    if (argc >= 3 && !strcmp(argv[1], "pack")) {
        unsigned long long seed = 12345;
        auto rnd = [&]() { seed = seed * 6364136223846793005ull + 1442695040888963407ull; return (unsigned)(seed >> 33); };
        const char alpha[] = "ACGTacgtNn>@+\n\r \tXYZ-*";
        const int rounds = atoi(argv[2]);
        unsigned long long total = 0;
        int levels_seen = 0;
        for (int t = 0; t < rounds; t++) {
            const int mode = t % 4;
            const size_t n = rnd() % (mode == 3 ? 12000 : 700);
            char *txt = (char *)malloc(n ? n : 1);
            for (size_t i = 0; i < n; i++)
                txt[i] = mode == 0 ? alpha[rnd() % (sizeof alpha - 1)]
                       : mode == 1 ? (char)(rnd() & 0xFF) : "ACGT\n>"[rnd() % (rnd() % 40 == 0 ? 6 : 4)];
            if (mode == 3) {
                // wrapped records: long runs of plain sequence (the 64-byte block path) broken by headers,
                // '+' lines, CRLF, lower case, N and the odd stray symbol
                const size_t width = 1 + rnd() % 120;
                size_t col = 0;
                bool crlf = rnd() % 4 == 0;
                for (size_t i = 0; i < n; i++) {
                    char c = "ACGTacgtN"[rnd() % (rnd() % 16 == 0 ? 9 : 4)];
                    if (rnd() % 500 == 0) c = alpha[rnd() % (sizeof alpha - 1)];
                    if (col == 0 && (i == 0 || rnd() % 60 == 0)) c = "+>@"[rnd() % 8 == 0 ? rnd() % 3 : 1];
                    if (col >= width) { c = '\n'; if (crlf && i && txt[i - 1] != '\r' && rnd() % 2) c = '\r'; }
                    txt[i] = c;
                    col = c == '\n' ? 0 : col + 1;
                }
            }
            const size_t cap = hs::pack_words_bound(n);
            uint64_t *seq = (uint64_t *)malloc(cap * 8 + 8);
            uint32_t *inv = (uint32_t *)malloc(cap * 4 + 4);
            hs::PackStats st;
            hs::set_pack_level(0);   // scalar: the statement of the format
            const uint64_t w = hs::pack_text_span(txt, n, seq, inv, &st);
            if (w > cap) { printf("packed %llu words into a bound of %zu\n", (unsigned long long)w, cap); return 1; }
            for (int level = 1; level <= 2; level++) {   // every SIMD level the host has must agree bit for bit
                hs::set_pack_level(level);
                if (hs::pack_level() != level) continue;
                levels_seen |= 1 << level;
                uint64_t *seq2 = (uint64_t *)malloc(cap * 8 + 8);
                uint32_t *inv2 = (uint32_t *)malloc(cap * 4 + 4);
                hs::PackStats st2;
                const uint64_t w2 = hs::pack_text_span(txt, n, seq2, inv2, &st2);
                if (w2 != w || memcmp(seq, seq2, w * 8) || memcmp(inv, inv2, w * 4) || st2.n_records != st.n_records ||
                    st2.n_seq_bases != st.n_seq_bases || st2.n_positions != st.n_positions) {
                    printf("pack level %d differs from the scalar packer (round %d, mode %d, %zu bytes)\n", level, t, mode, n);
                    return 4;
                }
                free(seq2); free(inv2);
            }
            hs::set_pack_level(-1);
            total += w;
            auto sp = hs::split_records(txt, n, 1 + (int)(rnd() % 5), 1 + rnd() % 64);
            size_t pos = 0;
            for (auto &p : sp) { if (p.first != pos || p.second < p.first) { printf("bad split\n"); return 2; } pos = p.second; }
            if (n && pos != n) { printf("split does not cover the text\n"); return 3; }
            // Cutting in front of a header the packer sees changes nothing: same records, bases, positions.  Holds
            // for FASTQ-first text (cut by the packer's own walk) and for FASTA without '@' records; a FASTA-first
            // file with FASTQ records inside is cut at "\n>" without looking at quality blocks (not a format).
            size_t first = 0;
            while (first < n && (txt[first] == '\n' || txt[first] == '\r')) first++;
            if (first < n && (txt[first] == '@' || !memchr(txt, '@', n))) {
                hs::PackStats parts_st;
                for (auto &p : sp) hs::pack_text_span(txt + p.first, p.second - p.first, seq, inv, &parts_st);
                if (parts_st.n_records != st.n_records || parts_st.n_seq_bases != st.n_seq_bases || parts_st.n_positions != st.n_positions) {
                    printf("packing the %zu spans of split_records differs from packing the text (round %d, mode %d, %zu bytes): "
                           "records %llu/%llu bases %llu/%llu\n", sp.size(), t, mode, n, (unsigned long long)parts_st.n_records,
                           (unsigned long long)st.n_records, (unsigned long long)parts_st.n_seq_bases, (unsigned long long)st.n_seq_bases);
                    return 5;
                }
            }
            free(txt); free(seq); free(inv);
        }
        printf("ok %llu levels %d\n", total, levels_seen);
        return 0;
    }
    return 64;
}
