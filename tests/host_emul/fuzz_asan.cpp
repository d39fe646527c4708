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
        for (int t = 0; t < rounds; t++) {
            const size_t n = rnd() % 700;
            char *txt = (char *)malloc(n ? n : 1);
            const int mode = t % 3;
            for (size_t i = 0; i < n; i++)
                txt[i] = mode == 0 ? alpha[rnd() % (sizeof alpha - 1)]
                       : mode == 1 ? (char)(rnd() & 0xFF) : "ACGT\n>"[rnd() % (rnd() % 40 == 0 ? 6 : 4)];
            const size_t cap = hs::pack_words_bound(n);
            uint64_t *seq = (uint64_t *)malloc(cap * 8 + 8);
            uint32_t *inv = (uint32_t *)malloc(cap * 4 + 4);
            hs::PackStats st;
            const uint64_t w = hs::pack_text_span(txt, n, seq, inv, &st);
            if (w > cap) { printf("packed %llu words into a bound of %zu\n", (unsigned long long)w, cap); return 1; }
            total += w;
            auto sp = hs::split_records(txt, n, 1 + (int)(rnd() % 5), 1 + rnd() % 64);
            size_t pos = 0;
            for (auto &p : sp) { if (p.first != pos || p.second < p.first) { printf("bad split\n"); return 2; } pos = p.second; }
            if (n && pos != n) { printf("split does not cover the text\n"); return 3; }
            free(txt); free(seq); free(inv);
        }
        printf("ok %llu\n", total);
        return 0;
    }
    return 64;
}
