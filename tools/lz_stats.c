/* lz_stats.c -- development aid: what the exhaustive greedy parse of src/LZ77.ts asks of an index.
 * For 64 KiB chunks of each synthetic kind it walks the reference's parse (App. A.1) and, at every parse position,
 * counts the candidates an index keyed by 3 / 4 / 5 bytes would have to look at. Host only.
 *   gcc -O2 -o /tmp/lz_stats tools/lz_stats.c zlib.ts_b200/csrc/zts_synth.c && /tmp/lz_stats
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

void zts_gen_text(uint8_t* out, size_t n, uint32_t seed);
void zts_gen_mixed(uint8_t* out, size_t n, uint32_t seed, uint32_t seg);

#define CH 65536
#define NBIN 10
static const int edges[NBIN] = {0, 1, 2, 4, 8, 16, 32, 64, 128, 1 << 30};
static int bin_of(int c)
{
    for (int i = 0; i < NBIN; ++i)
        if (c <= edges[i]) return i;
    return NBIN - 1;
}

typedef struct {
    long tokens, lits, matches, mlen_sum, len3;
    long h3[NBIN], h4[NBIN], h5[NBIN];    /* in-window earlier candidates sharing 3 / 4 / 5 bytes, at match positions */
    long c3_sum, c4_sum, c5_sum;
    long steps3, steps4, steps5;          /* sum of ceil(c / 32) */
    long no4_has3;                        /* positions with a 3-byte candidate but no 4-byte one */
    long bytes;
} stats;

static void chunk_stats(const uint8_t* s, int n, stats* st)
{
    int p = 0;
    st->bytes += n;
    while (p < n) {
        if (p + 3 >= n) {
            st->lits += n - p;
            st->tokens += n - p;
            break;
        }
        int best = 0, bq = -1, c3 = 0, c4 = 0, c5 = 0;
        int maxlen = n - p < 258 ? n - p : 258;
        int lo = p > 32768 ? p - 32768 : 0;
        for (int q = p - 1; q >= lo; --q) {
            if (s[q] != s[p] || s[q + 1] != s[p + 1] || s[q + 2] != s[p + 2]) continue;
            ++c3;
            int l = 3;
            while (l < maxlen && s[q + l] == s[p + l]) ++l;
            if (l >= 4) ++c4;
            if (l >= 5) ++c5;
            if (l > best) {
                best = l;
                bq = q;
            }
        }
        (void)bq;
        if (best >= 3) {
            st->matches++;
            st->tokens++;
            st->mlen_sum += best;
            if (best == 3) st->len3++;
            st->h3[bin_of(c3)]++;
            st->h4[bin_of(c4)]++;
            st->h5[bin_of(c5)]++;
            st->c3_sum += c3;
            st->c4_sum += c4;
            st->c5_sum += c5;
            st->steps3 += (c3 + 31) / 32;
            st->steps4 += (c4 + 31) / 32;
            st->steps5 += (c5 + 31) / 32;
            if (c4 == 0) st->no4_has3++;
            p += best;
        } else {
            st->lits++;
            st->tokens++;
            p++;
        }
    }
}

static void report(const char* name, const stats* st)
{
    printf("%-8s bytes %ld tokens %ld (%.2f B/tok) lits %ld matches %ld avg len %.2f len3 %.1f%% no4has3 %.1f%%\n", name,
           st->bytes, st->tokens, (double)st->bytes / st->tokens, st->lits, st->matches,
           (double)st->mlen_sum / (st->matches ? st->matches : 1), 100.0 * st->len3 / (st->matches ? st->matches : 1),
           100.0 * st->no4_has3 / (st->matches ? st->matches : 1));
    printf("   cand/match: c3 %.1f c4 %.1f c5 %.1f   32-steps/match: %.2f %.2f %.2f\n",
           (double)st->c3_sum / st->matches, (double)st->c4_sum / st->matches, (double)st->c5_sum / st->matches,
           (double)st->steps3 / st->matches, (double)st->steps4 / st->matches, (double)st->steps5 / st->matches);
    printf("   bins <=:   ");
    for (int i = 0; i < NBIN; ++i) printf("%7d", edges[i] > 100000 ? 99999 : edges[i]);
    printf("\n   c3 %%:      ");
    for (int i = 0; i < NBIN; ++i) printf("%7.1f", 100.0 * st->h3[i] / st->matches);
    printf("\n   c4 %%:      ");
    for (int i = 0; i < NBIN; ++i) printf("%7.1f", 100.0 * st->h4[i] / st->matches);
    printf("\n   c5 %%:      ");
    for (int i = 0; i < NBIN; ++i) printf("%7.1f", 100.0 * st->h5[i] / st->matches);
    printf("\n");
}

int main(int argc, char** argv)
{
    int nchunks = argc > 1 ? atoi(argv[1]) : 16;
    size_t n = (size_t)nchunks * CH;
    uint8_t* buf = malloc(n);
    stats st;
    /* text */
    zts_gen_text(buf, n, 1);
    memset(&st, 0, sizeof st);
    for (int c = 0; c < nchunks; ++c) chunk_stats(buf + (size_t)c * CH, CH, &st);
    report("text", &st);
    /* records */
    {
        uint32_t x = 7;
        for (size_t r = 0; r < n / 8; ++r) {
            uint32_t v = 7u * (uint32_t)r;
            x = x * 1664525u + 1013904223u;
            uint8_t rec[8] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24), (uint8_t)((x >> 16) & 15), 0, 0, 0};
            memcpy(buf + r * 8, rec, 8);
        }
    }
    memset(&st, 0, sizeof st);
    for (int c = 0; c < nchunks; ++c) chunk_stats(buf + (size_t)c * CH, CH, &st);
    report("records", &st);
    /* mixed */
    zts_gen_mixed(buf, n, 2, 4096);
    memset(&st, 0, sizeof st);
    for (int c = 0; c < nchunks; ++c) chunk_stats(buf + (size_t)c * CH, CH, &st);
    report("mixed", &st);
    free(buf);
    return 0;
}
