/*
 * zts_synth.c -- portable synthetic input generators.
 * Integer-only LCG text / mixed-entropy buffers used by tests and bench.py
 * (BASELINE.json configs C1..C5). Host code, no CUDA.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint32_t x;
} lcg;

static inline uint32_t u16(lcg* g)
{
    g->x = g->x * 1664525u + 1013904223u;
    return g->x >> 16;
}

/* text(n, seed): Zipf-ish words over "etaoinshrdlcumwfgypbvkjxqz" */
void zts_gen_text(uint8_t* out, size_t n, uint32_t seed)
{
    static const char A[] = "etaoinshrdlcumwfgypbvkjxqz";
    lcg g = {seed};
    uint8_t(*words)[12] = malloc(4096 * 12);
    uint8_t* wlen = malloc(4096);
    for (int w = 0; w < 4096; ++w) {
        int len = 2 + (int)(u16(&g) % 9);
        wlen[w] = (uint8_t)len;
        for (int i = 0; i < len; ++i) {
            uint32_t a = u16(&g), b = u16(&g);
            words[w][i] = (uint8_t)A[(((a * b) >> 16) * 26u) >> 16];
        }
    }
    size_t pos = 0;
    while (pos < n) {
        uint32_t a = u16(&g), b = u16(&g), c = u16(&g);
        uint32_t wi = (((a * b) >> 16) * c) >> 20;
        for (int i = 0; i < wlen[wi] && pos < n; ++i) out[pos++] = words[wi][i];
        uint32_t r = u16(&g) % 64;
        const char* sep = r == 0 ? ". " : r == 1 ? ",\n" : " ";
        for (int i = 0; sep[i] && pos < n; ++i) out[pos++] = (uint8_t)sep[i];
    }
    free(words);
    free(wlen);
}

/* mixed(n, seed, seg): uniform mix of random / text / byte-run / 8-byte-record segments */
void zts_gen_mixed(uint8_t* out, size_t n, uint32_t seed, uint32_t seg)
{
    lcg g = {seed};
    size_t tn = n < ((size_t)1 << 24) ? n : ((size_t)1 << 24);
    uint8_t* t = malloc(tn ? tn : 1);
    zts_gen_text(t, tn, seed ^ 0x5bd1e995u);
    size_t tp = 0, pos = 0;
    while (pos < n) {
        uint32_t k = u16(&g) & 3;
        if (k == 0) {
            for (uint32_t i = 0; i < seg; ++i) {
                uint8_t b = (uint8_t)(u16(&g) & 0xFF);
                if (pos < n) out[pos++] = b;
            }
        } else if (k == 1) {
            for (uint32_t i = 0; i < seg && tp + i < tn; ++i)
                if (pos < n) out[pos++] = t[tp + i];
            size_t m = tn > seg ? tn - seg : 1;
            tp = (tp + seg) % m;
        } else if (k == 2) {
            uint8_t b = (uint8_t)(u16(&g) & 0xFF);
            for (uint32_t i = 0; i < seg; ++i)
                if (pos < n) out[pos++] = b;
        } else {
            for (uint32_t r = 0; r < seg / 8; ++r) {
                uint32_t v = 7u * (uint32_t)(pos / 8);
                uint8_t rec[8] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24),
                                  (uint8_t)(u16(&g) & 15), 0, 0, 0};
                for (int i = 0; i < 8; ++i)
                    if (pos < n) out[pos++] = rec[i];
            }
        }
    }
    free(t);
}
