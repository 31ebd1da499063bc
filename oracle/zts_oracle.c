/*
 * zts_oracle.c -- CPU restatement of the zlib.ts hot path. TEST INFRASTRUCTURE ONLY
 * (see zts_oracle.h: pinned to the executed reference through oracle/minijs, and who may load it).
 *
 * Every function cites the reference lines it restates (paths under /root/reference).
 * JavaScript semantics that matter are made explicit: Uint16Array / Uint8Array wrap-around,
 * out-of-range typed-array writes being dropped, `undefined + x = NaN`, NaN comparisons false.
 */
#include "zts_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * CRC-32  (src/CRC32.ts:25-69)  reflected 0xEDB88320, init/xorout 0xFFFFFFFF
 * ---------------------------------------------------------------------------------------- */
static uint32_t g_crc_table[256];
static int g_crc_ready = 0;

static void crc_init(void) /* src/CRC32.ts:61-69 */
{
    for (uint32_t i = 0; i < 256; ++i) {
        uint32_t c = i;
        for (int j = 0; j < 8; ++j) c = (c & 1) ? (0xEDB88320u ^ (c >> 1)) : (c >> 1);
        g_crc_table[i] = c;
    }
    g_crc_ready = 1;
}

uint32_t zo_crc32_update(const uint8_t* data, size_t len, uint32_t crc)
{
    if (!g_crc_ready) crc_init();
    crc ^= 0xFFFFFFFFu;
    /* the reference unrolls by 8 (il & 7 first, then il >> 3 groups); byte order is unchanged */
    for (size_t i = 0; i < len; ++i) crc = (crc >> 8) ^ g_crc_table[(crc ^ data[i]) & 0xFF];
    return crc ^ 0xFFFFFFFFu;
}

uint32_t zo_crc32(const uint8_t* data, size_t len) { return zo_crc32_update(data, len, 0); }

/* ------------------------------------------------------------------------------------------
 * Adler-32  (src/Adler32.ts:28-48)  modulo deferred every OptimizationParameter=1024 bytes
 * ---------------------------------------------------------------------------------------- */
uint32_t zo_adler32_update(uint32_t adler, const uint8_t* data, size_t len)
{
    uint32_t s1 = adler & 0xFFFF, s2 = (adler >> 16) & 0xFFFF;
    size_t pos = 0;
    while (len > 0) {
        size_t tlen = len > 1024 ? 1024 : len;
        len -= tlen;
        do {
            s1 += data[pos++];
            s2 += s1;
        } while (--tlen);
        s1 %= 65521;
        s2 %= 65521;
    }
    return (s2 << 16) | s1;
}

uint32_t zo_adler32(const uint8_t* data, size_t len) { return zo_adler32_update(1, data, len); }

/* ------------------------------------------------------------------------------------------
 * LZ77  (src/LZ77.ts)
 * ---------------------------------------------------------------------------------------- */
#define LZ_MIN 3
#define LZ_MAX 258
#define LZ_WINDOW 0x8000

static const uint16_t kLenBase[29] = {3,   4,   5,   6,   7,   8,   9,   10,  11, 13,
                                      15,  17,  19,  23,  27,  31,  35,  43,  51, 59,
                                      67,  83,  99,  115, 131, 163, 195, 227, 258};
static const uint8_t kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2,
                                      2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
static const uint16_t kDistBase[30] = {1,    2,    3,    4,    5,    7,    9,    13,    17,    25,
                                       33,   49,   65,   97,   129,  193,  257,  385,   513,   769,
                                       1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
static const uint8_t kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2,  3,  3,  4,  4,  5,  5,  6,
                                       6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

/* src/LZ77.ts:20-53: [code, extra value, extra bits]; 227..257 -> 284, 258 -> 285 */
static void length_code(int length, uint16_t out[3])
{
    int idx = 28;
    if (length != 258) {
        idx = 27;
        while (kLenBase[idx] > length) --idx;
    }
    out[0] = (uint16_t)(257 + idx);
    out[1] = (uint16_t)(length - kLenBase[idx]);
    out[2] = kLenExtra[idx];
}

/* src/LZ77.ts:56-90 */
static void distance_code(int dist, uint16_t out[3])
{
    int idx = 29;
    while (kDistBase[idx] > dist) --idx;
    out[0] = (uint16_t)idx;
    out[1] = (uint16_t)(dist - kDistBase[idx]);
    out[2] = kDistExtra[idx];
}

typedef struct {
    int len, dist, valid;
} lz_match;

typedef struct {
    const uint8_t* in;
    size_t n;
    uint16_t* out; /* Uint16Array(2n): writes at pos >= cap are dropped (JS typed array) */
    size_t cap, pos;
    uint32_t* fl;
    uint32_t* fd;
    long skip;
    lz_match prev;
    /* exact-key candidate lists: hash heads + per-position links, filtered by the 24-bit key.
     * Equivalent to the reference's table[matchKey] arrays walked newest-first (:163-164) with
     * the window pruning of :223-225 applied as a distance test (lists are position-ordered). */
    int32_t* head;
    int32_t* link;
} lz_state;

#define LZ_HASH_BITS 16
static inline uint32_t lz_key(const uint8_t* p) { return ((uint32_t)p[0] << 16) | ((uint32_t)p[1] << 8) | p[2]; }
static inline uint32_t lz_hash(uint32_t key) { return (key * 2654435761u) >> (32 - LZ_HASH_BITS); }

static void lz_write_num(lz_state* s, uint16_t v) /* :131 */
{
    if (s->pos < s->cap) s->out[s->pos] = v;
    s->pos++;
}

static void lz_write_match(lz_state* s, lz_match m, int offset) /* :135-146 */
{
    uint16_t arr[6];
    length_code(m.len, arr);
    distance_code(m.dist, arr + 3);
    for (int i = 0; i < 6; ++i) lz_write_num(s, arr[i]);
    s->fl[arr[0]]++;
    s->fd[arr[3]]++;
    s->skip = (long)m.len + offset - 1;
    s->prev.valid = 0;
}

static void lz_insert(lz_state* s, size_t position) /* matchList.push(position) :218,:274 */
{
    /* positions whose key would be shorter than 3 bytes (:206) are never searched nor found */
    if (position + LZ_MIN > s->n) return;
    uint32_t h = lz_hash(lz_key(s->in + position));
    s->link[position] = s->head[h];
    s->head[h] = (int32_t)position;
}

/* :149-154 */
static int lz_max_match_test(const uint8_t* in, size_t m1, size_t m2, int len)
{
    for (int j = len; j > LZ_MIN; --j)
        if (in[m1 + j - 1] != in[m2 + j - 1]) return 0;
    return 1;
}

/* :157-194, candidates newest -> oldest; returns valid=0 when the (pruned) list is empty (:242) */
static lz_match lz_search(lz_state* s, size_t position)
{
    const uint8_t* in = s->in;
    const size_t n = s->n;
    const uint32_t key = lz_key(in + position);
    lz_match r = {0, 0, 0};
    size_t current = 0;
    int match_max = 0;
    for (int32_t q = s->head[lz_hash(key)]; q >= 0; q = s->link[q]) {
        if (position - (size_t)q > LZ_WINDOW) break; /* :223 */
        if (lz_key(in + q) != key) continue;         /* different table[] entry */
        if (!r.valid) {
            r.valid = 1;
            current = (size_t)q; /* :158 default */
        }
        size_t match = (size_t)q;
        int match_length = LZ_MIN;
        if (match_max > LZ_MIN) {
            if (!lz_max_match_test(in, match, position, match_max)) continue;
            match_length = match_max;
        }
        while (match_length < LZ_MAX && position + match_length < n &&
               in[match + match_length] == in[position + match_length])
            match_length++;
        if (match_length > match_max) {
            current = match;
            match_max = match_length;
        }
        if (match_length == LZ_MAX) break;
    }
    r.len = match_max;
    r.dist = (int)(position - current);
    return r;
}

/* LZ77.encode over in[0, n) whose first dict_len bytes are history only: they enter the match table exactly as
 * positions inside a match do (:217-220, the skipLength path) and produce no output -- the state the reference's
 * loop is in at position dict_len after a parse that happened to end there. dict_len = 0 is LZ77.encode itself. */
int zo_lz77_encode_dict(const uint8_t* in, size_t n, size_t dict_len, int lazy, uint16_t* tokens, size_t* ntok,
                        uint32_t fl[286], uint32_t fd[30])
{
    lz_state s;
    memset(&s, 0, sizeof s);
    if (dict_len > n || dict_len > 0x7FFFFFFF) return ZO_E_INPUT_BROKEN;
    s.skip = (long)dict_len;
    s.in = in;
    s.n = n;
    s.out = tokens;
    s.cap = 2 * n;
    s.fl = fl;
    s.fd = fd;
    memset(fl, 0, 286 * sizeof(uint32_t));
    memset(fd, 0, 30 * sizeof(uint32_t));
    fl[256] = 1; /* :127 */
    s.head = (int32_t*)malloc(sizeof(int32_t) << LZ_HASH_BITS);
    s.link = (int32_t*)malloc(sizeof(int32_t) * (n ? n : 1));
    if (!s.head || !s.link) {
        free(s.head);
        free(s.link);
        return ZO_E_NOMEM;
    }
    memset(s.head, 0xFF, sizeof(int32_t) << LZ_HASH_BITS);

    for (size_t position = 0; position < n; ++position) { /* :202 */
        if ((s.skip--) > 0) {                             /* :217-220 */
            lz_insert(&s, position);
            continue;
        }
        if (position + LZ_MIN >= n) { /* :228-239 */
            if (s.prev.valid) lz_write_match(&s, s.prev, -1);
            for (size_t i = position; i < n; ++i) {
                lz_write_num(&s, in[i]);
                fl[in[i]]++;
            }
            break;
        }
        lz_match longest = lz_search(&s, position);
        if (longest.valid) { /* :242-263 */
            if (s.prev.valid) {
                if (s.prev.len < longest.len) {
                    uint8_t tmp = in[position - 1];
                    lz_write_num(&s, tmp);
                    fl[tmp]++;
                    lz_write_match(&s, longest, 0);
                } else {
                    lz_write_match(&s, s.prev, -1);
                }
            } else if (longest.len < lazy) {
                s.prev = longest;
            } else {
                lz_write_match(&s, longest, 0);
            }
        } else if (s.prev.valid) { /* :264-266 */
            lz_write_match(&s, s.prev, -1);
        } else { /* :267-272 */
            lz_write_num(&s, in[position]);
            fl[in[position]]++;
        }
        lz_insert(&s, position); /* :274 */
    }
    lz_write_num(&s, 256); /* :278-279 */
    fl[256]++;
    *ntok = s.pos < s.cap ? s.pos : s.cap; /* subarray(0,pos) of a 2n-long array :281 */
    free(s.head);
    free(s.link);
    return ZO_OK;
}

int zo_lz77_encode(const uint8_t* in, size_t n, int lazy, uint16_t* tokens, size_t* ntok,
                   uint32_t fl[286], uint32_t fd[30])
{
    return zo_lz77_encode_dict(in, n, 0, lazy, tokens, ntok, fl, fd);
}

/* ------------------------------------------------------------------------------------------
 * Heap  (src/Heap.ts) -- max-heap of (value,index) pairs in one Uint16Array
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint16_t buf[4 * 286]; /* new Heap(2*HUFMAX) -> Uint16Array(size*2)  src/RawDeflate.ts:442 */
    int length;
} zo_heap;

static void heap_push(zo_heap* h, uint16_t index, uint16_t value) /* :49-81 */
{
    int current = h->length, parent;
    h->buf[h->length++] = value;
    h->buf[h->length++] = index;
    while (current > 0) {
        parent = ((current - 2) >> 2) << 1; /* :31 */
        if (h->buf[current] > h->buf[parent]) {
            uint16_t t = h->buf[current];
            h->buf[current] = h->buf[parent];
            h->buf[parent] = t;
            t = h->buf[current + 1];
            h->buf[current + 1] = h->buf[parent + 1];
            h->buf[parent + 1] = t;
            current = parent;
        } else {
            break;
        }
    }
}

static void heap_pop(zo_heap* h, uint16_t* index, uint16_t* value) /* :88-132 */
{
    uint16_t* heap = h->buf;
    *value = heap[0];
    *index = heap[1];
    h->length -= 2;
    heap[0] = heap[h->length];
    heap[1] = heap[h->length + 1];
    int parent = 0;
    for (;;) {
        int current = 2 * parent + 2; /* :40 */
        if (current >= h->length) break;
        if (current + 2 < h->length && heap[current + 2] > heap[current]) current += 2;
        if (heap[current] > heap[parent]) {
            uint16_t t = heap[parent];
            heap[parent] = heap[current];
            heap[current] = t;
            t = heap[parent + 1];
            heap[parent + 1] = heap[current + 1];
            heap[current + 1] = t;
        } else {
            break;
        }
        parent = current;
    }
}

/* ------------------------------------------------------------------------------------------
 * reversePackageMerge  (src/RawDeflate.ts:484-571)
 * value[][] holds JS numbers: `undefined` is NaN here (undefined + x = NaN; NaN > y is false).
 * type[][]: -1 is `undefined`.
 * ---------------------------------------------------------------------------------------- */
static uint64_t g_rpm_freq_oob = 0;
uint64_t zo_diag_rpm_freq_oob(void) { return g_rpm_freq_oob; }

typedef struct {
    int symbols, limit;
    uint8_t* code_length;
    int* type[16];
    int size[16];
    int current_position[16];
} rpm_ctx;

static void rpm_take_package(rpm_ctx* c, int j) /* :496-507 */
{
    int cp = c->current_position[j];
    int x = (cp >= 0 && cp < c->size[j]) ? c->type[j][cp] : -1;
    if (x == c->symbols && x >= 0) {
        rpm_take_package(c, j + 1);
        rpm_take_package(c, j + 1);
    } else if (x >= 0 && x < c->symbols) {
        c->code_length[x]--; /* Uint8Array element */
    } /* else: codeLength[undefined]-- / out of range: no effect on the typed array */
    c->current_position[j]++;
}

int zo_reverse_package_merge(const uint32_t* freqs, int symbols, int limit, uint8_t* code_length)
{
    uint16_t minimum_cost[16];
    int flag[16];
    double* value[16];
    rpm_ctx c;
    memset(&c, 0, sizeof c);
    memset(minimum_cost, 0, sizeof minimum_cost);
    if (limit < 1 || limit > 15) return ZO_E_BAD_TYPE;
    c.symbols = symbols;
    c.limit = limit;
    c.code_length = code_length;
    for (int i = 0; i < symbols; ++i) code_length[i] = (uint8_t)limit; /* :487 */

    minimum_cost[limit - 1] = (uint16_t)symbols; /* :509 */
    int32_t excess = (1 << limit) - symbols;     /* :511 */
    int32_t half = 1 << (limit - 1);
    for (int j = 0; j < limit; ++j) { /* :514-523 */
        if (excess < half) {
            flag[j] = 0;
        } else {
            flag[j] = 1;
            excess -= half;
        }
        excess <<= 1;
        if (limit - 2 - j >= 0) /* index -1 write is dropped */
            minimum_cost[limit - 2 - j] = (uint16_t)((minimum_cost[limit - 1 - j] >> 1) + symbols);
    }
    minimum_cost[0] = (uint16_t)flag[0]; /* :525 */
    for (int j = 1; j < limit; ++j)      /* :529-535 */
        if (minimum_cost[j] > 2 * minimum_cost[j - 1] + flag[j])
            minimum_cost[j] = (uint16_t)(2 * minimum_cost[j - 1] + flag[j]);

    int rc = ZO_OK;
    for (int j = 0; j < limit; ++j) {
        c.size[j] = minimum_cost[j];
        value[j] = (double*)malloc(sizeof(double) * (size_t)(c.size[j] + 1));
        c.type[j] = (int*)malloc(sizeof(int) * (size_t)(c.size[j] + 1));
        if (!value[j] || !c.type[j]) rc = ZO_E_NOMEM;
    }
    if (rc != ZO_OK) goto done;
    for (int j = 0; j < limit; ++j)
        for (int t = 0; t < c.size[j]; ++t) {
            value[j][t] = NAN;
            c.type[j][t] = -1;
        }

    for (int t = 0; t < minimum_cost[limit - 1]; ++t) { /* :538-541 */
        value[limit - 1][t] = (t < symbols) ? (double)freqs[t] : NAN;
        c.type[limit - 1][t] = t;
    }
    if (flag[limit - 1]) { /* :543-546 */
        if (symbols > 0) code_length[0]--;
        c.current_position[limit - 1]++;
    }
    for (int j = limit - 2; j >= 0; --j) { /* :548-568 */
        int i = 0;
        int next = c.current_position[j + 1];
        for (int t = 0; t < minimum_cost[j]; ++t) {
            double a = (next >= 0 && next < c.size[j + 1]) ? value[j + 1][next] : NAN;
            double b = (next + 1 >= 0 && next + 1 < c.size[j + 1]) ? value[j + 1][next + 1] : NAN;
            double weight = a + b;
            double fi;
            if (i < symbols) {
                fi = (double)freqs[i];
            } else {
                fi = NAN; /* freqs[i] === undefined: comparison false -> item branch */
                if (symbols >= 2) g_rpm_freq_oob++;
            }
            if (weight > fi) {
                value[j][t] = weight;
                c.type[j][t] = symbols;
                next += 2;
            } else {
                value[j][t] = fi;
                c.type[j][t] = i;
                i++;
            }
        }
        c.current_position[j] = 0;
        if (flag[j]) rpm_take_package(&c, j);
    }
done:
    for (int j = 0; j < limit; ++j) {
        free(value[j]);
        free(c.type[j]);
    }
    return rc;
}

/* src/RawDeflate.ts:440-474 */
int zo_get_lengths(const uint32_t* freqs, int nsym, int limit, uint8_t* lengths)
{
    zo_heap heap;
    heap.length = 0;
    memset(lengths, 0, (size_t)nsym);
    if (nsym > 286) return ZO_E_BAD_TYPE;
    int nodes = 0;
    for (int i = 0; i < nsym; ++i)
        if (freqs[i] > 0) {
            heap_push(&heap, (uint16_t)i, (uint16_t)freqs[i]); /* Uint16Array store: mod 65536 */
            nodes++;
        }
    if (nodes == 1) { /* :455-458 */
        uint16_t idx, val;
        heap_pop(&heap, &idx, &val);
        lengths[idx] = 1;
        return ZO_OK;
    }
    uint16_t index[286];
    uint32_t values[286];
    uint8_t code_length[286];
    for (int i = 0; i < nodes; ++i) { /* :461-465 */
        uint16_t idx, val;
        heap_pop(&heap, &idx, &val);
        index[i] = idx;
        values[i] = val;
    }
    int rc = zo_reverse_package_merge(values, nodes, limit, code_length);
    if (rc != ZO_OK) return rc;
    for (int i = 0; i < nodes; ++i) lengths[index[i]] = code_length[i]; /* :469-471 */
    return ZO_OK;
}

/* src/RawDeflate.ts:580-611 */
void zo_codes_from_lengths(const uint8_t* lengths, int n, uint16_t* codes)
{
    uint32_t count[17] = {0}, start_code[17] = {0};
    uint32_t code = 0;
    for (int i = 0; i < n; ++i)
        if (lengths[i] <= 16) count[lengths[i]]++;
    for (int i = 1; i <= 16; ++i) { /* MaxCodeLength = 16  :20 */
        start_code[i] = code;
        code += count[i];
        code <<= 1;
    }
    for (int i = 0; i < n; ++i) {
        uint16_t r = 0;
        if (lengths[i] > 0) { /* startCode[0] is undefined in JS; the loop below runs 0 times */
            code = start_code[lengths[i]];
            start_code[lengths[i]] += 1;
            for (int j = 0; j < lengths[i]; ++j) {
                r = (uint16_t)((r << 1) | (code & 1));
                code >>= 1;
            }
        }
        codes[i] = r;
    }
}

/* src/RawDeflate.ts:341-431 */
int zo_tree_symbols(int hlit, const uint8_t* litlen_lengths, int hdist, const uint8_t* dist_lengths,
                    uint32_t* result, uint8_t freqs[19])
{
    uint32_t src[286 + 30];
    int l = hlit + hdist, i, j = 0, n_result = 0;
    memset(freqs, 0, 19);
    for (i = 0; i < hlit; ++i) src[j++] = litlen_lengths[i];
    for (i = 0; i < hdist; ++i) src[j++] = dist_lengths[i];

    for (i = 0; i < l; i += j) {
        for (j = 1; i + j < l && src[i + j] == src[i]; ++j) {
        }
        int run_length = j;
        if (src[i] == 0) {
            if (run_length < 3) {
                while (run_length-- > 0) {
                    result[n_result++] = 0;
                    freqs[0]++;
                }
            } else {
                while (run_length > 0) {
                    int rpt = (run_length < 138 ? run_length : 138);
                    if (rpt > run_length - 3 && rpt < run_length) rpt = run_length - 3;
                    if (rpt <= 10) {
                        result[n_result++] = 17;
                        result[n_result++] = (uint32_t)(rpt - 3);
                        freqs[17]++;
                    } else {
                        result[n_result++] = 18;
                        result[n_result++] = (uint32_t)(rpt - 11);
                        freqs[18]++;
                    }
                    run_length -= rpt;
                }
            }
        } else {
            result[n_result++] = src[i];
            freqs[src[i]]++;
            run_length--;
            if (run_length < 3) {
                while (run_length-- > 0) {
                    result[n_result++] = src[i];
                    freqs[src[i]]++;
                }
            } else {
                while (run_length > 0) {
                    int rpt = (run_length < 6 ? run_length : 6);
                    if (rpt > run_length - 3 && rpt < run_length) rpt = run_length - 3;
                    result[n_result++] = 16;
                    result[n_result++] = (uint32_t)(rpt - 3);
                    freqs[16]++;
                    run_length -= rpt;
                }
            }
        }
    }
    return n_result;
}

/* ------------------------------------------------------------------------------------------
 * BitStream  (src/Bitstream.ts:62-130): LSB-first byte filling.
 *   reverse=true : the value's bits go out LSB first; reverse=false: MSB first.
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    uint8_t* buf;
    size_t cap, index;
    int bitindex;
    int overflow;
} zo_bits;

static void bits_write(zo_bits* s, uint32_t number, int b, int reverse)
{
    for (int i = 0; i < b; ++i) {
        uint32_t bit = reverse ? ((number >> i) & 1u) : ((number >> (b - i - 1)) & 1u);
        if (s->index >= s->cap) {
            s->overflow = 1;
            return;
        }
        if (s->bitindex == 0) s->buf[s->index] = 0;
        s->buf[s->index] |= (uint8_t)(bit << s->bitindex);
        if (++s->bitindex == 8) {
            s->bitindex = 0;
            s->index++;
        }
    }
}

static size_t bits_finish(zo_bits* s) /* :112-130 zero-pads the last byte */
{
    if (s->bitindex > 0) {
        s->index++;
        s->bitindex = 0;
    }
    return s->index;
}

/* HuffmanOrder, src/RawInflate.ts:14 */
static const uint8_t kHuffmanOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

size_t zo_raw_deflate_bound(size_t n)
{
    /* <= 15 bits per literal + header (~14 + 19*3 + 316*14 bits) + stored-block framing */
    return n * 2 + 5 * (n / 0xFFFF + 1) + 1024;
}

/* src/RawDeflate.ts:262-297 (dynamic) and :305-330 (fixed) */
static void emit_tokens(zo_bits* s, const uint16_t* tok, size_t ntok, int fixed,
                        const uint16_t* ll_codes, const uint8_t* ll_len,
                        const uint16_t* d_codes, const uint8_t* d_len)
{
    for (size_t index = 0; index < ntok; ++index) {
        uint16_t literal = tok[index];
        if (fixed) {
            /* FixedHuffmanTable :26-41, written MSB-first (reverse omitted => false) :313 */
            uint32_t code;
            int bits;
            if (literal <= 143) {
                code = literal + 0x030u;
                bits = 8;
            } else if (literal <= 255) {
                code = literal - 144u + 0x190u;
                bits = 9;
            } else if (literal <= 279) {
                code = literal - 256u;
                bits = 7;
            } else {
                code = literal - 280u + 0x0C0u;
                bits = 8;
            }
            bits_write(s, code, bits, 0);
        } else {
            bits_write(s, ll_codes[literal], ll_len[literal], 1); /* :279 */
        }
        if (literal > 256) {
            uint16_t ev = tok[++index], eb = tok[++index];
            bits_write(s, ev, eb, 1); /* length extra :284 / :318 */
            uint16_t code = tok[++index];
            if (fixed)
                bits_write(s, code, 5, 0); /* :320 */
            else
                bits_write(s, d_codes[code], d_len[code], 1); /* :287 */
            ev = tok[++index];
            eb = tok[++index];
            bits_write(s, ev, eb, 1); /* distance extra :289 / :322 */
        } else if (literal == 256) {
            break;
        }
    }
}

/* RawDeflate.compress of in[dict_len, n) with in[0, dict_len) as match history and a chosen BFINAL bit.
 * dict_len = 0, bfinal = 1 is the reference's function; the other settings restate the same block construction
 * for the chunk-joined / dictionary-primed streams the engine writes. */
static int raw_deflate_core(const uint8_t* in, size_t n, size_t dict_len, int bfinal, int type, int lazy, uint8_t* out,
                            size_t out_cap, size_t out_index, size_t* out_len)
{
    if (out_index > out_cap) return ZO_E_OUT_OVERFLOW;
    if (dict_len > n) return ZO_E_INPUT_BROKEN;
    if (type == ZO_NONE) { /* :93-100 + makeNocompressBlock :122-153 */
        size_t op = out_index;
        in += dict_len;
        n -= dict_len;
        for (size_t position = 0; position < n;) {
            size_t blen = n - position < 0xFFFF ? n - position : 0xFFFF;
            position += blen;
            if (op + 5 + blen > out_cap) return ZO_E_OUT_OVERFLOW;
            out[op++] = (uint8_t)((position == n && bfinal ? 1 : 0) | (ZO_NONE << 1));
            uint32_t len = (uint32_t)blen, nlen = len ^ 0xFFFFu;
            out[op++] = len & 0xFF;
            out[op++] = (len >> 8) & 0xFF;
            out[op++] = nlen & 0xFF;
            out[op++] = (nlen >> 8) & 0xFF;
            memcpy(out + op, in + position - blen, blen);
            op += blen;
        }
        *out_len = op;
        return ZO_OK;
    }
    if (type != ZO_FIXED && type != ZO_DYNAMIC) return ZO_E_BAD_TYPE; /* :110 */

    uint16_t* tok = (uint16_t*)malloc(sizeof(uint16_t) * (2 * n + 1));
    if (!tok) return ZO_E_NOMEM;
    uint32_t fl[286], fd[30];
    size_t ntok = 0;
    int rc = zo_lz77_encode_dict(in, n, dict_len, lazy, tok, &ntok, fl, fd);
    if (rc != ZO_OK) {
        free(tok);
        return rc;
    }
    zo_bits s = {out, out_cap, out_index, 0, 0};
    bits_write(&s, bfinal ? 1 : 0, 1, 1);     /* bfinal :165/:185 */
    bits_write(&s, (uint32_t)type, 2, 1);     /* btype  :166/:186 */

    if (type == ZO_FIXED) { /* :161-173 */
        emit_tokens(&s, tok, ntok, 1, NULL, NULL, NULL, NULL);
    } else { /* :181-251 */
        uint8_t ll_len[286], d_len[30], tree_len[19], tree_freq8[19];
        uint16_t ll_codes[286], d_codes[30], tree_codes[19];
        uint32_t tree_syms[2 * (286 + 30)], tree_freq32[19];
        rc = zo_get_lengths(fl, 286, 15, ll_len); /* :192 */
        if (rc == ZO_OK) rc = zo_get_lengths(fd, 30, 7, d_len); /* :194 */
        if (rc != ZO_OK) {
            free(tok);
            return rc;
        }
        zo_codes_from_lengths(ll_len, 286, ll_codes);
        zo_codes_from_lengths(d_len, 30, d_codes);
        int hlit, hdist, hclen;
        for (hlit = 286; hlit > 257 && ll_len[hlit - 1] == 0; hlit--) { /* :199 */
        }
        for (hdist = 30; hdist > 1 && d_len[hdist - 1] == 0; hdist--) { /* :200 */
        }
        int nsyms = zo_tree_symbols(hlit, ll_len, hdist, d_len, tree_syms, tree_freq8); /* :203 */
        for (int i = 0; i < 19; ++i) tree_freq32[i] = tree_freq8[i]; /* Uint8Array freqs :346 */
        rc = zo_get_lengths(tree_freq32, 19, 7, tree_len);           /* :204 */
        if (rc != ZO_OK) {
            free(tok);
            return rc;
        }
        uint8_t trans[19];
        for (int i = 0; i < 19; ++i) trans[i] = tree_len[kHuffmanOrder[i]]; /* :206-208 */
        for (hclen = 19; hclen > 4 && trans[hclen - 1] == 0; hclen--) {     /* :209 */
        }
        zo_codes_from_lengths(tree_len, 19, tree_codes); /* :211 */
        bits_write(&s, (uint32_t)(hlit - 257), 5, 1);    /* :214-216 */
        bits_write(&s, (uint32_t)(hdist - 1), 5, 1);
        bits_write(&s, (uint32_t)(hclen - 4), 4, 1);
        for (int i = 0; i < hclen; ++i) bits_write(&s, trans[i], 3, 1); /* :217-219 */
        for (int i = 0; i < nsyms; ++i) {                               /* :222-241 */
            uint32_t code = tree_syms[i];
            bits_write(&s, tree_codes[code], tree_len[code], 1);
            if (code >= 16) {
                int bitlen = code == 16 ? 2 : code == 17 ? 3 : 7;
                i++;
                bits_write(&s, tree_syms[i], bitlen, 1);
            }
        }
        emit_tokens(&s, tok, ntok, 0, ll_codes, ll_len, d_codes, d_len); /* :243-248 */
    }
    free(tok);
    *out_len = bits_finish(&s);
    return s.overflow ? ZO_E_OUT_OVERFLOW : ZO_OK;
}

int zo_raw_deflate(const uint8_t* in, size_t n, int type, int lazy, uint8_t* out, size_t out_cap,
                   size_t out_index, size_t* out_len)
{
    return raw_deflate_core(in, n, 0, 1, type, lazy, out, out_cap, out_index, out_len);
}

int zo_raw_deflate_dict(const uint8_t* in, size_t n, size_t dict_len, int bfinal, int type, uint8_t* out,
                        size_t out_cap, size_t* out_len)
{
    return raw_deflate_core(in, n, dict_len, bfinal, type, 0, out, out_cap, 0, out_len);
}

/* ------------------------------------------------------------------------------------------
 * buildHuffmanTable  (src/Huffman.ts:8-68): single-level table, entry = len<<16 | symbol
 * ---------------------------------------------------------------------------------------- */
int zo_build_huffman_table(const uint8_t* lengths, int n, uint32_t* table, int* max_len, int* min_len)
{
    int maxl = 0, minl = 1 << 30;
    for (int i = 0; i < n; ++i) {
        if (lengths[i] > maxl) maxl = lengths[i];
        if (lengths[i] < minl) minl = lengths[i];
    }
    if (maxl > 15) return ZO_E_CODE_LENGTH;
    uint32_t size = 1u << maxl;
    memset(table, 0, size * sizeof(uint32_t));
    uint32_t code = 0, skip = 2;
    for (int bit_length = 1; bit_length <= maxl;) {
        for (int i = 0; i < n; ++i) {
            if (lengths[i] == bit_length) {
                uint32_t reversed = 0, rtemp = code;
                for (int j = 0; j < bit_length; ++j) {
                    reversed = (reversed << 1) | (rtemp & 1);
                    rtemp >>= 1;
                }
                uint32_t value = ((uint32_t)bit_length << 16) | (uint32_t)i;
                for (uint32_t j = reversed; j < size; j += skip) table[j] = value;
                ++code;
            }
        }
        ++bit_length;
        code <<= 1;
        skip <<= 1;
    }
    *max_len = maxl;
    *min_len = minl;
    return ZO_OK;
}

/* ------------------------------------------------------------------------------------------
 * RawInflate  (src/RawInflate.ts:127-516), ADAPTIVE buffer strategy, caller-bounded output
 * ---------------------------------------------------------------------------------------- */
/* LengthCodeTable / LengthExtraTable / DistCodeTable / DistExtraTable  :17-42
 * (31 length entries: symbols 286/287 decode as 258 in the reference) */
static const uint16_t kInfLenBase[31] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23,  27, 31,
                                         35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258, 258, 258};
static const uint8_t kInfLenExtra[31] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
                                         3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0, 0, 0};

typedef struct {
    const uint8_t* in;
    size_t in_len, ip;
    uint8_t* out;
    size_t out_cap, op;
    uint32_t bitsbuf;
    int bitsbuflen;
    int flags;
    int err;
} inf_state;

typedef struct {
    uint32_t* table;
    int max_len;
} inf_table;

static uint32_t inf_read_bits(inf_state* s, int length) /* :177-207 */
{
    if (s->err) return 0;
    if (s->flags & 1) {
        /* literal end check, off by one */
        long need = ((long)length - s->bitsbuflen + 7) >> 3; /* arithmetic shift like JS */
        if ((long)s->ip + need >= (long)s->in_len) {
            s->err = ZO_E_INPUT_BROKEN;
            return 0;
        }
    }
    while (s->bitsbuflen < length) {
        if (s->ip >= s->in_len) {
            s->err = ZO_E_INPUT_BROKEN;
            return 0;
        }
        s->bitsbuf |= (uint32_t)s->in[s->ip++] << s->bitsbuflen;
        s->bitsbuflen += 8;
    }
    uint32_t octet = s->bitsbuf & ((1u << length) - 1);
    s->bitsbuf >>= length;
    s->bitsbuflen -= length;
    return octet;
}

static int inf_read_code(inf_state* s, const inf_table* t) /* :214-246 */
{
    if (s->err) return 0;
    while (s->bitsbuflen < t->max_len) {
        if (s->ip >= s->in_len) break;
        s->bitsbuf |= (uint32_t)s->in[s->ip++] << s->bitsbuflen;
        s->bitsbuflen += 8;
    }
    uint32_t cwl = t->table[s->bitsbuf & ((1u << t->max_len) - 1)];
    int code_length = (int)(cwl >> 16);
    if (code_length > s->bitsbuflen) {
        s->err = ZO_E_CODE_LENGTH | (code_length << 8); /* 'invalid code length: N' (:238): N rides above bit 7 */
        return 0;
    }
    if (code_length == 0) {
        /* reference: consumes 0 bits and returns symbol 0 forever (incomplete/empty code) */
        s->err = ZO_E_UNDEFINED;
        return 0;
    }
    s->bitsbuf >>= code_length;
    s->bitsbuflen -= code_length;
    return (int)(cwl & 0xFFFF);
}

static void inf_decode_huffman(inf_state* s, const inf_table* litlen, const inf_table* dist) /* :466-516 */
{
    int code;
    while (!s->err && (code = inf_read_code(s, litlen)) != 256) {
        if (s->err) return;
        if (code < 256) {
            if (s->op >= s->out_cap) {
                s->err = ZO_E_OUT_OVERFLOW;
                return;
            }
            s->out[s->op++] = (uint8_t)code;
            continue;
        }
        int ti = code - 257;
        if (ti > 30) {
            s->err = ZO_E_UNDEFINED;
            return;
        }
        uint32_t code_length = kInfLenBase[ti];
        if (kInfLenExtra[ti] > 0) code_length += inf_read_bits(s, kInfLenExtra[ti]);
        code = inf_read_code(s, dist);
        if (s->err) return;
        if (code > 29) { /* DistCodeTable[30..31] undefined in the reference */
            s->err = ZO_E_UNDEFINED;
            return;
        }
        uint32_t code_dist = kDistBase[code];
        if (kDistExtra[code] > 0) code_dist += inf_read_bits(s, kDistExtra[code]);
        if (s->err) return;
        if (code_dist > s->op) { /* reference reads output[-k] === undefined and stores 0 */
            s->err = ZO_E_UNDEFINED;
            return;
        }
        if (s->op + code_length > s->out_cap) {
            s->err = ZO_E_OUT_OVERFLOW;
            return;
        }
        while (code_length--) { /* :506-508 */
            s->out[s->op] = s->out[s->op - code_dist];
            s->op++;
        }
    }
    if (s->err) return;
    while (s->bitsbuflen >= 8) { /* :511-514 */
        s->bitsbuflen -= 8;
        s->ip--;
    }
}

int zo_raw_inflate(const uint8_t* in, size_t in_len, size_t index, uint8_t* out, size_t out_cap,
                   size_t* out_len, size_t* ip_out, int flags)
{
    inf_state s;
    memset(&s, 0, sizeof s);
    s.in = in;
    s.in_len = in_len;
    s.ip = index;
    s.out = out;
    s.out_cap = out_cap;
    s.flags = flags;
    uint32_t* tbl_a = (uint32_t*)malloc(sizeof(uint32_t) * 32768);
    uint32_t* tbl_b = (uint32_t*)malloc(sizeof(uint32_t) * 32768);
    uint32_t tbl_c[128];
    if (!tbl_a || !tbl_b) {
        free(tbl_a);
        free(tbl_b);
        return ZO_E_NOMEM;
    }
    int bfinal = 0;
    while (!bfinal && !s.err) { /* :128-130 */
        uint32_t header = inf_read_bits(&s, 3); /* :146 */
        if (s.err) break;
        if (header & 1) bfinal = 1;
        header >>= 1;
        if (header == 0) { /* parseUncompressedBlock :251-318 */
            s.bitsbuf = 0;
            s.bitsbuflen = 0;
            if (s.ip + 1 >= s.in_len) {
                s.err = ZO_E_STORED_LEN;
                break;
            }
            uint32_t len = s.in[s.ip] | ((uint32_t)s.in[s.ip + 1] << 8);
            s.ip += 2;
            if (s.ip + 1 >= s.in_len) {
                s.err = ZO_E_STORED_LEN | (1 << 8); /* '... header: NLEN' (:272); LEN (:266) has nothing above bit 7 */
                break;
            }
            s.ip += 2; /* NLEN read but never effectively verified (:277 is always false) */
            if (s.ip + len > s.in_len) {
                s.err = ZO_E_INPUT_BROKEN;
                break;
            }
            if (s.op + len > s.out_cap) {
                s.err = ZO_E_OUT_OVERFLOW;
                break;
            }
            memcpy(s.out + s.op, s.in + s.ip, len);
            s.op += len;
            s.ip += len;
        } else if (header == 1 || header == 2) {
            uint8_t ll[320], dl[32];
            int nl, nd, mn;
            inf_table tl, td;
            if (header == 1) { /* FixedLiteralLengthTable / FixedDistanceTable :45-61 */
                nl = 288;
                nd = 30;
                for (int i = 0; i < 288; ++i) ll[i] = i <= 143 ? 8 : i <= 255 ? 9 : i <= 279 ? 7 : 8;
                for (int i = 0; i < 30; ++i) dl[i] = 5;
            } else { /* parseDynamicHuffmanBlock :345-400 */
                int hlit = (int)inf_read_bits(&s, 5) + 257;
                int hdist = (int)inf_read_bits(&s, 5) + 1;
                int hclen = (int)inf_read_bits(&s, 4) + 4;
                uint8_t cl[19];
                memset(cl, 0, sizeof cl);
                for (int i = 0; i < hclen && !s.err; ++i) cl[kHuffmanOrder[i]] = (uint8_t)inf_read_bits(&s, 3);
                if (s.err) break;
                inf_table tc = {tbl_c, 0};
                if (zo_build_huffman_table(cl, 19, tbl_c, &tc.max_len, &mn) != ZO_OK) {
                    s.err = ZO_E_CODE_LENGTH;
                    break;
                }
                uint8_t length_table[288 + 32 + 8];
                memset(length_table, 0, sizeof length_table);
                int total = hlit + hdist, prev = 0;
                for (int i = 0; i < total && !s.err;) {
                    int code = inf_read_code(&s, &tc);
                    if (s.err) break;
                    int repeat;
                    switch (code) {
                        case 16:
                            repeat = 3 + (int)inf_read_bits(&s, 2);
                            while (repeat--) {
                                if (i < total) length_table[i] = (uint8_t)prev; /* OOB write dropped */
                                i++;
                            }
                            break;
                        case 17:
                            repeat = 3 + (int)inf_read_bits(&s, 3);
                            while (repeat--) {
                                if (i < total) length_table[i] = 0;
                                i++;
                            }
                            prev = 0;
                            break;
                        case 18:
                            repeat = 11 + (int)inf_read_bits(&s, 7);
                            while (repeat--) {
                                if (i < total) length_table[i] = 0;
                                i++;
                            }
                            prev = 0;
                            break;
                        default:
                            length_table[i++] = (uint8_t)code;
                            prev = code;
                            break;
                    }
                }
                if (s.err) break;
                nl = hlit;
                nd = hdist;
                memcpy(ll, length_table, (size_t)hlit);
                memcpy(dl, length_table + hlit, (size_t)hdist);
            }
            tl.table = tbl_a;
            td.table = tbl_b;
            if (zo_build_huffman_table(ll, nl, tbl_a, &tl.max_len, &mn) != ZO_OK ||
                zo_build_huffman_table(dl, nd, tbl_b, &td.max_len, &mn) != ZO_OK) {
                s.err = ZO_E_CODE_LENGTH;
                break;
            }
            inf_decode_huffman(&s, &tl, &td);
        } else {
            s.err = ZO_E_BTYPE; /* :168 */
        }
    }
    free(tbl_a);
    free(tbl_b);
    *out_len = s.op;
    *ip_out = s.ip;
    return s.err;
}
