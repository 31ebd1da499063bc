/*
 * zts_oracle_mt.c -- multi-threaded batch drivers over the CPU oracle (TEST / BASELINE INFRASTRUCTURE).
 * Used only by bench.py's cpu_baseline and --impl reference legs and by its C3 input generator ("zlib streams
 * produced by the reference"): the reference is single-threaded JavaScript; north_star asks for it on 1 thread and on
 * all host cores (worker_threads), i.e. independent chunks / streams sharded over threads, which is what these drivers
 * do with the C restatement.
 */
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#include "zts_oracle.h"

typedef struct {
    const uint8_t* in;
    size_t n, chunk;
    int type;
    volatile long* next;   /* shared work counter */
    size_t n_chunks;
    uint64_t out_bytes;
    int rc;
    /* inflate */
    const uint8_t* comp;
    const uint64_t* offs;
    const uint64_t* lens;
    uint8_t* out;
    size_t out_stride;
} mt_job;

static void* deflate_worker(void* arg)
{
    mt_job* j = (mt_job*)arg;
    size_t cap = zo_raw_deflate_bound(j->chunk);
    uint8_t* buf = (uint8_t*)malloc(cap);
    if (!buf) {
        j->rc = ZO_E_NOMEM;
        return NULL;
    }
    for (;;) {
        long k = __sync_fetch_and_add(j->next, 1);
        if ((size_t)k >= j->n_chunks) break;
        size_t off = (size_t)k * j->chunk;
        size_t len = j->n - off < j->chunk ? j->n - off : j->chunk;
        size_t olen = 0;
        int rc = zo_raw_deflate(j->in + off, len, j->type, 0, buf, cap, 0, &olen);
        if (rc) j->rc = rc;
        j->out_bytes += olen;
    }
    free(buf);
    return NULL;
}

/* RawDeflate of every `chunk`-byte piece of in[0..n), `threads` workers. Returns total compressed bytes. */
int zo_deflate_chunks_mt(const uint8_t* in, size_t n, size_t chunk, int type, int threads, uint64_t* total_out)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    mt_job jobs[256];
    volatile long next = 0;
    size_t n_chunks = (n + chunk - 1) / chunk;
    for (int t = 0; t < threads; ++t) {
        memset(&jobs[t], 0, sizeof(mt_job));
        jobs[t].in = in;
        jobs[t].n = n;
        jobs[t].chunk = chunk;
        jobs[t].type = type;
        jobs[t].next = &next;
        jobs[t].n_chunks = n_chunks;
        pthread_create(&th[t], NULL, deflate_worker, &jobs[t]);
    }
    uint64_t total = 0;
    int rc = 0;
    for (int t = 0; t < threads; ++t) {
        pthread_join(th[t], NULL);
        total += jobs[t].out_bytes;
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    *total_out = total;
    return rc;
}

static void* inflate_worker(void* arg)
{
    mt_job* j = (mt_job*)arg;
    for (;;) {
        long k = __sync_fetch_and_add(j->next, 1);
        if ((size_t)k >= j->n_chunks) break;
        size_t olen = 0, ip = 0;
        int rc = zo_raw_inflate(j->comp + j->offs[k], j->lens[k], 0, j->out + (size_t)k * j->out_stride, j->out_stride,
                                &olen, &ip, 0);
        if (rc) j->rc = rc;
        j->out_bytes += olen;
    }
    return NULL;
}

/* RawInflate of n_streams independent streams (offs/lens into comp) into out + k*out_stride. */
int zo_inflate_streams_mt(const uint8_t* comp, const uint64_t* offs, const uint64_t* lens, size_t n_streams,
                          uint8_t* out, size_t out_stride, int threads, uint64_t* total_out)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    mt_job jobs[256];
    volatile long next = 0;
    for (int t = 0; t < threads; ++t) {
        memset(&jobs[t], 0, sizeof(mt_job));
        jobs[t].comp = comp;
        jobs[t].offs = offs;
        jobs[t].lens = lens;
        jobs[t].out = out;
        jobs[t].out_stride = out_stride;
        jobs[t].next = &next;
        jobs[t].n_chunks = n_streams;
        pthread_create(&th[t], NULL, inflate_worker, &jobs[t]);
    }
    uint64_t total = 0;
    int rc = 0;
    for (int t = 0; t < threads; ++t) {
        pthread_join(th[t], NULL);
        total += jobs[t].out_bytes;
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    *total_out = total;
    return rc;
}

/* ---- zlib streams of independent chunks, kept: stream k = 78 9C | RawDeflate(chunk k) | Adler-32 BE (what
 * Zlib.Deflate.compress is meant to write, src/Deflate.ts:60-99) at out + k * slot, its length in lens[k]. */
typedef struct {
    const uint8_t* in;
    size_t n, chunk, slot;
    uint8_t* out;
    uint64_t* lens;
    volatile long* next;
    size_t n_chunks;
    int rc;
} keep_job;

static void* zlib_keep_worker(void* arg)
{
    keep_job* j = (keep_job*)arg;
    for (;;) {
        long k = __sync_fetch_and_add(j->next, 1);
        if ((size_t)k >= j->n_chunks) break;
        size_t off = (size_t)k * j->chunk;
        size_t len = j->n - off < j->chunk ? j->n - off : j->chunk;
        uint8_t* o = j->out + (size_t)k * j->slot;
        size_t olen = 0;
        o[0] = 0x78;
        o[1] = 0x9C;
        int rc = zo_raw_deflate(j->in + off, len, ZO_DYNAMIC, 0, o + 2, j->slot - 6, 0, &olen);
        if (rc) {
            j->rc = rc;
            j->lens[k] = 0;
            continue;
        }
        uint32_t a = zo_adler32_update(1u, j->in + off, len);
        o[2 + olen] = (uint8_t)(a >> 24);
        o[3 + olen] = (uint8_t)(a >> 16);
        o[4 + olen] = (uint8_t)(a >> 8);
        o[5 + olen] = (uint8_t)a;
        j->lens[k] = olen + 6;
    }
    return NULL;
}

int zo_zlib_chunks_keep_mt(const uint8_t* in, size_t n, size_t chunk, int threads, uint8_t* out, size_t slot,
                           uint64_t* lens)
{
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (slot < zo_raw_deflate_bound(chunk) + 6) return ZO_E_OUT_OVERFLOW;
    pthread_t th[256];
    keep_job jobs[256];
    volatile long next = 0;
    size_t n_chunks = (n + chunk - 1) / chunk;
    for (int t = 0; t < threads; ++t) {
        memset(&jobs[t], 0, sizeof(keep_job));
        jobs[t].in = in;
        jobs[t].n = n;
        jobs[t].chunk = chunk;
        jobs[t].slot = slot;
        jobs[t].out = out;
        jobs[t].lens = lens;
        jobs[t].next = &next;
        jobs[t].n_chunks = n_chunks;
        pthread_create(&th[t], NULL, zlib_keep_worker, &jobs[t]);
    }
    int rc = 0;
    for (int t = 0; t < threads; ++t) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    return rc;
}
