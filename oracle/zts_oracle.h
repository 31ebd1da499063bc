/*
 * zts_oracle.h -- CPU restatement of the zlib.ts hot path (TEST INFRASTRUCTURE).
 *
 * This is the parity oracle: a plain-C restatement of the reference's
 * RawDeflate / LZ77 / Heap / BitStream / RawInflate / Huffman / CRC32 / Adler32
 * (all under /root/reference/src, cited per function in zts_oracle.c).
 * It is NOT part of the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * PARITY PINNED TO THE EXECUTED REFERENCE: the reference ships no golden vectors
 * and the image has no JavaScript engine, so one was written as test
 * infrastructure (oracle/minijs/, C++): it runs /root/reference/dist/Zlib-main.js
 * UNMODIFIED. tests/golden/refjs_vectors.json holds what the reference computes
 * under it (tests/golden/make_refjs_vectors.py regenerates the file) and
 * tests/test_refjs.py holds this oracle to those vectors -- 240 fuzz inputs,
 * 64 KiB benchmark chunks, the 1 MiB single block of config C1, getLengths,
 * RawInflate's .ip, the zlib / gzip / zip containers -- plus a live differential
 * fuzz against the interpreter wherever the reference sources are present.
 * Further witnesses: oracle/js_model.py (independent Python model), SURVEY.md
 * Appendix C (reproduced by the executed reference bit for bit) and CPython zlib.
 */
#ifndef ZTS_ORACLE_H
#define ZTS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* CompressionType, src/RawDeflate.ts:12-17 */
enum { ZO_NONE = 0, ZO_FIXED = 1, ZO_DYNAMIC = 2 };

/* status codes; the text is the reference's thrown message where one exists */
enum {
    ZO_OK = 0,
    ZO_E_INPUT_BROKEN = 1,   /* 'input buffer is broken'            src/RawInflate.ts:188,282 */
    ZO_E_BTYPE = 2,          /* 'unknown BTYPE: 3'                  src/RawInflate.ts:168 */
    ZO_E_CODE_LENGTH = 3,    /* 'invalid code length: N'            src/RawInflate.ts:238 */
    ZO_E_OUT_OVERFLOW = 4,   /* caller's output capacity exceeded (reference grows instead) */
    ZO_E_STORED_LEN = 5,     /* 'invalid uncompressed block header: LEN' / NLEN   :266,:272 */
    ZO_E_BAD_TYPE = 6,       /* 'invalid compression type'          src/RawDeflate.ts:110 */
    ZO_E_NOMEM = 7,
    ZO_E_UNDEFINED = 8       /* reference would spin/produce undefined values (incomplete code hit) */
};

/* ---- checksums ---- */
uint32_t zo_crc32_update(const uint8_t* data, size_t len, uint32_t crc);   /* src/CRC32.ts:25-47 */
uint32_t zo_crc32(const uint8_t* data, size_t len);                       /* src/CRC32.ts:13 */
uint32_t zo_adler32_update(uint32_t adler, const uint8_t* data, size_t len); /* src/Adler32.ts:28-48 */
uint32_t zo_adler32(const uint8_t* data, size_t len);                     /* src/Adler32.ts:13 */

/* ---- LZ77 (src/LZ77.ts:196-283) ----
 * tokens: Uint16Array semantics, capacity 2*n entries (writes past it are dropped, as in JS).
 * *ntok receives the length of the returned subarray (min(pos, 2n)).                           */
int zo_lz77_encode(const uint8_t* in, size_t n, int lazy,
                   uint16_t* tokens, size_t* ntok,
                   uint32_t freqs_litlen[286], uint32_t freqs_dist[30]);

/* ---- Huffman construction ---- */
/* src/RawDeflate.ts:440-474 (+ Heap.ts). freqs as u32; the caller wraps u8 histograms itself. */
int zo_get_lengths(const uint32_t* freqs, int nsym, int limit, uint8_t* lengths);
/* src/RawDeflate.ts:484-571; freqs sorted as the heap pops them. */
int zo_reverse_package_merge(const uint32_t* freqs, int symbols, int limit, uint8_t* code_length);
/* src/RawDeflate.ts:580-611 (bit-reversed canonical codes). */
void zo_codes_from_lengths(const uint8_t* lengths, int n, uint16_t* codes);
/* src/RawDeflate.ts:341-431. codes: up to 316*2 entries; freqs: u8[19]. returns count. */
int zo_tree_symbols(int hlit, const uint8_t* litlen_lengths, int hdist, const uint8_t* dist_lengths,
                    uint32_t* codes, uint8_t freqs[19]);
/* diagnostic counters for the JS-undefined paths in reversePackageMerge */
uint64_t zo_diag_rpm_freq_oob(void);

/* ---- RawDeflate.compress (src/RawDeflate.ts:87-114) ----
 * Writes prefix-free output at out[out_index...]; bytes out[0..out_index) are left untouched
 * (the reference returns a buffer that still holds the caller's prefix).
 * *out_len receives the total length (== reference `.op`).                                   */
int zo_raw_deflate(const uint8_t* in, size_t n, int compression_type, int lazy,
                   uint8_t* out, size_t out_cap, size_t out_index, size_t* out_len);

/* The same block construction over in[dict_len, n) with in[0, dict_len) as LZ77 history (its positions enter the
 * match table like positions inside a match, src/LZ77.ts:217-220, and emit nothing) and a chosen BFINAL bit:
 * the restatement the dictionary-primed mode of the engine is checked against.
 * dict_len = 0, bfinal = 1 is zo_raw_deflate. */
int zo_raw_deflate_dict(const uint8_t* in, size_t n, size_t dict_len, int bfinal, int compression_type,
                        uint8_t* out, size_t out_cap, size_t* out_len);
int zo_lz77_encode_dict(const uint8_t* in, size_t n, size_t dict_len, int lazy, uint16_t* tokens, size_t* ntok,
                        uint32_t fl[286], uint32_t fd[30]);

/* upper bound on zo_raw_deflate output size for n input bytes (excluding out_index) */
size_t zo_raw_deflate_bound(size_t n);

/* ---- RawInflate.decompress (src/RawInflate.ts:127-516, ADAPTIVE buffer) ----
 * in/in_len is the WHOLE container buffer, index the first deflate byte.
 * *ip_out = first byte after the deflate data (reference `.ip`).
 * flags: bit0 = mirror the reference's readBits end-of-input off-by-one (App. B-7) exactly.   */
int zo_raw_inflate(const uint8_t* in, size_t in_len, size_t index,
                   uint8_t* out, size_t out_cap, size_t* out_len, size_t* ip_out, int flags);

/* decoder table builder, src/Huffman.ts:8-68; table must hold 1<<maxlen entries (<= 32768) */
int zo_build_huffman_table(const uint8_t* lengths, int n, uint32_t* table, int* max_len, int* min_len);

#ifdef __cplusplus
}
#endif
#endif
