// zts_inflate.cu -- batched RFC-1951 decoder, one warp per independent stream
// (replaces RawInflate.decompress / parseBlock / parseDynamicHuffmanBlock / decodeHuffmanAdaptive,
//  src/RawInflate.ts:127-516, and buildHuffmanTable, src/Huffman.ts:8-68).
//
// Design (B200): the decode state is warp-uniform -- every lane tracks the same bit buffer and
// output cursor, so table look-ups are shared-memory broadcasts and there is no divergence; the
// lanes differ only where they cooperate:
//   * input: the warp keeps a 128-byte window of the stream in registers (one coalesced 4-byte
//     load per lane), words are handed to the bit reader with a shuffle;
//   * decode and copy are split: up to 32 symbols are decoded back to back into one token per lane
//     without touching the output (the serial part is then only bit-buffer arithmetic and shared-
//     memory look-ups), then the batch is written: all literals with one store, all short matches
//     side by side (one per lane), long matches cooperatively (lane k copies byte k, k+32, ...).
//     A match that reads bytes produced inside the same batch cuts the batch into sub-batches that
//     run one after the other, so memory latency is paid per sub-batch, not per symbol;
//   * overlapping matches (dist < len, src/RawInflate.ts:506-508) read the periodic source
//     out[op - dist + k mod dist], which lies entirely before the match;
//   * table construction: canonical code assignment with ballots, no atomics.
// Tables per warp in shared memory: a 10-bit root for literal/length codes and an 8-bit root for
// distance codes (entry = base << 16 | kind << 8 | extra_bits << 4 | code_bits); codes longer than
// the root are resolved canonically (first_code / count per length + symbols sorted by code), which
// replaces the reference's 2^15-entry single-level table.
//
// Algorithmic bytes per stream: C (compressed, read once) + N (output, written once).
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <cstddef>
#include <chrono>

#include "zts_common.cuh"

#ifndef INF_WARPS_PER_CTA
#define INF_WARPS_PER_CTA 1  // a CTA holds its shared memory until its slowest stream is done: one stream per CTA wastes none
#endif
#define LIT_ROOT_BITS 10
#ifndef DIST_ROOT_BITS
#define DIST_ROOT_BITS 7
#endif
#ifndef INF_MIN_CTAS
#define INF_MIN_CTAS 32  // 64 registers; 32 one-warp CTAs per SM (6272 B of tables + 1 KB the system reserves per CTA)
#endif
#define CL_ROOT_BITS 7

#define KIND_LITERAL 0u
#define KIND_BASE 1u   // length or distance base + extra bits
#define KIND_EOB 2u
#define KIND_INVALID 3u
#define INF_TOK_MATCH 0x80000000u  // token of a match: INF_TOK_MATCH | len << 16 | dist; a literal: byte << 16 | (ignored) < 256
#define ROOT_UNRESOLVED (KIND_INVALID << 8)  // root-table entry of a bit pattern whose code is longer than the root (0 bits)

struct HuffTab {
    uint16_t first_code[16];  // canonical (MSB-first) first code of each length
    uint16_t first_idx[16];   // index into sorted[] of the first symbol of each length
    uint16_t count[16];
};

// 6272 bytes per stream: with the 1 KB the system reserves per CTA, 32 one-warp CTAs fill the 228 KB of an SM. What a
// dynamic block header needs only while it is read lives where the tables it leads to are built afterwards: the
// code-length code's root table in the distance root, its sorted symbols and counts in the distance code's, its 19
// lengths in the first bytes of the literal / length code's sorted symbols.
struct InfWarpSmem {
    uint32_t lit_root[1 << LIT_ROOT_BITS];
    uint32_t dist_root[1 << DIST_ROOT_BITS];  // (+ the code-length code's root table)
    uint16_t lit_sorted[288];                  // (+ the code-length code's lengths, 32 bytes)
    uint16_t dist_sorted[32];                  // (+ the code-length code's sorted symbols)
    HuffTab lit, dist;                         // (dist: + the code-length code's)
    uint8_t stage[320];                        // litlen ++ dist code lengths of a header; the token slots of a batch
    uint32_t ring[128];                        // input words on their way into the bit buffer (BitReader)
};
static_assert(CL_ROOT_BITS <= DIST_ROOT_BITS, "the code-length root table lives in the distance root table");
static_assert(sizeof(InfWarpSmem) == 6272, "32 x (sizeof + 1 KB) fills an SM");
static_assert(offsetof(InfWarpSmem, stage) % 4 == 0 && sizeof(InfWarpSmem::stage) >= 128,
              "the token slots of a batch alias the code-length staging area");

// LengthCodeTable / LengthExtraTable (src/RawInflate.ts:17-28; symbols 286/287 decode as 258 there)
__constant__ uint16_t c_len_base[31] = {3,  4,  5,  6,  7,  8,  9,  10, 11,  13,  15,  17,  19,  23,  27, 31,
                                        35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258, 258, 258};
__constant__ uint8_t c_len_extra[31] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2,
                                        3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0, 0, 0};
// DistCodeTable / DistExtraTable (src/RawInflate.ts:31-42)
__constant__ uint16_t c_dist_base[30] = {1,    2,    3,    4,    5,    7,    9,    13,    17,    25,
                                         33,   49,   65,   97,   129,  193,  257,  385,   513,   769,
                                         1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
__constant__ uint8_t c_dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2,  3,  3,  4,  4,  5,  5,  6,
                                         6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
// HuffmanOrder (src/RawInflate.ts:14)
__constant__ uint8_t c_huff_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

enum { TAB_LITLEN = 0, TAB_DIST = 1, TAB_CL = 2 };

__device__ __forceinline__ uint32_t make_entry(int which, uint32_t sym, uint32_t nbits)
{
    if (which == TAB_LITLEN) {
        if (sym < 256) return (sym << 16) | (KIND_LITERAL << 8) | nbits;
        if (sym == 256) return (KIND_EOB << 8) | nbits;
        // (bit 31 marks a match token: it travels with the length base into `len << 16 | dist`, so that a literal's
        // token is its table entry as it is -- byte << 16, code length below -- without a mask in the symbol loop)
        if (sym < 288) return INF_TOK_MATCH | ((uint32_t)c_len_base[sym - 257] << 16) | (KIND_BASE << 8) |
                              ((uint32_t)c_len_extra[sym - 257] << 4) | nbits;
        return (KIND_INVALID << 8) | nbits;
    }
    if (which == TAB_DIST) {
        if (sym < 30) return ((uint32_t)c_dist_base[sym] << 16) | (KIND_BASE << 8) |
                             ((uint32_t)c_dist_extra[sym] << 4) | nbits;
        return (KIND_INVALID << 8) | nbits;  // DistCodeTable[30..31] is undefined in the reference
    }
    return (sym << 16) | (KIND_LITERAL << 8) | nbits;
}

// Canonical decoder construction for one code (warp-cooperative, all lanes call it).
// Mirrors what buildHuffmanTable (src/Huffman.ts:8-68) produces for complete or incomplete codes:
// shorter lengths first, ascending symbol order inside a length. Returns false when the lengths
// are over-subscribed (the reference would silently overwrite table slots).
__device__ bool build_table(int which, const uint8_t* lens, int n, int root_bits, uint32_t* root,
                            uint16_t* sorted, HuffTab* tab)
{
    const unsigned lane = zts_lane();
    const unsigned lt = zts_lanemask_lt();
    // count per length: lane L accumulates count[L]
    uint32_t mycount = 0;
    for (int base = 0; base < n; base += 32) {
        int s = base + (int)lane;
        uint32_t l = s < n ? lens[s] : 0;
#pragma unroll
        for (uint32_t L = 1; L <= 15; ++L) {
            unsigned m = __ballot_sync(0xFFFFFFFFu, l == L);
            if (lane == L) mycount += __popc(m);
        }
    }
    if (lane < 16) tab->count[lane] = (uint16_t)(lane ? mycount : 0);
    __syncwarp();
    // first codes / first indices, Kraft check (every lane computes the same values)
    uint32_t code = 0, idx = 0, kraft = 0;
#pragma unroll
    for (uint32_t L = 1; L <= 15; ++L) {
        uint32_t c = tab->count[L];
        if (lane == L) {
            tab->first_code[L] = (uint16_t)code;
            tab->first_idx[L] = (uint16_t)idx;
        }
        code = (code + c) << 1;
        idx += c;
        kraft += c << (15 - L);
    }
    if (kraft > (1u << 15)) return false;
    // root table: 0 bits = "not resolved here" (long code or unused pattern); the kind field of such an entry says
    // "not a literal", so the symbol loop meets it behind the one test it makes anyway
    for (int i = (int)lane; i < (1 << root_bits); i += 32) root[i] = ROOT_UNRESOLVED;
    __syncwarp();
    // assign codes in (length, symbol) order; running offset per length kept uniformly in registers
    uint32_t offs[16];
#pragma unroll
    for (int L = 0; L < 16; ++L) offs[L] = 0;
    for (int base = 0; base < n; base += 32) {
        int s = base + (int)lane;
        uint32_t l = s < n ? lens[s] : 0;
        uint32_t my_rank = 0;
#pragma unroll
        for (uint32_t L = 1; L <= 15; ++L) {
            unsigned m = __ballot_sync(0xFFFFFFFFu, l == L);
            if (l == L) my_rank = offs[L] + __popc(m & lt);
            offs[L] += __popc(m);
        }
        if (l) {
            uint32_t slot = tab->first_idx[l] + my_rank;
            sorted[slot] = (uint16_t)s;
            uint32_t c = tab->first_code[l] + my_rank;  // MSB-first canonical code
            if ((int)l <= root_bits) {
                uint32_t r = __brev(c) >> (32 - l);  // as it appears in the LSB-first bit buffer
                uint32_t e = make_entry(which, (uint32_t)s, l);
                for (uint32_t j = r; j < (1u << root_bits); j += (1u << l)) root[j] = e;
            }
        }
    }
    __syncwarp();
    return true;
}

// canonical resolution of a code that the root table did not resolve; `bits` = bit buffer (LSB first)
__device__ __forceinline__ uint32_t slow_decode(int which, uint32_t bits, int root_bits, const uint16_t* sorted,
                                                const HuffTab* tab)
{
    uint32_t v = __brev(bits) >> 17;  // first 15 stream bits, MSB-first
    for (int L = root_bits + 1; L <= 15; ++L) {
        uint32_t c = v >> (15 - L);
        uint32_t d = c - tab->first_code[L];
        if (d < tab->count[L]) return make_entry(which, sorted[tab->first_idx[L] + d], (uint32_t)L);
    }
    return 0;  // undefined code
}

// Input words reach the bit buffer through a ring of 128 words in shared memory: a window of 32 words (one per lane,
// requested a window ahead) is appended whenever fewer than INF_RING_MIN unread words are left, which the symbol loop
// checks once per batch -- a batch of 32 symbols takes at most 32 x 48 bits, i.e. at most 50 words -- so that its refills are one shared-memory
// load at a running address.
#define INF_RING_WORDS 128u
#define INF_RING_MIN 52u
struct BitReader {
    const uint32_t* next_base;  // aligned address of the window behind the prefetched one
    const uint8_t* end;         // one past the last readable input byte
    uint32_t win_next;          // this lane's word of the next window to append (prefetched)
    uint32_t ring_s;            // shared-window address of this warp's ring
    uint32_t rd, wr;            // words consumed / appended since the last (re)start
    unsigned long long buf;
    int cnt;                    // valid bits in buf
    unsigned long long base_bits;    // stream bits consumed before the last (re)start
    uint32_t skip_bits;         // bits of the first word that precede the (re)start point
};

__device__ __forceinline__ uint32_t br_fetch(const BitReader& br, const uint32_t* base)
{
    const uint32_t* p = base + zts_lane();
    // a word is readable if it overlaps [.., end); bytes past `end` inside that word are never
    // consumed as data because every consumer checks the consumed-bit count against the item length
    return ((const uint8_t*)p < br.end) ? __ldg(p) : 0u;
}

// appends the prefetched window and requests the one behind it (all lanes; the slots it overwrites were read at
// least INF_RING_WORDS - 32 - INF_RING_MIN words ago)
__device__ __forceinline__ void br_append(BitReader& br)
{
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(br.ring_s + (((br.wr + zts_lane()) & (INF_RING_WORDS - 1u)) << 2)), "r"(br.win_next)
                 : "memory");
    br.wr += 32u;
    br.win_next = br_fetch(br, br.next_base);
    br.next_base += 32;
    __syncwarp();
}

__device__ __forceinline__ uint32_t br_ring_word(const BitReader& br, uint32_t i)
{
    uint32_t w;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(br.ring_s + ((i & (INF_RING_WORDS - 1u)) << 2)) : "memory");
    return w;
}

__device__ __forceinline__ void br_init(BitReader& br, const uint8_t* src, const uint8_t* end,
                                        unsigned long long base_bits)
{
    uintptr_t a = (uintptr_t)src;
    const uint32_t* base = (const uint32_t*)(a & ~(uintptr_t)3);
    br.end = end;
    br.base_bits = base_bits;
    br.skip_bits = (uint32_t)(a & 3) * 8;
    br.rd = br.wr = 0;
    __syncwarp();  // (a restart: every lane is done with the ring)
    br.win_next = br_fetch(br, base);
    br.next_base = base + 32;
    br_append(br);
    br_append(br);
    // first word: drop the bytes before the stream
    const uint32_t w = br_ring_word(br, 0);
    br.rd = 1;
    br.buf = (unsigned long long)(w >> br.skip_bits);
    br.cnt = 32 - (int)br.skip_bits;
}

__device__ __forceinline__ void br_refill(BitReader& br)
{
    if (br.cnt <= 32) {
        if (br.rd == br.wr) br_append(br);
        const uint32_t w = br_ring_word(br, br.rd);
        br.rd++;
        br.buf |= (unsigned long long)w << br.cnt;
        br.cnt += 32;
    }
}

__device__ __forceinline__ uint32_t br_take(BitReader& br, int n)
{
    uint32_t v = (uint32_t)br.buf & ((1u << n) - 1u);
    br.buf >>= n;
    br.cnt -= n;
    return v;
}

// stream bits consumed so far
__device__ __forceinline__ unsigned long long br_bits_used(const BitReader& br)
{
    return br.base_bits + (unsigned long long)br.rd * 32ull - br.skip_bits - (unsigned long long)br.cnt;
}

// ---- which error does the reference raise when the input ends early? ---------------------------------------------
// readBits throws 'input buffer is broken' (src/RawInflate.ts:188), readCodeByTable looks the code up in what is left
// (zero padded) and throws 'invalid code length: N' when the N bits of that code are not all there (:238). The decode
// loops only notice that a batch ran past the end; these helpers replay from a known bit position, one symbol at a
// time, and tell the two cases apart. Rare path, every lane computes the same.
// 32 bits of the stream at bit position `pos`, zero padded behind the end of the input
__device__ uint32_t inf_peek(const uint8_t* src, unsigned long long in_len, unsigned long long pos)
{
    const unsigned long long byte = pos >> 3;
    unsigned long long v = 0;
    for (uint32_t k = 0; k < 6; ++k)
        if (byte + k < in_len) v |= (unsigned long long)src[byte + k] << (8u * k);
    return (uint32_t)(v >> (pos & 7u));
}
// status of a Huffman code of the code-length alphabet read at `pos` (dynamic block header)
__device__ uint32_t inf_classify_cl(const InfWarpSmem* S, const uint8_t* src, unsigned long long in_len, unsigned long long pos)
{
    const uint32_t e = S->dist_root[inf_peek(src, in_len, pos) & ((1u << CL_ROOT_BITS) - 1u)];
    const uint32_t nb = e & 15u;
    if (nb && pos + nb > in_len * 8ull) return ZLB_ST_CODE_LENGTH | (nb << 8);
    return ZLB_ST_INPUT_BROKEN;  // the code was there: its repeat count was not
}
// replays the symbols of a block from `pos` until the input runs out
__device__ uint32_t inf_classify_symbols(const InfWarpSmem* S, const uint8_t* src, unsigned long long in_len, unsigned long long pos)
{
    const unsigned long long in_bits = in_len * 8ull;
    for (uint32_t k = 0; k < 64 && pos < in_bits + 64; ++k) {
        uint32_t bits = inf_peek(src, in_len, pos);
        uint32_t e = S->lit_root[bits & ((1u << LIT_ROOT_BITS) - 1u)];
        if ((e & 15u) == 0) e = slow_decode(TAB_LITLEN, bits, LIT_ROOT_BITS, S->lit_sorted, &S->lit);
        if (e == 0) return ZLB_ST_BAD_CODE;
        if (pos + (e & 15u) > in_bits) return ZLB_ST_CODE_LENGTH | ((e & 15u) << 8);
        pos += e & 15u;
        if (!(e & 0x300u)) continue;              // literal
        if ((e & 0x300u) != (KIND_BASE << 8)) break;  // end of block / invalid symbol: not an input problem
        pos += (e >> 4) & 15u;                    // length extra bits (readBits)
        if (pos > in_bits) return ZLB_ST_INPUT_BROKEN;
        bits = inf_peek(src, in_len, pos);
        uint32_t d = S->dist_root[bits & ((1u << DIST_ROOT_BITS) - 1u)];
        if ((d & 15u) == 0) d = slow_decode(TAB_DIST, bits, DIST_ROOT_BITS, S->dist_sorted, &S->dist);
        if (d == 0) return ZLB_ST_BAD_CODE;
        if (pos + (d & 15u) > in_bits) return ZLB_ST_CODE_LENGTH | ((d & 15u) << 8);
        pos += d & 15u;
        pos += (d >> 4) & 15u;                    // distance extra bits (readBits)
        if (pos > in_bits) return ZLB_ST_INPUT_BROKEN;
    }
    return ZLB_ST_INPUT_BROKEN;
}

__global__ void __launch_bounds__(INF_WARPS_PER_CTA * 32, INF_MIN_CTAS)
inflate_warp_kernel(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const zlb_item* __restrict__ items,
                    zlb_result* __restrict__ results, uint32_t n_items, uint32_t flags)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const unsigned lane = zts_lane();
    const unsigned warp = threadIdx.x >> 5;
    InfWarpSmem* S = reinterpret_cast<InfWarpSmem*>(smem_raw) + warp;
    const uint32_t item = blockIdx.x * INF_WARPS_PER_CTA + warp;
    if (item >= n_items) return;

    const zlb_item it = items[item];
    const uint8_t* src = in + it.in_off;
    uint8_t* dst = out + it.out_off;
    const unsigned long long in_bits = it.in_len * 8ull;
    const unsigned long long cap = it.out_cap;

    BitReader br;
    br.ring_s = (uint32_t)__cvta_generic_to_shared(S->ring);
    br_init(br, src, src + it.in_len, 0);

    unsigned long long op = 0;
    uint32_t status = ZLB_ST_OK;
    uint32_t blocks = 0;
    int have_fixed = 0;
    bool bfinal = false;

    while (!bfinal && status == ZLB_ST_OK) {
        // segment mode: a piece of a larger stream ends cleanly where a block ends and its input is used up
        if ((flags & ZLB_INFLATE_SEGMENT) && blocks > 0 && br_bits_used(br) == in_bits) break;
        br_refill(br);
        if (br_bits_used(br) + 3 > in_bits) {
            status = ZLB_ST_INPUT_BROKEN;
            break;
        }
        uint32_t hdr = br_take(br, 3);  // src/RawInflate.ts:146-152
        bfinal = hdr & 1u;
        uint32_t btype = hdr >> 1;
        blocks++;

        if (btype == 0) {
            // ---- stored block (src/RawInflate.ts:251-318): drop to the byte boundary, LEN, NLEN
            unsigned long long used = br_bits_used(br);
            unsigned long long byte_pos = (used + 7) >> 3;
            if (byte_pos + 4 > it.in_len) {
                // src/RawInflate.ts:266 (LEN) / :272 (NLEN): which of the two 16-bit fields is cut short
                status = byte_pos + 2 > it.in_len ? ZLB_ST_STORED_LEN : (ZLB_ST_STORED_LEN | (1u << 8));
                break;
            }
            const uint8_t* p = src + byte_pos;
            uint32_t len = (uint32_t)p[0] | ((uint32_t)p[1] << 8);
            uint32_t nlen = (uint32_t)p[2] | ((uint32_t)p[3] << 8);
            if ((flags & ZLB_INFLATE_CHECK_NLEN) && ((len ^ nlen) != 0xFFFFu)) {
                status = ZLB_ST_STORED_LEN;
                break;
            }
            byte_pos += 4;
            if (byte_pos + len > it.in_len) {
                status = ZLB_ST_INPUT_BROKEN;
                break;
            }
            if (op + len > cap) {
                status = ZLB_ST_OUT_OVERFLOW;
                break;
            }
            for (uint32_t k = lane; k < len; k += 32) dst[op + k] = src[byte_pos + k];
            op += len;
            br_init(br, src + byte_pos + len, src + it.in_len, (byte_pos + len) * 8ull);
            __syncwarp();
            continue;
        }
        if (btype == 3) {
            status = ZLB_ST_BTYPE;  // src/RawInflate.ts:168
            break;
        }

        if (btype == 1) {
            // ---- fixed Huffman tables (src/RawInflate.ts:45-61)
            if (have_fixed != 1) {
                for (int i = (int)lane; i < 288; i += 32) S->stage[i] = i <= 143 ? 8 : i <= 255 ? 9 : i <= 279 ? 7 : 8;
                for (int i = (int)lane; i < 30; i += 32) S->stage[288 + i] = 5;
                __syncwarp();
                build_table(TAB_LITLEN, S->stage, 288, LIT_ROOT_BITS, S->lit_root, S->lit_sorted, &S->lit);
                build_table(TAB_DIST, S->stage + 288, 30, DIST_ROOT_BITS, S->dist_root, S->dist_sorted, &S->dist);
                have_fixed = 1;
            }
        } else {
            // ---- dynamic Huffman header (src/RawInflate.ts:345-388)
            have_fixed = 0;
            br_refill(br);
            uint32_t hlit = br_take(br, 5) + 257;
            uint32_t hdist = br_take(br, 5) + 1;
            uint32_t hclen = br_take(br, 4) + 4;
            uint8_t* const cl_lens = reinterpret_cast<uint8_t*>(S->lit_sorted);  // (dead before that table is built)
            if (lane < 19) cl_lens[lane] = 0;
            __syncwarp();
            for (uint32_t i = 0; i < hclen; ++i) {
                br_refill(br);
                uint32_t v = br_take(br, 3);
                if (lane == 0) cl_lens[c_huff_order[i]] = (uint8_t)v;
            }
            __syncwarp();
            if (br_bits_used(br) > in_bits) {
                status = ZLB_ST_INPUT_BROKEN;
                break;
            }
            if (!build_table(TAB_CL, cl_lens, 19, CL_ROOT_BITS, S->dist_root, S->dist_sorted, &S->dist)) {
                status = ZLB_ST_BAD_LENGTHS;
                break;
            }
            // code lengths with repeat codes 16/17/18 (:361-383); over-long repeats are clipped like
            // the reference's out-of-range typed-array stores
            const uint32_t total = hlit + hdist;
            uint32_t i = 0, prev = 0;
            while (i < total && status == ZLB_ST_OK) {
                br_refill(br);
                const unsigned long long cl_pos = br_bits_used(br);
                uint32_t e = S->dist_root[(uint32_t)br.buf & ((1u << CL_ROOT_BITS) - 1u)];  // (the code-length code's table)
                uint32_t nb = e & 15u;
                if (nb == 0) {
                    status = ZLB_ST_BAD_CODE;
                    break;
                }
                br_take(br, (int)nb);
                uint32_t sym = e >> 16;
                uint32_t rep = 1, val = sym;
                if (sym == 16) {
                    rep = 3 + br_take(br, 2);
                    val = prev;
                } else if (sym == 17) {
                    rep = 3 + br_take(br, 3);
                    val = 0;
                } else if (sym == 18) {
                    rep = 11 + br_take(br, 7);
                    val = 0;
                }
                // lanes store the run cooperatively (rep <= 138)
                for (uint32_t k = lane; k < rep; k += 32)
                    if (i + k < total) S->stage[i + k] = (uint8_t)val;
                i += rep;
                prev = val;
                if (br_bits_used(br) > in_bits) status = inf_classify_cl(S, src, it.in_len, cl_pos);
            }
            if (status != ZLB_ST_OK) break;
            __syncwarp();
            // stage[0 .. total) holds litlen then dist lengths
            if (!build_table(TAB_LITLEN, S->stage, (int)hlit, LIT_ROOT_BITS, S->lit_root, S->lit_sorted, &S->lit) ||
                !build_table(TAB_DIST, S->stage + hlit, (int)hdist, DIST_ROOT_BITS, S->dist_root, S->dist_sorted,
                             &S->dist)) {
                status = ZLB_ST_BAD_LENGTHS;
                break;
            }
        }

        // ---- symbol loop (src/RawInflate.ts:466-516), 32 symbols per batch
        bool eob = false;
        while (!eob && status == ZLB_ST_OK) {
            // -- decode: token of symbol i ends up in lane i: literal = byte << 16 | junk < 256, match = INF_TOK_MATCH | len << 16 | dist.
            //    The tables are addressed through 32-bit shared-window addresses held in registers (a generic
            //    pointer to the per-warp slice gets rematerialised from %tid on every look-up otherwise).
            uint32_t mytok = 0, ntok = 0;
            uint32_t stop = 0;  // 1 = end of block, 2 = undefined code / symbol
            const unsigned long long batch_pos = br_bits_used(br);  // (only looked at when the input ends inside the batch)
            {
                // the address of the warp's slice is warp-uniform; routing it through a shuffle keeps it in a register
                // (ptxas otherwise recomputes it from %tid and %cluster_ctaid in front of every look-up: 9 instructions),
                // and everything in the slice is that register plus a constant
                const uint32_t smem_s = __shfl_sync(0xFFFFFFFFu, (uint32_t)__cvta_generic_to_shared(S), 0);
                const uint32_t lit_s = smem_s + (uint32_t)offsetof(InfWarpSmem, lit_root);
                const uint32_t dist_s = smem_s + (uint32_t)offsetof(InfWarpSmem, dist_root);
                // same for the end-of-input pointer of the bit reader (recomputed from the item table otherwise)
                br.end = reinterpret_cast<const uint8_t*>(__shfl_sync(0xFFFFFFFFu, (unsigned long long)br.end, 0));
                // with a whole batch worth of words in the ring the refills below need no test
                while (br.wr - br.rd < INF_RING_MIN) br_append(br);
                unsigned long long buf = br.buf;
                int cnt = br.cnt;
                const uint32_t ring_s = smem_s + (uint32_t)offsetof(InfWarpSmem, ring), ring_e = ring_s + 4u * INF_RING_WORDS;
                const uint32_t ra0 = ring_s + ((br.rd & (INF_RING_WORDS - 1u)) << 2);
                uint32_t ra = ra0;  // address of the next word
                // the batch's tokens go through shared memory (one store per symbol instead of a compare + select into the
                // lane that owns the slot); the code-length staging area is dead while symbols are decoded
                const uint32_t tok_s = smem_s + (uint32_t)offsetof(InfWarpSmem, stage);
                uint32_t ta = tok_s;
                const uint32_t tok_e = tok_s + 128u;
                // a word from the ring if 32 more bits fit (a length code: before it reads on, which is enough for its
                // extra bits and the distance, 5 + 15 + 13 bits)
#define INF_TAKE_WORD()                                                                       \
    if (cnt <= 32) {                                                                          \
        uint32_t w_;                                                                          \
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w_) : "r"(ra) : "memory");              \
        ra += 4u;                                                                             \
        ra = ra == ring_e ? ring_s : ra;                                                      \
        buf |= (unsigned long long)w_ << cnt;                                                 \
        cnt += 32;                                                                            \
    }
#define INF_SYMBOL(SLOT, TOP, TAIL) { \
                    if (TOP) INF_TAKE_WORD() \
                    uint32_t e; \
                    asm volatile("{\n\t.reg .b32 a;\n\tmad.lo.u32 a, %1, 4, %2;\n\tld.shared.u32 %0, [a];\n\t}" /* (index * 4 + base in one instruction) */ \
                                 : "=r"(e) : "r"((uint32_t)buf & ((1u << LIT_ROOT_BITS) - 1u)), "r"(lit_s) : "memory"); \
                    buf >>= (e & 15u); \
                    cnt -= (int)(e & 15u); \
                    uint32_t tokv = e; /* (a literal's entry is its token) */ \
                    if (e & 0x300u) { \
                        bool is_match = (e & 0x300u) == (KIND_BASE << 8); \
                        if (!is_match) { /* unresolved by the root table, end of block or an undefined symbol */ \
                            if ((e & 15u) == 0) { \
                                e = slow_decode(TAB_LITLEN, (uint32_t)buf, LIT_ROOT_BITS, S->lit_sorted, &S->lit); \
                                if (e == 0) { \
                                    stop = 2; \
                                    { ta += 4u * (SLOT); break; } \
                                } \
                                buf >>= (e & 15u); \
                                cnt -= (int)(e & 15u); \
                                tokv = e; \
                            } \
                            const uint32_t kind = e & 0x300u; \
                            if (kind > (KIND_BASE << 8)) { \
                                stop = (e & 0x100u) ? 2u : 1u; /* KIND_INVALID (3) / KIND_EOB (2) */ \
                                { ta += 4u * (SLOT); break; } \
                            } \
                            is_match = kind != 0u; \
                        } \
                        if (is_match) { \
                            INF_TAKE_WORD() \
                            const uint32_t xb = (e >> 4) & 15u; \
                            const uint32_t len = (e >> 16) + ((uint32_t)buf & ((1u << xb) - 1u)); \
                            buf >>= xb; \
                            cnt -= (int)xb; \
                            uint32_t d; \
                            asm volatile("{\n\t.reg .b32 a;\n\tmad.lo.u32 a, %1, 4, %2;\n\tld.shared.u32 %0, [a];\n\t}" \
                                         : "=r"(d) : "r"((uint32_t)buf & ((1u << DIST_ROOT_BITS) - 1u)), "r"(dist_s) : "memory"); \
                            if ((d & 0x300u) != (KIND_BASE << 8)) { \
                                if ((d & 15u) == 0) d = slow_decode(TAB_DIST, (uint32_t)buf, DIST_ROOT_BITS, S->dist_sorted, &S->dist); \
                                if ((d & 0x300u) != (KIND_BASE << 8)) { \
                                    stop = 2; \
                                    { ta += 4u * (SLOT); break; } \
                                } \
                            } \
                            buf >>= (d & 15u); \
                            cnt -= (int)(d & 15u); \
                            const uint32_t db = (d >> 4) & 15u; \
                            const uint32_t dist = (d >> 16) + ((uint32_t)buf & ((1u << db) - 1u)); \
                            buf >>= db; \
                            cnt -= (int)db; \
                            tokv = (len << 16) | dist; \
                            if (TAIL) INF_TAKE_WORD() \
                        } \
                    } \
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(ta + 4u * (SLOT)), "r"(tokv) : "memory"); }
                // Two symbols per trip (the token slots of a batch are an even number): one loop test for both, and the
                // second one does not ask for a word: behind a literal that was read right after that test at least 17
                // bits are left, enough for any code, and a match of the first symbol takes a word when it is done.
                do {
                    INF_SYMBOL(0u, true, true)
                    INF_SYMBOL(1u, false, false)
                    ta += 8u;
                } while (ta != tok_e);
#undef INF_SYMBOL
#undef INF_TAKE_WORD
                ntok = (ta - tok_s) >> 2;
                __syncwarp();
                if (lane < ntok) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(mytok) : "r"(tok_s + 4u * lane) : "memory");
                br.buf = buf;
                br.cnt = cnt;
                br.rd += ((ra - ra0) & (4u * INF_RING_WORDS - 1u)) >> 2;  // (fewer than INF_RING_WORDS words per batch)
            }
            if (stop == 1) eob = true;
            if (stop == 2) status = ZLB_ST_BAD_CODE;
            if (br_bits_used(br) > in_bits) {
                // the batch ran past the end of the input: nothing of it is trusted; which error it is -- a code or
                // extra bits cut short -- is found by replaying the batch
                status = inf_classify_symbols(S, src, it.in_len, batch_pos);
                break;
            }
            // -- output positions: exclusive prefix sum of the token lengths
            const bool mine = lane < ntok;
            const uint32_t dist = mytok & 0xFFFFu;  // (of a match)
            const bool is_match = mine && (mytok & INF_TOK_MATCH);
            const uint32_t len = !mine ? 0u : (is_match ? (mytok >> 16) & 0x7FFFu : 1u);
            uint32_t inc = len;
#pragma unroll
            for (int sft = 1; sft < 32; sft <<= 1) {
                const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, sft);
                if (lane >= (unsigned)sft) inc += t;
            }
            // Everything below works with 32-bit offsets from the batch's first byte (a batch is at most 32 x 258 bytes):
            // `rel` = where this token's bytes start; the output written so far and the room left are clamped to
            // values that decide the same (a distance is at most 32768, a batch at most 8256 bytes).
            const uint32_t rel = inc - len;
            uint8_t* const bdst = dst + op;  // first byte of the batch
            const uint32_t behind = op > 0x00FFFFFFull ? 0x00FFFFFFu : (uint32_t)op;         // bytes in front of the batch
            const uint32_t room = cap - op > 0x00FFFFFFull ? 0x00FFFFFFu : (uint32_t)(cap - op);  // (op <= cap always)
            // a distance beyond the start of the output (the reference would copy `undefined` -> 0) or an
            // output slot that is too small ends the batch before the offending token
            const bool bad_dist = is_match && dist > behind + rel;
            const bool over = mine && rel + len > room;
            const unsigned bad_mask = __ballot_sync(0xFFFFFFFFu, bad_dist || over);
            uint32_t nvalid = ntok;
            if (bad_mask) {
                nvalid = (uint32_t)__ffs((int)bad_mask) - 1u;
                const bool first_is_dist = __shfl_sync(0xFFFFFFFFu, (int)bad_dist, (int)nvalid) != 0;
                status = first_is_dist ? ZLB_ST_BAD_CODE : ZLB_ST_OUT_OVERFLOW;
            }
            const bool live = lane < nvalid;
            // -- literals: one store
            ZTS_ASSERT(!live || (op + rel + len <= cap && (!is_match || dist <= op + rel)));
            if (live && !is_match) bdst[rel] = (uint8_t)(mytok >> 16);
            __syncwarp();  // (a match may read the literals of its own batch)
            // -- matches, sub-batch by sub-batch: a sub-batch ends in front of the first match that reads what a match
            //    of the same sub-batch writes
            uint32_t start = 0;
            unsigned match_mask = __ballot_sync(0xFFFFFFFFu, live && is_match);
            // where this match's source ends (exclusive), as an offset from the batch's first byte: negative when it
            // lies in front of the batch
            const int src_end = (int)rel - (int)dist + (int)(len < dist ? len : dist);
            while (match_mask) {
                // first byte the matches of this sub-batch write (its first match's: match_mask holds lanes >= start)
                const int sub_start = (int)__shfl_sync(0xFFFFFFFFu, rel, __ffs((int)match_mask) - 1);
                // reads bytes written by this very sub-batch? (source end beyond that byte)
                const bool dep = live && is_match && lane >= start && src_end > sub_start;
                const unsigned dep_mask = __ballot_sync(0xFFFFFFFFu, dep);
                const uint32_t cut = dep_mask ? (uint32_t)__ffs((int)dep_mask) - 1u : 32u;  // behind the first match always
                const bool in_sub = live && is_match && lane >= start && lane < cut;
                const bool small = in_sub && len <= 8u && dist >= len;
                if (__any_sync(0xFFFFFFFFu, small)) {
                    // short non-overlapping matches: every lane copies its own, loads first, then stores
                    uint8_t v[8];
                    uint8_t* dp = bdst + rel;
                    const uint8_t* sp = dp - dist;
#pragma unroll
                    for (int k = 0; k < 8; ++k) v[k] = (small && (uint32_t)k < len) ? sp[k] : (uint8_t)0;
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        if (small && (uint32_t)k < len) dp[k] = v[k];
                }
                unsigned big_mask = __ballot_sync(0xFFFFFFFFu, in_sub && !small);
                while (big_mask) {
                    const int j = __ffs((int)big_mask) - 1;
                    big_mask &= big_mask - 1;
                    const uint32_t jl = __shfl_sync(0xFFFFFFFFu, len, j);
                    const uint32_t jd = __shfl_sync(0xFFFFFFFFu, dist, j);
                    const uint32_t jr = __shfl_sync(0xFFFFFFFFu, rel, j);
                    uint8_t* o = bdst + jr;
                    const uint8_t* s0 = o - jd;
                    if (jd >= jl) {
                        for (uint32_t k = lane; k < jl; k += 32) o[k] = s0[k];
                    } else {
                        // periodic source: byte k comes from s0[k mod dist]
                        uint32_t r = lane % jd;
                        const uint32_t step = 32u % jd;
                        for (uint32_t k = lane; k < jl; k += 32) {
                            o[k] = s0[r];
                            r += step;
                            if (r >= jd) r -= jd;
                        }
                    }
                }
                __syncwarp();  // the next sub-batch reads what this one wrote
                start = cut;
                match_mask = cut < 32u ? (match_mask >> cut) << cut : 0u;
            }
            // bytes of the tokens that were written (all of them unless the batch was cut short)
            op += nvalid ? __shfl_sync(0xFFFFFFFFu, inc, (int)nvalid - 1) : 0u;
            __syncwarp();
        }
        if (status == ZLB_ST_OK && br_bits_used(br) > in_bits) status = ZLB_ST_INPUT_BROKEN;
    }

    if (lane == 0) {
        zlb_result r = results[item];
        r.status = status;
        r.blocks = blocks | ((flags & ZLB_INFLATE_SEGMENT) && bfinal ? 0x80000000u : 0u);  // segment mode: saw BFINAL
        r.out_len = op;
        r.in_used = (br_bits_used(br) + 7) >> 3;  // whole unread bytes are given back (:511-514)
        results[item] = r;
    }
}

static int inflate_launch(zlb_ctx* ctx, cudaStream_t st, const uint8_t* d_in, uint8_t* d_out, const zlb_item* d_items,
                          zlb_result* d_results, size_t n, uint32_t flags)
{
    const size_t smem = sizeof(InfWarpSmem) * INF_WARPS_PER_CTA;
    ZTS_CUDA(ctx, cudaFuncSetAttribute(inflate_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    unsigned grid = (unsigned)((n + INF_WARPS_PER_CTA - 1) / INF_WARPS_PER_CTA);
    ZTS_LAUNCH(ctx, ZK_INFLATE,
               inflate_warp_kernel<<<grid, INF_WARPS_PER_CTA * 32, smem, st>>>(d_in, d_out, d_items, d_results,
                                                                              (uint32_t)n, flags));
    return ZLB_OK;
}

// ---- marker split (ZLB_INFLATE_SPLIT) ------------------------------------------------------------------------
// A stream written by zlb_deflate_batch is a sequence of independent pieces: every chunk was compressed on its own
// (no match reaches into an earlier chunk) and ends with an empty stored block, `00 00 FF FF` after byte alignment.
// One warp per piece decodes a 1 GiB stream in the time of a 64 KiB one. Nothing is assumed about the input: the
// pieces are decoded into scratch slots in segment mode, and only if every piece (a) ends exactly where the next
// one starts, (b) never reaches before its own first byte and (c) fits its slot is the result gathered; a false
// marker or a foreign stream with history across blocks fails one of these and the item is decoded serially.
#define SPLIT_MIN_BYTES (256u << 10)   // items at least this large are worth splitting
#define SPLIT_SLOT (128u << 10)        // scratch bytes per piece (pieces of this engine's streams are <= 64 KiB)
#define SPLIT_MAX_MARKS (1u << 22)
#define SPLIT_MIN_PIECE 64u            // compressed bytes a piece has at least (64 KiB of zeros deflate to 78 + the marker)

__global__ void __launch_bounds__(256)
marker_scan_kernel(const uint8_t* __restrict__ in, unsigned long long n, unsigned long long* __restrict__ marks,
                   uint32_t* __restrict__ count, uint32_t cap, unsigned long long tag)
{
    // thread t looks at 16 consecutive start offsets
    const unsigned long long base = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) * 16ull;
    if (base + 4 > n) return;
    uint8_t b[19];
#pragma unroll
    for (int k = 0; k < 19; ++k) b[k] = (base + k < n) ? in[base + k] : (uint8_t)0x55;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        if (b[k] == 0 && b[k + 1] == 0 && b[k + 2] == 0xFF && b[k + 3] == 0xFF && base + k + 4 <= n) {
            const uint32_t slot = atomicAdd(count, 1u);
            if (slot < cap) marks[slot] = tag | (base + k + 4);  // first byte after the marker (+ the item's tag)
        }
    }
}

struct ZtsGather {
    unsigned long long src, dst, len;
};

__global__ void __launch_bounds__(256)
segment_gather_kernel(const uint8_t* __restrict__ scratch, uint8_t* __restrict__ out, const ZtsGather* __restrict__ g)
{
    const ZtsGather e = g[blockIdx.x];
    const uint8_t* s = scratch + e.src;
    uint8_t* d = out + e.dst;
    // 16-byte copies where both sides allow it (slots are 16-byte aligned; the destination decides)
    const uint32_t head = (uint32_t)min((unsigned long long)((16u - ((uintptr_t)d & 15u)) & 15u), e.len);
    for (uint32_t i = threadIdx.x; i < head; i += 256) d[i] = s[i];
    const unsigned long long body = (e.len - head) & ~15ull;
    if ((((uintptr_t)(s + head)) & 15u) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(s + head);
        uint4* d4 = reinterpret_cast<uint4*>(d + head);
        for (unsigned long long i = threadIdx.x; i < body / 16; i += 256) d4[i] = s4[i];
    } else {
        for (unsigned long long i = threadIdx.x; i < body; i += 256) d[head + i] = s[head + i];
    }
    for (unsigned long long i = head + body + threadIdx.x; i < e.len; i += 256) d[i] = s[i];
}

static int inflate_launch(zlb_ctx* ctx, cudaStream_t st, const uint8_t* d_in, uint8_t* d_out, const zlb_item* d_items,
                          zlb_result* d_results, size_t n, uint32_t flags);

// Tries to decode the items h_items[big[*]] piecewise, all of them in one pass: one marker scan per item (queued back
// to back, one read-back for all), one segment-mode launch over the pieces of every item, one gather. done[k] tells
// whether big[k] worked (then res[k] is filled); the others have to go through the serial decoder. < 0 on API errors.
#define SPLIT_TAG_SHIFT 40  // marks carry the item's index above the byte offset (offsets < 1 TiB, < 16 M large items)
static int inflate_split_batch(zlb_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const zlb_item* h_items,
                               const std::vector<size_t>& big, uint32_t flags, std::vector<zlb_result>& res,
                               std::vector<char>& done)
{
    const size_t nb = big.size();
    res.assign(nb, zlb_result());
    done.assign(nb, 0);
    for (size_t k = 0; k < nb; ++k) memset(&res[k], 0, sizeof(zlb_result));
    if (nb >= (1ull << 24)) return ZLB_OK;
    cudaStream_t st = ctx->stream;
    int rc = zts_reserve(ctx, &ctx->d_misc, (size_t)SPLIT_MAX_MARKS * 8 + 256);
    if (rc) return rc;
    unsigned long long* d_marks = (unsigned long long*)((uint8_t*)ctx->d_misc.p + 256);
    uint32_t* d_count = (uint32_t*)ctx->d_misc.p;
    ZTS_CUDA(ctx, cudaMemsetAsync(d_count, 0, 4, st));
    for (size_t k = 0; k < nb; ++k) {
        const zlb_item& it = h_items[big[k]];
        if (it.in_len >= (1ull << SPLIT_TAG_SHIFT)) continue;
        const unsigned grid = (unsigned)((it.in_len / 16 + 255) / 256 + 1);
        ZTS_LAUNCH(ctx, ZK_MARKER_SCAN,
                   marker_scan_kernel<<<grid, 256, 0, st>>>(d_in + it.in_off, it.in_len, d_marks, d_count, SPLIT_MAX_MARKS,
                                                           (unsigned long long)k << SPLIT_TAG_SHIFT));
    }
    uint32_t count = 0;
    ZTS_CUDA(ctx, cudaMemcpyAsync(&count, d_count, 4, cudaMemcpyDeviceToHost, st));
    ZTS_CUDA(ctx, cudaStreamSynchronize(st));
    if (count == 0 || count > SPLIT_MAX_MARKS) return ZLB_OK;
    std::vector<unsigned long long> marks(count);
    ZTS_CUDA(ctx, cudaMemcpyAsync(marks.data(), d_marks, (size_t)count * 8, cudaMemcpyDeviceToHost, st));
    ZTS_CUDA(ctx, cudaStreamSynchronize(st));
    std::sort(marks.begin(), marks.end());  // by item, then by offset
    // pieces of every item: it starts at 0 and behind every marker that leaves at least one byte
    struct Piece {
        uint32_t item;  // index into big
        unsigned long long start;
    };
    std::vector<Piece> pieces;
    std::vector<size_t> first(nb + 1, 0);  // pieces of item k: [first[k], first[k + 1])
    {
        size_t mi = 0;
        for (size_t k = 0; k < nb; ++k) {
            first[k] = pieces.size();
            const unsigned long long n = h_items[big[k]].in_len;
            pieces.push_back({(uint32_t)k, 0ull});
            for (; mi < marks.size() && (marks[mi] >> SPLIT_TAG_SHIFT) == k; ++mi) {
                const unsigned long long m = marks[mi] & ((1ull << SPLIT_TAG_SHIFT) - 1);
                // a piece of this engine's streams holds a whole block: never closer than SPLIT_MIN_PIECE bytes to the
                // previous cut (a hostile stream of empty stored blocks would otherwise ask for a slot per 5 bytes)
                if (m < n && m >= pieces.back().start + SPLIT_MIN_PIECE) pieces.push_back({(uint32_t)k, m});
            }
            if (pieces.size() - first[k] < 2) pieces.resize(first[k]);  // no marker inside: nothing to gain
        }
        first[nb] = pieces.size();
    }
    const size_t np = pieces.size();
    if (np == 0) return ZLB_OK;
    // slots: a piece of L compressed bytes cannot inflate to more than ~1032 L (258 bytes per 2 bits at best), and a
    // piece of this engine's streams holds at most one 64 KiB chunk
    std::vector<zlb_item> seg(np);
    size_t scratch = 0, big_in = 0;
    for (size_t k = 0; k < nb; ++k) {
        const zlb_item& it = h_items[big[k]];
        if (first[k + 1] > first[k]) big_in += it.in_len;
        for (size_t j = first[k]; j < first[k + 1]; ++j) {
            seg[j].in_off = it.in_off + pieces[j].start;
            seg[j].in_len = (j + 1 < first[k + 1] ? pieces[j + 1].start : it.in_len) - pieces[j].start;
            size_t slot = (size_t)seg[j].in_len * 1032u + 512u;
            if (slot > SPLIT_SLOT) slot = SPLIT_SLOT;
            slot = (slot + 255) & ~(size_t)255;
            seg[j].out_off = scratch;
            seg[j].out_cap = slot;
            scratch += slot;
        }
    }
    {
        // the scratch is bounded: what the input could legitimately inflate to (2048 x, floor 64 MiB), a quarter of
        // what the device has free right now, 16 GiB at most. Whatever does not fit (or cannot be allocated) is simply
        // not split -- the one-warp decoder gives the same result.
        const size_t need = ((scratch + 255) & ~(size_t)255) + np * (sizeof(zlb_item) + sizeof(zlb_result) + sizeof(ZtsGather)) + 1024;
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = 0;
        size_t budget = (free_b + ctx->d_split.cap) / 4;
        if (budget > ((size_t)16 << 30)) budget = (size_t)16 << 30;
        size_t by_input = big_in * 2048u;
        if (by_input < ((size_t)64 << 20)) by_input = (size_t)64 << 20;
        if (budget > by_input) budget = by_input;
        if (need > budget) return ZLB_OK;
        rc = zts_reserve(ctx, &ctx->d_split, need);
        if (rc == ZLB_E_NOMEM) {
            ctx->err[0] = 0;
            return ZLB_OK;
        }
        if (rc) return rc;
    }
    uint8_t* d_scratch = (uint8_t*)ctx->d_split.p;
    zlb_item* d_seg = (zlb_item*)(d_scratch + ((scratch + 255) & ~(size_t)255));
    zlb_result* d_segres = (zlb_result*)(d_seg + np);
    ZtsGather* d_gather = (ZtsGather*)(d_segres + np);
    ZTS_CUDA(ctx, cudaMemcpyAsync(d_seg, seg.data(), np * sizeof(zlb_item), cudaMemcpyHostToDevice, st));
    ZTS_CUDA(ctx, cudaMemsetAsync(d_segres, 0, np * sizeof(zlb_result), st));
    rc = inflate_launch(ctx, st, d_in, d_scratch, d_seg, d_segres, np,
                        (flags & ZLB_INFLATE_CHECK_NLEN) | ZLB_INFLATE_SEGMENT);
    if (rc) return rc;
    std::vector<zlb_result> sres(np);
    ZTS_CUDA(ctx, cudaMemcpyAsync(sres.data(), d_segres, np * sizeof(zlb_result), cudaMemcpyDeviceToHost, st));
    ZTS_CUDA(ctx, cudaStreamSynchronize(st));
    // validate every item's chain of pieces; the valid ones are gathered in one launch
    std::vector<ZtsGather> gath;
    for (size_t k = 0; k < nb; ++k) {
        const zlb_item& it = h_items[big[k]];
        const size_t a = first[k], b = first[k + 1];
        if (b - a < 2) continue;
        unsigned long long total = 0, blocks = 0;
        bool ok = true;
        for (size_t j = a; j < b && ok; ++j) {
            const zlb_result& r = sres[j];
            const bool last = j + 1 == b;
            const bool fin = (r.blocks & 0x80000000u) != 0;
            if (r.status != ZLB_ST_OK) ok = false;                 // incl. a distance before the piece's first byte
            else if (!last && (fin || r.in_used != seg[j].in_len)) ok = false;
            else if (last && !fin) ok = false;
            total += r.out_len;
            blocks += r.blocks & 0x7FFFFFFFu;
        }
        if (!ok) continue;
        zlb_result& o = res[k];
        o.status = total > it.out_cap ? ZLB_ST_OUT_OVERFLOW : ZLB_ST_OK;
        o.out_len = total > it.out_cap ? 0 : total;
        o.in_used = pieces[b - 1].start + sres[b - 1].in_used;
        o.blocks = (uint32_t)blocks;
        done[k] = 1;
        if (total <= it.out_cap && total) {
            unsigned long long at = 0;
            for (size_t j = a; j < b; ++j) {
                if (sres[j].out_len) gath.push_back({seg[j].out_off, it.out_off + at, sres[j].out_len});
                at += sres[j].out_len;
            }
        }
    }
    if (!gath.empty()) {
        ZTS_CUDA(ctx, cudaMemcpyAsync(d_gather, gath.data(), gath.size() * sizeof(ZtsGather), cudaMemcpyHostToDevice, st));  // <= np entries
        ZTS_LAUNCH(ctx, ZK_GATHER, segment_gather_kernel<<<(unsigned)gath.size(), 256, 0, st>>>(d_scratch, d_out, d_gather));
        ZTS_CUDA(ctx, cudaStreamSynchronize(st));  // gath goes out of scope
    }
    return ZLB_OK;
}

// Host buffers of the "_host" entry point: items are cut into a few waves; the input of a wave travels while the
// previous waves are decoded (one compute stream per wave: a wave is latency-bound, so the waves run side by side) and its output
// travels back as soon as it is done.
struct InfHostIO {
    const uint8_t* h_in;   // where H2D copies read from: the caller's page-locked buffer, or the library's shadow
    uint8_t* h_out;        // where D2H copies write to
    size_t in_bytes, out_bytes;
    ZtsHostStage* stage;   // copy threads behind the shadows (pageable caller buffers), see zts_hoststage.cu
};

static int inflate_device(zlb_ctx* ctx, const uint8_t* d_in, uint8_t* d_out, const zlb_item* h_items,
                          zlb_result* h_results, size_t n, uint32_t flags, const InfHostIO* hio)
{
    int rc = zts_reserve(ctx, &ctx->d_items, n * sizeof(zlb_item) + 64);
    if (rc) return rc;
    rc = zts_reserve(ctx, &ctx->d_results, n * sizeof(zlb_result) + 64);
    if (rc) return rc;
    zlb_item* d_items = (zlb_item*)ctx->d_items.p;
    zlb_result* d_results = (zlb_result*)ctx->d_results.p;
    ZTS_CUDA(ctx, cudaMemcpyAsync(d_items, h_items, n * sizeof(zlb_item), cudaMemcpyHostToDevice, ctx->stream));
    ZTS_CUDA(ctx, cudaMemsetAsync(d_results, 0, n * sizeof(zlb_result), ctx->stream));
    uint32_t kinds = 0;
    if (flags & ZLB_INFLATE_WANT_CRC32) kinds |= ZLB_SUM_CRC32;
    if (flags & ZLB_INFLATE_WANT_ADLER32) kinds |= ZLB_SUM_ADLER32;

    // large items that may be cut at sync-flush markers are decoded piecewise, one after the other (each is
    // massively parallel by itself); whatever is left goes through the batch path below
    if ((flags & ZLB_INFLATE_SPLIT) && !(flags & ZLB_INFLATE_SEGMENT)) {
        std::vector<size_t> big;
        for (size_t i = 0; i < n; ++i)
            if (h_items[i].in_len >= SPLIT_MIN_BYTES) big.push_back(i);
        if (!big.empty()) {
            if (hio) {
                zts_stage_wait_in(hio->stage, 0, hio->in_bytes);
                ZTS_CUDA(ctx, cudaMemcpyAsync((void*)d_in, hio->h_in, hio->in_bytes, cudaMemcpyHostToDevice, ctx->stream));
            }
            std::vector<zlb_item> rest_items;
            std::vector<size_t> rest_idx;
            std::vector<char> is_done(n, 0);
            {
                std::vector<zlb_result> bres;
                std::vector<char> bdone;
                rc = inflate_split_batch(ctx, d_in, d_out, h_items, big, flags, bres, bdone);
                if (rc) return rc;
                bool any = false;
                for (size_t k = 0; k < big.size(); ++k)
                    if (bdone[k]) {
                        is_done[big[k]] = 1;
                        any = true;
                        ZTS_CUDA(ctx, cudaMemcpyAsync(d_results + big[k], &bres[k], sizeof(zlb_result), cudaMemcpyHostToDevice,
                                                      ctx->stream));
                    }
                if (any) ZTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // bres goes out of scope
            }
            for (size_t i = 0; i < n; ++i)
                if (!is_done[i]) {
                    rest_items.push_back(h_items[i]);
                    rest_idx.push_back(i);
                }
            if (!rest_items.empty()) {
                // the remaining items in one batch; their results are scattered back to their slots
                const size_t m = rest_items.size();
                rc = zts_reserve(ctx, &ctx->d_sums, m * (sizeof(zlb_item) + sizeof(zlb_result)) + 256);
                if (rc) return rc;
                zlb_item* d_ri = (zlb_item*)ctx->d_sums.p;
                zlb_result* d_rr = (zlb_result*)(d_ri + m);
                ZTS_CUDA(ctx, cudaMemcpyAsync(d_ri, rest_items.data(), m * sizeof(zlb_item), cudaMemcpyHostToDevice, ctx->stream));
                ZTS_CUDA(ctx, cudaMemsetAsync(d_rr, 0, m * sizeof(zlb_result), ctx->stream));
                rc = inflate_launch(ctx, ctx->stream, d_in, d_out, d_ri, d_rr, m, flags & ~(uint32_t)ZLB_INFLATE_SPLIT);
                if (rc) return rc;
                std::vector<zlb_result> rr(m);
                ZTS_CUDA(ctx, cudaMemcpyAsync(rr.data(), d_rr, m * sizeof(zlb_result), cudaMemcpyDeviceToHost, ctx->stream));
                ZTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
                for (size_t k = 0; k < m; ++k)
                    ZTS_CUDA(ctx, cudaMemcpyAsync(d_results + rest_idx[k], &rr[k], sizeof(zlb_result), cudaMemcpyHostToDevice, ctx->stream));
                ZTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            }
            if (kinds) {
                rc = zts_checksum_device(ctx, d_out, d_items, d_results, h_items, n, kinds, 1);
                if (rc) return rc;
            }
            ZTS_CUDA(ctx, cudaMemcpyAsync(h_results, d_results, n * sizeof(zlb_result), cudaMemcpyDeviceToHost, ctx->stream));
            if (hio) {
                ZTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));  // the sizes decide what travels
                rc = zts_copy_back(ctx, hio->stage, ctx->stream, d_out, hio->h_out, h_items, h_results, d_items, d_results, 0, n);
                if (rc) return rc;
            }
            return ZLB_OK;
        }
    }
    // waves only pay when the items are laid out in order on both sides (then a wave is one contiguous copy each way)
    size_t n_waves = 1;
    if (hio && n >= 1024) {
        n_waves = n >= 32768 ? 16 : n >= 16384 ? 8 : 4;  // (the last wave's output travels alone: keep it small)
        for (size_t i = 1; i < n && n_waves > 1; ++i)
            if (h_items[i].in_off < h_items[i - 1].in_off + h_items[i - 1].in_len ||
                h_items[i].out_off < h_items[i - 1].out_off + h_items[i - 1].out_cap)
                n_waves = 1;
    }
    if (!hio) {
        rc = inflate_launch(ctx, ctx->stream, d_in, d_out, d_items, d_results, n, flags);
        if (rc) return rc;
        if (kinds) {
            rc = zts_checksum_device(ctx, d_out, d_items, d_results, h_items, n, kinds, 1);
            if (rc) return rc;
        }
        ZTS_CUDA(ctx, cudaMemcpyAsync(h_results, d_results, n * sizeof(zlb_result), cudaMemcpyDeviceToHost, ctx->stream));
        return ZLB_OK;
    }

    // Host buffers: the input of wave k+1 travels on a copy stream while wave k is decoded, and as soon as the sizes of
    // wave k are known (its results are read back behind it, wave k+1 is already queued) what it wrote travels back on
    // the other copy stream. The waves are launched on up to four compute streams in turn: a stream is decoded by one
    // warp from start to end (milliseconds, whatever the batch size), so waves that do not fill the device by
    // themselves run side by side, staggered by their input copies, and their outputs leave as each one finishes;
    // waves that do fill it simply take over the SMs as their predecessor drains.
    rc = zts_host_streams(ctx);
    if (rc) return rc;
    rc = zts_reserve_pinned2(ctx, n * sizeof(zlb_result));
    if (rc) return rc;
    zlb_result* pin_res = (zlb_result*)ctx->h_pin2;
    const size_t per = (n + n_waves - 1) / n_waves;
    // events: 0 = item table on the device; 1 + 2k = input of wave k arrived; 2 + 2k = wave k decoded, results read back
    cudaEvent_t ev_tab = zts_sync_event(ctx, 0);
    ZTS_CUDA(ctx, cudaEventRecord(ev_tab, ctx->stream));
    auto wave_range = [&](size_t k, size_t& a, size_t& b) {
        a = k * per;
        b = a + per < n ? a + per : n;
        if (a > n) a = n;
    };
    auto wave_in = [&](size_t k) -> int {
        size_t a, b;
        wave_range(k, a, b);
        if (a < b) {
            uint64_t ilo, ihi;
            if (n_waves == 1) {  // items in any order: the whole blob
                ilo = 0;
                ihi = hio->in_bytes;
            } else {
                ilo = h_items[a].in_off;
                ihi = h_items[b - 1].in_off + h_items[b - 1].in_len;
            }
            if (ihi > ilo) {
                zts_stage_wait_in(hio->stage, ilo, ihi);
                ZTS_CUDA(ctx, cudaMemcpyAsync((void*)(d_in + ilo), hio->h_in + ilo, ihi - ilo, cudaMemcpyHostToDevice, ctx->s_in));
            }
        }
        ZTS_CUDA(ctx, cudaEventRecord(zts_sync_event(ctx, 1 + 2 * k), ctx->s_in));
        return ZLB_OK;
    };
    const bool trace = getenv("ZTS_TRACE_WAVES") != nullptr;  // development aid: when does the host see each wave
    const auto t_begin = std::chrono::steady_clock::now();
    auto wave_out = [&](size_t k) -> int {
        size_t a, b;
        wave_range(k, a, b);
        ZTS_CUDA(ctx, cudaEventSynchronize(zts_sync_event(ctx, 2 + 2 * k)));
        if (trace)
            fprintf(stderr, "inflate wave %zu of %zu decoded at %.2f ms\n", k, n_waves,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
        if (a >= b) return ZLB_OK;
        // The results are read back only now, on a stream of their own: a copy queued behind a kernel that is still
        // running holds up every later copy of its direction (the copy engine serves its queue in order), and the
        // outputs of the earlier waves would wait for the last wave's kernel.
        ZTS_CUDA(ctx, cudaMemcpyAsync(pin_res + a, d_results + a, (b - a) * sizeof(zlb_result), cudaMemcpyDeviceToHost, ctx->s_res));
        ZTS_CUDA(ctx, cudaStreamSynchronize(ctx->s_res));
        memcpy(h_results + a, pin_res + a, (b - a) * sizeof(zlb_result));
        return zts_copy_back(ctx, hio->stage, ctx->s_out, d_out, hio->h_out, h_items, h_results, d_items, d_results, a, b);
    };
    if ((rc = wave_in(0))) return rc;
    for (size_t k = 0; k < n_waves; ++k) {
        size_t a, b;
        wave_range(k, a, b);
        // checksums are computed on the context's stream (their tables are staged there): then every wave runs on it
        cudaStream_t st = (kinds || k % 4 == 0) ? ctx->stream : ctx->s_aux[k % 4 - 1];
        if (st != ctx->stream && k < 4) ZTS_CUDA(ctx, cudaStreamWaitEvent(st, ev_tab, 0));  // the item table
        ZTS_CUDA(ctx, cudaStreamWaitEvent(st, zts_sync_event(ctx, 1 + 2 * k), 0));
        if (a < b) {
            ctx->work = st;
            rc = inflate_launch(ctx, st, d_in, d_out, d_items + a, d_results + a, b - a, flags);
            ctx->work = ctx->stream;
            if (rc) return rc;
            if (kinds) {
                rc = zts_checksum_device(ctx, d_out, d_items + a, d_results + a, h_items + a, b - a, kinds, 1);
                if (rc) return rc;
            }
        }
        ZTS_CUDA(ctx, cudaEventRecord(zts_sync_event(ctx, 2 + 2 * k), st));
        if (k + 1 < n_waves && (rc = wave_in(k + 1))) return rc;  // behind the launches: may wait for the copy threads
    }
    // the call ends on ctx->stream -- joined only here: a join inside the loop would make the next wave queued on
    // ctx->stream wait for the waves in flight on the other streams
    for (size_t k = 0; k < n_waves; ++k)
        if (!(kinds || k % 4 == 0)) ZTS_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, zts_sync_event(ctx, 2 + 2 * k), 0));
    // every wave is queued (none waits for the host): their outputs leave in order, each as soon as its wave is done
    for (size_t k = 0; k < n_waves; ++k)
        if ((rc = wave_out(k))) return rc;
    ZTS_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
    if (trace)
        fprintf(stderr, "inflate outputs back at %.2f ms\n",
                std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count());
    return ZLB_OK;
}

extern "C" int zlb_inflate_batch(zlb_ctx* ctx, const void* d_in, void* d_out, const zlb_item* items,
                                 zlb_result* results, size_t n, uint32_t flags)
{
    if (!ctx || (!items && n) || (!results && n) || (!d_in && n) || (!d_out && n)) return ZLB_E_ARG;
    if (n == 0) return ZLB_OK;
    if (n > 0x7FFFFFFFull) return zts_fail(ctx, ZLB_E_ARG, "too many items");
    ZTS_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc = inflate_device(ctx, (const uint8_t*)d_in, (uint8_t*)d_out, items, results, n, flags, nullptr);
    if (rc) return rc;
    ZTS_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return ZLB_OK;
}

extern "C" int zlb_inflate_batch_host(zlb_ctx* ctx, const void* h_in, size_t in_bytes, void* h_out, size_t out_bytes,
                                      const zlb_item* items, zlb_result* results, size_t n, uint32_t flags)
{
    if (!ctx || (!items && n) || (!results && n) || (!h_in && in_bytes) || (!h_out && out_bytes)) return ZLB_E_ARG;
    if (n == 0) return ZLB_OK;
    if (n > 0x7FFFFFFFull) return zts_fail(ctx, ZLB_E_ARG, "too many items");
    ZTS_CUDA(ctx, cudaSetDevice(ctx->device));
    for (size_t i = 0; i < n; ++i) {
        if (items[i].in_off + items[i].in_len > in_bytes || items[i].out_off + items[i].out_cap > out_bytes)
            return zts_fail(ctx, ZLB_E_ARG, "item %zu out of range", i);
    }
    int rc = zts_reserve(ctx, &ctx->d_stage_in, in_bytes + 256);
    if (rc) return rc;
    rc = zts_reserve(ctx, &ctx->d_stage_out, out_bytes + 256);
    if (rc) return rc;
    ZtsHostStage* stage = nullptr;
    rc = zts_stage_begin(ctx, h_in, in_bytes, h_out, out_bytes, &stage);
    if (!rc) {
        InfHostIO hio = {zts_stage_in_ptr(stage), zts_stage_out_ptr(stage), in_bytes, out_bytes, stage};
        rc = inflate_device(ctx, (const uint8_t*)ctx->d_stage_in.p, (uint8_t*)ctx->d_stage_out.p, items, results, n, flags,
                            &hio);
        if (!rc && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = zts_fail(ctx, ZLB_E_CUDA, "synchronize failed");
    }
    zts_stage_end(ctx, stage);  // also on errors: the copy threads hold pointers into the caller's buffers
    return rc;
}
