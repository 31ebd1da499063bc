// zts_lz77.cu -- exact LZ77 tokenisation of one chunk per CTA
// (replaces LZ77.encode / searchLongestMatch / maxMatchTest, src/LZ77.ts:149-283, lazy = 0).
//
// What the reference computes: at every parse position p (p + 3 < n) take ALL
// earlier positions q with the same 3 bytes and p - q <= 32768, the longest match wins (<= 258,
// <= n - p), ties go to the nearest q; greedy: p += len, else literal. The reference walks
// per-key JS arrays newest-first; here the same function is evaluated as:
//
//   1. the chunk is staged in shared memory by a TMA bulk copy (cp.async.bulk + mbarrier);
//   2. all positions are radix-sorted (stable, 2 LSD passes of 8 bits, ballot-based ranking inside a warp)
//      by a 16-bit hash of their 3-byte key -> per-bucket position lists, ascending, contiguous; there is no table of
//      bucket starts;
//   3. one pass over the sorted index: a slot whose hash differs from its predecessor's starts a bucket, a position's
//      rank is its distance to the last start; every position gets an info word (L2-resident scratch): its slot and
//      its rank (how many earlier positions the bucket holds), and one bit in shared memory, "may have a candidate"
//      -- an earlier position with the same 3 bytes inside the window, found by walking back over the few hash
//      collisions in front of the position's slot;
//   4a. the chunk is cut into 1024 tiles of 64 positions and EVERY THREAD OWNS ONE: it parses its tile speculatively
//      from the tile's start, all by itself -- runs of positions without a candidate are literals and are emitted
//      together; at a position with candidates the thread walks the `rank` slots in front of the position's own slot,
//      newest first: one candidate per step with the exact comparison until a match is known, then four per step
//      through the tail-byte test (only a strictly longer match counts) with the exact comparison for the nearest
//      that passes; strictly longer wins, so ties stay with the nearest. A warp iterates "candidate steps for the
//      lanes that are searching, then one token / next-position step for the lanes that are not", so lanes with short
//      and long candidate lists do not wait for each other. Positions behind long candidate lists (ends of runs,
//      low-entropy records) are handed to the whole warp, 32 (then 128) candidates per step. A bounded search depth
//      (fast mode) and one-step lazy evaluation (ZLB_MODE_LAZY) are options of the same loop;
//   4b. the greedy parse is memoryless in p, so the true parse re-enters each tile at the true exit of the previous
//      one and only has to be re-parsed until it meets a speculatively parsed position (a visited-bit per
//      position); the remainder of the speculative tokens is reused. The owner of tile t does this for tile t + 1
//      as soon as its own tile is done, assuming tile t exits where its speculative parse did; the few tiles whose
//      predecessor exits elsewhere (a match that jumps a whole tile, chains inside long runs) are redone chain by
//      chain, one warp per chain, and a final check over the true exits verifies every splice;
//   5. the surviving tokens are copied into the chunk's contiguous token list and histogrammed (litlen / dist).
//
// Shared memory: 64 KiB chunk + 128 KiB sorted positions + 16 KiB radix offsets / visited bits + tile table, 8 KiB
// candidate bits, 8 KiB tile tables and counters.
#include <stddef.h>

#include "zts_deflate.cuh"

typedef uint8_t LZ_COUNT_T;    // tokens of a 64-position tile

struct LzSmem {
    // byte offsets into dynamic shared memory
    static constexpr uint32_t S_OFF = 0;                          // chunk bytes (+ shift, + slack)
    static constexpr uint32_t S_BYTES = LZ_MAX_CHUNK + 64;
    static constexpr uint32_t SORTED_OFF = S_OFF + S_BYTES;       // u16[65536]
    static constexpr uint32_t SORTED_BYTES = LZ_MAX_CHUNK * 2;
    static constexpr uint32_t BSTART_OFF = SORTED_OFF + SORTED_BYTES;  // 16 KiB: radix running offsets | visited bits + tile table
    static constexpr uint32_t BSTART_BYTES = 16384 + 16;
    static constexpr uint32_t AUX_OFF = BSTART_OFF + BSTART_BYTES;     // 8 KiB: may-have-a-candidate bits
    static constexpr uint32_t AUX_BYTES = 8192;
    static constexpr uint32_t MISC_OFF = AUX_OFF + AUX_BYTES;
    static constexpr uint32_t MISC_BYTES = 8192;
    static constexpr uint32_t TOTAL = MISC_OFF + MISC_BYTES;
};

// Per-tile tables. Positions are kept relative to the tile's first position (an exit lies at most a maximum match
// behind the tile: < 64 + 258; an entry at most 257 behind its first position), counts fit a byte (<= 64 tokens per
// tile). LzMisc lives in the misc area, LzTiles2 in the second half of the bucket-start area (dead once the info
// words are built).
struct LzMisc {
    unsigned long long mbar;
    uint32_t chunk;
    uint32_t n_tokens;
    uint32_t warp_tot[32];
    uint32_t warp_min[32];
    uint32_t hist[316];
    uint32_t start_mask[(LZ_NTILES + 31) / 32];  // tiles that start a chain of wrongly entered tiles
    uint16_t tile_start[LZ_NTILES];  // where the tile's speculative parse starts
    uint16_t entry_used[LZ_NTILES];  // the entry the tile was last parsed from
    LZ_COUNT_T spec_count[LZ_NTILES];
    LZ_COUNT_T fix_count[LZ_NTILES];
};
struct LzTiles2 {
    uint16_t spec_exit[LZ_NTILES];   // where the speculative parse of a tile ended (>= tile end)   } the two arrays double as
    uint16_t fix_exit[LZ_NTILES];    // exit of the tile for the entry it was last parsed from      } u32 tok_off[] in phase 5
    LZ_COUNT_T spec_from[LZ_NTILES]; // first speculative token that belongs to the true parse (4a: the position of the splice)
};
#define LZ_NO_SPLICE 0xFFu          // LzTiles2::spec_from during 4a: the re-entry parse did not meet the speculative one
#define LZ_EXIT_SPLICED 0xFFFFu     // LzTiles2::fix_exit during 4a: the tile exits where its speculative parse does

// The radix scratch T is the one global buffer a CTA keeps re-using (256 KiB per chunk, written and read twice):
// it is marked evict_last in L2 so that the streaming traffic next to it (input, tokens) does not push it out to
// DRAM, and the streaming token stores are marked evict_first.
__device__ __forceinline__ unsigned long long l2_policy_keep()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ unsigned long long l2_policy_stream()
{
    unsigned long long pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void st_u32_hint(uint32_t* p, uint32_t v, unsigned long long pol)
{
    asm volatile("st.global.cg.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
__device__ __forceinline__ uint32_t ld_u32_hint(const uint32_t* p, unsigned long long pol)
{
    uint32_t v;
    asm volatile("ld.global.cg.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// 4 bytes at byte offset i of S. The loads go through the 32-bit shared-memory address of S: the two aligned
// words around S + i and a funnel shift whose amount is the address itself times 8 (SHF takes it mod 32, i.e.
// the misalignment in bits) -- six instructions, none of which depends on where the chunk sits in the buffer.
// >= 8 readable bytes follow any i <= n.
struct LzS {
    const uint8_t* S;
    uint32_t sa;  // shared-space address of S
    __device__ __forceinline__ uint8_t operator[](uint32_t i) const { return S[i]; }
    // the same through the 32-bit shared-window address: in the warp search a load through the generic pointer makes
    // ptxas rebuild the window base (S2UR, UMOV, ULEA, two adds) in front of every byte it reads
    __device__ __forceinline__ uint8_t at(uint32_t i) const
    {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(sa + i));
        return (uint8_t)v;
    }
};
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;  // volatile: keeps its place relative to the barriers that separate staging from reading
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ uint32_t ld_u32(const LzS& V, uint32_t i)
{
    const uint32_t a = V.sa + i, w = a & ~3u;
    return __funnelshift_r(lds_u32(w), lds_u32(w + 4u), a << 3);
}

// 16 bytes at byte offset i of S as four little-endian words (>= 24 readable bytes follow any i <= n)
__device__ __forceinline__ void ld_u128(const LzS& V, uint32_t i, uint32_t (&o)[4])
{
    const uint32_t a = V.sa + i, w = a & ~3u, sh = a << 3;
    const uint32_t w0 = lds_u32(w), w1 = lds_u32(w + 4u), w2 = lds_u32(w + 8u), w3 = lds_u32(w + 12u), w4 = lds_u32(w + 16u);
    o[0] = __funnelshift_r(w0, w1, sh);
    o[1] = __funnelshift_r(w1, w2, sh);
    o[2] = __funnelshift_r(w2, w3, sh);
    o[3] = __funnelshift_r(w3, w4, sh);
}

__device__ __forceinline__ uint32_t lz_hash(uint32_t key3) { return (key3 * 0x9E3779B1u) >> (32 - LZ_HASH_BITS); }

// lanes holding the same NBITS-bit digit as this lane (invalid lanes only match each other). Same result as
// __match_any_sync, built from NBITS + 1 ballots: MATCH.ANY issues far too slowly on sm_100 to sit in a loop that
// runs once per 32 positions (ncu: the instruction behind it carried a fifth of the kernel's stall samples).
template <int NBITS>
__device__ __forceinline__ unsigned peers_of(uint32_t d, bool valid)
{
#ifdef ZTS_USE_MATCH
    return __match_any_sync(0xFFFFFFFFu, valid ? d : 0xFFFFFFFFu);
#else
    unsigned m = __ballot_sync(0xFFFFFFFFu, valid);
    unsigned peers = valid ? m : ~m;
#pragma unroll
    for (int b = 0; b < NBITS; ++b) {
        // test, vote, select, AND-XOR: the C++ form of this costs six instructions per bit
        asm volatile(
            "{\n\t.reg .pred p;\n\t.reg .b32 t, m, s;\n\t"
            "and.b32 t, %1, %2;\n\t"
            "setp.ne.u32 p, t, 0;\n\t"
            "vote.sync.ballot.b32 m, p, 0xffffffff;\n\t"
            "selp.b32 s, 0, 0xffffffff, p;\n\t"
            "lop3.b32 %0, %0, m, s, 0x60;\n\t}"   // peers & (m ^ s): m where the bit is set, ~m where it is not
            : "+r"(peers)
            : "r"(d), "r"(1u << b));
    }
    return peers;
#endif
}

// exclusive block scan (sum) over 1024 threads
__device__ __forceinline__ uint32_t block_excl_sum(uint32_t v, uint32_t* warp_tot, uint32_t* total)
{
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, inc, d);
        if (lane >= (unsigned)d) inc += t;
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_tot[lane], winc = w;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, winc, d);
            if (lane >= (unsigned)d) winc += t;
        }
        warp_tot[lane] = winc - w;
        if (lane == 31 && total) *total = winc;
    }
    __syncthreads();
    uint32_t r = warp_tot[warp] + inc - v;
    __syncthreads();
    return r;
}

// length of the match between candidate q and position p given that both share the 3-byte key test below;
// 0 when q has another key. pw / pw1 = the 8 bytes at p.
__device__ __forceinline__ uint32_t lz_match_len(const LzS& S, uint32_t q, uint32_t p, uint32_t pw, uint32_t pw1,
                                                 uint32_t maxlen)
{
    const uint32_t a = S.sa + q, w = a & ~3u, sh = a << 3;  // ld_u32(S, q) and ld_u32(S, q + 4) share a word
    const uint32_t w1 = lds_u32(w + 4u);
    const uint32_t x0 = __funnelshift_r(lds_u32(w), w1, sh) ^ pw;
    if (x0 & 0xFFFFFFu) return 0;  // another table[] key (src/LZ77.ts:204-214)
    uint32_t k = 3;
    if (x0 == 0) {
        k = 4;
        // most matches end within the next word; long ones continue 16 bytes per step
        const uint32_t x1 = __funnelshift_r(w1, lds_u32(w + 8u), sh) ^ pw1;
        if (x1) {
            k += (uint32_t)(__ffs((int)x1) - 1) >> 3;
        } else {
            k = 8;
            while (k < maxlen) {
                uint32_t a4[4], b4[4];
                ld_u128(S, q + k, a4);
                ld_u128(S, p + k, b4);
                const uint32_t y0 = a4[0] ^ b4[0], y1 = a4[1] ^ b4[1], y2 = a4[2] ^ b4[2], y3 = a4[3] ^ b4[3];
                if (y0 | y1 | y2 | y3) {
                    const uint32_t y = y0 ? y0 : y1 ? y1 : y2 ? y2 : y3;
                    const uint32_t wsel = y0 ? 0u : y1 ? 4u : y2 ? 8u : 12u;
                    k += wsel + ((uint32_t)(__ffs((int)y) - 1) >> 3);
                    break;
                }
                k += 16;
            }
        }
    }
    return min(k, maxlen);
}

// ---- per-position info ------------------------------------------------------------------------------------------
// Built once per chunk from the sorted index, one thread per slot:
//   P[p]      = slot of p in the sorted index | rank of p inside its bucket << 16  (rank = earlier bucket entries; the
//               candidates of p are the `rank` slots in front of its own). L2-resident scratch, 4 bytes per position.
//   hasbits   = one bit per position in shared memory: "may have a candidate" -- an earlier position with the same
//               3 bytes inside the window, found by walking back over the hash collisions in front of the slot.
#ifndef LZ_HAS_WALK
#define LZ_HAS_WALK 16u        // collisions walked over before a position is declared "may have a candidate"
#endif
#ifndef LZ_PRIV_CAP
#define LZ_PRIV_CAP 64u        // a search over more earlier bucket entries than this is handed to the whole warp
#endif
#ifndef LZ_WAIT_MUL
#define LZ_WAIT_MUL 0u         // ... or, if not 0, as soon as LZ_WAIT_MUL x (lanes waiting) >= lanes searching (measured: 0 is best)
#endif
#ifndef LZ_COOP_PER_LANE
#define LZ_COOP_PER_LANE 8u    // candidates per lane and step of the warp search's tail-byte filter
#endif
#ifndef LZ_KMAX
#define LZ_KMAX 8u             // candidate steps in a row before the lanes that wait for their next position are served
#endif

// Every warp walks the 2048 slots of the sorted index it ranked in the last radix sweep, 32 per step. Bucket starts
// are the slots whose hash differs from their predecessor's; a position's rank is its distance to the last start (a
// ballot + the start carried from the earlier steps) -- no table of bucket starts exists.
// P[] and `hasbits` (zeroed by the caller, barrier behind it) as described above.
__device__ __forceinline__ void lz_build_info(const LzS& S, const uint16_t* __restrict__ sorted, uint32_t m, uint32_t n,
                                              uint32_t* P, uint32_t* hasbits, unsigned long long keep)
{
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (uint32_t p = m + tid; p < n; p += LZ_THREADS) st_u32_hint(&P[p], 0u, keep);  // positions without a slot (P has LZ_MAX_CHUNK entries per CTA)
    const uint32_t i0 = warp * LZ_SORT_TILE;
    if (i0 >= m) return;
    // the bucket that holds slot i0 may begin in an earlier warp's range: look back for its first slot
    uint32_t carry = i0;          // first slot of the bucket that is open at the current step
    uint32_t hlast = 0xFFFFFFFFu; // hash of the slot in front of the current step
    if (i0) {
        const uint32_t h0 = lz_hash(ld_u32(S, sorted[i0]) & 0xFFFFFFu);
        hlast = lz_hash(ld_u32(S, sorted[i0 - 1]) & 0xFFFFFFu);
        if (hlast == h0) {
            uint32_t at = i0;  // slots [at, i0) are known to share h0
            for (;;) {
                const bool in = at > lane;  // slot at - 1 - lane exists
                const uint32_t hl = in ? lz_hash(ld_u32(S, sorted[at - 1 - lane]) & 0xFFFFFFu) : 0xFFFFFFFFu;
                const unsigned mism = __ballot_sync(0xFFFFFFFFu, hl != h0);
                if (mism) {
                    at -= (uint32_t)__ffs((int)mism) - 1u;
                    break;
                }
                at -= 32u;
            }
            carry = at;
        }
    }
    // the position of a slot is fetched two steps ahead, its key bytes one step ahead (two dependent shared-memory
    // round trips that would otherwise stand in front of every step)
    const uint32_t NOPOS = 0xFFFFFFFFu;
    uint32_t p_n1 = i0 + lane < m ? (uint32_t)sorted[i0 + lane] : NOPOS;
    uint32_t p_n2 = i0 + 32u + lane < m ? (uint32_t)sorted[i0 + 32u + lane] : NOPOS;
    uint32_t key_n1 = p_n1 != NOPOS ? ld_u32(S, p_n1) & 0xFFFFFFu : 0u;
    for (uint32_t it = 0; it < LZ_SORT_TILE / 32; ++it) {  // whole warps stay in the loop (warp votes below)
        const uint32_t i = i0 + it * 32 + lane;
        const bool valid = i < m;
        const uint32_t p = p_n1, key = key_n1;
        p_n1 = p_n2;
        key_n1 = p_n1 != NOPOS ? ld_u32(S, p_n1) & 0xFFFFFFu : 0u;
        p_n2 = (it + 2u < LZ_SORT_TILE / 32 && i + 64u < m) ? (uint32_t)sorted[i + 64u] : NOPOS;
        const uint32_t h = valid ? lz_hash(key) : 0xFFFFFFFEu;
        uint32_t hprev = __shfl_up_sync(0xFFFFFFFFu, h, 1);
        if (lane == 0) hprev = hlast;
        hlast = __shfl_sync(0xFFFFFFFFu, h, 31);
        const bool start = valid && (i == 0 || h != hprev);
        const unsigned sm = __ballot_sync(0xFFFFFFFFu, start);
        const unsigned below = sm & (zts_lanemask_lt() | (1u << lane));
        const uint32_t lo = below ? (i - lane) + (31u - (uint32_t)__clz((int)below)) : carry;  // first slot of i's bucket
        if (sm) carry = (i - lane) + (31u - (uint32_t)__clz((int)sm));
        bool has = false;
        if (valid) {
            if (p + 3u < n) {  // src/LZ77.ts:228: the last three positions are never searched
                // the nearest earlier position with the same 3 bytes sits a few slots back (hash collisions in
                // between); it decides: older ones are further away
                uint32_t j = i, steps = 0;
                while (j > lo) {
                    --j;
                    const uint32_t q = sorted[j];
                    if (S[q] == (uint8_t)key && (ld_u32(S, q) & 0xFFFFFFu) == key) {  // one byte decides for most collisions
                        has = p - q <= LZ_WINDOW;
                        break;
                    }
                    if (++steps >= LZ_HAS_WALK) {
                        has = true;  // undecided: the search will tell
                        break;
                    }
                }
            }
            ZTS_ASSERT(p < n && lo <= i && i < m);
            if (has) st_u32_hint(&P[p], i | ((i - lo) << 16), keep);  // (only read where the bit is set; a scattered store)
        }
        // inside a long bucket consecutive slots are consecutive positions (runs of one byte): one atomic per warp then
        const uint32_t w = p >> 5, w0 = __shfl_sync(0xFFFFFFFFu, w, 0);
        const uint32_t bit = has ? 1u << (p & 31u) : 0u;
        if (__all_sync(0xFFFFFFFFu, w == w0)) {
            const uint32_t bits = __reduce_or_sync(0xFFFFFFFFu, bit);
            if (lane == 0 && bits) atomicOr(&hasbits[w0], bits);
        } else if (bit) {
            atomicOr(&hasbits[w], bit);
        }
    }
}

// ---- warp-cooperative longest/nearest match search at position p (requires p + 3 < n) ----------
// returns (len << 16) | dist, or 0 when no candidate exists (src/LZ77.ts:157-194 + :242)
// The candidates are the slots [lo, cur) of the sorted index, newest (cur - 1) first;
// `depth` = how many of them are looked at: 0xFFFFFFFF = all, like the reference.
__device__ __forceinline__ uint32_t lz_search_from(const LzS& S, const uint16_t* __restrict__ sorted, uint32_t lo,
                                                   uint32_t cur, uint32_t p, uint32_t n, uint32_t depth)
{
    const unsigned lane = zts_lane();
    const uint32_t pw = ld_u32(S, p);
    const uint32_t maxlen = min(LZ_MAXLEN, n - p);
    // inside a run of one byte the distance-1 candidate reaches the maximum length: nothing is nearer and nothing
    // is longer, so that is the reference's answer (it stops at the first candidate there, :189 / end of input)
    {
        const uint32_t b0 = pw & 0xFFu;
        if (((pw ^ (pw >> 8)) & 0xFFFFu) == 0 && p > 0 && S.at(p - 1) == b0) {  // bytes 0..2 equal, and the one before
            const uint32_t splat = b0 * 0x01010101u;
            bool same = true;
#pragma unroll
            for (uint32_t w = 0; w < 3; ++w) {  // lane j checks bytes [12 j, 12 j + 12) of the 258 ahead
                const uint32_t k0 = lane * 12u + w * 4u;
                if (k0 < maxlen) {
                    uint32_t x = ld_u32(S, p + k0) ^ splat;
                    const uint32_t nbytes = maxlen - k0;
                    if (nbytes < 4u) x &= (1u << (8u * nbytes)) - 1u;
                    same = same && x == 0;
                }
            }
            if (__all_sync(0xFFFFFFFFu, same)) return (maxlen << 16) | 1u;
        }
    }
    const uint32_t pw1 = ld_u32(S, p + 4);
    uint32_t best = 0, best_len = 0;
    if (depth < cur - lo) lo = cur - depth;  // candidate budget
    // Positions ascend with the slot, so the candidates inside the window are the upper part of [lo, cur): it is cut
    // off once (binary search, every lane the same), and no step has to look at distances again (src/LZ77.ts:223)
    if (cur > lo && p - sorted[lo] > LZ_WINDOW) {
        uint32_t a = lo, b = cur;  // first slot inside the window lies in (a, b]
        while (b - a > 1u) {
            const uint32_t mid = (a + b) >> 1;
            if (p - sorted[mid] > LZ_WINDOW)
                a = mid;
            else
                b = mid;
        }
        lo = b;
    }
    // until a match is known: 32 candidates per step, exact comparison
    while (cur > lo && best_len < 3u) {
        const uint32_t cnt = min(32u, cur - lo);
        ZTS_ASSERT(cur >= cnt && cur <= LZ_MAX_CHUNK);
        uint32_t key = 0;
        if (lane < cnt) {
            const uint32_t q = sorted[cur - 1 - lane];  // lane 0 = newest candidate
            key = (lz_match_len(S, q, p, pw, pw1, maxlen) << 16) | q;
        }
        const uint32_t m = __reduce_max_sync(0xFFFFFFFFu, key);  // longest, then nearest
        if (m >> 16) {
            best_len = m >> 16;
            best = m;
        }
        cur -= cnt;
    }
    // behind a known match only a strictly longer one counts (:183): its byte at best_len must match, which is the
    // cheapest test and rejects nearly everything. 256 candidates per step through that test (eight independent
    // look-ups per lane in flight); the few that pass get the exact comparison
    while (cur > lo && best_len < maxlen) {  // :189 (258) or capped by the input end
        const uint32_t cnt = min(32u * LZ_COOP_PER_LANE, cur - lo);
        const uint8_t pt = S.at(p + best_len);
        uint32_t key = 0;
        unsigned hm = 0;  // which of this lane's candidates passed the test
        const uint16_t* sp = sorted + cur - 1u - lane;  // candidate k of this lane: sp[-32 k]
#pragma unroll
        for (uint32_t k = 0; k < LZ_COOP_PER_LANE; ++k) {
            const uint32_t o = k * 32u + lane;
            if (o < cnt) {
                ZTS_ASSERT(cur >= 1u + o && cur <= LZ_MAX_CHUNK);
                const uint32_t q = sp[-32 * (int)k];
                ZTS_ASSERT(q < p && p - q <= LZ_WINDOW);
                if (S.at(q + best_len) == pt) hm |= 1u << k;
            }
        }
        while (hm) {  // (rare: the position is read again)
            const uint32_t k = (uint32_t)__ffs((int)hm) - 1u;
            hm &= hm - 1u;
            const uint32_t q = sp[-32 * (int)k];
            key = max(key, (lz_match_len(S, q, p, pw, pw1, maxlen) << 16) | q);
        }
        if (__any_sync(0xFFFFFFFFu, key != 0u)) {
            const uint32_t m = __reduce_max_sync(0xFFFFFFFFu, key);
            if ((m >> 16) > best_len) {
                best_len = m >> 16;
                best = m;
            }
        }
        cur -= cnt;
    }
    if (best_len < 3) return 0;
    return (best_len << 16) | (p - (best & 0xFFFFu));
}

// one greedy step of the whole warp at parse position p: emits the token, returns the next parse position
__device__ __forceinline__ uint32_t lz_step(const LzS& S, const uint16_t* sorted, const uint32_t* P,
                                            const uint32_t* hasbits, unsigned long long keep, uint32_t p, uint32_t n,
                                            uint32_t depth, uint32_t* tok_out)
{
    uint32_t r = 0;
    if ((hasbits[p >> 5] >> (p & 31u)) & 1u) {  // never set for the last three positions (src/LZ77.ts:228)
        const uint32_t info = ld_u32_hint(&P[p], keep);
        const uint32_t slot = info & 0xFFFFu;
        r = lz_search_from(S, sorted, slot - (info >> 16), slot, p, n, depth);
    }
    if (r) {
        const uint32_t len = r >> 16, dist = r & 0xFFFFu;
        *tok_out = TOK_MATCH | ((len - 3) << 16) | (dist - 1);
        return p + len;
    }
    *tok_out = S[p];
    return p + 1;
}

// Where should the speculative parse of a tile that begins inside a run of one byte start? Inside such a run every
// match is a maximum-length distance-1 match (lz_search's first branch), so the true parse visits run_start + 1 +
// 258 k when it entered the run with a literal at its first byte -- the usual case. Starting the tile's speculative
// parse at the first such position >= t_begin makes the predecessor's exit land on it, and no chain of wrongly
// entered tiles forms across the run. Purely a heuristic choice of the start: any start >= t_begin is a valid
// speculative parse, the re-entry / chain / verification passes do not care where it came from.
// Returns t_begin when the tile does not begin inside a run, the run's first byte is out of reach (LZ_RUN_LOOKBACK),
// or the run ends before the aligned position.
#define LZ_LAZY_MAX 32u        // lazy evaluation (ZLB_MODE_LAZY): matches at least this long are taken at once
#define LZ_RUN_LOOKBACK 8192u
__device__ __forceinline__ uint32_t lz_run_aligned_start(const LzS& S, uint32_t t_begin, uint32_t t_end, uint32_t n)
{
    const unsigned lane = zts_lane();
    if (t_begin < 4u || t_begin + LZ_MAXLEN + 4u > n) return t_begin;
    const uint32_t w = ld_u32(S, t_begin - 1u);  // bytes t_begin - 1 .. t_begin + 2
    const uint32_t splat = (w & 0xFFu) * 0x01010101u;
    if (w != splat) return t_begin;
    // first byte of the run: look back 128 bytes per step (lane l: the word at pos - 4 (l + 1))
    uint32_t pos = t_begin - 1u;  // S[pos] is in the run
    const uint32_t lo = t_begin > LZ_RUN_LOOKBACK ? t_begin - LZ_RUN_LOOKBACK : 0u;
    uint32_t r0 = 0xFFFFFFFFu;
    while (pos > lo) {
        const uint32_t at = pos - 4u * (lane + 1u);  // wraps below zero for lanes past the start of the buffer
        const bool in = pos >= 4u * (lane + 1u) && at >= lo;
        const uint32_t x = in ? (ld_u32(S, at) ^ splat) : 0xFFFFFFFFu;  // out of reach counts as a mismatch
        const unsigned mm = __ballot_sync(0xFFFFFFFFu, x != 0u);
        if (mm) {
            const int src = __ffs((int)mm) - 1;  // nearest word with a mismatch
            const uint32_t xs = __shfl_sync(0xFFFFFFFFu, x, src);
            const bool ins = __shfl_sync(0xFFFFFFFFu, (int)in, src) != 0;
            if (!ins) break;                     // the run reaches past what we may look at: unknown start
            const uint32_t hb = (31u - (uint32_t)__clz((int)xs)) >> 3;  // highest mismatching byte of that word
            r0 = pos - 4u * ((uint32_t)src + 1u) + hb + 1u;
            break;
        }
        pos -= 128u;
    }
    if (r0 == 0xFFFFFFFFu) return t_begin;
    const uint32_t phase = (t_begin - (r0 + 1u)) % LZ_MAXLEN;
    if (phase == 0u) return t_begin;
    const uint32_t p0 = t_begin + (LZ_MAXLEN - phase);
    // the run has to reach the aligned position (and the 3 bytes behind it, or the match there would not be a run match)
    bool same = true;
#pragma unroll
    for (uint32_t k = 0; k < 3u; ++k) {
        const uint32_t o = 4u * (lane + 32u * k);
        if (t_begin + o < p0 + 3u) {
            uint32_t x = ld_u32(S, t_begin + o) ^ splat;
            const uint32_t nb = p0 + 3u - (t_begin + o);
            if (nb < 4u) x &= (1u << (8u * nb)) - 1u;
            same = same && x == 0u;
        }
    }
    if (!__all_sync(0xFFFFFFFFu, same)) return t_begin;
    (void)t_end;
    return p0;
}

// True parse of a tile entered at `entry` (>= the tile's begin is not required: entry may lie past it):
// re-parse until a position the speculative parse visited, from there its tokens are reused.
// Writes fix tokens, returns the exit; *nfix_out / *from_out describe the splice.
__device__ __forceinline__ uint32_t lz_resync_tile(const LzS& S, const uint16_t* sorted,
                                                   const uint32_t* P, const uint32_t* hasbits, unsigned long long keep,
                                                   uint32_t entry, uint32_t t_begin, uint32_t t_end, uint32_t n,
                                                   uint32_t depth, uint32_t* __restrict__ fix_out, const uint32_t* visited,
                                                   uint32_t spec_count, uint32_t spec_exit, uint32_t* nfix_out,
                                                   uint32_t* from_out)
{
    const unsigned lane = zts_lane();
    uint32_t nfix = 0, from = spec_count;
    uint32_t p = entry;
    while (p < t_end) {
        if ((visited[p >> 5] >> (p & 31)) & 1u) {
            // met the speculative parse: its tokens from this position on are the true ones
            uint32_t idx = 0;
            for (uint32_t wd = (t_begin >> 5) + lane; wd <= (p >> 5); wd += 32) {
                uint32_t bits = visited[wd];
                if (wd == (p >> 5)) bits &= (1u << (p & 31)) - 1u;
                idx += __popc(bits);
            }
            from = __reduce_add_sync(0xFFFFFFFFu, idx);
            p = spec_exit;
            break;
        }
        uint32_t tok;
        const uint32_t np = lz_step(S, sorted, P, hasbits, keep, p, n, depth, &tok);
        if (lane == 0) fix_out[nfix] = tok;
        nfix++;
        p = np;
    }
    *nfix_out = nfix;
    *from_out = from;
    return p;
}

__device__ __forceinline__ void hist_token(uint32_t tok, uint32_t* hist)
{
    if (tok & TOK_MATCH) {
        uint32_t ls, lb, lv, ds, db, dv;
        zts_len_code(((tok >> 16) & 0xFF) + 3, ls, lb, lv);
        zts_dist_code((tok & 0xFFFF) + 1, ds, db, dv);
        atomicAdd(&hist[257 + ls], 1u);
        atomicAdd(&hist[286 + ds], 1u);
    } else {
        atomicAdd(&hist[tok & 0xFF], 1u);
    }
}

// ---- stages 1 and 2 of both kernels: the chunk into shared memory, the position index over it ----------------
// On return S holds the chunk (+ zeroed slack), `sorted` the positions grouped by the 16-bit hash of their 3-byte key
// (ascending inside a bucket). `cnt16` = 16 KiB of scratch for the radix sort.
struct LzIndexed {
    LzS SV;
    uint32_t m;  // positions that own a 3-byte key
};

__device__ __forceinline__ LzIndexed lz_stage_and_index(const uint8_t* __restrict__ src, uint32_t n, uint8_t* Sbuf,
                                                        uint16_t* sorted, uint16_t* cnt16, LzMisc* M, uint32_t* T,
                                                        uint32_t& phase)
{
    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned long long keep = l2_policy_keep();
    // ---- 1. stage the chunk: 16-byte aligned body by TMA bulk copy, ragged ends by plain loads
    const uint32_t head = min(n, (uint32_t)((16u - ((uintptr_t)src & 15u)) & 15u));
    const uint32_t body = (n - head) & ~15u;
    const uint32_t tail = n - head - body;
    uint8_t* S = Sbuf + ((16u - head) & 15u);  // S + head is 16-byte aligned
    const LzS SV = {S, smem_u32(S)};
    if (tid == 0 && body) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&M->mbar)), "r"(body)
                     : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                smem_u32(S + head)),
            "l"(src + head), "r"(body), "r"(smem_u32(&M->mbar))
            : "memory");
    }
    if (tid < head) S[tid] = src[tid];
    if (tid < tail) S[head + body + tid] = src[head + body + tid];
    if (tid < 32) S[n + tid] = 0;  // slack read by the word-wise compares
    if (body) {
        uint32_t done = 0, spins = 0;
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\t"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                "selp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done)
                : "r"(smem_u32(&M->mbar)), "r"(phase)
                : "memory");
            if (!done && ++spins > (1u << 26)) __trap();  // never hang the device
        }
        phase ^= 1;
    }
    __syncthreads();

    const uint32_t m = n >= 3 ? n - 2 : 0;  // positions that own a 3-byte key

    // ---- 2. stable LSD radix sort of positions by lz_hash(key), low 8 bits then high 8 bits, three sweeps:
    //   A  count the low digits per (warp range, digit)            -- shared-memory atomics, order irrelevant
    //   B  scatter by low digit into T (global, L2 resident) as pos | hash << 16, stable (ballot ranking),
    //      and count the high digits per (destination range, digit) on the way
    //   C  scatter by high digit from T into `sorted`, stable
    // The u32 counters of A and B live in the `sorted` area, which is not written before sweep C.
    {
        uint32_t* cntA = reinterpret_cast<uint32_t*>(sorted);   // [32 ranges][256 digits]
        uint32_t* cntB = cntA + 32 * 256;                       // [32 ranges][256 digits]
        for (uint32_t i = tid; i < 2u * 32u * 256u; i += LZ_THREADS) cntA[i] = 0;
        __syncthreads();
        const uint32_t w_begin = warp * LZ_SORT_TILE;
        // sweep A
        if (w_begin < m) {
            for (uint32_t it = 0; it < LZ_SORT_TILE / 32; ++it) {
                const uint32_t i = w_begin + it * 32 + lane;
                if (i < m) atomicAdd(&cntA[warp * 256u + (lz_hash(ld_u32(SV, i) & 0xFFFFFFu) & 255u)], 1u);
            }
        }
        __syncthreads();
        // exclusive scan in (digit major, range minor) order: entry e = d * 32 + w  ->  u16 running offsets
        for (int pass = 0; pass < 2; ++pass) {
            uint32_t* cnt = pass ? cntB : cntA;
            if (pass) {
                // sweep B
                if (w_begin < m) {
                    uint16_t* wc = cnt16 + warp * 256u;
                    // the key bytes of the next step are fetched while this one is ranked (the ranking is a chain of
                    // shared-memory round trips through wc[])
                    uint32_t key_next = w_begin + lane < m ? ld_u32(SV, w_begin + lane) : 0u;
                    for (uint32_t it = 0; it < LZ_SORT_TILE / 32; ++it) {
                        const uint32_t i = w_begin + it * 32 + lane;
                        const bool v = i < m;
                        const uint32_t hh = v ? lz_hash(key_next & 0xFFFFFFu) : 0xFFFFu;
                        key_next = (it + 1 < LZ_SORT_TILE / 32 && i + 32u < m) ? ld_u32(SV, i + 32u) : 0u;
                        const uint32_t d = hh & 255u;
                        const unsigned peers = peers_of<8>(d, v);
                        uint32_t dst = 0;
                        if (v) dst = (uint32_t)wc[d] + __popc(peers & zts_lanemask_lt());
                        __syncwarp();
                        // (the first lane of a group has read the group's offset itself)
                        if (v && lane == (unsigned)(__ffs((int)peers) - 1)) wc[d] = (uint16_t)(dst + __popc(peers));
                        if (v) {
                            ZTS_ASSERT(dst < m);
                            st_u32_hint(&T[dst], i | (hh << 16), keep);
                            atomicAdd(&cntB[(dst >> 11) * 256u + (hh >> 8)], 1u);
                        }
                        __syncwarp();
                    }
                }
                __threadfence_block();
                __syncthreads();
            }
            uint32_t vals[8];
            uint32_t sum = 0;
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) {
                const uint32_t e = tid * 8 + k;
                vals[k] = cnt[(e & 31u) * 256u + (e >> 5)];
                sum += vals[k];
            }
            uint32_t base = block_excl_sum(sum, M->warp_tot, nullptr);
#pragma unroll
            for (uint32_t k = 0; k < 8; ++k) {
                const uint32_t e = tid * 8 + k;
                cnt16[(e & 31u) * 256u + (e >> 5)] = (uint16_t)base;
                base += vals[k];
            }
            __syncthreads();
        }
        // sweep C (overwrites the counters, which are dead now)
        if (w_begin < m) {
            uint16_t* wc = cnt16 + warp * 256u;
            // the entries come back from L2: the loads run four batches ahead of the ranking that consumes them
            uint32_t pre[4];
#pragma unroll
            for (uint32_t k = 0; k < 4; ++k) {
                const uint32_t i = w_begin + k * 32 + lane;
                pre[k] = i < m ? ld_u32_hint(&T[i], keep) : 0xFFFFFFFFu;
            }
#pragma unroll 4
            for (uint32_t it = 0; it < LZ_SORT_TILE / 32; ++it) {
                const uint32_t i = w_begin + it * 32 + lane;
                const bool v = i < m;
                const uint32_t e = pre[it & 3u];
                {
                    const uint32_t i4 = i + 4u * 32u;
                    pre[it & 3u] = (it + 4u < LZ_SORT_TILE / 32 && i4 < m) ? ld_u32_hint(&T[i4], keep) : 0xFFFFFFFFu;
                }
                const uint32_t d = e >> 24;
                const unsigned peers = peers_of<8>(d, v);
                uint32_t dst = 0;
                if (v) dst = (uint32_t)wc[d] + __popc(peers & zts_lanemask_lt());
                __syncwarp();
                if (v && lane == (unsigned)(__ffs((int)peers) - 1)) wc[d] = (uint16_t)(dst + __popc(peers));
                ZTS_ASSERT(!v || dst < m);
                if (v) sorted[dst] = (uint16_t)e;
                __syncwarp();
            }
        }
        __syncthreads();
    }

    LzIndexed r = {SV, m};
    return r;
}

__global__ void __launch_bounds__(LZ_THREADS, 1)
lz77_chunk_kernel(const uint8_t* __restrict__ in, const ZtsChunk* __restrict__ chunks, uint32_t n_chunks,
                  ZtsChunkInfo* __restrict__ info, uint32_t* __restrict__ tile_tok, uint32_t* __restrict__ list_out,
                  uint32_t* __restrict__ hist_out, uint32_t* __restrict__ sortT, uint32_t* __restrict__ work_counter,
                  uint32_t depth, uint32_t lazy)
{
    extern __shared__ __align__(128) unsigned char smem[];
    uint8_t* Sbuf = smem + LzSmem::S_OFF;
    uint16_t* sorted = reinterpret_cast<uint16_t*>(smem + LzSmem::SORTED_OFF);
    uint16_t* cnt16 = reinterpret_cast<uint16_t*>(smem + LzSmem::BSTART_OFF);  // [32 warps][256 digits] running offsets of the radix sort
    uint32_t* hasbits = reinterpret_cast<uint32_t*>(smem + LzSmem::AUX_OFF); // [2048] bit per position (+ 1 padding word: the misc area follows)
    // once the info words are built the bucket starts are dead: visited bits and the second tile table take their place
    uint32_t* visited = reinterpret_cast<uint32_t*>(smem + LzSmem::BSTART_OFF); // [2048] bit per position (+ slack: the table follows)
    LzTiles2* M2 = reinterpret_cast<LzTiles2*>(smem + LzSmem::BSTART_OFF + 8192);
    static_assert(8192 + sizeof(LzTiles2) <= LzSmem::BSTART_BYTES, "second tile table must fit behind the visited bits");
    LzMisc* M = reinterpret_cast<LzMisc*>(smem + LzSmem::MISC_OFF);

    const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* T = sortT + (size_t)blockIdx.x * LZ_MAX_CHUNK;  // per-CTA radix temp (L2 resident): pos | hash << 16
    const unsigned long long keep = l2_policy_keep();
    uint32_t phase = 0;

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&M->mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    for (;;) {
        if (tid == 0) M->chunk = atomicAdd(work_counter, 1u);
        __syncthreads();
        const uint32_t c = M->chunk;
        if (c >= n_chunks) break;
        const ZtsChunk ch = chunks[c];
        // primed mode: `base` bytes of history are staged and indexed in front of the chunk; positions below base
        // are candidates only (the state of src/LZ77.ts:202-275 after it has walked them), parsing starts at base
        const uint32_t base = ch.dict_len;
        const uint32_t n = base + ch.len;
        const uint8_t* src = in + ch.in_off - base;

        // ---- 1 + 2. stage the chunk (TMA bulk copy) and index its positions (radix sort by key hash)
        const LzIndexed ix = lz_stage_and_index(src, n, Sbuf, sorted, cnt16, M, T, phase);
        const LzS SV = ix.SV;
        const uint32_t m = ix.m;
        (void)m;
        const uint32_t t0 = base ? lz_tile_of(base) : 0u;  // first tile with anything to parse

        // ---- 3. per-position info: slot and bucket rank into the L2 scratch (which the radix sort no longer needs),
        //         may-have-a-candidate bits into shared memory
        uint32_t* P = T;
        for (uint32_t i = tid; i < LZ_MAX_CHUNK / 32; i += LZ_THREADS) hasbits[i] = 0;
        __syncthreads();
        lz_build_info(SV, sorted, m, n, P, hasbits, keep);
        const uint32_t n_tiles = lz_tile_count(n);
        __syncthreads();  // the bucket starts are dead from here on: their area holds the visited bits and the second tile table
        // Where the speculative parse of every tile starts. A tile that begins inside a run of one byte starts at the
        // run's 258-byte phase: inside such a run every match is a maximum-length distance-1 match, so the true parse
        // visits run_start + 1 + 258 k, and starting there makes the predecessor's exit land on a parsed position (no
        // chain of wrongly entered tiles across the run). Purely a heuristic choice: any start >= the tile's first
        // position is a valid speculative parse. The run boundaries ("change points": S[p] != S[p - 1]) around every
        // tile boundary come from two block scans over 64-byte blocks: the last change point at or before the end of a
        // block, the first one at or behind its start.
        {
            uint16_t* lastcp = reinterpret_cast<uint16_t*>(M2);     // [1024] (the second tile table is not in use yet)
            uint16_t* nextcp = lastcp + LZ_THREADS;                 // [1024]; 0xFFFF = none
            static_assert(4u * LZ_THREADS <= sizeof(LzTiles2), "scan arrays must fit the second tile table");
            const uint32_t lo = 64u * tid;
            uint32_t last = 0xFFFFu, next = 0xFFFFu;
            if (lo < n) {
                uint32_t prevb = lo ? SV[lo - 1] : ~(uint32_t)SV[0];  // position 0 is a change point
#pragma unroll 4
                for (uint32_t k = 0; k < 16; ++k) {
                    const uint32_t at = lo + 4u * k;
                    if (at >= n) break;
                    const uint32_t w = ld_u32(SV, at);
                    uint32_t x = w ^ ((w << 8) | (prevb & 0xFFu));  // byte j != 0  <=>  S[at + j] != S[at + j - 1]
                    const uint32_t nb = n - at;
                    if (nb < 4u) x &= (1u << (8u * nb)) - 1u;
                    if (x) {
                        if (next == 0xFFFFu) next = at + (((uint32_t)__ffs((int)x) - 1u) >> 3);
                        last = at + ((31u - (uint32_t)__clz((int)x)) >> 3);
                    }
                    prevb = w >> 24;
                }
            }
            // inclusive prefix "last" (max, with none = 0xFFFF treated as -1) and suffix "next" (min) over the threads
            int lastv = last == 0xFFFFu ? -1 : (int)last;
            uint32_t nextv = next;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int a = __shfl_up_sync(0xFFFFFFFFu, lastv, d);
                const uint32_t b = __shfl_down_sync(0xFFFFFFFFu, nextv, d);
                if ((int)lane >= d) lastv = max(lastv, a);
                if ((int)lane + d < 32) nextv = min(nextv, b);
            }
            if (lane == 31) M->warp_tot[warp] = (uint32_t)lastv;
            if (lane == 0) M->warp_min[warp] = nextv;
            __syncthreads();
            int carry = -1;
            uint32_t carry_n = 0xFFFFu;
            for (uint32_t w2 = 0; w2 < warp; ++w2) carry = max(carry, (int)M->warp_tot[w2]);
            for (uint32_t w2 = warp + 1; w2 < LZ_WARPS; ++w2) carry_n = min(carry_n, M->warp_min[w2]);
            lastcp[tid] = (uint16_t)max(lastv, carry);  // -1 -> 0xFFFF
            nextcp[tid] = (uint16_t)min(nextv, carry_n);
            __syncthreads();
            for (uint32_t t = t0 + tid; t < n_tiles; t += LZ_THREADS) {
                const uint32_t t_begin = lz_tile_begin(t);
                uint32_t p0 = max(t_begin, base);  // the first tile starts where the parse starts
                if (t != t0 && t_begin >= 4u && t_begin + LZ_MAXLEN + 4u <= n) {
                    const uint32_t r0 = lastcp[(t_begin - 1u) >> 6];   // first byte of the run that holds t_begin - 1
                    uint32_t nx = nextcp[t_begin >> 6];                // first change point at or behind t_begin
                    if (nx == 0xFFFFu) nx = n;
                    if (r0 != 0xFFFFu && nx >= t_begin + 3u) {
                        const uint32_t phase = (t_begin - (r0 + 1u)) % LZ_MAXLEN;
                        const uint32_t cand = t_begin + (LZ_MAXLEN - phase);
                        if (phase && cand + 3u <= nx) p0 = cand;       // the run reaches the aligned position and 3 bytes on
                    }
                }
                M->tile_start[t] = (uint16_t)(p0 - t_begin);  // < 258
            }
        }
        __syncthreads();  // the scan arrays are dead: the second tile table takes their place
        for (uint32_t i = tid; i < LZ_MAX_CHUNK / 32; i += LZ_THREADS) visited[i] = 0;
        if (tid < (LZ_NTILES + 31) / 32) M->start_mask[tid] = 0;
        __syncthreads();
        uint32_t* spec_c = tile_tok + (size_t)blockIdx.x * (2u * LZ_TOK_PER_CHUNK);  // per-CTA scratch: speculative tokens per tile,
        uint32_t* fix_c = spec_c + LZ_TOK_PER_CHUNK;                                  // re-parsed tokens per tile
        uint32_t* list_c = list_out + (size_t)c * LZ_LIST_PER_CHUNK;                   // the chunk's token list (phase 5)
        ZtsChunkInfo* ci = info + c;

        // ---- 4a. every thread parses the tile it owns from the tile's start (speculative), then carries on into the
        //      next tile from its own exit (the re-entry parse of that tile) until it meets a position the next tile's
        //      owner has visited, or leaves that tile. The visited bits of a tile are written by its owner alone and
        //      read by its predecessor while they are being written: a bit seen too late only means a few more
        //      re-parsed tokens (the parse is memoryless in p: both threads produce the same tokens from a common
        //      position on), the splice itself is taken from the final bits behind the barrier.
        //      A lane is in one of four states: looking for its next position (ST_ADV), walking a candidate list
        //      (ST_SRCH), waiting for the warp to search a long list for it (ST_COOP), holding a result (ST_RES).
        {
            enum { ST_ADV = 0, ST_SRCH = 1, ST_COOP = 2, ST_RES = 3, ST_DONE = 4 };
#ifdef LZ_NO_INTERLEAVE
            const uint32_t t = tid;
#else
            // the 32 tiles of a warp lie 2 KiB apart: neighbouring tiles cost alike (they hold the same kind of data),
            // so this spreads expensive stretches of the chunk over all the warps
            const uint32_t t = lane * LZ_WARPS + ((lane + warp) & 31u);  // (33 tiles between neighbouring lanes: their table words fall into different banks)
#endif
            const uint32_t t_begin = t * LZ_TILE_POS;
            uint32_t st = ST_DONE;
            bool resync = false;      // false: own tile; true: tile t + 1, entered at `entry`
            uint32_t p = 0, t_end = 0, entry = 0;
            uint32_t tok_i = 0, tok_0 = 0;   // next / first token slot of the parse in hand
            uint32_t* tbuf = spec_c;
            uint32_t pinfo = 0;       // P[p], requested as soon as p is known
            bool have_info = true;    // pinfo belongs to p
            uint32_t s_cur = 0, s_left = 0, best = 0, ptail = 0, pw = 0, pw1 = 0, maxlen = 0;
            uint32_t saved = 0;  // lazy evaluation: the match (len << 16 | q) found at p - 1, waiting for the search at p
#ifdef LZ_DEBUG_COUNTS
            uint32_t dbg_s[2] = {0, 0}, dbg_steps[2] = {0, 0}, dbg_coop[2] = {0, 0}, dbg_tok[2] = {0, 0}, dbg_iter = 0;
            if (tid < 16) M->hist[tid] = 0;
            __syncthreads();
#endif
            if (t >= t0 && t < n_tiles) {
                p = t_begin + M->tile_start[t];
                t_end = min(n, t_begin + LZ_TILE_POS);
                tok_0 = tok_i = lz_tok_off(t);
                if (p < n) pinfo = ld_u32_hint(&P[p], keep);
                st = ST_ADV;
            }
            for (;;) {
                // -- candidate steps of the searching lanes: up to LZ_KMAX in a row, then the lanes that wait are served
                {
                    unsigned busy = __ballot_sync(0xFFFFFFFFu, st == ST_SRCH);
                    uint32_t it = 0;
                    while (busy) {
                        if (st == ST_SRCH) {
#ifdef LZ_DEBUG_COUNTS
                            dbg_steps[resync]++;
#endif
                            // One code path for every searching lane. Behind a known match only a strictly longer one
                            // counts (:183): its byte at best_len must match -- four candidates per step go through that
                            // test alone and the nearest that passes gets the exact comparison. Before a match is known
                            // the step looks at one candidate, which "passes" unconditionally.
                            const uint32_t bl = best >> 16;
                            const bool known = bl >= 3u;
                            const uint32_t c4 = known ? min(4u, s_left) : 1u;
                            ZTS_ASSERT(s_left >= 1u && s_cur >= c4 && s_cur <= m);
                            // the four slots in front of the cursor and their tail bytes, whatever c4 is: slots in front
                            // of the list (down to three u16 in front of `sorted`: the chunk buffer's slack) are read and
                            // masked out -- no branch in the step
                            const uint16_t* sp = sorted + s_cur;
                            const uint32_t q0 = sp[-1], q1 = sp[-2], q2 = sp[-3], q3 = sp[-4];
                            uint32_t hm = (SV[q0 + bl] == ptail ? 1u : 0u) | (SV[q1 + bl] == ptail ? 2u : 0u) |
                                          (SV[q2 + bl] == ptail ? 4u : 0u) | (SV[q3 + bl] == ptail ? 8u : 0u);
                            hm = known ? hm & ((1u << c4) - 1u) : 1u;
                            // the candidate that decides: the nearest that passed, else the oldest looked at -- it tells
                            // whether the window ends here (ascending positions: older ones are outside it too,
                            // src/LZ77.ts:223)
                            const uint32_t j = hm ? (uint32_t)__ffs((int)hm) - 1u : c4 - 1u;
                            const uint32_t q = j == 0u ? q0 : j == 1u ? q1 : j == 2u ? q2 : q3;
                            s_cur -= j + 1u;
                            s_left -= j + 1u;
                            const bool inside = p - q <= LZ_WINDOW;
                            bool end = s_left == 0u || !inside;
                            ZTS_ASSERT(q < p && q + bl < n + 32u);
                            if (hm && inside) {
                                const uint32_t len = lz_match_len(SV, q, p, pw, pw1, maxlen);
                                if (len > bl) {
                                    best = (len << 16) | q;
                                    ptail = SV[p + len];
                                    end = end || len >= maxlen;  // :189 (258) or capped by the input end
                                }
                            }
                            if (end) st = ST_RES;
                        }
                        busy = __ballot_sync(0xFFFFFFFFu, st == ST_SRCH);
                        const unsigned waiting = __ballot_sync(0xFFFFFFFFu, st == ST_RES || st == ST_ADV);
                        if (++it >= LZ_KMAX || LZ_WAIT_MUL * __popc(waiting) >= __popc(busy)) break;
                    }
                }
                // -- searches behind long candidate lists: the whole warp, one request after the other
                {
                    unsigned cm = __ballot_sync(0xFFFFFFFFu, st == ST_COOP);
                    while (cm) {
                        const int src = __ffs((int)cm) - 1;
                        cm &= cm - 1u;
                        const uint32_t bp = __shfl_sync(0xFFFFFFFFu, p, src), bs = __shfl_sync(0xFFFFFFFFu, s_cur, src),
                                       br = __shfl_sync(0xFFFFFFFFu, s_left, src);
                        const uint32_t rr = lz_search_from(SV, sorted, bs - br, bs, bp, n, depth);
                        if (lane == (unsigned)src) {
                            best = rr ? (rr & 0xFFFF0000u) | (p - (rr & 0xFFFFu)) : 0u;
                            st = ST_RES;
                        }
                    }
                }
                // -- the token of a searched position
                if (st == ST_RES) {
                    bool emit = true;
                    uint32_t at = p, res = best;  // the token goes to position `at`
                    if (saved) {
                        // lazy evaluation: `saved` is the match at p - 1, `best` the one at p
                        at = p - 1u;
                        if ((best >> 16) > (saved >> 16)) {
                            res = 0;        // the later match is longer: p - 1 becomes a literal, p is looked at again
                            best = 0;       // (from the top: it may be the end of the tile or a visited position)
                        } else {
                            res = saved;
                        }
                        saved = 0;
                    } else if (lazy && best >= (3u << 16) && best < (LZ_LAZY_MAX << 16) &&
                               ((hasbits[(p + 1u) >> 5] >> ((p + 1u) & 31u)) & 1u)) {
                        // a short match, and the next position has candidates too: decide after looking there
                        saved = best;
                        p += 1u;
                        pinfo = ld_u32_hint(&P[p], keep);
                        emit = false;
                        const uint32_t rank = pinfo >> 16;
                        s_cur = pinfo & 0xFFFFu;
                        maxlen = min(LZ_MAXLEN, n - p);
                        best = 0;
                        s_left = min(rank, depth);
                        pw = ld_u32(SV, p);
                        pw1 = ld_u32(SV, p + 4u);
                        st = s_left ? ST_SRCH : ST_RES;  // (lazy is a fast-mode option: depth <= LZ_PRIV_CAP, no warp searches)
                    }
                    if (emit) {
                        uint32_t tok, np;
                        if (res >= (3u << 16)) {
                            const uint32_t len = res >> 16, dist = at - (res & 0xFFFFu);
                            tok = TOK_MATCH | ((len - 3u) << 16) | (dist - 1u);
                            np = at + len;
                        } else {
                            tok = SV[at];  // the bit was a "maybe" (hash collisions), or everything lies outside the window
                            np = at + 1u;
                        }
                        ZTS_ASSERT(tok_i < tok_0 + LZ_TILE_POS + 8u && at < n && (resync || at / LZ_TILE_POS == t));
#ifdef LZ_DEBUG_COUNTS
                        dbg_tok[resync]++;
#endif
                        tbuf[tok_i++] = tok;
                        if (!resync) visited[at >> 5] |= 1u << (at & 31u);
                        p = np;
                        if (p < n) pinfo = ld_u32_hint(&P[p], keep);
                        have_info = true;
                        st = ST_ADV;
                    }
                }
                // -- next position: end of the tile, the splice, a run of literals, or a search
                if (st == ST_ADV) {
                    bool fin = p >= t_end, spliced = false;
                    uint32_t vlimit = 32u;
                    if (!fin && resync) {
                        const volatile uint32_t* vv = visited;
                        const uint32_t vm = __funnelshift_r(vv[p >> 5], vv[(p >> 5) + 1u], p & 31u);
                        if (vm & 1u) {
                            fin = spliced = true;  // met the speculative parse: its tokens from this position on are the true ones
                        } else if (vm) {
                            vlimit = (uint32_t)__ffs((int)vm) - 1u;  // a run of literals stops in front of a visited position
                        }
                    }
                    if (fin) {
                        if (!resync) {
                            M2->spec_exit[t] = (uint16_t)(p - t_begin);
                            M->spec_count[t] = (LZ_COUNT_T)(tok_i - tok_0);
                            if (t + 1u < n_tiles) {
                                resync = true;
                                entry = p;
                                t_end = min(n, t_begin + 2u * LZ_TILE_POS);
                                tbuf = fix_c;
                                tok_0 = tok_i = lz_tok_off(t + 1u);
                            } else {
                                st = ST_DONE;
                            }
                        } else {
                            const uint32_t nb = t_begin + LZ_TILE_POS;  // first position of tile t + 1
                            M->entry_used[t + 1u] = (uint16_t)(entry - nb);
                            M->fix_count[t + 1u] = (LZ_COUNT_T)(tok_i - tok_0);
                            M2->fix_exit[t + 1u] = spliced ? (uint16_t)LZ_EXIT_SPLICED : (uint16_t)(p - nb);
                            M2->spec_from[t + 1u] = spliced ? (LZ_COUNT_T)(p - nb) : (LZ_COUNT_T)LZ_NO_SPLICE;
                            st = ST_DONE;
                        }
                    } else {
                        // may-have-a-candidate bits of the 32 positions from p on
                        const uint32_t w = p >> 5, off = p & 31u;
                        const uint32_t mm = __funnelshift_r(hasbits[w], hasbits[w + 1], off);  // w + 1 <= 2048: padding word
                        const uint32_t avail = min(min(32u, t_end - p), vlimit);
                        const uint32_t k = mm ? min((uint32_t)__ffs((int)mm) - 1u, avail) : avail;
                        if (k) {
                            // k positions without any candidate: k literals (src/LZ77.ts:267-272)
                            ZTS_ASSERT(tok_i + k <= tok_0 + LZ_TILE_POS + 8u && p + k <= t_end && t_end <= n);
                            {
                                uint32_t o = 0;  // single tokens up to a 16-byte boundary of the token slots, then four per store
                                for (; o < k && ((tok_i + o) & 3u); ++o) tbuf[tok_i + o] = SV[p + o];
                                for (; o + 4u <= k; o += 4u) {
                                    const uint32_t w4 = ld_u32(SV, p + o);
                                    *reinterpret_cast<uint4*>(tbuf + tok_i + o) =
                                        make_uint4(w4 & 0xFFu, (w4 >> 8) & 0xFFu, (w4 >> 16) & 0xFFu, w4 >> 24);
                                }
                                for (; o < k; ++o) tbuf[tok_i + o] = SV[p + o];
                            }
                            tok_i += k;
                            if (!resync) {  // visited bits [p, p + k): at most two words, this lane's own tile
                                const unsigned long long bits = (0xFFFFFFFFFFFFFFFFull >> (64u - k)) << off;
                                visited[w] |= (uint32_t)bits;
                                if (bits >> 32) visited[w + 1] |= (uint32_t)(bits >> 32);
                            }
                            p += k;
                            have_info = false;  // (fetched below if a search follows at once: literal-only stretches never ask)
                        }
                        if (k < avail) {  // p has candidates
                            if (!have_info) pinfo = ld_u32_hint(&P[p], keep);
                            have_info = true;
                            const uint32_t rank = pinfo >> 16;  // earlier entries of the bucket: the slots in front of p's own
                            s_cur = pinfo & 0xFFFFu;
                            ZTS_ASSERT(p + 3u < n && rank <= s_cur && s_cur < m && sorted[s_cur] == p);
                            maxlen = min(LZ_MAXLEN, n - p);
                            best = 0;
#ifdef LZ_DEBUG_COUNTS
                            dbg_s[resync]++;
                            if (rank > LZ_PRIV_CAP && depth > LZ_PRIV_CAP) dbg_coop[resync]++;
#endif
                            if (rank > LZ_PRIV_CAP && depth > LZ_PRIV_CAP) {
                                s_left = rank;
                                st = ST_COOP;
                            } else {
                                s_left = min(rank, depth);
                                pw = ld_u32(SV, p);
                                pw1 = ld_u32(SV, p + 4u);
                                st = s_left ? ST_SRCH : ST_RES;
                            }
                        }
                    }
                }
#ifdef LZ_DEBUG_COUNTS
                dbg_iter++;
#endif
                if (__all_sync(0xFFFFFFFFu, st == ST_DONE)) break;
            }
#ifdef LZ_DEBUG_COUNTS
            if (blockIdx.x < 4) {
                for (int k = 0; k < 2; ++k) {
                    atomicAdd(&M->hist[k * 4 + 0], dbg_s[k]);
                    atomicAdd(&M->hist[k * 4 + 1], dbg_steps[k]);
                    atomicAdd(&M->hist[k * 4 + 2], dbg_coop[k]);
                    atomicAdd(&M->hist[k * 4 + 3], dbg_tok[k]);
                }
                if (lane == 0) atomicAdd(&M->hist[8], dbg_iter);
                if (lane == 0) atomicMax(&M->hist[9], dbg_iter);
                __syncthreads();
                if (tid == 0)
                    printf("chunk %u: spec srch %u steps %u coop %u tok %u | fix srch %u steps %u coop %u tok %u | warp iters sum %u max %u\n", c,
                           M->hist[0], M->hist[1], M->hist[2], M->hist[3], M->hist[4], M->hist[5], M->hist[6], M->hist[7], M->hist[8], M->hist[9]);
                __syncthreads();
                if (tid < 16) M->hist[tid] = 0;
                __syncthreads();
            }
#endif
        }
        __threadfence_block();
        __syncthreads();
        // the splices, from the final visited bits; the first tile needs no re-entry
        if (tid >= t0 && tid < n_tiles) {
            const uint32_t t = tid;
            if (t == t0) {
                M->entry_used[t] = (uint16_t)(base - lz_tile_begin(t0));
                M->fix_count[t] = 0;
                M2->fix_exit[t] = M2->spec_exit[t];
                M2->spec_from[t] = 0;
            } else {
                const uint32_t x = M2->spec_from[t];
                if (x == LZ_NO_SPLICE) {
                    M2->spec_from[t] = M->spec_count[t];
                } else {
                    const unsigned long long vis = (unsigned long long)visited[2u * t] | ((unsigned long long)visited[2u * t + 1u] << 32);
                    M2->spec_from[t] = (LZ_COUNT_T)__popcll(vis & ((1ull << x) - 1ull));
                    M2->fix_exit[t] = M2->spec_exit[t];
                }
            }
        }
        __syncthreads();
        // ---- 4b. tiles entered at the wrong place form chains (consecutive tiles inside one long run of
        //          matches that never meets the speculative parse); chains are independent, one warp each
        for (uint32_t w = t0 + 1 + tid; w < n_tiles; w += LZ_THREADS) {
            const bool bad = lz_tile_begin(w - 1) + M2->fix_exit[w - 1] != lz_tile_begin(w) + M->entry_used[w];
            const bool bad_prev = w >= t0 + 2 && lz_tile_begin(w - 2) + M2->fix_exit[w - 2] != lz_tile_begin(w - 1) + M->entry_used[w - 1];
            if (bad && !bad_prev) atomicOr(&M->start_mask[w >> 5], 1u << (w & 31));
        }
        __syncthreads();
        for (uint32_t w = t0 + 1 + warp; w < n_tiles; w += LZ_WARPS) {
            if (!((M->start_mask[w >> 5] >> (w & 31)) & 1u)) continue;
            uint32_t end = w + 1;  // next chain start (or the end of the chunk)
            while (end < n_tiles && !((M->start_mask[end >> 5] >> (end & 31)) & 1u)) ++end;
            uint32_t entry = lz_tile_begin(w - 1) + M2->fix_exit[w - 1];
            for (uint32_t t = w; t < end; ++t) {
                if (entry == lz_tile_begin(t) + M->entry_used[t]) break;
                const uint32_t t_begin = lz_tile_begin(t), t_end = min(n, lz_tile_begin(t + 1));
                uint32_t nfix, from;
                const uint32_t ex = lz_resync_tile(SV, sorted, P, hasbits, keep, entry, t_begin, t_end, n, depth,
                                                   fix_c + lz_tok_off(t), visited, M->spec_count[t],
                                                   t_begin + M2->spec_exit[t], &nfix, &from);
                if (lane == 0) {
                    M->entry_used[t] = (uint16_t)(entry - t_begin);
                    M2->fix_exit[t] = (uint16_t)(ex - t_begin);
                    M->fix_count[t] = (LZ_COUNT_T)nfix;
                    M2->spec_from[t] = (LZ_COUNT_T)from;
                }
                __syncwarp();
                entry = ex;
            }
        }
        __syncthreads();
        // ---- 4c. walk the chain of true exits once; anything still entered at the wrong place is redone here
        //          (normally nothing: this pass is what makes 4a/4b safe to run optimistically)
        bool consistent = true;  // every tile entered where its predecessor exits: then the walk has nothing to find
        for (uint32_t w = t0 + 1 + tid; w < n_tiles; w += LZ_THREADS)
            consistent = consistent && lz_tile_begin(w - 1) + M2->fix_exit[w - 1] == lz_tile_begin(w) + M->entry_used[w];
        if (!__syncthreads_and(consistent) && warp == 0 && t0 < n_tiles) {
            uint32_t true_exit = lz_tile_begin(t0) + M2->fix_exit[t0];
            for (uint32_t w = t0 + 1; w < n_tiles; ++w) {
                if (true_exit == lz_tile_begin(w) + M->entry_used[w]) {
                    true_exit = lz_tile_begin(w) + M2->fix_exit[w];
                    continue;
                }
                const uint32_t t_begin = lz_tile_begin(w), t_end = min(n, lz_tile_begin(w + 1));
                uint32_t nfix, from;
                const uint32_t entry = true_exit;
                true_exit = lz_resync_tile(SV, sorted, P, hasbits, keep, entry, t_begin, t_end, n, depth, fix_c + lz_tok_off(w),
                                           visited, M->spec_count[w], t_begin + M2->spec_exit[w], &nfix, &from);
                if (lane == 0) {
                    M->entry_used[w] = (uint16_t)(entry - t_begin);
                    M2->fix_exit[w] = (uint16_t)(true_exit - t_begin);
                    M->fix_count[w] = (LZ_COUNT_T)nfix;
                    M2->spec_from[w] = (LZ_COUNT_T)from;
                }
            }
        }
        for (uint32_t i = tid; i < 316; i += LZ_THREADS) M->hist[i] = 0;
        if (tid == 0) M->n_tokens = 0;
        __threadfence_block();
        __syncthreads();

        // ---- 5. the chunk's token list + histograms (src/LZ77.ts:126-128,141-142,236,251,271,279): per tile the
        //         re-parsed tokens, then the speculative ones from the splice on, tiles in order -- one contiguous list,
        //         written once (streaming stores) and read once by the bit packer
        uint32_t* tok_off = reinterpret_cast<uint32_t*>(M2->spec_exit);  // the two exit tables are dead now: u32 per tile
        static_assert(offsetof(LzTiles2, fix_exit) == offsetof(LzTiles2, spec_exit) + sizeof(uint16_t) * LZ_NTILES,
                      "spec_exit and fix_exit must be adjacent");
        {
            uint32_t cnt = 0;
            if (tid < LZ_NTILES && tid >= t0 && tid < n_tiles)
                cnt = (uint32_t)M->fix_count[tid] + ((uint32_t)M->spec_count[tid] - (uint32_t)M2->spec_from[tid]);
            const uint32_t off = block_excl_sum(cnt, M->warp_tot, &M->n_tokens);
            if (tid < LZ_NTILES) tok_off[tid] = off;  // (block_excl_sum ends with a barrier: the exit tables were read before it)
        }
        __syncthreads();
        {
            const unsigned long long stream = l2_policy_stream();
            // a warp takes 32 consecutive tiles (their tokens are consecutive in the list); the first 64 tokens of the
            // next tile are requested before the current tile's are stored, so the trip to the L2 scratch overlaps
            const uint32_t NONE = 0xFFFFFFFFu;  // (no token has all bits set)
            auto fetch = [&](uint32_t w, uint32_t k) -> uint32_t {  // token k of tile w, NONE behind its end
                if (w >= n_tiles) return NONE;
                const uint32_t nf = M->fix_count[w], from = M2->spec_from[w], sc = M->spec_count[w];
                if (k >= nf + (sc - from)) return NONE;
                return k < nf ? fix_c[lz_tok_off(w) + k] : spec_c[lz_tok_off(w) + from + (k - nf)];
            };
            const uint32_t w_first = t0 + warp * 32u;
            uint32_t a0 = fetch(w_first, lane), a1 = fetch(w_first, lane + 32u);
            for (uint32_t w = w_first; w < min(n_tiles, w_first + 32u); ++w) {
                const uint32_t b0 = fetch(w + 1u < w_first + 32u ? w + 1u : n_tiles, lane),
                               b1 = fetch(w + 1u < w_first + 32u ? w + 1u : n_tiles, lane + 32u);
                uint32_t* dst = list_c + tok_off[w];
                ZTS_ASSERT(tok_off[w] <= n);
                if (a0 != NONE) {
                    st_u32_hint(&dst[lane], a0, stream);
                    hist_token(a0, M->hist);
                }
                if (a1 != NONE) {
                    st_u32_hint(&dst[lane + 32u], a1, stream);
                    hist_token(a1, M->hist);
                }
                if (__any_sync(0xFFFFFFFFu, a1 != NONE)) {  // a tile holds up to 64 re-parsed + 64 speculative tokens
                    for (uint32_t k = lane + 64u; k < 2u * LZ_TILE_POS; k += 32u) {
                        const uint32_t tok = fetch(w, k);
                        if (tok != NONE) {
                            st_u32_hint(&dst[k], tok, stream);
                            hist_token(tok, M->hist);
                        }
                    }
                }
                a0 = b0;
                a1 = b1;
            }
        }
        __syncthreads();
        for (uint32_t i = tid; i < 316; i += LZ_THREADS)
            hist_out[(size_t)c * 316 + i] = M->hist[i] + (i == 256 ? 2u : 0u);  // freqsLitLen[256] ends at 2
        if (tid == 0) {
            ci->n_tokens = M->n_tokens;
            ci->pad = 0;
        }
        __syncthreads();
    }
}

size_t zts_lz77_smem_bytes() { return LzSmem::TOTAL; }

size_t zts_lz77_scratch_bytes(int sm_count) { return (size_t)sm_count * 2u * LZ_TOK_PER_CHUNK * 4u; }

int zts_lz77_launch(zlb_ctx* ctx, const uint8_t* d_in, const ZtsChunk* d_chunks, uint32_t n_chunks,
                    ZtsChunkInfo* d_info, uint32_t* d_tile_tok, uint32_t* d_list, uint32_t* d_hist, uint32_t* d_sortT,
                    uint32_t* d_counter, uint32_t grid, uint32_t depth, uint32_t lazy)
{
    static_assert(sizeof(LzMisc) <= LzSmem::MISC_BYTES, "misc area too small");
    static_assert(LzSmem::TOTAL <= 232448, "exceeds 227 KiB of dynamic shared memory");
    if (grid > (uint32_t)ctx->sm_count) grid = (uint32_t)ctx->sm_count;  // the tile scratch holds one slot per SM
    ZTS_CUDA(ctx, cudaFuncSetAttribute(lz77_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)LzSmem::TOTAL));
    ZTS_CUDA(ctx, cudaMemsetAsync(d_counter, 0, sizeof(uint32_t), ctx->work));
    ZTS_LAUNCH(ctx, ZK_LZ77,
               lz77_chunk_kernel<<<grid, LZ_THREADS, LzSmem::TOTAL, ctx->work>>>(
                   d_in, d_chunks, n_chunks, d_info, d_tile_tok, d_list, d_hist, d_sortT, d_counter, depth, lazy));
    return ZLB_OK;
}
