// zts_huffman.cu -- per-chunk Huffman construction and dynamic-block header
// (replaces RawDeflate.getLengths / reversePackageMerge / getCodesFromLengths / getTreeSymbols and the
//  header part of makeDynamicHuffmanBlock, src/RawDeflate.ts:181-241,341-611, and Heap, src/Heap.ts).
//
// Byte-identity with the reference needs its exact tie-breaking: code lengths are handed to symbols
// in the pop order of its (value,index) max-heap over Uint16 frequencies, and the lengths themselves
// come from its "reverse package merge" including the JS `undefined`/NaN comparisons (SURVEY App. A.2,
// B-6). Both are restated literally; one warp owns one chunk and lane 0 runs the serial parts out of
// shared memory (a few hundred symbols; the kernel is latency- not bandwidth-bound and many chunks
// run concurrently). Output per chunk: code tables for the bit packer, the header bit string, and
// the exact size of the block, which is what the output-offset scan needs.
#include "zts_deflate.cuh"

#define HUF_MAXSYM 288
#define HUF_LVL 576  // >= 2 * symbols: a level holds at most symbols + symbols packages
#define HUF_UNDEF 0xFFFFFFFFu
#define HUF_PKW (HUF_LVL / 32)

struct HufWork {
    // val[0] doubles as the histogram being coded (read only while the heap is filled) and val[1] as
    // the heap (dead once everything is popped): both are free again when the level lists are built.
    // Three more tenants live in val[] while the level lists do not need the space (a code-length alphabet of
    // 19 symbols uses 38 entries of each list): see huf_clen / huf_tree_sym / huf_code_tmp below.
    uint32_t val[2][HUF_LVL];        // value[j+1], value[j] of reversePackageMerge
    uint32_t pkbits[15][HUF_PKW];    // per level: bit t set <=> item t is a package (type[j][t] === symbols)
    uint16_t nval[HUF_MAXSYM];       // frequencies in heap pop order (descending); Uint16 like the heap's values
    uint16_t nidx[HUF_MAXSYM];       // their symbols
    uint8_t ll_len[HUF_MAXSYM];
    uint8_t d_len[32];
    uint8_t t_len[20];
    uint32_t hdr_bits;
};
// 7.0 KiB: 28 chunks per SM (227 KiB less 1 KiB per CTA), so the 4096 chunks of a 256 MiB wave are all resident at
// once on 148 SMs -- the kernel is latency-bound and its duration is that of the slowest round of chunks.
static_assert(sizeof(HufWork) <= 7296, "HufWork must let 28 CTAs share an SM");

// code lengths in pop order, written by reversePackageMerge when its level lists are dead
__device__ __forceinline__ uint8_t* huf_clen(HufWork* W) { return reinterpret_cast<uint8_t*>(W->val[0]); }
// getTreeSymbols result (<= 632 bytes): outlives the code-length alphabet's getLengths, which touches val[0][0..38)
__device__ __forceinline__ uint8_t* huf_tree_sym(HufWork* W) { return reinterpret_cast<uint8_t*>(W->val[0] + 256); }
// scratch of getCodesFromLengths (288 codes) and the staged code-length-symbol counts: val[1][64..208)
__device__ __forceinline__ uint16_t* huf_code_tmp(HufWork* W) { return reinterpret_cast<uint16_t*>(W->val[1] + 64); }

// ---- Heap (src/Heap.ts:49-132): (value,index) pairs in one Uint16Array --------------------------
// A pair occupies heap[2k] (value) and heap[2k+1] (index): one aligned 32-bit word, value in the low half. The
// reference swaps the moving pair with its parent / larger child level by level; carrying it in a register and
// storing it once where it comes to rest leaves exactly the same array (every swap writes the other pair one level
// along the path, which is what the hole does), with one load and one store per level instead of four of each.
__device__ void heap_push(uint16_t* heap, int& length, uint16_t index, uint16_t value)
{
    uint32_t* h = reinterpret_cast<uint32_t*>(heap);
    const uint32_t x = (uint32_t)value | ((uint32_t)index << 16);
    int current = length;  // in Uint16 slots, as in the reference
    length += 2;
    while (current > 0) {
        const int parent = ((current - 2) >> 2) << 1;  // src/Heap.ts:31
        const uint32_t pp = h[parent >> 1];
        if ((x & 0xFFFFu) > (pp & 0xFFFFu)) {          // src/Heap.ts:62
            h[current >> 1] = pp;
            current = parent;
        } else {
            break;
        }
    }
    h[current >> 1] = x;
}

__device__ void heap_pop(uint16_t* heap, int& length, uint16_t& index, uint16_t& value)
{
    uint32_t* h = reinterpret_cast<uint32_t*>(heap);
    const uint32_t top = h[0];
    value = (uint16_t)top;
    index = (uint16_t)(top >> 16);
    length -= 2;
    const uint32_t x = h[length >> 1];  // the last pair moves to the root (src/Heap.ts:98-99) and sinks
    int parent = 0;
    for (;;) {
        int current = 2 * parent + 2;   // src/Heap.ts:38
        if (current >= length) break;
        uint32_t c = h[current >> 1];
        if (current + 2 < length) {     // the larger child; the left one on equal values (src/Heap.ts:108)
            const uint32_t c2 = h[(current + 2) >> 1];
            if ((c2 & 0xFFFFu) > (c & 0xFFFFu)) {
                c = c2;
                current += 2;
            }
        }
        if ((c & 0xFFFFu) > (x & 0xFFFFu)) {  // src/Heap.ts:113
            h[parent >> 1] = c;
        } else {
            break;
        }
        parent = current;
    }
    if (length > 0) h[parent >> 1] = x;
}

// ---- reversePackageMerge (src/RawDeflate.ts:484-571), warp-cooperative ---------------------------------
// The reference builds, from the deepest level up, value[j] = the first minimumCost[j] items of the merge
// of the symbols (descending freqs) with the packages value[j+1][next] + value[j+1][next+1], "package
// first iff weight > freqs[i]" (:553), and interleaves the recursive takePackage (:496-507). Two facts
// make this data-parallel without changing a single result (checked against the statement-by-statement
// model on 14 000 cases incl. ties, tests/test_oracle.py + tests/test_deflate_gpu.py):
//   1. when level j is built, currentPosition[j+1] is just flag[j+1]: takePackage only ever descends, so
//      the levels do not depend on the recursion, and a level is a plain merge of two sorted lists:
//      symbol i lands at i + #{k : W_k > f_i}, package k at k + #{i : f_i >= W_k} (two binary searches);
//   2. the items taken from a level are always a prefix of it: T[0] = flag[0], and with P[j] packages
//      among the first T[j] items of level j, T[j+1] = flag[j+1] + 2 P[j]; the S[j] = T[j] - P[j] symbols
//      taken there are symbols 0 .. S[j]-1, so codeLength[i] = limit - #{j : i < S[j]}.
// JS `undefined` (App. B-6): a sum with an out-of-range value[][] is NaN and compares false, i.e. "no
// package left"; once the symbols are exhausted too the reference stores undefined items, which can
// never be decremented -- modelled by HUF_UNDEF tail entries and the `defined` count.
__device__ void rpm_warp(const uint16_t* freqs, int symbols, int limit, uint8_t* code_length, HufWork* W)
{
    const int lane = (int)zts_lane();
    int flag[16], size[16];
    {
        uint16_t mc[16];
        for (int j = 0; j < 16; ++j) mc[j] = 0;
        mc[limit - 1] = (uint16_t)symbols;
        int excess = (1 << limit) - symbols;
        const int half = 1 << (limit - 1);
        for (int j = 0; j < limit; ++j) {
            if (excess < half) {
                flag[j] = 0;
            } else {
                flag[j] = 1;
                excess -= half;
            }
            excess <<= 1;
            if (limit - 2 - j >= 0) mc[limit - 2 - j] = (uint16_t)((mc[limit - 1 - j] >> 1) + symbols);
        }
        mc[0] = (uint16_t)flag[0];
        for (int j = 1; j < limit; ++j)
            if (mc[j] > 2 * mc[j - 1] + flag[j]) mc[j] = (uint16_t)(2 * mc[j - 1] + flag[j]);
        for (int j = 0; j < limit; ++j) size[j] = mc[j];
    }
    // deepest level: the symbols themselves
    {
        uint32_t* v = W->val[(limit - 1) & 1];
        for (int t = lane; t < size[limit - 1]; t += 32) v[t] = t < symbols ? freqs[t] : HUF_UNDEF;
        for (int w = lane; w < HUF_PKW; w += 32) W->pkbits[limit - 1][w] = 0;
    }
    int defined = min(size[limit - 1], symbols);
    const uint32_t fmin = freqs[symbols - 1];
    __syncwarp();
    for (int j = limit - 2; j >= 0; --j) {
        const uint32_t* pv = W->val[(j + 1) & 1];
        uint32_t* v = W->val[j & 1];
        uint32_t* bits = W->pkbits[j];
        const int n0 = flag[j + 1];
        const int K = defined > n0 ? (defined - n0) >> 1 : 0;  // packages that have two defined halves
        const int sz = size[j];
        for (int t = lane; t < sz; t += 32) v[t] = HUF_UNDEF;
        for (int w = lane; w < HUF_PKW; w += 32) bits[w] = 0;
        __syncwarp();
        for (int i = lane; i < symbols; i += 32) {
            const uint32_t fi = freqs[i];
            int lo = 0, hi = K;  // first package whose weight is not > f_i (weights descend)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (pv[n0 + 2 * mid] + pv[n0 + 2 * mid + 1] > fi)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            if (i + lo < sz) v[i + lo] = fi;
        }
        for (int k = lane; k < K; k += 32) {
            const uint32_t wk = pv[n0 + 2 * k] + pv[n0 + 2 * k + 1];
            int lo = 0, hi = symbols;  // first symbol with f_i < W_k (freqs descend)
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (freqs[mid] >= wk)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            // lo == symbols: every symbol precedes it, and past the last symbol the reference compares
            // against freqs[symbols] === undefined and never takes a package again
            if (lo < symbols && k + lo < sz) {
                v[k + lo] = wk;
                atomicOr(&bits[(k + lo) >> 5], 1u << ((k + lo) & 31));
            }
        }
        {
            int lo = 0, hi = K;  // packages heavier than the lightest symbol = packages that get placed
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (pv[n0 + 2 * mid] + pv[n0 + 2 * mid + 1] > fmin)
                    lo = mid + 1;
                else
                    hi = mid;
            }
            defined = min(sz, symbols + lo);
        }
        __syncwarp();
    }
    // items taken per level, top down
    int S[16];
    int T = flag[0];
    for (int j = 0; j < limit; ++j) {
        const int t = min(T, size[j]);
        int pc = 0;
        if (lane < HUF_PKW) {
            uint32_t b = W->pkbits[j][lane];
            const int lo_bit = lane * 32;
            if (t <= lo_bit)
                b = 0;
            else if (t < lo_bit + 32)
                b &= (1u << (t - lo_bit)) - 1u;
            pc = __popc(b);
        }
        const int P = (int)__reduce_add_sync(0xFFFFFFFFu, (unsigned)pc);
        S[j] = t - P;
        T = (j + 1 < limit ? flag[j + 1] : 0) + 2 * P;
    }
    for (int i = lane; i < symbols; i += 32) {
        int l = limit;
        for (int j = 0; j < limit; ++j) l -= (i < S[j]) ? 1 : 0;
        code_length[i] = (uint8_t)l;
    }
    __syncwarp();
}

// ---- getLengths (src/RawDeflate.ts:440-474): all lanes call it; freqs / lengths in shared memory ----------
__device__ void get_lengths(const uint32_t* freqs_in, int nsym, int limit, uint8_t* lengths, HufWork* W)
{
    const int lane = (int)zts_lane();
    uint32_t* freq = W->val[0];
    uint16_t* heap = reinterpret_cast<uint16_t*>(W->val[1]);
    for (int i = lane; i < nsym; i += 32) {
        freq[i] = freqs_in[i];
        lengths[i] = 0;
    }
    __syncwarp();
    int nodes = 0;
    if (lane == 0) {
        int hlen = 0;
        for (int i = 0; i < nsym; ++i)
            if (freq[i] > 0) {
                heap_push(heap, hlen, (uint16_t)i, (uint16_t)freq[i]);  // Uint16Array store: mod 65536
                nodes++;
            }
        if (nodes == 1) {
            uint16_t idx, val;
            heap_pop(heap, hlen, idx, val);
            lengths[idx] = 1;
        } else {
            for (int i = 0; i < nodes; ++i) {
                uint16_t idx, val;
                heap_pop(heap, hlen, idx, val);
                W->nidx[i] = idx;
                W->nval[i] = val;
            }
        }
    }
    nodes = __shfl_sync(0xFFFFFFFFu, nodes, 0);
    __syncwarp();
    if (nodes < 2) return;
    rpm_warp(W->nval, nodes, limit, huf_clen(W), W);
    for (int i = lane; i < nodes; i += 32) lengths[W->nidx[i]] = huf_clen(W)[i];
    __syncwarp();
}

// ---- getCodesFromLengths (src/RawDeflate.ts:580-611): canonical codes stored bit-reversed -------------
// warp-cooperative: lane L assigns the codes of length L in ascending symbol order
__device__ void codes_from_lengths(const uint8_t* lengths, int n, uint16_t* codes)
{
    const uint32_t lane = zts_lane();
    uint32_t mine = 0;  // count[lane]
    for (int i = 0; i < n; ++i) mine += (lengths[i] == lane) ? 1u : 0u;
    if (lane == 0 || lane > 16) mine = 0;
    // startCode[L] = (startCode[L-1] + count[L-1]) << 1
    uint32_t start = 0, code = 0;
    for (uint32_t L = 1; L <= 16; ++L) {
        const uint32_t cL = __shfl_sync(0xFFFFFFFFu, mine, (int)L);
        if (lane == L) start = code;
        code = (code + cL) << 1;
    }
    if (lane == 0)
        for (int i = 0; i < n; ++i)
            if (lengths[i] == 0) codes[i] = 0;
    if (lane >= 1 && lane <= 16) {
        for (int i = 0; i < n; ++i)
            if (lengths[i] == lane) {
                codes[i] = (uint16_t)(__brev(start) >> (32 - lane));
                start++;
            }
    }
    __syncwarp();
}

// ---- getTreeSymbols (src/RawDeflate.ts:341-431); serial (<= 316 lengths), result entries fit a byte ------
__device__ int tree_symbols(int hlit, const uint8_t* ll, int hdist, const uint8_t* dl, uint8_t* result, uint8_t* freqs)
{
    const int l = hlit + hdist;
    int n_result = 0, j;
    for (int i = 0; i < 19; ++i) freqs[i] = 0;
#define SRC(k) ((k) < hlit ? ll[(k)] : dl[(k) - hlit])
    for (int i = 0; i < l; i += j) {
        const uint32_t v = SRC(i);
        for (j = 1; i + j < l && SRC(i + j) == v; ++j) {
        }
        int run_length = j;
        if (v == 0) {
            if (run_length < 3) {
                while (run_length-- > 0) {
                    result[n_result++] = 0;
                    freqs[0]++;
                }
            } else {
                while (run_length > 0) {
                    int rpt = run_length < 138 ? run_length : 138;
                    if (rpt > run_length - 3 && rpt < run_length) rpt = run_length - 3;
                    if (rpt <= 10) {
                        result[n_result++] = 17;
                        result[n_result++] = (uint8_t)(rpt - 3);
                        freqs[17]++;
                    } else {
                        result[n_result++] = 18;
                        result[n_result++] = (uint8_t)(rpt - 11);
                        freqs[18]++;
                    }
                    run_length -= rpt;
                }
            }
        } else {
            result[n_result++] = (uint8_t)v;
            freqs[v]++;
            run_length--;
            if (run_length < 3) {
                while (run_length-- > 0) {
                    result[n_result++] = (uint8_t)v;
                    freqs[v]++;
                }
            } else {
                while (run_length > 0) {
                    int rpt = run_length < 6 ? run_length : 6;
                    if (rpt > run_length - 3 && rpt < run_length) rpt = run_length - 3;
                    result[n_result++] = 16;
                    result[n_result++] = (uint8_t)(rpt - 3);
                    freqs[16]++;
                    run_length -= rpt;
                }
            }
        }
    }
#undef SRC
    return n_result;
}

// LSB-first bit writer (BitStream.writeBits with reverse=true, src/Bitstream.ts:62-106): a 64-bit
// accumulator flushed a byte at a time into a zeroed buffer
struct HdrBits {
    uint8_t* buf;
    unsigned long long acc;
    uint32_t nacc;   // bits in acc
    uint32_t nbyte;  // bytes flushed
    __device__ __forceinline__ void put(uint32_t value, uint32_t nbits)
    {
        acc |= (unsigned long long)value << nacc;
        nacc += nbits;
        while (nacc >= 8) {
            buf[nbyte++] = (uint8_t)acc;
            acc >>= 8;
            nacc -= 8;
        }
    }
    __device__ __forceinline__ uint32_t finish()
    {
        const uint32_t total = nbyte * 8 + nacc;
        if (nacc) buf[nbyte] = (uint8_t)acc;
        return total;
    }
};

__constant__ uint8_t c_huff_order_enc[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
__constant__ uint8_t c_lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint8_t c_dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

// `smallest` (ZLB_MODE_SMALLEST, DYNAMIC only): the block is written as whichever of the reference's three block
// constructions is shortest for this chunk's tokens -- dynamic (:181-251), fixed (:161-173) or stored (:122-153).
__device__ void build_chunk(const uint32_t* hist_g, uint32_t chunk_flags, uint32_t chunk_len, int block_type,
                            bool smallest, ZtsChunkInfo* ci, ZtsChunkCodes* cc, HufWork* W)
{
    const unsigned lane = zts_lane();
    for (int i = (int)lane; i < HUF_MAXSYM; i += 32) {
        W->ll_len[i] = 0;
        if (i < 32) W->d_len[i] = 0;
    }
    __syncwarp();
    // the header bit string goes straight to the chunk's record in global memory (bytes past its last bit are
    // never read: the packer takes ceil(hdr_bits / 8) bytes, and the last one is zero above the last bit)
    HdrBits hb = {cc->hdr, 0ull, 0u, 0u};
    hb.put((chunk_flags & CHUNK_LAST) ? 1u : 0u, 1);  // BFINAL
    hb.put((uint32_t)block_type, 2);                  // BTYPE
    if (block_type == ZLB_FIXED) {
        // FixedHuffmanTable (src/RawDeflate.ts:26-41) = canonical code of these lengths; 5-bit distances
        for (int i = (int)lane; i < 288; i += 32) W->ll_len[i] = i <= 143 ? 8 : i <= 255 ? 9 : i <= 279 ? 7 : 8;
        if (lane < 30) W->d_len[lane] = 5;
        __syncwarp();
        if (lane == 0) W->hdr_bits = hb.finish();
    } else {
        get_lengths(hist_g, 286, 15, W->ll_len, W);        // :192
        get_lengths(hist_g + 286, 30, 7, W->d_len, W);     // :194
        int hlit, hdist, hclen;
        for (hlit = 286; hlit > 257 && W->ll_len[hlit - 1] == 0; hlit--) {
        }
        for (hdist = 30; hdist > 1 && W->d_len[hdist - 1] == 0; hdist--) {
        }
        int nsyms = 0;
        // the 19 code-length-symbol counts are staged in code_tmp (free until the codes are assigned);
        // get_lengths copies its input before it touches anything else
        uint32_t* tf = reinterpret_cast<uint32_t*>(huf_code_tmp(W));
        if (lane == 0) {
            uint8_t tf8[19];  // Uint8Array histogram (:346)
            nsyms = tree_symbols(hlit, W->ll_len, hdist, W->d_len, huf_tree_sym(W), tf8);  // :203
            for (int i = 0; i < 19; ++i) tf[i] = tf8[i];
        }
        nsyms = __shfl_sync(0xFFFFFFFFu, nsyms, 0);
        __syncwarp();
        get_lengths(tf, 19, 7, W->t_len, W);               // :204
        codes_from_lengths(W->t_len, 19, huf_code_tmp(W));
        if (lane == 0) {
            uint8_t trans[19];
            for (int i = 0; i < 19; ++i) trans[i] = W->t_len[c_huff_order_enc[i]];
            for (hclen = 19; hclen > 4 && trans[hclen - 1] == 0; hclen--) {
            }
            hb.put((uint32_t)(hlit - 257), 5);  // :214-216
            hb.put((uint32_t)(hdist - 1), 5);
            hb.put((uint32_t)(hclen - 4), 4);
            for (int i = 0; i < hclen; ++i) hb.put(trans[i], 3);
            for (int i = 0; i < nsyms; ++i) {  // :222-241
                const uint32_t code = huf_tree_sym(W)[i];
                hb.put(huf_code_tmp(W)[code], W->t_len[code]);
                if (code >= 16) {
                    const uint32_t bl = code == 16 ? 2 : code == 17 ? 3 : 7;
                    i++;
                    hb.put(huf_tree_sym(W)[i], bl);
                }
            }
            W->hdr_bits = hb.finish();
        }
    }
    __syncwarp();
    bool use_fixed = false, use_stored = false;
    uint32_t stored_bytes = 0;
    if (smallest && block_type == ZLB_DYNAMIC) {
        // byte sizes of the three encodings of this block alone (before any join marker)
        unsigned long long dyn = 0, fix = 0;
        for (int i = (int)lane; i < 286; i += 32) {
            uint32_t f = hist_g[i];
            if (i == 256) f = 1;
            const uint32_t ex = i > 256 ? c_lext[i - 257] : 0;
            dyn += (unsigned long long)f * (W->ll_len[i] + ex);
            fix += (unsigned long long)f * ((i <= 143 ? 8u : i <= 255 ? 9u : i <= 279 ? 7u : 8u) + ex);
        }
        if (lane < 30) {
            const unsigned long long f = hist_g[286 + lane];
            dyn += f * (W->d_len[lane] + c_dext[lane]);
            fix += f * (5u + c_dext[lane]);
        }
#pragma unroll
        for (int d = 16; d; d >>= 1) {
            dyn += __shfl_xor_sync(0xFFFFFFFFu, dyn, d);
            fix += __shfl_xor_sync(0xFFFFFFFFu, fix, d);
        }
        const unsigned long long dyn_bytes = (dyn + W->hdr_bits + 7) >> 3, fix_bytes = (fix + 3 + 7) >> 3;
        stored_bytes = chunk_len + 5u * ((chunk_len + 0xFFFEu) / 0xFFFFu);
        use_stored = chunk_len > 0 && stored_bytes < min(dyn_bytes, fix_bytes);
        use_fixed = !use_stored && fix_bytes < dyn_bytes;
        __syncwarp();
        if (use_fixed) {
            for (int i = (int)lane; i < 288; i += 32) W->ll_len[i] = i <= 143 ? 8 : i <= 255 ? 9 : i <= 279 ? 7 : 8;
            if (lane < 30) W->d_len[lane] = 5;
            __syncwarp();
            if (lane == 0) {
                HdrBits hf = {cc->hdr, 0ull, 0u, 0u};
                hf.put((chunk_flags & CHUNK_LAST) ? 1u : 0u, 1);
                hf.put((uint32_t)ZLB_FIXED, 2);
                W->hdr_bits = hf.finish();
            }
            __syncwarp();
        }
    }
    // code tables for the packer + exact body size
    codes_from_lengths(W->ll_len, (block_type == ZLB_FIXED || use_fixed) ? 288 : 286, huf_code_tmp(W));
    unsigned long long bits = 0;
    for (int i = (int)lane; i < 286; i += 32) {
        const uint32_t l = W->ll_len[i];
        cc->ll[i] = (uint32_t)huf_code_tmp(W)[i] | (l << 16);
        uint32_t f = hist_g[i];
        if (i == 256) f = 1;  // counted twice in the histogram, emitted once (src/LZ77.ts:127,279)
        const uint32_t ex = i > 256 ? c_lext[i - 257] : 0;
        bits += (unsigned long long)f * (l + ex);
    }
    __syncwarp();
    codes_from_lengths(W->d_len, 30, huf_code_tmp(W));
    if (lane < 30) {
        const uint32_t l = W->d_len[lane];
        cc->d[lane] = (uint32_t)huf_code_tmp(W)[lane] | (l << 16);
        bits += (unsigned long long)hist_g[286 + lane] * (l + c_dext[lane]);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) bits += __shfl_xor_sync(0xFFFFFFFFu, bits, d);
    if (lane == 0) {
        const uint32_t hbits = W->hdr_bits;
        ci->hdr_bits = hbits;
        ci->body_bits = bits;
        const unsigned long long total = hbits + bits;
        unsigned long long nbytes = (total + 7) >> 3;
        if (!(chunk_flags & CHUNK_LAST)) {
            // join: empty stored block that byte-aligns
            const uint32_t pad = (uint32_t)(nbytes * 8 - total);
            nbytes += (pad >= 3 ? 0 : 1) + 4;
        }
        ci->out_bytes = (uint32_t)nbytes;
        if (use_stored) {  // the packer copies the chunk's bytes instead; a join marker behind it is 5 whole bytes
            ci->hdr_bits = ZTS_HDR_STORED;
            ci->body_bits = 0;
            ci->out_bytes = stored_bytes + ((chunk_flags & CHUNK_LAST) ? 0u : 5u);
        }
    }
}

__global__ void __launch_bounds__(32)
huffman_build_kernel(const ZtsChunk* __restrict__ chunks, uint32_t n_chunks, const uint32_t* __restrict__ hist,
                     ZtsChunkInfo* __restrict__ info, ZtsChunkCodes* __restrict__ codes, int block_type, int smallest)
{
    __shared__ HufWork W;
    const uint32_t c = blockIdx.x;
    if (c >= n_chunks) return;
    build_chunk(hist + (size_t)c * 316, chunks[c].flags, chunks[c].len, block_type, smallest != 0, info + c, codes + c, &W);
}

// test hook: code lengths of one histogram
__global__ void __launch_bounds__(32)
huffman_lengths_kernel(const uint32_t* __restrict__ freqs, int nsym, int limit, uint8_t* __restrict__ lengths)
{
    __shared__ HufWork W;
    get_lengths(freqs, nsym, limit, W.ll_len, &W);
    for (int i = (int)threadIdx.x; i < nsym; i += 32) lengths[i] = W.ll_len[i];
}

int zts_huffman_launch(zlb_ctx* ctx, const ZtsChunk* d_chunks, uint32_t n_chunks, const uint32_t* d_hist,
                       ZtsChunkInfo* d_info, ZtsChunkCodes* d_codes, int block_type, int smallest)
{
    ZTS_LAUNCH(ctx, ZK_HUFFMAN,
               huffman_build_kernel<<<n_chunks, 32, 0, ctx->work>>>(d_chunks, n_chunks, d_hist, d_info, d_codes,
                                                                       block_type, smallest));
    return ZLB_OK;
}

int zts_huffman_lengths_debug(zlb_ctx* ctx, const uint32_t* d_freqs, int nsym, int limit, uint8_t* d_lengths)
{
    ZTS_LAUNCH(ctx, ZK_HUFFMAN, huffman_lengths_kernel<<<1, 32, 0, ctx->stream>>>(d_freqs, nsym, limit, d_lengths));
    return ZLB_OK;
}
