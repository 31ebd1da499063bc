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
#define HUF_LVL 576  // >= 2 * symbols
#define HUF_UNDEF 0xFFFFFFFFu
#define HUF_TUNDEF 0xFFFFu

struct HufWork {
    uint32_t val[2][HUF_LVL];       // value[j+1], value[j]
    uint16_t type[15][HUF_LVL];
    uint16_t heap[2 * HUF_MAXSYM];
    uint32_t nval[HUF_MAXSYM];
    uint16_t nidx[HUF_MAXSYM];
    uint8_t clen[HUF_MAXSYM];
    uint32_t freq[HUF_MAXSYM];
    uint32_t tree_sym[2 * (286 + 30)];
    uint16_t code_tmp[HUF_MAXSYM];
    uint8_t ll_len[HUF_MAXSYM];
    uint8_t d_len[32];
    uint8_t t_len[20];
    uint8_t hdr[ZTS_HDR_BYTES];
    uint32_t hdr_bits;
};

// ---- Heap (src/Heap.ts:49-132): (value,index) pairs in one Uint16Array --------------------------
__device__ void heap_push(uint16_t* heap, int& length, uint16_t index, uint16_t value)
{
    int current = length;
    heap[length++] = value;
    heap[length++] = index;
    while (current > 0) {
        int parent = ((current - 2) >> 2) << 1;
        if (heap[current] > heap[parent]) {
            uint16_t t = heap[current];
            heap[current] = heap[parent];
            heap[parent] = t;
            t = heap[current + 1];
            heap[current + 1] = heap[parent + 1];
            heap[parent + 1] = t;
            current = parent;
        } else {
            break;
        }
    }
}

__device__ void heap_pop(uint16_t* heap, int& length, uint16_t& index, uint16_t& value)
{
    value = heap[0];
    index = heap[1];
    length -= 2;
    heap[0] = heap[length];
    heap[1] = heap[length + 1];
    int parent = 0;
    for (;;) {
        int current = 2 * parent + 2;
        if (current >= length) break;
        if (current + 2 < length && heap[current + 2] > heap[current]) current += 2;
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

// ---- reversePackageMerge (src/RawDeflate.ts:484-571) ------------------------------------------------
// HUF_UNDEF plays JS `undefined`: undefined + x = NaN and every comparison with NaN is false.
__device__ void rpm(const uint32_t* freqs, int symbols, int limit, uint8_t* code_length, HufWork* W)
{
    uint16_t minimum_cost[16];
    int flag[16], size[16], cur[16];
    for (int i = 0; i < symbols; ++i) code_length[i] = (uint8_t)limit;
    for (int j = 0; j < 16; ++j) {
        minimum_cost[j] = 0;
        cur[j] = 0;
    }
    minimum_cost[limit - 1] = (uint16_t)symbols;
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
        if (limit - 2 - j >= 0) minimum_cost[limit - 2 - j] = (uint16_t)((minimum_cost[limit - 1 - j] >> 1) + symbols);
    }
    minimum_cost[0] = (uint16_t)flag[0];
    for (int j = 1; j < limit; ++j)
        if (minimum_cost[j] > 2 * minimum_cost[j - 1] + flag[j])
            minimum_cost[j] = (uint16_t)(2 * minimum_cost[j - 1] + flag[j]);
    for (int j = 0; j < limit; ++j) size[j] = minimum_cost[j];

    // deepest level: the symbols themselves
    {
        uint32_t* v = W->val[(limit - 1) & 1];
        uint16_t* ty = W->type[limit - 1];
        for (int t = 0; t < size[limit - 1]; ++t) {
            v[t] = t < symbols ? freqs[t] : HUF_UNDEF;
            ty[t] = (uint16_t)t;
        }
    }
    if (flag[limit - 1]) {
        if (symbols > 0) code_length[0]--;
        cur[limit - 1]++;
    }
    for (int j = limit - 2; j >= 0; --j) {
        const uint32_t* pv = W->val[(j + 1) & 1];
        uint32_t* v = W->val[j & 1];
        uint16_t* ty = W->type[j];
        const int psize = size[j + 1];
        int i = 0, next = cur[j + 1];
        for (int t = 0; t < size[j]; ++t) {
            const uint32_t a = next < psize ? pv[next] : HUF_UNDEF;
            const uint32_t b = next + 1 < psize ? pv[next + 1] : HUF_UNDEF;
            const uint32_t fi = i < symbols ? freqs[i] : HUF_UNDEF;
            const bool pkg = a != HUF_UNDEF && b != HUF_UNDEF && fi != HUF_UNDEF && (a + b) > fi;
            if (pkg) {
                v[t] = a + b;
                ty[t] = (uint16_t)symbols;
                next += 2;
            } else {
                v[t] = fi;
                ty[t] = (uint16_t)(i < 0xFFFF ? i : 0xFFFE);
                i++;
            }
        }
        cur[j] = 0;
        if (flag[j]) {
            // takePackage(j), recursion replaced by an explicit stack (:496-507)
            int stack[40], sp = 0;
            stack[sp++] = j;
            while (sp > 0) {
                const int lv = stack[--sp];
                const int cp = cur[lv];
                const uint32_t x = (cp < size[lv]) ? W->type[lv][cp] : HUF_TUNDEF;
                cur[lv]++;
                if (x == (uint32_t)symbols && lv + 1 < limit) {
                    stack[sp++] = lv + 1;
                    stack[sp++] = lv + 1;
                } else if (x < (uint32_t)symbols) {
                    code_length[x]--;
                }
            }
        }
    }
}

// ---- getLengths (src/RawDeflate.ts:440-474) ----------------------------------------------------------
__device__ void get_lengths(const uint32_t* freqs, int nsym, int limit, uint8_t* lengths, HufWork* W)
{
    int hlen = 0, nodes = 0;
    for (int i = 0; i < nsym; ++i) lengths[i] = 0;
    for (int i = 0; i < nsym; ++i)
        if (freqs[i] > 0) {
            heap_push(W->heap, hlen, (uint16_t)i, (uint16_t)freqs[i]);  // Uint16Array store: mod 65536
            nodes++;
        }
    if (nodes == 0) return;
    if (nodes == 1) {
        uint16_t idx, val;
        heap_pop(W->heap, hlen, idx, val);
        lengths[idx] = 1;
        return;
    }
    for (int i = 0; i < nodes; ++i) {
        uint16_t idx, val;
        heap_pop(W->heap, hlen, idx, val);
        W->nidx[i] = idx;
        W->nval[i] = val;
    }
    rpm(W->nval, nodes, limit, W->clen, W);
    for (int i = 0; i < nodes; ++i) lengths[W->nidx[i]] = W->clen[i];
}

// ---- getCodesFromLengths (src/RawDeflate.ts:580-611): canonical codes stored bit-reversed -------------
__device__ void codes_from_lengths(const uint8_t* lengths, int n, uint16_t* codes)
{
    uint32_t count[17], start_code[17];
    for (int i = 0; i <= 16; ++i) count[i] = 0;
    for (int i = 0; i < n; ++i) count[lengths[i]]++;
    uint32_t code = 0;
    for (int i = 1; i <= 16; ++i) {
        start_code[i] = code;
        code += count[i];
        code <<= 1;
    }
    for (int i = 0; i < n; ++i) {
        const uint32_t l = lengths[i];
        uint32_t r = 0;
        if (l) {
            const uint32_t c = start_code[l]++;
            r = __brev(c) >> (32 - l);
        }
        codes[i] = (uint16_t)r;
    }
}

// ---- getTreeSymbols (src/RawDeflate.ts:341-431) ----------------------------------------------------------
__device__ int tree_symbols(int hlit, const uint8_t* ll, int hdist, const uint8_t* dl, uint32_t* result, uint8_t* freqs)
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
            result[n_result++] = v;
            freqs[v]++;
            run_length--;
            if (run_length < 3) {
                while (run_length-- > 0) {
                    result[n_result++] = v;
                    freqs[v]++;
                }
            } else {
                while (run_length > 0) {
                    int rpt = run_length < 6 ? run_length : 6;
                    if (rpt > run_length - 3 && rpt < run_length) rpt = run_length - 3;
                    result[n_result++] = 16;
                    result[n_result++] = (uint32_t)(rpt - 3);
                    freqs[16]++;
                    run_length -= rpt;
                }
            }
        }
    }
#undef SRC
    return n_result;
}

// LSB-first bit append (BitStream.writeBits with reverse=true, src/Bitstream.ts:62-106)
__device__ __forceinline__ void put_bits(uint8_t* buf, uint32_t& bitpos, uint32_t value, uint32_t nbits)
{
    for (uint32_t k = 0; k < nbits; ++k) {
        if ((value >> k) & 1u) buf[bitpos >> 3] |= (uint8_t)(1u << (bitpos & 7));
        bitpos++;
    }
}

__constant__ uint8_t c_huff_order_enc[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
__constant__ uint8_t c_lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
__constant__ uint8_t c_dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

__device__ void build_chunk(const uint32_t* hist_g, uint32_t chunk_flags, int block_type, ZtsChunkInfo* ci,
                            ZtsChunkCodes* cc, HufWork* W)
{
    const unsigned lane = zts_lane();
    for (int i = (int)lane; i < ZTS_HDR_BYTES; i += 32) W->hdr[i] = 0;
    for (int i = (int)lane; i < HUF_MAXSYM; i += 32) {
        W->ll_len[i] = 0;
        if (i < 32) W->d_len[i] = 0;
    }
    __syncwarp();
    if (lane == 0) {
        uint32_t bp = 0;
        put_bits(W->hdr, bp, (chunk_flags & CHUNK_LAST) ? 1u : 0u, 1);  // BFINAL
        put_bits(W->hdr, bp, (uint32_t)block_type, 2);                  // BTYPE
        if (block_type == ZLB_FIXED) {
            // FixedHuffmanTable (src/RawDeflate.ts:26-41) = canonical code of these lengths; 5-bit distances
            for (int i = 0; i < 288; ++i) W->ll_len[i] = i <= 143 ? 8 : i <= 255 ? 9 : i <= 279 ? 7 : 8;
            for (int i = 0; i < 30; ++i) W->d_len[i] = 5;
        } else {
            for (int i = 0; i < 286; ++i) W->freq[i] = hist_g[i];
            get_lengths(W->freq, 286, 15, W->ll_len, W);  // :192
            for (int i = 0; i < 30; ++i) W->freq[i] = hist_g[286 + i];
            get_lengths(W->freq, 30, 7, W->d_len, W);     // :194
            int hlit, hdist, hclen;
            for (hlit = 286; hlit > 257 && W->ll_len[hlit - 1] == 0; hlit--) {
            }
            for (hdist = 30; hdist > 1 && W->d_len[hdist - 1] == 0; hdist--) {
            }
            uint8_t tf8[19];
            const int nsyms = tree_symbols(hlit, W->ll_len, hdist, W->d_len, W->tree_sym, tf8);  // :203
            for (int i = 0; i < 19; ++i) W->freq[i] = tf8[i];  // Uint8Array histogram (:346)
            get_lengths(W->freq, 19, 7, W->t_len, W);          // :204
            uint8_t trans[19];
            for (int i = 0; i < 19; ++i) trans[i] = W->t_len[c_huff_order_enc[i]];
            for (hclen = 19; hclen > 4 && trans[hclen - 1] == 0; hclen--) {
            }
            codes_from_lengths(W->t_len, 19, W->code_tmp);
            put_bits(W->hdr, bp, (uint32_t)(hlit - 257), 5);  // :214-216
            put_bits(W->hdr, bp, (uint32_t)(hdist - 1), 5);
            put_bits(W->hdr, bp, (uint32_t)(hclen - 4), 4);
            for (int i = 0; i < hclen; ++i) put_bits(W->hdr, bp, trans[i], 3);
            for (int i = 0; i < nsyms; ++i) {  // :222-241
                const uint32_t code = W->tree_sym[i];
                put_bits(W->hdr, bp, W->code_tmp[code], W->t_len[code]);
                if (code >= 16) {
                    const uint32_t bl = code == 16 ? 2 : code == 17 ? 3 : 7;
                    i++;
                    put_bits(W->hdr, bp, W->tree_sym[i], bl);
                }
            }
        }
        W->hdr_bits = bp;
    }
    __syncwarp();
    // code tables for the packer + exact body size
    if (lane == 0) codes_from_lengths(W->ll_len, block_type == ZLB_FIXED ? 288 : 286, W->code_tmp);
    __syncwarp();
    unsigned long long bits = 0;
    for (int i = (int)lane; i < 286; i += 32) {
        const uint32_t l = W->ll_len[i];
        cc->ll[i] = (uint32_t)W->code_tmp[i] | (l << 16);
        uint32_t f = hist_g[i];
        if (i == 256) f = 1;  // counted twice in the histogram, emitted once (src/LZ77.ts:127,279)
        const uint32_t ex = i > 256 ? c_lext[i - 257] : 0;
        bits += (unsigned long long)f * (l + ex);
    }
    __syncwarp();
    if (lane == 0) codes_from_lengths(W->d_len, 30, W->code_tmp);
    __syncwarp();
    if (lane < 30) {
        const uint32_t l = W->d_len[lane];
        cc->d[lane] = (uint32_t)W->code_tmp[lane] | (l << 16);
        bits += (unsigned long long)hist_g[286 + lane] * (l + c_dext[lane]);
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) bits += __shfl_xor_sync(0xFFFFFFFFu, bits, d);
    for (int i = (int)lane; i < ZTS_HDR_BYTES; i += 32) cc->hdr[i] = W->hdr[i];
    if (lane == 0) {
        const uint32_t hb = W->hdr_bits;
        ci->hdr_bits = hb;
        ci->body_bits = bits;
        const unsigned long long total = hb + bits;
        unsigned long long nbytes = (total + 7) >> 3;
        if (!(chunk_flags & CHUNK_LAST)) {
            // join: empty stored block that byte-aligns
            const uint32_t pad = (uint32_t)(nbytes * 8 - total);
            nbytes += (pad >= 3 ? 0 : 1) + 4;
        }
        ci->out_bytes = (uint32_t)nbytes;
    }
}

__global__ void __launch_bounds__(32)
huffman_build_kernel(const ZtsChunk* __restrict__ chunks, uint32_t n_chunks, const uint32_t* __restrict__ hist,
                     ZtsChunkInfo* __restrict__ info, ZtsChunkCodes* __restrict__ codes, int block_type)
{
    extern __shared__ __align__(16) unsigned char hsm[];
    HufWork* W = reinterpret_cast<HufWork*>(hsm);
    const uint32_t c = blockIdx.x;
    if (c >= n_chunks) return;
    build_chunk(hist + (size_t)c * 316, chunks[c].flags, block_type, info + c, codes + c, W);
}

// test hook: code lengths of one histogram
__global__ void __launch_bounds__(32)
huffman_lengths_kernel(const uint32_t* __restrict__ freqs, int nsym, int limit, uint8_t* __restrict__ lengths)
{
    extern __shared__ __align__(16) unsigned char hsm[];
    HufWork* W = reinterpret_cast<HufWork*>(hsm);
    if (threadIdx.x == 0) {
        for (int i = 0; i < nsym; ++i) W->freq[i] = freqs[i];
        get_lengths(W->freq, nsym, limit, W->ll_len, W);
        for (int i = 0; i < nsym; ++i) lengths[i] = W->ll_len[i];
    }
}

int zts_huffman_launch(zlb_ctx* ctx, const ZtsChunk* d_chunks, uint32_t n_chunks, const uint32_t* d_hist,
                       ZtsChunkInfo* d_info, ZtsChunkCodes* d_codes, int block_type)
{
    ZTS_CUDA(ctx, cudaFuncSetAttribute(huffman_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(HufWork)));
    ZTS_LAUNCH(ctx, ZK_HUFFMAN,
               huffman_build_kernel<<<n_chunks, 32, sizeof(HufWork), ctx->stream>>>(d_chunks, n_chunks, d_hist, d_info,
                                                                                    d_codes, block_type));
    return ZLB_OK;
}

int zts_huffman_lengths_debug(zlb_ctx* ctx, const uint32_t* d_freqs, int nsym, int limit, uint8_t* d_lengths)
{
    ZTS_CUDA(ctx, cudaFuncSetAttribute(huffman_lengths_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)sizeof(HufWork)));
    ZTS_LAUNCH(ctx, ZK_HUFFMAN,
               huffman_lengths_kernel<<<1, 32, sizeof(HufWork), ctx->stream>>>(d_freqs, nsym, limit, d_lengths));
    return ZLB_OK;
}
