"""js_model.py -- independent, statement-by-statement Python model of the reference's hot path with
JavaScript semantics made explicit (TEST INFRASTRUCTURE: pins oracle/zts_oracle.c; never shipped).

Why it exists: it was written before oracle/minijs could execute the reference itself, as a second witness for
the C oracle; it is kept as an independent statement-level model (the executed reference, tests/test_refjs.py, is
what pins parity now). This file is a second, separately written restatement that keeps the reference's own
control flow -- typed-array wrap-around, `undefined`/NaN comparisons, signed shifts, the per-bit
BitStream loop, the per-key candidate arrays of LZ77 -- so that a slip in either restatement shows
up as a disagreement. It is slow (pure Python) and only used on small inputs.

Each function cites the reference lines it follows (paths under /root/reference/src).
"""

UNDEF = None  # JS `undefined`


def _u8(x):
    return x & 0xFF


def _u16(x):
    return x & 0xFFFF


def _i32(x):
    x &= 0xFFFFFFFF
    return x - (1 << 32) if x & 0x80000000 else x


# ---------------------------------------------------------------------------------------------
# BitStream (Bitstream.ts)
# ---------------------------------------------------------------------------------------------
def _reverse_table():  # Bitstream.ts:134-147
    t = []
    for i0 in range(256):
        r, s, i = i0, 7, i0 >> 1
        while i:
            r = (r << 1) | (i & 1)
            s -= 1
            i >>= 1
        t.append((r << s) & 0xFF)
    return t


REV = _reverse_table()


class BitStream:
    def __init__(self, buffer=None, pos=0):  # Bitstream.ts:19-30
        self.buffer = bytearray(buffer) if buffer is not None else bytearray(0x8000)
        self.index = pos
        self.bitindex = 0
        if len(self.buffer) * 2 <= self.index:
            raise ValueError("invalid index")
        elif len(self.buffer) <= self.index:
            self.expand()

    def expand(self):  # :35-42
        self.buffer = self.buffer + bytearray(len(self.buffer))

    @staticmethod
    def rev32(n):  # :51-56 (result is a signed int32 in JS)
        n &= 0xFFFFFFFF
        return _i32((REV[n & 0xFF] << 24) | (REV[(n >> 8) & 0xFF] << 16) | (REV[(n >> 16) & 0xFF] << 8) | REV[(n >> 24) & 0xFF])

    def write_bits(self, number, b, reverse=False):  # :62-106
        buf, index, bitindex = self.buffer, self.index, self.bitindex
        current = buf[index]
        if reverse and b > 1:
            number = (self.rev32(number) >> (32 - b)) if b > 8 else (REV[number] >> (8 - b))
        if b + bitindex < 8:
            current = (current << b) | number
            bitindex += b
        else:
            for i in range(b):
                current = (current << 1) | ((number >> (b - i - 1)) & 1)
                bitindex += 1
                if bitindex == 8:
                    bitindex = 0
                    buf[index] = REV[current & 0xFF]
                    index += 1
                    current = 0
                    if index == len(buf):
                        self.expand()
                        buf = self.buffer
        buf[index] = current & 0xFF
        self.buffer, self.index, self.bitindex = buf, index, bitindex

    def finish(self):  # :112-130
        buf, index = self.buffer, self.index
        if self.bitindex > 0:
            buf[index] = (buf[index] << (8 - self.bitindex)) & 0xFF
            buf[index] = REV[buf[index]]
            index += 1
        return bytes(buf[:index])


# ---------------------------------------------------------------------------------------------
# LZ77 (LZ77.ts)
# ---------------------------------------------------------------------------------------------
_LEN_UPPER = [3, 4, 5, 6, 7, 8, 9, 10, 12, 14, 16, 18, 22, 26, 30, 34, 42, 50, 58, 66, 82, 98, 114, 130, 162, 194, 226, 257, 258]
_LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
_LEN_BITS = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
_DIST_UPPER = [1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096,
               6144, 8192, 12288, 16384, 24576, 32768]
_DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
              4097, 6145, 8193, 12289, 16385, 24577]
_DIST_BITS = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]


def length_code(length):  # LZ77.ts:20-53: first case whose bound holds
    if length < 3:
        raise ValueError("invalid length: %d" % length)
    for k, up in enumerate(_LEN_UPPER):
        if length <= up:
            return [257 + k, length - _LEN_BASE[k], _LEN_BITS[k]]
    raise ValueError("invalid length: %d" % length)


def distance_code(dist):  # LZ77.ts:56-90
    if dist < 1:
        raise ValueError("invalid distance")
    for k, up in enumerate(_DIST_UPPER):
        if dist <= up:
            return [k, dist - _DIST_BASE[k], _DIST_BITS[k]]
    raise ValueError("invalid distance")


class LZ77:
    MIN, MAX, WINDOW = 3, 258, 0x8000  # LZ77.ts:2,5,8

    def __init__(self, data, lazy):  # :120-129
        self.input = bytes(data)
        self.cap = len(self.input) * 2       # new Uint16Array(input.length * 2)
        self.output = [0] * self.cap
        self.pos = 0
        self.prev = None
        self.skip = 0
        self.lazy = lazy
        self.freqs_litlen = [0] * 286
        self.freqs_litlen[256] = 1
        self.freqs_dist = [0] * 30
        self.compares = 0

    def write_num(self, n):  # :131-133 (stores past the end of a typed array are dropped)
        if self.pos < self.cap:
            self.output[self.pos] = _u16(n)
        self.pos += 1

    def write_match(self, m, offset):  # :135-146
        arr = length_code(m[0]) + distance_code(m[1])
        for v in arr:
            self.write_num(v)
        self.freqs_litlen[arr[0]] += 1
        self.freqs_dist[arr[3]] += 1
        self.skip = m[0] + offset - 1
        self.prev = None

    def max_match_test(self, m1, m2, length):  # :149-154
        inp = self.input
        j = length
        while j > self.MIN:
            self.compares += 1
            if inp[m1 + j - 1] != inp[m2 + j - 1]:
                return False
            j -= 1
        return True

    def search(self, position, lst):  # :157-194
        inp, n = self.input, len(self.input)
        current, match_max = lst[-1], 0
        for i in range(len(lst)):
            m = lst[len(lst) - i - 1]
            ml = self.MIN
            if match_max > self.MIN:
                if not self.max_match_test(m, position, match_max):
                    continue
                ml = match_max
            while ml < self.MAX and position + ml < n and inp[m + ml] == inp[position + ml]:
                ml += 1
                self.compares += 1
            if ml > match_max:
                current, match_max = m, ml
            if ml == self.MAX:
                break
        return (match_max, position - current)

    def encode(self):  # :196-283
        table = {}
        inp, length = self.input, len(self.input)
        position = 0
        while position < length:
            key = 0
            for i in range(self.MIN):
                if position + i == length:
                    break
                key = (key << 8) | inp[position + i]
            lst = table.setdefault(key, [])
            was = self.skip
            self.skip -= 1
            if was > 0:
                lst.append(position)
                position += 1
                continue
            while lst and position - lst[0] > self.WINDOW:
                lst.pop(0)
            if position + self.MIN >= length:
                if self.prev:
                    self.write_match(self.prev, -1)
                for i in range(position, length):
                    self.write_num(inp[i])
                    self.freqs_litlen[inp[i]] += 1
                break
            if lst:
                longest = self.search(position, lst)
                if self.prev:
                    if self.prev[0] < longest[0]:
                        t = inp[position - 1]
                        self.write_num(t)
                        self.freqs_litlen[t] += 1
                        self.write_match(longest, 0)
                    else:
                        self.write_match(self.prev, -1)
                elif longest[0] < self.lazy:
                    self.prev = longest
                else:
                    self.write_match(longest, 0)
            elif self.prev:
                self.write_match(self.prev, -1)
            else:
                t = inp[position]
                self.write_num(t)
                self.freqs_litlen[t] += 1
            lst.append(position)
            position += 1
        self.write_num(256)
        self.freqs_litlen[256] += 1
        self.output = self.output[:min(self.pos, self.cap)]  # subarray(0, pos) clamps to the buffer
        return self.output


# ---------------------------------------------------------------------------------------------
# Heap (Heap.ts): (value, index) pairs in one Uint16Array
# ---------------------------------------------------------------------------------------------
class Heap:
    def __init__(self, size):
        self.buffer = [0] * (size * 2)
        self.length = 0
        self.nodes = 0

    def push(self, index, value):  # Heap.ts:49-81
        h = self.buffer
        current = self.length
        h[self.length] = _u16(value)
        self.length += 1
        h[self.length] = _u16(index)
        self.length += 1
        self.nodes += 1
        while current > 0:
            parent = ((current - 2) >> 2) << 1
            if h[current] > h[parent]:
                h[current], h[parent] = h[parent], h[current]
                h[current + 1], h[parent + 1] = h[parent + 1], h[current + 1]
                current = parent
            else:
                break
        return self.length

    def pop(self):  # Heap.ts:88-132
        h = self.buffer
        value, index = h[0], h[1]
        self.nodes -= 1
        self.length -= 2
        h[0], h[1] = h[self.length], h[self.length + 1]
        parent = 0
        while True:
            current = 2 * parent + 2
            if current >= self.length:
                break
            if current + 2 < self.length and h[current + 2] > h[current]:
                current += 2
            if h[current] > h[parent]:
                h[parent], h[current] = h[current], h[parent]
                h[parent + 1], h[current + 1] = h[current + 1], h[parent + 1]
            else:
                break
            parent = current
        return index, value


# ---------------------------------------------------------------------------------------------
# RawDeflate (RawDeflate.ts)
# ---------------------------------------------------------------------------------------------
HUFFMAN_ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]  # RawInflate.ts:14


def _get(arr, i):
    """JS array read: out of range (or a hole) is `undefined`."""
    if i is UNDEF or i < 0 or i >= len(arr):
        return UNDEF
    return arr[i]


def _add(a, b):
    """JS a + b on number|undefined: undefined makes NaN (modelled as None)."""
    if a is UNDEF or b is UNDEF:
        return UNDEF
    return a + b


def _gt(a, b):
    """JS a > b: false when either side is NaN/undefined."""
    if a is UNDEF or b is UNDEF:
        return False
    return a > b


def reverse_package_merge(freqs, symbols, limit, diag=None):  # RawDeflate.ts:484-571
    minimum_cost = [0] * limit          # Uint16Array(limit)
    flag = [UNDEF] * limit
    code_length = [limit & 0xFF] * symbols  # Uint8Array(symbols).fill(limit)
    value = [UNDEF] * limit
    typ = [UNDEF] * limit
    current_position = [0] * limit

    def mc_set(i, v):                   # typed-array store: index -1 is a no-op, value wraps to u16
        if 0 <= i < limit:
            minimum_cost[i] = _u16(v)

    def take_package(j):                # :496-507
        x = _get(typ[j], current_position[j]) if j < limit and typ[j] is not UNDEF else UNDEF
        if x is not UNDEF and x == symbols:
            take_package(j + 1)
            take_package(j + 1)
        elif x is not UNDEF and 0 <= x < symbols:
            code_length[x] = (code_length[x] - 1) & 0xFF
        elif x is not UNDEF and diag is not None:
            diag["oob_type"] = diag.get("oob_type", 0) + 1
        if j < limit:
            current_position[j] += 1

    mc_set(limit - 1, symbols)
    excess = (1 << limit) - symbols
    half = 1 << (limit - 1)
    for j in range(limit):
        if excess < half:
            flag[j] = False
        else:
            flag[j] = True
            excess -= half
        excess <<= 1
        mc_set(limit - 2 - j, (minimum_cost[limit - 1 - j] >> 1) + symbols)
    mc_set(0, int(flag[0]))
    value[0] = [UNDEF] * minimum_cost[0]
    typ[0] = [UNDEF] * minimum_cost[0]
    for j in range(1, limit):
        if minimum_cost[j] > 2 * minimum_cost[j - 1] + int(flag[j]):
            mc_set(j, 2 * minimum_cost[j - 1] + int(flag[j]))
        value[j] = [UNDEF] * minimum_cost[j]
        typ[j] = [UNDEF] * minimum_cost[j]
    for t in range(minimum_cost[limit - 1]):
        value[limit - 1][t] = _get(freqs, t)
        typ[limit - 1][t] = t
    if flag[limit - 1]:
        if symbols > 0:
            code_length[0] = (code_length[0] - 1) & 0xFF
        current_position[limit - 1] += 1
    for j in range(limit - 2, -1, -1):
        i = 0
        nxt = current_position[j + 1]
        for t in range(minimum_cost[j]):
            weight = _add(_get(value[j + 1], nxt), _get(value[j + 1], nxt + 1))
            fi = _get(freqs, i)
            if fi is UNDEF and diag is not None:
                diag["freq_oob"] = diag.get("freq_oob", 0) + 1
            if _gt(weight, fi):
                value[j][t] = weight
                typ[j][t] = symbols
                nxt += 2
            else:
                value[j][t] = fi
                typ[j][t] = i
                i += 1
        current_position[j] = 0
        if flag[j]:
            take_package(j)
    return code_length


def get_lengths(freqs, limit, diag=None):  # RawDeflate.ts:440-474
    n = len(freqs)
    heap = Heap(2 * 286)
    length = [0] * n
    for i in range(n):
        if freqs[i] > 0:
            heap.push(i, freqs[i])
    count = heap.nodes
    if count == 1:
        length[heap.pop()[0]] = 1
        return length
    nodes, values = [], []
    for _ in range(count):
        idx, val = heap.pop()
        nodes.append(idx)
        values.append(val)
    cl = reverse_package_merge(values, len(values), limit, diag)
    for i in range(len(nodes)):
        length[nodes[i]] = cl[i]
    return length


def codes_from_lengths(lengths):  # RawDeflate.ts:580-611
    count, start = {}, {}
    for l in lengths:
        count[l] = count.get(l, 0) + 1
    code = 0
    for i in range(1, 17):
        start[i] = code
        code += count.get(i, 0)
        code <<= 1
    codes = []
    for l in lengths:
        c = start.get(l)
        if l in start:
            start[l] += 1
        r = 0
        for _ in range(l):
            r = _u16((r << 1) | (c & 1))
            c >>= 1
        codes.append(r)
    return codes


def tree_symbols(hlit, litlen_lengths, hdist, dist_lengths):  # RawDeflate.ts:341-431
    src = list(litlen_lengths[:hlit]) + list(dist_lengths[:hdist])
    result, freqs = [], [0] * 19  # freqs is a Uint8Array
    i, l = 0, len(src)
    while i < l:
        j = 1
        while i + j < l and src[i + j] == src[i]:
            j += 1
        run = j
        if src[i] == 0:
            if run < 3:
                while run > 0:
                    run -= 1
                    result.append(0)
                    freqs[0] = _u8(freqs[0] + 1)
            else:
                while run > 0:
                    rpt = run if run < 138 else 138
                    if rpt > run - 3 and rpt < run:
                        rpt = run - 3
                    if rpt <= 10:
                        result += [17, rpt - 3]
                        freqs[17] = _u8(freqs[17] + 1)
                    else:
                        result += [18, rpt - 11]
                        freqs[18] = _u8(freqs[18] + 1)
                    run -= rpt
        else:
            result.append(src[i])
            freqs[src[i]] = _u8(freqs[src[i]] + 1)
            run -= 1
            if run < 3:
                while run > 0:
                    run -= 1
                    result.append(src[i])
                    freqs[src[i]] = _u8(freqs[src[i]] + 1)
            else:
                while run > 0:
                    rpt = run if run < 6 else 6
                    if rpt > run - 3 and rpt < run:
                        rpt = run - 3
                    result += [16, rpt - 3]
                    freqs[16] = _u8(freqs[16] + 1)
                    run -= rpt
        i += j
    return result, freqs


def _fixed_table():  # RawDeflate.ts:26-41
    t = []
    for i in range(288):
        if i <= 143:
            t.append((i + 0x030, 8))
        elif i <= 255:
            t.append((i - 144 + 0x190, 9))
        elif i <= 279:
            t.append((i - 256 + 0x000, 7))
        else:
            t.append((i - 280 + 0x0C0, 8))
    return t


FIXED_TABLE = _fixed_table()


def raw_deflate(data, compression_type=2, lazy=0, prefix=b""):
    """new RawDeflate(data, {compressionType, lazy, outputBuffer, outputIndex: len(prefix)}).compress()"""
    data = bytes(data)
    if compression_type == 0:  # RawDeflate.ts:93-100,122-153
        out = bytearray(prefix)
        pos, n = 0, len(data)
        while pos < n:
            blk = data[pos:pos + 0xFFFF]
            pos += len(blk)
            out.append((1 if pos == n else 0) | (0 << 1))
            out += bytes([len(blk) & 0xFF, len(blk) >> 8, (len(blk) ^ 0xFFFF) & 0xFF, (len(blk) ^ 0xFFFF) >> 8])
            out += blk
        return bytes(out)
    buf = bytearray(max(0x8000, len(prefix) + 1))
    buf[:len(prefix)] = prefix
    stream = BitStream(buf, len(prefix))
    stream.write_bits(1, 1, True)
    if compression_type == 1:  # :161-173, 305-330
        stream.write_bits(1, 2, True)
        lz = LZ77(data, lazy)
        tok = lz.encode()
        idx = 0
        while idx < len(tok):
            lit = tok[idx]
            c, b = FIXED_TABLE[lit]
            stream.write_bits(c, b)
            if lit > 0x100:
                stream.write_bits(tok[idx + 1], tok[idx + 2], True)
                stream.write_bits(tok[idx + 3], 5)
                stream.write_bits(tok[idx + 4], tok[idx + 5], True)
                idx += 5
            elif lit == 0x100:
                break
            idx += 1
        return stream.finish()
    if compression_type != 2:
        raise ValueError("invalid compression type")
    # :181-251
    stream.write_bits(2, 2, True)
    lz = LZ77(data, lazy)
    tok = lz.encode()
    ll_len = get_lengths(lz.freqs_litlen, 15)
    ll_codes = codes_from_lengths(ll_len)
    d_len = get_lengths(lz.freqs_dist, 7)
    d_codes = codes_from_lengths(d_len)
    hlit = 286
    while hlit > 257 and ll_len[hlit - 1] == 0:
        hlit -= 1
    hdist = 30
    while hdist > 1 and d_len[hdist - 1] == 0:
        hdist -= 1
    ts_codes, ts_freqs = tree_symbols(hlit, ll_len, hdist, d_len)
    t_len = get_lengths(ts_freqs, 7)
    trans = [t_len[HUFFMAN_ORDER[i]] for i in range(19)]
    hclen = 19
    while hclen > 4 and trans[hclen - 1] == 0:
        hclen -= 1
    t_codes = codes_from_lengths(t_len)
    stream.write_bits(hlit - 257, 5, True)
    stream.write_bits(hdist - 1, 5, True)
    stream.write_bits(hclen - 4, 4, True)
    for i in range(hclen):
        stream.write_bits(trans[i], 3, True)
    i = 0
    while i < len(ts_codes):
        code = ts_codes[i]
        stream.write_bits(t_codes[code], t_len[code], True)
        if code >= 16:
            i += 1
            stream.write_bits(ts_codes[i], {16: 2, 17: 3, 18: 7}[code], True)
        i += 1
    # dynamicHuffman :262-297
    idx = 0
    while idx < len(tok):
        lit = tok[idx]
        stream.write_bits(ll_codes[lit], ll_len[lit], True)
        if lit > 256:
            stream.write_bits(tok[idx + 1], tok[idx + 2], True)
            code = tok[idx + 3]
            stream.write_bits(d_codes[code], d_len[code], True)
            stream.write_bits(tok[idx + 4], tok[idx + 5], True)
            idx += 5
        elif lit == 256:
            break
        idx += 1
    return stream.finish()


# ---------------------------------------------------------------------------------------------
# Huffman.ts / RawInflate.ts
# ---------------------------------------------------------------------------------------------
def build_huffman_table(lengths):  # Huffman.ts:8-68
    max_len, min_len = 0, float("inf")
    for l in lengths:
        if l > max_len:
            max_len = l
        if l < min_len:
            min_len = l
    size = 1 << max_len
    table = [0] * size
    bit_length, code, skip = 1, 0, 2
    while bit_length <= max_len:
        for i, l in enumerate(lengths):
            if l == bit_length:
                rev, rt = 0, code
                for _ in range(bit_length):
                    rev = (rev << 1) | (rt & 1)
                    rt >>= 1
                v = (bit_length << 16) | i
                j = rev
                while j < size:
                    table[j] = v
                    j += skip
                code += 1
        bit_length += 1
        code <<= 1
        skip <<= 1
    return table, max_len, min_len


LENGTH_CODE = _LEN_BASE + [258, 258]           # RawInflate.ts:17-22
LENGTH_EXTRA = _LEN_BITS + [0, 0]              # :25-28
DIST_CODE = _DIST_BASE                         # :31-36
DIST_EXTRA = _DIST_BITS                        # :39-42


class InflateError(Exception):
    pass


class RawInflate:
    """ADAPTIVE buffer type only (the default); output growth is modelled by a Python bytearray."""

    def __init__(self, data, index=0):
        self.input = bytes(data)
        self.ip = index
        self.out = bytearray()
        self.bitsbuf = 0
        self.bitsbuflen = 0
        self.bfinal = False

    def read_bits(self, length):  # RawInflate.ts:177-207
        if self.ip + ((length - self.bitsbuflen + 7) >> 3) >= len(self.input):
            raise InflateError("input buffer is broken")
        while self.bitsbuflen < length:
            self.bitsbuf |= self.input[self.ip] << self.bitsbuflen
            self.ip += 1
            self.bitsbuflen += 8
        octet = self.bitsbuf & ((1 << length) - 1)
        self.bitsbuf >>= length
        self.bitsbuflen -= length
        return octet

    def read_code(self, table):  # :214-246
        tab, max_len, _ = table
        while self.bitsbuflen < max_len:
            if self.ip >= len(self.input):
                break
            self.bitsbuf |= self.input[self.ip] << self.bitsbuflen
            self.ip += 1
            self.bitsbuflen += 8
        cwl = tab[self.bitsbuf & ((1 << max_len) - 1)]
        cl = cwl >> 16
        if cl > self.bitsbuflen:
            raise InflateError("invalid code length: %d" % cl)
        self.bitsbuf >>= cl
        self.bitsbuflen -= cl
        return cwl & 0xFFFF

    def decompress(self):  # :127-140
        while not self.bfinal:
            self.parse_block()
        return bytes(self.out)

    def parse_block(self):  # :145-170
        hdr = self.read_bits(3)
        if hdr & 1:
            self.bfinal = True
        hdr >>= 1
        if hdr == 0:
            self.parse_stored()
        elif hdr == 1:
            self.decode(FIXED_LITLEN, FIXED_DIST)
        elif hdr == 2:
            self.parse_dynamic()
        else:
            raise InflateError("unknown BTYPE: %d" % hdr)

    def parse_stored(self):  # :251-318
        inp, n = self.input, len(self.input)
        self.bitsbuf = 0
        self.bitsbuflen = 0
        if self.ip + 1 >= n:
            raise InflateError("invalid uncompressed block header: LEN")
        ln = inp[self.ip] | (inp[self.ip + 1] << 8)
        self.ip += 2
        if self.ip + 1 >= n:
            raise InflateError("invalid uncompressed block header: NLEN")
        self.ip += 2  # NLEN is read but `len === ~nlen` can never hold
        if self.ip + ln > n:
            raise InflateError("input buffer is broken")
        self.out += inp[self.ip:self.ip + ln]
        self.ip += ln

    def parse_dynamic(self):  # :345-400
        hlit = self.read_bits(5) + 257
        hdist = self.read_bits(5) + 1
        hclen = self.read_bits(4) + 4
        cl = [0] * 19
        for i in range(hclen):
            cl[HUFFMAN_ORDER[i]] = self.read_bits(3)
        cl_table = build_huffman_table(cl)
        total = hlit + hdist
        lt = [0] * total

        def put(i, v):
            if i < total:
                lt[i] = v

        prev, i = 0, 0
        while i < total:
            code = self.read_code(cl_table)
            if code == 16:
                rep = 3 + self.read_bits(2)
                for _ in range(rep):
                    put(i, prev)
                    i += 1
            elif code == 17:
                rep = 3 + self.read_bits(3)
                for _ in range(rep):
                    put(i, 0)
                    i += 1
                prev = 0
            elif code == 18:
                rep = 11 + self.read_bits(7)
                for _ in range(rep):
                    put(i, 0)
                    i += 1
                prev = 0
            else:
                put(i, code)
                i += 1
                prev = code
        self.decode(build_huffman_table(lt[:hlit]), build_huffman_table(lt[hlit:]))

    def decode(self, litlen, dist):  # :466-516
        out = self.out
        while True:
            code = self.read_code(litlen)
            if code == 256:
                break
            if code < 256:
                out.append(code)
                continue
            ti = code - 257
            ln = LENGTH_CODE[ti]
            if LENGTH_EXTRA[ti] > 0:
                ln += self.read_bits(LENGTH_EXTRA[ti])
            code = self.read_code(dist)
            ds = DIST_CODE[code]
            if DIST_EXTRA[code] > 0:
                ds += self.read_bits(DIST_EXTRA[code])
            for _ in range(ln):
                k = len(out) - ds
                out.append(out[k] if k >= 0 else 0)  # reading before the buffer start gives undefined -> 0
        while self.bitsbuflen >= 8:
            self.bitsbuflen -= 8
            self.ip -= 1


FIXED_LITLEN = build_huffman_table([8 if i <= 143 else 9 if i <= 255 else 7 if i <= 279 else 8 for i in range(288)])
FIXED_DIST = build_huffman_table([5] * 30)


def raw_inflate(data, index=0):
    r = RawInflate(data, index)
    out = r.decompress()
    return out, r.ip


# ---------------------------------------------------------------------------------------------
# CRC32.ts / Adler32.ts
# ---------------------------------------------------------------------------------------------
def _crc_table():  # CRC32.ts:59-69
    t = []
    for i in range(256):
        c = i
        for _ in range(8):
            c = (0xEDB88320 ^ (c >> 1)) if (c & 1) else (c >> 1)
        t.append(c & 0xFFFFFFFF)
    return t


CRC_TABLE = _crc_table()


def crc32(data, crc=0):  # CRC32.ts:25-47
    crc ^= 0xFFFFFFFF
    for b in bytes(data):
        crc = (crc >> 8) ^ CRC_TABLE[(crc ^ b) & 0xFF]
    return (crc ^ 0xFFFFFFFF) & 0xFFFFFFFF


def adler32(data, adler=1):  # Adler32.ts:28-48 (modulo every 1024 bytes)
    s1, s2 = adler & 0xFFFF, (adler >> 16) & 0xFFFF
    data = bytes(data)
    i, n = 0, len(data)
    while n > 0:
        t = 1024 if n > 1024 else n
        n -= t
        for _ in range(t):
            s1 += data[i]
            s2 += s1
            i += 1
        s1 %= 65521
        s2 %= 65521
    return ((s2 << 16) | s1) & 0xFFFFFFFF
