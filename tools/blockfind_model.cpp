// CPU model of block_find_kernel (zlib.kotlin_b200/csrc/zk_inflate.cu): the bit positions of a DEFLATE stream at which a
// dynamic-Huffman block header that every check accepts begins.  Test infrastructure (tests/test_foreign_inflate.py compares the
// kernel's candidates with it) and the tool the filter cascade was sized with: `blockfind_model <file> [limit]` prints how
// many bit positions survive each stage.
//
// The checks are the ones the decoder applies when it parses such a header (zk_inflate.cu parse_huffman_header, which
// follows inflate/InflateStream.kt:299-348 and 102-152), plus: BFINAL = 0 (a chunk never starts at the last block) and a
// COMPLETE code-length code (libz never writes another one; an incomplete or over-subscribed one, which the reference's
// decoder would take, only costs the entry point).
//
// build: g++ -O2 -shared -fPIC -DBFM_LIB -o libblockfind_model.so blockfind_model.cpp   (or without -DBFM_LIB for the tool)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace {

const uint8_t kOrder[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

struct Bits {
    const uint8_t *p;
    uint64_t nbits;
    uint64_t get(uint64_t at, int n) const {  // n <= 32 bits at bit position `at`; bits past the end read as zero
        uint64_t v = 0;
        const uint64_t byte = at >> 3;
        for (int i = 0; i < 8; ++i) {
            const uint64_t b = byte + i;
            if (b * 8 < nbits) v |= (uint64_t)p[b] << (8 * i);
        }
        return (v >> (at & 7)) & ((1ull << n) - 1);
    }
};

// stage reached: 0 = not BFINAL 0 / BTYPE 2, 1 = HLIT / HDIST out of range, 2 = code-length code not complete,
// 3 = code lengths do not decode, 4 = literal/length or distance code rejected, 5 = accepted
int check(const Bits &s, uint64_t at) {
    if (at + 17 + 12 + 4 > s.nbits) return 0;
    if (s.get(at, 3) != 4) return 0;  // BFINAL 0, BTYPE 10b
    const int hlit = (int)s.get(at + 3, 5) + 257, hdist = (int)s.get(at + 8, 5) + 1, hclen = (int)s.get(at + 13, 4) + 4;
    if (hlit > 286 || hdist > 30) return 1;
    int clen[19] = {0};
    uint64_t pos = at + 17;
    int kraft = 0;
    for (int i = 0; i < hclen; ++i) {
        const int l = (int)s.get(pos, 3);
        pos += 3;
        clen[kOrder[i]] = l;
        if (l) kraft += 128 >> l;
    }
    if (kraft != 128) return 2;
    // canonical table of the code-length code, 7 bits
    uint8_t tl[128], tv[128];
    {
        int per[8] = {0}, next[8];
        for (int i = 0; i < 19; ++i) per[clen[i]]++;
        per[0] = 0;
        int code = 0;
        next[0] = 0;
        for (int b = 1; b <= 7; ++b) {
            code = (code + per[b - 1]) << 1;
            next[b] = code;
        }
        for (int sym = 0; sym < 19; ++sym) {
            const int l = clen[sym];
            if (!l) continue;
            const int c = next[l]++;
            int rev = 0;
            for (int b = 0; b < l; ++b) rev |= ((c >> b) & 1) << (l - 1 - b);
            for (int idx = rev; idx < 128; idx += 1 << l) {
                tl[idx] = (uint8_t)l;
                tv[idx] = (uint8_t)sym;
            }
        }
    }
    int per[2][16];
    memset(per, 0, sizeof per);
    int used[2] = {0, 0}, eob = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const int count = pass == 0 ? hlit : hdist;
        int i = 0, prev = 0;
        while (i < count) {
            if (pos + 1 > s.nbits) return 3;
            const int look = (int)s.get(pos, 7);
            const int l = tl[look], sym = tv[look];
            pos += l;
            if (sym <= 15) {
                if (sym) {
                    per[pass][sym]++;
                    used[pass]++;
                    if (pass == 0 && i == 256) eob = 1;
                }
                prev = sym;
                ++i;
                continue;
            }
            if (sym == 16 && i == 0) return 3;
            const int nb = sym == 16 ? 2 : sym == 17 ? 3 : 7;
            const int rep = (sym == 18 ? 11 : 3) + (int)s.get(pos, nb);
            pos += nb;
            if (i + rep > count) return 3;
            const int v = sym == 16 ? prev : 0;
            if (v) {
                per[pass][v] += rep;
                used[pass] += rep;
                if (pass == 0 && i <= 256 && 256 < i + rep) eob = 1;
            }
            prev = v;
            i += rep;
        }
    }
    if (pos > s.nbits) return 3;
    if (!eob || !used[1]) return 4;
    for (int pass = 0; pass < 2; ++pass) {
        int left = 1;
        for (int b = 1; b <= 15; ++b) {
            left = (left << 1) - per[pass][b];
            if (left < 0) return 4;
        }
        if (left != 0 && used[pass] > 1) return 4;
    }
    return 5;
}


// ---- the truth, for tests: walk the stream block by block (symbols only, no output) and list where blocks start -----------
struct Canon {
    int count[16], first_code[16], first_index[16];
    int sorted[320];
    void build(const int *len, int n) {
        memset(count, 0, sizeof count);
        for (int i = 0; i < n; ++i) count[len[i]]++;
        count[0] = 0;
        int code = 0, idx = 0;
        for (int b = 1; b < 16; ++b) {
            code = (code + count[b - 1]) << 1;
            first_code[b] = code;
            first_index[b] = idx;
            idx += count[b];
        }
        int fill[16];
        for (int b = 0; b < 16; ++b) fill[b] = first_index[b];
        for (int i = 0; i < n; ++i)
            if (len[i]) sorted[fill[len[i]]++] = i;
    }
    int decode(const Bits &s, uint64_t &pos) const {
        int code = 0;
        for (int l = 1; l < 16; ++l) {
            code = (code << 1) | (int)s.get(pos++, 1);
            if (code - first_code[l] >= 0 && code - first_code[l] < count[l]) return sorted[first_index[l] + code - first_code[l]];
        }
        return -1;
    }
};

const int kLenBase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
const int kLenExtra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
const int kDistExtra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};

// block starts (bit positions) of the raw deflate data beginning at bit `from`; types[i] = BTYPE | BFINAL << 2
long long walk(const Bits &s, uint64_t from, uint64_t *starts, int *types, uint64_t cap) {
    uint64_t pos = from, k = 0;
    for (;;) {
        if (pos + 3 > s.nbits) return -1;
        const int last = (int)s.get(pos, 1), type = (int)s.get(pos + 1, 2);
        if (k < cap) {
            starts[k] = pos;
            types[k] = type | last << 2;
        }
        ++k;
        pos += 3;
        if (type == 3) return -1;
        if (type == 0) {
            pos = (pos + 7) & ~7ull;
            const uint64_t len = s.get(pos, 16);
            pos += 32 + 8 * len;
        } else {
            int ll[320] = {0}, dl[32] = {0};
            if (type == 1) {
                for (int i = 0; i < 288; ++i) ll[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
                for (int i = 0; i < 32; ++i) dl[i] = 5;
            } else {
                const int hlit = (int)s.get(pos, 5) + 257, hdist = (int)s.get(pos + 5, 5) + 1, hclen = (int)s.get(pos + 10, 4) + 4;
                pos += 14;
                int cl[19] = {0};
                for (int i = 0; i < hclen; ++i) {
                    cl[kOrder[i]] = (int)s.get(pos, 3);
                    pos += 3;
                }
                Canon c;
                c.build(cl, 19);
                int all[320 + 32] = {0};
                int i = 0;
                while (i < hlit + hdist) {  // zlib reads the two sequences as one
                    const int sym = c.decode(s, pos);
                    if (sym < 0) return -1;
                    if (sym <= 15) all[i++] = sym;
                    else {
                        const int nb = sym == 16 ? 2 : sym == 17 ? 3 : 7;
                        const int rep = (sym == 18 ? 11 : 3) + (int)s.get(pos, nb);
                        pos += nb;
                        const int v = sym == 16 ? (i ? all[i - 1] : 0) : 0;
                        for (int r = 0; r < rep && i < hlit + hdist; ++r) all[i++] = v;
                    }
                }
                for (int j = 0; j < hlit; ++j) ll[j] = all[j];
                for (int j = 0; j < hdist; ++j) dl[j] = all[hlit + j];
            }
            Canon lc, dc;
            lc.build(ll, 288);
            dc.build(dl, 32);
            for (;;) {
                if (pos > s.nbits) return -1;
                const int sym = lc.decode(s, pos);
                if (sym < 0) return -1;
                if (sym < 256) continue;
                if (sym == 256) break;
                if (sym > 285) return -1;
                pos += kLenExtra[sym - 257];
                const int d = dc.decode(s, pos);
                if (d < 0 || d > 29) return -1;
                pos += kDistExtra[d];
            }
        }
        if (last) return (long long)k;
    }
}

}  // namespace

extern "C" {
long long bfm_walk(const uint8_t *p, unsigned long long n, unsigned long long from_bit, unsigned long long *starts, int *types,
                  unsigned long long cap) {
    Bits s{p, n * 8};
    return walk(s, from_bit, (uint64_t *)starts, types, cap);
}
// first accepted bit position in [from_bit, to_bit), or -1
long long bfm_first(const uint8_t *p, unsigned long long n, unsigned long long from_bit, unsigned long long to_bit) {
    Bits s{p, n * 8};
    for (uint64_t at = from_bit; at < to_bit && at < s.nbits; ++at)
        if (check(s, at) == 5) return (long long)at;
    return -1;
}
// all accepted positions in [from_bit, to_bit) -> out (capacity cap); returns the count
long long bfm_all(const uint8_t *p, unsigned long long n, unsigned long long from_bit, unsigned long long to_bit,
                  unsigned long long *out, unsigned long long cap) {
    Bits s{p, n * 8};
    uint64_t k = 0;
    for (uint64_t at = from_bit; at < to_bit && at < s.nbits; ++at)
        if (check(s, at) == 5) {
            if (k < cap) out[k] = at;
            ++k;
        }
    return (long long)k;
}
}

#ifndef BFM_LIB
int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 1;
    std::vector<uint8_t> buf;
    uint8_t tmp[65536];
    size_t r;
    while ((r = fread(tmp, 1, sizeof tmp, f)) > 0) buf.insert(buf.end(), tmp, tmp + r);
    fclose(f);
    uint64_t limit = argc > 2 ? strtoull(argv[2], nullptr, 0) : buf.size();
    if (limit > buf.size()) limit = buf.size();
    Bits s{buf.data(), (uint64_t)buf.size() * 8};
    uint64_t stage[6] = {0};
    for (uint64_t at = 0; at < limit * 8; ++at) {
        const int st = check(s, at);
        stage[st]++;
        if (st == 5) printf("accepted at bit %llu (byte %llu + %d)\n", (unsigned long long)at, (unsigned long long)(at >> 3), (int)(at & 7));
    }
    const double tot = (double)limit * 8;
    printf("positions %.0f: fail type %.4f, hlit/hdist %.4f, precode %.4f, lengths %.5f, trees %.6f, accepted %llu\n", tot, stage[0] / tot,
           stage[1] / tot, stage[2] / tot, stage[3] / tot, stage[4] / tot, (unsigned long long)stage[5]);
    return 0;
}
#endif
