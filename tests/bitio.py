"""Tiny LSB-first bit writer used to hand-build DEFLATE test streams
(the role StreamingBitWriter plays in the reference's own tests)."""


class BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def bits(self, value, count):
        self.acc |= (value & ((1 << count) - 1)) << self.n
        self.n += count
        while self.n >= 8:
            self.out.append(self.acc & 0xFF)
            self.acc >>= 8
            self.n -= 8
        return self

    def align(self):
        if self.n % 8:
            self.bits(0, 8 - self.n % 8)
        return self

    def raw(self, data):
        assert self.n == 0
        self.out += bytes(data)
        return self

    def done(self):
        self.align()
        return bytes(self.out)


def fixed_litlen_lengths():
    return [8] * 144 + [9] * 112 + [7] * 24 + [8] * 8


def zlib_header(flevel=2, fdict=0):
    cmf = 0x78
    flg = (flevel << 6) | (fdict << 5)
    flg |= (31 - ((cmf << 8) | flg) % 31) % 31
    return bytes([cmf, flg])
