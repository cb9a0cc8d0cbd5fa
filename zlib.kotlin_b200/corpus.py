"""Deterministic synthetic corpora for the deflate / inflate benchmarks and parity tests
(SURVEY.md section 8d).  Everything derives from splitmix64 over a stated seed, so the
bytes do not depend on the numpy version.

  text_like(n, seed)      Zipf(s=1) words from a 50 000-word lowercase vocabulary,
                          sentences of 4-18 words, ". " / ".\\n" / ", " punctuation, ~2 % digits
  mixed_entropy(n, seed)  1 MiB segments cycling text-like / uniform random / 4 KiB-period
                          repeats / byte runs of 64-4096 / little-endian u32 counters /
                          base64 of random
  edge_cases()            the small parity set (empty, 1-3 B, 258/259 runs, 32 KiB and
                          64 KiB boundaries, all-zero 1 MiB, the reference's stale-window input)
"""
import numpy as np

SEED_TEXT = 0x5EED0001
SEED_MIXED = 0x5EED0002
_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(seed, n, offset=0):
    """n outputs of splitmix64 started at `seed`, skipping the first `offset`."""
    with np.errstate(over="ignore"):
        idx = np.arange(offset + 1, offset + n + 1, dtype=np.uint64)
        z = np.uint64(seed & 0xFFFFFFFFFFFFFFFF) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


_VOCAB = None


def _vocabulary():
    """50 000 words x 4 trailing separators, flattened: (bytes, start[200000], length[200000])."""
    global _VOCAB
    if _VOCAB is not None:
        return _VOCAB
    nwords = 50_000
    r = splitmix64(0xC0FFEE, nwords * 13).reshape(nwords, 13)
    wlen = (1 + r[:, 0] % np.uint64(12)).astype(np.int64)
    # letter frequencies skewed towards "etaoin..." so literals are not uniform
    alphabet = np.frombuffer(b"eeeeeeetttttaaaaooooiiiinnnnsssshhhrrrddlllccuummwwffggyyppbbvkjxqz", dtype=np.uint8)
    letters = alphabet[(r[:, 1:] % np.uint64(len(alphabet))).astype(np.int64)]
    seps = [b" ", b", ", b". ", b".\n"]
    chunks, starts, lens = [], [], []
    at = 0
    for w in range(nwords):
        word = letters[w, : wlen[w]].tobytes()
        for s in seps:
            e = word + s
            chunks.append(e)
            starts.append(at)
            lens.append(len(e))
            at += len(e)
    flat = np.frombuffer(b"".join(chunks), dtype=np.uint8)
    cdf = np.cumsum(1.0 / np.arange(1, nwords + 1))
    cdf /= cdf[-1]
    _VOCAB = (flat, np.array(starts, dtype=np.int64), np.array(lens, dtype=np.int64), cdf)
    return _VOCAB


def _text_piece(n, seed):
    flat, starts, lens, cdf = _vocabulary()
    nwords = int(n / 4.5) + 64  # mean entry length is ~6.5 with Zipf weighting; over-provision
    out = np.empty(0, dtype=np.uint8)
    offset = 0
    while out.size < n:
        r = splitmix64(seed, nwords * 2, offset).reshape(nwords, 2)
        offset += nwords * 2
        u = (r[:, 0] >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        word = np.searchsorted(cdf, u).astype(np.int64)
        k = (r[:, 1] % np.uint64(1000)).astype(np.int64)
        # separator: 4-18 word sentences -> ~1/11 sentence ends; a few commas
        sep = np.zeros(nwords, dtype=np.int64)
        sep[k < 90] = 2
        sep[k < 25] = 3
        sep[(k >= 90) & (k < 150)] = 1
        entry = word * 4 + sep
        ln = lens[entry]
        total = int(ln.sum())
        dst_start = np.cumsum(ln) - ln
        within = np.arange(total, dtype=np.int64) - np.repeat(dst_start, ln)
        piece = flat[np.repeat(starts[entry], ln) + within].copy()
        # ~2 % digits: overwrite whole words picked by the random stream with decimal digits
        digit_word = (r[:, 1] >> np.uint64(20)) % np.uint64(50) == 0
        if digit_word.any():
            mask = np.repeat(digit_word, ln) & (piece >= 97) & (piece <= 122)
            piece[mask] = 48 + (piece[mask] % 10)
        out = np.concatenate([out, piece])
    return out[:n]


def text_like(n, seed=SEED_TEXT):
    """n bytes of text-like data; generated in 16 MiB pieces re-seeded per piece."""
    piece = 16 << 20
    parts = []
    for i, off in enumerate(range(0, n, piece)):
        parts.append(_text_piece(min(piece, n - off), seed + 0x1000 * i))
    return np.concatenate(parts) if parts else np.empty(0, dtype=np.uint8)


def _random_bytes(n, seed):
    return splitmix64(seed, (n + 7) // 8).view(np.uint8)[:n].copy()


def _segment(kind, n, seed):
    if kind == 0:
        return _text_piece(n, seed)
    if kind == 1:
        return _random_bytes(n, seed)
    if kind == 2:  # 4 KiB-period repeats of a text-like tile
        tile = _text_piece(4096, seed)
        return np.tile(tile, (n + 4095) // 4096)[:n]
    if kind == 3:  # runs of one byte, 64..4096 long
        nruns = n // 64 + 2
        r = splitmix64(seed, nruns)
        run_len = (64 + (r >> np.uint64(8)) % np.uint64(4033)).astype(np.int64)
        val = (r & np.uint64(0xFF)).astype(np.uint8)
        return np.repeat(val, run_len)[:n]
    if kind == 4:  # little-endian u32 counters
        start = int(splitmix64(seed, 1)[0] % np.uint64(1 << 30))
        return np.arange(start, start + (n + 3) // 4, dtype=np.uint32).view(np.uint8)[:n].copy()
    # base64 of random
    alphabet = np.frombuffer(b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/", dtype=np.uint8)
    six = _random_bytes(n, seed) & 0x3F
    out = alphabet[six]
    out[76::77] = 10  # MIME-style line breaks
    return out


def mixed_entropy(n, seed=SEED_MIXED):
    """n bytes: 1 MiB segments cycling the six kinds of SURVEY.md section 8(d)."""
    seg = 1 << 20
    parts = []
    for i, off in enumerate(range(0, n, seg)):
        parts.append(_segment(i % 6, min(seg, n - off), seed + 0x100 * i))
    return np.concatenate(parts) if parts else np.empty(0, dtype=np.uint8)


def edge_cases(chunk=None):
    """(name, bytes) pairs of the parity edge set."""
    rnd = _random_bytes(70_000, 99).tobytes()
    txt = _text_piece(70_000, 98).tobytes()
    cases = [
        ("empty", b""),
        ("one", b"x"),
        ("two", b"xy"),
        ("three", b"xyz"),
        ("run258", b"a" * 258),
        ("run259", b"a" * 259),
        ("run260_tail", b"a" * 260 + b"b"),
        ("n32767", txt[:32767]),
        ("n32768", txt[:32768]),
        ("n32769", txt[:32769]),
        ("n65535", txt[:65535]),
        ("n65536", txt[:65536]),
        ("n65537_random", rnd[:65537]),
        ("zeros_1MiB", bytes(1 << 20)),
        ("stale_window", b"abcabc\0\0\0"),
        ("hello", b"hello world"),
        ("all_bytes", bytes(range(256)) * 3),
        ("period3", b"abc" * 5000),
        ("period_long", txt[:300] * 200),
    ]
    if chunk:
        big = text_like(chunk + 1).tobytes()
        cases += [("chunk_minus1", big[: chunk - 1]), ("chunk", big[:chunk]), ("chunk_plus1", big)]
    return cases


def tail(kind, n, rank, nbytes):
    """Last `nbytes` of the corpus a rank with index `rank` generates for size n (bench.py's per-rank seeds),
    without generating the rest: only the final pieces / segments are built."""
    shift = 0x10000 * rank
    parts, have = [], 0
    if kind == "text":
        piece = 16 << 20
        i = (n - 1) // piece
        while have < nbytes and i >= 0:
            p = _text_piece(min(piece, n - i * piece), SEED_TEXT + shift + 0x1000 * i)
            parts.insert(0, p)
            have += p.size
            i -= 1
    else:
        seg = 1 << 20
        i = (n - 1) // seg
        while have < nbytes and i >= 0:
            p = _segment(i % 6, min(seg, n - i * seg), SEED_MIXED + shift + 0x100 * i)
            parts.insert(0, p)
            have += p.size
            i -= 1
    last = np.concatenate(parts)
    return last[-nbytes:].copy()
