"""Jump-ahead for the probe stream (numpy's legacy MT19937 generator, utils.py:213-216 / 255-258 of the reference draw every
probe element from it).  The state transition of MT19937 is linear over GF(2) on 19937 bits, so the word sequence X[t]
satisfies   X[J + w] = XOR_{i : g_i = 1} X[i + w],   g(t) = t^J mod phi(t),   phi = the characteristic polynomial
(Haramoto, Matsumoto, Nishimura, Panneton, L'Ecuyer 2008).  The device kernel (mt19937_jump_bits_kernel) applies one such
polynomial per set bit of the jump distance from the table   t^(2^b) mod phi,  b = 0..MAXB-1   built here:

    python -m deflatedmlmc_schwinger_b200.mtjump        # writes data/mt19937_jump_pow2.npy  (uint32 [MAXB][624])

phi is not typed in from anywhere: it is recovered by Berlekamp-Massey from 2 * 19937 bits of the generator's own output
and checked (degree 19937, annihilates a second, independent stretch of output).  `jump_host` is the plain-numpy statement
of what the kernel does, used by the CPU tests to pin the table against np.random itself."""
import os

import numpy as np

N, M, DEG = 624, 397, 19937
MAXB = 48
TABLE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "mt19937_jump_pow2.npy")


# ---- the raw (untempered) word sequence --------------------------------------------------------------------------------
def _twist(key):
    """next 624 raw words from the current 624 (vectorised in the same three runs as the device kernel)"""
    o = key
    nw = np.empty(N, dtype=np.uint32)

    def mix(a, b):
        y = (a & np.uint32(0x80000000)) | (b & np.uint32(0x7fffffff))
        return (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), np.uint32(0x9908b0df), np.uint32(0))
    nw[0:227] = o[397:624] ^ mix(o[0:227], o[1:228])
    nw[227:454] = nw[0:227] ^ mix(o[227:454], o[228:455])
    nw[454:623] = nw[227:396] ^ mix(o[454:623], o[455:624])
    nw[623] = nw[396] ^ mix(o[623:624], nw[0:1])[0]
    return nw


def raw_words(key, count):
    """X[0 .. count) with X[0..623] = key"""
    out = [np.asarray(key, dtype=np.uint32)]
    while N * len(out) < count:
        out.append(_twist(out[-1]))
    return np.concatenate(out)[:count]


def temper(y):
    y = y ^ (y >> np.uint32(11))
    y = y ^ ((y << np.uint32(7)) & np.uint32(0x9d2c5680))
    y = y ^ ((y << np.uint32(15)) & np.uint32(0xefc60000))
    return y ^ (y >> np.uint32(18))


# ---- GF(2)[t] on Python integers (bit i = coefficient of t^i) ---------------------------------------------------------
def _berlekamp_massey(bits):
    """minimal polynomial of a binary sequence; returns (C, L): C(t) = 1 + c_1 t + ... + c_L t^L  with
    s[n] = XOR_{i=1..L} c_i s[n-i]"""
    C, B, L, m = 1, 1, 0, 1
    W = 0                       # bit i = s[n - i] for the current n (window of the sequence, most recent at bit 0)
    for n, s in enumerate(bits):
        W = (W << 1) | int(s)
        d = bin(C & W).count("1") & 1 if not hasattr(int, "bit_count") else (C & W).bit_count() & 1
        if d == 0:
            m += 1
        elif 2 * L <= n:
            T = C
            C ^= B << m
            L, B, m = n + 1 - L, T, 1
        else:
            C ^= B << m
            m += 1
    return C, L


def _reverse_bits(C, L):
    return int(bin(C)[2:].zfill(L + 1)[::-1], 2)


def char_poly():
    """phi(t), degree 19937, as a Python integer"""
    rs = np.random.RandomState(20240531)
    key = np.asarray(rs.get_state()[1], dtype=np.uint32)
    X = raw_words(key, 1 + 2 * DEG + 64 + N)
    seq = ((X[1:] >> np.uint32(7)) & np.uint32(1)).astype(np.uint8)          # any bit of the words from X[1] on
    C, L = _berlekamp_massey(seq[:2 * DEG + 64])
    if L != DEG:
        raise Exception("Berlekamp-Massey found degree %d, expected %d" % (L, DEG))
    # connection polynomial C: s[n] = XOR c_i s[n-i]  <=>  characteristic polynomial phi(t) = t^L C(1/t)
    return _reverse_bits(C, L)


def _reduce(p, phi):
    while True:
        d = p.bit_length() - 1
        if d < DEG:
            return p
        # clear the whole overflow in one pass where possible: p = hi * t^DEG + lo;  t^DEG = phi - t^DEG  (mod phi)
        hi = p >> DEG
        lo = p & ((1 << DEG) - 1)
        tail = phi ^ (1 << DEG)
        acc = lo
        t = tail
        sh = 0
        while t:
            low = (t & -t).bit_length() - 1
            sh += low
            acc ^= hi << sh
            t >>= low + 1
            sh += 1
        p = acc


def _square(p):
    b = np.frombuffer(p.to_bytes((DEG + 8) // 8 + 1, "little"), dtype=np.uint8)
    bits = np.unpackbits(b, bitorder="little")
    sp = np.zeros(2 * bits.shape[0], dtype=np.uint8)
    sp[0::2] = bits
    return int.from_bytes(np.packbits(sp, bitorder="little").tobytes(), "little")


def poly_to_words(p):
    return np.frombuffer(p.to_bytes(4 * N, "little"), dtype=np.uint32).copy()


def make_table(path=TABLE):
    phi = char_poly()
    tab = np.zeros((MAXB, N), dtype=np.uint32)
    p = 2                                   # t^(2^0)
    for b in range(MAXB):
        tab[b] = poly_to_words(p)
        p = _reduce(_square(p), phi)
    np.save(path, tab)
    return tab


# ---- general jump polynomials (host): t^D mod phi by products of the table entries -----------------------------------------
_NFFT = 1 << 16


def _words_to_bits(w):
    return np.unpackbits(np.ascontiguousarray(w, dtype=np.uint32).view(np.uint8), bitorder="little")[:DEG].astype(np.float64)


def poly_mul(a_words, b_words, phi):
    """(a * b) mod phi for polynomials packed 32 coefficients per word: the product over GF(2) = the integer convolution of
    the coefficient vectors mod 2 (FFT in float64: every sum is an integer <= 19937, exact), reduced by the sparse phi"""
    fa = np.fft.rfft(_words_to_bits(a_words), _NFFT)
    fb = np.fft.rfft(_words_to_bits(b_words), _NFFT)
    c = np.rint(np.fft.irfft(fa * fb, _NFFT)).astype(np.int64) & 1
    prod = int.from_bytes(np.packbits(c[:2 * DEG].astype(np.uint8), bitorder="little").tobytes(), "little")
    return poly_to_words(_reduce(prod, phi))


_PHI = None
_POLY_CACHE = {}


def _phi():
    global _PHI
    if _PHI is None:
        _PHI = char_poly()
    return _PHI


def poly_for_distance(D):
    """t^D mod phi (uint32[624])"""
    D = int(D)
    if D in _POLY_CACHE:
        return _POLY_CACHE[D]
    tab = table()
    out = None
    b = 0
    d = D
    while d:
        if d & 1:
            out = tab[b].copy() if out is None else poly_mul(out, tab[b], _phi())
        d >>= 1
        b += 1
    if out is None:
        out = np.zeros(N, dtype=np.uint32)
        out[0] = 1
    if len(_POLY_CACHE) < 4096:
        _POLY_CACHE[D] = out
    return out


def chunk_polys(skip_before, chunk, nchunks, end_distance):
    """rows c < nchunks: t^(skip_before + c * chunk) mod phi; last row: t^end_distance mod phi   (uint32 [nchunks + 1][624])"""
    out = np.zeros((nchunks + 1, N), dtype=np.uint32)
    if nchunks > 0:
        out[0] = poly_for_distance(skip_before)
        step = poly_for_distance(chunk)
        for c in range(1, nchunks):
            out[c] = poly_mul(out[c - 1], step, _phi())
    out[nchunks] = poly_for_distance(end_distance)
    return out


_TAB = None


def table():
    global _TAB
    if _TAB is None:
        if not os.path.isfile(TABLE):
            raise Exception("missing " + TABLE + " (python -m deflatedmlmc_schwinger_b200.mtjump)")
        _TAB = np.load(TABLE)
        assert _TAB.shape == (MAXB, N) and _TAB.dtype == np.uint32
    return _TAB


# ---- host statement of the device algorithm ------------------------------------------------------------------------------
def apply_poly(arr, gwords):
    """arr = X[o .. o+623] (every word a true word of the sequence)  ->  X[o+J .. o+J+623] for g = t^J mod phi"""
    X = raw_words(arr, DEG + N)
    gbits = np.unpackbits(gwords.view(np.uint8), bitorder="little")[:DEG]
    idx = np.nonzero(gbits)[0]
    out = np.zeros(N, dtype=np.uint32)
    for i in idx:
        out ^= X[i:i + N]
    return out


def jump_host(key, pos, skip):
    """numpy state (key[624], pos) -> (arr, p): arr = 624 consecutive raw words such that the word `skip` outputs after
    the state's next output is arr[p].  Same decomposition as the kernel: shift the window by one word when pos >= 1 (the
    low 31 bits of key[0] of a freshly seeded state are not part of the sequence), jump by the bits >= 10 of the distance
    with the table, walk the rest."""
    key = np.asarray(key, dtype=np.uint32)
    tab = table()
    if pos >= 1:
        arr = raw_words(key, N + 1)[1:]
        J = pos - 1 + skip
    else:
        arr, J = key.copy(), skip
    hi, lo = J >> 10, J & 1023
    b = 10
    while hi:
        if hi & 1:
            arr = apply_poly(arr, tab[b])
        hi >>= 1
        b += 1
    X = raw_words(arr, lo + N)
    return X[lo:lo + N], 0


if __name__ == "__main__":
    import time
    t0 = time.time()
    tab = make_table()
    print("wrote", TABLE, tab.shape, "in %.1f s" % (time.time() - t0))
