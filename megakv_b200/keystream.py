"""Synthetic key streams for the bench and the parity tests (numpy, vectorised).

Keys are 64-bit; a request is derived from a key the way the reference's receiver does it
(src/mega_recv.c:350,361-362): hash = high 32 bits, sig = low 32 bits.  sig 0 is the table's
empty marker, so it is mapped to 1 (the reference's unused calc_signature does the same,
src/mega_common.c:58-59).  loc 0 means "miss" to the consumer (src/mega_send.c:411-414), so
locations start at 1.

* uniform: key_i = i-th output of splitmix64 started at state `seed`.
* zipf:    rank r drawn with Gray et al.'s method, rank -> key_r.  Two generators: `Zipf` (exact pow, any numpy
           Generator) and `RefZipf`, which restates the reference's own generator bit for bit -- src/zipf.h:44-183:
           the approximate pow, the 48-bit LCG, zetan summed in ascending order -- vectorised (the LCG by its closed
           form).  BASELINE configs[2] (SURVEY.md 8(d) Config 3) is quoted on the latter.
"""
import numpy as np

from .hashindex import IEL_DT, SEL_DT

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def splitmix64(seed, first, n):
    """outputs first .. first+n-1 of splitmix64 whose state starts at `seed`"""
    with np.errstate(over="ignore"):
        idx = np.arange(first + 1, first + n + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def keys_to_requests(keys, locs=None):
    keys = np.asarray(keys, dtype=np.uint64)
    sig = (keys & np.uint64(0xFFFFFFFF)).astype(np.uint32)
    sig[sig == 0] = 1
    hsh = (keys >> np.uint64(32)).astype(np.uint32)
    sel = np.empty(len(keys), dtype=SEL_DT)
    sel["sig"], sel["hash"] = sig, hsh
    if locs is None:
        return sel
    iel = np.empty(len(keys), dtype=IEL_DT)
    iel["sig"], iel["hash"], iel["loc"] = sig, hsh, np.asarray(locs, dtype=np.uint32)
    return iel, sel


def uniform_inserts(seed, first, n):
    """(ielem[n], selem[n]) for keys first..first+n-1; loc = key index + 1"""
    return keys_to_requests(splitmix64(seed, first, n), np.arange(first + 1, first + n + 1, dtype=np.uint64))


def uniform_queries(seed, population, n, rng):
    """n searches for keys drawn uniformly from the first `population` keys of the stream"""
    idx = rng.integers(0, population, size=n, dtype=np.int64)
    return keys_to_requests(_keys_at(seed, idx)), idx


def _keys_at(seed, idx):
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (idx.astype(np.uint64) + np.uint64(1)) * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def zetan(n, theta, exact_terms=1 << 20):
    """sum_{i=1..n} i^-theta without n terms: the first `exact_terms` summed, the rest by Euler-Maclaurin
    (integral + end-point + first derivative term; relative error < 1e-12 for theta < 1)"""
    m = int(min(n, exact_terms))
    k = np.arange(1, m + 1, dtype=np.float64)
    z = float(np.sum(1.0 / np.power(k, theta)))
    if n > m:
        a, b = float(m), float(n)
        z += (b ** (1.0 - theta) - a ** (1.0 - theta)) / (1.0 - theta)          # integral of x^-theta over [m, n]
        z += 0.5 * (b ** -theta - a ** -theta)                                    # end points (f(m) is already in the sum)
        z += (-theta) * (b ** (-theta - 1.0) - a ** (-theta - 1.0)) / 12.0        # B2/2! * (f'(n) - f'(m))
    return z


class Zipf:
    """Zipf(theta) ranks in [0, n) -- J. Gray et al., SIGMOD'94 (as src/zipf.h)."""

    def __init__(self, n, theta, rng):
        assert 0.0 < theta < 1.0
        self.n, self.theta, self.rng = n, theta, rng
        k = np.arange(1, n + 1, dtype=np.float64)
        self.zetan = float(np.sum(1.0 / np.power(k, theta)))
        zeta2 = 1.0 + 0.5 ** theta
        self.alpha = 1.0 / (1.0 - theta)
        self.eta = (1.0 - (2.0 / n) ** (1.0 - theta)) / (1.0 - zeta2 / self.zetan)
        self.thres = 1.0 + 0.5 ** theta

    def ranks(self, m):
        u = self.rng.random(m)
        uz = u * self.zetan
        r = (self.n * np.power(self.eta * (u - 1.0) + 1.0, self.alpha)).astype(np.int64)
        r[uz < self.thres] = 1
        r[uz < 1.0] = 0
        return np.clip(r, 0, self.n - 1)


def zipf_queries(seed, population, n, theta, rng):
    z = Zipf(population, theta, rng)
    idx = z.ranks(n)
    return keys_to_requests(_keys_at(seed, idx)), idx


# ---- the reference's generator, src/zipf.h:26-183 ("mehcached" zipf), vectorised

_LCG_A, _LCG_C, _M48 = np.uint64(0x5DEECE66D), np.uint64(0xB), np.uint64((1 << 48) - 1)


def ref_pow_approx(a, b):
    """src/zipf.h:44-71: a ** b for b >= 0 -- the integer part of b by repeated squaring, the fractional part by scaling
    the high word of the double around the constant 1072632447 (low word cleared).  `a` array of positive doubles, `b` scalar."""
    a = np.array(a, dtype=np.float64, ndmin=1)
    whole = int(b)
    hi = (a.view(np.uint64) >> np.uint64(32)).astype(np.int64)
    hi = np.trunc((b - float(whole)) * (hi - 1072632447).astype(np.float64) + 1072632447.0).astype(np.int64)
    frac = ((hi.astype(np.uint64) & np.uint64(0xFFFFFFFF)) << np.uint64(32)).view(np.float64)
    r = np.ones_like(a)
    sq = a.copy()
    while whole:
        if whole & 1:
            r = r * sq
        sq = sq * sq
        whole >>= 1
    return r * frac


def ref_zetan(n, theta, chunk=1 << 22):
    """src/zipf.h:103-115: sum_{i=1..n} 1 / pow_approx(i, theta), added in ascending order (np.cumsum is sequential, so the
    rounding is the C loop's)"""
    total = 0.0
    for lo in range(0, n, chunk):
        hi = min(n, lo + chunk)
        terms = np.empty(hi - lo + 1, dtype=np.float64)
        terms[0] = total
        terms[1:] = 1.0 / ref_pow_approx(np.arange(lo + 1, hi + 1, dtype=np.float64), theta)
        total = float(np.cumsum(terms)[-1])
    return total


def ref_lcg_states(x0, m):
    """states 1..m of x <- (x * 0x5deece66d + 0xb) mod 2^48 (src/zipf.h:117-126) by the closed form
    x_i = a^i x0 + c (a^(i-1) + .. + 1); uint64 arithmetic wraps mod 2^64, of which 2^48 is a divisor"""
    with np.errstate(over="ignore"):
        powers = np.cumprod(np.full(m, _LCG_A, dtype=np.uint64))                    # a^1 .. a^m
        geo = np.cumsum(np.concatenate([[np.uint64(1)], powers[:-1]]))              # 1 + a + .. + a^(i-1)
        return (powers * np.uint64(x0) + _LCG_C * geo) & _M48


class RefZipf:
    """mehcached_zipf_init / mehcached_zipf_next (src/zipf.h:73-183) for 0 <= theta < 1.  ranks(m) continues the stream."""

    def __init__(self, n, theta, rand_seed, zetan=None):
        assert 0.0 <= theta < 1.0 and n > 0
        self.n, self.theta, self.state = int(n), float(theta), int(rand_seed)
        if theta > 0.0:
            self.alpha = 1.0 / (1.0 - theta)
            self.thres = 1.0 + float(ref_pow_approx(0.5, theta)[0])
            self.zetan = ref_zetan(self.n, theta) if zetan is None else float(zetan)
            zeta2 = ref_zetan(2, theta)
            self.eta = (1.0 - float(ref_pow_approx(2.0 / float(self.n), 1.0 - theta)[0])) / (1.0 - zeta2 / self.zetan)

    def ranks(self, m):
        x = ref_lcg_states(self.state, m)
        self.state = int(x[-1]) if m else self.state
        u = x.astype(np.float64) / float((1 << 48) - 1)
        if self.theta == 0.0:
            return (float(self.n) * u).astype(np.uint64)
        uz = u * self.zetan
        r = (float(self.n) * ref_pow_approx(self.eta * (u - 1.0) + 1.0, self.alpha)).astype(np.uint64)
        r[uz < self.thres] = 1
        r[uz < 1.0] = 0
        return r


def ref_zipf_queries(seed, population, n, theta, rand_seed, zetan=None):
    """n searches whose key ranks come from the reference's generator; returns (selem[n], rank[n])"""
    idx = RefZipf(population, theta, rand_seed, zetan).ranks(n).astype(np.int64)
    return keys_to_requests(_keys_at(seed, idx)), idx
