"""Synthetic-workload generator: BLS12-381 points with known discrete logarithms, in plain Python
integers (independent of both the engine and the oracle).  Used by bench.py and the full-size
property tests to build deterministic batches:  P_i = (s + i*d) * G1,  Q_i = (s' + i*d') * G2.

Generation is O(n): repeated mixed additions in Jacobian coordinates + one batch inversion.
"""
import numpy as np

from . import layout as L

Q = L.Q
R = L.R_ORDER

# generators (g1.go:25-26, g2.go:26-29)
G1 = (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
      0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1)
G2 = ((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
       0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
      (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
       0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be))


class _Fq:
    zero, one = 0, 1
    @staticmethod
    def add(a, b): return (a + b) % Q
    @staticmethod
    def sub(a, b): return (a - b) % Q
    @staticmethod
    def mul(a, b): return a * b % Q
    @staticmethod
    def sqr(a): return a * a % Q
    @staticmethod
    def inv(a): return pow(a, -1, Q)
    @staticmethod
    def is_zero(a): return a == 0


class _Fq2:
    zero, one = (0, 0), (1, 0)
    @staticmethod
    def add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
    @staticmethod
    def sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
    @staticmethod
    def mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
    @staticmethod
    def sqr(a): return ((a[0] + a[1]) * (a[0] - a[1]) % Q, 2 * a[0] * a[1] % Q)
    @staticmethod
    def inv(a):
        t = pow(a[0] * a[0] + a[1] * a[1], -1, Q)
        return (a[0] * t % Q, -a[1] * t % Q)
    @staticmethod
    def is_zero(a): return a[0] == 0 and a[1] == 0


def _dbl(F, P):
    X, Y, Z = P
    if F.is_zero(Z):
        return P
    A = F.sqr(X); B = F.sqr(Y); C = F.sqr(B)
    D = F.sub(F.sub(F.sqr(F.add(X, B)), A), C); D = F.add(D, D)
    E = F.add(F.add(A, A), A); Fv = F.sqr(E)
    X3 = F.sub(Fv, F.add(D, D))
    C8 = F.add(C, C); C8 = F.add(C8, C8); C8 = F.add(C8, C8)
    Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
    Z3 = F.mul(Y, Z); Z3 = F.add(Z3, Z3)
    return (X3, Y3, Z3)


def _madd(F, P, q):
    """Jacobian P + affine q (q finite)"""
    X1, Y1, Z1 = P
    x2, y2 = q
    if F.is_zero(Z1):
        return (x2, y2, F.one)
    Z1Z1 = F.sqr(Z1)
    U2 = F.mul(x2, Z1Z1); S2 = F.mul(F.mul(y2, Z1), Z1Z1)
    H = F.sub(U2, X1); r = F.sub(S2, Y1)
    if F.is_zero(H):
        if F.is_zero(r):
            return _dbl(F, P)
        return (F.one, F.one, F.zero)
    HH = F.sqr(H); HHH = F.mul(H, HH); V = F.mul(X1, HH)
    X3 = F.sub(F.sub(F.sqr(r), HHH), F.add(V, V))
    Y3 = F.sub(F.mul(r, F.sub(V, X3)), F.mul(Y1, HHH))
    Z3 = F.mul(Z1, H)
    return (X3, Y3, Z3)


def _mul(F, q, k):
    """k * q for affine q, Jacobian result (double-and-add)"""
    acc = (F.one, F.one, F.zero)
    for bit in bin(k)[2:] if k else "":
        acc = _dbl(F, acc)
        if bit == "1":
            acc = _madd(F, acc, q)
    return acc


def _to_affine_batch(F, pts):
    """Montgomery's trick: one inversion for all z (points at infinity -> None)"""
    zs = [p[2] for p in pts]
    pref = []
    acc = F.one
    for z in zs:
        pref.append(acc)
        if not F.is_zero(z):
            acc = F.mul(acc, z)
    inv = F.inv(acc)
    out = [None] * len(pts)
    for i in range(len(pts) - 1, -1, -1):
        X, Y, Z = pts[i]
        if F.is_zero(Z):
            continue
        zi = F.mul(inv, pref[i])
        inv = F.mul(inv, Z)
        zi2 = F.sqr(zi)
        out[i] = (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))
    return out


def _progression(F, gen, s, d, n):
    s %= R; d %= R
    start = _mul(F, gen, s)
    step = _to_affine_batch(F, [_mul(F, gen, d)])[0]
    pts = [start]
    for _ in range(n - 1):
        pts.append(_madd(F, pts[-1], step) if step is not None else pts[-1])
    return _to_affine_batch(F, pts)


_MASK = (1 << 384) - 1


def _mont_rows(vals):
    """list of canonical ints -> (len, 6) u64 Montgomery limbs"""
    buf = b"".join((v * L.MONT_R % Q).to_bytes(48, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 6).copy()


def g1_points(affine_list):
    out = np.zeros(len(affine_list), dtype=L.G1_AFFINE)
    fin = [i for i, p in enumerate(affine_list) if p is not None]
    inf = [i for i, p in enumerate(affine_list) if p is None]
    if fin:
        out["x"][fin] = _mont_rows([affine_list[i][0] for i in fin])
        out["y"][fin] = _mont_rows([affine_list[i][1] for i in fin])
    for i in inf:
        out["y"][i] = L.fp_from_int(1); out["inf"][i] = 1          # G1AffineZero = (0, 1, true), g1.go:22
    return out


def g2_points(affine_list):
    out = np.zeros(len(affine_list), dtype=L.G2_AFFINE)
    fin = [i for i, p in enumerate(affine_list) if p is not None]
    inf = [i for i, p in enumerate(affine_list) if p is None]
    if fin:
        for c in (0, 1):
            out["x"][fin, c] = _mont_rows([affine_list[i][0][c] for i in fin])
            out["y"][fin, c] = _mont_rows([affine_list[i][1][c] for i in fin])
    for i in inf:
        out["y"][i, 0] = L.fp_from_int(1); out["inf"][i] = 1        # G2AffineZero, g2.go:24
    return out


def g1_progression(s, d, n):
    """numpy G1_AFFINE array of P_i = (s + i*d) * G1, i < n"""
    return g1_points(_progression(_Fq, G1, s, d, n))


def g2_progression(s, d, n):
    """numpy G2_AFFINE array of Q_i = (s + i*d) * G2, i < n"""
    return g2_points(_progression(_Fq2, G2, s, d, n))


def g1_mul(k):
    return g1_points(_to_affine_batch(_Fq, [_mul(_Fq, G1, k % R)]))


def g2_mul(k):
    return g2_points(_to_affine_batch(_Fq2, [_mul(_Fq2, G2, k % R)]))


def g1_neg(p):
    """negate affine points (numpy G1_AFFINE): y -> Q - y on the Montgomery limbs"""
    out = p.copy()
    for i in range(out.size):
        if not out["inf"].flat[i]:
            y = L.limbs_to_int(out["y"].reshape(-1, 6)[i])
            out["y"].reshape(-1, 6)[i] = np.array(L.int_to_limbs((Q - y) % Q), dtype=np.uint64)
    return out


def g2_neg(p):
    """negate affine G2 points (numpy G2_AFFINE): both coefficients of y"""
    out = p.copy()
    ys = out["y"].reshape(-1, 2, 6)
    for i in range(out.size):
        if not out["inf"].flat[i]:
            for c in range(2):
                y = L.limbs_to_int(ys[i, c])
                ys[i, c] = np.array(L.int_to_limbs((Q - y) % Q), dtype=np.uint64)
    return out


def splitmix_scalars(seed, n):
    """n deterministic canonical scalars < r as (n, 4) u64 (splitmix64 words, reduced mod r)"""
    x = seed & 0xFFFFFFFFFFFFFFFF
    vals = []
    for _ in range(n):
        w = 0
        for j in range(4):
            x = (x + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
            z = x
            z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
            z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
            z ^= z >> 31
            w |= z << (64 * j)
        vals.append(w % R)
    buf = b"".join(v.to_bytes(32, "little") for v in vals)
    return np.frombuffer(buf, dtype=np.uint64).reshape(-1, 4).copy(), vals
