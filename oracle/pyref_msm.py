"""Python big-int restatement of the InitialRandomnessHint path (smaller piece) -- ORACLE, test infrastructure.

Follows prover/gadget/hints.go:147-192:
    KrsGkr     = MultiExp(pubKGkr, scalarsPub)            (hints.go:182)
    KrsGkrPriv = MultiExp(privKGkrSigma, scalarsPriv)     (hints.go:183)
    KrsGkr    += KrsGkrPriv                               (hints.go:184)
    initialRandomness = DeriveRandomnessFromPoint(KrsGkr) (hints.go:147-159: legacy Keccak-256 of G1Affine.RawBytes(), fr.SetBytes)
The curve arithmetic lives in the un-vendored github.com/consensys/gnark-crypto v0.6.1-0.20220110145513-493bb1c180d9, packages
ecc/bn254 (G1Affine, MultiExp, marshal.go RawBytes) and ecc/bn254/fp; its published definitions are restated here:
BN254 G1 is y^2 = x^3 + 3 over Fp, generator (1, 2), the point at infinity is encoded as (0, 0) in affine coordinates, and
RawBytes is X || Y, 32 bytes each, big-endian, regular (non-Montgomery) form, with byte 0 = 0x40 and the rest zero for infinity.
A multi-exponentiation is a group element: whatever algorithm computes it, the affine result is unique, so parity is bit-exact.
"""
P = 21888242871839275222246405745257275088696311157297823662689037894645226208583  # base field
Q = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # scalar field (fr)
B = 3
G1 = (1, 2)
INF = (0, 0)


def is_on_curve(pt):
    x, y = pt
    return pt == INF or (y * y - x * x * x - B) % P == 0


def add(p1, p2):
    """affine addition with the special cases (infinity, doubling, inverse points)"""
    if p1 == INF:
        return p2
    if p2 == INF:
        return p1
    x1, y1 = p1
    x2, y2 = p2
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return INF
        lam = 3 * x1 * x1 * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return (x3, (lam * (x1 - x3) - y1) % P)


def neg(pt):
    return pt if pt == INF else (pt[0], (-pt[1]) % P)


def mul(k, pt):
    acc, base = INF, pt
    while k:
        if k & 1:
            acc = add(acc, base)
        base = add(base, base)
        k >>= 1
    return acc


def multi_exp(points, scalars):
    """G1Affine.MultiExp (gnark-crypto ecc/bn254/multiexp.go): sum_i scalars[i] * points[i]; scalars are fr values in regular form"""
    acc = INF
    for pt, s in zip(points, scalars):
        acc = add(acc, mul(s % Q, pt))
    return acc


# ---- legacy Keccak-256 (golang.org/x/crypto/sha3 NewLegacyKeccak256: rate 136, padding 0x01 ... 0x80)
_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
       0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
       0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
       0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M = (1 << 64) - 1


def _rol(v, n):
    return ((v << n) | (v >> (64 - n))) & _M if n else v


def _f1600(a):
    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = _rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ ((~b[(x + 1) % 5][y]) & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    return a


def keccak256(data: bytes) -> bytes:
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    while len(msg) % rate:
        msg.append(0)
    msg[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            a[i % 5][i // 5] ^= int.from_bytes(msg[off + 8 * i:off + 8 * i + 8], "little")
        a = _f1600(a)
    out = b"".join(a[i % 5][i // 5].to_bytes(8, "little") for i in range(4))
    return out


def raw_bytes(pt):
    """G1Affine.RawBytes (gnark-crypto ecc/bn254/marshal.go)"""
    if pt == INF:
        return bytes([0x40]) + bytes(63)
    return pt[0].to_bytes(32, "big") + pt[1].to_bytes(32, "big")


def derive_randomness_from_point(pt):
    """prover/gadget/hints.go:147-159 -> fr value (regular form)"""
    return int.from_bytes(keccak256(raw_bytes(pt)), "big") % Q


def initial_randomness(pub_points, pub_scalars, priv_points, priv_scalars):
    """prover/gadget/hints.go:162-192 -> (KrsGkrPriv, initialRandomness)"""
    krs = multi_exp(pub_points, pub_scalars)
    krs_priv = multi_exp(priv_points, priv_scalars)
    return krs_priv, derive_randomness_from_point(add(krs, krs_priv))


# ---- G2: y^2 = x^3 + 3/(9+u) over Fp2 = Fp[u]/(u^2+1); G2Affine.MultiExp (prover/gadget/prove.go:277).  Elements are pairs (a0, a1).
def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def f2_inv(a):
    n = pow(a[0] * a[0] + a[1] * a[1], -1, P)
    return (a[0] * n % P, (-a[1] * n) % P)


def f2_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def f2_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


B2 = f2_mul((3, 0), f2_inv((9, 1)))
G2 = ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
       11559732032986387107991004021392285783925812861821192530917403151452391805634),
      (8495653923123431417604973247489272438418190587263600148770280649306958101930,
       4082367875863433681332203403145435568316851327593401208105741076214120093531))  # EIP-197 / gnark-crypto bn254.Generators
INF2 = ((0, 0), (0, 0))


def g2_is_on_curve(pt):
    x, y = pt
    return pt == INF2 or f2_mul(y, y) == f2_add(f2_mul(f2_mul(x, x), x), B2)


def g2_add(p1, p2):
    if p1 == INF2:
        return p2
    if p2 == INF2:
        return p1
    (x1, y1), (x2, y2) = p1, p2
    if x1 == x2:
        if f2_add(y1, y2) == (0, 0):
            return INF2
        lam = f2_mul(f2_mul((3, 0), f2_mul(x1, x1)), f2_inv(f2_add(y1, y1)))
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), x1), x2)
    return (x3, f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1))


def g2_neg(pt):
    return pt if pt == INF2 else (pt[0], ((-pt[1][0]) % P, (-pt[1][1]) % P))


def g2_mul(k, pt):
    acc, base = INF2, pt
    while k:
        if k & 1:
            acc = g2_add(acc, base)
        base = g2_add(base, base)
        k >>= 1
    return acc


def g2_multi_exp(points, scalars):
    acc = INF2
    for pt, s in zip(points, scalars):
        acc = g2_add(acc, g2_mul(s % Q, pt))
    return acc
