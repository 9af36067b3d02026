"""Python big-int restatement of the FFT half of the Groth16 prover -- ORACLE, test infrastructure.

Follows prover/gadget/prove.go:310-366 (computeH) over the domain built by fft.NewDomain(len(r1cs.Constraints), 1, true)
(pkg/gnark/notinternal/backend/bn254/groth16/setup.go:98).  The transform itself lives in the un-vendored
github.com/consensys/gnark-crypto v0.6.1-0.20220110145513-493bb1c180d9, package ecc/bn254/fr/fft; its published definitions are
restated here from first principles (every function below is the DEFINITION, an O(n^2) sum -- it shares nothing with the product's
in-register radix-8 passes nor with the C oracle's iterative transform):

    Domain:   Cardinality n = next power of two >= m;  Generator w = g^(2^(28 - log n)) with g the 2^28-th root of unity
              19103219067921713944291392827692070036145651957329286315305642004821462161904 (= 5^((q-1)/2^28));
              FinerGenerator u = g^(2^(28 - log n - depth)), depth = 1, so u^2 = w and u^n = -1;  coset k = u^k * <w>.
    FFT(a, DIF, coset)         natural-order input, output in bit-reversed order:   out[rev(k)] = sum_j a[j] s^j w^(jk)
    FFT(a, DIT, coset)         bit-reversed input, natural-order output:            out[k] = sum_j a[rev(j)] s^j w^(jk)
    FFTInverse(a, DIF, coset)  natural in, bit-reversed out:                        out[rev(j)] = s^-j / n * sum_k a[k] w^(-jk)
    FFTInverse(a, DIT, coset)  bit-reversed in, natural out:                        out[j] = s^-j / n * sum_k a[rev(k)] w^(-jk)
    with s = u^coset (coset = 0: s = 1).
"""
Q = 21888242871839275222246405745257275088548364400416034343698204186575808495617
ROOT_OF_UNITY = 19103219067921713944291392827692070036145651957329286315305642004821462161904
MAX_ORDER_ROOT = 28


def next_pow2(m):
    n = 1
    while n < m:
        n <<= 1
    return n


class Domain:
    """fft.NewDomain(m, depth, _)"""

    def __init__(self, m, depth=1):
        self.n = next_pow2(m)
        self.log = self.n.bit_length() - 1
        assert self.log + depth <= MAX_ORDER_ROOT
        self.depth = depth
        self.generator = pow(ROOT_OF_UNITY, 1 << (MAX_ORDER_ROOT - self.log), Q)
        self.finer_generator = pow(ROOT_OF_UNITY, 1 << (MAX_ORDER_ROOT - self.log - depth), Q)
        self.generator_inv = pow(self.generator, -1, Q)
        self.cardinality_inv = pow(self.n, -1, Q)

    def shift(self, coset):
        return pow(self.finer_generator, coset, Q) if coset else 1


def rev(i, log):
    r = 0
    for _ in range(log):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


DIT, DIF = 0, 1


def fft(dom, a, decimation, coset=0):
    n, s, w = dom.n, dom.shift(coset), dom.generator
    assert len(a) == n
    nat = list(a) if decimation == DIF else [a[rev(j, dom.log)] for j in range(n)]
    out = [sum(nat[j] * pow(s, j, Q) * pow(w, j * k, Q) for j in range(n)) % Q for k in range(n)]
    return [out[rev(i, dom.log)] for i in range(n)] if decimation == DIF else out


def fft_inverse(dom, a, decimation, coset=0):
    n, s_inv, w_inv = dom.n, pow(dom.shift(coset), -1, Q), dom.generator_inv
    assert len(a) == n
    nat = list(a) if decimation == DIF else [a[rev(k, dom.log)] for k in range(n)]
    out = [pow(s_inv, j, Q) * dom.cardinality_inv * sum(nat[k] * pow(w_inv, j * k, Q) for k in range(n)) % Q for j in range(n)]
    return [out[rev(i, dom.log)] for i in range(n)] if decimation == DIF else out


def compute_h(a, b, c, dom):
    """prover/gadget/prove.go:310-366; values in, values out (the Go code ends with FromMont: h is handed to MultiExp in regular form)"""
    n = dom.n
    pad = lambda v: list(v) + [0] * (n - len(v))
    a, b, c = pad(a), pad(b), pad(c)
    a, b, c = (fft_inverse(dom, v, DIF, 0) for v in (a, b, c))
    a, b, c = (fft(dom, v, DIT, 1) for v in (a, b, c))
    minus_two_inv = pow(Q - 2, -1, Q)
    a = [(x * y - z) * minus_two_inv % Q for x, y, z in zip(a, b, c)]
    return fft_inverse(dom, a, DIF, 1)


def compute_h_by_division(a, b, c, dom):
    """the same h from its meaning, for a SATISFIED system (a[i] * b[i] = c[i] on the whole domain): interpolate A, B, C, divide
    A*B - C by X^n - 1 exactly; h[i] = coefficient rev(i) of the quotient.  Independent of the coset and of u."""
    n = dom.n
    pad = lambda v: list(v) + [0] * (n - len(v))

    def coef(v):  # coefficients in natural order, by definition
        v = pad(v)
        return [dom.cardinality_inv * sum(v[k] * pow(dom.generator_inv, j * k, Q) for k in range(n)) % Q for j in range(n)]

    A, B, C = coef(a), coef(b), coef(c)
    prod = [0] * (2 * n)
    for i, x in enumerate(A):
        for j, y in enumerate(B):
            prod[i + j] = (prod[i + j] + x * y) % Q
    for i, z in enumerate(C):
        prod[i] = (prod[i] - z) % Q
    # divide by X^n - 1: prod = H * (X^n - 1)  =>  H_j = prod[j + n] + H_(j + n) (H has degree < n, so H_j = prod[j + n]) and the
    # low half must cancel: prod[j] = -H_j
    H = prod[n:]
    assert all((prod[j] + H[j]) % Q == 0 for j in range(n)), "a * b != c on the domain: the quotient is not a polynomial"
    return [H[rev(i, dom.log)] for i in range(n)]
