"""CPU tests of the C-ABI library: it loads, exports every symbol include/gkrb200.h declares, its host-side
transcript pieces agree with the oracle, and device entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        return subprocess.run(["nvidia-smi", "-L"], capture_output=True, timeout=20).returncode == 0
    except Exception:
        return False


def test_library_exports_every_declared_symbol():
    import gkrb200
    hdr = open(os.path.join(ROOT, "include", "gkrb200.h")).read()
    names = sorted(set(re.findall(r"\b(gkrb200_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 20
    L = gkrb200.lib()
    for n in names:
        assert hasattr(L, n), "symbol %s declared in include/gkrb200.h is not exported" % n
    out = subprocess.run(["nm", "-D", "--defined-only", gkrb200._lib.SO_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (gkrb200_[a-z0-9_]+)", out))
    assert set(names) <= exported
    assert L.gkrb200_version().startswith(b"gkrb200")
    assert L.gkrb200_proof_vec_len(10) == 10243  # 1006*bn+183 (prover/gadget/hints.go:76-116)


def test_library_is_built_for_sm100a_only():
    import gkrb200
    out = subprocess.run(["cuobjdump", "--list-elf", gkrb200._lib.SO_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_a_device():
    import gkrb200
    if _has_gpu():
        pytest.skip("a GPU is present")
    with pytest.raises(gkrb200.GkrB200Error) as e:
        gkrb200.Context(device=0, max_bn=4)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_reference_the_oracle():
    """the product path must never import/link oracle/ (it is test infrastructure)"""
    pkg = os.path.join(ROOT, "gkr-mimc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "coracle" not in txt and "pyref" not in txt and "liboracle" not in txt and "gkr_oracle" not in txt, os.path.join(dirpath, f)
    out = subprocess.run(["ldd", os.path.join(pkg, "libgkrb200.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_host_transcript_matches_oracle(oracle):
    """common.GetChallenge / hash.MimcHash and poly.InterpolateOnRange run on the host inside the product"""
    import gkrb200
    rng = np.random.default_rng(3)
    for n in (0, 1, 3, 9, 91):
        x = oracle.to_mont([int.from_bytes(rng.bytes(40), "little") % oracle.Q for _ in range(n)]).reshape(n, 4)
        assert np.array_equal(gkrb200.common.GetChallenge(x), oracle.mimc_hash(x))
    # edge residues through the hand-written ADX multiplier and the unreduced in-round products of the host MiMC chain
    Q = oracle.Q
    edge = [0, 1, 2, Q - 1, Q - 2, (1 << 256) % Q, Q // 2, Q // 2 + 1, (1 << 253) - 1, Q - (1 << 64), (1 << 64) - 1, (1 << 192)]
    for rep in range(40):
        vals = [edge[(rep * 7 + i * 5) % len(edge)] for i in range(1 + rep % 9)]
        x = oracle.to_mont(vals).reshape(len(vals), 4)
        assert np.array_equal(gkrb200.common.GetChallenge(x), oracle.mimc_hash(x)), vals
        raw = np.array([[(v >> (64 * j)) & (2**64 - 1) for j in range(4)] for v in vals], dtype=np.uint64)  # same words read as Montgomery images
        assert np.array_equal(gkrb200.common.GetChallenge(raw), oracle.mimc_hash(raw)), vals
    assert oracle.from_mont(gkrb200.common.GetChallenge(oracle.to_mont([12])))[0] == \
        1808205620575546259657963589762746470347087906694759866517376279978241663265  # hash/hash_test.go:21-27
    for n in range(1, 13):
        v = oracle.random_fr_array(n + 5)[5:]
        assert np.array_equal(gkrb200.poly.InterpolateOnRange(v), oracle.interpolate_on_range(v))


def test_montgomery_conversions_and_random_array(oracle):
    import gkrb200
    vals = [0, 1, 5, oracle.Q - 1, 2**200 + 12345]
    reg = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        for j in range(4):
            reg[i, j] = (v >> (64 * j)) & (2**64 - 1)
    m = gkrb200.common.ToMontgomery(reg)
    assert np.array_equal(m, oracle.to_mont(vals))
    assert np.array_equal(gkrb200.common.FromMontgomery(m), reg)
    assert np.array_equal(gkrb200.common.RandomFrArray(1000), oracle.random_fr_array(1000))
    assert oracle.from_mont(gkrb200.common.SetUint64([0, 7, 2**64 - 1])) == [0, 7, 2**64 - 1]


def test_argument_errors_before_any_device_work():
    import gkrb200
    L = gkrb200.lib()
    out = np.zeros(4, dtype=np.uint64)
    assert L.gkrb200_mimc_hash(None, 3, out.ctypes.data_as(ctypes.c_void_p)) == -1
    assert b"null" in L.gkrb200_last_error()
    assert L.gkrb200_interpolate(out.ctypes.data_as(ctypes.c_void_p), 13, out.ctypes.data_as(ctypes.c_void_p)) == -1
    h = ctypes.c_void_p()
    assert L.gkrb200_init(ctypes.byref(h), 0, 99, None) == -1


def test_host_verifier_pieces_match_oracle(oracle):
    """poly.EvalUnivariate, poly.EvalEq, scalar fr.Element methods and sumcheck.Verify run on the host inside the product
    (sumcheck/verifier.go:28-65, poly/lagrange.go:31-39, poly/eq.go:19-32); checked against the oracle without a GPU."""
    import gkrb200
    from gkrb200.sumcheck import _scalar
    rng = np.random.default_rng(11)
    rnd = lambda n: oracle.to_mont([int.from_bytes(rng.bytes(40), "little") % oracle.Q for _ in range(n)]).reshape(n, 4)
    for n in (1, 2, 9, 91):
        c, x = rnd(n), rnd(1)[0]
        assert np.array_equal(gkrb200.poly.EvalUnivariate(c, x), oracle.eval_univariate(c, x))
    for n in (0, 1, 5, 22):
        q, h = rnd(n), rnd(n)
        assert np.array_equal(gkrb200.poly.EvalEq(q, h), oracle.eval_eq(q, h))
    Q = oracle.Q
    edge = oracle.to_mont([0, 1, Q - 1, Q // 2, (1 << 253) - 1, 2]).reshape(-1, 4)
    for a in list(edge) + list(rnd(4)):
        for b in list(edge) + list(rnd(2)):
            assert np.array_equal(_scalar(0, a, b), oracle.fr_mul(a, b))
            assert np.array_equal(_scalar(1, a, b), oracle.fr_add(a, b))
            assert np.array_equal(_scalar(2, a, b), oracle.fr_sub(a, b))
        a7 = oracle.fr_mul(oracle.fr_mul(oracle.fr_mul(a, a), oracle.fr_mul(a, a)), oracle.fr_mul(oracle.fr_mul(a, a), a))
        assert np.array_equal(_scalar(3, a, a), a7)
        inv = _scalar(4, a, a)
        if np.any(a):
            assert np.array_equal(oracle.fr_mul(inv, a), oracle.to_mont([1]).reshape(4))
        else:
            assert not np.any(inv)
    # sumcheck.Verify on the oracle prover's transcripts: cipher gate (1 claim) and 10-claim identity (sumcheck/testing.go:11-57)
    for bn in (0, 1, 4, 7):
        n = 1 << bn
        L = oracle.to_mont(list(range(n))).reshape(n, 4)
        q = oracle.random_fr_array(bn).reshape(1, bn, 4)
        ark = oracle.to_mont([145646]).reshape(4)
        claim = oracle.evaluation(1, ark, q, None, L, L).reshape(1, 4)
        proof, chal, fin = oracle.sumcheck_prove([L, L], q, claim, 1, ark)
        ch2, final, rho = gkrb200.sumcheck.Verify(claim, proof)
        orc, och, ofin, orho = oracle.sumcheck_verify(claim, proof)
        assert orc == 0 and np.array_equal(ch2, och) and np.array_equal(final, ofin) and np.array_equal(rho, orho)
        assert np.array_equal(ch2, chal)
        if bn:
            bad = proof.copy()
            bad[bn - 1, 3, 0] ^= np.uint64(1)
            with pytest.raises(gkrb200.GkrB200Error) as e:
                gkrb200.sumcheck.Verify(claim, bad)
            assert e.value.code == -6 and "round %d" % (bn - 1) in str(e.value)
        qs = oracle.to_mont([i * j + i for i in range(10) for j in range(bn)]).reshape(10, bn, 4)
        claims = np.stack([oracle.evaluation(0, None, qs[i:i + 1], None, L) for i in range(10)])
        proof, chal, fin = oracle.sumcheck_prove([L], qs, claims, 0)
        ch2, final, rho = gkrb200.sumcheck.Verify(claims, proof)
        orc, och, ofin, orho = oracle.sumcheck_verify(claims, proof)
        assert orc == 0 and np.array_equal(ch2, och) and np.array_equal(final, ofin) and np.array_equal(rho, orho)
    L_ = gkrb200.lib()
    out = np.zeros(4, dtype=np.uint64)
    assert L_.gkrb200_eval_univariate(None, 3, out.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p)) == -1
    assert L_.gkrb200_fr_scalar(9, out.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p)) == -1
    assert L_.gkrb200_sumcheck_verify(None, 0, None, 0, 9, None, out.ctypes.data_as(ctypes.c_void_p), None) == -1


def test_header_is_plain_c(tmp_path):
    """cgo compiles include/gkrb200.h as C: it must be valid C99 on its own (no C++, torch or CUDA types) and link against the library"""
    import gkrb200
    src = tmp_path / "use.c"
    src.write_text('#include "gkrb200.h"\n#include <stdio.h>\n'
                   'int main(void) {\n'
                   '    uint64_t twelve[4] = {12, 0, 0, 0}, m[4], h[4], r[4];\n'
                   '    if (gkrb200_to_montgomery(twelve, 1, m) || gkrb200_mimc_hash(m, 1, h) || gkrb200_from_montgomery(h, 1, r)) return 1;\n'
                   '    printf("%016llx%016llx%016llx%016llx %zu %s\\n", (unsigned long long)r[3], (unsigned long long)r[2], (unsigned long long)r[1],\n'
                   '           (unsigned long long)r[0], gkrb200_proof_vec_len(22), gkrb200_version());\n'
                   '    return 0;\n}\n')
    exe = tmp_path / "use"
    lib_dir = os.path.dirname(gkrb200._lib.SO_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L" + lib_dir, "-lgkrb200", "-Wl,-rpath," + lib_dir])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    hexval, veclen, _ = out.stdout.split(" ", 2)
    assert int(hexval, 16) == 1808205620575546259657963589762746470347087906694759866517376279978241663265  # hash/hash_test.go:21-27 through C
    assert int(veclen) == 1006 * 22 + 183


def test_check_mimc_circuit_accepts_only_the_mimc_wiring():
    """gkrb200_check_mimc_circuit (circuit/circuit.go:28-44 BuildCircuit / :70-79 IsInputLayer, examples/mimc.go:10-37): the Go shim
    hands over the Circuit it was called with; anything but examples.MimcCircuit() is GKRB200_ERR_ARG with a message naming the layer."""
    import gkrb200

    class NoDevice:
        world = 1

    c = gkrb200.MimcCircuit(NoDevice())
    c.check()  # the real circuit passes
    for mutate, what in (
            (lambda c: c.layers[50].In.__setitem__(1, 48), "layer 50 input 1"),         # wrong wiring
            (lambda c: c.layers[3].In.reverse(), "layer 3 input 0"),                    # inputs swapped
            (lambda c: setattr(c.layers[2], "gate_kind", 1), "layer 2 has gate kind"),  # identity layer given a cipher gate
            (lambda c: c.layers[1].In.append(0), "layer 1 has 1 inputs"),               # an input layer with inputs
            (lambda c: c.layers.pop(), "93 layers")):                                   # wrong depth
        c = gkrb200.MimcCircuit(NoDevice())
        mutate(c)
        with pytest.raises(gkrb200.GkrB200Error) as e:
            c.check()
        assert e.value.code == -1 and what in str(e.value), str(e.value)
    c = gkrb200.MimcCircuit(NoDevice())
    arks = gkrb200.context.fr_empty(94)
    for l in range(3, 94):
        arks[l] = gkrb200.common.Ark(l - 3)
    c.check(arks)
    arks[40, 2] ^= np.uint64(1)
    with pytest.raises(gkrb200.GkrB200Error) as e:
        c.check(arks)
    assert "layer 40" in str(e.value) and "Ark" in str(e.value)
    assert gkrb200.lib().gkrb200_check_mimc_circuit(94, None, None, None, None) == -1


def test_const_mul_table_and_the_fold_product_it_drives(oracle):
    """Host side of the constant-multiplier fold (fr_mul_const, csrc/fr_device.cuh): K_i = r * 2^(32i+64) * 2^-256 mod q, and the
    device algorithm restated limb for limb on those K_i -- 64 products sum_i d_i * K_i with every row at limb 0, then two
    Montgomery rows -- yields fr.Element.Mul(r, d) for edge and random operands (the GPU parity tests check the kernel itself)."""
    import random
    import gkrb200
    Q = oracle.Q
    R = 1 << 256
    rinv = pow(R, -1, Q)
    M32 = 0xFFFFFFFF
    qinv32 = (-pow(Q, -1, 1 << 32)) % (1 << 32)
    ql = [(Q >> (32 * i)) & M32 for i in range(8)]
    rnd = random.Random(11)

    def enc(v):
        return np.array([(v >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)], dtype=np.uint64)

    def chain(acc, start, xs, y, cin=0):
        carry = cin
        for k, x in enumerate(xs):
            p = x * y
            i = start + 2 * k
            t = acc[i] + (p & M32) + carry
            acc[i], carry = t & M32, t >> 32
            t = acc[i + 1] + (p >> 32) + carry
            acc[i + 1], carry = t & M32, t >> 32
        acc[start + 2 * len(xs)] += carry
        assert acc[start + 2 * len(xs)] <= M32  # the limb a chain carries out into never wraps

    vals = [0, 1, Q - 1, Q - 2, Q // 2, (1 << 253) - 1, R % Q] + [rnd.randrange(Q) for _ in range(40)]
    for r in vals[:12] + vals[-8:]:
        out = np.zeros(64, dtype=np.uint32)
        assert gkrb200.lib().gkrb200_const_mul_table(enc(r).ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p)) == 0
        K = out.reshape(8, 8).tolist()
        for i in range(8):
            assert sum(K[i][l] << (32 * l) for l in range(8)) == r * pow(2, 32 * i + 64, Q) * rinv % Q
        for d in vals:
            dl = [(d >> (32 * i)) & M32 for i in range(8)]
            P, Qd = [0] * 12, [0] * 12
            for i in range(8):
                chain(P, 0, K[i][0::2], dl[i])
                chain(Qd, 1, K[i][1::2], dl[i])
            m = ((P[0] + Qd[0]) * qinv32) & M32
            chain(P, 0, ql[0::2], m)
            chain(Qd, 1, ql[1::2], m, cin=(P[0] + Qd[0]) >> 32)
            m = ((Qd[1] + P[1]) * qinv32) & M32
            chain(Qd, 1, ql[0::2], m)
            chain(P, 2, ql[1::2], m, cin=(Qd[1] + P[1]) >> 32)
            u = sum((P[i] + Qd[i]) << (32 * (i - 2)) for i in range(2, 12))
            assert u < 2 * Q
            assert (u - Q if u >= Q else u) == r * d * rinv % Q
