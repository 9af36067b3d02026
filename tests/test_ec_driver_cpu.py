"""The Groth16-side library's OWN host driver on the CPU (gkr-mimc_b200/csrc/ec/ec.cu compiled by tests/emu/ec_hostbuild.cpp against a
stand-in for the CUDA runtime) driven by the DEVICE parity tests' own bodies (tests/test_zz*_gpu.py), against the same oracles.

Why: libgkrb200ec.so was written after the round's GPU budget was spent.  The kernel bodies are checked by tests/test_msm_cpu.py /
test_ntt_cpu.py / test_groth16_cpu.py; this file adds the ~700 lines of host driver around them -- staging buffers, grow-only
workspaces, base slots and their point kinds, error codes and messages, statistics, the resident fft domain, h staying "on the device"
between computeH and the multi-exponentiation, the Groth16 sequencing through the CUDA operations backend -- through the same C ABI
and the same Python mirror the GPU tests use.  The stand-in poisons every allocation and checks guard bands around it on every
synchronize / free.  TEST INFRASTRUCTURE: the host build lives under tests/emu/_build, is never loaded by the product package
(gkrb200.ec loads gkr-mimc_b200/libgkrb200ec.so only) and is swapped in here for the duration of this module alone.
"""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu")
SO = os.path.join(EMU, "_build", "libgkrb200ec_hostbuild.so")
EC_DIR = os.path.join(ROOT, "gkr-mimc_b200", "csrc", "ec")
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build_hostbuild():
    deps = [os.path.join(EMU, "ec_hostbuild.cpp"), os.path.join(EMU, "shim", "cuda_runtime.h"), os.path.join(ROOT, "include", "gkrb200_ec.h")] + \
           [os.path.join(EC_DIR, f) for f in os.listdir(EC_DIR)]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        os.makedirs(os.path.dirname(SO), exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-Wextra", "-Werror", "-Wno-unknown-pragmas", "-Wno-unused-function",
                               "-I", os.path.join(EMU, "shim"), "-o", SO, os.path.join(EMU, "ec_hostbuild.cpp")])
    return SO


@pytest.fixture(scope="module", autouse=True)
def host_driver():
    """swap the host build in for gkrb200.ec's library for this module only"""
    from gkrb200 import ec
    saved = ec._lib
    ec._lib = ec._bind(ctypes.CDLL(build_hostbuild()))
    yield ec
    ec._lib = saved


@pytest.fixture(scope="module")
def cmsm():
    import cmsm as m
    m.build()
    return m


@pytest.fixture(scope="module")
def cfft():
    import cfft as m
    m.build()
    return m


@pytest.fixture(scope="module")
def ecx(host_driver):
    c = host_driver.EcContext(device=0)
    yield c
    c.close()


import test_zz2_msm_gpu as gpu_msm  # noqa: E402  (the GPU tests' bodies; their module-level gpu mark does not travel with the functions)
import test_zz1_ntt_gpu as gpu_ntt  # noqa: E402
import test_zz3_groth16_gpu as gpu_g16  # noqa: E402


def test_g1_add(cmsm, ecx):
    gpu_msm.test_g1_add_on_the_device(cmsm, ecx)


@pytest.mark.parametrize("n", [0, 1, 2, 7, 64, 300])
def test_multiexp(cmsm, ecx, n):
    gpu_msm.test_multiexp_matches_oracle(cmsm, ecx, n)


def test_initial_randomness_hint(cmsm, ecx):
    gpu_msm.test_initial_randomness_hint_matches_oracle(cmsm, ecx)


def test_multiexp_argument_errors(cmsm, ecx):
    gpu_msm.test_multiexp_argument_errors(cmsm, ecx)


def test_multiexp_scalars_in_device_memory(cmsm, ecx, host_driver):
    """gkrb200ec_g1_multiexp_device: with the stand-in runtime a "device" pointer is a host pointer"""
    n = 200
    pts = cmsm.gen_points(n)
    vals = [(i * 0x9E3779B97F4A7C15 + 12345) % cmsm.Q for i in range(n)]
    ecx.SetBases(4, pts)
    mont = np.ascontiguousarray(cmsm.scalars_mont(vals))
    keep = mont.copy()
    got = ecx.MultiExpDevice(4, mont.ctypes.data, n, host_driver.SCALARS_MONTGOMERY)
    assert np.array_equal(got, cmsm.multiexp(pts, cmsm.scalars_regular(vals))) and np.array_equal(mont, keep)


def test_result_self_check_rejects_points_off_the_curve(cmsm, ecx, host_driver):
    """every point the device hands back is checked against the curve equation on the host: garbage in (or a device fault) is an error,
    not an output"""
    pts = cmsm.gen_points(4)
    bad = pts.copy()
    bad[2, 0] ^= np.uint64(2)  # x of one base point disturbed: (x', y) is not on y^2 = x^3 + 3
    sc = cmsm.scalars_regular([3, 5, 7, 11])
    assert np.array_equal(ecx.MultiExpPoints(pts, sc), cmsm.multiexp(pts, sc))
    with pytest.raises(host_driver.GkrB200EcError) as e:
        ecx.MultiExpPoints(bad, sc)
    assert e.value.code == -2 and "not on the curve" in str(e.value)
    with pytest.raises(host_driver.GkrB200EcError):
        ecx.Add(bad[2], pts[0])
    g2 = cmsm.g2_gen_points(3)
    bad2 = g2.copy()
    bad2[1, 5] ^= np.uint64(1)
    with pytest.raises(host_driver.GkrB200EcError) as e:
        ecx.MultiExpPointsG2(bad2, sc[:3])
    assert "not on the curve" in str(e.value)
    assert np.array_equal(ecx.MultiExpPointsG2(g2, sc[:3]), cmsm.g2_multiexp(g2, sc[:3]))  # the context survives
    # with a zero scalar on the bad point the sum never touches it: the result is a valid point and is returned
    sc0 = cmsm.scalars_regular([3, 5, 0, 11])
    assert np.array_equal(ecx.MultiExpPoints(bad, sc0), cmsm.multiexp(pts, sc0))


def test_g2_add(cmsm, ecx):
    gpu_msm.test_g2_add_on_the_device(cmsm, ecx)


@pytest.mark.parametrize("n", [0, 1, 7, 64])
def test_g2_multiexp(cmsm, ecx, n):
    gpu_msm.test_g2_multiexp_matches_oracle(cmsm, ecx, n)


@pytest.mark.parametrize("log", [0, 1, 2, 3, 4, 5, 6, 10, 11, 12])
def test_fft_variants(cfft, ecx, log):
    gpu_ntt.test_fft_variants_match_oracle(cfft, ecx, log)


@pytest.mark.parametrize("m", [1, 2, 5, 1000, 4096])
def test_compute_h(cfft, ecx, m):
    gpu_ntt.test_compute_h_matches_oracle(cfft, ecx, m)


def test_compute_h_result_feeds_the_multiexp(cfft, ecx):
    gpu_ntt.test_compute_h_result_feeds_the_multiexp_on_the_device(cfft, ecx)


def test_fft_argument_errors(cfft):
    gpu_ntt.test_fft_argument_errors(cfft)


@pytest.mark.parametrize("m,n_a,n_b,mont", [(5, 7, 6, 0), (100, 90, 77, 1)])
def test_compute_groth16_proof(m, n_a, n_b, mont):
    gpu_g16.test_compute_groth16_proof_matches_oracle(m, n_a, n_b, mont)


def test_descending_thread_order_gives_the_same_bytes(cmsm, cfft):
    """the same driver with every launch's threads run from the last to the first (a child process: the order is read once)"""
    code = (
        "import sys, ctypes, numpy as np\n"
        "sys.path[:0] = [%r, %r, %r]\n"
        "import cmsm, cfft\n"
        "from gkrb200 import ec\n"
        "ec._lib = ec._bind(ctypes.CDLL(%r))\n"
        "c = ec.EcContext(0)\n"
        "pts = cmsm.gen_points(300); sc = cmsm.scalars_regular([(i * i * 0x9E3779B97F4A7C15 + 7) %% cmsm.Q for i in range(300)])\n"
        "c.SetBases(0, pts)\n"
        "assert np.array_equal(c.MultiExp(0, sc), cmsm.multiexp(pts, sc))\n"
        "c.NewDomain(100); a = cmsm.scalars_mont(range(1, 101)); b = cmsm.scalars_mont(range(7, 107)); cc = cfft.mul_elementwise(a, b)\n"
        "assert np.array_equal(c.ComputeH(a, b, cc), cfft.compute_h(a, b, cc, 128))\n"
        "print('reverse ok')\n"
    ) % (os.path.join(ROOT, "gkr-mimc_b200"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), SO)
    env = dict(os.environ, EC_HOSTBUILD_REVERSE="1")
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=600)
    assert out.returncode == 0 and "reverse ok" in out.stdout, out.stderr[-2000:]


def _sharded_worker(rank, world, port, n, out):
    """one rank of the sharded multi-exponentiation: the host build of the driver stands in for the rank's GPU"""
    import torch.distributed as dist
    for p_ in (os.path.join(ROOT, "gkr-mimc_b200"), os.path.join(ROOT, "oracle")):
        sys.path.insert(0, p_)
    import cmsm
    from gkrb200 import ec
    ec._lib = ec._bind(ctypes.CDLL(SO))
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pts = cmsm.gen_points(n)
        vals = [(i * i * 0x9E3779B97F4A7C15 + 99) % cmsm.Q for i in range(n)]
        lo, hi = n * rank // world, n * (rank + 1) // world  # contiguous slices; any partition of the index set works
        with ec.EcContext(device=0) as c:
            c.SetBases(0, pts[lo:hi])
            got = c.MultiExpSharded(0, cmsm.scalars_regular(vals[lo:hi]))
        want = cmsm.multiexp(pts, cmsm.scalars_regular(vals))
        out.put((rank, bool(np.array_equal(got, want))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_multiexp_over_gloo(world):
    """EcContext.MultiExpSharded: every rank multiplies its slice, the partial sums are all-gathered and added -- identical bytes on
    every rank, equal to the oracle's full sum (world_size 2 and 4, gloo, the host build of the driver on each rank)"""
    import socket
    import torch.multiprocessing as mp
    build_hostbuild()
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, world, port, 203, q)) for r in range(world)]
    for p_ in procs:
        p_.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p_ in procs:
        p_.join(timeout=60)
    assert res == [(r, True) for r in range(world)]
