"""CPU: the host/device-shared pieces of libofg compiled for the host (libofg_hostcheck.so, libofg_synth.so)
against the oracle, scipy and Python's own int()/float()."""
import ctypes
import os
import random

import numpy as np
import pytest

from oracle import waterfall_oracle as wo
from tests.helpers import Golden

PKG = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "orthofinder_b200")


@pytest.fixture(scope="module")
def hc():
    lib = ctypes.CDLL(os.path.join(PKG, "libofg_hostcheck.so"))
    lib.ofg_hostcheck_parse.restype = ctypes.c_long
    return lib


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _fit_both(hc, lx, ly):
    lx, ly = np.ascontiguousarray(lx), np.ascontiguousarray(ly)
    o1, st, o2 = np.zeros(2), np.zeros(3, np.int32), np.zeros(6)
    i1 = wo.lib().ofg_oracle_lmfit(_p(lx), _p(ly), ctypes.c_long(lx.size), _p(o1), _p(st))
    i2 = hc.ofg_hostcheck_fit(_p(lx), _p(ly), ctypes.c_long(lx.size), _p(o2))
    return (i1, o1, st), (i2, o2)


def test_fit_core_matches_oracle_and_scipy(hc):
    from scipy.optimize import curve_fit
    import warnings
    rng = np.random.default_rng(11)
    f = lambda x, a, b: a * np.log10(x) + b
    used_iterative = 0
    for t in range(300):
        m = int(rng.choice([2, 3, 5, 40, 208, 1024, 1025, 3000]))
        if t % 4 == 0:  # far-from-start problems reach lmpar's iterative branch
            L = np.exp(rng.uniform(0, 30, m))
            y = rng.choice([1e6, -1e5, 50.0]) * np.log10(L) + rng.choice([1e7, 3.0]) + rng.standard_normal(m)
        else:
            L = np.clip(np.round(rng.lognormal(5.8, 0.6, m)), 30, 5000) * np.clip(np.round(rng.lognormal(5.8, 0.6, m)), 30, 5000)
            y = 0.46 * np.log10(L) + 0.4 + rng.choice([0, 1e-3, 0.05, 0.3]) * rng.standard_normal(m)
        lx = wo.log10_ref(L)
        (i1, o1, st), (i2, o2) = _fit_both(hc, lx, y)
        assert i1 == i2 and o1[0] == o2[0] and o1[1] == o2[1] and st[0] == o2[4]
        used_iterative += int(o2[5])
        if wo.numpy_uses_svml():
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                try:
                    popt, _, infod, _, ier = curve_fit(f, L, y, full_output=True)
                except RuntimeError:
                    assert i1 not in (1, 2, 3, 4)
                    continue
            assert popt[0] == o1[0] and popt[1] == o1[1] and ier == i1 and infod["nfev"] == st[0]
    assert used_iterative > 10


def test_log10_emulation_matches_numpy():
    if not wo.numpy_uses_svml():
        pytest.skip("numpy does not dispatch float64 log10 to SVML on this host (no AVX512_SKX)")
    rng = np.random.default_rng(5)
    xs = [rng.uniform(1, 2, 2_000_000),
          np.clip(np.round(rng.lognormal(5.8, 0.6, 2_000_000)), 30, 5000) * np.clip(np.round(rng.lognormal(5.8, 0.6, 2_000_000)), 30, 5000),
          np.round(rng.uniform(20, 5000, 2_000_000), 1), np.arange(1, 2_000_001, dtype=np.float64),
          np.exp(rng.uniform(-700, 700, 2_000_000))]
    for x in xs:
        assert np.array_equal(np.log10(x), wo.log10_ref(x))


def test_score_grammar_against_python_float(hc):
    random.seed(3)
    cases = ["694", " 694", "55.8", "1.2e+03", "1e5", "1.", ".5", "  12.25  ", "+3.5", "-3", "0", "0.0", "1e22", "1e-22",
             "000012.5", "12.5000000000000000000000", "100000000000000000000", "0.000000000000000000001"]
    for _ in range(20000):
        k = random.random()
        if k < 0.5:
            cases.append("%.*f" % (random.randint(0, 6), random.uniform(0, 100000)))
        elif k < 0.7:
            cases.append("%d" % random.randint(0, 10 ** random.randint(1, 15)))
        else:
            cases.append("%.*e" % (random.randint(0, 6), random.uniform(0, 1e6)))
    for s in cases:
        o = ctypes.c_double(0)
        b = s.encode() + b"\t"
        rc = hc.ofg_hostcheck_score(b, len(b), ctypes.byref(o))
        assert rc == 0 and o.value == float(s), (s, rc, o.value)
    for s in ["abc", "", "1e", "--1", "1 2", "1.2.3", "e5"]:
        o = ctypes.c_double(0)
        b = s.encode() + b"\n"
        assert hc.ofg_hostcheck_score(b, len(b), ctypes.byref(o)) == 2  # ValueError in the reference
    for s in ["inf", "nan", "1_0", "12345678901234567890", "1e400"]:  # float() accepts, device grammar refuses loudly
        o = ctypes.c_double(0)
        b = s.encode() + b"\n"
        assert hc.ofg_hostcheck_score(b, len(b), ctypes.byref(o)) == 6


def test_record_parser_matches_oracle_on_reference_fixture(hc):
    g = Golden("ref_AddTwoSpecies")
    total = 0
    for (p, j), txt in g.texts.items():
        q, s, sc = wo.parse_hits_text(txt)
        f0 = np.zeros(len(q) + 8, np.uint32)
        f1 = np.zeros_like(f0)
        v = np.zeros(len(q) + 8)
        k = ctypes.c_int(0)
        n = hc.ofg_hostcheck_parse(txt, len(txt), _p(f0), _p(f1), _p(v), ctypes.byref(k))
        assert n == len(q)
        assert np.array_equal(f0[:n], q) and np.array_equal(f1[:n], s) and np.array_equal(v[:n], sc)
        total += n
    assert total == 43221


@pytest.mark.parametrize("line,kind", [
    (b"0_0\t0_0\t100.00\t466\t", 2),                       # the reference's ExampleBadBlast fixture: truncated line
    (b"00\t0_1\t1\t1\t1\t1\t1\t1\t1\t1\t1\t50", 1),         # no '_' in the query id
    (b"0_a\t0_1\t1\t1\t1\t1\t1\t1\t1\t1\t1\t50", 1),
    (b"0_1", 1),
    (b"0_1\t0_2\t1\t1\t1\t1\t1\t1\t1\t1\t1\tx", 2),
])
def test_malformed_lines(hc, line, kind):
    f0, f1, v, k = np.zeros(4, np.uint32), np.zeros(4, np.uint32), np.zeros(4), ctypes.c_int(0)
    txt = b"0_0\t0_1\t1\t1\t1\t1\t1\t1\t1\t1\t1\t50\n" + line + b"\n"
    off = txt.index(b"\n") + 1
    n = hc.ofg_hostcheck_parse(txt, len(txt), _p(f0), _p(f1), _p(v), ctypes.byref(k))
    assert n == -(off + 1) and k.value == kind
    with pytest.raises(wo.HitFileError) as e:
        wo.parse_hits_text(txt)
    assert e.value.kind == kind and e.value.offset == off
