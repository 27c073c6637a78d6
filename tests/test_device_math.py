"""The fp64 kernels do not call libdevice for exp / atan2 / asin (mds_common.cuh: exp_, atan2_, asin_ for double): they
carry their own argument reductions and polynomials.  This CPU test reads the coefficient tables OUT OF THE HEADER and
evaluates the same schemes in numpy against numpy's functions, so a mistyped constant cannot hide behind the 1e-9 tolerances
of the GPU parity tests."""
import os
import re

import numpy as np

HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "multidronesim_b200", "csrc", "mds_common.cuh")


def _body(name):
    src = open(HDR).read()
    m = re.search(r"MDS_DEV double %s\(double[^)]*\) \{(.*?)\n\}" % name, src, re.S)
    assert m, name
    return m.group(1)


def _horner_coeffs(body):
    first = re.search(r"double p = (-?[0-9.e+-]+);", body)
    rest = re.findall(r"p = fma\(p, [rz], (-?[0-9.e+-]+)\);", body)
    return [float(first.group(1))] + [float(x) for x in rest]


def _horner(c, z):
    p = np.full_like(z, c[0])
    for k in c[1:]:
        p = p * z + k
    return p


def test_exp_restatement():
    body = _body("exp_")
    c = _horner_coeffs(body)
    assert len(c) == 14 and c[-1] == 1.0 and c[-2] == 1.0 and c[-3] == 0.5
    ln2_hi, ln2_lo = [-float(x) for x in re.findall(r"fma\(kd, (-[0-9.e+-]+), [xr]\)", body)]
    assert abs((ln2_hi + ln2_lo) - np.log(2.0)) < 1e-16
    x = -700.0 * np.random.default_rng(0).random(200000) ** 3
    k = np.rint(x * 1.4426950408889634)
    r = (x - k * ln2_hi) - k * ln2_lo
    got = np.ldexp(_horner(c, r), k.astype(np.int64))
    assert np.max(np.abs(got - np.exp(x)) / np.exp(x)) < 1e-15


def test_atan2_restatement():
    c = _horner_coeffs(_body("atan2_"))
    assert len(c) == 11
    rng = np.random.default_rng(1)
    y, x = rng.normal(size=200000) * 10.0 ** rng.integers(-4, 4, 200000), rng.normal(size=200000)
    ax, ay = np.abs(x), np.abs(y)
    mx, mn = np.maximum(ax, ay), np.minimum(ax, ay)
    mid = mn > 0.41421356237309503 * mx
    t = np.where(mid, mn - mx, mn) / np.where(mid, mn + mx, mx)
    z = t * t
    r = t + t * z * _horner(c, z)
    r = np.where(mid, r + 0.78539816339744831, r)
    r = np.where(ay > ax, 1.5707963267948966 - r, r)
    r = np.where(x < 0, 3.1415926535897932 - r, r)
    r = np.copysign(r, y)
    assert np.max(np.abs(r - np.arctan2(y, x))) < 2e-15


def test_asin_restatement():
    c = _horner_coeffs(_body("asin_"))
    assert len(c) == 13
    x = np.random.default_rng(2).uniform(-0.99999, 0.99999, 200000)
    a = np.abs(x)
    big = a > 0.5
    z = np.where(big, 0.5 * (1.0 - a), a * a)
    s = np.where(big, np.sqrt(z), a)
    r = s + s * z * _horner(c, z)
    r = np.copysign(np.where(big, 1.5707963267948966 - 2.0 * r, r), x)
    assert np.max(np.abs(r - np.arcsin(x))) < 2e-15
