"""CPU: host emulation of the table-driven device functions of bam_b200/csrc/bam_fastmath.cuh (tlog_pos, texp,
tsincos2pi_word) against long-double libm.  The emulation follows the device code operation by operation (fused
multiply-adds through extended precision); it pins the accuracy the parity tolerances of the fast kernels rest on."""
import numpy as np

LD = np.longdouble
N_LOG, N_EXP, N_SC = 256, 256, 256
# the coefficients of kFM / kFS as the device holds them
LN2 = 0.693147180559945309417232
C_EXP = 369.32993046757463        # 256 / ln2
C_HI, C_LO = -0.0027076061523985118, -2.1663774597568432e-11  # -ln2/256 split
MAGIC = 6755399441055744.0        # 1.5 * 2^52
# coefficients rounded to 20 mantissa bits (32-bit immediates of DFMA): FM_C5, FM_C24, FM_C120, FM_C720
C5, C24, C120, C720 = float.fromhex("0x1.9999ap-3"), float.fromhex("0x1.55555p-5"), float.fromhex("0x1.11111p-7"), float.fromhex("0x1.6c16cp-10")


def fma(a, b, c):
    return (LD(a) * LD(b) + LD(c)).astype(np.float64)


def tables():
    i = np.arange(N_LOG, dtype=LD)
    inv = (LD(1) / (LD(1) + (i + LD(0.5)) / LD(N_LOG))).astype(np.float64)
    mlog = (-np.log(inv.astype(LD))).astype(np.float64)
    ex = np.exp2(np.arange(N_EXP, dtype=LD) / LD(N_EXP)).astype(np.float64)
    pil = LD("3.141592653589793238462643383279502884")
    a = LD(2) * pil * (np.arange(N_SC, dtype=LD) + LD(0.5)) / LD(N_SC)
    return inv, mlog, ex, np.cos(a).astype(np.float64), np.sin(a).astype(np.float64)


INV, MLOG, EXPT, COST, SINT = tables()


def tlog_pos(x):
    bits = x.view(np.uint64)
    hi = (bits >> np.uint64(32)).astype(np.int64)
    idx = (hi >> 12) & 255
    m = ((bits & np.uint64(0x000FFFFFFFFFFFFF)) | np.uint64(0x3FF0000000000000)).view(np.float64)
    r = fma(m, INV[idx], -1.0)
    q = fma(C5, r, -0.25)
    q = fma(q, r, 1.0 / 3.0)
    q = fma(q, r, -0.5)
    pr = fma(q, r * r, r)
    s = fma(((hi >> 20) - 1023).astype(np.float64), LN2, MLOG[idx])
    return s + pr


def texp(x):
    t = fma(x, C_EXP, MAGIC)
    ki = (t.view(np.uint64) & np.uint64(0xFFFFFFFF)).astype(np.uint32).view(np.int32).astype(np.int64)
    kd = t - MAGIC
    r = fma(kd, C_HI, x)
    r = fma(kd, C_LO, r)
    T = EXPT[ki & 255]
    q = fma(C24, r, 1.0 / 6.0)
    q = fma(q, r, 0.5)
    q = fma(q, r, 1.0)
    res = fma(T, q * r, T)
    return np.ldexp(res, (ki >> 8).astype(np.int64))


def tsincos2pi_word(w):
    j = (w >> np.uint64(24)).astype(np.int64)
    low = (w & np.uint64(0xFFFFFF)).astype(np.float64)
    r = fma(low - 2.0**23, 1.4629180792671596e-09, 7.314590396335798e-10)
    r2 = r * r
    sr = fma(r, r2 * fma(r2, C120, -1.0 / 6.0), r)
    cm = r2 * fma(fma(r2, -C720, 1.0 / 24.0), r2, -0.5)
    C, S = COST[j], SINT[j]
    return S + fma(S, cm, C * sr), C + fma(C, cm, -(S * sr))


def test_tlog_accuracy():
    rng = np.random.default_rng(1)
    for lo, hi, tol in ((-3.0, 3.0, 6e-16), (-20.0, 20.0, 4e-15), (-300.0, 300.0, 6e-14)):
        x = np.exp(rng.uniform(lo, hi, 400_000))
        err = np.abs(tlog_pos(x).astype(LD) - np.log(x.astype(LD)))
        assert float(err.max()) <= tol, (lo, hi, float(err.max()))
    # the mantissa-interval edges of the 256-entry table and powers of two
    edge = np.concatenate([1.0 + np.arange(257) / 256.0, np.nextafter(1.0 + np.arange(1, 257) / 256.0, 0.0), 2.0 ** np.arange(-40, 40)])
    assert float(np.abs(tlog_pos(edge).astype(LD) - np.log(edge.astype(LD))).max()) <= 4e-15


def test_texp_accuracy():
    rng = np.random.default_rng(2)
    x = np.concatenate([rng.uniform(-700.0, 700.0, 400_000), rng.uniform(-2.0, 2.0, 200_000), [0.0, -0.0, 1e-300, 707.9, -707.9]])
    ref = np.exp(x.astype(LD))
    rel = np.abs((texp(x).astype(LD) - ref) / ref)
    assert float(rel.max()) <= 3e-16, float(rel.max())


def test_power_through_exp_of_log():
    """a (H - b)^c = texp(c tlog(H - b) + log a): the form the BaRatin fast kernels use, against long-double pow"""
    rng = np.random.default_rng(3)
    d = rng.uniform(1e-3, 8.0, 300_000)
    c = rng.uniform(1.0, 2.7, 300_000)
    la = np.log(rng.uniform(5.0, 200.0, 300_000))
    got = texp(fma(c, tlog_pos(d), la))
    ref = np.exp(la.astype(LD)) * d.astype(LD) ** c.astype(LD)
    assert float(np.abs((got.astype(LD) - ref) / ref).max()) <= 2e-14


def test_box_muller_angle():
    rng = np.random.default_rng(4)
    w = np.concatenate([rng.integers(0, 2**32, 1_000_000, dtype=np.uint64),
                        np.array([0, 2**32 - 1, 2**24 - 1, 2**24, 2**31, 2**31 - 1, 2**30, 3 * 2**30], dtype=np.uint64)])
    s, c = tsincos2pi_word(w)
    pil = LD("3.141592653589793238462643383279502884")
    th = LD(2) * pil * (w.astype(LD) + LD(0.5)) / LD(2.0**32)
    assert float(np.abs(s.astype(LD) - np.sin(th)).max()) <= 2.5e-16
    assert float(np.abs(c.astype(LD) - np.cos(th)).max()) <= 2.5e-16
    assert float(np.abs(s * s + c * c - 1.0).max()) <= 5e-16
