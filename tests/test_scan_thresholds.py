"""The two arithmetic arguments behind the cheap decisions of the 4096-row path, checked with numpy (no GPU):

* median_scan_q32 decides on a FLOAT copy of q = |F|^2 whenever the float lies outside
  [qlo (1 - 1e-6) rounded down, qhi (1 + 1e-6) rounded up] and looks at the exact double otherwise
  (steganosaurus_b200/csrc/tfft_kernels.cu): a float outside that band must never contradict the double;
* the sign-map column pass reads a bin as `Im > 0` unless |Im| <= 1e-9 |Re|, where it evaluates the reference's full
  formula (read_bit_from_bin S:734-746, steganosaurus_b200/csrc/tfft_pencil.cu): off the real axis the two must agree.
"""
import numpy as np


def f32_down(x):
    f = np.float32(x)
    return f if float(f) <= x else np.nextafter(f, np.float32(-np.inf))


def f32_up(x):
    f = np.float32(x)
    return f if float(f) >= x else np.nextafter(f, np.float32(np.inf))


def test_float_copy_never_contradicts_the_double():
    rng = np.random.default_rng(1)
    for qlo, qhi in ((1.93e8, 2.004e8), (3.1, 3.3), (1e-20, 1.5e-20), (7e30, 7.3e30)):
        flo, fhi = f32_down(qlo * (1.0 - 1e-6)), f32_up(qhi * (1.0 + 1e-6))
        # doubles crowded around both edges (within 1e-5 relative) and spread over the bracket's neighbourhood
        q = np.concatenate([qlo * (1.0 + rng.uniform(-1e-5, 1e-5, 200000)), qhi * (1.0 + rng.uniform(-1e-5, 1e-5, 200000)),
                            rng.uniform(0.5 * qlo, 2.0 * qhi, 200000)])
        qf = q.astype(np.float32)
        assert np.all(q[qf < flo] < qlo)          # counted as below on the float alone
        assert np.all(q[qf > fhi] > qhi)          # skipped on the float alone
        looked = (qf >= flo) & (qf <= fhi)        # settled on the exact value
        assert np.all(looked[(q >= qlo) & (q <= qhi)])   # every bracket member is looked at
        assert looked.mean() < 0.5                        # ... and the band stays narrow


def read_bit_full(re, im, alpha):
    """read_bit_from_bin S:734-746."""
    th = np.arctan2(im, re)

    def dist(a, b):
        d = np.fmod(a - b + np.pi, 2 * np.pi)
        d = np.where(d < 0, d + 2 * np.pi, d)
        return np.abs(d - np.pi)
    return (dist(th, alpha) <= dist(th, -alpha)).astype(np.uint8)


def test_sign_decision_equals_the_formula_off_the_real_axis():
    rng = np.random.default_rng(2)
    n = 400000
    mag = 10.0 ** rng.uniform(-6, 9, n)
    # angles crowded around 0 and pi (just outside the 1e-9 band) and uniform ones
    th = np.concatenate([rng.uniform(-np.pi, np.pi, n // 2), rng.choice([0.0, np.pi], n // 2) + rng.uniform(-1e-6, 1e-6, n // 2)])
    re, im = mag * np.cos(th), mag * np.sin(th)
    off = np.abs(im) > 1e-9 * np.abs(re)
    assert off.mean() > 0.99
    for alpha in (0.5, 0.18, 1e-6, 3.14159, 1.5):
        assert np.array_equal(read_bit_full(re[off], im[off], alpha), (im[off] > 0).astype(np.uint8)), alpha
    # on the axis the formula reads 1, which is why the kernel falls back to it there
    assert read_bit_full(np.array([1.0, -1.0, -1.0, 0.0]), np.array([0.0, 0.0, -0.0, 0.0]), 0.5).tolist() == [1, 1, 1, 1]
