"""Parity criteria (SPEC.md section T) shared by the GPU tests, smoke() and the benchmark.

* PSD bins (float32 FFT on both sides, different operation order): a bin P_k of a frame with mean
  power M may differ by  |dP| <= 1e-5 * P_k + 4e-6 * sqrt(P_k * M) + 1e-11 * M.
  The first term is the north-star's 1e-5 relative tolerance; the second is the float32 FFT noise
  floor, which is relative to the frame's total power, not to the bin (a 64K-point float32 FFT has
  ~1e-6 * rms(X) absolute error per bin whatever the algorithm); the third covers exact zeros.
* channel-rate samples out of the channeliser: |d| <= 1e-5 * rms(channel) + 4e-6 * rms(input) / sqrt(D).
* soft symbols through the full pipeline: counts equal, |d| <= SOFT_RTOL * rms(|ref|).
* soft symbols from identical channel-rate input (chain kernels alone): bit-identical.
* hard symbols: always bit-identical.
"""
import numpy as np

PSD_RTOL = 1e-5
PSD_FLOOR = 4e-6
SOFT_RTOL = 1e-5


def psd_err_ratio(got, ref):
    """max over bins of |dP| / tolerance (<= 1 passes)."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    M = ref.mean(axis=-1, keepdims=True)
    tol = PSD_RTOL * ref + PSD_FLOOR * np.sqrt(ref * M) + 1e-11 * M
    return float(np.max(np.abs(got - ref) / tol))


def assert_psd_close(got, ref):
    r = psd_err_ratio(got, ref)
    assert r <= 1.0, "PSD mismatch: worst bin at %.2f x tolerance" % r
    return r


def assert_channel_close(got, ref, x_rms, decim, rtol=1e-5):
    got = np.asarray(got, np.complex128)
    ref = np.asarray(ref, np.complex128)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if len(ref) == 0:
        return 0.0
    rms = np.sqrt(np.mean(np.abs(ref) ** 2))
    tol = rtol * rms + PSD_FLOOR * x_rms / np.sqrt(decim)
    worst = float(np.max(np.abs(got - ref)) / tol)
    assert worst <= 1.0, "channel output mismatch: %.2f x tolerance" % worst
    return worst


def assert_symbols_match(soft, hard, ref_soft, ref_hard, rtol=SOFT_RTOL, exact_soft=False):
    assert len(hard) == len(ref_hard), "symbol count differs: %d vs %d" % (len(hard), len(ref_hard))
    assert len(soft) == len(ref_soft)
    nbad = int(np.count_nonzero(np.asarray(hard) != np.asarray(ref_hard)))
    assert nbad == 0, "%d / %d hard symbols differ" % (nbad, len(hard))
    if len(soft) == 0:
        return 0.0
    if exact_soft:
        same = np.array_equal(np.asarray(soft).view(np.uint32), np.asarray(ref_soft).view(np.uint32))
        assert same, "soft symbols are not bit-identical (max |d| = %g)" % float(
            np.max(np.abs(np.asarray(soft) - np.asarray(ref_soft))))
        return 0.0
    rms = float(np.sqrt(np.mean(np.abs(ref_soft.astype(np.complex128)) ** 2)))
    err = float(np.max(np.abs(soft.astype(np.complex128) - ref_soft.astype(np.complex128))))
    assert err <= rtol * rms, "soft symbols differ: max |d| = %.3g vs tol %.3g" % (err, rtol * rms)
    return err / (rtol * rms)
