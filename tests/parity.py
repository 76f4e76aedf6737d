"""Parity criteria (SPEC.md section T) shared by the GPU tests, smoke() and the benchmark.

* PSD bins (float32 FFT on both sides, different operation order): a bin P_k of a frame with mean
  power M may differ by  |dP| <= 1e-5 * P_k + 4e-6 * sqrt(P_k * M) + 1e-11 * M.
  plus 2e-7 * sqrt(P_k * Pmax).  The first term is the north-star's 1e-5 relative tolerance; the second
  is the float32 FFT noise floor, which is relative to the frame's total power, not to the bin (a
  64K-point float32 FFT has ~1e-6 * rms(X) absolute error per bin whatever the algorithm); the
  Pmax term is the float32 dynamic range w.r.t. the strongest line (rounding of a line of amplitude
  |X|max lands, at ~eps * |X|max, on the bins that share late butterflies with it: measured on the
  oracle's own FFT against float64, see tests/test_oracle.py); the last covers exact zeros.
* channel-rate samples out of the channeliser: |d| <= 1e-5 * rms(channel) + 4e-6 * rms(input) / sqrt(D).
* soft symbols through the full pipeline: counts equal, |d| <= SOFT_RTOL * rms(|ref|).
* soft symbols from identical channel-rate input (chain kernels alone): bit-identical.
* hard symbols: always bit-identical.
"""
import numpy as np

PSD_RTOL = 1e-5
PSD_FLOOR = 4e-6
PSD_LINE = 2e-7
SOFT_RTOL = 1e-5


def psd_err_ratio(got, ref):
    """max over bins of |dP| / tolerance (<= 1 passes)."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    M = ref.mean(axis=-1, keepdims=True)
    Pmax = ref.max(axis=-1, keepdims=True)
    tol = PSD_RTOL * ref + PSD_FLOOR * np.sqrt(ref * M) + PSD_LINE * np.sqrt(ref * Pmax) + 1e-11 * M
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


def assert_symbols_match(soft, hard, ref_soft, ref_hard, rtol=SOFT_RTOL, exact_soft=False, skip=0,
                         chaos_ok=False):
    """skip: leading symbols excluded from the value comparison (counts are always compared).  The
    full-pipeline tests skip the symbols that come from the first half window of the channeliser:
    there the cross-fade weight sin^2 ramps up from 0, the channel signal is still below the float32
    noise floor of the 64K-point transforms, and the AGC amplifies that rounding noise to full scale
    on BOTH sides (SPEC.md section T)."""
    assert len(hard) == len(ref_hard), "symbol count differs: %d vs %d" % (len(hard), len(ref_hard))
    assert len(soft) == len(ref_soft)
    soft, hard, ref_soft, ref_hard = soft[skip:], hard[skip:], ref_soft[skip:], ref_hard[skip:]
    bad = np.flatnonzero(np.asarray(hard) != np.asarray(ref_hard))
    assert len(bad) == 0, "%d / %d hard symbols differ (first at %d)" % (len(bad), len(hard), bad[0] + skip)
    if len(soft) == 0:
        return 0.0
    if exact_soft:
        same = np.array_equal(np.asarray(soft).view(np.uint32), np.asarray(ref_soft).view(np.uint32))
        assert same, "soft symbols are not bit-identical (max |d| = %g)" % float(
            np.max(np.abs(np.asarray(soft) - np.asarray(ref_soft))))
        return 0.0
    rms = float(np.sqrt(np.mean(np.abs(ref_soft.astype(np.complex128)) ** 2)))
    d = np.abs(soft.astype(np.complex128) - ref_soft.astype(np.complex128))
    err = float(d.max())
    over = np.flatnonzero(d > rtol * rms)
    if chaos_ok and len(over):
        # The reference recurrences are discontinuous at a few thresholds (AGC hang counter reset, Gardner's
        # phi >= 0.5 test with the al = bnor*(phi-0.5) interpolation): a 1e-7 input perturbation can move one
        # sampling instant by a whole channel sample, after which the two runs re-converge only at the loop
        # time constant.  With inputs that differ in the last float bits (different FFT operation order)
        # the runs therefore agree to tolerance only up to the first such event.  Require: a long clean
        # prefix, a median error inside the tolerance, and identical hard decisions (checked above).
        assert over[0] >= min(250, len(d) // 4), "soft symbols diverge too early (index %d)" % (over[0] + skip)
        assert float(np.median(d)) <= rtol * rms, "median soft error %.3g above tol %.3g" % (np.median(d), rtol * rms)
        return err / (rtol * rms)
    assert err <= rtol * rms, ("soft symbols differ: max |d| = %.3g at %d (of %d; %d above tol %.3g, first at %d, "
                               "median |d| %.3g, p90 %.3g)") % (
        err, int(d.argmax()) + skip, len(d) + skip, len(over), rtol * rms, int(over[0]) + skip,
        float(np.median(d)), float(np.percentile(d, 90)))
    return err / (rtol * rms)
