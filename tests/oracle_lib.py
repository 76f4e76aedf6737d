"""ctypes bindings of the CPU oracle (oracle/_build/libsdoracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "libsdoracle.so")


def build_oracle(force=False):
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
    stale = (not os.path.exists(LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
    return LIB_PATH


c_float_p = C.POINTER(C.c_float)


class Cpx(C.Structure):
    _fields_ = [("re", C.c_float), ("im", C.c_float)]


class Ncqo(C.Structure):
    _fields_ = [("phi", C.c_float), ("omega", C.c_float)]


class TvParams(C.Structure):
    """sdo_tv_params"""
    _fields_ = [("enable_sync", C.c_int32), ("reverse", C.c_int32), ("interlace", C.c_int32), ("enable_agc", C.c_int32),
                ("x_off", C.c_float), ("dominance", C.c_int32), ("frame_lines", C.c_uint32),
                ("frame_spacing", C.c_float), ("enable_comb", C.c_int32), ("comb_reverse", C.c_int32),
                ("hsync_len", C.c_float), ("vsync_len", C.c_float), ("line_len", C.c_float),
                ("vsync_odd_trigger", C.c_uint32), ("t_tol", C.c_float), ("l_tol", C.c_float), ("g_tol", C.c_float),
                ("hsync_huge_err", C.c_float), ("hsync_max_err", C.c_float), ("hsync_min_err", C.c_float),
                ("hsync_len_tau", C.c_float), ("line_len_tau", C.c_float), ("agc_tau", C.c_float),
                ("hsync_fast_track_tau", C.c_float), ("hsync_slow_track_tau", C.c_float)]


class FftPlan(C.Structure):
    _fields_ = [("n", C.c_uint), ("log2n", C.c_uint), ("tw_re", c_float_p), ("tw_im", c_float_p),
                ("rev", C.POINTER(C.c_uint))]


class Filt(C.Structure):
    _fields_ = [("nb", C.c_uint), ("na", C.c_uint), ("b", c_float_p), ("a", c_float_p),
                ("x", C.c_void_p), ("y", C.c_void_p), ("xp", C.c_uint), ("yp", C.c_uint),
                ("gain", C.c_float)]


class AgcParams(C.Structure):
    _fields_ = [("threshold", C.c_float), ("slope_factor", C.c_float), ("hang_max", C.c_uint),
                ("delay_line_size", C.c_uint), ("mag_history_size", C.c_uint),
                ("fast_rise_t", C.c_float), ("fast_fall_t", C.c_float), ("slow_rise_t", C.c_float),
                ("slow_fall_t", C.c_float)]


class Agc(C.Structure):
    _fields_ = [("enabled", C.c_int), ("knee", C.c_float), ("gain_slope", C.c_float),
                ("fixed_gain", C.c_float), ("hang_max", C.c_uint), ("hang_n", C.c_uint),
                ("fast_alpha_rise", C.c_float), ("fast_alpha_fall", C.c_float),
                ("slow_alpha_rise", C.c_float), ("slow_alpha_fall", C.c_float),
                ("fast_level", C.c_float), ("slow_level", C.c_float), ("peak", C.c_float),
                ("delay_line_size", C.c_uint), ("delay_line_ptr", C.c_uint),
                ("mag_history_size", C.c_uint), ("mag_history_ptr", C.c_uint),
                ("delay_line", C.c_void_p), ("mag_history", C.c_void_p)]


class Pll(C.Structure):
    _fields_ = [("alpha", C.c_float), ("beta", C.c_float), ("ncqo", Ncqo)]


class Costas(C.Structure):
    _fields_ = [("kind", C.c_int), ("a", C.c_float), ("b", C.c_float), ("y_alpha", C.c_float),
                ("gain", C.c_float), ("lock", C.c_float), ("y", Cpx), ("z", Cpx), ("ncqo", Ncqo),
                ("af", Filt)]


class Clock(C.Structure):
    _fields_ = [("alpha", C.c_float), ("beta", C.c_float), ("bnor", C.c_float), ("bmin", C.c_float),
                ("bmax", C.c_float), ("phi", C.c_float), ("gain", C.c_float), ("e", C.c_float),
                ("halfcycle", C.c_int), ("x", Cpx * 3), ("prev", Cpx)]


class Sampler(C.Structure):
    _fields_ = [("bnor", C.c_float), ("period", C.c_float), ("phase", C.c_float),
                ("phase0_rel", C.c_float), ("phase0", C.c_float), ("prev", Cpx)]


class Decider(C.Structure):
    _fields_ = [("mode", C.c_int), ("bps", C.c_uint), ("intervals", C.c_uint), ("min", C.c_float),
                ("max", C.c_float), ("h", C.c_float)]


class InspConfig(C.Structure):
    _fields_ = [("insp_class", C.c_int), ("fs", C.c_float), ("agc_enabled", C.c_int),
                ("agc_gain_db", C.c_float), ("costas_order", C.c_uint), ("bits_per_symbol", C.c_uint),
                ("loop_bw", C.c_float), ("offset", C.c_float), ("fsk_phase", C.c_float),
                ("fsk_quad_demod", C.c_int), ("ask_use_pll", C.c_int), ("ask_channel", C.c_uint),
                ("mf_type", C.c_uint), ("mf_rolloff", C.c_float), ("clock_type", C.c_uint),
                ("baud", C.c_float), ("clock_gain", C.c_float), ("clock_phase", C.c_float),
                ("clock_running", C.c_int), ("audio_cutoff", C.c_float), ("audio_volume", C.c_float),
                ("audio_squelch_level", C.c_float), ("agc_ts", C.c_float),
                ("audio_sample_rate", C.c_uint), ("audio_demod", C.c_uint), ("audio_squelch", C.c_int),
                ("eq_type", C.c_uint), ("eq_rate", C.c_float), ("eq_locked", C.c_int)]


class AnChannel(C.Structure):
    _fields_ = [("f0", C.c_float), ("bw", C.c_float), ("guard", C.c_float), ("precise", C.c_int),
                ("insp", InspConfig)]


class AnParams(C.Structure):
    _fields_ = [("psd_size", C.c_uint), ("psd_window", C.c_int), ("st_window_size", C.c_uint),
                ("n_channels", C.c_uint), ("channels", C.POINTER(AnChannel))]


class AnCounts(C.Structure):
    _fields_ = [("n_frames", C.c_size_t), ("n_chan", C.POINTER(C.c_size_t)),
                ("n_sym", C.POINTER(C.c_size_t))]


class StChannelParams(C.Structure):
    pass


ON_DATA = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.POINTER(Cpx), C.c_size_t)
StChannelParams._fields_ = [("f0", C.c_float), ("delta_f", C.c_float), ("bw", C.c_float),
                            ("guard", C.c_float), ("precise", C.c_int), ("privdata", C.c_void_p),
                            ("on_data", ON_DATA)]


class SpectrumView(C.Structure):
    _fields_ = [("freq_min", C.c_double), ("freq_max", C.c_double), ("freq_range", C.c_double),
                ("fft_bandwidth", C.c_double), ("fft_rel_bw", C.c_float), ("spectrum_size", C.c_uint),
                ("psd", c_float_p), ("psd_accum", c_float_p), ("psd_count", c_float_p)]


WINDOW = {"none": 0, "hamming": 1, "hann": 2, "flat_top": 3, "blackmann_harris": 4}
INSP = {"psk": 0, "fsk": 1, "ask": 2, "audio": 3, "raw": 4}

_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(LIB_PATH)
        L.sdo_atan2f.restype = C.c_float
        L.sdo_atan2f.argtypes = [C.c_float, C.c_float]
        L.sdo_log10f.restype = C.c_float
        L.sdo_log10f.argtypes = [C.c_float]
        L.sdo_exp10f.restype = C.c_float
        L.sdo_exp10f.argtypes = [C.c_float]
        L.sdo_sincosf.argtypes = [C.c_float, c_float_p, c_float_p]
        L.sdo_ncqo_read.restype = Cpx
        L.sdo_agc_feed.restype = Cpx
        L.sdo_agc_feed.argtypes = [C.c_void_p, Cpx]
        L.sdo_costas_feed.restype = Cpx
        L.sdo_costas_feed.argtypes = [C.c_void_p, Cpx]
        L.sdo_pll_track.restype = Cpx
        L.sdo_pll_track.argtypes = [C.c_void_p, Cpx]
        L.sdo_filt_feed.restype = Cpx
        L.sdo_filt_feed.argtypes = [C.c_void_p, Cpx]
        L.sdo_clock_feed.argtypes = [C.c_void_p, Cpx, C.c_void_p]
        L.sdo_sampler_feed.argtypes = [C.c_void_p, Cpx, C.c_void_p]
        L.sdo_ncqo_init.argtypes = [C.c_void_p, C.c_float]
        L.sdo_ncqo_set_phase.argtypes = [C.c_void_p, C.c_float]
        L.sdo_pll_init.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.sdo_costas_init.argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_uint, C.c_float]
        L.sdo_clock_init.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.sdo_sampler_init.argtypes = [C.c_void_p, C.c_float]
        L.sdo_sampler_set_phase.argtypes = [C.c_void_p, C.c_float]
        L.sdo_decider_init.argtypes = [C.c_void_p, C.c_int, C.c_uint, C.c_float, C.c_float]
        L.sdo_decider_decide.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.sdo_taps_rrc.argtypes = [c_float_p, C.c_uint, C.c_float, C.c_float]
        L.sdo_taps_brickwall_lp.argtypes = [c_float_p, C.c_uint, C.c_float]
        L.sdo_butter_lp.argtypes = [C.c_uint, C.c_float, c_float_p, c_float_p]
        L.sdo_mf_span.restype = C.c_uint
        L.sdo_mf_span.argtypes = [C.c_float]
        L.sdo_agc_params_from_tau.argtypes = [C.c_void_p, C.c_float, C.c_float]
        L.sdo_window_fill.argtypes = [c_float_p, C.c_uint, C.c_int]
        L.sdo_fft_plan_init.argtypes = [C.c_void_p, C.c_uint]
        L.sdo_fft_exec.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.sdo_spec_plan_init.argtypes = [C.c_void_p, C.c_uint, C.c_int]
        L.sdo_spec_plan_free.argtypes = [C.c_void_p]
        L.sdo_spec_forward.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sdo_psd_frame_spec.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sdo_psd_frame.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.sdo_psd_shift_db.argtypes = [C.c_void_p, C.c_uint]
        L.sdo_averager_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_float]
        L.sdo_quad_demod.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.sdo_carrier_xlate.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.sdo_specttuner_new.restype = C.c_void_p
        L.sdo_specttuner_new.argtypes = [C.c_uint]
        L.sdo_specttuner_destroy.argtypes = [C.c_void_p]
        L.sdo_specttuner_open_channel.restype = C.c_void_p
        L.sdo_specttuner_open_channel.argtypes = [C.c_void_p, C.c_void_p]
        L.sdo_specttuner_feed_bulk.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.sdo_st_channel_geometry.argtypes = [C.c_uint, C.c_float, C.c_float, C.c_float,
                                              C.POINTER(C.c_uint), C.POINTER(C.c_uint),
                                              C.POINTER(C.c_uint)]
        L.sdo_st_filter_response.argtypes = [C.c_uint, C.c_uint, c_float_p]
        L.sdo_insp_config_default.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.sdo_inspector_new.restype = C.c_void_p
        L.sdo_inspector_new.argtypes = [C.c_void_p]
        L.sdo_inspector_destroy.argtypes = [C.c_void_p]
        L.sdo_inspector_feed.restype = C.c_size_t
        L.sdo_inspector_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
        L.sdo_inspector_decider.argtypes = [C.c_void_p, C.c_void_p]
        L.sdo_analyzer_new.restype = C.c_void_p
        L.sdo_analyzer_new.argtypes = [C.c_void_p]
        L.sdo_analyzer_destroy.argtypes = [C.c_void_p]
        L.sdo_analyzer_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                        C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                        C.c_void_p]
        L.sdo_baseline_run.restype = C.c_double
        L.sdo_baseline_run.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int,
                                       C.POINTER(C.c_uint64)]
        L.sdo_sview_init.argtypes = [C.c_void_p]
        L.sdo_sview_free.argtypes = [C.c_void_p]
        L.sdo_sview_set_range.argtypes = [C.c_void_p, C.c_double, C.c_double]
        L.sdo_sview_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_double, C.c_int]
        L.sdo_sview_interpolate.argtypes = [C.c_void_p]
        _lib = L
    return _lib


def _c64(a):
    a = np.ascontiguousarray(a, dtype=np.complex64)
    return a


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


# ----------------------------------------------------------------------------------------------
# convenient numpy-level wrappers
# ----------------------------------------------------------------------------------------------
def sincos(x):
    L = lib()
    x = np.asarray(x, np.float32)
    s = np.empty_like(x)
    c = np.empty_like(x)
    sv, cv = C.c_float(), C.c_float()
    for i, v in enumerate(x.ravel()):
        L.sdo_sincosf(C.c_float(float(v)), C.byref(sv), C.byref(cv))
        s.ravel()[i] = sv.value
        c.ravel()[i] = cv.value
    return s, c


def vec1(fn, x):
    x = np.asarray(x, np.float32)
    out = np.empty_like(x)
    for i, v in enumerate(x.ravel()):
        out.ravel()[i] = fn(C.c_float(float(v)))
    return out


def fft(x, sign=-1):
    L = lib()
    x = _c64(x)
    p = FftPlan()
    assert L.sdo_fft_plan_init(C.byref(p), len(x)) == 0
    out = np.empty_like(x)
    L.sdo_fft_exec(C.byref(p), ptr(x), ptr(out), sign)
    L.sdo_fft_plan_free(C.byref(p))
    return out


def window(n, kind):
    w = np.empty(n, np.float32)
    lib().sdo_window_fill(w.ctypes.data_as(c_float_p), n, WINDOW[kind] if isinstance(kind, str) else kind)
    return w


class SpecPlan(C.Structure):
    _fields_ = [("N", C.c_uint), ("N1", C.c_uint), ("N2", C.c_uint), ("four", C.c_int), ("kind", C.c_int),
                ("tw_a", C.c_void_p), ("tw_b", C.c_void_p), ("tw_n", C.c_void_p), ("scr", C.c_void_p),
                ("buf", C.c_void_p), ("tmp", C.c_void_p)]


def spec_fft(x, four=False):
    """Forward SPEC transform (F.2 / F.3 / F.4 by size) of complex64 x."""
    L = lib()
    x = _c64(x)
    p = SpecPlan()
    assert L.sdo_spec_plan_init(C.byref(p), len(x), int(four)) == 0
    out = np.empty_like(x)
    L.sdo_spec_forward(C.byref(p), ptr(x), None, ptr(out))
    L.sdo_spec_plan_free(C.byref(p))
    return out


def psd_frames(x, n, kind, spec=True):
    """PSD of consecutive non-overlapping n-sample frames of x -> [frames, n] float32.
    spec=True uses the SPEC transform (bit-identical to the CUDA path); False the radix-2 cross-check."""
    L = lib()
    x = _c64(x)
    if spec:
        p = SpecPlan()
        assert L.sdo_spec_plan_init(C.byref(p), n, 0) == 0
        w = window(n, kind)
        none = (kind == "none" or kind == 0)
        nf = len(x) // n
        out = np.empty((nf, n), np.float32)
        scratch = np.empty(n, np.complex64)
        for f in range(nf):
            L.sdo_psd_frame_spec(C.byref(p), None if none else ptr(w), ptr(x[f * n:(f + 1) * n]), ptr(out[f]),
                                 ptr(scratch))
        L.sdo_spec_plan_free(C.byref(p))
        return out
    p = FftPlan()
    assert L.sdo_fft_plan_init(C.byref(p), n) == 0
    w = window(n, kind)
    nf = len(x) // n
    out = np.empty((nf, n), np.float32)
    scratch = np.empty(2 * n, np.complex64)
    for f in range(nf):
        L.sdo_psd_frame(C.byref(p), ptr(w), ptr(x[f * n:(f + 1) * n]), ptr(out[f]), ptr(scratch))
    L.sdo_fft_plan_free(C.byref(p))
    return out


def insp_config(cls, fs, **kw):
    c = InspConfig()
    lib().sdo_insp_config_default(C.byref(c), INSP[cls], C.c_float(fs))
    for k, v in kw.items():
        assert hasattr(c, k), k
        setattr(c, k, v)
    return c


def inspector_run(cfg, x, chunk=None):
    """Run one inspector over channel-rate samples x; returns (soft complex64, hard uint8)."""
    L = lib()
    x = _c64(x)
    h = L.sdo_inspector_new(C.byref(cfg))
    out = np.empty(len(x) + 16, np.complex64)
    k = 0
    step = chunk or len(x)
    for s in range(0, len(x), step):
        seg = x[s:s + step]
        k += L.sdo_inspector_feed(h, ptr(seg), len(seg), ptr(out[k:]), len(out) - k)
    L.sdo_inspector_destroy(h)
    soft = out[:k].copy()
    d = Decider()
    L.sdo_inspector_decider(C.byref(cfg), C.byref(d))
    hard = np.empty(k, np.uint8)
    if k:
        L.sdo_decider_decide(C.byref(d), ptr(soft), ptr(hard), k)
    return soft, hard


def specttuner_run(x, window_size, channels, chunk=None):
    """channels: list of dict(f0, bw, guard, precise). Returns list of complex64 arrays."""
    L = lib()
    x = _c64(x)
    st = L.sdo_specttuner_new(window_size)
    outs = [[] for _ in channels]
    cbs = []
    for i, ch in enumerate(channels):
        def mk(i):
            def cb(chp, priv, data, n):
                outs[i].append(np.ctypeslib.as_array(C.cast(data, c_float_p), shape=(2 * n,)).copy()
                               .view(np.complex64))
                return 1
            return ON_DATA(cb)
        cb = mk(i)
        cbs.append(cb)
        p = StChannelParams(f0=ch["f0"], delta_f=0.0, bw=ch["bw"], guard=ch.get("guard", 1.0),
                            precise=int(ch.get("precise", 0)), privdata=None, on_data=cb)
        assert L.sdo_specttuner_open_channel(st, C.byref(p)), "open_channel failed"
    step = chunk or len(x)
    for s in range(0, len(x), step):
        seg = x[s:s + step]
        assert L.sdo_specttuner_feed_bulk(st, ptr(seg), len(seg))
    L.sdo_specttuner_destroy(st)
    return [np.concatenate(o) if o else np.zeros(0, np.complex64) for o in outs]


def channel_geometry(window_size, f0, bw, guard):
    c, s, w = C.c_uint(), C.c_uint(), C.c_uint()
    lib().sdo_st_channel_geometry(window_size, f0, bw, guard, C.byref(c), C.byref(s), C.byref(w))
    return c.value, s.value, w.value


def make_an_params(psd_size, psd_window, channels, st_window_size=0):
    """channels: list of (f0, bw, guard, precise, InspConfig)."""
    arr = (AnChannel * max(1, len(channels)))()
    for i, (f0, bw, guard, precise, ic) in enumerate(channels):
        arr[i].f0, arr[i].bw, arr[i].guard, arr[i].precise, arr[i].insp = f0, bw, guard, int(precise), ic
    p = AnParams(psd_size=psd_size, psd_window=WINDOW[psd_window] if isinstance(psd_window, str) else psd_window,
                 st_window_size=st_window_size, n_channels=len(channels), channels=arr)
    p._keep = arr
    return p


def analyzer_run(params, x, chunk=None, want_chan=True):
    """Full oracle pass. Returns dict(psd [F,N], chan [list], soft [list], hard [list])."""
    L = lib()
    x = _c64(x)
    K = params.n_channels
    N = params.psd_size
    a = L.sdo_analyzer_new(C.byref(params))
    assert a
    step = chunk or len(x)
    psd_l, chan_l, soft_l, hard_l = [], [[] for _ in range(K)], [[] for _ in range(K)], [[] for _ in range(K)]
    for s in range(0, len(x), step):
        seg = x[s:s + step]
        n = len(seg)
        nf = n // N + 2
        psd = np.empty((nf, N), np.float32)
        chan = [np.empty(n + 16, np.complex64) for _ in range(K)]
        soft = [np.empty(n + 16, np.complex64) for _ in range(K)]
        hard = [np.empty(n + 16, np.uint8) for _ in range(K)]
        pp = lambda lst: (C.c_void_p * max(1, K))(*[ptr(v) for v in lst]) if K else None
        nch = (C.c_size_t * max(1, K))()
        nsy = (C.c_size_t * max(1, K))()
        cnt = AnCounts(0, C.cast(nch, C.POINTER(C.c_size_t)), C.cast(nsy, C.POINTER(C.c_size_t)))
        rc = L.sdo_analyzer_feed(a, ptr(seg), n, ptr(psd), nf, pp(chan) if want_chan else None, n + 16,
                                 pp(soft), pp(hard), n + 16, C.byref(cnt))
        assert rc == 0
        psd_l.append(psd[:cnt.n_frames].copy())
        for k in range(K):
            if want_chan:
                chan_l[k].append(chan[k][:nch[k]].copy())
            soft_l[k].append(soft[k][:nsy[k]].copy())
            hard_l[k].append(hard[k][:nsy[k]].copy())
    L.sdo_analyzer_destroy(a)
    cat = lambda l, dt: np.concatenate(l) if l else np.zeros(0, dt)
    return dict(psd=np.concatenate(psd_l) if psd_l else np.zeros((0, N), np.float32),
                chan=[cat(c, np.complex64) for c in chan_l],
                soft=[cat(c, np.complex64) for c in soft_l],
                hard=[cat(c, np.uint8) for c in hard_l])


# ----------------------------------------------------------------------------------------------
# U: inspector spectrum sources and baud estimators
# ----------------------------------------------------------------------------------------------
SPECTSRC = {"none": 0, "psd": 1, "cyclo": 2, "fmspect": 3, "timediff": 4, "abstimediff": 5, "exp_2": 6,
            "exp_4": 7, "exp_8": 8, "fac": 9}
ESTIMATOR = {"baud-fac": 0, "baud-nonlinear": 1}


def spectsrc_frame(kind, ns, chan):
    """chan = the channel samples of one feed.  Returns the emitted spectrum or None."""
    L = lib()
    L.sdo_spectsrc_frame.restype = C.c_uint
    L.sdo_spectsrc_frame.argtypes = [C.c_int, C.c_uint, C.c_void_p, C.c_size_t, C.c_void_p]
    chan = _c64(chan)
    out = np.zeros(ns, np.float32)
    k = SPECTSRC[kind] if isinstance(kind, str) else kind
    n = L.sdo_spectsrc_frame(k, ns, ptr(chan), len(chan), ptr(out))
    return out[:n].copy() if n else None


def estimate_baud(estimator, ns, fs_ch, chan):
    L = lib()
    L.sdo_estimate_baud.restype = C.c_int
    L.sdo_estimate_baud.argtypes = [C.c_int, C.c_uint, C.c_float, C.c_void_p, C.c_size_t, c_float_p]
    chan = _c64(chan)
    v = C.c_float()
    e = ESTIMATOR[estimator] if isinstance(estimator, str) else estimator
    ok = L.sdo_estimate_baud(e, ns, C.c_float(fs_ch), ptr(chan), len(chan), C.byref(v))
    return v.value if ok else None


# ----------------------------------------------------------------------------------------------
# K: channel detector
# ----------------------------------------------------------------------------------------------
class OChannel(C.Structure):
    _fields_ = [("bin_lo", C.c_uint), ("bin_hi", C.c_uint), ("s0", C.c_float), ("n0", C.c_float), ("snr", C.c_float)]


class ChDet(C.Structure):
    _fields_ = [("n", C.c_uint), ("min_bins", C.c_uint), ("last_total", C.c_uint), ("alpha", C.c_float),
                ("gamma", C.c_float), ("snr", C.c_float), ("n0", C.c_float), ("primed", C.c_int),
                ("n0_primed", C.c_int), ("avg", c_float_p), ("tmp", c_float_p)]


class ChannelDetector:
    def __init__(self, n, alpha, gamma, snr, min_bins):
        self.L = lib()
        self.L.sdo_chdet_feed.restype = C.c_uint
        self.d = ChDet()
        assert self.L.sdo_chdet_init(C.byref(self.d), n, C.c_float(alpha), C.c_float(gamma), C.c_float(snr), min_bins) == 0
        self.n = n

    def feed(self, psd, cap=256):
        """psd [frames, n] float32 -> (list of (bin_lo, bin_hi, s0, n0, snr), total)"""
        psd = np.ascontiguousarray(psd, np.float32).reshape(-1, self.n)
        out = (OChannel * cap)()
        k = self.L.sdo_chdet_feed(C.byref(self.d), ptr(psd), psd.shape[0], out, cap)
        return [(o.bin_lo, o.bin_hi, o.s0, o.n0, o.snr) for o in out[:k]], self.d.last_total

    def close(self):
        self.L.sdo_chdet_free(C.byref(self.d))


# ----------------------------------------------------------------------------------------------
# Y: offline TimeWindow tasks (oracle/tasks.c)
# ----------------------------------------------------------------------------------------------
SPACE = {"amplitude": 0, "phase": 1, "frequency": 2}


def delayed_conj(x, delay):
    x = _c64(x)
    y = np.empty_like(x)
    L = lib()
    L.sdo_delayed_conj.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    L.sdo_delayed_conj.restype = None
    L.sdo_delayed_conj(ptr(x), ptr(y), len(x), delay)
    return y


def histogram_feed(x, space):
    x = _c64(x)
    out = np.empty(len(x), np.float32)
    L = lib()
    L.sdo_histogram_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    L.sdo_histogram_feed.restype = C.c_size_t
    k = L.sdo_histogram_feed(ptr(x), ptr(out), len(x), SPACE[space])
    return out[:k].copy()


def sample_manual(x, space, symbol_count, symbol_sync=0):
    x = _c64(x)
    out = np.empty(int(symbol_count), np.complex64)
    L = lib()
    L.sdo_sample_manual.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_size_t, C.c_double, C.c_void_p]
    L.sdo_sample_manual.restype = C.c_size_t
    k = L.sdo_sample_manual(ptr(x), len(x), SPACE[space], symbol_sync, float(symbol_count), ptr(out))
    return out[:k].copy()


def sample_zero_crossing(x, space, bnor, amplitude=False, threshold=0j, zc_angle=1 + 0j, cap=None):
    x = _c64(x)
    cap = cap or len(x)
    sym = np.empty(cap, np.uint8)
    L = lib()
    L.sdo_sample_zero_crossing.argtypes = [C.c_void_p, C.c_size_t, C.c_int, C.c_int, Cpx, Cpx, C.c_float,
                                           C.c_void_p, C.c_size_t]
    L.sdo_sample_zero_crossing.restype = C.c_size_t
    t, z = complex(threshold), complex(zc_angle)
    k = L.sdo_sample_zero_crossing(ptr(x), len(x), SPACE[space], int(amplitude), Cpx(t.real, t.imag),
                                   Cpx(z.real, z.imag), min(float(bnor), 1.0), ptr(sym), cap)
    return sym[:min(k, cap)].copy(), k


def carrier_detect(x, avg_rel_bw=0.01, dc_notch_rel_bw=0.0):
    x = _c64(x)
    L = lib()
    L.sdo_carrier_detect.argtypes = [C.c_void_p, C.c_size_t, C.c_double, C.c_double]
    L.sdo_carrier_detect.restype = C.c_float
    return float(L.sdo_carrier_detect(ptr(x), len(x), avg_rel_bw, dc_notch_rel_bw))


def decide(soft, mode, bps, vmin, vmax):
    soft = _c64(soft)
    d = Decider()
    L = lib()
    L.sdo_decider_init(C.byref(d), {"argument": 0, "modulus": 1}[mode], bps, C.c_float(vmin), C.c_float(vmax))
    sym = np.empty(len(soft), np.uint8)
    L.sdo_decider_decide(C.byref(d), ptr(soft), ptr(sym), len(soft))
    return sym


class SnrEstimatorState(C.Structure):
    _fields_ = [("sigma", C.c_float), ("alpha", C.c_float), ("hx", C.c_float), ("delta", C.c_float),
                ("bps", C.c_uint), ("intervals", C.c_uint), ("length", C.c_uint),
                ("gaussian", c_float_p), ("hi", c_float_p), ("htilde", c_float_p)]


class SnrEstimator:
    """Misc/SNREstimator.cpp restated (oracle/tasks.c)"""

    def __init__(self, bps, length, alpha=1.0, sigma=None):
        self.L = lib()
        self.L.sdo_snr_get.restype = C.c_float
        self.e = SnrEstimatorState()
        self.L.sdo_snr_init(C.byref(self.e))
        self.L.sdo_snr_set_bps(C.byref(self.e), bps)      # the length is learnt at the first feed, as in the GUI
        self.e.alpha = alpha
        if sigma is not None:
            self.e.sigma = sigma

    def feed(self, history):
        h = np.ascontiguousarray(history, np.uint32)
        self.L.sdo_snr_feed(C.byref(self.e), ptr(h), len(h))

    @property
    def sigma(self):
        return float(self.e.sigma)

    @property
    def snr(self):
        return float(self.L.sdo_snr_get(C.byref(self.e)))

    def model(self):
        return np.ctypeslib.as_array(self.e.hi, shape=(self.e.length,)).copy()

    def close(self):
        self.L.sdo_snr_free(C.byref(self.e))
