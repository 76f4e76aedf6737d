"""ctypes mirror of the suscan-style asynchronous analyzer (include/sigdigger_b200.h, sdb_analyzer_*).

Plays the role of `Suscan::Analyzer` + `Analyzer::AsyncThread` (Suscan/Analyzer.cpp:63-115, 601-638): construct,
queue requests, drain messages.  Message payloads are copied into Python objects and disposed at once."""
import ctypes as C

import numpy as np

from . import INSP, WINDOW, InspectorConfig, SdbError, last_error, load_library

MSG = {"SOURCE_INFO": 0, "SOURCE_INIT": 1, "CHANNEL": 2, "EOS": 3, "READ_ERROR": 4, "INTERNAL": 5, "SAMPLES": 6,
       "INSPECTOR": 7, "PSD": 8, "PARAMS": 9, "HALT": 0xffffffff, "TIMEOUT": 0xfffffffe}
MSG_NAME = {v: k for k, v in MSG.items()}
KIND = {0: "OPEN", 1: "SET_ID", 2: "GET_CONFIG", 3: "SET_CONFIG", 4: "ESTIMATOR", 5: "SPECTRUM", 6: "CLOSE",
        7: "INVALID_CHANNEL", 8: "WRONG_HANDLE", 9: "WRONG_OBJECT", 10: "WRONG_KIND"}


class Timeval(C.Structure):
    _fields_ = [("tv_sec", C.c_long), ("tv_usec", C.c_long)]


class SigutilsChannel(C.Structure):
    _fields_ = [("fc", C.c_double), ("ft", C.c_double), ("f_lo", C.c_double), ("f_hi", C.c_double), ("bw", C.c_float)]


class DetectorParams(C.Structure):
    _fields_ = [("window_size", C.c_uint64), ("window", C.c_int32), ("alpha", C.c_float), ("beta", C.c_float),
                ("gamma", C.c_float), ("snr", C.c_float)]


class AnalyzerParams(C.Structure):
    _fields_ = [("mode", C.c_int32), ("detector_params", DetectorParams), ("channel_update_int", C.c_float),
                ("psd_update_int", C.c_float), ("min_freq", C.c_double), ("max_freq", C.c_double)]


READ_FN = C.CFUNCTYPE(C.c_long, C.c_void_p, C.c_void_p, C.c_size_t)
BB_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64)
SETF_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_double)


class SourceConfig(C.Structure):
    _fields_ = [("samp_rate", C.c_double), ("freq", C.c_double), ("read_size", C.c_size_t), ("read", READ_FN),
                ("priv", C.c_void_p), ("data", C.c_void_p), ("length", C.c_size_t), ("loop", C.c_int32),
                ("device", C.c_int32), ("input_format", C.c_int32), ("set_frequency", SETF_FN)]


class PsdMsg(C.Structure):
    _fields_ = [("fc", C.c_int64), ("inspector_id", C.c_uint32), ("timestamp", Timeval), ("rt_time", Timeval),
                ("looped", C.c_int32), ("history_size", C.c_uint64), ("samp_rate", C.c_float),
                ("measured_samp_rate", C.c_float), ("psd_size", C.c_uint64), ("psd_data", C.POINTER(C.c_float))]


class SampleBatchMsg(C.Structure):
    _fields_ = [("inspector_id", C.c_uint32), ("samples", C.POINTER(C.c_float)), ("sample_count", C.c_uint64),
                ("symbols", C.POINTER(C.c_uint8))]


class InspectorMsg(C.Structure):
    _fields_ = [("kind", C.c_int32), ("inspector_id", C.c_uint32), ("req_id", C.c_uint32), ("handle", C.c_int32),
                ("class_name", C.c_char_p), ("channel", SigutilsChannel), ("config", InspectorConfig),
                ("fs", C.c_float), ("equiv_fs", C.c_float), ("bandwidth", C.c_float), ("lo", C.c_float),
                ("spectsrc_count", C.c_uint32), ("estimator_count", C.c_uint32), ("spectsrc_id", C.c_uint32),
                ("spectrum_data", C.POINTER(C.c_float)), ("spectrum_size", C.c_uint64), ("samp_rate", C.c_uint64),
                ("estimator_id", C.c_uint32), ("enabled", C.c_int32), ("value", C.c_float)]


class ChannelMsg(C.Structure):
    _fields_ = [("source", C.c_void_p), ("channel_count", C.c_uint32), ("channel_list", C.c_void_p)]


class StatusMsg(C.Structure):
    _fields_ = [("code", C.c_int32), ("err_msg", C.c_char_p)]


class Analyzer:
    def __init__(self, samp_rate, window_size=8192, window="hann", psd_update_int=0.04, data=None, read=None,
                 read_size=0, loop=False, freq=0.0, device=0, channel_update_int=0.0, alpha=0.25, beta=0.25,
                 gamma=0.5, snr=8.0, wide=False, min_freq=0.0, max_freq=0.0, set_frequency=None):
        self._L = load_library()
        p = AnalyzerParams()
        p.mode = 1 if wide else 0                      # SDB_ANALYZER_MODE_{CHANNEL, WIDE_SPECTRUM}
        p.min_freq, p.max_freq = min_freq, max_freq
        p.detector_params.window_size = window_size
        p.detector_params.window = WINDOW[window] if isinstance(window, str) else window
        p.detector_params.alpha, p.detector_params.beta = alpha, beta
        p.detector_params.gamma, p.detector_params.snr = gamma, snr
        p.psd_update_int = psd_update_int
        p.channel_update_int = channel_update_int
        self.params = p
        s = SourceConfig()
        s.samp_rate, s.freq, s.read_size, s.loop, s.device = samp_rate, freq, read_size, int(loop), device
        if data is not None:
            data = np.asarray(data)
            if data.dtype in (np.uint8, np.int8, np.int16):       # native capture formats: interleaved I,Q
                from . import FORMAT
                self._data = np.ascontiguousarray(data).reshape(-1)
                s.input_format = {np.dtype(np.uint8): FORMAT["u8"], np.dtype(np.int8): FORMAT["s8"],
                                  np.dtype(np.int16): FORMAT["s16"]}[self._data.dtype]
                s.data, s.length = self._data.ctypes.data, len(self._data) // 2
            else:
                self._data = np.ascontiguousarray(data, dtype=np.complex64)
                s.data, s.length = self._data.ctypes.data, len(self._data)
            s.read = READ_FN(0)
        else:
            self._cb = READ_FN(read)
            s.read = self._cb
        if set_frequency is not None:
            self._setf = SETF_FN(lambda priv, f: int(set_frequency(f) or 0))
            s.set_frequency = self._setf
        self._h = self._L.sdb_analyzer_new(C.byref(p), C.byref(s))
        if not self._h:
            raise SdbError(last_error() or "sdb_analyzer_new failed (no CUDA device or invalid parameters)")

    def close(self):
        if getattr(self, "_h", None):
            self._L.sdb_analyzer_destroy(self._h)
            self._h = None

    __del__ = close

    def halt(self):
        self._L.sdb_analyzer_req_halt(self._h)

    def open(self, cls, fc, bw, precise=False, req_id=0, parent=-1):
        """parent >= 0: sub-carrier inspector on that inspector's channel (fc relative to its centre)."""
        ch = SigutilsChannel(fc, 0.0, fc - bw / 2 - fc, fc + bw / 2 - fc, bw)
        ch.f_lo, ch.f_hi = -bw / 2, bw / 2          # as InspToolWidget.cpp:688-693 fills it
        if self._L.sdb_analyzer_open_ex_async(self._h, cls.encode(), C.byref(ch), int(precise), parent, req_id):
            raise SdbError("open_ex_async rejected")

    def set_inspector_id(self, handle, inspector_id, req_id=0):
        self._L.sdb_analyzer_set_inspector_id_async(self._h, handle, inspector_id, req_id)

    def set_inspector_config(self, handle, cfg, req_id=0):
        self._L.sdb_analyzer_set_inspector_config_async(self._h, handle, C.byref(cfg), req_id)

    def set_inspector_watermark(self, handle, watermark, req_id=0):
        self._L.sdb_analyzer_set_inspector_watermark_async(self._h, handle, int(watermark), req_id)

    def set_inspector_freq(self, handle, freq):
        self._L.sdb_analyzer_set_inspector_freq_overridable(self._h, handle, float(freq))

    def set_inspector_bandwidth(self, handle, bw):
        self._L.sdb_analyzer_set_inspector_bandwidth_overridable(self._h, handle, float(bw))

    # wide-spectrum (panoramic) mode controls, Panoramic/Scanner.cpp:396-503
    def set_hop_range(self, fmin, fmax):
        return self._L.sdb_analyzer_set_hop_range(self._h, float(fmin), float(fmax)) == 0

    def set_rel_bandwidth(self, rel_bw):
        self._L.sdb_analyzer_set_rel_bandwidth(self._h, float(rel_bw))

    def set_buffering_size(self, samples):
        self._L.sdb_analyzer_set_buffering_size(self._h, int(samples))

    def set_sweep_strategy(self, progressive=True):
        self._L.sdb_analyzer_set_sweep_strategy(self._h, 1 if progressive else 0)

    def set_spectrum_partitioning(self, continuous=False):
        self._L.sdb_analyzer_set_spectrum_partitioning(self._h, 1 if continuous else 0)

    def set_history_size(self, samples):
        return self._L.sdb_analyzer_set_history_size(self._h, int(samples)) == 0

    def replay(self, enabled=True):
        self._L.sdb_analyzer_replay(self._h, int(enabled))

    def seek(self, seconds):
        tv = Timeval(int(seconds), int(round((seconds - int(seconds)) * 1e6)))
        return self._L.sdb_analyzer_seek(self._h, C.byref(tv)) == 0

    def set_iq_reverse(self, enabled=True):
        self._L.sdb_analyzer_set_iq_reverse(self._h, int(enabled))

    def set_throttle(self, samp_rate, req_id=0):
        self._L.sdb_analyzer_set_throttle_async(self._h, int(samp_rate), req_id)

    def register_baseband_filter(self, fn):
        """fn(samples: complex64 ndarray view (writable), offset) -> truthy; runs on the worker thread."""
        def tramp(priv, a, ptr, length, offset):
            buf = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float)), shape=(2 * length,)).view(np.complex64)
            return 1 if fn(buf, offset) else 0
        cb = BB_FN(tramp)
        self._bb = getattr(self, "_bb", []) + [cb]          # keep the trampolines alive
        self._L.sdb_analyzer_register_baseband_filter(self._h, C.cast(cb, C.c_void_p), None)

    def set_spectrum_source(self, handle, spectsrc_id, req_id=0):
        self._L.sdb_analyzer_inspector_set_spectrum_async(self._h, handle, spectsrc_id, req_id)

    def estimator_cmd(self, handle, estimator_id, enabled=True, req_id=0):
        self._L.sdb_analyzer_inspector_estimator_cmd_async(self._h, handle, estimator_id, int(enabled), req_id)

    def close_inspector(self, handle, req_id=0):
        self._L.sdb_analyzer_close_async(self._h, handle, req_id)

    def read(self, timeout_ms=None):
        """-> (type name, payload dict); payload memory is disposed before returning."""
        t = C.c_uint32()
        if timeout_ms is None:
            ptr = self._L.sdb_analyzer_read(self._h, C.byref(t))
        else:
            ptr = self._L.sdb_analyzer_read_timeout(self._h, C.byref(t), timeout_ms)
        name = MSG_NAME.get(t.value, str(t.value))
        out = {}
        if ptr:
            if name == "PSD":
                m = C.cast(ptr, C.POINTER(PsdMsg)).contents
                out = dict(psd=np.ctypeslib.as_array(m.psd_data, shape=(m.psd_size,)).copy(), samp_rate=m.samp_rate,
                           measured_samp_rate=m.measured_samp_rate, fc=m.fc, looped=m.looped,
                           history_size=m.history_size,
                           timestamp=m.timestamp.tv_sec + 1e-6 * m.timestamp.tv_usec)
            elif name == "SAMPLES":
                m = C.cast(ptr, C.POINTER(SampleBatchMsg)).contents
                n = m.sample_count
                out = dict(inspector_id=m.inspector_id,
                           samples=np.ctypeslib.as_array(m.samples, shape=(2 * n,)).copy().view(np.complex64),
                           symbols=np.ctypeslib.as_array(m.symbols, shape=(n,)).copy())
            elif name == "INSPECTOR":
                m = C.cast(ptr, C.POINTER(InspectorMsg)).contents
                cfg = InspectorConfig()
                C.memmove(C.byref(cfg), C.byref(m.config), C.sizeof(cfg))
                out = dict(kind=KIND.get(m.kind, m.kind), req_id=m.req_id, handle=m.handle,
                           inspector_id=m.inspector_id, class_name=(m.class_name or b"").decode(), config=cfg,
                           fs=m.fs, equiv_fs=m.equiv_fs, bandwidth=m.bandwidth, lo=m.lo,
                           spectsrc_count=m.spectsrc_count, estimator_count=m.estimator_count,
                           spectsrc_id=m.spectsrc_id, estimator_id=m.estimator_id, enabled=m.enabled, value=m.value,
                           samp_rate=m.samp_rate, spectrum=None)
                if m.spectrum_data and m.spectrum_size:
                    out["spectrum"] = np.ctypeslib.as_array(m.spectrum_data, shape=(m.spectrum_size,)).copy()
            elif name == "CHANNEL":
                from . import DetectedChannel
                m = C.cast(ptr, C.POINTER(ChannelMsg)).contents
                arr = C.cast(m.channel_list, C.POINTER(DetectedChannel))
                out = dict(channels=[dict(fc=arr[i].fc, f_lo=arr[i].f_lo, f_hi=arr[i].f_hi, bw=arr[i].bw,
                                          snr=arr[i].snr, S0=arr[i].S0, N0=arr[i].N0) for i in range(m.channel_count)])
            elif name in ("EOS", "READ_ERROR", "HALT", "SOURCE_INIT"):
                m = C.cast(ptr, C.POINTER(StatusMsg)).contents
                out = dict(code=m.code, err_msg=(m.err_msg or b"").decode())
            self._L.sdb_analyzer_dispose_message(t.value, ptr)
        return name, out


__all__ = ["Analyzer", "MSG", "INSP"]
