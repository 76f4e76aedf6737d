"""Seeded synthetic complex-baseband generators (numpy) for tests and the benchmark.

Plays the role of suscan's `tonegen` / file sources that the reference selects from the GUI
(Default/SourceConfig/ToneGenSourcePage.cpp:81-87, FileSourcePage.cpp:80-104): raw complex float32
IQ at a nominal sample rate.  Nothing here is on the measured hot path.
"""
import numpy as np


def prbs(n, seed=0x7FFFFF, order=23):
    """PRBS-23 (x^23 + x^18 + 1) bit sequence, vectorised by blocks."""
    taps = {23: (23, 18), 15: (15, 14), 9: (9, 5)}[order]
    state = [(seed >> i) & 1 for i in range(order)]
    if not any(state):
        state[0] = 1
    out = np.empty(n + order, np.uint8)
    out[:order] = state
    a, b = taps
    # recurrence s[k] = s[k-a] ^ s[k-b]; generate in chunks of min(a,b) for vectorisation
    step = min(a, b)
    k = order
    while k < n + order:
        m = min(step, n + order - k)
        out[k:k + m] = out[k - a:k - a + m] ^ out[k - b:k - b + m]
        k += m
    return out[order:]


def rrc_taps(sps, beta, span):
    """Root-raised-cosine pulse sampled at sps samples/symbol over +-span symbols (float64), unit energy."""
    n = int(round(span * sps))
    t = np.arange(-n, n + 1, dtype=np.float64) / sps
    h = np.empty_like(t)
    for i, ti in enumerate(t):
        if abs(ti) < 1e-12:
            h[i] = 1.0 - beta + 4 * beta / np.pi
        elif abs(abs(4 * beta * ti) - 1.0) < 1e-9:
            h[i] = beta / np.sqrt(2) * ((1 + 2 / np.pi) * np.sin(np.pi / (4 * beta))
                                        + (1 - 2 / np.pi) * np.cos(np.pi / (4 * beta)))
        else:
            h[i] = (np.sin(np.pi * ti * (1 - beta)) + 4 * beta * ti * np.cos(np.pi * ti * (1 + beta))) \
                / (np.pi * ti * (1 - (4 * beta * ti) ** 2))
    return h / np.sqrt(np.sum(h * h))


def _shape(symbols, sps, beta, span, n):
    """Pulse-shape complex symbols at a (possibly non-integer) sps by polyphase evaluation -> n samples."""
    # fine grid: upsample by integer U >= sps*8 then pick nearest -- simpler: direct sum over neighbours
    h_os = 64
    taps = rrc_taps(h_os, beta, span)  # 64 samples / symbol
    half = (len(taps) - 1) // 2
    t = np.arange(n, dtype=np.float64) / sps                     # time in symbols
    k0 = np.floor(t).astype(np.int64)
    out = np.zeros(n, np.complex128)
    for d in range(-span, span + 1):
        k = k0 + d
        valid = (k >= 0) & (k < len(symbols))
        idx = np.rint((t - k) * h_os).astype(np.int64) + half
        ok = valid & (idx >= 0) & (idx < len(taps))
        out[ok] += symbols[k[ok]] * taps[idx[ok]]
    return out * np.sqrt(h_os)  # unit symbol amplitude -> unit peak-ish


def psk_signal(n, sps, order=4, beta=0.35, span=12, seed=1, bits=None):
    """RRC-shaped M-PSK at sps samples/symbol. Returns (x complex128, symbols_idx)."""
    bps = {2: 1, 4: 2, 8: 3}[order]
    nsym = int(n / sps) + 2 * span + 4
    if bits is None:
        bits = prbs(nsym * bps, seed=0x1000 + seed)
    b = bits[:nsym * bps].reshape(nsym, bps)
    idx = np.zeros(nsym, np.int64)
    for j in range(bps):
        idx = (idx << 1) | b[:, j]
    # constellation points at the centres of the decider's intervals (SPEC D)
    ph = -np.pi + (idx + 0.5) * (2 * np.pi / order)
    sym = np.exp(1j * ph)
    return _shape(sym, sps, beta, span, n), idx


def fsk_signal(n, sps, h=1.0, seed=2):
    """Binary CPFSK with modulation index h (rectangular pulses)."""
    nsym = int(n / sps) + 4
    bits = prbs(nsym, seed=0x1000 + seed)
    t = np.arange(n, dtype=np.float64) / sps
    k = np.floor(t).astype(np.int64)
    a = 2.0 * bits.astype(np.float64) - 1.0
    cum = np.concatenate([[0.0], np.cumsum(a)])
    phase = np.pi * h * (cum[k] + a[k] * (t - k))
    return np.exp(1j * phase), bits


def ask_signal(n, sps, levels=2, beta=0.35, span=12, seed=3, floor=0.15):
    nsym = int(n / sps) + 2 * span + 4
    bps = int(np.log2(levels))
    bits = prbs(nsym * bps, seed=0x1000 + seed)
    b = bits[:nsym * bps].reshape(nsym, bps)
    idx = np.zeros(nsym, np.int64)
    for j in range(bps):
        idx = (idx << 1) | b[:, j]
    amp = floor + (1.0 - floor) * idx / (levels - 1)
    return _shape(amp.astype(np.complex128), sps, beta, span, n), idx


def awgn(n, sigma, rng):
    return (rng.standard_normal(n) + 1j * rng.standard_normal(n)) * (sigma / np.sqrt(2.0))


def mix(x, fnor, phase=0.0):
    """Shift by fnor cycles/sample."""
    n = np.arange(len(x), dtype=np.float64)
    return x * np.exp(1j * (2 * np.pi * fnor * n + phase))


def multi_carrier(n, fs, carriers, noise_db=-60.0, seed=0):
    """Sum of carriers [(kind, f_hz, baud, amp_db, kwargs)] + AWGN -> complex64[n]."""
    rng = np.random.default_rng(seed)
    x = awgn(n, 10 ** (noise_db / 20), rng)
    meta = []
    for i, (kind, f, baud, amp_db, kw) in enumerate(carriers):
        sps = fs / baud
        if kind in ("bpsk", "qpsk", "8psk"):
            s, ref = psk_signal(n, sps, order={"bpsk": 2, "qpsk": 4, "8psk": 8}[kind], seed=seed * 131 + i, **kw)
        elif kind == "fsk":
            s, ref = fsk_signal(n, sps, seed=seed * 131 + i, **kw)
        elif kind == "ask":
            s, ref = ask_signal(n, sps, seed=seed * 131 + i, **kw)
        elif kind == "tone":
            s, ref = np.ones(n, np.complex128), None
        else:
            raise ValueError(kind)
        x = x + mix(s, f / fs, phase=0.1 * i) * 10 ** (amp_db / 20)
        meta.append(ref)
    return x.astype(np.complex64), meta


def tv_composite(line_len, hsync, short, lines, interlace, n_frames, picture, amp=0.8, noise=0.0, seed=1, n_eq=5):
    """Composite video with the sync tips HIGH (what TVProcessorTab::feed hands to the processor for a negatively
    modulated carrier): tip 1.0, blanking 0.7, picture 0.65 (black) ... 0.1 (white), times `amp`.
    picture[l] is sent on frame line l ([lines][C]).  Vertical intervals -- n_eq equalising pulses, n_eq broad pulses,
    n_eq equalising pulses, half a line apart -- start at the frame start and, when interlaced, half a frame later
    (mid-line for an odd line count).  line_len may be fractional.  Returns (float32 samples, (act0, act1)): the
    active part of a line spans [act0, act1) samples."""
    rng = np.random.default_rng(seed)
    total = int(line_len * lines * n_frames) + 10
    t = np.arange(total, dtype=np.float64)
    frame_len = line_len * lines
    tf = t - np.floor(t / frame_len) * frame_len
    ln = np.floor(tf / line_len).astype(np.int64)
    pos = tf - ln * line_len
    y = np.full(total, 0.7)
    act0, act1 = hsync * 2.2, line_len * 0.97
    cols = picture.shape[1]
    col = np.clip(((pos - act0) / (act1 - act0) * cols).astype(np.int64), 0, cols - 1)
    pic = (pos >= act0) & (pos < act1)
    y[pic] = 0.65 - 0.55 * picture[np.clip(ln, 0, lines - 1)[pic], col[pic]]
    y[pos < hsync] = 1.0
    half = line_len / 2
    for s0 in [0.0] + ([frame_len / 2] if interlace else []):
        m = (tf >= s0) & (tf < s0 + 3 * n_eq * half)
        u = tf[m] - s0
        k = np.floor(u / half).astype(np.int64)
        hp = u - k * half
        v = np.full(u.shape, 0.7)
        broad = (k >= n_eq) & (k < 2 * n_eq)
        v[(~broad) & (hp < short)] = 1.0
        v[broad & (hp < half - hsync)] = 1.0
        y[m] = v
    y *= amp
    if noise > 0:
        y += noise * rng.standard_normal(total)
    return y.astype(np.float32), (act0, act1)
