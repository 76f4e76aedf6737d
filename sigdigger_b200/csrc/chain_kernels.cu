// chain_kernels.cu -- per-channel inspector chains and Tasks/ primitives as serial recurrences, one
// GPU thread per (stream, channel) chain with the loop state carried in registers for the whole
// batch of channel samples (persistent across feeds through SdbChainState).
//
// COMPILED WITH -fmad=false: every expression below is a sequence of IEEE binary32 operations in
// source order, and the elementary functions are the fixed Cody-Waite + polynomial forms of
// SPEC.md section M, so soft symbols are bit-identical to the CPU oracle on identical input.
//
// Reference behaviour being replaced (all in suscan/sigutils, called from the reference at):
//   su_agc_feed            Tasks/AGCTask.cpp:70-73        su_costas_feed  Tasks/CostasRecoveryTask.cpp:58-61
//   su_pll_track           Tasks/PLLSyncTask.cpp:53-56    su_ncqo_read    Tasks/CarrierXlator.cpp:57-60
//   su_clock_detector_feed Tasks/WaveSampler.cpp:190-199  su_iir_filt_feed Tasks/WaveSampler.cpp:68-80
//   quadrature demod       Tasks/QuadDemodTask.cpp:44-60  Decider         Default/GenericInspector/InspectorUI.cpp:836-846
//   chain order            doc/SigDigger_User_Manual.pdf pp.50-52
#include "sdb_internal.h"
#include "../../include/sigdigger_b200.h"

#define PI_F   3.14159265358979323846f
#define TWOPI_F 6.28318530717958647692f

#include "sdb_math.h"

static __device__ __forceinline__ float wrap_once(float phi)
{
  if (phi >= TWOPI_F) phi = phi - TWOPI_F;
  else if (phi < 0.0f) phi = phi + TWOPI_F;
  return phi;
}

// NCQO read: y = exp(i phi); phi <- wrap(phi + omega)   (SPEC N)
static __device__ __forceinline__ float2 ncqo_read(float &phi, float omega)
{
  float s, c;
  d_sincosf(phi, &s, &c);
  phi = wrap_once(phi + omega);
  return make_float2(c, s);
}

// Small direct-form filter with circular lines of length n (SPEC I): used for the Costas arm filter
// and the audio low-pass.  Lines live in the state struct (registers / local memory).
static __device__ __forceinline__ float2 iir_feed(const float *b, const float *a, int n,
                                                  float *xre, float *xim, float *yre, float *yim,
                                                  unsigned &xp, unsigned &yp, float2 x)
{
  float accr = 0.0f, acci = 0.0f;
  xre[xp] = x.x; xim[xp] = x.y;
  unsigned p = xp;
  for (int i = 0; i < n; ++i) {
    accr = accr + b[i] * xre[p];
    acci = acci + b[i] * xim[p];
    p = p == 0 ? n - 1 : p - 1;
  }
  xp = xp + 1 == (unsigned) n ? 0 : xp + 1;
  if (n > 1) {
    p = yp;
    for (int i = 1; i < n; ++i) {
      accr = accr - a[i] * yre[p];
      acci = acci - a[i] * yim[p];
      p = p == 0 ? n - 1 : p - 1;
    }
    yp = yp + 1 == (unsigned) n ? 0 : yp + 1;
    yre[yp] = accr; yim[yp] = acci;
  }
  return make_float2(accr, acci);
}

// ------------------------------------------------------------------ chain blocks ----------------
struct AgcView {
  float *dl_re, *dl_im, *mh;     // strided pools in global memory
  int stride;
};

static __device__ __forceinline__ float2 agc_feed(const SdbChainCfg &c, SdbChainState &st, float *dl,
                                                  float *mh, float2 x)
{
  // delay line stores interleaved re, im
  float2 xd = make_float2(dl[2 * st.dl_ptr], dl[2 * st.dl_ptr + 1]);
  dl[2 * st.dl_ptr] = x.x; dl[2 * st.dl_ptr + 1] = x.y;
  if (++st.dl_ptr >= c.dl_size) st.dl_ptr = 0;

  float m = 10.0f * d_log10f(x.x * x.x + x.y * x.y + 1e-16f);
  float m_old = mh[st.mh_ptr];
  mh[st.mh_ptr] = m;
  if (++st.mh_ptr >= c.mh_size) st.mh_ptr = 0;

  if (m > st.peak) {
    st.peak = m;
  } else if (st.peak == m_old) {
    float pk = -160.0f;
    for (unsigned i = 0; i < c.mh_size; ++i) {
      float v = mh[i];
      if (pk < v) pk = v;
    }
    st.peak = pk;
  }
  float d = st.peak - st.fast_level;
  if (d > 0.0f) st.fast_level = st.fast_level + c.far_ * d;
  else          st.fast_level = st.fast_level + c.faf * d;
  d = st.peak - st.slow_level;
  if (d > 0.0f) {
    st.slow_level = st.slow_level + c.sar * d;
    st.hang_n = 0;
  } else if (st.hang_n >= c.hang_max) {
    st.slow_level = st.slow_level + c.saf * d;
  } else {
    ++st.hang_n;
  }
  float lvl = st.fast_level > st.slow_level ? st.fast_level : st.slow_level;
  float g = lvl < c.knee ? c.fixed_gain : d_db_to_mag(lvl * (c.gain_slope - 1.0f));
  g = g * 0.7f;
  xd.x = xd.x * g;
  xd.y = xd.y * g;
  return xd;
}

static __device__ __forceinline__ float sgnf(float v) { return v < 0.0f ? -1.0f : (v > 0.0f ? 1.0f : 0.0f); }

static __device__ __forceinline__ float2 costas_feed(const SdbChainCfg &c, SdbChainState &st, float2 x)
{
  float2 s = ncqo_read(st.c_phi, st.c_omega);
  float2 mixed = make_float2(x.x * s.x + x.y * s.y, x.y * s.x - x.x * s.y);
  float2 z = iir_feed(c.af_b, c.af_a, c.af_n, st.afx_re, st.afx_im, st.afy_re, st.afy_im, st.afxp,
                      st.afyp, mixed);
  float e = 0.0f, lr, li;
  switch (c.costas_kind) {
    case 1:
      e = -(z.x * z.y);
      break;
    case 2:
      lr = sgnf(z.x); li = sgnf(z.y);
      e = lr * z.y - li * z.x;
      break;
    case 3:
      lr = sgnf(z.x); li = sgnf(z.y);
      if (fabsf(z.x) >= fabsf(z.y)) e = lr * z.y - li * z.x * 0.41421356237309504f;
      else                          e = lr * z.y * 0.41421356237309504f - li * z.x;
      break;
    default:
      break;
  }
  st.c_lock = st.c_lock + c.c_a * (1.0f - e - st.c_lock);
  st.c_yre = st.c_yre + 1.0f * (z.x - st.c_yre);
  st.c_yim = st.c_yim + 1.0f * (z.y - st.c_yim);
  st.c_omega = st.c_omega + c.c_b * e;
  st.c_phi = wrap_once(st.c_phi + c.c_a * e);
  return make_float2(st.c_yre, st.c_yim);
}

static __device__ __forceinline__ float2 pll_track(const SdbChainCfg &c, SdbChainState &st, float2 x)
{
  float2 ref = ncqo_read(st.p_phi, st.p_omega);
  float2 mix = make_float2(x.x * ref.x + x.y * ref.y, x.y * ref.x - x.x * ref.y);
  float err = d_atan2f(x.y, x.x) - st.p_phi;
  if (err > PI_F) err = err - TWOPI_F;
  else if (err < -PI_F) err = err + TWOPI_F;
  st.p_omega = st.p_omega + c.pll_alpha * err;
  st.p_phi = wrap_once(st.p_phi + c.pll_beta * err);
  return mix;
}

static __device__ __forceinline__ bool clock_feed(const SdbChainCfg &c, SdbChainState &st, float2 v,
                                                  float2 &out)
{
  bool produced = false;
  st.k_phi = st.k_phi + st.k_bnor;
  if (st.k_phi >= 0.5f) {
    float al = st.k_bnor * (st.k_phi - 0.5f);
    float om = 1.0f - al;
    float pr = om * v.x + al * st.k_pr;
    float pi = om * v.y + al * st.k_pi;
    st.k_half = !st.k_half;
    st.k_phi = st.k_phi - 0.5f;
    if (!st.k_half) {
      st.k_x2r = st.k_x0r; st.k_x2i = st.k_x0i;
      st.k_x0r = pr; st.k_x0i = pi;
      float dr = st.k_x0r - st.k_x2r;
      float di = st.k_x0i - st.k_x2i;
      float e = c.clk_gain * (st.k_x1r * dr + st.k_x1i * di);
      st.k_phi = st.k_phi + c.clk_alpha * e;
      float bn = st.k_bnor + c.clk_beta * e;
      if (bn > 1.0f) bn = 1.0f;
      if (bn < 0.0f) bn = 0.0f;
      st.k_bnor = bn;
      out = make_float2(pr, pi);
      produced = true;
    } else {
      st.k_x1r = pr; st.k_x1i = pi;
    }
  }
  st.k_pr = v.x; st.k_pi = v.y;
  return produced;
}

static __device__ __forceinline__ bool sampler_feed(const SdbChainCfg &c, SdbChainState &st, float2 v,
                                                    float2 &out)
{
  bool sampled = false;
  if (c.smp_period >= 1.0f) {
    st.s_phase = st.s_phase + 1.0f;
    if (st.s_phase >= c.smp_period) st.s_phase = st.s_phase - c.smp_period;
    float ph = st.s_phase - c.smp_phase0;
    if (ph < 0.0f) ph = ph + c.smp_period;
    float fl = floorf(ph);
    if (fl == 0.0f) {
      float al = ph - fl, om = 1.0f - al;
      out = make_float2(om * st.s_pr + al * v.x, om * st.s_pi + al * v.y);
      sampled = true;
    }
  }
  st.s_pr = v.x; st.s_pi = v.y;
  return sampled;
}

static __device__ __forceinline__ unsigned char decide(const SdbChainCfg &c, float2 x)
{
  float v = c.dec_mode == 0 ? d_atan2f(x.y, x.x) : d_cabsf(x.x, x.y);
  float s = floorf((v - c.dec_min) / c.dec_h * (float) c.dec_intervals);
  int k = (int) s;
  if (!(s >= 0.0f)) k = 0;
  if (k > c.dec_intervals - 1) k = c.dec_intervals - 1;
  return (unsigned char) k;
}

// ------------------------------------------------------------------ the chain kernel ------------
// grid: ceil(S*K / block). thread -> chain (s, k).  Per-chain float pool layout:
//   [st_dl_off .. +2*dl_size) AGC delay line, [st_mh_off .. +mh_size) magnitude history,
//   [st_mf_off .. +2*mf_n) matched-filter delay line.
__global__ void __launch_bounds__(64) k_inspectors(const SdbChainCfg *__restrict__ cfgs, int n_channels,
                                                    int n_streams, SdbChainState *__restrict__ states,
                                                    float *__restrict__ pool, size_t pool_stride,
                                                    const float *__restrict__ taps_pool,
                                                    const SdbChannelDev *__restrict__ chans,
                                                    const float2 *__restrict__ chan_in,
                                                    size_t chan_stream_stride, uint32_t n_in,
                                                    float2 *__restrict__ soft, unsigned char *__restrict__ hard,
                                                    uint32_t *__restrict__ sym_counts, size_t sym_cap)
{
  const int chain = blockIdx.x * blockDim.x + threadIdx.x;
  if (chain >= n_channels * n_streams) return;
  const int s = chain / n_channels, k = chain - s * n_channels;
  const SdbChainCfg c = cfgs[k];
  SdbChainState st = states[chain];
  float *mypool = pool + (size_t) chain * pool_stride;
  float *dl = mypool + c.st_dl_off, *mh = mypool + c.st_mh_off, *mfl = mypool + c.st_mf_off;
  const float *taps = taps_pool + c.mf_off;
  const SdbChannelDev ch = chans[k];
  // number of channel samples this feed: n_in hops worth
  const uint32_t n = n_in * (uint32_t) ch.halfsz;
  const float2 *__restrict__ in = chan_in + (size_t) s * chan_stream_stride + ch.out_off;
  float2 *__restrict__ so = soft + (size_t) chain * sym_cap;
  unsigned char *__restrict__ ho = hard + (size_t) chain * sym_cap;
  uint32_t nout = 0;

  if (c.cls == SDB_INSP_RAW) {
    for (uint32_t i = 0; i < n && i < sym_cap; ++i) { so[i] = in[i]; ho[i] = 0; }
    sym_counts[chain] = n < sym_cap ? n : (uint32_t) sym_cap;
    return;
  }

  if (c.cls == SDB_INSP_AUDIO) {
    for (uint32_t i = 0; i < n; ++i) {
      float2 y = in[i];
      float v = 0.0f;
      if (c.have_agc) y = agc_feed(c, st, dl, mh, y);
      float p = y.x * y.x + y.y * y.y;
      st.sq_level = st.sq_level + c.sq_alpha * (p - st.sq_level);
      switch (c.audio_demod) {
        case SDB_AUDIO_AM:
          v = d_cabsf(y.x, y.y);
          st.dc = st.dc + c.dc_alpha * (v - st.dc);
          v = v - st.dc;
          break;
        case SDB_AUDIO_FM: {
          float dr = y.x * st.prev_re + y.y * st.prev_im;
          float di = y.y * st.prev_re - y.x * st.prev_im;
          v = d_atan2f(di, dr) * 0.318309886183790671538f;
          st.prev_re = y.x; st.prev_im = y.y;
          break;
        }
        case SDB_AUDIO_USB:
        case SDB_AUDIO_LSB: {
          float2 ph = ncqo_read(st.lo_phi, c.lo_omega);
          v = y.x * ph.x - y.y * ph.y;
          break;
        }
        default:
          break;
      }
      if (c.audio_squelch && !(st.sq_level > c.sq_thr)) v = 0.0f;
      float2 o = make_float2(v, 0.0f);
      if (c.alpf_n > 0) {
        // real-valued low-pass: reuse the complex helper with a zero imaginary line
        float xi[SDB_MAX_IIR] = {0, 0, 0, 0, 0}, yi[SDB_MAX_IIR] = {0, 0, 0, 0, 0};
        o = iir_feed(c.alpf_b, c.alpf_a, c.alpf_n, st.al_x, xi, st.al_y, yi, st.al_xp, st.al_yp, o);
      }
      st.rs_phase += c.rs_step;
      if (st.rs_phase >= 1.0) {
        st.rs_phase -= 1.0;
        float al = (float) (st.rs_phase / c.rs_step);
        if (al > 1.0f) al = 1.0f;
        if (nout < sym_cap) {
          so[nout] = make_float2(c.audio_volume * ((1.0f - al) * o.x + al * st.rs_prev), 0.0f);
          ho[nout] = 0;
          ++nout;
        }
      }
      st.rs_prev = o.x;
    }
    states[chain] = st;
    sym_counts[chain] = nout;
    return;
  }

  for (uint32_t i = 0; i < n; ++i) {
    float2 y = in[i], o;
    bool produced;

    if (c.have_lo) {
      float2 ph = ncqo_read(st.lo_phi, c.lo_omega);
      y = make_float2(y.x * ph.x + y.y * ph.y, y.y * ph.x - y.x * ph.y);
    }
    if (c.have_agc) {
      y = agc_feed(c, st, dl, mh, y);
      y.x = 2.0f * y.x; y.y = 2.0f * y.y;
    } else {
      y.x = c.gain2 * y.x; y.y = c.gain2 * y.y;
    }
    if (c.cls == SDB_INSP_PSK) {
      if (c.have_costas) y = costas_feed(c, st, y);
    } else if (c.cls == SDB_INSP_FSK) {
      float dr = y.x * st.prev_re + y.y * st.prev_im;
      float di = y.y * st.prev_re - y.x * st.prev_im;
      st.prev_re = y.x; st.prev_im = y.y;
      if (c.fsk_quad_demod) {
        y.x = d_atan2f(di, dr) * 0.318309886183790671538f;
        y.y = 0.0f;
      } else {
        y.x = dr * c.fsk_rot_re - di * c.fsk_rot_im;
        y.y = dr * c.fsk_rot_im + di * c.fsk_rot_re;
      }
    } else if (c.cls == SDB_INSP_ASK) {
      if (c.have_pll) y = pll_track(c, st, y);
      if (c.ask_channel == 0)      { y.x = d_cabsf(y.x, y.y); y.y = 0.0f; }
      else if (c.ask_channel == 1) { y.y = 0.0f; }
      else                         { y.x = y.y; y.y = 0.0f; }
    }

    if (c.have_mf) {
      // FIR, single accumulator, ascending tap index (SPEC I.1)
      mfl[2 * st.mf_ptr] = y.x; mfl[2 * st.mf_ptr + 1] = y.y;
      float accr = 0.0f, acci = 0.0f;
      unsigned p = st.mf_ptr;
      for (int t = 0; t < c.mf_n; ++t) {
        const float b = taps[t];
        accr = accr + b * mfl[2 * p];
        acci = acci + b * mfl[2 * p + 1];
        p = p == 0 ? c.mf_n - 1 : p - 1;
      }
      st.mf_ptr = st.mf_ptr + 1 == (unsigned) c.mf_n ? 0 : st.mf_ptr + 1;
      y = make_float2(accr, acci);
    }

    if (c.clock_type == 1) produced = clock_feed(c, st, y, o);
    else                   produced = sampler_feed(c, st, y, o);

    if (produced && c.clock_running && nout < sym_cap) {
      o.x = 0.75f * o.x; o.y = 0.75f * o.y;
      so[nout] = o;
      ho[nout] = decide(c, o);
      ++nout;
    }
  }
  states[chain] = st;
  sym_counts[chain] = nout;
}

cudaError_t sdb_launch_inspectors_n(const SdbLaunchCtx &c, const SdbChainCfg *cfg_dev, int n_channels,
                                    int n_streams, SdbChainState *state, float *pool, size_t pool_stride,
                                    const float *taps_pool, const SdbChannelDev *chans_dev,
                                    const float2 *chan_in, size_t chan_stream_stride, uint32_t n_hops,
                                    float2 *soft, uint8_t *hard, uint32_t *sym_counts, size_t sym_cap)
{
  const int chains = n_channels * n_streams;
  if (chains == 0) return cudaSuccess;
  const int block = 32;
  k_inspectors<<<(chains + block - 1) / block, block, 0, c.stream>>>(
      cfg_dev, n_channels, n_streams, state, pool, pool_stride, taps_pool, chans_dev, chan_in,
      chan_stream_stride, n_hops, soft, hard, sym_counts, sym_cap);
  if (c.launch_counter) ++*c.launch_counter;
  return cudaGetLastError();
}

// ------------------------------------------------------------------ Tasks/ primitives -----------
// One thread per buffer of the batch; same recurrences as above.
__global__ void k_task_xlate(const float2 *__restrict__ src, float2 *__restrict__ dst, size_t n, size_t batch,
                             float omega, float phi0)
{
  size_t b = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
  if (b >= batch) return;
  float phi = phi0;
  const float2 *x = src + b * n;
  float2 *y = dst + b * n;
  for (size_t i = 0; i < n; ++i) {
    float2 ph = ncqo_read(phi, omega);
    float2 v = x[i];
    y[i] = make_float2(v.x * ph.x - v.y * ph.y, v.x * ph.y + v.y * ph.x);
  }
}

// quadrature demod has no loop-carried state beyond x[p-1]: fully parallel over samples
__global__ void k_task_quad(const float2 *__restrict__ src, float2 *__restrict__ dst, size_t n, size_t batch)
{
  size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
  if (i >= n * batch) return;
  size_t p = i % n;
  if (p == 0) { dst[i] = make_float2(0.0f, 0.0f); return; }
  float2 x = src[i], pv = src[i - 1];
  float dr = x.x * pv.x + x.y * pv.y;
  float di = x.y * pv.x - x.x * pv.y;
  dst[i] = make_float2(0.0f, 0.318309886183790671538f * d_atan2f(di, dr));
}

__global__ void k_task_chain(const float2 *__restrict__ src, float2 *__restrict__ dst, size_t n, size_t batch,
                             SdbChainCfg c, int mode, float *pool, size_t pool_stride)
{
  size_t b = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
  if (b >= batch) return;
  SdbChainState st;
  memset(&st, 0, sizeof(st));
  st.fast_level = st.slow_level = st.peak = -160.0f;
  float *mypool = pool ? pool + b * pool_stride : nullptr;
  float *dl = mypool ? mypool + c.st_dl_off : nullptr, *mh = mypool ? mypool + c.st_mh_off : nullptr;
  const float2 *x = src + b * n;
  float2 *y = dst + b * n;
  for (size_t i = 0; i < n; ++i) {
    float2 v = x[i];
    if (mode == 0)      v = costas_feed(c, st, v);
    else if (mode == 1) v = pll_track(c, st, v);
    else                v = agc_feed(c, st, dl, mh, v);
    y[i] = v;
  }
}

cudaError_t sdb_launch_task_xlate(cudaStream_t s, const float2 *src, float2 *dst, size_t n, size_t batch,
                                  float omega, float phi0)
{
  k_task_xlate<<<(unsigned) ((batch + 31) / 32), 32, 0, s>>>(src, dst, n, batch, omega, phi0);
  return cudaGetLastError();
}
cudaError_t sdb_launch_task_quad(cudaStream_t s, const float2 *src, float2 *dst, size_t n, size_t batch)
{
  size_t tot = n * batch;
  k_task_quad<<<(unsigned) ((tot + 255) / 256), 256, 0, s>>>(src, dst, n, batch);
  return cudaGetLastError();
}
cudaError_t sdb_launch_task_chain(cudaStream_t s, const float2 *src, float2 *dst, size_t n, size_t batch,
                                  const SdbChainCfg &c, int mode, float *pool, size_t pool_stride)
{
  k_task_chain<<<(unsigned) ((batch + 31) / 32), 32, 0, s>>>(src, dst, n, batch, c, mode, pool, pool_stride);
  return cudaGetLastError();
}
