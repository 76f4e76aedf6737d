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

// ------------------------------------------------------------------ small IIR/FIR, shift registers --
// y[n] = sum_{i<N} b[i] x[n-i] - sum_{1<=i<N} a[i] y[n-i], single accumulator, ascending i (SPEC I.1).
// Lines are shift registers ([0] newest) so every index is a compile-time constant -> registers.
template <int N>
static __device__ __forceinline__ float2 iir_step(const float (&b)[SDB_MAX_IIR], const float (&a)[SDB_MAX_IIR],
                                                  float (&xr)[SDB_MAX_IIR], float (&xi)[SDB_MAX_IIR],
                                                  float (&yr)[SDB_MAX_IIR], float (&yi)[SDB_MAX_IIR], float2 in)
{
#pragma unroll
  for (int i = N - 1; i > 0; --i) { xr[i] = xr[i - 1]; xi[i] = xi[i - 1]; }
  xr[0] = in.x; xi[0] = in.y;
  float ar = 0.0f, ai = 0.0f;
#pragma unroll
  for (int i = 0; i < N; ++i) { ar = ar + b[i] * xr[i]; ai = ai + b[i] * xi[i]; }
  if (N > 1) {
#pragma unroll
    for (int i = 1; i < N; ++i) { ar = ar - a[i] * yr[i - 1]; ai = ai - a[i] * yi[i - 1]; }
#pragma unroll
    for (int i = N - 1; i > 0; --i) { yr[i] = yr[i - 1]; yi[i] = yi[i - 1]; }
    yr[0] = ar; yi[0] = ai;
  }
  return make_float2(ar, ai);
}

static __device__ __forceinline__ float2 iir_any(int n, const float (&b)[SDB_MAX_IIR], const float (&a)[SDB_MAX_IIR],
                                                 float (&xr)[SDB_MAX_IIR], float (&xi)[SDB_MAX_IIR],
                                                 float (&yr)[SDB_MAX_IIR], float (&yi)[SDB_MAX_IIR], float2 in)
{
  switch (n) {
    case 1:  return iir_step<1>(b, a, xr, xi, yr, yi, in);
    case 2:  return iir_step<2>(b, a, xr, xi, yr, yi, in);
    case 3:  return iir_step<3>(b, a, xr, xi, yr, yi, in);
    case 4:  return iir_step<4>(b, a, xr, xi, yr, yi, in);
    default: return iir_step<5>(b, a, xr, xi, yr, yi, in);
  }
}

static __device__ __forceinline__ float sgnf(float v) { return v < 0.0f ? -1.0f : (v > 0.0f ? 1.0f : 0.0f); }

// ------------------------------------------------------------------ per-block state, all in registers --
// AGC (SPEC A).  dl / mh are strided arrays (stride 32 floats: one column per lane).
struct AgcK {
  float knee, slope_m1, fixed_gain, far_, faf, sar, saf;
  unsigned hang_max, dl_size, mh_size;
};
struct AgcS { float fast, slow, peak; unsigned hang_n, dl_ptr, mh_ptr; };

static __device__ __forceinline__ float2 agc_step(const AgcK &k, AgcS &s, float *dl, float *mh, float2 x)
{
  float2 xd = make_float2(dl[(2 * s.dl_ptr) * 32], dl[(2 * s.dl_ptr + 1) * 32]);
  dl[(2 * s.dl_ptr) * 32] = x.x; dl[(2 * s.dl_ptr + 1) * 32] = x.y;
  if (++s.dl_ptr >= k.dl_size) s.dl_ptr = 0;
  float m = 10.0f * d_log10f(x.x * x.x + x.y * x.y + 1e-16f);
  float m_old = mh[s.mh_ptr * 32];
  mh[s.mh_ptr * 32] = m;
  if (++s.mh_ptr >= k.mh_size) s.mh_ptr = 0;
  if (m > s.peak) {
    s.peak = m;
  } else if (s.peak == m_old) {
    float pk = -160.0f;
    for (unsigned i = 0; i < k.mh_size; ++i) { float v = mh[i * 32]; if (pk < v) pk = v; }
    s.peak = pk;
  }
  float d = s.peak - s.fast;
  if (d > 0.0f) s.fast = s.fast + k.far_ * d;
  else          s.fast = s.fast + k.faf * d;
  d = s.peak - s.slow;
  if (d > 0.0f) { s.slow = s.slow + k.sar * d; s.hang_n = 0; }
  else if (s.hang_n >= k.hang_max) s.slow = s.slow + k.saf * d;
  else ++s.hang_n;
  float lvl = s.fast > s.slow ? s.fast : s.slow;
  float g = lvl < k.knee ? k.fixed_gain : d_db_to_mag(lvl * k.slope_m1);
  g = g * 0.7f;
  xd.x = xd.x * g; xd.y = xd.y * g;
  return xd;
}

struct CostasK { int kind, af_n; float a, b; float af_b[SDB_MAX_IIR], af_a[SDB_MAX_IIR]; };
struct CostasS { float phi, omega, lock, yre, yim; float xr[SDB_MAX_IIR], xi[SDB_MAX_IIR], yr[SDB_MAX_IIR], yi[SDB_MAX_IIR]; };

static __device__ __forceinline__ float2 costas_step(const CostasK &k, CostasS &s, float2 x)
{
  float2 n = ncqo_read(s.phi, s.omega);
  float2 mixed = make_float2(x.x * n.x + x.y * n.y, x.y * n.x - x.x * n.y);
  float2 z = iir_any(k.af_n, k.af_b, k.af_a, s.xr, s.xi, s.yr, s.yi, mixed);
  float e = 0.0f, lr, li;
  if (k.kind == 1) {
    e = -(z.x * z.y);
  } else if (k.kind == 2) {
    lr = sgnf(z.x); li = sgnf(z.y);
    e = lr * z.y - li * z.x;
  } else if (k.kind == 3) {
    lr = sgnf(z.x); li = sgnf(z.y);
    if (fabsf(z.x) >= fabsf(z.y)) e = lr * z.y - li * z.x * 0.41421356237309504f;
    else                          e = lr * z.y * 0.41421356237309504f - li * z.x;
  }
  s.lock = s.lock + k.a * (1.0f - e - s.lock);
  s.yre = s.yre + 1.0f * (z.x - s.yre);
  s.yim = s.yim + 1.0f * (z.y - s.yim);
  s.omega = s.omega + k.b * e;
  s.phi = wrap_once(s.phi + k.a * e);
  return make_float2(s.yre, s.yim);
}

static __device__ __forceinline__ float2 pll_step(float alpha, float beta, float &phi, float &omega, float2 x)
{
  float2 ref = ncqo_read(phi, omega);
  float2 mix = make_float2(x.x * ref.x + x.y * ref.y, x.y * ref.x - x.x * ref.y);
  float err = d_atan2f(x.y, x.x) - phi;
  if (err > PI_F) err = err - TWOPI_F;
  else if (err < -PI_F) err = err + TWOPI_F;
  omega = omega + alpha * err;
  phi = wrap_once(phi + beta * err);
  return mix;
}

struct ClockS { float phi, bnor, x0r, x0i, x1r, x1i, x2r, x2i, pr, pi; int half; };

static __device__ __forceinline__ bool clock_step(float gain, float alpha, float beta, ClockS &s, float2 v, float2 &out)
{
  bool produced = false;
  s.phi = s.phi + s.bnor;
  if (s.phi >= 0.5f) {
    float al = s.bnor * (s.phi - 0.5f);
    float om = 1.0f - al;
    float pr = om * v.x + al * s.pr;
    float pi = om * v.y + al * s.pi;
    s.half = !s.half;
    s.phi = s.phi - 0.5f;
    if (!s.half) {
      s.x2r = s.x0r; s.x2i = s.x0i;
      s.x0r = pr; s.x0i = pi;
      float dr = s.x0r - s.x2r;
      float di = s.x0i - s.x2i;
      float e = gain * (s.x1r * dr + s.x1i * di);
      s.phi = s.phi + alpha * e;
      float bn = s.bnor + beta * e;
      if (bn > 1.0f) bn = 1.0f;
      if (bn < 0.0f) bn = 0.0f;
      s.bnor = bn;
      out = make_float2(pr, pi);
      produced = true;
    } else {
      s.x1r = pr; s.x1i = pi;
    }
  }
  s.pr = v.x; s.pi = v.y;
  return produced;
}

static __device__ __forceinline__ bool sampler_step(float period, float phase0, float &phase, float &pr, float &pi,
                                                    float2 v, float2 &out)
{
  bool sampled = false;
  if (period >= 1.0f) {
    phase = phase + 1.0f;
    if (phase >= period) phase = phase - period;
    float ph = phase - phase0;
    if (ph < 0.0f) ph = ph + period;
    float fl = floorf(ph);
    if (fl == 0.0f) {
      float al = ph - fl, om = 1.0f - al;
      out = make_float2(om * pr + al * v.x, om * pi + al * v.y);
      sampled = true;
    }
  }
  pr = v.x; pi = v.y;
  return sampled;
}

static __device__ __forceinline__ unsigned char decide(int mode, float dmin, float dh, int intervals, float2 x)
{
  float v = mode == 0 ? d_atan2f(x.y, x.x) : d_cabsf(x.x, x.y);
  float s = floorf((v - dmin) / dh * (float) intervals);
  int k = (int) s;
  if (!(s >= 0.0f)) k = 0;
  if (k > intervals - 1) k = intervals - 1;
  return (unsigned char) k;
}

// ------------------------------------------------------------------ the chain kernel ------------
// One CTA = 32 chains x 4 stage-warps.  Warp w runs stage w of every chain of the CTA (lane = chain):
//   warp 0  gain      : manual offset LO, AGC / fixed gain                     -> ring A
//   warp 1  carrier   : Costas | PLL + component select | FSK discriminator | audio demodulators -> ring B
//   warp 2  filter    : RRC matched filter (FIR) | audio low-pass            -> ring C
//   warp 3  clock     : Gardner | manual sampler | audio resampler, x0.75, decision, output
// Stage w works on chunk (it - w) of CHUNK samples while stage w+1 works on the previous one, so the
// serial latency per sample is that of the slowest stage (the carrier loop), not the sum; loop state
// lives in registers for the whole feed.  Chains are ordered channel-major so a warp normally holds 32
// streams of the SAME channel (uniform configuration, no divergence).
#define CHUNK 32
// matched-filter line: 8 margin + 2 n slots (duplicated line, n <= 32 taps) or n slots (single line, up to 136);
// the slot count of a launch comes from the channel plan (SdbInspDyn, sdb_internal.h)
__device__ unsigned long long g_stage_cycles[8];

cudaError_t sdb_stage_cycles(unsigned long long out[8], int reset)
{
  cudaError_t e = cudaMemcpyFromSymbol(out, g_stage_cycles, sizeof(unsigned long long) * 8);
  if (e == cudaSuccess && reset) {
    unsigned long long z[8] = { 0 };
    e = cudaMemcpyToSymbol(g_stage_cycles, z, sizeof(z));
  }
  return e;
}
static __device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc)
{
  const unsigned sa = (unsigned) __cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gsrc));
}
static __device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> static __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Fixed part of the CTA's shared memory.  The per-chain state lines that depend on the channel plan follow it,
// sized per launch (SdbInspDyn): matched-filter line [mf_slots][32] float2, AGC delay line + magnitude history
// [agc_rows][32] float, CMA weights and line 2 x [SDB_EQ_LEN][32] float2 when some chain equalises.  cfg2 (19 taps,
// 36 AGC floats) needs 84 KB, cfg3 (75 taps, 144 AGC floats) 112 KB: both leave two CTAs per SM.  A chain whose
// lines do not fit keeps them in the global pool (correct, slower: a global round trip inside the recurrence).
struct ChainSmem {
  float2 tile[2][32][33];
  float2 ring[3][2][CHUNK][32];
  float  lvl[CHUNK][32];
};

// SPEC E: constant-modulus equaliser on the symbol stream.  Sums and updates run in index order.
static __device__ __forceinline__ float2 cma_step(float2 (*w)[32], float2 (*x)[32], int lane, float mu,
                                                  int locked, float2 in)
{
#pragma unroll
  for (int i = SDB_EQ_LEN - 1; i > 0; --i) x[i][lane] = x[i - 1][lane];
  x[0][lane] = in;
  float yr = 0.0f, yi = 0.0f;
#pragma unroll
  for (int i = 0; i < SDB_EQ_LEN; ++i) {
    const float2 wi = w[i][lane], xi = x[i][lane];
    yr = yr + (wi.x * xi.x - wi.y * xi.y);
    yi = yi + (wi.x * xi.y + wi.y * xi.x);
  }
  if (!locked) {
    const float y2 = yr * yr + yi * yi;
    const float er = yr * (y2 - 1.0f), ei = yi * (y2 - 1.0f);
#pragma unroll
    for (int i = 0; i < SDB_EQ_LEN; ++i) {
      const float2 xi = x[i][lane];
      float2 wi = w[i][lane];
      const float gr = xi.x * er + xi.y * ei;
      const float gi = xi.x * ei - xi.y * er;
      wi.x = wi.x - mu * gr;
      wi.y = wi.y - mu * gi;
      w[i][lane] = wi;
    }
  }
  return make_float2(yr, yi);
}

// Matched filter for mf_n <= 32 taps, NT = mf_n rounded up to a multiple of 8.  The line is kept twice
// (slot p and p + mf_n, after a margin of 8 slots) so that x[n-t] is always at slot base - t: every
// load has a compile-time offset from one base address and none depends on another.  Taps t >= mf_n
// are zero: adding +-0 to a +0-initialised accumulator never changes its bits, so the result is exactly
// the SPEC I.1 sum over mf_n taps.
template <int NT>
static __device__ __forceinline__ float2 mf_fir(const float2 (*mfh)[32], int lane, unsigned ptr, int mf_n,
                                                const float (&tp)[32])
{
  const float2 *base = &mfh[8 + ptr + mf_n][lane];
  float accr = 0.0f, acci = 0.0f;
  float2 v[NT];
#pragma unroll
  for (int t = 0; t < NT; ++t) v[t] = base[-t * 32];
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    accr = accr + tp[t] * v[t].x;
    acci = acci + tp[t] * v[t].y;
  }
  return make_float2(accr, acci);
}

__global__ void __launch_bounds__(128) k_inspectors(const SdbChainCfg *__restrict__ cfgs, int n_channels,
                                                     int n_streams, SdbChainState *__restrict__ states,
                                                     float *__restrict__ pool, size_t pool_stride,
                                                     const float *__restrict__ taps_pool,
                                                     const SdbChannelDev *__restrict__ chans,
                                                     const float2 *__restrict__ chan_in,
                                                     size_t chan_stream_stride, uint32_t n_hops,
                                                     float2 *__restrict__ soft, unsigned char *__restrict__ hard,
                                                     uint32_t *__restrict__ sym_counts, size_t sym_cap, int fresh,
                                                     const SdbInspDyn dyn)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ChainSmem &sm = *reinterpret_cast<ChainSmem *>(smem_raw);
  const int mf_slots = dyn.mf_slots, agc_rows = dyn.agc_rows;
  float2 (*s_mfh)[32] = reinterpret_cast<float2 (*)[32]>(smem_raw + sizeof(ChainSmem));
  float (*s_agc)[32] = reinterpret_cast<float (*)[32]>(smem_raw + sizeof(ChainSmem) + (size_t) mf_slots * 32 * sizeof(float2));
  float2 (*s_eqw)[32] = reinterpret_cast<float2 (*)[32]>(reinterpret_cast<unsigned char *>(s_agc) + (size_t) agc_rows * 32 * sizeof(float));
  float2 (*s_eqx)[32] = s_eqw + SDB_EQ_LEN;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int chains = n_channels * n_streams;
  const int g = blockIdx.x * 32 + lane;                // channel-major chain index
  const bool valid = g < chains;
  const int k = valid ? g / n_streams : 0, s = valid ? g - k * n_streams : 0;
  const int chain = s * n_channels + k;                // index into states / outputs (stream-major)
  const SdbChainCfg *__restrict__ cp = cfgs + k;
  const int cls = valid ? cp->cls : -1;
  const uint32_t n = valid ? n_hops * (uint32_t) chans[k].halfsz : 0;
  // per-CTA pool, interleaved [slot][lane]
  float *bpool = pool + (size_t) blockIdx.x * 32 * pool_stride + lane;
  // number of chunks: CTA-wide maximum
  __shared__ uint32_t s_nmax;
  if (threadIdx.x == 0) s_nmax = 0;
  __syncthreads();
  if (warp == 0) atomicMax(&s_nmax, n);
  __syncthreads();
  const uint32_t nchunks = (s_nmax + CHUNK - 1) / CHUNK;
  SdbChainState *__restrict__ stp = states + chain;

  // ------------------------------------------------------------------ stage set-up (registers)
  // stage 0
  AgcK ak; AgcS as; float lo_phi = 0.0f, lo_omega = 0.0f, gain2 = 1.0f;
  int have_agc = 0, have_lo = 0;
  float *dl = nullptr, *mh = nullptr;
  // stage 1
  CostasK ck; CostasS cs; float p_phi = 0, p_omega = 0, pll_a = 0, pll_b = 0, prev_re = 0, prev_im = 0;
  float rot_re = 1, rot_im = 0, dc = 0, sq_level = 0, dc_alpha = 0, sq_alpha = 0, sq_thr = 0;
  int have_costas = 0, have_pll = 0, quad = 0, ask_ch = 0, ademod = 0, asquelch = 0;
  // stage 2
  int mf_n = 0, have_mf = 0, alpf_n = 0; unsigned mf_ptr = 0; bool mf_smem = false; const float *taps = nullptr; float *mfl = nullptr;
  float al_b[SDB_MAX_IIR], al_a[SDB_MAX_IIR], al_x[SDB_MAX_IIR], al_xi[SDB_MAX_IIR], al_y[SDB_MAX_IIR], al_yi[SDB_MAX_IIR];
  float tp[32];     // matched-filter taps in registers when mf_n <= 32
  // stage 3
  ClockS ks; float clk_gain = 0, clk_alpha = 0, clk_beta = 0, smp_period = 0, smp_phase0 = 0, s_phase = 0, s_pr = 0, s_pi = 0;
  int clock_type = 1, clock_running = 1, dec_mode = 0, dec_int = 1; float dec_min = 0, dec_h = 1;
  float avol = 1, rs_prev = 0; double rs_step = 0, rs_phase = 0;
  int eq_type = 0, eq_locked = 0; float eq_mu = 0;
  uint32_t nout = 0;

  if (valid) {
    if (warp == 0) {
      have_agc = cp->have_agc; have_lo = cls != SDB_INSP_AUDIO ? cp->have_lo : 0;
      gain2 = cls == SDB_INSP_AUDIO ? 1.0f : cp->gain2;
      lo_omega = cp->lo_omega; lo_phi = stp->lo_phi;
      ak.knee = cp->knee; ak.slope_m1 = cp->gain_slope - 1.0f; ak.fixed_gain = cp->fixed_gain;
      ak.far_ = cp->far_; ak.faf = cp->faf; ak.sar = cp->sar; ak.saf = cp->saf;
      ak.hang_max = cp->hang_max; ak.dl_size = cp->dl_size; ak.mh_size = cp->mh_size;
      as.fast = stp->fast_level; as.slow = stp->slow_level; as.peak = stp->peak;
      as.hang_n = stp->hang_n; as.dl_ptr = stp->dl_ptr; as.mh_ptr = stp->mh_ptr;
      const bool in_smem = 2 * ak.dl_size + ak.mh_size <= (unsigned) agc_rows;
      float *gdl = bpool + (size_t) cp->st_dl_off * 32, *gmh = bpool + (size_t) cp->st_mh_off * 32;
      if (in_smem && have_agc) {
        dl = &s_agc[0][lane]; mh = &s_agc[2 * ak.dl_size][lane];
        for (unsigned i = 0; i < 2 * ak.dl_size; ++i) dl[i * 32] = fresh ? 0.0f : gdl[i * 32];
        for (unsigned i = 0; i < ak.mh_size; ++i) mh[i * 32] = fresh ? -160.0f : gmh[i * 32];
      } else {
        dl = gdl; mh = gmh;
        if (fresh && have_agc) {
          for (unsigned i = 0; i < 2 * ak.dl_size; ++i) dl[i * 32] = 0.0f;
          for (unsigned i = 0; i < ak.mh_size; ++i) mh[i * 32] = -160.0f;
        }
      }
    } else if (warp == 1) {
      have_costas = cp->have_costas; have_pll = cp->have_pll;
      ck.kind = cp->costas_kind; ck.af_n = cp->af_n; ck.a = cp->c_a; ck.b = cp->c_b;
#pragma unroll
      for (int i = 0; i < SDB_MAX_IIR; ++i) {
        ck.af_b[i] = cp->af_b[i]; ck.af_a[i] = cp->af_a[i];
        cs.xr[i] = stp->afx_re[i]; cs.xi[i] = stp->afx_im[i]; cs.yr[i] = stp->afy_re[i]; cs.yi[i] = stp->afy_im[i];
      }
      cs.phi = stp->c_phi; cs.omega = stp->c_omega; cs.lock = stp->c_lock; cs.yre = stp->c_yre; cs.yim = stp->c_yim;
      p_phi = stp->p_phi; p_omega = stp->p_omega; pll_a = cp->pll_alpha; pll_b = cp->pll_beta;
      prev_re = stp->prev_re; prev_im = stp->prev_im;
      rot_re = cp->fsk_rot_re; rot_im = cp->fsk_rot_im; quad = cp->fsk_quad_demod; ask_ch = cp->ask_channel;
      ademod = cp->audio_demod; asquelch = cp->audio_squelch; dc = stp->dc; sq_level = stp->sq_level;
      dc_alpha = cp->dc_alpha; sq_alpha = cp->sq_alpha; sq_thr = cp->sq_thr;
      if (cls == SDB_INSP_AUDIO) { lo_phi = stp->lo_phi; lo_omega = cp->lo_omega; }
    } else if (warp == 2) {
      have_mf = cp->have_mf; mf_n = cp->mf_n; mf_ptr = stp->mf_ptr; taps = taps_pool + cp->mf_off;
      alpf_n = cls == SDB_INSP_AUDIO ? cp->alpf_n : 0;
#pragma unroll
      for (int i = 0; i < SDB_MAX_IIR; ++i) {
        al_b[i] = cp->alpf_b[i]; al_a[i] = cp->alpf_a[i]; al_x[i] = stp->al_x[i]; al_y[i] = stp->al_y[i];
        al_xi[i] = 0.0f; al_yi[i] = 0.0f;
      }
#pragma unroll
      for (int t = 0; t < 32; ++t) tp[t] = (have_mf && t < mf_n) ? __ldg(taps + t) : 0.0f;
      float *gmf = bpool + (size_t) cp->st_mf_off * 32;
      for (int i = 0; i < mf_slots; ++i) s_mfh[i][lane] = make_float2(0.f, 0.f);
      mf_smem = have_mf && (mf_n <= 32 ? 8 + 2 * mf_n <= mf_slots : mf_n <= mf_slots);
      if (mf_smem) {
        mfl = reinterpret_cast<float *>(&s_mfh[0][lane]);   // float2 ring, stride 32 float2 = 64 floats
        for (int i = 0; i < mf_n; ++i) {
          float2 v = fresh ? make_float2(0.f, 0.f) : make_float2(gmf[(2 * i) * 32], gmf[(2 * i + 1) * 32]);
          if (mf_n <= 32) { s_mfh[8 + i][lane] = v; s_mfh[8 + i + mf_n][lane] = v; }   // duplicated line
          else s_mfh[i][lane] = v;
        }
      } else {
        mfl = gmf;
        if (fresh && have_mf) for (int i = 0; i < 2 * mf_n; ++i) gmf[i * 32] = 0.0f;
      }
    } else {
      ks.phi = stp->k_phi; ks.bnor = stp->k_bnor; ks.x0r = stp->k_x0r; ks.x0i = stp->k_x0i; ks.x1r = stp->k_x1r;
      ks.x1i = stp->k_x1i; ks.x2r = stp->k_x2r; ks.x2i = stp->k_x2i; ks.pr = stp->k_pr; ks.pi = stp->k_pi;
      ks.half = stp->k_half;
      clk_gain = cp->clk_gain; clk_alpha = cp->clk_alpha; clk_beta = cp->clk_beta;
      smp_period = cp->smp_period; smp_phase0 = cp->smp_phase0; s_phase = stp->s_phase; s_pr = stp->s_pr; s_pi = stp->s_pi;
      clock_type = cp->clock_type; clock_running = cp->clock_running;
      dec_mode = cp->dec_mode; dec_int = cp->dec_intervals; dec_min = cp->dec_min; dec_h = cp->dec_h;
      avol = cp->audio_volume; rs_prev = stp->rs_prev; rs_step = cp->rs_step; rs_phase = stp->rs_phase;
      eq_type = cp->eq_type; eq_locked = cp->eq_locked; eq_mu = cp->eq_mu;
      if (eq_type == 1) {
        for (int i = 0; i < SDB_EQ_LEN; ++i) {
          s_eqw[i][lane] = make_float2(stp->eq_wr[i], stp->eq_wi[i]);
          s_eqx[i][lane] = make_float2(stp->eq_xr[i], stp->eq_xi[i]);
        }
      }
    }
  }
  float2 *__restrict__ so = soft + (size_t) chain * sym_cap;
  unsigned char *__restrict__ ho = hard + (size_t) chain * sym_cap;
  // coalesced, asynchronous tile load: row r of the tile = CHUNK consecutive samples of the CTA's chain r
  // (= lane r's own chain); lane L copies column L of every row.  Double-buffered one chunk ahead.
  const unsigned long long rowp =
      valid ? (unsigned long long) (chan_in + (size_t) s * chan_stream_stride + chans[k].out_off) : 0ull;
  auto issue_tile = [&](uint32_t c) {
    const uint32_t b0 = c * CHUNK;
#pragma unroll 8
    for (int r = 0; r < 32; ++r) {
      const unsigned long long pr = __shfl_sync(0xffffffffu, rowp, r);
      const uint32_t nr = __shfl_sync(0xffffffffu, n, r);
      if (b0 + lane < nr) cp_async8(&sm.tile[c & 1][r][lane], (const float2 *) pr + b0 + lane);
    }
    cp_async_commit();
  };
  if (warp == 0 && nchunks > 0) issue_tile(0);
  __syncthreads();

  long long busy = 0;
  for (uint32_t it = 0; it < nchunks + 3; ++it) {
    const long long t_begin = clock64();
    if (warp == 0) {
      if (it < nchunks) {
        const uint32_t base = it * CHUNK;
        // chunk `it` was prefetched (cp.async) one iteration ago; start the copy of chunk it+1 now
        if (it + 1 < nchunks) {
          issue_tile(it + 1);
          cp_async_wait<1>();
        } else {
          cp_async_wait<0>();
        }
        __syncwarp();
        float2 (*out)[32] = sm.ring[0][it & 1];
        float2 (*tl)[33] = sm.tile[it & 1];
        const int cnt = base >= n ? 0 : (n - base < CHUNK ? (int) (n - base) : CHUNK);
        if (have_agc && cls != SDB_INSP_RAW) {
          // Same arithmetic as agc_step() per sample, regrouped so that the long independent parts
          // (log10 of the magnitudes, 10^x of the gains) of different samples overlap; only the
          // peak / level tracker in the middle is a true recurrence.
          // pass 1: LO, delay-line swap, magnitude [dB]
          unsigned dlp = as.dl_ptr;
#pragma unroll 4
          for (int i = 0; i < CHUNK; ++i) {
            if (i < cnt) {
              float2 y = tl[lane][i];
              if (have_lo) {
                float2 ph = ncqo_read(lo_phi, lo_omega);
                y = make_float2(y.x * ph.x + y.y * ph.y, y.y * ph.x - y.x * ph.y);
              }
              const unsigned dp = dlp;
              dlp = dlp + 1 >= ak.dl_size ? 0 : dlp + 1;
              const float2 xd = make_float2(dl[(2 * dp) * 32], dl[(2 * dp + 1) * 32]);
              dl[(2 * dp) * 32] = y.x; dl[(2 * dp + 1) * 32] = y.y;
              out[i][lane] = xd;
              sm.lvl[i][lane] = 10.0f * d_log10f(y.x * y.x + y.y * y.y + 1e-16f);
            }
          }
          as.dl_ptr = dlp;
          // pass 2: magnitude history, running peak, fast / slow levels (serial)
          for (int i = 0; i < cnt; ++i) {
            const float m = sm.lvl[i][lane];
            const float m_old = mh[as.mh_ptr * 32];
            mh[as.mh_ptr * 32] = m;
            if (++as.mh_ptr >= ak.mh_size) as.mh_ptr = 0;
            if (m > as.peak) {
              as.peak = m;
            } else if (as.peak == m_old) {
              float pk = -160.0f;
              for (unsigned q = 0; q < ak.mh_size; ++q) { float v = mh[q * 32]; if (pk < v) pk = v; }
              as.peak = pk;
            }
            float d = as.peak - as.fast;
            if (d > 0.0f) as.fast = as.fast + ak.far_ * d;
            else          as.fast = as.fast + ak.faf * d;
            d = as.peak - as.slow;
            if (d > 0.0f) { as.slow = as.slow + ak.sar * d; as.hang_n = 0; }
            else if (as.hang_n >= ak.hang_max) as.slow = as.slow + ak.saf * d;
            else ++as.hang_n;
            sm.lvl[i][lane] = as.fast > as.slow ? as.fast : as.slow;
          }
          // pass 3: gain and scaling of the delayed sample
#pragma unroll 4
          for (int i = 0; i < CHUNK; ++i) {
            if (i < cnt) {
              const float lvl = sm.lvl[i][lane];
              float g = lvl < ak.knee ? ak.fixed_gain : d_db_to_mag(lvl * ak.slope_m1);
              g = g * 0.7f;
              float2 y = out[i][lane];
              y.x = y.x * g; y.y = y.y * g;
              if (cls != SDB_INSP_AUDIO) { y.x = 2.0f * y.x; y.y = 2.0f * y.y; }
              out[i][lane] = y;
            }
          }
        } else {
          for (int i = 0; i < cnt; ++i) {
            float2 y = tl[lane][i];
            if (have_lo) {
              float2 ph = ncqo_read(lo_phi, lo_omega);
              y = make_float2(y.x * ph.x + y.y * ph.y, y.y * ph.x - y.x * ph.y);
            }
            if (cls != SDB_INSP_RAW && cls != SDB_INSP_AUDIO) { y.x = gain2 * y.x; y.y = gain2 * y.y; }
            out[i][lane] = y;
          }
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      if (it >= 1 && it < nchunks + 1) {
        const uint32_t c = it - 1, base = c * CHUNK;
        float2 (*in)[32] = sm.ring[0][c & 1];
        float2 (*out)[32] = sm.ring[1][c & 1];
        for (int i = 0; i < CHUNK; ++i) {
          if (base + i < n) {
            float2 y = in[i][lane];
            if (cls == SDB_INSP_PSK) {
              if (have_costas) y = costas_step(ck, cs, y);
            } else if (cls == SDB_INSP_FSK) {
              float dr = y.x * prev_re + y.y * prev_im;
              float di = y.y * prev_re - y.x * prev_im;
              prev_re = y.x; prev_im = y.y;
              if (quad) { y.x = d_atan2f(di, dr) * 0.318309886183790671538f; y.y = 0.0f; }
              else { y.x = dr * rot_re - di * rot_im; y.y = dr * rot_im + di * rot_re; }
            } else if (cls == SDB_INSP_ASK) {
              if (have_pll) y = pll_step(pll_a, pll_b, p_phi, p_omega, y);
              if (ask_ch == 0)      { y.x = d_cabsf(y.x, y.y); y.y = 0.0f; }
              else if (ask_ch == 1) { y.y = 0.0f; }
              else                  { y.x = y.y; y.y = 0.0f; }
            } else if (cls == SDB_INSP_AUDIO) {
              float v = 0.0f;
              float p = y.x * y.x + y.y * y.y;
              sq_level = sq_level + sq_alpha * (p - sq_level);
              if (ademod == SDB_AUDIO_AM) {
                v = d_cabsf(y.x, y.y);
                dc = dc + dc_alpha * (v - dc);
                v = v - dc;
              } else if (ademod == SDB_AUDIO_FM) {
                float dr = y.x * prev_re + y.y * prev_im;
                float di = y.y * prev_re - y.x * prev_im;
                v = d_atan2f(di, dr) * 0.318309886183790671538f;
                prev_re = y.x; prev_im = y.y;
              } else if (ademod == SDB_AUDIO_USB || ademod == SDB_AUDIO_LSB) {
                float2 ph = ncqo_read(lo_phi, lo_omega);
                v = y.x * ph.x - y.y * ph.y;
              }
              if (asquelch && !(sq_level > sq_thr)) v = 0.0f;
              y = make_float2(v, 0.0f);
            }
            out[i][lane] = y;
          }
        }
      }
    } else if (warp == 2) {
      if (it >= 2 && it < nchunks + 2) {
        const uint32_t c = it - 2, base = c * CHUNK;
        float2 (*in)[32] = sm.ring[1][c & 1];
        float2 (*out)[32] = sm.ring[2][c & 1];
        for (int i = 0; i < CHUNK; ++i) {
          if (base + i < n) {
            float2 y = in[i][lane];
            if (have_mf) {
              float accr = 0.0f, acci = 0.0f;
              if (mf_smem && mf_n <= 32) {
                // taps in registers, duplicated line, branch-free unrolled sum (see mf_fir)
                s_mfh[8 + mf_ptr][lane] = y;
                s_mfh[8 + mf_ptr + mf_n][lane] = y;
                float2 r;
                if (mf_n <= 8)       r = mf_fir<8>(s_mfh, lane, mf_ptr, mf_n, tp);
                else if (mf_n <= 16) r = mf_fir<16>(s_mfh, lane, mf_ptr, mf_n, tp);
                else if (mf_n <= 24) r = mf_fir<24>(s_mfh, lane, mf_ptr, mf_n, tp);
                else                 r = mf_fir<32>(s_mfh, lane, mf_ptr, mf_n, tp);
                accr = r.x; acci = r.y;
              } else if (mf_smem) {
                // single line, two contiguous runs (newest ... slot 0, then slot mf_n-1 ... oldest): same tap order
                // as SPEC I.1, but every address is affine in t, so the loads pipeline (the former per-tap
                // "p = p ? p-1 : mf_n-1" chain made the 38-tap ASK filter of cfg3 the slowest stage of the kernel)
                s_mfh[mf_ptr][lane] = y;
                const int p0 = (int) mf_ptr;
                const float2 *l0 = &s_mfh[p0][lane];
                int t = 0;
#pragma unroll 4
                for (; t <= p0; ++t) {
                  const float b = __ldg(taps + t);
                  const float2 v = l0[-t * 32];
                  accr = accr + b * v.x;
                  acci = acci + b * v.y;
                }
#pragma unroll 4
                for (; t < mf_n; ++t) {
                  const float b = __ldg(taps + t);
                  const float2 v = s_mfh[mf_n + p0 - t][lane];
                  accr = accr + b * v.x;
                  acci = acci + b * v.y;
                }
              } else {
                mfl[(2 * mf_ptr) * 32] = y.x; mfl[(2 * mf_ptr + 1) * 32] = y.y;
                unsigned p = mf_ptr;
                for (int t = 0; t < mf_n; ++t) {
                  const float b = __ldg(taps + t);
                  accr = accr + b * mfl[(2 * p) * 32];
                  acci = acci + b * mfl[(2 * p + 1) * 32];
                  p = p == 0 ? mf_n - 1 : p - 1;
                }
              }
              mf_ptr = mf_ptr + 1 == (unsigned) mf_n ? 0 : mf_ptr + 1;
              y = make_float2(accr, acci);
            } else if (alpf_n > 0) {
              y = iir_any(alpf_n, al_b, al_a, al_x, al_xi, al_y, al_yi, y);
            }
            out[i][lane] = y;
          }
        }
      }
    } else {
      if (it >= 3) {
        const uint32_t c = it - 3, base = c * CHUNK;
        float2 (*in)[32] = sm.ring[2][c & 1];
        for (int i = 0; i < CHUNK; ++i) {
          if (base + i < n) {
            float2 y = in[i][lane], o;
            if (cls == SDB_INSP_RAW) {
              if (nout < sym_cap) { so[nout] = y; ho[nout] = 0; ++nout; }
            } else if (cls == SDB_INSP_AUDIO) {
              rs_phase += rs_step;
              if (rs_phase >= 1.0) {
                rs_phase -= 1.0;
                float al = (float) (rs_phase / rs_step);
                if (al > 1.0f) al = 1.0f;
                if (nout < sym_cap) {
                  so[nout] = make_float2(avol * ((1.0f - al) * y.x + al * rs_prev), 0.0f);
                  ho[nout] = 0;
                  ++nout;
                }
              }
              rs_prev = y.x;
            } else {
              bool produced;
              if (clock_type == 1) produced = clock_step(clk_gain, clk_alpha, clk_beta, ks, y, o);
              else                 produced = sampler_step(smp_period, smp_phase0, s_phase, s_pr, s_pi, y, o);
              if (produced && eq_type == 1) o = cma_step(s_eqw, s_eqx, lane, eq_mu, eq_locked, o);
              if (produced && clock_running && nout < sym_cap) {
                o.x = 0.75f * o.x; o.y = 0.75f * o.y;
                so[nout] = o;
                ho[nout] = decide(dec_mode, dec_min, dec_h, dec_int, o);
                ++nout;
              }
            }
          }
        }
      }
    }
    busy += clock64() - t_begin;
    __syncthreads();
  }
  if (lane == 0) {   // stage balance bookkeeping (profiles/): busy cycles per stage, samples processed
    atomicAdd(&g_stage_cycles[warp], (unsigned long long) busy);
    if (warp == 0) atomicAdd(&g_stage_cycles[4], (unsigned long long) nchunks * CHUNK);
  }

  // ------------------------------------------------------------------ write state back
  if (valid) {
    if (warp == 0) {
      stp->fast_level = as.fast; stp->slow_level = as.slow; stp->peak = as.peak;
      stp->hang_n = as.hang_n; stp->dl_ptr = as.dl_ptr; stp->mh_ptr = as.mh_ptr;
      if (cls != SDB_INSP_AUDIO) stp->lo_phi = lo_phi;
      if (have_agc && 2 * ak.dl_size + ak.mh_size <= (unsigned) agc_rows) {
        float *gdl = bpool + (size_t) cp->st_dl_off * 32, *gmh = bpool + (size_t) cp->st_mh_off * 32;
        for (unsigned i = 0; i < 2 * ak.dl_size; ++i) gdl[i * 32] = dl[i * 32];
        for (unsigned i = 0; i < ak.mh_size; ++i) gmh[i * 32] = mh[i * 32];
      }
    } else if (warp == 1) {
#pragma unroll
      for (int i = 0; i < SDB_MAX_IIR; ++i) {
        stp->afx_re[i] = cs.xr[i]; stp->afx_im[i] = cs.xi[i]; stp->afy_re[i] = cs.yr[i]; stp->afy_im[i] = cs.yi[i];
      }
      stp->c_phi = cs.phi; stp->c_omega = cs.omega; stp->c_lock = cs.lock; stp->c_yre = cs.yre; stp->c_yim = cs.yim;
      stp->p_phi = p_phi; stp->p_omega = p_omega; stp->prev_re = prev_re; stp->prev_im = prev_im;
      stp->dc = dc; stp->sq_level = sq_level;
      if (cls == SDB_INSP_AUDIO) stp->lo_phi = lo_phi;
    } else if (warp == 2) {
      stp->mf_ptr = mf_ptr;
#pragma unroll
      for (int i = 0; i < SDB_MAX_IIR; ++i) { stp->al_x[i] = al_x[i]; stp->al_y[i] = al_y[i]; }
      if (mf_smem) {
        float *gmf = bpool + (size_t) cp->st_mf_off * 32;
        for (int i = 0; i < mf_n; ++i) {
          float2 v = mf_n <= 32 ? s_mfh[8 + i][lane] : s_mfh[i][lane];
          gmf[(2 * i) * 32] = v.x; gmf[(2 * i + 1) * 32] = v.y;
        }
      }
    } else {
      stp->k_phi = ks.phi; stp->k_bnor = ks.bnor; stp->k_x0r = ks.x0r; stp->k_x0i = ks.x0i; stp->k_x1r = ks.x1r;
      stp->k_x1i = ks.x1i; stp->k_x2r = ks.x2r; stp->k_x2i = ks.x2i; stp->k_pr = ks.pr; stp->k_pi = ks.pi;
      stp->k_half = ks.half; stp->s_phase = s_phase; stp->s_pr = s_pr; stp->s_pi = s_pi;
      stp->rs_prev = rs_prev; stp->rs_phase = rs_phase;
      if (eq_type == 1) {
        for (int i = 0; i < SDB_EQ_LEN; ++i) {
          stp->eq_wr[i] = s_eqw[i][lane].x; stp->eq_wi[i] = s_eqw[i][lane].y;
          stp->eq_xr[i] = s_eqx[i][lane].x; stp->eq_xi[i] = s_eqx[i][lane].y;
        }
      }
      sym_counts[chain] = nout;
    }
  }
}

cudaError_t sdb_launch_inspectors_n(const SdbLaunchCtx &c, const SdbChainCfg *cfg_dev, int n_channels,
                                    int n_streams, SdbChainState *state, float *pool, size_t pool_stride,
                                    const float *taps_pool, const SdbChannelDev *chans_dev,
                                    const float2 *chan_in, size_t chan_stream_stride, uint32_t n_hops,
                                    float2 *soft, uint8_t *hard, uint32_t *sym_counts, size_t sym_cap, int fresh,
                                    const SdbInspDyn &dyn)
{
  const int chains = n_channels * n_streams;
  if (chains == 0) return cudaSuccess;
  static std::atomic<unsigned long long> attr_done{ 0 };   // one bit per device: function attributes are per context
  if (sdb_first_on_device(attr_done)) {
    cudaFuncSetAttribute(k_inspectors, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);   // + static < 227 KB
  }
  const size_t smem = sizeof(ChainSmem) + (size_t) dyn.mf_slots * 32 * sizeof(float2) +
                      (size_t) dyn.agc_rows * 32 * sizeof(float) +
                      (dyn.use_eq ? 2 * (size_t) SDB_EQ_LEN * 32 * sizeof(float2) : 0);
  k_inspectors<<<(chains + 31) / 32, 128, smem, c.stream>>>(
      cfg_dev, n_channels, n_streams, state, pool, pool_stride, taps_pool, chans_dev, chan_in,
      chan_stream_stride, n_hops, soft, hard, sym_counts, sym_cap, fresh, dyn);
  if (c.launch_counter) ++*c.launch_counter;
  return cudaGetLastError();
}

// ------------------------------------------------------------------ Tasks/ primitives -----------
// One thread per buffer of the batch; same recurrences, state in registers.
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

// mode 0 Costas, 1 PLL, 2 AGC.  pool: per-buffer AGC arrays, interleaved by 32 buffers ([slot][lane]).
__global__ void k_task_chain(const float2 *__restrict__ src, float2 *__restrict__ dst, size_t n, size_t batch,
                             const SdbChainCfg *__restrict__ cp, int mode, float *pool, size_t pool_stride)
{
  size_t b = blockIdx.x * (size_t) blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const float2 *x = src + b * n;
  float2 *y = dst + b * n;
  if (mode == 0) {
    CostasK ck; CostasS cs;
    ck.kind = cp->costas_kind; ck.af_n = cp->af_n; ck.a = cp->c_a; ck.b = cp->c_b;
#pragma unroll
    for (int i = 0; i < SDB_MAX_IIR; ++i) {
      ck.af_b[i] = cp->af_b[i]; ck.af_a[i] = cp->af_a[i];
      cs.xr[i] = cs.xi[i] = cs.yr[i] = cs.yi[i] = 0.0f;
    }
    cs.phi = cs.omega = cs.lock = cs.yre = cs.yim = 0.0f;
    for (size_t i = 0; i < n; ++i) y[i] = costas_step(ck, cs, x[i]);
  } else if (mode == 1) {
    float phi = 0.0f, omega = 0.0f;
    const float a = cp->pll_alpha, be = cp->pll_beta;
    for (size_t i = 0; i < n; ++i) y[i] = pll_step(a, be, phi, omega, x[i]);
  } else {
    AgcK ak; AgcS as;
    ak.knee = cp->knee; ak.slope_m1 = cp->gain_slope - 1.0f; ak.fixed_gain = cp->fixed_gain;
    ak.far_ = cp->far_; ak.faf = cp->faf; ak.sar = cp->sar; ak.saf = cp->saf;
    ak.hang_max = cp->hang_max; ak.dl_size = cp->dl_size; ak.mh_size = cp->mh_size;
    as.fast = as.slow = as.peak = -160.0f; as.hang_n = as.dl_ptr = as.mh_ptr = 0;
    float *bp = pool + (b / 32) * 32 * pool_stride + (b % 32);
    float *dl = bp + (size_t) cp->st_dl_off * 32, *mh = bp + (size_t) cp->st_mh_off * 32;
    for (unsigned i = 0; i < 2 * ak.dl_size; ++i) dl[i * 32] = 0.0f;
    for (unsigned i = 0; i < ak.mh_size; ++i) mh[i * 32] = -160.0f;
    for (size_t i = 0; i < n; ++i) y[i] = agc_step(ak, as, dl, mh, x[i]);
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
                                  const SdbChainCfg *cfg_dev, int mode, float *pool, size_t pool_stride)
{
  k_task_chain<<<(unsigned) ((batch + 31) / 32), 32, 0, s>>>(src, dst, n, batch, cfg_dev, mode, pool, pool_stride);
  return cudaGetLastError();
}
