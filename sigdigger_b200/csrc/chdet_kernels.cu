// chdet_kernels.cu -- channel detector on the main PSD (SPEC.md section K).
//
// Replaces what su_channel_detector contributes to the analyzer loop: detector_params.{alpha, gamma, snr}
// (Suscan/AnalyzerParams.cpp:27-66) in, lists of struct sigutils_channel {fc, f_lo, f_hi, bw, snr, S0, N0}
// (Suscan/Messages/ChannelMessage.cpp:25-70, include/Suscan/Channel.h:26-32) out, once per feed
// (channel_update_int).  The upstream algorithm is not in the reference; SPEC K is this project's definition and
// is order-independent on purpose (exact order statistic instead of a mean, maxima instead of sums), so this
// parallel implementation is bit-identical to oracle/chdetect.c.
//
//   k_chdet_avg : one thread per (bin, stream): exponential average over the feed's frames, in frame order.
//   k_chdet_find: one 1024-thread CTA per stream: exact (N/4)-th smallest averaged bin by a 4-pass radix select on
//                 the float bit patterns (shared-memory histogram), noise-floor update, threshold, run detection
//                 with block scans (ordered start / end events in shared memory), width filter + ordered
//                 compaction, per-channel maximum by one warp per channel.
#include "sdb_internal.h"
#include "../../include/sigdigger_b200.h"
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>

#define CHDET_MAXRAW 8192      // raw runs examined per update (SPEC K.4)
#define CHDET_CAP    256       // channels reported per stream

struct SdbDetectedDev { unsigned bin_lo, bin_hi; float s0, n0, snr; };

__global__ void k_chdet_avg(float *__restrict__ avg, const float *__restrict__ psd, int N, int frames,
                            size_t stream_stride, float alpha, int primed)
{
  const int k = blockIdx.x * blockDim.x + threadIdx.x, s = blockIdx.y;
  if (k >= N) return;
  const float *__restrict__ p = psd + (size_t) s * stream_stride + k;
  float a = avg[(size_t) s * N + k];
  int f = 0;
  if (!primed) { a = p[0]; f = 1; }
  for (; f < frames; ++f) a = a + alpha * (p[(size_t) f * N] - a);
  avg[(size_t) s * N + k] = a;
}

// exclusive scan of one int per thread over a 1024-thread CTA; *total = sum.  s_w: 32 ints of scratch.
static __device__ int block_excl_scan(int v, int *total, int *s_w)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = s_w[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    s_w[lane] = w;
  }
  __syncthreads();
  const int off = warp ? s_w[warp - 1] : 0;
  *total = s_w[31];
  __syncthreads();
  return off + incl - v;
}

struct ChdetK {
  const float *avg; float *n0; int *n0_primed;
  int N; float gamma, snr; int min_bins;
  SdbDetectedDev *out; unsigned *count, *total;
};

__global__ void __launch_bounds__(1024) k_chdet_find(const ChdetK p)
{
  extern __shared__ int s_ev[];                 // starts[CHDET_MAXRAW], ends[CHDET_MAXRAW]
  int *starts = s_ev, *ends = s_ev + CHDET_MAXRAW;
  __shared__ unsigned hist[256];
  __shared__ unsigned s_bucket, s_rank;
  __shared__ int s_w[32];
  __shared__ float s_n0;
  const int tid = threadIdx.x, st = blockIdx.x, N = p.N, half = N >> 1;
  const float *__restrict__ avg = p.avg + (size_t) st * N;

  // ---- K.2: exact (N/4)-th smallest value (non-negative floats order like their bit patterns)
  unsigned prefix = 0, mask = 0, rank = (unsigned) N / 4;
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < N; i += 1024) {
      const unsigned u = __float_as_uint(avg[i]);
      if ((u & mask) == prefix) atomicAdd(&hist[(u >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned cum = 0, b = 0;
      for (; b < 255; ++b) { if (cum + hist[b] > rank) break; cum += hist[b]; }
      s_bucket = b; s_rank = rank - cum;
    }
    __syncthreads();
    prefix |= s_bucket << shift; mask |= 0xffu << shift; rank = s_rank;
    __syncthreads();
  }
  if (tid == 0) {
    const float inst = __uint_as_float(prefix);
    float n0 = p.n0[st];
    n0 = p.n0_primed[st] ? n0 + p.gamma * (inst - n0) : inst;
    p.n0[st] = n0; p.n0_primed[st] = 1;
    s_n0 = n0;
  }
  __syncthreads();
  const float n0 = s_n0, thr = n0 * p.snr;

  // ---- K.3: start / end events of the runs above the threshold, ascending frequency j = (k + N/2) mod N
  const int L = (N + 1023) / 1024;
  const int i0 = min(tid * L, N), i1 = min(i0 + L, N);
  int ns = 0, ne = 0;
  {
    bool prev = i0 > 0 && i0 < N && avg[(i0 - 1 + half) & (N - 1)] > thr;
    for (int i = i0; i < i1; ++i) {
      const bool cur = avg[(i + half) & (N - 1)] > thr;
      ns += cur && !prev; ne += !cur && prev;
      prev = cur;
    }
    if (i1 == N && i0 < N && prev) ++ne;       // a run that reaches the top edge ends at N
  }
  int tot_s, tot_e;
  int os = block_excl_scan(ns, &tot_s, s_w);
  int oe = block_excl_scan(ne, &tot_e, s_w);
  {
    bool prev = i0 > 0 && i0 < N && avg[(i0 - 1 + half) & (N - 1)] > thr;
    for (int i = i0; i < i1; ++i) {
      const bool cur = avg[(i + half) & (N - 1)] > thr;
      if (cur && !prev) { if (os < CHDET_MAXRAW) starts[os] = i; ++os; }
      if (!cur && prev) { if (oe < CHDET_MAXRAW) ends[oe] = i; ++oe; }
      prev = cur;
    }
    if (i1 == N && i0 < N && prev) { if (oe < CHDET_MAXRAW) ends[oe] = N; ++oe; }
  }
  __syncthreads();
  const int R = min(tot_s, CHDET_MAXRAW);

  // ---- K.4: width filter + ordered compaction
  SdbDetectedDev *__restrict__ out = p.out + (size_t) st * CHDET_CAP;
  int running = 0;
  for (int base = 0; base < R; base += 1024) {
    const int r = base + tid;
    const int keep = r < R && ends[r] - starts[r] >= p.min_bins;
    int tile;
    const int pos = running + block_excl_scan(keep, &tile, s_w);
    if (keep && pos < CHDET_CAP) { out[pos].bin_lo = (unsigned) starts[r]; out[pos].bin_hi = (unsigned) ends[r]; }
    running += tile;
  }
  __syncthreads();
  const int nch = min(running, CHDET_CAP);
  if (tid == 0) { p.count[st] = (unsigned) nch; p.total[st] = (unsigned) running; }
  // ---- peak level of every reported channel: one warp per channel
  for (int c = tid >> 5; c < nch; c += 32) {
    const int a = (int) out[c].bin_lo, b = (int) out[c].bin_hi;
    float m = 0.0f;
    for (int i = a + (tid & 31); i < b; i += 32) m = fmaxf(m, avg[(i + half) & (N - 1)]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0) { out[c].s0 = m; out[c].n0 = n0; out[c].snr = m / n0; }
  }
}

// ---------------------------------------------------------------------------------------------
// host object + C-ABI
// ---------------------------------------------------------------------------------------------
static thread_local std::string g_chdet_err;
extern "C" const char *sdb_last_error(void);

struct sdb_chdet {
  int device = 0; unsigned N = 0, S = 0, min_bins = 1;
  float alpha = 0, gamma = 0, snr = 0;
  float *d_avg = nullptr, *d_n0 = nullptr; int *d_n0_primed = nullptr;
  SdbDetectedDev *d_out = nullptr; unsigned *d_count = nullptr, *d_total = nullptr;
  bool primed = false;
};

extern "C" void sdb_chdet_destroy(sdb_chdet_t *d)
{
  if (!d) return;
  cudaSetDevice(d->device);
  cudaFree(d->d_avg); cudaFree(d->d_n0); cudaFree(d->d_n0_primed); cudaFree(d->d_out); cudaFree(d->d_count);
  cudaFree(d->d_total);
  delete d;
}

extern "C" sdb_chdet_t *sdb_chdet_new(int device, uint32_t n_bins, uint32_t n_streams, float alpha, float gamma,
                                      float snr, uint32_t min_bins)
{
  if (sdb_device_count() <= 0) return nullptr;                        // no CPU fallback
  if (n_bins < 16 || (n_bins & (n_bins - 1)) || n_bins > (1u << 20) || n_streams < 1) return nullptr;
  if (!(snr > 0.0f) || !(alpha >= 0.0f && alpha <= 1.0f) || !(gamma >= 0.0f && gamma <= 1.0f)) return nullptr;
  if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  sdb_chdet *d = new sdb_chdet();
  d->device = device; d->N = n_bins; d->S = n_streams; d->alpha = alpha; d->gamma = gamma; d->snr = snr;
  d->min_bins = min_bins < 1 ? 1 : min_bins;
  const size_t S = n_streams;
  bool ok = cudaMalloc(&d->d_avg, S * n_bins * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&d->d_n0, S * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&d->d_n0_primed, S * sizeof(int)) == cudaSuccess &&
            cudaMalloc(&d->d_out, S * CHDET_CAP * sizeof(SdbDetectedDev)) == cudaSuccess &&
            cudaMalloc(&d->d_count, S * sizeof(unsigned)) == cudaSuccess &&
            cudaMalloc(&d->d_total, S * sizeof(unsigned)) == cudaSuccess;
  if (ok) {
    cudaMemset(d->d_avg, 0, S * n_bins * sizeof(float));
    cudaMemset(d->d_n0, 0, S * sizeof(float));
    cudaMemset(d->d_n0_primed, 0, S * sizeof(int));
    cudaMemset(d->d_count, 0, S * sizeof(unsigned));
    cudaMemset(d->d_total, 0, S * sizeof(unsigned));
    static std::atomic<unsigned long long> attr_done{ 0 };   // one bit per device: function attributes are per context
    if (sdb_first_on_device(attr_done)) {
      cudaFuncSetAttribute(k_chdet_find, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * CHDET_MAXRAW * (int) sizeof(int));
    }
  }
  if (!ok || cudaGetLastError() != cudaSuccess) { sdb_chdet_destroy(d); return nullptr; }
  return d;
}

cudaError_t sdb_chdet_feed_stream(sdb_chdet *d, const float *psd_dev, uint32_t frames, size_t stream_stride,
                                  cudaStream_t stream, uint64_t *launch_counter)
{
  if (frames == 0) return cudaSuccess;
  dim3 grid((d->N + 255) / 256, d->S);
  k_chdet_avg<<<grid, 256, 0, stream>>>(d->d_avg, psd_dev, (int) d->N, (int) frames, stream_stride, d->alpha,
                                        d->primed ? 1 : 0);
  d->primed = true;
  ChdetK p;
  p.avg = d->d_avg; p.n0 = d->d_n0; p.n0_primed = d->d_n0_primed; p.N = (int) d->N; p.gamma = d->gamma; p.snr = d->snr;
  p.min_bins = (int) d->min_bins; p.out = d->d_out; p.count = d->d_count; p.total = d->d_total;
  k_chdet_find<<<d->S, 1024, 2 * CHDET_MAXRAW * sizeof(int), stream>>>(p);
  if (launch_counter) *launch_counter += 2;
  return cudaGetLastError();
}

extern "C" int sdb_chdet_feed_device(sdb_chdet_t *d, const float *psd_dev, uint32_t frames, size_t stream_stride)
{
  if (!d || !psd_dev) return -1;
  if (cudaSetDevice(d->device) != cudaSuccess) return -1;
  if (sdb_chdet_feed_stream(d, psd_dev, frames, stream_stride, nullptr, nullptr) != cudaSuccess) return -1;
  return cudaStreamSynchronize(nullptr) == cudaSuccess ? 0 : -1;
}

// every stream's list in three copies instead of three per stream (the panoramic sweep reads 128...1024 hops):
// centers[S] host, out[S][cap], counts[S] (clamped to cap)
extern "C" int sdb_chdet_read_all(sdb_chdet_t *d, double samp_rate, const double *centers, sdb_detected_channel *out,
                                  size_t cap, uint32_t *counts)
{
  if (!d || !centers || !out || !counts) return -1;
  if (cudaSetDevice(d->device) != cudaSuccess) return -1;
  const size_t S = d->S;
  std::vector<unsigned> cnt(S);
  std::vector<SdbDetectedDev> h(S * CHDET_CAP);
  if (cudaMemcpy(cnt.data(), d->d_count, S * sizeof(unsigned), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (cudaMemcpy(h.data(), d->d_out, h.size() * sizeof(SdbDetectedDev), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  const double df = samp_rate / (double) d->N, half = (double) (d->N / 2);
  for (size_t s = 0; s < S; ++s) {
    const size_t n = cnt[s] < cap ? cnt[s] : cap;
    counts[s] = (uint32_t) n;
    for (size_t i = 0; i < n; ++i) {
      const SdbDetectedDev &q = h[s * CHDET_CAP + i];
      sdb_detected_channel &c = out[s * cap + i];
      c.bin_lo = q.bin_lo; c.bin_hi = q.bin_hi;
      c.f_lo = centers[s] + ((double) q.bin_lo - half) * df;
      c.f_hi = centers[s] + ((double) q.bin_hi - half) * df;
      c.fc = 0.5 * (c.f_lo + c.f_hi); c.bw = c.f_hi - c.f_lo;
      c.S0 = q.s0; c.N0 = q.n0; c.snr = q.snr;
    }
  }
  return 0;
}

extern "C" long sdb_chdet_read(sdb_chdet_t *d, uint32_t stream, double samp_rate, double center_freq,
                               sdb_detected_channel *out, size_t cap, uint32_t *total)
{
  if (!d || stream >= d->S || (!out && cap)) return -1;
  if (cudaSetDevice(d->device) != cudaSuccess) return -1;
  unsigned cnt = 0, tot = 0;
  if (cudaMemcpy(&cnt, d->d_count + stream, sizeof(cnt), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (cudaMemcpy(&tot, d->d_total + stream, sizeof(tot), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  if (total) *total = tot;
  std::vector<SdbDetectedDev> h(cnt);
  if (cnt && cudaMemcpy(h.data(), d->d_out + (size_t) stream * CHDET_CAP, cnt * sizeof(SdbDetectedDev),
                        cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
  const size_t n = cnt < cap ? cnt : cap;
  const double df = samp_rate / (double) d->N, half = (double) (d->N / 2);
  for (size_t i = 0; i < n; ++i) {
    sdb_detected_channel &c = out[i];
    c.bin_lo = h[i].bin_lo; c.bin_hi = h[i].bin_hi;
    c.f_lo = center_freq + ((double) h[i].bin_lo - half) * df;       // bin edges: bin j spans [j, j+1) df
    c.f_hi = center_freq + ((double) h[i].bin_hi - half) * df;
    c.fc = 0.5 * (c.f_lo + c.f_hi); c.bw = c.f_hi - c.f_lo;
    c.S0 = h[i].s0; c.N0 = h[i].n0; c.snr = h[i].snr;
  }
  return (long) n;
}
