// panoramic.cu -- the panoramic scanner's sweep behind the C-ABI (BASELINE.json configs[4]).
//
// What the reference does on the GUI thread per PSD message of a wide-spectrum analyzer
// (Panoramic/Scanner.cpp:503-523: view.feed(psd, nullptr, fftSize, fc); constructor / hop plan :296-372), spread over
// the GPUs of one node, one process per GPU:
//
//     rank r : hops [lo_r, hi_r)  --PSD (engine)-->  SpectrumView projection  --contribution lists--+
//                                  `-> per-hop channel detector (optional)                          | gather to rank 0
//     rank 0 : accumulate(all contribution lists, in global hop order) + gap fill  <----------------+ (NCCL, NVLink)
//
// Per-bin state of the SpectrumView depends only on that bin's own contributions in hop order (SPEC V), so applying
// the gathered lists in rank order (= hop order: shards are contiguous) reproduces the reference's sequential
// feed() value by value.  This gather is the ONLY collective on the path (the PSDs never move): every rank packs
// {j0, nb, va, vc (, channel lists)} of its shard into one device buffer and rank 0 posts one ncclRecv per peer
// inside a group (ncclSend on the peers).  NCCL is bound at run time (dlopen "libnccl.so.2": the process that
// initialised torch.distributed has it loaded already); a single rank needs no NCCL at all.
#include "../../include/sigdigger_b200.h"
#include <cuda_runtime.h>
#include <nccl.h>
#include <dlfcn.h>

#include <algorithm>
#include <string>
#include <vector>
#include <string.h>

static thread_local std::string g_perr;
static int pfail(const std::string &m) { g_perr = m; return -1; }
extern "C" const char *sdb_panoramic_last_error(void) { return g_perr.c_str(); }

namespace {
struct Nccl {
  void *h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  bool load()
  {
    if (h) return true;
    const char *names[] = { getenv("SDB_NCCL_LIB"), "libnccl.so.2", "libnccl.so" };
    for (const char *n : names) { if (n && (h = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break; }
    if (!h) return false;
#define SYM(f, s) f = (decltype(f)) dlsym(h, s)
    SYM(GetUniqueId, "ncclGetUniqueId"); SYM(CommInitRank, "ncclCommInitRank"); SYM(CommDestroy, "ncclCommDestroy");
    SYM(GroupStart, "ncclGroupStart"); SYM(GroupEnd, "ncclGroupEnd"); SYM(Send, "ncclSend"); SYM(Recv, "ncclRecv");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    return GetUniqueId && CommInitRank && CommDestroy && GroupStart && GroupEnd && Send && Recv;
  }
};
Nccl g_nccl;

void shard(size_t n_hops, int world, int rank, size_t *lo, size_t *hi)
{
  const size_t base = n_hops / (size_t) world, rem = n_hops % (size_t) world;
  *lo = (size_t) rank * base + std::min<size_t>((size_t) rank, rem);
  *hi = *lo + base + ((size_t) rank < rem ? 1 : 0);
}
}  // namespace

struct sdb_panoramic {
  sdb_panoramic_params prm;
  int rank = 0, world = 1;
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;
  sdb_engine_t *eng = nullptr; size_t eng_hops = 0;
  sdb_sview_t *view = nullptr;
  // packed exchange buffers: one row block per rank, pl rows each
  size_t pl = 0, mb = 0, row_bytes = 0;
  unsigned char *d_send = nullptr, *d_recv = nullptr; size_t send_cap = 0, recv_cap = 0;
  std::vector<unsigned char> h_pack;
  float *d_db = nullptr; size_t db_cap = 0;
  int32_t *d_j0 = nullptr, *d_nb = nullptr; float *d_va = nullptr, *d_vc = nullptr; size_t lists_cap = 0;
  std::vector<std::vector<sdb_detected_channel>> channels;   // rank 0: per hop, after a sweep with the detector
  cudaEvent_t ev[4] = { nullptr, nullptr, nullptr, nullptr };
  sdb_panoramic_timing last{};
};

#define PCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return pfail(std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)
#define NCK(x) do { ncclResult_t r_ = (x); if (r_ != ncclSuccess) return pfail(std::string(#x) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "nccl error")); } while (0)

extern "C" int sdb_panoramic_unique_id(void *id128)
{
  if (!id128) return pfail("null id");
  if (!g_nccl.load()) return pfail("libnccl.so.2 not found (set SDB_NCCL_LIB)");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  NCK(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

extern "C" sdb_panoramic_t *sdb_panoramic_new(const sdb_panoramic_params *p, int rank, int world, const void *id128)
{
  if (!p || world < 1 || rank < 0 || rank >= world) { g_perr = "invalid arguments"; return nullptr; }
  if (sdb_device_count() <= 0) { g_perr = "no CUDA device: sigdigger_b200 has no CPU fallback"; return nullptr; }
  if (!(p->freq_max > p->freq_min) || !(p->fft_bandwidth > 0) || p->psd_size < 16 || (p->psd_size & (p->psd_size - 1))) {
    g_perr = "invalid sweep parameters"; return nullptr;
  }
  if (cudaSetDevice(p->device) != cudaSuccess) { g_perr = "cudaSetDevice failed"; return nullptr; }
  sdb_panoramic *s = new sdb_panoramic();
  s->prm = *p; s->rank = rank; s->world = world;
  if (s->prm.channel_cap == 0) s->prm.channel_cap = 64;
  if (world > 1) {
    if (!id128 || !g_nccl.load()) { g_perr = "NCCL unavailable or no unique id"; delete s; return nullptr; }
    ncclUniqueId id; memcpy(&id, id128, sizeof(id));
    if (g_nccl.CommInitRank(&s->comm, world, id, rank) != ncclSuccess) { g_perr = "ncclCommInitRank failed"; delete s; return nullptr; }
  }
  cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking);
  for (auto &e : s->ev) cudaEventCreate(&e);
  s->view = sdb_sview_new(p->device);
  if (!s->view || sdb_sview_set_range(s->view, p->freq_min, p->freq_max, p->fft_bandwidth, p->rel_bw > 0 ? p->rel_bw : 0.5f)) {
    g_perr = "SpectrumView set-up failed"; sdb_panoramic_destroy(s); return nullptr;
  }
  s->mb = sdb_sview_max_bins(s->view);
  return s;
}

extern "C" void sdb_panoramic_destroy(sdb_panoramic_t *s)
{
  if (!s) return;
  cudaSetDevice(s->prm.device);
  if (s->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(s->comm);
  if (s->eng) sdb_engine_destroy(s->eng);
  if (s->view) sdb_sview_destroy(s->view);
  cudaFree(s->d_send); cudaFree(s->d_recv); cudaFree(s->d_db);
  cudaFree(s->d_j0); cudaFree(s->d_nb); cudaFree(s->d_va); cudaFree(s->d_vc);
  for (auto &e : s->ev) if (e) cudaEventDestroy(e);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
}

// layout of one rank's packed block (pl rows): j0[pl] i32 | nb[pl] i32 | va[pl][mb] f32 | vc[pl][mb] f32
//                                              | ccnt[pl] i32 | chan[pl][cap] sdb_detected_channel   (detector only)
static size_t block_bytes(const sdb_panoramic *s)
{
  size_t b = s->pl * (2 * sizeof(int32_t) + 2 * s->mb * sizeof(float));
  if (s->prm.detect) b += s->pl * (sizeof(int32_t) + (size_t) s->prm.channel_cap * sizeof(sdb_detected_channel));
  return (b + 15) & ~(size_t) 15;
}

static int ensure(sdb_panoramic *s, size_t n_hops, size_t n_local)
{
  s->pl = (n_hops + (size_t) s->world - 1) / (size_t) s->world;
  const size_t bb = block_bytes(s);
  if (s->send_cap < bb) { cudaFree(s->d_send); PCK(cudaMalloc(&s->d_send, bb)); s->send_cap = bb; }
  if (s->rank == 0 && s->recv_cap < bb * (size_t) s->world) {
    cudaFree(s->d_recv); PCK(cudaMalloc(&s->d_recv, bb * (size_t) s->world)); s->recv_cap = bb * (size_t) s->world;
  }
  if (s->rank == 0 && s->lists_cap < n_hops) {
    cudaFree(s->d_j0); cudaFree(s->d_nb); cudaFree(s->d_va); cudaFree(s->d_vc);
    PCK(cudaMalloc(&s->d_j0, n_hops * sizeof(int32_t))); PCK(cudaMalloc(&s->d_nb, n_hops * sizeof(int32_t)));
    PCK(cudaMalloc(&s->d_va, n_hops * s->mb * sizeof(float))); PCK(cudaMalloc(&s->d_vc, n_hops * s->mb * sizeof(float)));
    s->lists_cap = n_hops;
  }
  if (n_local && (!s->eng || s->eng_hops != n_local)) {
    if (s->eng) sdb_engine_destroy(s->eng);
    sdb_engine_params ep; memset(&ep, 0, sizeof(ep));
    ep.n_streams = (uint32_t) n_local; ep.psd_size = s->prm.psd_size; ep.psd_window = s->prm.psd_window;
    ep.max_feed = s->prm.psd_size * (s->prm.frames_per_hop ? s->prm.frames_per_hop : 1); ep.device = s->prm.device;
    ep.flags = s->prm.detect ? 0 : SDB_FLAG_PSD_SHIFT_DB;       // the detector reads the linear PSD
    s->eng = sdb_engine_new(&ep, s->prm.fft_bandwidth);
    if (!s->eng) return pfail(sdb_last_error());
    if (s->prm.detect &&
        sdb_engine_set_channel_detector(s->eng, s->prm.det_alpha, 0.0f, s->prm.det_gamma, s->prm.det_snr,
                                        s->prm.det_min_bins ? s->prm.det_min_bins : 2))
      return pfail(sdb_last_error());
    if (sdb_engine_commit(s->eng)) return pfail(sdb_last_error());
    s->eng_hops = n_local;
    if ((s->prm.detect || s->prm.frames_per_hop > 1) && s->db_cap < n_local * s->prm.psd_size) {
      cudaFree(s->d_db); PCK(cudaMalloc(&s->d_db, n_local * s->prm.psd_size * sizeof(float))); s->db_cap = n_local * s->prm.psd_size;
    }
  }
  return 0;
}

// hops_local: this rank's contiguous shard [hi - lo][psd_size] (device pointer if on_device), one window per hop
static int sweep_impl(sdb_panoramic *s, const sdb_complex *hops_local, int on_device, const double *centers_all,
                      size_t n_hops)
{
  if (!s || !centers_all || n_hops == 0) return pfail("invalid arguments");
  PCK(cudaSetDevice(s->prm.device));
  size_t lo, hi;
  shard(n_hops, s->world, s->rank, &lo, &hi);
  const size_t n_local = hi - lo, N = s->prm.psd_size;
  if (n_local && !hops_local) return pfail("null hop buffer");
  if (ensure(s, n_hops, n_local)) return -1;
  const size_t bb = block_bytes(s), pl = s->pl, mb = s->mb;
  PCK(cudaEventRecord(s->ev[0], s->stream));
  PCK(cudaMemsetAsync(s->d_send, 0, bb, s->stream));
  int32_t *sj0 = (int32_t *) s->d_send, *snb = sj0 + pl;
  float *sva = (float *) (snb + pl), *svc = sva + pl * mb;
  int32_t *scnt = (int32_t *) (svc + pl * mb);
  sdb_detected_channel *sch = (sdb_detected_channel *) (scnt + pl);
  PCK(cudaStreamSynchronize(s->stream));
  if (n_local) {
    // frames_per_hop > 1: the detector averages over the hop's frames, the view takes the last one
    const size_t F = s->prm.frames_per_hop ? s->prm.frames_per_hop : 1, L = N * F;
    if (on_device ? sdb_engine_feed_device(s->eng, hops_local, L, L) : sdb_engine_feed_host(s->eng, hops_local, L, L))
      return pfail(sdb_last_error());
    if (sdb_engine_sync(s->eng)) return pfail(sdb_last_error());
    const float *psd = sdb_engine_psd_device(s->eng);
    if (s->prm.detect) {
      if (F == 1) {
        if (sdb_psd_shift_db_device(psd, s->d_db, n_local, (uint32_t) N)) return pfail(sdb_last_error());
      } else {
        for (size_t h = 0; h < n_local; ++h)
          if (sdb_psd_shift_db_device(psd + (h * F + F - 1) * N, s->d_db + h * N, 1, (uint32_t) N)) return pfail(sdb_last_error());
      }
      psd = s->d_db;
    } else if (F > 1) {
      PCK(cudaMemcpy2D(s->d_db, N * sizeof(float), psd + (F - 1) * N, L * sizeof(float), N * sizeof(float), n_local,
                       cudaMemcpyDeviceToDevice));
      psd = s->d_db;
    }
    if (sdb_sview_project(s->view, psd, N, centers_all + lo, n_local, 1)) return pfail(sdb_last_error());
    if (sdb_sview_contrib_copy(s->view, sj0, snb, sva, svc, n_local)) return pfail(sdb_last_error());
    if (s->prm.detect) {
      const size_t cap = s->prm.channel_cap;
      std::vector<int32_t> cnt(pl, 0);
      std::vector<sdb_detected_channel> ch(pl * cap);
      if (sdb_engine_read_all_channels(s->eng, centers_all + lo, ch.data(), cap, (uint32_t *) cnt.data()))
        return pfail(sdb_last_error());
      PCK(cudaMemcpy(scnt, cnt.data(), pl * sizeof(int32_t), cudaMemcpyHostToDevice));
      PCK(cudaMemcpy(sch, ch.data(), pl * cap * sizeof(sdb_detected_channel), cudaMemcpyHostToDevice));
    }
  }
  PCK(cudaDeviceSynchronize());
  PCK(cudaEventRecord(s->ev[1], s->stream));
  // ---- the one exchange of the path: gather the packed blocks on rank 0
  if (s->world > 1) {
    NCK(g_nccl.GroupStart());
    if (s->rank == 0) {
      for (int r = 1; r < s->world; ++r)
        NCK(g_nccl.Recv(s->d_recv + (size_t) r * bb, bb, ncclUint8, r, s->comm, s->stream));
    } else {
      NCK(g_nccl.Send(s->d_send, bb, ncclUint8, 0, s->comm, s->stream));
    }
    NCK(g_nccl.GroupEnd());
  }
  if (s->rank == 0) PCK(cudaMemcpyAsync(s->d_recv, s->d_send, bb, cudaMemcpyDeviceToDevice, s->stream));
  PCK(cudaEventRecord(s->ev[2], s->stream));
  PCK(cudaStreamSynchronize(s->stream));
  s->last.gather_bytes = s->world > 1 ? (uint64_t) bb * (uint64_t) (s->world - 1) : 0;
  if (s->rank == 0) {
    // unpack in rank order = global hop order, then accumulate + fill exactly as one sequential feed() series
    size_t row = 0;
    if (s->prm.detect) s->channels.assign(n_hops, {});
    for (int r = 0; r < s->world; ++r) {
      size_t rlo, rhi;
      shard(n_hops, s->world, r, &rlo, &rhi);
      const size_t cnt = rhi - rlo;
      if (!cnt) continue;
      const unsigned char *blk = s->d_recv + (size_t) r * bb;
      const int32_t *bj0 = (const int32_t *) blk, *bnb = bj0 + pl;
      const float *bva = (const float *) (bnb + pl), *bvc = bva + pl * mb;
      PCK(cudaMemcpyAsync(s->d_j0 + row, bj0, cnt * sizeof(int32_t), cudaMemcpyDeviceToDevice, s->stream));
      PCK(cudaMemcpyAsync(s->d_nb + row, bnb, cnt * sizeof(int32_t), cudaMemcpyDeviceToDevice, s->stream));
      PCK(cudaMemcpyAsync(s->d_va + row * mb, bva, cnt * mb * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
      PCK(cudaMemcpyAsync(s->d_vc + row * mb, bvc, cnt * mb * sizeof(float), cudaMemcpyDeviceToDevice, s->stream));
      if (s->prm.detect) {
        const size_t cap = s->prm.channel_cap;
        std::vector<int32_t> c(cnt); std::vector<sdb_detected_channel> ch(cnt * cap);
        const int32_t *bcnt = (const int32_t *) (bvc + pl * mb);
        const sdb_detected_channel *bch = (const sdb_detected_channel *) (bcnt + pl);
        PCK(cudaMemcpy(c.data(), bcnt, cnt * sizeof(int32_t), cudaMemcpyDeviceToHost));
        PCK(cudaMemcpy(ch.data(), bch, cnt * cap * sizeof(sdb_detected_channel), cudaMemcpyDeviceToHost));
        for (size_t h = 0; h < cnt; ++h) s->channels[row + h].assign(&ch[h * cap], &ch[h * cap] + c[h]);
      }
      row += cnt;
    }
    PCK(cudaStreamSynchronize(s->stream));
    if (sdb_sview_accumulate(s->view, s->d_j0, s->d_nb, s->d_va, s->d_vc, n_hops)) return pfail(sdb_last_error());
    PCK(cudaDeviceSynchronize());
  }
  PCK(cudaEventRecord(s->ev[3], s->stream));
  PCK(cudaEventSynchronize(s->ev[3]));
  float a = 0, b = 0, c = 0;
  cudaEventElapsedTime(&a, s->ev[0], s->ev[1]); cudaEventElapsedTime(&b, s->ev[1], s->ev[2]); cudaEventElapsedTime(&c, s->ev[2], s->ev[3]);
  s->last.psd_project_ms = a; s->last.gather_ms = b; s->last.accumulate_ms = c;
  s->last.n_hops_local = n_local;
  return 0;
}

extern "C" int sdb_panoramic_sweep_device(sdb_panoramic_t *s, const sdb_complex *hops_local_dev, const double *centers_all,
                                          size_t n_hops)
{ return sweep_impl(s, hops_local_dev, 1, centers_all, n_hops); }
extern "C" int sdb_panoramic_sweep_host(sdb_panoramic_t *s, const sdb_complex *hops_local, const double *centers_all,
                                        size_t n_hops)
{ return sweep_impl(s, hops_local, 0, centers_all, n_hops); }

extern "C" int sdb_panoramic_reset(sdb_panoramic_t *s) { return s ? sdb_sview_reset(s->view) : pfail("null"); }
extern "C" int sdb_panoramic_read(sdb_panoramic_t *s, float *psd, float *accum, float *count, size_t cap)
{
  if (!s) return pfail("null");
  if (s->rank != 0) return pfail("the stitched view lives on rank 0");
  return sdb_sview_read(s->view, psd, accum, count, cap) ? pfail(sdb_last_error()) : 0;
}
extern "C" uint32_t sdb_panoramic_size(const sdb_panoramic_t *s) { return s ? sdb_sview_size(s->view) : 0; }
extern "C" long sdb_panoramic_read_channels(sdb_panoramic_t *s, size_t hop, sdb_detected_channel *out, size_t cap)
{
  if (!s || s->rank != 0 || hop >= s->channels.size()) return -1;
  const size_t n = std::min(cap, s->channels[hop].size());
  if (n && out) memcpy(out, s->channels[hop].data(), n * sizeof(sdb_detected_channel));
  return (long) n;
}
extern "C" int sdb_panoramic_last_timing(const sdb_panoramic_t *s, sdb_panoramic_timing *t)
{
  if (!s || !t) return -1;
  *t = s->last;
  return 0;
}
extern "C" void sdb_panoramic_shard(size_t n_hops, int world, int rank, size_t *lo, size_t *hi) { shard(n_hops, world, rank, lo, hi); }
