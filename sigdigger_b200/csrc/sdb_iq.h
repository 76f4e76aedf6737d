// sdb_iq.h -- native SDR sample formats, converted to complex float32 inside the first load of the path
// (no separate conversion pass: 8/16-bit sources cost 2-4x fewer ingest bytes).
// Formats mirror SUSCAN_SOURCE_FORMAT_RAW_{FLOAT32, UNSIGNED8, SIGNED8, SIGNED16} as selectable at
// Default/SourceConfig/FileSourcePage.cpp:80-104.  SPEC.md section Q: u8 -> (v - 128) / 128, s8 -> v / 128,
// s16 -> v / 32768, all exact in binary32.
#pragma once
#include <cuda_runtime.h>

enum { SDB_FMT_F32 = 0, SDB_FMT_U8 = 1, SDB_FMT_S8 = 2, SDB_FMT_S16 = 3 };
// modifier bit: deliver (Q, I) instead of (I, Q) -- suscan_analyzer_set_iq_reverse (Suscan/Analyzer.cpp:238-244).
// A format with the bit set is not SDB_FMT_F32, so the float32 fast paths of the kernels fall through to sdb_ld_iq.
#define SDB_FMT_SWAP 0x10
#define SDB_FMT_BASE(fmt) ((fmt) & 0xf)

static __host__ __device__ inline size_t sdb_fmt_bytes(int fmt)
{
  return SDB_FMT_BASE(fmt) == SDB_FMT_F32 ? 8 : (SDB_FMT_BASE(fmt) == SDB_FMT_S16 ? 4 : 2);
}

#ifdef __CUDACC__
static __device__ __forceinline__ float2 sdb_ld_iq_base(const void *__restrict__ base, long idx, int fmt);
static __device__ __forceinline__ float2 sdb_ld_iq(const void *__restrict__ base, long idx, int fmt)
{
  const float2 v = sdb_ld_iq_base(base, idx, SDB_FMT_BASE(fmt));
  return (fmt & SDB_FMT_SWAP) ? make_float2(v.y, v.x) : v;
}
static __device__ __forceinline__ float2 sdb_ld_iq_base(const void *__restrict__ base, long idx, int fmt)
{
  if (fmt == SDB_FMT_F32) return __ldg(reinterpret_cast<const float2 *>(base) + idx);
  if (fmt == SDB_FMT_S16) {
    const short2 v = __ldg(reinterpret_cast<const short2 *>(base) + idx);
    return make_float2((float) v.x * (1.0f / 32768.0f), (float) v.y * (1.0f / 32768.0f));
  }
  if (fmt == SDB_FMT_U8) {
    const uchar2 v = __ldg(reinterpret_cast<const uchar2 *>(base) + idx);
    return make_float2(((float) v.x - 128.0f) * (1.0f / 128.0f), ((float) v.y - 128.0f) * (1.0f / 128.0f));
  }
  const char2 v = __ldg(reinterpret_cast<const char2 *>(base) + idx);
  return make_float2((float) v.x * (1.0f / 128.0f), (float) v.y * (1.0f / 128.0f));
}
#endif
