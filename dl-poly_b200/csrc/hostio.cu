// hostio.cu -- how the drop-in entry points move the caller's corePart records (include/dlpgpu.h: dlpgpu_link_cell_pairs,
// dlpgpu_two_body_forces, dlpgpu_vnl_check, dlpgpu_vnl_set_check, dlpgpu_spme_forces).
//
// The reference keeps positions, forces and the charge of an atom in one 64-byte record (particle.F90 corePart) and every
// force provider ADDS into parts%f (drivers.F90:655-660).  Of those 64 bytes the device needs 32 per step (x, y, z, charge
// of parts(1:nlast)) and hands back 24 (the force it computed for parts(1:natms)).  Two ways to move them:
//
// * whole records (dlpgpu_set_host_threads(ctx, 0), the default): parts(1:nlast) go up as they are from the caller's array,
//   which is page-locked on first sight; the device adds its forces to the copy and parts(1:natms) come back in one copy.
//   148 MB per 1 M-ion step, 2.7 ms at the 55 GB/s of the pool's boxes, no host work.  Strided DMA or zero-copy access to the
//   useful bytes is slower still (scripts/pcie_test.cu).
// * packed (dlpgpu_set_host_threads(ctx, n >= 1)): n host threads of the library (de)interleave the fields in chunks that
//   overlap with the DMA engine:
//     up:    workers copy {x, y, z} of a chunk of records into a page-locked staging buffer (non-temporal stores) and compare
//            the charge with the copy the device holds (charges only change place when the local order changes); the
//            calling thread queues the chunk's H2D copy -- and the chunk's charges if any differed -- as soon as the chunk is
//            packed; one small kernel merges coordinates and charges into the device's posq array;
//     down:  the device interleaves its force arrays into {fx, fy, fz} triples, the calling thread queues one D2H copy and
//            one event per chunk, workers add a chunk's triples into parts%f as soon as its event has completed -- the same
//            one rounding per component as the reference's accumulation into parts%f.  The device never sees parts%f.
//   56 MB per 1 M-ion step over PCIe, but 280 MB through the host's memory system: it pays when a rank has a dozen otherwise
//   idle cores (measured on the pool's 16-vCPU boxes, 1 M ions: 15.4 / 9.5 / 6.8 / 5.2 / 4.2 ms per step with 1 / 2 / 4 / 8 /
//   12 threads against 4.6 ms for whole records; profiles/r2_s48_*).
// ltype / ltg / lfrzn of link_cell_pairs are compared with the page-locked copy of the previous call chunk by chunk in both
// modes; only chunks that differ are copied again (in a serial run the local atoms never change their order).
// Counters of the bytes that really crossed PCIe are kept for bench.py (dlpgpu_transfer_bytes).
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include <sched.h>
#if defined(__x86_64__)
#include <emmintrin.h>
#endif

#include "common.cuh"
#include "hostpool.h"

namespace {

using dlp_hostpool::HostPool;
using dlp_hostpool::MAXCH;

struct HostIO {
  HostPool* pool = nullptr;
  int threads = -1;             // -1: not chosen yet, 0: whole records by DMA, >= 1: packed fields through that many workers
  double* up = nullptr;  size_t up_cap = 0;     // page-locked {x, y, z} per record
  double* chg = nullptr; size_t chg_cap = 0;    // page-locked copy of the charges the device holds (they change with the local order only)
  int chg_n = 0;                                // leading records for which chg and chg_dev hold the same, valid charges
  const double* chg_dev_seen = nullptr;
  DBuf<double> xyz3, chg_dev;                   // device side of the two
  double* dn = nullptr;  size_t dn_cap = 0;     // page-locked {fx, fy, fz} per local atom
  int* ints = nullptr;   size_t ints_cap = 0;   // page-locked ltype | ltg | lfrzn of the previous link_cell_pairs call
  int ints_n = -1;                              // records the device copies of the three arrays are valid for
  bool frzn_given = false;
  DBuf<double> f3;                              // device-side interleaved forces
  DBuf<dlpgpu_corepart> parts_dev;              // whole-record mode: device copy of the caller's records
  void* reg_ptr = nullptr; size_t reg_bytes = 0; bool reg_ours = false;   // the caller's array, page-locked by us
  cudaEvent_t ev[MAXCH] = {};
  std::atomic<int> avail[MAXCH];
  unsigned long long h2d = 0, d2h = 0;
  double t_up = 0.0, t_wait = 0.0, t_down = 0.0;   // seconds: uploads (call to last copy queued), download: call to first chunk on the host, first chunk to return
  int n_up = 0, n_down = 0;
};

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int default_threads() {   // whole records unless the user opts in: a library cannot know how many cores its rank may take
  if (const char* e = getenv("DLPGPU_HOST_THREADS")) { const int v = atoi(e); if (v >= 0) return std::min(v, 64); }
  return 0;
}

HostIO* io_of(dlpgpu_ctx* ctx) {
  if (!ctx->hostio) {
    HostIO* io = new HostIO();
    for (int c = 0; c < MAXCH; ++c) io->avail[c].store(0, std::memory_order_relaxed);
    ctx->hostio = io;
  }
  HostIO* io = static_cast<HostIO*>(ctx->hostio);
  if (io->threads < 0) io->threads = default_threads();
  if (!io->pool || io->pool->threads() != io->threads) {
    delete io->pool;
    io->pool = new HostPool(io->threads);
  }
  return io;
}

template <typename T>
cudaError_t grow_pinned(T*& p, size_t& cap, size_t n) {
  if (n <= cap) return cudaSuccess;
  if (p) cudaFreeHost(p);
  p = nullptr; cap = 0;
  const size_t ncap = n + n / 8 + 1024;
  cudaError_t e = cudaHostAlloc((void**)&p, ncap * sizeof(T), cudaHostAllocDefault);
  if (e == cudaSuccess) cap = ncap;
  return e;
}

// chunk length: a multiple of 4096 records, about 60 chunks for a large array (never more than MAXCH), at least 16 k records
int chunk_len(int n) {
  int len = ((n / 60 + 4095) / 4096) * 4096;
  return std::max(len, 16384);
}

// the caller's corePart array (config%parts) lives at one address for the whole run: page-lock it on first sight so that the
// whole-record copies run at full PCIe rate; memory that is already pinned (or cannot be pinned) is used as it is
void pin_host_parts(HostIO* io, const void* p, size_t bytes) {
  if (!p || bytes == 0) return;
  if (io->reg_ptr == p && io->reg_bytes >= bytes) return;
  if (io->reg_ptr && io->reg_ours) cudaHostUnregister(io->reg_ptr);
  io->reg_ptr = const_cast<void*>(p); io->reg_bytes = bytes;
  cudaError_t e = cudaHostRegister(const_cast<void*>(p), bytes, cudaHostRegisterDefault);
  io->reg_ours = (e == cudaSuccess);
  if (e != cudaSuccess) cudaGetLastError();
}

__global__ void k_merge_xyzq(int n, const double* __restrict__ xyz3, const double* __restrict__ chg, double4* __restrict__ posq) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  posq[i] = make_double4(xyz3[3 * (size_t)i], xyz3[3 * (size_t)i + 1], xyz3[3 * (size_t)i + 2], chg[i]);
}
__global__ void k_unpack_parts(const dlpgpu_corepart* __restrict__ parts, int n, double4* __restrict__ posq) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  dlpgpu_corepart p = parts[i];
  posq[i] = make_double4(p.xxx, p.yyy, p.zzz, p.chge);
}
__global__ void k_add_forces(dlpgpu_corepart* __restrict__ parts, int n, const double* __restrict__ fx, const double* __restrict__ fy,
                             const double* __restrict__ fz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  parts[i].fxx += fx[i]; parts[i].fyy += fy[i]; parts[i].fzz += fz[i];
}

__global__ void k_pack_f3(int n, const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz, double* __restrict__ f3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  f3[3 * (size_t)i] = fx[i]; f3[3 * (size_t)i + 1] = fy[i]; f3[3 * (size_t)i + 2] = fz[i];
}

}  // namespace

void dlp_hostio_release(dlpgpu_ctx* ctx) {
  if (!ctx->hostio) return;
  HostIO* io = static_cast<HostIO*>(ctx->hostio);
  delete io->pool;
  if (io->up) cudaFreeHost(io->up);
  if (io->chg) cudaFreeHost(io->chg);
  io->xyz3.release(ctx->stream); io->chg_dev.release(ctx->stream);
  if (io->dn) cudaFreeHost(io->dn);
  if (io->ints) cudaFreeHost(io->ints);
  io->f3.release(ctx->stream); io->parts_dev.release(ctx->stream);
  if (io->reg_ptr && io->reg_ours) cudaHostUnregister(io->reg_ptr);
  for (int c = 0; c < MAXCH; ++c) if (io->ev[c]) cudaEventDestroy(io->ev[c]);
  delete io;
  ctx->hostio = nullptr;
}

// parts(1:n) -> posq(1:n) on the device (x, y, z, chge); queued on the context's stream, the staging buffer is free again
// once the stream has passed the copies (every drop-in entry point synchronises before it returns)
int dlp_upload_parts(dlpgpu_ctx* ctx, int n, const dlpgpu_corepart* parts) {
  if (n <= 0) return 0;
  HostIO* io = io_of(ctx);
  if (io->threads == 0) {   // whole records
    pin_host_parts(io, parts, (size_t)n * sizeof(dlpgpu_corepart));
    CK(io->parts_dev.ensure(n, ctx->stream));
    CK(cudaMemcpyAsync(io->parts_dev.p, parts, (size_t)n * sizeof(dlpgpu_corepart), cudaMemcpyHostToDevice, ctx->stream));
    LAUNCH(ctx, k_unpack_parts, cdiv(n, 256), 256, 0, io->parts_dev.p, n, ctx->posq.p);
    io->h2d += (unsigned long long)n * sizeof(dlpgpu_corepart);
    ctx->tol_fresh = false; ctx->pub_fresh = false;
    ctx->parts_resident = 0; ctx->parts_current = false;
    return 0;
  }
  if ((size_t)3 * n > io->up_cap || (size_t)n > io->chg_cap) {
    CK(cudaStreamSynchronize(ctx->stream));
    CK(grow_pinned(io->up, io->up_cap, (size_t)3 * n));
    if ((size_t)n > io->chg_cap) { CK(grow_pinned(io->chg, io->chg_cap, (size_t)n)); io->chg_n = 0; }
  }
  CK(io->xyz3.ensure((size_t)3 * n, ctx->stream)); CK(io->chg_dev.ensure((size_t)n, ctx->stream));
  if (io->chg_dev_seen != io->chg_dev.p) io->chg_n = 0;   // the device array was re-allocated
  const int known = io->chg_n;                             // vnl_check sends natms records, the force call nlast: a prefix stays valid
  const int len = chunk_len(n), nch = cdiv(n, len);
  const double t0 = now_s();
  double* up = io->up;
  double* chg = io->chg;
  std::atomic<unsigned long long> changed{0};
  std::atomic<unsigned long long>* ch = &changed;
  io->pool->start(nch, [=](int c) {
    const int a = c * len, b = std::min(n, a + len);
    bool diff = b > known;
    int i = a;
#if defined(__x86_64__)
    // the staging buffer is only read by the DMA engine: non-temporal stores keep it out of the caches and save the read for
    // ownership; two records make three aligned 16-byte stores (chunks start at even records, 48 bytes per pair)
    for (; i + 1 < b; i += 2) {
      const dlpgpu_corepart& p0 = parts[i];
      const dlpgpu_corepart& p1 = parts[i + 1];
      double* q = up + 3 * (size_t)i;
      _mm_stream_pd(q, _mm_loadu_pd(&p0.xxx));
      _mm_stream_pd(q + 2, _mm_set_pd(p1.xxx, p0.zzz));
      _mm_stream_pd(q + 4, _mm_loadu_pd(&p1.yyy));
      if (std::memcmp(&chg[i], &p0.chge, sizeof(double)) != 0) { chg[i] = p0.chge; diff = true; }
      if (std::memcmp(&chg[i + 1], &p1.chge, sizeof(double)) != 0) { chg[i + 1] = p1.chge; diff = true; }
    }
#endif
    for (; i < b; ++i) {
      const dlpgpu_corepart& p = parts[i];
      double* q = up + 3 * (size_t)i;
      q[0] = p.xxx; q[1] = p.yyy; q[2] = p.zzz;
      if (std::memcmp(&chg[i], &p.chge, sizeof(double)) != 0) { chg[i] = p.chge; diff = true; }
    }
#if defined(__x86_64__)
    _mm_sfence();
#endif
    if (diff) ch->fetch_or(1ull << c, std::memory_order_relaxed);
  });
  cudaError_t e = cudaSuccess;
  unsigned long long moved = 0;
  for (int c = 0; c < nch; ++c) {
    io->pool->wait_chunk(c);
    const int a = c * len, b = std::min(n, a + len);
    if (e == cudaSuccess)
      e = cudaMemcpyAsync(io->xyz3.p + 3 * (size_t)a, up + 3 * (size_t)a, (size_t)(b - a) * 3 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
    moved += (unsigned long long)(b - a) * 3 * sizeof(double);
    if ((changed.load(std::memory_order_relaxed) >> c & 1ull) && e == cudaSuccess) {
      e = cudaMemcpyAsync(io->chg_dev.p + a, chg + a, (size_t)(b - a) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);
      moved += (unsigned long long)(b - a) * sizeof(double);
    }
  }
  io->pool->wait_all();
  CK(e);
  LAUNCH(ctx, k_merge_xyzq, cdiv(n, 256), 256, 0, n, io->xyz3.p, io->chg_dev.p, ctx->posq.p);
  io->chg_n = std::max(known, n); io->chg_dev_seen = io->chg_dev.p;
  io->t_up += now_s() - t0; io->n_up++;
  io->h2d += moved;
  ctx->tol_fresh = false; ctx->pub_fresh = false;
  ctx->parts_resident = 0; ctx->parts_current = false;
  return 0;
}

// ltype / ltg / lfrzn(1:n) of link_cell_pairs -> device arrays; chunks equal to what the device already holds are skipped
int dlp_upload_ints(dlpgpu_ctx* ctx, int n, const int* ltype, const int* ltg, const int* lfrzn) {
  if (n <= 0) return 0;
  HostIO* io = io_of(ctx);
  const bool fresh = n != io->ints_n || (lfrzn != nullptr) != io->frzn_given || (size_t)3 * n > io->ints_cap;
  if ((size_t)3 * n > io->ints_cap) {
    CK(cudaStreamSynchronize(ctx->stream));
    CK(grow_pinned(io->ints, io->ints_cap, (size_t)3 * n));
  }
  const int len = chunk_len(n), nch = cdiv(n, len);
  int* sh = io->ints;
  const int* src[3] = {ltype, ltg, lfrzn};
  int* dst[3] = {ctx->ltype.p, ctx->ltg.p, ctx->lfrzn.p};
  static_assert(MAXCH <= 64, "one bit per chunk");
  std::atomic<unsigned long long> changed[3];
  for (auto& w : changed) w.store(0);
  std::atomic<unsigned long long>* ch = changed;
  io->pool->start(nch, [=](int c) {
    const int a = c * len, b = std::min(n, a + len);
    const size_t bytes = (size_t)(b - a) * sizeof(int);
    for (int k = 0; k < 3; ++k) {
      if (!src[k]) continue;
      int* s = sh + (size_t)k * n + a;
      if (fresh || std::memcmp(s, src[k] + a, bytes) != 0) {
        std::memcpy(s, src[k] + a, bytes);
        ch[k].fetch_or(1ull << c, std::memory_order_relaxed);
      }
    }
  });
  cudaError_t e = cudaSuccess;
  unsigned long long moved = 0;
  for (int c = 0; c < nch; ++c) {
    io->pool->wait_chunk(c);
    const int a = c * len, b = std::min(n, a + len);
    const size_t bytes = (size_t)(b - a) * sizeof(int);
    for (int k = 0; k < 3; ++k)
      if (src[k] && (changed[k].load(std::memory_order_relaxed) >> c & 1ull) && e == cudaSuccess) {
        e = cudaMemcpyAsync(dst[k] + a, sh + (size_t)k * n + a, bytes, cudaMemcpyHostToDevice, ctx->stream);
        moved += bytes;
      }
  }
  io->pool->wait_all();
  CK(e);
  if (!lfrzn && (fresh || io->frzn_given)) CK(cudaMemsetAsync(ctx->lfrzn.p, 0, (size_t)n * sizeof(int), ctx->stream));
  io->ints_n = n; io->frzn_given = lfrzn != nullptr;
  io->h2d += moved;
  return 0;
}

// the device copies of ltype / ltg / lfrzn were overwritten by somebody else (native driver, SPME drop-in ...)
void dlp_hostio_ints_stale(dlpgpu_ctx* ctx) {
  if (ctx->hostio) static_cast<HostIO*>(ctx->hostio)->ints_n = -1;
}

// parts(1:natms)%f += the device's force arrays; returns when the caller's records are complete.  Everything the stream
// holds in front of the copies (the force kernels) overlaps with nothing here, but copy and host-side addition pipeline.
int dlp_download_add_forces(dlpgpu_ctx* ctx, int natms, dlpgpu_corepart* parts) {
  if (natms <= 0) return 0;
  HostIO* io = io_of(ctx);
  cudaStream_t s = ctx->stream;
  if (io->threads == 0) {   // whole records: the device copy still holds the caller's forces, the sum is formed there
    if (io->parts_dev.cap < (size_t)natms) return dlp_fail(ctx, DLPGPU_ERR_STATE, "drop-in forces: the records were never uploaded in whole-record mode");
    LAUNCH(ctx, k_add_forces, cdiv(natms, 256), 256, 0, io->parts_dev.p, natms, ctx->fx.p, ctx->fy.p, ctx->fz.p);
    CK(cudaMemcpyAsync(parts, io->parts_dev.p, (size_t)natms * sizeof(dlpgpu_corepart), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    io->d2h += (unsigned long long)natms * sizeof(dlpgpu_corepart);
    return 0;
  }
  if ((size_t)3 * natms > io->dn_cap) {
    CK(cudaStreamSynchronize(s));
    CK(grow_pinned(io->dn, io->dn_cap, (size_t)3 * natms));
  }
  CK(io->f3.ensure((size_t)3 * natms, s));
  const int len = chunk_len(natms), nch = cdiv(natms, len);
  for (int c = 0; c < nch; ++c) {
    if (!io->ev[c]) CK(cudaEventCreateWithFlags(&io->ev[c], cudaEventDisableTiming));
    io->avail[c].store(0, std::memory_order_relaxed);
  }
  LAUNCH(ctx, k_pack_f3, cdiv(natms, 256), 256, 0, natms, ctx->fx.p, ctx->fy.p, ctx->fz.p, io->f3.p);
  for (int c = 0; c < nch; ++c) {
    const int a = c * len, b = std::min(natms, a + len);
    CK(cudaMemcpyAsync(io->dn + 3 * (size_t)a, io->f3.p + 3 * (size_t)a, (size_t)(b - a) * 3 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaEventRecord(io->ev[c], s));
  }
  const double* dn = io->dn;
  std::atomic<int>* avail = io->avail;
  io->pool->start(nch, [=](int c) {
    unsigned spins = 0;
    int v;
    while ((v = avail[c].load(std::memory_order_acquire)) == 0) HostPool::pause(++spins);
    if (v < 0) return;   // the copy failed: leave the records alone
    const int a = c * len, b = std::min(natms, a + len);
    for (int i = a; i < b; ++i) {
      dlpgpu_corepart& p = parts[i];
      const double* f = dn + 3 * (size_t)i;
      p.fxx += f[0]; p.fyy += f[1]; p.fzz += f[2];
    }
  });
  cudaError_t e = cudaSuccess;
  const double t0 = now_s();
  double t1 = t0;
  for (int c = 0; c < nch; ++c) {
    if (e == cudaSuccess) e = cudaEventSynchronize(io->ev[c]);
    if (c == 0) t1 = now_s();
    io->avail[c].store(e == cudaSuccess ? 1 : -1, std::memory_order_release);
  }
  io->pool->wait_all();
  CK(e);
  io->t_wait += t1 - t0; io->t_down += now_s() - t1; io->n_down++;
  io->d2h += (unsigned long long)natms * 3 * sizeof(double);
  return 0;
}

extern "C" {

int dlpgpu_set_host_threads(dlpgpu_ctx* ctx, int nthreads) {
  if (!ctx || nthreads < 0 || nthreads > 64) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->hostio) io_of(ctx);
  HostIO* io = static_cast<HostIO*>(ctx->hostio);
  if (io->threads != nthreads) {
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->parts_resident = 0; ctx->parts_current = false;   // the two modes keep the uploaded records in different places
  }
  io->threads = nthreads;
  io_of(ctx);
  return 0;
}

int dlpgpu_transfer_bytes(dlpgpu_ctx* ctx, unsigned long long* h2d, unsigned long long* d2h, int reset) {
  if (!ctx) return DLPGPU_ERR_ARG;
  HostIO* io = ctx->hostio ? static_cast<HostIO*>(ctx->hostio) : nullptr;
  if (h2d) *h2d = io ? io->h2d : 0;
  if (d2h) *d2h = io ? io->d2h : 0;
  if (io && reset) { io->h2d = 0; io->d2h = 0; io->t_up = io->t_wait = io->t_down = 0.0; io->n_up = io->n_down = 0; }
  return 0;
}

int dlpgpu_transfer_times(dlpgpu_ctx* ctx, double out[5]) {
  if (!ctx || !out) return DLPGPU_ERR_ARG;
  HostIO* io = ctx->hostio ? static_cast<HostIO*>(ctx->hostio) : nullptr;
  out[0] = io ? io->t_up : 0.0; out[1] = io ? io->t_wait : 0.0; out[2] = io ? io->t_down : 0.0;
  out[3] = io ? io->n_up : 0; out[4] = io ? io->n_down : 0;
  return 0;
}

}  // extern "C"

int dlp_preload_hostio() {
  cudaFuncAttributes a;
  const void* ks[] = {(const void*)k_pack_f3, (const void*)k_unpack_parts, (const void*)k_add_forces, (const void*)k_merge_xyzq};
  for (const void* k : ks) if (cudaFuncGetAttributes(&a, k) != cudaSuccess) cudaGetLastError();
  return 0;
}
