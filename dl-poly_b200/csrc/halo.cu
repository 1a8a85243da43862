// halo.cu -- padding-driven rebuild test, halo build / refresh and particle migration kernels.
//
//   neighbours.F90:123-296  vnl_check (max displacement, minimum image via numerics.F90:1511-1600 images)
//   halo.F90:153-355        set_halo_particles        deport_data.F90:1673-1951  export_atomic_data
//   halo.F90:47-113         refresh_halo_positions    deport_data.F90:2301-2553  export_atomic_positions
//   deport_data.F90:2870-3202 relocate_particles      deport_data.F90:81-960     deport_atomic_data
//   numerics.F90:1851-1950  pbcshift_parts
//
// The wire format is the reference's (x,y,z,ltg,lsite,ixyz per halo atom on a build -- plus three doubles naming the
// atom's origin for the one-kernel refresh below --, 3 doubles on a staged refresh, the sender applies the periodic shift); selection, ordering (ascending local index, appended in stage order -x,+x,-y,+y,-z,+z) and the
// restack of staying atoms after migration reproduce the reference so local indices agree with a DL_POLY run.
// Compiled with -fmad=false: thresholds and shifts decide set membership.
#include <unistd.h>

#include "common.cuh"

#include <atomic>
#include <chrono>

namespace {

struct Dir { int kx, ky, kz, jxyz, kxyz; double xadd, yadd, zadd; int lwrap; };

Dir dir_settings(const dlpgpu_ctx* c, int mdir) {   // deport_data.F90:1728-1796
  Dir s{0, 0, 0, 0, 0, 0, 0, 0, 0};
  bool lsx = false, lex = false, lsy = false, ley = false, lsz = false, lez = false;
  switch (mdir) {
    case -1: s.kx = 1; s.jxyz = 1; s.kxyz = 3; lsx = (c->idx == 0); break;
    case 1: s.kx = 1; s.jxyz = 2; s.kxyz = 3; lex = (c->idx == c->nx - 1); break;
    case -2: s.ky = 1; s.jxyz = 10; s.kxyz = 30; lsy = (c->idy == 0); break;
    case 2: s.ky = 1; s.jxyz = 20; s.kxyz = 30; ley = (c->idy == c->ny - 1); break;
    case -3: s.kz = 1; s.jxyz = 100; s.kxyz = 300; lsz = (c->idz == 0); break;
    case 3: s.kz = 1; s.jxyz = 200; s.kxyz = 300; lez = (c->idz == c->nz - 1); break;
  }
  double uuu = 0.0; if (lsx) uuu = +1.0; if (lex) uuu = -1.0;
  double vvv = 0.0; if (lsy) vvv = +1.0; if (ley) vvv = -1.0;
  double www = 0.0; if (lsz) www = +1.0; if (lez) www = -1.0;
  s.lwrap = (std::fabs(uuu) + std::fabs(vvv) + std::fabs(www) > 0.5) ? 1 : 0;
  if (s.lwrap) {
    const double* cell = c->cell - 1;
    s.xadd = cell[1] * uuu + cell[4] * vvv + cell[7] * www;
    s.yadd = cell[2] * uuu + cell[5] * vvv + cell[8] * www;
    s.zadd = cell[3] * uuu + cell[6] * vvv + cell[9] * www;
  }
  return s;
}
int stage_of(int mdir) { return mdir == -1 ? 0 : mdir == 1 ? 1 : mdir == -2 ? 2 : mdir == 2 ? 3 : mdir == -3 ? 4 : mdir == 3 ? 5 : -1; }

void h_invert(const double* a0, double* b0) {
  const double* a = a0 - 1;
  double* b = b0 - 1;
  b[1] = a[5] * a[9] - a[6] * a[8]; b[2] = a[3] * a[8] - a[2] * a[9]; b[3] = a[2] * a[6] - a[3] * a[5];
  b[4] = a[6] * a[7] - a[4] * a[9]; b[5] = a[1] * a[9] - a[3] * a[7]; b[6] = a[3] * a[4] - a[1] * a[6];
  b[7] = a[4] * a[8] - a[5] * a[7]; b[8] = a[2] * a[7] - a[1] * a[8]; b[9] = a[1] * a[5] - a[2] * a[4];
  double d = a[1] * b[1] + a[4] * b[2] + a[7] * b[3];
  double r = 0.0;
  if (std::fabs(d) > 0.0) r = 1.0 / d;
  for (int i = 1; i <= 9; ++i) b[i] = r * b[i];
}
void h_widths(const double* aaa0, double* w3) {   // dcell bbb(7:9), numerics.F90:1344-1446
  const double* aaa = aaa0 - 1;
  double b1 = std::sqrt(aaa[1] * aaa[1] + aaa[2] * aaa[2] + aaa[3] * aaa[3]);
  double b2 = std::sqrt(aaa[4] * aaa[4] + aaa[5] * aaa[5] + aaa[6] * aaa[6]);
  double b3 = std::sqrt(aaa[7] * aaa[7] + aaa[8] * aaa[8] + aaa[9] * aaa[9]);
  double axb1 = aaa[2] * aaa[6] - aaa[3] * aaa[5], axb2 = aaa[3] * aaa[4] - aaa[1] * aaa[6], axb3 = aaa[1] * aaa[5] - aaa[2] * aaa[4];
  double bxc1 = aaa[5] * aaa[9] - aaa[6] * aaa[8], bxc2 = aaa[6] * aaa[7] - aaa[4] * aaa[9], bxc3 = aaa[4] * aaa[8] - aaa[5] * aaa[7];
  double cxa1 = aaa[8] * aaa[3] - aaa[9] * aaa[2], cxa2 = aaa[9] * aaa[1] - aaa[7] * aaa[3], cxa3 = aaa[7] * aaa[2] - aaa[8] * aaa[1];
  double vol = std::fabs(aaa[1] * bxc1 + aaa[2] * bxc2 + aaa[3] * bxc3);
  double d[4], x[4], y[4];
  d[1] = vol / std::sqrt(bxc1 * bxc1 + bxc2 * bxc2 + bxc3 * bxc3);
  d[2] = vol / std::sqrt(cxa1 * cxa1 + cxa2 * cxa2 + cxa3 * cxa3);
  d[3] = vol / std::sqrt(axb1 * axb1 + axb2 * axb2 + axb3 * axb3);
  x[1] = std::fabs(aaa[1]) / b1; y[1] = std::fabs(aaa[2]) / b1;
  x[2] = std::fabs(aaa[4]) / b2; y[2] = std::fabs(aaa[5]) / b2;
  x[3] = std::fabs(aaa[7]) / b3; y[3] = std::fabs(aaa[8]) / b3;
  if (x[1] >= x[2] && x[1] >= x[3]) { w3[0] = d[1]; if (y[2] >= y[3]) { w3[1] = d[2]; w3[2] = d[3]; } else { w3[1] = d[3]; w3[2] = d[2]; } }
  else if (x[2] >= x[1] && x[2] >= x[3]) { w3[0] = d[2]; if (y[1] >= y[3]) { w3[1] = d[1]; w3[2] = d[3]; } else { w3[1] = d[3]; w3[2] = d[1]; } }
  else { w3[0] = d[3]; if (y[1] >= y[2]) { w3[1] = d[1]; w3[2] = d[2]; } else { w3[1] = d[2]; w3[2] = d[1]; } }
}

struct Mat9 { double m[9]; };

// ---------------------------------------------------------------- vnl_check
__device__ __forceinline__ double vnl_displacement(int imcon, const Mat9& cell, const Mat9& rcell, double x, double y, double z) {
  {
    if (imcon == 1) {                                              // numerics.F90:1553-1563
      double aaa = 1.0 / cell.m[0];
      x = x - cell.m[0] * round(aaa * x) ; y = y - cell.m[0] * round(aaa * y); z = z - cell.m[0] * round(aaa * z);
    } else if (imcon == 2 || imcon == 0) {                         // :1563 (IMCON_ORTHORHOMBIC .or. IMCON_NOPBC: the reference folds both)
      double aaa = 1.0 / cell.m[0], bbb = 1.0 / cell.m[4], ccc = 1.0 / cell.m[8];
      x = x - cell.m[0] * round(aaa * x); y = y - cell.m[4] * round(bbb * y); z = z - cell.m[8] * round(ccc * z);
    } else if (imcon == 3) {
      double xss = rcell.m[0] * x + rcell.m[3] * y + rcell.m[6] * z;
      double yss = rcell.m[1] * x + rcell.m[4] * y + rcell.m[7] * z;
      double zss = rcell.m[2] * x + rcell.m[5] * y + rcell.m[8] * z;
      xss = xss - round(xss); yss = yss - round(yss); zss = zss - round(zss);
      x = cell.m[0] * xss + cell.m[3] * yss + cell.m[6] * zss;
      y = cell.m[1] * xss + cell.m[4] * yss + cell.m[7] * zss;
      z = cell.m[2] * xss + cell.m[5] * yss + cell.m[8] * zss;
    }
    return sqrt(x * x + y * y + z * z);                            // :166
  }
}
// maximum of r >= 0 over the block into *tol_bits (the bits of a non-negative double are order-preserving): one atomic per block --
// one per warp put 31 k atomics of a 1 M-atom pass on a single L2 address, which serialises them
__device__ __forceinline__ void block_max_to(double r, unsigned long long* tol_bits) {
  __shared__ double s_max[32];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) r = fmax(r, __shfl_xor_sync(DLP_FULL, r, d));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) s_max[w] = r;
  __syncthreads();
  if (w == 0) {
    r = lane < nw ? s_max[lane] : 0.0;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) r = fmax(r, __shfl_xor_sync(DLP_FULL, r, d));
    if (lane == 0 && r > 0.0) atomicMax(tol_bits, (unsigned long long)__double_as_longlong(r));
  }
}
__global__ void k_vnl_tol(int natms, int imcon, Mat9 cell, Mat9 rcell, const double4* __restrict__ posq, const double* __restrict__ xbg,
                          const double* __restrict__ ybg, const double* __restrict__ zbg, unsigned long long* __restrict__ tol_bits) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double r = 0.0;
  if (i < natms) {
    double4 p = posq[i];
    r = vnl_displacement(imcon, cell, rcell, p.x - xbg[i], p.y - ybg[i], p.z - zbg[i]);   // neighbours.F90:157-161
  }
  block_max_to(r, tol_bits);
}
// velocity-Verlet stage 1 (nve.F90:163-173, same contraction as forces.cu::k_vv) fused with what always follows it in the
// native driver: the displacement test of vnl_check on the new positions and the copy into the peer-visible buffer of the
// one-kernel halo refresh -- one pass over the atoms instead of three
__global__ void k_vv1_fused(int natms, double dt, int imcon, Mat9 cell, Mat9 rcell, const int* __restrict__ lsite,
                            const double* __restrict__ weight_site, double4* __restrict__ posq, double* vx, double* vy, double* vz,
                            const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz,
                            const double* __restrict__ xbg, const double* __restrict__ ybg, const double* __restrict__ zbg,
                            unsigned long long* __restrict__ tol_bits, double4* __restrict__ pub) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  double r = 0.0;
  if (i < natms) {
    const double hstep = 0.5 * dt;
    const double rm = 1.0 / weight_site[lsite[i] - 1];
    const double tmp = hstep * rm;
    const double a = __fma_rn(tmp, fx[i], vx[i]), b = __fma_rn(tmp, fy[i], vy[i]), c = __fma_rn(tmp, fz[i], vz[i]);
    vx[i] = a; vy[i] = b; vz[i] = c;
    double4 p = posq[i];
    p.x = __fma_rn(dt, a, p.x); p.y = __fma_rn(dt, b, p.y); p.z = __fma_rn(dt, c, p.z);
    posq[i] = p;
    if (pub) pub[i] = p;
    if (tol_bits) r = vnl_displacement(imcon, cell, rcell, p.x - xbg[i], p.y - ybg[i], p.z - zbg[i]);
  }
  if (tol_bits) block_max_to(r, tol_bits);
}

// ---------------------------------------------------------------- halo build
struct HaloThr { double ecwx, ecwy, ecwz, cwx, cwy, cwz; };

__global__ void k_halo_tag(int natms, Mat9 rcell, HaloThr t, const double4* __restrict__ posq, int* __restrict__ ixyz, int my_rank,
                           int* __restrict__ org_rank, int* __restrict__ org_idx, int* __restrict__ org_wrap) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natms) return;
  org_rank[i] = my_rank; org_idx[i] = i; org_wrap[i] = 13;   // 13 = no wrap on any axis
  double4 p = posq[i];
  double x = rcell.m[0] * p.x + rcell.m[3] * p.y + rcell.m[6] * p.z;   // halo.F90:265-267
  double y = rcell.m[1] * p.x + rcell.m[4] * p.y + rcell.m[7] * p.z;
  double z = rcell.m[2] * p.x + rcell.m[5] * p.y + rcell.m[8] * p.z;
  int v = 0;
  if (x <= t.ecwx) v += 1;
  if (x >= t.cwx) v += 2;
  if (y <= t.ecwy) v += 10;
  if (y >= t.cwy) v += 20;
  if (z <= t.ecwz) v += 100;
  if (z >= t.cwz) v += 200;
  ixyz[i] = v;
}

__device__ __forceinline__ int halo_sel(int v, Dir d) {   // deport_data.F90:1814-1828 -> 0: not selected, 1: j==jxyz, 2: both sides
  if (v <= 0) return 0;
  int ix = v % 10, iy = (v - ix) % 100, iz = (v - (ix + iy)) % 1000;
  int j = ix * d.kx + iy * d.ky + iz * d.kz;
  if (j == d.jxyz) return 1;
  if (j > d.jxyz && j % 3 == 0) return 2;
  return 0;
}
__global__ void k_halo_flag(int n, Dir d, const int* __restrict__ ixyz, int* __restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  flag[i] = (i < n && halo_sel(ixyz[i], d) != 0) ? 1 : 0;   // flag[n] = 0 pads the scan
}
__global__ void k_halo_pack(int n, Dir d, int cap, const int* __restrict__ flag, const int* __restrict__ pos, const double4* __restrict__ posq,
                            const int* __restrict__ ltg, const int* __restrict__ lsite, const int* __restrict__ ixyz,
                            const int* __restrict__ org_rank, const int* __restrict__ org_idx, const int* __restrict__ org_wrap, int wrap_add,
                            double* __restrict__ buf, int* __restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  int k = pos[i];
  idx[k] = i;
  if (k >= cap) return;
  double4 p = posq[i];
  double* b = buf + (size_t)k * DLP_HALO_W;
  b[6] = (double)org_rank[i]; b[7] = (double)org_idx[i]; b[8] = (double)(org_wrap[i] + wrap_add);
  if (!d.lwrap) { b[0] = p.x; b[1] = p.y; b[2] = p.z; }
  else { b[0] = p.x + d.xadd; b[1] = p.y + d.yadd; b[2] = p.z + d.zadd; }   // :1836-1844
  b[3] = (double)ltg[i];
  b[4] = (double)lsite[i];
  int v = ixyz[i];
  b[5] = (double)(v - (halo_sel(v, d) == 1 ? d.jxyz : d.kxyz));            // :1853
}
__global__ void k_halo_unpack(int count, int off, const double* __restrict__ buf, double4* __restrict__ posq, int* __restrict__ ltg,
                              int* __restrict__ lsite, int* __restrict__ ixyz, double* fx, double* fy, double* fz,
                              int* __restrict__ org_rank, int* __restrict__ org_idx, int* __restrict__ org_wrap) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const double* b = buf + (size_t)k * DLP_HALO_W;
  int i = off + k;
  org_rank[i] = __double2int_rn(b[6]); org_idx[i] = __double2int_rn(b[7]); org_wrap[i] = __double2int_rn(b[8]);
  posq[i] = make_double4(b[0], b[1], b[2], 0.0);
  ltg[i] = __double2int_rn(b[3]); lsite[i] = __double2int_rn(b[4]); ixyz[i] = __double2int_rn(b[5]);   // Nint, :1935-1940
  fx[i] = 0.0; fy[i] = 0.0; fz[i] = 0.0;
}
__global__ void k_assign_sites(int from, int to, const int* __restrict__ lsite, const int* __restrict__ type_site,
                               const double* __restrict__ charge_site, const int* __restrict__ freeze_site, double4* __restrict__ posq,
                               int* __restrict__ ltype, int* __restrict__ lfrzn) {
  int i = from + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= to) return;
  int s = lsite[i] - 1;
  ltype[i] = type_site[s];                      // halo.F90:296-302 / deport_data.F90:3062-3069
  posq[i].w = charge_site[s];
  lfrzn[i] = freeze_site[s];
}

// ---------------------------------------------------------------- halo refresh
__global__ void k_refresh_pack(int count, Dir d, const int* __restrict__ idx, const double4* __restrict__ posq, double* __restrict__ buf) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  double4 p = posq[idx[k]];
  double* b = buf + (size_t)k * 3;
  if (!d.lwrap) { b[0] = p.x; b[1] = p.y; b[2] = p.z; }
  else { b[0] = p.x + d.xadd; b[1] = p.y + d.yadd; b[2] = p.z + d.zadd; }   // deport_data.F90:2460-2468
}
__global__ void k_refresh_unpack(int count, int off, const double* __restrict__ buf, double4* __restrict__ posq) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const double* b = buf + (size_t)k * 3;
  double4 p = posq[off + k];
  p.x = b[0]; p.y = b[1]; p.z = b[2];
  posq[off + k] = p;
}

// ---------------------------------------------------------------- one-kernel halo refresh over peer memory
// refresh_halo_positions re-sends the same atoms through the same six dependent stages.  Every halo atom is, in the end, a
// copy of ONE local atom of some rank (possibly this one) plus the periodic shifts picked up on the way; the halo build
// records that origin.  Each rank publishes its local coordinates in a CUDA-IPC mapped buffer (double-buffered by step
// parity) and every rank fills its whole halo with one kernel of peer loads over NVLink -- no staging, no messages.
// The shifts are replayed in stage order (x, then y, then z) with the arithmetic of the staged exchange, so the
// coordinates carry the same bits.
__global__ void k_publish(int natms, const double4* __restrict__ posq, double4* __restrict__ pub) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < natms) pub[i] = posq[i];
}
__global__ void k_refresh_pull(int natms, int nlast, int parity, const unsigned long long* __restrict__ peers, Mat9 cell,
                               const int* __restrict__ org_rank, const int* __restrict__ org_idx, const int* __restrict__ org_wrap,
                               double4* __restrict__ posq) {
  int h = natms + blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= nlast) return;
  const double4* src = reinterpret_cast<const double4*>(peers[2 * org_rank[h] + parity]) + org_idx[h];
  const double2 a = __ldcg(reinterpret_cast<const double2*>(src));
  const double c = __ldcg(reinterpret_cast<const double*>(src) + 2);
  double x = a.x, y = a.y, z = c;
  const int w = org_wrap[h];
  const double u = (double)(w % 3 - 1), v = (double)((w / 3) % 3 - 1), ww = (double)(w / 9 - 1);
  // deport_data.F90:1786-1796 / :2460-2468 per stage: xadd = cell(1) uuu + cell(4) vvv + cell(7) www with one of them set
  if (u != 0.0) { x = x + (cell.m[0] * u + cell.m[3] * 0.0 + cell.m[6] * 0.0); y = y + (cell.m[1] * u + cell.m[4] * 0.0 + cell.m[7] * 0.0); z = z + (cell.m[2] * u + cell.m[5] * 0.0 + cell.m[8] * 0.0); }
  if (v != 0.0) { x = x + (cell.m[0] * 0.0 + cell.m[3] * v + cell.m[6] * 0.0); y = y + (cell.m[1] * 0.0 + cell.m[4] * v + cell.m[7] * 0.0); z = z + (cell.m[2] * 0.0 + cell.m[5] * v + cell.m[8] * 0.0); }
  if (ww != 0.0) { x = x + (cell.m[0] * 0.0 + cell.m[3] * 0.0 + cell.m[6] * ww); y = y + (cell.m[1] * 0.0 + cell.m[4] * 0.0 + cell.m[7] * ww); z = z + (cell.m[2] * 0.0 + cell.m[5] * 0.0 + cell.m[8] * ww); }
  double4 p = posq[h];
  p.x = x; p.y = y; p.z = z;
  posq[h] = p;
}

// ---------------------------------------------------------------- relocation
__global__ void k_pbcshift(int natms, int imcon, Mat9 cell, Mat9 rcell, double4* __restrict__ posq) {   // numerics.F90:1889-1950
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natms) return;
  const double half_minus = 0.499999999999999944488848768742172978818416595458984375;
  double4 p = posq[i];
  double xss, yss, zss;
  // Anint: round half away from zero
  if (imcon == 1) {
    double aaa = 1.0 / cell.m[0];
    xss = aaa * p.x; yss = aaa * p.y; zss = aaa * p.z;
  } else if (imcon == 2 || imcon == 0) {
    xss = (1.0 / cell.m[0]) * p.x; yss = (1.0 / cell.m[4]) * p.y; zss = (1.0 / cell.m[8]) * p.z;
  } else {
    xss = rcell.m[0] * p.x + rcell.m[3] * p.y + rcell.m[6] * p.z;
    yss = rcell.m[1] * p.x + rcell.m[4] * p.y + rcell.m[7] * p.z;
    zss = rcell.m[2] * p.x + rcell.m[5] * p.y + rcell.m[8] * p.z;
  }
  xss = xss - round(xss); if (xss >= half_minus) xss = -xss;
  yss = yss - round(yss); if (yss >= half_minus) yss = -yss;
  zss = zss - round(zss); if (zss >= half_minus) zss = -zss;
  if (imcon == 1) { p.x = cell.m[0] * xss; p.y = cell.m[0] * yss; p.z = cell.m[0] * zss; }
  else if (imcon == 2 || imcon == 0) { p.x = cell.m[0] * xss; p.y = cell.m[4] * yss; p.z = cell.m[8] * zss; }
  else {
    p.x = cell.m[0] * xss + cell.m[3] * yss + cell.m[6] * zss;
    p.y = cell.m[1] * xss + cell.m[4] * yss + cell.m[7] * zss;
    p.z = cell.m[2] * xss + cell.m[5] * yss + cell.m[8] * zss;
  }
  posq[i] = p;
}

struct DomI { int nx, ny, nz, idx, idy, idz; };
__global__ void k_reloc_tag(int natms, Mat9 rcell, DomI D, const double4* __restrict__ posq, int* __restrict__ ixyz) {   // deport_data.F90:2981-3025
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natms) return;
  const double half_plus = 0.50000000000000011102230246251565404236316680908203125;
  const double half_minus = 0.499999999999999944488848768742172978818416595458984375;
  double4 p = posq[i];
  double x = rcell.m[0] * p.x + rcell.m[3] * p.y + rcell.m[6] * p.z;
  double y = rcell.m[1] * p.x + rcell.m[4] * p.y + rcell.m[7] * p.z;
  double z = rcell.m[2] * p.x + rcell.m[5] * p.y + rcell.m[8] * p.z;
  int ipx = __double2int_rz((x + 0.5) * (double)D.nx), ipy = __double2int_rz((y + 0.5) * (double)D.ny), ipz = __double2int_rz((z + 0.5) * (double)D.nz);
  int v = 0;
  if (D.idx == 0) { if (x < -half_plus) v += 1; } else { if (ipx < D.idx) v += 1; }
  if (D.idx == D.nx - 1) { if (x >= half_minus) v += 2; } else { if (ipx > D.idx) v += 2; }
  if (D.idy == 0) { if (y < -half_plus) v += 10; } else { if (ipy < D.idy) v += 10; }
  if (D.idy == D.ny - 1) { if (y >= half_minus) v += 20; } else { if (ipy > D.idy) v += 20; }
  if (D.idz == 0) { if (z < -half_plus) v += 100; } else { if (ipz < D.idz) v += 100; }
  if (D.idz == D.nz - 1) { if (z >= half_minus) v += 200; } else { if (ipz > D.idz) v += 200; }
  ixyz[i] = v;
}
__global__ void k_reloc_flag(int n, Dir d, const int* __restrict__ ixyz, int* __restrict__ leave) {   // deport_data.F90:254-274
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  int f = 0;
  if (i < n) {
    int v = ixyz[i];
    if (v != 0) {
      int ix = v % 10, iy = (v - ix) % 100, iz = (v - (ix + iy)) % 1000;
      int j = ix * d.kx + iy * d.ky + iz * d.kz;
      if (j == d.jxyz) f = 1;   // the tag is decremented when the atom is packed (a too-small buffer retries the stage)
    }
  }
  leave[i] = f;
}
__global__ void k_reloc_pack(int n, int k_stay, Dir d, int cap, const int* __restrict__ leave, const int* __restrict__ lpos,
                             const double4* __restrict__ posq, const double* __restrict__ vx, const double* __restrict__ vy,
                             const double* __restrict__ vz, const double* __restrict__ fx, const double* __restrict__ fy,
                             const double* __restrict__ fz, const int* __restrict__ ltg, const int* __restrict__ lsite,
                             const int* __restrict__ ixyz, double* __restrict__ buf, int* __restrict__ hole_pos) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !leave[i]) return;
  int k = lpos[i];                       // rank among leavers, ascending index == the reference's buffer order
  if (i < k_stay) hole_pos[k] = i;       // holes below the new natms are the first ones (ascending)
  if (k >= cap) return;
  double4 p = posq[i];
  double* b = buf + (size_t)k * 12;
  if (!d.lwrap) { b[0] = p.x; b[1] = p.y; b[2] = p.z; }
  else { b[0] = p.x + d.xadd; b[1] = p.y + d.yadd; b[2] = p.z + d.zadd; }   // deport_data.F90:296-305
  b[3] = vx[i]; b[4] = vy[i]; b[5] = vz[i];
  b[6] = fx[i]; b[7] = fy[i]; b[8] = fz[i];
  b[9] = (double)ltg[i]; b[10] = (double)lsite[i]; b[11] = (double)(ixyz[i] - d.jxyz);
}
// restack (deport_data.F90:822-925): the r-th hole (ascending) below the new natms takes the r-th staying atom counted
// from the end.  stay_prefix[j] = j - lpos[j] staying atoms precede j.
__global__ void k_reloc_restack(int n, int k_stay, const int* __restrict__ leave, const int* __restrict__ lpos, const int* __restrict__ hole_pos,
                                double4* __restrict__ posq, double* vx, double* vy, double* vz, double* fx, double* fy, double* fz,
                                int* ltg, int* lsite, int* ixyz) {
  int j = k_stay + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || leave[j]) return;
  int stay_prefix = j - lpos[j];
  int r = k_stay - 1 - stay_prefix;
  int dst = hole_pos[r];
  posq[dst] = posq[j];
  vx[dst] = vx[j]; vy[dst] = vy[j]; vz[dst] = vz[j];
  fx[dst] = fx[j]; fy[dst] = fy[j]; fz[dst] = fz[j];
  ltg[dst] = ltg[j]; lsite[dst] = lsite[j]; ixyz[dst] = ixyz[j];
}
__global__ void k_reloc_unpack(int count, int off, const double* __restrict__ buf, double4* __restrict__ posq, double* vx, double* vy,
                               double* vz, double* fx, double* fy, double* fz, int* ltg, int* lsite, int* ixyz) {
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const double* b = buf + (size_t)k * 12;
  int i = off + k;
  posq[i] = make_double4(b[0], b[1], b[2], 0.0);
  vx[i] = b[3]; vy[i] = b[4]; vz[i] = b[5];
  fx[i] = b[6]; fy[i] = b[7]; fz[i] = b[8];
  ltg[i] = __double2int_rn(b[9]); lsite[i] = __double2int_rn(b[10]); ixyz[i] = __double2int_rn(b[11]);
}
__global__ void k_count_nonzero(int n, const int* __restrict__ ixyz, int* __restrict__ status) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && ixyz[i] != 0) atomicAdd(&status[2], 1);
}

Mat9 mat(const double* a) { Mat9 m; for (int i = 0; i < 9; ++i) m.m[i] = a[i]; return m; }

// ================================================================ fused device-side exchange over peer memory
// relocate_particles + set_halo_particles of one rebuild -- twelve dependent stages -- and the gmax of vnl_check without a
// single host round trip between the stages and without NCCL: every rank owns a CUDA-IPC exported region holding, per
// stage, a receive buffer of fixed capacity and a header {sequence number, atom count}.  The sender's pack kernel writes
// the payload straight into the RECEIVER's buffer over NVLink (or into its own when the decomposition has one domain in
// that direction, deport_data.F90:1884-1886), a one-thread kernel publishes count + sequence with a system-scope release,
// and the receiver's unpack kernel spins on its own header with acquire loads before it appends the atoms.  natms / nlast
// live on the device while the twelve stages run (kernels are launched over host-side upper bounds and read the live
// counts), the host learns them with ONE synchronisation at the end.  Selection rules, payloads, ordering and the periodic
// shifts are those of the staged routines above (same device functions), so local indices and bits agree with them.
struct XHdr { unsigned long long seq; long long count; };

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// true when the header carries `seq`; gives up after g_wait_ns nanoseconds (default 60 s, dlpgpu_dev_xchg_set_timeout) and raises
// DC_ERR bit 2 (a peer died / lost lock-step) -- long enough for a peer that is still in a first-touch list build
__device__ unsigned long long g_wait_ns = 60ull * 1000000000ull;
__device__ __forceinline__ unsigned long long x_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ bool x_wait(const unsigned long long* seq_word, unsigned long long seq, int* err) {
  const unsigned long long t0 = x_now_ns(), limit = g_wait_ns;
  for (long long it = 0;; ++it) {
    if (ld_acquire_sys(seq_word) == seq) return true;
    if ((it & 1023) == 1023) {
      if (*(volatile int*)err & 4) return false;   // somebody already gave up: do not wait again
      if (x_now_ns() - t0 > limit) break;
    }
    __nanosleep(200);
  }
  atomicOr(err, 4);
  return false;
}
// device counts block
#define DC_NATMS 0
#define DC_NLAST 1
#define DC_ERR 2      // bit 0: migration buffer overflow (error 43), 1: halo buffer overflow (54), 2: wait timeout, 3: atom arrays full
#define DC_LOST 3
#define DC_HSENT 8
#define DC_HRECV 16
#define DC_HOFF 24
#define DC_RSENT 32
#define DC_RRECV 40
#define DC_WORDS 64

__global__ void k_x_reloc_tag(const int* __restrict__ dc, Mat9 rcell, DomI D, const double4* __restrict__ posq, int* __restrict__ ixyz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dc[DC_NATMS]) return;
  const double half_plus = 0.50000000000000011102230246251565404236316680908203125;
  const double half_minus = 0.499999999999999944488848768742172978818416595458984375;
  double4 p = posq[i];
  double x = rcell.m[0] * p.x + rcell.m[3] * p.y + rcell.m[6] * p.z;
  double y = rcell.m[1] * p.x + rcell.m[4] * p.y + rcell.m[7] * p.z;
  double z = rcell.m[2] * p.x + rcell.m[5] * p.y + rcell.m[8] * p.z;
  int ipx = __double2int_rz((x + 0.5) * (double)D.nx), ipy = __double2int_rz((y + 0.5) * (double)D.ny), ipz = __double2int_rz((z + 0.5) * (double)D.nz);
  int v = 0;   // deport_data.F90:2981-3025
  if (D.idx == 0) { if (x < -half_plus) v += 1; } else { if (ipx < D.idx) v += 1; }
  if (D.idx == D.nx - 1) { if (x >= half_minus) v += 2; } else { if (ipx > D.idx) v += 2; }
  if (D.idy == 0) { if (y < -half_plus) v += 10; } else { if (ipy < D.idy) v += 10; }
  if (D.idy == D.ny - 1) { if (y >= half_minus) v += 20; } else { if (ipy > D.idy) v += 20; }
  if (D.idz == 0) { if (z < -half_plus) v += 100; } else { if (ipz < D.idz) v += 100; }
  if (D.idz == D.nz - 1) { if (z >= half_minus) v += 200; } else { if (ipz > D.idz) v += 200; }
  ixyz[i] = v;
}
// ---- stage kernels: count -> pack (+ signal) -> [restack] -> receive (+ bookkeeping): 3 launches per halo stage, 4 per
// migration stage.  A block owns XB_N consecutive atoms, eight per thread, so packed order = ascending local index (the
// reference's buffer order).  The last block to finish (a ticket in the counts block) does the one-thread work.
#define XB_T 256
#define XB_A 4   // 16 atoms per thread (a quarter of the blocks and of the ticket chain) is slower: 0.36 against 0.25 ms per 1 M-ion exchange
#define XB_N (XB_T * XB_A)
#define DC_TICKET 4
#define DC_NOLD 5
#define DC_TOT 6

__device__ __forceinline__ int reloc_sel(int v, const Dir& d) {   // deport_data.F90:254-274
  if (v == 0) return 0;
  int ix = v % 10, iy = (v - ix) % 100, iz = (v - (ix + iy)) % 1000;
  return (ix * d.kx + iy * d.ky + iz * d.kz == d.jxyz) ? 1 : 0;
}
__device__ __forceinline__ int x_block_sum(int v, int* s_w) {   // sum over the block, returned to every thread
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(DLP_FULL, v, o);
  __syncthreads();
  if (lane == 0) s_w[w] = v;
  __syncthreads();
  int t = 0;
#pragma unroll
  for (int k = 0; k < XB_T / 32; ++k) t += s_w[k];
  return t;
}
template <int HALO>
__global__ void __launch_bounds__(XB_T) k_x_count(const int* __restrict__ dc, Dir d, const int* __restrict__ ixyz, int* __restrict__ blocksum) {
  __shared__ int s_w[XB_T / 32];
  const int n = HALO ? dc[DC_NLAST] : dc[DC_NATMS];
  const int base = blockIdx.x * XB_N + threadIdx.x;
  int c = 0;
#pragma unroll
  for (int j = 0; j < XB_A; ++j) {
    const int i = base + j * XB_T;
    if (i < n) c += HALO ? (halo_sel(ixyz[i], d) != 0) : reloc_sel(ixyz[i], d);
  }
  const int t = x_block_sum(c, s_w);
  if (threadIdx.x == 0) blocksum[blockIdx.x] = t;
}
struct XAtoms {   // the per-atom arrays a stage moves
  double4* posq; double *vx, *vy, *vz, *fx, *fy, *fz;
  int *ltg, *lsite, *ixyz, *org_rank, *org_idx, *org_wrap;
};
template <int HALO>
__global__ void __launch_bounds__(XB_T)
k_x_pack(int* __restrict__ dc, Dir d, int cap, int q, const int* __restrict__ blocksum, XAtoms A, int wrap_add, double* __restrict__ buf,
         int* __restrict__ idx_or_hole, int* __restrict__ leave, int* __restrict__ lpos, XHdr* dst_hdr, unsigned long long seq) {
  __shared__ int s_w[XB_T / 32], s_w2[XB_T / 32];
  __shared__ int s_last;
  const int n = HALO ? dc[DC_NLAST] : dc[DC_NATMS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // this thread's XB_A CONSECUTIVE atoms (one pass, one scan: the packed position of an atom = selected atoms before it --
  // blocks before, warps before, lanes before, the thread's own earlier atoms)
  const int i0 = blockIdx.x * XB_N + threadIdx.x * XB_A;
  int code[XB_A];
  int mine = 0;
#pragma unroll
  for (int j = 0; j < XB_A; ++j) {
    const int i = i0 + j;
    code[j] = 0;
    if (i < n) code[j] = HALO ? halo_sel(A.ixyz[i], d) : reloc_sel(A.ixyz[i], d);
    mine += code[j] != 0;
  }
  // atoms selected in the blocks before this one, and in all blocks (one pair of block reductions)
  int before = 0, all = 0;
  for (int k = threadIdx.x; k < gridDim.x; k += XB_T) { const int v = blocksum[k]; all += v; if (k < blockIdx.x) before += v; }
  int incl = mine;   // inclusive scan of `mine` over the warp
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(DLP_FULL, incl, o); if (lane >= o) incl += v; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { before += __shfl_xor_sync(DLP_FULL, before, o); all += __shfl_xor_sync(DLP_FULL, all, o); }
  if (lane == 31) s_w[w] = incl;                 // the warp's selected atoms
  if (lane == 0) s_w2[w] = before;
  __syncthreads();
  int wbase = 0, before_blocks = 0;
#pragma unroll
  for (int k = 0; k < XB_T / 32; ++k) { if (k < w) wbase += s_w[k]; before_blocks += s_w2[k]; }
  __syncthreads();
  if (lane == 0) s_w[w] = all;
  __syncthreads();
  int total = 0;
#pragma unroll
  for (int k = 0; k < XB_T / 32; ++k) total += s_w[k];
  const int k_stay = n - total;
  int k = before_blocks + wbase + incl - mine;   // packed position of this thread's first selected atom
#pragma unroll
  for (int j = 0; j < XB_A; ++j) {
    const int i = i0 + j;
    if (i >= n) break;
    const int kk = k;
    if (code[j]) ++k;
    if (!HALO) { leave[i] = code[j] ? 1 : 0; lpos[i] = kk; }   // what the restack reads
    if (!code[j] || kk >= cap) continue;
    const double4 p = A.posq[i];
    if (HALO) {
      idx_or_hole[kk] = i;
      double* b = buf + (size_t)kk * DLP_HALO_W;
      b[6] = (double)A.org_rank[i]; b[7] = (double)A.org_idx[i]; b[8] = (double)(A.org_wrap[i] + wrap_add);
      if (!d.lwrap) { b[0] = p.x; b[1] = p.y; b[2] = p.z; }
      else { b[0] = p.x + d.xadd; b[1] = p.y + d.yadd; b[2] = p.z + d.zadd; }   // deport_data.F90:1836-1844
      b[3] = (double)A.ltg[i];
      b[4] = (double)A.lsite[i];
      b[5] = (double)(A.ixyz[i] - (code[j] == 1 ? d.jxyz : d.kxyz));           // :1853
    } else {
      if (i < k_stay) idx_or_hole[kk] = i;      // holes below the new natms are the first ones (ascending)
      double* b = buf + (size_t)kk * 12;
      if (!d.lwrap) { b[0] = p.x; b[1] = p.y; b[2] = p.z; }
      else { b[0] = p.x + d.xadd; b[1] = p.y + d.yadd; b[2] = p.z + d.zadd; }   // deport_data.F90:296-305
      b[3] = A.vx[i]; b[4] = A.vy[i]; b[5] = A.vz[i];
      b[6] = A.fx[i]; b[7] = A.fy[i]; b[8] = A.fz[i];
      b[9] = (double)A.ltg[i]; b[10] = (double)A.lsite[i]; b[11] = (double)(A.ixyz[i] - d.jxyz);
    }
  }
  // the last block publishes count + sequence to the receiver and settles the local counts.  The barrier orders the
  // block's payload stores before thread 0's device-scope fence, the ticket chains the blocks, and only the last block pays
  // for the system-scope fence in front of the release store (fences are cumulative along this chain).
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(&dc[DC_TICKET], 1) == (int)gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence_system();
    dc[DC_TICKET] = 0;
    int sent = total;
    if (total > cap) { atomicOr(&dc[DC_ERR], HALO ? 2 : 1); sent = cap; }
    if (HALO) dc[DC_HSENT + q] = sent;
    else {
      dc[DC_RSENT + q] = sent; dc[DC_NOLD] = n; dc[DC_TOT] = total;
      if (total <= cap) { dc[DC_NATMS] = k_stay; dc[DC_NLAST] = k_stay; }
    }
    dst_hdr->count = sent;
    __threadfence_system();
    st_release_sys(&dst_hdr->seq, seq);
  }
}
__global__ void k_x_reloc_restack(const int* __restrict__ dc, int cap, const int* __restrict__ leave, const int* __restrict__ lpos,
                                  const int* __restrict__ hole_pos, XAtoms A) {
  const int n = dc[DC_NOLD], total = dc[DC_TOT];
  if (total > cap) return;
  const int k_stay = n - total;
  int j = k_stay + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || leave[j]) return;
  int r = k_stay - 1 - (j - lpos[j]);   // deport_data.F90:822-925, see k_reloc_restack
  int dst = hole_pos[r];
  A.posq[dst] = A.posq[j];
  A.vx[dst] = A.vx[j]; A.vy[dst] = A.vy[j]; A.vz[dst] = A.vz[j];
  A.fx[dst] = A.fx[j]; A.fy[dst] = A.fy[j]; A.fz[dst] = A.fz[j];
  A.ltg[dst] = A.ltg[j]; A.lsite[dst] = A.lsite[j]; A.ixyz[dst] = A.ixyz[j];
}
template <int HALO>
__global__ void k_x_recv(const XHdr* __restrict__ hdr, unsigned long long seq, const double* __restrict__ buf, int capacity, int q,
                         int* __restrict__ dc, XAtoms A) {
  __shared__ int s_count, s_last;
  if (threadIdx.x == 0) s_count = x_wait(&hdr->seq, seq, &dc[DC_ERR]) ? (int)hdr->count : 0;
  __syncthreads();
  const int off = HALO ? dc[DC_NLAST] : dc[DC_NATMS];
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < s_count && off + k < capacity) {
    const int i = off + k;
    if (HALO) {
      const double* b = buf + (size_t)k * DLP_HALO_W;
      A.org_rank[i] = __double2int_rn(__ldcg(b + 6)); A.org_idx[i] = __double2int_rn(__ldcg(b + 7)); A.org_wrap[i] = __double2int_rn(__ldcg(b + 8));
      A.posq[i] = make_double4(__ldcg(b), __ldcg(b + 1), __ldcg(b + 2), 0.0);
      A.ltg[i] = __double2int_rn(__ldcg(b + 3)); A.lsite[i] = __double2int_rn(__ldcg(b + 4)); A.ixyz[i] = __double2int_rn(__ldcg(b + 5));   // Nint, :1935-1940
      A.fx[i] = 0.0; A.fy[i] = 0.0; A.fz[i] = 0.0;
    } else {
      const double* b = buf + (size_t)k * 12;
      A.posq[i] = make_double4(__ldcg(b), __ldcg(b + 1), __ldcg(b + 2), 0.0);
      A.vx[i] = __ldcg(b + 3); A.vy[i] = __ldcg(b + 4); A.vz[i] = __ldcg(b + 5);
      A.fx[i] = __ldcg(b + 6); A.fy[i] = __ldcg(b + 7); A.fz[i] = __ldcg(b + 8);
      A.ltg[i] = __double2int_rn(__ldcg(b + 9)); A.lsite[i] = __double2int_rn(__ldcg(b + 10)); A.ixyz[i] = __double2int_rn(__ldcg(b + 11));
    }
  }
  // every block has read `off` before it takes a ticket; the last one appends the received atoms to the counts
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&dc[DC_TICKET], 1) == (int)gridDim.x - 1);
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    dc[DC_TICKET] = 0;
    int count = s_count;
    if (off + count > capacity) { atomicOr(&dc[DC_ERR], 8); count = capacity - off; }
    if (HALO) { dc[DC_HOFF + q] = off; dc[DC_HRECV + q] = count; dc[DC_NLAST] = off + count; }
    else { dc[DC_RRECV + q] = count; dc[DC_NATMS] = off + count; dc[DC_NLAST] = off + count; }
  }
}
__global__ void k_x_reloc_end(int* dc, const int* __restrict__ ixyz, const int* __restrict__ lsite,
                              const int* __restrict__ type_site, const double* __restrict__ charge_site, const int* __restrict__ freeze_site,
                              double4* __restrict__ posq, int* __restrict__ ltype, int* __restrict__ lfrzn) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= *(volatile int*)&dc[DC_NATMS]) return;
  if (ixyz[i] != 0) atomicAdd(&dc[DC_LOST], 1);       // deport_data.F90:3056-3058
  int s = lsite[i] - 1;
  ltype[i] = type_site[s]; posq[i].w = charge_site[s]; lfrzn[i] = freeze_site[s];   // :3062-3069
}
__global__ void k_x_halo_tag(const int* __restrict__ dc, Mat9 rcell, HaloThr t, const double4* __restrict__ posq, int* __restrict__ ixyz,
                             int my_rank, int* __restrict__ org_rank, int* __restrict__ org_idx, int* __restrict__ org_wrap) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dc[DC_NATMS]) return;
  org_rank[i] = my_rank; org_idx[i] = i; org_wrap[i] = 13;
  double4 p = posq[i];
  double x = rcell.m[0] * p.x + rcell.m[3] * p.y + rcell.m[6] * p.z;   // halo.F90:265-267
  double y = rcell.m[1] * p.x + rcell.m[4] * p.y + rcell.m[7] * p.z;
  double z = rcell.m[2] * p.x + rcell.m[5] * p.y + rcell.m[8] * p.z;
  int v = 0;
  if (x <= t.ecwx) v += 1;
  if (x >= t.cwx) v += 2;
  if (y <= t.ecwy) v += 10;
  if (y >= t.cwy) v += 20;
  if (z <= t.ecwz) v += 100;
  if (z >= t.cwz) v += 200;
  ixyz[i] = v;
}
__global__ void k_x_halo_end(const int* __restrict__ dc, const int* __restrict__ lsite, const int* __restrict__ type_site,
                             const double* __restrict__ charge_site, const int* __restrict__ freeze_site, double4* __restrict__ posq,
                             int* __restrict__ ltype, int* __restrict__ lfrzn, double* xbg, double* ybg, double* zbg) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= dc[DC_NLAST]) return;
  double4 p = posq[i];
  if (i >= dc[DC_NATMS]) {   // halo.F90:296-302
    int s = lsite[i] - 1;
    ltype[i] = type_site[s]; p.w = charge_site[s]; lfrzn[i] = freeze_site[s];
    posq[i] = p;
  }
  xbg[i] = p.x; ybg[i] = p.y; zbg[i] = p.z;   // vnl_set_check, halo.F90:315 / neighbours.F90:337-341
}
// gmax of vnl_check (neighbours.F90:176) through the peers' mailboxes: lane r writes this rank's value into rank r's slot
// and waits for rank r's value in its own.  Slots alternate with the parity of the sequence number.
// The same message carries the 16 partial sums (energies, virials, stress) of the force call that completed before it on this
// stream, i.e. of the PREVIOUS step, and every rank adds the ranks' values in rank order: the gsum of two_body.F90:729 and
// drivers.F90:795 without an extra launch, collective or host round trip (deterministic: fixed summation order).
struct XGm { unsigned long long seq; unsigned long long tol; double sums[16]; };
__global__ void k_x_gmax(int rank, int nranks, unsigned long long seq, const unsigned long long* __restrict__ tol_bits,
                         const double* __restrict__ my_sums, const unsigned long long* __restrict__ peers, size_t off_gm,
                         unsigned long long* __restrict__ out /*[1 + 16]*/, int* __restrict__ dc,
                         unsigned long long* __restrict__ host_out /* pinned, mapped: [0..16] values, [17] error bits, [18] seq */) {
  const int r = threadIdx.x;
  unsigned long long v = 0;
  const size_t slot = ((seq & 1) * (size_t)nranks);
  if (r < nranks) {
    const unsigned long long mine = *tol_bits;
    XGm* dst = reinterpret_cast<XGm*>(peers[r] + off_gm) + slot + rank;
    dst->tol = mine;
#pragma unroll
    for (int k = 0; k < 16; ++k) dst->sums[k] = my_sums[k];
    __threadfence_system();
    st_release_sys(&dst->seq, seq);
    const XGm* src = reinterpret_cast<const XGm*>(peers[rank] + off_gm) + slot + r;
    if (x_wait(&src->seq, seq, &dc[DC_ERR])) v = ld_acquire_sys(&src->tol);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { unsigned long long o = __shfl_xor_sync(DLP_FULL, v, d); v = o > v ? o : v; }   // non-negative doubles order like their bits
  if (r == 0) { out[0] = v; host_out[0] = v; }
  __syncwarp();   // every peer's message has been acquired by the lane that waited for it
  if (r < 16) {
    double t = 0.0;
    for (int q = 0; q < nranks; ++q) {
      const XGm* src = reinterpret_cast<const XGm*>(peers[rank] + off_gm) + slot + q;
      t += __longlong_as_double((long long)ld_acquire_sys(reinterpret_cast<const unsigned long long*>(&src->sums[r])));
    }
    out[1 + r] = (unsigned long long)__double_as_longlong(t);
    host_out[1 + r] = (unsigned long long)__double_as_longlong(t);
  }
  // the host polls host_out[18] for this sequence number instead of synchronising the stream (no DMA, no wake-up): the values
  // above are written straight into its pinned memory and fenced before the flag
  __threadfence_system();
  __syncwarp();
  if (r == 0) {
    host_out[17] = (unsigned long long)(unsigned)dc[DC_ERR];
    __threadfence_system();
    st_release_sys(&host_out[18], seq);
  }
}

// ---------------------------------------------------------------- migration over the compact list of movers
// A rebuild moves a few hundred of a million atoms across the domain faces, and the six stages of deport_atomic_data each
// scanned every atom for them (count + pack + restack + receive: 24 launches, ~0.2 ms of launch latency).  Here the atoms with a
// non-zero relocation tag are compacted ONCE into an ascending index list M, and every stage is ONE single-block kernel over M:
// select (deport_data.F90:254-274), pack in ascending order into the receiver's buffer (:290-325), restack the tail into the
// holes (:822-925), publish, re-filter M (an atom that moved into a hole brings its tag along), wait for this rank's own
// message, append the received atoms (:927-960) and the movers among them.  Same buffers, order, counts and error flags as the
// scanning kernels, which is what the multi-rank parity tests hold it to.
#define DC_NMOV 7
#define XM_T 1024
__device__ __forceinline__ int xm_block_excl_scan(int v, int* s_w, int* total) {   // exclusive scan over the block + block total
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(DLP_FULL, incl, o); if (lane >= o) incl += u; }
  __syncthreads();
  if (lane == 31) s_w[w] = incl;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int k = 0; k < XM_T / 32; ++k) { const int u = s_w[k]; tot += u; if (k < w) base += u; }
  *total = tot;
  return base + incl - v;
}
// movers of the domain, ascending: per-block counts (k_x_mov_count), then the indices (k_x_mov_pack, the scan of k_x_pack)
__global__ void __launch_bounds__(XB_T) k_x_mov_count(const int* __restrict__ dc, const int* __restrict__ ixyz, int* __restrict__ blocksum) {
  __shared__ int s_w[XB_T / 32];
  const int n = dc[DC_NATMS];
  const int i0 = blockIdx.x * XB_N + threadIdx.x * XB_A;
  int c = 0;
#pragma unroll
  for (int j = 0; j < XB_A; ++j) if (i0 + j < n) c += ixyz[i0 + j] != 0;
  const int t = x_block_sum(c, s_w);
  if (threadIdx.x == 0) blocksum[blockIdx.x] = t;
}
__global__ void __launch_bounds__(XB_T)
k_x_mov_pack(int* __restrict__ dc, int capM, const int* __restrict__ blocksum, const int* __restrict__ ixyz, int* __restrict__ M) {
  __shared__ int s_w[XB_T / 32], s_w2[XB_T / 32];
  const int n = dc[DC_NATMS];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int i0 = blockIdx.x * XB_N + threadIdx.x * XB_A;
  int f[XB_A], mine = 0;
#pragma unroll
  for (int j = 0; j < XB_A; ++j) { f[j] = (i0 + j < n) && ixyz[i0 + j] != 0; mine += f[j]; }
  int before = 0, all = 0;
  for (int k = threadIdx.x; k < gridDim.x; k += XB_T) { const int v = blocksum[k]; all += v; if (k < blockIdx.x) before += v; }
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(DLP_FULL, incl, o); if (lane >= o) incl += v; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { before += __shfl_xor_sync(DLP_FULL, before, o); all += __shfl_xor_sync(DLP_FULL, all, o); }
  if (lane == 31) s_w[w] = incl;
  if (lane == 0) s_w2[w] = before;
  __syncthreads();
  int wbase = 0, bb = 0;
#pragma unroll
  for (int k = 0; k < XB_T / 32; ++k) { if (k < w) wbase += s_w[k]; bb += s_w2[k]; }
  __syncthreads();
  if (lane == 0) s_w[w] = all;
  __syncthreads();
  int k = bb + wbase + incl - mine;
#pragma unroll
  for (int j = 0; j < XB_A; ++j) if (f[j]) { if (k < capM) M[k] = i0 + j; ++k; }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    int total = 0;
    for (int q = 0; q < XB_T / 32; ++q) total += s_w[q];
    if (total > capM) { atomicOr(&dc[DC_ERR], 1); total = capM; }   // more movers than the migration buffers of all stages hold
    dc[DC_NMOV] = total;
  }
}
__device__ __forceinline__ void xm_move_atom(const XAtoms& A, int dst, int src) {
  A.posq[dst] = A.posq[src];
  A.vx[dst] = A.vx[src]; A.vy[dst] = A.vy[src]; A.vz[dst] = A.vz[src];
  A.fx[dst] = A.fx[src]; A.fy[dst] = A.fy[src]; A.fz[dst] = A.fz[src];
  A.ltg[dst] = A.ltg[src]; A.lsite[dst] = A.lsite[src]; A.ixyz[dst] = A.ixyz[src];
}
__global__ void __launch_bounds__(XM_T)
k_x_reloc_stage(int* __restrict__ dc, Dir d, int cap, int q, int* __restrict__ M, int capM, int* __restrict__ S, XAtoms A,
                double* __restrict__ buf_dst, XHdr* hdr_dst, const XHdr* __restrict__ hdr_me, const double* __restrict__ buf_me, int capacity,
                unsigned long long seq) {
  __shared__ int s_w[XM_T / 32];
  __shared__ int s_count;
  const int tid = threadIdx.x;
  const int n = dc[DC_NATMS], nm = dc[DC_NMOV];
  // ---- 1. the movers that leave in this direction, ascending: S[k]
  int total = 0;
  for (int base = 0; base < nm; base += XM_T) {
    const int e = base + tid;
    const int i = e < nm ? M[e] : -1;
    const int sel = (i >= 0 && i < n) ? reloc_sel(A.ixyz[i], d) : 0;
    int chunk;
    const int pos = total + xm_block_excl_scan(sel, s_w, &chunk);
    if (sel && pos < cap) S[pos] = i;
    total += chunk;
  }
  __syncthreads();
  const bool fits = total <= cap;
  const int sent = fits ? total : cap;
  const int k_stay = n - total;
  // ---- 2. pack (the tag loses this direction, :325)
  for (int k = tid; k < sent; k += XM_T) {
    const int i = S[k];
    const double4 p = A.posq[i];
    double* b = buf_dst + (size_t)k * 12;
    if (!d.lwrap) { b[0] = p.x; b[1] = p.y; b[2] = p.z; }
    else { b[0] = p.x + d.xadd; b[1] = p.y + d.yadd; b[2] = p.z + d.zadd; }   // deport_data.F90:296-305
    b[3] = A.vx[i]; b[4] = A.vy[i]; b[5] = A.vz[i];
    b[6] = A.fx[i]; b[7] = A.fy[i]; b[8] = A.fz[i];
    b[9] = (double)A.ltg[i]; b[10] = (double)A.lsite[i]; b[11] = (double)(A.ixyz[i] - d.jxyz);
  }
  __syncthreads();
  // ---- 3. restack (:822-925): the r-th hole (ascending; holes = the leavers below the new natms = a prefix of S) takes the r-th
  // staying atom counted from the end.  Sources sit at or above k_stay, holes below it.
  if (fits) {
    for (int j = k_stay + tid; j < n; j += XM_T) {
      int lo = 0, hi = total;                     // leavers before j
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (S[mid] < j) lo = mid + 1; else hi = mid; }
      if (lo < total && S[lo] == j) continue;     // j itself leaves
      const int r = k_stay - 1 - (j - lo);
      xm_move_atom(A, S[r], j);
    }
  }
  __syncthreads();
  // ---- 4. publish to the receiver, settle the local counts
  if (tid == 0) {
    if (!fits) atomicOr(&dc[DC_ERR], 1);
    dc[DC_RSENT + q] = sent;
    __threadfence_system();
    hdr_dst->count = sent;
    __threadfence_system();
    st_release_sys(&hdr_dst->seq, seq);
  }
  const int nat1 = fits ? k_stay : n;
  // ---- 5. the movers that are still here (a hole may have received one), ascending, compacted in place
  int nm2 = 0;
  for (int base = 0; base < nm; base += XM_T) {
    const int e = base + tid;
    const int i = e < nm ? M[e] : -1;
    const int keep = (i >= 0 && i < nat1 && A.ixyz[i] != 0) ? 1 : 0;
    int chunk;
    const int pos = nm2 + xm_block_excl_scan(keep, s_w, &chunk);   // the scan's barriers separate this chunk's reads from its writes
    if (keep) M[pos] = i;
    nm2 += chunk;
  }
  __syncthreads();
  // ---- 6. this rank's own message for the stage
  if (tid == 0) s_count = x_wait(&hdr_me->seq, seq, &dc[DC_ERR]) ? (int)hdr_me->count : 0;
  __syncthreads();
  int count = s_count;
  const bool room = nat1 + count <= capacity;
  if (!room) count = max(capacity - nat1, 0);
  for (int base = 0; base < count; base += XM_T) {
    const int k = base + tid;
    int mover = 0, i = -1;
    if (k < count) {
      i = nat1 + k;
      const double* b = buf_me + (size_t)k * 12;
      A.posq[i] = make_double4(__ldcg(b), __ldcg(b + 1), __ldcg(b + 2), 0.0);
      A.vx[i] = __ldcg(b + 3); A.vy[i] = __ldcg(b + 4); A.vz[i] = __ldcg(b + 5);
      A.fx[i] = __ldcg(b + 6); A.fy[i] = __ldcg(b + 7); A.fz[i] = __ldcg(b + 8);
      const int tag = __double2int_rn(__ldcg(b + 11));
      A.ltg[i] = __double2int_rn(__ldcg(b + 9)); A.lsite[i] = __double2int_rn(__ldcg(b + 10)); A.ixyz[i] = tag;
      mover = tag != 0;
    }
    int chunk;
    const int pos = nm2 + xm_block_excl_scan(mover, s_w, &chunk);
    if (mover && pos < capM) M[pos] = i;
    nm2 += chunk;
  }
  if (tid == 0) {
    if (!room) atomicOr(&dc[DC_ERR], 8);
    if (nm2 > capM) { atomicOr(&dc[DC_ERR], 1); nm2 = capM; }
    dc[DC_RRECV + q] = count;
    dc[DC_NATMS] = nat1 + count; dc[DC_NLAST] = nat1 + count;
    dc[DC_NMOV] = nm2;
  }
}

static size_t x_align(size_t v) { return (v + 255) & ~(size_t)255; }
struct XLayout { size_t off_gm, off_hdr, off_rbuf, off_hbuf, bytes; };
static XLayout x_layout(int nranks, int cap_r, int cap_h) {
  XLayout L;
  L.off_gm = 0;
  L.off_hdr = x_align((size_t)2 * nranks * sizeof(XGm));
  L.off_rbuf = x_align(L.off_hdr + 12 * sizeof(XHdr));
  L.off_hbuf = x_align(L.off_rbuf + (size_t)6 * cap_r * 12 * sizeof(double));
  L.bytes = x_align(L.off_hbuf + (size_t)6 * cap_h * DLP_HALO_W * sizeof(double));
  return L;
}


}  // namespace

// a peer rank that lives in this process on ANOTHER device: enable direct access once (no-op for the same device)
static int x_enable_peer(dlpgpu_ctx* ctx, int dev) {
  if (dev == ctx->device) return 0;
  int can = 0;
  CK(cudaDeviceCanAccessPeer(&can, ctx->device, dev));
  if (!can) return dlp_fail(ctx, DLPGPU_ERR_CUDA, "device %d cannot access the memory of device %d", ctx->device, dev);
  cudaError_t e = cudaDeviceEnablePeerAccess(dev, 0);
  if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return dlp_fail(ctx, DLPGPU_ERR_CUDA, "cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e));
  cudaGetLastError();
  return 0;
}

int dlp_vnl_check(dlpgpu_ctx* ctx, double* tol) {
  cudaStream_t s = ctx->stream;
  if (!ctx->have_bg) return dlp_fail(ctx, DLPGPU_ERR_STATE, "vnl_check: no checkpoint");
  double rc[9];
  h_invert(ctx->cell, rc);
  if (!ctx->tol_fresh) {   // otherwise the fused velocity-Verlet stage 1 already left the maximum there
    CK(cudaMemsetAsync(ctx->tol_bits.p, 0, sizeof(unsigned long long), s));
    if (ctx->natms > 0)
      LAUNCH(ctx, k_vnl_tol, cdiv(ctx->natms, 256), 256, 0, ctx->natms, ctx->imcon, mat(ctx->cell), mat(rc), ctx->posq.p, ctx->xbg.p,
             ctx->ybg.p, ctx->zbg.p, ctx->tol_bits.p);
  }
  ctx->tol_fresh = false;
  unsigned long long bits = 0;
  CK(cudaMemcpyAsync(&bits, ctx->tol_bits.p, sizeof bits, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  double r;
  std::memcpy(&r, &bits, sizeof r);
  *tol = r;
  return 0;
}

extern "C" {

int dlpgpu_dev_vnl_check(dlpgpu_ctx* ctx, double* tol) {
  if (!ctx || !tol) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  return dlp_vnl_check(ctx, tol);
}

int dlpgpu_dev_halo_begin(dlpgpu_ctx* ctx) {
  if (!ctx) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  // thresholds, halo.F90:190-249
  double cut = ctx->rx + 1.0e-6;
  double w3[3];
  h_widths(ctx->cell, w3);
  double nxr = (double)ctx->nx, nyr = (double)ctx->ny, nzr = (double)ctx->nz;
  int nlx = (int)(w3[0] / (cut * nxr)), nly = (int)(w3[1] / (cut * nyr)), nlz = (int)(w3[2] / (cut * nzr));
  if (nlx * nly * nlz == 0) return dlp_fail(ctx, DLPGPU_ERR_LINK_CELLS, "error 307: domain narrower than cutoff_extended");
  double xdc = (double)(nlx * ctx->nx), ydc = (double)(nly * ctx->ny), zdc = (double)(nlz * ctx->nz);
  double cwx = 1.0 / xdc, cwy = 1.0 / ydc, cwz = 1.0 / zdc;
  double ecwx = std::max(cwx, ctx->ecw[0]), ecwy = std::max(cwy, ctx->ecw[1]), ecwz = std::max(cwz, ctx->ecw[2]);
  double nx_recip = 1.0 / nxr, ny_recip = 1.0 / nyr, nz_recip = 1.0 / nzr;
  const double zero_plus = DBL_MIN;
  HaloThr t;
  t.ecwx = std::nextafter((-0.5 + ecwx) + (double)ctx->idx * nx_recip, DBL_MAX) + zero_plus;
  t.ecwy = std::nextafter((-0.5 + ecwy) + (double)ctx->idy * ny_recip, DBL_MAX) + zero_plus;
  t.ecwz = std::nextafter((-0.5 + ecwz) + (double)ctx->idz * nz_recip, DBL_MAX) + zero_plus;
  t.cwx = std::nextafter((-0.5 - cwx) + (double)(ctx->idx + 1) * nx_recip, -DBL_MAX) - zero_plus - (nlx == 1 ? cwx * 1.0e-10 : 0.0);
  t.cwy = std::nextafter((-0.5 - cwy) + (double)(ctx->idy + 1) * ny_recip, -DBL_MAX) - zero_plus - (nly == 1 ? cwy * 1.0e-10 : 0.0);
  t.cwz = std::nextafter((-0.5 - cwz) + (double)(ctx->idz + 1) * nz_recip, -DBL_MAX) - zero_plus - (nlz == 1 ? cwz * 1.0e-10 : 0.0);
  double rc[9];
  h_invert(ctx->cell, rc);
  ctx->nlast = ctx->natms;   // halo.F90:259
  if (ctx->natms > 0)
    LAUNCH(ctx, k_halo_tag, cdiv(ctx->natms, 256), 256, 0, ctx->natms, mat(rc), t, ctx->posq.p, ctx->ixyz.p, ctx->p2p_rank, ctx->org_rank.p,
           ctx->org_idx.p, ctx->org_wrap.p);
  ctx->halo_valid = false; ctx->list_valid = false;
  for (int q = 0; q < 6; ++q) { ctx->stage[q].count = 0; ctx->stage[q].recv_count = 0; ctx->stage[q].recv_off = ctx->natms; }
  return 0;
}

int dlpgpu_dev_halo_pack(dlpgpu_ctx* ctx, int mdir, double* sendbuf_dev, int capacity_atoms, int* count) {
  if (!ctx || !count || stage_of(mdir) < 0) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  Dir d = dir_settings(ctx, mdir);
  HaloStage& st = ctx->stage[stage_of(mdir)];
  int n = ctx->nlast;
  CK(ctx->flag.ensure((size_t)n + 2, s)); CK(ctx->scan_out.ensure((size_t)n + 2, s));
  LAUNCH(ctx, k_halo_flag, cdiv(n + 1, 256), 256, 0, n, d, ctx->ixyz.p, ctx->flag.p);
  int total = 0;
  CKRC(dlp_exclusive_scan(ctx, ctx->flag.p, ctx->scan_out.p, n, &total));
  st.count = total; st.lwrap = d.lwrap != 0; st.shift[0] = d.xadd; st.shift[1] = d.yadd; st.shift[2] = d.zadd;
  *count = total;
  if (total > capacity_atoms || (total > 0 && !sendbuf_dev))
    return dlp_fail(ctx, DLPGPU_ERR_BUFFER, "error 54: outgoing halo buffer too small (%d atoms > capacity %d)", total, capacity_atoms);
  CK(st.idx.ensure((size_t)total + 1, s));
  {
    // wrap code increment of this stage: the sender applies uuu/vvv/www = +-1 on its axis when it sits on the cell border
    int wrap_add = 0;
    if (d.lwrap) {
      const int sgn = (mdir < 0) ? +1 : -1;                 // -dir exports from the low border shift by +cell, +dir by -cell
      wrap_add = sgn * (d.kx ? 1 : d.ky ? 3 : 9);
    }
    if (total > 0)
      LAUNCH(ctx, k_halo_pack, cdiv(n, 256), 256, 0, n, d, capacity_atoms, ctx->flag.p, ctx->scan_out.p, ctx->posq.p, ctx->ltg.p, ctx->lsite.p,
             ctx->ixyz.p, ctx->org_rank.p, ctx->org_idx.p, ctx->org_wrap.p, wrap_add, sendbuf_dev, st.idx.p);
  }
  return 0;
}

int dlpgpu_dev_halo_unpack(dlpgpu_ctx* ctx, int mdir, const double* recvbuf_dev, int count) {
  if (ctx) dlp_hostio_ints_stale(ctx);
  if (!ctx || count < 0 || stage_of(mdir) < 0) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  HaloStage& st = ctx->stage[stage_of(mdir)];
  CKRC(dlp_ensure_atoms(ctx, ctx->nlast + count + 16));
  st.recv_off = ctx->nlast; st.recv_count = count;
  if (count > 0)
    LAUNCH(ctx, k_halo_unpack, cdiv(count, 256), 256, 0, count, ctx->nlast, recvbuf_dev, ctx->posq.p, ctx->ltg.p, ctx->lsite.p, ctx->ixyz.p,
           ctx->fx.p, ctx->fy.p, ctx->fz.p, ctx->org_rank.p, ctx->org_idx.p, ctx->org_wrap.p);
  ctx->nlast += count;
  return 0;
}

int dlpgpu_dev_halo_end(dlpgpu_ctx* ctx) {
  if (ctx) dlp_hostio_ints_stale(ctx);
  if (!ctx) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (ctx->nsites < 1) return dlp_fail(ctx, DLPGPU_ERR_STATE, "halo_end: sites not set");
  int nh = ctx->nlast - ctx->natms;
  if (nh > 0)
    LAUNCH(ctx, k_assign_sites, cdiv(nh, 256), 256, 0, ctx->natms, ctx->nlast, ctx->lsite.p, ctx->type_site.p, ctx->charge_site.p,
           ctx->freeze_site.p, ctx->posq.p, ctx->ltype.p, ctx->lfrzn.p);
  CKRC(dlp_vnl_set_check(ctx));   // halo.F90:315
  ctx->halo_valid = true;
  return 0;
}

int dlpgpu_dev_halo_serial(dlpgpu_ctx* ctx) {
  if (!ctx) return DLPGPU_ERR_ARG;
  if (ctx->nx * ctx->ny * ctx->nz != 1) return dlp_fail(ctx, DLPGPU_ERR_STATE, "halo_serial: domain decomposition is not 1x1x1");
  CKRC(dlpgpu_dev_halo_begin(ctx));
  const int mdirs[6] = {-1, 1, -2, 2, -3, 3};
  for (int q = 0; q < 6; ++q) {
    // jmove = imove (deport_data.F90:1884-1886): the rank is its own neighbour.  Size the buffer from the flag count.
    int cnt = 0;
    int cap = (int)(ctx->xfer.cap / DLP_HALO_W);
    int rc = dlpgpu_dev_halo_pack(ctx, mdirs[q], ctx->xfer.p, cap, &cnt);
    if (rc == DLPGPU_ERR_BUFFER) {
      CK(ctx->xfer.ensure((size_t)cnt * DLP_HALO_W + 64, ctx->stream));
      rc = dlpgpu_dev_halo_pack(ctx, mdirs[q], ctx->xfer.p, (int)(ctx->xfer.cap / DLP_HALO_W), &cnt);
    }
    if (rc) return rc;
    CKRC(dlpgpu_dev_halo_unpack(ctx, mdirs[q], ctx->xfer.p, cnt));
  }
  return dlpgpu_dev_halo_end(ctx);
}

int dlpgpu_dev_refresh_pack(dlpgpu_ctx* ctx, int mdir, double* sendbuf_dev, int* count) {
  if (!ctx || !count || stage_of(mdir) < 0) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->halo_valid) return dlp_fail(ctx, DLPGPU_ERR_STATE, "refresh: no halo has been built");
  HaloStage& st = ctx->stage[stage_of(mdir)];
  Dir d = dir_settings(ctx, mdir);
  *count = st.count;
  if (st.count > 0) {
    if (!sendbuf_dev) return DLPGPU_ERR_ARG;
    LAUNCH(ctx, k_refresh_pack, cdiv(st.count, 256), 256, 0, st.count, d, st.idx.p, ctx->posq.p, sendbuf_dev);
  }
  return 0;
}

int dlpgpu_dev_refresh_unpack(dlpgpu_ctx* ctx, int mdir, const double* recvbuf_dev, int count) {
  if (!ctx || stage_of(mdir) < 0) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  HaloStage& st = ctx->stage[stage_of(mdir)];
  if (count != st.recv_count)   // halo.F90:104-106
    return dlp_fail(ctx, DLPGPU_ERR_HALO_COUNT, "error 138: refreshed halo size %d differs from the built one %d", count, st.recv_count);
  if (count > 0) LAUNCH(ctx, k_refresh_unpack, cdiv(count, 256), 256, 0, count, st.recv_off, recvbuf_dev, ctx->posq.p);
  return 0;
}

// ---- peer-memory refresh
int dlpgpu_dev_p2p_init(dlpgpu_ctx* ctx, int rank, int nranks, int capacity_atoms, unsigned char handles_out[DLPGPU_P2P_BLOB]) {
  if (!ctx || rank < 0 || nranks < 1 || rank >= nranks || capacity_atoms < 1 || !handles_out) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (ctx->pub[0]) return dlp_fail(ctx, DLPGPU_ERR_STATE, "p2p_init: already initialised");
  ctx->p2p_rank = rank; ctx->p2p_nranks = nranks; ctx->pub_cap = capacity_atoms;
  std::memset(handles_out, 0, DLPGPU_P2P_BLOB);
  for (int b = 0; b < 2; ++b) {
    CK(cudaMalloc((void**)&ctx->pub[b], (size_t)capacity_atoms * sizeof(double4)));
    if (nranks > 1) {
      cudaIpcMemHandle_t h;
      CK(cudaIpcGetMemHandle(&h, ctx->pub[b]));
      static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
      std::memcpy(handles_out + 64 * b, &h, 64);
    }
  }
  {   // ranks that live in the SAME process (one host thread per rank) reach each other's buffers by plain pointers
    const long long pid = (long long)getpid(), dev = ctx->device;
    const unsigned long long p0 = (unsigned long long)(size_t)ctx->pub[0], p1 = (unsigned long long)(size_t)ctx->pub[1];
    std::memcpy(handles_out + 128, &pid, 8); std::memcpy(handles_out + 136, &p0, 8); std::memcpy(handles_out + 144, &p1, 8);
    std::memcpy(handles_out + 152, &dev, 8);
  }
  ctx->peer_pub.assign((size_t)2 * nranks, nullptr);
  ctx->peer_pub[2 * rank] = ctx->pub[0]; ctx->peer_pub[2 * rank + 1] = ctx->pub[1];
  ctx->p2p_ready = (nranks == 1);
  if (ctx->p2p_ready) {
    CK(ctx->peer_pub_dev.ensure(2, ctx->stream));
    CK(cudaMemcpy(ctx->peer_pub_dev.p, ctx->peer_pub.data(), 2 * sizeof(void*), cudaMemcpyHostToDevice));
  }
  return 0;
}

int dlpgpu_dev_p2p_open(dlpgpu_ctx* ctx, const unsigned char* all_handles /* nranks x DLPGPU_P2P_BLOB bytes, rank-major */) {
  if (!ctx || !all_handles) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->pub[0]) return dlp_fail(ctx, DLPGPU_ERR_STATE, "p2p_open: call p2p_init first");
  for (int r = 0; r < ctx->p2p_nranks; ++r) {
    if (r == ctx->p2p_rank) continue;
    const unsigned char* blob = all_handles + (size_t)DLPGPU_P2P_BLOB * r;
    long long pid = 0, dev = 0;
    std::memcpy(&pid, blob + 128, 8); std::memcpy(&dev, blob + 152, 8);
    if (pid == (long long)getpid()) {   // same process: the pointer itself (peer access between two devices enabled on demand)
      CKRC(x_enable_peer(ctx, (int)dev));
      for (int b = 0; b < 2; ++b) {
        unsigned long long q = 0;
        std::memcpy(&q, blob + 136 + 8 * b, 8);
        ctx->peer_pub[2 * r + b] = (double4*)(size_t)q;
      }
      ctx->peer_pub_local.resize((size_t)ctx->p2p_nranks, 0); ctx->peer_pub_local[r] = 1;
      continue;
    }
    for (int b = 0; b < 2; ++b) {
      cudaIpcMemHandle_t h;
      std::memcpy(&h, blob + 64 * b, 64);
      void* p = nullptr;
      CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
      ctx->peer_pub[2 * r + b] = (double4*)p;
    }
  }
  CK(ctx->peer_pub_dev.ensure((size_t)2 * ctx->p2p_nranks, ctx->stream));
  CK(cudaMemcpy(ctx->peer_pub_dev.p, ctx->peer_pub.data(), (size_t)2 * ctx->p2p_nranks * sizeof(void*), cudaMemcpyHostToDevice));
  ctx->p2p_ready = true;
  return 0;
}

// copies the local coordinates into the peer-visible buffer of the next parity; call once per step after the positions moved
int dlpgpu_dev_publish(dlpgpu_ctx* ctx) {
  if (!ctx) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->pub[0]) return dlp_fail(ctx, DLPGPU_ERR_STATE, "publish: p2p not initialised");
  if (ctx->natms > ctx->pub_cap) return dlp_fail(ctx, DLPGPU_ERR_BUFFER, "publish: %d local atoms exceed the peer buffer (%d)", ctx->natms, ctx->pub_cap);
  ctx->pub_parity ^= 1;
  if (ctx->natms > 0 && !ctx->pub_fresh)   // the fused velocity-Verlet stage 1 may have filled the buffer of the next parity already
    LAUNCH(ctx, k_publish, cdiv(ctx->natms, 256), 256, 0, ctx->natms, ctx->posq.p, ctx->pub[ctx->pub_parity]);
  ctx->pub_fresh = false;
  ctx->pub_valid = true;
  return 0;
}

// refresh_halo_positions in one kernel.  Every rank must have published this step's coordinates and a collective on the
// same streams (the gmax of vnl_check) must separate the publishes from the pulls.
int dlpgpu_dev_refresh_pull(dlpgpu_ctx* ctx) {
  if (!ctx) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->p2p_ready || !ctx->pub_valid) return dlp_fail(ctx, DLPGPU_ERR_STATE, "refresh_pull: peer buffers not ready / nothing published");
  if (!ctx->halo_valid) return dlp_fail(ctx, DLPGPU_ERR_STATE, "refresh: no halo has been built");
  const int nh = ctx->nlast - ctx->natms;
  if (nh > 0)
    LAUNCH(ctx, k_refresh_pull, cdiv(nh, 256), 256, 0, ctx->natms, ctx->nlast, ctx->pub_parity, ctx->peer_pub_dev.p, mat(ctx->cell),
           ctx->org_rank.p, ctx->org_idx.p, ctx->org_wrap.p, ctx->posq.p);
  return 0;
}

int dlpgpu_dev_halo_stage_counts(dlpgpu_ctx* ctx, int sent[6], int received[6]) {
  if (!ctx || !sent || !received) return DLPGPU_ERR_ARG;
  if (!ctx->halo_valid) return dlp_fail(ctx, DLPGPU_ERR_STATE, "halo_stage_counts: no halo has been built");
  for (int q = 0; q < 6; ++q) { sent[q] = ctx->stage[q].count; received[q] = ctx->stage[q].recv_count; }
  return 0;
}

int dlpgpu_dev_refresh_serial(dlpgpu_ctx* ctx) {
  if (!ctx) return DLPGPU_ERR_ARG;
  const int mdirs[6] = {-1, 1, -2, 2, -3, 3};
  for (int q = 0; q < 6; ++q) {
    int cnt = 0;
    CK(ctx->xfer.ensure((size_t)ctx->stage[q].count * 3 + 64, ctx->stream));
    CKRC(dlpgpu_dev_refresh_pack(ctx, mdirs[q], ctx->xfer.p, &cnt));
    CKRC(dlpgpu_dev_refresh_unpack(ctx, mdirs[q], ctx->xfer.p, cnt));
  }
  return 0;
}

int dlpgpu_dev_relocate_serial(dlpgpu_ctx* ctx) {
  if (!ctx) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  double rc[9];
  h_invert(ctx->cell, rc);
  if (ctx->natms > 0 && ctx->imcon != 0)
    LAUNCH(ctx, k_pbcshift, cdiv(ctx->natms, 256), 256, 0, ctx->natms, ctx->imcon, mat(ctx->cell), mat(rc), ctx->posq.p);
  ctx->nlast = ctx->natms;
  ctx->halo_valid = false; ctx->list_valid = false;
  ctx->tol_fresh = false; ctx->pub_fresh = false;
  return 0;
}

int dlpgpu_dev_relocate_begin(dlpgpu_ctx* ctx) {
  if (!ctx) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  double rc[9];
  h_invert(ctx->cell, rc);
  DomI D{ctx->nx, ctx->ny, ctx->nz, ctx->idx, ctx->idy, ctx->idz};
  ctx->nlast = ctx->natms;
  if (ctx->natms > 0) LAUNCH(ctx, k_reloc_tag, cdiv(ctx->natms, 256), 256, 0, ctx->natms, mat(rc), D, ctx->posq.p, ctx->ixyz.p);
  ctx->halo_valid = false; ctx->list_valid = false;
  ctx->tol_fresh = false; ctx->pub_fresh = false;
  return 0;
}

int dlpgpu_dev_relocate_pack(dlpgpu_ctx* ctx, int mdir, double* sendbuf_dev, int capacity_atoms, int* count) {
  if (!ctx || !count || stage_of(mdir) < 0) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  Dir d = dir_settings(ctx, mdir);
  int n = ctx->natms;
  CK(ctx->flag.ensure((size_t)n + 2, s)); CK(ctx->scan_out.ensure((size_t)n + 2, s));
  LAUNCH(ctx, k_reloc_flag, cdiv(n + 1, 256), 256, 0, n, d, ctx->ixyz.p, ctx->flag.p);
  int total = 0;
  CKRC(dlp_exclusive_scan(ctx, ctx->flag.p, ctx->scan_out.p, n, &total));
  *count = total;
  if (total == 0) return 0;
  if (total > capacity_atoms || !sendbuf_dev)
    return dlp_fail(ctx, DLPGPU_ERR_BUFFER, "error 43: outgoing migration buffer too small (%d atoms > capacity %d)", total, capacity_atoms);
  int k_stay = n - total;
  CK(ctx->hole_pos.ensure((size_t)total + 1, s));
  LAUNCH(ctx, k_reloc_pack, cdiv(n, 256), 256, 0, n, k_stay, d, capacity_atoms, ctx->flag.p, ctx->scan_out.p, ctx->posq.p, ctx->vx.p, ctx->vy.p,
         ctx->vz.p, ctx->fx.p, ctx->fy.p, ctx->fz.p, ctx->ltg.p, ctx->lsite.p, ctx->ixyz.p, sendbuf_dev, ctx->hole_pos.p);
  if (n - k_stay > 0 && k_stay > 0)
    LAUNCH(ctx, k_reloc_restack, cdiv(n - k_stay, 256), 256, 0, n, k_stay, ctx->flag.p, ctx->scan_out.p, ctx->hole_pos.p, ctx->posq.p,
           ctx->vx.p, ctx->vy.p, ctx->vz.p, ctx->fx.p, ctx->fy.p, ctx->fz.p, ctx->ltg.p, ctx->lsite.p, ctx->ixyz.p);
  ctx->natms = k_stay; ctx->nlast = k_stay;
  return 0;
}

int dlpgpu_dev_relocate_unpack(dlpgpu_ctx* ctx, int mdir, const double* recvbuf_dev, int count) {
  if (ctx) dlp_hostio_ints_stale(ctx);
  if (!ctx || count < 0 || stage_of(mdir) < 0) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (count == 0) return 0;
  CKRC(dlp_ensure_atoms(ctx, ctx->natms + count + 16));
  LAUNCH(ctx, k_reloc_unpack, cdiv(count, 256), 256, 0, count, ctx->natms, recvbuf_dev, ctx->posq.p, ctx->vx.p, ctx->vy.p, ctx->vz.p,
         ctx->fx.p, ctx->fy.p, ctx->fz.p, ctx->ltg.p, ctx->lsite.p, ctx->ixyz.p);
  ctx->natms += count; ctx->nlast = ctx->natms;
  return 0;
}

int dlpgpu_dev_relocate_end(dlpgpu_ctx* ctx, int* natms_now) {
  if (ctx) dlp_hostio_ints_stale(ctx);
  if (!ctx) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  if (ctx->nsites < 1) return dlp_fail(ctx, DLPGPU_ERR_STATE, "relocate_end: sites not set");
  int n = ctx->natms;
  CK(cudaMemsetAsync(ctx->status.p, 0, 8 * sizeof(int), s));
  if (n > 0) {
    LAUNCH(ctx, k_count_nonzero, cdiv(n, 256), 256, 0, n, ctx->ixyz.p, ctx->status.p);
    LAUNCH(ctx, k_assign_sites, cdiv(n, 256), 256, 0, 0, n, ctx->lsite.p, ctx->type_site.p, ctx->charge_site.p, ctx->freeze_site.p,
           ctx->posq.p, ctx->ltype.p, ctx->lfrzn.p);
  }
  int st[8];
  CK(cudaMemcpyAsync(st, ctx->status.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  if (natms_now) *natms_now = n;
  if (st[2] != 0)   // deport_data.F90:3056-3058: an atom still wants to leave after the six stages
    return dlp_fail(ctx, DLPGPU_ERR_LOST_ATOMS, "error 58: %d atoms moved further than one domain in a single relocation", st[2]);
  return 0;
}

}  // extern "C"

extern "C" {

int dlpgpu_dev_xchg_init(dlpgpu_ctx* ctx, int rank, int nranks, int cap_reloc_atoms, int cap_halo_atoms, unsigned char handle_out[DLPGPU_XCHG_BLOB]) {
  if (!ctx || rank < 0 || nranks < 1 || nranks > 32 || rank >= nranks || cap_reloc_atoms < 1 || cap_halo_atoms < 1 || !handle_out) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (ctx->xr) return dlp_fail(ctx, DLPGPU_ERR_STATE, "xchg_init: already initialised");
  const XLayout L = x_layout(nranks, cap_reloc_atoms, cap_halo_atoms);
  CK(cudaMalloc((void**)&ctx->xr, L.bytes));
  CK(cudaMemset(ctx->xr, 0, L.bytes));
  CK(cudaDeviceSynchronize());
  ctx->xr_rank = rank; ctx->xr_nranks = nranks; ctx->xr_cap_r = cap_reloc_atoms; ctx->xr_cap_h = cap_halo_atoms;
  std::memset(handle_out, 0, DLPGPU_XCHG_BLOB);
  if (nranks > 1) {
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->xr));
    std::memcpy(handle_out, &h, 64);
  }
  {
    const long long pid = (long long)getpid(), dev = ctx->device;
    const unsigned long long p0 = (unsigned long long)(size_t)ctx->xr;
    std::memcpy(handle_out + 64, &pid, 8); std::memcpy(handle_out + 72, &p0, 8); std::memcpy(handle_out + 80, &dev, 8);
  }
  ctx->peer_xr.assign(nranks, nullptr);
  ctx->peer_xr[rank] = ctx->xr;
  CK(ctx->dcnt.ensure(DC_WORDS, ctx->stream));
  CK(cudaMemset(ctx->dcnt.p, 0, DC_WORDS * sizeof(int)));
  if (!ctx->dc_pinned) CK(cudaHostAlloc((void**)&ctx->dc_pinned, DC_WORDS * sizeof(int), cudaHostAllocDefault));
  CK(ctx->gmax_out.ensure(32, ctx->stream));
  ctx->xr_ready = false;
  if (nranks == 1) {
    CK(ctx->peer_xr_dev.ensure(1, ctx->stream));
    CK(cudaMemcpy(ctx->peer_xr_dev.p, ctx->peer_xr.data(), sizeof(void*), cudaMemcpyHostToDevice));
    ctx->xr_ready = true;
  }
  return 0;
}

int dlpgpu_dev_set_rebuild_every(dlpgpu_ctx* ctx, int every) {
  if (!ctx || every < 0) return DLPGPU_ERR_ARG;
  ctx->rebuild_every = every; ctx->steps_since_rebuild = 0;
  return 0;
}

int dlpgpu_dev_xchg_last_ms(dlpgpu_ctx* ctx, double* ms) {
  if (!ctx || !ms) return DLPGPU_ERR_ARG;
  *ms = ctx->t_xchg;
  return 0;
}

int dlpgpu_dev_xchg_set_migration(dlpgpu_ctx* ctx, int scan_all) {
  if (!ctx || scan_all < 0 || scan_all > 1) return DLPGPU_ERR_ARG;
  ctx->xchg_scan_migration = scan_all;
  return 0;
}

int dlpgpu_dev_xchg_set_timeout(dlpgpu_ctx* ctx, double seconds) {
  if (!ctx || !(seconds > 0.0)) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  const unsigned long long ns = (unsigned long long)(seconds * 1.0e9);
  CK(cudaMemcpyToSymbol(g_wait_ns, &ns, sizeof ns));
  return 0;
}

int dlpgpu_dev_xchg_open(dlpgpu_ctx* ctx, const unsigned char* all_handles /* nranks x DLPGPU_XCHG_BLOB bytes */) {
  if (!ctx || !all_handles) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->xr) return dlp_fail(ctx, DLPGPU_ERR_STATE, "xchg_open: call xchg_init first");
  for (int r = 0; r < ctx->xr_nranks; ++r) {
    if (r == ctx->xr_rank) continue;
    const unsigned char* blob = all_handles + (size_t)DLPGPU_XCHG_BLOB * r;
    long long pid = 0, dev = 0;
    std::memcpy(&pid, blob + 64, 8); std::memcpy(&dev, blob + 80, 8);
    if (pid == (long long)getpid()) {
      CKRC(x_enable_peer(ctx, (int)dev));
      unsigned long long q = 0;
      std::memcpy(&q, blob + 72, 8);
      ctx->peer_xr[r] = (char*)(size_t)q;
      ctx->peer_xr_local.resize((size_t)ctx->xr_nranks, 0); ctx->peer_xr_local[r] = 1;
      continue;
    }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, blob, 64);
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ctx->peer_xr[r] = (char*)p;
  }
  CK(ctx->peer_xr_dev.ensure((size_t)ctx->xr_nranks, ctx->stream));
  CK(cudaMemcpy(ctx->peer_xr_dev.p, ctx->peer_xr.data(), (size_t)ctx->xr_nranks * sizeof(void*), cudaMemcpyHostToDevice));
  ctx->xr_ready = true;
  return 0;
}

// vnl_check + gmax (neighbours.F90:123-182) with the reduction done by the GPUs themselves
int dlpgpu_dev_xchg_gmax(dlpgpu_ctx* ctx, unsigned long long seq, double* tol) {
  if (!ctx || !tol) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->xr_ready) return dlp_fail(ctx, DLPGPU_ERR_STATE, "xchg_gmax: exchange region not ready");
  if (!ctx->have_bg) return dlp_fail(ctx, DLPGPU_ERR_STATE, "vnl_check: no checkpoint");
  cudaStream_t s = ctx->stream;
  double rc[9];
  h_invert(ctx->cell, rc);
  if (!ctx->tol_fresh) {
    CK(cudaMemsetAsync(ctx->tol_bits.p, 0, sizeof(unsigned long long), s));
    if (ctx->natms > 0)
      LAUNCH(ctx, k_vnl_tol, cdiv(ctx->natms, 256), 256, 0, ctx->natms, ctx->imcon, mat(ctx->cell), mat(rc), ctx->posq.p, ctx->xbg.p,
             ctx->ybg.p, ctx->zbg.p, ctx->tol_bits.p);
  }
  ctx->tol_fresh = false;
  const XLayout L = x_layout(ctx->xr_nranks, ctx->xr_cap_r, ctx->xr_cap_h);
  if (!ctx->gm_pinned) {
    CK(cudaHostAlloc((void**)&ctx->gm_pinned, 32 * sizeof(unsigned long long), cudaHostAllocMapped));
    std::memset(ctx->gm_pinned, 0xff, 32 * sizeof(unsigned long long));
    CK(cudaHostGetDevicePointer((void**)&ctx->gm_pinned_dev, ctx->gm_pinned, 0));
  }
  LAUNCH(ctx, k_x_gmax, 1, 32, 0, ctx->xr_rank, ctx->xr_nranks, seq, ctx->tol_bits.p, ctx->out_dev.p, ctx->peer_xr_dev.p, L.off_gm,
         ctx->gmax_out.p, ctx->dcnt.p, ctx->gm_pinned_dev);
  {   // wait for the kernel's flag in pinned memory; fall back to the stream if it does not show up (a failed launch)
    volatile unsigned long long* flag = ctx->gm_pinned + 18;
    const auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    while (*flag != seq) {
      if ((++spins & 0xfff) == 0) {
        if (cudaStreamQuery(s) != cudaErrorNotReady) break;          // the stream drained (or failed): the flag is final
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 600.0) break;
      }
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (*flag != seq) { CK(cudaStreamSynchronize(s)); CK(cudaGetLastError()); }
    if (*flag != seq) return dlp_fail(ctx, DLPGPU_ERR_STATE, "xchg_gmax: the gmax kernel did not report (sequence %llu)", seq);
  }
  int err = 0;
  std::memcpy(&err, ctx->gm_pinned + 17, sizeof err);
  if (err & 4) return dlp_fail(ctx, DLPGPU_ERR_STATE, "xchg_gmax: timed out waiting for a peer (ranks out of lock-step)");
  double r;
  std::memcpy(&r, ctx->gm_pinned, sizeof r);
  *tol = r;
  std::memcpy(ctx->gsum_prev, ctx->gm_pinned + 1, 16 * sizeof(double));   // all-reduced sums of the previous force call
  return 0;
}

// relocate_particles (deport_data.F90:2870-3202) + set_halo_particles (halo.F90:153-355) + vnl_set_check in one enqueue
int dlpgpu_dev_xchg_rebuild(dlpgpu_ctx* ctx, const int neigh[6], unsigned long long seq, int* natms_out, int* nlast_out) {
  if (!ctx || !neigh) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->xr_ready) return dlp_fail(ctx, DLPGPU_ERR_STATE, "xchg_rebuild: exchange region not ready");
  dlp_hostio_ints_stale(ctx);
  if (ctx->nsites < 1) return dlp_fail(ctx, DLPGPU_ERR_STATE, "xchg_rebuild: sites not set");
  cudaStream_t s = ctx->stream;
  const int nr = ctx->xr_nranks, cap_r = ctx->xr_cap_r, cap_h = ctx->xr_cap_h;
  const XLayout L = x_layout(nr, cap_r, cap_h);
  const bool multi = ctx->nx * ctx->ny * ctx->nz > 1;
  const int mdirs[6] = {-1, 1, -2, 2, -3, 3};
  const int ub_total = ctx->natms + (multi ? 6 * cap_r : 0) + 6 * cap_h + 16;
  CKRC(dlp_ensure_atoms(ctx, ub_total));
  const int capacity = ctx->capacity;
  CK(ctx->flag.ensure((size_t)ub_total + 2, s)); CK(ctx->scan_out.ensure((size_t)ub_total + 2, s));
  CK(ctx->hole_pos.ensure((size_t)cap_r + 1, s));
  double rc[9];
  h_invert(ctx->cell, rc);
  int* dc = ctx->dcnt.p;
  // The counts travel through PAGE-LOCKED memory.  A cudaMemcpyAsync to or from pageable memory waits inside the driver for
  // the stream's earlier work; when the ranks are threads of one process that wait can hold up the other threads' launches,
  // and this stream's receive kernels are waiting for exactly those (seen on B200 as a time-out of the whole exchange).
  int* h_dc = ctx->dc_pinned;
  const size_t dc_bytes = DC_WORDS * sizeof(int);
  std::memset(h_dc, 0, dc_bytes);
  h_dc[DC_NATMS] = ctx->natms; h_dc[DC_NLAST] = ctx->natms;
  CK(cudaMemcpyAsync(dc, h_dc, dc_bytes, cudaMemcpyHostToDevice, s));
  ctx->halo_valid = false; ctx->list_valid = false;
  ctx->tol_fresh = false; ctx->pub_fresh = false;
  int nub = ctx->natms;   // host-side upper bound of the live natms / nlast
  CK(ctx->scan_tmp.ensure((size_t)cdiv(ub_total, XB_N) + 64, s));   // per-block selection counts of a stage
  if (!ctx->ev_x[0]) { cudaEventCreate(&ctx->ev_x[0]); cudaEventCreate(&ctx->ev_x[1]); }
  cudaEventRecord(ctx->ev_x[0], s);
  XAtoms A{ctx->posq.p, ctx->vx.p, ctx->vy.p, ctx->vz.p, ctx->fx.p, ctx->fy.p, ctx->fz.p, ctx->ltg.p, ctx->lsite.p, ctx->ixyz.p,
           ctx->org_rank.p, ctx->org_idx.p, ctx->org_wrap.p};
  auto hdr_of = [&](int r, int stage) { return reinterpret_cast<XHdr*>(ctx->peer_xr[r] + L.off_hdr) + stage; };
  auto rbuf_of = [&](int r, int q) { return reinterpret_cast<double*>(ctx->peer_xr[r] + L.off_rbuf) + (size_t)q * cap_r * 12; };
  auto hbuf_of = [&](int r, int q) { return reinterpret_cast<double*>(ctx->peer_xr[r] + L.off_hbuf) + (size_t)q * cap_h * DLP_HALO_W; };
  const int me = ctx->xr_rank;
  // ---- relocate_particles
  if (!multi) {
    if (ctx->natms > 0 && ctx->imcon != 0)
      LAUNCH(ctx, k_pbcshift, cdiv(ctx->natms, 256), 256, 0, ctx->natms, ctx->imcon, mat(ctx->cell), mat(rc), ctx->posq.p);   // deport_data.F90:3190
  } else {
    DomI D{ctx->nx, ctx->ny, ctx->nz, ctx->idx, ctx->idy, ctx->idz};
    if (nub > 0) LAUNCH(ctx, k_x_reloc_tag, cdiv(nub, 256), 256, 0, dc, mat(rc), D, ctx->posq.p, ctx->ixyz.p);
    if (ctx->xchg_scan_migration) {   // dlpgpu_dev_xchg_set_migration(ctx, 1): every stage scans all atoms (the first implementation)
      for (int q = 0; q < 6; ++q) {
        const Dir d = dir_settings(ctx, mdirs[q]);
        const int dst = neigh[q];
        if (dst < 0 || dst >= nr) return DLPGPU_ERR_ARG;
        const int nb = std::max(1, cdiv(nub, XB_N));
        LAUNCH(ctx, k_x_count<0>, nb, XB_T, 0, dc, d, ctx->ixyz.p, ctx->scan_tmp.p);
        LAUNCH(ctx, k_x_pack<0>, nb, XB_T, 0, dc, d, cap_r, q, ctx->scan_tmp.p, A, 0, rbuf_of(dst, q), ctx->hole_pos.p, ctx->flag.p, ctx->scan_out.p,
               hdr_of(dst, q), seq);
        LAUNCH(ctx, k_x_reloc_restack, cdiv(cap_r, 256), 256, 0, dc, cap_r, ctx->flag.p, ctx->scan_out.p, ctx->hole_pos.p, A);
        LAUNCH(ctx, k_x_recv<0>, cdiv(cap_r, 256), 256, 0, hdr_of(me, q), seq, rbuf_of(me, q), capacity, q, dc, A);
        nub = std::min(nub + cap_r, capacity);
      }
    } else {   // the movers compacted once, one single-block kernel per stage (see k_x_reloc_stage)
      const int capM = 6 * cap_r;
      CK(ctx->movers.ensure((size_t)capM + 1, s)); CK(ctx->hole_pos.ensure((size_t)cap_r + 1, s));
      const int nb = std::max(1, cdiv(nub, XB_N));
      LAUNCH(ctx, k_x_mov_count, nb, XB_T, 0, dc, ctx->ixyz.p, ctx->scan_tmp.p);
      LAUNCH(ctx, k_x_mov_pack, nb, XB_T, 0, dc, capM, ctx->scan_tmp.p, ctx->ixyz.p, ctx->movers.p);
      for (int q = 0; q < 6; ++q) {
        const Dir d = dir_settings(ctx, mdirs[q]);
        const int dst = neigh[q];
        if (dst < 0 || dst >= nr) return DLPGPU_ERR_ARG;
        LAUNCH(ctx, k_x_reloc_stage, 1, XM_T, 0, dc, d, cap_r, q, ctx->movers.p, capM, ctx->hole_pos.p, A, rbuf_of(dst, q), hdr_of(dst, q),
               hdr_of(me, q), rbuf_of(me, q), capacity, seq);
        nub = std::min(nub + cap_r, capacity);
      }
    }
    LAUNCH(ctx, k_x_reloc_end, cdiv(nub, 256), 256, 0, dc, ctx->ixyz.p, ctx->lsite.p, ctx->type_site.p, ctx->charge_site.p,
           ctx->freeze_site.p, ctx->posq.p, ctx->ltype.p, ctx->lfrzn.p);
  }
  // ---- set_halo_particles: thresholds of halo.F90:190-249 (see dlpgpu_dev_halo_begin)
  {
    double cut = ctx->rx + 1.0e-6;
    double w3[3];
    h_widths(ctx->cell, w3);
    double nxr = (double)ctx->nx, nyr = (double)ctx->ny, nzr = (double)ctx->nz;
    int nlx = (int)(w3[0] / (cut * nxr)), nly = (int)(w3[1] / (cut * nyr)), nlz = (int)(w3[2] / (cut * nzr));
    if (nlx * nly * nlz == 0) return dlp_fail(ctx, DLPGPU_ERR_LINK_CELLS, "error 307: domain narrower than cutoff_extended");
    double xdc = (double)(nlx * ctx->nx), ydc = (double)(nly * ctx->ny), zdc = (double)(nlz * ctx->nz);
    double cwx = 1.0 / xdc, cwy = 1.0 / ydc, cwz = 1.0 / zdc;
    double ecwx = std::max(cwx, ctx->ecw[0]), ecwy = std::max(cwy, ctx->ecw[1]), ecwz = std::max(cwz, ctx->ecw[2]);
    double nx_recip = 1.0 / nxr, ny_recip = 1.0 / nyr, nz_recip = 1.0 / nzr;
    const double zero_plus = DBL_MIN;
    HaloThr t;
    t.ecwx = std::nextafter((-0.5 + ecwx) + (double)ctx->idx * nx_recip, DBL_MAX) + zero_plus;
    t.ecwy = std::nextafter((-0.5 + ecwy) + (double)ctx->idy * ny_recip, DBL_MAX) + zero_plus;
    t.ecwz = std::nextafter((-0.5 + ecwz) + (double)ctx->idz * nz_recip, DBL_MAX) + zero_plus;
    t.cwx = std::nextafter((-0.5 - cwx) + (double)(ctx->idx + 1) * nx_recip, -DBL_MAX) - zero_plus - (nlx == 1 ? cwx * 1.0e-10 : 0.0);
    t.cwy = std::nextafter((-0.5 - cwy) + (double)(ctx->idy + 1) * ny_recip, -DBL_MAX) - zero_plus - (nly == 1 ? cwy * 1.0e-10 : 0.0);
    t.cwz = std::nextafter((-0.5 - cwz) + (double)(ctx->idz + 1) * nz_recip, -DBL_MAX) - zero_plus - (nlz == 1 ? cwz * 1.0e-10 : 0.0);
    if (nub > 0)
      LAUNCH(ctx, k_x_halo_tag, cdiv(nub, 256), 256, 0, dc, mat(rc), t, ctx->posq.p, ctx->ixyz.p, ctx->p2p_rank, ctx->org_rank.p, ctx->org_idx.p,
             ctx->org_wrap.p);
  }
  for (int q = 0; q < 6; ++q) {
    const Dir d = dir_settings(ctx, mdirs[q]);
    const int dst = neigh[q];
    if (dst < 0 || dst >= nr) return DLPGPU_ERR_ARG;
    HaloStage& st = ctx->stage[q];
    st.lwrap = d.lwrap != 0; st.shift[0] = d.xadd; st.shift[1] = d.yadd; st.shift[2] = d.zadd;
    CK(st.idx.ensure((size_t)cap_h + 1, s));
    int wrap_add = 0;
    if (d.lwrap) wrap_add = ((mdirs[q] < 0) ? +1 : -1) * (d.kx ? 1 : d.ky ? 3 : 9);
    const int nb = std::max(1, cdiv(nub, XB_N));
    LAUNCH(ctx, k_x_count<1>, nb, XB_T, 0, dc, d, ctx->ixyz.p, ctx->scan_tmp.p);
    LAUNCH(ctx, k_x_pack<1>, nb, XB_T, 0, dc, d, cap_h, q, ctx->scan_tmp.p, A, wrap_add, hbuf_of(dst, q), st.idx.p, nullptr, nullptr,
           hdr_of(dst, 6 + q), seq);
    LAUNCH(ctx, k_x_recv<1>, cdiv(cap_h, 256), 256, 0, hdr_of(me, 6 + q), seq, hbuf_of(me, q), capacity, q, dc, A);
    nub = std::min(nub + cap_h, capacity);
  }
  if (nub > 0)
    LAUNCH(ctx, k_x_halo_end, cdiv(nub, 256), 256, 0, dc, ctx->lsite.p, ctx->type_site.p, ctx->charge_site.p, ctx->freeze_site.p, ctx->posq.p,
           ctx->ltype.p, ctx->lfrzn.p, ctx->xbg.p, ctx->ybg.p, ctx->zbg.p);
  cudaEventRecord(ctx->ev_x[1], s);
  CK(cudaMemcpyAsync(h_dc, dc, dc_bytes, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  { float ms = 0.f; if (cudaEventElapsedTime(&ms, ctx->ev_x[0], ctx->ev_x[1]) == cudaSuccess) ctx->t_xchg = ms; }
  ctx->natms = h_dc[DC_NATMS]; ctx->nlast = h_dc[DC_NLAST];
  if (natms_out) *natms_out = ctx->natms;
  if (nlast_out) *nlast_out = ctx->nlast;
  if (h_dc[DC_ERR] & 4)
    return dlp_fail(ctx, DLPGPU_ERR_STATE, "xchg_rebuild: timed out waiting for a peer (ranks out of lock-step); rank %d seq %llu, migration "
                    "stages sent %d %d %d %d %d %d received %d %d %d %d %d %d, halo stages sent %d %d %d %d %d %d received %d %d %d %d %d %d",
                    me, seq, h_dc[DC_RSENT], h_dc[DC_RSENT + 1], h_dc[DC_RSENT + 2], h_dc[DC_RSENT + 3], h_dc[DC_RSENT + 4], h_dc[DC_RSENT + 5],
                    h_dc[DC_RRECV], h_dc[DC_RRECV + 1], h_dc[DC_RRECV + 2], h_dc[DC_RRECV + 3], h_dc[DC_RRECV + 4], h_dc[DC_RRECV + 5],
                    h_dc[DC_HSENT], h_dc[DC_HSENT + 1], h_dc[DC_HSENT + 2], h_dc[DC_HSENT + 3], h_dc[DC_HSENT + 4], h_dc[DC_HSENT + 5],
                    h_dc[DC_HRECV], h_dc[DC_HRECV + 1], h_dc[DC_HRECV + 2], h_dc[DC_HRECV + 3], h_dc[DC_HRECV + 4], h_dc[DC_HRECV + 5]);
  if (h_dc[DC_ERR] & 1) return dlp_fail(ctx, DLPGPU_ERR_BUFFER, "error 43: outgoing migration buffer too small (capacity %d atoms per stage)", cap_r);
  if (h_dc[DC_ERR] & 2) return dlp_fail(ctx, DLPGPU_ERR_BUFFER, "error 54: outgoing halo buffer too small (capacity %d atoms per stage)", cap_h);
  if (h_dc[DC_ERR] & 8) return dlp_fail(ctx, DLPGPU_ERR_BUFFER, "xchg_rebuild: atom arrays full (%d)", capacity);
  if (h_dc[DC_LOST] != 0)
    return dlp_fail(ctx, DLPGPU_ERR_LOST_ATOMS, "error 58: %d atoms moved further than one domain in a single relocation", h_dc[DC_LOST]);
  for (int q = 0; q < 6; ++q) {
    ctx->stage[q].count = h_dc[DC_HSENT + q]; ctx->stage[q].recv_off = h_dc[DC_HOFF + q]; ctx->stage[q].recv_count = h_dc[DC_HRECV + q];
  }
  ctx->have_bg = true;
  ctx->halo_valid = true;
  return 0;
}

// velocity-Verlet stage 1 + vnl_check displacement maximum + publish in one pass (see k_vv1_fused); called by dlpgpu_dev_vv
int dlp_vv1_fused(dlpgpu_ctx* ctx, double dt) {
  cudaStream_t s = ctx->stream;
  const bool want_tol = ctx->have_bg;
  const bool want_pub = ctx->pub[0] != nullptr && ctx->natms <= ctx->pub_cap;
  double rc[9];
  h_invert(ctx->cell, rc);
  if (want_tol) CK(cudaMemsetAsync(ctx->tol_bits.p, 0, sizeof(unsigned long long), s));
  if (ctx->natms > 0)
    LAUNCH(ctx, k_vv1_fused, cdiv(ctx->natms, 256), 256, 0, ctx->natms, dt, ctx->imcon, mat(ctx->cell), mat(rc), ctx->lsite.p, ctx->weight_site.p,
           ctx->posq.p, ctx->vx.p, ctx->vy.p, ctx->vz.p, ctx->fx.p, ctx->fy.p, ctx->fz.p, ctx->xbg.p, ctx->ybg.p, ctx->zbg.p,
           want_tol ? ctx->tol_bits.p : nullptr, want_pub ? ctx->pub[ctx->pub_parity ^ 1] : nullptr);
  ctx->tol_fresh = want_tol;
  ctx->pub_fresh = want_pub;
  return 0;
}

// One MD step of the native driver around the path (md_vv, drivers.F90:1910-2290) enqueued from C, so that the device does
// not wait for an interpreter between the gmax decision and the force kernels: velocity-Verlet stage 1 (+ displacement
// maximum + publish), gmax through the peers' mailboxes (the step's only host synchronisation), then either
// relocate_particles + set_halo_particles + link_cell_pairs (neighbours.F90:182 says rebuild) or the one-kernel halo
// refresh, two_body_forces without waiting for its sums, velocity-Verlet stage 2.  The sums of the PREVIOUS step's force
// call are handed back (they completed before this step's synchronisation); *have_prev says whether there were any.
int dlpgpu_dev_md_step(dlpgpu_ctx* ctx, const int neigh[6], double dt, unsigned long long gseq, unsigned long long rseq, int* rebuilt,
                       double out_prev[16], int* have_prev, double* list_ms) {
  if (!ctx || !neigh || !rebuilt || !out_prev || !have_prev) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->xr_ready) return dlp_fail(ctx, DLPGPU_ERR_STATE, "md_step: exchange region not ready");
  if (ctx->nx * ctx->ny * ctx->nz > 1 && !(ctx->pub[0] && ctx->p2p_ready))
    return dlp_fail(ctx, DLPGPU_ERR_STATE, "md_step: several domains need the peer-memory halo refresh (p2p_init / p2p_open)");
  CKRC(dlpgpu_dev_vv(ctx, 1, dt));
  if (ctx->pub[0]) CKRC(dlpgpu_dev_publish(ctx));
  double tol = 0.0;
  CKRC(dlpgpu_dev_xchg_gmax(ctx, gseq, &tol));
  *have_prev = 0;
  if (ctx->res_pending) {
    CKRC(dlpgpu_dev_fetch_results(ctx, out_prev));       // this rank's partial sums (and the timing bookkeeping) ...
    std::memcpy(out_prev, ctx->gsum_prev, 16 * sizeof(double));   // ... replaced by the gsum over the ranks that rode on the gmax message
    *have_prev = 1;
  }
  const double half_minus = 0.499999999999999944488848768742172978818416595458984375;
  *rebuilt = (tol >= half_minus * ctx->padding) ? 1 : 0;                       // neighbours.F90:182
  if (ctx->rebuild_every > 0 && ++ctx->steps_since_rebuild >= ctx->rebuild_every) *rebuilt = 1;   // dlpgpu_dev_set_rebuild_every
  if (*rebuilt) ctx->steps_since_rebuild = 0;
  if (list_ms) *list_ms = 0.0;
  if (*rebuilt) {
    CKRC(dlpgpu_dev_xchg_rebuild(ctx, neigh, rseq, nullptr, nullptr));
    int ibig = 0;
    CKRC(dlp_build_lists(ctx, 0, &ibig));
    if (list_ms) *list_ms = ctx->t_list;
  } else if (ctx->pub[0]) {
    CKRC(dlpgpu_dev_refresh_pull(ctx));
  } else {
    CKRC(dlpgpu_dev_refresh_serial(ctx));
  }
  CKRC(dlp_two_body(ctx, 1, nullptr));
  return dlpgpu_dev_vv(ctx, 2, dt);
}

}  // extern "C"

// Function-level lazy loading (the CUDA 12 default) loads a kernel at its first launch, and that load can synchronise the
// whole context.  When several ranks drive ONE GPU from threads of one process, a peer's receive / gmax kernel may be
// spinning on the device at that moment, waiting for a message this rank can only send after the load: a deadlock (seen on
// B200; CUDA_MODULE_LOADING=EAGER cures it).  dlpgpu_create therefore loads every kernel of the library up front.
int dlp_preload_halo() {
  const void* ks[] = {(const void*)k_vnl_tol, (const void*)k_vv1_fused, (const void*)k_halo_tag, (const void*)k_halo_flag, (const void*)k_halo_pack,
                      (const void*)k_halo_unpack, (const void*)k_assign_sites, (const void*)k_refresh_pack, (const void*)k_refresh_unpack,
                      (const void*)k_publish, (const void*)k_refresh_pull, (const void*)k_pbcshift, (const void*)k_reloc_tag,
                      (const void*)k_reloc_flag, (const void*)k_reloc_pack, (const void*)k_reloc_restack, (const void*)k_reloc_unpack,
                      (const void*)k_count_nonzero, (const void*)k_x_reloc_tag, (const void*)k_x_count<0>, (const void*)k_x_count<1>,
                      (const void*)k_x_pack<0>, (const void*)k_x_pack<1>, (const void*)k_x_reloc_restack, (const void*)k_x_recv<0>,
                      (const void*)k_x_recv<1>, (const void*)k_x_reloc_end, (const void*)k_x_halo_tag, (const void*)k_x_halo_end,
                      (const void*)k_x_gmax, (const void*)k_x_mov_count, (const void*)k_x_mov_pack, (const void*)k_x_reloc_stage};
  cudaFuncAttributes a;
  for (const void* k : ks) if (cudaFuncGetAttributes(&a, k) != cudaSuccess) cudaGetLastError();
  return 0;
}
