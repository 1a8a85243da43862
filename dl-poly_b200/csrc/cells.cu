// cells.cu -- link-cell binning, stable cell sort and Verlet neighbour-list kernels.
//
// Follows neighbours.F90::link_cell_pairs (:356-1306).  Every floating-point expression that decides an integer
// (cell index, list membership) is evaluated with the reference's operation order and WITHOUT fused multiply-add:
// this translation unit is compiled with -fmad=false (see __graft_entry__.build), so `a*b + c*d` rounds each product.
//
// Device layout produced by a build ("sorted" = link-cell order, the reference's at_list/xxt/yyt/zzt, :803-823):
//   which_cell[i]  cell id of local index i (0 = residual halo)         lct_start[c]  first sorted slot of cell c (0-based)
//   at_list[s]     local index held by sorted slot s                    cell_s[s]     cell id of slot s
//   posq_s[s]      {x,y,z,chge} of slot s (refreshed every force call)  loc_slot[t]   slot of the t-th LOCAL atom
//   nbr[t][k]      full neighbour row of local atom t: slot | flags     xnbr[t][k]    excluded partners (ewald_excl_forces)
//   ref_list       optional reference-format half list (-3:max_list,1:natms), bit-exact incl. row order
#include "common.cuh"

namespace {

// ---------------------------------------------------------------- host geometry (neighbours.F90:401-601)
void h_dcell(const double* aaa0, double* b /*1..10*/) {   // numerics.F90:1344-1446
  const double* aaa = aaa0 - 1;
  b[1] = std::sqrt(aaa[1] * aaa[1] + aaa[2] * aaa[2] + aaa[3] * aaa[3]);
  b[2] = std::sqrt(aaa[4] * aaa[4] + aaa[5] * aaa[5] + aaa[6] * aaa[6]);
  b[3] = std::sqrt(aaa[7] * aaa[7] + aaa[8] * aaa[8] + aaa[9] * aaa[9]);
  double axb1 = aaa[2] * aaa[6] - aaa[3] * aaa[5], axb2 = aaa[3] * aaa[4] - aaa[1] * aaa[6], axb3 = aaa[1] * aaa[5] - aaa[2] * aaa[4];
  double bxc1 = aaa[5] * aaa[9] - aaa[6] * aaa[8], bxc2 = aaa[6] * aaa[7] - aaa[4] * aaa[9], bxc3 = aaa[4] * aaa[8] - aaa[5] * aaa[7];
  double cxa1 = aaa[8] * aaa[3] - aaa[9] * aaa[2], cxa2 = aaa[9] * aaa[1] - aaa[7] * aaa[3], cxa3 = aaa[7] * aaa[2] - aaa[8] * aaa[1];
  b[10] = std::fabs(aaa[1] * bxc1 + aaa[2] * bxc2 + aaa[3] * bxc3);
  double d[4], x[4], y[4];
  d[1] = b[10] / std::sqrt(bxc1 * bxc1 + bxc2 * bxc2 + bxc3 * bxc3);
  d[2] = b[10] / std::sqrt(cxa1 * cxa1 + cxa2 * cxa2 + cxa3 * cxa3);
  d[3] = b[10] / std::sqrt(axb1 * axb1 + axb2 * axb2 + axb3 * axb3);
  x[1] = std::fabs(aaa[1]) / b[1]; y[1] = std::fabs(aaa[2]) / b[1];
  x[2] = std::fabs(aaa[4]) / b[2]; y[2] = std::fabs(aaa[5]) / b[2];
  x[3] = std::fabs(aaa[7]) / b[3]; y[3] = std::fabs(aaa[8]) / b[3];
  if (x[1] >= x[2] && x[1] >= x[3]) { b[7] = d[1]; if (y[2] >= y[3]) { b[8] = d[2]; b[9] = d[3]; } else { b[8] = d[3]; b[9] = d[2]; } }
  else if (x[2] >= x[1] && x[2] >= x[3]) { b[7] = d[2]; if (y[1] >= y[3]) { b[8] = d[1]; b[9] = d[3]; } else { b[8] = d[3]; b[9] = d[1]; } }
  else { b[7] = d[3]; if (y[1] >= y[2]) { b[8] = d[1]; b[9] = d[2]; } else { b[8] = d[2]; b[9] = d[1]; } }
}
void h_invert(const double* a0, double* b0) {   // numerics.F90:1448-1509
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

int h_geometry(dlpgpu_ctx* ctx) {
  LCGeom& g = ctx->g;
  double celprp[11];
  h_dcell(ctx->cell, celprp);
  double det = std::min(celprp[7], std::min(celprp[8], celprp[9]));
  if (ctx->rx >= det / 2.0)   // :409-412
    return dlp_fail(ctx, DLPGPU_ERR_CUTOFF_HALF_CELL, "error 95: cutoff_extended %.6f >= half the minimum cell width %.6f", ctx->rx, det / 2.0);
  double cut = ctx->rx + 1.0e-6;   // smalldr, :416
  g.rcsq = ctx->rx * ctx->rx;      // :417
  double nx_recip = 1.0 / (double)ctx->nx, ny_recip = 1.0 / (double)ctx->ny, nz_recip = 1.0 / (double)ctx->nz;
  double dispx = nx_recip * celprp[7] / cut, dispy = ny_recip * celprp[8] / cut, dispz = nz_recip * celprp[9] / cut;
  int nlx = (int)dispx, nly = (int)dispy, nlz = (int)dispz;
  if (nlx * nly * nlz == 0) return dlp_fail(ctx, DLPGPU_ERR_LINK_CELLS, "error 307: link cell algorithm violation (domain narrower than cutoff_extended)");
  int nlp = 1;   // :441-454
  double nlr2 = (double)ctx->natms;
  det = nlr2 / (double)(nlx * nly * nlz);
  while (det > ctx->pdplnc) {
    nlp = nlp + 1;
    double rsq = (double)nlp;
    nlx = (int)(dispx * rsq); nly = (int)(dispy * rsq); nlz = (int)(dispz * rsq);
    det = nlr2 / (double)(nlx * nly * nlz);
  }
  g.nlx = nlx; g.nly = nly; g.nlz = nlz; g.nlp = nlp;
  g.sx = nlx + 2 * nlp; g.sy = nly + 2 * nlp; g.sz = nlz + 2 * nlp;
  long long nc = (long long)g.sx * g.sy * g.sz;
  if (nc > 0x3fffffffLL) return dlp_fail(ctx, DLPGPU_ERR_ARG, "too many link cells");
  g.ncells = (int)nc;
  g.idx = ctx->idx; g.idy = ctx->idy; g.idz = ctx->idz;
  g.xdc = (double)(nlx * ctx->nx); g.ydc = (double)(nly * ctx->ny); g.zdc = (double)(nlz * ctx->nz);   // :545-547
  g.jx = nlp - nlx * ctx->idx; g.jy = nlp - nly * ctx->idy; g.jz = nlp - nlz * ctx->idz;               // :554-556
  h_invert(ctx->cell, g.rcell);
  g.nir_r2 = (nlp - 1) * (nlp - 1);
  // semi-ball stencil :490-540 in reference order
  ctx->h_nix.clear(); ctx->h_niy.clear(); ctx->h_niz.clear(); ctx->h_nir.clear();
  int nlp2 = nlp * nlp, nlp3 = (nlp - 1) * (nlp - 1);
  int W = 2 * nlp + 1;
  ctx->h_xb.assign((size_t)W * W, -1);
  for (int iz = 0; iz <= nlp; ++iz) {
    int iz1 = (iz > 0) ? (iz - 1) * (iz - 1) : 0;
    int jz = iz * iz;
    for (int iy = -nlp; iy <= nlp; ++iy) {
      if (iz == 0 && iy < 0) continue;
      int a = std::abs(iy);
      int iy1 = (a > 0) ? (a - 1) * (a - 1) : 0;
      int ll = iz1 + iy1;
      if (ll > nlp2) continue;
      int jy = jz + iy * iy;
      for (int ix = -nlp; ix <= nlp; ++ix) {
        if (iz == 0 && iy == 0 && ix < 0) continue;
        int b = std::abs(ix);
        int ix1 = (b > 0) ? (b - 1) * (b - 1) : 0;
        if (ll + ix1 > nlp2) continue;
        int jxx = jy + ix * ix;
        ctx->h_nix.push_back(ix); ctx->h_niy.push_back(iy); ctx->h_niz.push_back(iz);
        ctx->h_nir.push_back(jxx < nlp3 ? 1 : 0);
        // symmetric half extent of row (iy,iz) and of its mirror (-iy,-iz)
        int& e1 = ctx->h_xb[(size_t)(iz + nlp) * W + (iy + nlp)];
        int& e2 = ctx->h_xb[(size_t)(-iz + nlp) * W + (-iy + nlp)];
        e1 = std::max(e1, b);
        e2 = std::max(e2, b);
      }
    }
  }
  g.nsbcll = (int)ctx->h_nix.size();
  return 0;
}

// ---------------------------------------------------------------- binning (neighbours.F90:612-799)
__device__ __forceinline__ bool f_equal(double a, double b) { return fabs(a - b) < DBL_EPSILON; }   // numerics.F90:3888-3893

__global__ void k_cell_index(LCGeom g, int natms, int nlast, const double4* __restrict__ posq, int* __restrict__ which_cell,
                             int* __restrict__ lct_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlast) return;
  const double half_plus = 0.50000000000000011102230246251565404236316680908203125;   // Nearest(0.5,+1)
  double4 p = posq[i];
  double x = g.rcell[0] * p.x + g.rcell[3] * p.y + g.rcell[6] * p.z;
  double y = g.rcell[1] * p.x + g.rcell[4] * p.y + g.rcell[7] * p.z;
  double z = g.rcell[2] * p.x + g.rcell[5] * p.y + g.rcell[8] * p.z;
  const int nlp = g.nlp, nlx = g.nlx, nly = g.nly, nlz = g.nlz;
  const int nlx0e = nlp - 1, nly0e = nlp - 1, nlz0e = nlp - 1;
  const int nlx1s = nlx + nlp, nly1s = nly + nlp, nlz1s = nlz + nlp;
  const int nlx1e = nlx + 2 * nlp - 1, nly1e = nly + 2 * nlp - 1, nlz1e = nlz + 2 * nlp - 1;
  int ix, iy, iz, icell;
  if (i < natms) {
    ix = __double2int_rz(g.xdc * (x + 0.5)) + g.jx;
    iy = __double2int_rz(g.ydc * (y + 0.5)) + g.jy;
    iz = __double2int_rz(g.zdc * (z + 0.5)) + g.jz;
    ix = max(min(ix, nlx1s - 1), nlx0e + 1);
    iy = max(min(iy, nly1s - 1), nly0e + 1);
    iz = max(min(iz, nlz1s - 1), nlz0e + 1);
    icell = 1 + ix + g.sx * (iy + g.sy * iz);
  } else {
    double dpx, dpy, dpz;
    if (x > -half_plus) { dpx = g.xdc * (x + 0.5); ix = __double2int_rz(dpx) + g.jx; }
    else { dpx = g.xdc * fabs(x + 0.5); ix = -__double2int_rz(dpx) + g.jx - 1; }
    if (y > -half_plus) { dpy = g.ydc * (y + 0.5); iy = __double2int_rz(dpy) + g.jy; }
    else { dpy = g.ydc * fabs(y + 0.5); iy = -__double2int_rz(dpy) + g.jy - 1; }
    if (z > -half_plus) { dpz = g.zdc * (z + 0.5); iz = __double2int_rz(dpz) + g.jz; }
    else { dpz = g.zdc * fabs(z + 0.5); iz = -__double2int_rz(dpz) + g.jz - 1; }
    if (ix >= 0 && iy >= 0 && iz >= 0) {
      bool lx0 = (ix > nlx0e), lx1 = (ix < nlx1s), ly0 = (iy > nly0e), ly1 = (iy < nly1s), lz0 = (iz > nlz0e), lz1 = (iz < nlz1s);
      if ((lx0 && lx1) && (ly0 && ly1) && (lz0 && lz1)) {   // halo atom kicked into the domain: put on the border (:699-756)
        double xa = fabs(dpx - (double)(nlx * g.idx)), x1 = fabs(dpx - (double)(nlx * (g.idx + 1)));
        dpx = fmin(xa, x1);
        double ya = fabs(dpy - (double)(nly * g.idy)), y1 = fabs(dpy - (double)(nly * (g.idy + 1)));
        dpy = fmin(ya, y1);
        double za = fabs(dpz - (double)(nlz * g.idz)), z1 = fabs(dpz - (double)(nlz * (g.idz + 1)));
        dpz = fmin(za, z1);
        if (dpx <= dpy && dpx <= dpz) {
          ix = (xa < x1) ? nlx0e : nlx1s;
          if (f_equal(dpx, dpy)) iy = (ya < y1) ? nly0e : nly1s;
          if (f_equal(dpx, dpz)) iz = (za < z1) ? nlz0e : nlz1s;
        } else if (dpy <= dpx && dpy <= dpz) {
          iy = (ya < y1) ? nly0e : nly1s;
          if (f_equal(dpy, dpz)) iz = (za < z1) ? nlz0e : nlz1s;
        } else {
          iz = (za < z1) ? nlz0e : nlz1s;
        }
      }
      bool out = (ix < 0) || (ix > nlx1e) || (iy < 0) || (iy > nly1e) || (iz < 0) || (iz > nlz1e);
      icell = out ? 0 : 1 + ix + g.sx * (iy + g.sy * iz);
    } else {
      icell = 0;
    }
  }
  which_cell[i] = icell;
  atomicAdd(&lct_count[icell], 1);
}

__global__ void k_cell_scatter(int nlast, const int* __restrict__ which_cell, const int* __restrict__ lct_start,
                               int* __restrict__ lct_fill, int* __restrict__ at_tmp) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlast) return;
  int c = which_cell[i];
  int pos = lct_start[c] + atomicAdd(&lct_fill[c], 1);
  at_tmp[pos] = i;
}

// Stable order inside each cell = ascending local index, as the reference's sequential counting sort gives (:811-823).
// One warp per cell, rank by counting.
// Cell 0 (the residual halo, :775-789) can hold a large share of the halo, so its stable order comes from a scan instead.
__global__ void k_cell0_flag(int nlast, const int* __restrict__ which_cell, int* __restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > nlast) return;
  flag[i] = (i < nlast && which_cell[i] == 0) ? 1 : 0;
}
__global__ void k_cell0_place(int nlast, const int* __restrict__ which_cell, const int* __restrict__ rank, int* __restrict__ at_list,
                              int* __restrict__ cell_s) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nlast || which_cell[i] != 0) return;
  at_list[rank[i]] = i;      // lct_start[0] == 0
  cell_s[rank[i]] = 0;
}
__global__ void k_cell_order(int ncells_p1, const int* __restrict__ lct_start, const int* __restrict__ at_tmp,
                             int* __restrict__ at_list, int* __restrict__ cell_s) {
  int c = 1 + ((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  int lane = threadIdx.x & 31;
  if (c >= ncells_p1) return;
  int s0 = lct_start[c], n = lct_start[c + 1] - s0;
  for (int a = lane; a < n; a += 32) {
    int e = at_tmp[s0 + a];
    int r = 0;
    for (int b = 0; b < n; ++b) r += (at_tmp[s0 + b] < e);
    at_list[s0 + r] = e;
    cell_s[s0 + a] = c;
  }
}

__global__ void k_sorted_static(int nlast, int natms, const int* __restrict__ at_list, const int* __restrict__ ltype,
                                const int* __restrict__ ltg, const int* __restrict__ lfrzn, int* __restrict__ type_s,
                                int* __restrict__ gid_s, int* __restrict__ frz_s, int2* __restrict__ info_s, int* __restrict__ is_local) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nlast) return;
  if (s == nlast) { is_local[s] = 0; return; }   // zero pad for the scan's total slot
  int i = at_list[s];
  type_s[s] = ltype[i]; gid_s[s] = ltg[i]; frz_s[s] = lfrzn[i];
  info_s[s] = make_int2(ltg[i], (ltype[i] & 0xffff) | (lfrzn[i] > 0 ? (1 << 16) : 0) | (i >= natms ? (1 << 17) : 0));
  is_local[s] = (i < natms) ? 1 : 0;
}
__global__ void k_loc_slot(int nlast, const int* __restrict__ is_local, const int* __restrict__ rank, int* __restrict__ loc_slot) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nlast && is_local[s]) loc_slot[rank[s]] = s;
}
__global__ void k_gather_posq(int nlast, const int* __restrict__ at_list, const double4* __restrict__ posq, double4* __restrict__ posq_s) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s < nlast) posq_s[s] = posq[at_list[s]];
  else if (s == nlast) posq_s[s] = make_double4(1e15, 1e15, 1e15, 0.0);   // the sentinel partner of padded list rows
}

// ---------------------------------------------------------------- list kernels
// The pair kernels read rows in passes of 16 or 32 entries without bounds checks and gather one pass ahead: every row is
// padded from its length n up to roundup32(n) + 32 with a sentinel partner (sorted slot nlast: see k_gather_posq).
#define DLP_ROW_PAD 64
__device__ __forceinline__ void dlp_pad_row(unsigned* row, int n, unsigned sentinel, int lane, int nlanes) {
  const int end = ((n + 31) & ~31) + 32;
  for (int k = n + lane; k < end; k += nlanes) row[k] = sentinel;
}
__device__ __forceinline__ double pair_rsq(const double4& a, double xi, double yi, double zi) {
  // neighbours.F90:991-992  rsq = (xxt(jj)-x_i)**2 + (yyt(jj)-y_i)**2 + (zzt(jj)-z_i)**2   (left-to-right, no FMA)
  double dx = a.x - xi, dy = a.y - yi, dz = a.z - zi;
  return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}
__device__ __forceinline__ bool nir_of(const LCGeom& g, int cj, int ix, int iy, int iz) {
  if (g.nir_r2 <= 0) return false;
  int c = cj - 1;
  int jx = c % g.sx, jz = c / (g.sx * g.sy), jy = c / g.sx - g.sy * jz;
  int dx = jx - ix, dy = jy - iy, dz = jz - iz;
  return dx * dx + dy * dy + dz * dz < g.nir_r2;
}

// Reference-format half list.  One warp per local atom; candidates are taken 32 at a time in the reference's visiting
// order (stencil order kk, then ascending sorted slot), accepted ones are appended with ballot/popc prefix compaction,
// so row contents AND order equal the sequential algorithm's.
__global__ void k_list_ref(LCGeom g, int natms, int max_list, const int* __restrict__ loc_slot, const int* __restrict__ at_list,
                           const int* __restrict__ lct_start, const int* __restrict__ cell_s, const double4* __restrict__ posq_s,
                           const int* __restrict__ nix, const int* __restrict__ niy, const int* __restrict__ niz,
                           const int* __restrict__ nir, int* __restrict__ list, int* __restrict__ status) {
  int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (t >= natms) return;
  const int ii = loc_slot[t];
  const int i = at_list[ii];
  const int ic = cell_s[ii];
  const int ix = (ic - 1) % g.sx, iz = (ic - 1) / (g.sx * g.sy), iy = (ic - 1) / g.sx - g.sy * iz;
  const double4 pi = posq_s[ii];
  int* row = list + (size_t)i * (max_list + 4) + 3;   // row[k] == list(k, i)
  int cnt = 0;
  const int nlp = g.nlp;
  const int lo_x = nlp - 1, hi_x = g.nlx + nlp, lo_y = nlp - 1, hi_y = g.nly + nlp, lo_z = nlp - 1, hi_z = g.nlz + nlp;

  auto scan_range = [&](int s0, int s1, bool nirflag) {
    for (int b = s0; b < s1; b += 32) {
      int jj = b + lane;
      bool ok = false;
      if (jj < s1) {
        if (nirflag) ok = true;
        else ok = pair_rsq(posq_s[jj], pi.x, pi.y, pi.z) <= g.rcsq;
      }
      unsigned m = __ballot_sync(DLP_FULL, ok);
      if (ok) {
        int ll = cnt + __popc(m & ((1u << lane) - 1)) + 1;
        if (ll <= max_list) row[ll] = at_list[jj] + 1;
        else { atomicOr(&status[0], 1); atomicMax(&status[1], ll); }
      }
      cnt += __popc(m);
    }
  };
  // pass 1: positive semi-ball (:877-1027)
  for (int kk = 0; kk < g.nsbcll; ++kk) {
    int jx = ix + nix[kk], jy = iy + niy[kk], jz = iz + niz[kk];
    int jc = 1 + jx + g.sx * (jy + g.sy * jz);
    int s0 = (jc != ic) ? lct_start[jc] : ii + 1;
    scan_range(s0, lct_start[jc + 1], nir[kk] != 0);
  }
  // pass 2: negative semi-ball, halo cells only, border cells only (:1033-1184)
  bool border = (ix - lo_x <= nlp) || (hi_x - ix <= nlp) || (iy - lo_y <= nlp) || (hi_y - iy <= nlp) || (iz - lo_z <= nlp) || (hi_z - iz <= nlp);
  if (border) {
    for (int kk = 1; kk < g.nsbcll; ++kk) {
      int jx = ix - nix[kk], jy = iy - niy[kk], jz = iz - niz[kk];
      if ((jx <= lo_x) || (jx >= hi_x) || (jy <= lo_y) || (jy >= hi_y) || (jz <= lo_z) || (jz >= hi_z)) {
        int jc = 1 + jx + g.sx * (jy + g.sy * jz);
        scan_range(lct_start[jc], lct_start[jc + 1], nir[kk] != 0);
      }
    }
  }
  if (lane == 0) { row[0] = cnt; row[-1] = cnt; row[-2] = cnt; row[-3] = cnt; }
}

__device__ __forceinline__ bool excl_match(int n, int ind_top, const int* __restrict__ list1 /*1-based view*/) {
  // numerics.F90:1048-1098 match()
  if (ind_top < 1) return false;
  int ind_old = 1, ind_now = 1;
  for (;;) {
    int v = list1[ind_now];
    if (n == v) return true;
    else if (n > v) {
      if (ind_old == ind_top) return false;
      ind_old = ind_now;
      ind_now = (ind_old + ind_top + 1) / 2;
    } else {
      ind_now = (ind_old + ind_now) / 2;
      if (ind_now == ind_old) return false;
    }
  }
}

// Row partition of the reference-format list: frozen-frozen pairs, then excluded pairs, swapped to the row tail with
// the reference's backwards sweep (:1198-1251) -- one thread per row, sequential like the original, so the order matches.
__global__ void k_row_partition(int natms, int max_list, int megfrz, int lbook, int max_exclude, int excl_by_gid,
                                const int* __restrict__ lfrzn, const int* __restrict__ ltg, const int* __restrict__ excl,
                                int* __restrict__ list) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natms) return;
  int* row = list + (size_t)i * (max_list + 4) + 3;
  if (row[0] > max_list) return;   // overflowed row: the build fails with error 106 anyway
  if (megfrz > 1) {
    int l_end = row[0], m_end = l_end;
    if (lfrzn[i] > 0) {
      for (int kk = l_end; kk >= 1; --kk) {
        int j = row[kk];
        if (lfrzn[j - 1] > 0) {
          if (kk < m_end) { row[kk] = row[m_end]; row[m_end] = j; }
          m_end = m_end - 1;
        }
      }
    }
    row[-2] = row[0];
    row[0] = m_end;
  } else {
    row[-2] = row[0];
  }
  if (lbook) {
    int l_end = row[0], m_end = l_end;
    const int* ex = excl + (size_t)(excl_by_gid ? (ltg[i] - 1) : i) * (max_exclude + 1);
    int ii = ex[0];
    if (ii > 0) {
      for (int kk = l_end; kk >= 1; --kk) {
        int j = row[kk];
        int jj = ltg[j - 1];
        if (excl_match(jj, ii, ex)) {
          if (kk < m_end) { row[kk] = row[m_end]; row[m_end] = j; }
          m_end = m_end - 1;
        }
      }
    }
    row[-1] = row[0];
    row[0] = m_end;
    row[-3] = row[0];
  } else {
    row[-1] = row[0];
    row[-3] = row[0];
  }
}

// Device-internal lists.  FULL: every partner (local or halo) of local atom t within the Verlet radius -- the
// symmetrised reference list restricted to local primaries, so no force scatter is needed.  HALF additionally keeps the
// reference's pair ownership (local-local pairs once) for the Newton's-third-law kernel.
// Candidate cells are exactly the reference's (semi-ball and its mirror); x-runs of one (dy,dz) row are contiguous in
// the sorted arrays, which is what makes 32-wide candidate batches dense.
template <bool HALF>
__global__ void k_list_dev(LCGeom g, int natms, int pitch, int xpitch, int megfrz, int lbook, int max_exclude, int excl_by_gid,
                           const int* __restrict__ loc_slot, const int* __restrict__ at_list, const int* __restrict__ lct_start,
                           const int* __restrict__ cell_s, const double4* __restrict__ posq_s, const int* __restrict__ gid_s,
                           const int* __restrict__ frz_s, const int* __restrict__ type_s, const int* __restrict__ pair_k, int ntypes,
                           const int* __restrict__ xb, const int* __restrict__ excl,
                           unsigned* __restrict__ nbr, int* __restrict__ nnbr, unsigned* __restrict__ xnbr, int* __restrict__ nxnbr,
                           int* __restrict__ status, unsigned sentinel) {
  int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (t >= natms) return;
  const int ii = loc_slot[t];
  const int i = at_list[ii];
  const int ic = cell_s[ii];
  const int ix = (ic - 1) % g.sx, iz = (ic - 1) / (g.sx * g.sy), iy = (ic - 1) / g.sx - g.sy * iz;
  const double4 pi = posq_s[ii];
  const int gid_i = gid_s[ii];
  const int frz_i = (megfrz > 1) ? frz_s[ii] : 0;
  const int type_i = type_s[ii];
  const int* ex = lbook ? excl + (size_t)(excl_by_gid ? (gid_i - 1) : i) * (max_exclude + 1) : nullptr;
  const int nex = lbook ? ex[0] : 0;
  unsigned* row = nbr + (size_t)t * pitch;
  unsigned* xrow = xnbr + (size_t)t * (xpitch > 0 ? xpitch : 1);
  int cnt = 0, xcnt = 0;
  const int nlp = g.nlp, W = 2 * nlp + 1;
  for (int dz = -nlp; dz <= nlp; ++dz) {
    for (int dy = -nlp; dy <= nlp; ++dy) {
      int b = xb[(dz + nlp) * W + (dy + nlp)];
      if (b < 0) continue;
      int xlo = -b, xhi = b;
      int jy = iy + dy, jz = iz + dz;
      int c0 = 1 + (ix + xlo) + g.sx * (jy + g.sy * jz);
      int c1 = 1 + (ix + xhi) + g.sx * (jy + g.sy * jz);
      int s0 = lct_start[c0], s1 = lct_start[c1 + 1];
      for (int bb = s0; bb < s1; bb += 32) {
        int jj = bb + lane;
        bool ok = false, isx = false;
        unsigned entry = 0;
        if (jj < s1 && jj != ii) {
          int cj = cell_s[jj];
          bool acc = nir_of(g, cj, ix, iy, iz) || (pair_rsq(posq_s[jj], pi.x, pi.y, pi.z) <= g.rcsq);
          if (acc) {
            int jref = at_list[jj];
            bool halo = jref >= natms;
            int gj = gid_s[jj];
            bool keep = true;
            if (HALF && !halo) {
              // pair ownership of the reference: the atom whose cell sees the other in the positive semi-ball; same cell: lower slot
              int cdx = (cj - 1) % g.sx - ix, cdz = (cj - 1) / (g.sx * g.sy) - iz, cdy = (cj - 1) / g.sx - g.sy * ((cj - 1) / (g.sx * g.sy)) - iy;
              bool pos = (cdz > 0) || (cdz == 0 && (cdy > 0 || (cdy == 0 && (cdx > 0 || (cdx == 0 && jj > ii)))));
              keep = pos;
            }
            if (keep) {
              if (frz_i > 0 && frz_s[jj] > 0) keep = false;   // frozen-frozen pairs never reach the force loops (:1198-1225)
            }
            if (keep) {
              const int kc = pair_k ? pair_k[(type_i - 1) * ntypes + (type_s[jj] - 1)] + 1 : 0;
              entry = (unsigned)jj | ((unsigned)kc << DLP_K_SHIFT) | (halo ? DLP_F_HALO : 0u) | ((halo && gid_i < gj) ? DLP_F_ECNT : 0u);
              if (nex > 0 && excl_match(gj, nex, ex)) isx = true; else ok = true;
            }
          }
        }
        unsigned m = __ballot_sync(DLP_FULL, ok), mx = __ballot_sync(DLP_FULL, isx);
        if (ok) {
          int ll = cnt + __popc(m & ((1u << lane) - 1));
          if (ll < pitch - DLP_ROW_PAD) row[ll] = entry; else { atomicOr(&status[0], 1); atomicMax(&status[1], ll + 1); }
        }
        if (isx) {
          int ll = xcnt + __popc(mx & ((1u << lane) - 1));
          if (ll < xpitch) xrow[ll] = entry; else { atomicOr(&status[0], 2); }
        }
        cnt += __popc(m);
        xcnt += __popc(mx);
      }
    }
  }
  if (lane == 0) { nnbr[t] = cnt; nxnbr[t] = xcnt; }
  dlp_pad_row(row, min(cnt, pitch - DLP_ROW_PAD), sentinel, lane, 32);
}


// ---------------------------------------------------------------- half list, one warp per link cell
// The reference visits, for every local atom, the positive semi-ball of cells (:877-1027) and -- for border cells -- the
// halo cells of the negative semi-ball (:1033-1184).  All atoms of one link cell share those candidate cells, so one warp
// takes a cell: the candidate slots (contiguous x-runs of the cell-sorted arrays, flattened through a small per-warp run
// table) are loaded ONCE per cell, 32 at a time, and tested against each atom of the cell, whose coordinates sit in shared
// memory and are read as broadcasts.  Accepted partners are appended with ballot/popc prefix compaction.  Entry format:
// see DLP_J_MASK.  Row order is irrelevant to the force kernel (sums are order-tolerant); membership is exact.
#define LC_WARPS 8
#define LC_MAXRUN 96
struct LCRow { int dy, dz, b; };

// Bounding box of the atoms of every link cell (halo cells included), single precision, rounded outward: what the run table of
// k_list_cell trims its x-runs against.  An empty cell gets an inverted box (never within reach of anything).
__global__ void k_cell_boxes(int ncells, const int* __restrict__ lct_start, const double4* __restrict__ posq_s,
                             float4* __restrict__ box /* [2 * (ncells + 1)]: lo, hi */) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > ncells) return;
  float lx = 3.0e38f, ly = 3.0e38f, lz = 3.0e38f, hx = -3.0e38f, hy = -3.0e38f, hz = -3.0e38f;
  if (c >= 1) {
    for (int s = lct_start[c]; s < lct_start[c + 1]; ++s) {
      const double4 p = posq_s[s];
      lx = fminf(lx, __double2float_rd(p.x)); ly = fminf(ly, __double2float_rd(p.y)); lz = fminf(lz, __double2float_rd(p.z));
      hx = fmaxf(hx, __double2float_ru(p.x)); hy = fmaxf(hy, __double2float_ru(p.y)); hz = fmaxf(hz, __double2float_ru(p.z));
    }
  }
  box[2 * c] = make_float4(lx, ly, lz, 0.f);
  box[2 * c + 1] = make_float4(hx, hy, hz, 0.f);
}
// squared distance between two boxes (0 when they overlap; huge when one is inverted)
__device__ __forceinline__ float box_dist2(const float4& alo, const float4& ahi, const float4& blo, const float4& bhi) {
  const float dx = fmaxf(fmaxf(blo.x - ahi.x, alo.x - bhi.x), 0.f), dy = fmaxf(fmaxf(blo.y - ahi.y, alo.y - bhi.y), 0.f),
              dz = fmaxf(fmaxf(blo.z - ahi.z, alo.z - bhi.z), 0.f);
  return dx * dx + dy * dy + dz * dz;
}

// Run table of a cell's candidate slots: potential run p < nrows: positive row p; p >= nrows: negative row (p-nrows)/2, low /
// high x part (halo cells only, neighbours.F90:1033-1184).  run0[r] = first slot of run r, pre[r] = candidates before run r;
// returns the number of candidates.  One warp; run0 / pre are that warp's shared arrays.
// box != nullptr: every run is trimmed to the cells whose atoms' bounding box lies within the Verlet radius of the bounding box of
// this cell's atoms (single precision, boxes rounded outward, radius widened: conservative, membership stays exact).  The
// stencil's semi-volume is 62.5 cells; about a third of its candidates sit in cells out of reach.
__device__ __forceinline__ int lc_run_table(const LCGeom& g, int nrows, const LCRow* __restrict__ rows, const int* __restrict__ lct_start,
                                            int cx, int cy, int cz, int* run0, int* pre, int lane, const float4* __restrict__ box = nullptr,
                                            float rc2_trim = 0.f) {
  const int nlp = g.nlp;
  float4 mylo = make_float4(0, 0, 0, 0), myhi = mylo;
  if (box) { const int ic = 1 + cx + g.sx * (cy + g.sy * cz); mylo = box[2 * ic]; myhi = box[2 * ic + 1]; }
  const int lo_x = nlp - 1, hi_x = g.nlx + nlp, lo_y = nlp - 1, hi_y = g.nly + nlp, lo_z = nlp - 1, hi_z = g.nlz + nlp;
  const bool border = (cx - lo_x <= nlp) || (hi_x - cx <= nlp) || (cy - lo_y <= nlp) || (hi_y - cy <= nlp) || (cz - lo_z <= nlp) || (hi_z - cz <= nlp);
  // ---- run table: potential run p < nrows: positive row p; p >= nrows: negative row (p-nrows)/2, low / high x part
  const int npot = border ? 3 * nrows : nrows;
  for (int p0 = 0; p0 < LC_MAXRUN; p0 += 32) {
    const int p = p0 + lane;
    int start = 0, len = 0;
    if (p < npot) {
      int xa, xbb, jy, jz;
      if (p < nrows) {
        const LCRow r = rows[p];
        jy = cy + r.dy; jz = cz + r.dz;
        xa = (r.dy == 0 && r.dz == 0) ? cx : cx - r.b;
        xbb = cx + r.b;
      } else {
        const LCRow r = rows[(p - nrows) >> 1];
        const int side = (p - nrows) & 1;
        jy = cy - r.dy; jz = cz - r.dz;
        xa = cx - r.b;
        xbb = (r.dy == 0 && r.dz == 0) ? cx - 1 : cx + r.b;
        const bool row_halo = (jy <= lo_y) || (jy >= hi_y) || (jz <= lo_z) || (jz >= hi_z);
        if (row_halo) { if (side == 1) xbb = xa - 1; }            // whole range is halo: side 0 takes it
        else if (side == 0) xbb = min(xbb, lo_x);                // low-x halo cells
        else xa = max(xa, hi_x);                                 // high-x halo cells
      }
      if (box) {   // cells are convex in x: the reachable ones of a row are contiguous
        const int rowbase = 1 + g.sx * (jy + g.sy * jz);
        while (xa <= xbb && box_dist2(mylo, myhi, box[2 * (rowbase + xa)], box[2 * (rowbase + xa) + 1]) > rc2_trim) ++xa;
        while (xa <= xbb && box_dist2(mylo, myhi, box[2 * (rowbase + xbb)], box[2 * (rowbase + xbb) + 1]) > rc2_trim) --xbb;
      }
      if (xa <= xbb) {
        const int c0 = 1 + xa + g.sx * (jy + g.sy * jz), c1 = 1 + xbb + g.sx * (jy + g.sy * jz);
        start = lct_start[c0];
        len = lct_start[c1 + 1] - start;
      }
    }
    int inc = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int v = __shfl_up_sync(DLP_FULL, inc, d); if (lane >= d) inc += v; }
    const int carry = (p0 == 0) ? 0 : pre[p0];
    run0[p] = start;
    pre[p + 1] = carry + inc;
    if (p0 == 0 && lane == 0) pre[0] = 0;
    __syncwarp();
  }
  return pre[LC_MAXRUN];
}

// SIMPLE 1: no exclusion lists, no frozen pairs, nlp < 3 -- the lean inner loop; 2: the lean loop with exclusion lists.
// RING (lean loops): candidates are additionally pruned one by one against the box of the pass's atoms and compacted into a
// shared-memory ring (fewer passes, dearer staging); without it they are staged 32 at a time as they come.
template <int SIMPLE, int RING = 0>
__global__ void __launch_bounds__(LC_WARPS * 32)
k_list_cell(LCGeom g, int natms, int pitch, int xpitch, int megfrz, int lbook, int max_exclude, int excl_by_gid, int nrows,
            const LCRow* __restrict__ rows, const int* __restrict__ at_list, const int* __restrict__ lct_start,
            const int* __restrict__ cell_s, const int* __restrict__ slot_rank, const double4* __restrict__ posq_s,
            const int2* __restrict__ info_s, const int* __restrict__ pair_k, int ntypes, const int* __restrict__ excl,
            unsigned* __restrict__ nbr, int* __restrict__ nnbr, unsigned* __restrict__ xnbr, int* __restrict__ nxnbr,
            int* __restrict__ status, unsigned long long* __restrict__ cnt64, unsigned sentinel, int prune,
            unsigned* __restrict__ fnbr, int* __restrict__ nfnbr, int fpitch, const float4* __restrict__ box, float rc2_trim) {
  __shared__ double4 s_pi[LC_WARPS][32];
  __shared__ double4 s_pj[LC_WARPS][(SIMPLE && RING) ? 64 : 1];   // ring of staged candidates (lean loops): {x, y, z, squared acceptance radius}
  __shared__ int4 s_cj[LC_WARPS][(SIMPLE && RING) ? 64 : 1];      // ... {slot | halo bit, packed vdW indices, global id, ordering threshold}
  __shared__ int2 s_info[LC_WARPS][32];
  __shared__ int s_run0[LC_WARPS][LC_MAXRUN];      // first slot of run r
  __shared__ int s_pre[LC_WARPS][LC_MAXRUN + 1];   // candidates before run r
  __shared__ int s_exn[LC_WARPS][32];                       // exclusion count / list of the cell's atoms (SIMPLE == 2)
  __shared__ unsigned long long s_exp[LC_WARPS][32];
  __shared__ int s_pk[256];                        // (type_i, type_j) -> vdW potential index + 1, when ntypes <= 16
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ unsigned s_kcp[8];                    // type_j -> its vdW potential indices + 1 against every type_i, 6 bits each (lean loops)
  const bool pk_smem = pair_k != nullptr && ntypes <= 16;
  if (pk_smem) for (int q = threadIdx.x; q < ntypes * ntypes; q += LC_WARPS * 32) s_pk[q] = pair_k[q] + 1;
  if (SIMPLE && threadIdx.x < 8) {
    unsigned kcp = 0;
    if (pair_k != nullptr && (int)threadIdx.x < ntypes)
      for (int ti = 0; ti < ntypes && ti < 5; ++ti) kcp |= (unsigned)(pair_k[ti * ntypes + threadIdx.x] + 1) << (6 * ti);
    s_kcp[threadIdx.x] = kcp;
  }
  __syncthreads();
  const int ncell_dom = g.nlx * g.nly * g.nlz;
  const int nlp = g.nlp;
  long long written = 0;
  int ovf = 0;   // longest row that did not fit (error 106), reported once per warp
  // the warps draw their link cells from a queue (cells hold 0 ... 15 atoms: dealt statically, eight to a block, the block waits
  // for its fullest cell); cells are handed out in order, so neighbouring warps still work on neighbouring cells
  for (;;) {
  int cd = 0;
  if (lane == 0) cd = (int)atomicAdd(&cnt64[2], 1ull);
  cd = __shfl_sync(DLP_FULL, cd, 0);
  if (cd >= ncell_dom) break;
  const int cx = cd % g.nlx + nlp, cy = (cd / g.nlx) % g.nly + nlp, cz = cd / (g.nlx * g.nly) + nlp;
  const int ic = 1 + cx + g.sx * (cy + g.sy * cz);
  const int s_own0 = lct_start[ic], s_own1 = lct_start[ic + 1];
  if (s_own1 == s_own0) continue;
  __syncwarp();
  const int total = lc_run_table(g, nrows, rows, lct_start, cx, cy, cz, s_run0[wid], s_pre[wid], lane, box, rc2_trim);
  for (int a0 = s_own0; a0 < s_own1; a0 += 32) {   // atoms of the cell, 32 at a time (one pass unless the cell is crowded)
    const int na = min(32, s_own1 - a0);
    int cnt = 0, xcnt = 0, fcnt = 0;
    int nex_l = 0;
    const int* ex_l = nullptr;
    __syncwarp();
    if (lane < na) {
      s_pi[wid][lane] = posq_s[a0 + lane];
      const int2 inf = info_s[a0 + lane];
      s_info[wid][lane] = inf;
      if (lbook) {
        ex_l = excl + (size_t)(excl_by_gid ? (inf.x - 1) : at_list[a0 + lane]) * (max_exclude + 1);
        nex_l = ex_l[0];
      }
    }
    __syncwarp();
    const int t0 = slot_rank[a0];
    if (SIMPLE) {
      // lean inner loop (ncu: the first version spent 118 instructions per 32 tests; this one keeps the per-atom work to the
      // distance test, one ballot, one shuffle and the store): lane = candidate; everything that depends on the candidate
      // only -- its vdW potential index against every possible type of the cell's atoms (6 bits each, ntypes <= 5), the
      // ordering rule (jj > ii or jj before the cell) as an integer threshold -- is formed once per 32 candidates.
      if (lane < na) s_info[wid][lane].y = 6 * ((s_info[wid][lane].y & 0xffff) - 1);   // shift of type_i in the packed kc word
      if (SIMPLE == 2) { s_exn[wid][lane] = lane < na ? nex_l : 0; s_exp[wid][lane] = (unsigned long long)(size_t)ex_l; }
      __syncwarp();
      unsigned* const xrow0 = xnbr + (size_t)t0 * xpitch;
      const unsigned ltmask = (1u << lane) - 1u;
      const int cap = pitch - DLP_ROW_PAD;
      unsigned* const row0 = nbr + (size_t)t0 * pitch;
      // the passes of one staged batch: every atom of the cell against the lanes' candidates
      auto passes = [&](const double4& pj, const unsigned jbits, const int2 infj, const int gid_j, const unsigned kcp, const int jrel,
                        const double rc_eff) {
        for (int a = 0; a < na; ++a) {
          const double4 pi = s_pi[wid][a];
          const bool acc = pair_rsq(pj, pi.x, pi.y, pi.z) <= rc_eff && jrel > a;
          unsigned m = __ballot_sync(DLP_FULL, acc);
          if (m == 0u) continue;
          bool isx = false;
          if (SIMPLE == 2) {   // partners on the atom's exclusion list go to its xnbr row (neighbours.F90:1229-1251)
            const int nex = s_exn[wid][a];
            if (nex > 0) {
              if (acc) isx = excl_match(infj.x, nex, reinterpret_cast<const int*>((size_t)s_exp[wid][a]));
              const unsigned mx = __ballot_sync(DLP_FULL, isx);
              if (mx) {
                const int cxa = __shfl_sync(DLP_FULL, xcnt, a);
                if (isx) {
                  const unsigned entry = jbits | ((s_info[wid][a].x < gid_j) ? DLP_F_ECNT : 0u);
                  const int ll = cxa + __popc(mx & ltmask);
                  if (ll < xpitch) xrow0[(unsigned)(a * xpitch + ll)] = entry; else atomicOr(&status[0], 2);
                }
                if (lane == a) xcnt += __popc(mx);
                m &= ~mx;
              }
            }
          }
          const int ca = __shfl_sync(DLP_FULL, cnt, a);
          if (acc && !isx) {
            const int2 infi = s_info[wid][a];
            const unsigned entry = jbits | (((kcp >> infi.y) & 63u) << DLP_K_SHIFT) | ((infi.x < gid_j) ? DLP_F_ECNT : 0u);
            const int ll = ca + __popc(m & ltmask);
            if (ll < cap) row0[(unsigned)(a * pitch + ll)] = entry;   // 32-bit offset inside the cell's block of rows
            else ovf = max(ovf, ll + 1);
          }
          if (lane == a) cnt += __popc(m);
        }
      };
      if (RING) {
        // Candidates are pruned while they are staged: a slot farther than the Verlet radius from the bounding box of the pass's
        // atoms cannot be anybody's partner (the stencil's semi-volume is 62.5 cells, the semi-ball around a cell's atoms about
        // half of that), and the survivors are compacted into a 64-entry ring in shared memory, from which the atoms take 32 at
        // a time.  The test runs in single precision relative to the box centre with the box and the radius widened by far more
        // than its rounding (coordinates within ~2 rx of the centre carry < 1e-5 A of fp32 error), so it never drops a slot whose
        // squared distance to some atom is <= rcsq; order is kept, so the rows are the same rows.
        const double4 pref = s_pi[wid][0];   // reference point of the single-precision box arithmetic
        float bx0, bx1, by0, by1, bz0, bz1;
        {
          const double4 pb = s_pi[wid][min(lane, na - 1)];
          bx0 = bx1 = (float)(pb.x - pref.x); by0 = by1 = (float)(pb.y - pref.y); bz0 = bz1 = (float)(pb.z - pref.z);
        }
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          bx0 = fminf(bx0, __shfl_xor_sync(DLP_FULL, bx0, d)); bx1 = fmaxf(bx1, __shfl_xor_sync(DLP_FULL, bx1, d));
          by0 = fminf(by0, __shfl_xor_sync(DLP_FULL, by0, d)); by1 = fmaxf(by1, __shfl_xor_sync(DLP_FULL, by1, d));
          bz0 = fminf(bz0, __shfl_xor_sync(DLP_FULL, bz0, d)); bz1 = fmaxf(bz1, __shfl_xor_sync(DLP_FULL, bz1, d));
        }
        const float bcx = 0.5f * (bx0 + bx1), bcy = 0.5f * (by0 + by1), bcz = 0.5f * (bz0 + bz1);
        const float bhx = 0.5f * (bx1 - bx0) + 1.0e-4f, bhy = 0.5f * (by1 - by0) + 1.0e-4f, bhz = 0.5f * (bz1 - bz0) + 1.0e-4f;
        const float rc_prune = prune ? (float)(g.rcsq * (1.0 + 1.0e-4) + 1.0e-3) : 3.0e38f;
        double4* const ring_p = s_pj[wid];
        int4* const ring_c = s_cj[wid];
        int run = 0, head = 0, fill = 0;   // ring: `fill` staged candidates from `head`
        for (int c0 = 0; c0 < total || fill > 0; c0 += 32) {
          if (c0 < total) {   // stage up to 32 more candidates; a lane's candidates only move forward, so its run index does too
            const int c = c0 + lane;
            int jj = -1;
            double4 pj = make_double4(0.0, 0.0, 0.0, g.rcsq);
            int2 infj = make_int2(0, 0);
            bool keep = false;
            if (c < total) {
              while (s_pre[wid][run + 1] <= c) ++run;          // pre[LC_MAXRUN] = total > c ends the walk
              jj = s_run0[wid][run] + (c - s_pre[wid][run]);
              pj = posq_s[jj]; infj = info_s[jj];
              // nlp == 2: the only stencil entry with the "no distance check" flag is the cell itself (neighbours.F90:537, :922-943)
              const bool own = jj >= s_own0 && jj < s_own1;
              pj.w = (g.nir_r2 > 0 && own) ? 1e301 : g.rcsq;
              const float dx = fmaxf(fabsf((float)(pj.x - pref.x) - bcx) - bhx, 0.0f), dy = fmaxf(fabsf((float)(pj.y - pref.y) - bcy) - bhy, 0.0f),
                          dz = fmaxf(fabsf((float)(pj.z - pref.z) - bcz) - bhz, 0.0f);
              keep = own || (dx * dx + dy * dy + dz * dz <= rc_prune);
            }
            const unsigned mk = __ballot_sync(DLP_FULL, keep);
            if (keep) {
              int4 cj;
              cj.x = (int)((unsigned)jj | (((infj.y >> 17) & 1) ? DLP_F_HALO : 0u));
              cj.y = (int)s_kcp[((infj.y & 0xffff) - 1) & 7];   // lean loops: ntypes <= 5
              cj.z = infj.x;
              // (jj > ii || jj < s_own0) with ii = a0 + a  <=>  jrel > a
              cj.w = jj < s_own0 ? 0x7fffffff : jj - a0;
              const int q = (head + fill + __popc(mk & ltmask)) & 63;
              ring_p[q] = pj; ring_c[q] = cj;
            }
            fill += __popc(mk);
            __syncwarp();
            if (fill < 32 && c0 + 32 < total) continue;          // wait for a full batch while there are candidates left
          }
          const int nb = min(32, fill);
          double4 pj = make_double4(1e300, 1e300, 1e300, g.rcsq);   // lanes without a candidate sit at 1e300: never within the cutoff
          int4 cj = make_int4(-1, 0, 0, -1);
          if (lane < nb) { const int q = (head + lane) & 63; pj = ring_p[q]; cj = ring_c[q]; }
          head = (head + nb) & 63; fill -= nb;
          if (c0 >= total) c0 -= 32;                                // draining the ring: no new candidates
          __syncwarp();
          const unsigned jbits = (unsigned)cj.x;
          const int2 infj = make_int2(cj.z, 0);
          const int gid_j = (jbits & DLP_F_HALO) ? cj.z : 0;       // energy ownership: halo partner and idi < ltg(jatm); gids are >= 1
          const unsigned kcp = (unsigned)cj.y;
          const int jrel = cj.w;
          const double rc_eff = pj.w;
          passes(pj, jbits, infj, gid_j, kcp, jrel, rc_eff);
        }
      } else {
        // plain staging: 32 candidates at a time as the (trimmed) runs deliver them
        int run = 0;
        for (int c0 = 0; c0 < total; c0 += 32) {
          const int c = c0 + lane;
          int jj = -1;
          double4 pj = make_double4(1e300, 1e300, 1e300, 0);   // lanes without a candidate sit at 1e300: never within the cutoff
          int2 infj = make_int2(0, 0);
          if (c < total) {   // run of candidate c: the lane's candidates only move forward, so its run index does too
            while (s_pre[wid][run + 1] <= c) ++run;          // pre[LC_MAXRUN] = total > c ends the walk
            jj = s_run0[wid][run] + (c - s_pre[wid][run]);
            pj = posq_s[jj]; infj = info_s[jj];
          }
          const bool halo_j = (infj.y >> 17) & 1;
          const unsigned jbits = (unsigned)jj | (halo_j ? DLP_F_HALO : 0u);
          const int gid_j = halo_j ? infj.x : 0;               // energy ownership: halo partner and idi < ltg(jatm); gids are >= 1
          const unsigned kcp = jj >= 0 ? s_kcp[((infj.y & 0xffff) - 1) & 7] : 0u;
          // (jj > ii || jj < s_own0) with ii = a0 + a  <=>  jrel > a
          const int jrel = jj < 0 ? -1 : (jj < s_own0 ? 0x7fffffff : jj - a0);
          // nlp == 2: the only stencil entry with the "no distance check" flag is the cell itself (neighbours.F90:537, :922-943)
          const double rc_eff = (g.nir_r2 > 0 && jj >= s_own0 && jj < s_own1) ? 1e301 : g.rcsq;
          passes(pj, jbits, infj, gid_j, kcp, jrel, rc_eff);
        }
      }
    } else
    {
    for (int c0 = 0; c0 < total; c0 += 32) {
      const int c = c0 + lane;
      const bool valid = c < total;
      int jj = 0;
      if (valid) {   // largest run r with pre[r] <= c
        int lo = 0, hi = LC_MAXRUN - 1;
        while (lo < hi) { const int mid = (lo + hi + 1) >> 1; if (s_pre[wid][mid] <= c) lo = mid; else hi = mid - 1; }
        jj = s_run0[wid][lo] + (c - s_pre[wid][lo]);
      }
      double4 pj = make_double4(0, 0, 0, 0);
      int2 infj = make_int2(0, 0);
      int cj = 0;
      if (valid) { pj = posq_s[jj]; infj = info_s[jj]; if (g.nir_r2 > 0) cj = cell_s[jj]; }
      const bool halo_j = (infj.y >> 17) & 1;
      const bool nir = valid && g.nir_r2 > 0 && nir_of(g, cj, cx, cy, cz);
      for (int a = 0; a < na; ++a) {
        const double4 pi = s_pi[wid][a];
        const int2 infi = s_info[wid][a];
        const int ii = a0 + a;
        bool acc = valid && (jj > ii || jj < s_own0) && (nir || pair_rsq(pj, pi.x, pi.y, pi.z) <= g.rcsq);
        // frozen-frozen pairs never reach the force loops (:1198-1225); they are kept in rows of their own for rdf_frzn_collect
        const bool isf = acc && megfrz > 1 && ((infi.y >> 16) & 1) && ((infj.y >> 16) & 1);
        if (isf) acc = false;
        if (megfrz > 1 && fnbr != nullptr) {
          const unsigned mf = __ballot_sync(DLP_FULL, isf);
          if (mf) {
            const int cfa = __shfl_sync(DLP_FULL, fcnt, a);
            if (isf) {
              const unsigned entry = (unsigned)jj | (halo_j ? DLP_F_HALO : 0u) | ((halo_j && infi.x < infj.x) ? DLP_F_ECNT : 0u);
              const int ll = cfa + __popc(mf & ((1u << lane) - 1));
              if (ll < fpitch) fnbr[(size_t)(t0 + a) * fpitch + ll] = entry; else atomicOr(&status[0], 1);
            }
            if (lane == a) fcnt += __popc(mf);
          }
        }
        bool isx = false;
        if (lbook) {
          const int nex = __shfl_sync(DLP_FULL, nex_l, a);
          const unsigned long long exq = __shfl_sync(DLP_FULL, (unsigned long long)(size_t)ex_l, a);
          if (acc && nex > 0 && excl_match(infj.x, nex, reinterpret_cast<const int*>((size_t)exq))) { isx = true; acc = false; }
        }
        const unsigned m = __ballot_sync(DLP_FULL, acc);
        const int ca = __shfl_sync(DLP_FULL, cnt, a);
        if (acc) {
          const int kc = pair_k ? pair_k[((infi.y & 0xffff) - 1) * ntypes + ((infj.y & 0xffff) - 1)] + 1 : 0;
          const unsigned entry = (unsigned)jj | ((unsigned)kc << DLP_K_SHIFT) | (halo_j ? DLP_F_HALO : 0u) |
                                 ((halo_j && infi.x < infj.x) ? DLP_F_ECNT : 0u);
          const int ll = ca + __popc(m & ((1u << lane) - 1));
          if (ll < pitch - DLP_ROW_PAD) nbr[(size_t)(t0 + a) * pitch + ll] = entry;
          else { atomicOr(&status[0], 1); atomicMax(&status[1], ll + 1); }
        }
        if (lane == a) cnt += __popc(m);
        if (lbook) {
          const unsigned mx = __ballot_sync(DLP_FULL, isx);
          const int cxa = __shfl_sync(DLP_FULL, xcnt, a);
          if (isx) {
            const unsigned entry = (unsigned)jj | (halo_j ? DLP_F_HALO : 0u) | ((halo_j && infi.x < infj.x) ? DLP_F_ECNT : 0u);
            const int ll = cxa + __popc(mx & ((1u << lane) - 1));
            if (ll < xpitch) xnbr[(size_t)(t0 + a) * xpitch + ll] = entry; else atomicOr(&status[0], 2);
          }
          if (lane == a) xcnt += __popc(mx);
        }
      }
    }
    }
    if (lane < na) { nnbr[t0 + lane] = cnt; nxnbr[t0 + lane] = xcnt; written += cnt; if (nfnbr != nullptr) nfnbr[t0 + lane] = fcnt; }
    for (int a = 0; a < na; ++a)   // sentinel padding the pair kernel relies on
      dlp_pad_row(nbr + (size_t)(t0 + a) * pitch, min(__shfl_sync(DLP_FULL, cnt, a), pitch - DLP_ROW_PAD), sentinel, lane, 32);
  }
  }   // cell queue
  if (ovf > 0) { atomicOr(&status[0], 1); atomicMax(&status[1], ovf); }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) written += __shfl_xor_sync(DLP_FULL, written, d);
  if (lane == 0) atomicAdd(&cnt64[0], (unsigned long long)written);
}

__global__ void k_bg_copy(int n, const double4* __restrict__ posq, double* xbg, double* ybg, double* zbg) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { double4 p = posq[i]; xbg[i] = p.x; ybg[i] = p.y; zbg[i] = p.z; }
}

}  // namespace

int dlp_gather_sorted(dlpgpu_ctx* ctx) {
  int n = ctx->list_nlast;
  if (n > 0) LAUNCH(ctx, k_gather_posq, cdiv(n + 1, 256), 256, 0, n, ctx->at_list.p, ctx->posq.p, ctx->posq_s.p);
  return 0;
}

namespace {
// The reference's half list gives the first atoms of a link cell up to four times the partners of the last ones (56 ... 221 per
// row in the NaCl melt), and the four rows that share a warp of k_pair_v2 run until the longest is done: 19 % of the lane
// slots idle (ncu: 26.1 active threads per instruction).  Ordering the rows of every window of DLP_ROW_WIN atoms by length
// lets a warp take four rows of (nearly) the same length; the kernel rotates the length classes over its warps pass by pass.
// (Cutting the windows from rows taken in the order of nlp^3 blocks of link cells instead of the sorted arrays' cell order -- a more
// compact neighbourhood per window -- changed nothing: 1.216 against 1.212-1.220 ms.)
__global__ void __launch_bounds__(DLP_ROW_WIN) k_row_perm(int natms, const int* __restrict__ nnbr, int* __restrict__ perm) {
  __shared__ int s_key[DLP_ROW_WIN];
  const int i = threadIdx.x, t = blockIdx.x * DLP_ROW_WIN + i;
  const int key = t < natms ? nnbr[t] : -1;
  s_key[i] = key;
  __syncthreads();
  int rank = 0;
#pragma unroll 8
  for (int j = 0; j < DLP_ROW_WIN; ++j) { const int kj = s_key[j]; rank += (kj > key) || (kj == key && j < i); }
  perm[blockIdx.x * DLP_ROW_WIN + rank] = t < natms ? t : -1;
}
}  // namespace

int dlp_build_lists(dlpgpu_ctx* ctx, int want_ref_list, int* ibig) {
  cudaStream_t s = ctx->stream;
  ctx->list_valid = false; ctx->ref_valid = false;
  if (ibig) *ibig = 0;
  if (ctx->max_list < 1) return dlp_fail(ctx, DLPGPU_ERR_STATE, "build: max_list not set");
  CKRC(h_geometry(ctx));
  LCGeom& g = ctx->g;
  const int natms = ctx->natms, nlast = ctx->nlast;
  cudaEventRecord(ctx->ev[0], s);
  // stencil upload
  size_t ns = ctx->h_nix.size();
  CK(ctx->st_nix.ensure(ns, s)); CK(ctx->st_niy.ensure(ns, s)); CK(ctx->st_niz.ensure(ns, s)); CK(ctx->st_nir.ensure(ns, s));
  CK(ctx->st_xb.ensure(ctx->h_xb.size(), s));
  // the stencil only changes with the link-cell geometry: upload it when it differs from what the device holds
  std::vector<int> key;
  key.reserve(4 * ns + ctx->h_xb.size());
  key.insert(key.end(), ctx->h_nix.begin(), ctx->h_nix.end()); key.insert(key.end(), ctx->h_niy.begin(), ctx->h_niy.end());
  key.insert(key.end(), ctx->h_niz.begin(), ctx->h_niz.end()); key.insert(key.end(), ctx->h_nir.begin(), ctx->h_nir.end());
  key.insert(key.end(), ctx->h_xb.begin(), ctx->h_xb.end());
  if (key != ctx->st_uploaded) {
    CK(cudaMemcpyAsync(ctx->st_nix.p, ctx->h_nix.data(), ns * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_niy.p, ctx->h_niy.data(), ns * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_niz.p, ctx->h_niz.data(), ns * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_nir.p, ctx->h_nir.data(), ns * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->st_xb.p, ctx->h_xb.data(), ctx->h_xb.size() * sizeof(int), cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));   // pageable host vectors: the copies are staged, but keep the key in step with the device
    ctx->st_uploaded.swap(key);
  }
  // buffers
  size_t nc = (size_t)g.ncells + 3;
  CK(ctx->which_cell.ensure(nlast + 1, s)); CK(ctx->at_list.ensure(nlast + 1, s)); CK(ctx->at_tmp.ensure(nlast + 1, s));
  CK(ctx->cell_s.ensure(nlast + 1, s)); CK(ctx->posq_s.ensure(nlast + 1, s));
  CK(ctx->type_s.ensure(nlast + 1, s)); CK(ctx->gid_s.ensure(nlast + 1, s)); CK(ctx->frz_s.ensure(nlast + 1, s));
  CK(ctx->info_s.ensure(nlast + 1, s));
  if (nlast >= DLP_MAX_SLOTS) return dlp_fail(ctx, DLPGPU_ERR_ARG, "more than %d resident atoms (local + halo) per GPU are not supported", DLP_MAX_SLOTS);
  CK(ctx->lct_count.ensure(nc, s)); CK(ctx->lct_start.ensure(nc, s)); CK(ctx->lct_fill.ensure(nc, s));
  CK(ctx->loc_slot.ensure(natms + 1, s));
  CK(ctx->flag.ensure((size_t)nlast + 2, s)); CK(ctx->scan_out.ensure((size_t)nlast + 2, s));
  CK(cudaMemsetAsync(ctx->lct_count.p, 0, nc * sizeof(int), s));
  CK(cudaMemsetAsync(ctx->lct_fill.p, 0, nc * sizeof(int), s));
  CK(cudaMemsetAsync(ctx->status.p, 0, 8 * sizeof(int), s));
  if (nlast > 0) LAUNCH(ctx, k_cell_index, cdiv(nlast, 256), 256, 0, g, natms, nlast, ctx->posq.p, ctx->which_cell.p, ctx->lct_count.p);
  // lct_start[c] for c = 0..ncells+1 (0-based slots): exclusive scan over counts of cells 0..ncells (+ zero pad)
  CKRC(dlp_exclusive_scan(ctx, ctx->lct_count.p, ctx->lct_start.p, g.ncells + 1, nullptr));
  if (nlast > 0) {
    LAUNCH(ctx, k_cell_scatter, cdiv(nlast, 256), 256, 0, nlast, ctx->which_cell.p, ctx->lct_start.p, ctx->lct_fill.p, ctx->at_tmp.p);
    LAUNCH(ctx, k_cell0_flag, cdiv(nlast + 1, 256), 256, 0, nlast, ctx->which_cell.p, ctx->flag.p);
    CKRC(dlp_exclusive_scan(ctx, ctx->flag.p, ctx->scan_out.p, nlast, nullptr));
    LAUNCH(ctx, k_cell0_place, cdiv(nlast, 256), 256, 0, nlast, ctx->which_cell.p, ctx->scan_out.p, ctx->at_list.p, ctx->cell_s.p);
    LAUNCH(ctx, k_cell_order, cdiv((long long)g.ncells * 32, 256), 256, 0, g.ncells + 1, ctx->lct_start.p, ctx->at_tmp.p,
           ctx->at_list.p, ctx->cell_s.p);
  }
  LAUNCH(ctx, k_sorted_static, cdiv(nlast + 1, 256), 256, 0, nlast, natms, ctx->at_list.p, ctx->ltype.p, ctx->ltg.p, ctx->lfrzn.p,
         ctx->type_s.p, ctx->gid_s.p, ctx->frz_s.p, ctx->info_s.p, ctx->flag.p);
  CKRC(dlp_exclusive_scan(ctx, ctx->flag.p, ctx->scan_out.p, nlast, nullptr));
  if (nlast > 0) {
    LAUNCH(ctx, k_loc_slot, cdiv(nlast, 256), 256, 0, nlast, ctx->flag.p, ctx->scan_out.p, ctx->loc_slot.p);
    LAUNCH(ctx, k_gather_posq, cdiv(nlast + 1, 256), 256, 0, nlast, ctx->at_list.p, ctx->posq.p, ctx->posq_s.p);
  }
  ctx->list_natms = natms; ctx->list_nlast = nlast;
  // device-internal lists
  // room for the sentinel padding (see dlp_pad_row).  Full-list mode keeps every local-local pair in BOTH rows, so a row can be
  // twice as long as the reference's half row that neigh%max_list was sized for (error 106 stays the reference's: it is raised by
  // the reference-format list, k_list_ref, against max_list itself)
  const int row_cap = ctx->force_mode == 0 ? 2 * ctx->max_list : ctx->max_list;
  ctx->pitch = ((row_cap + 31) / 32) * 32 + DLP_ROW_PAD;
  const unsigned sentinel = (unsigned)nlast | DLP_F_HALO;   // slot nlast of posq_s: a chargeless point 1e15 A away
  ctx->xpitch = ctx->lbook ? ((ctx->max_exclude + 31) / 32) * 32 : 0;
  const int wpb = 8;   // warps per block
  if (natms > 0) {
    CK(ctx->nnbr.ensure(natms + 1, s)); CK(ctx->nxnbr.ensure(natms + 1, s));
    CK(ctx->xnbr.ensure((size_t)natms * std::max(ctx->xpitch, 1) + 1, s));
    // frozen-frozen pairs (rdf_frzn_collect): rows of their own, only when the system has frozen atoms
    ctx->fpitch = ctx->megfrz > 1 ? ctx->pitch - DLP_ROW_PAD : 0;
    ctx->frz_rows_valid = false;
    if (ctx->fpitch > 0) { CK(ctx->fnbr.ensure((size_t)natms * ctx->fpitch + 1, s)); CK(ctx->nfnbr.ensure(natms + 1, s)); }
    cudaEventRecord(ctx->ev[2], s);
    if (ctx->force_mode == 0) {
      CK(ctx->nbr.ensure((size_t)natms * ctx->pitch + 256, s));
      LAUNCH(ctx, k_list_dev<false>, cdiv(natms, wpb), wpb * 32, 0, g, natms, ctx->pitch, ctx->xpitch, ctx->megfrz, ctx->lbook,
             ctx->max_exclude, ctx->excl_by_gid, ctx->loc_slot.p, ctx->at_list.p, ctx->lct_start.p, ctx->cell_s.p, ctx->posq_s.p,
             ctx->gid_s.p, ctx->frz_s.p, ctx->type_s.p, ctx->vdw_on ? ctx->pair_k.p : nullptr, ctx->ntypes, ctx->st_xb.p, ctx->excl.p,
             ctx->nbr.p, ctx->nnbr.p, ctx->xnbr.p, ctx->nxnbr.p, ctx->status.p, sentinel);
    } else {
      CK(ctx->nbr.ensure((size_t)natms * ctx->pitch + 256, s));
      // semi-ball rows (dy,dz) with their half x-extent, for the warp-per-cell kernel
      std::vector<LCRow> hrows;
      const int nlp = g.nlp, W = 2 * nlp + 1;
      for (int dz = 0; dz <= nlp; ++dz)
        for (int dy = -nlp; dy <= nlp; ++dy) {
          if (dz == 0 && dy < 0) continue;
          int b = ctx->h_xb[(size_t)(dz + nlp) * W + (dy + nlp)];
          if (b >= 0) hrows.push_back(LCRow{dy, dz, b});
        }
      CK(cudaMemsetAsync(ctx->cnt64.p, 0, 4 * sizeof(unsigned long long), s));
      if (3 * hrows.size() <= LC_MAXRUN) {
        CK(ctx->st_rows.ensure(hrows.size() * 3 + 3, s));
        CK(cudaMemcpyAsync(ctx->st_rows.p, hrows.data(), hrows.size() * sizeof(LCRow), cudaMemcpyHostToDevice, s));
        const int ncd = g.nlx * g.nly * g.nlz;
        const bool lean = ctx->megfrz <= 1 && g.nir_r2 <= 1 && ctx->ntypes <= 5;
        const bool simple = lean && !ctx->lbook;
        // run trimming against the cells' bounding boxes; not with the nlp >= 3 shortcut (neighbours.F90:537 admits whole cells
        // without a distance test, so a cell out of reach of the atoms' box may still have to be listed)
        const int lk = ctx->list_one_atom_per_pass;   // dlpgpu_set_list_kernel: 0 trimmed runs (default), 1 untrimmed, 2 trimmed + per-candidate ring
        const bool trim = lk != 1 && g.nir_r2 <= 1;
        const float rc2_trim = (float)(g.rcsq * (1.0 + 1.0e-4) + 1.0e-3);
        if (trim) {
          CK(ctx->cell_box.ensure((size_t)8 * (g.ncells + 2), s));
          LAUNCH(ctx, k_cell_boxes, cdiv(g.ncells + 1, 128), 128, 0, g.ncells, ctx->lct_start.p, ctx->posq_s.p, reinterpret_cast<float4*>(ctx->cell_box.p));
        }
const int lc_grid = std::max(1, std::min(cdiv(ncd, LC_WARPS), ctx->sm_count * 8));   // resident blocks; the warps draw cells from a queue
#define DLP_LC_ARGS g, natms, ctx->pitch, std::max(ctx->xpitch, 1), ctx->megfrz, ctx->lbook, ctx->max_exclude, ctx->excl_by_gid, \
               (int)hrows.size(), reinterpret_cast<const LCRow*>(ctx->st_rows.p), ctx->at_list.p, ctx->lct_start.p, ctx->cell_s.p, \
               ctx->scan_out.p, ctx->posq_s.p, ctx->info_s.p, ctx->vdw_on ? ctx->pair_k.p : nullptr, ctx->ntypes, ctx->excl.p, ctx->nbr.p, \
               ctx->nnbr.p, ctx->xnbr.p, ctx->nxnbr.p, ctx->status.p, ctx->cnt64.p, sentinel, 1, \
               ctx->fnbr.p, ctx->nfnbr.p, ctx->fpitch, trim ? reinterpret_cast<const float4*>(ctx->cell_box.p) : nullptr, rc2_trim
        if (simple && lk == 2) LAUNCH(ctx, (k_list_cell<1, 1>), lc_grid, LC_WARPS * 32, 0, DLP_LC_ARGS);
        else if (simple) LAUNCH(ctx, (k_list_cell<1, 0>), lc_grid, LC_WARPS * 32, 0, DLP_LC_ARGS);
        else if (lean && lk == 2) LAUNCH(ctx, (k_list_cell<2, 1>), lc_grid, LC_WARPS * 32, 0, DLP_LC_ARGS);
        else if (lean) LAUNCH(ctx, (k_list_cell<2, 0>), lc_grid, LC_WARPS * 32, 0, DLP_LC_ARGS);
        else { LAUNCH(ctx, (k_list_cell<0, 0>), lc_grid, LC_WARPS * 32, 0, DLP_LC_ARGS); ctx->frz_rows_valid = ctx->fpitch > 0; }
#undef DLP_LC_ARGS
      } else {   // very fine sub-celling (nlp >= 4): the per-atom kernel has no run-table limit
        LAUNCH(ctx, k_list_dev<true>, cdiv(natms, wpb), wpb * 32, 0, g, natms, ctx->pitch, ctx->xpitch, ctx->megfrz, ctx->lbook,
               ctx->max_exclude, ctx->excl_by_gid, ctx->loc_slot.p, ctx->at_list.p, ctx->lct_start.p, ctx->cell_s.p, ctx->posq_s.p,
               ctx->gid_s.p, ctx->frz_s.p, ctx->type_s.p, ctx->vdw_on ? ctx->pair_k.p : nullptr, ctx->ntypes, ctx->st_xb.p, ctx->excl.p,
               ctx->nbr.p, ctx->nnbr.p, ctx->xnbr.p, ctx->nxnbr.p, ctx->status.p, sentinel);
      }
    }
    cudaEventRecord(ctx->ev[3], s);
    if (want_ref_list) {
      CK(ctx->ref_list.ensure((size_t)natms * (ctx->max_list + 4) + 1, s));
      LAUNCH(ctx, k_list_ref, cdiv(natms, wpb), wpb * 32, 0, g, natms, ctx->max_list, ctx->loc_slot.p, ctx->at_list.p, ctx->lct_start.p,
             ctx->cell_s.p, ctx->posq_s.p, ctx->st_nix.p, ctx->st_niy.p, ctx->st_niz.p, ctx->st_nir.p, ctx->ref_list.p, ctx->status.p);
      LAUNCH(ctx, k_row_partition, cdiv(natms, 128), 128, 0, natms, ctx->max_list, ctx->megfrz, ctx->lbook, ctx->max_exclude,
             ctx->excl_by_gid, ctx->lfrzn.p, ctx->ltg.p, ctx->excl.p, ctx->ref_list.p);
    }
  }
  if (natms > 0) {
    const int nwin = cdiv(natms, DLP_ROW_WIN);
    CK(ctx->row_perm.ensure((size_t)nwin * DLP_ROW_WIN, s));
    LAUNCH(ctx, k_row_perm, nwin, DLP_ROW_WIN, 0, natms, ctx->nnbr.p, ctx->row_perm.p);
  }
  cudaEventRecord(ctx->ev[1], s);
  int st[8];
  unsigned long long c64[4] = {0, 0, 0, 0};
  CK(cudaMemcpyAsync(st, ctx->status.p, 8 * sizeof(int), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(c64, ctx->cnt64.p, sizeof c64, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  ctx->list_entries = c64[0] ? (long long)c64[0] : (long long)natms * (ctx->max_list / 4);   // row-length hint for the force kernel
  float ms = 0;
  cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[1]); ctx->t_list = ms;
  if (natms > 0) { cudaEventElapsedTime(&ms, ctx->ev[2], ctx->ev[3]); ctx->t_full = ms; }
  if (st[0]) {
    if (ibig) *ibig = st[1];
    return dlp_fail(ctx, DLPGPU_ERR_LIST_OVERFLOW, "error 106: neighbour list array exceeded (row length %d > max_list %d)", st[1], ctx->max_list);
  }
  ctx->list_valid = true;
  ctx->ref_valid = want_ref_list != 0 && natms > 0;
  return 0;
}

int dlp_vnl_set_check(dlpgpu_ctx* ctx) {
  int n = ctx->nlast;
  if (n > 0) LAUNCH(ctx, k_bg_copy, cdiv(n, 256), 256, 0, n, ctx->posq.p, ctx->xbg.p, ctx->ybg.p, ctx->zbg.p);
  ctx->have_bg = true; ctx->tol_fresh = false;
  return 0;
}

namespace {
__global__ void k_count_pairs(int natms, int pitch, const unsigned* __restrict__ nbr, const int* __restrict__ nnbr,
                              unsigned long long* __restrict__ out /*[2]: local partners, halo partners*/) {
  int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (t >= natms) return;
  int n = nnbr[t], nl = 0, nh = 0;
  for (int k = lane; k < n; k += 32) { if (nbr[(size_t)t * pitch + k] & DLP_F_HALO) ++nh; else ++nl; }
  for (int d = 16; d > 0; d >>= 1) { nl += __shfl_xor_sync(DLP_FULL, nl, d); nh += __shfl_xor_sync(DLP_FULL, nh, d); }
  if (lane == 0) { atomicAdd(&out[0], (unsigned long long)nl); atomicAdd(&out[1], (unsigned long long)nh); }
}
}  // namespace

// number of pairs the reference's half list holds for this domain (local-local once + local-halo), from the device list
extern "C" int dlpgpu_dev_list_pairs(dlpgpu_ctx* ctx, long long* pairs) {
  if (!ctx || !pairs) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->list_valid) return dlp_fail(ctx, DLPGPU_ERR_STATE, "list_pairs: no list");
  unsigned long long h[2] = {0, 0};
  CK(cudaMemsetAsync(ctx->tol_bits.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
  if (ctx->list_natms > 0)
    LAUNCH(ctx, k_count_pairs, cdiv(ctx->list_natms, 8), 256, 0, ctx->list_natms, ctx->pitch, ctx->nbr.p, ctx->nnbr.p, ctx->tol_bits.p);
  CK(cudaMemcpyAsync(h, ctx->tol_bits.p, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *pairs = (long long)(ctx->force_mode == 1 ? h[0] : h[0] / 2) + (long long)h[1];
  return 0;
}

extern "C" int dlpgpu_dev_link_cell_pairs(dlpgpu_ctx* ctx, int want_ref_list, int* ibig) {
  if (!ctx) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  return dlp_build_lists(ctx, want_ref_list, ibig);
}

int dlp_preload_cells() {   // see dlp_preload_halo
  const void* ks[] = {(const void*)k_cell_index, (const void*)k_cell_scatter, (const void*)k_cell0_flag, (const void*)k_cell0_place,
                      (const void*)k_cell_order, (const void*)k_sorted_static, (const void*)k_loc_slot, (const void*)k_gather_posq,
                      (const void*)k_list_ref, (const void*)k_row_partition, (const void*)k_list_dev<false>, (const void*)k_list_dev<true>,
                      (const void*)k_list_cell<0, 0>, (const void*)k_list_cell<1, 0>, (const void*)k_list_cell<2, 0>, (const void*)k_list_cell<1, 1>,
                      (const void*)k_list_cell<2, 1>, (const void*)k_cell_boxes, (const void*)k_bg_copy, (const void*)k_row_perm,
                      (const void*)k_count_pairs};
  cudaFuncAttributes a;
  for (const void* k : ks) if (cudaFuncGetAttributes(&a, k) != cudaSuccess) cudaGetLastError();
  return 0;
}
