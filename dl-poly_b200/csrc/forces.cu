// forces.cu -- the pair loops of two_body.F90::two_body_forces on the device-resident lists.
//
//   two_body.F90:339-352   distance gather                     vdw.F90:1790-2024   vdw_forces_tab
//   vdw.F90:1578-1788      vdw_forces_direct                   ewald_spole.F90:58-242   ewald_real_forces_coul
//   two_body.F90:552-606 + ewald_spole.F90:479-679  excluded-pair Ewald correction
//
// One warp per local atom (lanes = neighbours), warps persistent over atoms.  Per-pair arithmetic keeps the reference's
// operation order (this file is compiled with -fmad=false), so a pair term computed here has the same bits as the
// reference's; only the order in which pair terms are summed differs (the 1e-9 / 1e-10 allowance of the north star).
// FULL mode: each local-local pair is evaluated from both ends and weighted 1/2 in energy / virial / stress (exact:
// the two evaluations are bitwise mirror images), local-halo pairs once with the reference's global-id ownership rule.
// HALF mode: Newton's third law with fp64 RED atomics on the partner.
#include "common.cuh"

namespace {

struct FParams {
  int natms, pitch, xpitch, ntypes, max_grid, max_vdw, ew_n;
  int vdw_on, vdw_direct, vdw_fshift, ew_on, half, zero_forces, lbook;
  double rvdw, r_rvdw, vdw_rdr, rcut, ew_rdr, alpha, scaling;
};

constexpr double ZERO_PLUS = 2.2250738585072014e-308;   // Tiny(1.0_wp), constants.F90:189

__device__ __forceinline__ double powi6(double x) { double x2 = x * x; double x4 = x2 * x2; return x2 * x4; }   // __powidf2(x,6)

// two_body_potentials.F90 analytic forms used by vdw_forces_direct; p = param(1:7) of potential k
__device__ __forceinline__ void pot_direct(int key, const double* __restrict__ p, double r, double& e, double& g) {
  switch (key) {
    case 1: {   // 12-6  :307-317
      double r_6 = powi6(1.0 / r);
      e = (p[0] * r_6 - p[1]) * r_6;
      g = 6.0 * r_6 * (2.0 * p[0] * r_6 - p[1]);
      break;
    }
    case 2: {   // lj  :260-270
      double s6 = powi6(p[1] / r);
      e = 4.0 * p[0] * s6 * (s6 - 1.0);
      g = 24.0 * p[0] * s6 * (2.0 * s6 - 1.0);
      break;
    }
    case 4: {   // buckingham  :471-485
      double b = r / p[1];
      double t1 = p[0] * exp(-b);
      double t2 = -p[2] / powi6(r);
      e = t1 + t2;
      g = t1 * b + 6.0 * t2;
      break;
    }
    case 5: {   // bhm  :499-514
      double ri2 = 1.0 / (r * r);
      double t1 = p[0] * exp(p[1] * (p[2] - r));
      double t2 = -p[3] * (ri2 * (ri2 * ri2));          // r_inv_2**3 : y=x; x=x*x; y=y*x
      double q2 = ri2 * ri2;
      double t3 = -p[4] * (q2 * q2);                    // r_inv_2**4
      e = t1 + t2 + t3;
      g = (t1 * r * p[1] + 6.0 * t2 + 8.0 * t3);
      break;
    }
    case 12: {  // lj cohesive
      double s6 = powi6(p[1] / r);
      e = 4.0 * p[0] * s6 * (s6 - p[2]);
      g = 24.0 * p[0] * s6 * (2.0 * s6 - p[2]);
      break;
    }
    default: e = 0.0; g = 0.0;
  }
}

__device__ __forceinline__ double interp3(double g0, double g1, double g2, double ppp) {
  // numerics.F90:280-282 / vdw.F90:1918-1921
  double t1 = g0 + (g1 - g0) * ppp;
  double t2 = g1 + (g2 - g1) * (ppp - 1.0);
  return t1 + (t2 - t1) * ppp * 0.5;
}

template <bool SMEM>
__global__ void __launch_bounds__(256) k_pair_forces(FParams P, const int* __restrict__ loc_slot, const int* __restrict__ at_list,
                                                     const double4* __restrict__ posq_s, const int* __restrict__ type_s,
                                                     const unsigned* __restrict__ nbr, const int* __restrict__ nnbr,
                                                     const unsigned* __restrict__ xnbr, const int* __restrict__ nxnbr,
                                                     const int* __restrict__ pair_k_g, const int* __restrict__ ltp,
                                                     const double2* __restrict__ vdw_tab_g, const double* __restrict__ vdw_par,
                                                     const double2* __restrict__ ew_tab_g, double* __restrict__ fx,
                                                     double* __restrict__ fy, double* __restrict__ fz, double* __restrict__ fsx,
                                                     double* __restrict__ fsy, double* __restrict__ fsz, double* __restrict__ partial) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const double2* vdw_tab = vdw_tab_g;
  const double2* ew_tab = ew_tab_g;
  const int* pair_k = pair_k_g;
  if (SMEM) {
    double2* sv = reinterpret_cast<double2*>(smem_raw);
    size_t nv = P.vdw_on && !P.vdw_direct ? (size_t)P.max_vdw * (P.max_grid + 1) : 0;
    size_t ne = P.ew_on ? (size_t)P.ew_n + 1 : 0;
    double2* se = sv + nv;
    int* sp = reinterpret_cast<int*>(se + ne);
    for (size_t k = threadIdx.x; k < nv; k += blockDim.x) sv[k] = vdw_tab_g[k];
    for (size_t k = threadIdx.x; k < ne; k += blockDim.x) se[k] = ew_tab_g[k];
    for (int k = threadIdx.x; k < P.ntypes * P.ntypes; k += blockDim.x) sp[k] = P.vdw_on ? pair_k_g[k] : -1;
    __syncthreads();
    vdw_tab = sv; ew_tab = se; pair_k = sp;
  }
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int tstride = P.max_grid + 1;
  double acc[12];
#pragma unroll
  for (int k = 0; k < 12; ++k) acc[k] = 0.0;

  for (int t = gw; t < P.natms; t += nwarps) {
    const int ii = loc_slot[t];
    const double4 pi = posq_s[ii];
    const int ai = type_s[ii] - 1;
    const double qi_s = pi.w * P.scaling;                                     // ewald_spole.F90:114
    const bool coul_i = P.ew_on && !(fabs(qi_s) < ZERO_PLUS);                 // :117
    double fix = 0.0, fiy = 0.0, fiz = 0.0;
    const unsigned* row = nbr + (size_t)t * P.pitch;
    const int n = nnbr[t];
    for (int k = lane; k < n; k += 32) {
      const unsigned e = row[k];
      const int j = (int)(e & DLP_J_MASK);
      const bool halo = (e & DLP_F_HALO) != 0;
      const double w = P.half ? (halo ? ((e & DLP_F_ECNT) ? 1.0 : 0.0) : 1.0) : (halo ? ((e & DLP_F_ECNT) ? 1.0 : 0.0) : 0.5);
      const double4 pj = posq_s[j];
      const double xxt = pi.x - pj.x, yyt = pi.y - pj.y, zzt = pi.z - pj.z;   // two_body.F90:348-350
      const double rrr = sqrt(xxt * xxt + yyt * yyt + zzt * zzt);            // :351
      double gtx = 0.0, gty = 0.0, gtz = 0.0;   // pair force on i (sum of providers), for the HALF-mode scatter
      if (P.vdw_on) {
        const int kp = pair_k[ai * P.ntypes + (type_s[j] - 1)];
        if (kp >= 0 && rrr < P.rvdw) {                                        // vdw.F90:1892 / :1680
          const double r_rrr = 1.0 / rrr;
          const double rsq = rrr * rrr;
          const double r_rsq = r_rrr * r_rrr;
          double gamma, eng;
          if (!P.vdw_direct) {
            const double2* tb = vdw_tab + (size_t)kp * tstride;
            const int l = __double2int_rz(rrr * P.vdw_rdr);                   // :1909-1910
            const double ppp = rrr * P.vdw_rdr - (double)l;
            double2 a0 = tb[l], a1 = tb[l + 1], a2 = tb[l + 2];
            double gk = a0.x; if (l == 0) gk = gk * rrr;
            gamma = interp3(gk, a1.x, a2.x, ppp) * r_rsq;                     // :1914-1921
            eng = interp3(a0.y, a1.y, a2.y, ppp);                             // :1953-1960
            if (P.vdw_fshift) {                                               // :1922, :1962-1965
              const double2 c = tb[P.max_grid - 4];
              gamma = gamma - c.x * (r_rrr * P.r_rvdw);
              eng = eng + c.x * (rrr * P.r_rvdw - 1.0) - c.y;
            }
          } else {
            const double* pp = vdw_par + (size_t)kp * 10;
            double e0, g0;
            pot_direct(ltp[kp], pp, rrr, e0, g0);
            eng = e0 + pp[7] * rrr + pp[8];                                   // :1698-1699
            gamma = g0 * r_rsq - pp[7] * r_rrr;
          }
          const double f1 = gamma * xxt, f2 = gamma * yyt, f3 = gamma * zzt;
          fix = fix + f1; fiy = fiy + f2; fiz = fiz + f3;
          gtx += f1; gty += f2; gtz += f3;
          if (w != 0.0) {
            acc[0] += w * eng;
            acc[1] -= w * (gamma * rsq);
            acc[6] += w * (xxt * f1); acc[7] += w * (xxt * f2); acc[8] += w * (xxt * f3);
            acc[9] += w * (yyt * f2); acc[10] += w * (yyt * f3); acc[11] += w * (zzt * f3);
          }
        }
      }
      if (coul_i) {
        if (fabs(pj.w) > ZERO_PLUS && rrr < P.rcut) {                         // ewald_spole.F90:133
          const double prefac = qi_s * pj.w;
          const int l = __double2int_rz(rrr * P.ew_rdr);                      // :140-146
          const double diff = rrr * P.ew_rdr - (double)l;
          double2 a0 = ew_tab[l], a1 = ew_tab[l + 1], a2 = ew_tab[l + 2];
          double p1 = a0.x, q1 = a0.y;
          if (l == 0) { p1 = p1 * rrr; q1 = q1 * rrr; }
          const double erf_gamma = prefac * interp3(p1, a1.x, a2.x, diff);
          const double f1 = erf_gamma * xxt, f2 = erf_gamma * yyt, f3 = erf_gamma * zzt;
          fix = fix + f1; fiy = fiy + f2; fiz = fiz + f3;
          gtx += f1; gty += f2; gtz += f3;
          if (w != 0.0) {
            const double e_comp = prefac * interp3(q1, a1.y, a2.y, diff);     // :168-174
            acc[2] += w * e_comp;
            acc[3] -= w * (erf_gamma * (rrr * rrr));                          // :189
            acc[6] += w * (xxt * f1); acc[7] += w * (xxt * f2); acc[8] += w * (xxt * f3);
            acc[9] += w * (yyt * f2); acc[10] += w * (yyt * f3); acc[11] += w * (zzt * f3);
          }
        }
      }
      if (P.half && !halo) {   // Newton's third law: parts(jatm)%f -= f  (vdw.F90:1939-1941, ewald_spole.F90:159-161)
        atomicAdd(&fsx[j], -gtx); atomicAdd(&fsy[j], -gty); atomicAdd(&fsz[j], -gtz);
      }
    }
    // excluded pairs (two_body.F90:555-606 -> ewald_excl_forces)
    if (P.lbook && P.ew_on) {
      const int nx = nxnbr[t];
      if (nx > 0 && fabs(pi.w) > ZERO_PLUS) {                                 // ewald_spole.F90:541
        const double chgea = pi.w * P.scaling;
        const unsigned* xrow = xnbr + (size_t)t * P.xpitch;
        for (int k = lane; k < nx; k += 32) {
          const unsigned e = xrow[k];
          const int j = (int)(e & DLP_J_MASK);
          const bool halo = (e & DLP_F_HALO) != 0;
          const double w = P.half ? (halo ? ((e & DLP_F_ECNT) ? 1.0 : 0.0) : 1.0) : (halo ? ((e & DLP_F_ECNT) ? 1.0 : 0.0) : 0.5);
          const double4 pj = posq_s[j];
          const double xxt = pi.x - pj.x, yyt = pi.y - pj.y, zzt = pi.z - pj.z;
          const double rrr = sqrt(xxt * xxt + yyt * yyt + zzt * zzt);          // two_body.F90:576
          double chgprd = pj.w;
          if (fabs(chgprd) > ZERO_PLUS && rrr < P.rcut) {                     // :570
            const double a1 = 0.254829592, a2 = -0.284496736, a3 = 1.421413741, a4 = -1.453152027, a5 = 1.061405429,
                         pp = 0.3275911, r10 = 0.1, r216 = 1.0 / 216.0, r42 = 1.0 / 42.0, rr3 = 1.0 / 3.0;
            const double sqrpi = 1.7724538509055159;                          // Sqrt(pi), constants.F90:56
            chgprd = chgprd * chgea;
            const double rsq = rrr * rrr;
            const double alpr = rrr * P.alpha;
            const double alpr2 = alpr * alpr;
            double erfr, egamma;
            if (alpr < 1.0e-2) {                                              // :587-595
              erfr = 2.0 * chgprd * (P.alpha / sqrpi) * (1.0 + alpr2 * (-rr3 + alpr2 * (r10 + alpr2 * (-r42 + alpr2 * r216))));
              egamma = -4.0 * chgprd * ((P.alpha * (P.alpha * P.alpha)) / sqrpi) *
                       (rr3 + alpr2 * (-2.0 * r10 + alpr2 * (3.0 * r42 - 4.0 * alpr2 * r216)));
            } else {                                                          // :601-607
              const double ar = P.alpha * rrr;
              const double exp1 = exp(-(ar * ar));
              const double tt = 1.0 / (1.0 + pp * P.alpha * rrr);
              erfr = chgprd * (1.0 - tt * (a1 + tt * (a2 + tt * (a3 + tt * (a4 + tt * a5)))) * exp1) / rrr;
              egamma = -(erfr - 2.0 * chgprd * (P.alpha / sqrpi) * exp1) / rsq;
            }
            const double f1 = egamma * xxt, f2 = egamma * yyt, f3 = egamma * zzt;
            fix = fix + f1; fiy = fiy + f2; fiz = fiz + f3;
            if (P.half && !halo) { atomicAdd(&fsx[j], -f1); atomicAdd(&fsy[j], -f2); atomicAdd(&fsz[j], -f3); }
            if (w != 0.0) {
              acc[4] -= w * erfr;
              acc[5] -= w * (egamma * rsq);
              acc[6] += w * (xxt * f1); acc[7] += w * (xxt * f2); acc[8] += w * (xxt * f3);
              acc[9] += w * (yyt * f2); acc[10] += w * (yyt * f3); acc[11] += w * (zzt * f3);
            }
          }
        }
      }
    }
    // force on atom i: warp shuffle reduction, lane 0 commits
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      fix += __shfl_xor_sync(DLP_FULL, fix, d);
      fiy += __shfl_xor_sync(DLP_FULL, fiy, d);
      fiz += __shfl_xor_sync(DLP_FULL, fiz, d);
    }
    if (lane == 0) {
      if (P.half) {
        atomicAdd(&fsx[ii], fix); atomicAdd(&fsy[ii], fiy); atomicAdd(&fsz[ii], fiz);
      } else {
        const int i = at_list[ii];
        if (P.zero_forces) { fx[i] = fix; fy[i] = fiy; fz[i] = fiz; }
        else { fx[i] += fix; fy[i] += fiy; fz[i] += fiz; }
      }
    }
  }
  // energies / virials / stress: warp shuffle, then per-block in a fixed order
  __shared__ double red[8][12];
#pragma unroll
  for (int k = 0; k < 12; ++k) {
    double v = acc[k];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(DLP_FULL, v, d);
    if (lane == 0) red[warp][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 12) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
    partial[(size_t)blockIdx.x * 12 + threadIdx.x] = v;
  }
}

__global__ void k_final_reduce(int nblocks, const double* __restrict__ partial, double* __restrict__ out) {
  int k = threadIdx.x;
  if (k >= 12) return;
  double v = 0.0;
  for (int b = 0; b < nblocks; ++b) v += partial[(size_t)b * 12 + k];
  if (k < 6) out[k] = v;
  else {
    // strs1,2,3,5,6,9 -> stress(1:9) symmetric (vdw.F90:2014-2022)
    const int map1[6] = {0, 1, 2, 4, 5, 8};
    const int map2[6] = {-1, 3, 6, -1, 7, -1};
    out[6 + map1[k - 6]] = v;
    if (map2[k - 6] >= 0) out[6 + map2[k - 6]] = v;
  }
  if (k == 0) out[15] = 0.0;
}

__global__ void k_scatter_half(int nlast, int natms, int zero_forces, const int* __restrict__ at_list, const double* __restrict__ fsx,
                               const double* __restrict__ fsy, const double* __restrict__ fsz, double* fx, double* fy, double* fz) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nlast) return;
  int i = at_list[s];
  if (i >= natms) return;
  if (zero_forces) { fx[i] = fsx[s]; fy[i] = fsy[s]; fz[i] = fsz[s]; }
  else { fx[i] += fsx[s]; fy[i] += fsy[s]; fz[i] += fsz[s]; }
}

// nve.F90:163-173, :198-217
__global__ void k_vv(int natms, int stage, double dt, const int* __restrict__ lsite, const double* __restrict__ weight_site,
                     double4* __restrict__ posq, double* vx, double* vy, double* vz, const double* __restrict__ fx,
                     const double* __restrict__ fy, const double* __restrict__ fz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natms) return;
  double hstep = 0.5 * dt;
  double rm = 1.0 / weight_site[lsite[i] - 1];
  double tmp = hstep * rm;
  double a = vx[i] + tmp * fx[i], b = vy[i] + tmp * fy[i], c = vz[i] + tmp * fz[i];
  vx[i] = a; vy[i] = b; vz[i] = c;
  if (stage == 1) {
    double4 p = posq[i];
    p.x = p.x + dt * a; p.y = p.y + dt * b; p.z = p.z + dt * c;
    posq[i] = p;
  }
}

// DFMA throughput probe: 8 independent chains per thread, FMA allowed here on purpose
__global__ void k_dfma(int iters, double* out) {
  double a0 = threadIdx.x * 1e-9, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000000001, c = 1e-12;
  for (int i = 0; i < iters; ++i) {
    a0 = __fma_rn(a0, m, c); a1 = __fma_rn(a1, m, c); a2 = __fma_rn(a2, m, c); a3 = __fma_rn(a3, m, c);
    a4 = __fma_rn(a4, m, c); a5 = __fma_rn(a5, m, c); a6 = __fma_rn(a6, m, c); a7 = __fma_rn(a7, m, c);
  }
  double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (s == 123.456) out[0] = s;
}

}  // namespace

int dlp_two_body(dlpgpu_ctx* ctx, int zero_forces, double out[16]) {
  cudaStream_t s = ctx->stream;
  if (!ctx->list_valid) return dlp_fail(ctx, DLPGPU_ERR_STATE, "two_body: no valid neighbour list");
  if (ctx->natms != ctx->list_natms || ctx->nlast != ctx->list_nlast)
    return dlp_fail(ctx, DLPGPU_ERR_HALO_COUNT, "two_body: atom counts changed since the list build");
  const int natms = ctx->natms;
  cudaEventRecord(ctx->ev[4], s);
  CKRC(dlp_gather_sorted(ctx));
  FParams P{};
  P.natms = natms; P.pitch = ctx->pitch; P.xpitch = ctx->xpitch > 0 ? ctx->xpitch : 1; P.ntypes = std::max(ctx->ntypes, 1);
  P.max_grid = ctx->max_grid; P.max_vdw = ctx->max_vdw; P.ew_n = ctx->ew_n;
  P.vdw_on = ctx->vdw_on; P.vdw_direct = ctx->vdw_direct; P.vdw_fshift = ctx->vdw_fshift; P.ew_on = ctx->ew_on;
  P.half = ctx->force_mode == 1; P.zero_forces = zero_forces; P.lbook = ctx->lbook;
  P.rvdw = ctx->rvdw; P.r_rvdw = ctx->rvdw > 0 ? 1.0 / ctx->rvdw : 0.0; P.vdw_rdr = ctx->vdw_rdr; P.rcut = ctx->rcut;
  P.ew_rdr = ctx->ew_rdr; P.alpha = ctx->alpha; P.scaling = ctx->scaling;
  // shared-memory tables when they fit
  size_t nv = (ctx->vdw_on && !ctx->vdw_direct) ? (size_t)ctx->max_vdw * (ctx->max_grid + 1) : 0;
  size_t ne = ctx->ew_on ? (size_t)ctx->ew_n + 1 : 0;
  size_t smem = (nv + ne) * sizeof(double2) + (size_t)P.ntypes * P.ntypes * sizeof(int);
  bool use_smem = smem <= 200 * 1024;
  const int threads = 256;
  int bps = use_smem ? std::max(1, std::min(4, (int)((220 * 1024) / std::max(smem, (size_t)1)))) : 4;
  int blocks = std::max(1, std::min(cdiv(natms, threads / 32), ctx->sm_count * bps));
  CK(ctx->partial.ensure((size_t)blocks * 12 + 16, s));
  double *fsx = nullptr, *fsy = nullptr, *fsz = nullptr;
  if (P.half) {   // sorted-slot force accumulators of the atomics path
    CK(ctx->fsx.ensure(ctx->nlast + 1, s)); CK(ctx->fsy.ensure(ctx->nlast + 1, s)); CK(ctx->fsz.ensure(ctx->nlast + 1, s));
    CK(cudaMemsetAsync(ctx->fsx.p, 0, (size_t)ctx->nlast * sizeof(double), s));
    CK(cudaMemsetAsync(ctx->fsy.p, 0, (size_t)ctx->nlast * sizeof(double), s));
    CK(cudaMemsetAsync(ctx->fsz.p, 0, (size_t)ctx->nlast * sizeof(double), s));
    fsx = ctx->fsx.p; fsy = ctx->fsy.p; fsz = ctx->fsz.p;
  }
  cudaEventRecord(ctx->ev[6], s);
  if (natms > 0) {
    if (use_smem) {
      CK(cudaFuncSetAttribute(k_pair_forces<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      LAUNCH(ctx, k_pair_forces<true>, blocks, threads, smem, P, ctx->loc_slot.p, ctx->at_list.p, ctx->posq_s.p, ctx->type_s.p,
             ctx->nbr.p, ctx->nnbr.p, ctx->xnbr.p, ctx->nxnbr.p, ctx->pair_k.p, ctx->ltp.p, ctx->vdw_tab.p, ctx->vdw_par.p,
             ctx->ew_tab.p, ctx->fx.p, ctx->fy.p, ctx->fz.p, fsx, fsy, fsz, ctx->partial.p);
    } else {
      LAUNCH(ctx, k_pair_forces<false>, blocks, threads, 0, P, ctx->loc_slot.p, ctx->at_list.p, ctx->posq_s.p, ctx->type_s.p,
             ctx->nbr.p, ctx->nnbr.p, ctx->xnbr.p, ctx->nxnbr.p, ctx->pair_k.p, ctx->ltp.p, ctx->vdw_tab.p, ctx->vdw_par.p,
             ctx->ew_tab.p, ctx->fx.p, ctx->fy.p, ctx->fz.p, fsx, fsy, fsz, ctx->partial.p);
    }
  }
  cudaEventRecord(ctx->ev[7], s);
  if (natms > 0 && P.half)
    LAUNCH(ctx, k_scatter_half, cdiv(ctx->nlast, 256), 256, 0, ctx->nlast, natms, zero_forces, ctx->at_list.p, fsx, fsy, fsz, ctx->fx.p,
           ctx->fy.p, ctx->fz.p);
  LAUNCH(ctx, k_final_reduce, 1, 32, 0, natms > 0 ? blocks : 0, ctx->partial.p, ctx->out_dev.p);
  cudaEventRecord(ctx->ev[5], s);
  if (out) {
    CK(cudaMemcpyAsync(out, ctx->out_dev.p, 16 * sizeof(double), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]); ctx->t_force = ms;
    cudaEventElapsedTime(&ms, ctx->ev[6], ctx->ev[7]); ctx->t_pair = ms;
  }
  CK(cudaGetLastError());
  return 0;
}

extern "C" {

int dlpgpu_dev_two_body_forces(dlpgpu_ctx* ctx, int zero_forces, double out[16]) {
  if (!ctx) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  return dlp_two_body(ctx, zero_forces, out);
}

int dlpgpu_dev_vv(dlpgpu_ctx* ctx, int stage, double dt) {
  if (!ctx || (stage != 1 && stage != 2)) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (ctx->nsites < 1) return dlp_fail(ctx, DLPGPU_ERR_STATE, "vv: sites not set");
  if (ctx->natms > 0)
    LAUNCH(ctx, k_vv, cdiv(ctx->natms, 256), 256, 0, ctx->natms, stage, dt, ctx->lsite.p, ctx->weight_site.p, ctx->posq.p, ctx->vx.p,
           ctx->vy.p, ctx->vz.p, ctx->fx.p, ctx->fy.p, ctx->fz.p);
  return 0;
}

int dlpgpu_fp64_peak(dlpgpu_ctx* ctx, double seconds, double* tflops) {
  if (!ctx || !tflops) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  cudaStream_t s = ctx->stream;
  const int iters = 20000, threads = 256, blocks = ctx->sm_count * 8;
  LAUNCH(ctx, k_dfma, blocks, threads, 0, 1000, ctx->out_dev.p);   // warm-up
  CK(cudaStreamSynchronize(s));
  double best = 0.0, spent = 0.0;
  int reps = 0;
  while (spent < seconds * 1000.0 || reps < 3) {
    cudaEventRecord(ctx->ev[4], s);
    LAUNCH(ctx, k_dfma, blocks, threads, 0, iters, ctx->out_dev.p);
    cudaEventRecord(ctx->ev[5], s);
    CK(cudaStreamSynchronize(s));
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->ev[4], ctx->ev[5]);
    double fl = 2.0 * 8.0 * (double)iters * threads * blocks;
    best = std::max(best, fl / (ms * 1e-3) / 1e12);
    spent += ms;
    if (++reps > 200) break;
  }
  *tflops = best;
  return 0;
}

}  // extern "C"
