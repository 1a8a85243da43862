// =============================================================================
// dlp_oracle.cpp -- TEST INFRASTRUCTURE ONLY (not product code).
//
// CPU restatement ("port") of DL_POLY 5.1.0's short-range two-body hot path, written
// from the reference's Fortran sources statement-for-statement in operation order so
// that IEEE-754 double results are reproducible.  Compile with
//     g++ -O2 -ffp-contract=off -fno-fast-math      (see oracle/Makefile)
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this library.  The product (libdlpgpu.so) never links or calls it.
//
// Parity pin status (also stated in DESIGN.md):
//   * two_body_potentials + vdw_forces_direct : PINNED against the reference's own
//     known-answer vectors source/unit_tests/test_vdw.F90:46-59 (24 potentials).
//   * link_cell_pairs, vdw_forces_tab, ewald_real_forces_coul, ewald_excl_forces, halo /
//     deport, vnl_check : "parity unpinned" by the reference (it has no unit vectors and
//     its regression inputs are downloaded at build time, CMakeLists.txt:331-333; no
//     Fortran compiler exists in this image).  They are cross-checked here by an
//     independent O(N^2) minimum-image brute force (ora_brute_*, long double sums).
//
// Every routine cites the reference file:line it follows (paths under source/).
// Arrays keep the reference's 1-based indexing internally (slot 0 unused) to keep the
// restatement auditable against the Fortran.
// =============================================================================
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------- constants.F90
// constants.F90:53-58 (pi, sqrpi, rsqrpi), :100 (r4pie0), :189-204 (zero_plus, half_*, smalldr)
const double pi = 4.0 * std::atan(1.0);
const double sqrpi = std::sqrt(pi);
const double rsqrpi = 1.0 / std::sqrt(pi);
const double r4pie0 = 138935.4835e0;
const double zero_plus = DBL_MIN;                             // Tiny(1.0_wp)
const double half_plus = std::nextafter(0.5, 1.0);            // Nearest(0.5,+1)
const double half_minus = std::nextafter(0.5, 0.0);           // Nearest(0.5,-1)
const double smalldr = 1.0e-6;
const double delr_max = 0.01;                                 // constants.F90:139

// particle.F90:14-20  corePart, Sequence, 64 bytes
struct CorePart {
  double xxx, yyy, zzz, fxx, fyy, fzz, chge;
  int32_t pad1, pad2;
};
static_assert(sizeof(CorePart) == 64, "corePart must be 64 bytes");

inline int f_int(double x) { return (int)x; }                 // Fortran Int(): truncate
inline int f_nint(double x) { return (int)std::lround(x); }   // Fortran Nint(): half away from zero
inline double f_anint(double x) { return std::round(x); }     // Fortran Anint()

// libgcc __powidf2 -- what gfortran emits for real**integer
inline double powi(double x, int n) {
  unsigned m = (n < 0) ? (unsigned)(-n) : (unsigned)n;
  double y = (m & 1u) ? x : 1.0;
  while (m >>= 1) {
    x = x * x;
    if (m & 1u) y = y * x;
  }
  return n < 0 ? 1.0 / y : y;
}

// ---------------------------------------------------------------- numerics.F90
// numerics.F90:1344-1446  dcell
void dcell(const double* a /*1..9*/, double* b /*1..10*/) {
  const double* aaa = a;
  b[1] = std::sqrt(aaa[1] * aaa[1] + aaa[2] * aaa[2] + aaa[3] * aaa[3]);
  b[2] = std::sqrt(aaa[4] * aaa[4] + aaa[5] * aaa[5] + aaa[6] * aaa[6]);
  b[3] = std::sqrt(aaa[7] * aaa[7] + aaa[8] * aaa[8] + aaa[9] * aaa[9]);
  b[4] = (aaa[1] * aaa[4] + aaa[2] * aaa[5] + aaa[3] * aaa[6]) / (b[1] * b[2]);
  b[5] = (aaa[1] * aaa[7] + aaa[2] * aaa[8] + aaa[3] * aaa[9]) / (b[1] * b[3]);
  b[6] = (aaa[4] * aaa[7] + aaa[5] * aaa[8] + aaa[6] * aaa[9]) / (b[2] * b[3]);
  double axb1 = aaa[2] * aaa[6] - aaa[3] * aaa[5];
  double axb2 = aaa[3] * aaa[4] - aaa[1] * aaa[6];
  double axb3 = aaa[1] * aaa[5] - aaa[2] * aaa[4];
  double bxc1 = aaa[5] * aaa[9] - aaa[6] * aaa[8];
  double bxc2 = aaa[6] * aaa[7] - aaa[4] * aaa[9];
  double bxc3 = aaa[4] * aaa[8] - aaa[5] * aaa[7];
  double cxa1 = aaa[8] * aaa[3] - aaa[9] * aaa[2];
  double cxa2 = aaa[9] * aaa[1] - aaa[7] * aaa[3];
  double cxa3 = aaa[7] * aaa[2] - aaa[8] * aaa[1];
  b[10] = std::fabs(aaa[1] * bxc1 + aaa[2] * bxc2 + aaa[3] * bxc3);
  double d[4], x[4], y[4];
  d[1] = b[10] / std::sqrt(bxc1 * bxc1 + bxc2 * bxc2 + bxc3 * bxc3);
  d[2] = b[10] / std::sqrt(cxa1 * cxa1 + cxa2 * cxa2 + cxa3 * cxa3);
  d[3] = b[10] / std::sqrt(axb1 * axb1 + axb2 * axb2 + axb3 * axb3);
  x[1] = std::fabs(aaa[1]) / b[1]; y[1] = std::fabs(aaa[2]) / b[1];
  x[2] = std::fabs(aaa[4]) / b[2]; y[2] = std::fabs(aaa[5]) / b[2];
  x[3] = std::fabs(aaa[7]) / b[3]; y[3] = std::fabs(aaa[8]) / b[3];
  if (x[1] >= x[2] && x[1] >= x[3]) {
    b[7] = d[1];
    if (y[2] >= y[3]) { b[8] = d[2]; b[9] = d[3]; } else { b[8] = d[3]; b[9] = d[2]; }
  } else if (x[2] >= x[1] && x[2] >= x[3]) {
    b[7] = d[2];
    if (y[1] >= y[3]) { b[8] = d[1]; b[9] = d[3]; } else { b[8] = d[3]; b[9] = d[1]; }
  } else {
    b[7] = d[3];
    if (y[1] >= y[2]) { b[8] = d[1]; b[9] = d[2]; } else { b[8] = d[2]; b[9] = d[1]; }
  }
}

// numerics.F90:1448-1509  invert
void invert(const double* a /*1..9*/, double* b /*1..9*/, double& d) {
  b[1] = a[5] * a[9] - a[6] * a[8];
  b[2] = a[3] * a[8] - a[2] * a[9];
  b[3] = a[2] * a[6] - a[3] * a[5];
  b[4] = a[6] * a[7] - a[4] * a[9];
  b[5] = a[1] * a[9] - a[3] * a[7];
  b[6] = a[3] * a[4] - a[1] * a[6];
  b[7] = a[4] * a[8] - a[5] * a[7];
  b[8] = a[2] * a[7] - a[1] * a[8];
  b[9] = a[1] * a[5] - a[2] * a[4];
  d = a[1] * b[1] + a[4] * b[2] + a[7] * b[3];
  double r = 0.0;
  if (std::fabs(d) > 0.0) r = 1.0 / d;
  for (int i = 1; i <= 9; ++i) b[i] = r * b[i];
}

// numerics.F90:1511-1600  images (imcon 1,2(0),3 only; 4,5,7 are "NOT AVAILABLE in DD")
void images(int imcon, const double* cell, int pairs, double* xxx, double* yyy, double* zzz /*1-based*/) {
  if (imcon == 1) {
    double aaa = 1.0 / cell[1];
    for (int i = 1; i <= pairs; ++i) {
      xxx[i] = xxx[i] - cell[1] * f_anint(aaa * xxx[i]);
      yyy[i] = yyy[i] - cell[1] * f_anint(aaa * yyy[i]);
      zzz[i] = zzz[i] - cell[1] * f_anint(aaa * zzz[i]);
    }
  } else if (imcon == 2 || imcon == 0) {
    double aaa = 1.0 / cell[1], bbb = 1.0 / cell[5], ccc = 1.0 / cell[9];
    for (int i = 1; i <= pairs; ++i) {
      xxx[i] = xxx[i] - cell[1] * f_anint(aaa * xxx[i]);
      yyy[i] = yyy[i] - cell[5] * f_anint(bbb * yyy[i]);
      zzz[i] = zzz[i] - cell[9] * f_anint(ccc * zzz[i]);
    }
  } else if (imcon == 3) {
    double rcell[10], det;
    invert(cell, rcell, det);
    for (int i = 1; i <= pairs; ++i) {
      double xss = rcell[1] * xxx[i] + rcell[4] * yyy[i] + rcell[7] * zzz[i];
      double yss = rcell[2] * xxx[i] + rcell[5] * yyy[i] + rcell[8] * zzz[i];
      double zss = rcell[3] * xxx[i] + rcell[6] * yyy[i] + rcell[9] * zzz[i];
      xss = xss - f_anint(xss); yss = yss - f_anint(yss); zss = zss - f_anint(zss);
      xxx[i] = cell[1] * xss + cell[4] * yss + cell[7] * zss;
      yyy[i] = cell[2] * xss + cell[5] * yss + cell[8] * zss;
      zzz[i] = cell[3] * xss + cell[6] * yss + cell[9] * zss;
    }
  }
}

// numerics.F90:1851-1950  pbcshift_parts (imcon 1,2(0),3)
void pbcshift(int imcon, const double* cell, int natms, CorePart* parts /*1-based*/) {
  if (imcon == 1) {
    double aaa = 1.0 / cell[1];
    for (int i = 1; i <= natms; ++i) {
      double xss = aaa * parts[i].xxx, yss = aaa * parts[i].yyy, zss = aaa * parts[i].zzz;
      xss = xss - f_anint(xss); if (xss >= half_minus) xss = -xss;
      yss = yss - f_anint(yss); if (yss >= half_minus) yss = -yss;
      zss = zss - f_anint(zss); if (zss >= half_minus) zss = -zss;
      parts[i].xxx = cell[1] * xss; parts[i].yyy = cell[1] * yss; parts[i].zzz = cell[1] * zss;
    }
  } else if (imcon == 2 || imcon == 0) {
    double aaa = 1.0 / cell[1], bbb = 1.0 / cell[5], ccc = 1.0 / cell[9];
    for (int i = 1; i <= natms; ++i) {
      double xss = aaa * parts[i].xxx, yss = bbb * parts[i].yyy, zss = ccc * parts[i].zzz;
      xss = xss - f_anint(xss); if (xss >= half_minus) xss = -xss;
      yss = yss - f_anint(yss); if (yss >= half_minus) yss = -yss;
      zss = zss - f_anint(zss); if (zss >= half_minus) zss = -zss;
      parts[i].xxx = cell[1] * xss; parts[i].yyy = cell[5] * yss; parts[i].zzz = cell[9] * zss;
    }
  } else if (imcon == 3) {
    double rcell[10], det;
    invert(cell, rcell, det);
    for (int i = 1; i <= natms; ++i) {
      double xss = rcell[1] * parts[i].xxx + rcell[4] * parts[i].yyy + rcell[7] * parts[i].zzz;
      double yss = rcell[2] * parts[i].xxx + rcell[5] * parts[i].yyy + rcell[8] * parts[i].zzz;
      double zss = rcell[3] * parts[i].xxx + rcell[6] * parts[i].yyy + rcell[9] * parts[i].zzz;
      xss = xss - f_anint(xss); if (xss >= half_minus) xss = -xss;
      yss = yss - f_anint(yss); if (yss >= half_minus) yss = -yss;
      zss = zss - f_anint(zss); if (zss >= half_minus) zss = -zss;
      parts[i].xxx = cell[1] * xss + cell[4] * yss + cell[7] * zss;
      parts[i].yyy = cell[2] * xss + cell[5] * yss + cell[8] * zss;
      parts[i].zzz = cell[3] * xss + cell[6] * yss + cell[9] * zss;
    }
  }
}

// numerics.F90:1048-1098  match (binary search in ascending list(1:ind_top))
bool match(int n, int ind_top, const int* list /*1-based*/) {
  if (ind_top < 1) return false;
  int ind_old = 1, ind_now = 1;
  for (;;) {
    if (n == list[ind_now]) {
      return true;
    } else if (n > list[ind_now]) {
      if (ind_old == ind_top) return false;
      ind_old = ind_now;
      ind_now = (ind_old + ind_top + 1) / 2;
    } else {
      ind_now = (ind_old + ind_now) / 2;
      if (ind_now == ind_old) return false;
    }
  }
}

// numerics.F90:3888-3893 equal_real_wp: Abs(a-b) < epsilon_wp (constants.F90:200 Epsilon(1.0_wp))
inline bool f_equal(double a, double b) { return std::fabs(a - b) < DBL_EPSILON; }

// numerics.F90:3647-3683  calc_erfc_n / calc_erfc_deriv_n (Abramowitz-Stegun 7.1.26)
inline double calc_erfc(double x) {
  const double a1 = 0.254829592, a2 = -0.284496736, a3 = 1.421413741, a4 = -1.453152027,
               a5 = 1.061405429, pp = 0.3275911;
  double tt = 1.0 / (1.0 + pp * x);
  return tt * (a1 + tt * (a2 + tt * (a3 + tt * (a4 + tt * a5)))) * std::exp(-(x * x));
}
inline double calc_erfc_deriv(double x) { return 2.0 * std::exp(-(x * x)) * rsqrpi; }

// ---------------------------------------------------------------- two_body_potentials.F90
struct EG { double energy, gamma; };

// keypot numbering = vdw.F90:64-117 (VDW_* constants)
EG pot_energy(int key, const double* p /*1..7*/, double r);

EG lj126(const double* p, double r) {            // two_body_potentials.F90:307-317
  double r_6 = powi(1.0 / r, 6);
  return {(p[1] * r_6 - p[2]) * r_6, 6.0 * r_6 * (2.0 * p[1] * r_6 - p[2])};
}
EG lj(double eps, double sig, double r) {        // :260-270
  double sor6 = powi(sig / r, 6);
  return {4.0 * eps * sor6 * (sor6 - 1.0), 24.0 * eps * sor6 * (2.0 * sor6 - 1.0)};
}
EG buck(double A, double rho, double C, double r) {   // :471-485
  double b = r / rho;
  double t1 = A * std::exp(-b);
  double t2 = -C / powi(r, 6);
  return {t1 + t2, t1 * b + 6.0 * t2};
}
EG morse(double e0, double r0, double k, double r) {  // :424-435
  double t = std::exp(-k * (r - r0));
  return {e0 * (powi(1.0 - t, 2) - 1.0), -2.0 * r * e0 * k * (1.0 - t) * t};
}
EG zbl(double k, double ia, double r) {               // :718-743
  const double zb[5] = {0, 0.18175, 0.50986, 0.28022, 0.02817};
  const double zc[5] = {0, 3.1998, 0.94229, 0.40290, 0.20162};
  const double zbl_ab = 0.52917721067;
  double a = (std::pow(k, 0.23) + std::pow(ia, 0.23)) / (zbl_ab * 0.88534);
  double kk = k * ia * r4pie0;
  double e = 0.0, g = 0.0;
  double x = r * a, ir = 1.0 / r;
  for (int i = 1; i <= 4; ++i) {
    double t1 = zb[i] * std::exp(-x * zc[i]);
    e = e + t1;
    g = g - zc[i] * t1;
  }
  e = kk * e * ir;
  g = e - a * kk * g;
  return {e, g};
}
EG fm(double rm, double ic, double r) {               // :754-775
  double c = 1.0 / ic;
  if (r < rm) {
    double t = std::exp(-(rm - r) * c) * 0.5;
    return {1.0 - t, r * c * t};
  } else {
    double t = std::exp(-(r - rm) * c) * 0.5;
    return {t, r * c * t};
  }
}
EG mdf(double ri, double rc, double r) {              // MDF_energy
  if (r < ri) return {1.0, 0.0};
  if (r > rc) return {0.0, 0.0};
  double rci = powi(rc - ri, 5);
  double e = powi(rc - r, 3) *
             (10.0 * powi(ri, 2) - 5.0 * rc * ri - 15.0 * r * ri + powi(rc, 2) + 3.0 * r * rc + 6 * powi(r, 2)) / rci;
  double g = 30.0 * r * powi(r - rc, 2) * powi(r - ri, 2) / rci;
  return {e, g};
}

EG pot_energy(int key, const double* p, double r) {
  switch (key) {
    case 1: return lj126(p, r);
    case 2: return lj(p[1], p[2], r);
    case 3: {  // n_m :330-345
      double a = p[4] / r, b = 1.0 / (p[2] - p[3]);
      double r_n = powi(a, (int)p[2]), r_m = powi(a, (int)p[3]);
      return {p[1] * (p[3] * r_n - p[2] * r_m) * b, p[1] * p[3] * p[2] * (r_n - r_m) * b};
    }
    case 4: return buck(p[1], p[2], p[3], r);
    case 5: {  // bhm :499-514  a,b,sig,c,d
      double r_inv_2 = powi(r, -2);
      double t1 = p[1] * std::exp(p[2] * (p[3] - r));
      double t2 = -p[4] * powi(r_inv_2, 3);
      double t3 = -p[5] * powi(r_inv_2, 4);
      return {t1 + t2 + t3, (t1 * r * p[2] + 6.0 * t2 + 8.0 * t3)};
    }
    case 6: {  // hbond 12-10
      double r_inv_2 = powi(r, -2);
      double fac12 = p[1] * powi(r_inv_2, 6), fac10 = -p[2] * powi(r_inv_2, 5);
      return {fac12 + fac10, (12.0 * fac12 + 10.0 * fac10)};
    }
    case 7: {  // nm_shift :362-407  e0,n,m,r0,r_trunc
      double e0 = p[1], n = p[2], m = p[3], r0 = p[4], rt = p[5];
      if (r <= rt) {
        double r_inv = powi(r, -1);
        int n_int = f_nint(n), m_int = f_nint(m);
        double t = n - m, b = 1.0 / t, c = rt / r0, c_inv = r0 / rt;
        double beta = c * std::pow((powi(c, m_int + 1) - 1.0) / (powi(c, n_int + 1) - 1.0), b);
        double alpha = -t / (m * powi(beta, n_int) * (1.0 + (n * c_inv - n - 1.0) * powi(c_inv, n_int)) -
                             n * powi(beta, m_int) * (1.0 + (m * c_inv - m - 1.0) * powi(c_inv, m_int)));
        double e1 = e0 * alpha;
        double a = r0 * r_inv;
        double e = e1 * (m * powi(beta, n_int) * (powi(a, n_int) - powi(1.0 * c_inv, n_int)) -
                         n * powi(beta, m_int) * (powi(a, m_int) - powi(1.0 * c_inv, m_int)) +
                         n * m * ((r / rt - 1.0) * (powi(beta * c_inv, n_int) - powi(beta * c_inv, m_int)))) * b;
        double g = e1 * m * n * (powi(beta, n_int) * powi(a, n_int) - powi(beta, m_int) * powi(a, m_int) -
                                 r / rt * (powi(beta * c_inv, n_int) - powi(beta * c_inv, m_int))) * b;
        return {e, g};
      }
      return {0.0, 0.0};
    }
    case 8: return morse(p[1], p[2], p[3], r);
    case 9: {  // wca  eps,sig,d,cut
      if (r < p[4] || std::fabs(r - p[3]) < 1.0e-10) {
        double s6 = powi(p[2] / (r - p[3]), 6);
        return {4.0 * p[1] * s6 * (s6 - 1.0) + p[1], 24.0 * p[1] * s6 * (2.0 * s6 - 1.0) * r / (r - p[3])};
      }
      return {0.0, 0.0};
    }
    case 10: {  // dpd a,rc
      if (r < p[2]) {
        double t2 = r / p[2];
        double t1 = 0.5 * p[1] * p[2] * (1.0 - t2);
        return {t1 * (1.0 - t2), 2.0 * t1 * t2};
      }
      return {0.0, 0.0};
    }
    case 11: {  // amoeba eps,sig
      double rho = r / p[2];
      double t1 = 1.0 / (0.07 + rho);
      double t2 = 1.0 / (0.12 + powi(rho, 7));
      double t3 = p[1] * powi(1.07 * t1, 7);
      double t = t3 * ((1.12 * t2) - 2.0);
      return {t, 7.0 * (t1 * t + 1.12 * t3 * powi(t2, 2) * powi(rho, 6)) * rho};
    }
    case 12: {  // lj_coh eps,sig,coh
      double sor6 = powi(p[2] / r, 6);
      return {4.0 * p[1] * sor6 * (sor6 - p[3]), 24.0 * p[1] * sor6 * (2.0 * sor6 - p[3])};
    }
    case 13: {  // morse12 e0,r0,kk,c
      double t1 = std::exp(-p[3] * (r - p[2]));
      double t2 = p[4] * powi(r, -12);
      return {p[1] * t1 * (t1 - 2.0) + t2, -2.0 * r * p[1] * p[3] * (1.0 - t1) * t1 + 12.0 * t2};
    }
    case 14: {  // rydberg a,b,c
      double kk = r / p[3];
      double t1 = std::exp(-kk);
      return {(p[1] + p[2] * r) * t1, kk * t1 * (p[1] - p[2] * p[3] + p[2] * r)};
    }
    case 15: return zbl(p[1], p[2], r);
    case 16: {  // zbls
      EG z = zbl(p[1], p[2], r), f = fm(p[3], p[4], r), m = morse(p[5], p[6], p[7], r);
      return {f.energy * z.energy + (1.0 - f.energy) * m.energy,
              f.energy * z.gamma + f.gamma * z.energy + (1.0 - f.energy) * m.gamma - f.gamma * m.energy};
    }
    case 17: {  // zblb
      EG z = zbl(p[1], p[2], r), f = fm(p[3], p[4], r), b = buck(p[5], p[6], p[7], r);
      return {f.energy * z.energy + (1.0 - f.energy) * b.energy,
              f.energy * z.gamma + f.gamma * z.energy + (1.0 - f.energy) * b.gamma - f.gamma * b.energy};
    }
    case 18: {  // mlj  eps,sig,ri,rc
      EG l = lj(p[1], p[2], r), m = mdf(p[3], p[4], r);
      return {l.energy * m.energy, l.gamma * m.energy + m.gamma * l.energy};
    }
    case 19: {  // mbuck A,rho,C,ri,rc
      EG b = buck(p[1], p[2], p[3], r), m = mdf(p[4], p[5], r);
      return {b.energy * m.energy, b.gamma * m.energy + m.gamma * b.energy};
    }
    case 20: {  // mlj126 a,b,ri,rc
      EG l = lj126(p, r), m = mdf(p[3], p[4], r);
      return {l.energy * m.energy, l.gamma * m.energy + m.gamma * l.energy};
    }
    case 21: {  // ljf ea,sig2,rc2
      double r2 = r * r;
      if (r2 > p[3]) return {0.0, 0.0};
      double ir = 1.0 / r2, st = p[2] * ir, rct = p[3] * ir;
      double x = p[1] * powi(rct - 1.0, 2);
      return {x * (st - 1.0), 4.0 * p[1] * rct * (rct - 1.0) * (st - 1.0) + 2.0 * x * st};
    }
    case 22: {  // sanderson A,L,d
      double b = std::pow((r - p[2]) / p[3], 2.0);
      double t = p[1] * std::exp(-b);
      return {-t, -2.0 * (r - p[2]) * r * t / std::pow(p[3], 2.0)};
    }
    case 23: {  // ndpd a,b,n,rc
      if (r < p[4]) {
        double t2 = r / p[4];
        double t1 = p[1] * p[4] * (1.0 - t2);
        double t0 = p[2] * std::pow(1.0 - t2, p[3] - 1.0);
        return {t1 * (1.0 - t2) * (t0 / (p[3] + 1.0) - 0.5), t1 * t2 * (t0 - 1.0)};
      }
      return {0.0, 0.0};
    }
    case 24: {  // sw eps,A,B,sig,p,q,aa
      double e = p[4] / (r - p[7] * p[4]);
      if (r < p[7] * p[4]) {
        double p_r = p[4] / r;
        double c = p[3] * std::pow(p_r, p[5]);
        double exp_e = std::exp(e);
        double t = p[2] * p[1] * (c - std::pow(p_r, p[6])) * exp_e;
        return {t, p[2] * p[1] * (p[5] * c - p[6] * std::pow(p_r, p[6])) * exp_e + t * r * e / (r - p[7] * p[4])};
      }
      return {0.0, 0.0};
    }
    default: return {0.0, 0.0};
  }
}

// ---------------------------------------------------------------- force-field container
struct Vdw {
  int ntypes = 0, n_vdw = 0, max_vdw = 0, max_grid = 0, max_param = 7;
  std::vector<int> list;            // list(1:ntab) key -> k          (vdw.F90 vdws%list)
  std::vector<int> ltp;             // ltp(1:max_vdw)
  std::vector<double> param;        // param(1:7, 1:max_vdw)
  std::vector<double> tab_potential, tab_force;   // (0:max_grid, 1:max_vdw) column-major
  std::vector<double> afs, bfs;     // (1:max_vdw)
  double cutoff = 0.0, dlrpot = 0.0, rdr = 0.0;
  bool l_force_shift = false, l_direct = false, no_vdw = true;
  double& tp(int i, int k) { return tab_potential[(size_t)(k - 1) * (max_grid + 1) + i]; }
  double& tf(int i, int k) { return tab_force[(size_t)(k - 1) * (max_grid + 1) + i]; }
  const double* par(int k) const { return &param[(size_t)(k - 1) * 7] - 1; }  // 1-based view
};
// coul_spole.F90: the direct-space Coulomb variants of two_body.F90:480-514 (everything except Ewald)
enum { COUL_NONE = 0, COUL_CP = 1, COUL_DDDP = 2, COUL_FSCP = 3, COUL_RFP = 4 };
struct Coulomb {
  int kind = COUL_NONE;
  bool damp = false;                       // electro%damp: Fennell-Gezelter damping through the erfc tables of `Ewald`
  double scaling = 0.0;                    // r4pie0/eps
  double force_shift = 0.0, energy_shift = 0.0;   // coul_spole.F90:186-202 / :413-414
  double rf[3] = {0, 0, 0};                // electro%reaction_field(0:2), :407-409
};
struct Ewald {
  bool active = false;
  double alpha = 0.0, scaling = 0.0;   // scaling = r4pie0/eps (two_body.F90:188)
  int nsamples = 0;
  double spacing = 0.0, recip_spacing = 0.0;
  std::vector<double> erfc, erfc_deriv;    // table(1:nsamples); slot 0 = out-of-bounds slot (numerics.F90:235)
};

// electrostatic.F90:88-127 erfcgen + numerics.F90:215-248 init_interp_table
void erfcgen(double rcut, double alpha, int nsamples, double* erfc_t /*1-based, [0] set to 0*/, double* deriv_t,
             double& spacing, double& recip) {
  spacing = rcut / (double)(nsamples - 4);
  recip = 1.0 / spacing;
  erfc_t[0] = 0.0; deriv_t[0] = 0.0;   // reference: out of bounds (quirk 4); never read when r >= spacing
  for (int i = 1; i <= nsamples; ++i) {
    double x = (double)i * spacing;
    erfc_t[i] = calc_erfc(alpha * x) / x;
  }
  for (int i = 1; i <= nsamples; ++i) {
    double rrr = (double)i * spacing;
    double rsq = rrr * rrr;
    double e = calc_erfc(alpha * rrr) / rrr;
    deriv_t[i] = (e + alpha * calc_erfc_deriv(alpha * rrr)) / rsq;
  }
}

// vdw.F90:1397-1576 vdw_generate (one potential k)
void vdw_generate_one(Vdw& v, int ivdw) {
  double dlrpot = v.cutoff / (double)(v.max_grid - 4);
  int keypot = v.ltp[ivdw];
  const double* prm = v.par(ivdw);
  if (keypot != 0) {
    for (int i = 1; i <= v.max_grid; ++i) {
      double r = (double)i * dlrpot;
      EG z = pot_energy(keypot, prm, r);
      v.tp(i, ivdw) = z.energy;
      v.tf(i, ivdw) = z.gamma;
    }
    v.tp(0, ivdw) = DBL_MAX;   // Huge()
    v.tf(0, ivdw) = DBL_MAX;
  }
  switch (keypot) {
    case 8: {  // VDW_MORSE :1463-1469
      double e0 = prm[1], r0 = prm[2], kk = prm[3];
      double t1 = std::exp(+kk * r0);
      v.tf(0, ivdw) = -2.0 * e0 * kk * (1.0 - t1) * t1;
      break;
    }
    case 10: v.tf(0, ivdw) = prm[1]; break;                 // VDW_DPD
    case 23: v.tf(0, ivdw) = prm[1] * prm[2]; break;        // VDW_NDPD
    case 14: v.tp(0, ivdw) = prm[1]; v.tf(0, ivdw) = 0.0; break;   // VDW_RYDBERG
    case 18: case 19: case 20: {                            // *_MDF: rc := vdws%cutoff  :1488-1550
      double q[8] = {0};
      int nn = (keypot == 19) ? 4 : 3;
      for (int i = 1; i <= nn; ++i) q[i] = prm[i];
      q[nn + 1] = v.cutoff;
      for (int i = 1; i <= v.max_grid; ++i) {
        double r = (double)i * dlrpot;
        EG z = pot_energy(keypot, q, r);
        v.tp(i, ivdw) = z.energy;
        v.tf(i, ivdw) = z.gamma;
      }
      v.tp(0, ivdw) = DBL_MAX; v.tf(0, ivdw) = DBL_MAX;
      break;
    }
    default: break;
  }
  // :1555-1566 force-shift loop computes and discards (quirk 2) -- nothing to do.
  if (std::fabs(v.tp(0, ivdw)) <= zero_plus)                 // :1570-1572
    v.tp(0, ivdw) = std::copysign(DBL_MIN, v.tp(0, ivdw));
}

// vdw.F90:969-1049 vdw_direct_fs_generate
void vdw_direct_fs_generate(Vdw& v) {
  v.afs.assign(v.max_vdw + 1, 0.0);
  v.bfs.assign(v.max_vdw + 1, 0.0);
  if (!v.l_force_shift) return;
  for (int ivdw = 1; ivdw <= v.n_vdw; ++ivdw) {
    int keypot = v.ltp[ivdw];
    EG z_dz = pot_energy(keypot, v.par(ivdw), v.cutoff);
    double z = z_dz.energy, dz = z_dz.gamma;
    if (keypot == 7 || keypot == 10 || keypot == 23 || keypot == 21 || keypot == 20) { z = 0.0; dz = 0.0; }
    v.afs[ivdw] = dz / v.cutoff;
    v.bfs[ivdw] = -z - dz;
  }
}

// vdw.F90:1196-1341: re-grid one TABLE array (buffer(1:ngrid) as read from file) onto tab(0:max_grid)
// is_force selects the index-0 extrapolation (:1208 vs :1269).  engunit multiply (:1332-1341) applied last.
void vdw_table_regrid(const double* buffer_in /*1..ngrid*/, int ngrid, double delpot_in, double rvdw, int max_grid,
                      bool is_force, double engunit, double* tab /*0..max_grid*/) {
  std::vector<double> buffer(ngrid + 1, 0.0);   // buffer(0) is never assigned in the reference (allocated 0:ngrid)
  for (int i = 1; i <= ngrid; ++i) buffer[i] = buffer_in[i];
  double delpot = delpot_in;
  double dlrpot = rvdw / (double)(max_grid - 4);
  if (std::fabs(delpot - dlrpot) <= 1.0e-8) delpot = dlrpot;      // :1116-1119
  bool remake = false;
  double rdr = 0.0;
  if (std::fabs(1.0 - (delpot / dlrpot)) > 1.0e-8) { remake = true; rdr = 1.0 / delpot; }
  for (int i = 0; i <= max_grid; ++i) tab[i] = 0.0;
  if (!is_force) tab[0] = 2.0 * buffer[1] - buffer[2];
  else tab[0] = (2.0 * buffer[1] - 0.5 * buffer[2]) / delpot;
  if (remake) {
    for (int i = 1; i <= max_grid - 4; ++i) {
      double rrr = (double)i * dlrpot;
      int l = f_int(rrr * rdr);
      double ppp = rrr * rdr - (double)l;
      double vk = buffer[l], vk1, vk2;
      if (l + 2 > ngrid) {
        if (l + 1 > ngrid) { vk1 = 2.0 * buffer[l] - buffer[l - 1]; vk2 = 2.0 * vk1 - buffer[l]; }
        else { vk1 = buffer[l + 1]; vk2 = 2.0 * buffer[l + 1] - buffer[l]; }
      } else { vk1 = buffer[l + 1]; vk2 = buffer[l + 2]; }
      double t1 = vk + (vk1 - vk) * ppp;
      double t2 = vk1 + (vk2 - vk1) * (ppp - 1.0);
      tab[i] = t1 + (t2 - t1) * ppp * 0.5;
    }
  } else {
    for (int i = 1; i <= max_grid - 4; ++i) tab[i] = buffer[i];
    tab[max_grid - 3] = 2.0 * tab[max_grid - 4] - tab[max_grid - 5];
  }
  tab[max_grid - 2] = 2.0 * tab[max_grid - 3] - tab[max_grid - 4];
  if (!is_force && std::fabs(tab[0]) <= zero_plus) tab[0] = std::copysign(DBL_MIN, tab[0]);   // :1319-1321
  for (int i = 0; i <= max_grid; ++i) tab[i] = tab[i] * engunit;
}

// ---------------------------------------------------------------- domain / world
struct Sites {   // site.F90 arrays indexed by lsite
  std::vector<int> type_site, freeze_site;   // 1-based
  std::vector<double> charge_site;
};

struct Dom {
  int idnode = 0;
  int nx = 1, ny = 1, nz = 1, idx = 0, idy = 0, idz = 0;
  double nx_real = 1, ny_real = 1, nz_real = 1, nx_recip = 1, ny_recip = 1, nz_recip = 1;
  int map[27] = {0};
  int natms = 0, nlast = 0;
  std::vector<CorePart> parts;           // 1-based
  std::vector<int> ltg, lsite, ltype, lfrzn, ixyz;
  std::vector<double> vxx, vyy, vzz;
  std::vector<double> xbg, ybg, zbg;
  // verlet list: list(-3:max_list, 1:natms)
  int max_list = 0;
  std::vector<int> list;
  int max_exclude = 0;
  std::vector<int> list_excl;            // (0:max_exclude, 1:natms)
  // link-cell diagnostics of the last build
  int nlx = 0, nly = 0, nlz = 0, nlp = 0, ncells = 0, nsbcll = 0;
  std::vector<int> which_cell, at_list, lct_start;
  int ibig = 0;
  bool list_safe = true;
  // results of the last two_body call
  double stress[10] = {0};
  double engvdw = 0, virvdw = 0, engcpe_rl = 0, vircpe_rl = 0, engcpe_ex = 0, vircpe_ex = 0;
  // stats%pp_energy(1:natms), stats%pp_stress(1:9, 1:natms) of the last two_body call (World::collect_pp), 1-based atoms
  std::vector<double> pp_energy, pp_stress;
  inline void pp_add(int atom, double e_half, const double* x, const double* f) {   // statistics.F90:2616-2625 calculate_stress, * 0.5
    pp_energy[atom] += e_half;
    double* t = &pp_stress[(size_t)atom * 9];
    for (int c = 0; c < 3; ++c) { t[c] += x[0] * f[c] * 0.5; t[3 + c] += x[1] * f[c] * 0.5; t[6 + c] += x[2] * f[c] * 0.5; }
  }
  // halo exchange scratch
  std::vector<double> sendbuf;
  inline int& L(int k, int i) { return list[(size_t)(i - 1) * (max_list + 4) + (k + 3)]; }
  inline int& LE(int k, int i) { return list_excl[(size_t)(i - 1) * (max_exclude + 1) + k]; }
  void ensure(int n) {
    if ((int)parts.size() < n + 1) {
      size_t m = (size_t)(n + 1) + (size_t)(n / 4) + 16;
      parts.resize(m); ltg.resize(m); lsite.resize(m); ltype.resize(m); lfrzn.resize(m); ixyz.resize(m);
      vxx.resize(m); vyy.resize(m); vzz.resize(m);
    }
  }
};

struct World {
  int P = 1, imcon = 1;
  int nthreads = 1;   // host threads the world-level loops over domains may use (one domain = one would-be MPI rank)
  double cell[10] = {0};
  int megatm = 0, megfrz = 0;
  bool lbook = false;
  double rcut = 0, padding = 0, rx = 0, pdplnc = 50.0;
  double ecw[4] = {0, 0, 0, 0};          // SPME negative-direction halo widths (reduced); 0 => link-cell width
  Vdw vdw;
  Ewald ew;
  Coulomb coul;
  Sites sites;
  int max_exclude = 0;
  std::vector<int> excl_global;          // (0:max_exclude, 1:megatm) by global id
  std::vector<Dom> d;
  bool collect_pp = false;               // stats%collect_pp (statistics.F90:227): per-particle energy / stress
  bool update = true;                    // neigh%update
  double neighskip[6] = {0, 0, 0, 0, 999999999.0, 0};   // statistics.F90:185-186 (index 0 unused)
  bool newstart = true;
  bool l_str = true;                     // strict_checks (control.F90:4421-4427: default on)
  int bspline = 0;                       // > 0 <=> SPME is on (vnl_check's test = 2 % instead of 4 %)
  std::string err;
};

// domains.F90:63-258 map_domains ; numerics.F90:3563-3645 factor/get_nth_prime ; domains.F90:260-335
const int primes_tab[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37, 41, 43, 47, 53, 59, 61, 67, 71, 73, 79, 83, 89, 97};
inline int get_nth_prime(int n) { return (n >= 1 && n <= (int)(sizeof(primes_tab) / sizeof(int))) ? primes_tab[n - 1] : -1; }
const int max_factor = 10;
void factor(int n, int* facs /*1..max_factor*/) {
  for (int i = 1; i <= max_factor; ++i) facs[i] = 0;
  int left = n;
  for (int i = 1; i <= max_factor - 1; ++i) {
    int p = get_nth_prime(i);
    if (p <= 0) break;
    while (p * (left / p) == left) { left = left / p; facs[i] = facs[i] + 1; }
  }
  facs[max_factor] = left;
}
int get_n_factors(const int* f) {
  int nf = 1;
  for (int i = 1; i <= max_factor - 1; ++i) nf *= (f[i] + 1);
  if (f[max_factor] != 1) nf *= 2;
  return nf;
}
int get_nth_factor(const int* f, int n) {
  int nfacs = get_n_factors(f);
  if (n > nfacs) return -1;
  int nt = (f[max_factor] != 1 && n > nfacs / 2) ? n - nfacs / 2 : n;
  nt = nt - 1;
  int fac_counts[max_factor + 1] = {0};
  int dim_prod = 1;
  for (int i = 1; i <= max_factor - 2; ++i) dim_prod *= (f[i] + 1);
  for (int i = max_factor - 1; i >= 2; --i) {
    fac_counts[i] = nt / dim_prod;
    nt = nt - fac_counts[i] * dim_prod;
    dim_prod = dim_prod / (f[i - 1] + 1);
  }
  fac_counts[1] = nt;
  int r = 1;
  for (int i = 1; i <= max_factor - 1; ++i)
    for (int k = 0; k < fac_counts[i]; ++k) r *= get_nth_prime(i);
  if (f[max_factor] != 1 && n > nfacs / 2) r *= f[max_factor];
  return r;
}
inline int idcube(int i, int j, int k, int nx, int ny) { return i + nx * (j + ny * k); }

void map_domains(int imcon, double wx, double wy, double wz, int mxnode, int idnode, Dom& dom) {
  const double tol = 1.0e-6;
  if (mxnode == 1) {
    dom.nx = dom.ny = dom.nz = 1;
  } else {
    int limx = (imcon != 0) ? INT32_MAX : 2, limy = limx;
    int limz = (imcon != 0 && imcon != 6) ? INT32_MAX : 2;
    double min_S = DBL_MAX;
    dom.nx = dom.ny = dom.nz = -1;
    int P = mxnode;
    int pfx[max_factor + 1], pfy[max_factor + 1];
    factor(P, pfx);
    int nfx = get_n_factors(pfx);
    for (int i = 1; i <= nfx; ++i) {
      int nx = get_nth_factor(pfx, i);
      if (nx > limx) continue;
      double dx = wx / (double)nx;
      int Pyz = P / nx;
      factor(Pyz, pfy);
      int nfy = get_n_factors(pfy);
      for (int j = 1; j <= nfy; ++j) {
        int ny = get_nth_factor(pfy, j);
        if (ny > limy) continue;
        double dy = wy / (double)ny;
        int nz = Pyz / ny;
        if (nz > limz) continue;
        double dz = wz / (double)nz;
        double S = 2.0 * (dx * dy + dy * dz + dz * dx);
        auto take = [&]() { min_S = S; dom.nx = nx; dom.ny = ny; dom.nz = nz; };
        if (min_S - S > tol) {
          take();
        } else if (std::fabs(min_S - S) < tol) {
          int mnew = std::max(nx, std::max(ny, nz)), mold = std::max(dom.nx, std::max(dom.ny, dom.nz));
          if (mnew < mold) take();
          else if (mnew == mold) {
            if (nx < dom.nx) take();
            else if (nx == dom.nx && ny < dom.ny) take();
          }
        }
      }
    }
  }
  dom.idnode = idnode;
  dom.nx_real = dom.nx; dom.nx_recip = 1.0 / dom.nx_real;
  dom.ny_real = dom.ny; dom.ny_recip = 1.0 / dom.ny_real;
  dom.nz_real = dom.nz; dom.nz_recip = 1.0 / dom.nz_real;
  dom.idz = idnode / (dom.nx * dom.ny);
  dom.idy = idnode / dom.nx - dom.idz * dom.ny;
  dom.idx = idnode % dom.nx;
  int nx = dom.nx, ny = dom.ny, nz = dom.nz, idx = dom.idx, idy = dom.idy, idz = dom.idz;
  int jdz = nz + idz, jdy = ny + idy, jdx = nx + idx;
  int xm = (jdx - 1) % nx, xp = (idx + 1) % nx, ym = (jdy - 1) % ny, yp = (idy + 1) % ny, zm = (jdz - 1) % nz,
      zp = (idz + 1) % nz;
  int* m = dom.map;
  m[1] = idcube(xm, idy, idz, nx, ny); m[2] = idcube(xp, idy, idz, nx, ny);
  m[3] = idcube(idx, ym, idz, nx, ny); m[4] = idcube(idx, yp, idz, nx, ny);
  m[5] = idcube(idx, idy, zm, nx, ny); m[6] = idcube(idx, idy, zp, nx, ny);
  m[7] = idcube(xm, yp, idz, nx, ny); m[8] = idcube(xp, ym, idz, nx, ny);
  m[9] = idcube(xm, ym, idz, nx, ny); m[10] = idcube(xp, yp, idz, nx, ny);
  m[11] = idcube(xm, idy, zp, nx, ny); m[12] = idcube(xp, idy, zm, nx, ny);
  m[13] = idcube(xm, idy, zm, nx, ny); m[14] = idcube(xp, idy, zp, nx, ny);
  m[15] = idcube(idx, ym, zp, nx, ny); m[16] = idcube(idx, yp, zm, nx, ny);
  m[17] = idcube(idx, ym, zm, nx, ny); m[18] = idcube(idx, yp, zp, nx, ny);
  m[19] = idcube(xm, ym, zm, nx, ny); m[20] = idcube(xp, yp, zp, nx, ny);
  m[21] = idcube(xm, ym, zp, nx, ny); m[22] = idcube(xp, yp, zm, nx, ny);
  m[23] = idcube(xm, yp, zm, nx, ny); m[24] = idcube(xp, ym, zp, nx, ny);
  m[25] = idcube(xm, yp, zp, nx, ny); m[26] = idcube(xp, ym, zm, nx, ny);
}

// halo.F90:153-302 set_halo_particles -- part 1: thresholds + ixyz tagging for one domain
void halo_tag(World& w, Dom& dom) {
  double cut = w.rx + 1.0e-6;
  double celprp[11];
  dcell(w.cell, celprp);
  int nlx = f_int(celprp[7] / (cut * dom.nx_real));
  int nly = f_int(celprp[8] / (cut * dom.ny_real));
  int nlz = f_int(celprp[9] / (cut * dom.nz_real));
  double xdc = (double)(nlx * dom.nx), ydc = (double)(nly * dom.ny), zdc = (double)(nlz * dom.nz);
  double cwx = 1.0 / xdc, cwy = 1.0 / ydc, cwz = 1.0 / zdc;
  // halo.F90:219-233: SPME may ask for a wider negative-direction halo (w.ecw = num_spline_pad/kmax); Max(cw, ecw)
  double ecwx = std::max(cwx, w.ecw[1]), ecwy = std::max(cwy, w.ecw[2]), ecwz = std::max(cwz, w.ecw[3]);
  ecwx = std::nextafter((-0.5 + ecwx) + (double)dom.idx * dom.nx_recip, DBL_MAX) + zero_plus;
  ecwy = std::nextafter((-0.5 + ecwy) + (double)dom.idy * dom.ny_recip, DBL_MAX) + zero_plus;
  ecwz = std::nextafter((-0.5 + ecwz) + (double)dom.idz * dom.nz_recip, DBL_MAX) + zero_plus;
  cwx = std::nextafter((-0.5 - cwx) + (double)(dom.idx + 1) * dom.nx_recip, -DBL_MAX) - zero_plus -
        (nlx == 1 ? cwx * 1.0e-10 : 0.0);
  cwy = std::nextafter((-0.5 - cwy) + (double)(dom.idy + 1) * dom.ny_recip, -DBL_MAX) - zero_plus -
        (nly == 1 ? cwy * 1.0e-10 : 0.0);
  cwz = std::nextafter((-0.5 - cwz) + (double)(dom.idz + 1) * dom.nz_recip, -DBL_MAX) - zero_plus -
        (nlz == 1 ? cwz * 1.0e-10 : 0.0);
  double rcell[10], det;
  invert(w.cell, rcell, det);
  dom.nlast = dom.natms;
  for (int i = 1; i <= dom.nlast; ++i) {
    dom.ixyz[i] = 0;
    const CorePart& p = dom.parts[i];
    double x = rcell[1] * p.xxx + rcell[4] * p.yyy + rcell[7] * p.zzz;
    double y = rcell[2] * p.xxx + rcell[5] * p.yyy + rcell[8] * p.zzz;
    double z = rcell[3] * p.xxx + rcell[6] * p.yyy + rcell[9] * p.zzz;
    if (x <= ecwx) dom.ixyz[i] += 1;
    if (x >= cwx) dom.ixyz[i] += 2;
    if (y <= ecwy) dom.ixyz[i] += 10;
    if (y >= cwy) dom.ixyz[i] += 20;
    if (z <= ecwz) dom.ixyz[i] += 100;
    if (z >= cwz) dom.ixyz[i] += 200;
  }
}

struct DirSet { int kx, ky, kz, jxyz, kxyz, jd, kd; double xadd, yadd, zadd; bool lwrap; };
// deport_data.F90:1728-1796 direction settings (shared by export_atomic_data/positions and deport)
DirSet dir_settings(const World& w, const Dom& dom, int mdir) {
  DirSet s{0, 0, 0, 0, 0, 0, 0, 0, 0, 0, false};
  bool lsx = false, lex = false, lsy = false, ley = false, lsz = false, lez = false;
  switch (mdir) {
    case -1: s.kx = 1; s.jxyz = 1; s.kxyz = 3; lsx = (dom.idx == 0); s.jd = dom.map[1]; s.kd = dom.map[2]; break;
    case 1: s.kx = 1; s.jxyz = 2; s.kxyz = 3; lex = (dom.idx == dom.nx - 1); s.jd = dom.map[2]; s.kd = dom.map[1]; break;
    case -2: s.ky = 1; s.jxyz = 10; s.kxyz = 30; lsy = (dom.idy == 0); s.jd = dom.map[3]; s.kd = dom.map[4]; break;
    case 2: s.ky = 1; s.jxyz = 20; s.kxyz = 30; ley = (dom.idy == dom.ny - 1); s.jd = dom.map[4]; s.kd = dom.map[3]; break;
    case -3: s.kz = 1; s.jxyz = 100; s.kxyz = 300; lsz = (dom.idz == 0); s.jd = dom.map[5]; s.kd = dom.map[6]; break;
    case 3: s.kz = 1; s.jxyz = 200; s.kxyz = 300; lez = (dom.idz == dom.nz - 1); s.jd = dom.map[6]; s.kd = dom.map[5]; break;
  }
  double uuu = 0.0; if (lsx) uuu = +1.0; if (lex) uuu = -1.0;
  double vvv = 0.0; if (lsy) vvv = +1.0; if (ley) vvv = -1.0;
  double www = 0.0; if (lsz) www = +1.0; if (lez) www = -1.0;
  s.lwrap = (std::fabs(uuu) + std::fabs(vvv) + std::fabs(www) > 0.5);
  if (s.lwrap) {
    s.xadd = w.cell[1] * uuu + w.cell[4] * vvv + w.cell[7] * www;
    s.yadd = w.cell[2] * uuu + w.cell[5] * vvv + w.cell[8] * www;
    s.zadd = w.cell[3] * uuu + w.cell[6] * vvv + w.cell[9] * www;
  }
  return s;
}

// deport_data.F90:1673-1951 export_atomic_data, all ranks of the world "simultaneously"
void export_atomic_data(World& w, int mdir) {
  const int iadd = 6;
#pragma omp parallel for schedule(static, 1) num_threads(w.nthreads) if (w.nthreads > 1)
  for (int rr = 0; rr < (int)w.d.size(); ++rr) {   // pack phase (every rank packs before anybody unpacks = MPI semantics)
    Dom& dom = w.d[rr];
    DirSet s = dir_settings(w, dom, mdir);
    dom.sendbuf.clear();
    for (int i = 1; i <= dom.nlast; ++i) {
      if (dom.ixyz[i] > 0) {
        int ix = dom.ixyz[i] % 10;
        int iy = (dom.ixyz[i] - ix) % 100;
        int iz = (dom.ixyz[i] - (ix + iy)) % 1000;
        int j = ix * s.kx + iy * s.ky + iz * s.kz;
        if (j == s.jxyz || (j > s.jxyz && j % 3 == 0)) {
          const CorePart& p = dom.parts[i];
          if (!s.lwrap) {
            dom.sendbuf.push_back(p.xxx); dom.sendbuf.push_back(p.yyy); dom.sendbuf.push_back(p.zzz);
          } else {
            dom.sendbuf.push_back(p.xxx + s.xadd); dom.sendbuf.push_back(p.yyy + s.yadd); dom.sendbuf.push_back(p.zzz + s.zadd);
          }
          dom.sendbuf.push_back((double)dom.ltg[i]);
          dom.sendbuf.push_back((double)dom.lsite[i]);
          dom.sendbuf.push_back((double)(dom.ixyz[i] - (j == s.jxyz ? s.jxyz : s.kxyz)));
        }
      }
    }
  }
#pragma omp parallel for schedule(static, 1) num_threads(w.nthreads) if (w.nthreads > 1)
  for (int rr = 0; rr < (int)w.d.size(); ++rr) {   // unpack phase: receive from kdnode
    Dom& dom = w.d[rr];
    DirSet s = dir_settings(w, dom, mdir);
    const std::vector<double>& buf = w.d[s.kd].sendbuf;
    int jmove = (int)buf.size();
    dom.ensure(dom.nlast + jmove / iadd);
    int j = 0;
    for (int i = 1; i <= jmove / iadd; ++i) {
      dom.nlast = dom.nlast + 1;
      CorePart& p = dom.parts[dom.nlast];
      p.xxx = buf[j + 0]; p.yyy = buf[j + 1]; p.zzz = buf[j + 2];
      p.fxx = p.fyy = p.fzz = 0.0;
      dom.ltg[dom.nlast] = f_nint(buf[j + 3]);
      dom.lsite[dom.nlast] = f_nint(buf[j + 4]);
      dom.ixyz[dom.nlast] = f_nint(buf[j + 5]);
      j = j + iadd;
    }
  }
}

// neighbours.F90:305-343 vnl_set_check
void vnl_set_check(Dom& dom) {
  dom.xbg.resize(dom.nlast + 1); dom.ybg.resize(dom.nlast + 1); dom.zbg.resize(dom.nlast + 1);
  for (int i = 1; i <= dom.nlast; ++i) {
    dom.xbg[i] = dom.parts[i].xxx; dom.ybg[i] = dom.parts[i].yyy; dom.zbg[i] = dom.parts[i].zzz;
  }
}

// halo.F90:153-355 set_halo_particles (world-wide)
void set_halo_particles(World& w) {
#pragma omp parallel for schedule(static, 1) num_threads(w.nthreads) if (w.nthreads > 1)
  for (int rr = 0; rr < (int)w.d.size(); ++rr) halo_tag(w, w.d[rr]);
  export_atomic_data(w, -1); export_atomic_data(w, 1);
  export_atomic_data(w, -2); export_atomic_data(w, 2);
  export_atomic_data(w, -3); export_atomic_data(w, 3);
#pragma omp parallel for schedule(static, 1) num_threads(w.nthreads) if (w.nthreads > 1)
  for (int rr = 0; rr < (int)w.d.size(); ++rr) {
    Dom& dom = w.d[rr];
    for (int i = dom.natms + 1; i <= dom.nlast; ++i) {          // halo.F90:296-302
      dom.ltype[i] = w.sites.type_site[dom.lsite[i]];
      dom.parts[i].chge = w.sites.charge_site[dom.lsite[i]];
      dom.lfrzn[i] = w.sites.freeze_site[dom.lsite[i]];
    }
    vnl_set_check(dom);                                           // halo.F90:315
  }
}

// deport_data.F90:2301-2553 export_atomic_positions (world-wide); mlast per domain
void export_atomic_positions(World& w, int mdir, std::vector<int>& mlast) {
  const int iadd = 3;
#pragma omp parallel for schedule(static, 1) num_threads(w.nthreads) if (w.nthreads > 1)
  for (int r = 0; r < (int)w.d.size(); ++r) {
    Dom& dom = w.d[r];
    DirSet s = dir_settings(w, dom, mdir);
    dom.sendbuf.clear();
    for (int i = 1; i <= mlast[r]; ++i) {
      if (dom.ixyz[i] > 0) {
        int ix = dom.ixyz[i] % 10;
        int iy = (dom.ixyz[i] - ix) % 100;
        int iz = (dom.ixyz[i] - (ix + iy)) % 1000;
        int j = ix * s.kx + iy * s.ky + iz * s.kz;
        if (j == s.jxyz || (j > s.jxyz && j % 3 == 0)) {
          const CorePart& p = dom.parts[i];
          if (!s.lwrap) {
            dom.sendbuf.push_back(p.xxx); dom.sendbuf.push_back(p.yyy); dom.sendbuf.push_back(p.zzz);
          } else {
            dom.sendbuf.push_back(p.xxx + s.xadd); dom.sendbuf.push_back(p.yyy + s.yadd); dom.sendbuf.push_back(p.zzz + s.zadd);
          }
        }
      }
    }
  }
#pragma omp parallel for schedule(static, 1) num_threads(w.nthreads) if (w.nthreads > 1)
  for (int r = 0; r < (int)w.d.size(); ++r) {
    Dom& dom = w.d[r];
    DirSet s = dir_settings(w, dom, mdir);
    const std::vector<double>& buf = w.d[s.kd].sendbuf;
    int jmove = (int)buf.size(), j = 0;
    for (int i = 1; i <= jmove / iadd; ++i) {
      mlast[r] = mlast[r] + 1;
      CorePart& p = dom.parts[mlast[r]];
      p.xxx = buf[j + 0]; p.yyy = buf[j + 1]; p.zzz = buf[j + 2];
      j = j + iadd;
    }
  }
}

// halo.F90:47-113 refresh_halo_positions; returns 0 ok, 138 on count mismatch
int refresh_halo_positions(World& w) {
  std::vector<int> mlast(w.d.size());
  for (size_t r = 0; r < w.d.size(); ++r) mlast[r] = w.d[r].natms;
  export_atomic_positions(w, -1, mlast); export_atomic_positions(w, 1, mlast);
  export_atomic_positions(w, -2, mlast); export_atomic_positions(w, 2, mlast);
  export_atomic_positions(w, -3, mlast); export_atomic_positions(w, 3, mlast);
  for (size_t r = 0; r < w.d.size(); ++r)
    if (mlast[r] != w.d[r].nlast) return 138;
  return 0;
}

// neighbours.F90:182-284: the update decision, the padding re-tune of the 'no strict' regime and the skip statistics, for a
// displacement maximum that is already global.  celprp from dcell(cell); dims = (nx, ny, nz); mxnode = nx ny nz.
// Returns 0, or 307 (the reference's error number).  The KIM clause (:261-266) is outside this path.
int vnl_decide(bool l_str, double tolg, int bspline, double cutoff, double& padding, double& cutoff_extended, const double* cell,
               int nx, int ny, int nz, bool& update, bool& newstart, double* ns /*1..5*/, double& width) {
  update = (tolg >= half_minus * padding);                                   // :182
  double celprp[11];
  dcell(cell, celprp);                                                       // :186
  width = std::min(celprp[7], std::min(celprp[8], celprp[9]));               // :187
  double cut = cutoff_extended + smalldr;                                    // :191
  double nx_recip = 1.0 / (double)nx, ny_recip = 1.0 / (double)ny, nz_recip = 1.0 / (double)nz;
  int mxnode = nx * ny * nz;
  int ilx = f_int(nx_recip * celprp[7] / cut);                               // :195-197
  int ily = f_int(ny_recip * celprp[8] / cut);
  int ilz = f_int(nz_recip * celprp[9] / cut);
  double m6 = 0.05, m7 = 0.005, m8 = 0.02, m9 = 0.95;                         // :199-202
  double tol = std::min(m6, m7 * cutoff);                                    // :204
  double test;
  if (bspline > 0) test = m8; else test = m8 * 2.0;                          // :206-210
  cut = std::min(nx_recip * celprp[7], std::min(ny_recip * celprp[8], nz_recip * celprp[9])) - smalldr;   // :212-214
  if (ilx * ily * ilz == 0) {                                                // :216
    if (cut < cutoff) return 307;                                            // :217-219
    else {
      if (cut < cutoff_extended) {                                           // :221
        if (l_str) return 307;                                               // :222-224
        else {
          if (cut >= cutoff) {                                               // :226-232
            padding = std::min(m9 * (cut - cutoff), test * cutoff);
            padding = (double)f_int(100.0 * padding) / 100.0;
            if (padding < tol) padding = 0.0;
            cutoff_extended = cutoff + padding;
            update = true;
          }
        }
      }
    }
  } else {                                                                   // :236
    if (update && (!l_str)) {                                                // :237
      if (f_int((double)std::min(ilx, std::min(ily, ilz)) / (1.0 + test)) >= 2) {   // :238
        cut = test * cutoff;
      } else {
        if (mxnode > 1) {                                                    // :241-245
          cut = std::min(m9 * (std::min(nx_recip * celprp[7] / (double)ilx,
                                        std::min(ny_recip * celprp[8] / (double)ily, nz_recip * celprp[9] / (double)ilz)) -
                               cutoff - smalldr),
                         test * cutoff);
        } else {
          cut = m9 * (0.5 * width - cutoff - smalldr);                       // :247
        }
      }
      cut = (double)f_int(100.0 * cut) / 100.0;                              // :250
      if ((!(cut < tol)) && cut - padding > 0.005) {                         // :251-257
        padding = cut;
        cutoff_extended = cutoff + padding;
      }
    }
  }
  if (update) {                                                              // :270-284
    ns[3] = ns[2] * ns[3];
    ns[2] = ns[2] + 1.0;
    ns[3] = ns[3] / ns[2] + ns[1] / ns[2];
    if (!newstart) ns[4] = std::min(ns[1], ns[4]); else newstart = false;
    ns[5] = std::max(ns[1], ns[5]);
    ns[1] = 0.0;
  } else {
    ns[1] = ns[1] + 1.0;
  }
  return 0;
}

// neighbours.F90:123-296 vnl_check
bool vnl_check(World& w, double* tol_out) {
  if (!(w.padding > 0.0)) {   // unconditional_update false => returns leaving update=.true. (neighbours.F90:141)
    w.update = true;
    if (tol_out) *tol_out = 0.0;
    return true;
  }
  double tol = 0.0;
#pragma omp parallel for schedule(static, 1) num_threads(w.nthreads) if (w.nthreads > 1) reduction(max : tol)
  for (int rr = 0; rr < (int)w.d.size(); ++rr) {
    Dom& dom = w.d[rr];
    int n = dom.natms;
    std::vector<double> x(n + 1), y(n + 1), z(n + 1);
    for (int i = 1; i <= n; ++i) {
      x[i] = dom.parts[i].xxx - dom.xbg[i];
      y[i] = dom.parts[i].yyy - dom.ybg[i];
      z[i] = dom.parts[i].zzz - dom.zbg[i];
    }
    images(w.imcon, w.cell, n, x.data(), y.data(), z.data());
    for (int i = 1; i <= n; ++i) {
      double r = std::sqrt(x[i] * x[i] + y[i] * y[i] + z[i] * z[i]);
      tol = std::max(tol, r);   // Maxval + gmax
    }
  }
  if (tol_out) *tol_out = tol;
  double width;
  int rc = vnl_decide(w.l_str, tol, w.bspline, w.rcut, w.padding, w.rx, w.cell, w.d[0].nx, w.d[0].ny, w.d[0].nz, w.update, w.newstart,
                      w.neighskip, width);
  if (rc != 0) { w.err = "vnl_check: error 307"; w.update = true; }
  return w.update;
}

// ---------------------------------------------------------------- neighbours.F90:356-1306 link_cell_pairs
int link_cell_pairs(World& w, Dom& dom) {
  double celprp[11];
  dcell(w.cell, celprp);
  double det = std::min(celprp[7], std::min(celprp[8], celprp[9]));
  if (w.rx >= det / 2.0) return 95;
  double cut = w.rx + smalldr;
  double rcsq = w.rx * w.rx;
  double dispx = dom.nx_recip * celprp[7] / cut;
  double dispy = dom.ny_recip * celprp[8] / cut;
  double dispz = dom.nz_recip * celprp[9] / cut;
  int nlx = f_int(dispx), nly = f_int(dispy), nlz = f_int(dispz);
  if (nlx * nly * nlz == 0) return 307;
  int nlp = 1;
  double nlr2 = (double)dom.natms;
  det = nlr2 / (double)(nlx * nly * nlz);
  while (det > w.pdplnc) {
    nlp = nlp + 1;
    double rsq = (double)nlp;
    nlx = f_int(dispx * rsq); nly = f_int(dispy * rsq); nlz = f_int(dispz * rsq);
    det = nlr2 / (double)(nlx * nly * nlz);
  }
  int ncells = (nlx + 2 * nlp) * (nly + 2 * nlp) * (nlz + 2 * nlp);
  int nlp2 = (1 + (1 + 2 * nlp) * (1 + 2 * nlp) * (1 + 2 * nlp)) / 2;
  int nlp3 = nlx * nly * nlz;
  int nlp4 = nlp3 - std::max(nlx - 2 * nlp, 0) * std::max(nly - 2 * nlp, 0) * std::max(nlz - 2 * nlp, 0);
  std::vector<int> nix(nlp2 + 1, 0), niy(nlp2 + 1, 0), niz(nlp2 + 1, 0);
  std::vector<char> nir(nlp2 + 1, 0);
  std::vector<int> cell_dom(nlp3 + 1), cell_bor(nlp4 + 1);
  cell_dom[0] = nlp3; cell_bor[0] = nlp4;
  // semi-ball stencil :490-540
  {
    int nlp2s = nlp * nlp, nlp3s = (nlp - 1) * (nlp - 1);
    int nsbcll = 0;
    for (int iz = 0; iz <= nlp; ++iz) {
      int iz1 = (iz > 0) ? (iz - 1) * (iz - 1) : 0;
      int jz = iz * iz;
      for (int iy = -nlp; iy <= nlp; ++iy) {
        if (iz == 0 && iy < 0) continue;
        int a = std::abs(iy);
        int iy1 = (a > 0) ? (a - 1) * (a - 1) : 0;
        int ll = iz1 + iy1;
        if (ll > nlp2s) continue;
        int jy = jz + iy * iy;
        for (int ix = -nlp; ix <= nlp; ++ix) {
          if (iz == 0 && iy == 0 && ix < 0) continue;
          int b = std::abs(ix);
          int ix1 = (b > 0) ? (b - 1) * (b - 1) : 0;
          if (ll + ix1 > nlp2s) continue;
          int jx = jy + ix * ix;
          nsbcll = nsbcll + 1;
          nix[nsbcll] = ix; niy[nsbcll] = iy; niz[nsbcll] = iz;
          nir[nsbcll] = (jx < nlp3s);
        }
      }
    }
    dom.nsbcll = nsbcll;
  }
  const int nsbcll = dom.nsbcll;
  double xdc = (double)(nlx * dom.nx), ydc = (double)(nly * dom.ny), zdc = (double)(nlz * dom.nz);
  int jx = nlp - nlx * dom.idx, jy = nlp - nly * dom.idy, jz = nlp - nlz * dom.idz;
  int nlx0s = 0, nly0s = 0, nlz0s = 0;
  int nlx0e = nlp - 1, nly0e = nlp - 1, nlz0e = nlp - 1;
  int nlx1s = nlx + nlp, nly1s = nly + nlp, nlz1s = nlz + nlp;
  int nlx1e = nlx + 2 * nlp - 1, nly1e = nly + 2 * nlp - 1, nlz1e = nlz + 2 * nlp - 1;
  const int sx = nlx + 2 * nlp, sy = nly + 2 * nlp;
  std::vector<int> lct_count(ncells + 1, 0), lct_start(ncells + 2, 0), lct_where(ncells + 2, 0);
  std::vector<int>& which_cell = dom.which_cell;
  std::vector<int>& at_list = dom.at_list;
  which_cell.assign(dom.nlast + 1, 0);
  at_list.assign(dom.nlast + 1, 0);
  std::vector<double> xxt(dom.nlast + 1), yyt(dom.nlast + 1), zzt(dom.nlast + 1);
  double rcell[10], dt;
  invert(w.cell, rcell, dt);
  const CorePart* parts = dom.parts.data();
  for (int i = 1; i <= dom.natms; ++i) {   // :612-650
    double x = rcell[1] * parts[i].xxx + rcell[4] * parts[i].yyy + rcell[7] * parts[i].zzz;
    double y = rcell[2] * parts[i].xxx + rcell[5] * parts[i].yyy + rcell[8] * parts[i].zzz;
    double z = rcell[3] * parts[i].xxx + rcell[6] * parts[i].yyy + rcell[9] * parts[i].zzz;
    int ix = f_int(xdc * (x + 0.5)) + jx;
    int iy = f_int(ydc * (y + 0.5)) + jy;
    int iz = f_int(zdc * (z + 0.5)) + jz;
    ix = std::max(std::min(ix, nlx1s - 1), nlx0e + 1);
    iy = std::max(std::min(iy, nly1s - 1), nly0e + 1);
    iz = std::max(std::min(iz, nlz1s - 1), nlz0e + 1);
    int icell = 1 + ix + sx * (iy + sy * iz);
    lct_count[icell] += 1;
    which_cell[i] = icell;
  }
  for (int i = dom.natms + 1; i <= dom.nlast; ++i) {   // :652-799
    double x = rcell[1] * parts[i].xxx + rcell[4] * parts[i].yyy + rcell[7] * parts[i].zzz;
    double y = rcell[2] * parts[i].xxx + rcell[5] * parts[i].yyy + rcell[8] * parts[i].zzz;
    double z = rcell[3] * parts[i].xxx + rcell[6] * parts[i].yyy + rcell[9] * parts[i].zzz;
    int ix, iy, iz;
    double dpx, dpy, dpz;
    if (x > -half_plus) { dpx = xdc * (x + 0.5); ix = f_int(dpx) + jx; }
    else { dpx = xdc * std::fabs(x + 0.5); ix = -f_int(dpx) + jx - 1; }
    if (y > -half_plus) { dpy = ydc * (y + 0.5); iy = f_int(dpy) + jy; }
    else { dpy = ydc * std::fabs(y + 0.5); iy = -f_int(dpy) + jy - 1; }
    if (z > -half_plus) { dpz = zdc * (z + 0.5); iz = f_int(dpz) + jz; }
    else { dpz = zdc * std::fabs(z + 0.5); iz = -f_int(dpz) + jz - 1; }
    int icell;
    if (ix >= nlx0s && iy >= nly0s && iz >= nlz0s) {
      bool lx0 = (ix > nlx0e), lx1 = (ix < nlx1s), ly0 = (iy > nly0e), ly1 = (iy < nly1s), lz0 = (iz > nlz0e),
           lz1 = (iz < nlz1s);
      if ((lx0 && lx1) && (ly0 && ly1) && (lz0 && lz1)) {   // :699-756 put on border in the halo
        double xa = std::fabs(dpx - (double)(nlx * dom.idx));
        double x1 = std::fabs(dpx - (double)(nlx * (dom.idx + 1)));
        dpx = std::min(xa, x1);
        double ya = std::fabs(dpy - (double)(nly * dom.idy));
        double y1 = std::fabs(dpy - (double)(nly * (dom.idy + 1)));
        dpy = std::min(ya, y1);
        double za = std::fabs(dpz - (double)(nlz * dom.idz));
        double z1 = std::fabs(dpz - (double)(nlz * (dom.idz + 1)));
        dpz = std::min(za, z1);
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
      lx0 = (ix < nlx0s); lx1 = (ix > nlx1e); ly0 = (iy < nly0s); ly1 = (iy > nly1e); lz0 = (iz < nlz0s); lz1 = (iz > nlz1e);
      if (!(lx0 || lx1 || ly0 || ly1 || lz0 || lz1)) icell = 1 + ix + sx * (iy + sy * iz);
      else icell = 0;
    } else {
      icell = 0;
    }
    lct_count[icell] += 1;
    which_cell[i] = icell;
  }
  lct_start[0] = 1;   // :803-806
  for (int icell = 1; icell <= ncells + 1; ++icell) lct_start[icell] = lct_start[icell - 1] + lct_count[icell - 1];
  lct_where = lct_start;
  for (int i = 1; i <= dom.nlast; ++i) {   // :811-823
    int j = lct_where[which_cell[i]];
    at_list[j] = i;
    xxt[j] = parts[i].xxx; yyt[j] = parts[i].yyy; zzt[j] = parts[i].zzz;
    lct_where[which_cell[i]] += 1;
  }
  {  // :825-857 cell_dom / cell_bor
    int n3 = 0, n4 = 0;
    for (int iz = nlz0e + 1; iz <= nlz1s - 1; ++iz) {
      int iz1 = iz - nlz0e, iz2 = iz - nlz1s;
      for (int iy = nly0e + 1; iy <= nly1s - 1; ++iy) {
        int iy1 = iy - nly0e, iy2 = iy - nly1s;
        for (int ix = nlx0e + 1; ix <= nlx1s - 1; ++ix) {
          int ix1 = ix - nlx0e, ix2 = ix - nlx1s;
          int ic = 1 + ix + sx * (iy + sy * iz);
          n3 += 1;
          cell_dom[n3] = ic;
          if ((ix1 >= 1 && ix1 <= nlp) || (ix2 <= -1 && ix2 >= -nlp) || (iy1 >= 1 && iy1 <= nlp) ||
              (iy2 <= -1 && iy2 >= -nlp) || (iz1 >= 1 && iz1 <= nlp) || (iz2 <= -1 && iz2 >= -nlp)) {
            n4 += 1;
            cell_bor[n4] = ic;
          }
        }
      }
    }
  }
  const int ml = dom.max_list;
  dom.list.assign((size_t)std::max(dom.natms, 1) * (ml + 4), 0);
  int ibig = 0;
  bool safe = true;
  for (int ipass = 1; ipass <= 2; ++ipass) {   // :876-1027 (ipass=1), :1032-1184 (ipass=2)
    const std::vector<int>& cells = (ipass == 1) ? cell_dom : cell_bor;
    for (int icell = 1; icell <= cells[0]; ++icell) {
      int ic = cells[icell];
      int ix = (ic - 1) % sx;
      int iz = (ic - 1) / (sx * sy);
      int iy = (ic - 1) / sx - sy * iz;
      for (int ii = lct_start[ic]; ii <= lct_start[ic + 1] - 1; ++ii) {
        int i = at_list[ii];
        int j_start = 0;
        for (int kk = ipass; kk <= nsbcll; ++kk) {
          int jxx, jyy, jzz;
          if (ipass == 1) { jxx = ix + nix[kk]; jyy = iy + niy[kk]; jzz = iz + niz[kk]; }
          else {
            jxx = ix - nix[kk]; jyy = iy - niy[kk]; jzz = iz - niz[kk];
            if (!((jxx <= nlx0e) || (jxx >= nlx1s) || (jyy <= nly0e) || (jyy >= nly1s) || (jzz <= nlz0e) ||
                  (jzz >= nlz1s)))
              continue;
          }
          int jc = 1 + jxx + sx * (jyy + sy * jzz);
          if (jc != ic) j_start = lct_start[jc];
          else if (ipass == 1) j_start = ii + 1;
          int& cnt = dom.L(0, i);
          if (nir[kk]) {
            for (int jj = j_start; jj <= lct_start[jc + 1] - 1; ++jj) {
              int j = at_list[jj];
              int ll = cnt + 1;
              if (ll <= ml) dom.L(ll, i) = j; else { ibig = std::max(ibig, ll); safe = false; }
              cnt = ll;
            }
          } else {
            const double xi = parts[i].xxx, yi = parts[i].yyy, zi = parts[i].zzz;
            for (int jj = j_start; jj <= lct_start[jc + 1] - 1; ++jj) {
              int j = at_list[jj];
              double dx = xxt[jj] - xi, dy = yyt[jj] - yi, dz = zzt[jj] - zi;
              double rsq = dx * dx + dy * dy + dz * dz;   // :991-992 (** 2 then left-to-right +)
              if (rsq <= rcsq) {
                int ll = cnt + 1;
                if (ll <= ml) dom.L(ll, i) = j; else { safe = false; ibig = std::max(ibig, ll); }
                cnt = ll;
              }
            }
          }
        }
      }
    }
  }
  dom.ibig = ibig;
  dom.list_safe = safe;
  dom.nlx = nlx; dom.nly = nly; dom.nlz = nlz; dom.nlp = nlp; dom.ncells = ncells;
  dom.lct_start = lct_start;
  if (!safe) return 106;
  // :1198-1225 rear down frozen pairs
  if (w.megfrz > 1) {
    for (int i = 1; i <= dom.natms; ++i) {
      int l_end = dom.L(0, i), m_end = l_end;
      if (dom.lfrzn[i] > 0) {
        for (int kk = l_end; kk >= 1; --kk) {
          int j = dom.L(kk, i);
          if (dom.lfrzn[j] > 0) {
            if (kk < m_end) { dom.L(kk, i) = dom.L(m_end, i); dom.L(m_end, i) = j; }
            m_end = m_end - 1;
          }
        }
      }
      dom.L(-2, i) = dom.L(0, i);
      dom.L(0, i) = m_end;
    }
  } else {
    for (int i = 1; i <= dom.natms; ++i) dom.L(-2, i) = dom.L(0, i);
  }
  // :1229-1306 rear down excluded pairs
  if (w.lbook) {
    for (int i = 1; i <= dom.natms; ++i) {
      int l_end = dom.L(0, i), m_end = l_end;
      int ii = dom.LE(0, i);
      if (ii > 0) {
        const int* ex = &dom.LE(0, i);   // ex[1..ii]
        for (int kk = l_end; kk >= 1; --kk) {
          int j = dom.L(kk, i);
          int jj = dom.ltg[j];
          if (match(jj, ii, ex)) {
            if (kk < m_end) { dom.L(kk, i) = dom.L(m_end, i); dom.L(m_end, i) = j; }
            m_end = m_end - 1;
          }
        }
      }
      dom.L(-1, i) = dom.L(0, i);
      dom.L(0, i) = m_end;
    }
    for (int i = 1; i <= dom.natms; ++i) dom.L(-3, i) = dom.L(0, i);   // no CHARMM (:1297-1299)
  } else {
    for (int i = 1; i <= dom.natms; ++i) { dom.L(-1, i) = dom.L(0, i); dom.L(-3, i) = dom.L(0, i); }
  }
  return 0;
}

// ---------------------------------------------------------------- per-atom kernels
struct Acc { double eng, vir; };

// vdw.F90:1790-2024 vdw_forces_tab
Acc vdw_forces_tab(World& w, Dom& dom, int iatm, const double* xxt, const double* yyt, const double* zzt, const double* rrt) {
  Vdw& v = w.vdw;
  double engvdw = 0.0, virvdw = 0.0;
  double strs1 = 0, strs2 = 0, strs3 = 0, strs5 = 0, strs6 = 0, strs9 = 0;
  int idi = dom.ltg[iatm], ai = dom.ltype[iatm];
  CorePart* parts = dom.parts.data();
  double fix = parts[iatm].fxx, fiy = parts[iatm].fyy, fiz = parts[iatm].fzz;
  for (int mm = 1; mm <= dom.L(0, iatm); ++mm) {
    int jatm = dom.L(mm, iatm);
    int aj = dom.ltype[jatm];
    int key = (ai > aj) ? ai * (ai - 1) / 2 + aj : aj * (aj - 1) / 2 + ai;
    int k = v.list[key];
    if (std::fabs(v.tp(0, k)) < zero_plus) continue;
    double rrr = rrt[mm];
    int ityp = v.ltp[k];
    if (ityp != -1 && rrr < v.cutoff) {
      double r_rrr = 1.0 / rrr;
      double r_rvdw = 1.0 / v.cutoff;
      double rsq = rrr * rrr;
      double r_rsq = r_rrr * r_rrr;
      double r_rrv = r_rrr * r_rvdw;
      double rscl = rrr * r_rvdw;
      int l = f_int(rrr * v.rdr);
      double ppp = rrr * v.rdr - (double)l;
      double gk = v.tf(l, k); if (l == 0) gk = gk * rrr;
      double gk1 = v.tf(l + 1, k), gk2 = v.tf(l + 2, k);
      double t1 = gk + (gk1 - gk) * ppp;
      double t2 = gk1 + (gk2 - gk1) * (ppp - 1.0);
      double gamma = (t1 + (t2 - t1) * ppp * 0.5) * r_rsq;
      if (v.l_force_shift) gamma = gamma - v.tf(v.max_grid - 4, k) * r_rrv;
      double eng_pp = 0.0;   // vdw.F90:1905: eng = 0 at the top of every pair; it stays 0 for a halo partner this rank does not own
      double fx = gamma * xxt[mm], fy = gamma * yyt[mm], fz = gamma * zzt[mm];
      fix = fix + fx; fiy = fiy + fy; fiz = fiz + fz;
      if (jatm <= dom.natms) {
        parts[jatm].fxx = parts[jatm].fxx - fx;
        parts[jatm].fyy = parts[jatm].fyy - fy;
        parts[jatm].fzz = parts[jatm].fzz - fz;
      }
      if (jatm <= dom.natms || idi < dom.ltg[jatm]) {
        double vk = v.tp(l, k), vk1 = v.tp(l + 1, k), vk2 = v.tp(l + 2, k);
        t1 = vk + (vk1 - vk) * ppp;
        t2 = vk1 + (vk2 - vk1) * (ppp - 1.0);
        double eng = t1 + (t2 - t1) * ppp * 0.5;
        if (v.l_force_shift) eng = eng + v.tf(v.max_grid - 4, k) * (rscl - 1.0) - v.tp(v.max_grid - 4, k);
        engvdw = engvdw + eng;
        eng_pp = eng;
        virvdw = virvdw - gamma * rsq;
        strs1 = strs1 + xxt[mm] * fx; strs2 = strs2 + xxt[mm] * fy; strs3 = strs3 + xxt[mm] * fz;
        strs5 = strs5 + yyt[mm] * fy; strs6 = strs6 + yyt[mm] * fz; strs9 = strs9 + zzt[mm] * fz;
      }
      if (w.collect_pp) {   // vdw.F90:1987-2001
        const double x3[3] = {xxt[mm], yyt[mm], zzt[mm]}, f3[3] = {fx, fy, fz};
        dom.pp_add(iatm, eng_pp * 0.5, x3, f3);
        if (jatm <= dom.natms) dom.pp_add(jatm, eng_pp * 0.5, x3, f3);
      }
    }
  }
  parts[iatm].fxx = fix; parts[iatm].fyy = fiy; parts[iatm].fzz = fiz;
  double* s = dom.stress;
  s[1] += strs1; s[2] += strs2; s[3] += strs3; s[4] += strs2; s[5] += strs5; s[6] += strs6; s[7] += strs3; s[8] += strs6; s[9] += strs9;
  return {engvdw, virvdw};
}

// vdw.F90:1578-1788 vdw_forces_direct
Acc vdw_forces_direct(World& w, Dom& dom, int iatm, const double* xxt, const double* yyt, const double* zzt, const double* rrt) {
  Vdw& v = w.vdw;
  double engvdw = 0.0, virvdw = 0.0;
  double strs1 = 0, strs2 = 0, strs3 = 0, strs5 = 0, strs6 = 0, strs9 = 0;
  int idi = dom.ltg[iatm], ai = dom.ltype[iatm];
  CorePart* parts = dom.parts.data();
  double fix = parts[iatm].fxx, fiy = parts[iatm].fyy, fiz = parts[iatm].fzz;
  for (int mm = 1; mm <= dom.L(0, iatm); ++mm) {
    int jatm = dom.L(mm, iatm);
    int aj = dom.ltype[jatm];
    int key = (ai > aj) ? ai * (ai - 1) / 2 + aj : aj * (aj - 1) / 2 + ai;
    int k = v.list[key];
    double rrr = rrt[mm];
    int ityp = v.ltp[k];
    if (ityp != -1 && rrr < v.cutoff) {
      double r_rrr = 1.0 / rrr;
      double rsq = rrr * rrr;
      double r_rsq = r_rrr * r_rrr;
      EG eg = pot_energy(ityp, v.par(k), rrr);
      double eng = eg.energy + v.afs[k] * rrr + v.bfs[k];
      double gamma = eg.gamma * r_rsq - v.afs[k] * r_rrr;
      double fx = gamma * xxt[mm], fy = gamma * yyt[mm], fz = gamma * zzt[mm];
      fix = fix + fx; fiy = fiy + fy; fiz = fiz + fz;
      if (jatm > dom.natms && idi >= dom.ltg[jatm] && !w.collect_pp) eng = 0.0;   // vdw.F90:1707
      if (jatm <= dom.natms) {
        parts[jatm].fxx = parts[jatm].fxx - fx;
        parts[jatm].fyy = parts[jatm].fyy - fy;
        parts[jatm].fzz = parts[jatm].fzz - fz;
      }
      if (jatm <= dom.natms || idi < dom.ltg[jatm]) {
        engvdw = engvdw + eng;
        virvdw = virvdw - gamma * rsq;
        strs1 = strs1 + xxt[mm] * fx; strs2 = strs2 + xxt[mm] * fy; strs3 = strs3 + xxt[mm] * fz;
        strs5 = strs5 + yyt[mm] * fy; strs6 = strs6 + yyt[mm] * fz; strs9 = strs9 + zzt[mm] * fz;
      }
      if (w.collect_pp) {   // vdw.F90:1741-1755
        const double x3[3] = {xxt[mm], yyt[mm], zzt[mm]}, f3[3] = {fx, fy, fz};
        dom.pp_add(iatm, eng * 0.5, x3, f3);
        if (jatm <= dom.natms) dom.pp_add(jatm, eng * 0.5, x3, f3);
      }
    }
  }
  parts[iatm].fxx = fix; parts[iatm].fyy = fiy; parts[iatm].fzz = fiz;
  double* s = dom.stress;
  s[1] += strs1; s[2] += strs2; s[3] += strs3; s[4] += strs2; s[5] += strs5; s[6] += strs6; s[7] += strs3; s[8] += strs6; s[9] += strs9;
  return {engvdw, virvdw};
}

// ewald_spole.F90:58-242 ewald_real_forces_coul
Acc ewald_real_forces_coul(World& w, Dom& dom, int iatm, const double* x_pos, const double* y_pos, const double* z_pos,
                           const double* mod_dr_ij) {
  Ewald& e = w.ew;
  double engcpe_rl = 0.0, vircpe_rl = 0.0;
  double st[7] = {0}, ft[4] = {0};
  CorePart* parts = dom.parts.data();
  int global_id_i = dom.ltg[iatm];
  double atom_coeffs_i = parts[iatm].chge * e.scaling;
  if (std::fabs(atom_coeffs_i) < zero_plus) return {0.0, 0.0};
  const double* td = e.erfc_deriv.data();
  const double* te = e.erfc.data();
  for (int m = 1; m <= dom.L(0, iatm); ++m) {
    int jatm = dom.L(m, iatm);
    int global_id_j = dom.ltg[jatm];
    double mod_r_ij = mod_dr_ij[m];
    double prefac = parts[jatm].chge;
    if (std::fabs(prefac) > zero_plus && mod_r_ij < w.rcut) {
      double px = x_pos[m], py = y_pos[m], pz = z_pos[m];
      prefac = atom_coeffs_i * prefac;
      int nsi = f_int(mod_r_ij * e.recip_spacing);
      double diff = mod_r_ij * e.recip_spacing - (double)nsi;
      double p1 = td[nsi], p2 = td[nsi + 1], p3 = td[nsi + 2];
      if (nsi == 0) p1 = p1 * mod_r_ij;
      double tm1 = p1 + (p2 - p1) * diff;
      double tm2 = p2 + (p3 - p2) * (diff - 1.0);
      double erf_gamma = prefac * (tm1 + (tm2 - tm1) * diff * 0.5);
      double fcx = erf_gamma * px, fcy = erf_gamma * py, fcz = erf_gamma * pz;
      ft[1] = ft[1] + fcx; ft[2] = ft[2] + fcy; ft[3] = ft[3] + fcz;
      double e_comp = 0.0;
      if (jatm <= dom.natms || global_id_i < global_id_j || w.collect_pp) {   // ewald_spole.F90:155
        if (jatm <= dom.natms) {
          parts[jatm].fxx = parts[jatm].fxx - fcx;
          parts[jatm].fyy = parts[jatm].fyy - fcy;
          parts[jatm].fzz = parts[jatm].fzz - fcz;
        }
        nsi = f_int(mod_r_ij * e.recip_spacing);
        diff = mod_r_ij * e.recip_spacing - (double)nsi;
        p1 = te[nsi]; p2 = te[nsi + 1]; p3 = te[nsi + 2];
        if (nsi == 0) p1 = p1 * mod_r_ij;
        tm1 = p1 + (p2 - p1) * diff;
        tm2 = p2 + (p3 - p2) * (diff - 1.0);
        e_comp = prefac * (tm1 + (tm2 - tm1) * diff * 0.5);
      }
      if (jatm <= dom.natms || global_id_i < global_id_j) {
        engcpe_rl = engcpe_rl + e_comp;
        vircpe_rl = vircpe_rl - erf_gamma * (mod_r_ij * mod_r_ij);
        st[1] = st[1] + px * fcx; st[2] = st[2] + px * fcy; st[3] = st[3] + px * fcz;
        st[4] = st[4] + py * fcy; st[5] = st[5] + py * fcz; st[6] = st[6] + pz * fcz;
      }
      if (w.collect_pp) {   // ewald_spole.F90:205-215
        const double x3[3] = {px, py, pz}, f3[3] = {fcx, fcy, fcz};
        dom.pp_add(iatm, e_comp * 0.5, x3, f3);
        if (jatm <= dom.natms) dom.pp_add(jatm, e_comp * 0.5, x3, f3);
      }
    }
  }
  parts[iatm].fxx = parts[iatm].fxx + ft[1];
  parts[iatm].fyy = parts[iatm].fyy + ft[2];
  parts[iatm].fzz = parts[iatm].fzz + ft[3];
  double* s = dom.stress;
  s[1] += st[1]; s[2] += st[2]; s[3] += st[3]; s[4] += st[2]; s[5] += st[4]; s[6] += st[5]; s[7] += st[3]; s[8] += st[5]; s[9] += st[6];
  return {engcpe_rl, vircpe_rl};
}

// numerics.F90:265-290 calc_interp (the electro%calc_erfc / calc_erfc_deriv of the damped variants)
static inline double interp3_table(const double* t, double recip_spacing, double r) {
  int nsi = f_int(r * recip_spacing);
  double diff = r * recip_spacing - (double)nsi;
  double p1 = t[nsi], p2 = t[nsi + 1], p3 = t[nsi + 2];
  if (nsi == 0) p1 = p1 * r;
  double tm1 = p1 + (p2 - p1) * diff;
  double tm2 = p2 + (p3 - p2) * (diff - 1.0);
  return tm1 + (tm2 - tm1) * diff * 0.5;
}

// coul_spole.F90:567-729 coul_cp_forces, :731-897 coul_dddp_forces, :141-350 coul_fscp_forces, :352-565 coul_rfp_forces
Acc coul_direct_forces(World& w, Dom& dom, int iatm, const double* xxt, const double* yyt, const double* zzt, const double* rrt) {
  const Coulomb& c = w.coul;
  const Ewald& e = w.ew;   // tables of the damped variants
  double engcpe = 0.0, vircpe = 0.0;
  double st[7] = {0};
  CorePart* parts = dom.parts.data();
  int idi = dom.ltg[iatm];
  double chgea = parts[iatm].chge;
  if (!(std::fabs(chgea) > zero_plus)) return {0.0, 0.0};
  chgea = chgea * c.scaling;
  double fix = parts[iatm].fxx, fiy = parts[iatm].fyy, fiz = parts[iatm].fzz;
  const double cutoff_2 = w.rcut * w.rcut;
  for (int m = 1; m <= dom.L(0, iatm); ++m) {
    int jatm = dom.L(m, iatm);
    double chgprd = parts[jatm].chge;
    double rrr = rrt[m];
    if (std::fabs(chgprd) > zero_plus && rrr < w.rcut) {
      chgprd = chgprd * chgea;
      double rsq = rrr * rrr;
      double egamma = 0.0, coul = 0.0;
      switch (c.kind) {
        case COUL_CP:                                             // :647-649
          coul = chgprd / rrr;
          egamma = coul / (rrr * rrr);
          break;
        case COUL_DDDP:                                           // :812-815
          coul = chgprd / rsq;
          egamma = 2.0 * coul / rsq;
          break;
        case COUL_FSCP:                                           // :258-262
          if (c.damp) egamma = (interp3_table(e.erfc_deriv.data(), e.recip_spacing, rrr) - c.force_shift / rrr) * chgprd;
          else egamma = chgprd * (1.0 / rsq - c.force_shift) / rrr;
          break;
        case COUL_RFP:                                            // :470-475
          if (c.damp) egamma = (interp3_table(e.erfc_deriv.data(), e.recip_spacing, rrr) - c.force_shift / rrr - c.rf[0]) * chgprd;
          else egamma = chgprd * (1.0 / rsq / rrr - c.rf[0]);
          break;
      }
      double fx = egamma * xxt[m], fy = egamma * yyt[m], fz = egamma * zzt[m];
      fix = fix + fx; fiy = fiy + fy; fiz = fiz + fz;
      if (jatm <= dom.natms) {
        parts[jatm].fxx = parts[jatm].fxx - fx;
        parts[jatm].fyy = parts[jatm].fyy - fy;
        parts[jatm].fzz = parts[jatm].fzz - fz;
      }
      if (jatm <= dom.natms || idi < dom.ltg[jatm]) {
        if (c.kind == COUL_FSCP) {                                // :292-301
          if (c.damp) coul = (interp3_table(e.erfc.data(), e.recip_spacing, rrr) + c.force_shift * rrr + c.energy_shift) * chgprd;
          else coul = chgprd * (1.0 / rrr + c.force_shift * rrr + c.energy_shift);
        } else if (c.kind == COUL_RFP) {                          // :503-508
          if (c.damp) coul = (interp3_table(e.erfc.data(), e.recip_spacing, rrr) + c.force_shift * rrr + c.energy_shift +
                              c.rf[2] * (rsq - cutoff_2)) * chgprd;
          else coul = chgprd * (1.0 / rrr + c.rf[2] * rsq - c.rf[1]);
        }
        engcpe = engcpe + coul;
        if (c.kind == COUL_FSCP || c.kind == COUL_RFP) vircpe = vircpe - egamma * rsq;
        st[1] += xxt[m] * fx; st[2] += xxt[m] * fy; st[3] += xxt[m] * fz;
        st[4] += yyt[m] * fy; st[5] += yyt[m] * fz; st[6] += zzt[m] * fz;
      }
    }
  }
  parts[iatm].fxx = fix; parts[iatm].fyy = fiy; parts[iatm].fzz = fiz;
  if (c.kind == COUL_CP) vircpe = -engcpe;                         // :722
  if (c.kind == COUL_DDDP) vircpe = -2.0 * engcpe;                 // :890
  double* s = dom.stress;
  s[1] += st[1]; s[2] += st[2]; s[3] += st[3]; s[4] += st[2]; s[5] += st[4]; s[6] += st[5]; s[7] += st[3]; s[8] += st[5]; s[9] += st[6];
  return {engcpe, vircpe};
}

// ewald_spole.F90:479-679 ewald_excl_forces
Acc ewald_excl_forces(World& w, Dom& dom, int iatm, const double* xxt, const double* yyt, const double* zzt, const double* rrt) {
  const double a1 = 0.254829592, a2 = -0.284496736, a3 = 1.421413741, a4 = -1.453152027, a5 = 1.061405429,
               pp = 0.3275911, r10 = 0.1, r216 = 1.0 / 216.0, r42 = 1.0 / 42.0, rr3 = 1.0 / 3.0;
  double alpha = w.ew.alpha;
  double engcpe_ex = 0.0, vircpe_ex = 0.0;
  double strs1 = 0, strs2 = 0, strs3 = 0, strs5 = 0, strs6 = 0, strs9 = 0;
  CorePart* parts = dom.parts.data();
  int idi = dom.ltg[iatm];
  double chgea = parts[iatm].chge;
  if (std::fabs(chgea) > zero_plus) {
    chgea = chgea * w.ew.scaling;
    double fix = parts[iatm].fxx, fiy = parts[iatm].fyy, fiz = parts[iatm].fzz;
    int limit = dom.L(-1, iatm) - dom.L(0, iatm);
    for (int m = 1; m <= limit; ++m) {
      int jatm = dom.L(dom.L(0, iatm) + m, iatm);
      double chgprd = parts[jatm].chge;
      double rrr = rrt[m];
      if (std::fabs(chgprd) > zero_plus && rrr < w.rcut) {
        chgprd = chgprd * chgea;
        double rsq = rrr * rrr;
        double alpr = rrr * alpha;
        double alpr2 = alpr * alpr;
        double erfr, egamma;
        if (alpr < 1.0e-2) {
          erfr = 2.0 * chgprd * (alpha / sqrpi) * (1.0 + alpr2 * (-rr3 + alpr2 * (r10 + alpr2 * (-r42 + alpr2 * r216))));
          egamma = -4.0 * chgprd * (powi(alpha, 3) / sqrpi) *
                   (rr3 + alpr2 * (-2.0 * r10 + alpr2 * (3.0 * r42 - 4.0 * alpr2 * r216)));
        } else {
          double ar = alpha * rrr;
          double exp1 = std::exp(-(ar * ar));
          double tt = 1.0 / (1.0 + pp * alpha * rrr);
          erfr = chgprd * (1.0 - tt * (a1 + tt * (a2 + tt * (a3 + tt * (a4 + tt * a5)))) * exp1) / rrr;
          egamma = -(erfr - 2.0 * chgprd * (alpha / sqrpi) * exp1) / rsq;
        }
        double fx = egamma * xxt[m], fy = egamma * yyt[m], fz = egamma * zzt[m];
        fix = fix + fx; fiy = fiy + fy; fiz = fiz + fz;
        if (jatm <= dom.natms) {
          parts[jatm].fxx = parts[jatm].fxx - fx;
          parts[jatm].fyy = parts[jatm].fyy - fy;
          parts[jatm].fzz = parts[jatm].fzz - fz;
        }
        if (jatm <= dom.natms || idi < dom.ltg[jatm]) {
          engcpe_ex = engcpe_ex - erfr;
          vircpe_ex = vircpe_ex - egamma * rsq;
          strs1 = strs1 + xxt[m] * fx; strs2 = strs2 + xxt[m] * fy; strs3 = strs3 + xxt[m] * fz;
          strs5 = strs5 + yyt[m] * fy; strs6 = strs6 + yyt[m] * fz; strs9 = strs9 + zzt[m] * fz;
        }
      }
    }
    parts[iatm].fxx = fix; parts[iatm].fyy = fiy; parts[iatm].fzz = fiz;
    double* s = dom.stress;
    s[1] += strs1; s[2] += strs2; s[3] += strs3; s[4] += strs2; s[5] += strs5; s[6] += strs6; s[7] += strs3; s[8] += strs6; s[9] += strs9;
  }
  return {engcpe_ex, vircpe_ex};
}

// rdfs.F90:146-212 rdf_collect (active pairs) and :880-946 rdf_excl_collect (excluded pairs), as two_body_forces calls them
// (two_body.F90:523, :581) with the same distance arrays.  rdf(ll,kk) is (1:max_grid, 1:n_pairs), column-major, counts.
void rdf_collect_domain(World& w, Dom& dom, const int* rdf_list /*1-based keys*/, int n_pairs, int max_grid, double* rdf) {
  const double rdelr = (double)max_grid / w.rcut;
  CorePart* parts = dom.parts.data();
  auto collect = [&](int iatm, int first, int count) {
    int idi = dom.ltg[iatm], ai = dom.ltype[iatm];
    for (int m = 1; m <= count; ++m) {
      int jatm = dom.L(first + m, iatm);
      int aj = dom.ltype[jatm];
      if (jatm <= dom.natms || idi < dom.ltg[jatm]) {
        int keyrdf = (std::max(ai, aj) * (std::max(ai, aj) - 1)) / 2 + std::min(ai, aj);
        int kk = rdf_list[keyrdf - 1];
        if (kk > 0 && kk <= n_pairs) {
          double xx = parts[iatm].xxx - parts[jatm].xxx, yy = parts[iatm].yyy - parts[jatm].yyy, zz = parts[iatm].zzz - parts[jatm].zzz;
          double rrr = std::sqrt(xx * xx + yy * yy + zz * zz);
          if (rrr < w.rcut) {
            int ll = std::min(1 + f_int(rrr * rdelr), max_grid);
            rdf[(size_t)(kk - 1) * max_grid + (ll - 1)] += 1.0;
          }
        }
      }
    }
  };
  for (int i = 1; i <= dom.natms; ++i) {
    collect(i, 0, dom.L(0, i));                                         // rdf_collect
    if (w.lbook) collect(i, dom.L(0, i), dom.L(-1, i) - dom.L(0, i));   // rdf_excl_collect
    if (w.megfrz != 0) collect(i, dom.L(-1, i), dom.L(-2, i) - dom.L(-1, i));   // rdf_frzn_collect (two_body.F90:615-652, rdfs.F90:948-1018)
  }
}

// two_body.F90:339-525 + :552-606 : the two outer loops (per domain).  Forces are ADDED into parts%f.
void two_body_forces(World& w, Dom& dom) {
  const int ml = dom.max_list;
  std::vector<double> xxt(ml + 1), yyt(ml + 1), zzt(ml + 1), rrt(ml + 1);
  for (int k = 1; k <= 9; ++k) dom.stress[k] = 0.0;
  if (w.collect_pp) { dom.pp_energy.assign((size_t)dom.natms + 1, 0.0); dom.pp_stress.assign(((size_t)dom.natms + 1) * 9, 0.0); }
  double engvdw = 0, virvdw = 0, engcpe_rl = 0, vircpe_rl = 0, engcpe_ex = 0, vircpe_ex = 0;
  CorePart* parts = dom.parts.data();
  for (int i = 1; i <= dom.natms; ++i) {
    int limit = dom.L(0, i);
    for (int k = 1; k <= limit; ++k) {
      int j = dom.L(k, i);
      xxt[k] = parts[i].xxx - parts[j].xxx;
      yyt[k] = parts[i].yyy - parts[j].yyy;
      zzt[k] = parts[i].zzz - parts[j].zzz;
      rrt[k] = std::sqrt(xxt[k] * xxt[k] + yyt[k] * yyt[k] + zzt[k] * zzt[k]);
    }
    if (!w.vdw.no_vdw && w.vdw.n_vdw > 0) {
      Acc a = w.vdw.l_direct ? vdw_forces_direct(w, dom, i, xxt.data(), yyt.data(), zzt.data(), rrt.data())
                             : vdw_forces_tab(w, dom, i, xxt.data(), yyt.data(), zzt.data(), rrt.data());
      engvdw = engvdw + a.eng; virvdw = virvdw + a.vir;
    }
    if (w.ew.active) {
      Acc a = ewald_real_forces_coul(w, dom, i, xxt.data(), yyt.data(), zzt.data(), rrt.data());
      engcpe_rl = engcpe_rl + a.eng; vircpe_rl = vircpe_rl + a.vir;
    } else if (w.coul.kind != COUL_NONE) {   // two_body.F90:480-514
      Acc a = coul_direct_forces(w, dom, i, xxt.data(), yyt.data(), zzt.data(), rrt.data());
      engcpe_rl = engcpe_rl + a.eng; vircpe_rl = vircpe_rl + a.vir;
    }
  }
  if (w.lbook && w.ew.active) {   // two_body.F90:552-606
    for (int i = 1; i <= dom.natms; ++i) {
      int limit = dom.L(-1, i) - dom.L(0, i);
      if (limit > 0) {
        for (int k = 1; k <= limit; ++k) {
          int j = dom.L(dom.L(0, i) + k, i);
          xxt[k] = parts[i].xxx - parts[j].xxx;
          yyt[k] = parts[i].yyy - parts[j].yyy;
          zzt[k] = parts[i].zzz - parts[j].zzz;
        }
        for (int k = 1; k <= limit; ++k) rrt[k] = std::sqrt(xxt[k] * xxt[k] + yyt[k] * yyt[k] + zzt[k] * zzt[k]);
        Acc a = ewald_excl_forces(w, dom, i, xxt.data(), yyt.data(), zzt.data(), rrt.data());
        engcpe_ex = engcpe_ex + a.eng; vircpe_ex = vircpe_ex + a.vir;
      }
    }
  }
  dom.engvdw = engvdw; dom.virvdw = virvdw; dom.engcpe_rl = engcpe_rl; dom.vircpe_rl = vircpe_rl;
  dom.engcpe_ex = engcpe_ex; dom.vircpe_ex = vircpe_ex;
}

// deport_data.F90:2870-3202 relocate_particles + :81-960 deport_atomic_data (positions, velocities, forces,
// ltg, lsite, ixyz and list_excl rows travel; the other bookkeeping payloads are outside the hot path)
int relocate_particles(World& w) {
  if (w.P == 1) {
    pbcshift(w.imcon, w.cell, w.d[0].natms, w.d[0].parts.data());   // :3190
    return 0;
  }
  double rcell[10], det;
  invert(w.cell, rcell, det);
#pragma omp parallel for schedule(static, 1) num_threads(w.nthreads) if (w.nthreads > 1)
  for (int rr = 0; rr < (int)w.d.size(); ++rr) {   // :2981-3025
    Dom& dom = w.d[rr];
    for (int i = 1; i <= dom.natms; ++i) {
      dom.ixyz[i] = 0;
      const CorePart& p = dom.parts[i];
      double x = rcell[1] * p.xxx + rcell[4] * p.yyy + rcell[7] * p.zzz;
      double y = rcell[2] * p.xxx + rcell[5] * p.yyy + rcell[8] * p.zzz;
      double z = rcell[3] * p.xxx + rcell[6] * p.yyy + rcell[9] * p.zzz;
      int ipx = f_int((x + 0.5) * dom.nx_real), ipy = f_int((y + 0.5) * dom.ny_real), ipz = f_int((z + 0.5) * dom.nz_real);
      if (dom.idx == 0) { if (x < -half_plus) dom.ixyz[i] += 1; } else { if (ipx < dom.idx) dom.ixyz[i] += 1; }
      if (dom.idx == dom.nx - 1) { if (x >= half_minus) dom.ixyz[i] += 2; } else { if (ipx > dom.idx) dom.ixyz[i] += 2; }
      if (dom.idy == 0) { if (y < -half_plus) dom.ixyz[i] += 10; } else { if (ipy < dom.idy) dom.ixyz[i] += 10; }
      if (dom.idy == dom.ny - 1) { if (y >= half_minus) dom.ixyz[i] += 20; } else { if (ipy > dom.idy) dom.ixyz[i] += 20; }
      if (dom.idz == 0) { if (z < -half_plus) dom.ixyz[i] += 100; } else { if (ipz < dom.idz) dom.ixyz[i] += 100; }
      if (dom.idz == dom.nz - 1) { if (z >= half_minus) dom.ixyz[i] += 200; } else { if (ipz > dom.idz) dom.ixyz[i] += 200; }
    }
  }
  const int mdirs[6] = {-1, 1, -2, 2, -3, 3};
  const int rec = 12;   // x,y,z,vx,vy,vz,fx,fy,fz,ltg,lsite,ixyz  (deport_data.F90:290-325)
  for (int q = 0; q < 6; ++q) {
    int mdir = mdirs[q];
    std::vector<std::vector<int>> excl_rows(w.d.size());
#pragma omp parallel for schedule(static, 1) num_threads(w.nthreads) if (w.nthreads > 1)
    for (int r = 0; r < (int)w.d.size(); ++r) {
      Dom& dom = w.d[r];
      DirSet s = dir_settings(w, dom, mdir);
      dom.sendbuf.clear();
      std::vector<int> ind_on(1, 0), ind_off(1, 0);
      for (int i = 1; i <= dom.natms; ++i) {
        bool stay = false;
        if (dom.ixyz[i] == 0) stay = true;
        else {
          int ix = dom.ixyz[i] % 10;
          int iy = (dom.ixyz[i] - ix) % 100;
          int iz = (dom.ixyz[i] - (ix + iy)) % 1000;
          int j = ix * s.kx + iy * s.ky + iz * s.kz;
          if (j == s.jxyz) dom.ixyz[i] = dom.ixyz[i] - s.jxyz; else stay = true;
        }
        if (stay) { ind_on.push_back(i); ind_on[0]++; }
        else {
          ind_off.push_back(i); ind_off[0]++;
          const CorePart& p = dom.parts[i];
          double px = p.xxx, py = p.yyy, pz = p.zzz;
          if (s.lwrap) { px = p.xxx + s.xadd; py = p.yyy + s.yadd; pz = p.zzz + s.zadd; }
          double recd[rec] = {px, py, pz, dom.vxx[i], dom.vyy[i], dom.vzz[i], p.fxx, p.fyy, p.fzz,
                              (double)dom.ltg[i], (double)dom.lsite[i], (double)dom.ixyz[i]};
          dom.sendbuf.insert(dom.sendbuf.end(), recd, recd + rec);
          if (w.lbook)
            for (int k = 0; k <= dom.max_exclude; ++k) excl_rows[r].push_back(dom.LE(k, i));
        }
      }
      // restack :822-925
      int k = ind_on[0], l = ind_off[0];
      for (int ii = 1; ii <= l; ++ii) {
        int keep = ind_off[ii];
        if (k < 1 || keep > ind_on[k]) break;
        int i = ind_on[k - ii + 1];
        dom.parts[keep] = dom.parts[i];
        dom.vxx[keep] = dom.vxx[i]; dom.vyy[keep] = dom.vyy[i]; dom.vzz[keep] = dom.vzz[i];
        dom.ltg[keep] = dom.ltg[i]; dom.lsite[keep] = dom.lsite[i]; dom.ixyz[keep] = dom.ixyz[i];
        if (w.lbook)
          for (int kk = 0; kk <= dom.max_exclude; ++kk) dom.LE(kk, keep) = dom.LE(kk, i);
      }
      dom.natms = k;   // keep
    }
#pragma omp parallel for schedule(static, 1) num_threads(w.nthreads) if (w.nthreads > 1)
    for (int r = 0; r < (int)w.d.size(); ++r) {   // receive
      Dom& dom = w.d[r];
      DirSet s = dir_settings(w, dom, mdir);
      const std::vector<double>& buf = w.d[s.kd].sendbuf;
      const std::vector<int>& ex = excl_rows[s.kd];
      int nin = (int)buf.size() / rec;
      dom.ensure(dom.natms + nin);
      if (w.lbook && (int)dom.list_excl.size() < (dom.natms + nin) * (dom.max_exclude + 1))
        dom.list_excl.resize((size_t)(dom.natms + nin + 64) * (dom.max_exclude + 1), 0);
      for (int n = 0; n < nin; ++n) {
        int i = ++dom.natms;
        const double* b = &buf[(size_t)n * rec];
        CorePart& p = dom.parts[i];
        p.xxx = b[0]; p.yyy = b[1]; p.zzz = b[2];
        dom.vxx[i] = b[3]; dom.vyy[i] = b[4]; dom.vzz[i] = b[5];
        p.fxx = b[6]; p.fyy = b[7]; p.fzz = b[8];
        dom.ltg[i] = f_nint(b[9]); dom.lsite[i] = f_nint(b[10]); dom.ixyz[i] = f_nint(b[11]);
        if (w.lbook)
          for (int k = 0; k <= dom.max_exclude; ++k) dom.LE(k, i) = ex[(size_t)n * (dom.max_exclude + 1) + k];
      }
    }
  }
  int total = 0;
  for (Dom& dom : w.d) {
    for (int i = 1; i <= dom.natms; ++i) {
      if (dom.ixyz[i] != 0) return 58;
      dom.ltype[i] = w.sites.type_site[dom.lsite[i]];          // :3062-3069
      dom.parts[i].chge = w.sites.charge_site[dom.lsite[i]];
      dom.lfrzn[i] = w.sites.freeze_site[dom.lsite[i]];
    }
    dom.nlast = dom.natms;
    total += dom.natms;
  }
  if (total != w.megatm) return 58;
  return 0;
}

}  // namespace

// =============================================================================
// C API (ctypes).  Plain pointers, 0-based arrays on the outside.
// =============================================================================
// ---------------------------------------------------------------- vdw.F90:617-967 vdw_lrc
// Long-range corrections to energy and virial in a 3D periodic system, for the potentials of this path (TABLE entries carry
// theirs in param(1:2), vdw.F90:1166-1170).  num_type / numfrz: atoms and frozen atoms per type, already global (gsum, :672).
// list / ltp / param as in Vdw (1-based views); returns elrc, vlrc.
void vdw_lrc(int ntype_atom, const int* list /*1..*/, const int* ltp /*1..*/, const double* param /*(1:7,1:max_vdw)*/, double rvdw,
             bool l_force_shift, int imcon, double volm, const double* num_type /*1..*/, const double* numfrz /*1..*/, double* elrc_out,
             double* vlrc_out) {
  double twopi = 2.0 * pi;
  double plrc = 0.0, elrc = 0.0;                                             // :659-660
  if (!l_force_shift) {                                                      // :662
    if (imcon != 0 && imcon != 6) {                                          // :676
      int ivdw = 0;
      for (int i = 1; i <= ntype_atom; ++i) {
        for (int j = 1; j <= i; ++j) {
          double eadd = 0.0, padd = 0.0;
          ivdw = ivdw + 1;
          int k = list[ivdw];
          int keypot = ltp[k];
          const double* prm = param + (size_t)(k - 1) * 7 - 1;               // prm[1..7]
          double r = rvdw;
          if (keypot == 0) {                                                 // VDW_TAB :691-696
            eadd = prm[1];
            padd = -prm[2];
          } else if (keypot == 1) {                                          // VDW_12_6 :698-707
            double a = prm[1], b = prm[2];
            eadd = a / (9.0 * powi(r, 9)) - b / (3.0 * powi(r, 3));
            padd = 12.0 * a / (9.0 * powi(r, 9)) - 6.0 * b / (3.0 * powi(r, 3));
          } else if (keypot == 2) {                                          // VDW_LENNARD_JONES :709-718
            double eps = prm[1], sig = prm[2];
            eadd = 4.0 * eps * (powi(sig, 12) / (9.0 * powi(r, 9)) - powi(sig, 6) / (3.0 * powi(r, 3)));
            padd = 8.0 * eps * (6.0 * powi(sig, 12) / (9.0 * powi(r, 9)) - powi(sig, 6) / (powi(r, 3)));
          } else if (keypot == 4) {                                          // VDW_BUCKINGHAM :735-743
            double c = prm[3];
            eadd = -c / (3.0 * powi(r, 3));
            padd = -2.0 * c / (powi(r, 3));
          } else if (keypot == 5) {                                          // VDW_BORN_HUGGINS_MEYER :745-754
            double c = prm[4], d = prm[5];
            eadd = -c / (3.0 * powi(r, 3)) - d / (5.0 * powi(r, 5));
            padd = -2.0 * c / (powi(r, 3)) - 8.0 * d / (5.0 * powi(r, 5));
          }                                                                  // VDW_NULL and the rest: 0
          if (i != j) {                                                      // :936-939
            eadd = eadd * 2.0;
            padd = padd * 2.0;
          }
          double denprd = twopi * (num_type[i] * num_type[j] - numfrz[i] * numfrz[j]) / powi(volm, 2);   // :941
          elrc = elrc + volm * denprd * eadd;                                // :945
          plrc = plrc + denprd * padd / 3.0;                                 // :946
        }
      }
    }
  }
  *elrc_out = elrc;
  *vlrc_out = plrc * (-3.0 * volm);                                          // :962
}

// two_body.F90:672-790, the terms this path feeds: in = engvdw, virvdw, engcpe_rl, vircpe_rl, engcpe_ex, vircpe_ex AFTER gsum
// (:729), plus the reciprocal-space pair (engcpe_rc, vircpe_rc) the caller's SPME produced (0 without it).  Adds to
// tot = {engcpe, vircpe, engsrp, virsrp} and to this rank's stress(1:9).
void two_body_epilogue(const double* in6, double engcpe_rc, double vircpe_rc, bool spme_or_poisson, double sumchg, double alpha, double eps,
                       double volm, double elrc, double vlrc, int mxnode, double* tot4, double* stress /*1..9*/) {
  double engcpe_nz = 0.0, vircpe_nz = 0.0;
  if (spme_or_poisson) {
    if (std::fabs(sumchg) > 1.0e-6) {                                        // :691-696 (Fuchs)
      double factor_nz = -0.5 * (pi * r4pie0 / eps) * powi(sumchg / alpha, 2);
      engcpe_nz = factor_nz / volm;
      vircpe_nz = -3.0 * engcpe_nz;
    }
  }
  double engvdw = in6[0], virvdw = in6[1], engcpe_rl = in6[2], vircpe_rl = in6[3], engcpe_ex = in6[4], vircpe_ex = in6[5];
  double engcpe_ch = 0.0, vircpe_ch = 0.0, engcpe_fr = 0.0, vircpe_fr = 0.0, vircpe_dt = 0.0;   // core-shell, frozen, multipoles: not this path
  tot4[0] = tot4[0] + engcpe_rc + engcpe_rl + engcpe_ch + engcpe_ex + engcpe_fr + engcpe_nz;    // :766
  tot4[1] = tot4[1] + vircpe_rc + vircpe_rl + vircpe_ch + vircpe_ex + vircpe_fr + vircpe_nz + vircpe_dt;   // :767
  double tmp = -vircpe_nz / (3.0 * (double)mxnode);                          // :772-775
  stress[1] = stress[1] + tmp;
  stress[5] = stress[5] + tmp;
  stress[9] = stress[9] + tmp;
  tot4[2] = tot4[2] + (engvdw + elrc);                                       // :780 (no KIM, no metal)
  tot4[3] = tot4[3] + (virvdw + vlrc);                                       // :781
  tmp = -(vlrc + 0.0) / (3.0 * (double)mxnode);                              // :787-790
  stress[1] = stress[1] + tmp;
  stress[5] = stress[5] + tmp;
  stress[9] = stress[9] + tmp;
}

extern "C" {

void ora_dcell(const double* cell9, double* out10) {
  double a[10], b[11];
  for (int i = 0; i < 9; ++i) a[i + 1] = cell9[i];
  dcell(a, b);
  for (int i = 0; i < 10; ++i) out10[i] = b[i + 1];
}
void ora_invert(const double* a9, double* b9, double* det) {
  double a[10], b[10];
  for (int i = 0; i < 9; ++i) a[i + 1] = a9[i];
  invert(a, b, *det);
  for (int i = 0; i < 9; ++i) b9[i] = b[i + 1];
}
void ora_vdw_lrc(int ntype_atom, const int* list0, const int* ltp0, const double* param, double rvdw, int l_force_shift, int imcon,
                 double volm, const double* num_type0, const double* numfrz0, double* elrc, double* vlrc) {
  vdw_lrc(ntype_atom, list0 - 1, ltp0 - 1, param, rvdw, l_force_shift != 0, imcon, volm, num_type0 - 1, numfrz0 - 1, elrc, vlrc);
}
void ora_two_body_epilogue(const double* in6, double engcpe_rc, double vircpe_rc, int spme, double sumchg, double alpha, double eps,
                           double volm, double elrc, double vlrc, int mxnode, double* tot4, double* stress9) {
  two_body_epilogue(in6, engcpe_rc, vircpe_rc, spme != 0, sumchg, alpha, eps, volm, elrc, vlrc, mxnode, tot4, stress9 - 1);
}
int ora_match(int n, int ind_top, const int* list0) { return match(n, ind_top, list0 - 1) ? 1 : 0; }
void ora_images(int imcon, const double* cell9, int n, double* x, double* y, double* z) {
  double c[10];
  for (int i = 0; i < 9; ++i) c[i + 1] = cell9[i];
  images(imcon, c, n, x - 1, y - 1, z - 1);
}
double ora_calc_erfc(double x) { return calc_erfc(x); }
int ora_max_grid(double rcut) { return std::max(1004, f_nint(rcut / delr_max) + 4); }   // bounds.F90:811,820
// control.F90:1709-1710
double ora_ewald_alpha(double precision, double rcut) {
  double tol = std::sqrt(std::fabs(std::log(precision * rcut)));
  return std::sqrt(std::fabs(std::log(precision * rcut * tol))) / rcut;
}
// bounds.F90:907 (fdens = density * densvar-factor supplied by caller)
int ora_max_list(double fdens, double rx) { return f_nint(fdens * (7.5 / 3.0) * pi * (rx * rx * rx)); }

void ora_pot_energy(int keypot, const double* param7, double r, double* e, double* g) {
  EG z = pot_energy(keypot, param7 - 1, r);
  *e = z.energy; *g = z.gamma;
}
// erfc tables: out arrays have nsamples+1 entries, index i = table(i), index 0 = 0 (out-of-bounds slot)
void ora_erfcgen(double rcut, double alpha, int nsamples, double* erfc_t, double* deriv_t, double* recip_spacing) {
  double sp, rc;
  erfcgen(rcut, alpha, nsamples, erfc_t, deriv_t, sp, rc);
  *recip_spacing = rc;
}
// vdw_generate for one potential: outputs (0:max_grid)
void ora_vdw_generate(int keypot, const double* param7, double rvdw, int max_grid, double* tab_pot, double* tab_force) {
  Vdw v;
  v.n_vdw = v.max_vdw = 1; v.max_grid = max_grid; v.cutoff = rvdw;
  v.ltp.assign(2, keypot);
  v.param.assign(7, 0.0);
  for (int i = 0; i < 7; ++i) v.param[i] = param7[i];
  v.tab_potential.assign(max_grid + 1, 0.0); v.tab_force.assign(max_grid + 1, 0.0);
  vdw_generate_one(v, 1);
  std::memcpy(tab_pot, v.tab_potential.data(), sizeof(double) * (max_grid + 1));
  std::memcpy(tab_force, v.tab_force.data(), sizeof(double) * (max_grid + 1));
}
void ora_vdw_table_regrid(const double* buffer0 /*ngrid values*/, int ngrid, double delpot, double rvdw, int max_grid,
                          int is_force, double engunit, double* tab) {
  vdw_table_regrid(buffer0 - 1, ngrid, delpot, rvdw, max_grid, is_force != 0, engunit, tab);
}
void ora_vdw_direct_fs(int keypot, const double* param7, double rvdw, double* afs, double* bfs) {
  Vdw v;
  v.n_vdw = v.max_vdw = 1; v.cutoff = rvdw; v.l_force_shift = true;
  v.ltp.assign(2, keypot);
  v.param.assign(7, 0.0);
  for (int i = 0; i < 7; ++i) v.param[i] = param7[i];
  vdw_direct_fs_generate(v);
  *afs = v.afs[1]; *bfs = v.bfs[1];
}

// test_vdw.F90:93-177 fake 2-atom system + vdw_forces_direct call (:83) for one potential.
void ora_kat_vdw_direct(int keypot, const double* param7, double* eng, double* vir) {
  World w;
  w.P = 1; w.rcut = 10.0; w.rx = 10.0;
  Vdw& v = w.vdw;
  v.no_vdw = false; v.l_direct = true; v.ntypes = 1; v.n_vdw = 1; v.max_vdw = 1; v.cutoff = 10.0;
  v.list.assign(2, 1); v.ltp.assign(2, keypot);
  v.param.assign(7, 0.0);
  for (int i = 0; i < 7; ++i) v.param[i] = param7[i];
  v.afs.assign(2, 0.0); v.bfs.assign(2, 0.0);
  w.d.resize(1);
  Dom& dom = w.d[0];
  dom.natms = 10; dom.ensure(2); dom.max_list = 2;
  dom.list.assign(2 * (2 + 4), 0);
  dom.L(0, 1) = 1; dom.L(1, 1) = 2;
  for (int i = 1; i <= 2; ++i) { dom.parts[i] = CorePart{0, 0, 0, 0, 0, 0, 0, 0, 0}; dom.ltype[i] = 1; dom.ltg[i] = i; }
  dom.parts[2].xxx = 1.0;
  double ones[3] = {0.0, 1.0, 1.0};
  Acc a = vdw_forces_direct(w, dom, 1, ones, ones, ones, ones);
  *eng = a.eng; *vir = a.vir;
}

// ---- world API -----------------------------------------------------------------------------------------------
void* ora_world_create(int P, const double* cell9, int imcon) {
  World* w = new World();
  w->P = P; w->imcon = imcon;
  for (int i = 0; i < 9; ++i) w->cell[i + 1] = cell9[i];
  double celprp[11];
  dcell(w->cell, celprp);
  w->d.resize(P);
  for (int r = 0; r < P; ++r) map_domains(imcon, celprp[7], celprp[8], celprp[9], P, r, w->d[r]);
  return w;
}
void ora_world_destroy(void* h) { delete (World*)h; }
void ora_world_dd(void* h, int rank, int* out6, int* map26) {
  World* w = (World*)h;
  Dom& d = w->d[rank];
  out6[0] = d.nx; out6[1] = d.ny; out6[2] = d.nz; out6[3] = d.idx; out6[4] = d.idy; out6[5] = d.idz;
  if (map26) for (int i = 0; i < 26; ++i) map26[i] = d.map[i + 1];
}
void ora_world_set_cutoffs(void* h, double rcut, double padding, double pdplnc, const double* ecw3) {
  World* w = (World*)h;
  w->rcut = rcut; w->padding = padding; w->rx = rcut + padding; w->pdplnc = pdplnc;
  for (int i = 0; i < 3; ++i) w->ecw[i + 1] = ecw3 ? ecw3[i] : 0.0;
}
void ora_world_set_sites(void* h, int nsites, const int* type_site, const double* charge_site, const int* freeze_site) {
  World* w = (World*)h;
  w->sites.type_site.assign(nsites + 1, 0); w->sites.charge_site.assign(nsites + 1, 0.0); w->sites.freeze_site.assign(nsites + 1, 0);
  for (int i = 0; i < nsites; ++i) {
    w->sites.type_site[i + 1] = type_site[i];
    w->sites.charge_site[i + 1] = charge_site[i];
    w->sites.freeze_site[i + 1] = freeze_site[i];
  }
}
// tables exactly as the C-ABI of the product receives them (column-major (0:max_grid,1:max_vdw))
void ora_world_set_vdw(void* h, int ntypes, const int* vdw_list /*ntab*/, int max_vdw, int n_vdw, const int* ltp, int max_grid,
                       const double* tab_potential, const double* tab_force, double rvdw, int force_shift, int direct,
                       const double* param /*7*max_vdw*/, const double* afs, const double* bfs) {
  World* w = (World*)h;
  Vdw& v = w->vdw;
  v.no_vdw = (n_vdw <= 0);
  v.ntypes = ntypes; v.n_vdw = n_vdw; v.max_vdw = max_vdw; v.max_grid = max_grid; v.cutoff = rvdw;
  v.l_force_shift = force_shift != 0; v.l_direct = direct != 0;
  int ntab = ntypes * (ntypes + 1) / 2;
  v.list.assign(ntab + 1, 0);
  for (int i = 0; i < ntab; ++i) v.list[i + 1] = vdw_list[i];
  v.ltp.assign(max_vdw + 1, -1);
  for (int i = 0; i < max_vdw; ++i) v.ltp[i + 1] = ltp[i];
  size_t n = (size_t)(max_grid + 1) * max_vdw;
  if (tab_potential) { v.tab_potential.assign(tab_potential, tab_potential + n); v.tab_force.assign(tab_force, tab_force + n); }
  else { v.tab_potential.assign(n, 0.0); v.tab_force.assign(n, 0.0); }
  v.param.assign((size_t)7 * max_vdw, 0.0);
  if (param) v.param.assign(param, param + (size_t)7 * max_vdw);
  v.afs.assign(max_vdw + 1, 0.0); v.bfs.assign(max_vdw + 1, 0.0);
  for (int i = 0; i < max_vdw; ++i) { if (afs) v.afs[i + 1] = afs[i]; if (bfs) v.bfs[i + 1] = bfs[i]; }
  if (max_grid > 4) { v.dlrpot = v.cutoff / (double)(v.max_grid - 4); v.rdr = 1.0 / v.dlrpot; }   // vdw.F90:1836-1837
}
void ora_world_set_ewald(void* h, int active, double alpha, double scaling, int nsamples, const double* erfc_t,
                         const double* deriv_t /* nsamples+1 each, [0] unused */, double recip_spacing) {
  World* w = (World*)h;
  Ewald& e = w->ew;
  e.active = active != 0; e.alpha = alpha; e.scaling = scaling; e.nsamples = nsamples; e.recip_spacing = recip_spacing;
  if (active) { e.erfc.assign(erfc_t, erfc_t + nsamples + 1); e.erfc_deriv.assign(deriv_t, deriv_t + nsamples + 1); }
}
// direct-space Coulomb variant (kind 1..4 as COUL_*); damped variants carry the erfc tables generated with alpha = damping
void ora_world_set_coulomb(void* h, int kind, int damp, double scaling, double force_shift, double energy_shift, const double* rf3,
                           int nsamples, const double* erfc_t, const double* deriv_t, double recip_spacing) {
  World* w = (World*)h;
  Coulomb& c = w->coul;
  c.kind = kind; c.damp = damp != 0; c.scaling = scaling; c.force_shift = force_shift; c.energy_shift = energy_shift;
  for (int k = 0; k < 3; ++k) c.rf[k] = rf3 ? rf3[k] : 0.0;
  Ewald& e = w->ew;
  e.active = false;
  if (c.damp) {
    e.nsamples = nsamples; e.recip_spacing = recip_spacing;
    e.erfc.assign(erfc_t, erfc_t + nsamples + 1); e.erfc_deriv.assign(deriv_t, deriv_t + nsamples + 1);
  }
}
// exclusions by global id: rows (0:max_exclude) for gid 1..megatm ; row[0]=count, sorted ascending ids
void ora_world_set_excl(void* h, int max_exclude, const int* excl /* (max_exclude+1)*megatm */, int megatm) {
  World* w = (World*)h;
  w->lbook = true; w->max_exclude = max_exclude;
  w->excl_global.assign(excl, excl + (size_t)(max_exclude + 1) * megatm);
}
void ora_world_set_max_list(void* h, int max_list) {
  World* w = (World*)h;
  for (Dom& d : w->d) d.max_list = max_list;
}

// configuration.F90:1183-1205: fold into [-0.5,0.5), recompute Cartesian from cell*s, assign to idm; local order =
// CONFIG order filtered by domain.  fold=0 keeps positions exactly as given (C-ABI harness convention).
int ora_world_load(void* h, int megatm, const double* xyz /*3*megatm*/, const double* vel /*or null*/, const int* lsite,
                   int fold) {
  World* w = (World*)h;
  w->megatm = megatm;
  double rcell[10], det;
  invert(w->cell, rcell, det);
  for (Dom& d : w->d) { d.natms = 0; d.nlast = 0; }
  w->megfrz = 0;
  Dom& d0 = w->d[0];
  for (int g = 1; g <= megatm; ++g) {
    double axx = xyz[3 * (g - 1)], ayy = xyz[3 * (g - 1) + 1], azz = xyz[3 * (g - 1) + 2];
    double sxx = rcell[1] * axx + rcell[4] * ayy + rcell[7] * azz;
    double syy = rcell[2] * axx + rcell[5] * ayy + rcell[8] * azz;
    double szz = rcell[3] * axx + rcell[6] * ayy + rcell[9] * azz;
    sxx = sxx - f_anint(sxx); if (sxx >= half_minus) sxx = -sxx;
    syy = syy - f_anint(syy); if (syy >= half_minus) syy = -syy;
    szz = szz - f_anint(szz); if (szz >= half_minus) szz = -szz;
    if (fold) {
      axx = w->cell[1] * sxx + w->cell[4] * syy + w->cell[7] * szz;
      ayy = w->cell[2] * sxx + w->cell[5] * syy + w->cell[8] * szz;
      azz = w->cell[3] * sxx + w->cell[6] * syy + w->cell[9] * szz;
    }
    int ipx = f_int((sxx + 0.5) * d0.nx_real), ipy = f_int((syy + 0.5) * d0.ny_real), ipz = f_int((szz + 0.5) * d0.nz_real);
    ipx = std::min(std::max(ipx, 0), d0.nx - 1); ipy = std::min(std::max(ipy, 0), d0.ny - 1); ipz = std::min(std::max(ipz, 0), d0.nz - 1);
    int idm = ipx + d0.nx * (ipy + d0.ny * ipz);
    Dom& d = w->d[idm];
    int i = ++d.natms;
    d.ensure(i);
    d.parts[i] = CorePart{axx, ayy, azz, 0, 0, 0, w->sites.charge_site[lsite[g - 1]], 0, 0};
    d.vxx[i] = vel ? vel[3 * (g - 1)] : 0.0; d.vyy[i] = vel ? vel[3 * (g - 1) + 1] : 0.0; d.vzz[i] = vel ? vel[3 * (g - 1) + 2] : 0.0;
    d.ltg[i] = g; d.lsite[i] = lsite[g - 1];
    d.ltype[i] = w->sites.type_site[lsite[g - 1]];
    d.lfrzn[i] = w->sites.freeze_site[lsite[g - 1]];
    d.ixyz[i] = 0;
    if (d.lfrzn[i] > 0) w->megfrz++;
  }
  for (Dom& d : w->d) d.nlast = d.natms;
  return 0;
}
// (re)build each domain's list_excl rows for its local atoms from the global table (build_excl.F90 product, by ltg)
static void world_fill_excl(World* w) {
  if (!w->lbook) return;
  for (Dom& d : w->d) {
    d.max_exclude = w->max_exclude;
    d.list_excl.assign((size_t)(d.natms + 64) * (d.max_exclude + 1), 0);
    for (int i = 1; i <= d.natms; ++i)
      for (int k = 0; k <= d.max_exclude; ++k) d.LE(k, i) = w->excl_global[(size_t)(d.ltg[i] - 1) * (w->max_exclude + 1) + k];
  }
}
int ora_world_relocate(void* h) {
  World* w = (World*)h;
  if (w->lbook) for (Dom& d : w->d) if (d.list_excl.empty()) { world_fill_excl(w); break; }
  return relocate_particles(*w);
}
int ora_world_set_halo(void* h) {
  World* w = (World*)h;
  set_halo_particles(*w);
  return 0;
}
int ora_world_refresh_halo(void* h) { return refresh_halo_positions(*(World*)h); }
int ora_world_vnl_check(void* h, double* tol) { return vnl_check(*(World*)h, tol) ? 1 : 0; }
void ora_world_set_strict(void* h, int l_str, int bspline) { ((World*)h)->l_str = l_str != 0; ((World*)h)->bspline = bspline; }
void ora_world_cutoffs(void* h, double* out3) { World* w = (World*)h; out3[0] = w->rcut; out3[1] = w->padding; out3[2] = w->rx; }
// stateless form of the decision: io = {padding, cutoff_extended}, flags = {update, newstart} (in / out), ns5 = neighskip(1:5)
int ora_vnl_decide(int l_str, double tolg, int bspline, double cutoff, double* io2, const double* cell9, const int* dims3, int* flags2,
                   double* ns5, double* width) {
  bool update = flags2[0] != 0, newstart = flags2[1] != 0;
  double ns[6] = {0, ns5[0], ns5[1], ns5[2], ns5[3], ns5[4]};
  double cell[10];
  for (int i = 1; i <= 9; ++i) cell[i] = cell9[i - 1];
  int rc = vnl_decide(l_str != 0, tolg, bspline, cutoff, io2[0], io2[1], cell, dims3[0], dims3[1], dims3[2], update, newstart, ns, *width);
  flags2[0] = update; flags2[1] = newstart;
  for (int i = 0; i < 5; ++i) ns5[i] = ns[i + 1];
  return rc;
}
void ora_world_neighskip(void* h, double* out5) { for (int i = 0; i < 5; ++i) out5[i] = ((World*)h)->neighskip[i + 1]; }

static void par_for_domains(World* w, int nthreads, void (*fn)(World*, int)) {
  int P = (int)w->d.size();
  int nt = std::max(1, std::min(nthreads, P));
#pragma omp parallel for schedule(static, 1) num_threads(nt) if (nt > 1)
  for (int r = 0; r < P; ++r) fn(w, r);
}
// host threads for the world-level loops over domains of every other phase (halo, migration, vnl_check, integrator)
void ora_world_set_threads(void* h, int nthreads) { ((World*)h)->nthreads = std::max(1, nthreads); }
static int g_rc[4096];
int ora_world_link_cell_pairs(void* h, int nthreads) {
  World* w = (World*)h;
  if (w->lbook) {
    bool need = false;
    for (Dom& d : w->d) if ((int)d.list_excl.size() < (d.natms) * (w->max_exclude + 1) || d.max_exclude != w->max_exclude) need = true;
    if (need) world_fill_excl(w);
  }
  par_for_domains(w, nthreads, [](World* ww, int r) { g_rc[r] = link_cell_pairs(*ww, ww->d[r]); });
  for (size_t r = 0; r < w->d.size(); ++r) if (g_rc[r]) return g_rc[r];
  return 0;
}
// zero_forces!=0: parts(:)%f = 0 first (drivers.F90:655-660).  out: per-world sums [engvdw,virvdw,engcpe_rl,vircpe_rl,
// engcpe_ex,vircpe_ex] + stress(9) (the gsum of two_body.F90:729 / drivers.F90:795)
int ora_world_set_collect_pp(void* h, int on) { ((World*)h)->collect_pp = on != 0; return 0; }
// pp_energy(1:natms) and pp_stress(1:9, 1:natms) (column-major, 9 per atom) of domain `rank` after the last two_body call
int ora_dom_get_pp(void* h, int rank, double* pp_energy, double* pp_stress) {
  World* w = (World*)h;
  Dom& d = w->d[rank];
  if ((int)d.pp_energy.size() < d.natms + 1) return 1;
  for (int i = 1; i <= d.natms; ++i) {
    pp_energy[i - 1] = d.pp_energy[i];
    for (int c = 0; c < 9; ++c) pp_stress[(size_t)(i - 1) * 9 + c] = d.pp_stress[(size_t)i * 9 + c];
  }
  return 0;
}
int ora_world_two_body(void* h, int nthreads, int zero_forces, double* out15) {
  World* w = (World*)h;
  if (zero_forces) {
    const int nt0 = std::max(1, std::min(nthreads, (int)w->d.size()));
#pragma omp parallel for schedule(static, 1) num_threads(nt0) if (nt0 > 1)
    for (int rr = 0; rr < (int)w->d.size(); ++rr) {
      Dom& d = w->d[rr];
      for (int i = 1; i <= d.nlast; ++i) { d.parts[i].fxx = 0; d.parts[i].fyy = 0; d.parts[i].fzz = 0; }
    }
  }
  par_for_domains(w, nthreads, [](World* ww, int r) { two_body_forces(*ww, ww->d[r]); });
  if (out15) {
    for (int i = 0; i < 15; ++i) out15[i] = 0.0;
    for (Dom& d : w->d) {
      out15[0] += d.engvdw; out15[1] += d.virvdw; out15[2] += d.engcpe_rl; out15[3] += d.vircpe_rl;
      out15[4] += d.engcpe_ex; out15[5] += d.vircpe_ex;
      for (int k = 1; k <= 9; ++k) out15[5 + k] += d.stress[k];
    }
  }
  return 0;
}
void ora_world_rdf_collect(void* h, const int* rdf_list, int n_pairs, int max_grid, double* rdf /* max_grid*n_pairs, += */) {
  World* w = (World*)h;
  for (Dom& d : w->d) rdf_collect_domain(*w, d, rdf_list, n_pairs, max_grid, rdf);
}
// per-domain accessors
void ora_dom_counts(void* h, int rank, int* out /*natms,nlast,max_list,max_exclude,nlx,nly,nlz,nlp,ncells,nsbcll,ibig*/) {
  Dom& d = ((World*)h)->d[rank];
  int v[11] = {d.natms, d.nlast, d.max_list, d.max_exclude, d.nlx, d.nly, d.nlz, d.nlp, d.ncells, d.nsbcll, d.ibig};
  for (int i = 0; i < 11; ++i) out[i] = v[i];
}
void ora_dom_get_parts(void* h, int rank, void* parts_out /*nlast*64B*/) {
  Dom& d = ((World*)h)->d[rank];
  std::memcpy(parts_out, &d.parts[1], sizeof(CorePart) * d.nlast);
}
void ora_dom_set_parts(void* h, int rank, const void* parts_in, int n) {
  Dom& d = ((World*)h)->d[rank];
  std::memcpy(&d.parts[1], parts_in, sizeof(CorePart) * n);
}
void ora_dom_get_ints(void* h, int rank, int* ltg, int* lsite, int* ltype, int* lfrzn, int* ixyz) {
  Dom& d = ((World*)h)->d[rank];
  for (int i = 1; i <= d.nlast; ++i) {
    if (ltg) ltg[i - 1] = d.ltg[i];
    if (lsite) lsite[i - 1] = d.lsite[i];
    if (ltype) ltype[i - 1] = d.ltype[i];
    if (lfrzn) lfrzn[i - 1] = d.lfrzn[i];
    if (ixyz) ixyz[i - 1] = d.ixyz[i];
  }
}
void ora_dom_get_vel(void* h, int rank, double* v3) {
  Dom& d = ((World*)h)->d[rank];
  for (int i = 1; i <= d.natms; ++i) { v3[3 * (i - 1)] = d.vxx[i]; v3[3 * (i - 1) + 1] = d.vyy[i]; v3[3 * (i - 1) + 2] = d.vzz[i]; }
}
void ora_dom_set_vel(void* h, int rank, const double* v3) {
  Dom& d = ((World*)h)->d[rank];
  for (int i = 1; i <= d.natms; ++i) { d.vxx[i] = v3[3 * (i - 1)]; d.vyy[i] = v3[3 * (i - 1) + 1]; d.vzz[i] = v3[3 * (i - 1) + 2]; }
}
// list in the reference layout (-3:max_list, 1:natms), column-major => row i contiguous, (max_list+4) ints per atom
void ora_dom_get_list(void* h, int rank, int* list_out) {
  Dom& d = ((World*)h)->d[rank];
  std::memcpy(list_out, d.list.data(), sizeof(int) * (size_t)d.natms * (d.max_list + 4));
}
void ora_dom_get_list_excl(void* h, int rank, int* out /*(max_exclude+1)*natms*/) {
  World* w = (World*)h;
  Dom& d = w->d[rank];
  if (w->lbook && ((int)d.list_excl.size() < d.natms * (w->max_exclude + 1) || d.max_exclude != w->max_exclude)) world_fill_excl(w);
  std::memcpy(out, d.list_excl.data(), sizeof(int) * (size_t)d.natms * (d.max_exclude + 1));
}
void ora_dom_get_cells(void* h, int rank, int* which_cell /*nlast*/, int* at_list /*nlast*/, int* lct_start /*ncells+2*/) {
  Dom& d = ((World*)h)->d[rank];
  for (int i = 1; i <= d.nlast; ++i) { which_cell[i - 1] = d.which_cell[i]; at_list[i - 1] = d.at_list[i]; }
  for (int i = 0; i <= d.ncells + 1; ++i) lct_start[i] = d.lct_start[i];
}
void ora_dom_get_results(void* h, int rank, double* out15) {
  Dom& d = ((World*)h)->d[rank];
  out15[0] = d.engvdw; out15[1] = d.virvdw; out15[2] = d.engcpe_rl; out15[3] = d.vircpe_rl; out15[4] = d.engcpe_ex; out15[5] = d.vircpe_ex;
  for (int k = 1; k <= 9; ++k) out15[5 + k] = d.stress[k];
}
void ora_dom_get_bg(void* h, int rank, double* xbg, double* ybg, double* zbg) {
  Dom& d = ((World*)h)->d[rank];
  for (int i = 1; i <= d.nlast; ++i) { xbg[i - 1] = d.xbg[i]; ybg[i - 1] = d.ybg[i]; zbg[i - 1] = d.zbg[i]; }
}
// simple NVE velocity-Verlet stages (nve.F90:163-173, :198-217) so the CPU baseline can advance a trajectory:
// stage 1: v += (dt/2m) f ; x += dt v      stage 2: v += (dt/2m) f       (weight per type)
void ora_world_vv(void* h, int stage, double dt, const double* weight_by_type /*1-based via type-1*/) {
  World& w = *(World*)h;
#pragma omp parallel for schedule(static, 1) num_threads(w.nthreads) if (w.nthreads > 1)
  for (int rr = 0; rr < (int)w.d.size(); ++rr) {
    Dom& d = w.d[rr];
    for (int i = 1; i <= d.natms; ++i) {
      double hstep = 0.5 * dt, rm = 1.0 / weight_by_type[d.ltype[i] - 1];
      CorePart& p = d.parts[i];
      double tmp = hstep * rm;
      d.vxx[i] = d.vxx[i] + tmp * p.fxx; d.vyy[i] = d.vyy[i] + tmp * p.fyy; d.vzz[i] = d.vzz[i] + tmp * p.fzz;
      if (stage == 1) { p.xxx = p.xxx + dt * d.vxx[i]; p.yyy = p.yyy + dt * d.vyy[i]; p.zzz = p.zzz + dt * d.vzz[i]; }
    }
  }
}

// ---- independent second opinion: O(N^2) minimum-image brute force ---------------------------------------------
// Pair set by definition |r_ij|^2 <= rx^2 (min image, orthorhombic/cubic), returned as sorted (gi<gj) pairs.
// Returns number of pairs; pairs_out may be null to query size.  band_out counts pairs within rel 1e-12 of rx^2.
long ora_brute_pairs(int n, const double* xyz, const double* cell9, double rx, int* pairs_out, long cap, long* band_out) {
  long np = 0, band = 0;
  long double c[10], rc[10];   // general cell: minimum image in reduced coordinates (see ora_world_brute_forces)
  for (int k = 1; k <= 9; ++k) c[k] = cell9[k - 1];
  rc[1] = c[5] * c[9] - c[6] * c[8]; rc[2] = c[3] * c[8] - c[2] * c[9]; rc[3] = c[2] * c[6] - c[3] * c[5];
  rc[4] = c[6] * c[7] - c[4] * c[9]; rc[5] = c[1] * c[9] - c[3] * c[7]; rc[6] = c[3] * c[4] - c[1] * c[6];
  rc[7] = c[4] * c[8] - c[5] * c[7]; rc[8] = c[2] * c[7] - c[1] * c[8]; rc[9] = c[1] * c[5] - c[2] * c[4];
  {
    long double det = c[1] * rc[1] + c[4] * rc[2] + c[7] * rc[3];
    for (int k = 1; k <= 9; ++k) rc[k] /= det;
  }
  long double rc2 = (long double)rx * rx;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      long double dx = (long double)xyz[3 * i] - xyz[3 * j], dy = (long double)xyz[3 * i + 1] - xyz[3 * j + 1],
                  dz = (long double)xyz[3 * i + 2] - xyz[3 * j + 2];
      {
        long double sx = rc[1] * dx + rc[4] * dy + rc[7] * dz, sy = rc[2] * dx + rc[5] * dy + rc[8] * dz,
                    sz = rc[3] * dx + rc[6] * dy + rc[9] * dz;
        long double nx = roundl(sx), ny = roundl(sy), nz = roundl(sz);
        dx -= c[1] * nx + c[4] * ny + c[7] * nz; dy -= c[2] * nx + c[5] * ny + c[8] * nz; dz -= c[3] * nx + c[6] * ny + c[9] * nz;
      }
      long double r2 = dx * dx + dy * dy + dz * dz;
      if (fabsl(r2 - rc2) <= 1e-12L * rc2) band++;
      if (r2 <= rc2) {
        if (pairs_out && np < cap) { pairs_out[2 * np] = i + 1; pairs_out[2 * np + 1] = j + 1; }
        np++;
      }
    }
  if (band_out) *band_out = band;
  return np;
}
// Brute-force forces/energies with the world's tables (tabulated or direct vdW + Ewald real + exclusion correction),
// every unique pair once, minimum image, long double accumulation.  excl via world excl_global.
void ora_world_brute_forces(void* h, int n, const double* xyz, const int* lsite, double* f_out /*3n*/, double* out6) {
  World* w = (World*)h;
  Vdw& v = w->vdw; Ewald& e = w->ew;
  // minimum image in reduced coordinates of the general (parallelepiped) cell, in long double: exact for every pair closer
  // than half the smallest perpendicular width, which cutoff_extended always is (neighbours.F90:409-412)
  long double c[10], rc[10];
  for (int k = 1; k <= 9; ++k) c[k] = w->cell[k];
  {
    rc[1] = c[5] * c[9] - c[6] * c[8]; rc[2] = c[3] * c[8] - c[2] * c[9]; rc[3] = c[2] * c[6] - c[3] * c[5];
    rc[4] = c[6] * c[7] - c[4] * c[9]; rc[5] = c[1] * c[9] - c[3] * c[7]; rc[6] = c[3] * c[4] - c[1] * c[6];
    rc[7] = c[4] * c[8] - c[5] * c[7]; rc[8] = c[2] * c[7] - c[1] * c[8]; rc[9] = c[1] * c[5] - c[2] * c[4];
    long double det = c[1] * rc[1] + c[4] * rc[2] + c[7] * rc[3];
    for (int k = 1; k <= 9; ++k) rc[k] /= det;
  }
  std::vector<long double> F(3 * (size_t)n, 0.0L);
  long double ev = 0, vv = 0, ec = 0, vc = 0, ex = 0, vx = 0;
  for (int i = 0; i < n; ++i)
    for (int j = i + 1; j < n; ++j) {
      long double dxl = (long double)xyz[3 * i] - xyz[3 * j], dyl = (long double)xyz[3 * i + 1] - xyz[3 * j + 1],
                  dzl = (long double)xyz[3 * i + 2] - xyz[3 * j + 2];
      {
        long double sx = rc[1] * dxl + rc[4] * dyl + rc[7] * dzl, sy = rc[2] * dxl + rc[5] * dyl + rc[8] * dzl,
                    sz = rc[3] * dxl + rc[6] * dyl + rc[9] * dzl;
        long double nx = roundl(sx), ny = roundl(sy), nz = roundl(sz);
        dxl -= c[1] * nx + c[4] * ny + c[7] * nz; dyl -= c[2] * nx + c[5] * ny + c[8] * nz; dzl -= c[3] * nx + c[6] * ny + c[9] * nz;
      }
      long double r2 = dxl * dxl + dyl * dyl + dzl * dzl;
      if (r2 > (long double)w->rx * w->rx) continue;
      double dx = (double)dxl, dy = (double)dyl, dz = (double)dzl;
      double rrr = std::sqrt((double)r2);
      int gi = i + 1, gj = j + 1;
      bool excluded = false;
      if (w->lbook) {
        const int* row = &w->excl_global[(size_t)(gi - 1) * (w->max_exclude + 1)];
        excluded = match(gj, row[0], row);
      }
      int ai = w->sites.type_site[lsite[i]], aj = w->sites.type_site[lsite[j]];
      double qi = w->sites.charge_site[lsite[i]], qj = w->sites.charge_site[lsite[j]];
      bool frozen = (w->megfrz > 1) && w->sites.freeze_site[lsite[i]] > 0 && w->sites.freeze_site[lsite[j]] > 0;
      if (frozen) continue;
      long double g_tot = 0.0L;
      if (!excluded) {
        if (!v.no_vdw) {
          int key = (ai > aj) ? ai * (ai - 1) / 2 + aj : aj * (aj - 1) / 2 + ai;
          int k = v.list[key];
          bool defined = v.l_direct ? true : !(std::fabs(v.tp(0, k)) < zero_plus);
          if (defined && v.ltp[k] != -1 && rrr < v.cutoff) {
            double gamma, eng;
            if (v.l_direct) {
              EG eg = pot_energy(v.ltp[k], v.par(k), rrr);
              eng = eg.energy + v.afs[k] * rrr + v.bfs[k];
              gamma = eg.gamma / (rrr * rrr) - v.afs[k] / rrr;
            } else {
              int l = f_int(rrr * v.rdr);
              double ppp = rrr * v.rdr - (double)l;
              double gk = v.tf(l, k); if (l == 0) gk *= rrr;
              double gk1 = v.tf(l + 1, k), gk2 = v.tf(l + 2, k);
              double t1 = gk + (gk1 - gk) * ppp, t2 = gk1 + (gk2 - gk1) * (ppp - 1.0);
              gamma = (t1 + (t2 - t1) * ppp * 0.5) / (rrr * rrr);
              double vk = v.tp(l, k), vk1 = v.tp(l + 1, k), vk2 = v.tp(l + 2, k);
              t1 = vk + (vk1 - vk) * ppp; t2 = vk1 + (vk2 - vk1) * (ppp - 1.0);
              eng = t1 + (t2 - t1) * ppp * 0.5;
              if (v.l_force_shift) {
                gamma -= v.tf(v.max_grid - 4, k) / (rrr * v.cutoff);
                eng += v.tf(v.max_grid - 4, k) * (rrr / v.cutoff - 1.0) - v.tp(v.max_grid - 4, k);
              }
            }
            ev += eng; vv -= (long double)gamma * rrr * rrr; g_tot += gamma;
          }
        }
        if (e.active && std::fabs(qi * e.scaling) >= zero_plus && std::fabs(qj) > zero_plus && rrr < w->rcut) {
          double prefac = qi * e.scaling * qj;
          int nsi = f_int(rrr * e.recip_spacing);
          double diff = rrr * e.recip_spacing - (double)nsi;
          const double* td = e.erfc_deriv.data(); const double* te = e.erfc.data();
          double p1 = td[nsi], p2 = td[nsi + 1], p3 = td[nsi + 2]; if (nsi == 0) p1 *= rrr;
          double tm1 = p1 + (p2 - p1) * diff, tm2 = p2 + (p3 - p2) * (diff - 1.0);
          double gam = prefac * (tm1 + (tm2 - tm1) * diff * 0.5);
          p1 = te[nsi]; p2 = te[nsi + 1]; p3 = te[nsi + 2]; if (nsi == 0) p1 *= rrr;
          tm1 = p1 + (p2 - p1) * diff; tm2 = p2 + (p3 - p2) * (diff - 1.0);
          double en = prefac * (tm1 + (tm2 - tm1) * diff * 0.5);
          ec += en; vc -= (long double)gam * rrr * rrr; g_tot += gam;
        }
      } else if (e.active && std::fabs(qi) > zero_plus && std::fabs(qj) > zero_plus && rrr < w->rcut) {
        double chgprd = qj * (qi * e.scaling);
        double alpr = rrr * e.alpha;
        double exp1 = std::exp(-(alpr * alpr));
        double erfr = chgprd * (1.0 - calc_erfc(alpr)) / rrr;   // same A&S polynomial as the reference
        double egamma = -(erfr - 2.0 * chgprd * (e.alpha / sqrpi) * exp1) / (rrr * rrr);
        ex -= erfr; vx -= (long double)egamma * rrr * rrr; g_tot += egamma;
      }
      F[3 * i] += g_tot * dx; F[3 * i + 1] += g_tot * dy; F[3 * i + 2] += g_tot * dz;
      F[3 * j] -= g_tot * dx; F[3 * j + 1] -= g_tot * dy; F[3 * j + 2] -= g_tot * dz;
    }
  for (size_t k = 0; k < F.size(); ++k) f_out[k] = (double)F[k];
  out6[0] = (double)ev; out6[1] = (double)vv; out6[2] = (double)ec; out6[3] = (double)vc; out6[4] = (double)ex; out6[5] = (double)vx;
}

}  // extern "C"
