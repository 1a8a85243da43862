// C++ host side of the short-range two-body path (see dlpoly_host.hpp).  Compiled with -ffp-contract=off: the table
// generators and the decisions below follow the reference's un-fused IEEE arithmetic.
#include "dlpoly_host.hpp"

#include <algorithm>
#include <cstring>
#include <fstream>
#include <limits>
#include <sstream>

namespace dlpoly {

void error(int kode, const std::string& message) { throw dlpoly_error(kode, message); }

namespace {

inline long f_nint(double x) { return x >= 0.0 ? (long)std::floor(x + 0.5) : -(long)std::floor(-x + 0.5); }   // Fortran Nint
inline int f_int(double x) { return (int)x; }                                                                // Fortran Int

// real ** integer as gfortran evaluates it (libgcc __powidf2: square-and-multiply from the low bit)
double powi(double x, int n) {
  unsigned m = n < 0 ? 0u - (unsigned)n : (unsigned)n;
  double y = (m & 1u) ? x : 1.0;
  while (m >>= 1) {
    x = x * x;
    if (m & 1u) y = y * x;
  }
  return n < 0 ? 1.0 / y : y;
}

// parse.F90 word_2_real: Fortran real literals may carry a D exponent
double word_2_real(std::string w) {
  for (char& c : w)
    if (c == 'd' || c == 'D') c = 'e';
  char* end = nullptr;
  const double v = std::strtod(w.c_str(), &end);
  return end == w.c_str() ? 0.0 : v;
}

}   // namespace

// ---------------------------------------------------------------- numerics.F90:1344-1446
void dcell(const double a[9], double b[10]) {
  b[0] = std::sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
  b[1] = std::sqrt(a[3] * a[3] + a[4] * a[4] + a[5] * a[5]);
  b[2] = std::sqrt(a[6] * a[6] + a[7] * a[7] + a[8] * a[8]);
  b[3] = (a[0] * a[3] + a[1] * a[4] + a[2] * a[5]) / (b[0] * b[1]);
  b[4] = (a[0] * a[6] + a[1] * a[7] + a[2] * a[8]) / (b[0] * b[2]);
  b[5] = (a[3] * a[6] + a[4] * a[7] + a[5] * a[8]) / (b[1] * b[2]);
  const double axb[3] = {a[1] * a[5] - a[2] * a[4], a[2] * a[3] - a[0] * a[5], a[0] * a[4] - a[1] * a[3]};
  const double bxc[3] = {a[4] * a[8] - a[5] * a[7], a[5] * a[6] - a[3] * a[8], a[3] * a[7] - a[4] * a[6]};
  const double cxa[3] = {a[7] * a[2] - a[8] * a[1], a[8] * a[0] - a[6] * a[2], a[6] * a[1] - a[7] * a[0]};
  b[9] = std::fabs(a[0] * bxc[0] + a[1] * bxc[1] + a[2] * bxc[2]);
  const double d[3] = {b[9] / std::sqrt(bxc[0] * bxc[0] + bxc[1] * bxc[1] + bxc[2] * bxc[2]),
                       b[9] / std::sqrt(cxa[0] * cxa[0] + cxa[1] * cxa[1] + cxa[2] * cxa[2]),
                       b[9] / std::sqrt(axb[0] * axb[0] + axb[1] * axb[1] + axb[2] * axb[2])};
  // the width that belongs to x is the one of the lattice vector most aligned with x, then y among the other two
  double x[3], y[3];
  for (int v = 0; v < 3; ++v) {
    x[v] = std::fabs(a[3 * v]) / b[v];
    y[v] = std::fabs(a[3 * v + 1]) / b[v];
  }
  int first;
  if (x[0] >= x[1] && x[0] >= x[2]) first = 0;
  else if (x[1] >= x[0] && x[1] >= x[2]) first = 1;
  else first = 2;
  const int p = first == 0 ? 1 : 0, q = first == 2 ? 1 : 2;   // the two remaining vectors, ascending
  b[6] = d[first];
  if (y[p] >= y[q]) { b[7] = d[p]; b[8] = d[q]; }
  else { b[7] = d[q]; b[8] = d[p]; }
}

// ---------------------------------------------------------------- numerics.F90:1448-1509 (adjugate scaled by 1 / det)
void invert(const double a[9], double b[9], double& d) {
  b[0] = a[4] * a[8] - a[5] * a[7];
  b[1] = a[2] * a[7] - a[1] * a[8];
  b[2] = a[1] * a[5] - a[2] * a[4];
  b[3] = a[5] * a[6] - a[3] * a[8];
  b[4] = a[0] * a[8] - a[2] * a[6];
  b[5] = a[2] * a[3] - a[0] * a[5];
  b[6] = a[3] * a[7] - a[4] * a[6];
  b[7] = a[1] * a[6] - a[0] * a[7];
  b[8] = a[0] * a[4] - a[1] * a[3];
  d = a[0] * b[0] + a[3] * b[1] + a[6] * b[2];
  const double r = std::fabs(d) > 0.0 ? 1.0 / d : 0.0;
  for (int k = 0; k < 9; ++k) b[k] = r * b[k];
}

// ---------------------------------------------------------------- domains.F90:63-258
void map_domains(int imcon, double wx, double wy, double wz, int idnode, int mxnode, domains_type& dom) {
  dom.mxnode = mxnode;
  dom.idnode = idnode;
  if (mxnode == 1) {
    dom.nx = dom.ny = dom.nz = 1;
  } else {
    const double tol = 1.0e-6;
    const int huge = std::numeric_limits<int>::max();
    const int limx = imcon != 0 ? huge : 2, limy = limx, limz = (imcon != 0 && imcon != 6) ? huge : 2;
    double min_s = std::numeric_limits<double>::max();
    int bx = -1, by = -1, bz = -1;
    for (int nx = 1; nx <= mxnode; ++nx) {
      if (mxnode % nx != 0 || nx > limx) continue;
      const double dx = wx / (double)nx;
      const int pyz = mxnode / nx;
      for (int ny = 1; ny <= pyz; ++ny) {
        if (pyz % ny != 0 || ny > limy) continue;
        const int nz = pyz / ny;
        if (nz > limz) continue;
        const double dy = wy / (double)ny, dz = wz / (double)nz;
        const double s = 2.0 * (dx * dy + dy * dz + dz * dx);
        bool take = false;
        if (min_s - s > tol) {
          take = true;
        } else if (std::fabs(min_s - s) < tol) {   // degenerate: fewest ranks along any axis, then least along x, then y
          const int mnew = std::max(nx, std::max(ny, nz)), mold = std::max(bx, std::max(by, bz));
          take = mnew < mold || (mnew == mold && (nx < bx || (nx == bx && ny < by)));
        }
        if (take) { min_s = s; bx = nx; by = ny; bz = nz; }
      }
    }
    if (bx == -1 || by == -1 || bz == -1) error(520, "no domain decomposition found");
    dom.nx = bx; dom.ny = by; dom.nz = bz;
  }
  dom.nx_recip = 1.0 / (double)dom.nx;
  dom.ny_recip = 1.0 / (double)dom.ny;
  dom.nz_recip = 1.0 / (double)dom.nz;
  dom.idz = idnode / (dom.nx * dom.ny);
  dom.idy = idnode / dom.nx - dom.idz * dom.ny;
  dom.idx = idnode % dom.nx;
  // map(1:26): faces -x,+x,-y,+y,-z,+z, then the xy, xz, yz edges, then the corners, in the reference's order
  static const signed char off[26][3] = {
      {-1, 0, 0}, {1, 0, 0}, {0, -1, 0}, {0, 1, 0}, {0, 0, -1}, {0, 0, 1},
      {-1, 1, 0}, {1, -1, 0}, {-1, -1, 0}, {1, 1, 0},
      {-1, 0, 1}, {1, 0, -1}, {-1, 0, -1}, {1, 0, 1},
      {0, -1, 1}, {0, 1, -1}, {0, -1, -1}, {0, 1, 1},
      {-1, -1, -1}, {1, 1, 1}, {-1, -1, 1}, {1, 1, -1}, {-1, 1, -1}, {1, -1, 1}, {-1, 1, 1}, {1, -1, -1}};
  for (int k = 0; k < 26; ++k) {
    const int i = (dom.idx + off[k][0] + dom.nx) % dom.nx, j = (dom.idy + off[k][1] + dom.ny) % dom.ny,
              l = (dom.idz + off[k][2] + dom.nz) % dom.nz;
    dom.map[k] = i + dom.nx * (j + dom.ny * l);   // idcube
    dom.map_unique[k] = 0;
  }
  for (int i = 0; i < 26; ++i) {
    if (idnode == dom.map[i]) dom.map_unique[i] = 1;
    for (int j = i + 1; j < 26; ++j)
      if (dom.map[i] == dom.map[j]) dom.map_unique[j] = 1;
  }
}

// ---------------------------------------------------------------- sizes
int max_grid_of(double rcut) { return (int)std::max(1004L, f_nint(rcut / delr_max) + 4); }
int max_list_of(double fdens, double rx) { return (int)f_nint(fdens * (7.5 / 3.0) * pi * powi(rx, 3)); }

// ---------------------------------------------------------------- vdw_type
void vdw_type::init(int ntype, double rvdw, bool force_shift, bool direct) {
  ntype_atom = ntype;
  cutoff = rvdw;
  l_force_shift = force_shift;
  l_direct = direct;
  max_grid = max_grid_of(rvdw);
  list.assign((size_t)ntype * (ntype + 1) / 2, 0);
  n_vdw = 0;
  ltp.clear();
  param.clear();
}

int vdw_type::add(int ai, int aj, int keypot, const double* prm, int nprm) {
  const int k = key(ai, aj);
  if (k < 1 || k > (int)list.size()) error(81, "vdw pair refers to an unknown atom type");
  if (list[k - 1] != 0) error(15, "duplicate vdw potential for one pair");   // ffield.F90:3868
  ++n_vdw;
  list[k - 1] = n_vdw;
  ltp.push_back(keypot);
  for (int p = 0; p < 7; ++p) param.push_back(p < nprm ? prm[p] : 0.0);
  if (keypot == VDW_TAB) l_tab = true;
  return n_vdw;
}

void vdw_type::finalise() {
  const int ntab = (int)list.size();
  max_vdw = n_vdw < ntab ? n_vdw + 1 : std::max(n_vdw, 1);
  for (int& v : list)
    if (v == 0) v = n_vdw + 1;
  ltp.resize(max_vdw, VDW_NULL);
  param.resize((size_t)7 * max_vdw, 0.0);
  afs.assign(max_vdw, 0.0);
  bfs.assign(max_vdw, 0.0);
  tab_potential.assign((size_t)(max_grid + 1) * max_vdw, 0.0);
  tab_force.assign((size_t)(max_grid + 1) * max_vdw, 0.0);
}

// two_body_potentials.F90:260-270 (lj), :307-317 (12-6), :471-485 (buck), :499-514 (bhm): energy and gamma = -r dU/dr
void pair_potential(int keypot, const double* p, double r, double& e, double& g) {
  switch (keypot) {
    case VDW_12_6: {
      const double r6 = powi(1.0 / r, 6);
      e = (p[0] * r6 - p[1]) * r6;
      g = 6.0 * r6 * (2.0 * p[0] * r6 - p[1]);
      return;
    }
    case VDW_LENNARD_JONES: {
      const double s6 = powi(p[1] / r, 6);
      e = 4.0 * p[0] * s6 * (s6 - 1.0);
      g = 24.0 * p[0] * s6 * (2.0 * s6 - 1.0);
      return;
    }
    case VDW_BUCKINGHAM: {
      const double b = r / p[1];
      const double t1 = p[0] * std::exp(-b), t2 = -p[2] / powi(r, 6);
      e = t1 + t2;
      g = t1 * b + 6.0 * t2;
      return;
    }
    case VDW_BORN_HUGGINS_MEYER: {
      const double ri2 = powi(r, -2);
      const double t1 = p[0] * std::exp(p[1] * (p[2] - r)), t2 = -p[3] * powi(ri2, 3), t3 = -p[4] * powi(ri2, 4);
      e = t1 + t2 + t3;
      g = t1 * r * p[1] + 6.0 * t2 + 8.0 * t3;
      return;
    }
    default:
      error(150, "pair potential key " + std::to_string(keypot) + " is not evaluated by the short-range GPU path");
  }
}

// vdw.F90:1397-1576
void vdw_generate(vdw_type& v) {
  const double dlrpot = v.cutoff / (double)(v.max_grid - 4);
  const double huge = std::numeric_limits<double>::max();
  for (int k = 0; k < v.n_vdw; ++k) {
    if (v.ltp[k] == VDW_TAB || v.ltp[k] == VDW_NULL) continue;
    double* tp = &v.tab_potential[(size_t)k * (v.max_grid + 1)];
    double* tf = &v.tab_force[(size_t)k * (v.max_grid + 1)];
    for (int i = 1; i <= v.max_grid; ++i) pair_potential(v.ltp[k], &v.param[(size_t)7 * k], (double)i * dlrpot, tp[i], tf[i]);
    tp[0] = huge;
    tf[0] = huge;
  }
}

// vdw.F90:969-1049
void vdw_direct_fs_generate(vdw_type& v) {
  for (int k = 0; k < v.n_vdw; ++k) {
    if (v.ltp[k] == VDW_TAB || v.ltp[k] == VDW_NULL) continue;
    double z, dz;
    pair_potential(v.ltp[k], &v.param[(size_t)7 * k], v.cutoff, z, dz);
    v.afs[k] = dz / v.cutoff;
    v.bfs[k] = -z - dz;
  }
}

// vdw.F90:1051-1370
void vdw_table_read(vdw_type& v, const std::string& filename, double engunit) {
  std::ifstream f(filename);
  std::string record, word;
  if (!f || !std::getline(f, record)) error(24, "end of file in TABLE " + filename);   // header record
  if (!std::getline(f, record)) error(24, "end of file in TABLE");
  std::istringstream rs(record);
  std::string w1, w2, w3;
  rs >> w1 >> w2 >> w3;
  double delpot = word_2_real(w1);
  const double cutpot = word_2_real(w2);
  const int ngrid = (int)f_nint(word_2_real(w3));
  const double dlrpot = v.cutoff / (double)(v.max_grid - 4);
  bool safe = false;
  if (std::fabs(delpot - dlrpot) <= 1.0e-8) { safe = true; delpot = dlrpot; }
  if (delpot > delr_max && !safe) error(22, "TABLE radial increment exceeds delr_max");
  const bool remake = std::fabs(1.0 - (delpot / dlrpot)) > 1.0e-8;
  const double rdr = 1.0 / delpot;
  if (ngrid < v.max_grid - 4) error(0, "Transfer buffer too small in vdw_table_read");
  if (cutpot < v.cutoff) error(0, "Cutoff too large for TABLE file");
  std::vector<double> buffer((size_t)ngrid + 3, 0.0);   // buffer(0:ngrid)
  const int g = v.max_grid;

  auto read_array = [&]() {
    for (int i = 1; i <= ngrid; ++i) {
      if (!(f >> word)) error(24, "end of file in TABLE");
      buffer[i] = word_2_real(word);
    }
    f.ignore(std::numeric_limits<std::streamsize>::max(), '\n');
  };
  auto regrid = [&](double* tab) {   // tab(1 : max_grid-2) from buffer; tab(0) is the caller's
    if (remake) {
      for (int i = 1; i <= g - 4; ++i) {
        const double rrr = (double)i * dlrpot;
        const int l = f_int(rrr * rdr);
        const double ppp = rrr * rdr - (double)l;
        const double vk = buffer[l];
        double vk1, vk2;
        if (l + 2 > ngrid) {        // linear extrapolation just beyond the file's last point
          if (l + 1 > ngrid) { vk1 = 2.0 * buffer[l] - buffer[l - 1]; vk2 = 2.0 * vk1 - buffer[l]; }
          else { vk1 = buffer[l + 1]; vk2 = 2.0 * buffer[l + 1] - buffer[l]; }
        } else { vk1 = buffer[l + 1]; vk2 = buffer[l + 2]; }
        const double t1 = vk + (vk1 - vk) * ppp, t2 = vk1 + (vk2 - vk1) * (ppp - 1.0);
        tab[i] = t1 + (t2 - t1) * ppp * 0.5;
      }
    } else {
      for (int i = 1; i <= g - 4; ++i) tab[i] = buffer[i];
      tab[g - 3] = 2.0 * tab[g - 4] - tab[g - 5];
    }
    tab[g - 2] = 2.0 * tab[g - 3] - tab[g - 4];
  };

  for (int ivdw = 1; ivdw <= v.n_vdw; ++ivdw) {
    if (v.ltp[ivdw - 1] != VDW_TAB) continue;
    if (!std::getline(f, record)) error(24, "end of file in TABLE");
    std::istringstream ls(record);
    std::string atom1, atom2, we, wv;
    ls >> atom1 >> atom2 >> we >> wv;
    v.param[(size_t)7 * (ivdw - 1) + 0] = word_2_real(we) * engunit;   // elrc, vlrc of the pair
    v.param[(size_t)7 * (ivdw - 1) + 1] = word_2_real(wv) * engunit;
    int katom1 = 0, katom2 = 0;
    for (int j = 0; j < (int)v.unique_atom.size(); ++j) {
      if (atom1 == v.unique_atom[j]) katom1 = j + 1;
      if (atom2 == v.unique_atom[j]) katom2 = j + 1;
    }
    if (katom1 == 0 || katom2 == 0) error(81, "****" + atom1 + "***" + atom2 + "**** entry in TABLE");
    if (v.list[vdw_type::key(katom1, katom2) - 1] != ivdw) error(0, "Incompatible FIELD and TABLE file potentials");
    double* tp = &v.tab_potential[(size_t)(ivdw - 1) * (g + 1)];
    double* tf = &v.tab_force[(size_t)(ivdw - 1) * (g + 1)];
    read_array();
    tp[0] = 2.0 * buffer[1] - buffer[2];
    regrid(tp);
    read_array();
    tf[0] = (2.0 * buffer[1] - 0.5 * buffer[2]) / delpot;
    regrid(tf);
    if (std::fabs(tp[0]) <= zero_plus) tp[0] = std::copysign(zero_plus, tp[0]);   // "something has been defined"
  }
  for (int ivdw = 1; ivdw <= v.n_vdw; ++ivdw) {
    if (v.ltp[ivdw - 1] != VDW_TAB) continue;
    double* tp = &v.tab_potential[(size_t)(ivdw - 1) * (g + 1)];
    double* tf = &v.tab_force[(size_t)(ivdw - 1) * (g + 1)];
    for (int i = 0; i <= g; ++i) { tp[i] *= engunit; tf[i] *= engunit; }
    if (v.l_force_shift) tp[g - 3] = tp[g - 2] = tf[g - 3] = tf[g - 2] = 0.0;
  }
}

// vdw.F90:617-967
void vdw_lrc(const std::vector<double>& num_type, const std::vector<double>& numfrz, vdw_type& v, int imcon, double volm) {
  v.elrc = v.vlrc = 0.0;
  if (v.l_force_shift || imcon == 0 || imcon == 6) return;   // shifted potentials vanish at the cutoff; 3D periodic systems only
  const double r = v.cutoff, r3 = powi(r, 3), r5 = powi(r, 5), r9 = powi(r, 9);
  double plrc = 0.0;
  int ivdw = 0;
  for (int i = 1; i <= v.ntype_atom; ++i) {
    for (int j = 1; j <= i; ++j) {
      const int k = v.list[ivdw++] - 1;
      const double* p = &v.param[(size_t)7 * k];
      double eadd = 0.0, padd = 0.0;   // integrals of u r^2 and of (r du/dr) r^2 beyond the cutoff
      switch (v.ltp[k]) {
        case VDW_TAB: eadd = p[0]; padd = -p[1]; break;                        // the TABLE file's own corrections
        case VDW_12_6:
          eadd = p[0] / (9.0 * r9) - p[1] / (3.0 * r3);
          padd = 12.0 * p[0] / (9.0 * r9) - 6.0 * p[1] / (3.0 * r3);
          break;
        case VDW_LENNARD_JONES:
          eadd = 4.0 * p[0] * (powi(p[1], 12) / (9.0 * r9) - powi(p[1], 6) / (3.0 * r3));
          padd = 8.0 * p[0] * (6.0 * powi(p[1], 12) / (9.0 * r9) - powi(p[1], 6) / r3);
          break;
        case VDW_BUCKINGHAM:
          eadd = -p[2] / (3.0 * r3);
          padd = -2.0 * p[2] / r3;
          break;
        case VDW_BORN_HUGGINS_MEYER:
          eadd = -p[3] / (3.0 * r3) - p[4] / (5.0 * r5);
          padd = -2.0 * p[3] / r3 - 8.0 * p[4] / (5.0 * r5);
          break;
        default: break;                                                        // VDW_NULL: no potential for this pair
      }
      if (i != j) { eadd = eadd * 2.0; padd = padd * 2.0; }                    // unlike pairs count twice
      const double denprd = 2.0 * pi * (num_type[i - 1] * num_type[j - 1] - numfrz[i - 1] * numfrz[j - 1]) / powi(volm, 2);
      v.elrc = v.elrc + volm * denprd * eadd;
      plrc = plrc + denprd * padd / 3.0;
    }
  }
  v.vlrc = plrc * (-3.0 * volm);
}

// configuration.F90:1183-1205
void read_config_fold(std::vector<double>& xyz, const double cell[9], const domains_type& dom, std::vector<int>& owner) {
  double rc[9], det;
  invert(cell, rc, det);
  const size_t n = xyz.size() / 3;
  owner.assign(n, 0);
  const double hm = half_minus();
  auto fold = [hm](double s) {
    // Anint: round half away from zero
    const double t = std::trunc(s);
    const double a = std::fabs(s - t) >= 0.5 ? t + (s < 0.0 ? -1.0 : 1.0) : t;
    s = s - a;
    return s >= hm ? -s : s;
  };
  auto ip = [](double s, int nd) { return std::min(std::max((int)((s + 0.5) * (double)nd), 0), nd - 1); };
  for (size_t i = 0; i < n; ++i) {
    const double ax = xyz[3 * i], ay = xyz[3 * i + 1], az = xyz[3 * i + 2];
    const double sx = fold(rc[0] * ax + rc[3] * ay + rc[6] * az);
    const double sy = fold(rc[1] * ax + rc[4] * ay + rc[7] * az);
    const double sz = fold(rc[2] * ax + rc[5] * ay + rc[8] * az);
    xyz[3 * i] = cell[0] * sx + cell[3] * sy + cell[6] * sz;
    xyz[3 * i + 1] = cell[1] * sx + cell[4] * sy + cell[7] * sz;
    xyz[3 * i + 2] = cell[2] * sx + cell[5] * sy + cell[8] * sz;
    owner[i] = ip(sx, dom.nx) + dom.nx * (ip(sy, dom.ny) + dom.ny * ip(sz, dom.nz));
  }
}

// two_body.F90:672-790
void two_body_totals(stats_type& st, const vdw_type& v, const electrostatic_type& el, const ewald_type& ew, const configuration_type& c,
                     int mxnode, double engvdw, double virvdw, double engcpe_rc, double vircpe_rc, double engcpe_rl, double vircpe_rl,
                     double engcpe_ex, double vircpe_ex) {
  double engcpe_nz = 0.0, vircpe_nz = 0.0;
  if (el.key == ELECTROSTATIC_SPME && std::fabs(c.sumchg) > 1.0e-6) {          // Fuchs, Proc. R. Soc. A 151 (1935) 585
    const double factor_nz = -0.5 * (pi * r4pie0 / el.eps) * powi(c.sumchg / ew.alpha, 2);
    engcpe_nz = factor_nz / c.volm;
    vircpe_nz = -3.0 * engcpe_nz;
  }
  const double zero = 0.0;   // core-shell (ch), frozen (fr) and multipole (dt) terms are not part of this path
  st.engcpe = st.engcpe + engcpe_rc + engcpe_rl + zero + engcpe_ex + zero + engcpe_nz;
  st.vircpe = st.vircpe + vircpe_rc + vircpe_rl + zero + vircpe_ex + zero + vircpe_nz + zero;
  st.engsrp = st.engsrp + (engvdw + v.elrc);
  st.virsrp = st.virsrp + (virvdw + v.vlrc);
  for (const double corr : {-vircpe_nz / (3.0 * (double)mxnode), -(v.vlrc + 0.0) / (3.0 * (double)mxnode)}) {
    st.stress[0] = st.stress[0] + corr;
    st.stress[4] = st.stress[4] + corr;
    st.stress[8] = st.stress[8] + corr;
  }
}

// ---------------------------------------------------------------- electrostatics
double calc_erfc(double x) {   // numerics.F90:3659-3665 (Abramowitz-Stegun 7.1.26)
  const double a1 = 0.254829592, a2 = -0.284496736, a3 = 1.421413741, a4 = -1.453152027, a5 = 1.061405429, pp = 0.3275911;
  const double tt = 1.0 / (1.0 + pp * x);
  return tt * (a1 + tt * (a2 + tt * (a3 + tt * (a4 + tt * a5)))) * std::exp(-(x * x));
}

double ewald_alpha_from_precision(double precision, double rcut) {   // control.F90:1709-1710
  const double tol = std::sqrt(std::fabs(std::log(precision * rcut)));
  return std::sqrt(std::fabs(std::log(precision * rcut * tol))) / rcut;
}

void electrostatic_type::erfcgen(double rcut, double alpha) {
  const int n = max_grid_of(rcut);
  const double rsqrpi = 1.0 / std::sqrt(pi);
  for (interp_table* t : {&erfc, &erfc_deriv}) {
    t->nsamples = n;
    t->table.assign((size_t)n + 1, 0.0);
    t->spacing = rcut / (double)(n - 4);
    t->recip_spacing = 1.0 / t->spacing;
  }
  for (int i = 1; i <= n; ++i) {
    const double r = (double)i * erfc.spacing;
    const double e = calc_erfc(alpha * r) / r;
    const double ar = alpha * r;
    erfc.table[i] = e;
    erfc_deriv.table[i] = (e + alpha * (2.0 * std::exp(-(ar * ar)) * rsqrpi)) / (r * r);
  }
  erfc.end_sample = erfc.table[n - 4];
  erfc_deriv.end_sample = erfc_deriv.table[n - 4];
}

void coul_setup(electrostatic_type& el, double rcut) {
  el.force_shift = el.energy_shift = 0.0;
  el.reaction_field[0] = el.reaction_field[1] = el.reaction_field[2] = 0.0;
  el.damp = el.damping > 0.0 && (el.key == ELECTROSTATIC_COULOMB_FORCE_SHIFT || el.key == ELECTROSTATIC_COULOMB_REACTION_FIELD);
  if (el.key == ELECTROSTATIC_COULOMB_REACTION_FIELD) {   // coul_spole.F90:405-409
    const double b0 = 2.0 * (el.eps - 1.0) / (2.0 * el.eps + 1.0);
    el.reaction_field[0] = b0 / powi(rcut, 3);
    el.reaction_field[1] = (1.0 + 0.5 * b0) / rcut;
    el.reaction_field[2] = 0.5 * el.reaction_field[0];
  }
  if (el.damp) {                                          // :192-195, :411-414
    el.erfcgen(rcut, el.damping);
    el.force_shift = el.erfc_deriv.end_sample * rcut;
    el.energy_shift = -(el.erfc.end_sample + el.force_shift * rcut);
  } else if (el.key == ELECTROSTATIC_COULOMB_FORCE_SHIFT) {   // :198-199
    el.force_shift = 1.0 / powi(rcut, 2);
    el.energy_shift = -2.0 / rcut;
  }
}

// ---------------------------------------------------------------- neighbours.F90:182-284
bool vnl_decide(bool l_str, double tol_global, int bspline, neighbours_type& neigh, stats_type& stat, const domains_type& domain,
                const configuration_type& config, double& width) {
  neigh.update = tol_global >= half_minus() * neigh.padding;
  double celprp[10];
  dcell(config.cell, celprp);
  width = std::min(celprp[6], std::min(celprp[7], celprp[8]));
  double cut = neigh.cutoff_extended + smalldr;
  const double wdx = domain.nx_recip * celprp[6], wdy = domain.ny_recip * celprp[7], wdz = domain.nz_recip * celprp[8];
  const int ilx = f_int(wdx / cut), ily = f_int(wdy / cut), ilz = f_int(wdz / cut);
  const double m6 = 0.05, m7 = 0.005, m8 = 0.02, m9 = 0.95;
  const double tol = std::min(m6, m7 * neigh.cutoff);
  const double test = bspline > 0 ? m8 : m8 * 2.0;
  cut = std::min(wdx, std::min(wdy, wdz)) - smalldr;
  const char* msg307 = "neigh%cutoff <= Min(domain width) < neigh%cutoff_extended = neigh%cutoff + neigh%padding";
  if (ilx * ily * ilz == 0) {
    if (cut < neigh.cutoff) error(307, msg307);
    if (cut < neigh.cutoff_extended) {
      if (l_str) error(307, msg307);
      if (cut >= neigh.cutoff) {   // re-set the padding with some slack
        neigh.padding = std::min(m9 * (cut - neigh.cutoff), test * neigh.cutoff);
        neigh.padding = (double)f_int(100.0 * neigh.padding) / 100.0;
        if (neigh.padding < tol) neigh.padding = 0.0;
        neigh.cutoff_extended = neigh.cutoff + neigh.padding;
        neigh.update = true;
      }
    }
  } else if (neigh.update && !l_str) {   // push the limits when up for an update in a 'no strict' regime
    if (f_int((double)std::min(ilx, std::min(ily, ilz)) / (1.0 + test)) >= 2) {
      cut = test * neigh.cutoff;
    } else if (domain.mxnode > 1) {
      cut = std::min(m9 * (std::min(wdx / (double)ilx, std::min(wdy / (double)ily, wdz / (double)ilz)) - neigh.cutoff - smalldr),
                     test * neigh.cutoff);
    } else {
      cut = m9 * (0.5 * width - neigh.cutoff - smalldr);
    }
    cut = (double)f_int(100.0 * cut) / 100.0;
    if (!(cut < tol) && cut - neigh.padding > 0.005) {
      neigh.padding = cut;
      neigh.cutoff_extended = neigh.cutoff + neigh.padding;
    }
  }
  double* ns = stat.neighskip;   // ns[0..4] == neighskip(1:5)
  if (neigh.update) {
    ns[2] = ns[1] * ns[2];
    ns[1] = ns[1] + 1.0;
    ns[2] = ns[2] / ns[1] + ns[0] / ns[1];
    if (!neigh.newstart) ns[3] = std::min(ns[0], ns[3]);
    else neigh.newstart = false;
    ns[4] = std::max(ns[0], ns[4]);
    ns[0] = 0.0;
  } else {
    ns[0] = ns[0] + 1.0;
  }
  return neigh.update;
}

void exchange_capacities(const double cell[9], int megatm, const domains_type& d, double rcut, double padding, double safety,
                         int& cap_r, int& cap_h) {
  double celprp[10];
  dcell(cell, celprp);
  const double rho = (double)megatm / celprp[9], rx = rcut + padding;
  const double wid[3] = {celprp[6] / d.nx, celprp[7] / d.ny, celprp[8] / d.nz};
  double lw[3], full[3];
  for (int a = 0; a < 3; ++a) {
    lw[a] = wid[a] / std::max(f_int(wid[a] / (rx + smalldr)), 1);
    full[a] = wid[a] + 2.0 * lw[a];
  }
  const double face = std::max(full[1] * full[2] * lw[0], std::max(full[0] * full[2] * lw[1], full[0] * full[1] * lw[2]));
  const double area = std::max(wid[1] * wid[2], std::max(wid[0] * wid[2], wid[0] * wid[1]));
  cap_h = (int)(safety * rho * face) + 4096;
  cap_r = (int)(safety * rho * area * std::max(padding, 0.05 * rx)) + 4096;
}

// ---------------------------------------------------------------- gpu_short_range
gpu_short_range::gpu_short_range(int device) {
  const int rc = dlpgpu_create(&ctx_, device);
  if (rc != 0 || !ctx_) error(rc ? rc : DLPGPU_ERR_CUDA, "dlpgpu_create failed on device " + std::to_string(device) +
                                                              " (no CUDA GPU visible?); the short-range path has no CPU fallback");
}

gpu_short_range::~gpu_short_range() {
  if (ctx_) dlpgpu_destroy(ctx_);
}

void gpu_short_range::ck(int rc) const {
  if (rc != 0) error(rc, dlpgpu_last_error(ctx_));
}

long long gpu_short_range::launch_count() const { return dlpgpu_launch_count(ctx_); }

void gpu_short_range::init(const domains_type& d, const configuration_type& c, const neighbours_type& n) {
  const int dd[6] = {d.nx, d.ny, d.nz, d.idx, d.idy, d.idz};
  ck(dlpgpu_set_domain(ctx_, dd));
  ck(dlpgpu_set_cell(ctx_, c.cell, c.imcon));
  ck(dlpgpu_set_cutoffs(ctx_, n.cutoff, n.padding, n.pdplnc));
  dom_ = d;
  std::memcpy(cell_, c.cell, sizeof cell_);
  rcut_ = n.cutoff;
  padding_ = n.padding;
}

void gpu_short_range::set_forcefield(const vdw_type& v, const electrostatic_type& el, const ewald_type& ew, double rcut) {
  (void)rcut;
  ck(dlpgpu_set_vdw(ctx_, v.ntype_atom, v.list.data(), v.max_vdw, v.n_vdw, v.ltp.data(), v.max_grid, v.tab_potential.data(),
                    v.tab_force.data(), v.cutoff, v.l_force_shift ? 1 : 0, v.l_direct ? 1 : 0, v.param.data(), v.afs.data(),
                    v.bfs.data()));
  const double scaling = r4pie0 / el.eps;   // two_body.F90:188
  if (el.key == ELECTROSTATIC_SPME && ew.active) {
    ck(dlpgpu_set_ewald(ctx_, 1, ew.alpha, scaling, el.erfc.nsamples, el.erfc.table.data(), el.erfc_deriv.table.data(),
                        el.erfc.recip_spacing));
  } else if (el.key == ELECTROSTATIC_NULL) {
    ck(dlpgpu_set_ewald(ctx_, 0, 0.0, 0.0, 0, nullptr, nullptr, 0.0));
  } else {
    int kind = 0;
    switch (el.key) {
      case ELECTROSTATIC_COULOMB: kind = DLPGPU_COUL_CP; break;
      case ELECTROSTATIC_DDDP: kind = DLPGPU_COUL_DDDP; break;
      case ELECTROSTATIC_COULOMB_FORCE_SHIFT: kind = DLPGPU_COUL_FSCP; break;
      case ELECTROSTATIC_COULOMB_REACTION_FIELD: kind = DLPGPU_COUL_RFP; break;
      default: error(DLPGPU_ERR_ARG, "electrostatics key " + std::to_string(el.key) + " is not part of the short-range GPU path");
    }
    if (el.damp)
      ck(dlpgpu_set_coulomb(ctx_, kind, 1, scaling, el.force_shift, el.energy_shift, el.reaction_field, el.erfc.nsamples,
                            el.erfc.table.data(), el.erfc_deriv.table.data(), el.erfc.recip_spacing));
    else
      ck(dlpgpu_set_coulomb(ctx_, kind, 0, scaling, el.force_shift, el.energy_shift, el.reaction_field, 0, nullptr, nullptr, 0.0));
  }
}

void gpu_short_range::vnl_check(bool l_str, double& width, neighbours_type& neigh, stats_type& stat, const domains_type& domain,
                                const configuration_type& config, int bspline, gmax_fn gmax, void* gmax_user) {
  if (!neigh.unconditional_update) return;   // neighbours.F90:141: padding == 0 leaves update = .true.
  double tol = 0.0;
  ck(dlpgpu_vnl_check(ctx_, config.natms, config.parts.data(), &tol));   // :157-174 on the device
  if (gmax) tol = gmax(tol, gmax_user);                                  // :176
  const double padding_before = neigh.padding;
  vnl_decide(l_str, tol, bspline, neigh, stat, domain, config, width);
  if (neigh.padding != padding_before) ck(dlpgpu_set_cutoffs(ctx_, neigh.cutoff, neigh.padding, neigh.pdplnc));
}

void gpu_short_range::set_host_threads(int nthreads) { ck(dlpgpu_set_host_threads(ctx_, nthreads)); }

void gpu_short_range::vnl_set_check(neighbours_type& neigh, const configuration_type& config) {
  if (!neigh.unconditional_update) return;
  neigh.newjob = false;
  ck(dlpgpu_vnl_set_check(ctx_, config.nlast, config.parts.data()));
}

void gpu_short_range::link_cell_pairs(bool lbook, int megfrz, neighbours_type& neigh, const configuration_type& c, bool want_host_list) {
  if (want_host_list) neigh.list.assign((size_t)(neigh.max_list + 4) * c.natms, 0);
  int ibig = 0;
  const int rc = dlpgpu_link_cell_pairs(ctx_, c.natms, c.nlast, c.parts.data(), c.ltype.data(), c.ltg.data(),
                                        c.lfrzn.empty() ? nullptr : c.lfrzn.data(), lbook ? 1 : 0, megfrz, neigh.max_exclude,
                                        (lbook && !neigh.list_excl.empty()) ? neigh.list_excl.data() : nullptr, neigh.max_list,
                                        want_host_list ? neigh.list.data() : nullptr, &ibig);
  if (rc == DLPGPU_ERR_LIST_OVERFLOW)   // neighbours.F90:1189-1194: warning(290) with the largest row, then error(106)
    error(106, "neighbour list array exceeded: largest row " + std::to_string(ibig) + " > max_list " + std::to_string(neigh.max_list));
  ck(rc);
  neigh.newjob = false;   // the call took the vnl_set_check snapshot (halo.F90:315)
}

void gpu_short_range::two_body_forces(configuration_type& c, stats_type& stats, double& engvdw, double& virvdw, double& engcpe_rl,
                                      double& vircpe_rl, double& engcpe_ex, double& vircpe_ex, bool list_just_built) {
  if (list_just_built) ck(dlpgpu_parts_unchanged_since_list(ctx_));
  double out[16];
  ck(dlpgpu_two_body_forces(ctx_, c.natms, c.nlast, c.parts.data(), out));
  engvdw += out[0]; virvdw += out[1];
  engcpe_rl += out[2]; vircpe_rl += out[3];
  engcpe_ex += out[4]; vircpe_ex += out[5];
  for (int k = 0; k < 9; ++k) stats.stress[k] += out[6 + k];
}

void gpu_short_range::rdf_collect(int ntype_atom, const std::vector<int>& rdf_list, int n_pairs, int max_grid, std::vector<double>& rdf) {
  if (rdf.size() != (size_t)n_pairs * max_grid) rdf.assign((size_t)n_pairs * max_grid, 0.0);
  ck(dlpgpu_rdf_collect(ctx_, ntype_atom, rdf_list.data(), n_pairs, max_grid, rdf.data()));
}

// ---- native device-resident mode
void gpu_short_range::dev_setup(const configuration_type& c, const neighbours_type& n, const std::vector<int>& type_site,
                                const std::vector<double>& charge_site, const std::vector<int>& freeze_site,
                                const std::vector<double>& weight_site, const std::vector<int>* excl_by_gid, int max_exclude) {
  ck(dlpgpu_dev_set_sites(ctx_, (int)type_site.size(), type_site.data(), charge_site.data(), freeze_site.data(), weight_site.data()));
  if (excl_by_gid) ck(dlpgpu_dev_set_excl(ctx_, c.megatm, max_exclude, excl_by_gid->data()));
  ck(dlpgpu_dev_set_list_capacity(ctx_, n.max_list, c.megfrz));
  int cap_r = 0, cap_h = 0;
  exchange_capacities(cell_, c.megatm, dom_, rcut_, padding_, 2.0, cap_r, cap_h);
  unsigned char handle[DLPGPU_XCHG_BLOB];
  ck(dlpgpu_dev_xchg_init(ctx_, dom_.idnode, dom_.mxnode, cap_r, cap_h, handle));   // one domain: no IPC involved
}

void gpu_short_range::dev_load(const std::vector<double>& xyz, const std::vector<double>& vel, const std::vector<int>& ltg,
                               const std::vector<int>& lsite) {
  ck(dlpgpu_dev_load_atoms(ctx_, (int)ltg.size(), xyz.data(), vel.empty() ? nullptr : vel.data(), ltg.data(), lsite.data(), 0));
}

void gpu_short_range::dev_first_forces(double out[16]) {
  int ibig = 0;
  ck(dlpgpu_dev_relocate_serial(ctx_));
  ck(dlpgpu_dev_halo_serial(ctx_));
  ck(dlpgpu_dev_link_cell_pairs(ctx_, 0, &ibig));
  ck(dlpgpu_dev_two_body_forces(ctx_, 1, out));
}

void gpu_short_range::dev_md_step(double dt, bool& rebuilt, double out_prev[16], bool& have_prev) {
  static const int self[6] = {0, 0, 0, 0, 0, 0};   // one domain: the rank is its own neighbour in every direction
  int reb = 0, have = 0;
  double list_ms = 0.0;
  ck(dlpgpu_dev_md_step(ctx_, self, dt, ++gseq_, rseq_ + 1, &reb, out_prev, &have, &list_ms));
  if (reb) ++rseq_;
  rebuilt = reb != 0;
  have_prev = have != 0;
}

void gpu_short_range::dev_fetch_results(double out[16]) { ck(dlpgpu_dev_fetch_results(ctx_, out)); }

void gpu_short_range::dev_get(configuration_type& c, std::vector<double>* vel) {
  ck(dlpgpu_dev_counts(ctx_, &c.natms, &c.nlast));
  c.parts.resize(c.nlast);
  ck(dlpgpu_dev_get_parts(ctx_, c.parts.data(), c.nlast));
  c.ltg.resize(c.nlast); c.lsite.resize(c.nlast); c.ltype.resize(c.nlast); c.lfrzn.resize(c.nlast);
  std::vector<int> ixyz(c.nlast);
  ck(dlpgpu_dev_get_ints(ctx_, c.nlast, c.ltg.data(), c.lsite.data(), c.ltype.data(), c.lfrzn.data(), ixyz.data()));
  if (vel) {
    vel->resize((size_t)3 * c.natms);
    ck(dlpgpu_dev_get_vel(ctx_, c.natms, vel->data()));
  }
}

}   // namespace dlpoly
