// =============================================================================
// dlpoly_host.hpp -- C++ host side of the short-range two-body path, above the C ABI (include/dlpgpu.h).
//
// The reference is compiled Fortran; the build image has no Fortran compiler, so the host logic a DL_POLY rank runs around
// the two redirected call sites is mirrored here in C++ with the reference's names, argument meaning and error numbers:
//
//   map_domains              domains.F90:63-258
//   dcell / invert           numerics.F90:1344-1509
//   vdw_generate             vdw.F90:1397-1576       vdw_direct_fs_generate  vdw.F90:969-1049
//   vdw_table_read           vdw.F90:1051-1370       erfcgen                 electrostatic.F90:88-127
//   vnl_check                neighbours.F90:123-296  vnl_set_check           neighbours.F90:305-343
//   link_cell_pairs          neighbours.F90:356      (call site drivers.F90:675-679)
//   two_body_forces          two_body.F90:339-606    (the two pair loops + the sums of :672-790 that belong to them)
//   rdf_collect              rdfs.F90:146-212, :880-946
//
// Everything numeric on the GPU goes through libdlpgpu.so; nothing here evaluates a pair.  There is no CPU fallback:
// gpu_short_range's constructor raises when dlpgpu_create fails.
// Errors: error(kode, message) of errors_warnings.F90:840 aborts all ranks in the reference; here it throws
// dlpoly::dlpoly_error carrying kode.
// =============================================================================
#ifndef DLPOLY_HOST_HPP
#define DLPOLY_HOST_HPP

#include <cfloat>
#include <cmath>
#include <stdexcept>
#include <string>
#include <vector>

#include "dlpgpu.h"

namespace dlpoly {

struct dlpoly_error : std::runtime_error {
  int kode;
  dlpoly_error(int k, const std::string& m) : std::runtime_error("DL_POLY error " + std::to_string(k) + ": " + m), kode(k) {}
};
[[noreturn]] void error(int kode, const std::string& message = "");

using corePart = dlpgpu_corepart;   // particle.F90:14-20

// constants.F90:53-58, 100, 139, 189-204
constexpr double pi = 3.14159265358979312;
constexpr double sqrpi = 1.7724538509055159;
constexpr double r4pie0 = 138935.4835;
constexpr double delr_max = 0.01;
constexpr double smalldr = 1.0e-6;
constexpr double zero_plus = DBL_MIN;
inline double half_minus() { return std::nextafter(0.5, 0.0); }

// vdw.F90:64-117 (the potential keys this path evaluates)
enum : int { VDW_NULL = -1, VDW_TAB = 0, VDW_12_6 = 1, VDW_LENNARD_JONES = 2, VDW_BUCKINGHAM = 4, VDW_BORN_HUGGINS_MEYER = 5 };

// ---------------------------------------------------------------- numerics.F90
void dcell(const double aaa[9], double bbb[10]);          // bbb[0..9] == Fortran bbb(1:10); widths are bbb[6..8]
void invert(const double a[9], double b[9], double& d);

// ---------------------------------------------------------------- domains.F90
struct domains_type {
  int nx = 1, ny = 1, nz = 1, idx = 0, idy = 0, idz = 0;
  double nx_recip = 1.0, ny_recip = 1.0, nz_recip = 1.0;
  int map[26] = {0};          // ranks of the 26 neighbours, map[0..5] = -x,+x,-y,+y,-z,+z   (domains.F90:206-243)
  int map_unique[26] = {0};   // 1 where the entry repeats an earlier one or is the rank itself (:251-256)
  int mxnode = 1, idnode = 0;
};
void map_domains(int imcon, double wx, double wy, double wz, int idnode, int mxnode, domains_type& domain);

// ---------------------------------------------------------------- vdw.F90 vdw_type (the members this path reads)
struct vdw_type {
  int ntype_atom = 0;                 // sites%ntype_atom
  int n_vdw = 0, max_vdw = 0, max_grid = 0;
  double cutoff = 0.0;
  bool l_force_shift = false, l_direct = false, l_tab = false;
  std::vector<int> list;              // list(1:ntype(ntype+1)/2) -> potential number (1-based)
  std::vector<int> ltp;               // ltp(1:max_vdw)
  std::vector<double> param;          // param(1:7, 1:max_vdw), column-major
  std::vector<double> tab_potential;  // (0:max_grid, 1:max_vdw), column-major
  std::vector<double> tab_force;
  std::vector<double> afs, bfs;       // force-shift constants of vdw_forces_direct
  std::vector<std::string> unique_atom;   // sites%unique_atom (labels TABLE entries are matched against)
  double elrc = 0.0, vlrc = 0.0;      // long-range corrections (vdw_lrc)

  void init(int ntype_atom, double rvdw, bool force_shift, bool direct);   // bounds.F90:820 max_grid
  // read_field's vdw block (ffield.F90:3620-3960): one potential per unordered type pair; returns its number
  int add(int atom_i, int atom_j, int keypot, const double* prm, int nprm);
  void finalise();                    // ffield.F90:3939-3954: undefined pairs -> VDW_NULL slot past n_vdw
  static int key(int ai, int aj) { const int hi = ai > aj ? ai : aj, lo = ai > aj ? aj : ai; return hi * (hi - 1) / 2 + lo; }
};
int max_grid_of(double rcut);                                            // bounds.F90:811,820
int max_list_of(double fdens, double cutoff_extended);                   // bounds.F90:907
void pair_potential(int keypot, const double* prm, double r, double& energy, double& gamma);   // two_body_potentials.F90
void vdw_generate(vdw_type& vdws);
void vdw_direct_fs_generate(vdw_type& vdws);
void vdw_table_read(vdw_type& vdws, const std::string& filename, double engunit = 1.0);
// vdw.F90:617-967: num_type / numfrz = atoms / frozen atoms per type over the whole system (after gsum); sets elrc, vlrc
void vdw_lrc(const std::vector<double>& num_type, const std::vector<double>& numfrz, vdw_type& vdws, int imcon, double volm);

// ---------------------------------------------------------------- electrostatic.F90 / ewald.F90
struct interp_table {                 // numerics.F90:47-60
  std::vector<double> table;          // table[i] == Fortran table(i), i = 1..nsamples; table[0] unused
  int nsamples = 0;
  double spacing = 0.0, recip_spacing = 0.0, end_sample = 0.0;
};
enum : int { ELECTROSTATIC_NULL = 0, ELECTROSTATIC_SPME = 1, ELECTROSTATIC_DDDP = 2, ELECTROSTATIC_COULOMB = 3,
             ELECTROSTATIC_COULOMB_FORCE_SHIFT = 4, ELECTROSTATIC_COULOMB_REACTION_FIELD = 5 };   // electrostatic.F90:18-28
struct electrostatic_type {
  int key = ELECTROSTATIC_NULL;
  double eps = 1.0, damping = 0.0;
  bool damp = false;
  double force_shift = 0.0, energy_shift = 0.0, reaction_field[3] = {0.0, 0.0, 0.0};
  interp_table erfc, erfc_deriv;
  void erfcgen(double rcut, double alpha);   // electrostatic.F90:88-127
};
struct ewald_type { bool active = false; double alpha = 0.0; };
double ewald_alpha_from_precision(double precision, double rcut);        // control.F90:1709-1710
double calc_erfc(double x);                                               // numerics.F90:3647-3667
// coul_spole.F90:186-202 (fscp) / :399-417 (rfp): the constants the direct-space variants keep in electrostatic_type
void coul_setup(electrostatic_type& electro, double rcut);

// ---------------------------------------------------------------- neighbours.F90 neighbours_type, statistics
struct neighbours_type {
  double cutoff = 0.0, padding = 0.0, cutoff_extended = 0.0, pdplnc = 50.0;
  bool unconditional_update = false;   // .true. when padding > 0 (bounds.F90:1366)
  bool update = true, newstart = true, newjob = true;
  int max_list = 0, max_exclude = 0;
  std::vector<int> list;               // list(-3:max_list, 1:natms), column-major: row i contiguous
  std::vector<int> list_excl;          // list_excl(0:max_exclude, 1:natms)
  int& l(int k, int i) { return list[(size_t)(i - 1) * (max_list + 4) + (k + 3)]; }
};
struct stats_type {
  double neighskip[5] = {0.0, 0.0, 0.0, 999999999.0, 0.0};   // cycles, accesses, average, minimum, maximum (statistics.F90:185-186)
  double engsrp = 0.0, virsrp = 0.0, engcpe = 0.0, vircpe = 0.0;
  double stress[9] = {0.0};
};
struct configuration_type {
  int imcon = 1, natms = 0, nlast = 0, megatm = 0, megfrz = 0;
  double cell[9] = {0.0};
  double volm = 0.0, sumchg = 0.0;     // cell volume, total system charge
  std::vector<corePart> parts;         // parts(1:nlast)
  std::vector<int> ltg, lsite, ltype, lfrzn;
};

// configuration.F90:1183-1205 (read_config): fold every atom into the reduced cell [-0.5, 0.5), recompute its Cartesian position
// from cell . s (so positions carry the bits a DL_POLY run starts from) and assign it to its domain
// idm = ipx + nx (ipy + ny ipz), ip = Int((s + 0.5) n).  xyz(3, n) in / out; owner(n) out (0-based rank).
void read_config_fold(std::vector<double>& xyz, const double cell[9], const domains_type& domain, std::vector<int>& owner);

// The end of two_body_forces (two_body.F90:672-790) for the terms of this path, AFTER the caller's gsum of the six partial sums
// (:729): Fuchs' net-charge correction (SPME only), stats%engcpe / vircpe / engsrp / virsrp incl. the long-range corrections,
// and the per-rank share of the corrections on the stress diagonal.  engcpe_rc / vircpe_rc: the caller's reciprocal-space sums.
void two_body_totals(stats_type& stats, const vdw_type& vdws, const electrostatic_type& electro, const ewald_type& ewld,
                     const configuration_type& config, int mxnode, double engvdw, double virvdw, double engcpe_rc, double vircpe_rc,
                     double engcpe_rl, double vircpe_rl, double engcpe_ex, double vircpe_ex);

// The padding / update decision of vnl_check for a displacement maximum that is already global (after gmax): everything of
// neighbours.F90:182-284 except the KIM clause.  bspline > 0 <=> SPME is on.  Returns neigh.update.
bool vnl_decide(bool l_str, double tol_global, int bspline, neighbours_type& neigh, stats_type& stat, const domains_type& domain,
                const configuration_type& config, double& width);

// Per-stage receive capacities (atoms) of the device-side exchange, identical on every rank: the halo slab of the widest
// face (link-cell width, halo.F90:219-239, grown by the layers received in the earlier directions) at the mean density times
// `safety`; migration: a layer of one padding thickness across that face (an atom moves < padding / 2 between rebuilds).
void exchange_capacities(const double cell[9], int megatm, const domains_type& domain, double rcut, double padding, double safety,
                         int& cap_reloc_atoms, int& cap_halo_atoms);

// ---------------------------------------------------------------- the GPU engine of one rank
// gmax over ranks for vnl_check; the default (one rank) is the identity.  An MPI host passes its own.
using gmax_fn = double (*)(double local, void* user);

class gpu_short_range {
 public:
  explicit gpu_short_range(int device);          // raises 9001 when no GPU / library context can be created
  ~gpu_short_range();
  gpu_short_range(const gpu_short_range&) = delete;
  gpu_short_range& operator=(const gpu_short_range&) = delete;

  // once, after set_bounds / read_field / vdw_generate / erfcgen (dlp_gpu_init + dlp_gpu_set_forcefield)
  void init(const domains_type& domain, const configuration_type& config, const neighbours_type& neigh);
  void set_forcefield(const vdw_type& vdws, const electrostatic_type& electro, const ewald_type& ewld, double rcut);
  // how config.parts travels (include/dlpgpu.h: dlpgpu_set_host_threads): 0 = whole corePart records by DMA (default), n >= 1 =
  // n host threads of the library copy x, y, z up and ADD the returned forces into parts%f (worth it from ~12 spare cores)
  void set_host_threads(int nthreads);

  // neighbours.F90:123-296
  void vnl_check(bool l_str, double& width, neighbours_type& neigh, stats_type& stat, const domains_type& domain,
                 const configuration_type& config, int bspline, gmax_fn gmax = nullptr, void* gmax_user = nullptr);
  // neighbours.F90:305-343 (link_cell_pairs takes the same snapshot itself; this is the stand-alone call)
  void vnl_set_check(neighbours_type& neigh, const configuration_type& config);
  // neighbours.F90:356-1306; want_host_list fills neigh.list in the reference's format
  void link_cell_pairs(bool lbook, int megfrz, neighbours_type& neigh, const configuration_type& config, bool want_host_list);
  // two_body.F90:339-606: adds the pair forces to config.parts(1:natms) and this rank's partial sums to the arguments /
  // stats.stress; list_just_built: parts unchanged since link_cell_pairs (skips the upload)
  void two_body_forces(configuration_type& config, stats_type& stats, double& engvdw, double& virvdw, double& engcpe_rl,
                       double& vircpe_rl, double& engcpe_ex, double& vircpe_ex, bool list_just_built = false);
  // rdfs.F90:146-212 / :880-946 on the device list; rdf(1:max_grid, 1:n_pairs) column-major, incremented
  void rdf_collect(int ntype_atom, const std::vector<int>& rdf_list, int n_pairs, int max_grid, std::vector<double>& rdf);

  // ---- native device-resident mode (no per-step host buffers): a single-domain NVE run around the path
  void dev_setup(const configuration_type& config, const neighbours_type& neigh, const std::vector<int>& type_site,
                 const std::vector<double>& charge_site, const std::vector<int>& freeze_site, const std::vector<double>& weight_site,
                 const std::vector<int>* excl_by_gid, int max_exclude);
  void dev_load(const std::vector<double>& xyz, const std::vector<double>& vel, const std::vector<int>& ltg, const std::vector<int>& lsite);
  void dev_first_forces(double out[16]);    // relocate + halo + list + forces of the starting configuration
  // md_vv (drivers.F90:1910-2290) around the path, one step; out_prev / have_prev as dlpgpu_dev_md_step
  void dev_md_step(double dt, bool& rebuilt, double out_prev[16], bool& have_prev);
  void dev_fetch_results(double out[16]);
  void dev_get(configuration_type& config, std::vector<double>* vel);

  dlpgpu_ctx* handle() { return ctx_; }
  long long launch_count() const;

 private:
  void ck(int rc) const;
  dlpgpu_ctx* ctx_ = nullptr;
  unsigned long long gseq_ = 0, rseq_ = 0;
  domains_type dom_;
  double cell_[9] = {0.0}, rcut_ = 0.0, padding_ = 0.0;
};

}   // namespace dlpoly
#endif
