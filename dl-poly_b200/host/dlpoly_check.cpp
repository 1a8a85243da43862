// dlpoly_check -- drives the C++ host side (dlpoly_host.hpp) from a binary bundle and writes its results to another one, so
// that tests/test_host_cpp.py can hold it against the oracle.  Modes:
//   host    <in> <out>   CPU only: dcell / invert / map_domains / table generation / vnl_check decision logic
//   dropin  <in> <out>   one DL_POLY domain through link_cell_pairs + vnl_check + two_body_forces (host buffers, like the
//                        Fortran call sites of INTEGRATION.md)
//   md      <in> <out>   native device-resident NVE run of one domain (md_vv around the path)
// Bundle record: char name[24]; char dtype ('d' double, 'i' int32, 'b' bytes); char pad[7]; int64 count; payload.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "dlpoly_host.hpp"

using namespace dlpoly;

namespace {

struct Record {
  char dtype = 'b';
  std::vector<char> data;
  int64_t count = 0;
};

struct Bundle {
  std::map<std::string, Record> rec;
  std::vector<std::string> order;

  void read(const char* path) {
    FILE* f = std::fopen(path, "rb");
    if (!f) error(DLPGPU_ERR_ARG, std::string("cannot open ") + path);
    char head[32];
    while (std::fread(head, 1, 32, f) == 32) {
      int64_t count;
      if (std::fread(&count, 8, 1, f) != 1) break;
      Record r;
      r.dtype = head[24];
      r.count = count;
      const size_t w = r.dtype == 'd' ? 8 : (r.dtype == 'i' ? 4 : 1);
      r.data.resize((size_t)count * w);
      if (count && std::fread(r.data.data(), 1, r.data.size(), f) != r.data.size()) error(DLPGPU_ERR_ARG, "truncated bundle");
      head[23] = 0;
      rec[head] = std::move(r);
    }
    std::fclose(f);
  }
  bool has(const std::string& n) const { return rec.count(n) != 0; }
  const Record& get(const std::string& n) const {
    auto it = rec.find(n);
    if (it == rec.end()) error(DLPGPU_ERR_ARG, "bundle has no record '" + n + "'");
    return it->second;
  }
  std::vector<double> d(const std::string& n) const {
    const Record& r = get(n);
    std::vector<double> v((size_t)r.count);
    std::memcpy(v.data(), r.data.data(), r.data.size());
    return v;
  }
  std::vector<int> i(const std::string& n) const {
    const Record& r = get(n);
    std::vector<int> v((size_t)r.count);
    std::memcpy(v.data(), r.data.data(), r.data.size());
    return v;
  }
  std::string s(const std::string& n) const {
    const Record& r = get(n);
    return std::string(r.data.begin(), r.data.end());
  }
  double d1(const std::string& n) const { return d(n).at(0); }
  int i1(const std::string& n) const { return i(n).at(0); }

  void put(const std::string& n, char dtype, const void* p, int64_t count) {
    Record r;
    r.dtype = dtype;
    r.count = count;
    const size_t w = dtype == 'd' ? 8 : (dtype == 'i' ? 4 : 1);
    r.data.resize((size_t)count * w);
    if (count) std::memcpy(r.data.data(), p, r.data.size());
    if (!rec.count(n)) order.push_back(n);
    rec[n] = std::move(r);
  }
  void put(const std::string& n, const std::vector<double>& v) { put(n, 'd', v.data(), (int64_t)v.size()); }
  void put(const std::string& n, const std::vector<int>& v) { put(n, 'i', v.data(), (int64_t)v.size()); }
  void write(const char* path) const {
    FILE* f = std::fopen(path, "wb");
    if (!f) error(DLPGPU_ERR_ARG, std::string("cannot write ") + path);
    for (const std::string& n : order) {
      const Record& r = rec.at(n);
      char head[32] = {0};
      std::strncpy(head, n.c_str(), 23);
      head[24] = r.dtype;
      std::fwrite(head, 1, 32, f);
      std::fwrite(&r.count, 8, 1, f);
      if (r.count) std::fwrite(r.data.data(), 1, r.data.size(), f);
    }
    std::fclose(f);
  }
};

// what read_field / set_bounds leave behind for this path, from the bundle's force-field records
struct ForceField {
  vdw_type vdws;
  electrostatic_type electro;
  ewald_type ewld;
};

void build_forcefield(const Bundle& in, ForceField& ff) {
  const double rcut = in.d1("rcut"), rvdw = in.d1("rvdw");
  ff.vdws.init(in.i1("ntypes"), rvdw, in.i1("force_shift") != 0, in.i1("direct") != 0);
  if (in.has("unique_atom")) {
    std::string all = in.s("unique_atom"), cur;
    for (char c : all) {
      if (c == '\n') { ff.vdws.unique_atom.push_back(cur); cur.clear(); }
      else cur.push_back(c);
    }
    if (!cur.empty()) ff.vdws.unique_atom.push_back(cur);
  }
  const std::vector<int> pairs = in.i("pot_pairs");       // (ai, aj, keypot) per potential, FIELD order
  const std::vector<double> prm = in.d("pot_param");      // 7 per potential
  for (size_t k = 0; k < pairs.size() / 3; ++k) ff.vdws.add(pairs[3 * k], pairs[3 * k + 1], pairs[3 * k + 2], &prm[7 * k], 7);
  ff.vdws.finalise();
  vdw_generate(ff.vdws);
  if (ff.vdws.l_tab) vdw_table_read(ff.vdws, in.s("table_file"), in.has("engunit") ? in.d1("engunit") : 1.0);
  if (ff.vdws.l_force_shift && ff.vdws.l_direct) vdw_direct_fs_generate(ff.vdws);
  ff.electro.key = in.i1("electro_key");
  ff.electro.eps = in.d1("eps");
  ff.electro.damping = in.d1("damping");
  if (ff.electro.key == ELECTROSTATIC_SPME) {
    ff.ewld.active = true;
    ff.ewld.alpha = in.has("ew_alpha") ? in.d1("ew_alpha") : ewald_alpha_from_precision(in.d1("ew_precision"), rcut);
    ff.electro.erfcgen(rcut, ff.ewld.alpha);
  } else if (ff.electro.key != ELECTROSTATIC_NULL) {
    coul_setup(ff.electro, rcut);
  }
}

void put_forcefield(Bundle& out, const ForceField& ff) {
  out.put("vdw_list", ff.vdws.list);
  out.put("ltp", ff.vdws.ltp);
  out.put("param", ff.vdws.param);
  out.put("tab_potential", ff.vdws.tab_potential);
  out.put("tab_force", ff.vdws.tab_force);
  out.put("afs", ff.vdws.afs);
  out.put("bfs", ff.vdws.bfs);
  out.put("vdw_sizes", std::vector<int>{ff.vdws.n_vdw, ff.vdws.max_vdw, ff.vdws.max_grid});
  out.put("erfc", ff.electro.erfc.table);
  out.put("erfc_deriv", ff.electro.erfc_deriv.table);
  out.put("electro", std::vector<double>{ff.ewld.alpha, ff.electro.erfc.recip_spacing, ff.electro.force_shift, ff.electro.energy_shift,
                                         ff.electro.reaction_field[0], ff.electro.reaction_field[1], ff.electro.reaction_field[2],
                                         ff.electro.damp ? 1.0 : 0.0});
}

void fill_config(const Bundle& in, configuration_type& c, neighbours_type& n, domains_type& dom) {
  const std::vector<double> cell = in.d("cell");
  std::memcpy(c.cell, cell.data(), sizeof c.cell);
  c.imcon = in.i1("imcon");
  c.megatm = in.i1("megatm");
  c.megfrz = in.has("megfrz") ? in.i1("megfrz") : 0;
  n.cutoff = in.d1("rcut");
  n.padding = in.d1("padding");
  n.cutoff_extended = n.cutoff + n.padding;
  n.pdplnc = in.has("pdplnc") ? in.d1("pdplnc") : 50.0;
  n.unconditional_update = n.padding > 0.0;                    // bounds.F90:1366
  n.max_list = in.has("max_list") ? in.i1("max_list") : 0;
  double celprp[10];
  dcell(c.cell, celprp);
  const int mxnode = in.has("mxnode") ? in.i1("mxnode") : 1, idnode = in.has("idnode") ? in.i1("idnode") : 0;
  map_domains(c.imcon, celprp[6], celprp[7], celprp[8], idnode, mxnode, dom);
}

// ---------------------------------------------------------------- mode host
int mode_host(const Bundle& in, Bundle& out) {
  const std::vector<double> cell = in.d("cell");
  double celprp[10], rcell[9], det;
  dcell(cell.data(), celprp);
  invert(cell.data(), rcell, det);
  out.put("celprp", 'd', celprp, 10);
  out.put("rcell", 'd', rcell, 9);
  out.put("det", 'd', &det, 1);
  if (in.has("dd_cases")) {   // (imcon, mxnode, idnode) triples with widths from dd_widths (3 per case)
    const std::vector<int> cases = in.i("dd_cases");
    const std::vector<double> wid = in.d("dd_widths");
    std::vector<int> res;
    for (size_t k = 0; k < cases.size() / 3; ++k) {
      domains_type dom;
      map_domains(cases[3 * k], wid[3 * k], wid[3 * k + 1], wid[3 * k + 2], cases[3 * k + 2], cases[3 * k + 1], dom);
      const int six[6] = {dom.nx, dom.ny, dom.nz, dom.idx, dom.idy, dom.idz};
      res.insert(res.end(), six, six + 6);
      res.insert(res.end(), dom.map, dom.map + 26);
      res.insert(res.end(), dom.map_unique, dom.map_unique + 26);
    }
    out.put("dd_results", res);
  }
  if (in.has("pot_pairs") && !in.has("bad_table_file")) {
    ForceField ff;
    build_forcefield(in, ff);
    put_forcefield(out, ff);
    out.put("sizes", std::vector<int>{max_grid_of(in.d1("rcut")), max_list_of(in.d1("fdens"), in.d1("rcut") + in.d1("padding"))});
    const double a = ewald_alpha_from_precision(in.has("ew_precision") ? in.d1("ew_precision") : 1.0e-6, in.d1("rcut"));
    out.put("alpha_from_precision", 'd', &a, 1);
  }
  if (in.has("cap_cases")) {   // exchange_capacities: (mxnode, megatm) pairs with (rcut, padding) pairs
    const std::vector<int> cs = in.i("cap_cases");
    const std::vector<double> cp2 = in.d("cap_cutoffs");
    std::vector<int> res;
    double w[10];
    dcell(cell.data(), w);
    for (size_t k = 0; k < cs.size() / 2; ++k) {
      domains_type dom;
      map_domains(in.i1("imcon"), w[6], w[7], w[8], 0, cs[2 * k], dom);
      int cr = 0, ch = 0;
      exchange_capacities(cell.data(), cs[2 * k + 1], dom, cp2[2 * k], cp2[2 * k + 1], 2.0, cr, ch);
      res.push_back(cr);
      res.push_back(ch);
    }
    out.put("cap_results", res);
  }
  if (in.has("fold_xyz")) {   // read_config: fold + domain assignment
    std::vector<double> xyz = in.d("fold_xyz");
    std::vector<int> owner;
    domains_type dom;
    double cp[10];
    dcell(cell.data(), cp);
    map_domains(in.i1("imcon"), cp[6], cp[7], cp[8], 0, in.i1("mxnode"), dom);
    read_config_fold(xyz, cell.data(), dom, owner);
    out.put("folded_xyz", xyz);
    out.put("owner", owner);
  }
  if (in.has("num_type")) {   // vdw_lrc + the end of two_body_forces on given (already global) partial sums
    ForceField ff;
    build_forcefield(in, ff);
    configuration_type c;
    c.imcon = in.i1("imcon");
    c.volm = in.d1("volm");
    c.sumchg = in.d1("sumchg");
    vdw_lrc(in.d("num_type"), in.d("numfrz"), ff.vdws, c.imcon, c.volm);
    out.put("lrc", std::vector<double>{ff.vdws.elrc, ff.vdws.vlrc});
    const std::vector<double> p = in.d("partial_sums");   // engvdw, virvdw, engcpe_rl, vircpe_rl, engcpe_ex, vircpe_ex, engcpe_rc, vircpe_rc
    stats_type st;
    const std::vector<double> s0 = in.d("stress_in");
    for (int k = 0; k < 9; ++k) st.stress[k] = s0[k];
    two_body_totals(st, ff.vdws, ff.electro, ff.ewld, c, in.i1("mxnode"), p[0], p[1], p[6], p[7], p[2], p[3], p[4], p[5]);
    out.put("totals", std::vector<double>{st.engcpe, st.vircpe, st.engsrp, st.virsrp});
    out.put("stress_out", 'd', st.stress, 9);
  }
  if (in.has("vnl_tols")) {   // a sequence of global displacement maxima through vnl_check's decision logic
    configuration_type c;
    neighbours_type n;
    domains_type dom;
    stats_type st;
    fill_config(in, c, n, dom);
    const std::vector<double> tols = in.d("vnl_tols");
    const bool l_str = in.i1("l_str") != 0;
    const int bspline = in.i1("bspline");
    std::vector<double> trace;
    std::vector<int> kode;
    for (double t : tols) {
      double width = 0.0;
      int k = 0;
      try {
        vnl_decide(l_str, t, bspline, n, st, dom, c, width);
      } catch (const dlpoly_error& e) {
        k = e.kode;
      }
      kode.push_back(k);
      const double row[10] = {n.update ? 1.0 : 0.0, n.padding, n.cutoff_extended, width, st.neighskip[0], st.neighskip[1],
                              st.neighskip[2], st.neighskip[3], st.neighskip[4], n.newstart ? 1.0 : 0.0};
      trace.insert(trace.end(), row, row + 10);
      if (k) break;
    }
    out.put("vnl_trace", trace);
    out.put("vnl_kode", kode);
  }
  if (in.has("bad_table_file")) {   // error paths of vdw_table_read: the kode each file raises
    ForceField ff;
    int k = -1;
    try {
      Bundle b2 = in;
      Record r = in.get("bad_table_file");
      b2.rec["table_file"] = r;
      build_forcefield(b2, ff);
    } catch (const dlpoly_error& e) {
      k = e.kode;
    }
    out.put("bad_table_kode", std::vector<int>{k});
  }
  return 0;
}

// ---------------------------------------------------------------- mode dropin
int mode_dropin(const Bundle& in, Bundle& out) {
  configuration_type c;
  neighbours_type n;
  domains_type dom;
  stats_type st;
  ForceField ff;
  fill_config(in, c, n, dom);
  build_forcefield(in, ff);
  c.natms = in.i1("natms");
  c.nlast = in.i1("nlast");
  const Record& pr = in.get("parts");
  c.parts.resize((size_t)c.nlast);
  if (pr.data.size() != c.parts.size() * sizeof(corePart)) error(DLPGPU_ERR_ARG, "parts record has the wrong size");
  std::memcpy(c.parts.data(), pr.data.data(), pr.data.size());
  c.ltype = in.i("ltype");
  c.ltg = in.i("ltg");
  c.lfrzn = in.i("lfrzn");
  const bool lbook = in.i1("lbook") != 0;
  if (lbook) {
    n.max_exclude = in.i1("max_exclude");
    n.list_excl = in.i("list_excl");
  }
  gpu_short_range gpu(in.has("device") ? in.i1("device") : 0);
  gpu.init(dom, c, n);
  gpu.set_forcefield(ff.vdws, ff.electro, ff.ewld, n.cutoff);
  if (in.has("force_mode")) {
    const int rc = dlpgpu_set_force_mode(gpu.handle(), in.i1("force_mode"));
    if (rc) error(rc, dlpgpu_last_error(gpu.handle()));
  }
  // calculate_forces (drivers.F90:675-679 then two_body_forces): list, forces on the records just uploaded
  gpu.link_cell_pairs(lbook, c.megfrz, n, c, /*want_host_list=*/true);
  double engvdw = 0.0, virvdw = 0.0, engcpe_rl = 0.0, vircpe_rl = 0.0, engcpe_ex = 0.0, vircpe_ex = 0.0;
  gpu.two_body_forces(c, st, engvdw, virvdw, engcpe_rl, vircpe_rl, engcpe_ex, vircpe_ex, /*list_just_built=*/true);
  out.put("list", n.list);
  out.put("parts", 'b', c.parts.data(), (int64_t)(c.parts.size() * sizeof(corePart)));
  out.put("sums", std::vector<double>{engvdw, virvdw, engcpe_rl, vircpe_rl, engcpe_ex, vircpe_ex});
  out.put("stress", 'd', st.stress, 9);
  // vnl_check against the checkpoint link_cell_pairs took: shift every local atom by `vnl_shift` first
  if (in.has("vnl_shift")) {
    const std::vector<double> sh = in.d("vnl_shift");
    for (int i = 0; i < c.natms; ++i) { c.parts[i].xxx += sh[0]; c.parts[i].yyy += sh[1]; c.parts[i].zzz += sh[2]; }
    double width = 0.0;
    gpu.vnl_check(in.i1("l_str") != 0, width, n, st, dom, c, in.i1("bspline"));
    out.put("vnl", std::vector<double>{n.update ? 1.0 : 0.0, n.padding, n.cutoff_extended, width, st.neighskip[0], st.neighskip[1]});
  }
  if (in.has("rdf_list")) {
    std::vector<double> rdf;
    gpu.rdf_collect(ff.vdws.ntype_atom, in.i("rdf_list"), in.i1("rdf_pairs"), in.i1("rdf_grid"), rdf);
    out.put("rdf", rdf);
  }
  const long long nl = gpu.launch_count();
  out.put("launches", std::vector<int>{(int)nl});
  return 0;
}

// ---------------------------------------------------------------- mode md
int mode_md(const Bundle& in, Bundle& out) {
  configuration_type c;
  neighbours_type n;
  domains_type dom;
  ForceField ff;
  fill_config(in, c, n, dom);
  build_forcefield(in, ff);
  gpu_short_range gpu(in.has("device") ? in.i1("device") : 0);
  gpu.init(dom, c, n);
  gpu.set_forcefield(ff.vdws, ff.electro, ff.ewld, n.cutoff);
  std::vector<int> excl;
  const bool lbook = in.has("excl_by_gid");
  if (lbook) excl = in.i("excl_by_gid");
  gpu.dev_setup(c, n, in.i("type_site"), in.d("charge_site"), in.i("freeze_site"), in.d("weight_site"), lbook ? &excl : nullptr,
                lbook ? in.i1("max_exclude") : 0);
  if (in.has("force_mode")) {
    const int rc = dlpgpu_set_force_mode(gpu.handle(), in.i1("force_mode"));
    if (rc) error(rc, dlpgpu_last_error(gpu.handle()));
  }
  std::vector<double> xyz = in.d("xyz");   // CONFIG positions: folded into the cell like read_config does
  std::vector<int> owner;
  read_config_fold(xyz, c.cell, dom, owner);
  gpu.dev_load(xyz, in.has("vel") ? in.d("vel") : std::vector<double>(), in.i("ltg"), in.i("lsite"));
  const int nsteps = in.i1("nsteps");
  const double dt = in.d1("timestep");
  std::vector<double> sums((size_t)16 * (nsteps + 1), 0.0);
  std::vector<int> rebuilt_at;
  gpu.dev_first_forces(&sums[0]);
  for (int s = 1; s <= nsteps; ++s) {
    bool reb = false, have = false;
    double prev[16];
    gpu.dev_md_step(dt, reb, prev, have);
    if (have && s >= 2) std::memcpy(&sums[(size_t)16 * (s - 1)], prev, sizeof prev);
    if (reb) rebuilt_at.push_back(s);
  }
  if (nsteps > 0) gpu.dev_fetch_results(&sums[(size_t)16 * nsteps]);
  std::vector<double> vel;
  gpu.dev_get(c, &vel);
  out.put("sums", sums);
  out.put("rebuilt_at", rebuilt_at);
  out.put("counts", std::vector<int>{c.natms, c.nlast});
  out.put("parts", 'b', c.parts.data(), (int64_t)(c.parts.size() * sizeof(corePart)));
  out.put("ltg", c.ltg);
  out.put("vel", vel);
  return 0;
}

}   // namespace

int main(int argc, char** argv) {
  if (argc != 4) {
    std::fprintf(stderr, "usage: dlpoly_check host|dropin|md <in.bundle> <out.bundle>\n");
    return 2;
  }
  try {
    Bundle in, out;
    in.read(argv[2]);
    const std::string mode = argv[1];
    int rc;
    if (mode == "host") rc = mode_host(in, out);
    else if (mode == "dropin") rc = mode_dropin(in, out);
    else if (mode == "md") rc = mode_md(in, out);
    else { std::fprintf(stderr, "unknown mode %s\n", argv[1]); return 2; }
    out.write(argv[3]);
    return rc;
  } catch (const dlpoly_error& e) {
    std::fprintf(stderr, "%s\n", e.what());
    return e.kode == 0 ? 1 : (e.kode > 255 ? 255 : e.kode);   // like error(): message, then abort with the number
  }
}
