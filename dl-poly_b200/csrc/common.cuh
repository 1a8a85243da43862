// common.cuh -- context, device buffers and launch helpers shared by the libdlpgpu translation units.
#pragma once
#include <cuda_runtime.h>

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/dlpgpu.h"

#define DLP_WARP 32
#define DLP_FULL 0xffffffffu

// entries of the device-internal lists: sorted slot of the partner in the low 24 bits, then 6 bits "vdW potential of this
// type pair + 1" (0: no vdW interaction, vdw.F90:1875-1892 resolved at build time), then two flags
#define DLP_J_MASK 0x00ffffffu
#define DLP_K_SHIFT 24
#define DLP_K_MASK 0x3fu
#define DLP_MAX_SLOTS (1 << 24)
#define DLP_MAX_KCODE 62
#define DLP_F_HALO 0x40000000u   // partner is a halo atom (one-sided pair)
#define DLP_F_ECNT 0x80000000u   // halo partner whose pair energy is counted here (idi < ltg(jatm))

template <class T>
struct DBuf {
  T* p = nullptr;
  size_t cap = 0;
  // grow to at least n elements; keep=true preserves the first `keep_n` elements.  Stream-ordered allocation (cudaMallocAsync /
  // cudaFreeAsync on the context's stream): growing a buffer never synchronises the DEVICE, which matters when several ranks
  // drive one GPU from threads of one process -- a peer's receive kernel may be spinning on this rank's next message, and a
  // device-wide implicit synchronisation (cudaFree) would wait for it for ever.
  cudaError_t ensure(size_t n, cudaStream_t s = 0, bool keep = false, size_t keep_n = 0) {
    if (n <= cap) return cudaSuccess;
    size_t ncap = n + n / 8 + 64;
    T* q = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&q, ncap * sizeof(T), s);
    if (e != cudaSuccess) return e;
    if (keep && p && keep_n) {
      e = cudaMemcpyAsync(q, p, std::min(keep_n, cap) * sizeof(T), cudaMemcpyDeviceToDevice, s);
      if (e != cudaSuccess) return e;
    }
    if (p) cudaFreeAsync(p, s);
    p = q;
    cap = ncap;
    return cudaSuccess;
  }
  // stream-ordered like ensure(): cudaFree would wait for every stream of the device, including a peer rank's spinning receive
  void release(cudaStream_t s = 0) {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
    cap = 0;
  }
};

// link-cell geometry of one build (neighbours.F90:401-601), filled on the host with the reference's arithmetic
struct LCGeom {
  double rcell[9];
  double xdc, ydc, zdc;
  double rcsq;
  int jx, jy, jz;
  int nlx, nly, nlz, nlp;
  int sx, sy, sz;     // cells per direction incl. halo layers: nlx + 2 nlp ...
  int ncells;         // sx*sy*sz ; cell ids 1..ncells, 0 = residual halo
  int idx, idy, idz;
  int nsbcll;
  int nir_r2;         // (nlp-1)^2: offsets with ix^2+iy^2+iz^2 < nir_r2 skip the distance test (neighbours.F90:537)
};

#define DLP_HALO_W 9   // doubles per atom on a halo build: the reference's six (x,y,z,ltg,lsite,ixyz) + origin rank, index, wraps

struct HaloStage {    // one export_atomic_data direction as recorded at halo build, replayed by the refresh
  int count = 0;      // atoms sent (== received from the opposite neighbour in a serial run)
  int recv_off = 0;   // first local index (0-based) the received atoms were appended at
  int recv_count = 0;
  double shift[3] = {0, 0, 0};
  bool lwrap = false;
  DBuf<int> idx;      // source indices (0-based), ascending
};

struct dlpgpu_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  long long launches = 0;
  int sm_count = 148;

  // ---- setup
  int nx = 1, ny = 1, nz = 1, idx = 0, idy = 0, idz = 0;
  double cell[9] = {0};
  int imcon = 1;
  double rcut = 0, padding = 0, rx = 0, pdplnc = 50.0;
  double ecw[3] = {0, 0, 0};
  int force_mode = 1;     // 1: half list + fp64 RED (Newton 3), 0: full list without atomics

  // vdw
  bool vdw_on = false, vdw_fshift = false, vdw_direct = false;
  int ntypes = 0, n_vdw = 0, max_vdw = 0, max_grid = 0;
  double rvdw = 0, vdw_rdr = 0, thr_vdw = 0;   // thr_vdw: rsq < thr_vdw  <=>  Sqrt(rsq) < rvdw
  DBuf<int> pair_k;        // [ntypes*ntypes] -> potential index k (0-based) or -1 (no interaction)
  DBuf<int> ltp;           // [max_vdw]
  DBuf<double2> vdw_tab;   // [max_vdw][max_grid+1] {tab_force, tab_potential}
  DBuf<double> vdw_par;    // [max_vdw][10] param(1:7), afs, bfs, pad
  // ewald
  bool ew_on = false;
  int coul_kind = 0;        // DLPGPU_COUL_*: direct-space Coulomb variant instead of Ewald (coul_spole.F90)
  bool coul_damp = false;
  double coul_fs = 0, coul_es = 0, coul_rf[3] = {0, 0, 0};
  double alpha = 0, scaling = 0, ew_rdr = 0, thr_coul = 0;
  int ew_n = 0;
  DBuf<double2> ew_tab;    // [ew_n+1] {erfc_deriv, erfc}
  // quadratic-form tables of the pair kernel: [max_vdw][max_grid+1] then [ew_n+1] entries {g_f, g_e, h_f, h_e}
  DBuf<double> tab4;
  size_t tab4_entries = 0;
  // k_pair_v2 layout: [g units: table 0 = Ewald (or zeros), table k = vdW potential k, then 2 zero entries][h units likewise]
  DBuf<double> tab2;
  int tab2_ne = 0, tab2_ts = 0, tab2_zero = 0;
  DBuf<double> tab2s;                 // copy of tab2's g units with the 8-bit completion of the fp32 energy h parked in g_energy's low bits
  DBuf<float> tab2h;                  // float4 second differences {vdW force, vdW energy, Ewald force, Ewald energy} per (potential, l)
  cudaTextureObject_t tab2h_tex = 0;
  void* hostio = nullptr;             // hostio.cu: worker pool, page-locked staging buffers and byte counters of the drop-in entry points
  int* dc_pinned = nullptr;           // page-locked image of the exchange's device counts (dlpgpu_dev_xchg_init allocates it)
  std::vector<int> st_uploaded;       // the stencil arrays the device currently holds (dlp_build_lists uploads on change)
  DBuf<int> row_perm;                 // rows of each window of DLP_ROW_WIN atoms ordered by length (longest first, -1 = no row): k_pair_v2 deals them to its warps
  DBuf<float> cell_box;               // {lo, hi} float4 per link cell: bounding boxes of the cells' atoms (k_cell_boxes)
  DBuf<unsigned> fnbr;                // frozen-frozen partners per row (kept for rdf_frzn_collect only), pitch fpitch
  DBuf<int> nfnbr;
  int fpitch = 0;
  bool frz_rows_valid = false;
  // SPME reciprocal space (spme.cu)
  int spme_k[3] = {0, 0, 0}, spme_n = 0, spme_kmax = 0, spme_plan = 0;
  bool spme_plan_valid = false;
  int spme_plan_d2z = 0, spme_plan_z2d = 0;   // real-to-complex pair of plans (orthogonal cells), created on first use
  bool spme_r2c_valid = false;
  DBuf<double> spme_grid, spme_rgrid, spme_norm2, spme_fraw, spme_tot;
  bool collect_pp = false;            // dlpgpu_set_collect_pp: stats%collect_pp
  int pp_natms = -1;
  DBuf<double> pp_pos, pp_neg, pp_energy, pp_stress;   // per-particle sums of the last force call (see k_pair_forces<.., PP>)
  int last_pair_kernel = 0;           // 1 k_pair_forces, 2 k_pair_v2 (dlpgpu_pair_kernel_used)
  int list_one_atom_per_pass = 0;     // dlpgpu_set_list_kernel: 1 = k_list_cell<1> instead of k_list_cell8 (diagnostic)
  cudaTextureObject_t tab2_tex = 0;   // the same buffer as 16-byte texels: table reads through the texture pipe (see pair2)
  int ew_off = 0;
  bool tab4_valid = false;
  std::vector<double> h_vdw_f, h_vdw_e, h_ew_d, h_ew_e;   // host copies the tab4 build reads
  bool no_fast = false;    // dlpgpu_set_pair_kernel: always use the general pair kernel

  // sites (native mode)
  int nsites = 0;
  DBuf<int> type_site, freeze_site;
  DBuf<double> charge_site, weight_site;
  // exclusions
  int lbook = 0, megfrz = 0, max_exclude = 0;
  int excl_by_gid = 0;     // rows indexed by global id (native) or by local index (drop-in)
  DBuf<int> excl;          // [(max_exclude+1) * rows]

  // ---- atoms in DL_POLY local order (0-based here)
  int natms = 0, nlast = 0, capacity = 0;
  DBuf<double4> posq;      // x,y,z,chge
  DBuf<double> fx, fy, fz;
  DBuf<double> fsx, fsy, fsz;   // force_mode 1: fsx = row sums [3][natms], fsy = blocked j-side accumulators, fsz unused
  DBuf<double> vx, vy, vz;
  DBuf<int> ltg, lsite, ltype, lfrzn, ixyz;
  // origin of every resident atom, carried through the halo build: owning rank, local index there, and the periodic
  // wraps applied on the way ((u+1) + 3(v+1) + 9(w+1), u,v,w in {-1,0,1}) -- what the one-kernel halo refresh replays
  DBuf<int> org_rank, org_idx, org_wrap;
  // peer-visible copies of the local coordinates (double-buffered by step parity), CUDA-IPC mapped on every rank
  double4* pub[2] = {nullptr, nullptr};
  int pub_cap = 0, pub_parity = 0;
  int p2p_rank = 0, p2p_nranks = 0;
  bool p2p_ready = false, pub_valid = false;
  bool tol_fresh = false, pub_fresh = false;   // left behind by the fused velocity-Verlet stage 1 (dlp_vv1_fused)
  std::vector<double4*> peer_pub;   // [2 * nranks]
  std::vector<char> peer_pub_local, peer_xr_local;   // [nranks] 1: the peer lives in this process (plain pointer, nothing to close)
  DBuf<unsigned long long> peer_pub_dev;   // the same table on the device
  // fused device-side exchange (dlpgpu_dev_xchg_*): one CUDA-IPC exported region with the gmax mailboxes and the per-stage
  // receive buffers of migration and halo build, the peers' regions, and the device-resident atom counts
  char* xr = nullptr;
  int xr_rank = 0, xr_nranks = 0, xr_cap_r = 0, xr_cap_h = 0;
  bool xr_ready = false;
  std::vector<char*> peer_xr;
  DBuf<unsigned long long> peer_xr_dev;
  DBuf<int> dcnt;
  DBuf<unsigned long long> gmax_out;   // [1 + 16]: gmax bits, then the gsum of the previous force call's 16 sums
  unsigned long long* gm_pinned = nullptr;       // mapped pinned memory k_x_gmax reports into (the host polls its flag)
  unsigned long long* gm_pinned_dev = nullptr;
  double gsum_prev[16] = {0};
  int rebuild_every = 0, steps_since_rebuild = 0;   // dlpgpu_dev_set_rebuild_every: forced cadence on top of the padding test
  DBuf<double> xbg, ybg, zbg;
  bool have_bg = false;

  // ---- link cells / lists of the last build
  LCGeom g{};
  bool list_valid = false;
  int list_natms = 0, list_nlast = 0;
  int max_list = 0;
  DBuf<int> which_cell, at_list, at_tmp, lct_count, lct_start, lct_fill;
  DBuf<int> cell_s;        // cell id per sorted slot
  DBuf<int> loc_slot;      // [natms] sorted slot of the t-th local atom in sorted order
  DBuf<int> flag, scan_out, scan_tmp;
  DBuf<double4> posq_s;    // sorted copy, refreshed every force call
  DBuf<int> type_s, gid_s, frz_s;
  DBuf<int2> info_s;       // per sorted slot {global id, type | frozen<<16 | halo<<17}
  DBuf<int> st_rows;       // semi-ball rows {dy,dz,half x-extent} of the warp-per-cell list kernel
  DBuf<int> st_nix, st_niy, st_niz, st_nir;   // semi-ball stencil in reference order
  DBuf<int> st_xb;         // [(2nlp+1)^2] half x-extent per (dy,dz) row of the full ball, -1 = row absent
  std::vector<int> h_nix, h_niy, h_niz, h_nir, h_xb;
  // reference-format half list (optional)
  DBuf<int> ref_list;      // (-3:max_list, 1:natms)
  bool ref_valid = false;
  // device-internal lists, row t = t-th local atom in sorted order
  int pitch = 0, xpitch = 0;
  DBuf<unsigned> nbr;      // [natms][pitch]
  DBuf<int> nnbr;          // [natms]
  DBuf<unsigned> xnbr;     // [natms][xpitch] excluded partners
  DBuf<int> nxnbr;
  DBuf<unsigned> hnbr;     // half list (force_mode 1): [natms][pitch]
  DBuf<int> nhnbr;
  DBuf<double> xfer;       // internal exchange buffer of the serial halo / refresh
  DBuf<int> hole_pos;
  cudaEvent_t ev_x[2] = {nullptr, nullptr};   // around the kernels of the last dlpgpu_dev_xchg_rebuild
  double t_xchg = 0.0;
  DBuf<int> movers;                   // ascending indices of the atoms with a non-zero relocation tag (k_x_reloc_stage)
  int xchg_scan_migration = 0;        // dlpgpu_dev_xchg_set_migration: 1 = the scanning migration stages (diagnostic)
  DBuf<int> status;        // [8] device status words: 0 overflow flag, 1 ibig, 2 lost atoms, 3 spare
  DBuf<unsigned long long> cnt64;   // [4] 0: entries of the device list
  long long list_entries = 0;
  // halo replay
  HaloStage stage[6];
  bool halo_valid = false;
  // reductions
  DBuf<int> rdf_list;                 // rdf%list on the device (dlpgpu_rdf_collect)
  DBuf<unsigned long long> rdf_hist;
  DBuf<double> partial;    // [blocks][16]
  DBuf<double> out_dev;    // [16]
  DBuf<unsigned long long> tol_bits;
  // staging for the host-buffer entry points
  DBuf<dlpgpu_corepart> parts_dev;
  int parts_resident = 0;          // records of the caller's parts array whose coordinates the device holds (set by link_cell_pairs)
  bool parts_current = false;      // dlpgpu_parts_unchanged_since_list: the next two_body_forces may skip its upload
  // timings
  cudaEvent_t ev[8] = {nullptr};
  cudaEvent_t ev_res = nullptr;     // behind the asynchronous result copy of dlpgpu_dev_two_body_forces(out = NULL)
  double* out_pinned = nullptr;
  bool res_pending = false;
  double t_list = 0, t_force = 0, t_pair = 0, t_full = 0;
};

int dlp_fail(dlpgpu_ctx* ctx, int code, const char* fmt, ...);

#define CK(expr)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return dlp_fail(ctx, DLPGPU_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
  } while (0)
#define CKRC(expr)          \
  do {                      \
    int _rc = (expr);       \
    if (_rc != 0) return _rc; \
  } while (0)
#define DLP_ROW_WIN 64   // rows per block pass of k_pair_v2 (512 threads / 8 lanes per row)
#define LAUNCH(ctx, kern, grid, block, smem, ...)                    \
  do {                                                               \
    kern<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);   \
    (ctx)->launches++;                                               \
  } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// every translation unit loads its kernels up front (see dlp_preload_halo)
int dlp_preload_halo();
int dlp_preload_hostio();
// hostio.cu: the caller's corePart records <-> device arrays (packed, chunked, overlapped with the DMA engine)
int dlp_upload_parts(dlpgpu_ctx* ctx, int n, const dlpgpu_corepart* parts);
int dlp_upload_ints(dlpgpu_ctx* ctx, int n, const int* ltype, const int* ltg, const int* lfrzn);
int dlp_download_add_forces(dlpgpu_ctx* ctx, int natms, dlpgpu_corepart* parts);
void dlp_hostio_ints_stale(dlpgpu_ctx* ctx);
void dlp_hostio_release(dlpgpu_ctx* ctx);
int dlp_preload_cells();
int dlp_preload_forces();
int dlp_preload_ctx();
int dlp_preload_spme();
void dlp_spme_release(dlpgpu_ctx* ctx);
// util.cu
int dlp_exclusive_scan(dlpgpu_ctx* ctx, const int* in_dev, int* out_dev, int n, int* total_host /*nullable*/);
int dlp_ensure_atoms(dlpgpu_ctx* ctx, int n);
// cells.cu
int dlp_build_lists(dlpgpu_ctx* ctx, int want_ref_list, int* ibig);
int dlp_gather_sorted(dlpgpu_ctx* ctx);
// forces.cu
int dlp_two_body(dlpgpu_ctx* ctx, int zero_forces, double out[16]);
int dlp_build_tab4(dlpgpu_ctx* ctx);
// halo.cu
int dlp_vnl_set_check(dlpgpu_ctx* ctx);
extern "C" int dlp_vv1_fused(dlpgpu_ctx* ctx, double dt);
int dlp_vnl_check(dlpgpu_ctx* ctx, double* tol);
