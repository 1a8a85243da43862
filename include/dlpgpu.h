/* =============================================================================
 * dlpgpu.h -- C ABI of libdlpgpu.so: DL_POLY 5.1.0's short-range two-body path on one NVIDIA B200 (sm_100a).
 *
 * This is the drop-in boundary.  The reference has no plugin registry; force providers are direct Fortran calls
 * (precedent: the OpenKIM coupling, source/kim.F90:880-999, called from source/two_body.F90:269-279).  The entry
 * points below are what an ISO_C_BINDING module (fortran/dlp_gpu_binding.F90) binds to replace
 *
 *   source/drivers.F90:675-679     Call link_cell_pairs(...)            -> dlpgpu_link_cell_pairs
 *   source/two_body.F90:339-525    Do i = 1, natms (vdw + ewald real)   \
 *   source/two_body.F90:552-606    Do i = 1, natms (excluded pairs)     -> dlpgpu_two_body_forces
 *   source/neighbours.F90:157-171  max displacement in vnl_check        -> dlpgpu_vnl_check
 *   source/neighbours.F90:337-341  vnl_set_check snapshot               -> dlpgpu_vnl_set_check
 *
 * Conventions
 *   - every function returns 0 on success or a non-zero code; where DL_POLY has a numbered error for the condition
 *     the same number is returned (errors_warnings.F90), so the Fortran wrapper can `Call error(code)`.
 *     dlpgpu_last_error(ctx) gives a message.  No exceptions cross the boundary.
 *   - all pointers are plain host pointers owned by the caller unless the name ends in `_dev`.
 *   - index values inside arrays keep the Fortran convention (1-based local indices, global ids from 1).
 *   - arrays are passed in Fortran memory order; shapes are written in Fortran notation in the comments.
 *   - one context per MPI rank / GPU; not thread-safe; calls are synchronous on return.
 *   - outputs are per-rank partial sums: the caller still performs gsum(buffer) (two_body.F90:729) and
 *     gsum(stress) (drivers.F90:795), so MPI semantics are unchanged.
 * ============================================================================= */
#ifndef DLPGPU_H
#define DLPGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dlpgpu_ctx dlpgpu_ctx;

/* source/particle.F90:14-20  Type corePart (Sequence): 64 bytes */
typedef struct dlpgpu_corepart {
  double xxx, yyy, zzz;
  double fxx, fyy, fzz;
  double chge;
  int32_t pad1, pad2;
} dlpgpu_corepart;

/* error codes (DL_POLY error numbers where one exists) */
#define DLPGPU_OK 0
#define DLPGPU_ERR_CUTOFF_HALF_CELL 95  /* neighbours.F90:409-412  cutoff_extended >= half min cell width  */
#define DLPGPU_ERR_LIST_OVERFLOW 106    /* neighbours.F90:1189-1194 neighbour list array exceeded          */
#define DLPGPU_ERR_HALO_COUNT 138       /* halo.F90:104-106        refreshed halo size mismatch            */
#define DLPGPU_ERR_LINK_CELLS 307       /* neighbours.F90:431      no link cells fit the domain            */
#define DLPGPU_ERR_LOST_ATOMS 58        /* deport_data.F90:3056-3058                                       */
#define DLPGPU_ERR_BUFFER 54            /* deport_data.F90:1871-1876 outgoing transfer buffer too small    */
#define DLPGPU_ERR_CUDA 9001
#define DLPGPU_ERR_ARG 9002
#define DLPGPU_ERR_STATE 9003

/* ---------------------------------------------------------------- lifecycle */
int dlpgpu_create(dlpgpu_ctx** ctx, int device);
int dlpgpu_destroy(dlpgpu_ctx* ctx);
const char* dlpgpu_last_error(const dlpgpu_ctx* ctx);
int dlpgpu_version(void);
/* number of CUDA kernels this context has launched since creation (bench.py's gpu_launches) */
long long dlpgpu_launch_count(const dlpgpu_ctx* ctx);
/* CUDA stream every kernel of the context is launched on (cudaStream_t as void*) */
void* dlpgpu_stream(dlpgpu_ctx* ctx);

/* ---------------------------------------------------------------- setup (once, or when the quantity changes) */
/* domains.F90:35-57  dd = {nx, ny, nz, idx, idy, idz} of this rank (map_domains stays on the host) */
int dlpgpu_set_domain(dlpgpu_ctx* ctx, const int dd[6]);
/* configuration.F90: cell(1:9) row-major lattice vectors, imcon (0,1,2,3 supported: numerics.F90:1511-1600) */
int dlpgpu_set_cell(dlpgpu_ctx* ctx, const double cell[9], int imcon);
/* neighbours.F90:62-84  neigh%cutoff, neigh%padding (cutoff_extended = sum), neigh%pdplnc */
int dlpgpu_set_cutoffs(dlpgpu_ctx* ctx, double rcut, double padding, double pdplnc);
/* vdw.F90 vdw_type:  list(1:ntypes(ntypes+1)/2), ltp(1:max_vdw), tab_potential/tab_force(0:max_grid,1:max_vdw),
 * cutoff, l_force_shift, l_direct, param(1:7,1:max_vdw), afs/bfs(1:max_vdw).  n_vdw<=0 switches vdW off. */
int dlpgpu_set_vdw(dlpgpu_ctx* ctx, int ntypes, const int* vdw_list, int max_vdw, int n_vdw, const int* ltp, int max_grid,
                   const double* tab_potential, const double* tab_force, double rvdw, int force_shift, int direct,
                   const double* param, const double* afs, const double* bfs);
/* ewald_type%alpha, spme_data(0)%scaling = r4pie0/eps (two_body.F90:188), electro%erfc / erfc_deriv interp_tables:
 * arrays hold nsamples+1 doubles with element i == Fortran table(i); element 0 is never read for r >= spacing
 * (numerics.F90:235).  active=0 switches electrostatics off (coul_method off). */
int dlpgpu_set_ewald(dlpgpu_ctx* ctx, int active, double alpha, double scaling, int nsamples, const double* erfc_tab,
                     const double* erfc_deriv_tab, double recip_spacing);
/* The direct-space Coulomb variants of coul_spole.F90, dispatched at two_body.F90:480-514 instead of the Ewald term:
 * kind 1 = coul_cp_forces (1/r), 2 = coul_dddp_forces (distance-dependent dielectric), 3 = coul_fscp_forces (force-shifted),
 * 4 = coul_rfp_forces (reaction field).  scaling = r4pie0/eps; force_shift / energy_shift = electro%force_shift /
 * energy_shift (coul_spole.F90:186-202); reaction_field = electro%reaction_field(0:2) (:407-409).  damp != 0 (kinds 3, 4):
 * the damped forms read electro%erfc / erfc_deriv (generated with alpha = electro%damping), passed like dlpgpu_set_ewald's.
 * Replaces any earlier dlpgpu_set_ewald. */
#define DLPGPU_COUL_CP 1
#define DLPGPU_COUL_DDDP 2
#define DLPGPU_COUL_FSCP 3
#define DLPGPU_COUL_RFP 4
int dlpgpu_set_coulomb(dlpgpu_ctx* ctx, int kind, int damp, double scaling, double force_shift, double energy_shift,
                       const double reaction_field[3], int nsamples, const double* erfc_tab, const double* erfc_deriv_tab,
                       double recip_spacing);

/* ---------------------------------------------------------------- drop-in entry points (host buffers) */
/* neighbours.F90:356-1306.  parts(1:nlast), ltype/ltg/lfrzn(1:nlast), list_excl(0:max_exclude,1:natms) (may be NULL
 * when lbook==0).  list_out(-3:max_list,1:natms) receives the reference-format half list (may be NULL: the list
 * then stays device-resident only).  *ibig = largest row length seen when the list overflows (error 106).
 * Also takes the vnl_set_check snapshot of parts(1:nlast) (halo.F90:315 runs at the same positions). */
int dlpgpu_link_cell_pairs(dlpgpu_ctx* ctx, int natms, int nlast, const dlpgpu_corepart* parts, const int* ltype,
                           const int* ltg, const int* lfrzn, int lbook, int megfrz, int max_exclude,
                           const int* list_excl, int max_list, int* list_out, int* ibig);
/* two_body.F90:339-525 and :552-606 for the device-resident list.  parts(1:nlast): xyz and chge are read, fxx/fyy/fzz
 * of parts(1:natms) are incremented.  out[0..5] = engvdw, virvdw, engcpe_rl, vircpe_rl, engcpe_ex, vircpe_ex;
 * out[6..14] = this rank's contribution to stats%stress(1:9); out[15] = 0. */
int dlpgpu_two_body_forces(dlpgpu_ctx* ctx, int natms, int nlast, dlpgpu_corepart* parts, double out[16]);
/* How the drop-in entry points move corePart records (csrc/hostio.cu).  nthreads = 0 (default; or the environment variable
 * DLPGPU_HOST_THREADS): whole records -- parts(1:nlast) go up as they are from the caller's array, which the library
 * page-locks on first sight, the device adds its forces to that copy and parts(1:natms) come back: 64 nlast + 64 natms bytes
 * per step, no host work.  nthreads >= 1 (at most 64): packed -- that many host threads of the library copy x, y, z (24 of a
 * record's 64 bytes; the charges only where they differ from what the device holds) into page-locked staging buffers and ADD the returned forces (24 bytes per local atom) into
 * parts(1:natms)%f, chunk by chunk under the DMA transfers; the caller's array is left pageable.  Half the PCIe traffic for
 * about twice the traffic through the host's memory: worth it from about a dozen otherwise idle cores per rank.
 * dlpgpu_transfer_bytes reports the bytes the entry points have copied over PCIe since the counters were last reset. */
int dlpgpu_set_host_threads(dlpgpu_ctx* ctx, int nthreads);
int dlpgpu_transfer_bytes(dlpgpu_ctx* ctx, unsigned long long* h2d, unsigned long long* d2h, int reset);
/* packed mode, host clock since the last reset: out[0] = seconds inside the uploads (first record packed to last copy queued),
 * out[1] = seconds a force download waited for its first chunk (the force kernels), out[2] = seconds from there to the last
 * record updated, out[3], out[4] = number of uploads / downloads. */
int dlpgpu_transfer_times(dlpgpu_ctx* ctx, double out[5]);
/* The caller asserts that positions and charges of parts(1:nlast) have not been written since the last
 * dlpgpu_link_cell_pairs: the next dlpgpu_two_body_forces then works on what that call left on the device and skips its own
 * upload.  In whole-record mode (the default, see dlpgpu_set_host_threads) the assertion COVERS parts%f TOO: the device adds
 * its forces to the copy of the records it holds and returns parts(1:natms) = that copy + the pair forces, so contributions
 * that tersoff_forces, three_body_forces or four_body_forces added in between (drivers.F90:675-700) would be overwritten --
 * with any of them active do not make the assertion.  In packed mode the device never sees parts%f (the library's host side
 * adds the returned forces), so only positions and charges matter.  One-shot: consumed by the next dlpgpu_two_body_forces. */
int dlpgpu_parts_unchanged_since_list(dlpgpu_ctx* ctx);
/* SPME reciprocal-space Coulomb term, ewald_spme_forces_coul (ewald_spole.F90:244-477; SURVEY section 8f row 4, beyond the hot path
 * of the north star): B-spline charge spreading (ewald_general.F90:517-576), forward transform, the reference's influence
 * function inside its spherical k cutoff with the stress kernel (ewald_spole.F90:1257-1386), backward transform, force / energy
 * gather with the net force removed (ewald_general.F90:717-869), self interaction (spme.F90:159-231).  dlpgpu_dev_spme_forces serves
 * one domain (mxnode = 1; DLPGPU_ERR_STATE otherwise; several domains: the staged calls below); the 3-D transforms are cuFFT's, loaded on first use
 * (real-to-complex on half the spectrum in orthogonal cells, complex in parallelepiped cells: the reference's + K / 2 convention on
 * the Nyquist planes is not even in m there).
 * set_spme: kdim = ewld%kspace%k_vec_dim (after adjust_kmax), nsplines = bspline%num_splines (3..12); alpha, the Coulomb
 * scaling and the cell come from dlpgpu_set_ewald / dlpgpu_set_cell.
 * dev_spme_forces works on the device-resident atoms (1:natms), ADDS the reciprocal forces to the device force arrays and
 * returns out[0] = engcpe_rc (reciprocal energy + self interaction), out[1] = vircpe_rc, out[2..10] = what the routine adds to
 * stats%stress(1:9), out[11] = the reciprocal energy alone, out[12] = the self interaction.  megatm = atoms in the system. */
int dlpgpu_set_spme(dlpgpu_ctx* ctx, const int kdim[3], int nsplines);
int dlpgpu_dev_spme_forces(dlpgpu_ctx* ctx, int megatm, double out[16]);
/* The same over several domains, with a REPLICATED grid instead of the reference's distributed one (exchange_grid + the parallel
 * DaFT transform, ewald_spole.F90:336-420, parallel_fft.F90): every rank spreads its own atoms onto a grid of the whole cell
 * (spread; grid_dev = K1 K2 K3 doubles of DEVICE memory owned by the caller, z fastest, zeroed by the call), the host side sums
 * the ranks' grids with the collective it has (NCCL all-reduce over NVLink, dl-poly_b200/dd.py), every rank transforms the
 * whole grid and gathers the forces of its own atoms (solve_gather; ftot_local = the rank's raw net force), and after the sum of
 * those three doubles over the ranks the net force is removed and the forces are ADDED to the device force arrays (finish; out
 * as for dlpgpu_dev_spme_forces, this rank's share: the ranks' outs add up under gsum -- the k-space stress, formed from the whole
 * grid on every rank, is reported as 1 / nranks of itself). */
int dlpgpu_dev_spme_spread(dlpgpu_ctx* ctx, double* grid_dev);
int dlpgpu_dev_spme_solve_gather(dlpgpu_ctx* ctx, double* grid_dev, double ftot_local[3]);
int dlpgpu_dev_spme_finish(dlpgpu_ctx* ctx, int megatm, const double ftot_global[3], int nranks, double out[16]);
/* the same with the caller's corePart array (drop-in for the call at two_body.F90:298-302 when comm%mxnode == 1): parts(1:natms) are
 * uploaded, the reciprocal forces come back and are ADDED to parts(1:natms)%f.  It reuses the device atom arrays, so the
 * neighbour list held by the context is invalidated: call it BEFORE dlpgpu_link_cell_pairs of the step, or from a context of its own. */
int dlpgpu_spme_forces(dlpgpu_ctx* ctx, int natms, dlpgpu_corepart* parts, int megatm, double out[16]);
/* stats%collect_pp (statistics.F90:227, set by the per-particle / heat-flux options): while on, two_body_forces (drop-in and
 * dev_) also books, for every pair, half of its energy and half of its stress tensor r (x) f on each LOCAL partner, following
 * the reference path by path -- vdw_forces_direct (vdw.F90:1707, :1741-1755: the pair energy for every pair),
 * vdw_forces_tab (:1905, :1987-2001: the energy only where this rank owns it, i.e. a local partner or idi < ltg(jatm)),
 * ewald_real_forces_coul (ewald_spole.F90:155, :205-215: every pair); ewald_excl_forces books nothing per particle.  The general pair
 * kernel runs (not the fast one).  Needs the half-list force mode; not available with the coul_spole.F90 variants
 * (DLPGPU_ERR_STATE).  dlpgpu_get_pp ADDS the sums of the last force call into pp_energy(1:natms) and pp_stress(1:9,1:natms)
 * (column-major: nine per atom, calculate_stress order, statistics.F90:2616-2625). */
int dlpgpu_set_collect_pp(dlpgpu_ctx* ctx, int on);
int dlpgpu_get_pp(dlpgpu_ctx* ctx, int natms, double* pp_energy, double* pp_stress);
/* rdfs.F90:146-212 rdf_collect, :880-946 rdf_excl_collect and :948-1018 rdf_frzn_collect for the atoms and list of the last
 * dlpgpu_two_body_forces / dlpgpu_dev_two_body_forces call (two_body_forces calls them at two_body.F90:523, :581, :649):
 * rdf_list(1:ntypes(ntypes+1)/2) = rdf%list (0 or > n_pairs: pair not collected), rdf = rdf%rdf(1:max_grid,1:n_pairs)
 * column-major on the host, INCREMENTED by this rank's counts.  Frozen-frozen pairs never reach the force loops
 * (neighbours.F90:1198-1225) but are kept in device rows of their own for this call when megfrz > 1. */
int dlpgpu_rdf_collect(dlpgpu_ctx* ctx, int ntypes, const int* rdf_list, int n_pairs, int max_grid, double* rdf);
/* neighbours.F90:157-171: tol = max_i |r_i - r_bg,i| (minimum image) over parts(1:natms); caller does gmax + test */
int dlpgpu_vnl_check(dlpgpu_ctx* ctx, int natms, const dlpgpu_corepart* parts, double* tol);
int dlpgpu_vnl_set_check(dlpgpu_ctx* ctx, int nlast, const dlpgpu_corepart* parts);

/* ---------------------------------------------------------------- native device-resident mode
 * The same kernels, driven without a Fortran host (bench.py, multi-GPU engine): atoms live on the device in
 * DL_POLY's local order 1..natms | natms+1..nlast; halo build / refresh / migration run as pack / unpack kernels
 * around a transport the caller owns (torch.distributed NCCL send/recv on `*_dev` buffers). */
int dlpgpu_dev_set_sites(dlpgpu_ctx* ctx, int nsites, const int* type_site, const double* charge_site,
                         const int* freeze_site, const double* weight_site);
/* exclusion rows by GLOBAL id: excl(0:max_exclude, 1:megatm), row 0 = count, ids ascending (build_excl.F90:1181-1184) */
int dlpgpu_dev_set_excl(dlpgpu_ctx* ctx, int megatm, int max_exclude, const int* excl_by_gid);
/* halo.F90:219-233 reduced-space width of the negative-direction halo requested by SPME (0 = link-cell width) */
int dlpgpu_dev_set_halo_width(dlpgpu_ctx* ctx, const double ecw[3]);
int dlpgpu_dev_set_list_capacity(dlpgpu_ctx* ctx, int max_list, int megfrz);
int dlpgpu_dev_load_atoms(dlpgpu_ctx* ctx, int natms, const double* xyz, const double* vel, const int* ltg,
                          const int* lsite, int capacity_atoms);
int dlpgpu_dev_counts(dlpgpu_ctx* ctx, int* natms, int* nlast);
int dlpgpu_dev_zero_forces(dlpgpu_ctx* ctx);
/* nve.F90:163-173 (stage 1) and :198-217 (stage 2) velocity-Verlet half steps -- the trajectory driver of bench.py */
int dlpgpu_dev_vv(dlpgpu_ctx* ctx, int stage, double dt);
int dlpgpu_dev_vnl_check(dlpgpu_ctx* ctx, double* tol);
/* set_halo_particles (halo.F90:153-355) split around the transport: begin (tag ixyz) -> for mdir in -1,1,-2,2,-3,3:
 * pack (export_atomic_data select + pack, deport_data.F90:1810-1866), [exchange], unpack (:1921-1943) -> end
 * (types/charges from lsite, vnl_set_check).  *count = atoms packed.  DLPGPU_HALO_WIDTH doubles per atom: the
 * reference's six (x,y,z,ltg,lsite,ixyz) followed by origin rank, origin local index and the periodic wraps applied so
 * far, which is what dlpgpu_dev_refresh_pull replays. */
#define DLPGPU_HALO_WIDTH 9
int dlpgpu_dev_halo_begin(dlpgpu_ctx* ctx);
int dlpgpu_dev_halo_pack(dlpgpu_ctx* ctx, int mdir, double* sendbuf_dev, int capacity_atoms, int* count);
int dlpgpu_dev_halo_unpack(dlpgpu_ctx* ctx, int mdir, const double* recvbuf_dev, int count);
int dlpgpu_dev_halo_end(dlpgpu_ctx* ctx);
/* refresh_halo_positions (halo.F90:47-113): same atoms in the same order, 3 doubles per atom */
int dlpgpu_dev_refresh_pack(dlpgpu_ctx* ctx, int mdir, double* sendbuf_dev, int* count);
int dlpgpu_dev_refresh_unpack(dlpgpu_ctx* ctx, int mdir, const double* recvbuf_dev, int count);
/* refresh_halo_positions (halo.F90:47-113) as ONE kernel over NVLink peer memory instead of six staged messages: every
 * rank keeps a double-buffered, CUDA-IPC exported copy of its local coordinates; p2p_init allocates it and returns the two
 * IPC handles (a DLPGPU_P2P_BLOB), p2p_open maps the buffers of all ranks (blobs gathered rank-major).  Per
 * step: dlpgpu_dev_publish after the positions moved, a collective on the same streams (the gmax of vnl_check), then
 * dlpgpu_dev_refresh_pull fills the whole halo from the owners' buffers, replaying the periodic shifts in stage order
 * (same bits as the staged exchange).  With nranks == 1 no IPC is involved. */
#define DLPGPU_P2P_BLOB 192   /* two CUDA-IPC handles (2 x 64 B) + {pid, pointer 0, pointer 1, device}: ranks that are threads of one
                               * process reach each other's buffers by the plain pointers, processes by the IPC handles */
#define DLPGPU_XCHG_BLOB 128  /* one CUDA-IPC handle + {pid, pointer, device} */
int dlpgpu_dev_p2p_init(dlpgpu_ctx* ctx, int rank, int nranks, int capacity_atoms, unsigned char handles_out[DLPGPU_P2P_BLOB]);
int dlpgpu_dev_p2p_open(dlpgpu_ctx* ctx, const unsigned char* all_handles);
int dlpgpu_dev_publish(dlpgpu_ctx* ctx);
int dlpgpu_dev_refresh_pull(dlpgpu_ctx* ctx);
/* relocate_particles + set_halo_particles (+ vnl_set_check) and the gmax of vnl_check as device-side exchanges over NVLink
 * peer memory: no NCCL, no host synchronisation between the twelve dependent stages of a rebuild.  xchg_init allocates this
 * rank's CUDA-IPC exported region (gmax mailboxes + one fixed-capacity receive buffer and {sequence, count} header per
 * stage; capacities in atoms per stage, identical on every rank) and returns its handle (a DLPGPU_XCHG_BLOB); xchg_open maps the regions
 * of all ranks (handles gathered rank-major).  xchg_rebuild enqueues everything and synchronises ONCE at the end (natms /
 * nlast live on the device meanwhile); neigh = map(1:6) of domains.F90:206-211 (0-based ranks; the rank itself where the
 * decomposition has one domain in that direction).  seq must be a fresh, rank-uniform sequence number per call (ranks run
 * in lock-step: every rank calls xchg_rebuild / xchg_gmax in the same order).  Errors: 43 / 54 (a stage exceeded its
 * capacity), 58 (lost atoms), DLPGPU_ERR_STATE on a peer time-out.  With nranks == 1 no IPC is involved. */
int dlpgpu_dev_xchg_init(dlpgpu_ctx* ctx, int rank, int nranks, int cap_reloc_atoms, int cap_halo_atoms, unsigned char handle_out[DLPGPU_XCHG_BLOB]);
int dlpgpu_dev_xchg_open(dlpgpu_ctx* ctx, const unsigned char* all_handles);
/* Diagnostic: how the six migration stages of xchg_rebuild find their atoms.  0 (default): the atoms with a relocation tag are
 * compacted once and every stage is one single-block kernel over that list; 1: every stage scans all atoms (count, pack,
 * restack, receive kernels).  Same buffers, order and counts either way; must be the same on every rank. */
int dlpgpu_dev_xchg_set_migration(dlpgpu_ctx* ctx, int scan_all);
/* device time (CUDA events on the context's stream, ms) of the kernels of the last dlpgpu_dev_xchg_rebuild, peer waits included */
int dlpgpu_dev_xchg_last_ms(dlpgpu_ctx* ctx, double* ms);
/* how long a receive / gmax kernel waits for its peer before the call fails with DLPGPU_ERR_STATE (default 60 s; per device) */
int dlpgpu_dev_xchg_set_timeout(dlpgpu_ctx* ctx, double seconds);
int dlpgpu_dev_xchg_rebuild(dlpgpu_ctx* ctx, const int neigh[6], unsigned long long seq, int* natms, int* nlast);
int dlpgpu_dev_xchg_gmax(dlpgpu_ctx* ctx, unsigned long long seq, double* tol);
/* One MD step of the native driver around the path (md_vv, drivers.F90:1910-2290), enqueued from C: dev_vv(1) + publish,
 * xchg_gmax (the only host synchronisation), then xchg_rebuild + link_cell_pairs when tol >= half_minus * padding
 * (neighbours.F90:182; *rebuilt = 1, *list_ms = time of the list build) or the one-kernel halo refresh, two_body_forces
 * without waiting for its sums, dev_vv(2).  out_prev / *have_prev: the sums of the previous step's force call, if one was
 * pending, ALREADY SUMMED OVER ALL RANKS: the 16 partial sums ride on the gmax message and every rank adds them in rank
 * order (the gsum of two_body.F90:729 and drivers.F90:795, deterministic).  The last step's sums come from
 * dlpgpu_dev_fetch_results as this rank's partial sums.  gseq / rseq: as for xchg_gmax / xchg_rebuild. */
int dlpgpu_dev_md_step(dlpgpu_ctx* ctx, const int neigh[6], double dt, unsigned long long gseq, unsigned long long rseq,
                       int* rebuilt, double out_prev[16], int* have_prev, double* list_ms);
/* dlpgpu_dev_md_step rebuilds at least every `every` steps in addition to the padding-driven test of neighbours.F90:182
 * (0 = off, the default; must be the same on every rank).  For trajectories that are held still on purpose (measurement). */
int dlpgpu_dev_set_rebuild_every(dlpgpu_ctx* ctx, int every);
/* atoms sent / received in each of the six stages of the last halo build (order -x,+x,-y,+y,-z,+z) */
int dlpgpu_dev_halo_stage_counts(dlpgpu_ctx* ctx, int sent[6], int received[6]);
/* single-domain shortcuts (mxnode == 1: the neighbour is the rank itself, deport_data.F90:1884-1886) */
int dlpgpu_dev_halo_serial(dlpgpu_ctx* ctx);
int dlpgpu_dev_refresh_serial(dlpgpu_ctx* ctx);
/* relocate_particles (deport_data.F90:2870-3202): serial = pbcshift; DD = tag, then per direction pack (12 doubles per
 * atom: x,v,f,ltg,lsite,ixyz) / unpack, then end (lost-atom check is the caller's gsum) */
int dlpgpu_dev_relocate_serial(dlpgpu_ctx* ctx);
int dlpgpu_dev_relocate_begin(dlpgpu_ctx* ctx);
int dlpgpu_dev_relocate_pack(dlpgpu_ctx* ctx, int mdir, double* sendbuf_dev, int capacity_atoms, int* count);
int dlpgpu_dev_relocate_unpack(dlpgpu_ctx* ctx, int mdir, const double* recvbuf_dev, int count);
int dlpgpu_dev_relocate_end(dlpgpu_ctx* ctx, int* natms_now);
/* link_cell_pairs / two_body_forces on the resident atoms.  want_ref_list!=0 additionally materialises the
 * reference-format half list on the device (for dlpgpu_dev_get_list). */
int dlpgpu_dev_link_cell_pairs(dlpgpu_ctx* ctx, int want_ref_list, int* ibig);
int dlpgpu_dev_two_body_forces(dlpgpu_ctx* ctx, int zero_forces, double out[16]);
/* out == NULL above enqueues the call without waiting; the sums are collected later (after any later synchronisation of the
 * context's stream they are already there) -- one host round trip per MD step instead of two */
int dlpgpu_dev_fetch_results(dlpgpu_ctx* ctx, double out[16]);
/* pairs in the reference's half list for this domain (local-local once + local-halo), counted from the device list */
int dlpgpu_dev_list_pairs(dlpgpu_ctx* ctx, long long* pairs);
/* read-back (tests, diagnostics) */
int dlpgpu_dev_get_parts(dlpgpu_ctx* ctx, dlpgpu_corepart* parts_out, int n);
int dlpgpu_dev_get_ints(dlpgpu_ctx* ctx, int n, int* ltg, int* lsite, int* ltype, int* lfrzn, int* ixyz);
int dlpgpu_dev_get_vel(dlpgpu_ctx* ctx, int n, double* vel3);
int dlpgpu_dev_get_list(dlpgpu_ctx* ctx, int natms, int max_list, int* list_out);
/* link-cell diagnostics of the last build: info = {nlx,nly,nlz,nlp,ncells,nsbcll}; arrays may be NULL */
int dlpgpu_dev_get_cells(dlpgpu_ctx* ctx, int info[6], int* which_cell, int* at_list, int* lct_start);
/* the device-internal full neighbour rows of local atom i (1-based local index): partners as 1-based local
 * indices; n_main/n_excl counts.  Test hook for "full list == symmetrised reference list". */
int dlpgpu_dev_get_full_row(dlpgpu_ctx* ctx, int i, int* n_main, int* main_out, int* n_excl, int* excl_out, int cap);

/* ---------------------------------------------------------------- measurement helpers */
/* DFMA micro-benchmark: sustained fp64 FMA throughput of this GPU in TFLOP/s (2 flop per FMA) */
int dlpgpu_fp64_peak(dlpgpu_ctx* ctx, double seconds, double* tflops);
/* device time (ms, CUDA events on the context's stream) of the last list build and the last force evaluation,
 * and of their dominant kernels: t[0]=list total, t[1]=force total, t[2]=pair-force kernel, t[3]=full-list kernel */
int dlpgpu_last_timings(dlpgpu_ctx* ctx, double t[4]);
/* Two notes on the tabulated fast path (k_pair_v2): a pair closer than ONE grid step of the interpolation tables (r < rcut / (mxgrid - 4),
 * ~0.01 A) is evaluated in the first full interval (l clamped to 1) where the reference would read its Huge(1.0) / r-scaled entry 0;
 * no physical configuration reaches it, and the general kernel (dlpgpu_set_pair_kernel(ctx, 1)) follows the reference there.
 * And: in full-list mode the device rows hold up to 2 max_list partners (each local-local pair sits in both rows). */
/* 1 (default, common.cuh force_mode): half list + fp64 RED atomics (Newton's third law) -- the reference's own pair count.
 * 0: full list without atomics (every local-local pair evaluated from both ends; bitwise reproducible forces). */
int dlpgpu_set_force_mode(dlpgpu_ctx* ctx, int mode);
/* Diagnostic: which pair kernel two_body_forces uses.  0 (default): automatic -- the fast tabulated kernel (k_pair_v2) where it
 * applies, the general kernel otherwise.  1: always the general kernel (k_pair_forces: the reference's operation order
 * statement by statement, vdw.F90:1790-2024 / ewald_spole.F90:58-242).  The parity tests hold both against the oracle and
 * against each other; results agree within the north-star bars whichever runs. */
int dlpgpu_set_pair_kernel(dlpgpu_ctx* ctx, int which);
/* which kernel the last two_body_forces call ran (1 general, 2 k_pair_v2; 0 none); *packed_table_error is reserved (0) */
int dlpgpu_pair_kernel_used(dlpgpu_ctx* ctx, int* which, double* packed_table_error);
/* Diagnostic: how the warp-per-cell kernel builds the device half list.  0 (default): the x-runs of candidate cells are trimmed
 * against the bounding boxes of the cells' atoms; 1: untrimmed runs; 2: trimmed runs plus a per-candidate prune with compaction
 * into a shared-memory ring.  Same rows either way (members and order); the parity tests compare them. */
int dlpgpu_set_list_kernel(dlpgpu_ctx* ctx, int which);

#ifdef __cplusplus
}
#endif
#endif /* DLPGPU_H */
