// ctx.cu -- context lifecycle, setup entry points, host-buffer (drop-in) wrappers and read-back of libdlpgpu.
#include <cstdarg>
#include <cstdlib>

#include "common.cuh"

int dlp_fail(dlpgpu_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx) ctx->err = buf;
  return code;
}

// ------------------------------------------------------------------ exclusive scan (int32), hand-written
// 3-phase: per-block scan of 2048 items -> scan of block sums (recursive) -> add offsets.
namespace {
constexpr int SCAN_T = 512, SCAN_I = 4, SCAN_B = SCAN_T * SCAN_I;

__global__ void k_scan_block(const int* __restrict__ in, int* __restrict__ out, int* __restrict__ sums, int n) {
  __shared__ int wsum[SCAN_T / 32];
  int base = blockIdx.x * SCAN_B + threadIdx.x * SCAN_I;
  int v[SCAN_I], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_I; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int inc = s;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    int t = __shfl_up_sync(DLP_FULL, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) wsum[w] = inc;
  __syncthreads();
  if (w == 0) {
    int x = (lane < SCAN_T / 32) ? wsum[lane] : 0;
    int y = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(DLP_FULL, y, d);
      if (lane >= d) y += t;
    }
    if (lane < SCAN_T / 32) wsum[lane] = y - x;   // exclusive warp offsets
    if (lane == SCAN_T / 32 - 1 && sums) sums[blockIdx.x] = y;
  }
  __syncthreads();
  int run = wsum[w] + inc - s;
#pragma unroll
  for (int k = 0; k < SCAN_I; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
}
__global__ void k_scan_add(int* __restrict__ out, const int* __restrict__ offs, int n) {
  int i = blockIdx.x * SCAN_B + threadIdx.x;
  int o = offs[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_I; ++k) {
    int j = i + k * SCAN_T;
    if (j < n) out[j] += o;
  }
}
}  // namespace

static int scan_rec(dlpgpu_ctx* ctx, const int* in, int* out, int n, int* tmp, size_t tmp_off) {
  int nb = cdiv(n, SCAN_B);
  int* sums = tmp + tmp_off;
  LAUNCH(ctx, k_scan_block, nb, SCAN_T, 0, in, out, sums, n);
  if (nb > 1) {
    int* sums_scanned = sums + nb;
    CKRC(scan_rec(ctx, sums, sums_scanned, nb, tmp, tmp_off + 2 * (size_t)nb));
    LAUNCH(ctx, k_scan_add, nb, SCAN_T, 0, out, sums_scanned, n);
  }
  return 0;
}

// exclusive scan of in[0..n) into out[0..n); additionally out[n] = total.  in/out must have n+1 elements.
int dlp_exclusive_scan(dlpgpu_ctx* ctx, const int* in, int* out, int n, int* total_host) {
  if (n < 0) return dlp_fail(ctx, DLPGPU_ERR_ARG, "scan: n<0");
  int m = n + 1;   // scanning one extra (zero-padded by caller) element yields the total at out[n]
  CK(ctx->scan_tmp.ensure(4 * (size_t)cdiv(m, SCAN_B) + 64, ctx->stream));
  CKRC(scan_rec(ctx, in, out, m, ctx->scan_tmp.p, 0));
  if (total_host) {
    CK(cudaMemcpyAsync(total_host, out + n, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

int dlp_ensure_atoms(dlpgpu_ctx* ctx, int n) {
  if (n <= ctx->capacity) return 0;
  size_t keep = (size_t)ctx->nlast;
  cudaStream_t s = ctx->stream;
  CK(ctx->posq.ensure(n, s, true, keep));
  CK(ctx->fx.ensure(n, s, true, keep));
  CK(ctx->fy.ensure(n, s, true, keep));
  CK(ctx->fz.ensure(n, s, true, keep));
  CK(ctx->vx.ensure(n, s, true, keep));
  CK(ctx->vy.ensure(n, s, true, keep));
  CK(ctx->vz.ensure(n, s, true, keep));
  CK(ctx->ltg.ensure(n, s, true, keep));
  CK(ctx->lsite.ensure(n, s, true, keep));
  CK(ctx->ltype.ensure(n, s, true, keep));
  CK(ctx->lfrzn.ensure(n, s, true, keep));
  CK(ctx->ixyz.ensure(n, s, true, keep));
  CK(ctx->org_rank.ensure(n, s, true, keep));
  CK(ctx->org_idx.ensure(n, s, true, keep));
  CK(ctx->org_wrap.ensure(n, s, true, keep));
  CK(ctx->xbg.ensure(n, s, true, keep));
  CK(ctx->ybg.ensure(n, s, true, keep));
  CK(ctx->zbg.ensure(n, s, true, keep));
  CK(ctx->flag.ensure((size_t)n + 2, s));
  CK(ctx->scan_out.ensure((size_t)n + 2, s));
  ctx->capacity = (int)std::min(std::min(ctx->posq.cap, ctx->fx.cap), std::min(ctx->ltg.cap, ctx->xbg.cap));
  return 0;
}

// smallest double x with Sqrt(x) >= rc, so that (rsq < x) <=> (Sqrt(rsq) < rc) for IEEE round-to-nearest sqrt
static double sqrt_threshold(double rc) {
  double x = rc * rc;
  while (std::sqrt(x) >= rc) x = std::nextafter(x, 0.0);
  while (std::sqrt(x) < rc) x = std::nextafter(x, DBL_MAX);
  return x;
}

extern "C" {

int dlpgpu_version(void) { return 100; }

int dlpgpu_create(dlpgpu_ctx** out, int device) {
  if (!out) return DLPGPU_ERR_ARG;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) return DLPGPU_ERR_CUDA;
  dlpgpu_ctx* ctx = new dlpgpu_ctx();
  ctx->device = device;
  if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return DLPGPU_ERR_CUDA; }
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return DLPGPU_ERR_CUDA; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
  for (int i = 0; i < 8; ++i) cudaEventCreate(&ctx->ev[i]);
  cudaEventCreate(&ctx->ev_res);
  {   // once per device and process
    static bool loaded[64] = {false};
    if (device < 64 && !loaded[device]) {   // keep freed DBuf memory in the stream-ordered pool instead of returning it at every synchronisation
      cudaMemPool_t pool;
      unsigned long long keep = ~0ull;
      if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      cudaGetLastError();
    }
    if (device < 64 && !loaded[device]) { dlp_preload_ctx(); dlp_preload_cells(); dlp_preload_forces(); dlp_preload_halo(); dlp_preload_spme(); dlp_preload_hostio(); loaded[device] = true; }
  }
  if (ctx->status.ensure(8, ctx->stream) != cudaSuccess || ctx->out_dev.ensure(16, ctx->stream) != cudaSuccess ||
      ctx->tol_bits.ensure(2, ctx->stream) != cudaSuccess || ctx->cnt64.ensure(4, ctx->stream) != cudaSuccess) { delete ctx; return DLPGPU_ERR_CUDA; }
  *out = ctx;
  return 0;
}

int dlpgpu_destroy(dlpgpu_ctx* ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  // buffers are released with the context; DBuf has no destructor on purpose (explicit lifetime)
  DBuf<int>* ib[] = {&ctx->pair_k, &ctx->ltp, &ctx->type_site, &ctx->freeze_site, &ctx->excl, &ctx->ltg, &ctx->lsite, &ctx->ltype,
                     &ctx->lfrzn, &ctx->ixyz, &ctx->org_rank, &ctx->org_idx, &ctx->org_wrap, &ctx->which_cell, &ctx->at_list, &ctx->at_tmp, &ctx->lct_count, &ctx->lct_start,
                     &ctx->lct_fill, &ctx->cell_s, &ctx->loc_slot, &ctx->flag, &ctx->scan_out, &ctx->scan_tmp, &ctx->type_s,
                     &ctx->gid_s, &ctx->frz_s, &ctx->st_nix, &ctx->st_niy, &ctx->st_niz, &ctx->st_nir, &ctx->st_xb, &ctx->ref_list,
                     &ctx->nnbr, &ctx->nxnbr, &ctx->nhnbr, &ctx->status};
  for (auto* b : ib) b->release(ctx->stream);
  DBuf<double>* db[] = {&ctx->vdw_par, &ctx->charge_site, &ctx->weight_site, &ctx->fx, &ctx->fy, &ctx->fz, &ctx->vx, &ctx->vy,
                        &ctx->vz, &ctx->xbg, &ctx->ybg, &ctx->zbg, &ctx->partial, &ctx->out_dev};
  for (auto* b : db) b->release(ctx->stream);
  ctx->vdw_tab.release(ctx->stream); ctx->ew_tab.release(ctx->stream); ctx->posq.release(ctx->stream); ctx->posq_s.release(ctx->stream);
  ctx->nbr.release(ctx->stream); ctx->xnbr.release(ctx->stream); ctx->hnbr.release(ctx->stream); ctx->tol_bits.release(ctx->stream); ctx->parts_dev.release(ctx->stream);
  if (ctx->tab2_tex) { cudaDestroyTextureObject(ctx->tab2_tex); ctx->tab2_tex = 0; }
  if (ctx->tab2h_tex) { cudaDestroyTextureObject(ctx->tab2h_tex); ctx->tab2h_tex = 0; }
  dlp_spme_release(ctx);
  ctx->fnbr.release(ctx->stream); ctx->nfnbr.release(ctx->stream); ctx->row_perm.release(ctx->stream); ctx->cell_box.release(ctx->stream); ctx->movers.release(ctx->stream);
  ctx->pp_pos.release(ctx->stream); ctx->pp_neg.release(ctx->stream); ctx->pp_energy.release(ctx->stream); ctx->pp_stress.release(ctx->stream);
  ctx->tab2h.release(ctx->stream); ctx->tab2s.release(ctx->stream); ctx->rdf_list.release(ctx->stream); ctx->rdf_hist.release(ctx->stream);
  ctx->tab4.release(ctx->stream); ctx->tab2.release(ctx->stream); ctx->cnt64.release(ctx->stream); ctx->info_s.release(ctx->stream); ctx->st_rows.release(ctx->stream);
  for (int i = 0; i < 6; ++i) ctx->stage[i].idx.release(ctx->stream);
  for (int r = 0; r < ctx->p2p_nranks; ++r)
    if (r != ctx->p2p_rank && !(ctx->peer_pub_local.size() > (size_t)r && ctx->peer_pub_local[r]))
      for (int b = 0; b < 2; ++b)
        if (ctx->peer_pub.size() > (size_t)(2 * r + b) && ctx->peer_pub[2 * r + b]) cudaIpcCloseMemHandle(ctx->peer_pub[2 * r + b]);
  for (int b = 0; b < 2; ++b) if (ctx->pub[b]) cudaFree(ctx->pub[b]);
  ctx->peer_pub_dev.release(ctx->stream);
  for (int r = 0; r < ctx->xr_nranks; ++r)
    if (r != ctx->xr_rank && ctx->peer_xr.size() > (size_t)r && ctx->peer_xr[r] &&
        !(ctx->peer_xr_local.size() > (size_t)r && ctx->peer_xr_local[r])) cudaIpcCloseMemHandle(ctx->peer_xr[r]);
  if (ctx->xr) cudaFree(ctx->xr);
  ctx->peer_xr_dev.release(ctx->stream); ctx->dcnt.release(ctx->stream); ctx->gmax_out.release(ctx->stream);
  dlp_hostio_release(ctx);
  for (int i = 0; i < 8; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->ev_res) cudaEventDestroy(ctx->ev_res);
  for (int i = 0; i < 2; ++i) if (ctx->ev_x[i]) cudaEventDestroy(ctx->ev_x[i]);
  if (ctx->out_pinned) cudaFreeHost(ctx->out_pinned);
  if (ctx->gm_pinned) cudaFreeHost(ctx->gm_pinned);
  if (ctx->dc_pinned) cudaFreeHost(ctx->dc_pinned);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamDestroy(ctx->stream);
  delete ctx;
  return 0;
}

const char* dlpgpu_last_error(const dlpgpu_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
long long dlpgpu_launch_count(const dlpgpu_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* dlpgpu_stream(dlpgpu_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int dlpgpu_set_domain(dlpgpu_ctx* ctx, const int dd[6]) {
  if (!ctx || !dd) return DLPGPU_ERR_ARG;
  if (dd[0] < 1 || dd[1] < 1 || dd[2] < 1 || dd[3] < 0 || dd[3] >= dd[0] || dd[4] < 0 || dd[4] >= dd[1] || dd[5] < 0 ||
      dd[5] >= dd[2])
    return dlp_fail(ctx, DLPGPU_ERR_ARG, "set_domain: inconsistent decomposition");
  ctx->nx = dd[0]; ctx->ny = dd[1]; ctx->nz = dd[2]; ctx->idx = dd[3]; ctx->idy = dd[4]; ctx->idz = dd[5];
  ctx->list_valid = false;
  return 0;
}

int dlpgpu_set_cell(dlpgpu_ctx* ctx, const double cell[9], int imcon) {
  if (!ctx || !cell) return DLPGPU_ERR_ARG;
  if (!(imcon == 0 || imcon == 1 || imcon == 2 || imcon == 3)) return dlp_fail(ctx, DLPGPU_ERR_ARG, "imcon %d not supported", imcon);
  for (int i = 0; i < 9; ++i) ctx->cell[i] = cell[i];
  ctx->imcon = imcon;
  return 0;
}

int dlpgpu_set_cutoffs(dlpgpu_ctx* ctx, double rcut, double padding, double pdplnc) {
  if (!ctx) return DLPGPU_ERR_ARG;
  if (!(rcut > 0.0) || padding < 0.0) return dlp_fail(ctx, DLPGPU_ERR_ARG, "set_cutoffs: rcut must be > 0, padding >= 0");
  ctx->rcut = rcut; ctx->padding = padding; ctx->rx = rcut + padding;   // bounds.F90:1298
  ctx->pdplnc = pdplnc > 0.0 ? pdplnc : 50.0;
  ctx->thr_coul = sqrt_threshold(rcut);
  ctx->list_valid = false;
  return 0;
}

int dlpgpu_set_vdw(dlpgpu_ctx* ctx, int ntypes, const int* vdw_list, int max_vdw, int n_vdw, const int* ltp, int max_grid,
                   const double* tab_potential, const double* tab_force, double rvdw, int force_shift, int direct,
                   const double* param, const double* afs, const double* bfs) {
  if (!ctx) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  ctx->tab4_valid = false; ctx->list_valid = false;   // list entries carry the potential index of their type pair
  if (n_vdw <= 0) { ctx->vdw_on = false; ctx->ntypes = ntypes; return 0; }
  if (max_vdw > DLP_MAX_KCODE) return dlp_fail(ctx, DLPGPU_ERR_ARG, "set_vdw: more than %d vdW potentials are not supported", DLP_MAX_KCODE);
  if (ntypes < 1 || !vdw_list || !ltp || max_vdw < n_vdw) return dlp_fail(ctx, DLPGPU_ERR_ARG, "set_vdw: bad arguments");
  if (!direct && (!tab_potential || !tab_force || max_grid < 8)) return dlp_fail(ctx, DLPGPU_ERR_ARG, "set_vdw: tables missing");
  if (direct && !param) return dlp_fail(ctx, DLPGPU_ERR_ARG, "set_vdw: direct evaluation needs param");
  if (direct)   // vdw_forces_direct evaluates analytic forms only: keys 1..24 of vdw.F90:71-113 (a TABLE-file potential has none)
    for (int k = 0; k < n_vdw; ++k)
      if (ltp[k] != -1 && (ltp[k] < 1 || ltp[k] > 24))
        return dlp_fail(ctx, DLPGPU_ERR_ARG, "set_vdw: potential %d has key %d, which vdw_method direct cannot evaluate", k + 1, ltp[k]);
  ctx->ntypes = ntypes; ctx->n_vdw = n_vdw; ctx->max_vdw = max_vdw; ctx->max_grid = max_grid;
  ctx->rvdw = rvdw; ctx->vdw_fshift = force_shift != 0; ctx->vdw_direct = direct != 0;
  if (max_grid > 4) {   // vdw.F90:1836-1837
    double dlrpot = rvdw / (double)(max_grid - 4);
    ctx->vdw_rdr = 1.0 / dlrpot;
  }
  ctx->thr_vdw = sqrt_threshold(rvdw);
  const double zero_plus = DBL_MIN;
  // dense (ai,aj) -> k map.  vdw.F90:1875-1892: key, k = list(key); skipped when |tab_potential(0,k)| < zero_plus
  // (tabulated path only) or ltp(k) == VDW_NULL.
  std::vector<int> pk((size_t)ntypes * ntypes, -1);
  for (int ai = 1; ai <= ntypes; ++ai)
    for (int aj = 1; aj <= ntypes; ++aj) {
      int key = (ai > aj) ? ai * (ai - 1) / 2 + aj : aj * (aj - 1) / 2 + ai;
      int k = vdw_list[key - 1];
      if (k < 1 || k > max_vdw) continue;
      if (ltp[k - 1] == -1) continue;
      if (!direct && std::fabs(tab_potential[(size_t)(k - 1) * (max_grid + 1)]) < zero_plus) continue;
      pk[(size_t)(ai - 1) * ntypes + (aj - 1)] = k - 1;
    }
  CK(ctx->pair_k.ensure(pk.size(), ctx->stream));
  CK(cudaMemcpyAsync(ctx->pair_k.p, pk.data(), pk.size() * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(ctx->ltp.ensure(max_vdw, ctx->stream));
  CK(cudaMemcpyAsync(ctx->ltp.p, ltp, max_vdw * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  std::vector<double> par((size_t)max_vdw * 10, 0.0);
  for (int k = 0; k < max_vdw; ++k) {
    if (param) for (int q = 0; q < 7; ++q) par[(size_t)k * 10 + q] = param[(size_t)k * 7 + q];
    if (afs) par[(size_t)k * 10 + 7] = afs[k];
    if (bfs) par[(size_t)k * 10 + 8] = bfs[k];
  }
  CK(ctx->vdw_par.ensure(par.size(), ctx->stream));
  CK(cudaMemcpyAsync(ctx->vdw_par.p, par.data(), par.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (tab_potential && tab_force && max_grid > 4) {
    size_t n = (size_t)max_vdw * (max_grid + 1);
    std::vector<double2> t(n);
    for (size_t i = 0; i < n; ++i) t[i] = make_double2(tab_force[i], tab_potential[i]);
    ctx->h_vdw_f.assign(tab_force, tab_force + n);
    ctx->h_vdw_e.assign(tab_potential, tab_potential + n);
    CK(ctx->vdw_tab.ensure(n, ctx->stream));
    CK(cudaMemcpyAsync(ctx->vdw_tab.p, t.data(), n * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->vdw_on = true;
  return 0;
}

int dlpgpu_set_ewald(dlpgpu_ctx* ctx, int active, double alpha, double scaling, int nsamples, const double* erfc_tab,
                     const double* erfc_deriv_tab, double recip_spacing) {
  if (!ctx) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  ctx->tab4_valid = false;
  ctx->coul_kind = 0; ctx->coul_damp = false;
  if (!active) { ctx->ew_on = false; return 0; }
  if (!erfc_tab || !erfc_deriv_tab || nsamples < 10) return dlp_fail(ctx, DLPGPU_ERR_ARG, "set_ewald: tables missing");
  ctx->alpha = alpha; ctx->scaling = scaling; ctx->ew_n = nsamples; ctx->ew_rdr = recip_spacing;
  std::vector<double2> t((size_t)nsamples + 1);
  for (int i = 0; i <= nsamples; ++i) t[i] = make_double2(erfc_deriv_tab[i], erfc_tab[i]);
  ctx->h_ew_d.assign(erfc_deriv_tab, erfc_deriv_tab + nsamples + 1);
  ctx->h_ew_e.assign(erfc_tab, erfc_tab + nsamples + 1);
  CK(ctx->ew_tab.ensure(t.size(), ctx->stream));
  CK(cudaMemcpyAsync(ctx->ew_tab.p, t.data(), t.size() * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->ew_on = true;
  return 0;
}

int dlpgpu_set_coulomb(dlpgpu_ctx* ctx, int kind, int damp, double scaling, double force_shift, double energy_shift,
                       const double reaction_field[3], int nsamples, const double* erfc_tab, const double* erfc_deriv_tab,
                       double recip_spacing) {
  if (!ctx || kind < DLPGPU_COUL_CP || kind > DLPGPU_COUL_RFP) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  ctx->tab4_valid = false;
  ctx->ew_on = false;
  ctx->coul_kind = kind; ctx->coul_damp = damp != 0 && (kind == DLPGPU_COUL_FSCP || kind == DLPGPU_COUL_RFP);
  ctx->scaling = scaling; ctx->coul_fs = force_shift; ctx->coul_es = energy_shift;
  for (int k = 0; k < 3; ++k) ctx->coul_rf[k] = reaction_field ? reaction_field[k] : 0.0;
  if (ctx->coul_damp) {
    if (!erfc_tab || !erfc_deriv_tab || nsamples < 10) return dlp_fail(ctx, DLPGPU_ERR_ARG, "set_coulomb: damped variant without erfc tables");
    ctx->ew_n = nsamples; ctx->ew_rdr = recip_spacing;
    std::vector<double2> t((size_t)nsamples + 1);
    for (int i = 0; i <= nsamples; ++i) t[i] = make_double2(erfc_deriv_tab[i], erfc_tab[i]);
    ctx->h_ew_d.assign(erfc_deriv_tab, erfc_deriv_tab + nsamples + 1);
    ctx->h_ew_e.assign(erfc_tab, erfc_tab + nsamples + 1);
    CK(ctx->ew_tab.ensure(t.size(), ctx->stream));
    CK(cudaMemcpyAsync(ctx->ew_tab.p, t.data(), t.size() * sizeof(double2), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

int dlpgpu_set_force_mode(dlpgpu_ctx* ctx, int mode) {
  if (!ctx || mode < 0 || mode > 1) return DLPGPU_ERR_ARG;
  if (mode != ctx->force_mode) ctx->list_valid = false;
  ctx->force_mode = mode;
  return 0;
}

int dlpgpu_set_pair_kernel(dlpgpu_ctx* ctx, int which) {
  if (!ctx || which < 0 || which > 1) return DLPGPU_ERR_ARG;
  ctx->no_fast = which == 1;
  return 0;
}

int dlpgpu_set_list_kernel(dlpgpu_ctx* ctx, int which) {
  if (!ctx || which < 0 || which > 2) return DLPGPU_ERR_ARG;
  ctx->list_one_atom_per_pass = which;
  return 0;
}

int dlpgpu_last_timings(dlpgpu_ctx* ctx, double t[4]) {
  if (!ctx || !t) return DLPGPU_ERR_ARG;
  t[0] = ctx->t_list; t[1] = ctx->t_force; t[2] = ctx->t_pair; t[3] = ctx->t_full;
  return 0;
}

// ------------------------------------------------------------------ native setup
int dlpgpu_dev_set_sites(dlpgpu_ctx* ctx, int nsites, const int* type_site, const double* charge_site, const int* freeze_site,
                         const double* weight_site) {
  if (!ctx || nsites < 1 || !type_site || !charge_site) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  std::vector<int> fr(nsites, 0);
  std::vector<double> wt(nsites, 1.0);
  if (freeze_site) fr.assign(freeze_site, freeze_site + nsites);
  if (weight_site) wt.assign(weight_site, weight_site + nsites);
  ctx->nsites = nsites;
  CK(ctx->type_site.ensure(nsites, ctx->stream)); CK(ctx->freeze_site.ensure(nsites, ctx->stream));
  CK(ctx->charge_site.ensure(nsites, ctx->stream)); CK(ctx->weight_site.ensure(nsites, ctx->stream));
  CK(cudaMemcpyAsync(ctx->type_site.p, type_site, nsites * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->freeze_site.p, fr.data(), nsites * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->charge_site.p, charge_site, nsites * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->weight_site.p, wt.data(), nsites * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int dlpgpu_dev_set_excl(dlpgpu_ctx* ctx, int megatm, int max_exclude, const int* excl_by_gid) {
  if (!ctx || megatm < 1 || max_exclude < 0 || !excl_by_gid) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  size_t n = (size_t)megatm * (max_exclude + 1);
  CK(ctx->excl.ensure(n, ctx->stream));
  CK(cudaMemcpyAsync(ctx->excl.p, excl_by_gid, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->lbook = 1; ctx->max_exclude = max_exclude; ctx->excl_by_gid = 1;
  ctx->list_valid = false;
  return 0;
}

int dlpgpu_dev_set_halo_width(dlpgpu_ctx* ctx, const double ecw[3]) {
  if (!ctx || !ecw) return DLPGPU_ERR_ARG;
  for (int i = 0; i < 3; ++i) ctx->ecw[i] = ecw[i];
  return 0;
}

int dlpgpu_dev_set_list_capacity(dlpgpu_ctx* ctx, int max_list, int megfrz) {
  if (!ctx || max_list < 1) return DLPGPU_ERR_ARG;
  ctx->max_list = max_list; ctx->megfrz = megfrz;
  ctx->list_valid = false;
  return 0;
}

int dlpgpu_dev_counts(dlpgpu_ctx* ctx, int* natms, int* nlast) {
  if (!ctx) return DLPGPU_ERR_ARG;
  if (natms) *natms = ctx->natms;
  if (nlast) *nlast = ctx->nlast;
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------ AoS <-> device layout kernels
namespace {
__global__ void k_pack_parts(dlpgpu_corepart* __restrict__ parts, int n, const double4* __restrict__ posq,
                             const double* __restrict__ fx, const double* __restrict__ fy, const double* __restrict__ fz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 q = posq[i];
  dlpgpu_corepart p;
  p.xxx = q.x; p.yyy = q.y; p.zzz = q.z; p.fxx = fx[i]; p.fyy = fy[i]; p.fzz = fz[i]; p.chge = q.w; p.pad1 = 0; p.pad2 = 0;
  parts[i] = p;
}
__global__ void k_load_atoms(int n, const double* __restrict__ xyz, const double* __restrict__ vel, const int* __restrict__ lsite,
                             const int* __restrict__ type_site, const double* __restrict__ charge_site,
                             const int* __restrict__ freeze_site, double4* __restrict__ posq, double* vx, double* vy, double* vz,
                             int* ltype, int* lfrzn, int* ixyz, double* fx, double* fy, double* fz) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int s = lsite[i] - 1;
  posq[i] = make_double4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], charge_site[s]);
  vx[i] = vel ? vel[3 * i] : 0.0; vy[i] = vel ? vel[3 * i + 1] : 0.0; vz[i] = vel ? vel[3 * i + 2] : 0.0;
  ltype[i] = type_site[s]; lfrzn[i] = freeze_site[s]; ixyz[i] = 0;
  fx[i] = 0.0; fy[i] = 0.0; fz[i] = 0.0;
}
__global__ void k_zero3(int n, double* a, double* b, double* c) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { a[i] = 0.0; b[i] = 0.0; c[i] = 0.0; }
}
}  // namespace

static int upload_parts(dlpgpu_ctx* ctx, int n, const dlpgpu_corepart* parts) { return dlp_upload_parts(ctx, n, parts); }   // hostio.cu

extern "C" {

int dlpgpu_dev_load_atoms(dlpgpu_ctx* ctx, int natms, const double* xyz, const double* vel, const int* ltg, const int* lsite,
                          int capacity_atoms) {
  if (!ctx || natms < 0 || !xyz || !ltg || !lsite) return DLPGPU_ERR_ARG;
  if (ctx->nsites < 1) return dlp_fail(ctx, DLPGPU_ERR_STATE, "load_atoms: call dlpgpu_dev_set_sites first");
  CK(cudaSetDevice(ctx->device));
  ctx->natms = 0; ctx->nlast = 0;
  dlp_hostio_ints_stale(ctx);
  CKRC(dlp_ensure_atoms(ctx, std::max(capacity_atoms, natms + 16)));
  DBuf<double> tx, tv;
  CK(tx.ensure((size_t)3 * natms + 1, ctx->stream));
  CK(cudaMemcpyAsync(tx.p, xyz, (size_t)3 * natms * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  if (vel) {
    CK(tv.ensure((size_t)3 * natms + 1, ctx->stream));
    CK(cudaMemcpyAsync(tv.p, vel, (size_t)3 * natms * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  }
  CK(cudaMemcpyAsync(ctx->ltg.p, ltg, (size_t)natms * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->lsite.p, lsite, (size_t)natms * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  if (natms > 0)
    LAUNCH(ctx, k_load_atoms, cdiv(natms, 256), 256, 0, natms, tx.p, vel ? tv.p : nullptr, ctx->lsite.p, ctx->type_site.p,
           ctx->charge_site.p, ctx->freeze_site.p, ctx->posq.p, ctx->vx.p, ctx->vy.p, ctx->vz.p, ctx->ltype.p, ctx->lfrzn.p,
           ctx->ixyz.p, ctx->fx.p, ctx->fy.p, ctx->fz.p);
  CK(cudaStreamSynchronize(ctx->stream));
  tx.release(); tv.release();
  ctx->natms = natms; ctx->nlast = natms;
  ctx->list_valid = false; ctx->halo_valid = false; ctx->have_bg = false; ctx->tol_fresh = false; ctx->pub_fresh = false;
  return 0;
}

int dlpgpu_dev_zero_forces(dlpgpu_ctx* ctx) {
  if (!ctx) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (ctx->nlast > 0) LAUNCH(ctx, k_zero3, cdiv(ctx->nlast, 256), 256, 0, ctx->nlast, ctx->fx.p, ctx->fy.p, ctx->fz.p);
  return 0;
}

// ------------------------------------------------------------------ drop-in (host buffer) entry points
int dlpgpu_link_cell_pairs(dlpgpu_ctx* ctx, int natms, int nlast, const dlpgpu_corepart* parts, const int* ltype, const int* ltg,
                           const int* lfrzn, int lbook, int megfrz, int max_exclude, const int* list_excl, int max_list,
                           int* list_out, int* ibig) {
  if (!ctx || natms < 0 || nlast < natms || !parts || !ltype || !ltg || max_list < 1) return DLPGPU_ERR_ARG;
  if (lbook && max_exclude > 0 && !list_excl) return dlp_fail(ctx, DLPGPU_ERR_ARG, "link_cell_pairs: lbook set but list_excl is NULL");
  CK(cudaSetDevice(ctx->device));
  ctx->natms = 0; ctx->nlast = 0;
  CKRC(dlp_ensure_atoms(ctx, nlast + 16));
  CKRC(upload_parts(ctx, nlast, parts));
  ctx->parts_resident = nlast;
  CKRC(dlp_upload_ints(ctx, nlast, ltype, ltg, lfrzn));
  ctx->lbook = lbook ? 1 : 0; ctx->megfrz = megfrz; ctx->max_exclude = lbook ? max_exclude : 0; ctx->excl_by_gid = 0;
  if (lbook && list_excl) {
    size_t n = (size_t)natms * (max_exclude + 1);
    CK(ctx->excl.ensure(n + 1, ctx->stream));
    CK(cudaMemcpyAsync(ctx->excl.p, list_excl, n * sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
  }
  ctx->natms = natms; ctx->nlast = nlast; ctx->max_list = max_list;
  ctx->halo_valid = false;
  CKRC(dlp_vnl_set_check(ctx));
  int big = 0;
  int rc = dlp_build_lists(ctx, list_out != nullptr, &big);
  if (ibig) *ibig = big;
  if (rc) return rc;
  if (list_out) {
    CK(cudaMemcpyAsync(list_out, ctx->ref_list.p, (size_t)natms * (max_list + 4) * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  return 0;
}

int dlpgpu_two_body_forces(dlpgpu_ctx* ctx, int natms, int nlast, dlpgpu_corepart* parts, double out[16]) {
  if (!ctx || !parts || !out) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->list_valid) return dlp_fail(ctx, DLPGPU_ERR_STATE, "two_body_forces: no valid neighbour list (call link_cell_pairs)");
  if (natms != ctx->list_natms || nlast != ctx->list_nlast)
    return dlp_fail(ctx, DLPGPU_ERR_HALO_COUNT, "two_body_forces: natms/nlast (%d/%d) differ from the list build (%d/%d)", natms, nlast,
                    ctx->list_natms, ctx->list_nlast);
  if (ctx->parts_current && ctx->parts_resident >= nlast) {
    // the records link_cell_pairs uploaded (and unpacked into the device arrays) are still the caller's
  } else {
    CKRC(upload_parts(ctx, nlast, parts));
  }
  ctx->parts_current = false;
  // every force provider ADDS (drivers.F90:655-660): the device computes this provider's forces from zero, they come back as
  // {fx, fy, fz} triples behind the kernels and the host adds them into parts(1:natms)%f (hostio.cu) -- whatever other
  // providers put there since the upload stays.  The sums are fetched last, so the copies start without a host round trip.
  CKRC(dlp_two_body(ctx, 1, nullptr));
  CKRC(dlp_download_add_forces(ctx, natms, parts));
  CKRC(dlpgpu_dev_fetch_results(ctx, out));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

// ewald_spme_forces_coul with the caller's corePart array (one domain): parts(1:natms) go up, the reciprocal forces are ADDED
// to parts%f on the device and the records come back; out as for dlpgpu_dev_spme_forces
int dlpgpu_spme_forces(dlpgpu_ctx* ctx, int natms, dlpgpu_corepart* parts, int megatm, double out[16]) {
  if (!ctx || !parts || !out || natms < 0 || megatm < 1) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CKRC(dlp_ensure_atoms(ctx, natms + 16));
  CKRC(upload_parts(ctx, natms, parts));
  dlp_hostio_ints_stale(ctx);
  const int keep_n = ctx->natms, keep_l = ctx->nlast;
  ctx->natms = natms;
  if (natms > 0) LAUNCH(ctx, k_zero3, cdiv(natms, 256), 256, 0, natms, ctx->fx.p, ctx->fy.p, ctx->fz.p);
  const int rc = dlpgpu_dev_spme_forces(ctx, megatm, out);
  ctx->natms = keep_n; ctx->nlast = keep_l;
  if (rc) return rc;
  CKRC(dlp_download_add_forces(ctx, natms, parts));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->list_valid = false;   // the device force / position arrays were reused: the next short-range call starts from link_cell_pairs
  return 0;
}

int dlpgpu_parts_unchanged_since_list(dlpgpu_ctx* ctx) {
  if (!ctx) return DLPGPU_ERR_ARG;
  if (!ctx->list_valid || ctx->parts_resident < ctx->list_nlast)
    return dlp_fail(ctx, DLPGPU_ERR_STATE, "parts_unchanged_since_list: no list build has uploaded the parts array");
  ctx->parts_current = true;
  return 0;
}

int dlpgpu_vnl_set_check(dlpgpu_ctx* ctx, int nlast, const dlpgpu_corepart* parts) {
  if (!ctx || !parts || nlast < 0) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  CKRC(dlp_ensure_atoms(ctx, nlast + 16));
  CKRC(upload_parts(ctx, nlast, parts));
  if (ctx->nlast < nlast) ctx->nlast = nlast;
  CKRC(dlp_vnl_set_check(ctx));
  CK(cudaStreamSynchronize(ctx->stream));   // the caller's array is page-locked, so the upload is a real DMA: it has to be over before the caller may touch parts again
  return 0;
}

int dlpgpu_vnl_check(dlpgpu_ctx* ctx, int natms, const dlpgpu_corepart* parts, double* tol) {
  if (!ctx || !parts || !tol) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->have_bg) return dlp_fail(ctx, DLPGPU_ERR_STATE, "vnl_check: no checkpoint (vnl_set_check / link_cell_pairs not called)");
  if (natms > ctx->capacity) return dlp_fail(ctx, DLPGPU_ERR_ARG, "vnl_check: natms exceeds the checkpoint size");
  CKRC(upload_parts(ctx, natms, parts));
  int keep = ctx->natms;
  ctx->natms = natms;
  int rc = dlp_vnl_check(ctx, tol);
  ctx->natms = keep;
  return rc;
}

// ------------------------------------------------------------------ read-back
int dlpgpu_dev_get_parts(dlpgpu_ctx* ctx, dlpgpu_corepart* parts_out, int n) {
  if (!ctx || !parts_out || n < 0 || n > ctx->nlast) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (n == 0) return 0;
  CK(ctx->parts_dev.ensure(n, ctx->stream));
  LAUNCH(ctx, k_pack_parts, cdiv(n, 256), 256, 0, ctx->parts_dev.p, n, ctx->posq.p, ctx->fx.p, ctx->fy.p, ctx->fz.p);
  CK(cudaMemcpyAsync(parts_out, ctx->parts_dev.p, (size_t)n * sizeof(dlpgpu_corepart), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int dlpgpu_dev_get_ints(dlpgpu_ctx* ctx, int n, int* ltg, int* lsite, int* ltype, int* lfrzn, int* ixyz) {
  if (!ctx || n < 0 || n > ctx->nlast) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  size_t b = (size_t)n * sizeof(int);
  if (ltg) CK(cudaMemcpyAsync(ltg, ctx->ltg.p, b, cudaMemcpyDeviceToHost, ctx->stream));
  if (lsite) CK(cudaMemcpyAsync(lsite, ctx->lsite.p, b, cudaMemcpyDeviceToHost, ctx->stream));
  if (ltype) CK(cudaMemcpyAsync(ltype, ctx->ltype.p, b, cudaMemcpyDeviceToHost, ctx->stream));
  if (lfrzn) CK(cudaMemcpyAsync(lfrzn, ctx->lfrzn.p, b, cudaMemcpyDeviceToHost, ctx->stream));
  if (ixyz) CK(cudaMemcpyAsync(ixyz, ctx->ixyz.p, b, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int dlpgpu_dev_get_vel(dlpgpu_ctx* ctx, int n, double* vel3) {
  if (!ctx || !vel3 || n < 0 || n > ctx->natms) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  std::vector<double> t((size_t)3 * n + 3);
  CK(cudaMemcpyAsync(t.data(), ctx->vx.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(t.data() + n, ctx->vy.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(t.data() + 2 * (size_t)n, ctx->vz.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (int i = 0; i < n; ++i) { vel3[3 * i] = t[i]; vel3[3 * i + 1] = t[n + i]; vel3[3 * i + 2] = t[2 * (size_t)n + i]; }
  return 0;
}

int dlpgpu_dev_get_list(dlpgpu_ctx* ctx, int natms, int max_list, int* list_out) {
  if (!ctx || !list_out) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->ref_valid || natms != ctx->list_natms || max_list != ctx->max_list)
    return dlp_fail(ctx, DLPGPU_ERR_STATE, "get_list: reference-format list not materialised for these sizes");
  CK(cudaMemcpyAsync(list_out, ctx->ref_list.p, (size_t)natms * (max_list + 4) * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int dlpgpu_dev_get_cells(dlpgpu_ctx* ctx, int info[6], int* which_cell, int* at_list, int* lct_start) {
  if (!ctx) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->list_valid) return dlp_fail(ctx, DLPGPU_ERR_STATE, "get_cells: no list build yet");
  const LCGeom& g = ctx->g;
  if (info) { info[0] = g.nlx; info[1] = g.nly; info[2] = g.nlz; info[3] = g.nlp; info[4] = g.ncells; info[5] = g.nsbcll; }
  int n = ctx->list_nlast;
  if (which_cell) CK(cudaMemcpyAsync(which_cell, ctx->which_cell.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (at_list) CK(cudaMemcpyAsync(at_list, ctx->at_list.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  if (lct_start) CK(cudaMemcpyAsync(lct_start, ctx->lct_start.p, (size_t)(g.ncells + 2) * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int dlpgpu_dev_get_full_row(dlpgpu_ctx* ctx, int i, int* n_main, int* main_out, int* n_excl, int* excl_out, int cap) {
  if (!ctx || i < 1 || i > ctx->list_natms || !n_main || !main_out || !n_excl || !excl_out) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (!ctx->list_valid) return dlp_fail(ctx, DLPGPU_ERR_STATE, "get_full_row: no list");
  // host-side search of the row that belongs to local atom i: download loc_slot/at_list once per call (test hook only)
  int natms = ctx->list_natms, nlast = ctx->list_nlast;
  std::vector<int> loc(natms), atl(nlast), nn(natms), nx(natms);
  CK(cudaMemcpy(loc.data(), ctx->loc_slot.p, natms * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(atl.data(), ctx->at_list.p, nlast * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(nn.data(), ctx->nnbr.p, natms * sizeof(int), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(nx.data(), ctx->nxnbr.p, natms * sizeof(int), cudaMemcpyDeviceToHost));
  int t = -1;
  for (int k = 0; k < natms; ++k) if (atl[loc[k]] == i - 1) { t = k; break; }
  if (t < 0) return dlp_fail(ctx, DLPGPU_ERR_STATE, "get_full_row: atom not found");
  std::vector<unsigned> row(ctx->pitch), xrow(std::max(ctx->xpitch, 1));
  CK(cudaMemcpy(row.data(), ctx->nbr.p + (size_t)t * ctx->pitch, ctx->pitch * sizeof(unsigned), cudaMemcpyDeviceToHost));
  if (ctx->xpitch > 0) CK(cudaMemcpy(xrow.data(), ctx->xnbr.p + (size_t)t * ctx->xpitch, ctx->xpitch * sizeof(unsigned), cudaMemcpyDeviceToHost));
  *n_main = nn[t]; *n_excl = nx[t];
  for (int k = 0; k < nn[t] && k < cap; ++k) main_out[k] = atl[row[k] & DLP_J_MASK] + 1;
  for (int k = 0; k < nx[t] && k < cap; ++k) excl_out[k] = atl[xrow[k] & DLP_J_MASK] + 1;
  return 0;
}

}  // extern "C"

int dlp_preload_ctx() {   // see dlp_preload_halo
  const void* ks[] = {(const void*)k_scan_block, (const void*)k_scan_add, (const void*)k_pack_parts,
                      (const void*)k_load_atoms, (const void*)k_zero3};
  cudaFuncAttributes a;
  for (const void* k : ks) if (cudaFuncGetAttributes(&a, k) != cudaSuccess) cudaGetLastError();
  return 0;
}
