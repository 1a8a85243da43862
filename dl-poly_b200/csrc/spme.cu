// spme.cu -- SPME reciprocal-space Coulomb term (beyond the north star's hot path, first version):
//
//   ewald_spole.F90:244-477     ewald_spme_forces_coul (driver: scaled coordinates, scale factors, energy / stress assembly)
//   bspline.F90:73-190          bspline_coeffs_gen  -> |b(m)|^2 per dimension (host, once per grid)
//   bspline.F90:192-306         bspline_splines_gen -> values and first derivatives, recomputed per atom in both kernels
//   ewald_general.F90:517-576   spme_construct_charge_array        -> k_spme_spread (fp64 RED onto the grid)
//   ewald_spole.F90:1257-1386   spme_construct_potential_grid_coul -> cuFFT Z2Z + k_spme_influence (+ stress kernel) + cuFFT Z2Z
//   ewald_general.F90:717-869   spme_calc_force_energy             -> k_spme_gather + k_spme_finish (net force removed)
//   spme.F90:159-231            spme_self_interaction
//
// One domain holds the whole grid (mxnode = 1): what the reference assembles from the + side halo images is the periodic wrap of
// the spline footprint here.  The 3-D transforms are cuFFT's (a library FFT; the reference uses its own DaFT / gpfa code), complex
// to complex so that the Nyquist planes carry exactly the reference's per-element kernel.  cuFFT is loaded on first use, so the
// short-range library does not depend on it.
#include "common.cuh"

#include <dlfcn.h>

namespace {

constexpr int SPME_MAXN = 12;
constexpr double SQRPI = 1.7724538509055160273;   // constants.F90:56
constexpr double ZERO_PLUS = 2.2250738585072014e-308;

struct SpmeP {
  int K[3];
  int n, natms;
  double rc[9];   // inverse cell, rc[d + 3 j] = rcell(d + 1 + 3 j) of invert(): s_d = rc[d] x + rc[d + 3] y + rc[d + 6] z
};

// v[j] = M_n(w + j), d[j] = M_n'(w + j), j = 0..n-1 (cardinal B-spline of order n, Cox-de Boor): grid point Int(u) - j carries
// v[j], which is derivs(:, 0, n - j, i) of bspline_splines_gen
__device__ void bspline_fill(double w, int n, double* v, double* d) {
  v[0] = w; v[1] = 1.0 - w;
  for (int j = 2; j < n; ++j) v[j] = 0.0;
  for (int k = 3; k <= n; ++k) {
    if (k == n)
      for (int j = n - 1; j >= 0; --j) d[j] = v[j] - (j > 0 ? v[j - 1] : 0.0);   // M_n' = M_{n-1}(x) - M_{n-1}(x - 1)
    const double r = 1.0 / (double)(k - 1);
    for (int j = k - 1; j >= 1; --j) v[j] = ((w + j) * v[j] + ((double)k - w - j) * v[j - 1]) * r;
    v[0] = w * v[0] * r;
  }
}

struct AtomSpl {
  int idx[3];
  double q;
  double v[3][SPME_MAXN], d[3][SPME_MAXN];
  // periodic wrap of the footprint, resolved once per atom and dimension instead of once per grid point (an integer modulo per
  // point made both kernels issue-bound: 1125 warp instructions per atom in the gather): element offsets of plane ix, row iy, iz
  long long ox[SPME_MAXN];
  int oy[SPME_MAXN], oz[SPME_MAXN];
};

// threads 0..2 of an atom's group fill its three spline sets
__device__ __forceinline__ int wrapk(int i, int K) { i %= K; return i < 0 ? i + K : i; }

__device__ __forceinline__ void atom_splines(const SpmeP& P, const double4& p, int dim, AtomSpl& s) {
  const double u = (double)P.K[dim] * (P.rc[dim] * p.x + P.rc[dim + 3] * p.y + P.rc[dim + 6] * p.z + 0.5);   // ewald_spole.F90:311-318
  const double t = trunc(u);
  s.idx[dim] = (int)t;
  bspline_fill(u - t, P.n, s.v[dim], s.d[dim]);
  int w = wrapk((int)t, P.K[dim]);
  for (int j = 0; j < P.n; ++j) {   // grid index Int(u) - j, wrapped: one step down at a time
    if (dim == 0) s.ox[j] = (long long)w * P.K[1] * P.K[2];
    else if (dim == 1) s.oy[j] = w * P.K[2];
    else s.oz[j] = w;
    w = w == 0 ? P.K[dim] - 1 : w - 1;
  }
  if (dim == 0) s.q = p.w;
}

constexpr int SPME_GROUP = 64;      // threads per atom: an n x n face of its footprint (n <= 8), looped for larger orders
constexpr int SPME_NG = 4;          // groups per block
constexpr int SPME_AGP = 8;         // atoms a group takes per block pass: the spline sets of all SPME_NG * SPME_AGP atoms of a pass are
constexpr int SPME_APB = SPME_NG * SPME_AGP;   // ... filled together (one thread per atom and dimension), then each group walks its atoms

// spme_construct_charge_array: Q(j,k,l) += q vx vy vz over the n^3 footprint (real grid, z fastest); also sum q^2 (self interaction)
template <int NF>   // NF: the spline order when it is known at compile time (8, the default of the reference), 0: P.n
__global__ void __launch_bounds__(SPME_GROUP * SPME_NG)
k_spme_spread(SpmeP P, const double4* __restrict__ posq, double* __restrict__ rgrid, double* __restrict__ totals) {
  const int n = NF ? NF : P.n;
  __shared__ AtomSpl s_a[SPME_APB];
  const int g = threadIdx.x / SPME_GROUP, t = threadIdx.x % SPME_GROUP;
  double q2 = 0.0;
  for (int base = blockIdx.x * SPME_APB; base < P.natms; base += gridDim.x * SPME_APB) {
    __syncthreads();
    if (threadIdx.x < 3 * SPME_APB) {
      const int a = base + threadIdx.x / 3;
      if (a < P.natms) atom_splines(P, posq[a], threadIdx.x % 3, s_a[threadIdx.x / 3]);
    }
    __syncthreads();
    for (int w = 0; w < SPME_AGP; ++w) {
      const int sl = g * SPME_AGP + w, a = base + sl;
      if (a >= P.natms) break;
      const AtomSpl& s = s_a[sl];
      if (t == 0) q2 += s.q * s.q;
      if (!(fabs(s.q) > ZERO_PLUS)) continue;                                  // ewald_general.F90:536
      for (int f = t; f < n * n; f += SPME_GROUP) {
        const int py = f / n, pz = f % n;
        double* const col = rgrid + (s.oy[py] + s.oz[pz]);
        const double fyz = s.q * s.v[2][pz] * s.v[1][py];
#pragma unroll
        for (int px = 0; px < n; ++px) atomicAdd(col + s.ox[px], fyz * s.v[0][px]);
      }
    }
  }
  // sum of q^2: one add per block
  __shared__ double s_q2[SPME_NG];
  if (t == 0) s_q2[g] = q2;
  __syncthreads();
  if (threadIdx.x == 0) { double v = 0.0; for (int k = 0; k < SPME_NG; ++k) v += s_q2[k]; atomicAdd(&totals[10], v); }
}

__global__ void k_spme_real_to_complex(size_t n, const double* __restrict__ r, double2* __restrict__ c) {
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) c[e] = make_double2(r[e], 0.0);
}
__global__ void k_spme_complex_to_real(size_t n, const double2* __restrict__ c, double* __restrict__ r) {   // extended_potential_grid is Real
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) r[e] = c[e].x;
}

// spme_construct_potential_grid_coul between the two transforms: potential_component = B(m) S(m) exp(-x^2) / (sqrt(pi) x^2) inside
// the spherical cutoff, x = pi |m| / alpha, and the stress kernel sum m_a m_b Re[comp (-2 (1 + x^2) / m^2) conj(S)]
// HALF: the grid is the non-redundant half of the transform of a real array (l = 0 .. K3 / 2, the layout of cufftExecD2Z): an element
// with 0 < l < K3 / 2 (or l = (K3 - 1) / 2 ... for odd K3) stands for itself and its conjugate partner, whose kernel value and stress
// term are the same in an orthogonal cell, so its stress term counts twice
template <bool HALF>
__global__ void k_spme_influence(SpmeP P, double conv, double test_fac, double cut2, const double* __restrict__ norm2, int kmax,
                                 double2* __restrict__ grid, double* __restrict__ totals) {
  const int K2l = HALF ? P.K[2] / 2 + 1 : P.K[2];   // stored extent of the fastest dimension
  const size_t ntot = (size_t)P.K[0] * P.K[1] * K2l;
  double st[6] = {0, 0, 0, 0, 0, 0};
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < ntot; e += (size_t)gridDim.x * blockDim.x) {
    const int l = (int)(e % K2l), k = (int)((e / K2l) % P.K[1]), j = (int)(e / ((size_t)K2l * P.K[1]));
    const int jj = 2 * j > P.K[0] ? j - P.K[0] : j, kk = 2 * k > P.K[1] ? k - P.K[1] : k, ll = 2 * l > P.K[2] ? l - P.K[2] : l;
    // recip_pos = jj rcell(1:9:3) + kk rcell(2:9:3) + ll rcell(3:9:3)
    const double m0 = jj * P.rc[0] + kk * P.rc[1] + ll * P.rc[2], m1 = jj * P.rc[3] + kk * P.rc[4] + ll * P.rc[5],
                 m2 = jj * P.rc[6] + kk * P.rc[7] + ll * P.rc[8];
    const double k2 = m0 * m0 + m1 * m1 + m2 * m2;
    double2 c = make_double2(0.0, 0.0);
    if (k2 <= cut2 && k2 > test_fac) {
      const double2 S = grid[e];
      const double x2 = k2 * conv * conv;
      const double fac = norm2[j] * norm2[kmax + k] * norm2[2 * kmax + l] * exp(-x2) / (SQRPI * x2);
      c = make_double2(fac * S.x, fac * S.y);
      const double pv = (c.x * S.x + c.y * S.y) * (-2.0 * ((1.0 + x2) / k2));
      st[0] += m0 * m0 * pv; st[1] += m0 * m1 * pv; st[2] += m0 * m2 * pv; st[3] += m1 * m1 * pv; st[4] += m1 * m2 * pv; st[5] += m2 * m2 * pv;
      if (HALF && l != 0 && 2 * l != P.K[2]) {
        // the conjugate partner (K1 - j, K2 - k, K3 - l) that the half spectrum leaves out: same kernel value, its own m vector -- a
        // Nyquist index keeps + K / 2 (the reference's 2 j > K test), every other component changes sign
        const int jp = 2 * j == P.K[0] ? jj : -jj, kp = 2 * k == P.K[1] ? kk : -kk, lp = -ll;
        const double p0 = jp * P.rc[0] + kp * P.rc[1] + lp * P.rc[2], p1 = jp * P.rc[3] + kp * P.rc[4] + lp * P.rc[5],
                     p2 = jp * P.rc[6] + kp * P.rc[7] + lp * P.rc[8];
        st[0] += p0 * p0 * pv; st[1] += p0 * p1 * pv; st[2] += p0 * p2 * pv; st[3] += p1 * p1 * pv; st[4] += p1 * p2 * pv; st[5] += p2 * p2 * pv;
      }
    }
    grid[e] = c;
  }
  __shared__ double red[8][6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 6; ++q) {
    double v = st[q];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(DLP_FULL, v, d);
    if (lane == 0) red[warp][q] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) v += red[w][threadIdx.x];
    atomicAdd(&totals[4 + threadIdx.x], v);
  }
}

// spme_calc_force_energy: per atom, energy and the three force sums over its footprint of the (real) potential grid
// (one warp per atom, two columns per lane and no cross-warp stage, measured slower: 2.89 against 2.56 ms per call)
template <int NF>
__global__ void __launch_bounds__(SPME_GROUP * SPME_NG)
k_spme_gather(SpmeP P, double kmx, double kmy, double kmz, const double4* __restrict__ posq, const double* __restrict__ rgrid,
              double* __restrict__ fraw, double* __restrict__ totals) {
  const int n = NF ? NF : P.n;
  __shared__ AtomSpl s_a[SPME_APB];
  __shared__ double s_red[SPME_NG][SPME_GROUP / 32][4];
  const int g = threadIdx.x / SPME_GROUP, t = threadIdx.x % SPME_GROUP;
  double te = 0.0, tf0 = 0.0, tf1 = 0.0, tf2 = 0.0;   // block totals, kept by thread 0 of each group
  for (int base = blockIdx.x * SPME_APB; base < P.natms; base += gridDim.x * SPME_APB) {
    __syncthreads();
    if (threadIdx.x < 3 * SPME_APB) {
      const int a = base + threadIdx.x / 3;
      if (a < P.natms) atom_splines(P, posq[a], threadIdx.x % 3, s_a[threadIdx.x / 3]);
    }
    __syncthreads();
    for (int w = 0; w < SPME_AGP; ++w) {   // the two warps of a group stay together through the named barrier below
      const int sl = g * SPME_AGP + w, a = base + sl;
      const AtomSpl& s = s_a[sl];
      const bool live = a < P.natms && fabs(s.q) > ZERO_PLUS;                  // ewald_general.F90:779
      double e = 0.0, f0 = 0.0, f1 = 0.0, f2 = 0.0;
      if (live) {
        for (int f = t; f < n * n; f += SPME_GROUP) {
          const int py = f / n, pz = f % n;
          const double* const col = rgrid + (s.oy[py] + s.oz[pz]);
          const double y0 = s.v[1][py], z0 = s.v[2][pz], y1 = s.d[1][py], z1 = s.d[2][pz];
          double sx0 = 0.0, sx1 = 0.0;                                         // sum over x of phi vx, phi vx'
#pragma unroll
          for (int px = 0; px < n; ++px) {
            const double phi = col[s.ox[px]];
            sx0 += phi * s.v[0][px]; sx1 += phi * s.d[0][px];
          }
          e += y0 * z0 * sx0;
          f0 += y0 * z0 * sx1; f1 += y1 * z0 * sx0; f2 += y0 * z1 * sx0;
        }
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        e += __shfl_xor_sync(DLP_FULL, e, d); f0 += __shfl_xor_sync(DLP_FULL, f0, d);
        f1 += __shfl_xor_sync(DLP_FULL, f1, d); f2 += __shfl_xor_sync(DLP_FULL, f2, d);
      }
      if ((t & 31) == 0) { double* r = s_red[g][t >> 5]; r[0] = e; r[1] = f0; r[2] = f1; r[3] = f2; }
      asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(SPME_GROUP) : "memory");
      if (t == 0 && a < P.natms) {
        double r[4] = {0, 0, 0, 0};
        for (int ww = 0; ww < SPME_GROUP / 32; ++ww) for (int c = 0; c < 4; ++c) r[c] += s_red[g][ww][c];
        // curr_force_temp = q (sum) recip_kmax; forces(:, i) = -curr_force_temp; force_total -= curr_force_temp
        const double c0 = s.q * r[1] * kmx, c1 = s.q * r[2] * kmy, c2 = s.q * r[3] * kmz;
        fraw[a] = live ? -c0 : 0.0; fraw[(size_t)P.natms + a] = live ? -c1 : 0.0; fraw[2 * (size_t)P.natms + a] = live ? -c2 : 0.0;
        if (live) { te += s.q * r[0]; tf0 -= c0; tf1 -= c1; tf2 -= c2; }
      }
      asm volatile("bar.sync %0, %1;" ::"r"(g + 1), "r"(SPME_GROUP) : "memory");
    }
  }
  __syncthreads();
  if (t == 0) { double* r = s_red[g][0]; r[0] = te; r[1] = tf0; r[2] = tf1; r[3] = tf2; }
  __syncthreads();
  if (threadIdx.x < 4) {
    double v = 0.0;
    for (int k = 0; k < SPME_NG; ++k) v += s_red[k][0][threadIdx.x];
    atomicAdd(&totals[threadIdx.x], v);
  }
}

// forces = (fraw - force_total / megatm) * scale * 2, added to the force arrays (ewald_general.F90:862-866, ewald_spole.F90:437-446)
__global__ void k_spme_finish(int natms, double inv_megatm, double scale2, const double* __restrict__ fraw,
                              const double* __restrict__ totals, double* fx, double* fy, double* fz) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= natms) return;
  fx[i] += (fraw[i] - totals[1] * inv_megatm) * scale2;
  fy[i] += (fraw[(size_t)natms + i] - totals[2] * inv_megatm) * scale2;
  fz[i] += (fraw[2 * (size_t)natms + i] - totals[3] * inv_megatm) * scale2;
}

// numerics.F90 invert: rcell = adjugate / determinant, laid out so that s_d = rcell(d) x + rcell(d + 3) y + rcell(d + 6) z
void invert9(const double* a0, double* b0, double* det) {
  const double* a = a0 - 1;
  double* b = b0 - 1;
  b[1] = a[5] * a[9] - a[6] * a[8]; b[2] = a[3] * a[8] - a[2] * a[9]; b[3] = a[2] * a[6] - a[3] * a[5];
  b[4] = a[6] * a[7] - a[4] * a[9]; b[5] = a[1] * a[9] - a[3] * a[7]; b[6] = a[3] * a[4] - a[1] * a[6];
  b[7] = a[4] * a[8] - a[5] * a[7]; b[8] = a[2] * a[7] - a[1] * a[8]; b[9] = a[1] * a[5] - a[2] * a[4];
  const double d = a[1] * b[1] + a[4] * b[2] + a[7] * b[3];
  const double r = std::fabs(d) > 0.0 ? 1.0 / d : 0.0;
  for (int i = 1; i <= 9; ++i) b[i] = r * b[i];
  *det = d;
}

// ---- cuFFT, loaded on first use
typedef int (*fn_plan3d)(int*, int, int, int, int);
typedef int (*fn_setstream)(int, cudaStream_t);
typedef int (*fn_execz2z)(int, void*, void*, int);
typedef int (*fn_execr)(int, void*, void*);
typedef int (*fn_destroy)(int);
struct CufftApi { void* lib = nullptr; fn_plan3d plan3d = nullptr; fn_setstream set_stream = nullptr; fn_execz2z exec = nullptr; fn_destroy destroy = nullptr;
                  fn_execr exec_d2z = nullptr, exec_z2d = nullptr; };
CufftApi g_fft;
constexpr int CUFFT_Z2Z_TYPE = 0x69, CUFFT_D2Z_TYPE = 0x6a, CUFFT_Z2D_TYPE = 0x6c, CUFFT_FWD = -1, CUFFT_INV = 1;

bool load_cufft() {
  if (g_fft.lib) return true;
  const char* names[] = {"libcufft.so.11", "libcufft.so", "/usr/local/cuda/lib64/libcufft.so.11", "/usr/local/cuda/lib64/libcufft.so", "libcufft.so.12"};
  for (const char* nm : names) { g_fft.lib = dlopen(nm, RTLD_NOW | RTLD_LOCAL); if (g_fft.lib) break; }
  if (!g_fft.lib) return false;
  g_fft.plan3d = (fn_plan3d)dlsym(g_fft.lib, "cufftPlan3d");
  g_fft.set_stream = (fn_setstream)dlsym(g_fft.lib, "cufftSetStream");
  g_fft.exec = (fn_execz2z)dlsym(g_fft.lib, "cufftExecZ2Z");
  g_fft.destroy = (fn_destroy)dlsym(g_fft.lib, "cufftDestroy");
  g_fft.exec_d2z = (fn_execr)dlsym(g_fft.lib, "cufftExecD2Z");
  g_fft.exec_z2d = (fn_execr)dlsym(g_fft.lib, "cufftExecZ2D");
  return g_fft.plan3d && g_fft.set_stream && g_fft.exec && g_fft.destroy;
}

}  // namespace

void dlp_spme_release(dlpgpu_ctx* ctx) {
  if (ctx->spme_plan_valid && g_fft.destroy) g_fft.destroy(ctx->spme_plan);
  ctx->spme_plan_valid = false;
  if (ctx->spme_r2c_valid && g_fft.destroy) { g_fft.destroy(ctx->spme_plan_d2z); g_fft.destroy(ctx->spme_plan_z2d); }
  ctx->spme_r2c_valid = false;
  ctx->spme_grid.release(ctx->stream); ctx->spme_rgrid.release(ctx->stream); ctx->spme_norm2.release(ctx->stream); ctx->spme_fraw.release(ctx->stream); ctx->spme_tot.release(ctx->stream);
}

extern "C" {

int dlpgpu_set_spme(dlpgpu_ctx* ctx, const int kdim[3], int nsplines) {
  if (!ctx || !kdim) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (nsplines < 3 || nsplines > SPME_MAXN) return dlp_fail(ctx, DLPGPU_ERR_ARG, "set_spme: spline order %d outside 3..%d", nsplines, SPME_MAXN);
  for (int d = 0; d < 3; ++d)
    if (kdim[d] < nsplines || kdim[d] > 2048) return dlp_fail(ctx, DLPGPU_ERR_ARG, "set_spme: grid dimension %d not supported", kdim[d]);
  if (!load_cufft()) return dlp_fail(ctx, DLPGPU_ERR_STATE, "set_spme: cuFFT (libcufft.so) could not be loaded");
  if (ctx->spme_plan_valid) { g_fft.destroy(ctx->spme_plan); ctx->spme_plan_valid = false; }
  if (ctx->spme_r2c_valid) { g_fft.destroy(ctx->spme_plan_d2z); g_fft.destroy(ctx->spme_plan_z2d); ctx->spme_r2c_valid = false; }
  for (int d = 0; d < 3; ++d) ctx->spme_k[d] = kdim[d];
  ctx->spme_n = nsplines;
  const int n = nsplines, kmax = std::max(kdim[0], std::max(kdim[1], kdim[2]));
  // bspline_coeffs_gen: cspline(k + 2) = M_n(k + 1), b(i) = w^{i (n-1)} / sum_k cspline(k + 2) w^{i k}, norm2 = |b|^2
  std::vector<double> cs(n + 1, 0.0);
  cs[2] = 1.0;
  for (int k = 3; k <= n; ++k)
    for (int j = k; j >= 2; --j) cs[j] = ((double)(j - 1) * cs[j] + (double)(k - j + 1) * cs[j - 1]) / (double)(k - 1);
  std::vector<double> nrm((size_t)3 * kmax, 0.0);
  const double twopi = 6.283185307179586476925287;
  for (int d = 0; d < 3; ++d) {
    const int K = kdim[d];
    for (int i = 0; i < K; ++i) {
      double re = 0.0, im = 0.0;
      for (int k = 0; k <= n - 2; ++k) {
        const double arg = twopi * (double)((long long)i * k % K) / (double)K;
        re += cs[k + 2] * std::cos(arg); im += cs[k + 2] * std::sin(arg);
      }
      nrm[(size_t)d * kmax + i] = 1.0 / (re * re + im * im);   // |w^{..}| = 1
    }
  }
  CK(ctx->spme_norm2.ensure(nrm.size(), ctx->stream));
  CK(cudaMemcpyAsync(ctx->spme_norm2.p, nrm.data(), nrm.size() * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->spme_kmax = kmax;
  const size_t ntot = (size_t)kdim[0] * kdim[1] * kdim[2];
  CK(ctx->spme_grid.ensure(2 * ntot, ctx->stream));
  CK(ctx->spme_rgrid.ensure(ntot, ctx->stream));
  CK(ctx->spme_tot.ensure(16, ctx->stream));
  int plan = 0;
  if (g_fft.plan3d(&plan, kdim[0], kdim[1], kdim[2], CUFFT_Z2Z_TYPE) != 0) return dlp_fail(ctx, DLPGPU_ERR_CUDA, "set_spme: cufftPlan3d failed");
  if (g_fft.set_stream(plan, ctx->stream) != 0) return dlp_fail(ctx, DLPGPU_ERR_CUDA, "set_spme: cufftSetStream failed");
  ctx->spme_plan = plan; ctx->spme_plan_valid = true;
  return 0;
}

}  // extern "C"

// ---- the stages of ewald_spme_forces_coul.  One domain runs them back to back (dlpgpu_dev_spme_forces); several domains keep a
// REPLICATED grid: every rank spreads its own atoms onto a grid of the whole cell, the host side sums the grids over the ranks
// (one all-reduce over NVLink: 46 MB at 180^3), every rank transforms the whole grid itself (0.3 ms at 180^3 -- less than the
// transposes of a distributed transform would cost on one NVSwitch node, and 180 GB of HBM hold the grid many times over) and
// gathers the forces of its own atoms; what the reference exchanges through exchange_grid and its parallel DaFT transform
// (ewald_spole.F90:336-420, parallel_fft.F90) shrinks to that one sum plus the three doubles of the net force.
namespace {
struct SpmeGeom { SpmeP P; double inv[9], det, scale, kmx, kmy, kmz, cut2, conv, test_fac; size_t ntot; int blocks_a, blocks_g; };

int spme_geometry(dlpgpu_ctx* ctx, const char* who, SpmeGeom& G) {
  if (!ctx->spme_plan_valid) return dlp_fail(ctx, DLPGPU_ERR_STATE, "%s: call dlpgpu_set_spme first", who);
  if (!ctx->ew_on) return dlp_fail(ctx, DLPGPU_ERR_STATE, "%s: no Ewald parameters (dlpgpu_set_ewald)", who);
  SpmeP& P = G.P;
  for (int d = 0; d < 3; ++d) P.K[d] = ctx->spme_k[d];
  P.n = ctx->spme_n; P.natms = ctx->natms;
  invert9(ctx->cell, G.inv, &G.det);
  if (std::fabs(G.det) < 1.0e-6) return dlp_fail(ctx, 120, "%s: singular cell", who);
  for (int k = 0; k < 9; ++k) P.rc[k] = G.inv[k];
  G.ntot = (size_t)P.K[0] * P.K[1] * P.K[2];
  G.blocks_a = std::max(1, std::min(cdiv(std::max(ctx->natms, 1), SPME_APB), ctx->sm_count * 8));
  G.blocks_g = std::max(1, std::min(cdiv((long long)G.ntot, 256), ctx->sm_count * 16));
  // the spherical cutoff of the reference: 0.525 min_d(K_d * width_d of the reciprocal cell), ewald_spole.F90:1287-1291
  double w[3];
  {
    const double* a = G.inv; const double* b = G.inv + 3; const double* c = G.inv + 6;   // rows of rcell as lattice vectors (dcell)
    auto cross = [](const double* u, const double* v, double* o) { o[0] = u[1] * v[2] - u[2] * v[1]; o[1] = u[2] * v[0] - u[0] * v[2]; o[2] = u[0] * v[1] - u[1] * v[0]; };
    auto nrm3 = [](const double* u) { return std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]); };
    double bc[3], ca[3], ab[3];
    cross(b, c, bc); cross(c, a, ca); cross(a, b, ab);
    const double vol = std::fabs(a[0] * bc[0] + a[1] * bc[1] + a[2] * bc[2]);
    w[0] = vol / nrm3(bc); w[1] = vol / nrm3(ca); w[2] = vol / nrm3(ab);
  }
  const double cut = 0.5 * 1.05 * std::min(P.K[0] * w[0], std::min(P.K[1] * w[1], P.K[2] * w[2]));
  const double pi = 3.14159265358979323846264338327950288;
  G.cut2 = cut * cut;
  G.conv = pi / ctx->alpha; G.test_fac = (1.0e-6 / G.conv) * (1.0e-6 / G.conv);
  // recip_kmax = Matmul(Reshape(rcell, [3, 3]), k_vec_dim_real): component a = sum_b rcell(a + 3 (b - 1)) K_b  (ewald_general.F90:755-756)
  G.kmx = G.inv[0] * P.K[0] + G.inv[3] * P.K[1] + G.inv[6] * P.K[2]; G.kmy = G.inv[1] * P.K[0] + G.inv[4] * P.K[1] + G.inv[7] * P.K[2];
  G.kmz = G.inv[2] * P.K[0] + G.inv[5] * P.K[1] + G.inv[8] * P.K[2];
  G.scale = pi * SQRPI * (1.0 / (ctx->alpha * ctx->alpha)) * (0.5 / std::fabs(G.det)) * ctx->scaling;   // ewald_spole.F90:299, pot_order 1
  return 0;
}

// spme_construct_charge_array of this rank's atoms onto rgrid (zeroed here); sum q^2 of the rank goes to the totals
int spme_spread(dlpgpu_ctx* ctx, const SpmeGeom& G, double* rgrid) {
  cudaStream_t s = ctx->stream;
  CK(cudaMemsetAsync(rgrid, 0, G.ntot * sizeof(double), s));
  CK(cudaMemsetAsync(ctx->spme_tot.p, 0, 16 * sizeof(double), s));
  if (ctx->natms > 0) {
    if (G.P.n == 8) LAUNCH(ctx, k_spme_spread<8>, G.blocks_a, SPME_GROUP * SPME_NG, 0, G.P, ctx->posq.p, rgrid, ctx->spme_tot.p);
    else LAUNCH(ctx, k_spme_spread<0>, G.blocks_a, SPME_GROUP * SPME_NG, 0, G.P, ctx->posq.p, rgrid, ctx->spme_tot.p);
  }
  return 0;
}
// charge grid of the WHOLE system -> potential grid (in place); the stress kernel sums go to the totals.  Orthogonal cells take the
// real-to-complex transforms (half the spectrum: the kernel is even in m there, Nyquist planes included -- a triclinic cell breaks
// that on the Nyquist planes, where the reference takes + K / 2 for both members of a conjugate pair, so it keeps the complex path)
int spme_solve(dlpgpu_ctx* ctx, const SpmeGeom& G, double* rgrid) {
  double2* grid = reinterpret_cast<double2*>(ctx->spme_grid.p);
  const double* c = ctx->cell;
  const bool ortho = c[1] == 0.0 && c[2] == 0.0 && c[3] == 0.0 && c[5] == 0.0 && c[6] == 0.0 && c[7] == 0.0;
  if (ortho && g_fft.exec_d2z && g_fft.exec_z2d) {
    if (!ctx->spme_r2c_valid) {
      int p1 = 0, p2 = 0;
      if (g_fft.plan3d(&p1, G.P.K[0], G.P.K[1], G.P.K[2], CUFFT_D2Z_TYPE) != 0) return dlp_fail(ctx, DLPGPU_ERR_CUDA, "spme: cufftPlan3d (D2Z) failed");
      if (g_fft.plan3d(&p2, G.P.K[0], G.P.K[1], G.P.K[2], CUFFT_Z2D_TYPE) != 0) { g_fft.destroy(p1); return dlp_fail(ctx, DLPGPU_ERR_CUDA, "spme: cufftPlan3d (Z2D) failed"); }
      if (g_fft.set_stream(p1, ctx->stream) != 0 || g_fft.set_stream(p2, ctx->stream) != 0) { g_fft.destroy(p1); g_fft.destroy(p2); return dlp_fail(ctx, DLPGPU_ERR_CUDA, "spme: cufftSetStream failed"); }
      ctx->spme_plan_d2z = p1; ctx->spme_plan_z2d = p2; ctx->spme_r2c_valid = true;
    }
    const size_t nhalf = (size_t)G.P.K[0] * G.P.K[1] * (G.P.K[2] / 2 + 1);
    const int blocks_h = std::max(1, std::min(cdiv((long long)nhalf, 256), ctx->sm_count * 16));
    if (g_fft.exec_d2z(ctx->spme_plan_d2z, rgrid, grid) != 0) return dlp_fail(ctx, DLPGPU_ERR_CUDA, "spme: forward FFT (D2Z) failed");
    LAUNCH(ctx, k_spme_influence<true>, blocks_h, 256, 0, G.P, G.conv, G.test_fac, G.cut2, ctx->spme_norm2.p, ctx->spme_kmax, grid, ctx->spme_tot.p);
    if (g_fft.exec_z2d(ctx->spme_plan_z2d, grid, rgrid) != 0) return dlp_fail(ctx, DLPGPU_ERR_CUDA, "spme: backward FFT (Z2D) failed");
    return 0;
  }
  LAUNCH(ctx, k_spme_real_to_complex, G.blocks_g, 256, 0, G.ntot, rgrid, grid);
  if (g_fft.exec(ctx->spme_plan, grid, grid, CUFFT_FWD) != 0) return dlp_fail(ctx, DLPGPU_ERR_CUDA, "spme: forward FFT failed");
  LAUNCH(ctx, k_spme_influence<false>, G.blocks_g, 256, 0, G.P, G.conv, G.test_fac, G.cut2, ctx->spme_norm2.p, ctx->spme_kmax, grid, ctx->spme_tot.p);
  if (g_fft.exec(ctx->spme_plan, grid, grid, CUFFT_INV) != 0) return dlp_fail(ctx, DLPGPU_ERR_CUDA, "spme: backward FFT failed");
  LAUNCH(ctx, k_spme_complex_to_real, G.blocks_g, 256, 0, G.ntot, grid, rgrid);
  return 0;
}
// spme_calc_force_energy for this rank's atoms: raw forces into spme_fraw, energy and raw force total of the rank into the totals
int spme_gather(dlpgpu_ctx* ctx, const SpmeGeom& G, const double* rgrid) {
  CK(ctx->spme_fraw.ensure((size_t)3 * std::max(ctx->natms, 1), ctx->stream));
  if (ctx->natms > 0) {
    if (G.P.n == 8)
      LAUNCH(ctx, k_spme_gather<8>, G.blocks_a, SPME_GROUP * SPME_NG, 0, G.P, G.kmx, G.kmy, G.kmz, ctx->posq.p, rgrid, ctx->spme_fraw.p, ctx->spme_tot.p);
    else
      LAUNCH(ctx, k_spme_gather<0>, G.blocks_a, SPME_GROUP * SPME_NG, 0, G.P, G.kmx, G.kmy, G.kmz, ctx->posq.p, rgrid, ctx->spme_fraw.p, ctx->spme_tot.p);
  }
  return 0;
}
// forces += (raw - net force of ALL atoms / megatm) * 2 scale; sums of this rank.  ftot: the net raw force over all ranks (nullptr: the
// rank's own, i.e. one domain); stress_share: the k-space stress sums were formed from the whole grid on every rank, each reports 1 / nranks
int spme_finish(dlpgpu_ctx* ctx, const SpmeGeom& G, int megatm, const double* ftot, double stress_share, double out[16]) {
  cudaStream_t s = ctx->stream;
  if (ftot) CK(cudaMemcpyAsync(ctx->spme_tot.p + 1, ftot, 3 * sizeof(double), cudaMemcpyHostToDevice, s));
  if (ctx->natms > 0)
    LAUNCH(ctx, k_spme_finish, cdiv(ctx->natms, 256), 256, 0, ctx->natms, 1.0 / (double)megatm, G.scale * 2.0, ctx->spme_fraw.p, ctx->spme_tot.p,
           ctx->fx.p, ctx->fy.p, ctx->fz.p);
  double t[16];
  CK(cudaMemcpyAsync(t, ctx->spme_tot.p, 16 * sizeof(double), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  CK(cudaGetLastError());
  for (int k = 0; k < 16; ++k) out[k] = 0.0;
  const double eng = t[0] * G.scale;
  const double self = -t[10] * ctx->scaling * ctx->alpha / SQRPI;            // spme.F90:220-224
  // stress_temp(beta, alpha) column-major, symmetric: [xx xy xz / xy yy yz / xz yz zz] * scale, diagonal += eng (:448-451)
  const double sx[9] = {t[4], t[5], t[6], t[5], t[7], t[8], t[6], t[8], t[9]};
  for (int k = 0; k < 9; ++k) out[2 + k] = sx[k] * G.scale * stress_share + ((k % 4 == 0) ? eng : 0.0);
  out[0] = eng + self;
  out[1] = -(out[2] + out[6] + out[10]);
  out[11] = eng; out[12] = self;
  return 0;
}
}  // namespace

extern "C" {

// out[0] = engcpe_rc (reciprocal energy + self interaction), out[1] = vircpe_rc, out[2..10] = the nine stress contributions
// (stats%stress += ...), out[11] = the reciprocal energy alone, out[12] = the self interaction
int dlpgpu_dev_spme_forces(dlpgpu_ctx* ctx, int megatm, double out[16]) {
  if (!ctx || !out || megatm < 1) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  if (ctx->nx * ctx->ny * ctx->nz > 1)
    return dlp_fail(ctx, DLPGPU_ERR_STATE, "spme_forces: one call serves one domain (mxnode = 1); several domains run the stages dlpgpu_dev_spme_spread / "
                                           "_solve / _gather / _finish around a sum of the charge grids");
  SpmeGeom G{};
  CKRC(spme_geometry(ctx, "spme_forces", G));
  double* rgrid = ctx->spme_rgrid.p;   // the charge grid, later the real potential grid
  CKRC(spme_spread(ctx, G, rgrid));
  CKRC(spme_solve(ctx, G, rgrid));
  CKRC(spme_gather(ctx, G, rgrid));
  return spme_finish(ctx, G, megatm, nullptr, 1.0, out);
}

// the stages for several domains (replicated grid, see above).  grid_dev: K1 K2 K3 doubles of DEVICE memory owned by the caller
// (z fastest), so that the host side can sum it over the ranks with whatever collective it has (NCCL all-reduce, peer copies).
int dlpgpu_dev_spme_spread(dlpgpu_ctx* ctx, double* grid_dev) {
  if (!ctx || !grid_dev) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  SpmeGeom G{};
  CKRC(spme_geometry(ctx, "spme_spread", G));
  return spme_spread(ctx, G, grid_dev);
}
// grid_dev = the SUM of the ranks' charge grids; on return the potential grid, and ftot_local = the rank's raw net force
int dlpgpu_dev_spme_solve_gather(dlpgpu_ctx* ctx, double* grid_dev, double ftot_local[3]) {
  if (!ctx || !grid_dev || !ftot_local) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  SpmeGeom G{};
  CKRC(spme_geometry(ctx, "spme_solve_gather", G));
  CKRC(spme_solve(ctx, G, grid_dev));
  CKRC(spme_gather(ctx, G, grid_dev));
  CK(cudaMemcpyAsync(ftot_local, ctx->spme_tot.p + 1, 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
// ftot_global = sum of ftot_local over the ranks; out as for dlpgpu_dev_spme_forces, this rank's share (the ranks' outs add up)
int dlpgpu_dev_spme_finish(dlpgpu_ctx* ctx, int megatm, const double ftot_global[3], int nranks, double out[16]) {
  if (!ctx || !ftot_global || !out || megatm < 1 || nranks < 1) return DLPGPU_ERR_ARG;
  CK(cudaSetDevice(ctx->device));
  SpmeGeom G{};
  CKRC(spme_geometry(ctx, "spme_finish", G));
  return spme_finish(ctx, G, megatm, ftot_global, 1.0 / (double)nranks, out);
}

}  // extern "C"

int dlp_preload_spme() {
  const void* ks[] = {(const void*)k_spme_spread<8>, (const void*)k_spme_spread<0>, (const void*)k_spme_influence<true>, (const void*)k_spme_influence<false>, (const void*)k_spme_gather<8>, (const void*)k_spme_gather<0>, (const void*)k_spme_finish, (const void*)k_spme_real_to_complex,
                      (const void*)k_spme_complex_to_real};
  cudaFuncAttributes a;
  for (const void* k : ks) if (cudaFuncGetAttributes(&a, k) != cudaSuccess) cudaGetLastError();
  return 0;
}
