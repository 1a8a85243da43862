// ubench.cu -- B200 micro-benchmarks behind the pair-kernel design decisions in DESIGN.md (development tool, not product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench.bin scripts/ubench.cu
// Measures: global RED.ADD.F64 throughput (random / windowed addresses), shared-memory fp64 CAS-add throughput,
// random LDS.128 throughput, windowed 32-byte gathers through L1/L2, DFMA peak.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// each thread: iters x 3 RED to f[3*j..], j = base + hash % window
__global__ void k_red(double* f, int n, int window, int iters) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  int warp = tid >> 5;
  long long base = ((long long)warp * 7) % (n - window);   // warps walk the array like consecutive atoms
  unsigned h = hash(tid + 12345);
  for (int it = 0; it < iters; ++it) {
    h = hash(h + it);
    int j = (int)(base + h % window);
    double v = 1e-9 * (h & 255);
    atomicAdd(&f[j], v);
    atomicAdd(&f[n + j], v);
    atomicAdd(&f[2 * n + j], v);
  }
}
// same with an interleaved xyz layout (3 consecutive doubles of one 32-byte slot)
__global__ void k_red4(double* f, int n, int window, int iters) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  int warp = tid >> 5;
  long long base = ((long long)warp * 7) % (n - window);
  unsigned h = hash(tid + 12345);
  for (int it = 0; it < iters; ++it) {
    h = hash(h + it);
    size_t j = (size_t)(base + h % window) * 4;
    double v = 1e-9 * (h & 255);
    atomicAdd(&f[j], v);
    atomicAdd(&f[j + 1], v);
    atomicAdd(&f[j + 2], v);
  }
}

__global__ void k_smem_cas(double* out, int window, int iters) {
  extern __shared__ double s[];
  for (int i = threadIdx.x; i < window; i += blockDim.x) s[i] = 0.0;
  __syncthreads();
  unsigned h = hash(blockIdx.x * blockDim.x + threadIdx.x + 777);
  for (int it = 0; it < iters; ++it) {
    h = hash(h + it);
    int j = h % (window / 3);
    double v = 1e-9 * (h & 255);
    atomicAdd(&s[3 * j], v);
    atomicAdd(&s[3 * j + 1], v);
    atomicAdd(&s[3 * j + 2], v);
  }
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x] = s[0] + s[window - 1];
}

__global__ void k_lds128(double* out, int entries, int iters, int stride_mode) {
  extern __shared__ double2 t[];
  for (int i = threadIdx.x; i < entries; i += blockDim.x) t[i] = make_double2(i, 2 * i);
  __syncthreads();
  unsigned h = hash(blockIdx.x * blockDim.x + threadIdx.x + 99);
  double a = 0, b = 0;
  for (int it = 0; it < iters; ++it) {
    h = hash(h + it);
    int l = stride_mode ? (int)((threadIdx.x * 7 + it * 131) % (entries - 3)) : (int)(h % (entries - 3));
    double2 x0 = t[l], x1 = t[l + 1], x2 = t[l + 2];
    a += x0.x + x1.x + x2.x;
    b += x0.y + x1.y + x2.y;
  }
  if (a + b == 1.2345) out[0] = a;
}

// 32-byte gathers: each warp reads j = base + hash % window (window in atoms)
__global__ void k_gather(const double4* __restrict__ p, double* out, int n, int window, int iters, int sorted_runs) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  int warp = tid >> 5, lane = tid & 31;
  long long base = ((long long)warp * 7) % (n - window - 64);
  unsigned h = hash(warp + 4242);
  double a = 0;
  for (int it = 0; it < iters; ++it) {
    h = hash(h + it);
    int j;
    if (sorted_runs) j = (int)(base + (h % window) + 2 * lane);   // one warp reads a strided run starting at a random point
    else j = (int)(base + hash(h + lane) % window);
    double4 v = p[j];
    a += v.x + v.y + v.z + v.w;
  }
  if (a == 1.2345) out[0] = a;
}

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


// RED in runs: the lanes of a warp hit consecutive (stride `st`) atoms of one random run, SoA components
__global__ void k_red_runs(double* f, int n, int st, int iters) {
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  int warp = tid >> 5, lane = tid & 31;
  unsigned h = hash(warp + 999);
  for (int it = 0; it < iters; ++it) {
    h = hash(h + it);
    int j = (int)(h % (n - 32 * st - 1)) + lane * st;
    double v = 1e-9 * (h & 255);
    atomicAdd(&f[j], v);
    atomicAdd(&f[n + j], v);
    atomicAdd(&f[2 * n + j], v);
  }
}
// random 16-byte table reads through L1 (LDG), optionally mixed with the same number of LDS reads
__global__ void k_ldg_tab(const double2* __restrict__ tab, double* out, int entries, int iters, int mix) {
  extern __shared__ double2 t[];
  if (mix) { for (int i = threadIdx.x; i < entries; i += blockDim.x) t[i] = tab[i]; __syncthreads(); }
  unsigned h = hash(blockIdx.x * blockDim.x + threadIdx.x + 99);
  double a = 0, b = 0;
  for (int it = 0; it < iters; ++it) {
    h = hash(h + it);
    int l = (int)(h % (entries - 3));
    double2 x0 = __ldg(&tab[l]), x1 = __ldg(&tab[l + 1]), x2 = __ldg(&tab[l + 2]);
    a += x0.x + x1.x + x2.x; b += x0.y + x1.y + x2.y;
    if (mix) {
      int m = (int)((h >> 7) % (entries - 3));
      double2 y0 = t[m], y1 = t[m + 1], y2 = t[m + 2];
      a += y0.x + y1.x + y2.x; b += y0.y + y1.y + y2.y;
    }
  }
  if (a + b == 1.2345) out[0] = a;
}
// LDS.128 with R-fold replication: copy c = lane % R of entry e lives at 16-byte unit e*R + c
__global__ void k_lds_rep(double* out, int entries, int R, int iters) {
  extern __shared__ double2 t[];
  for (int i = threadIdx.x; i < entries * R; i += blockDim.x) t[i] = make_double2(i, 2 * i);
  __syncthreads();
  unsigned h = hash(blockIdx.x * blockDim.x + threadIdx.x + 99);
  int c = threadIdx.x % R;
  double a = 0, b = 0;
  for (int it = 0; it < iters; ++it) {
    h = hash(h + it);
    int l = (int)(h % (entries - 3));
    double2 x0 = t[l * R + c], x1 = t[(l + 1) * R + c], x2 = t[(l + 2) * R + c];
    a += x0.x + x1.x + x2.x; b += x0.y + x1.y + x2.y;
  }
  if (a + b == 1.2345) out[0] = a;
}

template <class F>
float timeit(F f, int reps = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a);
    f();
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp pr;
  CK(cudaGetDeviceProperties(&pr, dev));
  int sms = pr.multiProcessorCount;
  printf("device %s  SMs %d  clock %d MHz  L2 %d MB\n", pr.name, sms, pr.clockRate / 1000, pr.l2CacheSize >> 20);
  const int n = 1 << 20;
  double* f; CK(cudaMalloc(&f, (size_t)4 * n * sizeof(double))); CK(cudaMemset(f, 0, (size_t)4 * n * sizeof(double)));
  double* out; CK(cudaMalloc(&out, 1 << 20));
  double4* p; CK(cudaMalloc(&p, (size_t)(n + 65536) * sizeof(double4))); CK(cudaMemset(p, 0, (size_t)(n + 65536) * sizeof(double4)));
  {
    int iters = 2000, blocks = sms * 8, threads = 256;
    float ms = timeit([&] { k_dfma<<<blocks, threads>>>(iters, out); });
    printf("DFMA peak: %.2f TFLOP/s\n", 2.0 * 8 * iters * (double)threads * blocks / (ms * 1e-3) / 1e12);
  }
  for (int window : {1000, 4000, 1 << 19}) {
    int iters = 64, blocks = sms * 16, threads = 256;
    float ms = timeit([&] { k_red<<<blocks, threads>>>(f, n, window, iters); });
    double cnt = 3.0 * iters * (double)threads * blocks;
    printf("global RED.F64 SoA  window %7d atoms: %.1f G atomics/s\n", window, cnt / (ms * 1e-3) / 1e9);
    ms = timeit([&] { k_red4<<<blocks, threads>>>(f, n, window, iters); });
    printf("global RED.F64 AoS4 window %7d atoms: %.1f G atomics/s\n", window, cnt / (ms * 1e-3) / 1e9);
  }
  for (int window : {1536, 6144}) {
    int iters = 256, blocks = sms * 2, threads = 512;
    size_t sm = window * sizeof(double);
    float ms = timeit([&] { k_smem_cas<<<blocks, threads, sm>>>(out, window, iters); });
    double cnt = 3.0 * iters * (double)threads * blocks;
    printf("shared fp64 CAS-add window %5d doubles: %.1f G atomics/s  (%.2f per clk per SM @1.9GHz)\n", window, cnt / (ms * 1e-3) / 1e9,
           cnt / (ms * 1e-3) / sms / 1.9e9);
  }
  for (int mode : {0, 1}) {
    int entries = 4 * 1205, iters = 512, blocks = sms, threads = 1024;
    size_t sm = entries * sizeof(double2);
    CK(cudaFuncSetAttribute(k_lds128, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    float ms = timeit([&] { k_lds128<<<blocks, threads, sm>>>(out, entries, iters, mode); });
    double cnt = 3.0 * iters * (double)threads * blocks;
    printf("LDS.128 %s: %.1f G loads/s  = %.2f warp-instr per clk per SM @1.9GHz (%.1f clk per warp LDS.128)\n", mode ? "strided" : "random ",
           cnt / (ms * 1e-3) / 1e9, cnt / 32 / (ms * 1e-3) / sms / 1.9e9, 1.0 / (cnt / 32 / (ms * 1e-3) / sms / 1.9e9));
  }
  for (int runs : {0, 1})
    for (int window : {1500, 20000, 1 << 19}) {
      int iters = 256, blocks = sms * 8, threads = 256;
      float ms = timeit([&] { k_gather<<<blocks, threads>>>(p, out, n, window, iters, runs); });
      double cnt = (double)iters * threads * blocks;
      printf("gather double4 %s window %7d: %.1f G loads/s = %.2f TB/s (%.1f clk per warp-load per SM)\n", runs ? "runs  " : "random", window,
             cnt / (ms * 1e-3) / 1e9, cnt * 32 / (ms * 1e-3) / 1e12, 1.0 / (cnt / 32 / (ms * 1e-3) / sms / 1.9e9));
    }

  for (int st : {1, 2, 4}) {
    int iters = 64, blocks = sms * 16, threads = 256;
    float ms = timeit([&] { k_red_runs<<<blocks, threads>>>(f, n, st, iters); });
    double cnt = 3.0 * iters * (double)threads * blocks;
    printf("global RED.F64 runs stride %d: %.1f G atomics/s\n", st, cnt / (ms * 1e-3) / 1e9);
  }
  {
    int entries = 1205;
    double2* tab; CK(cudaMalloc(&tab, entries * sizeof(double2))); CK(cudaMemset(tab, 0, entries * sizeof(double2)));
    for (int mix : {0, 1}) {
      int iters = 512, blocks = sms, threads = 1024;
      size_t sm = entries * sizeof(double2);
      float ms = timeit([&] { k_ldg_tab<<<blocks, threads, sm>>>(tab, out, entries, iters, mix); });
      double cnt = 3.0 * iters * (double)threads * blocks;
      printf("LDG.128 random L1-resident table%s: %.1f clk per warp LDG.128%s\n", mix ? " + equal LDS.128 stream" : "",
             1.0 / (cnt / 32 / (ms * 1e-3) / sms / 1.9e9), mix ? " (per LDG+LDS pair)" : "");
    }
  }
  for (int R : {1, 2, 4, 8}) {
    int entries = 1205, iters = 512, blocks = sms, threads = 1024;
    size_t sm = (size_t)entries * R * sizeof(double2);
    CK(cudaFuncSetAttribute(k_lds_rep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    float ms = timeit([&] { k_lds_rep<<<blocks, threads, sm>>>(out, entries, R, iters); });
    double cnt = 3.0 * iters * (double)threads * blocks;
    printf("LDS.128 random, %d-fold replicated table: %.1f clk per warp LDS.128\n", R, 1.0 / (cnt / 32 / (ms * 1e-3) / sms / 1.9e9));
  }
  return 0;
}
