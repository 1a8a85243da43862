// pcie_test.cu -- how fast can the x,y,z / fx,fy,fz fields of DL_POLY's 64-byte corePart records cross PCIe? (development tool)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/pcie_test.bin scripts/pcie_test.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
struct Part { double x, y, z, fx, fy, fz, q; int p1, p2; };
__global__ void k_pull_xyz(int n, const Part* __restrict__ h, double4* __restrict__ d) {   // 32 B of every 64 B record
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double4 v = *reinterpret_cast<const double4*>(&h[i]);
  d[i] = v;
}
__global__ void k_pull_all(int n, const double4* __restrict__ h, double4* __restrict__ d) {  // whole records, coalesced
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 2 * n) d[i] = h[i];
}
__global__ void k_push_f(int n, const double* __restrict__ f, Part* __restrict__ h) {   // 24 B into every 64 B record
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  h[i].fx = f[i]; h[i].fy = f[n + i]; h[i].fz = f[2 * n + i];
}
__global__ void k_push_half(int n, const double4* __restrict__ d, Part* __restrict__ h) {   // upper 32 B (fx.. ) hmm: fx,fy,fz,q
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double4 v = d[i];
  double* p = &h[i].fx;   // offset 24: not 32-byte aligned -> three 8-byte stores
  p[0] = v.x; p[1] = v.y; p[2] = v.z;
}
template <class F> float timeit(F f, int reps = 5) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); CK(cudaDeviceSynchronize());
  float best = 1e9;
  for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaDeviceSynchronize()); float ms; cudaEventElapsedTime(&ms, a, b); best = ms < best ? ms : best; }
  return best;
}
int main() {
  const int n = 1260000;
  Part* h; CK(cudaHostAlloc((void**)&h, (size_t)n * 64, cudaHostAllocMapped));
  for (int i = 0; i < n; ++i) { h[i].x = i; h[i].y = 2 * i; h[i].z = 3 * i; h[i].q = 1; }
  char* d; CK(cudaMalloc((void**)&d, (size_t)n * 64));
  double* f; CK(cudaMalloc((void**)&f, (size_t)n * 24)); CK(cudaMemset(f, 0, (size_t)n * 24));
  Part* hd; CK(cudaHostGetDevicePointer((void**)&hd, h, 0));
  float t;
  t = timeit([&] { cudaMemcpyAsync(d, h, (size_t)n * 64, cudaMemcpyHostToDevice); });
  printf("H2D contiguous 64 B/atom          %7.3f ms  %6.1f GB/s\n", t, n * 64.0 / t / 1e6);
  t = timeit([&] { cudaMemcpy2DAsync(d, 24, h, 64, 24, n, cudaMemcpyHostToDevice); });
  printf("H2D 2D copy 24 of 64 B            %7.3f ms  %6.1f GB/s useful\n", t, n * 24.0 / t / 1e6);
  t = timeit([&] { cudaMemcpy2DAsync(d, 32, h, 64, 32, n, cudaMemcpyHostToDevice); });
  printf("H2D 2D copy 32 of 64 B            %7.3f ms  %6.1f GB/s useful\n", t, n * 32.0 / t / 1e6);
  t = timeit([&] { k_pull_xyz<<<(n + 255) / 256, 256>>>(n, hd, (double4*)d); });
  printf("zero-copy pull 32 of 64 B         %7.3f ms  %6.1f GB/s useful\n", t, n * 32.0 / t / 1e6);
  t = timeit([&] { k_pull_all<<<(2 * n + 255) / 256, 256>>>(n, (const double4*)hd, (double4*)d); });
  printf("zero-copy pull 64 of 64 B         %7.3f ms  %6.1f GB/s\n", t, n * 64.0 / t / 1e6);
  const int m = 1000000;
  t = timeit([&] { cudaMemcpyAsync(h, d, (size_t)m * 64, cudaMemcpyDeviceToHost); });
  printf("D2H contiguous 64 B/atom          %7.3f ms  %6.1f GB/s\n", t, m * 64.0 / t / 1e6);
  t = timeit([&] { cudaMemcpy2DAsync((char*)h + 24, 64, d, 24, 24, m, cudaMemcpyDeviceToHost); });
  printf("D2H 2D copy 24 into 64 B          %7.3f ms  %6.1f GB/s useful\n", t, m * 24.0 / t / 1e6);
  t = timeit([&] { k_push_f<<<(m + 255) / 256, 256>>>(m, f, hd); });
  printf("zero-copy push 24 into 64 B       %7.3f ms  %6.1f GB/s useful\n", t, m * 24.0 / t / 1e6);
  // both directions at once (two streams)
  cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
  t = timeit([&] { cudaMemcpyAsync(d, h, (size_t)n * 32, cudaMemcpyHostToDevice, s1); cudaMemcpyAsync((char*)h + (size_t)n * 32, d + (size_t)n * 32, (size_t)n * 32, cudaMemcpyDeviceToHost, s2); cudaStreamSynchronize(s1); cudaStreamSynchronize(s2); });
  printf("H2D + D2H concurrently, 40 MB each %7.3f ms  %6.1f GB/s per direction\n", t, n * 32.0 / t / 1e6);
  return 0;
}
