// ubench2.cu -- second set of B200 micro-benchmarks for the pair kernel (development tool, not product):
//   does a table read through the texture path (tex1Dfetch<int4>) use a data pipe separate from LDS?  random LDS.64 cost?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench2.bin scripts/ubench2.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ unsigned hash(unsigned x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

// MODE bit0: LDS.128 random stream (2 per iteration), bit1: tex int4 fetch stream (2 per iteration), bit2: LDS.64 x2 instead
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_mix(cudaTextureObject_t tex, const double2* __restrict__ tab, int n, int iters, double* out) {
  extern __shared__ __align__(16) double2 s[];
  for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = tab[i];
  __syncthreads();
  unsigned h = hash(blockIdx.x * blockDim.x + threadIdx.x + 99);
  double acc = 0.0;
  int u = h % n, v = (h >> 8) % n;
  for (int it = 0; it < iters; ++it) {
    if (MODE & 1) {
      double2 a = s[u], b = s[v];
      acc += a.x + b.y;
      u = (u + __double2loint(a.y) + 977) % n;
    }
    if (MODE & 4) {
      const double* sd = reinterpret_cast<const double*>(s);
      double a = sd[2 * u], b = sd[2 * v + 1];
      acc += a + b;
      u = (u + __double2loint(a) + 977) % n;
    }
    if (MODE & 2) {
      int4 a = tex1Dfetch<int4>(tex, v), b = tex1Dfetch<int4>(tex, (v + 613) % n);
      acc += __hiloint2double(a.y, a.x) + __hiloint2double(b.w, b.z);
      v = (v + a.z + 1201) % n;
    }
    if (!(MODE & 2)) v = (v + 1201) % n;
  }
  if (acc == 1.2345) out[0] = acc;
}

template <int MODE>
void run(const char* name, cudaTextureObject_t tex, const double2* tab, int n, int sms) {
  int iters = 4000;
  size_t smem = (size_t)n * 16;
  CK(cudaFuncSetAttribute(k_mix<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  double* out; CK(cudaMalloc(&out, 8));
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  k_mix<MODE><<<sms, 512, smem>>>(tex, tab, n, 100, out);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(a);
  k_mix<MODE><<<sms, 512, smem>>>(tex, tab, n, iters, out);
  cudaEventRecord(b);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, a, b);
  double clk = ms * 1e-3 * 1.965e9;                     // per SM
  double warp_iters = 16.0 * iters;                     // warps per SM x iterations
  printf("%-60s %8.3f ms  %6.1f clk per warp-iteration per SM\n", name, ms, clk / warp_iters);
  cudaFree(out);
}

int main() {
  int dev = 0; cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, dev));
  const int n = 4820;   // NaCl: 4 tables x 1205 entries of 16 B (g units)
  std::vector<double2> h(n);
  for (int i = 0; i < n; ++i) { h[i].x = 1e-3 * i; h[i].y = (double)((i * 7919) % 1000); }
  double2* tab; CK(cudaMalloc(&tab, n * 16)); CK(cudaMemcpy(tab, h.data(), n * 16, cudaMemcpyHostToDevice));
  cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = tab;
  rd.res.linear.desc = cudaCreateChannelDesc<int4>(); rd.res.linear.sizeInBytes = n * 16;
  cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
  cudaTextureObject_t tex; CK(cudaCreateTextureObject(&tex, &rd, &td, nullptr));
  run<1>("2 x LDS.128 random", tex, tab, n, p.multiProcessorCount);
  run<2>("2 x tex1Dfetch<int4> random (77 KB table)", tex, tab, n, p.multiProcessorCount);
  run<3>("2 x LDS.128 + 2 x tex int4", tex, tab, n, p.multiProcessorCount);
  run<4>("2 x LDS.64 random", tex, tab, n, p.multiProcessorCount);
  run<5>("2 x LDS.128 + 2 x LDS.64", tex, tab, n, p.multiProcessorCount);
  return 0;
}
