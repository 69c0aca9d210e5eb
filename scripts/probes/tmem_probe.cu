// tmem_probe.cu -- can tensor memory (TMEM) serve as per-thread private scratch for a
// non-tensor-core kernel?  Each warp owns a 32-lane quarter x 64 columns; every thread
// writes 32 doubles (64 x 32-bit columns) with tcgen05.st, reads them back with tcgen05.ld,
// verifies, and the loop is timed with clock64.   nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void
tmem_st32(uint32_t taddr, const uint32_t (&v)[32])
{
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
               "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
               "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
               :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
               "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
               "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
               "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]),
               "r"(v[31]) : "memory");
}
__device__ __forceinline__ void
tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "
               "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
               "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]),
                 "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]),
                 "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]),
                 "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
               : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(512, 1)
tmem_probe(int iters, unsigned long long *cycles, int *errors)
{
  __shared__ uint32_t tmem_base;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (warp == 0)
    {
      const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(&tmem_base);
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(saddr), "r"(256));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = tmem_base;
  // lane quarter of this warp in bits 31:16, column offset in bits 15:0
  const uint32_t taddr = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
  uint32_t a[32], b[32];
  for (int i = 0; i < 32; ++i)
    {
      a[i] = 0x9e3779b9u * (uint32_t)(tid * 64 + i) + blockIdx.x;
      b[i] = 0x85ebca6bu * (uint32_t)(tid * 64 + 32 + i) + blockIdx.x;
    }
  tmem_st32(taddr, a);
  tmem_st32(taddr + 32, b);
  asm volatile("tcgen05.wait::st.sync.aligned;");
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
    {
      tmem_ld32(taddr, a);
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      for (int i = 0; i < 32; ++i)
        a[i] += 1u;
      tmem_st32(taddr, a);
      tmem_ld32(taddr + 32, b);
      asm volatile("tcgen05.wait::ld.sync.aligned;");
      for (int i = 0; i < 32; ++i)
        b[i] += 3u;
      tmem_st32(taddr + 32, b);
      asm volatile("tcgen05.wait::st.sync.aligned;");
    }
  const long long t1 = clock64();
  __syncthreads();
  tmem_ld32(taddr, a);
  tmem_ld32(taddr + 32, b);
  asm volatile("tcgen05.wait::ld.sync.aligned;");
  int bad = 0;
  for (int i = 0; i < 32; ++i)
    {
      bad += a[i] != 0x9e3779b9u * (uint32_t)(tid * 64 + i) + blockIdx.x + (uint32_t)iters;
      bad += b[i] != 0x85ebca6bu * (uint32_t)(tid * 64 + 32 + i) + blockIdx.x + 3u * (uint32_t)iters;
    }
  if (bad)
    atomicAdd(errors, bad);
  if (tid == 0 && blockIdx.x == 0)
    *cycles = (unsigned long long)(t1 - t0);
  __syncthreads();
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(256));
  (void)lane;
}

int
main()
{
  unsigned long long *d_cyc, h_cyc = 0;
  int                *d_err, h_err = 0;
  cudaMalloc(&d_cyc, 8);
  cudaMalloc(&d_err, 4);
  cudaMemset(d_err, 0, 4);
  const int iters = 1000;
  tmem_probe<<<148, 512>>>(iters, d_cyc, d_err);
  cudaError_t e = cudaDeviceSynchronize();
  printf("launch: %s\n", cudaGetErrorString(e));
  cudaMemcpy(&h_cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
  cudaMemcpy(&h_err, d_err, 4, cudaMemcpyDeviceToHost);
  const double bytes = 512.0 * 256.0 * 2.0; // per iteration per SM: ld + st of 256 B per thread
  printf("errors %d, cycles/iter %.1f, TMEM ld+st bytes/clk/SM %.1f\n", h_err, (double)h_cyc / iters,
         bytes * iters / (double)h_cyc);
  return h_err != 0 || e != cudaSuccess;
}
