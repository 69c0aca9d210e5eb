// onchip_peaks.cu -- the two on-chip ceilings the SMEM-resident solve kernels sit under, MEASURED on the
// device the bench runs on ("FP64 vector peak: not measured -- builder must
// microbenchmark"):
//   (1) FP64 FMA throughput: every thread runs 8 independent DFMA chains in registers;
//   (2) shared-memory crossbar bandwidth: conflict-free LDS.128 / LDS.64 / STS.128 sweeps;
//   (3) HBM copy bandwidth (double2 grid-stride copy of 2 GiB) for reference beside MEASURED_PEAKS.json.
// Every figure is reported per SM per clock (clock64 inside the kernel) and as a device total
// (CUDA events), as one JSON object on stdout.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o onchip_peaks onchip_peaks.cu && ./onchip_peaks
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x)                                                                                 \
  do                                                                                          \
    {                                                                                         \
      cudaError_t e_ = (x);                                                                   \
      if (e_ != cudaSuccess)                                                                  \
        {                                                                                     \
          fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                            \
          return 1;                                                                           \
        }                                                                                     \
    }                                                                                         \
  while (0)

// ---------------------------------------------------------------------------------- FP64 FMA
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
dfma_kernel(int iters, double seed, double *out, unsigned long long *cycles)
{
  double a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
    a[i] = seed + threadIdx.x * 1e-9 + i;
  const double m = 1.0 - 1e-9, c = 1e-9;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          a[i] = fma(a[i], m, c);
    }
  __syncthreads();
  const long long t1 = clock64();
  double          s  = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    s += a[i];
  if (s == 12345.678)
    out[0] = s;
  if (threadIdx.x == 0)
    cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

// ---------------------------------------------------------------------------------- shared memory
// mode 0: LDS.128 (double2 per lane, consecutive), 1: LDS.64, 2: STS.128
template <int THREADS, int MODE>
__global__ void __launch_bounds__(THREADS)
smem_kernel(int iters, double *out, unsigned long long *cycles)
{
  extern __shared__ __align__(16) unsigned long long sm[];
  const int words = 16384; // 128 KB
  for (int i = threadIdx.x; i < words; i += THREADS)
    sm[i] = 0x9e3779b97f4a7c15ull * i;
  __syncthreads();
  // integer accumulation: the loads must not wait on an FP64 dependency chain
  unsigned long long acc0 = 0, acc1 = 0;
  const long long    t0   = clock64();
  for (int it = 0; it < iters; ++it)
    {
#pragma unroll
      for (int r = 0; r < 16; ++r)
        {
          if (MODE == 0)
            {
              // lane-consecutive 16-byte accesses, a different 8/16 KB block per repetition
              const int        idx = ((r * THREADS + threadIdx.x) * 2 + it * 2) & (words - 2);
              const ulonglong2 v   = *reinterpret_cast<const ulonglong2 *>(sm + idx);
              acc0 ^= v.x;
              acc1 ^= v.y;
            }
          else if (MODE == 1)
            {
              const int idx = (r * THREADS + threadIdx.x + it) & (words - 1);
              acc0 ^= sm[idx];
            }
          else
            {
              // (the address depends on the iteration so that no store is dead)
              const int idx = ((r * THREADS + threadIdx.x) * 2 + it * 2) & (words - 2);
              *reinterpret_cast<ulonglong2 *>(sm + idx) = make_ulonglong2(acc0 + it, acc1 + r);
            }
        }
    }
  __syncthreads();
  const long long t1 = clock64();
  if ((acc0 ^ acc1 ^ sm[threadIdx.x]) == 0x123456789abcdefull)
    out[0] = (double)acc0;
  if (threadIdx.x == 0)
    cycles[blockIdx.x] = (unsigned long long)(t1 - t0);
}

// ---------------------------------------------------------------------------------- HBM copy
__global__ void __launch_bounds__(256)
copy_kernel(const double2 *__restrict__ src, double2 *__restrict__ dst, size_t n)
{
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

static double
mean_cycles(const unsigned long long *c, int n)
{
  double s = 0;
  for (int i = 0; i < n; ++i)
    s += (double)c[i];
  return s / n;
}

int
main()
{
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int           sms = prop.multiProcessorCount;
  double             *d_out;
  unsigned long long *d_cyc, *h_cyc;
  CK(cudaMalloc(&d_out, 64));
  CK(cudaMalloc(&d_cyc, sizeof(unsigned long long) * sms * 4));
  h_cyc = (unsigned long long *)malloc(sizeof(unsigned long long) * sms * 4);
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float ms;

  printf("{\"device\": \"%s\", \"sms\": %d", prop.name, sms);

  // ---- FP64: 512 and 1024 threads per SM (1 CTA per SM), best of 5
  {
    const int iters = 4000;
    double    best_tf = 0, best_per_clk = 0;
    int       best_threads = 0;
    for (int cfg = 0; cfg < 2; ++cfg)
      for (int rep = 0; rep < 5; ++rep)
        {
          const int threads = cfg ? 1024 : 512;
          CK(cudaEventRecord(e0));
          if (cfg)
            dfma_kernel<1024><<<sms, 1024>>>(iters, 1.0, d_out, d_cyc);
          else
            dfma_kernel<512><<<sms, 512>>>(iters, 1.0, d_out, d_cyc);
          CK(cudaEventRecord(e1));
          CK(cudaEventSynchronize(e1));
          CK(cudaEventElapsedTime(&ms, e0, e1));
          CK(cudaMemcpy(h_cyc, d_cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost));
          const double fmas    = (double)iters * 64 * threads; // per SM
          const double per_clk = fmas / mean_cycles(h_cyc, sms);
          const double tf      = 2.0 * fmas * sms / (ms * 1e-3) / 1e12;
          if (tf > best_tf)
            best_tf = tf, best_per_clk = per_clk, best_threads = threads;
        }
    printf(", \"fp64\": {\"tflops\": %.3f, \"dfma_per_clk_per_sm\": %.2f, \"threads_per_sm\": %d}", best_tf,
           best_per_clk, best_threads);
  }

  // ---- shared memory: 1 CTA of 512 / 1024 threads per SM, 128 KB touched
  {
    const int   iters = 2000;
    const char *names[3] = {"lds128", "lds64", "sts128"};
    const int   bytes_per_access[3] = {16, 8, 16};
    CK(cudaFuncSetAttribute(smem_kernel<1024, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    CK(cudaFuncSetAttribute(smem_kernel<1024, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    CK(cudaFuncSetAttribute(smem_kernel<1024, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    CK(cudaFuncSetAttribute(smem_kernel<512, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    CK(cudaFuncSetAttribute(smem_kernel<512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    CK(cudaFuncSetAttribute(smem_kernel<512, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 131072));
    printf(", \"smem\": {");
    for (int mode = 0; mode < 3; ++mode)
      {
        double best_per_clk = 0, best_tbs = 0;
        int    best_threads = 0;
        for (int cfg = 0; cfg < 2; ++cfg)
          for (int rep = 0; rep < 5; ++rep)
            {
              const int threads = cfg ? 1024 : 512;
              CK(cudaEventRecord(e0));
              if (cfg)
                {
                  if (mode == 0)
                    smem_kernel<1024, 0><<<sms, 1024, 131072>>>(iters, d_out, d_cyc);
                  else if (mode == 1)
                    smem_kernel<1024, 1><<<sms, 1024, 131072>>>(iters, d_out, d_cyc);
                  else
                    smem_kernel<1024, 2><<<sms, 1024, 131072>>>(iters, d_out, d_cyc);
                }
              else
                {
                  if (mode == 0)
                    smem_kernel<512, 0><<<sms, 512, 131072>>>(iters, d_out, d_cyc);
                  else if (mode == 1)
                    smem_kernel<512, 1><<<sms, 512, 131072>>>(iters, d_out, d_cyc);
                  else
                    smem_kernel<512, 2><<<sms, 512, 131072>>>(iters, d_out, d_cyc);
                }
              CK(cudaEventRecord(e1));
              CK(cudaEventSynchronize(e1));
              CK(cudaEventElapsedTime(&ms, e0, e1));
              CK(cudaMemcpy(h_cyc, d_cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost));
              const double bytes   = (double)iters * 16 * threads * bytes_per_access[mode]; // per SM
              const double per_clk = bytes / mean_cycles(h_cyc, sms);
              const double tbs     = bytes * sms / (ms * 1e-3) / 1e12;
              if (per_clk > best_per_clk)
                best_per_clk = per_clk, best_tbs = tbs, best_threads = threads;
            }
        printf("%s\"%s\": {\"bytes_per_clk_per_sm\": %.1f, \"device_tb_s\": %.2f, \"threads_per_sm\": %d}",
               mode ? ", " : "", names[mode], best_per_clk, best_tbs, best_threads);
      }
    printf("}");
  }

  // ---- HBM copy, 2 GiB read + 2 GiB write, best of 5
  {
    const size_t n = (size_t)1 << 27; // double2 elements = 2 GiB
    double2     *a, *b;
    CK(cudaMalloc(&a, n * sizeof(double2)));
    CK(cudaMalloc(&b, n * sizeof(double2)));
    CK(cudaMemset(a, 1, n * sizeof(double2)));
    double best = 0;
    for (int rep = 0; rep < 6; ++rep)
      {
        CK(cudaEventRecord(e0));
        copy_kernel<<<sms * 16, 256>>>(a, b, n);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaEventElapsedTime(&ms, e0, e1));
        const double gbs = 2.0 * n * sizeof(double2) / (ms * 1e-3) / 1e9;
        if (rep > 0 && gbs > best)
          best = gbs;
      }
    printf(", \"hbm_copy_gbs\": %.1f", best);
    cudaFree(a);
    cudaFree(b);
  }
  int clk_khz = 0;
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  printf(", \"sm_clock_mhz_attr\": %.0f}\n", clk_khz / 1e3);
  return 0;
}
