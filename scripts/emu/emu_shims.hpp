// emu_shims.hpp -- host stand-ins for the CUDA execution model, shared by the emulation programs in this
// directory: every CUDA thread is an OS thread; __syncthreads / __syncwarp / warp shuffles are std::barriers,
// DSMEM is ordinary memory, st.async + mbarrier complete_tx and tensor memory are emulated.  Development and
// test tooling only: nothing in the library, bench.py or the GPU tests uses it.
#pragma once
#include <barrier>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

struct EmuDim
{
  unsigned x = 0, y = 0, z = 0;
};
thread_local EmuDim threadIdx, blockIdx, blockDim, gridDim;

namespace emu
{
  struct Warp
  {
    std::barrier<> bar{32};
    double         buf[32];
  };
  struct Cta
  {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<double>             smem;
    std::vector<std::unique_ptr<Warp>> warps;
  };
  struct Cluster
  {
    std::unique_ptr<std::barrier<>> bar;
    std::vector<Cta>                ctas;
  };
  thread_local Cluster *t_cluster = nullptr;
  thread_local int      t_rank    = 0;
  inline Cta &
  cta()
  {
    return t_cluster->ctas[t_rank];
  }
  inline double *
  smem()
  {
    return cta().smem.data();
  }
  // bar.sync ID, COUNT: a barrier over the first COUNT threads of the CTA (created on first use)
  inline std::mutex &
  named_mutex()
  {
    static std::mutex m;
    return m;
  }
  inline void
  named_barrier(int id, int count)
  {
    static std::map<std::pair<const void *, int>, std::unique_ptr<std::barrier<>>> bars;
    std::barrier<> *b;
    {
      std::lock_guard<std::mutex> lk(named_mutex());
      auto &slot = bars[{(const void *)&cta(), id}];
      if (!slot)
        slot = std::make_unique<std::barrier<>>(count);
      b = slot.get();
    }
    b->arrive_and_wait();
  }
} // namespace emu

inline void
__syncthreads()
{
  emu::cta().bar->arrive_and_wait();
}
inline void
__syncwarp()
{
  emu::cta().warps[threadIdx.x >> 5]->bar.arrive_and_wait();
}
inline double
__shfl_xor_sync(unsigned, double v, int off)
{
  emu::Warp &w   = *emu::cta().warps[threadIdx.x >> 5];
  const int  ln  = threadIdx.x & 31;
  w.buf[ln]      = v;
  w.bar.arrive_and_wait();
  const double o = w.buf[ln ^ off];
  w.bar.arrive_and_wait();
  return o;
}
inline double
__shfl_sync(unsigned, double v, int src)
{
  emu::Warp &w   = *emu::cta().warps[threadIdx.x >> 5];
  const int  ln  = threadIdx.x & 31;
  w.buf[ln]      = v;
  w.bar.arrive_and_wait();
  const double o = w.buf[src];
  w.bar.arrive_and_wait();
  return o;
}
inline long long
clock64()
{
  return 0;
}
inline double
rsqrt(double x)
{
  return 1.0 / std::sqrt(x);
}
inline int
min(int a, int b)
{
  return a < b ? a : b;
}
#include <atomic>
#include <mutex>
static std::mutex g_atomic_mutex;
inline int
atomicMin(int *p, int v)
{
  std::lock_guard<std::mutex> lk(g_atomic_mutex);
  const int                   o = *p;
  if (v < o)
    *p = v;
  return o;
}
template <class T>
inline T
atomicAdd(T *p, T v)
{
  std::lock_guard<std::mutex> lk(g_atomic_mutex);
  const T                     o = *p;
  *p += v;
  return o;
}

namespace cooperative_groups
{
  struct cluster_group
  {
    unsigned
    block_rank() const
    {
      return (unsigned)emu::t_rank;
    }
    void
    sync() const
    {
      emu::t_cluster->bar->arrive_and_wait();
    }
    template <class T>
    T *
    map_shared_rank(T *p, int rank) const
    {
      const size_t off = (const char *)p - (const char *)emu::smem();
      return (T *)((char *)emu::t_cluster->ctas[rank].smem.data() + off);
    }
  };
  inline cluster_group
  this_cluster()
  {
    return {};
  }
} // namespace cooperative_groups

// distributed-shared-memory primitives of msb_solve_cluster.cu: st.async + mbarrier complete_tx
#include <map>
namespace dsm
{
  struct State
  {
    long                  tx = 0;
    int                   pending = 1, count = 1;
    std::atomic<unsigned> phase{0};
  };
  static std::mutex                  mu;
  static std::map<const void *, State> tab; // keyed by the mbarrier's address in its CTA's memory
  template <class T>
  inline T *
  map_rank(T *p, int rank)
  {
    const size_t off = (const char *)p - (const char *)emu::smem();
    return (T *)((char *)emu::t_cluster->ctas[rank].smem.data() + off);
  }
  inline void
  try_complete(State &s)
  {
    if (s.pending == 0 && s.tx == 0)
      {
        s.pending = s.count;
        s.phase.fetch_add(1, std::memory_order_release);
        s.phase.notify_all();
      }
  }
  inline void
  mbar_init(uint64_t *mb, int count)
  {
    std::lock_guard<std::mutex> lk(mu);
    State                      &s = tab[mb];
    s.tx = 0, s.pending = s.count = count, s.phase.store(0);
  }
  inline void
  mbar_fence_init()
  {}
  inline void
  mbar_expect(uint64_t *mb, uint32_t bytes)
  {
    std::lock_guard<std::mutex> lk(mu);
    State                      &s = tab.at(mb);
    s.tx += bytes, s.pending -= 1;
    try_complete(s);
  }
  inline void
  push(double *dst, uint64_t *mb, int rank, double v)
  {
    *map_rank(dst, rank) = v;
    std::lock_guard<std::mutex> lk(mu);
    State                      &s = tab.at(map_rank(mb, rank));
    s.tx -= 8;
    try_complete(s);
  }
  inline void
  mbar_wait(uint64_t *mb, uint32_t parity)
  {
    State *s;
    {
      std::lock_guard<std::mutex> lk(mu);
      s = &tab.at(mb);
    }
    for (;;)
      {
        const unsigned p = s->phase.load(std::memory_order_acquire);
        if ((p & 1u) != parity)
          return;
        s->phase.wait(p, std::memory_order_acquire);
      }
  }
} // namespace dsm

// tensor memory: per-thread private columns (8 doubles = 16 columns)
namespace tmm
{
  thread_local double t_mem[64];
  inline void
  ld8(uint32_t a, double (&d)[8])
  {
    for (int i = 0; i < 8; ++i)
      d[i] = t_mem[a / 2 + i];
  }
  inline void
  st8(uint32_t a, const double (&d)[8])
  {
    for (int i = 0; i < 8; ++i)
      t_mem[a / 2 + i] = d[i];
  }
  inline void
  wait_st()
  {}
  inline uint32_t
  alloc(uint32_t *, int, int, uint32_t &base)
  {
    for (double &v : t_mem)
      v = NAN; // poison
    base = 0;
    emu::cta().bar->arrive_and_wait();
    return 0;
  }
  inline void
  release(uint32_t, int)
  {
    emu::cta().bar->arrive_and_wait();
  }
} // namespace tmm

#include <cuda_runtime.h>
#undef __launch_bounds__
#define __launch_bounds__(...)
