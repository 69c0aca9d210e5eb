// msb_tmem.cuh -- tensor memory (TMEM) as per-thread private storage for the on-chip solve kernels.
// 32x32b shape: a warp owns a 32-lane quarter, a thread its lane's 32-bit columns; 8 doubles = 16
// consecutive columns.  No tensor-core instruction is issued anywhere in this library.
// Measured on B200 (scripts/probes/tmem_probe.cu): tcgen05.st + tcgen05.ld round trips are bit exact
// and sustain ~556 B/clk/SM.
// MSB_EMU (scripts/emu/*.cpp: the kernels compiled for the host) replaces every function by a
// thread-private array.
#pragma once

#include <stdint.h>

#ifndef MSB_EMU
namespace msb
{
  namespace tmem
  {
    __device__ __forceinline__ void
    ld8(uint32_t taddr, double (&d)[8])
    {
      uint32_t v[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                   "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                     "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
                     "=r"(v[14]), "=r"(v[15])
                   : "r"(taddr)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 8; ++i)
        d[i] = __hiloint2double((int)v[2 * i + 1], (int)v[2 * i]);
    }

    // two independent loads in flight, one wait
    __device__ __forceinline__ void
    ld8x2(uint32_t ta, uint32_t tb, double (&da)[8], double (&db)[8])
    {
      uint32_t v[16], w[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                   "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                     "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
                     "=r"(v[14]), "=r"(v[15])
                   : "r"(ta)
                   : "memory");
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                   "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]),
                     "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]),
                     "=r"(w[14]), "=r"(w[15])
                   : "r"(tb)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 8; ++i)
        {
          da[i] = __hiloint2double((int)v[2 * i + 1], (int)v[2 * i]);
          db[i] = __hiloint2double((int)w[2 * i + 1], (int)w[2 * i]);
        }
    }

    // three independent loads in flight, one wait
    __device__ __forceinline__ void
    ld8x3(uint32_t ta, uint32_t tb, uint32_t tc, double (&da)[8], double (&db)[8], double (&dc)[8])
    {
      uint32_t v[16], w[16], z[16];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                   "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
                     "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
                     "=r"(v[14]), "=r"(v[15])
                   : "r"(ta)
                   : "memory");
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                   "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]),
                     "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]),
                     "=r"(w[14]), "=r"(w[15])
                   : "r"(tb)
                   : "memory");
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                   "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                   : "=r"(z[0]), "=r"(z[1]), "=r"(z[2]), "=r"(z[3]), "=r"(z[4]), "=r"(z[5]), "=r"(z[6]),
                     "=r"(z[7]), "=r"(z[8]), "=r"(z[9]), "=r"(z[10]), "=r"(z[11]), "=r"(z[12]), "=r"(z[13]),
                     "=r"(z[14]), "=r"(z[15])
                   : "r"(tc)
                   : "memory");
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int i = 0; i < 8; ++i)
        {
          da[i] = __hiloint2double((int)v[2 * i + 1], (int)v[2 * i]);
          db[i] = __hiloint2double((int)w[2 * i + 1], (int)w[2 * i]);
          dc[i] = __hiloint2double((int)z[2 * i + 1], (int)z[2 * i]);
        }
    }

    __device__ __forceinline__ void
    st8(uint32_t taddr, const double (&d)[8])
    {
      uint32_t v[16];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        {
          v[2 * i]     = (uint32_t)__double2loint(d[i]);
          v[2 * i + 1] = (uint32_t)__double2hiint(d[i]);
        }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
                   "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
                   :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]),
                   "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]),
                   "r"(v[15])
                   : "memory");
    }

    __device__ __forceinline__ void
    wait_st()
    {
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }

    // Allocates `cols` columns for the CTA (warp 0 issues), publishes the base address through *slot
    // (shared memory) and returns it to every thread.  Contains a block barrier.
    __device__ __forceinline__ uint32_t
    alloc(uint32_t *slot, int warp, int cols)
    {
      if (warp == 0)
        {
          const uint32_t saddr = (uint32_t)__cvta_generic_to_shared(slot);
          asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(saddr), "r"(cols));
          asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncthreads();
      asm volatile("tcgen05.fence::after_thread_sync;");
      return *slot;
    }

    // Contains a block barrier.
    __device__ __forceinline__ void
    release(uint32_t base, int warp, int cols)
    {
      asm volatile("tcgen05.fence::before_thread_sync;");
      __syncthreads();
      if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(cols));
    }

    // this warp's 32-lane quarter (address bits 31:16)
    __device__ __forceinline__ uint32_t
    lane_quarter(int warp)
    {
      return (uint32_t)((warp & 3) * 32) << 16;
    }
  } // namespace tmem

  // bar.sync over the first COUNT threads of the CTA only (named barrier ID)
  template <int ID, int COUNT>
  __device__ __forceinline__ void
  named_barrier()
  {
    asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory");
  }
} // namespace msb
#else
// ---- host emulation: a thread-private array of 512 columns (256 doubles); the lane-quarter bits of an
// address are ignored because the storage is already private to the thread
namespace msb
{
  namespace tmem
  {
    inline double *
    priv()
    {
      thread_local double t[256];
      return t;
    }
    // tcgen05.ld / st / wait are .sync.aligned: every lane of the warp must execute them together.  The
    // emulation enforces it with a warp barrier, so a call from divergent code deadlocks HERE instead of
    // hanging the GPU (it did: a tcgen05.st inside `if (colok)` cost ten GPU-minutes in round 2).
    inline void
    ld8(uint32_t a, double (&d)[8])
    {
      __syncwarp();
      for (int i = 0; i < 8; ++i)
        d[i] = priv()[(a & 0xffffu) / 2 + i];
    }
    inline void
    ld8x2(uint32_t a, uint32_t b, double (&da)[8], double (&db)[8])
    {
      ld8(a, da);
      ld8(b, db);
    }
    inline void
    ld8x3(uint32_t a, uint32_t b, uint32_t c, double (&da)[8], double (&db)[8], double (&dc)[8])
    {
      ld8(a, da);
      ld8(b, db);
      ld8(c, dc);
    }
    inline void
    st8(uint32_t a, const double (&d)[8])
    {
      __syncwarp();
      for (int i = 0; i < 8; ++i)
        priv()[(a & 0xffffu) / 2 + i] = d[i];
    }
    inline void
    wait_st()
    {
      __syncwarp();
    }
    inline uint32_t
    alloc(uint32_t *, int, int)
    {
      for (int i = 0; i < 256; ++i)
        priv()[i] = NAN; // poison: the kernel must initialise what it reads
      __syncthreads();
      return 0;
    }
    inline void
    release(uint32_t, int, int)
    {
      __syncthreads();
    }
    inline uint32_t
    lane_quarter(int warp)
    {
      return (uint32_t)((warp & 3) * 32) << 16;
    }
  } // namespace tmem
  template <int ID, int COUNT>
  inline void
  named_barrier()
  {
    emu::named_barrier(ID, COUNT);
  }
} // namespace msb
#endif
