// msb_solve_cluster.cu -- thread-block-cluster / distributed-shared-memory tier: the multilevel
// preconditioned CG of the streamed tier (msb_solve_stream.cu, same mathematics, same stopping
// rule) for local meshes whose vectors do not fit ONE SM but do fit the shared memory of a
// cluster of SMs: n = 128, the reference's default run (main.cxx:23-25, n_refine_local = 7).
//
// One cluster of CS = n/16 CTAs per coarse cell (8 at n = 128, the portable maximum).  CTA c of
// the cluster owns the slab of 16 fine node rows [16c, 16c+16), a thread 4 consecutive rows of one
// column; all vectors of the solve live on chip for the whole iteration.  Two flavours:
//   <L, 4, true> (default)  all four bases of the cell in ONE pass
//     shared memory  the search direction p with one halo row either side; a staging copy of the
//                    residual that also parks q = K p; the slab's part of coarse levels 1 and 2
//                    (+ halo rows), a full copy of levels >= 3, reciprocal Galerkin diagonals
//     tensor memory  per-thread private columns (tcgen05.ld/st.32x32b, no MMA is issued): the
//                    thread's own 33 stencil coefficients, 1/diag and x
//     registers      r, z
//   <L, 2, false> (variant 4)  two passes of two bases; the five coefficient arrays of the slab
//                    and 1/diag in shared memory, x, r, z / q in registers, no tensor memory
// What crosses CTAs goes through DSMEM, always as a PUSH: st.async.shared::cluster with
// mbarrier complete_tx into the consumer's shared memory; the consumer arms its own mbarrier
// with the byte count it expects and waits on it -- point-to-point, no cluster-wide barrier and
// no fence in the iteration (cg::cluster.sync() compiles to MEMBAR.ALL.GPU + barrier + CCTL.IVALL
// and measured ~1900 cycles per use with 8 x 512 threads; six of them were 45 % of an iteration):
//     halo rows of z (-> p), r, r_1, r_2 to the neighbouring slab; the two level-3 rows a CTA owns to
//     every CTA of the cluster (levels >= 3 are then swept redundantly by all CTAs: no serial
//     coarse chain across the cluster); the per-CTA partial dot products to every CTA, which
//     adds them with the same butterfly so that all CTAs take bitwise identical decisions.
// The three all-reduces of an iteration are the only cluster-wide synchronisation points; every
// buffer a CTA pushes into was last read by its owner before an all-reduce the pusher has
// already completed (see the hazard table in DESIGN.md 3.4).
// HBM traffic = read the stencil once, write Phi.
//
// Replaces, for these local meshes, the per-basis sequence of the reference:
// diffusion_problem_basis.tpp:450-465 (condense, solve_iterative :293-317, distribute :308).
#ifndef MSB_EMU // scripts/emu/cluster_emu.cpp compiles this file for the host with its own shims
#  include <cooperative_groups.h>
#endif
#include <limits.h>
#include <math.h>

#define MSB_STAGE_ARRAY g_msb_stage_cycles_cl
#include "msb_bpx_common.cuh"

#ifdef MSB_STAGE_TIMERS
__device__ unsigned long long g_msb_stage_cycles_cl[16];
extern "C" int
msb_debug_stage_cycles_cl(unsigned long long *out, int reset)
{
  cudaError_t e = cudaMemcpyFromSymbol(out, g_msb_stage_cycles_cl, sizeof(unsigned long long) * 16);
  if (e == cudaSuccess && reset)
    {
      unsigned long long z[16] = {0};
      e = cudaMemcpyToSymbol(g_msb_stage_cycles_cl, z, sizeof z);
    }
  return (int)e;
}
#endif

namespace cg = cooperative_groups;

// ---- distributed-shared-memory primitives (device: PTX; MSB_EMU: scripts/emu/cluster_emu.cpp)
#ifndef MSB_EMU
namespace dsm
{
  __device__ __forceinline__ uint32_t
  saddr(const void *p)
  {
    return (uint32_t)__cvta_generic_to_shared(p);
  }
  // shared::cluster address of CTA `rank`'s copy of a shared::cta address
  __device__ __forceinline__ uint32_t
  peer(uint32_t a, int rank)
  {
    uint32_t r;
    asm("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
    return r;
  }
  // store one double into CTA `rank`'s copy of *dst; its mbarrier *mb counts the 8 bytes
  __device__ __forceinline__ void
  push(double *dst, uint64_t *mb, int rank, double v)
  {
    asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                 :
                 : "r"(peer(saddr(dst), rank)), "l"(__double_as_longlong(v)), "r"(peer(saddr(mb), rank))
                 : "memory");
  }
  __device__ __forceinline__ void
  mbar_init(uint64_t *mb, int count)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" : : "r"(saddr(mb)), "r"(count) : "memory");
  }
  __device__ __forceinline__ void
  mbar_fence_init()
  {
    asm volatile("fence.mbarrier_init.release.cluster;" : : : "memory");
  }
  // the consumer's single arrival of a phase + the bytes it expects from its producers
  __device__ __forceinline__ void
  mbar_expect(uint64_t *mb, uint32_t bytes)
  {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" : : "r"(saddr(mb)), "r"(bytes) : "memory");
  }
  __device__ __forceinline__ void
  mbar_wait(uint64_t *mb, uint32_t parity)
  {
    asm volatile("{\n\t"
                 ".reg .pred P1;\n\t"
                 "DSM_WAIT:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, 0x989680;\n\t"
                 "@P1 bra DSM_DONE;\n\t"
                 "bra DSM_WAIT;\n\t"
                 "DSM_DONE:\n\t"
                 "}"
                 :
                 : "r"(saddr(mb)), "r"(parity)
                 : "memory");
  }
} // namespace dsm

// ---- tensor memory as per-thread private storage (32x32b shape: a warp owns a 32-lane quarter, a
// thread its lane's columns; no tensor-core instruction is issued).  8 doubles = 16 columns.
namespace tmm
{
  constexpr int COLS = 512; // the whole tensor memory of the SM (one CTA per SM)
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
  // all threads of the CTA; returns this thread's base address (lane quarter of its warp in bits
  // 31:16, the column block of its warp in bits 15:0: warps w, w+4, w+8, .. share a lane quarter)
  __device__ __forceinline__ uint32_t
  alloc(uint32_t *s_base, int warp, int cols_per_warp, uint32_t &base_out)
  {
    if (warp == 0)
      {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                       (uint32_t)__cvta_generic_to_shared(s_base)),
                     "r"(COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
      }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    base_out = *s_base;
    return base_out + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * cols_per_warp);
  }
  __device__ __forceinline__ void
  release(uint32_t base, int warp)
  {
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(COLS));
  }
} // namespace tmm
#endif

namespace msb
{
  namespace clus
  {
    constexpr int ROWS = 16; // fine node rows per CTA

    struct Params
    {
      const double *corners; // [C][8]
      const double *q1coef;  // [C][16]
      const double *sten;    // [C][6][N]
      const double *dinv;    // [C][cn] reciprocal Galerkin diagonals, levels 1.. (stream_galerkin_kernel)
      double       *phi;     // [C][4][N]
      int32_t      *iters;   // [C][4]
      double       *res;     // [C][4]
      int32_t      *fail;
      double        tol2;
      int           max_iter;
      int           cn;
      int           cell0; // first cell of this launch
      int           split; // 1: one cluster per (cell, pass) instead of one per cell (the tail of a small shard)
    };

    // NBP = bases per pass.  TM = false: NBP = 2, coefficients in shared memory, x and q in registers.
    // TM = true: NBP = 4 (one pass), the thread's own stencil coefficients, 1/diag and x in tensor
    // memory, q parked in the residual staging buffer: frees 100 KB of shared memory for the vectors
    // of two more bases and takes the coefficient loads off the shared-memory crossbar.
    template <int L, int NBP, bool TM>
    struct Lay
    {
      static constexpr int n = 1 << L, np = n + 1, N = np * np, CS = n / ROWS, T = 4 * n, NW = T / 32;
      static constexpr int W = n, W1 = n / 2, W2 = n / 4, W3 = n / 8, NP3 = W3 + 1;
      static constexpr int LV = L - 1; // coarse levels 1..LV
      __host__ __device__ static constexpr int
      npl(int l)
      {
        return (n >> l) + 1;
      }
      // node offset of level l inside the packed per-cell coarse arrays (make_levels, msb_solve_stream.cu)
      __host__ __device__ static constexpr int
      goff(int l)
      {
        int o = 0;
        for (int k = 1; k < l; ++k)
          o += npl(k) * npl(k);
        return o;
      }
      // threads the direct restriction to the levels > 3 occupies (whole warps per level)
      __host__ __device__ static constexpr int
      deep_threads()
      {
        int b = 0;
        for (int m = 1; m <= LV - 3; ++m)
          {
            const int nin = npl(3 + m) - 2, items = NBP * nin * nin, G = m == 1 ? 1 : m == 2 ? 4 : 16;
            b += ((items * G + 31) / 32) * 32;
          }
        return b;
      }
      static constexpr int cn  = goff(LV + 1);
      static constexpr int cn3 = goff(LV + 1) - goff(3); // nodes of levels 3..LV (full copies)
      // shared-memory map, in doubles
      static constexpr int o_cf  = 0;                              // [5][17][W]   rows y0-1 .. y0+15
      static constexpr int o_d0  = o_cf + (TM ? 0 : 5 * 17 * W);   // [16][W]      1/KC, own rows
      static constexpr int o_p   = o_d0 + (TM ? 0 : 16 * W);       // [NBP][18][W] rows y0-1 .. y0+16 (+ pad)
      static constexpr int o_rs  = o_p + NBP * 18 * W + W;         // [NBP][17][W] rows y0-1 .. y0+15 (+ pad)
      static constexpr int o_v1  = o_rs + NBP * 17 * W + W;        // [NBP][10][W1] level-1 rows 8c-1 .. 8c+8 (+ pad)
      static constexpr int o_d1  = o_v1 + NBP * 10 * W1 + W1;      // [9][W1]      level-1 rows 8c .. 8c+8
      static constexpr int o_v2  = o_d1 + 9 * W1;                  // [NBP][6][W2] level-2 rows 4c-1 .. 4c+4 (+ pad)
      static constexpr int o_d2  = o_v2 + NBP * 6 * W2 + W2;       // [5][W2]
      static constexpr int o_v3  = o_d2 + 5 * W2;                  // [NBP][cn3]   levels >= 3, full, stride npl
      static constexpr int o_d3  = o_v3 + NBP * cn3;               // [cn3]
      static constexpr int o_red = o_d3 + cn3;                     // [3][CS][NBP] partial dot products
      static constexpr int o_buf = o_red + 3 * CS * NBP;           // [3][NBP][NW] block reduction scratch
      static constexpr int BUFS  = NBP * NW > 32 ? NBP * NW : 32;  // one block-reduction scratch (bpx::block_sum4 needs >= 32)
      static constexpr int o_zh  = o_buf + 3 * BUFS;               // [NBP][2][W]  z of the rows y0-1 and y0+16
      static constexpr int o_mb  = o_zh + NBP * 2 * W;             // [6] mbarriers (64 bit each) + TMEM base
      static constexpr int total = o_mb + 8;
      // tensor-memory map of a thread (32-bit columns): x [NBP][4] | 8 coefficients per own row | kS of
      // row 0 and 1/diag of the 4 rows
      static constexpr int XOFF = 0, COFF = 4 * NBP * 2, EOFF = COFF + 4 * 16, TCOLS = 128;
      static_assert(!TM || (NBP == 4 && EOFF + 16 <= TCOLS && T / 32 <= 16), "four warps share a lane quarter");
      static_assert(NBP % 2 == 0 && 4 % NBP == 0, "bases are handled in pairs");
      static constexpr size_t smem_bytes = sizeof(double) * (size_t)total;
      static_assert(L >= 5 && L <= 7, "cluster tier: 32 <= n <= 128 (cluster of 2..8 CTAs)");
      static_assert(NW <= 32, "block_sum: one lane per warp partial");
      static_assert(deep_threads() <= T, "direct restriction: one segment of whole warps per level");
    };

    // full weighting of one coarse node from a buffer with row stride `stride`, centred at s
    __device__ __forceinline__ double
    restrict_node(const double *s, int stride)
    {
      const double a = fma(0.5, s[-stride - 1] + s[-stride + 1], s[-stride]);
      const double b = fma(0.5, s[-1] + s[1], s[0]);
      const double c = fma(0.5, s[stride - 1] + s[stride + 1], s[stride]);
      return fma(0.5, a + c, b);
    }

    template <int L, int NBP, bool TM>
    __global__ void __launch_bounds__((Lay<L, NBP, TM>::T), 1)
    solve_cluster_kernel(Params P)
    {
      using Y = Lay<L, NBP, TM>;
      constexpr int n = Y::n, np = Y::np, N = Y::N, CS = Y::CS, T = Y::T, NW = Y::NW;
      constexpr int W = Y::W, W1 = Y::W1, W2 = Y::W2, W3 = Y::W3, NP3 = Y::NP3, LV = Y::LV, cn3 = Y::cn3;
#ifndef MSB_EMU
      extern __shared__ __align__(16) double sm[];
#else
      double *sm = emu::smem();
#endif
      cg::cluster_group cluster = cg::this_cluster();
      const int rank = (int)cluster.block_rank();
      // split launches (the tail of a shard whose last wave of clusters would leave most SMs idle): cluster -> (cell, pass)
      const int cid     = blockIdx.x / CS;
      const int cell    = P.cell0 + (P.split ? cid / (4 / NBP) : cid);
      const int pass_lo = P.split ? cid % (4 / NBP) : 0, pass_hi = P.split ? pass_lo + 1 : 4 / NBP;
      const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
      const int jx = tid & (n - 1), g = tid >> L; // column, group of 4 rows
      const int y0 = rank * ROWS;
      const bool colact = jx >= 1;

      double *cf = sm + Y::o_cf, *d0 = sm + Y::o_d0, *pS = sm + Y::o_p, *rS = sm + Y::o_rs;
      double *v1 = sm + Y::o_v1, *d1 = sm + Y::o_d1, *v2 = sm + Y::o_v2, *d2 = sm + Y::o_d2;
      double *v3 = sm + Y::o_v3, *d3 = sm + Y::o_d3, *red = sm + Y::o_red, *buf = sm + Y::o_buf;
      double *zh = sm + Y::o_zh;
      uint64_t *mb = reinterpret_cast<uint64_t *>(sm + Y::o_mb);
      // one mbarrier per kind of exchange, one phase per use; bit i of ph = parity to wait for next
      enum { MB_V1 = 0, MB_V2, MB_V3, MB_RZ, MB_PQ, MB_RR };
      unsigned  ph = 0;
      const int nn = (rank > 0) + (rank + 1 < CS); // neighbouring slabs
      // wait until every byte the producers of this exchange push into this CTA has landed
      auto await = [&](int which, int bytes) {
        if (tid == 0)
          dsm::mbar_expect(mb + which, (uint32_t)bytes);
        dsm::mbar_wait(mb + which, (ph >> which) & 1u);
        ph ^= 1u << which;
      };

      const double *S  = P.sten + (size_t)cell * ST_NARR * N;
      const double *c  = P.corners + 8 * (size_t)cell;
      const double *q1 = P.q1coef + 16 * (size_t)cell;

      ST_DECL
      // ------------------------------------------------------------------ prologue (once per cell)
      for (int i = tid; i < Y::total; i += T)
        sm[i] = 0.0;
      __syncthreads();
      if (tid == 0)
        {
          for (int i = 0; i < 6; ++i)
            dsm::mbar_init(mb + i, 1);
          dsm::mbar_fence_init();
        }
      uint32_t tm = 0, tm_base = 0; // this thread's tensor-memory base address
      if constexpr (TM)
        {
          tm = tmm::alloc(reinterpret_cast<uint32_t *>(mb + 6), warp, Y::TCOLS, tm_base);
          // the thread's own coefficients straight from HBM (lanes = consecutive columns: coalesced)
          double e8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
          for (int i = 0; i < 4; ++i)
            {
              const int y = y0 + 4 * g + i, t = y * np + jx;
              double    c8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
              if (colact && y >= 1)
                {
                  c8[0] = S[(size_t)ST_KC * N + t];           // kc
                  c8[1] = S[(size_t)ST_KE * N + t];           // kE
                  c8[2] = S[(size_t)ST_KE * N + t - 1];       // kW
                  c8[3] = S[(size_t)ST_KN * N + t];           // kN
                  c8[4] = S[(size_t)ST_KD1 * N + t];          // kNE
                  c8[5] = S[(size_t)ST_KD1 * N + t - np - 1]; // kSW
                  c8[6] = S[(size_t)ST_KD2 * N + t - 1];      // kNW
                  c8[7] = S[(size_t)ST_KD2 * N + t - np];     // kSE
                  e8[1 + i] = 1.0 / c8[0];
                  if (i == 0)
                    e8[0] = S[(size_t)ST_KN * N + t - np];    // kS of the first row (then: kN of the row below)
                }
              tmm::st8(tm + Y::COFF + 16 * i, c8);
            }
          tmm::st8(tm + Y::EOFF, e8);
          tmm::wait_st();
        }
      else
        {
          for (int t = tid; t < 17 * W; t += T)
            {
              const int r = t >> L, x = t & (W - 1), y = y0 - 1 + r;
              if (y >= 0)
                {
                  const int gi = y * np + x;
#pragma unroll
                  for (int a = 0; a < 5; ++a)
                    cf[(a * 17 + r) * W + x] = S[(size_t)a * N + gi];
                  if (r >= 1 && y >= 1 && x >= 1)
                    d0[(r - 1) * W + x] = 1.0 / S[(size_t)ST_KC * N + gi];
                }
            }
        }
      {
        const double *dg = P.dinv + (size_t)cell * P.cn;
        for (int t = tid; t < 9 * W1; t += T)
          {
            const int r = t / W1, X = t % W1;
            d1[t] = dg[Y::goff(1) + (8 * rank + r) * Y::npl(1) + X];
          }
        for (int t = tid; t < 5 * W2; t += T)
          {
            const int r = t / W2, X = t % W2;
            d2[t] = dg[Y::goff(2) + (4 * rank + r) * Y::npl(2) + X];
          }
        for (int t = tid; t < cn3; t += T)
          d3[t] = dg[Y::goff(3) + t];
      }
      // nobody may push into a peer's shared memory before that peer has zeroed it and set up its mbarriers
      cluster.sync();
      ST_MARK(0)

      // sum of one value per basis over the whole cluster, the same bits in every thread of every CTA
      // (slot 0: r.z, 1: p.q, 2: r.r; `extra` = bytes of halo rows that travel with the partials)
      auto allreduce = [&](double(&v)[NBP], int slot, int extra) {
        // transposing exchanges: 2 (4) values for the shuffle count of one butterfly
        if constexpr (NBP == 2 && NW <= 16)
          bpx::block_sum2<NW>(v[0], v[1], buf + slot * Y::BUFS, warp, lane);
        else if constexpr (NBP == 4 && NW <= 16)
          bpx::block_sum4<NW>(v, buf + slot * Y::BUFS, warp, lane);
        else
          bpx::block_sum<NBP, NW>(v, buf + slot * Y::BUFS, warp, lane);
        if (tid < CS)
          {
#pragma unroll
            for (int k = 0; k < NBP; ++k)
              dsm::push(red + (slot * CS + rank) * NBP + k, mb + MB_RZ + slot, tid, v[k]);
          }
        await(MB_RZ + slot, 8 * (CS * NBP + extra));
        // lane -> (CTA j = lane % CS, basis k = lane / CS): one load per lane, a butterfly over the CS
        // partials (the same tree in every warp of every CTA: identical bits), one broadcast per basis
        static_assert(CS * NBP <= 32, "one lane per partial");
        double s = lane < CS * NBP ? red[(slot * CS + (lane % CS)) * NBP + lane / CS] : 0.0;
#pragma unroll
        for (int off = CS / 2; off > 0; off >>= 1)
          s += __shfl_xor_sync(0xffffffffu, s, off);
#pragma unroll
        for (int k = 0; k < NBP; ++k)
          v[k] = __shfl_sync(0xffffffffu, s, k * CS);
      };

      for (int pass = pass_lo; pass < pass_hi; ++pass)
        {
          double x[TM ? 1 : NBP][4], r[NBP][4], z[NBP][4]; // TM: x lives in tensor memory
          double rr[NBP], rz[NBP], beta[NBP];
          int    itc[NBP];
          bool   done[NBP];

          // ---- x = g on the boundary (distribute, basis.tpp:308), written straight to Phi
          for (int i = rank * T + tid; i < 4 * n; i += CS * T)
            {
              const int side = i / n, o = i % n;
              const int bx = side == 0 ? o : side == 1 ? n : side == 2 ? n - o : 0;
              const int by = side == 0 ? 0 : side == 1 ? o : side == 2 ? n : n - o;
              double    px, py;
              fine_vertex(c, n, bx, by, px, py);
#pragma unroll
              for (int k = 0; k < NBP; ++k)
                P.phi[((size_t)cell * 4 + NBP * pass + k) * N + by * np + bx] =
                  basis_q1_value(q1, NBP * pass + k, px, py);
            }
          // ---- initial guess x_0 = g, the coarse Q1 shape function, on the interior nodes too (the exact
          //      solution for a constant coefficient: 38.6 -> 35 iterations on the reference's default run);
          //      r_0 = b - K_II g_I = -(K g) on interior rows (b = -K_IB g_B: condense)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            {
              const int y = y0 + 4 * g + i;
#pragma unroll
              for (int k = 0; k < NBP; ++k)
                {
                  r[k][i] = 0.0, z[k][i] = 0.0;
                  if constexpr (!TM)
                    x[k][i] = 0.0;
                }
              if (colact && y >= 1)
                {
                  double rv[NBP];
#pragma unroll
                  for (int k = 0; k < NBP; ++k)
                    rv[k] = 0.0;
#pragma unroll
                  for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                    for (int dx = -1; dx <= 1; ++dx)
                      {
                        const double kij = bpx::sten_get(S, np, N, jx, y, dx, dy);
                        double       px, py;
                        fine_vertex(c, n, jx + dx, y + dy, px, py);
#pragma unroll
                        for (int k = 0; k < NBP; ++k)
                          {
                            const double gv = basis_q1_value(q1, NBP * pass + k, px, py);
                            rv[k] -= kij * gv;
                            if constexpr (!TM)
                              if (dx == 0 && dy == 0)
                                x[k][i] = gv;
                          }
                      }
#pragma unroll
                  for (int k = 0; k < NBP; ++k)
                    r[k][i] = rv[k];
                }
            }

          if constexpr (TM)
            {
#pragma unroll
              for (int kp = 0; kp < NBP / 2; ++kp)
                {
                  double x8[8]; // x of bases 2kp, 2kp+1, rows 0..3
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    {
                      const int y = y0 + 4 * g + i;
                      double    px, py;
                      fine_vertex(c, n, colact ? jx : 1, y >= 1 ? y : 1, px, py);
#pragma unroll
                      for (int kk = 0; kk < 2; ++kk)
                        x8[4 * kk + i] = (colact && y >= 1) ? basis_q1_value(q1, NBP * pass + 2 * kp + kk, px, py) : 0.0;
                    }
                  tmm::st8(tm + Y::XOFF + 16 * kp, x8);
                }
              tmm::wait_st();
            }
          // residual to the staging buffer (+ the slab's last row into the upper neighbour's halo)
          auto stage_r = [&]() {
            if (colact)
              {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                  for (int k = 0; k < NBP; ++k)
                    rS[(k * 17 + 4 * g + i + 1) * W + jx] = r[k][i];
                if (g == 3 && rank + 1 < CS)
                  {
#pragma unroll
                    for (int k = 0; k < NBP; ++k)
                      dsm::push(rS + (k * 17 + 0) * W + jx, mb + MB_RR, rank + 1, r[k][3]);
                  }
              }
          };

          {
            double acc[NBP];
#pragma unroll
            for (int k = 0; k < NBP; ++k)
              {
                acc[k] = 0.0;
#pragma unroll
                for (int i = 0; i < 4; ++i)
                  acc[k] = fma(r[k][i], r[k][i], acc[k]);
              }
            stage_r();
            allreduce(acc, 2, rank > 0 ? NBP * (W - 1) : 0);
#pragma unroll
            for (int k = 0; k < NBP; ++k)
              rr[k] = acc[k], rz[k] = 1.0, beta[k] = 0.0, itc[k] = -1, done[k] = false;
          }
          ST_MARK(1)

          int it = 0;
          while (true)
            {
              // ---- stopping rule of SolverControl (basis.tpp:297): ||r||_2 <= tol, every iteration
              bool all_done = true;
#pragma unroll
              for (int k = 0; k < NBP; ++k)
                {
                  if (!done[k] && rr[k] <= P.tol2)
                    done[k] = true, itc[k] = it;
                  all_done = all_done && done[k];
                }
              if (all_done || it >= P.max_iter)
                break;

              // ---- z = M^-1 r: additive multilevel preconditioner (exact Galerkin diagonals)
              // level 1 rows 8c .. 8c+7 from the staged residual (own rows + lower halo)
              for (int t = tid; t < NBP * 8 * W1; t += T)
                {
                  const int X = t & (W1 - 1), qr = (t / W1) & 7, k = t / (8 * W1);
                  if (X >= 1 && 8 * rank + qr >= 1)
                    {
                      const double v = restrict_node(rS + (k * 17 + 2 * qr + 1) * W + 2 * X, W);
                      v1[(k * 10 + qr + 1) * W1 + X] = v;
                      if (qr == 0 && rank > 0)
                        dsm::push(v1 + (k * 10 + 9) * W1 + X, mb + MB_V1, rank - 1, v);
                      if (qr == 7 && rank + 1 < CS)
                        dsm::push(v1 + (k * 10 + 0) * W1 + X, mb + MB_V1, rank + 1, v);
                    }
                }
              __syncthreads();
              await(MB_V1, 8 * nn * NBP * (W1 - 1));
              ST_MARK(2)
              // level 2 rows 4c .. 4c+3
              for (int t = tid; t < NBP * 4 * W2; t += T)
                {
                  const int X = t & (W2 - 1), qr = (t / W2) & 3, k = t / (4 * W2);
                  if (X >= 1 && 4 * rank + qr >= 1)
                    {
                      const double v = restrict_node(v1 + (k * 10 + 2 * qr + 1) * W1 + 2 * X, W1);
                      v2[(k * 6 + qr + 1) * W2 + X] = v;
                      if (qr == 0 && rank > 0)
                        dsm::push(v2 + (k * 6 + 5) * W2 + X, mb + MB_V2, rank - 1, v);
                      if (qr == 3 && rank + 1 < CS)
                        dsm::push(v2 + (k * 6 + 0) * W2 + X, mb + MB_V2, rank + 1, v);
                    }
                }
              __syncthreads();
              await(MB_V2, 8 * nn * NBP * (W2 - 1));
              ST_MARK(3)
              // level 3 rows 2c, 2c+1 -> every CTA of the cluster
              for (int t = tid; t < NBP * 2 * W3; t += T)
                {
                  const int X = t & (W3 - 1), qr = (t / W3) & 1, k = t / (2 * W3);
                  const int Y3 = 2 * rank + qr;
                  if (X >= 1 && Y3 >= 1)
                    {
                      const double v = restrict_node(v2 + (k * 6 + 2 * qr + 1) * W2 + 2 * X, W2);
#pragma unroll
                      for (int j = 0; j < CS; ++j)
                        dsm::push(v3 + k * cn3 + Y3 * NP3 + X, mb + MB_V3, j, v);
                    }
                }
              await(MB_V3, 8 * NBP * (W3 - 1) * (W3 - 1));
              ST_MARK(4)
              // levels > 3, redundantly in every CTA.  The preconditioner is additive, so the levels are
              // independent: r_l = (P_3..l)^T r_3 directly (nested hat functions of half width 2^(l-3)),
              // scaled by 1/D_l, by different warps at the same time; then ONE sweep of level 3 adds the
              // direct bilinear interpolation of every deeper level.  Two barriers instead of 2 (LV-3) + 1.
              {
                int base = 0; // first thread of the level's segment (whole warps)
#pragma unroll
                for (int m = 1; m <= LV - 3; ++m)
                  {
                    const int l = 3 + m, npl = Y::npl(l), nin = npl - 2, items = NBP * nin * nin;
                    const int lo = Y::goff(l) - Y::goff(3), h = 1 << m;
                    const int G   = m == 1 ? 1 : m == 2 ? 4 : 16; // lanes per coarse node, each sums whole window rows
                    const int seg = ((items * G + 31) / 32) * 32;
                    if (tid >= base && tid < base + seg)
                      {
                        const int u = tid - base, item = u / G, sub = u % G;
                        const int k = item / (nin * nin), q2 = item % (nin * nin), cx = 1 + q2 % nin, cy = 1 + q2 / nin;
                        double    acc = 0.0;
                        if (item < items)
                          {
                            const double *src = v3 + k * cn3 + (cy * h) * NP3 + cx * h;
                            if (m == 1)
                              acc = restrict_node(src, NP3);
                            else
                              {
                                const double rh = 1.0 / h;
                                for (int wr = sub; wr < 2 * h - 1; wr += G)
                                  {
                                    const int dy  = wr - (h - 1);
                                    double    row = 0.0;
                                    for (int dx = -(h - 1); dx <= h - 1; ++dx)
                                      row = fma(1.0 - abs(dx) * rh, src[dy * NP3 + dx], row);
                                    acc = fma(1.0 - abs(dy) * rh, row, acc);
                                  }
                              }
                          }
#pragma unroll
                        for (int off = G / 2; off > 0; off >>= 1)
                          acc += __shfl_xor_sync(0xffffffffu, acc, off);
                        if (item < items && sub == 0)
                          v3[k * cn3 + lo + cy * npl + cx] = acc * d3[lo + cy * npl + cx];
                      }
                    base += seg;
                  }
                __syncthreads();
                const int nin = NP3 - 2;
                for (int t = tid; t < NBP * nin * nin; t += T)
                  {
                    const int k = t / (nin * nin), u = t % (nin * nin), fx = 1 + u % nin, fy = 1 + u / nin;
                    const int i = fy * NP3 + fx;
                    double    v = v3[k * cn3 + i] * d3[i];
#pragma unroll
                    for (int m = 1; m <= LV - 3; ++m)
                      {
                        const int     l = 3 + m, npc = Y::npl(l), h = 1 << m;
                        const double *vc = v3 + k * cn3 + Y::goff(l) - Y::goff(3);
                        const int     cx = fx >> m, cy = fy >> m;
                        const double  tx = (fx & (h - 1)) * (1.0 / h), ty = (fy & (h - 1)) * (1.0 / h);
                        const double  a = fma(tx, vc[cy * npc + cx + 1] - vc[cy * npc + cx], vc[cy * npc + cx]);
                        const double  b =
                          fma(tx, vc[(cy + 1) * npc + cx + 1] - vc[(cy + 1) * npc + cx], vc[(cy + 1) * npc + cx]);
                        v += fma(ty, b - a, a);
                      }
                    v3[k * cn3 + i] = v;
                  }
                __syncthreads();
              }
              ST_MARK(5)
              // level 2: own rows + the upper halo row (needed by level-1 rows of this slab)
              for (int t = tid; t < NBP * 5 * W2; t += T)
                {
                  const int X = t & (W2 - 1), qr = (t / W2) % 5, k = t / (5 * W2);
                  const int Y2 = 4 * rank + qr;
                  if (X >= 1 && Y2 >= 1 && Y2 <= W2 - 1)
                    {
                      const int     i  = (k * 6 + qr + 1) * W2 + X;
                      const double *vc = v3 + k * cn3;
                      const int     xl = X >> 1, xh = (X + 1) >> 1, yl = Y2 >> 1, yh = (Y2 + 1) >> 1;
                      v2[i] = v2[i] * d2[qr * W2 + X] +
                              0.25 * ((vc[yl * NP3 + xl] + vc[yl * NP3 + xh]) + (vc[yh * NP3 + xl] + vc[yh * NP3 + xh]));
                    }
                }
              __syncthreads();
              // level 1: own rows + the upper halo row (needed by the last fine row of this slab)
              for (int t = tid; t < NBP * 9 * W1; t += T)
                {
                  const int X = t & (W1 - 1), qr = (t / W1) % 9, k = t / (9 * W1);
                  const int Y1 = 8 * rank + qr;
                  if (X >= 1 && Y1 >= 1 && Y1 <= W1 - 1)
                    {
                      const int     i  = (k * 10 + qr + 1) * W1 + X;
                      const double *vc = v2 + k * 6 * W2;
                      const int     xl = X >> 1, xh = (X + 1) >> 1;
                      const int     yl = (Y1 >> 1) - 4 * rank + 1, yh = ((Y1 + 1) >> 1) - 4 * rank + 1;
                      v1[i] = v1[i] * d1[qr * W1 + X] +
                              0.25 * ((vc[yl * W2 + xl] + vc[yl * W2 + xh]) + (vc[yh * W2 + xl] + vc[yh * W2 + xh]));
                    }
                }
              __syncthreads();
              ST_MARK(6)
              // fine level: z = r / D + P z_1, r.z
              {
                double acc[NBP];
#pragma unroll
                for (int k = 0; k < NBP; ++k)
                  acc[k] = 0.0;
                double di[4];
                if constexpr (TM)
                  {
                    double e8[8];
                    tmm::ld8(tm + Y::EOFF, e8); // all lanes: tcgen05.ld is warp-collective
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                      di[i] = e8[1 + i];
                  }
                else
                  {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                      di[i] = d0[(4 * g + i) * W + (colact ? jx : 1)];
                  }
                if (colact)
                  {
                    const int xl = jx >> 1, xh = (jx + 1) >> 1;
#pragma unroll
                    for (int k = 0; k < NBP; ++k)
                      {
                        // the 4 fine rows of a thread lie under 3 level-1 rows (buffer rows 2g+1 .. 2g+3)
                        double c1[3][2];
#pragma unroll
                        for (int a = 0; a < 3; ++a)
                          {
                            const double *vc = v1 + (k * 10 + 2 * g + 1 + a) * W1;
                            c1[a][0] = vc[xl], c1[a][1] = vc[xh];
                          }
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                          {
                            const int    al = i >> 1, ah = (i + 1) >> 1;
                            const double cv = 0.25 * ((c1[al][0] + c1[al][1]) + (c1[ah][0] + c1[ah][1]));
                            const double zv = fma(r[k][i], di[i], cv);
                            z[k][i]         = zv;
                            acc[k]          = fma(r[k][i], zv, acc[k]);
                          }
                      }
                    // z of the slab's first / last row to the neighbours: they update their halo copy
                    // of p themselves once beta is known (saves the barrier a push of p would need)
                    if (g == 0 && rank > 0)
                      {
#pragma unroll
                        for (int k = 0; k < NBP; ++k)
                          dsm::push(zh + (k * 2 + 1) * W + jx, mb + MB_RZ, rank - 1, z[k][0]);
                      }
                    if (g == 3 && rank + 1 < CS)
                      {
#pragma unroll
                        for (int k = 0; k < NBP; ++k)
                          dsm::push(zh + (k * 2 + 0) * W + jx, mb + MB_RZ, rank + 1, z[k][3]);
                      }
                  }
                allreduce(acc, 0, nn * NBP * (W - 1));
#pragma unroll
                for (int k = 0; k < NBP; ++k)
                  {
                    beta[k] = (it == 0 || done[k]) ? 0.0 : bpx::fast_div(acc[k], rz[k]);
                    rz[k]   = acc[k];
                  }
              }
              ++it;
              ST_MARK(7)

              // ---- p = z + beta p on the slab and, with the neighbours' z, on the two halo rows
              // (the same fma on the same bits as in the owning CTA)
              if (colact)
                {
#pragma unroll
                  for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int k = 0; k < NBP; ++k)
                      {
                        const int idx = (k * 18 + 4 * g + i + 1) * W + jx;
                        pS[idx]       = fma(beta[k], pS[idx], z[k][i]);
                      }
                  if (g == 1 && rank > 0)
                    {
#pragma unroll
                      for (int k = 0; k < NBP; ++k)
                        {
                          const int idx = (k * 18 + 0) * W + jx;
                          pS[idx]       = fma(beta[k], pS[idx], zh[(k * 2 + 0) * W + jx]);
                        }
                    }
                  if (g == 2 && rank + 1 < CS)
                    {
#pragma unroll
                      for (int k = 0; k < NBP; ++k)
                        {
                          const int idx = (k * 18 + 17) * W + jx;
                          pS[idx]       = fma(beta[k], pS[idx], zh[(k * 2 + 1) * W + jx]);
                        }
                    }
                }
              __syncthreads();
              ST_MARK(8)

              // ---- q = K p (9-point stencil, coefficients shared by the bases of the pass), p.q
              double q[TM ? 1 : NBP][4], pq[NBP]; // TM: q is parked in the residual staging buffer
#pragma unroll
              for (int k = 0; k < NBP; ++k)
                pq[k] = 0.0;
              if constexpr (TM)
                {
                  double e8[8];
                  tmm::ld8(tm + Y::EOFF, e8);
#pragma unroll
                  for (int kp = 0; kp < NBP / 2; ++kp) // two bases at a time: 18 window registers
                    {
                      double w[2][3][3];
                      if (colact)
                        {
#pragma unroll
                          for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                            for (int a = 0; a < 2; ++a)
#pragma unroll
                              for (int dx = 0; dx < 3; ++dx)
                                w[kk][a][dx] = pS[((2 * kp + kk) * 18 + 4 * g + a) * W + jx + dx - 1];
                        }
                      double kS = e8[0];
#pragma unroll
                      for (int i = 0; i < 4; ++i)
                        {
                          double c8[8]; // kc kE kW kN kNE kSW kNW kSE of this row
                          tmm::ld8(tm + Y::COFF + 16 * i, c8);
                          if (colact)
                            {
                              const int  cr     = 4 * g + i + 1;
                              const bool rowact = y0 + 4 * g + i >= 1;
#pragma unroll
                              for (int kk = 0; kk < 2; ++kk)
                                {
                                  const int k = 2 * kp + kk;
#pragma unroll
                                  for (int dx = 0; dx < 3; ++dx)
                                    w[kk][2][dx] = pS[(k * 18 + cr + 1) * W + jx + dx - 1];
                                  double yv = c8[0] * w[kk][1][1];
                                  yv        = fma(c8[1], w[kk][1][2], yv);
                                  yv        = fma(c8[2], w[kk][1][0], yv);
                                  yv        = fma(c8[3], w[kk][2][1], yv);
                                  yv        = fma(kS, w[kk][0][1], yv);
                                  yv        = fma(c8[4], w[kk][2][2], yv);
                                  yv        = fma(c8[5], w[kk][0][0], yv);
                                  yv        = fma(c8[6], w[kk][2][0], yv);
                                  yv        = fma(c8[7], w[kk][0][2], yv);
                                  if (!rowact)
                                    yv = 0.0;
                                  rS[(k * 17 + cr) * W + jx] = yv;
                                  pq[k]                      = fma(w[kk][1][1], yv, pq[k]);
#pragma unroll
                                  for (int dx = 0; dx < 3; ++dx)
                                    w[kk][0][dx] = w[kk][1][dx], w[kk][1][dx] = w[kk][2][dx];
                                }
                            }
                          kS = c8[3];
                        }
                    }
                }
              else if (colact)
                {
#pragma unroll
                  for (int k = 0; k < NBP; ++k)
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                      q[k][i] = 0.0;
                  double w[NBP][3][3];
#pragma unroll
                  for (int k = 0; k < NBP; ++k)
#pragma unroll
                    for (int a = 0; a < 2; ++a)
#pragma unroll
                      for (int dx = 0; dx < 3; ++dx)
                        w[k][a][dx] = pS[(k * 18 + 4 * g + a) * W + jx + dx - 1];
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    {
                      const int cr = 4 * g + i + 1;
#pragma unroll
                      for (int k = 0; k < NBP; ++k)
#pragma unroll
                        for (int dx = 0; dx < 3; ++dx)
                          w[k][2][dx] = pS[(k * 18 + cr + 1) * W + jx + dx - 1];
                      const double kc = cf[(0 * 17 + cr) * W + jx];
                      const double kE = cf[(1 * 17 + cr) * W + jx], kW = cf[(1 * 17 + cr) * W + jx - 1];
                      const double kN = cf[(2 * 17 + cr) * W + jx], kS = cf[(2 * 17 + cr - 1) * W + jx];
                      const double kNE = cf[(3 * 17 + cr) * W + jx], kSW = cf[(3 * 17 + cr - 1) * W + jx - 1];
                      const double kNW = cf[(4 * 17 + cr) * W + jx - 1], kSE = cf[(4 * 17 + cr - 1) * W + jx];
                      const bool   rowact = y0 + 4 * g + i >= 1;
#pragma unroll
                      for (int k = 0; k < NBP; ++k)
                        {
                          double yv = kc * w[k][1][1];
                          yv        = fma(kE, w[k][1][2], yv);
                          yv        = fma(kW, w[k][1][0], yv);
                          yv        = fma(kN, w[k][2][1], yv);
                          yv        = fma(kS, w[k][0][1], yv);
                          yv        = fma(kNE, w[k][2][2], yv);
                          yv        = fma(kSW, w[k][0][0], yv);
                          yv        = fma(kNW, w[k][2][0], yv);
                          yv        = fma(kSE, w[k][0][2], yv);
                          if (!rowact)
                            yv = 0.0;
                          q[k][i] = yv;
                          pq[k]   = fma(w[k][1][1], yv, pq[k]);
#pragma unroll
                          for (int dx = 0; dx < 3; ++dx)
                            w[k][0][dx] = w[k][1][dx], w[k][1][dx] = w[k][2][dx];
                        }
                    }
                }
              allreduce(pq, 1, 0);
              ST_MARK(9)

              // ---- x += alpha p ; r -= alpha q ; r.r
              {
                double alpha[NBP], acc[NBP];
#pragma unroll
                for (int k = 0; k < NBP; ++k)
                  alpha[k] = done[k] ? 0.0 : bpx::fast_div(rz[k], pq[k]), acc[k] = 0.0;
                if constexpr (TM)
                  {
#pragma unroll
                    for (int kp = 0; kp < NBP / 2; ++kp)
                      {
                        double x8[8]; // x of bases 2kp, 2kp+1, rows 0..3
                        tmm::ld8(tm + Y::XOFF + 16 * kp, x8);
                        if (colact)
                          {
#pragma unroll
                            for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                              for (int i = 0; i < 4; ++i)
                                {
                                  const int    k  = 2 * kp + kk;
                                  const double pv = pS[(k * 18 + 4 * g + i + 1) * W + jx];
                                  const double qv = rS[(k * 17 + 4 * g + i + 1) * W + jx];
                                  x8[4 * kk + i]  = fma(alpha[k], pv, x8[4 * kk + i]);
                                  const double rn = fma(-alpha[k], qv, r[k][i]);
                                  r[k][i]         = rn;
                                  acc[k]          = fma(rn, rn, acc[k]);
                                }
                          }
                        tmm::st8(tm + Y::XOFF + 16 * kp, x8);
                      }
                    tmm::wait_st();
                  }
                else if (colact)
                  {
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                      for (int k = 0; k < NBP; ++k)
                        {
                          const double pv = pS[(k * 18 + 4 * g + i + 1) * W + jx];
                          x[k][i]         = fma(alpha[k], pv, x[k][i]);
                          const double rn = fma(-alpha[k], q[k][i], r[k][i]);
                          r[k][i]         = rn;
                          acc[k]          = fma(rn, rn, acc[k]);
                        }
                  }
                stage_r();
                allreduce(acc, 2, rank > 0 ? NBP * (W - 1) : 0);
#pragma unroll
                for (int k = 0; k < NBP; ++k)
                  if (!done[k])
                    rr[k] = acc[k];
              }
              ST_MARK(10)
            }

          // ---- results of the pass
          if constexpr (TM)
            {
#pragma unroll
              for (int kp = 0; kp < NBP / 2; ++kp)
                {
                  double x8[8];
                  tmm::ld8(tm + Y::XOFF + 16 * kp, x8);
                  if (colact)
                    {
#pragma unroll
                      for (int kk = 0; kk < 2; ++kk)
#pragma unroll
                        for (int i = 0; i < 4; ++i)
                          if (y0 + 4 * g + i >= 1)
                            P.phi[((size_t)cell * 4 + NBP * pass + 2 * kp + kk) * N + (y0 + 4 * g + i) * np + jx] =
                              x8[4 * kk + i];
                    }
                }
            }
          else if (colact)
            {
#pragma unroll
              for (int i = 0; i < 4; ++i)
                {
                  const int y = y0 + 4 * g + i;
                  if (y >= 1)
                    {
#pragma unroll
                      for (int k = 0; k < NBP; ++k)
                        P.phi[((size_t)cell * 4 + NBP * pass + k) * N + y * np + jx] = x[k][i];
                    }
                }
            }
          if (rank == 0 && tid == 0)
            {
#pragma unroll
              for (int k = 0; k < NBP; ++k)
                {
                  const int sidx = cell * 4 + NBP * pass + k;
                  P.iters[sidx]  = itc[k] >= 0 ? itc[k] : it;
                  P.res[sidx]    = sqrt(rr[k]);
                  if (!(rr[k] <= P.tol2))
                    atomicMin(P.fail, sidx);
                }
            }
          // a fast CTA must not start the next pass (or exit) while a peer still reads what it pushed
          cluster.sync();
          ST_MARK(11)
        }
      ST_FLUSH
      if constexpr (TM)
        tmm::release(tm_base, warp);
    }

#ifndef MSB_EMU
    // n_clusters = cells of the launch (P.split = 0) or cells x passes (P.split = 1).  max_active != nullptr: only
    // report how many clusters of this kernel are co-resident on the device.
    template <int L, int NBP, bool TM>
    static cudaError_t
    launch(const Params &P, int n_clusters, cudaStream_t st, int *max_active = nullptr)
    {
      using Y = Lay<L, NBP, TM>;
      cudaError_t e = cudaFuncSetAttribute(solve_cluster_kernel<L, NBP, TM>,
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Y::smem_bytes);
      if (e != cudaSuccess)
        return e;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim            = dim3((unsigned)(Y::CS * (n_clusters > 0 ? n_clusters : 1)), 1, 1);
      cfg.blockDim           = dim3(Y::T, 1, 1);
      cfg.dynamicSmemBytes   = Y::smem_bytes;
      cfg.stream             = st;
      cudaLaunchAttribute at[1];
      at[0].id               = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = Y::CS;
      at[0].val.clusterDim.y = 1;
      at[0].val.clusterDim.z = 1;
      cfg.attrs              = at;
      cfg.numAttrs           = 1;
      if (max_active)
        return cudaOccupancyMaxActiveClusters(max_active, solve_cluster_kernel<L, NBP, TM>, &cfg);
      return cudaLaunchKernelEx(&cfg, solve_cluster_kernel<L, NBP, TM>, P);
    }
#endif
  } // namespace clus

#ifndef MSB_EMU
  bool
  cluster_tier_supported(int l)
  {
    return l >= 5 && l <= 7;
  }

  // the solve of all cells of the shard; s.d_dinv must hold the reciprocal Galerkin diagonals
  // (launch_solve_streamed computes them before dispatching here).  tmem = true: all four bases of a
  // cell in one pass, coefficients and x in tensor memory; false: two passes of two bases, shared
  // memory and registers only.
  cudaError_t
  launch_solve_cluster(const Shard &s, double tol, int max_iter, bool tmem, cudaStream_t st, int *n_launches)
  {
    clus::Params P;
    P.corners  = s.d_corners;
    P.q1coef   = s.d_q1coef;
    P.sten     = s.d_sten;
    P.dinv     = s.d_dinv;
    P.phi      = s.d_phi;
    P.iters    = s.d_iters;
    P.res      = s.d_res;
    P.fail     = s.d_fail;
    P.tol2     = tol * tol;
    P.max_iter = max_iter;
    P.cn       = (int)streamed_coarse_nodes(s.l);
    P.cell0    = 0;
    P.split    = 0;
    auto run = [&](bool tm, const clus::Params &Q, int n_clusters, int *max_active) -> cudaError_t {
      switch (s.l)
        {
          case 5:
            return tm ? clus::launch<5, 4, true>(Q, n_clusters, st, max_active) : clus::launch<5, 2, false>(Q, n_clusters, st, max_active);
          case 6:
            return tm ? clus::launch<6, 4, true>(Q, n_clusters, st, max_active) : clus::launch<6, 2, false>(Q, n_clusters, st, max_active);
          case 7:
            return tm ? clus::launch<7, 4, true>(Q, n_clusters, st, max_active) : clus::launch<7, 2, false>(Q, n_clusters, st, max_active);
        }
      return cudaErrorInvalidValue;
    };
    // Tail balancing.  Only `ma` clusters are co-resident (15 on this B200: 120 of its 148 SMs), so a shard runs in
    // ceil(cells / ma) waves and a short last wave leaves most of the GPU idle -- the reference's own default run has
    // 64 cells: 4 full waves + 4 clusters.  The cells of a short last wave go to TWO clusters each, one per pair of bases
    // (the two-bases flavour, one pass per cluster): the wave then takes ~0.55 of a full one (cfg1: 95.8 -> 103.1 k solves/s).
    int tail = 0;
    if (tmem && s.variant != 8) // (variant 8: no tail balancing, for the A/B)
      {
        static int  ma_of_l[8] = {0, 0, 0, 0, 0, 0, 0, 0}; // (one device type per process: queried once per local-mesh size)
        int         ma = ma_of_l[s.l];
        cudaError_t eq = cudaSuccess;
        if (ma == 0)
          {
            eq = run(true, P, 1, &ma);
            if (eq == cudaSuccess)
              ma_of_l[s.l] = ma;
          }
        if (eq == cudaSuccess && ma > 0)
          {
            const int t = s.n_cells % ma;
            if (t > 0 && 2 * t <= ma)
              tail = t;
          }
        else
          (void)cudaGetLastError();
      }
    cudaError_t e = cudaSuccess;
    if (s.n_cells - tail > 0)
      {
        e = run(tmem, P, s.n_cells - tail, nullptr);
        if (e == cudaSuccess)
          ++*n_launches;
      }
    if (e == cudaSuccess && tail > 0)
      {
        clus::Params Q = P;
        Q.cell0        = s.n_cells - tail;
        Q.split        = 1;
        e              = run(false, Q, 2 * tail, nullptr);
        if (e == cudaSuccess)
          ++*n_launches;
      }
    return e;
  }
#endif
} // namespace msb
