// msb_solve_fused.cu -- the WHOLE basis stage of one coarse cell in one CTA (64 x 64 local mesh).
//
// run() of the reference's basis object (diffusion_problem_basis.tpp:438-474) is
//   assemble_system (:159-242)  ->  4 x { condense (:461), PCG (:293-317), distribute (:308) }
//   ->  assemble_global_element_matrix (:245-285).
// Round 1 ran it as three kernels: the stencil was written to HBM (203 KB per cell), re-read by the solve
// and re-read together with Phi by the element-matrix kernel: ~53 GB of DRAM traffic per 65 536-cell
// step where only Phi (8.9 GB) has to leave the chip.  This kernel keeps everything between the cell's
// corner points and its (Phi, M, b) on the SM:
//
//   prologue  the 9-point stencil of the cell is computed node by node straight into shared memory
//             (2x2 Gauss, full tensor coefficient, separable sine tables: the arithmetic of
//             assemble_kernel, same summation order), then the Galerkin hierarchy, the exact inverse of
//             the 7x7 level (tensor memory) and the symmetric diagonal scaling, in place;
//   solve     the multilevel-preconditioned CG of msb_solve_bpx_tm.cu (two bases in flight, x / p / sqrt(d)
//             in tensor memory) with three changes, all measured leads of the round-1 profile:
//              - the four coefficient arrays are stored as two arrays of PAIRS {E,D2} and {N,D1}: a stencil
//                row needs 4 LDS.128 instead of 7 LDS.64 (the crossbar moves a 128-bit access at twice the
//                byte rate of a 64-bit one: scripts/probes/onchip_peaks.cu, 255 vs 128 B/clk/SM);
//              - r.z is NOT reduced after the fine prolongation: with z = r + D^1/2 P z_1 it equals
//                |r|^2 + u_1 . z_1 (u_1 = restricted residual, z_1 = level-1 correction), and both parts are
//                known when level 1 is prolonged.  beta is therefore available BEFORE the fine
//                prolongation, which is fused with the direction update: one block barrier, one
//                reduction round trip and one pass over the strip fewer per iteration;
//              - the pre-summed residual strips are stored with a 5-entry pad between the even and the
//                odd columns (bank-conflict free, see Presum);
//   epilogue  M = Phi^T K Phi and b = Phi^T F from data that is still on chip.  K phi_j vanishes on interior
//             rows up to the final residual (<= 1e-12), so
//                 M_ij = sum_{a on the boundary} g_i(a) (K phi_j)(a),
//             (K phi_j)(a) = sum_nbr K_ab (phi_j(b) - g_j(a)) using the zero row sums of K; the interior
//             neighbours enter through the scaled iterate, K_ac (phi_j(c) - g_j(a)) =
//             Ahat_ac (xhat_j(c) - sqrt(d_c) g_j(a)): no division, no second pass over Phi.
//
// Scaling convention: S = diag(s), s = d^-1/2 on interior nodes and 1 on boundary nodes; the stored edge
// coefficients are S K S.  An interior-boundary edge is thus scaled by its interior end only -- exactly the
// factor the condensed right-hand side -D^-1/2 K_IB g_B needs -- and boundary-boundary edges stay unscaled.
//
// Eligibility (msb_run decides on the host): every coarse cell axis-aligned (what the reference's refined
// hyper_cube produces) and an analytic coefficient kind.  Anything else takes the three-kernel path.
// No tensor-core instruction is issued; TMEM is used as memory only.
#include <limits.h>

#define MSB_STAGE_ARRAY g_msb_stage_cycles_fu
#include "msb_bpx_common.cuh"
#include "msb_coeff.cuh"
#include "msb_fused.cuh"

#ifdef MSB_STAGE_TIMERS
__device__ unsigned long long g_msb_stage_cycles_fu[16];
extern "C" int
msb_debug_stage_cycles_fu(unsigned long long *out, int reset)
{
  cudaError_t e = cudaMemcpyFromSymbol(out, g_msb_stage_cycles_fu, sizeof(unsigned long long) * 16);
  if (e == cudaSuccess && reset)
    {
      unsigned long long z[16] = {0};
      e = cudaMemcpyToSymbol(g_msb_stage_cycles_fu, z, sizeof z);
    }
  return (int)e;
}
#endif

namespace msb
{
  namespace fused
  {
    using namespace bpx;

    // 16-byte pair access
    __device__ __forceinline__ void
    ld2(const double *p, int idx, double &a, double &b)
    {
      const double2 t = *reinterpret_cast<const double2 *>(p + 2 * (size_t)idx);
      a = t.x, b = t.y;
    }
    __device__ __forceinline__ void
    st2(double *p, int idx, double a, double b)
    {
      *reinterpret_cast<double2 *>(p + 2 * (size_t)idx) = make_double2(a, b);
    }

    // ---- One entry of a fine element matrix on an hx x hy rectangle from the tensor coefficient at the 2x2 Gauss
    // points q = qx + 2 qy (abscissae g_qx, g_qy):
    //   K_ij = sum_q [ dNx_i dNx_j a00 hy/(4hx) + (dNx_i dNy_j + dNy_i dNx_j) (a01+a10)/8 + dNy_i dNy_j a11 hx/(4hy) ]_q .
    // The x-gradients of the bilinear shape functions depend on eta_q only and the y-gradients on xi_q only, so the
    // first and the last term need the coefficient only SUMMED over qx resp. qy:
    //   S[0] = c00_0 + c00_1 (eta = g0), S[1] = c00_2 + c00_3 (eta = g1), S[2] = c11_0 + c11_2 (xi = g0), S[3] = c11_1 + c11_3
    // -- 4 fused multiply-adds per entry for an isotropic coefficient (c01 = 0), 8 for a full tensor, instead
    // of 12 (assemble_kernel's order; the results differ in the last bits).
    template <int I, int J, bool ISO>
    __device__ __forceinline__ double
    kentry(const double (&S)[4], const double (&c01)[4])
    {
      constexpr double G0 = 0.21132486540518711775, G1 = 0.78867513459481288225;
      constexpr double dx0[4] = {-(1 - G0), (1 - G0), -G0, G0}, dx1[4] = {-(1 - G1), (1 - G1), -G1, G1}; // dNx at eta = g0, g1
      constexpr double dy0[4] = {-(1 - G0), -G0, (1 - G0), G0}, dy1[4] = {-(1 - G1), -G1, (1 - G1), G1}; // dNy at xi = g0, g1
      double           k = dx0[I] * dx0[J] * S[0];
      k                  = fma(dx1[I] * dx1[J], S[1], k);
      k                  = fma(dy0[I] * dy0[J], S[2], k);
      k                  = fma(dy1[I] * dy1[J], S[3], k);
      if constexpr (!ISO)
        {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            {
              const double *dx = (q >> 1) ? dx1 : dx0, *dy = (q & 1) ? dy1 : dy0;
              k = fma(dx[I] * dy[J] + dy[I] * dx[J], c01[q], k);
            }
        }
      return k;
    }

    // The four edge-coefficient arrays of the n x n fine cells, cell index i = y n + x:
    //   E  (x,y)-(x+1,y)    D2 (x+1,y)-(x,y+1)    N  (x,y)-(x,y+1)    D1 (x,y)-(x+1,y+1)
    // MSB_FUSED_PAIRED = 0: four planes, sA = [E | D2], sB = [N | D1]: a stencil row costs 7 LDS.64.
    // = 1: interleaved pairs sA[i] = {E, D2}, sB[i] = {N, D1}: 4 LDS.128 per row, one of the four half used -- measured
    // slower (profiles/r02c_ab_flavours.txt), kept for the A/B only.
    // = 2: four planes whose ROWS are interleaved in pairs, entry (x, y) of a plane at ((y / 2) n + x) 2 + (y & 1): the
    // thread that marches up its strip fetches the coefficients of two consecutive rows with one LDS.128 -- 35 LDS.128
    // per strip of 8 rows instead of 59 LDS.64, every byte used.  Measured 4 % SLOWER on both instantiations (5920 cells:
    // target 14.13 vs 13.58 ms, cfg4 7.34 vs 7.04 ms): inside this kernel a 128-bit access costs its four 128-byte
    // wavefronts, whatever the stand-alone probe (scripts/probes/onchip_peaks.cu: 255 B/clk/SM for LDS.128) suggests.
    // Kept for the A/B only.
#ifndef MSB_FUSED_PAIRED
#  define MSB_FUSED_PAIRED 0
#endif
    template <int n>
    struct Coef
    {
      static constexpr bool PAIRED = MSB_FUSED_PAIRED == 1, ROWPAIR = MSB_FUSED_PAIRED == 2;
      __device__ static __forceinline__ int
      rp(int i) // row-pair position of cell i = y n + x inside a plane
      {
        const int y = i / n, x = i % n;
        return (((y >> 1) * n + x) << 1) + (y & 1);
      }
      __device__ static __forceinline__ int
      e(int i)
      {
        return PAIRED ? 2 * i : (ROWPAIR ? rp(i) : i);
      }
      __device__ static __forceinline__ int
      hi(int i) // D2 in sA, D1 in sB
      {
        return PAIRED ? 2 * i + 1 : n * n + (ROWPAIR ? rp(i) : i);
      }
      // both rows 2 m, 2 m + 1 of column x of a plane (ROWPAIR)
      __device__ static __forceinline__ double2
      pair(const double *plane, int m, int x)
      {
        return *reinterpret_cast<const double2 *>(plane + ((m * n + x) << 1));
      }
    };

    // edge coefficient of node (x,y) towards (x+ex, y+ey), for edges that live inside the n x n cell arrays
    // (every edge with at least one interior end)
    template <int n>
    __device__ __forceinline__ double
    eget(const double *sA, const double *sB, int x, int y, int ex, int ey)
    {
      using K = Coef<n>;
      if (ey == 0)
        return sA[K::e(y * n + (ex > 0 ? x : x - 1))];
      if (ex == 0)
        return sB[K::e((ey > 0 ? y : y - 1) * n + x)];
      if (ex == ey)
        return sB[K::hi((ey > 0 ? y : y - 1) * n + (ex > 0 ? x : x - 1))];
      return sA[K::hi(ey > 0 ? y * n + x - 1 : (y - 1) * n + x)];
    }
    // ... and any edge of the mesh, including the boundary-boundary edges of the top row / right column
    template <int n>
    __device__ __forceinline__ double
    eget_full(const double *sA, const double *sB, const double *sEb, const double *sNb, int x, int y, int ex, int ey)
    {
      if (ey == 0 && y == n)
        return sEb[ex > 0 ? x : x - 1];
      if (ex == 0 && x == n)
        return sNb[ey > 0 ? y : y - 1];
      return eget<n>(sA, sB, x, y, ex, ey);
    }

    // Warp-level sums of TEN values for the element-matrix epilogue: three transposing exchanges leave one of the
    // first eight values per group of four lanes (9 double shuffles instead of 40), the last two share one
    // butterfly (5 instead of 10).  Lane l with (l & 3) == 0 returns the warp total of value (l >> 2) & 7 in `e`;
    // lanes 0 and 16 return the totals of values 8 and 9 in `t`.
    __device__ __forceinline__ void
    warp_sum10(const double (&v)[10], int lane, double &e, double &t)
    {
      const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
      double     a[4], c[2];
#pragma unroll
      for (int i = 0; i < 4; ++i)
        a[i] = (b4 ? v[4 + i] : v[i]) + __shfl_xor_sync(0xffffffffu, b4 ? v[i] : v[4 + i], 16);
#pragma unroll
      for (int i = 0; i < 2; ++i)
        c[i] = (b3 ? a[2 + i] : a[i]) + __shfl_xor_sync(0xffffffffu, b3 ? a[i] : a[2 + i], 8);
      e = (b2 ? c[1] : c[0]) + __shfl_xor_sync(0xffffffffu, b2 ? c[0] : c[1], 4);
      e += __shfl_xor_sync(0xffffffffu, e, 2);
      e += __shfl_xor_sync(0xffffffffu, e, 1);
      t = (b4 ? v[9] : v[8]) + __shfl_xor_sync(0xffffffffu, b4 ? v[8] : v[9], 16);
#pragma unroll
      for (int off = 8; off > 0; off >>= 1)
        t += __shfl_xor_sync(0xffffffffu, t, off);
    }

    // NL = 6: 64 x 64 local mesh, 512 threads, one CTA per SM, the inverse of the 7x7 level in tensor memory.
    // NL = 5: 32 x 32 local mesh, 128 threads, THREE CTAs per SM (69 KB of shared memory each: the inverse is kept as
    // a packed lower triangle in shared memory, 9.8 instead of 19.2 KB).
    template <int NL_>
    struct Cfg
    {
      static constexpr int NL = NL_, NRHS = 2;
      static_assert(NL == 5 || NL == 6, "fused stage: 32 x 32 and 64 x 64 local meshes");
      using L                    = Levels<NL>;
      static constexpr int n     = 1 << NL, np = n + 1, N = np * np;
      static constexpr int RPT   = 8;
      static constexpr int WX    = (n - 1 + 31) / 32;
      static constexpr int WY    = n / RPT;
      static constexpr int NWARP = WX * WY;
      static constexpr int THREADS = 32 * NWARP;
      static constexpr int NCH   = RPT / 4; // TMEM chunks of 4 rows x 2 bases = 8 doubles
      static constexpr int CN    = L::CN;
      static constexpr int PAD   = 5;
      using PS                   = Presum<NL, NRHS, RPT, PAD>;
      // per-thread TMEM columns: x | p_old | rhat or q (flavours 1, 2) | sqrt(d)
      static constexpr int XOFF = 0, POFF = 2 * RPT * NRHS, ROFF = 4 * RPT * NRHS, SOFF = 6 * RPT * NRHS;
      static constexpr int TCOLS = SOFF + 2 * RPT;
      using X7 = Exact7<THREADS>;
      static constexpr bool GI_TMEM = NL == 6; // where the 49 x 49 inverse of the 7x7 level lives
      static constexpr int  MOFF = (NWARP / 4) * TCOLS, MCOLS = GI_TMEM ? ((X7::WARPS + 3) / 4) * 16 * X7::NCHK : 0;
      static constexpr int  TMEM_COLS = NL == 6 ? 512 : 128; // per CTA (a power of two)
      static constexpr int  RED = 10 * NWARP + 8; // the widest reduction: 10 values (element matrix columns)
      // shared memory, in doubles
      static constexpr int o_A = 0, o_B = o_A + 2 * n * n, o_P = o_B + 2 * n * n, o_V = o_P + NRHS * N;
      static constexpr int o_red = o_V + NRHS * CN, o_eb = o_red + RED, o_nb = o_eb + n, o_gi = o_nb + n;
      static constexpr int o_di  = o_gi + (GI_TMEM ? 0 : EXACT7_TRI + 1); // (NL = 5: packed lower triangle)
      static constexpr size_t smem_bytes = sizeof(double) * (size_t)o_di + sizeof(float) * (size_t)CN;
      // prologue scratch inside the vector buffers (o_P .. o_red is one contiguous area)
      static constexpr int o_kc  = o_P + 5 * L::lvl_off(2); // fine diagonal, N doubles (dead before level 2 is built)
      static constexpr int o_tab = o_kc + N;                // sine tables, 8 n doubles
      // scratch of the banded factorisation (and, NL = 6, of the inverse before it moves to tensor memory)
      static constexpr int o_x7  = o_V;
      static_assert(o_tab + 8 * n <= o_red, "sine tables");
      static_assert(5 * CN <= NRHS * N, "Galerkin scratch must fit the vector buffer");
      static_assert((GI_TMEM ? 49 * 49 : 0) + EXACT7_SCRATCH <= NRHS * CN, "scratch of the 7x7 inverse must fit the coarse vectors");
      static_assert(NRHS * PS::ENTRIES <= NRHS * N, "pre-summed staging must fit the vector buffer");
      static_assert(4 * 4 * n <= NRHS * CN, "Dirichlet table must fit the coarse vectors");
      static_assert((NWARP + 3) / 4 * TCOLS + MCOLS <= TMEM_COLS, "tensor memory columns");
      static_assert(smem_bytes <= 232448, "shared memory");
      static_assert(NL == 6 || 3 * (smem_bytes + 1024) <= 233472, "NL = 5: three CTAs per SM");
    };

    // RMODE (A/B flavours, FusedParams::flavor; measured on 5920 target cells, profiles/r02c_ab_flavours.txt):
    //   0  residual and q = Ahat p in registers                                   17.78 ms   <- default
    //   1  residual in tensor memory (fewest registers, two more TMEM round trips) 18.16 ms
    //   2  residual in registers, q through tensor memory four rows at a time      17.93 ms
    // A flattened coarse chain (level 2 and the 7x7 level restricted straight from level 1 by different warps,
    // both interpolated straight back: two block barriers instead of five) was 3 % SLOWER than
    // bpx::coarse_correction in every flavour and is not kept: the stages it removes are short, the ones it
    // fattens (all threads) are not.
    // SPLIT: one CTA per (cell, pair of bases) for the short last wave of a small shard (launch_stage_fused); a template
    // parameter, not a run-time flag: with run-time loop bounds the ordinary kernel was 1.8 % slower on the target workload.
    template <int NL_, int RMODE, bool SPLIT = false>
    __global__ void __launch_bounds__(Cfg<NL_>::THREADS, NL_ == 6 ? 1 : 3)
    solve_fused_kernel(FusedParams P)
    {
      constexpr bool RTMEM = RMODE == 1, QTMEM = RMODE == 2;
      using C             = Cfg<NL_>;
      using L             = typename C::L;
      using PS            = typename C::PS;
      constexpr int THREADS = C::THREADS, NL = C::NL, NRHS = C::NRHS, n = C::n, np = C::np, N = C::N;
      constexpr int NWARP = C::NWARP, WX = C::WX, RPT = C::RPT, NCH = C::NCH, CN = C::CN;

#ifndef MSB_EMU
      extern __shared__ __align__(16) double smem[];
      __shared__ uint32_t s_tmem_base;
      uint32_t *tmem_slot = &s_tmem_base;
#else
      double   *smem      = emu::smem();
      uint32_t *tmem_slot = nullptr;
#endif
      double *sA   = smem + C::o_A;   // [n*n] pairs {E, D2}
      double *sB   = smem + C::o_B;   // [n*n] pairs {N, D1}
      double *sP   = smem + C::o_P;   // [N][2]: p (stencil) or pre-summed u (restriction), zero halo
      double *sV   = smem + C::o_V;   // [CN][2] coarse residuals / corrections
      double *sRed = smem + C::o_red; // reduction buffer
      double *sEb  = smem + C::o_eb;  // [n] E edges of the top boundary row (unscaled)
      double *sNb  = smem + C::o_nb;  // [n] N edges of the right boundary column (unscaled)
      float  *sDi  = reinterpret_cast<float *>(smem + C::o_di); // [CN] 1 / Galerkin diagonal
      double *sKC  = smem + C::o_kc;  // prologue only
      double *tsx  = smem + C::o_tab, *tsy = tsx + 4 * n;

      const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
      // split launches (the short last wave of a small shard, launch_stage_fused): CTA -> (cell, pair of bases); the
      // prologue is then paid by both CTAs of a cell
      const int cell   = SPLIT ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
      const int grp_lo = SPLIT ? (int)(blockIdx.x & 1) : 0, grp_hi = SPLIT ? grp_lo + 1 : 4 / Cfg<NL_>::NRHS;

      const double *crn = P.corners + 8 * (size_t)cell;
      const double *q1  = P.q1coef + 16 * (size_t)cell;
      const double  hx = (crn[2] - crn[0]) / n, hy = (crn[5] - crn[1]) / n;

      ST_DECL
      // ---------------------------------------------------------------- tensor memory
      const uint32_t tmem_base = tmem::alloc(tmem_slot, warp, C::TMEM_COLS);
      // this warp's 32-lane quarter (bits 31:16) and this warp's column block (bits 15:0)
      const uint32_t tm = tmem_base + tmem::lane_quarter(warp) + (uint32_t)((warp >> 2) * C::TCOLS);

      // ================================================================ prologue: assemble_system on chip
      // (0) sine tables of a separable coefficient: one value per (fine column / row, Gauss point), the
      //     abscissae by the expressions of assemble_kernel (bilinear image of the reference point)
      const bool separable = P.coef.separable();
      if (separable)
        {
          const double g0 = 0.5 - 0.5 / sqrt(3.0), g1 = 0.5 + 0.5 / sqrt(3.0);
          for (int t = tid; t < 8 * n; t += THREADS)
            {
              const bool   isx = t < 4 * n;
              const int    u = isx ? t : t - 4 * n, k = u >> 2, q = u & 3;
              const int    ix = isx ? k : 0, iy = isx ? 0 : k;
              const double xi = (q & 1) ? g1 : g0, eta = (q >> 1) ? g1 : g0;
              const double Nv[4] = {(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta};
              double       xq = 0, yq = 0;
#pragma unroll
              for (int vv = 0; vv < 4; ++vv)
                {
                  double px, py;
                  fine_vertex(crn, n, ix + (vv & 1), iy + (vv >> 1), px, py);
                  xq += px * Nv[vv];
                  yq += py * Nv[vv];
                }
              (isx ? tsx : tsy)[q * n + k] = P.coef.sine_term(isx ? xq : yq); // [Gauss point][column]: conflict-free reads
            }
          __syncthreads();
        }
      // (1) node stencils: K_e entries of the <= 4 adjacent fine cells, gathered in the order SW, SE, NW, NE
      {
        const double rxx = 0.25 * hy / hx, ryy = 0.25 * hx / hy;
        // every kind but the reference's rotated tensor is a scalar times the identity
        const bool iso = P.coef.kind != MSB_COEFF_REFERENCE;
        auto node_stencil = [&](auto iso_c, int t) {
          constexpr bool ISO = decltype(iso_c)::value;
          const int      jx = t % np, jy = t / np;
          double         kc = 0, kE = 0, kN = 0, kd1 = 0, kd2 = 0;
          auto cell_sums = [&](int ix, int iy, double(&S)[4], double(&c01)[4]) {
            double c00[4], c11[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
              {
                constexpr double G0 = 0.21132486540518711775, G1 = 0.78867513459481288225;
                double           a00, a01, a10, a11;
                if (separable)
                  P.coef.from_sines(tsx[q * n + ix], tsy[q * n + iy], a00, a01, a10, a11);
                else
                  P.coef(crn[0] + (ix + ((q & 1) ? G1 : G0)) * hx, crn[1] + (iy + ((q >> 1) ? G1 : G0)) * hy, a00, a01,
                         a10, a11);
                c00[q] = a00 * rxx, c01[q] = 0.125 * (a01 + a10), c11[q] = a11 * ryy;
              }
            S[0] = c00[0] + c00[1], S[1] = c00[2] + c00[3], S[2] = c11[0] + c11[2], S[3] = c11[1] + c11[3];
          };
          double S[4], c01[4];
          if (jx > 0 && jy > 0) // SW cell: the node is its vertex 3
            {
              cell_sums(jx - 1, jy - 1, S, c01);
              kc += kentry<3, 3, ISO>(S, c01);
            }
          if (jx < n && jy > 0) // SE cell: vertex 2
            {
              cell_sums(jx, jy - 1, S, c01);
              kc += kentry<2, 2, ISO>(S, c01);
              kE += kentry<2, 3, ISO>(S, c01);
            }
          if (jx > 0 && jy < n) // NW cell: vertex 1
            {
              cell_sums(jx - 1, jy, S, c01);
              kc += kentry<1, 1, ISO>(S, c01);
              kN += kentry<1, 3, ISO>(S, c01);
            }
          if (jx < n && jy < n) // NE cell: vertex 0
            {
              cell_sums(jx, jy, S, c01);
              kc += kentry<0, 0, ISO>(S, c01);
              kE += kentry<0, 1, ISO>(S, c01);
              kN += kentry<0, 2, ISO>(S, c01);
              kd1 = kentry<0, 3, ISO>(S, c01);
              kd2 = kentry<1, 2, ISO>(S, c01);
            }
          sKC[t] = kc;
          if (jx < n && jy < n)
            {
              using K = Coef<n>;
              sA[K::e(jy * n + jx)] = kE, sA[K::hi(jy * n + jx)] = kd2;
              sB[K::e(jy * n + jx)] = kN, sB[K::hi(jy * n + jx)] = kd1;
            }
          else if (jx < n)
            sEb[jx] = kE; // top boundary row: boundary-boundary edges
          else if (jy < n)
            sNb[jy] = kN; // right boundary column
        };
        if (iso)
          for (int t = tid; t < N; t += THREADS)
            node_stencil(std::true_type{}, t);
        else
          for (int t = tid; t < N; t += THREADS)
            node_stencil(std::false_type{}, t);
      }
      __syncthreads();
      ST_MARK(12)
      // (2) level 1 of the Galerkin hierarchy of the UNSCALED interior operator from the fine stencil
      double *G = sP;
      {
        constexpr int npl = (n >> 1) + 1, Nl = npl * npl, nin = npl - 2;
        for (int t = tid; t < 5 * Nl; t += THREADS)
          G[t] = 0.0;
        for (int t = tid; t < Nl; t += THREADS)
          sDi[t] = 0.0f;
        __syncthreads();
        auto fine = [&](int ix, int iy, int ex, int ey) {
          return (ex == 0 && ey == 0) ? sKC[iy * np + ix] : eget<n>(sA, sB, ix, iy, ex, ey);
        };
        for (int t = tid; t < nin * nin; t += THREADS)
          {
            const int X = 1 + t % nin, Y = 1 + t / nin, i = Y * npl + X;
            double    a[5];
            galerkin_row_of(fine, X, Y, a);
            G[ST_KC * Nl + i] = a[0];
            if (X < nin)
              G[ST_KE * Nl + i] = a[1];
            if (Y < nin)
              G[ST_KN * Nl + i] = a[2];
            if (X < nin && Y < nin)
              G[ST_KD1 * Nl + i] = a[3];
            if (X > 1 && Y < nin)
              G[ST_KD2 * Nl + i - 1] = a[4];
            sDi[i] = (float)(1.0 / a[0]);
          }
      }
      // ---------------------------------------------------------------- ownership
      const int  wx = warp % WX, wy = warp / WX;
      const int  X  = 1 + 32 * wx + lane;
      const int  Y0 = 1 + RPT * wy;
      const bool colok = X <= n - 1;
      // (3) sqrt(d) of the owned DoFs -> tensor memory (kept for all four bases of the cell) ...
      {
        double sq8[8];
#pragma unroll
        for (int j = 0; j < 8; ++j)
          {
            const int y = Y0 + j;
            sq8[j]      = (colok && y <= n - 1) ? sqrt(sKC[y * np + X]) : 0.0;
          }
        tmem::st8(tm + C::SOFF, sq8);
      }
      __syncthreads();
      //     ... and s = d^-1/2 (interior) / 1 (boundary) in place
      for (int t = tid; t < N; t += THREADS)
        {
          const int jx = t % np, jy = t / np;
          sKC[t] = (jx == 0 || jx == n || jy == 0 || jy == n) ? 1.0 : rsqrt(sKC[t]);
        }
      __syncthreads();
      // (4) S K S in place
      for (int i = tid; i < n * n; i += THREADS)
        {
          const int    x = i % n, y = i / n, g = y * np + x;
          const double s00 = sKC[g], s10 = sKC[g + 1], s01 = sKC[g + np], s11 = sKC[g + np + 1];
          using K = Coef<n>;
          sA[K::e(i)] *= s00 * s10;
          sA[K::hi(i)] *= s10 * s01;
          sB[K::e(i)] *= s00 * s01;
          sB[K::hi(i)] *= s00 * s11;
        }
      __syncthreads();
      ST_MARK(13)
      // (5) levels 2 .. of the hierarchy (the fine diagonal is dead: its place is the scratch of these levels)
      {
        const double *Sf  = G;
        int           npf = (n >> 1) + 1, Nf = npf * npf, goff = 5 * Nf;
#pragma unroll 1
        for (int l = 2; l <= L::LEVELS; ++l)
          {
            const int npl = (n >> l) + 1, Nl = npl * npl, nin = npl - 2;
            double   *Gl = G + goff;
            for (int t = tid; t < 5 * Nl; t += THREADS)
              Gl[t] = 0.0;
            for (int t = tid; t < Nl; t += THREADS)
              sDi[goff / 5 + t] = 0.0f;
            __syncthreads();
            for (int t = tid; t < nin * nin; t += THREADS)
              {
                const int Xc = 1 + t % nin, Yc = 1 + t / nin, i = Yc * npl + Xc;
                double    a[5];
                galerkin_row(Sf, npf, Nf, Xc, Yc, a);
                Gl[ST_KC * Nl + i] = a[0];
                if (Xc < nin)
                  Gl[ST_KE * Nl + i] = a[1];
                if (Yc < nin)
                  Gl[ST_KN * Nl + i] = a[2];
                if (Xc < nin && Yc < nin)
                  Gl[ST_KD1 * Nl + i] = a[3];
                if (Xc > 1 && Yc < nin)
                  Gl[ST_KD2 * Nl + i - 1] = a[4];
                sDi[goff / 5 + i] = (float)(1.0 / a[0]);
              }
            __syncthreads();
            Sf   = Gl;
            npf  = npl;
            Nf   = Nl;
            goff += 5 * Nl;
          }
      }
      ST_MARK(14)
      // (6) exact coarse solve: the 49 x 49 inverse of the 7x7-level operator, built in the coarse-vector
      //     area and parked in tensor memory
      using X7            = typename C::X7;
      [[maybe_unused]] const uint32_t tmat =
        tmem_base + tmem::lane_quarter(warp) + (uint32_t)(C::MOFF + (warp >> 2) * 16 * X7::NCHK);
      [[maybe_unused]] const double *sGiv = smem + C::o_gi; // NL = 5: the inverse stays in shared memory
      if constexpr (C::GI_TMEM)
        {
          double *sGi = smem + C::o_x7, *sBand = sGi + 49 * 49;
          exact7_build<THREADS>(G + 5 * L::lvl_off(L::LW + 1), sGi, sBand, tid);
          if (warp < X7::WARPS)
            {
#pragma unroll
              for (int c = 0; c < X7::NCHK; ++c)
                {
                  double g[8];
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    g[i] = X7::fetch(sGi, tid, c, i);
                  tmem::st8(tmat + 16 * c, g);
                }
            }
          tmem::wait_st();
          __syncthreads();
        }
      else
        exact7_build<THREADS, true>(G + 5 * L::lvl_off(L::LW + 1), smem + C::o_gi, smem + C::o_x7, tid);
      ST_MARK(0)

      // boundary nodes in walking order (4 n of them)
      auto boundary_node = [&](int t, int &jx, int &jy) {
        if (t < n)
          jx = t, jy = 0;
        else if (t < 2 * n)
          jx = n, jy = t - n;
        else if (t < 3 * n)
          jx = n - (t - 2 * n), jy = n;
        else
          jx = 0, jy = n - (t - 3 * n);
      };
      // the pre-summed residual strips share the vector buffer with p and overwrite parts of p's zero halo:
      // restore the halo whenever p is rewritten
      auto zero_halo = [&]() {
        const double zero2[NRHS] = {0.0, 0.0};
        for (int t = tid; t < 4 * n; t += THREADS)
          {
            int jx, jy;
            boundary_node(t, jx, jy);
            stv<NRHS>(sP, jy * np + jx, zero2);
          }
      };

      // Dirichlet data: the four coarse Q1 shape functions (BasisQ1, basis_q1.tpp:116-133) at the 4 n boundary
      // nodes, tabulated in the coarse-vector area whenever that is free (right-hand side, epilogue)
      double *sG = sV; // [4 n][4]
      auto fill_boundary_table = [&]() {
        for (int t = tid; t < 4 * n; t += THREADS)
          {
            int jx, jy;
            boundary_node(t, jx, jy);
            double px, py;
            fine_vertex(crn, n, jx, jy, px, py);
            st2(sG, 2 * t, basis_q1_value(q1, 0, px, py), basis_q1_value(q1, 1, px, py));
            st2(sG, 2 * t + 1, basis_q1_value(q1, 2, px, py), basis_q1_value(q1, 3, px, py));
          }
      };
      // position of boundary node (bx, by) in the walking order of boundary_node()
      auto boundary_index = [&](int bx, int by) {
        return by == 0 ? bx : (bx == n ? n + by : (by == n ? 3 * n - bx : 4 * n - by));
      };

#pragma unroll 1
      for (int grp = grp_lo; grp < grp_hi; ++grp)
        {
          const int rhs0 = grp * NRHS;
          // (a) Initial guess: the coarse Q1 shape function itself, x_0 = g on the interior nodes (the exact
          //     solution for a constant coefficient; for the oscillating coefficients of the target workload it
          //     saves 4-5 of 26 iterations, profiles/r02e_*).  In the scaled variables what = S^-1 g (sqrt(d) g
          //     inside, g on the boundary) and rhat_0 = -S K g = -(S K S) what: ONE pass of the stencil sweep
          //     below over what, boundary values in the halo, gives the initial residual -- iteration 0 runs
          //     it with p = what and alpha = 1 (x_0 = 0 + 1 * what, rhat_0 = 0 - 1 * Ahat what, beta = 0).
          //     The condensed right-hand side -D^-1/2 K_IB g_B is the boundary part of that product.
          [[maybe_unused]] double rreg[RPT][NRHS]; // the residual when it lives in registers (!RTMEM)
          {
            double sq8[8];
            tmem::ld8(tm + C::SOFF, sq8);
#pragma unroll
            for (int c = 0; c < NCH; ++c)
              {
                double w8[8];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj)
                  {
                    const int j = 4 * c + jj, y = Y0 + j;
                    double    px, py, w[NRHS];
                    fine_vertex(crn, n, colok ? X : n - 1, y <= n - 1 ? y : n - 1, px, py);
#pragma unroll
                    for (int k = 0; k < NRHS; ++k)
                      {
                        w[k]           = sq8[j] * basis_q1_value(q1, rhs0 + k, px, py); // sqrt(d) = 0 beyond the mesh
                        w8[2 * jj + k] = w[k];
                        rreg[j][k]     = 0.0;
                      }
                    if (colok && y <= n - 1)
                      stv<NRHS>(sP, y * np + X, w);
                  }
                const double zero8[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                tmem::st8(tm + C::POFF + 16 * c, w8);
                tmem::st8(tm + C::XOFF + 16 * c, zero8);
                if constexpr (RMODE != 0)
                  tmem::st8(tm + C::ROFF + 16 * c, zero8); // rhat (1) / q of the threads beyond the last column (2)
              }
            for (int t = tid; t < 4 * n; t += THREADS)
              {
                int jx, jy;
                boundary_node(t, jx, jy);
                double px, py, w[NRHS];
                fine_vertex(crn, n, jx, jy, px, py);
#pragma unroll
                for (int k = 0; k < NRHS; ++k)
                  w[k] = basis_q1_value(q1, rhs0 + k, px, py);
                stv<NRHS>(sP, jy * np + jx, w);
              }
            for (int i = tid; i < NRHS * CN; i += THREADS)
              sV[i] = 0.0;
            tmem::wait_st();
          }
          __syncthreads();
          ST_MARK(1)

          double rho[NRHS] = {1.0, 1.0}, exact[NRHS] = {0.0, 0.0}, alpha[NRHS] = {0.0, 0.0};
          bool   done[NRHS] = {false, false};
          int    kit[NRHS]  = {0, 0};
          bool   all_done   = false;
          int    it         = 0;

          // ------------------------------------------------------------ one PCG iteration =
          //   [stencil, alpha, r update]   (it = 0: the initial residual, see (a))
          //   stage u = D^1/2 rhat -> coarse levels -> (r.z, |r|^2) -> beta -> p = z + beta p_old, x += alpha p_old
#pragma unroll 1
          for (;;)
            {
              double q[RPT][NRHS];
                {
                  // ---- q = Ahat p for both bases: one set of coefficient loads per stencil row
                  double pq[NRHS] = {0.0, 0.0};
#pragma unroll
                  for (int j = 0; j < RPT; ++j)
                    q[j][0] = q[j][1] = 0.0;
                  {
                    // Every lane runs the sweep (tcgen05.st is .sync.aligned: no divergence around it): the one
                    // lane beyond the last interior column (X = n) re-reads column n-1 and its results are
                    // discarded.
                    const int Xc = colok ? X : n - 1;
                    double    a0[NRHS], a1[NRHS], a2[NRHS];
                    double    b0[NRHS], b1[NRHS], b2[NRHS];
                    ldv<NRHS>(sP, (Y0 - 1) * np + Xc - 1, a0);
                    ldv<NRHS>(sP, (Y0 - 1) * np + Xc, a1);
                    ldv<NRHS>(sP, (Y0 - 1) * np + Xc + 1, a2);
                    ldv<NRHS>(sP, Y0 * np + Xc - 1, b0);
                    ldv<NRHS>(sP, Y0 * np + Xc, b1);
                    ldv<NRHS>(sP, Y0 * np + Xc + 1, b2);
                    // couplings towards the row below the current one, carried up the strip
                    using K = Coef<n>;
                    // ROWPAIR: the strip starts on an odd row, so rows (Y0 - 1, Y0), (Y0 + 1, Y0 + 2), ... are pairs
                    [[maybe_unused]] double2 pE, pW, pD2, pNW, pN, pNE, pD1w;
                    [[maybe_unused]] const int m0 = (Y0 - 1) >> 1;
                    [[maybe_unused]] auto load_pair = [&](int m) {
                      pE = K::pair(sA, m, Xc), pW = K::pair(sA, m, Xc - 1);
                      pD2 = K::pair(sA + n * n, m, Xc), pNW = K::pair(sA + n * n, m, Xc - 1);
                      pN = K::pair(sB, m, Xc);
                      pNE = K::pair(sB + n * n, m, Xc), pD1w = K::pair(sB + n * n, m, Xc - 1);
                    };
                    double cS, cSE, cSW;
                    if constexpr (K::ROWPAIR)
                      {
                        load_pair(m0);
                        cS = pN.x, cSE = pD2.x, cSW = pD1w.x;
                      }
                    else
                      {
                        cS  = sB[K::e((Y0 - 1) * n + Xc)];      // N(X, y-1)
                        cSE = sA[K::hi((Y0 - 1) * n + Xc)];     // D2(X, y-1)
                        cSW = sB[K::hi((Y0 - 1) * n + Xc - 1)]; // D1(X-1, y-1)
                      }
#pragma unroll
                    for (int j = 0; j < RPT; ++j)
                      {
                        const int y = Y0 + j;
                        if (y <= n - 1) // (uniform over the warp)
                          {
                            double c0[NRHS], c1[NRHS], c2[NRHS];
                            ldv<NRHS>(sP, (y + 1) * np + Xc - 1, c0);
                            ldv<NRHS>(sP, (y + 1) * np + Xc, c1);
                            ldv<NRHS>(sP, (y + 1) * np + Xc + 1, c2);
                            double cE, d2o, cW, cNW, cN, cNE, d1w;
                            if constexpr (K::ROWPAIR)
                              {
                                if (j & 1) // row Y0 + j is even: the first row of the next pair
                                  load_pair(m0 + (j + 1) / 2);
                                constexpr bool second = true;
                                (void)second;
                                if (j & 1)
                                  cE = pE.x, cW = pW.x, d2o = pD2.x, cNW = pNW.x, cN = pN.x, cNE = pNE.x, d1w = pD1w.x;
                                else
                                  cE = pE.y, cW = pW.y, d2o = pD2.y, cNW = pNW.y, cN = pN.y, cNE = pNE.y, d1w = pD1w.y;
                              }
                            else if constexpr (K::PAIRED)
                              {
                                double unused;
                                ld2(sA, y * n + Xc, cE, d2o);
                                ld2(sA, y * n + Xc - 1, cW, cNW);
                                ld2(sB, y * n + Xc, cN, cNE);
                                ld2(sB, y * n + Xc - 1, unused, d1w);
                              }
                            else
                              {
                                const int i = y * n + Xc;
                                cE = sA[i], cW = sA[i - 1];
                                d2o = sA[n * n + i], cNW = sA[n * n + i - 1];
                                cN = sB[i];
                                cNE = sB[n * n + i], d1w = sB[n * n + i - 1];
                              }
#pragma unroll
                            for (int k = 0; k < NRHS; ++k)
                              {
                                double t = b1[k];
                                t        = fma(cE, b2[k], t);
                                t        = fma(cW, b0[k], t);
                                t        = fma(cN, c1[k], t);
                                t        = fma(cS, a1[k], t);
                                t        = fma(cNE, c2[k], t);
                                t        = fma(cSW, a0[k], t);
                                t        = fma(cNW, c0[k], t);
                                t        = fma(cSE, a2[k], t);
                                t        = colok ? t : 0.0;
                                q[j][k]  = t;
                                pq[k]    = fma(b1[k], t, pq[k]);
                                a0[k] = b0[k], a1[k] = b1[k], a2[k] = b2[k];
                                b0[k] = c0[k], b1[k] = c1[k], b2[k] = c2[k];
                              }
                            cS = cN, cSE = d2o, cSW = d1w;
                          }
                        if constexpr (QTMEM)
                          {
                            if ((j & 3) == 3) // rows 4c .. 4c+3 are complete (zero beyond the mesh)
                              {
                                double q8[8];
#pragma unroll
                                for (int jj = 0; jj < 4; ++jj)
                                  q8[2 * jj] = q[j - 3 + jj][0], q8[2 * jj + 1] = q[j - 3 + jj][1];
                                tmem::st8(tm + C::ROFF + 16 * (j >> 2), q8);
                              }
                          }
                      }
                    if constexpr (QTMEM)
                      tmem::wait_st();
                  }
                  ST_MARK(2)
                  block_sum2<NWARP>(pq[0], pq[1], sRed, warp, lane);
                  ST_MARK(3)
#pragma unroll
                  for (int k = 0; k < NRHS; ++k)
                    alpha[k] = it == 0 ? 1.0 : (done[k] ? 0.0 : fast_div(rho[k], pq[k]));
                  if constexpr (RMODE == 0)
                    {
#pragma unroll
                      for (int j = 0; j < RPT; ++j)
#pragma unroll
                        for (int k = 0; k < NRHS; ++k)
                          rreg[j][k] = fma(-alpha[k], q[j][k], rreg[j][k]);
                    }
                }

              // ---- rhat -= alpha q (tensor memory); u = D^1/2 rhat, pre-summed per strip into the vector buffer
              //      (p is dead there: every warp has passed the barrier of the p.q reduction); |u|^2 for the
              //      stopping rule, |rhat|^2 for r.z
              double rz[NRHS] = {0.0, 0.0}, rr[NRHS] = {0.0, 0.0};
              {
                double    acc[NRHS];
                const int pc = PS::col(X);
                double    sq8[8], ra[8], rb[8];
                if constexpr (RTMEM)
                  tmem::ld8x3(tm + C::SOFF, tm + C::ROFF, tm + C::ROFF + 16, sq8, ra, rb);
                else
                  {
                    if constexpr (QTMEM)
                      {
                        double qa[8], qb[8];
                        tmem::ld8x3(tm + C::SOFF, tm + C::ROFF, tm + C::ROFF + 16, sq8, qa, qb);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                          for (int k = 0; k < NRHS; ++k)
                            q[jj][k] = qa[2 * jj + k], q[4 + jj][k] = qb[2 * jj + k];
                      }
                    else
                      tmem::ld8(tm + C::SOFF, sq8);
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                      for (int k = 0; k < NRHS; ++k)
                        ra[2 * jj + k] = rreg[jj][k], rb[2 * jj + k] = rreg[4 + jj][k];
                  }
                if constexpr (RMODE != 0)
                  {
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                      for (int k = 0; k < NRHS; ++k)
                        {
                          ra[2 * jj + k] = fma(-alpha[k], q[jj][k], ra[2 * jj + k]);
                          rb[2 * jj + k] = fma(-alpha[k], q[4 + jj][k], rb[2 * jj + k]);
                        }
                    if constexpr (RTMEM)
                      {
                        tmem::st8(tm + C::ROFF, ra);
                        tmem::st8(tm + C::ROFF + 16, rb);
                      }
                    else
                      {
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
#pragma unroll
                          for (int k = 0; k < NRHS; ++k)
                            rreg[jj][k] = ra[2 * jj + k], rreg[4 + jj][k] = rb[2 * jj + k];
                      }
                  }
#pragma unroll
                for (int jj = 0; jj < 8; ++jj)
                  {
                    double u[NRHS];
#pragma unroll
                    for (int k = 0; k < NRHS; ++k)
                      {
                        const double rv = jj < 4 ? ra[2 * jj + k] : rb[2 * (jj - 4) + k];
                        u[k]  = sq8[jj] * rv; // zero beyond the mesh
                        rr[k] = fma(u[k], u[k], rr[k]);
                        rz[k] = fma(rv, rv, rz[k]);
                      }
                    PS::push(sP, jj, u, acc, pc, wy, colok);
                  }
                if constexpr (RTMEM)
                  tmem::wait_st(); // rhat is re-read after the coarse levels
              }
              __syncthreads();
              ST_MARK(4)
              // ---- coarse levels; r.z = |rhat|^2 + u_1 . z_1 is completed while level 1 is prolonged
                {
                  // The index arithmetic of the coarse stages is loop invariant; hoisted out of the PCG loop it
                  // stays live across the whole iteration and is SPILLED (the kernel sits at the 128-register
                  // cap), and with 227 KB of the L1 carved out as shared memory a spill load goes to L2: eight of
                  // them in front of one coarse stage cost ~3500 cycles per iteration.  Laundering the thread
                  // index through an empty asm keeps the arithmetic inside the loop.
                  int tl = tid;
#ifndef MSB_EMU
                  asm volatile("" : "+r"(tl));
#endif
                  coarse_correction<NL, NRHS, THREADS, RPT, true, C::PAD>(
                  sP, sV, sDi, tl, tl >> 5, tl & 31,
                  [&](int st_k) {
                    (void)st_k;
                    ST_MARK(st_k)
                  },
                  [&](int c, double(&g)[8]) {
                    if constexpr (C::GI_TMEM)
                      tmem::ld8(tmat + 16 * c, g);
                    else
                      {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                          g[i] = X7::fetch_tri(sGiv, tl, c, i);
                      }
                  },
                  rz);
                }
              ST_MARK(11)
              {
                double four[4] = {rz[0], rz[1], rr[0], rr[1]};
                block_sum4<NWARP>(four, sRed + 2 * NWARP, warp, lane); // (its barrier publishes z_1)
                rz[0] = four[0], rz[1] = four[1], rr[0] = four[2], rr[1] = four[3];
              }
              ST_MARK(8)

              double beta[NRHS];
              all_done = true;
#pragma unroll
              for (int k = 0; k < NRHS; ++k)
                {
                  beta[k] = (done[k] || it == 0) ? 0.0 : fast_div(rz[k], rho[k]);
                  if (!done[k])
                    {
                      rho[k]   = rz[k];
                      exact[k] = rr[k];
                      if (rr[k] <= P.tol2)
                        {
                          done[k] = true;
                          kit[k]  = it;
                        }
                      else if (it >= P.max_iter)
                        kit[k] = it;
                    }
                  all_done = all_done && done[k];
                }
              const bool last = all_done || it >= P.max_iter;

              // ---- zhat = rhat + D^1/2 (P z_1) (coarse-row averages cached in registers), fused with
              //      x += alpha p_old ; p = zhat + beta p_old (x and p_old live in tensor memory)
              {
                constexpr int np1 = L::lvl_np(1);
                const int     xl = X >> 1, xh = (X + 1) >> 1, cr0 = (Y0 - 1) >> 1;
                double        sq8[8];
                tmem::ld8(tm + C::SOFF, sq8);
                double h[5][NRHS];
#pragma unroll
                for (int m = 0; m <= 4; ++m)
                  {
                    const int cr = cr0 + m <= n / 2 ? cr0 + m : n / 2;
                    double    a[NRHS] = {0.0, 0.0}, b[NRHS] = {0.0, 0.0};
                    if (colok && !last)
                      {
                        ldv<NRHS>(sV, cr * np1 + xl, a);
                        ldv<NRHS>(sV, cr * np1 + xh, b);
                      }
#pragma unroll
                    for (int k = 0; k < NRHS; ++k)
                      h[m][k] = 0.5 * (a[k] + b[k]);
                  }
#pragma unroll
                for (int c = 0; c < NCH; ++c)
                  {
                    double x8[8], p8[8], r8[8];
                    if constexpr (RTMEM)
                      tmem::ld8x3(tm + C::XOFF + 16 * c, tm + C::POFF + 16 * c, tm + C::ROFF + 16 * c, x8, p8, r8);
                    else
                      {
                        tmem::ld8x2(tm + C::XOFF + 16 * c, tm + C::POFF + 16 * c, x8, p8);
#pragma unroll
                        for (int jj = 0; jj < 4; ++jj)
                          r8[2 * jj] = rreg[4 * c + jj][0], r8[2 * jj + 1] = rreg[4 * c + jj][1];
                      }
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj)
                      {
                        const int j = 4 * c + jj, y = Y0 + j;
                        double    pn[NRHS];
#pragma unroll
                        for (int k = 0; k < NRHS; ++k)
                          {
                            const double cc = (j & 1) ? h[(j + 1) / 2][k] : 0.5 * (h[j / 2][k] + h[j / 2 + 1][k]);
                            const double z  = fma(sq8[j], cc, r8[2 * jj + k]);
                            const double po = p8[2 * jj + k];
                            x8[2 * jj + k]  = fma(alpha[k], po, x8[2 * jj + k]);
                            pn[k]           = (done[k] && beta[k] == 0.0) ? po : fma(beta[k], po, z);
                            p8[2 * jj + k]  = pn[k];
                          }
                        if (colok && y <= n - 1 && !last)
                          stv<NRHS>(sP, y * np + X, pn);
                      }
                    tmem::st8(tm + C::XOFF + 16 * c, x8);
                    if (!last)
                      tmem::st8(tm + C::POFF + 16 * c, p8);
                  }
              }
              if (last)
                {
                  tmem::wait_st();
                  __syncthreads(); // the reduction buffer is reused by the epilogue
                  break;
                }
              zero_halo();
              tmem::wait_st();
              __syncthreads();
              ST_MARK(9)
              ++it;
            }

          // -------------------------------------------------------------- epilogue
          // distribute() (basis.tpp:308): interior phi = D^-1/2 xhat, boundary phi = g; and the columns
          // rhs0, rhs0+1 of M plus b[rhs0], b[rhs0+1] (basis.tpp:245-285) from the boundary rows of K phi
          double *out = P.phi + ((size_t)cell * 4 + rhs0) * N;
          double  macc[4][NRHS], bsum[NRHS] = {0.0, 0.0};
#pragma unroll
          for (int i = 0; i < 4; ++i)
            macc[i][0] = macc[i][1] = 0.0;
          fill_boundary_table(); // (the coarse vectors are dead: the loop ended with a block barrier)
          __syncthreads();
          {
            double sq8[8], xa[8], xb[8];
            tmem::ld8(tm + C::SOFF, sq8);
            tmem::ld8x2(tm + C::XOFF, tm + C::XOFF + 16, xa, xb); // rows 0..3, 4..7
#pragma unroll
            for (int jj = 0; jj < 8; ++jj)
              {
                const int y = Y0 + jj;
                if (colok && y <= n - 1)
                  {
                    const int    i = y * np + X;
                    const double s = fast_div(1.0, sq8[jj]);
                    double       xh[NRHS];
#pragma unroll
                    for (int k = 0; k < NRHS; ++k)
                      {
                        xh[k]                  = jj < 4 ? xa[2 * jj + k] : xb[2 * (jj - 4) + k];
                        const double phi       = s * xh[k];
                        out[(size_t)k * N + i] = phi;
                        bsum[k] += phi;
                      }
                    if (X == 1 || X == n - 1 || y == 1 || y == n - 1)
                      {
#pragma unroll
                        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                          for (int dx = -1; dx <= 1; ++dx)
                            {
                              const int bx = X + dx, by = y + dy;
                              if ((dx == 0 && dy == 0) || !(bx == 0 || bx == n || by == 0 || by == n))
                                continue;
                              const double kij = eget<n>(sA, sB, X, y, dx, dy); // Ahat_ac = K_ac d_c^-1/2
                              double       g[4];
                              const int    bi = boundary_index(bx, by);
                              ld2(sG, 2 * bi, g[0], g[1]);
                              ld2(sG, 2 * bi + 1, g[2], g[3]);
                              const double gj[NRHS] = {grp ? g[2] : g[0], grp ? g[3] : g[1]};
#pragma unroll
                              for (int k = 0; k < NRHS; ++k)
                                {
                                  const double t = kij * fma(-sq8[jj], gj[k], xh[k]);
#pragma unroll
                                  for (int i2 = 0; i2 < 4; ++i2)
                                    macc[i2][k] = fma(g[i2], t, macc[i2][k]);
                                }
                            }
                      }
                  }
              }
          }
          for (int t = tid; t < 4 * n; t += THREADS)
            {
              int jx, jy;
              boundary_node(t, jx, jy);
              double ga[4];
              ld2(sG, 2 * t, ga[0], ga[1]);
              ld2(sG, 2 * t + 1, ga[2], ga[3]);
              const double gaj[NRHS] = {grp ? ga[2] : ga[0], grp ? ga[3] : ga[1]};
              double       T[NRHS]   = {0.0, 0.0};
              // the boundary neighbours of a boundary node (every pair of 8-neighbours shares a fine cell, so
              // every such pair is coupled); boundary-boundary edges are stored unscaled
#pragma unroll
              for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx)
                  {
                    const int bx = jx + dx, by = jy + dy;
                    if ((dx == 0 && dy == 0) || bx < 0 || bx > n || by < 0 || by > n ||
                        !(bx == 0 || bx == n || by == 0 || by == n))
                      continue;
                    const double kab = eget_full<n>(sA, sB, sEb, sNb, jx, jy, dx, dy);
                    double       gb0, gb1;
                    ld2(sG, 2 * boundary_index(bx, by) + grp, gb0, gb1);
                    T[0] = fma(kab, gb0 - gaj[0], T[0]);
                    T[1] = fma(kab, gb1 - gaj[1], T[1]);
                  }
              const bool   corner = (jx == 0 || jx == n) && (jy == 0 || jy == n);
              const double w      = corner ? 0.25 : 0.5;
#pragma unroll
              for (int k = 0; k < NRHS; ++k)
                {
                  out[(size_t)k * N + jy * np + jx] = gaj[k];
                  bsum[k] = fma(w, gaj[k], bsum[k]);
#pragma unroll
                  for (int i2 = 0; i2 < 4; ++i2)
                    macc[i2][k] = fma(ga[i2], T[k], macc[i2][k]);
                }
            }
          {
            double ten[10], e, t;
#pragma unroll
            for (int i2 = 0; i2 < 4; ++i2)
              ten[2 * i2] = macc[i2][0], ten[2 * i2 + 1] = macc[i2][1];
            ten[8] = bsum[0], ten[9] = bsum[1];
            warp_sum10(ten, lane, e, t);
            if ((lane & 3) == 0)
              sRed[((lane >> 2) & 7) * NWARP + warp] = e;
            if ((lane & 15) == 0)
              sRed[(8 + (lane >> 4)) * NWARP + warp] = t;
            __syncthreads();
            // ten threads add the warp partials in a fixed order (deterministic) and write the results
            if (tid < 10)
              {
                double sum = 0.0;
#pragma unroll
                for (int w = 0; w < NWARP; ++w)
                  sum += sRed[tid * NWARP + w];
                if (tid < 8)
                  P.M[16 * (size_t)cell + 4 * (tid >> 1) + rhs0 + (tid & 1)] = sum;
                else
                  P.b[4 * (size_t)cell + rhs0 + (tid - 8)] = P.rhs_value * hx * hy * sum;
              }
            if (tid == 0)
              {
#pragma unroll
                for (int k = 0; k < NRHS; ++k)
                  {
                    const int sidx = cell * 4 + rhs0 + k;
                    P.iters[sidx]  = kit[k];
                    P.res[sidx]    = sqrt(exact[k]);
                    if (!done[k])
                      atomicMin(P.fail, P.fail_base + sidx);
                  }
              }
          }
          __syncthreads(); // shared buffers are reused by the next pair of bases
          ST_MARK(10)
        }
      ST_FLUSH
      tmem::release(tmem_base, warp, C::TMEM_COLS);
    }
  } // namespace fused

#ifndef MSB_EMU
  namespace fused
  {
    template <int NL>
    static cudaError_t
    launch(const FusedParams &P, cudaStream_t st)
    {
      using C = Cfg<NL>;
      void (*kern)(FusedParams);
      switch (P.flavor)
        {
          case 1:
            kern = solve_fused_kernel<NL, 1>;
            break;
          case 2:
            kern = solve_fused_kernel<NL, 2>;
            break;
          default:
            kern = solve_fused_kernel<NL, 0>;
        }
      if (P.split) // (split launches always run flavour 0)
        kern = solve_fused_kernel<NL, 0, true>;
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::smem_bytes);
      if (e != cudaSuccess)
        return e;
      kern<<<P.split ? 2 * P.n_cells : P.n_cells, C::THREADS, C::smem_bytes, st>>>(P);
      return cudaGetLastError();
    }
  } // namespace fused

  cudaError_t
  launch_solve_fused(const FusedParams &P, int l, cudaStream_t st)
  {
    return l == 5 ? fused::launch<5>(P, st) : fused::launch<6>(P, st);
  }
#endif
} // namespace msb
