// megabeast_b200 device code: the ensemble-likelihood kernels for sm_100a.
//
// Reference semantics (what every kernel must reproduce, in fp64):
//   /root/reference/megabeast/singlepop_dust_model.py:131-148  physmod = prod of active priors
//   /root/reference/megabeast/helpers.py:133-164               expected number of stars
//   /root/reference/megabeast/singlepop_dust_model.py:182-186  Poisson term
//   /root/reference/megabeast/singlepop_dust_model.py:194-223  per-star gather, sum, log, sum
//
// Design (DESIGN.md has the long form).  BEAST grids are Cartesian: every free parameter's
// column holds a few tens of distinct values ("classes"), and physmod[g,w] depends on the row
// only through its class codes:
//
//     physmod[g,w] = FO[cO(g)][w] * FI[cI(g)][w]
//
// FO = "outer" factor table (age class x the dust parameters kept outside), FI = "inner" factor
// table (the remaining dust parameters).  The classic split is outer = age alone (FA), inner = all
// dust parameters (FB); grids whose dust classes do not fit shared memory as ONE product table
// (production BEAST grids: 100-200 A(V) values x 9 R(V)) keep the parameters separate: inner = A(V),
// outer = age x R(V) (x f_A), and a star's entries are sorted into runs of equal outer class.
//
// K0  mb_k0_factors      one block per walker: prior models on the classes -> factor tables FO, FI
//                        (TABLE modes, tabulated models, pixel batches; in FACTOR mode K2 builds
//                        the tables itself in its prologue and K0 is not launched)
// K1  mb_k1_expand       TABLE modes: physmod table T[u][w] = FA*FB          (HBM write bound)
// K1e mb_k1_elementwise  ELEMENTWISE mode: physmod[g][w] from the grid's physics columns, row by
//                        row (any grid, no class coding); K1b mb_k1b_binsums adds its per-tile
//                        N-stars partial sums per (age, Z) bin
// K2  mb_k2_stars<..>    prologue: (FACTOR) the factor tables, built in shared memory while the first
//                        entry chunks are already on their way; every block takes some (age, Z) bins of
//                        the N-stars statistics x FI x FO -> per-bin sums (one warp per bin when the
//                        tables are small).  Then the hot loop:
//                        stream (weight, code) entries once for all walkers, look the factors up
//                        (smem) or gather T rows (L2/HBM), per-star fp64 sum, log via
//                        mantissa/exponent product.  K3 is fused into its tail: warp -> block ->
//                        grid (-> GPUs over NVLink peer memory) reduction in two-word fixed point;
//                        the last block also adds the bins -> expected N -> Poisson term
// KPX mb_k2_pixels<..>   pixel batches: one block per pixel at a time, same inner loop
// KB  mb_build_entries   once per mb_create: (star, sparse point) pairs -> the sorted, merged, flagged
//                        entry stream (one warp per star, bitonic sort); mb_compact_entries packs it
//
// Lanes of a warp are WALKERS, a warp walks one star's entries at a time, so every control
// decision in K2 (run boundary, end of star) is warp-uniform.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#include "../../include/megabeast_b200.h"

namespace mb {

// ---- entry stream --------------------------------------------------------------------------
// One finite (star, sparse point) pair = an 8-byte weight and a 4-byte code in two parallel arrays:
//   ea[i] = vals * prior_weight[r] * grid_weight[r] * completeness[r]   (walker independent,
//           folded at create time)
//   ec[i] = FACTOR:  [31] first entry of a run of equal age class (or of a star),
//                    [30] first entry of a star, [29:20] age class, [19:0] dust class
//           TABLE*:  [31] first entry of a star, [30:0] table row: class combination cA*nB+cB
//                    (TABLE_UNIQUE) or grid row (TABLE_GRID, the literal gather)
constexpr uint32_t kNewRun = 0x80000000u;
constexpr uint32_t kNewStar = 0x40000000u;
constexpr uint32_t kRowMask = 0x7fffffffu;
// grid-row class code (host side and K1 expand): cA << kAgeShift | cB
constexpr int kAgeShift = 20;
constexpr uint32_t kAgeMask = 0x3ffu;       // <= 1024 age classes
constexpr uint32_t kDustMask = 0xfffffu;    // <= 2^20 dust classes

struct ModelDev {        // device copy of what K0 needs
  int32_t kind[MB_NPARAM];
  int32_t nnodes[MB_NPARAM];
  int32_t voff[MB_NPARAM];    // offset of the parameter's variables in phi
  int32_t ncls[MB_NPARAM];
  int32_t clsoff[MB_NPARAM];  // offset into clsx
  int32_t radix[MB_NPARAM];   // dust class code = sum_p digit_p * radix[p]  (p = Av, Rv, fA)
  double nodes[MB_NPARAM][MB_MAX_NODES];
};

// ---- prior models (restating beast.physicsmodel.priormodel; anchors in oracle/priors.py) --------
__device__ __forceinline__ double lognorm(double x, double mode, double sigma, double amp) {
  // megabeast/old/ensemble_model.py:13-47: mu = ln(mode) + sigma^2, 0 for x <= 0
  if (!(x > 0.0)) return 0.0;
  const double inv_sqrt_2pi = 0.3989422804014326779399460599343819;
  double mu = log(mode) + sigma * sigma;
  double norm = inv_sqrt_2pi / (x * sigma);
  double t = (log(x) - mu) / sigma;
  return amp * norm * exp(-0.5 * (t * t));
}

__device__ double prior_eval(const ModelDev& m, int p, const double* __restrict__ th, double x) {
  switch (m.kind[p]) {
    case MB_LOGNORMAL:
      return lognorm(x, th[0], th[1], 1.0);
    case MB_TWO_LOGNORMAL: {
      double n1 = 1.0 - 1.0 / (th[4] + 1.0), n2 = 1.0 / (th[4] + 1.0);
      return lognorm(x, th[0], th[2], n1) + lognorm(x, th[1], th[3], n2);
    }
    case MB_BINS_HISTO: {  // left-constant steps; x == last node -> last height
      int n = m.nnodes[p], i = 0;
      while (i + 1 < n && x >= m.nodes[p][i + 1]) ++i;
      return th[i];
    }
    case MB_BINS_INTERP: {  // np.interp
      int n = m.nnodes[p];
      if (x <= m.nodes[p][0]) return th[0];
      if (x >= m.nodes[p][n - 1]) return th[n - 1];
      int i = 0;
      while (x >= m.nodes[p][i + 1]) ++i;
      double slope = (th[i + 1] - th[i]) / (m.nodes[p][i + 1] - m.nodes[p][i]);
      return slope * (x - m.nodes[p][i]) + th[i];
    }
    case MB_EXPONENTIAL:
      return exp(-1.0 * pow(10.0, x) / (th[0] * 1e9));
    default:  // MB_FLAT
      return 1.0;
  }
}

// deterministic block-wide sum (shuffle tree inside a warp, fixed order across warps)
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* scratch) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();  // scratch may still be read from the previous call
  if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
#pragma unroll
  for (int i = 0; i < THREADS / 32; ++i) t += scratch[i];
  return t;
}

// ---- K0: per-walker factor tables ----------------------------------------------------------------
// One block per walker column.  Phase 1: every free parameter's prior model evaluated on its
// classes (singlepop_dust_model.py:146-148 restricted to the distinct column values) into a
// shared-memory look-up table -- a few tens of fp64 exp/log evaluations per walker.  Phase 2:
// the table rows (nrows = outer rows then inner rows), each the product of up to four look-up rows
// listed in `rowsrc` (the reference's left-to-right order), then the nAge plain age-factor rows
// (the `ageprior` of helpers.py:135), walker minor.  Columns w >= W (padding up to the pass width)
// repeat the last real walker so no lane of K2 ever sees garbage.  The model description travels
// in the kernel parameters (constant bank).  Pixel batches: P independent ensembles, tables laid
// out [p][class][Wpad], phi [p][W][ndim].
// tab[p]: host-evaluated table [columns][ncls[p]] for MB_TABULATED parameters (else unused).
struct K0Args {
  ModelDev model;
  const double* clsx;
  const double* phi;          // [P][W][ndim], already offset to the pass's first walker
  const double* tab[MB_NPARAM];
  double* F;                  // [P][nrows + nAge][Wpad]: outer rows, inner rows, plain age rows
  const uint16_t* rowsrc;     // [nrows][4] look-up rows (class offsets included) multiplied into a table row; 0xffff = none
  int32_t nrows, nAge, W, Wpad, ndim, ncls_total;
  const double* box;          // null, or [3][ndim] = lo, hi, park: a column with any phi outside [lo, hi] is built
                              // from `park` instead (mb_lnprob_pixels_box; the host reports it as -inf)
};
constexpr int kK0Threads = 128;

__global__ void __launch_bounds__(kK0Threads) mb_k0_factors(const __grid_constant__ K0Args a) {
  extern __shared__ double lut[];  // [sum of ncls over the parameters], offsets = model.clsoff
  // programmatic dependent launch: K2's blocks may be scheduled while this grid still runs; they
  // wait (griddepcontrol.wait) for this grid's completion before touching the tables
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const ModelDev& m = a.model;
  const int pix = blockIdx.x / a.Wpad, w = blockIdx.x - pix * a.Wpad;
  const int wr = w < a.W ? w : a.W - 1;
  const size_t col = (size_t)pix * a.W + wr;
  const double* ph = a.phi + col * a.ndim;
  if (a.box) {  // block-uniform: every thread tests the same ndim values
    bool outside = false;
    for (int k = 0; k < a.ndim; ++k) {
      const double v = ph[k];
      outside |= v < a.box[k] || v > a.box[a.ndim + k];  // a NaN is inside, as in singlepop_dust_model.py:255-271
    }
    if (outside) ph = a.box + 2 * a.ndim;
  }
  for (int i = threadIdx.x; i < a.ncls_total; i += kK0Threads) {
    int p = 0;
#pragma unroll
    for (int q = 1; q < MB_NPARAM; ++q)
      if (i >= m.clsoff[q]) p = q;
    const int c = i - m.clsoff[p];
    double v = 1.0;
    if (m.kind[p] == MB_TABULATED) v = a.tab[p][col * m.ncls[p] + c];
    else if (m.kind[p] != MB_FIXED) v = prior_eval(m, p, ph + m.voff[p], a.clsx[i]);
    lut[i] = v;
  }
  __syncthreads();
  double* F = a.F + (size_t)pix * (a.nrows + a.nAge) * a.Wpad + w;
  for (int c = threadIdx.x; c < a.nrows + a.nAge; c += kK0Threads) {
    double v = 1.0;
    if (c < a.nrows) {
      // same left-to-right product order as the reference's `cur_physmod *=` loop (:146-148)
      const ushort4 r = __ldg(reinterpret_cast<const ushort4*>(a.rowsrc) + c);
      if (r.x != 0xffff) v *= lut[r.x];
      if (r.y != 0xffff) v *= lut[r.y];
      if (r.z != 0xffff) v *= lut[r.z];
      if (r.w != 0xffff) v *= lut[r.w];
    } else if (m.kind[MB_P_LOGA] != MB_FIXED) {
      v = lut[m.clsoff[MB_P_LOGA] + (c - a.nrows)];
    }
    F[(size_t)c * a.Wpad] = v;
  }
}

// ---- N-stars / Poisson term, spread over the blocks of K2 by (age, Z) bin -------------------------
// helpers.py:133-162 needs, per (age, Z) bin, S = sum_g physmod*grid_weight and
// C = sum_g physmod*grid_weight*completeness over the bin's rows.  Inside a bin the age factor is
// constant, so with the create-time statistics Ht[bin][cD] = (sum grid_weight, sum
// grid_weight*completeness) over the bin's rows of dust class cD = od * nI + i (outer-dust digit,
// inner class):
//     S_bin = sum_od FO[age(bin) * nOd + od] * sum_i Ht[bin][od * nI + i].x * FI[i]     (C_bin with .y)
// Block b of K2 computes the bins b, b + gridDim.x, ... for ALL walkers of the pass (Ht is read
// once per launch, not once per walker) and leaves (S_bin, term_bin) in `Pb`; the last block of the
// launch adds the bins in fixed order: normalisation (:133-134; it cancels in C/S and only decides
// the degenerate cases), skip of empty bins (:156), expected number of stars, Poisson term
// (singlepop_dust_model.py:182-186).
struct PoissonArgs {
  const double* Ht;           // [nbins][nOd * nI][2]
  const int32_t* bin_agecls;  // [nbins] row of `fage` holding the bin's plain age factor; also the age class
  const double* coef;         // [nbins] metprior * massmult / avemass   (helpers.py:148-151)
  double* Pb;                 // [nbins][2][Wpad] scratch: S_bin, term_bin (rewritten by every launch)
  const double* fage;         // plain age factors [..][Wpad] in global memory, or null: read them from the
                              // tables / look-up rows in shared memory (fage_smem_off)
  double n_stars;             // S of the fit (:183-185)
  double lgamma_n1;           // lgamma(S + 1), a constant of the fit
  int32_t nbins;              // 0: no N-stars term
  int32_t nwarps;             // warps of the block that take part (their partial sums must fit the scratch)
  int32_t scratch_off;        // byte offset of the reduction scratch [kPoisBins][nwarps][2][Wpad] in dynamic shared memory
  int32_t fage_smem_off;      // byte offset of the plain age-factor rows in dynamic shared memory (fage == null)
  const double* SC;           // ELEMENTWISE: bin sums [nbins][Wpad][2] already formed by K1b (else null)
  // few dust classes: ONE WARP per bin (lanes = walkers, no cross-warp sums, no block barriers); the warp's
  // first bin's statistics are staged into shared memory (cp.async at kernel start)
  int32_t warp_private;       // 1: this scheme
  int32_t hts_off;            // byte offset of the staged statistics [hts_slices][nOd * nI][2]
  int32_t hts_slices;         // warps that have a slice
  int32_t bins_first;         // 1 (default): the bins are claimed right after the table build by the first warps
                              // to get there; 0 (MB_BINS_LAST=1): after the star loop by the first warps to finish
};
constexpr int kPoisBins = 2;  // bins a block works on at a time (loads of both in flight together)
// classes of a bin whose loads are in flight together in the one-warp-per-bin sum: the warp shares the load/store
// pipe with 23 warps streaming stars, every dependent round trip costs ~200 cycles there
#ifndef MB_BIN_UNROLL
#define MB_BIN_UNROLL 4
#endif
constexpr int kBinUnroll = MB_BIN_UNROLL;

// F: outer rows then inner rows, element (row, w) at F[row * Wpad + w]; shared memory (FACTOR mode with
// fp64 tables) or global.  All threads of the block must call.  Lane l of a warp owns walkers
// l, l + 32, ... (8-byte loads, conflict free).
template <int NW, int THREADS>
__device__ void poisson_bins(const PoissonArgs& pa, const double* F, int nO, int nOd, int nI,
                             unsigned char* smem_raw, bool scratch_is_private) {
  constexpr int Wpad = 32 * NW;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pw = pa.nwarps;
  double* red = reinterpret_cast<double*>(smem_raw + pa.scratch_off);  // [kPoisBins][pw][2][Wpad]
  const double* fage = pa.fage ? pa.fage : reinterpret_cast<const double*>(smem_raw + pa.fage_smem_off);
  const double2* Ht = reinterpret_cast<const double2*>(pa.Ht);
  const double* FI = F + (size_t)nO * Wpad + lane;
  const double2* hts = reinterpret_cast<const double2*>(smem_raw + pa.hts_off);
  int pass = -1;
  for (int b0 = blockIdx.x; b0 < pa.nbins; b0 += kPoisBins * gridDim.x) {
    if (++pass > 0) __syncthreads();  // the scratch is reused by the block's next pass
    // this pass: bins b0 and b0 + gridDim.x (the second may not exist)
    const int nb = (b0 + (int)gridDim.x < pa.nbins) ? 2 : 1;
    // the combining threads fetch their bin's constants now: the latency hides behind the sums
    const int cw = threadIdx.x % Wpad, cb = threadIdx.x / Wpad;
    int c_acls = 0;
    double c_coef = 0.0;
    if (cb < nb) {
      c_acls = __ldg(pa.bin_agecls + b0 + cb * (int)gridDim.x);
      c_coef = __ldg(pa.coef + b0 + cb * (int)gridDim.x);
    }
    if (warp < pw) {
      double S[kPoisBins][NW], C[kPoisBins][NW];
#pragma unroll
      for (int k = 0; k < kPoisBins; ++k)
#pragma unroll
        for (int j = 0; j < NW; ++j) { S[k][j] = 0.0; C[k][j] = 0.0; }
      if (pa.SC) {
        // ELEMENTWISE: S and C were summed over the grid rows themselves (mb_k1b_binsums)
        if (warp == 0) {
#pragma unroll
          for (int k = 0; k < kPoisBins; ++k)
            if (k < nb) {
#pragma unroll
              for (int j = 0; j < NW; ++j) {
                const double2 v = __ldg(reinterpret_cast<const double2*>(pa.SC) +
                                        (size_t)(b0 + k * (int)gridDim.x) * Wpad + lane + 32 * j);
                S[k][j] = v.x; C[k][j] = v.y;
              }
            }
        }
      } else {
        int acls[kPoisBins];
#pragma unroll
        for (int k = 0; k < kPoisBins; ++k) acls[k] = k < nb ? __ldg(pa.bin_agecls + b0 + k * (int)gridDim.x) : 0;
        for (int od = 0; od < nOd; ++od) {
          double s[kPoisBins][NW], c[kPoisBins][NW];
#pragma unroll
          for (int k = 0; k < kPoisBins; ++k)
#pragma unroll
            for (int j = 0; j < NW; ++j) { s[k][j] = 0.0; c[k][j] = 0.0; }
          // (slices 2 pass, 2 pass + 1 of the statistics staged in shared memory at kernel start, if they fit)
          const double2* hrow0 = (2 * pass < pa.hts_slices ? hts + (size_t)(2 * pass) * nOd * nI
                                                           : Ht + (size_t)b0 * nOd * nI) + (size_t)od * nI;
          const double2* hrow1 = nb > 1 ? (2 * pass + 1 < pa.hts_slices ? hts + (size_t)(2 * pass + 1) * nOd * nI
                                                                        : Ht + (size_t)(b0 + (int)gridDim.x) * nOd * nI) + (size_t)od * nI
                                        : hrow0;
          int i = warp;
          // two classes x two bins: four independent (warp-uniform) loads of the statistics in flight
          for (; i + pw < nI; i += 2 * pw) {
            const double2 h00 = hrow0[i], h01 = hrow0[i + pw];
            const double2 h10 = hrow1[i], h11 = hrow1[i + pw];
            const double* f0 = FI + (size_t)i * Wpad;
            const double* f1 = FI + (size_t)(i + pw) * Wpad;
#pragma unroll
            for (int j = 0; j < NW; ++j) {
              const double fv0 = f0[32 * j], fv1 = f1[32 * j];
              s[0][j] = fma(h00.x, fv0, s[0][j]); c[0][j] = fma(h00.y, fv0, c[0][j]);
              s[1][j] = fma(h10.x, fv0, s[1][j]); c[1][j] = fma(h10.y, fv0, c[1][j]);
              s[0][j] = fma(h01.x, fv1, s[0][j]); c[0][j] = fma(h01.y, fv1, c[0][j]);
              s[1][j] = fma(h11.x, fv1, s[1][j]); c[1][j] = fma(h11.y, fv1, c[1][j]);
            }
          }
          for (; i < nI; i += pw) {
            const double2 h0 = hrow0[i], h1 = hrow1[i];
            const double* f = FI + (size_t)i * Wpad;
#pragma unroll
            for (int j = 0; j < NW; ++j) {
              const double fv = f[32 * j];
              s[0][j] = fma(h0.x, fv, s[0][j]); c[0][j] = fma(h0.y, fv, c[0][j]);
              s[1][j] = fma(h1.x, fv, s[1][j]); c[1][j] = fma(h1.y, fv, c[1][j]);
            }
          }
#pragma unroll
          for (int k = 0; k < kPoisBins; ++k) {
            const double* fo = F + (size_t)(acls[k] * nOd + od) * Wpad + lane;
#pragma unroll
            for (int j = 0; j < NW; ++j) {
              const double fv = fo[32 * j];
              S[k][j] = fma(fv, s[k][j], S[k][j]); C[k][j] = fma(fv, c[k][j], C[k][j]);
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < kPoisBins; ++k)
#pragma unroll
        for (int j = 0; j < NW; ++j) {
          red[(((size_t)k * pw + warp) * 2 + 0) * Wpad + lane + 32 * j] = S[k][j];
          red[(((size_t)k * pw + warp) * 2 + 1) * Wpad + lane + 32 * j] = C[k][j];
        }
    }
    __syncthreads();
    if (cb < nb) {   // thread (bin cb, walker cw)
      double Sb = 0.0, Cb = 0.0;
      for (int wp = 0; wp < pw; ++wp) {  // fixed warp order
        Sb += red[(((size_t)cb * pw + wp) * 2 + 0) * Wpad + cw];
        Cb += red[(((size_t)cb * pw + wp) * 2 + 1) * Wpad + cw];
      }
      // helpers.py:152-162: ageprior * metprior * massmult / avemass * C / S, bins with S <= 0 skipped
      const double fa = fage[(size_t)c_acls * Wpad + cw];
      const double term = Sb > 0.0 ? (fa * c_coef) * (Cb / Sb) : 0.0;
      const int b = b0 + cb * (int)gridDim.x;
      pa.Pb[((size_t)b * 2 + 0) * Wpad + cw] = Sb;
      pa.Pb[((size_t)b * 2 + 1) * Wpad + cw] = term;
    }
  }
  // the entry stage shares the scratch: the combining threads must be done before any warp streams.
  // With a private scratch the other warps go straight on to their stars.
  if (!scratch_is_private) __syncthreads();
}

// One warp per bin: the block's bins b = blockIdx.x + k * gridDim.x are claimed one at a time (shared
// counter) by warps that have finished their stars.  Lanes are walkers, so a lane ends with its
// walkers' complete (S, C): no scratch, no barrier.  For tables of a few hundred rows (a bin is ~nI
// broadcast loads + row reads).
template <int NW, int THREADS>
__device__ void poisson_bins_private(const PoissonArgs& pa, const double* F, int nO, int nOd, int nI,
                                     unsigned char* smem_raw, int* bin_ctr) {
  constexpr int Wpad = 32 * NW;
  const int lane = threadIdx.x & 31;
  const double* fage = pa.fage ? pa.fage : reinterpret_cast<const double*>(smem_raw + pa.fage_smem_off);
  const double2* Ht = reinterpret_cast<const double2*>(pa.Ht);
  const double* FI = F + (size_t)nO * Wpad + lane;
  const int nD = nOd * nI;
  while (true) {
    int k = 0;
    if (lane == 0) k = atomicAdd(bin_ctr, 1);
    k = __shfl_sync(0xffffffffu, k, 0);
    const int b = (int)blockIdx.x + k * (int)gridDim.x;
    if (b >= pa.nbins) break;
    const double2* hts = reinterpret_cast<const double2*>(smem_raw + pa.hts_off) + (size_t)k * nD;
    const int acls = __ldg(pa.bin_agecls + b);
    const double coef = __ldg(pa.coef + b);
    double S[NW], C[NW];
#pragma unroll
    for (int j = 0; j < NW; ++j) { S[j] = 0.0; C[j] = 0.0; }
    if (pa.SC) {
#pragma unroll
      for (int j = 0; j < NW; ++j) {
        const double2 v = __ldg(reinterpret_cast<const double2*>(pa.SC) + (size_t)b * Wpad + lane + 32 * j);
        S[j] = v.x; C[j] = v.y;
      }
    } else {
      // (slice k was staged at kernel start by warp k and made visible by the table build's barrier)
      const double2* hrow = k < pa.hts_slices ? hts : Ht + (size_t)b * nD;
      for (int od = 0; od < nOd; ++od) {
        double s[NW], c[NW];
#pragma unroll
        for (int j = 0; j < NW; ++j) { s[j] = 0.0; c[j] = 0.0; }
#pragma unroll kBinUnroll
        for (int i = 0; i < nI; ++i) {
          const double2 h = hrow[(size_t)od * nI + i];   // warp-uniform: a broadcast
          const double* f = FI + (size_t)i * Wpad;
#pragma unroll
          for (int j = 0; j < NW; ++j) {
            const double fv = f[32 * j];
            s[j] = fma(h.x, fv, s[j]); c[j] = fma(h.y, fv, c[j]);
          }
        }
        const double* fo = F + (size_t)(acls * nOd + od) * Wpad + lane;
#pragma unroll
        for (int j = 0; j < NW; ++j) {
          const double fv = fo[32 * j];
          S[j] = fma(fv, s[j], S[j]); C[j] = fma(fv, c[j], C[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NW; ++j) {
      const int w = lane + 32 * j;
      // helpers.py:152-162: ageprior * metprior * massmult / avemass * C / S, bins with S <= 0 skipped
      const double fa = fage[(size_t)acls * Wpad + w];
      const double term = S[j] > 0.0 ? (fa * coef) * (C[j] / S[j]) : 0.0;
      pa.Pb[((size_t)b * 2 + 0) * Wpad + w] = S[j];
      pa.Pb[((size_t)b * 2 + 1) * Wpad + w] = term;
    }
  }
}

// Last block of the launch: add the bins in fixed order -> Poisson term of every walker of the pass.
// Two steps so that the loads of the first are in flight together with the caller's other loads.
struct PoisPart { double nrm, pred; };
template <int NW, int THREADS>
__device__ __forceinline__ PoisPart poisson_finish_load(const PoissonArgs& pa) {
  constexpr int Wpad = 32 * NW, kParts = THREADS / Wpad;
  const int w = threadIdx.x % Wpad, part = threadIdx.x / Wpad;
  PoisPart p{0.0, 0.0};
#pragma unroll 4
  for (int b = part; b < pa.nbins; b += kParts) {
    p.nrm += __ldcg(pa.Pb + ((size_t)b * 2 + 0) * Wpad + w);
    p.pred += __ldcg(pa.Pb + ((size_t)b * 2 + 1) * Wpad + w);
  }
  return p;
}
// psum: scratch [THREADS / Wpad][2][Wpad] doubles.  All threads of the block must call.
template <int NW, int THREADS>
__device__ void poisson_finish(const PoissonArgs& pa, PoisPart p, double* psum, double* out_row, int W) {
  constexpr int Wpad = 32 * NW, kParts = THREADS / Wpad;
  const int w = threadIdx.x % Wpad, part = threadIdx.x / Wpad;
  psum[((size_t)part * 2 + 0) * Wpad + w] = p.nrm;
  psum[((size_t)part * 2 + 1) * Wpad + w] = p.pred;
  __syncthreads();
  if (threadIdx.x < W) {
    double nrm = 0.0, pred = 0.0;
    for (int q = 0; q < kParts; ++q) {  // fixed order
      nrm += psum[((size_t)q * 2 + 0) * Wpad + threadIdx.x];
      pred += psum[((size_t)q * 2 + 1) * Wpad + threadIdx.x];
    }
    // helpers.py:133-134 divides every weight by the grand total first: a total of 0 (0/0 = nan
    // weights), nan or inf leaves no bin with S/total > 0 -> every bin skipped, pred = 0
    if (!(nrm > 0.0) || !(nrm < INFINITY)) pred = 0.0;
    double pois = 0.0;
    if (pa.nbins > 0) pois = pa.n_stars * log(pred) - pred - pa.lgamma_n1;  // :182-186
    out_row[threadIdx.x] = pois;
  }
  __syncthreads();
}

// ---- K1: materialise the prior-weight table (TABLE modes) -------------------------------------
// T[u][w] = FA[cA(u)][w] * FB[cB(u)][w].  rowcode == nullptr: u enumerates the distinct class
// combinations (u = cA*nB + cB); else u is a grid row and rowcode[u] = cA<<20 | cB.
// Reads 4 B/row (+ tables from L2), writes 8*Wpad B/row: HBM-write bound.
__global__ void __launch_bounds__(256) mb_k1_expand(const double* __restrict__ FA,
                                                    const double* __restrict__ FB,
                                                    const uint32_t* __restrict__ rowcode,
                                                    double* __restrict__ T, int64_t U, int nB,
                                                    int Wpad) {
  // one thread = one (row, walker pair): 16-byte stores, walker minor -> fully coalesced.
  // Wpad/2 is 16, 32 or 64: shifts and masks, no 64-bit division in the loop
  const int half = Wpad >> 1, hshift = 31 - __clz(half);
  const int w2 = (threadIdx.x & (half - 1)) * 2;
  const int64_t rows_per_step = ((int64_t)gridDim.x * blockDim.x) >> hshift;
  const double2* FA2 = reinterpret_cast<const double2*>(FA + w2);
  const double2* FB2 = reinterpret_cast<const double2*>(FB + w2);
  for (int64_t u = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> hshift; u < U; u += rows_per_step) {
    uint32_t cA, cB;
    if (rowcode) {
      uint32_t code = __ldg(rowcode + u);
      cA = (code >> kAgeShift) & kAgeMask;
      cB = code & kDustMask;
    } else {
      cA = (uint32_t)u / (uint32_t)nB;          // U = nA * nB <= 2^27 here
      cB = (uint32_t)u - cA * (uint32_t)nB;
    }
    const double2 fa = __ldg(FA2 + (size_t)cA * half);
    const double2 fb = __ldg(FB2 + (size_t)cB * half);
    __stcs(reinterpret_cast<double2*>(T + (size_t)u * Wpad + w2), make_double2(fa.x * fb.x, fa.y * fb.y));
  }
}

// ---- K1e / K1b: the prior weights evaluated ELEMENTWISE over the grid's physics columns ---------------
// singlepop_dust_model.py:131-148 taken literally: physmod[g][w] = prod_p f_p(col_p[g]; theta_p(phi_w))
// for every grid row, no class coding -- the path for grids whose columns are not a small set of
// distinct values (non-Cartesian grids).  A block owns tiles of kK1eRows rows:
//   phase 1  one thread per (row, parameter): what depends on the row only -- ln x and 1/x for the
//            lognormals, the interval (and offset) for the binned models, 10^x for the exponential;
//            column loads are coalesced along the rows
//   phase 2  one thread per (row, walker pair): the per-walker part (one fp64 exp per lognormal),
//            product in the reference's left-to-right order, 16-byte streaming stores, walker minor
// Per-walker constants (mu, 1/sigma, inv_sqrt_2pi/sigma, ...) are derived from phi once per block.
// Reads 8 P bytes per row, writes 8 Wpad bytes per row: HBM-write / fp64-exp bound.
constexpr int kK1eThreads = 256;
constexpr int kK1eRows = 256;
constexpr int kK1eWalkerVals = 6;   // per (parameter, walker) derived constants
struct K1eArgs {
  ModelDev model;
  const double* col[MB_NPARAM];  // device columns of the free parameters [G]
  const double* tab[MB_NPARAM];  // MB_TABULATED dust parameters: the host callable's values on the whole column,
                                 // [W][G] (every grid row is its own class here); null otherwise
  const double* phi;             // [W][ndim]
  double* T;                     // [G][Wpad]
  // N-stars term (helpers.py:133-162), optional: rows are visited in (age, Z)-bin order through
  // `rowlist`, tiles never straddle a bin, and every tile leaves its partial bin sums
  //   partial[tile][w] = (sum physmod grid_weight, sum physmod grid_weight completeness)
  // so the table is never read back; mb_k1b_binsums adds the tiles of each bin in fixed order
  const int32_t* rowlist;        // [G] grid rows in bin order (null: natural order, no partial sums)
  const int64_t* tile_start;     // [ntiles + 1] positions in rowlist (null: tiles of kK1eRows rows)
  const double* gw;              // grid_weight [G]
  const double* compl_;          // completeness [G]
  double* partial;               // [ntiles][Wpad][2]
  double* Fage;                  // [nbins][Wpad]: age factor of each (age, Z) bin
  const double* bin_age;         // [nbins] logA of the bin
  int64_t G, ntiles;
  int32_t W, Wpad, ndim, nbins;
};

// exp(x) for x <= 0 (K1e's arguments are -t^2/2), NaN passed through.  The library's scheme --
// k = round(x / ln 2), r = x - k ln 2 in two parts, exp(r) by a degree-13 polynomial (|r| <= 0.347:
// truncation 4e-18), scaled by 2^k in one or two exact steps so that results below 2^-1022 round
// once into the subnormals -- with the coefficients held in REGISTERS by the caller: K1e evaluates
// four of these per (row, walker pair), and the inlined library exp re-materialised its constants
// each time (a third of that kernel's instructions).  Against numpy's exp on 2 x 10^6 arguments
// in [-746, 0]: <= 1 ulp in the normal range, <= 1 unit in the subnormals, the same exact zeros.
struct ExpCoef { double c[14]; };
__constant__ double kExpPoly[14] = {
    1.0, 1.0, 0.5, 1.0 / 6, 1.0 / 24, 1.0 / 120, 1.0 / 720, 1.0 / 5040, 1.0 / 40320, 1.0 / 362880,
    1.0 / 3628800, 1.0 / 39916800, 1.0 / 479001600, 1.0 / 6227020800.0};
__device__ __forceinline__ double exp_nonpos(double x, const ExpCoef& e) {
  // branch free: arguments below -746 are clamped (the two-step scaling then rounds p 2^-1076 to
  // exactly zero, as exp underflows below -745.14); NaN is restored by a select at the end
  const double xc = fmax(x, -746.0);
  const double shifter = 6755399441055744.0;    // 1.5 * 2^52: the low word of x*log2e + shifter is round(x*log2e)
  const double t = fma(xc, 1.4426950408889634074, shifter);
  const int k = __double2loint(t);
  const double kd = t - shifter;
  double r = fma(-kd, 6.93147180369123816490e-01, xc);
  r = fma(-kd, 1.90821492927058770002e-10, r);
  double p = e.c[13];
#pragma unroll
  for (int i = 12; i >= 0; --i) p = fma(p, r, e.c[i]);
  // 2^k in two exact power-of-two factors: only the last multiply can round (into the subnormals)
  const int k1 = k >> 1, k2 = k - k1;
  const double res = (p * __hiloint2double((k1 + 1023) << 20, 0)) * __hiloint2double((k2 + 1023) << 20, 0);
  return x != x ? x : res;
}

// derived per-walker constants of parameter p -> c[0..5]
__device__ __forceinline__ void k1e_walker_consts(const ModelDev& m, int p, const double* th, double* c) {
  const double inv_sqrt_2pi = 0.3989422804014326779399460599343819;
  switch (m.kind[p]) {
    case MB_LOGNORMAL:
      c[0] = log(th[0]) + th[1] * th[1]; c[1] = 1.0 / th[1]; c[2] = inv_sqrt_2pi / th[1];
      break;
    case MB_TWO_LOGNORMAL: {
      const double n1 = 1.0 - 1.0 / (th[4] + 1.0), n2 = 1.0 / (th[4] + 1.0);
      c[0] = log(th[0]) + th[2] * th[2]; c[1] = 1.0 / th[2]; c[2] = n1 * inv_sqrt_2pi / th[2];
      c[3] = log(th[1]) + th[3] * th[3]; c[4] = 1.0 / th[3]; c[5] = n2 * inv_sqrt_2pi / th[3];
      break;
    }
    case MB_EXPONENTIAL:
      c[0] = th[0] * 1e9;
      break;
    default:
      break;
  }
}
// row part: two doubles and an interval index / validity flag
__device__ __forceinline__ void k1e_row_part(const ModelDev& m, int p, double x, double* r0, double* r1, int* ri) {
  *r0 = 0.0; *r1 = 0.0; *ri = 0;
  switch (m.kind[p]) {
    case MB_LOGNORMAL: case MB_TWO_LOGNORMAL:
      if (x > 0.0) { *r0 = log(x); *r1 = 1.0 / x; *ri = 1; }
      break;
    case MB_BINS_HISTO: {
      int n = m.nnodes[p], i = 0;
      while (i + 1 < n && x >= m.nodes[p][i + 1]) ++i;
      *ri = i;
      break;
    }
    case MB_BINS_INTERP: {
      int n = m.nnodes[p];
      if (x <= m.nodes[p][0]) { *ri = 0; *r1 = -1.0; }               // th[0]
      else if (x >= m.nodes[p][n - 1]) { *ri = n - 1; *r1 = -1.0; }  // th[n-1]
      else {
        int i = 0;
        while (x >= m.nodes[p][i + 1]) ++i;
        *ri = i; *r0 = x - m.nodes[p][i]; *r1 = m.nodes[p][i + 1] - m.nodes[p][i];
      }
      break;
    }
    case MB_EXPONENTIAL:
      *r0 = pow(10.0, x);
      break;
    default:
      break;
  }
}
__device__ __forceinline__ double k1e_eval(int kind, const double* th, const double* c, double r0, double r1, int ri) {
  switch (kind) {
    case MB_LOGNORMAL: {
      if (!ri) return 0.0;
      const double t = (r0 - c[0]) * c[1];
      return (c[2] * r1) * exp(-0.5 * (t * t));
    }
    case MB_TWO_LOGNORMAL: {
      if (!ri) return 0.0;
      const double t1 = (r0 - c[0]) * c[1], t2 = (r0 - c[3]) * c[4];
      return (c[2] * r1) * exp(-0.5 * (t1 * t1)) + (c[5] * r1) * exp(-0.5 * (t2 * t2));
    }
    case MB_BINS_HISTO:
      return th[ri];
    case MB_BINS_INTERP:
      if (r1 < 0.0) return th[ri];
      return (th[ri + 1] - th[ri]) / r1 * r0 + th[ri];
    case MB_EXPONENTIAL:
      return exp(-1.0 * r0 / c[0]);
    default:  // MB_FLAT
      return 1.0;
  }
}

// LN = the common shape (docs example): every free dust parameter is a single lognormal.  Its
// per-walker constants then live in registers for the whole kernel (no shared-memory reads, no
// dispatch on the model kind per evaluation) and exp runs with register-resident coefficients.
template <bool LN>
__global__ void __launch_bounds__(kK1eThreads, LN ? 2 : 3) mb_k1_elementwise(const __grid_constant__ K1eArgs a) {
  extern __shared__ __align__(16) double k1e_smem[];
  const ModelDev& m = a.model;
  const int Wpad = a.Wpad, ndim = a.ndim;
  // smem: phi [Wpad][ndim] | consts [NPARAM][Wpad][6] | row parts r0, r1 [NPARAM][rows] | weights
  // gw, gw*compl [rows] | combine [threads][4] | ri [NPARAM][rows] | rows [rows]
  double* sphi = k1e_smem;
  double* scon = sphi + (size_t)Wpad * ndim;
  double* sr0 = scon + (size_t)MB_NPARAM * Wpad * kK1eWalkerVals;
  double* sr1 = sr0 + MB_NPARAM * kK1eRows;
  double* sgw = sr1 + MB_NPARAM * kK1eRows;
  double* sgc = sgw + kK1eRows;
  double* scomb = sgc + kK1eRows;
  int* sri = reinterpret_cast<int*>(scomb + 4 * kK1eThreads);
  int* srow = sri + MB_NPARAM * kK1eRows;
  for (int i = threadIdx.x; i < Wpad * ndim; i += kK1eThreads) {
    const int w = i / ndim, k = i - w * ndim;
    sphi[i] = a.phi[(size_t)(w < a.W ? w : a.W - 1) * ndim + k];  // padding columns repeat the last walker
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MB_NPARAM * Wpad; i += kK1eThreads) {
    const int p = i / Wpad, w = i - p * Wpad;
    if (m.kind[p] != MB_FIXED) k1e_walker_consts(m, p, sphi + (size_t)w * ndim + m.voff[p], scon + (size_t)i * kK1eWalkerVals);
  }
  __syncthreads();
  // age factor of every (age, Z) bin for the Poisson term (block 0 only; tiny)
  if (a.Fage && blockIdx.x == 0) {
    for (int i = threadIdx.x; i < a.nbins * Wpad; i += kK1eThreads) {
      const int b = i / Wpad, w = i - b * Wpad;
      double v = 1.0;
      if (m.kind[MB_P_LOGA] != MB_FIXED) {
        double r0, r1; int ri;
        k1e_row_part(m, MB_P_LOGA, a.bin_age[b], &r0, &r1, &ri);
        v = k1e_eval(m.kind[MB_P_LOGA], sphi + (size_t)w * ndim + m.voff[MB_P_LOGA],
                     scon + ((size_t)MB_P_LOGA * Wpad + w) * kK1eWalkerVals, r0, r1, ri);
      }
      a.Fage[i] = v;
    }
  }
  const int half = Wpad >> 1;
  const int rows_per_pass = kK1eThreads / half;
  const int pair = threadIdx.x % half, rlane = threadIdx.x / half;
  int kinds[MB_NPARAM];
#pragma unroll
  for (int p = 0; p < MB_NPARAM; ++p) kinds[p] = m.kind[p];
  // LN: this thread's two walkers' lognormal constants (mu, 1/sigma, inv_sqrt_2pi/sigma) per dust
  // parameter, and the exp coefficients
  double lc[LN ? 3 : 1][2][3];
  ExpCoef ec;
  if (LN) {
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int w = 0; w < 2; ++w)
#pragma unroll
        for (int k = 0; k < 3; ++k)
          lc[LN ? j : 0][w][k] = kinds[MB_P_AV + j] == MB_LOGNORMAL
                                     ? scon[((size_t)(MB_P_AV + j) * Wpad + 2 * pair + w) * kK1eWalkerVals + k] : 0.0;
#pragma unroll
    for (int i = 0; i < 14; ++i) ec.c[i] = kExpPoly[i];
  }
  for (int64_t tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const int64_t i0 = a.tile_start ? a.tile_start[tile] : tile * kK1eRows;
    const int nrows = (int)((a.tile_start ? a.tile_start[tile + 1] : min(a.G, i0 + kK1eRows)) - i0);
    __syncthreads();  // row parts of the previous tile are no longer read
    for (int r = threadIdx.x; r < nrows; r += kK1eThreads) {
      const int g = a.rowlist ? __ldg(a.rowlist + i0 + r) : (int)(i0 + r);
      srow[r] = g;
      if (a.partial) {
        const double w = __ldg(a.gw + g);
        sgw[r] = w; sgc[r] = w * __ldg(a.compl_ + g);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < MB_NPARAM * kK1eRows; i += kK1eThreads) {
      const int p = i / kK1eRows, r = i - p * kK1eRows;
      if (m.kind[p] != MB_FIXED && m.kind[p] != MB_TABULATED && r < nrows) {
        double r0, r1; int ri;
        k1e_row_part(m, p, __ldg(a.col[p] + srow[r]), &r0, &r1, &ri);
        sr0[i] = r0; sr1[i] = r1; sri[i] = ri;
      }
    }
    __syncthreads();
    double s0 = 0.0, s1 = 0.0, c0 = 0.0, c1 = 0.0;
    // physmod of one grid row for this thread's two walkers
    auto eval_row = [&](int r, double& v0, double& v1) {
      v0 = 1.0; v1 = 1.0;
      // LN: the lognormal dust factors share ONE exp per walker -- the product of
      // (c2 / x) exp(-t^2 / 2) over the parameters is (prod c2 / x) exp(-sum t^2 / 2); the fp64 exp is
      // what bounds this kernel.  Where the summed exponent is below -700 (or not a number) the
      // factors are evaluated and multiplied one by one as the reference does, so that underflow to
      // subnormals / zero happens exactly where numpy's happens.
      double e0 = 0.0, e1 = 0.0, pf0 = 1.0, pf1 = 1.0;
#pragma unroll
      for (int p = 0; p < MB_NPARAM; ++p) {  // logA, Av, Rv, fA: the reference's `cur_physmod *=` order
        if (kinds[p] == MB_FIXED) continue;
        const int i = p * kK1eRows + r;
        if (LN && p == MB_P_LOGA && kinds[p] == MB_BINS_HISTO) {  // step function: the row's interval picks the height
          const double* th = sphi + (size_t)(2 * pair) * ndim + m.voff[MB_P_LOGA] + sri[i];
          v0 = th[0]; v1 = th[ndim];
          continue;
        }
        if (LN && p >= MB_P_AV) {
          const double lx = sr0[i], rx = sr1[i];
          const double (&c)[2][3] = lc[LN ? p - MB_P_AV : 0];
          // x <= 0: the row part holds ln x = 0 and 1/x = 0, so the factor is exactly 0 as it must be
          const double t0 = (lx - c[0][0]) * c[0][1], t1 = (lx - c[1][0]) * c[1][1];
          e0 = fma(-0.5 * t0, t0, e0); e1 = fma(-0.5 * t1, t1, e1);
          pf0 *= c[0][2] * rx; pf1 *= c[1][2] * rx;
          continue;
        }
        if (kinds[p] == MB_TABULATED) {  // (padding columns repeat the last walker)
          const size_t g = (size_t)srow[r];
          v0 *= __ldg(a.tab[p] + (size_t)min(2 * pair, a.W - 1) * a.G + g);
          v1 *= __ldg(a.tab[p] + (size_t)min(2 * pair + 1, a.W - 1) * a.G + g);
          continue;
        }
        const double* th0 = sphi + (size_t)(2 * pair) * ndim + m.voff[p];
        const double* cc = scon + ((size_t)p * Wpad + 2 * pair) * kK1eWalkerVals;
        const double r0 = sr0[i], r1 = sr1[i];
        const int ri = sri[i];
        v0 *= k1e_eval(kinds[p], th0, cc, r0, r1, ri);
        v1 *= k1e_eval(kinds[p], th0 + ndim, cc + kK1eWalkerVals, r0, r1, ri);
      }
      if (LN) {
        if (e0 > -700.0 && e1 > -700.0) {
          v0 *= pf0 * exp_nonpos(e0, ec);
          v1 *= pf1 * exp_nonpos(e1, ec);
        } else {  // rare: factor by factor, in the reference's order
#pragma unroll
          for (int p = MB_P_AV; p < MB_NPARAM; ++p) {
            if (kinds[p] != MB_LOGNORMAL) continue;
            const int i = p * kK1eRows + r;
            const double lx = sr0[i], rx = sr1[i];
            const double (&c)[2][3] = lc[LN ? p - MB_P_AV : 0];
            const double t0 = (lx - c[0][0]) * c[0][1], t1 = (lx - c[1][0]) * c[1][1];
            v0 *= (c[0][2] * rx) * exp_nonpos(-0.5 * (t0 * t0), ec);
            v1 *= (c[1][2] * rx) * exp_nonpos(-0.5 * (t1 * t1), ec);
          }
        }
      }
    };
    auto emit_row = [&](int r, double v0, double v1) {
      __stcs(reinterpret_cast<double2*>(a.T + (size_t)srow[r] * Wpad + 2 * pair), make_double2(v0, v1));
      if (a.partial) {
        s0 = fma(v0, sgw[r], s0); s1 = fma(v1, sgw[r], s1);
        c0 = fma(v0, sgc[r], c0); c1 = fma(v1, sgc[r], c1);
      }
    };
    // two rows per iteration: two independent exp chains in flight per thread
    int r = rlane;
    for (; r + rows_per_pass < nrows; r += 2 * rows_per_pass) {
      double va0, va1, vb0, vb1;
      eval_row(r, va0, va1);
      eval_row(r + rows_per_pass, vb0, vb1);
      emit_row(r, va0, va1);
      emit_row(r + rows_per_pass, vb0, vb1);
    }
    if (r < nrows) {
      double v0, v1;
      eval_row(r, v0, v1);
      emit_row(r, v0, v1);
    }
    if (a.partial) {  // combine the row lanes in fixed order -> this tile's partial bin sums
      double* mine = scomb + 4 * threadIdx.x;
      mine[0] = s0; mine[1] = s1; mine[2] = c0; mine[3] = c1;
      __syncthreads();
      if (rlane == 0) {
        for (int l = 1; l < rows_per_pass; ++l) {
          const double* o = scomb + 4 * (l * half + pair);
          s0 += o[0]; s1 += o[1]; c0 += o[2]; c1 += o[3];
        }
        double2* out = reinterpret_cast<double2*>(a.partial) + (size_t)tile * Wpad + 2 * pair;
        out[0] = make_double2(s0, c0);
        out[1] = make_double2(s1, c1);
      }
    }
  }
}

// K1b: bin sums S, C (helpers.py:133-162) = the tiles' partial sums added per bin in tile order.
// One block per bin, thread = (walker, tile lane); lanes combined in fixed order.
constexpr int kK1bThreads = 256;
__global__ void __launch_bounds__(kK1bThreads) mb_k1b_binsums(const double* __restrict__ partial,
                                                              const int64_t* __restrict__ bin_tile,
                                                              double* __restrict__ SC, int Wpad) {
  __shared__ double2 part[kK1bThreads];
  const int b = blockIdx.x;
  const int w = threadIdx.x % Wpad, lane = threadIdx.x / Wpad, nlanes = kK1bThreads / Wpad;
  const double2* P = reinterpret_cast<const double2*>(partial);
  double s = 0.0, c = 0.0;
  for (int64_t t = bin_tile[b] + lane; t < bin_tile[b + 1]; t += nlanes) {
    const double2 v = __ldg(P + (size_t)t * Wpad + w);
    s += v.x; c += v.y;
  }
  part[threadIdx.x] = make_double2(s, c);
  __syncthreads();
  if (lane == 0) {
    for (int l = 1; l < nlanes; ++l) { s += part[l * Wpad + w].x; c += part[l * Wpad + w].y; }
    reinterpret_cast<double2*>(SC)[(size_t)b * Wpad + w] = make_double2(s, c);
  }
}

// ---- K2: the hot kernel --------------------------------------------------------------------------
// log of a running product: ln(prod L_s) = ln(P) + E ln 2 with P kept in [1, 2^k).
// One DMUL + a few integer ops per star instead of a ~50-instruction fp64 log.
struct LogAcc {
  double P;        // running product of mantissas, in [1, 2^k)
  int32_t E;       // running sum of BIASED exponents (bias removed in take_value)
  int32_t n;       // stars folded in since the last renormalisation of P (<= 512) / since init
  int32_t nb;      // stars folded in since take_value (for the bias)
  int32_t bad;     // stars whose sum was not finite -> ValueError (:216-217)
  int32_t isnan;   // a negative sum: np.log gives nan, no exception
  __device__ __forceinline__ void init() { P = 1.0; E = 0; n = 0; nb = 0; bad = 0; isnan = 0; }
  __device__ __forceinline__ void renorm() {
    int hi = __double2hiint(P);
    int e = ((hi >> 20) & 0x7ff);
    if (e != 0x7ff && e != 0) {  // leave inf/nan alone
      E += e - 1023;
      P = __hiloint2double((hi & 0x800fffff) | (1023 << 20), __double2loint(P));
    }
    n = 0;
  }
  // ln of the product so far; the product restarts, the bad / nan flags persist
  __device__ __forceinline__ double take_value() {
    renorm();
    double v = log(P) + (double)(E - 1023 * nb) * 0.693147180559945309417232121458;
    P = 1.0; E = 0; n = 0; nb = 0;
    return v;
  }
};
// Segment log-sums are accumulated as TWO 64-bit fixed-point words: hi counts units of 2^-24, lo
// the remainder in units of 2^-68 (|lo| <= 2^43 per segment).  Integer addition is exact, so the
// total is independent of the order in which warps, blocks and GPUs contribute -- bitwise
// reproducible under dynamic scheduling -- and the absolute resolution (3.4e-21 per segment) keeps
// 1e-9 RELATIVE accuracy down to totals of ~1e-11.  Range: |total| < 2^39 = 5.5e11 in hi; a star's
// ln L lies in [-745, 710], so that takes > 7e8 stars per job, more than 16 GPUs can hold; lo has
// room for 2^20 segments (a launch has at most a few thousand per GPU).
constexpr double kFixHiScale = 16777216.0;          // 2^24
constexpr double kFixLoScale = 17592186044416.0;    // 2^44 (on top of 2^24)
constexpr int kAccRows = 4;                         // hi, lo, non-finite stars, nan flags
struct Fix2 { long long hi, lo; };
__device__ __forceinline__ Fix2 to_fix2(double v) {
  Fix2 f;
  const double vs = v * kFixHiScale;                // exact (power of two)
  f.hi = __double2ll_rn(vs);
  f.lo = __double2ll_rn((vs - (double)f.hi) * kFixLoScale);  // the difference is exact
  return f;
}
// The two words as doubles, after carrying the whole units of lo into hi: lo is then in [0, 2^44) and
// converts exactly, hi converts exactly while |total| < 2^29 = 5.4e8 -- so words of several shards can
// be added in floating point without rounding, and hi + lo is the only rounding of the whole sum.
__device__ __forceinline__ double fix_hi_to_double(long long hi, long long lo) {
  return (double)(hi + (lo >> 44)) * (1.0 / kFixHiScale);
}
__device__ __forceinline__ double fix_lo_to_double(long long lo) {
  return (double)(lo & ((1ll << 44) - 1)) * (1.0 / (kFixHiScale * kFixLoScale));
}
enum { kCtrSegment = 0, kCtrBlocks = 1, kCtrWords = 2 };

// Rare star outcomes, kept out of line: exact zero (skipped, :218-221), non-finite (counted ->
// ValueError, :216-217), negative (np.log gives nan), subnormal (rescaled by 2^64 first).
__device__ __noinline__ LogAcc star_done_slow(LogAcc acc, double L) {
  if (L == 0.0) return acc;
  int hi = __double2hiint(L);
  int e = (hi >> 20) & 0x7ff;
  if (e == 0x7ff) { acc.bad++; return acc; }
  if (hi < 0) { acc.isnan = 1; return acc; }
  L *= 18446744073709551616.0;  // subnormal: 2^64
  hi = __double2hiint(L);
  acc.E += ((hi >> 20) & 0x7ff) - 64;
  acc.nb++;
  acc.P *= __hiloint2double((hi & 0x000fffff) | (1023 << 20), __double2loint(L));
  if (++acc.n >= 512) acc.renorm();
  return acc;
}

// End of a star: fold its integrated probability L into the walker's accumulator.  The common
// case (L positive, normal) is a handful of integer ops and one DMUL.
__device__ __forceinline__ void star_done(LogAcc& acc, double L) {
  const int hi = __double2hiint(L);
  const unsigned e = (unsigned)hi >> 20;         // sign + exponent field
  if (e - 1u < 2046u) {                          // 1 <= e <= 2046: positive, normal, finite
    acc.E += (int)e;
    acc.nb++;
    acc.P *= __hiloint2double((hi & 0x000fffff) | (1023 << 20), __double2loint(L));
    if (++acc.n >= 512) acc.renorm();
  } else {
    acc = star_done_slow(acc, L);
  }
}

enum { kModeFactor = 1, kModeTable = 2, kModeTableGrid = 3 };

// ---- star shards on several GPUs: reduction over NVLink peer memory, fused into K2's tail --------
// Every GPU owns an exchange buffer (cudaMalloc, opened by the other ranks through CUDA IPC):
//   words [2 parities][MB_MAX_PEERS sources][2 * kXWords] u64
// The last block of rank r's K2 sends its shard totals -- (hi, lo) fixed-point log-sum, non-finite and
// nan counts per walker, kXWords 64-bit values -- into slot r of EVERY other rank's buffer as
// 8-byte words that carry their own flag: [63:32] = sequence number of the launch, [31:0] = one
// half of a value.  An aligned 8-byte store is a single transaction, so a reader that sees the
// sequence number of its launch in a word holds valid data: no fence, no separate flag, one NVLink
// one-way trip (the protocol NCCL calls LL).  Every rank then polls the words of all other sources in
// its own buffer and adds the values to its own.  Integer sums: the order is irrelevant, every rank
// gets bitwise the same totals.  Two parities: a rank can be at most one launch ahead of the slowest
// one (it needs everyone's words of launch n to leave launch n), so launch n + 2 never overwrites
// words of launch n that are still being read.
constexpr int kXWords = kAccRows * 128;
__host__ __device__ constexpr size_t xbuf_word_off(int par, int src) {   // in u64 words
  return ((size_t)par * MB_MAX_PEERS + src) * 2 * kXWords;
}
// rendezvous flags (mb_peer_rendezvous): [2 parities][MB_MAX_PEERS sources] u32 after the words
__host__ __device__ constexpr size_t xbuf_rdv_off(int par, int src) {   // in u32 words
  return 2 * (size_t)2 * MB_MAX_PEERS * 2 * kXWords + (size_t)par * MB_MAX_PEERS + src;
}
constexpr size_t kXBufBytes = 8 * (size_t)2 * MB_MAX_PEERS * 2 * kXWords + 4 * (size_t)2 * MB_MAX_PEERS + 64;
struct PeerArgs {
  int32_t world, rank;   // world <= 1: single shard, no exchange
  uint32_t seq;          // sequence number of this launch, identical on all ranks
  unsigned long long timeout_ns;  // how long the last block waits for a peer's shard (MB_PEER_TIMEOUT_MS)
  unsigned long long* xbuf[MB_MAX_PEERS];  // exchange buffer of every rank, as mapped on THIS GPU
};

#ifndef MB_PHI_INLINE
#define MB_PHI_INLINE 1152
#endif
// Device-side rendezvous of the ranks of a peer group: every rank raises its flag in all peers and
// waits for theirs, so all ranks leave within one NVLink latency of the last arrival.  Used to align
// the ranks ahead of a timed launch (bench.py) -- the likelihood itself never needs it.
__global__ void mb_rendezvous(PeerArgs peer) {
  const int t = threadIdx.x, par = (int)(peer.seq & 1u);
  if (t >= peer.world || t == peer.rank) return;
  *(reinterpret_cast<volatile unsigned int*>(peer.xbuf[t]) + xbuf_rdv_off(par, peer.rank)) = peer.seq;
  volatile unsigned int* mine = reinterpret_cast<volatile unsigned int*>(peer.xbuf[peer.rank]) + xbuf_rdv_off(par, t);
  unsigned long long t0 = 0ull;
  unsigned int spins = 0;
  while (*mine != peer.seq) {
    if ((++spins & 0x3ffu) == 0u) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
      if (t0 == 0ull) t0 = t1;
      if (t1 - t0 > peer.timeout_ns) break;  // give up quietly: this is only an alignment aid
    }
  }
}

constexpr int kPhiInline = MB_PHI_INLINE;  // doubles of phi carried in K2's kernel parameters (e.g. 128 walkers x 9)
struct K2Args {
  const double* ea;     // entry weights  [n_entries]
  const void* ec;       // entry codes    [n_entries]: uint32, or uint16 for the C16 kernels (padded by one)
  const int64_t* seg;   // [nseg+1] entry offsets, star aligned
  int32_t nseg;
  const double* FA;     // FACTOR: outer rows [nA][Wpad] then inner rows [nB][Wpad], contiguous (staged to smem)
  const double* T;      // TABLE*: [U][Wpad]
  unsigned long long* gacc;  // [kAccRows][Wpad] fixed-point log-sum (hi, lo), non-finite stars, nan flags (zero between launches)
  unsigned int* counters;    // [kCtrWords] next segment, blocks done (zero between launches)
  double* out;          // [MB_OUT_ROWS][Wtot], written at columns w0 .. w0+W-1
  int32_t nA, nB, Wpad, W, w0, Wtot;   // nA = rows of the OUTER table (age classes x nOd), nB = rows of the INNER table
  int32_t nOd;               // outer-dust combinations per age class (1: the classic age | dust split)
  int32_t tab_off;           // byte offset of the factor tables in dynamic shared memory
  int32_t scratch_off;       // byte offset of the table-build scratch: 0 (it shares the entry stage) or right
                             // after the stage (`separate` layout: the stage can be primed before the build)
  int32_t separate;          // 1: stage and scratch are disjoint
  int32_t use_tma;           // 1: the entry ring is filled by bulk copies (TMA) completing on mbarriers
  int32_t slab_off;          // byte offset of the block-reduction slabs (0, or tab_off: over the dead tables)
  PoissonArgs pois;          // N-stars term: every block computes some (age, Z) bins before the star loop
  PeerArgs peer;             // cross-GPU reduction of the shard totals
  unsigned long long* timeline;  // optional [gridDim.x][kTimelinePoints] %globaltimer stamps (profiling)
  // FACTOR mode, fp64 tables, analytic models: every block builds the factor tables itself in its
  // prologue (K0 fused into K2: no separate launch, no dependent-kernel gap); phi comes with the
  // kernel parameters when it is small enough, else from `phi` (device or mapped host memory)
  volatile unsigned long long* done_flag;  // optional (mapped host memory): set to done_seq after the rows
  unsigned long long done_seq;
  int32_t build_tables, phi_inline, ndim, ncls_total;
  const double* clsx;        // class values [ncls_total]
  const uint16_t* rowsrc;    // [nA + nB][4] the look-up rows multiplied into a table row, as SLOTS (see lutslot);
                             // 0xffff = none, first = 0xfffe: the row is a look-up row evaluated in place
  const uint16_t* lutslot;   // [ncls_total] where a look-up row lives while the tables are built: bit 15 set =
                             // row (& 0x7fff) of the tables themselves (a parameter alone in its table is
                             // evaluated in place), else row of the scratch look-up table
  int32_t nscratch;          // rows of the scratch look-up table
  const double* phi;         // [W][ndim], already offset to this pass's first walker
  ModelDev model;
  double phi_in[kPhiInline];
};
// -DMB_WARP_STAMPS (measurement builds only, scripts/k2_warp_stamps.py): per warp, after the 8 block stamps --
// table build left, star loop entered (after its N-stars bin, if it took one), star loop left, entries of its
// first segment (as block start + 1000 ns per entry)
#ifdef MB_WARP_STAMPS
constexpr int kWarpStampWarps = 24, kWarpStampPoints = 4;
#else
constexpr int kWarpStampWarps = 0, kWarpStampPoints = 0;
#endif
constexpr int kTimelinePoints = 8 + kWarpStampWarps * kWarpStampPoints;
                                    // start, tables staged, N-stars bins done, stars done, block reduced,
                                    // (last block:) totals + Poisson term + sent to peers, peers received, end
__device__ __forceinline__ void stamp(unsigned long long* timeline, int point) {
  if (timeline && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    timeline[(size_t)blockIdx.x * kTimelinePoints + point] = t;
  }
}

// block shape: ONE block per SM -- 768 threads (<= 85 registers) for up to 64 walkers per pass,
// 512 threads (<= 128 registers) for 128; one copy of the factor tables per SM
#ifndef MB_K2_THREADS_LO
#define MB_K2_THREADS_LO 768
#endif
__host__ __device__ constexpr int k2_threads(int NW) { return NW <= 2 ? MB_K2_THREADS_LO : 512; }
__host__ __device__ constexpr int k2_min_blocks(int NW) { return 1; }
constexpr int kChunk = 32;                          // entries staged per warp per step
#ifndef MB_K2_STAGES
#define MB_K2_STAGES 4
#endif
constexpr int kStages = MB_K2_STAGES;                    // chunks in flight per warp (ring of buffers)
constexpr int kStageBytes = kStages * kChunk * (8 + 4);  // per warp: weights ring + codes ring
// The 128-walker kernel issues twice the factor loads per entry, so two chunks in flight hide the
// stream as well as four (measured on B200: 215 us with 2, 3 or 4) -- and the smaller ring lets
// 128 walkers x 200 dust classes (config 4: 210 KB of tables) run as ONE pass.
__host__ __device__ constexpr int k2_stages(int NW) { return NW == 4 ? 2 : kStages; }
__host__ __device__ constexpr int k2_stage_bytes(int NW) { return k2_stages(NW) * kChunk * (8 + 4); }

__device__ __forceinline__ void cp_async8(uint32_t smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(uint32_t smem, const void* gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(smem), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}
// staged entry data (ordered with the cp.async waits): warp-uniform addresses = broadcasts
__device__ __forceinline__ uint4 lds_stage_u4(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double2 lds_stage_d2(uint32_t addr) {
  double2 v;
  asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double lds_stage_d(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint2 lds_stage_u2(uint32_t addr) {
  uint2 v;
  asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_stage_u16(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];\n" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds_stage_u(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ double2 lds_f64x2(uint32_t addr) {  // read-only factor tables
  double2 v;
  asm("ld.shared.v2.f64 {%0,%1}, [%2];\n" : "=d"(v.x), "=d"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm("ld.shared.f64 %0, [%1];\n" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint2 lds_u32x2(uint32_t addr) {
  uint2 v;
  asm("ld.shared.v2.u32 {%0,%1}, [%2];\n" : "=r"(v.x), "=r"(v.y) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm("ld.shared.u32 %0, [%1];\n" : "=r"(v) : "r"(addr));
  return v;
}
// high word of a double rounded to 21 significant bits (round to nearest on the dropped word)
__device__ __forceinline__ uint32_t round_to_hi_word(double v) {
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  if (((b >> 52) & 0x7ff) != 0x7ff) b += 0x80000000ull;  // leave inf / nan alone
  return (uint32_t)(b >> 32);
}

// Launder a loop-invariant address through an opaque move: without it ptxas, short of registers,
// re-derives the shared-memory addresses from %tid / %cgaid / kernel parameters inside the hot
// loop (~20 instructions per group of four entries in the first build of this kernel).
__device__ __forceinline__ uint32_t keep_in_register(uint32_t x) {
  uint32_t y;
  asm volatile("mov.u32 %0, %1;\n" : "=r"(y) : "r"(x));
  return y;
}

// Per-lane state of the NW walkers a lane owns while its warp walks stars.
// Walker mapping: NW == 1: w = lane;  NW >= 2: for q < NW/2, w = 64 q + 2 lane + {0,1}
// (one 16-byte shared/global load per q, conflict free).
// TF = element type of the FACTOR-mode tables in shared memory: double (default, exact) or a
// 32-bit type (optional: half the shared-memory traffic; the element is the HIGH WORD of the
// factor rounded to 21 significant bits, so it is used as a double without any conversion
// instruction -- F2F.F64.F32 runs on the quarter-rate XU pipe and made a true fp32 table no
// faster than the fp64 one; sums are still fp64).
// OB > 0: the FACTOR code word packed in 16 bits -- [15] new run, [14] new star, then OB bits of
// outer class and 14 - OB bits of inner class (OB = 4: <= 16 outer x 1,024 inner classes, the classic
// age | dust split; OB = 6: <= 64 x 256, the separable split of production dust axes) -- so that
// four codes are ONE 8-byte broadcast load instead of a 16-byte one (one shared-memory wavefront
// less per four entries).  An even entry's code arrives with the next one in its upper half; the
// masks below never look there.  OB = 0: 32-bit code words.
template <int NW, int MODE, typename TF = double, int OB = 0>
struct Lanes {
  static constexpr bool C16 = OB > 0;
  static constexpr uint32_t kInnerBits16 = 14 - OB;
  static_assert(!C16 || (MODE == 1 && sizeof(TF) == 8), "16-bit codes: FACTOR mode with fp64 tables");
  // dust() = inner class (row of the inner table), age() = outer class (row of the outer table)
  __device__ __forceinline__ static uint32_t dust(uint32_t code) { return C16 ? (code & ((1u << kInnerBits16) - 1u)) : (code & kDustMask); }
  __device__ __forceinline__ static uint32_t age(uint32_t code) { return C16 ? ((code >> kInnerBits16) & ((1u << OB) - 1u)) : ((code >> kAgeShift) & kAgeMask); }
  __device__ __forceinline__ static bool is_run(uint32_t code) { return C16 ? (code & 0x8000u) != 0 : (int32_t)code < 0; }
  __device__ __forceinline__ static bool is_star(uint32_t code) { return C16 ? (code & 0x4000u) != 0 : (code & kNewStar) != 0; }
  static constexpr int NQ = (NW + 1) / 2;
  static constexpr bool kF32 = sizeof(TF) == 4;
  static_assert(!kF32 || MODE == 1, "float tables exist in FACTOR mode only");
  // one table row: Wpad = 32 NW walkers
  static constexpr uint32_t rowbytes = 32u * NW * (MODE == 1 ? (uint32_t)sizeof(TF) : 8u);
  double run[NW], run2[NW];       // two partial sums of the current run (halves the DFMA chain)
  double star[NW], fa[NW];
  LogAcc acc[NW];
  uint32_t tabA, tabB;            // FACTOR: shared-memory byte addresses (lane offset included)
  const char* Tlane;              // TABLE*: global table base + lane offset

  __device__ __forceinline__ void init(uint32_t tabA_, uint32_t tabB_, const char* Tlane_) {
    tabA = keep_in_register(tabA_); tabB = keep_in_register(tabB_); Tlane = Tlane_;
#pragma unroll
    for (int i = 0; i < NW; ++i) { run[i] = 0.0; run2[i] = 0.0; star[i] = 0.0; fa[i] = 1.0; acc[i].init(); }
  }
  // factors of one entry for this lane's walkers
  // FACTOR: one table row element pair / element of this lane (byte address given)
  __device__ __forceinline__ static void lds_row(uint32_t addr, double* f) {
    if (kF32) {  // 32-bit elements = high words of the (rounded) doubles: no conversion needed
      if (NW == 1) f[0] = __hiloint2double((int)lds_u32(addr), 0);
      else {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          uint2 v = lds_u32x2(addr + 256u * q);
          f[2 * q] = __hiloint2double((int)v.x, 0); f[2 * q + 1] = __hiloint2double((int)v.y, 0);
        }
      }
    } else {
      if (NW == 1) f[0] = lds_f64(addr);
      else {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          double2 v = lds_f64x2(addr + 512u * q);
          f[2 * q] = v.x; f[2 * q + 1] = v.y;
        }
      }
    }
  }
  __device__ __forceinline__ void load_factors(uint32_t code, double* f) const {
    if (MODE == kModeFactor) {
      lds_row(dust(code) * rowbytes + tabB, f);
    } else {
      const char* p = Tlane + (size_t)(code & kRowMask) * rowbytes;
      if (NW == 1) f[0] = __ldg(reinterpret_cast<const double*>(p));
      else {
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          double2 v = __ldg(reinterpret_cast<const double2*>(p + 512 * q));
          f[2 * q] = v.x; f[2 * q + 1] = v.y;
        }
      }
    }
  }
  // boundary before an entry (warp-uniform): close the run (and the star), fetch the new age factor
  __device__ __forceinline__ void boundary(uint32_t code) {
#pragma unroll
    for (int i = 0; i < NW; ++i) { star[i] = fma(fa[i], run[i] + run2[i], star[i]); run[i] = 0.0; run2[i] = 0.0; }
    if (MODE != kModeFactor || is_star(code)) {
#pragma unroll
      for (int i = 0; i < NW; ++i) { star_done(acc[i], star[i]); star[i] = 0.0; }
    }
    if (MODE == kModeFactor) lds_row(age(code) * rowbytes + tabA, fa);
  }
  __device__ __forceinline__ void mac(double a, const double* f) {
#pragma unroll
    for (int i = 0; i < NW; ++i) run[i] = fma(a, f[i], run[i]);
  }
  __device__ __forceinline__ void mac2(double a, const double* f) {
#pragma unroll
    for (int i = 0; i < NW; ++i) run2[i] = fma(a, f[i], run2[i]);
  }
  __device__ __forceinline__ void step(double a, uint32_t code, const double* f) {
    if (is_run(code)) boundary(code);  // the new-run flag, warp-uniform
    mac(a, f);
  }
  // a streamed range ends on a star boundary: close the last star
  __device__ __forceinline__ void close() {
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      star_done(acc[i], fma(fa[i], run[i] + run2[i], star[i]));
      star[i] = 0.0; run[i] = 0.0; run2[i] = 0.0;
    }
  }
};

// ---- the entry ring: chunks of kChunk entries, STAGES slots per warp ---------------------------------
// Two ways to fill a slot.  TMA = false (the default build, -DMB_RING=0): per-lane cp.async (LDGSTS)
// with commit groups.  TMA = true (-DMB_RING=1): ONE lane issues two bulk copies (cp.async.bulk, the
// TMA engine: 256 B of weights, 64 or 128 B of codes) that complete on the slot's mbarrier -- no LSU
// instructions per lane, the staging writes do not go through the load/store pipe that the factor
// reads saturate.  Bulk copies need 16-byte aligned addresses and sizes, so chunks start on a multiple
// of 8 (16-bit codes) or 4 (32-bit codes) entries and are always copied whole: a range that starts
// in between skips up to 7 leading entries of its first chunk, and the entry arrays are padded at
// the end (kEntryPad).  Both are correct (the GPU tests pass with either); the bulk-copy ring measured
// 0.7 % slower on B200 (profiles/r2_ab_ring.txt), so it is not the default.
constexpr int kEntryPad = 64;   // entries allocated past the end of the entry arrays
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok && ++spins > (1u << 28)) __trap();  // a copy that never lands must not hang the GPU
  } while (!ok);
}

// One chunk of kChunk entries into ring slot `slot` of this warp's stage.
template <bool C16, bool TMA>
__device__ __forceinline__ void issue_chunk(const double* __restrict__ ga, const void* __restrict__ gcodes,
                                            int64_t p, const int64_t end, int slot, uint32_t st_a,
                                            uint32_t st_c, uint32_t bar0, int lane) {
  constexpr uint32_t kCodeBytes = C16 ? 2u : 4u;
  if (TMA) {
    if (lane == 0 && p < end) {
      const uint32_t bar = bar0 + 8u * slot;
      mbar_expect_tx(bar, kChunk * (8u + kCodeBytes));
      bulk_g2s(st_a + slot * (kChunk * 8u), ga + p, kChunk * 8u, bar);
      bulk_g2s(st_c + slot * (kChunk * kCodeBytes), static_cast<const char*>(gcodes) + p * kCodeBytes, kChunk * kCodeBytes, bar);
    }
  } else {
    const uint32_t sa_lane = st_a + lane * 8u, sc_lane = st_c + lane * 4u;
    if (p + lane < end) {
      cp_async8(sa_lane + slot * (kChunk * 8u), ga + p + lane);
      if (!C16) cp_async4(sc_lane + slot * (kChunk * 4u), static_cast<const uint32_t*>(gcodes) + p + lane);
    }
    if (C16 && lane < kChunk / 2 && p + 2 * lane < end)  // two codes per lane (the array is padded)
      cp_async4(sc_lane + slot * (kChunk * 2u), static_cast<const uint16_t*>(gcodes) + p + 2 * lane);
    cp_async_commit();
  }
}
template <bool C16, bool TMA>
__host__ __device__ constexpr int chunk_align() { return TMA ? (C16 ? 8 : 4) : (C16 ? 2 : 1); }

// The first STAGES - 1 chunks of a range, issued ahead of time (K2 does this for every warp's first
// segment before it builds the factor tables: the cold-HBM latency of the entry stream then hides
// behind the table build).  stream_range(..., primed = true) must follow with the same arguments.
template <int STAGES, int OB, bool TMA>
__device__ __forceinline__ void stream_prime(const double* __restrict__ ga, const void* __restrict__ gcodes,
                                             int64_t pos, const int64_t end, const uint32_t st_a,
                                             const uint32_t st_c, const uint32_t bar0, const int lane) {
  constexpr bool C16 = OB > 0;
  pos -= pos & (chunk_align<C16, TMA>() - 1);
#pragma unroll
  for (int k = 0; k < STAGES - 1; ++k)
    issue_chunk<C16, TMA>(ga, gcodes, pos + (int64_t)k * kChunk, end, k, st_a, st_c, bar0, lane);
}

// Stream entries [pos, end) through this warp's ring (weights at st_a, codes at st_c) and fold them
// into L.  Per group of four entries the warp issues three broadcast loads (four codes; two weight
// pairs) = 1.25-1.5 shared-memory wavefronts per entry, then per entry one IMAD (table row address),
// NW/2 x LDS.128 (factors), NW x DFMA and one sign test.  `par`: the ring's mbarrier phase bits (TMA).
template <int NW, int MODE, typename TF, int STAGES = kStages, int OB = 0, bool TMA = false>
__device__ __forceinline__ void stream_range(Lanes<NW, MODE, TF, OB>& L, const double* __restrict__ ga,
                                             const void* __restrict__ gcodes, int64_t pos, const int64_t end,
                                             const uint32_t st_a, const uint32_t st_c, const int lane,
                                             const bool primed = false, const uint32_t bar0 = 0, uint32_t* par = nullptr) {
  constexpr bool C16 = OB > 0;
  constexpr uint32_t kCodeBytes = C16 ? 2u : 4u;
  // chunks start on an aligned entry; a range that starts in between stages the entries before it too
  // and skips them: its first chunk handles entries one by one up to the next multiple of four
  int lead = (int)(pos & (chunk_align<C16, TMA>() - 1));
  pos -= lead;
  const uint32_t sa = keep_in_register(st_a), sc = keep_in_register(st_c);
  // ring of STAGES chunk buffers: STAGES - 1 chunks are in flight while one is processed
  auto issue = [&](int64_t p, int slot) { issue_chunk<C16, TMA>(ga, gcodes, p, end, slot, sa, sc, bar0, lane); };
  if (!primed) {
#pragma unroll
    for (int k = 0; k < STAGES - 1; ++k) issue(pos + (int64_t)k * kChunk, k);
  }
  int slot = 0;
  while (pos < end) {
    int ahead = slot + STAGES - 1;
    if (ahead >= STAGES) ahead -= STAGES;
    issue(pos + (int64_t)(STAGES - 1) * kChunk, ahead);
    if (TMA) {
      mbar_wait(bar0 + 8u * slot, (*par >> slot) & 1u);
      *par ^= 1u << slot;
    } else {
      cp_async_wait<STAGES - 1>();
      __syncwarp();
    }
    const uint32_t pa = sa + slot * (kChunk * 8u), pc = sc + slot * (kChunk * kCodeBytes);
    const int n = (int)min((int64_t)kChunk, end - pos);
    int i = lead;
    lead = 0;
    auto single = [&](int j) {
      const uint32_t c = C16 ? lds_stage_u16(pc + 2u * j) : lds_stage_u(pc + 4u * j);
      const double a = lds_stage_d(pa + 8u * j);
      double f[NW];
      L.load_factors(c, f);
      L.step(a, c, f);
    };
    for (; i < n && (i & 3); ++i) single(i);
    for (; i + 4 <= n; i += 4) {
      uint4 c;
      bool any_boundary;
      if (C16) {  // four codes in one 8-byte broadcast
        const uint2 cc = lds_stage_u2(pc + 2u * i);
        c = make_uint4(cc.x, cc.x >> 16, cc.y, cc.y >> 16);
        any_boundary = ((cc.x | cc.y) & 0x80008000u) != 0;
      } else {
        c = lds_stage_u4(pc + 4u * i);
        any_boundary = (int32_t)(c.x | c.y | c.z | c.w) < 0;
      }
      const double2 a01 = lds_stage_d2(pa + 8u * i), a23 = lds_stage_d2(pa + 8u * i + 16u);
      double f0[NW], f1[NW], f2[NW], f3[NW];
      L.load_factors(c.x, f0); L.load_factors(c.y, f1); L.load_factors(c.z, f2); L.load_factors(c.w, f3);
      // (the vote makes the uniformity of the branch visible to the compiler: no BSSY/BSYNC pair)
      if (__any_sync(0xffffffffu, any_boundary)) {  // a run / star boundary in the group
        L.step(a01.x, c.x, f0); L.step(a01.y, c.y, f1); L.step(a23.x, c.z, f2); L.step(a23.y, c.w, f3);
      } else {                                     // common case: four plain multiply-adds
        L.mac(a01.x, f0); L.mac2(a01.y, f1); L.mac(a23.x, f2); L.mac2(a23.y, f3);
      }
    }
    for (; i < n; ++i) single(i);
    __syncwarp();   // every lane is done with the slot before it is refilled
    if (++slot == STAGES) slot = 0;
    pos += kChunk;
  }
  if (!TMA) cp_async_wait<0>();
}

// Prologue: factor tables (built in place or copied), then this block's (age, Z) bins of the N-stars
// term.  Hot loop: every warp streams star-aligned segments of the entry stream for all walkers.
// Tail: reduction over stars (K3), Poisson term and, with peers, the reduction over GPUs.
template <int NW, int MODE, typename TF = double, int OB = 0>
__global__ void __launch_bounds__(k2_threads(NW), k2_min_blocks(NW)) mb_k2_stars(const __grid_constant__ K2Args a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int kK2Warps = k2_threads(NW) / 32;
  constexpr bool kF32 = sizeof(TF) == 4;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int Wpad = 32 * NW;  // == a.Wpad
  constexpr uint32_t rowbytes = Lanes<NW, MODE, TF, OB>::rowbytes;
  const uint32_t lanebytes = (NW == 1 ? lane : 2 * lane) * (MODE == kModeFactor ? (uint32_t)sizeof(TF) : 8u);

  // smem layout: [front: per-warp entry stage; also table-build / Poisson / reduction scratch]
  //              [outer | inner tables at tab_off (FACTOR only)]
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
  constexpr int STAGES = k2_stages(NW);
  const uint32_t st_a = smem0 + (uint32_t)warp * k2_stage_bytes(NW);
  const uint32_t st_c = st_a + STAGES * kChunk * 8u;
  const uint32_t tabA = smem0 + (uint32_t)a.tab_off + lanebytes;
  const uint32_t tabB = tabA + (uint32_t)a.nA * rowbytes;
  __shared__ unsigned int ticket;
  __shared__ int timed_out;
  __shared__ int bin_ctr;       // next of the block's N-stars bins (claimed by warps after their stars)
  if (threadIdx.x == 0) bin_ctr = 0;
  stamp(a.timeline, 0);
  // launched with programmatic stream serialization: wait here for K0 (the tables) to complete
  asm volatile("griddepcontrol.wait;" ::: "memory");
  // every warp's first segment: its first chunks are requested now, so that the cold-HBM latency of
  // the entry stream hides behind the table build (needs a stage that the build does not use as scratch)
  // (the segment bounds are requested first; the chunks themselves once the table build has issued
  // its own first loads, so that the two cold-memory latencies overlap instead of adding up)
  const int nwarps_total = (int)gridDim.x * kK2Warps;
  int s = (int)blockIdx.x * kK2Warps + warp;
  bool primed = false;
  int64_t seg_lo = 0, seg_hi = 0;
  const bool want_prime = a.separate && s < a.nseg;
  if (want_prime) { seg_lo = __ldg(a.seg + s); seg_hi = __ldg(a.seg + s + 1); }
  // the ring's mbarriers (bulk-copy mode): one per slot, armed and waited on by this warp alone
  __shared__ __align__(8) unsigned long long ring_bar[kK2Warps * kStages];
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&ring_bar[warp * kStages]);
  uint32_t ring_par = 0;
  // MB_RING: 0 = only the cp.async ring is compiled (default), 1 = only the bulk-copy (TMA) ring, 2 = both
  // (run-time switch MB_NO_TMA).  Measured on B200, same box, config 3 / its 1/8 shard (us per step):
  // 127.45 / 37.88 with 0, 128.3 / 38.0 with 1, 128.0-129.2 / 37.95-38.5 with 2 (profiles/r2_ab_ring.txt):
  // the bulk copies take ~1.5 % of the shared-memory wavefronts off the load/store pipe, but the
  // mbarrier try_wait per chunk costs more than cp.async.wait_group does once the data has landed.
#ifndef MB_RING
#define MB_RING 0
#endif
  const bool tma = MB_RING == 1 || (MB_RING == 2 && a.use_tma != 0);
  if (tma) {
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < STAGES; ++k) mbar_init(bar0 + 8u * k, 1u);
      asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncwarp();
  }
  auto prime_now = [&]() {
    if (want_prime && !primed) {
      if (tma) stream_prime<STAGES, OB, true>(a.ea, a.ec, seg_lo, seg_hi, st_a, st_c, bar0, lane);
      else stream_prime<STAGES, OB, false>(a.ea, a.ec, seg_lo, seg_hi, st_a, st_c, bar0, lane);
      primed = true;
    }
  };
  // N-stars term: the statistics of the block's first bins (b = blockIdx.x + k gridDim.x, slice k) on
  // their way to shared memory, so that the sums over the dust classes later run without global latency
  bool hts_pending = false;
  if (a.pois.nbins > 0 && !a.pois.SC && a.pois.hts_slices > 0) {
    const int nD = a.nOd * a.nB;
    for (int k = 0; k < a.pois.hts_slices; ++k) {
      const int b = (int)blockIdx.x + k * (int)gridDim.x;
      if (b >= a.pois.nbins) break;
      const double2* src = reinterpret_cast<const double2*>(a.pois.Ht) + (size_t)b * nD;
      const uint32_t dst = smem0 + (uint32_t)a.pois.hts_off + (uint32_t)k * (uint32_t)nD * 16u;
      for (int i = threadIdx.x; i < nD; i += blockDim.x)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + 16u * i), "l"(src + i));
    }
    cp_async_commit();
    hts_pending = true;
  }
  // (this copy group is the warp's oldest: waiting for it leaves the primed chunks in flight)
  auto hts_wait = [&]() {
    if (hts_pending) {
      if (primed && !tma) cp_async_wait<STAGES - 1>(); else cp_async_wait<0>();
      hts_pending = false;
    }
  };
  // this block's slices of the N-stars statistics: pull them into L2 while the tables are built
  if (a.pois.nbins > 0 && !a.pois.SC && a.pois.hts_slices == 0) {
    const int per_bin = a.nOd * a.nB * 16;  // bytes
    for (int b = blockIdx.x; b < a.pois.nbins; b += gridDim.x) {
      const char* base = reinterpret_cast<const char*>(a.pois.Ht) + (size_t)b * per_bin;
      for (int o = threadIdx.x * 128; o < per_bin; o += blockDim.x * 128)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + o));
    }
  }
  if (MODE == kModeFactor && !kF32 && a.build_tables) {
    // K0 fused: the prior models on the classes (singlepop_dust_model.py:146-148 restricted to the
    // distinct column values) for all walkers of the pass -> look-up rows (in the still idle stage
    // area, or directly in the tables for a parameter that is alone in its table), then the product
    // rows straight into shared memory.  The fp64 pipe bounds this prologue (every block repeats
    // it), so the lognormals are arranged for few operations: ln x and 1/x once per class,
    // mu = ln(mode) + sigma^2, 1/sigma and amp/(sigma sqrt(2 pi)) once per walker, then one
    // multiply-add chain and one branch-free exp per (class, walker) -- no divisions there.
    // (mb_k0_factors evaluates the textbook form; the two agree to an ulp or two.)
    const ModelDev& m = a.model;
    double* lut = reinterpret_cast<double*>(smem_raw + a.scratch_off);  // [nscratch][Wpad]
    double* tab = reinterpret_cast<double*>(smem_raw + a.tab_off);  // [nA + nB][Wpad]
    const int ncls = a.ncls_total, nrows = a.nA + a.nB;
    double* lxs = lut + (size_t)a.nscratch * Wpad;                  // [ncls][3] ln x, 1/x, x of the class value
    double* cw = lxs + 3 * ncls;                                    // [NPARAM][Wpad][6] per-walker lognormal constants
    ushort4* srow = reinterpret_cast<ushort4*>(cw + 6 * MB_NPARAM * Wpad);  // [nrows] look-up rows of a table row
    unsigned short* sslot = reinterpret_cast<unsigned short*>(srow + nrows);  // [ncls] (padded to 8 bytes)
    // phi: carried in the kernel parameters (copied to shared memory with indexed constant loads --
    // reads of the parameter space through a generic pointer are slow) or in device memory
    // The later phases read it from shared memory in both cases: a first touch of device memory there
    // would be a cold miss queued behind the entry chunks requested in between.
    const double* ph = a.phi;
    double* sphi = reinterpret_cast<double*>(sslot + ((ncls + 3) & ~3));
    {
      const int n = a.W * a.ndim;
      if (a.phi_inline) {
        for (int i = threadIdx.x; i < n; i += blockDim.x) sphi[i] = a.phi_in[i];
        __syncthreads();
        ph = sphi;
      } else {
        for (int i = threadIdx.x; i < n; i += blockDim.x) sphi[i] = __ldg(a.phi + i);  // visible after the next barrier
      }
    }
    const double inv_sqrt_2pi = 0.3989422804014326779399460599343819;
    // (the row lists ride along with the first phase: their L2 latency hides behind the logarithms)
    for (int i = (int)blockDim.x - 1 - (int)threadIdx.x; i < nrows; i += blockDim.x)
      srow[i] = __ldg(reinterpret_cast<const ushort4*>(a.rowsrc) + i);
    for (int i = (int)blockDim.x - 1 - (int)threadIdx.x; i < ncls; i += blockDim.x)
      sslot[i] = __ldg(a.lutslot + i);
    for (int idx = threadIdx.x; idx < ncls + MB_NPARAM * Wpad; idx += blockDim.x) {
      if (idx < ncls) {
        const double x = __ldg(a.clsx + idx);
        lxs[3 * idx] = x > 0.0 ? log(x) : 0.0;
        lxs[3 * idx + 1] = x > 0.0 ? 1.0 / x : 0.0;  // x <= 0: the lognormal is exactly 0 there
        lxs[3 * idx + 2] = x;
      } else {
        const int j = idx - ncls, p = j / Wpad, w = j - p * Wpad;
        const double* th = ph + (size_t)(w < a.W ? w : a.W - 1) * a.ndim + m.voff[p];
        double* c = cw + 6 * j;
        if (m.kind[p] == MB_LOGNORMAL) {
          const double isig = 1.0 / th[1];
          c[0] = log(th[0]) + th[1] * th[1]; c[1] = isig; c[2] = inv_sqrt_2pi * isig;
        } else if (m.kind[p] == MB_TWO_LOGNORMAL) {
          const double n1 = 1.0 - 1.0 / (th[4] + 1.0), n2 = 1.0 / (th[4] + 1.0);
          const double isig1 = 1.0 / th[2], isig2 = 1.0 / th[3];
          c[0] = log(th[0]) + th[2] * th[2]; c[1] = isig1; c[2] = n1 * inv_sqrt_2pi * isig1;
          c[3] = log(th[1]) + th[3] * th[3]; c[4] = isig2; c[5] = n2 * inv_sqrt_2pi * isig2;
        }
      }
    }
    prime_now();
    __syncthreads();
    ph = sphi;
    {
      ExpCoef ec;
#pragma unroll
      for (int i = 0; i < 14; ++i) ec.c[i] = kExpPoly[i];
      const int nlut = ncls * Wpad;
#pragma unroll 3
      for (int idx = threadIdx.x; idx < nlut; idx += blockDim.x) {
        const int i = idx / Wpad, w = idx - i * Wpad;
        int p = 0;
#pragma unroll
        for (int q = 1; q < MB_NPARAM; ++q)
          if (i >= m.clsoff[q]) p = q;
        const double* c = cw + 6 * (p * Wpad + w);
        const double lx = lxs[3 * i], rx = lxs[3 * i + 1];
        double v = 1.0;
        if (m.kind[p] == MB_LOGNORMAL) {
          const double t = (lx - c[0]) * c[1];
          v = (c[2] * rx) * exp_nonpos(-0.5 * (t * t), ec);
        } else if (m.kind[p] == MB_TWO_LOGNORMAL) {
          const double t1 = (lx - c[0]) * c[1], t2 = (lx - c[3]) * c[4];
          v = (c[2] * rx) * exp_nonpos(-0.5 * (t1 * t1), ec) + (c[5] * rx) * exp_nonpos(-0.5 * (t2 * t2), ec);
        } else if (m.kind[p] != MB_FIXED) {
          // padding columns repeat the last walker
          v = prior_eval(m, p, ph + (size_t)(w < a.W ? w : a.W - 1) * a.ndim + m.voff[p], lxs[3 * i + 2]);
        }
        const unsigned slot = sslot[i];
        ((slot & 0x8000u) ? tab + (size_t)(slot & 0x7fffu) * Wpad : lut + (size_t)slot * Wpad)[w] = v;
      }
    }
    __syncthreads();
    const int ntab = nrows * Wpad;
#pragma unroll 3
    for (int idx = threadIdx.x; idx < ntab; idx += blockDim.x) {
      const int r = idx / Wpad, w = idx - r * Wpad;
      // look-up rows of the table row (precomputed: no integer divisions here), multiplied in the
      // reference's left-to-right order (`cur_physmod *=`, :146-148)
      const ushort4 sr = srow[r];
      if (sr.x == 0xfffe) continue;  // evaluated in place
      double v = 1.0;
      auto lrow = [&](unsigned slot) -> const double* {
        return (slot & 0x8000u) ? tab + (size_t)(slot & 0x7fffu) * Wpad : lut + (size_t)slot * Wpad;
      };
      if (sr.x != 0xffff) v *= lrow(sr.x)[w];
      if (sr.y != 0xffff) v *= lrow(sr.y)[w];
      if (sr.z != 0xffff) v *= lrow(sr.z)[w];
      if (sr.w != 0xffff) v *= lrow(sr.w)[w];
      tab[idx] = v;
    }
    hts_wait();
    __syncthreads();
  } else if (MODE == kModeFactor) {
    const int n2 = (a.nA + a.nB) * Wpad / 2;
    const double2* src = reinterpret_cast<const double2*>(a.FA);
    if (kF32) {
      uint2* dst = reinterpret_cast<uint2*>(smem_raw + a.tab_off);
      for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        const double2 v = __ldg(src + i);
        dst[i] = make_uint2(round_to_hi_word(v.x), round_to_hi_word(v.y));
      }
    } else {
      double2* dst = reinterpret_cast<double2*>(smem_raw + a.tab_off);
      for (int i = threadIdx.x; i < n2; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    prime_now();
    hts_wait();
    __syncthreads();
  } else {
    prime_now();
    if (a.pois.nbins > 0 && (a.pois.warp_private || a.pois.hts_slices > 0)) { hts_wait(); __syncthreads(); }  // (also publishes bin_ctr)
  }
  stamp(a.timeline, 1);
#ifdef MB_WARP_STAMPS
  unsigned long long* wst = a.timeline ? a.timeline + (size_t)blockIdx.x * kTimelinePoints + 8 + warp * kWarpStampPoints : nullptr;
  auto wstamp = [&](int k) {
    if (wst && lane == 0 && warp < kWarpStampWarps) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
      wst[k] = t;
    }
  };
  wstamp(0);
#endif
  // N-stars prologue: this block's (age, Z) bins for all walkers of the pass, on the tables it just
  // staged (always the fp64 ones: from shared memory when they are there, else from global memory)
  const double* Ftab = (MODE == kModeFactor && !kF32) ? reinterpret_cast<const double*>(smem_raw + a.tab_off) : a.FA;
  if (a.pois.nbins > 0 && !a.pois.warp_private)
    poisson_bins<NW, k2_threads(NW)>(a.pois, Ftab, a.nA, a.nOd, a.nB, smem_raw, a.separate != 0);
  if (a.pois.nbins > 0 && a.pois.warp_private && a.pois.bins_first)
    poisson_bins_private<NW, k2_threads(NW)>(a.pois, Ftab, a.nA, a.nOd, a.nB, smem_raw, &bin_ctr);

  stamp(a.timeline, 2);
#ifdef MB_WARP_STAMPS
  wstamp(1);
  if (wst && lane == 0 && warp < kWarpStampWarps)
    wst[3] = a.timeline[(size_t)blockIdx.x * kTimelinePoints] + 1000ull * (s < a.nseg ? (unsigned long long)(a.seg[s + 1] - a.seg[s]) : 0ull);
#endif
  Lanes<NW, MODE, TF, OB> L;
  L.init(tabA, tabB, reinterpret_cast<const char*>(a.T) + lanebytes);

  // Segments: every warp starts with its own index (no atomic storm at kernel start); further
  // ones are claimed from a shared counter, one atomic per segment issued a segment ahead.  The
  // result does not depend on who processed what: a segment's value depends only on its own
  // entries, and segment values are added as fixed point.
  long long fixhi[NW], fixlo[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) { fixhi[i] = 0; fixlo[i] = 0; }
  // (ptxas turns the claim -- written as atomicAdd() or as a PTX atom -- into its warp-aggregated form, leader
  // atomic + SHFL of the result, so the warp WAITS for the atomic's round trip where it is issued: ~1.3 us per
  // warp at the head of the loop with 3,552 warps hitting one address (ncu source page,
  // profiles/r2_k2_source_hotspots.txt).  With one segment per warp -- the default -- there is nothing to
  // claim, and the atomic is skipped.)
  const bool more_segments = a.nseg > nwarps_total;
  int s_next = 0;
  while (s < a.nseg) {
    if (more_segments && lane == 0) s_next = nwarps_total + (int)atomicAdd(a.counters + kCtrSegment, 1u);
    if (tma) stream_range<NW, MODE, TF, STAGES, OB, true>(L, a.ea, a.ec, a.seg[s], a.seg[s + 1], st_a, st_c, lane, primed, bar0, &ring_par);
    else stream_range<NW, MODE, TF, STAGES, OB, false>(L, a.ea, a.ec, a.seg[s], a.seg[s + 1], st_a, st_c, lane, primed);
    primed = false;
    L.close();
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const Fix2 f = to_fix2(L.acc[i].take_value());
      fixhi[i] += f.hi; fixlo[i] += f.lo;
    }
    s = more_segments ? __shfl_sync(0xffffffffu, s_next, 0) : a.nseg;
  }
  // Few dust classes: the block's (age, Z) bins are taken by whichever warps finish their stars first --
  // warps wait here for the block's slowest one anyway (segments end on star boundaries), so the
  // N-stars term costs the critical path nothing.  The tables are still intact (the reduction slabs
  // move over them only after the barrier below).
#ifdef MB_WARP_STAMPS
  wstamp(2);
#endif
  if (a.pois.nbins > 0 && a.pois.warp_private && !a.pois.bins_first)
    poisson_bins_private<NW, k2_threads(NW)>(a.pois, Ftab, a.nA, a.nOd, a.nB, smem_raw, &bin_ctr);

  // ---- K3 (fused): warp -> block (slabs) -> grid (global atomics) reduction over stars,
  // all in integers; the last block to finish converts and writes the result rows.
  __syncthreads();
  stamp(a.timeline, 3);
  // every warp parks its per-walker integers in its own slab of the (now idle) stage, then
  // kAccRows * Wpad threads add the slabs: no shared-memory atomics (64-bit ones are CAS loops)
  // (in the front region, or -- when that is too small -- over the factor tables, dead by now)
  unsigned long long* slab = reinterpret_cast<unsigned long long*>(smem_raw + a.slab_off);  // [warps][kAccRows][Wpad]
#pragma unroll
  for (int i = 0; i < NW; ++i) {
    const int w = (NW == 1) ? lane : 64 * (i >> 1) + 2 * lane + (i & 1);
    slab[((size_t)warp * kAccRows + 0) * Wpad + w] = (unsigned long long)fixhi[i];
    slab[((size_t)warp * kAccRows + 1) * Wpad + w] = (unsigned long long)fixlo[i];
    slab[((size_t)warp * kAccRows + 2) * Wpad + w] = (unsigned long long)L.acc[i].bad;
    slab[((size_t)warp * kAccRows + 3) * Wpad + w] = (unsigned long long)L.acc[i].isnan;
  }
  if (threadIdx.x == 0) timed_out = 0;
  __syncthreads();
  unsigned long long mine = 0ull;
  if (threadIdx.x < kAccRows * Wpad) {
    for (int wp = 0; wp < kK2Warps; ++wp) mine += slab[(size_t)wp * kAccRows * Wpad + threadIdx.x];
    if (mine != 0ull) atomicAdd(a.gacc + threadIdx.x, mine);
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) ticket = atomicAdd(a.counters + kCtrBlocks, 1u);
  __syncthreads();
  stamp(a.timeline, 4);
  if (ticket != gridDim.x - 1) return;
  __threadfence();
  // last block of this GPU: take the shard totals, leave the accumulators zero for the next launch.
  // (the loads of the totals and of the per-bin N-stars results are issued together)
  unsigned long long* tot = slab;  // [kAccRows][Wpad]; the Poisson scratch follows it
  unsigned long long mytot = 0ull;
  if (threadIdx.x < kAccRows * Wpad) mytot = __ldcg(a.gacc + threadIdx.x);
  const PoisPart ppart = poisson_finish_load<NW, k2_threads(NW)>(a.pois);
  if (threadIdx.x < kAccRows * Wpad) {
    tot[threadIdx.x] = mytot;
    a.gacc[threadIdx.x] = 0ull;
  }
  if (threadIdx.x == 0) { a.counters[kCtrSegment] = 0u; a.counters[kCtrBlocks] = 0u; }
  const bool peers = a.peer.world > 1;
  const int par = (int)(a.peer.seq & 1u), me = a.peer.rank, world = a.peer.world;
  __syncthreads();   // tot[] complete
  if (peers) {
    // my totals to every other rank, each 64-bit value as two self-flagged words (see above)
    for (int j = threadIdx.x; j < 2 * kAccRows * Wpad; j += blockDim.x) {
      const unsigned long long v = tot[j >> 1];
      const unsigned long long word = ((unsigned long long)a.peer.seq << 32) | (unsigned long long)(uint32_t)(v >> (32 * (j & 1)));
      for (int p = 0; p < world; ++p)
        if (p != me)
          asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(a.peer.xbuf[p] + xbuf_word_off(par, me) + j), "l"(word) : "memory");
    }
  }
  // Poisson term of every walker: the bins of all blocks, added in fixed order (identical on every
  // rank); with peers this hides behind the NVLink flight of the words just sent
  poisson_finish<NW, k2_threads(NW)>(a.pois, ppart, reinterpret_cast<double*>(tot + kAccRows * Wpad),
                                     a.out + (size_t)MB_OUT_POISSON * a.Wtot + a.w0, a.W);
  stamp(a.timeline, 5);
  if (peers) {
    // thread i adds value i of every other source as it arrives (bounded wait)
    if (threadIdx.x < kAccRows * Wpad) {
      unsigned long long sum = tot[threadIdx.x];
      unsigned long long t0 = 0ull;
      for (int r = 0; r < world; ++r) {
        if (r == me) continue;
        const unsigned long long* src = a.peer.xbuf[me] + xbuf_word_off(par, r) + 2 * threadIdx.x;
        unsigned long long w0, w1;
        unsigned int spins = 0;
        while (true) {
          asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
          if ((uint32_t)(w0 >> 32) == a.peer.seq && (uint32_t)(w1 >> 32) == a.peer.seq) break;
          if ((++spins & 0x3ffu) == 0u) {
            unsigned long long t1;
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            if (t0 == 0ull) t0 = t1;
            if (t1 - t0 > a.peer.timeout_ns) { timed_out = 1; break; }
          }
        }
        sum += (w0 & 0xffffffffull) | (w1 << 32);
      }
      tot[threadIdx.x] = sum;
    }
    __syncthreads();
  }
  stamp(a.timeline, 6);
  for (int w = threadIdx.x; w < a.W; w += blockDim.x) {
    const unsigned long long nbad = tot[2 * Wpad + w], nnan = tot[3 * Wpad + w];
    // each word converts exactly (|word| < 2^53): the caller's STARSUM + STARSUM_LO is the only rounding
    const bool bad_sum = nnan || timed_out;
    a.out[(size_t)MB_OUT_STARSUM * a.Wtot + a.w0 + w] = bad_sum ? nan("") : fix_hi_to_double((long long)tot[w], (long long)tot[Wpad + w]);
    a.out[(size_t)MB_OUT_STARSUM_LO * a.Wtot + a.w0 + w] = bad_sum ? 0.0 : fix_lo_to_double((long long)tot[Wpad + w]);
    a.out[(size_t)MB_OUT_NONFINITE * a.Wtot + a.w0 + w] = timed_out ? -1.0 : (double)nbad;
  }
  if (a.done_flag) {
    // host path: the result rows live in mapped host memory; publish "rows complete" there so the
    // host can pick them up without waiting for the stream's completion machinery
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *a.done_flag = a.done_seq;
  }
  stamp(a.timeline, 7);
}

// ---- pixel batches: many independent small ensembles over one shared grid -----------------------
// SURVEY.md section 8(f) "next #3" (origin: old/megabeast_fit.py:97-182, a serial loop over sky
// pixels with one fit each).  Every pixel has its own stars and its own W walkers; all pixels of
// all walkers are evaluated in ONE launch.  A block owns one pixel at a time (pixels claimed
// dynamically): it stages that pixel's factor tables, computes the N-stars / Poisson terms of
// all W walkers at once, streams the pixel's entries (12 star-aligned pieces, one per warp) and
// writes the pixel's result rows -- no cross-block reduction.
constexpr int kPixThreads = 384;
constexpr int kPixWarps = kPixThreads / 32;
struct KPixArgs {
  const double* ea;           // entry weights
  const uint32_t* ec;         // entry codes (FACTOR layout)
  const int64_t* pix_seg;     // [P][kPixWarps + 1] entry offsets of the pixel's warp pieces
  const double* F;            // [P][nA + nB + nA][Wpad] (K0's layout: outer = age rows, inner = dust rows, plain age rows)
  const int32_t* pix_nstars;  // [P] S of each pixel's Poisson term
  double* out;                // [P][MB_OUT_ROWS][W]
  unsigned int* counters;     // [kCtrWords]: next pixel, blocks done (zero between launches)
  PoissonArgs pois;           // shared statistics; n_stars is per pixel
  int32_t P, nA, nB, Wpad, W, tab_off;
};

// Poisson terms of ALL walkers of one ensemble: tasks = (group of 8 walkers) x (chunk of 32 bins),
// lanes = bins (Ht read coalesced), the 8 walkers' factors are warp-uniform broadcasts.  One
// pass: pred = sum over bins with S > 0 of coef * FA * C / S; the grand total only decides
// the all-zero case (helpers.py:133-134,156: 0/0 weights skip every bin -> pred = 0).
template <int THREADS>
__device__ void poisson_all_walkers(const PoissonArgs& pa, const double* sF, int nA, int nB, int Wpad,
                                    double n_stars, double* red /* [tasks][16] */, double* spois /* [Wpad] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = THREADS / 32;
  const int nbins = pa.nbins, ngroups = Wpad / 8, nchunks = (nbins + 31) / 32;
  const double2* Ht = reinterpret_cast<const double2*>(pa.Ht);
  const double* sFB = sF + (size_t)nA * Wpad;
  for (int task = warp; task < ngroups * nchunks; task += nwarps) {
    const int wg = task / nchunks, b = (task - wg * nchunks) * 32 + lane;
    double S[8], C[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { S[k] = 0.0; C[k] = 0.0; }
    if (b < nbins) {
      for (int cb = 0; cb < nB; ++cb) {
        const double2 h = __ldg(Ht + (size_t)cb * nbins + b);
        const double2* f2 = reinterpret_cast<const double2*>(sFB + (size_t)cb * Wpad + wg * 8);
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          const double2 f = f2[k2];
          S[2 * k2] = fma(h.x, f.x, S[2 * k2]);         C[2 * k2] = fma(h.y, f.x, C[2 * k2]);
          S[2 * k2 + 1] = fma(h.x, f.y, S[2 * k2 + 1]); C[2 * k2 + 1] = fma(h.y, f.y, C[2 * k2 + 1]);
        }
      }
    }
    double v[16];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double nrm = 0.0, pred = 0.0;
      if (b < nbins) {
        const double fa = sF[(size_t)pa.bin_agecls[b] * Wpad + wg * 8 + k];
        const double s = fa * S[k], c = fa * C[k];
        nrm = s;
        if (s > 0.0) pred = (fa * pa.coef[b]) * c / s;
      }
      v[k] = nrm; v[8 + k] = pred;
    }
#pragma unroll
    for (int k = 0; k < 16; ++k)
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0)
#pragma unroll
      for (int k = 0; k < 16; ++k) red[(size_t)task * 16 + k] = v[k];
  }
  __syncthreads();
  for (int w = threadIdx.x; w < Wpad; w += THREADS) {
    const int wg = w / 8, k = w % 8;
    double nrm = 0.0, pred = 0.0;
    for (int ch = 0; ch < nchunks; ++ch) {  // fixed chunk order
      nrm += red[(size_t)(wg * nchunks + ch) * 16 + k];
      pred += red[(size_t)(wg * nchunks + ch) * 16 + 8 + k];
    }
    if (!(nrm > 0.0) && !(nrm != nrm)) pred = 0.0;  // all weights zero: every bin skipped
    spois[w] = n_stars * log(pred) - pred - lgamma(n_stars + 1.0);  // :182-186
  }
  __syncthreads();
}

template <int NW>
__global__ void __launch_bounds__(kPixThreads, (NW <= 2 ? 2 : 1)) mb_k2_pixels(KPixArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int Wpad = 32 * NW;  // == a.Wpad
  constexpr uint32_t rowbytes = (uint32_t)Wpad * 8u;
  const uint32_t lanebytes = (NW == 1 ? lane : 2 * lane) * 8u;
  // smem: [front: entry stage / Poisson task scratch][tables][spois Wpad][sacc kAccRows x Wpad]
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
  const uint32_t st_a = smem0 + (uint32_t)warp * kStageBytes;
  const uint32_t st_c = st_a + kStages * kChunk * 8u;
  double* sF = reinterpret_cast<double*>(smem_raw + a.tab_off);
  double* spois = sF + (size_t)(a.nA + a.nB) * Wpad;
  unsigned long long* sacc = reinterpret_cast<unsigned long long*>(spois + Wpad);
  const uint32_t tabA = smem0 + (uint32_t)a.tab_off + lanebytes;
  const uint32_t tabB = tabA + (uint32_t)a.nA * rowbytes;
  __shared__ int cur_pixel;

  bool first = true;
  while (true) {
    if (threadIdx.x == 0)
      cur_pixel = first ? (int)blockIdx.x : (int)gridDim.x + (int)atomicAdd(a.counters + kCtrSegment, 1u);
    first = false;
    __syncthreads();
    const int pix = cur_pixel;
    if (pix >= a.P) break;
    {  // this pixel's factor tables
      const int n2 = (a.nA + a.nB) * Wpad / 2;
      const double2* src = reinterpret_cast<const double2*>(a.F + (size_t)pix * (a.nA + a.nB + a.nA) * Wpad);
      double2* dst = reinterpret_cast<double2*>(sF);
      for (int i = threadIdx.x; i < n2; i += kPixThreads) dst[i] = __ldg(src + i);
      for (int i = threadIdx.x; i < kAccRows * Wpad; i += kPixThreads) sacc[i] = 0ull;
    }
    __syncthreads();
    const double n_stars = (double)a.pix_nstars[pix];
    if (a.pois.nbins > 0)
      poisson_all_walkers<kPixThreads>(a.pois, sF, a.nA, a.nB, Wpad, n_stars,
                                       reinterpret_cast<double*>(smem_raw), spois);
    else {
      for (int w = threadIdx.x; w < Wpad; w += kPixThreads) spois[w] = 0.0;
      __syncthreads();
    }

    // this warp's piece of the pixel's entry stream (same inner loop as mb_k2_stars)
    Lanes<NW, kModeFactor, double> L;
    L.init(tabA, tabB, nullptr);
    stream_range<NW, kModeFactor, double>(L, a.ea, a.ec, a.pix_seg[(size_t)pix * (kPixWarps + 1) + warp],
                                  a.pix_seg[(size_t)pix * (kPixWarps + 1) + warp + 1], st_a, st_c, lane);
    L.close();
#pragma unroll
    for (int i = 0; i < NW; ++i) {
      const int w = (NW == 1) ? lane : 64 * (i >> 1) + 2 * lane + (i & 1);
      const Fix2 f = to_fix2(L.acc[i].take_value());
      if (f.hi) atomicAdd(sacc + w, (unsigned long long)f.hi);
      if (f.lo) atomicAdd(sacc + Wpad + w, (unsigned long long)f.lo);
      if (L.acc[i].bad) atomicAdd(sacc + 2 * Wpad + w, (unsigned long long)L.acc[i].bad);
      if (L.acc[i].isnan) atomicAdd(sacc + 3 * Wpad + w, 1ull);
    }
    __syncthreads();
    double* out = a.out + (size_t)pix * MB_OUT_ROWS * a.W;
    for (int w = threadIdx.x; w < a.W; w += kPixThreads) {
      out[(size_t)MB_OUT_STARSUM * a.W + w] = sacc[3 * Wpad + w] ? nan("") : fix_hi_to_double((long long)sacc[w], (long long)sacc[Wpad + w]);
      out[(size_t)MB_OUT_STARSUM_LO * a.W + w] = sacc[3 * Wpad + w] ? 0.0 : fix_lo_to_double((long long)sacc[Wpad + w]);
      out[(size_t)MB_OUT_NONFINITE * a.W + w] = (double)sacc[2 * Wpad + w];
      out[(size_t)MB_OUT_POISSON * a.W + w] = spois[w];
    }
    __syncthreads();  // tables, sacc and the stage are reused by the next pixel
  }
  // the last block out re-arms the counters for the next launch
  __shared__ unsigned int ticket;
  if (threadIdx.x == 0) ticket = atomicAdd(a.counters + kCtrBlocks, 1u);
  __syncthreads();
  if (ticket == gridDim.x - 1 && threadIdx.x == 0) { a.counters[kCtrSegment] = 0u; a.counters[kCtrBlocks] = 0u; }
}

// ---- one-time entry build on the device (SURVEY.md 8(f) "next #1") -------------------------------------
// What mb_create does with a star's sparse points, for all stars at once: mask the non-finite values
// (singlepop_dust_model.py:196 -- their indices are never dereferenced), fold the three static gathers into
// one weight vals*prior_weight[r]*grid_weight[r]*completeness[r] (:207-213, same multiplication order),
// look the class code of row r up, sort the star's entries by (outer class, table row) so that equal
// outer classes form runs, pre-sum entries that share all class codes (in the order of their sparse
// points: deterministic), flag run and star starts.  One warp per star; the star's points are sorted
// by a bitonic network in shared memory (N = 32, 64 or 128 slots).  Entries land in a padded
// per-star temporary; mb_compact_entries packs them once the host has prefix-summed the counts.
struct BuildArgs {
  const int64_t* indxs;      // [K][S] this handle's columns of the (n_lnps, n_stars) arrays
  const double* vals;
  int64_t K, S, G;
  const uint32_t* rowcode;   // [G] outer class << 20 | inner class; null: no class codes (by-row modes on any grid)
  const double* pw;          // prior_weight, grid_weight, completeness [G]
  const double* gw;
  const double* cm;
  int32_t mode;              // MB_MODE_*: what the code word's low bits hold
  int32_t nB;                // TABLE_UNIQUE: row = age class * nB + dust class
  int32_t merge;
  double* tmp_a;             // [S][K] weights
  uint32_t* tmp_c;           // [S][K] 32-bit code words
  int32_t* cnt;              // [S] entries kept, [S] finite points, [S] runs
  int32_t* nraw;
  int32_t* nrun;
  int32_t* err_star;         // smallest star with an index out of bounds (INT_MAX: none)
};
constexpr int kBuildWarps = 4;

template <int N>
__global__ void __launch_bounds__(32 * kBuildWarps) mb_build_entries(const BuildArgs a) {
  __shared__ unsigned long long skey[kBuildWarps][N];
  __shared__ double sw[kBuildWarps][N];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long* key = skey[warp];
  double* wt = sw[warp];
  for (int64_t s = (int64_t)blockIdx.x * kBuildWarps + warp; s < a.S; s += (int64_t)gridDim.x * kBuildWarps) {
    int nfin = 0;
    for (int k = lane; k < N; k += 32) {
      unsigned long long kv = ~0ull;  // padding and masked points sort last
      double w = 0.0;
      if (k < a.K) {
        const double v = a.vals[(size_t)k * a.S + s];
        if (isfinite(v)) {
          int64_t r = a.indxs[(size_t)k * a.S + s];
          if (r < 0) r += a.G;  // numpy negative indexing
          if (r < 0 || r >= a.G) {
            atomicMin(a.err_star, (int)s);
          } else {
            w = v * a.pw[r] * a.gw[r] * a.cm[r];
            const uint32_t rc = a.rowcode ? a.rowcode[r] : 0u;
            const uint32_t cO = (rc >> kAgeShift) & kAgeMask, cI = rc & kDustMask;
            const uint32_t idx = a.mode == MB_MODE_FACTOR ? cI : a.mode == MB_MODE_TABLE_UNIQUE ? cO * (uint32_t)a.nB + cI : (uint32_t)r;
            kv = ((unsigned long long)cO << 40) | ((unsigned long long)idx << 8) | (unsigned long long)k;
            ++nfin;
          }
        }
      }
      key[k] = kv; wt[k] = w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) nfin += __shfl_xor_sync(0xffffffffu, nfin, o);
    __syncwarp();
    // bitonic sort of (key, weight) pairs, ascending; keys are distinct (they end in the point number)
    for (int k = 2; k <= N; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int t = lane; t < N / 2; t += 32) {
          const int i = ((t / j) * 2 * j) + (t % j), l = i + j;
          const bool up = (i & k) == 0;
          const unsigned long long ki = key[i], kl = key[l];
          if ((ki > kl) == up) {
            key[i] = kl; key[l] = ki;
            const double wi = wt[i]; wt[i] = wt[l]; wt[l] = wi;
          }
        }
        __syncwarp();
      }
    // heads of groups of equal class codes (all entries are heads when duplicates are not merged)
    int base = 0, runs = 0;
    for (int i0 = 0; i0 < nfin; i0 += 32) {
      const int i = i0 + lane;
      bool head = false, newrun = false;
      double wsum = 0.0;
      unsigned long long c = 0ull;
      if (i < nfin) {
        c = key[i] >> 8;
        const unsigned long long cp = i > 0 ? key[i - 1] >> 8 : ~0ull;
        head = !a.merge || i == 0 || c != cp;
        newrun = i == 0 || (c >> 32) != (cp >> 32);
        if (head) {
          wsum = wt[i];
          if (a.merge)
            for (int j = i + 1; j < nfin && (key[j] >> 8) == c; ++j) wsum += wt[j];  // in point order
        }
      }
      const unsigned hm = __ballot_sync(0xffffffffu, head);
      if (head) {
        const int pos = base + __popc(hm & ((1u << lane) - 1u));
        const uint32_t cO = (uint32_t)(c >> 32), idx = (uint32_t)c;
        uint32_t code;
        if (a.mode == MB_MODE_FACTOR) code = (cO << kAgeShift) | idx | (i == 0 ? (kNewRun | kNewStar) : newrun ? kNewRun : 0u);
        else code = idx | (i == 0 ? kNewRun : 0u);
        a.tmp_a[(size_t)s * a.K + pos] = wsum;
        a.tmp_c[(size_t)s * a.K + pos] = code;
      }
      base += __popc(hm);
      runs += __popc(__ballot_sync(0xffffffffu, head && newrun));
    }
    if (lane == 0) { a.cnt[s] = base; a.nraw[s] = nfin; a.nrun[s] = a.mode == MB_MODE_FACTOR ? runs : 0; }
    __syncwarp();
  }
}

// pack the per-star temporaries into the entry stream: offs[s] = entries before star s
__global__ void __launch_bounds__(256) mb_compact_entries(const double* __restrict__ tmp_a, const uint32_t* __restrict__ tmp_c,
                                                          const int32_t* __restrict__ cnt, const int64_t* __restrict__ offs,
                                                          int64_t S, int64_t K, double* __restrict__ ea,
                                                          uint32_t* __restrict__ ec, uint16_t* __restrict__ ec16, int ob) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t s = wid; s < S; s += nw) {
    const int n = cnt[s];
    const int64_t off = offs[s];
    for (int i = lane; i < n; i += 32) {
      ea[off + i] = tmp_a[(size_t)s * K + i];
      const uint32_t c = tmp_c[(size_t)s * K + i];
      if (ec16)
        ec16[off + i] = (uint16_t)(((c & kNewRun) ? 0x8000u : 0u) | ((c & kNewStar) ? 0x4000u : 0u) |
                                   (((c >> kAgeShift) & kAgeMask) << (14 - ob)) | (c & kDustMask));
      else ec[off + i] = c;
    }
  }
}

// ---- one-time ingest: helpers.py:33-41 ---------------------------------------------------------
__global__ void mb_ingest_lnp(int64_t n, const int64_t* __restrict__ indxs, double* __restrict__ vals,
                              int64_t G, const double* __restrict__ pw, int* __restrict__ err) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x) {
    double v = vals[i];
    if (isfinite(v)) {
      int64_t r = indxs[i];
      if (r < 0) r += G;  // numpy negative indexing
      if (r < 0 || r >= G) { atomicExch(err, 1); continue; }
      vals[i] = exp(v) / pw[r];
    }
  }
}

}  // namespace mb
