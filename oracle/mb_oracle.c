/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not product code; never linked into
 * libmegabeast_b200.so.  Used by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
 *
 * Plain-C restatement of the reference's ensemble likelihood for a batch of walkers:
 *
 *   /root/reference/megabeast/singlepop_dust_model.py:131-148   prior weights over the grid
 *   /root/reference/megabeast/helpers.py:133-164                expected number of stars
 *   /root/reference/megabeast/singlepop_dust_model.py:182-186   Poisson term
 *   /root/reference/megabeast/singlepop_dust_model.py:194-223   per-star gather, sum, log
 *
 * It deliberately does what the reference does -- materialise physmod[G] for each walker,
 * accumulate the (age, Z) bin sums over every grid row, gather physmod and the three static
 * columns at each star's finite entries -- and shares no code or data layout with the CUDA
 * path, so agreement between the two is evidence about both.  Differences from the numpy
 * text are confined to summation order (sequential here, pairwise in numpy): <= 1e-13
 * relative, checked against oracle/lnprob_port.py and the golden vectors in
 * tests/test_oracle.py.
 *
 * The prior-model formulas restate the `beast` package (absent from the container and the
 * reference tree) exactly as oracle/priors.py does: parity UNPINNED for that part.
 *
 * Parallelism: OpenMP over walkers (each thread owns a physmod scratch array).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { MBO_FIXED = 0, MBO_FLAT = 1, MBO_LOGNORMAL = 2, MBO_TWO_LOGNORMAL = 3,
       MBO_BINS_HISTO = 4, MBO_BINS_INTERP = 5, MBO_EXPONENTIAL = 6 };

#define MBO_NPARAM 4      /* logA, Av, Rv, fA -- M_ini can never be free in the reference */
#define MBO_MAX_NODES 64

typedef struct {
  int32_t kind[MBO_NPARAM];          /* MBO_FIXED == contributes no factor (:135) */
  int32_t nnodes[MBO_NPARAM];        /* bins_histo / bins_interp */
  double nodes[MBO_NPARAM][MBO_MAX_NODES];
  double fixed_tail[MBO_NPARAM];     /* unused; keeps the struct layout obvious */
} mbo_model;

static int nvars_of(const mbo_model* m, int p) {
  switch (m->kind[p]) {
    case MBO_LOGNORMAL: return 2;
    case MBO_TWO_LOGNORMAL: return 5;
    case MBO_BINS_HISTO: case MBO_BINS_INTERP: return m->nnodes[p];
    case MBO_EXPONENTIAL: return 1;
    default: return 0;
  }
}

static double lognorm(double x, double mode, double sigma, double amp) {
  /* megabeast/old/ensemble_model.py:13-47 */
  if (!(x > 0.0)) return 0.0;
  double mu = log(mode) + sigma * sigma;
  double norm = (1.0 / sqrt(2.0 * M_PI)) / (x * sigma);
  double t = (log(x) - mu) / sigma;
  return amp * norm * exp(-0.5 * (t * t));
}

static double eval_prior(const mbo_model* m, int p, const double* th, double x) {
  switch (m->kind[p]) {
    case MBO_FLAT: return 1.0;
    case MBO_LOGNORMAL: return lognorm(x, th[0], th[1], 1.0);
    case MBO_TWO_LOGNORMAL: {
      double n1 = 1.0 - 1.0 / (th[4] + 1.0), n2 = 1.0 / (th[4] + 1.0);
      return lognorm(x, th[0], th[2], n1) + lognorm(x, th[1], th[3], n2);
    }
    case MBO_BINS_HISTO: {   /* left-constant steps, last node -> last height */
      int n = m->nnodes[p], i = 0;
      const double* nd = m->nodes[p];
      while (i + 1 < n && x >= nd[i + 1]) ++i;
      return th[i];
    }
    case MBO_BINS_INTERP: {  /* np.interp: clamped linear */
      int n = m->nnodes[p];
      const double* nd = m->nodes[p];
      if (x <= nd[0]) return th[0];
      if (x >= nd[n - 1]) return th[n - 1];
      int i = 0;
      while (x >= nd[i + 1]) ++i;
      double slope = (th[i + 1] - th[i]) / (nd[i + 1] - nd[i]);
      return slope * (x - nd[i]) + th[i];
    }
    case MBO_EXPONENTIAL: return exp(-1.0 * pow(10.0, x) / (th[0] * 1e9));
    default: return 1.0;
  }
}

/*
 * cols[p]      : grid column of parameter p (length G) or NULL when kind[p] == MBO_FIXED
 * bin_of_row   : (age, Z) bin of every grid row, age-major (helpers.py:143-145)
 * age_of_bin   : index into ages[] for every bin;  coef[b] = metprior * massmult / avemass
 * indxs, vals  : (K, S) C-order exactly as the reference holds them
 * status[w]    : bit0 = some star non-finite (ValueError, :216-217)
 *                bit1 = total exactly 0.0 (exit(), :225-227)
 * returns 0, or -1 on allocation failure.
 */
int mbo_lnlike_batch(int64_t G, const double* const* cols, const double* prior_weight,
                     const double* grid_weight, const double* completeness,
                     const int32_t* bin_of_row, int32_t nbins, const int32_t* age_of_bin,
                     const double* coef, int32_t nages, const double* ages,
                     const mbo_model* model, int64_t S, int64_t K, const int64_t* indxs,
                     const double* vals, int32_t W, int32_t ndim, const double* phi,
                     double* lnlike, int32_t* status, int32_t nthreads) {
  int off[MBO_NPARAM], k = 0;
  for (int p = 0; p < MBO_NPARAM; ++p) { off[p] = k; k += nvars_of(model, p); }
  if (k != ndim) return -2;
  /* :71-74; nbins == 0: the caller wants the sum over stars alone (bench.py adds shard sums of several ranks) */
  const int poisson = model->kind[0] != MBO_FIXED && nbins > 0;
  int fail = 0;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    double* physmod = (double*)malloc(sizeof(double) * (size_t)G);
    double* sb = (double*)malloc(sizeof(double) * (size_t)(nbins > 0 ? nbins : 1) * 2);
    if (!physmod || !sb) {
#pragma omp atomic write
      fail = 1;
    } else {
#pragma omp for schedule(dynamic, 1)
      for (int w = 0; w < W; ++w) {
        const double* ph = phi + (size_t)w * ndim;
        for (int64_t g = 0; g < G; ++g) physmod[g] = 1.0;
        for (int p = 0; p < MBO_NPARAM; ++p) {
          if (model->kind[p] == MBO_FIXED) continue;
          const double* c = cols[p];
          for (int64_t g = 0; g < G; ++g) physmod[g] *= eval_prior(model, p, ph + off[p], c[g]);
        }
        double total = 0.0;
        if (poisson) {
          double nrm = 0.0;
          for (int64_t g = 0; g < G; ++g) nrm += physmod[g] * grid_weight[g];
          memset(sb, 0, sizeof(double) * (size_t)nbins * 2);
          for (int64_t g = 0; g < G; ++g) {
            double wgt = physmod[g] * grid_weight[g] / nrm;   /* helpers.py:133-134 */
            sb[2 * bin_of_row[g]] += wgt;
            sb[2 * bin_of_row[g] + 1] += wgt * completeness[g];
          }
          double pred = 0.0;
          for (int b = 0; b < nbins; ++b) {
            double s = sb[2 * b];
            if (s > 0) {                                          /* helpers.py:156 */
              double nexp = eval_prior(model, 0, ph + off[0], ages[age_of_bin[b]]) * coef[b];
              pred += nexp * sb[2 * b + 1] / s;
            }
          }
          total = (double)S * log(pred) - pred - lgamma((double)S + 1.0);   /* :182-186 */
        }
        int st = 0;
        for (int64_t s = 0; s < S; ++s) {
          double acc = 0.0;
          for (int64_t kk = 0; kk < K; ++kk) {
            double v = vals[kk * S + s];
            if (!isfinite(v)) continue;                          /* :196-198 */
            int64_t r = indxs[kk * S + s];
            if (r < 0) r += G;                                    /* numpy indexing counts negative indices from the end */
            acc += v * physmod[r] * prior_weight[r] * grid_weight[r] * completeness[r];
          }
          if (!isfinite(acc)) { st |= 1; continue; }
          if (acc != 0.0) total += log(acc);                     /* :218-223 */
        }
        if (total == 0.0) st |= 2;
        lnlike[w] = total;
        status[w] = st;
      }
    }
    free(physmod);
    free(sb);
  }
  (void)nages;
  return fail ? -1 : 0;
}

int mbo_sizeof_model(void) { return (int)sizeof(mbo_model); }
