/*
 * A plain C client of the C ABI (include/megabeast_b200.h): no Python, no torch, no CUDA headers.
 * Reads one problem from a binary file written by tests/test_c_client.py, evaluates the ensemble
 * likelihood of W walkers through mb_create / mb_lnlike / mb_destroy and prints one line per
 * walker: "<lnlike %.17g> <status>".  Exit code 3 + the library's message when mb_create fails
 * (e.g. no CUDA device: the library has no CPU fallback).
 *
 * File layout (little endian): int64 G, K, S, W, ndim, nnodes, n_ages, n_mets; double avemass;
 * then doubles: nodes[nnodes], logA[G], Z[G], Av[G], Rv[G], prior_weight[G], grid_weight[G],
 * completeness[G], ages[n_ages], mets[n_mets], massmult[n_ages*n_mets], phi[W*ndim], vals[K*S];
 * then int64 indxs[K*S].  Model: logA bins_histo (free), Av and Rv lognormal (free), fA fixed.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "megabeast_b200.h"

static double* read_doubles(FILE* f, long long n) {
  double* p = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
  if (!p || fread(p, sizeof(double), (size_t)n, f) != (size_t)n) { fprintf(stderr, "short read\n"); exit(2); }
  return p;
}

int main(int argc, char** argv) {
  if (argc < 2) { fprintf(stderr, "usage: %s problem.bin\n", argv[0]); return 2; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  long long hdr[8];
  double avemass;
  if (fread(hdr, sizeof(long long), 8, f) != 8 || fread(&avemass, sizeof(double), 1, f) != 1) return 2;
  const long long G = hdr[0], K = hdr[1], S = hdr[2], W = hdr[3], ndim = hdr[4], nnodes = hdr[5],
                  n_ages = hdr[6], n_mets = hdr[7];
  double* nodes = read_doubles(f, nnodes);
  double* logA = read_doubles(f, G);
  double* Z = read_doubles(f, G);
  double* Av = read_doubles(f, G);
  double* Rv = read_doubles(f, G);
  double* pw = read_doubles(f, G);
  double* gw = read_doubles(f, G);
  double* compl_ = read_doubles(f, G);
  double* ages = read_doubles(f, n_ages);
  double* mets = read_doubles(f, n_mets);
  double* massmult = read_doubles(f, n_ages * n_mets);
  double* phi = read_doubles(f, W * ndim);
  double* vals = read_doubles(f, K * S);
  int64_t* indxs = (int64_t*)malloc(sizeof(int64_t) * (size_t)(K * S > 0 ? K * S : 1));
  if (fread(indxs, sizeof(int64_t), (size_t)(K * S), f) != (size_t)(K * S)) { fprintf(stderr, "short read\n"); return 2; }
  fclose(f);

  mb_model_desc model;
  memset(&model, 0, sizeof model);
  model.kind[MB_P_LOGA] = MB_BINS_HISTO;
  model.nnodes[MB_P_LOGA] = (int32_t)nnodes;
  memcpy(model.nodes[MB_P_LOGA], nodes, sizeof(double) * (size_t)nnodes);
  model.kind[MB_P_AV] = MB_LOGNORMAL;
  model.kind[MB_P_RV] = MB_LOGNORMAL;
  model.kind[MB_P_FA] = MB_FIXED;

  mb_grid_desc grid;
  memset(&grid, 0, sizeof grid);
  grid.G = G;
  grid.col[MB_P_LOGA] = logA; grid.col[MB_P_AV] = Av; grid.col[MB_P_RV] = Rv;
  grid.logA = logA; grid.Z = Z;
  grid.prior_weight = pw; grid.grid_weight = gw; grid.completeness = compl_;

  mb_stars_desc stars;
  memset(&stars, 0, sizeof stars);
  stars.K = K; stars.S = S; stars.indxs = indxs; stars.vals = vals;
  stars.star_begin = 0; stars.star_end = S; stars.n_stars_total = S;

  mb_massmult_desc mm;
  mm.n_ages = (int32_t)n_ages; mm.n_mets = (int32_t)n_mets;
  mm.ages = ages; mm.mets = mets; mm.massmult = massmult; mm.avemass = avemass;

  mb_options opt;
  memset(&opt, 0, sizeof opt);
  opt.merge_duplicates = 1;

  mb_handle* h = NULL;
  if (mb_create(&grid, &stars, &model, &mm, &opt, &h)) {
    fprintf(stderr, "mb_create failed: %s\n", mb_last_error());
    return 3;
  }
  double* lnlike = (double*)malloc(sizeof(double) * (size_t)W);
  int32_t* status = (int32_t*)malloc(sizeof(int32_t) * (size_t)W);
  if (mb_lnlike(h, phi, (int32_t)W, (int32_t)ndim, lnlike, status)) {
    fprintf(stderr, "mb_lnlike failed: %s\n", mb_last_error());
    return 4;
  }
  mb_info info;
  mb_get_info(h, &info);
  fprintf(stderr, "abi %d mode %d entries %lld launches %d\n", info.abi_version, info.mode,
          (long long)info.n_entries, info.launches_last_eval);
  for (long long w = 0; w < W; ++w) printf("%.17g %d\n", lnlike[w], status[w]);
  mb_destroy(h);
  return 0;
}
