/*
 * mcmc_main.c -- the `mcmc` command line over the C ABI (host code stays in C).
 *
 * Drop-in for the reference binary's contract (C_Implementation/mcmc.c:102-210, called by
 * script.py:44-45 as `./mcmc <chain_index> < dataset.txt` with GSL_RNG_SEED in the environment):
 *   - argv[1] = chain index -> files go to Chains/chain_XX/ (two digits, like mcmc.c:148-178)
 *   - stdin   = dataset in the reference's .txt format
 *   - 1000 burn-in + 1000 sampling calls of 10 sweeps, one thinned sample per call
 *   - writes chain_data.csv, exp_data.csv, taxa.csv, sites.csv, hard_sites.csv
 *   - stderr echoes "GSL_RNG_SEED=<n>" like gsl_rng_env_setup(); exit 0 / 1
 * Differences, all deliberate: the random stream is the structured Philox stream keyed by
 * (GSL_RNG_SEED, chain index) instead of GSL's MT19937; Chains/chain_XX/ is created when missing
 * (the reference dereferences a NULL FILE*); `mcmc` without arguments prints usage instead of
 * crashing at atoi(NULL) (mcmc.c:153).
 *
 * Batch mode (replaces script.py's Pool over 100 processes by one call over all the chains, on one GPU or
 * sharded over the GPUs of the box -- ser_multi_*):
 *   mcmc --chains N [--gpus G] [--first I] [--burn B] [--samples S] [--seed X] [--dataset file]
 *        [--chains-dir DIR] [--select K] [--po file.csv] [--device D] [--manycd 0|1]
 * Replay mode: SER_TAPE_IN=<file of raw doubles> mcmc <idx> < dataset.txt
 */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>

#include "../../include/seriation_b200.h"

static void die(const char *what)
{
  fprintf(stderr, "mcmc: %s: %s\n", what, ser_last_error());
  exit(1);
}

static void usage(const char *argv0)
{
  fprintf(stderr,
          "usage: %s <chain_index> < dataset.txt          (reference-compatible single chain)\n"
          "       %s [manycd Tburnin T] < dataset.txt      (manycd 1: per-taxon c, d; chain 0)\n"
          "       %s --chains N [--gpus G] [--first I] [--burn B] [--samples S] [--seed X] [--dataset F]\n"
          "              [--chains-dir DIR] [--select K] [--po out.csv] [--device D] [--manycd 0|1]\n",
          argv0, argv0, argv0);
  exit(1);
}

static double *read_tape(const char *path, uint64_t *len)
{
  FILE *f = fopen(path, "rb");
  long sz;
  double *buf;
  if (!f) { fprintf(stderr, "mcmc: cannot open tape %s\n", path); exit(1); }
  fseek(f, 0, SEEK_END);
  sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  buf = (double *)malloc((size_t)sz + 8);
  if (!buf || fread(buf, 1, (size_t)sz, f) != (size_t)sz) { fprintf(stderr, "mcmc: cannot read tape %s\n", path); exit(1); }
  fclose(f);
  *len = (uint64_t)sz / sizeof(double);
  return buf;
}

int main(int argc, char **argv)
{
  int n_chains = 1, first = 0, burn = 1000, samples = 1000, select_k = 0, device = 0, batch = 0, manycd = 0, gpus = 1, i;
  unsigned long seed = 0;
  const char *dataset = NULL, *chains_dir = "Chains", *po_path = NULL, *tape_path = getenv("SER_TAPE_IN");
  const char *env_seed = getenv("GSL_RNG_SEED");
  ser_dataset *ds = NULL;
  ser_run *run = NULL;   /* single-chain replay */
  ser_multi *multi = NULL; /* everything else: the chains of the call over `gpus` devices */
  ser_run_config cfg;
  int32_t N, M, nh, bad = 0, devs[8];

  if (env_seed) {
    seed = strtoul(env_seed, NULL, 0);
    fprintf(stderr, "GSL_RNG_SEED=%lu\n", seed);
  }
  if (argc == 1) usage(argv[0]);
  if (argv[1][0] == '-' && argv[1][1] == '-') {
    batch = 1;
    for (i = 1; i < argc; i++) {
      const char *a = argv[i], *v = (i + 1 < argc) ? argv[i + 1] : NULL;
      if (!v) usage(argv[0]);
      if (!strcmp(a, "--chains")) n_chains = atoi(v);
      else if (!strcmp(a, "--gpus")) gpus = atoi(v);
      else if (!strcmp(a, "--first")) first = atoi(v);
      else if (!strcmp(a, "--burn")) burn = atoi(v);
      else if (!strcmp(a, "--samples")) samples = atoi(v);
      else if (!strcmp(a, "--seed")) seed = strtoul(v, NULL, 0);
      else if (!strcmp(a, "--dataset")) dataset = v;
      else if (!strcmp(a, "--chains-dir")) chains_dir = v;
      else if (!strcmp(a, "--select")) select_k = atoi(v);
      else if (!strcmp(a, "--po")) po_path = v;
      else if (!strcmp(a, "--device")) device = atoi(v);
      else if (!strcmp(a, "--manycd")) manycd = atoi(v) != 0;
      else usage(argv[0]);
      i++;
    }
    if (n_chains < 1 || burn < 0 || samples < 0 || gpus < 1 || gpus > 8 || n_chains < gpus) usage(argv[0]);
  } else if (argc == 2) {
    char *end = NULL;
    const long idx = strtol(argv[1], &end, 10); /* mcmc.c:115,153 */
    if (end == argv[1] || *end || idx < 0 || idx > 99) {
      /* the reference builds "Chains/chain_XX" from two digits (mcmc.c:148-178); an index outside 0..99 has no
       * directory there -- refuse it up front instead of simulating 20 000 sweeps and writing nothing */
      fprintf(stderr, "mcmc: chain index '%s' must be an integer in 0..99 (Chains/chain_XX)\n", argv[1]);
      return 1;
    }
    first = (int)idx;
  } else if (argc == 4) {
    if (!(sscanf(argv[1], "%d", &manycd) == 1 && sscanf(argv[2], "%d", &burn) == 1 && burn >= 0 &&
          sscanf(argv[3], "%d", &samples) == 1 && samples >= 0))
      usage(argv[0]);
    manycd = manycd != 0; /* mcmc_readmodel tests it as a flag (mcmc.c:363) */
  } else {
    usage(argv[0]);
  }

  if (ser_dataset_read_txt(dataset, &ds)) { fprintf(stderr, "%s\n", ser_last_error()); return 1; } /* mcmc_readmodel's messages */
  ser_dataset_dims(ds, &N, &M, &nh);

  memset(&cfg, 0, sizeof(cfg));
  cfg.struct_size = (uint32_t)sizeof(cfg);
  cfg.n_chains = n_chains;
  cfg.chain_offset = first;
  cfg.sweeps_per_call = 10;
  cfg.mode = tape_path ? SER_MODE_REPLAY : SER_MODE_FREE;
  cfg.seed = (uint32_t)seed;
  cfg.store = (n_chains <= 100) ? SER_STORE_FULL : SER_STORE_PI;
  cfg.max_samples = samples;
  cfg.device = device;
  cfg.manycd = manycd;
  for (i = 0; i < 8; i++) devs[i] = device + i;

  if (tape_path) { /* replay of one recorded chain (validation against the reference) */
    uint64_t offs[2] = {0, 0};
    double *tape;
    if (n_chains != 1 || gpus != 1) { fprintf(stderr, "mcmc: SER_TAPE_IN drives exactly one chain\n"); return 1; }
    if (ser_run_create(ds, &cfg, &run)) die("ser_run_create");
    tape = read_tape(tape_path, &offs[1]);
    if (ser_run_set_tapes(run, tape, offs)) die("ser_run_set_tapes");
    free(tape);
    if (ser_run_init(run)) die("ser_run_init");
    if (ser_run_advance_both(run, burn, samples)) die("sweeps"); /* mcmc.c:140-143, :180-185 */
    if (ser_run_sync(run)) die("ser_run_sync");
    if (ser_run_check(run, &bad)) { fprintf(stderr, "main: error. (%s)\n", ser_last_error()); return 1; } /* mcmc.c:199-204 */
  } else {
    if (ser_multi_create(ds, &cfg, gpus, devs, &multi)) die("ser_multi_create");
    if (ser_multi_init(multi)) die("ser_multi_init");
    if (ser_multi_advance(multi, burn, samples)) die("sweeps"); /* mcmc.c:140-143, :180-185 */
    if (ser_multi_sync(multi)) die("ser_multi_sync");
    if (ser_multi_check(multi, &bad)) { fprintf(stderr, "main: error. (%s)\n", ser_last_error()); return 1; }
  }

  /* Chains/chain_XX/ for the chains whose index has a two-digit directory */
  if (cfg.store == SER_STORE_FULL) {
    int skipped = 0;
    mkdir(chains_dir, 0777);
    for (i = 0; i < n_chains; i++) {
      char dir[1024];
      const int idx = first + i;
      ser_run *owner = run;
      int32_t local = i;
      if (idx < 0 || idx > 99) { skipped++; continue; } /* the reference's directory name has two digits (mcmc.c:148-178) */
      if (multi && ser_multi_locate(multi, idx, &owner, &local)) die("ser_multi_locate");
      snprintf(dir, sizeof(dir), "%s/chain_%02d", chains_dir, idx);
      if (mkdir(dir, 0777) && errno != EEXIST) { fprintf(stderr, "mcmc: cannot create %s\n", dir); return 1; }
      if (ser_write_chain_files(owner, local, dir)) die("ser_write_chain_files");
    }
    if (skipped) fprintf(stderr, "mcmc: %d chain(s) outside the index range 0..99 have no Chains/chain_XX directory and were not written\n", skipped);
  } else if (batch && !select_k) {
    fprintf(stderr, "mcmc: %d chains is more than the 100 the reference's file layout holds: no chain files are written; "
                    "use --select K [--po FILE] for the cross-chain summaries\n", n_chains);
  }

  if (batch && multi) {
    double *e = (double *)malloc(sizeof(double) * (size_t)n_chains), *ec = (double *)malloc(sizeof(double) * (size_t)n_chains);
    double *ed = (double *)malloc(sizeof(double) * (size_t)n_chains), ms = 0.0;
    int32_t ns = 0, peer = 0;
    if (!e || !ec || !ed) { fprintf(stderr, "mcmc: out of memory\n"); return 1; }
    if (ser_multi_chain_stats(multi, e, ec, ed, &ns)) die("ser_multi_chain_stats");
    ser_multi_elapsed_ms(multi, &ms, 0);
    ser_multi_layout(multi, NULL, NULL, &peer);
    printf("chains %d  gpus %d%s  sites %d  taxa %d  hard %d  sweeps/chain %d  gpu_ms %.1f  sweeps/s %.0f\n", n_chains, gpus,
           gpus > 1 ? (peer ? " (peer stores)" : " (nccl)") : "", N, M, nh, (burn + samples) * 10, ms,
           ms > 0 ? (double)n_chains * (burn + samples) * 10 / (ms * 1e-3) : 0.0);
    if (select_k > 0) {
      int32_t *chosen = (int32_t *)malloc(sizeof(int32_t) * (size_t)select_k), nchosen = 0;
      int32_t *counts = po_path ? (int32_t *)calloc((size_t)select_k * N * N, sizeof(int32_t)) : NULL;
      double mn, sd;
      if (!chosen || (po_path && !counts)) { fprintf(stderr, "mcmc: out of memory\n"); return 1; }
      /* E[-logL] -> selection -> pair-order counts in one step on the device(s) (script.py:70-99, :155-189) */
      if (ser_multi_cross_chain(multi, select_k, chosen, &nchosen, &mn, &sd, counts)) die("ser_multi_cross_chain");
      printf("selection: min E[-logL] %.6f  sigma %.6f  chosen", mn, sd);
      for (i = 0; i < nchosen; i++) printf(" %d", chosen[i]);
      printf("\n");
      if (nchosen > 0) {
        double sc = 0.0, sdd = 0.0;
        for (i = 0; i < nchosen; i++) { sc += ec[chosen[i] - first]; sdd += ed[chosen[i] - first]; }
        printf("E[c] %.6f  E[d] %.6f over the chosen chains\n", sc / nchosen, sdd / nchosen);
      }
      if (po_path && nchosen > 0) {
        double *po = (double *)malloc(sizeof(double) * (size_t)N * N);
        FILE *f;
        int r, c;
        if (!po) { fprintf(stderr, "mcmc: out of memory\n"); return 1; }
        if (ser_po_finalize(counts, nchosen, N, select_k, 1, po)) die("ser_po_finalize");
        if (!(f = fopen(po_path, "w"))) { fprintf(stderr, "mcmc: cannot open %s\n", po_path); return 1; }
        for (r = 0; r < N; r++)
          for (c = 0; c < N; c++) fprintf(f, "%.6f%c", po[(size_t)r * N + c], c + 1 < N ? ',' : '\n');
        fclose(f);
        free(po);
      }
      free(chosen); free(counts);
    }
    free(e); free(ec); free(ed);
  }
  ser_run_destroy(run);
  ser_multi_destroy(multi);
  ser_dataset_free(ds);
  return 0;
}
