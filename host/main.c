/* ./GetHI <param_file> -- same command line, banner, stage order and output files as the reference's
 * driver (src/main_gh.c:24-80).
 *
 * One process per GPU.  Without a launcher the program is its own: GH_NGPUS=P ./GetHI file forks P ranks
 * (before any CUDA call), rank 0 creates the NCCL id and hands it to the others through pipes.  Under
 * torchrun / mpirun-style launchers: GH_RANK + GH_NRANKS (or RANK + WORLD_SIZE) and GH_UNIQUE_ID_FILE, see
 * run_under_launcher below. */
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>
#include "gh_host.h"

static int run_rank(const char *fname, int rank, int nranks, int device, const void *uid)
{
  gh_mpi_init(rank, nranks, device, uid);
  setbuf(stdout, NULL);
  print_info("\n");
  print_info("|-------------------------------------------------|\n");
  print_info("|                      GetHI                      |\n");
  print_info("|-------------------------------------------------|\n\n");
  if (NodeThis == 0) timer(4);
  gh_phase(NULL);
  ParamGetHI *par = read_run_params(fname);
  gh_phase("run-parameter banner");
  print_info("Seed : %u\n", par->seed_rng);
  create_d_and_vr_fields(par);
  if (par->do_psources) { /* src/main_gh.c:55-59: Poisson-sample the sources from the Gaussian field, before get_HI */
    setup_psources(par);
    get_point_sources(par);
  }
  get_HI(par);
  gh_phase("create_d_and_vr_fields + get_HI");
  if (par->do_psources) { /* the source maps are needed before the writers start: blocking form, then src/main_gh.c:64-65 */
    mk_T_maps(par);
    mk_psources_maps(par);
  } else {
    mk_T_maps_begin(par); /* non-blocking mk_T_maps ... */
  }
  write_maps(par);      /* ... every rank writes the shells it owns as they arrive from the device */
  gh_phase("mk_T_maps + write_maps");
  if (NodeThis == 0) timer(5);
  print_info("\n");
  print_info("|-------------------------------------------------|\n\n");
  param_gethi_free(par);
  return 0;
}

/* Launcher-provided layout (torchrun / mpirun / srun style): GH_NRANKS (or WORLD_SIZE) > 1 together with GH_RANK (or
 * RANK); the device is GH_DEVICE, else LOCAL_RANK, else the rank.  The NCCL id travels through a file both sides can
 * see, GH_UNIQUE_ID_FILE: rank 0 writes it (to a temporary name, then rename, so a reader never sees half of it),
 * the others wait for it to appear (120 s). */
static int env_int(const char *a, const char *b, int dflt)
{
  const char *v = getenv(a);
  if (!v && b) v = getenv(b);
  return v ? atoi(v) : dflt;
}

static int run_under_launcher(const char *fname, int rank, int nranks)
{
  const int device = env_int("GH_DEVICE", "LOCAL_RANK", rank);
  const char *idf = getenv("GH_UNIQUE_ID_FILE");
  unsigned char uid[GH_CUDA_UNIQUE_ID_BYTES];
  if (!idf) { fprintf(stderr, "Node %d, Fatal: GH_NRANKS/WORLD_SIZE > 1 needs GH_UNIQUE_ID_FILE\n", rank); return 1; }
  if (rank == 0) {
    char tmp[1024];
    snprintf(tmp, sizeof(tmp), "%s.tmp.%ld", idf, (long)getpid());
    if (gh_cuda_get_unique_id(uid)) { fprintf(stderr, "Node 0, Fatal: %s\n", gh_cuda_last_error()); return 1; }
    FILE *f = fopen(tmp, "wb");
    if (!f || fwrite(uid, 1, sizeof(uid), f) != sizeof(uid) || fclose(f) || rename(tmp, idf)) {
      fprintf(stderr, "Node 0, Fatal: cannot write %s\n", idf);
      return 1;
    }
  } else {
    int ok = 0;
    for (int tries = 0; tries < 1200 && !ok; tries++) {
      FILE *f = fopen(idf, "rb");
      if (f) {
        ok = fread(uid, 1, sizeof(uid), f) == sizeof(uid);
        fclose(f);
      }
      if (!ok) usleep(100000);
    }
    if (!ok) { fprintf(stderr, "Node %d, Fatal: no NCCL id in %s after 120 s\n", rank, idf); return 1; }
  }
  return run_rank(fname, rank, nranks, device, uid);
}

int gh_main(int argc, char **argv)
{
  if (argc != 2) {
    fprintf(stderr, "Usage: ./GetHI file_name\n");
    exit(0);
  }
  const int launched = env_int("GH_NRANKS", "WORLD_SIZE", 1);
  if (launched > 1 && !getenv("GH_NGPUS")) return run_under_launcher(argv[1], env_int("GH_RANK", "RANK", 0), launched);
  const char *env = getenv("GH_NGPUS");
  const int nranks = env ? atoi(env) : 1;
  if (nranks <= 1) return run_rank(argv[1], 0, 1, getenv("GH_DEVICE") ? atoi(getenv("GH_DEVICE")) : 0, NULL);

  /* fork first, touch CUDA / NCCL only in the children.  pipes[r] carries the NCCL id from rank 0 to rank r; every
   * process closes the ends it does not use, so a reader sees end-of-file as soon as rank 0 dies without writing */
  int (*pipes)[2] = malloc(sizeof(int[2]) * nranks);
  for (int r = 1; r < nranks; r++)
    if (pipe(pipes[r])) { perror("pipe"); return 1; }
  pid_t *pids = calloc(nranks, sizeof(pid_t));
  for (int r = 0; r < nranks; r++) {
    pid_t pid = fork();
    if (pid < 0) {
      perror("fork");
      for (int q = 0; q < r; q++) kill(pids[q], SIGTERM);
      return 1;
    }
    if (pid == 0) {
      unsigned char uid[GH_CUDA_UNIQUE_ID_BYTES];
      for (int q = 1; q < nranks; q++) {
        if (r != 0) close(pipes[q][1]);            /* only rank 0 writes */
        if (q != r) close(pipes[q][0]);            /* rank r reads its own pipe only */
      }
      if (r == 0) {
        if (gh_cuda_get_unique_id(uid)) { fprintf(stderr, "Node 0, Fatal: %s\n", gh_cuda_last_error()); _exit(1); }
        for (int q = 1; q < nranks; q++) {
          if (write(pipes[q][1], uid, sizeof(uid)) != (ssize_t)sizeof(uid)) _exit(1);
          close(pipes[q][1]);
        }
      } else {
        if (read(pipes[r][0], uid, sizeof(uid)) != (ssize_t)sizeof(uid)) {
          fprintf(stderr, "Node %d, Fatal: rank 0 went away before handing out the NCCL id\n", r);
          _exit(1);
        }
        close(pipes[r][0]);
      }
      _exit(run_rank(argv[1], r, nranks, r, uid));
    }
    pids[r] = pid;
  }
  for (int r = 1; r < nranks; r++) { close(pipes[r][0]); close(pipes[r][1]); }
  /* first abnormal exit takes the other ranks down with it (they would otherwise wait in a collective for ever) */
  int bad = 0, left = nranks;
  while (left > 0) {
    int st = 0;
    const pid_t done = wait(&st);
    if (done < 0) break;
    left--;
    if ((!WIFEXITED(st) || WEXITSTATUS(st)) && !bad) {
      bad = 1;
      for (int r = 0; r < nranks; r++)
        if (pids[r] != done) kill(pids[r], SIGTERM);
    }
  }
  free(pids);
  free(pipes);
  return bad;
}

#ifndef GH_NO_MAIN
int main(int argc, char **argv) { return gh_main(argc, argv); }
#endif
