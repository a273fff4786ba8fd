/* ./GetHI <param_file> -- same command line, banner, stage order and output files as the reference's
 * driver (src/main_gh.c:24-80).
 *
 * One process per GPU.  Without a launcher the program is its own: GH_NGPUS=P ./GetHI file forks P ranks
 * (before any CUDA call), rank 0 creates the NCCL id and hands it to the others through pipes.  Under
 * torchrun / mpirun-style launchers set GH_RANK, GH_NRANKS and GH_UNIQUE_ID_FILE instead. */
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
  ParamGetHI *par = read_run_params(fname);
  print_info("Seed : %u\n", par->seed_rng);
  create_d_and_vr_fields(par);
  get_HI(par);
  mk_T_maps_begin(par); /* non-blocking mk_T_maps ... */
  write_maps(par);      /* ... every rank writes the shells it owns as they arrive from the device */
  if (NodeThis == 0) timer(5);
  print_info("\n");
  print_info("|-------------------------------------------------|\n\n");
  param_gethi_free(par);
  return 0;
}

int gh_main(int argc, char **argv)
{
  if (argc != 2) {
    fprintf(stderr, "Usage: ./GetHI file_name\n");
    exit(0);
  }
  const char *env = getenv("GH_NGPUS");
  const int nranks = env ? atoi(env) : 1;
  if (nranks <= 1) return run_rank(argv[1], 0, 1, getenv("GH_DEVICE") ? atoi(getenv("GH_DEVICE")) : 0, NULL);

  /* fork first, touch CUDA / NCCL only in the children */
  int (*pipes)[2] = malloc(sizeof(int[2]) * nranks);
  for (int r = 1; r < nranks; r++)
    if (pipe(pipes[r])) { perror("pipe"); return 1; }
  pid_t *pids = calloc(nranks, sizeof(pid_t));
  for (int r = 0; r < nranks; r++) {
    pid_t pid = fork();
    if (pid < 0) { perror("fork"); return 1; }
    if (pid == 0) {
      unsigned char uid[GH_CUDA_UNIQUE_ID_BYTES];
      if (r == 0) {
        if (gh_cuda_get_unique_id(uid)) { fprintf(stderr, "Node 0, Fatal: %s\n", gh_cuda_last_error()); _exit(1); }
        for (int q = 1; q < nranks; q++)
          if (write(pipes[q][1], uid, sizeof(uid)) != (ssize_t)sizeof(uid)) _exit(1);
      } else {
        if (read(pipes[r][0], uid, sizeof(uid)) != (ssize_t)sizeof(uid)) _exit(1);
      }
      _exit(run_rank(argv[1], r, nranks, r, uid));
    }
    pids[r] = pid;
  }
  int bad = 0;
  for (int r = 0; r < nranks; r++) {
    int st = 0;
    waitpid(pids[r], &st, 0);
    if (!WIFEXITED(st) || WEXITSTATUS(st)) bad = 1;
  }
  free(pids);
  free(pipes);
  return bad;
}

#ifndef GH_NO_MAIN
int main(int argc, char **argv) { return gh_main(argc, argv); }
#endif
