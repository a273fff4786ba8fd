/* Process runtime of the host layer: rank bookkeeping, rank-0 logging, fatal errors, stage timer.
 * Same observable behaviour as reference src/common_gh.c:31-120 and src/common.c:77-131 (timer lines
 * ">    Relative time ellapsed %.1lf ms" are what downstream scripts parse). */
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "gh_host.h"

int NodeThis = 0, NNodes = 1;
static int g_device = 0;
static unsigned char g_uid[GH_CUDA_UNIQUE_ID_BYTES];
static int g_have_uid = 0;

void gh_mpi_init(int rank, int nranks, int device, const void *unique_id)
{
  NodeThis = rank;
  NNodes = nranks;
  g_device = device;
  g_have_uid = unique_id != NULL;
  if (unique_id) memcpy(g_uid, unique_id, sizeof(g_uid));
}

int gh_runtime_device(void) { return g_device; }
const void *gh_runtime_unique_id(void) { return g_have_uid ? g_uid : NULL; }

void print_info(const char *fmt, ...)
{
  if (NodeThis != 0) return;
  va_list ap;
  va_start(ap, fmt);
  vprintf(fmt, ap);
  va_end(ap);
}

void report_error(int level, const char *fmt, ...)
{
  char msg[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof(msg), fmt, ap);
  va_end(ap);
  if (level) {
    fprintf(stderr, "Node %d, Fatal: %s", NodeThis, msg);
    exit(level);
  }
  fprintf(stderr, "Node %d, Warning: %s", NodeThis, msg);
}

static double now_s(void)
{
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

/* GH_HOST_TIMING=1: wall time of each host phase on stderr (rank 0), for the T_total breakdown of bench.py */
void gh_phase(const char *name)
{
  static double last = 0;
  static int on = -1;
  if (on < 0) on = getenv("GH_HOST_TIMING") != NULL;
  const double t = now_s();
  if (on && NodeThis == 0 && name && last > 0) fprintf(stderr, "[gh_host] %s %.1f ms\n", name, 1000 * (t - last));
  last = t;
}

void timer(int i)
{
  static double rel0, abs0;
  const double t = now_s();
  if (i == 0) rel0 = t;
  else if (i == 1) printf(">    Relative time ellapsed %.1lf ms\n", 1000 * (t - rel0));
  else if (i == 2) { printf(">    Relative time ellapsed %.1lf ms\n", 1000 * (t - rel0)); rel0 = now_s(); }
  else if (i == 4) abs0 = t;
  else if (i == 5) printf(">    Total time ellapsed %.1lf ms\n", 1000 * (t - abs0));
}
