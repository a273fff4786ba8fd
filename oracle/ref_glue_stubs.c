/* TEST INFRASTRUCTURE ONLY -- the two point-source entry points of the reference's grid_tools.c / pixelize.c, which
 * the GPU hot path does not provide (do_psources=1 is out of scope): the reference's own main_gh.c references them,
 * so the drop-in link (oracle/Makefile, _ref/GetHI_gpu) needs the symbols. */
#include "common_gh.h"

void get_point_sources(ParamGetHI *par)
{
  (void)par;
  report_error(1, "do_psources=1 is not supported by the GPU hot path\n");
}

void mk_psources_maps(ParamGetHI *par)
{
  (void)par;
  report_error(1, "do_psources=1 is not supported by the GPU hot path\n");
}
