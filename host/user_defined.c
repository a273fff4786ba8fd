/* User-defined HI model, the counterpart of reference src/user_defined.c:27-35 ("edit this file"): the neutral
 * hydrogen fraction and the HI bias as functions of redshift.  cosmo_set samples both on the radial table grid and
 * hands the samples to the device (gh_cuda_params.frac_HI_arr / bias_HI_arr), so an edit here changes what the
 * get_HI kernel computes -- nothing about them is compiled into the CUDA library. */
#include <math.h>
#include "gh_host.h"

double fraction_HI(double z) { return 0.008 * pow(1 + z, 0.6); }

double bias_HI(double z) { return 0.904 + 0.135 * pow(1 + z, 1.696); }
