// ba_lm.cu -- placeholder while the eval path gets its first GPU run; replaced by the real solver.
#include <cmath>
#include "ba_internal.h"
namespace ba {
int lm_prepare(ba_handle*) { return BA_OK; }
void lm_release(ba_handle*) {}
}
extern "C" {
void ba_lm_default_params(ba_lm_params* p) {
  const double eps = 2.220446049250313e-16;
  p->restol = p->ortol = p->rtol = cbrt(eps);
  p->satol = p->srtol = p->oatol = p->atol = sqrt(eps);
  p->nu_d = 3; p->nu_m = 3; p->lambda = 30; p->delta_d = 2;
  p->ite_max = 200; p->linesearch = 0; p->pcg_max_iter = 500; p->pcg_tol = 1e-13;
}
int ba_lm_step(ba_handle* h, const double*, double, double, int32_t, double*, double*, double*, double*, int32_t*) {
  if (h) h->err = "ba_lm_step: not built yet";
  return BA_ERR_ARG;
}
int ba_lm_solve(ba_handle* h, double*, const ba_lm_params*, ba_lm_stats*, ba_iter_cb, void*) {
  if (h) h->err = "ba_lm_solve: not built yet";
  return BA_ERR_ARG;
}
}
