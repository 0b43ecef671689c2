/* Plain-C consumer of include/bagpu.h: proves the header is valid C99 and that libbagpu.so links and answers
 * from a non-Python host (what a Julia ccall or a C++ embedder sees).  Host-only entry points: no GPU needed.
 * Built and run by tests/test_host.py::test_plain_c_consumer. */
#include <stdio.h>
#include <string.h>
#include "bagpu.h"

int main(void) {
  ba_lm_params p;
  ba_lm_default_params(&p);
  if (p.nu_d != 3.0 || p.nu_m != 3.0 || p.lambda != 30.0 || p.delta_d != 2.0 || p.ite_max != 200) return 1;
  if (!strstr(ba_version(), "sm_100a")) return 2;
  /* 3 points with 2, 3, 2 observations, point-major: cuts for 2 ranks fall on a point boundary */
  int64_t pnt[7] = {1, 1, 2, 2, 2, 3, 3}, cuts[3];
  if (ba_partition_observations(7, pnt, 2, cuts) != BA_OK) return 3;
  if (cuts[0] != 0 || cuts[2] != 7 || (cuts[1] != 2 && cuts[1] != 5)) return 4;
  int64_t bad[7] = {2, 1, 2, 2, 2, 3, 3};
  if (ba_partition_observations(7, bad, 2, cuts) != BA_ERR_UNSORTED) return 5;
  /* argument validation happens before any CUDA call */
  int64_t cam[2] = {1, 9};
  double pt2d[4] = {0, 0, 0, 0};
  ba_handle* h = NULL;
  if (ba_create(2, 3, 2, cam, pnt, pt2d, 0, &h) != BA_ERR_ARG || !h) return 6;
  if (!strstr(ba_last_error(h), "out of range")) return 7;
  ba_destroy(h);
  printf("abi ok: %s\n", ba_version());
  return 0;
}
