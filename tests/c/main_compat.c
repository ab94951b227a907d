/* Downstream-compile test, modelled on the program the reference's CI generates
 * (python/kdsource/test.py:97-124): open a source, take w_crit, sample, destroy.
 * Usage: main_compat source.xml N  -> prints one particle per line. */
#include <stdio.h>
#include <stdlib.h>

#include "kdsource_compat.h"

int main(int argc, char** argv) {
  if (argc != 3) return 1;
  KDSource* kds = KDS_open(argv[1]);
  const long n = atol(argv[2]);
  printf("npts %lld J %g kernel %c order %d\n", kds->plist->npts, kds->J, kds->kernel,
         kds->geom->ord);
  double w_crit = KDS_w_mean(kds, 1000, NULL);
  printf("w_crit %.17g\n", w_crit);
  mcpl_particle_t part;
  for (long i = 0; i < n; ++i) {
    KDS_sample2(kds, &part, 1, w_crit, NULL, 1);
    printf("P %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %d\n", part.ekin, part.position[0],
           part.position[1], part.position[2], part.direction[0], part.direction[1],
           part.direction[2], part.time, part.weight, part.pdgcode);
  }
  /* w_crit <= 0: the list itself, in order, unperturbed, original weights */
  for (long i = 0; i < 3; ++i) {
    KDS_sample2(kds, &part, 0, -1, NULL, 1);
    printf("S %.9g %.9g %.9g %.9g\n", part.ekin, part.position[0], part.position[1], part.weight);
  }
  KDS_destroy(kds);
  printf("Successfully sampled particles.\n");
  return 0;
}
