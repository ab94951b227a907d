/* MultiSource host program (clib/src/kdsource.c:375-473): MS_open -> MS_w_mean ->
 * N x MS_sample2, then a biased KDS_sample2 / KDS_w_mean on the first source and a
 * host walk over PList_get / Geom_perturb / PList_next / Geom_next.
 * Usage: main_ms a.xml b.xml wa wb N */
#include <stdio.h>
#include <stdlib.h>

#include "kdsource_compat.h"

static double bias_fn(const mcpl_particle_t* p) { return p->ekin > 0.05 ? 3.0 : 0.5; }

static unsigned long long xs = 88172645463325252ull;
static double my_rng(void* st) {
  (void)st;
  xs ^= xs << 13;
  xs ^= xs >> 7;
  xs ^= xs << 17;
  return (double)(xs >> 11) / 9007199254740992.0;
}

int main(int argc, char** argv) {
  if (argc != 6) return 1;
  const char* files[2] = {argv[1], argv[2]};
  const double ws[2] = {atof(argv[3]), atof(argv[4])};
  const long n = atol(argv[5]);
  MultiSource* ms = MS_open(2, files, ws);
  printf("MS len %d J %.17g cdf %.17g %.17g\n", ms->len, ms->J, ms->cdf[0], ms->cdf[1]);
  const double w_crit = MS_w_mean(ms, 1000, NULL);
  printf("w_crit %.17g\n", w_crit);
  mcpl_particle_t part;
  for (long i = 0; i < n; ++i) {
    MS_sample2(ms, &part, 1, w_crit, NULL, 1);
    printf("P %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.9g %.17g %d\n", part.ekin, part.position[0],
           part.position[1], part.position[2], part.direction[0], part.direction[1],
           part.direction[2], part.time, part.weight, part.pdgcode);
  }
  /* bias: picks follow w * bias, weight 1 / bias; unperturbed so the pick is visible */
  KDSource* k0 = ms->s[0];
  printf("wb_mean %.17g\n", KDS_w_mean(k0, 9, bias_fn));
  for (long i = 0; i < n; ++i) {
    KDS_rand_sample2(my_rng, NULL, k0, &part, 0, 1.0, bias_fn, 1);
    printf("B %.9g %.17g\n", part.ekin, part.weight);
  }
  /* a host that walks the list itself, as KDS_rand_sample2 does with w_crit <= 0 */
  for (int i = 0; i < 5; ++i) {
    PList_get(k0->plist, &part);
    const double e0 = part.ekin;
    Geom_perturb(my_rng, NULL, k0->geom, &part);
    printf("W %.9g %.9g %.9g %d %d\n", e0, part.ekin, part.weight, PList_next(k0->plist, 1),
           Geom_next(k0->geom, 1));
  }
  MS_destroy(ms);
  printf("Successfully sampled particles.\n");
  return 0;
}
