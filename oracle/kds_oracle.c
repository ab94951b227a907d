/*
 * kds_oracle.c -- CPU restatement (fp64, plain C) of KDSource's kernel-density
 * hot path.  TEST INFRASTRUCTURE ONLY: nothing in the product package may
 * include, link or call this file.  Allowed users: tests/, bench.py's
 * cpu_baseline / --impl reference legs, and __graft_entry__.smoke().
 *
 * Every function cites the reference lines it restates (paths relative to
 * /root/reference).  Parity status:
 *   - bandwidth rules, K-fold split, kNN batches, CV score formula: pinned by
 *     running the reference's own python/kdsource/kde.py in place
 *     (tests/golden/make_golden.py, fixtures under tests/golden/).
 *   - metric perturbation maps and uniform->deviate maps: pinned by calling
 *     the reference's own C functions compiled in place (oracle/_ref).
 *   - the Gaussian kernel sum itself restates KDEpy's published formula
 *     (KDEpy>=1.1.12, pinned in pyproject.toml:13; source not in the tree):
 *     PARITY UNPINNED against KDEpy itself; the formula is cross-checked
 *     against scikit-learn's exact weighted KernelDensity and scipy's
 *     gaussian_kde (tests/test_oracle.py::test_kde_sum_vs_independent_libraries)
 *     and through analytic checks.
 *   - KDSource.fit / evaluate / plot marginals around that sum: pinned by the
 *     reference's own kdsource.py + plist.py run in place (ref_kdsource.npz).
 *   - index pick by prefix-sum + binary search and the Philox counter layout
 *     are this project's own contract (the reference walks the list
 *     sequentially, clib/src/kdsource.c:282-330); agreement with the
 *     reference sampler is statistical (KS / chi-square tests).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.1415926535897932384626433832795028841971694

/* ------------------------------------------------------------------ */
/* A.  Python-side hot path                                            */
/* ------------------------------------------------------------------ */

/* bw_silv: python/kdsource/kde.py:18-38 */
double orc_bw_silv(int dim, double n_eff) {
  return pow(4.0 / (2.0 + dim), 1.0 / (4.0 + dim)) *
         pow(n_eff, -1.0 / (4.0 + dim));
}

/*
 * Gaussian kernel sum, the arithmetic of KDEpy TreeKDE/NaiveKDE.evaluate as
 * called from python/kdsource/kdsource.py:239 and kde.py:146-151:
 *   out[m] = sum_i wn_i (2 pi)^(-d/2) h_i^-d exp(-|q_m - x_i|^2 / (2 h_i^2))
 * wn = weights normalised to sum 1 by the caller (KDEpy fit()).
 * cutoff > 0: hard support radius (sources with |q-x| > cutoff are skipped),
 * the deterministic stand-in for TreeKDE's ball query.
 */
void orc_kde_sum(const double *data, const double *wn, const double *bw,
                 int64_t N, int d, const double *pts, int64_t M, double cutoff,
                 double *out, int nthreads) {
  const double lognorm = -0.5 * d * log(2.0 * ORC_PI);
  double *coef = (double *)malloc(sizeof(double) * (size_t)N);
  double *ih2 = (double *)malloc(sizeof(double) * (size_t)N);
  for (int64_t i = 0; i < N; i++) {
    coef[i] = wn[i] * exp(lognorm - d * log(bw[i]));
    ih2[i] = 0.5 / (bw[i] * bw[i]);
  }
  const double c2 = cutoff > 0 ? cutoff * cutoff : INFINITY;
#ifdef _OPENMP
  if (nthreads > 0)
    omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic, 16)
  for (int64_t m = 0; m < M; m++) {
    const double *q = pts + m * d;
    double s = 0.0;
    for (int64_t i = 0; i < N; i++) {
      const double *x = data + i * d;
      double r2 = 0.0;
      for (int k = 0; k < d; k++) {
        double t = q[k] - x[k];
        r2 += t * t;
      }
      if (r2 <= c2)
        s += coef[i] * exp(-r2 * ih2[i]);
    }
    out[m] = s;
  }
  free(coef);
  free(ih2);
}

/* sklearn KFold(n_splits=K) without shuffle (used at kde.py:134-136):
 * contiguous folds, the first N%K folds one element longer. */
void orc_kfold_bounds(int64_t N, int K, int64_t *start /* K+1 */) {
  int64_t base = N / K, rem = N % K, s = 0;
  for (int f = 0; f < K; f++) {
    start[f] = s;
    s += base + (f < rem ? 1 : 0);
  }
  start[K] = s;
}

/*
 * _kde_cv_score for G bandwidth factors at once: kde.py:106-153 evaluated for
 * bw_grid[g] = grid[g] * seed (kde.py:204).  For fold f with test set T and
 * train set R:
 *    S_i   = sum_{j in R} (w_j / sum_R w) K_{h_j}(x_i - x_j)
 *    score = mean_f mean_{i in T_f} ( w_i * ln S_i )
 * seed has one value per particle (a scalar seed is broadcast by the caller).
 * weights == NULL follows the unweighted branch (kde.py:150-151).
 */
void orc_mlcv_scores(const double *data, const double *weights,
                     const double *seed, int64_t N, int d, const double *grid,
                     int G, int K, double *scores, int nthreads) {
  int64_t *fs = (int64_t *)malloc(sizeof(int64_t) * (size_t)(K + 1));
  orc_kfold_bounds(N, K, fs);
  const double lognorm = -0.5 * d * log(2.0 * ORC_PI);
  double *wtrain = (double *)malloc(sizeof(double) * (size_t)K);
  double wtot = 0.0;
  for (int64_t i = 0; i < N; i++)
    wtot += weights ? weights[i] : 1.0;
  for (int f = 0; f < K; f++) {
    /* sum of train weights, accumulated over the train set itself */
    double s = 0.0;
    for (int64_t i = 0; i < fs[f]; i++)
      s += weights ? weights[i] : 1.0;
    for (int64_t i = fs[f + 1]; i < N; i++)
      s += weights ? weights[i] : 1.0;
    wtrain[f] = s;
  }
  (void)wtot;
  double *acc = (double *)calloc((size_t)G * (size_t)K, sizeof(double));
#ifdef _OPENMP
  if (nthreads > 0)
    omp_set_num_threads(nthreads);
#endif
  for (int f = 0; f < K; f++) {
    const int64_t t0 = fs[f], t1 = fs[f + 1];
#pragma omp parallel
    {
      double *S = (double *)malloc(sizeof(double) * (size_t)G);
      double *loc = (double *)calloc((size_t)G, sizeof(double));
#pragma omp for schedule(dynamic, 8)
      for (int64_t i = t0; i < t1; i++) {
        for (int g = 0; g < G; g++)
          S[g] = 0.0;
        const double *q = data + i * d;
        for (int64_t j = 0; j < N; j++) {
          if (j >= t0 && j < t1)
            continue;
          const double *x = data + j * d;
          double r2 = 0.0;
          for (int k = 0; k < d; k++) {
            double t = q[k] - x[k];
            r2 += t * t;
          }
          const double wj = (weights ? weights[j] : 1.0) / wtrain[f];
          for (int g = 0; g < G; g++) {
            const double h = grid[g] * seed[j];
            S[g] += wj * exp(lognorm - d * log(h) - r2 / (2.0 * h * h));
          }
        }
        const double wi = weights ? weights[i] : 1.0;
        for (int g = 0; g < G; g++)
          loc[g] += wi * log(S[g]);
      }
#pragma omp critical
      for (int g = 0; g < G; g++)
        acc[(size_t)g * K + f] += loc[g];
      free(S);
      free(loc);
    }
  }
  for (int g = 0; g < G; g++) {
    double s = 0.0;
    for (int f = 0; f < K; f++)
      s += acc[(size_t)g * K + f] / (double)(fs[f + 1] - fs[f]);
    scores[g] = s / K;
  }
  free(acc);
  free(wtrain);
  free(fs);
}

/*
 * Brute-force kNN inside contiguous batches, the arithmetic of
 * NearestNeighbors(n_neighbors=k+1).fit(b).kneighbors(b) at kde.py:93-97.
 * Distances are sqrt(sum_j (a_j - b_j)^2) accumulated for j = 0..d-1 (the form
 * sklearn's kd-tree reproduces bit for bit); neighbours
 * are ordered by (distance, index) so ties are deterministic.
 * batch_off has nb+1 entries (np.array_split boundaries, kde.py:91).
 */
void orc_knn(const double *data, int64_t N, int d, const int64_t *batch_off,
             int nb, int kk, double *dist, int64_t *idx, int nthreads) {
  (void)N;
#ifdef _OPENMP
  if (nthreads > 0)
    omp_set_num_threads(nthreads);
#endif
  for (int b = 0; b < nb; b++) {
    const int64_t b0 = batch_off[b], b1 = batch_off[b + 1];
#pragma omp parallel for schedule(dynamic, 32)
    for (int64_t i = b0; i < b1; i++) {
      double *bd = dist + i * kk;
      int64_t *bi = idx + i * kk;
      int cnt = 0;
      const double *q = data + i * d;
      for (int64_t j = b0; j < b1; j++) {
        const double *x = data + j * d;
        double r2 = 0.0;
        for (int k = 0; k < d; k++) {
          double t = q[k] - x[k];
          r2 += t * t;
        }
        /* insertion into the sorted list of the kk best (r2, j) pairs */
        if (cnt == kk && !(r2 < bd[kk - 1]))
          continue; /* j increases, so equal r2 never displaces */
        int p = cnt < kk ? cnt : kk - 1;
        while (p > 0 && bd[p - 1] > r2) {
          bd[p] = bd[p - 1];
          bi[p] = bi[p - 1];
          p--;
        }
        bd[p] = r2;
        bi[p] = j - b0;
        if (cnt < kk)
          cnt++;
      }
      for (int p = 0; p < kk; p++)
        bd[p] = p < cnt ? sqrt(bd[p]) : INFINITY;
      for (int p = cnt; p < kk; p++)
        bi[p] = -1;
    }
  }
}

/* ------------------------------------------------------------------ */
/* B.  C-side hot path: uniform -> deviate maps, metric perturbations  */
/* ------------------------------------------------------------------ */

/* Philox4x32-10 (Salmon et al., SC'11), the counter-based generator named by
 * north_star.  ctr = (c0,c1,c2,c3), key = (k0,k1). */
static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
  const uint32_t n3 = (uint32_t)p0;
  c[0] = n0;
  c[1] = n1;
  c[2] = n2;
  c[3] = n3;
}

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2],
                       uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  uint32_t k[2] = {key[0], key[1]};
  for (int r = 0; r < 10; r++) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  memcpy(out, c, sizeof(uint32_t) * 4);
}

/*
 * Per-particle uniform stream.  Two sources:
 *  - Philox: counter (p_lo, p_hi, block, 0), key = seed.  Block 0 words 0,1
 *    form the 64-bit pick number; every later word w yields one uniform
 *    u = ((w >> 9) + 0.5) * 2^-23 (exact in fp32 and fp64, never 0 or 1).
 *  - external doubles: U[p*slots + (n % slots)], slot 0 is the pick.
 */
typedef struct {
  int external;
  const double *ext;
  int slots;
  int next;
  uint32_t key[2];
  uint32_t ctr[4];
  uint32_t buf[4];
  int have; /* words left in buf */
  /* uniform-consumption contract of the counter-based mode (include/kdsb200.h,
   * kdsb200_resample): 1 = the reference's polar method, 2 = Box-Muller pairs */
  int contract;
  int have_spare;
  double spare;
} orc_stream;

static void stream_init(orc_stream *s, uint64_t seed, uint64_t p,
                        const double *ext, int slots) {
  memset(s, 0, sizeof(*s));
  if (ext) {
    s->external = 1;
    s->ext = ext + (size_t)p * (size_t)slots;
    s->slots = slots;
    s->next = 0;
  } else {
    s->key[0] = (uint32_t)seed;
    s->key[1] = (uint32_t)(seed >> 32);
    s->ctr[0] = (uint32_t)p;
    s->ctr[1] = (uint32_t)(p >> 32);
    s->ctr[2] = 0;
    s->ctr[3] = 0;
    orc_philox4x32_10(s->ctr, s->key, s->buf);
    s->have = 4;
  }
}

static uint32_t stream_word(orc_stream *s) {
  if (s->have == 0) {
    s->ctr[2] += 1;
    orc_philox4x32_10(s->ctr, s->key, s->buf);
    s->have = 4;
  }
  uint32_t w = s->buf[4 - s->have];
  s->have--;
  return w;
}

/* index pick: T in [0,total), first i with cdf[i] > T */
static uint64_t stream_pick_target(orc_stream *s, uint64_t total) {
  if (s->external) {
    double u = s->ext[s->next % s->slots];
    s->next++;
    uint64_t T = (uint64_t)(u * (double)total);
    if (T >= total)
      T = total - 1;
    return T;
  }
  uint64_t hi = stream_word(s), lo = stream_word(s);
  uint64_t R = (hi << 32) | lo;
  return (uint64_t)(((unsigned __int128)R * (unsigned __int128)total) >> 64);
}

static double stream_uniform(orc_stream *s) {
  if (s->external) {
    double u = s->ext[s->next % s->slots];
    s->next++;
    return u;
  }
  uint32_t w = stream_word(s);
  return ((double)(w >> 9) + 0.5) * (1.0 / 8388608.0);
}

/* kds_rand_norm (Marsaglia polar, second value discarded):
 * clib/src/utils.c:9-29 */
static double rand_norm(orc_stream *s) {
  if (!s->external && s->contract == 2) {
    /* contract 2 (this library's own; no counterpart in the reference): Box-Muller, both
     * values of a pair used, two uniforms per pair, no rejection */
    if (s->have_spare) {
      s->have_spare = 0;
      return s->spare;
    }
    const double u1 = stream_uniform(s), u2 = stream_uniform(s);
    const double r = sqrt(-2.0 * log(u1));
    const double a = (2.0 * ORC_PI) * u2;
    s->spare = r * sin(a);
    s->have_spare = 1;
    return r * cos(a);
  }
  double t, g1, g2;
  do {
    g1 = 2.0 * stream_uniform(s) - 1.0;
    g2 = 2.0 * stream_uniform(s) - 1.0;
    t = g1 * g1 + g2 * g2;
  } while (t >= 1.0 || t == 0.0);
  t = sqrt((-2.0 * log(t)) / t);
  return g1 * t;
}

/* kds_rand_epan: clib/src/utils.c:32-46 */
static double rand_epan(orc_stream *s) {
  double x1 = stream_uniform(s) * 2.0 - 1.0;
  double x2 = stream_uniform(s) * 2.0 - 1.0;
  double x3 = stream_uniform(s) * 2.0 - 1.0;
  if (fabs(x3) >= fabs(x2) && fabs(x3) >= fabs(x1))
    return x2;
  return x3;
}

/* kds_rand_box: clib/src/utils.c:49-52 */
static double rand_box(orc_stream *s) { return stream_uniform(s) * 2.0 - 1.0; }

/* kds_rand_type: clib/src/utils.c:55-67 (unknown kernel -> 0) */
static double rand_type(orc_stream *s, char kernel) {
  if (kernel == 'g')
    return rand_norm(s);
  if (kernel == 'e')
    return rand_epan(s);
  if (kernel == 'b')
    return rand_box(s);
  return 0.0;
}

/* rotv (Rodrigues, axis-angle): clib/src/utils.c:82-103 */
static void rotv(double *v, const double *rv, int inverse) {
  double theta = sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
  if (theta == 0)
    return;
  if (inverse)
    theta = -theta;
  double kdotv = rv[0] * v[0] + rv[1] * v[1] + rv[2] * v[2];
  double kx[3] = {rv[1] * v[2] - rv[2] * v[1], rv[2] * v[0] - rv[0] * v[2],
                  rv[0] * v[1] - rv[1] * v[0]};
  double out[3];
  for (int i = 0; i < 3; i++)
    out[i] = v[i] * cos(theta) + (kx[i] / fabs(theta)) * sin(theta) +
             rv[i] * kdotv / (theta * theta) * (1 - cos(theta));
  for (int i = 0; i < 3; i++)
    v[i] = out[i];
}

/* metric ids: order of _metric_names, clib/src/geom.c:13-16 */
enum {
  M_ENERGY = 0,
  M_LETHARGY,
  M_WAVELENGTH,
  M_VOL,
  M_SURFXY,
  M_SURFR,
  M_SURFR2,
  M_SURFCIRCLE,
  M_GUIDE,
  M_ISOTROP,
  M_POLAR,
  M_POLARMU,
  M_TIME,
  M_DECADE
};

typedef struct {
  int type;
  int dim;
  float scaling[6]; /* float, as Metric_create casts: geom.c:35-38 */
  int nps;
  double params[8];
} orc_metric;

typedef struct {
  int order;
  orc_metric ms[8];
  char kernel;
  int has_trasl, has_rot;
  double trasl[3], rot[3];
} orc_geom;

/* particle: ekin, pos[3], dir[3], time */
typedef struct {
  double ekin, pos[3], dir[3], time;
} orc_part;

static void reflect(double *x, double lo, double hi) {
  /* geom.c:242-247 and siblings */
  while ((*x < lo) || (*x > hi)) {
    if (*x < lo)
      *x += 2 * (lo - *x);
    else
      *x -= 2 * (*x - hi);
  }
}

/* _vMF_perturb: clib/src/geom.c:557-573 */
static void vmf_perturb(orc_stream *s, double bw, double *dx, double *dy,
                        double *dz) {
  double xi = stream_uniform(s);
  double w = 1;
  w += bw * bw * log(xi + (1 - xi) * exp(-2 / (bw * bw)));
  double phi = 2. * ORC_PI * stream_uniform(s);
  double uv = sqrt(1 - w * w), u = uv * cos(phi), v = uv * sin(phi);
  double dx0 = *dx, dy0 = *dy, dz0 = *dz;
  if (dz0 > 0) {
    *dx = u * dz0 + w * dx0 - (v * dx0 - u * dy0) * dy0 / (1 + dz0);
    *dy = v * dz0 + w * dy0 + (v * dx0 - u * dy0) * dx0 / (1 + dz0);
  } else {
    *dx = u * dz0 + w * dx0 + (v * dx0 - u * dy0) * dy0 / (1 - dz0);
    *dy = v * dz0 + w * dy0 - (v * dx0 - u * dy0) * dx0 / (1 - dz0);
  }
  *dz = w * dz0 - u * dx0 - v * dy0;
}

static void guide_perturb(orc_stream *s, const orc_metric *m, orc_part *p,
                          double bw, char kernel);

/* one metric's perturbation; geom.c:192-650 */
static void metric_perturb(orc_stream *s, const orc_metric *m, orc_part *p,
                           double bw, char kernel) {
  switch (m->type) {
  case M_ENERGY: /* E_perturb geom.c:192-198 */
    p->ekin += bw * m->scaling[0] * rand_type(s, kernel);
    if (p->ekin < 0)
      p->ekin *= -1;
    break;
  case M_LETHARGY: { /* Let_perturb geom.c:200-210 */
    double E = p->ekin;
    E *= exp(bw * m->scaling[0] * rand_type(s, kernel));
    while (E > m->params[0]) {
      E = p->ekin;
      E *= exp(bw * m->scaling[0] * rand_type(s, kernel));
    }
    p->ekin = E;
  } break;
  case M_WAVELENGTH: { /* wl_perturb geom.c:212-221 */
    double wl = 9.045 / sqrt(p->ekin * 1e9);
    double wl2 = wl + bw * m->scaling[0] * rand_type(s, kernel);
    p->ekin = 81.82e-9 / (wl2 * wl2);
    if (p->ekin < 0)
      p->ekin *= -1;
  } break;
  case M_TIME: /* t_perturb geom.c:223-227 */
    p->time += bw * m->scaling[0] * rand_type(s, kernel);
    break;
  case M_DECADE: /* Dec_perturb geom.c:228-233 */
    p->time *= pow(10, bw * m->scaling[0] * rand_type(s, kernel));
    break;
  case M_VOL: /* Vol_perturb geom.c:235-265 */
    for (int a = 0; a < 3; a++) {
      p->pos[a] += bw * m->scaling[a] * rand_type(s, kernel);
      reflect(&p->pos[a], m->params[2 * a], m->params[2 * a + 1]);
    }
    break;
  case M_SURFXY: /* SurfXY_perturb geom.c:266-287 */
    for (int a = 0; a < 2; a++) {
      p->pos[a] += bw * m->scaling[a] * rand_type(s, kernel);
      reflect(&p->pos[a], m->params[2 * a], m->params[2 * a + 1]);
    }
    break;
  case M_SURFR: { /* SurfR_perturb geom.c:288-319 */
    double rho = sqrt(p->pos[0] * p->pos[0] + p->pos[1] * p->pos[1]);
    double psi = atan2(p->pos[1], p->pos[0]);
    double rho2 = rho + bw * m->scaling[0] * rand_type(s, kernel);
    double psi2 =
        psi + bw * m->scaling[1] * ORC_PI / 180 * rand_type(s, kernel);
    reflect(&rho2, m->params[0], m->params[1]);
    reflect(&psi2, m->params[2], m->params[3]);
    p->pos[0] = rho2 * cos(psi2);
    p->pos[1] = rho2 * sin(psi2);
  } break;
  case M_SURFR2: { /* SurfR2_perturb geom.c:320-354 */
    double rmin = m->params[0] * m->params[0];
    double rmax = m->params[1] * m->params[1];
    double rho = p->pos[0] * p->pos[0] + p->pos[1] * p->pos[1];
    double psi = atan2(p->pos[1], p->pos[0]);
    double rho2 = rho + bw * m->scaling[0] * rand_type(s, kernel);
    double psi2 =
        psi + bw * m->scaling[1] * ORC_PI / 180 * rand_type(s, kernel);
    reflect(&rho2, rmin, rmax);
    reflect(&psi2, m->params[2], m->params[3]);
    p->pos[0] = sqrt(rho2) * cos(psi2);
    p->pos[1] = sqrt(rho2) * sin(psi2);
  } break;
  case M_SURFCIRCLE: { /* SurfCircle_perturb geom.c:355-385 */
    double x = p->pos[0] + bw * m->scaling[0] * rand_type(s, kernel);
    double y = p->pos[1] + bw * m->scaling[1] * rand_type(s, kernel);
    double rho2 = sqrt(x * x + y * y);
    double psi2 = atan2(y, x);
    reflect(&rho2, m->params[0], m->params[1]);
    reflect(&psi2, m->params[2], m->params[3]);
    p->pos[0] = rho2 * cos(psi2);
    p->pos[1] = rho2 * sin(psi2);
  } break;
  case M_GUIDE:
    guide_perturb(s, m, p, bw, kernel);
    break;
  case M_ISOTROP: { /* Isotrop_perturb geom.c:575-600 (kernel ignored) */
    double sc = bw * m->scaling[0];
    if (isinf(sc)) {
      p->dir[2] = -1 + 2. * stream_uniform(s);
      double dxy = sqrt(1 - p->dir[2] * p->dir[2]);
      double phi = 2. * ORC_PI * stream_uniform(s);
      p->dir[0] = dxy * cos(phi);
      p->dir[1] = dxy * sin(phi);
    } else if (sc > 0) {
      double dx = p->dir[0], dy = p->dir[1], dz = p->dir[2];
      vmf_perturb(s, sc, &p->dir[0], &p->dir[1], &p->dir[2]);
      int keepx = (int)m->params[0], keepy = (int)m->params[1],
          keepz = (int)m->params[2];
      if (keepx && (dx * p->dir[0] < 0))
        p->dir[0] *= -1;
      if (keepy && (dy * p->dir[1] < 0))
        p->dir[1] *= -1;
      if (keepz && (dz * p->dir[2] < 0))
        p->dir[2] *= -1;
    }
  } break;
  case M_POLAR: { /* Polar_perturb geom.c:602-620; the reflected theta is
                     computed and then overwritten, as in the reference */
    double theta = acos(p->dir[2]);
    double phi = atan2(p->dir[1], p->dir[0]);
    double theta2 =
        theta + bw * m->scaling[0] * ORC_PI / 180 * rand_type(s, kernel);
    theta = theta2;
    phi += bw * m->scaling[1] * ORC_PI / 180 * rand_type(s, kernel);
    p->dir[0] = sin(theta) * cos(phi);
    p->dir[1] = sin(theta) * sin(phi);
    p->dir[2] = cos(theta);
  } break;
  case M_POLARMU: { /* PolarMu_perturb geom.c:622-650 */
    double mu = p->dir[2];
    double phi = atan2(p->dir[1], p->dir[0]);
    double mu2 = mu + bw * m->scaling[0] * rand_type(s, kernel);
    if (mu >= 0) {
      while ((mu2 < 0) || (mu2 > 1)) {
        if (mu2 < 0)
          mu2 *= -1;
        if (mu2 > 1)
          mu2 -= 2 * (mu2 - 1);
      }
    } else {
      while ((mu2 < -1) || (mu2 > 0)) {
        if (mu2 < -1)
          mu2 += 2 * (-1 - mu2);
        if (mu2 > 0)
          mu2 *= -1;
      }
    }
    mu = mu2;
    phi += bw * m->scaling[1] * ORC_PI / 180 * rand_type(s, kernel);
    p->dir[0] = sqrt(1 - mu * mu) * cos(phi);
    p->dir[1] = sqrt(1 - mu * mu) * sin(phi);
    p->dir[2] = mu;
  } break;
  default:
    break;
  }
}

/* Guide_perturb: clib/src/geom.c:386-555 */
static void guide_perturb(orc_stream *s, const orc_metric *m, orc_part *p,
                          double bw, char kernel) {
  double x = p->pos[0], y = p->pos[1], z = p->pos[2];
  double dx = p->dir[0], dy = p->dir[1], dz = p->dir[2], dx2, dz2;
  const double xw = m->params[0], yh = m->params[1], zmax = m->params[2],
               rc = m->params[3];
  double t = 0, mu = 0, mu2 = 0, phi = 0;
  int mirror;
  if (rc != 0) {
    double r = sqrt((rc + x) * (rc + x) + z * z);
    x = copysign(1, rc) * r - rc;
    z = fabs(rc) * asin(z / r);
    dx2 = dx;
    dz2 = dz;
    dx = dx2 * cos(z / rc) + dz2 * sin(z / rc);
    dz = -dx2 * sin(z / rc) + dz2 * cos(z / rc);
  }
  if ((y / yh > -x / xw) && (y / yh < x / xw))
    mirror = 0;
  else if ((y / yh > x / xw) && (y / yh > -x / xw))
    mirror = 1;
  else if ((y / yh < -x / xw) && (y / yh > x / xw))
    mirror = 2;
  else
    mirror = 3;
  switch (mirror) {
  case 0:
    t = 0.5 * yh + y;
    mu = dx;
    phi = atan2(-dy, dz);
    break;
  case 1:
    t = 1.0 * yh + 0.5 * xw - x;
    mu = dy;
    phi = atan2(dx, dz);
    break;
  case 2:
    t = 1.5 * yh + 1.0 * xw - y;
    mu = -dx;
    phi = atan2(dy, dz);
    break;
  case 3:
    t = 2.0 * yh + 1.5 * xw + x;
    mu = -dy;
    phi = atan2(-dx, dz);
    break;
  }
  z += bw * m->scaling[0] * rand_type(s, kernel);
  while ((z < 0) || (z > zmax)) {
    if (z < 0)
      z *= -1;
    else
      z -= 2 * (z - zmax);
  }
  t += bw * m->scaling[1] * rand_type(s, kernel);
  {
    double lo = 0, hi = 0;
    switch (mirror) {
    case 0:
      lo = 0;
      hi = yh;
      break;
    case 1:
      lo = yh;
      hi = yh + xw;
      break;
    case 2:
      lo = yh + xw;
      hi = 2 * yh + xw;
      break;
    case 3:
      lo = 2 * yh + xw;
      hi = 2 * yh + 2 * xw;
      break;
    }
    if (mirror == 0) {
      while ((t < 0) || (t > yh)) {
        if (t < 0)
          t *= -1;
        else
          t -= 2 * (t - yh);
      }
    } else {
      reflect(&t, lo, hi);
    }
  }
  if (isinf(m->scaling[2]))
    mu = -1 + 2. * stream_uniform(s);
  else {
    mu2 = mu + bw * m->scaling[2] * rand_type(s, kernel);
    if (mu >= 0) {
      while ((mu2 < 0) || (mu2 > 1)) {
        if (mu2 < 0)
          mu2 *= -1;
        else
          mu2 -= 2 * (mu2 - 1);
      }
    } else {
      while ((mu2 < -1) || (mu2 > 0)) {
        if (mu2 < -1)
          mu2 += 2 * (-1 - mu2);
        else
          mu2 *= -1;
      }
    }
    mu = mu2;
  }
  if (isinf(m->scaling[3]))
    phi = 2. * ORC_PI * stream_uniform(s);
  else
    phi += bw * m->scaling[3] * ORC_PI / 180 * rand_type(s, kernel);
  switch (mirror) {
  case 0:
    x = xw / 2;
    y = t - 0.5 * yh;
    dx = mu;
    dz = sqrt(1 - mu * mu) * cos(phi);
    dy = -sqrt(1 - mu * mu) * sin(phi);
    break;
  case 1:
    y = yh / 2;
    x = -t + yh + 0.5 * xw;
    dy = mu;
    dz = sqrt(1 - mu * mu) * cos(phi);
    dx = sqrt(1 - mu * mu) * sin(phi);
    break;
  case 2:
    x = -xw / 2;
    y = -t + 1.5 * yh + xw;
    dx = -mu;
    dz = sqrt(1 - mu * mu) * cos(phi);
    dy = sqrt(1 - mu * mu) * sin(phi);
    break;
  case 3:
    y = -yh / 2;
    x = t - 2 * yh - 1.5 * xw;
    dy = -mu;
    dz = sqrt(1 - mu * mu) * cos(phi);
    dx = -sqrt(1 - mu * mu) * sin(phi);
    break;
  }
  if (rc != 0) {
    double r = (rc + x) * copysign(1, rc), ang = z / rc;
    x = copysign(1, rc) * r * cos(ang) - rc;
    z = r * sin(fabs(ang));
    dx2 = dx;
    dz2 = dz;
    dx = dx2 * cos(ang) - dz2 * sin(ang);
    dz = dx2 * sin(ang) + dz2 * cos(ang);
  }
  p->pos[0] = x;
  p->pos[1] = y;
  p->pos[2] = z;
  p->dir[0] = dx;
  p->dir[1] = dy;
  p->dir[2] = dz;
}

/* Geom_perturb: clib/src/geom.c:139-158 */
static void geom_perturb(orc_stream *s, const orc_geom *g, orc_part *p,
                         double bw) {
  if (g->has_trasl)
    for (int i = 0; i < 3; i++)
      p->pos[i] -= g->trasl[i];
  if (g->has_rot) {
    rotv(p->pos, g->rot, 1);
    rotv(p->dir, g->rot, 1);
  }
  for (int i = 0; i < g->order; i++)
    metric_perturb(s, &g->ms[i], p, bw, g->kernel);
  if (g->has_rot) {
    rotv(p->pos, g->rot, 0);
    rotv(p->dir, g->rot, 0);
  }
  if (g->has_trasl)
    for (int i = 0; i < 3; i++)
      p->pos[i] += g->trasl[i];
}

/* test hook: run one metric perturbation on one particle with a scripted
 * uniform stream (external doubles), so the maps can be pinned against the
 * reference's own compiled functions.  Returns uniforms consumed. */
int orc_metric_perturb_scripted(int type, int dim, const double *scaling,
                                int nps, const double *params, char kernel,
                                double bw, double *part8 /* in/out */,
                                const double *uniforms, int nuni) {
  orc_metric m;
  memset(&m, 0, sizeof(m));
  m.type = type;
  m.dim = dim;
  for (int i = 0; i < dim && i < 6; i++)
    m.scaling[i] = (float)scaling[i];
  m.nps = nps;
  for (int i = 0; i < nps && i < 8; i++)
    m.params[i] = params[i];
  orc_stream s;
  stream_init(&s, 0, 0, uniforms, nuni);
  orc_part p = {part8[0], {part8[1], part8[2], part8[3]},
                {part8[4], part8[5], part8[6]}, part8[7]};
  metric_perturb(&s, &m, &p, bw, kernel);
  part8[0] = p.ekin;
  for (int i = 0; i < 3; i++) {
    part8[1 + i] = p.pos[i];
    part8[4 + i] = p.dir[i];
  }
  part8[7] = p.time;
  return s.next;
}

/* scripted rotv / deviate hooks (pinned against utils.c) */
void orc_rotv(double *v, const double *rv, int inverse) { rotv(v, rv, inverse); }
double orc_rand_type_scripted(char kernel, const double *uniforms, int nuni,
                              int *consumed) {
  orc_stream s;
  stream_init(&s, 0, 0, uniforms, nuni);
  double v = rand_type(&s, kernel);
  *consumed = s.next;
  return v;
}

/* ------------------------------------------------------------------ */
/* C.  Resampling contract of this project                             */
/* ------------------------------------------------------------------ */

/* fixed-point weights: c_i = floor(w_i * 2^62 / sum(w)) (at least 1 for any
 * w_i > 0), cdf = inclusive prefix sum in uint64 (exactly associative, so a
 * parallel scan on the device reproduces it bit for bit).  sum_w is supplied by
 * the caller (numpy's np.sum of the weights on both sides). */
void orc_weights_cdf(const double *w, int64_t N, double sum_w, uint64_t *cdf) {
  const double scale = 4611686018427387904.0 / sum_w; /* 2^62 / sum */
  uint64_t acc = 0;
  for (int64_t i = 0; i < N; i++) {
    uint64_t c = (uint64_t)(w[i] * scale);
    if (c == 0 && w[i] > 0)
      c = 1;
    acc += c;
    cdf[i] = acc;
  }
}

static int64_t upper_bound_u64(const uint64_t *cdf, int64_t N, uint64_t T) {
  int64_t lo = 0, hi = N;
  while (lo < hi) {
    int64_t mid = lo + ((hi - lo) >> 1);
    if (cdf[mid] > T)
      hi = mid;
    else
      lo = mid + 1;
  }
  return lo < N ? lo : N - 1;
}

/* MCPL adaptive projection packing (libmcpl mcpl_unitvect_pack_adaptproj,
 * invoked through mcpl_add_particle at clib/src/resamplemcpl.c:42) */
static void pack36(const orc_part *p, double weight, int32_t pdg,
                   unsigned char *out) {
  float rec[8];
  double ax = fabs(p->dir[0]), ay = fabs(p->dir[1]), az = fabs(p->dir[2]);
  double p0, p1, sgnsrc;
  if (az < fmax(ax, ay)) {
    double invz = p->dir[2] != 0.0 ? 1.0 / p->dir[2] : INFINITY;
    if (ax >= ay) {
      p0 = invz;
      p1 = p->dir[1];
      sgnsrc = p->dir[0];
    } else {
      p0 = p->dir[0];
      p1 = invz;
      sgnsrc = p->dir[1];
    }
  } else {
    p0 = p->dir[0];
    p1 = p->dir[1];
    sgnsrc = p->dir[2];
  }
  rec[0] = (float)p->pos[0];
  rec[1] = (float)p->pos[1];
  rec[2] = (float)p->pos[2];
  rec[3] = (float)p0;
  rec[4] = (float)p1;
  rec[5] = (float)copysign(fabs(p->ekin), sgnsrc);
  rec[6] = (float)p->time;
  rec[7] = (float)weight;
  memcpy(out, rec, 32);
  memcpy(out + 32, &pdg, 4);
}

/*
 * Resample `count` particles with global output indices first..first+count-1.
 * src8: N x 8 doubles [ekin,x,y,z,ux,uy,uz,t] (valid particles only, PList
 * trasl/rot/x2z already applied as PList_get does, clib/src/plist.c:69-89);
 * bw: per-source bandwidth (NULL -> bw_scalar), the value Geom_next would
 * have loaded for that list entry (geom.c:160-177).
 * Outputs (each optional): 36-byte MCPL-3 records, picked index, full
 * precision particle (8 doubles).
 * Emitted weight is 1 (kdsource.c:325 with bias == NULL).
 */
void orc_resample(const double *src8, const uint64_t *cdf, const float *bw,
                  double bw_scalar, int64_t N, const orc_geom *g, int32_t pdg,
                  uint64_t seed, uint64_t first, uint64_t count,
                  const double *ext_uniforms, int slots, int perturb, int contract,
                  unsigned char *out36, int64_t *out_idx, double *out_parts) {
  const uint64_t total = cdf[N - 1];
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < (int64_t)count; n++) {
    const uint64_t p = first + (uint64_t)n;
    orc_stream s;
    stream_init(&s, seed, ext_uniforms ? (uint64_t)n : p, ext_uniforms, slots);
    s.contract = contract;
    const uint64_t T = stream_pick_target(&s, total);
    const int64_t i = upper_bound_u64(cdf, N, T);
    const double *r = src8 + i * 8;
    orc_part q = {r[0], {r[1], r[2], r[3]}, {r[4], r[5], r[6]}, r[7]};
    const double h = bw ? (double)bw[i] : bw_scalar;
    if (perturb)
      geom_perturb(&s, g, &q, h);
    if (out_idx)
      out_idx[n] = i;
    if (out_parts) {
      double *o = out_parts + n * 8;
      o[0] = q.ekin;
      for (int k = 0; k < 3; k++) {
        o[1 + k] = q.pos[k];
        o[4 + k] = q.dir[k];
      }
      o[7] = q.time;
    }
    if (out36)
      pack36(&q, 1.0, pdg, out36 + n * 36);
  }
}

size_t orc_sizeof_geom(void) { return sizeof(orc_geom); }
