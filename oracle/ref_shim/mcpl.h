/* Shim for <mcpl.h> (libmcpl is not installed here).  Declares only what the
 * reference's clib/src/{kdsource,geom,plist,utils}.c use; the functions are
 * implemented by ref_harness.c over an in-memory particle array.
 * TEST INFRASTRUCTURE (oracle/_ref build) -- not part of the product. */
#ifndef ORC_SHIM_MCPL_H
#define ORC_SHIM_MCPL_H
#include <stdint.h>
typedef struct {
  double ekin;
  double polarisation[3];
  double position[3];
  double direction[3];
  double time;
  double weight;
  int32_t pdgcode;
  uint32_t userflags;
} mcpl_particle_t;
typedef struct { void *internal; } mcpl_file_t;
typedef struct { void *internal; } mcpl_outfile_t;
mcpl_file_t mcpl_open_file(const char *filename);
uint64_t mcpl_hdr_nparticles(mcpl_file_t);
const mcpl_particle_t *mcpl_read(mcpl_file_t);
void mcpl_rewind(mcpl_file_t);
void mcpl_close_file(mcpl_file_t);
#endif
