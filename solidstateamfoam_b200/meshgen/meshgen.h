/* C interface of libssam_meshgen.so: synthetic meshes in OpenFOAM conventions and OpenFOAM case I/O. */
#ifndef SSAM_MESHGEN_H
#define SSAM_MESHGEN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef struct smg_mesh smg_mesh;
smg_mesh* smg_hex_block(int nx, int ny, int nz, const double* origin, const double* len, const double* grading,
                        const double* affine, const int32_t* sideNbrRank, int myRank);
smg_mesh* smg_poly_refined(int nbx, int nby, int nbz, const double* origin, const double* len, double r1, double d1,
                           double r2, double d2, const double* affine);
int smg_decompose(const smg_mesh* g, const int32_t* cellProc, int nProcs, smg_mesh** out);
void smg_free(smg_mesh* m);
void smg_sizes(const smg_mesh* m, int32_t* out4); /* nCells, nInternalFaces, nFaces, nPatches */
const int32_t* smg_owner(const smg_mesh*);
const int32_t* smg_neighbour(const smg_mesh*);
const int32_t* smg_patch_start(const smg_mesh*);
const int32_t* smg_patch_size(const smg_mesh*);
const int32_t* smg_patch_kind(const smg_mesh*);
const int32_t* smg_patch_nbr_rank(const smg_mesh*);
const double* smg_Sf(const smg_mesh*);
const double* smg_Cf(const smg_mesh*);
const double* smg_magSf(const smg_mesh*);
const double* smg_weights(const smg_mesh*);
const double* smg_delta_coeffs(const smg_mesh*);
const double* smg_V(const smg_mesh*);
const double* smg_C(const smg_mesh*);
const char* smg_patch_name(const smg_mesh*, int p);
/* OpenFOAM case I/O */
smg_mesh* smg_read_polymesh(const char* polyMeshDir); /* NULL on error, see smg_io_error() */
int smg_write_polymesh(const smg_mesh*, const char* polyMeshDir, int binary);
int smg_couple_processor_patches(smg_mesh** meshes, int n);
int smg_read_field(const smg_mesh*, const char* path, int ncomp, double* internal, double* boundary, int32_t* patchBc);
int smg_write_field(const smg_mesh*, const char* path, const char* name, int ncomp, const double* internal,
                    const double* boundary, const int32_t* patchBc, int binary);
const char* smg_io_error(void);
#ifdef __cplusplus
}
#endif
#endif
