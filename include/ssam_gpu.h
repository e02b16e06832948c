/*
 * ssam_gpu.h -- C-ABI of the B200-native constitutive + heat-source update
 *               (grad(U) -> strain rate -> viscoplastic nu -> two-phase mixture -> heat sources)
 *
 * Drop-in boundary for ONE hot path of kckincaid/solidStateAMFoam's chtMultiRegionInterFoam.
 * The reference has no FFI: the path sits behind OpenFOAM run-time-selectable C++ virtuals.
 * Every entry point below names the reference interface (file:line under /root/reference) it
 * replaces; INTEGRATION.md shows the OpenFOAM-side shim that forwards to it.
 *
 * Conventions
 *   - plain C, opaque handle, int status (0 = ok); no exceptions cross the boundary;
 *     ssam_last_error() returns the message of the last failing call on that context.
 *   - scalar = double (WM_PRECISION_OPTION=DP), label = int32 (WM_LABEL_SIZE=32).
 *   - host buffers are caller-owned and borrowed for the duration of the call; host layout is
 *     OpenFOAM's contiguous AoS Field<T>: vector = 3 doubles (x,y,z), tensor = 9 doubles row-major.
 *   - every field has an internal part (nCells) and a boundary part (nFaces-nInternalFaces values in
 *     boundary-face order = patch order), like a GeometricField.
 *   - one context per rank / GPU; calls on one context are serialised by the caller (as in OpenFOAM).
 *   - there is NO CPU fallback: every compute entry point launches sm_100a kernels or fails.
 */
#ifndef SSAM_GPU_H
#define SSAM_GPU_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define SSAM_API __attribute__((visibility("default")))
#else
#define SSAM_API
#endif

typedef struct ssam_ctx ssam_ctx;
typedef int ssam_status;

enum {
    SSAM_OK = 0,
    SSAM_ERR_INVALID = 1,   /* bad argument / call order                                   */
    SSAM_ERR_CUDA = 2,      /* CUDA runtime error (message has the cudaError string)       */
    SSAM_ERR_UNKNOWN_MODEL = 3, /* mirrors FatalError "Unknown ... model" (stickModelNew.C:46-54) */
    SSAM_ERR_MISSING_FIELD = 4, /* mirrors objectRegistry::lookupObject failure (constantSlipStick.C:72) */
    SSAM_ERR_COMM = 5,      /* NCCL failure / communicator missing                          */
    SSAM_ERR_UNITS = 6      /* mirrors "Unknown conductivity function units"
                               (incompressibleTwoPhaseThermalMixture.C:235-238,305-308)     */
};

/* ---- patches ---------------------------------------------------------------------------------- */
enum { SSAM_PATCH_GENERIC = 0, SSAM_PATCH_PROCESSOR = 1, SSAM_PATCH_EMPTY = 2 };

/* how snGrad(U) is formed on a non-coupled patch (gaussGrad boundary correction):
 * VALUE: deltaCoeffs*(U_b - U_P) (fixedValue family incl. constant/exponentialRotationalSlip);
 * ZERO_GRADIENT: 0. Processor patches ignore this. */
enum { SSAM_BC_VALUE = 0, SSAM_BC_ZERO_GRADIENT = 1 };

/* ---- mesh --------------------------------------------------------------------------------------
 * Replaces the fvMesh accessors the path reads: owner()/neighbour() and Sf()/weights()/V() inside
 * fvc::grad(U) (called at fluid/TEqn.H:3, stickModel.C:60, NortonHoff.C:62), C()/Cf()/magSf() and
 * patch faceCells in constantSlipStick.C:110-146, patch Cf in
 * constantRotationalSlipFvPatchVectorField.C:109-139.  Geometry is an input (never recomputed). */
typedef struct ssam_mesh_desc {
    int32_t nCells, nInternalFaces, nFaces, nPatches;
    const int32_t* owner;        /* [nFaces]                                                       */
    const int32_t* neighbour;    /* [nInternalFaces]                                               */
    const int32_t* patchStart;   /* [nPatches] first (global) face of the patch                    */
    const int32_t* patchSize;    /* [nPatches]                                                     */
    const int32_t* patchKind;    /* [nPatches] SSAM_PATCH_*                                        */
    const int32_t* patchNbrRank; /* [nPatches] neighbour rank of a processor patch, else -1        */
    const double* Sf;            /* [3*nFaces] face area vectors (owner->neighbour / outward)      */
    const double* weights;       /* [nFaces] linear weights: internal faces, then patch weights    */
    const double* V;             /* [nCells]                                                       */
    const double* C;             /* [3*nCells] cell centres                                        */
    const double* Cf_b;          /* [3*nBnd] boundary face centres                                 */
    const double* magSf_b;       /* [nBnd]                                                         */
    const double* deltaCoeffs_b; /* [nBnd]                                                         */
} ssam_mesh_desc;

/* ---- fields ------------------------------------------------------------------------------------
 * The registry names the reference looks up ("temp", "epsilonDotEq", "muEff", "sigmay", alphaName,
 * pName: SheppardWright.C:55,58; constantSlipStick.C:54-75) become field ids. */
enum {
    SSAM_F_U = 0,        /* vector, in                                                             */
    SSAM_F_T = 1,        /* "temp", in                                                             */
    SSAM_F_ALPHA1 = 2,   /* alpha.<phase1>, in                                                     */
    SSAM_F_P = 3,        /* pName, in (friction only)                                              */
    SSAM_F_EPSDOT = 4,   /* "epsilonDotEq" = sqrt(2/3)|symm(grad U)|  (fluid/TEqn.H:3)             */
    SSAM_F_NU1 = 5,      /* phase-1 viscosityModel nu_                                             */
    SSAM_F_NU2 = 6,      /* phase-2 viscosityModel nu_                                             */
    SSAM_F_SIGMAF = 7,   /* "sigmaf"                                                               */
    SSAM_F_SIGMAY = 8,   /* "sigmay"                                                               */
    SSAM_F_NU = 9,       /* mixture nu_ (incompressibleTwoPhaseThermalMixture.C:53)                */
    SSAM_F_MUEFF = 10,   /* "muEff" = thermo.mu() (fluid/TEqn.H:6)                                 */
    SSAM_F_RHO = 11,     /* thermo.rho()                                                           */
    SSAM_F_CP = 12,      /* thermo.cp()  (fluid/TEqn.H:12)                                         */
    SSAM_F_K = 13,       /* thermo.k()   (fluid/TEqn.H:15)                                         */
    SSAM_F_QVISC = 14,   /* friction.qvisc()                                                       */
    SSAM_F_QFRIC = 15,   /* friction.qfric()                                                       */
    SSAM_F_GRADU = 16,   /* tensor, fvc::grad(U) (debug / parity / calculateStrain.H)              */
    SSAM_F_STRAINRATE = 17, /* viscosityModel::strainRate() = sqrt(2)|symm(grad U)| (NortonHoff.C:62) */
    SSAM_F_RHO_SOLVER = 18, /* rho   == alpha1*rho1 + alpha2*rho2        (alphaEqnSubCycle.H:41)    */
    SSAM_F_RHOCP = 19,      /* rhoCp == alpha1*rho1*cp1 + alpha2*rho2*cp2 (alphaEqnSubCycle.H:42)   */
    /* fluid/calculateStrain.H post-processing */
    SSAM_F_PRGH = 20,       /* p_rgh, in                                                            */
    SSAM_F_Z = 21,          /* Z = epsilonDotEq*exp(Q/(R*temp))                 (calculateStrain.H:12) */
    SSAM_F_SIGMAF_PP = 22,  /* the solver's own "sigmaf" (createStressStrain.H:36, calculateStrain.H:13) */
    SSAM_F_DEPSEQ = 23,     /* dEpsilonEq = sqrt(4/3*J2)                        (calculateStrain.H:35) */
    SSAM_F_EPSEQ = 24,      /* epsilonEq, in/out: += dEpsilonEq                 (calculateStrain.H:40) */
    SSAM_F_SIGMAVM = 25,    /* von Mises stress                                 (calculateStrain.H:92) */
    /* interfaceProperties::correct() (called at
     * immiscibleIncompressibleTwoPhaseThermalMixture.H:80, consumed by alphaEqn.H:154 and UEqn.H:31) */
    SSAM_F_GRADALPHA = 26,  /* vector, fvc::grad(alpha1, "nHat")                                    */
    SSAM_F_CURVATURE = 27,  /* interfaceProperties::K_ = -fvc::div(nHatf_)                          */
    /* explicit part of turbulence.divDevRhoReff(rho, U) (fluid/UEqn.H:13; OpenFOAM linearViscousStress):
     * -fvc::div((rho*nuEff)*dev2(T(fvc::grad(U)))), laminar nuEff = thermo.nu()                      */
    SSAM_F_DIVDEVRHOREFF = 28, /* vector                                                              */
    SSAM_F_K1 = 29,         /* thermo.k1()  (incompressibleTwoPhaseThermalMixture.C:205-273)          */
    SSAM_F_K2 = 30,         /* thermo.k2()  (:276-343)                                                */
    SSAM_F_COUNT = 31
};

/* ---- model parameters (what the reference reads from dictionaries) ---------- */
enum { SSAM_VISC_NEWTONIAN = 0, SSAM_VISC_SHEPPARD_WRIGHT = 1, SSAM_VISC_FKS = 2, SSAM_VISC_NORTON_HOFF = 3 };

typedef struct ssam_viscosity_params {
    int32_t model; /* SSAM_VISC_* */
    int32_t _pad;
    /* Newtonian: nu */
    double nu;
    /* SheppardWright (SheppardWright.C:132-138; limiters :104-106) */
    double sigmaR, Q, R, A, n;
    /* FKS (FKS.C:199-213) */
    double a1, a2, b1, b2, b3, e0, c1, c2, Tm, Ta;
    /* NortonHoff (NortonHoff.C:82-88) */
    double mu0, m, strainRateEffMin;
    /* common */
    double rho, nuMin, nuMax, eDotMin;
} ssam_viscosity_params;

enum { SSAM_K_CONSTANT = 0, SSAM_K_CUBIC = 1 };
enum { SSAM_UNITS_KELVIN = 0, SSAM_UNITS_CELSIUS = 1 };

typedef struct ssam_conductivity {
    int32_t kModel; /* SSAM_K_*  (kModel word, incompressibleTwoPhaseThermalMixture.C:108-109) */
    int32_t units;  /* SSAM_UNITS_* (:223-239); any other value -> SSAM_ERR_UNITS             */
    double k;       /* constant model (:261)                                                   */
    double k0, k1, k2, k3, Tmin, Tmax; /* cubic model (:215-221)                               */
} ssam_conductivity;

typedef struct ssam_mixture_params {
    double rho1, rho2, cp1, cp2; /* incompressibleTwoPhaseThermalMixture.C:100-106 */
    ssam_conductivity k1, k2;
} ssam_mixture_params;

enum { SSAM_STICK_CONSTANT = 0, SSAM_STICK_EXPONENTIAL = 1 };

typedef struct ssam_stick_params {
    int32_t model;     /* SSAM_STICK_*: constantSlipStick / exponentialSlipStick                  */
    int32_t fricPatch; /* patch index of `fricPatch`; <0 skips qfric (NortonHoff case, SURVEY D-1).
                        * Multi-rank: the first friction evaluation per mesh and patch all-reduces the patch centroid
                        * (constantSlipStick.C:133-135), so EVERY rank must pass the same index, including ranks that
                        * hold none of the patch's faces (decomposePar keeps the patch there, empty) - exactly as every
                        * OpenFOAM rank calls reduce().  A rank passing <0 while others name a patch stalls them. */
    double betaTQ;
    double omega[3];
    double delta, muf;                    /* constantSlipStick.C:165-168                          */
    double omega0, delta0, Rs, mu0, lambda; /* exponentialSlipStick.C ctor                         */
} ssam_stick_params;

enum { SSAM_ROTSLIP_CONSTANT = 0, SSAM_ROTSLIP_EXPONENTIAL = 1 };

typedef struct ssam_rotslip_params {
    int32_t model; /* SSAM_ROTSLIP_* */
    int32_t patch;
    double omega[3];
    double Vt[3];
    double delta;              /* constantRotationalSlip   */
    double omega0, delta0, R0; /* exponentialRotationalSlip */
} ssam_rotslip_params;

/* fluid/calculateStrain.H: constants looked up in the phase-1 dictionary (:6-10) + runTime.deltaT() */
typedef struct ssam_stress_strain_params {
    double Q, R, sigmaR, A, n;
    double deltaT;
} ssam_stress_strain_params;

/* everything one fused update needs (benchmark path) */
typedef struct ssam_update_params {
    ssam_viscosity_params visc1, visc2;
    ssam_mixture_params mix;
    ssam_stick_params stick;
} ssam_update_params;

/* ---- lifecycle --------------------------------------------------------------------------------- */
SSAM_API ssam_status ssam_create(ssam_ctx** ctx, int device);
SSAM_API ssam_status ssam_destroy(ssam_ctx* ctx);
SSAM_API const char* ssam_last_error(const ssam_ctx* ctx); /* ctx may be NULL (creation errors) */
SSAM_API ssam_status ssam_sync(ssam_ctx* ctx);
/* Launch on the caller's CUDA stream (cudaStream_t as void*; NULL = the context's own stream) and
 * choose whether compute calls block until done (default 1, OpenFOAM semantics). */
SSAM_API ssam_status ssam_set_stream(ssam_ctx* ctx, void* cuda_stream);
SSAM_API ssam_status ssam_set_blocking(ssam_ctx* ctx, int blocking);
SSAM_API int ssam_abi_version(void);
/* kernels launched by this context since creation (bench.py's gpu_launches) */
SSAM_API int64_t ssam_launch_count(const ssam_ctx* ctx);
/* Fused cell-kernel variant: 1 (default) = tiled, one CTA per tile: inputs of 256 consecutive cells staged in shared
 * memory by 1-D bulk async copies; 2 = persistent self-pipelined tiles (the bulk copies of a CTA's next tile land
 * while it runs the FP64 chain of the current one; dynamic tile scheduling) -- measured slower than 1 on B200, kept
 * as an option; 0 = one thread per cell with direct global loads.  Results are bit-identical. */
SSAM_API ssam_status ssam_set_cell_kernel(ssam_ctx* ctx, int variant);
/* Cell layout of the device planes, to be chosen before ssam_mesh_set:
 *   SSAM_LAYOUT_CALLER  (0) the caller's cell numbering;
 *   SSAM_LAYOUT_AUTO    (1, default) tile-compact when it removes at least one out-of-tile neighbour per cell;
 *   SSAM_LAYOUT_COMPACT (2) always tile-compact.
 * Tile-compact: runs of 64 consecutive cells are regrouped so that 256 consecutive device cells form a compact
 * brick of the mesh.  ssam_field_upload/download and ssam_update_fused_host apply the permutation, every
 * addressing table above stays in the caller's numbering, per-cell face order (and so every result bit) is
 * unchanged; only ssam_field_device_ptr exposes the device order (SSAM_ADDR_CELL_ORDER tells which cell is where).
 * ssam_update_fused_host is chunk-pipelined in both layouts; it overlaps best with SSAM_LAYOUT_CALLER. */
enum { SSAM_LAYOUT_CALLER = 0, SSAM_LAYOUT_AUTO = 1, SSAM_LAYOUT_COMPACT = 2 };
SSAM_API ssam_status ssam_set_cell_layout(ssam_ctx* ctx, int mode);
/* Host-only helper (no context, no GPU): the tile-compact order ssam_mesh_set would compute for this mesh
 * (order[i] = caller cell held at device position i), the out-of-tile neighbours per cell before and after, and
 * whether SSAM_LAYOUT_AUTO would adopt it. */
SSAM_API ssam_status ssam_compact_cell_order(const ssam_mesh_desc* mesh, int32_t* order, double* remote_caller,
                                             double* remote_compact, int* would_use);
/* 1 when the current mesh is held in tile-compact order */
SSAM_API int ssam_cell_layout_is_compact(const ssam_ctx* ctx);
/* Per-kernel timing for the roofline figure: when enabled, every fused update brackets its cell
 * kernel (the dominant kernel) with CUDA events on the launching stream.  ssam_profile_get
 * synchronises, returns the summed cell-kernel milliseconds and the number of launches since the
 * last call, and resets both. */
SSAM_API ssam_status ssam_profile_enable(ssam_ctx* ctx, int enable);
SSAM_API ssam_status ssam_profile_get(ssam_ctx* ctx, double* cells_ms_total, int64_t* n_launches);

/* ---- mesh + integer addressing ------------------------------------------------------------------ */
SSAM_API ssam_status ssam_mesh_set(ssam_ctx* ctx, const ssam_mesh_desc* mesh);
/* per-patch snGrad kind of U, [nPatches] of SSAM_BC_*; default SSAM_BC_VALUE */
SSAM_API ssam_status ssam_patch_u_bc_set(ssam_ctx* ctx, const int32_t* kinds);

/* Derived addressing, built on the GPU by ssam_mesh_set; getters exist for the bit-exact integer
 * parity checks ("integer work ... must be bit-exact", BASELINE.json north_star). */
enum {
    SSAM_ADDR_OWNER_START = 0,  /* [nCells+1] first owned internal face of each cell (lduAddressing ownerStart) */
    SSAM_ADDR_EXTRA_START = 1,  /* [nCells+1] offsets into the extra list                                     */
    SSAM_ADDR_EXTRA_FACE = 2,   /* [nExtra] neighbour-side internal faces (ascending) then boundary faces
                                   (ascending) of each cell                                                    */
    SSAM_ADDR_EXTRA_OTHER = 3,  /* [nExtra] owner cell of that face; -1 generic boundary, -2 coupled boundary  */
    SSAM_ADDR_CELL_FACES_START = 4, /* [nCells+1] full cell->face CSR, ascending face index                  */
    SSAM_ADDR_CELL_FACES = 5,   /* [nIF*2+nBnd] face index, bit 31 set when the cell is the neighbour        */
    SSAM_ADDR_HALO_SEND_CELLS = 6, /* [nProcFaces] faceCells of all processor patches, patch order           */
    SSAM_ADDR_HALO_PATCH_OFFSETS = 7, /* [nProcPatches+1] offsets of each processor patch in the send list   */
    SSAM_ADDR_FRIC_CELLS = 8,   /* [patchSize] faceCells of the last friction patch used                     */
    SSAM_ADDR_FACE_COLOUR = 9,  /* [nInternalFaces] greedy face colouring (scatter variant)                  */
    SSAM_ADDR_CELL_ORDER = 10,  /* [nCells] caller's cell label held at each position of the device planes
                                   (identity unless the tile-compact layout is active, see ssam_set_cell_layout) */
    SSAM_ADDR_FACE_ORDER = 11,  /* [nInternalFaces] caller's face index held at each internal-face position of the device
                                   face fields (identity unless the tile-compact layout is active)              */
    SSAM_ADDR_COUNT = 12
};
SSAM_API ssam_status ssam_addressing_size(ssam_ctx* ctx, int which, int64_t* n);
SSAM_API ssam_status ssam_addressing_get(ssam_ctx* ctx, int which, int32_t* out, int64_t n);

/* ---- field transfer ----------------------------------------------------------------------------- */
/* host AoS -> device SoA (transposed on the device); either pointer may be NULL to skip that part */
SSAM_API ssam_status ssam_field_upload(ssam_ctx* ctx, int field, const double* internal, const double* boundary);
SSAM_API ssam_status ssam_field_download(ssam_ctx* ctx, int field, double* internal, double* boundary);
/* device-resident use: pointer to component plane `comp` (length nCells+nBnd: internal then boundary).  The
 * internal part is in the library's cell order: entry i belongs to the caller's cell SSAM_ADDR_CELL_ORDER[i]
 * (identity with SSAM_LAYOUT_CALLER); the boundary part is always in the caller's boundary-face order.
 * ssam_face_field_device_ptr likewise holds the internal faces in the library's face order (SSAM_ADDR_FACE_ORDER). */
SSAM_API ssam_status ssam_field_device_ptr(ssam_ctx* ctx, int field, int comp, void** ptr, int64_t* len);
/* 1 if the field currently holds data (uploaded or computed) */
SSAM_API int ssam_field_is_set(const ssam_ctx* ctx, int field);

/* ---- boundary conditions feeding the gradient ---------------------------------------------------
 * constantRotationalSlipFvPatchVectorField::updateCoeffs (…C:109-139) and
 * exponentialRotationalSlipFvPatchVectorField::updateCoeffs (…C:119-153): writes U_b on `patch`. */
SSAM_API ssam_status ssam_bc_rotational_slip(ssam_ctx* ctx, const ssam_rotslip_params* p);

/* ---- staged entry points (drop-in; each equals one reference call) ------------------------------ */
/* fvc::grad(U) with "Gauss linear" incl. boundary correction -> SSAM_F_GRADU (calculateStrain.H:17) */
SSAM_API ssam_status ssam_grad_u(ssam_ctx* ctx);
/* out_field = factor*mag(symm(fvc::grad(U))):  fluid/TEqn.H:3 (factor sqrt(2/3) -> SSAM_F_EPSDOT),
 * stickModel.C:58-61 / viscosityModel::strainRate() (factor sqrt(2) -> SSAM_F_STRAINRATE) */
SSAM_API ssam_status ssam_strain_rate(ssam_ctx* ctx, double factor, int out_field);
/* viscosityModel::correct() of phase 1 or 2: SheppardWright.H:133-138, FKS.H:156-161,
 * NortonHoff.H:126-129 (recomputes strainRate() from U), Newtonian (constant). Reads T, EPSDOT. */
SSAM_API ssam_status ssam_viscosity_correct(ssam_ctx* ctx, int phase, const ssam_viscosity_params* p);
/* incompressibleTwoPhaseThermalMixture::calcNu tail (…C:46-53): NU = mu()/(la*rho1+(1-la)*rho2) */
SSAM_API ssam_status ssam_mixture_calc_nu(ssam_ctx* ctx, const ssam_mixture_params* p);
/* thermo.correct() = nuModel1.correct(); nuModel2.correct(); calcNu tail
 * (immiscibleIncompressibleTwoPhaseThermalMixture.H:77-81 minus interfaceProperties::correct()) */
SSAM_API ssam_status ssam_thermo_correct(ssam_ctx* ctx, const ssam_viscosity_params* v1,
                                         const ssam_viscosity_params* v2, const ssam_mixture_params* mix);
/* thermo.mu()/rho()/cp()/k() (…C:133-153, 384-400, 365-381, 346-362) -> MUEFF / RHO / CP / K */
SSAM_API ssam_status ssam_mixture_mu(ssam_ctx* ctx, const ssam_mixture_params* p);
SSAM_API ssam_status ssam_mixture_rho(ssam_ctx* ctx, const ssam_mixture_params* p);
SSAM_API ssam_status ssam_mixture_cp(ssam_ctx* ctx, const ssam_mixture_params* p);
SSAM_API ssam_status ssam_mixture_k(ssam_ctx* ctx, const ssam_mixture_params* p);
/* k1() / k2(): the phase conductivity alone -> SSAM_F_K1 / SSAM_F_K2 (phase = 1 or 2) */
SSAM_API ssam_status ssam_mixture_k_phase(ssam_ctx* ctx, const ssam_mixture_params* p, int phase);
/* solver-side rho, rhoCp (multiRegionVoF/alphaEqnSubCycle.H:41-42) -> RHO_SOLVER, RHOCP */
SSAM_API ssam_status ssam_solver_rho_rhocp(ssam_ctx* ctx, const ssam_mixture_params* p);
/* frictionModel::correct() -> stickModel::correct(): qvisc_, qfric_ (constantSlipStick.H:133-138,
 * exponentialSlipStick.H:139-144). Reads MUEFF, EPSDOT, ALPHA1, SIGMAY, P. */
SSAM_API ssam_status ssam_friction_correct(ssam_ctx* ctx, const ssam_stick_params* p);
/* patch centroid of the friction patch after the cross-rank sum (constantSlipStick.C:118-138) */
SSAM_API ssam_status ssam_friction_patch_centre(ssam_ctx* ctx, double centre[3], double* area);

/* fluid/calculateStrain.H:1-35,77-94 (the explicit part; the epsilonEq transport solve :45-58 stays in
 * OpenFOAM): another fvc::grad(U) -> GRADU, then Z, SIGMAF_PP, DEPSEQ, EPSEQ += DEPSEQ, SIGMAVM.
 * Reads U, EPSDOT, T, MUEFF, PRGH, EPSEQ. */
SSAM_API ssam_status ssam_stress_strain(ssam_ctx* ctx, const ssam_stress_strain_params* p);

/* Explicit part of the momentum equation's viscous term (fluid/UEqn.H:13 `turbulence.divDevRhoReff(rho, U)`, for
 * the laminar model OpenFOAM's linearViscousStress::divDevRhoReff):
 *   DIVDEVRHOREFF = -fvc::div((rho*nuEff)*dev2(T(fvc::grad(U))))     with rho = RHO_SOLVER, nuEff = NU
 * (Gauss linear; the implicit -fvm::laplacian(rho*nuEff, U) stays in OpenFOAM).  Reads GRADU (ssam_grad_u or the
 * fused update with flag 1), RHO_SOLVER, NU incl. their boundary values. */
SSAM_API ssam_status ssam_div_dev_rho_reff_explicit(ssam_ctx* ctx);

/* interfaceProperties::correct() = calculateK(), the other half of thermo.correct()
 * (immiscibleIncompressibleTwoPhaseThermalMixture.H:77-81).  The class itself is OpenFOAM v1806 library code
 * (src/transportModels/interfaceProperties/interfaceProperties.C), restated in oracle/oracle.cpp:
 *   gradAlpha = fvc::grad(alpha1)  (Gauss linear)      -> GRADALPHA
 *   nHatfv    = interpolate(gradAlpha)/(mag(interpolate(gradAlpha)) + deltaN);  nHatf = nHatfv & Sf
 *   K         = -fvc::div(nHatf)                       -> CURVATURE
 * Reads ALPHA1 (internal + boundary values; flags bit 1 (value 2): refresh processor-patch values first).
 * No alphaContactAngle patches (the reference's cases have none): correctContactAngle is a no-op.
 * nHatf_ is a surfaceScalarField (alphaEqn.H:154 `phic*thermo.nHatf()`): nInternalFaces + nBoundaryFaces values. */
SSAM_API ssam_status ssam_interface_correct(ssam_ctx* ctx, double deltaN, unsigned flags);

/* ---- surfaceScalarFields: nInternalFaces values, then nBoundaryFaces values in patch order --------- */
enum {
    SSAM_FF_NHATF = 0, /* interfaceProperties::nHatf_                                                   */
    SSAM_FF_MUF = 1,   /* incompressibleTwoPhaseThermalMixture::muf()  (…Mixture.C:163-180)              */
    SSAM_FF_NUF = 2,   /* incompressibleTwoPhaseThermalMixture::nuf()  (…Mixture.C:183-202)              */
    SSAM_FF_KF = 3,    /* fvc::interpolate(kf): the face conductivity fvm::laplacian(kf, temp) uses (fluid/TEqn.H:22)  */
    SSAM_FF_PHI = 4,   /* the face flux phi, IN (ssam_face_field_upload)                                            */
    SSAM_FF_PHIC = 5,  /* the alpha-equation compression coefficient phic (multiRegionVoF/alphaEqn.H:58-87)          */
    SSAM_FF_COUNT = 6
};
SSAM_API ssam_status ssam_face_field_download(ssam_ctx* ctx, int face_field, double* internal_faces,
                                              double* boundary_faces);
/* surfaceScalarField -> device slot (inputs such as the face flux phi): nInternalFaces + nBoundaryFaces values */
SSAM_API ssam_status ssam_face_field_upload(ssam_ctx* ctx, int which, const double* internal_faces,
                                            const double* boundary_faces);
/* The interface-compression coefficient of the alpha equation (multiRegionVoF/alphaEqn.H:58-87) -> SSAM_FF_PHIC:
 *   phic = cAlpha*mag(phi/magSf)
 *   icAlpha > 0:  phic *= (1 - icAlpha);  phic += (cAlpha*icAlpha)*fvc::interpolate(mag(U))            (:61-65)
 *   scAlpha > 0:  phic += scAlpha*mag(mesh.delta() & fvc::interpolate(symm(fvc::grad(U))))             (:68-72)
 *   phic = 0 on non-coupled boundary faces                                                            (:79-87)
 * Needs SSAM_FF_PHI, SSAM_F_U (icAlpha > 0) and SSAM_F_GRADU incl. its boundary / processor values (scAlpha > 0:
 * call ssam_grad_u first).  magSf = mag(Sf), mesh.delta() = C[neighbour] - C[owner] on internal faces and the
 * cell-to-cell vector across processor faces (neighbour cell centres exchanged once per mesh). */
SSAM_API ssam_status ssam_alpha_compression_coeff(ssam_ctx* ctx, double cAlpha, double icAlpha, double scAlpha);
SSAM_API double* ssam_face_field_device_ptr(ssam_ctx* ctx, int face_field);
/* muf()/nuf(): alpha1f = min(max(fvc::interpolate(alpha1),0),1);
 *   muf = alpha1f*rho1*interpolate(nu1) + (1-alpha1f)*rho2*interpolate(nu2);  nuf = muf/(alpha1f*rho1+(1-alpha1f)*rho2)
 * linear interpolation; reads ALPHA1, NU1, NU2 (processor-patch boundary values as left by the last update). */
/* fvc::interpolate(vf) with the linear scheme for any scalar field (internal faces w*(P-N)+N, processor faces
 * w*P+(1-w)*N_b, other boundary faces the boundary value) -> face field slot `face_field`. */
SSAM_API ssam_status ssam_face_interpolate(ssam_ctx* ctx, int field, int face_field);
SSAM_API ssam_status ssam_mixture_muf(ssam_ctx* ctx, const ssam_mixture_params* p);
SSAM_API ssam_status ssam_mixture_nuf(ssam_ctx* ctx, const ssam_mixture_params* p);
/* min(vf).value(), max(vf).value() over internal and boundary values and over all ranks, as printed by
 * fluid/TEqn.H:42-43 ("Min/max T"); scalar fields only. Synchronises the stream. */
SSAM_API ssam_status ssam_field_min_max(ssam_ctx* ctx, int field, double* min_value, double* max_value);
SSAM_API ssam_status ssam_interface_nhatf_download(ssam_ctx* ctx, double* internal_faces, double* boundary_faces);
SSAM_API double* ssam_interface_nhatf_device_ptr(ssam_ctx* ctx);
/* deltaN_ = 1e-8/pow(average(mesh.V()), 1.0/3.0) (interfaceProperties constructor); average over all ranks
 * when a communicator is set. */
SSAM_API ssam_status ssam_interface_delta_n(ssam_ctx* ctx, double* deltaN);

/* ---- fused entry point (the benchmarked path) ---------------------------------------------------
 * One pass: grad(U) -> EPSDOT -> nuModel1/2.correct() (NU1,SIGMAF,SIGMAY,NU2) -> NU -> MUEFF ->
 * QVISC,QFRIC -> RHO,CP,K.  Per-field results equal the staged calls issued in that order on the
 * same inputs.  `flags` (ssam_fused_flags): SSAM_FUSED_WRITE_GRAD also writes SSAM_F_GRADU; SSAM_FUSED_HALO_INPUTS
 * refreshes the processor-patch values of U, temp and alpha1 from the neighbour ranks before the gradient (what
 * U.correctBoundaryConditions() does in the solver, pEqn.H:73) -- a multi-rank caller that does not exchange them
 * itself (ssam_halo_exchange) MUST set it, otherwise the gradient reads stale halo values. */
typedef enum { SSAM_FUSED_WRITE_GRAD = 1, SSAM_FUSED_HALO_INPUTS = 2 } ssam_fused_flags;
SSAM_API ssam_status ssam_update_fused(ssam_ctx* ctx, const ssam_update_params* p, int flags);
/* Performs every allocation and lazy set-up ssam_update_fused(ctx, p, flags) would do (output planes, friction scratch,
 * halo transport) without enqueuing an exchange.  Optional in one-process-per-GPU runs; required on every rank of an
 * in-process rank group (ssam_comm_init_local) before its first collective update. */
SSAM_API ssam_status ssam_prepare_fused(ssam_ctx* ctx, const ssam_update_params* p, int flags);
/* Host-staged drop-in of the same update: uploads U,T,alpha1,p (AoS host buffers, internal+boundary),
 * runs the fused kernels, downloads the listed output fields. Used for the e2e metric. */
typedef struct ssam_host_io {
    const double* U_i; const double* U_b;
    const double* T_i; const double* T_b;
    const double* alpha1_i; const double* alpha1_b;
    const double* p_i; const double* p_b;   /* may be NULL when qfric is skipped */
    int32_t nOut;
    const int32_t* outFields;  /* [nOut] SSAM_F_* */
    double* const* out_i;      /* [nOut] host internal buffers */
    double* const* out_b;      /* [nOut] host boundary buffers (entries may be NULL) */
} ssam_host_io;
SSAM_API ssam_status ssam_update_fused_host(ssam_ctx* ctx, const ssam_update_params* p, const ssam_host_io* io);
/* ssam_update_fused_host pipelines upload / cell kernel / download over contiguous chunks of tiles on three
 * streams (PCIe is full duplex).  n < 0: automatic (default); 0 or 1: unpipelined; otherwise the requested
 * number of chunks, reduced until a chunk is at least as deep as the mesh bandwidth. */
SSAM_API ssam_status ssam_set_host_chunks(ssam_ctx* ctx, int n);

/* ---- test hook ---------------------------------------------------------------------------------------
 * Evaluates one of the device math routines of csrc/fastmath.cuh on host arrays (tests/test_gpu_math.py):
 * fn 0 log, 1 exp, 2 pow(x,y), 3 asinh, 4 exact division x/y, 5 operator / , 6 x/x, 7 asinh(exp(x)). */
SSAM_API ssam_status ssam_debug_math(ssam_ctx* ctx, int fn, const double* x, const double* y, double* out, int64_t n);
/* Measured FP64 issue ceiling of the device: thread-level FP64 instructions per second of a DFMA-only micro-kernel
 * (8 independent chains per thread, 8 CTAs of 256 threads per SM, best of 4 timed launches, CUDA events).  The
 * fused update is co-limited by FP64 issue; bench.py prints roofline.fp64_frac against this. */
SSAM_API ssam_status ssam_measure_fp64_peak(ssam_ctx* ctx, double* fp64_instr_per_s);

/* ---- multi-GPU: processor-patch halos over NCCL --------------------------------------------------
 * Replaces processorFvPatchField initEvaluate/evaluate (OpenFOAM, triggered by
 * correctBoundaryConditions()) and the reduce() pair in constantSlipStick.C:134-135. */
#define SSAM_UNIQUE_ID_BYTES 128
SSAM_API ssam_status ssam_comm_unique_id(void* id_out /* SSAM_UNIQUE_ID_BYTES */);
SSAM_API ssam_status ssam_comm_init(ssam_ctx* ctx, const void* id, int rank, int nRanks);
/* boundary values on processor patches <- neighbour rank's adjacent cell values, for every field
 * whose bit (1<<SSAM_F_*) is set in field_mask */
SSAM_API ssam_status ssam_halo_exchange(ssam_ctx* ctx, uint32_t field_mask);
/* In-process rank group: all `nRanks` contexts live in THIS process (on one GPU or on several with peer access);
 * context r becomes rank r.  Same peer-memory exchange as across processes, the receive slots being linked by plain
 * device pointers; waits are stream memory operations (no kernel ever spins on another rank), so the ranks may share
 * one GPU.  The caller must issue each collective step (ssam_update_fused, ssam_halo_exchange, ...) on EVERY context
 * with ssam_set_blocking(ctx, 0) before synchronising any of them -- a blocking call on one rank would wait for a
 * neighbour whose work has not been enqueued yet.  Used by the single-GPU multi-rank parity tests and by solvers that
 * drive several regions / partitions from one process. */
SSAM_API ssam_status ssam_comm_init_local(ssam_ctx** ctxs, int nRanks);
/* Transport of the processor-patch exchanges: 1 = peer memory over NVLink (the pack kernel stores straight into the
 * neighbour's IPC-mapped receive slot and publishes a flag; the receiver waits with a stream memory operation),
 * 2 = grouped ncclSend/ncclRecv, 0 = not decided yet (before the first exchange) or no processor patches.
 * One process per GPU: NCCL by default; SSAM_HALO=p2p in the environment selects peer memory (falls back to NCCL when
 * a neighbour's memory cannot be mapped).  In-process rank groups (ssam_comm_init_local) always use peer memory.
 * Either way ssam_update_fused runs everything that depends on another rank -- exchange of U, temp, alpha1, the tiles
 * that touch a processor patch, exchange of epsilonDotEq -- on a high-priority side stream beside one launch of all
 * interior tiles. */
SSAM_API int ssam_halo_transport(const ssam_ctx* ctx);
/* Bytes this rank stored into its neighbours' memory (or sent through NCCL) during the last ssam_update_fused. */
SSAM_API int64_t ssam_halo_bytes_last_update(const ssam_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* SSAM_GPU_H */
