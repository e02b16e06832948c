// TEST INFRASTRUCTURE ONLY -- C harness around the reference's own model sources.
//
// oracle/Makefile (target `ref`) compiles this file together with the UNMODIFIED reference translation units
//   viscoPlasticModels/{SheppardWright,FKS,NortonHoff}/*.C
//   frictionModels/{frictionModel,incompressibleFrictionModel,stickModels/*}/*.C
//   chtMultiRegionInterFoam/incompressibleTwoPhaseThermalMixture/incompressibleTwoPhaseThermalMixture.C
// where they lie under /root/reference, against the stand-in OpenFOAM headers in oracle/ofstub/, into
// oracle/_ref/libssam_ref.so.  The harness builds the stand-in fvMesh from the same ssam_mesh_desc the oracle
// and the GPU library take, registers the input fields under the names the reference looks up, constructs the
// reference's mixture and friction model the way createFluidFields.H does (:168-172, :502-506), and exposes
// their member functions.  tests/test_reference_sources.py compares oracle/oracle.cpp against it bit for bit and
// tests/golden/gen_reference_golden.py freezes its outputs into tests/golden/ for machines without /root/reference.
//
// What is pinned this way: the reference's expression trees.  What is not: OpenFOAM's own field algebra and
// fvc:: operators (stand-ins, see ofstub.H), and the OpenFOAM library classes the reference derives from
// (viscosityModel, twoPhaseMixture, the Newtonian model below), which are restated here.
#include <cstring>
#include <memory>

#include "../include/ssam_gpu.h"
#include "ofstub.H"
// reference headers (found through -I, never copied)
#include "constantRotationalSlipFvPatchVectorField.H"
#include "exponentialRotationalSlipFvPatchVectorField.H"
#include "incompressibleFrictionModel.H"
#include "incompressibleTwoPhaseThermalMixture.H"
#include "stickModel.H"

namespace Foam {

const word calculatedFvPatchScalarField::typeName("calculated");

// ---- OpenFOAM library pieces the reference links against -----------------------------------------------------
defineTypeNameAndDebug(viscosityModel, 0);
defineRunTimeSelectionTable(viscosityModel, dictionary);

// viscosityModelNew.C: selects on the keyword "transportModel"
autoPtr<viscosityModel> viscosityModel::New(const word& name, const dictionary& viscosityProperties,
                                            const volVectorField& U, const surfaceScalarField& phi) {
    const word modelType(viscosityProperties.lookup("transportModel"));
    Info << "Selecting incompressible transport model " << modelType << endl;
    auto cstrIter = dictionaryConstructorTablePtr_->cfind(modelType);
    if (!cstrIter.found()) {
        FatalErrorInFunction << "Unknown viscosityModel type " << modelType << nl << nl << "Valid viscosityModels are : " << endl
                             << dictionaryConstructorTablePtr_->sortedToc() << exit(FatalError);
    }
    return autoPtr<viscosityModel>(cstrIter()(name, viscosityProperties, U, phi));
}

namespace viscosityModels {
// OpenFOAM's Newtonian model (phase 2 of the reference's cases): nu_ = nu0 everywhere, correct() does nothing
class Newtonian : public viscosityModel {
    dimensionedScalar nu0_;
    volScalarField nu_;

public:
    TypeName("Newtonian");
    Newtonian(const word& name, const dictionary& viscosityProperties, const volVectorField& U,
              const surfaceScalarField& phi)
        : viscosityModel(name, viscosityProperties, U, phi),
          nu0_("nu", dimViscosity, viscosityProperties_),
          nu_(IOobject(name, U_.time().timeName(), U_.db(), IOobject::NO_READ, IOobject::NO_WRITE), U_.mesh(), nu0_) {}
    virtual tmp<volScalarField> nu() const { return nu_; }
    virtual tmp<scalarField> nu(const label patchi) const { return nu_.boundaryField()[patchi]; }
    virtual void correct() {}
};
defineTypeNameAndDebug(Newtonian, 0);
addToRunTimeSelectionTable(viscosityModel, Newtonian, dictionary);
}  // namespace viscosityModels
}  // namespace Foam

using namespace Foam;

namespace {
struct RefCase {
    fvMesh mesh;
    std::map<std::string, std::unique_ptr<volScalarField>> inputs;  // registered input fields
    std::unique_ptr<volVectorField> U, Ut;
    std::unique_ptr<surfaceScalarField> phi;
    std::unique_ptr<incompressibleTwoPhaseThermalMixture> thermo;
    std::unique_ptr<incompressibleFrictionModel> friction;
    std::string err, info;
    int nTot = 0;
};

template <class Fn>
int guarded(RefCase* c, Fn fn) {
    try {
        fn();
        c->info = infoStream().str();
        return 0;
    } catch (const std::exception& e) {
        c->err = e.what();
        c->info = infoStream().str();
        return 1;
    }
}
}  // namespace

extern "C" {

typedef struct ref_case ref_case;
#define RC(h) (reinterpret_cast<RefCase*>(h))

ref_case* ref_create(const ssam_mesh_desc* m, const char* const* patchNames, const char* caseDir) {
    RefCase* c = new RefCase;
    fvMesh& mesh = c->mesh;
    mesh.runTime.caseDir = caseDir;
    mesh.nCells = m->nCells;
    mesh.nInternalFaces = m->nInternalFaces;
    mesh.nBoundaryFaces = m->nFaces - m->nInternalFaces;
    c->nTot = mesh.nCells + mesh.nBoundaryFaces;
    mesh.owner.assign(m->owner, m->owner + m->nFaces);
    mesh.neighbour.assign(m->neighbour, m->neighbour + m->nInternalFaces);
    mesh.weights.assign(m->weights, m->weights + m->nFaces);
    for (int p = 0; p < m->nPatches; ++p) {
        polyPatch pp;
        pp.name = patchNames[p];
        pp.start = m->patchStart[p] - m->nInternalFaces;
        pp.n = m->patchSize[p];
        pp.coupled = m->patchKind[p] == SSAM_PATCH_PROCESSOR;
        pp.cells.assign(m->owner + m->patchStart[p], m->owner + m->patchStart[p] + m->patchSize[p]);
        mesh.bMesh.patches.push_back(pp);
    }
    // mesh.C(): cell centres, boundary values = face centres; mesh.Cf()/magSf(): only boundary values are used
    mesh.Cptr.reset(new volVectorField(mesh, "C"));
    for (int i = 0; i < mesh.nCells; ++i) mesh.Cptr->values()[size_t(i)] = vector(m->C[3 * i], m->C[3 * i + 1], m->C[3 * i + 2]);
    mesh.CfPtr.reset(new surfaceVectorField(mesh, "Cf"));
    mesh.magSfPtr.reset(new surfaceScalarField(mesh, "magSf"));
    for (int b = 0; b < mesh.nBoundaryFaces; ++b) {
        const vector cf(m->Cf_b[3 * b], m->Cf_b[3 * b + 1], m->Cf_b[3 * b + 2]);
        mesh.Cptr->values()[size_t(mesh.nCells + b)] = cf;
        mesh.CfPtr->values()[size_t(mesh.nInternalFaces + b)] = cf;
        mesh.magSfPtr->values()[size_t(mesh.nInternalFaces + b)] = m->magSf_b[b];
    }
    {
        fvMesh::GeometryStash& g = mesh.stash;
        const int nF = m->nFaces, nIF = m->nInternalFaces, nC = m->nCells, nB = nF - nIF, nP = m->nPatches;
        g.owner.assign(m->owner, m->owner + nF);
        g.neighbour.assign(m->neighbour, m->neighbour + nIF);
        g.patchStart.assign(m->patchStart, m->patchStart + nP);
        g.patchSize.assign(m->patchSize, m->patchSize + nP);
        g.patchKind.assign(m->patchKind, m->patchKind + nP);
        g.patchNbrRank.assign(m->patchNbrRank, m->patchNbrRank + nP);
        g.patchUbc.assign(size_t(nP), SSAM_BC_VALUE);
        g.Sf.assign(m->Sf, m->Sf + 3 * size_t(nF));
        g.weights.assign(m->weights, m->weights + nF);
        g.V.assign(m->V, m->V + nC);
        g.C.assign(m->C, m->C + 3 * size_t(nC));
        g.Cf_b.assign(m->Cf_b, m->Cf_b + 3 * size_t(nB));
        g.magSf_b.assign(m->magSf_b, m->magSf_b + nB);
        g.deltaCoeffs_b.assign(m->deltaCoeffs_b, m->deltaCoeffs_b + nB);
    }
    c->U.reset(new volVectorField(mesh, "U"));
    c->Ut.reset(new volVectorField(mesh, "Ut"));
    c->phi.reset(new surfaceScalarField(mesh, "phi"));
    return reinterpret_cast<ref_case*>(c);
}

// U (internal then boundary values, AoS) and the snGrad kind of each patch of U: only the GPU-backed models of
// integration/openfoam need them (the CPU models get mag(symm(grad U)) from the oracle through "__magSymmGradU")
int ref_set_U(ref_case* h, const double* values, const int32_t* patchKinds) {
    RefCase* c = RC(h);
    return guarded(c, [&]() {
        for (int i = 0; i < c->nTot; ++i) c->U->values()[size_t(i)] = vector(values[3 * i], values[3 * i + 1], values[3 * i + 2]);
        if (patchKinds) c->mesh.stash.patchUbc.assign(patchKinds, patchKinds + c->mesh.bMesh.size());
    });
}

#ifdef SSAM_WITH_GPU_SHIM
}  // extern "C"
#include "ssamGpuMesh.H"
extern "C" {
#endif

void ref_destroy(ref_case* h) {
    RefCase* c = RC(h);
    c->friction.reset();
    c->thermo.reset();
#ifdef SSAM_WITH_GPU_SHIM
    ssamGpuMesh::release(c->mesh);
#endif
    delete c;
}
const char* ref_last_error(ref_case* h) { return RC(h)->err.c_str(); }
const char* ref_info(ref_case* h) { return RC(h)->info.c_str(); }

// registers (first call) or overwrites a volScalarField the reference looks up by name: "temp", "epsilonDotEq",
// "muEff", pName, "__magSymmGradU"; kind 1 = a field file that a MUST_READ constructor finds (alpha.<phase1>, qfric)
int ref_set_scalar(ref_case* h, const char* name, const double* values, int asFile) {
    RefCase* c = RC(h);
    return guarded(c, [&]() {
        if (asFile) {
            c->mesh.files[name].assign(values, values + c->nTot);
            if (c->thermo && c->thermo->alpha1().name() == name) c->thermo->alpha1().values().assign(values, values + c->nTot);
            return;
        }
        auto it = c->inputs.find(name);
        if (it == c->inputs.end()) {
            volScalarField init(c->mesh, name);
            c->inputs[name].reset(new volScalarField(IOobject(name, "0", c->mesh, IOobject::NO_READ, IOobject::AUTO_WRITE), init));
            it = c->inputs.find(name);
        }
        it->second->values().assign(values, values + c->nTot);
    });
}

// createFluidFields.H:168-172 (mixture) and :502-506 (friction model); `withFriction` 0 skips the latter
int ref_construct(ref_case* h, int withFriction) {
    RefCase* c = RC(h);
    return guarded(c, [&]() {
        infoStream().str("");
        c->thermo.reset(new incompressibleTwoPhaseThermalMixture(*c->U, *c->phi));
        if (withFriction) c->friction.reset(new incompressibleFrictionModel(*c->U, *c->Ut));
    });
}

int ref_thermo_correct(ref_case* h) {
    RefCase* c = RC(h);
    return guarded(c, [&]() { c->thermo->correct(); });
}
int ref_friction_correct(ref_case* h) {
    RefCase* c = RC(h);
    return guarded(c, [&]() { c->friction->correct(); });
}

// evaluates a member function / registered field into out: vol fields nCells+nBnd values, face fields nFaces
int ref_eval(ref_case* h, const char* what, double* out) {
    RefCase* c = RC(h);
    return guarded(c, [&]() {
        const std::string w(what);
        auto copyV = [&](const volScalarField& f) { std::memcpy(out, f.values().data(), f.values().size() * sizeof(double)); };
        auto copyS = [&](const surfaceScalarField& f) { std::memcpy(out, f.values().data(), f.values().size() * sizeof(double)); };
        if (w == "mu") copyV(c->thermo->mu()());
        else if (w == "nu") copyV(c->thermo->nu()());
        else if (w == "k") copyV(c->thermo->k()());
        else if (w == "k1") copyV(c->thermo->k1()());
        else if (w == "k2") copyV(c->thermo->k2()());
        else if (w == "cp") copyV(c->thermo->cp()());
        else if (w == "rho") copyV(c->thermo->rho()());
        else if (w == "nu1") copyV(c->thermo->nuModel1().nu()());
        else if (w == "nu2") copyV(c->thermo->nuModel2().nu()());
        else if (w == "muf") copyS(c->thermo->muf()());
        else if (w == "nuf") copyS(c->thermo->nuf()());
        else if (w == "qvisc") copyV(c->friction->qvisc()());
        else if (w == "qfric") copyV(c->friction->qfric()());
        else copyV(c->mesh.lookupObject<volScalarField>(w));  // "sigmaf", "sigmay", inputs
    });
}

// boundaryConditions/*RotationalSlip*: constructs the patch field from the patch's entry in <caseDir>/0/U
// (dictionary constructor, as fvPatchField::New does), calls updateCoeffs() and copies U_b out ([3*n], AoS)
int ref_bc_rotational_slip(ref_case* h, const char* patchName, double* out) {
    RefCase* c = RC(h);
    return guarded(c, [&]() {
        dictionary Udict;
        Udict.read(c->mesh.runTime.caseDir + "/0/U");
        const dictionary& pd = Udict.subDict("boundaryField").subDict(patchName);
        const label pi = c->mesh.boundaryMesh().findPatchID(patchName);
        if (pi < 0) throw FoamFatal(std::string("no patch ") + patchName);
        const polyPatch& pp = c->mesh.boundaryMesh()[pi];
        fvPatch patch;
        patch.mesh_ = &c->mesh;
        patch.index_ = pi;
        for (label i = 0; i < pp.n; ++i) patch.Cf_.push_back(c->mesh.Cf().values()[size_t(c->mesh.nInternalFaces + pp.start + i)]);
        DimensionedField<vector, volMesh> iF;
        const word type(pd.lookup("type"));
        std::unique_ptr<fvPatchVectorField> f;
        if (type == constantRotationalSlipFvPatchVectorField::typeName)
            f.reset(new constantRotationalSlipFvPatchVectorField(patch, iF, pd));
        else if (type == exponentialRotationalSlipFvPatchVectorField::typeName)
            f.reset(new exponentialRotationalSlipFvPatchVectorField(patch, iF, pd));
        else
            throw FoamFatal("Unknown patchField type " + type);
        f->updateCoeffs();
        for (label i = 0; i < pp.n; ++i)
            for (int k = 0; k < 3; ++k) out[3 * i + k] = (*f)[size_t(i)][k];
    });
}

// fluid/calculateStrain.H, included verbatim with the solver's variable names in scope.  Its fvm:: transport solve
// of epsilonEq is a no-op here (ofstub.H), as it is outside ssam_stress_strain.  gradU: [9*nTot] from the oracle.
int ref_calculate_strain(ref_case* h, double deltaT, const double* gradUIn, const double* pRgh, double* epsilonEqInOut,
                         double* ZOut, double* sigmafOut, double* sigmavmOut) {
    RefCase* c = RC(h);
    return guarded(c, [&]() {
        fvMesh& mesh = c->mesh;
        const size_t n = size_t(c->nTot);
        mesh.gradUValues.resize(n);
        for (size_t i = 0; i < n; ++i)
            for (int k = 0; k < 9; ++k) mesh.gradUValues[i].c[k] = gradUIn[9 * i + k];
        mesh.runTime.deltaTValue = deltaT;
        Time& runTime = mesh.runTime;
        incompressibleTwoPhaseThermalMixture& thermo = *c->thermo;
        const volVectorField& U = *c->U;
        const volScalarField& epsilonDotEq = mesh.lookupObject<volScalarField>("epsilonDotEq");
        const volScalarField& temp = mesh.lookupObject<volScalarField>("temp");
        const volScalarField& muEff = mesh.lookupObject<volScalarField>("muEff");
        volScalarField Z(mesh, "Z"), sigmaf(mesh, "sigmafPP"), sigmavm(mesh, "sigmavm"), epsilonEq(mesh, "epsilonEq"),
            p_rgh(mesh, "p_rgh"), rho(mesh, "rho");
        surfaceScalarField rhoPhi(mesh, "rhoPhi");
        dimensionedScalar gammaEpsilon("gammaEpsilon", dimless, 0.0);
        const bool finalIter = true;
        p_rgh.values().assign(pRgh, pRgh + n);
        epsilonEq.values().assign(epsilonEqInOut, epsilonEqInOut + n);
        (void)runTime; (void)U; (void)finalIter;
#include "calculateStrain.H"
        std::memcpy(epsilonEqInOut, epsilonEq.values().data(), n * sizeof(double));
        std::memcpy(ZOut, Z.values().data(), n * sizeof(double));
        std::memcpy(sigmafOut, sigmaf.values().data(), n * sizeof(double));
        std::memcpy(sigmavmOut, sigmavm.values().data(), n * sizeof(double));
    });
}

#ifdef SSAM_GPU_MIXTURE
}  // extern "C"
#include "ssamTEqnUpdate.H"
extern "C" {

// fluid/TEqn.H:3-15 through integration/openfoam/ssamTEqnUpdate.H on the case's own registered fields
// ("epsilonDotEq", "muEff"; the results cpf / kf are registered as "cpf" / "kf").  fused = 0: the staged device-resident
// sequence; 1: one ssam_update_fused with the TEqn outputs only; 2: ... with every output.
int ref_teqn_update(ref_case* h, int fused) {
    RefCase* c = RC(h);
    return guarded(c, [&]() {
        for (const char* nm : {"cpf", "kf"})
            if (c->inputs.find(nm) == c->inputs.end()) {
                volScalarField init(c->mesh, nm);
                c->inputs[nm].reset(new volScalarField(IOobject(nm, "0", c->mesh, IOobject::NO_READ, IOobject::AUTO_WRITE), init));
            }
        volScalarField& eps = *c->inputs.at("epsilonDotEq");
        volScalarField& muEff = *c->inputs.at("muEff");
        if (fused)
            ssamFusedUpdate(*c->U, *c->thermo, eps, muEff, *c->inputs.at("cpf"), *c->inputs.at("kf"), fused == 2);
        else
            ssamTEqnUpdate(*c->U, *c->thermo, c->friction.get(), eps, muEff, *c->inputs.at("cpf"), *c->inputs.at("kf"));
    });
}
#endif

#ifdef SSAM_WITH_GPU_SHIM
// bytes the shim moved over PCIe for this case's mesh since the last reset: out[0] = host->device, out[1] = device->host
void ref_gpu_traffic(ref_case* h, double* out, int reset) {
    const ssamGpuMesh& gpu = ssamGpuMesh::New(RC(h)->mesh);
    out[0] = double(gpu.bytesH2D());
    out[1] = double(gpu.bytesD2H());
    if (reset) gpu.resetTraffic();
}
// number of live GPU contexts (one per mesh / region)
int ref_gpu_contexts() { return int(ssamGpuMesh::table().size()); }
#endif

}  // extern "C"
