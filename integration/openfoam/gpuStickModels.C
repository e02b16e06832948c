#include "gpuStickModels.H"
#include "addToRunTimeSelectionTable.H"
#include "surfaceFields.H"

#include <cstring>

namespace Foam
{
namespace stickModels
{
    defineTypeNameAndDebug(constantSlipStick, 0);
    addToRunTimeSelectionTable(stickModel, constantSlipStick, dictionary);
    defineTypeNameAndDebug(exponentialSlipStick, 0);
    addToRunTimeSelectionTable(stickModel, exponentialSlipStick, dictionary);
}
}

// The reference constructs qvisc_ from calcQvisc() and reads qfric_ from the time directory (MUST_READ); the first
// correct() overwrites both.  Same here: qvisc from the device, qfric from its file.
Foam::stickModels::gpuStickBase::gpuStickBase
(
    const word& name,
    const dictionary& stickProperties,
    const volVectorField& U,
    const volVectorField& Ut,
    const word& coeffsName
)
:
    stickModel(name, stickProperties, U, Ut),
    coeffs_(stickProperties.optionalSubDict(coeffsName)),
    alphaName_(stickProperties.lookup("alphaName")),
    pName_(stickProperties.lookup("pName")),
    patch_(stickProperties.lookup("fricPatch")),
    qvisc_
    (
        IOobject("qvisc", U_.time().timeName(), U_.db(), IOobject::NO_READ, IOobject::AUTO_WRITE),
        U_.mesh(),
        dimensionedScalar("qvisc", dimPower/dimVolume, 0.0)
    ),
    qfric_
    (
        IOobject("qfric", U_.time().timeName(), U_.mesh(), IOobject::MUST_READ, IOobject::AUTO_WRITE),
        U_.mesh()
    )
{
    ssamGpuMesh::New(U_.mesh()).setStickModel(this);
}


void Foam::stickModels::gpuStickBase::params(ssam_stick_params& p) const
{
    std::memset(&p, 0, sizeof(p));
    fill(p);
    p.fricPatch = U_.mesh().boundaryMesh().findPatchID(patch_);
}


void Foam::stickModels::gpuStickBase::gpuCorrect(bool withQfric)
{
    const fvMesh& mesh = U_.mesh();
    ssamGpuMesh& gpu = ssamGpuMesh::New(mesh);
    ssam_stick_params p;
    std::memset(&p, 0, sizeof(p));
    fill(p);
    p.fricPatch = withQfric ? mesh.boundaryMesh().findPatchID(patch_) : -1;
    gpu.upload(SSAM_F_MUEFF, mesh.lookupObject<volScalarField>("muEff"));
    gpu.upload(SSAM_F_EPSDOT, mesh.lookupObject<volScalarField>("epsilonDotEq"));
    if (withQfric)
    {
        gpu.upload(SSAM_F_ALPHA1, mesh.lookupObject<volScalarField>(alphaName_));
        gpu.upload(SSAM_F_SIGMAY, mesh.lookupObject<volScalarField>("sigmay"));
        gpu.upload(SSAM_F_P, mesh.lookupObject<volScalarField>(pName_));
    }
    gpu.check(ssam_friction_correct(gpu.ctx(), &p), "stickModel::correct");
    gpu.download(SSAM_F_QVISC, qvisc_);
    if (withQfric)
    {
        gpu.download(SSAM_F_QFRIC, qfric_);
        double c[3], area;
        gpu.check(ssam_friction_patch_centre(gpu.ctx(), c, &area), "patch centre");
        Info<< "Center for patch " << patch_ << " found to be at " << vector(c[0], c[1], c[2]) << endl;
    }
}


bool Foam::stickModels::gpuStickBase::read(const dictionary& stickProperties)
{
    stickModel::read(stickProperties);
    coeffs_ = stickProperties.optionalSubDict(type() + "Coeffs");
    coeffs_.lookup("alphaName") >> alphaName_;
    coeffs_.lookup("pName") >> pName_;
    coeffs_.lookup("fricPatch") >> patch_;
    return true;
}


Foam::stickModels::constantSlipStick::constantSlipStick
(
    const word& name,
    const dictionary& stickProperties,
    const volVectorField& U,
    const volVectorField& Ut
)
:
    gpuStickBase(name, stickProperties, U, Ut, typeName + "Coeffs")
{
    gpuCorrect(false);  // qvisc_ as the reference's constructor leaves it
}

void Foam::stickModels::constantSlipStick::fill(ssam_stick_params& p) const
{
    p.model = SSAM_STICK_CONSTANT;
    p.betaTQ = dimensionedScalar("betaTQ", dimless, coeffs_).value();
    p.delta = dimensionedScalar("delta", dimless, coeffs_).value();
    p.muf = dimensionedScalar("muf", dimless, coeffs_).value();
    const vector w = dimensionedVector("omega", dimless/dimTime, coeffs_).value();
    p.omega[0] = w.x(); p.omega[1] = w.y(); p.omega[2] = w.z();
}


Foam::stickModels::exponentialSlipStick::exponentialSlipStick
(
    const word& name,
    const dictionary& stickProperties,
    const volVectorField& U,
    const volVectorField& Ut
)
:
    gpuStickBase(name, stickProperties, U, Ut, typeName + "Coeffs")
{
    gpuCorrect(false);
}

void Foam::stickModels::exponentialSlipStick::fill(ssam_stick_params& p) const
{
    p.model = SSAM_STICK_EXPONENTIAL;
    p.betaTQ = dimensionedScalar("betaTQ", dimless, coeffs_).value();
    p.omega0 = dimensionedScalar("omega0", dimless/dimTime, coeffs_).value();
    p.delta0 = dimensionedScalar("delta0", dimless, coeffs_).value();
    p.Rs = dimensionedScalar("Rs", dimLength, coeffs_).value();
    p.mu0 = dimensionedScalar("mu0", dimless, coeffs_).value();
    p.lambda = dimensionedScalar("lambda", dimTime/dimLength, coeffs_).value();
    const vector w = dimensionedVector("omega", dimless/dimTime, coeffs_).value();
    p.omega[0] = w.x(); p.omega[1] = w.y(); p.omega[2] = w.z();
}
