// GPU-backed translation unit of the reference's incompressibleTwoPhaseThermalMixture: build THIS file INSTEAD of
// chtMultiRegionInterFoam/incompressibleTwoPhaseThermalMixture/incompressibleTwoPhaseThermalMixture.C, against the
// reference's own, unmodified incompressibleTwoPhaseThermalMixture.H.  Class layout, constructor, TypeName and every
// public member stay the reference's; only the bodies change: each property is one C-ABI call on fields that stay
// on the device (ssamGpuMesh.H), the result is downloaded into the tmp<> the solver consumes.
//
//   calcNu()                  …Mixture.C:41-54    nuModel1_->correct(); nuModel2_->correct(); ssam_mixture_calc_nu
//   mu(), mu(patchI)          :133-161            ssam_mixture_mu
//   muf(), nuf()              :163-202            ssam_mixture_muf / ssam_mixture_nuf
//   k1(), k2(), k()           :205-362            ssam_mixture_k_phase / ssam_mixture_k (cubic k(T), units, clamp)
//   cp(), rho()               :365-400            ssam_mixture_cp / ssam_mixture_rho
//   read()                    :402-440            unchanged logic (dictionary only)
// Parameters are looked up in the phase dictionaries on every call, like the reference does.
#include "incompressibleTwoPhaseThermalMixture.H"
#include "addToRunTimeSelectionTable.H"
#include "surfaceFields.H"
#include "fvc.H"

#include <cstring>

#include "ssamGpuMesh.H"
#include "ssamMixtureParams.H"

// * * * * * * * * * * * * * * Static Data Members * * * * * * * * * * * * * //

namespace Foam
{
    defineTypeNameAndDebug(incompressibleTwoPhaseThermalMixture, 0);
}

namespace
{
using namespace Foam;

// alpha1, and the two phase viscosities as their models hold them, on the device (uploaded only if they changed)
void stageViscosities(const ssamGpuMesh& gpu, const volScalarField& alpha1, const viscosityModel& m1, const viscosityModel& m2)
{
    gpu.upload(SSAM_F_ALPHA1, alpha1);
    ssamStageViscosity(gpu, 1, m1);
    ssamStageViscosity(gpu, 2, m2);
}

tmp<volScalarField> fetch(const ssamGpuMesh& gpu, int field, const word& name, const volScalarField& like)
{
    tmp<volScalarField> t(new volScalarField(name, like));
    gpu.deviceWrote(field);
    gpu.download(field, t.ref());
    return t;
}
}  // namespace


// * * * * * * * * * * * Protected Member Functions  * * * * * * * * * * * * //

void Foam::incompressibleTwoPhaseThermalMixture::calcNu()
{
    nuModel1_->correct();
    nuModel2_->correct();

    const ssamGpuMesh& gpu = ssamGpuMesh::New(U_.mesh());
    ssam_mixture_params p;
    ssamMixtureParams(*this, p, false);
    stageViscosities(gpu, alpha1_, nuModel1_(), nuModel2_());
    gpu.check(ssam_mixture_calc_nu(gpu.ctx(), &p), "incompressibleTwoPhaseThermalMixture::calcNu");
    gpu.deviceWrote(SSAM_F_NU);
    gpu.download(SSAM_F_NU, nu_);
}


// * * * * * * * * * * * * * * * * Constructors  * * * * * * * * * * * * * * //

Foam::incompressibleTwoPhaseThermalMixture::incompressibleTwoPhaseThermalMixture
(
    const volVectorField& U,
    const surfaceScalarField& phi
)
:
    IOdictionary
    (
        IOobject
        (
            "transportProperties",
            U.time().constant(),
            U.db(),
            IOobject::MUST_READ_IF_MODIFIED,
            IOobject::NO_WRITE
        )
    ),
    twoPhaseMixture(U.mesh(), *this),
    nuModel1_(viscosityModel::New("nu1", subDict(phase1Name_), U, phi)),
    nuModel2_(viscosityModel::New("nu2", subDict(phase2Name_), U, phi)),
    rho1_("rho", dimDensity, nuModel1_->viscosityProperties()),
    rho2_("rho", dimDensity, nuModel2_->viscosityProperties()),
    cp1_("cp", dimEnergy/dimMass/dimTemperature, nuModel1_->viscosityProperties()),
    cp2_("cp", dimEnergy/dimMass/dimTemperature, nuModel2_->viscosityProperties()),
    kModel1_(nuModel1_->viscosityProperties().lookupOrDefault<word>("kModel", "constant")),
    kModel2_(nuModel2_->viscosityProperties().lookupOrDefault<word>("kModel", "constant")),
    U_(U),
    phi_(phi),
    nu_
    (
        IOobject("nu", U_.time().timeName(), U_.db()),
        U_.mesh(),
        dimensionedScalar(dimViscosity, Zero),
        calculatedFvPatchScalarField::typeName
    )
{
    calcNu();
}


// * * * * * * * * * * * * * * Member Functions  * * * * * * * * * * * * * * //

Foam::tmp<Foam::volScalarField>
Foam::incompressibleTwoPhaseThermalMixture::mu() const
{
    const ssamGpuMesh& gpu = ssamGpuMesh::New(U_.mesh());
    ssam_mixture_params p;
    ssamMixtureParams(*this, p, false);
    stageViscosities(gpu, alpha1_, nuModel1_(), nuModel2_());
    gpu.check(ssam_mixture_mu(gpu.ctx(), &p), "incompressibleTwoPhaseThermalMixture::mu");
    return fetch(gpu, SSAM_F_MUEFF, "mu", nu_);
}


Foam::tmp<Foam::scalarField>
Foam::incompressibleTwoPhaseThermalMixture::mu(const label patchI) const
{
    return mu()().boundaryField()[patchI];
}


Foam::tmp<Foam::surfaceScalarField>
Foam::incompressibleTwoPhaseThermalMixture::muf() const
{
    const ssamGpuMesh& gpu = ssamGpuMesh::New(U_.mesh());
    ssam_mixture_params p;
    ssamMixtureParams(*this, p, false);
    stageViscosities(gpu, alpha1_, nuModel1_(), nuModel2_());
    gpu.check(ssam_mixture_muf(gpu.ctx(), &p), "incompressibleTwoPhaseThermalMixture::muf");
    tmp<surfaceScalarField> t(new surfaceScalarField("muf", phi_));
    gpu.downloadFaceField(SSAM_FF_MUF, t.ref());
    return t;
}


Foam::tmp<Foam::surfaceScalarField>
Foam::incompressibleTwoPhaseThermalMixture::nuf() const
{
    const ssamGpuMesh& gpu = ssamGpuMesh::New(U_.mesh());
    ssam_mixture_params p;
    ssamMixtureParams(*this, p, false);
    stageViscosities(gpu, alpha1_, nuModel1_(), nuModel2_());
    gpu.check(ssam_mixture_nuf(gpu.ctx(), &p), "incompressibleTwoPhaseThermalMixture::nuf");
    tmp<surfaceScalarField> t(new surfaceScalarField("nuf", phi_));
    gpu.downloadFaceField(SSAM_FF_NUF, t.ref());
    return t;
}


Foam::tmp<Foam::volScalarField>
Foam::incompressibleTwoPhaseThermalMixture::k1() const
{
    const ssamGpuMesh& gpu = ssamGpuMesh::New(U_.mesh());
    ssam_mixture_params p;
    ssamMixtureParams(*this, p, true, 1);
    gpu.upload(SSAM_F_T, U_.mesh().lookupObject<volScalarField>("temp"));
    gpu.check(ssam_mixture_k_phase(gpu.ctx(), &p, 1), "incompressibleTwoPhaseThermalMixture::k1");
    return fetch(gpu, SSAM_F_K1, "k1", nu_);
}


Foam::tmp<Foam::volScalarField>
Foam::incompressibleTwoPhaseThermalMixture::k2() const
{
    const ssamGpuMesh& gpu = ssamGpuMesh::New(U_.mesh());
    ssam_mixture_params p;
    ssamMixtureParams(*this, p, true, 2);
    gpu.upload(SSAM_F_T, U_.mesh().lookupObject<volScalarField>("temp"));
    gpu.check(ssam_mixture_k_phase(gpu.ctx(), &p, 2), "incompressibleTwoPhaseThermalMixture::k2");
    return fetch(gpu, SSAM_F_K2, "k2", nu_);
}


Foam::tmp<Foam::volScalarField>
Foam::incompressibleTwoPhaseThermalMixture::k() const
{
    const ssamGpuMesh& gpu = ssamGpuMesh::New(U_.mesh());
    ssam_mixture_params p;
    ssamMixtureParams(*this, p, true);
    gpu.upload(SSAM_F_ALPHA1, alpha1_);
    gpu.upload(SSAM_F_T, U_.mesh().lookupObject<volScalarField>("temp"));
    gpu.check(ssam_mixture_k(gpu.ctx(), &p), "incompressibleTwoPhaseThermalMixture::k");
    return fetch(gpu, SSAM_F_K, "k", nu_);
}


Foam::tmp<Foam::volScalarField>
Foam::incompressibleTwoPhaseThermalMixture::cp() const
{
    const ssamGpuMesh& gpu = ssamGpuMesh::New(U_.mesh());
    ssam_mixture_params p;
    ssamMixtureParams(*this, p, false);
    gpu.upload(SSAM_F_ALPHA1, alpha1_);
    gpu.check(ssam_mixture_cp(gpu.ctx(), &p), "incompressibleTwoPhaseThermalMixture::cp");
    return fetch(gpu, SSAM_F_CP, "cp", nu_);
}


Foam::tmp<Foam::volScalarField>
Foam::incompressibleTwoPhaseThermalMixture::rho() const
{
    const ssamGpuMesh& gpu = ssamGpuMesh::New(U_.mesh());
    ssam_mixture_params p;
    ssamMixtureParams(*this, p, false);
    gpu.upload(SSAM_F_ALPHA1, alpha1_);
    gpu.check(ssam_mixture_rho(gpu.ctx(), &p), "incompressibleTwoPhaseThermalMixture::rho");
    return fetch(gpu, SSAM_F_RHO, "rho", nu_);
}


bool Foam::incompressibleTwoPhaseThermalMixture::read()
{
    if (regIOobject::read())
    {
        if
        (
            nuModel1_().read(subDict(phase1Name_ == "1" ? "phase1" : phase1Name_))
         && nuModel2_().read(subDict(phase2Name_ == "2" ? "phase2" : phase2Name_))
        )
        {
            nuModel1_->viscosityProperties().lookup("rho") >> rho1_;
            nuModel2_->viscosityProperties().lookup("rho") >> rho2_;
            nuModel1_->viscosityProperties().lookup("kModel") >> kModel1_;
            nuModel2_->viscosityProperties().lookup("kModel") >> kModel2_;
            nuModel1_->viscosityProperties().lookup("cp") >> cp1_;
            nuModel2_->viscosityProperties().lookup("cp") >> cp2_;
            return true;
        }
        return false;
    }
    return false;
}

// ************************************************************************* //
