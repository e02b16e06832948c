#include "gpuViscosityModels.H"
#include "addToRunTimeSelectionTable.H"
#include "surfaceFields.H"

#include <cstring>

namespace Foam
{
namespace viscosityModels
{
    defineTypeNameAndDebug(SheppardWright, 0);
    addToRunTimeSelectionTable(viscosityModel, SheppardWright, dictionary);
    defineTypeNameAndDebug(FKS, 0);
    addToRunTimeSelectionTable(viscosityModel, FKS, dictionary);
    defineTypeNameAndDebug(NortonHoff, 0);
    addToRunTimeSelectionTable(viscosityModel, NortonHoff, dictionary);
}
}

Foam::viscosityModels::gpuViscoPlasticBase::gpuViscoPlasticBase
(
    const word& name,
    const dictionary& viscosityProperties,
    const volVectorField& U,
    const surfaceScalarField& phi,
    const word& coeffsName,
    bool hasSigma
)
:
    viscosityModel(name, viscosityProperties, U, phi),
    coeffs_(viscosityProperties.optionalSubDict(coeffsName)),
    nu_
    (
        IOobject(name, U_.time().timeName(), U_.db(), IOobject::NO_READ, IOobject::AUTO_WRITE),
        U_.mesh(),
        dimensionedScalar("nu", dimViscosity, 0.0)
    )
{
    ssamGpuMesh::New(U_.mesh()).setViscModel(phase(), this);
    if (hasSigma)
    {
        sigmaf_ = autoPtr<volScalarField>(new volScalarField
        (
            IOobject("sigmaf", U_.time().timeName(), U_.db(), IOobject::NO_READ, IOobject::AUTO_WRITE),
            U_.mesh(),
            dimensionedScalar("sigmaf", dimPressure, 0.0)
        ));
        sigmay_ = autoPtr<volScalarField>(new volScalarField
        (
            IOobject("sigmay", U_.time().timeName(), U_.db(), IOobject::NO_READ, IOobject::AUTO_WRITE),
            U_.mesh(),
            dimensionedScalar("sigmay", dimPressure, 0.0)
        ));
    }
}


void Foam::viscosityModels::gpuViscoPlasticBase::gpuCorrect()
{
    const fvMesh& mesh = U_.mesh();
    ssamGpuMesh& gpu = ssamGpuMesh::New(mesh);
    ssam_viscosity_params p;
    std::memset(&p, 0, sizeof(p));
    fill(p);
    if (p.model == SSAM_VISC_NORTON_HOFF)
    {
        gpu.uploadU(U_);  // NortonHoff takes its strain rate from U itself: strainRate() = sqrt(2)*mag(symm(grad(U)))
    }
    else
    {
        gpu.upload(SSAM_F_T, mesh.lookupObject<volScalarField>("temp"));
        gpu.upload(SSAM_F_EPSDOT, mesh.lookupObject<volScalarField>("epsilonDotEq"));
    }
    gpu.check(ssam_viscosity_correct(gpu.ctx(), phase(), &p), "viscosityModel::correct");
    if (p.model == SSAM_VISC_NORTON_HOFF) gpu.deviceWrote(SSAM_F_STRAINRATE);
    gpu.download(phase() == 1 ? SSAM_F_NU1 : SSAM_F_NU2, nu_);
    if (sigmaf_.valid())
    {
        gpu.download(SSAM_F_SIGMAF, sigmaf_());
        gpu.download(SSAM_F_SIGMAY, sigmay_());
    }
}


void Foam::viscosityModels::gpuViscoPlasticBase::params(ssam_viscosity_params& p) const
{
    std::memset(&p, 0, sizeof(p));
    fill(p);
}


bool Foam::viscosityModels::gpuViscoPlasticBase::read(const dictionary& viscosityProperties)
{
    viscosityModel::read(viscosityProperties);
    coeffs_ = viscosityProperties.optionalSubDict(type() + "Coeffs");
    return true;
}


// ---- SheppardWright: sigmaR, Q, R, A, n, rho; limiters looked up on every call with the reference's defaults ----

Foam::viscosityModels::SheppardWright::SheppardWright
(
    const word& name,
    const dictionary& viscosityProperties,
    const volVectorField& U,
    const surfaceScalarField& phi
)
:
    gpuViscoPlasticBase(name, viscosityProperties, U, phi, typeName + "Coeffs", true)
{
    gpuCorrect();
}

void Foam::viscosityModels::SheppardWright::fill(ssam_viscosity_params& p) const
{
    p.model = SSAM_VISC_SHEPPARD_WRIGHT;
    p.sigmaR = dimensionedScalar("sigmaR", dimPressure, coeffs_).value();
    p.Q = dimensionedScalar("Q", dimEnergy/dimMoles, coeffs_).value();
    p.R = dimensionedScalar("R", dimEnergy/dimMoles/dimTemperature, coeffs_).value();
    p.A = dimensionedScalar("A", dimless/dimTime, coeffs_).value();
    p.n = dimensionedScalar("n", dimless, coeffs_).value();
    p.rho = dimensionedScalar("rho", dimDensity, coeffs_).value();
    p.nuMin = coeffs_.lookupOrDefault("nuMin", ROOTVSMALL);
    p.nuMax = coeffs_.lookupOrDefault("nuMax", ROOTVGREAT);
    p.eDotMin = coeffs_.lookupOrDefault("eDotMin", ROOTVSMALL);
}


// ---- FKS ---------------------------------------------------------------------------------------------------------

Foam::viscosityModels::FKS::FKS
(
    const word& name,
    const dictionary& viscosityProperties,
    const volVectorField& U,
    const surfaceScalarField& phi
)
:
    gpuViscoPlasticBase(name, viscosityProperties, U, phi, typeName + "Coeffs", true)
{
    gpuCorrect();
}

void Foam::viscosityModels::FKS::fill(ssam_viscosity_params& p) const
{
    p.model = SSAM_VISC_FKS;
    p.a1 = dimensionedScalar("a1", dimPressure, coeffs_).value();
    p.a2 = dimensionedScalar("a2", dimPressure, coeffs_).value();
    p.b1 = dimensionedScalar("b1", dimless, coeffs_).value();
    p.b2 = dimensionedScalar("b2", dimless, coeffs_).value();
    p.b3 = dimensionedScalar("b3", dimless, coeffs_).value();
    p.e0 = dimensionedScalar("e0", dimless/dimTime, coeffs_).value();
    p.c1 = dimensionedScalar("c1", dimless, coeffs_).value();
    p.c2 = dimensionedScalar("c2", dimless, coeffs_).value();
    p.Tm = dimensionedScalar("Tm", dimTemperature, coeffs_).value();
    p.Ta = dimensionedScalar("Ta", dimTemperature, coeffs_).value();
    p.rho = dimensionedScalar("rho", dimDensity, coeffs_).value();
    p.nuMin = coeffs_.lookupOrDefault("nuMin", ROOTVSMALL);
    p.nuMax = coeffs_.lookupOrDefault("nuMax", ROOTVGREAT);
    p.eDotMin = coeffs_.lookupOrDefault("eDotMin", ROOTVSMALL);
}


// ---- NortonHoff --------------------------------------------------------------------------------------------------

Foam::viscosityModels::NortonHoff::NortonHoff
(
    const word& name,
    const dictionary& viscosityProperties,
    const volVectorField& U,
    const surfaceScalarField& phi
)
:
    gpuViscoPlasticBase(name, viscosityProperties, U, phi, typeName + "Coeffs", false)
{
    gpuCorrect();
}

void Foam::viscosityModels::NortonHoff::fill(ssam_viscosity_params& p) const
{
    p.model = SSAM_VISC_NORTON_HOFF;
    p.mu0 = dimensionedScalar("mu0", dimViscosity*dimDensity, coeffs_).value();
    p.m = dimensionedScalar("m", dimless, coeffs_).value();
    p.rho = dimensionedScalar("rho", dimDensity, coeffs_).value();
    p.nuMin = dimensionedScalar("nuMin", dimViscosity, coeffs_).value();
    p.nuMax = dimensionedScalar("nuMax", dimViscosity, coeffs_).value();
    p.strainRateEffMin = dimensionedScalar("strainRateEffMin", dimless/dimTime, coeffs_).value();
    p.eDotMin = ROOTVSMALL;
}
