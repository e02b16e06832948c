// ssam_case_driver: replays chtMultiRegionInterFoam's call sequence for the hot path
// through the host-side mirror classes, on a case directory laid out like an OpenFOAM case:
//
//   <case>/constant/transportProperties     phases (...); <phase> { transportModel ...; ...Coeffs {...} rho cp kModel ... }
//   <case>/constant/frictionProperties      stickModel ...; alphaName pName fricPatch; ...Coeffs {...}
//   <case>/0/U                              boundaryField { <patch> { type constantRotationalSlip; omega (...); ... } }
//   <case>/mesh.bin, fields.bin, patches.txt binary mesh / field dump written by solidstateamfoam_b200.cases
// or, when <case>/constant/polyMesh/points exists, a real OpenFOAM case:
//   <case>/constant/polyMesh/{points,faces,owner,neighbour,boundary}   ascii or binary
//   <case>/0/{U,temp,alpha.<phase1>,<pName>}                           vol fields; U's boundaryField also
//                                                                     selects the rotational-slip BCs
//   results are then also written as OpenFOAM fields into <case>/<nSteps>/
//
// usage: ssam_case_driver <case dir> <nSteps> <out.bin> [device]
// Per step (nOuterCorrectors 1, one alpha sub-cycle): U BC updateCoeffs(); thermo.correct()
// (alphaEqn.H:219); rho/rhoCp (alphaEqnSubCycle.H:41-42); thermo.correct() (solveFluid.H:52);
// TEqn.H:3-15: epsilonDotEq, muEff = thermo.mu(), friction.correct(), cp, k.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>

#include "../meshgen/meshgen.h"
#include "ssamFoam.H"

using namespace ssam;

template <class T>
static std::vector<T> readArr(std::ifstream& in, size_t n) {
    std::vector<T> v(n);
    if (n) in.read(reinterpret_cast<char*>(v.data()), std::streamsize(n * sizeof(T)));
    if (!in) throw FatalError("mesh/field file truncated");
    return v;
}

int main(int argc, char** argv) {
    if (argc < 4) {
        std::fprintf(stderr, "usage: %s <case dir> <nSteps> <out.bin> [device]\n", argv[0]);
        return 2;
    }
    const std::string dir = argv[1];
    const int nSteps = std::atoi(argv[2]);
    const std::string outPath = argv[3];
    const int device = argc > 4 ? std::atoi(argv[4]) : 0;
    try {
        // ---- dictionaries ----
        const dictionary transportProperties = dictionary::fromFile(dir + "/constant/transportProperties");
        const dictionary frictionProperties = dictionary::fromFile(dir + "/constant/frictionProperties");
        const dictionary Udict = dictionary::fromFile(dir + "/0/U");
        const word phase1 = transportProperties.lookupWordList("phases")[0];

        // ---- mesh + fields ----
        fvMeshGpu mesh(device);
        std::vector<word> patchNames;
        int nc = 0, nif = 0, nf = 0, np = 0, nb = 0;
        smg_mesh* foamMesh = nullptr;
        const bool foamCase = bool(std::ifstream(dir + "/constant/polyMesh/points"));
        struct In { int field; int ncomp; const char* name; };
        const In ins[] = {{SSAM_F_U, 3, "U"}, {SSAM_F_T, 1, "temp"}, {SSAM_F_ALPHA1, 1, nullptr}, {SSAM_F_P, 1, nullptr},
                          {SSAM_F_EPSDOT, 1, "epsilonDotEq"}, {SSAM_F_MUEFF, 1, "muEff"}};
        if (foamCase) {
            foamMesh = smg_read_polymesh((dir + "/constant/polyMesh").c_str());
            if (!foamMesh) throw FatalError(smg_io_error());
            int32_t sz[4];
            smg_sizes(foamMesh, sz);
            nc = sz[0]; nif = sz[1]; nf = sz[2]; np = sz[3]; nb = nf - nif;
            for (int p = 0; p < np; ++p) patchNames.push_back(smg_patch_name(foamMesh, p));
            const word names[4] = {"U", "temp", "alpha." + phase1, frictionProperties.lookupWord("pName")};
            std::vector<int32_t> ubc(np, SSAM_BC_VALUE), bcTmp(np);
            std::vector<std::vector<double>> fi(4), fb(4);
            for (int i = 0; i < 4; ++i) {
                fi[i].resize(size_t(nc) * ins[i].ncomp);
                fb[i].resize(size_t(nb) * ins[i].ncomp);
                if (smg_read_field(foamMesh, (dir + "/0/" + names[i]).c_str(), ins[i].ncomp, fi[i].data(), fb[i].data(),
                                   i == 0 ? ubc.data() : bcTmp.data()))
                    throw FatalError(smg_io_error());
            }
            ssam_mesh_desc md{nc, nif, nf, np, smg_owner(foamMesh), smg_neighbour(foamMesh), smg_patch_start(foamMesh),
                              smg_patch_size(foamMesh), smg_patch_kind(foamMesh), smg_patch_nbr_rank(foamMesh),
                              smg_Sf(foamMesh), smg_weights(foamMesh), smg_V(foamMesh), smg_C(foamMesh),
                              smg_Cf(foamMesh) + 3 * size_t(nif), smg_magSf(foamMesh) + nif,
                              smg_delta_coeffs(foamMesh) + nif};
            mesh.setMesh(md, patchNames, ubc.data());
            for (int i = 0; i < 4; ++i) {
                mesh.check(ssam_field_upload(mesh.ctx(), ins[i].field, fi[i].data(), fb[i].data()), "ssam_field_upload");
                if (ins[i].name) mesh.checkIn(ins[i].name, ins[i].field);
            }
            // epsilonDotEq / muEff as createFluidFields.H:129-164 creates them: uniform initial value
            std::vector<double> one_i(nc, 1.0), one_b(nb, 1.0);
            for (int i = 4; i < 6; ++i) {
                mesh.check(ssam_field_upload(mesh.ctx(), ins[i].field, one_i.data(), one_b.data()), "ssam_field_upload");
                mesh.checkIn(ins[i].name, ins[i].field);
            }
        } else {
            std::ifstream mf(dir + "/mesh.bin", std::ios::binary);
            if (!mf) throw FatalError("cannot open " + dir + "/mesh.bin");
            auto hdr = readArr<int32_t>(mf, 4);
            nc = hdr[0]; nif = hdr[1]; nf = hdr[2]; np = hdr[3]; nb = nf - nif;
            auto owner = readArr<int32_t>(mf, nf), neighbour = readArr<int32_t>(mf, nif);
            auto pStart = readArr<int32_t>(mf, np), pSize = readArr<int32_t>(mf, np), pKind = readArr<int32_t>(mf, np),
                 pNbr = readArr<int32_t>(mf, np), pUbc = readArr<int32_t>(mf, np);
            auto Sf = readArr<double>(mf, 3 * size_t(nf)), w = readArr<double>(mf, nf), V = readArr<double>(mf, nc),
                 C = readArr<double>(mf, 3 * size_t(nc)), Cfb = readArr<double>(mf, 3 * size_t(nb)),
                 magSfb = readArr<double>(mf, nb), dcb = readArr<double>(mf, nb);
            {
                std::ifstream pf(dir + "/patches.txt");
                word s;
                while (pf >> s) patchNames.push_back(s);
            }
            ssam_mesh_desc md{nc, nif, nf, np, owner.data(), neighbour.data(), pStart.data(), pSize.data(), pKind.data(),
                              pNbr.data(), Sf.data(), w.data(), V.data(), C.data(), Cfb.data(), magSfb.data(), dcb.data()};
            mesh.setMesh(md, patchNames, pUbc.data());
            // U, temp, alpha.<phase1>, p, initial epsilonDotEq / muEff (createFluidFields.H:76-164)
            std::ifstream ff(dir + "/fields.bin", std::ios::binary);
            if (!ff) throw FatalError("cannot open " + dir + "/fields.bin");
            for (const In& in : ins) {
                auto fi = readArr<double>(ff, size_t(nc) * in.ncomp), fb = readArr<double>(ff, size_t(nb) * in.ncomp);
                mesh.check(ssam_field_upload(mesh.ctx(), in.field, fi.data(), fb.data()), "ssam_field_upload");
                if (in.name) mesh.checkIn(in.name, in.field);
            }
        }

        // ---- velocity BCs (fvPatchField RTS on the patch "type") ----
        std::unique_ptr<constantRotationalSlipFvPatchVectorField> bcConst;
        std::unique_ptr<exponentialRotationalSlipFvPatchVectorField> bcExp;
        const dictionary& bf = Udict.subDict("boundaryField");
        for (int p = 0; p < np; ++p) {
            if (!bf.isDict(patchNames[p])) continue;
            const dictionary& pd = bf.subDict(patchNames[p]);
            const word type = pd.lookupWord("type");
            if (type == constantRotationalSlipFvPatchVectorField::typeName_())
                bcConst.reset(new constantRotationalSlipFvPatchVectorField(mesh, p, pd));
            else if (type == exponentialRotationalSlipFvPatchVectorField::typeName_())
                bcExp.reset(new exponentialRotationalSlipFvPatchVectorField(mesh, p, pd));
        }
        if (bcConst) bcConst->updateCoeffs();
        if (bcExp) bcExp->updateCoeffs();

        // ---- models, in createFluidFields.H order: mixture (:168-172), friction model (:502-506) ----
        immiscibleIncompressibleTwoPhaseThermalMixture thermo(mesh, transportProperties);
        incompressibleFrictionModel friction(mesh, frictionProperties);
        std::printf("viscosity models: %s / %s\n", thermo.nuModel1().type().c_str(), thermo.nuModel2().type().c_str());
        const bool hasSigmay = mesh.foundObject("sigmay");

        // ---- time loop ----
        for (int step = 0; step < nSteps; ++step) {
            if (bcConst) bcConst->updateCoeffs();
            if (bcExp) bcExp->updateCoeffs();
            thermo.correct();                                     // alphaEqn.H:219
            mesh.check(ssam_solver_rho_rhocp(mesh.ctx(), &thermo.params()), "rho/rhoCp");  // alphaEqnSubCycle.H:41-42
            thermo.correct();                                     // solveFluid.H:52
            mesh.check(ssam_strain_rate(mesh.ctx(), std::sqrt(2.0 / 3.0), SSAM_F_EPSDOT), "epsilonDotEq");  // TEqn.H:3
            thermo.mu();                                          // TEqn.H:6   muEff = thermo.mu()
            mesh.checkIn("muEff", SSAM_F_MUEFF);
            if (hasSigmay)
                friction.correct();                               // TEqn.H:9
            else
                std::printf("friction.correct() skipped: no sigmay registered (NortonHoff, SURVEY D-1)\n");
            thermo.cp();                                          // TEqn.H:12
            thermo.k();                                           // TEqn.H:15
            double tMin = 0, tMax = 0;
            mesh.minMax(SSAM_F_T, tMin, tMax);                    // TEqn.H:42-43
            std::printf("Min/max T:%.17g %.17g\n", tMin, tMax);
        }

        // ---- write the AUTO_WRITE fields ----
        std::ofstream out(outPath, std::ios::binary);
        std::vector<int> outs = {SSAM_F_EPSDOT, SSAM_F_NU1, SSAM_F_NU2, SSAM_F_NU, SSAM_F_MUEFF, SSAM_F_CP, SSAM_F_K,
                                 SSAM_F_RHO_SOLVER, SSAM_F_RHOCP, SSAM_F_CURVATURE};
        if (hasSigmay) {
            outs.push_back(SSAM_F_SIGMAF);
            outs.push_back(SSAM_F_SIGMAY);
            outs.push_back(SSAM_F_QVISC);
            outs.push_back(SSAM_F_QFRIC);
        }
        for (int f : outs) {
            std::vector<double> fi, fb;
            volScalarFieldRef{&mesh, f}.download(fi, fb);
            const int32_t id = f;
            out.write(reinterpret_cast<const char*>(&id), sizeof(id));
            out.write(reinterpret_cast<const char*>(fi.data()), std::streamsize(fi.size() * sizeof(double)));
            out.write(reinterpret_cast<const char*>(fb.data()), std::streamsize(fb.size() * sizeof(double)));
        }
        if (foamCase) {
            // AUTO_WRITE: the same fields as OpenFOAM volScalarFields into <case>/<nSteps>/
            const char* foamNames[SSAM_F_COUNT] = {"U", "temp", "alpha1", "p", "epsilonDotEq", "nu1", "nu2", "sigmaf",
                                                    "sigmay", "nu", "muEff", "rho", "cp", "k", "qvisc", "qfric",
                                                    "gradU", "strainRate", "rho", "rhoCp"};
            foamNames[SSAM_F_CURVATURE] = "interfaceProperties:K";
            for (int f : outs) {
                std::vector<double> fi, fb;
                volScalarFieldRef{&mesh, f}.download(fi, fb);
                const std::string path = dir + "/" + std::to_string(nSteps) + "/" + foamNames[f];
                if (smg_write_field(foamMesh, path.c_str(), foamNames[f], 1, fi.data(), fb.data(), nullptr, 0))
                    throw FatalError(smg_io_error());
            }
            smg_free(foamMesh);
        }
        std::printf("End\n");
    } catch (const FatalError& e) {
        std::fprintf(stderr, "--> FOAM FATAL ERROR:\n%s\n", e.what());
        return 1;
    }
    return 0;
}
