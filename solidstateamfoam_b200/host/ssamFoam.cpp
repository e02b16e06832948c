// Implementation of the host-side mirror (see ssamFoam.H).  Every method is a thin forwarder to one
// C-ABI call; the arithmetic lives in the CUDA kernels only.
#include "ssamFoam.H"

#include <cctype>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>

namespace ssam {

namespace {
constexpr double ROOTVSMALL = 1e-150, ROOTVGREAT = 1e150;

std::string sortedToc(const std::vector<word>& keys) {
    std::ostringstream os;
    os << keys.size() << "(";
    for (size_t i = 0; i < keys.size(); ++i) os << (i ? " " : "") << keys[i];
    os << ")";
    return os.str();
}
}  // namespace

// ---- dictionary parser ----------------------------------------------------------------------------------------
class dictParser {
public:
    explicit dictParser(const std::string& text) {
        // "format binary;" in the FoamFile header: nonuniform List<T> payloads are raw bytes, skipped here
        const size_t h = text.find("FoamFile");
        if (h != std::string::npos) {
            const size_t e = text.find('}', h);
            binary_ = e != std::string::npos && text.substr(h, e - h).find("binary") != std::string::npos;
        }
        tokenize(text);
    }
    void parseInto(dictionary& d, bool top) {
        while (pos_ < tok_.size()) {
            const std::string t = tok_[pos_];
            if (t == "}") {
                if (top) throw FatalError("dictionary: unexpected '}'");
                ++pos_;
                return;
            }
            const word key = t;
            ++pos_;
            if (pos_ < tok_.size() && tok_[pos_] == "{") {
                ++pos_;
                dictionary sub;
                sub.name_ = d.name_ + "." + key;
                parseInto(sub, false);
                d.subs_[key] = sub;
                continue;
            }
            std::vector<std::string> vals;
            while (pos_ < tok_.size() && tok_[pos_] != ";") vals.push_back(tok_[pos_++]);
            if (pos_ >= tok_.size()) throw FatalError("dictionary: missing ';' after keyword " + key);
            ++pos_;
            d.entries_[key] = vals;
        }
        if (!top) throw FatalError("dictionary: missing '}'");
    }

private:
    void tokenize(const std::string& s) {
        size_t i = 0;
        while (i < s.size()) {
            const char c = s[i];
            if (std::isspace((unsigned char)c)) {
                ++i;
            } else if (c == '/' && i + 1 < s.size() && s[i + 1] == '/') {
                while (i < s.size() && s[i] != '\n') ++i;
            } else if (c == '/' && i + 1 < s.size() && s[i + 1] == '*') {
                i = s.find("*/", i + 2);
                i = (i == std::string::npos) ? s.size() : i + 2;
            } else if (c == '{' || c == '}' || c == '(' || c == ')' || c == ';') {
                tok_.push_back(std::string(1, c));
                ++i;
                if (binary_ && c == '(' && tok_.size() >= 3) {
                    // ... List<vector> N ( <N*ncomp*8 raw bytes> )
                    const std::string& type = tok_[tok_.size() - 3];
                    int nc = type == "List<scalar>" ? 1 : type == "List<vector>" ? 3 : type == "List<tensor>" ? 9 : 0;
                    char* end = nullptr;
                    const long n = std::strtol(tok_[tok_.size() - 2].c_str(), &end, 10);
                    if (nc && *end == '\0' && n >= 0) i += size_t(n) * nc * 8;
                }
            } else {
                size_t j = i;
                while (j < s.size() && !std::isspace((unsigned char)s[j]) && std::string("{}();").find(s[j]) == std::string::npos)
                    ++j;
                tok_.push_back(s.substr(i, j - i));
                i = j;
            }
        }
    }
    std::vector<std::string> tok_;
    size_t pos_ = 0;
    bool binary_ = false;
};

dictionary dictionary::fromString(const std::string& text, const std::string& name) {
    dictionary d;
    d.name_ = name;
    dictParser p(text);
    p.parseInto(d, true);
    // FoamFile header, if any, is just another sub-dictionary
    return d;
}

dictionary dictionary::fromFile(const std::string& path) {
    std::ifstream in(path);
    if (!in) throw FatalError("cannot open file " + path);
    std::stringstream ss;
    ss << in.rdbuf();
    return fromString(ss.str(), path);
}

static const std::vector<std::string>& tokensOf(const std::map<word, std::vector<std::string>>& e, const word& key,
                                                const std::string& dictName) {
    auto it = e.find(key);
    if (it == e.end())
        throw FatalError("keyword " + key + " is undefined in dictionary \"" + dictName + "\"");
    return it->second;
}

double dictionary::lookupScalar(const word& key) const {
    const auto& t = tokensOf(entries_, key, name_);
    // dimensionedScalar entries may read "name [dims] value" or "[dims] value" or "value": take the last number
    for (auto it = t.rbegin(); it != t.rend(); ++it) {
        char* end = nullptr;
        const double v = std::strtod(it->c_str(), &end);
        if (end != it->c_str() && *end == '\0') return v;
    }
    throw FatalError("keyword " + key + " in dictionary \"" + name_ + "\" is not a scalar");
}

word dictionary::lookupWord(const word& key) const {
    const auto& t = tokensOf(entries_, key, name_);
    if (t.empty()) throw FatalError("keyword " + key + " has no value in dictionary \"" + name_ + "\"");
    return t[0];
}

vector dictionary::lookupVector(const word& key) const {
    const auto& t = tokensOf(entries_, key, name_);
    std::vector<double> v;
    for (const auto& s : t) {
        char* end = nullptr;
        const double x = std::strtod(s.c_str(), &end);
        if (end != s.c_str() && *end == '\0') v.push_back(x);
    }
    if (v.size() < 3) throw FatalError("keyword " + key + " in dictionary \"" + name_ + "\" is not a vector");
    return {v[v.size() - 3], v[v.size() - 2], v[v.size() - 1]};
}

std::vector<word> dictionary::lookupWordList(const word& key) const {
    std::vector<word> out;
    for (const auto& s : tokensOf(entries_, key, name_))
        if (s != "(" && s != ")") out.push_back(s);
    return out;
}

const dictionary& dictionary::subDict(const word& key) const {
    auto it = subs_.find(key);
    if (it == subs_.end()) throw FatalError("keyword " + key + " is undefined in dictionary \"" + name_ + "\" (subDict)");
    return it->second;
}

// ---- mesh / registry ---------------------------------------------------------------------------------------------
fvMeshGpu::fvMeshGpu(int device) {
    ssam_status st = ssam_create(&ctx_, device);
    if (st != SSAM_OK) throw FatalError(std::string("ssam_create: ") + ssam_last_error(nullptr));
    // the registry names of the reference (createFluidFields.H:111-164, SheppardWright.C:55-58)
    registry_["U"] = SSAM_F_U;
}
fvMeshGpu::~fvMeshGpu() { ssam_destroy(ctx_); }

void fvMeshGpu::check(ssam_status st, const char* where) const {
    if (st != SSAM_OK) throw FatalError(std::string(where) + ": " + ssam_last_error(ctx_));
}

void fvMeshGpu::setMesh(const ssam_mesh_desc& m, const std::vector<word>& patchNames, const int32_t* patchUbc) {
    check(ssam_mesh_set(ctx_, &m), "ssam_mesh_set");
    if (patchUbc) check(ssam_patch_u_bc_set(ctx_, patchUbc), "ssam_patch_u_bc_set");
    patchNames_ = patchNames;
    nCells_ = m.nCells;
    nBnd_ = m.nFaces - m.nInternalFaces;
}

int fvMeshGpu::findPatchID(const word& name) const {
    for (size_t i = 0; i < patchNames_.size(); ++i)
        if (patchNames_[i] == name) return int(i);
    return -1;
}

int fvMeshGpu::lookupObject(const word& name) const {
    auto it = registry_.find(name);
    if (it == registry_.end() || !ssam_field_is_set(ctx_, it->second))
        throw FatalError("request for volScalarField " + name + " from objectRegistry region0 failed");
    return it->second;
}

void volScalarFieldRef::download(std::vector<double>& internal, std::vector<double>& boundary) const {
    internal.resize(mesh->nCells());
    boundary.resize(mesh->nBoundaryFaces());
    mesh->check(ssam_field_download(mesh->ctx(), field, internal.data(), boundary.data()), "ssam_field_download");
}

// ---- interfaceProperties -------------------------------------------------------------------------------------------
double interfaceProperties::deltaN() const {
    if (deltaN_ < 0) imesh_.check(ssam_interface_delta_n(imesh_.ctx(), &deltaN_), "interfaceProperties::deltaN");
    return deltaN_;
}

void interfaceProperties::correct() {
    imesh_.check(ssam_interface_correct(imesh_.ctx(), deltaN(), 0), "interfaceProperties::correct");
}

void interfaceProperties::nHatf(std::vector<double>& internalFaces, std::vector<double>& boundaryFaces,
                                int nInternalFaces) const {
    internalFaces.resize(size_t(nInternalFaces));
    boundaryFaces.resize(size_t(imesh_.nBoundaryFaces()));
    imesh_.check(ssam_interface_nhatf_download(imesh_.ctx(), internalFaces.data(), boundaryFaces.data()),
                 "interfaceProperties::nHatf");
}

// ---- viscosityModel ------------------------------------------------------------------------------------------------
std::map<word, viscosityModel::ctor>& viscosityModel::dictionaryConstructorTable() {
    static std::map<word, ctor> table = {
#define SSAM_ADD(Name)                                                                                  \
    {#Name, [](const word& n, const dictionary& d, fvMeshGpu& m, int ph) {                              \
         return std::unique_ptr<viscosityModel>(new viscosityModels::Name(n, d, m, ph));                \
     }},
        SSAM_ADD(SheppardWright) SSAM_ADD(FKS) SSAM_ADD(NortonHoff) SSAM_ADD(Newtonian)
#undef SSAM_ADD
    };
    return table;
}

std::unique_ptr<viscosityModel> viscosityModel::New(const word& name, const dictionary& viscosityProperties,
                                                    fvMeshGpu& mesh, int phase) {
    const word modelType(viscosityProperties.lookupWord("transportModel"));
    auto& table = dictionaryConstructorTable();
    auto it = table.find(modelType);
    if (it == table.end()) {
        std::vector<word> keys;
        for (auto& kv : table) keys.push_back(kv.first);
        throw FatalError("Unknown viscosityModel type " + modelType + "\n\nValid viscosityModels :\n" + sortedToc(keys));
    }
    return it->second(name, viscosityProperties, mesh, phase);
}

void viscosityModel::correct() {
    // limiters are looked up on every call with their defaults (SheppardWright.C:104-106, FKS.C:165-167)
    if (p_.model == SSAM_VISC_SHEPPARD_WRIGHT || p_.model == SSAM_VISC_FKS) {
        const dictionary& c = viscosityProperties_.optionalSubDict(type() + "Coeffs");
        p_.nuMin = c.lookupOrDefault("nuMin", ROOTVSMALL);
        p_.nuMax = c.lookupOrDefault("nuMax", ROOTVGREAT);
        p_.eDotMin = c.lookupOrDefault("eDotMin", ROOTVSMALL);
        mesh_.lookupObject("temp");          // lookupObject<volScalarField>("temp")
        mesh_.lookupObject("epsilonDotEq");  // lookupObject<volScalarField>("epsilonDotEq")
    }
    mesh_.check(ssam_viscosity_correct(mesh_.ctx(), phase_, &p_), "viscosityModel::correct");
    if (p_.model == SSAM_VISC_SHEPPARD_WRIGHT || p_.model == SSAM_VISC_FKS) {
        mesh_.checkIn("sigmaf", SSAM_F_SIGMAF);  // registered AUTO_WRITE fields (SheppardWright.C:153-177)
        mesh_.checkIn("sigmay", SSAM_F_SIGMAY);
    }
}

namespace viscosityModels {

SheppardWright::SheppardWright(const word& name, const dictionary& d, fvMeshGpu& mesh, int phase)
    : viscosityModel(name, d, mesh, phase) {
    p_.model = SSAM_VISC_SHEPPARD_WRIGHT;
    const dictionary& c = d.optionalSubDict("SheppardWrightCoeffs");
    p_.sigmaR = c.lookupScalar("sigmaR");
    p_.Q = c.lookupScalar("Q");
    p_.R = c.lookupScalar("R");
    p_.A = c.lookupScalar("A");
    p_.n = c.lookupScalar("n");
    p_.rho = c.lookupScalar("rho");
    correct();  // nu_(calcNu()), sigmaf_(sigmaf()), sigmay_(sigmay()) in the constructor (:140-177)
}
bool SheppardWright::read(const dictionary& d) {
    viscosityProperties_ = d;
    const dictionary& c = d.optionalSubDict("SheppardWrightCoeffs");
    p_.sigmaR = c.lookupScalar("sigmaR");
    p_.Q = c.lookupScalar("Q");
    p_.R = c.lookupScalar("R");
    p_.A = c.lookupScalar("A");
    p_.n = c.lookupScalar("n");  // rho is NOT re-read (SheppardWright.C:192-196)
    return true;
}

FKS::FKS(const word& name, const dictionary& d, fvMeshGpu& mesh, int phase) : viscosityModel(name, d, mesh, phase) {
    p_.model = SSAM_VISC_FKS;
    read(d);
    correct();
}
bool FKS::read(const dictionary& d) {
    viscosityProperties_ = d;
    const dictionary& c = d.optionalSubDict("FKSCoeffs");
    p_.a1 = c.lookupScalar("a1");
    p_.a2 = c.lookupScalar("a2");
    p_.b1 = c.lookupScalar("b1");
    p_.b2 = c.lookupScalar("b2");
    p_.b3 = c.lookupScalar("b3");
    p_.e0 = c.lookupScalar("e0");
    p_.c1 = c.lookupScalar("c1");
    p_.c2 = c.lookupScalar("c2");
    p_.Tm = c.lookupScalar("Tm");
    p_.Ta = c.lookupScalar("Ta");
    p_.rho = c.lookupScalar("rho");
    return true;
}

NortonHoff::NortonHoff(const word& name, const dictionary& d, fvMeshGpu& mesh, int phase)
    : viscosityModel(name, d, mesh, phase) {
    p_.model = SSAM_VISC_NORTON_HOFF;
    read(d);
    correct();
}
bool NortonHoff::read(const dictionary& d) {
    viscosityProperties_ = d;
    const dictionary& c = d.optionalSubDict("NortonHoffCoeffs");
    p_.mu0 = c.lookupScalar("mu0");
    p_.m = c.lookupScalar("m");
    p_.rho = c.lookupScalar("rho");
    p_.nuMin = c.lookupScalar("nuMin");  // all required (NortonHoff.C:82-87)
    p_.nuMax = c.lookupScalar("nuMax");
    p_.strainRateEffMin = c.lookupScalar("strainRateEffMin");
    return true;
}

Newtonian::Newtonian(const word& name, const dictionary& d, fvMeshGpu& mesh, int phase)
    : viscosityModel(name, d, mesh, phase) {
    p_.model = SSAM_VISC_NEWTONIAN;
    read(d);
    correct();
}
bool Newtonian::read(const dictionary& d) {
    viscosityProperties_ = d;
    p_.nu = d.lookupScalar("nu");
    return true;
}

}  // namespace viscosityModels

// ---- mixture -----------------------------------------------------------------------------------------------------------
incompressibleTwoPhaseThermalMixture::incompressibleTwoPhaseThermalMixture(fvMeshGpu& mesh,
                                                                           const dictionary& transportProperties)
    : mesh_(mesh), dict_(transportProperties) {
    mix_ = ssam_mixture_params();
    const std::vector<word> phases = dict_.lookupWordList("phases");  // twoPhaseMixture: phases (p1 p2);
    if (phases.size() != 2) throw FatalError("twoPhaseMixture: expected phases (phase1 phase2)");
    phase1Name_ = phases[0];
    phase2Name_ = phases[1];
    mesh_.checkIn("alpha." + phase1Name_, SSAM_F_ALPHA1);
    nuModel1_ = viscosityModel::New("nu1", dict_.subDict(phase1Name_), mesh_, 1);
    nuModel2_ = viscosityModel::New("nu2", dict_.subDict(phase2Name_), mesh_, 2);
    mix_.rho1 = nuModel1_->viscosityProperties().lookupScalar("rho");
    mix_.rho2 = nuModel2_->viscosityProperties().lookupScalar("rho");
    mix_.cp1 = nuModel1_->viscosityProperties().lookupScalar("cp");
    mix_.cp2 = nuModel2_->viscosityProperties().lookupScalar("cp");
    calcNu();
}

void incompressibleTwoPhaseThermalMixture::calcNu() {
    nuModel1_->correct();
    nuModel2_->correct();
    mesh_.check(ssam_mixture_calc_nu(mesh_.ctx(), &mix_), "incompressibleTwoPhaseThermalMixture::calcNu");
    mesh_.checkIn("nu", SSAM_F_NU);
}

volScalarFieldRef incompressibleTwoPhaseThermalMixture::mu() const {
    mesh_.check(ssam_mixture_mu(mesh_.ctx(), &mix_), "mu()");
    return {&mesh_, SSAM_F_MUEFF};
}

static void readConductivity(const dictionary& d, ssam_conductivity& k, int phase) {
    const word kModel = d.lookupOrDefault("kModel", word("constant"));
    k = ssam_conductivity();
    if (kModel == "cubic") {  // any other word selects the constant model (...Mixture.C:212,256)
        k.kModel = SSAM_K_CUBIC;
        k.k0 = d.lookupScalar("k0");
        k.k1 = d.lookupScalar("k1");
        k.k2 = d.lookupScalar("k2");
        k.k3 = d.lookupScalar("k3");
        k.Tmax = d.lookupScalar("Tmax");
        k.Tmin = d.lookupScalar("Tmin");
        const word units = d.lookupOrDefault("units", word("Kelvin"));
        if (units == "Kelvin")
            k.units = SSAM_UNITS_KELVIN;
        else if (units == "Celsius")
            k.units = SSAM_UNITS_CELSIUS;
        else
            throw FatalError("Unknown conductivity function units " + units + " for phase " + std::to_string(phase) +
                             ". Choose Kelvin or Celsius.");
    } else {
        k.kModel = SSAM_K_CONSTANT;
        k.k = d.lookupScalar("k");
    }
}

static void faceVisc(const fvMeshGpu& mesh, const ssam_mixture_params& mix, bool nuf, std::vector<double>& fi,
                     std::vector<double>& fb, int nInternalFaces) {
    mesh.check(nuf ? ssam_mixture_nuf(mesh.ctx(), &mix) : ssam_mixture_muf(mesh.ctx(), &mix), nuf ? "nuf()" : "muf()");
    fi.resize(size_t(nInternalFaces));
    fb.resize(size_t(mesh.nBoundaryFaces()));
    mesh.check(ssam_face_field_download(mesh.ctx(), nuf ? SSAM_FF_NUF : SSAM_FF_MUF, fi.data(), fb.data()),
               "ssam_face_field_download");
}
void incompressibleTwoPhaseThermalMixture::muf(std::vector<double>& fi, std::vector<double>& fb, int nIF) const {
    faceVisc(mesh_, mix_, false, fi, fb, nIF);
}
void incompressibleTwoPhaseThermalMixture::nuf(std::vector<double>& fi, std::vector<double>& fb, int nIF) const {
    faceVisc(mesh_, mix_, true, fi, fb, nIF);
}

void incompressibleTwoPhaseThermalMixture::readThermal() {
    // k1()/k2() look their constants up on every call (:215-223,:261)
    readConductivity(nuModel1_->viscosityProperties(), mix_.k1, 1);
    readConductivity(nuModel2_->viscosityProperties(), mix_.k2, 2);
}

volScalarFieldRef incompressibleTwoPhaseThermalMixture::k() const {
    mesh_.lookupObject("temp");
    const_cast<incompressibleTwoPhaseThermalMixture*>(this)->readThermal();
    mesh_.check(ssam_mixture_k(mesh_.ctx(), &mix_), "k()");
    return {&mesh_, SSAM_F_K};
}
volScalarFieldRef incompressibleTwoPhaseThermalMixture::cp() const {
    mesh_.check(ssam_mixture_cp(mesh_.ctx(), &mix_), "cp()");
    return {&mesh_, SSAM_F_CP};
}
volScalarFieldRef incompressibleTwoPhaseThermalMixture::rho() const {
    mesh_.check(ssam_mixture_rho(mesh_.ctx(), &mix_), "rho()");
    return {&mesh_, SSAM_F_RHO};
}

bool incompressibleTwoPhaseThermalMixture::read(const dictionary& transportProperties) {
    dict_ = transportProperties;
    if (!nuModel1_->read(dict_.subDict(phase1Name_)) || !nuModel2_->read(dict_.subDict(phase2Name_))) return false;
    mix_.rho1 = nuModel1_->viscosityProperties().lookupScalar("rho");
    mix_.rho2 = nuModel2_->viscosityProperties().lookupScalar("rho");
    nuModel1_->viscosityProperties().lookupWord("kModel");  // required by read() (:426-427)
    nuModel2_->viscosityProperties().lookupWord("kModel");
    mix_.cp1 = nuModel1_->viscosityProperties().lookupScalar("cp");
    mix_.cp2 = nuModel2_->viscosityProperties().lookupScalar("cp");
    return true;
}

// ---- stick models ------------------------------------------------------------------------------------------------------
std::map<word, stickModel::ctor>& stickModel::dictionaryConstructorTable() {
    static std::map<word, ctor> table = {
        {"constantSlipStick",
         [](const word& n, const dictionary& d, fvMeshGpu& m) {
             return std::unique_ptr<stickModel>(new stickModels::constantSlipStick(n, d, m));
         }},
        {"exponentialSlipStick",
         [](const word& n, const dictionary& d, fvMeshGpu& m) {
             return std::unique_ptr<stickModel>(new stickModels::exponentialSlipStick(n, d, m));
         }},
    };
    return table;
}

std::unique_ptr<stickModel> stickModel::New(const word& name, const dictionary& stickProperties, fvMeshGpu& mesh) {
    const word modelType(stickProperties.lookupWord("stickModel"));  // stickModelNew.C:40
    std::printf("Selecting stick model %s\n", modelType.c_str());
    auto& table = dictionaryConstructorTable();
    auto it = table.find(modelType);
    if (it == table.end()) {
        std::vector<word> keys;
        for (auto& kv : table) keys.push_back(kv.first);
        throw FatalError("Unknown stickModel type " + modelType + "\n\nValid stickModels :\n" + sortedToc(keys));
    }
    return it->second(name, stickProperties, mesh);
}

void stickModel::readCommon(const dictionary& top) {
    alphaName_ = top.lookupWord("alphaName");
    pName_ = top.lookupWord("pName");
    patch_ = top.lookupWord("fricPatch");
    mesh_.checkIn(pName_, SSAM_F_P);
    p_.fricPatch = mesh_.findPatchID(patch_);
}

void stickModel::correct() {
    // calcQvisc(): lookupObject "muEff", "epsilonDotEq"; calcQfric(): alphaName_, "sigmay", pName_
    mesh_.lookupObject("muEff");
    mesh_.lookupObject("epsilonDotEq");
    mesh_.lookupObject(alphaName_);
    mesh_.lookupObject("sigmay");
    mesh_.lookupObject(pName_);
    if (p_.fricPatch < 0)
        throw FatalError("fricPatch " + patch_ + " not found (the reference indexes boundaryMesh()[-1], SURVEY D-7)");
    mesh_.check(ssam_friction_correct(mesh_.ctx(), &p_), "stickModel::correct");
    double c[3], area;
    mesh_.check(ssam_friction_patch_centre(mesh_.ctx(), c, &area), "r()");
    std::printf("Center for patch %s found to be at (%.12g %.12g %.12g)\n", patch_.c_str(), c[0], c[1], c[2]);
}

namespace stickModels {

constantSlipStick::constantSlipStick(const word& name, const dictionary& d, fvMeshGpu& mesh) : stickModel(name, d, mesh) {
    p_.model = SSAM_STICK_CONSTANT;
    readCommon(d);  // top-level dictionary in the constructor (constantSlipStick.C:161-163)
    const dictionary& c = d.optionalSubDict("constantSlipStickCoeffs");
    p_.betaTQ = c.lookupScalar("betaTQ");
    p_.delta = c.lookupScalar("delta");
    const vector om = c.lookupVector("omega");
    for (int i = 0; i < 3; ++i) p_.omega[i] = om[i];
    p_.muf = c.lookupScalar("muf");
}
bool constantSlipStick::read(const dictionary& d) {
    stickProperties_ = d;
    const dictionary& c = d.optionalSubDict("constantSlipStickCoeffs");
    readCommon(c);  // read() looks these up in the Coeffs dictionary (:210-212, SURVEY D-8)
    p_.betaTQ = c.lookupScalar("betaTQ");
    p_.delta = c.lookupScalar("delta");
    const vector om = c.lookupVector("omega");
    for (int i = 0; i < 3; ++i) p_.omega[i] = om[i];
    p_.muf = c.lookupScalar("muf");
    return true;
}

exponentialSlipStick::exponentialSlipStick(const word& name, const dictionary& d, fvMeshGpu& mesh)
    : stickModel(name, d, mesh) {
    p_.model = SSAM_STICK_EXPONENTIAL;
    readCommon(d);
    read(d);
    readCommon(d);
}
bool exponentialSlipStick::read(const dictionary& d) {
    stickProperties_ = d;
    const dictionary& c = d.optionalSubDict("exponentialSlipStickCoeffs");
    p_.betaTQ = c.lookupScalar("betaTQ");
    const vector om = c.lookupVector("omega");
    for (int i = 0; i < 3; ++i) p_.omega[i] = om[i];
    p_.omega0 = c.lookupScalar("omega0");
    p_.delta0 = c.lookupScalar("delta0");
    p_.Rs = c.lookupScalar("Rs");
    p_.mu0 = c.lookupScalar("mu0");
    p_.lambda = c.lookupScalar("lambda");
    return true;
}

}  // namespace stickModels

incompressibleFrictionModel::incompressibleFrictionModel(fvMeshGpu& mesh, const dictionary& frictionProperties)
    : dict_(frictionProperties), stickModelPtr_(stickModel::New("qvisc", dict_, mesh)) {}

// ---- boundary conditions -----------------------------------------------------------------------------------------------
constantRotationalSlipFvPatchVectorField::constantRotationalSlipFvPatchVectorField(fvMeshGpu& mesh, int patchi,
                                                                                   const dictionary& dict)
    : mesh_(mesh) {
    p_ = ssam_rotslip_params();
    p_.model = SSAM_ROTSLIP_CONSTANT;
    p_.patch = patchi;
    const vector om = dict.lookupVector("omega"), vt = dict.lookupVector("Vt");
    for (int i = 0; i < 3; ++i) {
        p_.omega[i] = om[i];
        p_.Vt[i] = vt[i];
    }
    p_.delta = dict.lookupScalar("delta");
}
void constantRotationalSlipFvPatchVectorField::updateCoeffs() {
    mesh_.check(ssam_bc_rotational_slip(mesh_.ctx(), &p_), "constantRotationalSlip::updateCoeffs");
}

exponentialRotationalSlipFvPatchVectorField::exponentialRotationalSlipFvPatchVectorField(fvMeshGpu& mesh, int patchi,
                                                                                         const dictionary& dict)
    : mesh_(mesh) {
    p_ = ssam_rotslip_params();
    p_.model = SSAM_ROTSLIP_EXPONENTIAL;
    p_.patch = patchi;
    const vector om = dict.lookupVector("omega"), vt = dict.lookupVector("Vt");
    for (int i = 0; i < 3; ++i) {
        p_.omega[i] = om[i];
        p_.Vt[i] = vt[i];
    }
    p_.omega0 = dict.lookupScalar("omega0");
    p_.delta0 = dict.lookupScalar("delta0");
    p_.R0 = dict.lookupScalar("R0");
}
void exponentialRotationalSlipFvPatchVectorField::updateCoeffs() {
    mesh_.check(ssam_bc_rotational_slip(mesh_.ctx(), &p_), "exponentialRotationalSlip::updateCoeffs");
}

}  // namespace ssam
