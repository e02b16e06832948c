// Synthetic finite-volume mesh generators in OpenFOAM conventions.
//
// Stand-in for blockMesh / snappyHexMesh / decomposePar, none of which exist offline
// (SURVEY.md Appendix E).  Produces exactly the arrays the C-ABI (include/ssam_gpu.h,
// ssam_mesh_desc) takes as *inputs*: owner/neighbour addressing in upper-triangular
// order, patch tables, Sf, Cf, magSf, linear weights, deltaCoeffs, V, C.
//
// Conventions: faces [0,nIF) internal, then boundary faces
// grouped by patch; owner[f] < neighbour[f]; Sf points owner -> neighbour (outward on
// the boundary); physical patches first, processor patches after, ordered by
// neighbour rank; both sides of a processor patch list their faces in the same order.
//
// Geometry formulas restate OpenFOAM v1806 (from memory):
//   internal weights  w = |Sf.(C_N-Cf)| / (|Sf.(Cf-C_P)| + |Sf.(C_N-Cf)|)
//   generic boundary  w = 1, deltaCoeffs = 1/|Cf-C_P|
//   processor patch   w = nbrDist/(ownDist+nbrDist) with face-normal distances
//                     (processorFvPatch::makeWeights), deltaCoeffs = 1/|C_N-C_P|
//
// All cells are axis-aligned boxes in a reference space, optionally graded per axis
// and mapped by an affine matrix A (sheared hex => non-orthogonal mesh), so the exact
// centroids/areas/volumes are available in closed form.

#include <sys/stat.h>

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <cstdint>
#include <cstring>
#include <map>
#include <string>
#include <vector>

namespace {

constexpr double VSMALL = 1e-300;

enum PatchKind : int32_t { GENERIC = 0, PROCESSOR = 1, EMPTY = 2 };

struct Mesh {
    int32_t nCells = 0, nIF = 0, nFaces = 0;
    std::vector<int32_t> owner, neighbour;
    std::vector<int32_t> patchStart, patchSize, patchKind, patchNbrRank;
    std::vector<std::string> patchName;
    std::vector<double> Sf, Cf, magSf, w, deltaCoeffs;  // per face (3*nFaces for vectors)
    std::vector<double> V, C;                           // per cell
    // decomposition bookkeeping (empty for undecomposed meshes)
    std::vector<int32_t> cellGlobal;  // local cell -> global cell
    std::vector<int32_t> faceGlobal;  // local face -> global face (sign in flip)
    std::vector<int32_t> faceFlip;    // 1 if the local face is the flipped global face
    // optional point/face topology (hex blocks; polyMesh reader): what constant/polyMesh/{points,faces} hold
    std::vector<double> points;          // 3*nPoints
    std::vector<int32_t> faceVertStart;  // nFaces+1
    std::vector<int32_t> faceVerts;
};

struct Vec3 {
    double x, y, z;
};
inline Vec3 operator-(Vec3 a, Vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline double dot(Vec3 a, Vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline Vec3 cross(Vec3 a, Vec3 b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline double mag(Vec3 a) { return std::sqrt(dot(a, a)); }

struct Affine {
    double a[9];
    Vec3 apply(Vec3 p) const {
        return {a[0] * p.x + a[1] * p.y + a[2] * p.z, a[3] * p.x + a[4] * p.y + a[5] * p.z,
                a[6] * p.x + a[7] * p.y + a[8] * p.z};
    }
    Vec3 col(int j) const { return {a[j], a[3 + j], a[6 + j]}; }
    double det() const { return dot(col(0), cross(col(1), col(2))); }
};

// Area vector of the reference rectangle normal to axis d (extents du, dv on the two
// cyclically following axes) after the affine map: (A e_u du) x (A e_v dv).
inline Vec3 mappedArea(const Affine& A, int d, double du, double dv) {
    Vec3 eu = A.col((d + 1) % 3), ev = A.col((d + 2) % 3);
    Vec3 s = cross(eu, ev);
    return {s.x * du * dv, s.y * du * dv, s.z * du * dv};
}

inline void put3(std::vector<double>& v, size_t i, Vec3 p) {
    v[3 * i] = p.x;
    v[3 * i + 1] = p.y;
    v[3 * i + 2] = p.z;
}
inline Vec3 get3(const std::vector<double>& v, size_t i) { return {v[3 * i], v[3 * i + 1], v[3 * i + 2]}; }

// weights / deltaCoeffs / magSf from Sf, Cf, C (OpenFOAM surfaceInterpolation::makeWeights,
// makeDeltaCoeffs restated).  Processor patches are filled by the caller.
void finishInternalAndGenericGeometry(Mesh& m) {
    m.magSf.resize(m.nFaces);
    m.w.resize(m.nFaces);
    m.deltaCoeffs.resize(m.nFaces);
    for (int64_t f = 0; f < m.nFaces; ++f) {
        Vec3 S = get3(m.Sf, f), cf = get3(m.Cf, f);
        m.magSf[f] = mag(S);
        Vec3 cp = get3(m.C, m.owner[f]);
        if (f < m.nIF) {
            Vec3 cn = get3(m.C, m.neighbour[f]);
            double SfdOwn = std::fabs(dot(S, cf - cp));
            double SfdNei = std::fabs(dot(S, cn - cf));
            m.w[f] = SfdNei / (SfdOwn + SfdNei);
            m.deltaCoeffs[f] = 1.0 / mag(cn - cp);
        } else {
            m.w[f] = 1.0;
            m.deltaCoeffs[f] = 1.0 / mag(cf - cp);
        }
    }
}

// geometric grading like blockMesh simpleGrading: ratio = last cell size / first.
std::vector<double> gradedCoords(int n, double x0, double len, double ratio) {
    std::vector<double> xs(n + 1);
    if (n == 1 || ratio == 1.0) {
        for (int i = 0; i <= n; ++i) xs[i] = x0 + len * (double(i) / double(n));
        return xs;
    }
    double r = std::pow(ratio, 1.0 / double(n - 1));
    double sum = 0, s = 1;
    std::vector<double> acc(n + 1, 0.0);
    for (int i = 0; i < n; ++i) {
        sum += s;
        acc[i + 1] = sum;
        s *= r;
    }
    for (int i = 0; i <= n; ++i) xs[i] = x0 + len * (acc[i] / sum);
    return xs;
}

const char* kSideName[6] = {"xmin", "xmax", "ymin", "ymax", "zmin", "zmax"};

// ---------------------------------------------------------------------------------
// Structured hex block in blockMesh cell order c = i + nx*(j + ny*k).
// sideNbrRank[s] < 0  => physical (generic) patch, else processor patch to that rank.
// ---------------------------------------------------------------------------------
Mesh* hexBlock(int nx, int ny, int nz, const double* origin, const double* len, const double* grading,
               const double* affine, const int32_t* sideNbrRank, int myRank) {
    Mesh* mp = new Mesh;
    Mesh& m = *mp;
    Affine A;
    std::memcpy(A.a, affine, sizeof(A.a));
    std::vector<double> xs = gradedCoords(nx, origin[0], len[0], grading[0]);
    std::vector<double> ys = gradedCoords(ny, origin[1], len[1], grading[1]);
    std::vector<double> zs = gradedCoords(nz, origin[2], len[2], grading[2]);

    const int64_t nc = int64_t(nx) * ny * nz;
    const int64_t nifx = int64_t(nx - 1) * ny * nz, nify = int64_t(nx) * (ny - 1) * nz,
                  nifz = int64_t(nx) * ny * (nz - 1);
    const int64_t nIF = nifx + nify + nifz;
    const int64_t nb = 2 * (int64_t(ny) * nz + int64_t(nx) * nz + int64_t(nx) * ny);
    m.nCells = int32_t(nc);
    m.nIF = int32_t(nIF);
    m.nFaces = int32_t(nIF + nb);
    m.owner.resize(m.nFaces);
    m.neighbour.resize(m.nIF);
    m.Sf.resize(3 * size_t(m.nFaces));
    m.Cf.resize(3 * size_t(m.nFaces));
    m.V.resize(nc);
    m.C.resize(3 * size_t(nc));
    const double detA = A.det();

    auto cid = [&](int i, int j, int k) { return int32_t(i + int64_t(nx) * (j + int64_t(ny) * k)); };
    // points and face vertex lists (right-hand rule: the face normal points owner -> neighbour / outward)
    auto pid = [&](int i, int j, int k) { return int32_t(i + int64_t(nx + 1) * (j + int64_t(ny + 1) * k)); };
    m.points.resize(3 * size_t(nx + 1) * (ny + 1) * (nz + 1));
    for (int k = 0; k <= nz; ++k)
        for (int j = 0; j <= ny; ++j)
            for (int i = 0; i <= nx; ++i) put3(m.points, pid(i, j, k), A.apply({xs[i], ys[j], zs[k]}));
    m.faceVertStart.assign(1, 0);
    m.faceVerts.reserve(4 * size_t(m.nFaces));
    // quad normal to `axis` at grid index (i,j,k) of its low corner; positive = normal along +axis
    auto quad = [&](int axis, int i, int j, int k, bool positive) {
        int32_t v[4];
        if (axis == 0) {
            v[0] = pid(i, j, k); v[1] = pid(i, j + 1, k); v[2] = pid(i, j + 1, k + 1); v[3] = pid(i, j, k + 1);
        } else if (axis == 1) {
            v[0] = pid(i, j, k); v[1] = pid(i, j, k + 1); v[2] = pid(i + 1, j, k + 1); v[3] = pid(i + 1, j, k);
        } else {
            v[0] = pid(i, j, k); v[1] = pid(i + 1, j, k); v[2] = pid(i + 1, j + 1, k); v[3] = pid(i, j + 1, k);
        }
        if (positive)
            for (int q = 0; q < 4; ++q) m.faceVerts.push_back(v[q]);
        else
            for (int q = 0; q < 4; ++q) m.faceVerts.push_back(v[(4 - q) % 4]);
        m.faceVertStart.push_back(int32_t(m.faceVerts.size()));
    };

    int64_t f = 0;
    for (int k = 0; k < nz; ++k)
        for (int j = 0; j < ny; ++j)
            for (int i = 0; i < nx; ++i) {
                const int32_t c = cid(i, j, k);
                const double dx = xs[i + 1] - xs[i], dy = ys[j + 1] - ys[j], dz = zs[k + 1] - zs[k];
                const double cx = 0.5 * (xs[i] + xs[i + 1]), cy = 0.5 * (ys[j] + ys[j + 1]),
                             cz = 0.5 * (zs[k] + zs[k + 1]);
                put3(m.C, c, A.apply({cx, cy, cz}));
                m.V[c] = detA * (dx * dy * dz);
                if (i + 1 < nx) {
                    m.owner[f] = c;
                    m.neighbour[f] = cid(i + 1, j, k);
                    put3(m.Sf, f, mappedArea(A, 0, dy, dz));
                    put3(m.Cf, f, A.apply({xs[i + 1], cy, cz}));
                    quad(0, i + 1, j, k, true);
                    ++f;
                }
                if (j + 1 < ny) {
                    m.owner[f] = c;
                    m.neighbour[f] = cid(i, j + 1, k);
                    put3(m.Sf, f, mappedArea(A, 1, dz, dx));
                    put3(m.Cf, f, A.apply({cx, ys[j + 1], cz}));
                    quad(1, i, j + 1, k, true);
                    ++f;
                }
                if (k + 1 < nz) {
                    m.owner[f] = c;
                    m.neighbour[f] = cid(i, j, k + 1);
                    put3(m.Sf, f, mappedArea(A, 2, dx, dy));
                    put3(m.Cf, f, A.apply({cx, cy, zs[k + 1]}));
                    quad(2, i, j, k + 1, true);
                    ++f;
                }
            }

    // patch order: the six physical patches in side order on every rank -- decomposePar keeps every
    // physical patch on every processor mesh, with zero faces where the side is a processor boundary,
    // so patch indices (fricPatch!) are the same on all ranks -- then processor sides by neighbour rank
    std::vector<std::pair<int, bool>> order;  // (side, asProcessorPatch)
    for (int s = 0; s < 6; ++s) order.push_back({s, false});
    std::vector<int> procSides;
    for (int s = 0; s < 6; ++s)
        if (sideNbrRank[s] >= 0) procSides.push_back(s);
    std::stable_sort(procSides.begin(), procSides.end(),
                     [&](int a, int b) { return sideNbrRank[a] < sideNbrRank[b]; });
    for (int s : procSides) order.push_back({s, true});

    for (auto& so : order) {
        const int s = so.first;
        const bool isProc = so.second;
        m.patchStart.push_back(int32_t(f));
        m.patchKind.push_back(isProc ? PROCESSOR : GENERIC);
        m.patchNbrRank.push_back(isProc ? sideNbrRank[s] : -1);
        m.patchName.push_back(isProc ? ("procBoundary" + std::to_string(myRank) + "to" +
                                        std::to_string(sideNbrRank[s]))
                                     : std::string(kSideName[s]));
        const int axis = s / 2;
        const bool hi = (s % 2) == 1;
        const double sgn = hi ? 1.0 : -1.0;
        const int64_t f0 = f;
        if (!isProc && sideNbrRank[s] >= 0) {
            m.patchSize.push_back(0);  // physical patch with no faces on this rank
            continue;
        }
        if (axis == 0) {
            const int i = hi ? nx - 1 : 0;
            for (int k = 0; k < nz; ++k)
                for (int j = 0; j < ny; ++j) {
                    const double dy = ys[j + 1] - ys[j], dz = zs[k + 1] - zs[k];
                    Vec3 S = mappedArea(A, 0, dy, dz);
                    m.owner[f] = cid(i, j, k);
                    put3(m.Sf, f, {sgn * S.x, sgn * S.y, sgn * S.z});
                    put3(m.Cf, f,
                         A.apply({hi ? xs[nx] : xs[0], 0.5 * (ys[j] + ys[j + 1]), 0.5 * (zs[k] + zs[k + 1])}));
                    quad(0, hi ? nx : 0, j, k, hi);
                    ++f;
                }
        } else if (axis == 1) {
            const int j = hi ? ny - 1 : 0;
            for (int k = 0; k < nz; ++k)
                for (int i = 0; i < nx; ++i) {
                    const double dx = xs[i + 1] - xs[i], dz = zs[k + 1] - zs[k];
                    Vec3 S = mappedArea(A, 1, dz, dx);
                    m.owner[f] = cid(i, j, k);
                    put3(m.Sf, f, {sgn * S.x, sgn * S.y, sgn * S.z});
                    put3(m.Cf, f,
                         A.apply({0.5 * (xs[i] + xs[i + 1]), hi ? ys[ny] : ys[0], 0.5 * (zs[k] + zs[k + 1])}));
                    quad(1, i, hi ? ny : 0, k, hi);
                    ++f;
                }
        } else {
            const int k = hi ? nz - 1 : 0;
            for (int j = 0; j < ny; ++j)
                for (int i = 0; i < nx; ++i) {
                    const double dx = xs[i + 1] - xs[i], dy = ys[j + 1] - ys[j];
                    Vec3 S = mappedArea(A, 2, dx, dy);
                    m.owner[f] = cid(i, j, k);
                    put3(m.Sf, f, {sgn * S.x, sgn * S.y, sgn * S.z});
                    put3(m.Cf, f,
                         A.apply({0.5 * (xs[i] + xs[i + 1]), 0.5 * (ys[j] + ys[j + 1]), hi ? zs[nz] : zs[0]}));
                    quad(2, i, j, hi ? nz : 0, hi);
                    ++f;
                }
        }
        m.patchSize.push_back(int32_t(f - f0));
    }
    finishInternalAndGenericGeometry(m);

    // Processor sides of a slab partition: the neighbour block continues the same graded
    // grid, so mirror the adjacent cell across the face with the *neighbour's* cell size.
    // Only uniform continuation is supported (nbr cell thickness = own boundary cell
    // thickness scaled by nbrScale == 1), which is what the slab bench generator needs.
    for (size_t p = 0; p < m.patchStart.size(); ++p) {
        if (m.patchKind[p] != PROCESSOR) continue;
        for (int64_t ff = m.patchStart[p]; ff < m.patchStart[p] + m.patchSize[p]; ++ff) {
            Vec3 S = get3(m.Sf, ff), cf = get3(m.Cf, ff), cp = get3(m.C, m.owner[ff]);
            Vec3 cn = {2 * cf.x - cp.x, 2 * cf.y - cp.y, 2 * cf.z - cp.z};
            const double ms = m.magSf[ff];
            Vec3 nf = {S.x / ms, S.y / ms, S.z / ms};
            Vec3 nS = {-S.x, -S.y, -S.z};
            const double nms = mag(nS) + VSMALL;
            Vec3 nnf = {nS.x / nms, nS.y / nms, nS.z / nms};
            const double nbrDist = dot(nnf, cf - cn);
            const double ownDist = dot(nf, cf - cp);
            m.w[ff] = nbrDist / (ownDist + nbrDist);
            m.deltaCoeffs[ff] = 1.0 / mag(cn - cp);
        }
    }
    return mp;
}

// ---------------------------------------------------------------------------------
// "snappy-style" polyhedral mesh: n^3 base hex grid with up to two 2:1 octree
// refinement levels inside cylinders under the tool (axis = z through the top-face
// centre).  Coarse/fine interfaces are split into separate faces, so cells have 6..24
// faces and the coarse/fine faces are non-orthogonal.
// ---------------------------------------------------------------------------------
struct FaceRec {
    int32_t own, nei;  // nei < 0 => boundary, patch = -1-nei
    int8_t dir;        // normal axis
    double pos;        // coordinate along dir
    double u0, u1, v0, v1;
};

Mesh* polyRefined(int nbx, int nby, int nbz, const double* origin, const double* len, double r1, double d1,
                  double r2, double d2, const double* affine) {
    Mesh* mp = new Mesh;
    Mesh& m = *mp;
    Affine A;
    std::memcpy(A.a, affine, sizeof(A.a));
    const double hx = len[0] / nbx, hy = len[1] / nby, hz = len[2] / nbz;
    const double axx = origin[0] + 0.5 * len[0], axy = origin[1] + 0.5 * len[1], ztop = origin[2] + len[2];
    auto bid = [&](int I, int J, int K) { return int64_t(I) + int64_t(nbx) * (J + int64_t(nby) * K); };
    const int64_t nb = int64_t(nbx) * nby * nbz;
    std::vector<int8_t> L(nb, 0);
    for (int K = 0; K < nbz; ++K)
        for (int J = 0; J < nby; ++J)
            for (int I = 0; I < nbx; ++I) {
                const double cx = origin[0] + (I + 0.5) * hx, cy = origin[1] + (J + 0.5) * hy,
                             cz = origin[2] + (K + 0.5) * hz;
                const double rr = std::hypot(cx - axx, cy - axy), depth = ztop - cz;
                int8_t l = 0;
                if (rr <= r1 && depth <= d1) l = 1;
                if (rr <= r2 && depth <= d2) l = 2;
                L[bid(I, J, K)] = l;
            }
    // 2:1 balance across faces: neighbours of a level-2 base cell must be >= level 1
    for (int K = 0; K < nbz; ++K)
        for (int J = 0; J < nby; ++J)
            for (int I = 0; I < nbx; ++I) {
                if (L[bid(I, J, K)] != 2) continue;
                const int dI[6] = {-1, 1, 0, 0, 0, 0}, dJ[6] = {0, 0, -1, 1, 0, 0}, dK[6] = {0, 0, 0, 0, -1, 1};
                for (int s = 0; s < 6; ++s) {
                    int I2 = I + dI[s], J2 = J + dJ[s], K2 = K + dK[s];
                    if (I2 < 0 || J2 < 0 || K2 < 0 || I2 >= nbx || J2 >= nby || K2 >= nbz) continue;
                    if (L[bid(I2, J2, K2)] == 0) L[bid(I2, J2, K2)] = 1;
                }
            }
    std::vector<int64_t> off(nb + 1, 0);
    for (int64_t b = 0; b < nb; ++b) off[b + 1] = off[b] + (int64_t(1) << (3 * L[b]));
    const int64_t nc = off[nb];
    m.nCells = int32_t(nc);
    m.V.resize(nc);
    m.C.resize(3 * size_t(nc));
    const double detA = A.det();

    std::vector<FaceRec> internal;
    std::vector<std::vector<FaceRec>> bnd(6);
    internal.reserve(size_t(nc) * 3 + 1024);
    std::vector<FaceRec> mine;

    for (int K = 0; K < nbz; ++K)
        for (int J = 0; J < nby; ++J)
            for (int I = 0; I < nbx; ++I) {
                const int64_t b = bid(I, J, K);
                const int l = L[b], n = 1 << l;
                const double sx = hx / n, sy = hy / n, sz = hz / n;
                const double bx = origin[0] + I * hx, by = origin[1] + J * hy, bz = origin[2] + K * hz;
                for (int k = 0; k < n; ++k)
                    for (int j = 0; j < n; ++j)
                        for (int i = 0; i < n; ++i) {
                            const int32_t c = int32_t(off[b] + i + n * (j + n * k));
                            const double x0 = bx + i * sx, y0 = by + j * sy, z0 = bz + k * sz;
                            const double x1 = bx + (i + 1) * sx, y1 = by + (j + 1) * sy, z1 = bz + (k + 1) * sz;
                            put3(m.C, c, A.apply({0.5 * (x0 + x1), 0.5 * (y0 + y1), 0.5 * (z0 + z1)}));
                            m.V[c] = detA * ((x1 - x0) * (y1 - y0) * (z1 - z0));
                            mine.clear();
                            // loc[3] = local index, for axis d the in-plane axes are u=(d+1)%3, v=(d+2)%3
                            const int loc[3] = {i, j, k};
                            const int Bidx[3] = {I, J, K};
                            const int nB[3] = {nbx, nby, nbz};
                            const double lo[3] = {x0, y0, z0}, hi[3] = {x1, y1, z1};
                            for (int d = 0; d < 3; ++d) {
                                const int u = (d + 1) % 3, v = (d + 2) % 3;
                                // minus-side boundary
                                if (loc[d] == 0 && Bidx[d] == 0)
                                    bnd[2 * d].push_back({c, -1, int8_t(d), lo[d], lo[u], hi[u], lo[v], hi[v]});
                                if (loc[d] + 1 < n) {
                                    int l2[3] = {i, j, k};
                                    l2[d] += 1;
                                    const int32_t c2 = int32_t(off[b] + l2[0] + n * (l2[1] + n * l2[2]));
                                    mine.push_back({c, c2, int8_t(d), hi[d], lo[u], hi[u], lo[v], hi[v]});
                                } else if (Bidx[d] + 1 >= nB[d]) {
                                    bnd[2 * d + 1].push_back({c, -1, int8_t(d), hi[d], lo[u], hi[u], lo[v], hi[v]});
                                } else {
                                    int B2[3] = {I, J, K};
                                    B2[d] += 1;
                                    const int64_t b2 = bid(B2[0], B2[1], B2[2]);
                                    const int lb = L[b2], n2 = 1 << lb;
                                    if (lb == l) {
                                        int l2[3] = {i, j, k};
                                        l2[d] = 0;
                                        const int32_t c2 = int32_t(off[b2] + l2[0] + n2 * (l2[1] + n2 * l2[2]));
                                        mine.push_back({c, c2, int8_t(d), hi[d], lo[u], hi[u], lo[v], hi[v]});
                                    } else if (lb == l - 1) {
                                        int l2[3] = {i >> 1, j >> 1, k >> 1};
                                        l2[d] = 0;
                                        const int32_t c2 = int32_t(off[b2] + l2[0] + n2 * (l2[1] + n2 * l2[2]));
                                        mine.push_back({c, c2, int8_t(d), hi[d], lo[u], hi[u], lo[v], hi[v]});
                                    } else {  // lb == l + 1: four fine faces
                                        const double mu = 0.5 * (lo[u] + hi[u]), mv = 0.5 * (lo[v] + hi[v]);
                                        for (int b_v = 0; b_v < 2; ++b_v)
                                            for (int b_u = 0; b_u < 2; ++b_u) {
                                                int l2[3] = {2 * i, 2 * j, 2 * k};
                                                l2[d] = 0;
                                                l2[u] += b_u;
                                                l2[v] += b_v;
                                                const int32_t c2 =
                                                    int32_t(off[b2] + l2[0] + n2 * (l2[1] + n2 * l2[2]));
                                                mine.push_back({c, c2, int8_t(d), hi[d], b_u ? mu : lo[u],
                                                                b_u ? hi[u] : mu, b_v ? mv : lo[v], b_v ? hi[v] : mv});
                                            }
                                    }
                                }
                            }
                            std::sort(mine.begin(), mine.end(),
                                      [](const FaceRec& a, const FaceRec& b2) { return a.nei < b2.nei; });
                            for (auto& r : mine) internal.push_back(r);
                        }
            }
    m.nIF = int32_t(internal.size());
    size_t nbf = 0;
    for (auto& v : bnd) nbf += v.size();
    m.nFaces = int32_t(internal.size() + nbf);
    m.owner.resize(m.nFaces);
    m.neighbour.resize(m.nIF);
    m.Sf.resize(3 * size_t(m.nFaces));
    m.Cf.resize(3 * size_t(m.nFaces));
    auto emit = [&](int64_t f, const FaceRec& r, double sgn) {
        const int d = r.dir, u = (d + 1) % 3, v = (d + 2) % 3;
        Vec3 S = mappedArea(A, d, r.u1 - r.u0, r.v1 - r.v0);
        put3(m.Sf, f, {sgn * S.x, sgn * S.y, sgn * S.z});
        double p[3];
        p[d] = r.pos;
        p[u] = 0.5 * (r.u0 + r.u1);
        p[v] = 0.5 * (r.v0 + r.v1);
        put3(m.Cf, f, A.apply({p[0], p[1], p[2]}));
        m.owner[f] = r.own;
    };
    int64_t f = 0;
    for (auto& r : internal) {
        emit(f, r, 1.0);
        m.neighbour[f] = r.nei;
        ++f;
    }
    for (int s = 0; s < 6; ++s) {
        m.patchStart.push_back(int32_t(f));
        m.patchSize.push_back(int32_t(bnd[s].size()));
        m.patchKind.push_back(GENERIC);
        m.patchNbrRank.push_back(-1);
        m.patchName.push_back(kSideName[s]);
        for (auto& r : bnd[s]) {
            emit(f, r, (s % 2) ? 1.0 : -1.0);
            ++f;
        }
    }
    finishInternalAndGenericGeometry(m);
    return mp;
}

// ---------------------------------------------------------------------------------
// decomposePar-style decomposition for an arbitrary cellProc list.
// ---------------------------------------------------------------------------------
void decompose(const Mesh& g, const int32_t* cellProc, int nProcs, Mesh** out) {
    std::vector<std::vector<int32_t>> cells(nProcs);
    std::vector<int32_t> localCell(g.nCells);
    for (int32_t c = 0; c < g.nCells; ++c) {
        const int p = cellProc[c];
        localCell[c] = int32_t(cells[p].size());
        cells[p].push_back(c);
    }
    // global patch of each boundary face
    std::vector<int32_t> facePatch(g.nFaces - g.nIF);
    size_t nPhysical = 0;
    for (size_t p = 0; p < g.patchStart.size(); ++p) {
        for (int32_t f = g.patchStart[p]; f < g.patchStart[p] + g.patchSize[p]; ++f) facePatch[f - g.nIF] = int32_t(p);
        nPhysical = p + 1;
    }
    for (int p = 0; p < nProcs; ++p) {
        Mesh* mp = new Mesh;
        Mesh& m = *mp;
        m.nCells = int32_t(cells[p].size());
        m.cellGlobal = cells[p];
        m.V.resize(m.nCells);
        m.C.resize(3 * size_t(m.nCells));
        for (int32_t lc = 0; lc < m.nCells; ++lc) {
            m.V[lc] = g.V[cells[p][lc]];
            put3(m.C, lc, get3(g.C, cells[p][lc]));
        }
        std::vector<int32_t> internalF;
        std::vector<std::vector<int32_t>> physF(nPhysical);
        std::map<int, std::vector<std::pair<int32_t, int32_t>>> procF;  // nbr -> (global face, flip)
        for (int32_t f = 0; f < g.nIF; ++f) {
            const int po = cellProc[g.owner[f]], pn = cellProc[g.neighbour[f]];
            if (po == p && pn == p)
                internalF.push_back(f);
            else if (po == p)
                procF[pn].push_back({f, 0});
            else if (pn == p)
                procF[po].push_back({f, 1});
        }
        for (int32_t f = g.nIF; f < g.nFaces; ++f)
            if (cellProc[g.owner[f]] == p) physF[facePatch[f - g.nIF]].push_back(f);

        m.nIF = int32_t(internalF.size());
        size_t nf = internalF.size();
        for (auto& v : physF) nf += v.size();
        for (auto& kv : procF) nf += kv.second.size();
        m.nFaces = int32_t(nf);
        m.owner.resize(nf);
        m.neighbour.resize(m.nIF);
        m.Sf.resize(3 * nf);
        m.Cf.resize(3 * nf);
        m.magSf.resize(nf);
        m.w.resize(nf);
        m.deltaCoeffs.resize(nf);
        m.faceGlobal.resize(nf);
        m.faceFlip.assign(nf, 0);
        int64_t lf = 0;
        for (int32_t f : internalF) {
            m.owner[lf] = localCell[g.owner[f]];
            m.neighbour[lf] = localCell[g.neighbour[f]];
            put3(m.Sf, lf, get3(g.Sf, f));
            put3(m.Cf, lf, get3(g.Cf, f));
            m.magSf[lf] = g.magSf[f];
            m.w[lf] = g.w[f];
            m.deltaCoeffs[lf] = g.deltaCoeffs[f];
            m.faceGlobal[lf] = f;
            ++lf;
        }
        for (size_t gp = 0; gp < nPhysical; ++gp) {
            m.patchStart.push_back(int32_t(lf));
            m.patchSize.push_back(int32_t(physF[gp].size()));
            m.patchKind.push_back(g.patchKind[gp]);
            m.patchNbrRank.push_back(-1);
            m.patchName.push_back(g.patchName[gp]);
            for (int32_t f : physF[gp]) {
                m.owner[lf] = localCell[g.owner[f]];
                put3(m.Sf, lf, get3(g.Sf, f));
                put3(m.Cf, lf, get3(g.Cf, f));
                m.magSf[lf] = g.magSf[f];
                m.w[lf] = g.w[f];
                m.deltaCoeffs[lf] = g.deltaCoeffs[f];
                m.faceGlobal[lf] = f;
                ++lf;
            }
        }
        for (auto& kv : procF) {
            m.patchStart.push_back(int32_t(lf));
            m.patchSize.push_back(int32_t(kv.second.size()));
            m.patchKind.push_back(PROCESSOR);
            m.patchNbrRank.push_back(kv.first);
            m.patchName.push_back("procBoundary" + std::to_string(p) + "to" + std::to_string(kv.first));
            for (auto& fr : kv.second) {
                const int32_t f = fr.first;
                const bool flip = fr.second != 0;
                const int32_t gOwn = flip ? g.neighbour[f] : g.owner[f];
                const int32_t gNbr = flip ? g.owner[f] : g.neighbour[f];
                Vec3 S = get3(g.Sf, f);
                if (flip) S = {-S.x, -S.y, -S.z};
                Vec3 cf = get3(g.Cf, f), cp = get3(g.C, gOwn), cn = get3(g.C, gNbr);
                m.owner[lf] = localCell[gOwn];
                put3(m.Sf, lf, S);
                put3(m.Cf, lf, cf);
                const double ms = mag(S);
                m.magSf[lf] = ms;
                Vec3 nfv = {S.x / ms, S.y / ms, S.z / ms};
                Vec3 nS = {-S.x, -S.y, -S.z};
                const double nms = mag(nS) + VSMALL;
                Vec3 nnf = {nS.x / nms, nS.y / nms, nS.z / nms};
                const double nbrDist = dot(nnf, cf - cn);
                const double ownDist = dot(nfv, cf - cp);
                m.w[lf] = nbrDist / (ownDist + nbrDist);
                m.deltaCoeffs[lf] = 1.0 / mag(cn - cp);
                m.faceGlobal[lf] = f;
                m.faceFlip[lf] = flip ? 1 : 0;
                ++lf;
            }
        }
        out[p] = mp;
    }
}

#include "foam_io.inc"

}  // namespace

extern "C" {

typedef struct smg_mesh smg_mesh;

smg_mesh* smg_hex_block(int nx, int ny, int nz, const double* origin, const double* len, const double* grading,
                        const double* affine, const int32_t* sideNbrRank, int myRank) {
    return reinterpret_cast<smg_mesh*>(hexBlock(nx, ny, nz, origin, len, grading, affine, sideNbrRank, myRank));
}

smg_mesh* smg_poly_refined(int nbx, int nby, int nbz, const double* origin, const double* len, double r1, double d1,
                           double r2, double d2, const double* affine) {
    return reinterpret_cast<smg_mesh*>(polyRefined(nbx, nby, nbz, origin, len, r1, d1, r2, d2, affine));
}

int smg_decompose(const smg_mesh* g, const int32_t* cellProc, int nProcs, smg_mesh** out) {
    decompose(*reinterpret_cast<const Mesh*>(g), cellProc, nProcs, reinterpret_cast<Mesh**>(out));
    return 0;
}

void smg_free(smg_mesh* m) { delete reinterpret_cast<Mesh*>(m); }

void smg_sizes(const smg_mesh* mm, int32_t* out4) {
    const Mesh& m = *reinterpret_cast<const Mesh*>(mm);
    out4[0] = m.nCells;
    out4[1] = m.nIF;
    out4[2] = m.nFaces;
    out4[3] = int32_t(m.patchStart.size());
}

#define SMG_GETTER(name, type, member)                                  \
    const type* smg_##name(const smg_mesh* mm) {                        \
        return reinterpret_cast<const Mesh*>(mm)->member.data();        \
    }
SMG_GETTER(owner, int32_t, owner)
SMG_GETTER(neighbour, int32_t, neighbour)
SMG_GETTER(patch_start, int32_t, patchStart)
SMG_GETTER(patch_size, int32_t, patchSize)
SMG_GETTER(patch_kind, int32_t, patchKind)
SMG_GETTER(patch_nbr_rank, int32_t, patchNbrRank)
SMG_GETTER(Sf, double, Sf)
SMG_GETTER(Cf, double, Cf)
SMG_GETTER(magSf, double, magSf)
SMG_GETTER(weights, double, w)
SMG_GETTER(delta_coeffs, double, deltaCoeffs)
SMG_GETTER(V, double, V)
SMG_GETTER(C, double, C)
SMG_GETTER(cell_global, int32_t, cellGlobal)
SMG_GETTER(face_global, int32_t, faceGlobal)
SMG_GETTER(face_flip, int32_t, faceFlip)

SMG_GETTER(points, double, points)
SMG_GETTER(face_vert_start, int32_t, faceVertStart)
SMG_GETTER(face_verts, int32_t, faceVerts)
int64_t smg_n_points(const smg_mesh* mm) { return int64_t(reinterpret_cast<const Mesh*>(mm)->points.size() / 3); }

// ---- OpenFOAM case I/O (foam_io.inc) ----
int smg_write_polymesh(const smg_mesh* mm, const char* polyMeshDir, int binary) {
    try {
        writePolyMesh(*reinterpret_cast<const Mesh*>(mm), polyMeshDir, binary != 0);
        return 0;
    } catch (const std::exception& e) {
        g_ioError = e.what();
        return 1;
    }
}
smg_mesh* smg_read_polymesh(const char* polyMeshDir) {
    try {
        return reinterpret_cast<smg_mesh*>(readPolyMesh(polyMeshDir));
    } catch (const std::exception& e) {
        g_ioError = e.what();
        return nullptr;
    }
}
int smg_couple_processor_patches(smg_mesh** meshes, int n) {
    try {
        coupleProcessorPatches(reinterpret_cast<Mesh**>(meshes), n);
        return 0;
    } catch (const std::exception& e) {
        g_ioError = e.what();
        return 1;
    }
}
int smg_write_field(const smg_mesh* mm, const char* path, const char* name, int ncomp, const double* internal,
                    const double* boundary, const int32_t* patchBc, int binary) {
    try {
        writeVolField(*reinterpret_cast<const Mesh*>(mm), path, name, ncomp, internal, boundary, patchBc, binary != 0);
        return 0;
    } catch (const std::exception& e) {
        g_ioError = e.what();
        return 1;
    }
}
int smg_read_field(const smg_mesh* mm, const char* path, int ncomp, double* internal, double* boundary,
                   int32_t* patchBc) {
    try {
        readVolField(*reinterpret_cast<const Mesh*>(mm), path, ncomp, internal, boundary, patchBc);
        return 0;
    } catch (const std::exception& e) {
        g_ioError = e.what();
        return 1;
    }
}
const char* smg_io_error() { return g_ioError.c_str(); }

const char* smg_patch_name(const smg_mesh* mm, int p) {
    return reinterpret_cast<const Mesh*>(mm)->patchName[p].c_str();
}
int32_t smg_has_global_maps(const smg_mesh* mm) {
    return reinterpret_cast<const Mesh*>(mm)->cellGlobal.empty() ? 0 : 1;
}

}  // extern "C"
