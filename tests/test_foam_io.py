"""OpenFOAM case ingest: constant/polyMesh (ascii + binary), processorN directories,
volScalarField / volVectorField files.  OpenFOAM is not available offline, so these are round trips of the
synthetic meshes through OpenFOAM's file formats: the geometry the reader computes from points and faces
(OpenFOAM's face / cell centre and volume algorithms) must equal the generator's closed-form geometry."""
import os

import numpy as np
import pytest

from solidstateamfoam_b200 import cases
from solidstateamfoam_b200 import mesh as M
from solidstateamfoam_b200 import params as P

SHEAR = (1.0, 0.3, 0.1, 0.0, 1.0, 0.2, 0.0, 0.0, 1.0)


def same_mesh(a: M.Mesh, b: M.Mesh, rtol=1e-13):
    for k in ("nCells", "nInternalFaces", "nFaces"):
        assert getattr(a, k) == getattr(b, k), k
    assert a.patchName == b.patchName
    for k in ("owner", "neighbour", "patchStart", "patchSize", "patchKind", "patchNbrRank"):
        assert np.array_equal(getattr(a, k), getattr(b, k)), k
    scale = {"Sf": np.abs(b.Sf).max(), "Cf": np.abs(b.Cf).max() + 1e-300, "C": np.abs(b.C).max() + 1e-300}
    for k in ("Sf", "Cf", "C"):
        assert np.abs(getattr(a, k) - getattr(b, k)).max() <= rtol * scale[k], k
    for k in ("magSf", "V", "weights", "deltaCoeffs"):
        assert np.allclose(getattr(a, k), getattr(b, k), rtol=rtol * 10, atol=0), k


@pytest.mark.parametrize("binary", [False, True])
@pytest.mark.parametrize("affine,grading", [(M.IDENTITY, (1, 1, 1)), (SHEAR, (2.0, 0.5, 3.0))])
def test_polymesh_round_trip(tmp_path, binary, affine, grading):
    h = M.hex_block_handle(7, 5, 4, length=(1.4, 1.0, 0.6), affine=affine, grading=grading)
    d = str(tmp_path / "constant" / "polyMesh")
    h.write_polymesh(d, binary=binary)
    for name in ("points", "faces", "owner", "neighbour", "boundary"):
        assert os.path.exists(os.path.join(d, name))
    txt = open(os.path.join(d, "boundary")).read()
    assert "FoamFile" in txt and "polyBoundaryMesh" in txt and "startFace" in txt
    r = M.read_polymesh(d)
    same_mesh(r.mesh(), h.mesh())


def test_decomposed_case_round_trip(tmp_path):
    R = 3
    hs = []
    for r in range(R):
        side = [-1] * 6
        if r > 0:
            side[4] = r - 1
        if r < R - 1:
            side[5] = r + 1
        h = M.hex_block_handle(5, 4, 3, origin=(0, 0, 0.3 * r), length=(1.0, 0.8, 0.3), affine=SHEAR,
                               side_nbr_rank=tuple(side), my_rank=r)
        h.write_polymesh(str(tmp_path / f"processor{r}" / "constant" / "polyMesh"))
        hs.append(h)
    txt = open(tmp_path / "processor1" / "constant" / "polyMesh" / "boundary").read()
    assert "procBoundary1to0" in txt and "neighbProcNo    2;" in txt
    rs = M.read_decomposed_case(str(tmp_path), R)
    for r in range(R):
        # incl. the processor-patch weights / deltaCoeffs recomputed from the neighbour side
        same_mesh(rs[r].mesh(), hs[r].mesh(), rtol=1e-12)


@pytest.mark.parametrize("binary", [False, True])
def test_field_round_trip_and_oracle_on_ingested_case(tmp_path, binary):
    """write a cfg1 case to disk in OpenFOAM format, read everything back, run the oracle on both"""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import oracle_update
    from oracle import pyoracle as O

    case = cases.cfg1(dims=(8, 6, 5), affine=SHEAR)
    dims = case.mesh.meta["dims"]
    h = M.hex_block_handle(*dims, length=case.mesh.meta["length"], affine=SHEAR)
    pm = str(tmp_path / "constant" / "polyMesh")
    h.write_polymesh(pm, binary=binary)
    t0 = str(tmp_path / "0")
    h.write_field(os.path.join(t0, "U"), "U", case.U_i, case.U_b, case.patch_u_bc, binary)
    h.write_field(os.path.join(t0, "temp"), "temp", case.T_i, case.T_b, None, binary)
    h.write_field(os.path.join(t0, "alpha.metal"), "alpha.metal", case.alpha1_i, case.alpha1_b, None, binary)
    h.write_field(os.path.join(t0, "p"), "p", case.p_i, case.p_b, None, binary)
    r = M.read_polymesh(pm)
    U_i, U_b, bc = r.read_field(os.path.join(t0, "U"), 3)
    T_i, T_b, _ = r.read_field(os.path.join(t0, "temp"), 1)
    a_i, a_b, _ = r.read_field(os.path.join(t0, "alpha.metal"), 1)
    p_i, p_b, _ = r.read_field(os.path.join(t0, "p"), 1)
    assert np.array_equal(bc, case.patch_u_bc)
    for got, ref in ((U_i, case.U_i), (U_b, case.U_b), (T_i, case.T_i), (T_b, case.T_b), (a_i, case.alpha1_i),
                     (p_b, case.p_b)):
        assert np.array_equal(got, ref)  # 17 significant digits / raw bytes: exact
    # the ingested case drives the same update
    import copy
    ing = copy.copy(case)
    ing.mesh = r.mesh()
    ing.U_i, ing.U_b, ing.T_i, ing.T_b, ing.alpha1_i, ing.alpha1_b, ing.p_i, ing.p_b = U_i, U_b, T_i, T_b, a_i, a_b, p_i, p_b
    a = oracle_update(case)
    b = oracle_update(ing)
    for f in case.outputs:
        x, y = a.field(f), b.field(f)
        assert np.allclose(x, y, rtol=1e-9, atol=0), P.FIELD_NAMES[f]  # geometry agrees to ~1e-13


def test_reader_errors(tmp_path):
    with pytest.raises(M.FoamIOError) as e:
        M.read_polymesh(str(tmp_path / "nowhere"))
    assert "cannot open file" in str(e.value)
    h = M.hex_block_handle(2, 2, 2)
    d = str(tmp_path / "constant" / "polyMesh")
    h.write_polymesh(d)
    p = os.path.join(d, "owner")
    open(p, "w").write(open(p).read().replace("label", "label", 1).rsplit("\n", 4)[0])  # truncate
    with pytest.raises(M.FoamIOError):
        M.read_polymesh(d)
    with pytest.raises(M.FoamIOError) as e:
        M.MeshHandle(M.poly_refined(4, 4, 4, _keep=True)).write_polymesh(str(tmp_path / "x"))
    assert "no point/face topology" in str(e.value)


# ---- hand-written OpenFOAM-v1806-style files (tests/golden/foam_v1806_2cell) ----------------------------------------
# Two unit cubes side by side, written the way v1806's own writers lay the files out (banner, `note` entry, `inGroups`,
# one-line short lists `1(1)`, `0()`, an empty physical patch kept by decomposePar, `nonuniform 0()`), NOT by this
# repo's writer; the expected numbers below are worked out by hand.
FIX = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "foam_v1806_2cell")


def test_reads_hand_written_v1806_polymesh():
    m = M.read_polymesh(os.path.join(FIX, "constant", "polyMesh")).mesh()
    assert (m.nCells, m.nInternalFaces, m.nFaces) == (2, 1, 11)
    assert m.patchName == ["left", "right", "walls"]
    assert list(m.patchStart) == [1, 2, 3] and list(m.patchSize) == [1, 1, 8] and list(m.patchKind) == [0, 0, 0]
    assert list(m.owner) == [0, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1] and list(m.neighbour) == [1]
    assert np.allclose(m.V, [1.0, 1.0], rtol=1e-15)
    assert np.allclose(m.C, [[0.5, 0.5, 0.5], [1.5, 0.5, 0.5]], rtol=1e-15)
    expect_sf = np.array([[1, 0, 0], [-1, 0, 0], [1, 0, 0], [0, -1, 0], [0, -1, 0], [0, 1, 0], [0, 1, 0], [0, 0, -1],
                          [0, 0, -1], [0, 0, 1], [0, 0, 1]], dtype=float)
    assert np.allclose(m.Sf, expect_sf, atol=1e-15)
    assert np.allclose(m.Cf[0], [1.0, 0.5, 0.5]) and np.allclose(m.Cf[2], [2.0, 0.5, 0.5])
    assert m.weights[0] == pytest.approx(0.5, rel=1e-15) and np.all(m.weights[1:] == 1.0)
    assert m.deltaCoeffs[0] == pytest.approx(1.0, rel=1e-15)          # 1/|C_N - C_P|
    assert np.allclose(m.deltaCoeffs[1:], 2.0, rtol=1e-15)              # 1/|Cf - C_P| on the boundary


def test_reads_hand_written_v1806_fields():
    h = M.read_polymesh(os.path.join(FIX, "constant", "polyMesh"))
    ti, tb, bc = h.read_field(os.path.join(FIX, "0", "temp"), 1)
    assert list(ti) == [300.0, 400.0]
    assert tb[0] == 293.0                      # fixedValue
    assert tb[1] == 400.0                      # zeroGradient on `right`: the adjacent cell's value
    assert list(tb[2:]) == [300.0, 400.0] * 4  # zeroGradient walls, faces alternate between the two cells
    assert list(bc) == [P.BC_VALUE, P.BC_ZERO_GRADIENT, P.BC_ZERO_GRADIENT]
    ui, ub, ubc = h.read_field(os.path.join(FIX, "0", "U"), 3)
    assert np.array_equal(ui, [[1, 0, 0], [2, 0.5, 0]])
    assert np.array_equal(ub[0], [1, 0, 0]) and np.array_equal(ub[1], [2, 0.5, 0])
    assert np.array_equal(ub[2:], np.zeros((8, 3)))  # noSlip = fixedValue (0 0 0)
    assert list(ubc) == [P.BC_VALUE, P.BC_ZERO_GRADIENT, P.BC_VALUE]


def test_reads_hand_written_v1806_decomposed_case():
    hs = M.read_decomposed_case(FIX, 2)
    for r, h in enumerate(hs):
        m = h.mesh()
        assert (m.nCells, m.nInternalFaces, m.nFaces) == (1, 0, 6)
        assert m.patchName == ["left", "right", "walls", f"procBoundary{r}to{1 - r}"]
        assert list(m.patchKind) == [0, 0, 0, 1] and m.patchNbrRank[3] == 1 - r
        assert list(m.patchSize) == ([1, 0, 4, 1] if r == 0 else [0, 1, 4, 1])
        f = 5  # the processor face
        assert np.allclose(m.Sf[f], [1.0 if r == 0 else -1.0, 0, 0], atol=1e-15)
        # completed from the neighbour side, like processorFvPatch::makeWeights / makeDeltaCoeffs
        assert m.weights[f] == pytest.approx(0.5, rel=1e-15)
        assert m.deltaCoeffs[f] == pytest.approx(1.0, rel=1e-15)
        ti, tb, bc = h.read_field(os.path.join(FIX, f"processor{r}", "0", "temp"), 1)
        assert list(ti) == [(300.0, 400.0)[r]]
        assert tb[-1] == (400.0, 300.0)[r]     # processor patch: the neighbour cell's value as written


# ---- the same two cubes in v1806's BINARY layout, emitted by an independent writer (this test, numpy.tobytes) ----------
# binary lists are `N` newline `(` raw little-endian bytes `)` (List<T>::writeList + OSstream::write(const char*, n));
# faces become a faceCompactList = offsets labelList + flat labels; the header carries arch "LSB;label=32;scalar=64";
# constant/polyMesh/boundary stays ascii in every format.
BANNER = (
    "/*--------------------------------*- C++ -*----------------------------------*\\\n"
    "| =========                 |                                                 |\n"
    "| \\\\      /  F ield         | OpenFOAM: The Open Source CFD Toolbox           |\n"
    "|  \\\\    /   O peration     | Version:  v1806                                 |\n"
    "|   \\\\  /    A nd           | Web:      www.OpenFOAM.com                      |\n"
    "|    \\\\/     M anipulation  |                                                 |\n"
    "\\*---------------------------------------------------------------------------*/\n")
RULE = "// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //\n"
TAIL = b"\n\n// ************************************************************************* //\n"


def _foam_header(cls, obj, location='"constant/polyMesh"', note=None):
    s = BANNER + "FoamFile\n{\n    version     2.0;\n    format      binary;\n    class       " + cls + ";\n"
    s += '    arch        "LSB;label=32;scalar=64";\n'
    if note:
        s += f'    note        "{note}";\n'
    s += f"    location    {location};\n    object      {obj};\n}}\n" + RULE + "\n\n"
    return s.encode()


def _binary_list(a):
    a = np.ascontiguousarray(a)
    return f"{a.shape[0]}\n(".encode() + a.tobytes() + b")\n"


def test_reads_binary_polymesh_from_an_independent_writer(tmp_path):
    pts = np.array([[i, j, k] for k in (0, 1) for j in (0, 1) for i in (0, 1, 2)], dtype="<f8")
    faces = [[1, 4, 10, 7], [0, 6, 9, 3], [2, 5, 11, 8], [0, 1, 7, 6], [1, 2, 8, 7], [3, 9, 10, 4], [4, 10, 11, 5],
             [0, 3, 4, 1], [1, 4, 5, 2], [6, 7, 10, 9], [7, 8, 11, 10]]
    owner = np.array([0, 0, 1, 0, 1, 0, 1, 0, 1, 0, 1], dtype="<i4")
    note = "nPoints:12  nCells:2  nFaces:11  nInternalFaces:1"
    pm = tmp_path / "constant" / "polyMesh"
    pm.mkdir(parents=True)
    (pm / "points").write_bytes(_foam_header("vectorField", "points") + _binary_list(pts) + TAIL)
    offsets = np.cumsum([0] + [len(f) for f in faces]).astype("<i4")
    flat = np.array([v for f in faces for v in f], dtype="<i4")
    (pm / "faces").write_bytes(_foam_header("faceCompactList", "faces") + _binary_list(offsets) + b"\n" +
                               _binary_list(flat) + TAIL)
    (pm / "owner").write_bytes(_foam_header("labelList", "owner", note=note) + _binary_list(owner) + TAIL)
    (pm / "neighbour").write_bytes(_foam_header("labelList", "neighbour", note=note) +
                                   _binary_list(np.array([1], dtype="<i4")) + TAIL)
    with open(os.path.join(FIX, "constant", "polyMesh", "boundary"), "rb") as fh:
        (pm / "boundary").write_bytes(fh.read())  # ascii in every format

    h = M.read_polymesh(str(pm))
    m = h.mesh()
    ref = M.read_polymesh(os.path.join(FIX, "constant", "polyMesh")).mesh()
    same_mesh(m, ref, rtol=0.0)  # same numbers as the ascii files, bit for bit

    # a binary volScalarField / volVectorField: `nonuniform List<scalar>` followed by the binary list
    t0 = tmp_path / "0"
    t0.mkdir()
    body = (b"dimensions      [0 0 0 1 0 0 0];\n\ninternalField   nonuniform List<scalar> \n" +
            _binary_list(np.array([300.0, 400.0], dtype="<f8")) + b";\n\nboundaryField\n{\n"
            b"    left\n    {\n        type            fixedValue;\n        value           uniform 293;\n    }\n"
            b"    right\n    {\n        type            zeroGradient;\n    }\n"
            b"    walls\n    {\n        type            fixedValue;\n        value           nonuniform List<scalar> \n" +
            _binary_list(np.arange(8, dtype="<f8") + 500.0) + b";\n    }\n}\n")
    (t0 / "temp").write_bytes(_foam_header("volScalarField", "temp", location='"0"') + body + TAIL)
    ti, tb, bc = h.read_field(str(t0 / "temp"), 1)
    assert list(ti) == [300.0, 400.0]
    assert tb[0] == 293.0 and tb[1] == 400.0 and list(tb[2:]) == [500.0 + i for i in range(8)]
    assert list(bc) == [P.BC_VALUE, P.BC_ZERO_GRADIENT, P.BC_VALUE]
    ubody = (b"dimensions      [0 1 -1 0 0 0 0];\n\ninternalField   nonuniform List<vector> \n" +
             _binary_list(np.array([[1, 0, 0], [2, 0.5, 0]], dtype="<f8")) + b";\n\nboundaryField\n{\n"
             b"    left\n    {\n        type            fixedValue;\n        value           uniform (1 0 0);\n    }\n"
             b"    right\n    {\n        type            zeroGradient;\n    }\n"
             b"    walls\n    {\n        type            noSlip;\n    }\n}\n")
    (t0 / "U").write_bytes(_foam_header("volVectorField", "U", location='"0"') + ubody + TAIL)
    ui, ub, ubc = h.read_field(str(t0 / "U"), 3)
    assert np.array_equal(ui, [[1, 0, 0], [2, 0.5, 0]])
    assert np.array_equal(ub[0], [1, 0, 0]) and np.array_equal(ub[1], [2, 0.5, 0]) and not ub[2:].any()
    assert list(ubc) == [P.BC_VALUE, P.BC_ZERO_GRADIENT, P.BC_VALUE]
