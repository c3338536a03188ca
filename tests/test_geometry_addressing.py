"""lduAddressing and the geometric fields the kernels read, WITHOUT handing
the solver the oracle's geometry first (tests/parity.py does that for the stage tests): owner /
neighbour and the cell -> face ELL table against an independent numpy construction from the polyMesh
(integers: exact), V, Sf, |Sf|, linear weights, nonOrthDeltaCoeffs, the non-orthogonal correction
vectors, the cell-centre differences, gh and ghf against the oracle's OpenFOAM-order arithmetic
(bit for bit), on tets, prisms, hexes and an unstructured Delaunay mesh."""
import numpy as np
import pytest

import bench
import oracle
from openfoam_tpp_b200 import meshgen as mg
from openfoam_tpp_b200 import solver as sv


def _meshes():
    C = bench.CASE
    yield "tets", mg.cylinder_mesh(C["H"], C["D"], 5, 6, "flat", "tet")
    yield "prisms", mg.cylinder_mesh(C["H"], C["D"], 5, 6, "flat", "prism")
    yield "hexes", mg.box_mesh(5, 6, 7, lo=(-0.1, -0.1, 0.0), hi=(0.1, 0.1, 0.208), cell="hex", top_patch="atmosphere")
    yield "delaunay", mg.unstructured_cylinder_mesh(C["H"], C["D"], 0.03, seed=3, iters=12)
    yield "shuffled tets", mg.shuffled(mg.cylinder_mesh(C["H"], C["D"], 4, 5, "flat", "tet"), 5)


def _ell_reference(mesh):
    """(face << 1) | neighbourSide and the other cell, per cell in ascending face order"""
    nC, nI, nF = mesh.n_cells, mesh.n_internal, mesh.n_faces
    own, nei = mesh.owner.astype(np.int64), mesh.neighbour.astype(np.int64)
    cell = np.concatenate([own, nei])
    code = np.concatenate([np.arange(nF) << 1, (np.arange(nI) << 1) | 1])
    other = np.concatenate([np.concatenate([nei, -np.ones(nF - nI, dtype=np.int64)]), own[:nI]])
    o = np.lexsort((code, cell))
    cell, code, other = cell[o], code[o], other[o]
    start = np.searchsorted(cell, np.arange(nC))
    slot = np.arange(cell.size) - start[cell]
    return cell, slot, code, other, int(slot.max()) + 1


def _check(lib, renumber=None):
    import os

    if renumber is None:
        for r in ("0", "1", "-1"):
            os.environ["TPP_RENUMBER"] = r
            try:
                _check(lib, r)
            finally:
                os.environ.pop("TPP_RENUMBER", None)
        return
    for name, mesh in _meshes():
        cfg = bench.make_config(mesh)
        g = sv.Solver(mesh, cfg, device=0, lib_path=lib)
        o = oracle.Oracle(mesh, cfg)
        nC, nCp, W, nI, nB, nG = g.get_int("layout")
        assert (nC, nI, nB, nG) == (mesh.n_cells, mesh.n_internal, mesh.n_faces - mesh.n_internal, 0), name
        # the library may renumber cells and internal faces internally (Morton order; a shuffled or
        # Delaunay mesh is, a layer-by-layer one is not): integer tables are compared through its maps,
        # which must be permutations, keep the boundary faces in place and leave every cell's slots in
        # FILE face order (that is what keeps the floating-point sums bit-exact)
        cfo, ffo = g.get_int("cellFileOf").astype(np.int64), g.get_int("faceFileOf").astype(np.int64)
        assert np.array_equal(np.sort(cfo), np.arange(nC)) and np.array_equal(np.sort(ffo[:nI]), np.arange(nI)) and np.array_equal(ffo[nI:], np.arange(nI, nI + nB)), name
        if renumber == "0":
            assert np.array_equal(cfo, np.arange(nC)) and np.array_equal(ffo, np.arange(nI + nB)), name
        if renumber == "1":
            assert not np.array_equal(cfo, np.arange(nC)), name
        assert np.array_equal(cfo[g.get_int("owner")], mesh.owner[ffo]) and np.array_equal(cfo[g.get_int("neighbour")], mesh.neighbour[ffo[:nI]]), name
        cell, slot, code, other, Wref = _ell_reference(mesh)
        assert W == Wref and nCp >= nC and nCp % 32 == 0, (name, W, Wref)
        cf, cn = g.get_int("cf").reshape(W, nCp).astype(np.int64), g.get_int("cn").reshape(W, nCp).astype(np.int64)
        ref_cf, ref_cn = -np.ones((W, nC), dtype=np.int64), -np.ones((W, nC), dtype=np.int64)
        ref_cf[slot, cell], ref_cn[slot, cell] = code, other
        got_cf = np.where(cf[:, :nC] >= 0, (ffo[np.maximum(cf[:, :nC], 0) >> 1] << 1) | (cf[:, :nC] & 1), -1)
        got_cn = np.where(cn[:, :nC] >= 0, cfo[np.maximum(cn[:, :nC], 0)], -1)
        assert np.array_equal(got_cf, ref_cf[:, cfo]) and np.array_equal(got_cn, ref_cn[:, cfo]), name
        assert np.all(cf[:, nC:] == -1) and np.all(cn[:, nC:] == -1)
        # geometry: bit for bit against the oracle's face-loop arithmetic
        Cc, Cf = o.get("C").reshape(-1, 3), o.get("Cf").reshape(-1, 3)
        own, nei = mesh.owner.astype(np.int64), mesh.neighbour.astype(np.int64)
        for nm in ("V", "Sf", "magSf", "w", "dc"):
            a, b = g.get(nm), o.get(nm)
            assert np.array_equal(a[: b.size], b[: a.size]), (name, nm, np.abs(a[: b.size] - b[: a.size]).max())
        assert np.array_equal(g.get("corrVec")[: 3 * nI], o.get("corrVec")[: 3 * nI]), name
        assert np.array_equal(g.get("dPN")[: 3 * nI], (Cc[nei] - Cc[own[:nI]]).reshape(-1)), name
        gv = np.asarray(cfg.g, dtype=np.float64)
        assert np.array_equal(g.get("gh"), gv[0] * Cc[:, 0] + gv[1] * Cc[:, 1] + gv[2] * Cc[:, 2]), name
        assert np.array_equal(g.get("ghf"), gv[0] * Cf[:, 0] + gv[1] * Cf[:, 1] + gv[2] * Cf[:, 2]), name
        g.close()


def test_geometry_and_addressing_bit_exact_emu(emu_lib):
    _check(emu_lib)


@pytest.mark.gpu
def test_geometry_and_addressing_bit_exact_gpu(gpu_lib):
    _check(None)
