"""Stage-by-stage parity driver: the CUDA path (or its host emulation) against the CPU
oracle on identical inputs.  Integer/index work and every explicit kernel must agree bit for
bit; the pressure solve agrees to the solver tolerance (different preconditioners)."""
import numpy as np

STATE = ["alpha", "alpha_b", "U", "U_b", "p_rgh", "p_rgh_b", "pGrad_b", "p", "rho", "rho_b", "phi", "Uf", "U0", "U0_b", "rho0", "Uf0", "meshPhi", "alphaPhi", "rhoPhi"]


def sync_geometry(g, o, mesh, cfg):
    """Give the GPU solver the oracle's (recomputed-from-points) geometry so later stages see
    identical inputs; the rigid-transform path is compared separately."""
    nI = mesh.n_internal
    own, nei = mesh.owner.astype(np.int64), mesh.neighbour.astype(np.int64)
    C = o.get("C").reshape(-1, 3)
    Cf = o.get("Cf").reshape(-1, 3)
    g.set("Sf", o.get("Sf"))
    g.set("magSf", o.get("magSf"))
    g.set("w", o.get("w"))
    g.set("dc", o.get("dc"))
    g.set("corrVec", o.get("corrVec")[: 3 * nI])
    g.set("dPN", (C[nei] - C[own[:nI]]).reshape(-1))
    g.set("V", o.get("V"))
    gv = cfg.g
    g.set("gh", gv[0] * C[:, 0] + gv[1] * C[:, 1] + gv[2] * C[:, 2])
    g.set("ghf", gv[0] * Cf[:, 0] + gv[1] * Cf[:, 1] + gv[2] * Cf[:, 2])


def sync_state(g, o, names=STATE):
    for nm in names:
        g.set(nm, o.get(nm))


def compare(g, o, names, exact=True, rtol=0.0, report=None):
    bad = []
    for nm in names:
        a, b = g.get(nm), o.get(nm)
        n = min(a.size, b.size)
        a, b = a[:n], b[:n]
        if exact:
            ok = np.array_equal(a, b)
        else:
            ok = np.all(np.abs(a - b) <= rtol * max(np.abs(b).max(), 1e-300))
        err = float(np.abs(a - b).max()) if n else 0.0
        if report is not None:
            report.append((nm, ok, err, float(np.abs(b).max()) if n else 0.0))
        if not ok:
            bad.append((nm, err, float(np.abs(b).max())))
    return bad


ALPHA_OUT = ["alpha", "alpha_b", "grad:gradAlpha", "phiBD", "phiCorr", "lambda", "alphaPhiUn"]


def get_pair(g, o, name):
    gn, on = (name.split(":") + [name])[:2] if ":" in name else (name, name)
    return g.get(gn), o.get(on)
