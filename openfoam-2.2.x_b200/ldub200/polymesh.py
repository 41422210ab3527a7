"""polyMesh import: OpenFOAM's on-disk mesh -> the LDU addressing and coupled patches the solver takes.

The data format on the input side of the path.  What fvMesh does in memory
(finiteVolume/fvMesh/fvMeshLduAddressing.H:83-102: lowerAddr = owner of the internal faces,
upperAddr = neighbour, patchAddr = owner of the patch faces) is done here from the files

    constant/polyMesh/owner        labelList, one owner cell per face (internal faces first)
    constant/polyMesh/neighbour    labelList, neighbour cell of every internal face
    constant/polyMesh/boundary     polyBoundaryMesh: name { type; nFaces; startFace; [myProcNo; neighbProcNo;] }
    constant/polyMesh/points,faces (optional) geometry for face areas / Laplacian coefficients

of a case, or of every processorN/ directory of a case decomposed by decomposePar
(applications/utilities/parallelProcessing/decomposePar): one mesh region per rank, its `processor`
patches (processorPolyPatch.H:58-63) become the interfaces, both sides listing the cut faces in
the same order.  `cyclic` patches (neighbourPatch entry, cyclicPolyPatch.H) become the two halves
of a cyclic pair.  ASCII files, plain or .gz; binary-format files are refused.
"""
from __future__ import annotations

import gzip
import re
from pathlib import Path

import numpy as np


class FoamFileError(ValueError):
    pass


def _read_text(path: Path) -> str:
    path = Path(path)
    if not path.exists() and Path(str(path) + ".gz").exists():
        path = Path(str(path) + ".gz")
    raw = gzip.open(path, "rb").read() if path.suffix == ".gz" else path.read_bytes()
    text = raw.decode("latin-1")
    head = text[:2000]
    m = re.search(r"format\s+(\w+)\s*;", head)
    if m and m.group(1) != "ascii":
        raise FoamFileError(f"{path}: format {m.group(1)} is not supported (write the case with writeFormat ascii)")
    return text


def _strip_comments(text: str) -> str:
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return re.sub(r"//[^\n]*", "", text)


def _body(text: str) -> str:
    """what follows the FoamFile { ... } header"""
    text = _strip_comments(text)
    m = re.search(r"FoamFile\s*\{.*?\}", text, flags=re.S)
    return text[m.end():] if m else text


def read_label_list(path) -> np.ndarray:
    """labelList file (owner, neighbour): `N ( a b c ... )`"""
    body = _body(_read_text(Path(path)))
    m = re.search(r"(\d+)\s*\(", body)
    if not m:
        raise FoamFileError(f"{path}: no list found")
    n = int(m.group(1))
    end = body.rindex(")")
    vals = np.array(body[m.end():end].split(), dtype=np.int64)
    if vals.size != n:
        raise FoamFileError(f"{path}: header says {n} entries, found {vals.size}")
    return vals.astype(np.int32)


def read_points(path) -> np.ndarray:
    """vectorField file: `N ( (x y z) ... )` -> [N,3]"""
    body = _body(_read_text(Path(path)))
    m = re.search(r"(\d+)\s*\(", body)
    n = int(m.group(1))
    vals = np.array(body[m.end():body.rindex(")")].replace("(", " ").replace(")", " ").split(), dtype=np.float64)
    if vals.size != 3 * n:
        raise FoamFileError(f"{path}: header says {n} points, found {vals.size / 3}")
    return vals.reshape(n, 3)


def read_faces(path):
    """faceList file: `N ( k(p0 p1 ...) ... )` -> list of int arrays"""
    body = _body(_read_text(Path(path)))
    m = re.search(r"(\d+)\s*\(", body)
    n = int(m.group(1))
    faces = [np.array(g.split(), dtype=np.int64) for g in re.findall(r"\d+\s*\(([^()]*)\)", body[m.end():])]
    if len(faces) != n:
        raise FoamFileError(f"{path}: header says {n} faces, found {len(faces)}")
    return faces


def read_boundary(path):
    """polyBoundaryMesh -> list of dict(name, type, nFaces, startFace, + the patch's other scalar entries)"""
    body = _body(_read_text(Path(path)))
    m = re.search(r"(\d+)\s*\(", body)
    if not m:
        raise FoamFileError(f"{path}: no patch list found")
    n = int(m.group(1))
    patches = []
    for pm in re.finditer(r"([A-Za-z_][\w.:-]*)\s*\{([^{}]*)\}", body[m.end():]):
        entries = dict(name=pm.group(1))
        for em in re.finditer(r"(\w+)\s+([^;]*);", pm.group(2)):
            key, val = em.group(1), em.group(2).strip()
            entries[key] = int(val) if re.fullmatch(r"-?\d+", val) else val
        for need in ("type", "nFaces", "startFace"):
            if need not in entries:
                raise FoamFileError(f"{path}: patch {entries['name']} has no {need}")
        patches.append(entries)
    if len(patches) != n:
        raise FoamFileError(f"{path}: header says {n} patches, found {len(patches)}")
    return patches


def read_poly_mesh(mesh_dir, geometry=False):
    """constant/polyMesh -> dict(nCells, nFaces (internal), lower, upper, owner (all faces), patches,
    [points, faces])."""
    mesh_dir = Path(mesh_dir)
    owner = read_label_list(mesh_dir / "owner")
    neighbour = read_label_list(mesh_dir / "neighbour")
    patches = read_boundary(mesh_dir / "boundary")
    n_int = neighbour.size
    n_cells = int(owner.max()) + 1 if owner.size else 0
    lower, upper = owner[:n_int].copy(), neighbour.copy()
    if n_int and (np.any(lower >= upper) or np.any(np.diff(lower) < 0)):
        raise FoamFileError(f"{mesh_dir}: internal faces are not in upper-triangular order")
    for p in patches:
        if p["startFace"] < n_int or p["startFace"] + p["nFaces"] > owner.size:
            raise FoamFileError(f"{mesh_dir}: patch {p['name']} lies outside the boundary faces")
        p["faceCells"] = owner[p["startFace"]:p["startFace"] + p["nFaces"]].copy()
    mesh = dict(nCells=n_cells, nFaces=n_int, lower=lower, upper=upper, owner=owner, patches=patches)
    if geometry:
        mesh["points"] = read_points(mesh_dir / "points")
        mesh["faces"] = read_faces(mesh_dir / "faces")
    return mesh


# --------------------------------------------------------------------------- #
# geometry -> the coefficients fvm::laplacian would assemble
# --------------------------------------------------------------------------- #
def face_geometry(points, faces):
    """face centres and area vectors (fan about the vertex average, area-weighted centres; the
    construction of primitiveMeshFaceCentresAndAreas.C:60-125)"""
    nf = len(faces)
    Cf, Sf = np.zeros((nf, 3)), np.zeros((nf, 3))
    by_size = {}
    for i, f in enumerate(faces):
        by_size.setdefault(f.size, []).append(i)
    for k, idx in by_size.items():
        idx = np.array(idx)
        P = points[np.array([faces[i] for i in idx])]             # [n,k,3]
        if k == 3:
            Cf[idx] = P.mean(axis=1)
            Sf[idx] = 0.5 * np.cross(P[:, 1] - P[:, 0], P[:, 2] - P[:, 0])
            continue
        centre = P.mean(axis=1)
        sumN, sumA, sumAc = np.zeros((idx.size, 3)), np.zeros(idx.size), np.zeros((idx.size, 3))
        for j in range(k):
            a, b = P[:, j], P[:, (j + 1) % k]
            n = np.cross(b - a, centre - a)
            c3 = a + b + centre
            area = np.linalg.norm(n, axis=1)
            sumN += n
            sumA += area
            sumAc += area[:, None] * c3
        Cf[idx] = sumAc / (3.0 * sumA[:, None])
        Sf[idx] = 0.5 * sumN
    return Cf, Sf


def laplacian_system(mesh, dirichlet_types=("patch",), variable=False, asym=0.0):
    """The system fvm::laplacian(gamma, p) == source assembles on the mesh (orthogonal part:
    upper = gamma |Sf| / |d|, diag = -sum, gaussLaplacianScheme.C:50-88), Dirichlet on the patches
    whose type is in dirichlet_types, zero gradient elsewhere (`empty` patches take no part).
    Same dict layout as ldub200.meshes.laplacian_system, plus faceWeights for faceAreaPair
    (faceAreaPairGAMGAgglomeration.C:58-66)."""
    Cf, Sf = face_geometry(mesh["points"], mesh["faces"])
    owner, n, nf = mesh["owner"], mesh["nCells"], mesh["nFaces"]
    magSf = np.linalg.norm(Sf, axis=1)
    # cell centres: area-weighted mean of the face centres
    C = np.zeros((n, 3))
    wsum = np.zeros(n)
    np.add.at(C, owner, magSf[:, None] * Cf)
    np.add.at(wsum, owner, magSf)
    np.add.at(C, mesh["upper"], magSf[:nf, None] * Cf[:nf])
    np.add.at(wsum, mesh["upper"], magSf[:nf])
    C /= wsum[:, None]
    l, u = mesh["lower"].astype(np.int64), mesh["upper"].astype(np.int64)
    d = np.linalg.norm(C[u] - C[l], axis=1)
    gamma = 1.0 + (0.3 * np.sin(0.01 * np.arange(nf)) if variable else 0.0)
    upper = gamma * magSf[:nf] / d
    lower = upper * (1.0 - asym * (0.5 + 0.5 * np.cos(0.02 * np.arange(nf)))) if asym else None
    diag = np.zeros(n)
    np.subtract.at(diag, l, upper if lower is None else lower)
    np.subtract.at(diag, u, upper)
    n_dirichlet = 0
    for p in mesh["patches"]:
        if p["type"] in dirichlet_types and p["nFaces"]:
            fs = np.arange(p["startFace"], p["startFace"] + p["nFaces"])
            db = np.linalg.norm(Cf[fs] - C[owner[fs]], axis=1)
            np.subtract.at(diag, owner[fs], magSf[fs] / db)
            n_dirichlet += p["nFaces"]
    if n_dirichlet == 0 and n:      # fvMatrix::setReference(0, 0)
        diag[0] += diag[0]
    s = np.sqrt(magSf[:nf])[:, None]
    weights = np.linalg.norm(Sf[:nf] / s * np.array([1.0, 1.01, 1.02]), axis=1)
    return dict(nCells=n, nFaces=nf, lower=mesh["lower"], upper=mesh["upper"], diag=diag, upperCoef=upper,
                lowerCoef=lower, source=np.sin(0.37 * np.arange(n)) * np.abs(diag).mean(), psi0=np.zeros(n),
                faceWeights=weights)


# --------------------------------------------------------------------------- #
# decomposed cases: processorN/constant/polyMesh
# --------------------------------------------------------------------------- #
def read_decomposed_case(case_dir):
    """processor0..N-1 of a decomposePar'd case -> one region per rank: dict(nCells, nFaces, lower,
    upper, patches, interfaces=[dict(nbrRegion, nbrInterface, faceCells, patch)]) in rank order.
    Interface lists hold the coupled patches (processor, cyclic) in patch order."""
    case_dir = Path(case_dir)
    dirs = sorted((d for d in case_dir.iterdir() if re.fullmatch(r"processor\d+", d.name)),
                  key=lambda d: int(d.name[9:]))
    if not dirs or [int(d.name[9:]) for d in dirs] != list(range(len(dirs))):
        raise FoamFileError(f"{case_dir}: processor directories are not 0..N-1")
    regions = [read_poly_mesh(d / "constant" / "polyMesh") for d in dirs]
    for r, reg in enumerate(regions):
        reg["interfaces"] = []
        for p in reg["patches"]:
            if p["type"] == "processor":
                if p.get("myProcNo") != r:
                    raise FoamFileError(f"{dirs[r]}: patch {p['name']} says myProcNo {p.get('myProcNo')}")
                reg["interfaces"].append(dict(nbrRegion=int(p["neighbProcNo"]), faceCells=p["faceCells"],
                                              patch=p["name"]))
            elif p["type"] == "cyclic":
                reg["interfaces"].append(dict(nbrRegion=r, faceCells=p["faceCells"], patch=p["name"],
                                              neighbourPatch=p["neighbourPatch"]))
    for r, reg in enumerate(regions):
        seen = {}
        for it in reg["interfaces"]:
            s = it["nbrRegion"]
            if "neighbourPatch" in it:
                match = [k for k, jt in enumerate(reg["interfaces"]) if jt["patch"] == it["neighbourPatch"]]
            else:
                # k-th patch towards rank s <-> k-th patch of rank s towards here
                k = seen.get(s, 0)
                seen[s] = k + 1
                back = [j for j, jt in enumerate(regions[s]["interfaces"])
                        if jt["nbrRegion"] == r and "neighbourPatch" not in jt]
                match = back[k:k + 1]
            if not match:
                raise FoamFileError(f"processor{r}: no counterpart for patch {it['patch']}")
            other = regions[s]["interfaces"][match[0]]
            if other["faceCells"].size != it["faceCells"].size:
                raise FoamFileError(f"processor{r}: patch {it['patch']} and its counterpart differ in size")
            it["nbrInterface"] = match[0]
    return regions


_HEADER = """FoamFile
{{
    version     2.0;
    format      ascii;
    class       {cls};
    location    "constant/polyMesh";
    object      {obj};
}}

"""


def _write_label_list(path, obj, vals):
    with open(path, "w") as fh:
        fh.write(_HEADER.format(cls="labelList", obj=obj))
        fh.write(f"{len(vals)}\n(\n")
        fh.write("\n".join(str(int(v)) for v in vals))
        fh.write("\n)\n")


def write_decomposed_case(case_dir, regions):
    """Write owner / neighbour / boundary of every region (ldub200.decompose layout) as
    processorN/constant/polyMesh — the connectivity decomposePar writes (no points/faces: the LDU
    path needs none).  Patch faces come after the internal faces, one `processor` patch per
    interface to another region, `cyclic` pairs for interfaces inside a region."""
    case_dir = Path(case_dir)
    for r, reg in enumerate(regions):
        d = case_dir / f"processor{r}" / "constant" / "polyMesh"
        d.mkdir(parents=True, exist_ok=True)
        owner = [np.asarray(reg["lower"], dtype=np.int64)]
        start = int(np.asarray(reg["lower"]).size)
        entries = []
        its = reg.get("interfaces", [])
        for i, it in enumerate(its):
            fc = np.asarray(it["faceCells"], dtype=np.int64)
            if it["nbrRegion"] == r:
                name = f"cyclic_{i}"
                body = (f"    type cyclic;\n    nFaces {fc.size};\n    startFace {start};\n"
                        f"    neighbourPatch cyclic_{it['nbrInterface']};\n")
            else:
                name = f"procBoundary{r}to{it['nbrRegion']}" + (f"_{i}" if sum(
                    1 for jt in its if jt["nbrRegion"] == it["nbrRegion"]) > 1 else "")
                body = (f"    type processor;\n    nFaces {fc.size};\n    startFace {start};\n"
                        f"    myProcNo {r};\n    neighbProcNo {it['nbrRegion']};\n")
            entries.append(f"{name}\n{{\n{body}}}\n")
            owner.append(fc)
            start += fc.size
        _write_label_list(d / "owner", "owner", np.concatenate(owner))
        _write_label_list(d / "neighbour", "neighbour", reg["upper"])
        with open(d / "boundary", "w") as fh:
            fh.write(_HEADER.format(cls="polyBoundaryMesh", obj="boundary"))
            fh.write(f"{len(entries)}\n(\n" + "\n".join(entries) + ")\n")
