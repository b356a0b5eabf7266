"""Reader / writer for the reference's text hierarchy format (``.dat``).

Format restated from the reference's importer (``include/lexls/tools.h:273-453``) and MATLAB exporter
(``interfaces/matlab-octave/tests/implementation/export_hierarchy.m:82-156``)::

    #HierType  100 equalities [A b] | 200 inequalities [A lb ub] | 210 = 200 + trailing activation type
    #nVar n    #nObj P    #nCtr m_0 .. m_{P-1}    #ObjType 100 (simple bounds, level 0 only) | 200 (general)
    #OBJECTIVE k   then m_k rows: general -> n + {1|2} numbers [+type]; simple -> var_index(1-based) {b | lb ub} [+type]
    #SolGuess n numbers (optional)   #Solution n numbers (optional)

Header keys may come in any order; extra columns are ignored.  Variable indices of a simple-bounds level
are 1-based in the file (the MEX front-end subtracts 1, ``interfaces/matlab-octave/lexlsi.cpp:412``);
:func:`read_hierarchy` returns them 0-based.
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

HIER_EQUALITIES = 100
HIER_INEQUALITIES = 200
HIER_INEQUALITIES_WITH_AS = 210
OBJ_SIMPLE_BOUNDS = 100
OBJ_GENERAL = 200


@dataclass
class Hierarchy:
    hier_type: int
    nVar: int
    ObjDim: List[int]
    ObjType: List[int]                      # 0 general, 1 simple bounds (lexls enum values)
    objectives: List[np.ndarray]            # general: m_k x (n + 1|2); simple: m_0 x (1|2) bounds only
    var_index: Optional[np.ndarray] = None  # 0-based, for a simple-bounds level 0
    active_set_guess: Optional[List[np.ndarray]] = None
    solution_guess: Optional[np.ndarray] = None
    solution: Optional[np.ndarray] = None
    extra: dict = field(default_factory=dict)


def read_hierarchy(path) -> Hierarchy:
    with open(path, "r") as f:
        lines = [ln.rstrip("\n") for ln in f]
    pos = 0
    header = {}
    keys = {"#nVar", "#nObj", "#nCtr", "#HierType", "#ObjType"}
    # header, any order (tools.h:318-354)
    while pos < len(lines) and len(header) < len(keys):
        key = "".join(lines[pos].split())
        pos += 1
        if key in keys:
            header[key] = lines[pos].split()
            pos += 1
    if len(header) < len(keys):
        raise ValueError("At least one required parameters is not set.")
    hier_type = int(header["#HierType"][0])
    nVar = int(header["#nVar"][0])
    nObj = int(header["#nObj"][0])
    ObjDim = [int(t) for t in header["#nCtr"]]
    raw_types = [int(t) for t in header["#ObjType"]]
    if len(ObjDim) != nObj or len(raw_types) != nObj:
        raise ValueError("Wrong number of objectives.")
    if hier_type == HIER_EQUALITIES:
        nb = 1
    elif hier_type in (HIER_INEQUALITIES, HIER_INEQUALITIES_WITH_AS):
        nb = 2
    else:
        raise ValueError("Unsupported type of hierarchy.")
    with_as = hier_type == HIER_INEQUALITIES_WITH_AS

    objectives, guesses, var_index = [], [], None
    k = 0
    while pos < len(lines) and k < nObj:
        key = "".join(lines[pos].split())
        pos += 1
        if not key.startswith("#OBJECTIVE"):
            continue
        simple = raw_types[k] == OBJ_SIMPLE_BOUNDS
        if simple and k != 0:
            raise ValueError("Simple constraints are supported only in the first objective.")
        ncols = (1 + nb) if simple else (nVar + nb)
        M = np.zeros((ObjDim[k], ncols))
        g = np.zeros(ObjDim[k], dtype=np.uint8)
        for r in range(ObjDim[k]):
            tok = lines[pos].split()
            pos += 1
            if len(tok) < ncols:
                raise ValueError("Not enough data.")
            M[r, :] = [float(t) for t in tok[:ncols]]
            if with_as and len(tok) > ncols:
                t = int(float(tok[ncols]))
                if t not in (0, 1, 2, 3):
                    raise ValueError("Unsupported constraint activation type.")
                g[r] = t
        if simple:
            var_index = np.rint(M[:, 0]).astype(np.int64) - 1
            if (var_index < 0).any():
                raise ValueError("variable indices in the file are 1-based")
            var_index = var_index.astype(np.uint32)
            M = M[:, 1:]
        objectives.append(M)
        guesses.append(g)
        k += 1
    if k != nObj:
        raise ValueError("The number of objectives is lower than expected.")

    def read_vec(p):
        vals = []
        while p < len(lines) and len(vals) < nVar:
            vals.extend(float(t) for t in lines[p].split())
            p += 1
        if len(vals) < nVar:
            raise ValueError("Could not read a solution vector.")
        return np.array(vals[:nVar]), p

    sol_guess = sol = None
    while pos < len(lines):
        key = lines[pos].strip()
        pos += 1
        if key == "#SolGuess":
            sol_guess, pos = read_vec(pos)
        elif key == "#Solution":
            sol, pos = read_vec(pos)
    return Hierarchy(hier_type, nVar, ObjDim, [1 if t == OBJ_SIMPLE_BOUNDS else 0 for t in raw_types], objectives,
                     var_index, guesses if with_as else None, sol_guess, sol)


def write_hierarchy(path, h: Hierarchy):
    """Inverse of :func:`read_hierarchy` (same layout as export_hierarchy.m)."""
    with_as = h.active_set_guess is not None and h.hier_type == HIER_INEQUALITIES_WITH_AS
    with open(path, "w") as f:
        f.write("#HierType\n%d\n\n#nVar\n%d\n\n#nObj\n%d\n\n" % (h.hier_type, h.nVar, len(h.ObjDim)))
        f.write("#nCtr\n%s\n\n" % " ".join(str(m) for m in h.ObjDim))
        f.write("#ObjType\n%s\n\n" % " ".join(str(OBJ_SIMPLE_BOUNDS if t == 1 else OBJ_GENERAL) for t in h.ObjType))
        for k, M in enumerate(h.objectives):
            f.write("#OBJECTIVE %d\n" % k)
            for r in range(M.shape[0]):
                row = []
                if h.ObjType[k] == 1:
                    row.append("%d" % (int(h.var_index[r]) + 1))
                row.extend(repr(float(v)) for v in M[r])
                if with_as:
                    row.append("%d" % int(h.active_set_guess[k][r]))
                f.write(" " + " ".join(row) + "\n")
            f.write("\n")
        for name, vec in (("#SolGuess", h.solution_guess), ("#Solution", h.solution)):
            if vec is not None:
                f.write(name + "\n")
                for v in vec:
                    f.write(" %s\n" % repr(float(v)))
                f.write("\n")
