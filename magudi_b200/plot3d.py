"""Serial PLOT3D reader / writer (multi-block, whole format, Fortran unformatted records): the files the hot path is
fed from and writes to -- ``.xyz`` grids (with IBLANK), ``.q`` solutions, ``.f`` function files -- so that
``magudi.inp`` / ``bc.dat`` / PLOT3D cases run unchanged around the CUDA library.

Host-side I/O only (no numerical work); the format is the reference's:

* record markers are 4-byte integers (8 with ``USE_EXTENDED_PLOT3D``): ``src/PLOT3DFormat.c:22-31``
* ``[4][nGrids:int32][4]``, then ``[12 nGrids][ni nj nk : int32 x 3 per grid][12 nGrids]`` -- function files carry
  a fourth integer per grid, the number of scalars (``16 nGrids``): ``src/PLOT3DFormat.c:230-262``
* grid record per block: x, y, z (each ni*nj*nk doubles, i fastest) and IBLANK (int32): ``:263-272``
* solution: per block a record of 4 doubles (aux; aux[0] = timestep as a real, aux[3] = time,
  ``src/RegionImpl.f90:1598-1609``) and a record of ALWAYS five components rho, rho u, rho v, rho w, rho E; in 2-D
  the fourth slot (1-D: the third and fourth) is left unused: ``src/PLOT3DHelperImpl.f90:757-777``
* function: per block one record of nScalars components: ``src/PLOT3DFormat.c:288-296``

Arrays follow the library's layout: ``(N, nComp)`` with ``N = ni*nj*nk`` and point index ``i + ni*(j + nj*k)``.
"""
from __future__ import annotations

import os
import struct

import numpy as np

GRID_FILE, SOLUTION_FILE, FUNCTION_FILE = 0, 1, 2


class Plot3DError(Exception):
    pass


def _marker(extended):
    return ("q", 8) if extended else ("i", 4)


def detect_format(filename, include_function_files=True, extended=False):
    """``plot3dDetectFormat`` (``src/PLOT3DFormat.c:45-214``): returns a dict with nGrids, nDimensions, nScalars,
    isEndiannessNative, hasIblank, fileType and gridSizes (nGrids x 3)."""
    code, msz = _marker(extended)
    size = os.path.getsize(filename)
    with open(filename, "rb") as f:
        raw = f.read(msz)
        if len(raw) < msz:
            raise Plot3DError(f"{filename}: file is empty")
        end = "="
        if struct.unpack(end + code, raw)[0] != 4:
            end = ">" if struct.pack("=i", 1) == struct.pack("<i", 1) else "<"
            if struct.unpack(end + code, raw)[0] != 4:
                raise Plot3DError(f"{filename}: not a multi-block whole-format PLOT3D file")
        # record markers must be consistent up to EOF
        f.seek(0)
        pos = 0
        records = []
        while pos < size:
            n = struct.unpack(end + code, f.read(msz))[0]
            f.seek(n, 1)
            tail = f.read(msz)
            if len(tail) < msz or struct.unpack(end + code, tail)[0] != n:
                raise Plot3DError(f"{filename}: leading / trailing record sizes do not match")
            records.append((pos + msz, n))
            pos += n + 2 * msz
        f.seek(records[0][0])
        nGrids = struct.unpack(end + "i", f.read(4))[0]
        if nGrids <= 0:
            raise Plot3DError(f"{filename}: number of grids must be positive")
        off, n = records[1]
        if n == 12 * nGrids:
            fileType, per = GRID_FILE, 3
        elif n == 16 * nGrids and include_function_files:
            fileType, per = FUNCTION_FILE, 4
        else:
            raise Plot3DError(f"{filename}: invalid size header")
        f.seek(off)
        hdr = np.frombuffer(f.read(n), dtype=end + "i4").reshape(nGrids, per)
        if np.any(hdr <= 0):
            raise Plot3DError(f"{filename}: grid sizes must be positive")
        sizes = hdr[:, :3].astype(int)
        nScalars, hasIblank = 0, False
        if fileType == FUNCTION_FILE:
            nScalars = int(hdr[0, 3])
            if np.any(hdr[:, 3] != nScalars):
                raise Plot3DError(f"{filename}: number of scalars differs between grids")
        else:
            if len(records) > 2 and records[2][1] == 32:
                fileType = SOLUTION_FILE
            for i in range(nGrids):
                npts = int(np.prod(sizes[i]))
                rec = records[2 + (2 * i + 1 if fileType == SOLUTION_FILE else i)][1]
                if fileType == SOLUTION_FILE:
                    ok = rec == 5 * 8 * npts
                else:
                    ok = rec in (3 * 8 * npts, (3 * 8 + 4) * npts)
                    hasIblank = rec == (3 * 8 + 4) * npts
                if not ok:
                    raise Plot3DError(f"{filename}: unexpected record size for grid {i + 1}")
    nD = 1
    for j in range(3):
        if np.any(sizes[:, j] > 1):
            nD = j + 1
    return dict(nGrids=nGrids, nDimensions=nD, nScalars=nScalars, isEndiannessNative=(end == "="),
                hasIblank=hasIblank, fileType=fileType, gridSizes=sizes, endian=end, records=records)


def _write_header(f, sizes, nScalars, fileType, extended):
    code, _ = _marker(extended)
    nGrids = len(sizes)
    f.write(struct.pack(code, 4) + struct.pack("i", nGrids) + struct.pack(code, 4))
    per = 4 if fileType == FUNCTION_FILE else 3
    f.write(struct.pack(code, 4 * per * nGrids))
    for s in sizes:
        s3 = list(s) + [1] * (3 - len(s))
        f.write(struct.pack("3i", *s3))
        if fileType == FUNCTION_FILE:
            f.write(struct.pack("i", nScalars))
    f.write(struct.pack(code, 4 * per * nGrids))


def _size3(s):
    return tuple(int(v) for v in (list(s) + [1] * (3 - len(s))))


def write_grid(filename, coordinates, sizes, iblank=None, extended=False):
    """coordinates: list (one per block) of (N, nD) arrays; iblank: list of (N,) int arrays or None (all 1)."""
    code, _ = _marker(extended)
    with open(filename, "wb") as f:
        _write_header(f, sizes, 0, GRID_FILE, extended)
        for b, (xyz, s) in enumerate(zip(coordinates, sizes)):
            n = int(np.prod(_size3(s)))
            xyz = np.asarray(xyz, dtype=np.float64).reshape(n, -1)
            rec = (3 * 8 + 4) * n
            f.write(struct.pack(code, rec))
            for d in range(3):
                col = xyz[:, d] if d < xyz.shape[1] else np.zeros(n)
                f.write(np.ascontiguousarray(col, dtype="=f8").tobytes())
            ib = np.ones(n, dtype="=i4") if iblank is None or iblank[b] is None else np.asarray(iblank[b], dtype="=i4")
            f.write(np.ascontiguousarray(ib.reshape(n)).tobytes())
            f.write(struct.pack(code, rec))


def read_grid(filename, extended=False):
    """Returns (coordinates, iblank, sizes): per block (N, nD) doubles, (N,) int32, (ni, nj, nk)."""
    fmt = detect_format(filename, extended=extended)
    if fmt["fileType"] != GRID_FILE:
        raise Plot3DError(f"{filename}: not a grid file")
    nD, end = fmt["nDimensions"], fmt["endian"]
    coords, iblanks = [], []
    with open(filename, "rb") as f:
        for b in range(fmt["nGrids"]):
            off, _ = fmt["records"][2 + b]
            n = int(np.prod(fmt["gridSizes"][b]))
            f.seek(off)
            xyz = np.frombuffer(f.read(3 * 8 * n), dtype=end + "f8").reshape(3, n).T
            coords.append(np.array(xyz[:, :nD], dtype=np.float64, order="F"))
            if fmt["hasIblank"]:
                iblanks.append(np.frombuffer(f.read(4 * n), dtype=end + "i4").astype(np.int32))
            else:
                iblanks.append(np.ones(n, dtype=np.int32))
    return coords, iblanks, [tuple(s) for s in fmt["gridSizes"]]


def write_solution(filename, solutions, sizes, aux=None, extended=False):
    """solutions: list of (N, nD + 2) conserved-variable arrays; aux: list of 4 doubles per block
    (timestep, -, -, time).  Five components are always written; the unused momentum slots stay zero."""
    code, _ = _marker(extended)
    with open(filename, "wb") as f:
        _write_header(f, sizes, 0, SOLUTION_FILE, extended)
        for b, (q, s) in enumerate(zip(solutions, sizes)):
            n = int(np.prod(_size3(s)))
            q = np.asarray(q, dtype=np.float64).reshape(n, -1)
            a = np.zeros(4) if aux is None else np.asarray(aux[b], dtype=np.float64).reshape(4)
            f.write(struct.pack(code, 32) + a.astype("=f8").tobytes() + struct.pack(code, 32))
            f.write(struct.pack(code, 5 * 8 * n))
            nU = q.shape[1]
            slots = [np.zeros(n)] * 5
            slots = list(slots)
            for c in range(nU - 1):
                slots[c] = q[:, c]
            slots[4] = q[:, nU - 1]
            for c in range(5):
                f.write(np.ascontiguousarray(slots[c], dtype="=f8").tobytes())
            f.write(struct.pack(code, 5 * 8 * n))


def read_solution(filename, extended=False):
    """Returns (solutions, aux, sizes): per block (N, nD + 2) conserved variables and the 4 aux doubles."""
    fmt = detect_format(filename, extended=extended)
    if fmt["fileType"] != SOLUTION_FILE:
        raise Plot3DError(f"{filename}: not a solution file")
    nD, end = fmt["nDimensions"], fmt["endian"]
    out, auxs = [], []
    with open(filename, "rb") as f:
        for b in range(fmt["nGrids"]):
            n = int(np.prod(fmt["gridSizes"][b]))
            f.seek(fmt["records"][2 + 2 * b][0])
            auxs.append(np.frombuffer(f.read(32), dtype=end + "f8").astype(np.float64))
            f.seek(fmt["records"][3 + 2 * b][0])
            q5 = np.frombuffer(f.read(5 * 8 * n), dtype=end + "f8").reshape(5, n)
            q = np.empty((n, nD + 2), order="F")
            q[:, :nD + 1] = q5[:nD + 1].T
            q[:, nD + 1] = q5[4]
            out.append(q)
    return out, auxs, [tuple(s) for s in fmt["gridSizes"]]


def write_function(filename, functions, sizes, extended=False):
    """functions: list of (N, nScalars) arrays (same nScalars on every block)."""
    code, _ = _marker(extended)
    nS = np.asarray(functions[0]).reshape(int(np.prod(_size3(sizes[0]))), -1).shape[1]
    with open(filename, "wb") as f:
        _write_header(f, sizes, nS, FUNCTION_FILE, extended)
        for fn, s in zip(functions, sizes):
            n = int(np.prod(_size3(s)))
            a = np.asarray(fn, dtype=np.float64).reshape(n, -1)
            if a.shape[1] != nS:
                raise Plot3DError("function file: every block must hold the same number of scalars")
            f.write(struct.pack(code, nS * 8 * n))
            f.write(np.asfortranarray(a, dtype="=f8").T.tobytes())
            f.write(struct.pack(code, nS * 8 * n))


def read_function(filename, extended=False):
    fmt = detect_format(filename, extended=extended)
    if fmt["fileType"] != FUNCTION_FILE:
        raise Plot3DError(f"{filename}: not a function file")
    end, nS = fmt["endian"], fmt["nScalars"]
    out = []
    with open(filename, "rb") as f:
        for b in range(fmt["nGrids"]):
            n = int(np.prod(fmt["gridSizes"][b]))
            f.seek(fmt["records"][2 + b][0])
            a = np.frombuffer(f.read(nS * 8 * n), dtype=end + "f8").reshape(nS, n).T
            out.append(np.array(a, dtype=np.float64, order="F"))
    return out, [tuple(s) for s in fmt["gridSizes"]]
