"""PLOT3D reader / writer (magudi_b200/plot3d.py): round trips, the record layout of src/PLOT3DFormat.c, and -- in
the development container, where the reference tree exists -- a cross-check against the reference's own Python
reader (utils/magudi_utils/src/magudi_utils/plot3dnasa.py), which reads the files this writer produces."""
import os
import struct
import sys

import numpy as np
import pytest

from magudi_b200 import plot3d as p3d

REF_UTILS = "/root/reference/utils/magudi_utils/src"


def _blocks(rng, sizes, ncomp):
    return [rng.random((int(np.prod(s)), ncomp)) for s in sizes]


@pytest.mark.parametrize("sizes", [[(7, 5, 1)], [(6, 5, 4), (3, 4, 5)], [(9, 1, 1)]])
def test_round_trips(tmp_path, sizes):
    rng = np.random.default_rng(1)
    nD = max(j + 1 for s in sizes for j in range(3) if s[j] > 1)
    xyz = _blocks(rng, sizes, nD)
    ib = [rng.integers(0, 2, int(np.prod(s))).astype(np.int32) for s in sizes]
    g = str(tmp_path / "a.xyz")
    p3d.write_grid(g, xyz, sizes, ib)
    x2, ib2, s2 = p3d.read_grid(g)
    assert s2 == [tuple(s) for s in sizes]
    for a, b in zip(xyz, x2):
        assert np.array_equal(a, b)
    for a, b in zip(ib, ib2):
        assert np.array_equal(a, b)
    q = _blocks(rng, sizes, nD + 2)
    aux = [np.array([12.0, 0.0, 0.0, 0.6]) for _ in sizes]
    qf = str(tmp_path / "a.q")
    p3d.write_solution(qf, q, sizes, aux)
    q2, aux2, _ = p3d.read_solution(qf)
    for a, b in zip(q, q2):
        assert np.array_equal(a, b)
    assert np.array_equal(aux2[0], aux[0])
    fn = _blocks(rng, sizes, 3)
    ff = str(tmp_path / "a.f")
    p3d.write_function(ff, fn, sizes)
    f2, _ = p3d.read_function(ff)
    for a, b in zip(fn, f2):
        assert np.array_equal(a, b)
    assert p3d.detect_format(g)["fileType"] == p3d.GRID_FILE and p3d.detect_format(g)["hasIblank"]
    assert p3d.detect_format(qf)["fileType"] == p3d.SOLUTION_FILE
    assert p3d.detect_format(ff)["fileType"] == p3d.FUNCTION_FILE and p3d.detect_format(ff)["nScalars"] == 3


def test_solution_record_layout_2d(tmp_path):
    """Five slots always; 2-D leaves the fourth (rho w) unused (src/PLOT3DHelperImpl.f90:757-777)."""
    sizes = [(4, 3, 1)]
    q = np.arange(12 * 4, dtype=float).reshape(12, 4, order="F") + 1.0
    f = str(tmp_path / "b.q")
    p3d.write_solution(f, [q], sizes, [np.array([3.0, 0, 0, 0.25])])
    raw = open(f, "rb").read()
    # [4][1][4] [12][4 3 1][12] [32][aux][32] [480][5 x 12 doubles][480]
    assert struct.unpack("iii", raw[:12]) == (4, 1, 4)
    assert struct.unpack("i3ii", raw[12:32]) == (12, 4, 3, 1, 12)
    assert struct.unpack("i", raw[32:36])[0] == 32 and struct.unpack("4d", raw[36:68]) == (3.0, 0.0, 0.0, 0.25)
    assert struct.unpack("i", raw[72:76])[0] == 5 * 8 * 12
    body = np.frombuffer(raw[76:76 + 480], dtype="f8").reshape(5, 12)
    assert np.array_equal(body[0], q[:, 0]) and np.array_equal(body[2], q[:, 2])
    assert np.all(body[3] == 0.0) and np.array_equal(body[4], q[:, 3])
    assert len(raw) == 76 + 480 + 4


def test_corrupt_files_are_rejected(tmp_path):
    f = str(tmp_path / "c.xyz")
    open(f, "wb").write(struct.pack("iii", 8, 1, 8))
    with pytest.raises(p3d.Plot3DError):
        p3d.detect_format(f)


@pytest.mark.skipif(not os.path.isdir(REF_UTILS), reason="reference tree not present on this box")
def test_reference_reader_reads_our_files(tmp_path, monkeypatch):
    # the reference utility predates NumPy 2 (binary np.fromstring was removed): shim the removed call, nothing else
    monkeypatch.setattr(np, "fromstring", lambda s, dtype=float, **kw: np.frombuffer(s, dtype=dtype).copy())
    sys.path.insert(0, REF_UTILS)
    try:
        from magudi_utils import plot3dnasa as ref
    except Exception as exc:      # pragma: no cover
        pytest.skip(f"reference utilities not importable: {exc}")
    rng = np.random.default_rng(2)
    sizes = [(6, 5, 4), (3, 4, 5)]
    xyz = _blocks(rng, sizes, 3)
    q = _blocks(rng, sizes, 5)
    fn = _blocks(rng, sizes, 2)
    g, qf, ff = (str(tmp_path / n) for n in ("r.xyz", "r.q", "r.f"))
    p3d.write_grid(g, xyz, sizes)
    p3d.write_solution(qf, q, sizes, [np.array([5.0, 0, 0, 1.5])] * 2)
    p3d.write_function(ff, fn, sizes)
    G = ref.Grid(g, forceread=True)
    S = ref.Solution(qf, forceread=True)
    F = ref.Function(ff, forceread=True)
    assert G.has_iblank
    for b, s in enumerate(sizes):
        assert tuple(G.get_size(b)) == s
        assert np.array_equal(np.asarray(G.xyz[b]).reshape(-1, 3, order="F"), xyz[b])
        assert np.all(np.asarray(G.iblank[b]) == 1)
        assert np.array_equal(np.asarray(S.q[b]).reshape(-1, 5, order="F"), q[b])
        assert np.array_equal(np.asarray(F.f[b]).reshape(-1, 2, order="F"), fn[b])
