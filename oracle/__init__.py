"""CPU oracle: a NumPy restatement of magudi's RHS / adjoint / RK4 hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``magudi_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline``
/ ``--impl reference`` legs of ``bench.py`` use it, and only as the checker
(or the timed CPU baseline), never as the product path.

Every function cites the reference file:line it restates (paths relative to
the upstream repository root, dreamer2368/magudi).

PARITY STATUS: **parity unpinned**.  The reference ships no golden vectors and
cannot be compiled in this environment (no Fortran compiler, no MPI).  The
oracle is pinned instead by the reference's own *property* tests (stencil
order conditions, SBP property, dissipation self-adjointness, adjoint
relation, flux-Jacobian finite differences, SAT consistency), restated in
``tests/`` at the reference's tolerances, and by an independent copy of the
first-derivative tables in the reference's Python utilities
(``utils/magudi_utils/src/magudi_utils/SummationByParts.py``), see
``tests/golden/``.
"""
