"""CPU oracle for the MonolithicFEMVLFS.jl hot path (TEST INFRASTRUCTURE ONLY).

This package restates, in NumPy/SciPy, the algorithm the reference delegates to
Gridap 0.17 (`Project.toml:24`): model topology and DOF numbering, Lagrangian
reference elements, Gauss quadrature, cell-wise integration of the monolithic
potential-flow / beam / plate weak forms, COO -> CSC assembly with explicit
zeros, sparse LU solve and Newmark stepping.

PARITY STATUS: "parity unpinned" at the Gridap boundary. Julia/Gridap are not
installable in this environment and the reference's own tests assert nothing
(`test/runtests.jl:1-5`).  The oracle is pinned against what the reference does
hold: the golden mesh pair `models/multi_geo_coarse.msh` <-> `.json` (topology
numbering), the analytic travelling-wave solution and convergence rates
(`src/Periodic_Beam.jl:44-45`, `scripts/5-1-1-...jl:140-141`), analytic energies
(`src/Periodic_Beam.jl:129-132`, `src/Periodic_Beam_FS.jl:177-180`), the digitised literature
curves the reference plots against (`data/Ref_data/Khabakhpasheva`, `data/Ref_data/Liu`: the
Liu case on the reference's own `models/Liu.json` lands within 0.009 of them) and structural
identities.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import this package.  The product package
(`monolithicfemvlfs.jl_b200/`) never does.
"""
