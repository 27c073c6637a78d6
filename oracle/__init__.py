"""CPU oracle for the batched drone hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it, and there only as the checker or as the timed
CPU baseline.  The product package ``multidronesim_b200`` never imports it.

Parity status (see DESIGN.md section "Oracle"):

* trajectories, controllers, low-level PID, linear models, nonlinear xdot,
  conversions, CBF row builder: restated in numpy fp64 and PINNED against the
  reference's own Python modules imported in the build container
  (``oracle/ref_import.py`` + ``oracle/make_golden.py`` ->
  ``tests/golden/*.npz``).
* env step (upstream gym-pybullet-drones ``BaseAviary``/``CtrlAviary``,
  un-vendored and un-pinned by the reference), ``DSLPIDControl`` and the QP
  solver (cvxopt 1.3.2, not installed): PARITY UNPINNED -- restated from the
  published algorithms; the QP oracle is validated by a KKT
  certificate instead of by cvxopt output.
"""
