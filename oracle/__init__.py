"""CPU oracle for the GP-ODE vector-field hot path of IlzeAmandaA/VAE-GP-ODE.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product path: only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may
import it, and only as the checker or the CPU baseline being timed.  The CUDA library never calls it.

Parity status: the per-evaluation arithmetic (``field.py``) is pinned against outputs of the live
reference modules (``/root/reference/experiments/model/core/{kernels,svpy}.py``) through
``oracle/reference_harness.py`` + ``oracle/make_golden.py``; the frozen vectors live in
``tests/golden/``.  The fixed-grid solver (``solvers.py``) restates torchdiffeq's published
algorithm -- torchdiffeq is an un-vendored, un-pinned third-party dependency of the reference
(``experiments/model/core/flow.py:3-4``) and is not installed in this image, and the reference holds
no test or golden vector for it: **parity unpinned at the solver boundary** (see DESIGN.md).
"""
