"""TEST INFRASTRUCTURE ONLY - the CPU oracle for the gail-carla learning hot path.

Nothing in the shipped package (``gail_carla_b200``) may import this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker or as the timed CPU
baseline - never as the thing that is shipped.

* ``oracle.ref_path``  - functional torch-CPU (fp32, autograd) restatement of the
  reference's five hot-path files, one function per reference function, each
  citing the reference ``file:line`` it follows.
* ``oracle.abi_emu``   - per-entry-point CPU statement of what each C-ABI op in
  ``include/gail_carla_b200.h`` computes (layouts included), used to unit-test
  the CUDA kernels op by op and to drive the host-side classes on CPU in tests.

Pinning: the reference has no tests / golden vectors, so the
restatement is pinned against the reference itself: ``tests/golden/make_golden.py``
imports the unmodified reference modules from ``/root/reference`` in the build
container, runs them and this restatement on identical seeded inputs, asserts
agreement, and commits the reference's outputs as ``tests/golden/*.npz``.
"""
