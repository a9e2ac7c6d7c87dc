"""CPU oracle for the GPflux SVGP hot path.  TEST INFRASTRUCTURE ONLY.

PARITY UNPINNED: the reference (GPflux 0.4.4) holds no arithmetic for this path; the maths lives
in GPflow >= 2.9.2 / TensorFlow / TFP, none of which is installed here, and the reference's tests
store no golden numbers.  This package restates GPflow's published algorithm
 twice, independently:

* ``oracle.svgp_numpy``  - numpy/scipy float64, forward only;
* ``oracle.svgp_torch``  - torch float64 on the CPU with autograd (gradient ground truth).

``oracle.natgrad_torch`` restates GPflow's natural-gradient step the same way.  ``oracle.philox_numpy``
is different in kind: it checks an ADDITION of this repository (noise generated inside the sampling
kernel), not a reference behaviour, and is pinned to three Philox4x32-10 known-answer vectors recalled
from Random123's published kat_vectors file, which is not available offline (see its header).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package.  Nothing under ``gpflux_b200/`` does.
"""
