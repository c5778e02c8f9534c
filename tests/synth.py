"""Synthetic workloads live in the package (ministark_b200/synth.py) so that bench.py does
not depend on the test tree; the tests keep importing them from here."""
from ministark_b200.synth import *  # noqa: F401,F403
from ministark_b200.synth import P, SynthAir, root_of_unity, splitmix64, synth_linear_matrix, synth_trace  # noqa: F401
