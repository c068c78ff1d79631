"""GPU: BASELINE.json's strategy-level configs (#1 DYCORS/Ackley-10, #2 SRBF/Levy-30 async 8 workers,
#5 SOP/Rastrigin-20 with predict_deriv), proposal by proposal.

tools/make_strategy_traces.py ran the UNMODIFIED reference strategies with the reference surrogate and stored,
for every candidate_dycors / candidate_srbf call they made, the arguments, the RNG state and the returned points.
Here every call is replayed through pysot_b200's candidate functions on the GPU with the same arguments and the
same RNG stream (the recorded points are fed back, so a tie cannot derail later proposals).

Bar: the proposed points are identical; a different pick is accepted only where the reference's own best and
second-best merits are within 1e-12 (north_star).
"""
import numpy as np
import pytest

from trace_replay import GpuBackend, replay

pytestmark = pytest.mark.gpu


def test_config1_dycors_ackley10_serial():
    ncalls, mismatches, ties, _, _ = replay("trace_dycors.npz", GpuBackend())
    assert ncalls > 400 and not mismatches, (ncalls, ties, mismatches[:5])


def test_config2_srbf_levy30_async8():
    ncalls, mismatches, ties, _, _ = replay("trace_srbf.npz", GpuBackend())
    assert ncalls > 1500 and not mismatches, (ncalls, ties, mismatches[:5])


def test_config5_sop_rastrigin20_with_gradients():
    ncalls, mismatches, ties, worst_pred, worst_deriv = replay("trace_sop.npz", GpuBackend())
    assert ncalls > 400 and not mismatches, (ncalls, ties, mismatches[:5])
    # the reference extends its LU incrementally, the GPU refits: the reference's own reorder noise
    assert worst_pred < 1e-9 and worst_deriv < 1e-9, (worst_pred, worst_deriv)
