"""GPU parity of the whole hot path: one training-iteration slice (tools/learn.py:137-223,269) through the drop-in
classes on CUDA vs the outputs of the UNMODIFIED reference stored in tests/golden/update_*.npz.

Tolerances (fp32 reference vs tcgen05 kind::tf32 contractions with fp32 accumulation), set from the deviations measured
on B200 (profiles/r02_parity_probe.txt, scaled error |got-ref| / (1e-3 + |ref|)):
  * quantities produced BEFORE any optimiser step (bootstrap value, compute_loss before): rtol 5e-3 (+ atol 5e-4) in
    every case; observed 1e-4..2.2e-3 - this is the TF32 level of a forward pass;
  * quantities produced AFTER optimiser steps inherit Adam's sign-like first steps (|step| = lr whatever |g|): a gradient
    element whose sign is decided by TF32 noise moves its parameter by 2*lr, so a T*N = 12-sample case (update_tiny,
    update_tiny2, update_unclipped) shows 2e-2..6e-2 on the critic statistics while the B = 256 case (update_mid: two
    critic and two PPO minibatches with the BC mix, multi-tile / split-K shapes) stays at 1e-3..2e-2; per-case rtol below;
  * post-update parameters: max |diff| <= 2.5*lr*steps + 1e-3*|p|, and mean |diff| <= mean_frac*lr with mean_frac 0.15
    for the B >= 128 cases (observed 0.03 / 0.08) and 0.5 for the 12-sample cases (observed 0.05..0.25).
The gradient-level comparison (tests/test_grads_gpu.py) is the sharper instrument; this file checks the whole slice.
"""
import pytest

import test_host_cpu as H

pytestmark = pytest.mark.gpu


CASES = [("update_tiny", 2e-2, 0.5), ("update_tiny2", 5e-2, 0.5), ("update_c1", 2e-2, 0.15), ("update_unclipped", 8e-2, 0.5),
         ("update_mid", 2e-2, 0.15)]


@pytest.mark.parametrize("name,tol,mean_frac", CASES)
def test_update_iteration_matches_reference_gpu(name, tol, mean_frac):
    out = H.run_update_case(name, "cuda")
    H.check_update_case(*out, tol=tol, mean_frac=mean_frac, pre_tol=5e-3)


@pytest.mark.parametrize("name,tol,mean_frac", [("update_tiny", 2e-2, 0.5), ("update_mid", 2e-2, 0.15)])
def test_update_iteration_uint8_store_matches_reference_gpu(name, tol, mean_frac):
    """Rollout in the byte store + uint8 expert batches (lossless: observations are uint8/255 by construction)."""
    import torch
    out = H.run_update_case(name, "cuda", obs_dtype=torch.uint8, expert_u8=True)
    H.check_update_case(*out, tol=tol, mean_frac=mean_frac, pre_tol=5e-3)


def test_cpu_tensors_fail_loudly():
    """No CPU fallback: CPU tensors or a CPU-resident module raise instead of silently computing elsewhere."""
    import torch
    from types import SimpleNamespace as NS
    import gail_carla_b200 as G
    from gail_carla_b200 import _abi as A
    with pytest.raises(RuntimeError):
        A.reward_epilogue(torch.zeros(4), torch.zeros(4), 4)
    pol = G.Policy((3, 192, 192), NS(shape=(4,)), NS(shape=(2,)), True, [-1.4, -3.2], False)
    with pytest.raises(RuntimeError):
        pol.get_value(torch.zeros(1, 3, 192, 192), torch.zeros(1, 4))
