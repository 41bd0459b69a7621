"""Gradient-level check of the hand-derived backward passes (host logic on the emulated ABI, fp32 throughout) against
torch autograd on the oracle restatement: algo/ppo.py:64-114 (PPO minibatch, with and without the BC mix, clipped and
unclipped value loss) and algo/wdgail.py:112-139 (critic minibatch incl. the gradient penalty's double backward).
Tolerance: fp32 re-association only - cosine >= 0.99999, rel-Frobenius <= 2e-4 per parameter tensor."""
import pytest

import grad_cases as GC


@pytest.mark.parametrize("B,Be,clipped", [(6, 0, True), (5, 3, True), (4, 0, False)])
def test_policy_gradients_match_autograd_cpu(emulated_abi, B, Be, clipped):
    got, ref = GC.policy_grads("cpu", B, Be, clipped=clipped)
    GC.compare(got, ref, 0.99999, 2e-4, "policy")


def test_critic_gradients_match_autograd_cpu(emulated_abi):
    got, ref, gs, rs = GC.critic_grads("cpu", 4)
    GC.compare(got, ref, 0.99999, 2e-4, "critic")
    assert abs(gs["wd"] - rs["wd"]) <= 1e-5 + 1e-4 * abs(rs["wd"])
    assert abs(gs["gp"] - rs["gp"]) <= 1e-5 + 1e-4 * abs(rs["gp"])
