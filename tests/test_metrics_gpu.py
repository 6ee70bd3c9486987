"""Loss / metric consumers (protoasnet_b200.metrics) against golden vectors produced by the reference's own classes
(oracle/gen_golden_losses.py): values, gradients and the prototype usage counters."""
import glob
import os

import numpy as np
import pytest
import torch

from protoasnet_b200 import metrics

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses_*.npz")))


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_losses_and_counters_match_reference(path):
    z = np.load(path)
    n, P, K, abstain, p, n_specific = (int(v) for v in z["meta"])
    reduction = str(z["reduction"])
    sim = torch.from_numpy(z["sim"]).cuda().requires_grad_(True)
    tgt = torch.from_numpy(z["target"]).cuda()
    occ = torch.from_numpy(z["occ"]).cuda().requires_grad_(True)
    cl = metrics.ClusterRoiFeat(0.8, num_classes=K, reduction=reduction).compute(sim, tgt)
    sp = metrics.SeparationRoiFeat(0.08, num_classes=K, reduction=reduction, abstain_class=bool(abstain)).compute(sim, tgt)
    ln = metrics.L_norm(p=p, loss_weight=1e-3, reduction=reduction).compute(occ, dim=tuple(range(-(occ.dim() - 3), 0)))
    (cl + sp + ln).backward()
    np.testing.assert_allclose(float(cl.detach()), float(z["cluster"]), rtol=1e-6)
    np.testing.assert_allclose(float(sp.detach()), float(z["separation"]), rtol=1e-6)
    np.testing.assert_allclose(float(ln.detach()), float(z["lnorm"]), rtol=1e-5)
    np.testing.assert_allclose(sim.grad.cpu().numpy(), z["grad_sim"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(occ.grad.cpu().numpy(), z["grad_occ"], rtol=1e-5, atol=1e-9)
    dc = metrics.DiversityCounters(P, n_specific=n_specific, abstain_class=bool(abstain))
    half = n // 2
    dc.update(sim.detach()[:half])
    dc.update(sim.detach()[half:])          # counters accumulate across steps
    count, cums = dc.result()
    assert np.array_equal(count, z["count"].astype(np.int64))
    np.testing.assert_allclose(cums, z["cumsum"], rtol=1e-6)


def test_zero_weight_and_fallthrough_cases():
    sim = torch.rand((4, 8), device="cuda")
    tgt = torch.tensor([0, 1, 0, 1], device="cuda")
    assert float(metrics.ClusterRoiFeat(0, num_classes=2).compute(sim, tgt)) == 0.0
    # the masked last-layer regulariser of the reference (XProtoNet_Base.py:81, :359) is not an occurrence map: plain torch
    w = torch.rand((4, 8), device="cuda")
    mask = (torch.rand((4, 8), device="cuda") > 0.5).float()
    got = metrics.L_norm(mask=mask, p=1, loss_weight=1e-4).compute(w)
    assert abs(float(got) - 1e-4 * float((mask * w).abs().sum())) < 1e-9
    with pytest.raises(Exception):
        metrics.ClusterRoiFeat(1.0, num_classes=2).compute(sim.cpu(), tgt.cpu())
