"""-m gpu: the batch memory plan of get_log_ZSMC (chunk= / return_Xs= / accumulate_grad=) is invisible in the results:
streaming the batch k sequences at a time through re-used output buffers reproduces the whole-batch call bit for bit (the
in-kernel Philox is keyed by the global sequence index), and accumulate_grad leaves the whole-batch gradient in .grad.
Replaces the O(T N B Dx) TensorArrays the reference keeps alive (src/SMC/SVO.py:106-108)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _build(cls_name, **over):
    from psvo_b200.flags import default_flags
    from psvo_b200.model import SSM
    from psvo_b200.SMC.AESMC import AESMC
    from psvo_b200.SMC.PSVO import PSVO
    torch.manual_seed(0)
    F = default_flags(Dx=3, Dy=2, n_particles=96, n_particles_for_BSim_proposal=5, q0_layers="16", q1_layers="16", q2_layers="16",
                      f_layers="16", g_layers="16", y_smoother_Dhs="8", X0_smoother_Dhs="8", seed=3,
                      AESMC=cls_name == "AESMC", PSVO=cls_name == "PSVO", **over)
    model = SSM(F)
    smc = (AESMC if cls_name == "AESMC" else PSVO)(model, F)
    return F, model, smc


@pytest.mark.parametrize("cls_name", ["AESMC", "PSVO"])
def test_chunked_call_equals_whole_batch(cls_name):
    F, model, smc = _build(cls_name)
    g = torch.Generator().manual_seed(1)
    obs = torch.randn(7, 9, 2, generator=g).to(DEV)
    with torch.no_grad():
        whole, logw = smc.get_log_ZSMC(obs, None, seed=11)
        for chunk in (1, 3, 4):
            z, log = smc.get_log_ZSMC(obs, None, seed=11, chunk=chunk)
            assert float(z) == float(whole), (chunk, float(z), float(whole))
            assert torch.equal(log["Xs"], logw["Xs"])
        z, log = smc.get_log_ZSMC(obs, None, seed=11, chunk=3, return_Xs=False)
        assert float(z) == float(whole) and log["Xs"] is None
    assert smc.seq_offset == 0


@pytest.mark.parametrize("cls_name", ["AESMC", "PSVO"])
def test_accumulate_grad_equals_whole_batch_gradient(cls_name):
    F, model, smc = _build(cls_name)
    g = torch.Generator().manual_seed(2)
    obs = torch.randn(6, 8, 2, generator=g).to(DEV)
    with torch.no_grad():
        smc.get_log_ZSMC(obs[:1], None, seed=1)                    # lazy build
    for p in model.parameters():
        p.grad = None
    z, _ = smc.get_log_ZSMC(obs, None, seed=21)
    (-z).backward()
    ref = {n: p.grad.clone() for n, p in model.named_parameters() if p.grad is not None}
    for p in model.parameters():
        p.grad = None
    zc, log = smc.get_log_ZSMC(obs, None, seed=21, chunk=4, return_Xs=False, accumulate_grad=True)
    assert not zc.requires_grad and abs(float(zc) - float(z)) <= 1e-6 * abs(float(z))
    got = {n: p.grad for n, p in model.named_parameters() if p.grad is not None}
    assert set(got) == set(ref)
    for n in ref:
        scale = float(ref[n].abs().max().clamp_min(1e-6))
        assert float((got[n] - ref[n]).abs().max()) <= 2e-5 * scale, n


def test_differentiable_chunked_call_keeps_graph():
    F, model, smc = _build("AESMC")
    g = torch.Generator().manual_seed(3)
    obs = torch.randn(5, 6, 2, generator=g).to(DEV)
    with torch.no_grad():
        smc.get_log_ZSMC(obs[:1], None, seed=1)
    for p in model.parameters():
        p.grad = None
    z, _ = smc.get_log_ZSMC(obs, None, seed=5)
    (-z).backward()
    ref = [p.grad.clone() for p in model.parameters() if p.grad is not None]
    for p in model.parameters():
        p.grad = None
    zc, _ = smc.get_log_ZSMC(obs, None, seed=5, chunk=2)             # no accumulate_grad: an ordinary differentiable scalar
    assert zc.requires_grad
    (-zc).backward()
    got = [p.grad for p in model.parameters() if p.grad is not None]
    for a, b in zip(got, ref):
        assert float((a - b).abs().max()) <= 2e-5 * float(b.abs().max().clamp_min(1e-6))
