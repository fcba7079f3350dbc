"""CPU test of SURVEY 8a row a12 / 8f row f2: psvo_b200.encoder.BiRNN (torch.nn.LSTM), loaded with the LSTMBlockCell kernels and
biases the reference's own code ran with (exported into the golden vectors by tests/golden/make_golden.py), reproduces the
features the reference fed to the hot path: pre_obs / pre_x0 of SVO.preprocess_obs_w_bRNN (src/SMC/SVO.py:333-369) and bs_feat
of PSVO.BS_preprocess_obs (src/SMC/PSVO.py:205-216).  Pins the gate order (i, c~, f, o -> torch's i, f, g, o), forget_bias = 1,
stacked vs MultiRNNCell wiring (src/model.py:162-186), the shared-vs-separate X0 RNN and the 4*Dh X0 feature of the stacked form."""
import numpy as np
import pytest
import torch

from golden_io import case_names, load_case

RNN_CASES = [c for c in case_names() if c.startswith("svo") or "psvo" in c or c.startswith("c1_svo")]


def _cells(z, key):
    out, i = [], 0
    while "rnn.%s.%d.kernel" % (key, i) in z:
        assert float(z["rnn.%s.%d.forget_bias" % (key, i)]) == 1.0
        out.append((z["rnn.%s.%d.kernel" % (key, i)], z["rnn.%s.%d.bias" % (key, i)]))
        i += 1
    return out


@pytest.mark.parametrize("name", RNN_CASES)
def test_encoder_reproduces_reference_features(name):
    from psvo_b200.flags import default_flags
    from psvo_b200.model import SSM
    from psvo_b200.SMC.SVO import SVO
    from psvo_b200.SMC.PSVO import PSVO
    z, P, F, d = load_case(name, dtype=torch.float32)
    cls = str(z["cls"])
    Dx, Dy = int(z["Xs"].shape[-1]), int(z["obs"].shape[-1])
    FL = default_flags(Dx=Dx, Dy=Dy, n_particles=int(z["N"]), y_smoother_Dhs=str(z["y_smoother_Dhs"]),
                       X0_smoother_Dhs=str(z["X0_smoother_Dhs"]), use_stack_rnn=bool(z["use_stack_rnn"]),
                       X0_use_separate_RNN=bool(z["X0_use_separate_RNN"]), SVO=cls == "SVO", PSVO=cls == "PSVO",
                       PSVOwR=cls == "PSVOwR", AESMC=False)
    model = SSM(FL)
    y_smoother, x0_smoother = model.bRNN
    y_smoother.load_tf_cells(_cells(z, "y_f"), _cells(z, "y_b"))
    if x0_smoother is not None and _cells(z, "x0_f"):      # PSVO never runs its X0 RNN (smooth_obs=False): cells stay unbuilt
        x0_smoother.load_tf_cells(_cells(z, "x0_f"), _cells(z, "x0_b"))
    smc = (SVO if cls == "SVO" else PSVO)(model, FL)
    obs = d["obs"]
    with torch.no_grad():
        pre_x0, pre_obs = smc.preprocess_obs_w_bRNN(obs)
    if cls == "SVO":
        np.testing.assert_allclose(pre_obs.numpy(), z["pre_obs"], rtol=2e-5, atol=2e-6)
        np.testing.assert_allclose(pre_x0.numpy(), z["pre_x0"], rtol=2e-5, atol=2e-6)
        width = 2 * int(str(z["X0_smoother_Dhs"]).split(",")[-1])
        if bool(z["use_stack_rnn"]):
            assert pre_x0.shape[-1] == 2 * width          # the 4*Dh quirk (SVO.py:362-367: outputs_fw = outputs_bw = outputs)
        else:
            assert pre_x0.shape[-1] == width
    else:
        np.testing.assert_allclose(pre_obs.numpy(), z["bs_feat"], rtol=2e-5, atol=2e-6)


def test_load_tf_lstm_cell_against_explicit_recurrence():
    """the mapping itself, on a random cell: torch.nn.LSTM loaded from a TF-layout kernel == LSTMBlockCell's published recurrence"""
    from psvo_b200.encoder import load_tf_lstm_cell
    g = torch.Generator().manual_seed(0)
    d_in, u, B, T = 3, 5, 2, 7
    kernel = torch.randn(d_in + u, 4 * u, generator=g, dtype=torch.float64) * 0.4
    bias = torch.randn(4 * u, generator=g, dtype=torch.float64) * 0.3
    x = torch.randn(B, T, d_in, generator=g, dtype=torch.float64)
    lstm = load_tf_lstm_cell(torch.nn.LSTM(d_in, u, batch_first=True).double(), kernel, bias, forget_bias=1.0)
    got, _ = lstm(x)
    c = torch.zeros(B, u, dtype=torch.float64); h = torch.zeros(B, u, dtype=torch.float64)
    outs = []
    for t in range(T):
        zt = torch.cat([x[:, t], h], -1) @ kernel + bias
        i, ci, f, o = torch.chunk(zt, 4, -1)
        c = torch.sigmoid(f + 1.0) * c + torch.sigmoid(i) * torch.tanh(ci)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    np.testing.assert_allclose(got.detach().numpy(), torch.stack(outs, 1).numpy(), rtol=1e-12, atol=1e-12)
