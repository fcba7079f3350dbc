"""Bidirectional LSTM observation encoder: the step just BEFORE the hot path.
Built on torch.nn.LSTM (cuDNN); its outputs are the kernel's boundary inputs pre_obs / pre_x0 and it receives
their gradients from psvo_smc_backward.

Reference: SSM.init_RNNs (src/model.py:162-186) + SVO.preprocess_obs_w_bRNN (src/SMC/SVO.py:333-369).
  use_stack_rnn=True  : tf.contrib.rnn.stack_bidirectional_dynamic_rnn -- layer k reads concat(fw,bw) of layer k-1
  use_stack_rnn=False : MultiRNNCell per direction + tf.nn.bidirectional_dynamic_rnn
LSTMBlockCell has forget_bias=1: reproduced by initialising the forget-gate bias to 1."""
import torch


def _lstm(d_in, d_h):
    m = torch.nn.LSTM(d_in, d_h, batch_first=True)
    with torch.no_grad():
        m.bias_ih_l0.zero_(); m.bias_hh_l0.zero_()
        m.bias_ih_l0[d_h:2 * d_h].fill_(1.0)          # torch gate order i,f,g,o
    return m


def load_tf_lstm_cell(lstm, kernel, bias, forget_bias=1.0):
    """Copy a tf.contrib.rnn.LSTMBlockCell into a one-layer torch.nn.LSTM.
    TF: kernel [d_in + units, 4 units] applied to concat(x, h), gate order (i, c~, f, o), forget_bias added to f at run time.
    torch: weight_ih [4 units, d_in], weight_hh [4 units, units], gate order (i, f, g = c~, o), biases b_ih + b_hh."""
    kernel = torch.as_tensor(kernel)
    bias = torch.as_tensor(bias)
    u = lstm.hidden_size
    d_in = kernel.shape[0] - u
    assert d_in == lstm.input_size and kernel.shape[1] == 4 * u and bias.shape[0] == 4 * u
    i, c, f, o = torch.chunk(kernel, 4, dim=1)
    bi, bc, bf, bo = torch.chunk(bias, 4, dim=0)
    W = torch.cat([i, f, c, o], 1)                       # torch order
    b = torch.cat([bi, bf + forget_bias, bc, bo], 0)
    with torch.no_grad():
        lstm.weight_ih_l0.copy_(W[:d_in].t().to(lstm.weight_ih_l0.dtype))
        lstm.weight_hh_l0.copy_(W[d_in:].t().to(lstm.weight_hh_l0.dtype))
        lstm.bias_ih_l0.copy_(b.to(lstm.bias_ih_l0.dtype))
        lstm.bias_hh_l0.zero_()
    return lstm


class BiRNN(torch.nn.Module):
    def __init__(self, d_in, Dhs, stack=True):
        super().__init__()
        self.stack, self.Dhs = stack, list(Dhs)
        self.fw, self.bw = torch.nn.ModuleList(), torch.nn.ModuleList()
        d = d_in
        for h in self.Dhs:
            self.fw.append(_lstm(d, h))
            self.bw.append(_lstm(d, h))
            d = 2 * h if stack else h
        self.out_width = 2 * self.Dhs[-1]

    def load_tf_cells(self, cells_fw, cells_bw, forget_bias=1.0):
        """cells_*: per layer (kernel, bias) of the reference's LSTMBlockCells (src/model.py:164-179)."""
        assert len(cells_fw) == len(self.fw) and len(cells_bw) == len(self.bw)
        for m, (k, b) in zip(self.fw, cells_fw):
            load_tf_lstm_cell(m, k, b, forget_bias)
        for m, (k, b) in zip(self.bw, cells_bw):
            load_tf_lstm_cell(m, k, b, forget_bias)
        return self

    def forward(self, obs):
        """obs (B,T,D) -> (outputs_fw, outputs_bw) each (B,T,.) as the reference's `outputs` unpacks
        (SVO.py:362-366): with stacking both are the concatenated (B,T,2Dh) tensor."""
        if self.stack:
            x = obs
            for f, b in zip(self.fw, self.bw):
                of, _ = f(x)
                ob, _ = b(torch.flip(x, [1]))
                x = torch.cat([of, torch.flip(ob, [1])], -1)
            return x, x
        xf, xb = obs, torch.flip(obs, [1])
        for f, b in zip(self.fw, self.bw):
            xf, _ = f(xf)
            xb, _ = b(xb)
        return xf, torch.flip(xb, [1])
