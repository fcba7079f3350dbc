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
