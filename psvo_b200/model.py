"""SSM: container of the state-space model's networks and distributions, with the attribute names the SMC
objective classes read (src/model.py:17-61,67-192): q0/q1/q2/f/g_tran and _dist, Bsim_q_init_dist, q1_inv_dist,
BSim_q2_dist, bRNN, X0_transformer, Dx, Dy, use_bootstrap, use_2_q, poisson_emission.
Placeholders (src/model.py:63-65) have no eager equivalent: get_log_ZSMC takes tensors."""
import torch

from .distribution.mvn import tf_mvn
from .distribution.poisson import tf_poisson
from .encoder import BiRNN
from .flags import parse_layers
from .transformation.MLP import MLP_transformation


class SSM(torch.nn.Module):
    def __init__(self, FLAGS):
        super().__init__()
        F = FLAGS
        self.Dx, self.Dy = F.Dx, F.Dy
        self.time, self.batch_size = getattr(F, "time", None), getattr(F, "batch_size", None)
        self.q0_layers, self.q1_layers, self.q2_layers = parse_layers(F.q0_layers), parse_layers(F.q1_layers), parse_layers(F.q2_layers)
        self.f_layers, self.g_layers = parse_layers(F.f_layers), parse_layers(F.g_layers)
        self.q0_sigma_init, self.q0_sigma_min = F.q0_sigma_init, F.q0_sigma_min
        self.q1_sigma_init, self.q1_sigma_min = F.q1_sigma_init, F.q1_sigma_min
        self.q2_sigma_init, self.q2_sigma_min = F.q2_sigma_init, F.q2_sigma_min
        self.f_sigma_init, self.f_sigma_min = F.f_sigma_init, F.f_sigma_min
        self.g_sigma_init, self.g_sigma_min = F.f_sigma_init, F.g_sigma_min      # sic: src/model.py:37
        self.y_smoother_Dhs, self.X0_smoother_Dhs = parse_layers(F.y_smoother_Dhs), parse_layers(F.X0_smoother_Dhs)
        self.output_cov, self.diag_cov = F.output_cov, F.diag_cov
        self.use_bootstrap, self.use_2_q, self.poisson_emission = F.use_bootstrap, F.use_2_q, F.poisson_emission
        self.X0_use_separate_RNN, self.use_stack_rnn = F.X0_use_separate_RNN, F.use_stack_rnn
        self.PSVO, self.PSVOwR, self.SVO = F.PSVO, F.PSVOwR, F.SVO
        self.BSim_use_single_RNN = F.BSim_use_single_RNN
        self.init_trans()
        self.init_dist()
        self.init_RNNs()

    def _mlp(self, layers, Dout, name, Din=None):
        return MLP_transformation(layers, Dout, output_cov=self.output_cov, diag_cov=self.diag_cov, name=name, Din=Din)

    def init_trans(self):                                                     # src/model.py:67-111
        self.q0_tran = self._mlp(self.q0_layers, self.Dx, "q0_tran")
        self.q1_tran = self._mlp(self.q1_layers, self.Dx, "q1_tran", Din=self.Dx)
        self.q2_tran = self._mlp(self.q2_layers, self.Dx, "q2_tran") if self.use_2_q else None
        if self.PSVO or self.PSVOwR:
            self.BSim_q_init_tran = self._mlp(self.q0_layers, self.Dx, "BSim_q_init_tran")
            self.q1_inv_tran = self._mlp(self.q1_layers, self.Dx, "q1_inv_tran", Din=self.Dx)
            self.BSim_q2_tran = self._mlp(self.q2_layers, self.Dx, "BSim_q2_tran")
        # bootstrap: f IS q1 (same parameters, gradients accumulate into one tensor)
        self.f_tran = self.q1_tran if self.use_bootstrap else self._mlp(self.f_layers, self.Dx, "f_tran", Din=self.Dx)
        self.g_tran = self._mlp(self.g_layers, self.Dy, "g_tran", Din=self.Dx)

    def init_dist(self):                                                      # src/model.py:113-160
        self.q0_dist = tf_mvn(self.q0_tran, self.q0_sigma_init, self.q0_sigma_min, name="q0_dist")
        self.q1_dist = tf_mvn(self.q1_tran, self.q1_sigma_init, self.q1_sigma_min, name="q1_dist")
        self.q2_dist = tf_mvn(self.q2_tran, self.q2_sigma_init, self.q2_sigma_min, name="q2_dist") if self.use_2_q else None
        if self.PSVO or self.PSVOwR:
            self.Bsim_q_init_dist = tf_mvn(self.BSim_q_init_tran, self.q0_sigma_init, self.q0_sigma_min, name="BSim_q_init_dist")
            self.q1_inv_dist = tf_mvn(self.q1_inv_tran, self.q1_sigma_init, self.q1_sigma_min, name="q1_inv_dist")
            self.BSim_q2_dist = tf_mvn(self.BSim_q2_tran, self.q2_sigma_init, self.q2_sigma_min, name="BSim_q2_dist")
        self.f_dist = self.q1_dist if self.use_bootstrap else tf_mvn(self.f_tran, self.f_sigma_init, self.f_sigma_min, name="f_dist")
        if self.poisson_emission:
            self.g_dist = tf_poisson(self.g_tran, name="g_dist")
        else:
            self.g_dist = tf_mvn(self.g_tran, self.g_sigma_init, self.g_sigma_min, name="g_dist")

    def init_RNNs(self):                                                      # src/model.py:162-192
        if self.SVO or self.PSVO or self.PSVOwR:
            self.y_smoother = BiRNN(self.Dy, self.y_smoother_Dhs, stack=self.use_stack_rnn)
            self.X0_smoother = BiRNN(self.Dy, self.X0_smoother_Dhs, stack=self.use_stack_rnn) \
                if self.X0_use_separate_RNN else None
            self.bRNN = (self.y_smoother, self.X0_smoother)
        else:
            self.bRNN = None
        if not (self.use_bootstrap and self.use_2_q):
            self.X0_transformer = MLP_transformation([], self.Dx, name="X0_transformer",
                                                     kernel_initializer="he_uniform")        # linear Dense(Dx)

    # ---- parameters in the layout the fused kernels consume (names of include/psvo.h psvo_params)
    def dist_params(self):
        P = {"q0": self.q0_dist.dist_params(), "q1": self.q1_dist.dist_params(), "g": self.g_dist.dist_params()}
        if self.use_2_q:
            P["q2"] = self.q2_dist.dist_params()
        if not self.use_bootstrap:
            P["f"] = self.f_dist.dist_params()
        if self.PSVO or self.PSVOwR:
            P["bsim_q_init"] = self.Bsim_q_init_dist.dist_params()
            P["q1_inv"] = self.q1_inv_dist.dist_params()
            P["bsim_q2"] = self.BSim_q2_dist.dist_params()
        # lazily built networks (Keras infers the input width at first call) only exist once they have been fed
        return {k: v for k, v in P.items() if len(v.Ws) > 0}
