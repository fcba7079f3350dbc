"""PSVO: particle-smoothing objective -- forward filter + FFBSi-style backward simulation with an M-sample proposal.

Interface of src/SMC/PSVO.py: PSVO(model, FLAGS), get_log_ZSMC(obs, hidden) -> (log_ZSMC, {"Xs": bw_Xs (B,T,N,Dx)}),
backward_simulation_w_proposal(Xs, log_Ws, obs), compute_log_ZSMC(f, g, Omega), BS_preprocess_obs(obs).
The objective is differentiable end to end: psvo_ffbsi_backward chained into psvo_smc_backward (ops.SmoothingFn),
with the bi-LSTM features and the MLP parameters receiving gradients exactly where TF autodiff would put them."""
import torch

from .. import functional as Fn
from .. import ops
from .SVO import SVO


class PSVO(SVO):
    with_resampling = False

    def __init__(self, model, FLAGS, name="log_ZSMC"):
        SVO.__init__(self, model, FLAGS, name="log_ZSMC")
        self.n_particles_for_BSim_proposal = FLAGS.n_particles_for_BSim_proposal
        self.smooth_obs = False
        self.BSim_use_single_RNN = FLAGS.BSim_use_single_RNN
        if self.BSim_use_single_RNN:
            raise NotImplementedError("BSim_use_single_RNN=True is broken in the reference (self.y_smoother_f is unset "
                                      "when smooth_obs=False, src/SMC/PSVO.py:208-212)")
        self.q1_inv = model.q1_inv_dist
        self.BSim_q_init = model.Bsim_q_init_dist
        self.BSim_q2 = model.BSim_q2_dist

    def get_log_ZSMC(self, obs, hidden=None, eps=None, uniforms=None, forced_ancestors=None, seed=None,
                     eps_bw=None, uniforms_pick=None, uniforms_res=None, forced_pick=None, forced_anc2=None, out=None,
                     chunk=None, return_Xs=True, accumulate_grad=False):
        """chunk / return_Xs / accumulate_grad: the batch memory plan of SVO.get_log_ZSMC"""
        per_seq = dict(eps=eps, uniforms=uniforms, forced_ancestors=forced_ancestors, eps_bw=eps_bw, uniforms_pick=uniforms_pick,
                       uniforms_res=uniforms_res, forced_pick=forced_pick, forced_anc2=forced_anc2)
        if chunk is not None and 0 < int(chunk) < obs.shape[0]:
            return self._chunked(obs, int(chunk), return_Xs, accumulate_grad, out, seed, per_seq)
        logZ_b, log = self._filter_objective(obs, seed=seed, out=out, **per_seq)
        if accumulate_grad:
            (-logZ_b.mean()).backward()
            logZ_b = logZ_b.detach()
        return logZ_b.mean(), (log if return_Xs else {"Xs": None})                        # PSVO.py:65

    def _filter_objective(self, obs, eps=None, uniforms=None, forced_ancestors=None, seed=None, eps_bw=None,
                          uniforms_pick=None, uniforms_res=None, forced_pick=None, forced_anc2=None, out=None):
        """Per-sequence smoothing objective logZ_b (B,) and the log dict (PSVO.py:39-48)."""
        if obs.dim() != 3:
            raise ValueError("obs must be (batch_size, time, Dy)")
        m = self.model
        self.batch_size, self.time, self.Dx = obs.shape[0], obs.shape[1], m.Dx
        pre_X0, pre_obs = self.preprocess_obs(obs)                                        # smooth_obs=False: raw obs
        self.preprocessed_X0, self.preprocessed_obs = pre_X0, pre_obs
        self._build_nets(pre_X0.shape[-1], pre_obs.shape[-1], obs.device)
        _, feats = self.BS_preprocess_obs(obs)                                            # PSVO.py:80
        feats = feats.contiguous()
        m.BSim_q_init_tran.build(feats.shape[-1], obs.device)
        m.BSim_q2_tran.build(feats.shape[-1], obs.device)
        m.to(obs.device)
        if seed is None:
            seed = (self.seed << 20) + self._calls
            self._calls += 1
        self._last_seed = seed
        flags = Fn.flag_bits(m.use_bootstrap, m.use_2_q, True, m.poisson_emission, self.with_resampling)
        logZ_b, Xout, last = ops.smoothing_objective(
            m.dist_params(), flags, obs, pre_obs, pre_X0, feats, self.n_particles, self.n_particles_for_BSim_proposal,
            eps=eps, u=uniforms, forced_anc=forced_ancestors, eps_bw=eps_bw, u_pick=uniforms_pick, u_res=uniforms_res,
            forced_pick=forced_pick, forced_anc2=forced_anc2, seed=seed, seq_offset=self.seq_offset, out=out)
        self.last = last
        return logZ_b, {"Xs": Xout}                                                       # PSVO.py:43, PSVOwR.py:43

    def _filter(self, obs, eps, u, forced, seed, out):
        if seed is None:
            seed = (self.seed << 20) + self._calls
            self._calls += 1
        self._last_seed = seed
        return SVO._filter(self, obs, eps, u, forced, seed, out)

    def BS_preprocess_obs(self, obs):
        """PSVO.py:205-216: bi-LSTM features for the backward proposal (the X0 feature is discarded)."""
        return self.preprocess_obs_w_bRNN(obs)

    def backward_simulation_w_proposal(self, Xs, log_Ws, obs, lse=None, eps_bw=None, uniforms_pick=None,
                                       uniforms_res=None, forced_pick=None, forced_anc2=None, seed=0):
        """Xs (B,T,N,Dx) pre-resampling particles, log_Ws (B,T,N) from the forward filter.  Returns the dict of
        psvo_ffbsi_forward outputs: bw_Xs, pick, logZ_b and f_lp/g_lp/Omega (PSVO) or bw_log_W/bw_X_anc/anc2 (PSVOwR)."""
        m = self.model
        if lse is None:
            lse = Fn.logsumexp(log_Ws)
        _, feats = self.BS_preprocess_obs(obs)
        feats = feats.contiguous()
        m.BSim_q_init_tran.build(feats.shape[-1], obs.device)
        m.BSim_q2_tran.build(feats.shape[-1], obs.device)
        m.to(obs.device)
        flags = Fn.flag_bits(m.use_bootstrap, m.use_2_q, True, m.poisson_emission, self.with_resampling)
        return Fn.ffbsi_forward_raw(m.dist_params(), flags, obs, feats.detach(), self.preprocessed_X0.detach(), Xs, log_Ws,
                                    lse, self.n_particles_for_BSim_proposal, eps_bw=eps_bw, u_pick=uniforms_pick,
                                    u_res=uniforms_res, forced_pick=forced_pick, forced_anc2=forced_anc2, seed=seed,
                                    seq_offset=self.seq_offset)

    @staticmethod
    def compute_log_ZSMC(f_log_probs, g_log_probs, bw_log_Omegas):
        """PSVO.py:52-67 on (time, n_particles, batch_size) tensors."""
        T, N, B = f_log_probs.shape
        joint = (f_log_probs + g_log_probs).sum(0)
        prop = bw_log_Omegas.sum(0)
        lse = Fn.logsumexp_ad((joint - prop).t().contiguous())
        return (lse - torch.log(torch.tensor(float(N), device=lse.device))).mean()
