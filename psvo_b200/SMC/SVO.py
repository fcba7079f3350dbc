"""SVO: forward particle-filter objective log_ZSMC, computed by the fused sm_100a kernels.

Public interface of src/SMC/SVO.py: SVO(model, FLAGS, name), get_log_ZSMC(obs, hidden) -> (log_ZSMC, {"Xs"}),
SMC(hidden, obs), resample_X, get_resample_idx, compute_log_ZSMC, sample_from_2_dist, preprocess_obs,
n_step_prediction, get_nextX; attributes n_particles, resample_particles, smooth_obs.
Eager tensors replace TF placeholders; results are laid out sequence-major (B,T,N,.) -- the reference's
particle-major (T,N,B,.) tensors are available as permuted views from SMC()."""
import torch

from .. import functional as Fn
from .. import ops


class SVO:
    def __init__(self, model, FLAGS, name="log_ZSMC"):
        self.model = model
        self.q0, self.q1, self.q2 = model.q0_dist, model.q1_dist, model.q2_dist
        self.f, self.g = model.f_dist, model.g_dist
        self.n_particles = FLAGS.n_particles
        self.q_uses_true_X = FLAGS.q_uses_true_X
        if self.q_uses_true_X:
            raise NotImplementedError("q_uses_true_X debug proposal (src/SMC/SVO.py:234-241) is not built")
        self.X0_use_separate_RNN = FLAGS.X0_use_separate_RNN
        self.use_stack_rnn = FLAGS.use_stack_rnn
        self.smooth_obs = True
        self.resample_particles = True
        # EXTRA (not in the reference, which only has multinomial resampling): FLAGS.systematic_resampling=True switches the
        # forward filter to systematic resampling (lower-variance ancestry; the objective stays a valid SMC bound)
        self.systematic_resampling = bool(getattr(FLAGS, "systematic_resampling", False))
        self.name = name
        self.seed = int(getattr(FLAGS, "seed", 0))
        self._calls = 0
        self.seq_offset = 0            # global index of this rank's first sequence (multi-GPU sharding)

    # ------------------------------------------------------------------ public API
    def get_log_ZSMC(self, obs, hidden=None, eps=None, uniforms=None, forced_ancestors=None, seed=None, out=None,
                     chunk=None, return_Xs=True, accumulate_grad=False):
        """obs (B,T,Dy) CUDA tensor; hidden is accepted for signature parity (only q_uses_true_X reads it).
        Extra keyword arguments (parity / reproducibility): eps (B,T,N,Dx), uniforms (B,T,N),
        forced_ancestors (B,T,N), seed (Philox key; default: FLAGS.seed advanced per call).
        Memory plan for large batches (the reference keeps O(T N B Dx) TensorArrays alive, SVO.py:106-108; config 5 at
        B = 1024 is ~570 GB of filter outputs): chunk=k streams the batch through the kernels k sequences at a time, re-using
        one set of output buffers; return_Xs=False skips collecting log["Xs"]; accumulate_grad=True back-propagates every
        chunk's share of -log_ZSMC at once (gradients accumulate in .grad, the returned scalar is detached) so that a training
        step never holds more than one chunk.  Results do not depend on the chunking (Philox is keyed by the global sequence).
        Returns (log_ZSMC scalar, differentiable unless accumulate_grad; {"Xs": (B,T,N,Dx) or None})."""
        if chunk is not None and 0 < int(chunk) < obs.shape[0]:
            return self._chunked(obs, int(chunk), return_Xs, accumulate_grad, out, seed,
                                 dict(eps=eps, uniforms=uniforms, forced_ancestors=forced_ancestors))
        logZ_b, log = self._filter_objective(obs, eps=eps, uniforms=uniforms, forced_ancestors=forced_ancestors,
                                             seed=seed, out=out)
        if accumulate_grad:
            (-logZ_b.mean()).backward()
            logZ_b = logZ_b.detach()
        return logZ_b.mean(), (log if return_Xs else {"Xs": None})            # SVO.py:308-309

    def _chunked(self, obs, chunk, return_Xs, accumulate_grad, out, seed, per_seq):
        """get_log_ZSMC over slices of the batch; per_seq: keyword tensors with a leading batch axis (sliced alongside obs)."""
        B = obs.shape[0]
        if seed is None:
            seed = (self.seed << 20) + self._calls
            self._calls += 1
        needs_graph = torch.is_grad_enabled() and not accumulate_grad and any(p.requires_grad for p in self.model.parameters())
        bufs = out if out is not None else {}
        base, zs, xs = self.seq_offset, [], []
        try:
            for i in range(0, B, chunk):
                sl = slice(i, min(B, i + chunk))
                self.seq_offset = base + i
                kw = {k: (None if v is None else v[sl]) for k, v in per_seq.items()}
                # a differentiable result keeps every chunk's tensors alive for autograd: fresh buffers then
                reuse = None if needs_graph else bufs
                logZ_b, log = self._filter_objective(obs[sl], seed=seed, out=reuse, **kw)
                if reuse is not None and isinstance(getattr(self, "last", None), dict):
                    raw = self.last.get("raw") or (self.last.get("filter") if "filter" in self.last else None)
                    if isinstance(raw, dict):
                        bufs.update({k: v for k, v in raw.items() if torch.is_tensor(v)})
                if accumulate_grad:
                    (-logZ_b.sum() / float(B)).backward()
                    logZ_b = logZ_b.detach()
                zs.append(logZ_b.clone())
                if return_Xs:
                    xs.append(log["Xs"].clone() if reuse is not None else log["Xs"])
        finally:
            self.seq_offset = base
        return torch.cat(zs).mean(), {"Xs": torch.cat(xs, 0) if return_Xs else None}

    def _filter_objective(self, obs, eps=None, uniforms=None, forced_ancestors=None, seed=None, out=None):
        """Per-sequence objective logZ_b (B,) (log_ZSMC = its batch mean) and the log dict; the sequence-sharded
        multi-GPU wrapper (psvo_b200.distributed) reduces logZ_b over ranks itself."""
        logZ_b, aux = self._filter(obs, eps, uniforms, forced_ancestors, seed, out)
        self.last = aux
        return logZ_b, {"Xs": aux["X_anc"]}                                   # SVO.py:51-56

    def SMC(self, hidden, obs, q_cov=1.0, **kw):
        """Reference layout: Xs, X_ancestors (T,N,B,Dx), log_Ws (T,N,B) (permuted views, no copy)."""
        _, aux = self._filter(obs, kw.get("eps"), kw.get("uniforms"), kw.get("forced_ancestors"), kw.get("seed"), None)
        self.last = aux
        return aux["Xs"].permute(1, 2, 0, 3), aux["X_anc"].permute(1, 2, 0, 3), aux["log_Ws"].permute(1, 2, 0)

    def _filter(self, obs, eps, u, forced, seed, out):
        if obs.dim() != 3:
            raise ValueError("obs must be (batch_size, time, Dy)")
        self.batch_size, self.time, self.Dx = obs.shape[0], obs.shape[1], self.model.Dx
        pre_X0, pre_obs = self.preprocess_obs(obs)
        self.preprocessed_X0, self.preprocessed_obs = pre_X0, pre_obs
        self._build_nets(pre_X0.shape[-1], pre_obs.shape[-1], obs.device)
        if seed is None:
            seed = (self.seed << 20) + self._calls
            self._calls += 1
        flags = Fn.flag_bits(self.model.use_bootstrap, self.model.use_2_q, self.resample_particles,
                             self.model.poisson_emission, systematic=self.systematic_resampling)
        P = self.model.dist_params()
        return ops.smc_filter(P, flags, obs, pre_obs, pre_X0, self.n_particles, eps=eps, u=u, forced_anc=forced,
                              seed=seed, seq_offset=self.seq_offset, out=out)

    def _build_nets(self, Dx0, Dobs, device):
        m = self.model
        m.q0_tran.build(Dx0, device)
        if m.use_2_q:
            m.q2_tran.build(Dobs, device)
        m.to(device)

    # ------------------------------------------------------------------ helpers with the reference's names
    def preprocess_obs(self, obs):
        """SVO.py:313-331.  Returns (preprocessed_X0 (B,Dx0), preprocessed_obs (B,T,Dobs))."""
        if not self.smooth_obs:
            pre_obs, pre_X0 = obs, obs[:, 0]
        else:
            pre_X0, pre_obs = self.preprocess_obs_w_bRNN(obs)
        if not (self.model.use_bootstrap and self.model.use_2_q):
            pre_X0 = self.model.X0_transformer.transform(pre_X0.contiguous())[0]
        return pre_X0.contiguous(), pre_obs.contiguous()

    def preprocess_obs_w_bRNN(self, obs):
        """SVO.py:333-369 (including the quirk that with use_stack_rnn the X0 feature is 4*Dh wide)."""
        y_smoother, X0_smoother = self.model.bRNN
        y_smoother.to(obs.device)
        out_fw, out_bw = y_smoother(obs)
        pre_obs = out_fw if self.use_stack_rnn else torch.cat([out_fw, out_bw], -1)
        if self.X0_use_separate_RNN:
            X0_smoother.to(obs.device)
            out_fw, out_bw = X0_smoother(obs)
        pre_X0 = torch.cat([out_fw[:, -1], out_bw[:, 0]], -1)
        return pre_X0, pre_obs

    def sample_from_2_dist(self, dist1, dist2, d1_input, d2_input, sample_size=()):
        """Diagonal branch of SVO.py:182-197,228-232 as a stand-alone (non-fused) helper; differentiable like the original."""
        d1, d2 = dist1.get_mvn(d1_input), dist2.get_mvn(d2_input)
        i1, i2 = 1.0 / d1.scale_diag, 1.0 / d2.scale_diag
        cov = 1.0 / (i1 + i2)
        mean = cov * (i1 * d1.loc + i2 * d2.loc)
        ss = () if sample_size == () else (sample_size,)
        X = mean + cov * torch.randn(ss + tuple(mean.shape), device=mean.device)
        q_lp = Fn.mvn_log_prob_ad(X, mean.expand_as(X), cov)
        f_lp = Fn.mvn_log_prob_ad(X, d1.loc.expand_as(X), d1.scale_diag)
        return X, q_lp, f_lp

    def get_resample_idx(self, log_W, sample_size=(), uniforms=None):
        """SVO.py:266-300: log_W (K, batch...) -> ancestor indices (sample_size, batch...) (int32)."""
        K, batch = log_W.shape[0], tuple(log_W.shape[1:])
        n = 1 if sample_size == () else int(sample_size)
        logits = log_W.movedim(0, -1).contiguous()
        if uniforms is None:
            uniforms = torch.rand(batch + (n,), device=log_W.device)
        idx = Fn.resample(logits, uniforms)                                   # (batch..., n)
        idx = idx.movedim(-1, 0)
        return idx[0] if sample_size == () else idx

    def resample_X(self, X, log_W, sample_size=(), resample_particles=True, uniforms=None):
        """SVO.py:243-264: X (K, batch..., D) or a list of such."""
        if not resample_particles:
            return X
        if log_W.shape[0] == 1:
            assert sample_size == 1
            return X
        idx = self.get_resample_idx(log_W, sample_size, uniforms)
        idx_b = idx if sample_size != () else idx[None]                       # (n, batch...)

        def one(item):
            extra = item.dim() - log_W.dim()
            src = item if extra == 1 else item[..., None]
            src = src.movedim(0, -2).contiguous()                              # (batch..., K, D)
            got = Fn.gather(src, idx_b.movedim(0, -1).contiguous())            # (batch..., n, D)
            got = got.movedim(-2, 0)
            got = got if extra == 1 else got[..., 0]
            return got if sample_size != () else got[0]
        return [one(i) for i in X] if isinstance(X, list) else one(X)

    @staticmethod
    def compute_log_ZSMC(log_Ws):
        """SVO.py:302-311: log_Ws (time, n_particles, batch_size) -> scalar."""
        T, N, B = log_Ws.shape
        lse = Fn.logsumexp_ad(log_Ws.permute(0, 2, 1).contiguous())            # (T,B)
        return lse.sum(0).mean()

    def n_step_prediction(self, n_steps, hidden, obs):
        """SVO.py:371-399: k-step-ahead predictions from the particle mean.  hidden (B,T,N,Dx), obs (B,T,Dy)."""
        B, T = hidden.shape[0], hidden.shape[1]
        assert n_steps < T, "n_steps = {} >= time".format(n_steps)
        x = hidden.mean(2)
        y_hat = []
        with torch.no_grad():
            for _ in range(n_steps):
                y_hat.append(self.g.mean(x.contiguous()))
                x = self.f.mean(x[:, :-1].contiguous())
            y_hat.append(self.g.mean(x.contiguous()))
        return y_hat, [obs[:, k:] for k in range(n_steps + 1)]

    def get_nextX(self, X):
        with torch.no_grad():
            return self.f.mean(X.contiguous())
