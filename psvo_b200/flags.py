"""FLAGS namespace with the reference's flag names and defaults (src/runner_flag.py:22-153).
`SVO(model, FLAGS)` etc. accept any object with these attributes (argparse.Namespace, tf.app.flags.FLAGS, ...)."""
import types

DEFAULTS = dict(
    # training hyper-parameters (runner_flag.py:22-28)
    Dx=2, Dy=1, n_particles=16, batch_size=1, lr=3e-3, epoch=200, seed=2, time=5,
    # networks (runner_flag.py:53-64) -- comma-separated hidden widths, "" = no hidden layer
    q0_layers="32", q1_layers="32", q2_layers="32", f_layers="32", g_layers="32",
    q0_sigma_init=5, q0_sigma_min=1, q1_sigma_init=5, q1_sigma_min=1, q2_sigma_init=5, q2_sigma_min=1,
    f_sigma_init=5, f_sigma_min=1, g_sigma_init=5, g_sigma_min=1,
    output_cov=False, diag_cov=False,
    # bidirectional RNN encoder (runner_flag.py:75-83)
    y_smoother_Dhs="32", X0_smoother_Dhs="32", X0_use_separate_RNN=True, use_stack_rnn=True,
    # state-space model (runner_flag.py:88-98)
    use_bootstrap=True, q_uses_true_X=False, use_2_q=True, poisson_emission=False,
    # objective (runner_flag.py:103-113)
    PSVO=True, PSVOwR=False, SVO=False, AESMC=False, IWAE=False,
    n_particles_for_BSim_proposal=16, BSim_use_single_RNN=False,
    MSE_steps=30,
    # EXTRA (no reference counterpart, off by default): systematic instead of multinomial resampling in the forward filter
    systematic_resampling=False,
    # trainer (runner_flag.py:118-153)
    early_stop_patience=200, lr_reduce_patience=30, lr_reduce_factor=0.7071067811865475, min_lr=3e-4, print_freq=1,
    save_model=False,
)


def default_flags(**over):
    d = dict(DEFAULTS)
    unknown = set(over) - set(d)
    if unknown:
        raise TypeError("unknown flags: %s" % sorted(unknown))
    d.update(over)
    return types.SimpleNamespace(**d)


def parse_layers(s):
    """'64,64' -> [64, 64]; '' -> [] (the reference's int('') crash at src/model.py:27-31 is not reproduced)."""
    if isinstance(s, (list, tuple)):
        return [int(v) for v in s]
    s = str(s).strip()
    return [int(v) for v in s.split(",")] if s else []
