"""Eager training loop around the fused SMC objective -- the caller of the hot path.

Mirrors src/trainer.py (class trainer: train / evaluate_and_save_metrics / adjust_lr / evaluate_R_square) without the
TF1 session: the objective object (SVO / PSVO / PSVOwR / AESMC / IWAE) is called per mini-batch, -log_ZSMC is
back-propagated through psvo_smc_backward / psvo_ffbsi_backward, Adam uses TensorFlow's defaults
(beta1 0.9, beta2 0.999, eps 1e-8; src/trainer.py:117).  Plotting and TensorBoard are out of scope; metrics can be
written as the reference's per-epoch pickles (metric_{epoch}.p) and history.json."""
import json
import math
import os
import pickle
import time

import numpy as np
import torch


class StopTraining(Exception):
    pass


class trainer:
    def __init__(self, model, SMC, FLAGS, device="cuda"):
        self.model, self.SMC, self.FLAGS = model, SMC, FLAGS
        self.device = torch.device(device)
        self.Dx, self.Dy, self.n_particles = FLAGS.Dx, FLAGS.Dy, FLAGS.n_particles
        self.MSE_steps = getattr(FLAGS, "MSE_steps", 30)
        self.batch_size = FLAGS.batch_size
        self.lr = getattr(FLAGS, "lr", 3e-3)
        self.epoch = getattr(FLAGS, "epoch", 200)
        self.early_stop_patience = getattr(FLAGS, "early_stop_patience", 200)      # src/runner_flag.py:118-127
        self.lr_reduce_factor = getattr(FLAGS, "lr_reduce_factor", 1 / math.sqrt(2))
        self.lr_reduce_patience = getattr(FLAGS, "lr_reduce_patience", 30)
        self.min_lr = getattr(FLAGS, "min_lr", self.lr / 10)
        self.bestCost, self.early_stop_count, self.lr_reduce_count = 0, 0, 0
        self.log_ZSMC_trains, self.log_ZSMC_tests, self.R_square_trains, self.R_square_tests = [], [], [], []
        self.save_res, self.RLT_DIR = False, None
        self.seed = int(getattr(FLAGS, "seed", 0))
        self.optimizer = None

    def init_data_saving(self, RLT_DIR):
        self.save_res, self.RLT_DIR = True, RLT_DIR
        os.makedirs(RLT_DIR, exist_ok=True)

    # ------------------------------------------------------------------ training
    def _to_dev(self, a):
        return torch.as_tensor(np.asarray(a), dtype=torch.float32).to(self.device)

    def train(self, obs_train, obs_test, hidden_train=None, hidden_test=None, print_freq=10):
        """Returns (metrics dict, log dict) like src/trainer.py:100-195.  obs_* are (n, time, Dy) arrays."""
        self.obs_train, self.obs_test = self._to_dev(obs_train), self._to_dev(obs_test)
        self.MSE_steps = min(self.MSE_steps, self.obs_train.shape[1] - 1)          # src/runner.py:58
        self._lazy_build()
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=self.lr, betas=(0.9, 0.999), eps=1e-8)
        gen = torch.Generator().manual_seed(self.seed)
        log = {}
        for i in range(self.epoch):
            start = time.time()
            if i == 0:
                self.evaluate_and_save_metrics(i)
            perm = torch.randperm(self.obs_train.shape[0], generator=gen).to(self.device)     # sklearn shuffle, trainer.py:146
            for j in range(0, len(perm), self.batch_size):
                batch = self.obs_train[perm[j:j + self.batch_size]].contiguous()
                for group in self.optimizer.param_groups:
                    group["lr"] = self.lr
                self.optimizer.zero_grad(set_to_none=True)
                log_ZSMC, log = self.SMC.get_log_ZSMC(batch, None)
                (-log_ZSMC).backward()                                                         # trainer.py:118
                self.optimizer.step()
            if (i + 1) % print_freq == 0:
                try:
                    self.evaluate_and_save_metrics(i)
                    self.adjust_lr(i, print_freq)
                except StopTraining:
                    break
            print("epoch {:<4} took {:.3f} seconds".format(i + 1, time.time() - start))
        metrics = {"log_ZSMC_trains": self.log_ZSMC_trains, "log_ZSMC_tests": self.log_ZSMC_tests,
                   "R_square_trains": self.R_square_trains, "R_square_tests": self.R_square_tests}
        if self.save_res:
            with open(os.path.join(self.RLT_DIR, "history.json"), "w") as f:
                json.dump({k: [np.asarray(v).tolist() for v in vs] for k, vs in metrics.items()}, f, indent=2)
        return metrics, log

    def _lazy_build(self):
        """Keras infers Dense input widths at the first call; do one forward pass so every parameter exists before
        the optimizer is created."""
        with torch.no_grad():
            self.SMC.get_log_ZSMC(self.obs_train[:self.batch_size].contiguous(), None)

    # ------------------------------------------------------------------ evaluation (trainer.py:200-242, 287-335)
    def evaluate(self, obs):
        """log_ZSMC of every mini-batch and the k-step predictions over the whole set."""
        vals, y_hats, ys = [], None, None
        with torch.no_grad():
            for i in range(0, obs.shape[0], self.batch_size):
                batch = obs[i:i + self.batch_size].contiguous()
                log_ZSMC, log = self.SMC.get_log_ZSMC(batch, None)
                vals.append(float(log_ZSMC))
                y_hat, y = self.SMC.n_step_prediction(self.MSE_steps, log["Xs"], batch)
                if y_hats is None:
                    y_hats, ys = [[a.cpu().numpy()] for a in y_hat], [[a.cpu().numpy()] for a in y]
                else:
                    for k in range(len(y_hat)):
                        y_hats[k].append(y_hat[k].cpu().numpy()); ys[k].append(y[k].cpu().numpy())
        return np.array(vals), [np.concatenate(a) for a in y_hats], [np.concatenate(a) for a in ys]

    def evaluate_and_save_metrics(self, iter_num):
        lz_tr, yh_tr, y_tr = self.evaluate(self.obs_train)
        lz_te, yh_te, y_te = self.evaluate(self.obs_test)
        log_ZSMC_train, log_ZSMC_test = float(np.mean(lz_tr)), float(np.mean(lz_te))
        R_square_train, R_square_test = self.evaluate_R_square(yh_tr, y_tr), self.evaluate_R_square(yh_te, y_te)
        print("\niter", iter_num + 1)
        print("Train log_ZSMC: {:>7.3f}, valid log_ZSMC: {:>7.3f}".format(log_ZSMC_train, log_ZSMC_test))
        print("Train, Valid k-step Rsq:\n", R_square_train, "\n", R_square_test)
        if not math.isfinite(log_ZSMC_train):                                                  # trainer.py:221-223
            print("Nan in log_ZSMC, stop training")
            raise StopTraining()
        self.log_ZSMC_trains.append(log_ZSMC_train); self.log_ZSMC_tests.append(log_ZSMC_test)
        self.R_square_trains.append(R_square_train); self.R_square_tests.append(R_square_test)
        if self.save_res:
            with open(os.path.join(self.RLT_DIR, "metric_{}.p".format(iter_num + 1)), "wb") as f:
                pickle.dump({"log_ZSMC_train": log_ZSMC_train, "log_ZSMC_test": log_ZSMC_test,
                             "R_square_train": R_square_train, "R_square_test": R_square_test}, f)
        return log_ZSMC_train, log_ZSMC_test, R_square_train, R_square_test

    def adjust_lr(self, iter_num, print_freq):
        """Early stopping and learning-rate decay on the held-out log_ZSMC (trainer.py:244-270)."""
        best = int(np.argmax(self.log_ZSMC_tests))
        if self.bestCost != best:
            self.early_stop_count, self.lr_reduce_count, self.bestCost = 0, 0, best
        print("best valid cost on iter: {}\n".format(self.bestCost * print_freq))
        if self.bestCost != len(self.log_ZSMC_tests) - 1:
            self.early_stop_count += 1
            if self.early_stop_count * print_freq == self.early_stop_patience:
                print("valid cost not improving. stopping training...")
                raise StopTraining()
            self.lr_reduce_count += 1
            if self.lr_reduce_count * print_freq == self.lr_reduce_patience:
                self.lr_reduce_count = 0
                self.lr = max(self.lr * self.lr_reduce_factor, self.min_lr)
                print("valid cost not improving. reduce learning rate to {}".format(self.lr))
        elif self.save_res and getattr(self.FLAGS, "save_model", False):
            torch.save(self.model.state_dict(), os.path.join(self.RLT_DIR, "model_epoch_{}.pt".format(iter_num + 1)))

    @staticmethod
    def evaluate_R_square(y_hat, y):
        """k-step R^2 = 1 - sum (y_hat - y)^2 / sum (y - mean_over_sequences y)^2   (trainer.py:322-335)."""
        R_square = np.zeros(len(y_hat))
        for i, (y_hat_i, y_i) in enumerate(zip(y_hat, y)):
            MSE = np.sum((y_hat_i - y_i) ** 2)
            y_i_var = np.sum((y_i - np.mean(y_i, axis=0, keepdims=True)) ** 2)
            R_square[i] = 1 - MSE / y_i_var
        return R_square
