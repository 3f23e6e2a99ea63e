"""Stand-in for the reference trainer's call sequence (model/traintest_MegaCRN.py:101-155), for machines that do not have
the reference checkout (the GPU test box): the SAME sequence of calls into the drop-in module that ``traintest_model()``
and ``evaluate()`` make --

    model = MegaCRN(...).to(device)                                  traintest:27-31
    optimizer = torch.optim.Adam(model.parameters(), lr, eps)        :104      (plain torch optimiser: in-place updates)
    lr_scheduler = MultiStepLR(optimizer, milestones, gamma)         :105
    per batch:  optimizer.zero_grad(); model(x, ycov, y, batches_seen)          :115-117
                masked MAE on inverse-scaled values + lamb * TripletMarginLoss + lamb1 * MSELoss   :118-125
                loss.item(); loss.backward(); clip_grad_norm_(5); optimizer.step()                  :126-130
    per epoch:  lr_scheduler.step(); evaluate('val'); evaluate('test'); torch.save(state_dict) on improvement   :132-150
    finally:    fresh model, load_state_dict(torch.load(...)), evaluate('test')                     :152-155

-- written against plain PyTorch autograd (``loss.backward()`` through the module's autograd.Function), NOT the fused
``train_step``: it is what the unmodified reference script does to the module.  When the reference checkout is available,
``megacrn_b200.launch_traintest`` runs the reference's own file instead; this module exists so that the same behaviour is
exercised by ``pytest -m gpu`` on a box without it.  Nothing here is reference code: data are synthetic arrays of the
``train/val/test.npz`` schema (x, y float [S, 12, N, 2]: channel 0 speed, channel 1 time of day;
generate_training_data.py:29-53).
"""
from __future__ import annotations

import argparse
import os
import time

import numpy as np
import torch
import torch.nn as nn

from .data import DeviceDataLoader
from .MegaCRN import MegaCRN


def synthetic_npz(num_samples, num_nodes, seq_len=12, horizon=12, seed=0, zero_fraction=0.02):
    """x, y [S, T, N, 2] with the METR-LA schema: speeds with a daily profile (and a few zeros = missing readings, which
    the masked losses ignore), time of day in [0, 1)."""
    rng = np.random.default_rng(seed)
    t0 = rng.integers(0, 288, size=num_samples)
    steps = np.arange(seq_len + horizon)
    tod = ((t0[:, None] + steps[None, :]) % 288) / 288.0                          # [S, T]
    base = 55.0 + 10.0 * np.sin(2 * np.pi * tod)[:, :, None] + 4.0 * rng.standard_normal((1, 1, num_nodes))
    speed = base + 2.0 * rng.standard_normal((num_samples, seq_len + horizon, num_nodes))
    speed[rng.random(speed.shape) < zero_fraction] = 0.0
    full = np.stack([speed, np.broadcast_to(tod[:, :, None], speed.shape)], axis=-1)
    return full[:, :seq_len].copy(), full[:, seq_len:].copy()


def masked_mae(y_pred, y_true):
    """utils.masked_mae_loss semantics (model/utils.py:126-133)."""
    mask = (y_true != 0).float()
    mask = mask / mask.mean()
    loss = torch.abs(y_pred - y_true) * mask
    loss = torch.where(torch.isnan(loss), torch.zeros_like(loss), loss)
    return loss.mean()


class StandInTrainer:
    def __init__(self, num_nodes=207, seq_len=12, horizon=12, rnn_units=64, mem_num=20, mem_dim=64, cheb_k=3, batch_size=64,
                 lr=0.01, eps=1e-3, milestones=(50, 100), gamma=0.1, lamb=0.01, lamb1=0.01, max_grad_norm=5.0,
                 cl_decay_steps=2000, device="cuda:0", n_train=256, n_val=64, n_test=64, seed=0, workdir="."):
        self.device = torch.device(device)
        self.kw = dict(num_nodes=num_nodes, input_dim=1, output_dim=1, horizon=horizon, rnn_units=rnn_units, num_layers=1,
                       mem_num=mem_num, mem_dim=mem_dim, cheb_k=cheb_k, cl_decay_steps=cl_decay_steps,
                       use_curriculum_learning=True)
        self.lamb, self.lamb1, self.max_grad_norm = lamb, lamb1, max_grad_norm
        xs, ys = synthetic_npz(n_train + n_val + n_test, num_nodes, seq_len, horizon, seed)
        self.mean, self.std = float(xs[:n_train, ..., 0].mean()), float(xs[:n_train, ..., 0].std())     # traintest:274
        cut = {"train": (0, n_train), "val": (n_train, n_train + n_val), "test": (n_train + n_val, len(xs))}
        self.loaders = {}
        for k, (a, b) in cut.items():
            x, y = xs[a:b].copy(), ys[a:b].copy()
            x[..., 0] = (x[..., 0] - self.mean) / self.std                                             # traintest:276-277
            y[..., 0] = (y[..., 0] - self.mean) / self.std
            self.loaders[k] = DeviceDataLoader(x, y, batch_size, shuffle=(k == "train"), device=self.device)
        self.model = self.get_model()
        self.optimizer = torch.optim.Adam(self.model.parameters(), lr=lr, eps=eps)
        self.scheduler = torch.optim.lr_scheduler.MultiStepLR(self.optimizer, milestones=list(milestones), gamma=gamma)
        self.path = os.path.join(workdir, "standin_best.pt")
        self.batches_seen = 0
        self.log = []

    def get_model(self):
        return MegaCRN(**self.kw).to(self.device)

    def _loss(self, outs, y):
        output, _h_att, query, pos, neg = outs
        y_pred, y_true = output * self.std + self.mean, y * self.std + self.mean
        l1 = masked_mae(y_pred, y_true)
        l2 = nn.TripletMarginLoss(margin=1.0)(query, pos.detach(), neg.detach())
        l3 = nn.MSELoss()(query, pos.detach())
        return l1 + self.lamb * l2 + self.lamb1 * l3, l1

    def evaluate(self, model, mode):
        model = model.eval()
        losses, maes = [], []
        with torch.no_grad():
            for x, y, ycov in self.loaders[mode].get_iterator():
                loss, l1 = self._loss(model(x, ycov), y)
                losses.append(loss.item()); maes.append(l1.item())
        return float(np.mean(losses)), float(np.mean(maes))

    def train_epoch(self):
        model = self.model.train()
        losses = []
        for x, y, ycov in self.loaders["train"].get_iterator():
            self.optimizer.zero_grad()
            loss, _ = self._loss(model(x, ycov, y, self.batches_seen), y)
            losses.append(loss.item())
            self.batches_seen += 1
            loss.backward()
            torch.nn.utils.clip_grad_norm_(model.parameters(), self.max_grad_norm)
            self.optimizer.step()
        self.scheduler.step()
        return float(np.mean(losses))

    def fit(self, epochs, patience=20, verbose=False):
        best, wait = float("inf"), 0
        for ep in range(epochs):
            t0 = time.time()
            train_loss = self.train_epoch()
            val_loss, _ = self.evaluate(self.model, "val")
            test_loss, _ = self.evaluate(self.model, "test")
            self.log.append(dict(epoch=ep + 1, train_loss=train_loss, val_loss=val_loss, test_loss=test_loss,
                                 lr=self.optimizer.param_groups[0]["lr"], seconds=time.time() - t0))
            if verbose:
                print("Epoch [{}/{}] ({}) train_loss: {:.4f}, val_loss: {:.4f}, test_loss: {:.4f}, lr: {:.6f}, {:.1f}s".format(
                    ep + 1, epochs, self.batches_seen, train_loss, val_loss, test_loss, self.log[-1]["lr"], self.log[-1]["seconds"]), flush=True)
            if val_loss < best:
                best, wait = val_loss, 0
                torch.save(self.model.state_dict(), self.path)
            else:
                wait += 1
                if wait == patience:
                    break
        fresh = self.get_model()
        fresh.load_state_dict(torch.load(self.path))
        return best, self.evaluate(fresh, "val")[0], self.evaluate(fresh, "test")


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--num_nodes", type=int, default=207)
    ap.add_argument("--epochs", type=int, default=2)
    ap.add_argument("--batch_size", type=int, default=64)
    ap.add_argument("--n_train", type=int, default=512)
    ap.add_argument("--workdir", default=".")
    a = ap.parse_args(argv)
    np.random.seed(0); torch.manual_seed(0)
    tr = StandInTrainer(num_nodes=a.num_nodes, batch_size=a.batch_size, n_train=a.n_train, n_val=128, n_test=128, workdir=a.workdir)
    best, reloaded, (test_loss, test_mae) = tr.fit(a.epochs, verbose=True)
    print(f"best val_loss {best:.4f}; reloaded checkpoint val_loss {reloaded:.4f}; test loss {test_loss:.4f} mae {test_mae:.4f}")


if __name__ == "__main__":
    main()
