"""Training harness with the reference's interface (src/utils/train_utils.py:8-163): checkpoint
dictionaries with the same keys, `Trainer(model, model_fn, optimizer, ckpt_dir, lr_scheduler, ...)`.
Plain control flow around the accelerated modules; multi-GPU runs wrap the model in DDP
(planar_optical_flow_b200.parallel) and only rank 0 writes checkpoints."""
import os

import torch
from torch.nn.utils import clip_grad_norm_


def _unwrap(model):
    return model.module if hasattr(model, "module") and isinstance(
        model, (torch.nn.DataParallel, torch.nn.parallel.DistributedDataParallel)) else model


def checkpoint_state(model=None, optimizer=None, epoch=None, it=None):
    """{'epoch', 'it', 'model_state', 'optimizer_state'} — what reference checkpoints contain (:8-18)."""
    return {"epoch": epoch, "it": it,
            "model_state": _unwrap(model).state_dict() if model is not None else None,
            "optimizer_state": optimizer.state_dict() if optimizer is not None else None}


def save_checkpoint(state=None, filename="checkpoint"):
    torch.save(state, "{}.pth".format(filename))


def load_checkpoint(model=None, optimizer=None, filename="checkpoint", logger=None):
    """Returns (it, epoch); raises FileNotFoundError like the reference (:23-37)."""
    if not os.path.isfile(filename):
        print("Could not find %s" % filename)
        raise FileNotFoundError(filename)
    ckpt = torch.load(filename, map_location="cpu")
    if model is not None and ckpt.get("model_state") is not None:
        _unwrap(model).load_state_dict(ckpt["model_state"])
    if optimizer is not None and ckpt.get("optimizer_state") is not None:
        optimizer.load_state_dict(ckpt["optimizer_state"])
    return ckpt.get("it", 0.0), ckpt.get("epoch", -1)


def lr_scheduler():
    return 0.01


class ConstantLR:
    """What `optim.Adam(lr=tu.lr_scheduler())` amounts to in bin/train_dr_spaam.py:83."""

    def __init__(self, optimizer):
        self._optim = optimizer

    def step(self, epoch):
        pass

    def get_lr(self):
        return self._optim.param_groups[0]["lr"]


class LucasScheduler:
    """`v0` until epoch `e0`, exponential decay to `v1` at `e1`, `v1` afterwards (:42-68)."""

    def __init__(self, optimizer, e0, v0, e1, v1, eNone=float("inf")):
        self.e0, self.v0, self.e1, self.v1, self.eNone = e0, v0, e1, v1, eNone
        self._optim = optimizer

    def step(self, epoch):
        if epoch < self.e0:
            lr = self.v0
        elif epoch < self.e1:
            lr = self.v0 * (self.v1 / self.v0) ** ((epoch - self.e0) / (self.e1 - self.e0))
        else:
            lr = self.v1
        for group in self._optim.param_groups:
            group["lr"] = lr

    def get_lr(self):
        return self._optim.param_groups[0]["lr"]


class _NullWriter:
    def add_scalar(self, *a, **k):
        pass

    def flush(self):
        pass

    def close(self):
        pass


def create_tb_logger(root_dir, tb_log_dir="tensorboard"):
    """TensorBoard writer if the package is importable, else a no-op writer."""
    try:
        from torch.utils.tensorboard import SummaryWriter

        return SummaryWriter(log_dir=os.path.join(root_dir, tb_log_dir))
    except Exception:      # noqa: BLE001
        return _NullWriter()


class Trainer:
    def __init__(self, model, model_fn, optimizer, ckpt_dir, lr_scheduler, model_fn_eval=None, grad_norm_clip=1.0,
                 tb_logger=None, logger=None, is_main=True):
        self.model, self.model_fn, self.model_fn_eval = model, model_fn, model_fn_eval
        self.optimizer, self.ckpt_dir, self.grad_norm_clip = optimizer, ckpt_dir, grad_norm_clip
        self.lr_scheduler = lr_scheduler if lr_scheduler is not None else ConstantLR(optimizer)
        self.tb_logger = tb_logger if tb_logger is not None else _NullWriter()
        self.logger = logger
        self.is_main = is_main
        self._epoch = self._it = 0

    def train(self, num_epochs, train_loader, eval_loader=None, eval_frequency=1, ckpt_save_interval=5,
              lr_scheduler_each_iter=True, starting_epoch=0, starting_iteration=0, max_iters=None):
        self._it = starting_iteration
        last = None
        for self._epoch in range(starting_epoch, num_epochs):
            if not lr_scheduler_each_iter:
                self.lr_scheduler.step(self._epoch)
            running, n_batches = 0.0, max(len(train_loader), 1)
            sampler = getattr(train_loader, "sampler", None)
            if hasattr(sampler, "set_epoch"):          # DistributedSampler: a different shard order every epoch
                sampler.set_epoch(self._epoch)
            for cur_it, batch in enumerate(train_loader):
                if lr_scheduler_each_iter:
                    self.lr_scheduler.step(self._epoch + cur_it / n_batches)
                self.tb_logger.add_scalar("Learning_rate", self.lr_scheduler.get_lr(), self._it)
                last = self._train_it(batch)
                running += last
                self.tb_logger.add_scalar("Train_loss", last, self._it)
                self._it += 1
                if max_iters is not None and self._it - starting_iteration >= max_iters:
                    return last
            done = self._epoch + 1
            if self.is_main:
                print("Current Epoch: %d  [Learning rate: %s]  Epoch loss: %.6f" % (done, self.lr_scheduler.get_lr(), running / n_batches))
            self.tb_logger.add_scalar("Epoch_loss", running / n_batches, self._epoch)
            if self.is_main and done % ckpt_save_interval == 0:
                name = os.path.join(self.ckpt_dir, "ckpt_e{}".format(done))
                print("Saving checkpoint to {}".format(name))
                save_checkpoint(checkpoint_state(self.model, self.optimizer, done, self._it), filename=name)
            if eval_loader is not None and self.model_fn_eval is not None and done % eval_frequency == 0:
                with torch.no_grad():
                    metrics = self.model_fn_eval(self.model, eval_loader)
                if self.is_main:
                    print("Validation:", metrics)
            self.tb_logger.flush()
        return last

    def _train_it(self, batch):
        self.model.train()
        self.optimizer.zero_grad(set_to_none=True)
        out = self.model_fn(self.model, batch)
        loss = out[0] if isinstance(out, tuple) else out
        loss.backward()
        if self.grad_norm_clip > 0:
            clip_grad_norm_(self.model.parameters(), self.grad_norm_clip)
        self.optimizer.step()
        return loss.item()
