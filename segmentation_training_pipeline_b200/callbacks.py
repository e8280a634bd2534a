"""Host-side training controls of the stage loop: the YAML `callbacks:` block (reference README.md:147-158, 441-452;
schemas/callbacks.raml:22-48 -> keras.callbacks.EarlyStopping / ReduceLROnPlateau [DEP keras 2.2.4] and Brad Kenstler's
CyclicLR [DEP], README.md:437).  They run between replays of the captured step graph: a new learning rate is one 4-byte
host->device copy (Trainer.set_lr), nothing is re-captured."""
from __future__ import annotations

import math
from typing import Dict, List, Optional


def _auto_mode(mode: str, monitor: str) -> str:
    if mode in ("min", "max"):
        return mode
    return "max" if ("acc" in monitor or monitor.startswith("fmeasure")) else "min"   # keras 2.2.4 rule


class Callback:
    stop_training = False

    def on_train_begin(self, trainer):
        pass

    def on_batch_begin(self, trainer, iteration: int):
        pass

    def on_epoch_end(self, trainer, epoch: int, logs: Dict[str, float]):
        pass


class EarlyStopping(Callback):
    def __init__(self, monitor="val_loss", min_delta=0.0, patience=0, verbose=0, mode="auto", **_):
        self.monitor, self.patience, self.verbose = monitor, int(patience), verbose
        self.mode = _auto_mode(mode, monitor)
        self.min_delta = abs(float(min_delta)) * (1.0 if self.mode == "max" else -1.0)
        self.wait, self.best, self.stopped_epoch = 0, (-math.inf if self.mode == "max" else math.inf), None

    def on_train_begin(self, trainer):
        self.wait, self.stop_training = 0, False
        self.best = -math.inf if self.mode == "max" else math.inf

    def on_epoch_end(self, trainer, epoch, logs):
        cur = logs.get(self.monitor)
        if cur is None:
            return
        better = (cur - self.min_delta > self.best) if self.mode == "max" else (cur - self.min_delta < self.best)
        if better:
            self.best, self.wait = cur, 0
        else:
            self.wait += 1
            if self.wait >= self.patience:
                self.stopped_epoch, self.stop_training = epoch, True


class ReduceLROnPlateau(Callback):
    def __init__(self, monitor="val_loss", factor=0.1, patience=10, verbose=0, mode="auto", min_delta=1e-4, cooldown=0,
                 min_lr=0.0, **_):
        if factor >= 1.0:
            raise ValueError("ReduceLROnPlateau does not support a factor >= 1.0.")
        self.monitor, self.factor, self.patience, self.verbose = monitor, float(factor), int(patience), verbose
        self.min_delta, self.cooldown, self.min_lr = float(min_delta), int(cooldown), float(min_lr)
        self.mode = _auto_mode(mode, monitor)
        self.on_train_begin(None)

    def on_train_begin(self, trainer):
        self.best = -math.inf if self.mode == "max" else math.inf
        self.cooldown_counter, self.wait = 0, 0

    def on_epoch_end(self, trainer, epoch, logs):
        cur = logs.get(self.monitor)
        if cur is None:
            return
        if self.cooldown_counter > 0:
            self.cooldown_counter -= 1
            self.wait = 0
        better = (cur > self.best + self.min_delta) if self.mode == "max" else (cur < self.best - self.min_delta)
        if better:
            self.best, self.wait = cur, 0
        elif self.cooldown_counter <= 0:
            self.wait += 1
            if self.wait >= self.patience:
                old = trainer.get_lr()
                if old > self.min_lr:
                    trainer.set_lr(max(old * self.factor, self.min_lr))
                    self.cooldown_counter, self.wait = self.cooldown, 0


class CyclicLR(Callback):
    def __init__(self, base_lr=0.001, max_lr=0.006, step_size=2000.0, mode="triangular", gamma=1.0, **_):
        self.base_lr, self.max_lr, self.step_size, self.mode, self.gamma = float(base_lr), float(max_lr), float(step_size), mode, float(gamma)
        if mode not in ("triangular", "triangular2", "exp_range"):
            raise ValueError("CyclicLR: unknown mode " + str(mode))
        self.clr_iterations = 0

    def clr(self) -> float:
        it = self.clr_iterations
        cycle = math.floor(1 + it / (2 * self.step_size))
        x = abs(it / self.step_size - 2 * cycle + 1)
        scale = 1.0 if self.mode == "triangular" else (1.0 / (2.0 ** (cycle - 1)) if self.mode == "triangular2" else self.gamma ** it)
        return self.base_lr + (self.max_lr - self.base_lr) * max(0.0, 1 - x) * scale

    def on_train_begin(self, trainer):
        trainer.set_lr(self.base_lr if self.clr_iterations == 0 else self.clr())

    def on_batch_begin(self, trainer, iteration):
        if iteration > 0:          # keras: on_batch_end of batch k-1 sets the rate batch k runs with
            self.clr_iterations += 1
            trainer.set_lr(self.clr())


class LRVariator(Callback):
    """musket_core's LRVariator (reference README.md:454, listed with ReduceLROnPlateau as the way "to modify learning rate on
    the fly") [DEP musket_core, unpinned]: the rate moves from `fromVal` to `toVal` over `relSize` epochs' worth of batches
    (or `absSize` batches) following `style` -- linear, const, cos / cos+ / cos- or sin / sin+ / sin- -- and then stays at
    `toVal`.  fromVal defaults to the rate the stage starts with."""

    STYLES = ("linear", "const", "cos", "cos+", "cos-", "sin", "sin+", "sin-")

    def __init__(self, relSize=1.0, toVal=0.0, fromVal=None, style="linear", absSize=None, steps_per_epoch=None, **_):
        if style not in self.STYLES:
            raise ValueError("LRVariator: unknown style %r (known: %s)" % (style, ", ".join(self.STYLES)))
        self.relSize, self.toVal, self.fromVal, self.style = float(relSize), float(toVal), fromVal, style
        self.absSize, self.steps_per_epoch = absSize, steps_per_epoch
        self.total, self.it = None, 0

    def shape(self, t: float) -> float:
        """progress t in [0, 1] -> fraction of the way from fromVal to toVal"""
        st = self.style
        if st == "linear":
            return t
        if st == "const":
            return 0.0 if t < 1.0 else 1.0
        if st in ("cos", "cos+"):
            return 0.5 * (1.0 - math.cos(math.pi * t))          # slow start, slow end
        if st == "cos-":
            return 1.0 - math.cos(0.5 * math.pi * t)            # slow start
        if st in ("sin", "sin+"):
            return math.sin(0.5 * math.pi * t)                  # fast start
        return 1.0 - math.sin(0.5 * math.pi * (1.0 - t))        # sin-

    def on_train_begin(self, trainer):
        if self.fromVal is None:
            self.fromVal = trainer.get_lr()
        self.fromVal = float(self.fromVal)
        spe = self.steps_per_epoch or getattr(trainer, "steps_per_epoch", None) or 1
        self.total = int(self.absSize) if self.absSize else max(1, int(round(self.relSize * spe)))
        self.it = 0
        trainer.set_lr(self.fromVal)

    def on_batch_begin(self, trainer, iteration):
        t = min(1.0, self.it / float(self.total))
        trainer.set_lr(self.fromVal + (self.toVal - self.fromVal) * self.shape(t))
        self.it += 1


_REGISTRY = {"EarlyStopping": EarlyStopping, "ReduceLROnPlateau": ReduceLROnPlateau, "CyclicLR": CyclicLR,
             "LRVariator": LRVariator}


def build(spec: Optional[dict], extra: Optional[dict] = None) -> List[Callback]:
    """YAML mapping {Name: {kwargs}} -> callback objects; unknown names raise (nothing is silently ignored)."""
    out: List[Callback] = []
    for block in (spec, extra):
        for name, kw in (block or {}).items():
            if name not in _REGISTRY:
                raise NotImplementedError("callback '%s' is not implemented (known: %s)" % (name, sorted(_REGISTRY)))
            out.append(_REGISTRY[name](**(kw or {})))
    return out
