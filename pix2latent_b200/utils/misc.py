"""Small host helpers used by the optimizers and examples (reference: pix2latent/utils/misc.py)."""
import os
import random
import sys
import warnings

import numpy as np
import torch
import torch.nn as nn


def set_seed(i):
    """Seeds torch / numpy / random. Like the reference (misc.py:17-22) this does NOT seed CMA."""
    torch.manual_seed(i)
    np.random.seed(i)
    random.seed(i)


def to_numpy(x):
    return x.detach().cpu().numpy()


def to_onehot(c, num_classes=1000):
    onehot = torch.zeros((1, num_classes))
    onehot[:, c] = 1.0
    return onehot


class HiddenPrints:
    """Context manager that silences print()."""

    def __enter__(self):
        self._stdout = sys.stdout
        sys.stdout = open(os.devnull, "w")

    def __exit__(self, exc_type, exc_val, exc_tb):
        sys.stdout.close()
        sys.stdout = self._stdout


_ANSI = {"HEADER": 95, "OKBLUE": 94, "OKGREEN": 92, "WARNING": 93, "FAIL": 91, "ENDC": 0, "BOLD": 1, "UNDERLINE": 4,
         "b": 94, "blue": 94, "g": 92, "green": 92, "y": 93, "yellow": 93, "r": 91, "red": 91, "c": 36, "cyan": 36,
         "lb": 94, "lightblue": 94, "p": 95, "pink": 95, "o": 33, "orange": 33, "lc": 96, "lightcyan": 96, "end": 0}
# attribute access (bcolors.red, bcolors.ENDC, ...) as the reference's scripts use it
bcolors = type("bcolors", (), {k: "\033[%dm" % v for k, v in _ANSI.items()})


def color_str(string, color):
    code = getattr(bcolors, color, None)
    if code is None:
        warnings.warn("Unknown color {}".format(color))
        return string
    return code + str(string) + bcolors.end


def cprint(print_str, color):
    print(color_str(print_str, color))


def color_loss(loss):
    """loss as a 5-decimal string, coloured by magnitude (red >= 0.5 > yellow >= 0.1 > green >= 0.01 > cyan)"""
    for bound, name in ((0.01, "cyan"), (0.1, "green"), (0.5, "yellow")):
        if loss < bound:
            return color_str("{:.5f}".format(loss), name)
    return color_str("{:.5f}".format(loss), "red")


def progress_print(phase, i, j, color="c", t=None):
    msg = "({}) progress {:.0f}% [{}/{}]".format(color_str(phase, color), (100. * i) / j, i, j)
    if t is not None:
        msg += " ({:.3f} sec/iter)".format(t)
    print(msg)


def replace_to_inplace_relu(model):
    for name, child in model.named_children():
        if isinstance(child, nn.ReLU):
            setattr(model, name, nn.ReLU(inplace=True))
        else:
            replace_to_inplace_relu(child)


def remove_spectral_norm(model):
    for n, m in model.named_modules():
        if isinstance(m, (nn.Conv2d, nn.Linear)):
            try:
                torch.nn.utils.remove_spectral_norm(m)
            except Exception:
                print("{} has no spectral_norm.".format(n))
