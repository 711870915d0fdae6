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


class bcolors:
    HEADER = '\033[95m'
    b = blue = OKBLUE = '\033[94m'
    g = green = OKGREEN = '\033[92m'
    y = yellow = WARNING = '\033[93m'
    r = red = FAIL = '\033[91m'
    c = cyan = '\033[36m'
    lb = lightblue = '\033[94m'
    p = pink = '\033[95m'
    o = orange = '\033[33m'
    lc = lightcyan = '\033[96m'
    end = ENDC = '\033[0m'
    BOLD = '\033[1m'
    UNDERLINE = '\033[4m'


def color_str(string, color):
    if not hasattr(bcolors, color):
        warnings.warn("Unknown color {}".format(color))
        return string
    return "{}{}{}".format(getattr(bcolors, color), string, bcolors.end)


def cprint(print_str, color):
    print(color_str(print_str, color))


def color_loss(loss):
    c = "red"
    if loss < 0.5:
        c = "yellow"
    if loss < 0.1:
        c = "green"
    if loss < 0.01:
        c = "cyan"
    return "{}{:.5f}{}".format(getattr(bcolors, c), loss, bcolors.end)


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
