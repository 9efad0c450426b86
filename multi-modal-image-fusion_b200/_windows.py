"""Window buffers the drop-in modules register (state_dict compatibility with the reference)."""
from math import exp

import torch


def gauss_1d(win_size, sigma):
    g = torch.FloatTensor([exp(-(x - win_size // 2) ** 2 / (2.0 * sigma ** 2)) for x in range(win_size)])
    return g / g.sum()


def loss_window(win_size):
    """(1,1,k,k) buffer with the reference's window-size -> sigma rule (loss.py:33-39)."""
    sigma = 1.5 if win_size == 11 else 0.15 * (win_size - 1)
    w = gauss_1d(win_size, sigma).unsqueeze(1)
    return torch.mm(w, w.t()).unsqueeze(0).unsqueeze(0)
