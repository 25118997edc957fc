"""CPU tests of host-side logic that needs no GPU: EMA decay schedule, discriminator tap tables / plane geometry."""
import math
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "stable-audio-tools_b200"))


def test_ema_decay_schedule_matches_published_formula():
    from b200sat.optim import ema_decay_at
    assert ema_decay_at(0) == 0.0 and ema_decay_at(1) == 0.0 and ema_decay_at(2) == 0.0
    for step in (3, 10, 1000, 123456):
        epoch = step - 1 - 1
        assert abs(ema_decay_at(step) - min(1.0 - (1.0 + epoch) ** -0.75, 0.9999)) < 1e-12
    assert ema_decay_at(10 ** 9) == 0.9999
    assert abs(ema_decay_at(50, beta=0.99, inv_gamma=2.0, power=1.0, update_after_step=10) - min(1 - (1 + 39 / 2.0) ** -1.0, 0.99)) < 1e-12


def test_discriminator_tap_tables_are_the_2d_taps_of_the_flattened_plane():
    """Row shift of tap (it, jf) of a (3 x 9, dilation (d, 1)) conv on a plane with row pitch Fp = F + 8 must move exactly `d` frames per
    time tap and one bin per frequency tap, be antisymmetric under tap reversal (the data-gradient conv reuses the table) and come in
    runs of 9 consecutive rows (what the shared-window path detects)."""
    for n_fft in (2048, 128):
        F = n_fft // 2 + 1
        Fp = F + 8
        for d in (1, 2, 4):
            offs = [(k // 9 - 1) * d * Fp + (k % 9 - 4) for k in range(27)]
            for k in range(27):
                it, jf = k // 9 - 1, k % 9 - 4
                assert offs[k] == it * d * Fp + jf
                assert offs[26 - k] == -offs[k]
            for g in range(3):
                assert all(offs[g * 9 + j] == offs[g * 9] + j for j in range(9))
            # a valid bin shifted by any tap stays inside its frame's row group: no wrap into the neighbouring frame
            for f in (0, F - 1):
                for jf in range(-4, 5):
                    assert 0 <= f + 4 + jf < Fp
        offs33 = [(k // 3 - 1) * Fp + (k % 3 - 1) for k in range(9)]
        assert all(offs33[8 - k] == -offs33[k] for k in range(9))


def test_hann_window_power_used_by_the_stft_front_end():
    """csrc/discriminator.cu normalises by 1/sqrt(3 n / 8): the sum of squares of the periodic Hann window."""
    import torch
    for n in (128, 512, 2048):
        w = torch.hann_window(n, periodic=True, dtype=torch.float64)
        assert abs(w.pow(2).sum().item() - 0.375 * n) < 1e-9 * n
