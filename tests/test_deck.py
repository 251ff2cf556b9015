"""The synthetic deck generator: deterministic, CPU and CUDA builds bit-identical."""
import zlib

import numpy as np
import pytest

from util import deck_frames, deck_truth


def test_deck_is_deterministic():
    a = deck_frames(40, 2)
    b = deck_frames(41, 1)
    assert np.array_equal(a[1], b[0])
    assert zlib.crc32(deck_frames(0, 1).tobytes()) == zlib.crc32(deck_frames(0, 1, threads=1).tobytes())


def test_deck_truth_is_luhn_valid(oracle):
    for f in (0, 8, 16, 24):
        digits, quad = deck_truth(f)
        assert len(digits) in (15, 16) and oracle.luhn(digits)
        assert oracle.card_type(digits) in (2, 4)  # amex / visa
        base = np.array([106, 105, 533, 105, 106, 374, 533, 374], float)
        assert np.abs(quad - base).max() <= 8.0


def test_other_resolutions_detect(oracle):
    for (w, h) in ((1280, 720),):
        fr = deck_frames(0, 1, w, h)
        d = oracle.detect_edges(fr[0])
        assert d.all_found


@pytest.mark.gpu
def test_cuda_deck_equals_cpu_deck():
    from util import deck_frames_cuda
    for (w, h) in ((640, 480), (1280, 720)):
        g = deck_frames_cuda(5, 3, w, h).cpu().numpy()
        assert np.array_equal(g, deck_frames(5, 3, w, h))
