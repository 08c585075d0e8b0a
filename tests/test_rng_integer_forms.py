"""CPU: the integer forms the CUDA kernels use for the reference's float RNG transforms
(tg_device.cuh: noisy_from_k, the JUMP and flip thresholds) are exactly CPython's float results
for every 53-bit draw k (u = k / 2**53), including the round-half-even ties."""
import random


def py_noisy(val, u):                      # _treasure_game_impl.py:361-366
    mid = val / 2.0
    if val < mid:
        return int(round(val + (mid - val) * u))
    return int(round(mid + (val - mid) * u))


def k_noisy(k, negative):                  # tg_device.cuh noisy_from_k
    q = k >> 1
    m = q + (k & q & 1)
    return (-4 if negative else 2) + int(m > (1 << 50)) + int(m >= (3 << 50))


def test_integer_forms_match_cpython_floats():
    rng = random.Random(1)
    ks = [0, 1, 2, 3, (1 << 53) - 1, (1 << 53) - 2, (1 << 53) - 3, 7205759403792794, 7205759403792795]
    for base in (1 << 50, 3 << 50, 1 << 51, 3 << 51, 1 << 52):
        ks += [base + d for d in range(-6, 7)]
    ks += [rng.getrandbits(53) for _ in range(300000)]
    for k in ks:
        u = k / 9007199254740992.0
        assert u * 9007199254740992.0 == k
        assert py_noisy(-4, u) == k_noisy(k, True)
        assert py_noisy(4, u) == k_noisy(k, False)
        assert (u > 0.25) == (k > (1 << 51))                       # _treasure_game_impl.py:318
        assert (0 + (1 - 0) * u <= 0.8) == (k <= 7205759403792794)  # _objects.py:119
