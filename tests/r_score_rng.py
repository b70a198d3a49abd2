"""R's `set.seed(seed); runif(k)` on numpy's own MT19937 (test infrastructure, independent of the library's mt_seed /
mt_unif): Randomize() scrambles the seed 50 times with the LCG 69069*s + 1, RNG_Init fills 625 words with the same LCG
(the first is the dummy position word), position 624 forces a fresh block; unif_rand = word * 2.3283064365386963e-10
with fixup() keeping the value inside (0, 1) (R core src/main/RNG.c)."""
import numpy as np


class RUnif:
    def __init__(self, seed: int):
        s = np.uint32(seed)
        with np.errstate(over="ignore"):
            for _ in range(50):
                s = np.uint32(69069) * s + np.uint32(1)
            words = np.empty(625, dtype=np.uint32)
            for i in range(625):
                s = np.uint32(69069) * s + np.uint32(1)
                words[i] = s
        self.bg = np.random.MT19937()
        self.bg.state = {"bit_generator": "MT19937", "state": {"key": words[1:].copy(), "pos": 624}}

    def __call__(self, k: int) -> np.ndarray:
        u = self.bg.random_raw(k).astype(np.float64) * 2.3283064365386963e-10
        half = 0.5 * 2.328306437080797e-10
        u[u <= 0.0] = half
        u[(1.0 - u) <= 0.0] = 1.0 - half
        return u
