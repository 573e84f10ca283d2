"""CPU oracle (test infrastructure only) for the temperature-ladder sampler: whisper.cpp whisper_sample_token(best = false),

    std::discrete_distribution<> dist(probs.begin(), probs.end());  id = dist(decoder.rng);      // decoder.rng: std::mt19937

restated from the C++ standard (mt19937) and libstdc++'s implementation of discrete_distribution / generate_canonical<double, 53>:
weights -> double, normalised by their sequential sum, cumulative sums (last forced to 1), draw p = (g1 + g2 * 2^32) / 2^64 from two
32-bit outputs, index = lower_bound(cumulative, p).  The raw generator is numpy's MT19937 under legacy (init_genrand) seeding, which
is the seeding std::mt19937(seed) performs; test_oracle_decoder pins it to the standard's 10000th-output known answer."""
import numpy as np


class Mt19937:
    def __init__(self, seed):
        self.bg = np.random.MT19937()
        self.bg._legacy_seeding(int(seed))

    def next_u32(self):
        return int(self.bg.random_raw(1)[0])

    def canonical(self):
        """std::generate_canonical<double, 53>(mt19937): two outputs, low word first."""
        g1 = float(self.next_u32())
        g2 = float(self.next_u32())
        r = (g1 + g2 * 4294967296.0) / 18446744073709551616.0
        return r if r < 1.0 else float(np.nextafter(1.0, 0.0))


def draw(logprobs, rng):
    lp = np.asarray(logprobs, np.float32)
    probs = np.where(np.isneginf(lp), np.float32(0.0), np.exp(lp, dtype=np.float32)).astype(np.float64)
    total = np.cumsum(probs)[-1]  # sequential accumulation, as std::accumulate
    cp = np.cumsum(probs / total)
    cp[-1] = 1.0
    return int(np.searchsorted(cp, rng.canonical(), side="left"))


def sample_discrete(logprobs, seed, n_draws):
    rng = Mt19937(seed)
    return np.array([draw(logprobs, rng) for _ in range(n_draws)], np.int32)
