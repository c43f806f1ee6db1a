"""MAD module samplers — same factory API as the reference (Sampler/sampler_factory.py:4-90).

get_sampler(name, blocks_to_fetch, fixed_id) -> object with .sample(distribution) returning the indices of
the network portions to train this frame.  Host-side numpy policy (it is ~5 floats per frame).
"""
import numpy as np


class meta_sampler(object):
    """Sampler for MAD adaptation."""

    def __init__(self, blocks_to_fetch):
        self._blocks_to_fetch = blocks_to_fetch

    def sample(self, distribution):
        raise NotImplementedError


class fixed_sampler(meta_sampler):
    """Always the same group (no sampling). Accepts an int or the driver's one-element list for fixed_id."""

    def __init__(self, blocks_to_fetch, fixed_id):
        super(fixed_sampler, self).__init__(blocks_to_fetch)
        if isinstance(fixed_id, (list, tuple, np.ndarray)):   # argparse nargs='+' quirk, Stereo_Online_Adaptation.py:304
            if len(fixed_id) > 1:
                print('WARNING: fixed sampler trains ONE group per frame; using fixedID {} and ignoring {}'.format(fixed_id[0], list(fixed_id[1:])))
            fixed_id = fixed_id[0]
        self._fixed_id = int(fixed_id)

    def sample(self, distribution):
        return [self._fixed_id]


class random_sampler(meta_sampler):
    def sample(self, distribution):
        return np.random.choice(range(distribution.shape[0]), size=self._blocks_to_fetch, replace=False)


class argmax_sampler(meta_sampler):
    def sample(self, distribution):
        k = self._blocks_to_fetch
        return np.argpartition(np.squeeze(distribution), -k)[-k:]


class sequential_sampler(meta_sampler):
    """Round robin."""

    def __init__(self, blocks_to_fetch):
        super(sequential_sampler, self).__init__(blocks_to_fetch)
        self._sample_counter = 0

    def sample(self, distribution):
        n = distribution.shape[0]
        base = self._sample_counter % n
        self._sample_counter += 1
        return [(base + i) % n for i in range(self._blocks_to_fetch)]


class probabilistic_sampler(meta_sampler):
    def sample(self, distribution):
        return np.random.choice(range(distribution.shape[0]), size=self._blocks_to_fetch, replace=False,
                                p=np.squeeze(distribution))


SAMPLER_FACTORY = {
    'FIXED': fixed_sampler,
    'RANDOM': random_sampler,
    'ARGMAX': argmax_sampler,
    'SEQUENTIAL': sequential_sampler,
    'PROBABILITY': probabilistic_sampler,
}

AVAILABLE_SAMPLER = SAMPLER_FACTORY.keys()


def get_sampler(name, blocks_to_fetch, fixed_id=0):
    assert (name in AVAILABLE_SAMPLER)
    if name == 'FIXED':
        return SAMPLER_FACTORY[name](blocks_to_fetch, fixed_id)
    return SAMPLER_FACTORY[name](blocks_to_fetch)
