"""Golden vectors produced by the REFERENCE's own Python code, imported from /root/reference in the build container.

Only the TensorFlow-free parts of the hot path can be imported here: Sampler/sampler_factory.py and the
block_config/*.json train configurations (a15).  This script drives them with seeded numpy RNG and writes
tests/golden/reference_sampler.json; tests/test_host_cpu.py replays the same seeds through this repository's mirror and
compares exactly.  The reference path is read ONLY here, never by a test.

    python tests/golden/make_reference_golden.py
"""
import importlib.util
import json
import os

import numpy as np

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_sampler.json')


def load_reference_sampler():
    spec = importlib.util.spec_from_file_location('ref_sampler_factory', os.path.join(REF, 'Sampler', 'sampler_factory.py'))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def softmax(x):
    e = np.exp(x - np.max(x))
    return e / e.sum()


def main():
    ref = load_reference_sampler()
    cases = []
    rng = np.random.default_rng(2024)
    dists = [np.ones(5) / 5, softmax(np.array([0.3, -0.1, 0.0, 0.2, -0.4])), softmax(rng.normal(size=5)),
             softmax(rng.normal(size=6) * 3.0)]
    for name in ('FIXED', 'RANDOM', 'ARGMAX', 'SEQUENTIAL', 'PROBABILITY'):
        for blocks in (1, 2, 3):
            for di, d in enumerate(dists):
                seed = 1000 + 37 * blocks + di
                np.random.seed(seed)
                s = ref.get_sampler(name, blocks, fixed_id=2)
                draws = [[int(v) for v in s.sample(d)] for _ in range(7)]
                if name == 'ARGMAX':                     # argpartition order inside the top-k is unspecified
                    draws = [sorted(v) for v in draws]
                cases.append({'name': name, 'blocks': blocks, 'distribution': [float(v) for v in d], 'seed': seed,
                              'draws': draws})
    configs = {}
    for fn in sorted(os.listdir(os.path.join(REF, 'block_config'))):
        if fn.endswith('.json'):
            configs[fn] = json.load(open(os.path.join(REF, 'block_config', fn)))
    out = {'source': 'CVLAB-Unibo/Real-time-self-adaptive-deep-stereo: Sampler/sampler_factory.py, block_config/*.json',
           'available_sampler': sorted(ref.AVAILABLE_SAMPLER), 'cases': cases, 'block_config': configs}
    json.dump(out, open(OUT, 'w'), indent=1)
    print('wrote', OUT, len(cases), 'sampler cases,', len(configs), 'block configs')


if __name__ == '__main__':
    main()
