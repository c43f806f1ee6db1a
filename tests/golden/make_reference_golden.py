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


def driver_cli(script='Stereo_Online_Adaptation.py'):
    """The argparse definition of the reference's Stereo_Online_Adaptation.py, captured by running the script as __main__
    with `tensorflow` replaced by the eager shim and `parse_args` intercepted (nothing else of the script executes)."""
    import argparse
    import runpy
    import sys
    import types
    root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    sys.path.insert(0, root)
    from oracle import tf1_shim
    sys.modules['tensorflow'] = tf1_shim.as_module()
    mpl = types.ModuleType('matplotlib'); mpl.__path__ = []
    mpl.cm = types.ModuleType('matplotlib.cm'); mpl.pyplot = types.ModuleType('matplotlib.pyplot')
    mpl.colors = types.ModuleType('matplotlib.colors'); mpl.colors.LinearSegmentedColormap = object
    for k, v in (('matplotlib', mpl), ('matplotlib.cm', mpl.cm), ('matplotlib.pyplot', mpl.pyplot), ('matplotlib.colors', mpl.colors)):
        sys.modules.setdefault(k, v)
    for k in [k for k in sys.modules if k.split('.')[0] in ('Nets', 'Losses', 'Data_utils', 'Sampler')]:
        del sys.modules[k]
    captured = {}

    class Stop(Exception):
        pass

    def fake_parse(self, *a, **k):
        captured['parser'] = self
        raise Stop()

    orig = argparse.ArgumentParser.parse_args
    argparse.ArgumentParser.parse_args = fake_parse
    sys.path.insert(0, REF)
    old_argv = sys.argv
    try:
        sys.argv = [script]
        runpy.run_path(os.path.join(REF, script), run_name='__main__')
    except Stop:
        pass
    finally:
        sys.argv = old_argv
        argparse.ArgumentParser.parse_args = orig
        sys.path.remove(REF)
    acts = []
    for a in captured['parser']._actions:
        if a.dest == 'help':
            continue
        acts.append({'flags': list(a.option_strings), 'dest': a.dest, 'required': bool(a.required), 'default': a.default,
                     'type': getattr(a.type, '__name__', None), 'nargs': a.nargs, 'choices': sorted(a.choices) if a.choices else None,
                     'store_true': isinstance(a, argparse._StoreTrueAction)})
    return acts


def main_cli():
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_driver_cli.json')
    json.dump({'source': 'CVLAB-Unibo/Real-time-self-adaptive-deep-stereo: Stereo_Online_Adaptation.py (argparse actions)',
               'actions': driver_cli()}, open(out_path, 'w'), indent=1)
    print('wrote', out_path)
    out_path = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_continual_cli.json')
    json.dump({'source': 'CVLAB-Unibo/Real-time-self-adaptive-deep-stereo: Stereo_Continual_Adaptation.py (argparse actions)',
               'actions': driver_cli('Stereo_Continual_Adaptation.py')}, open(out_path, 'w'), indent=1)
    print('wrote', out_path)


if __name__ == '__main__':
    main_cli()
