"""Checkpoint import with the reference's matching rules, without TensorFlow.

Mirrors Data_utils/weights_utils.py of the reference: `get_var_to_restore_list` (:4-38: skip variables whose name
contains a `mask` entry, strip every `ignore_list` substring from the checkpoint key, prepend `prefix`, keep the key if
the result names a model variable) and `check_for_weights_or_restore_them` (:41-80: newest checkpoint of a log directory,
else the given initial weights, which may be a directory).  `tf.train.NewCheckpointReader` / `Saver.restore` are replaced
by madstereo.tf_checkpoint (TF "V2" bundles) and by `.npz` archives of name -> array.

`model` below is anything with `get_variable_names()` / `load_weights(dict)`: madstereo.adaptation.OnlineAdaptation.
"""
import os
import re

import numpy as np

from madstereo.tf_checkpoint import CheckpointReader, CheckpointError  # noqa: F401


def _open(ckpt_path):
    """-> (shape map, getter).  Accepts a TF V2 checkpoint prefix or an .npz archive."""
    if ckpt_path.endswith('.npz'):
        z = np.load(ckpt_path)
        return {k: list(z[k].shape) for k in z.files}, (lambda k: np.asarray(z[k]))
    r = CheckpointReader(ckpt_path)
    return r.get_variable_to_shape_map(), r.get_tensor


def get_var_to_restore_list(ckpt_path, mask=[], prefix="", ignore_list=[], variables=()):
    """{checkpoint key: model variable name} -- the reference returns the tf.Variable, here its name (without ':0')."""
    variables_dict = {}
    for name in variables:
        name = name[:-2] if name.endswith(':0') else name
        if not any(m in name for m in mask):
            variables_dict[name] = name
    var_to_shape_map, _get = _open(ckpt_path)
    var_to_restore = {}
    for key in var_to_shape_map:
        t_key = key
        for ig in ignore_list:
            t_key = t_key.replace(ig, '')
        if prefix + t_key in variables_dict:
            var_to_restore[key] = variables_dict[prefix + t_key]
    return var_to_restore


def load_weights(ckpt_path, variables, mask=[], prefix="", ignore_list=[]):
    """{model variable name: array} for every checkpoint entry that matches (what Saver.restore would assign)."""
    mapping = get_var_to_restore_list(ckpt_path, mask, prefix, ignore_list, variables)
    _shapes, get = _open(ckpt_path)
    return {var: get(key) for key, var in mapping.items()}


def latest_checkpoint(logdir):
    """tf.train.latest_checkpoint: the prefix named by `<logdir>/checkpoint` (model_checkpoint_path: "...")."""
    state = os.path.join(logdir, 'checkpoint')
    if not os.path.exists(state):
        return None
    m = re.search(r'^model_checkpoint_path:\s*"(.*)"\s*$', open(state).read(), re.M)
    if not m:
        return None
    p = m.group(1)
    p = p if os.path.isabs(p) else os.path.join(logdir, p)
    return p if os.path.exists(p + '.index') else None


def check_for_weights_or_restore_them(logdir, model, initial_weights=None, prefix='', ignore_list=[]):
    """-> (restored?, step).  `model` replaces the reference's `session` argument."""
    names = model.get_variable_names()
    ckpt = latest_checkpoint(logdir) if logdir else None
    if ckpt:
        print('Found valid checkpoint file: {}'.format(ckpt))
        model.load_weights(load_weights(ckpt, names, [], prefix=""), strict=False)
        step = int(ckpt.split('-')[-1]) if ckpt.split('-')[-1].isdigit() else 0
        return True, step
    elif initial_weights is not None:
        if os.path.isdir(initial_weights):
            found = latest_checkpoint(initial_weights)
            if found is None:
                raise Exception('no usable checkpoint in directory {} (missing `checkpoint` state file or its .index)'.format(initial_weights))
            initial_weights = found
        w = load_weights(initial_weights, names, [], prefix=prefix, ignore_list=ignore_list)
        print('Found {} variables to restore in {}'.format(len(w), initial_weights))
        if len(w) > 0:
            model.load_weights(w, strict=False)
            return True, 0
        return False, 0
    else:
        print('Unable to restore any weight')
        return False, 0
