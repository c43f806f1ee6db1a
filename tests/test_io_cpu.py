"""Host-side I/O next to the hot path: TF V2 checkpoint import/export without TensorFlow, the
reference's checkpoint-key matching rules, the list-file / image reader and the driver's output files."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'real-time-self-adaptive-deep-stereo_b200')
sys.path.insert(0, ROOT)
sys.path.insert(0, PKG)

from madstereo import tf_checkpoint as tfc  # noqa: E402
from Data_utils import weights_utils  # noqa: E402


def test_reader_on_hand_assembled_tensorbundle():
    """An INDEPENDENT file: tests/golden/tensorbundle_handmade.* was assembled byte by byte from the LevelDB table format
    and tensor_bundle.proto by tests/golden/make_tensorbundle_fixture.py, which shares no code with the module under
    test (own varint / crc32c / protobuf emitters; shared key prefixes, two data blocks, int32 scalar)."""
    import json
    prefix = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'tensorbundle_handmade')
    want = json.load(open(prefix + '.json'))
    r = tfc.CheckpointReader(prefix, verify=True)                     # block and tensor checksums verified
    assert sorted(r.get_variable_to_shape_map()) == sorted(want)
    for name, w in want.items():
        assert r.has_tensor(name)
        got = r.get_tensor(name)
        assert list(got.shape) == w['shape'] and str(got.dtype) == w['dtype'].lstrip('<')
        assert np.array_equal(got.ravel(), np.asarray(w['values'], dtype=got.dtype))
    assert int(r.get_tensor('global_step')) == 4200
    assert r.header['num_shards'] == 1


def test_crc32c_lanes_match_the_bytewise_definition():
    rng = np.random.default_rng(1)
    for n in (262144 + 17, 1000003):
        d = rng.integers(0, 256, n, dtype=np.uint8)
        ref = tfc._crc_scalar(d, 0xFFFFFFFF) ^ 0xFFFFFFFF
        assert tfc.crc32c(d) == ref
        assert tfc.crc32c(d[n // 3:], tfc.crc32c(d[:n // 3])) == ref    # incremental form


def test_crc32c_known_answers_and_masking():
    assert tfc.crc32c(b'123456789') == 0xE3069283               # the standard CRC-32C check value
    assert tfc.crc32c(b'') == 0
    assert tfc.crc32c(bytes(32)) == 0x8A9136AA                   # RFC 3720 B.4: 32 zero bytes
    for v in (0, 1, 0xE3069283, 0xFFFFFFFF):
        assert tfc.unmask_crc(tfc.mask_crc(v)) == v


def _tensors(seed=0):
    rng = np.random.default_rng(seed)
    t = {'model/gc-read-pyramid/conv%d/weights' % i: rng.standard_normal((3, 3, 4, 5)).astype(np.float32) for i in range(1, 40)}
    t['model/gc-read-pyramid/conv1/biases'] = rng.standard_normal(5).astype(np.float32)
    t['model/G6/fgc-volume-filtering-6/disp-1/weights/Momentum'] = rng.standard_normal((3, 3, 2, 2)).astype(np.float32)
    t['global_step'] = np.array(1234, dtype=np.int64)
    t['empty'] = np.zeros((0, 3), dtype=np.float32)
    return t


def test_checkpoint_round_trip_multi_block(tmp_path):
    t = _tensors()
    prefix = str(tmp_path / 'model.ckpt-100')
    tfc.write_checkpoint(prefix, t, block_size=512)             # small blocks: several data blocks + a real index block
    assert os.path.exists(prefix + '.index') and os.path.exists(prefix + '.data-00000-of-00001')
    r = tfc.CheckpointReader(prefix, verify=True)                # verifies block and tensor CRC32C
    shapes = r.get_variable_to_shape_map()
    assert set(shapes) == set(t)
    assert shapes['global_step'] == [] and shapes['empty'] == [0, 3]
    for k, v in t.items():
        got = r.get_tensor(k)
        assert got.dtype == v.dtype and got.shape == v.shape and np.array_equal(got, v), k
    assert r.get_variable_to_dtype_map()['global_step'] == np.int64
    assert not r.has_tensor('nope')
    with pytest.raises(tfc.CheckpointError):
        r.get_tensor('nope')


def test_checkpoint_rejects_foreign_and_corrupt_files(tmp_path):
    p = str(tmp_path / 'x')
    open(p + '.index', 'wb').write(b'not a table' * 10)
    with pytest.raises(tfc.CheckpointError):
        tfc.CheckpointReader(p)
    with pytest.raises(tfc.CheckpointError):
        tfc.CheckpointReader(str(tmp_path / 'missing'))
    prefix = str(tmp_path / 'c')
    tfc.write_checkpoint(prefix, {'a': np.arange(100, dtype=np.float32)})
    raw = bytearray(open(prefix + '.data-00000-of-00001', 'rb').read())
    raw[17] ^= 0x40
    open(prefix + '.data-00000-of-00001', 'wb').write(bytes(raw))
    assert tfc.CheckpointReader(prefix).get_tensor('a').shape == (100,)          # unverified read still works
    with pytest.raises(tfc.CheckpointError):
        tfc.CheckpointReader(prefix, verify=True).get_tensor('a')                # checksum catches the flipped bit


def test_restore_list_follows_reference_matching_rules(tmp_path):
    """Data_utils/weights_utils.py:4-38 -- mask skips model variables, ignore_list strips substrings from checkpoint keys,
    prefix is prepended before the lookup; optimizer slots and unknown keys are ignored."""
    t = _tensors()
    prefix = str(tmp_path / 'm')
    tfc.write_checkpoint(prefix, t)
    variables = ['model/gc-read-pyramid/conv1/weights:0', 'model/gc-read-pyramid/conv1/biases:0',
                 'model/gc-read-pyramid/conv2/weights:0', 'model/context-1/weights:0']
    m = weights_utils.get_var_to_restore_list(prefix, variables=variables)
    assert m == {'model/gc-read-pyramid/conv1/weights': 'model/gc-read-pyramid/conv1/weights',
                 'model/gc-read-pyramid/conv1/biases': 'model/gc-read-pyramid/conv1/biases',
                 'model/gc-read-pyramid/conv2/weights': 'model/gc-read-pyramid/conv2/weights'}
    m = weights_utils.get_var_to_restore_list(prefix, mask=['conv2'], variables=variables)
    assert 'model/gc-read-pyramid/conv2/weights' not in m and len(m) == 2
    stripped = ['gc-read-pyramid/conv1/weights', 'gc-read-pyramid/conv1/biases']
    m = weights_utils.get_var_to_restore_list(prefix, ignore_list=['model/'], variables=stripped)
    assert m == {'model/gc-read-pyramid/conv1/weights': stripped[0], 'model/gc-read-pyramid/conv1/biases': stripped[1]}
    m = weights_utils.get_var_to_restore_list(prefix, prefix='net/', variables=['net/model/gc-read-pyramid/conv3/weights'])
    assert m == {'model/gc-read-pyramid/conv3/weights': 'net/model/gc-read-pyramid/conv3/weights'}
    w = weights_utils.load_weights(prefix, variables)
    assert set(w) == set(v[:-2] for v in variables[:3]) and np.array_equal(w[variables[0][:-2]], t[variables[0][:-2]])
    # an .npz archive of name -> array is accepted wherever a checkpoint prefix is
    np.savez(str(tmp_path / 'w.npz'), **{k: v for k, v in t.items() if 'conv1/' in k})
    assert set(weights_utils.load_weights(str(tmp_path / 'w.npz'), variables)) == {variables[0][:-2], variables[1][:-2]}


class _Model:
    def __init__(self, names):
        self.names, self.loaded = names, None

    def get_variable_names(self):
        return self.names

    def load_weights(self, w, strict=True):
        self.loaded, self.strict = w, strict


def test_check_for_weights_or_restore_them(tmp_path):
    t = _tensors()
    names = ['model/gc-read-pyramid/conv1/weights', 'model/gc-read-pyramid/conv1/biases']
    logdir = tmp_path / 'log'; logdir.mkdir()
    init = tmp_path / 'init'; init.mkdir()
    tfc.write_checkpoint(str(init / 'weights.ckpt'), t)
    open(str(init / 'checkpoint'), 'w').write('model_checkpoint_path: "weights.ckpt"\nall_model_checkpoint_paths: "weights.ckpt"\n')
    m = _Model(names)
    assert weights_utils.check_for_weights_or_restore_them(str(logdir), m) == (False, 0) and m.loaded is None
    assert weights_utils.check_for_weights_or_restore_them(str(logdir), m, initial_weights=str(init)) == (True, 0)
    assert set(m.loaded) == set(names) and m.strict is False
    tfc.write_checkpoint(str(logdir / 'model.ckpt-4200'), t)
    open(str(logdir / 'checkpoint'), 'w').write('model_checkpoint_path: "model.ckpt-4200"\n')
    assert weights_utils.check_for_weights_or_restore_them(str(logdir), _Model(names), initial_weights=str(init)) == (True, 4200)
    assert weights_utils.check_for_weights_or_restore_them(str(logdir), _Model(['zzz']), initial_weights=None)[0] is True


# ---------------------------------------------------------------------------------------------------
# input pipeline
# ---------------------------------------------------------------------------------------------------
def _write_dataset(tmp_path, n=5, h=20, w=30, gt16=True):
    import cv2
    from Data_utils import data_reader as dr  # noqa: F401
    rng = np.random.default_rng(3)
    lines, frames = ['# left;right;gt', ''], []
    for i in range(n):
        l = rng.integers(0, 256, (h, w, 3), dtype=np.uint8); r = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        g = rng.integers(0, 40 * 256, (h, w + 4), dtype=np.uint16) if gt16 else rng.integers(0, 200, (h, w + 4), dtype=np.uint8)
        pl, pr, pg = (str(tmp_path / ('%s_%d.png' % (k, i))) for k in 'lrg')
        cv2.imwrite(pl, l[:, :, ::-1]); cv2.imwrite(pr, r[:, :, ::-1]); cv2.imwrite(pg, g)      # cv2 writes BGR
        lines.append('%s;%s,%s' % (pl, pr, pg))                                                   # both separators
        frames.append((l, r, g))
    lst = str(tmp_path / 'list.csv')
    open(lst, 'w').write('\n'.join(lines) + '\n')
    return lst, frames


def test_list_file_and_image_decoding(tmp_path):
    from Data_utils import data_reader as dr
    lst, frames = _write_dataset(tmp_path)
    left, right, gt, conf = dr.read_list_file(lst)
    assert len(left) == len(right) == len(gt) == 5 and conf == []
    img = dr.read_image_from_disc(left[0])
    assert img.dtype == np.float32 and np.array_equal(img, frames[0][0].astype(np.float32))      # RGB order, 0..255
    g = dr.read_gt_from_disc(gt[0])
    assert g.shape == (20, 34, 1) and np.allclose(g[..., 0], frames[0][2] / 256.0)                 # 16-bit PNG / 256
    with pytest.raises(Exception):
        dr.dataset(str(tmp_path / 'missing.csv'))


def test_crop_or_pad_matches_oracle_restatement():
    import torch
    from Data_utils import data_reader as dr
    from oracle import tf1_ops as T
    rng = np.random.default_rng(0)
    for (h, w, th, tw) in [(10, 13, 7, 9), (7, 9, 10, 13), (10, 9, 7, 13), (5, 5, 5, 5), (11, 12, 6, 17), (1, 1, 4, 3)]:
        x = rng.standard_normal((h, w, 3)).astype(np.float32)
        assert np.array_equal(dr.resize_image_with_crop_or_pad(x, th, tw), T.crop_or_pad(torch.tensor(x)[None], th, tw)[0].numpy())


def test_dataset_batches_in_order_with_prefetch_thread(tmp_path):
    from Data_utils import data_reader as dr
    lst, frames = _write_dataset(tmp_path, n=5, h=20, w=30)
    ds = dr.dataset(lst, batch_size=2, crop_shape=[16, 36], num_epochs=1, augment=False, is_training=False, shuffle=False, prefetch=2)
    assert len(ds) == 5 and ds.get_max_steps() == 2                                               # drop_remainder
    batches = list(ds)
    assert len(batches) == 2
    l, r, g = batches[1]
    assert l.shape == (2, 16, 36, 3) and g.shape == (2, 16, 36, 1) and l.dtype == np.float32
    want = dr.resize_image_with_crop_or_pad(frames[2][0].astype(np.float32), 16, 36)             # centre crop rows, zero-pad cols
    assert np.array_equal(l[0], want)
    gt_cropped = (frames[3][2][:, :30, None] / 256.0).astype(np.float32)                           # gt cut to the image width first
    assert np.allclose(g[1], dr.resize_image_with_crop_or_pad(gt_cropped, 16, 36))
    with pytest.raises(NotImplementedError):
        dr.dataset(lst, augment=True)


def test_pfm_reader(tmp_path):
    from Data_utils import data_reader as dr
    a = np.arange(12, dtype=np.float32).reshape(3, 4)
    p = str(tmp_path / 'd.pfm')
    with open(p, 'wb') as f:
        f.write(b'Pf\n4 3\n-1.0\n'); f.write(np.flipud(a).astype('<f4').tobytes())
    d, scale = dr.readPFM(p)
    assert scale == 1.0 and d.shape == (3, 4, 1) and np.array_equal(d[..., 0], a)
    assert np.array_equal(dr.read_gt_from_disc(p)[..., 0], a)


# ---------------------------------------------------------------------------------------------------
# driver: command line and output files
# ---------------------------------------------------------------------------------------------------
def _driver():
    import importlib.util
    spec = importlib.util.spec_from_file_location('soa_driver', os.path.join(PKG, 'Stereo_Online_Adaptation.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_driver_command_line_equals_the_reference(tmp_path):
    """tests/golden/reference_driver_cli.json holds the argparse actions of the reference's Stereo_Online_Adaptation.py
    (captured by tests/golden/make_reference_golden.py running that script under the TF shim)."""
    import argparse
    import json
    g = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'reference_driver_cli.json')))['actions']
    mine = []
    for a in _driver().build_parser()._actions:
        if a.dest == 'help':
            continue
        mine.append({'flags': list(a.option_strings), 'dest': a.dest, 'required': bool(a.required), 'default': a.default,
                     'type': getattr(a.type, '__name__', None), 'nargs': a.nargs, 'choices': sorted(a.choices) if a.choices else None,
                     'store_true': isinstance(a, argparse._StoreTrueAction)})
    assert mine == g


class _FakeAdapt:
    def __init__(self):
        self.calls, self.reset_counter, self.fetch_counter, self.sample_distribution = [], 1, [2, 1, 0, 0, 0], np.array([0.5, 0, 0, 0, -0.25])

    def step(self, left, right, gt=None, want_disp_mask=0, prefetch=None):
        self.calls.append((float(left.mean()), want_disp_mask, prefetch is not None))
        k = len(self.calls)
        return {'loss': 0.1 * k, 'train_loss': 0.0, 'epe': 1.0 * k, 'bad3': 0.01 * k, 'blocks': [0], 'reset': False}


def test_driver_loop_and_output_files(tmp_path):
    import argparse
    import cv2
    d = _driver()
    out = tmp_path / 'out'; (out / 'disparities').mkdir(parents=True)
    args = argparse.Namespace(logDispStep=2, output=str(out))
    frames = [(np.full((1, 4, 6, 3), float(i), np.float32), np.zeros((1, 4, 6, 3), np.float32), np.zeros((1, 4, 6, 1), np.float32)) for i in range(3)]
    ad = _FakeAdapt()
    disp = np.linspace(0, 300, 24, dtype=np.float32).reshape(1, 4, 6, 1)
    epe, bad3, exec_time, step = d.run_loop(ad, frames, args, max_steps=3, get_disparity=lambda: disp, log=lambda s: None)
    assert step == 3 and epe == [1.0, 2.0, 3.0] and bad3 == [0.01, 0.02, 0.03]
    assert [c[0] for c in ad.calls] == [0.0, 1.0, 2.0]                                      # frames in order
    assert [c[1] for c in ad.calls] == [0b100000, 0, 0b100000]                             # disparity fetched on logDispStep
    assert [c[2] for c in ad.calls] == [True, True, False]                                 # next frame prefetched except at the end
    png = cv2.imread(str(out / 'disparities' / 'disparity_2.png'), cv2.IMREAD_UNCHANGED)
    assert png.dtype == np.uint16 and png.shape == (4, 6)
    assert np.array_equal(png, (np.clip(disp[0, ..., 0], 0, 256) * 256.0).astype(np.uint16))   # x256, clipped at MAX_DISP
    d.write_stats(str(out / 'stats.csv'), epe, bad3, 0.5, step, ad.reset_counter, 5, ad.fetch_counter, ad.sample_distribution)
    d.write_series(str(out / 'series.csv'), epe, bad3, 0.5, step)
    lines = open(str(out / 'stats.csv')).read().splitlines()
    assert lines[0] == 'Metrics,cumulative,average' and lines[1] == 'EPE,6.0,2.0'
    assert lines[3].startswith('time,0.5,') and lines[4].startswith('FPS,6.0') and lines[5] == '#resets,1'
    assert lines[6] == 'Blocks,0,1,2,3,4,final' and lines[7] == 'fetch_counter,2,1,0,0,0' and lines[8] == ',0.5,0.0,0.0,0.0,-0.25'
    s = open(str(out / 'series.csv')).read().splitlines()
    assert s[0] == 'Iteration,Time,EPE,bad3' and s[1] == '0,0.0,1.0,0.01' and s[3].startswith('2,0.3333')


class _FakeHandle:
    def __init__(self, h, w):
        self._h, self._w = h, w

    def numpy(self):
        return np.full((1, self._h, self._w, 1), 7.5, np.float32)


class _FakeNet:
    def __init__(self, h, w):
        self._d = [_FakeHandle(h, w) for _ in range(6)]

    def get_disparities(self):
        return self._d


class _FakeAdapt2(_FakeAdapt):
    def __init__(self, names):
        super().__init__()
        self.names, self.loaded = names, None

    def get_variable_names(self):
        return self.names

    def load_weights(self, w, strict=True):
        self.loaded, self.strict = dict(w), strict


def test_driver_main_end_to_end_with_a_stub_engine(tmp_path):
    """Everything of the driver except the CUDA engine: list file -> decoded, cropped frames -> checkpoint restore through the
    TF-free reader -> loop -> stats.csv / series.csv / disparity PNGs."""
    import cv2
    d = _driver()
    lst, frames = _write_dataset(tmp_path, n=4, h=20, w=30)
    names = ['model/gc-read-pyramid/conv1/weights', 'model/gc-read-pyramid/conv1/biases', 'model/context-1/weights']
    tfc.write_checkpoint(str(tmp_path / 'w.ckpt'), {k: v for k, v in _tensors().items() if 'conv1/' in k})
    out = tmp_path / 'run'; (out / 'disparities').mkdir(parents=True)
    cfg = os.path.join(PKG, 'block_config', 'MadNet_full.json')
    args = d.build_parser().parse_args(['-l', lst, '-o', str(out), '--weights', str(tmp_path / 'w.ckpt'), '--modelName', 'MADNet',
                                        '--blockConfig', cfg, '--mode', 'MAD', '--sampleMode', 'SEQUENTIAL', '--imageShape', '16', '32',
                                        '--logDispStep', '3'])
    made = {}

    def build(a, train_config):
        assert len(train_config) == 5 and a.imageShape == [16, 32]
        made['adapt'] = _FakeAdapt2(names)
        return _FakeNet(16, 32), made['adapt']

    d.main(args, build=build)
    ad = made['adapt']
    assert set(ad.loaded) == set(names[:2]) and ad.strict is False                 # only what the checkpoint holds
    assert len(ad.calls) == 4
    want0 = float(np.asarray(__import__('Data_utils.data_reader', fromlist=['x']).resize_image_with_crop_or_pad(
        frames[0][0].astype(np.float32), 16, 32)).mean())
    assert abs(ad.calls[0][0] - want0) < 1e-3
    stats = open(str(out / 'stats.csv')).read().splitlines()
    assert stats[1].startswith('EPE,10.0,2.5') and stats[6] == 'Blocks,0,1,2,3,4,final'
    assert len(open(str(out / 'series.csv')).read().splitlines()) == 5
    png = cv2.imread(str(out / 'disparities' / 'disparity_3.png'), cv2.IMREAD_UNCHANGED)
    assert png.shape == (16, 32) and int(png[0, 0]) == int(7.5 * 256)
    with pytest.raises(SystemExit):
        d.main(d.build_parser().parse_args(['-l', lst, '-o', str(out), '--weights', 'x', '--blockConfig', cfg, '--reprojectionScale', '2']), build=build)


# ---------------------------------------------------------------------------------------------------- continual adaptation
def _continual():
    import importlib
    return importlib.import_module('Stereo_Continual_Adaptation')


def test_continual_driver_command_line_equals_the_reference():
    """tests/golden/reference_continual_cli.json: the argparse actions of the reference's Stereo_Continual_Adaptation.py."""
    import argparse
    import json
    g = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'reference_continual_cli.json')))['actions']
    mine = []
    for a in _continual().build_parser()._actions:
        if a.dest == 'help':
            continue
        mine.append({'flags': list(a.option_strings), 'dest': a.dest, 'required': bool(a.required), 'default': a.default,
                     'type': getattr(a.type, '__name__', None), 'nargs': a.nargs, 'choices': sorted(a.choices) if a.choices else None,
                     'store_true': isinstance(a, argparse._StoreTrueAction)})
    assert mine == g


def test_proxy_loss_oracle_known_answers():
    """Losses/loss_factory.py:304-351 + :28-38: valid = !(proxy <= 0 | proxy >= 192); w * sum(valid*|d-p|) / sum(valid)."""
    import torch
    from oracle import tf1_ops as T
    d = torch.tensor([[[[1.0], [5.0], [7.0], [300.0]]]])
    p = torch.tensor([[[[2.0], [0.0], [4.0], [192.0]]]])                 # pixels 1 (proxy 0) and 3 (proxy 192) are invalid
    assert abs(float(T.proxy_loss(d, p, 0.1)) - 0.1 * (1.0 + 3.0) / 2.0) < 1e-7
    assert float(T.proxy_loss(d, torch.zeros_like(p), 0.1)) == 0.0     # no valid pixel (the reference: nan)
    d.requires_grad_(True)
    T.proxy_loss(d, p, 0.01).backward()
    assert torch.allclose(d.grad.ravel(), torch.tensor([-0.005, 0.0, 0.005, 0.0]))


def test_reader_with_proxies_and_continual_loop(tmp_path):
    """List format left;right;gt;proxy (continual_data_reader.py:55-78), 16-bit proxies / 256, and the driver loop with its
    output files around a stub adaptation object (EPE / D1 as :243-249)."""
    import cv2
    from Data_utils import data_reader
    rng = np.random.default_rng(3)
    lines = []
    for i in range(3):
        l = rng.integers(0, 255, (20, 30, 3), dtype=np.uint8); r = rng.integers(0, 255, (20, 30, 3), dtype=np.uint8)
        gt = (rng.uniform(1, 40, (20, 30)) * 256).astype(np.uint16); px = (rng.uniform(0, 60, (20, 34)) * 256).astype(np.uint16)
        names = [str(tmp_path / ('%s_%d.png' % (k, i))) for k in ('l', 'r', 'g', 'p')]
        for n, img in zip(names, (l, r, gt, px)):
            cv2.imwrite(n, img)
        lines.append(';'.join(names))
    lst = tmp_path / 'list.csv'
    lst.write_text('\n'.join(lines) + '\n')
    ds = data_reader.dataset(str(lst), batch_size=1, crop_shape=[16, 24], num_epochs=1, is_training=False, shuffle=False, proxies=True)
    batches = list(ds)
    assert len(batches) == 3 and [b.shape for b in batches[0][:4]] == [(1, 16, 24, 3), (1, 16, 24, 3), (1, 16, 24, 1), (1, 16, 24, 1)]
    px0 = cv2.imread(lines[0].split(';')[3], -1).astype(np.float32)[:, :30] / 256.0      # cut to the image width, centre crop
    assert np.allclose(batches[0][3][0, :, :, 0], px0[2:18, 3:27])
    assert float(batches[0][4][0, 0]) == 30.0

    C = _continual()

    class Stub:
        fetch_counter = [0] * 5

        def __init__(self):
            self.calls = []

        def step(self, left, right, gt=None, want_disp_mask=0, prefetch=None, proxy=None):
            assert proxy is not None and proxy.shape == (1, 16, 24, 1) and want_disp_mask == 0b100000
            self.calls.append(float(proxy.mean()))
            return {'loss': 0.1}

    out = tmp_path / 'out'
    (out / 'disparities').mkdir(parents=True)
    args = C.build_parser().parse_args(['-l', str(lst), '-o', str(out), '--weights', 'w', '--blockConfig', 'b', '--logDispStep', '2'])
    stub = Stub()
    disp = np.full((1, 16, 24, 1), 7.6, np.float32)
    avg, d1, steps, _ = C.run_loop(stub, batches, args, get_disparity=lambda: disp, log=lambda *a: None)
    assert steps == 3 and len(stub.calls) == 3
    e0, o0 = C.frame_errors(disp[-1], batches[0][2][-1])
    assert avg[0] == e0 and d1[0] == o0 and 0.0 <= o0 <= 100.0
    C.write_outputs(args, avg, d1)
    assert (out / 'overall.csv').read_text().splitlines()[0] == 'EPE\tD1'
    assert (out / 'series.csv').read_text().splitlines()[1].startswith('0 & ')
    assert (out / 'histogram.csv').exists()
    saved = cv2.imread(str(out / 'disparities' / 'disparity_2.png'), -1)
    assert saved.dtype == np.uint16 and int(saved[0, 0]) == 7 * 256                          # cast to uint16 first, then x 256 (:279-280)
