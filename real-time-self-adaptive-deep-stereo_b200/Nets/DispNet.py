"""DispNet-C on the B200 engine — host-side mirror of the reference class (Nets/DispNet.py:9-152).

The DispNet engine graph is not wired into libmadstereo in this build; constructing it raises loudly
(there is no fallback path).  Argument validation follows the reference (DispNet.py:23-37).
"""
from Nets import Stereo_net

MAX_DISP = 40


class DispNet(Stereo_net.StereoNet):
    _valid_args = [
        ("left_img", "device buffer for left image batch"),
        ("right_img", "device buffer for right image batch"),
        ("correlation", "flag to enable the use of the correlation layer"),
    ] + Stereo_net.StereoNet._valid_args
    _netName = "Dispnet"

    def __init__(self, **kwargs):
        super(DispNet, self).__init__(**kwargs)

    def _validate_args(self, args):
        super(DispNet, self)._validate_args(args)
        if ("left_img" not in args) or ("right_img" not in args):
            raise Exception('Missing input op for left and right images')
        if "correlation" not in args:
            print('WARNING: Correlation unspecified, setting to True')
            args['correlation'] = True
        return args

    def _preprocess_inputs(self, args):
        self._left_input_batch = args['left_img']
        self._right_input_batch = args['right_img']

    def _build_network(self, args):
        raise NotImplementedError('DispNet graph is not available in this build of libmadstereo')
