"""Drop-in for the reference `Nets` package (Nets/__init__.py:1-13): same factory, same names."""
from Nets import DispNet as _DispNet
from Nets import MadNet as _MadNet

STEREO_FACTORY = {
    _DispNet.DispNet._netName: _DispNet.DispNet,
    _MadNet.MadNet._netName: _MadNet.MadNet,
}


def get_stereo_net(name, args):
    if name not in STEREO_FACTORY:
        raise Exception('Unrecognized network name: {}'.format(name))
    return STEREO_FACTORY[name](**args)
