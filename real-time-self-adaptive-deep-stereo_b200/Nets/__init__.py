"""Drop-in for the reference `Nets` package (Nets/__init__.py:1-13): same factory name, keys and error."""
from Nets import DispNet as _DispNet
from Nets import MadNet as _MadNet

# keyed by the class's _netName, like the reference: 'Dispnet', 'MADNet' (argparse choices of the drivers)
STEREO_FACTORY = {cls._netName: cls for cls in (_DispNet.DispNet, _MadNet.MadNet)}


def get_stereo_net(name, args):
    """Build the network called `name` from the argument dict the drivers assemble (Stereo_Online_Adaptation.py:55-62)."""
    try:
        builder = STEREO_FACTORY[name]
    except KeyError:
        raise Exception('Unrecognized network name: {}'.format(name))
    return builder(**args)
