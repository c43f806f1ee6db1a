"""Host-side mirror of the reference's abstract StereoNet (Nets/Stereo_net.py:6-222).

Same construction protocol (_validate_args -> _preprocess_inputs -> _build_network), same public getters.
What differs by design: layers / disparities / variables are handles onto device buffers owned by the C++
engine (torch CUDA tensors used purely as containers) instead of TF graph nodes, and the graph itself is
executed by libmadstereo, not by a TF session.
"""
import abc
from collections import OrderedDict


class Variable(object):
    """Stand-in for a tf.Variable: name + views of weight / gradient / momentum inside the engine arenas."""

    def __init__(self, name, engine, key):
        self.name = name + ':0'
        self.op_name = name
        self._engine, self._key = engine, key

    @property
    def shape(self):
        return tuple(self.value().shape)

    def value(self):
        return self._engine.param_views()[self._key]

    def grad(self):
        return self._engine.param_views(self._engine.grads)[self._key]

    def momentum(self):
        return self._engine.param_views(self._engine.momentum)[self._key]

    def __repr__(self):
        return "<Variable '%s' shape=%s>" % (self.name, self.shape)


class LayerHandle(object):
    """Stand-in for a TF op output: lazily resolves to the engine tensor of that name."""

    def __init__(self, net, name, shape, tensor_name=None):
        self._net, self.name, self.shape = net, name, tuple(shape)
        self._tensor_name = tensor_name or name

    def get_shape(self):
        return self.shape

    def tensor(self):
        return self._net.engine.tensor(self._tensor_name)

    def numpy(self):
        return self.tensor().detach().cpu().numpy().copy()

    def __repr__(self):
        return "<Layer '%s' shape=%s>" % (self.name, self.shape)


class StereoNet(object):
    __metaclass__ = abc.ABCMeta
    _valid_args = [
        ("split_layer", "name of the layer where the network will be splitted"),
        ("sequence", "flag to use network on a video sequence instead of on single images"),
        ("train_portion", "one among 'BEGIN' or 'END' specify which portion of the network will be trained"),
        ("is_training", "boolean or placeholder to specify if the network is in train or inference mode"),
    ]
    _netName = "stereoNet"

    @classmethod
    def getPossibleArsg(cls):
        return cls._valid_args

    def __init__(self, **kwargs):
        self._layers = OrderedDict()
        self._disparities = []
        self._placeholders = []
        self._trainable_variables = OrderedDict()
        self._layer_to_var = {}
        self.engine = None
        print('=' * 50)
        print('Starting Creation of {}'.format(self._netName))
        print('=' * 50)
        args = self._validate_args(kwargs)
        print('Args Validated, setting up graph')
        self._preprocess_inputs(args)
        print('Meta op to preprocess data created')
        self._build_network(args)
        print('Network ready')
        print('=' * 50)

    # -- helpers used by subclasses ---------------------------------------------------------------
    def _add_to_layers(self, name, handle, variables=()):
        self._layers[name] = handle
        self._layer_to_var[name] = list(variables)
        if self._train_beginning or self._split_layers_list != [None]:
            for v in variables:
                self._trainable_variables[v] = True

    def __str__(self):
        ss = ""
        for k, l in self._layers.items():
            if l in self._disparities:
                ss += "Prediction Layer {}: {}\n".format(k, str(l.shape))
            else:
                ss += "Layer {}: {}\n".format(k, str(l.shape))
        return ss

    def __repr__(self):
        return self.__str__()

    def __getitem__(self, key):
        return self._layers[key]

    @abc.abstractmethod
    def _preprocess_inputs(self, args):
        pass

    @abc.abstractmethod
    def _build_network(self, args):
        pass

    @abc.abstractmethod
    def _validate_args(self, args):
        portion_options = ['BEGIN', 'END']
        if 'split_layers' not in args:
            print('WARNING: no split points selected, the network will flow without interruption')
            args['split_layers'] = [None]
        if 'train_portion' not in args:
            print('WARNING: train_portion not specified, using default END')
            args['train_portion'] = 'END' if args['split_layers'] != [None] else 'BEGIN'
        elif args['train_portion'] not in portion_options:
            raise Exception('Invalid portion options {}'.format(args['train_portion']))
        if 'sequence' not in args:
            print('WARNING: sequence flag not setted, configuring the network for single image adaptation')
            args['sequence'] = False
        if 'is_training' not in args:
            print('WARNING: flag for trainign not setted, using default False')
            args['is_training'] = False
        if args['split_layers'] != [None]:
            raise Exception('split_layers other than [None] is not supported by the B200 engine '
                            '(the adaptation drivers always pass [None], Stereo_Online_Adaptation.py:57)')
        self._split_layers_list = args['split_layers']
        self._train_beginning = (args['train_portion'] == 'BEGIN')
        self._sequence = args['sequence']
        self._isTraining = False

    # -- public API (Nets/Stereo_net.py:166-222) ----------------------------------------------------
    def get_placeholders(self):
        return self._placeholders

    def get_placeholder(self, name):
        raise Exception('Unable to find placeholder for layer {}'.format(name + '_placeholder'))

    def get_all_layers(self):
        return self._layers

    def get_layers_names(self):
        return self._layers.keys()

    def get_disparities(self):
        return self._disparities

    def get_trainable_variables(self):
        return list(self._trainable_variables.keys())

    def get_variables(self, layer_name):
        if layer_name in self._layers and layer_name not in self._layer_to_var:
            return []
        return self._layer_to_var[layer_name]
