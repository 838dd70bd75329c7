"""The reference-side weight bridge (scripts/keras_weights_bridge.py) on stand-in layers with Keras' interface (class name,
.name, get_weights / set_weights): Keras itself is absent here.  The walk must recover the tensors of model.generate_model()'s
graph (model.py:30-56) whatever order the two heads' layers come in, and round-trip them."""
import importlib.util
import os

import numpy as np

from oracle import ref_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("bridge", os.path.join(ROOT, "scripts", "keras_weights_bridge.py"))
bridge = importlib.util.module_from_spec(spec)
spec.loader.exec_module(bridge)


def _layer(kind, weights, name="layer"):
    cls = type(kind, (), {"get_weights": lambda self: [w.copy() for w in self._w],
                          "set_weights": lambda self, ws: setattr(self, "_w", [np.asarray(w) for w in ws])})
    obj = cls()
    obj._w, obj.name = [np.asarray(w) for w in weights], name
    return obj


def _fake_model(w, blocks, heads_interleaved):
    bn = lambda t: _layer("BatchNormalization", [t[0], t[1], t[2], t[3]])
    layers = [_layer("InputLayer", []), _layer("Conv2D", [w["stem.w"]]), bn(w["stem.bn"]), _layer("LeakyReLU", [])]
    for i in range(blocks):
        layers += [_layer("Conv2D", [w[f"res{i}.w1"]]), bn(w[f"res{i}.bn1"]), _layer("LeakyReLU", []),
                   _layer("Conv2D", [w[f"res{i}.w2"]]), bn(w[f"res{i}.bn2"]), _layer("Add", []), _layer("LeakyReLU", [])]
    value = [_layer("Conv2D", [w["value.conv"]]), bn(w["value.bn"]), _layer("LeakyReLU", []), _layer("Flatten", []),
             _layer("Dense", [w["value.fc1"]], "dense"), _layer("LeakyReLU", []), _layer("Dense", [w["value.fc2"]], "value_head")]
    policy = [_layer("Conv2D", [w["policy.conv"]]), bn(w["policy.bn"]), _layer("LeakyReLU", []), _layer("Flatten", []),
              _layer("Dense", [w["policy.fc"]], "policy_head")]
    if heads_interleaved:                      # a topological order may interleave the two heads
        tail = [x for pair in zip(value, policy) for x in pair] + value[len(policy):]
    else:
        tail = value + policy
    return type("Model", (), {"layers": layers + tail})()


def test_export_recovers_every_tensor_and_import_round_trips():
    for filters, blocks, inter in ((48, 4, False), (16, 2, True), (8, 0, True)):
        w = ref_model.init_weights(filters, blocks, seed=filters, random_bn=True)
        out = bridge.export_weights(_fake_model(w, blocks, inter))
        assert set(out) == set(w)
        for k in w:
            assert np.array_equal(out[k].reshape(np.asarray(w[k]).shape), w[k]), k
        blank = {k: np.zeros_like(v) for k, v in w.items()}
        model = _fake_model(blank, blocks, inter)
        bridge.import_weights(model, out)
        back = bridge.export_weights(model)
        assert all(np.array_equal(back[k], out[k]) for k in out)
