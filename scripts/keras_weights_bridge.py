"""Reference-side bridge: run THIS in the reference's own TensorFlow / Keras environment to move weights
between a Keras model built by the reference's model.generate_model() (model.py:30-70) and the .npz layout chessbot_b200
loads (tensor names and layouts of include/chessbot_b200.h: cb_net_set_tensor / cb_train_set_tensor).

    python keras_weights_bridge.py export checkpoints/keras/model.keras weights.npz     # Keras -> chessbot_b200
    python keras_weights_bridge.py import weights.npz checkpoints/keras/model.keras     # chessbot_b200 -> Keras (in place)

Keras layouts are kept as they are (Conv2D kernels HWIO, Dense kernels [in, out], BatchNormalization as the rows gamma, beta,
moving_mean, moving_variance), so the mapping is purely by layer: the 3x3 convolutions and the C-channel normalisations in
graph order are stem, res0.w1 / bn1, res0.w2 / bn2, ...; the 1x1 convolutions with 1 / 2 filters and the 1- / 2-channel
normalisations are the value / policy heads; Dense(C) is value.fc1, the layers named "value_head" / "policy_head" the outputs.

NOT RUN in the build container of chessbot_b200 (TensorFlow and Keras are absent there); the layer walk below uses only
`model.layers`, `layer.get_weights()` / `set_weights()` and class names, which are stable across Keras 2 and 3.
The consumer side (`chessbot_b200.model.load_weights`, `Trainer.set_weights`) is covered by the repo's tests."""
import sys

import numpy as np


def _walk(model):
    """-> (ordered slots [(name, layer)], ...) following the structure of model.py:30-56"""
    convs3, bns_c, slots = [], [], {}
    filters = None
    for layer in model.layers:
        kind = type(layer).__name__
        if kind == "Conv2D":
            k = layer.get_weights()[0]
            if k.shape[0] == 3:
                convs3.append(layer)
                filters = k.shape[-1]
            elif k.shape[-1] == 1:
                slots["value.conv"] = layer
            elif k.shape[-1] == 2:
                slots["policy.conv"] = layer
        elif kind == "BatchNormalization":
            c = layer.get_weights()[0].shape[0]
            if c == 1:
                slots["value.bn"] = layer
            elif c == 2:
                slots["policy.bn"] = layer
            else:
                bns_c.append(layer)
        elif kind == "Dense":
            if layer.name == "value_head":
                slots["value.fc2"] = layer
            elif layer.name == "policy_head":
                slots["policy.fc"] = layer
            else:
                slots["value.fc1"] = layer
    if len(convs3) != len(bns_c) or len(convs3) % 2 != 1:
        raise ValueError(f"not the graph of model.generate_model(): {len(convs3)} 3x3 convolutions, {len(bns_c)} wide normalisations")
    slots["stem.w"], slots["stem.bn"] = convs3[0], bns_c[0]
    for i in range((len(convs3) - 1) // 2):
        slots[f"res{i}.w1"], slots[f"res{i}.bn1"] = convs3[1 + 2 * i], bns_c[1 + 2 * i]
        slots[f"res{i}.w2"], slots[f"res{i}.bn2"] = convs3[2 + 2 * i], bns_c[2 + 2 * i]
    return slots, filters, (len(convs3) - 1) // 2


def export_weights(model):
    slots, filters, blocks = _walk(model)
    out = {}
    for name, layer in slots.items():
        w = layer.get_weights()
        out[name] = np.stack(w).astype(np.float32) if name.endswith(("bn", "bn1", "bn2")) else np.asarray(w[0], np.float32)
    return out


def import_weights(model, tensors):
    slots, _, _ = _walk(model)
    for name, layer in slots.items():
        t = np.asarray(tensors[name], np.float32)
        if name.endswith(("bn", "bn1", "bn2")):
            layer.set_weights([t[0], t[1], t[2], t[3]])
        else:
            layer.set_weights([t.reshape(layer.get_weights()[0].shape)])


if __name__ == "__main__":
    import keras
    mode, src, dst = sys.argv[1:4]
    if mode == "export":
        np.savez(dst, **export_weights(keras.models.load_model(src)))
    elif mode == "import":
        model = keras.models.load_model(dst)
        with np.load(src) as f:
            import_weights(model, {k: f[k] for k in f.files})
        model.save(dst)
    else:
        raise SystemExit(__doc__)
