"""tf.keras placeholder: just enough for the reference's model classes to be *defined*.

TEST INFRASTRUCTURE ONLY (see ../__init__.py).  Backbones are out of scope; nothing here
computes anything.  `Model.__init__` accepts and ignores Keras kwargs so that
`class RetinaNet(tf.keras.Model)` can be instantiated by the golden-vector generator.
"""
import types


class Model:
    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)


class _Layer:
    def __init__(self, *args, **kwargs):
        pass


def _missing(*args, **kwargs):
    raise NotImplementedError("tf.keras graph building is out of scope for the stub")


layers = types.ModuleType("tensorflow.keras.layers")
layers.Layer = _Layer
for _n in ("Conv2D", "BatchNormalization", "UpSampling2D", "MaxPool2D", "ReLU",
           "Add", "Concatenate", "Dense", "Input", "Conv2DTranspose"):
    setattr(layers, _n, _missing)

applications = types.ModuleType("tensorflow.keras.applications")
Input = _missing
