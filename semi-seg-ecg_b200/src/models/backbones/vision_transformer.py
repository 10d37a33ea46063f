"""ViT backbones are outside the accelerated hot path (out of scope -- the
north star names the convolutional segmentor).  The registry names exist so that configs fail with
a clear message instead of a KeyError."""

__all__ = ["vit_tiny", "vit_small", "vit_base"]


def _unsupported(name):
    def ctor(*args, **kwargs):
        raise NotImplementedError(
            f"{name}: the B200 hot path accelerates the 1-D ResNet + FCNHead segmentor only; "
            "ViT backbones (reference src/models/backbones/vision_transformer.py) are out of scope.")
    ctor.__name__ = name
    return ctor


vit_tiny = _unsupported("vit_tiny")
vit_small = _unsupported("vit_small")
vit_base = _unsupported("vit_base")
