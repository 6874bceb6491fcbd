"""``fashionmnist/sample.py`` surface: the same four generation modes as mnist/sample.py:66-112 (the conv decoders return
[n,1,28,28] logits)."""
from ..mnist.sample import generate  # noqa: F401
