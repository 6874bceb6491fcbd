"""mvae-b200: B200-native (sm_100a) implementation of the MVAE training-step hot path of
mhw32/multimodal-vae-public behind that project's Python surface.  See DESIGN.md."""
__version__ = "0.1.0"
