"""CelebA-19 flavour (image + 18 single-attribute modalities): drop-in ``model`` / ``train`` surfaces."""
