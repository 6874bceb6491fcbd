"""Drop-in surface for the reference's ``celeba/`` experiment."""
