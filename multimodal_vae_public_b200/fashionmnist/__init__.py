"""Drop-in surface for the reference's ``fashionmnist/`` experiment."""
