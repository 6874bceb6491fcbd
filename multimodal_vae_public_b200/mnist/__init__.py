"""Drop-in surface for the reference's ``mnist/`` experiment: ``model.MVAE`` & co, ``train.elbo_loss`` & co."""
