"""Host-side mirrors of the reference's `model` package for the inference hot path
(model/dit.py, model/dpmsolver.py, model/autoencoder.py, model/attention)."""
