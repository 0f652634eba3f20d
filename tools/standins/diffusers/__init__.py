"""Stand-in for the absent third-party `diffusers` package: train_JPDVT.py:27 imports AutoencoderKL and never uses it."""
