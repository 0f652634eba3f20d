class AutoencoderKL:  # imported by train_JPDVT.py:27, never instantiated
    pass
