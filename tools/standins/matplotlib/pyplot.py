def imsave(*args, **kwargs):
    return None
