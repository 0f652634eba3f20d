def imsave(*args, **kwargs):  # side effect dropped on purpose
    return None
