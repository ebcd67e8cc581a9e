def expect(*args, **kwargs):
    return None
