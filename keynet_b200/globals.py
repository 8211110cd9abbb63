"""Process-wide switches (reference: keynet/globals.py).  Only `verbose` is consumed by the keyed path."""

GLOBAL = {'VERBOSE': False}


def backend():
    return 'b200'


def verbose(b=None):
    if b is not None:
        GLOBAL['VERBOSE'] = bool(b)
    return GLOBAL['VERBOSE']
