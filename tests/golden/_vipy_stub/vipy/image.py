class Image(object):
    def __init__(self, *a, **k):
        raise NotImplementedError('vipy.image.Image is not available in the stub (image file I/O is out of scope)')
