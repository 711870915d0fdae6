"""Interface of a target transformation (reference: pix2latent/transform/base_transform.py:1-31): callable on a batch
of images with a batch of parameters, with default / identity parameters and explicit forward / inverse maps.
Concrete transforms override what they support; everything else reports which method of which class is missing."""


def _missing(name, what):
    def method(self, *args, **kwargs):
        raise NotImplementedError("{}.{} ({}) is not implemented".format(type(self).__name__, name, what))
    method.__name__ = name
    method.__doc__ = what
    return method


class TransformTemplate():
    pass


for _name, _what in (("__call__", "apply the transformation to the images"),
                     ("get_default_param", "parameter the search starts from"),
                     ("get_identity_param", "parameter of the identity transformation"),
                     ("transform", "forward map"),
                     ("invert_transform", "inverse map")):
    setattr(TransformTemplate, _name, _missing(_name, _what))
del _name, _what
