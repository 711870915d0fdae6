from .biggan import BigGAN


def __getattr__(name):
    if name == "StyleGAN2":
        from .stylegan2 import StyleGAN2
        return StyleGAN2
    raise AttributeError(name)
