from .injector import StragglerInjector

__all__ = ["StragglerInjector"]
