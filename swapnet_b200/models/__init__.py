"""Drop-in `models` package: the plugin surface of /root/reference/models/__init__.py:5-44
(`create_model`, `get_options_modifier`, `find_model_using_name`, `BaseModel`) backed by the B200
engines.  `dropin/models` re-exports this package under the top-level name `models` so that the
reference's train.py / options code import it unchanged.
"""
import importlib

from .base_model import BaseModel

__all__ = ["BaseModel", "find_model_using_name", "get_options_modifier", "create_model"]


def find_model_using_name(model_name):
    """`<name>` -> class `<Name>Model` (case-insensitive) defined in `<name>_model.py`."""
    module = importlib.import_module(f"{__name__}.{model_name}_model")
    wanted = (model_name.replace("_", "") + "model").lower()
    for attr, cls in vars(module).items():
        if attr.lower() == wanted and isinstance(cls, type) and issubclass(cls, BaseModel):
            return cls
    print(f"In {model_name}_model.py, there should be a subclass of BaseModel with class name that matches "
          f"{wanted} in lowercase.")
    exit(0)  # the reference's error convention (models/__init__.py:20-22)


def get_options_modifier(model_name):
    return find_model_using_name(model_name).modify_commandline_options


def create_model(opt):
    instance = find_model_using_name(opt.model)(opt)
    print("model [%s] was created" % type(instance).__name__)
    return instance
