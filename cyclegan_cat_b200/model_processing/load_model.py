"""Config loading -- mirrors `model_processing/load_model.py:9-32` of the reference.

`bunch.Bunch` is not installed here; `Bunch` below is the same idea (a dict whose
top-level keys are also attributes; nested mappings stay plain dicts, which is
what `model.py` relies on: `model_config.generator["type"]`, `train_config.g_opt`).
"""
import yaml


class Bunch(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v

    def __delattr__(self, k):
        del self[k]


def yaml2namespace(yaml_path: str) -> Bunch:
    with open(yaml_path, 'r') as f:
        model_config_dict = yaml.load(f, yaml.FullLoader)
    return Bunch(model_config_dict)


def namespace2yaml(yaml_path: str, namespace: Bunch):
    with open(yaml_path, 'w') as f:
        yaml.dump(dict(namespace), f)
