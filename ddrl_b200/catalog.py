"""``ModelCatalog.register_custom_model`` boundary (models/__init__.py:7-13).

When ``ray`` is importable the registration is forwarded to ``ray.rllib.models.ModelCatalog`` so that
``config['model']['custom_model'] = 'ffn' | 'gnn' | 'cup' | 'fc_glorot_uniform_init'`` resolves to the B200 classes
inside an unmodified ``tune.run("PPO", ...)`` (train_experiment_1_architecture_on_flat.py:138,201-211).  Without ray
(this container) a local registry with the same two calls is used by the in-repo learner and the tests."""
from __future__ import annotations

from typing import Any, Callable, Dict

_REGISTRY: Dict[str, Any] = {}

try:  # pragma: no cover - ray is absent in the build container
    from ray.rllib.models import ModelCatalog as _RayCatalog  # type: ignore
except Exception:
    _RayCatalog = None


class ModelCatalog:
    @staticmethod
    def register_custom_model(model_name: str, model_class: Callable) -> None:
        _REGISTRY[model_name] = model_class
        if _RayCatalog is not None:
            _RayCatalog.register_custom_model(model_name, model_class)

    @staticmethod
    def get_custom_model(model_name: str):
        if model_name not in _REGISTRY:
            raise KeyError(f"custom model {model_name!r} is not registered (known: {sorted(_REGISTRY)})")
        return _REGISTRY[model_name]

    @staticmethod
    def get_model_v2(obs_space, action_space, num_outputs: int, model_config: dict, framework: str = "torch",
                     name: str = "default_model", **kw):
        """Same call RLlib's policy builder makes; resolves ``model_config['custom_model']``."""
        cls = ModelCatalog.get_custom_model(model_config["custom_model"])
        return cls(obs_space, action_space, num_outputs, model_config, name)


def register_custom_model(model_name: str, model_class: Callable) -> None:
    ModelCatalog.register_custom_model(model_name, model_class)
