"""`DreamLLMConfig` with the reference's plugin-registration contract (SURVEY.md §8b "Plugin registration / instantiation").

Mirror of omni/models/dreamllm/configuration_dreamllm.py:64-278 without the omegaconf / hydra dependency:

* the LLaMA hyper-parameters the decoder reads (:168-214), `special_tokens2ids_dict`, `plugins_init_kwargs`, `plugins_type`,
  `loss_weight_lm = 1.0`, `loss_weight_vm = 10.0`, `loss_scale_schedule`, `log_attentions`, `log_hidden_states` (:215-223);
* `update_plugins(ConfigAndInitKwargs(_class_=Cls, _name_=str, _plugin_type_=str, **init_kwargs))` stores
  `{"_target_": "module.Cls", **init_kwargs}` under `plugins_init_kwargs[name]` (:237-255) — swapping an implementation is changing
  `_class_` (or the `_target_` string of a saved config.json);
* `update_special_tokens2ids_dict` (:225-235), `reset_plugins_init_kwargs` (:274-278);
* `save_pretrained` / `from_pretrained` read and write a plain `config.json` (the PretrainedConfig file format, so configs exported by the
  reference load here: unknown keys are kept as attributes).

`deep_instantiate` restates omni/config/instantiate.py:86-136 for plain dict / list containers.
"""
from __future__ import annotations

import copy
import importlib
import json
import os
from pydoc import locate

CLASS_KEY, NAME_KEY, PLUGIN_TYPE_KEY = "_class_", "_name_", "_plugin_type_"       # configuration_dreamllm.py:25-27
CONFIG_NAME = "config.json"


def ConfigAndInitKwargs(**kwargs) -> dict:
    """The reference's `ConfigAndInitKwargs` is a TypeAlias of dict (:44); config files call it like a constructor
    (projects/dreamllm/configs/common.py:12-56)."""
    return dict(**kwargs)


def _target_string(cls) -> str:
    if isinstance(cls, str):
        return cls
    return cls.__module__ + "." + cls.__qualname__


def _locate(name: str):
    obj = locate(name)
    if obj is None:                       # pydoc.locate swallows ImportError of the leaf module: retry for a real message
        mod, _, attr = name.rpartition(".")
        obj = getattr(importlib.import_module(mod), attr)
    return obj


def deep_instantiate(cfg):
    """Recursively build objects described by `{"_target_": "pkg.mod.Class" | callable, **kwargs}` (omni/config/instantiate.py:86-136)."""
    if isinstance(cfg, (list, tuple)):
        return [deep_instantiate(x) for x in cfg]
    if isinstance(cfg, dict):
        if "_target_" in cfg:
            kw = {k: deep_instantiate(v) for k, v in cfg.items()}
            cls = kw.pop("_target_")
            if isinstance(cls, str):
                name = cls
                cls = _locate(name)
                assert cls is not None, name
            assert callable(cls), f"_target_ {cls} does not define a callable object"
            return cls(**kw)
        return {k: deep_instantiate(v) for k, v in cfg.items()}
    return cfg


class DreamLLMConfig:
    model_type = "dreamllm"
    keys_to_ignore_at_inference = ["past_key_values"]

    def __init__(self, vocab_size=32000, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32,
                 num_key_value_heads=None, hidden_act="silu", max_position_embeddings=2048, initializer_range=0.02, rms_norm_eps=1e-6,
                 use_cache=True, pad_token_id=None, bos_token_id=1, eos_token_id=2, pretraining_tp=1, tie_word_embeddings=False,
                 rope_theta=10000.0, rope_scaling=None, attention_bias=False, special_tokens2ids_dict=None, plugins_init_kwargs=None,
                 plugins_type=None, loss_weight_lm=1.0, loss_weight_vm=10.0, loss_scale_schedule="none", log_attentions=False,
                 log_hidden_states=False, **kwargs):
        llama = dict(vocab_size=vocab_size, max_position_embeddings=max_position_embeddings, hidden_size=hidden_size,
                     intermediate_size=intermediate_size, num_hidden_layers=num_hidden_layers, num_attention_heads=num_attention_heads,
                     num_key_value_heads=num_attention_heads if num_key_value_heads is None else num_key_value_heads,
                     hidden_act=hidden_act, initializer_range=initializer_range, rms_norm_eps=rms_norm_eps, pretraining_tp=pretraining_tp,
                     use_cache=use_cache, rope_theta=rope_theta, rope_scaling=rope_scaling, attention_bias=attention_bias,
                     pad_token_id=pad_token_id, bos_token_id=bos_token_id, eos_token_id=eos_token_id,
                     tie_word_embeddings=tie_word_embeddings)
        dream = dict(  # the reference's mutable `{}` defaults are shared between instances; fresh dicts here
            special_tokens2ids_dict={} if special_tokens2ids_dict is None else special_tokens2ids_dict,
            plugins_init_kwargs={} if plugins_init_kwargs is None else plugins_init_kwargs,
            plugins_type={} if plugins_type is None else plugins_type,
            loss_weight_lm=loss_weight_lm, loss_weight_vm=loss_weight_vm, loss_scale_schedule=loss_scale_schedule,
            log_attentions=log_attentions, log_hidden_states=log_hidden_states)
        for name, value in {**llama, **dream}.items():
            setattr(self, name, value)
        self._rope_scaling_validation()
        for k, v in kwargs.items():                     # PretrainedConfig keeps unknown kwargs as attributes
            setattr(self, k, v)

    @classmethod
    def vicuna_7b(cls, **kw):
        base = dict(vocab_size=32008, hidden_size=4096, intermediate_size=11008, num_hidden_layers=32, num_attention_heads=32)
        base.update(kw)                      # e.g. num_hidden_layers=2 for a reduced-depth dev run
        return cls(**base)

    # ---------------------------------------------------------------------------------------------- reference methods
    def update_special_tokens2ids_dict(self, tokens_dict: dict, tokenizer):
        """:225-235 — `{"additional_special_tokens": ["<im_start>", ...], "bos_token": "<s>"}`: a list value becomes a nested
        `{token: id}` table under its key, a single token is stored under the token string itself."""
        table = self.special_tokens2ids_dict
        for key, value in tokens_dict.items():
            if isinstance(value, list):
                table.setdefault(key, {}).update(zip(value, tokenizer.convert_tokens_to_ids(value)))
            else:
                table[value] = tokenizer.convert_tokens_to_ids(value)

    def update_plugins(self, init_kwargs: dict) -> str:
        """:237-255."""
        init_kwargs = dict(init_kwargs)
        cls = init_kwargs.pop(CLASS_KEY, None)
        name = init_kwargs.pop(NAME_KEY, None)
        plugin_type = init_kwargs.pop(PLUGIN_TYPE_KEY, None)
        assert cls is not None and name is not None and plugin_type is not None, \
            f"`init_kwargs` must have `{CLASS_KEY}`, `{NAME_KEY}` and `{PLUGIN_TYPE_KEY}` fields"
        lazy_init = {"_target_": _target_string(cls), **copy.deepcopy(init_kwargs)}
        if name not in self.plugins_init_kwargs.keys():
            self.plugins_init_kwargs[name] = lazy_init
        else:
            self.plugins_init_kwargs[name].update(lazy_init)
        self.plugins_type[name] = plugin_type
        return name

    def _rope_scaling_validation(self):
        """:257-272 — None, or exactly {"type": "linear" | "dynamic", "factor": float > 1}; same messages as the reference."""
        rs = self.rope_scaling
        if rs is None:
            return
        if not (isinstance(rs, dict) and len(rs) == 2):
            raise ValueError(f"`rope_scaling` must be a dictionary with with two fields, `type` and `factor`, got {rs}")
        kind, factor = rs.get("type"), rs.get("factor")
        if kind not in ("linear", "dynamic"):
            raise ValueError(f"`rope_scaling`'s type field must be one of ['linear', 'dynamic'], got {kind}")
        if not (isinstance(factor, float) and factor > 1.0):
            raise ValueError(f"`rope_scaling`'s factor field must be an float > 1, got {factor}")

    def reset_plugins_init_kwargs(self, pretrained_plugin_model_name_or_path: str = None):
        """:274-278."""
        for plugin_name in self.plugins_init_kwargs.keys():
            self.plugins_init_kwargs[plugin_name]["pretrained_model_name_or_path"] = pretrained_plugin_model_name_or_path

    # ---------------------------------------------------------------------------------------------- (de)serialisation
    def to_dict(self) -> dict:
        d = {k: copy.deepcopy(v) for k, v in self.__dict__.items() if not k.startswith("_")}
        d["model_type"] = self.model_type
        return d

    def to_json_string(self) -> str:
        return json.dumps(self.to_dict(), indent=2, sort_keys=True) + "\n"

    def save_pretrained(self, save_directory: str):
        os.makedirs(save_directory, exist_ok=True)
        with open(os.path.join(save_directory, CONFIG_NAME), "w") as f:
            f.write(self.to_json_string())

    @classmethod
    def from_dict(cls, config_dict: dict, **kwargs):
        d = dict(config_dict)
        d.pop("model_type", None)
        d.pop("transformers_version", None)
        rp = d.pop("rope_parameters", None)            # transformers >= 5 spelling of rope_theta / rope_scaling
        if isinstance(rp, dict):
            d.setdefault("rope_theta", rp.get("rope_theta", 10000.0))
            if rp.get("rope_type", "default") != "default":
                d.setdefault("rope_scaling", {"type": rp["rope_type"], "factor": rp.get("factor")})
        d.update(kwargs)
        return cls(**d)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, return_unused_kwargs: bool = False, **kwargs):
        path = pretrained_model_name_or_path
        f = os.path.join(path, CONFIG_NAME) if os.path.isdir(path) else path
        if not os.path.isfile(f):
            raise OSError(f"no {CONFIG_NAME} under {path!r} (hub ids cannot be resolved: this build has no network)")
        with open(f) as fh:
            cfg = cls.from_dict(json.load(fh))
        return (cfg, kwargs) if return_unused_kwargs else cfg

    def __repr__(self):
        return f"{self.__class__.__name__} {self.to_json_string()}"
