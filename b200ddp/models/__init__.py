"""Model registry: the reference hard-codes ``FooModel()`` (``ddp.py:311``); ``--model`` selects here."""
from .foo import FooModel, BranchyFooModel
from .resnet import ResNet, resnet50, resnet152
from .bert import BertConfig, BertModel, BertForMaskedLM, bert_base

MODEL_REGISTRY = {
    "foo": FooModel,
    "resnet50": resnet50,
    "resnet152": resnet152,
    "bert-base": bert_base,
}


def build_model(name: str, **kwargs):
    try:
        return MODEL_REGISTRY[name](**kwargs)
    except KeyError:
        raise ValueError(f"unknown model {name!r}; choose from {sorted(MODEL_REGISTRY)}") from None


__all__ = ["FooModel", "BranchyFooModel", "ResNet", "resnet50", "resnet152", "BertConfig", "BertModel",
           "BertForMaskedLM", "bert_base", "MODEL_REGISTRY", "build_model"]
