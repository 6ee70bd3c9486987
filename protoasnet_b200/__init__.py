"""protoasnet_b200 -- B200-native prototype head + push for ProtoASNet.

Public API mirrors the reference's model/push interface:
    from protoasnet_b200 import Video_XProtoNet, XProtoNet, construct_Video_XProtoNet, construct_XProtoNet, MODELS, build
    from protoasnet_b200 import push_prototypes, load_data_and_model_products
"""
from .head import (MODELS, FeatureInput, HostPipeline, Video_XProtoNet, XProtoNet, build, construct_Video_XProtoNet,  # noqa: F401
                   construct_XProtoNet)
from .push import push_prototypes, push_resident  # noqa: F401
from .explain import collect_model_products, load_data_and_model_products  # noqa: F401
from . import metrics  # noqa: F401

__all__ = ["MODELS", "FeatureInput", "HostPipeline", "Video_XProtoNet", "XProtoNet", "build", "construct_Video_XProtoNet",
           "construct_XProtoNet", "push_prototypes", "push_resident", "collect_model_products",
           "load_data_and_model_products", "metrics"]
