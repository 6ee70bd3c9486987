"""Writes tests/golden/state_dict_keys.json: the state_dict key -> shape tables of the reference models built from their
string backbones (``pretrained=False``), so that checkpoint-layout parity (strict ``load_state_dict``, src/agents/base.py:128,
src/agents/XProtoNet_e2e.py:95) can be tested where /root/reference does not exist.  TEST INFRASTRUCTURE (see oracle/__init__.py).

    PYTHONPATH=/root/reference python oracle/gen_golden_keys.py
"""
import json
import os
import sys

sys.path.insert(0, "/root/reference")
from src.models.Video_XProtoNet import construct_Video_XProtoNet  # noqa: E402
from src.models.XProtoNet import construct_XProtoNet  # noqa: E402

out = {}
for name, m in (
    ("video_resnet2p1d_18_last-3", construct_Video_XProtoNet("resnet2p1d_18", pretrained=False, backbone_last_layer_num=-3)),
    ("video_resnet2p1d_18_last-2", construct_Video_XProtoNet("resnet2p1d_18", pretrained=False, backbone_last_layer_num=-2)),
    ("image_resnet18", construct_XProtoNet("resnet18", pretrained=False, prototype_shape=(40, 512, 1, 1), num_classes=4)),
):
    out[name] = {k: list(v.shape) for k, v in m.state_dict().items()}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "state_dict_keys.json")
json.dump(out, open(path, "w"), indent=0, sort_keys=True)
print({k: len(v) for k, v in out.items()})
