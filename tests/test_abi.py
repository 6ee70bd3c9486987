"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol include/pasn.h declares,
argument validation works without a GPU, and the Python mirror keeps the reference's interface."""
import ctypes as C
import os
import re

import pytest
import torch

import protoasnet_b200 as pasn
from protoasnet_b200 import _lib, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "pasn.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pasn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _header_symbols()
    assert declared == sorted(_lib.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert lib.pasn_abi_version() == 2
    assert lib.pasn_strerror(0) == b"ok"
    assert b"workspace" in lib.pasn_strerror(-2)


def test_gemm_descriptor_mirror_matches_the_library():
    """The ctypes mirror of pasn::tcg::Gemm used by the GEMM unit tests and tools/probe_chain_gemms.py has the library's size
    (a field added on one side only would shift every later field silently)."""
    from tests.test_tc_gemm_gpu import Gemm
    assert _lib.load().pasn_debug_tc_gemm_desc_bytes() == C.sizeof(Gemm)


def test_argument_validation_without_gpu():
    lib = _lib.load()
    bad = _lib.PasnDims(1, 8, 7, 4, 2, 9, 0, 0, 0, 0)          # odd D
    assert lib.pasn_head_workspace_bytes(C.byref(bad)) == 0
    ok = _lib.PasnDims(4, 8, 6, 4, 2, 9, 0, 0, 0, 1)
    assert lib.pasn_head_workspace_bytes(C.byref(ok)) > 0
    w = _lib.PasnWeights()
    st = lib.pasn_head_forward(None, C.byref(w), None, C.byref(bad), None, None, None, None, None, None, None, 0, None)
    assert st == -1
    assert lib.pasn_push_init(None, 4, None) == -1
    # dtype / layout / path enums are range-checked
    for field, val in (("dtype", 5), ("layout", 3), ("path", 9), ("occ_act", 1)):
        d = _lib.PasnDims(4, 8, 6, 4, 2, 9, 0, 0, 0, 0)
        setattr(d, field, val)
        assert lib.pasn_head_workspace_bytes(C.byref(d)) == 0


def test_module_mirrors_reference_interface():
    dims = synth.CONFIGS["cfg1_video_yml"]
    m = pasn.build({"name": "Video_XProtoNet", "checkpoint_path": "", "base_architecture": pasn.FeatureInput(dims.C),
                    "pretrained": False, "prototype_shape": "(40, 256, 1, 1, 1)", "num_classes": 4,
                    "backbone_last_layer_num": -3})
    sd = m.state_dict()
    expect = {
        "prototype_vectors": (40, 256, 1, 1, 1), "ones": (40, 256, 1, 1, 1),
        "add_on_layers.0.weight": (256, 256, 1, 1, 1), "add_on_layers.0.bias": (256,),
        "add_on_layers.2.weight": (256, 256, 1, 1, 1), "add_on_layers.2.bias": (256,),
        "occurrence_module.0.weight": (256, 256, 1, 1, 1), "occurrence_module.0.bias": (256,),
        "occurrence_module.2.weight": (128, 256, 1, 1, 1), "occurrence_module.2.bias": (128,),
        "occurrence_module.4.weight": (40, 128, 1, 1, 1), "last_layer.weight": (4, 40),
    }
    assert {k: tuple(v.shape) for k, v in sd.items()} == expect
    assert all(v.dtype == torch.float32 for v in sd.values())
    # reference initialisation: zero biases, block-identity last layer, prototypes in [0,1)
    assert float(sd["add_on_layers.0.bias"].abs().max()) == 0.0
    assert torch.equal(sd["last_layer.weight"], m.prototype_class_identity.t())
    assert 0 <= float(sd["prototype_vectors"].min()) and float(sd["prototype_vectors"].max()) < 1
    for name in ("forward", "push_forward", "compute_occurence_map", "get_occurence_map_absolute_val",
                 "set_last_layer_incorrect_connection"):
        assert callable(getattr(m, name))
    for name in ("cnn_backbone", "add_on_layers", "occurrence_module", "last_layer", "prototype_class_identity",
                 "num_prototypes", "num_classes", "prototype_shape"):
        assert hasattr(m, name)
    # state_dict round trip with numpy-generated parameters
    m.load_state_dict({k: torch.from_numpy(v) for k, v in synth.make_head_params(dims).items()})


def test_image_model_and_registry():
    m = pasn.MODELS["XProtoNet"](pasn.FeatureInput(512), pretrained=False, prototype_shape=(40, 512, 1, 1), num_classes=4)
    assert tuple(m.state_dict()["occurrence_module.4.weight"].shape) == (40, 256, 1, 1)
    with pytest.raises(AssertionError):
        pasn.construct_Video_XProtoNet(pasn.FeatureInput(8), prototype_shape=(10, 8, 1, 1, 1), num_classes=4)  # P % K


def test_no_cpu_fallback_and_grad_guard():
    m = pasn.construct_Video_XProtoNet(pasn.FeatureInput(24), prototype_shape=(8, 16, 1, 1, 1), num_classes=4)
    x = torch.zeros(1, 24, 2, 3, 3)
    with torch.no_grad(), pytest.raises(_lib.PasnError):
        m(x)                                    # CPU tensor: refuses, never falls back
    with pytest.raises(NotImplementedError):
        m(x)                                    # grad enabled: explicit refusal unless autograd_mode='composite'
    m.autograd_mode = "composite"
    logits, sim, occ = m(x)
    assert logits.requires_grad and occ.shape == (1, 8, 1, 2, 3, 3)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "protoasnet_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


@pytest.mark.parametrize("name", ["video_resnet2p1d_18_last-3", "video_resnet2p1d_18_last-2", "image_resnet18"])
def test_checkpoint_layout_matches_reference_models(name):
    """state_dict keys and shapes of the models built from the reference's string backbones equal the reference's own
    (tests/golden/state_dict_keys.json, generated by oracle/gen_golden_keys.py from the reference constructors): reference
    checkpoints load with strict=True (src/agents/base.py:128, src/agents/XProtoNet_e2e.py:95)."""
    import json
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))[name]
    if name.startswith("video"):
        m = pasn.construct_Video_XProtoNet("resnet2p1d_18", pretrained=False, backbone_last_layer_num=int(name[-2:]))
    else:
        m = pasn.construct_XProtoNet("resnet18", pretrained=False, prototype_shape=(40, 512, 1, 1), num_classes=4)
    own = {k: list(v.shape) for k, v in m.state_dict().items()}
    assert own == ref
    m.load_state_dict({k: torch.zeros(v) for k, v in ref.items()}, strict=True)
