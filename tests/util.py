"""Shared helpers for the GPU parity tests (they call the CUDA path through the C ABI via protoasnet_b200)."""
import json

import numpy as np
import torch

import protoasnet_b200 as pasn
from protoasnet_b200 import synth

FP32_RTOL = 1e-5   # north_star: logits and similarities within 1e-5 relative in fp32
BF16_RTOL = 1e-3   # ... and 1e-3 relative in bf16
# fp32 feature maps on the tensor cores (tiled path: bf16 hi/lo split, 16 significant bits per operand, fp32 accumulation):
# logits / similarities / distances keep north_star's 1e-5; the big intermediate tensors that are also outputs
# (features_extracted, occurrence_map, pushed prototype_vectors) are judged against their scale at this fraction.
# kernel_path = PASN_PATH_GENERIC (CUDA-core FFMA) is the path that meets 1e-6 of scale on those as well.
FP32_TC_FEAT_ATOL = 3e-5


def feat_atol(model, x):
    """atol_frac for features_extracted / occurrence_map of ``model`` fed ``x``: None (default, rtol/10) unless an fp32 input
    is served by a tensor-core path."""
    import ctypes as C
    from protoasnet_b200 import _lib
    if x.dtype != torch.float32:
        return None
    dims = model._rt.make_dims(x, model.kernel_path)[0]
    return FP32_TC_FEAT_ATOL if _lib.load().pasn_tcgen05_supported(C.byref(dims)) else None


def build_model(dims: synth.HeadDims, sd_np, device="cuda", path=None):
    if dims.ndim == 3:
        m = pasn.construct_Video_XProtoNet(pasn.FeatureInput(dims.C), pretrained=False,
                                           prototype_shape=dims.prototype_shape, num_classes=dims.K)
    else:
        m = pasn.construct_XProtoNet(pasn.FeatureInput(dims.C), pretrained=False,
                                     prototype_shape=dims.prototype_shape, num_classes=dims.K)
    own = m.state_dict()
    assert set(own.keys()) == set(sd_np.keys())
    m.load_state_dict({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in sd_np.items()})
    m = m.to(device).eval()
    if path is not None:
        m.kernel_path = path
    return m


def assert_close(got, ref, rtol, name="", atol_frac=None):
    """|got - ref| <= atol + rtol*|ref| with atol = atol_frac * max|ref| (default rtol/10: near-zero entries of a
    tensor are judged against the tensor's scale, everything else relatively)."""
    got = got.detach().float().cpu().numpy() if isinstance(got, torch.Tensor) else np.asarray(got)
    ref = np.asarray(ref, dtype=np.float64)
    scale = float(np.abs(ref).max()) if ref.size else 1.0
    frac = rtol * 0.1 if atol_frac is None else atol_frac
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=frac * scale + 1e-30, err_msg=name)


def load_golden(path):
    z = np.load(path)
    return z, json.loads(str(z["recipe"]))
